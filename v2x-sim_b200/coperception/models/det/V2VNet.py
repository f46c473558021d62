"""coperception.models.det.V2VNet on the sm_100a path (reference: CP/models/det/V2VNet.py:8-120)."""
import torch

from ._base import B200DetModel
from ._schema import BackboneParams, Conv2dGRUParams


class V2VNet(B200DetModel):
    """V2VNet (https://arxiv.org/abs/2008.07519): per-agent BEV encoder, ``gnn_iter_times`` rounds of
    warp -> neighbour mean -> ConvGRU at layer 3, decoder, detection heads.

    Constructor and forward signatures, parameter names and the returned dict match the reference
    (V2VNet.py:14-24, :47, :120); u_encoder / decoder each carry a full Backbone parameter set
    (IntermediateModelBase.py:24-25)."""

    def __init__(self, config, gnn_iter_times, layer, layer_channel, in_channels=13, num_agent=5, compress_level=0,
                 only_v2i=False):
        super().__init__(config, layer, in_channels, num_agent=num_agent, only_v2i=only_v2i)
        if layer not in (1, 2, 3, 4) or layer_channel != (32, 64, 128, 256, 512)[layer]:
            raise NotImplementedError("v2x_b200 V2VNet communicates at layer 1..4 with layer_channel = 64 / 128 / 256 / 512 "
                                      "(the reference scripts use layer 3, train_codet.py:106-114)")
        self.u_encoder = BackboneParams(in_channels, compress_level)
        self.decoder = BackboneParams(in_channels)
        self.layer_channel = layer_channel
        self.gnn_iter_num = gnn_iter_times
        self.convgru = Conv2dGRUParams(layer_channel * 2, layer_channel, 3)
        self.compress_level = compress_level

    def forward(self, bevs, trans_matrices, num_agent_tensor, batch_size=1):
        """bevs [A*B,1,256,256,13] (agent-major), trans_matrices [B,A,A,4,4], num_agent_tensor [B,A]
        -> {"loc": [A*B,256,256,6,1,6], "cls": [A*B,393216,2]} (fp32, on bevs.device)."""
        from v2x_b200 import nets
        dev = bevs.device
        if dev.type != "cuda":
            raise RuntimeError("v2x_b200 V2VNet needs CUDA tensors (no CPU fallback); got %s" % dev)
        assert bevs.shape[0] == batch_size * self.agent_num, "bevs must hold batch_size * num_agent maps"
        if self.training:
            # train-mode forward (BatchNorm batch statistics) with the backward pass behind torch.autograd, so the
            # reference's FaFModule.step (loss.backward() / optimizer.step(), CoDetModule.py:283-291) drives it
            if self.layer != 3 or self.compress_level > 3:
                raise NotImplementedError("training on the sm_100a path: layer 3, compress_level 0..3")
            from v2x_b200.train import V2VNetTrainStep
            loc, cls = V2VNetTrainStep.apply(self, bevs, trans_matrices, num_agent_tensor, int(batch_size), *self.parameters())
            return {"loc": loc, "cls": cls}
        self._check_eval()
        # bool / uint8 occupancy grids (the dataset's array before .astype(np.float32)) are expanded on device
        mode = "u8" if bevs.dtype in (torch.uint8, torch.bool) else "f32"
        key = ("v2v", int(batch_size), dev.index, self.precision, mode)
        plan = self._get_plan(key, lambda: nets.V2VNetDetPlan(
            self._state(), int(batch_size), self.agent_num, gnn_iter=self.gnn_iter_num, planes=self._planes(),
            device=dev, only_v2i=self.only_v2i, input_mode=mode, layer=self.layer))
        if mode == "u8":
            bevs = bevs.view(torch.uint8) if bevs.dtype == torch.bool else bevs
        else:
            bevs = bevs.to(torch.float32)
        return plan.forward(bevs, trans_matrices.to(torch.float64), num_agent_tensor.to(torch.int64))

    def forward_voxels(self, voxel_rows, trans_matrices, num_agent_tensor, batch_size=1, capacity=0):
        """Extension of the reference surface: ``voxel_rows`` int32 [n, 4] = (map, i0, i1, i2), the
        dataset's sparse ``voxel_indices_0`` of every map prefixed by its agent-major map index; the dense BEV
        (scatter + np.rot90(., 3), V2XSimDet.py:294-299) is built on device.  Same result dict as forward()."""
        from v2x_b200 import nets
        self._check_eval()
        dev = trans_matrices.device
        if dev.type != "cuda":
            raise RuntimeError("v2x_b200 V2VNet needs CUDA tensors (no CPU fallback); got %s" % dev)
        cap = int(capacity) or max(32768 * batch_size * self.agent_num, int(voxel_rows.shape[0]))
        key = ("v2v", int(batch_size), dev.index, self.precision, "voxels", cap)
        plan = self._get_plan(key, lambda: nets.V2VNetDetPlan(
            self._state(), int(batch_size), self.agent_num, gnn_iter=self.gnn_iter_num, planes=self._planes(),
            device=dev, only_v2i=self.only_v2i, input_mode="voxels", voxel_capacity=cap, layer=self.layer))
        out = plan.forward(voxel_rows, trans_matrices.to(torch.float64), num_agent_tensor.to(torch.int64))
        plan.check_voxels()
        return out
