"""coperception.models.det.FaFNet on the sm_100a path (reference: CP/models/det/FaFNet.py:4-39)."""
import torch

from ._base import B200DetModel
from ._schema import BackboneParams


class FaFNet(B200DetModel):
    """No-fusion STPN detector: lower-bound (own BEV) / upper-bound (early-fused BEV) depending on the input.
    Parameters live under ``stpn.*`` (NonIntermediateModelBase.py:24)."""

    def __init__(self, config, layer=3, in_channels=13, kd_flag=True, num_agent=5, compress_level=0):
        super().__init__(config, layer, in_channels, kd_flag, num_agent=num_agent)
        self.stpn = BackboneParams(config.map_dims[2], compress_level)

    def forward(self, bevs, maps=None, vis=None, batch_size=None):
        """Called with the fusion-model argument list by FaFModule (CoDetModule.py:254-256): ``maps`` /
        ``vis`` receive trans_matrices / num_agent_tensor and are ignored, as in the reference (Q12)."""
        from v2x_b200 import nets, ops
        dev = bevs.device
        if self.training:
            # train-mode forward (BatchNorm batch statistics + running-buffer update) with a backward pass behind
            # torch.autograd: FaFModule.step's loss.backward() / optimizer.step() drive it (CoDetModule.py:283-291)
            if self.kd_flag == 1 or (hasattr(self.stpn, "com_compresser") and self.stpn.com_compresser.out_channels < 32):
                raise NotImplementedError("training with kd_flag == 1 / compress_level > 3 is not built on the sm_100a path")
            if dev.type != "cuda":
                raise RuntimeError("v2x_b200 FaFNet needs CUDA tensors (no CPU fallback); got %s" % dev)
            from v2x_b200.train import FaFNetTrainStep
            loc, cls = FaFNetTrainStep.apply(self, bevs, *self.parameters())
            return {"loc": loc, "cls": cls}
        self._check_eval()
        if dev.type != "cuda":
            raise RuntimeError("v2x_b200 FaFNet needs CUDA tensors (no CPU fallback); got %s" % dev)
        n = int(bevs.shape[0])
        plan = self._get_plan(("faf", n, dev.index, self.precision),
                              lambda: nets.FaFNetPlan(self._state(), n, planes=self._planes(), device=dev))
        result = plan.forward(bevs.to(torch.float32))
        if self.kd_flag == 1:
            # (result, x_8, x_7, x_6, x_5, x_3) as FaFNet.py:36-37; x_7/x_6/x_5 are stored 2x-upsampled
            # x_3 is encoded_layers[3], i.e. AFTER com_compresser / com_decompresser when compress_level > 0
            # (Backbone.py:138-141): the workspace holds that map as "x3d"
            x3 = "x3d" if "x3d" in plan.ws else "x3"
            f = {k: ops.act_to_float(plan.ws[k]) for k in ("x8", "x7u", "x6u", "x5u", x3)}
            return result, f["x8"], f["x7u"][:, :, ::2, ::2], f["x6u"][:, :, ::2, ::2], f["x5u"][:, :, ::2, ::2], f[x3]
        return result
