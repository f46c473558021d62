"""coperception.models.seg.FusionBase on the sm_100a path (reference: CP/models/seg/FusionBase.py:6-84): UNet encoder ->
cross-agent fuse of the layer-4 maps -> UNet decoder; subclasses pick the fuse rule (``KIND``)."""
import torch

from .SegModelBase import SegModelBase


class FusionBase(SegModelBase):
    KIND = None

    def __init__(self, n_channels, n_classes, num_agent=5, kd_flag=False, compress_level=0, only_v2i=False):
        super().__init__(n_channels, n_classes, num_agent=num_agent, compress_level=compress_level, only_v2i=only_v2i)
        self.neighbor_feat_list = None
        self.tg_agent = None
        self.current_num_agent = None
        self.kd_flag = kd_flag
        self.only_v2i = only_v2i

    def fusion(self):
        raise NotImplementedError("Please implement this method for specific fusion strategies")

    def forward(self, x, trans_matrices, num_agent_tensor):
        """x [A*B,13,256,256] -> logits [A*B,n_classes,256,256]; with kd_flag also (x9, x8, x7, x6, x5, feat_mat)."""
        from v2x_b200 import nets_seg
        if self.KIND is None:
            self.fusion()
        if self.training and self.KIND in ("mean", "sum", "max", "cat", "agent"):
            # train-mode forward with a backward pass behind torch.autograd (v2x_b200/train.py::SegTrainStep)
            batch = int(x.shape[0]) // self.num_agent
            return self._train_forward(x, (trans_matrices, num_agent_tensor, batch, self.num_agent, bool(self.only_v2i),
                                           self.KIND))
        self._check(x)
        batch = int(x.shape[0]) // self.num_agent
        plan = self._get_plan((self.KIND, batch, x.device.index, self.precision),
                              lambda: nets_seg.SegFusionPlan(self._state(), self.KIND, batch, self.num_agent,
                                                             planes=self._planes(), device=x.device,
                                                             only_v2i=self.only_v2i))
        logits = plan.forward(x.to(torch.float32).contiguous(), trans_matrices.to(torch.float64),
                              num_agent_tensor.to(torch.int64))
        if self.kd_flag:
            return (logits, *plan.kd_layers())
        return logits
