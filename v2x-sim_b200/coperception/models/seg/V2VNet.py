"""coperception.models.seg.V2VNet on the sm_100a path (reference: CP/models/seg/V2VNet.py:8-92)."""
import torch

from ..det._schema import Conv2dGRUParams
from .SegModelBase import SegModelBase


class V2VNet(SegModelBase):
    def __init__(self, n_channels, n_classes, num_agent=5, compress_level=0, only_v2i=False):
        super().__init__(n_channels, n_classes, num_agent=num_agent, compress_level=compress_level, only_v2i=only_v2i)
        self.layer_channel = 512
        self.gnn_iter_num = 1
        self.convgru = Conv2dGRUParams(self.layer_channel * 2, self.layer_channel, 3)

    def forward(self, x, trans_matrices, num_agent_tensor):
        from v2x_b200 import nets_seg
        batch = int(x.shape[0]) // self.num_agent
        if self.training:
            return self._train_forward(x, (trans_matrices, num_agent_tensor, batch, self.num_agent, bool(self.only_v2i)))
        self._check(x)
        plan = self._get_plan(("v2v", batch, x.device.index, self.precision),
                              lambda: nets_seg.SegV2VNetPlan(self._state(), batch, self.num_agent, planes=self._planes(),
                                                             device=x.device, only_v2i=self.only_v2i))
        return plan.forward(x.to(torch.float32).contiguous(), trans_matrices.to(torch.float64),
                            num_agent_tensor.to(torch.int64))
