"""coperception.models.seg.UNet on the sm_100a path (reference: CP/models/seg/UNet.py:5-44)."""
import torch

from .SegModelBase import SegModelBase


class UNet(SegModelBase):
    def __init__(self, n_channels, n_classes, bilinear=True, num_agent=5, kd_flag=False, compress_level=0):
        super().__init__(n_channels, n_classes, bilinear, num_agent=num_agent, compress_level=compress_level)
        if kd_flag:
            raise NotImplementedError("kd_flag outputs are not exported by the sm_100a seg path yet")
        self.kd_flag = kd_flag

    def forward(self, x):
        """x [N,13,256,256] fp32 -> logits [N,n_classes,256,256]."""
        from v2x_b200 import nets_seg
        self._check(x)
        n = int(x.shape[0])
        plan = self._get_plan(("unet", n, x.device.index, self.precision),
                              lambda: nets_seg.SegUNetPlan(self._state(), n, planes=self._planes(), device=x.device))
        return plan.forward(x.to(torch.float32).contiguous())
