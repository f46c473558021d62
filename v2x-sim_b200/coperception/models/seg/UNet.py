"""coperception.models.seg.UNet on the sm_100a path (reference: CP/models/seg/UNet.py:5-44)."""
import torch

from .SegModelBase import SegModelBase


class UNet(SegModelBase):
    def __init__(self, n_channels, n_classes, bilinear=True, num_agent=5, kd_flag=False, compress_level=0):
        super().__init__(n_channels, n_classes, bilinear, num_agent=num_agent, compress_level=compress_level)
        self.kd_flag = kd_flag

    def forward(self, x):
        """x [N,13,256,256] fp32 -> logits [N,n_classes,256,256]."""
        from v2x_b200 import nets_seg
        if self.training:
            return self._train_forward(x)
        self._check(x)
        n = int(x.shape[0])
        plan = self._get_plan(("unet", n, x.device.index, self.precision),
                              lambda: nets_seg.SegUNetPlan(self._state(), n, planes=self._planes(), device=x.device))
        logits = plan.forward(x.to(torch.float32).contiguous())
        if self.kd_flag:   # (logits, x9, x8, x7, x6, x5, x4) as UNet.py:41-42
            from v2x_b200 import ops
            return (logits, *[ops.act_to_float(plan.ws[k]) for k in ("x9", "x8", "x7", "x6", "x5")],
                    ops.act_to_float(plan.x4))
        return logits
