"""coperception.models.det.TeacherNet on the sm_100a path (reference: CP/models/det/TeacherNet.py:4-13)."""
import torch

from ._base import B200DetModel
from ._schema import BackboneParams


class TeacherNet(B200DetModel):
    """The early-fusion teacher of DiscoNet's distillation: a bare STPN whose forward returns the decoder layers and
    the two deepest encoder layers, (x_8, x_7, x_6, x_5, x_3, x_4) (STPN_KD.forward, Backbone.py:251-257)."""

    def __init__(self, config):
        super().__init__(config)
        self.stpn = BackboneParams(config.map_dims[2], 0)

    def forward(self, bevs, maps=None, vis=None):
        from v2x_b200 import nets, ops
        self._check_eval()
        dev = bevs.device
        if dev.type != "cuda":
            raise RuntimeError("v2x_b200 TeacherNet needs CUDA tensors (no CPU fallback); got %s" % dev)
        n = int(bevs.shape[0])
        plan = self._get_plan(("teacher", n, dev.index, self.precision),
                              lambda: nets.FaFNetPlan(self._state(), n, planes=self._planes(), device=dev, heads=False))
        plan.forward(bevs.to(torch.float32))
        f = {k: ops.act_to_float(plan.ws[k]) for k in ("x8", "x7u", "x6u", "x5u", "x3", "x4u")}
        half = lambda t: t[:, :, ::2, ::2].contiguous()  # noqa: E731  (stored nearest-upsampled: one value per 2x2 block)
        return f["x8"], half(f["x7u"]), half(f["x6u"]), half(f["x5u"]), f["x3"], half(f["x4u"])
