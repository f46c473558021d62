"""coperception.models.seg.DiscoNet on the sm_100a path (reference: CP/models/seg/DiscoNet.py:8-126)."""
from ..det._fusion import PairWeightNet
from .FusionBase import FusionBase


class DiscoNet(FusionBase):
    KIND = "disco"

    def __init__(self, n_channels, n_classes, num_agent, kd_flag=True, compress_level=0, only_v2i=False):
        super().__init__(n_channels, n_classes, num_agent, kd_flag=kd_flag, compress_level=compress_level,
                         only_v2i=only_v2i)
        self.pixel_weighted_fusion = PairWeightNet(512)
