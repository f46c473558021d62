"""coperception.models.seg.AgentWiseWeightedFusion on the sm_100a path
(reference: CP/models/seg/AgentWiseWeightedFusion.py:8-79)."""
from ..det._fusion import PairWeightNet
from .FusionBase import FusionBase


class AgentWiseWeightedFusion(FusionBase):
    KIND = "agent"

    def __init__(self, n_channels, n_classes, num_agent=5, compress_level=0, only_v2i=False):
        super().__init__(n_channels, n_classes, num_agent=num_agent, compress_level=compress_level, only_v2i=only_v2i)
        self.agent_weighted_fusion = PairWeightNet(512, agent_wise=True)
