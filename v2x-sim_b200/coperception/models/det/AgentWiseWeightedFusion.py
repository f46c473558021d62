"""coperception.models.det.AgentWiseWeightedFusion on the sm_100a path
(reference: CP/models/det/AgentWiseWeightedFusion.py:7-76)."""
from ._fusion import FusionBase, PairWeightNet


class AgentWiseWeightedFusion(FusionBase):
    """One learned scalar weight per (target, list member) pair, softmax over the list, weighted sum."""
    KIND = "agent"

    def __init__(self, config, layer=3, in_channels=13, kd_flag=True, num_agent=5, compress_level=0, only_v2i=False):
        super().__init__(config, layer, in_channels, kd_flag, num_agent, compress_level, only_v2i)
        self.agent_weighted_fusion = PairWeightNet(256, agent_wise=True)
