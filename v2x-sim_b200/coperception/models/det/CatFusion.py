"""coperception.models.det.CatFusion on the sm_100a path (reference: CP/models/det/CatFusion.py:7-41)."""
import torch.nn as nn

from ._fusion import FusionBase


class ModulationLayer3(nn.Module):
    def __init__(self):
        super().__init__()
        self._conv1_1 = nn.Conv2d(512, 256, kernel_size=1, stride=1, padding=0)
        self._bn1_1 = nn.BatchNorm2d(256)


class CatFusion(FusionBase):
    """cat([target, mean over the list]) -> 1x1 conv + BN + ReLU (CatFusion.py:22-26): a two-source tensor-core conv."""
    KIND = "cat"

    def __init__(self, config, layer=3, in_channels=13, kd_flag=True, num_agent=5, compress_level=0, only_v2i=False):
        super().__init__(config, layer, in_channels, kd_flag, num_agent, compress_level, only_v2i)
        self._modulation_layer_3 = ModulationLayer3()
