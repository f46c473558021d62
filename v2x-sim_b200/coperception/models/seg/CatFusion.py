"""coperception.models.seg.CatFusion on the sm_100a path (reference: CP/models/seg/CatFusion.py:8-34)."""
import torch.nn as nn

from .FusionBase import FusionBase


class ModulationLayer3(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1_1 = nn.Conv2d(1024, 512, kernel_size=1, stride=1, padding=0)
        self.bn1_1 = nn.BatchNorm2d(512)


class CatFusion(FusionBase):
    KIND = "cat"

    def __init__(self, n_channels, n_classes, num_agent, compress_level, only_v2i):
        super().__init__(n_channels, n_classes, num_agent=num_agent, compress_level=compress_level, only_v2i=only_v2i)
        self.modulation_layer_3 = ModulationLayer3()
