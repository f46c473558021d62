"""Parameter containers with the reference's module / parameter names, so checkpoints load with
``strict=True`` (SURVEY.md section 8(b) "state_dict").  These modules hold nn.Parameters and BN
buffers only; the arithmetic runs in libv2x_b200.so on operands packed from them.

Name / shape schema follows CP/models/det/backbone/Backbone.py:9-87 (Backbone, Conv3D),
CP/models/det/base/DetModelBase.py:268-351 (heads) and
CP/utils/convolutional_rnn/module.py:62-124 (Conv2dGRU parameters and their U(-1/sqrt(C), 1/sqrt(C)) init).
"""
import math

import torch
import torch.nn as nn


class Conv3D(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv3d = nn.Conv3d(cin, cout, kernel_size=(1, 1, 1), stride=1, padding=(0, 0, 0))
        self.bn3d = nn.BatchNorm3d(cout)


class BackboneParams(nn.Module):
    """All parameters of one reference ``Backbone`` (encoder AND decoder halves, as LidarEncoder /
    LidarDecoder / STPN_KD each carry the full set -- SURVEY.md Q13)."""

    _CONVS = [("conv1_1", 32, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128),
              ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv4_1", 256, 512), ("conv4_2", 512, 512),
              ("conv5_1", 768, 256), ("conv5_2", 256, 256), ("conv6_1", 384, 128), ("conv6_2", 128, 128),
              ("conv7_1", 192, 64), ("conv7_2", 64, 64), ("conv8_1", 96, 32), ("conv8_2", 32, 32)]

    def __init__(self, height_feat_size=13, compress_level=0):
        super().__init__()
        self.conv_pre_1 = nn.Conv2d(height_feat_size, 32, 3, 1, 1)
        self.conv_pre_2 = nn.Conv2d(32, 32, 3, 1, 1)
        self.bn_pre_1 = nn.BatchNorm2d(32)
        self.bn_pre_2 = nn.BatchNorm2d(32)
        self.conv3d_1 = Conv3D(64, 64)
        self.conv3d_2 = Conv3D(128, 128)
        for name, cin, cout in self._CONVS:
            setattr(self, name, nn.Conv2d(cin, cout, 3, 2 if name.endswith("_1") and name[4] in "1234" else 1, 1))
        for name, _, cout in self._CONVS:
            setattr(self, "bn" + name[4:], nn.BatchNorm2d(cout))
        self.compress_level = compress_level
        if compress_level > 0:
            assert compress_level <= 8
            cc = 256 // (2 ** compress_level)
            self.com_compresser = nn.Conv2d(256, cc, 1, 1)
            self.bn_compress = nn.BatchNorm2d(cc)
            self.com_decompresser = nn.Conv2d(cc, 256, 1, 1)
            self.bn_decompress = nn.BatchNorm2d(256)


class ClassificationHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        channel = 32
        anchors = len(config.anchor_size)
        self.conv1 = nn.Conv2d(channel, channel, 3, 1, 1)
        self.conv2 = nn.Conv2d(channel, config.category_num * anchors, 1, 1, 0)
        self.bn1 = nn.BatchNorm2d(channel)


class SingleRegressionHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        channel = 32
        anchors = len(config.anchor_size)
        out_seq_len = 1 if config.only_det else config.pred_len
        self.box_prediction = nn.Sequential(
            nn.Conv2d(channel, channel, 3, 1, 1), nn.BatchNorm2d(channel), nn.ReLU(),
            nn.Conv2d(channel, anchors * config.box_code_size * out_seq_len, 1, 1, 0))


class Conv2dGRUParams(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3):
        super().__init__()
        g = 3 * out_channels
        self.weight_ih_l0 = nn.Parameter(torch.empty(g, in_channels, kernel_size, kernel_size))
        self.weight_hh_l0 = nn.Parameter(torch.empty(g, out_channels, kernel_size, kernel_size))
        self.bias_ih_l0 = nn.Parameter(torch.empty(g))
        self.bias_hh_l0 = nn.Parameter(torch.empty(g))
        stdv = 1.0 / math.sqrt(out_channels)
        for p in self.parameters():
            p.data.uniform_(-stdv, stdv)


def check_config(config):
    """The sm_100a heads implement the reference's default detection configuration only."""
    unsupported = []
    if getattr(config, "use_map", False):
        unsupported.append("use_map")
    if getattr(config, "use_vis", False):
        unsupported.append("use_vis")
    if getattr(config, "motion_state", False):
        unsupported.append("motion_state")
    if not getattr(config, "binary", True) or not getattr(config, "only_det", True):
        unsupported.append("binary/only_det != True")
    if unsupported:
        raise NotImplementedError("v2x_b200 detection heads do not support: " + ", ".join(unsupported))
