"""Parameter schema + host logic of the segmentation models (reference: CP/models/seg/SegModelBase.py:6-151).

The nn.Modules below only hold parameters / BN buffers under the reference's names (inc.double_conv.{0,1,3,4},
down*.maxpool_conv.1.double_conv.*, up*.conv.double_conv.*, outc.conv) so reference checkpoints load strictly; the
arithmetic runs in libv2x_b200.so."""
import os

import torch
import torch.nn as nn

from ..det._base import PlanCacheMixin


class DoubleConv(nn.Module):
    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid_channels = mid_channels or out_channels
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1), nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True),
            nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))


class Down(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))


class Up(nn.Module):
    def __init__(self, in_channels, out_channels, bilinear=True):
        super().__init__()
        self.conv = DoubleConv(in_channels, out_channels, in_channels // 2)


class OutConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)


class SegModelBase(PlanCacheMixin, nn.Module):
    def __init__(self, n_channels, n_classes, bilinear=True, num_agent=5, compress_level=0, only_v2i=False):
        super().__init__()
        if not bilinear:
            raise NotImplementedError("v2x_b200 seg models implement the reference default (bilinear upsampling)")
        self.n_channels, self.n_classes, self.bilinear = n_channels, n_classes, bilinear
        self.num_agent, self.only_v2i, self.compress_level = num_agent, only_v2i, compress_level
        self.inc = DoubleConv(n_channels, 64)
        self.down1, self.down2, self.down3, self.down4 = Down(64, 128), Down(128, 256), Down(256, 512), Down(512, 512)
        self.up1, self.up2, self.up3, self.up4 = Up(1024, 256), Up(512, 128), Up(256, 64), Up(128, 64)
        self.outc = OutConv(64, n_classes)
        if compress_level > 0:   # SegModelBase.py:29-43
            assert compress_level <= 9
            cc = 512 // (2 ** compress_level)
            self.com_compresser = nn.Conv2d(512, cc, kernel_size=1, stride=1)
            self.bn_compress = nn.BatchNorm2d(cc)
            self.com_decompresser = nn.Conv2d(cc, 512, kernel_size=1, stride=1)
            self.bn_decompress = nn.BatchNorm2d(512)
        self._init_plan_cache()

    def _check(self, x):
        if self.training:
            raise NotImplementedError("the sm_100a path trains every seg model except DiscoNet; this model only implements "
                                      "inference (model.eval())")
        if x.device.type != "cuda":
            raise RuntimeError("v2x_b200 seg models need CUDA tensors (no CPU fallback); got %s" % x.device)
        self._warn_no_grad_graph()

    def _train_forward(self, x, fuse=None):
        """Train-mode forward with a backward pass behind torch.autograd (v2x_b200/train.py::SegTrainStep)."""
        if self.compress_level > 4 or getattr(self, "kd_flag", False):
            raise NotImplementedError("training with compress_level > 4 (fewer than 32 compressed channels) / kd_flag is "
                                      "not built on the sm_100a path")
        if x.device.type != "cuda":
            raise RuntimeError("v2x_b200 seg models need CUDA tensors (no CPU fallback); got %s" % x.device)
        from v2x_b200.train import SegTrainStep
        return SegTrainStep.apply(self, fuse, x, *self.parameters())
