"""Shared host logic of the intermediate-fusion baselines on the sm_100a path (reference:
CP/models/det/base/FusionBase.py:4-75, IntermediateModelBase.py:5-25): encoder -> cross-agent fuse at layer 3 ->
decoder -> heads; the fuse rule is the only thing the subclasses differ in (``KIND``)."""
import torch
import torch.nn as nn

from ._base import B200DetModel
from ._schema import BackboneParams


class PairWeightNet(nn.Module):
    """Parameters of PixelWeightedFusionSoftmax (DiscoNet.py:132-147) / AgentWeightedFusion
    (AgentWiseWeightedFusion.py:44-64): 1x1 convs 2C -> 128 -> 32 -> 8 -> 1 with BN on the first three; the agent-wise
    variant adds the 32x32 ``conv1_5``."""

    def __init__(self, channel, agent_wise=False):
        super().__init__()
        self.conv1_1 = nn.Conv2d(channel * 2, 128, kernel_size=1, stride=1, padding=0)
        self.bn1_1 = nn.BatchNorm2d(128)
        self.conv1_2 = nn.Conv2d(128, 32, kernel_size=1, stride=1, padding=0)
        self.bn1_2 = nn.BatchNorm2d(32)
        self.conv1_3 = nn.Conv2d(32, 8, kernel_size=1, stride=1, padding=0)
        self.bn1_3 = nn.BatchNorm2d(8)
        self.conv1_4 = nn.Conv2d(8, 1, kernel_size=1, stride=1, padding=0)
        if agent_wise:
            self.conv1_5 = nn.Conv2d(1, 1, kernel_size=32, stride=1, padding=0)


class FusionBase(B200DetModel):
    KIND = None

    def __init__(self, config, layer=3, in_channels=13, kd_flag=True, num_agent=5, compress_level=0, only_v2i=False):
        super().__init__(config, layer, in_channels, kd_flag, num_agent=num_agent, only_v2i=only_v2i)
        if layer not in (0, 1, 2, 3, 4):
            raise NotImplementedError("fusion models fuse at layer 0..4 (DetModelBase.py:71-92)")
        self.compress_level = compress_level
        self.u_encoder = BackboneParams(in_channels, compress_level)
        self.decoder = BackboneParams(in_channels)
        self.num_agent = 0   # the reference overwrites this per scene (FusionBase.py:16,41)

    def _run(self, bevs, trans_matrices, num_agent_tensor, batch_size):
        from v2x_b200 import nets
        if self.KIND is None:
            raise NotImplementedError("Please implement this method for specific fusion strategies")
        dev = bevs.device
        if self.training and self.KIND in ("mean", "sum", "max", "cat", "agent", "disco"):
            # train-mode forward with a backward pass behind torch.autograd (v2x_b200/train.py::FusionTrainStep), every
            # fuse rule, with or without the kd_flag outputs
            if self.layer != 3 or self.compress_level > 3:
                raise NotImplementedError("training on the sm_100a path: layer 3, compress_level 0..3")
            if dev.type != "cuda":
                raise RuntimeError("v2x_b200 fusion models need CUDA tensors (no CPU fallback); got %s" % dev)
            from v2x_b200.train import FusionTrainStep
            outs = FusionTrainStep.apply(self, self.KIND, bevs, trans_matrices, num_agent_tensor, int(batch_size),
                                         *self.parameters())
            self._train_kd = tuple(outs[2:])       # (x_8, x_7, x_6, x_5, fused) when kd_flag == 1, else ()
            return None, {"loc": outs[0], "cls": outs[1]}
        self._check_eval()
        if dev.type != "cuda":
            raise RuntimeError("v2x_b200 fusion models need CUDA tensors (no CPU fallback); got %s" % dev)
        assert bevs.shape[0] == batch_size * self.agent_num, "bevs must hold batch_size * num_agent maps"
        key = (self.KIND, int(batch_size), dev.index, self.precision)
        plan = self._get_plan(key, lambda: nets.FusionDetPlan(
            self._state(), self.KIND, int(batch_size), self.agent_num, planes=self._planes(), device=dev,
            only_v2i=self.only_v2i, layer=self.layer))
        result = plan.forward(bevs.to(torch.float32), trans_matrices.to(torch.float64), num_agent_tensor.to(torch.int64))
        return plan, result

    def forward(self, bevs, trans_matrices, num_agent_tensor, batch_size=1):
        """Same contract as FusionBase.forward (FusionBase.py:23-75): the result dict, or with ``kd_flag == 1`` the
        tuple (result, x_8, x_7, x_6, x_5, fused layer-3 maps)."""
        plan, result = self._run(bevs, trans_matrices, num_agent_tensor, batch_size)
        if plan is None:       # model.train(): the train step
            return (result, *self._train_kd) if self.kd_flag == 1 else result
        if self.kd_flag == 1:
            return (result, *plan.kd_layers())
        return result
