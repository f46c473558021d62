"""Shared host logic of the drop-in detection modules: plan cache, precision mode, input staging."""
import os

import torch
import torch.nn as nn

from ._schema import ClassificationHead, SingleRegressionHead, check_config

_PLANES = {"bf16": 1, "bf16x3": 2}


class B200DetModel(nn.Module):
    """nn.Module surface of DetModelBase (CP/models/det/base/DetModelBase.py:26-51) whose forward runs on
    libv2x_b200.so.  ``precision``: "bf16" (default; activations/weights stored in bf16, fp32 accumulate)
    or "bf16x3" (hi/lo split operands, 3 MMAs, ~fp32-grade parity) -- also settable through the
    V2X_PRECISION environment variable."""

    def __init__(self, config, layer=3, in_channels=13, kd_flag=True, p_com_outage=0.0, num_agent=5, only_v2i=False):
        super().__init__()
        check_config(config)
        self.motion_state = config.motion_state
        self.out_seq_len = 1 if config.only_det else config.pred_len
        self.box_code_size = config.box_code_size
        self.category_num = config.category_num
        self.use_map = config.use_map
        self.anchor_num_per_loc = len(config.anchor_size)
        self.classification = ClassificationHead(config)
        self.regression = SingleRegressionHead(config)
        self.agent_num = num_agent
        self.kd_flag = kd_flag
        self.layer = layer
        self.p_com_outage = p_com_outage
        self.only_v2i = only_v2i
        self.precision = os.environ.get("V2X_PRECISION", "bf16")
        self.use_cuda_graph = os.environ.get("V2X_CUDA_GRAPH", "1") != "0"
        self._plans = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    # packed operands are caches derived from the nn.Parameters
    def invalidate(self):
        self._plans = {}

    def train(self, mode=True):
        self.invalidate()
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def _planes(self):
        if self.precision not in _PLANES:
            raise ValueError("precision must be one of %s" % list(_PLANES))
        return _PLANES[self.precision]

    def _check_eval(self):
        if self.training:
            raise NotImplementedError(
                "the sm_100a path implements inference (model.eval()); the training/backward step "
                "(SURVEY.md section 8(f1)) is not built yet")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            pass  # outputs are produced by custom kernels and carry no autograd graph

    def _state(self):
        return {k: v.detach() for k, v in self.state_dict().items()}

    def _get_plan(self, key, factory):
        plan = self._plans.get(key)
        if plan is None:
            plan = factory()
            if self.use_cuda_graph:
                plan.capture()
            self._plans[key] = plan
        return plan
