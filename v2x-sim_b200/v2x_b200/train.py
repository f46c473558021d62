"""Training step of the detection backbone on the sm_100a kernels: train-mode forward (BatchNorm with
batch statistics over all A*B maps + running-buffer update) and the backward pass, as a tape of kernel launches.

What runs where
  forward   conv (v2x_conv_fwd, raw output z = W * x + b, fp16 hi/lo, 3 tensor-core passes) -> v2x_bn_stats_fwd ->
            v2x_bn_finalize (scale / shift + running buffers) -> v2x_bn_relu_apply_fwd
  backward  v2x_bn_relu_bwd (ReLU mask + BN backward, d gamma / d beta) -> v2x_conv_wgrad (filter gradient) ->
            data gradient = v2x_conv_fwd with the transposed, 180-degree-rotated filter (transforms.dgrad_weights_stride1;
            stride-2 layers through v2x_resample2 zero-stuffing) -> v2x_act_add where a map feeds two consumers;
            nearest-upsample backward = v2x_resample2 mode 2
torch is plumbing (parameter storage, the autograd.Function boundary, the loss); no FLOP of the path runs in ATen.

Reference semantics: CP/models/det/backbone/Backbone.py:89-242 in .train() mode, DetModelBase.py:226-351 (heads), entered
through CP/utils/CoDetModule.py:217-291 (``loss.backward()``).  The loss stays the reference's python: it hands
d(loss)/d(loc), d(loss)/d(cls) to ``backward``.

Gradients are carried in the fp16 hi/lo act format scaled by a power of two S (chosen per step from the upstream
gradient's magnitude, like a static AMP loss scale) so they sit in fp16's normal range; parameter gradients are unscaled
in fp32 by the reducing kernels.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import torch

from . import ops
from ._lib import EPI_ACT, EPI_F32_NCHW, EPI_F32_SPLIT, check
from .ops import ConvLaunch, _ptr, _stream
from .transforms import dgrad_weights_stride1

BN_EPS, BN_MOMENTUM = 1e-5, 0.1
WGRAD_TC = os.environ.get("V2X_WGRAD", "tc") != "cuda"       # A/B switch: tensor-core vs CUDA-core weight gradient
WGRAD_TC_MIN_C = int(os.environ.get("V2X_WGRAD_TC_MIN_C", "16"))
PLANES = 2     # the training path always runs in the fp16 hi/lo format with 3 tensor-core passes (parity mode)


class Var:
    """A map on the tape: forward value (act [P, N, H, W, C]) and its accumulated gradient (same format, scaled by S)."""

    def __init__(self, act: torch.Tensor, c_log: Optional[int] = None):
        self.act = act
        self.c_log = c_log if c_log is not None else act.shape[-1]
        self.grad: Optional[torch.Tensor] = None

    def add_grad(self, lib, g: torch.Tensor):
        if self.grad is None:
            self.grad = g
        else:
            check(lib.v2x_act_add(_ptr(self.grad), _ptr(g), g[0].numel(), PLANES, _stream()), "v2x_act_add")


class Tape:
    """One training step: forward ops push their backward closures; ``backward`` replays them in reverse."""

    def __init__(self, params: Dict[str, torch.Tensor], buffers: Dict[str, torch.Tensor], device):
        self.lib = ops.require_gpu()
        self.p, self.b, self.dev = params, buffers, device
        self.grads: Dict[str, torch.Tensor] = {}
        self.back: List = []
        self.scale = 1.0

    # ---- helpers ----
    def _f32(self, *shape):
        return torch.zeros(shape, dtype=torch.float32, device=self.dev)

    def _f64(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.dev)

    def _param_grad(self, name):
        if name not in self.grads:
            self.grads[name] = torch.zeros_like(self.p[name], dtype=torch.float32)
        return self.grads[name]

    # ---- ops ----
    def upsample2(self, x: Var) -> Var:
        p, n, h, w, c = x.act.shape
        out = ops.empty_act(p, n, 2 * h, 2 * w, c, self.dev)
        check(self.lib.v2x_resample2(_ptr(x.act), _ptr(out), n, 2 * h, 2 * w, c, p, 1, _stream()), "v2x_resample2(up)")
        y = Var(out, x.c_log)

        def bwd():
            if y.grad is None:
                return
            g = ops.empty_act(p, n, h, w, c, self.dev)
            check(self.lib.v2x_resample2(_ptr(y.grad), _ptr(g), n, h, w, c, p, 2, _stream()), "v2x_resample2(sum)")
            x.add_grad(self.lib, g)
            y.grad = None
        self.back.append(bwd)
        return y

    def maxpool2(self, x: Var) -> Var:
        """nn.MaxPool2d(2) (SegModelBase.py:113) with its backward."""
        lib = self.lib
        y = Var(ops.maxpool2(x.act), x.c_log)
        p, n, h2, w2, c = x.act.shape

        def bwd():
            if y.grad is None:
                return
            g = torch.empty_like(x.act)
            check(lib.v2x_maxpool2_bwd(_ptr(x.act), _ptr(y.grad), _ptr(g), n, h2 // 2, w2 // 2, c, p, _stream()), "v2x_maxpool2_bwd")
            x.add_grad(lib, g)
            y.grad = None
        self.back.append(bwd)
        return y

    def upsample_bilinear2(self, x: Var) -> Var:
        """nn.Upsample(scale_factor=2, bilinear, align_corners=True) (SegModelBase.py:125) with its backward."""
        lib = self.lib
        y = Var(ops.upsample_bilinear2(x.act), x.c_log)
        p, n, h, w, c = x.act.shape

        def bwd():
            if y.grad is None:
                return
            dx = torch.empty((n, h, w, c), dtype=torch.float32, device=self.dev)
            check(lib.v2x_upsample_bilinear2_bwd(_ptr(y.grad), _ptr(dx), n, h, w, c, p, _stream()), "v2x_upsample_bilinear2_bwd")
            x.add_grad(lib, ops.pack_input(dx, c, p))
            y.grad = None
        self.back.append(bwd)
        return y

    def _conv_raw(self, wname, bname, srcs: List[Var], stride):
        """z = conv(cat(srcs), W) + b (no BN fold, no ReLU) as one act; returns (z act, cins logical)."""
        w = self.p[wname]
        w4 = w.reshape(w.shape[0], w.shape[1], *(w.shape[-2:] if w.dim() >= 4 and w.shape[-1] == 3 else (1, 1)))
        cins = [s.c_log for s in srcs]
        pc = ops.pack_conv(w4, self.p[bname] if bname else None, None, cins=cins, stride=stride, planes=PLANES,
                           device=self.dev, tap_pack=False, mmas=3)
        z = ops.conv(pc, [s.act for s in srcs], relu=False)
        return z, w4, cins

    def _conv_backward(self, wname, w4, srcs: List[Var], cins, stride, dz: torch.Tensor, need_input_grad):
        """Filter gradient (v2x_conv_wgrad per concat source) and data gradients (v2x_conv_fwd with the transposed,
        rotated filter; stride 2 through zero-stuffing) of z = conv(cat(srcs), w4)."""
        lib = self.lib
        p, n, ho, wo, co_pad = dz.shape
        co, ci_total, k = w4.shape[0], w4.shape[1], w4.shape[2]
        taps = k * k
        dw = self._param_grad(wname)
        ci_off = 0
        for s, c_log in zip(srcs, cins):
            # tensor-core wgrad (MN-major operands straight from the NHWC tensors); the CUDA-core kernel covers what it
            # cannot take: the stride-2 layer with fewer than 64 input channels (conv1_1)
            ci_phys = s.act.shape[-1]
            use_tc = WGRAD_TC and (stride == 1 or ci_phys % 64 == 0) and min(ci_phys, co_pad) >= WGRAD_TC_MIN_C
            fn = lib.v2x_conv_wgrad_tc if use_tc else lib.v2x_conv_wgrad
            check(fn(_ptr(dz), _ptr(s.act), n, ho, wo, co_pad, ci_phys, p, stride, taps, _ptr(dw),
                     co, c_log, ci_off, ci_total, 1.0 / self.scale, _stream()), "v2x_conv_wgrad")
            ci_off += c_log
        ci_off = 0
        dzin = dz
        if stride == 2 and any(need_input_grad):
            dzin = ops.empty_act(p, n, 2 * ho, 2 * wo, co_pad, self.dev)
            check(lib.v2x_resample2(_ptr(dz), _ptr(dzin), n, 2 * ho, 2 * wo, co_pad, p, 0, _stream()), "v2x_resample2(stuff)")
        for s, c_log, need in zip(srcs, cins, need_input_grad):
            if need:
                wt = dgrad_weights_stride1(w4[:, ci_off:ci_off + c_log])          # [c_log, co, k, k]
                pcd = ops.pack_conv(wt, None, None, cins=[co], stride=1, planes=PLANES, device=self.dev, tap_pack=False,
                                    mmas=3, cout_pad=s.act.shape[-1])
                g = ops.empty_act(p, n, s.act.shape[2], s.act.shape[3], s.act.shape[-1], self.dev)
                # very deep K (the seg GRU's data gradient: 9 x 1536 channels): an N tile of 128 keeps the hi/lo weight ring
                # of the halo + streamed-weight mode inside shared memory (the per-tap path is limited to 160 k-blocks)
                bn = 128 if (k * k * co_pad // 64 > 160 and pcd.cout % 128 == 0) else None
                ConvLaunch(pcd, [dzin], relu=False, out0=g, epilogue=EPI_ACT, block_n=bn)()
                s.add_grad(lib, g)
            ci_off += c_log

    def cbr(self, conv: str, bn: str, srcs: List[Var], stride=1, need_input_grad=None, bias=True) -> Var:
        """conv + BatchNorm(train) + ReLU (Backbone.py:102-136 in .train() mode)."""
        lib = self.lib
        need_input_grad = need_input_grad or [True] * len(srcs)
        z, w4, cins = self._conv_raw(conv + ".weight", conv + ".bias" if bias else None, srcs, stride)
        p, n, h, w, c = z.shape
        npix = n * h * w
        gamma, beta = self.p[bn + ".weight"], self.p[bn + ".bias"]
        rm, rv = self.b[bn + ".running_mean"], self.b[bn + ".running_var"]
        s0, s1 = self._f64(c), self._f64(c)
        check(lib.v2x_bn_stats_fwd(_ptr(z), npix, c, p, _ptr(s0), _ptr(s1), _stream()), "v2x_bn_stats_fwd")
        scale, shift, mean, invstd = (self._f32(c) for _ in range(4))
        check(lib.v2x_bn_finalize(_ptr(s0), _ptr(s1), npix, _ptr(gamma), _ptr(beta), BN_EPS, BN_MOMENTUM, _ptr(rm), _ptr(rv),
                                  _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), c, _stream()), "v2x_bn_finalize")
        nbt = self.b.get(bn + ".num_batches_tracked")
        if nbt is not None:
            nbt += 1
        y_act = torch.empty_like(z)
        check(lib.v2x_bn_relu_apply_fwd(_ptr(z), _ptr(y_act), npix, c, p, _ptr(scale), _ptr(shift), 1, _stream()),
              "v2x_bn_relu_apply_fwd")
        y = Var(y_act)

        def bwd():
            if y.grad is None:
                return
            dz = torch.empty_like(z)
            d1, d2 = self._f64(c), self._f64(c)
            check(lib.v2x_bn_relu_bwd(_ptr(y.grad), _ptr(z), _ptr(dz), npix, c, p, _ptr(scale), _ptr(shift), _ptr(mean),
                                      _ptr(invstd), 1, _ptr(d1), _ptr(d2), _stream()), "v2x_bn_relu_bwd")
            inv = 1.0 / self.scale
            check(lib.v2x_scale_to_f32(_ptr(d1), _ptr(self._param_grad(bn + ".bias")), c, inv, 1, _stream()), "d beta")
            check(lib.v2x_scale_to_f32(_ptr(d2), _ptr(self._param_grad(bn + ".weight")), c, inv, 1, _stream()), "d gamma")
            if bias:
                self._param_grad(conv + ".bias")    # a conv bias in front of a train-mode BN has an exactly zero gradient
            self._conv_backward(conv + ".weight", w4, srcs, cins, stride, dz, need_input_grad)
            y.grad = None
        self.back.append(bwd)
        return y

    def conv1x1_out(self, conv: str, src: Var, out_f32: torch.Tensor, nchw=False) -> "OutVar":
        """Final 1x1 conv with bias and fp32 output: NHWC (heads' conv2 / box_prediction.3, DetModelBase.py:283-329) or, with
        ``nchw``, [N, C, H, W] (the seg models' OutConv logits, SegModelBase.py:145-151)."""
        lib = self.lib
        w = self.p[conv + ".weight"]
        w4 = w.reshape(w.shape[0], w.shape[1], 1, 1)
        co = w4.shape[0]
        co_pad = 32 if co <= 32 else ((co + 63) // 64) * 64    # a valid N tile of v2x_conv_fwd for any kc
        pc = ops.pack_conv(w4, self.p[conv + ".bias"], None, cins=[src.c_log], planes=PLANES, device=self.dev, mmas=3,
                           cout_pad=co_pad)
        if nchw:
            ConvLaunch(pc, [src.act], epilogue=EPI_F32_NCHW, relu=False, out0=out_f32, block_n=min(co_pad, 64))()
        else:
            ConvLaunch(pc, [src.act], epilogue=EPI_F32_SPLIT, relu=False, out0=out_f32, split=co, block_n=min(co_pad, 64))()
        ov = OutVar(out_f32)

        def bwd():
            if ov.upstream is None:
                return
            p, n, h, wd, _ = src.act.shape
            if nchw:
                up = (ov.upstream.reshape(n, co, h, wd).to(torch.float32) * self.scale).contiguous()
                dz = ops.pack_input_nchw(up, co_pad, PLANES)
            else:
                up = (ov.upstream.reshape(n, h, wd, co).to(torch.float32) * self.scale).contiguous()
                dz = ops.pack_input(up, co_pad, PLANES)
            s0, s1 = self._f64(co_pad), self._f64(co_pad)
            check(lib.v2x_bn_stats_fwd(_ptr(dz), n * h * wd, co_pad, PLANES, _ptr(s0), _ptr(s1), _stream()), "bias grad")
            db = self._f32(co_pad)
            check(lib.v2x_scale_to_f32(_ptr(s0), _ptr(db), co_pad, 1.0 / self.scale, 0, _stream()), "bias grad")
            self._param_grad(conv + ".bias").add_(db[:co])
            dw = self._param_grad(conv + ".weight")
            check(lib.v2x_conv_wgrad(_ptr(dz), _ptr(src.act), n, h, wd, co_pad, src.act.shape[-1], PLANES, 1, 1, _ptr(dw), co,
                                     src.c_log, 0, w4.shape[1], 1.0 / self.scale, _stream()), "v2x_conv_wgrad(1x1)")
            wt = w4.transpose(0, 1).contiguous()                     # [cin, co, 1, 1]
            pcd = ops.pack_conv(wt, None, None, cins=[co], cin_pads=[co_pad], planes=PLANES, device=self.dev, mmas=3)
            g = torch.empty_like(src.act)
            ConvLaunch(pcd, [dz], relu=False, out0=g, epilogue=EPI_ACT)()
            src.add_grad(lib, g)
        self.back.append(bwd)
        return ov

    def warp_mean(self, x: Var, trans, num_agent, batch, agents, only_v2i=False, include_self=False) -> Var:
        """Neighbour mean of the warped maps (det V2VNet.py:85-98: self excluded; seg V2VNet.py:55-74: self included) with
        its backward (grid_sample backward)."""
        lib = self.lib
        out = ops.warp_mean(x.act, trans, num_agent, batch, agents, include_self=include_self, only_v2i=only_v2i)
        m = Var(out)
        p, n, h, w, c = x.act.shape

        def bwd():
            if m.grad is None:
                return
            dx = torch.empty((n, h, w, c), dtype=torch.float32, device=self.dev)
            check(lib.v2x_warp_mean_bwd(_ptr(m.grad), _ptr(dx), _ptr(trans), _ptr(num_agent), batch, agents, h, w, c, p,
                                        int(include_self), int(only_v2i), _stream()), "v2x_warp_mean_bwd")
            x.add_grad(lib, ops.pack_input(dx, c, p))
            m.grad = None
        self.back.append(bwd)
        return m

    def warp_reduce(self, x: Var, trans, num_agent, batch, agents, mode: str, only_v2i=False) -> Var:
        """Mean / Sum / Max fusion of the warped member maps (FusionBase.py:41-63 with MeanFusion.py:11-12,
        SumFusion.py:20-21, MaxFusion.py:20-21; absent agent slots keep their own map) with its backward."""
        lib = self.lib
        out = ops.warp_reduce(x.act, trans, num_agent, batch, agents, mode, only_v2i=only_v2i)
        y = Var(out)
        p, n, h, w, c = x.act.shape

        def bwd():
            if y.grad is None:
                return
            dx = torch.empty((n, h, w, c), dtype=torch.float32, device=self.dev)
            check(lib.v2x_warp_reduce_bwd(_ptr(y.grad), _ptr(x.act), _ptr(dx), _ptr(None), _ptr(trans), _ptr(num_agent), batch,
                                          agents, h, w, c, p, ops.REDUCE_MODES[mode], int(only_v2i), _stream()),
                  "v2x_warp_reduce_bwd")
            x.add_grad(lib, ops.pack_input(dx, c, p))
            y.grad = None
        self.back.append(bwd)
        return y

    def cat_modulate(self, x: Var, mean: Var, num_agent, batch, agents, conv="_modulation_layer_3._conv1_1",
                     bn="_modulation_layer_3._bn1_1") -> Var:
        """CatFusion's ModulationLayer3 in train mode (CatFusion.py:23-41): relu(bn(conv1x1(cat[tg, mean]))) evaluated by the
        reference ONCE PER PRESENT AGENT with a batch of one map, in the order scene-major / agent-minor (FusionBase.py:40-63)
        -- so the BatchNorm statistics are per map (1024 pixels) and the running buffers take one momentum update per call,
        in that order.  The 1x1 conv runs as one launch over all maps; statistics / normalise / backward run per map.
        Absent agent slots keep their own map."""
        lib = self.lib
        z, w4, cins = self._conv_raw(conv + ".weight", conv + ".bias", [x, mean], 1)
        p, n, h, w, c = z.shape
        hw = h * w
        gamma, beta = self.p[bn + ".weight"], self.p[bn + ".bias"]
        rm, rv = self.b[bn + ".running_mean"], self.b[bn + ".running_var"]
        na = num_agent[:, 0].tolist()                     # host copy: the call order is data dependent (one sync per step)
        order = [batch * i + b for b in range(batch) for i in range(min(int(na[b]), agents))]
        present = set(order)
        out = x.act.clone()                                # absent slots: own map
        saved = {}
        for u in order:
            zu = z[:, u:u + 1].contiguous()
            s0, s1 = self._f64(c), self._f64(c)
            check(lib.v2x_bn_stats_fwd(_ptr(zu), hw, c, p, _ptr(s0), _ptr(s1), _stream()), "v2x_bn_stats_fwd")
            scale, shift, mu, invstd = (self._f32(c) for _ in range(4))
            check(lib.v2x_bn_finalize(_ptr(s0), _ptr(s1), hw, _ptr(gamma), _ptr(beta), BN_EPS, BN_MOMENTUM, _ptr(rm), _ptr(rv),
                                      _ptr(scale), _ptr(shift), _ptr(mu), _ptr(invstd), c, _stream()), "v2x_bn_finalize")
            yu = torch.empty_like(zu)
            check(lib.v2x_bn_relu_apply_fwd(_ptr(zu), _ptr(yu), hw, c, p, _ptr(scale), _ptr(shift), 1, _stream()),
                  "v2x_bn_relu_apply_fwd")
            out[:, u:u + 1].copy_(yu)
            saved[u] = (zu, scale, shift, mu, invstd)
        nbt = self.b.get(bn + ".num_batches_tracked")
        if nbt is not None:
            nbt += len(order)
        y = Var(out)

        def bwd():
            if y.grad is None:
                return
            inv = 1.0 / self.scale
            dz = torch.zeros_like(z)
            for u in order:
                zu, scale, shift, mu, invstd = saved[u]
                dyu = y.grad[:, u:u + 1].contiguous()
                dzu = torch.empty_like(zu)
                d1, d2 = self._f64(c), self._f64(c)
                check(lib.v2x_bn_relu_bwd(_ptr(dyu), _ptr(zu), _ptr(dzu), hw, c, p, _ptr(scale), _ptr(shift), _ptr(mu),
                                          _ptr(invstd), 1, _ptr(d1), _ptr(d2), _stream()), "v2x_bn_relu_bwd")
                check(lib.v2x_scale_to_f32(_ptr(d1), _ptr(self._param_grad(bn + ".bias")), c, inv, 1, _stream()), "d beta")
                check(lib.v2x_scale_to_f32(_ptr(d2), _ptr(self._param_grad(bn + ".weight")), c, inv, 1, _stream()), "d gamma")
                dz[:, u:u + 1].copy_(dzu)
            self._param_grad(conv + ".bias")      # in front of a train-mode BN: exactly zero gradient
            self._conv_backward(conv + ".weight", w4, [x, mean], cins, 1, dz, [True, True])
            absent = [u for u in range(n) if u not in present]
            if absent:                            # pass-through of the slots the fuse did not rewrite
                g = torch.zeros_like(x.act)
                g[:, absent] = y.grad[:, absent]
                x.add_grad(lib, g)
            y.grad = None
        self.back.append(bwd)
        return y

    def warp_member(self, x: Var, k: int, trans, num_agent, batch, agents, only_v2i=False) -> Var:
        """Agent k's map warped into every target's frame (identity where k is the target itself): one member slot of the
        neighbour lists FusionBase builds (DetModelBase.py:171-209), as one launch over all targets, with its backward."""
        lib = self.lib
        onehot = torch.zeros((batch, agents, agents), dtype=torch.float32, device=self.dev)
        onehot[:, :, k] = 1.0
        y = Var(ops.warp_weighted(x.act, trans, num_agent, onehot, batch, agents, per_pixel=False, only_v2i=only_v2i))
        p_, n, h, w, c = x.act.shape

        def bwd():
            if y.grad is None:
                return
            dx = torch.empty((n, h, w, c), dtype=torch.float32, device=self.dev)
            check(lib.v2x_warp_reduce_bwd(_ptr(y.grad), _ptr(x.act), _ptr(dx), _ptr(onehot), _ptr(trans), _ptr(num_agent), batch,
                                          agents, h, w, c, p_, 3, int(only_v2i), _stream()), "v2x_warp_reduce_bwd(member)")
            x.add_grad(lib, ops.pack_input(dx, c, p_))
            y.grad = None
        self.back.append(bwd)
        return y

    def conv_plain(self, conv: str, srcs: List[Var]) -> Var:
        """1x1 / 3x3 conv + bias without BatchNorm or ReLU (the consumer normalises per call), with its backward.  Its bias
        sits in front of a train-mode BatchNorm: exactly zero gradient."""
        z, w4, cins = self._conv_raw(conv + ".weight", conv + ".bias", srcs, 1)
        y = Var(z)

        def bwd():
            if y.grad is None:
                return
            self._param_grad(conv + ".bias")
            self._conv_backward(conv + ".weight", w4, srcs, cins, 1, y.grad, [True] * len(srcs))
            y.grad = None
        self.back.append(bwd)
        return y

    def disco_fuse(self, x: Var, trans, num_agent, batch, agents, only_v2i=False, prefix="pixel_weighted_fusion.") -> Var:
        """DiscoNet's pixel-wise weighted fuse in train mode (DiscoNet.py:80-107, 132-155; kd_flag = 0): for every present
        target i and list member k the weight net maps cat[tg_i, warp_{k->i}(x_k)] to a per-pixel score; the members are mixed
        with the per-pixel softmax of the scores, and -- unlike AgentWise -- the weight net IS trained through it.
        On the tape: the member warps (warp_member), the weight net's first layer as one 1x1-conv launch per member slot
        (2C -> 128 on every pixel of every pair: 94% of its FLOPs; filter gradient on tcgen05) and the fuse
        (v2x_warp_weighted_fwd / v2x_warp_weighted_bwd: map gradient + softmax backward per pixel).  Between them the per-call
        tail of the weight net -- BatchNorm with per-call statistics (one map per call, running buffers updated call by call
        in the reference's order), 128 -> 32 -> 8 -> 1 per pixel: 9 MFLOP per call -- is a torch-autograd island in fp32."""
        import torch.nn.functional as F
        lib = self.lib
        p_, n, h, w, c = x.act.shape
        hw = h * w
        na = [min(int(v), agents) for v in num_agent[:, 0].tolist()]
        zs = []
        for k in range(agents):
            wk = self.warp_member(x, k, trans, num_agent, batch, agents, only_v2i=only_v2i)
            zs.append(self.conv_plain(prefix + "conv1_1", [x, wk]))
        tail_names = [prefix + t for t in ("bn1_1.weight", "bn1_1.bias", "conv1_2.weight", "conv1_2.bias", "bn1_2.weight",
                                           "bn1_2.bias", "conv1_3.weight", "conv1_3.bias", "bn1_3.weight", "bn1_3.bias",
                                           "conv1_4.weight", "conv1_4.bias")]
        calls = []                                   # (b, i, k, score map [hw]) in the reference's call order
        with torch.enable_grad():
            z1 = [ops.act_to_float(z.act).detach().requires_grad_(True) for z in zs]      # [units, 128, h, w] fp32 each
            tp = {k: self.p[k].detach().to(torch.float32).requires_grad_(True) for k in tail_names}
            for b in range(batch):
                for i in range(na[b]):
                    members = [i] + [k for k in range(na[b]) if k != i and not (only_v2i and i != 0 and k != 0)]
                    for k in members:
                        t = z1[k][batch * i + b: batch * i + b + 1]
                        for li, cname in ((1, None), (2, "conv1_2"), (3, "conv1_3")):
                            if cname is not None:
                                t = F.conv2d(t, tp[prefix + cname + ".weight"], tp[prefix + cname + ".bias"])
                            bn = prefix + "bn1_%d" % li
                            t = F.relu(F.batch_norm(t, self.b[bn + ".running_mean"], self.b[bn + ".running_var"], tp[bn + ".weight"],
                                                    tp[bn + ".bias"], True, BN_MOMENTUM, BN_EPS))
                            nbt = self.b.get(bn + ".num_batches_tracked")
                            if nbt is not None:
                                nbt += 1
                        t = F.relu(F.conv2d(t, tp[prefix + "conv1_4.weight"], tp[prefix + "conv1_4.bias"]))
                        calls.append((b, i, k, t.reshape(hw)))
        scores = torch.zeros((batch, agents, agents, hw), dtype=torch.float32, device=self.dev)
        for b, i, k, sm in calls:
            scores[b, i, k] = sm.detach()
        holder: Dict[str, torch.Tensor] = {}

        def island_bwd():
            ds = holder.get("dscores")
            if ds is None or not calls:
                return
            leaves = z1 + [tp[k] for k in tail_names]
            grads = torch.autograd.grad([sm for _, _, _, sm in calls], leaves, grad_outputs=[ds[b, i, k] for b, i, k, _ in calls],
                                        allow_unused=True)
            for z, gz in zip(zs, grads[:len(z1)]):
                if gz is not None:
                    z.add_grad(lib, ops.pack_input_nchw((gz * self.scale).contiguous(), int(gz.shape[1]), PLANES))
            for k, gk in zip(tail_names, grads[len(z1):]):
                if gk is not None:
                    self._param_grad(k).add_(gk)
        self.back.append(island_bwd)
        out = ops.warp_weighted(x.act, trans, num_agent, scores, batch, agents, per_pixel=True, only_v2i=only_v2i)
        y = Var(out)

        def bwd():
            if y.grad is None:
                return
            dx = torch.empty((n, h, w, c), dtype=torch.float32, device=self.dev)
            dscores = torch.empty((batch, agents, agents, hw), dtype=torch.float32, device=self.dev)
            check(lib.v2x_warp_weighted_bwd(_ptr(y.grad), _ptr(x.act), _ptr(dx), _ptr(dscores), _ptr(scores), _ptr(trans),
                                            _ptr(num_agent), batch, agents, h, w, c, p_, int(only_v2i), _stream()),
                  "v2x_warp_weighted_bwd")
            holder["dscores"] = dscores * (1.0 / self.scale)
            x.add_grad(lib, ops.pack_input(dx, c, p_))
            y.grad = None
        self.back.append(bwd)
        return y

    def agent_weighted_fuse(self, x: Var, trans, num_agent, batch, agents, only_v2i=False,
                            prefix="agent_weighted_fusion.") -> Var:
        """AgentWiseWeightedFusion in train mode (AgentWiseWeightedFusion.py:24-43, 44-76; seg twin): for every present target
        i and list member k (self first, then the other present agents in ascending order) the weight net maps
        cat[tg_i, warp_{k->i}(x_k)] to one scalar; the members are mixed with the softmax of those scalars.  The reference
        rebuilds the scalars with ``torch.tensor([...])``, which DETACHES them: the weight net gets no gradient and the fuse
        back-propagates to the maps only, with constant coefficients (v2x_warp_reduce_bwd mode 3).  The weight net still runs
        in train mode -- BatchNorm statistics per call (one map), running buffers updated call by call in that order.
        Its first layer (512 -> 128 over every pixel of every pair, 94% of its FLOPs) runs here as one 1x1-conv launch per
        member slot over the warped maps (v2x_warp_weighted_fwd with a one-hot coefficient); the per-call tail
        (BN, 128 -> 32 -> 8 -> 1 per pixel, the 32x32 conv1_5: 9 MFLOP per call, forward only) is evaluated with torch in
        fp32, in the reference's call order, on the module's own running buffers."""
        import torch.nn.functional as F
        lib = self.lib
        p_, n, h, w, c = x.act.shape
        na = [min(int(v), agents) for v in num_agent[:, 0].tolist()]
        f32 = lambda k: self.p[prefix + k].detach().to(torch.float32)   # noqa: E731
        z1 = []
        for k in range(agents):                     # member slot k: agent k's map warped into every target's frame
            onehot = torch.zeros((batch, agents, agents), dtype=torch.float32, device=self.dev)
            onehot[:, :, k] = 1.0
            wk = Var(ops.warp_weighted(x.act, trans, num_agent, onehot, batch, agents, per_pixel=False, only_v2i=only_v2i))
            zk, _, _ = self._conv_raw(prefix + "conv1_1.weight", prefix + "conv1_1.bias", [x, wk], 1)
            z1.append(ops.act_to_float(zk))         # [units, 128, h, w] fp32
        w5f = torch.flip(f32("conv1_5.weight").reshape(h, w), (0,))      # the 32x32 "valid" conv sees the H-flipped map
        coef = torch.zeros((batch, agents, agents), dtype=torch.float32, device=self.dev)
        with torch.no_grad():
            for b in range(batch):
                for i in range(na[b]):
                    members = [i] + [k for k in range(na[b]) if k != i and not (only_v2i and i != 0 and k != 0)]
                    scal = []
                    for k in members:
                        t = z1[k][batch * i + b: batch * i + b + 1]
                        for li, cname in ((1, None), (2, "conv1_2"), (3, "conv1_3")):
                            if cname is not None:
                                t = F.conv2d(t, f32(cname + ".weight"), f32(cname + ".bias"))
                            bn = prefix + "bn1_%d" % li
                            t = F.relu(F.batch_norm(t, self.b[bn + ".running_mean"], self.b[bn + ".running_var"], f32("bn1_%d.weight" % li),
                                                    f32("bn1_%d.bias" % li), True, BN_MOMENTUM, BN_EPS))
                            nbt = self.b.get(bn + ".num_batches_tracked")
                            if nbt is not None:
                                nbt += 1
                        t = F.relu(F.conv2d(t, f32("conv1_4.weight"), f32("conv1_4.bias")))
                        scal.append(F.relu((t.reshape(h, w) * w5f).sum() + f32("conv1_5.bias").reshape(())))
                    soft = torch.softmax(torch.stack(scal), 0)
                    for m, k in enumerate(members):
                        coef[b, i, k] = soft[m]
        out = ops.warp_weighted(x.act, trans, num_agent, coef, batch, agents, per_pixel=False, only_v2i=only_v2i)
        y = Var(out)

        def bwd():
            if y.grad is None:
                return
            dx = torch.empty((n, h, w, c), dtype=torch.float32, device=self.dev)
            check(lib.v2x_warp_reduce_bwd(_ptr(y.grad), _ptr(x.act), _ptr(dx), _ptr(coef), _ptr(trans), _ptr(num_agent), batch,
                                          agents, h, w, c, p_, 3, int(only_v2i), _stream()), "v2x_warp_reduce_bwd(coef)")
            x.add_grad(lib, ops.pack_input(dx, c, p_))
            y.grad = None
        self.back.append(bwd)
        return y

    def gated_fuse(self, x: Var, coef: torch.Tensor, dcoef_out: Dict[str, torch.Tensor], trans, num_agent, batch, agents,
                   warp_flag=1, only_v2i=False) -> Var:
        """when2com fuse out[b,q] = sum_k coef[b,k,q] * val[b,k,q] (When2com.py:199-225, 397-412) with its backward: the map
        gradient goes onto the tape, d(loss)/d(coef) (unscaled, fp32 [B,A,A]) is left in ``dcoef_out["dcoef"]`` for the
        attention island that produced ``coef``."""
        lib = self.lib
        out = ops.warp_gated(x.act, trans, num_agent, coef, batch, agents, warp_flag=warp_flag, only_v2i=only_v2i)
        y = Var(out)
        p, n, h, w, c = x.act.shape

        def bwd():
            if y.grad is None:
                return
            dx = torch.empty((n, h, w, c), dtype=torch.float32, device=self.dev)
            dcoef = torch.empty((batch, agents, agents), dtype=torch.float32, device=self.dev)
            check(lib.v2x_warp_gated_bwd(_ptr(y.grad), _ptr(x.act), _ptr(dx), _ptr(dcoef), _ptr(coef), _ptr(trans),
                                         _ptr(num_agent), batch, agents, h, w, c, p, int(warp_flag), int(only_v2i), _stream()),
                  "v2x_warp_gated_bwd")
            dcoef_out["dcoef"] = dcoef * (1.0 / self.scale)
            x.add_grad(lib, ops.pack_input(dx, c, p))
            y.grad = None
        self.back.append(bwd)
        return y

    def gru_round(self, h: Var, mean: Var, x_pass: Var, num_agent, batch, agents, prefix="convgru.") -> Var:
        """One zero-hidden ConvGRU step on cat([h, mean]) (V2VNet.py:99-101; functional.py:84-105).  The kernels run in the
        un-flipped domain, so the filter rows are mirrored; the filter gradient is accumulated in that
        mirrored form under the key ``<prefix>weight_ih_l0`` and un-mirrored once at the end of the step."""
        lib = self.lib
        w_used = torch.flip(self.p[prefix + "weight_ih_l0"], (2,)).contiguous()      # index permutation, no arithmetic
        c = w_used.shape[0] // 3
        b_ih, b_hh = self.p[prefix + "bias_ih_l0"], self.p[prefix + "bias_hh_l0"]
        pc = ops.pack_conv(w_used, b_ih, None, cins=[c, w_used.shape[1] - c], planes=PLANES, device=self.dev, tap_pack=False,
                           mmas=3)
        a = ops.conv(pc, [h.act, mean.act], relu=False)
        p, n, hh, ww, _ = a.shape
        npix, hw = n * hh * ww, hh * ww
        out = torch.empty_like(h.act)
        bhh = b_hh.to(torch.float32).contiguous()
        check(lib.v2x_gru_gates_fwd(_ptr(a), _ptr(bhh), _ptr(x_pass.act), _ptr(out), npix, hw, c, p, _ptr(num_agent), batch, agents,
                                    _stream()), "v2x_gru_gates_fwd")
        y = Var(out)

        def bwd():
            if y.grad is None:
                return
            da = torch.empty_like(a)
            dpass = torch.empty_like(h.act)
            dbhn = self._f64(c)
            check(lib.v2x_gru_gates_bwd(_ptr(y.grad), _ptr(a), _ptr(bhh), _ptr(da), _ptr(dpass), npix, hw, c, p, _ptr(num_agent),
                                        batch, agents, _ptr(dbhn), _stream()), "v2x_gru_gates_bwd")
            inv = 1.0 / self.scale
            s0, s1 = self._f64(3 * c), self._f64(3 * c)
            check(lib.v2x_bn_stats_fwd(_ptr(da), npix, 3 * c, p, _ptr(s0), _ptr(s1), _stream()), "sum da")
            g_ih, g_hh = self._param_grad(prefix + "bias_ih_l0"), self._param_grad(prefix + "bias_hh_l0")
            check(lib.v2x_scale_to_f32(_ptr(s0), _ptr(g_ih), 3 * c, inv, 1, _stream()), "d b_ih")
            check(lib.v2x_scale_to_f32(_ptr(s0), _ptr(g_hh), 2 * c, inv, 1, _stream()), "d b_hh (r, z)")
            check(lib.v2x_scale_to_f32(_ptr(dbhn), _ptr(g_hh[2 * c:]), c, inv, 1, _stream()), "d b_hh (n)")
            self._param_grad(prefix + "weight_hh_l0")     # conv over the all-zero hidden state: exactly zero gradient
            self._conv_backward(prefix + "weight_ih_l0", w_used, [h, mean], [c, w_used.shape[1] - c], 1, da, [True, True])
            h.add_grad(lib, dpass)
            y.grad = None
        self.back.append(bwd)
        return y

    def backward(self):
        for fn in reversed(self.back):
            fn()
        self.back = []


class OutVar:
    def __init__(self, t):
        self.value = t
        self.upstream: Optional[torch.Tensor] = None


def choose_scale(*upstreams) -> float:
    """Power-of-two gradient scale S that puts the largest upstream magnitude near 2^9 (one host sync per step)."""
    m = max(float(u.abs().max().item()) for u in upstreams if u is not None)
    if not math.isfinite(m) or m <= 0.0:
        return 1.0
    return float(2.0 ** round(math.log2(512.0 / m)))


# =====================================================================================================================
# network builders
# =====================================================================================================================
MIN_COMPRESSED_CHANNELS = 32    # narrowest train-mode BatchNorm the tape's kernels take without channel padding


def compress_pair(t: Tape, pre: str, x: Var) -> Var:
    """com_compresser -> BN -> ReLU -> com_decompresser -> BN -> ReLU of the communicated layer in train mode
    (Backbone.py:138-141, SegModelBase.py:29-43 as applied at UNet.py:30-32): two 1x1 ``cbr`` steps on the tape.  The
    eval plans zero-pad a compressed width below 32 (nets.pack_compress_pair); a train-mode BatchNorm over padded channels
    would need padded affine / running buffers, so training takes >= 32 compressed channels (det compress_level <= 3,
    seg <= 4) and refuses the rest."""
    cc = int(t.p[pre + "com_compresser.weight"].shape[0])
    if cc < MIN_COMPRESSED_CHANNELS:
        raise NotImplementedError("training with %d compressed channels is not built on the sm_100a path (>= %d: "
                                  "compress_level <= 3 for det, <= 4 for seg)" % (cc, MIN_COMPRESSED_CHANNELS))
    c = t.cbr(pre + "com_compresser", pre + "bn_compress", [x])
    return t.cbr(pre + "com_decompresser", pre + "bn_decompress", [c])


def backbone_encode(t: Tape, pre: str, x_in: Var):
    """Backbone.encode in train mode (Backbone.py:89-143); returns [x, x_1, x_2, x_3, x_4].  With compress_level > 0 the
    returned x_3 went through the compresser pair; x_4 was computed from the uncompressed map (:131-141)."""
    x = t.cbr(pre + "conv_pre_1", pre + "bn_pre_1", [x_in], need_input_grad=[False])
    x0 = t.cbr(pre + "conv_pre_2", pre + "bn_pre_2", [x])
    x = t.cbr(pre + "conv1_1", pre + "bn1_1", [x0], stride=2)
    x = t.cbr(pre + "conv1_2", pre + "bn1_2", [x])
    x1 = t.cbr(pre + "conv3d_1.conv3d", pre + "conv3d_1.bn3d", [x])
    x = t.cbr(pre + "conv2_1", pre + "bn2_1", [x1], stride=2)
    x = t.cbr(pre + "conv2_2", pre + "bn2_2", [x])
    x2 = t.cbr(pre + "conv3d_2.conv3d", pre + "conv3d_2.bn3d", [x])
    x = t.cbr(pre + "conv3_1", pre + "bn3_1", [x2], stride=2)
    x3 = t.cbr(pre + "conv3_2", pre + "bn3_2", [x])
    x = t.cbr(pre + "conv4_1", pre + "bn4_1", [x3], stride=2)
    x4 = t.cbr(pre + "conv4_2", pre + "bn4_2", [x])
    if pre + "com_compresser.weight" in t.p:
        x3 = compress_pair(t, pre, x3)
    return [x0, x1, x2, x3, x4]


def backbone_decode(t: Tape, pre: str, x0, x1, x2, x3, x4, kd: bool = False):
    """Backbone.decode in train mode (Backbone.py:145-242): cat((up2(deep), skip)) -> two CBRs, four times.
    ``kd``: return [x_8, x_7, x_6, x_5] (what STPN_KD / the kd_flag forwards hand to the distillation loss)."""
    x = t.cbr(pre + "conv5_1", pre + "bn5_1", [t.upsample2(x4), x3])
    x5 = t.cbr(pre + "conv5_2", pre + "bn5_2", [x])
    x = t.cbr(pre + "conv6_1", pre + "bn6_1", [t.upsample2(x5), x2])
    x6 = t.cbr(pre + "conv6_2", pre + "bn6_2", [x])
    x = t.cbr(pre + "conv7_1", pre + "bn7_1", [t.upsample2(x6), x1])
    x7 = t.cbr(pre + "conv7_2", pre + "bn7_2", [x])
    x = t.cbr(pre + "conv8_1", pre + "bn8_1", [t.upsample2(x7), x0])
    x8 = t.cbr(pre + "conv8_2", pre + "bn8_2", [x])
    return [x8, x7, x6, x5] if kd else x8


def det_heads(t: Tape, x8: Var, n: int):
    """ClassificationHead / SingleRegressionHead in train mode (DetModelBase.py:268-351): fp32 NHWC outputs."""
    dev = t.dev
    h, w = x8.act.shape[2], x8.act.shape[3]
    n_cls = t.p["classification.conv2.weight"].shape[0]
    n_loc = t.p["regression.box_prediction.3.weight"].shape[0]
    cls = torch.empty((n, h, w, n_cls), dtype=torch.float32, device=dev)
    loc = torch.empty((n, h, w, n_loc), dtype=torch.float32, device=dev)
    c1 = t.cbr("classification.conv1", "classification.bn1", [x8])
    o_cls = t.conv1x1_out("classification.conv2", c1, cls)
    r1 = t.cbr("regression.box_prediction.0", "regression.box_prediction.1", [x8])
    o_loc = t.conv1x1_out("regression.box_prediction.3", r1, loc)
    return o_loc, o_cls


class FaFNetTrainStep(torch.autograd.Function):
    """One train-mode forward of FaFNet / STPN (FaFNet.py:28-39) with its backward, behind torch.autograd so the
    reference's ``loss.backward()`` / Adam step (CoDetModule.py:283-291) drive it unchanged.
    Inputs: (module, bevs, *parameters in named_parameters() order); outputs (loc, cls) in the reference layout."""

    @staticmethod
    def forward(ctx, module, bevs, *params):
        names = [k for k, _ in module.named_parameters()]
        p = {k: v.detach() for k, v in zip(names, params)}
        b = {k: v for k, v in module.named_buffers()}
        dev = bevs.device
        n = int(bevs.shape[0])
        tape = Tape(p, b, dev)
        x_in = Var(ops.pack_input(bevs.reshape(n, 256, 256, -1).to(torch.float32).contiguous(), 16, PLANES), c_log=int(bevs.shape[-1]))
        enc = backbone_encode(tape, "stpn.", x_in)
        x8 = backbone_decode(tape, "stpn.", *enc)
        o_loc, o_cls = det_heads(tape, x8, n)
        ctx.tape, ctx.names, ctx.o_loc, ctx.o_cls = tape, names, o_loc, o_cls
        ctx.shapes = [v.shape for v in params]
        loc = o_loc.value.view(n, 256, 256, 6, 1, 6)
        cls = o_cls.value.view(n, -1, 2)
        return loc, cls

    @staticmethod
    def backward(ctx, dloc, dcls):
        tape = ctx.tape
        ups = [u for u in (dloc, dcls) if u is not None]
        tape.scale = choose_scale(*ups)
        ctx.o_loc.upstream, ctx.o_cls.upstream = dloc, dcls
        tape.backward()
        grads = []
        for k, shape in zip(ctx.names, ctx.shapes):
            g = tape.grads.get(k)
            grads.append(None if g is None else g.reshape(shape))
        ctx.tape = None
        return (None, None, *grads)


class V2VNetTrainStep(torch.autograd.Function):
    """One train-mode forward of det V2VNet (V2VNet.py:47-120: encoder -> 3 x [warp / neighbour mean / ConvGRU] at layer 3
    -> decoder -> heads) with its backward.  Inputs: (module, bevs, trans_matrices, num_agent_tensor, batch_size,
    *parameters in named_parameters() order)."""

    @staticmethod
    def forward(ctx, module, bevs, trans, nat, batch, *params):
        names = [k for k, _ in module.named_parameters()]
        p = {k: v.detach() for k, v in zip(names, params)}
        b = {k: v for k, v in module.named_buffers()}
        dev = bevs.device
        n = int(bevs.shape[0])
        agents = n // batch
        tape = Tape(p, b, dev)
        trans = trans.to(device=dev, dtype=torch.float64).contiguous()
        nat = nat.to(device=dev, dtype=torch.int64).contiguous()
        x_in = Var(ops.pack_input(bevs.reshape(n, 256, 256, -1).to(torch.float32).contiguous(), 16, PLANES), c_log=int(bevs.shape[-1]))
        x0, x1, x2, x3, x4 = backbone_encode(tape, "u_encoder.", x_in)
        mean = tape.warp_mean(x3, trans, nat, batch, agents, only_v2i=bool(module.only_v2i))
        h = x3
        for _ in range(module.gnn_iter_num):
            h = tape.gru_round(h, mean, x3, nat, batch, agents)
        x8 = backbone_decode(tape, "decoder.", x0, x1, x2, h, x4)
        o_loc, o_cls = det_heads(tape, x8, n)
        ctx.tape, ctx.names, ctx.o_loc, ctx.o_cls = tape, names, o_loc, o_cls
        ctx.shapes = [v.shape for v in params]
        return o_loc.value.view(n, 256, 256, 6, 1, 6), o_cls.value.view(n, -1, 2)

    @staticmethod
    def backward(ctx, dloc, dcls):
        tape = ctx.tape
        tape.scale = choose_scale(*[u for u in (dloc, dcls) if u is not None])
        ctx.o_loc.upstream, ctx.o_cls.upstream = dloc, dcls
        tape.backward()
        k = "convgru.weight_ih_l0"
        if k in tape.grads:
            tape.grads[k] = torch.flip(tape.grads[k], (2,))      # back from the mirrored-row form the kernels ran in
        grads = []
        for name, shape in zip(ctx.names, ctx.shapes):
            g = tape.grads.get(name)
            grads.append(None if g is None else g.reshape(shape))
        ctx.tape = None
        return (None, None, None, None, None, *grads)


class FusionTrainStep(torch.autograd.Function):
    """One train-mode forward of the intermediate-fusion baselines MeanFusion / SumFusion / MaxFusion / CatFusion
    (FusionBase.py:23-75: encoder -> fuse of the warped member maps at layer 3 -> decoder -> heads) with its backward.
    Inputs: (module, kind, bevs, trans_matrices, num_agent_tensor, batch_size, *parameters in named_parameters() order)."""

    @staticmethod
    def forward(ctx, module, kind, bevs, trans, nat, batch, *params):
        names = [k for k, _ in module.named_parameters()]
        p = {k: v.detach() for k, v in zip(names, params)}
        b = {k: v for k, v in module.named_buffers()}
        dev = bevs.device
        n = int(bevs.shape[0])
        agents = n // batch
        tape = Tape(p, b, dev)
        trans = trans.to(device=dev, dtype=torch.float64).contiguous()
        nat = nat.to(device=dev, dtype=torch.int64).contiguous()
        x_in = Var(ops.pack_input(bevs.reshape(n, 256, 256, -1).to(torch.float32).contiguous(), 16, PLANES), c_log=int(bevs.shape[-1]))
        x0, x1, x2, x3, x4 = backbone_encode(tape, "u_encoder.", x_in)
        if kind == "disco":
            fused = tape.disco_fuse(x3, trans, nat, batch, agents, only_v2i=bool(module.only_v2i))
        elif kind == "agent":
            fused = tape.agent_weighted_fuse(x3, trans, nat, batch, agents, only_v2i=bool(module.only_v2i))
        elif kind == "cat":     # CatFusion.py:23-27: mean of the member stack, then the modulation layer on cat[tg, mean]
            mean = tape.warp_reduce(x3, trans, nat, batch, agents, "mean", only_v2i=bool(module.only_v2i))
            fused = tape.cat_modulate(x3, mean, nat, batch, agents)
        else:
            fused = tape.warp_reduce(x3, trans, nat, batch, agents, kind, only_v2i=bool(module.only_v2i))
        kd = int(getattr(module, "kd_flag", 0)) == 1
        dec = backbone_decode(tape, "decoder.", x0, x1, x2, fused, x4, kd=kd)
        x8 = dec[0] if kd else dec
        o_loc, o_cls = det_heads(tape, x8, n)
        ctx.tape, ctx.names, ctx.o_loc, ctx.o_cls = tape, names, o_loc, o_cls
        ctx.shapes = [v.shape for v in params]
        # kd_flag == 1 (FusionBase.py:72-73, DiscoNet.py:125-127): the decoder maps x_8, x_7, x_6, x_5 and the fused layer
        # also leave the step, as fp32 NCHW tensors WITH gradients -- FaFModule.get_kd_loss (CoDetModule.py:257-260) pulls the
        # student's x_5, x_6, x_7 and fused layer towards the teacher's
        ctx.kd_vars = (dec + [fused]) if kd else []
        extra = tuple(ops.act_to_float(v.act)[:, :v.c_log].contiguous() for v in ctx.kd_vars)
        return (o_loc.value.view(n, 256, 256, 6, 1, 6), o_cls.value.view(n, -1, 2), *extra)

    @staticmethod
    def backward(ctx, dloc, dcls, *dkd):
        tape = ctx.tape
        tape.scale = choose_scale(*[u for u in (dloc, dcls, *dkd) if u is not None])
        ctx.o_loc.upstream, ctx.o_cls.upstream = dloc, dcls
        for v, g in zip(ctx.kd_vars, dkd):          # distillation gradients enter the tape at the maps they belong to
            if g is not None:
                gs = (g.to(torch.float32) * tape.scale).contiguous()
                v.add_grad(tape.lib, ops.pack_input_nchw(gs, int(v.act.shape[-1]), PLANES))
        tape.backward()
        grads = [None if tape.grads.get(name) is None else tape.grads[name].reshape(shape)
                 for name, shape in zip(ctx.names, ctx.shapes)]
        ctx.tape = None
        return (None, None, None, None, None, None, *grads)


ISLAND_PREFIXES = ("key_net.", "query_net.", "attention_net.")


def handshake_island(tape: Tape, q: Var, p: Dict[str, torch.Tensor], names, batch: int, agents: int):
    """The when2com handshake between the policy features ``q`` (on the tape) and the fuse kernel: KmGenerator key / query
    MLPs on ``features.view(-1, 4096)`` (When2com.py:415-430; for the seg model's 256 x 8 x 8 features that view makes FOUR
    rows per map and rows 0..A*B-1 are taken as the agents' keys / queries, When2Com_UNet.py:207-226), the
    32 -> 1024 linear layer on the queries, key . query scores and the softmax over the keys (When2com.py:374-412).
    Evaluated with torch autograd in fp32 (about 0.1 GFLOP per step).  Returns (coef [B, key, query] fp32 for
    v2x_warp_gated_fwd, holder): the fuse's backward leaves d(loss)/d(coef) in ``holder["dcoef"]`` and the closure
    registered here -- it must sit on the tape BEFORE the fuse -- pushes it back to ``q`` and to the island's parameters."""
    import torch.nn.functional as F
    lib = tape.lib
    island_names = [k for k in names if k.startswith(ISLAND_PREFIXES)]
    with torch.enable_grad():
        feat = ops.act_to_float(q.act).detach().requires_grad_(True)          # [N, 256, h, w] fp32, NCHW like the reference
        ip = {k: p[k].detach().to(torch.float32).requires_grad_(True) for k in island_names}
        flat = feat.reshape(-1, ip["key_net.fc.0.weight"].shape[1])           # KmGenerator: features_map.view(-1, n_feat)

        def mlp(pre):
            hid = F.relu(F.linear(flat, ip[pre + "fc.0.weight"], ip[pre + "fc.0.bias"]))
            hid = F.relu(F.linear(hid, ip[pre + "fc.2.weight"], ip[pre + "fc.2.bias"]))
            return F.linear(hid, ip[pre + "fc.4.weight"], ip[pre + "fc.4.bias"])
        keys, querys = mlp("key_net."), mlp("query_net.")
        key_mat = torch.stack([keys[batch * i: batch * (i + 1)] for i in range(agents)], 1)        # [B, A, key_size]
        query_mat = torch.stack([querys[batch * i: batch * (i + 1)] for i in range(agents)], 1)    # [B, A, query_size]
        query = F.linear(query_mat, ip["attention_net.linear.weight"], ip["attention_net.linear.bias"])
        attn = torch.softmax(torch.bmm(key_mat, query.transpose(2, 1)), dim=1)                      # [B, key, query]
    coef = attn.detach().contiguous()
    holder: Dict[str, torch.Tensor] = {}

    def island_bwd():
        dcoef = holder.get("dcoef")
        if dcoef is None:
            return
        leaves = [feat] + [ip[k] for k in island_names]
        grads = torch.autograd.grad(attn, leaves, grad_outputs=dcoef, allow_unused=True)
        if grads[0] is not None:
            if os.environ.get("V2X_TRAIN_DEBUG"):
                print("[when2com island] |d feat| max %.3e, x scale %.3e = %.3e; |dcoef| max %.3e" % (
                    float(grads[0].abs().max()), tape.scale, float(grads[0].abs().max()) * tape.scale, float(dcoef.abs().max())))
            g = (grads[0] * tape.scale).contiguous()
            q.add_grad(lib, ops.pack_input_nchw(g, int(g.shape[1]), PLANES))
        for k, gk in zip(island_names, grads[1:]):
            if gk is not None:
                tape._param_grad(k).add_(gk)
    tape.back.append(island_bwd)
    return coef, holder


class When2comTrainStep(torch.autograd.Function):
    """One train-mode forward of det When2com / who2com with ``training=True`` (When2com.py:150-332: image encoder, policy
    encoder PolicyNet4 = a second LidarEncoder + five conv/BN/ReLU, key / query MLPs, key-query attention with a softmax
    over the keys, ONE decoder pass on the attention-weighted fuse, heads) with its backward.
    The conv stacks, the fuse and their backward (99.99% of the step's FLOPs) run in the sm_100a kernels on the tape.  The
    handshake itself -- three-layer MLPs on [units, 4096] vectors, a 32 -> 1024 linear layer and a 5 x 5 softmax per scene,
    about 0.1 GFLOP per step -- is an "island" evaluated with torch autograd in fp32: its inputs (the policy features) come
    off the tape, its output (the attention coefficients) feeds v2x_warp_gated_fwd, and in the backward
    v2x_warp_gated_bwd's d(coef) is pushed through the island to the policy features and to the island's parameters.
    Inputs: (module, bevs, trans_matrices, num_agent_tensor, batch_size, *parameters in named_parameters() order)."""

    @staticmethod
    def forward(ctx, module, bevs, trans, nat, batch, *params):
        names = [k for k, _ in module.named_parameters()]
        p = {k: v.detach() for k, v in zip(names, params)}
        b = {k: v for k, v in module.named_buffers()}
        dev = bevs.device
        n = int(bevs.shape[0])
        agents = n // batch
        tape = Tape(p, b, dev)
        lib = tape.lib
        trans = trans.to(device=dev, dtype=torch.float64).contiguous()
        nat = nat.to(device=dev, dtype=torch.int64).contiguous()
        x_in = Var(ops.pack_input(bevs.reshape(n, 256, 256, -1).to(torch.float32).contiguous(), 16, PLANES), c_log=int(bevs.shape[-1]))
        x0, x1, x2, x3, x4 = backbone_encode(tape, "u_encoder.", x_in)
        # policy branch (PolicyNet4, When2com.py:335-359): its own encoder, then conv1..conv5 (strides 1, 1, 2, 1, 2)
        q = backbone_encode(tape, "query_key_net.lidar_encoder.", x_in)[4]
        for i, stride in enumerate((1, 1, 2, 1, 2), start=1):
            pre = "query_key_net.conv%d.cbr_unit." % i
            q = tape.cbr(pre + "0", pre + "1", [q], stride=stride)
        coef, holder = handshake_island(tape, q, p, names, batch, agents)
        fused = tape.gated_fuse(x3, coef, holder, trans, nat, batch, agents, warp_flag=int(module.warp_flag),
                                only_v2i=bool(module.only_v2i))
        x8 = backbone_decode(tape, "decoder.", x0, x1, x2, fused, x4)
        o_loc, o_cls = det_heads(tape, x8, n)
        ctx.tape, ctx.names, ctx.o_loc, ctx.o_cls = tape, names, o_loc, o_cls
        ctx.shapes = [v.shape for v in params]
        return o_loc.value.view(n, 256, 256, 6, 1, 6), o_cls.value.view(n, -1, 2)

    @staticmethod
    def backward(ctx, dloc, dcls):
        tape = ctx.tape
        tape.scale = choose_scale(*[u for u in (dloc, dcls) if u is not None])
        ctx.o_loc.upstream, ctx.o_cls.upstream = dloc, dcls
        tape.backward()
        grads = [None if tape.grads.get(name) is None else tape.grads[name].reshape(shape)
                 for name, shape in zip(ctx.names, ctx.shapes)]
        ctx.tape = None
        return (None, None, None, None, None, *grads)


# =====================================================================================================================
# segmentation models (CP/models/seg/SegModelBase.py:91-151, UNet.py:24-44, seg/V2VNet.py:25-92) in train mode
# =====================================================================================================================
def seg_double_conv(t: Tape, p: str, srcs, need_input_grad=None) -> Var:
    x = t.cbr(p + "0", p + "1", srcs, need_input_grad=need_input_grad)
    return t.cbr(p + "3", p + "4", [x])


def seg_encode(t: Tape, x_in: Var, pre: str = ""):
    x1 = seg_double_conv(t, pre + "inc.double_conv.", [x_in], need_input_grad=[False])
    x2 = seg_double_conv(t, pre + "down1.maxpool_conv.1.double_conv.", [t.maxpool2(x1)])
    x3 = seg_double_conv(t, pre + "down2.maxpool_conv.1.double_conv.", [t.maxpool2(x2)])
    x4 = seg_double_conv(t, pre + "down3.maxpool_conv.1.double_conv.", [t.maxpool2(x3)])
    if pre + "com_compresser.weight" in t.p:     # the model's own encoder only; PolicyNet4 has no compresser
        x4 = compress_pair(t, pre, x4)
    return x1, x2, x3, x4


def seg_decode(t: Tape, feat: Var, x1, x2, x3, n: int):
    """down4 / up1..4 (cat([skip, bilinear_up])) / outc -> fp32 NCHW logits."""
    x5 = seg_double_conv(t, "down4.maxpool_conv.1.double_conv.", [t.maxpool2(feat)])
    x = seg_double_conv(t, "up1.conv.double_conv.", [feat, t.upsample_bilinear2(x5)])
    x = seg_double_conv(t, "up2.conv.double_conv.", [x3, t.upsample_bilinear2(x)])
    x = seg_double_conv(t, "up3.conv.double_conv.", [x2, t.upsample_bilinear2(x)])
    x = seg_double_conv(t, "up4.conv.double_conv.", [x1, t.upsample_bilinear2(x)])
    n_cls = t.p["outc.conv.weight"].shape[0]
    logits = torch.empty((n, n_cls, x.act.shape[2], x.act.shape[3]), dtype=torch.float32, device=t.dev)
    return t.conv1x1_out("outc.conv", x, logits, nchw=True)


class SegTrainStep(torch.autograd.Function):
    """One train-mode forward of seg UNet (``fuse`` = None), seg V2VNet (``fuse`` = (trans, num_agent, batch, agents,
    only_v2i): one GNN round on the 512-channel layer-4 map, neighbour mean INCLUDING self) or seg Mean / Sum / Max fusion
    (``fuse`` = (..., kind)) with its backward, so the
    reference's SegModule.step (CP/utils/SegModule.py:45-120: loss -> backward -> optimizer) drives it unchanged.
    Inputs: (module, fuse, x [N,13,256,256], *parameters in named_parameters() order) -> logits [N, classes, 256, 256]."""

    @staticmethod
    def forward(ctx, module, fuse, x, *params):
        names = [k for k, _ in module.named_parameters()]
        p = {k: v.detach() for k, v in zip(names, params)}
        b = {k: v for k, v in module.named_buffers()}
        dev = x.device
        n = int(x.shape[0])
        tape = Tape(p, b, dev)
        x_in = Var(ops.pack_input_nchw(x.to(torch.float32).contiguous(), 16, PLANES), c_log=int(x.shape[1]))
        x1, x2, x3, x4 = seg_encode(tape, x_in)
        feat = x4
        if fuse is not None:
            trans, nat, batch, agents, only_v2i = fuse[:5]
            kind = fuse[5] if len(fuse) > 5 else "v2v"
            trans = trans.to(device=dev, dtype=torch.float64).contiguous()
            nat = nat.to(device=dev, dtype=torch.int64).contiguous()
            if kind == "v2v":
                mean = tape.warp_mean(x4, trans, nat, batch, agents, only_v2i=only_v2i, include_self=True)
                for _ in range(module.gnn_iter_num):
                    feat = tape.gru_round(feat, mean, x4, nat, batch, agents)
            elif kind == "when2com":
                # seg When2Com_UNet with training=True (When2Com_UNet.py:144-307): policy branch = its own inc / down1..3
                # + conv1..conv5 (strides 1, 1, 2, 1, 2) -> [N, 256, 8, 8]; handshake island; attention-weighted fuse of x4
                q = seg_encode(tape, x_in, "query_key_net.")[3]
                for i, stride in enumerate((1, 1, 2, 1, 2), start=1):
                    cpre = "query_key_net.conv%d.cbr_unit." % i
                    q = tape.cbr(cpre + "0", cpre + "1", [q], stride=stride)
                coef, holder = handshake_island(tape, q, p, names, batch, agents)
                feat = tape.gated_fuse(x4, coef, holder, trans, nat, batch, agents, warp_flag=int(fuse[6]), only_v2i=only_v2i)
            elif kind == "disco":   # seg DiscoNet (kd_flag False): trained per-pixel weights (see Tape.disco_fuse)
                feat = tape.disco_fuse(x4, trans, nat, batch, agents, only_v2i=only_v2i)
            elif kind == "agent":   # seg AgentWiseWeightedFusion: detached per-pair weights (see Tape.agent_weighted_fuse)
                feat = tape.agent_weighted_fuse(x4, trans, nat, batch, agents, only_v2i=only_v2i)
            elif kind == "cat":   # seg CatFusion (seg/CatFusion.py:8-34): mean, then the per-agent modulation layer
                mean = tape.warp_reduce(x4, trans, nat, batch, agents, "mean", only_v2i=only_v2i)
                feat = tape.cat_modulate(x4, mean, nat, batch, agents, conv="modulation_layer_3.conv1_1",
                                         bn="modulation_layer_3.bn1_1")
            else:   # seg MeanFusion / SumFusion / MaxFusion (seg/FusionBase.py:25-84): parameter-free fuse of the layer-4 maps
                feat = tape.warp_reduce(x4, trans, nat, batch, agents, kind, only_v2i=only_v2i)
        out = seg_decode(tape, feat, x1, x2, x3, n)
        ctx.tape, ctx.names, ctx.out = tape, names, out
        ctx.shapes = [v.shape for v in params]
        return out.value

    @staticmethod
    def backward(ctx, dlogits):
        tape = ctx.tape
        tape.scale = choose_scale(dlogits)
        ctx.out.upstream = dlogits
        tape.backward()
        k = "convgru.weight_ih_l0"
        if k in tape.grads:
            tape.grads[k] = torch.flip(tape.grads[k], (2,))
        grads = [None if tape.grads.get(name) is None else tape.grads[name].reshape(shape)
                 for name, shape in zip(ctx.names, ctx.shapes)]
        ctx.tape = None
        return (None, None, None, *grads)
