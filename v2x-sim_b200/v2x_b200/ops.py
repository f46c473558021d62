"""Torch-facing wrappers of the C ABI: tensors in, tensors out, kernels from libv2x_b200.so.

torch is plumbing here (device memory, streams); every FLOP of the path runs in the
hand-written sm_100a kernels.  An ``act`` is a 16-bit tensor ``[planes, N, H, W, C]``
(planes = 1: one bf16 plane; planes = 2: fp16 hi/lo split, see include/v2x_b200.h).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import EPI_ACT, EPI_F32_NCHW, EPI_F32_SPLIT, EPI_GRU, EPI_TAIL_F32_SPLIT, ConvParams, V2XError, check

BN_EPS = 1e-5  # nn.BatchNorm2d default (reference never overrides it)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def require_gpu():
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise V2XError("no CUDA device: the v2x_b200 path has no CPU fallback")
    if not lib.v2x_device_ok():
        raise V2XError("current device is not compute capability 10.x (sm_100a kernels only)")
    return lib


def act_dtype(planes):
    """Storage format follows the plane count: one bf16 plane, or fp16 hi + fp16 lo (include/v2x_b200.h)."""
    return torch.bfloat16 if planes == 1 else torch.float16


def empty_act(planes, n, h, w, c, device):
    return torch.empty((planes, n, h, w, c), dtype=act_dtype(planes), device=device)


FMT_BF16, FMT_F16X2, FMT_F16 = 1, 2, 3


def weight_fmt(planes, mmas):
    """(storage format, planes) of a packed conv operand: bf16 | fp16 hi/lo (3 MMA passes) | one fp16 plane (1 or 2)."""
    if planes == 1:
        return FMT_BF16, 1
    return (FMT_F16X2, 2) if mmas == 3 else (FMT_F16, 1)


def act_to_float(act: torch.Tensor) -> torch.Tensor:
    """act [P,N,H,W,C] -> fp32 NCHW through the CUDA export kernel."""
    lib = require_gpu()
    p, n, h, w, c = act.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=act.device)
    check(lib.v2x_act_to_nchw_f32(_ptr(act), _ptr(out), n, h, w, c, p, _stream()), "v2x_act_to_nchw_f32")
    return out


def pack_input(x: torch.Tensor, c_pad: int, planes: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 NHWC ``[..., H, W, C]`` (any leading dims) -> act ``[planes, N, H, W, c_pad]``."""
    lib = require_gpu()
    assert x.dtype == torch.float32 and x.is_contiguous() and x.is_cuda
    h, w, c = x.shape[-3:]
    n = x.numel() // (h * w * c)
    if out is None:
        out = empty_act(planes, n, h, w, c_pad, x.device)
    check(lib.v2x_pack_input(_ptr(x), _ptr(out), n * h * w, c, c_pad, planes, _stream()), "v2x_pack_input")
    return out


def pack_input_u8(x: torch.Tensor, c_pad: int, planes: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bool / uint8 NHWC occupancy ``[..., H, W, C]`` (non-zero = 1.0) -> act ``[planes, N, H, W, c_pad]``."""
    lib = require_gpu()
    assert x.dtype in (torch.uint8, torch.bool) and x.is_contiguous() and x.is_cuda
    h, w, c = x.shape[-3:]
    n = x.numel() // (h * w * c)
    if out is None:
        out = empty_act(planes, n, h, w, c_pad, x.device)
    check(lib.v2x_pack_input_u8(_ptr(x), _ptr(out), n * h * w, c, c_pad, planes, _stream()), "v2x_pack_input_u8")
    return out


def voxelize(idx: torch.Tensor, count: torch.Tensor, out: torch.Tensor, c: int, bad: torch.Tensor, rot90=True):
    """Sparse voxel rows (map, i0, i1, i2) int32 [capacity, 4] (``count`` valid, a device int32 scalar) -> the dense
    occupancy act ``out`` [planes, N, H, W, c_pad], incl. the dataset's np.rot90(., 3) (V2XSimDet.py:294-299)."""
    lib = require_gpu()
    assert idx.dtype == torch.int32 and idx.is_cuda and idx.is_contiguous() and idx.shape[1] == 4
    assert count.dtype == torch.int32 and count.is_cuda and bad.dtype == torch.int32 and bad.is_cuda
    planes, n, h, w, c_pad = out.shape
    check(lib.v2x_voxelize_fwd(_ptr(idx), _ptr(count), idx.shape[0], _ptr(out), n, h, w, c, c_pad, planes, int(rot90),
                               _ptr(bad), _stream()), "v2x_voxelize_fwd")
    return out


@dataclass
class PackedConv:
    """Packed operand of one fused conv launch (weights [PW, cout_pad, k_total] in the format of
    ``weight_fmt(planes, mmas)``, bias fp32).  ``planes`` = storage planes of the act tensors it is applied to,
    ``mmas`` = tensor-core passes per k-step (1 for planes == 1; 1..3 for planes == 2, see v2x_b200/precision.py)."""
    weights: torch.Tensor
    bias: torch.Tensor
    cins: List[int]          # padded channels per source
    taps: int
    stride: int
    cout: int
    cout_pad: int
    planes: int
    gru_bhn: Optional[torch.Tensor] = None
    keep: list = field(default_factory=list)
    gru_pre_act: bool = False   # weights carry 192 identity K columns for the pre-activation window (v2x_conv_params.gru_pre_act)
    tap_pack: bool = False      # weights are [P, 96, 3 * sum(cins)]: horizontal taps packed into N (v2x_conv_params.tap_pack)
    mmas: int = 0               # 0 = the format's default (1 for bf16, 3 for fp16 hi/lo)

    @property
    def k_total(self):
        return self.taps * sum(self.cins)


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


TAP_PACK = not os.environ.get("V2X_NO_TAP_PACK")   # A/B switch for the tap-packed 32-channel convs (csrc/conv_pack3.cu)
TAP_PACK_SPLIT = not os.environ.get("V2X_NO_TAP_PACK_SPLIT")   # ... for the shallow N = 32 layers in the fp16 hi/lo modes


def pack_conv_tap_packed(weight, bias, bn, *, cins: Sequence[int], planes=1, device=None, mmas=0) -> PackedConv:
    """3x3 stride-1 conv with 32 | cout in the tap-packed form of csrc/conv_pack3.cu: the operand is
    [planes, cout/32 * 96, 3 * sum(cins)] with row = (group, kw, co) and k = (source, kh, ci) -- built by handing the
    regular packer the weights rearranged as a "3-tap" conv with 3 * cout outputs (BN scale folded per row; rows
    (group, kw = 0, co) carry the folded bias)."""
    lib = require_gpu()
    device = device or weight.device
    cout, cin_total = weight.shape[0], weight.shape[1]
    assert cout % 32 == 0 and tuple(weight.shape[2:]) == (3, 3) and sum(cins) == cin_total
    g = cout // 32
    cin_pads = [((c + 15) // 16) * 16 for c in cins]
    from .transforms import tap_packed_rows
    w = tap_packed_rows(_f32(weight, device))   # [cout, ci, kh, kw] -> rows (g, kw, co) x [ci, kh]; proven in tests/test_transforms.py
    rep = lambda t: None if t is None else _f32(t, device).view(g, 1, 32).expand(g, 3, 32).reshape(-1).contiguous()  # noqa: E731
    b = rep(bias)
    bnp = [rep(t) for t in bn] if bn is not None else [None] * 4
    k_total = 3 * sum(cin_pads)
    mmas = (3 if planes == 2 else 1) if not mmas else mmas
    fmt, pw = weight_fmt(planes, mmas)
    dst = torch.zeros((pw, 3 * cout, k_total), dtype=act_dtype(planes), device=device)
    dst_bias = torch.zeros((3 * cout,), dtype=torch.float32, device=device)
    ci_lo, k_off = 0, 0
    for s, (c, cp) in enumerate(zip(cins, cin_pads)):
        check(lib.v2x_pack_conv_weights(_ptr(w), _ptr(b), _ptr(bnp[0]), _ptr(bnp[1]), _ptr(bnp[2]), _ptr(bnp[3]), BN_EPS,
                                        3 * cout, cin_total, 3, ci_lo, ci_lo + c, cp, 0, 0, _ptr(dst), _ptr(dst_bias),
                                        fmt, 3 * cout, k_total, 0, k_off, int(s == 0), _stream()),
              "v2x_pack_conv_weights(tap_pack)")
        ci_lo += c
        k_off += 3 * cp
    pc = PackedConv(dst, dst_bias, cin_pads, 9, 1, cout, 3 * cout, planes, mmas=mmas)
    pc.tap_pack = True
    return pc


def pack_conv(weight, bias, bn=None, *, cins: Sequence[int], cin_pads: Optional[Sequence[int]] = None, stride=1,
              planes=1, device=None, vflip=False, gru=False, cout_pad: Optional[int] = None,
              tap_pack: Optional[bool] = None, mmas: int = 0) -> PackedConv:
    """BN-fold + reorder + bf16 split of one conv's weights, on device.

    weight: OIHW (or OI111 for the 1x1x1 Conv3D) fp32; ``cins`` = logical channels of each concat
    source (sum == weight.shape[1]); bn = (gamma, beta, running_mean, running_var) or None.
    """
    lib = require_gpu()
    device = device or weight.device
    cout, cin_total = weight.shape[0], weight.shape[1]
    taps = int(weight.numel() // (cout * cin_total))
    assert taps in (1, 9) and sum(cins) == cin_total
    # tap packing pays where the layer is tensor-pipe bound, i.e. K is deep enough (conv8_1: 96 channels); the 13 / 32
    # channel layers are epilogue / HBM bound and keep the leaner N = 32 epilogue (measured, profiles/r01_v10)
    eligible = taps == 9 and stride == 1 and cout in (32, 64) and cin_pads is None and cout_pad is None and not vflip \
        and not gru and weight.dim() == 4
    # 64-output layers with deep K (conv7_1: 192 channels) additionally stop re-streaming their 221 KB of weights from
    # L2 for every M tile: each 32-channel group keeps its 110 KB operand resident
    mmas = (3 if planes == 2 else 1) if not mmas else mmas
    fmt, pw = weight_fmt(planes, mmas)
    deep = cin_total >= (64 if cout == 32 else 128) and 3 * cin_total * 96 * 2 * pw <= 112 * 1024  # + >= 3 stages
    # fp16 hi/lo storage: with 2-3 tensor-core passes per k-step even the shallow 32-output layers (conv_pre_*, conv8_2)
    # are tensor-pipe bound at N = 32 (40 cycles per MMA for 16 cycles of math), so they take the tap-packed form too
    # (measured r02: conv_pre_2 254 -> 193 us, conv8_2 258 -> 206 us; the 13-channel conv_pre_1 loses, K is too shallow)
    split_n32 = (planes == 2 and mmas >= 2 and cout == 32 and cin_total >= 32 and TAP_PACK_SPLIT
                 and 3 * cin_total * 96 * 2 * pw <= 112 * 1024)
    if eligible and (tap_pack is True or (tap_pack is None and TAP_PACK and (deep or split_n32))):
        return pack_conv_tap_packed(weight, bias, bn, cins=cins, planes=planes, device=device, mmas=mmas)
    cin_pads = list(cin_pads) if cin_pads is not None else [((c + 15) // 16) * 16 for c in cins]
    cout_pad = cout_pad or cout   # zero rows up to a multiple of the N tile (e.g. the 8-class seg logits)
    k_total = taps * sum(cin_pads)
    w = _f32(weight, device).reshape(cout, cin_total, taps)
    b = _f32(bias, device) if bias is not None else None
    bnp = [_f32(t, device) for t in bn] if bn is not None else [None] * 4
    dst = torch.zeros((pw, cout_pad, k_total), dtype=act_dtype(planes), device=device)
    dst_bias = torch.zeros((cout_pad,), dtype=torch.float32, device=device)
    ci_lo, k_off = 0, 0
    for s, (c, cp) in enumerate(zip(cins, cin_pads)):
        check(lib.v2x_pack_conv_weights(_ptr(w), _ptr(b), _ptr(bnp[0]), _ptr(bnp[1]), _ptr(bnp[2]), _ptr(bnp[3]),
                                        BN_EPS, cout, cin_total, taps, ci_lo, ci_lo + c, cp, int(vflip),
                                        3 if gru else 0, _ptr(dst), _ptr(dst_bias), fmt, cout_pad, k_total, 0, k_off,
                                        int(s == 0 and not gru), _stream()), "v2x_pack_conv_weights")
        ci_lo += c
        k_off += taps * cp
    return PackedConv(dst, dst_bias, cin_pads, taps, stride, cout, cout_pad, planes, mmas=mmas)


def pack_gru(w_ih, b_ih, b_hh, *, planes=1, device=None, vflip=True, mmas=0) -> PackedConv:
    """Zero-hidden ConvGRU operand: W_ih rows gate-interleaved per 64 channels, filter rows mirrored so
    the GRU runs in the un-flipped domain.  W_hh is never needed (hidden == 0)."""
    lib = require_gpu()
    device = device or w_ih.device
    c = w_ih.shape[0] // 3
    pc = pack_conv(w_ih, None, None, cins=[c, w_ih.shape[1] - c], planes=planes, device=device, vflip=vflip, gru=True,
                   mmas=mmas)
    bhn = torch.empty((c,), dtype=torch.float32, device=device)
    bi, bh = _f32(b_ih, device), _f32(b_hh, device)
    check(lib.v2x_pack_gru_bias(_ptr(bi), _ptr(bh), c, _ptr(pc.bias), _ptr(bhn), _stream()), "v2x_pack_gru_bias")
    pc.gru_bhn = bhn
    return pc


def pack_gru_split(w_ih, b_ih, b_hh, *, planes=1, device=None, vflip=True, pre_act=False, mmas=0):
    """The zero-hidden ConvGRU split in two operands: ``gru_m`` convolves the round-invariant neighbour mean
    (W_ih[:, C:], carries the combined bias) once per frame into gate pre-activations; ``gru_h`` convolves the agent's
    own state (W_ih[:, :C]) in every GNN round and adds them.  Same arithmetic as conv(cat([h, mean])) up to fp32
    summation order; one third fewer FLOPs over three rounds.

    pre_act=False: pre-activations are fp32 (an EPI_F32_SPLIT launch of ``gru_m``) and are added in ``gru_h``'s gate
    epilogue (``gru_add``).  pre_act=True: they are a bf16 act tensor (EPI_ACT launch, relu=False) that ``gru_h`` takes
    as a second source: 192 identity columns appended to its packed weights make every N tile accumulate its own window
    on the tensor core, so the epilogue issues no global loads (``v2x_conv_params.gru_pre_act``)."""
    lib = require_gpu()
    device = device or w_ih.device
    c = w_ih.shape[0] // 3
    gru_h = pack_conv(w_ih[:, :c], None, None, cins=[c], planes=planes, device=device, vflip=vflip, gru=True, mmas=mmas)
    gru_m = pack_conv(w_ih[:, c:], None, None, cins=[w_ih.shape[1] - c], planes=planes, device=device, vflip=vflip,
                      gru=True, mmas=mmas)
    bhn = torch.empty((c,), dtype=torch.float32, device=device)
    bi, bh = _f32(b_ih, device), _f32(b_hh, device)
    check(lib.v2x_pack_gru_bias(_ptr(bi), _ptr(bh), c, _ptr(gru_m.bias), _ptr(bhn), _stream()), "v2x_pack_gru_bias")
    gru_h.gru_bhn = bhn     # gru_h.bias stays zero: the bias rides in gru_m's output
    if pre_act:
        assert planes == 1, "gru_pre_act is a bf16 (planes == 1) feature"
        eye = torch.zeros((planes, 3 * c, 192), dtype=torch.bfloat16, device=device)
        eye[0] = torch.eye(192, dtype=torch.bfloat16, device=device).repeat(3 * c // 192, 1)   # row n -> column n % 192
        gru_h.weights = torch.cat([gru_h.weights, eye], dim=2).contiguous()
        gru_h.cins = [c, 192]
        gru_h.gru_pre_act = True
    return gru_h, gru_m


def pack_heads(cls1_w, cls1_b, cls_bn, reg1_w, reg1_b, reg_bn, cls2_w, cls2_b, reg2_w, reg2_b, *, planes=1,
               device=None, mmas=0):
    """Detection heads (DetModelBase.py:268-351) as two launches: one 32->64 3x3 conv (cls.conv1|reg.0 rows
    stacked, BN folded, ReLU) and one block-diagonal 64->48 1x1 conv (cls.conv2 on channels 0..31, reg.3 on 32..63)."""
    lib = require_gpu()
    device = device or cls1_w.device
    c = cls1_w.shape[1]
    n1 = cls1_w.shape[0] + reg1_w.shape[0]
    mmas = (3 if planes == 2 else 1) if not mmas else mmas
    fmt1, pw1 = weight_fmt(planes, mmas)
    fmt2, pw2 = weight_fmt(planes, 3 if planes == 2 else 1)   # the 1x1 tail always runs at full precision (0.5% of the FLOPs)
    w1 = torch.zeros((pw1, n1, 9 * c), dtype=act_dtype(planes), device=device)
    b1 = torch.zeros((n1,), dtype=torch.float32, device=device)
    row = 0
    for w, b, bn in ((cls1_w, cls1_b, cls_bn), (reg1_w, reg1_b, reg_bn)):
        wf = _f32(w, device).reshape(w.shape[0], c, 9)
        bnp = [_f32(t, device) for t in bn]
        check(lib.v2x_pack_conv_weights(_ptr(wf), _ptr(_f32(b, device)), _ptr(bnp[0]), _ptr(bnp[1]), _ptr(bnp[2]),
                                        _ptr(bnp[3]), BN_EPS, w.shape[0], c, 9, 0, c, c, 0, 0, _ptr(w1), _ptr(b1),
                                        fmt1, n1, 9 * c, row, 0, 1, _stream()), "v2x_pack_conv_weights(head1)")
        row += w.shape[0]
    head1 = PackedConv(w1, b1, [c], 9, 1, n1, n1, planes, mmas=mmas)
    n2 = cls2_w.shape[0] + reg2_w.shape[0]
    w2 = torch.zeros((pw2, n2, n1), dtype=act_dtype(planes), device=device)
    b2 = torch.zeros((n2,), dtype=torch.float32, device=device)
    row, koff = 0, 0
    for w, b in ((cls2_w, cls2_b), (reg2_w, reg2_b)):
        ci = w.shape[1]
        wf = _f32(w, device).reshape(w.shape[0], ci, 1)
        check(lib.v2x_pack_conv_weights(_ptr(wf), _ptr(_f32(b, device)), None, None, None, None, BN_EPS, w.shape[0],
                                        ci, 1, 0, ci, ci, 0, 0, _ptr(w2), _ptr(b2), fmt2, n2, n1, row, koff, 1,
                                        _stream()), "v2x_pack_conv_weights(head2)")
        row += w.shape[0]
        koff += ci
    head2 = PackedConv(w2, b2, [n1], 1, 1, n2, n2, planes)
    return head1, head2, cls2_w.shape[0]


def pick_block_n(cout: int, m_tiles: int, sms: int = 148) -> int:
    """N tile of the persistent conv kernel: grid = (sms // n_tiles, n_tiles) CTAs, each walking
    ceil(m_tiles / ctas_x) M tiles.  Cost per 16-deep k-step of a 128 x BN tile ~ max(tensor math BN/2,
    shared-memory operand reads (128 + BN) / 4) cycles; pick the BN minimising rounds * cost."""
    cands = [bn for bn in (256, 128, 64, 32) if cout % bn == 0]
    if not cands:
        return 48 if cout % 48 == 0 else 32
    best, best_cost = None, None
    for bn in cands:
        n_tiles = cout // bn
        ctas_x = max(1, min(m_tiles, sms // n_tiles))
        rounds = -(-m_tiles // ctas_x)
        cost = rounds * max(bn / 2.0, (128 + bn) / 4.0)
        if best_cost is None or cost < best_cost:
            best, best_cost = bn, cost
    return best


class ConvLaunch:
    """A fully-bound v2x_conv_fwd call (parameters resolved once, replayed every step)."""

    def __init__(self, pc: PackedConv, srcs: Sequence[torch.Tensor], *, epilogue=EPI_ACT, relu=True, upsample2x=False,
                 out0: torch.Tensor = None, out1: torch.Tensor = None, out_c_off=0, split=0, block_n=None,
                 passthrough=None, num_agent=None, batch=0, agents=0, map_offset=0, gru_add=None, crosscheck=False,
                 tail: Optional[PackedConv] = None):
        self.lib = require_gpu()
        planes, n, h_in, w_in, _ = srcs[0].shape
        assert planes == pc.planes
        for i, (s, cp) in enumerate(zip(srcs, pc.cins)):
            want = pc.cout if (pc.gru_pre_act and i == 1) else cp    # the pre-activation source spans all cout channels
            assert s.shape[-1] == want and s.is_contiguous() and s.dtype == act_dtype(planes), (s.shape, s.dtype, want)
        assert not pc.gru_pre_act or (epilogue == EPI_GRU and len(srcs) == 2)
        h_out, w_out = h_in // pc.stride, w_in // pc.stride
        p = ConvParams()
        p.src[0] = srcs[0].data_ptr()
        p.src[1] = srcs[1].data_ptr() if len(srcs) > 1 else None
        p.cin[0] = pc.cins[0]
        p.cin[1] = pc.cins[1] if len(srcs) > 1 else 0
        p.n_maps, p.h_out, p.w_out, p.stride, p.taps, p.planes = n, h_out, w_out, pc.stride, pc.taps, planes
        p.weights, p.bias = pc.weights.data_ptr(), pc.bias.data_ptr()
        p.cout, p.cout_pad = pc.cout, pc.cout_pad
        m_tiles = n * (-(-h_out // 8)) * (-(-w_out // 16))
        if epilogue == EPI_GRU:
            p.block_n = 192
        else:
            p.block_n = block_n or pick_block_n(pc.cout, m_tiles)
        p.epilogue, p.relu, p.upsample2x = epilogue, int(relu), int(upsample2x)
        p.out0 = out0.data_ptr()
        p.out1 = out1.data_ptr() if out1 is not None else None
        p.out_c_total = out0.shape[-1] if epilogue in (EPI_ACT, EPI_GRU) else 0
        p.out_c_off, p.split = out_c_off, split
        p.gru_bhn = pc.gru_bhn.data_ptr() if pc.gru_bhn is not None else None
        p.passthrough = passthrough.data_ptr() if passthrough is not None else None
        p.num_agent = num_agent.data_ptr() if num_agent is not None else None
        p.batch, p.agents, p.map_offset = batch, agents, map_offset
        p.gru_add = gru_add.data_ptr() if gru_add is not None else None
        p.gru_pre_act = int(pc.gru_pre_act)
        p.mmas = pc.mmas
        if pc.tap_pack:
            assert epilogue == EPI_ACT and not upsample2x and h_out % 8 == 0, "tap-packed convs: plain EPI_ACT, 8 | H"
            p.tap_pack, p.block_n = 1, 96
        if tail is not None:   # fused 1x1 conv on the ReLU output (EPI_TAIL_F32_SPLIT)
            assert epilogue == EPI_TAIL_F32_SPLIT and tail.taps == 1 and tail.cins == [pc.cout] and tail.planes == planes
            p.tail_weights, p.tail_bias = tail.weights.data_ptr(), tail.bias.data_ptr()
            p.tail_cout, p.tail_cout_pad = tail.cout, tail.cout_pad
        self.p = p
        self.keep = (pc, list(srcs), out0, out1, passthrough, num_agent, gru_add, tail)
        if crosscheck and pc.tap_pack:
            raise V2XError("the CUDA-core cross-check kernel takes the regular operand layout: pack with tap_pack=False")
        self.fn = self.lib.v2x_conv_fwd_crosscheck if crosscheck else self.lib.v2x_conv_fwd
        k_eff = pc.taps * pc.cins[0] + 192 if pc.gru_pre_act else pc.taps * sum(pc.cins)
        self.mma_passes = (pc.mmas or 3) if planes == 2 else 1    # tensor-core passes per k-step (bench.py's pipe FLOPs)
        self.flops = 2.0 * n * h_out * w_out * pc.cout * k_eff
        if tail is not None:
            self.flops += 2.0 * n * h_out * w_out * tail.cout * pc.cout

    def __call__(self):
        check(self.fn(C.byref(self.p), _stream()), "v2x_conv_fwd")

    @property
    def label(self):
        """Geometry tag for profiles: e.g. "96>32 k3 s1 256x256 n40 N96 x2" (x = tensor-core passes)."""
        p = self.p
        cin = "%d" % p.cin[0] + ("+%d" % p.cin[1] if p.cin[1] else "")
        epi = {EPI_ACT: "", EPI_F32_SPLIT: " f32", EPI_GRU: " gru", EPI_F32_NCHW: " nchw", EPI_TAIL_F32_SPLIT: " +tail"}[p.epilogue]
        return "%s>%d k%d s%d %dx%d n%d N%d x%d%s%s" % (cin, p.cout, 3 if p.taps == 9 else 1, p.stride, p.h_out, p.w_out,
                                                        p.n_maps, p.block_n, self.mma_passes, " up2" if p.upsample2x else "", epi)


def conv(pc: PackedConv, srcs, *, relu=True, upsample2x=False, out=None, block_n=None, crosscheck=False):
    """One-off fused conv + (folded BN) + ReLU -> act."""
    planes, n, h_in, w_in, _ = srcs[0].shape
    up = 2 if upsample2x else 1
    if out is None:
        out = empty_act(planes, n, h_in // pc.stride * up, w_in // pc.stride * up, pc.cout, srcs[0].device)
    ConvLaunch(pc, srcs, relu=relu, upsample2x=upsample2x, out0=out, block_n=block_n, crosscheck=crosscheck)()
    return out


def warp_mean(x: torch.Tensor, trans: torch.Tensor, num_agent: torch.Tensor, batch: int, agents: int, *,
              include_self=False, only_v2i=False, out=None, unit_offset=0, unit_count=0) -> torch.Tensor:
    """Cross-agent bilinear warp + neighbour mean of agent-major maps ``x`` [P, A*B, H, W, C].
    ``unit_offset/unit_count`` restrict the computed targets to a slice of units (unit-sharded plans)."""
    lib = require_gpu()
    planes, n, h, w, c = x.shape
    assert n == batch * agents
    assert trans.dtype == torch.float64 and trans.is_cuda and trans.is_contiguous()
    assert num_agent.dtype == torch.int64 and num_agent.is_cuda and num_agent.is_contiguous()
    if out is None:
        out = torch.empty_like(x) if unit_count <= 0 else empty_act(planes, unit_count, h, w, c, x.device)
    check(lib.v2x_warp_mean_fwd(_ptr(x), _ptr(out), _ptr(trans), _ptr(num_agent), batch, agents, h, w, c, planes,
                                int(include_self), int(only_v2i), unit_offset, unit_count, _stream()), "v2x_warp_mean_fwd")
    return out


# ---- when2com / who2com ----------------------------------------------------------------------
def linear(x, w, b, *, relu=False, out=None, act_input=False, rows=None):
    """fp32 linear layer through v2x_linear_fwd.  ``act_input``: x is an act [P, maps, H, W, C] flattened in NCHW
    order and viewed as rows of ``in_f`` values (KmGenerator's ``view(-1, n_feat)``; when H*W*C is a multiple of in_f
    each map spans several rows -- the seg quirk); ``rows`` = how many leading rows to compute.
    Otherwise x is fp32 [rows, in_f]."""
    lib = require_gpu()
    out_f, in_f = w.shape
    maps, split = 0, 1
    if act_input:
        planes, maps, h, wd, c = x.shape
        assert (h * wd * c) % in_f == 0
        split = h * wd * c // in_f
        mode, hw = 1, h * wd
        rows = maps if rows is None else rows
    else:
        rows, planes, hw, c, mode = x.shape[0], 1, 0, 0, 0
        assert x.dtype == torch.float32 and x.shape[1] == in_f and x.is_contiguous()
    if out is None:
        out = torch.empty((rows, out_f), dtype=torch.float32, device=w.device)
    check(lib.v2x_linear_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(out), rows, in_f, out_f, int(relu), mode, hw, c, planes,
                             split, maps, _stream()), "v2x_linear_fwd")
    return out


GATE_MODES = {"softmax": 0, "activated": 1, "argmax_test": 2}


def attn_scores(keys, querys, w, bw, batch, agents, gate_mode, attn=None, coef=None):
    lib = require_gpu()
    dev = keys.device
    if attn is None:
        attn = torch.empty((batch, agents, agents), dtype=torch.float32, device=dev)
    if coef is None:
        coef = torch.empty((batch, agents, agents), dtype=torch.float32, device=dev)
    check(lib.v2x_attn_scores_fwd(_ptr(keys), _ptr(querys), _ptr(w), _ptr(bw), _ptr(attn), _ptr(coef), batch, agents,
                                  keys.shape[1], querys.shape[1], gate_mode, _stream()), "v2x_attn_scores_fwd")
    return attn, coef


def warp_gated(x, trans, num_agent, coef, batch, agents, *, warp_flag=1, only_v2i=False, out=None, unit_offset=0,
               unit_count=0, x_unit_offset=0):
    """when2com fuse: out[b,q] = sum_k coef[b,k,q] * val[b,k,q] without materialising val_mat.
    ``unit_offset/unit_count`` restrict the computed targets to a slice of the agent-major units (unit-sharded plans);
    ``x`` then holds units [x_unit_offset, x_unit_offset + x.shape[1])."""
    lib = require_gpu()
    planes, n, h, w, c = x.shape
    full = unit_count <= 0
    assert (n == batch * agents if full else True) and coef.dtype == torch.float32 and coef.is_contiguous()
    if out is None:
        out = torch.empty_like(x) if full else empty_act(planes, unit_count, h, w, c, x.device)
    check(lib.v2x_warp_gated_fwd(_ptr(x), _ptr(out), _ptr(trans), _ptr(num_agent), _ptr(coef), batch, agents, h, w, c,
                                 planes, int(warp_flag), int(only_v2i), unit_offset, 0 if full else unit_count,
                                 0 if full else x_unit_offset, 0 if full else n, _stream()), "v2x_warp_gated_fwd")
    return out


# ---- segmentation UNet pieces ----------------------------------------------------------------
def pack_input_nchw(x: torch.Tensor, c_pad: int, planes: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 NCHW [N, C, H, W] -> act [planes, N, H, W, c_pad]."""
    lib = require_gpu()
    assert x.dtype == torch.float32 and x.is_contiguous() and x.is_cuda and x.dim() == 4
    n, c, h, w = x.shape
    if out is None:
        out = empty_act(planes, n, h, w, c_pad, x.device)
    check(lib.v2x_pack_input_nchw(_ptr(x), _ptr(out), n, c, h, w, c_pad, planes, _stream()), "v2x_pack_input_nchw")
    return out


def maxpool2(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = require_gpu()
    planes, n, h, w, c = x.shape
    if out is None:
        out = empty_act(planes, n, h // 2, w // 2, c, x.device)
    check(lib.v2x_maxpool2_fwd(_ptr(x), _ptr(out), n, h // 2, w // 2, c, planes, _stream()), "v2x_maxpool2_fwd")
    return out


def upsample_bilinear2(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = require_gpu()
    planes, n, h, w, c = x.shape
    if out is None:
        out = empty_act(planes, n, 2 * h, 2 * w, c, x.device)
    check(lib.v2x_upsample_bilinear2_fwd(_ptr(x), _ptr(out), n, h, w, c, planes, _stream()), "v2x_upsample_bilinear2_fwd")
    return out


# ---- intermediate-fusion baselines (FusionBase family) ---------------------------------------
REDUCE_MODES = {"mean": 0, "sum": 1, "max": 2}


def warp_reduce(x, trans, num_agent, batch, agents, mode, *, only_v2i=False, out=None):
    """out[b,i] = mean/sum/max over {x[b,i]} U {warp_{j->i} x[b,j]} (Mean/Sum/MaxFusion.fusion)."""
    lib = require_gpu()
    planes, n, h, w, c = x.shape
    assert n == batch * agents and x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    check(lib.v2x_warp_reduce_fwd(_ptr(x), _ptr(out), _ptr(trans), _ptr(num_agent), batch, agents, h, w, c, planes,
                                  REDUCE_MODES[mode], int(only_v2i), _stream()), "v2x_warp_reduce_fwd")
    return out


def fold_bn_1x1(w, b, bn, device):
    """BN(eval)-folded fp32 matrix / bias of a 1x1 conv: (W * s, (b - mean) * s + beta); bn may be None."""
    w = _f32(w, device).reshape(w.shape[0], -1)
    b = _f32(b, device)
    if bn is None:
        return w.contiguous(), b.contiguous()
    gamma, beta, mean, var = [_f32(t, device) for t in bn]
    s = gamma / torch.sqrt(var + BN_EPS)
    return (w * s[:, None]).contiguous(), ((b - mean) * s + beta).contiguous()


def pack_pair_conv1(w, b, bn, *, planes=1, device=None) -> PackedConv:
    """First layer of the pair-weight nets (conv1_1 over cat[tg, nb], 2C -> 128, + BN) as ONE C -> 256 1x1 operand:
    rows [0,128) = the tg half (carries the folded bias), rows [128,256) = the neighbour half (zero bias).  conv1_1 is
    linear and pointwise, hence commutes with the bilinear warp (see v2x_pair_score_fwd)."""
    device = device or w.device
    c = w.shape[1] // 2
    gamma, beta, mean, var = bn
    w2 = torch.cat([w[:, :c], w[:, c:]], 0)
    b2 = torch.cat([b, mean])                       # (b - mean) == 0 for the neighbour half
    bn2 = (torch.cat([gamma, gamma]), torch.cat([beta, torch.zeros_like(beta)]), torch.cat([mean, mean]),
           torch.cat([var, var]))
    return pack_conv(w2, b2, bn2, cins=[c], planes=planes, device=device)


def pair_score(q, trans, num_agent, batch, agents, mlp, *, only_v2i=False, out=None):
    """scores[b,i,k,p] of the DiscoNet / AgentWise weight nets; ``mlp`` = (w2,b2,w3,b3,w4,b4) folded fp32."""
    lib = require_gpu()
    planes, n, h, w, c = q.shape
    assert n == batch * agents and c == 256 and q.is_contiguous()
    w2, b2, w3, b3, w4, b4 = mlp
    assert tuple(w2.shape) == (32, 128) and tuple(w3.shape) == (8, 32) and w4.numel() == 8 and b4.numel() == 1
    if out is None:
        out = torch.zeros((batch, agents, agents, h * w), dtype=torch.float32, device=q.device)
    check(lib.v2x_pair_score_fwd(_ptr(q), _ptr(out), _ptr(trans), _ptr(num_agent), _ptr(w2), _ptr(b2), _ptr(w3),
                                 _ptr(b3), _ptr(w4), _ptr(b4), batch, agents, h, w, planes, int(only_v2i), _stream()),
          "v2x_pair_score_fwd")
    return out


def agent_softmax(scores, w5f, b5, num_agent, *, out=None):
    lib = require_gpu()
    batch, agents, _, hw = scores.shape
    assert w5f.numel() == hw and w5f.is_contiguous()
    if out is None:
        out = torch.empty((batch, agents, agents), dtype=torch.float32, device=scores.device)
    check(lib.v2x_agent_softmax_fwd(_ptr(scores), _ptr(w5f), _ptr(b5), _ptr(num_agent), _ptr(out), batch, agents, hw,
                                    _stream()), "v2x_agent_softmax_fwd")
    return out


def warp_weighted(x, trans, num_agent, coef, batch, agents, *, per_pixel, only_v2i=False, out=None):
    """out[b,i] = sum_k c[b,i,k] * warp_{k->i}(x[b,k]); per_pixel: c = softmax_k of scores [B,A,A,H*W]."""
    lib = require_gpu()
    planes, n, h, w, c = x.shape
    assert n == batch * agents and coef.dtype == torch.float32 and coef.is_contiguous()
    assert coef.numel() == batch * agents * agents * (h * w if per_pixel else 1)
    if out is None:
        out = torch.empty_like(x)
    check(lib.v2x_warp_weighted_fwd(_ptr(x), _ptr(out), _ptr(trans), _ptr(num_agent), _ptr(coef), int(per_pixel), batch,
                                    agents, h, w, c, planes, int(only_v2i), _stream()), "v2x_warp_weighted_fwd")
    return out


def restore_absent(x, out, num_agent, batch, agents):
    """out[unit] = x[unit] for agent slots absent from their scene."""
    lib = require_gpu()
    planes, n, h, w, c = x.shape
    assert out.shape == x.shape and n == batch * agents
    check(lib.v2x_restore_absent_fwd(_ptr(x), _ptr(out), _ptr(num_agent), batch, agents, h * w * c, planes, _stream()),
          "v2x_restore_absent_fwd")
    return out
