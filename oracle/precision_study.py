"""TEST INFRASTRUCTURE ONLY -- CPU emulation of the operand precisions the sm_100a conv kernels can run in.

Every conv of the restated V2VNet det forward is re-run with its activations / weights rounded the way a candidate
kernel mode stores them (fp32 accumulation, like the TMEM accumulators), and the end-to-end error against the fp32
oracle is reported in the metric of tests/test_gpu_nets.py (max-abs error / max-abs value).  This is what decided
which modes were built; it costs no GPU time.

  python -m oracle.precision_study            # table of modes
"""
from __future__ import annotations

import sys

import torch
import torch.nn.functional as F

from . import restate, synth

_real_conv2d = F.conv2d


def q_bf16(x):
    return x.bfloat16().float()


def q_fp16(x):
    return x.half().float()


def q_split(q):
    def f(x):
        hi = q(x)
        return hi + q(x - hi)
    return f


def ident(x):
    return x


QUANT = {"f32": ident, "bf16": q_bf16, "fp16": q_fp16, "bf16x2": q_split(q_bf16), "fp16x2": q_split(q_fp16)}

# candidate kernel modes: (activation storage, weight storage, MMAs per k-step)
MODES = {
    "bf16":          ("bf16", "bf16", 1),
    "fp16":          ("fp16", "fp16", 1),
    "bf16x3":        ("bf16x2", "bf16x2", 3),
    "fp16a2":        ("fp16x2", "fp16", 2),     # activations hi/lo, weights single fp16
    "fp16w2":        ("fp16", "fp16x2", 2),     # activations single fp16, weights hi/lo
    "fp16x3":        ("fp16x2", "fp16x2", 3),
}


class emulate:
    """with emulate(policy): every F.conv2d inside oracle.restate rounds (x, w) per ``policy(x, w) -> (qa, qw)``."""

    def __init__(self, policy):
        self.policy = policy

    def __enter__(self):
        pol = self.policy

        def conv2d(x, w, b=None, *a, **k):
            qa, qw = pol(x, w)
            return _real_conv2d(QUANT[qa](x), QUANT[qw](w), b, *a, **k)
        restate.F.conv2d = conv2d
        return self

    def __exit__(self, *exc):
        restate.F.conv2d = _real_conv2d
        return False


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def heavy(x, w):
    """The tensor-bound launches of the V2VNet det step: K = 9 * cin >= 1728 (GRU, conv5_1 .. conv8_1 are cat inputs)."""
    cin = w.shape[1]
    return cin * w.shape[2] * w.shape[3] >= 1700 and cin not in (256, 512) or w.shape[0] == 768


def run(seed=0, batch=1, agents=5):
    sd = synth.v2vnet_det_state(seed)
    bevs, trans, nat = synth.make_scene(batch, agents, seed)
    with torch.no_grad():
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=agents, gnn_iter=3)
    rows = []
    policies = {name: (lambda x, w, m=m: (m[0], m[1])) for name, m in MODES.items()}
    # mixed: tensor-bound layers in a 2-MMA mode, everything else at full split precision
    for light, hv in (("fp16x3", "fp16w2"), ("fp16x3", "fp16a2"), ("fp16w2", "fp16x3"), ("fp16", "fp16x3"),
                      ("fp16x3", "fp16")):
        policies["light=%s heavy=%s" % (light, hv)] = (
            lambda x, w, l=MODES[light], h=MODES[hv]: (h[0], h[1]) if heavy(x, w) else (l[0], l[1]))
    for name, pol in policies.items():
        with torch.no_grad(), emulate(pol):
            out = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=agents, gnn_iter=3)
        flips = int((out["cls"].argmax(-1) != ref["cls"].argmax(-1)).sum())
        rows.append((name, rel_err(out["loc"], ref["loc"]), rel_err(out["cls"], ref["cls"]), flips))
        print("%-34s loc %.2e  cls %.2e  argmax flips %d / %d" % (rows[-1] + (ref["cls"].numel() // 2,)), flush=True)
    return rows


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "fp8"):
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 0)


def per_layer(seed=0, batch=1, agents=5, cheap=("fp16a2", "fp16w2", "fp16")):
    """Error contribution of each conv: everything at fp16x3 except ONE layer in a cheaper mode."""
    sd = synth.v2vnet_det_state(seed)
    bevs, trans, nat = synth.make_scene(batch, agents, seed)
    names = {v.data_ptr(): k for k, v in sd.items() if v.dim() >= 4}
    with torch.no_grad():
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=agents, gnn_iter=3)
    used = []

    def name_of(w):
        return names.get(w.data_ptr()) or names.get(w._base.data_ptr() if w._base is not None else 0, "?")

    def probe(x, w):
        if name_of(w) not in used:
            used.append(name_of(w))
        return ("f32", "f32")
    with torch.no_grad(), emulate(probe):
        restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=agents, gnn_iter=3)
    res = {}
    for layer in used:
        for m in cheap:
            def pol(x, w, layer=layer, m=m):
                return MODES[m][:2] if name_of(w) == layer else MODES["fp16x3"][:2]
            with torch.no_grad(), emulate(pol):
                out = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=agents, gnn_iter=3)
            res[(layer, m)] = max(rel_err(out["loc"], ref["loc"]), rel_err(out["cls"], ref["cls"]))
        print("%-42s " % layer + "  ".join("%s %.2e" % (m, res[(layer, m)]) for m in cheap), flush=True)
    return res


# ---------------------------------------------------------------------------------------------------------------------
# Study for the NEXT kernel mode: keep the main pass in fp16 and run the two correction
# passes of fp16x3 -- hi * w_lo and a_lo * w_hi, each 2^-11 of the product -- as FP8 (e5m2) tensor-core passes, which issue
# at twice the fp16 rate on sm_100a (kind::f8f6f4): 1 + 1/2 + 1/2 = 2 pass-equivalents instead of 3.
# e5m2 spans 2^-16 .. 2^15 (31 binades, 2 mantissa bits): with a static power-of-two split per layer (the activation side
# scaled down by 2^s, the weight-lo side up by 2^s, so the product needs no rescaling in the accumulator) both operands of
# each correction pass fit.  `python -m oracle.precision_study fp8` emulates exactly that on the CPU.
# ---------------------------------------------------------------------------------------------------------------------
def q_e5m2(x, shift):
    """x -> e5m2(x * 2^shift) * 2^-shift (round to nearest even, saturating to the largest finite e5m2)."""
    s = float(2.0 ** shift)
    y = (x * s).clamp(-57344.0, 57344.0)
    return y.to(torch.float8_e5m2).float() / s


def fp8_corrected_conv(x, w, b, args, kwargs, shift_w=8, shift_a=8):
    x_hi, w_hi = q_fp16(x), q_fp16(w)
    x_lo, w_lo = q_fp16(x - x_hi), q_fp16(w - w_hi)
    main = _real_conv2d(x_hi, w_hi, b, *args, **kwargs)
    c1 = _real_conv2d(q_e5m2(x_hi, -shift_w), q_e5m2(w_lo, shift_w), None, *args, **kwargs)      # hi * w_lo
    c2 = _real_conv2d(q_e5m2(x_lo, shift_a), q_e5m2(w_hi, -shift_a), None, *args, **kwargs)      # a_lo * w_hi
    return main + c1 + c2


class emulate_fp8:
    """Every conv as fp16 main pass + two e5m2 correction passes, except where ``keep(x, w)`` names another MODES entry."""

    def __init__(self, keep=None, shift_w=8, shift_a=8):
        self.keep, self.sw, self.sa = keep, shift_w, shift_a

    def __enter__(self):
        def conv2d(x, w, b=None, *a, **k):
            m = self.keep(x, w) if self.keep else None
            if m is not None:
                return _real_conv2d(QUANT[MODES[m][0]](x), QUANT[MODES[m][1]](w), b, *a, **k)
            return fp8_corrected_conv(x, w, b, a, k, self.sw, self.sa)
        restate.F.conv2d = conv2d
        return self

    def __exit__(self, *exc):
        restate.F.conv2d = _real_conv2d
        return False


def run_fp8(seed=0):
    """V2VNet det (5 agents) and FaFNet (2 maps): end-to-end error of the fp8-corrected mode next to fp16x3 / fp16a2 / fp16."""
    sd = synth.v2vnet_det_state(seed)
    bevs, trans, nat = synth.make_scene(1, 5, seed)
    fsd = synth.fafnet_state(seed)
    fbev = synth.make_bevs(2, seed)
    with torch.no_grad():
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=1, agent_num=5, gnn_iter=3)
        fref = restate.fafnet_forward(fbev, fsd)
    is_gru = lambda x, w: "fp16" if w.shape[0] == 768 else None       # noqa: E731  (the ConvGRU stays 1 pass, as in `mixed`)
    cases = [("fp16x3 (3 passes)", emulate(lambda x, w: MODES["fp16x3"][:2])),
             ("fp16a2 (2 passes, weights single-rounded)", emulate(lambda x, w: MODES["fp16a2"][:2])),
             ("fp16 (1 pass)", emulate(lambda x, w: MODES["fp16"][:2])),
             ("fp16 + 2 x e5m2 corrections (2 pass-equivalents)", emulate_fp8()),
             ("same, ConvGRU at 1 fp16 pass (1.7 pass-equivalents on V2VNet)", emulate_fp8(keep=is_gru)),
             ("same, shifts 6 / 6", emulate_fp8(shift_w=6, shift_a=6)),
             ("same, shifts 10 / 10", emulate_fp8(shift_w=10, shift_a=10))]
    for name, ctx in cases:
        with torch.no_grad(), ctx:
            out = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=1, agent_num=5, gnn_iter=3)
            fout = restate.fafnet_forward(fbev, fsd)
        print("%-66s v2v loc %.2e cls %.2e | faf loc %.2e cls %.2e" % (
            name, rel_err(out["loc"], ref["loc"]), rel_err(out["cls"], ref["cls"]),
            rel_err(fout["loc"], fref["loc"]), rel_err(fout["cls"], fref["cls"])), flush=True)
    # the other BASELINE models: seg When2Com_UNet (softmax fuse: no thresholded gate in the way) and seg UNet
    wsd = synth.seg_when2com_state(seed)
    wx, wtrans, wnat = synth.make_seg_scene(1, 5, seed)
    usd = synth.seg_unet_state(seed)
    with torch.no_grad():
        wref = restate.seg_when2com_forward(wx, wtrans, wnat, wsd, agent_num=5, warp_flag=1, inference="softmax")
        uref = restate.seg_unet_forward(wx[:2], usd)
    for name, mk in (("fp16x3 (3 passes)", lambda: emulate(lambda x, w: MODES["fp16x3"][:2])),
                     ("fp16 (1 pass)", lambda: emulate(lambda x, w: MODES["fp16"][:2])),
                     ("fp16 + 2 x e5m2 corrections (2 pass-equivalents)", lambda: emulate_fp8())):
        with torch.no_grad(), mk():
            wout = restate.seg_when2com_forward(wx, wtrans, wnat, wsd, agent_num=5, warp_flag=1, inference="softmax")
            uout = restate.seg_unet_forward(wx[:2], usd)
        print("%-66s seg when2com logits %.2e | seg unet logits %.2e" % (name, rel_err(wout, wref), rel_err(uout, uref)), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "fp8":
    run_fp8()
