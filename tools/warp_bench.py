"""Times the cross-agent warp + fuse kernels on the bench shapes: the shared-memory-staged kernel (csrc/warp_staged.cuh)
against the direct one-warp-per-pixel gathers (the default; V2X_WARP_STAGED=1 selects the staged kernel).  CUDA events on the launching stream, inputs of one
step (L2-resident, as in the real step: the producer conv has just written them).  Prints one JSON line per case.
usage: python tools/warp_bench.py [--iters 200]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "v2x-sim_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    from v2x_b200 import ops, synthetic
    ops.require_gpu()
    dev = torch.device("cuda")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    cases = [("v2v_det mean", "mean", 8, 5, 256, 32, 2), ("v2v_det mean bf16", "mean", 8, 5, 256, 32, 1),
             ("w2c_det gated", "gated", 8, 5, 256, 32, 2), ("w2c_seg gated", "gated", 4, 5, 512, 32, 2),
             ("seg_v2v mean+self", "mean_self", 4, 5, 512, 32, 2), ("fusion reduce max", "max", 8, 5, 256, 32, 2),
             ("disco weighted", "weighted", 8, 5, 256, 32, 2)]
    for name, kind, B, A, C, HW, planes in cases:
        g = torch.Generator().manual_seed(1)
        x = ops.pack_input(torch.randn((A * B, HW, HW, C), generator=g).to(dev), C, planes)
        out = torch.empty_like(x)
        trans = synthetic.make_trans_matrices(B, A, 3).to(dev)
        nat = torch.full((B, A), A, dtype=torch.long, device=dev)
        coef = torch.rand((B, A, A), generator=g).to(dev)
        scores = torch.randn((B, A, A, HW * HW), generator=g).to(dev)
        fn = {"mean": lambda: ops.warp_mean(x, trans, nat, B, A, out=out),
              "mean_self": lambda: ops.warp_mean(x, trans, nat, B, A, include_self=True, out=out),
              "gated": lambda: ops.warp_gated(x, trans, nat, coef, B, A, warp_flag=1, out=out),
              "max": lambda: ops.warp_reduce(x, trans, nat, B, A, "max", out=out),
              "weighted": lambda: ops.warp_weighted(x, trans, nat, scores, B, A, per_pixel=True, out=out)}[kind]
        os.environ["V2X_WARP_STAGED"] = "0"
        t_direct = timed(fn, args.iters)
        os.environ["V2X_WARP_STAGED"] = "1"
        t_staged = timed(fn, args.iters)
        # algorithmic bytes: every source map read once + every output map written once (warp kernel alone)
        alg = 2 * x.numel() * 2
        rec = {"case": name, "maps": A * B, "C": C, "HxW": "%dx%d" % (HW, HW), "planes": planes,
               "direct_us": round(t_direct, 2), "staged_us": round(t_staged, 2), "speedup": round(t_direct / t_staged, 2),
               "algorithmic_MB": round(alg / 1e6, 1), "staged_GBps_algorithmic": round(alg / t_staged / 1e3, 1),
               "hbm_peak_GBps": peaks.get("hbm_gbs")}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
