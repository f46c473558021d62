"""Per-launch ablation of the V2VNet forward: time every launch with one pipeline role removed
(no MMA / no TMA / no stores) to see what bounds each layer.  Profiling aid; run on the GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "v2x-sim_b200")]
import torch
from v2x_b200 import synthetic as synth
from v2x_b200 import _lib, nets

NAMES = ["pack_in","pre_1","pre_2","c1_1","c1_2","c3d_1","c2_1","c2_2","c3d_2","c3_1","c3_2","c4_1","c4_2","warp","gru_m","gru1","gru2","gru3","c5_1","c5_2","c6_1","c6_2","c7_1","c7_2","c8_1","c8_2","head1","head2"]

def time_launches(plan, iters=5):
    n = len(plan.launches); tot = [0.0] * n
    for it in range(iters + 1):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i, l in enumerate(plan.launches):
            l(); evs[i + 1].record()
        torch.cuda.synchronize()
        if it:
            for i in range(n): tot[i] += evs[i].elapsed_time(evs[i + 1]) * 1e3 / iters
    return tot

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    planes = sys.argv[2] if len(sys.argv) > 2 else "bf16"   # precision mode name
    lib = _lib.load()
    sd = synth.v2vnet_det_state(0)
    bevs, trans, nat = synth.make_scene(B, 5, 0)
    plan = nets.V2VNetDetPlan(sd, B, 5, planes=planes)
    plan.set_inputs(bevs.cuda(), trans.cuda(), nat.cuda())
    res = {}
    for mode, tag in ((0, "full"), (1, "noMMA"), (2, "noTMA"), (3, "noST"), (4, "noEPI"), (0, "full2")):
        lib.v2x_set_debug_mode(mode)
        # launches bake the mode at call time (fill_dev), so just re-run
        res[tag] = time_launches(plan)
    lib.v2x_set_debug_mode(0)
    print("%-8s %9s %9s %9s %9s %9s %9s   GF      TF/s" % ("layer", "full", "noMMA", "noTMA", "noST", "noEPI", "full2"))
    for i, name in enumerate(NAMES[:len(plan.launches)]):
        fl = getattr(plan.launches[i], "flops", 0.0)
        print("%-8s %9.0f %9.0f %9.0f %9.0f %9.0f %9.0f   %6.1f %7.0f" % (name, res["full"][i], res["noMMA"][i], res["noTMA"][i], res["noST"][i], res["noEPI"][i], res["full2"][i], fl / 1e9, fl / (res["full"][i] * 1e-6) / 1e12 if fl else 0))
    print("total", {k: round(sum(v)) for k, v in res.items()})
    json.dump({"names": NAMES, "us": res}, open(os.path.join(ROOT, "gpurun_out", "ablate_B%d_%s.json" % (B, planes)), "w"))

if __name__ == "__main__":
    main()
