"""One eager forward step of a bench config, for profilers:
   ncu ... --profile-from-start off python tools/prof_step.py --config w2c_seg --precision mixed
The plan is built and warmed up first; only ONE eager pass over the launch list sits between cudaProfilerStart/Stop.
Also writes gpurun_out/prof_<config>_<precision>_launches.json: per launch its label, algorithmic conv FLOPs and
tensor-core passes (the python-side launch list, in issue order), which tools/ncu_report.py joins with the ncu rows."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "v2x-sim_b200")]
import torch  # noqa: E402

from v2x_b200 import nets, nets_seg  # noqa: E402
from v2x_b200 import synthetic as synth  # noqa: E402


def build(config, precision, units):
    dev = torch.device("cuda")
    if config == "v2v_det":
        sd = synth.v2vnet_det_state(0)
        bevs, trans, nat = synth.make_scene(units, 5, 0)
        plan = nets.V2VNetDetPlan(sd, units, 5, gnn_iter=3, planes=precision, device=dev)
        plan.set_inputs(bevs.to(dev), trans.to(dev), nat.to(dev))
    elif config in ("faf_lower", "faf_upper_dp"):
        n = units * (6 if config == "faf_upper_dp" else 1)
        plan = nets.FaFNetPlan(synth.fafnet_state(0), n, planes=precision, device=dev)
        plan.set_bevs(synth.make_bevs(n, 0).to(dev))
    elif config == "w2c_seg":
        x, trans, nat = synth.make_seg_scene(units, 5, 0)
        plan = nets_seg.SegWhen2comPlan(synth.seg_when2com_state(0), units, 5, planes=precision, device=dev, warp_flag=1,
                                        inference="activated")
        plan.set_x(x.to(dev))
        plan.trans.copy_(trans.to(dev))
        plan.num_agent.copy_(nat.to(dev))
    else:
        raise ValueError(config)
    return plan


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="v2v_det")
    ap.add_argument("--precision", default="mixed")
    ap.add_argument("--units", type=int, default=0)
    args = ap.parse_args()
    units = args.units or {"v2v_det": 8, "faf_lower": 40, "faf_upper_dp": 4, "w2c_seg": 4}[args.config]
    plan = build(args.config, args.precision, units)
    os.environ["V2X_NO_SIDE_STREAM"] = "1"     # serial issue order = launch-list order (ncu serialises kernels anyway)
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    plan.run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    rows = []
    for i, l in enumerate(plan.launches):
        rows.append({"i": i, "label": getattr(l, "label", "aux"), "flops": getattr(l, "flops", 0.0),
                     "passes": getattr(l, "mma_passes", 0)})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "prof_%s_%s_launches.json" % (args.config, args.precision)), "w") as f:
        json.dump({"config": args.config, "precision": args.precision, "units": units, "launches": rows}, f)
    print("profiled one step of %s/%s: %d launches" % (args.config, args.precision, len(rows)))


if __name__ == "__main__":
    main()
