"""GPU experiment behind the "mixed" precision policy (not a pytest module): for the V2VNet det and FaFNet fixtures, run
the sm_100a plans with ONE layer (or a candidate set) at fewer tensor-core passes -- everything else at fp16x3 -- and
print the end-to-end error against the fp32 CPU oracle in the parity tests' metric.

  python tests/precision_sweep.py [single|combos]      -> gpurun_out/precision_sweep.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "v2x-sim_b200")]
import torch  # noqa: E402

from oracle import restate, synth  # noqa: E402
from v2x_b200 import nets, ops  # noqa: E402

LAYERS = ["conv_pre_1", "conv_pre_2", "conv1_1", "conv1_2", "conv3d_1", "conv2_1", "conv2_2", "conv3d_2", "conv3_1",
          "conv3_2", "conv4_1", "conv4_2", "gru", "conv5_1", "conv5_2", "conv6_1", "conv6_2", "conv7_1", "conv7_2",
          "conv8_1", "conv8_2", "heads"]


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "single"
    cases = {}
    sd = synth.v2vnet_det_state(0)
    bevs, trans, nat = synth.make_scene(1, 5, 0)
    with torch.no_grad():
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=1, agent_num=5, gnn_iter=3, stages=True)
    cases["v2v"] = (lambda: nets.V2VNetDetPlan(sd, 1, 5, gnn_iter=3, planes="mixed"),
                    lambda p: p.forward(bevs.cuda(), trans.cuda(), nat.cuda()), ref, "h3", ref["fused"])
    sdf = synth.fafnet_state(0)
    bf = synth.make_bevs(2, 0)
    with torch.no_grad():
        reff = restate.fafnet_forward(bf, sdf, stages=True)
    cases["faf"] = (lambda: nets.FaFNetPlan(sdf, 2, planes="mixed"), lambda p: p.forward(bf.cuda()), reff, "x8", reff["dec"][0])

    def run(policy):
        os.environ["V2X_MIXED_POLICY"] = ",".join("%s=%d" % kv for kv in policy.items())
        out = {}
        for name, (mk, fwd, r, stage, stage_ref) in cases.items():
            plan = mk()
            o = fwd(plan)
            torch.cuda.synchronize()
            out[name] = dict(loc=rel(o["loc"], r["loc"]), cls=rel(o["cls"], r["cls"]),
                             stage=rel(ops.act_to_float(plan.ws[stage]), stage_ref))
        return out

    results = []
    if what == "single":
        sweeps = [({}, "all fp16x3")] + [({l: m}, "%s=%d" % (l, m)) for l in LAYERS for m in ((2, 1) if l == "gru" else (2,))]
    else:
        base = {"gru": 1}
        sweeps = [(dict(base), "gru=1"), ({"gru": 2}, "gru=2")]
        for extra in (["conv5_1"], ["conv7_1"], ["conv8_1"], ["conv6_1"], ["conv5_1", "conv7_1"], ["conv5_1", "conv8_1"],
                      ["conv7_1", "conv8_1"], ["conv5_1", "conv6_1"], ["conv5_1", "conv7_1", "conv8_1"],
                      ["conv5_1", "conv6_1", "conv7_1"], ["conv5_1", "conv6_1", "conv7_1", "conv8_1"]):
            pol = dict(base)
            pol.update({l: 2 for l in extra})
            sweeps.append((pol, "gru=1 + " + "+".join(extra) + "=2"))
    for pol, tag in sweeps:
        r = run(pol)
        results.append(dict(policy=pol, tag=tag, **{k + "_" + m: v[m] for k, v in r.items() for m in v}))
        print("%-46s v2v loc %.2e cls %.2e h3 %.2e | faf loc %.2e cls %.2e x8 %.2e" % (
            tag, r["v2v"]["loc"], r["v2v"]["cls"], r["v2v"]["stage"], r["faf"]["loc"], r["faf"]["cls"], r["faf"]["stage"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "precision_sweep_%s.json" % what), "w"), indent=1)


if __name__ == "__main__":
    main()
