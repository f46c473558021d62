"""GPU parity of the whole-forward plans against the CPU oracle and the live-reference fixtures.

north-star tolerance: fp activations within 1e-3 relative (max-abs error / max-abs value), argmax indices identical.
Modes (v2x_b200/precision.py): "mixed" -- the DEFAULT of every drop-in module and the mode bench.py reports -- and
"fp16x3" are both held to 1e-3 on every stage tensor and on loc / cls; "bf16" (one bf16 plane, the raw-throughput mode)
is outside that contract by construction and is held to a bf16-sized bound instead.  Every measured number goes to
gpurun_out/parity.json (committed copy: profiles/r02_parity.json)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MODES = ["mixed", "fp16x3", "bf16"]
REL_TOL = {"mixed": 1e-3, "fp16x3": 1e-3, "bf16": 5e-2}
# argmax flips tolerated as a fraction of positions (all must be near-ties, see argmax_flips): measured 3e-6 (fp16x3),
# 2e-4 (mixed), 5e-3 (bf16)
FLIP_TOL = {"mixed": 5e-4, "fp16x3": 1e-4, "bf16": 2e-2}


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def argmax_flips(out_cls, ref_cls, max_abs_err):
    """(all flips, flips where the reference margin |l0 - l1| exceeds 2x the measured max-abs error).
    A different summation order can legitimately flip near-ties (the reference's own cuDNN vs oneDNN
    runs differ the same way); a flip at a margin above the error bound would be a real defect."""
    out_cls, ref_cls = out_cls.detach().float().cpu(), ref_cls.detach().float().cpu()
    flip = out_cls.argmax(-1) != ref_cls.argmax(-1)
    margin = (ref_cls[..., 0] - ref_cls[..., 1]).abs()
    return int(flip.sum()), int((flip & (margin > 2 * max_abs_err)).sum())


def _golden_sub(t, g, name):
    from oracle.gen_golden import STRIDE
    sub = t.detach().float().cpu().contiguous().view(-1)[::STRIDE].numpy()
    ref = g[name + ".sub"]
    return float(np.abs(sub - ref).max() / np.abs(ref).max())


def _check_outputs(tag, mode, out, ref, g, parity_log, stages=None):
    rec = {}
    for name, (got, want) in (stages or {}).items():
        rec["stage_" + name] = rel_err(got, want)
        assert rec["stage_" + name] < REL_TOL[mode], (tag, mode, name, rec["stage_" + name])
    for k in ("loc", "cls"):
        e, eg = rel_err(out[k], ref[k]), _golden_sub(out[k], g, k)
        rec[k], rec[k + "_golden"] = e, eg
        assert out[k].shape == ref[k].shape
        assert e < REL_TOL[mode] and eg < REL_TOL[mode], (tag, mode, k, e, eg)
    err = (out["cls"].cpu() - ref["cls"]).abs().max().item()
    flips, bad = argmax_flips(out["cls"], ref["cls"], err)
    n = ref["cls"].numel() // 2
    rec.update(argmax_flips=flips, argmax_flips_outside_margin=bad, argmax_positions=n)
    parity_log(tag, mode, **rec)
    print("%s %s: %s" % (tag, mode, {k: ("%.2e" % v if isinstance(v, float) else v) for k, v in rec.items()}))
    assert bad == 0
    assert flips <= FLIP_TOL[mode] * n


@pytest.mark.parametrize("mode", MODES)
def test_fafnet_forward(mode, golden_dir, parity_log):
    from oracle import restate, synth
    from v2x_b200 import nets, ops
    g = np.load(os.path.join(golden_dir, "fafnet_n2_seed0.npz"))
    n, seed = [int(v) for v in g["meta"]]
    sd = synth.fafnet_state(seed)
    bevs = synth.make_bevs(n, seed)
    with torch.no_grad():
        ref = restate.fafnet_forward(bevs, sd, stages=True)
    plan = nets.FaFNetPlan(sd, n, planes=mode)
    out = plan.forward(bevs.cuda())
    torch.cuda.synchronize()
    stages = {"x3": (ops.act_to_float(plan.ws["x3"]), ref["enc"][3]), "x8": (ops.act_to_float(plan.ws["x8"]), ref["dec"][0])}
    _check_outputs("fafnet_n2_seed0", mode, out, ref, g, parity_log, stages)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", ["v2vnet_det_A5B1_seed0", "v2vnet_det_A5B2_seed1_present53"])
def test_v2vnet_det_forward(tag, mode, golden_dir, parity_log):
    from oracle import restate, synth
    from v2x_b200 import nets, ops
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    batch, a, seed, gnn = [int(v) for v in g["meta"]]
    present = [int(v) for v in g["present"]] if "present" in g.files else None
    sd = synth.v2vnet_det_state(seed)
    bevs, trans, nat = synth.make_scene(batch, a, seed, present=present)
    with torch.no_grad():
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=a, gnn_iter=gnn, stages=True)
    plan = nets.V2VNetDetPlan(sd, batch, a, gnn_iter=gnn, planes=mode)
    out = plan.forward(bevs.cuda(), trans.cuda(), nat.cuda())
    torch.cuda.synchronize()
    want = {"x0": ref["enc"][0], "x1": ref["enc"][1], "x2": ref["enc"][2], "x3": ref["enc"][3], "h%d" % gnn: ref["fused"],
            "x8": ref["x8"]}
    stages = {name: (ops.act_to_float(plan.ws[name]), r) for name, r in want.items()}
    _check_outputs(tag, mode, out, ref, g, parity_log, stages)


def test_v2vnet_graph_replay_matches_eager():
    """The CUDA-graph replay is the same launch list: results must be bit-identical to the eager run."""
    from oracle import synth
    from v2x_b200 import nets
    sd = synth.v2vnet_det_state(0)
    bevs, trans, nat = synth.make_scene(1, 5, 0)
    plan = nets.V2VNetDetPlan(sd, 1, 5, planes=1)
    out = plan.forward(bevs.cuda(), trans.cuda(), nat.cuda())
    eager = {k: v.clone() for k, v in out.items()}
    plan.capture()
    out = plan.forward(bevs.cuda(), trans.cuda(), nat.cuda())
    torch.cuda.synchronize()
    for k in eager:
        assert torch.equal(eager[k], out[k])


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", ["when2com_det_warp_activated_seed2", "when2com_det_nowarp_argmax_seed3_present4",
                                 "when2com_det_warp_softmax_B2_seed4"])
def test_when2com_det_forward(tag, mode, golden_dir, parity_log):
    check_when2com_det(tag, mode, golden_dir, parity_log)


def check_when2com_det(tag, mode, golden_dir, parity_log):
    """det When2com / who2com through the drop-in module, vs oracle + live-reference fixture.
    Constructor options beside the defaults (has_query, sparse, layer) come from the fixture (tests/test_gpu_zz_options.py).

    The eval gate is DISCRETE (p > 0.2 / argmax over keys, When2com.py:94-148): a gate that the oracle's own attention
    puts within the path's measured attention error of its threshold may legitimately land on the other side.  Stated
    bound: the attention map is within ATTN_TOL of the oracle's; every gate the path decides differently is such a
    near-threshold gate; and whenever all gates agree the outputs meet the mode's tolerance."""
    from coperception.models.det import When2com
    from oracle import restate, synth
    from v2x_b200 import default_det_config
    ATTN_TOL = {"mixed": 1e-3, "fp16x3": 1e-3, "bf16": 5e-2}
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    batch, a, seed, warp = [int(v) for v in g["meta"]]
    inference = str(g["inference"])
    present = [int(v) for v in g["present"]] if "present" in g.files else None
    opts = [int(v) for v in g["options"]] if "options" in g.files else [1, 0, 3]
    has_query, sparse, layer = bool(opts[0]), bool(opts[1]), opts[2]
    sd = synth.when2com_det_state(seed, has_query=has_query)
    bevs, trans, nat = synth.make_scene(batch, a, seed, present=present)
    with torch.no_grad():
        ref = restate.when2com_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=a, warp_flag=warp,
                                           inference=inference, stages=True, has_query=has_query, layer=layer)
    model = When2com(default_det_config(), layer=layer, warp_flag=warp, num_agent=a, has_query=has_query, sparse=sparse)
    model.load_state_dict(sd, strict=True)
    model.precision = mode
    model = model.cuda().eval()
    with torch.no_grad():
        out = model(bevs.cuda(), trans.cuda(), nat.cuda(), training=False, MO_flag=True, inference=inference,
                    batch_size=batch)
    torch.cuda.synchronize()
    plan = model._plan_list()[0]
    attn, attn_ref = plan.attn.cpu(), ref["attn"]
    attn_err = (attn - attn_ref).abs().max().item()
    assert attn_err < ATTN_TOL[mode], (tag, mode, attn_err)
    eye = 0.001 * torch.eye(a).unsqueeze(0)
    p_out, p_ref = attn + eye, attn_ref + eye
    if inference == "activated":
        differ = (p_out > 0.2) != (p_ref > 0.2)
        near = (p_ref - 0.2).abs() <= 2 * attn_err
    elif inference == "argmax_test":
        differ = p_out.argmax(1) != p_ref.argmax(1)
        top2 = p_ref.topk(2, dim=1).values
        near = (top2[:, 0] - top2[:, 1]) <= 2 * attn_err
    else:
        differ = near = torch.zeros(1, dtype=torch.bool)
    assert not bool((differ & ~near).any()), "a gate flipped although the oracle's attention is clear of the threshold"
    gates_agree = not bool(differ.any())
    rec = dict(attn_err=attn_err, gates_agree=gates_agree)
    for k in ("loc", "cls"):
        rec[k], rec[k + "_golden"] = rel_err(out[k], ref[k]), _golden_sub(out[k], g, k)
        assert out[k].shape == ref[k].shape
        if gates_agree:
            assert rec[k] < REL_TOL[mode] and rec[k + "_golden"] < REL_TOL[mode], (tag, mode, k, rec)
    if gates_agree:
        err = (out["cls"].cpu() - ref["cls"]).abs().max().item()
        flips, bad = argmax_flips(out["cls"], ref["cls"], err)
        rec.update(argmax_flips=flips, argmax_flips_outside_margin=bad)
        assert bad == 0
    parity_log(tag, mode, **rec)
    print("when2com %s %s: %s" % (tag, mode, rec))
    if mode != "bf16":
        assert gates_agree, "the 1e-3 modes must reproduce the oracle's gates on these fixtures"


@pytest.mark.parametrize("mode", MODES)
def test_v2vnet_map_parity_planted(mode, parity_log):
    """mAP parity (BASELINE metric: "mAP@0.5 parity vs ref", north-star: within 0.1): planted-head weights so that
    ~150 anchors per agent pass the reference's 0.7 score filter, then the reference's evaluation chain
    restated in oracle/postproc.py (softmax -> decode -> corners -> polygon NMS -> area AP) is applied to the oracle's
    and to the sm_100a path's (loc, cls).  Tolerance in mAP POINTS (the reference prints mAP x 100): the 1e-3 modes
    ("mixed", "fp16x3") within 0.1 point; bf16 -- outside the parity contract -- within 2 points."""
    from oracle import postproc as pp, restate, synth
    from v2x_b200 import nets
    seed = 0
    sd0 = synth.v2vnet_det_state(seed)
    bevs, trans, nat = synth.make_scene(1, 5, seed)
    with torch.no_grad():
        ref0 = restate.v2vnet_det_forward(bevs, trans, nat, sd0, batch_size=1, agent_num=5, gnn_iter=3)
        sd = synth.plant_detections(sd0, ref0["cls"], per_agent=150)
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=1, agent_num=5, gnn_iter=3)
    plan = nets.V2VNetDetPlan(sd, 1, 5, gnn_iter=3, planes=mode)
    out = plan.forward(bevs.cuda(), trans.cuda(), nat.cuda())
    torch.cuda.synchronize()
    det_ref, sel_ref = pp.detections_of(ref["loc"].numpy(), ref["cls"].numpy())
    det_out, sel_out = pp.detections_of(out["loc"].float().cpu().numpy(), out["cls"].float().cpu().numpy())
    assert all(len(s) > 20 for s in sel_ref)
    gts = synth.make_gt_from_detections(det_ref, seed=1)
    res = {}
    for thr in (0.5, 0.7):
        res[thr] = (pp.eval_map(det_ref, gts, thr)[0], pp.eval_map(det_out, gts, thr)[0])
    same = sum(int(set(a.tolist()) == set(b.tolist())) for a, b in zip(sel_ref, sel_out))
    common = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(sel_ref, sel_out))
    total = sum(len(a) for a in sel_ref)
    print("mAP %s: ref/out @0.5 %.4f/%.4f @0.7 %.4f/%.4f; NMS picks identical for %d/5 agents, %d/%d common"
          % (mode, res[0.5][0], res[0.5][1], res[0.7][0], res[0.7][1], same, common, total))
    parity_log("v2vnet_map_planted_seed0", mode, map50_ref_points=100 * res[0.5][0], map50_out_points=100 * res[0.5][1],
               map70_ref_points=100 * res[0.7][0], map70_out_points=100 * res[0.7][1], nms_common=common, nms_total=total)
    assert 0.2 < res[0.5][0] <= 1.0
    tol_points = 0.1 if mode != "bf16" else 2.0
    assert 100 * abs(res[0.5][0] - res[0.5][1]) <= tol_points and 100 * abs(res[0.7][0] - res[0.7][1]) <= tol_points
    if mode != "bf16":
        # a score within the path's error of the 0.7 filter may enter/leave the candidate set; everything else is exact
        assert common >= 0.98 * total
