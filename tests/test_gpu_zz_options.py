"""GPU: constructor options of the drop-in modules beside the reference scripts' defaults.

This file sorts LAST on purpose.  It was written after the round's GPU budget was spent: everything here has been checked
as far as a CPU can check it -- the oracle is pinned to live-reference fixtures for every case
(tests/test_oracle_golden.py), every launch list and training tape is dry-run against a fake library that validates each
ctypes call (tests/test_plan_dryrun.py), the library builds -- but its first run on a B200 is the driver's round-end run.
Each case only recombines kernels the files before it already exercise (the same launches with a different operand, layer
or channel count); the parity bounds are the ones of those files.

Because these cases have never met the hardware, they are (a) ISOLATED: every group runs in a child pytest process with a
time limit, so a fault in one of them (a sticky CUDA error, a kernel that never returns) ends with that child and cannot
disturb this session or leave the GPU busy; and (b) RECORDED rather than gated: a pass shows as XPASS, a miss as XFAIL with
the child's output in gpurun_out/first_run_<group>.log, and neither hides the cases before this file that have run.
Remove the marker of a group once a round-end log shows XPASS.
"""
import os

import pytest

pytestmark = pytest.mark.gpu

MODES = ["mixed", "fp16x3", "bf16"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = os.environ.get("V2X_ZZ_CHILD") == "1"
in_child_only = pytest.mark.skipif(not CHILD, reason="runs in a child process: test_first_hardware_run")
first_hardware_run = in_child_only      # the numerics cases below carry this marker

GROUPS = ["test_when2com_det_options", "test_seg_when2com_options", "test_v2vnet_compressed_train_step_matches_oracle",
          "test_seg_unet_compressed_train_step_matches_oracle", "test_v2vnet_layer4", "test_sum_fusion_layer4"]
CHILD_TIME_LIMIT_S = 300      # a group takes 0.5 - 2 min; six groups bound the worst case at 30 min


@pytest.mark.skipif(CHILD, reason="the parent-side launcher")
@pytest.mark.xfail(strict=False, reason="first B200 run of this group (written after the round's GPU budget was spent; oracle "
                   "pinned and host wiring dry-run on the CPU)")
@pytest.mark.parametrize("group", GROUPS)
def test_first_hardware_run(group):
    import torch
    from child_run import ran_and_passed, run_in_child
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    log = os.path.join(ROOT, "gpurun_out", "first_run_%s.log" % group)
    rc, tail = run_in_child(__file__, group + " and not first_hardware_run", log, CHILD_TIME_LIMIT_S, marker="gpu",
                            env=dict(V2X_ZZ_CHILD="1", V2X_PARITY_FILE="parity_first_run_%s.json" % group))
    print(tail)
    assert ran_and_passed(rc, tail), "child run of %s ended with %s; see %s" % (group, rc, log)


# ---- host-logic refusals first: they launch no kernel, so nothing a later first-run case does can disturb them ----------
def test_when2com_refuses_what_the_reference_cannot_run():
    """MO_flag=False and layer=2 + argmax_test raise in the reference itself (profiles/r02_reference_option_probe.txt)."""
    import torch
    from coperception.models.det import When2com
    from oracle import synth
    from v2x_b200 import default_det_config
    bevs, trans, nat = synth.make_scene(1, 5, 0)
    m = When2com(default_det_config(), layer=2, num_agent=5).cuda().eval()
    with torch.no_grad():
        with pytest.raises(NotImplementedError):
            m(bevs.cuda(), trans.cuda(), nat.cuda(), training=False, MO_flag=False, batch_size=1)
        with pytest.raises(NotImplementedError):
            m(bevs.cuda(), trans.cuda(), nat.cuda(), training=False, inference="argmax_test", batch_size=1)


def test_training_refuses_fewer_than_32_compressed_channels():
    import torch
    from coperception.models.det import V2VNet
    from oracle import synth
    from v2x_b200 import default_det_config
    bevs, trans, nat = synth.make_scene(1, 5, 0)
    m = V2VNet(default_det_config(), 3, 3, 256, num_agent=5, compress_level=4).cuda().train()
    with pytest.raises(NotImplementedError):
        m(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)


# ---- numerics cases ------------------------------------------------------------------------------------------------
@first_hardware_run
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", ["when2com_det_noquery_activated_seed41",
                                 "when2com_det_layer2_sparse_activated_seed42_present4",
                                 "when2com_det_layer2_noquery_nowarp_softmax_B2_seed43",
                                 "when2com_det_layer4_activated_seed48"])
def test_when2com_det_options(tag, mode, golden_dir, parity_log):
    """has_query=False (ones as every agent's query, When2com.py:241-245), sparse=True (a no-op in the reference,
    :374-412), layer=2 and layer=4 (:167-190; the fused 16x16 map reaches the decoder through DetPlan.upsample2), against
    the oracle and the live-reference fixtures."""
    from test_gpu_nets import check_when2com_det
    check_when2com_det(tag, mode, golden_dir, parity_log)


@first_hardware_run
@pytest.mark.parametrize("mode", MODES)
def test_seg_when2com_options(mode, golden_dir, parity_log):
    """seg When2Com_UNet(has_query=False, sparse=True) (When2Com_UNet.py:219-225, 408-446)."""
    from test_gpu_seg import check_seg_model
    check_seg_model("seg_when2com_noquery_sparse_activated_seed44", "when2com", mode, golden_dir, parity_log)


@first_hardware_run
def test_v2vnet_compressed_train_step_matches_oracle(golden_dir, parity_log):
    """det V2VNet(compress_level=2) in .train(): the compresser pair of the communicated layer (Backbone.py:138-141) with
    train-mode BatchNorm on the tape; x_4 is computed from the uncompressed x_3, so x_3 collects two gradients."""
    import os

    import numpy as np
    import torch
    from coperception.models.det import V2VNet
    from oracle import restate
    from oracle.gen_golden import make_upstream, train_case
    from test_gpu_train import _check_buffers, _check_grads, _rel
    from v2x_b200 import default_det_config
    tag, seed = "train_step_v2vnet_c2_seed45", 45
    golden = np.load(os.path.join(golden_dir, tag + ".npz"))
    sd, (bevs, trans, nat), _ = train_case("v2vnet_c2", seed)
    sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    up = make_upstream({"loc": (bevs.shape[0], 256, 256, 6, 1, 6), "cls": (bevs.shape[0], 256 * 256 * 6, 2)}, seed)
    out_ref, grads_ref, sd_after = restate.train_step_vjp(
        lambda s: restate.v2vnet_det_forward(bevs.double(), trans, nat, s, batch_size=1, agent_num=5, gnn_iter=3,
                                             compress_level=2), sd64, up)
    model = V2VNet(default_det_config(), 3, 3, 256, num_agent=5, compress_level=2)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    out = model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
    for k in ("loc", "cls"):
        e = _rel(out[k], out_ref[k])
        print("v2vnet compress 2 train forward", k, "rel_err %.3e" % e)
        assert e < 1e-3
    torch.autograd.backward([out["cls"], out["loc"]], [up["cls"].float().cuda(), up["loc"].float().cuda()])
    torch.cuda.synchronize()
    got = {k: p.grad for k, p in model.named_parameters()}
    for k, g in got.items():
        assert (g is not None) == (k in grads_ref), k
    assert "u_encoder.com_compresser.weight" in grads_ref and "u_encoder.bn_decompress.weight" in grads_ref
    _check_grads(tag, got, grads_ref, golden, parity_log)
    _check_buffers(tag, dict(model.named_buffers()), sd_after, parity_log, golden=golden)


@first_hardware_run
def test_seg_unet_compressed_train_step_matches_oracle(golden_dir, parity_log):
    """seg UNet(compress_level=3) in .train() (UNet.py:30-32: the pair replaces x4 for the whole decoder)."""
    import os

    import numpy as np
    import torch
    from coperception.models.seg import UNet
    from oracle import restate
    from oracle.gen_golden import make_upstream, train_case
    from test_gpu_train import _check_buffers, _check_grads, _rel
    tag, seed = "train_step_seg_unet_c3_seed46", 46
    golden = np.load(os.path.join(golden_dir, tag + ".npz"))
    sd, (x,), _ = train_case("seg_unet_c3", seed)
    sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    up = make_upstream({"logits": (x.shape[0], 8, 256, 256)}, seed)
    out_ref, grads_ref, sd_after = restate.train_step_vjp(lambda s: {"logits": restate.seg_unet_forward(x.double(), s)}, sd64, up)
    model = UNet(13, 8, compress_level=3)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    out = model(x.cuda())
    e = _rel(out, out_ref["logits"])
    print("seg unet compress 3 train forward logits rel_err %.3e" % e)
    assert out.shape == out_ref["logits"].shape and e < 1e-3
    out.backward(up["logits"].float().cuda())
    torch.cuda.synchronize()
    got = {k: p.grad for k, p in model.named_parameters()}
    for k, g in got.items():
        assert (g is not None) == (k in grads_ref), k
    _check_grads(tag, got, grads_ref, golden, parity_log, min_params=40)
    _check_buffers(tag, dict(model.named_buffers()), sd_after, parity_log, golden=golden, rtol=2e-5)


@first_hardware_run
@pytest.mark.parametrize("mode", MODES)
def test_v2vnet_layer4(mode, golden_dir, parity_log):
    """det V2VNet communicating at layer 4 (DetModelBase.py:71-92): 512-channel maps at 16 x 16, two GNN rounds, one
    absent agent slot; the decoder reads the fused map 2x-upsampled (Backbone.py:176)."""
    import os

    import numpy as np
    import torch
    from coperception.models.det import V2VNet
    from oracle import restate, synth
    from test_gpu_nets import REL_TOL, _golden_sub, rel_err
    from v2x_b200 import default_det_config
    tag = "layer4_v2vnet_det_seed47"
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    sd = synth.v2vnet_det_state(47, layer_channel=512)
    bevs, trans, nat = synth.make_scene(1, 5, 47, present=[4])
    with torch.no_grad():
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=1, agent_num=5, gnn_iter=2, layer=4)
    m = V2VNet(default_det_config(), 2, 4, 512, num_agent=5)
    m.load_state_dict(sd, strict=True)
    m.precision = mode
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
    torch.cuda.synchronize()
    rec = {}
    for k in ("loc", "cls"):
        rec[k], rec[k + "_golden"] = rel_err(out[k], ref[k]), _golden_sub(out[k], g, k)
    parity_log(tag, mode, **rec)
    print("v2vnet@4 %s: %s" % (mode, rec))
    assert all(v < REL_TOL[mode] for v in rec.values()), rec


@first_hardware_run
@pytest.mark.parametrize("mode", MODES)
def test_sum_fusion_layer4(mode, golden_dir, parity_log):
    """SumFusion at layer 4 (FusionBase.py:23-75 with DetModelBase.py:71-92): 512-channel 16 x 16 maps, two absent slots."""
    import os

    import numpy as np
    import torch
    from coperception.models.det import SumFusion
    from oracle import restate, synth
    from test_gpu_nets import REL_TOL, _golden_sub, rel_err
    from v2x_b200 import default_det_config
    tag = "layer4_sum_det_seed49"
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    sd = synth.fusion_det_state("sum", 49)
    bevs, trans, nat = synth.make_scene(1, 5, 49, present=[3])
    with torch.no_grad():
        ref = restate.fusion_det_forward("sum", bevs, trans, nat, sd, batch_size=1, agent_num=5, layer=4)
    m = SumFusion(default_det_config(), layer=4, kd_flag=0, num_agent=5)
    m.load_state_dict(sd, strict=True)
    m.precision = mode
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
    torch.cuda.synchronize()
    rec = {}
    for k in ("loc", "cls"):
        rec[k], rec[k + "_golden"] = rel_err(out[k], ref[k]), _golden_sub(out[k], g, k)
    parity_log(tag, mode, **rec)
    print("sum@4 %s: %s" % (mode, rec))
    assert all(v < REL_TOL[mode] for v in rec.values()), rec
