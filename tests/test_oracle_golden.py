"""CPU: the oracle restatement (oracle/restate.py) against fixtures made from the LIVE reference
modules (oracle/gen_golden.py -> tests/golden/*.npz).  Tolerance 2e-4 absolute on O(1) activations
(the restatement uses functional ops instead of nn.Modules; measured difference 1.5e-5)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate, synth
from oracle.gen_golden import STRIDE, argmax_checksum

ATOL = 2e-4


def _check(name, t, g):
    t = t.detach().float().contiguous()
    assert list(t.shape) == list(g[name + ".shape"]), name
    sub = t.view(-1)[::STRIDE].numpy()
    ref = g[name + ".sub"]
    err = np.abs(sub - ref).max()
    assert err < ATOL, (name, err)
    assert abs(t.double().abs().sum().item() - float(g[name + ".abssum"])) < 1e-5 * float(g[name + ".abssum"]) + 1e-3


def test_warp_small(golden_dir):
    g = np.load(os.path.join(golden_dir, "warp_small_seed3.npz"))
    local = torch.from_numpy(g["local"])
    trans = torch.from_numpy(g["trans"])
    C = local.shape[2]
    for k, (b, j, i) in enumerate(g["pairs"]):
        out = restate.feature_transformation(local, int(b), int(j), int(i), trans, (1, C, 32, 32))
        assert np.abs(out.numpy() - g["out"][k]).max() < 1e-5


def test_convgru_small(golden_dir):
    g = np.load(os.path.join(golden_dir, "convgru_small_seed4.npz"))
    sd = {"convgru." + k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    x = torch.from_numpy(g["x"])[0]
    y = restate.convgru_zero_hidden(x, sd)
    assert np.abs(y.numpy() - g["y"][0]).max() < 1e-5


def test_fafnet(golden_dir):
    g = np.load(os.path.join(golden_dir, "fafnet_n2_seed0.npz"))
    n, seed = [int(v) for v in g["meta"]]
    sd = synth.fafnet_state(seed)
    with torch.no_grad():
        r = restate.fafnet_forward(synth.make_bevs(n, seed), sd)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)


def test_v2vnet_stages(golden_dir):
    g = np.load(os.path.join(golden_dir, "v2vnet_det_stages_seed0.npz"))
    _, a, seed = [int(v) for v in g["meta"]]
    sd = synth.v2vnet_det_state(seed)
    bevs, trans, nat = synth.make_scene(1, a, seed)
    with torch.no_grad():
        r = restate.v2vnet_det_forward(bevs, trans, nat, sd, stages=True)
    for i in range(5):
        _check("enc%d" % i, r["enc"][i], g)
    _check("fused", r["fused"], g)
    _check("x8", r["x8"], g)


@pytest.mark.parametrize("tag", ["v2vnet_det_A5B1_seed0", "v2vnet_det_A5B2_seed1_present53"])
def test_v2vnet_det(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    batch, a, seed, gnn = [int(v) for v in g["meta"]]
    present = [int(v) for v in g["present"]] if "present" in g.files else None
    sd = synth.v2vnet_det_state(seed)
    bevs, trans, nat = synth.make_scene(batch, a, seed, present=present)
    with torch.no_grad():
        r = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=a, gnn_iter=gnn)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)
    cnt, chk = argmax_checksum(r["cls"])
    # argmax can legitimately flip where the two logits agree to ~1e-5; allow a handful
    assert np.abs(cnt - g["cls.argmax_count"]).max() <= 4


WHEN2COM_DET_FIXTURES = ["when2com_det_warp_activated_seed2", "when2com_det_nowarp_argmax_seed3_present4",
                         "when2com_det_warp_softmax_B2_seed4",
                         # constructor options beside the scripts' defaults (oracle/gen_golden.py::gen_options)
                         "when2com_det_noquery_activated_seed41", "when2com_det_layer2_sparse_activated_seed42_present4",
                         "when2com_det_layer2_noquery_nowarp_softmax_B2_seed43", "when2com_det_layer4_activated_seed48"]


def when2com_options(g):
    """(has_query, sparse, layer) of a when2com fixture (absent = the defaults True, False, 3)."""
    o = [int(v) for v in g["options"]] if "options" in g.files else [1, 0, 3]
    return bool(o[0]), bool(o[1]), (o[2] if len(o) > 2 else 3)


@pytest.mark.parametrize("tag", WHEN2COM_DET_FIXTURES)
def test_when2com_det(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    batch, a, seed, warp = [int(v) for v in g["meta"]]
    inference = str(g["inference"])
    present = [int(v) for v in g["present"]] if "present" in g.files else None
    has_query, _, layer = when2com_options(g)      # `sparse` never reaches the arithmetic (When2com.py:374-412)
    sd = synth.when2com_det_state(seed, has_query=has_query)
    bevs, trans, nat = synth.make_scene(batch, a, seed, present=present)
    with torch.no_grad():
        r = restate.when2com_det_forward(bevs, trans, nat, sd, batch_size=batch, agent_num=a, warp_flag=warp,
                                         inference=inference, has_query=has_query, layer=layer)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)


@pytest.mark.parametrize("tag,kind", [("seg_unet_seed0", "unet"), ("seg_v2vnet_seed1_present4", "v2vnet"),
                                      ("seg_when2com_warp_activated_seed2", "when2com"),
                                      ("seg_when2com_nowarp_activated_seed3", "when2com"),
                                      ("seg_when2com_noquery_sparse_activated_seed44", "when2com")])
def test_seg_models(golden_dir, tag, kind):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    batch, a, seed, warp = [int(v) for v in g["meta"]]
    present = [int(v) for v in g["present"]] if "present" in g.files else None
    x, trans, nat = synth.make_seg_scene(batch, a, seed, present=present)
    with torch.no_grad():
        if kind == "unet":
            r = restate.seg_unet_forward(x, synth.seg_unet_state(seed))
        elif kind == "v2vnet":
            r = restate.seg_v2vnet_forward(x, trans, nat, synth.seg_v2vnet_state(seed), agent_num=a)
        else:
            has_query = when2com_options(g)[0]
            r = restate.seg_when2com_forward(x, trans, nat, synth.seg_when2com_state(seed, has_query=has_query),
                                             agent_num=a, warp_flag=warp, inference=str(g["inference"]), has_query=has_query)
    _check("logits", r, g)


def _fusion_fixtures(family):
    from oracle.gen_golden import FUSION_FIXTURES
    return [f for f in FUSION_FIXTURES if f[1] == family]


@pytest.mark.parametrize("fx", _fusion_fixtures("det"), ids=lambda f: f[0])
def test_fusion_det(golden_dir, fx):
    """FusionBase family (Mean/Max/Sum/Cat/AgentWise/DiscoNet det) restatement vs the live reference modules."""
    tag, _, kind, batch, seed, present, v2i = fx
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    assert [int(v) for v in g["meta"]] == [batch, 5, seed, int(v2i)] and str(g["kind"]) == kind
    sd = synth.fusion_det_state(kind, seed)
    bevs, trans, nat = synth.make_scene(batch, 5, seed, present=present)
    with torch.no_grad():
        r = restate.fusion_det_forward(kind, bevs, trans, nat, sd, batch_size=batch, agent_num=5, only_v2i=v2i)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)


@pytest.mark.parametrize("fx", _fusion_fixtures("seg"), ids=lambda f: f[0])
def test_fusion_seg(golden_dir, fx):
    tag, _, kind, batch, seed, present, v2i = fx
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    sd = synth.seg_fusion_state(kind, seed)
    x, trans, nat = synth.make_seg_scene(batch, 5, seed, present=present)
    with torch.no_grad():
        r = restate.seg_fusion_forward(kind, x, trans, nat, sd, agent_num=5, only_v2i=v2i)
    _check("logits", r, g)


def test_teacher(golden_dir):
    g = np.load(os.path.join(golden_dir, "teacher_n1_seed3.npz"))
    n, seed = [int(v) for v in g["meta"]]
    with torch.no_grad():
        r = restate.teacher_forward(synth.make_bevs(n, seed), synth.fafnet_state(seed))
    for name, t in zip(("x8", "x7", "x6", "x5", "x3", "x4"), r):
        _check(name, t, g)


def test_fusion_weights_are_not_degenerate():
    """The synthetic pair-weight nets must produce non-uniform fusion weights, else the fixtures could not tell a
    wrong (target, member) pairing from a right one."""
    sd = synth.fusion_det_state("disco", 10)
    g = torch.Generator().manual_seed(0)
    tg, nb = torch.rand((256, 32, 32), generator=g), torch.rand((256, 32, 32), generator=g)
    with torch.no_grad():
        s0 = restate._pair_weight_net(torch.cat([tg, tg]).unsqueeze(0), sd, "pixel_weighted_fusion.")
        s1 = restate._pair_weight_net(torch.cat([tg, nb]).unsqueeze(0), sd, "pixel_weighted_fusion.")
    assert (s0 > 0).float().mean() > 0.5 and (s0 - s1).abs().mean() > 1e-2


def test_densify_voxels_is_scatter_then_rot90():
    """oracle.synth.densify_voxels (V2XSimDet.py:294-302): voxel (i0, i1, z) lands on pixel (i1, 255 - i0, z)."""
    lists = synth.make_voxel_indices(1, seed=2, points=500)
    d = synth.densify_voxels(lists[0])
    assert d.dtype == np.float32 and d.shape == (256, 256, 13)
    assert all(d[i1, 255 - i0, z] == 1.0 for i0, i1, z in lists[0])
    assert int(d.sum()) == len(np.unique(lists[0], axis=0))


def test_compress_level(golden_dir):
    """compress_level > 0: restatement vs the live reference (V2VNet det level 2, DiscoNet det level 6, seg UNet level 3
    with its kd_flag tuple)."""
    g = np.load(os.path.join(golden_dir, "compress_v2vnet_det_l2_seed15.npz"))
    sd = synth.v2vnet_det_state(15, compress_level=2)
    bevs, trans, nat = synth.make_scene(1, 5, 15, present=[4])
    with torch.no_grad():
        r = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=1, agent_num=5, gnn_iter=3, compress_level=2)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)
    g = np.load(os.path.join(golden_dir, "compress_disco_det_l6_seed16.npz"))
    sd = synth.fusion_det_state("disco", 16, compress_level=6)
    bevs, trans, nat = synth.make_scene(1, 5, 16)
    with torch.no_grad():
        r = restate.fusion_det_forward("disco", bevs, trans, nat, sd, batch_size=1, agent_num=5)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)
    g = np.load(os.path.join(golden_dir, "compress_seg_unet_l3_kd_seed17.npz"))
    sd = synth.seg_unet_state(17, compress_level=3)
    x, _, _ = synth.make_seg_scene(1, 2, 17)
    with torch.no_grad():
        _check("logits", restate.seg_unet_forward(x, sd), g)


def test_other_communication_layers(golden_dir):
    """Fusion at encoder layers other than 3 (DetModelBase.py:71-92, 211-224): restatement vs the live reference."""
    g = np.load(os.path.join(golden_dir, "layer2_v2vnet_det_seed18.npz"))
    sd = synth.v2vnet_det_state(18, layer_channel=128)
    bevs, trans, nat = synth.make_scene(1, 5, 18, present=[4])
    with torch.no_grad():
        r = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=1, agent_num=5, gnn_iter=2, layer=2)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)
    g = np.load(os.path.join(golden_dir, "layer4_v2vnet_det_seed47.npz"))       # 512 channels at 16 x 16
    bevs, trans, nat = synth.make_scene(1, 5, 47, present=[4])
    with torch.no_grad():
        r = restate.v2vnet_det_forward(bevs, trans, nat, synth.v2vnet_det_state(47, layer_channel=512), batch_size=1,
                                       agent_num=5, gnn_iter=2, layer=4)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)
    g = np.load(os.path.join(golden_dir, "layer4_sum_det_seed49.npz"))
    bevs, trans, nat = synth.make_scene(1, 5, 49, present=[3])
    with torch.no_grad():
        r = restate.fusion_det_forward("sum", bevs, trans, nat, synth.fusion_det_state("sum", 49), batch_size=1, agent_num=5,
                                       layer=4)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)
    g = np.load(os.path.join(golden_dir, "layer2_disco_det_seed19.npz"))
    bevs, trans, nat = synth.make_scene(1, 5, 19)
    with torch.no_grad():
        r = restate.fusion_det_forward("disco", bevs, trans, nat, synth.fusion_det_state("disco", 19, channel=128),
                                       batch_size=1, agent_num=5, layer=2)
    _check("loc", r["loc"], g)
    g = np.load(os.path.join(golden_dir, "layer1_max_det_seed20.npz"))
    bevs, trans, nat = synth.make_scene(1, 5, 20, present=[3])
    with torch.no_grad():
        r = restate.fusion_det_forward("max", bevs, trans, nat, synth.fusion_det_state("max", 20), batch_size=1,
                                       agent_num=5, layer=1)
    _check("loc", r["loc"], g)
    _check("cls", r["cls"], g)


def _train_cases():
    from oracle.gen_golden import TRAIN_CASES
    return [(tag, kind, seed) for tag, (kind, seed) in TRAIN_CASES.items()]


@pytest.mark.parametrize("tag,kind,seed", _train_cases(), ids=[c[0] for c in _train_cases()])
def test_train_step_oracle(golden_dir, tag, kind, seed):
    """SURVEY 8(f1) oracle pin: the restatement in training mode (BatchNorm batch statistics + running-buffer updates)
    and its autograd vector-Jacobian product vs one training step of the LIVE reference module in .train() mode, for
    det V2VNet / FaFNet / When2com (training=True) / DiscoNet and seg UNet / V2VNet.
    Both sides run in float64 (fixtures made under oracle.ref_loader.float64_shim): in float32 the BatchNorm backward of
    these seeded random nets amplifies summation-order noise to ~1% per gradient element, which would hide a wrong
    algorithm; in float64 the outputs, the gradient of EVERY parameter that receives one and every BatchNorm running
    buffer must agree to 1e-6 relative (measured 1e-12 .. 2e-8; the when2com path has one float32 constant)."""
    from oracle.gen_golden import grad_stride, make_upstream, train_case
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    sd, inputs, keys = train_case(kind, seed)
    sd = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    wrap = lambda t: t if isinstance(t, dict) else {"logits": t}   # noqa: E731
    fwd = {
        "v2vnet": lambda w: restate.v2vnet_det_forward(*inputs, w, batch_size=1, agent_num=5, gnn_iter=3),
        "v2vnet_c2": lambda w: restate.v2vnet_det_forward(*inputs, w, batch_size=1, agent_num=5, gnn_iter=3, compress_level=2),
        "seg_unet_c3": lambda w: wrap(restate.seg_unet_forward(inputs[0], w)),
        "fafnet": lambda w: restate.fafnet_forward(inputs[0], w),
        "when2com": lambda w: restate.when2com_det_forward(*inputs, w, batch_size=1, agent_num=5, warp_flag=1, training=True),
        "disco": lambda w: restate.fusion_det_forward("disco", *inputs, w, batch_size=1, agent_num=5),
        "cat": lambda w: restate.fusion_det_forward("cat", *inputs, w, batch_size=1, agent_num=5),
        "agent": lambda w: restate.fusion_det_forward("agent", *inputs, w, batch_size=1, agent_num=5),
        "mean": lambda w: restate.fusion_det_forward("mean", *inputs, w, batch_size=1, agent_num=5),
        "sum": lambda w: restate.fusion_det_forward("sum", *inputs, w, batch_size=1, agent_num=5),
        "max": lambda w: restate.fusion_det_forward("max", *inputs, w, batch_size=1, agent_num=5),
        "seg_when2com": lambda w: wrap(restate.seg_when2com_forward(*inputs, w, agent_num=5, warp_flag=1, training=True)),
        "seg_unet": lambda w: wrap(restate.seg_unet_forward(inputs[0], w)),
        "seg_v2vnet": lambda w: wrap(restate.seg_v2vnet_forward(*inputs, w, agent_num=5)),
        "seg_mean": lambda w: wrap(restate.seg_fusion_forward("mean", *inputs, w, agent_num=5)),
        "seg_max": lambda w: wrap(restate.seg_fusion_forward("max", *inputs, w, agent_num=5)),
        "seg_cat": lambda w: wrap(restate.seg_fusion_forward("cat", *inputs, w, agent_num=5)),
        "seg_agent": lambda w: wrap(restate.seg_fusion_forward("agent", *inputs, w, agent_num=5)),
        "seg_disco": lambda w: wrap(restate.seg_fusion_forward("disco", *inputs, w, agent_num=5)),
    }[kind]
    out, grads, after = restate.train_step_vjp(fwd, sd, make_upstream({k: g[k + ".shape"] for k in keys}, seed))
    for name in keys:
        assert out[name].dtype == torch.float64 and list(out[name].shape) == list(g[name + ".shape"])
        sub = out[name].contiguous().view(-1)[::STRIDE].numpy()
        assert np.abs(sub - g[name + ".sub"]).max() < 1e-6 * np.abs(g[name + ".sub"]).max(), name
    ref_grads = sorted(k[5:-4] for k in g.files if k.startswith("grad.") and k.endswith(".sub"))
    assert len(ref_grads) > (50 if kind.startswith("seg_") else 70) and set(ref_grads) == set(grads), set(ref_grads) ^ set(grads)
    for k in ref_grads:
        gr, ref = grads[k], g["grad." + k + ".sub"]
        sub = gr.reshape(-1)[::grad_stride(gr.numel())].numpy()
        # (a conv bias feeding a train-mode BatchNorm has an exactly-zero gradient: both sides hold 1e-16 round-off)
        assert np.abs(sub - ref).max() < 1e-6 * np.abs(ref).max() + 1e-11, k
        assert abs(gr.norm().item() - float(g["grad." + k + ".norm"])) < 1e-6 * float(g["grad." + k + ".norm"]) + 1e-11, k
    bn = [k[3:] for k in g.files if k.startswith("bn.")]
    assert len(bn) >= 20
    for k in bn:
        assert np.abs(after[k].numpy() - g["bn." + k]).max() < 1e-8, k
    assert restate._TRAIN is False      # the switch is restored: eval-mode behaviour is untouched
