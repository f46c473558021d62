"""GPU parity of the intermediate-fusion baselines (SURVEY 8(f4) / 8(b) module list): Mean / Max / Sum / Cat /
AgentWiseWeighted fusion and DiscoNet, det and seg, plus TeacherNet -- ops against torch, whole models through the
drop-in ``coperception.models.{det,seg}`` classes against the CPU oracle and the live-reference fixtures.
Tolerance: 1e-3 relative (max-abs error / max-abs value) for bf16x3; argmax identical outside the error margin."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _to_act(x_nchw, planes=2):
    from v2x_b200 import ops
    return ops.pack_input_nchw(x_nchw.cuda().contiguous(), x_nchw.shape[1], planes)


def _agent_major(local):
    """[B,A,C,H,W] -> [A*B,C,H,W] (agent-major, DetModelBase.agents_to_batch without the flip)."""
    return torch.cat([local[:, i] for i in range(local.shape[1])], 0)


def _scene(batch, agents, c, seed, present):
    from oracle import synth
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((batch * agents, c, 32, 32), generator=g)
    trans = synth.make_trans_matrices(batch, agents, seed, present=present)
    nat = torch.full((batch, agents), agents, dtype=torch.int64)
    if present is not None:
        for b in range(batch):
            nat[b, :] = present[b]
    return x, trans, nat


@pytest.mark.parametrize("c", [64, 512])
@pytest.mark.parametrize("mode", ["mean", "sum", "max"])
@pytest.mark.parametrize("only_v2i", [False, True])
def test_warp_reduce(mode, c, only_v2i):
    """v2x_warp_reduce_fwd vs the oracle's flip -> warp list -> reduce -> flip (FusionBase.py:41-63)."""
    from oracle import restate
    from v2x_b200 import ops
    batch, agents = 2, 5
    x, trans, nat = _scene(batch, agents, c, 3, [5, 3])
    ref = restate.fusion_stage(mode, x, trans, nat, {}, batch, agents, only_v2i=only_v2i)
    out = ops.act_to_float(ops.warp_reduce(_to_act(x), trans.cuda(), nat.cuda(), batch, agents, mode, only_v2i=only_v2i))
    assert rel_err(out, ref) < 2e-4
    # absent agent slots keep their own map bit-for-bit at bf16x3 precision
    absent = [batch * i + 1 for i in range(3, agents)]
    assert rel_err(out[absent], x[absent]) < 1e-4


@pytest.mark.parametrize("kind", ["disco", "agent"])
def test_pair_weight_fuse(kind):
    """pair_score (+ agent_softmax) + warp_weighted vs the oracle's fuse_rule on random layer maps (C = 256)."""
    from oracle import restate, synth
    from v2x_b200 import nets, ops
    batch, agents, c = 2, 5, 256
    x, trans, nat = _scene(batch, agents, c, 4, [4, 5])
    x = x.abs()  # post-ReLU-like features
    sd = synth.fusion_det_state(kind, 21)
    with torch.no_grad():
        ref = restate.fusion_stage(kind, x, trans, nat, sd, batch, agents)

    class P:  # a bare launch recorder with the plan interface FuseStage needs
        device, planes = torch.device("cuda"), 2

        def __init__(self):
            self.launches = []

        def act(self, name, h, w, ch):
            return ops.empty_act(2, batch * agents, h, w, ch, self.device)

        def add(self, l):
            self.launches.append(l)

    p = P()
    st = nets.FuseStage(p, kind, sd, _to_act(x), trans.cuda(), nat.cuda(), batch, agents)
    for l in p.launches:
        l()
    torch.cuda.synchronize()
    e = rel_err(ops.act_to_float(st.out), ref)
    print("fuse stage %s rel_err %.3e" % (kind, e))
    assert e < 5e-4


def _flips(out_cls, ref_cls):
    out_cls, ref_cls = out_cls.detach().float().cpu(), ref_cls.detach().float().cpu()
    err = (out_cls - ref_cls).abs().max().item()
    flip = out_cls.argmax(-1) != ref_cls.argmax(-1)
    margin = (ref_cls[..., 0] - ref_cls[..., 1]).abs()
    return int(flip.sum()), int((flip & (margin > 2 * err)).sum())


def _det_fixtures():
    from oracle.gen_golden import FUSION_FIXTURES
    return [f for f in FUSION_FIXTURES if f[1] == "det"]


def _seg_fixtures():
    from oracle.gen_golden import FUSION_FIXTURES
    return [f for f in FUSION_FIXTURES if f[1] == "seg"]


DET_CLASSES = {"mean": "MeanFusion", "max": "MaxFusion", "sum": "SumFusion", "cat": "CatFusion",
               "agent": "AgentWiseWeightedFusion", "disco": "DiscoNet"}


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("fx", _det_fixtures(), ids=lambda f: f[0])
def test_fusion_det_models(fx, precision, golden_dir):
    import coperception.models.det as det
    from oracle import restate, synth
    from oracle.gen_golden import STRIDE
    from v2x_b200 import default_det_config
    tag, _, kind, batch, seed, present, v2i = fx
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    sd = synth.fusion_det_state(kind, seed)
    bevs, trans, nat = synth.make_scene(batch, 5, seed, present=present)
    with torch.no_grad():
        ref = restate.fusion_det_forward(kind, bevs, trans, nat, sd, batch_size=batch, agent_num=5, only_v2i=v2i,
                                         stages=True)
    model = getattr(det, DET_CLASSES[kind])(default_det_config(), layer=3, kd_flag=0, num_agent=5, only_v2i=v2i)
    model.load_state_dict(sd, strict=True)
    model.precision = precision
    model = model.cuda().eval()
    with torch.no_grad():
        out = model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=batch)
    if kind == "disco":
        out, weights = out
        assert len(weights) == int(nat[:, 0].sum()) and weights[0][0].shape == (32, 32)
    torch.cuda.synchronize()
    tol = 1e-3 if precision == "bf16x3" else 5e-2
    for k in ("loc", "cls"):
        sub = out[k].detach().float().cpu().contiguous().view(-1)[::STRIDE].numpy()
        eg = float(np.abs(sub - g[k + ".sub"]).max() / np.abs(g[k + ".sub"]).max())
        e = rel_err(out[k], ref[k])
        print("fusion det %s %s %s rel_err=%.3e golden=%.3e" % (tag, precision, k, e, eg))
        assert out[k].shape == ref[k].shape and e < tol and eg < tol
    flips, bad = _flips(out["cls"], ref["cls"])
    print("fusion det %s %s argmax flips %d (outside margin %d)" % (tag, precision, flips, bad))
    assert bad == 0
    if precision == "bf16x3":
        assert flips <= 1e-4 * ref["cls"].numel() / 2
        # kd_flag == 1 output tuple: (result, x_8, x_7, x_6, x_5, fused) (FusionBase.py:72-73)
        model.kd_flag = 1
        with torch.no_grad():
            r, x8, x7, x6, x5, fused = model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=batch)
        for name, t, rt in (("x8", x8, ref["dec"][0]), ("x7", x7, ref["dec"][1]), ("x6", x6, ref["dec"][2]),
                            ("x5", x5, ref["dec"][3]), ("fused", fused, ref["fused"])):
            assert t.shape == rt.shape and rel_err(t, rt) < 1e-3, name


@pytest.mark.parametrize("fx", _seg_fixtures(), ids=lambda f: f[0])
def test_fusion_seg_models(fx, golden_dir):
    import coperception.models.seg as seg
    from oracle import restate, synth
    from oracle.gen_golden import STRIDE
    tag, _, kind, batch, seed, present, v2i = fx
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    sd = synth.seg_fusion_state(kind, seed)
    x, trans, nat = synth.make_seg_scene(batch, 5, seed, present=present)
    with torch.no_grad():
        ref = restate.seg_fusion_forward(kind, x, trans, nat, sd, agent_num=5, only_v2i=v2i, stages=True)
    cls = getattr(seg, DET_CLASSES[kind])
    if kind == "disco":
        model = cls(13, 8, 5, kd_flag=False, only_v2i=v2i)
    else:
        model = cls(13, 8, 5, 0, v2i)
    model.load_state_dict(sd, strict=True)
    model.precision = "bf16x3"
    model = model.cuda().eval()
    with torch.no_grad():
        out = model(x.cuda(), trans.cuda(), nat.cuda())
    torch.cuda.synchronize()
    sub = out.detach().float().cpu().contiguous().view(-1)[::STRIDE].numpy()
    eg = float(np.abs(sub - g["logits.sub"]).max() / np.abs(g["logits.sub"]).max())
    e = rel_err(out, ref["logits"])
    print("fusion seg %s rel_err=%.3e golden=%.3e" % (tag, e, eg))
    assert out.shape == ref["logits"].shape and e < 1e-3 and eg < 1e-3
    o, r = out.float().cpu(), ref["logits"]
    err = (o - r).abs().max().item()
    flip = o.argmax(1) != r.argmax(1)
    top2 = r.topk(2, dim=1).values
    assert int((flip & ((top2[:, 0] - top2[:, 1]) > 2 * err)).sum()) == 0
    model.kd_flag = True
    with torch.no_grad():
        tup = model(x.cuda(), trans.cuda(), nat.cuda())
    assert len(tup) == 7 and rel_err(tup[6], ref["fused"]) < 1e-3


def test_teacher_net(golden_dir):
    from coperception.models.det import TeacherNet
    from oracle import restate, synth
    from v2x_b200 import default_det_config
    g = np.load(os.path.join(golden_dir, "teacher_n1_seed3.npz"))
    n, seed = [int(v) for v in g["meta"]]
    sd = synth.fafnet_state(seed)
    bevs = synth.make_bevs(n, seed)
    with torch.no_grad():
        ref = restate.teacher_forward(bevs, sd)
    m = TeacherNet(default_det_config())
    m.load_state_dict(sd, strict=True)
    m.precision = "bf16x3"
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(bevs.cuda())
    assert len(out) == 6
    for name, t, r in zip(("x8", "x7", "x6", "x5", "x3", "x4"), out, ref):
        assert t.shape == r.shape and rel_err(t, r) < 1e-3, name


def test_compress_level_models(golden_dir):
    """compress_level > 0 (optional compress / decompress of the communicated layer): V2VNet det level 2,
    DiscoNet det level 6 (4 channels, zero-padded operand), seg UNet level 3 with its kd_flag tuple -- vs oracle + fixtures."""
    import coperception.models.det as det
    import coperception.models.seg as seg
    from oracle import restate, synth
    from oracle.gen_golden import STRIDE
    from v2x_b200 import default_det_config

    def golden_err(t, g, name):
        sub = t.detach().float().cpu().contiguous().view(-1)[::STRIDE].numpy()
        return float(np.abs(sub - g[name + ".sub"]).max() / np.abs(g[name + ".sub"]).max())

    # V2VNet det
    g = np.load(os.path.join(golden_dir, "compress_v2vnet_det_l2_seed15.npz"))
    sd = synth.v2vnet_det_state(15, compress_level=2)
    bevs, trans, nat = synth.make_scene(1, 5, 15, present=[4])
    with torch.no_grad():
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=1, agent_num=5, gnn_iter=3, compress_level=2)
    m = det.V2VNet(default_det_config(), 3, 3, 256, num_agent=5, compress_level=2)
    m.load_state_dict(sd, strict=True)
    m.precision = "bf16x3"
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
    for k in ("loc", "cls"):
        assert rel_err(out[k], ref[k]) < 1e-3 and golden_err(out[k], g, k) < 1e-3, k
    # DiscoNet det, 4 compressed channels
    g = np.load(os.path.join(golden_dir, "compress_disco_det_l6_seed16.npz"))
    sd = synth.fusion_det_state("disco", 16, compress_level=6)
    bevs, trans, nat = synth.make_scene(1, 5, 16)
    with torch.no_grad():
        ref = restate.fusion_det_forward("disco", bevs, trans, nat, sd, batch_size=1, agent_num=5)
    m = det.DiscoNet(default_det_config(), layer=3, kd_flag=0, num_agent=5, compress_level=6)
    m.load_state_dict(sd, strict=True)
    m.precision = "bf16x3"
    m = m.cuda().eval()
    with torch.no_grad():
        out, _ = m(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
    for k in ("loc", "cls"):
        assert rel_err(out[k], ref[k]) < 1e-3 and golden_err(out[k], g, k) < 1e-3, k
    # seg UNet level 3, kd tuple
    g = np.load(os.path.join(golden_dir, "compress_seg_unet_l3_kd_seed17.npz"))
    sd = synth.seg_unet_state(17, compress_level=3)
    x, _, _ = synth.make_seg_scene(1, 2, 17)
    m = seg.UNet(13, 8, kd_flag=True, compress_level=3)
    m.load_state_dict(sd, strict=True)
    m.precision = "bf16x3"
    m = m.cuda().eval()
    with torch.no_grad():
        tup = m(x.cuda())
    assert len(tup) == 7
    for name, t in zip(("logits", "x9", "x8", "x7", "x6", "x5", "x4"), tup):
        assert list(t.shape) == list(g[name + ".shape"]) and golden_err(t, g, name) < 1e-3, name


def test_other_communication_layers(golden_dir):
    """communication at encoder layers other than 3 -- V2VNet at layer 2 (128 ch @ 64x64, two GNN rounds),
    DiscoNet at layer 2, MaxFusion at layer 1 (64 ch @ 128x128) -- vs oracle and live-reference fixtures."""
    import coperception.models.det as det
    from oracle import restate, synth
    from oracle.gen_golden import STRIDE
    from v2x_b200 import default_det_config

    def check(out, ref, g, what):
        for k in ("loc", "cls"):
            sub = out[k].detach().float().cpu().contiguous().view(-1)[::STRIDE].numpy()
            eg = float(np.abs(sub - g[k + ".sub"]).max() / np.abs(g[k + ".sub"]).max())
            e = rel_err(out[k], ref[k])
            print("layers %s %s rel_err=%.3e golden=%.3e" % (what, k, e, eg))
            assert e < 1e-3 and eg < 1e-3, (what, k)

    g = np.load(os.path.join(golden_dir, "layer2_v2vnet_det_seed18.npz"))
    sd = synth.v2vnet_det_state(18, layer_channel=128)
    bevs, trans, nat = synth.make_scene(1, 5, 18, present=[4])
    with torch.no_grad():
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=1, agent_num=5, gnn_iter=2, layer=2)
    m = det.V2VNet(default_det_config(), 2, 2, 128, num_agent=5)
    m.load_state_dict(sd, strict=True)
    m.precision = "bf16x3"
    m = m.cuda().eval()
    with torch.no_grad():
        check(m(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1), ref, g, "v2vnet@2")
    m.precision = "bf16"
    m.invalidate()
    with torch.no_grad():
        out = m(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
    assert rel_err(out["loc"], ref["loc"]) < 5e-2

    g = np.load(os.path.join(golden_dir, "layer2_disco_det_seed19.npz"))
    sd = synth.fusion_det_state("disco", 19, channel=128)
    bevs, trans, nat = synth.make_scene(1, 5, 19)
    with torch.no_grad():
        ref = restate.fusion_det_forward("disco", bevs, trans, nat, sd, batch_size=1, agent_num=5, layer=2)
    m = det.DiscoNet(default_det_config(), layer=2, kd_flag=0, num_agent=5)
    m.load_state_dict(sd, strict=True)
    m.precision = "bf16x3"
    m = m.cuda().eval()
    with torch.no_grad():
        out, w = m(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
    check(out, ref, g, "disco@2")
    assert w[0][0].shape == (64, 64)

    g = np.load(os.path.join(golden_dir, "layer1_max_det_seed20.npz"))
    sd = synth.fusion_det_state("max", 20)
    bevs, trans, nat = synth.make_scene(1, 5, 20, present=[3])
    with torch.no_grad():
        ref = restate.fusion_det_forward("max", bevs, trans, nat, sd, batch_size=1, agent_num=5, layer=1)
    m = det.MaxFusion(default_det_config(), layer=1, kd_flag=0, num_agent=5)
    m.load_state_dict(sd, strict=True)
    m.precision = "bf16x3"
    m = m.cuda().eval()
    with torch.no_grad():
        check(m(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1), ref, g, "max@1")
