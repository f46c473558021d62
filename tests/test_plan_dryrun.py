"""CPU dry run of the host logic above the C ABI: every plan constructor and every training tape is executed on CPU
tensors against a FAKE library whose entry points do nothing but validate their arguments.

What this checks without a GPU: that each plan / train step issues a launch list at all (shapes, channel padding, the
asserts in ops.ConvLaunch, workspace wiring of every option: communication layer, compress_level, has_query, kd outputs,
input modes), that every ctypes call matches the arity and argument types declared in v2x_b200/_lib.py::SYMBOLS (each fake
entry point is a ctypes callback with the declared signature, so a wrong argument count or type raises exactly like the
real foreign function would), and that every parameter the reference gives a gradient gets a gradient tensor of the
right shape from the training tapes.  It does NOT check arithmetic: no kernel runs, the product path still has no CPU
fallback (the fake lives here, under tests/, and is installed by monkeypatching for the duration of one test).
"""
import ctypes as C

import pytest
import torch

from oracle import synth
from v2x_b200 import _lib


class FakeLib:
    """Entry points with the signatures of include/v2x_b200.h that only count calls."""

    def __init__(self):
        self.calls = {}
        self._keep = []
        for name, res, args in _lib.SYMBOLS:
            if res is C.c_char_p or not args:
                continue
            proto = C.CFUNCTYPE(res, *args)

            def body(*a, _n=name):
                self.calls[_n] = self.calls.get(_n, 0) + 1
                return 0
            fn = proto(body)
            self._keep.append(fn)
            setattr(self, name, fn)

    def v2x_version(self):
        return 1

    def v2x_device_ok(self):
        return 1

    def v2x_last_error(self):
        return b"fake"


@pytest.fixture
def fake(monkeypatch):
    from v2x_b200 import ops, train
    lib = FakeLib()
    monkeypatch.setattr(ops, "require_gpu", lambda: lib)
    monkeypatch.setattr(ops, "_stream", lambda: C.c_void_p(0))
    monkeypatch.setattr(train, "_stream", lambda: C.c_void_p(0))
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)
    return lib


def _issue(plan):
    assert plan.launches, "empty launch list"
    for launch in plan.launches:
        launch()


def test_fake_library_rejects_a_wrong_call(fake):
    with pytest.raises((TypeError, C.ArgumentError)):
        fake.v2x_maxpool2_fwd(None, None, 1, 2, 3)           # too few arguments
    with pytest.raises((TypeError, C.ArgumentError)):
        fake.v2x_maxpool2_fwd(None, None, 1.5, 2, 3, 4, 5, None)    # a float where the header declares int32_t


@pytest.mark.parametrize("kw", [dict(), dict(layer=2), dict(layer=1), dict(layer=4), dict(layer=4, planes="bf16"), dict(compress=2),
                                dict(compress=6), dict(input_mode="u8"),
                                dict(input_mode="voxels"), dict(only_v2i=True, planes="bf16"), dict(planes="fp16x3")],
                         ids=lambda k: "-".join("%s=%s" % kv for kv in k.items()) or "default")
def test_v2vnet_det_plan_wiring(fake, kw):
    from v2x_b200 import nets
    kw = dict(kw)
    compress, layer = kw.pop("compress", 0), kw.get("layer", 3)
    sd = synth.v2vnet_det_state(1, layer_channel=(32, 64, 128, 256, 512)[layer], compress_level=compress)
    plan = nets.V2VNetDetPlan(sd, 2, 5, device="cpu", **dict(dict(planes="mixed"), **kw))
    _issue(plan)
    assert fake.calls["v2x_conv_fwd"] >= 25 and fake.calls["v2x_warp_mean_fwd"] == 1
    out = plan.result()
    assert tuple(out["loc"].shape) == (10, 256, 256, 6, 1, 6) and tuple(out["cls"].shape) == (10, 256 * 256 * 6, 2)
    assert ("x3d" in plan.ws) == (compress > 0)
    if layer == 4:     # the GNN block runs on the plain 16x16 x_4; the decoder reads the fused map through the upsampling store
        assert tuple(plan.ws["mean"].shape[2:]) == (16, 16, 512) and tuple(plan.ws["x4u"].shape[2:]) == (32, 32, 512)
        assert not plan.gru_h.gru_pre_act and plan.side_lo < 0


def test_fafnet_and_teacher_plan_wiring(fake):
    from v2x_b200 import nets
    plan = nets.FaFNetPlan(synth.fafnet_state(2), 3, planes="mixed", device="cpu")
    _issue(plan)
    assert tuple(plan.result()["cls"].shape) == (3, 256 * 256 * 6, 2)
    plan = nets.FaFNetPlan(synth.fafnet_state(2, compress_level=3), 2, planes="mixed", device="cpu", heads=False)
    _issue(plan)
    assert "x3d" in plan.ws and not hasattr(plan, "cls")


@pytest.mark.parametrize("kw", [dict(), dict(inference="softmax"), dict(inference="argmax_test", warp_flag=0),
                                dict(training_pass_only=True), dict(has_query=False), dict(layer=2), dict(layer=4),
                                dict(layer=2, has_query=False, inference="softmax", warp_flag=0)],
                         ids=lambda k: "-".join("%s=%s" % kv for kv in k.items()) or "default")
def test_when2com_det_plan_wiring(fake, kw):
    from v2x_b200 import nets
    has_query = kw.get("has_query", True)
    sd = synth.when2com_det_state(3, has_query=has_query)
    plan = nets.When2comDetPlan(sd, 1, 5, planes="mixed", device="cpu", **kw)
    _issue(plan)
    two_pass = not kw.get("training_pass_only") and kw.get("inference", "activated") != "softmax"
    assert fake.calls["v2x_warp_gated_fwd"] == (2 if two_pass else 1)
    assert fake.calls["v2x_linear_fwd"] == (6 if has_query else 3)       # key MLP (+ query MLP)
    layer = kw.get("layer", 3)
    assert tuple(plan.ws["fuse1"].shape[2:]) == {2: (64, 64, 128), 3: (32, 32, 256), 4: (16, 16, 512)}[layer]
    if layer == 4:
        assert tuple(plan.ws["fuse1u"].shape[2:]) == (32, 32, 512) and tuple(plan.ws["fuse2u"].shape[2:]) == (32, 32, 512)
    if not has_query:
        assert bool((plan.querys == 1).all()) and tuple(plan.querys.shape) == (5, 32)


def test_when2com_det_plan_refuses_what_the_reference_cannot_run(fake):
    from v2x_b200 import nets, ops
    sd = synth.when2com_det_state(3)
    with pytest.raises(ops.V2XError):
        nets.When2comDetPlan(sd, 1, 5, device="cpu", layer=2, inference="argmax_test")
    with pytest.raises(ops.V2XError):
        nets.When2comDetPlan(sd, 1, 5, device="cpu", layer=4, inference="argmax_test")
    with pytest.raises(ops.V2XError):
        nets.When2comDetPlan(sd, 1, 5, device="cpu", layer=1)
    with pytest.raises(ops.V2XError):      # the ConvGRU tile needs a multiple of 64 channels: layer 0 has 32
        nets.V2VNetDetPlan(synth.v2vnet_det_state(3, layer_channel=32), 1, 5, device="cpu", layer=0)


@pytest.mark.parametrize("kind", ["mean", "sum", "max", "cat", "agent", "disco"])
def test_fusion_det_plan_wiring(fake, kind):
    from v2x_b200 import nets
    sd = synth.fusion_det_state(kind, 4)
    plan = nets.FusionDetPlan(sd, kind, 1, 5, planes="mixed", device="cpu")
    _issue(plan)
    assert tuple(plan.fused.shape) == (2, 5, 32, 32, 256)
    if kind in ("mean", "sum", "max"):      # parameter-free rules also run at the other layers
        plan = nets.FusionDetPlan(sd, kind, 1, 5, planes="mixed", device="cpu", layer=1)
        _issue(plan)
        assert tuple(plan.fused.shape) == (2, 5, 128, 128, 64)
        plan = nets.FusionDetPlan(sd, kind, 1, 5, planes="mixed", device="cpu", layer=4)
        _issue(plan)
        assert tuple(plan.fused.shape) == (2, 5, 16, 16, 512) and tuple(plan.ws["x4u"].shape) == (2, 5, 32, 32, 512)


def test_seg_plan_wiring(fake):
    from v2x_b200 import nets_seg
    plan = nets_seg.SegUNetPlan(synth.seg_unet_state(5, compress_level=3), 2, planes="mixed", device="cpu")
    _issue(plan)
    assert tuple(plan.logits.shape) == (2, 8, 256, 256)
    plan = nets_seg.SegV2VNetPlan(synth.seg_v2vnet_state(5), 1, 5, planes="mixed", device="cpu")
    _issue(plan)
    for hq in (True, False):
        plan = nets_seg.SegWhen2comPlan(synth.seg_when2com_state(5, has_query=hq), 1, 5, planes="mixed", device="cpu",
                                        has_query=hq)
        _issue(plan)
    for kind in ("mean", "max", "sum", "cat", "agent", "disco"):
        plan = nets_seg.SegFusionPlan(synth.seg_fusion_state(kind, 5), kind, 1, 5, planes="mixed", device="cpu")
        _issue(plan)
        assert tuple(plan.logits.shape) == (5, 8, 256, 256)


# ---- training tapes -------------------------------------------------------------------------------------------------
def _grads_cover(module, grads, expect_none=()):
    names = [k for k, _ in module.named_parameters()]
    assert len(grads) == len(names)
    n = 0
    for k, p, g in zip(names, module.parameters(), grads):
        if g is None:
            assert k.startswith(tuple(expect_none)), "no gradient for %s" % k
            continue
        assert tuple(g.shape) == tuple(p.shape), k
        n += 1
    return n


def _det_step(fn, module, args, n_maps):
    """Run an autograd.Function train step forward + backward on CPU tensors through the fake library."""
    params = [p.detach().requires_grad_(True) for p in module.parameters()]
    outs = fn.apply(module, *args, *params)
    loc, cls = outs[0], outs[1]
    assert tuple(loc.shape) == (n_maps, 256, 256, 6, 1, 6) and tuple(cls.shape) == (n_maps, 256 * 256 * 6, 2)
    torch.autograd.backward([loc, cls], [torch.full_like(loc, 1e-3), torch.full_like(cls, 1e-3)])
    return [p.grad for p in params]


@pytest.mark.parametrize("compress", [0, 2, 3])
def test_v2vnet_train_tape_wiring(fake, compress):
    from coperception.models.det import V2VNet
    from v2x_b200 import default_det_config
    from v2x_b200.train import V2VNetTrainStep
    m = V2VNet(default_det_config(), 3, 3, 256, num_agent=5, compress_level=compress).train()
    bevs, trans, nat = synth.make_scene(1, 5, 6)
    grads = _det_step(V2VNetTrainStep, m, (bevs, trans, nat, 1), 5)
    # the unused halves of the two Backbones (decoder convs of u_encoder, encoder convs of decoder) and W_hh get none
    n = _grads_cover(m, grads, expect_none=("u_encoder.conv5", "u_encoder.bn5", "u_encoder.conv6", "u_encoder.bn6",
                                            "u_encoder.conv7", "u_encoder.bn7", "u_encoder.conv8", "u_encoder.bn8",
                                            "decoder.conv_pre", "decoder.bn_pre", "decoder.conv1", "decoder.bn1",
                                            "decoder.conv2", "decoder.bn2", "decoder.conv3", "decoder.bn3", "decoder.conv4",
                                            "decoder.bn4", "convgru.weight_hh"))
    names = [k for k, _ in m.named_parameters()]
    got = {k for k, g in zip(names, grads) if g is not None}
    assert n >= 96 and ("u_encoder.com_compresser.weight" in got) == (compress > 0)
    if compress:
        assert {"u_encoder.bn_compress.weight", "u_encoder.com_decompresser.weight", "u_encoder.bn_decompress.bias"} <= got
        assert int(m.u_encoder.bn_compress.num_batches_tracked) == 1
    assert fake.calls["v2x_conv_wgrad_tc"] + fake.calls.get("v2x_conv_wgrad", 0) >= 30


def test_train_tape_refuses_narrow_compression(fake):
    from coperception.models.det import V2VNet
    from v2x_b200 import default_det_config
    from v2x_b200.train import V2VNetTrainStep
    m = V2VNet(default_det_config(), 3, 3, 256, num_agent=5, compress_level=4).train()    # 16 compressed channels
    bevs, trans, nat = synth.make_scene(1, 5, 6)
    with pytest.raises(NotImplementedError):
        V2VNetTrainStep.apply(m, bevs, trans, nat, 1, *m.parameters())


def test_fafnet_and_when2com_train_tape_wiring(fake):
    from coperception.models.det import FaFNet, When2com
    from v2x_b200 import default_det_config
    from v2x_b200.train import FaFNetTrainStep, When2comTrainStep
    m = FaFNet(default_det_config(), kd_flag=0, num_agent=5, compress_level=1).train()
    grads = _det_step(FaFNetTrainStep, m, (synth.make_bevs(2, 7),), 2)
    assert _grads_cover(m, grads) == len(grads)          # FaFNet uses every parameter of its one Backbone
    m = When2com(default_det_config(), layer=3, warp_flag=1, num_agent=5).train()
    bevs, trans, nat = synth.make_scene(1, 5, 7)
    grads = _det_step(When2comTrainStep, m, (bevs, trans, nat, 1), 5)
    names = [k for k, _ in m.named_parameters()]
    got = {k for k, g in zip(names, grads) if g is not None}
    assert {"key_net.fc.0.weight", "query_net.fc.4.bias", "attention_net.linear.weight",
            "query_key_net.conv5.cbr_unit.0.weight", "query_key_net.lidar_encoder.conv4_2.weight"} <= got


@pytest.mark.parametrize("kind", ["mean", "max", "cat", "agent", "disco"])
def test_fusion_train_tape_wiring(fake, kind):
    from coperception.models import det as det_models
    from v2x_b200 import default_det_config
    from v2x_b200.train import FusionTrainStep
    cls = {"mean": "MeanFusion", "max": "MaxFusion", "cat": "CatFusion", "agent": "AgentWiseWeightedFusion",
           "disco": "DiscoNet"}[kind]
    m = getattr(det_models, cls)(default_det_config(), layer=3, kd_flag=0, num_agent=5, compress_level=2).train()
    bevs, trans, nat = synth.make_scene(1, 5, 8, present=[4])
    grads = _det_step(FusionTrainStep, m, (kind, bevs, trans, nat, 1), 5)
    names = [k for k, _ in m.named_parameters()]
    got = {k for k, g in zip(names, grads) if g is not None}
    assert "u_encoder.com_decompresser.weight" in got and "decoder.conv8_2.weight" in got


def test_train_tape_option_combinations(fake):
    """kd_flag = 1 outputs with upstream gradients, only_v2i, when2com without the warp and with compression."""
    from coperception.models import det as det_models
    from v2x_b200 import default_det_config
    from v2x_b200.train import FusionTrainStep, When2comTrainStep
    cfg = default_det_config()
    bevs, trans, nat = synth.make_scene(1, 5, 8, present=[4])
    for cls, kind in (("MaxFusion", "max"), ("DiscoNet", "disco")):
        m = getattr(det_models, cls)(cfg, layer=3, kd_flag=1, num_agent=5, only_v2i=True).train()
        params = [p.detach().requires_grad_(True) for p in m.parameters()]
        outs = FusionTrainStep.apply(m, kind, bevs, trans, nat, 1, *params)
        assert len(outs) == 7      # loc, cls, x_8, x_7, x_6, x_5, fused (FusionBase.py:72-73)
        assert [tuple(o.shape[1:]) for o in outs[2:]] == [(32, 256, 256), (64, 128, 128), (128, 64, 64), (256, 32, 32),
                                                          (256, 32, 32)]
        torch.autograd.backward(list(outs), [torch.full_like(o, 1e-3) for o in outs])
        assert params[0].grad is not None
    m = det_models.When2com(cfg, layer=3, warp_flag=0, num_agent=5, compress_level=2).train()
    grads = _det_step(When2comTrainStep, m, (bevs, trans, nat, 1), 5)
    names = [k for k, _ in m.named_parameters()]
    got = {k for k, g in zip(names, grads) if g is not None}
    assert {"u_encoder.com_compresser.weight", "attention_net.linear.bias"} <= got


@pytest.mark.parametrize("case", ["unet", "unet_c3", "v2v", "when2com", "mean", "cat"])
def test_seg_train_tape_wiring(fake, case):
    from coperception.models import seg as seg_models
    from v2x_b200 import default_det_config
    from v2x_b200.train import SegTrainStep
    x, trans, nat = synth.make_seg_scene(1, 5, 9)
    fuse = None
    if case in ("unet", "unet_c3"):
        m, x = seg_models.UNet(13, 8, compress_level=3 if case == "unet_c3" else 0), x[:2]
    elif case == "v2v":
        m, fuse = seg_models.V2VNet(13, 8, num_agent=5), (trans, nat, 1, 5, False)
    elif case == "when2com":
        m = seg_models.When2Com_UNet(default_det_config(), n_classes=8, in_channels=13, warp_flag=1, num_agent=5)
        fuse = (trans, nat, 1, 5, False, "when2com", 1)
    else:
        m = seg_models.MeanFusion(13, 8, num_agent=5) if case == "mean" else seg_models.CatFusion(13, 8, 5, 0, False)
        fuse = (trans, nat, 1, 5, False, case)
    m.train()
    params = [p.detach().requires_grad_(True) for p in m.parameters()]
    out = SegTrainStep.apply(m, fuse, x, *params)
    assert tuple(out.shape) == (x.shape[0], 8, 256, 256)
    out.backward(torch.full_like(out, 1e-3))
    names = [k for k, _ in m.named_parameters()]
    got = {k for k, p in zip(names, params) if p.grad is not None}
    assert "inc.double_conv.0.weight" in got and "outc.conv.bias" in got
    assert ("com_compresser.weight" in got) == (case == "unet_c3")


# ---- unit-sharded plans (one rank of a gloo group; the collectives run for real on the CPU tensors) -----------------
@pytest.fixture
def gloo_single(tmp_path):
    import torch.distributed as dist
    if dist.is_initialized():
        pytest.skip("a process group is already initialised in this process")
    dist.init_process_group("gloo", rank=0, world_size=1, store=dist.FileStore(str(tmp_path / "store"), 1))
    yield dist
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["allgather", "neighbours"])
def test_v2vnet_sharded_plan_wiring(fake, gloo_single, exchange):
    from v2x_b200 import nets
    plan = nets.V2VNetDetShardedPlan(synth.v2vnet_det_state(1), 2, 5, 0, 1, planes="mixed", device="cpu", exchange=exchange)
    bevs, trans, nat = synth.make_scene(2, 5, 1)
    out = plan.forward(bevs, trans, nat)
    assert tuple(out["cls"].shape) == (10, 256 * 256 * 6, 2) and plan.exchange_planes == 1
    assert len(plan._segments()) == 3 and sum(len(s) for s in plan._segments()) == len(plan.launches)


@pytest.mark.parametrize("warp_flag", [0, 1])
def test_when2com_sharded_plan_wiring(fake, gloo_single, warp_flag):
    from v2x_b200 import nets
    plan = nets.When2comDetShardedPlan(synth.when2com_det_state(3), 1, 5, 0, 1, planes="mixed", device="cpu", warp_flag=warp_flag)
    bevs, trans, nat = synth.make_scene(1, 5, 3)
    out = plan.forward(bevs, trans, nat)
    assert tuple(out["loc"].shape) == (5, 256, 256, 6, 1, 6)
    assert hasattr(plan, "x3_all") == (not warp_flag)     # feature maps cross the wire only without the warp


def test_plan_classes_only_read_attributes_they_define():
    """Static check (AST) over the product package: every ``self.x`` a class reads is assigned somewhere in the class, one
    of its in-package bases or subclasses, or comes from a base outside the package.  (A ``capture()`` reading a
    ``self.peer`` its class never set survived a round because no CPU test could execute it.)"""
    import ast
    import os
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    classes = {}
    for fn in sorted(os.listdir(root)):
        if not fn.endswith(".py"):
            continue
        for node in ast.walk(ast.parse(open(os.path.join(root, fn)).read())):
            if not isinstance(node, ast.ClassDef):
                continue
            assigned, read = set(), {}
            for b in node.body:
                if isinstance(b, (ast.FunctionDef, ast.ClassDef)):
                    assigned.add(b.name)
                elif isinstance(b, ast.Assign):
                    assigned.update(t.id for t in b.targets if isinstance(t, ast.Name))
                elif isinstance(b, ast.AnnAssign) and isinstance(b.target, ast.Name):
                    assigned.add(b.target.id)
            for n in ast.walk(node):
                if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name) and n.value.id == "self":
                    if isinstance(n.ctx, ast.Store):
                        assigned.add(n.attr)
                    else:
                        read.setdefault(n.attr, n.lineno)
            bases = [b.id if isinstance(b, ast.Name) else getattr(b, "attr", "?") for b in node.bases]
            classes[node.name] = (bases, assigned, read, fn)

    def lineage(name, seen=None):     # the class, its in-package ancestors; None if a base lives outside the package
        seen = seen or set()
        if name in seen:
            return set()
        seen.add(name)
        out = {name}
        for b in classes[name][0]:
            if b not in classes:
                return None
            up = lineage(b, seen)
            if up is None:
                return None
            out |= up
        return out

    problems = []
    for name, (bases, assigned, read, fn) in classes.items():
        line = lineage(name)
        if line is None:
            continue
        family = set(line)
        family |= {n for n in classes if (lineage(n) or set()) & {name}}       # subclasses may set what a base reads
        have = set().union(*(classes[n][1] for n in family))
        problems += ["%s:%d %s.%s" % (fn, ln, name, a) for a, ln in read.items() if a not in have]
    assert not problems, problems
