"""CPU tests of the host-side logic added in round 2: precision modes / per-layer pass policy, the plan-cache fingerprint
and output-aliasing rules of the drop-in modules, the staged-reference recipe, bench.py's config table."""
import importlib.util
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_precision_modes_and_policy(monkeypatch):
    from v2x_b200 import precision
    assert precision.resolve(1).name == "bf16" and precision.resolve(2).name == "fp16x3"
    assert precision.resolve("bf16x3").name == "fp16x3"          # round-1 name kept as an alias
    assert precision.resolve(None).name == precision.DEFAULT == os.environ.get("V2X_PRECISION", "mixed")
    with pytest.raises(ValueError):
        precision.resolve("fp8")
    bf, x3, mx = precision.resolve("bf16"), precision.resolve("fp16x3"), precision.resolve("mixed")
    assert (bf.planes, x3.planes, mx.planes) == (1, 2, 2)
    for layer in ("conv_pre_1", "conv5_1", "gru", "heads"):
        assert bf.mmas(layer) == 1 and x3.mmas(layer) == 3
    # the measured policy: GRU 1 pass, conv5_1 2 passes, everything else 3
    assert mx.mmas("gru") == 1 and mx.mmas("conv5_1") == 2
    assert all(mx.mmas(l) == 3 for l in ("conv_pre_1", "conv4_2", "conv6_1", "conv7_1", "conv8_1", "conv8_2", "heads"))
    monkeypatch.setenv("V2X_MIXED_POLICY", "gru=2,conv8_1=2")
    assert mx.mmas("gru") == 2 and mx.mmas("conv8_1") == 2 and mx.mmas("conv5_1") == 3
    assert precision.resolve(mx) is mx and mx == precision.Precision("mixed") and hash(mx) == hash(precision.Precision("mixed"))


def test_weight_format_of_each_pass_count():
    from v2x_b200 import ops
    assert ops.weight_fmt(1, 1) == (ops.FMT_BF16, 1)
    assert ops.weight_fmt(2, 3) == (ops.FMT_F16X2, 2)
    assert ops.weight_fmt(2, 2) == (ops.FMT_F16, 1) and ops.weight_fmt(2, 1) == (ops.FMT_F16, 1)
    assert ops.act_dtype(1) == torch.bfloat16 and ops.act_dtype(2) == torch.float16


def _v2v():
    from coperception.models.det import V2VNet
    from v2x_b200 import default_det_config
    return V2VNet(default_det_config(), 3, 3, 256, num_agent=5)


def test_default_precision_and_fingerprint_on_cpu():
    """The plan cache is keyed by a fingerprint of every parameter's (pointer, version): tracked in-place edits change it,
    reads do not; `verify_weights` adds a value checksum that also sees edits through .data."""
    m = _v2v()
    assert m.precision == os.environ.get("V2X_PRECISION", "mixed") and m.alias_outputs is False
    f0 = m._fingerprint()
    assert m._fingerprint() == f0
    with torch.no_grad():
        m.classification.conv2.bias.add_(1.0)
    f1 = m._fingerprint()
    assert f1 != f0
    m.classification.conv2.bias.data.add_(1.0)       # bypasses the version counter
    assert m._fingerprint() == f1
    m.verify_weights = True
    f2 = m._fingerprint()
    m.classification.conv2.bias.data.add_(1.0)
    assert m._fingerprint() != f2
    sd = m.state_dict()
    m._plans["x"] = (0, object())
    m.load_state_dict(sd)
    assert m._plans == {}                             # load_state_dict invalidates


def test_outputs_are_cloned_unless_aliased():
    m = _v2v()
    static = torch.zeros(4)
    m._static_ptrs.add(static.data_ptr())
    other = torch.ones(3)
    out = m._out({"cls": static.view(2, 2), "loc": other})
    assert out["cls"].data_ptr() != static.data_ptr() and out["loc"] is other
    tup = m._out((static, (other, static)))
    assert tup[0].data_ptr() != static.data_ptr() and tup[1][0] is other and tup[1][1].data_ptr() != static.data_ptr()
    m.alias_outputs = True
    assert m._out({"cls": static})["cls"] is static


def test_forward_refuses_cpu_tensors_and_eval_warning_is_once():
    m = _v2v().eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros((5, 1, 256, 256, 13)), torch.zeros((1, 5, 5, 4, 4)), torch.full((1, 5), 5), batch_size=1)


def test_training_is_offered_where_built_and_refused_elsewhere():
    """.train() forwards are routed to the training tape (which needs CUDA tensors: no CPU fallback) for every model that
    trains on the path; what is not built -- FaFNet's kd tuple, fewer than 32 compressed channels, seg DiscoNet, a When2com
    without its query net or asked for the gated inference pass in .train() -- raises NotImplementedError before touching
    the device."""
    from coperception.models.det import DiscoNet, FaFNet, MeanFusion, When2com
    from coperception.models.seg import DiscoNet as SegDiscoNet, MeanFusion as SegMeanFusion
    from v2x_b200 import default_det_config
    cfg = default_det_config()
    det_args = (torch.zeros((5, 1, 256, 256, 13)), torch.zeros((1, 5, 5, 4, 4)), torch.full((1, 5), 5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FaFNet(cfg, kd_flag=0).train()(torch.zeros((1, 1, 256, 256, 13)))
    for model in (When2com(cfg, layer=3), MeanFusion(cfg, layer=3, kd_flag=0), MeanFusion(cfg, layer=3, kd_flag=1),
                  DiscoNet(cfg, layer=3, kd_flag=1)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model.train()(*det_args, batch_size=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SegMeanFusion(13, 8, num_agent=5).train()(torch.zeros((5, 13, 256, 256)), det_args[1], det_args[2])
    with pytest.raises(NotImplementedError):
        FaFNet(cfg, kd_flag=1).train()(torch.zeros((1, 1, 256, 256, 13)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):       # compressed training: >= 32 compressed channels
        MeanFusion(cfg, layer=3, kd_flag=0, compress_level=2).train()(*det_args, batch_size=1)
    with pytest.raises(NotImplementedError):
        MeanFusion(cfg, layer=3, kd_flag=0, compress_level=4).train()(*det_args, batch_size=1)
    with pytest.raises(NotImplementedError):
        When2com(cfg, layer=3, has_query=False).train()(*det_args, batch_size=1)
    with pytest.raises(NotImplementedError):
        SegDiscoNet(13, 8, 5, kd_flag=False).train()(torch.zeros((5, 13, 256, 256)), det_args[1], det_args[2])


def test_make_ref_stages_the_reference_and_ref_loader_finds_it(tmp_path, monkeypatch):
    from oracle import make_ref
    if not os.path.isdir(os.path.join(make_ref.SRC, "models", "det")):
        pytest.skip("reference tree only exists in the build container")
    monkeypatch.setattr(make_ref, "DST", str(tmp_path / "_ref" / "coperception"))
    make_ref.main(quiet=True)
    for rel in ("models/det/V2VNet.py", "models/seg/When2Com_UNet.py", "configs/Config.py", "utils/convolutional_rnn/module.py"):
        assert os.path.exists(os.path.join(make_ref.DST, rel)), rel
    assert not any(f.endswith(".so") for _, _, fs in os.walk(make_ref.DST) for f in fs)


def test_bench_config_table():
    spec = importlib.util.spec_from_file_location("bench_cfg", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert set(bench.CONFIGS) >= {"v2v_det", "faf_lower", "w2c_seg", "faf_upper_dp", "w2c_det"}
    # algorithmic GFLOP per unit, SURVEY.md section 8(d)
    assert bench.CONFIGS["v2v_det"]["gflop"] == 264.51 and bench.CONFIGS["faf_lower"]["gflop"] == 31.16
    assert bench.CONFIGS["w2c_seg"]["gflop"] == 581.38 and abs(bench.CONFIGS["faf_upper_dp"]["gflop"] - 6 * 31.16) < 1e-9
    assert set(bench.DTYPE_OF) == {"mixed", "fp16x3", "bf16"}


def test_warp_staged_box_rule_covers_every_tap():
    """csrc/warp_staged.cuh stages, per (8x8 output tile, member term), the bounding box of the tile's footprint in the
    source map: the box of the four CORNER samples widened by 0.01 px, +1 for the bilinear taps, clipped to the map.  This
    emulates that rule in float32 exactly as the kernel evaluates it (affine_grid + grid_sample, align_corners=False, theta'
    of the un-flipped domain) and checks, over random rigid poses and map sizes incl. partial tiles, that (a) every in-map tap
    of every output pixel lies inside the box -- so the kernel's direct-gather fallback for stray taps never triggers -- and
    (b) the box never exceeds the 192 staged pixels (kWsCap), so rigid poses always take the staged path."""
    import numpy as np
    f32 = np.float32

    def sample(t, ow, oh, W, H):
        gx = (f32(2) * ow.astype(f32) + f32(1)) / f32(W) - f32(1)
        gy = (f32(2) * oh.astype(f32) + f32(1)) / f32(H) - f32(1)
        sx = t[0] * gx + t[1] * gy + t[2]
        sy = t[3] * gx + t[4] * gy + t[5]
        return ((sx + f32(1)) * f32(W) - f32(1)) * f32(0.5), ((sy + f32(1)) * f32(H) - f32(1)) * f32(0.5)

    rng = np.random.default_rng(0)
    stray, worst = 0, 0
    for _ in range(400):
        W = H = int(rng.choice([32, 64, 20, 28]))
        yaw = rng.uniform(-np.pi, np.pi)
        tx, ty = rng.uniform(-60, 60, 2)
        c, s = np.cos(yaw), np.sin(yaw)
        t = np.array([c, s, -tx / 32, -s, c, ty / 32], dtype=f32)     # [[T00, -T01, -T03/32], [-T10, T11, +T13/32]], T01 = -s
        for th in range((H + 7) // 8):
            for tw in range((W + 7) // 8):
                oh0, ow0 = th * 8, tw * 8
                cx, cy = sample(t, np.array([ow0, ow0 + 7, ow0, ow0 + 7]), np.array([oh0, oh0, oh0 + 7, oh0 + 7]), W, H)
                bx0 = max(0, int(np.floor(cx.min() - f32(0.01))))
                bx1 = min(W - 1, int(np.floor(cx.max() + f32(0.01))) + 1)
                by0 = max(0, int(np.floor(cy.min() - f32(0.01))))
                by1 = min(H - 1, int(np.floor(cy.max() + f32(0.01))) + 1)
                pw, ph = np.meshgrid(np.arange(ow0, min(ow0 + 8, W)), np.arange(oh0, min(oh0 + 8, H)))
                ix, iy = sample(t, pw.ravel(), ph.ravel(), W, H)
                x0, y0 = np.floor(ix).astype(int), np.floor(iy).astype(int)
                empty = bx0 > bx1 or by0 > by1
                for dx in (0, 1):
                    for dy in (0, 1):
                        xx, yy = x0 + dx, y0 + dy
                        in_map = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
                        in_box = (xx >= bx0) & (xx <= bx1) & (yy >= by0) & (yy <= by1) if not empty else np.zeros_like(in_map)
                        stray += int((in_map & ~in_box).sum())
                if not empty:
                    worst = max(worst, (bx1 - bx0 + 1) * (by1 - by0 + 1))
    assert stray == 0
    assert worst <= 192, worst


def test_peer_exchange_protocol_is_deadlock_free_and_never_overwrites_unread_maps():
    """Discrete-event model of the device-side x_3 exchange (csrc/peer_kernels.cu): every rank runs, in its own stream
    order, begin (step += 1; wait consumed[r] >= step - 1 for all r) -> push (write its slot of EVERY region, then publish
    ready[rank] = step everywhere) -> wait (ready[r] >= step for all r) -> read (the fuse kernel reads its own region) ->
    done (publish consumed[rank] = step everywhere).  Ranks are interleaved by a random scheduler (any rank whose next
    kernel is not blocked may run: that is all stream order guarantees).  Checked over many schedules and world sizes:
    no schedule deadlocks, every read sees exactly the maps of its own step from every rank, and no slot is overwritten
    between the time it was published and the time its owner region's rank has read it (the region is single-buffered)."""
    import random
    for world in (1, 2, 3, 4, 8):
        for trial in range(60):
            rnd = random.Random(1000 * world + trial)
            steps = 6
            ready = [[0] * world for _ in range(world)]       # ready[region][writer]
            consumed = [[0] * world for _ in range(world)]    # consumed[region][reader]
            slot = [[0] * world for _ in range(world)]        # slot[region][writer] = step whose maps are stored there
            pending_read = [[False] * world for _ in range(world)]   # published but not yet read by the region's rank
            step = [0] * world
            pc = [0] * world                                   # 0 begin, 1 push, 2 wait, 3 read, 4 done
            done_steps = [0] * world
            guard = 0
            while min(done_steps) < steps:
                guard += 1
                assert guard < 100000
                runnable = []
                for r in range(world):
                    if done_steps[r] >= steps:
                        continue
                    if pc[r] == 0 and not all(consumed[r][q] >= step[r] for q in range(world) if q != r):
                        continue                                # begin: peers have not consumed my previous maps yet
                    if pc[r] == 2 and not all(ready[r][q] >= step[r] for q in range(world) if q != r):
                        continue                                # wait: not every rank's maps of this step have landed
                    runnable.append(r)
                assert runnable, "deadlock: world %d trial %d state %s %s" % (world, trial, pc, step)
                r = rnd.choice(runnable)
                if pc[r] == 0:
                    step[r] += 1
                elif pc[r] == 1:
                    for region in range(world):
                        assert not pending_read[region][r], "rank %d overwrote maps rank %d has not read" % (r, region)
                        slot[region][r] = step[r]
                        pending_read[region][r] = True
                        ready[region][r] = step[r]
                elif pc[r] == 3:
                    assert all(slot[r][q] == step[r] for q in range(world)), (world, trial, r, slot[r], step[r])
                    for q in range(world):
                        pending_read[r][q] = False
                elif pc[r] == 4:
                    for region in range(world):
                        consumed[region][r] = step[r]
                    done_steps[r] += 1
                pc[r] = (pc[r] + 1) % 5
            assert step == [steps] * world


def test_child_runner_isolates_pass_fail_and_hang(tmp_path):
    """tests/child_run.py (used by the never-run-on-hardware GPU cases): a passing selection passes, a failing one
    reports its exit status, one that only skips is not a pass, and one that never returns is killed at the time limit."""
    import time
    sys_path = os.path.join(ROOT, "tests")
    import sys
    if sys_path not in sys.path:
        sys.path.insert(0, sys_path)
    from child_run import ran_and_passed, run_in_child
    f = tmp_path / "test_tmp_cases.py"
    f.write_text("import time, pytest\n"
                 "def test_ok():\n    assert True\n"
                 "def test_bad():\n    assert False\n"
                 "def test_skip():\n    pytest.skip('no')\n"
                 "def test_hang():\n    time.sleep(600)\n")
    rc, tail = run_in_child(str(f), "test_ok", str(tmp_path / "ok.log"), 120)
    assert ran_and_passed(rc, tail), tail
    rc, tail = run_in_child(str(f), "test_bad", str(tmp_path / "bad.log"), 120)
    assert rc == 1 and not ran_and_passed(rc, tail)
    rc, tail = run_in_child(str(f), "test_skip", str(tmp_path / "skip.log"), 120)
    assert rc == 0 and not ran_and_passed(rc, tail)
    t0 = time.time()
    rc, tail = run_in_child(str(f), "test_hang", str(tmp_path / "hang.log"), 8)
    assert rc == "time limit of 8 s" and time.time() - t0 < 60 and not ran_and_passed(rc, tail)


def test_plan_cache_is_bounded_and_least_recently_used():
    """The module keeps at most ``max_plans`` plans (each owns GBs of static workspace): a hit refreshes its key, the
    least recently used one is dropped, the set of static output pointers follows the live plans."""
    m = _v2v()
    m.use_cuda_graph = False
    m.max_plans = 2

    class P:
        def __init__(self, tag):
            self.tag, self.loc = tag, torch.zeros(4)

    built = []

    def factory(tag):
        def f():
            built.append(tag)
            return P(tag)
        return f
    a = m._get_plan("a", factory("a"))
    b = m._get_plan("b", factory("b"))
    assert m._get_plan("a", factory("a2")) is a and built == ["a", "b"]          # hit: nothing rebuilt, "a" is now freshest
    c = m._get_plan("c", factory("c"))
    assert list(m._plans) == ["a", "c"]                                          # "b" was the least recently used
    assert m._static_ptrs == {a.loc.data_ptr(), c.loc.data_ptr()} and b.loc.data_ptr() not in m._static_ptrs
    with torch.no_grad():
        m.classification.conv2.bias.add_(1.0)                                    # weights changed: the key is rebuilt
    a2 = m._get_plan("a", factory("a3"))
    assert a2 is not a and built[-1] == "a3" and list(m._plans) == ["c", "a"]
