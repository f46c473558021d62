"""CPU, world_size = 2 over gloo: the unit-sharding host logic of the multi-GPU path (v2x_b200/sharding.py).

Each rank encodes only its slice of the agent-major units (oracle encoder as the compute stand-in), the ranks
all-gather x_3 through the same ``all_gather_units`` the GPU plan uses, fuse + decode their own units, and the
concatenation must equal the unsharded oracle forward: one exchange per forward is sufficient."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "v2x-sim_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import restate, synth
        from v2x_b200 import sharding
        torch.set_num_threads(2)
        B, A = 2, 2                      # 4 units, 2 per rank: rank 0 = agent 0 of both scenes, rank 1 = agent 1
        sd = synth.v2vnet_det_state(5)
        bevs, trans, nat = synth.make_scene(B, A, seed=5)
        off, n = sharding.unit_range(B * A, rank, world)
        with torch.no_grad():
            enc = restate.encode(bevs[off:off + n], sd, "u_encoder.")
            x3_all, _ = sharding.all_gather_units(enc[3])
            fused_all = restate.v2vnet_fuse(x3_all, trans, nat, sd, B, agent_num=A, gnn_iter=2)
            dec_in = list(enc)
            dec_in[3] = fused_all[off:off + n]
            out = restate.heads(restate.decode(*dec_in, sd, "decoder.")[0], sd)
            full = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=B, agent_num=A, gnn_iter=2)
        err = max((out[k] - full[k][off:off + n]).abs().max().item() for k in ("loc", "cls"))
        # act-plane tensors take the per-plane path of all_gather_units
        planes = torch.full((2, n, 2, 2, 8), float(rank), dtype=torch.bfloat16)
        g, _ = sharding.all_gather_units(planes)
        ok_planes = g.shape == (2, n * world, 2, 2, 8) and bool((g[:, :n] == 0).all()) and bool((g[:, n:] == 1).all())
        ret[rank] = (err, ok_planes)
    finally:
        dist.destroy_process_group()


def test_unit_range():
    from v2x_b200 import sharding
    assert sharding.unit_range(40, 3, 4) == (30, 10)
    with pytest.raises(ValueError):
        sharding.unit_range(5, 0, 4)     # one 5-agent scene cannot be split over 4 ranks


def test_two_rank_sharded_forward_matches_unsharded():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    for r in range(2):
        err, ok_planes = ret[r]
        assert err < 1e-4, (r, err)
        assert ok_planes


@pytest.mark.parametrize("world,batch_total,agents", [(2, 2, 2), (2, 4, 5), (4, 8, 5), (4, 32, 5), (8, 64, 5), (8, 8, 5),
                                                      (5, 3, 5), (8, 16, 6)])
def test_neighbour_exchange_plan_is_complete_and_consistent(world, batch_total, agents):
    """Simulate the targeted x_3 exchange on one process: after it, every rank holds its own maps and every map of the
    other agents of its scenes, each received exactly once, and the k-th send of s to r pairs with r's k-th receive from s."""
    import numpy as np
    from v2x_b200 import sharding
    units = batch_total * agents
    n = units // world
    truth = np.arange(units)
    plans = [sharding.neighbour_exchange_plan(batch_total, agents, r, world) for r in range(world)]
    for r in range(world):
        sends, recvs = plans[r]
        buf = np.full(units, -1)
        buf[r * n:(r + 1) * n] = truth[r * n:(r + 1) * n]
        for peer in range(world):
            from_peer = [(gs, c) for (p, gs, c) in recvs if p == peer]
            to_me = [(ls, c) for (q, ls, c) in plans[peer][0] if q == r] if peer != r else []
            assert len(from_peer) == len(to_me)
            for (gs, c), (ls, c2) in zip(from_peer, to_me):
                assert c == c2 and gs == peer * n + ls       # pairs up in issue order
                assert (buf[gs:gs + c] == -1).all()          # nothing is received twice
                buf[gs:gs + c] = truth[peer * n + ls:peer * n + ls + c]
        for u in range(r * n, (r + 1) * n):
            b = u % batch_total
            for j in range(agents):
                assert buf[batch_total * j + b] == batch_total * j + b
        received = int((buf >= 0).sum()) - n
        assert received <= n * (agents - 1)


def _worker_exchange(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "v2x-sim_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from v2x_b200 import sharding
        B, A = 4, 5
        n = B * A // world
        full = torch.arange(2 * B * A * 3, dtype=torch.float32).view(2, B * A, 1, 1, 3).to(torch.bfloat16)
        local = full[:, rank * n:(rank + 1) * n].contiguous()
        out = torch.full_like(full, -1.0)
        for w in sharding.exchange_neighbour_units(local, out, B, A):
            w.wait()
        ok = True
        for u in range(rank * n, (rank + 1) * n):
            for j in range(A):
                g = B * j + u % B
                ok = ok and bool(torch.equal(out[:, g], full[:, g]))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_neighbour_exchange_over_gloo_world4():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_worker_exchange, args=(r, 4, port, ret)) for r in range(4)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(4))


def _worker_plans(rank, world, port, ret):
    """Both unit-sharded PLAN classes on two ranks: kernels replaced by a fake library that validates every ctypes call
    (tests/test_plan_dryrun.py), collectives real (gloo).  What crosses the wire is marked by rank so the exchange wiring
    -- which buffer, which unit rows, which planes -- is checked exactly."""
    import ctypes as C
    for p in (ROOT, os.path.join(ROOT, "v2x-sim_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import synth
        from test_plan_dryrun import FakeLib
        from v2x_b200 import nets, ops
        torch.set_num_threads(2)
        fake = FakeLib()
        ops.require_gpu = lambda: fake
        ops._stream = lambda: C.c_void_p(0)
        torch.Tensor.is_cuda = property(lambda self: True)
        B, A = 2, 5                       # 10 units, 5 per rank
        res = {}
        for exchange in ("allgather", "neighbours"):
            plan = nets.V2VNetDetShardedPlan(synth.v2vnet_det_state(1), B, A, rank, world, planes="mixed", device="cpu",
                                             exchange=exchange)
            bevs, trans, nat = synth.make_scene(B, A, 1)
            plan.x3_local.fill_(float(rank + 1))          # stands in for this rank's encoder output
            plan.x3_all.zero_()
            out = plan.forward(bevs[plan.offset:plan.offset + plan.n], trans, nat)
            n = plan.n
            hi, lo = plan.x3_all[0].float(), plan.x3_all[1].float()
            own = slice(rank * n, rank * n + n)
            other = slice((1 - rank) * n, (1 - rank) * n + n)
            res[exchange] = (tuple(out["cls"].shape) == (n, 256 * 256 * 6, 2)
                             and bool((hi[own] == rank + 1).all()) and bool((lo[own] == rank + 1).all())      # own units: both planes
                             and bool((hi[other] == 2 - rank).all()) and bool((lo[other] == 0).all()))       # remote: hi plane only
        for warp_flag in (1, 0):
            plan = nets.When2comDetShardedPlan(synth.when2com_det_state(3), B, A, rank, world, planes="mixed", device="cpu",
                                               warp_flag=warp_flag)
            bevs, trans, nat = synth.make_scene(B, A, 3)
            plan.keys_local.fill_(float(rank + 1))
            plan.querys_local.fill_(float(10 * (rank + 1)))
            plan.x3_local.fill_(float(rank + 1))
            out = plan.forward(bevs[plan.offset:plan.offset + plan.n], trans, nat)
            n = plan.n
            ok = tuple(plan.keys.shape) == (B * A, 1024) and tuple(plan.querys.shape) == (B * A, 32)
            for r in range(world):
                ok = ok and bool((plan.keys[r * n:(r + 1) * n] == r + 1).all()) and bool((plan.querys[r * n:(r + 1) * n] == 10 * (r + 1)).all())
            if not warp_flag:             # without the warp every agent's map is read: x_3 crosses the wire too
                ok = ok and all(bool((plan.x3_all[:, r * n:(r + 1) * n].float() == r + 1).all()) for r in range(world))
            else:
                ok = ok and not hasattr(plan, "x3_all")
            res["w2c%d" % warp_flag] = ok and tuple(out["loc"].shape) == (n, 256, 256, 6, 1, 6)
        ret[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_plans_exchange_the_right_rows():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 30100 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker_plans, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    for r in range(2):
        assert all(ret[r].values()), (r, dict(ret[r]))
