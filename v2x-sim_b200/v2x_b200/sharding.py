"""Unit sharding of the collaborative forward across ranks.

A *unit* is one (scene, agent) map; units are numbered agent-major (``u = B * agent + scene``, the layout of the
reference's batches, train_codet.py:287-334).  Encoder, decoder and heads are independent per unit; the fuse step of a
unit needs the layer-3 maps of the other agents of its scene -- and, because V2VNet always warps the ORIGINAL encoder
maps (V2VNet.py:85-94), exactly one exchange per forward: an all-gather of x_3.

Every rank owns a contiguous slice of ``units / world`` units.  ``all_gather_units`` is the only data-path collective
(NCCL on GPUs, gloo in the CPU tests); it is issued asynchronously so the x_4 branch of the encoder overlaps it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def unit_range(units: int, rank: int, world: int):
    """(offset, count) of the contiguous unit slice of ``rank``; units must divide evenly."""
    if units % world != 0:
        raise ValueError("agents * batch (%d units) must be a multiple of the world size %d "
                         "(shard a batch of B = 0 mod world scenes, SURVEY 8(e))" % (units, world))
    n = units // world
    return rank * n, n


def all_gather_units(local: torch.Tensor, out: torch.Tensor = None, group=None, async_op=False):
    """All-gather per-unit tensors ``local`` [n_loc, ...] (or act planes [P, n_loc, ...]) into [n_loc * world, ...]
    (resp. [P, n_loc * world, ...]) in rank order == global unit order.  Returns (out, work-or-None list)."""
    world = dist.get_world_size(group)
    planar = local.dim() == 5 and local.dtype in (torch.bfloat16, torch.float16)  # act layout [P, N, H, W, C]
    if planar:
        p, n = local.shape[0], local.shape[1]
        if out is None:
            out = local.new_empty((p, n * world) + tuple(local.shape[2:]))
        works = [dist.all_gather_into_tensor(out[i], local[i].contiguous(), group=group, async_op=async_op)
                 for i in range(p)]
    else:
        if out is None:
            out = local.new_empty((local.shape[0] * world,) + tuple(local.shape[1:]))
        works = [dist.all_gather_into_tensor(out, local.contiguous(), group=group, async_op=async_op)]
    return out, (works if async_op else None)


# ---------------------------------------------------------------------------------------------------------------------
# Targeted exchange: a unit's fuse step only reads the maps of the OTHER AGENTS OF ITS SCENE (DetModelBase.py:171-209), so
# a rank needs (agents - 1) maps per local unit -- 4/5 of what it owns, whatever the world size -- while an all-gather
# delivers (world - 1) x what it owns.  At 8 ranks that is 80 MB instead of 147 MB received per rank and step.
# ---------------------------------------------------------------------------------------------------------------------
def _needed_units(batch_total: int, agents: int, rank: int, world: int):
    n = batch_total * agents // world
    local = range(rank * n, rank * n + n)
    scenes = {u % batch_total for u in local}
    mine = set(local)
    return sorted(u for u in (batch_total * j + b for b in scenes for j in range(agents)) if u not in mine)


def _runs(units, n):
    """Sorted unit ids -> maximal runs (owner, start, count) of consecutive ids with the same owner rank (u // n)."""
    out = []
    for u in units:
        if out and out[-1][1] + out[-1][2] == u and out[-1][0] == u // n:
            out[-1][2] += 1
        else:
            out.append([u // n, u, 1])
    return [tuple(r) for r in out]


def neighbour_exchange_plan(batch_total: int, agents: int, rank: int, world: int):
    """(sends, recvs) of ``rank``: sends = [(peer, local_start, count)], recvs = [(peer, global_start, count)], both in
    ascending unit order per peer, so the k-th send of s to r pairs with the k-th receive of r from s."""
    units = batch_total * agents
    if units % world != 0:
        raise ValueError("agents * batch (%d units) must be a multiple of the world size %d" % (units, world))
    n = units // world
    recvs = _runs(_needed_units(batch_total, agents, rank, world), n)
    sends = []
    for q in range(world):
        if q == rank:
            continue
        for owner, start, count in _runs(_needed_units(batch_total, agents, q, world), n):
            if owner == rank:
                sends.append((q, start - rank * n, count))
    return sends, recvs


def exchange_neighbour_units(local: torch.Tensor, out: torch.Tensor, batch_total: int, agents: int, group=None):
    """Fill ``out`` (the global-unit-indexed buffer of act planes [P, units, ...]) with this rank's own maps and the maps
    of the other agents of its scenes, by point-to-point sends / receives of contiguous unit runs (NCCL send/recv in one
    group on GPUs, gloo in the CPU tests).  Entries no local unit reads are left untouched.  Returns the work handles."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    assert local.dim() == 5 and out.dim() == 5 and out.shape[1] == local.shape[1] * world
    n = local.shape[1]
    sends, recvs = neighbour_exchange_plan(batch_total, agents, rank, world)
    out[:, rank * n:rank * n + n].copy_(local, non_blocking=True)
    ops = []
    for p in range(local.shape[0]):
        for peer, ls, c in sends:
            ops.append(dist.P2POp(dist.isend, local[p, ls:ls + c], peer, group))
        for peer, gs, c in recvs:
            ops.append(dist.P2POp(dist.irecv, out[p, gs:gs + c], peer, group))
    return dist.batch_isend_irecv(ops) if ops else []
