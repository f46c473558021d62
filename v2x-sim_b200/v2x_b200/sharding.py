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
    planar = local.dim() == 5 and local.dtype == torch.bfloat16  # act layout [P, N, H, W, C]
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
