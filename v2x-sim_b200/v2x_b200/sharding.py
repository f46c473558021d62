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


# ---------------------------------------------------------------------------------------------------------------------
# Device-side exchange over NVLink peer memory (csrc/peer_kernels.cu): every rank owns a peer-visible region
# [flag block | payload]; ranks push their maps straight into each other's payload and synchronise through the flag
# blocks, so the forward needs no host-issued collective and is ONE CUDA graph.  torch.distributed is only used once, at
# set-up, to hand the 64-byte CUDA IPC handles around.
# ---------------------------------------------------------------------------------------------------------------------
PEER_HANDLE_BYTES = 64
PEER_FLAG_BYTES = 4096
PEER_TIMEOUT_MS = 20000


class _RawCudaBuffer:
    """Exposes a raw device allocation through __cuda_array_interface__ so torch can view it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerRegion:
    """This rank's peer-visible region plus the mapped regions of every other rank of the (single-node) group.

    ``payload(shape, dtype)`` views the local payload as a tensor; ``payload_ptr(r)`` is the device address of rank r's
    payload as seen from THIS process; ``begin/push/wait/done`` enqueue the four kernels of one exchange step on the
    current stream (capturable into a CUDA graph); ``check()`` raises if a wait ever timed out."""

    def __init__(self, payload_bytes: int, rank: int, world: int, group=None, device="cuda"):
        import ctypes as C
        from . import ops
        from ._lib import check
        if world > 8:
            raise ops.V2XError("the peer-memory exchange supports up to 8 ranks (one NVSwitch domain)")
        self.lib = ops.require_gpu()
        self.rank, self.world, self.group, self.device = rank, world, group, torch.device(device)
        self.nbytes = PEER_FLAG_BYTES + ((int(payload_bytes) + 255) // 256) * 256
        ptr, handle = C.c_void_p(), (C.c_uint8 * PEER_HANDLE_BYTES)()
        with torch.cuda.device(self.device):
            check(self.lib.v2x_peer_alloc(self.nbytes, C.byref(ptr), handle), "v2x_peer_alloc")
            self.local_ptr = int(ptr.value)
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.device)
            if world > 1:
                gathered = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(gathered, mine, group=group)
            else:
                gathered = [mine]
            self.ptrs = []
            for r in range(world):
                if r == rank:
                    self.ptrs.append(self.local_ptr)
                    continue
                h = (C.c_uint8 * PEER_HANDLE_BYTES)(*gathered[r].cpu().tolist())
                p = C.c_void_p()
                check(self.lib.v2x_peer_open(h, C.byref(p)), "v2x_peer_open(rank %d)" % r)
                self.ptrs.append(int(p.value))
            self._raw = _RawCudaBuffer(self.local_ptr, self.nbytes)
            self._bytes = torch.as_tensor(self._raw, device=self.device)
            assert self._bytes.data_ptr() == self.local_ptr, "torch copied the peer region instead of viewing it"
            self.step = torch.zeros(1, dtype=torch.int64, device=self.device)
            self.counter = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
            torch.cuda.synchronize(self.device)
        if world > 1:
            dist.barrier(group=group)        # every region is zeroed and mapped before anyone pushes
        self._regions = (C.c_void_p * world)(*self.ptrs)

    def payload(self, shape, dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        nb = n * torch.empty((), dtype=dtype).element_size()
        assert PEER_FLAG_BYTES + nb <= self.nbytes
        return self._bytes[PEER_FLAG_BYTES:PEER_FLAG_BYTES + nb].view(dtype).view(*shape)

    def payload_ptr(self, r: int) -> int:
        return self.ptrs[r] + PEER_FLAG_BYTES

    def begin(self):
        from . import ops
        from ._lib import check
        check(self.lib.v2x_peer_begin(self.local_ptr, self.step.data_ptr(), self.rank, self.world, PEER_TIMEOUT_MS,
                                      self.err.data_ptr(), ops._stream()), "v2x_peer_begin")

    def push(self, src: torch.Tensor, dst_elem_offset: int, dst_plane_stride: int, dst_planes):
        """src: act planes [P, n, ...] (contiguous); lands at element ``dst_elem_offset`` of plane 0 of every rank's
        payload (16-bit elements), ``dst_planes[r]`` planes for rank r."""
        import ctypes as C
        from . import ops
        from ._lib import check
        assert src.is_contiguous() and src.element_size() == 2
        per_plane = src[0].numel()
        dst = (C.c_void_p * self.world)(*[self.payload_ptr(r) + 2 * dst_elem_offset for r in range(self.world)])
        planes = (C.c_int32 * self.world)(*[int(p) for p in dst_planes])
        check(self.lib.v2x_peer_push(src.data_ptr(), per_plane, per_plane, dst, planes, dst_plane_stride, self._regions,
                                     self.step.data_ptr(), self.counter.data_ptr(), self.rank, self.world, ops._stream()),
              "v2x_peer_push")

    def wait(self):
        from . import ops
        from ._lib import check
        check(self.lib.v2x_peer_wait(self.local_ptr, self.step.data_ptr(), self.rank, self.world, PEER_TIMEOUT_MS,
                                     self.err.data_ptr(), ops._stream()), "v2x_peer_wait")

    def done(self):
        from . import ops
        from ._lib import check
        check(self.lib.v2x_peer_done(self._regions, self.step.data_ptr(), self.rank, self.world, ops._stream()),
              "v2x_peer_done")

    def check(self):
        """Host-side check (synchronises): raises if any wait of this rank timed out since the last check."""
        from . import ops
        code = int(self.err.item())
        if code:
            self.err.zero_()
            raise ops.V2XError("peer exchange: rank %d timed out waiting for rank %d (%s flag) after %d ms" %
                               (self.rank, code % 100, "consumed" if code < 200 else "ready", PEER_TIMEOUT_MS))

    def close(self):
        """Unmap the other ranks' regions, then (after a barrier: nobody still maps it) free the local one.  Collective."""
        if getattr(self, "ptrs", None) is None:
            return
        from ._lib import check
        torch.cuda.synchronize(self.device)
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                check(self.lib.v2x_peer_close(p), "v2x_peer_close")
        if self.world > 1:
            dist.barrier(group=self.group)
        self._bytes = self._raw = None
        check(self.lib.v2x_peer_free(self.local_ptr), "v2x_peer_free")
        self.ptrs = None
