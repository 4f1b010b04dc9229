"""Multi-GPU Bloom-filter radix join, one process per GPU.

Two implementations of the same sharding live here:
  * DistGroup / DistJoinGraph (bottom of the file, the product path): the collective join runs inside the library
    (hwbrj_dist_join): fused partition + all-to-all and fused filter-slice build + all-gather over NVLink peer memory,
    device-side barriers; torch.distributed only carries the handles at start-up.
  * dist_join (NCCL reference path): the same steps spelled out with torch.distributed all-to-all / all-gather /
    all-reduce between the library's building-block kernels. It is what the peer-memory join is validated against, and
    its orchestration is exercised on CPU with the gloo backend (tests/test_dist_gloo.py).

Sharding. Rank g holds a contiguous chunk of R and of S (the GPU analogue of the reference's per-thread chunks,
parallel_radix_join_bloom.c:1646-1672). Every key has one OWNER rank, a pure function of the key, so owners join
independently:
  * sliceable filter (BASIC with k <= 1, or BLOCKED): owner = the rank holding the 1/G slice of the filter that
    contains the key's bits -- each GPU builds exactly its slice from the R tuples routed to it and the slices are
    all-gathered into a replicated filter;
  * otherwise (BASIC k > 1, or no filter): owner = top bits of crapwow(42,key); each GPU builds a full-size
    partial filter from its routed R tuples, partials are all-gathered and OR-ed (NCCL has no bitwise-OR op).
Steps: (1) partition local R by owner, all-to-all; (2) build filter slice / partial, all-gather (+OR);
(3) pre-filter the LOCAL S chunk with the replicated filter, so only survivors cross NVLink; (4) partition
survivors by owner, all-to-all; (5) local radix join of owned R with owned survivors; (6) all-reduce of
{matches, filtered, checksums}. The three scalars are identical for every G and equal to the single-GPU / CPU
oracle values.

`ops` abstracts the local compute so that the sharding and exchange logic can be exercised on CPU with the gloo
backend (tests provide an oracle-backed ops object); the only production implementation is CudaOps below.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _native as N
from .api import BLOCKED, BloomFilterArgs

MASK64 = (1 << 64) - 1


def sliceable(bloom: Optional[BloomFilterArgs], world: int) -> bool:
    if bloom is None:
        return False
    if bloom.variant == BLOCKED:
        return bloom.m // bloom.B >= world
    return bloom.k <= 1 and bloom.m >= world * 8


class CudaOps:
    """Local compute on this rank's GPU. Tuples travel as torch.int64 tensors (8 bytes per tuple): torch owns the
    memory and the collectives, the kernels get raw device pointers on torch's current stream."""

    def __init__(self, device: torch.device):
        self.device = device
        self.L = N.load()
        if self.L.hwbrj_set_device(device.index) != 0:
            raise RuntimeError("hwbrj_set_device failed")
        self.L.hwbrj_set_quiet(1)
        self.L.hwbrj_set_stream(torch.cuda.current_stream(device).cuda_stream)

    # -- buffers
    def empty_tuples(self, n: int) -> torch.Tensor:
        return torch.empty(max(n, 1) + 8, dtype=torch.int64, device=self.device)[:n]  # +64 B slack for 16-byte bulk loads

    def _wrap(self, t: torch.Tensor) -> int:
        return self.L.hwbrj_rel_wrap(t.data_ptr(), t.numel())

    def view_int64(self, ptr: int, n: int) -> torch.Tensor:
        """torch view of library-owned device memory (plumbing for zeroing / reading small control words)"""
        class _View:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}
        return torch.as_tensor(_View(), device=self.device)

    def generate_shard(self, kind: int, n: int, r: int, q: float, seed: int, begin: int, count: int) -> torch.Tensor:
        h = self.L.hwbrj_rel_generate_shard(kind, n, r, q, seed, begin, count)
        out = self.empty_tuples(count)
        if count:
            src = self.L.hwbrj_rel_ptr(h)

            class _View:  # the library buffer seen by torch (plumbing): a device-to-device copy into torch memory
                __cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (src, False), "version": 3}
            out.copy_(torch.as_tensor(_View(), device=self.device))
            torch.cuda.current_stream(self.device).synchronize()
        self.L.hwbrj_rel_free(h)
        return out

    # -- kernels
    def owner_partition(self, rel: torch.Tensor, world: int, slice_args: Optional[BloomFilterArgs]):
        out = self.empty_tuples(rel.numel())
        counts = (C.c_uint64 * world)()
        h = self._wrap(rel)
        cargs = slice_args.to_c() if slice_args is not None else None
        rc = self.L.hwbrj_owner_partition(h, world, C.byref(cargs) if cargs is not None else None, out.data_ptr(), counts)
        self.L.hwbrj_rel_free(h)
        if rc != 0:
            raise RuntimeError("hwbrj_owner_partition failed")
        return out, [int(c) for c in counts]

    def filter_build(self, rel: torch.Tensor, bloom: BloomFilterArgs) -> torch.Tensor:
        filt = torch.empty(max(bloom.m // 8, 16), dtype=torch.uint8, device=self.device)
        h = self._wrap(rel)
        cargs = bloom.to_c()
        rc = self.L.hwbrj_filter_build(h, C.byref(cargs), filt.data_ptr(), 1)
        self.L.hwbrj_rel_free(h)
        if rc != 0:
            raise RuntimeError("hwbrj_filter_build failed")
        return filt

    def filter_or(self, dst: torch.Tensor, src: torch.Tensor) -> None:
        if dst.numel() % 16 == 0:
            self.L.hwbrj_filter_or(dst.data_ptr(), src.data_ptr(), dst.numel())
        else:
            dst |= src

    def filter_probe(self, filt: torch.Tensor, rel: torch.Tensor, bloom: BloomFilterArgs) -> torch.Tensor:
        out = self.empty_tuples(rel.numel())
        h = self._wrap(rel)
        cargs = bloom.to_c()
        n = self.L.hwbrj_filter_probe(filt.data_ptr(), h, C.byref(cargs), out.data_ptr())
        self.L.hwbrj_rel_free(h)
        if n < 0:
            raise RuntimeError("hwbrj_filter_probe failed")
        return out[:n]

    def join(self, R: torch.Tensor, S: torch.Tensor) -> dict:
        hr, hs = self._wrap(R), self._wrap(S)
        st = N.StatsT()
        rc = self.L.hwbrj_join_device(hr, hs, None, C.byref(st))
        self.L.hwbrj_rel_free(hr)
        self.L.hwbrj_rel_free(hs)
        if rc != 0:
            raise RuntimeError("hwbrj_join_device failed")
        return st.as_dict()


def _all_gather_counts(counts, world, device, group):
    t = torch.tensor(counts, dtype=torch.int64, device=device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return [o.tolist() for o in outs]  # matrix[src][dst]


def exchange(ops, send: torch.Tensor, counts, group=None) -> torch.Tensor:
    """all-to-all of tuples grouped by destination (send is laid out as [to rank 0 | to rank 1 | ...])"""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return send
    matrix = _all_gather_counts(counts, world, send.device, group)
    recv_counts = [matrix[src][rank] for src in range(world)]
    out = ops.empty_tuples(sum(recv_counts))
    dist.all_to_all_single(out, send, output_split_sizes=recv_counts, input_split_sizes=list(counts), group=group)
    return out


def combine_filter(ops, filt: torch.Tensor, bloom: BloomFilterArgs, is_sliced: bool, group=None) -> torch.Tensor:
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return filt
    nbytes = filt.numel()
    if is_sliced:
        sl = bloom.m // 8 // world
        mine = filt[rank * sl:(rank + 1) * sl].clone()
        parts = [filt[g * sl:(g + 1) * sl] for g in range(world)]
        dist.all_gather(parts, mine, group=group)  # writes every rank's slice into place: a replicated filter
        return filt
    parts = [torch.empty_like(filt) if g != rank else filt for g in range(world)]
    mine = filt.clone()
    dist.all_gather(parts, mine, group=group)
    for g in range(world):
        if g != rank:
            ops.filter_or(filt, parts[g])
    assert filt.numel() == nbytes
    return filt


def _reduce_scalars(vals_u64, device, group):
    """exact sums mod 2^64: every value travels as two 32-bit halves in int64 lanes"""
    halves = []
    for v in vals_u64:
        v &= MASK64
        halves += [v & 0xFFFFFFFF, v >> 32]
    t = torch.tensor(halves, dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    h = t.tolist()
    return [((h[2 * i + 1] << 32) + h[2 * i]) & MASK64 for i in range(len(vals_u64))]


class PhaseTimer:
    """CUDA events on the (shared) current stream around each phase; no-op on CPU tensors."""

    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.marks = []

    def mark(self, name: str):
        if self.enabled:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def phases_ms(self) -> dict:
        out = {}
        for (_, a), (name, b) in zip(self.marks, self.marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


def dist_join(ops, Rshard: torch.Tensor, Sshard: torch.Tensor, bloom: Optional[BloomFilterArgs], group=None,
              time_phases: bool = False) -> dict:
    """Collective join of the ranks' shards; returns the global scalars (identical on every rank)."""
    if bloom is not None:
        bloom.check()
    world = dist.get_world_size(group)
    is_sliced = sliceable(bloom, world)
    slice_args = bloom if is_sliced else None
    info = {"sliced_filter": is_sliced, "world": world}
    tm = PhaseTimer(time_phases and Rshard.is_cuda)

    tm.mark("start")
    Rsend, cntR = ops.owner_partition(Rshard, world, slice_args)
    tm.mark("route_r_partition")
    Rown = exchange(ops, Rsend, cntR, group)
    tm.mark("route_r_all_to_all")
    info["r_sent"] = sum(cntR) - cntR[dist.get_rank(group)]
    filtered_local = 0
    if bloom is not None:
        filt = ops.filter_build(Rown, bloom)
        tm.mark("filter_build")
        filt = combine_filter(ops, filt, bloom, is_sliced, group)
        tm.mark("filter_all_gather")
        Ssurv = ops.filter_probe(filt, Sshard, bloom)
        tm.mark("s_probe")
        filtered_local = int(Ssurv.numel())
    else:
        Ssurv = Sshard
    Ssend, cntS = ops.owner_partition(Ssurv, world, slice_args)
    tm.mark("route_s_partition")
    Sown = exchange(ops, Ssend, cntS, group)
    tm.mark("route_s_all_to_all")
    info["s_sent"] = sum(cntS) - cntS[dist.get_rank(group)]
    st = ops.join(Rown, Sown)
    tm.mark("local_join")
    if tm.enabled:
        torch.cuda.synchronize()
        info["phases_ms"] = tm.phases_ms()
    vals = _reduce_scalars([st["matches"], filtered_local, st["checksum_pair"], st["checksum_rpay"],
                            st["checksum_spay"], st["checksum_key"], info["r_sent"], info["s_sent"]],
                           Rshard.device, group)
    return {"matches": vals[0], "filtered": vals[1] if bloom is not None else -1, "checksum_pair": vals[2],
            "checksum_rpay": vals[3], "checksum_spay": vals[4], "checksum_key": vals[5],
            "tuples_over_nvlink_r": vals[6], "tuples_over_nvlink_s": vals[7], "local": st, **info}


# ----------------------------------------------------------------------------------------------------------------
# NVLink peer-memory path: the whole collective join runs below the C ABI (hwbrj_dist_*). torch.distributed is only the
# launcher's channel for the 128-byte handles; the exchanges are peer-memory loads/stores issued by the library's own
# kernels, ranks meet at device-side barriers, and nothing here touches the data.
# ----------------------------------------------------------------------------------------------------------------
class DistGroup:
    """This rank's membership in a group of GPUs that join together (hwbrj_dist_t).

    The library allocates one symmetric block per rank (receive buffers for the level-1 routing of R and of the filter
    survivors, the replicated filter, gathered histogram / result rows, barrier flags) and exports a handle; the handles
    travel through torch.distributed.all_gather and every rank maps its peers (CUDA IPC). `join` is then ONE call into
    the library per rank."""

    def __init__(self, ops: "CudaOps", cap_r: int, cap_s: int, max_filter_bytes: int, group=None):
        self.ops, self.group = ops, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        L = ops.L
        buf = (C.c_ubyte * N.DIST_HANDLE_BYTES)()
        self.h = L.hwbrj_dist_create(self.rank, self.world, int(cap_r), int(cap_s), int(max_filter_bytes), buf)
        ok = 1 if self.h else 0  # a failed allocation is agreed on collectively below: nobody is left waiting
        mine = torch.frombuffer(bytearray(buf), dtype=torch.uint8).to(ops.device)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        blob = b"".join(t.cpu().numpy().tobytes() for t in allh)
        if ok and L.hwbrj_dist_connect(self.h, blob) != 0:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=ops.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            if self.h:
                L.hwbrj_dist_destroy(self.h)
                self.h = None
            raise RuntimeError("peer memory of the GPU group cannot be allocated or mapped on at least one rank")

    def join(self, Rshard: torch.Tensor, Sshard: torch.Tensor, bloom: Optional[BloomFilterArgs], r_total: int) -> Optional[dict]:
        """Collective join of the ranks' chunks; the scalars are global and identical on every rank. Returns None when
        a receive buffer was too small (heavily skewed owners) or a peer did not arrive."""
        if bloom is not None:
            bloom.check()
        L = self.ops.L
        L.hwbrj_set_stream(torch.cuda.current_stream(self.ops.device).cuda_stream)
        hr, hs = self.ops._wrap(Rshard), self.ops._wrap(Sshard)
        cargs = bloom.to_c() if bloom is not None else None
        st = N.StatsT()
        rc = L.hwbrj_dist_join(self.h, hr, hs, C.byref(cargs) if cargs is not None else None, int(r_total), C.byref(st))
        L.hwbrj_rel_free(hr)
        L.hwbrj_rel_free(hs)
        if rc == -2:
            return None
        if rc != 0:
            raise RuntimeError(f"hwbrj_dist_join rc={rc}")
        d = st.as_dict()
        return {"matches": d["matches"], "filtered": d["filtered"], "checksum_pair": d["checksum_pair"],
                "checksum_rpay": d["checksum_rpay"], "checksum_spay": d["checksum_spay"], "checksum_key": d["checksum_key"],
                "world": self.world, "owned_r": d["owned_r"], "owned_s": d["owned_s"], "local": d,
                "path": "nvlink-peer-memory (in-library)"}

    def filter_bytes(self, nbytes: int) -> bytes:
        """this rank's copy of the replicated filter (tests: it must be byte-identical on every rank)"""
        ptr = self.ops.L.hwbrj_dist_filter(self.h)
        return self.ops.view_int64(ptr, nbytes // 8).cpu().numpy().tobytes()

    def close(self):
        if self.h:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)  # nobody unmaps memory a peer may still write to
            self.ops.L.hwbrj_dist_destroy(self.h)
            self.h = None
            dist.barrier(group=self.group)


class DistJoinGraph:
    """DistGroup.join captured as ONE CUDA graph for fixed chunk tensors (kernels, NVLink stores and device-side barriers;
    no host work between them); replay() runs one join and returns the global scalars (None on overflow / time-out)."""

    def __init__(self, grp: DistGroup, Rshard: torch.Tensor, Sshard: torch.Tensor, bloom: Optional[BloomFilterArgs], r_total: int):
        if bloom is not None:
            bloom.check()
        self.grp, self.bloom = grp, bloom
        self.keep = [Rshard, Sshard]
        ops, L = grp.ops, grp.ops.L
        self.out8 = torch.zeros(8, dtype=torch.int64, device=ops.device)
        cargs = bloom.to_c() if bloom is not None else None
        cref = C.byref(cargs) if cargs is not None else None
        hr, hs = ops._wrap(Rshard), ops._wrap(Sshard)

        def enqueue():
            L.hwbrj_set_stream(torch.cuda.current_stream(ops.device).cuda_stream)
            n = L.hwbrj_dist_join_async(grp.h, hr, hs, cref, int(r_total), self.out8.data_ptr())
            if n < 0:
                raise RuntimeError("hwbrj_dist_join_async failed")
            return n
        side = torch.cuda.Stream(device=ops.device)
        side.wait_stream(torch.cuda.current_stream(ops.device))
        with torch.cuda.stream(side):  # warm-up on a side stream: workspace allocations happen here, not during capture
            for _ in range(2):
                enqueue()
        torch.cuda.current_stream(ops.device).wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.launches = enqueue()
        L.hwbrj_set_stream(torch.cuda.current_stream(ops.device).cuda_stream)
        L.hwbrj_rel_free(hr)
        L.hwbrj_rel_free(hs)

    def replay(self) -> Optional[dict]:
        self.graph.replay()
        v = [x & MASK64 for x in self.out8.tolist()]  # the only host synchronisation of the join
        if v[6]:
            return None
        return {"matches": v[0], "filtered": v[5] if self.bloom is not None else -1, "checksum_pair": v[1],
                "checksum_rpay": v[2], "checksum_spay": v[3], "checksum_key": v[4], "world": self.grp.world,
                "path": "nvlink-peer-memory (in-library) + cuda-graph", "local": {"kernel_launches": self.launches}}
