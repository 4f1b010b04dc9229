"""Multi-GPU Bloom-filter radix join: one process per GPU, torch.distributed (NCCL over NVLink /
NVSwitch) for the exchange steps, the library's CUDA kernels for everything else.

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
# NVLink peer-memory path: the exchanges are done BY the partitioning kernel (fused partition + all-to-all)
# ----------------------------------------------------------------------------------------------------------------
class PeerFabric:
    """Per-rank receive buffers that every peer of the NVLink domain can store into.

    The buffers are allocated by the library (cudaMalloc) and exported as CUDA IPC handles; the 64-byte handles
    travel through torch.distributed and each rank maps its peers' buffers. `hwbrj_route_peer` then writes every
    tuple straight into its owner's buffer, claiming space from the owner's cursor with a system-scope atomic over
    NVLink: no send buffers, no counts on the host, no collective for the data. Ranks are separated by tiny
    stream-ordered all-reduces (route -> consume)."""

    # two sets of control words {R cursor u64, S cursor u64, overflow u32, pad}: join i uses set i%2 and zeroes the other
    # one, which no peer touches before join i+1 -- and every rank enters join i+1 only after the barriers of join i,
    # i.e. after this rank's zeroing (stream order). So no extra "cursors are zero" barrier is needed.
    CTRL_BYTES = 256
    SET_BYTES = 32

    def __init__(self, ops: "CudaOps", cap_r: int, cap_s: int, group=None):
        self.ops, self.group = ops, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.cap_r, self.cap_s = int(cap_r), int(cap_s)
        L = ops.L
        self.local = [L.hwbrj_symm_alloc((self.cap_r + 8) * 8), L.hwbrj_symm_alloc((self.cap_s + 8) * 8),
                      L.hwbrj_symm_alloc(self.CTRL_BYTES)]
        handles = torch.zeros(3 * N_IPC, dtype=torch.uint8)
        ok = 1 if all(self.local) else 0  # a failed allocation is reported through the collective flag below as well
        for i, p in enumerate(self.local):
            buf = (C.c_ubyte * N_IPC)()
            if not ok or L.hwbrj_ipc_export(p, buf) != 0:
                ok = 0  # keep going: the failure is agreed on collectively below, nobody is left waiting
            handles[i * N_IPC:(i + 1) * N_IPC] = torch.frombuffer(bytearray(buf), dtype=torch.uint8)
        handles = handles.to(ops.device)
        allh = [torch.empty_like(handles) for _ in range(self.world)]
        dist.all_gather(allh, handles, group=group)
        self.peer = []  # peer[g] = [recvR, recvS, ctrl] pointers valid in this process
        for g in range(self.world):
            if g == self.rank:
                self.peer.append(list(self.local))
                continue
            hb = allh[g].cpu().numpy().tobytes()
            ptrs = []
            for i in range(3):
                raw = (C.c_ubyte * N_IPC).from_buffer_copy(hb[i * N_IPC:(i + 1) * N_IPC])
                p = L.hwbrj_ipc_open(raw) if ok else None
                if not p:
                    ok = 0
                ptrs.append(p)
            self.peer.append(ptrs)
        # agree collectively: if any rank could not map a peer buffer, every rank gives up on the peer path together
        flag = torch.tensor([ok], dtype=torch.int32, device=ops.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            raise RuntimeError("CUDA IPC peer mapping unavailable on at least one rank")
        self._bar = torch.zeros(1, dtype=torch.int32, device=ops.device)
        self.ctrl_view = ops.view_int64(self.local[2], self.CTRL_BYTES // 8)
        self.parity = 0
        self.ctrl_view.zero_()
        self.barrier()

    def barrier(self):
        """stream-ordered barrier across ranks (no host synchronisation)"""
        dist.all_reduce(self._bar, group=self.group)

    def ptr_array(self, which: int, byte_offset: int = 0):
        arr = (C.c_void_p * self.world)()
        for g in range(self.world):
            arr[g] = self.peer[g][which] + byte_offset
        return arr

    def begin_join(self) -> int:
        """switch to the other set of control words and zero the one the NEXT join will use; returns the byte offset
        of the active set"""
        self.parity ^= 1
        nxt = (self.parity ^ 1) * (self.SET_BYTES // 8)
        self.ctrl_view[nxt:nxt + self.SET_BYTES // 8].zero_()
        return self.parity * self.SET_BYTES

    def close(self):
        L = self.ops.L
        torch.cuda.synchronize()
        self.barrier()
        torch.cuda.synchronize()
        for g in range(self.world):
            if g != self.rank:
                for p in self.peer[g]:
                    L.hwbrj_ipc_close(p)
        self.barrier()
        torch.cuda.synchronize()
        for p in self.local:
            L.hwbrj_symm_free(p)


N_IPC = 64  # HWBRJ_IPC_HANDLE_BYTES


def dist_join_peer(ops: "CudaOps", fabric: PeerFabric, Rshard: torch.Tensor, Sshard: torch.Tensor,
                   bloom: Optional[BloomFilterArgs], r_total: int, s_total: int, time_phases: bool = False) -> Optional[dict]:
    """Collective join with the exchanges fused into the partitioning kernels (NVLink peer stores). Returns None when
    a receive buffer overflowed (heavily skewed owners): the caller then uses dist_join (NCCL all-to-all)."""
    if bloom is not None:
        bloom.check()
    group, world, rank = fabric.group, fabric.world, fabric.rank
    L = ops.L
    is_sliced = sliceable(bloom, world)
    slice_args = bloom if is_sliced else None
    cargs = slice_args.to_c() if slice_args is not None else None
    cref = C.byref(cargs) if cargs is not None else None
    tm = PhaseTimer(time_phases)
    tm.mark("start")
    coff = fabric.begin_join()
    ctrl = fabric.local[2] + coff
    # (1) R: fused partition + all-to-all
    h = ops._wrap(Rshard)
    rc = L.hwbrj_route_peer(h, world, cref, fabric.ptr_array(0), fabric.ptr_array(2, coff), fabric.cap_r, ctrl + 16)
    L.hwbrj_rel_free(h)
    if rc != 0:
        raise RuntimeError("hwbrj_route_peer(R) failed")
    fabric.barrier()
    tm.mark("route_r_fused")
    Rown = L.hwbrj_rel_wrap_counted(fabric.local[0], fabric.cap_r, ctrl, max(r_total // world, 1))
    if os.environ.get("HWBRJ_DIST_OVERLAP_R") == "1" and L.hwbrj_join_prepare_r(Rown) < 0:  # experimental, see below
        raise RuntimeError("hwbrj_join_prepare_r failed")
    # (2) filter slice / partial + combine, (3) local pre-filter
    if bloom is not None:
        filt = torch.empty(max(bloom.m // 8, 16), dtype=torch.uint8, device=ops.device)
        bc = bloom.to_c()
        if L.hwbrj_filter_build(Rown, C.byref(bc), filt.data_ptr(), 1) != 0:
            raise RuntimeError("hwbrj_filter_build failed")
        tm.mark("filter_build")
        filt = combine_filter(ops, filt, bloom, is_sliced, group)
        tm.mark("filter_all_gather")
        surv = ops.empty_tuples(Sshard.numel())
        cnt = torch.zeros(1, dtype=torch.int64, device=ops.device)
        hs = ops._wrap(Sshard)
        if L.hwbrj_filter_probe_async(filt.data_ptr(), hs, C.byref(bc), surv.data_ptr(), cnt.data_ptr()) != 0:
            raise RuntimeError("hwbrj_filter_probe_async failed")
        L.hwbrj_rel_free(hs)
        tm.mark("s_probe")
        hsurv = L.hwbrj_rel_wrap_counted(surv.data_ptr(), surv.numel(), cnt.data_ptr(), surv.numel())
    else:
        cnt = None
        hsurv = ops._wrap(Sshard)
    # (4) survivors: fused partition + all-to-all
    rc = L.hwbrj_route_peer(hsurv, world, cref, fabric.ptr_array(1), fabric.ptr_array(2, coff + 8), fabric.cap_s, ctrl + 16)
    L.hwbrj_rel_free(hsurv)
    if rc != 0:
        raise RuntimeError("hwbrj_route_peer(S) failed")
    fabric.barrier()
    tm.mark("route_s_fused")
    # (5) local join of owned R and owned survivors; counts stay on the device
    Sown = L.hwbrj_rel_wrap_counted(fabric.local[1], fabric.cap_s, ctrl + 8, max(s_total // world, 1))
    st = N.StatsT()
    rc = L.hwbrj_join_device(Rown, Sown, None, C.byref(st))  # synchronises the stream to fetch the scalars
    L.hwbrj_rel_free(Rown)
    L.hwbrj_rel_free(Sown)
    if rc != 0:
        raise RuntimeError("hwbrj_join_device failed")
    tm.mark("local_join")
    c = fabric.ctrl_view[coff // 8:coff // 8 + 4].tolist()
    filtered_local = int(cnt.item()) if cnt is not None else 0
    overflow = c[2] & 0xFFFFFFFF
    vals = _reduce_scalars([st.matches, filtered_local, st.checksum_pair, st.checksum_rpay, st.checksum_spay,
                            st.checksum_key, overflow, c[0], c[1]], ops.device, group)
    if vals[6]:
        return None
    out = {"matches": vals[0], "filtered": vals[1] if bloom is not None else -1, "checksum_pair": vals[2],
           "checksum_rpay": vals[3], "checksum_spay": vals[4], "checksum_key": vals[5], "sliced_filter": is_sliced,
           "world": world, "r_owned_total": vals[7], "s_owned_total": vals[8], "local": st.as_dict(),
           "path": "nvlink-peer-stores"}
    if tm.enabled:
        torch.cuda.synchronize()
        out["phases_ms"] = tm.phases_ms()
    return out


# ----------------------------------------------------------------------------------------------------------------
# The same pipeline as ONE CUDA graph: kernels, NVLink peer stores, barriers, filter all-gather and the final
# all-reduce are captured once and replayed per join, so there is no launch gap between the ~20 short kernels.
# ----------------------------------------------------------------------------------------------------------------
def _peer_pipeline_async(ops: "CudaOps", fabric: PeerFabric, Rshard, Sshard, bloom, r_total, s_total, keep):
    """dist_join_peer without any host synchronisation: returns the all-reduced int64 vector (device)
    [matches, filtered, cpair, crpay, cspay, ckey, overflow, r_owned, s_owned]. `keep` collects the tensors that
    must stay alive as long as the captured graph."""
    group, world = fabric.group, fabric.world
    L = ops.L
    L.hwbrj_set_stream(torch.cuda.current_stream(ops.device).cuda_stream)
    is_sliced = sliceable(bloom, world)
    slice_args = bloom if is_sliced else None
    cargs = slice_args.to_c() if slice_args is not None else None
    cref = C.byref(cargs) if cargs is not None else None
    ctrl = fabric.local[2]  # set 0 only: the graph starts with an explicit reset + barrier
    fabric.ctrl_view[0:4].zero_()
    fabric.barrier()
    h = ops._wrap(Rshard)
    if L.hwbrj_route_peer(h, world, cref, fabric.ptr_array(0), fabric.ptr_array(2, 0), fabric.cap_r, ctrl + 16) != 0:
        raise RuntimeError("hwbrj_route_peer(R) failed")
    L.hwbrj_rel_free(h)
    fabric.barrier()
    Rown = L.hwbrj_rel_wrap_counted(fabric.local[0], fabric.cap_r, ctrl, max(r_total // world, 1))
    cnt = torch.zeros(1, dtype=torch.int64, device=ops.device)
    keep.append(cnt)
    if os.environ.get("HWBRJ_DIST_OVERLAP_R") == "1":
        # experimental: the owned R is partitioned on the library's side stream while this stream builds, gathers and
        # probes the filter; the local join at the end picks the partitions up (hwbrj_join_prepare_r)
        if L.hwbrj_join_prepare_r(Rown) < 0:
            raise RuntimeError("hwbrj_join_prepare_r failed")
    if bloom is not None:
        filt = torch.empty(max(bloom.m // 8, 16), dtype=torch.uint8, device=ops.device)
        bc = bloom.to_c()
        if L.hwbrj_filter_build(Rown, C.byref(bc), filt.data_ptr(), 1) != 0:
            raise RuntimeError("hwbrj_filter_build failed")
        filt = combine_filter(ops, filt, bloom, is_sliced, group)
        surv = ops.empty_tuples(Sshard.numel())
        keep += [filt, surv]
        hs = ops._wrap(Sshard)
        if L.hwbrj_filter_probe_async(filt.data_ptr(), hs, C.byref(bc), surv.data_ptr(), cnt.data_ptr()) != 0:
            raise RuntimeError("hwbrj_filter_probe_async failed")
        L.hwbrj_rel_free(hs)
        hsurv = L.hwbrj_rel_wrap_counted(surv.data_ptr(), surv.numel(), cnt.data_ptr(), surv.numel())
    else:
        hsurv = ops._wrap(Sshard)
    if L.hwbrj_route_peer(hsurv, world, cref, fabric.ptr_array(1), fabric.ptr_array(2, 8), fabric.cap_s, ctrl + 16) != 0:
        raise RuntimeError("hwbrj_route_peer(S) failed")
    L.hwbrj_rel_free(hsurv)
    fabric.barrier()
    Sown = L.hwbrj_rel_wrap_counted(fabric.local[1], fabric.cap_s, ctrl + 8, max(s_total // world, 1))
    out6 = torch.zeros(8, dtype=torch.int64, device=ops.device)
    keep.append(out6)
    launches = L.hwbrj_join_device_async(Rown, Sown, None, out6.data_ptr())
    L.hwbrj_rel_free(Rown)
    L.hwbrj_rel_free(Sown)
    if launches < 0:
        raise RuntimeError("hwbrj_join_device_async failed")
    cv = fabric.ctrl_view
    vec = torch.stack([out6[0], cnt[0], out6[1], out6[2], out6[3], out6[4], cv[2] & 0xFFFFFFFF, cv[0], cv[1]])
    dist.all_reduce(vec, group=group)  # int64 lanes wrap modulo 2^64 exactly like the uint64 sums they carry
    keep.append(vec)
    return vec, is_sliced, launches


class PeerJoinGraph:
    """dist_join_peer captured as a CUDA graph for fixed shard tensors; replay() runs one join and returns the global
    scalars (or None on receive-buffer overflow)."""

    def __init__(self, ops: "CudaOps", fabric: PeerFabric, Rshard, Sshard, bloom, r_total, s_total):
        if bloom is not None:
            bloom.check()
        self.ops, self.bloom, self.keep = ops, bloom, [Rshard, Sshard]
        side = torch.cuda.Stream(device=ops.device)
        side.wait_stream(torch.cuda.current_stream(ops.device))
        with torch.cuda.stream(side):  # warm-up on a side stream: allocations, attributes, NCCL connections
            for _ in range(2):
                _peer_pipeline_async(ops, fabric, Rshard, Sshard, bloom, r_total, s_total, [])
        torch.cuda.current_stream(ops.device).wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.vec, self.is_sliced, self.launches = _peer_pipeline_async(ops, fabric, Rshard, Sshard, bloom, r_total,
                                                                           s_total, self.keep)
        ops.L.hwbrj_set_stream(torch.cuda.current_stream(ops.device).cuda_stream)
        self.world = fabric.world

    def replay(self) -> Optional[dict]:
        self.graph.replay()
        v = [x & MASK64 for x in self.vec.tolist()]  # the only host synchronisation of the join
        if v[6]:
            return None
        return {"matches": v[0], "filtered": v[1] if self.bloom is not None else -1, "checksum_pair": v[2],
                "checksum_rpay": v[3], "checksum_spay": v[4], "checksum_key": v[5], "sliced_filter": self.is_sliced,
                "world": self.world, "r_owned_total": v[7], "s_owned_total": v[8], "path": "nvlink-peer-stores+cuda-graph",
                "local": {"kernel_launches": self.launches}}
