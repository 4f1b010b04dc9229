"""bench.py's N>1 arm: launched by torchrun with one rank per GPU (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* from the
environment). Strong scaling: the BASELINE workload is split into contiguous chunks over the ranks; `value` is
(|R|+|S|) over the max-over-ranks device time of the collective join, inputs resident in HBM.

torch is the launcher here: process group for the start-up handle exchange and the max-over-ranks reductions of the
timings, CUDA events and the CUDA graph. The join itself -- kernels, NVLink peer stores, device-side barriers -- is one
call into the library per rank (hwbrj_dist_join / hwbrj_dist_join_async)."""
from __future__ import annotations

import json
import os
import sys
import threading
import time

import torch
import torch.distributed as dist


def bind_to_gpu_numa(local: int) -> str:
    """Run this rank (and first-touch its pinned host buffers) on the CPUs of the NUMA node its GPU hangs off: with eight
    ranks copying from pinned memory at once, buffers that all sit on one socket are limited by that socket's DRAM and
    by the inter-socket link instead of by the eight PCIe links. Returns a description for the bench line."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        cpus = open(f"{base}/local_cpulist").read().strip()
        node = open(f"{base}/numa_node").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= os.sched_getaffinity(0)
        if not ids:
            return f"gpu {bdf}: numa {node}, no usable cpu in {cpus}"
        os.sched_setaffinity(0, ids)
        return f"gpu {bdf}: numa node {node}, {len(ids)} cpus"
    except Exception as exc:  # not fatal: the copies are merely slower
        return f"unbound ({exc})"


def main(args, wl, METRIC, UNIT, measured_peak, ClockSampler) -> int:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        if rank == 0:
            print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun with one rank per GPU"}))
        return 1
    watchdog = threading.Timer(float(os.environ.get("HWBRJ_BENCH_WATCHDOG_S", "420")), lambda: os._exit(3))
    watchdog.daemon = True  # never let a stuck rank hold the box
    watchdog.start()
    # NCCL prints its version banner on stdout; the contract is ONE JSON line there, so everything else goes to stderr
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if os.environ.get("HWBRJ_NUMA_BIND", "1") == "1" else "off"
    dist.init_process_group("nccl", device_id=device)
    from . import BloomFilterArgs
    from .dist import CudaOps, DistGroup, DistJoinGraph, dist_join
    ops = CudaOps(device)
    r, s, q, variant, m, k, B, desc = wl
    bloom = BloomFilterArgs(variant, m, k, B) if variant is not None else None

    def chunk(n):
        per = (n // world) & ~1  # even chunk starts keep every chunk 16-byte aligned
        lo = rank * per
        return lo, (n - lo if rank == world - 1 else per)
    rlo, rcnt = chunk(r)
    slo, scnt = chunk(s)
    Rsh = ops.generate_shard(0, r, r, 1.0, 1, rlo, rcnt)
    Ssh = ops.generate_shard(2 if q < 0 else 1, s, r, -q if q < 0 else q, 2, slo, scnt)  # q < 0: Zipf exponent -q

    def allmax(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # The GPU group: receive buffers sized for 25 % imbalance of R (hash owners of distinct keys) and for the worst case
    # of the probe side (every S tuple survives and one owner gets them all: a Zipf hot key), so nothing can overflow.
    use_peer = os.environ.get("HWBRJ_DIST_PATH", "peer") == "peer"
    grp = None
    if use_peer:
        try:  # fails on ALL ranks together when peer memory cannot be mapped -> NCCL reference path
            grp = DistGroup(ops, int(r / world * 1.25) + 65536, s + 2, max(m // 8, 16) if bloom is not None else 16)
        except Exception as exc:
            print(f"[bench] NVLink peer path unavailable ({exc}); using the NCCL all-to-all path", flush=True)

    def step(Rt, St):
        if grp is not None:
            out = grp.join(Rt, St, bloom, r)
            if out is not None:
                return out
        out = dist_join(ops, Rt, St, bloom)
        out["path"] = "nccl-all-to-all"
        return out

    # device-resident leg: the whole collective join replayed as one CUDA graph
    graph = None
    if grp is not None and os.environ.get("HWBRJ_DIST_GRAPH", "1") == "1":
        try:
            graph = DistJoinGraph(grp, Rsh, Ssh, bloom, r)
            if graph.replay() is None:
                graph = None
        except Exception as exc:  # capture not possible here: run the eager pipeline
            print(f"[bench] CUDA graph capture unavailable ({exc}); using the eager pipeline", flush=True)
            graph = None

    def resident_step():
        if graph is not None:
            out = graph.replay()
            if out is not None:
                return out
        return step(Rsh, Ssh)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a few hundred ms before its first sample: spawn it before the warm-up
    for _ in range(max(args.warmup, 3)):
        res = resident_step()
    if rank == 0:
        sampler.begin()  # waits for the first sample, then opens the sampling window
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        res = resident_step()
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = allmax(ev0.elapsed_time(ev1)) / args.steps
    value = (r + s) / (ms_per_step * 1e-3) / 1e6

    # per-phase device times and per-GPU load: eager joins (the library's own CUDA events on its stream), mean over the
    # steps, max over the ranks; the ranks run in lock step because every phase boundary that matters is a device barrier
    phases, loads = {}, None
    if grp is not None:
        acc = {}
        n_eager = min(args.steps, 5)
        for _ in range(n_eager):
            e = step(Rsh, Ssh)
            for kk, vv in e["local"].items():
                if kk.startswith("ms_"):
                    acc[kk] = acc.get(kk, 0.0) + vv / n_eager
        phases = {kk: allmax(vv) for kk, vv in sorted(acc.items())}
        t = torch.zeros(2 * world, dtype=torch.int64, device=device)
        t[2 * rank], t[2 * rank + 1] = e["owned_r"], e["owned_s"]
        dist.all_reduce(t)
        loads = {"owned_r_per_gpu": t[0::2].tolist(), "owned_s_per_gpu": t[1::2].tolist()}
        launches = e["local"]["kernel_launches"]
    else:
        launches = res.get("local", {}).get("kernel_launches", 0) + 8

    # ---- host-buffer leg: pinned host shards -> device, collective join, scalars back ----
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    hR = torch.empty(Rsh.numel(), dtype=torch.int64).pin_memory()
    hS = torch.empty(Ssh.numel(), dtype=torch.int64).pin_memory()
    hR.copy_(Rsh)
    hS.copy_(Ssh)
    dR, dS = ops.empty_tuples(Rsh.numel()), ops.empty_tuples(Ssh.numel())

    def host_step():
        dR.copy_(hR, non_blocking=True)
        dS.copy_(hS, non_blocking=True)
        return step(dR, dS)
    host_step()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hres = host_step()
    torch.cuda.synchronize()
    e2e_s = allmax((time.perf_counter() - t0) / e2e_steps)
    assert hres["matches"] == res["matches"]

    if rank == 0:
        peak, peak_src = measured_peak()
        F = res["filtered"] if bloom is not None else s
        b_alg = (24 * r + 8 * s + 16 * F + 2 * (m // 8)) if bloom is not None else (24 * r + 24 * s)
        probe_ms = phases.get("ms_probe") if bloom is not None else None
        roofline = {"bound": "hbm", "kernel": "k_probe_compact (K2) on every rank's S chunk", "unit": "GB/s",
                    "peak": peak * world, "peak_source": peak_src + f" x {world} GPUs",
                    # every rank reads its S chunk, writes its survivors and reads the whole replicated filter
                    "achieved": (8 * s + 8 * F + world * (m // 8)) / (probe_ms * 1e-3) / 1e9 if probe_ms else None,
                    "traffic": None,
                    "note": "phase times: mean of eager joins, library CUDA events, max over ranks; whole_join from graph replays",
                    "whole_join": {"algorithmic_bytes": b_alg, "achieved": b_alg / (ms_per_step * 1e-3) / 1e9,
                                   "frac": b_alg / (ms_per_step * 1e-3) / 1e9 / (peak * world)}}
        if roofline["achieved"]:
            roofline["frac"] = roofline["achieved"] / roofline["peak"]
        # NVLink: (G-1)/G of R and of the survivors leave their rank; every rank stores its filter slice into G-1 peers
        moved = (r + F) * (world - 1) // world
        filt_bytes = (world - 1) * (m // 8) if bloom is not None else 0
        nvl_bytes = 8 * moved + filt_bytes
        nvlink = {"bytes_total_per_step": nvl_bytes, "bytes_per_gpu_per_step": nvl_bytes // world,
                  "gbps_per_gpu_over_whole_step": nvl_bytes / world / (ms_per_step * 1e-3) / 1e9,
                  "frac_of_900_GBps_per_direction": nvl_bytes / world / (ms_per_step * 1e-3) / 1e9 / 900.0}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": desc, "name": args.workload, "r": r, "s": s, "q": q if q >= 0 else None, "zipf": -q if q < 0 else None,
                           "bloom": None if bloom is None else {"variant": "basic" if variant == 0 else "blocked", "m": m, "k": k, "B": B},
                           "sharding": "contiguous input chunks per rank; a key's owner = high bits of its partition id "
                                       "(the rank holding its filter slice)",
                           "exchange": res.get("path", "nccl-all-to-all"),
                           "l2": "per-rank inputs exceed the 126 MB L2; no flush needed"},
                "results": {"matches": res["matches"], "filtered": res["filtered"], "checksum_pair": res["checksum_pair"]},
                "phases_ms_max_over_ranks": phases, "per_gpu_load": loads, "wall_ms_per_step": wall / args.steps * 1e3,
                "nvlink": nvlink, "roofline": roofline, "cpu_baseline": None,
                "e2e": {"value": (r + s) / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * (r + s),
                        "d2h_bytes_per_step": 8 * 16 * world, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                        "api": "hwbrj_dist_join (C ABI) on chunks copied from pinned host memory", "host_numa_binding_rank0": numa},
                "gpu_launches": int(args.steps * world * launches), "clocks": clocks,
                "cuda_graph": graph is not None}
        os.write(saved_stdout, (json.dumps(line) + "\n").encode())
    # No collective teardown: every rank has passed the last all-reduce, rank 0 has printed its line. Tearing down the
    # captured graph, the IPC mappings and the NCCL communicator in lock step gains nothing here and a rank that
    # waits for a peer which is already gone would hang the launcher, so the processes simply exit.
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)
