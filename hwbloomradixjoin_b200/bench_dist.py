"""bench.py's N>1 arm: launched by torchrun with one rank per GPU (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* from the
environment). Strong scaling: the BASELINE workload is split into contiguous chunks over the ranks; `value` is
(|R|+|S|) over the max-over-ranks device time of the collective join, inputs resident in HBM."""
from __future__ import annotations

import json
import os
import sys
import threading
import statistics
import time

import torch
import torch.distributed as dist


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
    dist.init_process_group("nccl", device_id=device)
    from . import BloomFilterArgs
    from .dist import CudaOps, PeerFabric, PeerJoinGraph, dist_join, dist_join_peer
    ops = CudaOps(device)
    r, s, q, variant, m, k, B, desc = wl
    bloom = BloomFilterArgs(variant, m, k, B) if variant is not None else None

    def chunk(n):
        per = n // world
        lo = rank * per
        return lo, (n - lo if rank == world - 1 else per)
    rlo, rcnt = chunk(r)
    slo, scnt = chunk(s)
    Rsh = ops.generate_shard(0, r, r, 1.0, 1, rlo, rcnt)
    Ssh = ops.generate_shard(2 if q < 0 else 1, s, r, -q if q < 0 else q, 2, slo, scnt)  # q < 0: Zipf exponent -q

    # exchanges fused into the partitioning kernels (NVLink peer stores); NCCL all-to-all is the fallback
    use_peer = os.environ.get("HWBRJ_DIST_PATH", "peer") == "peer"
    fabric = None
    if use_peer:
        try:  # the constructor fails on ALL ranks together when peer memory cannot be mapped -> NCCL all-to-all path
            # receive capacity per rank: 20 % over the even share; a Zipf probe relation sends its hot keys (all of
            # them survive the filter) to single owners, so leave 2.5x there (overflow falls back to NCCL anyway)
            s_slack = 2.5 if q < 0 else 1.2
            fabric = PeerFabric(ops, int(r / world * 1.2) + 65536, int(s / world * s_slack) + 65536)
        except Exception as exc:
            print(f"[bench] NVLink peer path unavailable ({exc}); using NCCL all-to-all", flush=True)

    def step(Rt, St, time_phases=False):
        if fabric is not None:
            out = dist_join_peer(ops, fabric, Rt, St, bloom, r, s, time_phases=time_phases)
            if out is not None:
                return out
        out = dist_join(ops, Rt, St, bloom, time_phases=time_phases)
        out["path"] = "nccl-all-to-all"
        return out

    # device-resident leg: the whole collective join replayed as one CUDA graph (kernels + NVLink stores + NCCL)
    graph = None
    if fabric is not None and os.environ.get("HWBRJ_DIST_GRAPH", "1") == "1":
        try:
            graph = PeerJoinGraph(ops, fabric, Rsh, Ssh, bloom, r, s)
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
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # max over ranks
    ms_per_step = ms.item() / args.steps
    value = (r + s) / (ms_per_step * 1e-3) / 1e6
    phased = step(Rsh, Ssh, time_phases=True)

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
    e2e_t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=device)
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    assert hres["matches"] == res["matches"]

    if rank == 0:
        peak, peak_src = measured_peak()
        F = res["filtered"] if bloom is not None else s
        b_alg = (24 * r + 8 * s + 16 * F + 2 * (m // 8)) if bloom is not None else (24 * r + 24 * s)
        ph = phased.get("phases_ms", {})
        probe_ms = ph.get("s_probe")
        roofline = {"bound": "hbm", "kernel": "k_probe_compact (K2) on the rank's S chunk", "unit": "GB/s",
                    "peak": peak * world, "peak_source": peak_src + f" x {world} GPUs",
                    "achieved": (8 * s + 8 * F + world * (m // 8)) / (probe_ms * 1e-3) / 1e9 if probe_ms else None,
                    "traffic": None,
                    "whole_join": {"algorithmic_bytes": b_alg, "achieved": b_alg / (ms_per_step * 1e-3) / 1e9,
                                   "frac": b_alg / (ms_per_step * 1e-3) / 1e9 / (peak * world)}}
        if roofline["achieved"]:
            roofline["frac"] = roofline["achieved"] / roofline["peak"]
        moved = (r + F) * (world - 1) // world  # hash owners: (G-1)/G of R and of the survivors leave their rank
        nvl_bytes = 8 * moved + (world - 1) * (m // 8 if bloom is not None else 0) * (1 if res["sliced_filter"] else world)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": desc, "name": args.workload, "r": r, "s": s, "q": q if q >= 0 else None, "zipf": -q if q < 0 else None,
                           "bloom": None if bloom is None else {"variant": "basic" if variant == 0 else "blocked", "m": m, "k": k, "B": B},
                           "sharding": "contiguous chunks per rank; owner = filter-slice rank" if res["sliced_filter"] else "contiguous chunks per rank; owner = crapwow top bits",
                           "exchange": res.get("path", "nccl-all-to-all"),
                           "l2": "per-rank inputs exceed the 126 MB L2; no flush needed"},
                "results": {"matches": res["matches"], "filtered": res["filtered"], "checksum_pair": res["checksum_pair"]},
                "phases_ms_rank0": ph, "wall_ms_per_step": wall / args.steps * 1e3, "nvlink_bytes_total": nvl_bytes,
                "roofline": roofline, "cpu_baseline": None,
                "e2e": {"value": (r + s) / e2e_t.item() / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * (r + s),
                        "d2h_bytes_per_step": 8 * 16 * world, "ms_per_step": e2e_t.item() * 1e3, "steps": e2e_steps,
                        "api": "hwbloomradixjoin_b200.dist.dist_join on pinned host shards"},
                "gpu_launches": int(args.steps * world * (phased["local"]["kernel_launches"] + 8)), "clocks": clocks,
                "cuda_graph": graph is not None}
        os.write(saved_stdout, (json.dumps(line) + "\n").encode())
    # No collective teardown: every rank has passed the last all-reduce, rank 0 has printed its line. Tearing down the
    # captured graph, the IPC mappings and the NCCL communicator in lock step gains nothing here and a rank that
    # waits for a peer which is already gone would hang the launcher, so the processes simply exit.
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)
