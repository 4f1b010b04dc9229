"""world_size-2/4 CPU tests (gloo) of the multi-GPU host logic in hwbloomradixjoin_b200/dist.py: with an
oracle-backed local-compute object the sharded pipeline must give the same scalars as the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(oracle, zipf):
    R = oracle.gen_R(60_000, nthreads=2)
    if zipf:  # BASELINE config 5: hot foreign keys, every probe tuple has a partner and passes the filter
        S = oracle.gen_zipf(300_001, 60_000, 1.0)
    else:
        S = oracle.gen_S(400_001, 60_000, 0.05, nthreads=2)
    return R, S


def _worker(rank, world, port, cases, out_q, zipf=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from dist_oracle_ops import OracleOps, _t
    from hwbloomradixjoin_b200 import BloomFilterArgs
    from hwbloomradixjoin_b200.dist import dist_join
    ops = OracleOps()
    R, S = _inputs(oracle, zipf)
    # contiguous chunks like the reference's per-thread split (last rank takes the remainder, :1646-1670)
    def chunk(a):
        per = a.shape[0] // world
        lo = rank * per
        hi = a.shape[0] if rank == world - 1 else lo + per
        return _t(a[lo:hi])
    results = []
    for case in cases:
        bloom = BloomFilterArgs(*case) if case is not None else None
        res = dist_join(ops, chunk(R), chunk(S), bloom)
        results.append({k: res[k] for k in ("matches", "filtered", "checksum_pair", "checksum_rpay", "checksum_spay",
                                            "checksum_key", "sliced_filter", "tuples_over_nvlink_s")})
    if rank == 0:
        out_q.put(results)
    dist.barrier()
    dist.destroy_process_group()


CASES = [(0, 1 << 19, 1, 512), (0, 1 << 19, 3, 512), (1, 1 << 19, 4, 256), (1, 1 << 19, 2, 1 << 19), None]


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_join_matches_single_process_oracle(world, oracle_mod):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, CASES, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    R = oracle_mod.gen_R(60_000, nthreads=2)
    S = oracle_mod.gen_S(400_001, 60_000, 0.05, nthreads=2)
    for case, got in zip(CASES, results):
        if case is None:
            exp = oracle_mod.join(R, S, False)
            assert got["filtered"] == -1
        else:
            exp = oracle_mod.join(R, S, True, *case)
            assert got["filtered"] == exp["filtered"], case
        for f in ("matches", "checksum_pair", "checksum_rpay", "checksum_spay", "checksum_key"):
            assert got[f] == exp[f], (case, f)
        # BASIC k>1 cannot be sliced by filter range; a BLOCKED filter with a single block cannot either
        expect_sliced = case is not None and ((case[0] == 0 and case[2] <= 1) or (case[0] == 1 and case[1] // case[3] >= world))
        assert got["sliced_filter"] == expect_sliced
        if case is not None:  # the filter cuts the exchanged volume: only survivors cross the wire
            assert got["tuples_over_nvlink_s"] <= exp["filtered"]


def test_sharded_join_with_zipf_skew(oracle_mod):
    """owners receive very different numbers of probe tuples (the hottest key alone is ~9 % of S): the exchange must
    cope with ragged counts, and since every Zipf key exists in R all of S matches and survives the filter"""
    world, cases = 4, [(0, 1 << 19, 1, 512), (1, 1 << 19, 4, 256), None]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q, True)) for r in range(world)]
    for p in procs:
        p.start()
    results = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    R, S = _inputs(oracle_mod, True)
    for case, got in zip(cases, results):
        exp = oracle_mod.join(R, S, case is not None, *(case or ()))
        assert got["matches"] == exp["matches"] == S.shape[0]
        assert got["filtered"] == (S.shape[0] if case is not None else -1)
        for f in ("checksum_pair", "checksum_rpay", "checksum_spay", "checksum_key"):
            assert got[f] == exp[f], (case, f)


def test_scalar_reduction_is_exact_mod_2_64():
    from hwbloomradixjoin_b200.dist import MASK64
    vals = [MASK64, 1 << 63, 12345678901234567890]
    halves = []
    for v in vals:
        halves += [v & 0xFFFFFFFF, v >> 32]
    world = 8
    summed = [h * world for h in halves]
    rec = [((summed[2 * i + 1] << 32) + summed[2 * i]) & MASK64 for i in range(len(vals))]
    assert rec == [(v * world) & MASK64 for v in vals]


def test_sliceable_rule():
    from hwbloomradixjoin_b200 import BloomFilterArgs
    from hwbloomradixjoin_b200.dist import sliceable
    assert sliceable(BloomFilterArgs(0, 1 << 30, 1, 512), 8)
    assert not sliceable(BloomFilterArgs(0, 1 << 30, 2, 512), 8)
    assert sliceable(BloomFilterArgs(1, 1 << 30, 6, 512), 8)
    assert not sliceable(BloomFilterArgs(1, 1 << 10, 6, 1 << 10), 2)
    assert not sliceable(None, 4)
