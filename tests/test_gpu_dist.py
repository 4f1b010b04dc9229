"""2-GPU test of the sharded join (-m gpu; skipped on a single-GPU box): torchrun with two ranks over NCCL must
reproduce the single-GPU scalars. One rank per GPU -- ranks are never stacked on one device."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["HWBRJ_ROOT"])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from hwbloomradixjoin_b200 import BloomFilterArgs
from hwbloomradixjoin_b200.dist import CudaOps, PeerFabric, PeerJoinGraph, dist_join, dist_join_peer
ops = CudaOps(dev)
r, s, q = 2_000_000, 16_000_000, 0.01
per_r, per_s = r // world, s // world
R = ops.generate_shard(0, r, r, 1.0, 1, rank * per_r, r - rank * per_r if rank == world - 1 else per_r)
S = ops.generate_shard(1, s, r, q, 2, rank * per_s, s - rank * per_s if rank == world - 1 else per_s)
out, peer, graphed = [], [], []
fabric = PeerFabric(ops, int(r / world * 1.25) + 65536, int(s / world * 1.25) + 65536)
for case in [(0, 1 << 24, 1, 512), (0, 1 << 24, 3, 512), (1, 1 << 24, 4, 256), None]:
    bloom = BloomFilterArgs(*case) if case else None
    res = dist_join(ops, R, S, bloom)
    out.append({k: res[k] for k in ("matches", "filtered", "checksum_pair", "checksum_key", "sliced_filter", "tuples_over_nvlink_s")})
    for rep in range(2):  # twice: the cursors must be reset correctly between joins
        pres = dist_join_peer(ops, fabric, R, S, bloom, r, s)
    peer.append({k: pres[k] for k in ("matches", "filtered", "checksum_pair", "checksum_key", "r_owned_total", "s_owned_total")})
    pg = PeerJoinGraph(ops, fabric, R, S, bloom, r, s)  # the same pipeline captured as one CUDA graph
    for rep in range(3):
        gres = pg.replay()
    graphed.append({k: gres[k] for k in ("matches", "filtered", "checksum_pair", "checksum_key", "r_owned_total", "s_owned_total")})
    del pg
# a receive buffer that is too small must be reported (None), never silently truncate
small = PeerFabric(ops, 1000, 1000)
overflowed = dist_join_peer(ops, small, R, S, None, r, s) is None
small.close()
fabric.close()
if rank == 0:
    print("RESULT " + json.dumps(out))
    print("PEER " + json.dumps(peer))
    print("GRAPH " + json.dumps(graphed))
    print("OVERFLOW " + json.dumps(overflowed))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_gpu_join_equals_single_gpu(Hgpu, tmp_path):
    _two_gpu_join_equals_single_gpu(Hgpu, tmp_path, {})


@pytest.mark.skipif(os.environ.get("HWBRJ_TEST_EXPERIMENTAL") != "1", reason="experimental: set HWBRJ_TEST_EXPERIMENTAL=1")
def test_two_gpu_join_with_precounted_routing(Hgpu, tmp_path):
    """HWBRJ_ROUTE_PRECOUNT=1: one remote claim per owner (k_route_claim) instead of one per (tile, owner)"""
    _two_gpu_join_equals_single_gpu(Hgpu, tmp_path, {"HWBRJ_ROUTE_PRECOUNT": "1"})


@pytest.mark.skipif(os.environ.get("HWBRJ_TEST_EXPERIMENTAL") != "1", reason="experimental: set HWBRJ_TEST_EXPERIMENTAL=1")
def test_two_gpu_join_with_r_partitioned_ahead(Hgpu, tmp_path):
    """HWBRJ_DIST_OVERLAP_R=1: the owned R is partitioned on a side stream under the filter exchange and the S probe
    (hwbrj_join_prepare_r), eager and captured in the CUDA graph"""
    _two_gpu_join_equals_single_gpu(Hgpu, tmp_path, {"HWBRJ_DIST_OVERLAP_R": "1"})


@pytest.mark.skipif(os.environ.get("HWBRJ_TEST_EXPERIMENTAL") != "1", reason="experimental: set HWBRJ_TEST_EXPERIMENTAL=1")
def test_two_gpu_join_with_both_experimental_paths(Hgpu, tmp_path):
    _two_gpu_join_equals_single_gpu(Hgpu, tmp_path, {"HWBRJ_DIST_OVERLAP_R": "1", "HWBRJ_ROUTE_PRECOUNT": "1"})


def _two_gpu_join_equals_single_gpu(Hgpu, tmp_path, extra_env):
    if Hgpu.device_count() < 2:
        pytest.skip("needs 2 GPUs (one rank per GPU)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, HWBRJ_ROOT=ROOT, **extra_env)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][0]
    got = json.loads(line[7:])
    peer = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("PEER ")][0][5:])
    graphed = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("GRAPH ")][0][6:])
    assert graphed == peer
    assert json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("OVERFLOW ")][0][9:]) is True
    r, s, q = 2_000_000, 16_000_000, 0.01
    dR = Hgpu.DeviceRelation.generate(0, r, r, 1.0, 1)
    dS = Hgpu.DeviceRelation.generate(1, s, r, q, 2)
    for case, g in zip([(0, 1 << 24, 1, 512), (0, 1 << 24, 3, 512), (1, 1 << 24, 4, 256), None], got):
        one = Hgpu.join_device(dR, dS, Hgpu.BloomFilterArgs(*case) if case else None)
        assert (g["matches"], g["filtered"], g["checksum_pair"], g["checksum_key"]) == \
               (one.totalresults, one.filtered, one.checksum_pair, one.checksum_key), case
        if case:
            assert g["tuples_over_nvlink_s"] <= one.filtered
    for case, g in zip([(0, 1 << 24, 1, 512), (0, 1 << 24, 3, 512), (1, 1 << 24, 4, 256), None], peer):
        one = Hgpu.join_device(dR, dS, Hgpu.BloomFilterArgs(*case) if case else None)
        assert (g["matches"], g["filtered"], g["checksum_pair"], g["checksum_key"]) == \
               (one.totalresults, one.filtered, one.checksum_pair, one.checksum_key), ("peer path", case)
        assert g["r_owned_total"] == r and g["s_owned_total"] == (one.filtered if case else s)
