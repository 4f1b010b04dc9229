"""Multi-GPU tests of the sharded join (-m gpu; skipped on a single-GPU box). Both ways of forming a GPU group must
reproduce the single-GPU scalars bit for bit:
  * one process per GPU (torchrun, two ranks; handles over torch.distributed): the in-library peer-memory join
    (hwbrj_dist_join), the same captured as a CUDA graph, and the NCCL reference path;
  * one process driving several GPUs through the reference's own entry points (hwbrj_set_gpus + BPRO/PRO).
One rank per GPU -- ranks are never stacked on one device."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [(0, 1 << 24, 1, 512), (0, 1 << 24, 3, 512), (1, 1 << 24, 4, 256), (0, 1 << 24, 0, 512), None]

WORKER = r'''
import json, os, sys, hashlib, torch, torch.distributed as dist
sys.path.insert(0, os.environ["HWBRJ_ROOT"])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from hwbloomradixjoin_b200 import BloomFilterArgs
from hwbloomradixjoin_b200.dist import CudaOps, DistGroup, DistJoinGraph, dist_join
ops = CudaOps(dev)
CASES = json.loads(os.environ["HWBRJ_CASES"])
r, s, q = 2_000_000, 16_000_000, 0.01
per_r, per_s = r // world, s // world
R = ops.generate_shard(0, r, r, 1.0, 1, rank * per_r, r - rank * per_r if rank == world - 1 else per_r)
S = ops.generate_shard(1, s, r, q, 2, rank * per_s, s - rank * per_s if rank == world - 1 else per_s)
Z = ops.generate_shard(2, s // 4, r, 1.0, 3, rank * (per_s // 4), per_s // 4)  # Zipf probe chunk: skewed owners
KEYS = ("matches", "filtered", "checksum_pair", "checksum_key")
out, peer, graphed, owned, filt = [], [], [], [], []
grp = DistGroup(ops, int(r / world * 1.25) + 65536, s + 2, 1 << 21)
for case in CASES:
    bloom = BloomFilterArgs(*case) if case else None
    res = dist_join(ops, R, S, bloom)
    out.append({k: res[k] for k in KEYS + ("tuples_over_nvlink_s",)})
    for rep in range(2):  # twice: nothing of the first join may leak into the second
        pres = grp.join(R, S, bloom, r)
    peer.append({k: pres[k] for k in KEYS})
    t = torch.tensor([pres["owned_r"], pres["owned_s"]], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    owned.append(t.tolist())
    if bloom is not None:  # the replicated filter is byte-identical on every rank
        dig = hashlib.sha256(grp.filter_bytes(bloom.m // 8)).hexdigest()
        alld = [None] * world
        dist.all_gather_object(alld, dig)
        filt.append(alld)
    pg = DistJoinGraph(grp, R, S, bloom, r)  # the same pipeline captured as one CUDA graph
    for rep in range(3):
        gres = pg.replay()
    graphed.append({k: gres[k] for k in KEYS})
    del pg
zres = grp.join(R, Z, BloomFilterArgs(0, 1 << 24, 1, 512), r)
zipf = {k: zres[k] for k in KEYS}
# a receive buffer that is too small must be reported (None), never overrun
small = DistGroup(ops, 1000, 1000, 1 << 21)
overflowed = small.join(R, S, None, r) is None
again = small.join(R, S, None, r) is None  # and the group stays usable (and keeps failing) afterwards
small.close()
grp.close()
if rank == 0:
    print("RESULT " + json.dumps({"nccl": out, "peer": peer, "graph": graphed, "owned": owned, "filters": filt,
                                  "zipf": zipf, "overflow": [overflowed, again]}))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("dist_parts", ["0", "4"])
def test_two_gpu_join_equals_single_gpu(Hgpu, tmp_path, dist_parts):
    """dist_parts: groups of owned level-1 bins (HWBRJ_DIST_PARTS). "0" = chosen by size (one group at these sizes),
    "4" = four groups pipelined on two streams, as the large workloads run on 2 GPUs"""
    if Hgpu.device_count() < 2:
        pytest.skip("needs 2 GPUs (one rank per GPU)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, HWBRJ_ROOT=ROOT, HWBRJ_CASES=json.dumps(CASES), HWBRJ_DIST_PARTS=dist_parts)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    got = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][0][7:])
    assert got["graph"] == got["peer"]
    assert got["overflow"] == [True, True]
    for digests in got["filters"]:
        assert len(set(digests)) == 1
    r, s, q = 2_000_000, 16_000_000, 0.01
    dR = Hgpu.DeviceRelation.generate(0, r, r, 1.0, 1)
    dS = Hgpu.DeviceRelation.generate(1, s, r, q, 2)
    for case, g, pr, own in zip(CASES, got["nccl"], got["peer"], got["owned"]):
        one = Hgpu.join_device(dR, dS, Hgpu.BloomFilterArgs(*case) if case else None)
        want = (one.totalresults, one.filtered, one.checksum_pair, one.checksum_key)
        assert (g["matches"], g["filtered"], g["checksum_pair"], g["checksum_key"]) == want, ("nccl path", case)
        assert (pr["matches"], pr["filtered"], pr["checksum_pair"], pr["checksum_key"]) == want, ("peer path", case)
        if case:
            assert g["tuples_over_nvlink_s"] <= one.filtered
        assert own == [r, one.filtered if case else s], case  # every tuple has exactly one owner
    # Zipf chunks: rank i generated positions [i*per/4, (i+1)*per/4) of the same global relation of s/4 tuples
    per = s // 2 // 4
    from hwbloomradixjoin_b200 import _native
    parts = [Hgpu.DeviceRelation(_native.load().hwbrj_rel_generate_shard(2, s // 4, r, 1.0, 3, i * per, per)) for i in range(2)]
    Z = np.concatenate([p_.download() for p_ in parts])
    one = Hgpu.join_device(dR, Hgpu.DeviceRelation.upload(Z), Hgpu.BloomFilterArgs(0, 1 << 24, 1, 512))
    z = got["zipf"]
    assert (z["matches"], z["filtered"], z["checksum_pair"], z["checksum_key"]) == \
           (one.totalresults, one.filtered, one.checksum_pair, one.checksum_key)
    assert z["matches"] == Z.shape[0]


@pytest.mark.parametrize("gpus", [2, 4, 8])
def test_reference_entry_points_on_several_gpus(Hgpu, oracle_mod, gpus):
    """hwbrj_set_gpus(n): BPRO / PRO shard the host relations over n GPUs of this process (peer access instead of IPC)
    and must return exactly what one GPU and the oracle return"""
    if Hgpu.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    code = r'''
import json, sys
sys.path.insert(0, %(root)r)
import oracle
import hwbloomradixjoin_b200 as H
H.set_quiet(True)
R = oracle.gen_R(300_001); S = oracle.gen_S(2_500_003, 300_001, 0.02)
out = []
for gpus in (1, %(gpus)d, %(gpus)d):
    H.set_gpus(gpus)
    row = []
    for case in [(0, 1 << 22, 1, 512), (1, 1 << 22, 4, 256), (0, 1 << 22, 3, 512), None]:
        res = H.BPRO(R, S, 4, H.BloomFilterArgs(*case)) if case else H.PRO(R, S, 4)
        row.append([res.totalresults, res.filtered, res.checksum_pair, res.checksum_key, res.stats["n_gpus"]])
    out.append(row)
exp = []
for case in [(0, 1 << 22, 1, 512), (1, 1 << 22, 4, 256), (0, 1 << 22, 3, 512), None]:
    o = oracle.join(R, S, True, *case) if case else oracle.join(R, S, False)
    exp.append([o["matches"], o["filtered"] if case else -1, o["checksum_pair"], o["checksum_key"]])
print(json.dumps({"out": out, "exp": exp}))
''' % {"root": ROOT, "gpus": gpus}
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    for row, n in zip(d["out"], (1, gpus, gpus)):
        assert [x[:4] for x in row] == d["exp"]
        assert all(x[4] == n for x in row)
