"""TEST-ONLY local-compute backend for hwbloomradixjoin_b200.dist: the oracle on CPU tensors, so that the sharding
and exchange logic (owner function, counts, all-to-all, filter combine, all-reduce) runs under gloo without a GPU.
The product never imports this file; its only backend is dist.CudaOps."""
import numpy as np
import torch

import oracle


def _np(t: torch.Tensor) -> np.ndarray:
    return t.numpy().view(oracle.TUPLE)


def _t(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy())


class OracleOps:
    def empty_tuples(self, n):
        return torch.empty(n, dtype=torch.int64)

    def _owner(self, keys: np.ndarray, world: int, slice_args):
        g = world.bit_length() - 1
        if slice_args is not None and slice_args.variant == 1:
            nblocks = slice_args.m // slice_args.B
            return (oracle.hash_many(0, 42, keys) & np.uint32(nblocks - 1)) >> np.uint32(nblocks.bit_length() - 1 - g)
        if slice_args is not None:
            return (oracle.hash_many(2, 42, keys) & np.uint32(slice_args.m - 1)) >> np.uint32(slice_args.m.bit_length() - 1 - g)
        return oracle.hash_many(2, 42, keys) >> np.uint32(32 - g) if g else np.zeros(keys.shape[0], np.uint32)

    def owner_partition(self, rel, world, slice_args):
        a = _np(rel)
        own = self._owner(a["key"], world, slice_args).astype(np.int64)
        order = np.argsort(own, kind="stable")
        counts = np.bincount(own, minlength=world).tolist()
        return _t(a[order]), counts

    def filter_build(self, rel, bloom):
        bm = oracle.bloom_build(_np(rel), bloom.variant, bloom.m, bloom.k, bloom.B)
        return torch.from_numpy(bm.copy())

    def filter_or(self, dst, src):
        dst |= src

    def filter_probe(self, filt, rel, bloom):
        n, surv = oracle.bloom_filter(filt.numpy(), _np(rel), bloom.variant, bloom.m, bloom.k, bloom.B, want_survivors=True)
        return _t(surv)

    def join(self, R, S):
        return oracle.join(_np(R), _np(S), False)
