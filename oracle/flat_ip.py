"""Retrieval half of the oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates the faiss-cpu calls the reference makes:
  faiss.normalize_L2(x)              src/trainer.py:240, app/recommendation_service.py:71
  faiss.IndexFlatIP(d) / .add / .ntotal / .search
                                     src/trainer.py:241-242, app/recommendation_service.py:50,72
  tf.linalg.matvec(items, user)      app/recommendation_service.py:91

Two scorers are kept apart on purpose:
  * search_chain  -- the PARITY oracle: the stated fp32 FMA chain (oracle/chain.c),
    tie-break (score desc, index asc), -1 / -FLT_MAX padding.
  * search_blas   -- the CPU BASELINE: what faiss-cpu does above 20 queries
    (blocked sgemm on the host BLAS + a per-query heap), restated from its
    documentation.  Same top-k sets up to fp32 reassociation near ties; timed, not
    used as the checker.
"""

import ctypes

import numpy as np

from .build import load_chain_lib

_F32P = ctypes.POINTER(ctypes.c_float)
_I64P = ctypes.POINTER(ctypes.c_int64)
_I32P = ctypes.POINTER(ctypes.c_int)
NEG_PAD = np.float32(-3.4028234663852886e38)  # -FLT_MAX, faiss's IP heap neutral


def _f32(a):
    return a.ctypes.data_as(_F32P)


def _check_f32_2d(x, name):
    if not isinstance(x, np.ndarray) or x.dtype != np.float32 or x.ndim != 2 or not x.flags.c_contiguous:
        raise TypeError(f"{name} must be a C-contiguous float32 [n, d] array")


def normalize_L2(x: np.ndarray) -> None:
    """In place, returns None -- as faiss.normalize_L2.  All-zero rows are skipped."""
    _check_f32_2d(x, "x")
    load_chain_lib().oracle_normalize_l2(_f32(x), x.shape[0], x.shape[1])


def scores_chain(xb: np.ndarray, q: np.ndarray) -> np.ndarray:
    _check_f32_2d(xb, "xb")
    _check_f32_2d(q, "q")
    out = np.empty((q.shape[0], xb.shape[0]), dtype=np.float32)
    load_chain_lib().oracle_scores_chain(_f32(xb), xb.shape[0], xb.shape[1], _f32(q), q.shape[0], _f32(out))
    return out


def search_chain(xb: np.ndarray, q: np.ndarray, k: int, id_base: int = 0):
    """IndexFlatIP.search restated with the stated chain arithmetic -> (D f32[nq,k], I i64[nq,k])."""
    _check_f32_2d(xb, "xb")
    _check_f32_2d(q, "q")
    if q.shape[1] != xb.shape[1]:
        raise ValueError("dimension mismatch")
    nq = q.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    if xb.shape[0] == 0:
        D[:] = NEG_PAD
        I[:] = -1
        return D, I
    load_chain_lib().oracle_search_chain(_f32(xb), xb.shape[0], xb.shape[1], _f32(q), nq, int(k),
                                         int(id_base), _f32(D), I.ctypes.data_as(_I64P))
    return D, I


def search_blas(xb: np.ndarray, q: np.ndarray, k: int, bs_x: int = 4096, bs_y: int = 16384):
    """CPU baseline: faiss's exhaustive_inner_product_blas shape -- query blocks of
    bs_x, database blocks of bs_y, sgemm per block, heap update per block.  faiss's own
    database block is 1024 rows; 16384 is used here because it is the fastest setting for
    this numpy/OpenMP orchestration (8x faster than 1024 on the 8-core host), i.e. the
    baseline is tuned in the CPU's favour."""
    _check_f32_2d(xb, "xb")
    _check_f32_2d(q, "q")
    lib = load_chain_lib()
    nq, n = q.shape[0], xb.shape[0]
    hs = lib.oracle_hit_size()
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    for i0 in range(0, nq, bs_x):
        i1 = min(nq, i0 + bs_x)
        heaps = np.zeros((i1 - i0) * k * hs, dtype=np.uint8)
        cnts = np.zeros(i1 - i0, dtype=np.int32)
        qb = q[i0:i1]
        for j0 in range(0, n, bs_y):
            j1 = min(n, j0 + bs_y)
            S = np.ascontiguousarray(qb @ xb[j0:j1].T)
            lib.oracle_heap_addn(heaps.ctypes.data, cnts.ctypes.data_as(_I32P), i1 - i0, int(k), _f32(S),
                                 j1 - j0, j0)
        lib.oracle_heap_finish(heaps.ctypes.data, cnts.ctypes.data_as(_I32P), i1 - i0, int(k),
                               _f32(D[i0:i1]), I[i0:i1].ctypes.data_as(_I64P))
    return D, I


def search_fp64(xb: np.ndarray, q: np.ndarray, k: int):
    """fp64 scores, same ordering rule; used to classify near-ties in parity reports."""
    S = q.astype(np.float64) @ xb.astype(np.float64).T
    n = xb.shape[0]
    idx = np.broadcast_to(np.arange(n, dtype=np.int64), S.shape)
    order = np.lexsort((idx, -S), axis=1)[:, :k]
    D = np.take_along_axis(S, order, axis=1)
    if n < k:
        pad = k - n
        D = np.concatenate([D, np.full((q.shape[0], pad), float(NEG_PAD))], axis=1)
        order = np.concatenate([order, np.full((q.shape[0], pad), -1, dtype=np.int64)], axis=1)
    return D, order.astype(np.int64)


def merge_topk(Dparts: np.ndarray, Iparts: np.ndarray):
    """Merge [nparts, nq, k] per-shard lists (global ids) into [nq, k]."""
    Dparts = np.ascontiguousarray(Dparts, dtype=np.float32)
    Iparts = np.ascontiguousarray(Iparts, dtype=np.int64)
    nparts, nq, k = Dparts.shape
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    load_chain_lib().oracle_merge_topk(_f32(Dparts), Iparts.ctypes.data_as(_I64P), nparts, nq, k, _f32(D),
                                       I.ctypes.data_as(_I64P))
    return D, I


def matvec(items: np.ndarray, user: np.ndarray) -> np.ndarray:
    """tf.linalg.matvec(item_embeddings[n,d], user_embedding[d]) -- raw dots, chain arithmetic."""
    items = np.ascontiguousarray(items, dtype=np.float32)
    user = np.ascontiguousarray(user, dtype=np.float32).reshape(1, -1)
    return scores_chain(items, user)[0]


class IndexFlatIP:
    """faiss.IndexFlatIP restated: flat fp32 store, row id = insertion order."""

    def __init__(self, d: int):
        self.d = int(d)
        self._xb = np.empty((0, self.d), dtype=np.float32)
        self.is_trained = True

    @property
    def ntotal(self) -> int:
        return self._xb.shape[0]

    def add(self, x: np.ndarray) -> None:
        _check_f32_2d(x, "x")
        if x.shape[1] != self.d:
            raise ValueError("dimension mismatch")
        self._xb = np.ascontiguousarray(np.concatenate([self._xb, x], axis=0))

    def search(self, x: np.ndarray, k: int):
        _check_f32_2d(x, "x")
        if x.shape[1] != self.d:
            raise ValueError("dimension mismatch")
        return search_chain(self._xb, x, k)

    def reconstruct_n(self, i0: int, n: int) -> np.ndarray:
        return self._xb[i0:i0 + n].copy()
