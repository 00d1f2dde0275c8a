"""FAISS-shaped flat inner-product index over the catalog-scoring kernels (K4, K5).

Mirrors the faiss-cpu surface the reference touches, with the same names and return conventions:
  faiss.IndexFlatIP(d), .add(x), .search(x, k) -> (D, I), .ntotal, .d     src/trainer.py:241-242,
                                                                           app/recommendation_service.py:47-50,72
  faiss.normalize_L2(x)  (in place, returns None)                          src/trainer.py:240, app/...:71
  faiss.write_index(index, path) / faiss.read_index(path)                  src/trainer.py:243, app/...:47

numpy in -> numpy out, exactly like faiss (the C library stages host buffers through the GPU inside
the call); torch CUDA tensors in -> torch CUDA tensors out with no host round trip.  The catalog lives
in HBM behind an opaque C handle.  Results: per query the k largest inner products, score descending
then row ascending, padded with (-FLT_MAX, -1) when the index holds fewer than k rows.
"""

import ctypes
import struct
from typing import Tuple, Union

import numpy as np
import torch

from . import _lib

Array = Union[np.ndarray, torch.Tensor]


def _check_2d_f32(x, name, d=None):
    if isinstance(x, np.ndarray):
        if x.dtype != np.float32 or x.ndim != 2 or not x.flags.c_contiguous:
            raise TypeError(f"{name} must be a C-contiguous float32 [n, d] array")
    elif isinstance(x, torch.Tensor):
        if x.dtype != torch.float32 or x.dim() != 2:
            raise TypeError(f"{name} must be a float32 [n, d] tensor")
    else:
        raise TypeError(f"{name} must be a numpy array or a torch tensor")
    if d is not None and x.shape[1] != d:
        raise ValueError(f"{name} has dimension {x.shape[1]}, index has {d}")


def normalize_L2(x: Array) -> None:
    """In place x_i /= ||x_i||_2 (zero rows untouched); returns None like faiss.normalize_L2."""
    _check_2d_f32(x, "x")
    if isinstance(x, np.ndarray) and not x.flags.writeable:
        raise ValueError("x must be writeable")
    b = _lib.dl(x)
    _lib.check(_lib.load().rs_normalize_l2(b.ptr, _lib.stream_ptr()))


class IndexFlatIP:
    def __init__(self, d: int):
        self._h = ctypes.c_void_p()
        self._lib = _lib.load()
        _lib.check(self._lib.rs_index_create(int(d), ctypes.byref(self._h)))
        self.d = int(d)
        self.is_trained = True
        self.metric_type = 0  # faiss.METRIC_INNER_PRODUCT

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.rs_index_free(h)
            except Exception:
                pass
            self._h = ctypes.c_void_p()

    @property
    def ntotal(self) -> int:
        out = ctypes.c_int64(0)
        _lib.check(self._lib.rs_index_ntotal(self._h, ctypes.byref(out)))
        return int(out.value)

    def add(self, x: Array) -> None:
        _check_2d_f32(x, "x", self.d)
        b = _lib.dl(x)
        _lib.check(self._lib.rs_index_add(self._h, b.ptr, _lib.stream_ptr()))

    def search(self, x: Array, k: int, *, mode: str = "auto", id_base: int = 0) -> Tuple[Array, Array]:
        """-> (D float32 [n, k], I int64 [n, k]); `mode` in {"auto", "exact", "tensor"} (same results)."""
        _check_2d_f32(x, "x", self.d)
        k = int(k)
        if k < 1:
            raise ValueError("k must be >= 1")
        n = x.shape[0]
        if isinstance(x, np.ndarray):
            D = np.empty((n, k), dtype=np.float32)
            I = np.empty((n, k), dtype=np.int64)
        else:
            if not x.is_cuda:
                raise TypeError("torch queries must be CUDA tensors (use numpy for host data)")
            x = x.contiguous()
            D = torch.empty((n, k), dtype=torch.float32, device=x.device)
            I = torch.empty((n, k), dtype=torch.int64, device=x.device)
        if n == 0:
            return D, I
        xb, Db, Ib = _lib.dl(x), _lib.dl(D), _lib.dl(I)
        _lib.check(self._lib.rs_index_search(self._h, xb.ptr, k, int(id_base), _lib.mode_code(mode), Db.ptr, Ib.ptr,
                                             _lib.stream_ptr()))
        return D, I

    def search_deferred(self, x: "torch.Tensor", k: int, *, mode: str = "auto", id_base: int = 0, out=None):
        """search() without the host synchronisation: -> (D, I, need_exact int32 [n]).  Rows with need_exact == 1 are not
        final; the caller re-searches those queries with mode="exact" (sharded.ShardedIndexFlatIP does, after its
        exchange).  `out` = (D, I, need_exact) writes into caller-provided CUDA tensors (views of an exchange buffer)."""
        _check_2d_f32(x, "x", self.d)
        k = int(k)
        if k < 1:
            raise ValueError("k must be >= 1")
        if isinstance(x, np.ndarray) or not x.is_cuda:
            raise TypeError("search_deferred takes CUDA tensors")
        x = x.contiguous()
        n = x.shape[0]
        if out is None:
            out = (torch.empty((n, k), dtype=torch.float32, device=x.device),
                   torch.empty((n, k), dtype=torch.int64, device=x.device),
                   torch.zeros((n,), dtype=torch.int32, device=x.device))
        D, I, F = out
        if n == 0:
            return D, I, F
        xb, Db, Ib, Fb = _lib.dl(x), _lib.dl(D), _lib.dl(I), _lib.dl(F)
        _lib.check(self._lib.rs_index_search_deferred(self._h, xb.ptr, k, int(id_base), _lib.mode_code(mode), Db.ptr, Ib.ptr,
                                                      Fb.ptr, _lib.stream_ptr()))
        return D, I, F

    # ---- shared-threshold search of one shard of a partitioned catalog (include/recsys_b200.h) ----
    @staticmethod
    def tensor_plan(d: int, ntotal: int, nq: int, k: int) -> Tuple[bool, int]:
        """(tensor path applies to an index of `ntotal` rows for this (nq, k), rows its threshold sample covers)."""
        rows, ok = ctypes.c_int64(0), ctypes.c_int32(0)
        _lib.check(_lib.load().rs_tensor_plan(int(d), int(ntotal), int(nq), int(k), ctypes.byref(rows), ctypes.byref(ok)))
        return bool(ok.value), int(rows.value)

    @staticmethod
    def threshold_rank(k: int, rho: float) -> int:
        r = ctypes.c_int32(0)
        _lib.check(_lib.load().rs_tensor_threshold_rank(int(k), float(rho), ctypes.byref(r)))
        return int(r.value)

    def shared_begin(self, x: "torch.Tensor", k: int, top: "torch.Tensor"):
        """Sample pass; fills top [n, R] with each query's R largest sample maxima.  -> session for shared_finish."""
        _check_2d_f32(x, "x", self.d)
        sess = ctypes.c_void_p()
        xb, tb = _lib.dl(x), _lib.dl(top)
        _lib.check(self._lib.rs_index_shared_begin(self._h, xb.ptr, int(k), tb.ptr, ctypes.byref(sess), _lib.stream_ptr()))
        return sess

    def shared_finish(self, sess, tops: "torch.Tensor", rank: int, *, id_base: int = 0, out) -> None:
        """tops [shards, n, R] gathered from every shard; out = (D, I, need_exact, bound) CUDA tensors.  Ends the session."""
        D, I, F, B = out
        tb, Db, Ib, Fb, Bb = _lib.dl(tops), _lib.dl(D), _lib.dl(I), _lib.dl(F), _lib.dl(B)
        _lib.check(self._lib.rs_index_shared_finish(sess, tb.ptr, int(rank), int(id_base), Db.ptr, Ib.ptr, Fb.ptr, Bb.ptr,
                                                    _lib.stream_ptr()))

    def fallbacks(self):
        """(last, total) tensor-mode queries that were re-run on the exact scan (same results either way)."""
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(self._lib.rs_index_fallbacks(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def debug_tensor_scores(self, q: torch.Tensor) -> torch.Tensor:
        """Raw bf16 tensor-core scores [nq, ntotal] (validation of the tcgen05 tiles; small catalogs)."""
        out = torch.empty((q.shape[0], self.ntotal), dtype=torch.float32, device=q.device)
        qb, ob = _lib.dl(q.contiguous()), _lib.dl(out)
        _lib.check(self._lib.rs_debug_tensor_scores(self._h, qb.ptr, ob.ptr, _lib.stream_ptr()))
        return out

    def reconstruct_n(self, i0: int = 0, n: int = None) -> np.ndarray:
        """Stored vectors [i0, i0+n) back on the host (faiss.IndexFlat.reconstruct_n)."""
        nt = self.ntotal
        n = nt - i0 if n is None else n
        if i0 < 0 or n < 0 or i0 + n > nt:
            raise ValueError("range out of bounds")
        out = np.empty((n, self.d), dtype=np.float32)
        if n:
            b = _lib.dl(out)
            _lib.check(self._lib.rs_index_reconstruct(self._h, int(i0), b.ptr, _lib.stream_ptr()))
        return out


# ---- faiss.write_index / read_index for IndexFlatIP (section 8f: artifact formats) ------------------
# Binary layout of faiss's index_write.cpp for a flat index, restated from the faiss sources'
# documentation (faiss is not installable here; the layout is unverified against a real file):
#   fourcc "IxFI" | int32 d | int64 ntotal | int64 dummy (1<<20) x2 | uint8 is_trained | int32 metric_type
#   | uint64 count (= ntotal * d floats) | float32 data
_FOURCC_IP = b"IxFI"


def write_index(index: IndexFlatIP, path: str) -> None:
    xb = index.reconstruct_n(0, index.ntotal)
    with open(path, "wb") as f:
        f.write(_FOURCC_IP)
        f.write(struct.pack("<i", index.d))
        f.write(struct.pack("<q", index.ntotal))
        f.write(struct.pack("<qq", 1 << 20, 1 << 20))
        f.write(struct.pack("<B", 1))
        f.write(struct.pack("<i", 0))
        f.write(struct.pack("<Q", xb.size))
        f.write(np.ascontiguousarray(xb, dtype="<f4").tobytes())


def read_index(path: str) -> IndexFlatIP:
    with open(path, "rb") as f:
        fourcc = f.read(4)
        if fourcc != _FOURCC_IP:
            raise ValueError(f"unsupported index type {fourcc!r}: only IndexFlatIP ('IxFI') is on the serving path")
        (d,) = struct.unpack("<i", f.read(4))
        (ntotal,) = struct.unpack("<q", f.read(8))
        f.read(16)
        f.read(1)
        (metric,) = struct.unpack("<i", f.read(4))
        if metric != 0:
            raise ValueError("IxFI file with a non inner-product metric")
        (count,) = struct.unpack("<Q", f.read(8))
        if count != ntotal * d:
            raise ValueError("corrupt index file: vector count does not match the header")
        xb = np.frombuffer(f.read(count * 4), dtype="<f4").reshape(ntotal, d)
    index = IndexFlatIP(d)
    if ntotal:
        index.add(np.ascontiguousarray(xb, dtype=np.float32))
    return index
