"""Operator-level Python entry points: one function per C-ABI compute call.

All tensors are torch CUDA tensors (torch only supplies device memory and the stream); every function
launches hand-written sm_100a kernels from librecsys_b200.so on torch's current stream and raises on
any failure -- there is no eager/PyTorch fallback.
"""

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise TypeError("expected CUDA torch tensors (no CPU path exists)")


def pack_strings(ids, width: Optional[int] = None):
    """Strings -> the fixed-width byte matrix the device lookup reads: uint8 [n, W] numpy, one NUL-padded row per string
    (UTF-8), W a multiple of 16.  Accepts a sequence of str / bytes or a numpy `S` / `U` array.  Pure host code."""
    import numpy as np
    if isinstance(ids, np.ndarray) and ids.dtype.kind == "S":
        a = ids.reshape(-1)
    elif isinstance(ids, np.ndarray) and ids.dtype.kind == "U":
        a = np.char.encode(ids.reshape(-1), "utf-8")
    else:
        a = np.array([s if isinstance(s, bytes) else str(s).encode("utf-8") for s in ids], dtype=np.bytes_)
        if a.size == 0:
            a = np.zeros((0,), dtype="S1")
    w = max(int(a.dtype.itemsize), 1)
    want = int(width) if width is not None else (w + 15) // 16 * 16
    if want < w:
        raise ValueError(f"width {want} is smaller than the longest string ({w} bytes)")
    out = np.zeros((a.shape[0], want), dtype=np.uint8)
    if a.shape[0]:
        out[:, :w] = np.ascontiguousarray(a).view(np.uint8).reshape(a.shape[0], w)
    return out


def pack_str_list(ids):
    """Fast packing of a list / tuple of Python `str` through an Arrow string array (one C-level pass instead of a Python
    loop): the same uint8 [n, W] matrix pack_strings gives, or None when that is not possible -- pyarrow missing, entries
    that are not `str`, a string longer than 4,096 bytes, or a string with a NUL byte inside (the byte layout ends a row at its first NUL, a dict key does
    not: such ids must go through the host dict to keep the two lookups identical)."""
    import numpy as np
    try:
        import pyarrow as pa
        a = pa.array(ids, type=pa.string())
    except Exception:  # noqa: BLE001 -- ImportError, ArrowInvalid / ArrowTypeError for non-str entries
        return None
    if a.null_count:
        return None
    n = len(a)
    bufs = a.buffers()
    off = np.frombuffer(bufs[1], dtype=np.int32, count=n + 1, offset=a.offset * 4).astype(np.int64)
    lens = np.diff(off)
    w = int(lens.max()) if n else 1
    if w > 4096:                                   # wider than the device lookup accepts (rs_vocab_lookup): host dict
        return None
    out = np.zeros((n, max(16, (w + 15) // 16 * 16)), dtype=np.uint8)
    if n and off[-1] > off[0]:
        data = np.frombuffer(bufs[2], dtype=np.uint8)
        rows = np.repeat(np.arange(n), lens)
        cols = np.arange(off[0], off[-1]) - np.repeat(off[:-1], lens)
        out[rows, cols] = data[off[0]:off[-1]]
        if not np.array_equal((out != 0).sum(axis=1), lens):
            return None
    return out


class DeviceVocabulary:
    """K0 handle: the StringLookup table on the device (csrc/string_lookup.cu).  `lookup(ids)` takes a CUDA uint8 [B, W]
    byte matrix (see pack_strings) and returns int64 [B] rows: vocabulary[i] -> i + 1, anything else -> 0."""

    def __init__(self, vocabulary, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("recsys_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        words = pack_str_list(vocabulary) if isinstance(vocabulary, (list, tuple)) else None
        if words is None:
            words = pack_strings(vocabulary)
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rs_vocab_create(words.ctypes.data_as(ctypes.c_void_p), int(words.shape[0]),
                                                 int(words.shape[1]), ctypes.byref(self._h)))

    def __len__(self) -> int:
        n = ctypes.c_int64()
        _lib.check(self._lib.rs_vocab_size(self._h, ctypes.byref(n)))
        return int(n.value)

    def lookup(self, ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _require_cuda(ids, out)
        if ids.dtype != torch.uint8 or ids.dim() != 2:
            raise TypeError("ids must be a uint8 [B, W] byte matrix (ops.pack_strings)")
        if out is None:
            out = torch.empty((ids.shape[0],), dtype=torch.int64, device=ids.device)
        ib, ob = _lib.dl(ids), _lib.dl(out)
        _lib.check(self._lib.rs_vocab_lookup(self._h, ib.ptr, ob.ptr, _lib.stream_ptr()))
        return out

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.rs_vocab_free(h)
            except Exception:  # noqa: BLE001 -- interpreter shutdown
                pass


def gather_concat(tables: Sequence[torch.Tensor], ids: Sequence[torch.Tensor],
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1: out[b] = concat_t tables[t][ids[t][b]]  (Embedding lookups + tf.concat, src/models.py:85,89,128)."""
    _require_cuda(*tables, *ids, out)
    lib = _lib.load()
    n = len(tables)
    if out is None:
        out = torch.empty((ids[0].shape[0], sum(int(t.shape[1]) for t in tables)), dtype=torch.float32,
                          device=tables[0].device)
    tb = [_lib.dl(t) for t in tables]
    ib = [_lib.dl(i) for i in ids]
    tarr = (_lib._DLP * n)(*[b.ptr for b in tb])
    iarr = (_lib._DLP * n)(*[b.ptr for b in ib])
    ob = _lib.dl(out)
    _lib.check(lib.rs_gather_concat(n, tarr, iarr, ob.ptr, _lib.stream_ptr()))
    return out


def dense(x: torch.Tensor, kernel: torch.Tensor, bias: Optional[torch.Tensor], activation=None,
          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K2: act(x @ kernel + bias) with the Keras (in, out) kernel layout (src/models.py:27,76-77)."""
    _require_cuda(x, kernel, bias, out)
    if out is None:
        out = torch.empty((x.shape[0], kernel.shape[1]), dtype=torch.float32, device=x.device)
    xb, kb, bb, ob = _lib.dl(x), _lib.dl(kernel), _lib.dl(bias), _lib.dl(out)
    _lib.check(_lib.load().rs_dense_forward(xb.ptr, kb.ptr, _lib.ptr(bb), _lib.act_code(activation), ob.ptr,
                                            _lib.stream_ptr()))
    return out


def dense_x3_prepare(kernel: torch.Tensor):
    """Keras kernel f32 [in, out] -> (wt_hi, wt_lo) fp16 [out, in]: the split B operand of the fp32 tensor mode."""
    _require_cuda(kernel)
    K, N = int(kernel.shape[0]), int(kernel.shape[1])
    hi = torch.empty((N, K), dtype=torch.float16, device=kernel.device)
    lo = torch.empty((N, K), dtype=torch.float16, device=kernel.device)
    kb, hb, lb = _lib.dl(kernel), _lib.dl(hi), _lib.dl(lo)
    _lib.check(_lib.load().rs_dense_x3_prepare(kb.ptr, hb.ptr, lb.ptr, _lib.stream_ptr()))
    return hi, lo


def dense_x3_supported(in_dim: int, out_dim: int, activation=None) -> bool:
    try:
        act = _lib.act_code(activation)
    except ValueError:
        return False
    return bool(_lib.load().rs_dense_x3_supported(int(in_dim), int(out_dim), act))


def dense_x3_layers(layers, activation=None) -> None:
    """K2 at fp32 accuracy on tcgen05 (3 x fp16 split products, csrc/dense_x3.cu): one layer, or two independent layers of
    the same shape in ONE launch (the two towers).  Each layer is a dict with wt_hi, wt_lo (dense_x3_prepare), optional
    bias, an input -- x (f32 [B, in]; with ids: x is an embedding table and the layer reads x[ids]) or x_hi + x_lo (the
    split planes a previous layer wrote) -- and outputs y (f32) and / or y_hi + y_lo (split planes for the next layer)."""
    keys = ("x", "ids", "x_hi", "x_lo", "wt_hi", "wt_lo", "bias", "y", "y_hi", "y_lo")
    n = len(layers)
    arr = (_lib.X3Layer * n)()
    keep = []
    for i, L in enumerate(layers):
        for k in keys:
            t = L.get(k)
            _require_cuda(t)
            b = _lib.dl(t)
            keep.append(b)
            setattr(arr[i], k, _lib.ptr(b))
    _lib.check(_lib.load().rs_dense_forward_x3(n, arr, _lib.act_code(activation), _lib.stream_ptr()))


def dense_x3(x: torch.Tensor, wt_hi: torch.Tensor, wt_lo: torch.Tensor, bias: Optional[torch.Tensor], activation=None,
             out: Optional[torch.Tensor] = None, ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One fp32-in / fp32-out layer on the fp32 tensor kernel; same contract as dense().  ids: x is an embedding table
    and the layer reads x[ids] (K1 fused into the layer)."""
    if out is None:
        out = torch.empty((x.shape[0] if ids is None else ids.shape[0], wt_hi.shape[0]), dtype=torch.float32, device=x.device)
    dense_x3_layers([{"x": x, "ids": ids, "wt_hi": wt_hi, "wt_lo": wt_lo, "bias": bias, "y": out}], activation)
    return out


def split_planes(rows: int, width: int, device) -> tuple:
    """(hi, lo) fp16 [rows, width] planes: the form activations take between two layers of the fp32 tensor mode."""
    both = torch.empty((2, rows, width), dtype=torch.float16, device=device)
    return both[0], both[1]


def dense_bf16(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor], activation=None,
               out_bf16: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None, want_bf16: bool = True):
    """K2 on tcgen05: x bf16 [B, in], wt bf16 [out, in] (Keras kernel transposed) -> (y_bf16 or None, y_f32 or None)."""
    _require_cuda(x, wt, bias, out_bf16, out_f32)
    if out_bf16 is None and want_bf16:
        out_bf16 = torch.empty((x.shape[0], wt.shape[0]), dtype=torch.bfloat16, device=x.device)
    xb, wb, bb, ob, of = _lib.dl(x), _lib.dl(wt), _lib.dl(bias), _lib.dl(out_bf16), _lib.dl(out_f32)
    _lib.check(_lib.load().rs_dense_forward_bf16(xb.ptr, wb.ptr, _lib.ptr(bb), _lib.act_code(activation), _lib.ptr(ob),
                                                 _lib.ptr(of), _lib.stream_ptr()))
    return out_bf16, out_f32


def chain_forward_bf16(layer_sets, acts, tables=None, ids=None, x=None, y_f32=None, y_bf16=None):
    """Fused Dense chain(s) on tcgen05 (chain_tc.cu).  layer_sets: one list per problem of (wt_bf16 [out,in], bias or
    None); acts: activation names per layer; tower chain: tables/ids per problem; deep chain: x per problem;
    outputs per problem (None entries allowed, at least one of y_f32[p] / y_bf16[p])."""
    n_prob, n_layers = len(layer_sets), len(layer_sets[0])
    keep = []

    def arr(objs):
        if objs is None:
            return None
        bs = [_lib.dl(o) for o in objs]
        keep.append(bs)
        return (_lib._DLP * len(bs))(*[_lib.ptr(b) for b in bs])

    wts = arr([w for ls in layer_sets for (w, _) in ls])
    biases = arr([b for ls in layer_sets for (_, b) in ls])
    act_arr = (ctypes.c_int * n_layers)(*[_lib.act_code(a) for a in acts])
    _lib.check(_lib.load().rs_chain_forward_bf16(n_prob, n_layers, wts, biases, act_arr, arr(tables), arr(ids), arr(x),
                                                 arr(y_f32), arr(y_bf16), _lib.stream_ptr()))


def gather_bf16(table: torch.Tensor, ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1 for the tensor path: fp32 embedding rows -> bf16 activations [B, E]."""
    _require_cuda(table, ids, out)
    if out is None:
        out = torch.empty((ids.shape[0], table.shape[1]), dtype=torch.bfloat16, device=table.device)
    tb, ib, ob = _lib.dl(table), _lib.dl(ids), _lib.dl(out)
    _lib.check(_lib.load().rs_gather_bf16(tb.ptr, ib.ptr, ob.ptr, _lib.stream_ptr()))
    return out


def dcn_cross(x0: torch.Tensor, cross_w: torch.Tensor, cross_b: torch.Tensor,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K3a: the DCN-v1 cross stack (src/models.py:38-44). cross_w, cross_b: [n_layers, D]."""
    _require_cuda(x0, cross_w, cross_b, out)
    if out is None:
        out = torch.empty_like(x0)
    a, w, b, o = _lib.dl(x0), _lib.dl(cross_w), _lib.dl(cross_b), _lib.dl(out)
    _lib.check(_lib.load().rs_dcn_cross_forward(a.ptr, w.ptr, b.ptr, o.ptr, _lib.stream_ptr()))
    return out


def cross_v2_combine(x0: torch.Tensor, u: torch.Tensor, xl: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """DCN-v2 cross layer, elementwise part (EXTENSION, no reference counterpart): out = x0 * u + xl with u = x_l W + b."""
    _require_cuda(x0, u, xl, out)
    if out is None:
        out = torch.empty_like(x0)
    a, b, c, o = _lib.dl(x0), _lib.dl(u), _lib.dl(xl), _lib.dl(out)
    _lib.check(_lib.load().rs_cross_v2_combine(a.ptr, b.ptr, c.ptr, o.ptr, _lib.stream_ptr()))
    return out


def heads(z: torch.Tensor, head_w: torch.Tensor, head_b: torch.Tensor):
    """K3b: rating = z.w_r + b_r, ctr = sigmoid(z.w_c + b_c) (src/models.py:119-120,131). -> ([B,1], [B,1])"""
    _require_cuda(z, head_w, head_b)
    rating = torch.empty((z.shape[0], 1), dtype=torch.float32, device=z.device)
    ctr = torch.empty((z.shape[0], 1), dtype=torch.float32, device=z.device)
    a, w, b, r, c = _lib.dl(z), _lib.dl(head_w), _lib.dl(head_b), _lib.dl(rating), _lib.dl(ctr)
    _lib.check(_lib.load().rs_heads_forward(a.ptr, w.ptr, b.ptr, r.ptr, c.ptr, _lib.stream_ptr()))
    return rating, ctr


def cross_heads(x0: torch.Tensor, deep: torch.Tensor, cross_w: torch.Tensor, cross_b: torch.Tensor, head_w: torch.Tensor,
                head_b: torch.Tensor):
    """K3a + K3b in one pass (no [B, 384] concat): cross stack on x0 [B,256], heads on (cross | deep [B,128]) -> ([B,1], [B,1])."""
    _require_cuda(x0, deep, cross_w, cross_b, head_w, head_b)
    rating = torch.empty((x0.shape[0], 1), dtype=torch.float32, device=x0.device)
    ctr = torch.empty((x0.shape[0], 1), dtype=torch.float32, device=x0.device)
    bs = [_lib.dl(t) for t in (x0, deep, cross_w, cross_b, head_w, head_b, rating, ctr)]
    _lib.check(_lib.load().rs_cross_heads_forward(*[b.ptr for b in bs], _lib.stream_ptr()))
    return rating, ctr


def normalize_l2_(x) -> None:
    """K4: faiss.normalize_L2 in place; x is a CUDA torch tensor or a host numpy array."""
    b = _lib.dl(x)
    _lib.check(_lib.load().rs_normalize_l2(b.ptr, _lib.stream_ptr()))


def matvec(items: torch.Tensor, user: torch.Tensor) -> torch.Tensor:
    """K6: tf.linalg.matvec(items[n,d], user[d]) (app/recommendation_service.py:91)."""
    _require_cuda(items, user)
    out = torch.empty((items.shape[0],), dtype=torch.float32, device=items.device)
    a, u, o = _lib.dl(items), _lib.dl(user), _lib.dl(out)
    _lib.check(_lib.load().rs_matvec(a.ptr, u.ptr, o.ptr, _lib.stream_ptr()))
    return out


def topk_merge(Dparts: torch.Tensor, Iparts: torch.Tensor):
    """K7: merge [nparts, nq, k] per-shard lists (global ids) -> ([nq,k] f32, [nq,k] i64)."""
    _require_cuda(Dparts, Iparts)
    nparts, nq, k = Dparts.shape
    D = torch.empty((nq, k), dtype=torch.float32, device=Dparts.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=Dparts.device)
    a, b, c, d = _lib.dl(Dparts), _lib.dl(Iparts), _lib.dl(D), _lib.dl(I)
    _lib.check(_lib.load().rs_topk_merge(a.ptr, b.ptr, c.ptr, d.ptr, _lib.stream_ptr()))
    return D, I


def topk_verify(D: torch.Tensor, I: torch.Tensor, bounds: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    """After the merge of a shared-threshold sharded search: int32 [nq], 1 = the query needs the exact re-search
    (a shard overflowed, the merged list is short of k, or its k-th score < some shard's tau + eps)."""
    _require_cuda(D, I, bounds, flags)
    need = torch.empty((D.shape[0],), dtype=torch.int32, device=D.device)
    a, b, c, d, e = _lib.dl(D), _lib.dl(I), _lib.dl(bounds), _lib.dl(flags), _lib.dl(need)
    _lib.check(_lib.load().rs_topk_verify(a.ptr, b.ptr, c.ptr, d.ptr, e.ptr, _lib.stream_ptr()))
    return need
