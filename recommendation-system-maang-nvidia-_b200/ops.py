"""Operator-level Python entry points: one function per C-ABI compute call.

All tensors are torch CUDA tensors (torch only supplies device memory and the stream); every function
launches hand-written sm_100a kernels from librecsys_b200.so on torch's current stream and raises on
any failure -- there is no eager/PyTorch fallback.
"""

from typing import Optional, Sequence

import torch

from . import _lib


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise TypeError("expected CUDA torch tensors (no CPU path exists)")


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


def dcn_cross(x0: torch.Tensor, cross_w: torch.Tensor, cross_b: torch.Tensor,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K3a: the DCN-v1 cross stack (src/models.py:38-44). cross_w, cross_b: [n_layers, D]."""
    _require_cuda(x0, cross_w, cross_b, out)
    if out is None:
        out = torch.empty_like(x0)
    a, w, b, o = _lib.dl(x0), _lib.dl(cross_w), _lib.dl(cross_b), _lib.dl(out)
    _lib.check(_lib.load().rs_dcn_cross_forward(a.ptr, w.ptr, b.ptr, o.ptr, _lib.stream_ptr()))
    return out


def heads(z: torch.Tensor, head_w: torch.Tensor, head_b: torch.Tensor):
    """K3b: rating = z.w_r + b_r, ctr = sigmoid(z.w_c + b_c) (src/models.py:119-120,131). -> ([B,1], [B,1])"""
    _require_cuda(z, head_w, head_b)
    rating = torch.empty((z.shape[0], 1), dtype=torch.float32, device=z.device)
    ctr = torch.empty((z.shape[0], 1), dtype=torch.float32, device=z.device)
    a, w, b, r, c = _lib.dl(z), _lib.dl(head_w), _lib.dl(head_b), _lib.dl(rating), _lib.dl(ctr)
    _lib.check(_lib.load().rs_heads_forward(a.ptr, w.ptr, b.ptr, r.ptr, c.ptr, _lib.stream_ptr()))
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
