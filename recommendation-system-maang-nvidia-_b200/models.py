"""Host-side mirror of the reference's Keras classes for the ranking / encoding forward.

Same class names, constructor arguments, call signatures, attribute names and returned dict keys as
/root/reference/src/models.py, so code written against the reference (`model.encoder(...)`,
`model.dcn(x)`, `MultiTaskModel(...)(features)`) runs against this module unchanged:

  DeepCrossNetwork   src/models.py:14-55
  MultiTowerModel    src/models.py:59-102
  MultiTaskModel     src/models.py:106-131 (call only; compute_loss and the tfrs tasks are training)

Differences, all forced by the runtime: tensors are torch CUDA tensors instead of tf tensors; string
ids are Python/numpy strings (torch has no string tensor) looked up on the host through a dict
(StringLookup semantics: OOV -> 0, vocab[i] -> i+1); an *integer* tensor passed as `user_id` /
`movie_id` is taken as already-looked-up row numbers (an extension used by the bench).  Every layer
runs hand-written sm_100a kernels through the C ABI (ops.py); there is no CPU path.
"""

from typing import Any, Dict, List, Optional, Sequence

import os

import numpy as np
import torch

from . import ops
from .config import ModelConfig


def _dev(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("recsys_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class StringLookup:
    """keras.layers.StringLookup(vocabulary=..., mask_token=None): OOV -> 0, vocabulary[i] -> i + 1."""

    DEVICE_MIN_IDS = 8192          # lists of str at least this long ...
    DEVICE_MIN_WORDS = 200_000     # ... against vocabularies at least this large are looked up on the device (K0)

    def __init__(self, vocabulary: Sequence[str], mask_token=None, device=None):
        if mask_token is not None:
            raise ValueError("only mask_token=None is supported (as the reference uses, src/models.py:70,73)")
        self.vocabulary = list(vocabulary)
        self._index = {s: i + 1 for i, s in enumerate(self.vocabulary)}
        self.device = _dev(device)
        self._device_vocab = None          # K0 table, built by the first lookup_device call

    def vocabulary_size(self) -> int:
        return len(self.vocabulary) + 1

    def lookup_host(self, ids) -> np.ndarray:
        if isinstance(ids, (str, bytes)):
            ids = [ids]
        get = self._index.get
        return np.fromiter((get(s.decode() if isinstance(s, bytes) else str(s), 0) for s in ids), dtype=np.int64,
                           count=len(ids))

    def lookup_device(self, ids) -> torch.Tensor:
        """The same mapping computed on the GPU (K0, csrc/string_lookup.cu) for ids that arrive as bytes: a numpy `S`
        array, or a uint8 [B, W] byte matrix (numpy or torch, host or device; NUL-padded rows, ops.pack_strings).  The
        table is built on first use.  int64 [B] on the device, equal to lookup_host element for element."""
        if self._device_vocab is None:
            self._device_vocab = ops.DeviceVocabulary(self.vocabulary, device=self.device)
        if isinstance(ids, np.ndarray) and ids.dtype.kind in "SU":
            ids = ops.pack_strings(ids)
        if isinstance(ids, np.ndarray):
            ids = torch.from_numpy(np.ascontiguousarray(ids))
        if not isinstance(ids, torch.Tensor) or ids.dtype != torch.uint8 or ids.dim() != 2:
            raise TypeError("lookup_device takes a numpy S array or a uint8 [B, W] byte matrix")
        with torch.cuda.device(self.device):
            return self._device_vocab.lookup(ids.to(self.device, non_blocking=True))

    def __call__(self, ids) -> torch.Tensor:
        if isinstance(ids, torch.Tensor):
            if ids.dtype == torch.uint8 and ids.dim() == 2:
                return self.lookup_device(ids)
            if ids.dtype not in (torch.int32, torch.int64):
                raise TypeError("tensor ids must be int32/int64 row numbers or a uint8 [B, W] byte matrix")
            return ids.to(self.device)
        if isinstance(ids, np.ndarray) and ids.dtype.kind == "S":
            return self.lookup_device(ids)
        # large batches against large vocabularies: a dict of millions of strings misses the CPU caches on every probe
        # (65,536 ids against 1M words: 40 ms), packing the batch through Arrow and looking it up on the device does not
        # (10 ms + 0.02 ms).  Small vocabularies (the reference's 6,040 users) stay on the host dict, which is faster there.
        if (isinstance(ids, (list, tuple)) and len(ids) >= self.DEVICE_MIN_IDS and len(self.vocabulary) >= self.DEVICE_MIN_WORDS
                and os.environ.get("RS_DEVICE_LOOKUP", "1") != "0"):
            packed = ops.pack_str_list(ids)
            if packed is not None:
                return self.lookup_device(packed)
        rows = self.lookup_host(ids)
        return torch.from_numpy(rows).to(self.device, non_blocking=True)


class Embedding:
    """keras.layers.Embedding(input_dim, output_dim): a row gather (K1)."""

    def __init__(self, input_dim: int, output_dim: int, device=None, seed: Optional[int] = None):
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        g = torch.Generator(device="cpu")
        g.manual_seed(0 if seed is None else seed)
        w = (torch.rand((self.input_dim, self.output_dim), generator=g, dtype=torch.float32) - 0.5) * 0.1
        self._version = 0
        self.embeddings = w.to(_dev(device))

    # every assignment of the table bumps `_version`; handles that copy or borrow weights (the fused ranker, CUDA-graph
    # sessions) compare version signatures instead of data pointers, which the caching allocator may reuse
    @property
    def embeddings(self) -> torch.Tensor:
        return self._embeddings

    @embeddings.setter
    def embeddings(self, w: torch.Tensor) -> None:
        self._embeddings = w
        self._version += 1

    def __call__(self, rows: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return ops.gather_concat([self.embeddings], [rows], out=out)

    def get_weights(self):
        return [self.embeddings.cpu().numpy()]

    def set_weights(self, ws):
        (w,) = ws
        w = np.asarray(w, dtype=np.float32)
        if w.shape != (self.input_dim, self.output_dim):
            raise ValueError(f"embedding shape {w.shape} != {(self.input_dim, self.output_dim)}")
        self.embeddings = torch.from_numpy(np.ascontiguousarray(w)).to(self.embeddings.device)


class Dense:
    """keras.layers.Dense(units, activation): act(x @ kernel + bias), kernel (in, out); built lazily (K2)."""

    def __init__(self, units: int, activation=None, name: Optional[str] = None, device=None, **_ignored):
        self.units, self.activation, self.name = int(units), activation, name
        self._version = 0
        self._wt_cache = None
        self._x3_cache = None
        self._kernel: Optional[torch.Tensor] = None
        self._bias: Optional[torch.Tensor] = None
        self._device = _dev(device)

    # assigning kernel / bias (build, set_weights, or directly) bumps `_version`: the bf16 copy below and every handle
    # built from this layer (fused ranker, head packing, CUDA-graph sessions) are keyed on it
    @property
    def kernel(self) -> Optional[torch.Tensor]:
        return self._kernel

    @kernel.setter
    def kernel(self, k: Optional[torch.Tensor]) -> None:
        self._kernel = k
        self._version += 1
        self._wt_cache = None
        self._x3_cache = None

    @property
    def bias(self) -> Optional[torch.Tensor]:
        return self._bias

    @bias.setter
    def bias(self, b: Optional[torch.Tensor]) -> None:
        self._bias = b
        self._version += 1

    def build(self, input_dim: int, seed: int = 0):
        lim = float(np.sqrt(6.0 / (input_dim + self.units)))
        g = torch.Generator(device="cpu")
        g.manual_seed(seed)
        k = (torch.rand((input_dim, self.units), generator=g, dtype=torch.float32) * 2 - 1) * lim
        self.kernel = k.to(self._device)
        self.bias = torch.zeros((self.units,), dtype=torch.float32, device=self._device)

    def __call__(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.kernel is None:
            self.build(int(x.shape[-1]))
        if x3_chain_applies([self], int(x.shape[0]), int(x.shape[1]), x=x, out=out):
            return x3_chain([([self], {"x": x})], [out])[0]
        return ops.dense(x, self.kernel, self.bias, self.activation, out=out)

    # fp32 mode has two kernels with the same stated tolerance (1e-5 of tensor scale): the FFMA GEMM (dense.cu) and, from
    # X3_MIN_ROWS rows, the tcgen05 kernel that splits every fp32 operand into two fp16 numbers (dense_x3.cu).
    # RS_FP32_TENSOR=0 keeps everything on the FFMA kernel.
    X3_MIN_ROWS = 512

    def wt_x3(self):
        """(hi, lo) fp16 [out, in] images of the kernel for the fp32 tensor mode (cached per weight version)."""
        c = self._x3_cache
        if c is None or c[0] != self._version:
            c = (self._version, ops.dense_x3_prepare(self.kernel))
            self._x3_cache = c
        return c[1]

    def wt_bf16(self) -> torch.Tensor:
        """The kernel transposed to [out, in] bf16 -- the K-major B operand of the tcgen05 path (cached)."""
        c = self._wt_cache
        if c is None or c[0] != self._version:
            c = (self._version, self.kernel.t().contiguous().to(torch.bfloat16))
            self._wt_cache = c
        return c[1]

    def call_bf16(self, x_bf16: torch.Tensor, out_bf16=None, out_f32=None, want_bf16=True):
        if self.kernel is None:
            self.build(int(x_bf16.shape[-1]))
        return ops.dense_bf16(x_bf16, self.wt_bf16(), self.bias, self.activation, out_bf16=out_bf16, out_f32=out_f32,
                              want_bf16=want_bf16)

    def get_weights(self):
        return [self.kernel.cpu().numpy(), self.bias.cpu().numpy()]

    def set_weights(self, ws):
        k, b = ws
        k = np.ascontiguousarray(np.asarray(k, dtype=np.float32))
        b = np.ascontiguousarray(np.asarray(b, dtype=np.float32)).reshape(-1)
        if k.shape[1] != self.units or b.shape[0] != self.units:
            raise ValueError("Dense weight shapes do not match units")
        self.kernel = torch.from_numpy(k).to(self._device)
        self.bias = torch.from_numpy(b).to(self._device)


def x3_chain_applies(layers, rows: int, in_dim: int, x: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> bool:
    """Whether a chain of Dense layers runs on the fp32 tensor kernel (dense_x3.cu): ranking-size batch, every layer shape
    served, and the fp32 input / output views aligned the way the kernel's 128-bit loads and 256-bit stores need."""
    if rows < Dense.X3_MIN_ROWS or os.environ.get("RS_FP32_TENSOR", "1") == "0":
        return False
    d = in_dim
    for layer in layers:
        if layer.kernel is not None and int(layer.kernel.shape[0]) != d:
            return False
        if not ops.dense_x3_supported(d, layer.units, layer.activation):
            return False
        d = layer.units
    if x is not None and (x.dtype != torch.float32 or x.stride(1) != 1 or x.stride(0) % 4 or x.data_ptr() % 16):
        return False
    if out is not None and (out.stride(1) != 1 or out.stride(0) % 8 or out.data_ptr() % 32):
        return False
    return True


def x3_chain(problems, outs, split_outs=None):
    """Dense chains on the fp32 tensor kernel.  problems: one or two (layers, first_input) pairs with identical layer shapes
    -- both run in the same launches (the two towers); first_input: {"x": f32 rows}, {"x": table, "ids": rows} or {"x_hi",
    "x_lo"}.  Between layers the activations travel as split fp16 planes (ops.split_planes); the last layer writes fp32 into
    outs[p] (allocated when None) and, when split_outs[p] = (hi, lo) is given, the split planes for a following chain."""
    n_layers = len(problems[0][0])
    cur = [dict(first) for _, first in problems]
    results = list(outs)
    for li in range(n_layers):
        last = li == n_layers - 1
        batch = []
        nxt = []
        for pi, (layers, _) in enumerate(problems):
            layer = layers[li]
            if layer.kernel is None:
                layer.build(int(layers[li - 1].units) if li else int(cur[pi]["x"].shape[1]))
            hi, lo = layer.wt_x3()
            src = cur[pi]
            rows = int(src["ids"].shape[0]) if src.get("ids") is not None else int((src["x"] if "x" in src else src["x_hi"]).shape[0])
            L = dict(src, wt_hi=hi, wt_lo=lo, bias=layer.bias)
            if last:
                if results[pi] is None:
                    results[pi] = torch.empty((rows, layer.units), dtype=torch.float32, device=hi.device)
                L["y"] = results[pi]
                if split_outs is not None and split_outs[pi] is not None:
                    L["y_hi"], L["y_lo"] = split_outs[pi]
            else:
                L["y_hi"], L["y_lo"] = ops.split_planes(rows, layer.units, hi.device)
                nxt.append({"x_hi": L["y_hi"], "x_lo": L["y_lo"]})
            batch.append(L)
        ops.dense_x3_layers(batch, problems[0][0][li].activation)
        cur = nxt
    return results


class Sequential:
    """keras.Sequential of Dense layers; `out` lets the last layer write into a wider buffer."""

    def __init__(self, layers: List[Dense]):
        self.layers = list(layers)

    def __call__(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n = len(self.layers)
        if n and x3_chain_applies(self.layers, int(x.shape[0]), int(x.shape[1]), x=x, out=out):
            return x3_chain([(self.layers, {"x": x})], [out])[0]
        for i, layer in enumerate(self.layers):
            x = layer(x, out=out if i == n - 1 else None)
        return x

    def call_on_rows(self, embedding: "Embedding", rows: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """tower(embedding(rows)) (src/models.py:85,89): at ranking batches the gather rides in the first layer's loads."""
        table = embedding.embeddings
        if (self.layers and rows.dtype in (torch.int32, torch.int64) and rows.is_contiguous()
                and x3_chain_applies(self.layers, int(rows.shape[0]), int(table.shape[1]), x=table, out=out)):
            return x3_chain([(self.layers, {"x": table, "ids": rows})], [out])[0]
        return self(embedding(rows), out=out)

    def call_bf16(self, x_bf16: torch.Tensor, out_f32: Optional[torch.Tensor] = None,
                  out_bf16: Optional[torch.Tensor] = None, want_bf16_out: bool = False):
        """tcgen05 chain: bf16 activations between layers; the last layer writes fp32 (and optionally bf16)."""
        n = len(self.layers)
        y32 = None
        for i, layer in enumerate(self.layers):
            last = i == n - 1
            if last:
                if out_f32 is None:
                    out_f32 = torch.empty((x_bf16.shape[0], layer.units), dtype=torch.float32, device=x_bf16.device)
                x_bf16, y32 = layer.call_bf16(x_bf16, out_bf16=out_bf16, out_f32=out_f32,
                                              want_bf16=want_bf16_out or out_bf16 is not None)
            else:
                x_bf16, _ = layer.call_bf16(x_bf16)
        return y32, x_bf16


class DeepCrossNetwork:
    """Deep & Cross Network (src/models.py:14-55): rank-1 cross stack + deep relu MLP, concatenated."""

    def __init__(self, cross_layers: int = 3, deep_layers: List[int] = None, dropout_rate: float = 0.2,
                 l2_reg: float = 1e-5, device=None, **kwargs):
        self.cross_layers = cross_layers
        self.deep_layers = deep_layers or [256, 128, 64]
        self.dropout_rate = dropout_rate   # accepted and unused, exactly as in the reference (no Dropout layer)
        self.l2_reg = l2_reg               # training-only regulariser
        self._device = _dev(device)
        self.cross_weights: List[torch.Tensor] = []
        self.cross_biases: List[torch.Tensor] = []
        self._cross_w: Optional[torch.Tensor] = None   # stacked [L, D] storage the kernel reads
        self._cross_b: Optional[torch.Tensor] = None
        self.deep_nets = [Dense(units, activation="relu", device=self._device) for units in self.deep_layers]
        self.built = False

    def build(self, input_shape):
        input_dim = int(input_shape[-1])
        L = self.cross_layers
        lim = float(np.sqrt(6.0 / (input_dim + 1)))
        g = torch.Generator(device="cpu")
        g.manual_seed(7)
        w = (torch.rand((L, input_dim), generator=g, dtype=torch.float32) * 2 - 1) * lim
        self._set_cross(w.to(self._device), torch.zeros((L, input_dim), dtype=torch.float32, device=self._device))
        self.built = True

    def _set_cross(self, w: torch.Tensor, b: torch.Tensor):
        self._cross_w, self._cross_b = w.contiguous(), b.contiguous()
        self._version = getattr(self, "_version", 0) + 1
        # the reference's variable shapes: cross_w_i (D, 1), cross_b_i (D,)  (src/models.py:34-35)
        self.cross_weights = [self._cross_w[i].unsqueeze(1) for i in range(w.shape[0])]
        self.cross_biases = [self._cross_b[i] for i in range(w.shape[0])]

    def _cross(self, D: int):
        """The stacked [L, D] cross kernels / biases; with zero cross layers the width comes from the input."""
        if self._cross_w.shape[0] == 0 and int(self._cross_w.shape[1]) != D:
            z = torch.zeros((0, D), dtype=torch.float32, device=self._device)
            self._set_cross(z, z.clone())
        return self._cross_w, self._cross_b

    def set_cross_weights(self, cross_w: Sequence[np.ndarray], cross_b: Sequence[np.ndarray]):
        w = np.stack([np.asarray(x, dtype=np.float32).reshape(-1) for x in cross_w]) if len(cross_w) else \
            np.zeros((0, 0), dtype=np.float32)
        b = np.stack([np.asarray(x, dtype=np.float32).reshape(-1) for x in cross_b]) if len(cross_b) else \
            np.zeros((0, 0), dtype=np.float32)
        self.cross_layers = w.shape[0]
        self._set_cross(torch.from_numpy(w).to(self._device), torch.from_numpy(b).to(self._device))
        self.built = True

    def __call__(self, inputs: torch.Tensor, training=None) -> torch.Tensor:
        if not self.built:
            self.build(inputs.shape)
        B, D = int(inputs.shape[0]), int(inputs.shape[1])
        for i, layer in enumerate(self.deep_nets):
            if layer.kernel is None:
                layer.build(D if i == 0 else self.deep_layers[i - 1], seed=20 + i)
        out = torch.empty((B, D + self.deep_layers[-1]), dtype=torch.float32, device=inputs.device)
        ops.dcn_cross(inputs, *self._cross(D), out=out[:, :D])                      # src/models.py:38-44
        Sequential(self.deep_nets)(inputs, out=out[:, D:])                           # :46-48, concat :50
        return out

    def get_config(self):
        return {"cross_layers": self.cross_layers, "deep_layers": self.deep_layers,
                "dropout_rate": self.dropout_rate, "l2_reg": self.l2_reg}


class MultiTowerModel:
    """Two-tower encoder (src/models.py:59-102)."""

    def __init__(self, config: ModelConfig, user_vocab: List[str], item_vocab: List[str],
                 feature_specs: Dict[str, Any] = None, device=None, precision: str = "fp32", **kwargs):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' (exact FFMA path) or 'bf16' (tcgen05 path)")
        self.precision = precision
        self.config = config
        self.user_vocab = user_vocab
        self.item_vocab = item_vocab
        self.feature_specs = feature_specs   # stored and unused, as in the reference (src/models.py:68)
        dev = _dev(device)
        self.device = dev
        E = config.embedding_dim
        self.user_lookup = StringLookup(vocabulary=user_vocab, mask_token=None, device=dev)
        self.user_embedding = Embedding(len(user_vocab) + 1, E, device=dev, seed=1)
        self.item_lookup = StringLookup(vocabulary=item_vocab, mask_token=None, device=dev)
        self.item_embedding = Embedding(len(item_vocab) + 1, E, device=dev, seed=2)
        self.user_tower = Sequential([Dense(u, activation="relu", device=dev) for u in config.user_tower_dims]
                                     + [Dense(E, device=dev)])
        self.item_tower = Sequential([Dense(u, activation="relu", device=dev) for u in config.item_tower_dims]
                                     + [Dense(E, device=dev)])
        for tower, dims, s in ((self.user_tower, config.user_tower_dims, 100), (self.item_tower, config.item_tower_dims, 200)):
            fan_in = E
            for i, layer in enumerate(tower.layers):
                layer.build(fan_in, seed=s + i)
                fan_in = layer.units

    def _bf16_ok(self, n: int) -> bool:
        # the tensor path is for ranking-size batches; tiny batches (the /recommend path) stay on the exact path
        return self.precision == "bf16" and n >= 128

    def encode_user(self, ids, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        rows = self.user_lookup(ids)
        if self._bf16_ok(int(rows.shape[0])):
            return self.user_tower.call_bf16(ops.gather_bf16(self.user_embedding.embeddings, rows), out_f32=out)[0]
        return self.user_tower.call_on_rows(self.user_embedding, rows, out=out)            # src/models.py:85

    def encode_item(self, ids, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        rows = self.item_lookup(ids)
        if self._bf16_ok(int(rows.shape[0])):
            return self.item_tower.call_bf16(ops.gather_bf16(self.item_embedding.embeddings, rows), out_f32=out)[0]
        return self.item_tower.call_on_rows(self.item_embedding, rows, out=out)            # src/models.py:89

    def __call__(self, features: Dict[str, Any], training=None) -> Dict[str, Optional[torch.Tensor]]:
        user_emb = None
        item_emb = None
        if "user_id" in features:
            user_emb = self.encode_user(features["user_id"])
        if "movie_id" in features:
            item_emb = self.encode_item(features["movie_id"])
        return {"user_embedding": user_emb, "item_embedding": item_emb}

    def get_config(self):
        return {"config": self.config.to_dict(), "user_vocab": self.user_vocab, "item_vocab": self.item_vocab,
                "feature_specs": self.feature_specs}

    @classmethod
    def from_config(cls, config):
        config = dict(config)
        config["config"] = ModelConfig(**config.pop("config"))
        return cls(**config)


class _FusedRanker:
    """rs_ranker handle: the whole bf16-mode ranking forward behind ONE C call per batch (chain_tc.cu)."""

    def __init__(self, model: "MultiTaskModel"):
        import ctypes
        from . import _lib
        enc, dcn = model.encoder, model.dcn
        self._lib = _lib.load()
        self._tables = (enc.user_embedding.embeddings, enc.item_embedding.embeddings)   # borrowed by the handle
        tw = [l.wt_bf16() for t in (enc.user_tower, enc.item_tower) for l in t.layers]
        tb = [l.bias for t in (enc.user_tower, enc.item_tower) for l in t.layers]
        dw = [l.wt_bf16() for l in dcn.deep_nets]
        db = [l.bias for l in dcn.deep_nets]
        hw, hb = model._heads_packed()
        dcn._cross(2 * model.config.embedding_dim)
        keep = []

        def arr(objs):
            bs = [_lib.dl(o) for o in objs]
            keep.append(bs)
            return (_lib._DLP * len(bs))(*[_lib.ptr(b) for b in bs])

        singles = [_lib.dl(t) for t in (self._tables[0], self._tables[1], dcn._cross_w, dcn._cross_b, hw, hb)]
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.rs_ranker_create(singles[0].ptr, singles[1].ptr, arr(tw), arr(tb), arr(dw), arr(db),
                                              singles[2].ptr, singles[3].ptr, singles[4].ptr, singles[5].ptr,
                                              ctypes.byref(self._h)))

    def forward(self, rows_u: torch.Tensor, rows_i: torch.Tensor, from_embedding_rows: bool = False):
        """rows_*: int row numbers [B] -- or, with from_embedding_rows, the looked-up f32 [B,128] rows themselves."""
        from . import _lib
        B = int(rows_u.shape[0])
        dev = rows_u.device
        combined = torch.empty((B, 256), dtype=torch.float32, device=dev)
        logits = torch.empty((2, B, 1), dtype=torch.float32, device=dev)
        rating, ctr = logits[0], logits[1]
        a, b, c, d, e = _lib.dl(rows_u), _lib.dl(rows_i), _lib.dl(combined), _lib.dl(rating), _lib.dl(ctr)
        fn = self._lib.rs_ranker_forward_rows if from_embedding_rows else self._lib.rs_ranker_forward
        _lib.check(fn(self._h, a.ptr, b.ptr, c.ptr, d.ptr, e.ptr, _lib.stream_ptr()))
        return combined, rating, ctr

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.rs_ranker_free(h)
            except Exception:
                pass


class RankingSession:
    """MultiTaskModel.call for ONE fixed batch size, replayed as one CUDA graph (bf16 tensor mode: two launches; fp32
    mode: seven launches of the fp32 tensor kernel, whose eager host work is longer than the kernels at batch 65,536).

    The eager call costs ~50 us of host work per batch (output allocation, DLPack export, two launches, the stream-
    ordered scratch allocation), which is as long as the two kernels themselves at batch 65,536.  A session owns
    static input buffers (`user_rows`, `item_rows`: int32 row numbers, what StringLookup emits) and static outputs
    (`outputs`: the four-key dict of MultiTaskModel.call, views of graph-owned memory that the next replay
    overwrites), captures the forward once and replays it:

        sess = model.ranking_session(65536)
        sess.user_rows.copy_(rows_u, non_blocking=True); sess.item_rows.copy_(rows_i, non_blocking=True)
        out = sess.run()                      # or: out = sess({"user_id": rows_u, "movie_id": rows_i})

    Same kernels, same arithmetic as the eager call (bit-identical outputs, tests/test_gpu_ranking_bf16.py).  A session
    snapshots the model's weights at capture time and re-captures itself when the model's weights_signature() changes."""

    def __init__(self, model: "MultiTaskModel", batch: int, ids_dtype=torch.int32):
        self.model, self.batch = model, int(batch)
        dev = model.device
        self.user_rows = torch.zeros((self.batch,), dtype=ids_dtype, device=dev)
        self.item_rows = torch.zeros((self.batch,), dtype=ids_dtype, device=dev)
        self._capture()

    def _capture(self) -> None:
        model, dev = self.model, self.model.device
        feats = {"user_id": self.user_rows, "movie_id": self.item_rows}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):           # warm-up off the capture: builds the handle, sets kernel attributes
            for _ in range(2):
                model(feats)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.outputs = model(feats)
        # the captured launches point into the ranker handle's own copies of the weights: keep that handle alive for as
        # long as the graph can be replayed, and remember which weights it was captured with
        self._ranker = getattr(model, "_ranker", None)
        self._sig = model.weights_signature()

    def run(self) -> Dict[str, torch.Tensor]:
        if self._sig != self.model.weights_signature():      # weights changed since capture: capture again (rare)
            self._capture()
        self._graph.replay()
        return self.outputs

    def __call__(self, data, training=None) -> Dict[str, torch.Tensor]:
        features, _ = data if isinstance(data, tuple) else (data, None)
        enc = self.model.encoder
        for dst, key, lookup in ((self.user_rows, "user_id", enc.user_lookup), (self.item_rows, "movie_id", enc.item_lookup)):
            rows = features[key]
            if not isinstance(rows, torch.Tensor):
                rows = torch.from_numpy(lookup.lookup_host(rows))
            if int(rows.shape[0]) != self.batch:
                raise ValueError(f"this session was captured for batch {self.batch}, got {int(rows.shape[0])}")
            dst.copy_(rows, non_blocking=True)
        return self.run()


class MultiTaskModel:
    """Ranking forward of the multi-task model (src/models.py:106-131)."""

    def __init__(self, config: ModelConfig, user_vocab: List[str], item_vocab: List[str],
                 feature_specs: Dict[str, Any] = None, class_weights: Dict[int, float] = None, device=None,
                 precision: str = "fp32", **kwargs):
        self.config = config
        dev = _dev(device)
        self.device = dev
        self.precision = precision
        self.encoder = MultiTowerModel(config, user_vocab, item_vocab, feature_specs, device=dev, precision=precision)
        self.class_weights = class_weights
        self.dcn = DeepCrossNetwork(config.cross_layers, config.dnn_dims, config.dropout_rate, config.l2_reg, device=dev)
        self.rating_head = Dense(1, name="rating_pred", device=dev)
        self.ctr_head = Dense(1, activation="sigmoid", name="ctr_pred", device=dev)
        W = 2 * config.embedding_dim + config.dnn_dims[-1]
        self.rating_head.build(W, seed=31)
        self.ctr_head.build(W, seed=32)
        self._head_cache = None

    def weights_signature(self) -> tuple:
        """Version counters of every weight the forward reads.  Handles that copy or borrow weights (the fused ranker,
        the head packing, CUDA-graph sessions) remember the signature they were built from and are rebuilt when it
        changes -- so Keras-style `layer.set_weights(...)`, `set_cross_weights` or a plain attribute assignment can never
        leave a stale copy in use.  (Versions, not data pointers: the caching allocator reuses addresses.)"""
        enc, dcn = self.encoder, self.dcn
        layers = list(enc.user_tower.layers) + list(enc.item_tower.layers) + list(dcn.deep_nets) + [self.rating_head, self.ctr_head]
        return (enc.user_embedding._version, enc.item_embedding._version, getattr(dcn, "_version", 0),
                tuple(l._version for l in layers))

    def _heads_packed(self):
        """[2, W] / [2] packing of the two Dense(1) heads for the fused heads kernel (K3b)."""
        key = (self.rating_head._version, self.ctr_head._version)
        if self._head_cache is None or self._head_cache[0] != key:
            hw = torch.stack([self.rating_head.kernel[:, 0], self.ctr_head.kernel[:, 0]]).contiguous()
            hb = torch.cat([self.rating_head.bias, self.ctr_head.bias]).contiguous()
            self._head_cache = (key, hw, hb)
        return self._head_cache[1], self._head_cache[2]

    def __call__(self, data, training=None) -> Dict[str, torch.Tensor]:
        features, _ = data if isinstance(data, tuple) else (data, None)                    # src/models.py:126
        enc = self.encoder
        E = self.config.embedding_dim
        rows_u = enc.user_lookup(features["user_id"])
        rows_i = enc.item_lookup(features["movie_id"])
        B = int(rows_u.shape[0])
        if self.precision == "bf16" and B >= 1:
            return self._call_bf16(rows_u, rows_i)
        # the towers write straight into the two halves of the DCN input: the concat (:128) is free
        combined = torch.empty((B, 2 * E), dtype=torch.float32, device=self.device)
        hw, hb = self._heads_packed()
        dcn = self.dcn
        if not dcn.built:
            dcn.build(combined.shape)
        cw, cb = dcn._cross(2 * E)
        ut, it = enc.user_tower.layers, enc.item_tower.layers
        tables = (enc.user_embedding.embeddings, enc.item_embedding.embeddings)
        ids_ok = all(r.dtype in (torch.int32, torch.int64) and r.is_contiguous() for r in (rows_u, rows_i)) and int(rows_i.shape[0]) == B
        if (ids_ok and 2 * E == 256 and dcn.deep_layers[-1] == 128 and cw.shape[0] <= 4
                and [l.units for l in ut] == [l.units for l in it] and [l.activation for l in ut] == [l.activation for l in it]
                and x3_chain_applies(ut, B, E, x=tables[0], out=combined[:, :E]) and x3_chain_applies(it, B, E, x=tables[1], out=combined[:, E:])
                and x3_chain_applies(dcn.deep_nets, B, 2 * E)):
            # fp32 tensor mode (dense_x3.cu), 7 launches: both towers layer by layer in the same launches (gather fused into
            # the first, split fp16 planes between layers), the deep net fed from the split copy of the tower outputs, then
            # cross stack + heads in ONE pass over (x0, deep) -- the [B, 384] concat of :50 is never written
            chi, clo = ops.split_planes(B, 2 * E, self.device)
            user_emb, item_emb = x3_chain([(ut, {"x": tables[0], "ids": rows_u}), (it, {"x": tables[1], "ids": rows_i})],
                                          [combined[:, :E], combined[:, E:]],                       # :127 / :85, :89
                                          split_outs=[(chi[:, :E], clo[:, :E]), (chi[:, E:], clo[:, E:])])
            deep = x3_chain([(dcn.deep_nets, {"x_hi": chi, "x_lo": clo})], [None])[0]               # :46-48
            rating, ctr = ops.cross_heads(combined, deep, cw, cb, hw, hb)                           # :38-44, :131
        else:
            user_emb = enc.user_tower.call_on_rows(enc.user_embedding, rows_u, out=combined[:, :E])   # :127 / :85
            item_emb = enc.item_tower.call_on_rows(enc.item_embedding, rows_i, out=combined[:, E:])   # :127 / :89
            dcn_out = dcn(combined, training=training)                                     # :129
            rating, ctr = ops.heads(dcn_out, hw, hb)                                       # :131
        return {"user_embedding": user_emb, "item_embedding": item_emb,
                "rating_prediction": rating, "ctr_prediction": ctr}

    def _call_bf16(self, rows_u: torch.Tensor, rows_i: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Ranking forward with every Dense on tcgen05 (bf16 operands, fp32 accumulation).  The tower outputs are
        produced in fp32 (model outputs, DCN cross input) and in bf16 (deep-net input); cross and heads stay fp32."""
        enc, E, dev = self.encoder, self.config.embedding_dim, self.device
        B = int(rows_u.shape[0])
        D = 2 * E
        combined = torch.empty((B, D), dtype=torch.float32, device=dev)
        combined_bf16 = torch.empty((B, D), dtype=torch.bfloat16, device=dev)
        dcn = self.dcn
        if not dcn.built:
            dcn.build(combined.shape)
        dcn_out = torch.empty((B, D + dcn.deep_layers[-1]), dtype=torch.float32, device=dev)
        cfg = self.config
        fused = (E == 128 and list(cfg.user_tower_dims) == [256, 128, 64] and list(cfg.item_tower_dims) == [256, 128, 64]
                 and list(dcn.deep_layers) == [256, 128] and int(rows_i.shape[0]) == B)
        dcn._cross(D)
        if fused and int(dcn._cross_w.shape[1]) == 256:
            combined, rating, ctr = self._fused_ranker().forward(rows_u.contiguous(), rows_i.contiguous())
            return {"user_embedding": combined[:, :E], "item_embedding": combined[:, E:],
                    "rating_prediction": rating, "ctr_prediction": ctr}
        if fused:
            # both towers in ONE kernel (gather + 4 layers each, activations in TMEM), then the deep net in one kernel
            towers = [[(l.wt_bf16(), l.bias) for l in t.layers] for t in (enc.user_tower, enc.item_tower)]
            ops.chain_forward_bf16(towers, ["relu", "relu", "relu", None],
                                   tables=[enc.user_embedding.embeddings, enc.item_embedding.embeddings],
                                   ids=[rows_u, rows_i], y_f32=[combined[:, :E], combined[:, E:]],
                                   y_bf16=[combined_bf16[:, :E], combined_bf16[:, E:]])
            ops.chain_forward_bf16([[(l.wt_bf16(), l.bias) for l in dcn.deep_nets]], ["relu", "relu"], x=[combined_bf16],
                                   y_f32=[dcn_out[:, D:]])
        else:
            enc.user_tower.call_bf16(ops.gather_bf16(enc.user_embedding.embeddings, rows_u), out_f32=combined[:, :E],
                                     out_bf16=combined_bf16[:, :E])
            enc.item_tower.call_bf16(ops.gather_bf16(enc.item_embedding.embeddings, rows_i), out_f32=combined[:, E:],
                                     out_bf16=combined_bf16[:, E:])
            Sequential(dcn.deep_nets).call_bf16(combined_bf16, out_f32=dcn_out[:, D:])
        ops.dcn_cross(combined, dcn._cross_w, dcn._cross_b, out=dcn_out[:, :D])
        hw, hb = self._heads_packed()
        rating, ctr = ops.heads(dcn_out, hw, hb)
        return {"user_embedding": combined[:, :E], "item_embedding": combined[:, E:],
                "rating_prediction": rating, "ctr_prediction": ctr}

    def _fused_ranker(self) -> "_FusedRanker":
        """The rs_ranker handle for the CURRENT weights (rebuilt whenever weights_signature() has changed)."""
        sig = self.weights_signature()
        r = getattr(self, "_ranker", None)
        if r is None or getattr(self, "_ranker_sig", None) != sig:
            self.dcn._cross(2 * self.config.embedding_dim)
            r = self._ranker = _FusedRanker(self)
            self._ranker_sig = self.weights_signature()      # _cross() above may have bumped the DCN version
        return r

    def ranking_session(self, batch: int, ids_dtype=torch.int32) -> RankingSession:
        """A CUDA-graph replay of this forward for one fixed batch size (see RankingSession)."""
        return RankingSession(self, batch, ids_dtype)

    def rank_from_embedding_rows(self, user_rows: torch.Tensor, item_rows: torch.Tensor) -> Dict[str, torch.Tensor]:
        """The same forward, fed with already-looked-up embedding rows f32 [B, E] (the output of
        sharded.ShardedEmbedding.lookup when the tables are partitioned over GPUs; bf16 tensor mode only)."""
        E = self.config.embedding_dim
        combined, rating, ctr = self._fused_ranker().forward(user_rows, item_rows, from_embedding_rows=True)
        return {"user_embedding": combined[:, :E], "item_embedding": combined[:, E:],
                "rating_prediction": rating, "ctr_prediction": ctr}

    def compute_loss(self, data, training=False):
        raise NotImplementedError("training (tfrs tasks, src/models.py:133-148) is out of scope of the serving path")

    def get_config(self):
        return {"config": self.config.to_dict(), "user_vocab": self.encoder.user_vocab,
                "item_vocab": self.encoder.item_vocab, "feature_specs": self.encoder.feature_specs,
                "class_weights": self.class_weights}

    @classmethod
    def from_config(cls, config):
        config = dict(config)
        config["config"] = ModelConfig(**config.pop("config"))
        return cls(**config)


def load_weights(model, weights: Dict) -> None:
    """Install a synth.make_weights()-style dict into a MultiTowerModel or MultiTaskModel."""
    enc = model.encoder if isinstance(model, MultiTaskModel) else model
    enc.user_embedding.set_weights([weights["user_embedding"]])
    enc.item_embedding.set_weights([weights["item_embedding"]])
    for tower, key in ((enc.user_tower, "user_tower"), (enc.item_tower, "item_tower")):
        if len(tower.layers) != len(weights[key]):
            raise ValueError(f"{key}: layer count mismatch")
        for layer, (k, b) in zip(tower.layers, weights[key]):
            layer.set_weights([k, b])
    if isinstance(model, MultiTaskModel):
        model.dcn.set_cross_weights(weights["cross_w"], weights["cross_b"])
        if len(model.dcn.deep_nets) != len(weights["deep"]):
            raise ValueError("deep: layer count mismatch")
        for layer, (k, b) in zip(model.dcn.deep_nets, weights["deep"]):
            layer.set_weights([k, b])
        model.rating_head.set_weights(list(weights["rating_head"]))
        model.ctr_head.set_weights(list(weights["ctr_head"]))
        model._head_cache = None
        model._ranker = None          # (redundant with the version signature; kept so the old handle is freed at once)
