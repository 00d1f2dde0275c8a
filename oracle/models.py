"""Model half of the oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of /root/reference/src/models.py, line by line:
  DeepCrossNetwork   src/models.py:14-55
  MultiTowerModel    src/models.py:59-102
  MultiTaskModel     src/models.py:106-131   (call only; compute_loss is training)

Third-party (Keras/TF) semantics restated from their public documentation,
because TensorFlow cannot be imported here (primitive arithmetic UNPINNED; the
composition below is held to a run of src/models.py itself, tests/test_reference_run.py):
  StringLookup(vocabulary=v, mask_token=None): OOV -> 0, v[i] -> i+1, int64
  Embedding(V, E): row gather
  Dense(units, activation): x @ kernel + bias, kernel stored (in, out)
  relu: max(x, 0);  sigmoid: 1 / (1 + exp(-x))
  tf.tensordot(a[B,D,1], w[D,1], axes=(1,0)) -> [B,1,1]

`dtype` selects the arithmetic: np.float32 is the reference's CPU arithmetic
(mixed_float16 is only enabled when TF sees a GPU, scripts/train.py:31-36);
np.float64 is the tolerance yardstick.

The weights are plain numpy arrays in a dict (see `synth.make_weights` in the
product package for the layout); Keras layer objects are not restated.
"""

from typing import Dict, List, Optional, Sequence

import numpy as np


def string_lookup(vocab_index: Dict[str, int], ids: Sequence[str]) -> np.ndarray:
    """keras.layers.StringLookup(vocabulary=vocab, mask_token=None) (src/models.py:70,73)."""
    return np.fromiter((vocab_index.get(s, 0) for s in ids), dtype=np.int64, count=len(ids))


def make_vocab_index(vocab: List[str]) -> Dict[str, int]:
    return {s: i + 1 for i, s in enumerate(vocab)}


def dense(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray, activation: Optional[str] = None) -> np.ndarray:
    y = x @ kernel.astype(x.dtype) + bias.astype(x.dtype)
    if activation == "relu":
        y = np.maximum(y, 0)
    elif activation == "sigmoid":
        y = 1.0 / (1.0 + np.exp(-y))
    elif activation is not None:
        raise ValueError(activation)
    return y.astype(x.dtype)


class DeepCrossNetwork:
    """src/models.py:14-55.  cross_w[i]: (D, 1), cross_b[i]: (D,), deep: [(kernel, bias), ...]."""

    def __init__(self, cross_w, cross_b, deep, dtype=np.float32):
        self.cross_weights = [np.asarray(w) for w in cross_w]
        self.cross_biases = [np.asarray(b) for b in cross_b]
        self.deep_nets = [(np.asarray(k), np.asarray(b)) for k, b in deep]
        self.cross_layers = len(self.cross_weights)
        self.dtype = dtype

    def __call__(self, inputs, training=None):
        dt = self.dtype
        x0 = np.asarray(inputs).astype(dt)
        xl = x0
        for i in range(self.cross_layers):                       # src/models.py:40
            xl_T = np.expand_dims(xl, axis=2)                     # :41  [B, D, 1]
            w_xl = np.tensordot(xl_T, self.cross_weights[i].astype(dt), axes=(1, 0))  # :42  [B, 1, 1]
            cross_term = np.squeeze(w_xl, axis=2)                 # :43  [B, 1]
            xl = x0 * cross_term + self.cross_biases[i].astype(dt) + xl   # :44  bias outside the product
        deep_out = x0                                             # :46  deep stack reads the *inputs*
        for kernel, bias in self.deep_nets:                       # :47-48  every deep layer is relu
            deep_out = dense(deep_out, kernel, bias, "relu")
        return np.concatenate([xl, deep_out], axis=1).astype(dt)  # :50  cross part first


class MultiTowerModel:
    """src/models.py:59-102.  Either side is optional; the missing side is None."""

    def __init__(self, weights, user_vocab: List[str], item_vocab: List[str], dtype=np.float32):
        self.w = weights
        self.user_vocab = list(user_vocab)
        self.item_vocab = list(item_vocab)
        self._uidx = make_vocab_index(self.user_vocab)
        self._iidx = make_vocab_index(self.item_vocab)
        self.dtype = dtype

    def _tower(self, layers, x):
        n = len(layers)
        for li, (kernel, bias) in enumerate(layers):              # :76-77  relu on all but the last Dense
            x = dense(x, kernel, bias, "relu" if li < n - 1 else None)
        return x

    def lookup_user(self, ids):
        return string_lookup(self._uidx, ids)

    def lookup_item(self, ids):
        return string_lookup(self._iidx, ids)

    def encode_rows(self, side: str, rows: np.ndarray) -> np.ndarray:
        table = self.w[f"{side}_embedding"]
        emb = table[rows].astype(self.dtype)                      # Embedding = row gather
        return self._tower(self.w[f"{side}_tower"], emb)

    def __call__(self, features, training=None):
        user_emb = None
        item_emb = None
        if "user_id" in features:                                 # :84-85
            user_emb = self.encode_rows("user", self.lookup_user(features["user_id"]))
        if "movie_id" in features:                                # :88-89
            item_emb = self.encode_rows("item", self.lookup_item(features["movie_id"]))
        return {"user_embedding": user_emb, "item_embedding": item_emb}


class MultiTaskModel:
    """src/models.py:106-131 -- the ranking forward (call) only."""

    def __init__(self, weights, user_vocab, item_vocab, dtype=np.float32):
        self.encoder = MultiTowerModel(weights, user_vocab, item_vocab, dtype)
        self.dcn = DeepCrossNetwork(weights["cross_w"], weights["cross_b"], weights["deep"], dtype)
        self.rating_head = weights["rating_head"]
        self.ctr_head = weights["ctr_head"]
        self.dtype = dtype

    def __call__(self, data, training=None):
        features, _ = data if isinstance(data, tuple) else (data, None)   # :126
        emb = self.encoder(features, training=training)                    # :127
        combined = np.concatenate([emb["user_embedding"], emb["item_embedding"]], axis=1)  # :128
        dcn_out = self.dcn(combined, training=training)                    # :129
        return {**emb,
                "rating_prediction": dense(dcn_out, *self.rating_head),    # :119,131
                "ctr_prediction": dense(dcn_out, *self.ctr_head, activation="sigmoid")}  # :120,131

    def rank_rows(self, user_rows: np.ndarray, item_rows: np.ndarray):
        """Same forward from already-looked-up row indices (the bench's integer-id entry)."""
        u = self.encoder.encode_rows("user", user_rows)
        it = self.encoder.encode_rows("item", item_rows)
        dcn_out = self.dcn(np.concatenate([u, it], axis=1))
        return {"user_embedding": u, "item_embedding": it,
                "rating_prediction": dense(dcn_out, *self.rating_head),
                "ctr_prediction": dense(dcn_out, *self.ctr_head, activation="sigmoid")}
