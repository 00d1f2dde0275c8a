"""EXTENSIONS: model-side pieces the north-star names but the reference does not contain.

    * DCN-v2 matrix cross  x_{l+1} = x0 * (x_l W_l + b_l) + x_l  with W_l in R^{D x D}  -- the reference's cross layer is the
      rank-1 DCN-v1 form (src/models.py:38-44); the matrix form is the one whose product really is a GEMM.
    * extra categorical tables next to the id embeddings (age / gender / occupation codes on the user side, any integer-coded
      item feature on the item side: the columns src/preprocessing.py:47,50-52 emits and the reference model never reads).

There is NO reference counterpart for either: they are tested against a numpy fp64 restatement of the formulas above (test
infrastructure, named in tests/test_gpu_extensions.py) plus closed-form known-answer cases in the same test file.  Everything runs on
this repo's existing kernels -- the multi-table gather fused with the concat (rs_gather_concat), Dense on the fp32 tensor
kernel from 512 rows (rs_dense_forward_x3: the D x D cross products are K = N = 256 GEMMs on tcgen05) or the FFMA kernel
below that, rs_cross_v2_combine for the elementwise rest, rs_heads_forward -- in fp32 mode (tolerance 1e-5 of tensor scale).
"""

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .config import ModelConfig
from .models import Dense, Embedding, Sequential, StringLookup, _dev


class DeepCrossNetworkV2:
    """DCN-v2: matrix cross stack + deep relu MLP, concatenated -- the constructor and call signature of
    models.DeepCrossNetwork (src/models.py:14-55), with cross_kernels[l] (D, D) / cross_biases[l] (D,) in place of the
    rank-1 vectors."""

    def __init__(self, cross_layers: int = 3, deep_layers: List[int] = None, dropout_rate: float = 0.2,
                 l2_reg: float = 1e-5, device=None, **kwargs):
        self.cross_layers = cross_layers
        self.deep_layers = deep_layers or [256, 128, 64]
        self.dropout_rate, self.l2_reg = dropout_rate, l2_reg
        self._device = _dev(device)
        self.cross_nets: List[Dense] = []
        self.deep_nets = [Dense(units, activation="relu", device=self._device) for units in self.deep_layers]
        self.built = False

    def build(self, input_shape):
        D = int(input_shape[-1])
        self.cross_nets = [Dense(D, device=self._device) for _ in range(self.cross_layers)]
        for i, layer in enumerate(self.cross_nets):
            layer.build(D, seed=40 + i)
        self.built = True

    def set_cross_weights(self, kernels: Sequence[np.ndarray], biases: Sequence[np.ndarray]):
        self.cross_layers = len(kernels)
        self.cross_nets = []
        for k, b in zip(kernels, biases):
            layer = Dense(int(np.asarray(k).shape[1]), device=self._device)
            layer.set_weights([k, b])
            self.cross_nets.append(layer)
        self.built = True

    def __call__(self, inputs: torch.Tensor, training=None) -> torch.Tensor:
        if not self.built:
            self.build(inputs.shape)
        B, D = int(inputs.shape[0]), int(inputs.shape[1])
        for i, layer in enumerate(self.deep_nets):
            if layer.kernel is None:
                layer.build(D if i == 0 else self.deep_layers[i - 1], seed=20 + i)
        out = torch.empty((B, D + self.deep_layers[-1]), dtype=torch.float32, device=inputs.device)
        xl = inputs
        n = len(self.cross_nets)
        for i, layer in enumerate(self.cross_nets):
            u = layer(xl)                                                     # x_l W_l + b_l: a D x D GEMM (tcgen05 from 512 rows)
            xl = ops.cross_v2_combine(inputs, u, xl, out=out[:, :D] if i == n - 1 else None)
        if n == 0:
            out[:, :D].copy_(inputs)
        Sequential(self.deep_nets)(inputs, out=out[:, D:])
        return out

    def get_config(self):
        return {"cross_layers": self.cross_layers, "deep_layers": self.deep_layers,
                "dropout_rate": self.dropout_rate, "l2_reg": self.l2_reg}


class ExtendedMultiTaskModel:
    """Two towers with extra categorical tables + DCN (v1 or v2) + the two heads; same output dict as
    models.MultiTaskModel.call.  extra_user_features / extra_item_features: {feature name: (number of codes, embedding
    width)}; the features dict then carries an integer tensor per name next to user_id / movie_id.  A tower's input is
    concat(id embedding, extra embeddings...) -- one rs_gather_concat launch over all its tables."""

    def __init__(self, config: ModelConfig, user_vocab: List[str], item_vocab: List[str],
                 extra_user_features: Optional[Dict[str, Tuple[int, int]]] = None,
                 extra_item_features: Optional[Dict[str, Tuple[int, int]]] = None, cross: str = "v2", device=None):
        if cross not in ("v1", "v2"):
            raise ValueError("cross must be 'v1' (the reference's rank-1 layer) or 'v2' (matrix cross)")
        from .models import DeepCrossNetwork
        self.config, self.device = config, _dev(device)
        dev, E = self.device, config.embedding_dim
        self.user_lookup = StringLookup(vocabulary=user_vocab, mask_token=None, device=dev)
        self.item_lookup = StringLookup(vocabulary=item_vocab, mask_token=None, device=dev)
        self.user_embedding = Embedding(len(user_vocab) + 1, E, device=dev, seed=1)
        self.item_embedding = Embedding(len(item_vocab) + 1, E, device=dev, seed=2)
        self.extra_user = {k: Embedding(n, w, device=dev, seed=50 + i) for i, (k, (n, w)) in enumerate((extra_user_features or {}).items())}
        self.extra_item = {k: Embedding(n, w, device=dev, seed=70 + i) for i, (k, (n, w)) in enumerate((extra_item_features or {}).items())}
        self.user_tower = Sequential([Dense(u, activation="relu", device=dev) for u in config.user_tower_dims] + [Dense(E, device=dev)])
        self.item_tower = Sequential([Dense(u, activation="relu", device=dev) for u in config.item_tower_dims] + [Dense(E, device=dev)])
        for tower, extras, s in ((self.user_tower, self.extra_user, 100), (self.item_tower, self.extra_item, 200)):
            fan_in = E + sum(e.output_dim for e in extras.values())
            for i, layer in enumerate(tower.layers):
                layer.build(fan_in, seed=s + i)
                fan_in = layer.units
        dcn_cls = DeepCrossNetworkV2 if cross == "v2" else DeepCrossNetwork
        self.dcn = dcn_cls(config.cross_layers, config.dnn_dims, config.dropout_rate, config.l2_reg, device=dev)
        self.rating_head = Dense(1, name="rating_pred", device=dev)
        self.ctr_head = Dense(1, name="ctr_pred", device=dev)
        W = 2 * E + config.dnn_dims[-1]
        self.rating_head.build(W, seed=300)
        self.ctr_head.build(W, seed=301)

    def _side(self, lookup, emb, extras, tower, features, key, out):
        rows = lookup(features[key])
        tables = [emb.embeddings] + [e.embeddings for e in extras.values()]
        ids = [rows] + [torch.as_tensor(features[k], device=self.device) for k in extras]
        if len(tables) == 1:
            return tower.call_on_rows(emb, rows, out=out)
        return tower(ops.gather_concat(tables, ids), out=out)

    def __call__(self, data, training=None) -> Dict[str, torch.Tensor]:
        features, _ = data if isinstance(data, tuple) else (data, None)
        E = self.config.embedding_dim
        B = int(self.user_lookup(features["user_id"]).shape[0])
        combined = torch.empty((B, 2 * E), dtype=torch.float32, device=self.device)
        user_emb = self._side(self.user_lookup, self.user_embedding, self.extra_user, self.user_tower, features, "user_id", combined[:, :E])
        item_emb = self._side(self.item_lookup, self.item_embedding, self.extra_item, self.item_tower, features, "movie_id", combined[:, E:])
        dcn_out = self.dcn(combined, training=training)
        hw = torch.stack([self.rating_head.kernel[:, 0], self.ctr_head.kernel[:, 0]]).contiguous()
        hb = torch.cat([self.rating_head.bias, self.ctr_head.bias]).contiguous()
        rating, ctr = ops.heads(dcn_out, hw, hb)
        return {"user_embedding": user_emb, "item_embedding": item_emb, "rating_prediction": rating, "ctr_prediction": ctr}

    # weights as one dict of numpy arrays (the test restatement takes the same dict)
    def get_weights(self) -> Dict[str, Any]:
        g = lambda layers: [tuple(layer.get_weights()) for layer in layers]   # noqa: E731
        w = {"user_embedding": self.user_embedding.get_weights()[0], "item_embedding": self.item_embedding.get_weights()[0],
             "extra_user": {k: e.get_weights()[0] for k, e in self.extra_user.items()},
             "extra_item": {k: e.get_weights()[0] for k, e in self.extra_item.items()},
             "user_tower": g(self.user_tower.layers), "item_tower": g(self.item_tower.layers), "deep": g(self.dcn.deep_nets),
             "rating_head": tuple(self.rating_head.get_weights()), "ctr_head": tuple(self.ctr_head.get_weights())}
        if isinstance(self.dcn, DeepCrossNetworkV2):
            if not self.dcn.built:
                self.dcn.build((0, 2 * self.config.embedding_dim))
            w["cross_v2"] = g(self.dcn.cross_nets)
        else:
            if not self.dcn.built:
                self.dcn.build((0, 2 * self.config.embedding_dim))
            w["cross_w"] = [x.cpu().numpy().reshape(-1) for x in self.dcn.cross_weights]
            w["cross_b"] = [x.cpu().numpy() for x in self.dcn.cross_biases]
        return w
