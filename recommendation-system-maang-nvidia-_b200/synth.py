"""Deterministic synthetic weights and inputs of the BASELINE shapes (numpy only; no GPU, no oracle).

The reference ships no trained artifacts, so tests and the bench use random-init
weights drawn the way Keras initialises them: Embedding ~ U(-0.05, 0.05), Dense / cross kernels
glorot-uniform, and small non-zero biases (Keras uses zeros; non-zero exercises the bias paths).

The weight dict is the interchange layout used by `load_weights` in models.py and by the oracle:
  user_embedding [Vu+1, E], item_embedding [Vi+1, E]        (row 0 = the StringLookup OOV bucket)
  user_tower / item_tower / deep : list of (kernel (in, out), bias (out,))
  cross_w : list of (D, 1);  cross_b : list of (D,)
  rating_head / ctr_head : (kernel (D + deep[-1], 1), bias (1,))
"""

from typing import Dict, List

import numpy as np

from .config import ModelConfig


def glorot_uniform(rng, fan_in: int, fan_out: int) -> np.ndarray:
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


def _dense_stack(rng, dims: List[int]):
    out = []
    for i in range(len(dims) - 1):
        out.append((glorot_uniform(rng, dims[i], dims[i + 1]),
                    rng.normal(0.0, 0.01, size=(dims[i + 1],)).astype(np.float32)))
    return out


def make_weights(config: ModelConfig, n_users: int, n_items: int, seed: int = 10) -> Dict:
    rng = np.random.default_rng(seed)
    E = config.embedding_dim
    D = 2 * E
    w = {
        "user_embedding": rng.uniform(-0.05, 0.05, size=(n_users + 1, E)).astype(np.float32),
        "item_embedding": rng.uniform(-0.05, 0.05, size=(n_items + 1, E)).astype(np.float32),
        "user_tower": _dense_stack(rng, [E] + list(config.user_tower_dims) + [E]),
        "item_tower": _dense_stack(rng, [E] + list(config.item_tower_dims) + [E]),
        "cross_w": [glorot_uniform(rng, D, 1) for _ in range(config.cross_layers)],
        "cross_b": [rng.normal(0.0, 0.01, size=(D,)).astype(np.float32) for _ in range(config.cross_layers)],
        "deep": _dense_stack(rng, [D] + list(config.dnn_dims)),
    }
    W = D + config.dnn_dims[-1]
    w["rating_head"] = (glorot_uniform(rng, W, 1), rng.normal(0.0, 0.01, size=(1,)).astype(np.float32))
    w["ctr_head"] = (glorot_uniform(rng, W, 1), rng.normal(0.0, 0.01, size=(1,)).astype(np.float32))
    return w


def make_vocab(n: int, prefix: str = "") -> List[str]:
    """Vocabulary the way the reference builds it: sorted *strings* (src/trainer.py:81-82)."""
    return sorted(f"{prefix}{i}" for i in range(1, n + 1))


def make_ids(rng, n_rows: int, batch: int, oov_frac: float = 0.01) -> np.ndarray:
    """Row numbers in [1, n_rows] with a fraction of OOV (row 0), int64 -- i.e. StringLookup output."""
    ids = rng.integers(1, n_rows + 1, size=batch, dtype=np.int64)
    ids[rng.random(batch) < oov_frac] = 0
    return ids


def make_unit_rows(n: int, d: int, seed: int, block: int = 1 << 20) -> np.ndarray:
    """Catalog / query rows ~ N(0,1), then scaled to unit length in fp64 and rounded to fp32."""
    out = np.empty((n, d), dtype=np.float32)
    for b0 in range(0, n, block):
        b1 = min(n, b0 + block)
        rng = np.random.default_rng(seed + b0 // block)
        x = rng.standard_normal((b1 - b0, d))
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        out[b0:b1] = x.astype(np.float32)
    return out


def hash_table_rows(rows: np.ndarray, dim: int, seed: int) -> np.ndarray:
    """Host replay of rs_hash_table_fill: T[r, c] = (u32(mix(seed, r*dim + c)) * 2^-32 - 0.5) * 0.1, bit for bit."""
    with np.errstate(over="ignore"):
        g = rows.astype(np.uint64)[:, None] * np.uint64(dim) + np.arange(dim, dtype=np.uint64)[None, :]
        x = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + g
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xFF51AFD7ED558CCD)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xC4CEB9FE1A85EC53)
        x ^= x >> np.uint64(33)
    u = (x & np.uint64(0xFFFFFFFF)).astype(np.uint32).astype(np.float32)
    return ((u * np.float32(2.3283064365386963e-10) - np.float32(0.5)) * np.float32(0.1)).astype(np.float32)


def zipf_ids(rng, n_rows: int, batch: int, alpha: float = 1.05, perm_seed: int = 4001) -> np.ndarray:
    """Bounded-Zipf ranks (exponent alpha) mapped to rows through a fixed random permutation, int64."""
    ranks = np.arange(1, n_rows + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** -alpha)
    cdf /= cdf[-1]
    r = np.searchsorted(cdf, rng.random(batch)).astype(np.int64)
    perm = np.random.default_rng(perm_seed).permutation(n_rows)
    return perm[np.minimum(r, n_rows - 1)].astype(np.int64)
