"""Catalog build and offline evaluation on the GPU path.

Mirrors the trainer-side helpers that sit next to the serving path:
  ProductionTrainer._get_item_embeddings   src/trainer.py:221-226   item tower over the vocabulary in 512-row batches
  ProductionTrainer._build_faiss           src/trainer.py:236-245   normalize_L2 -> IndexFlatIP.add -> item_map
  ProductionTrainer._evaluate              src/trainer.py:195-219   recall@k of the true item among the top-k by raw dot

Everything numeric runs through this repo's kernels (tower forward, K4 normalise, K5 search); the reference's
`np.dot` + `argpartition` evaluation becomes one exact top-k search over an un-normalised index.
"""

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .index import IndexFlatIP, normalize_L2, write_index
from .models import MultiTowerModel


def get_item_embeddings(encoder: MultiTowerModel, item_vocab: Sequence[str], batch: int = 512) -> torch.Tensor:
    """Item tower over the whole vocabulary -> float32 [len(item_vocab), E] on the device (src/trainer.py:221-226)."""
    outs = []
    for i in range(0, len(item_vocab), batch):
        outs.append(encoder({"movie_id": list(item_vocab[i:i + batch])}, training=False)["item_embedding"])
    return torch.cat(outs).contiguous() if outs else torch.empty((0, encoder.config.embedding_dim), device=encoder.device)


def build_item_index(encoder: MultiTowerModel, item_vocab: Sequence[str], normalize: bool = True,
                     batch: int = 512) -> Tuple[IndexFlatIP, Dict[str, str]]:
    """-> (index, item_map) exactly as _build_faiss leaves them: row i of the index is item_vocab[i]."""
    embs = get_item_embeddings(encoder, item_vocab, batch)
    if normalize:
        normalize_L2(embs)
    index = IndexFlatIP(int(embs.shape[1]))
    if embs.shape[0]:
        index.add(embs)
    return index, {str(i): item for i, item in enumerate(item_vocab)}


def save_index_artifacts(index: IndexFlatIP, item_map: Dict[str, str], out_dir) -> None:
    """faiss.idx + item_map.json as src/trainer.py:243-245 writes them."""
    import json
    from pathlib import Path
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    write_index(index, str(out / "faiss.idx"))
    with open(out / "item_map.json", "w") as f:
        json.dump(item_map, f)


def evaluate_recall(encoder: MultiTowerModel, item_vocab: Sequence[str], user_ids: Sequence[str],
                    true_items: Sequence[str], ks: Sequence[int] = (5, 10, 20, 50)) -> Dict[str, float]:
    """recall@k over (user, true item) pairs, scores = raw dot of the two tower outputs (src/trainer.py:201-213)."""
    index, _ = build_item_index(encoder, item_vocab, normalize=False)
    q = encoder({"user_id": list(user_ids)}, training=False)["user_embedding"].contiguous()
    kmax = min(max(ks), max(1, len(item_vocab)))
    _, I = index.search(q, kmax)
    I = I.cpu().numpy()
    pos = {item: i for i, item in enumerate(item_vocab)}
    truth = np.array([pos.get(t, -2) for t in true_items], dtype=np.int64)
    out = {}
    for k in ks:
        kk = min(k, kmax)
        out[f"recall@{k}"] = float(np.mean((I[:, :kk] == truth[:, None]).any(axis=1))) if len(truth) else 0.0
    return out


def ranked_item_ids(index: IndexFlatIP, item_map: Dict[str, str], queries, k: int) -> List[List[str]]:
    """Top-k as ranked item-id strings per query: the List[List[str]] that AdvancedMetrics consumes (src/evaluation.py:25-81)."""
    _, I = index.search(queries, k)
    I = I.cpu().numpy() if isinstance(I, torch.Tensor) else I
    return [[item_map[str(int(j))] for j in row if j >= 0] for row in I]
