"""Service half of the oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates app/recommendation_service.py:64-102 on top of the oracle encoder and
the oracle flat index: membership test -> user tower -> normalize_L2 -> search ->
dict list (recommend), both towers -> raw matvec (score), first-k cold start.
"""

from typing import Dict, List

import numpy as np

from . import flat_ip


class RecommendationService:
    version = "1.2.0"                                             # app/recommendation_service.py:26

    def __init__(self, encoder, index: flat_ip.IndexFlatIP, user_vocab: List[str], item_map: Dict[str, str]):
        self.encoder_model = encoder
        self.faiss_index = index
        self.user_vocab = list(user_vocab)
        self.item_map = dict(item_map)

    def is_ready(self) -> bool:                                   # :61-62
        return all([self.encoder_model, self.faiss_index, self.user_vocab, self.item_map])

    def recommend(self, user_id: str, k: int = 10) -> List[Dict]:
        if user_id not in self.user_vocab:                        # :65
            return self._get_popular_items(k)                     # :67
        emb = self.encoder_model({"user_id": [user_id]}, training=False)["user_embedding"]  # :68-69
        q = np.ascontiguousarray(emb, dtype=np.float32)           # :70
        flat_ip.normalize_L2(q)                                   # :71
        distances, indices = self.faiss_index.search(q, k)        # :72
        recs = []
        for i, idx in enumerate(indices[0]):                      # :74
            if idx != -1:                                         # :75
                recs.append({"item_id": self.item_map[str(idx)], "score": float(distances[0][i]), "rank": i + 1})
        return recs

    def score(self, user_id: str, item_ids: List[str]) -> Dict[str, float]:
        if user_id not in self.user_vocab:                        # :84-85
            raise ValueError(f"User '{user_id}' not found in vocabulary.")
        emb = self.encoder_model({"user_id": [user_id], "movie_id": list(item_ids)}, training=False)  # :86-88
        scores = flat_ip.matvec(emb["item_embedding"], emb["user_embedding"][0])   # :91  raw, un-normalised
        return {item_id: float(s) for item_id, s in zip(item_ids, scores)}         # :92

    def _get_popular_items(self, k: int) -> List[Dict]:           # :94-102
        return [{"item_id": self.item_map[str(i)], "score": 1.0 - (i * 0.05), "rank": i + 1}
                for i in range(min(k, len(self.item_map)))]

    def get_model_info(self) -> Dict:                             # :104-110
        return {"version": self.version, "model_path": "<oracle>", "num_users": len(self.user_vocab),
                "faiss_index_items": self.faiss_index.ntotal if self.faiss_index else 0}
