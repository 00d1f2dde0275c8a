"""RecommendationService -- the reference's serving API (app/recommendation_service.py:18-110) over the
B200 kernels.  Same constructor, methods, return types and error behaviour, so app/main.py can import it
unchanged: `recommend(user_id, k)` -> [{"item_id", "score", "rank"}], `score(user_id, item_ids)` ->
{item_id: float}, ValueError on an unknown user in score(), the first-k "popular items" cold start for
an unknown user in recommend(), FileNotFoundError from load() when the model directory is missing.

The whole recommend() path stays on the GPU: host dict lookup -> K1 gather -> K2 user tower ->
K4 normalise -> K5 catalog scan + top-k; only the k (score, row) pairs come back to the host.

Artifacts.  The reference saves `encoder.keras` (Keras v3 zip holding an HDF5 weight file),
`vocabs.json`, `item_map.json` and `faiss.idx` (src/trainer.py:228-245).  load() reads all of them as the
reference writes them; the HDF5 member of encoder.keras goes through this repo's own minimal reader
(keras_io.py -- h5py is not available here, and the reader could not be checked against a file written by
h5py itself).  If `encoder_weights.npz` exists (keys: user_embedding, item_embedding,
user_tower_{i}_kernel/bias, item_tower_{i}_kernel/bias; `export_encoder_npz` documents the mapping) it wins.
`from_parts` builds a service from in-memory objects (tests, bench).
"""

import json
import logging
import threading
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

from .config import ModelConfig
from .index import IndexFlatIP, normalize_L2, read_index
from .models import MultiTowerModel, load_weights
from . import ops

logger = logging.getLogger(__name__)


def export_encoder_npz(path, weights: Dict) -> None:
    """Write the encoder half of a weight dict as the .npz side-file load() reads."""
    out = {"user_embedding": weights["user_embedding"], "item_embedding": weights["item_embedding"]}
    for side in ("user", "item"):
        for i, (k, b) in enumerate(weights[f"{side}_tower"]):
            out[f"{side}_tower_{i}_kernel"] = k
            out[f"{side}_tower_{i}_bias"] = b
    np.savez(path, **out)


def _weights_from_npz(z) -> Dict:
    w = {"user_embedding": z["user_embedding"], "item_embedding": z["item_embedding"]}
    for side in ("user", "item"):
        layers = []
        i = 0
        while f"{side}_tower_{i}_kernel" in z:
            layers.append((z[f"{side}_tower_{i}_kernel"], z[f"{side}_tower_{i}_bias"]))
            i += 1
        w[f"{side}_tower"] = layers
    return w


class RecommendationService:
    def __init__(self, model_dir: str = "outputs/models/experiment_001"):
        self.model_dir = Path(model_dir)
        self.version = "1.2.0"
        self.user_vocab: List[str] = []
        self.item_map: Dict[str, str] = {}
        self.encoder_model: Optional[MultiTowerModel] = None
        self.faiss_index: Optional[IndexFlatIP] = None
        self._user_set = frozenset()
        self.single_user_graph = True          # recommend(): one CUDA-graph replay per request (see _single_user)
        self._graphs: Dict[int, dict] = {}
        self._graph_lock = threading.Lock()

    @classmethod
    def from_parts(cls, encoder: MultiTowerModel, index: IndexFlatIP, user_vocab: List[str], item_map: Dict[str, str],
                   model_dir: str = "<memory>") -> "RecommendationService":
        s = cls(model_dir)
        s.encoder_model, s.faiss_index = encoder, index
        # a plain dict is copied; any other mapping (a lazy row -> item-id view of a 10M-item catalog, say) is used as is:
        # recommend() only ever does item_map[str(row)] and len(item_map)
        s.user_vocab, s.item_map = list(user_vocab), (dict(item_map) if isinstance(item_map, dict) else item_map)
        s._user_set = frozenset(s.user_vocab)
        return s

    def load(self):
        logger.info(f"Loading model artifacts from {self.model_dir}...")
        if not self.model_dir.exists():
            raise FileNotFoundError(f"Model directory not found: {self.model_dir}")
        with open(self.model_dir / "vocabs.json", "r") as f:
            vocabs = json.load(f)
        self.user_vocab = vocabs["users"]
        self._user_set = frozenset(self.user_vocab)
        self.faiss_index = read_index(str(self.model_dir / "faiss.idx"))
        with open(self.model_dir / "item_map.json", "r") as f:
            self.item_map = json.load(f)
        cfg_path = self.model_dir / "config.json"
        config = ModelConfig(**json.load(open(cfg_path))) if cfg_path.exists() else ModelConfig()
        item_vocab = vocabs.get("items") or [self.item_map[str(i)] for i in range(len(self.item_map))]
        npz, keras_zip = self.model_dir / "encoder_weights.npz", self.model_dir / "encoder.keras"
        if npz.exists():
            w = _weights_from_npz(np.load(npz))
        elif keras_zip.exists():
            from .keras_io import load_keras_encoder           # the reference's own artifact (src/trainer.py:230)
            w = load_keras_encoder(keras_zip)
        else:
            raise FileNotFoundError(f"neither {keras_zip} nor {npz} found in the model directory")
        config.embedding_dim = int(w["user_embedding"].shape[1])
        config.user_tower_dims = [int(k.shape[1]) for k, _ in w["user_tower"][:-1]]
        config.item_tower_dims = [int(k.shape[1]) for k, _ in w["item_tower"][:-1]]
        self.encoder_model = MultiTowerModel(config, self.user_vocab, item_vocab, {})
        load_weights(self.encoder_model, w)
        logger.info(f"Loaded {len(self.user_vocab)} users, index with {self.faiss_index.ntotal} items.")

    def is_ready(self) -> bool:
        return all([self.encoder_model, self.faiss_index, self.user_vocab, self.item_map])

    def _device_ctx(self):
        """Makes the encoder's GPU current for the calling thread: a threaded HTTP server (or the micro-batcher's worker
        thread, serving.py) calls in from threads whose current device is still 0."""
        import contextlib
        import torch
        dev = getattr(self.encoder_model, "device", None)
        if dev is None or not torch.cuda.is_available():
            return contextlib.nullcontext()
        return torch.cuda.device(dev)

    def _single_user(self, user_id: str, k: int):
        """The /recommendations request (app/recommendation_service.py:64-72 for ONE known user) as one CUDA-graph replay:
        row number up (8 bytes, pinned), gather -> user tower -> normalize_L2 -> catalog search as captured launches, k
        scores + k rows + the search's "not final" flag down in one synchronisation.  Same kernels, same arithmetic and
        the same results as the eager calls below -- it removes their ~ten launches' host work and two of the three
        host synchronisations from the request's latency.  Returns None when it does not apply (no CUDA index, another
        thread is inside it, a flagged query): the caller then takes the eager path."""
        enc, index = self.encoder_model, self.faiss_index
        if not self.single_user_graph or not isinstance(index, IndexFlatIP) or k < 1 or index.ntotal == 0:
            return None
        if not self._graph_lock.acquire(blocking=False):
            return None
        try:
            import torch
            sig = (id(enc), id(index), int(index.ntotal), enc.user_embedding._version,
                   tuple(layer._version for layer in enc.user_tower.layers))
            entry = self._graphs.get(k)
            if entry is None or entry["sig"] != sig:
                dev = enc.device
                rows_host = torch.zeros((1,), dtype=torch.int64).pin_memory()
                rows = torch.zeros((1,), dtype=torch.int64, device=dev)

                def step():
                    q = enc.user_tower.call_on_rows(enc.user_embedding, rows)          # src/models.py:85
                    normalize_L2(q)
                    return index.search_deferred(q, k)

                side = torch.cuda.Stream(device=dev)                # capture off the legacy default stream
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(2):                              # lazy one-time setup (bf16 shadow) outside the capture
                        step()
                    side.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side):
                        D, I, need = step()
                torch.cuda.current_stream(dev).wait_stream(side)
                entry = self._graphs[k] = {
                    "sig": sig, "graph": graph, "rows_host": rows_host, "rows": rows, "D": D, "I": I, "need": need,
                    "D_host": torch.empty((1, k), dtype=torch.float32).pin_memory(),
                    "I_host": torch.empty((1, k), dtype=torch.int64).pin_memory(),
                    "need_host": torch.zeros((1,), dtype=torch.int32).pin_memory()}
            entry["rows_host"][0] = int(enc.user_lookup.lookup_host([user_id])[0])
            entry["rows"].copy_(entry["rows_host"], non_blocking=True)
            entry["graph"].replay()
            entry["D_host"].copy_(entry["D"], non_blocking=True)
            entry["I_host"].copy_(entry["I"], non_blocking=True)
            entry["need_host"].copy_(entry["need"], non_blocking=True)
            torch.cuda.current_stream(enc.device).synchronize()
            if int(entry["need_host"][0]) != 0:
                return None
            return entry["D_host"].numpy().copy(), entry["I_host"].numpy().copy()
        finally:
            self._graph_lock.release()

    def recommend(self, user_id: str, k: int = 10) -> List[Dict]:
        if user_id not in self._user_set:          # dict/set membership instead of the reference's O(V) list scan
            logger.warning(f"User '{user_id}' not in vocabulary. Applying cold-start strategy.")
            return self._get_popular_items(k)
        with self._device_ctx():
            fast = self._single_user(user_id, int(k))
            if fast is not None:
                distances, indices = fast
            else:
                q = self.encoder_model({"user_id": [user_id]}, training=False)["user_embedding"]   # [1, E] on device
                normalize_L2(q)
                distances, indices = self.faiss_index.search(q, k)
                distances = distances.cpu().numpy()        # the only device -> host traffic: k scores + k rows
                indices = indices.cpu().numpy()
        recommendations = []
        for i, idx in enumerate(indices[0]):
            if idx != -1:
                recommendations.append({"item_id": self.item_map[str(idx)], "score": float(distances[0][i]),
                                        "rank": i + 1})
        return recommendations

    def recommend_batch(self, user_ids: List[str], k: int = 10) -> List[Dict]:
        """Batched variant with the reference's return shape (app/model_service.py:200-228): one entry per input id, in
        input order -- {"user_id", "recommendations", "status": "success"} or, when that user's recommendations could
        not be produced, {"user_id", "recommendations": [], "status": "error", "error": message}.  Duplicate ids give
        duplicate entries.  Unknown users get the cold-start list (status "success", as recommend() would return).
        Underneath it is ONE user-tower pass + ONE catalog pass for all known users of the call."""
        known = [u for u in dict.fromkeys(user_ids) if u in self._user_set]
        recs: Dict[str, List[Dict]] = {}
        error: Optional[str] = None
        if known:
            try:
                with self._device_ctx():
                    q = self.encoder_model({"user_id": known}, training=False)["user_embedding"]
                    normalize_L2(q)
                    D, I = self.faiss_index.search(q, k)
                    D, I = D.cpu().numpy(), I.cpu().numpy()
                for r, u in enumerate(known):
                    recs[u] = [{"item_id": self.item_map[str(idx)], "score": float(D[r][i]), "rank": i + 1}
                               for i, idx in enumerate(I[r]) if idx != -1]
            except Exception as e:  # noqa: BLE001 -- the reference catches per user; the batched pass fails as a whole
                error = str(e)
        results = []
        for u in user_ids:
            if u in recs:
                results.append({"user_id": u, "recommendations": [dict(r) for r in recs[u]], "status": "success"})
            elif u in self._user_set:
                results.append({"user_id": u, "recommendations": [], "status": "error", "error": error or "unknown error"})
            else:
                try:
                    results.append({"user_id": u, "recommendations": self._get_popular_items(k), "status": "success"})
                except Exception as e:  # noqa: BLE001
                    results.append({"user_id": u, "recommendations": [], "status": "error", "error": str(e)})
        return results

    def score(self, user_id: str, item_ids: List[str]) -> Dict[str, float]:
        if user_id not in self._user_set:
            raise ValueError(f"User '{user_id}' not found in vocabulary.")
        with self._device_ctx():
            emb = self.encoder_model({"user_id": [user_id], "movie_id": list(item_ids)}, training=False)
            scores = ops.matvec(emb["item_embedding"], emb["user_embedding"][0].contiguous())   # raw, un-normalised
            return {item_id: float(s) for item_id, s in zip(item_ids, scores.cpu().numpy())}

    def _get_popular_items(self, k: int) -> List[Dict]:
        return [{"item_id": self.item_map[str(i)], "score": 1.0 - (i * 0.05), "rank": i + 1}
                for i in range(min(k, len(self.item_map)))]

    def get_model_info(self) -> Dict:
        return {"version": self.version, "model_path": str(self.model_dir), "num_users": len(self.user_vocab),
                "faiss_index_items": self.faiss_index.ntotal if self.faiss_index else 0}
