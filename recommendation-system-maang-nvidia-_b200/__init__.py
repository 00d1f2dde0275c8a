"""recsys_b200 -- B200-native (sm_100a) drop-in for the serving/scoring hot path of
OnlyAhad13/Recommendation-System-MAANG-NVIDIA-: MultiTower + DCN forward and exact two-tower top-K.

The directory is named `recommendation-system-maang-nvidia-_b200` (not an importable identifier); the
repo-root module `recsys_b200.py` loads it under the package name `recsys_b200`.

    import recsys_b200 as rs
    index = rs.IndexFlatIP(128); index.add(items); D, I = index.search(queries, 100)   # faiss-shaped
    model = rs.MultiTaskModel(rs.ModelConfig(), user_vocab, item_vocab, {})            # Keras-shaped
    out = model({"user_id": [...], "movie_id": [...]})
    service = rs.RecommendationService.from_parts(model.encoder, index, user_vocab, item_map)
    service.recommend("42", k=10)
"""

from . import _lib, ops, synth  # noqa: F401
from .config import ModelConfig  # noqa: F401
from .index import IndexFlatIP, normalize_L2, read_index, write_index  # noqa: F401
from .models import (  # noqa: F401
    DeepCrossNetwork,
    Dense,
    Embedding,
    MultiTaskModel,
    MultiTowerModel,
    RankingSession,
    Sequential,
    StringLookup,
    load_weights,
)
from .service import RecommendationService, export_encoder_npz  # noqa: F401
from .serving import MicroBatcher, create_app  # noqa: F401
from .extensions import DeepCrossNetworkV2, ExtendedMultiTaskModel  # noqa: F401  (north-star extensions, no reference counterpart)
from .catalog import build_item_index, evaluate_recall, get_item_embeddings, ranked_item_ids, save_index_artifacts  # noqa: F401

__version__ = "0.1.0"
