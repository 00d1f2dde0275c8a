"""ModelConfig -- same field names and defaults as the reference's src/config.py:9-61, so a
`config.json` written by the reference trainer (src/trainer.py:233-234) loads unchanged.  Only the
shape fields matter on the serving/scoring path; the training fields are carried, not used."""

from dataclasses import asdict, dataclass
from typing import List


@dataclass
class ModelConfig:
    embedding_dim: int = 128
    user_tower_dims: List[int] = None
    item_tower_dims: List[int] = None

    cross_layers: int = 3
    dnn_dims: List[int] = None
    dropout_rate: float = 0.2
    l2_reg: float = 1e-4

    batch_size: int = 2048
    learning_rate_retrieval: float = 0.001
    learning_rate_ranking: float = 0.0001
    epochs_retrieval: int = 20
    epochs_ranking: int = 5
    warmup_steps: int = 1000

    num_hard_negatives: int = 5
    num_random_negatives: int = 10
    negative_sampling_strategy: str = "mixed"

    retrieval_weight: float = 1.0
    ctr_weight: float = 2.0
    rating_weight: float = 0.2

    eval_topk: List[int] = None

    mixed_precision: bool = True
    distributed_strategy: str = "none"

    def __post_init__(self):
        if self.user_tower_dims is None:
            self.user_tower_dims = [256, 128, 64]
        if self.item_tower_dims is None:
            self.item_tower_dims = [256, 128, 64]
        if self.dnn_dims is None:
            self.dnn_dims = [256, 128]
        if self.eval_topk is None:
            self.eval_topk = [5, 10, 20, 50]

    def to_dict(self):
        return asdict(self)
