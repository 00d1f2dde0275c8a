"""Restatement of the top-K consumer, AdvancedMetrics (src/evaluation.py:25-104).  TEST INFRASTRUCTURE ONLY.

The reference module imports keras at module scope, so it only loads here over the stand-ins of tests/golden (which is how
tests/test_reference_run.py holds this copy to the reference's own values); the hand-off test (the drop-in's recommend()
output feeds these functions unchanged) uses this copy of the signatures:
predictions: List[List[str]] ranked item ids per user; ground_truth: List[str] one true item per user.
"""

from typing import List

import numpy as np


def recall_at_k(predictions: List[List[str]], ground_truth: List[str], k: int) -> float:      # :25-32
    r = [1.0 if t in p[:k] else 0.0 for p, t in zip(predictions, ground_truth)]
    return float(np.mean(r)) if r else 0.0


def precision_at_k(predictions: List[List[str]], ground_truth: List[str], k: int) -> float:   # :34-41
    r = [1.0 / len(p[:k]) if t in p[:k] else 0.0 for p, t in zip(predictions, ground_truth)]
    return float(np.mean(r)) if r else 0.0


def ndcg_at_k(predictions: List[List[str]], ground_truth: List[str], k: int) -> float:        # :43-56
    out = []
    for p, t in zip(predictions, ground_truth):
        top = p[:k]
        out.append((1.0 / np.log2(top.index(t) + 2)) / (1.0 / np.log2(2)) if t in top else 0.0)
    return float(np.mean(out)) if out else 0.0


def map_at_k(predictions: List[List[str]], ground_truth: List[str], k: int) -> float:         # :58-69
    out = [1.0 / (p[:k].index(t) + 1) if t in p[:k] else 0.0 for p, t in zip(predictions, ground_truth)]
    return float(np.mean(out)) if out else 0.0


def mrr(predictions: List[List[str]], ground_truth: List[str]) -> float:                      # :71-81
    out = [1.0 / (p.index(t) + 1) if t in p else 0.0 for p, t in zip(predictions, ground_truth)]
    return float(np.mean(out)) if out else 0.0


def coverage(recommendations: List[List[str]], all_items: List[str]) -> float:               # :83-91
    if not all_items:
        return 0.0
    seen = set()
    for rec in recommendations:
        seen.update(rec)
    return len(seen) / len(all_items)


def diversity(recommendations: List[List[str]]) -> float:                                    # :93-104
    if not recommendations:
        return 0.0
    return float(np.mean([len(set(rec)) / len(rec) if len(rec) > 1 else 0.0 for rec in recommendations]))
