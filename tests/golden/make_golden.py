"""Generate the committed golden fixtures from the oracle.

    python tests/golden/make_golden.py

The reference's own stack (TensorFlow, faiss) cannot be imported in this image, so these fixtures are
outputs of the CPU restatement in oracle/ (see oracle/__init__.py; the fixtures that come from the
reference's own files are made by make_reference_golden.py).  Inputs and weights are regenerated from seeds by recsys_b200.synth; only ids and
expected outputs are stored, which keeps the files small.  The fixtures (a) freeze the oracle so a later
edit cannot silently change it, and (b) travel to the GPU box, where the CUDA path is checked against them.
"""

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
import recsys_b200 as rs  # noqa: E402
from oracle import flat_ip  # noqa: E402
from oracle.models import MultiTaskModel  # noqa: E402

OUT = Path(__file__).resolve().parent

RANK = dict(n_users=300, n_items=200, batch=96, weight_seed=21, id_seed=22)
RETR = dict(n=20000, d=128, nq=8, k=100, catalog_seed=31, query_seed=32)
RETR64 = dict(n=5000, d=64, nq=5, k=10, catalog_seed=41, query_seed=42)


def ranking_case():
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, RANK["n_users"], RANK["n_items"], seed=RANK["weight_seed"])
    uv, iv = rs.synth.make_vocab(RANK["n_users"]), rs.synth.make_vocab(RANK["n_items"])
    rng = np.random.default_rng(RANK["id_seed"])
    ur = rs.synth.make_ids(rng, RANK["n_users"], RANK["batch"], oov_frac=0.05)
    ir = rs.synth.make_ids(rng, RANK["n_items"], RANK["batch"], oov_frac=0.05)
    o32 = MultiTaskModel(w, uv, iv, np.float32).rank_rows(ur, ir)
    o64 = MultiTaskModel(w, uv, iv, np.float64).rank_rows(ur, ir)
    out = {"user_rows": ur, "item_rows": ir}
    for k_, v in o32.items():
        out[f"f32_{k_}"] = v.astype(np.float32)
    for k_, v in o64.items():
        out[f"f64_{k_}"] = v.astype(np.float64)
    np.savez_compressed(OUT / "ranking_case.npz", **out)


def retrieval_case(spec, name):
    xb = rs.synth.make_unit_rows(spec["n"], spec["d"], spec["catalog_seed"])
    q = rs.synth.make_unit_rows(spec["nq"], spec["d"], spec["query_seed"])
    D, I = flat_ip.search_chain(xb, q, spec["k"])
    D64, I64 = flat_ip.search_fp64(xb, q, spec["k"])
    np.savez_compressed(OUT / name, D=D, I=I, D64=D64, I64=I64, xb_checksum=np.float64(xb.astype(np.float64).sum()),
                        q_checksum=np.float64(q.astype(np.float64).sum()))


if __name__ == "__main__":
    oracle.build_chain_lib()
    ranking_case()
    retrieval_case(RETR, "retrieval_d128.npz")
    retrieval_case(RETR64, "retrieval_d64.npz")
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))
