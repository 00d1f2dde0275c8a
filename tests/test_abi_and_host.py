"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the product fails loudly
without a GPU, the oracle reproduces the committed golden fixtures, and host-side logic works."""

import ctypes
import re
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(GOLD))


def test_library_exports_every_symbol_in_the_header(rs):
    header = (ROOT / "include" / "recsys_b200.h").read_text()
    declared = set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = rs._lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/recsys_b200.h but not exported"
    assert declared == set(rs._lib.EXPORTED)
    assert lib.rs_version() == 1


def test_no_gpu_means_loud_failure_not_fallback(rs):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="librecsys_b200"):
        rs.IndexFlatIP(128)
    with pytest.raises(RuntimeError):
        rs.MultiTowerModel(rs.ModelConfig(), ["1"], ["2"], {})
    x = np.ones((2, 128), np.float32)
    with pytest.raises(RuntimeError):
        rs.normalize_L2(x)
    np.testing.assert_array_equal(x, 1.0)      # and nothing was computed on the host


def test_product_never_imports_the_oracle():
    pkg = ROOT / "recommendation-system-maang-nvidia-_b200"
    for p in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if p.is_file() and p.suffix in (".py", ".cu", ".cuh"):
            text = p.read_text()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), p
            assert "oracle/" not in text.replace("oracle/chain.c)", "") or p.suffix != ".py", p


def test_dlpack_marshalling_of_views(rs):
    a = np.arange(24, dtype=np.float32).reshape(4, 6)
    b = rs._lib.dl(a)
    t = b.ptr.contents
    assert (t.ndim, t.shape[0], t.shape[1], t.dtype.code, t.dtype.bits, t.device.device_type) == (2, 4, 6, 2, 32, 1)
    v = torch.arange(24, dtype=torch.float32).reshape(4, 6)[:, 2:4]
    bv = rs._lib.dl(v)          # keep the borrow alive while its DLTensor is read
    t = bv.ptr.contents
    assert t.shape[1] == 2 and t.strides[0] == 6 and t.strides[1] == 1
    data = ctypes.cast(t.data + t.byte_offset, ctypes.POINTER(ctypes.c_float))
    assert data[0] == 2.0
    ids = torch.arange(5, dtype=torch.int64)
    bi = rs._lib.dl(ids)
    t = bi.ptr.contents
    assert (t.dtype.code, t.dtype.bits) == (0, 64)


def test_string_lookup_and_config(rs):
    cfg = rs.ModelConfig()
    assert (cfg.embedding_dim, cfg.user_tower_dims, cfg.dnn_dims, cfg.cross_layers) == (128, [256, 128, 64], [256, 128], 3)
    assert rs.ModelConfig(**cfg.to_dict()) == cfg
    from recsys_b200.models import StringLookup
    sl = StringLookup.__new__(StringLookup)
    sl.vocabulary = ["1", "10", "2"]
    sl._index = {s: i + 1 for i, s in enumerate(sl.vocabulary)}
    np.testing.assert_array_equal(sl.lookup_host(["2", b"10", "zz", 1]), [3, 2, 0, 1])


def test_index_file_format_roundtrip_layout(rs, tmp_path):
    # writer/reader agree on the byte layout without touching the GPU (the GPU leg is in test_gpu_parity)
    import struct
    d, n = 32, 3
    xb = np.arange(n * d, dtype=np.float32).reshape(n, d)
    p = tmp_path / "faiss.idx"
    with open(p, "wb") as f:
        f.write(b"IxFI" + struct.pack("<i", d) + struct.pack("<q", n) + struct.pack("<qq", 1 << 20, 1 << 20)
                + struct.pack("<B", 1) + struct.pack("<i", 0) + struct.pack("<Q", n * d) + xb.tobytes())
    raw = p.read_bytes()
    assert len(raw) == 4 + 4 + 8 + 16 + 1 + 4 + 8 + n * d * 4
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            rs.read_index(str(p))
    bad = tmp_path / "bad.idx"
    bad.write_bytes(b"IxF2" + raw[4:])
    with pytest.raises(ValueError):
        rs.read_index(str(bad))


def test_oracle_reproduces_golden_fixtures(rs, oracle_mod):
    import make_golden as mg
    from oracle import flat_ip
    from oracle.models import MultiTaskModel
    g = np.load(GOLD / "ranking_case.npz")
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, mg.RANK["n_users"], mg.RANK["n_items"], seed=mg.RANK["weight_seed"])
    uv, iv = rs.synth.make_vocab(mg.RANK["n_users"]), rs.synth.make_vocab(mg.RANK["n_items"])
    out = MultiTaskModel(w, uv, iv, np.float64).rank_rows(g["user_rows"], g["item_rows"])
    for k_, v in out.items():
        np.testing.assert_allclose(v, g[f"f64_{k_}"], rtol=1e-12, atol=1e-14)
    for spec, name in ((mg.RETR, "retrieval_d128.npz"), (mg.RETR64, "retrieval_d64.npz")):
        g = np.load(GOLD / name)
        xb = rs.synth.make_unit_rows(spec["n"], spec["d"], spec["catalog_seed"])
        q = rs.synth.make_unit_rows(spec["nq"], spec["d"], spec["query_seed"])
        assert float(xb.astype(np.float64).sum()) == float(g["xb_checksum"])
        D, I = flat_ip.search_chain(xb, q, spec["k"])
        np.testing.assert_array_equal(I, g["I"])
        np.testing.assert_array_equal(D, g["D"])


def test_metrics_consume_recommendations_unchanged(oracle_mod):
    from oracle import evaluation as ev
    recs = [[{"item_id": "7", "score": 0.9, "rank": 1}, {"item_id": "3", "score": 0.5, "rank": 2}],
            [{"item_id": "1", "score": 0.8, "rank": 1}, {"item_id": "2", "score": 0.1, "rank": 2}]]
    preds = [[r["item_id"] for r in user] for user in recs]
    truth = ["3", "9"]
    assert ev.recall_at_k(preds, truth, 2) == 0.5
    assert ev.precision_at_k(preds, truth, 2) == 0.25
    assert ev.map_at_k(preds, truth, 2) == 0.25
    assert ev.mrr(preds, truth) == 0.25
    assert abs(ev.ndcg_at_k(preds, truth, 2) - 0.5 / np.log2(3)) < 1e-12


def test_x3_shape_support_is_host_logic(rs):
    """rs_dense_x3_supported needs no GPU: the fp32 tensor kernel serves in / out widths of 64, 128, 256 with no or relu
    activation; everything else stays on the FFMA kernel."""
    ok = rs.ops.dense_x3_supported
    assert ok(128, 256, "relu") and ok(256, 128, None) and ok(64, 64, "relu") and ok(256, 256, None)
    assert not ok(100, 128, None) and not ok(128, 512, None) and not ok(128, 128, "sigmoid") and not ok(32, 128, None)
