"""Known-answer tests that pin the oracle.  CPU only.

The reference ships no golden vectors and TF/FAISS cannot run here, so the restatement is pinned by
cases whose answers can be worked out by hand from the reference source.
"""

import numpy as np
import pytest

import oracle
from oracle import flat_ip
from oracle.models import DeepCrossNetwork, MultiTaskModel, MultiTowerModel, dense, make_vocab_index, string_lookup


def test_dcn_all_ones_doubles_each_layer():
    # x0 = 1, w_l = 1/D, b_l = 0  =>  s_l = mean(x_l), x_{l+1} = x0 * mean(x_l) + x_l  =>  x_l = 2^l exactly
    D = 256
    dcn = DeepCrossNetwork([np.full((D, 1), 1.0 / D, np.float32)] * 3, [np.zeros(D, np.float32)] * 3, deep=[])
    x0 = np.ones((5, D), np.float32)
    out = dcn(x0)
    assert out.shape == (5, 2 * D)      # zero deep layers: deep_out stays `inputs` (src/models.py:46-50)
    np.testing.assert_array_equal(out[:, :D], np.full((5, D), 8.0, np.float32))
    np.testing.assert_array_equal(out[:, D:], x0)


def test_dcn_zero_cross_layers_and_identity_deep():
    D = 128
    dcn = DeepCrossNetwork([], [], deep=[(np.eye(D, dtype=np.float32), np.zeros(D, np.float32))])
    x0 = np.abs(np.random.default_rng(0).standard_normal((3, D)).astype(np.float32))
    out = dcn(x0)
    np.testing.assert_array_equal(out[:, :D], x0)       # zero cross layers: x_L = x0
    np.testing.assert_array_equal(out[:, D:], x0)       # identity Dense + relu on non-negative inputs


def test_dcn_bias_sits_outside_the_product():
    # one layer, w = 0  =>  s = 0  =>  x_1 = x0 * 0 + b + x0 = x0 + b   (bias NOT multiplied by x0)
    D = 128
    b = np.arange(D, dtype=np.float32)
    dcn = DeepCrossNetwork([np.zeros((D, 1), np.float32)], [b], deep=[(np.zeros((D, 4), np.float32), np.zeros(4, np.float32))])
    x0 = np.full((2, D), 3.0, np.float32)
    out = dcn(x0)
    np.testing.assert_array_equal(out[:, :D], x0 + b)
    np.testing.assert_array_equal(out[:, D:], np.zeros((2, 4), np.float32))


def test_dcn_cross_term_is_a_scalar_per_row():
    rng = np.random.default_rng(1)
    D = 256
    w = rng.standard_normal((D, 1)).astype(np.float32)
    b = rng.standard_normal(D).astype(np.float32)
    x0 = rng.standard_normal((7, D)).astype(np.float32)
    dcn = DeepCrossNetwork([w], [b], deep=[(np.zeros((D, 1), np.float32), np.zeros(1, np.float32))], dtype=np.float64)
    out = dcn(x0)[:, :D]
    s = x0.astype(np.float64) @ w.astype(np.float64)        # [7, 1]
    np.testing.assert_allclose(out, x0 * s + b + x0, rtol=1e-12)


def test_string_lookup_semantics():
    vocab = ["1", "10", "2"]                                   # sorted as *strings* (src/trainer.py:81-82)
    idx = make_vocab_index(vocab)
    np.testing.assert_array_equal(string_lookup(idx, ["1", "2", "10", "nope", ""]), [1, 3, 2, 0, 0])
    assert string_lookup(idx, []).dtype == np.int64


def test_identity_towers_return_embedding_rows():
    E = 8
    eye = [(np.eye(E, dtype=np.float32), np.zeros(E, np.float32))] * 4
    table = np.abs(np.random.default_rng(2).standard_normal((4, E)).astype(np.float32))
    w = {"user_embedding": table, "item_embedding": table * 2, "user_tower": eye, "item_tower": eye}
    m = MultiTowerModel(w, ["a", "b", "c"], ["x", "y", "z"])
    out = m({"user_id": ["b", "zzz"]})
    assert out["item_embedding"] is None                      # missing side -> None (src/models.py:80-91)
    np.testing.assert_array_equal(out["user_embedding"], table[[2, 0]])      # OOV -> row 0
    out = m({"movie_id": ["z"]})
    assert out["user_embedding"] is None
    np.testing.assert_array_equal(out["item_embedding"], 2 * table[[3]])


def test_last_tower_layer_is_linear():
    E = 4
    neg = [(-np.eye(E, dtype=np.float32), np.zeros(E, np.float32))]
    table = np.ones((2, E), np.float32)
    w = {"user_embedding": table, "item_embedding": table, "user_tower": neg, "item_tower": neg + neg}
    m = MultiTowerModel(w, ["a"], ["x"])
    np.testing.assert_array_equal(m({"user_id": ["a"]})["user_embedding"], -table[[1]])   # single layer = linear
    np.testing.assert_array_equal(m({"movie_id": ["x"]})["item_embedding"], np.zeros((1, E)))  # relu(-1)=0 then linear


def test_multitask_heads_and_keys():
    import recsys_b200 as rs
    cfg = rs.ModelConfig(embedding_dim=32, user_tower_dims=[16], item_tower_dims=[16], dnn_dims=[8], cross_layers=2)
    w = rs.synth.make_weights(cfg, 5, 6, seed=3)
    uv, iv = rs.synth.make_vocab(5), rs.synth.make_vocab(6)
    m = MultiTaskModel(w, uv, iv, dtype=np.float64)
    feats = {"user_id": ["1", "3"], "movie_id": ["2", "nope"]}
    out = m((feats, {"rating": None}))                        # tuple input accepted (src/models.py:126)
    assert set(out) == {"user_embedding", "item_embedding", "rating_prediction", "ctr_prediction"}
    assert out["rating_prediction"].shape == (2, 1) and out["ctr_prediction"].shape == (2, 1)
    z = m.dcn(np.concatenate([out["user_embedding"], out["item_embedding"]], axis=1))
    assert z.shape == (2, 64 + 8)
    np.testing.assert_allclose(out["rating_prediction"], z @ w["rating_head"][0] + w["rating_head"][1], rtol=1e-12)
    np.testing.assert_allclose(out["ctr_prediction"], 1 / (1 + np.exp(-(z @ w["ctr_head"][0] + w["ctr_head"][1]))), rtol=1e-12)
    assert np.all((out["ctr_prediction"] > 0) & (out["ctr_prediction"] < 1))


def test_fp32_and_fp64_oracles_agree():
    import recsys_b200 as rs
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, 50, 40, seed=4)
    uv, iv = rs.synth.make_vocab(50), rs.synth.make_vocab(40)
    rng = np.random.default_rng(5)
    ur, ir = rs.synth.make_ids(rng, 50, 256), rs.synth.make_ids(rng, 40, 256)
    o32 = MultiTaskModel(w, uv, iv, np.float32).rank_rows(ur, ir)
    o64 = MultiTaskModel(w, uv, iv, np.float64).rank_rows(ur, ir)
    for key in o32:
        scale = np.abs(o64[key]).max()
        assert np.abs(o32[key] - o64[key]).max() <= 2e-6 * scale + 1e-9, key


# ---------------------------------------------------------------- retrieval ------------------------
def test_one_hot_catalog_returns_largest_query_coordinates():
    d = 128
    xb = np.eye(d, dtype=np.float32)
    q = np.random.default_rng(6).permutation(d).astype(np.float32)[None, :]
    D, I = flat_ip.search_chain(xb, q, 10)
    np.testing.assert_array_equal(I[0], np.argsort(-q[0])[:10])
    np.testing.assert_array_equal(D[0], np.sort(q[0])[::-1][:10])


def test_duplicate_rows_tie_break_is_index_ascending():
    rng = np.random.default_rng(7)
    base = rng.standard_normal((10, 64)).astype(np.float32)
    xb = np.ascontiguousarray(np.concatenate([base, base, base])[rng.permutation(30)])
    q = rng.standard_normal((3, 64)).astype(np.float32)
    D, I = flat_ip.search_chain(xb, q, 30)
    for r in range(3):
        for j in range(29):
            assert D[r, j] > D[r, j + 1] or (D[r, j] == D[r, j + 1] and I[r, j] < I[r, j + 1])
    assert sorted(I[0].tolist()) == list(range(30))


def test_k_larger_than_n_pads_with_minus_one():
    xb = np.random.default_rng(8).standard_normal((3, 32)).astype(np.float32)
    q = xb[:1].copy()
    D, I = flat_ip.search_chain(xb, q, 5)
    assert I[0, 3] == -1 and I[0, 4] == -1 and set(I[0, :3].tolist()) == {0, 1, 2}
    assert D[0, 3] == flat_ip.NEG_PAD and D[0, 4] == flat_ip.NEG_PAD
    D, I = flat_ip.search_chain(np.empty((0, 32), np.float32), q, 2)
    assert (I == -1).all()


def test_normalize_l2_in_place_and_zero_rows():
    x = np.array([[3.0, 4.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], [1.0, 1.0, 1.0, 1.0]], np.float32)
    assert flat_ip.normalize_L2(x) is None
    np.testing.assert_array_equal(x[0], np.array([0.6, 0.8, 0, 0], np.float32))
    np.testing.assert_array_equal(x[1], 0)
    np.testing.assert_array_equal(x[2], 0.5)


def test_chain_matches_fp64_and_blas_variants():
    import recsys_b200 as rs
    xb = rs.synth.make_unit_rows(20000, 128, 11)
    q = rs.synth.make_unit_rows(16, 128, 12)
    D, I = flat_ip.search_chain(xb, q, 100)
    D64, I64 = flat_ip.search_fp64(xb, q, 100)
    Db, Ib = flat_ip.search_blas(xb, q, 100)
    assert np.abs(D - D64).max() < 1e-6
    # index sets agree except where fp32 reassociation flips a near-tie (gap below 1e-6)
    for r in range(16):
        for other in (I64[r], Ib[r]):
            diff = set(I[r]) ^ set(other)
            assert len(diff) <= 2
    S = flat_ip.scores_chain(xb, q)
    assert S[3, I[3, 0]] == D[3, 0]
    assert oracle.load_chain_lib().oracle_dot_chain is not None


def test_merge_topk_equals_search_over_concatenation():
    import recsys_b200 as rs
    xb = rs.synth.make_unit_rows(3000, 64, 13)
    q = rs.synth.make_unit_rows(5, 64, 14)
    D, I = flat_ip.search_chain(xb, q, 20)
    parts = [flat_ip.search_chain(np.ascontiguousarray(xb[a:b]), q, 20, id_base=a) for a, b in ((0, 1000), (1000, 1007), (1007, 3000))]
    Dm, Im = flat_ip.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    np.testing.assert_array_equal(Im, I)
    np.testing.assert_array_equal(Dm, D)


def test_service_sequence():
    import recsys_b200 as rs
    cfg = rs.ModelConfig(embedding_dim=32, user_tower_dims=[16], item_tower_dims=[16])
    uv, iv = rs.synth.make_vocab(20), rs.synth.make_vocab(30)
    w = rs.synth.make_weights(cfg, 20, 30, seed=15)
    enc = MultiTowerModel(w, uv, iv)
    embs = np.ascontiguousarray(enc({"movie_id": iv})["item_embedding"], dtype=np.float32)
    flat_ip.normalize_L2(embs)
    index = flat_ip.IndexFlatIP(32)
    index.add(embs)
    item_map = {str(i): it for i, it in enumerate(iv)}
    svc = oracle.RecommendationService(enc, index, uv, item_map)
    recs = svc.recommend(uv[3], 7)
    assert [r["rank"] for r in recs] == list(range(1, 8))
    assert all(-1.0001 <= r["score"] <= 1.0001 for r in recs)            # cosine: both sides normalised
    assert recs == sorted(recs, key=lambda r: -r["score"])
    cold = svc.recommend("unknown-user", 3)
    assert [c["item_id"] for c in cold] == [iv[0], iv[1], iv[2]] and cold[1]["score"] == 0.95
    sc = svc.score(uv[3], [iv[0], "not-an-item"])
    assert set(sc) == {iv[0], "not-an-item"}                             # unknown item scored via OOV row 0
    with pytest.raises(ValueError):
        svc.score("unknown-user", [iv[0]])


def test_extensions_restatement_known_answers():
    """oracle/extensions.py (the checker of the north-star extensions, which have no reference counterpart): DCN-v2 with
    W = c I, b = 0 on x0 = 1 gives x_l = (1 + c)^l; with W = 0 the bias sits INSIDE the product (x_L = x0 (1 + L b)); extra
    tables are concatenated behind the id embedding; out-of-range codes read as zeros."""
    from oracle import extensions as oext
    D, E = 8, 4
    eye_tower = [(np.eye(E), np.zeros(E))]
    w = {"user_embedding": np.ones((3, E)), "item_embedding": np.ones((3, E)), "extra_user": {}, "extra_item": {},
         "user_tower": eye_tower, "item_tower": eye_tower, "deep": [(np.zeros((D, 2)), np.zeros(2))],
         "rating_head": (np.ones((D + 2, 1)), np.zeros(1)), "ctr_head": (np.zeros((D + 2, 1)), np.zeros(1)),
         "cross_v2": [(0.5 * np.eye(D), np.zeros(D))] * 3}
    out = oext.forward(w, np.array([1, 2]), np.array([0, 1]))
    np.testing.assert_allclose(out["rating_prediction"], D * 1.5 ** 3)          # sum of x_3 = 8 * (1.5)^3, deep part is zero
    np.testing.assert_allclose(out["ctr_prediction"], 0.5)
    b = np.arange(D) * 0.1
    w["cross_v2"] = [(np.zeros((D, D)), b)] * 2
    out = oext.forward(w, np.array([1]), np.array([1]))
    np.testing.assert_allclose(out["rating_prediction"], (1 + 2 * b).sum())
    # an extra user table widens the tower input: identity-like tower on the first E inputs, extra columns weighted by 2
    w["extra_user"] = {"age_id": np.array([[1.0, 1.0], [3.0, 3.0]])}
    w["user_tower"] = [(np.vstack([np.eye(E), 2 * np.ones((2, E))]), np.zeros(E))]
    w["cross_v2"] = []
    out = oext.forward(w, np.array([1, 1, 1]), np.array([1, 1, 1]), {"age_id": np.array([0, 1, 7])})
    np.testing.assert_allclose(out["user_embedding"][:, 0], [1 + 4, 1 + 12, 1 + 0])   # code 7 is outside the table: zeros
