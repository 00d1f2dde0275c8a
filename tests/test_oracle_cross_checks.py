"""The oracle (numpy / C restatement of the reference's TF + FAISS path) against an INDEPENDENT implementation of the
same layer semantics: PyTorch's CPU operators -- the only other implementation importable in this image (TensorFlow,
Keras and faiss are absent, so the primitives' arithmetic stays unpinned with respect to those binaries; this narrows
the room for a restatement error, it does not close it -- the composition is pinned separately, test_reference_run.py).  torch.nn.functional.{embedding, linear, relu}, torch.sigmoid
and torch.topk are written by other people against the same definitions (Embedding = row gather, Dense = x @ W + b
with W stored (in, out) in Keras and (out, in) in torch, DCN-v1 cross per the paper's formula)."""

import numpy as np
import torch
import torch.nn.functional as F

import recsys_b200 as rs
from oracle import flat_ip
from oracle.models import MultiTaskModel as OracleMultiTask


def _torch_forward(w, ur, ir):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double()   # noqa: E731

    def tower(side, rows):
        x = F.embedding(torch.from_numpy(rows), t(w[f"{side}_embedding"]))
        layers = w[f"{side}_tower"]
        for i, (kernel, bias) in enumerate(layers):
            x = F.linear(x, t(kernel).T, t(bias))                      # torch keeps (out, in)
            if i < len(layers) - 1:
                x = F.relu(x)
        return x

    u, it = tower("user", ur), tower("item", ir)
    x0 = torch.cat([u, it], dim=1)
    xl = x0
    for cw, cb in zip(w["cross_w"], w["cross_b"]):                     # DCN-v1: x_{l+1} = x0 * (x_l . w_l) + b_l + x_l
        xl = x0 * (xl @ t(cw)) + t(cb) + xl
    deep = x0
    for kernel, bias in w["deep"]:
        deep = F.relu(F.linear(deep, t(kernel).T, t(bias)))
    z = torch.cat([xl, deep], dim=1)
    rating = F.linear(z, t(w["rating_head"][0]).T, t(w["rating_head"][1]))
    ctr = torch.sigmoid(F.linear(z, t(w["ctr_head"][0]).T, t(w["ctr_head"][1])))
    return {"user_embedding": u, "item_embedding": it, "rating_prediction": rating, "ctr_prediction": ctr}


def test_ranking_oracle_matches_torch_cpu_ops():
    for cross_layers in (0, 3):
        cfg = rs.ModelConfig()
        cfg.cross_layers = cross_layers
        w = rs.synth.make_weights(cfg, 200, 150, seed=40 + cross_layers)
        rng = np.random.default_rng(2)
        w["cross_b"] = [(0.05 * rng.standard_normal(b.shape)).astype(np.float32) for b in w["cross_b"]]
        ur, ir = rs.synth.make_ids(rng, 200, 300), rs.synth.make_ids(rng, 150, 300)
        want = _torch_forward(w, ur, ir)
        got = OracleMultiTask(w, rs.synth.make_vocab(200), rs.synth.make_vocab(150), np.float64).rank_rows(ur, ir)
        for key, ref in want.items():
            ref = ref.numpy()
            assert got[key].shape == ref.shape, key
            assert np.abs(got[key] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), key


def test_retrieval_oracle_matches_torch_topk_and_normalize():
    xb = rs.synth.make_unit_rows(5000, 64, 11) * np.float32(1.7)       # not unit length: normalisation has work to do
    q = rs.synth.make_unit_rows(33, 64, 12)
    a = xb.copy()
    flat_ip.normalize_L2(a)
    b = F.normalize(torch.from_numpy(xb).double(), dim=1).numpy()
    assert np.abs(a - b).max() <= 2e-7                                  # fp32 chain vs fp64: rounding only
    D, I = flat_ip.search_chain(a, q, 20)
    s64 = torch.from_numpy(q).double() @ torch.from_numpy(a).double().T
    Dt, It = torch.topk(s64, 20, dim=1)                                 # descending, like IndexFlatIP.search
    # fp32 chain scores can reorder near-ties of the fp64 scores: compare as sets unless the gap is clear
    for r in range(q.shape[0]):
        gaps = (Dt[r, :-1] - Dt[r, 1:]).numpy()
        if gaps.min() > 1e-5:
            np.testing.assert_array_equal(I[r], It[r].numpy())
        else:
            assert set(I[r].tolist()) == set(It[r].tolist()) or abs(float(Dt[r, -1]) - float(D[r, -1])) < 1e-5
        assert np.abs(D[r] - Dt[r].numpy()).max() <= 1e-5
    # padding and merge semantics against a plain concatenate + topk
    Dp, Ip = flat_ip.search_chain(a[:7], q, 10)
    assert (Ip[:, 7:] == -1).all() and (Dp[:, 7:] == -np.finfo(np.float32).max).all()
    halves = [flat_ip.search_chain(a[:2500], q, 20), flat_ip.search_chain(a[2500:], q, 20, id_base=2500)]
    Dm, Im = flat_ip.merge_topk(np.stack([h[0] for h in halves]), np.stack([h[1] for h in halves]))
    np.testing.assert_array_equal(Im, I)
    np.testing.assert_array_equal(Dm, D)
