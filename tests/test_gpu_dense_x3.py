"""GPU tests of the fp32 tensor mode (csrc/dense_x3.cu): Keras Dense on tcgen05 with every fp32 operand split into two
fp16 numbers (v = hi + lo * 2^-11; x.w = hi.hi + 2^-11 (hi.lo + lo.hi), fp32 accumulation).

Tolerances, all against fp64 arithmetic on the SAME fp32 inputs (numpy restatement of src/models.py:27 -- y = act(x @ kernel
+ bias)):
  * one layer: |y - y64| <= 3e-6 * max|y64|  (X3_LAYER_TOL; the dropped lo.lo term is 2^-22 relative per product and the
    fp32 accumulation of a K <= 256 sum rounds at 2^-24 per add; measured values are printed)
  * the FFMA kernel (rs_dense_forward) under the same measure, printed beside it for comparison
  * MultiTaskModel forward in fp32 mode (now on this kernel from 512 rows): 1e-5 of each tensor's scale, the tolerance
    north_star states for fp32 -- tests/test_gpu_parity.py holds those cases; here the two fp32 kernels are compared.
"""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

X3_LAYER_TOL = 3e-6


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ref64(x, w, b, act):
    y = x.astype(np.float64) @ w.astype(np.float64) + (0 if b is None else b.astype(np.float64))
    return np.maximum(y, 0) if act == "relu" else y


@pytest.mark.parametrize("K,N", [(128, 256), (256, 128), (128, 64), (64, 128), (256, 256), (128, 128), (256, 64), (64, 64), (64, 256)])
@pytest.mark.parametrize("B,act", [(128, "relu"), (1000, None), (4099, "relu"), (128 * 148 * 2 + 77, None)])
def test_dense_x3_matches_fp64(rs, K, N, B, act):
    rng = np.random.default_rng(K * 7 + N * 3 + B)
    x = rng.standard_normal((B, K)).astype(np.float32)
    x[rng.integers(0, B, 5), rng.integers(0, K, 5)] = 0.0
    x[0, :4] = [1e-6, -3e-5, 250.0, -1.0e3]                     # tiny values (lo in fp16's low range) and large ones
    w = rs.synth.glorot_uniform(rng, K, N)
    b = rng.normal(0, 0.1, N).astype(np.float32)
    assert rs.ops.dense_x3_supported(K, N, act)
    hi, lo = rs.ops.dense_x3_prepare(dev(w))
    y = rs.ops.dense_x3(dev(x), hi, lo, dev(b), act)
    want = ref64(x, w, b, act)
    scale = np.abs(want).max()
    err = np.abs(y.cpu().numpy().astype(np.float64) - want).max() / scale
    y_ffma = rs.ops.dense(dev(x), dev(w), dev(b), act)
    err_ffma = np.abs(y_ffma.cpu().numpy().astype(np.float64) - want).max() / scale
    print(f"x3 K={K} N={N} B={B}: max err / scale = {err:.2e} (FFMA kernel: {err_ffma:.2e})")
    assert err <= X3_LAYER_TOL


def test_dense_x3_strided_in_and_out(rs):
    """Column slices of wider buffers on both sides (how the towers write into the concatenated DCN input), no bias."""
    rng = np.random.default_rng(2)
    B = 1500
    wide_in = rng.standard_normal((B, 384)).astype(np.float32)
    w = rs.synth.glorot_uniform(rng, 128, 128)
    xin = dev(wide_in)[:, 128:256]
    wide_out = torch.full((B, 256), 7.0, device="cuda")
    hi, lo = rs.ops.dense_x3_prepare(dev(w))
    rs.ops.dense_x3(xin, hi, lo, None, None, out=wide_out[:, 128:])
    want = ref64(wide_in[:, 128:256], w, None, None)
    got = wide_out.cpu().numpy()
    assert np.abs(got[:, 128:] - want).max() <= X3_LAYER_TOL * np.abs(want).max()
    assert np.all(got[:, :128] == 7.0)                          # nothing outside the slice was touched


def test_fp32_model_forward_uses_x3_and_agrees_with_ffma(rs, monkeypatch):
    """MultiTaskModel in fp32 mode at a ranking batch: the tensor-core kernel (default) and the FFMA kernel
    (RS_FP32_TENSOR=0) are two implementations of one tolerance; both must sit within 1e-5 of the fp64 oracle, and
    the unsupported-shape / small-batch cases must silently stay on the FFMA kernel."""
    from oracle.models import MultiTaskModel as OMT
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, 6040, 3706, seed=10)
    uv, iv = rs.synth.make_vocab(6040), rs.synth.make_vocab(3706)
    m = rs.MultiTaskModel(cfg, uv, iv, {})
    rs.load_weights(m, w)
    rng = np.random.default_rng(100)
    B = 8192 + 33
    ur, ir = rs.synth.make_ids(rng, 6040, B), rs.synth.make_ids(rng, 3706, B)
    feats = {"user_id": dev(ur), "movie_id": dev(ir)}
    want = OMT(w, uv, iv, np.float64).rank_rows(ur, ir)
    n0 = rs._lib.launch_count()
    out_x3 = {k_: v.clone() for k_, v in m(feats).items()}
    n_x3 = rs._lib.launch_count() - n0
    monkeypatch.setenv("RS_FP32_TENSOR", "0")
    out_ffma = m(feats)
    for k_ in want:
        scale = np.abs(want[k_]).max()
        e3 = np.abs(out_x3[k_].cpu().numpy().astype(np.float64) - want[k_]).max() / scale
        ef = np.abs(out_ffma[k_].cpu().numpy().astype(np.float64) - want[k_]).max() / scale
        print(f"fp32 mode {k_}: tensor kernel {e3:.2e}, FFMA kernel {ef:.2e} (of tensor scale)")
        assert e3 <= 1e-5 and ef <= 1e-5, k_
    assert n_x3 > 0
    # a layer width outside {64,128,256} is not served by the tensor kernel
    assert not rs.ops.dense_x3_supported(100, 128, None) and not rs.ops.dense_x3_supported(128, 128, "sigmoid")


def test_split_plane_chain_and_two_problems_per_launch(rs):
    """What MultiTaskModel's fp32 mode does: two towers of the same shape in the same launches, the Embedding gather fused
    into the first layer (int32 and int64 ids, an out-of-range id reads zeros), split fp16 planes between the layers, fp32
    AND split output from the last layer, and a following chain fed from those planes -- against fp64 numpy."""
    from recsys_b200.models import Dense, x3_chain
    rng = np.random.default_rng(11)
    B, E = 5000, 128
    dims = [256, 128, 64, 128]
    acts = ["relu", "relu", "relu", None]
    tables = [rng.standard_normal((700, E)).astype(np.float32) * 0.3, rng.standard_normal((450, E)).astype(np.float32) * 0.3]
    ids = [rng.integers(0, 700, B).astype(np.int32), rng.integers(0, 450, B).astype(np.int64)]
    ids[0][7] = 5000                                             # outside the table: a row of zeros
    towers, weights = [], []
    for t in range(2):
        layers, ws, d = [], [], E
        for u, a in zip(dims, acts):
            L = Dense(u, activation=a)
            k, b = rs.synth.glorot_uniform(rng, d, u), rng.normal(0, 0.05, u).astype(np.float32)
            L.set_weights([k, b])
            layers.append(L)
            ws.append((k, b))
            d = u
        towers.append(layers)
        weights.append(ws)
    combined = torch.zeros((B, 2 * E), device="cuda")
    chi, clo = rs.ops.split_planes(B, 2 * E, "cuda")
    outs = x3_chain([(towers[0], {"x": dev(tables[0]), "ids": dev(ids[0])}), (towers[1], {"x": dev(tables[1]), "ids": dev(ids[1])})],
                    [combined[:, :E], combined[:, E:]], split_outs=[(chi[:, :E], clo[:, :E]), (chi[:, E:], clo[:, E:])])
    want = []
    for t in range(2):
        x = tables[t][np.clip(ids[t], 0, len(tables[t]) - 1)].astype(np.float64)
        x[ids[t] >= len(tables[t])] = 0
        for (k, b), a in zip(weights[t], acts):
            x = x @ k.astype(np.float64) + b
            x = np.maximum(x, 0) if a == "relu" else x
        want.append(x)
    for t in range(2):
        got = outs[t].cpu().numpy().astype(np.float64)
        err = np.abs(got - want[t]).max() / np.abs(want[t]).max()
        print(f"tower {t}: 4-layer chain max err / scale = {err:.2e}")
        assert err <= 1e-5
    # the split copy of the outputs reproduces them to 2^-22 relative, and feeds the next chain
    rec = chi.float().cpu().numpy().astype(np.float64) + clo.float().cpu().numpy().astype(np.float64) / 2048.0
    assert np.abs(rec - combined.cpu().numpy()).max() <= 2.0 ** -21 * np.abs(combined.cpu().numpy()).max()
    deep = []
    d = 2 * E
    for u in (256, 128):
        L = Dense(u, activation="relu")
        k, b = rs.synth.glorot_uniform(rng, d, u), rng.normal(0, 0.05, u).astype(np.float32)
        L.set_weights([k, b])
        deep.append((L, k, b))
        d = u
    y = x3_chain([([L for L, _, _ in deep], {"x_hi": chi, "x_lo": clo})], [None])[0]
    x = np.concatenate(want, axis=1)
    for _, k, b in deep:
        x = np.maximum(x @ k.astype(np.float64) + b, 0)
    err = np.abs(y.cpu().numpy().astype(np.float64) - x).max() / np.abs(x).max()
    print(f"deep chain fed from split planes: max err / scale = {err:.2e}")
    assert err <= 1e-5
