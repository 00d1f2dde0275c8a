"""GPU tests of the EXTENSIONS (recsys_b200.extensions): DCN-v2 matrix cross and towers with extra categorical tables.
The reference has neither, so the checker is oracle/extensions.py -- a numpy fp64 restatement of the stated formulas --
plus closed-form known-answer cases.  Tolerance: 1e-5 of each tensor's scale (fp32 mode; the D x D cross products run on
the fp32 tensor kernel from 512 rows and on the FFMA kernel below)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-5


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(got, want, name):
    got = got.cpu().numpy().astype(np.float64)
    assert got.shape == want.shape, name
    assert np.abs(got - want).max() <= TOL * max(np.abs(want).max(), 1e-30), f"{name}: {np.abs(got - want).max():.3e} of {np.abs(want).max():.3e}"


def test_cross_v2_known_answers(rs):
    """W_l = c I, b_l = 0  =>  x_{l+1} = x0 * (c x_l) + x_l; with x0 = 1 (all ones): x_l = (1 + c)^l.  And W = 0 with bias
    b: x_{l+1} = x0 * b + x_l = x0 (1 + l b) -- the bias sits INSIDE the product in v2 (outside in the reference's v1)."""
    from recsys_b200.extensions import DeepCrossNetworkV2
    D, c = 256, 0.5
    for B in (3, 700):                                          # FFMA kernel / fp32 tensor kernel
        net = DeepCrossNetworkV2(cross_layers=3, deep_layers=[256, 128])
        net.set_cross_weights([c * np.eye(D, dtype=np.float32)] * 3, [np.zeros(D, np.float32)] * 3)
        x0 = torch.ones((B, D), device="cuda")
        out = net(x0)
        assert out.shape == (B, D + 128)
        np.testing.assert_allclose(out[:, :D].cpu().numpy(), (1 + c) ** 3, rtol=1e-6)
        rng = np.random.default_rng(B)
        x = rng.standard_normal((B, D)).astype(np.float32)
        b = rng.standard_normal(D).astype(np.float32)
        net.set_cross_weights([np.zeros((D, D), np.float32)] * 2, [b] * 2)
        got = net(dev(x))[:, :D].cpu().numpy().astype(np.float64)
        want = x.astype(np.float64) * (1 + 2 * b.astype(np.float64))
        assert np.abs(got - want).max() <= TOL * np.abs(want).max()
    zero = DeepCrossNetworkV2(cross_layers=0, deep_layers=[256, 128])
    x = dev(np.random.default_rng(1).standard_normal((5, D)).astype(np.float32))
    assert torch.equal(zero(x)[:, :D], x)


@pytest.mark.parametrize("cross", ["v2", "v1"])
@pytest.mark.parametrize("B", [9, 4096 + 17])
def test_extended_model_matches_the_restatement(rs, cross, B):
    """Extra tables (age 7 x 32, gender 2 x 32, occupation 21 x 64 -> tower input 256; one extra item table 18 x 128) + DCN
    (v2 / v1) + heads against oracle/extensions.py; string user ids with an out-of-vocabulary entry, int item rows."""
    from oracle import extensions as oext
    from recsys_b200.extensions import ExtendedMultiTaskModel
    cfg = rs.ModelConfig()
    uv, iv = rs.synth.make_vocab(300), rs.synth.make_vocab(200)
    m = ExtendedMultiTaskModel(cfg, uv, iv, extra_user_features={"age_id": (7, 32), "gender_id": (2, 32), "occupation_id": (21, 64)},
                               extra_item_features={"genre_id": (18, 128)}, cross=cross)
    rng = np.random.default_rng(B)
    if cross == "v1":       # non-trivial v1 cross vectors / biases (build() starts them at glorot / zero)
        m.dcn.set_cross_weights([rng.standard_normal(256).astype(np.float32) * 0.05 for _ in range(3)],
                                [rng.standard_normal(256).astype(np.float32) * 0.05 for _ in range(3)])
    ur, ir = rs.synth.make_ids(rng, 300, B), rs.synth.make_ids(rng, 200, B)
    users = [uv[i - 1] if i > 0 else "no-such-user" for i in ur]
    feats = {"user_id": users, "movie_id": dev(ir), "age_id": dev(rng.integers(0, 7, B)), "gender_id": dev(rng.integers(0, 2, B)),
             "occupation_id": dev(rng.integers(0, 21, B)), "genre_id": dev(rng.integers(0, 18, B))}
    out = m(feats)
    w = m.get_weights()
    want = oext.forward(w, ur, ir, {k: feats[k].cpu().numpy() for k in ("age_id", "gender_id", "occupation_id")},
                        {"genre_id": feats["genre_id"].cpu().numpy()})
    for k_ in want:
        close(out[k_], want[k_], f"{cross} B={B} {k_}")
