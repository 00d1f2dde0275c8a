"""GPU tests of the tcgen05 retrieval path (tc_filter.cu): the raw tensor-core tiles against a bf16 numpy
model, and the filter + exact re-score pipeline against the oracle -- bit-exact scores and indices, the
same bar as the exact scan."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def to_bf16(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 round-to-nearest-even, returned as fp32 (what cvt.rn.bf16x2.f32 does)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n,nq", [(256, 128), (1000, 3), (5000, 130), (70001, 257)])
def test_tcgen05_tiles_match_bf16_model(rs, n, nq):
    """Every (query, row) score out of TMEM equals the fp64 dot of the bf16-rounded operands to fp32 accuracy:
    validates the TMA swizzle, the UMMA descriptors, the K stepping and the TMEM lane/column mapping."""
    rng = np.random.default_rng(n + nq)
    xb = (rng.standard_normal((n, 128)) * rng.uniform(0.5, 2.0, (n, 1))).astype(np.float32)
    q = rng.standard_normal((nq, 128)).astype(np.float32)
    index = rs.IndexFlatIP(128)
    index.add(xb)
    got = index.debug_tensor_scores(dev(q)).cpu().numpy()
    want = to_bf16(q).astype(np.float64) @ to_bf16(xb).astype(np.float64).T
    err = np.abs(got - want).max()
    assert err < 2e-4 * np.abs(want).max(), f"max err {err}"
    # and the rigorous filter bound holds against the exact fp32 scores: |s_hat - s| <= c * |q| * max|x|
    exact = q.astype(np.float64) @ xb.astype(np.float64).T
    bound = (2.0 ** -8 * 1.0625) * np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(xb, axis=1).max()
    assert (np.abs(got - exact) <= bound).all()


def _check(rs, xb, q, k, expect_fallback=None):
    from oracle import flat_ip
    index = rs.IndexFlatIP(128)
    index.add(dev(xb))
    D, I = index.search(dev(q), k, mode="tensor")
    Dw, Iw = flat_ip.search_chain(xb, q, k)
    np.testing.assert_array_equal(I.cpu().numpy(), Iw)
    np.testing.assert_array_equal(D.cpu().numpy(), Dw)
    last, total = index.fallbacks()
    if expect_fallback is True:
        assert last > 0
    elif expect_fallback is False:
        assert last == 0, f"{last} of {q.shape[0]} queries fell back to the exact scan"
    return index


@pytest.mark.parametrize("nq,k", [(1, 10), (7, 1), (128, 100), (129, 128), (300, 100), (1024, 10)])
def test_tensor_search_is_bit_exact(rs, nq, k):
    xb = rs.synth.make_unit_rows(200_000, 128, 77)
    q = rs.synth.make_unit_rows(nq, 128, 78 + nq)
    _check(rs, xb, q, k, expect_fallback=False)


def test_tensor_search_odd_catalog_sizes(rs):
    q = rs.synth.make_unit_rows(40, 128, 5)
    for n in (65536, 65537, 100_003, 131_071):
        _check(rs, rs.synth.make_unit_rows(n, 128, n), q, 100)


def test_tensor_search_non_unit_norms_and_planted_neighbours(rs):
    rng = np.random.default_rng(9)
    n = 150_000
    xb = (rng.standard_normal((n, 128)) * rng.lognormal(0, 0.5, (n, 1))).astype(np.float32)
    q = (rng.standard_normal((64, 128)) * rng.lognormal(0, 0.5, (64, 1))).astype(np.float32)
    for i in range(64):                       # 20 planted near-duplicates of every query direction
        rows = rng.integers(0, n, 20)
        xb[rows] = q[i] * rng.uniform(0.8, 1.2, (20, 1)).astype(np.float32) + 0.01 * rng.standard_normal((20, 128)).astype(np.float32)
    _check(rs, xb, q, 50)


def test_tensor_search_ties_and_degenerate_catalogs_fall_back_but_stay_exact(rs):
    rng = np.random.default_rng(10)
    base = rs.synth.make_unit_rows(1000, 128, 11)
    xb = np.ascontiguousarray(np.tile(base, (80, 1)))            # every row 80 times: massive exact ties
    q = rs.synth.make_unit_rows(9, 128, 12)
    _check(rs, xb, q, 100)
    # a catalog whose scores are packed inside the filter margin: lists overflow -> exact-scan fallback
    centre = rs.synth.make_unit_rows(1, 128, 13)
    xb = (centre + 1e-4 * rng.standard_normal((80_000, 128))).astype(np.float32)
    _check(rs, xb, q, 10, expect_fallback=True)
    _check(rs, np.zeros((70_000, 128), np.float32), q, 10)      # all scores 0: rows 0..9 win


def test_every_chunk_a_hit_chunk_stresses_the_extraction_rings(rs):
    """tc_filter_kernel hands hit chunks (chunk maximum >= tau) to two extraction warps through 8-slot rings per epilogue
    warp.  With every score equal, every lane of every epilogue warp reports a hit in every chunk: the rings are always
    full, the producers wait for slots and hand over in several passes -- and the answer must still be the exact scan's
    (here: the candidate lists overflow, every query is flagged and repeated exactly)."""
    q = rs.synth.make_unit_rows(300, 128, 21)
    _check(rs, np.zeros((70_000, 128), np.float32), q, 10)
    # half of the rows identical to a query's best match: dense hits for some query rows only, sparse for the others
    xb = rs.synth.make_unit_rows(66_000, 128, 22)
    xb[::2] = xb[0]
    _check(rs, xb, q, 100)


def test_auto_mode_equals_exact_mode_and_host_path(rs):
    xb = rs.synth.make_unit_rows(120_000, 128, 21)
    q = rs.synth.make_unit_rows(200, 128, 22)
    index = rs.IndexFlatIP(128)
    index.add(xb)
    De, Ie = index.search(q, 100, mode="exact")
    Da, Ia = index.search(q, 100, mode="auto")
    Dt, It = index.search(q, 100, mode="tensor")
    np.testing.assert_array_equal(Ia, Ie)
    np.testing.assert_array_equal(Da, De)
    np.testing.assert_array_equal(It, Ie)
    np.testing.assert_array_equal(Dt, De)
    index.add(rs.synth.make_unit_rows(5000, 128, 23))            # growing the index refreshes the bf16 shadow
    D2, I2 = index.search(q[:20], 10, mode="tensor")
    D3, I3 = index.search(q[:20], 10, mode="exact")
    np.testing.assert_array_equal(I2, I3)
    np.testing.assert_array_equal(D2, D3)
    with pytest.raises(RuntimeError):
        rs.IndexFlatIP(64).search(np.zeros((1, 64), np.float32), 3, mode="tensor")


def test_full_size_catalog_properties(rs):
    """BASELINE config 3 at full size (10M x 128, 1,024 queries, top-100): the tensor path equals the exact
    scan on every query, lists are sorted by (score desc, row asc), and the returned scores equal the
    stated-arithmetic scores of the returned rows."""
    dev0 = torch.device("cuda", 0)
    N, nq, k = 10_000_000, 1024, 100
    index = rs.IndexFlatIP(128)
    for b0 in range(0, N, 1 << 20):
        g = torch.Generator(device=dev0)
        g.manual_seed(3000 + b0 // (1 << 20))
        x = torch.randn((min(1 << 20, N - b0), 128), generator=g, device=dev0)
        rs.normalize_L2(x)
        index.add(x)
    g = torch.Generator(device=dev0)
    g.manual_seed(3999)
    q = torch.randn((nq, 128), generator=g, device=dev0)
    rs.normalize_L2(q)
    Dt, It = index.search(q, k, mode="tensor")
    last, _ = index.fallbacks()
    assert last <= 2, f"{last} fallbacks"
    De, Ie = index.search(q[:64].contiguous(), k, mode="exact")
    assert torch.equal(It[:64], Ie) and torch.equal(Dt[:64], De)
    assert bool(((Dt[:, :-1] > Dt[:, 1:]) | ((Dt[:, :-1] == Dt[:, 1:]) & (It[:, :-1] < It[:, 1:]))).all())
    assert int(It.min()) >= 0 and int(It.max()) < N
    rows = torch.from_numpy(index.reconstruct_n(int(It[5, 0]), 1)).cuda()
    np.testing.assert_array_equal(rs.ops.matvec(rows, q[5].contiguous()).cpu().numpy(), Dt[5, :1].cpu().numpy())
    # batch-1 latency mode (config 5): top-10 for single queries, both modes identical
    for i in (0, 1, 2):
        D1, I1 = index.search(q[i:i + 1].contiguous(), 10, mode="exact")
        D2, I2 = index.search(q[i:i + 1].contiguous(), 10, mode="tensor")
        assert torch.equal(I1, I2) and torch.equal(D1, D2)
        assert torch.equal(I1[0], It[i, :10])


def test_deferred_search_flags_exactly_what_the_synchronous_search_re_runs(rs):
    """rs_index_search_deferred: no host synchronisation; rows with need_exact == 0 are final (bit-identical to the
    oracle), rows with need_exact == 1 are the ones rs_index_search would have re-run on the exact scan."""
    from oracle import flat_ip
    xb = rs.synth.make_unit_rows(200_000, 128, 77)
    q = rs.synth.make_unit_rows(300, 128, 91)
    index = rs.IndexFlatIP(128)
    index.add(dev(xb))
    D, I, F = index.search_deferred(dev(q), 100)
    assert F.dtype == torch.int32 and int(F.sum()) == 0
    Dw, Iw = flat_ip.search_chain(xb, q, 100)
    np.testing.assert_array_equal(I.cpu().numpy(), Iw)
    np.testing.assert_array_equal(D.cpu().numpy(), Dw)
    # exact mode and shapes outside the tensor path never flag
    _, _, F = index.search_deferred(dev(q), 100, mode="exact")
    assert int(F.sum()) == 0
    # scores packed inside the filter margin: every query is flagged, and the exact re-search repairs it
    rng = np.random.default_rng(10)
    centre = rs.synth.make_unit_rows(1, 128, 13)
    xb = (centre + 1e-4 * rng.standard_normal((80_000, 128))).astype(np.float32)
    q9 = rs.synth.make_unit_rows(9, 128, 12)
    index = rs.IndexFlatIP(128)
    index.add(dev(xb))
    D, I, F = index.search_deferred(dev(q9), 10, id_base=1000)
    assert int(F.sum()) > 0
    idx = F.nonzero().flatten()
    D[idx], I[idx] = index.search(dev(q9)[idx].contiguous(), 10, mode="exact", id_base=1000)
    Dw, Iw = flat_ip.search_chain(xb, q9, 10, id_base=1000)
    np.testing.assert_array_equal(I.cpu().numpy(), Iw)
    np.testing.assert_array_equal(D.cpu().numpy(), Dw)
    # outputs may be views of one exchange buffer (what the sharded search passes)
    from recsys_b200.sharded import exchange_layout
    nq, k = 9, 10
    off_I, off_F, total = exchange_layout(nq, k)
    buf = torch.zeros((total,), dtype=torch.uint8, device="cuda")
    out = (buf[:nq * k * 4].view(torch.float32).view(nq, k), buf[off_I:off_I + nq * k * 8].view(torch.int64).view(nq, k),
           buf[off_F:off_F + nq * 4].view(torch.int32))
    index.search_deferred(dev(q9), k, mode="exact", id_base=1000, out=out)
    np.testing.assert_array_equal(out[1].cpu().numpy(), Iw)


def test_topk_merge_reads_parts_out_of_a_strided_exchange_buffer(rs):
    from oracle import flat_ip
    from recsys_b200.sharded import exchange_layout
    G, nq, k = 4, 9, 30
    xb = rs.synth.make_unit_rows(6000, 64, 12)
    q = rs.synth.make_unit_rows(nq, 64, 13)
    bounds = np.linspace(0, 6000, G + 1).astype(int)
    off_I, off_F, total = exchange_layout(nq, k)
    recv = torch.zeros((G, total), dtype=torch.uint8, device="cuda")
    for g, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
        Dl, Il = flat_ip.search_chain(np.ascontiguousarray(xb[a:b]), q, k, id_base=int(a))
        recv[g, :nq * k * 4].view(torch.float32).view(nq, k).copy_(dev(Dl))
        recv[g, off_I:off_I + nq * k * 8].view(torch.int64).view(nq, k).copy_(dev(Il))
    Dg = recv[:, :nq * k * 4].view(torch.float32).view(G, nq, k)
    Ig = recv[:, off_I:off_I + nq * k * 8].view(torch.int64).view(G, nq, k)
    assert not Dg.is_contiguous()
    D, I = rs.ops.topk_merge(Dg, Ig)
    Dw, Iw = flat_ip.search_chain(xb, q, k)
    np.testing.assert_array_equal(I.cpu().numpy(), Iw)
    np.testing.assert_array_equal(D.cpu().numpy(), Dw)


def test_shared_threshold_protocol_over_simulated_shards(rs):
    """The sharded catalog's shared-threshold search with the shards as four indexes on one GPU (the all-gathers become
    torch.stack): begin -> tops -> finish -> merge -> verify must give the single-index result bit for bit, with no
    query flagged, while single shards keep fewer than k candidates."""
    from oracle import flat_ip
    G, n_shard, nq, k = 4, 100_000, 300, 100
    xb = rs.synth.make_unit_rows(G * n_shard, 128, 501)
    q = rs.synth.make_unit_rows(nq, 128, 502)
    qd = dev(q)
    shards = []
    for g in range(G):
        ix = rs.IndexFlatIP(128)
        ix.add(dev(xb[g * n_shard:(g + 1) * n_shard]))
        shards.append(ix)
    ok, rows = rs.IndexFlatIP.tensor_plan(128, n_shard, nq, k)
    assert ok and rows >= 64 * 256
    assert rs.IndexFlatIP.tensor_plan(128, 1000, nq, k)[0] is False and rs.IndexFlatIP.tensor_plan(64, n_shard, nq, k)[0] is False
    rank = rs.IndexFlatIP.threshold_rank(k, G * rows / (G * n_shard))
    assert 4 < rank <= 128

    def run(rank_):
        tops, sess = [], []
        for ix in shards:
            top = torch.empty((nq, rank), dtype=torch.float32, device="cuda")
            sess.append(ix.shared_begin(qd, k, top))
            tops.append(top)
        tops = torch.stack(tops)
        outs = []
        for g, ix in enumerate(shards):
            out = (torch.empty((nq, k), dtype=torch.float32, device="cuda"), torch.empty((nq, k), dtype=torch.int64, device="cuda"),
                   torch.empty((nq,), dtype=torch.int32, device="cuda"), torch.empty((nq,), dtype=torch.float32, device="cuda"))
            ix.shared_finish(sess[g], tops, rank_, id_base=g * n_shard, out=out)
            outs.append(out)
        D, I = rs.ops.topk_merge(torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]))
        need = rs.ops.topk_verify(D, I, torch.stack([o[3] for o in outs]), torch.stack([o[2] for o in outs]))
        return D, I, need, outs

    D, I, need, outs = run(rank)
    assert int(need.sum()) == 0
    Dw, Iw = flat_ip.search_chain(xb, q, k)
    np.testing.assert_array_equal(I.cpu().numpy(), Iw)
    np.testing.assert_array_equal(D.cpu().numpy(), Dw)
    # the top lists really are each query's largest sample maxima: the shared tau is below every shard's best score
    per_shard = torch.stack([(o[1] >= 0).sum(dim=1) for o in outs]).float()
    assert float(per_shard.sum(dim=0).min()) >= k            # together the shards hold at least k candidates per query
    assert bool((per_shard < k).any())                       # ... while single shards come back short (padded lists)
    # a threshold guessed far too high (rank 1 = the largest sample maximum): lists come back short, verify flags them
    D, I, need, _ = run(1)
    assert int(need.sum()) > 0.9 * nq
    ok_rows = (need == 0).cpu().numpy()
    np.testing.assert_array_equal(I.cpu().numpy()[ok_rows], Iw[ok_rows])     # whatever verify lets through is exact


@pytest.mark.parametrize("n,nq,k", [(3706, 4096, 10), (3706, 77, 28), (256, 5, 1), (257, 130, 2), (1000, 300, 7),
                                    (4095, 64, 31), (16384, 64, 128), (32768, 200, 100), (20000, 1, 10)])
def test_small_catalog_tensor_path_is_bit_exact(rs, n, nq, k):
    """Catalogs of 256 .. 32,768 rows (the reference's own catalog: 3,706 items, BASELINE configs[0]) on the tensor path:
    every tile scored twice, threshold = k-th largest 32-row chunk maximum minus 2 eps -- against the oracle's fp32 FMA
    chain, bit for bit, with no query falling back to the exact scan."""
    xb = rs.synth.make_unit_rows(n, 128, n + 1)
    q = rs.synth.make_unit_rows(nq, 128, n + nq)
    _check(rs, xb, q, k, expect_fallback=False)


def test_small_catalog_ties_negative_scores_and_auto_mode(rs):
    rng = np.random.default_rng(31)
    base = rs.synth.make_unit_rows(1000, 128, 32)
    xb = np.ascontiguousarray(np.tile(base, (3, 1)))             # 3,000 rows, every row 3 times: exact ties across chunks
    q = rs.synth.make_unit_rows(33, 128, 33)
    _check(rs, xb, q, 20)
    # all scores negative: the zero-filled rows past the catalog (score 0) must not raise the threshold
    d = rs.synth.make_unit_rows(1, 128, 34)
    xb = (-np.abs(rng.standard_normal((3706, 1))).astype(np.float32) - 0.5) * d
    _check(rs, xb.astype(np.float32), np.repeat(d, 7, axis=0), 10, expect_fallback=False)
    # auto mode picks the tensor path for the small catalog and equals the exact scan
    xb = rs.synth.make_unit_rows(3706, 128, 35)
    index = rs.IndexFlatIP(128)
    index.add(xb)
    qq = rs.synth.make_unit_rows(500, 128, 36)
    De, Ie = index.search(qq, 10, mode="exact")
    Da, Ia = index.search(qq, 10, mode="auto")
    np.testing.assert_array_equal(Ia, Ie)
    np.testing.assert_array_equal(Da, De)
    with pytest.raises(RuntimeError):                            # the top-k must be sparse in a small catalog: k <= rows / 128
        index.search(qq, 100, mode="tensor")
    Da, Ia = index.search(qq, 100, mode="auto")                  # ... auto then stays on the exact scan
    De, Ie = index.search(qq, 100, mode="exact")
    np.testing.assert_array_equal(Ia, Ie)
    tiny = rs.IndexFlatIP(128)
    tiny.add(xb[:100])
    Dt, It = tiny.search(qq[:4], 10)                             # auto: exact scan
    Dw, Iw = tiny.search(qq[:4], 10, mode="exact")
    np.testing.assert_array_equal(It, Iw)


def test_small_degenerate_catalog_auto_mode_stays_exact_and_fast(rs):
    """Item embeddings that all point the same way pack every score inside the bf16 filter's margin: the small-catalog
    tensor path repeats its queries on the exact scan (results stay exact); AUTO notices and stays on the scan until rows
    are added; an explicit tensor request still tries."""
    rng = np.random.default_rng(41)
    centre = rs.synth.make_unit_rows(1, 128, 42)
    xb = (centre + 1e-4 * rng.standard_normal((3706, 128))).astype(np.float32)
    q = rs.synth.make_unit_rows(256, 128, 43)
    index = rs.IndexFlatIP(128)
    index.add(xb)
    De, Ie = index.search(q, 10, mode="exact")
    total0 = index.fallbacks()[1]
    D1, I1 = index.search(q, 10)                      # AUTO: tries the tensor path, falls back
    np.testing.assert_array_equal(I1, Ie)
    np.testing.assert_array_equal(D1, De)
    total1 = index.fallbacks()[1]
    assert total1 > total0
    D2, I2 = index.search(q, 10)                      # AUTO again: straight to the exact scan, no new fallbacks
    np.testing.assert_array_equal(I2, Ie)
    assert index.fallbacks()[1] == total1
    index.search(q, 10, mode="tensor")                # explicit request: tries again
    assert index.fallbacks()[1] > total1
    index.add(rs.synth.make_unit_rows(300, 128, 44))  # new rows: AUTO gives the tensor path another chance
    t2 = index.fallbacks()[1]
    index.search(q, 10)
    assert index.fallbacks()[1] > t2


@pytest.mark.parametrize("nq,k,n", [(1, 10, 200_000), (2, 100, 200_000), (1, 128, 65_536), (2, 1, 100_003), (1, 100, 65_537)])
def test_bf16_shadow_scan_for_one_or_two_queries_is_bit_exact(rs, nq, k, n, monkeypatch):
    """The batch-1 / batch-2 path (scan_bf16_topk_kernel: bf16 shadow for the bulk, exact fp32 re-score of every row that
    could still make the list) against the oracle's FMA chain, and against the seven-launch tensor path it replaces."""
    xb = rs.synth.make_unit_rows(n, 128, n + k)
    q = rs.synth.make_unit_rows(nq, 128, n + nq)
    index = _check(rs, xb, q, k, expect_fallback=False)
    n0 = rs._lib.launch_count()
    D, I = index.search(dev(q), k)
    assert rs._lib.launch_count() - n0 == 2                      # scan + merge
    monkeypatch.setenv("RS_BF16_SCAN", "0")                      # read once per process: only documents the knob here
    De, Ie = index.search(dev(q), k, mode="exact")
    assert torch.equal(I, Ie) and torch.equal(D, De)


def test_bf16_shadow_scan_ties_and_non_unit_rows(rs):
    rng = np.random.default_rng(51)
    base = rs.synth.make_unit_rows(1000, 128, 52)
    xb = np.ascontiguousarray(np.tile(base, (80, 1)))            # every row 80 times: exact ties, row-ascending order decides
    q = rs.synth.make_unit_rows(2, 128, 53)
    _check(rs, xb, q[:1], 100)
    _check(rs, xb, q, 10)
    xb = (rng.standard_normal((90_000, 128)) * rng.lognormal(0, 0.7, (90_000, 1))).astype(np.float32)
    qq = (rng.standard_normal((1, 128)) * 3.0).astype(np.float32)
    _check(rs, xb, qq, 50)
    _check(rs, np.zeros((70_000, 128), np.float32), q[:1], 10)  # all scores 0: rows 0..9 win
