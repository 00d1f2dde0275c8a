"""Multi-rank host logic of the sharded catalog search on CPU (gloo, world_size 2 and 3).

The product's local scan and merge are CUDA kernels; here the *plumbing* of
`recsys_b200.sharded.ShardedIndexFlatIP` (shard bounds, global row offsets, all-gather layout, every
rank ending with the same lists) runs over gloo with the oracle plugged in as the local searcher and
merger -- the oracle being the checker is exactly what tests are allowed to do.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, nq, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import recsys_b200 as rs
        from oracle import flat_ip
        from recsys_b200.sharded import ShardedIndexFlatIP, shard_bounds

        xb = rs.synth.make_unit_rows(n, 64, 5)
        q = rs.synth.make_unit_rows(nq, 64, 6)
        bounds = shard_bounds(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        shard = np.ascontiguousarray(xb[lo:hi])

        def local_search(qq, kk):
            D, I = flat_ip.search_chain(shard, qq, kk, id_base=lo)
            return torch.from_numpy(D), torch.from_numpy(I)

        def merge(Dg, Ig):
            D, I = flat_ip.merge_topk(Dg.numpy(), Ig.numpy())
            return torch.from_numpy(D), torch.from_numpy(I)

        sidx = ShardedIndexFlatIP(64, n, local_search=local_search, merge=merge)
        assert (sidx.row_begin, sidx.row_end) == (lo, hi) and sidx.world == world and sidx.ntotal == n
        D, I = sidx.search(q, k)
        Dw, Iw = flat_ip.search_chain(xb, q, k)
        np.testing.assert_array_equal(I.numpy(), Iw)
        np.testing.assert_array_equal(D.numpy(), Dw)
        np.save(os.path.join(out_dir, f"I_{rank}.npy"), I.numpy())

        # the fused exchange (one all-gather of scores | rows | flags, merge out of the strided buffer, flagged queries
        # re-searched after the exchange).  The stand-in for the deferred tensor-mode search returns GARBAGE for the
        # queries it flags -- rank 1 flags query 2, the last rank flags query 5 -- so only the fallback can fix them.
        bad = {1: [2], world - 1: [5]}

        def local_deferred(qq, kk, out):
            Dl, Il = local_search(qq, kk)
            out[0].copy_(Dl)
            out[1].copy_(Il)
            out[2].zero_()
            for i in bad.get(rank, []):
                out[0][i] = 123.0
                out[1][i] = lo
                out[2][i] = 1

        calls = []

        def merge_strided(Dg, Ig):
            calls.append((tuple(Dg.shape), Dg.stride(0), Ig.stride(0)))
            return merge(Dg.contiguous(), Ig.contiguous())

        fused = ShardedIndexFlatIP(64, n, local_search=local_search, merge=merge_strided, local_deferred=local_deferred)
        D2, I2 = fused.search(q, k)
        np.testing.assert_array_equal(I2.numpy(), Iw)
        np.testing.assert_array_equal(D2.numpy(), Dw)
        assert fused.fallbacks == len({i for v in bad.values() for i in v})
        from recsys_b200.sharded import exchange_layout
        off_I, off_F, total = exchange_layout(nq, k)
        assert calls[0] == ((world, nq, k), total // 4, total // 8)      # parts are slices of the exchange buffer
        assert off_I % 16 == 0 and total % 16 == 0 and off_F == off_I + nq * k * 8

        # shared-threshold protocol (begin -> all-gather of the top lists -> finish -> fused exchange -> merge ->
        # verify), with numpy stand-ins for the kernels: the filter here is an fp32 matmul whose scores may differ
        # from the chain scores by rounding, covered by the 1e-5 margin in `bound` exactly as eps covers bf16.
        FLT_MAX = np.finfo(np.float32).max

        def begin(qq, kk, top):
            scores = qq @ shard.T
            sample = scores[:, ::5]
            R = top.shape[1]
            top.copy_(torch.from_numpy(-np.sort(-sample, axis=1)[:, :R].copy()))
            return scores

        def finish(scores, tops, rank_, out):
            allv = tops.numpy().transpose(1, 0, 2).reshape(scores.shape[0], -1)
            tau = -np.sort(-allv, axis=1)[:, rank_ - 1]
            for i in range(scores.shape[0]):
                rows = np.nonzero(scores[i] >= tau[i])[0]
                Dl = np.full(k, -FLT_MAX, np.float32)
                Il = np.full(k, -1, np.int64)
                if len(rows):
                    d, ix = flat_ip.search_chain(np.ascontiguousarray(shard[rows]), q[i:i + 1], min(k, len(rows)))
                    Dl[:len(rows)][:k] = d[0][:k]
                    Il[:min(k, len(rows))] = rows[ix[0]] + lo
                out[0][i] = torch.from_numpy(Dl)
                out[1][i] = torch.from_numpy(Il)
            out[2].zero_()
            out[3].copy_(torch.from_numpy((tau + 1e-5).astype(np.float32)))

        def verify(D_, I_, Bg, Fg):
            b = Bg.numpy().max(axis=0)
            bad_ = (Fg.numpy() != 0).any(axis=0) | (I_.numpy()[:, k - 1] < 0) | (D_.numpy()[:, k - 1] < b)
            return torch.from_numpy(bad_.astype(np.int32))

        for rank_fn, expect_fallbacks in ((lambda kk, rho: int(np.ceil((2 + np.sqrt(4 + 2 * kk * rho)) ** 2)), False),
                                          (lambda kk, rho: 2, True)):       # rank 2: tau far too high -> repaired
            sh = ShardedIndexFlatIP(64, n, local_search=local_search, merge=merge_strided, local_deferred=local_deferred,
                                    shared=(begin, finish), plan=lambda rows, nq_, kk: (True, (rows + 4) // 5),
                                    rank=rank_fn, verify=verify)
            if k > 128 and not expect_fallbacks:
                assert sh._shared_rank(nq, k) == 0      # rank too large for the kernels: per-shard thresholds
                D3, I3 = sh.search(q, k)
                np.testing.assert_array_equal(I3.numpy(), Iw)
                continue
            assert sh._shared_rank(nq, k) > 0
            D3, I3 = sh.search(q, k)
            np.testing.assert_array_equal(I3.numpy(), Iw)
            np.testing.assert_array_equal(D3.numpy(), Dw)
            assert (sh.fallbacks > 0) == expect_fallbacks, sh.fallbacks

        # exchange="p2p" host logic (csrc/peer.cu does this with NVLink stores on the GPU): the messages travel through
        # an object with publish / gathered / done -- here a gloo stand-in that keeps the region layout [world, slot]
        # -- the tops carry the rank padded to 4 columns, and the consumers read the parts out of the region in place.
        class FakePeer:
            log = []

            def __init__(self, nbytes):
                assert nbytes % 16 == 0
                self.slot, self.region, self.pending = nbytes, torch.zeros((world, nbytes), dtype=torch.uint8), None

            def publish(self, msg):
                assert msg.dtype == torch.uint8 and msg.dim() == 1 and msg.numel() <= self.slot
                self.pending = torch.zeros(self.slot, dtype=torch.uint8)
                self.pending[:msg.numel()] = msg

            def gathered(self):
                dist.all_gather_into_tensor(self.region.view(-1), self.pending)
                return self.region

            def done(self):
                FakePeer.log.append(self.slot)

            def close(self):
                pass

        if k <= 128:
            sh = ShardedIndexFlatIP(64, n, local_search=local_search, merge=merge_strided, local_deferred=local_deferred,
                                    shared=(begin, finish), plan=lambda rows, nq_, kk: (True, (rows + 4) // 5),
                                    rank=lambda kk, rho: int(np.ceil((2 + np.sqrt(4 + 2 * kk * rho)) ** 2)), verify=verify)
            sh.exchange, sh._peer_factory = "p2p", FakePeer
            r_ = sh._shared_rank(nq, k)
            for _ in range(2):
                D4, I4 = sh.search(q, k)
                np.testing.assert_array_equal(I4.numpy(), Iw)
                np.testing.assert_array_equal(D4.numpy(), Dw)
            r_pad = (r_ + 3) // 4 * 4
            assert FakePeer.log == [nq * r_pad * 4, total] * 2          # tops then lists, every step; created once
            assert len(sh._peers) == 1

            # a rank that cannot map its peers makes EVERY rank fall back to the NCCL / gloo all-gathers
            def broken(nbytes):
                if rank == world - 1:
                    raise RuntimeError("librecsys_b200: cudaIpcOpenMemHandle failed (stand-in)")
                return FakePeer(nbytes)
            sh2 = ShardedIndexFlatIP(64, n, local_search=local_search, merge=merge_strided, local_deferred=local_deferred)
            sh2.exchange, sh2._peer_factory = "p2p", broken
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                D5, I5 = sh2.search(q, k)
            np.testing.assert_array_equal(I5.numpy(), Iw)
            assert sh2.exchange == "nccl" and not sh2._peers
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,k", [(2, 3001, 10), (3, 1000, 400)])
def test_sharded_search_over_gloo(tmp_path, world, n, k):
    import oracle
    oracle.build_chain_lib()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, 7, k, str(tmp_path)), nprocs=world, join=True)
    lists = [np.load(tmp_path / f"I_{r}.npy") for r in range(world)]
    for other in lists[1:]:
        np.testing.assert_array_equal(other, lists[0])       # every rank holds the same merged lists


def test_shard_bounds_cover_the_catalog():
    from recsys_b200.sharded import gather_bytes, shard_bounds
    for n in (0, 1, 7, 10_000_000):
        for g in (1, 2, 4, 8):
            b = shard_bounds(n, g)
            assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
            assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1
    assert gather_bytes(8, 1024, 100) == 8 * (1024 * 100 * 12 + 1024 * 8)


def _emb_worker(rank, world, port, n_rows, batch):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import recsys_b200 as rs
        from recsys_b200.sharded import ShardedEmbedding

        dim = 16
        local_rows = np.arange(rank, n_rows, world)
        table = torch.from_numpy(rs.synth.hash_table_rows(local_rows, dim, seed=7))

        def count(ids, G):
            return torch.bincount(ids % G, minlength=G)

        def scatter(ids, G, base):
            owner = (ids % G).numpy()
            order = np.argsort(owner, kind="stable")
            send = torch.from_numpy((ids.numpy() // G)[order])
            perm = np.empty(len(order), dtype=np.int64)
            perm[order] = np.arange(len(order))
            return send, torch.from_numpy(perm)

        emb = ShardedEmbedding(n_rows, dim, local_table=table, plan=(count, scatter), gather=lambda t, r: t[r])
        assert emb.n_local == len(local_rows)
        rng = np.random.default_rng(100 + rank)
        ids = rs.synth.zipf_ids(rng, n_rows, batch)
        got = emb.lookup(torch.from_numpy(ids)).numpy()
        np.testing.assert_array_equal(got, rs.synth.hash_table_rows(ids, dim, seed=7))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_embedding_lookup_over_gloo(world):
    mp.spawn(_emb_worker, args=(world, _free_port(), 1000, 257), nprocs=world, join=True)


def test_zipf_ids_and_hash_rows_are_deterministic():
    import recsys_b200 as rs
    a = rs.synth.zipf_ids(np.random.default_rng(1), 5000, 4096)
    b = rs.synth.zipf_ids(np.random.default_rng(1), 5000, 4096)
    np.testing.assert_array_equal(a, b)
    assert a.min() >= 0 and a.max() < 5000
    top = np.bincount(a, minlength=5000).max()
    assert top > 4096 * 0.02                      # skewed: the hottest row takes a few percent of the batch
    t = rs.synth.hash_table_rows(np.array([0, 1, 49_999_999]), 128, 3)
    assert t.shape == (3, 128) and np.abs(t).max() <= 0.05 and t.std() > 0.02
