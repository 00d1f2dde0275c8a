#!/usr/bin/env python
"""Multi-GPU parity check of the sharded catalog search (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

Every rank builds its row shard; the sharded search (local tensor-path scan -> NCCL all-gather -> merge
kernel) must equal, bit for bit, a single-GPU search over the whole catalog (built on every rank here).
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402
from recsys_b200.sharded import ShardedIndexFlatIP  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, nq, k = 1_500_003, 300, 100
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    xb = torch.randn((N, 128), generator=g, device=dev)
    rs.normalize_L2(xb)
    q = torch.randn((nq, 128), generator=g, device=dev)
    rs.normalize_L2(q)
    full = rs.IndexFlatIP(128)
    full.add(xb)
    Dw, Iw = full.search(q, k)
    sidx = ShardedIndexFlatIP(128, N)
    sidx.add_local(xb[sidx.row_begin:sidx.row_end].contiguous())
    # default: threshold shared by all shards (one extra small all-gather), verification after the merge
    assert sidx._shared_rank(nq, k) > 0
    D, I = sidx.search(q, k)
    assert torch.equal(I, Iw) and torch.equal(D, Dw), f"rank {rank}: shared-threshold sharded search differs"
    assert sidx.fallbacks == 0
    sidx.shared_threshold = False
    sd = sidx.local.search_deferred
    for mode in ("tensor", "exact"):
        # fused exchange: deferred local search -> one all-gather (scores | rows | flags) -> merge out of the buffer
        sidx._local_deferred = lambda q_, k_, out, m=mode: sd(q_, k_, id_base=sidx.row_begin, mode=m, out=out)
        D, I = sidx.search(q, k)
        assert torch.equal(I, Iw) and torch.equal(D, Dw), f"rank {rank}: sharded {mode} search differs"
    assert sidx.fallbacks == 0

    # a shard that flags queries (and returns garbage for them) is repaired by the post-exchange exact re-search
    def flagged(q_, k_, out):
        sd(q_, k_, id_base=sidx.row_begin, out=out)
        mine = [3 + rank, 100]
        out[0][mine] = 9.0
        out[1][mine] = sidx.row_begin
        out[2][mine] = 1
    sidx._local_deferred = flagged
    D, I = sidx.search(q, k)
    assert torch.equal(I, Iw) and torch.equal(D, Dw), f"rank {rank}: flagged queries were not repaired"
    assert sidx.fallbacks == world + 1
    # the two-collective path (no deferred search) stays available
    sidx._local_deferred = None
    D, I = sidx.search(q, k)
    assert torch.equal(I, Iw) and torch.equal(D, Dw)
    sidx._local_deferred = lambda q_, k_, out: sd(q_, k_, id_base=sidx.row_begin, out=out)
    sidx.shared_threshold = True
    # the same step replayed as a CUDA graph (kernels + both NCCL collectives captured once per (nq, k))
    sidx.use_graph = True
    for _ in range(3):
        D, I = sidx.search(q, k)
        assert torch.equal(I, Iw) and torch.equal(D, Dw), f"rank {rank}: graph-replayed sharded search differs"
    q2 = torch.roll(q, 7, dims=0).contiguous()                      # new queries through the captured graph
    D, I = sidx.search(q2, k)
    assert torch.equal(I, torch.roll(Iw, 7, dims=0)) and torch.equal(D, torch.roll(Dw, 7, dims=0))
    sidx.use_graph = False
    # both exchanges of the step through NVLink peer memory (csrc/peer.cu) instead of NCCL: eager, graph, other shapes
    sidx.exchange = "p2p"
    for _ in range(3):
        D, I = sidx.search(q, k)
        assert torch.equal(I, Iw) and torch.equal(D, Dw), f"rank {rank}: peer-memory sharded search differs"
    sidx.use_graph = True
    for shift in (0, 0, 5, 11):
        qs = torch.roll(q, shift, dims=0).contiguous()
        D, I = sidx.search(qs, k)
        assert torch.equal(I, torch.roll(Iw, shift, dims=0)) and torch.equal(D, torch.roll(Dw, shift, dims=0)), \
            f"rank {rank}: graph-replayed peer-memory search differs (shift {shift})"
    sidx.use_graph = False
    sidx.shared_threshold = False                                   # per-shard thresholds: one exchange per step
    D, I = sidx.search(q, k)
    assert torch.equal(I, Iw) and torch.equal(D, Dw), f"rank {rank}: peer-memory search (per-shard thresholds) differs"
    sidx.shared_threshold = True
    D7, I7 = sidx.search(q[:7].contiguous(), 10)                    # ragged shape: rank padded to 4 columns
    assert torch.equal(I7, Iw[:7, :10]) and torch.equal(D7, Dw[:7, :10])
    assert sidx.fallbacks == world + 1
    if rank == 0:
        print("peer-memory exchange OK (eager, CUDA graph, per-shard thresholds, ragged shapes)", flush=True)
    # batch-1 (config 5)
    D1, I1 = sidx.search(q[:1].contiguous(), 10)
    assert torch.equal(I1, Iw[:1, :10])
    checks = ["shared-threshold search (NCCL)", "per-shard thresholds tensor / exact (NCCL)", "flagged-query repair",
              "two-collective path", "CUDA-graph replay (NCCL)", "peer-memory exchange eager / graph / per-shard / ragged",
              "batch-1"]

    # ---- the C entry points a non-Python host would call (csrc/sharded.cu), driven directly ----
    import ctypes
    from recsys_b200 import _lib
    from recsys_b200.sharded import Comm, ShardedEmbedding, lookup_many
    lib = _lib.load()
    comm = Comm.get(dev)
    st = _lib.stream_ptr

    def native(qq, kk, mode, need):
        Dn = torch.empty((qq.shape[0], kk), dtype=torch.float32, device=dev)
        In = torch.empty((qq.shape[0], kk), dtype=torch.int64, device=dev)
        bq, bD, bI, bN = _lib.dl(qq), _lib.dl(Dn), _lib.dl(In), _lib.dl(need)
        _lib.check(lib.rs_index_search_sharded(sidx.local._h, comm._h, N, bq.ptr, kk, _lib.mode_code(mode), bD.ptr, bI.ptr,
                                               bN.ptr if need is not None else None, st()))
        return Dn, In

    need = torch.full((nq,), 7, dtype=torch.int32, device=dev)
    for mode in ("auto", "exact"):
        Dn, In = native(q, k, mode, need)
        assert torch.equal(In, Iw) and torch.equal(Dn, Dw) and int(need.abs().sum()) == 0, f"rank {rank}: rs_index_search_sharded({mode}) differs"
    Dn, In = native(q, k, "auto", None)                              # the library resolves flags itself
    assert torch.equal(In, Iw) and torch.equal(Dn, Dw)
    os.environ["RS_DEBUG_FLAG_QUERIES"] = "3,100,299"               # test hook: treat these queries as flagged
    Dn, In = native(q, k, "auto", None)
    del os.environ["RS_DEBUG_FLAG_QUERIES"]
    assert torch.equal(In, Iw) and torch.equal(Dn, Dw), f"rank {rank}: library-side repair differs"
    assert sidx.local.fallbacks()[0] == 3, sidx.local.fallbacks()
    Dn, In = native(q[:5].contiguous(), 17, "auto", None)            # a second (nq, k): new regions, collectively
    assert torch.equal(In, Iw[:5, :17]) and torch.equal(Dn, Dw[:5, :17])
    checks.append("rs_index_search_sharded: auto / exact / library-side repair / second shape")

    # sharded embedding tables: rows == host replay of the hash-generated table, skewed ranks, two tables concurrently
    import time
    import numpy as np
    rng = np.random.default_rng(100 + rank)
    tabs = []
    for n_rows, seed in ((300_007, 5), (70_001, 6)):
        t = ShardedEmbedding(n_rows, 128, local_table=ShardedEmbedding.fill_synthetic(n_rows, 128, rank, world, seed, dev))
        t.enable_peer_exchange(4096, dedup=seed == 5)
        tabs.append((t, n_rows, seed))
    for it in range(12):
        if it % 3 == rank % 3:
            time.sleep(0.02 * (1 + rank))                            # ranks drift apart: waits must cover it
        pairs, want = [], []
        for t, n_rows, seed in tabs:
            B = int(rng.integers(1, 4097))
            ids = rs.synth.zipf_ids(rng, n_rows, B).astype(np.int64)
            pairs.append((t, torch.from_numpy(ids).to(dev)))
            want.append(rs.synth.hash_table_rows(ids, 128, seed=seed))
        got = lookup_many(pairs) if it % 2 else [t.lookup(i) for t, i in pairs]
        for g_, w_ in zip(got, want):
            assert np.array_equal(g_.cpu().numpy(), w_), f"rank {rank}: sharded lookup differs from the host replay (iteration {it})"
    for t, _, _ in tabs:
        t.close()
    checks.append("rs_embedding_a2a_lookup: 12 skewed iterations x 2 tables (sequential and concurrent) == host replay")
    comm.release()
    sidx.close()
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU parity OK on {world} GPUs: sharded == single-GPU, bit-exact (N={N}, nq={nq}, k={k})", flush=True)
        import json
        out = Path(__file__).resolve().parent.parent / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / f"r02_multi_gpu_check_{world}gpu.json").write_text(json.dumps(
            {"n_gpus": world, "catalog_rows": N, "nq": nq, "k": k, "bit_identical_to_single_gpu": True, "checks": checks,
             "device": torch.cuda.get_device_name(dev)}, indent=1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
