#!/usr/bin/env python
"""BASELINE config 4: Zipf-skewed lookups into embedding tables sharded over the GPUs (row mod G), all-to-all
exchanges feeding the DCN ranking forward -- timed three ways in one run: NCCL all-to-alls (host-sized splits),
peer-memory exchange (csrc/peer.cu) without and with per-batch de-duplication.  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29520 \
        tools/bench_config4.py [--users 50000000 --items 10000000 --batch 65536]

Checks (every rank): looked-up rows are bit-identical to the host replay of the synthetic table; the ranking
forward fed with the exchanged rows equals the forward over an unsharded copy (small tables only).  Then times
lookup and lookup+ranking with CUDA events (max over ranks) and prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402
from recsys_b200.sharded import ShardedEmbedding, lookup_many  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=50_000_000)
    ap.add_argument("--items", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    E, B = 128, args.batch
    tables = {}
    for name, V, seed in (("user", args.users, 41), ("item", args.items, 42)):
        shard = ShardedEmbedding.fill_synthetic(V, E, rank, world, seed, dev)
        tables[name] = ShardedEmbedding(V, E, local_table=shard)
    rng = np.random.default_rng(4100 + rank)
    ids = {"user": rs.synth.zipf_ids(rng, min(args.users, 5_000_000), B), "item": rs.synth.zipf_ids(rng, min(args.items, 5_000_000), B)}
    # spread the Zipf head over the whole id range (the generator's permutation covers the first 5M ids)
    ids["user"] = (ids["user"].astype(np.int64) * 9) % args.users
    ids["item"] = (ids["item"].astype(np.int64) * 3) % args.items
    ids_d = {k_: torch.from_numpy(v).to(dev) for k_, v in ids.items()}

    # ---- parity: exchanged rows == host replay, bit for bit ----
    for name, seed in (("user", 41), ("item", 42)):
        got = tables[name].lookup(ids_d[name]).cpu().numpy()
        want = rs.synth.hash_table_rows(ids[name], E, seed)
        assert np.array_equal(got, want), f"rank {rank}: {name} rows differ from the host replay"

    # ---- ranking forward fed from the sharded lookups ----
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, 8, 8, seed=10)        # tower / DCN weights; the 9-row tables are placeholders
    model = rs.MultiTaskModel(cfg, rs.synth.make_vocab(8), rs.synth.make_vocab(8), {}, device=dev, precision="bf16")
    rs.load_weights(model, w)

    def step_lookup():       # peer-memory tables are looked up concurrently (two streams); NCCL tables one after the other
        return lookup_many([(tables["user"], ids_d["user"]), (tables["item"], ids_d["item"])])

    def step_full():
        u, it = step_lookup()
        return model.rank_from_embedding_rows(u, it)

    out = step_full()
    from oracle.models import MultiTaskModel as OMT   # checker only
    w2 = dict(w)
    w2["user_embedding"] = rs.synth.hash_table_rows(ids["user"][:512], E, 41)
    w2["item_embedding"] = rs.synth.hash_table_rows(ids["item"][:512], E, 42)
    ref = OMT(w2, [], [], np.float64).rank_rows(np.arange(512), np.arange(512))
    for key in ("rating_prediction", "ctr_prediction", "user_embedding"):
        got = out[key][:512].cpu().numpy().astype(np.float64)
        rel = np.abs(got - ref[key]).max() / np.abs(ref[key]).max()
        assert rel <= 1e-2, f"rank {rank}: {key} off by {rel:.2e} (bf16 tolerance 1e-2)"

    def timed(fn, steps):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    res = {}
    res["nccl"] = (timed(step_lookup, args.steps), timed(step_full, args.steps))
    for label, dedup in (("peer", False), ("peer_dedup", True)):
        for name, seed in (("user", 41), ("item", 42)):
            tables[name].close()
            tables[name].enable_peer_exchange(B, dedup=dedup)
            for _ in range(2):       # twice: the second pass re-uses the regions
                got = tables[name].lookup(ids_d[name]).cpu().numpy()
                assert np.array_equal(got, rs.synth.hash_table_rows(ids[name], E, seed)), f"rank {rank}: {name} rows differ ({label})"
        res[label] = (timed(step_lookup, args.steps), timed(step_full, args.steps))
    # the peer-memory step has no host synchronisation: capture lookups + ranking once, replay as one CUDA graph
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        step_full()
    torch.cuda.current_stream(dev).wait_stream(side)
    dist.barrier()
    torch.cuda.synchronize()
    g_lookup, g_full = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_lookup):
        rows_static = step_lookup()
    with torch.cuda.graph(g_full):
        out_static = step_full()
    g_full.replay()
    torch.cuda.synchronize()
    for key in ("rating_prediction", "ctr_prediction"):
        assert torch.equal(out_static[key], out[key]), f"rank {rank}: graph replay differs from the eager step ({key})"
    res["peer_dedup_graph"] = (timed(g_lookup.replay, args.steps), timed(g_full.replay, args.steps))
    del g_lookup, g_full, rows_static, out_static
    uniq = torch.tensor([float(len(np.unique(ids["user"])) + len(np.unique(ids["item"]))) / (2 * B)], device=dev, dtype=torch.float64)
    dist.all_reduce(uniq, op=dist.ReduceOp.SUM)
    for t in tables.values():
        t.close()
    if rank == 0:
        remote = (world - 1) / world
        line = {
            "workload": f"config 4: tables {args.users:,} x 128 + {args.items:,} x 128 sharded row mod {world}, Zipf(1.05) ids, "
                        f"local batch {B}, all-to-all lookups feeding the bf16 DCN ranking forward",
            "n_gpus": world, "parity": "looked-up rows bit-identical to host replay on all three paths; logits within 1e-2 of fp64 oracle",
            "distinct_id_fraction": round(float(uniq.item()) / world, 4),
            "nvlink_bytes_per_sample_each_way_no_dedup": round(2 * (8 + 4 * E) * remote, 1)}
        for label, (ms_lookup, ms_full) in res.items():
            line[label] = {"lookup_ms": round(ms_lookup, 4), "lookup_samples_per_s": round(world * B / ms_lookup * 1e3),
                           "lookup_plus_ranking_ms": round(ms_full, 4), "samples_per_s": round(world * B / ms_full * 1e3),
                           "nvlink_GBps_per_gpu_effective": round(2 * 4 * E * remote * B / ms_lookup / 1e6, 1)}
        print(json.dumps(line))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
