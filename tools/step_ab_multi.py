#!/usr/bin/env python
"""A/B of the sharded search step (10M x 128 catalog, 1,024 queries, top-100; BASELINE configs[2]) under torchrun:
the lists / tops published by the kernels that build them (7 launches) vs separate publish launches (9), copy-out vs
graph-owned outputs.
Device ms per step, max over ranks, CUDA-graph replay.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/step_ab_multi.py
"""
import json
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
    N, nq, k = 10_000_000, 1024, 100
    base = ShardedIndexFlatIP(128, N)
    for b0 in range(base.row_begin, base.row_end, 1 << 20):
        n = min(1 << 20, base.row_end - b0)
        g = torch.Generator(device=dev).manual_seed(3000 + b0)
        x = torch.randn((n, 128), generator=g, device=dev)
        rs.normalize_L2(x)
        base.add_local(x)
    gq = torch.Generator(device=dev).manual_seed(3999)
    q = torch.randn((nq, 128), generator=gq, device=dev)
    rs.normalize_L2(q)
    res = {}
    ref = None
    for name, env, reuse in (("fused_publish", "1", True), ("separate_publish", "0", True), ("fused_publish_copy_out", "1", False),
                             ("fused_publish_again", "1", True)):
        os.environ["RS_PEER_FUSED_PUBLISH"] = env
        sidx = ShardedIndexFlatIP(128, N, local_index=base.local)
        sidx.exchange, sidx.use_graph, sidx.reuse_outputs = "p2p", True, reuse
        for _ in range(30):
            D, I = sidx.search(q, k)
        if ref is None:
            ref = (D.clone(), I.clone())
        assert torch.equal(I, ref[1]) and torch.equal(D, ref[0])
        times = []
        for rep in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100):
                sidx.search(q, k)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 100], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(round(float(t.item()), 4))
        res[name] = times
        sidx._graphs.clear()
        torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"n_gpus": world, "ms_per_step": res}, indent=1), flush=True)
        out = Path(__file__).resolve().parent.parent / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / f"r02_step_ab2_{world}gpu.json").write_text(json.dumps({"n_gpus": world, "ms_per_step": res}, indent=1))
    dist.barrier()
    base.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
