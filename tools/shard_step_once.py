"""One sharded-search step (nq = 1,024, top-100) over a 1.25M-row shard behind a communicator of one rank -- the per-GPU
kernel sequence of the 8-way split 10M catalog (with ~8x the candidates per query: the threshold comes from this shard
alone) -- for ncu launch lists."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402
from recsys_b200.sharded import ShardedIndexFlatIP  # noqa: E402

dev = torch.device("cuda", 0)
N = 1_250_000
g = torch.Generator(device=dev).manual_seed(1)
xb = torch.randn((N, 128), generator=g, device=dev)
rs.normalize_L2(xb)
index = rs.IndexFlatIP(128)
index.add(xb)
q = torch.randn((1024, 128), generator=g, device=dev)
rs.normalize_L2(q)
sidx = ShardedIndexFlatIP(128, N, local_index=index)
sidx.world = 1
sidx.exchange = "p2p"
for _ in range(3):
    D, I, need = sidx._native_step(q, 100)
    torch.cuda.synchronize()
Dw, Iw = index.search(q, 100, mode="exact")
assert torch.equal(I, Iw) and torch.equal(D, Dw) and not bool(need.any())
if len(sys.argv) > 1 and sys.argv[1] == "time":   # device ms per step (eager launches; the step is GPU-bound at this size)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            sidx._native_step(q, 100)
        e1.record()
        torch.cuda.synchronize()
        print("ms per step: %.4f" % (e0.elapsed_time(e1) / 100))
