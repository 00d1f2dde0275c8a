#!/usr/bin/env python
"""One small launch of every kernel family, for compute-sanitizer:
    python tools/sanitize_smoke.py && compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402

dev = torch.device("cuda", 0)
xb = torch.from_numpy(rs.synth.make_unit_rows(70_001, 128, 1)).to(dev)
q = torch.from_numpy(rs.synth.make_unit_rows(130, 128, 2)).to(dev)
index = rs.IndexFlatIP(128)
index.add(xb)
De, Ie = index.search(q, 10, mode="exact")
Dt, It = index.search(q, 10, mode="tensor")
assert torch.equal(Ie, It) and torch.equal(De, Dt)
D1, I1 = index.search(q[:1].contiguous(), 100, mode="tensor")
Dm, Im = rs.ops.topk_merge(torch.stack([De, De]), torch.stack([Ie, Ie + 70_001]))
s = index.debug_tensor_scores(q[:3].contiguous())
cfg = rs.ModelConfig()
w = rs.synth.make_weights(cfg, 300, 200, seed=5)
ids_u = torch.from_numpy(rs.synth.make_ids(np.random.default_rng(1), 300, 300)).to(dev)
ids_i = torch.from_numpy(rs.synth.make_ids(np.random.default_rng(2), 200, 300)).to(dev)
outs = {}
for precision in ("fp32", "bf16"):
    m = rs.MultiTaskModel(cfg, rs.synth.make_vocab(300), rs.synth.make_vocab(200), {}, precision=precision)
    rs.load_weights(m, w)
    outs[precision] = m({"user_id": ids_u, "movie_id": ids_i})
    m.encoder({"user_id": ["1", "zz"]})
err = (outs["fp32"]["rating_prediction"] - outs["bf16"]["rating_prediction"]).abs().max().item()
assert err < 1e-2
x = torch.randn((300, 128), device=dev).to(torch.bfloat16)
wt = torch.randn((64, 128), device=dev).to(torch.bfloat16)
rs.ops.dense_bf16(x, wt, None, "relu")
from recsys_b200.sharded import ShardedEmbedding  # noqa: E402
emb = ShardedEmbedding(1000, 128, local_table=ShardedEmbedding.fill_synthetic(1000, 128, 0, 1, 7, dev))
emb.lookup(torch.randint(0, 1000, (500,), device=dev))
torch.cuda.synchronize()
print("sanitize smoke OK; kernels launched:", rs._lib.launch_count())
