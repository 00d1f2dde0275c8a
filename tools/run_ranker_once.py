#!/usr/bin/env python
"""A few bf16-mode ranking forwards at one batch size (for ncu captures of the chain kernels)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
cfg = rs.ModelConfig()
w = rs.synth.make_weights(cfg, 6040, 3706, seed=10)
model = rs.MultiTaskModel(cfg, rs.synth.make_vocab(6040), rs.synth.make_vocab(3706), {}, precision="bf16")
rs.load_weights(model, w)
ur = torch.randint(0, 6041, (B,), device=dev)
ir = torch.randint(0, 3707, (B,), device=dev)
for _ in range(4):
    out = model({"user_id": ur, "movie_id": ir})
torch.cuda.synchronize()
print("ok", float(out["rating_prediction"].sum()))
