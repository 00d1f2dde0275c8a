"""One batch-1 top-10 search over a 1.25M-row index (a shard of the 8-way split 10M catalog) -- for ncu launch lists."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
xb = torch.randn((1_250_000, 128), generator=g, device="cuda")
rs.normalize_L2(xb)
index = rs.IndexFlatIP(128)
index.add(xb)
q = torch.randn((1, 128), generator=g, device="cuda")
rs.normalize_L2(q)
for mode in ("auto", "exact", "auto", "exact"):
    index.search(q, 10, mode=mode)
    torch.cuda.synchronize()
