#!/usr/bin/env python
"""Exact top-10 over the config-1 sized catalog (3,706 x 128) for several query counts: device ms per search.
The exact fp32 scan (with the small-catalog grid rule of rs::search_exact_device) against the small-catalog tensor path
(tc_filter.cu: every tile scored twice on tcgen05, chunk-maximum threshold, exact re-score) -- same results."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402
from tools.microbench import timeit  # noqa: E402

xb = torch.from_numpy(rs.synth.make_unit_rows(3706, 128, 1)).cuda()
q = torch.from_numpy(rs.synth.make_unit_rows(4096, 128, 2)).cuda()
idx = rs.IndexFlatIP(128)
idx.add(xb)
for nq in (4096, 256, 64, 8, 1):
    qq = q[:nq].contiguous()
    print(nq, "queries: exact scan", round(timeit(lambda: idx.search(qq, 10, mode="exact"), 20), 4), "ms; tensor path (auto)",
          round(timeit(lambda: idx.search(qq, 10), 20), 4), "ms", flush=True)
