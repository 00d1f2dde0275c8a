#!/usr/bin/env python
"""Device time of tc_filter_kernel (10M x 128 catalog, 1,024 queries, top-100) under the RS_TC_DEBUG ablation knobs
(4 = no survivors at all, 8 = survivors queued but dropped at the flush; results are then WRONG -- timing experiments only)."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent

def child():
    import torch
    sys.path.insert(0, str(ROOT))
    import recsys_b200 as rs
    dev = torch.device("cuda", 0)
    N = int(os.environ.get("N_ROWS", 10_000_000))
    index = rs.IndexFlatIP(128)
    for b0 in range(0, N, 1 << 20):
        n = min(1 << 20, N - b0)
        g = torch.Generator(device=dev).manual_seed(3000 + b0)
        x = torch.randn((n, 128), generator=g, device=dev)
        rs.normalize_L2(x)
        index.add(x)
    gq = torch.Generator(device=dev).manual_seed(3999)
    q = torch.randn((1024, 128), generator=gq, device=dev)
    rs.normalize_L2(q)
    for _ in range(3):
        index.search_deferred(q, 100, mode="tensor")   # no repair of flagged queries: the knobs flag all of them
    torch.cuda.synchronize()
    rs._lib.prof_enable(True)
    n_it = 20
    for _ in range(n_it):
        index.search_deferred(q, 100, mode="tensor")
    torch.cuda.synchronize()
    out = {}
    for name, slot in (("filter", rs._lib.PROF_FILTER), ("rescore", rs._lib.PROF_RESCORE), ("sample", rs._lib.PROF_SAMPLE),
                       ("threshold", rs._lib.PROF_THRESHOLD)):
        ms, n = rs._lib.prof_read(slot)
        out[name + "_ms"] = round(ms / n_it, 4)
    rs._lib.prof_enable(False)
    print(json.dumps(out))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for dbg in os.environ.get("KNOBS", "0,4,8").split(","):
            r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, RS_TC_DEBUG=dbg), capture_output=True, text=True)
            print(f"RS_TC_DEBUG={dbg}", r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
        for dbg in [x for x in os.environ.get("SEL_KNOBS", "").split(",") if x]:   # select_rescore_kernel: 1 = no re-score, 2 = no rank loop
            r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, RS_SEL_DEBUG=dbg), capture_output=True, text=True)
            print(f"RS_SEL_DEBUG={dbg}", r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
