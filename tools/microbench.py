#!/usr/bin/env python
"""Device-time micro-benchmarks of the retrieval and ranking kernels (CUDA events, warm-up, per-shape lines).

    python tools/microbench.py [--n-items 10000000] [--modes exact,auto]

Prints one JSON line per (nq, k, mode) with ms, queries/s, effective GB/s over the fp32 catalog and TFLOP/s.
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-items", type=int, default=10_000_000)
    ap.add_argument("--modes", default="exact")
    ap.add_argument("--shapes", default="1:10,1:100,8:10,8:100,64:100,1024:100")
    ap.add_argument("--ranking", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    N, d = args.n_items, 128
    index = rs.IndexFlatIP(d)
    blk = 1 << 20
    for b0 in range(0, N if args.shapes else 0, blk):
        g = torch.Generator(device=dev)
        g.manual_seed(3000 + b0 // blk)
        x = torch.randn((min(blk, N - b0), d), generator=g, device=dev)
        rs.normalize_L2(x)
        index.add(x)
    for shape in (args.shapes.split(",") if args.shapes else []):
        nq, k = map(int, shape.split(":"))
        q = torch.randn((nq, d), device=dev)
        rs.normalize_L2(q)
        for mode in args.modes.split(","):
            try:
                rs._lib.prof_enable(True)
                iters = 20 if nq <= 64 else 5
                ms = timeit(lambda: index.search(q, k, mode=mode), iters)
                prof = {n: rs._lib.prof_read(s) for n, s in (("scan", 0), ("merge", 1), ("filter", 2), ("rescore", 3), ("sample", 5))}
                rs._lib.prof_enable(False)
                print(json.dumps({"nq": nq, "k": k, "mode": mode, "ms": round(ms, 4), "qps": round(nq / ms * 1e3, 1),
                                  "catalog_GBps": round(N * d * 4 / ms / 1e6, 1),
                                  "TFLOPs": round(2.0 * nq * N * d / ms / 1e9, 2),
                                  "kernels_ms_per_call": {n: round(v[0] / max(1, v[1]) * (v[1] / (iters + 3)), 4) for n, v in prof.items() if v[1]}}),
                      flush=True)
            except RuntimeError as e:
                print(json.dumps({"nq": nq, "k": k, "mode": mode, "error": str(e)[:200]}), flush=True)
    if args.ranking:
        cfg = rs.ModelConfig()
        w = rs.synth.make_weights(cfg, 6040, 3706, seed=10)
        for precision in ("fp32", "bf16"):
            model = rs.MultiTaskModel(cfg, rs.synth.make_vocab(6040), rs.synth.make_vocab(3706), {}, precision=precision)
            rs.load_weights(model, w)
            for B in (4096, 65536, 262144):
                ur = torch.randint(0, 6041, (B,), device=dev)
                ir = torch.randint(0, 3707, (B,), device=dev)
                rs._lib.prof_enable(True)
                ms = timeit(lambda: model({"user_id": ur, "movie_id": ir}), 10)
                tc_ms, tc_n = rs._lib.prof_read(4)
                rs._lib.prof_enable(False)
                print(json.dumps({"ranking_batch": B, "precision": precision, "ms": round(ms, 4),
                                  "samples_per_s": round(B / ms * 1e3), "gemm_TFLOPs": round(524288.0 * B / ms / 1e9, 2),
                                  "dense_tc_ms_per_call": round(tc_ms / 13, 4), "dense_tc_launches_per_call": tc_n / 13}),
                      flush=True)


if __name__ == "__main__":
    main()
