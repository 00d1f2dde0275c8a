#!/usr/bin/env python
"""Timing of the fp32 ranking mode: the tcgen05 fp16-split kernel (dense_x3.cu, default) against the FFMA kernel
(RS_FP32_TENSOR=0), whole MultiTaskModel forward and layer by layer, CUDA events, batch 65,536 (BASELINE configs[1]).

    python tools/time_fp32_mode.py
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    B = int(os.environ.get("B", 65536))
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, 6040, 3706, seed=10)
    model = rs.MultiTaskModel(cfg, rs.synth.make_vocab(6040), rs.synth.make_vocab(3706), {})
    rs.load_weights(model, w)
    g = torch.Generator(device=dev).manual_seed(1)
    feats = {"user_id": torch.randint(0, 6041, (B,), device=dev, generator=g),
             "movie_id": torch.randint(0, 3707, (B,), device=dev, generator=g)}
    res = {"batch": B}
    for name, env in (("tensor", "1"), ("ffma", "0")):
        os.environ["RS_FP32_TENSOR"] = env
        model(feats)
        ms = timeit(lambda: model(feats), 20)
        res[f"forward_ms_{name}"] = round(ms, 4)
        res[f"samples_per_s_{name}"] = round(B / ms * 1e3)
        # the same forward as one CUDA graph: the eager call's host work (7-10 launches with their DLPack marshalling) is
        # longer than the kernels at this batch
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = model(feats)
        gms = timeit(g.replay, 30)
        res[f"graph_ms_{name}"] = round(gms, 4)
        res[f"graph_samples_per_s_{name}"] = round(B / gms * 1e3)
        del g, out
    os.environ["RS_FP32_TENSOR"] = "1"
    layers = {}
    rng = np.random.default_rng(0)
    for K, N in ((128, 256), (256, 128), (128, 64), (64, 128), (256, 256)):
        x = torch.randn((B, K), device=dev)
        wk = torch.from_numpy(rs.synth.glorot_uniform(rng, K, N)).to(dev)
        b = torch.zeros((N,), device=dev)
        hi, lo = rs.ops.dense_x3_prepare(wk)
        y = torch.empty((B, N), device=dev)
        t3 = timeit(lambda: rs.ops.dense_x3(x, hi, lo, b, "relu", out=y), 30)
        tf = timeit(lambda: rs.ops.dense(x, wk, b, "relu", out=y), 10)
        gb = B * (K + N) * 4 / 1e9
        layers[f"{K}x{N}"] = {"x3_ms": round(t3, 4), "ffma_ms": round(tf, 4), "x3_GBps": round(gb / t3 * 1e3, 1),
                              "x3_tensor_TFLOPs": round(3 * 2 * B * K * N / t3 / 1e9, 1)}
    res["layers"] = layers
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
