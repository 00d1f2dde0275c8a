#!/usr/bin/env python
"""A/B of the bf16 ranking forward: deep+cross+heads fused (default) vs three kernels (RS_FUSED_HEADS=0).
Prints device time per batch and the largest difference between the two paths' outputs."""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, 6040, 3706, seed=10)
    model = rs.MultiTaskModel(cfg, rs.synth.make_vocab(6040), rs.synth.make_vocab(3706), {}, precision="bf16")
    rs.load_weights(model, w)
    for B in (1, 100, 4096, 65536, 262144):
        ur = torch.randint(0, 6041, (B,), device=dev)
        ir = torch.randint(0, 3707, (B,), device=dev)
        res = {}
        for fused in ("1", "0"):
            os.environ["RS_FUSED_HEADS"] = fused
            out = model({"user_id": ur, "movie_id": ir})
            torch.cuda.synchronize()
            ms = timeit(lambda: model({"user_id": ur, "movie_id": ir}), 20)
            res[fused] = (ms, out["rating_prediction"].clone(), out["ctr_prediction"].clone())
        dr = (res["1"][1] - res["0"][1]).abs().max().item()
        dc = (res["1"][2] - res["0"][2]).abs().max().item()
        print(json.dumps({"B": B, "fused_ms": round(res["1"][0], 4), "unfused_ms": round(res["0"][0], 4),
                          "fused_samples_per_s": round(B / res["1"][0] * 1e3),
                          "max_abs_diff_rating": dr, "max_abs_diff_ctr": dc,
                          "rating_scale": res["0"][1].abs().max().item()}), flush=True)


if __name__ == "__main__":
    main()
