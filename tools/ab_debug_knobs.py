#!/usr/bin/env python
"""Timing of the bf16 ranking forward at B=65,536 under the RS_CHAIN_DEBUG ablation knobs of tower_sp_kernel
(16 = no fp32 output, 64 = no interleaved bf16 output; results are then WRONG -- timing experiments only)."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent

def child():
    import torch
    sys.path.insert(0, str(ROOT))
    import recsys_b200 as rs
    from tools.microbench import timeit
    dev = torch.device("cuda", 0)
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, 6040, 3706, seed=10)
    model = rs.MultiTaskModel(cfg, rs.synth.make_vocab(6040), rs.synth.make_vocab(3706), {}, precision="bf16")
    rs.load_weights(model, w)
    B = 65536
    g = torch.Generator(device=dev).manual_seed(B)
    ur = torch.randint(0, 6041, (B,), device=dev, generator=g)
    ir = torch.randint(0, 3707, (B,), device=dev, generator=g)
    model({"user_id": ur, "movie_id": ir})
    rs._lib.prof_enable(True)
    for _ in range(20):
        model({"user_id": ur, "movie_id": ir})
    torch.cuda.synchronize()
    k_ms, k_n = rs._lib.prof_read(rs._lib.PROF_RANK)
    d_ms, d_n = rs._lib.prof_read(rs._lib.PROF_DEEP)
    rs._lib.prof_enable(False)
    sess = model.ranking_session(B)
    sess.user_rows.copy_(ur); sess.item_rows.copy_(ir)
    print(json.dumps({"tower_ms": round(k_ms / 20, 4), "deep_ms": round(d_ms / 20, 4), "graph_ms": round(timeit(sess.run, 50), 4)}))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for dbg in ("0", "16", "64", "80"):
            for sp in ("2", "1"):
                r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, RS_CHAIN_DEBUG=dbg, RS_TOWER_SP=sp), capture_output=True, text=True)
                print("RS_CHAIN_DEBUG=" + dbg, "RS_TOWER_SP=" + sp, r.stdout.strip().splitlines()[-1] if r.returncode == 0 else r.stderr[-1500:], flush=True)
