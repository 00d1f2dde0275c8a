#!/usr/bin/env python
"""A/B of the round-2 ranking kernels against the round-1 ones (one subprocess per setting: the knobs are read once per
process).  Settings: RS_TOWER_SP (2 = tower_sp_kernel with a team of warps per slot, 1 = with the shared event order, 0 = tower_pp_kernel), RS_DEEP_V2 (pipelined deep_heads kernel vs the
round-1 one).  Prints device ms per bf16 ranking forward (eager and CUDA-graph replay), per-kernel ms from the library's
event brackets, and checks that every setting produces bit-identical outputs (same arithmetic, different scheduling).

    python tools/ab_ranker_r02.py [--trace]      # --trace: also one RS_CHAIN_TRACE=1 forward per setting (stderr)
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SETTINGS = [("sp2_deep1", {"RS_TOWER_SP": "2", "RS_DEEP_V2": "1"}),
            ("sp1_deep1", {"RS_TOWER_SP": "1", "RS_DEEP_V2": "1"}),
            ("sp1_deep0", {"RS_TOWER_SP": "1", "RS_DEEP_V2": "0"}),
            ("sp0_deep1", {"RS_TOWER_SP": "0", "RS_DEEP_V2": "1"}),
            ("sp0_deep0", {"RS_TOWER_SP": "0", "RS_DEEP_V2": "0"}),
            ("pdl1", {"RS_PDL": "1"}), ("pdl0", {"RS_PDL": "0"})]   # programmatic dependent launch of the deep kernel on / off


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
    res = {}
    sizes = (1, 127, 1000, 4096, 65536, 65537, 262144) if not os.environ.get("RS_CHAIN_TRACE") else (65536,)
    for B in sizes:
        g = torch.Generator(device=dev).manual_seed(B)
        ur = torch.randint(0, 6041, (B,), device=dev, generator=g)
        ir = torch.randint(0, 3707, (B,), device=dev, generator=g)
        out = model({"user_id": ur, "movie_id": ir})
        torch.cuda.synchronize()
        if os.environ.get("RS_CHAIN_TRACE"):
            continue
        chk = [float(out[k].double().sum()) for k in ("user_embedding", "item_embedding", "rating_prediction", "ctr_prediction")]
        chk += [float(out[k].double().abs().sum()) for k in ("user_embedding", "item_embedding", "rating_prediction", "ctr_prediction")]
        chk += [float(out[k].double().abs().max()) for k in ("rating_prediction", "ctr_prediction")]
        ms = timeit(lambda: model({"user_id": ur, "movie_id": ir}), 30)
        rs._lib.prof_enable(True)
        for _ in range(10):
            model({"user_id": ur, "movie_id": ir})
        torch.cuda.synchronize()
        k_ms, k_n = rs._lib.prof_read(rs._lib.PROF_RANK)
        d_ms, d_n = rs._lib.prof_read(rs._lib.PROF_DEEP)
        rs._lib.prof_enable(False)
        sess = model.ranking_session(B)
        sess.user_rows.copy_(ur)
        sess.item_rows.copy_(ir)
        gms = timeit(sess.run, 50)
        res[B] = {"ms": round(ms, 4), "graph_ms": round(gms, 4), "tower_ms": round(k_ms / 10, 4), "deep_ms": round(d_ms / 10, 4), "launches": (k_n + d_n) // 10,
                  "graph_samples_per_s": round(B / gms * 1e3), "gemm_tflops": round(524288.0 * B / gms / 1e9, 1), "checksums": chk}
    print(json.dumps(res))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        return child()
    trace = "--trace" in sys.argv
    outs = {}
    only = os.environ.get("AB_SETTINGS")
    settings = [x for x in SETTINGS if not only or x[0] in only.split(",")]
    for name, env_add in settings:
        env = dict(os.environ, **env_add)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(name, "FAILED\n", r.stdout[-2000:], r.stderr[-4000:])
            sys.exit(1)
        outs[name] = json.loads(r.stdout.strip().splitlines()[-1])
        print(name, json.dumps({b: {k: v for k, v in d.items() if k != "checksums"} for b, d in outs[name].items()}), flush=True)
        if trace:
            r = subprocess.run([sys.executable, __file__, "child"], env=dict(env, RS_CHAIN_TRACE="1"), capture_output=True, text=True)
            print(f"--- trace {name}\n" + r.stderr[-6000:], flush=True)
    base = outs[settings[-1][0]]
    # different kernels order their fp32 additions differently (bias first vs last, packed dots): checksums agree to ~1e-6
    rel = max(abs(outs[m][b]["checksums"][i] - base[b]["checksums"][i]) / max(1e-30, abs(base[b]["checksums"][i]))
              for m in outs for b in outs[m] for i in range(len(base[b]["checksums"])))
    print("largest relative difference of an output checksum between the settings:", rel)
    sys.exit(0 if rel < 1e-4 else 2)


if __name__ == "__main__":
    main()
