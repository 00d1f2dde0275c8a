#!/usr/bin/env python
"""A/B of tower_pp_kernel's data-movement modes (RS_TOWER_STAGE): 2 = staged gather + TMA stores (default),
1 = staged gather + coalesced STG, 0 = direct thread-per-row loads/stores.  One subprocess per setting (the knob
is read once per process).  Prints device ms per bf16 ranking forward and checks that all settings produce
bit-identical outputs (same arithmetic, only the data movement differs)."""
import json
import os
import subprocess
import sys
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
    res = {}
    for B in (1000, 65536, 262144):
        g = torch.Generator(device=dev).manual_seed(B)
        ur = torch.randint(0, 6041, (B,), device=dev, generator=g)
        ir = torch.randint(0, 3707, (B,), device=dev, generator=g)
        out = model({"user_id": ur, "movie_id": ir})
        torch.cuda.synchronize()
        ms = timeit(lambda: model({"user_id": ur, "movie_id": ir}), 30)
        chk = [float(out[k].double().sum()) for k in ("user_embedding", "item_embedding", "rating_prediction", "ctr_prediction")]
        sess = model.ranking_session(B)
        sess.user_rows.copy_(ur)
        sess.item_rows.copy_(ir)
        gms = timeit(sess.run, 30)
        res[B] = {"ms": round(ms, 4), "graph_ms": round(gms, 4), "graph_samples_per_s": round(B / gms * 1e3), "checksums": chk}
    print(json.dumps(res))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        return child()
    outs = {}
    for stage in ("2", "1", "0"):
        env = dict(os.environ, RS_TOWER_STAGE=stage)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stderr[-3000:])
            sys.exit(1)
        outs[stage] = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"RS_TOWER_STAGE={stage}", json.dumps(outs[stage]), flush=True)
    same = all(outs[m][b]["checksums"] == outs["0"][b]["checksums"] for m in outs for b in outs[m])
    print("outputs identical between the settings:", same)
    sys.exit(0 if same else 2)


if __name__ == "__main__":
    main()
