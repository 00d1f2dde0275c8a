"""Time (and, under ncu, profile) the device StringLookup alone: the `string_lookup` block of bench.py without the rest.

    python tools/string_lookup_once.py [--words 1048576] [--ids 4194304] [--steps 20]
    ncu --set full --clock-control none --import-source on -k regex:string_lookup -c 1 -o gpurun_out/k0 python tools/string_lookup_once.py --steps 1
"""

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
import recsys_b200 as rs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--words", type=int, default=1 << 20)
    ap.add_argument("--ids", type=int, default=1 << 22)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    out = bench.bench_string_lookup(rs, torch, device, timed, bench.peaks(), n_words=a.words, B=a.ids, steps=a.steps)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
