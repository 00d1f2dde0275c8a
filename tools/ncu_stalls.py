#!/usr/bin/env python
"""Top warp-stall sites of one kernel from an ncu report (--set full --import-source on):
    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep <kernel regex> [N]"""
import csv, io, subprocess, sys

def main():
    rep, rx = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    data = []
    for r in rows[hi + 1:]:
        if r and r[0] == "Address":
            break          # next kernel instance
        if len(r) == len(hdr):
            data.append(r)
    ia, isrc, iall, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda x: int(x) if x.strip().isdigit() else 0
    tot = sum(num(r[iall]) for r in data)
    print(f"# {rx}: {len(data)} instructions, {tot} stall samples")
    agg = {}
    for r in data:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + num(r[i])
    print("# by reason:", ", ".join(f"{k[6:]} {v / max(1, tot):.1%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -num(r[iall]))[:n]:
        st = sorted(((hdr[i][6:], num(r[i])) for i in stall_cols if num(r[i]) > 0), key=lambda kv: -kv[1])[:3]
        print(f"{r[ia][-5:]} {num(r[iall]) / max(1, tot):6.1%} exec={r[iex]:>8} {r[isrc][:64]:64s} {st}")

if __name__ == "__main__":
    main()
