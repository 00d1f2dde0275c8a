#!/usr/bin/env python
"""Which kernels of the built librecsys_b200.so really contain Blackwell tensor-core / TMA instructions: counts of the
SASS mnemonics `UTC*MMA` (tcgen05.mma), `LDTM` / `STTM` (tcgen05.ld / st), `UTMALDG` / `UTMASTG` (TMA tensor load /
store) and `UBLKCP` (cp.async.bulk) per kernel, from `cuobjdump -sass` (runs without a GPU).

    python tools/sass_evidence.py > profiles/rNN_sass_mnemonics.txt
"""
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "recommendation-system-maang-nvidia-_b200" / "librecsys_b200.so"
PAT = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP)\b")


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:  # noqa: BLE001
        return {n: n for n in names}


def collect(lib=LIB):
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    counts, name = defaultdict(lambda: defaultdict(int)), None
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
        elif name:
            m = PAT.search(line)
            if m:
                counts[name][m.group(1)] += 1
    pretty = demangle(list(counts))
    out = {}
    for k, v in counts.items():
        short = re.sub(r"\(anonymous namespace\)::", "", pretty[k]).split("(")[0].replace("void ", "")
        out[short] = dict(v)
    return out


if __name__ == "__main__":
    res = collect(Path(sys.argv[1]) if len(sys.argv) > 1 else LIB)
    print("# cuobjdump -sass librecsys_b200.so: tcgen05 / TMA mnemonics per kernel (count of static instructions)")
    for k in sorted(res):
        print(f"{k:60s} " + "  ".join(f"{m}={n}" for m, n in sorted(res[k].items())))
