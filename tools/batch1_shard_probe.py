#!/usr/bin/env python
"""What one GPU of an 8-way sharded 10M-row catalog does for a batch-1 top-10 query (BASELINE configs[4] at G = 8): a
1.25M-row shard behind a communicator of one rank (the whole protocol runs, exchanges included, without NVLink hops).
Device ms per query for the tensor path (shared threshold), the exact fp32 scan, both through rs_index_search_sharded, eager
and as a CUDA-graph replay.

    python tools/batch1_shard_probe.py
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402
from recsys_b200 import _lib  # noqa: E402
from recsys_b200.sharded import Comm  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    N, k = 1_250_000, 10
    g = torch.Generator(device=dev).manual_seed(1)
    xb = torch.randn((N, 128), generator=g, device=dev)
    rs.normalize_L2(xb)
    index = rs.IndexFlatIP(128)
    index.add(xb)
    comm = Comm.get(dev)
    lib = _lib.load()
    res = {}
    for nq in (1, 8):
        q = torch.randn((nq, 128), generator=g, device=dev)
        rs.normalize_L2(q)
        D = torch.empty((nq, k), device=dev)
        I = torch.empty((nq, k), dtype=torch.int64, device=dev)
        need = torch.empty((nq,), dtype=torch.int32, device=dev)
        bq, bD, bI, bN = _lib.dl(q), _lib.dl(D), _lib.dl(I), _lib.dl(need)

        def call(mode):
            _lib.check(lib.rs_index_search_sharded(index._h, comm._h, N, bq.ptr, k, _lib.mode_code(mode), bD.ptr, bI.ptr, bN.ptr,
                                                   _lib.stream_ptr()))
        for mode in ("auto", "exact"):
            call(mode)
            torch.cuda.synchronize()
            want = index.search(q, k, mode="exact")
            assert torch.equal(I, want[1]) and torch.equal(D, want[0])
            eager = timeit(lambda: call(mode), 50)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                call(mode)
            rep = timeit(gr.replay, 100)
            res[f"nq{nq}_{mode}"] = {"eager_ms": round(eager, 4), "graph_ms": round(rep, 4)}
            del gr
        res[f"nq{nq}_plain_index_search_auto_ms"] = round(timeit(lambda: index.search(q, k), 50), 4)
        res[f"nq{nq}_plain_index_search_exact_ms"] = round(timeit(lambda: index.search(q, k, mode="exact"), 50), 4)
    comm.release()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
