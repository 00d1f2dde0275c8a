#!/usr/bin/env python
"""bench.py -- headline benchmark of the serving/scoring hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU path, timed on the host cores

Metric (BASELINE.json): queries/sec of exact top-100 retrieval over a 10M-item, d=128 catalog in
1,024-query batches (configs[2]; the catalog is 5.12 GB and fits one GPU, at N>1 it is row-sharded, the
per-shard lists are exchanged over NVLink and merged).  A "step" is one 1,024-query batch over the whole catalog.

One JSON line.  Top-level keys follow the driver's contract for configs[2]; every other BASELINE config rides in
the same line:
  value / ms_per_step   queries/s with the query batch already in HBM, timed WITHOUT the library's per-kernel event
                        brackets (those run in a second pass that feeds `roofline` / `kernels`)
  e2e                   N=1: through `IndexFlatIP.search(numpy array)` -- host query buffer in, host (D, I) out, the
                        staging copies inside the C call; N>1: pinned host queries up, sharded search, (D, I) down
  parity_check          results of the timed path compared bit for bit on this box: against the CPU oracle's exact
                        fp32 chain on a 1M-row slice (N=1) / against the exact-scan fallback path on every shard (N>1)
  ranking               configs[1]: MultiTaskModel forward at batch 65,536 per GPU (bf16 tensor mode + fp32 mode),
                        roofline, e2e, torch-CPU baseline
  batch1_latency        configs[4]: single-query top-10 latency p50/p99 (device, host round trip, and through
                        RecommendationService.recommend), roofline, CPU baseline
  config1               configs[0]: MovieLens-1M-shaped forward + exact top-10 (times only; parity is in tests/)
  config4               configs[3] (N>1): Zipf lookups into 50M + 10M row tables sharded row mod N over NVLink peer
                        memory, feeding the ranking forward; rows checked bit for bit against the host replay
  cpu_baseline          the CPU restatement (oracle.search_blas: blocked sgemm + heap, faiss-cpu's >20-query path)
                        on a bounded slice of the same workload, all host cores
Under --impl reference the CPU restatement is the timed arm (TensorFlow / faiss cannot be installed in this
image; kind = "port").
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "queries/sec exact top-100 over 10M items"
UNIT = "queries/s"
NVLINK_PEER_GBS = 770.0      # measured peer copy per direction per GPU on this pool (/opt/skills/guides/B200_PROFILING.md)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-items", type=int, default=10_000_000)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--mode", default="auto", choices=["auto", "exact", "tensor"])
    ap.add_argument("--rank-batch", type=int, default=65536)
    ap.add_argument("--no-ranking", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch1", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--c4-users", type=int, default=50_000_000)
    ap.add_argument("--c4-items", type=int, default=10_000_000)
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def host_threads() -> int:
    return os.cpu_count() or 1


class all_host_cores:
    """The CPU legs use every host core whatever OMP_NUM_THREADS says (torch.distributed.run exports OMP_NUM_THREADS=1,
    which would silently make the CPU arm ~3x slower than the same arm at N=1)."""

    def __enter__(self):
        try:
            import torch
            self._torch_threads = torch.get_num_threads()       # read BEFORE the limits below change libgomp's setting
        except Exception:  # noqa: BLE001
            self._torch_threads = None
        try:
            from threadpoolctl import threadpool_limits
            self._ctx = threadpool_limits(limits=host_threads())     # OpenBLAS (numpy) and libgomp (oracle/chain.c, torch)
            self._ctx.__enter__()
        except Exception:  # noqa: BLE001
            self._ctx = None
        if self._torch_threads is not None:
            import torch
            torch.set_num_threads(host_threads())
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        if self._torch_threads is not None:
            import torch
            torch.set_num_threads(self._torch_threads)
        return False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.t0 = index, [], None, None

    def prestart(self):
        """Spawn nvidia-smi early (it needs ~1 s before its first line); samples only count from start() on."""
        if self.proc is not None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def start(self):
        self.prestart()
        self.t0 = time.monotonic()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        rows = list(self.rows)
        under_load = [x for x in rows if self.t0 is None or x[0] >= self.t0]   # drop idle-clock samples from before start()
        for t, r in (under_load or rows[-3:]):
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm / baselines
def cpu_prepare(args, sample_rows: int):
    import oracle
    import recsys_b200 as rs
    oracle.build_chain_lib()
    rows = min(sample_rows, args.n_items)
    xb = rs.synth.make_unit_rows(rows, args.dim, 3000)
    q = rs.synth.make_unit_rows(args.nq, args.dim, 3999)
    with all_host_cores():
        oracle.search_blas(xb[:4096], q[:64], args.k)           # warm the BLAS threads
    return xb, q


def cpu_step(args, xb, q) -> float:
    """One bounded CPU step: nq queries x the catalog slice; returns seconds scaled to the full catalog."""
    import oracle
    with all_host_cores():
        t0 = time.perf_counter()
        oracle.search_blas(xb, q, args.k)
        dt = time.perf_counter() - t0
    return dt * (args.n_items / xb.shape[0])


def cpu_describe(args, rows, scaled_s):
    return {"value": args.nq / scaled_s, "unit": UNIT, "cores": host_threads(), "kind": "port",
            "sample": f"{args.nq} queries x {rows:,}-row slice ({rows / args.n_items:.0%} of the catalog), blocked "
                      f"OpenBLAS sgemm + per-query heap (faiss-cpu's >20-query path restated; TF/faiss are not "
                      f"installable here), time scaled x{args.n_items / rows:.1f} to the full catalog; thread count set "
                      f"explicitly to all {host_threads()} host cores"}


def cpu_batch1(args, xb, n_queries=8, k=10):
    """configs[4] on the host: one query at a time over the catalog slice (faiss-cpu scans sequentially below 20 queries;
    here the slice goes through one multi-threaded sgemv + a partial sort, the CPU's best case), scaled to 10M rows."""
    rng = np.random.default_rng(5555)
    qs = rng.standard_normal((n_queries, args.dim)).astype(np.float32)
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    ms = []
    with all_host_cores():
        for i in range(n_queries):
            t0 = time.perf_counter()
            s = xb @ qs[i]
            top = np.argpartition(-s, k)[:k]
            top[np.argsort(-s[top])]
            ms.append((time.perf_counter() - t0) * 1e3 * (args.n_items / xb.shape[0]))
    return {"ms_p50": float(np.percentile(ms, 50)), "unit": "ms per query", "cores": host_threads(), "kind": "port",
            "sample": f"{n_queries} single queries x {xb.shape[0]:,}-row slice (sgemv + argpartition), time scaled x"
                      f"{args.n_items / xb.shape[0]:.1f} to the full catalog"}


def cpu_ranking(rs, B, reps=3, whole_batch=False):
    """configs[1] on the host: the reference's forward (src/models.py:37-50,79-91,125-131) as a torch-CPU fp32 F.linear chain
    with the same synthetic weights, all host cores.  Bounded: `reps` batches of B/4 samples."""
    import torch
    import torch.nn.functional as F
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, 6040, 3706, seed=10)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))     # noqa: E731
    Bs = B if whole_batch else max(1024, B // 4)
    rng = np.random.default_rng(200)
    ur = torch.from_numpy(rs.synth.make_ids(rng, 6040, Bs))
    ir = torch.from_numpy(rs.synth.make_ids(rng, 3706, Bs))
    ue, ie = t(w["user_embedding"]), t(w["item_embedding"])
    towers = {s: [(t(k).t().contiguous(), t(b)) for k, b in w[f"{s}_tower"]] for s in ("user", "item")}
    deep = [(t(k).t().contiguous(), t(b)) for k, b in w["deep"]]
    cw = [t(x).reshape(-1) for x in w["cross_w"]]
    cb = [t(x) for x in w["cross_b"]]
    heads = [(t(w[h][0]).t().contiguous(), t(w[h][1])) for h in ("rating_head", "ctr_head")]

    def tower(x, layers):
        for i, (k, b) in enumerate(layers):
            x = F.linear(x, k, b)
            if i < len(layers) - 1:
                x = torch.relu(x)
        return x

    def forward():
        u = tower(F.embedding(ur, ue), towers["user"])
        it = tower(F.embedding(ir, ie), towers["item"])
        x0 = torch.cat([u, it], dim=1)
        xl = x0
        for wv, bv in zip(cw, cb):
            xl = x0 * (xl @ wv).unsqueeze(1) + bv + xl
        d = x0
        for k, b in deep:
            d = torch.relu(F.linear(d, k, b))
        z = torch.cat([xl, d], dim=1)
        return F.linear(z, *heads[0]), torch.sigmoid(F.linear(z, *heads[1]))

    with all_host_cores(), torch.no_grad():
        forward()
        t0 = time.perf_counter()
        for _ in range(reps):
            forward()
        dt = (time.perf_counter() - t0) / reps
    return {"value": Bs / dt, "unit": "samples/s", "cores": host_threads(), "kind": "port",
            "sample": f"{reps} batches of {Bs:,} samples ({'the whole' if Bs == B else 'a quarter of the'} {B:,} batch), torch-CPU fp32 F.linear chain following "
                      f"src/models.py:37-50,79-91,125-131 (TensorFlow is not installable here), torch.set_num_threads({host_threads()})"}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    xb, q = cpu_prepare(args, args.cpu_sample_rows)
    times = []
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        dt = cpu_step(args, xb, q)
        if i >= args.warmup:
            times.append(dt)
        if time.perf_counter() - t_start > 240 and times:
            break
    scaled = float(np.mean(times))
    value = args.nq / scaled
    rows = xb.shape[0]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * scaled, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"exact top-{args.k} over {args.n_items:,} x {args.dim} fp32 catalog, "
                                   f"{args.nq}-query batches (BASELINE configs[2]); each step = one bounded slice "
                                   f"of {rows:,} rows, time scaled to the full catalog",
                       "n_items": args.n_items, "dim": args.dim, "nq": args.nq, "k": args.k},
            "cpu_baseline": cpu_describe(args, rows, scaled),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def catalog_block(torch, device, dim, block_index, block=1 << 20):
    """Synthetic catalog rows ~ N(0,1) of one 1M-row block (seed 3000 + block), not yet normalised."""
    g = torch.Generator(device=device)
    g.manual_seed(3000 + block_index)
    return torch.randn((block, dim), generator=g, device=device, dtype=torch.float32)


def build_catalog(rs, torch, index, row_begin, row_end, dim, device, block=1 << 20):
    """Unit-length catalog rows generated on the device per 1M-row block with a per-block seed, so the catalog is the same
    for every shard count."""
    b0 = (row_begin // block) * block
    while b0 < row_end:
        x = catalog_block(torch, device, dim, b0 // block, block)
        lo, hi = max(row_begin, b0), min(row_end, b0 + block)
        x = x[lo - b0:hi - b0].contiguous()
        rs.normalize_L2(x)                       # K4 on the device
        index.add(x)                             # copies into the handle's HBM buffer
        del x
        b0 += block


class _LazyItemMap:
    """item_map of a 10M-item catalog without materialising 10M dict entries: row number (as str) -> item id."""

    def __init__(self, n):
        self._n = n

    def __getitem__(self, key):
        return "item_" + key

    def __len__(self):
        return self._n


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    # stdout must carry exactly ONE JSON line: native libraries (NCCL's version banner) write to fd 1 directly, so
    # fd 1 is pointed at stderr for the duration of the run and restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import recsys_b200 as rs
    from recsys_b200.sharded import ShardedIndexFlatIP, gather_bytes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.gpus != world and rank == 0:
        print(f"note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.prestart()                   # nvidia-smi streams by the time the warm-up begins

    pk = peaks()
    N, nq, k, dim = args.n_items, args.nq, args.k, args.dim

    # ---- the catalog shard of this rank, resident in HBM behind the index handle ----
    sidx = ShardedIndexFlatIP(dim, N)
    build_catalog(rs, torch, sidx.local, sidx.row_begin, sidx.row_end, dim, device)
    assert sidx.local.ntotal == sidx.row_end - sidx.row_begin
    local_search = sidx.local.search
    sidx._local_search = lambda q_, k_: local_search(q_, k_, id_base=sidx.row_begin, mode=args.mode)
    local_deferred = sidx.local.search_deferred
    sidx._local_deferred = lambda q_, k_, out: local_deferred(q_, k_, id_base=sidx.row_begin, mode=args.mode, out=out)
    sidx.shared_threshold = args.mode != "exact" and os.environ.get("RS_SHARED_THRESHOLD", "1") != "0"
    sidx.use_graph = os.environ.get("RS_SHARDED_GRAPH", "1") != "0"   # one CUDA-graph replay per step (also on one GPU)
    sidx.reuse_outputs = True    # results are graph-owned buffers, valid until the second-next search (no copy-out per step)
    # both exchanges of a step through NVLink peer memory (csrc/peer.cu; "nccl" = two NCCL all-gathers)
    sidx.exchange = os.environ.get("RS_SHARDED_EXCHANGE", "p2p") if world > 1 else "nccl"

    gq = torch.Generator(device=device)
    gq.manual_seed(3999)
    q_dev = torch.randn((nq, dim), generator=gq, device=device, dtype=torch.float32)
    rs.normalize_L2(q_dev)
    q_pinned = torch.empty((nq, dim), dtype=torch.float32).pin_memory()
    q_pinned.copy_(q_dev.cpu())
    q_host_np = q_pinned.numpy()             # a numpy view of the pinned buffer: what a faiss caller hands to search()
    D_host = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    I_host = torch.empty((nq, k), dtype=torch.int64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return sidx.search(q_dev, k)

    # The device-resident serving loop (RS_BENCH_PIPELINE=1; off by default): submit() queues step i (one CUDA-graph replay), the step's "repeat exactly" flag
    # travels to pinned host memory on a side stream, and the host looks at it -- finish(), the repair of flagged queries --
    # after queueing step i+1, so the GPU never waits for the host between steps.  Every step's flag is read and honoured
    # inside the timed region (the last one by the drain before the closing event).
    flag_stream = torch.cuda.Stream(device=device)
    flag_host = [torch.zeros((1,), dtype=torch.bool).pin_memory() for _ in range(2)]
    f_searched = [torch.cuda.Event() for _ in range(2)]
    f_landed = [torch.cuda.Event() for _ in range(2)]
    dpipe = {"i": 0, "open": None}

    def finish_pipe(entry):
        ticket, s = entry
        f_landed[s].synchronize()
        if bool(flag_host[s][0]):
            sidx.finish(ticket)

    def step_pipe():
        s = dpipe["i"] & 1
        dpipe["i"] += 1
        ticket = sidx.submit(q_dev, k)
        if ticket[0] == "done":
            return
        f_searched[s].record()
        with torch.cuda.stream(flag_stream):
            flag_stream.wait_event(f_searched[s])
            flag_host[s].copy_(ticket[3].reshape(1), non_blocking=True)
            f_landed[s].record()
        prev, dpipe["open"] = dpipe["open"], (ticket, s)
        if prev is not None:
            finish_pipe(prev)

    def drain_pipe():
        prev, dpipe["open"] = dpipe["open"], None
        if prev is not None:
            finish_pipe(prev)

    if world == 1:
        def step_e2e():
            # the drop-in's real host path: numpy in -> numpy out, H2D / D2H staged inside the C call (rs_index_search)
            return sidx.local.search(q_host_np, k, mode=args.mode)
    else:
        def step_e2e_serial():
            qd = q_pinned.to(device, non_blocking=True)              # H2D of this step's inputs (pinned)
            D, I = sidx.search(qd, k)
            D_host.copy_(D, non_blocking=True)                       # D2H of this step's result
            I_host.copy_(I, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return D_host, I_host

        # The serving loop of the sharded index: submit() queues a step without looking at it, finish() is the step's
        # one host synchronisation.  Two batches in flight: while step i's (D, I, flag) travel to pinned host memory on a
        # copy stream, step i+1 (its own H2D included) already runs; the host completes step i after queueing step i+1.
        # Every step's H2D and D2H are inside the timed region (the last step is drained before the closing event).
        copy_stream = torch.cuda.Stream(device=device)
        outs = [(torch.empty((nq, k), dtype=torch.float32).pin_memory(), torch.empty((nq, k), dtype=torch.int64).pin_memory(),
                 torch.zeros((1,), dtype=torch.bool).pin_memory()) for _ in range(2)]
        searched = [torch.cuda.Event() for _ in range(2)]
        landed = [torch.cuda.Event() for _ in range(2)]
        pipe = {"i": 0, "open": None}

        def complete(entry):
            ticket, s = entry
            landed[s].synchronize()
            if bool(outs[s][2][0]):                                  # a flagged query (never on this data): repair, copy again
                D, I = sidx.finish(ticket)
                outs[s][0].copy_(D)
                outs[s][1].copy_(I)
            return outs[s][0], outs[s][1]

        def step_e2e():
            s = pipe["i"] & 1
            pipe["i"] += 1
            qd = q_pinned.to(device, non_blocking=True)              # H2D of this step's inputs (pinned)
            ticket = sidx.submit(qd, k)
            searched[s].record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(searched[s])
                outs[s][0].copy_(ticket[1], non_blocking=True)       # D2H of this step's result
                outs[s][1].copy_(ticket[2], non_blocking=True)
                outs[s][2].copy_(ticket[3].reshape(1), non_blocking=True)
                landed[s].record()
            prev, pipe["open"] = pipe["open"], (ticket, s)
            return complete(prev) if prev is not None else None

        def drain_e2e():
            prev, pipe["open"] = pipe["open"], None
            if prev is not None:
                complete(prev)
            torch.cuda.current_stream().wait_stream(copy_stream)

    def timed(fn, steps, drain=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        if drain is not None:
            drain()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    if rank == 0:
        sampler.start()                      # samples cover the warm-up and the timed region (both under load)
    # >= W warm-up steps, extended to ~1 s of load so clocks settle and the sampler has data.  The step count
    # must be identical on every rank (each step contains an exchange), so it is derived from a MAX-reduced time.
    n_warm = max(3, args.warmup)
    for _ in range(3):        # one-time setup (lazy shadow build, graph capture) stays out of the warm-up time estimate below
        step_device()
    barrier()
    pipelined = bool(sidx.use_graph and sidx.reuse_outputs) and os.environ.get("RS_BENCH_PIPELINE", "0") == "1"
    step_timed, drain_timed = (step_pipe, drain_pipe) if pipelined else (step_device, None)
    warm_ms = timed(step_timed, n_warm, drain=drain_timed)
    extra = int(min(2000, max(0, (1000.0 - warm_ms) / max(warm_ms / n_warm, 1e-3))))
    if extra:
        timed(step_timed, extra, drain=drain_timed)
        n_warm += extra
    barrier()
    # ---- the timed region: no per-kernel event brackets, nothing but the steps ----
    total_ms = timed(step_timed, args.steps, drain=drain_timed)
    clocks = sampler.stop() if rank == 0 else None
    # ---- second pass, same steps launched eagerly with the library's event brackets: per-kernel times + launch count
    #      (a CUDA-graph replay runs no library host code, so neither would see it) ----
    graphed = bool(getattr(sidx, "use_graph", False))
    if graphed:
        sidx.use_graph = False
        timed(step_device, 2)
    rs._lib.prof_enable(True)
    launches0 = rs._lib.launch_count()
    timed(step_device, args.steps)
    launches = rs._lib.launch_count() - launches0
    prof = {name: rs._lib.prof_read(slot) for name, slot in
            (("scan", rs._lib.PROF_SCAN), ("merge", rs._lib.PROF_MERGE), ("filter", rs._lib.PROF_FILTER),
             ("rescore", rs._lib.PROF_RESCORE), ("sample", rs._lib.PROF_SAMPLE), ("threshold", rs._lib.PROF_THRESHOLD),
             ("tops", rs._lib.PROF_TOPS), ("exchange", rs._lib.PROF_EXCHANGE))}
    rs._lib.prof_enable(False)
    if graphed:
        sidx.use_graph = True

    e2e_serial_ms = None
    if world == 1:
        for _ in range(max(3, args.warmup)):
            step_e2e()
        e2e_ms = timed(step_e2e, args.steps)
    else:
        for _ in range(max(3, args.warmup)):
            step_e2e_serial()
        e2e_serial_ms = timed(step_e2e_serial, args.steps)
        for _ in range(max(3, args.warmup)):
            step_e2e()
        drain_e2e()
        e2e_ms = timed(step_e2e, args.steps, drain=drain_e2e)

    ms_per_step = total_ms / args.steps
    value = nq / (ms_per_step / 1e3)
    e2e_value = nq / (e2e_ms / args.steps / 1e3)

    lt = torch.tensor([launches], device=device, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(lt)
    n_local = sidx.row_end - sidx.row_begin

    # ---- roofline of the dominant kernel (largest summed device time on this rank) ----
    dom = max(prof, key=lambda n_: prof[n_][0])
    dom_ms, dom_n = prof[dom]
    roofline = None
    if dom_n:
        t = dom_ms / dom_n / 1e3
        if dom == "filter":
            flops = 2.0 * nq * n_local * dim
            ach = flops / t / 1e12
            peak = pk["bf16_tflops_sustained"]
            roofline = {"kernel": "tc_filter_kernel", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": ach / peak, "traffic": None, "peak_source": pk["source"] + " (sustained bf16)",
                        "algorithmic": "2*nq*n_local*d flops per launch", "launches": dom_n,
                        "ms_per_launch": dom_ms / dom_n,
                        "hbm_frac": (n_local * dim * 2 / t / 1e9) / pk["hbm_gbs"],
                        "frac_of_burst_peak": ach / pk["bf16_tflops"],
                        "note": "frac is taken against the SUSTAINED cuBLAS bf16 rate (a kernel timed inside a step); it can exceed 1 "
                                "because the bracketed pass launches eagerly (the gaps between launches let the SM clock recover above "
                                "what 4 s of back-to-back GEMMs settle at under the 1 kW cap) and because this kernel spends fewer "
                                "joules per flop than the library GEMM (scores never leave TMEM); frac_of_burst_peak is the same "
                                "rate against the best-of-10 burst figure"}
        else:
            byts = float(n_local) * dim * 4
            ach = byts / t / 1e9
            roofline = {"kernel": "scan_topk_kernel" if dom == "scan" else dom, "bound": "hbm", "achieved": ach,
                        "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
                        "peak_source": pk["source"],
                        "algorithmic": "n_local*d*4 bytes per launch (one pass over the fp32 catalog shard; the "
                                       "exact-fp32 mode re-streams it per 8-query group and is FFMA-bound at nq=1024)",
                        "launches": dom_n, "ms_per_launch": dom_ms / dom_n,
                        "fp32_tflops": 2.0 * nq * n_local * dim / t / 1e12}
    traffic_db = {}
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        traffic_db = json.loads(tf.read_text())
    if roofline:
        ent = traffic_db.get(roofline["kernel"])
        if ent and ent.get("n_local") == n_local and ent.get("nq") == nq:
            roofline["traffic"] = ent["dram_bytes_per_launch"]
            roofline["traffic_source"] = ent["source"]
    kernel_ms = {n_: {"ms_total": round(v[0], 4), "launches": v[1]} for n_, v in prof.items() if v[1]}

    # ---- parity on this box, on the path that was just timed ----
    parity = None
    if not args.no_parity:
        try:
            parity = check_parity(rs, torch, dist, sidx, args, device, world, rank, q_dev)
        except Exception as e:  # noqa: BLE001
            parity = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and isinstance(parity, dict):     # single-GPU models and index: one line (N=1) carries it
            try:
                parity["reference_run"] = reference_run_check(rs, torch, device)
            except Exception as e:  # noqa: BLE001
                parity["reference_run"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- second headline: DCN ranking forward at batch 65,536 (configs[1]) ----
    ranking = None
    if not args.no_ranking:
        ranking = bench_ranking(rs, torch, dist, args, device, world, rank, pk, timed, traffic_db)

    cpu, xb_h = None, None
    if rank == 0 and not args.no_cpu_baseline:
        xb_h, q_h = cpu_prepare(args, args.cpu_sample_rows)
        if world == 1:
            cpu = cpu_describe(args, xb_h.shape[0], cpu_step(args, xb_h, q_h))

    batch1 = None
    if not args.no_batch1:
        batch1 = bench_batch1(rs, torch, dist, sidx, args, device, world, rank, pk, xb_h)
    del xb_h

    # configs[0], the reference's own CPU-runnable case (MovieLens-1M-shaped, batch 4096): times only, on rank 0 at N=1;
    # its outputs are compared with the oracle by tests/test_gpu_parity.py.  Never allowed to take the headline down.
    config1 = None
    if world == 1 and not args.no_ranking:
        try:
            config1 = bench_config1(rs, torch, device, timed, cpu=not args.no_cpu_baseline)
        except Exception as e:  # noqa: BLE001
            config1 = {"error": f"{type(e).__name__}: {e}"}

    # K0: StringLookup on the device, rank 0 at N=1; reported, never allowed to take the line down
    string_lookup = None
    if world == 1 and not args.no_ranking:
        try:
            string_lookup = bench_string_lookup(rs, torch, device, timed, pk, cpu=not args.no_cpu_baseline)
        except Exception as e:  # noqa: BLE001
            string_lookup = {"error": f"{type(e).__name__}: {e}"}

    # configs[3]: sharded embedding tables feeding the ranking forward (needs >= 2 GPUs).  Every rank runs it (it is
    # collective); a failure is reported, not raised -- but it is raised on EVERY rank or on none (the flag is reduced)
    config4 = None
    if world > 1 and not args.no_config4 and not args.no_ranking:
        config4 = bench_config4(rs, torch, dist, args, device, world, rank)

    if rank == 0:
        sharding = "none"
        if world > 1:
            how = ("stores into NVLink peer memory by the kernel that builds the message (sample_top_kernel / select_rescore_kernel; "
                   "its last CTA raises the flags and waits for all ranks, the threshold / merge kernels acknowledge), the whole step "
                   "seven launches sequenced by one C call (rs_index_search_sharded); no NCCL on the step"
                   if sidx.exchange == "p2p" else "NCCL all-gathers")
            sharding = (f"rows/{world} contiguous; per step: sample pass -> exchange of each shard's largest sample maxima ({how}) -> "
                        f"shared-threshold filter + exact re-score -> one fused exchange of scores | rows | flags | bounds ({how}) -> "
                        f"merge -> verify" + ("; the step is one CUDA-graph replay (per-kernel times and gpu_launches from the same "
                                              "number of eager steps)" if graphed else ""))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"exact top-{k} over {N:,} x {dim} fp32 catalog, {nq}-query batches "
                                       f"(BASELINE configs[2])", "n_items": N, "dim": dim, "nq": nq, "k": k,
                           "mode": args.mode,
                           "arithmetic": "bf16 tcgen05 pre-filter (fp32 accumulate) + exact fp32 FMA-chain re-score: results "
                                         "bit-identical to the fp32 scan",
                           "sharding": sharding,
                           "timing": "value: device-resident queries, each step one CUDA-graph replay of the search (the deferred C "
                                     "entry: no host synchronisation inside a step) " + (
                                         "whose 'repeat exactly' flag goes to pinned host memory on a side stream and is read by the host "
                                         "after it has queued the next step (RS_BENCH_PIPELINE=1: submit / finish, every flag read and "
                                         "honoured inside the timed region, the last by a drain before the closing event)" if pipelined else
                                         "followed by one read of the per-query 'repeat exactly' flags (RS_BENCH_PIPELINE=1 reads it after "
                                         "queueing the next step: measured equal on 1 and 2 GPUs)") +
                                     "; no per-kernel event brackets inside the timed region; roofline / kernels: a second "
                                     "pass of the same steps launched eagerly with the library's event brackets.  The replayed steps keep the GPU busy "
                                     "without gaps, so under the board's power cap the SM clock settles lower (see clocks) than in the "
                                     "e2e loop, whose host copies leave the GPU idle part of every step -- e2e can therefore read "
                                     "higher than value",
                           "l2": "inputs larger than L2 (catalog shard %.2f GB per GPU)" % (n_local * dim * 4 / 1e9)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nq * dim * 4,
                        "d2h_bytes_per_step": nq * k * 12, "ms_per_step": e2e_ms / args.steps,
                        "path": ("IndexFlatIP.search(numpy float32 [nq, d]) -> numpy (D, I): host buffers in and out, H2D / D2H staged "
                                 "inside rs_index_search (the faiss-shaped drop-in call)") if world == 1 else
                                ("pinned host queries -> ShardedIndexFlatIP.submit (device tensors) -> pinned host (D, I, flag) on a copy "
                                 "stream -> finish: the sharded index's serving loop with two batches in flight (step i's D2H "
                                 "overlaps step i+1; every step's H2D and D2H are inside the timed region, the last step is drained "
                                 "before the closing event); `serial` = the same without overlap (search, copy, synchronise per "
                                 "step); the numpy entry point of the single-GPU index is not sharded")},
                "gpu_launches": int(lt.item()), "clocks": clocks, "roofline": roofline, "kernels": kernel_ms,
                "parity_check": parity, "cpu_baseline": cpu, "ranking": ranking, "batch1_latency": batch1,
                "config1": config1, "config4": config4, "string_lookup": string_lookup}
        if world > 1:
            line["exchange_bytes_per_step_per_rank"] = gather_bytes(world, nq, k)
            line["e2e"]["serial"] = {"value": nq / (e2e_serial_ms / args.steps / 1e3), "ms_per_step": e2e_serial_ms / args.steps}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        sidx.close()                         # captured graphs hold peer-memory / NCCL work: release them before the group goes away
        dist.barrier()
        dist.destroy_process_group()


def check_parity(rs, torch, dist, sidx, args, device, world, rank, q_dev, n_q=16):
    """Bit-for-bit checks of the timed path, on this box.  Nothing here is timed."""
    k, dim = args.k, args.dim
    q16 = q_dev[:n_q].contiguous()
    out = {"queries": n_q}
    # (1) the timed path against the exact-scan path over the whole (sharded) catalog
    D, I = sidx.search(q16, k)
    if world > 1:
        De, Ie = sidx._gather_merge(*sidx._local_exact(q16, k))
        out["against"] = ("exact fp32 scan of every shard (mode='exact') -> NCCL all-gather -> merge: the repair path of the "
                          "sharded search, over the full 10M-row catalog")
    else:
        De, Ie = sidx.local.search(q16, k, mode="exact")
        out["against"] = "exact fp32 scan (mode='exact') over the full 10M-row catalog"
    eq = bool(torch.equal(D, De) and torch.equal(I, Ie))
    if world > 1:
        flag = torch.tensor([1 if eq else 0], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        eq = bool(flag.item())
    out["bit_equal"] = eq
    # (2) against the CPU oracle's stated fp32 chain (oracle.search_chain) on a 1M-row slice: rank 0, own small index
    if rank == 0:
        import oracle
        from oracle import flat_ip
        oracle.build_chain_lib()
        rows = min(1 << 20, args.n_items)
        x = catalog_block(torch, device, dim, 0)[:rows].contiguous()
        rs.normalize_L2(x)
        small = rs.IndexFlatIP(dim)
        small.add(x)
        Ds, Is = small.search(q16, k, mode=args.mode, id_base=12345)
        xh, qh = x.cpu().numpy(), q16.cpu().numpy()
        del x
        with all_host_cores():
            Dw, Iw = flat_ip.search_chain(xh, qh, k, id_base=12345)
        out["oracle_slice"] = {"rows": rows, "id_base": 12345,
                               "bit_equal": bool(np.array_equal(Ds.cpu().numpy(), Dw) and np.array_equal(Is.cpu().numpy(), Iw)),
                               "against": "oracle.flat_ip.search_chain (single fp32 FMA chain per score, tie-break score desc / row asc) on the host"}
    return out


def reference_run_check(rs, torch, device):
    """The CUDA path against tests/golden/reference_run.*: outputs of the reference's OWN files (src/models.py,
    src/trainer.py's catalog build, app/recommendation_service.py) run over numpy stand-ins of TF / faiss by
    tests/golden/make_reference_golden.py (the same checks as tests/test_reference_run.py, here as a flag in the line)."""
    golden = ROOT / "tests" / "golden"
    if str(golden) not in sys.path:
        sys.path.insert(0, str(golden))
    import make_reference_golden as mrg
    arrays, doc = mrg.load_committed()
    cfg, w, uv, iv, users, items = mrg.synth_inputs()
    model = rs.MultiTaskModel(cfg, uv, iv, {}, device=device)
    rs.load_weights(model, w)
    out = model({"user_id": users, "movie_id": items}, training=False)
    worst = 0.0
    for key in ("user_embedding", "item_embedding", "rating_prediction", "ctr_prediction"):
        want = arrays[f"f64_{key}"]
        worst = max(worst, float(np.abs(out[key].cpu().numpy().astype(np.float64) - want).max() / np.abs(want).max()))
    index, item_map = rs.build_item_index(model.encoder, iv)
    svc = rs.RecommendationService.from_parts(model.encoder, index, uv, item_map)
    same_items, score_err = True, 0.0
    for u in mrg.SPEC["users"]:
        got, want = svc.recommend(u, mrg.SPEC["k"]), doc["recommend"][u]
        same_items = same_items and [x["item_id"] for x in got] == [x["item_id"] for x in want]
        score_err = max(score_err, max(abs(a["score"] - b["score"]) for a, b in zip(got, want)))
    return {"against": "tests/golden/reference_run.*: the reference's own src/models.py / trainer catalog build / "
                       "recommendation_service.py run over numpy stand-ins of TF and faiss (float64 primitives)",
            "ranking_max_err_over_scale": worst, "ranking_within_1e-5": bool(worst <= 1e-5),
            "recommend_users": len(mrg.SPEC["users"]), "recommend_items_equal": bool(same_items),
            "recommend_max_score_err": score_err, "cold_start_equal": svc.recommend("nobody", 7) == doc["cold_start"]}


def bench_ranking(rs, torch, dist, args, device, world, rank, pk, timed, traffic_db):
    """MultiTaskModel forward (ids -> towers -> DCN -> heads) at batch 65,536 per GPU (data-parallel replicas; weak
    scaling, no collective).  Headline = bf16 tensor mode (tcgen05 fused chains, tolerance 5e-3 of scale vs the fp64
    oracle, tests/test_gpu_ranking_bf16.py); the fp32 FFMA mode (1e-5) is reported beside it."""
    cfg = rs.ModelConfig()
    n_users, n_items, B = 6040, 3706, args.rank_batch
    w = rs.synth.make_weights(cfg, n_users, n_items, seed=10)
    rng = np.random.default_rng(200 + rank)
    # row numbers as int32 (what a StringLookup emits fits easily; halves the upload), outputs as two contiguous rows
    ur = torch.from_numpy(rs.synth.make_ids(rng, n_users, B).astype(np.int32)).pin_memory()
    ir = torch.from_numpy(rs.synth.make_ids(rng, n_items, B).astype(np.int32)).pin_memory()
    ur_d, ir_d = ur.to(device), ir.to(device)
    out_host = torch.empty((2, B), dtype=torch.float32).pin_memory()
    emb_host = torch.empty((2, B, 128), dtype=torch.float32).pin_memory()   # one contiguous pinned block per embedding
    flops = 524288.0 * B
    res = {}
    for precision in ("bf16", "fp32"):
        model = rs.MultiTaskModel(cfg, rs.synth.make_vocab(n_users), rs.synth.make_vocab(n_items), {}, device=device,
                                  precision=precision)
        rs.load_weights(model, w)

        def step_eager():
            return model({"user_id": ur_d, "movie_id": ir_d})

        # bf16 mode: the forward is replayed as one CUDA graph over static buffers (RankingSession) -- the eager call's
        # host work (~50 us: allocation, DLPack export, two launches) is as long as the kernels at this batch size.
        # The eager figure is reported beside it.  fp32 mode (7 launches of the fp32 tensor kernel) is replayed the same way.
        sess = model.ranking_session(B)

        def run_once(full):
            if sess is not None:
                sess.user_rows.copy_(ur, non_blocking=True)
                sess.item_rows.copy_(ir, non_blocking=True)
                o = sess.run()
            else:
                o = model({"user_id": ur.to(device, non_blocking=True), "movie_id": ir.to(device, non_blocking=True)})
            out_host[0].copy_(o["rating_prediction"].view(-1), non_blocking=True)
            out_host[1].copy_(o["ctr_prediction"].view(-1), non_blocking=True)
            if full:     # what MultiTaskModel.call returns in full: the two [B,128] embeddings as well
                emb_host[0].copy_(o["user_embedding"], non_blocking=True)
                emb_host[1].copy_(o["item_embedding"], non_blocking=True)
            torch.cuda.current_stream().synchronize()

        if sess is not None:
            sess.user_rows.copy_(ur_d)
            sess.item_rows.copy_(ir_d)
            step = sess.run
        else:
            step = step_eager

        steps = max(args.steps, 20)
        for _ in range(5):
            step()
            step_eager()
        ms = timed(step, steps) / steps
        # per-kernel device time and the launch count come from eager steps (a graph replay runs no library host code)
        rs._lib.prof_enable(True)
        l0 = rs._lib.launch_count()
        eager_ms = timed(step_eager, steps) / steps
        launches = rs._lib.launch_count() - l0
        tw_ms, tw_n = rs._lib.prof_read(rs._lib.PROF_RANK)
        dp_ms, dp_n = rs._lib.prof_read(rs._lib.PROF_DEEP)
        rs._lib.prof_enable(False)
        e2e = {}
        for label, full in (("logits", False), ("all_outputs", True)):
            for _ in range(3):
                run_once(full)
            e2e[label] = timed(lambda: run_once(full), steps) / steps
        res[precision] = {"value": world * B / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms,
                          "eager_ms_per_step": eager_ms, "cuda_graph": sess is not None,
                          "e2e": {"value": world * B / (e2e["logits"] / 1e3), "unit": "samples/s", "h2d_bytes_per_step": 2 * B * 4,
                                  "d2h_bytes_per_step": B * 8,
                                  "note": "int32 row ids up, the two logit vectors down; MultiTaskModel.call also returns the two [B,128] "
                                          "fp32 embeddings, which stay on the device here -- `e2e_all_outputs` brings them down too"},
                          "e2e_all_outputs": {"value": world * B / (e2e["all_outputs"] / 1e3), "unit": "samples/s",
                                              "h2d_bytes_per_step": 2 * B * 4, "d2h_bytes_per_step": B * 8 + B * 1024},
                          "gemm_tflops_per_gpu": flops / (ms / 1e3) / 1e12,
                          "gpu_launches_per_step": launches // max(1, steps)}
        if precision == "bf16" and (tw_n or dp_n):
            # the graph replay IS the two chain kernels back to back (nothing else is in the graph), so the step time is the
            # kernels' device time including the gap between them; the event-bracketed eager launches give the split
            t = ms / 1e3
            ent = traffic_db.get("ranking_bf16") or {}
            res[precision]["roofline"] = {"kernel": "tower_sp_kernel + deep_heads2_kernel (chain_tc.cu)", "bound": "tensor",
                                          "achieved": flops / t / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                          "frac": flops / t / 1e12 / pk["bf16_tflops_sustained"],
                                          "traffic": ent.get("dram_bytes_per_launch") if ent.get("batch") == B else None,
                                          "traffic_source": ent.get("source") if ent.get("batch") == B else None,
                                          "algorithmic": "524,288 GEMM flop per sample; duration = one CUDA-graph replay of the two kernels",
                                          "ms_per_step": ms, "tower_ms_eager_bracketed": tw_ms / steps, "deep_ms_eager_bracketed": dp_ms / steps,
                                          "launches_per_step": (tw_n + dp_n) / steps, "peak_source": pk["source"] + " (sustained bf16)"}
    # the FFMA kernels the fp32 tensor kernel replaced at this batch (RS_FP32_TENSOR=0), same model, same tolerance
    os.environ["RS_FP32_TENSOR"] = "0"
    try:
        for _ in range(3):
            step_eager()
        ffma_ms = timed(step_eager, max(5, args.steps // 2)) / max(5, args.steps // 2)
        res["fp32_ffma"] = {"value": world * B / (ffma_ms / 1e3), "ms_per_step": ffma_ms, "gemm_tflops_per_gpu": flops / (ffma_ms / 1e3) / 1e12}
    finally:
        os.environ.pop("RS_FP32_TENSOR", None)
    head = res["bf16"]
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_ranking(rs, B)
        except Exception as e:  # noqa: BLE001
            cpu = {"error": f"{type(e).__name__}: {e}"}
    return {"metric": "DCN ranking forward samples/sec", "value": head["value"], "unit": "samples/s",
            "ms_per_step": head["ms_per_step"], "batch_per_gpu": B, "n_gpus": world, "scaling": "weak",
            "mode": "bf16 tensor mode (tcgen05, fp32 accumulate; tolerance 5e-3 of tensor scale)", "dtype": "bf16",
            "config": {"workload": f"MultiTaskModel forward, 3 cross layers + MLP, emb 128, batch {B} per GPU (BASELINE configs[1])"},
            "e2e": head["e2e"], "e2e_all_outputs": head["e2e_all_outputs"], "gemm_tflops_per_gpu": head["gemm_tflops_per_gpu"],
            "tensor_frac_of_sustained_bf16": head["gemm_tflops_per_gpu"] / pk["bf16_tflops_sustained"],
            "roofline": head.get("roofline"), "gpu_launches_per_step": head["gpu_launches_per_step"],
            "cuda_graph": head["cuda_graph"], "eager_ms_per_step": head["eager_ms_per_step"], "cpu_baseline": cpu,
            "fp32_mode": {"value": res["fp32"]["value"], "ms_per_step": res["fp32"]["ms_per_step"], "dtype": "f32",
                          "arithmetic": "fp32 operands split into two fp16 numbers, three tcgen05 products per K step, fp32 "
                                        "accumulation (csrc/dense_x3.cu); cross stack and heads in fp32 FMA",
                          "tolerance": "1e-5 of tensor scale (measured <= 1.1e-6 per layer, 6.5e-7 whole model: tests/test_gpu_dense_x3.py)",
                          "cuda_graph": res["fp32"]["cuda_graph"], "eager_ms_per_step": res["fp32"]["eager_ms_per_step"],
                          "gpu_launches_per_step": res["fp32"]["gpu_launches_per_step"], "e2e": res["fp32"]["e2e"],
                          "gemm_tflops_per_gpu": res["fp32"]["gemm_tflops_per_gpu"],
                          "tensor_tflops_issued_per_gpu": 3 * res["fp32"]["gemm_tflops_per_gpu"],
                          "ffma_kernel": res.get("fp32_ffma")}}


def cpu_config1(rs, n_items=3706, B=4096, k=10, reps=3):
    """configs[0] on the host -- the one case the reference itself runs on a CPU: the forward as cpu_ranking's torch chain at
    the whole 4,096 batch, and exact top-10 of 4,096 unit queries over a 3,706-row catalog through the sgemm + heap
    restatement of faiss-cpu's batched path."""
    import oracle
    fwd = cpu_ranking(rs, B, reps=reps, whole_batch=True)
    xb = rs.synth.make_unit_rows(n_items, 128, 7)
    q = rs.synth.make_unit_rows(B, 128, 8)
    with all_host_cores():
        oracle.search_blas(xb, q[:64], k)
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle.search_blas(xb, q, k)
        dt = (time.perf_counter() - t0) / reps
    return {"forward_samples_per_s": fwd["value"], "forward_ms": B / fwd["value"] * 1e3, "top10_queries_per_s": B / dt,
            "top10_ms": dt * 1e3, "cores": host_threads(), "kind": "port",
            "sample": f"{reps} whole batches each: {fwd['sample']}; top-10: OpenBLAS sgemm + per-query heap over all {n_items:,} rows"}


def bench_config1(rs, torch, device, timed, steps=20, cpu=True):
    """BASELINE configs[0]: 6,040 users / 3,706 items, emb 128, batch 4,096 -- MultiTaskModel forward (fp32 exact mode, the
    reference's arithmetic) and exact top-10 over the item-tower catalog for the batch's user embeddings.  The whole
    problem sits in L2, so this is a launch/latency figure, not a roofline one."""
    cfg = rs.ModelConfig()
    n_users, n_items, B = 6040, 3706, 4096
    uv, iv = rs.synth.make_vocab(n_users), rs.synth.make_vocab(n_items)
    model = rs.MultiTaskModel(cfg, uv, iv, {}, device=device)
    rs.load_weights(model, rs.synth.make_weights(cfg, n_users, n_items, seed=10))
    rng = np.random.default_rng(100)
    ur = torch.from_numpy(rs.synth.make_ids(rng, n_users, B)).to(device)
    ir = torch.from_numpy(rs.synth.make_ids(rng, n_items, B)).to(device)
    items = model.encoder({"movie_id": torch.arange(1, n_items + 1, device=device)})["item_embedding"].contiguous()
    rs.normalize_L2(items)
    index = rs.IndexFlatIP(cfg.embedding_dim)
    index.add(items)

    def forward():
        return model({"user_id": ur, "movie_id": ir})

    q = forward()["user_embedding"].contiguous()
    rs.normalize_L2(q)

    def search():
        return index.search(q, 10)

    for _ in range(3):
        forward()
        search()
    f_ms = timed(forward, steps) / steps
    s_ms = timed(search, steps) / steps
    fb = index.fallbacks()[0]
    # the same forward as one CUDA-graph replay (7 launches of the fp32 tensor kernel: the eager call is host-bound here)
    sess = model.ranking_session(B, ids_dtype=torch.int64)
    sess.user_rows.copy_(ur)
    sess.item_rows.copy_(ir)
    for _ in range(3):
        sess.run()
    g_ms = timed(sess.run, steps) / steps
    # retrieval over a catalog with spread-out rows of the same shape (unit Gaussian directions): the synthetic model's item
    # embeddings all point almost the same way, which packs every score inside the bf16 filter's margin -- the small-catalog
    # tensor path then declines (AUTO stays on the exact scan after the first attempt); a trained tower does not do that
    spread = rs.IndexFlatIP(cfg.embedding_dim)
    spread.add(torch.from_numpy(rs.synth.make_unit_rows(n_items, cfg.embedding_dim, 7)).to(device))
    qs = torch.from_numpy(rs.synth.make_unit_rows(B, cfg.embedding_dim, 8)).to(device)
    for _ in range(3):
        spread.search(qs, 10)
    t_ms = timed(lambda: spread.search(qs, 10), steps) / steps
    e_ms = timed(lambda: spread.search(qs, 10, mode="exact"), steps) / steps
    return {"workload": "MovieLens-1M-shaped synthetic (6,040 users / 3,706 items, emb 128, batch 4096): MultiTaskModel forward "
                        "in fp32 mode + exact top-10 over the 3,706-item catalog (BASELINE configs[0])",
            "forward_ms": g_ms, "forward_samples_per_s": B / (g_ms / 1e3), "forward_eager_ms": f_ms, "forward_cuda_graph": True,
            "top10_ms": s_ms, "top10_queries_per_s": B / (s_ms / 1e3),
            "top10_path": "exact fp32 scan (the synthetic towers' item embeddings are nearly collinear: the tensor path declined)"
                          if fb or index.fallbacks()[1] else "small-catalog tensor path (tcgen05 pre-filter + exact re-score)",
            "top10_spread_catalog": {"tensor_path_ms": t_ms, "exact_scan_ms": e_ms, "queries_per_s": B / (t_ms / 1e3),
                                     "fallbacks": spread.fallbacks()[1],
                                     "note": "same shape, unit Gaussian rows and queries: tcgen05 pre-filter + exact re-score, results "
                                             "bit-identical to the scan (tests/test_gpu_tensor_path.py)"},
            "cpu_baseline": cpu_config1(rs, n_items, B) if cpu else None,
            "note": "problem is L2-resident: launch/latency-bound, no roofline fraction"}


def bench_string_lookup(rs, torch, device, timed, pk, n_words=1 << 20, B=1 << 22, steps=20, cpu=True):
    """K0: keras.layers.StringLookup (src/models.py:70,73,85,89) on the device -- 4M string ids (16-byte rows, 1 % out of
    vocabulary) against a 1M-word vocabulary.  Byte work: per id 16 B read + 8 B written (the table -- 32 MB of
    tagged slots + 16 MB of words -- is probed at random, mostly from L2); checked id for id against the host dict the models use otherwise."""
    vocab = rs.synth.make_vocab(n_words)
    words = rs.ops.pack_strings(vocab)
    rng = np.random.default_rng(500)
    pick = rng.integers(0, n_words, B)
    ids_h = words[pick]                                            # uint8 [B, 16]
    oov = rng.random(B) < 0.01
    ids_h[oov, 0] = ord("x")                                       # no word starts with a letter
    want = np.where(oov, 0, pick + 1).astype(np.int64)
    dv = rs.ops.DeviceVocabulary(vocab, device=device)
    ids_pin = torch.from_numpy(ids_h).pin_memory()
    ids_d = ids_pin.to(device)
    rows_d = torch.empty((B,), dtype=torch.int64, device=device)
    rows_pin = torch.empty((B,), dtype=torch.int64).pin_memory()

    def step():
        dv.lookup(ids_d, out=rows_d)

    def step_e2e():
        ids_d.copy_(ids_pin, non_blocking=True)
        dv.lookup(ids_d, out=rows_d)
        rows_pin.copy_(rows_d, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        step_e2e()
    equal = bool(np.array_equal(rows_pin.numpy(), want))
    d_ms = timed(step, steps) / steps
    e_ms = timed(step_e2e, steps) / steps
    alg = B * (ids_h.shape[1] + 8)
    ent = {}
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        ent = json.loads(tf.read_text()).get("string_lookup_kernel") or {}
    same_shape = ent.get("ids") == B and ent.get("n_words") == n_words
    out = {"workload": f"{B:,} string ids (uint8 [B, {ids_h.shape[1]}], 1 % out of vocabulary) -> int64 rows, vocabulary of {n_words:,} words",
           "ids_per_s": B / (d_ms / 1e3), "ms_per_step": d_ms,
           "e2e": {"ids_per_s": B / (e_ms / 1e3), "ms_per_step": e_ms, "h2d_bytes_per_step": int(ids_h.nbytes),
                   "d2h_bytes_per_step": B * 8, "path": "pinned byte matrix -> device -> DeviceVocabulary.lookup -> pinned int64 rows"},
           "roofline": {"bound": "hbm", "achieved": alg / (d_ms / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": alg / (d_ms / 1e3) / 1e9 / pk["hbm_gbs"],
                        "traffic": ent.get("dram_bytes_per_launch") if same_shape else None,
                        "traffic_source": ent.get("source") if same_shape else None,
                        "note": "algorithmic bytes = ids read + rows written; the table (32 MB of tagged slots + 16 MB of words) is "
                                "read at random and mostly L2-resident, it is not counted.  The kernel is bound by the latency of "
                                "three dependent loads per id (row -> bucket -> word) and by L1 wavefronts of the divergent "
                                "probes, not by bytes"},
           "parity": {"ids": B, "equal_to_host_dict": equal}, "gpu_launches_per_step": 1}
    if cpu:
        look = rs.StringLookup(vocab, device=device)
        n = 1 << 18
        sample = [vocab[i] if not o else "x" + vocab[i] for i, o in zip(pick[:n], oov[:n])]
        t0 = time.perf_counter()
        got = look.lookup_host(sample)
        dt = time.perf_counter() - t0
        out["parity"]["sample_equal_to_lookup_host"] = bool(np.array_equal(got, want[:n]))
        # the model-facing call on a LIST of Python strings (65,536 ids): Arrow packing + device table (the automatic path
        # for large batches against large vocabularies) against the host dict, both ending with the rows on the device
        m = 1 << 16
        times = {}
        for label, env in (("device_table_ms", "1"), ("host_dict_ms", "0")):
            os.environ["RS_DEVICE_LOOKUP"] = env
            best = None
            for _ in range(4):
                t0 = time.perf_counter()
                rows = look(sample[:m])
                torch.cuda.synchronize()
                dt_ = time.perf_counter() - t0
                best = dt_ if best is None else min(best, dt_)
            times[label] = best * 1e3
            times[label.replace("_ms", "_equal")] = bool(np.array_equal(rows.cpu().numpy(), want[:m]))
        os.environ.pop("RS_DEVICE_LOOKUP", None)
        out["str_list_call"] = dict(times, ids=m, note="StringLookup.__call__(list of str): best of 4, host wall time incl. the "
                                                        "upload; the first device call also builds the table (excluded by best-of)")
        out["cpu_baseline"] = {"value": n / dt, "unit": "ids/s", "cores": 1, "kind": "port",
                               "sample": f"{n:,} of the same ids as Python strings through the host dict (the lookup the models use "
                                         f"for string lists; stands for the TF hash-table op)"}
    return out


def bench_batch1(rs, torch, dist, sidx, args, device, world, rank, pk, xb_h, n_queries=200, k=10):
    """BASELINE config 5: single-query top-10 over the (sharded) catalog, one query at a time.  Device time per query
    (CUDA events around each call), host round trip (host query in, host (D, I) out), and the whole service call
    `RecommendationService.recommend(user_id, 10)` (app/recommendation_service.py:64-81: membership test, user tower,
    normalise, search, the D2H of k results, the Python result list), p50 / p99."""
    g = torch.Generator(device=device)
    g.manual_seed(5555)
    qs = torch.randn((n_queries, args.dim), generator=g, device=device, dtype=torch.float32)
    rs.normalize_L2(qs)
    qs_host = qs.cpu().pin_memory()
    for i in range(20):
        sidx.search(qs[i:i + 1].contiguous(), k)
    torch.cuda.synchronize()
    dev_ms, wall_ms, svc_ms = [], [], []
    for i in range(n_queries):
        q1 = qs[i:i + 1].contiguous()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sidx.search(q1, k)
        e1.record()
        e1.synchronize()
        dev_ms.append(e0.elapsed_time(e1))
    for i in range(n_queries):
        t0 = time.perf_counter()
        D, I = sidx.search(qs_host[i:i + 1].to(device, non_blocking=True), k)
        D.cpu(), I.cpu()
        wall_ms.append((time.perf_counter() - t0) * 1e3)
    # the service call: a MovieLens-sized user tower in front of the same catalog (every rank calls it in lockstep at N>1)
    svc_err = None
    try:
        cfg = rs.ModelConfig()
        uv = rs.synth.make_vocab(6040)
        enc = rs.MultiTowerModel(cfg, uv, rs.synth.make_vocab(8), {}, device=device)
        rs.load_weights(enc, rs.synth.make_weights(cfg, 6040, 8, seed=10))
        svc = rs.RecommendationService.from_parts(enc, sidx if world > 1 else sidx.local, uv, _LazyItemMap(args.n_items))
        users = [uv[j] for j in np.random.default_rng(77).integers(0, len(uv), size=n_queries)]
        for u in users[:20]:
            svc.recommend(u, k)
        for u in users:
            t0 = time.perf_counter()
            recs = svc.recommend(u, k)
            svc_ms.append((time.perf_counter() - t0) * 1e3)
        assert len(recs) == k and recs[0]["rank"] == 1
    except Exception as e:  # noqa: BLE001
        svc_err = f"{type(e).__name__}: {e}"
        svc_ms = [float("nan")]
    dm = torch.tensor([float(np.percentile(dev_ms, 50)), float(np.percentile(dev_ms, 99)),
                       float(np.percentile(wall_ms, 50)), float(np.percentile(wall_ms, 99)),
                       float(np.percentile(svc_ms, 50)), float(np.percentile(svc_ms, 99))], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dm, op=dist.ReduceOp.MAX)
    dm = dm.tolist()
    n_local = sidx.row_end - sidx.row_begin
    # the batch-1 search streams the bf16 shadow of the shard once (the fp32 rows are touched only for the ~k re-scored
    # candidates): those are the bytes the kernel really reads, so the roofline is stated on them
    shadow_bytes = float(n_local) * args.dim * 2
    ach = shadow_bytes / (dm[0] / 1e3) / 1e9
    out = {"workload": f"batch-1 top-{k} over {args.n_items:,} x {args.dim} (BASELINE configs[4]), {n_queries} queries one at a time",
           "device_ms_p50": dm[0], "device_ms_p99": dm[1], "e2e_ms_p50": dm[2], "e2e_ms_p99": dm[3],
           "recommend_e2e_ms_p50": dm[4], "recommend_e2e_ms_p99": dm[5],
           "recommend_path": "RecommendationService.recommend(user_id, 10): dict membership -> StringLookup (host) -> row number up -> "
                             "user tower (fp32 kernels) -> normalize_L2 -> IndexFlatIP search -> D2H of k (score, row) pairs -> "
                             "result list" + ("; on one GPU the device part is one CUDA-graph replay per request "
                                              "(service._single_user: same kernels and results as the eager calls)" if world == 1 else
                                              "; sharded index: eager calls"),
           "queries_per_s_at_p50": 1e3 / dm[2],
           "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                        "algorithmic": "n_local*d*2 bytes per query: one pass over the bf16 shadow of the catalog shard (the fp32 "
                                       "catalog, n_local*d*4 = 2x, is read only for the re-scored candidates); whole-search device "
                                       "time p50 (all launches of a search, not the filter kernel alone)",
                        "traffic": None, "peak_source": pk["source"]}}
    if svc_err:
        out["recommend_error"] = svc_err
    if rank == 0 and world == 1 and xb_h is not None:
        try:
            out["cpu_baseline"] = cpu_batch1(args, xb_h, k=k)
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def bench_config4(rs, torch, dist, args, device, world, rank, steps=20):
    """BASELINE configs[3]: Zipf-skewed lookups into a 50M x 128 and a 10M x 128 table sharded `row mod N` over the GPUs,
    feeding the bf16 ranking forward; local batch 65,536 per GPU.  Lookups go through NVLink peer memory (csrc/peer.cu:
    de-duplicated ids to the owners, the owner-side gather written straight into the requester's region), both tables
    concurrently, and the whole lookups + ranking step is replayed as one CUDA graph.  Parity: every looked-up row of
    every rank is compared bit for bit with the host replay of the hash-generated table."""
    from recsys_b200.sharded import ShardedEmbedding, lookup_many
    E, B = 128, args.rank_batch
    stage, err = "setup", None
    res = {}
    tables = {}
    try:
        for name, V, seed in (("user", args.c4_users, 41), ("item", args.c4_items, 42)):
            shard = ShardedEmbedding.fill_synthetic(V, E, rank, world, seed, device)
            tables[name] = ShardedEmbedding(V, E, local_table=shard)
        rng = np.random.default_rng(4100 + rank)
        ids = {"user": rs.synth.zipf_ids(rng, min(args.c4_users, 5_000_000), B), "item": rs.synth.zipf_ids(rng, min(args.c4_items, 5_000_000), B)}
        # spread the Zipf head over the whole id range (the generator's permutation covers the first 5M ids)
        ids["user"] = (ids["user"].astype(np.int64) * 9) % args.c4_users
        ids["item"] = (ids["item"].astype(np.int64) * 3) % args.c4_items
        ids_d = {k_: torch.from_numpy(v).to(device) for k_, v in ids.items()}
        cfg = rs.ModelConfig()
        w = rs.synth.make_weights(cfg, 8, 8, seed=10)        # tower / DCN weights; the 9-row tables are placeholders
        model = rs.MultiTaskModel(cfg, rs.synth.make_vocab(8), rs.synth.make_vocab(8), {}, device=device, precision="bf16")
        rs.load_weights(model, w)
        stage = "peer regions"
        for name in ("user", "item"):
            tables[name].enable_peer_exchange(B, dedup=True)
    except Exception as e:  # noqa: BLE001
        err = f"{stage}: {type(e).__name__}: {e}"
    ok = torch.tensor([0 if err else 1], device=device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not int(ok.item()):
        for t in tables.values():
            try:
                t.close()
            except Exception:  # noqa: BLE001
                pass
        return {"error": err or "setup failed on another rank"}

    def step_lookup():
        return lookup_many([(tables["user"], ids_d["user"]), (tables["item"], ids_d["item"])])

    def step_full():
        u, it = step_lookup()
        return model.rank_from_embedding_rows(u, it)

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=device, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- parity: the exchanged rows == the host replay, every row of every rank, twice (the second pass re-uses the regions)
    rows_equal = True
    for _ in range(2):
        u, it = step_lookup()
        rows_equal &= bool(np.array_equal(u.cpu().numpy(), rs.synth.hash_table_rows(ids["user"], E, 41)))
        rows_equal &= bool(np.array_equal(it.cpu().numpy(), rs.synth.hash_table_rows(ids["item"], E, 42)))
    out = step_full()
    flag = torch.tensor([1 if rows_equal else 0], device=device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    rows_equal = bool(flag.item())
    # logits of the first 512 samples of rank 0 against the fp64 oracle forward over the replayed rows
    logits_rel = None
    if rank == 0:
        from oracle.models import MultiTaskModel as OMT   # checker only
        w2 = dict(w)
        w2["user_embedding"] = rs.synth.hash_table_rows(ids["user"][:512], E, 41)
        w2["item_embedding"] = rs.synth.hash_table_rows(ids["item"][:512], E, 42)
        ref = OMT(w2, [], [], np.float64).rank_rows(np.arange(512), np.arange(512))
        logits_rel = max(float(np.abs(out[key][:512].cpu().numpy().astype(np.float64) - ref[key]).max() / np.abs(ref[key]).max())
                         for key in ("rating_prediction", "ctr_prediction", "user_embedding"))
    res["eager"] = (timed(step_lookup), timed(step_full))
    # the step has no host synchronisation: capture lookups + ranking once, replay as one CUDA graph
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        step_full()
    torch.cuda.current_stream(device).wait_stream(side)
    dist.barrier()
    torch.cuda.synchronize()
    g_lookup, g_full = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_lookup):
        rows_static = step_lookup()
    with torch.cuda.graph(g_full):
        out_static = step_full()
    g_full.replay()
    torch.cuda.synchronize()
    graph_equal = all(bool(torch.equal(out_static[key], out[key])) for key in ("rating_prediction", "ctr_prediction"))
    res["graph"] = (timed(g_lookup.replay), timed(g_full.replay))
    del g_lookup, g_full, rows_static, out_static
    uniq = torch.tensor([float(len(np.unique(ids["user"])) + len(np.unique(ids["item"]))) / (2 * B)], device=device, dtype=torch.float64)
    dist.all_reduce(uniq, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize()
    dist.barrier()
    for t in tables.values():
        t.close()
    if rank != 0:
        return None
    remote = (world - 1) / world
    lookup_ms, full_ms = res["graph"]
    # bytes that cross NVLink per GPU and direction for one step, WITHOUT counting the de-duplication saving (the
    # algorithmic figure of SURVEY.md section 8d): rows back 2 tables x 512 B, ids out 2 x 8 B, for the remote share
    payload = 2 * (4 * E + 8) * remote * B
    ach = payload / (lookup_ms / 1e3) / 1e9
    return {"workload": f"tables {args.c4_users:,} x 128 + {args.c4_items:,} x 128 (fp32, {4 * E * (args.c4_users + args.c4_items) / world / 1e9:.2f} GB per GPU) "
                        f"sharded row mod {world}, Zipf(1.05) ids, local batch {B} per GPU, peer-memory all-to-all lookups feeding the bf16 "
                        f"ranking forward (BASELINE configs[3])",
            "n_gpus": world, "scaling": "weak", "value": world * B / (full_ms / 1e3), "unit": "samples/s",
            "lookup_ms": lookup_ms, "lookup_plus_ranking_ms": full_ms, "lookup_samples_per_s": world * B / (lookup_ms / 1e3),
            "eager_lookup_ms": res["eager"][0], "eager_lookup_plus_ranking_ms": res["eager"][1],
            "cuda_graph": True, "dedup": True, "distinct_id_fraction": float(uniq.item()) / world,
            "roofline": {"bound": "nvlink", "achieved": ach, "peak": NVLINK_PEER_GBS, "unit": "GB/s per GPU per direction", "frac": ach / NVLINK_PEER_GBS,
                         "algorithmic": "(2 x 512 B rows back + 2 x 8 B ids out) x (N-1)/N remote share x local batch, per GPU and direction, "
                                        "without the de-duplication saving; divided by the lookup time (graph replay)",
                         "peak_source": "measured peer copy on this pool (B200_PROFILING.md); nominal 900", "traffic": None},
            "parity": {"rows_bit_identical_to_host_replay": rows_equal, "rows_checked_per_rank": 4 * B,
                       "graph_replay_equals_eager": graph_equal, "logits_max_rel_err_vs_fp64_oracle": logits_rel,
                       "logits_tolerance": 1e-2}}


if __name__ == "__main__":
    main()
