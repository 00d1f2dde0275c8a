#!/usr/bin/env python
"""bench.py -- headline benchmark of the serving/scoring hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU path, timed on the host cores

Metric (BASELINE.json): queries/sec of exact top-100 retrieval over a 10M-item, d=128 catalog in
1,024-query batches (configs[2]; the catalog is 5.12 GB and fits one GPU, at N>1 it is row-sharded and the
per-shard lists are all-gathered over NCCL and merged).  A "step" is one 1,024-query batch over the whole
catalog.  The same JSON line also carries, under "ranking", the second headline: DCN ranking forward
samples/sec at batch 65,536 (configs[1]).

`value`  : queries/s with the query batch already in HBM (catalog resident, as in faiss after add()).
`e2e`    : the same through the faiss-shaped public call with HOST queries in and HOST (D, I) out.
`roofline`: the dominant kernel, timed with CUDA events on its own stream inside the timed region.
`cpu_baseline`: the CPU restatement (oracle.search_blas: blocked sgemm + heap, what faiss-cpu does for
            > 20 queries) on a bounded slice of the same workload, on this box's host cores.
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
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


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


# ----------------------------------------------------------------------------- CPU arm / baseline
def cpu_prepare(args, sample_rows: int):
    import oracle
    import recsys_b200 as rs
    oracle.build_chain_lib()
    rows = min(sample_rows, args.n_items)
    xb = rs.synth.make_unit_rows(rows, args.dim, 3000)
    q = rs.synth.make_unit_rows(args.nq, args.dim, 3999)
    oracle.search_blas(xb[:4096], q[:64], args.k)           # warm the BLAS threads
    return xb, q


def cpu_step(args, xb, q) -> float:
    """One bounded CPU step: nq queries x the catalog slice; returns seconds scaled to the full catalog."""
    import oracle
    t0 = time.perf_counter()
    oracle.search_blas(xb, q, args.k)
    return (time.perf_counter() - t0) * (args.n_items / xb.shape[0])


def cpu_describe(args, rows, scaled_s):
    return {"value": args.nq / scaled_s, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"{args.nq} queries x {rows:,}-row slice ({rows / args.n_items:.0%} of the catalog), blocked "
                      f"OpenBLAS sgemm + per-query heap (faiss-cpu's >20-query path restated; TF/faiss are not "
                      f"installable here), time scaled x{args.n_items / rows:.1f} to the full catalog"}


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
def build_catalog(rs, torch, index, row_begin, row_end, dim, device, block=1 << 20):
    """Synthetic catalog rows ~ N(0,1) -> unit length, generated on the device per 1M-row block with a
    per-block seed (3000 + block), so the catalog is the same for every shard count."""
    b0 = (row_begin // block) * block
    while b0 < row_end:
        g = torch.Generator(device=device)
        g.manual_seed(3000 + b0 // block)
        x = torch.randn((block, dim), generator=g, device=device, dtype=torch.float32)
        lo, hi = max(row_begin, b0), min(row_end, b0 + block)
        x = x[lo - b0:hi - b0].contiguous()
        rs.normalize_L2(x)                       # K4 on the device
        index.add(x)                             # copies into the handle's HBM buffer
        del x
        b0 += block


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
    sidx.use_graph = world > 1 and os.environ.get("RS_SHARDED_GRAPH", "1") != "0"   # one CUDA-graph replay per step
    # both exchanges of a step through NVLink peer memory (csrc/peer.cu; "nccl" = two NCCL all-gathers)
    sidx.exchange = os.environ.get("RS_SHARDED_EXCHANGE", "p2p") if world > 1 else "nccl"

    gq = torch.Generator(device=device)
    gq.manual_seed(3999)
    q_dev = torch.randn((nq, dim), generator=gq, device=device, dtype=torch.float32)
    rs.normalize_L2(q_dev)
    q_pinned = torch.empty((nq, dim), dtype=torch.float32).pin_memory()
    q_pinned.copy_(q_dev.cpu())
    D_host = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    I_host = torch.empty((nq, k), dtype=torch.int64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return sidx.search(q_dev, k)

    def step_e2e():
        qd = q_pinned.to(device, non_blocking=True)              # H2D of this step's inputs (pinned)
        D, I = sidx.search(qd, k)
        D_host.copy_(D, non_blocking=True)                       # D2H of this step's result
        I_host.copy_(I, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return D_host, I_host

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    if rank == 0:
        sampler.start()                      # samples cover the warm-up and the timed region (both under load)
    # >= W warm-up steps, extended to ~1 s of load so clocks settle and the sampler has data.  The step count
    # must be identical on every rank (each step contains a collective), so it is derived from a MAX-reduced time.
    n_warm = max(3, args.warmup)
    warm_ms = timed(step_device, n_warm)
    extra = int(min(2000, max(0, (1000.0 - warm_ms) / max(warm_ms / n_warm, 1e-3))))
    if extra:
        timed(step_device, extra)
        n_warm += extra
    barrier()
    graphed = bool(getattr(sidx, "use_graph", False)) and world > 1
    if graphed:
        # the timed steps replay one CUDA graph per step (no library host code runs, so neither the launch counter nor
        # the per-kernel event timing sees them): time them plainly, then run the same number of eager steps for the
        # per-kernel figures -- same kernels, same inputs
        total_ms = timed(step_device, args.steps)
        clocks = sampler.stop() if rank == 0 else None
        sidx.use_graph = False
        timed(step_device, 2)
    rs._lib.prof_enable(True)
    launches0 = rs._lib.launch_count()
    eager_ms = timed(step_device, args.steps)
    launches = rs._lib.launch_count() - launches0
    prof = {name: rs._lib.prof_read(slot) for name, slot in
            (("scan", rs._lib.PROF_SCAN), ("merge", rs._lib.PROF_MERGE), ("filter", rs._lib.PROF_FILTER),
             ("rescore", rs._lib.PROF_RESCORE), ("sample", rs._lib.PROF_SAMPLE))}
    rs._lib.prof_enable(False)
    if graphed:
        sidx.use_graph = True
    else:
        total_ms = eager_ms
        clocks = sampler.stop() if rank == 0 else None

    for _ in range(max(3, args.warmup)):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

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
                        "hbm_frac": (n_local * dim * 2 / t / 1e9) / pk["hbm_gbs"]}
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
    if roofline:
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            ent = json.loads(tf.read_text()).get(roofline["kernel"])
            if ent and ent.get("n_local") == n_local and ent.get("nq") == nq:
                roofline["traffic"] = ent["dram_bytes_per_launch"]
                roofline["traffic_source"] = ent["source"]
    kernel_ms = {n_: {"ms_total": round(v[0], 4), "launches": v[1]} for n_, v in prof.items() if v[1]}

    # ---- second headline: DCN ranking forward at batch 65,536 (configs[1]) ----
    ranking = None
    if not args.no_ranking:
        ranking = bench_ranking(rs, torch, dist, args, device, world, rank, pk, timed)

    batch1 = None
    if not args.no_batch1:
        batch1 = bench_batch1(rs, torch, dist, sidx, args, device, world)

    # configs[0], the reference's own CPU-runnable case (MovieLens-1M-shaped, batch 4096): times only, on rank 0 at N=1;
    # its outputs are compared with the oracle by tests/test_gpu_parity.py.  Never allowed to take the headline down.
    config1 = None
    if world == 1 and not args.no_ranking:
        try:
            config1 = bench_config1(rs, torch, device, timed)
        except Exception as e:  # noqa: BLE001
            config1 = {"error": f"{type(e).__name__}: {e}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        xb_h, q_h = cpu_prepare(args, args.cpu_sample_rows)
        cpu = cpu_describe(args, xb_h.shape[0], cpu_step(args, xb_h, q_h))
        del xb_h

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"exact top-{k} over {N:,} x {dim} fp32 catalog, {nq}-query batches "
                                       f"(BASELINE configs[2])", "n_items": N, "dim": dim, "nq": nq, "k": k,
                           "mode": args.mode, "arithmetic": "bf16 tcgen05 pre-filter (fp32 accumulate) + exact fp32 FMA-chain re-score: results bit-identical to the fp32 scan", "sharding": (f"rows/{world} contiguous; per step: sample -> all-gather of the top lists -> shared-threshold filter + exact re-score -> "
                                        f"one fused all-gather -> merge -> verify" + ("; the step is one CUDA-graph replay (per-kernel times and gpu_launches from "
                                        "the same number of eager steps)" if graphed else "")) if world > 1 else "none",
                           "l2": "inputs larger than L2 (catalog shard %.2f GB per GPU)" % (n_local * dim * 4 / 1e9)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nq * dim * 4,
                        "d2h_bytes_per_step": nq * k * 12, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(lt.item()), "clocks": clocks, "roofline": roofline, "kernels": kernel_ms,
                "cpu_baseline": cpu, "ranking": ranking, "batch1_latency": batch1, "config1": config1}
        if world > 1:
            line["exchange_bytes_per_step_per_rank"] = gather_bytes(world, nq, k)
            line["config"]["exchange"] = ("NVLink peer memory: each rank stores its message into every peer's region and raises a flag "
                                          "(rs_peer_publish / wait / ack, no NCCL on the step)" if sidx.exchange == "p2p"
                                          else "two NCCL all-gathers per step")
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        sidx.close()                         # captured graphs hold NCCL work: release them before the group goes away
        dist.barrier()
        dist.destroy_process_group()


def bench_ranking(rs, torch, dist, args, device, world, rank, pk, timed):
    """MultiTaskModel forward (ids -> towers -> DCN -> heads) at batch 65,536 per GPU (data-parallel replicas; weak
    scaling, no collective).  Headline = bf16 tensor mode (tcgen05 fused chains, tolerance 1e-2 of scale vs the fp64
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
        # The eager figure is reported beside it; fp32 mode (1.1 ms of kernels) is timed eagerly.
        sess = model.ranking_session(B) if precision == "bf16" else None
        if sess is not None:
            sess.user_rows.copy_(ur_d)
            sess.item_rows.copy_(ir_d)
            step = sess.run

            def step_e2e():
                sess.user_rows.copy_(ur, non_blocking=True)
                sess.item_rows.copy_(ir, non_blocking=True)
                o = sess.run()
                out_host[0].copy_(o["rating_prediction"].view(-1), non_blocking=True)
                out_host[1].copy_(o["ctr_prediction"].view(-1), non_blocking=True)
                torch.cuda.current_stream().synchronize()
        else:
            step = step_eager

            def step_e2e():
                o = model({"user_id": ur.to(device, non_blocking=True), "movie_id": ir.to(device, non_blocking=True)})
                out_host[0].copy_(o["rating_prediction"].view(-1), non_blocking=True)
                out_host[1].copy_(o["ctr_prediction"].view(-1), non_blocking=True)
                torch.cuda.current_stream().synchronize()

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
        tc_ms, tc_n = rs._lib.prof_read(rs._lib.PROF_RANK)
        rs._lib.prof_enable(False)
        for _ in range(3):
            step_e2e()
        e2e_ms = timed(step_e2e, steps) / steps
        res[precision] = {"value": world * B / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms,
                          "eager_ms_per_step": eager_ms, "cuda_graph": sess is not None,
                          "e2e": {"value": world * B / (e2e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": 2 * B * 4,
                                  "d2h_bytes_per_step": B * 8},
                          "gemm_tflops_per_gpu": flops / (ms / 1e3) / 1e12,
                          "gpu_launches_per_step": launches // max(1, steps)}
        if precision == "bf16" and tc_n:
            t = tc_ms / steps / 1e3          # device time of the tcgen05 chain kernels per step
            res[precision]["roofline"] = {"kernel": "tower_pp_kernel + deep_heads_kernel (chain_tc.cu)", "bound": "tensor",
                                          "achieved": flops / t / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                          "frac": flops / t / 1e12 / pk["bf16_tflops_sustained"], "traffic": None,
                                          "algorithmic": "524,288 GEMM flop per sample", "ms_per_step": tc_ms / steps,
                                          "launches_per_step": tc_n / steps, "peak_source": pk["source"] + " (sustained bf16)"}
    head = res["bf16"]
    return {"metric": "DCN ranking forward samples/sec", "value": head["value"], "unit": "samples/s",
            "ms_per_step": head["ms_per_step"], "batch_per_gpu": B, "n_gpus": world, "scaling": "weak",
            "mode": "bf16 tensor mode (tcgen05, fp32 accumulate; tolerance 1e-2 of tensor scale)", "dtype": "bf16",
            "config": {"workload": f"MultiTaskModel forward, 3 cross layers + MLP, emb 128, batch {B} per GPU (BASELINE configs[1])"},
            "e2e": head["e2e"], "gemm_tflops_per_gpu": head["gemm_tflops_per_gpu"],
            "tensor_frac_of_sustained_bf16": head["gemm_tflops_per_gpu"] / pk["bf16_tflops_sustained"],
            "roofline": head.get("roofline"), "gpu_launches_per_step": head["gpu_launches_per_step"],
            "cuda_graph": head["cuda_graph"], "eager_ms_per_step": head["eager_ms_per_step"],
            "fp32_mode": {"value": res["fp32"]["value"], "ms_per_step": res["fp32"]["ms_per_step"], "dtype": "f32",
                          "tolerance": "1e-5 of tensor scale", "e2e": res["fp32"]["e2e"],
                          "gemm_tflops_per_gpu": res["fp32"]["gemm_tflops_per_gpu"]}}


def bench_config1(rs, torch, device, timed, steps=20):
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
    return {"workload": "MovieLens-1M-shaped synthetic (6,040 users / 3,706 items, emb 128, batch 4096): MultiTaskModel forward "
                        "in fp32 exact mode + exact top-10 over the 3,706-item catalog (BASELINE configs[0])",
            "forward_ms": f_ms, "forward_samples_per_s": B / (f_ms / 1e3), "top10_ms": s_ms,
            "top10_queries_per_s": B / (s_ms / 1e3), "note": "problem is L2-resident: launch/latency-bound, no roofline fraction"}


def bench_batch1(rs, torch, dist, sidx, args, device, world, n_queries=200, k=10):
    """BASELINE config 5: single-query top-10 over the (sharded) catalog, one query at a time.  Device time per query
    (CUDA events around each call) and end-to-end wall time (host query in, host (D, I) out), p50 / p99."""
    g = torch.Generator(device=device)
    g.manual_seed(5555)
    qs = torch.randn((n_queries, args.dim), generator=g, device=device, dtype=torch.float32)
    rs.normalize_L2(qs)
    qs_host = qs.cpu().pin_memory()
    for i in range(20):
        sidx.search(qs[i:i + 1].contiguous(), k)
    torch.cuda.synchronize()
    dev_ms, wall_ms = [], []
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
    dm = torch.tensor([float(np.percentile(dev_ms, 50)), float(np.percentile(dev_ms, 99)),
                       float(np.percentile(wall_ms, 50)), float(np.percentile(wall_ms, 99))], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dm, op=dist.ReduceOp.MAX)
    dm = dm.tolist()
    return {"workload": f"batch-1 top-{k} over {args.n_items:,} x {args.dim} (BASELINE configs[4]), {n_queries} queries one at a time",
            "device_ms_p50": dm[0], "device_ms_p99": dm[1], "e2e_ms_p50": dm[2], "e2e_ms_p99": dm[3],
            "queries_per_s_at_p50": 1e3 / dm[2]}


if __name__ == "__main__":
    main()
