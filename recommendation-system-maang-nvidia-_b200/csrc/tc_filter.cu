// tc_filter.cu -- K5 (tensor mode): tcgen05 bf16 pre-filter + exact fp32 re-score for batched retrieval.
//
// Same results as the exact scan (flat_index.cu) -- bit-exact scores, index-exact top-k -- at tensor-core
// speed.  The idea: a low-precision score can only *rank* approximately, but it bounds the exact score:
//   |s_hat(q,x) - s(q,x)| <= eps_q := c * ||q|| * max_i ||x_i||,   c = 2^-8 * 1.0625
// (bf16 round-to-nearest of both operands, unit roundoff 2^-9 each, Cauchy-Schwarz; the slack covers the
// fp32 accumulation of either side).  Hence every true top-k item has s_hat >= s_hat_(k) - 2*eps_q, where
// s_hat_(k) is the k-th largest low-precision score.  The pipeline per query batch:
//   1. prep:    queries -> bf16 (K-major tile layout for TMA), 2*eps_q
//   2. sample:  tc_filter_kernel<SAMPLE> over ~131k catalog rows -> per-query block maxima -> a threshold
//               guess tau_q (the r-th largest sample value; r chosen so that tau_q is below
//               s_hat_(k) - 2*eps_q with overwhelming probability)
//   3. filter:  tc_filter_kernel<FILTER> over the whole bf16 shadow catalog: S = Q * X^T on tcgen05
//               (128 x 256 x 128 tiles, accumulators in TMEM, operands TMA-staged with 128B swizzle); the
//               epilogue reads the accumulator tile straight out of TMEM, rejects against tau_q in
//               registers (3-input max tree, one compare per 32 scores) and appends the rare survivors to
//               per-query candidate lists.  Scores never reach HBM.
//   4. select:  per query: sort candidates by s_hat, VERIFY  s_hat_(k) - 2*eps_q >= tau_q  (this is what
//               makes the guess rigorous: if it holds, the list provably contains the whole top-k), keep
//               the prefix with s_hat >= s_hat_(k) - 2*eps_q, re-score those rows with the stated fp32 FMA
//               chain from the fp32 catalog, sort by (score desc, row asc), emit top-k.
//   5. queries that fail verification (or overflow their list) are re-run on the exact scan.  Rare.
//
// tc_filter_kernel roles (384 threads, 1 CTA per SM, persistent over item tiles):
//   warp 0 lane*: TMA producer        warp 1 lane*: tcgen05.mma issuer (+ TMEM alloc)
//   warps 4-11:   epilogue (TMEM lane quarter = warp % 4, column half = (warp - 4) / 4)
// Pipelines: smem full/empty per item stage (TMA <-> MMA), TMEM full/empty per accumulator (MMA <->
// epilogue; two 256-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1).
// Tensor roofline: 2*nq*N*128 flops per batch.  HBM: N*128*2 bytes (bf16 shadow), re-read from L2 by the
// CTA groups that hold different query tiles.
#include <cfloat>
#include <memory>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "flat_index.cuh"
#include "peer.cuh"
#include "tcgen05.cuh"

namespace {

using namespace rsd;

constexpr int BM = 128;                  // queries per tile == TMEM lanes
constexpr int BN = 256;                  // catalog rows per tile == TMEM columns per accumulator
constexpr int DK = 128;                  // feature dimension of the tensor path
constexpr int R_MAX = 2;                 // query tiles resident in shared memory per CTA
constexpr int NSTAGE = 2;                // catalog tile stages
constexpr int EPI_WARPS = 16;              // 4 TMEM lane quarters x EPI_SLICES column slices
constexpr int EPI_SLICES = EPI_WARPS / 4;
constexpr int SLICE_COLS = BN / EPI_SLICES; // accumulator columns per epilogue warp
constexpr int TC_THREADS = 128 + EPI_WARPS * 32;
constexpr int QTILE_BYTES = BM * DK * 2;  // 32 KB: two [128 x 64] bf16 swizzled halves
constexpr int STAGE_BYTES = BN * DK * 2;  // 64 KB: two [256 x 64] bf16 swizzled halves
constexpr int WQ_CAP = 64;               // survivor records per extraction-warp staging queue
constexpr int RING = 8;                  // hit chunks an epilogue warp may have handed over and not yet extracted
constexpr int RING_ENTRY = 144;          // 32 scores + {query row, first catalog row, threshold, -}
constexpr int CTRL_BYTES = 512;          // mbarriers, TMEM slot, ring heads / tails
constexpr int TC_SMEM = R_MAX * QTILE_BYTES + NSTAGE * STAGE_BYTES + CTRL_BYTES + EPI_WARPS * RING * RING_ENTRY +
                        2 * WQ_CAP * 16 + 1024;
static_assert(RING * RING_ENTRY >= WQ_CAP * 16 && (RING & (RING - 1)) == 0, "a warp's ring doubles as its dense-mode queue");
constexpr int CAND_CAP = 4096;           // candidate list capacity per query
constexpr int SEL_NT = 256;
constexpr float C_EPS = 0.00390625f * 1.0625f;

enum { MODE_FILTER = 0, MODE_SAMPLE = 1, MODE_DUMP = 2, MODE_SAMPLE_FINE = 3, MODE_FILTER_DENSE = 4 };

struct TcParams {
  int64_t n_items;
  int nq;
  int n_qtiles;
  int64_t n_visit;      // item tiles this launch visits
  int64_t tile_step;    // visited tile v -> catalog tile v * tile_step
  int ctas_per_group;
  int n_groups;
  const float* thr;     // [nq_pad]      FILTER
  uint64_t* cand;       // [nq_pad][cap] FILTER
  int* cand_cnt;        // [nq_pad]      FILTER
  int cap;
  int* overflow;        // FILTER: set when a staging queue overran
  unsigned* smax;       // [nq_pad][smax_ld] SAMPLE: ordered-uint maxima; slot = (visit % (smax_ld/S))*S + slice
  int smax_ld;
  float* dump;          // [nq_pad][dump_ld] DUMP (debug: every low-precision score)
  int64_t dump_ld;
  int debug;            // perf experiments only: 1 = epilogue skips TMEM reads, 2 = issuer skips the MMAs, 4 = no survivors, 8 = no flush
};

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_filter_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                 const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* qsm = base;
  uint8_t* xsm = base + R_MAX * QTILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xsm + NSTAGE * STAGE_BYTES);
  uint64_t* qfull = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = full + NSTAGE;
  uint64_t* tfull = empty + NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  // FILTER: hit chunks travel from the epilogue warps to the two extraction warps through per-warp rings
  volatile int* heads = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + 128);   // [EPI_WARPS] entries written
  volatile int* tails = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + 192);   // [EPI_WARPS] entries consumed
  volatile int* epi_done = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + 256);
  uint8_t* rings = reinterpret_cast<uint8_t*>(bars) + CTRL_BYTES;       // [EPI_WARPS][RING][RING_ENTRY]
  uint4* wqueue = reinterpret_cast<uint4*>(rings + EPI_WARPS * RING * RING_ENTRY);   // [2][WQ_CAP]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.x / p.ctas_per_group;
  const int j = blockIdx.x % p.ctas_per_group;
  if (group >= p.n_groups) return;
  const int qt0 = group * R_MAX;
  const int nr = min(R_MAX, p.n_qtiles - qt0);
  const int64_t cpg = p.ctas_per_group;
  const int64_t n_my = j < p.n_visit ? (p.n_visit - j + cpg - 1) / cpg : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_x);
    mbar_init(qfull, 1);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], EPI_WARPS);   // lane 0 of every epilogue warp arrives once it has drained its part
    }
    fence_mbar_init();
  }
  if (threadIdx.x < EPI_WARPS) {
    heads[threadIdx.x] = 0;
    tails[threadIdx.x] = 0;
  }
  if (threadIdx.x == 0) *epi_done = 0;
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer ----------------
      mbar_arrive_expect_tx(qfull, static_cast<uint32_t>(nr) * QTILE_BYTES);
      for (int r = 0; r < nr; ++r)
        for (int h = 0; h < 2; ++h)
          tma_load_2d(qsm + r * QTILE_BYTES + h * (QTILE_BYTES / 2), &tmap_q, qfull, h * 64, (qt0 + r) * BM, kEvictLast);
      for (int64_t i = 0; i < n_my; ++i) {
        const int s = static_cast<int>(i % NSTAGE);
        const uint32_t ph = static_cast<uint32_t>((i / NSTAGE) & 1);
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        const int64_t tile = (j + i * cpg) * p.tile_step;
        for (int h = 0; h < 2; ++h)
          tma_load_2d(xsm + s * STAGE_BYTES + h * (STAGE_BYTES / 2), &tmap_x, &full[s], h * 64,
                      static_cast<int>(tile * BN), kEvictNormal);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, BN);
      mbar_wait(qfull, 0);
      // operand descriptors = one base descriptor per region + small offsets: keeps the issuing thread's per-MMA
      // overhead to an add (it is a single thread feeding the whole tensor pipe)
      const uint64_t qdesc0 = umma_desc_k_sw128(smem_u32(qsm)), xdesc0 = umma_desc_k_sw128(smem_u32(xsm));
      uint32_t acc_it = 0;
      for (int64_t i = 0; i < n_my; ++i) {
        const int s = static_cast<int>(i % NSTAGE);
        const uint32_t ph = static_cast<uint32_t>((i / NSTAGE) & 1);
        mbar_wait(&full[s], ph);
        tc_fence_after_sync();
        for (int r = 0; r < nr; ++r, ++acc_it) {
          const uint32_t a = acc_it & 1u, aph = (acc_it >> 1) & 1u;
          mbar_wait(&tempty[a], aph ^ 1u);
          tc_fence_after_sync();
#pragma unroll
          for (int k = 0; k < DK / 16; ++k) {
            if (p.debug & 2) break;
            const int h = k >> 2;
            const uint32_t ko = static_cast<uint32_t>(k & 3) * 32u;
            const uint64_t da = qdesc0 + static_cast<uint64_t>((r * QTILE_BYTES + h * (QTILE_BYTES / 2) + static_cast<int>(ko)) >> 4);
            const uint64_t db = xdesc0 + static_cast<uint64_t>((s * STAGE_BYTES + h * (STAGE_BYTES / 2) + static_cast<int>(ko)) >> 4);
            umma_bf16_ss(tmem_base + a * BN, da, db, idesc, k > 0 ? 1u : 0u);
          }
          umma_commit(&tfull[a]);
        }
        umma_commit(&empty[s]);
      }
    }
  } else if (MODE == MODE_FILTER && (warp == 2 || warp == 3)) {
    // ---------------- survivor extraction ----------------
    // Survivors are rare (~1e-4 of the scores) but finding them inside a 32-score chunk is divergent, branchy code: done
    // by the epilogue lane that owns the row it cost 15 % of the kernel (a lone active lane walking compare chains between
    // two TMEM loads).  The epilogue only DETECTS a hit chunk (chunk maximum >= tau) and hands its 32 scores over through a
    // shared-memory ring; here 32 lanes test one score each, queue the survivors, and append a full queue to the per-query
    // candidate lists with 32 global atomics in flight (one round trip per 32 records).  A list that overruns is seen by
    // select_rescore_kernel (count > cap): that query is repeated exactly.
    const int cw = warp - 2;                       // epilogue warps [8 cw, 8 cw + 8) are this warp's producers
    uint4* wq = wqueue + cw * WQ_CAP;
    int qcnt = 0;
    int tail[EPI_WARPS / 2];
#pragma unroll
    for (int pw = 0; pw < EPI_WARPS / 2; ++pw) tail[pw] = 0;
    auto flush = [&]() {
      __syncwarp();
      for (int t = lane; t < qcnt; t += 32) {
        const uint4 r = wq[t];
        const int pos = atomicAdd(&p.cand_cnt[r.z], 1);
        if (pos < p.cap) p.cand[(int64_t)r.z * p.cap + pos] = (static_cast<uint64_t>(r.y) << 32) | r.x;
      }
      __syncwarp();
      qcnt = 0;
    };
    for (;;) {
      const bool finished = *epi_done == EPI_WARPS;   // read BEFORE the scan: nothing is published after it
      __threadfence_block();
      bool any = false;
#pragma unroll
      for (int pw = 0; pw < EPI_WARPS / 2; ++pw) {
        const int ew = cw * (EPI_WARPS / 2) + pw;
        const int head = heads[ew];
        if (head != tail[pw]) {   // uniform
          any = true;
          __threadfence_block();
          while (tail[pw] != head) {
            const uint8_t* ent = rings + (ew * RING + (tail[pw] & (RING - 1))) * RING_ENTRY;
            const float v = reinterpret_cast<const float*>(ent)[lane];
            const int4 hdr = *reinterpret_cast<const int4*>(ent + 128);
            const int64_t idx = static_cast<int64_t>(static_cast<uint32_t>(hdr.y)) + lane;
            const bool hit = v >= __int_as_float(hdr.z) && idx < p.n_items && !(p.debug & 16);
            const unsigned hb = __ballot_sync(0xFFFFFFFFu, hit);
            if (hit) {
              const uint64_t key = pack_key(v, static_cast<uint32_t>(idx));
              wq[qcnt + __popc(hb & ((1u << lane) - 1u))] =
                  make_uint4(static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), static_cast<uint32_t>(hdr.x), 0u);
            }
            qcnt += __popc(hb);
            ++tail[pw];
            __syncwarp();
            if (lane == 0) tails[ew] = tail[pw];      // the slot may be overwritten
            if (qcnt > WQ_CAP - 32) flush();
          }
        }
      }
      if (!any) {
        if (finished) break;
        __nanosleep(200);
      }
    }
    if (qcnt > 0) flush();
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    const int ew = warp - 4;
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter + 32) are this warp's
    const int slice = ew >> 2;             // columns [SLICE_COLS*slice, SLICE_COLS*(slice+1)) of the accumulator
    const int row_in_tile = quarter * 32 + lane;
    uint8_t* ring = rings + ew * RING * RING_ENTRY;   // FILTER: this warp's ring of hit chunks
    int head = 0;                                    // entries written to it (warp-uniform)
    // FILTER_DENSE: the ring space is this warp's survivor queue instead (WQ_CAP records), its head word the count
    uint4* dq = reinterpret_cast<uint4*>(ring);
    int* dq_cnt = const_cast<int*>(heads) + ew;
    auto dense_flush = [&](int n) {
      for (int t = lane; t < n; t += 32) {
        const uint4 r = dq[t];
        const int pos = atomicAdd(&p.cand_cnt[r.z], 1);
        if (pos < p.cap) p.cand[(int64_t)r.z * p.cap + pos] = (static_cast<uint64_t>(r.y) << 32) | r.x;
      }
      __syncwarp();
      if (lane == 0) *dq_cnt = 0;
      __syncwarp();
    };
    float thr[R_MAX];
    int qrow[R_MAX];
#pragma unroll
    for (int r = 0; r < R_MAX; ++r) {
      qrow[r] = (qt0 + r) * BM + row_in_tile;
      thr[r] = INFINITY;
      if ((MODE == MODE_FILTER || MODE == MODE_FILTER_DENSE) && r < nr && qrow[r] < p.nq && !(p.debug & 4)) thr[r] = p.thr[qrow[r]];
    }
    uint32_t acc_it = 0;
    for (int64_t i = 0; i < n_my; ++i) {
      const int64_t visit = j + i * cpg;
      const int64_t item0 = visit * p.tile_step * BN + slice * SLICE_COLS;
      for (int r = 0; r < nr; ++r, ++acc_it) {
        const uint32_t a = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        mbar_wait(&tfull[a], aph);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + a * BN + slice * SLICE_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
        float hmax = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < SLICE_COLS / 32; ++c) {
          if (p.debug & 1) break;
          float v[32];
          __syncwarp();
          tmem_ld_32x32(taddr + c * 32, v);
          if (c == SLICE_COLS / 32 - 1) {   // last chunk is in registers: hand the accumulator back before processing it
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[a]);
          }
          // group maxima over 8 consecutive scores with 3-input max: half an instruction per score
          float g[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            g[u] = max3(max3(v[8 * u], v[8 * u + 1], v[8 * u + 2]), max3(v[8 * u + 3], v[8 * u + 4], v[8 * u + 5]),
                        fmaxf(v[8 * u + 6], v[8 * u + 7]));
          const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
          if (MODE == MODE_FILTER) {
            // hit chunks (chunk maximum >= tau: ~7 % of a warp's chunks) go to the extraction warps as they are
            const bool hit = m >= thr[r] && !((p.debug & 32) && c == 0) && !((p.debug & 64) && c == 1);
            unsigned rem = __ballot_sync(0xFFFFFFFFu, hit);
            while (rem) {   // uniform; one pass unless the ring is nearly full
              const int room = RING - (head - __shfl_sync(0xFFFFFFFFu, tails[ew], 0));
              if (room <= 0) continue;   // the extraction warp is behind (adversarial data): wait for a slot
              const int mine = __popc(rem & ((1u << lane) - 1u));
              if (((rem >> lane) & 1u) && mine < room) {
                uint8_t* ent = ring + ((head + mine) & (RING - 1)) * RING_ENTRY;
#pragma unroll
                for (int e = 0; e < 32; e += 4)
                  *reinterpret_cast<float4*>(ent + e * 4) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                *reinterpret_cast<int4*>(ent + 128) =
                    make_int4(qrow[r], static_cast<int>(item0 + c * 32), __float_as_int(thr[r]), 0);
              }
              const int took = min(__popc(rem), room);
              for (int t = 0; t < took; ++t) rem &= rem - 1u;
              head += took;
              __syncwarp();
              if (lane == 0) {
                __threadfence_block();
                heads[ew] = head;
              }
            }
          } else if (MODE == MODE_FILTER_DENSE) {
            // small catalogs (every tile scored, k of a few thousand rows): most chunks hold a survivor, two extraction
            // warps would be the bottleneck -- every epilogue lane extracts its own and queues them per warp
            if (m >= thr[r]) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (g[u] >= thr[r]) {
#pragma unroll
                  for (int e = 8 * u; e < 8 * u + 8; ++e) {   // unrolled so v[] stays in registers
                    const int64_t idx = item0 + c * 32 + e;
                    if (v[e] >= thr[r] && idx < p.n_items) {
                      const int slot = atomicAdd(dq_cnt, 1);
                      if (slot < WQ_CAP) {
                        const uint64_t key = pack_key(v[e], static_cast<uint32_t>(idx));
                        dq[slot] = make_uint4(static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32),
                                              static_cast<uint32_t>(qrow[r]), 0u);
                      } else {
                        *p.overflow = 1;   // queue overrun (adversarial data): the host re-runs the chunk exactly
                      }
                    }
                  }
                }
              }
            }
          } else if (MODE == MODE_SAMPLE) {
            hmax = fmaxf(hmax, m);
          } else if (MODE == MODE_SAMPLE_FINE) {
            // small catalogs: EVERY tile is visited once and every 32-score chunk keeps its own maximum (no sharing of
            // slots, no atomics); rows past the catalog (zero-filled by TMA: score 0) must not count
            float mm = m;
            const int64_t i0 = item0 + c * 32;
            if (i0 + 32 > p.n_items) {
              mm = -INFINITY;
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (i0 + e < p.n_items) mm = fmaxf(mm, v[e]);
            }
            p.smax[(int64_t)qrow[r] * p.smax_ld + visit * (BN / 32) + slice * (SLICE_COLS / 32) + c] = f32_to_ordered(mm);
          } else {
            if (qrow[r] < p.nq) {
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const int64_t idx = item0 + c * 32 + e;
                if (idx < p.n_items) p.dump[(int64_t)qrow[r] * p.dump_ld + idx] = v[e];
              }
            }
          }
        }
        if (MODE == MODE_SAMPLE)   // several sampled tiles share a slot: keep the maximum
          atomicMax(&p.smax[(int64_t)qrow[r] * p.smax_ld + (visit % (p.smax_ld / EPI_SLICES)) * EPI_SLICES + slice],
                    f32_to_ordered(hmax));
        if (p.debug & 1) {   // (experiment: no TMEM reads) still release the accumulator
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[a]);
        }
        __syncwarp();
        if (MODE == MODE_FILTER_DENSE) {
          const int n = min(*dq_cnt, WQ_CAP);   // uniform: read after the __syncwarp above
          if (n >= 32) dense_flush(n);
        }
      }
    }
    if (MODE == MODE_FILTER_DENSE) {
      __syncwarp();
      const int n = min(*dq_cnt, WQ_CAP);
      if (n > 0) dense_flush(n);
    }
    if (MODE == MODE_FILTER) {   // every entry of this warp is published: tell the extraction warps
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        atomicAdd(const_cast<int*>(epi_done), 1);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// small kernels around the filter
// ------------------------------------------------------------------------------------------------

// fp32 catalog -> bf16 shadow (round to nearest even) + max row norm; one warp per row, d = 128
__global__ void __launch_bounds__(256) build_shadow_kernel(const float* __restrict__ x, uint16_t* __restrict__ xb,
                                                           int64_t row0, int64_t n, unsigned int* max_norm_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float best = 0.0f;
  for (int64_t r = row0 + warp; r < n; r += nwarps) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * DK) + lane);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(xb + r * DK)[lane] = o;
    float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o2);
    best = fmaxf(best, s);
  }
  if (lane == 0 && best > 0.0f) atomicMax(max_norm_bits, __float_as_uint(sqrtf(best) * 1.00001f));
}

// queries -> bf16 rows (zero rows for the padding) and 2*eps_q
__global__ void __launch_bounds__(128) prep_queries_kernel(const float* __restrict__ q, int nq, int nq_pad,
                                                           uint16_t* __restrict__ qb, float* __restrict__ eps2,
                                                           const unsigned int* max_norm_bits) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= nq_pad) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < nq) v = __ldg(reinterpret_cast<const float4*>(q + (int64_t)row * DK) + lane);
  uint2 o;
  o.x = pack_bf16x2(v.x, v.y);
  o.y = pack_bf16x2(v.z, v.w);
  reinterpret_cast<uint2*>(qb + (int64_t)row * DK)[lane] = o;
  float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o2);
  if (lane == 0) eps2[row] = 2.0f * C_EPS * (sqrtf(s) * 1.00001f) * __uint_as_float(*max_norm_bits);
}

// tau_q = the `rank`-th largest of the query's sample maxima; also clears the candidate counters
__global__ void __launch_bounds__(256) threshold_kernel(const unsigned* __restrict__ smax, int n_vals, int rank,
                                                        float* __restrict__ thr, int* __restrict__ cand_cnt,
                                                        int* __restrict__ flags) {
  __shared__ unsigned vals[1024];
  __shared__ unsigned hist[256];
  __shared__ unsigned state[2];
  const int tid = threadIdx.x;
  const int qi = blockIdx.x;
  for (int i = tid; i < n_vals; i += 256) vals[i] = smax[(int64_t)qi * n_vals + i];
  __syncthreads();
  const uint32_t kth = block_kth_largest<256>([&](int i) { return vals[i]; }, n_vals, rank, hist, state, tid);
  if (tid == 0) {
    thr[qi] = ordered_to_f32(kth);
    cand_cnt[qi] = 0;
    flags[qi] = 0;
  }
}

// Small catalogs (every tile sampled, one maximum per 32-row chunk): tau_q = (k-th largest chunk maximum) - 2 eps_q.  At
// least k rows score >= the k-th largest chunk maximum, so s_hat_(k) >= it, and every row of the true top-k has
// s_hat >= s_hat_(k) - 2 eps >= tau: the candidate list provably holds the top-k (select_rescore_kernel's verification
// s_hat_(k) - 2 eps >= tau then always passes).  Fewer real chunks than k: no threshold (the catalog is below 4,096 rows).
__global__ void __launch_bounds__(256) threshold_small_kernel(const unsigned* __restrict__ smax, int n_vals, int n_real, int k,
                                                              const float* __restrict__ eps2, float* __restrict__ thr,
                                                              int* __restrict__ cand_cnt, int* __restrict__ flags) {
  __shared__ unsigned vals[1024];
  __shared__ unsigned hist[256];
  __shared__ unsigned state[2];
  const int tid = threadIdx.x;
  const int qi = blockIdx.x;
  for (int i = tid; i < n_vals; i += 256) vals[i] = smax[(int64_t)qi * n_vals + i];
  __syncthreads();
  float t = -FLT_MAX;
  if (k <= n_real) {   // uniform
    const uint32_t kth = block_kth_largest<256>([&](int i) { return vals[i]; }, n_vals, k, hist, state, tid);
    t = ordered_to_f32(kth) - eps2[qi];
  }
  if (tid == 0) {
    thr[qi] = t;
    cand_cnt[qi] = 0;
    flags[qi] = 0;
  }
}

constexpr int SEL_MAX_M = 1024;   // most survivors of the margin cut that one query may re-score

// Where the kernels of a sharded step put what the other ranks need.  bases == nullptr: plain local outputs.  Otherwise the
// kernel PUBLISHES: it writes its part into slot `rank` of every rank's peer-memory region (csrc/peer.cu) -- what a separate
// peer_publish_kernel launch would copy there -- after the peers have acknowledged the previous step, and its last CTA
// raises this rank's flag in every region and then waits until every rank's message of this step has arrived here (one
// warp spins; the consumer kernel queued behind finds its data present).
struct PeerOut {
  uint8_t* const* bases = nullptr;
  rsp::PeerLayout PL{};
  int64_t off_I = 0, off_F = 0, off_B = 0;   // lists message: scores f32 [nq, k] at 0 | rows int64 [nq, k] | flags | bounds
};

__device__ __forceinline__ void publish_done(const PeerOut& out, uint64_t seq, int* s_last) {
  if (threadIdx.x == 0) *s_last = 0;
  __syncthreads();
  rsp::peer_finish(out.bases, out.PL, seq, nullptr, nullptr, 0, s_last);
  __syncthreads();
  if (*s_last && threadIdx.x < out.PL.world)
    rsp::spin_until(reinterpret_cast<const uint64_t*>(out.bases[out.PL.rank] + out.PL.off_flags) + threadIdx.x, seq,
                    out.PL.spin_budget, "a peer's message");
}

// step 4: verify, cut, exact re-score, final order.  One CTA per query.
__global__ void __launch_bounds__(SEL_NT) select_rescore_kernel(
    const uint64_t* __restrict__ cand, const int* __restrict__ cand_cnt, int cap, const float* __restrict__ thr,
    const float* __restrict__ eps2, const float* __restrict__ q, const float* __restrict__ xb, int k, int64_t id_base,
    float* __restrict__ D, int64_t ldD, int64_t* __restrict__ I, int64_t ldI, int* __restrict__ flags,
    int* __restrict__ n_flagged, float* __restrict__ bound, const int* __restrict__ overflow, int* __restrict__ need_exact,
    const PeerOut out, const int debug) {
  // bound == nullptr: the threshold was this index's own guess and is verified here (s_hat_(k) - 2 eps >= tau).
  // bound != nullptr ("shared threshold" mode of a partitioned catalog): tau came from the union of all shards'
  // samples, so this shard may hold fewer than k survivors.  It returns the exact top-min(c, k) of its survivors
  // (padded with D = -FLT_MAX, I = -1) and bound[q] = tau + eps; the caller verifies AFTER merging the shards that
  // the k-th merged score >= every shard's bound (a row that failed a shard's filter has s_hat < tau, hence an exact
  // score < tau + eps, and cannot belong to the top-k then).
  // need_exact != nullptr: need_exact[q] = 1 when the row of D / I is not final (the guess was not verified, a list or a
  // staging queue of the filter overran) -- what the host reads from `flags` in the synchronous form.
  // out.bases != nullptr: D / I / need_exact / bound are not used; the query's part of the lists message goes to every rank.
  __shared__ uint64_t keys[CAND_CAP];
  __shared__ __align__(16) uint32_t sel[SEL_MAX_M];
  __shared__ __align__(16) float qs[DK];
  __shared__ unsigned hist[256];
  __shared__ unsigned state[2];
  __shared__ int s_m;
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int qi = blockIdx.x;
  uint64_t seq = 0;
  if (out.bases) seq = rsp::peer_begin(out.bases[out.PL.rank], out.PL);
  const int c_raw = cand_cnt[qi];
  const int c = min(c_raw, cap);
  if (tid < DK) qs[tid] = q[(int64_t)qi * DK + tid];
  if (tid == 0) s_m = 0;
  for (int i = tid; i < c; i += SEL_NT) keys[i] = cand[(int64_t)qi * cap + i];
  __syncthreads();
  const bool shared_thr = bound != nullptr || out.bases != nullptr;
  bool ok = c_raw <= cap && (shared_thr || c >= k);
  float cutoff = -INFINITY;
  if (ok && c >= k) {   // uniform
    const uint32_t kth = block_kth_largest<SEL_NT>([&](int i) { return static_cast<uint32_t>(keys[i] >> 32); }, c, k, hist,
                                                   state, tid);
    cutoff = ordered_to_f32(kth) - eps2[qi];
    if (!shared_thr) ok = cutoff >= thr[qi];   // the guess was low enough: the list provably holds the whole top-k
  }
  if (ok) {
    // survivors of the margin cut: s_hat >= s_hat_(k) - 2 eps
    for (int i = tid; i < c; i += SEL_NT)
      if (key_score(keys[i]) >= cutoff) {
        const int pos = atomicAdd(&s_m, 1);
        if (pos < SEL_MAX_M) sel[pos] = key_row(keys[i]);
      }
    __syncthreads();
    ok = s_m <= SEL_MAX_M;
  }
  // the sorted list is staged in shared memory (sel[] is free after the re-score) and leaves the CTA in coalesced stores
  float* oD = reinterpret_cast<float*>(sel);            // [128]
  int64_t* oI = reinterpret_cast<int64_t*>(sel + 128);  // [128]
  if (ok) {
    const int m = s_m;
    // (The re-score is a gather of m random 512-byte rows out of the fp32 catalog: 2/3 of this kernel's time at 10M rows,
    // ~1.3 TB/s.  Asking L2 for all lines of all rows up front with prefetch.global.L2 measured 5 % slower: the rate of
    // the random gather itself bounds it, not the per-thread latency chain.)
    // exact re-score with the stated fp32 FMA chain; keys[] is free again and receives the exact keys
    for (int i = tid; i < m; i += SEL_NT) {
      const uint32_t row = sel[i];
      const float4* xr = reinterpret_cast<const float4*>(xb + (int64_t)row * DK);
      float acc = 0.0f;
#pragma unroll 16
      for (int c4 = 0; c4 < ((debug & 1) ? 0 : DK / 4); ++c4) {   // (experiment 1: no re-score loads / chain)
        const float4 xv = __ldg(xr + c4);
        const float4 qv = *reinterpret_cast<const float4*>(qs + c4 * 4);
        acc = fmaf(qv.x, xv.x, acc);
        acc = fmaf(qv.y, xv.y, acc);
        acc = fmaf(qv.z, xv.z, acc);
        acc = fmaf(qv.w, xv.w, acc);
      }
      keys[i] = pack_key(acc, row);
    }
    __syncthreads();
    // final order by counting: keys are distinct (the row is part of the key), so rank = #better keys
    for (int i = tid; i < m; i += SEL_NT) {
      const uint64_t mine = keys[i];
      int rank = (debug & 2) ? i : 0;                           // (experiment 2: no rank loop)
      for (int jx = 0; jx < ((debug & 2) ? 0 : m); ++jx) rank += keys[jx] > mine ? 1 : 0;
      if (rank < k) {
        oD[rank] = key_score(mine);
        oI[rank] = id_base + static_cast<int64_t>(key_row(mine));
      }
    }
    for (int i = m + tid; i < k; i += SEL_NT) {   // shared-threshold mode only: fewer than k survivors in this shard
      oD[i] = -FLT_MAX;
      oI[i] = -1;
    }
    __syncthreads();
  } else if (tid == 0) {
    flags[qi] = 1;
    atomicAdd(n_flagged, 1);
  }
  const int need = (!ok || *overflow != 0) ? 1 : 0;
  const float bnd = thr[qi] + 0.5f * eps2[qi];
  if (!out.bases) {
    if (ok)
      for (int i = tid; i < k; i += SEL_NT) {
        D[qi * ldD + i] = oD[i];
        I[qi * ldI + i] = oI[i];
      }
    if (tid == 0) {
      if (bound) bound[qi] = bnd;
      if (need_exact) need_exact[qi] = need;
    }
    return;
  }
  for (int g = 0; g < out.PL.world; ++g) {
    uint8_t* slot = out.bases[g] + out.PL.rank * out.PL.slot_bytes;
    if (ok) {
      float* gD = reinterpret_cast<float*>(slot) + (int64_t)qi * k;
      int64_t* gI = reinterpret_cast<int64_t*>(slot + out.off_I) + (int64_t)qi * k;
      for (int i = tid; i < k; i += SEL_NT) gD[i] = oD[i];
      for (int i = tid; i < k; i += SEL_NT) gI[i] = oI[i];
    }
    if (tid == 0) {
      reinterpret_cast<int*>(slot + out.off_F)[qi] = need;
      reinterpret_cast<float*>(slot + out.off_B)[qi] = bnd;
    }
  }
  publish_done(out, seq, &s_last);
}

// The `R` largest of each query's sample maxima, as floats, in no particular order (what a shard contributes to the
// shared threshold of a partitioned catalog).  R <= n_vals <= 1024.  pub.bases != nullptr: the list goes
// into slot `rank` of every rank's tops region (row stride R floats) instead of `out` -- see PeerOut.
__global__ void __launch_bounds__(256) sample_top_kernel(const unsigned* __restrict__ smax, int n_vals, int R,
                                                         float* __restrict__ out, int64_t ld_out, const PeerOut pub) {
  __shared__ unsigned vals[1024];
  __shared__ unsigned hist[256];
  __shared__ unsigned state[2];
  __shared__ float top[1024];
  __shared__ int n_gt;
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int qi = blockIdx.x;
  uint64_t seq = 0;
  if (pub.bases) seq = rsp::peer_begin(pub.bases[pub.PL.rank], pub.PL);
  for (int i = tid; i < n_vals; i += 256) vals[i] = smax[(int64_t)qi * n_vals + i];
  if (tid == 0) n_gt = 0;
  __syncthreads();
  const uint32_t kth = block_kth_largest<256>([&](int i) { return vals[i]; }, n_vals, R, hist, state, tid);
  for (int i = tid; i < n_vals; i += 256)
    if (vals[i] > kth) top[atomicAdd(&n_gt, 1)] = ordered_to_f32(vals[i]);   // fewer than R of these
  __syncthreads();
  for (int i = n_gt + tid; i < R; i += 256) top[i] = ordered_to_f32(kth);     // the rest are ties at the R-th value
  __syncthreads();
  if (!pub.bases) {
    for (int i = tid; i < R; i += 256) out[(int64_t)qi * ld_out + i] = top[i];
    return;
  }
  for (int g = 0; g < pub.PL.world; ++g) {
    float* o = reinterpret_cast<float*>(pub.bases[g] + pub.PL.rank * pub.PL.slot_bytes) + (int64_t)qi * R;
    for (int i = tid; i < R; i += 256) o[i] = top[i];
  }
  publish_done(pub, seq, &s_last);
}

// tau_q = the `rank`-th largest of the union of the shards' top-R lists: tops[g * part_ld + q * R + i]
// peer_bases != nullptr: `tops` is the local peer-memory region of a tops exchange (csrc/peer.cu); the kernel itself waits
// for every rank's message at entry and acknowledges the step to the peers once all CTAs have read their part (what
// separate rs_peer_wait / rs_peer_ack launches would do).  The region is written by other GPUs: plain loads only.
__global__ void __launch_bounds__(256) threshold_shared_kernel(const float* tops, int n_parts, int64_t part_ld,
                                                               int R, int rank, float* __restrict__ thr,
                                                               int* __restrict__ cand_cnt, int* __restrict__ flags,
                                                               uint8_t* const* __restrict__ peer_bases, const rsp::PeerLayout PL) {
  __shared__ unsigned vals[1024];
  __shared__ unsigned hist[256];
  __shared__ unsigned state[2];
  const int tid = threadIdx.x;
  const int qi = blockIdx.x;
  const int n = n_parts * R;
  uint64_t seq = 0;
  if (peer_bases) seq = rsp::peer_arrived(peer_bases[PL.rank], PL);
  for (int i = tid; i < n; i += 256) vals[i] = f32_to_ordered(tops[(i / R) * part_ld + (int64_t)qi * R + (i % R)]);
  __syncthreads();
  if (peer_bases) rsp::peer_consumed(peer_bases, PL, seq);
  const uint32_t kth = block_kth_largest<256>([&](int i) { return vals[i]; }, n, rank, hist, state, tid);
  if (tid == 0) {
    thr[qi] = ordered_to_f32(kth);
    cand_cnt[qi] = 0;
    flags[qi] = 0;
  }
}

struct Workspace {
  uint16_t* qb = nullptr;
  float* eps2 = nullptr;
  float* thr = nullptr;
  unsigned* smax = nullptr;
  int* cnt = nullptr;
  int* flags = nullptr;   // [nq_pad + 4]: per-query flag, then {n_flagged, queue overflow, -, -}
  uint64_t* cand = nullptr;
  cudaStream_t st;
  explicit Workspace(cudaStream_t s) : st(s) {}
  ~Workspace() {
    cudaFreeAsync(qb, st);
    cudaFreeAsync(eps2, st);
    cudaFreeAsync(thr, st);
    cudaFreeAsync(smax, st);
    cudaFreeAsync(cnt, st);
    cudaFreeAsync(flags, st);
    cudaFreeAsync(cand, st);
  }
};

void ensure_shadow(rs_index* h, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->shadow_rows == h->ntotal && h->xb_bf16) return;
  if (!h->xb_bf16) {
    RS_CUDA(cudaMalloc(&h->xb_bf16, sizeof(uint16_t) * (size_t)h->capacity * DK));
    RS_CUDA(cudaMalloc(&h->row_norm, sizeof(float) * 4));   // [0] holds max ||x|| as float bits
    RS_CUDA(cudaMemsetAsync(h->row_norm, 0, sizeof(float) * 4, st));
    h->shadow_rows = 0;
  }
  const int64_t todo = h->ntotal - h->shadow_rows;
  const int64_t want = (todo * 32 + 255) / 256;
  const int blocks = static_cast<int>(std::min<int64_t>(want, (int64_t)rs::sm_count() * 16));
  build_shadow_kernel<<<blocks, 256, 0, st>>>(h->xb, h->xb_bf16, h->shadow_rows, h->ntotal,
                                              reinterpret_cast<unsigned int*>(h->row_norm));
  RS_LAUNCH_CHECK();
  RS_CUDA(cudaMemcpyAsync(&h->max_norm, h->row_norm, sizeof(float), cudaMemcpyDeviceToHost, st));
  RS_CUDA(cudaStreamSynchronize(st));
  h->shadow_rows = h->ntotal;
  h->tmap_bf16 = rs::make_tmap_2d(h->xb_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)h->ntotal, DK,
                                  (uint64_t)DK * 2, BN, 64);
}

template <int MODE>
void launch_filter(const CUtensorMap& tq, const CUtensorMap& tx, TcParams p, int device, cudaStream_t st) {
  static std::once_flag once[64];
  std::call_once(once[device & 63], [] {
    cudaFuncSetAttribute(tc_filter_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  });
  if (const char* e = getenv("RS_TC_DEBUG")) p.debug = atoi(e);
  const int grid = p.n_groups * p.ctas_per_group;
  tc_filter_kernel<MODE><<<grid, TC_THREADS, TC_SMEM, st>>>(tq, tx, p);
  RS_LAUNCH_CHECK();
}

}  // namespace

namespace rs {

// Catalogs of 256 .. 32,768 rows (the reference's own catalog has 3,706 items): no sampling -- every tile is scored twice
// (chunk maxima -> threshold, then the filter), both passes out of L2.  The filter's survivor path is built for RARE
// survivors (64-record staging queue per warp and tile), so the top-k itself must be sparse in the catalog: k <= N / 128
// (k = 10 over 3,706 rows: 0.27 %); denser requests stay on the exact scan.
bool tensor_small_catalog(int64_t ntotal, int k) { return ntotal >= 256 && ntotal <= 32768 && ntotal >= 128ll * k; }

void ensure_tensor_shadow(rs_index* h, cudaStream_t st) { ensure_shadow(h, st); }

bool tensor_path_supported(const rs_index* h, int64_t nq, int k) {
  return h->d == DK && k >= 1 && k <= 128 && (h->ntotal >= 65536 || tensor_small_catalog(h->ntotal, k)) && nq >= 1 &&
         h->ntotal <= 0x7FFFFF00ll;
}

// flags layout of a search chunk: [0, nq) per-query "guess not verified", [nq_pad] their count, [nq_pad + 1] a staging
// queue of the filter overran (nothing is trusted then)

// Rows covered by the threshold sample of an index with `ntotal` rows: ~5 % of its 256-row tiles (64..2048 tiles),
// never the zero-padded tail tile.
static void sample_plan(int64_t ntotal, int64_t* n_sample, int64_t* sample_step) {
  const int64_t ntiles = (ntotal + BN - 1) / BN;
  *n_sample = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(2048, std::max<int64_t>(64, ntiles / 20)), (ntiles - 1) / 2));
  *sample_step = (ntiles - 1) / *n_sample;
}
int64_t tensor_sample_rows(int64_t ntotal) {
  int64_t n_sample, step;
  sample_plan(ntotal, &n_sample, &step);
  return n_sample * BN;
}
// rank of the sample maximum used as the threshold guess, for a sample covering the fraction rho of the rows
int tensor_threshold_rank(int k, double rho) {
  return static_cast<int>(std::ceil(std::pow(2.0 + std::sqrt(4.0 + 2.0 * k * rho), 2.0)));
}
int64_t tensor_max_queries() { return (int64_t)sm_count() * R_MAX * BM; }

}  // namespace rs

// One tensor-path search of <= tensor_max_queries() queries, split where the threshold is chosen:
//   tc_begin   prep queries, sample pass (block maxima of ~5 % of the tiles)
//   ...        threshold: this index's own guess (threshold_kernel) or one shared by all shards of a partitioned
//              catalog (sample_top_kernel -> all-gather by the caller -> threshold_shared_kernel)
//   tc_finish  filter pass, bucket, select + exact re-score
struct rs_tc_session {
  rs_index* h;
  const float* q;
  int nq, k, nq_pad, n_vals;
  int64_t ntiles, n_sample;
  cudaStream_t st;
  Workspace ws;
  CUtensorMap tq;
  TcParams p;
  rs_tc_session(cudaStream_t s) : st(s), ws(s) {}
};

namespace rs {

std::unique_ptr<rs_tc_session> tc_begin(rs_index* h, const float* qc, int cq, int k, cudaStream_t st) {
  RS_CHECK(tensor_path_supported(h, cq, k) && cq <= tensor_max_queries(), "tensor path does not apply to this shape");
  ensure_shadow(h, st);
  const int sms = sm_count();
  auto sp = std::make_unique<rs_tc_session>(st);
  rs_tc_session& s = *sp;
  s.h = h; s.q = qc; s.nq = cq; s.k = k;
  const int nq_pad = s.nq_pad = (cq + BM - 1) / BM * BM;
  const int n_qtiles = nq_pad / BM;
  s.ntiles = (h->ntotal + BN - 1) / BN;
  const bool small = h->ntotal < 65536;
  int64_t sample_step;
  if (small) {
    s.n_sample = s.ntiles;
    sample_step = 1;
  } else {
    sample_plan(h->ntotal, &s.n_sample, &sample_step);
  }
  const int n_vals = s.n_vals = small ? static_cast<int>(s.ntiles * (BN / 32))
                                      : static_cast<int>(std::min<int64_t>(s.n_sample, 1024 / EPI_SLICES) * EPI_SLICES);
  Workspace& ws = s.ws;
  RS_CUDA(cudaMallocAsync(&ws.qb, sizeof(uint16_t) * (size_t)nq_pad * DK, st));
  RS_CUDA(cudaMallocAsync(&ws.eps2, sizeof(float) * nq_pad, st));
  RS_CUDA(cudaMallocAsync(&ws.thr, sizeof(float) * nq_pad, st));
  RS_CUDA(cudaMallocAsync(&ws.smax, sizeof(unsigned) * (size_t)nq_pad * n_vals, st));
  RS_CUDA(cudaMemsetAsync(ws.smax, 0, sizeof(unsigned) * (size_t)nq_pad * n_vals, st));
  RS_CUDA(cudaMallocAsync(&ws.cnt, sizeof(int) * nq_pad, st));
  RS_CUDA(cudaMallocAsync(&ws.flags, sizeof(int) * (nq_pad + 4), st));
  RS_CUDA(cudaMallocAsync(&ws.cand, sizeof(uint64_t) * (size_t)nq_pad * CAND_CAP, st));
  RS_CUDA(cudaMemsetAsync(ws.flags + nq_pad, 0, sizeof(int) * 4, st));
  {
    ProfScope prof(RS_PROF_THRESHOLD, st);
    prep_queries_kernel<<<(nq_pad * 32 + 127) / 128, 128, 0, st>>>(qc, cq, nq_pad, ws.qb, ws.eps2,
                                                                  reinterpret_cast<unsigned int*>(h->row_norm));
    RS_LAUNCH_CHECK();
  }
  s.tq = make_tmap_2d(ws.qb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)nq_pad, DK, (uint64_t)DK * 2, BM, 64);
  TcParams& p = s.p;
  p = TcParams{};
  p.n_items = h->ntotal;
  p.nq = cq;
  p.n_qtiles = n_qtiles;
  p.n_groups = (n_qtiles + R_MAX - 1) / R_MAX;
  p.ctas_per_group = std::max(1, sms / p.n_groups);
  p.thr = ws.thr;
  p.cand = ws.cand;
  p.cand_cnt = ws.cnt;
  p.cap = CAND_CAP;
  p.smax = ws.smax;
  p.smax_ld = n_vals;
  p.overflow = ws.flags + nq_pad + 1;
  {
    ProfScope prof(RS_PROF_SAMPLE, st);
    TcParams ps = p;
    ps.n_visit = s.n_sample;
    ps.tile_step = sample_step;
    if (small) launch_filter<MODE_SAMPLE_FINE>(s.tq, h->tmap_bf16, ps, h->device, st);
    else launch_filter<MODE_SAMPLE>(s.tq, h->tmap_bf16, ps, h->device, st);
  }
  return sp;
}

// the threshold must have been written to s.ws.thr (and cnt / flags cleared) by one of the threshold kernels
static int sel_debug() {   // RS_SEL_DEBUG: timing experiments on select_rescore_kernel only (results are then WRONG)
  const char* e = getenv("RS_SEL_DEBUG");
  return e ? atoi(e) : 0;
}

// expect: candidates per query this index is expected to keep (threshold rank / sampled fraction; a shard's share of it
// under a shared threshold).  It picks the survivor path: a warp's 32 x 32-score chunk holds a hit with probability
// ~1,024 * expect / rows -- the ring hand-off to the two extraction warps wins while hit chunks are rare (6.8 % for 660
// candidates over 10M rows, and the same for each shard of a shared-threshold search); from ~15 % on the two warps fall
// behind the sixteen producers (1.25M rows behind one threshold of their own: 54 %, 0.54 vs 0.385 ms) and every epilogue
// lane extracts its own hits instead.
void tc_finish(rs_tc_session& s, int64_t id_base, float* D, int64_t ldD, int64_t* I, int64_t ldI, float* bound, int* need_exact,
               double expect, const PeerOut& out = PeerOut{}) {
  rs_index* h = s.h;
  cudaStream_t st = s.st;
  {
    ProfScope prof(RS_PROF_FILTER, st);
    TcParams pm = s.p;
    pm.n_visit = s.ntiles;
    pm.tile_step = 1;
    static const int force = [] { const char* e = getenv("RS_TC_SURVIVORS"); return e ? atoi(e) : 0; }();   // 1 = rings, 2 = inline (A/B)
    const bool dense = force == 2 || (force != 1 && (h->ntotal < 65536 || 1024.0 * expect > 0.15 * double(h->ntotal)));
    if (dense) launch_filter<MODE_FILTER_DENSE>(s.tq, h->tmap_bf16, pm, h->device, st);
    else launch_filter<MODE_FILTER>(s.tq, h->tmap_bf16, pm, h->device, st);
  }
  ProfScope prof(RS_PROF_RESCORE, st);
  select_rescore_kernel<<<s.nq, SEL_NT, 0, st>>>(s.ws.cand, s.ws.cnt, CAND_CAP, s.ws.thr, s.ws.eps2, s.q, h->xb, s.k, id_base,
                                                 D, ldD, I, ldI, s.ws.flags, s.ws.flags + s.nq_pad, bound,
                                                 s.ws.flags + s.nq_pad + 1, need_exact, out, sel_debug());
  RS_LAUNCH_CHECK();
}

void search_tensor_device(rs_index* h, const float* q, int64_t nq, int k, int64_t id_base, float* D, int64_t ldD,
                          int64_t* I, int64_t ldI, cudaStream_t st, int* need_exact) {
  RS_CHECK(tensor_path_supported(h, nq, k), "tensor path does not apply to this shape");
  if (bf16_scan_applies(h, nq, k)) {   // batch-1 / batch-2: two launches instead of seven, same bytes, always exact
    search_bf16_scan_device(h, q, nq, k, id_base, D, ldD, I, ldI, st, need_exact);
    return;
  }
  // query chunks so that every CTA group gets at least one SM
  const int64_t max_chunk = tensor_max_queries();
  for (int64_t q0 = 0; q0 < nq; q0 += max_chunk) {
    const int cq = static_cast<int>(std::min<int64_t>(max_chunk, nq - q0));
    const float* qc = q + q0 * DK;
    auto sp = tc_begin(h, qc, cq, k, st);
    rs_tc_session& s = *sp;
    const int nq_pad = s.nq_pad;
    Workspace& ws = s.ws;
    if (h->ntotal < 65536) {
      ProfScope prof(RS_PROF_THRESHOLD, st);
      threshold_small_kernel<<<cq, 256, 0, st>>>(ws.smax, s.n_vals, static_cast<int>((h->ntotal + 31) / 32), k, ws.eps2, ws.thr,
                                                 ws.cnt, ws.flags);
      RS_LAUNCH_CHECK();
    } else {
      const double rho = double(s.n_sample * BN) / double(h->ntotal);
      const int rank = std::max(1, std::min(tensor_threshold_rank(k, rho), s.n_vals / 2));
      ProfScope prof(RS_PROF_THRESHOLD, st);
      threshold_kernel<<<cq, 256, 0, st>>>(ws.smax, s.n_vals, rank, ws.thr, ws.cnt, ws.flags);
      RS_LAUNCH_CHECK();
    }
    const double rho_own = double(s.n_sample * BN) / double(h->ntotal);
    const double expect = h->ntotal < 65536 ? double(h->ntotal) : double(tensor_threshold_rank(k, rho_own)) / rho_own;
    tc_finish(s, id_base, D + q0 * ldD, ldD, I + q0 * ldI, ldI, nullptr, need_exact ? need_exact + q0 : nullptr, expect);
    if (need_exact) continue;   // deferred: the caller resolves flagged queries (after its own exchange step)
    // step 5: the (rare) queries whose guess could not be verified go through the exact scan
    std::vector<int> flags(nq_pad + 4);
    RS_CUDA(cudaMemcpyAsync(flags.data(), ws.flags, sizeof(int) * (nq_pad + 4), cudaMemcpyDeviceToHost, st));
    RS_CUDA(cudaStreamSynchronize(st));
    if (flags[nq_pad + 1] != 0) {
      // a staging queue of the filter overran (scores packed inside the margin): nothing is trusted
      search_exact_device(h, qc, cq, k, id_base, D + q0 * ldD, ldD, I + q0 * ldI, ldI, st);
      flags[nq_pad] = cq;
    } else if (flags[nq_pad] > 0) {
      for (int i = 0; i < cq; ++i)
        if (flags[i])
          search_exact_device(h, qc + (int64_t)i * DK, 1, k, id_base, D + (q0 + i) * ldD, ldD, I + (q0 + i) * ldI, ldI, st);
    }
    if (q0 == 0) h->last_fallbacks.store(flags[nq_pad], std::memory_order_relaxed);
    else h->last_fallbacks.fetch_add(flags[nq_pad], std::memory_order_relaxed);
    h->total_fallbacks.fetch_add(flags[nq_pad], std::memory_order_relaxed);
  }
}

// ---- shared-threshold search of one shard of a partitioned catalog (see rs_index_shared_begin in the header) ----
static PeerOut peer_out(rs_peer* region, int64_t off_I = 0, int64_t off_F = 0, int64_t off_B = 0) {
  PeerOut o;
  if (region) {
    RS_CHECK(region->connected, "peer region is not connected");
    o.bases = region->d_bases;
    o.PL = rsp::layout_of(region);
    o.off_I = off_I;
    o.off_F = off_F;
    o.off_B = off_B;
  }
  return o;
}

rs_tc_session* shared_begin(rs_index* h, const float* q, int64_t nq, int k, int R, float* top, int64_t ld_top, cudaStream_t st,
                            rs_peer* publish_tops) {
  RS_CHECK(nq <= tensor_max_queries(), "too many queries for one shared-threshold search");
  RS_CHECK(h->ntotal >= 65536, "shared-threshold searches need shards of at least 65,536 rows");
  auto sp = tc_begin(h, q, static_cast<int>(nq), k, st);
  RS_CHECK(R >= 1 && R <= sp->n_vals, "R must be in [1, number of sample maxima]");
  RS_CHECK(!publish_tops || nq * R * 4 <= publish_tops->slot_bytes, "tops message longer than the region's slot");
  ProfScope prof(RS_PROF_TOPS, st);
  sample_top_kernel<<<static_cast<unsigned>(nq), 256, 0, st>>>(sp->ws.smax, sp->n_vals, R, top, ld_top, peer_out(publish_tops));
  RS_LAUNCH_CHECK();
  return sp.release();
}

void shared_finish(rs_tc_session* sp, const float* tops, int n_parts, int64_t part_ld, int R, int rank, int64_t id_base,
                   float* D, int64_t ldD, int64_t* I, int64_t ldI, int* need_exact, float* bound, rs_peer* tops_region,
                   rs_peer* publish_lists, int64_t off_I, int64_t off_F, int64_t off_B) {
  std::unique_ptr<rs_tc_session> own(sp);
  rs_tc_session& s = *sp;
  RS_CHECK(n_parts >= 1 && R >= 1 && n_parts * R <= 1024, "shards x R must be <= 1024");
  RS_CHECK(rank >= 1 && rank <= n_parts * R, "rank must be in [1, shards x R]");
  {
    ProfScope prof(RS_PROF_THRESHOLD, s.st);
    threshold_shared_kernel<<<s.nq, 256, 0, s.st>>>(tops, n_parts, part_ld, R, rank, s.ws.thr, s.ws.cnt, s.ws.flags,
                                                    tops_region ? tops_region->d_bases : nullptr,
                                                    tops_region ? rsp::layout_of(tops_region) : rsp::PeerLayout{});
    RS_LAUNCH_CHECK();
  }
  RS_CHECK(!publish_lists || off_B + (int64_t)s.nq * 4 <= publish_lists->slot_bytes, "lists message longer than the region's slot");
  RS_CHECK(publish_lists || bound != nullptr, "bound is NULL");
  // the union's rank-th sample maximum keeps ~rank / rho candidates in the whole catalog, 1 / n_parts of them here
  const double rho_own = double(s.n_sample * BN) / double(s.h->ntotal);
  tc_finish(s, id_base, D, ldD, I, ldI, bound, need_exact, double(rank) / rho_own / n_parts,
            peer_out(publish_lists, off_I, off_F, off_B));
}

void tc_session_free(rs_tc_session* sp) { delete sp; }

void tensor_scores_debug(rs_index* h, const float* q, int64_t nq, float* out, cudaStream_t st) {
  RS_CHECK(h->d == DK && h->ntotal >= 1, "needs d = 128 and a non-empty index");
  ensure_shadow(h, st);
  const int sms = sm_count();
  const int nq_pad = static_cast<int>((nq + BM - 1) / BM * BM);
  RS_CHECK(nq_pad / BM <= sms * R_MAX, "too many queries for the debug entry");
  Workspace ws(st);
  RS_CUDA(cudaMallocAsync(&ws.qb, sizeof(uint16_t) * (size_t)nq_pad * DK, st));
  RS_CUDA(cudaMallocAsync(&ws.eps2, sizeof(float) * nq_pad, st));
  prep_queries_kernel<<<(nq_pad * 32 + 127) / 128, 128, 0, st>>>(q, (int)nq, nq_pad, ws.qb, ws.eps2,
                                                                reinterpret_cast<unsigned int*>(h->row_norm));
  RS_LAUNCH_CHECK();
  const CUtensorMap tq = make_tmap_2d(ws.qb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)nq_pad, DK, (uint64_t)DK * 2,
                                      BM, 64);
  TcParams p{};
  p.n_items = h->ntotal;
  p.nq = (int)nq;
  p.n_qtiles = nq_pad / BM;
  p.n_groups = (p.n_qtiles + R_MAX - 1) / R_MAX;
  p.ctas_per_group = std::max(1, sms / p.n_groups);
  p.n_visit = (h->ntotal + BN - 1) / BN;
  p.tile_step = 1;
  p.dump = out;
  p.dump_ld = h->ntotal;
  launch_filter<MODE_DUMP>(tq, h->tmap_bf16, p, h->device, st);
}

}  // namespace rs

// ------------------------------------------------------------------------------------------------
// C ABI of the shared-threshold search (declared in include/recsys_b200.h)
// ------------------------------------------------------------------------------------------------
extern "C" {

int rs_tensor_plan(int d, int64_t ntotal, int64_t nq, int k, int64_t* sampled_rows, int* supported) {
  return rs::guard([&] {
    RS_CHECK(sampled_rows && supported, "output pointer is NULL");
    *supported = (d == DK && k >= 1 && k <= 128 && ntotal >= 65536 && ntotal <= 0x7FFFFF00ll && nq >= 1 &&
                  nq <= rs::tensor_max_queries()) ? 1 : 0;
    *sampled_rows = ntotal > 0 ? rs::tensor_sample_rows(ntotal) : 0;
  });
}

int rs_tensor_threshold_rank(int k, double rho, int* rank) {
  return rs::guard([&] {
    RS_CHECK(rank != nullptr && k >= 1 && rho > 0.0 && rho <= 1.0, "rank is NULL, k < 1 or rho outside (0, 1]");
    *rank = rs::tensor_threshold_rank(k, rho);
  });
}

int rs_index_shared_begin(rs_index_t* h, const DLTensor* q, int k, DLTensor* top, rs_tc_session_t** out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr && out != nullptr, "index or out is NULL");
    require(q, "q", 2, kDLFloat, 32);
    require(top, "top", 2, kDLFloat, 32);
    const int64_t nq = q->shape[0];
    RS_CHECK(q->shape[1] == h->d && nq >= 1, "q must be [nq >= 1, d] with d == index dimension");
    RS_CHECK(on_device(q) && on_device(top) && is_compact(q) && q->device.device_id == h->device, "q / top must be CUDA tensors on the index's device, q contiguous");
    RS_CHECK(top->shape[0] == nq && top->shape[1] >= 1 && (top->shape[1] == 1 || stride_of(top, 1) == 1), "top must be [nq, R] with unit column stride");
    require_aligned16(data_ptr<float>(q), "q");
    DeviceGuard g(h->device);
    *out = shared_begin(h, data_ptr<float>(q), nq, k, static_cast<int>(top->shape[1]), data_ptr<float>(top),
                        row_stride_2d(top, "top"), static_cast<cudaStream_t>(stream));
  });
}

int rs_index_shared_finish(rs_tc_session_t* s, const DLTensor* tops, int rank, int64_t id_base, DLTensor* D, DLTensor* I,
                           DLTensor* need_exact, DLTensor* bound, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(s != nullptr, "session is NULL");
    std::unique_ptr<rs_tc_session> own(s);   // freed on every path
    RS_CHECK(static_cast<cudaStream_t>(stream) == s->st, "finish must use the stream begin was called with");
    require(tops, "tops", 3, kDLFloat, 32);
    require(D, "D", 2, kDLFloat, 32);
    require(I, "I", 2, kDLInt, 64);
    require(need_exact, "need_exact", 1, kDLInt, 32);
    require(bound, "bound", 1, kDLFloat, 32);
    const int64_t nq = s->nq, k = s->k, G = tops->shape[0], R = tops->shape[2];
    RS_CHECK(tops->shape[1] == nq && on_device(tops), "tops must be a CUDA [shards, nq, R] tensor");
    int64_t part_ld = nq * R;
    if (tops->strides) {
      RS_CHECK((R <= 1 || tops->strides[2] == 1) && (nq <= 1 || tops->strides[1] == R), "tops: each shard's [nq, R] block must be contiguous");
      if (G > 1) part_ld = tops->strides[0];
    }
    RS_CHECK(D->shape[0] == nq && D->shape[1] == k && I->shape[0] == nq && I->shape[1] == k, "D and I must be [nq, k]");
    RS_CHECK(need_exact->shape[0] == nq && bound->shape[0] == nq && (nq <= 1 || (stride_of(need_exact, 0) == 1 && stride_of(bound, 0) == 1)),
             "need_exact / bound must be contiguous [nq] tensors");
    RS_CHECK(on_device(D) && on_device(I) && on_device(need_exact) && on_device(bound), "outputs must be CUDA tensors");
    DeviceGuard g(s->h->device);
    shared_finish(own.release(), data_ptr<float>(tops), static_cast<int>(G), part_ld, static_cast<int>(R), rank, id_base,
                  data_ptr<float>(D), row_stride_2d(D, "D"), data_ptr<int64_t>(I), row_stride_2d(I, "I"),
                  data_ptr<int>(need_exact), data_ptr<float>(bound));
  });
}

int rs_tc_session_free(rs_tc_session_t* s) {
  return rs::guard([&] { rs::tc_session_free(s); });
}

}  // extern "C"
