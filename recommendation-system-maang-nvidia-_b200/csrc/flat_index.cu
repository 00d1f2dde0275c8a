// flat_index.cu -- K4/K5/K6/K7 (exact fp32 mode): the FAISS IndexFlatIP replacement.
//
// Replaces faiss.IndexFlatIP.add/search (src/trainer.py:241-242, app/recommendation_service.py:72),
// faiss.normalize_L2 (src/trainer.py:240, app/recommendation_service.py:71) and
// tf.linalg.matvec (app/recommendation_service.py:91).
//
// scan_topk_kernel -- the catalog scan.  HBM-bound: the fp32 catalog ([N, d] row-major, the layout
// faiss keeps) is streamed once per group of QT queries through a 3-stage shared-memory ring filled by
// TMA (cp.async.bulk.tensor, 128-byte swizzle, one mbarrier per stage).  One thread scores one catalog
// row of the tile against the QT resident queries with the stated arithmetic -- a single fp32
// accumulator per (query, row), fused multiply-add, feature index ascending -- which the oracle
// (oracle/chain.c) reproduces bit for bit.  The swizzle makes the row-per-thread 128-bit shared loads
// bank-conflict free.  Scores never go to HBM: each is compared in registers with the CTA's running
// k-th key and only the rare survivors enter the shared-memory candidate buffer (topk.cuh).
// Algorithmic bytes: N*d*4 per query group (+ nq*d*4 queries, + nq*P*k*8 partial lists).
//
// merge_topk_kernel -- one CTA per query merges the P per-CTA lists (or the G per-shard lists after
// the all-gather) in rank-major order with early termination.
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <mutex>

#include "flat_index.cuh"
#include "peer.cuh"
#include "topk.cuh"

namespace {

using namespace rsd;

constexpr int TILE_ROWS = 128;  // catalog rows per pipeline stage == compute threads
constexpr int SCAN_NT = 128;
constexpr int STAGES = 3;
constexpr int MERGE_NT = 256;

template <int D, int QT, int CAP>
struct ScanSmem {
  static constexpr int kStageBytes = TILE_ROWS * D * 4;
  static constexpr int kKeysOff = STAGES * kStageBytes;
  static constexpr int kQOff = kKeysOff + QT * CAP * 8;
  static constexpr int kBarOff = kQOff + QT * D * 4;
  static constexpr int kCntOff = kBarOff + STAGES * 8;
  static constexpr int kTotal = kCntOff + QT * 4;
  static constexpr int kDynamic = kTotal + 1024;  // slack for the 1024-byte alignment of the ring
};

template <int D, int QT, int CAP>
__global__ void __launch_bounds__(SCAN_NT, 1)
scan_topk_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ q, int64_t n, int nq, int k,
                 uint64_t* __restrict__ partial) {
  using L = ScanSmem<D, QT, CAP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* keys = reinterpret_cast<uint64_t*>(base + L::kKeysOff);
  float* qs = reinterpret_cast<float*>(base + L::kQOff);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + L::kBarOff);
  int* cnt = reinterpret_cast<int*>(base + L::kCntOff);

  const int tid = threadIdx.x;
  const int qbase = blockIdx.y * QT;
  const int64_t ntiles = (n + TILE_ROWS - 1) / TILE_ROWS;
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (tid == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < QT * D; i += SCAN_NT) {
    const int g = i / D;
    qs[i] = (qbase + g < nq) ? q[(int64_t)(qbase + g) * D + (i - g * D)] : 0.0f;
  }
  for (int i = tid; i < QT * CAP; i += SCAN_NT) keys[i] = kEmptyKey;
  if (tid < QT) cnt[tid] = 0;
  __syncthreads();

  auto issue = [&](int stage, int64_t it) {
    const int64_t tile = blockIdx.x + it * gridDim.x;
    mbar_arrive_expect_tx(&full[stage], L::kStageBytes);
#pragma unroll
    for (int j = 0; j < D / 32; ++j)
      tma_load_2d(base + stage * L::kStageBytes + j * (TILE_ROWS * 128), &tmap, &full[stage], j * 32,
                  static_cast<int>(tile * TILE_ROWS), kEvictFirst);
  };
  if (tid == 0)
    for (int s = 0; s < STAGES && s < my_tiles; ++s) issue(s, s);

  uint64_t thr[QT];
#pragma unroll
  for (int g = 0; g < QT; ++g) thr[g] = kEmptyKey;

  const uint32_t sw = static_cast<uint32_t>(tid & 7) << 4;
  for (int64_t it = 0; it < my_tiles; ++it) {
    const int stage = static_cast<int>(it % STAGES);
    mbar_wait(&full[stage], static_cast<uint32_t>((it / STAGES) & 1));

    // ---- score: row `tid` of the tile against QT queries, one FMA chain each ----
    float acc[QT];
#pragma unroll
    for (int g = 0; g < QT; ++g) acc[g] = 0.0f;
    const uint8_t* srow = base + stage * L::kStageBytes + tid * 128;
#pragma unroll
    for (int j = 0; j < D / 32; ++j) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 xv = *reinterpret_cast<const float4*>(srow + j * (TILE_ROWS * 128) + ((c << 4) ^ sw));
#pragma unroll
        for (int g = 0; g < QT; ++g) {
          const float4 qv = *reinterpret_cast<const float4*>(qs + g * D + j * 32 + c * 4);
          acc[g] = fmaf(qv.x, xv.x, acc[g]);
          acc[g] = fmaf(qv.y, xv.y, acc[g]);
          acc[g] = fmaf(qv.z, xv.z, acc[g]);
          acc[g] = fmaf(qv.w, xv.w, acc[g]);
        }
      }
    }

    // ---- select: threshold-reject in registers, survivors to the candidate buffers ----
    const int64_t row = (blockIdx.x + it * gridDim.x) * TILE_ROWS + tid;
    uint64_t key[QT];
    uint32_t pend = 0;
#pragma unroll
    for (int g = 0; g < QT; ++g) {
      key[g] = pack_key(acc[g], static_cast<uint32_t>(row));
      if (row < n && qbase + g < nq && key[g] > thr[g]) pend |= 1u << g;
    }
    for (;;) {
      bool saw_full = false;
#pragma unroll
      for (int g = 0; g < QT; ++g) {
        if (pend & (1u << g)) {
          const int pos = atomicAdd(&cnt[g], 1);
          if (pos < CAP) {
            keys[g * CAP + pos] = key[g];
            pend &= ~(1u << g);
          }
          saw_full |= pos >= CAP - 1;
        }
      }
      // this barrier also orders every thread's reads of the stage before the refill below
      if (!__syncthreads_or(pend != 0 || saw_full)) break;
#pragma unroll
      for (int g = 0; g < QT; ++g) {
        if (cnt[g] >= CAP) {  // uniform: nobody touches cnt between the barrier and the prune
          thr[g] = prune_buffer<CAP, SCAN_NT>(keys + g * CAP, &cnt[g], k, tid);
          if ((pend & (1u << g)) && !(key[g] > thr[g])) pend &= ~(1u << g);
        }
      }
      __syncthreads();
    }
    if (tid == 0 && it + STAGES < my_tiles) issue(stage, it + STAGES);
  }

  // ---- emit this CTA's sorted list for each query ----
  const int P = gridDim.x;
#pragma unroll 1
  for (int g = 0; g < QT; ++g) {
    if (qbase + g >= nq) break;
    prune_buffer<CAP, SCAN_NT>(keys + g * CAP, &cnt[g], k, tid);
    uint64_t* dst = partial + ((int64_t)(qbase + g) * P + blockIdx.x) * k;
    for (int j = tid; j < k; j += SCAN_NT) dst[j] = keys[g * CAP + j];
  }
}

// ------------------------------------------------------------------------------------------------
// scan_bf16_topk_kernel -- the same scan for ONE or TWO queries (the /recommendations path, BASELINE configs[4]) at half
// the bytes: the catalog is read from the bf16 shadow (what the tensor path keeps anyway); a CTA keeps every row whose
// low-precision score is within 2 eps of its running k-th low-precision score (eps = c |q| max|x| bounds |s_hat - s|, as
// in tc_filter.cu), and when its part of the catalog is done re-scores those few rows from the fp32 catalog with the
// stated FMA chain and emits its exact top-k.  Every row of a CTA's true top-k has s_hat >= s_hat_(k) - 2 eps (the tensor
// path's argument), so the emitted lists -- and after the merge the results -- are bit-identical to scan_topk_kernel's.
// Two launches per search (scan, merge) against the tensor path's seven: at batch 1 the launches, not the bytes, were
// most of the step on a shard of an 8-way split catalog.  A first version re-scored candidates inline: ~80 dependent
// global loads per CTA, each stalling the whole CTA at the per-tile barrier, made it slower than the fp32 scan.
// A CTA whose margin holds more rows than its buffer (scores packed inside 2 eps: degenerate catalogs) raises ovf[query];
// the caller then repeats that query on the fp32 scan.
// ------------------------------------------------------------------------------------------------
constexpr int BSCAN_STAGES = 3;
constexpr int BSCAN_ROWS = 256;                          // rows per stage
constexpr int BSCAN_NT = 512;                            // two threads per row (one per 64-feature half): 16 warps per SM -- with
                                                         // 8 the kernel was latency-bound at 4.7 TB/s (ncu: 24 % issue utilisation)
constexpr int BSCAN_STAGE_BYTES = BSCAN_ROWS * 128 * 2;  // [256 rows x 128 bf16] as two 128-byte-swizzled [256 x 64] boxes
constexpr int BSCAN_CAP = 1024;                          // candidate buffer per query
constexpr int BSCAN_SOFT = 512;                          // prune as soon as this many are held: the first prune (after two
                                                         // tiles) sets the cut, and the sorts stay short; a prune must leave
                                                         // room for one tile below it, and two tiles fit above it

template <int QT>
struct BScanSmem {
  static constexpr int kKeysOff = BSCAN_STAGES * BSCAN_STAGE_BYTES;
  static constexpr int kQOff = kKeysOff + QT * BSCAN_CAP * 8;
  static constexpr int kBarOff = kQOff + QT * 128 * 4;
  static constexpr int kCntOff = kBarOff + BSCAN_STAGES * 8;
  static constexpr int kEpsOff = kCntOff + QT * 4;
  static constexpr int kPartOff = kEpsOff + QT * 4 + 16;
  static constexpr int kTotal = kPartOff + QT * BSCAN_ROWS * 4;
  static constexpr int kDynamic = kTotal + 1024;
};

// Collective.  Sorts the buffer (descending), keeps every entry whose score is within eps2 of the k-th (all entries while
// fewer than k are held), returns the new k-th key (kEmptyKey while fewer than k).  *cnt = entries kept; more than
// CAP - 2 tiles' worth: *ovf = 1 and only the best k are kept (the list is no longer trusted).
__device__ __forceinline__ uint64_t prune_margin(uint64_t* keys, int* cnt, int* kept, int k, float eps2, int* ovf, int tid) {
  const int n_valid = min(*cnt, BSCAN_CAP);
  for (int i = n_valid + tid; i < BSCAN_CAP; i += BSCAN_NT) keys[i] = kEmptyKey;
  if (tid == 0) *kept = 0;
  __syncthreads();
  block_sort_desc_upto<BSCAN_CAP, BSCAN_NT>(keys, n_valid, tid);
  uint64_t thr = kEmptyKey;
  int m = n_valid;
  if (n_valid >= k) {   // uniform
    thr = keys[k - 1];
    const float cutv = key_score(thr) - eps2;
    int mine = 0;
    for (int i = tid; i < n_valid; i += BSCAN_NT) mine += key_score(keys[i]) >= cutv ? 1 : 0;
    if (mine) atomicAdd(kept, mine);
    __syncthreads();
    m = *kept;
    if (m > BSCAN_SOFT - BSCAN_ROWS) {
      if (tid == 0) *ovf = 1;
      m = k;
    }
  }
  __syncthreads();
  if (tid == 0) *cnt = m;
  __syncthreads();
  return thr;
}

template <int QT>
__global__ void __launch_bounds__(BSCAN_NT, 1)
scan_bf16_topk_kernel(const __grid_constant__ CUtensorMap tmap16, const float* __restrict__ xb, const float* __restrict__ q,
                      int64_t n, int nq, int k, float eps_scale, uint64_t* __restrict__ partial, int* __restrict__ ovf) {
  using L = BScanSmem<QT>;
  constexpr int D = 128;
  constexpr int CAP = BSCAN_CAP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* keys = reinterpret_cast<uint64_t*>(base + L::kKeysOff);
  float* qs = reinterpret_cast<float*>(base + L::kQOff);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + L::kBarOff);
  int* cnt = reinterpret_cast<int*>(base + L::kCntOff);
  float* eps_s = reinterpret_cast<float*>(base + L::kEpsOff);   // [QT] 2 eps
  int* kept = reinterpret_cast<int*>(eps_s + QT);               // scratch of prune_margin
  float* part = reinterpret_cast<float*>(base + L::kPartOff);   // [QT][256]: the second half's partial sums

  const int tid = threadIdx.x;
  const int trow = tid & (BSCAN_ROWS - 1);                       // this thread's row of the tile ...
  const int half = tid >> 8;                                     // ... and 64-feature half (warps 0-7 / 8-15: no bank conflicts)
  const int64_t ntiles = (n + BSCAN_ROWS - 1) / BSCAN_ROWS;
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (tid == 0) {
    tma_prefetch_desc(&tmap16);
    for (int s = 0; s < BSCAN_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < QT * D; i += BSCAN_NT) {
    const int g = i / D;
    qs[i] = (g < nq) ? q[(int64_t)g * D + (i - g * D)] : 0.0f;
  }
  for (int i = tid; i < QT * CAP; i += BSCAN_NT) keys[i] = kEmptyKey;
  if (tid < QT) cnt[tid] = 0;
  __syncthreads();
  if (tid < 32) {   // 2 eps_g = 2 eps_scale |q_g| (a hair above: the norm is rounded up)
    for (int g = 0; g < QT; ++g) {
      float s = 0.0f;
      for (int i = tid; i < D; i += 32) s = fmaf(qs[g * D + i], qs[g * D + i], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      if (tid == 0) eps_s[g] = 2.0f * eps_scale * (sqrtf(s) * 1.00001f);
    }
  }
  __syncthreads();

  auto issue = [&](int stage, int64_t it) {
    const int64_t tile = blockIdx.x + it * gridDim.x;
    mbar_arrive_expect_tx(&full[stage], BSCAN_STAGE_BYTES);
#pragma unroll
    for (int h = 0; h < 2; ++h)
      tma_load_2d(base + stage * BSCAN_STAGE_BYTES + h * (BSCAN_ROWS * 128), &tmap16, &full[stage], h * 64,
                  static_cast<int>(tile * BSCAN_ROWS), kEvictFirst);
  };
  if (tid == 0)
    for (int s = 0; s < BSCAN_STAGES && s < my_tiles; ++s) issue(s, s);

  float cut[QT];        // keep rows with s_hat >= cut = (running k-th s_hat) - 2 eps (-inf until k rows are held)
  float eps2[QT];
#pragma unroll
  for (int g = 0; g < QT; ++g) {
    cut[g] = -INFINITY;
    eps2[g] = eps_s[g];
  }

  for (int64_t it = 0; it < my_tiles; ++it) {
    const int stage = static_cast<int>(it % BSCAN_STAGES);
    mbar_wait(&full[stage], static_cast<uint32_t>((it / BSCAN_STAGES) & 1));

    // ---- low-precision score of row `tid`: bf16 row x fp32 query, fp32 accumulation ----
    // four partial sums per query over this thread's half of the row: the order of a low-precision score does not matter,
    // the latency of one long chain does
    float a4[QT][4];
#pragma unroll
    for (int g = 0; g < QT; ++g)
#pragma unroll
      for (int u = 0; u < 4; ++u) a4[g][u] = 0.0f;
    const uint8_t* srow = base + stage * BSCAN_STAGE_BYTES + half * (BSCAN_ROWS * 128) + trow * 128;
    const uint32_t sw = static_cast<uint32_t>(trow & 7) << 4;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 xv = *reinterpret_cast<const uint4*>(srow + ((c << 4) ^ sw));
      const uint32_t w[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int g = 0; g < QT; ++g) {
        const float4 qa = *reinterpret_cast<const float4*>(qs + g * D + half * 64 + c * 8);
        const float4 qb = *reinterpret_cast<const float4*>(qs + g * D + half * 64 + c * 8 + 4);
        a4[g][0] = fmaf(qa.x, __uint_as_float(w[0] << 16), a4[g][0]);
        a4[g][1] = fmaf(qa.y, __uint_as_float(w[0] & 0xFFFF0000u), a4[g][1]);
        a4[g][2] = fmaf(qa.z, __uint_as_float(w[1] << 16), a4[g][2]);
        a4[g][3] = fmaf(qa.w, __uint_as_float(w[1] & 0xFFFF0000u), a4[g][3]);
        a4[g][0] = fmaf(qb.x, __uint_as_float(w[2] << 16), a4[g][0]);
        a4[g][1] = fmaf(qb.y, __uint_as_float(w[2] & 0xFFFF0000u), a4[g][1]);
        a4[g][2] = fmaf(qb.z, __uint_as_float(w[3] << 16), a4[g][2]);
        a4[g][3] = fmaf(qb.w, __uint_as_float(w[3] & 0xFFFF0000u), a4[g][3]);
      }
    }
    float acc[QT];
#pragma unroll
    for (int g = 0; g < QT; ++g) {
      acc[g] = (a4[g][0] + a4[g][1]) + (a4[g][2] + a4[g][3]);
      if (half == 1) part[g * BSCAN_ROWS + trow] = acc[g];
    }
    __syncthreads();
    if (half == 0) {
#pragma unroll
      for (int g = 0; g < QT; ++g) acc[g] += part[g * BSCAN_ROWS + trow];
    }

    // ---- rows inside the margin go to the candidate buffer with their low-precision key ----
    const int64_t row = (blockIdx.x + it * gridDim.x) * BSCAN_ROWS + trow;
    uint32_t pend = 0;
#pragma unroll
    for (int g = 0; g < QT; ++g)
      if (half == 0 && row < n && g < nq && acc[g] >= cut[g]) pend |= 1u << g;
    for (;;) {
      bool saw_full = false;
#pragma unroll
      for (int g = 0; g < QT; ++g) {
        if (pend & (1u << g)) {
          const int pos = atomicAdd(&cnt[g], 1);
          if (pos < CAP) {
            keys[g * CAP + pos] = pack_key(acc[g], static_cast<uint32_t>(row));
            pend &= ~(1u << g);
          }
          saw_full |= pos >= BSCAN_SOFT - 1;
        }
      }
      // this barrier also orders every thread's reads of the stage before the refill below
      if (!__syncthreads_or(pend != 0 || saw_full)) break;
#pragma unroll
      for (int g = 0; g < QT; ++g) {
        if (cnt[g] >= BSCAN_SOFT) {  // uniform: nobody touches cnt between the barrier and the prune
          const uint64_t thr = prune_margin(keys + g * CAP, &cnt[g], kept, k, eps2[g], ovf + g, tid);
          if (thr != kEmptyKey) cut[g] = key_score(thr) - eps2[g];
          if ((pend & (1u << g)) && !(acc[g] >= cut[g])) pend &= ~(1u << g);
        }
      }
      __syncthreads();
    }
    if (tid == 0 && it + BSCAN_STAGES < my_tiles) issue(stage, it + BSCAN_STAGES);
  }

  // ---- exact scores of the kept rows (fp32 catalog, the stated chain: one accumulator, k ascending), exact order, emit ----
  const int P = gridDim.x;
#pragma unroll 1
  for (int g = 0; g < QT; ++g) {
    if (g >= nq) break;
    uint64_t* kg = keys + g * CAP;
    prune_margin(kg, &cnt[g], kept, k, eps2[g], ovf + g, tid);
    const int m = cnt[g];
    for (int i = tid; i < CAP; i += BSCAN_NT) {
      uint64_t key = kEmptyKey;
      if (i < m) {
        const uint32_t r = key_row(kg[i]);
        const float4* xr = reinterpret_cast<const float4*>(xb + static_cast<int64_t>(r) * D);
        float ex = 0.0f;
#pragma unroll 16
        for (int c4 = 0; c4 < D / 4; ++c4) {
          const float4 xv = __ldg(xr + c4);
          const float4 qv = *reinterpret_cast<const float4*>(qs + g * D + c4 * 4);
          ex = fmaf(qv.x, xv.x, ex);
          ex = fmaf(qv.y, xv.y, ex);
          ex = fmaf(qv.z, xv.z, ex);
          ex = fmaf(qv.w, xv.w, ex);
        }
        key = pack_key(ex, r);
      }
      kg[i] = key;
    }
    __syncthreads();
    block_sort_desc_upto<CAP, BSCAN_NT>(kg, m, tid);
    uint64_t* dst = partial + ((int64_t)g * P + blockIdx.x) * k;
    for (int j = tid; j < k; j += BSCAN_NT) dst[j] = kg[j];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// merge of sorted lists.  Source A: packed keys [nq][P][k] (rows local to the index).
// Source B: (score, id) pairs [P][nq][k] (global ids, -1 = padding) after an all-gather.
// ------------------------------------------------------------------------------------------------
// after the merge of a shared-threshold search: a query needs the exact re-search iff a shard overflowed, the merged
// list is short of k, or its k-th score does not clear every shard's bound (tau + eps)
__global__ void verify_topk_kernel(const float* __restrict__ D, int64_t ldD, const int64_t* __restrict__ I, int64_t ldI, int k,
                                   const float* __restrict__ bounds, int64_t ld_b, const int* __restrict__ flags, int64_t ld_f,
                                   int G, int64_t nq, int* __restrict__ need_exact) {
  const int64_t qi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  float b = -INFINITY;
  int f = 0;
  for (int g = 0; g < G; ++g) {
    b = fmaxf(b, bounds[g * ld_b + qi]);
    f |= flags[g * ld_f + qi];
  }
  const bool ok = f == 0 && I[qi * ldI + (k - 1)] >= 0 && D[qi * ldD + (k - 1)] >= b;
  need_exact[qi] = ok ? 0 : 1;
}

struct KeySource {
  const uint64_t* keys;  // A
  const float* Dp;       // B
  const int64_t* Ip;     // B
  int64_t nq;
  int64_t part_ld_D, part_ld_I;   // B: elements between consecutive parts (nq * k when the parts are packed)
  __device__ __forceinline__ uint64_t load(int64_t qi, int p, int P, int r, int k) const {
    if (keys) return keys[((int64_t)qi * P + p) * k + r];
    const int64_t o = qi * k + r;
    const int64_t id = Ip[p * part_ld_I + o];
    return id < 0 ? kEmptyKey : pack_key(Dp[p * part_ld_D + o], static_cast<uint32_t>(id));
  }
};

// What the sharded search (sharded.cu) folds into the merge: the wait for / acknowledgement of the peer-memory region the
// parts live in, and the per-query verdict that rs_topk_verify computes (OR of the shards' flags; with bounds also the
// shared-threshold verification).  All pointers null: a plain merge.
struct MergeExtras {
  uint8_t* const* peer_bases = nullptr;
  rsp::PeerLayout PL{};
  const float* bounds = nullptr;
  int64_t ld_b = 0;
  const int* flags = nullptr;
  int64_t ld_f = 0;
  int* need_exact = nullptr;
};

template <int CAP>
__global__ void __launch_bounds__(MERGE_NT)
merge_topk_kernel(const KeySource src, int P, int k, int64_t id_base, float* __restrict__ D, int64_t ldD,
                  int64_t* __restrict__ I, int64_t ldI, const MergeExtras ex) {
  __shared__ uint64_t keys[CAP];
  __shared__ int cnt;
  const int tid = threadIdx.x;
  const int64_t qi = blockIdx.x;
  uint64_t seq = 0;
  if (ex.peer_bases) seq = rsp::peer_arrived(ex.peer_bases[ex.PL.rank], ex.PL);
  for (int i = tid; i < CAP; i += MERGE_NT) keys[i] = kEmptyKey;
  if (tid == 0) cnt = 0;
  __syncthreads();
  uint64_t thr = kEmptyKey;
  const int64_t total = (int64_t)P * k;
  int quiet = 0;  // consecutive elements (rank-major) in which no list offered a key
  for (int64_t e0 = 0; e0 < total; e0 += MERGE_NT) {
    const int64_t e = e0 + tid;
    uint64_t key = kEmptyKey;
    if (e < total) {
      const int r = static_cast<int>(e / P);
      const int p = static_cast<int>(e - (int64_t)r * P);
      key = src.load(qi, p, P, r, k);
    }
    const bool any = offer_one<CAP, MERGE_NT>(keys, &cnt, k, key != kEmptyKey, key, thr, tid);
    quiet = any ? 0 : quiet + MERGE_NT;
    if (quiet >= P) break;  // every list's next entry already lost; lists are sorted, so we are done
  }
  prune_buffer<CAP, MERGE_NT>(keys, &cnt, k, tid);
  for (int j = tid; j < k; j += MERGE_NT) {
    const uint64_t key = keys[j];
    if (key == kEmptyKey) {
      D[qi * ldD + j] = -FLT_MAX;
      I[qi * ldI + j] = -1;
    } else {
      D[qi * ldD + j] = key_score(key);
      I[qi * ldI + j] = id_base + static_cast<int64_t>(key_row(key));
    }
  }
  if (ex.need_exact && tid == 0) {   // verify_topk_kernel's verdict for this query
    float b = -INFINITY;
    int f = 0;
    for (int g = 0; g < P; ++g) {
      if (ex.bounds) b = fmaxf(b, ex.bounds[g * ex.ld_b + qi]);
      f |= ex.flags[g * ex.ld_f + qi];
    }
    const uint64_t kth = keys[k - 1];
    const bool ok = f == 0 && (!ex.bounds || (kth != kEmptyKey && key_score(kth) >= b));
    ex.need_exact[qi] = ok ? 0 : 1;
  }
  if (ex.peer_bases) rsp::peer_consumed(ex.peer_bases, ex.PL, seq);
}

__global__ void fill_pad_kernel(float* D, int64_t ldD, int64_t* I, int64_t ldI, int64_t nq, int k) {
  const int64_t total = nq * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k;
    const int j = static_cast<int>(i - r * k);
    D[r * ldD + j] = -FLT_MAX;
    I[r * ldI + j] = -1;
  }
}

// K4: one thread per row, the stated chain over x_j^2 then x_j *= 1/sqrt(nr); zero rows untouched
__global__ void __launch_bounds__(128) normalize_l2_kernel(float* __restrict__ x, int64_t ld, int64_t n, int d) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  float* row = x + r * ld;
  float nr = 0.0f;
  for (int j = 0; j < d; ++j) nr = fmaf(row[j], row[j], nr);
  if (nr > 0.0f) {
    const float inv = 1.0f / sqrtf(nr);
    for (int j = 0; j < d; ++j) row[j] *= inv;
  }
}

// K6: one thread per item row, the stated chain (n <= 100 on the /score path)
__global__ void __launch_bounds__(128) matvec_kernel(const float* __restrict__ items, int64_t ld,
                                                     const float* __restrict__ user, float* __restrict__ out,
                                                     int64_t n, int d) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float* row = items + r * ld;
  float acc = 0.0f;
  for (int j = 0; j < d; ++j) acc = fmaf(user[j], row[j], acc);
  out[r] = acc;
}

// The same merge for short inputs (P * k <= 4096 keys, k <= 128: the batch-1 path, 148 lists of 10, or 8 shards' lists of
// 100) without a sort: the lists are sorted, so the element at position r - 1 of a list has r elements at or above it.  With
// r = ceil(k / P) and m = ceil(k / r), tau = the m-th largest of those elements has m r >= k elements at or above it, hence
// the k-th largest key overall is >= tau and only keys >= tau can be in the result: a handful, ranked by counting.
// merge_topk_kernel's rounds of offer / prune cost ~13 us here -- a fifth of the whole catalog scan of a shard; a bitonic
// sort of all keys 25 us; a radix select 11 us (a thousand shared-memory atomics on two or three histogram bins).
constexpr int SMALL_MERGE_KEYS = 4096;
constexpr int SMALL_MERGE_TOP = 512;
__global__ void __launch_bounds__(MERGE_NT)
merge_small_kernel(const KeySource src, int P, int k, int64_t id_base, float* __restrict__ D, int64_t ldD,
                   int64_t* __restrict__ I, int64_t ldI, const MergeExtras ex) {
  __shared__ uint64_t all[SMALL_MERGE_KEYS];
  __shared__ uint64_t top[SMALL_MERGE_TOP];
  __shared__ uint64_t tau_s, kth_s;
  __shared__ int n_top;
  const int tid = threadIdx.x;
  const int64_t qi = blockIdx.x;
  uint64_t seq = 0;
  if (ex.peer_bases) seq = rsp::peer_arrived(ex.peer_bases[ex.PL.rank], ex.PL);
  const int total = P * k;
  if (tid == 0) {
    n_top = 0;
    tau_s = kEmptyKey;
    kth_s = kEmptyKey;
  }
  for (int e = tid; e < total; e += MERGE_NT) {
    const int p = e / k, r = e - p * k;
    all[e] = src.load(qi, p, P, r, k);
  }
  __syncthreads();
  const int r = (k + P - 1) / P;
  const int m = (k + r - 1) / r;                   // <= P
  for (int p = tid; p < P; p += MERGE_NT) {        // rank of list p's element r - 1 among the P of them (keys are distinct
    const uint64_t mine = all[p * k + r - 1];      // except for padding, which ranks last in any order)
    int rank = 0;
    for (int o = 0; o < P; ++o) {
      const uint64_t other = all[o * k + r - 1];
      rank += (other > mine || (other == mine && o < p)) ? 1 : 0;
    }
    if (rank == m - 1) tau_s = mine;
  }
  __syncthreads();
  const uint64_t tau = tau_s;
  for (int e = tid; e < total; e += MERGE_NT) {
    const uint64_t key = all[e];
    if (key != kEmptyKey && key >= tau) {
      const int pos = atomicAdd(&n_top, 1);
      if (pos < SMALL_MERGE_TOP) top[pos] = key;
    }
  }
  __syncthreads();
  int nt = n_top;
  if (nt > SMALL_MERGE_TOP) {   // uniform; lists that are near copies of each other: sort everything (slow, rare)
    int pow2 = 128;
    while (pow2 < total) pow2 <<= 1;
    for (int e = total + tid; e < pow2; e += MERGE_NT) all[e] = kEmptyKey;
    __syncthreads();
    block_sort_desc_upto<SMALL_MERGE_KEYS, MERGE_NT>(all, total, tid);
    for (int j = tid; j < k; j += MERGE_NT) {
      const uint64_t key = j < total ? all[j] : kEmptyKey;
      D[qi * ldD + j] = key == kEmptyKey ? -FLT_MAX : key_score(key);
      I[qi * ldI + j] = key == kEmptyKey ? -1 : id_base + static_cast<int64_t>(key_row(key));
      if (j == k - 1) kth_s = key;
    }
  } else {
    for (int j = nt + tid; j < k; j += MERGE_NT) {   // fewer than k real keys in all lists together: padding
      D[qi * ldD + j] = -FLT_MAX;
      I[qi * ldI + j] = -1;
    }
    for (int c = tid; c < nt; c += MERGE_NT) {
      const uint64_t mine = top[c];
      int rank = 0;
      for (int o = 0; o < nt; ++o) rank += top[o] > mine ? 1 : 0;
      if (rank < k) {
        D[qi * ldD + rank] = key_score(mine);
        I[qi * ldI + rank] = id_base + static_cast<int64_t>(key_row(mine));
        if (rank == k - 1) kth_s = mine;
      }
    }
  }
  __syncthreads();
  if (ex.need_exact && tid == 0) {   // verify_topk_kernel's verdict for this query
    float b = -INFINITY;
    int f = 0;
    for (int g = 0; g < P; ++g) {
      if (ex.bounds) b = fmaxf(b, ex.bounds[g * ex.ld_b + qi]);
      f |= ex.flags[g * ex.ld_f + qi];
    }
    const uint64_t kthk = kth_s;
    const bool ok = f == 0 && (!ex.bounds || (kthk != kEmptyKey && key_score(kthk) >= b));
    ex.need_exact[qi] = ok ? 0 : 1;
  }
  if (ex.peer_bases) rsp::peer_consumed(ex.peer_bases, ex.PL, seq);
}

template <int CAP>
void launch_merge(const KeySource& src, int64_t nq, int P, int k, int64_t id_base, float* D, int64_t ldD, int64_t* I,
                  int64_t ldI, cudaStream_t st, const MergeExtras& ex) {
  rs::ProfScope prof(RS_PROF_MERGE, st);
  merge_topk_kernel<CAP><<<static_cast<unsigned>(nq), MERGE_NT, 0, st>>>(src, P, k, id_base, D, ldD, I, ldI, ex);
  RS_LAUNCH_CHECK();
}

void merge_dispatch(const KeySource& src, int64_t nq, int P, int k, int64_t id_base, float* D, int64_t ldD, int64_t* I,
                    int64_t ldI, cudaStream_t st, const MergeExtras& ex = MergeExtras{}) {
  RS_CHECK(k >= 1 && k <= 1024, "k must be in [1, 1024]");
  RS_CHECK(nq <= 0x7FFFFFFF, "too many queries");
  // short inputs -- the batch-1 scan's 148 lists of 10, the sharded search's G lists of k: position thresholds + rank counting
  // instead of rounds of offer / prune (a 256-key bitonic sort each)
  if (k <= 128 && (int64_t)P * k <= SMALL_MERGE_KEYS && (nq <= 64 || P <= 16)) {
    rs::ProfScope prof(RS_PROF_MERGE, st);
    merge_small_kernel<<<static_cast<unsigned>(nq), MERGE_NT, 0, st>>>(src, P, k, id_base, D, ldD, I, ldI, ex);
    RS_LAUNCH_CHECK();
    return;
  }
  if (k <= 128) launch_merge<256>(src, nq, P, k, id_base, D, ldD, I, ldI, st, ex);
  else if (k <= 256) launch_merge<512>(src, nq, P, k, id_base, D, ldD, I, ldI, st, ex);
  else if (k <= 512) launch_merge<1024>(src, nq, P, k, id_base, D, ldD, I, ldI, st, ex);
  else launch_merge<2048>(src, nq, P, k, id_base, D, ldD, I, ldI, st, ex);
}

template <int D, int QT, int CAP>
void launch_scan(const rs_index* h, const float* q, int64_t nq, int k, int P, uint64_t* partial, cudaStream_t st) {
  using L = ScanSmem<D, QT, CAP>;
  static std::once_flag once[64];
  std::call_once(once[h->device & 63], [] {
    cudaFuncSetAttribute(scan_topk_kernel<D, QT, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
  });
  const int64_t groups = (nq + QT - 1) / QT;
  RS_CHECK(groups <= 65535, "too many queries for one search call (limit 524,280)");
  dim3 grid(static_cast<unsigned>(P), static_cast<unsigned>(groups));
  rs::ProfScope prof(RS_PROF_SCAN, st);
  scan_topk_kernel<D, QT, CAP><<<grid, SCAN_NT, L::kDynamic, st>>>(h->tmap_f32, q, h->ntotal, static_cast<int>(nq), k,
                                                                  partial);
  RS_LAUNCH_CHECK();
}

template <int D>
void scan_dispatch(const rs_index* h, const float* q, int64_t nq, int k, int P, uint64_t* partial, cudaStream_t st) {
  if (k > 128) return launch_scan<D, 1, 2048>(h, q, nq, k, P, partial, st);
  if (nq >= 8) return launch_scan<D, 8, 256>(h, q, nq, k, P, partial, st);
  if (nq > 2) return launch_scan<D, 4, 256>(h, q, nq, k, P, partial, st);
  if (nq == 2) return launch_scan<D, 2, 256>(h, q, nq, k, P, partial, st);
  return launch_scan<D, 1, 256>(h, q, nq, k, P, partial, st);
}

void rebuild_tmaps(rs_index* h) {
  if (h->ntotal == 0) return;
  h->tmap_f32 = rs::make_tmap_2d(h->xb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)h->ntotal, (uint64_t)h->d,
                                 (uint64_t)h->d * 4, TILE_ROWS, 32);
}

}  // namespace

namespace rs {

void merge_lists_device(const float* Dparts, int64_t part_ld_D, const int64_t* Iparts, int64_t part_ld_I, int P, int64_t nq,
                        int k, float* D, int64_t ldD, int64_t* I, int64_t ldI, rs_peer* lists_region, const float* bounds,
                        int64_t ld_b, const int* flags, int64_t ld_f, int* need_exact, cudaStream_t st) {
  RS_CHECK(P >= 1 && P <= 4096, "nparts must be in [1, 4096]");
  RS_CHECK(need_exact == nullptr || flags != nullptr, "need_exact needs the shards' flags");
  if (nq == 0) return;
  MergeExtras ex;
  if (lists_region) {
    ex.peer_bases = lists_region->d_bases;
    ex.PL = rsp::layout_of(lists_region);
  }
  ex.bounds = bounds;
  ex.ld_b = ld_b;
  ex.flags = flags;
  ex.ld_f = ld_f;
  ex.need_exact = need_exact;
  KeySource src{nullptr, Dparts, Iparts, nq, part_ld_D, part_ld_I};
  merge_dispatch(src, nq, P, k, 0, D, ldD, I, ldI, st, ex);
}

void search_exact_device(rs_index* h, const float* q, int64_t nq, int k, int64_t id_base, float* D, int64_t ldD,
                         int64_t* I, int64_t ldI, cudaStream_t st) {
  RS_CHECK(k >= 1 && k <= 1024, "k must be in [1, 1024]");
  if (nq == 0) return;
  if (h->ntotal == 0) {
    fill_pad_kernel<<<64, 256, 0, st>>>(D, ldD, I, ldI, nq, k);
    RS_LAUNCH_CHECK();
    return;
  }
  RS_CHECK(h->ntotal <= 0xFFFFFFFFll, "catalog shards above 2^32 rows are not supported");
  const int64_t ntiles = (h->ntotal + TILE_ROWS - 1) / TILE_ROWS;
  int P = static_cast<int>(std::min<int64_t>(ntiles, sm_count()));
  if (ntiles < sm_count()) {
    // Small catalog, many queries (BASELINE config 1: 4,096 queries x 3,706 rows): one CTA per (tile, query group) is
    // 14,848 CTAs that each pay the pipeline start-up for a single 128-row tile.  Give every query group just enough
    // row slices to fill the machine about twice; a CTA then walks several tiles and fewer lists need merging.
    const int qt = k > 128 ? 1 : nq >= 8 ? 8 : nq > 2 ? 4 : static_cast<int>(nq);   // as scan_dispatch picks it
    const int64_t groups = (nq + qt - 1) / qt;
    P = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(P, (2 * (int64_t)sm_count() + groups - 1) / groups)));
  }
  uint64_t* partial = nullptr;
  RS_CUDA(cudaMallocAsync(&partial, sizeof(uint64_t) * (size_t)nq * P * k, st));
  try {
    switch (h->d) {
      case 32: scan_dispatch<32>(h, q, nq, k, P, partial, st); break;
      case 64: scan_dispatch<64>(h, q, nq, k, P, partial, st); break;
      case 128: scan_dispatch<128>(h, q, nq, k, P, partial, st); break;
      default: RS_CHECK(false, "index dimension must be 32, 64 or 128 (got " + std::to_string(h->d) + ")");
    }
    KeySource src{partial, nullptr, nullptr, nq};
    merge_dispatch(src, nq, P, k, id_base, D, ldD, I, ldI, st);
  } catch (...) {
    cudaFreeAsync(partial, st);
    throw;
  }
  RS_CUDA(cudaFreeAsync(partial, st));
}

// exact top-k of one or two queries through the bf16 shadow (scan_bf16_topk_kernel); same results as search_exact_device
bool bf16_scan_applies(const rs_index* h, int64_t nq, int k) {
  static const bool on = [] { const char* e = getenv("RS_BF16_SCAN"); return !(e && e[0] == '0'); }();
  return on && h->d == 128 && nq >= 1 && nq <= 2 && k >= 1 && k <= 128 && h->ntotal >= 65536 && h->ntotal <= 0x7FFFFF00ll;
}

template <int QT>
static void launch_bscan(const rs_index* h, const CUtensorMap& tm, const float* q, int nq, int k, int P, float eps_scale,
                         uint64_t* partial, int* ovf, cudaStream_t st) {
  using L = BScanSmem<QT>;
  static std::once_flag once[64];
  std::call_once(once[h->device & 63], [] {
    cudaFuncSetAttribute(scan_bf16_topk_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
  });
  rs::ProfScope prof(RS_PROF_SCAN, st);
  scan_bf16_topk_kernel<QT><<<P, BSCAN_NT, L::kDynamic, st>>>(tm, h->xb, q, h->ntotal, nq, k, eps_scale, partial, ovf);
  RS_LAUNCH_CHECK();
}

// need_exact != nullptr (device int32 [nq]): receives the per-query "margin overflowed, list not trusted" flags, nothing is
// synchronised.  nullptr: the flags are read back (one stream synchronise) and flagged queries repeated on the fp32 scan.
void search_bf16_scan_device(rs_index* h, const float* q, int64_t nq, int k, int64_t id_base, float* D, int64_t ldD, int64_t* I,
                             int64_t ldI, cudaStream_t st, int* need_exact) {
  RS_CHECK(bf16_scan_applies(h, nq, k), "the bf16-shadow scan serves 1-2 queries, d = 128, k <= 128, >= 65,536 rows");
  ensure_tensor_shadow(h, st);
  const int64_t ntiles = (h->ntotal + BSCAN_ROWS - 1) / BSCAN_ROWS;
  const int P = static_cast<int>(std::min<int64_t>(ntiles, sm_count()));
  const CUtensorMap& tm = h->tmap_bf16;   // box [256 rows, 64 bf16], 128-byte swizzle: the tensor path's own map
  // |s_hat - s| <= 2^-9 |q| |x| for bf16-rounded rows and fp32 queries; the tensor path's constant (both operands rounded,
  // 6 % slack for the fp32 accumulations) is used unchanged
  const float eps_scale = 0.00390625f * 1.0625f * h->max_norm;
  uint64_t* partial = nullptr;
  int* ovf = nullptr;
  RS_CUDA(cudaMallocAsync(&partial, sizeof(uint64_t) * (size_t)nq * P * k + 16, st));
  ovf = need_exact ? need_exact : reinterpret_cast<int*>(partial + (size_t)nq * P * k);
  int h_ovf[2] = {0, 0};
  try {
    RS_CUDA(cudaMemsetAsync(ovf, 0, sizeof(int) * static_cast<size_t>(nq), st));
    if (nq == 1) launch_bscan<1>(h, tm, q, 1, k, P, eps_scale, partial, ovf, st);
    else launch_bscan<2>(h, tm, q, 2, k, P, eps_scale, partial, ovf, st);
    KeySource src{partial, nullptr, nullptr, nq};
    merge_dispatch(src, nq, P, k, id_base, D, ldD, I, ldI, st);
    if (!need_exact) RS_CUDA(cudaMemcpyAsync(h_ovf, ovf, sizeof(int) * static_cast<size_t>(nq), cudaMemcpyDeviceToHost, st));
  } catch (...) {
    cudaFreeAsync(partial, st);
    throw;
  }
  RS_CUDA(cudaFreeAsync(partial, st));
  if (need_exact) return;
  RS_CUDA(cudaStreamSynchronize(st));
  int fb = 0;
  for (int64_t i = 0; i < nq; ++i)
    if (h_ovf[i]) {
      search_exact_device(h, q + i * h->d, 1, k, id_base, D + i * ldD, ldD, I + i * ldI, ldI, st);
      ++fb;
    }
  h->last_fallbacks.store(fb, std::memory_order_relaxed);
  h->total_fallbacks.fetch_add(fb, std::memory_order_relaxed);
}

}  // namespace rs

extern "C" {

int rs_index_create(int d, rs_index_t** out) {
  return rs::guard([&] {
    RS_CHECK(out != nullptr, "out is NULL");
    RS_CHECK(d == 32 || d == 64 || d == 128, "index dimension must be 32, 64 or 128");
    rs::sm_count();  // fails loudly when there is no usable sm_100 device: no CPU fallback
    auto* h = new rs_index();
    h->d = d;
    RS_CUDA(cudaGetDevice(&h->device));
    *out = h;
  });
}

int rs_index_free(rs_index_t* h) {
  return rs::guard([&] {
    if (!h) return;
    {
      rs::DeviceGuard g(h->device);
      cudaFree(h->xb);
      cudaFree(h->xb_bf16);
      cudaFree(h->row_norm);
    }
    delete h;
  });
}

int rs_index_ntotal(const rs_index_t* h, int64_t* out) {
  return rs::guard([&] {
    RS_CHECK(h && out, "NULL argument");
    *out = h->ntotal;
  });
}

int rs_index_dim(const rs_index_t* h, int* out) {
  return rs::guard([&] {
    RS_CHECK(h && out, "NULL argument");
    *out = h->d;
  });
}

int rs_index_reconstruct(const rs_index_t* h, int64_t i0, DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "index is NULL");
    require(out, "out", 2, kDLFloat, 32);
    RS_CHECK(out->shape[1] == h->d && is_compact(out), "out must be a contiguous [n, d] tensor");
    const int64_t n = out->shape[0];
    RS_CHECK(i0 >= 0 && n >= 0 && i0 + n <= h->ntotal, "row range out of bounds");
    RS_CHECK(on_device(out) || on_host(out), "out must be a CUDA or host tensor");
    if (n == 0) return;
    DeviceGuard g(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RS_CUDA(cudaMemcpyAsync(data_ptr<float>(out), h->xb + i0 * h->d, sizeof(float) * (size_t)n * h->d,
                            on_device(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (on_host(out)) RS_CUDA(cudaStreamSynchronize(st));
  });
}

int rs_index_fallbacks(const rs_index_t* h, int64_t* last, int64_t* total) {
  return rs::guard([&] {
    RS_CHECK(h && last && total, "NULL argument");
    *last = h->last_fallbacks.load(std::memory_order_relaxed);
    *total = h->total_fallbacks.load(std::memory_order_relaxed);
  });
}

int rs_debug_tensor_scores(rs_index_t* h, const DLTensor* q, DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "index is NULL");
    require(q, "q", 2, kDLFloat, 32);
    require(out, "out", 2, kDLFloat, 32);
    RS_CHECK(on_device(q) && on_device(out) && is_compact(q) && is_compact(out), "q and out must be contiguous CUDA tensors");
    RS_CHECK(q->shape[1] == h->d && out->shape[0] == q->shape[0] && out->shape[1] == h->ntotal, "shape mismatch");
    DeviceGuard g(h->device);
    tensor_scores_debug(h, data_ptr<float>(q), q->shape[0], data_ptr<float>(out), static_cast<cudaStream_t>(stream));
  });
}

int rs_index_add(rs_index_t* h, const DLTensor* x, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "index is NULL");
    require(x, "x", 2, kDLFloat, 32);
    RS_CHECK(x->shape[1] == h->d, "x must be [n, d] with d == index dimension");
    RS_CHECK(on_device(x) || on_host(x), "x must be a CUDA or host tensor");
    const int64_t n = x->shape[0];
    if (n == 0) return;
    const int64_t ldx = row_stride_2d(x, "x");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t need = h->ntotal + n;
    RS_CHECK(need <= 0xFFFFFFFFll, "catalog shards above 2^32 rows are not supported");
    if (need > h->capacity) {
      int64_t cap = h->capacity == 0 ? need : std::max(need, h->capacity + h->capacity / 2);
      float* nb = nullptr;
      RS_CUDA(cudaMalloc(&nb, sizeof(float) * (size_t)cap * h->d));
      if (h->ntotal > 0)
        RS_CUDA(cudaMemcpyAsync(nb, h->xb, sizeof(float) * (size_t)h->ntotal * h->d, cudaMemcpyDeviceToDevice, st));
      RS_CUDA(cudaStreamSynchronize(st));
      cudaFree(h->xb);
      h->xb = nb;
      h->capacity = cap;
      cudaFree(h->xb_bf16);
      cudaFree(h->row_norm);
      h->xb_bf16 = nullptr;
      h->row_norm = nullptr;
      h->shadow_rows = 0;
    }
    RS_CUDA(cudaMemcpy2DAsync(h->xb + h->ntotal * h->d, sizeof(float) * h->d, data_ptr<float>(x), sizeof(float) * ldx,
                              sizeof(float) * h->d, (size_t)n, on_device(x) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                              st));
    if (on_host(x)) RS_CUDA(cudaStreamSynchronize(st));
    h->ntotal = need;
    h->small_tensor_off.store(false, std::memory_order_relaxed);
    rebuild_tmaps(h);
  });
}

int rs_index_search(rs_index_t* h, const DLTensor* q, int k, int64_t id_base, int mode, DLTensor* D, DLTensor* I,
                    void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "index is NULL");
    require(q, "q", 2, kDLFloat, 32);
    require(D, "D", 2, kDLFloat, 32);
    require(I, "I", 2, kDLInt, 64);
    RS_CHECK(q->shape[1] == h->d, "q must be [nq, d] with d == index dimension");
    RS_CHECK(k >= 1 && k <= 1024, "k must be in [1, 1024]");
    const int64_t nq = q->shape[0];
    RS_CHECK(D->shape[0] == nq && D->shape[1] == k && I->shape[0] == nq && I->shape[1] == k, "D and I must be [nq, k]");
    RS_CHECK(mode == RS_EXACT_FP32 || mode == RS_TENSOR_BF16 || mode == RS_AUTO, "unknown mode");
    const bool dev = on_device(q) && on_device(D) && on_device(I);
    const bool host = on_host(q) && on_host(D) && on_host(I);
    RS_CHECK(dev || host, "q, D, I must be all CUDA tensors or all host tensors");
    if (nq == 0) return;
    DeviceGuard g(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bool tensor = false;
    if (mode == RS_TENSOR_BF16) {
      RS_CHECK(tensor_path_supported(h, nq, k), "tensor mode is not applicable to this shape (needs d=128, k<=128 and ntotal>=65536, or 256<=ntotal<=32768 with k<=ntotal/128)");
      tensor = true;
    } else if (mode == RS_AUTO) {
      // large catalogs: faster than the fp32 scan at every batch size (bf16 shadow = half the bytes); small (L2-resident)
      // ones: from 128 queries -- below that the scan's two launches beat the tensor path's seven (measured, config 1 shape)
      tensor = tensor_path_supported(h, nq, k) &&
               (h->ntotal >= 65536 || (nq >= 128 && !h->small_tensor_off.load(std::memory_order_relaxed)));
    }
    auto run = [&](const float* qp, float* Dp, int64_t ldD, int64_t* Ip, int64_t ldI) {
      if (tensor) {
        search_tensor_device(h, qp, nq, k, id_base, Dp, ldD, Ip, ldI, st);
        if (mode == RS_AUTO && h->ntotal < 65536 && 2 * h->last_fallbacks.load(std::memory_order_relaxed) >= nq)
          h->small_tensor_off.store(true, std::memory_order_relaxed);
      } else {
        search_exact_device(h, qp, nq, k, id_base, Dp, ldD, Ip, ldI, st);
      }
    };
    if (dev) {
      RS_CHECK(is_compact(q), "q must be contiguous");
      RS_CHECK(q->device.device_id == h->device, "q is on a different device than the index");
      require_aligned16(data_ptr<float>(q), "q");
      run(data_ptr<float>(q), data_ptr<float>(D), row_stride_2d(D, "D"), data_ptr<int64_t>(I), row_stride_2d(I, "I"));
      return;
    }
    // host tensors: stage through the device on `st` and return once the results are back
    RS_CHECK(is_compact(q) && is_compact(D) && is_compact(I), "host q, D, I must be contiguous");
    float* dq = nullptr;
    float* dD = nullptr;
    int64_t* dI = nullptr;
    RS_CUDA(cudaMallocAsync(&dq, sizeof(float) * (size_t)nq * h->d, st));
    RS_CUDA(cudaMallocAsync(&dD, sizeof(float) * (size_t)nq * k, st));
    RS_CUDA(cudaMallocAsync(&dI, sizeof(int64_t) * (size_t)nq * k, st));
    try {
      RS_CUDA(cudaMemcpyAsync(dq, data_ptr<float>(q), sizeof(float) * (size_t)nq * h->d, cudaMemcpyHostToDevice, st));
      run(dq, dD, k, dI, k);
      RS_CUDA(cudaMemcpyAsync(data_ptr<float>(D), dD, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
      RS_CUDA(cudaMemcpyAsync(data_ptr<int64_t>(I), dI, sizeof(int64_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
    } catch (...) {
      cudaFreeAsync(dq, st);
      cudaFreeAsync(dD, st);
      cudaFreeAsync(dI, st);
      cudaStreamSynchronize(st);
      throw;
    }
    cudaFreeAsync(dq, st);
    cudaFreeAsync(dD, st);
    cudaFreeAsync(dI, st);
    RS_CUDA(cudaStreamSynchronize(st));
  });
}

int rs_index_search_deferred(rs_index_t* h, const DLTensor* q, int k, int64_t id_base, int mode, DLTensor* D, DLTensor* I,
                             DLTensor* need_exact, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "index is NULL");
    require(q, "q", 2, kDLFloat, 32);
    require(D, "D", 2, kDLFloat, 32);
    require(I, "I", 2, kDLInt, 64);
    require(need_exact, "need_exact", 1, kDLInt, 32);
    RS_CHECK(q->shape[1] == h->d, "q must be [nq, d] with d == index dimension");
    RS_CHECK(k >= 1 && k <= 1024, "k must be in [1, 1024]");
    const int64_t nq = q->shape[0];
    RS_CHECK(D->shape[0] == nq && D->shape[1] == k && I->shape[0] == nq && I->shape[1] == k, "D and I must be [nq, k]");
    RS_CHECK(need_exact->shape[0] == nq && (nq <= 1 || stride_of(need_exact, 0) == 1), "need_exact must be a contiguous [nq] tensor");
    RS_CHECK(mode == RS_EXACT_FP32 || mode == RS_TENSOR_BF16 || mode == RS_AUTO, "unknown mode");
    RS_CHECK(on_device(q) && on_device(D) && on_device(I) && on_device(need_exact), "the deferred search takes CUDA tensors only");
    if (nq == 0) return;
    DeviceGuard g(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RS_CHECK(is_compact(q), "q must be contiguous");
    RS_CHECK(q->device.device_id == h->device, "q is on a different device than the index");
    require_aligned16(data_ptr<float>(q), "q");
    const bool supported = tensor_path_supported(h, nq, k);
    RS_CHECK(mode != RS_TENSOR_BF16 || supported, "tensor mode is not applicable to this shape (needs d=128, k<=128 and ntotal>=65536, or 256<=ntotal<=32768 with k<=ntotal/128)");
    if (mode != RS_EXACT_FP32 && supported) {
      search_tensor_device(h, data_ptr<float>(q), nq, k, id_base, data_ptr<float>(D), row_stride_2d(D, "D"),
                           data_ptr<int64_t>(I), row_stride_2d(I, "I"), st, data_ptr<int>(need_exact));
    } else {
      search_exact_device(h, data_ptr<float>(q), nq, k, id_base, data_ptr<float>(D), row_stride_2d(D, "D"),
                          data_ptr<int64_t>(I), row_stride_2d(I, "I"), st);
      RS_CUDA(cudaMemsetAsync(data_ptr<int>(need_exact), 0, sizeof(int) * (size_t)nq, st));
    }
  });
}

int rs_normalize_l2(DLTensor* x, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(x, "x", 2, kDLFloat, 32);
    const int64_t n = x->shape[0];
    const int d = static_cast<int>(x->shape[1]);
    if (n == 0 || d == 0) return;
    const int64_t ld = row_stride_2d(x, "x");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sm_count();
    const unsigned blocks = static_cast<unsigned>((n + 127) / 128);
    if (on_device(x)) {
      normalize_l2_kernel<<<blocks, 128, 0, st>>>(data_ptr<float>(x), ld, n, d);
      RS_LAUNCH_CHECK();
      return;
    }
    RS_CHECK(on_host(x), "x must be a CUDA or host tensor");
    float* dx = nullptr;
    RS_CUDA(cudaMallocAsync(&dx, sizeof(float) * (size_t)n * d, st));
    try {
      RS_CUDA(cudaMemcpy2DAsync(dx, sizeof(float) * d, data_ptr<float>(x), sizeof(float) * ld, sizeof(float) * d, (size_t)n,
                                cudaMemcpyHostToDevice, st));
      normalize_l2_kernel<<<blocks, 128, 0, st>>>(dx, d, n, d);
      RS_LAUNCH_CHECK();
      RS_CUDA(cudaMemcpy2DAsync(data_ptr<float>(x), sizeof(float) * ld, dx, sizeof(float) * d, sizeof(float) * d, (size_t)n,
                                cudaMemcpyDeviceToHost, st));
    } catch (...) {
      cudaFreeAsync(dx, st);
      cudaStreamSynchronize(st);
      throw;
    }
    cudaFreeAsync(dx, st);
    RS_CUDA(cudaStreamSynchronize(st));
  });
}

int rs_matvec(const DLTensor* items, const DLTensor* user, DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(items, "items", 2, kDLFloat, 32);
    require(user, "user", 1, kDLFloat, 32);
    require(out, "out", 1, kDLFloat, 32);
    RS_CHECK(on_device(items) && on_device(user) && on_device(out), "all tensors must be CUDA tensors");
    const int64_t n = items->shape[0];
    const int d = static_cast<int>(items->shape[1]);
    RS_CHECK(user->shape[0] == d && out->shape[0] == n, "shape mismatch: items [n, d], user [d], out [n]");
    RS_CHECK((d <= 1 || stride_of(user, 0) == 1) && (n <= 1 || stride_of(out, 0) == 1), "user and out must be contiguous");
    if (n == 0) return;
    sm_count();
    matvec_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        data_ptr<float>(items), row_stride_2d(items, "items"), data_ptr<float>(user), data_ptr<float>(out), n, d);
    RS_LAUNCH_CHECK();
  });
}

int rs_topk_verify(const DLTensor* D, const DLTensor* I, const DLTensor* bounds, const DLTensor* flags,
                   DLTensor* need_exact, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(D, "D", 2, kDLFloat, 32);
    require(I, "I", 2, kDLInt, 64);
    require(bounds, "bounds", 2, kDLFloat, 32);
    require(flags, "flags", 2, kDLInt, 32);
    require(need_exact, "need_exact", 1, kDLInt, 32);
    const int64_t nq = D->shape[0], k = D->shape[1], G = bounds->shape[0];
    RS_CHECK(I->shape[0] == nq && I->shape[1] == k && bounds->shape[1] == nq && flags->shape[0] == G && flags->shape[1] == nq &&
                 need_exact->shape[0] == nq, "shapes: D, I [nq, k]; bounds, flags [shards, nq]; need_exact [nq]");
    RS_CHECK(on_device(D) && on_device(I) && on_device(bounds) && on_device(flags) && on_device(need_exact), "all tensors must be CUDA tensors");
    RS_CHECK((nq <= 1 || (stride_of(bounds, 1) == 1 && stride_of(flags, 1) == 1 && stride_of(need_exact, 0) == 1)) && k >= 1,
             "bounds / flags rows and need_exact must have unit stride");
    if (nq == 0) return;
    sm_count();
    verify_topk_kernel<<<static_cast<unsigned>((nq + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        data_ptr<float>(D), row_stride_2d(D, "D"), data_ptr<int64_t>(I), row_stride_2d(I, "I"), static_cast<int>(k),
        data_ptr<float>(bounds), G > 1 ? stride_of(bounds, 0) : nq, data_ptr<int>(flags), G > 1 ? stride_of(flags, 0) : nq,
        static_cast<int>(G), nq, data_ptr<int>(need_exact));
    RS_LAUNCH_CHECK();
  });
}

int rs_topk_merge(const DLTensor* Dparts, const DLTensor* Iparts, DLTensor* D, DLTensor* I, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(Dparts, "Dparts", 3, kDLFloat, 32);
    require(Iparts, "Iparts", 3, kDLInt, 64);
    require(D, "D", 2, kDLFloat, 32);
    require(I, "I", 2, kDLInt, 64);
    RS_CHECK(on_device(Dparts) && on_device(Iparts) && on_device(D) && on_device(I), "all tensors must be CUDA tensors");
    const int64_t P = Dparts->shape[0], nq = Dparts->shape[1], k = Dparts->shape[2];
    RS_CHECK(Iparts->shape[0] == P && Iparts->shape[1] == nq && Iparts->shape[2] == k, "Iparts must match Dparts");
    RS_CHECK(D->shape[0] == nq && D->shape[1] == k && I->shape[0] == nq && I->shape[1] == k, "D and I must be [nq, k]");
    // each part is a packed [nq, k] block; the parts themselves may be strided (slices of one exchange buffer)
    auto part_ld = [&](const DLTensor* t) -> int64_t {
      if (!t->strides) return nq * k;
      RS_CHECK((k <= 1 || t->strides[2] == 1) && (nq <= 1 || t->strides[1] == k) && (P <= 1 || t->strides[0] >= nq * k),
               "Dparts / Iparts: each part must be a contiguous [nq, k] block");
      return P <= 1 ? nq * k : t->strides[0];
    };
    RS_CHECK(P >= 1 && P <= 4096, "nparts must be in [1, 4096]");
    if (nq == 0) return;
    sm_count();
    KeySource src{nullptr, data_ptr<float>(Dparts), data_ptr<int64_t>(Iparts), nq, part_ld(Dparts), part_ld(Iparts)};
    merge_dispatch(src, nq, static_cast<int>(P), static_cast<int>(k), 0, data_ptr<float>(D), row_stride_2d(D, "D"),
                   data_ptr<int64_t>(I), row_stride_2d(I, "I"), static_cast<cudaStream_t>(stream));
  });
}

}  // extern "C"
