// dense_x3.cu -- Keras Dense, y = act(x @ kernel + bias), at fp32 accuracy ON THE TENSOR CORES ("fp32 tensor mode").
// Replaces keras.layers.Dense in user_tower / item_tower / deep_nets (src/models.py:26-29,76-77) -- and the Embedding
// gather in front of a tower (src/models.py:85,89) -- for the precision the reference actually computes in (fp32), at
// ranking batch sizes.  The FFMA kernel (dense.cu) stays the small-batch path.
//
// Arithmetic: every fp32 operand is split into two fp16 numbers, v = hi + lo * 2^-11 with hi = fp16(v) and
// lo = fp16((v - hi) * 2^11) (the scaling keeps lo out of fp16's subnormal range); hi + lo * 2^-11 carries 22 significant
// bits.  x.w = hi_x hi_w + 2^-11 (hi_x lo_w + lo_x hi_w) + O(2^-22): three tcgen05.mma (kind::f16, fp32 accumulate)
// per K step, the main term and the two correction terms into SEPARATE tensor-memory accumulators, combined in the
// epilogue (main + 2^-11 corr).  Dropped: lo_x lo_w, 2^-22 relative -- below fp32's own rounding of a K-long sum.
// Measured against fp64 in tests/test_gpu_dense_x3.py (<= 1.1e-6 of tensor scale per layer, the same as the FFMA kernel);
// the stated tolerance of the fp32 mode is 1e-5 of tensor scale.  Range: |v| < 65504 (fp16).
//
// Activations BETWEEN layers travel in the split form -- two fp16 planes [B, width] (hi, lo), the same bytes as fp32 --
// written by the producing layer's epilogue, so the consuming layer's operands are plain TMA loads (3 x 32 KB in flight per
// SM, no registers, no conversion on the load path).  A first version converted fp32 rows in the producer warps: one
// 32 KB chunk in flight per SM made it latency-bound at ~2.4 TB/s (ncu: long-scoreboard stalls, 19 % of DRAM peak).  The
// converting producer remains for layers whose input is fp32 -- the first tower layer, which gathers embedding rows by
// id straight out of the table -- with two chunks in flight.
//
// One persistent CTA per SM, weight-stationary; up to two independent layers of the same shape per launch (the two
// towers): blockIdx.x % n_prob picks the problem.
//   * the layer's kernel, pre-split and transposed to [out][in] fp16 hi / lo images (rs_dense_x3_prepare, once per weight
//     change), is TMA-loaded once per CTA: <= 128 KB (N tile 128 x K 256 x 2 images, or both N halves of a K = 128 layer);
//   * a [128 rows x 64] K chunk of the input = hi image | lo image, 32 KB, in a 2-4 stage ring in the 128-byte swizzled
//     K-major layout UMMA reads: filled by TMA (split input) or by 8 converting warps (fp32 / gathered input);
//   * one thread issues 3 x 4 MMAs per chunk (M128 x N<=128 x K16) into one of two TMEM slots (main | corr, 256 columns);
//   * 4 epilogue warps combine, add the bias, apply the activation and hand fp32 rows and / or the split planes to TMA
//     stores through two 4 KB staging buffers per warp.
// Per sample and layer: 3 x 2 x in x out tensor flop; bytes: 4 in read + 4 out written.
#include <algorithm>
#include <mutex>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace {

using namespace rsd;

constexpr int BM = 128;
constexpr int X3_PROD_WARPS = 8;                 // converting kernels: warps 0-7 split fp32 rows
// Warp roles.  Registers are allocated per four warps: 13 warps -> 16 (128 registers per thread; the converting producers
// hold two chunks in flight) and 6 -> 8 (TMA input).  Epilogue warps start at a multiple of 4: a warp may only touch the
// TMEM lane quarter warp % 4.
template <bool TMA_IN>
struct X3Roles {
  static constexpr int EPI0 = TMA_IN ? 4 : 8;
  static constexpr int EPI_WARPS = 4;
  static constexpr int ISSUER = TMA_IN ? 1 : 12;          // TMA input: warp 0 = TMA producer, warps 2-3 idle
  static constexpr int THREADS = TMA_IN ? 8 * 32 : 13 * 32;
};
constexpr int X3_MAX_STAGES = 4;                 // ring depth is a launch parameter: whatever fits beside the kernel images (2..4)
constexpr int X3_STAGE_BYTES = 2 * BM * 128;     // hi image [128 x 64 fp16] | lo image
constexpr int X3_OUT_BYTES = 4096;               // one output piece: fp32 [32 rows x 32] or the hi | lo planes of it
constexpr int X3_OUT_STAGING = 4 * 2 * X3_OUT_BYTES;   // two pieces per epilogue warp
constexpr float kLoScale = 2048.0f;              // 2^11
constexpr float kLoInv = 1.0f / 2048.0f;

struct X3Problem {
  int64_t rows;
  const float* x;        // fp32 rows [rows, K] (row stride ldx), or the embedding table when ids != nullptr; nullptr: split input
  int64_t ldx;
  const void* ids;       // input row b = x[ids[b]]; ids outside [0, table_rows) read as zeros
  int ids64;
  int64_t table_rows;
  const float* bias;     // [N] or nullptr
  float* y;              // fp32 output [rows, N] (row stride ldy) or nullptr
  int64_t ldy;
  __half* y_hi;          // split output planes [rows, N] (row stride ldy16) or nullptr
  __half* y_lo;
  int64_t ldy16;
};

struct X3Params {
  X3Problem prob[2];
  int n_prob;
  int act;
  int nh;                // N tiles per CTA (1 or 2); grid.y covers the rest
  int stages;            // ring depth (2..4)
};

struct X3Maps {
  CUtensorMap w_hi[2], w_lo[2];   // per problem: kernel images [N, K], box [NT, 64]
  CUtensorMap x_hi[2], x_lo[2];   // per problem: split input planes [rows, K], box [128, 64] (TMA input only)
  CUtensorMap y[2];               // per problem: fp32 output [rows, N], box [32 rows, 32 floats], 128-byte swizzle
  CUtensorMap y_hi[2], y_lo[2];   // per problem: split output planes [rows, N], box [32 rows, 32 fp16], 64-byte swizzle
};

// instruction descriptor, kind::f16: D = f32, A = B = fp16 (format 0), both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16_f32(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// K = 256 layers keep the 128 KB of kernel images and a 3-stage ring (their MMA time per chunk is long enough that two
// stages leave bubbles: measured 34 -> 39 us) and store straight from registers; narrower layers have room for the
// TMA-store staging.
template <int K>
constexpr bool x3_out_tma() { return K < 256; }
template <int K, int NT>
constexpr int x3_smem(int nh, int stages) {
  return nh * (K / 64) * NT * 128 * 2 + stages * X3_STAGE_BYTES + (x3_out_tma<K>() ? X3_OUT_STAGING : 0) + 1024 /*bias*/ +
         256 /*barriers*/ + 1024 /*align*/;
}

// v -> (hi, lo) with v ~= hi + lo * 2^-11 (scalar form: the weight-split kernel)
__device__ __forceinline__ void split16(float v, __half& h, __half& l) {
  h = __float2half_rn(v);
  l = __float2half_rn((v - __half2float(h)) * kLoScale);
}
// The same for a pair, packed: (a, b) -> hi = {fp16(a), fp16(b)}, lo likewise (a in the low half).  cvt.rn.f16x2.f32 is one
// full-rate F2FP for two values; the scalar F2F.F16.F32 the compiler emits for __float2half_rn runs at 16 per clock and SM
// and was the bottleneck of the first version (3 conversions per element in the producers and again in the epilogue).
__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void split16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = cvt_f16x2(a, b);
  const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = cvt_f16x2((a - back.x) * kLoScale, (b - back.y) * kLoScale);
}

template <int K, int NT, bool TMA_IN>
__global__ void __launch_bounds__(X3Roles<TMA_IN>::THREADS, 1)
dense_x3_kernel(const __grid_constant__ X3Maps maps, const __grid_constant__ X3Params p) {
  constexpr int KC = K / 64;
  constexpr int X3_EPI0 = X3Roles<TMA_IN>::EPI0, X3_EPI_WARPS = X3Roles<TMA_IN>::EPI_WARPS, X3_ISSUER = X3Roles<TMA_IN>::ISSUER;
  constexpr int X3_THREADS = X3Roles<TMA_IN>::THREADS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nh = p.nh;
  const int w_img = nh * KC * NT * 128;          // bytes of one weight image (hi or lo)
  uint8_t* w_hi = base;
  uint8_t* w_lo = base + w_img;
  uint8_t* ring = base + 2 * w_img;
  const uint32_t X3_STAGES = static_cast<uint32_t>(p.stages);
  constexpr bool OUT_TMA = x3_out_tma<K>();
  uint8_t* out_stage = ring + X3_STAGES * X3_STAGE_BYTES;                        // [4 epilogue warps][2][4 KB] (OUT_TMA)
  float* bias_s = reinterpret_cast<float*>(out_stage + (OUT_TMA ? X3_OUT_STAGING : 0));   // [nh * NT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 256);
  uint64_t* wfull = bars;
  uint64_t* full = bars + 1;                     // [STAGES] producers -> issuer
  uint64_t* empty = full + X3_MAX_STAGES;        // [STAGES] commit -> producers
  uint64_t* tfull = empty + X3_MAX_STAGES;       // [2] commit -> epilogue
  uint64_t* tempty = tfull + 2;                  // [2] epilogue -> issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int pi = blockIdx.x % p.n_prob;
  const int cta = blockIdx.x / p.n_prob;
  const int ncta = (gridDim.x - pi + p.n_prob - 1) / p.n_prob;
  const X3Problem& pr = p.prob[pi];
  const int64_t ntiles = (pr.rows + BM - 1) / BM;
  const int64_t n_my = cta < ntiles ? (ntiles - cta + ncta - 1) / ncta : 0;
  // Work items.  TMA input: (row tile, N tile) pairs, the N tiles of a row tile back to back, two TMEM slots (main | corr
  // each) so the epilogue of one item overlaps the MMAs of the next.  Converting producers: ONE item per row tile -- both N
  // tiles are computed from the same converted chunk (the conversion, not the tensor pipe, is what that kernel is short
  // of: ncu showed it issue-bound with the rows converted once per N tile), into all four accumulators of the one slot.
  constexpr bool SHARED_A = !TMA_IN;
  const int halves = SHARED_A ? nh : 1;          // N tiles computed per item
  const int64_t n_items = SHARED_A ? n_my : n_my * nh;
  const int n0 = blockIdx.y * nh * NT;           // first output column of this CTA

  if (warp == X3_ISSUER && lane == 0) {
    tma_prefetch_desc(&maps.w_hi[pi]);
    tma_prefetch_desc(&maps.w_lo[pi]);
    if (TMA_IN) {
      tma_prefetch_desc(&maps.x_hi[pi]);
      tma_prefetch_desc(&maps.x_lo[pi]);
    }
    mbar_init(wfull, 1);
    if (pr.y) tma_prefetch_desc(&maps.y[pi]);
    if (pr.y_hi) {
      tma_prefetch_desc(&maps.y_hi[pi]);
      tma_prefetch_desc(&maps.y_lo[pi]);
    }
    for (uint32_t s = 0; s < X3_STAGES; ++s) {
      mbar_init(&full[s], TMA_IN ? 1 : X3_PROD_WARPS);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], X3_EPI_WARPS);
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < nh * NT; i += X3_THREADS) bias_s[i] = pr.bias ? pr.bias[n0 + i] : 0.0f;
  if (warp == X3_ISSUER) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == X3_ISSUER) {
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, 2 * w_img);
      for (int hh = 0; hh < nh; ++hh)
        for (int kc = 0; kc < KC; ++kc) {
          const int off = (hh * KC + kc) * NT * 128;
          tma_load_2d(w_hi + off, &maps.w_hi[pi], wfull, kc * 64, n0 + hh * NT, kEvictLast);
          tma_load_2d(w_lo + off, &maps.w_lo[pi], wfull, kc * 64, n0 + hh * NT, kEvictLast);
        }
      constexpr uint32_t idesc = umma_idesc_f16_f32(BM, NT);
      mbar_wait(wfull, 0);
      const uint64_t whi0 = umma_desc_k_sw128(smem_u32(w_hi)), wlo0 = umma_desc_k_sw128(smem_u32(w_lo));
      const uint64_t ring0 = umma_desc_k_sw128(smem_u32(ring));
      uint32_t it = 0;
      for (int64_t i = 0; i < n_items; ++i) {
        const uint32_t a = SHARED_A ? 0u : (static_cast<uint32_t>(i) & 1u);
        const uint32_t aph = SHARED_A ? (static_cast<uint32_t>(i) & 1u) : ((static_cast<uint32_t>(i) >> 1) & 1u);
        mbar_wait(&tempty[a], aph ^ 1u);
        tc_fence_after_sync();
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const uint32_t s = it % X3_STAGES, ph = (it / X3_STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          for (int h2 = 0; h2 < halves; ++h2) {
            const int hh = SHARED_A ? h2 : static_cast<int>(i % nh);
            const uint32_t d_main = tmem_base + (SHARED_A ? h2 * 256 : a * 256), d_corr = d_main + 128;
            const uint32_t woff = static_cast<uint32_t>((hh * KC + kc) * NT * 128);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t a_hi = ring0 + static_cast<uint64_t>((s * X3_STAGE_BYTES + k4 * 32) >> 4);
              const uint64_t a_lo = a_hi + static_cast<uint64_t>((BM * 128) >> 4);
              const uint64_t b_hi = whi0 + static_cast<uint64_t>((woff + k4 * 32) >> 4);
              const uint64_t b_lo = wlo0 + static_cast<uint64_t>((woff + k4 * 32) >> 4);
              const uint32_t acc = (kc | k4) ? 1u : 0u;
              umma_bf16_ss(d_main, a_hi, b_hi, idesc, acc);
              umma_bf16_ss(d_corr, a_hi, b_lo, idesc, acc);
              umma_bf16_ss(d_corr, a_lo, b_hi, idesc, 1u);
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[a]);
      }
    }
  } else if (TMA_IN && warp == 0) {
    // ---- TMA producer: the hi / lo planes of the previous layer's output are operand images already ----
    if (elect_one()) {
      uint32_t it = 0;
      for (int64_t i = 0; i < n_items; ++i) {
        const int64_t tile = cta + (i / nh) * ncta;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const uint32_t s = it % X3_STAGES, ph = (it / X3_STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], X3_STAGE_BYTES);
          uint8_t* st_hi = ring + s * X3_STAGE_BYTES;
          tma_load_2d(st_hi, &maps.x_hi[pi], &full[s], kc * 64, static_cast<int>(tile * BM), kEvictFirst);
          tma_load_2d(st_hi + BM * 128, &maps.x_lo[pi], &full[s], kc * 64, static_cast<int>(tile * BM), kEvictFirst);
        }
      }
    }
  } else if (!TMA_IN && warp < X3_PROD_WARPS) {
    // ---- converting producers: fp32 rows (or gathered table rows) -> (hi, lo) fp16 images of a [128 x 64] K chunk ----
    // a half-warp covers one row's 64 floats (16 x float4); a warp instruction covers 2 rows, a warp 16 rows of the chunk.
    // Two chunks are in flight in registers (bufA / bufB) while a third is being split.
    const int rsub = lane >> 4;        // which of the 2 rows
    const int j = lane & 15;           // float4 index within the row's chunk
    const int64_t total = n_items * KC;
    // Source row of each of this thread's 8 rows of a tile: the row itself, the table row its id names, or -1 (past the
    // batch / id outside the table: reads as zeros).  The ids of the NEXT tile are fetched raw when the load stream enters a
    // tile and only looked at when it enters the next one (a first version loaded and range-checked them per chunk: the
    // id -> row dependency was two exposed L2 latencies per chunk, 60 % of the kernel's stall samples).
    int rid[8];
    uint32_t raw_lo[8], raw_hi[8];
    int64_t ld_tile = -1;
    auto fetch_raw = [&](int64_t tile) {
      if (!pr.ids) return;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t row = tile * BM + warp * 16 + e * 2 + rsub;
        raw_lo[e] = 0xFFFFFFFFu;
        raw_hi[e] = 0xFFFFFFFFu;             // -1: invalid
        if (tile < ntiles && row < pr.rows) {
          if (pr.ids64) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(pr.ids) + row);
            raw_lo[e] = v.x;
            raw_hi[e] = v.y;
          } else {
            raw_lo[e] = __ldg(static_cast<const uint32_t*>(pr.ids) + row);
          }
        }
      }
    };
    auto promote = [&](int64_t tile) {   // raw ids of `tile` (or its row numbers) -> rid
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t row = tile * BM + warp * 16 + e * 2 + rsub;
        if (pr.ids) {
          const int64_t id = pr.ids64 ? static_cast<int64_t>((static_cast<uint64_t>(raw_hi[e]) << 32) | raw_lo[e])
                                      : static_cast<int64_t>(static_cast<int32_t>(raw_lo[e]));
          rid[e] = (row < pr.rows && id >= 0 && id < pr.table_rows) ? static_cast<int>(id) : -1;
        } else {
          rid[e] = row < pr.rows ? static_cast<int>(row) : -1;
        }
      }
    };
    auto load_chunk = [&](float4 (&buf)[8], int64_t g) {
      const int64_t tile = cta + (g / KC) * ncta;     // converting producers: one item (all N tiles) per row tile
      const int kc = static_cast<int>(g % KC);
      if (tile != ld_tile) {
        if (ld_tile < 0) fetch_raw(tile);
        promote(tile);
        ld_tile = tile;
        fetch_raw(tile + ncta);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e)
        buf[e] = rid[e] >= 0 ? __ldg(reinterpret_cast<const float4*>(pr.x + static_cast<int64_t>(rid[e]) * pr.ldx + kc * 64) + j)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto convert_chunk = [&](const float4 (&buf)[8], int64_t g) {
      const uint32_t s = static_cast<uint32_t>(g % X3_STAGES), ph = static_cast<uint32_t>((g / X3_STAGES) & 1);
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* st_hi = ring + s * X3_STAGE_BYTES;
      uint8_t* st_lo = st_hi + BM * 128;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int r = warp * 16 + e * 2 + rsub;                       // row within the tile
        uint2 h2, l2;
        split16x2(buf[e].x, buf[e].y, h2.x, l2.x);
        split16x2(buf[e].z, buf[e].w, h2.y, l2.y);
        // 16-byte chunk c = j / 2 of the 128-byte row, XOR-swizzled with the row's position in its 8-row group
        const uint32_t off = static_cast<uint32_t>(r * 128 + (((j >> 1) ^ (r & 7)) << 4) + (j & 1) * 8);
        *reinterpret_cast<uint2*>(st_hi + off) = h2;
        *reinterpret_cast<uint2*>(st_lo + off) = l2;
      }
      fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    };
    float4 bufA[8], bufB[8];
    if (total > 0) load_chunk(bufA, 0);
    if (total > 1) load_chunk(bufB, 1);
    for (int64_t g = 0; g < total; g += 2) {
      convert_chunk(bufA, g);
      if (g + 2 < total) load_chunk(bufA, g + 2);
      if (g + 1 < total) {
        convert_chunk(bufB, g + 1);
        if (g + 3 < total) load_chunk(bufB, g + 3);
      }
    }
  } else if (warp >= X3_EPI0 && warp < X3_EPI0 + X3_EPI_WARPS) {
    // ---- epilogue: main + 2^-11 corr + bias -> activation -> fp32 rows and / or the split planes ----
    // Every output piece (32 rows x 32 columns of one warp) goes through one of the warp's two 4 KB staging buffers and
    // leaves as a TMA store; rows past the batch are clipped by the tensor map.  The first version stored straight from
    // registers, one row per thread: each 256-bit store instruction touched 32 different lines, and the LSU's one tag
    // lookup per cycle made the stores the longest phase of every layer (128 KB per tile = ~16 k cycles).
    const int quarter = warp & 3;
    const int ew = warp - X3_EPI0;
    uint8_t* const stg0 = out_stage + ew * 2 * X3_OUT_BYTES;
    uint32_t nst = 0;
    auto stage_acquire = [&]() -> uint8_t* {
      if (lane == 0) bulk_wait_read<1>();        // the store that last read this buffer (two pieces ago) is done with it
      __syncwarp();
      return stg0 + (nst++ & 1u) * X3_OUT_BYTES;
    };
    for (int64_t i = 0; i < n_items; ++i) {
      const int64_t tile = cta + (SHARED_A ? i : i / nh) * ncta;
      const int row0 = static_cast<int>(tile * BM) + quarter * 32;
      const uint32_t a = SHARED_A ? 0u : (static_cast<uint32_t>(i) & 1u);
      const uint32_t aph = SHARED_A ? (static_cast<uint32_t>(i) & 1u) : ((static_cast<uint32_t>(i) >> 1) & 1u);
      mbar_wait(&tfull[a], aph);
      tc_fence_after_sync();
#pragma unroll 1
      for (int h2 = 0; h2 < halves; ++h2) {
      const int hh = SHARED_A ? h2 : static_cast<int>(i % nh);
      const uint32_t taddr = tmem_base + (SHARED_A ? h2 * 256 : a * 256) + (static_cast<uint32_t>(quarter * 32) << 16);
      const int col_base = n0 + hh * NT;
#pragma unroll 1
      for (int c = 0; c < NT / 32; ++c) {
        float m[32], cr[32];
        __syncwarp();
        tmem_ld_32x32(taddr + c * 32, m);
        tmem_ld_32x32(taddr + 128 + c * 32, cr);
        if (c == NT / 32 - 1 && h2 == halves - 1) {   // the accumulators are in registers: the issuer may start the next item
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[a]);
        }
        const float* bs = bias_s + hh * NT + c * 32;
#pragma unroll
        for (int e = 0; e < 32; ++e) m[e] = fmaf(cr[e], kLoInv, m[e]) + bs[e];
        if (p.act == RS_ACT_RELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) m[e] = fmaxf(m[e], 0.0f);
        }
        const int col = col_base + c * 32;
        if constexpr (!OUT_TMA) {
          const int64_t row = static_cast<int64_t>(row0) + lane;
          if (row < pr.rows) {
            if (pr.y) {
              float* yrow = pr.y + row * pr.ldy + col;
#pragma unroll
              for (int e = 0; e < 32; e += 8)
                st_global_256(yrow + e, m[e], m[e + 1], m[e + 2], m[e + 3], m[e + 4], m[e + 5], m[e + 6], m[e + 7]);
            }
            if (pr.y_hi) {
              uint32_t ph[16], pl[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) split16x2(m[2 * e], m[2 * e + 1], ph[e], pl[e]);
              __half* hrow = pr.y_hi + row * pr.ldy16 + col;
              __half* lrow = pr.y_lo + row * pr.ldy16 + col;
              st_global_256(hrow, ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7]);
              st_global_256(hrow + 16, ph[8], ph[9], ph[10], ph[11], ph[12], ph[13], ph[14], ph[15]);
              st_global_256(lrow, pl[0], pl[1], pl[2], pl[3], pl[4], pl[5], pl[6], pl[7]);
              st_global_256(lrow + 16, pl[8], pl[9], pl[10], pl[11], pl[12], pl[13], pl[14], pl[15]);
            }
          }
          continue;
        }
        if (pr.y) {
          uint8_t* buf = stage_acquire();
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)           // row `lane` of the piece: 128 bytes, 16-byte chunks XOR-swizzled
            *reinterpret_cast<float4*>(buf + lane * 128 + ((jj ^ (lane & 7)) << 4)) =
                make_float4(m[4 * jj], m[4 * jj + 1], m[4 * jj + 2], m[4 * jj + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&maps.y[pi], buf, col, row0);
            bulk_commit();
          }
        }
        if (pr.y_hi) {
          uint32_t ph[16], pl[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) split16x2(m[2 * e], m[2 * e + 1], ph[e], pl[e]);
          uint8_t* buf = stage_acquire();
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {         // row `lane`: 64 bytes per plane, 64-byte swizzle (chunk ^= (row >> 1) & 3)
            const uint32_t off = static_cast<uint32_t>(lane * 64 + ((cc ^ ((lane >> 1) & 3)) << 4));
            *reinterpret_cast<uint4*>(buf + off) = make_uint4(ph[4 * cc], ph[4 * cc + 1], ph[4 * cc + 2], ph[4 * cc + 3]);
            *reinterpret_cast<uint4*>(buf + 2048 + off) = make_uint4(pl[4 * cc], pl[4 * cc + 1], pl[4 * cc + 2], pl[4 * cc + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&maps.y_hi[pi], buf, col, row0);
            tma_store_2d(&maps.y_lo[pi], buf + 2048, col, row0);
            bulk_commit();
          }
        }
      }
    }
    }
    if (OUT_TMA && lane == 0) bulk_wait_all();   // the staging buffers (and this CTA) outlive every store
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == X3_ISSUER) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// Keras kernel f32 [in, out] (row stride ld) -> hi / lo fp16 images [out, in], contiguous
__global__ void __launch_bounds__(256) x3_split_kernel(const float* __restrict__ w, int64_t ld, int K, int N,
                                                       __half* __restrict__ hi, __half* __restrict__ lo) {
  const int64_t total = static_cast<int64_t>(K) * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<int64_t>(n) * K);
    split16(w[k * ld + n], hi[i], lo[i]);
  }
}

template <int K, int NT, bool TMA_IN>
void launch_x3(const X3Maps& maps, X3Params p, int N, cudaStream_t st) {
  // N tiles per CTA: as many as keep both weight images within 128 KB; ring depth: what fits beside them
  const int n_tiles = N / NT;
  const int fit = std::max(1, (128 * 1024) / ((K / 64) * NT * 128 * 2));
  p.nh = std::min(n_tiles, std::min(fit, 2));
  while (n_tiles % p.nh) --p.nh;
  p.stages = X3_MAX_STAGES;
  while (p.stages > 2 && x3_smem<K, NT>(p.nh, p.stages) > 232448) --p.stages;
  const int smem = x3_smem<K, NT>(p.nh, p.stages);
  RS_CHECK(smem <= 232448, "fp32 tensor-mode Dense: the layer does not fit the shared memory of an SM");
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(dense_x3_kernel<K, NT, TMA_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  });
  int64_t max_tiles = 0;
  for (int i = 0; i < p.n_prob; ++i) max_tiles = std::max<int64_t>(max_tiles, (p.prob[i].rows + BM - 1) / BM);
  const int gy = n_tiles / p.nh;
  const int per_prob = std::max(1, rs::sm_count() / (gy * p.n_prob));
  const int gx = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(max_tiles, per_prob))) * p.n_prob;
  rs::ProfScope prof(RS_PROF_RANK, st);
  dense_x3_kernel<K, NT, TMA_IN><<<dim3(gx, gy), X3Roles<TMA_IN>::THREADS, smem, st>>>(maps, p);
  RS_LAUNCH_CHECK();
}

template <bool TMA_IN>
void dispatch_x3(const X3Maps& maps, const X3Params& p, int K, int N, cudaStream_t st) {
  if (N >= 128) {
    switch (K) {
      case 64: launch_x3<64, 128, TMA_IN>(maps, p, N, st); break;
      case 128: launch_x3<128, 128, TMA_IN>(maps, p, N, st); break;
      default: launch_x3<256, 128, TMA_IN>(maps, p, N, st); break;
    }
  } else {
    switch (K) {
      case 64: launch_x3<64, 64, TMA_IN>(maps, p, N, st); break;
      case 128: launch_x3<128, 64, TMA_IN>(maps, p, N, st); break;
      default: launch_x3<256, 64, TMA_IN>(maps, p, N, st); break;
    }
  }
}

}  // namespace

extern "C" int rs_dense_x3_prepare(const DLTensor* kernel, DLTensor* wt_hi, DLTensor* wt_lo, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(kernel, "kernel", 2, kDLFloat, 32);
    require(wt_hi, "wt_hi", 2, kDLFloat, 16);
    require(wt_lo, "wt_lo", 2, kDLFloat, 16);
    const int64_t K = kernel->shape[0], N = kernel->shape[1];
    RS_CHECK(on_device(kernel) && on_device(wt_hi) && on_device(wt_lo), "all tensors must be CUDA tensors");
    RS_CHECK(wt_hi->shape[0] == N && wt_hi->shape[1] == K && wt_lo->shape[0] == N && wt_lo->shape[1] == K && is_compact(wt_hi) && is_compact(wt_lo),
             "wt_hi / wt_lo must be contiguous [out, in] fp16 tensors");
    if (K * N == 0) return;
    sm_count();
    const int blocks = static_cast<int>(std::min<int64_t>((K * N + 255) / 256, 1024));
    x3_split_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(data_ptr<float>(kernel), row_stride_2d(kernel, "kernel"),
                                                                          static_cast<int>(K), static_cast<int>(N),
                                                                          data_ptr<__half>(wt_hi), data_ptr<__half>(wt_lo));
    RS_LAUNCH_CHECK();
  });
}

extern "C" int rs_dense_x3_supported(int64_t in, int64_t out, int act) {
  const bool k_ok = in == 64 || in == 128 || in == 256;
  const bool n_ok = out == 64 || out == 128 || out == 256;
  return (k_ok && n_ok && (act == RS_ACT_NONE || act == RS_ACT_RELU)) ? 1 : 0;
}

extern "C" int rs_dense_forward_x3(int n_layers, const rs_x3_layer_t* layers, int act, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(layers != nullptr && (n_layers == 1 || n_layers == 2), "one or two layers per call");
    X3Params p{};
    X3Maps maps{};
    p.n_prob = n_layers;
    p.act = act;
    int64_t K = 0, N = 0;
    bool tma_in = false;
    bool any_rows = false;
    for (int i = 0; i < n_layers; ++i) {
      const rs_x3_layer_t& L = layers[i];
      X3Problem& q = p.prob[i];
      require(L.wt_hi, "wt_hi", 2, kDLFloat, 16);
      require(L.wt_lo, "wt_lo", 2, kDLFloat, 16);
      RS_CHECK(on_device(L.wt_hi) && on_device(L.wt_lo) && is_compact(L.wt_hi) && is_compact(L.wt_lo) && L.wt_lo->shape[0] == L.wt_hi->shape[0] &&
                   L.wt_lo->shape[1] == L.wt_hi->shape[1], "wt_hi / wt_lo must be contiguous CUDA [out, in] fp16 tensors (rs_dense_x3_prepare)");
      const int64_t Ki = L.wt_hi->shape[1], Ni = L.wt_hi->shape[0];
      const bool split_in = L.x_hi != nullptr;
      if (i == 0) {
        K = Ki;
        N = Ni;
        tma_in = split_in;
        RS_CHECK(rs_dense_x3_supported(K, N, act), "fp32 tensor-mode Dense supports (in, out) in {64,128,256}^2 with activation none / relu");
      }
      RS_CHECK(Ki == K && Ni == N && split_in == tma_in, "the layers of one call must have the same shape and input form");
      int64_t B = 0;
      if (split_in) {
        RS_CHECK(L.x == nullptr && L.ids == nullptr && L.x_lo != nullptr, "give either x (fp32) or x_hi + x_lo (split planes)");
        require(L.x_hi, "x_hi", 2, kDLFloat, 16);
        require(L.x_lo, "x_lo", 2, kDLFloat, 16);
        B = L.x_hi->shape[0];
        RS_CHECK(on_device(L.x_hi) && on_device(L.x_lo) && L.x_hi->shape[1] == K && L.x_lo->shape[0] == B && L.x_lo->shape[1] == K,
                 "x_hi / x_lo must be CUDA fp16 [B, in] tensors");
        const int64_t ldh = row_stride_2d(L.x_hi, "x_hi"), ldl = row_stride_2d(L.x_lo, "x_lo");
        RS_CHECK(ldh % 8 == 0 && ldl % 8 == 0, "x_hi / x_lo row strides must be multiples of 8 elements");
        if (B > 0) {
          maps.x_hi[i] = make_tmap_2d(data_ptr<void>(L.x_hi), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)B, (uint64_t)K, (uint64_t)ldh * 2, BM, 64);
          maps.x_lo[i] = make_tmap_2d(data_ptr<void>(L.x_lo), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)B, (uint64_t)K, (uint64_t)ldl * 2, BM, 64);
        }
      } else {
        require(L.x, "x", 2, kDLFloat, 32);
        RS_CHECK(on_device(L.x) && L.x->shape[1] == K, "x must be a CUDA f32 [B, in] tensor");
        B = L.x->shape[0];
        q.x = data_ptr<float>(L.x);
        q.ldx = row_stride_2d(L.x, "x");
        RS_CHECK(q.ldx % 4 == 0, "x row stride must be a multiple of 4 floats");
        require_aligned16(q.x, "x");
        if (L.ids) {
          RS_CHECK(L.ids->ndim == 1 && (is_dtype(L.ids, kDLInt, 32) || is_dtype(L.ids, kDLInt, 64)) && on_device(L.ids) &&
                       (L.ids->shape[0] <= 1 || stride_of(L.ids, 0) == 1), "ids must be a contiguous CUDA 1-D int32 / int64 tensor");
          q.ids = data_ptr<void>(L.ids);
          q.ids64 = is_dtype(L.ids, kDLInt, 64) ? 1 : 0;
          q.table_rows = L.x->shape[0];
          RS_CHECK(q.table_rows <= 0x7FFFFFFF, "embedding tables above 2^31 rows are not supported by the fused gather");
          B = L.ids->shape[0];
        }
      }
      q.rows = B;
      RS_CHECK(B <= 0x7FFFFFFF, "too many rows");
      any_rows = any_rows || B > 0;
      RS_CHECK(L.y != nullptr || (L.y_hi != nullptr && L.y_lo != nullptr), "give y (fp32) and / or y_hi + y_lo (split planes)");
      if (L.y) {
        require(L.y, "y", 2, kDLFloat, 32);
        RS_CHECK(on_device(L.y) && L.y->shape[0] == B && L.y->shape[1] == N, "y must be a CUDA f32 [B, out] tensor");
        q.y = data_ptr<float>(L.y);
        q.ldy = row_stride_2d(L.y, "y");
        RS_CHECK(q.ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(q.y) & 31) == 0, "y must be 32-byte aligned with a row stride that is a multiple of 8");
        if (B > 0) maps.y[i] = make_tmap_2d(q.y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)B, (uint64_t)N, (uint64_t)q.ldy * 4, 32, 32);
      }
      if (L.y_hi) {
        require(L.y_hi, "y_hi", 2, kDLFloat, 16);
        require(L.y_lo, "y_lo", 2, kDLFloat, 16);
        RS_CHECK(on_device(L.y_hi) && on_device(L.y_lo) && L.y_hi->shape[0] == B && L.y_hi->shape[1] == N && L.y_lo->shape[0] == B &&
                     L.y_lo->shape[1] == N, "y_hi / y_lo must be CUDA fp16 [B, out] tensors");
        q.y_hi = data_ptr<__half>(L.y_hi);
        q.y_lo = data_ptr<__half>(L.y_lo);
        q.ldy16 = row_stride_2d(L.y_hi, "y_hi");
        RS_CHECK(row_stride_2d(L.y_lo, "y_lo") == q.ldy16 && q.ldy16 % 16 == 0 && (reinterpret_cast<uintptr_t>(q.y_hi) & 31) == 0 &&
                     (reinterpret_cast<uintptr_t>(q.y_lo) & 31) == 0, "y_hi / y_lo: equal row strides (multiple of 16), 32-byte aligned");
        if (B > 0) {
          maps.y_hi[i] = make_tmap_2d_sw64(q.y_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)B, (uint64_t)N, (uint64_t)q.ldy16 * 2, 32, 32);
          maps.y_lo[i] = make_tmap_2d_sw64(q.y_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)B, (uint64_t)N, (uint64_t)q.ldy16 * 2, 32, 32);
        }
      }
      if (L.bias) {
        require(L.bias, "bias", 1, kDLFloat, 32);
        RS_CHECK(on_device(L.bias) && L.bias->shape[0] == N && (N <= 1 || stride_of(L.bias, 0) == 1), "bias must be a contiguous CUDA [out] tensor");
        q.bias = data_ptr<float>(L.bias);
      }
      const uint32_t NT = N >= 128 ? 128 : 64;
      maps.w_hi[i] = make_tmap_2d(data_ptr<void>(L.wt_hi), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)N, (uint64_t)K, (uint64_t)K * 2, NT, 64);
      maps.w_lo[i] = make_tmap_2d(data_ptr<void>(L.wt_lo), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)N, (uint64_t)K, (uint64_t)K * 2, NT, 64);
    }
    if (!any_rows) return;
    for (int i = 0; i < n_layers; ++i)   // an empty problem next to a non-empty one: its CTAs find no tiles, but the maps must be valid
      if (p.prob[i].rows == 0) {
        maps.x_hi[i] = maps.x_lo[i] = maps.y[i] = maps.y_hi[i] = maps.y_lo[i] = maps.w_hi[i];
        p.prob[i].y = nullptr;
        p.prob[i].y_hi = p.prob[i].y_lo = nullptr;
      }
    sm_count();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (tma_in) dispatch_x3<true>(maps, p, static_cast<int>(K), static_cast<int>(N), st);
    else dispatch_x3<false>(maps, p, static_cast<int>(K), static_cast<int>(N), st);
  });
}
