// chain_tc.cu -- K1+K2 / K3 fused (bf16 tensor mode): a whole Dense chain in ONE kernel on tcgen05, with the
// activations living in tensor memory between layers.
//   tower chain  (src/models.py:76-77,85,89): ids -> Embedding gather -> Dense256 relu -> Dense128 relu ->
//                                             Dense64 relu -> Dense128        (both towers in one launch)
//   deep chain   (src/models.py:26-29,46-48): x0[B,256] -> Dense256 relu -> Dense128 relu
//
// Layout of the work, per CTA (one CTA per SM, persistent over 128-row tiles):
//   * ALL of the chain's kernels (transposed to [out][in] bf16, the K-major B operand) are TMA-loaded once and
//     stay resident in shared memory: 160 KB for a tower, 192 KB for the deep net.
//   * The A operand of every layer is read from TENSOR MEMORY (tcgen05.mma with A in TMEM): the epilogue warps
//     gather the embedding rows (or load the bf16 input rows), convert to bf16 and tcgen05.st them as layer 0's
//     A operand; after each layer they read the fp32 accumulator (tcgen05.ld), add bias, apply relu, pack to
//     bf16 and tcgen05.st the result as the next layer's A operand.  Activations never touch shared memory or
//     HBM: per sample the kernel reads the ids + one 512-byte embedding row and writes the 128-wide output.
//   * TMEM columns (512 per CTA) are re-used along the chain:  A0 [0,64|128)  acc1 [256,512)  A1 [0,128)
//     acc2 [128,256)  A2 [0,64)  acc3 [64,128)  A3 [128,160)  acc4 [256,384).
//   * warp 0: TMA (weights), warp 1: the single tcgen05.mma issuer, warps 4-19: epilogue (TMEM lane quarter =
//     warp % 4, column slice = (warp-4)/4).  Two mbarriers alternate per layer: a_ready (epilogue -> MMA, the A
//     operand is in TMEM) and acc_ready (tcgen05.commit -> epilogue).
// Tensor work per sample: 81,920 MAC per tower, 98,304 MAC for the deep net.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"

namespace {

using namespace rsd;

constexpr int BM = 128;
constexpr int EPI_WARPS = 16;
constexpr int NTHREADS = 128 + EPI_WARPS * 32;
constexpr int MAX_L = 4;

// a_ready / acc_ready hand-offs: the chain is strictly serial per tile, so every wait blocks
#ifndef RS_CHAIN_SPIN
#define RS_CHAIN_SPIN 0   // measured: no difference on B200 (the hand-off cost is not wake-up latency)
#endif
__device__ __forceinline__ void chain_wait(uint64_t* bar, uint32_t parity) {
#if RS_CHAIN_SPIN
  mbar_wait_spin(bar, parity);
#else
  mbar_wait(bar, parity);
#endif
}

struct CfgTower {
  static constexpr int L = 4;
  static constexpr bool GATHER = true;
  __host__ __device__ static constexpr int K(int l) { return l == 0 ? 128 : l == 1 ? 256 : l == 2 ? 128 : 64; }
  __host__ __device__ static constexpr int N(int l) { return l == 0 ? 256 : l == 1 ? 128 : l == 2 ? 64 : 128; }
  __host__ __device__ static constexpr int ACOL(int l) { return l == 3 ? 128 : 0; }
  __host__ __device__ static constexpr int DCOL(int l) { return l == 0 ? 256 : l == 1 ? 128 : l == 2 ? 64 : 256; }
};
struct CfgDeep {
  static constexpr int L = 2;
  static constexpr bool GATHER = false;
  __host__ __device__ static constexpr int K(int l) { return 256; }
  __host__ __device__ static constexpr int N(int l) { return l == 0 ? 256 : 128; }
  __host__ __device__ static constexpr int ACOL(int l) { return 0; }
  __host__ __device__ static constexpr int DCOL(int l) { return l == 0 ? 256 : 128; }
};

template <class C>
__host__ __device__ constexpr int w_off(int l) {  // byte offset of layer l's kernel in shared memory
  int o = 0;
  for (int i = 0; i < l; ++i) o += C::N(i) * C::K(i) * 2;
  return o;
}
template <class C>
__host__ __device__ constexpr int b_off(int l) {  // float offset of layer l's bias
  int o = 0;
  for (int i = 0; i < l; ++i) o += C::N(i);
  return o;
}
template <class C>
constexpr int chain_smem() {
  return w_off<C>(C::L) + b_off<C>(C::L) * 4 + 256 /*barriers*/ + 1024 /*align*/;
}

struct ChainMaps {
  CUtensorMap w[2][MAX_L];   // [problem][layer]
  CUtensorMap y[2];          // tower_pp_kernel, ChainProblem::y_tma: the fp32 output, box 16 rows x 32 floats, 128B swizzle
};

struct ChainProblem {
  int64_t rows;
  const float* bias[MAX_L];
  const float* table;    // GATHER: fp32 [V, 128]
  const uint16_t* table16;   // tower_pp_kernel: bf16 copy of `table` ([V, 128] packed) or nullptr
  int64_t table_ld;
  int64_t table_rows;
  const void* ids;       // GATHER: int32 / int64 [rows]
  int ids64;
  const uint16_t* x;     // !GATHER: bf16 [rows, 256]
  int64_t x_ld;
  float* y_f32;          // [rows, N_last] or nullptr
  int64_t ld_f32;
  uint16_t* y_bf16;      // [rows, N_last] or nullptr
  int64_t ld_bf16;
  // tower_pp_kernel only, for the fused ranker (all nullptr otherwise):
  //  * y_il16: bf16 copy of the output in the lane-interleaved layout deep_heads_kernel loads its A operand from.
  //    16-byte chunk c (8 features) of row (tile T, quarter q, lane) is uint4 number ((T*4 + q)*32 + c)*32 + lane, so
  //    one 128-bit access per lane is a contiguous 512 bytes per warp on both the writing and the reading side.
  uint16_t* y_il16;
  int il_chunk0;         // first 16-byte chunk of this problem's columns (0 for the user tower, 16 for the item tower)
  int y_tma;             // ChainMaps::y[problem] describes y_f32: the kernel may store it with TMA
};

struct ChainParams {
  ChainProblem prob[2];
  int n_prob;            // 1 or 2 (both towers in one launch)
  int act[MAX_L];
  int debug;             // perf experiments: 1 = epilogue skips TMEM traffic, 2 = issuer skips the MMAs
  int lag;               // tower_pp_kernel: MMA steps of slot 0 issued before slot 1 may start (RS_TOWER_LAG, default 3)
  long long* trace;      // RS_CHAIN_TRACE=1 (tower_pp_kernel): clock64 stamps of block 0, slot 0, warp 4 ([round][16])
  const void* zero_row;  // tower_sp_kernel: 512 zero bytes in global memory (what an invalid row id gathers)
  // tower_sp_kernel: the biases BY VALUE ([problem][b_off(layer) + column]).  Kernel parameters live in the constant bank,
  // so the epilogue reads them with LDC (same address in every lane) instead of LDS: the shared-memory / L1 pipe is the
  // one the gather warps saturate with their thread-per-row loads (32 lines per instruction), and bias loads queued
  // behind them made an epilogue step take 2 k cycles instead of 0.6 k (RS_CHAIN_TRACE)
  float biasv[2][576];
};

template <class C>
__global__ void __launch_bounds__(NTHREADS, 1)
chain_tc_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = base;
  float* bias_s = reinterpret_cast<float*>(base + w_off<C>(C::L));
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + b_off<C>(C::L));
  uint64_t* wfull = bars;
  uint64_t* a_ready = bars + 1;
  uint64_t* acc_ready = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pi = blockIdx.x % p.n_prob;                 // which problem (tower) this CTA works on
  const int cta = blockIdx.x / p.n_prob;
  const int ncta = (gridDim.x - pi + p.n_prob - 1) / p.n_prob;
  const ChainProblem& pr = p.prob[pi];
  const int64_t ntiles = (pr.rows + BM - 1) / BM;
  const int64_t n_my = cta < ntiles ? (ntiles - cta + ncta - 1) / ncta : 0;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < C::L; ++l) tma_prefetch_desc(&maps.w[pi][l]);
    mbar_init(wfull, 1);
    mbar_init(a_ready, EPI_WARPS);
    mbar_init(acc_ready, 1);
    fence_mbar_init();
  }
#pragma unroll
  for (int l = 0; l < C::L; ++l)
    for (int i = threadIdx.x; i < C::N(l); i += NTHREADS) bias_s[b_off<C>(l) + i] = pr.bias[l] ? pr.bias[l][i] : 0.0f;
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
      mbar_arrive_expect_tx(wfull, w_off<C>(C::L));
#pragma unroll
      for (int l = 0; l < C::L; ++l)
        for (int kc = 0; kc < C::K(l) / 64; ++kc)
          tma_load_2d(wsm + w_off<C>(l) + kc * (C::N(l) * 128), &maps.w[pi][l], wfull, kc * 64, 0, kEvictLast);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer: one layer at a time, A from TMEM, B from the resident weights ----------------
      mbar_wait(wfull, 0);
      // one descriptor for the start of the weight region; every operand descriptor is this plus a compile-time
      // offset (the single issuing thread must not spend its time rebuilding descriptors)
      const uint64_t wdesc0 = umma_desc_k_sw128(smem_u32(wsm));
      uint32_t use = 0;
      for (int64_t i = 0; i < n_my; ++i) {
#pragma unroll
        for (int l = 0; l < C::L; ++l, ++use) {
          chain_wait(a_ready, use & 1u);
          tc_fence_after_sync();
          const uint32_t idesc = umma_idesc_bf16_f32(BM, C::N(l));
#pragma unroll
          for (int k = 0; k < C::K(l) / 16; ++k) {
            if (p.debug & 2) break;
            const uint64_t db = wdesc0 + static_cast<uint64_t>((w_off<C>(l) + (k >> 2) * (C::N(l) * 128) + (k & 3) * 32) >> 4);
            umma_bf16_ts(tmem_base + C::DCOL(l), tmem_base + C::ACOL(l) + k * 8, db, idesc, k > 0 ? 1u : 0u);
          }
          umma_commit(acc_ready);
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue warps: produce A0, then bias/activation/repack between layers ----------------
    const int quarter = warp & 3;
    const int slice = (warp - 4) >> 2;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t use = 0;
    // layer-0 A operand of one tile: this thread's row, this warp's feature slice, as packed bf16 pairs.  The
    // loads for tile i+1 are issued right after tile i's operand is stored, so their latency (id -> row, two
    // dependent global round trips) hides behind tile i's layers instead of sitting on the critical path.
    uint32_t a0[C::GATHER ? 16 : 32];
    auto load_a0 = [&](int64_t tile_) {
      const int64_t row_ = tile_ * BM + quarter * 32 + lane;
#pragma unroll
      for (int e = 0; e < (C::GATHER ? 16 : 32); ++e) a0[e] = 0u;
      if (row_ >= pr.rows) return;
      if constexpr (C::GATHER) {
        const int64_t id = !pr.ids ? row_   // no ids: the "table" already holds this batch's rows in order
                           : pr.ids64 ? static_cast<const int64_t*>(pr.ids)[row_]
                                      : static_cast<int64_t>(static_cast<const int32_t*>(pr.ids)[row_]);
        if (id >= 0 && id < pr.table_rows) {
          const float4* src = reinterpret_cast<const float4*>(pr.table + id * pr.table_ld + slice * 32);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 f = __ldg(src + e);
            a0[2 * e] = pack_bf16x2(f.x, f.y);
            a0[2 * e + 1] = pack_bf16x2(f.z, f.w);
          }
        }
      } else {
        const uint4* src = reinterpret_cast<const uint4*>(pr.x + row_ * pr.x_ld + slice * 64);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint4 a = __ldg(src + e);
          a0[4 * e] = a.x; a0[4 * e + 1] = a.y; a0[4 * e + 2] = a.z; a0[4 * e + 3] = a.w;
        }
      }
    };
    if (n_my > 0) load_a0(cta);
    for (int64_t i = 0; i < n_my; ++i) {
      const int64_t tile = cta + i * ncta;
      const int64_t row = tile * BM + quarter * 32 + lane;
      __syncwarp();
      if constexpr (C::GATHER) {
        uint32_t (&r)[16] = reinterpret_cast<uint32_t (&)[16]>(a0);
        tmem_st_32x32_x16(tmem_base + C::ACOL(0) + slice * 16 + lane_sel, r);
      } else {
        uint32_t (&r0)[16] = reinterpret_cast<uint32_t (&)[16]>(a0[0]);
        uint32_t (&r1)[16] = reinterpret_cast<uint32_t (&)[16]>(a0[16]);
        tmem_st_32x32_x16(tmem_base + C::ACOL(0) + slice * 32 + lane_sel, r0);
        tmem_st_32x32_x16(tmem_base + C::ACOL(0) + slice * 32 + 16 + lane_sel, r1);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      if (i + 1 < n_my) load_a0(tile + ncta);   // prefetch the next tile's operand

#pragma unroll
      for (int l = 0; l < C::L; ++l, ++use) {
        const int N = C::N(l);
        const int slices = N >= 128 ? 4 : 2;
        const int cols = N / slices;                       // accumulator columns per warp: 64, 32 or 32
        chain_wait(acc_ready, use & 1u);
        tc_fence_after_sync();
        if (slice < slices && !(p.debug & 1)) {
#pragma unroll 1
          for (int c = 0; c < cols / 32; ++c) {
            const int col0 = slice * cols + c * 32;
            float v[32];
            __syncwarp();
            if (!(p.debug & 8)) tmem_ld_32x32(tmem_base + C::DCOL(l) + col0 + lane_sel, v);
            else {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = static_cast<float>(e + lane);
            }
            const float* bs = bias_s + b_off<C>(l) + col0;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
              v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
            }
            if (p.act[l] == RS_ACT_RELU) {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.0f);
            }
            if (l + 1 < C::L) {
              uint32_t r[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
              if (!(p.debug & 4)) tmem_st_32x32_x16(tmem_base + C::ACOL(l + 1 < C::L ? l + 1 : l) + (col0 >> 1) + lane_sel, r);
            } else if (row < pr.rows) {
              if (pr.y_f32) {
                float* dst = pr.y_f32 + row * pr.ld_f32 + col0;
#pragma unroll
                for (int e = 0; e < 32; e += 8)
                  st_global_256(dst + e, v[e], v[e + 1], v[e + 2], v[e + 3], v[e + 4], v[e + 5], v[e + 6], v[e + 7]);
              }
              if (pr.y_bf16) {
                uint16_t* dst = pr.y_bf16 + row * pr.ld_bf16 + col0;
#pragma unroll
                for (int e = 0; e < 32; e += 16)
                  st_global_256(static_cast<void*>(dst + e), pack_bf16x2(v[e], v[e + 1]), pack_bf16x2(v[e + 2], v[e + 3]),
                                pack_bf16x2(v[e + 4], v[e + 5]), pack_bf16x2(v[e + 6], v[e + 7]),
                                pack_bf16x2(v[e + 8], v[e + 9]), pack_bf16x2(v[e + 10], v[e + 11]),
                                pack_bf16x2(v[e + 12], v[e + 13]), pack_bf16x2(v[e + 14], v[e + 15]));
              }
            }
          }
        }
        if (l + 1 < C::L) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_ready);
        }
      }
      // the next tile's A0 overwrites TMEM columns this tile's last accumulator does not use, and the last
      // accumulator is not overwritten before the next tile's layer L (after further a_ready/acc_ready rounds)
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
// tower_pp_kernel: the tower chain with TWO 128-row tiles in flight per CTA ("ping-pong").  A single tile's chain
// is a strict sequence MMA -> epilogue -> MMA ..., so with one tile the tensor pipe idles during every epilogue.
// Here each tile owns one half of TMEM (256 columns) and 8 of the 16 epilogue warps; the one MMA issuer alternates
// between the two tiles, so tile X's epilogue overlaps tile Y's MMAs.  To fit a tile into 256 columns layer 1
// (N = 256) runs as two N = 128 halves that share one accumulator:
//   columns relative to the tile's base R:  A0 [0,64)  acc [64,192)  A1a [192,256)  A1b [0,64)  A2 [192,256)
//                                           acc3 [0,64)  A3 [64,96)  acc4 [128,256)
//   steps: L1a (A0 x W1[0:128])  L1b (A0 x W1[128:256])  L2 (A1a|A1b x W2)  L3 (A2 x W3)  L4 (A3 x W4)
//
// Data movement, MODE (RS_TOWER_STAGE=0/1/2 for A/B runs, default 2).  A tcgen05 epilogue thread owns one ROW, so a
// gather load or an output store issued per thread touches 32 different 128-byte lines per instruction.
//   0  direct: thread-per-row loads and 256-bit stores.
//   1  staged: gather and fp32 output go through a per-warp shared-memory buffer (16-byte chunks XOR-swizzled by row,
//      every access conflict-free); the global side is issued 8 lanes per row, 4 lines per instruction.  The gather
//      keeps its "in flight for a whole tile" property: the raw chunks land in registers one tile ahead (ids two tiles
//      ahead) and are only transposed when the tile starts.  Measured: same speed as 0 -- the output step is bound
//      by the SM's global-store rate (~32 B/clk), not by the number of lines per instruction.
//   2  staged + TMA stores: the outputs are written to the stage buffers only and leave through the async proxy
//      (cp.async.bulk[.tensor] shared -> global, bulk groups), two 2 KB buffers per warp in alternation, so the
//      epilogue warps go back to the chain while the stores drain.  Uses all 227 KB of shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int TOWER_STAGE_BYTES = 2048;   // per epilogue warp and buffer (MODE 1: one buffer, MODE 2: two)
// warp 0: TMA, warp 1: MMA issuer, warps 4-19: epilogue (register allocation is per 4 warps, so dropping the two idle
// warps would not raise the 96-register cap; setmaxnreg made ptxas spill more, not less)
constexpr int PP_THREADS = NTHREADS;
constexpr int PP_EPI0 = 4;

constexpr int TOWER_FIXED_SMEM = w_off<CfgTower>(CfgTower::L) + b_off<CfgTower>(CfgTower::L) * 4 + 256;   // kernels, biases, barriers
constexpr int TOWER_STAGE_OFF = (TOWER_FIXED_SMEM + 1023) / 1024 * 1024;
constexpr int tower_pp_smem(int mode) {   // modes 0/1 carry 1 KB of alignment slack; mode 2 relies on __align__(1024)
  return mode == 0 ? TOWER_FIXED_SMEM + 1024
         : mode == 1 ? TOWER_STAGE_OFF + EPI_WARPS * TOWER_STAGE_BYTES + 1024
                     : TOWER_STAGE_OFF + EPI_WARPS * 2 * TOWER_STAGE_BYTES;
}
static_assert(tower_pp_smem(2) <= 232448, "tower_pp_kernel mode 2 must fit the 227 KB of shared memory per CTA");

template <int MODE>
__global__ void __launch_bounds__(PP_THREADS, 1)
tower_pp_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  using C = CfgTower;
  constexpr bool STAGED = MODE >= 1;
  constexpr bool TMAST = MODE == 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (TMAST && base != smem_raw) __trap();   // no room for slack in mode 2
  uint8_t* wsm = base;
  float* bias_s = reinterpret_cast<float*>(base + w_off<C>(C::L));
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + b_off<C>(C::L));
  uint64_t* wfull = bars;
  uint64_t* a_ready = bars + 1;      // [2]
  uint64_t* acc_ready = bars + 3;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  uint8_t* stage = base + TOWER_STAGE_OFF;   // STAGED: [EPI_WARPS][1 or 2][TOWER_STAGE_BYTES], 1024-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pi = blockIdx.x % p.n_prob;
  const int cta = blockIdx.x / p.n_prob;
  const int ncta = (gridDim.x - pi + p.n_prob - 1) / p.n_prob;
  const ChainProblem& pr = p.prob[pi];
  const int64_t ntiles = (pr.rows + BM - 1) / BM;
  const int64_t n_my = cta < ntiles ? (ntiles - cta + ncta - 1) / ncta : 0;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < C::L; ++l) tma_prefetch_desc(&maps.w[pi][l]);
    mbar_init(wfull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_ready[s], EPI_WARPS / 2);
      mbar_init(&acc_ready[s], 1);
    }
    fence_mbar_init();
  }
#pragma unroll
  for (int l = 0; l < C::L; ++l)
    for (int i = threadIdx.x; i < C::N(l); i += PP_THREADS) bias_s[b_off<C>(l) + i] = pr.bias[l] ? pr.bias[l][i] : 0.0f;
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
      mbar_arrive_expect_tx(wfull, w_off<C>(C::L));
#pragma unroll
      for (int l = 0; l < C::L; ++l)
        for (int kc = 0; kc < C::K(l) / 64; ++kc)
          tma_load_2d(wsm + w_off<C>(l) + kc * (C::N(l) * 128), &maps.w[pi][l], wfull, kc * 64, 0, kEvictLast);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer: step s of tile slot 0, step s of tile slot 1, step s+1 of slot 0, ... ----------------
      mbar_wait(wfull, 0);
      const uint64_t wdesc0 = umma_desc_k_sw128(smem_u32(wsm));
      // Ready-driven: whichever slot has its A operand in TMEM gets its next step issued.  Slot 1 starts only after
      // slot 0 has had `p.lag` steps issued, which puts the two tiles half a period apart for the rest of the kernel:
      // one slot's output step (bound by the SM's global-store rate) then overlaps the other slot's MMA / TMEM steps
      // instead of both bursts sharing the store path at the same time.
      int64_t left[2] = {(n_my + 1) / 2, n_my / 2};
      uint32_t use[2] = {0, 0};
      int step_of[2] = {0, 0};
      int lag = left[1] > 0 ? p.lag : 0;
      auto issue = [&](int slot, int step) {
        const uint32_t R = tmem_base + slot * 256;
        if (step <= 1) {          // L1 half: A0 [R, R+64) x W1 rows [128*step, +128), K = 128
          const uint32_t idesc = umma_idesc_bf16_f32(BM, 128);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t db = wdesc0 + static_cast<uint64_t>((w_off<C>(0) + (k >> 2) * (256 * 128) + step * (128 * 128) + (k & 3) * 32) >> 4);
            umma_bf16_ts(R + 64, R + k * 8, db, idesc, k > 0 ? 1u : 0u);
          }
        } else if (step == 2) {   // L2: A1a [R+192, +64) then A1b [R, +64), W2 (N = 128), K = 256
          const uint32_t idesc = umma_idesc_bf16_f32(BM, 128);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const uint64_t db = wdesc0 + static_cast<uint64_t>((w_off<C>(1) + (k >> 2) * (128 * 128) + (k & 3) * 32) >> 4);
            umma_bf16_ts(R + 64, k < 8 ? R + 192 + k * 8 : R + (k - 8) * 8, db, idesc, k > 0 ? 1u : 0u);
          }
        } else if (step == 3) {   // L3: A2 [R+192, +64), W3 (N = 64), K = 128 -> acc3 [R, +64)
          const uint32_t idesc = umma_idesc_bf16_f32(BM, 64);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t db = wdesc0 + static_cast<uint64_t>((w_off<C>(2) + (k >> 2) * (64 * 128) + (k & 3) * 32) >> 4);
            umma_bf16_ts(R, R + 192 + k * 8, db, idesc, k > 0 ? 1u : 0u);
          }
        } else {                  // L4: A3 [R+64, +32), W4 (N = 128), K = 64 -> acc4 [R+128, +128)
          const uint32_t idesc = umma_idesc_bf16_f32(BM, 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t db = wdesc0 + static_cast<uint64_t>((w_off<C>(3) + (k & 3) * 32) >> 4);
            umma_bf16_ts(R + 128, R + 64 + k * 8, db, idesc, k > 0 ? 1u : 0u);
          }
        }
      };
      uint32_t idle = 0;
      while (left[0] > 0 || left[1] > 0) {
        if (++idle > (1u << 27)) {   // seconds without progress: fail loudly instead of hanging the device
          printf("recsys_b200: tower_pp_kernel MMA issuer made no progress (block %d)\n", blockIdx.x);
          __trap();
        }
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
          if (left[slot] == 0) continue;
          if (slot == 1 && lag > 0 && left[0] > 0) continue;
          if (!mbar_test_wait(&a_ready[slot], use[slot] & 1u)) continue;
          idle = 0;
          tc_fence_after_sync();
          issue(slot, step_of[slot]);
          umma_commit(&acc_ready[slot]);
          ++use[slot];
          if (slot == 0 && lag > 0) --lag;
          if (++step_of[slot] == 5) {
            step_of[slot] = 0;
            --left[slot];
          }
        }
      }
    }
  } else if (warp >= PP_EPI0) {
    // ---------------- epilogue: 8 warps per tile slot ----------------
    const int ew = warp - PP_EPI0;
    const int slot = ew >> 3;
    const int quarter = warp & 3;
    const int half = (ew >> 2) & 1;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t R = tmem_base + slot * 256 + lane_sel;
    const int64_t n_slot = slot == 0 ? (n_my + 1) / 2 : n_my / 2;
    uint32_t use = 0;
    // A0 of one tile: features [64*half, +64) of this thread's embedding row -> 32 packed columns; prefetched one
    // tile ahead so the id -> row global round trips overlap the previous tile's layers
    uint32_t a0[32];
    auto load_a0 = [&](int64_t tile_) {
      const int64_t row_ = tile_ * BM + quarter * 32 + lane;
#pragma unroll
      for (int e = 0; e < 32; ++e) a0[e] = 0u;
      if (row_ >= pr.rows) return;
      const int64_t id = !pr.ids ? row_
                         : pr.ids64 ? static_cast<const int64_t*>(pr.ids)[row_]
                                    : static_cast<int64_t>(static_cast<const int32_t*>(pr.ids)[row_]);
      if (id >= 0 && id < pr.table_rows) {
        if (pr.table16) {
          // bf16 copy of the table: the eight loads land in the very registers tcgen05.st takes next round, so
          // nothing waits for them here and they stay in flight behind this tile's layers
          const uint4* src = reinterpret_cast<const uint4*>(pr.table16 + id * 128 + half * 64);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint4 u = __ldg(src + e);
            a0[4 * e] = u.x; a0[4 * e + 1] = u.y; a0[4 * e + 2] = u.z; a0[4 * e + 3] = u.w;
          }
        } else {
          const float4* src = reinterpret_cast<const float4*>(pr.table + id * pr.table_ld + half * 64);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float4 f = __ldg(src + e);
            a0[2 * e] = pack_bf16x2(f.x, f.y);
            a0[2 * e + 1] = pack_bf16x2(f.z, f.w);
          }
        }
      }
    };
    // STAGED gather (bf16 table only): a0[4e..4e+3] = 16-byte chunk (lane & 7) of the half-row of tile row
    // 4e + (lane >> 3) -- eight lanes read one 128-byte half-row, so a load instruction touches 4 lines.  The row ids
    // are loaded one per lane (coalesced) a tile earlier still and handed round by shuffles.
    const bool staged_gather = STAGED && pr.table16 != nullptr;
    uint8_t* const stg0 = stage + ew * (TMAST ? 2 : 1) * TOWER_STAGE_BYTES;
    // MODE 2: lane 0 issues every bulk store of the warp, one bulk group each; store number n reads buffer n & 1, so a
    // buffer may be refilled once at most one group (the other buffer's) is still reading
    uint32_t nst = 0;
    auto stage_acquire = [&]() -> uint8_t* {
      if (TMAST) {
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        return stg0 + (nst & 1u) * TOWER_STAGE_BYTES;
      }
      __syncwarp();
      return stg0;
    };
    const int l8 = lane & 7, l3 = lane >> 3;
    auto load_id = [&](int64_t tile_) -> int64_t {
      const int64_t row_ = tile_ * BM + quarter * 32 + lane;
      if (row_ >= pr.rows) return -1;
      return !pr.ids ? row_
             : pr.ids64 ? static_cast<const int64_t*>(pr.ids)[row_]
                        : static_cast<int64_t>(static_cast<const int32_t*>(pr.ids)[row_]);
    };
    // The loads are UNCONDITIONAL (an invalid id reads row 0) and land directly in a0; which chunks are real is kept
    // in vmask and applied when the tile starts.  (Written as "u = 0; if (ok) u = load" the compiler loads into a
    // temporary and moves it under the predicate right away, a use of the loaded value.)  Issuing the eight loads
    // still costs the warp 2-3 k cycles per tile in every form measured -- conditional or not, all at once or two
    // per step, thread-per-row or 8 lanes per row: 32 KB of L2 reads per slot and tile against what one SM keeps
    // in flight with all of its shared memory carved out.
    uint32_t vmask = 0;
    auto load_a0_staged = [&](int64_t my_id) {
      vmask = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t id = __shfl_sync(0xffffffffu, my_id, 4 * e + l3);
        const bool ok = id >= 0 && id < pr.table_rows;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(pr.table16 + (ok ? id : 0) * 128 + half * 64) + l8);
        a0[4 * e] = u.x; a0[4 * e + 1] = u.y; a0[4 * e + 2] = u.z; a0[4 * e + 3] = u.w;
        vmask |= (ok ? 1u : 0u) << e;
      }
    };
    int64_t id_next = -1;   // STAGED: this lane's row id for the tile after the one whose rows are in flight
    if (n_slot > 0) {
      if (staged_gather) {
        load_a0_staged(load_id(cta + slot * ncta));
        if (n_slot > 1) id_next = load_id(cta + (2 + slot) * ncta);
      } else {
        load_a0(cta + slot * ncta);
      }
    }
    for (int64_t r = 0; r < n_slot; ++r) {
      const int64_t tile = cta + (2 * r + slot) * ncta;
      const int64_t row = tile * BM + quarter * 32 + lane;
      const bool tr = p.trace && blockIdx.x == 0 && r < 8 && (ew & 7) == 0 && lane == 0;
      auto stamp = [&](int k) { if (tr) p.trace[(slot * 8 + r) * 16 + k] = clock64(); };
      stamp(0);
      if (staged_gather) {
        // transpose through the stage buffer, 32 features (four 16-byte chunks) of all 32 rows at a time: chunk c of
        // row rr sits at rr*64 + ((c ^ ((rr >> 1) & 3)) << 4) -- conflict-free for the 8-lane phases of both sides
        uint8_t* const stg = stage_acquire();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t t[16];
          __syncwarp();
          if ((l8 >> 2) == h) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int rr = 4 * e + l3;
              const bool ok = (vmask >> e) & 1u;
              *reinterpret_cast<uint4*>(stg + rr * 64 + (((l8 & 3) ^ ((rr >> 1) & 3)) << 4)) =
                  ok ? make_uint4(a0[4 * e], a0[4 * e + 1], a0[4 * e + 2], a0[4 * e + 3]) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 u = *reinterpret_cast<const uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
            t[4 * j] = u.x; t[4 * j + 1] = u.y; t[4 * j + 2] = u.z; t[4 * j + 3] = u.w;
          }
          tmem_st_32x32_x16(R + half * 32 + h * 16, t);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[slot]);
        stamp(1);
        if (r + 1 < n_slot) {
          load_a0_staged(id_next);
          id_next = r + 2 < n_slot ? load_id(tile + 4 * ncta) : -1;
        }
        stamp(2);
      } else {
        uint32_t (&r0)[16] = reinterpret_cast<uint32_t (&)[16]>(a0[0]);
        uint32_t (&r1)[16] = reinterpret_cast<uint32_t (&)[16]>(a0[16]);
        __syncwarp();
        tmem_st_32x32_x16(R + half * 32, r0);
        tmem_st_32x32_x16(R + half * 32 + 16, r1);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[slot]);
        stamp(1);
        if (r + 1 < n_slot) load_a0(tile + 2 * ncta);
        stamp(2);
      }
#pragma unroll
      for (int step = 0; step < 5; ++step, ++use) {
        // accumulator columns, bias, destination of the repacked activations for this step
        const int acc_col = step <= 2 ? 64 : step == 3 ? 0 : 128;
        const int ncols = step == 3 ? 64 : 128;                 // accumulator width
        const int per_warp = ncols / 2;                          // 64 or 32 columns per warp
        const int layer = step <= 1 ? 0 : step - 1;
        const int bias0 = b_off<C>(layer) + (step == 1 ? 128 : 0);
        const int dst_col = step == 0 ? 192 : step == 1 ? 0 : step == 2 ? 192 : 64;   // A1a, A1b, A2, A3
        chain_wait(&acc_ready[slot], use & 1u);
        tc_fence_after_sync();
        stamp(3 + 2 * step);
#pragma unroll 1
        for (int c = 0; c < per_warp / 32; ++c) {
          const int col0 = half * per_warp + c * 32;
          float v[32];
          __syncwarp();
          tmem_ld_32x32(R + acc_col + col0, v);
          const float* bs = bias_s + bias0 + col0;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
            v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
          }
          if (step < 4) {
            uint32_t rr[16];
            if (p.act[layer] == RS_ACT_RELU) {
#pragma unroll
              for (int e = 0; e < 16; ++e) rr[e] = pack_bf16x2_relu(v[2 * e], v[2 * e + 1]);
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) rr[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
            }
            tmem_st_32x32_x16(R + dst_col + (col0 >> 1), rr);
          } else {
            if (p.act[layer] == RS_ACT_RELU) {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.0f);
            }
            if (pr.y_il16 && !(p.debug & 64)) {   // rows past the end of the batch land in the scratch's padding
              uint4* gdst = reinterpret_cast<uint4*>(pr.y_il16) + ((tile * 4 + quarter) * 32 + pr.il_chunk0 + (col0 >> 3)) * 32;
              uint4* dst = gdst + lane;
              uint8_t* buf = nullptr;
              if (TMAST) {   // this warp's four 512-byte chunks are 2 KB contiguous: one bulk store
                buf = stage_acquire();
                dst = reinterpret_cast<uint4*>(buf) + lane;
              }
#pragma unroll
              for (int e = 0; e < 4; ++e)
                dst[e * 32] = make_uint4(pack_bf16x2(v[8 * e], v[8 * e + 1]), pack_bf16x2(v[8 * e + 2], v[8 * e + 3]),
                                         pack_bf16x2(v[8 * e + 4], v[8 * e + 5]), pack_bf16x2(v[8 * e + 6], v[8 * e + 7]));
              if (TMAST) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  bulk_store_1d(gdst, buf, 2048);
                  bulk_commit();
                }
                ++nst;
              }
            }
            if (TMAST && pr.y_f32 && pr.y_tma && !(p.debug & 16)) {
              // fp32 output: 16 rows x 128 B per buffer in the 128B-swizzle layout, one TMA store each (rows past the
              // end of the batch are clipped by the tensor map)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint8_t* buf = stage_acquire();
                if ((lane >> 4) == h) {
                  const int rr = lane & 15;
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(buf + rr * 128 + ((j ^ (rr & 7)) << 4)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&maps.y[pi], buf, col0, static_cast<int>(tile * BM + quarter * 32 + h * 16));
                  bulk_commit();
                }
                ++nst;
              }
            } else if (STAGED && pr.y_f32 && !(p.debug & 16)) {
              // fp32 output through the stage buffer, 16 rows at a time; stored 8 lanes per 128-byte row segment
              if (TMAST) {
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
              }
              uint8_t* const stg = stg0;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                __syncwarp();
                if ((lane >> 4) == h) {
                  const int rr = lane & 15;
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stg + rr * 128 + ((j ^ (rr & 7)) << 4)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int rr = 4 * i + l3;
                  const int64_t grow = tile * BM + quarter * 32 + h * 16 + rr;
                  const float4 x = *reinterpret_cast<const float4*>(stg + rr * 128 + ((l8 ^ (rr & 7)) << 4));
                  if (grow < pr.rows) *reinterpret_cast<float4*>(pr.y_f32 + grow * pr.ld_f32 + col0 + 4 * l8) = x;
                }
              }
            }
            if (row < pr.rows) {
              if (!STAGED && pr.y_f32 && !(p.debug & 16)) {
                float* dst = pr.y_f32 + row * pr.ld_f32 + col0;
#pragma unroll
                for (int e = 0; e < 32; e += 8)
                  st_global_256(dst + e, v[e], v[e + 1], v[e + 2], v[e + 3], v[e + 4], v[e + 5], v[e + 6], v[e + 7]);
              }
              if (pr.y_bf16) {
                uint16_t* dst = pr.y_bf16 + row * pr.ld_bf16 + col0;
#pragma unroll
                for (int e = 0; e < 32; e += 16)
                  st_global_256(static_cast<void*>(dst + e), pack_bf16x2(v[e], v[e + 1]), pack_bf16x2(v[e + 2], v[e + 3]),
                                pack_bf16x2(v[e + 4], v[e + 5]), pack_bf16x2(v[e + 6], v[e + 7]),
                                pack_bf16x2(v[e + 8], v[e + 9]), pack_bf16x2(v[e + 10], v[e + 11]),
                                pack_bf16x2(v[e + 12], v[e + 13]), pack_bf16x2(v[e + 14], v[e + 15]));
              }
            }
          }
        }
        if (step < 4) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[slot]);
        }
        stamp(4 + 2 * step);
      }
    }
    if (TMAST && lane == 0) bulk_wait_all();   // the stage buffers (and this CTA) outlive every store
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// tower_sp_kernel ("shared ping-pong", round 2).  Same tensor-memory layout, same five MMA steps per tile and the same
// arithmetic as tower_pp_kernel (bit-identical outputs); what changed is who does what, after the round-1 traces and
// tools/ubench_tc.cu showed where a 20 k-cycle tile went (MMA 2.6 k):
//   * ALL 16 epilogue warps serve BOTH tile slots in a fixed order (slot 0 step s, slot 1 step s, slot 0 step s+1 ..).
//     A step's accumulator (128 columns) is 32 columns per warp -- one tcgen05.ld, one tcgen05.st -- so a step's
//     epilogue takes ~450 cycles instead of ~1000 (8 warps x 64 columns), and while the warps work on one slot the
//     tensor pipe runs the other slot's step: a half-step costs max(MMA, epilogue + hand-off) instead of their sum.
//   * The embedding gather has its own four warps (one per TMEM lane quarter): thread-per-row loads of the whole 256-byte
//     bf16 row, held in registers until the slot's A0 columns are free, then four tcgen05.st.  It used to cost the
//     epilogue warps ~3.9 k cycles per tile (issue + transpose) on the critical path.
//   * The MMA issuer follows the same fixed order with blocking (suspending) waits instead of polling both slots.
//   * The output step is bound by the SM's store path (32 B/clk measured, tools/ubench_tc.cu: 96 KB per tile = 3.1 k
//     cycles): it still goes through two 2 KB stage buffers per warp and TMA / bulk stores, but a warp releases the
//     accumulator columns as soon as its tcgen05.ld has landed, so the next tile's first MMA step overlaps the stores.
// Barriers per slot: a0_ready (gather -> issuer), a0_free (epilogue step 3 -> gather: columns [0,64) are dead),
// a_ready (epilogue -> issuer: next step's A operand is in TMEM / after step 4: the accumulator columns are free),
// acc_ready (tcgen05.commit -> epilogue).
// ------------------------------------------------------------------------------------------------
// warps 0-3 gather, warps 4-19 epilogue (TMEM lane quarter = warp % 4 for both), warp 20: one thread loads the kernels
// (TMA) and then issues every MMA.  21 warps = 672 threads leave 96 registers per thread: with the usual 4 + 4 + 16 warps
// (80 registers) the gather rows and a handful of loop scalars spilled, and a spill is an L2 round trip here -- the
// 227 KB of shared memory leave the SM almost no L1.
constexpr int SP_THREADS = 672;
constexpr int SP_EPI0 = 4;
constexpr int SP_EPI_WARPS = 16;
constexpr int SP_ISSUER = SP_EPI0 + SP_EPI_WARPS;

// TEAMS = false: all 16 epilogue warps serve both slots in a fixed event order (32 accumulator columns per warp and step).
// TEAMS = true: eight warps per slot (64 columns per warp and step, as two 32-column chunks) and a polling issuer: a
// team that sits in its output step -- waiting for the store path, 96 KB per tile at ~20 B/clk -- no longer holds up the
// other slot's chain steps, which is what bounded the shared order (two output steps of ~3.3 k cycles per 15 k round).
template <bool TEAMS, bool TRACE>
__global__ void __launch_bounds__(SP_THREADS, 1)
tower_sp_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  using C = CfgTower;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  if ((smem_u32(base) & 1023u) != 0) __trap();   // all 227 KB are in use: no room for alignment slack
  uint8_t* wsm = base;
  float* bias_s = reinterpret_cast<float*>(base + w_off<C>(C::L));
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + b_off<C>(C::L));
  uint64_t* wfull = bars;
  uint64_t* a0_ready = bars + 1;    // [2]
  uint64_t* a0_free = bars + 3;     // [2]
  uint64_t* a_ready = bars + 5;     // [2]
  uint64_t* acc_ready = bars + 7;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  uint8_t* stage = base + TOWER_STAGE_OFF;   // [16 epilogue warps][2][TOWER_STAGE_BYTES]

  // broadcast from lane 0: the compiler then knows the warp index is uniform and keeps everything derived from it (role
  // branches, bias addresses in the constant bank) on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int pi = blockIdx.x % p.n_prob;
  const int cta = blockIdx.x / p.n_prob;
  const int ncta = (gridDim.x - pi + p.n_prob - 1) / p.n_prob;
  const ChainProblem& pr = p.prob[pi];
  // 32-bit tile arithmetic throughout (a tile is 128 rows: 2^31 rows is far beyond any batch): every register counts,
  // the cap is 80 and a spill is an L2 round trip
  const int ntiles = static_cast<int>((pr.rows + BM - 1) / BM);
  const int n_my = cta < ntiles ? (ntiles - cta + ncta - 1) / ncta : 0;
  const int n_slot0 = (n_my + 1) / 2, n_slot1 = n_my / 2;   // tiles of slot 0 / slot 1 (tile j of this CTA -> slot j & 1)

  if (warp == SP_ISSUER && lane == 0) {
    for (int l = 0; l < C::L; ++l) tma_prefetch_desc(&maps.w[pi][l]);
    tma_prefetch_desc(&maps.y[pi]);
    mbar_init(wfull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a0_ready[s], 4);
      mbar_init(&a0_free[s], TEAMS ? SP_EPI_WARPS / 2 : SP_EPI_WARPS);
      mbar_init(&a_ready[s], TEAMS ? SP_EPI_WARPS / 2 : SP_EPI_WARPS);
      mbar_init(&acc_ready[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == SP_ISSUER) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: the deep kernel queued behind this one may be scheduled as SMs free up -- its CTAs
  // load their kernels and set up tensor memory while the last tower tiles finish elsewhere, and wait
  // (griddepcontrol.wait) before they touch anything this kernel writes
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == SP_ISSUER) {
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, w_off<C>(C::L));
#pragma unroll
      for (int l = 0; l < C::L; ++l)
        for (int kc = 0; kc < C::K(l) / 64; ++kc)
          tma_load_2d(wsm + w_off<C>(l) + kc * (C::N(l) * 128), &maps.w[pi][l], wfull, kc * 64, 0, kEvictLast);
      // ---------------- MMA issuer ----------------
      // Event order (shared with the epilogue warps): slot 0 step k of its tile r, then slot 1 step (k + 3) % 5 -- slot 1
      // runs two steps behind slot 0, so the two output steps (96 KB each into a store path that drains ~20 B/clk) are
      // half a round apart instead of back to back.  The slot is a RUN-TIME value and the steps share code where they
      // can: this kernel's first version unrolled (step x slot) and was 68 KB of SASS -- twice the 32 KB instruction
      // cache -- and the single issuing thread then paid an instruction-cache miss per MMA group (~470 cycles from an
      // epilogue's arrive to the next MMA, RS_CHAIN_TRACE).
      mbar_wait(wfull, 0);
      uint64_t wdesc0 = umma_desc_k_sw128(smem_u32(wsm));
      constexpr uint32_t idesc128 = umma_idesc_bf16_f32(BM, 128);
      constexpr uint32_t idesc64 = umma_idesc_bf16_f32(BM, 64);
      auto issue = [&](int slot, int step) {
        const uint32_t R = tmem_base + slot * 256;
        asm volatile("" : "+l"(wdesc0));   // keeps the 36 operand descriptors from being hoisted (and spilled) as loop invariants
        switch (step) {
          case 0:
          case 1: {   // L1 half: A0 [R, R+64) x W1 rows [128*step, +128), K = 128 -> acc [R+64, +128)
            const uint64_t wd = wdesc0 + static_cast<uint64_t>((w_off<C>(0) + step * (128 * 128)) >> 4);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16_ts(R + 64, R + kk * 8, wd + static_cast<uint64_t>(((kk >> 2) * (256 * 128) + (kk & 3) * 32) >> 4), idesc128,
                           kk > 0 ? 1u : 0u);
            break;
          }
          case 2: {   // L2: A1a [R+192, +64) then A1b [R, +64), W2 (N = 128), K = 256 -> acc
#pragma unroll
            for (int kk = 0; kk < 16; ++kk)
              umma_bf16_ts(R + 64, kk < 8 ? R + 192 + kk * 8 : R + (kk - 8) * 8,
                           wdesc0 + static_cast<uint64_t>((w_off<C>(1) + (kk >> 2) * (128 * 128) + (kk & 3) * 32) >> 4), idesc128,
                           kk > 0 ? 1u : 0u);
            break;
          }
          case 3: {   // L3: A2 [R+192, +64), W3 (N = 64), K = 128 -> acc3 [R, +64)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16_ts(R, R + 192 + kk * 8,
                           wdesc0 + static_cast<uint64_t>((w_off<C>(2) + (kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4), idesc64,
                           kk > 0 ? 1u : 0u);
            break;
          }
          default: {  // L4: A3 [R+64, +32), W4 (N = 128), K = 64 -> acc4 [R+128, +128)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ts(R + 128, R + 64 + kk * 8, wdesc0 + static_cast<uint64_t>((w_off<C>(3) + (kk & 3) * 32) >> 4), idesc128,
                           kk > 0 ? 1u : 0u);
            break;
          }
        }
      };
      if constexpr (TEAMS) {
        // ready-driven: whichever slot has its next operand in tensor memory gets its next step issued
        int left[2] = {n_slot0, n_slot1}, stp[2] = {0, 0}, til[2] = {0, 0};
        uint32_t use[2] = {0, 0};
        uint32_t idle = 0;
        while (left[0] > 0 || left[1] > 0) {
          if (++idle > (1u << 27)) __trap();   // seconds without progress: fail instead of hanging the device
#pragma unroll
          for (int slot = 0; slot < 2; ++slot) {
            if (left[slot] == 0) continue;
            const int step = stp[slot], t = til[slot];
            const bool need_a = step != 0 || t > 0;   // step 0 of a later tile: the previous tile's last accumulator has been read
            if (step == 0 && !mbar_test_wait(&a0_ready[slot], static_cast<uint32_t>(t & 1))) continue;
            if (need_a && !mbar_test_wait(&a_ready[slot], use[slot] & 1u)) continue;
            if (need_a) ++use[slot];
            idle = 0;
            tc_fence_after_sync();
            if (TRACE && blockIdx.x == 0 && t < 4) p.trace[80 + (t * 2 + slot) * 5 + step] = clock64();
            issue(slot, step);
            umma_commit(&acc_ready[slot]);
            if (++stp[slot] == 5) {
              stp[slot] = 0;
              ++til[slot];
              --left[slot];
            }
          }
        }
      } else {
      uint32_t use0 = 0, use1 = 0;
      for (int r = 0; r <= n_slot0; ++r) {
#pragma unroll 1
        for (int k = 0; k < 5; ++k) {
#pragma unroll 1
          for (int slot = 0; slot < 2; ++slot) {
            const int step = slot == 0 ? k : (k >= 2 ? k - 2 : k + 3);
            const int t = slot == 0 ? r : (k < 2 ? r - 1 : r);   // this slot's tile number
            if (t < 0 || t >= (slot == 0 ? n_slot0 : n_slot1)) continue;
            uint32_t& use = slot == 0 ? use0 : use1;
            if (step == 0) mbar_wait(&a0_ready[slot], static_cast<uint32_t>(t & 1));
            if (step != 0 || t > 0) {   // step 0: the previous tile's last accumulator has been read
              mbar_wait(&a_ready[slot], use & 1u);
              ++use;
            }
            tc_fence_after_sync();
            if (TRACE && blockIdx.x == 0 && t < 4) p.trace[80 + (t * 2 + slot) * 5 + step] = clock64();
            issue(slot, step);
            umma_commit(&acc_ready[slot]);
          }
        }
      }
      }
    }
  } else if (warp < SP_EPI0) {
    // ---------------- gather: one embedding row per thread -> A0 of the slot, as soon as its columns are free ----------------
    const int quarter = warp & 3;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
#pragma unroll 1
    for (int j = 0; j < n_my; ++j) {
      const int slot = j & 1;
      const int r = j >> 1;
      const int64_t row_ = static_cast<int64_t>(cta + j * ncta) * BM + quarter * 32 + lane;
      int64_t id = -1;
      if (row_ < pr.rows)
        id = !pr.ids ? row_
             : pr.ids64 ? static_cast<const int64_t*>(pr.ids)[row_]
                        : static_cast<int64_t>(static_cast<const int32_t*>(pr.ids)[row_]);
      const bool ok = id >= 0 && id < pr.table_rows;   // rows past the batch and out-of-range ids read as zeros
      const uint32_t R = tmem_base + slot * 256 + lane_sel;
      const bool tr = TRACE && blockIdx.x == 0 && j < 8 && quarter == 0 && lane == 0;
      if (tr) p.trace[120 + j * 3] = clock64();
      if (pr.table16) {
        // the whole 256-byte bf16 row: sixteen 128-bit loads in flight, untouched until the columns are free.  An invalid
        // row reads a row of zeros (p.zero_row) instead of being masked afterwards: no selects, no extra live registers.
        uint32_t a0[64];
        const uint4* src = ok ? reinterpret_cast<const uint4*>(pr.table16 + id * 128) : reinterpret_cast<const uint4*>(p.zero_row);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const uint4 u = __ldg(src + e);
          a0[4 * e] = u.x; a0[4 * e + 1] = u.y; a0[4 * e + 2] = u.z; a0[4 * e + 3] = u.w;
        }
        if (r > 0) mbar_wait(&a0_free[slot], static_cast<uint32_t>((r - 1) & 1));
        tc_fence_after_sync();
        if (tr) p.trace[120 + j * 3 + 1] = clock64();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st_32x32_x16_nowait(R + c * 16, a0 + c * 16);
        tmem_st_wait();
      } else {
        // fp32 rows (no bf16 copy of the table; looked-up rows of a sharded table): half a row at a time
        const float4* src = ok ? reinterpret_cast<const float4*>(pr.table + id * pr.table_ld) : reinterpret_cast<const float4*>(p.zero_row);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t a[32];
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            float4 f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __ldg(src + h * 16 + b * 8 + e);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              a[b * 16 + 2 * e] = pack_bf16x2(f[e].x, f[e].y);
              a[b * 16 + 2 * e + 1] = pack_bf16x2(f[e].z, f[e].w);
            }
          }
          if (h == 0) {
            if (r > 0) mbar_wait(&a0_free[slot], static_cast<uint32_t>((r - 1) & 1));
            tc_fence_after_sync();
            if (tr) p.trace[120 + j * 3 + 1] = clock64();
          }
          __syncwarp();
          tmem_st_32x32_x16_nowait(R + h * 32, a);
          tmem_st_32x32_x16_nowait(R + h * 32 + 16, a + 16);
          tmem_st_wait();
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a0_ready[slot]);
      if (tr) p.trace[120 + j * 3 + 2] = clock64();
    }
  } else {
    // ---------------- epilogue: 16 warps, both slots, 32 accumulator columns per warp and step ----------------
    const int ew = warp - SP_EPI0;
    const int quarter = warp & 3;
    const int cg = ew >> 2;   // column group: accumulator columns [32 cg, +32) (step 3: [16 cg, +16))
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    uint8_t* const stg0 = stage + ew * 2 * TOWER_STAGE_BYTES;
    uint32_t nst = 0;   // bulk store number: store n reads buffer n & 1
    auto stage_acquire = [&]() -> uint8_t* {
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      return stg0 + (nst & 1u) * TOWER_STAGE_BYTES;
    };
    const float* const bv = p.biasv[pi];   // constant bank
    // One 32-column accumulator chunk of a chain step: bias, activation, bf16 -> the next layer's A operand.
    auto chunk32 = [&](uint32_t src, const float* bs, bool relu, uint32_t dst) {
      float v[32];
      __syncwarp();
      tmem_ld_32x32(src, v);
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
        v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
      }
      uint32_t rr[16];
      if (relu) {
#pragma unroll
        for (int e = 0; e < 16; ++e) rr[e] = pack_bf16x2_relu(v[2 * e], v[2 * e + 1]);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) rr[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
      }
      tmem_st_32x32_x16(dst, rr);
    };
    auto arrive = [&](uint64_t* bar) {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // One 32-column chunk of the last accumulator: fp32 rows through the warp's two stage buffers (TMA stores, rows past
    // the batch are clipped by the tensor map) and the lane-interleaved bf16 copy for the deep kernel straight from
    // registers (coalesced 128-bit stores; a third staged piece would make the warp wait for a buffer every time).
    // `release` != nullptr: arrive there once the chunk is in registers (the accumulator columns are free again).
    auto out_chunk = [&](uint32_t R, int tile, int col0, uint64_t* release) {
      const float* bs = bv + b_off<C>(3) + col0;
      float v[32];
      __syncwarp();
      tmem_ld_32x32(R + 128 + col0, v);
      if (release) arrive(release);
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
        v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
      }
      if (p.act[3] == RS_ACT_RELU) {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.0f);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (p.debug & 16) break;   // perf experiments only: no fp32 output
        uint8_t* buf = stage_acquire();
        if ((lane >> 4) == h) {
          const int rw = lane & 15;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            *reinterpret_cast<float4*>(buf + rw * 128 + ((jj ^ (rw & 7)) << 4)) =
                make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&maps.y[pi], buf, col0, tile * BM + quarter * 32 + h * 16);
          bulk_commit();
        }
        ++nst;
      }
      if (pr.y_il16 && !(p.debug & 64)) {
        uint4* dst = reinterpret_cast<uint4*>(pr.y_il16) + ((static_cast<int64_t>(tile) * 4 + quarter) * 32 + pr.il_chunk0 + (col0 >> 3)) * 32 + lane;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          dst[e * 32] = make_uint4(pack_bf16x2(v[8 * e], v[8 * e + 1]), pack_bf16x2(v[8 * e + 2], v[8 * e + 3]),
                                   pack_bf16x2(v[8 * e + 4], v[8 * e + 5]), pack_bf16x2(v[8 * e + 6], v[8 * e + 7]));
      }
    };
    if constexpr (TEAMS) {
      // ---------------- eight warps per slot: 64 accumulator columns per warp and step ----------------
      const int slot = ew >> 3;
      const int hf = (ew >> 2) & 1;
      const uint32_t R = tmem_base + slot * 256 + lane_sel;
      const int n_slot = slot == 0 ? n_slot0 : n_slot1;
      uint32_t cnt = 0;
      for (int r = 0; r < n_slot; ++r) {
        const int tile = cta + (2 * r + slot) * ncta;
#pragma unroll 1
        for (int step = 0; step < 5; ++step) {
          const bool tr = TRACE && blockIdx.x == 0 && r < 4 && (ew & 7) == 0 && lane == 0;
          chain_wait(&acc_ready[slot], cnt & 1u);
          ++cnt;
          tc_fence_after_sync();
          if (tr) p.trace[((r * 2 + slot) * 5 + step) * 2] = clock64();
          if (step <= 2) {
            const bool relu = p.act[step == 2 ? 1 : 0] == RS_ACT_RELU;
            const int b0 = (step == 2 ? b_off<C>(1) : step * 128) + hf * 64;
            const uint32_t dst = R + (step == 1 ? 0 : 192) + hf * 32;   // A1a, A1b, A2
#pragma unroll 1
            for (int c = 0; c < 2; ++c) chunk32(R + 64 + hf * 64 + c * 32, bv + b0 + c * 32, relu, dst + c * 16);
            arrive(&a_ready[slot]);
          } else if (step == 3) {
            // acc3 [0,64): 32 columns per warp; once they are in registers columns [0,64) are dead and the gather warps may
            // write the next A0 (chunk32 arrives too late for that: split by hand)
            float v[32];
            __syncwarp();
            tmem_ld_32x32(R + hf * 32, v);
            arrive(&a0_free[slot]);
            const float* bs = bv + b_off<C>(2) + hf * 32;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
              v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
            }
            uint32_t rr[16];
            if (p.act[2] == RS_ACT_RELU) {
#pragma unroll
              for (int e = 0; e < 16; ++e) rr[e] = pack_bf16x2_relu(v[2 * e], v[2 * e + 1]);
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) rr[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
            }
            tmem_st_32x32_x16(R + 64 + hf * 16, rr);   // A3
            arrive(&a_ready[slot]);
          } else {
            out_chunk(R, tile, hf * 64, nullptr);
            out_chunk(R, tile, hf * 64 + 32, &a_ready[slot]);   // the accumulator may be overwritten by the next tile's L1
          }
          if (tr) p.trace[((r * 2 + slot) * 5 + step) * 2 + 1] = clock64();
        }
      }
    } else {
      // ---------------- all 16 warps, both slots, the issuer's fixed event order ----------------
      uint32_t cnt0 = 0, cnt1 = 0;
      for (int rr_ = 0; rr_ <= n_slot0; ++rr_) {
#pragma unroll 1
        for (int k = 0; k < 5; ++k) {
#pragma unroll 1
          for (int slot = 0; slot < 2; ++slot) {
            const int step = slot == 0 ? k : (k >= 2 ? k - 2 : k + 3);
            const int r = slot == 0 ? rr_ : (k < 2 ? rr_ - 1 : rr_);
            if (r < 0 || r >= (slot == 0 ? n_slot0 : n_slot1)) continue;
            const uint32_t R = tmem_base + slot * 256 + lane_sel;
            const int tile = cta + (2 * r + slot) * ncta;
            const bool tr = TRACE && blockIdx.x == 0 && r < 4 && ew == 0 && lane == 0;
            uint32_t& cnt = slot == 0 ? cnt0 : cnt1;
            chain_wait(&acc_ready[slot], cnt & 1u);
            ++cnt;
            tc_fence_after_sync();
            if (tr) p.trace[((r * 2 + slot) * 5 + step) * 2] = clock64();
            if (step <= 2) {
              chunk32(R + 64 + cg * 32, bv + (step == 2 ? b_off<C>(1) : step * 128) + cg * 32, p.act[step == 2 ? 1 : 0] == RS_ACT_RELU,
                      R + (step == 1 ? 0 : 192) + cg * 16);   // A1a, A1b, A2
              arrive(&a_ready[slot]);
            } else if (step == 3) {
              const float* bs = bv + b_off<C>(2) + cg * 16;
              float v[16];
              __syncwarp();
              tmem_ld_32x32_x16(R + cg * 16, v);
              arrive(&a0_free[slot]);   // columns [0,64) are dead: the gather warps may write the next A0
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
                v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
              }
              uint32_t rr[8];
              if (p.act[2] == RS_ACT_RELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) rr[e] = pack_bf16x2_relu(v[2 * e], v[2 * e + 1]);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) rr[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
              }
              tmem_st_32x32_x8(R + 64 + cg * 8, rr);   // A3
              arrive(&a_ready[slot]);
            } else {
              out_chunk(R, tile, cg * 32, &a_ready[slot]);
            }
            if (tr) p.trace[((r * 2 + slot) * 5 + step) * 2 + 1] = clock64();
          }
        }
      }
    }
    if (lane == 0) bulk_wait_all();   // the stage buffers (and this CTA) outlive every store
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == SP_ISSUER) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// RS_TOWER_SP: 0 = tower_pp_kernel (round 1), 1 (default) = tower_sp_kernel with the shared event order, 2 = with teams
// (measured equal within noise: the teams overlap the output steps but their chain steps are twice as long)
int tower_sp_mode() {
  static const int m = [] { const char* e = getenv("RS_TOWER_SP"); return e && e[0] >= '0' && e[0] <= '2' ? e[0] - '0' : 1; }();
  return m;
}
bool use_tower_sp() { return tower_sp_mode() != 0; }

// the new kernel serves the fused ranker's call shape: fp32 output through a TMA map (+ optional interleaved bf16 copy)
bool tower_sp_applies(const ChainParams& p) {
  for (int i = 0; i < p.n_prob; ++i)
    if (!p.prob[i].y_f32 || !p.prob[i].y_tma || p.prob[i].y_bf16) return false;
  return true;
}

void launch_tower_sp(const ChainMaps& maps, const ChainParams& p, cudaStream_t st) {
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  static cudaError_t attr_err[64];
  std::call_once(once[dev & 63], [dev] {
    attr_err[dev & 63] = cudaFuncSetAttribute(tower_sp_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tower_pp_smem(2));
    cudaFuncSetAttribute(tower_sp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tower_pp_smem(2));
    cudaFuncSetAttribute(tower_sp_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tower_pp_smem(2));
    cudaFuncSetAttribute(tower_sp_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tower_pp_smem(2));
    if (attr_err[dev & 63] != cudaSuccess) cudaGetLastError();
  });
  RS_CHECK(attr_err[dev & 63] == cudaSuccess, "tower_sp_kernel needs 227 KB of shared memory per block");
  int64_t max_tiles = 0;
  for (int i = 0; i < p.n_prob; ++i) max_tiles = std::max<int64_t>(max_tiles, (p.prob[i].rows + BM - 1) / BM);
  const int sms = rs::sm_count();
  int grid = static_cast<int>(std::min<int64_t>((max_tiles + 1) / 2 * p.n_prob, sms - (sms % p.n_prob)));
  grid = std::max(grid, p.n_prob);
  ChainParams q = p;
  static void* zero_row[64];   // per device: 512 zero bytes (what an invalid row id gathers); lives as long as the library
  static std::once_flag zonce[64];
  std::call_once(zonce[dev & 63], [dev] {
    if (cudaMalloc(&zero_row[dev & 63], 512) == cudaSuccess) cudaMemset(zero_row[dev & 63], 0, 512);
    else zero_row[dev & 63] = nullptr;
  });
  RS_CHECK(zero_row[dev & 63] != nullptr, "could not allocate the zero row");
  q.zero_row = zero_row[dev & 63];
  static const int dbg = [] { const char* e = getenv("RS_CHAIN_DEBUG"); return e ? atoi(e) : 0; }();
  q.debug = dbg;   // perf experiments: 16 / 64 skip the fp32 / interleaved bf16 output
  const char* tr_env = getenv("RS_CHAIN_TRACE");
  if (tr_env && tr_env[0] == '1') {   // debugging aid: phase stamps of CTA 0, relative to the first one
    RS_CUDA(cudaMalloc(&q.trace, 256 * sizeof(long long)));
    RS_CUDA(cudaMemsetAsync(q.trace, 0, 256 * sizeof(long long), st));
    if (tower_sp_mode() == 2) tower_sp_kernel<true, true><<<grid, SP_THREADS, tower_pp_smem(2), st>>>(maps, q);
    else tower_sp_kernel<false, true><<<grid, SP_THREADS, tower_pp_smem(2), st>>>(maps, q);
    RS_LAUNCH_CHECK();
    long long host[256];
    RS_CUDA(cudaStreamSynchronize(st));
    RS_CUDA(cudaMemcpy(host, q.trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(q.trace);
    long long t0 = host[120] ? host[120] : host[0];
    for (int t = 0; t < 8; ++t) {   // (round, slot)
      if (!host[t * 10]) continue;
      fprintf(stderr, "tower_sp round %d slot %d: issue", t / 2, t % 2);
      for (int k = 0; k < 5; ++k) fprintf(stderr, " %lld", host[80 + t * 5 + k] - t0);
      fprintf(stderr, " | epilogue (acc seen, done) x5:");
      for (int k = 0; k < 10; ++k) fprintf(stderr, " %lld", host[t * 10 + k] - t0);
      fprintf(stderr, "\n");
    }
    for (int j = 0; j < 8; ++j)
      if (host[120 + j * 3]) fprintf(stderr, "tower_sp gather tile %d: start %lld columns free %lld A0 ready %lld\n", j, host[120 + j * 3] - t0,
                                     host[121 + j * 3] - t0, host[122 + j * 3] - t0);
    return;
  }
  rs::ProfScope prof(RS_PROF_RANK, st);
  if (tower_sp_mode() == 2) tower_sp_kernel<true, false><<<grid, SP_THREADS, tower_pp_smem(2), st>>>(maps, q);
  else tower_sp_kernel<false, false><<<grid, SP_THREADS, tower_pp_smem(2), st>>>(maps, q);
  RS_LAUNCH_CHECK();
}

void launch_tower_pp(const ChainMaps& maps, const ChainParams& p, cudaStream_t st) {
  if (use_tower_sp() && tower_sp_applies(p)) return launch_tower_sp(maps, p, st);
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  static int mode_ok[64];   // highest mode whose shared-memory request the device accepted
  std::call_once(once[dev & 63], [dev] {
    cudaFuncSetAttribute(tower_pp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tower_pp_smem(0));
    cudaFuncSetAttribute(tower_pp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tower_pp_smem(1));
    const cudaError_t e2 = cudaFuncSetAttribute(tower_pp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tower_pp_smem(2));
    if (e2 != cudaSuccess) cudaGetLastError();
    mode_ok[dev & 63] = e2 == cudaSuccess ? 2 : 1;
  });
  static const int want = [] { const char* e = getenv("RS_TOWER_STAGE"); return e && e[0] >= '0' && e[0] <= '2' ? e[0] - '0' : 2; }();
  const int mode = std::min(want, mode_ok[dev & 63]);
  auto kern = mode == 2 ? tower_pp_kernel<2> : mode == 1 ? tower_pp_kernel<1> : tower_pp_kernel<0>;
  const int smem = tower_pp_smem(mode);
  int64_t max_tiles = 0;
  for (int i = 0; i < p.n_prob; ++i) max_tiles = std::max<int64_t>(max_tiles, (p.prob[i].rows + BM - 1) / BM);
  const int sms = rs::sm_count();
  // two tiles per CTA at a time: do not spread fewer than 2 tiles per CTA over more CTAs than needed
  int grid = static_cast<int>(std::min<int64_t>((max_tiles + 1) / 2 * p.n_prob, sms - (sms % p.n_prob)));
  grid = std::max(grid, p.n_prob);
  static const int lag = [] { const char* e = getenv("RS_TOWER_LAG"); return e && e[0] >= '0' && e[0] <= '5' ? e[0] - '0' : 3; }();
  ChainParams q = p;
  q.lag = lag;
  const char* tr_env = getenv("RS_CHAIN_TRACE");
  if (tr_env && tr_env[0] == '1') {   // debugging aid: per-round phase stamps of CTA 0 (both slots), relative to the first one
    if (const char* e = getenv("RS_CHAIN_DEBUG")) q.debug = atoi(e);   // 16 / 64: skip the fp32 / interleaved bf16 output
    RS_CUDA(cudaMalloc(&q.trace, 16 * 16 * sizeof(long long)));
    RS_CUDA(cudaMemsetAsync(q.trace, 0, 16 * 16 * sizeof(long long), st));
    kern<<<grid, PP_THREADS, smem, st>>>(maps, q);
    RS_LAUNCH_CHECK();
    long long host[16 * 16];
    RS_CUDA(cudaStreamSynchronize(st));
    RS_CUDA(cudaMemcpy(host, q.trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(q.trace);
    for (int t = 0; t < 16; ++t) {
      if (!host[t * 16 + 1]) continue;
      fprintf(stderr, "tower_pp slot %d round %d:", t / 8, t % 8);
      for (int k = 0; k < 13; ++k) fprintf(stderr, " %lld", host[t * 16 + k] ? host[t * 16 + k] - host[0] : -1);
      fprintf(stderr, "\n");
    }
    return;
  }
  rs::ProfScope prof(RS_PROF_RANK, st);
  kern<<<grid, PP_THREADS, smem, st>>>(maps, q);
  RS_LAUNCH_CHECK();
}

bool use_tower_pp() {
  const char* e = getenv("RS_TOWER_PP");
  return !(e && e[0] == '0');
}

template <class C>
void launch_chain(const ChainMaps& maps, ChainParams p, cudaStream_t st) {
  if (const char* e = getenv("RS_CHAIN_DEBUG")) p.debug = atoi(e);
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(chain_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain_smem<C>());
  });
  int64_t max_tiles = 0;
  for (int i = 0; i < p.n_prob; ++i) max_tiles = std::max<int64_t>(max_tiles, (p.prob[i].rows + BM - 1) / BM);
  const int sms = rs::sm_count();
  int grid = static_cast<int>(std::min<int64_t>(max_tiles * p.n_prob, sms - (sms % p.n_prob)));
  grid = std::max(grid, p.n_prob);
  rs::ProfScope prof(RS_PROF_RANK, st);
  chain_tc_kernel<C><<<grid, NTHREADS, chain_smem<C>(), st>>>(maps, p);
  RS_LAUNCH_CHECK();
}

}  // namespace

// One entry for both chain shapes.  n_prob = 1 or 2 independent problems with the same layer shapes (the two
// towers).  Per problem p: wts[p*n_layers + l] bf16 [N_l, K_l], biases[p*n_layers + l] f32 [N_l] or NULL.
// Tower chain: tables[p] f32 [V,128] + ids[p]; deep chain: x[p] bf16 [B,256].
extern "C" int rs_chain_forward_bf16(int n_prob, int n_layers, const DLTensor* const* wts, const DLTensor* const* biases,
                                     const int* acts, const DLTensor* const* tables, const DLTensor* const* ids,
                                     const DLTensor* const* x, DLTensor* const* y_f32, DLTensor* const* y_bf16,
                                     void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(n_prob == 1 || n_prob == 2, "n_prob must be 1 or 2");
    RS_CHECK(n_layers == 4 || n_layers == 2, "supported chains: 4 layers (tower) or 2 layers (deep net)");
    const bool tower = n_layers == 4;
    ChainMaps maps;
    ChainParams p{};
    p.n_prob = n_prob;
    for (int l = 0; l < n_layers; ++l) {
      RS_CHECK(acts[l] == RS_ACT_NONE || acts[l] == RS_ACT_RELU, "chain activations must be relu or linear");
      p.act[l] = acts[l];
    }
    sm_count();
    for (int q = 0; q < n_prob; ++q) {
      ChainProblem& pr = p.prob[q];
      for (int l = 0; l < n_layers; ++l) {
        const DLTensor* w = wts[q * n_layers + l];
        require(w, "wt", 2, kDLBfloat, 16);
        const int K = tower ? CfgTower::K(l) : CfgDeep::K(l);
        const int N = tower ? CfgTower::N(l) : CfgDeep::N(l);
        RS_CHECK(on_device(w) && is_compact(w) && w->shape[0] == N && w->shape[1] == K,
                 "fused chain supports only 128-256-128-64-128 (tower) and 256-256-128 (deep); layer " + std::to_string(l) +
                     " has shape [" + std::to_string(w->shape[0]) + ", " + std::to_string(w->shape[1]) + "]");
        maps.w[q][l] = make_tmap_2d(data_ptr<void>(w), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)N, (uint64_t)K,
                                    (uint64_t)K * 2, static_cast<uint32_t>(N), 64);
        const DLTensor* b = biases ? biases[q * n_layers + l] : nullptr;
        if (b) {
          require(b, "bias", 1, kDLFloat, 32);
          RS_CHECK(on_device(b) && b->shape[0] == N && stride_of(b, 0) == 1, "bias must be a contiguous CUDA [out] tensor");
          pr.bias[l] = data_ptr<float>(b);
        }
      }
      const int n_last = tower ? CfgTower::N(3) : CfgDeep::N(1);
      int64_t rows = -1;
      if (tower) {
        const DLTensor* tb = tables[q];
        const DLTensor* id = ids[q];
        require(tb, "table", 2, kDLFloat, 32);
        RS_CHECK(id && id->ndim == 1 && (is_dtype(id, kDLInt, 32) || is_dtype(id, kDLInt, 64)), "ids must be 1-D int32/int64");
        RS_CHECK(on_device(tb) && on_device(id) && tb->shape[1] == 128, "table must be a CUDA [V, 128] tensor");
        RS_CHECK(id->shape[0] <= 1 || stride_of(id, 0) == 1, "ids must be contiguous");
        pr.table = data_ptr<float>(tb);
        pr.table_ld = row_stride_2d(tb, "table");
        pr.table_rows = tb->shape[0];
        RS_CHECK(pr.table_ld % 4 == 0, "table row stride must be a multiple of 4");
        require_aligned16(pr.table, "table");
        pr.ids = data_ptr<void>(id);
        pr.ids64 = is_dtype(id, kDLInt, 64) ? 1 : 0;
        rows = id->shape[0];
      } else {
        const DLTensor* xx = x[q];
        require(xx, "x", 2, kDLBfloat, 16);
        RS_CHECK(on_device(xx) && xx->shape[1] == 256, "x must be a CUDA bf16 [B, 256] tensor");
        pr.x = data_ptr<uint16_t>(xx);
        pr.x_ld = row_stride_2d(xx, "x");
        RS_CHECK(pr.x_ld % 8 == 0, "x row stride must be a multiple of 8");
        require_aligned16(pr.x, "x");
        rows = xx->shape[0];
      }
      pr.rows = rows;
      const DLTensor* yf = y_f32 ? y_f32[q] : nullptr;
      const DLTensor* yb = y_bf16 ? y_bf16[q] : nullptr;
      RS_CHECK(yf || yb, "each problem needs at least one output");
      if (yf) {
        require(yf, "y_f32", 2, kDLFloat, 32);
        RS_CHECK(on_device(yf) && yf->shape[0] == rows && yf->shape[1] == n_last, "y_f32 must be [B, out]");
        pr.y_f32 = data_ptr<float>(yf);
        pr.ld_f32 = row_stride_2d(yf, "y_f32");
        RS_CHECK(pr.ld_f32 % 8 == 0 && (reinterpret_cast<uintptr_t>(pr.y_f32) & 31) == 0, "y_f32 must be 32-byte aligned with a row stride multiple of 8");
      }
      if (yb) {
        require(yb, "y_bf16", 2, kDLBfloat, 16);
        RS_CHECK(on_device(yb) && yb->shape[0] == rows && yb->shape[1] == n_last, "y_bf16 must be [B, out]");
        pr.y_bf16 = data_ptr<uint16_t>(yb);
        pr.ld_bf16 = row_stride_2d(yb, "y_bf16");
        RS_CHECK(pr.ld_bf16 % 16 == 0 && (reinterpret_cast<uintptr_t>(pr.y_bf16) & 31) == 0, "y_bf16 must be 32-byte aligned with a row stride multiple of 16");
      }
    }
    bool any = false;
    for (int q = 0; q < n_prob; ++q) any |= p.prob[q].rows > 0;
    if (!any) return;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (tower && use_tower_pp()) launch_tower_pp(maps, p, st);
    else if (tower) launch_chain<CfgTower>(maps, p, st);
    else launch_chain<CfgDeep>(maps, p, st);
  });
}

// ================================================================================================
// rs_ranker: MultiTaskModel.call (src/models.py:125-131) behind one handle and ONE C call per batch.
// The handle owns bf16 copies of the Dense kernels ([out][in]), the fp32 biases, cross vectors and heads, and the
// TMA maps over them; the embedding tables are borrowed (they can be tens of GB) and must outlive the handle.
//   forward: chain<Tower> (both towers, gather fused) -> chain<Deep> -> cross_heads_kernel
// ================================================================================================
namespace {

// DCN-v1 cross stack + both heads for D = 256, deep width 128, without materialising the [B,384] concat:
// one warp per row; x0/x_l live in registers (8 per lane), deep output 4 per lane.  All cross / head vectors are
// hoisted into registers once per warp (L is a template parameter), so a row costs three 128-bit loads.
template <int L>
__global__ void __launch_bounds__(256) cross_heads_kernel(const float* __restrict__ x0p, int64_t ldx,
                                                          const float* __restrict__ deep, int64_t ldd,
                                                          const float* __restrict__ cw, const float* __restrict__ cb,
                                                          const float* __restrict__ hw, const float* __restrict__ hb,
                                                          float* __restrict__ rating, float* __restrict__ ctr, int64_t B) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int D = 256, W = 384;
  float4 wv[L > 0 ? L : 1][2], bv[L > 0 ? L : 1][2], wr[3], wc[3];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      wv[l][c] = __ldg(reinterpret_cast<const float4*>(cw + (int64_t)l * D) + lane + 32 * c);
      bv[l][c] = __ldg(reinterpret_cast<const float4*>(cb + (int64_t)l * D) + lane + 32 * c);
    }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    wr[c] = __ldg(reinterpret_cast<const float4*>(hw) + lane + 32 * c);
    wc[c] = __ldg(reinterpret_cast<const float4*>(hw + W) + lane + 32 * c);
  }
  const float hb0 = hb[0], hb1 = hb[1];
  auto dot4 = [](const float4& a, const float4& b, float s) {
    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    return s;
  };
  // the next row's three loads are issued before this row's arithmetic (one row in flight per warp left the kernel
  // latency-bound at ~2.6 TB/s: ~110 registers per thread allow only 16 warps per SM)
  float4 nx0[2], ndv;
  if (warp < B) {
#pragma unroll
    for (int c = 0; c < 2; ++c) nx0[c] = __ldg(reinterpret_cast<const float4*>(x0p + warp * ldx) + lane + 32 * c);
    ndv = __ldg(reinterpret_cast<const float4*>(deep + warp * ldd) + lane);
  }
  for (int64_t r = warp; r < B; r += nwarps) {
    float4 x0[2], xl[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      x0[c] = nx0[c];
      xl[c] = x0[c];
    }
    const float4 dv = ndv;
    if (r + nwarps < B) {
#pragma unroll
      for (int c = 0; c < 2; ++c) nx0[c] = __ldg(reinterpret_cast<const float4*>(x0p + (r + nwarps) * ldx) + lane + 32 * c);
      ndv = __ldg(reinterpret_cast<const float4*>(deep + (r + nwarps) * ldd) + lane);
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
      float s = dot4(xl[1], wv[l][1], dot4(xl[0], wv[l][0], 0.0f));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        xl[c].x = (x0[c].x * s + bv[l][c].x) + xl[c].x;   // bias outside the product (src/models.py:44)
        xl[c].y = (x0[c].y * s + bv[l][c].y) + xl[c].y;
        xl[c].z = (x0[c].z * s + bv[l][c].z) + xl[c].z;
        xl[c].w = (x0[c].w * s + bv[l][c].w) + xl[c].w;
      }
    }
    float sr = dot4(dv, wr[2], dot4(xl[1], wr[1], dot4(xl[0], wr[0], 0.0f)));
    float sc = dot4(dv, wc[2], dot4(xl[1], wc[1], dot4(xl[0], wc[0], 0.0f)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_xor_sync(0xFFFFFFFFu, sr, o);
      sc += __shfl_xor_sync(0xFFFFFFFFu, sc, o);
    }
    if (lane == 0) {
      rating[r] = sr + hb0;
      ctr[r] = 1.0f / (1.0f + expf(-(sc + hb1)));
    }
  }
}


// ------------------------------------------------------------------------------------------------
// deep_heads_kernel: the deep chain, the DCN-v1 cross stack and both heads in ONE kernel (src/models.py:38-48 and
// 119-120,131).  Same tcgen05 pipeline as chain_tc_kernel<CfgDeep>; what is fused around it:
//   * the layer-0 A operand comes from tower_pp_kernel's lane-interleaved bf16 copy of x0 (ChainProblem::y_il16):
//     eight fully coalesced 128-bit loads per thread that land directly in the registers tcgen05.st takes, issued a
//     whole tile ahead and not touched until then, so their latency and the HBM burst are off the critical path;
//   * the x0 dots of the cross stack, P_l = x0 . w_l (l < LC), Hr = x0 . head_r[:256], Hc = x0 . head_c[:256], ride on
//     the tensor pipe: layer 0 gets 16 extra output columns (one more N = 16 MMA group on the same A operand) whose
//     "kernel" rows are the bf16 hi and lo halves of each vector (w = hi + lo to ~2^-17, so only x0's own bf16
//     rounding -- the one the deep net sees anyway -- is left), accumulated in fp32 in 16 spare TMEM columns.  The
//     tower kernel used to take these dots on the CUDA cores (6 x 128 FMAs per row fed by broadcast shared-memory
//     loads); at 128 B/clk of shared-memory return bandwidth that was 6 k cycles per pair of tiles, the largest
//     single item of its output step (measured with RS_CHAIN_TRACE=1 RS_CHAIN_DEBUG=..);
//   * the last epilogue dots the fp32 deep output (bias + relu applied) with head_r[256:], head_c[256:] instead of
//     writing it;
//   * one thread per row then runs the cross recurrence in closed form.  x_{l+1} = x0 (x_l . w_l) + b_l + x_l keeps
//     x_l = a_l x0 + c_l with c_l = b_0 + .. + b_{l-1}, so  s_l = a_l P_l + c_l . w_l,  a_{l+1} = a_l + s_l, and
//     rating = a_L Hr + c_L . head_r[:256] + deep . head_r[256:] + bias  (ctr likewise, through a sigmoid).
//     The c_l dots are per-model constants (ranker_consts_kernel).
// Per sample the kernel reads 512 bytes and writes 8.
// ------------------------------------------------------------------------------------------------
struct DeepHeadsParams {
  int64_t rows;
  const uint16_t* x_il16;  // bf16 x0, lane-interleaved (ChainProblem::y_il16)
  const float* bias[2];
  const float* hw;         // [2, 384]
  const float* hb;         // [2]
  const float* consts;     // [LC + 2]: c_l . w_l for l < LC, c_L . head_r[:256], c_L . head_c[:256]  (ranker_consts_kernel)
  float* rating;           // [rows]
  float* ctr;              // [rows]
  long long* trace;        // RS_CHAIN_TRACE=1: clock64 stamps of CTA 0 ([tile][16]: 0-7 epilogue warp 4, 8-11 MMA issuer)
  // deep_heads2_kernel: biases (256 + 128) and the deep part of the two head rows ([2][128]) BY VALUE -- kernel parameters
  // sit in the constant bank, so the epilogues read them with LDC instead of LDS (see ChainParams::biasv)
  float biasv[384];
  float hdv[256];
};

// per-model constants of the closed-form cross recurrence, computed once when the ranker handle is created
__global__ void ranker_consts_kernel(const float* __restrict__ cw, const float* __restrict__ cb,
                                     const float* __restrict__ hw, int L, float* __restrict__ out) {
  const int lane = threadIdx.x;
  float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  for (int j = lane; j < 256; j += 32) {
    float c = 0.0f;
    for (int l = 0; l < L; ++l) {
      acc[l] = fmaf(c, cw[l * 256 + j], acc[l]);
      c += cb[l * 256 + j];
    }
    acc[4] = fmaf(c, hw[j], acc[4]);
    acc[5] = fmaf(c, hw[384 + j], acc[5]);
  }
#pragma unroll
  for (int t = 0; t < 6; ++t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xFFFFFFFFu, acc[t], o);
  }
  if (lane == 0) {
    for (int l = 0; l < L; ++l) out[l] = acc[l];
    out[L] = acc[4];
    out[L + 1] = acc[5];
  }
}

constexpr int DOTW_ROWS = 16;                       // extra layer-0 output columns: (hi, lo) x (LC + 2 <= 6 vectors), zero-padded
constexpr int DOTW_BYTES = DOTW_ROWS * 256 * 2;     // bf16 [16][256], K-major, SWIZZLE_128B: 4 K chunks of [16][64]

// the 16 extra layer-0 rows of the deep net: row 2t = bf16(v_t), row 2t+1 = bf16(v_t - bf16(v_t)) for the vectors
// v_t = cross kernel t (t < L), then head_r[:256], head_c[:256]; rows beyond 2(L + 2) are zero
__global__ void ranker_dotw_kernel(const float* __restrict__ cw, const float* __restrict__ hw, int L, uint16_t* __restrict__ out) {
  const int j = threadIdx.x;   // 256 threads: one feature each
  for (int t = 0; t < DOTW_ROWS / 2; ++t) {
    float v = 0.0f;
    if (t < L) v = cw[t * 256 + j];
    else if (t < L + 2) v = hw[(t - L) * 384 + j];
    const uint32_t hi = pack_bf16x2(v, 0.0f) & 0xFFFFu;
    const float hif = __uint_as_float(hi << 16);
    const uint32_t lo = pack_bf16x2(v - hif, 0.0f) & 0xFFFFu;
    out[(2 * t) * 256 + j] = static_cast<uint16_t>(hi);
    out[(2 * t + 1) * 256 + j] = static_cast<uint16_t>(lo);
  }
}

template <int LC>
constexpr int deep_heads_smem() {
  return w_off<CfgDeep>(2) + DOTW_BYTES + (384 + 256 + 8 + 2 * 512) * 4 + 256 + 1024;
}

template <int LC>
__global__ void __launch_bounds__(NTHREADS, 1)
deep_heads_kernel(const __grid_constant__ ChainMaps maps, const DeepHeadsParams p) {
  using C = CfgDeep;
  constexpr int NP = LC + 2;   // dots over x0 per row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = base;
  uint8_t* dotw_s = base + w_off<C>(2);       // the 16 extra layer-0 "kernel" rows (maps.w[0][2])
  float* bias_s = reinterpret_cast<float*>(dotw_s + DOTW_BYTES);
  float* hd_s = bias_s + 384;                 // [2][128]: the deep part of the two head rows
  float* const_s = hd_s + 256;                // [LC] c_l . w_l, then c_L . head_r, c_L . head_c
  float* part = const_s + 8;                  // [2][4 slices][128 rows]: deep . head_r[256:], deep . head_c[256:] partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + 2 * 512);
  uint64_t* wfull = bars;
  uint64_t* a_ready = bars + 1;
  uint64_t* acc_ready = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int ncta = gridDim.x;
  const int64_t ntiles = (p.rows + BM - 1) / BM;
  const int64_t n_my = cta < ntiles ? (ntiles - cta + ncta - 1) / ncta : 0;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < 3; ++l) tma_prefetch_desc(&maps.w[0][l]);
    mbar_init(wfull, 1);
    mbar_init(a_ready, EPI_WARPS);
    mbar_init(acc_ready, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 384; i += NTHREADS) {
    const int l = i < 256 ? 0 : 1;
    const float* b = p.bias[l];
    bias_s[i] = b ? b[i - l * 256] : 0.0f;
  }
  for (int i = threadIdx.x; i < 256; i += NTHREADS) hd_s[i] = p.hw[(i >> 7) * 384 + 256 + (i & 127)];
  if (threadIdx.x < NP) const_s[threadIdx.x] = p.consts[threadIdx.x];
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
      mbar_arrive_expect_tx(wfull, w_off<C>(2) + DOTW_BYTES);
#pragma unroll
      for (int l = 0; l < 2; ++l)
        for (int kc = 0; kc < 4; ++kc)
          tma_load_2d(wsm + w_off<C>(l) + kc * (C::N(l) * 128), &maps.w[0][l], wfull, kc * 64, 0, kEvictLast);
      for (int kc = 0; kc < 4; ++kc)
        tma_load_2d(dotw_s + kc * (DOTW_ROWS * 128), &maps.w[0][2], wfull, kc * 64, 0, kEvictLast);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(wfull, 0);
      const uint64_t wdesc0 = umma_desc_k_sw128(smem_u32(wsm));
      uint32_t use = 0;
      for (int64_t i = 0; i < n_my; ++i) {
#pragma unroll
        for (int l = 0; l < 2; ++l, ++use) {
          chain_wait(a_ready, use & 1u);
          tc_fence_after_sync();
          if (p.trace && cta == 0 && i < 8) p.trace[i * 16 + 8 + 2 * l] = clock64();
          const uint32_t idesc = umma_idesc_bf16_f32(BM, C::N(l));
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const uint64_t db = wdesc0 + static_cast<uint64_t>((w_off<C>(l) + (k >> 2) * (C::N(l) * 128) + (k & 3) * 32) >> 4);
            umma_bf16_ts(tmem_base + C::DCOL(l), tmem_base + C::ACOL(l) + k * 8, db, idesc, k > 0 ? 1u : 0u);
          }
          if (l == 0) {   // the x0 dots: same A operand, 16 more output columns -> TMEM [128, 144) (free until layer 1)
            const uint32_t idesc_d = umma_idesc_bf16_f32(BM, DOTW_ROWS);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const uint64_t db = wdesc0 + static_cast<uint64_t>((w_off<C>(2) + (k >> 2) * (DOTW_ROWS * 128) + (k & 3) * 32) >> 4);
              umma_bf16_ts(tmem_base + 128, tmem_base + k * 8, db, idesc_d, k > 0 ? 1u : 0u);
            }
          }
          umma_commit(acc_ready);
          if (p.trace && cta == 0 && i < 8) p.trace[i * 16 + 9 + 2 * l] = clock64();
        }
      }
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    const int slice = (warp - 4) >> 2;
    const int trow = quarter * 32 + lane;                 // this thread's row inside a tile
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const float hb0 = p.hb[0], hb1 = p.hb[1];
    uint32_t use = 0;
    // One tile ahead, in registers: the A operand (this row's 64-feature slice, 8 x 128 bits).  Nothing reads it
    // before the next iteration.
    uint4 a0[8];
    float pd[NP];   // slice-0 warps: this row's x0 dots of the tile in progress (read from TMEM after layer 0)
#pragma unroll
    for (int t = 0; t < NP; ++t) pd[t] = 0.0f;
    // unconditional: x_il16 is padded to whole tiles and the tower kernel writes every row of every tile (rows past the
    // end of the batch hold the towers' output for an all-zero embedding; their results are never stored), and a
    // "valid ? load : 0" select would be a use of the loaded value right here
    auto load_next = [&](int64_t tile_) {
      const uint4* src = reinterpret_cast<const uint4*>(p.x_il16) + ((tile_ * 4 + quarter) * 32 + slice * 8) * 32 + lane;
#pragma unroll
      for (int e = 0; e < 8; ++e) a0[e] = __ldg(src + e * 32);
    };
    // cross recurrence + heads for one finished tile: run by the slice-0 warp of each quarter once the other three
    // warps of the quarter have published their partial dots (named barrier 1 + quarter, 128 threads)
    auto finish = [&](int64_t tile_) {
      if (slice != 0) {   // bar.arrive orders this thread's earlier shared-memory writes before the barrier completes
        asm volatile("bar.arrive %0, 128;" ::"r"(1 + quarter) : "memory");
        return;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      float s[NP + 2];
#pragma unroll
      for (int t = 0; t < NP; ++t) s[t] = pd[t];
#pragma unroll
      for (int t = 0; t < 2; ++t)
        s[NP + t] = (part[(t * 4 + 0) * 128 + trow] + part[(t * 4 + 1) * 128 + trow]) +
                    (part[(t * 4 + 2) * 128 + trow] + part[(t * 4 + 3) * 128 + trow]);
      float a = 1.0f;
#pragma unroll
      for (int l = 0; l < LC; ++l) a += fmaf(a, s[l], const_s[l]);
      const float zr = fmaf(a, s[LC], const_s[LC]) + s[NP] + hb0;
      const float zc = fmaf(a, s[LC + 1], const_s[LC + 1]) + s[NP + 1] + hb1;
      const int64_t row_ = tile_ * BM + trow;
      if (row_ < p.rows) {
        p.rating[row_] = zr;
        p.ctr[row_] = __fdividef(1.0f, 1.0f + __expf(-zc));   // MUFU only: an IEEE division / expf has a slow path that is an ABI
                                                                // call, and a call inside the epilogue spills everything live
      }
    };
    if (n_my > 0) load_next(cta);
    for (int64_t i = 0; i < n_my; ++i) {
      const int64_t tile = cta + i * ncta;
      const bool tr = p.trace && cta == 0 && i < 8 && warp == 4 && lane == 0;
      auto stamp = [&](int k) { if (tr) p.trace[i * 16 + k] = clock64(); };
      __syncwarp();
      stamp(0);
      {
        uint32_t (&r0)[16] = reinterpret_cast<uint32_t (&)[16]>(a0[0]);
        uint32_t (&r1)[16] = reinterpret_cast<uint32_t (&)[16]>(a0[4]);
        tmem_st_32x32_x16(tmem_base + slice * 32 + lane_sel, r0);
        tmem_st_32x32_x16(tmem_base + slice * 32 + 16 + lane_sel, r1);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      stamp(1);
      // the previous tile's rows, while this tile's layer 0 runs; before the loads below are issued, so that the
      // fence in front of the named barrier has nothing outstanding to wait for
      if (i > 0) finish(tile - ncta);
      stamp(12);
      if (i + 1 < n_my) load_next(tile + ncta);   // in flight for the whole tile
      stamp(2);
      if (i + 2 < n_my && warp == 4 && lane == 0) {   // and the tile after that on its way from HBM to L2
        const int64_t t2 = tile + 2 * ncta;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.x_il16 + t2 * (BM * 256)), "r"(BM * 256 * 2) : "memory");
      }
      stamp(3);

      // layer 0: 256 accumulator columns, 64 per warp -> bias, relu, bf16 -> A operand of layer 1
      chain_wait(acc_ready, use & 1u);
      ++use;
      tc_fence_after_sync();
      stamp(4);
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = slice * 64 + c * 32;
        float v[32];
        __syncwarp();
        tmem_ld_32x32(tmem_base + C::DCOL(0) + col0 + lane_sel, v);
        const float* bs = bias_s + col0;
        uint32_t r[16];
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
          r[e / 2] = pack_bf16x2_relu(v[e] + b4.x, v[e + 1] + b4.y);
          r[e / 2 + 1] = pack_bf16x2_relu(v[e + 2] + b4.z, v[e + 3] + b4.w);
        }
        tmem_st_32x32_x16(tmem_base + (col0 >> 1) + lane_sel, r);
      }
      if (slice == 0) {   // the x0 dots of this row: columns [128, 144) hold (hi, lo) pairs
        float d[16];
        tmem_ld_32x32_x16(tmem_base + 128 + lane_sel, d);
#pragma unroll
        for (int t = 0; t < NP; ++t) pd[t] = d[2 * t] + d[2 * t + 1];
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      stamp(5);

      // layer 1: 128 columns, 32 per warp -> bias, relu, dotted with the deep part of both heads
      chain_wait(acc_ready, use & 1u);
      ++use;
      tc_fence_after_sync();
      stamp(6);
      {
        const int col0 = slice * 32;
        float v[32];
        __syncwarp();
        tmem_ld_32x32(tmem_base + C::DCOL(1) + col0 + lane_sel, v);
        const float* bs = bias_s + 256 + col0;
        const float* wr = hd_s + col0;
        const float* wc = hd_s + 128 + col0;
        float dr = 0.0f, dq = 0.0f;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
          const float4 r4 = *reinterpret_cast<const float4*>(wr + e);
          const float4 c4 = *reinterpret_cast<const float4*>(wc + e);
          const float y0 = fmaxf(v[e] + b4.x, 0.0f), y1 = fmaxf(v[e + 1] + b4.y, 0.0f);
          const float y2 = fmaxf(v[e + 2] + b4.z, 0.0f), y3 = fmaxf(v[e + 3] + b4.w, 0.0f);
          dr = fmaf(y0, r4.x, dr); dr = fmaf(y1, r4.y, dr); dr = fmaf(y2, r4.z, dr); dr = fmaf(y3, r4.w, dr);
          dq = fmaf(y0, c4.x, dq); dq = fmaf(y1, c4.y, dq); dq = fmaf(y2, c4.z, dq); dq = fmaf(y3, c4.w, dq);
        }
        part[(0 * 4 + slice) * 128 + trow] = dr;
        part[(1 * 4 + slice) * 128 + trow] = dq;
      }
      stamp(7);
      // the next tile's A0 (columns [0,128)) does not overlap this tile's last accumulator ([128,256)), which is
      // fully read above; `part` is re-written only after every warp has passed the next tile's layer-0 barrier,
      // whose MMA needs the slice-0 warps' layer-1 operand arrival, made after their finish()
    }
    if (n_my > 0) finish(cta + (n_my - 1) * ncta);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int LC>
void launch_deep_heads(const ChainMaps& maps, const DeepHeadsParams& p, cudaStream_t st) {
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(deep_heads_kernel<LC>, cudaFuncAttributeMaxDynamicSharedMemorySize, deep_heads_smem<LC>());
  });
  const int64_t ntiles = (p.rows + BM - 1) / BM;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ntiles, rs::sm_count())));
  rs::ProfScope prof(RS_PROF_DEEP, st);
  deep_heads_kernel<LC><<<grid, NTHREADS, deep_heads_smem<LC>(), st>>>(maps, p);
  RS_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// deep_heads2_kernel (round 2): the same arithmetic as deep_heads_kernel (bit-identical outputs), pipelined INSIDE a tile.
// deep_heads_kernel runs a tile as MMA(L0) -> epilogue -> MMA(L1) -> epilogue, one after the other (8.2 k cycles per tile
// for 3.3 k of MMA, RS_CHAIN_TRACE), and two tiles in flight do not fit tensor memory (A0 128 + accumulator 272 columns).
// Here layer 0 runs as two N halves into two accumulator regions and layer 1 as two K halves:
//     L0a: A0 x [W1 rows 0-127 | 16 dot rows] -> X   (N = 144: the x0 dots ride in the same MMAs; they used to be 16
//                                                     extra N = 16 instructions, which cost 45 cycles each like any MMA
//                                                     with N <= 64 -- tools/ubench_tc.cu -- 720 cycles per tile)
//     L0b: A0 x  W1 rows 128-255                -> Y     || epilogue a: X -> bias, relu, bf16 -> A1a; dots -> registers
//     L1a: A1a x W2[:, 0:128]                   -> X     || epilogue b: Y -> A1b (over A0's first half, dead by then)
//     L1b: A1b x W2[:, 128:256]                 -> X (accumulate)
//     epilogue 1: next tile's A0 -> tensor memory first (the issuer starts its L0a at once), then X -> heads.
// X / Y swap every tile, so the next tile's L0a never waits for this tile's last epilogue.
// Tensor memory columns: A0 [0,128)  RA [128,272)  RB [272,416)  A1a [416,480).
// ------------------------------------------------------------------------------------------------
constexpr int DH2_W1B_OFF = 128 * 128 + DOTW_ROWS * 128;          // within a K chunk: W1 rows 0-127 | dot rows | W1 rows 128-255
constexpr int DH2_KC_BYTES = DH2_W1B_OFF + 128 * 128;             // 34,816 bytes per 64-wide K chunk of layer 0
constexpr int DH2_W2_OFF = 4 * DH2_KC_BYTES;
constexpr int DH2_W_BYTES = DH2_W2_OFF + 4 * 128 * 128;           // 204,800: what deep_heads_kernel keeps resident as well
constexpr int DH2_A0 = 0, DH2_RA = 128, DH2_RB = 272, DH2_A1A = 416;
// warps 0-15 epilogue (TMEM lane quarter = warp % 4, column slice = warp / 4), warp 16: one thread loads the kernels (TMA) and
// issues every MMA.  17 warps = 544 threads: 120 registers per thread, nothing spills (a spill is an L2 round trip here)
constexpr int DH2_ISSUER = EPI_WARPS;
constexpr int DH2_THREADS = (EPI_WARPS + 1) * 32;
static_assert(DH2_KC_BYTES % 1024 == 0 && DH2_W1B_OFF % 1024 == 0, "operand tiles must stay 1024-byte aligned (128B swizzle)");

template <int LC>
constexpr int deep_heads2_smem() {
  return DH2_W_BYTES + (384 + 256 + 8 + 2 * 512) * 4 + 256 + 1024;
}

template <int LC, bool TRACE>
__global__ void __launch_bounds__(DH2_THREADS, 1)
deep_heads2_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ DeepHeadsParams p) {
  constexpr int NP = LC + 2;   // dots over x0 per row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = base;
  float* bias_s = reinterpret_cast<float*>(base + DH2_W_BYTES);
  float* hd_s = bias_s + 384;                 // [2][128]: the deep part of the two head rows
  float* const_s = hd_s + 256;                // [LC] c_l . w_l, then c_L . head_r, c_L . head_c
  float* part = const_s + 8;                  // [2][4 slices][128 rows]: deep . head_r[256:], deep . head_c[256:] partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + 2 * 512);
  uint64_t* wfull = bars;
  uint64_t* a0_ready = bars + 1;     // 16 epilogue warps: this tile's A0 is in tensor memory
  uint64_t* acc_a = bars + 2;        // commit: L0a (and the dots) done
  uint64_t* acc_b = bars + 3;        // commit: L0b done
  uint64_t* a1a_ready = bars + 4;    // 16: A1a written (X has been read)
  uint64_t* a1b_ready = bars + 5;    // 16: A1b written (Y has been read)
  uint64_t* acc1_ready = bars + 6;   // commit: layer 1 done
  uint64_t* acc1_read = bars + 7;    // 16: X has been read again -> the next tile may use it as its Y
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  // broadcast from lane 0: a provably uniform warp index keeps the role branches and the constant-bank addresses derived
  // from it (biases, head rows) on the uniform datapath -- the adds then take their bias straight from uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int ncta = gridDim.x;
  const int ntiles = static_cast<int>((p.rows + BM - 1) / BM);   // 32-bit tile arithmetic: registers are scarce
  const int n_my = cta < ntiles ? (ntiles - cta + ncta - 1) / ncta : 0;

  if (warp == DH2_ISSUER && lane == 0) {
    tma_prefetch_desc(&maps.w[1][0]);
    tma_prefetch_desc(&maps.w[0][1]);
    tma_prefetch_desc(&maps.w[0][2]);
    mbar_init(wfull, 1);
    mbar_init(a0_ready, EPI_WARPS);
    mbar_init(acc_a, 1);
    mbar_init(acc_b, 1);
    mbar_init(a1a_ready, EPI_WARPS);
    mbar_init(a1b_ready, EPI_WARPS);
    mbar_init(acc1_ready, 1);
    mbar_init(acc1_read, EPI_WARPS);
    fence_mbar_init();
  }
  if (threadIdx.x < NP) const_s[threadIdx.x] = p.consts[threadIdx.x];
  if (warp == DH2_ISSUER) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == DH2_ISSUER) {
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, DH2_W_BYTES);
      for (int kc = 0; kc < 4; ++kc) {   // maps.w[1][0]: W1 [256, 256] with a 128-row box
        tma_load_2d(wsm + kc * DH2_KC_BYTES, &maps.w[1][0], wfull, kc * 64, 0, kEvictLast);
        tma_load_2d(wsm + kc * DH2_KC_BYTES + 128 * 128, &maps.w[0][2], wfull, kc * 64, 0, kEvictLast);
        tma_load_2d(wsm + kc * DH2_KC_BYTES + DH2_W1B_OFF, &maps.w[1][0], wfull, kc * 64, 128, kEvictLast);
      }
      for (int kc = 0; kc < 4; ++kc)
        tma_load_2d(wsm + DH2_W2_OFF + kc * (128 * 128), &maps.w[0][1], wfull, kc * 64, 0, kEvictLast);
      mbar_wait(wfull, 0);
      uint64_t wdesc0 = umma_desc_k_sw128(smem_u32(wsm));
      constexpr uint32_t idesc_a = umma_idesc_bf16_f32(BM, 128 + DOTW_ROWS);
      constexpr uint32_t idesc_h = umma_idesc_bf16_f32(BM, 128);
      for (int i = 0; i < n_my; ++i) {
        // opaque to the optimiser: otherwise all 48 operand descriptors (wdesc0 + constant) are hoisted out of the loop as
        // loop invariants, 96 registers' worth, and come back as spill loads on the issue path
        asm volatile("" : "+l"(wdesc0));
        const uint32_t ph = static_cast<uint32_t>(i & 1);
        const uint32_t X = tmem_base + (ph ? DH2_RB : DH2_RA), Y = tmem_base + (ph ? DH2_RA : DH2_RB);
        mbar_wait(a0_ready, ph);
        tc_fence_after_sync();
        if (TRACE && cta == 0 && i < 8) p.trace[i * 16 + 8] = clock64();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint64_t db = wdesc0 + static_cast<uint64_t>(((k >> 2) * DH2_KC_BYTES + (k & 3) * 32) >> 4);
          umma_bf16_ts(X, tmem_base + DH2_A0 + k * 8, db, idesc_a, k > 0 ? 1u : 0u);
        }
        umma_commit(acc_a);
        if (i > 0) {   // Y is the previous tile's X: its layer-1 accumulator must have been read
          mbar_wait(acc1_read, static_cast<uint32_t>((i - 1) & 1));
          tc_fence_after_sync();
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint64_t db = wdesc0 + static_cast<uint64_t>(((k >> 2) * DH2_KC_BYTES + DH2_W1B_OFF + (k & 3) * 32) >> 4);
          umma_bf16_ts(Y, tmem_base + DH2_A0 + k * 8, db, idesc_h, k > 0 ? 1u : 0u);
        }
        umma_commit(acc_b);
        if (TRACE && cta == 0 && i < 8) p.trace[i * 16 + 9] = clock64();
        mbar_wait(a1a_ready, ph);
        tc_fence_after_sync();
        if (TRACE && cta == 0 && i < 8) p.trace[i * 16 + 10] = clock64();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t db = wdesc0 + static_cast<uint64_t>((DH2_W2_OFF + (k >> 2) * (128 * 128) + (k & 3) * 32) >> 4);
          umma_bf16_ts(X, tmem_base + DH2_A1A + k * 8, db, idesc_h, k > 0 ? 1u : 0u);
        }
        mbar_wait(a1b_ready, ph);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 8; k < 16; ++k) {
          const uint64_t db = wdesc0 + static_cast<uint64_t>((DH2_W2_OFF + (k >> 2) * (128 * 128) + (k & 3) * 32) >> 4);
          umma_bf16_ts(X, tmem_base + DH2_A0 + (k - 8) * 8, db, idesc_h, 1u);
        }
        umma_commit(acc1_ready);
        if (TRACE && cta == 0 && i < 8) p.trace[i * 16 + 11] = clock64();
      }
    }
  } else {
    const int quarter = warp & 3;
    const int slice = warp >> 2;
    const int trow = quarter * 32 + lane;                 // this thread's row inside a tile
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t T = tmem_base + lane_sel;
    const float hb0 = p.hb[0], hb1 = p.hb[1];
    // One tile ahead, in registers: this row's 64-feature slice of x0 (8 x 128 bits, lane-interleaved layout).  Nothing
    // reads it before it is stored to tensor memory.  Unconditional: x_il16 is padded to whole tiles.
    uint4 a0[8];
    float pd[NP];
#pragma unroll
    for (int t = 0; t < NP; ++t) pd[t] = 0.0f;
    auto load_next = [&](int tile_) {
      const uint4* src = reinterpret_cast<const uint4*>(p.x_il16) + ((static_cast<int64_t>(tile_) * 4 + quarter) * 32 + slice * 8) * 32 + lane;
#pragma unroll
      for (int e = 0; e < 8; ++e) a0[e] = __ldg(src + e * 32);
    };
    auto store_a0 = [&]() {
      uint32_t (&r0)[16] = reinterpret_cast<uint32_t (&)[16]>(a0[0]);
      uint32_t (&r1)[16] = reinterpret_cast<uint32_t (&)[16]>(a0[4]);
      __syncwarp();
      tmem_st_32x32_x16(T + DH2_A0 + slice * 32, r0);
      tmem_st_32x32_x16(T + DH2_A0 + slice * 32 + 16, r1);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(a0_ready);
    };
    // cross recurrence + heads of one finished tile: the slice-0 warp of each quarter, once the other three warps of the
    // quarter have published their partial dots (named barrier 1 + quarter, 128 threads)
    auto finish = [&](int tile_) {
      if (slice != 0) {   // bar.arrive orders this thread's earlier shared-memory writes before the barrier completes
        asm volatile("bar.arrive %0, 128;" ::"r"(1 + quarter) : "memory");
        return;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      float s[NP + 2];
#pragma unroll
      for (int t = 0; t < NP; ++t) s[t] = pd[t];
#pragma unroll
      for (int t = 0; t < 2; ++t)
        s[NP + t] = (part[(t * 4 + 0) * 128 + trow] + part[(t * 4 + 1) * 128 + trow]) +
                    (part[(t * 4 + 2) * 128 + trow] + part[(t * 4 + 3) * 128 + trow]);
      float a = 1.0f;
#pragma unroll
      for (int l = 0; l < LC; ++l) a += fmaf(a, s[l], const_s[l]);
      const float zr = fmaf(a, s[LC], const_s[LC]) + s[NP] + hb0;
      const float zc = fmaf(a, s[LC + 1], const_s[LC + 1]) + s[NP + 1] + hb1;
      const int64_t row_ = static_cast<int64_t>(tile_) * BM + trow;
      if (row_ < p.rows) {
        p.rating[row_] = zr;
        p.ctr[row_] = __fdividef(1.0f, 1.0f + __expf(-zc));   // MUFU only: an IEEE division / expf has a slow path that is an ABI
                                                                // call, and a call inside the epilogue spills everything live
      }
    };
    // x0 is the tower kernel's output: with a programmatic launch this CTA may have started before that kernel finished
    // (no-op otherwise).  Everything above -- barriers, tensor memory, the issuer's kernel loads -- overlapped its tail.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (n_my > 0) {
      load_next(cta);
      store_a0();
    }
    for (int i = 0; i < n_my; ++i) {
      const int tile = cta + i * ncta;
      const uint32_t ph = static_cast<uint32_t>(i & 1);
      const uint32_t X = T + (ph ? DH2_RB : DH2_RA), Y = T + (ph ? DH2_RA : DH2_RB);
      const bool tr = TRACE && cta == 0 && i < 8 && warp == 0 && lane == 0;
      auto stamp = [&](int k) { if (tr) p.trace[i * 16 + k] = clock64(); };
      // ---- epilogues a / b (one code path): 128 deep columns each -> bias, relu, bf16 -> A1a / A1b.  A1b lands on the
      //      first half of A0 (L0b has completed).  Epilogue a of the slice-0 warps also takes this row's x0 dots. ----
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        chain_wait(half ? acc_b : acc_a, ph);
        tc_fence_after_sync();
        stamp(2 * half);
        {
          float v[32];
          __syncwarp();
          tmem_ld_32x32((half ? Y : X) + slice * 32, v);
          const float* bs = p.biasv + half * 128 + slice * 32;   // constant bank
          uint32_t r[16];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
            r[e / 2] = pack_bf16x2_relu(v[e] + b4.x, v[e + 1] + b4.y);
            r[e / 2 + 1] = pack_bf16x2_relu(v[e + 2] + b4.z, v[e + 3] + b4.w);
          }
          tmem_st_32x32_x16(T + (half ? DH2_A0 : DH2_A1A) + slice * 16, r);
        }
        if (half == 0 && slice == 0) {
          float d[16];
          tmem_ld_32x32_x16(X + 128, d);
#pragma unroll
          for (int t = 0; t < NP; ++t) pd[t] = d[2 * t] + d[2 * t + 1];
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(half ? a1b_ready : a1a_ready);
        stamp(2 * half + 1);
      }
      // ---- the next tile's rows: issued here, landing while layer 1 runs (~900 cycles; the rows were prefetched into L2 a
      //      tile earlier), so that the 32 registers are live only across the wait below -- not across the epilogues,
      //      where every register is needed (this kernel has no L1 to spill into) ----
      if (i + 1 < n_my) load_next(tile + ncta);
      if (i + 2 < n_my && warp == 0 && lane == 0) {
        const int64_t t2 = tile + 2 * ncta;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.x_il16 + t2 * (BM * 256)), "r"(BM * 256 * 2) : "memory");
      }
      // ---- layer 1 done: hand the next tile's A0 to the tensor pipe first (its L0 then runs under this tile's last
      //      epilogue), then finish this tile on the CUDA cores ----
      chain_wait(acc1_ready, ph);
      tc_fence_after_sync();
      stamp(4);
      if (i + 1 < n_my) store_a0();
      stamp(5);
      {
        const int col0 = slice * 32;
        float v[32];
        __syncwarp();
        tmem_ld_32x32(X + col0, v);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc1_read);
        const float* bs = p.biasv + 256 + col0;   // constant bank, like the two head rows
        const float* wr = p.hdv + col0;
        const float* wc = p.hdv + 128 + col0;
        // packed fp32 arithmetic (FADD2 / FFMA2, two lanes per instruction): this loop is 3 fma-pipe operations per
        // accumulator element and was the longest phase of a tile (1.8 k cycles, RS_CHAIN_TRACE); each dot is kept as an
        // (even, odd) pair of partial sums
        float2 dr2 = make_float2(0.0f, 0.0f), dq2 = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
          const float4 r4 = *reinterpret_cast<const float4*>(wr + e);
          const float4 c4 = *reinterpret_cast<const float4*>(wc + e);
          float2 y01 = __fadd2_rn(make_float2(v[e], v[e + 1]), make_float2(b4.x, b4.y));
          float2 y23 = __fadd2_rn(make_float2(v[e + 2], v[e + 3]), make_float2(b4.z, b4.w));
          y01.x = fmaxf(y01.x, 0.0f); y01.y = fmaxf(y01.y, 0.0f);
          y23.x = fmaxf(y23.x, 0.0f); y23.y = fmaxf(y23.y, 0.0f);
          dr2 = __ffma2_rn(y01, make_float2(r4.x, r4.y), dr2);
          dq2 = __ffma2_rn(y01, make_float2(c4.x, c4.y), dq2);
          dr2 = __ffma2_rn(y23, make_float2(r4.z, r4.w), dr2);
          dq2 = __ffma2_rn(y23, make_float2(c4.z, c4.w), dq2);
        }
        const float dr = dr2.x + dr2.y, dq = dq2.x + dq2.y;
        part[(0 * 4 + slice) * 128 + trow] = dr;
        part[(1 * 4 + slice) * 128 + trow] = dq;
      }
      stamp(6);
      // `part` is re-written by the next tile's last epilogue, which runs behind a1b_ready of that tile: every warp,
      // the slice-0 warps after their finish() below, has arrived there by then
      finish(tile);
      stamp(7);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == DH2_ISSUER) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

bool use_deep_v2() {
  static const bool on = [] { const char* e = getenv("RS_DEEP_V2"); return !(e && e[0] == '0'); }();
  return on;
}

template <int LC>
void launch_deep_heads2(const ChainMaps& maps, const DeepHeadsParams& p, cudaStream_t st) {
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(deep_heads2_kernel<LC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, deep_heads2_smem<LC>());
    cudaFuncSetAttribute(deep_heads2_kernel<LC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, deep_heads2_smem<LC>());
  });
  const int64_t ntiles = (p.rows + BM - 1) / BM;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ntiles, rs::sm_count())));
  rs::ProfScope prof(RS_PROF_DEEP, st);
  if (p.trace) {
    deep_heads2_kernel<LC, true><<<grid, DH2_THREADS, deep_heads2_smem<LC>(), st>>>(maps, p);
    RS_LAUNCH_CHECK();
    return;
  }
  // programmatic dependent launch behind the tower kernel (RS_PDL=0: plain stream order)
  static const bool pdl = [] { const char* e = getenv("RS_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(DH2_THREADS);
  cfg.dynamicSmemBytes = deep_heads2_smem<LC>();
  cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl ? 1 : 0;
  RS_CUDA(cudaLaunchKernelEx(&cfg, deep_heads2_kernel<LC, false>, maps, p));
  RS_LAUNCH_CHECK();
}

bool use_fused_heads() {
  const char* e = getenv("RS_FUSED_HEADS");
  return !(e && e[0] == '0');
}

}  // namespace

struct rs_ranker {
  int device = 0;
  int cross_layers = 0;
  const float* table[2] = {nullptr, nullptr};   // borrowed
  uint16_t* table16[2] = {nullptr, nullptr};    // owned bf16 copies of small tables (what the tower kernel gathers from)
  int64_t table_ld[2] = {0, 0}, table_rows[2] = {0, 0};
  uint16_t* wts = nullptr;      // all bf16 kernels, [out][in] each: user tower, item tower, deep
  float* f32 = nullptr;         // biases (user 576, item 576, deep 384), cross_w [L,256], cross_b [L,256], head_w [2,384], head_b [2]
  size_t off_w[2][4], off_dw[2], off_dotw;
  size_t off_b[2][4], off_db[2], off_cw, off_cb, off_hw, off_hb, off_const;
  ChainMaps tower_maps, deep_maps;
  float h_tower_bias[2][576] = {};   // host copy of the tower biases: tower_sp_kernel takes them as kernel parameters
  float h_deep_bias[384] = {};       // likewise for deep_heads2_kernel: deep biases and the deep part of the head rows
  float h_hd[256] = {};
  ~rs_ranker() {
    cudaFree(wts);
    cudaFree(f32);
    cudaFree(table16[0]);
    cudaFree(table16[1]);
  }
};

namespace {

// tables up to this many rows are copied to bf16 when the handle is created (<= 256 MB per table)
constexpr int64_t kShadowTableRows = 1 << 20;

__global__ void table_to_bf16_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, uint16_t* __restrict__ dst) {
  const int64_t n = rows * 64;   // bf16 pairs
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i >> 6;
    const int c = static_cast<int>(i & 63) * 2;
    const float2 f = *reinterpret_cast<const float2*>(src + r * ld + c);
    reinterpret_cast<uint32_t*>(dst)[i] = pack_bf16x2(f.x, f.y);
  }
}

void copy_in(void* dst, const DLTensor* t, size_t bytes, cudaStream_t st) {
  RS_CUDA(cudaMemcpyAsync(dst, rs::data_ptr<void>(t), bytes, cudaMemcpyDeviceToDevice, st));
}

}  // namespace

extern "C" int rs_ranker_create(const DLTensor* user_table, const DLTensor* item_table, const DLTensor* const* tower_wt,
                                const DLTensor* const* tower_b, const DLTensor* const* deep_wt,
                                const DLTensor* const* deep_b, const DLTensor* cross_w, const DLTensor* cross_b,
                                const DLTensor* head_w, const DLTensor* head_b, rs_ranker** out) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(out != nullptr, "out is NULL");
    sm_count();
    auto h = std::unique_ptr<rs_ranker>(new rs_ranker());
    RS_CUDA(cudaGetDevice(&h->device));
    const DLTensor* tabs[2] = {user_table, item_table};
    for (int s = 0; s < 2; ++s) {
      require(tabs[s], "table", 2, kDLFloat, 32);
      RS_CHECK(on_device(tabs[s]) && tabs[s]->shape[1] == 128, "embedding tables must be CUDA [V, 128] tensors");
      h->table[s] = data_ptr<float>(tabs[s]);
      h->table_ld[s] = row_stride_2d(tabs[s], "table");
      h->table_rows[s] = tabs[s]->shape[0];
      RS_CHECK(h->table_ld[s] % 4 == 0, "table row stride must be a multiple of 4");
      require_aligned16(h->table[s], "table");
    }
    // sizes
    size_t wb = 0, fb = 0;
    for (int s = 0; s < 2; ++s)
      for (int l = 0; l < 4; ++l) {
        h->off_w[s][l] = wb;
        wb += (size_t)CfgTower::N(l) * CfgTower::K(l);
        h->off_b[s][l] = fb;
        fb += CfgTower::N(l);
      }
    for (int l = 0; l < 2; ++l) {
      h->off_dw[l] = wb;
      wb += (size_t)CfgDeep::N(l) * CfgDeep::K(l);
      h->off_db[l] = fb;
      fb += CfgDeep::N(l);
    }
    h->off_dotw = wb;
    wb += (size_t)DOTW_ROWS * 256;
    require(cross_w, "cross_w", 2, kDLFloat, 32);
    require(cross_b, "cross_b", 2, kDLFloat, 32);
    require(head_w, "head_w", 2, kDLFloat, 32);
    require(head_b, "head_b", 1, kDLFloat, 32);
    const int64_t L = cross_w->shape[0];
    RS_CHECK(L >= 0 && L <= 4 && cross_w->shape[1] == 256 && cross_b->shape[0] == L && cross_b->shape[1] == 256 &&
                 is_compact(cross_w) && is_compact(cross_b), "cross_w / cross_b must be contiguous [n_layers, 256]");
    RS_CHECK(head_w->shape[0] == 2 && head_w->shape[1] == 384 && is_compact(head_w) && head_b->shape[0] == 2,
             "head_w must be contiguous [2, 384], head_b [2]");
    h->cross_layers = static_cast<int>(L);
    h->off_cw = fb; fb += (size_t)L * 256;
    h->off_cb = fb; fb += (size_t)L * 256;
    h->off_hw = fb; fb += 2 * 384;
    h->off_hb = fb; fb += 4;
    h->off_const = fb; fb += 8;
    RS_CUDA(cudaMalloc(&h->wts, wb * sizeof(uint16_t)));
    RS_CUDA(cudaMalloc(&h->f32, fb * sizeof(float)));
    RS_CUDA(cudaMemset(h->f32, 0, fb * sizeof(float)));
    cudaStream_t st = nullptr;
    auto take_w = [&](const DLTensor* w, int N, int K, size_t off, CUtensorMap* map) {
      require(w, "wt", 2, kDLBfloat, 16);
      RS_CHECK(on_device(w) && is_compact(w) && w->shape[0] == N && w->shape[1] == K,
               "the fused ranker supports towers 128-256-128-64-128 and deep net 256-256-128 (kernels as bf16 [out, in])");
      copy_in(h->wts + off, w, (size_t)N * K * 2, st);
      *map = make_tmap_2d(h->wts + off, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)N, (uint64_t)K, (uint64_t)K * 2,
                          static_cast<uint32_t>(N), 64);
    };
    auto take_b = [&](const DLTensor* b, int N, size_t off) {
      if (!b) return;
      require(b, "bias", 1, kDLFloat, 32);
      RS_CHECK(on_device(b) && b->shape[0] == N && (N <= 1 || stride_of(b, 0) == 1), "bias must be a contiguous CUDA [out] tensor");
      copy_in(h->f32 + off, b, (size_t)N * 4, st);
    };
    for (int s = 0; s < 2; ++s)
      for (int l = 0; l < 4; ++l) {
        take_w(tower_wt[s * 4 + l], CfgTower::N(l), CfgTower::K(l), h->off_w[s][l], &h->tower_maps.w[s][l]);
        take_b(tower_b ? tower_b[s * 4 + l] : nullptr, CfgTower::N(l), h->off_b[s][l]);
      }
    for (int l = 0; l < 2; ++l) {
      take_w(deep_wt[l], CfgDeep::N(l), CfgDeep::K(l), h->off_dw[l], &h->deep_maps.w[0][l]);
      take_b(deep_b ? deep_b[l] : nullptr, CfgDeep::N(l), h->off_db[l]);
    }
    // deep_heads2_kernel loads layer 0's kernel as two 128-row halves (the 16 dot rows sit between them in shared memory)
    h->deep_maps.w[1][0] = make_tmap_2d(h->wts + h->off_dw[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 256, 256, 256 * 2, 128, 64);
    RS_CHECK(on_device(cross_w) && on_device(cross_b) && on_device(head_w) && on_device(head_b), "weights must be CUDA tensors");
    if (L > 0) {
      copy_in(h->f32 + h->off_cw, cross_w, (size_t)L * 256 * 4, st);
      copy_in(h->f32 + h->off_cb, cross_b, (size_t)L * 256 * 4, st);
    }
    copy_in(h->f32 + h->off_hw, head_w, 2 * 384 * 4, st);
    copy_in(h->f32 + h->off_hb, head_b, 2 * 4, st);
    for (int s = 0; s < 2; ++s) {
      if (h->table_rows[s] < 1 || h->table_rows[s] > kShadowTableRows) continue;
      RS_CUDA(cudaMalloc(&h->table16[s], sizeof(uint16_t) * (size_t)h->table_rows[s] * 128));
      const int blocks = static_cast<int>(std::min<int64_t>((h->table_rows[s] * 64 + 255) / 256, (int64_t)sm_count() * 8));
      table_to_bf16_kernel<<<blocks, 256, 0, st>>>(h->table[s], h->table_ld[s], h->table_rows[s], h->table16[s]);
      RS_LAUNCH_CHECK();
    }
    ranker_dotw_kernel<<<1, 256, 0, st>>>(h->f32 + h->off_cw, h->f32 + h->off_hw, static_cast<int>(L), h->wts + h->off_dotw);
    RS_LAUNCH_CHECK();
    h->deep_maps.w[0][2] = make_tmap_2d(h->wts + h->off_dotw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)DOTW_ROWS, 256,
                                        256 * 2, DOTW_ROWS, 64);
    ranker_consts_kernel<<<1, 32, 0, st>>>(h->f32 + h->off_cw, h->f32 + h->off_cb, h->f32 + h->off_hw, static_cast<int>(L),
                                           h->f32 + h->off_const);
    RS_LAUNCH_CHECK();
    RS_CUDA(cudaStreamSynchronize(st));
    for (int sdx = 0; sdx < 2; ++sdx)
      RS_CUDA(cudaMemcpy(h->h_tower_bias[sdx], h->f32 + h->off_b[sdx][0], sizeof(float) * 576, cudaMemcpyDeviceToHost));
    RS_CUDA(cudaMemcpy(h->h_deep_bias, h->f32 + h->off_db[0], sizeof(float) * 384, cudaMemcpyDeviceToHost));
    for (int t = 0; t < 2; ++t)
      RS_CUDA(cudaMemcpy(h->h_hd + t * 128, h->f32 + h->off_hw + t * 384 + 256, sizeof(float) * 128, cudaMemcpyDeviceToHost));
    *out = h.release();
  });
}

extern "C" int rs_ranker_free(rs_ranker* h) {
  return rs::guard([&] {
    if (!h) return;
    delete h;
  });
}

// user_rows / item_rows: int32|int64 [B]; combined: f32 [B, 256] (user embedding | item embedding, the DCN input);
// rating, ctr: f32 [B, 1]
static int ranker_forward_impl(rs_ranker* h, const DLTensor* user_rows, const DLTensor* item_rows,
                               const DLTensor* user_emb, const DLTensor* item_emb, DLTensor* combined, DLTensor* rating,
                               DLTensor* ctr, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "ranker is NULL");
    const DLTensor* ids[2] = {user_rows, item_rows};
    const DLTensor* embs[2] = {user_emb, item_emb};
    require(combined, "combined", 2, kDLFloat, 32);
    require(rating, "rating", 2, kDLFloat, 32);
    require(ctr, "ctr", 2, kDLFloat, 32);
    const int64_t B = combined->shape[0];
    RS_CHECK(on_device(combined) && combined->shape[1] == 256 && is_compact(combined), "combined must be a contiguous CUDA [B, 256] tensor");
    RS_CHECK(on_device(rating) && on_device(ctr) && rating->shape[0] == B && ctr->shape[0] == B && rating->shape[1] == 1 &&
                 ctr->shape[1] == 1 && is_compact(rating) && is_compact(ctr), "rating / ctr must be contiguous CUDA [B, 1] tensors");
    require_aligned16(data_ptr<float>(combined), "combined");
    if (B == 0) return;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint16_t* x_bf16 = nullptr;   // unfused path only: [B, 256] bf16 copy of the tower outputs, the deep net's input
    float* deep_out = nullptr;    // unfused path only: [B, 128]
    const bool pp = use_tower_pp();
    const bool fused = use_fused_heads() && pp;   // default: deep chain + cross + heads in one kernel (needs tower_pp's extra outputs)
    const int64_t tiles = (B + BM - 1) / BM;
    uint16_t* x_il16 = nullptr;   // fused path: bf16 x0, lane-interleaved       (ChainProblem::y_il16)
    if (fused) {
      RS_CUDA(cudaMallocAsync(&x_il16, sizeof(uint16_t) * (size_t)tiles * BM * 256, st));
    }
    if (!fused) {
      RS_CUDA(cudaMallocAsync(&x_bf16, sizeof(uint16_t) * (size_t)B * 256, st));
      RS_CUDA(cudaMallocAsync(&deep_out, sizeof(float) * (size_t)B * 128, st));
    }
    try {
      ChainParams p{};
      p.n_prob = 2;
      p.act[0] = p.act[1] = p.act[2] = RS_ACT_RELU;
      p.act[3] = RS_ACT_NONE;
      for (int s = 0; s < 2; ++s) {
        ChainProblem& pr = p.prob[s];
        pr.rows = B;
        for (int l = 0; l < 4; ++l) pr.bias[l] = h->f32 + h->off_b[s][l];
        if (embs[s]) {   // already-looked-up embedding rows (sharded tables): identity gather
          require(embs[s], "embedding rows", 2, kDLFloat, 32);
          RS_CHECK(on_device(embs[s]) && embs[s]->shape[0] == B && embs[s]->shape[1] == 128, "embedding rows must be CUDA [B, 128]");
          pr.table = data_ptr<float>(embs[s]);
          pr.table_ld = row_stride_2d(embs[s], "embedding rows");
          pr.table_rows = B;
          RS_CHECK(pr.table_ld % 4 == 0, "embedding row stride must be a multiple of 4");
          require_aligned16(pr.table, "embedding rows");
          pr.ids = nullptr;
        } else {
          RS_CHECK(ids[s] && ids[s]->ndim == 1 && ids[s]->shape[0] == B && on_device(ids[s]) &&
                       (is_dtype(ids[s], kDLInt, 32) || is_dtype(ids[s], kDLInt, 64)) && (B <= 1 || stride_of(ids[s], 0) == 1),
                   "row ids must be contiguous CUDA int32/int64 [B] tensors");
          pr.table = h->table[s];
          pr.table16 = h->table16[s];
          pr.table_ld = h->table_ld[s];
          pr.table_rows = h->table_rows[s];
          pr.ids = data_ptr<void>(ids[s]);
          pr.ids64 = is_dtype(ids[s], kDLInt, 64) ? 1 : 0;
        }
        pr.y_f32 = data_ptr<float>(combined) + s * 128;
        pr.ld_f32 = 256;
        pr.y_bf16 = fused ? nullptr : x_bf16 + s * 128;
        pr.ld_bf16 = 256;
        pr.y_il16 = x_il16;
        pr.il_chunk0 = s * 16;
      }
      if (pp) {
        std::memcpy(p.biasv, h->h_tower_bias, sizeof(p.biasv));
        ChainMaps maps = h->tower_maps;
        for (int s = 0; s < 2; ++s) {
          maps.y[s] = make_tmap_2d(p.prob[s].y_f32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)B, 128, 256 * 4, 16, 32);
          p.prob[s].y_tma = 1;
        }
        launch_tower_pp(maps, p, st);
      } else {
        launch_chain<CfgTower>(h->tower_maps, p, st);
      }
      if (fused) {
        DeepHeadsParams q{};
        q.rows = B;
        q.x_il16 = x_il16;
        for (int l = 0; l < 2; ++l) q.bias[l] = h->f32 + h->off_db[l];
        q.hw = h->f32 + h->off_hw;
        q.hb = h->f32 + h->off_hb;
        q.consts = h->f32 + h->off_const;
        q.rating = data_ptr<float>(rating);
        q.ctr = data_ptr<float>(ctr);
        std::memcpy(q.biasv, h->h_deep_bias, sizeof(q.biasv));
        std::memcpy(q.hdv, h->h_hd, sizeof(q.hdv));
        const char* tr_env = getenv("RS_CHAIN_TRACE");
        if (tr_env && tr_env[0] == '1') {
          RS_CUDA(cudaMalloc(&q.trace, 8 * 16 * sizeof(long long)));
          RS_CUDA(cudaMemsetAsync(q.trace, 0, 8 * 16 * sizeof(long long), st));
        }
        const bool v2 = use_deep_v2();
        switch (h->cross_layers) {
          case 0: v2 ? launch_deep_heads2<0>(h->deep_maps, q, st) : launch_deep_heads<0>(h->deep_maps, q, st); break;
          case 1: v2 ? launch_deep_heads2<1>(h->deep_maps, q, st) : launch_deep_heads<1>(h->deep_maps, q, st); break;
          case 2: v2 ? launch_deep_heads2<2>(h->deep_maps, q, st) : launch_deep_heads<2>(h->deep_maps, q, st); break;
          case 3: v2 ? launch_deep_heads2<3>(h->deep_maps, q, st) : launch_deep_heads<3>(h->deep_maps, q, st); break;
          case 4: v2 ? launch_deep_heads2<4>(h->deep_maps, q, st) : launch_deep_heads<4>(h->deep_maps, q, st); break;
          default: RS_CHECK(false, "the fused ranker supports up to 4 cross layers");
        }
        cudaFreeAsync(x_il16, st);
        if (q.trace) {   // debugging aid: per-tile phase stamps of CTA 0, relative to the first one
          long long host[8 * 16];
          RS_CUDA(cudaStreamSynchronize(st));
          RS_CUDA(cudaMemcpy(host, q.trace, sizeof(host), cudaMemcpyDeviceToHost));
          cudaFree(q.trace);
          for (int t = 0; t < 8 && host[t * 16]; ++t) {
            fprintf(stderr, "deep_heads tile %d:", t);
            for (int k = 0; k < 13; ++k) fprintf(stderr, " %lld", host[t * 16 + k] ? host[t * 16 + k] - host[0] : -1);
            fprintf(stderr, "\n");
          }
        }
        return;
      }
      ChainParams d{};
      d.n_prob = 1;
      d.act[0] = d.act[1] = RS_ACT_RELU;
      d.prob[0].rows = B;
      for (int l = 0; l < 2; ++l) d.prob[0].bias[l] = h->f32 + h->off_db[l];
      d.prob[0].x = x_bf16;
      d.prob[0].x_ld = 256;
      d.prob[0].y_f32 = deep_out;
      d.prob[0].ld_f32 = 128;
      launch_chain<CfgDeep>(h->deep_maps, d, st);
      const int64_t want = (B * 32 + 255) / 256;
      const int blocks = static_cast<int>(std::min<int64_t>(want, (int64_t)sm_count() * 8));
      const float* x0 = data_ptr<float>(combined);
      const float* cwp = h->f32 + h->off_cw;
      const float* cbp = h->f32 + h->off_cb;
      const float* hwp = h->f32 + h->off_hw;
      const float* hbp = h->f32 + h->off_hb;
      float* rp = data_ptr<float>(rating);
      float* cp = data_ptr<float>(ctr);
      switch (h->cross_layers) {
        case 0: cross_heads_kernel<0><<<blocks, 256, 0, st>>>(x0, 256, deep_out, 128, cwp, cbp, hwp, hbp, rp, cp, B); break;
        case 1: cross_heads_kernel<1><<<blocks, 256, 0, st>>>(x0, 256, deep_out, 128, cwp, cbp, hwp, hbp, rp, cp, B); break;
        case 2: cross_heads_kernel<2><<<blocks, 256, 0, st>>>(x0, 256, deep_out, 128, cwp, cbp, hwp, hbp, rp, cp, B); break;
        case 3: cross_heads_kernel<3><<<blocks, 256, 0, st>>>(x0, 256, deep_out, 128, cwp, cbp, hwp, hbp, rp, cp, B); break;
        case 4: cross_heads_kernel<4><<<blocks, 256, 0, st>>>(x0, 256, deep_out, 128, cwp, cbp, hwp, hbp, rp, cp, B); break;
        default: RS_CHECK(false, "the fused ranker supports up to 4 cross layers");
      }
      RS_LAUNCH_CHECK();
    } catch (...) {
      if (x_il16) cudaFreeAsync(x_il16, st);
      if (x_bf16) cudaFreeAsync(x_bf16, st);
      if (deep_out) cudaFreeAsync(deep_out, st);
      throw;
    }
    if (x_bf16) cudaFreeAsync(x_bf16, st);
    if (deep_out) cudaFreeAsync(deep_out, st);
  });
}

extern "C" int rs_ranker_forward(rs_ranker* h, const DLTensor* user_rows, const DLTensor* item_rows, DLTensor* combined,
                                 DLTensor* rating, DLTensor* ctr, void* stream) {
  return ranker_forward_impl(h, user_rows, item_rows, nullptr, nullptr, combined, rating, ctr, stream);
}

// same forward from already-looked-up embedding rows f32 [B,128] (the output of a sharded-table lookup)
extern "C" int rs_ranker_forward_rows(rs_ranker* h, const DLTensor* user_emb, const DLTensor* item_emb, DLTensor* combined,
                                      DLTensor* rating, DLTensor* ctr, void* stream) {
  return ranker_forward_impl(h, nullptr, nullptr, user_emb, item_emb, combined, rating, ctr, stream);
}

// K3a + K3b without the [B, 384] concat (src/models.py:38-44,50,119-120,131): DCN-v1 cross stack on x0 and both heads on
// (cross output | deep output), one warp per row, fp32 FMA arithmetic in the reference's order (bias outside the product).
extern "C" int rs_cross_heads_forward(const DLTensor* x0, const DLTensor* deep, const DLTensor* cross_w, const DLTensor* cross_b,
                                      const DLTensor* head_w, const DLTensor* head_b, DLTensor* rating, DLTensor* ctr, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(x0, "x0", 2, kDLFloat, 32);
    require(deep, "deep", 2, kDLFloat, 32);
    require(cross_w, "cross_w", 2, kDLFloat, 32);
    require(cross_b, "cross_b", 2, kDLFloat, 32);
    require(head_w, "head_w", 2, kDLFloat, 32);
    require(head_b, "head_b", 1, kDLFloat, 32);
    RS_CHECK(rating != nullptr && ctr != nullptr && is_dtype(rating, kDLFloat, 32) && is_dtype(ctr, kDLFloat, 32), "rating / ctr must be f32 tensors");
    const int64_t B = x0->shape[0], L = cross_w->shape[0];
    RS_CHECK(on_device(x0) && on_device(deep) && on_device(cross_w) && on_device(cross_b) && on_device(head_w) && on_device(head_b) &&
                 on_device(rating) && on_device(ctr), "all tensors must be CUDA tensors");
    RS_CHECK(x0->shape[1] == 256 && deep->shape[0] == B && deep->shape[1] == 128, "x0 must be [B, 256] and deep [B, 128]");
    RS_CHECK(L <= 4 && (L == 0 || cross_w->shape[1] == 256) && cross_b->shape[0] == L && (L == 0 || cross_b->shape[1] == 256) &&
                 is_compact(cross_w) && is_compact(cross_b), "cross_w / cross_b must be contiguous [L <= 4, 256]");
    RS_CHECK(head_w->shape[0] == 2 && head_w->shape[1] == 384 && is_compact(head_w) && head_b->shape[0] == 2, "head_w [2, 384], head_b [2]");
    RS_CHECK(rating->shape[0] == B && ctr->shape[0] == B && is_compact(rating) && is_compact(ctr), "rating / ctr must be contiguous with B rows");
    const int64_t ldx = row_stride_2d(x0, "x0"), ldd = row_stride_2d(deep, "deep");
    RS_CHECK(ldx % 4 == 0 && ldd % 4 == 0, "row strides must be multiples of 4 floats");
    require_aligned16(data_ptr<float>(x0), "x0");
    require_aligned16(data_ptr<float>(deep), "deep");
    if (L > 0) {
      require_aligned16(data_ptr<float>(cross_w), "cross_w");
      require_aligned16(data_ptr<float>(cross_b), "cross_b");
    }
    require_aligned16(data_ptr<float>(head_w), "head_w");
    if (B == 0) return;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t want = (B * 32 + 255) / 256;
    const int blocks = static_cast<int>(std::min<int64_t>(want, (int64_t)sm_count() * 8));
    const float *xp = data_ptr<float>(x0), *dp = data_ptr<float>(deep), *cwp = data_ptr<float>(cross_w), *cbp = data_ptr<float>(cross_b);
    const float *hwp = data_ptr<float>(head_w), *hbp = data_ptr<float>(head_b);
    float *rp = data_ptr<float>(rating), *cp = data_ptr<float>(ctr);
    switch (L) {
      case 0: cross_heads_kernel<0><<<blocks, 256, 0, st>>>(xp, ldx, dp, ldd, cwp, cbp, hwp, hbp, rp, cp, B); break;
      case 1: cross_heads_kernel<1><<<blocks, 256, 0, st>>>(xp, ldx, dp, ldd, cwp, cbp, hwp, hbp, rp, cp, B); break;
      case 2: cross_heads_kernel<2><<<blocks, 256, 0, st>>>(xp, ldx, dp, ldd, cwp, cbp, hwp, hbp, rp, cp, B); break;
      case 3: cross_heads_kernel<3><<<blocks, 256, 0, st>>>(xp, ldx, dp, ldd, cwp, cbp, hwp, hbp, rp, cp, B); break;
      default: cross_heads_kernel<4><<<blocks, 256, 0, st>>>(xp, ldx, dp, ldd, cwp, cbp, hwp, hbp, rp, cp, B); break;
    }
    RS_LAUNCH_CHECK();
  });
}
