// dense_tc.cu -- K2/K3 (bf16 tensor mode): Keras Dense on tcgen05, y = act(x @ kernel + bias).
// Replaces keras.layers.Dense in user_tower / item_tower / deep_nets (src/models.py:26-29,76-77) at
// ranking batch sizes, where the layers really are dense contractions (B x in x out GEMMs).
//
// Weight-stationary persistent kernel, one CTA per SM:
//   * the layer's kernel, transposed to [out][in] bf16 (K-major B operand), is TMA-loaded ONCE per CTA and
//     stays in shared memory (<= 128 KB for the 256x256 layer);
//   * activation tiles of 128 rows stream through a ring of 16 KB [128 x 64] bf16 chunks (TMA, 128B swizzle,
//     full/empty mbarriers per chunk), so the ring depth does not depend on the layer's width;
//   * one thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> f32, M128 x N<=256 x K16) into one of two TMEM
//     accumulators; 16 epilogue warps read the other accumulator (tcgen05.ld), add the bias, apply the
//     activation and store bf16 (the next layer's input) and/or fp32 (the model's outputs, possibly a
//     column slice of a wider buffer -- this is how the towers write into the concatenated DCN input).
// Per-sample work: 2*in*out flop; bytes: in*2 read + out*2 (bf16) / out*4 (fp32) written.  Between layers the
// bf16 activations of a 65,536 batch (<= 32 MB) stay L2-resident (126 MB L2), so the chain is bound by L2
// bandwidth and the tensor pipe rather than HBM.
#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"

namespace {

using namespace rsd;

constexpr int BM = 128;
constexpr int CHUNK_BYTES = BM * 128;      // [128 rows x 64 bf16] = 16 KB
constexpr int EPI_WARPS = 16;
constexpr int NTHREADS = 128 + EPI_WARPS * 32;

struct DenseParams {
  int64_t rows;          // B
  int act;
  const float* bias;     // [N] or nullptr
  uint16_t* y_bf16;      // [B, N] row stride ld_bf16 (elements) or nullptr
  int64_t ld_bf16;
  float* y_f32;          // [B, N] row stride ld_f32 (elements) or nullptr
  int64_t ld_f32;
};

template <int K, int N>
struct DenseCfg {
  static constexpr int KC = K / 64;                         // 64-wide k chunks
  static constexpr int W_BYTES = N * K * 2;
  static constexpr int kAux = 1024 /*bias*/ + 512 /*barriers*/ + 1024 /*align*/;
  static constexpr int kFit = (232448 - kAux - W_BYTES) / CHUNK_BYTES;    // 227 KB of shared memory per CTA
  static constexpr int RING = kFit >= 8 ? 8 : kFit;
  static constexpr int SLICES = N / 32 >= 4 ? 4 : N / 32;   // column slices handled by different warps
  static constexpr int SLICE_COLS = N / SLICES;
  static constexpr int SMEM = W_BYTES + RING * CHUNK_BYTES + kAux;
  static_assert(RING >= KC + 1, "activation ring too shallow");
};

template <int K, int N>
__global__ void __launch_bounds__(NTHREADS, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const DenseParams p) {
  using C = DenseCfg<K, N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = base;                                   // KC chunks of [N rows x 128 B]
  uint8_t* xsm = base + C::W_BYTES;                      // RING chunks of [128 rows x 128 B]
  float* bias_s = reinterpret_cast<float*>(xsm + C::RING * CHUNK_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 256);
  uint64_t* wfull = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = full + C::RING;
  uint64_t* tfull = empty + C::RING;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t ntiles = (p.rows + BM - 1) / BM;
  const int64_t n_my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  constexpr uint32_t TMEM_COLS = (2 * N <= 32) ? 32 : (2 * N <= 64) ? 64 : (2 * N <= 128) ? 128 : (2 * N <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    mbar_init(wfull, 1);
    for (int s = 0; s < C::RING; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], C::SLICES * 4);
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < N; i += NTHREADS) bias_s[i] = p.bias ? p.bias[i] : 0.0f;
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ---- TMA producer: the weights once, then activation chunks ----
      mbar_arrive_expect_tx(wfull, C::W_BYTES);
      for (int kc = 0; kc < C::KC; ++kc) tma_load_2d(wsm + kc * (N * 128), &tmap_w, wfull, kc * 64, 0, kEvictLast);
      uint32_t it = 0;
      for (int64_t i = 0; i < n_my; ++i) {
        const int64_t tile = blockIdx.x + i * gridDim.x;
        for (int kc = 0; kc < C::KC; ++kc, ++it) {
          const uint32_t s = it % C::RING, ph = (it / C::RING) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], CHUNK_BYTES);
          tma_load_2d(xsm + s * CHUNK_BYTES, &tmap_x, &full[s], kc * 64, static_cast<int>(tile * BM), kEvictFirst);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---- MMA issuer ----
      constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, N);
      mbar_wait(wfull, 0);
      const uint64_t wdesc0 = umma_desc_k_sw128(smem_u32(wsm)), xdesc0 = umma_desc_k_sw128(smem_u32(xsm));
      uint32_t it = 0;
      for (int64_t i = 0; i < n_my; ++i) {
        const uint32_t a = static_cast<uint32_t>(i) & 1u, aph = (static_cast<uint32_t>(i) >> 1) & 1u;
        mbar_wait(&tempty[a], aph ^ 1u);
        tc_fence_after_sync();
        for (int kc = 0; kc < C::KC; ++kc, ++it) {
          const uint32_t s = it % C::RING, ph = (it / C::RING) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t da = xdesc0 + static_cast<uint64_t>((s * CHUNK_BYTES + k4 * 32) >> 4);
            const uint64_t db = wdesc0 + static_cast<uint64_t>((kc * (N * 128) + k4 * 32) >> 4);
            umma_bf16_ss(tmem_base + a * N, da, db, idesc, (kc | k4) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[a]);
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: TMEM -> bias + activation -> bf16 / fp32 stores ----
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int slice = ew >> 2;
    if (slice < C::SLICES) {
      for (int64_t i = 0; i < n_my; ++i) {
        const int64_t tile = blockIdx.x + i * gridDim.x;
        const int64_t row = tile * BM + quarter * 32 + lane;
        const uint32_t a = static_cast<uint32_t>(i) & 1u, aph = (static_cast<uint32_t>(i) >> 1) & 1u;
        mbar_wait(&tfull[a], aph);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + a * N + slice * C::SLICE_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < C::SLICE_COLS / 32; ++c) {
          float v[32];
          __syncwarp();
          tmem_ld_32x32(taddr + c * 32, v);
          const int col0 = slice * C::SLICE_COLS + c * 32;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + col0 + e);
            v[e] += b4.x;
            v[e + 1] += b4.y;
            v[e + 2] += b4.z;
            v[e + 3] += b4.w;
          }
          if (p.act == RS_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.0f);
          } else if (p.act == RS_ACT_SIGMOID) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = 1.0f / (1.0f + __expf(-v[e]));
          }
          if (row < p.rows) {
            if (p.y_bf16) {
              uint4* dst = reinterpret_cast<uint4*>(p.y_bf16 + row * p.ld_bf16 + col0);
#pragma unroll
              for (int e = 0; e < 32; e += 8)
                dst[e >> 3] = make_uint4(pack_bf16x2(v[e], v[e + 1]), pack_bf16x2(v[e + 2], v[e + 3]),
                                         pack_bf16x2(v[e + 4], v[e + 5]), pack_bf16x2(v[e + 6], v[e + 7]));
            }
            if (p.y_f32) {
              float4* dst = reinterpret_cast<float4*>(p.y_f32 + row * p.ld_f32 + col0);
#pragma unroll
              for (int e = 0; e < 32; e += 4) dst[e >> 2] = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
            }
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[a]);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// embedding rows (fp32 table) -> bf16 activations [B, E]; one thread per 8 features
__global__ void __launch_bounds__(256) gather_bf16_kernel(const float* __restrict__ table, int64_t ld, int64_t n_rows,
                                                          const void* __restrict__ ids, int ids64, uint16_t* __restrict__ out,
                                                          int64_t ld_out, int64_t batch, int chunks) {
  const int64_t total = batch * chunks;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / chunks;
    const int c = static_cast<int>(i - b * chunks);
    const int64_t row = ids64 ? static_cast<const int64_t*>(ids)[b] : static_cast<int64_t>(static_cast<const int32_t*>(ids)[b]);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (row >= 0 && row < n_rows) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(table + row * ld) + 2 * c);
      const float4 d = __ldg(reinterpret_cast<const float4*>(table + row * ld) + 2 * c + 1);
      o = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
    }
    *reinterpret_cast<uint4*>(out + b * ld_out + 8 * (int64_t)c) = o;
  }
}

template <int K, int N>
void launch_dense(const CUtensorMap& tx, const CUtensorMap& tw, const DenseParams& p, cudaStream_t st) {
  using C = DenseCfg<K, N>;
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(dense_tc_kernel<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  });
  const int64_t ntiles = (p.rows + BM - 1) / BM;
  const int grid = static_cast<int>(std::min<int64_t>(ntiles, rs::sm_count()));
  rs::ProfScope prof(RS_PROF_RANK, st);
  dense_tc_kernel<K, N><<<grid, NTHREADS, C::SMEM, st>>>(tx, tw, p);
  RS_LAUNCH_CHECK();
}

}  // namespace

extern "C" int rs_dense_forward_bf16(const DLTensor* x, const DLTensor* wt, const DLTensor* bias, int act,
                                     DLTensor* y_bf16, DLTensor* y_f32, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(x, "x", 2, kDLBfloat, 16);
    require(wt, "wt", 2, kDLBfloat, 16);
    RS_CHECK(on_device(x) && on_device(wt), "x and wt must be CUDA tensors");
    RS_CHECK(act == RS_ACT_NONE || act == RS_ACT_RELU || act == RS_ACT_SIGMOID, "unknown activation code");
    const int64_t B = x->shape[0], K = x->shape[1], N = wt->shape[0];
    RS_CHECK(wt->shape[1] == K && is_compact(wt), "wt must be a contiguous [out, in] bf16 tensor (the Keras kernel transposed)");
    RS_CHECK(y_bf16 || y_f32, "at least one output is required");
    DenseParams p{};
    p.rows = B;
    p.act = act;
    if (bias) {
      require(bias, "bias", 1, kDLFloat, 32);
      RS_CHECK(on_device(bias) && bias->shape[0] == N && (N <= 1 || stride_of(bias, 0) == 1), "bias must be a contiguous CUDA [out] tensor");
      p.bias = data_ptr<float>(bias);
    }
    if (y_bf16) {
      require(y_bf16, "y_bf16", 2, kDLBfloat, 16);
      RS_CHECK(on_device(y_bf16) && y_bf16->shape[0] == B && y_bf16->shape[1] == N, "y_bf16 must be a CUDA [B, out] tensor");
      p.y_bf16 = data_ptr<uint16_t>(y_bf16);
      p.ld_bf16 = row_stride_2d(y_bf16, "y_bf16");
      RS_CHECK(p.ld_bf16 % 8 == 0, "y_bf16 row stride must be a multiple of 8 elements");
      require_aligned16(p.y_bf16, "y_bf16");
    }
    if (y_f32) {
      require(y_f32, "y_f32", 2, kDLFloat, 32);
      RS_CHECK(on_device(y_f32) && y_f32->shape[0] == B && y_f32->shape[1] == N, "y_f32 must be a CUDA [B, out] tensor");
      p.y_f32 = data_ptr<float>(y_f32);
      p.ld_f32 = row_stride_2d(y_f32, "y_f32");
      RS_CHECK(p.ld_f32 % 4 == 0, "y_f32 row stride must be a multiple of 4 elements");
      require_aligned16(p.y_f32, "y_f32");
    }
    if (B == 0) return;
    const int64_t ldx = row_stride_2d(x, "x");
    RS_CHECK(ldx % 8 == 0, "x row stride must be a multiple of 8 elements");
    sm_count();
    const CUtensorMap tx = make_tmap_2d(data_ptr<void>(x), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)B, (uint64_t)K,
                                        (uint64_t)ldx * 2, BM, 64);
    const CUtensorMap tw = make_tmap_2d(data_ptr<void>(wt), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)N, (uint64_t)K,
                                        (uint64_t)K * 2, static_cast<uint32_t>(N), 64);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int key = static_cast<int>(K) * 1000 + static_cast<int>(N);
    switch (key) {
      case 128 * 1000 + 256: launch_dense<128, 256>(tx, tw, p, st); break;
      case 256 * 1000 + 128: launch_dense<256, 128>(tx, tw, p, st); break;
      case 128 * 1000 + 64: launch_dense<128, 64>(tx, tw, p, st); break;
      case 64 * 1000 + 128: launch_dense<64, 128>(tx, tw, p, st); break;
      case 256 * 1000 + 256: launch_dense<256, 256>(tx, tw, p, st); break;
      case 128 * 1000 + 128: launch_dense<128, 128>(tx, tw, p, st); break;
      case 256 * 1000 + 64: launch_dense<256, 64>(tx, tw, p, st); break;
      case 64 * 1000 + 64: launch_dense<64, 64>(tx, tw, p, st); break;
      default:
        RS_CHECK(false, "bf16 tensor-core Dense supports (in, out) in {64,128,256} x {64,128,256} with in*out <= 65536; got (" +
                            std::to_string(K) + ", " + std::to_string(N) + ")");
    }
  });
}

extern "C" int rs_gather_bf16(const DLTensor* table, const DLTensor* ids, DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(table, "table", 2, kDLFloat, 32);
    require(out, "out", 2, kDLBfloat, 16);
    RS_CHECK(ids != nullptr && ids->ndim == 1 && (is_dtype(ids, kDLInt, 32) || is_dtype(ids, kDLInt, 64)), "ids must be 1-D int32/int64");
    RS_CHECK(on_device(table) && on_device(out) && on_device(ids), "all tensors must be CUDA tensors");
    const int64_t B = out->shape[0], E = out->shape[1];
    RS_CHECK(ids->shape[0] == B && table->shape[1] == E && E % 8 == 0, "shape mismatch (E must be a multiple of 8)");
    RS_CHECK(B <= 1 || stride_of(ids, 0) == 1, "ids must be contiguous");
    const int64_t ld = row_stride_2d(table, "table"), ldo = row_stride_2d(out, "out");
    RS_CHECK(ld % 4 == 0 && ldo % 8 == 0, "row strides must keep 16-byte alignment");
    require_aligned16(data_ptr<float>(table), "table");
    require_aligned16(data_ptr<uint16_t>(out), "out");
    if (B == 0) return;
    const int chunks = static_cast<int>(E / 8);
    const int64_t want = (B * chunks + 255) / 256;
    const int blocks = static_cast<int>(std::min<int64_t>(want, (int64_t)sm_count() * 16));
    gather_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        data_ptr<float>(table), ld, table->shape[0], data_ptr<void>(ids), is_dtype(ids, kDLInt, 64) ? 1 : 0,
        data_ptr<uint16_t>(out), ldo, B, chunks);
    RS_LAUNCH_CHECK();
  });
}
