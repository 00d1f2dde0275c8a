// dense_x3.cu -- Keras Dense, y = act(x @ kernel + bias), at fp32 accuracy ON THE TENSOR CORES ("fp32 tensor mode").
// Replaces keras.layers.Dense in user_tower / item_tower / deep_nets (src/models.py:26-29,76-77) for the precision the
// reference actually computes in (fp32), at ranking batch sizes.  The FFMA kernel (dense.cu) stays the small-batch path.
//
// Arithmetic: every fp32 operand is split into two fp16 numbers, v = hi + lo * 2^-11 with hi = fp16(v) and
// lo = fp16((v - hi) * 2^11) (the scaling keeps lo out of fp16's subnormal range); hi + lo * 2^-11 carries 22 significant
// bits.  x.w = hi_x hi_w + 2^-11 (hi_x lo_w + lo_x hi_w) + O(2^-22): three tcgen05.mma (kind::f16, fp32 accumulate)
// per K step, the main term and the two correction terms into SEPARATE tensor-memory accumulators, combined in the
// epilogue (main + 2^-11 corr).  Dropped: lo_x lo_w, 2^-22 relative -- below fp32's own rounding of a K-long sum.
// Measured against fp64 in tests/test_gpu_dense_x3.py; the stated tolerance of the fp32 mode is 1e-5 of tensor scale.
// Range: |v| < 65504 (fp16); activations and kernels of a recommender are O(1).
//
// One persistent CTA per SM, weight-stationary:
//   * the layer's kernel, pre-split and transposed to [out][in] fp16 hi / lo images (rs_dense_x3_prepare, once per weight
//     change), is TMA-loaded once per CTA: <= 128 KB (N tile 128 x K 256 x 2 images, or both N halves of a K = 128 layer);
//   * 8 producer warps read fp32 activation rows (coalesced 128-bit loads, next chunk in flight while the current one is
//     split), split them and store the hi / lo images of a [128 rows x 64] K chunk into a 3-stage ring in the 128-byte
//     swizzled K-major layout UMMA reads (what TMA would have written had the data been fp16 already);
//   * one thread issues 3 x 4 MMAs per chunk (M128 x N<=128 x K16) into one of two TMEM slots (main | corr, 256 columns);
//   * 4 epilogue warps combine, add the bias, apply the activation and store fp32 rows (256-bit stores).
// Per sample and layer: 3 x 2 x in x out tensor flop; bytes: 4 in read + 4 out written -- HBM-bound at these widths.
#include <algorithm>
#include <mutex>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace {

using namespace rsd;

constexpr int BM = 128;
constexpr int X3_PROD_WARPS = 8;
constexpr int X3_EPI0 = 8;                       // warps 8-11: epilogue (TMEM lane quarter = warp & 3)
constexpr int X3_ISSUER = 12;
constexpr int X3_THREADS = 13 * 32;
constexpr int X3_STAGES = 3;
constexpr int X3_STAGE_BYTES = 2 * BM * 128;     // hi image [128 x 64 fp16] | lo image
constexpr float kLoScale = 2048.0f;              // 2^11
constexpr float kLoInv = 1.0f / 2048.0f;

struct X3Params {
  const float* x;
  int64_t ldx;
  int64_t rows;
  float* y;
  int64_t ldy;
  const float* bias;   // [N] or nullptr
  int act;
  int nh;              // N tiles per CTA (1 or 2); grid.y covers the rest
};

// instruction descriptor, kind::f16: D = f32, A = B = fp16 (format 0), both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16_f32(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

template <int K, int NT>
constexpr int x3_smem(int nh) {
  return nh * (K / 64) * NT * 128 * 2 + X3_STAGES * X3_STAGE_BYTES + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*align*/;
}

template <int K, int NT>
__global__ void __launch_bounds__(X3_THREADS, 1)
dense_x3_kernel(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo, const X3Params p) {
  constexpr int KC = K / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nh = p.nh;
  const int w_img = nh * KC * NT * 128;          // bytes of one weight image (hi or lo)
  uint8_t* w_hi = base;
  uint8_t* w_lo = base + w_img;
  uint8_t* ring = base + 2 * w_img;
  float* bias_s = reinterpret_cast<float*>(ring + X3_STAGES * X3_STAGE_BYTES);   // [nh * NT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 256);
  uint64_t* wfull = bars;
  uint64_t* full = bars + 1;                     // [STAGES] producers -> issuer
  uint64_t* empty = full + X3_STAGES;            // [STAGES] commit -> producers
  uint64_t* tfull = empty + X3_STAGES;           // [2] commit -> epilogue
  uint64_t* tempty = tfull + 2;                  // [2] epilogue -> issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int64_t ntiles = (p.rows + BM - 1) / BM;
  const int64_t n_my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t n_items = n_my * nh;             // work items: (row tile, N tile), the N tiles of a row tile back to back
  const int n0 = blockIdx.y * nh * NT;           // first output column of this CTA

  if (warp == X3_ISSUER && lane == 0) {
    tma_prefetch_desc(&tmap_hi);
    tma_prefetch_desc(&tmap_lo);
    mbar_init(wfull, 1);
    for (int s = 0; s < X3_STAGES; ++s) {
      mbar_init(&full[s], X3_PROD_WARPS);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < nh * NT; i += X3_THREADS) bias_s[i] = p.bias ? p.bias[n0 + i] : 0.0f;
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
          tma_load_2d(w_hi + off, &tmap_hi, wfull, kc * 64, n0 + hh * NT, kEvictLast);
          tma_load_2d(w_lo + off, &tmap_lo, wfull, kc * 64, n0 + hh * NT, kEvictLast);
        }
      constexpr uint32_t idesc = umma_idesc_f16_f32(BM, NT);
      mbar_wait(wfull, 0);
      const uint64_t whi0 = umma_desc_k_sw128(smem_u32(w_hi)), wlo0 = umma_desc_k_sw128(smem_u32(w_lo));
      const uint64_t ring0 = umma_desc_k_sw128(smem_u32(ring));
      uint32_t it = 0;
      for (int64_t i = 0; i < n_items; ++i) {
        const int hh = static_cast<int>(i % nh);
        const uint32_t a = static_cast<uint32_t>(i) & 1u, aph = (static_cast<uint32_t>(i) >> 1) & 1u;
        mbar_wait(&tempty[a], aph ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_main = tmem_base + a * 256, d_corr = d_main + 128;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const uint32_t s = it % X3_STAGES, ph = (it / X3_STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
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
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[a]);
      }
    }
  } else if (warp < X3_PROD_WARPS) {
    // ---- producers: fp32 rows -> (hi, lo) fp16 images of a [128 x 64] K chunk, 128-byte swizzled ----
    // a half-warp covers one row's 64 floats (16 x float4); a warp instruction covers 2 rows, a warp 16 rows of the chunk
    const int rsub = lane >> 4;        // which of the 2 rows
    const int j = lane & 15;           // float4 index within the row's chunk
    float4 buf[8];
    auto load_chunk = [&](int64_t item, int kc) {
      const int64_t tile = blockIdx.x + (item / nh) * gridDim.x;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t row = tile * BM + warp * 16 + e * 2 + rsub;
        buf[e] = row < p.rows ? __ldg(reinterpret_cast<const float4*>(p.x + row * p.ldx + kc * 64) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    const int64_t total = n_items * KC;
    if (total > 0) load_chunk(0, 0);
    for (int64_t g = 0; g < total; ++g) {
      float4 cur[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) cur[e] = buf[e];
      if (g + 1 < total) load_chunk((g + 1) / KC, static_cast<int>((g + 1) % KC));
      const uint32_t s = static_cast<uint32_t>(g % X3_STAGES), ph = static_cast<uint32_t>((g / X3_STAGES) & 1);
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* st_hi = ring + s * X3_STAGE_BYTES;
      uint8_t* st_lo = st_hi + BM * 128;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int r = warp * 16 + e * 2 + rsub;                       // row within the tile
        const float v[4] = {cur[e].x, cur[e].y, cur[e].z, cur[e].w};
        __half h[4], l[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          h[t] = __float2half_rn(v[t]);
          l[t] = __float2half_rn((v[t] - __half2float(h[t])) * kLoScale);
        }
        // 16-byte chunk c = j / 2 of the 128-byte row, XOR-swizzled with the row's position in its 8-row group
        const uint32_t off = static_cast<uint32_t>(r * 128 + (((j >> 1) ^ (r & 7)) << 4) + (j & 1) * 8);
        uint2 ph2, pl2;
        ph2.x = static_cast<uint32_t>(__half_as_ushort(h[0])) | (static_cast<uint32_t>(__half_as_ushort(h[1])) << 16);
        ph2.y = static_cast<uint32_t>(__half_as_ushort(h[2])) | (static_cast<uint32_t>(__half_as_ushort(h[3])) << 16);
        pl2.x = static_cast<uint32_t>(__half_as_ushort(l[0])) | (static_cast<uint32_t>(__half_as_ushort(l[1])) << 16);
        pl2.y = static_cast<uint32_t>(__half_as_ushort(l[2])) | (static_cast<uint32_t>(__half_as_ushort(l[3])) << 16);
        *reinterpret_cast<uint2*>(st_hi + off) = ph2;
        *reinterpret_cast<uint2*>(st_lo + off) = pl2;
      }
      fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
  } else if (warp >= X3_EPI0 && warp < X3_EPI0 + 4) {
    // ---- epilogue: main + 2^-11 corr + bias -> activation -> fp32 rows ----
    const int quarter = warp & 3;
    for (int64_t i = 0; i < n_items; ++i) {
      const int hh = static_cast<int>(i % nh);
      const int64_t tile = blockIdx.x + (i / nh) * gridDim.x;
      const int64_t row = tile * BM + quarter * 32 + lane;
      const uint32_t a = static_cast<uint32_t>(i) & 1u, aph = (static_cast<uint32_t>(i) >> 1) & 1u;
      mbar_wait(&tfull[a], aph);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + a * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
      float* yrow = p.y + row * p.ldy + n0 + hh * NT;
#pragma unroll 1
      for (int c = 0; c < NT / 32; ++c) {
        float m[32], cr[32];
        __syncwarp();
        tmem_ld_32x32(taddr + c * 32, m);
        tmem_ld_32x32(taddr + 128 + c * 32, cr);
        const float* bs = bias_s + hh * NT + c * 32;
#pragma unroll
        for (int e = 0; e < 32; ++e) m[e] = fmaf(cr[e], kLoInv, m[e]) + bs[e];
        if (p.act == RS_ACT_RELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) m[e] = fmaxf(m[e], 0.0f);
        }
        if (row < p.rows) {
#pragma unroll
          for (int e = 0; e < 32; e += 8)
            st_global_256(yrow + c * 32 + e, m[e], m[e + 1], m[e + 2], m[e + 3], m[e + 4], m[e + 5], m[e + 6], m[e + 7]);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
    }
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
    const float v = w[k * ld + n];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn((v - __half2float(h)) * kLoScale);
  }
}

template <int K, int NT>
void launch_x3(const CUtensorMap& thi, const CUtensorMap& tlo, X3Params p, int N, cudaStream_t st) {
  // N tiles per CTA: as many as keep both weight images within 128 KB
  const int n_tiles = N / NT;
  const int fit = std::max(1, (128 * 1024) / ((K / 64) * NT * 128 * 2));
  p.nh = std::min(n_tiles, std::min(fit, 2));
  while (n_tiles % p.nh) --p.nh;
  const int smem = x3_smem<K, NT>(p.nh);
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(dense_x3_kernel<K, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, x3_smem<K, NT>(2) <= 232448 ? x3_smem<K, NT>(2) : x3_smem<K, NT>(1));
  });
  const int64_t ntiles = (p.rows + BM - 1) / BM;
  const int gy = n_tiles / p.nh;
  const int gx = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ntiles, rs::sm_count() / gy)));
  rs::ProfScope prof(RS_PROF_RANK, st);
  dense_x3_kernel<K, NT><<<dim3(gx, gy), X3_THREADS, smem, st>>>(thi, tlo, p);
  RS_LAUNCH_CHECK();
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

extern "C" int rs_dense_forward_x3(const DLTensor* x, const DLTensor* wt_hi, const DLTensor* wt_lo, const DLTensor* bias, int act,
                                   DLTensor* y, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(x, "x", 2, kDLFloat, 32);
    require(wt_hi, "wt_hi", 2, kDLFloat, 16);
    require(wt_lo, "wt_lo", 2, kDLFloat, 16);
    require(y, "y", 2, kDLFloat, 32);
    RS_CHECK(on_device(x) && on_device(wt_hi) && on_device(wt_lo) && on_device(y), "all tensors must be CUDA tensors");
    const int64_t B = x->shape[0], K = x->shape[1], N = wt_hi->shape[0];
    RS_CHECK(wt_hi->shape[1] == K && wt_lo->shape[0] == N && wt_lo->shape[1] == K && is_compact(wt_hi) && is_compact(wt_lo),
             "wt_hi / wt_lo must be contiguous [out, in] fp16 tensors (rs_dense_x3_prepare)");
    RS_CHECK(rs_dense_x3_supported(K, N, act), "fp32 tensor-mode Dense supports (in, out) in {64,128,256}^2 with activation none / relu");
    RS_CHECK(y->shape[0] == B && y->shape[1] == N, "y must be [B, out]");
    X3Params p{};
    p.x = data_ptr<float>(x);
    p.ldx = row_stride_2d(x, "x");
    p.rows = B;
    p.y = data_ptr<float>(y);
    p.ldy = row_stride_2d(y, "y");
    p.act = act;
    RS_CHECK(p.ldx % 4 == 0 && p.ldy % 8 == 0, "x row stride must be a multiple of 4 floats, y row stride a multiple of 8");
    require_aligned16(p.x, "x");
    RS_CHECK((reinterpret_cast<uintptr_t>(p.y) & 31) == 0, "y must be 32-byte aligned");
    if (bias) {
      require(bias, "bias", 1, kDLFloat, 32);
      RS_CHECK(on_device(bias) && bias->shape[0] == N && (N <= 1 || stride_of(bias, 0) == 1), "bias must be a contiguous CUDA [out] tensor");
      p.bias = data_ptr<float>(bias);
    }
    if (B == 0) return;
    sm_count();
    const int NT = N >= 128 ? 128 : 64;
    const CUtensorMap thi = make_tmap_2d(data_ptr<void>(wt_hi), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)N, (uint64_t)K,
                                         (uint64_t)K * 2, static_cast<uint32_t>(NT), 64);
    const CUtensorMap tlo = make_tmap_2d(data_ptr<void>(wt_lo), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)N, (uint64_t)K,
                                         (uint64_t)K * 2, static_cast<uint32_t>(NT), 64);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Ni = static_cast<int>(N);
    if (NT == 128) {
      switch (K) {
        case 64: launch_x3<64, 128>(thi, tlo, p, Ni, st); break;
        case 128: launch_x3<128, 128>(thi, tlo, p, Ni, st); break;
        default: launch_x3<256, 128>(thi, tlo, p, Ni, st); break;
      }
    } else {
      switch (K) {
        case 64: launch_x3<64, 64>(thi, tlo, p, Ni, st); break;
        case 128: launch_x3<128, 64>(thi, tlo, p, Ni, st); break;
        default: launch_x3<256, 64>(thi, tlo, p, Ni, st); break;
      }
    }
  });
}
