// dense.cu -- K2 (fp32 mode): one Keras Dense layer, y = act(x @ kernel + bias).
// Replaces keras.layers.Dense in user_tower / item_tower / deep_nets (src/models.py:26-29,76-77).
//
// This is the fp32-exact arithmetic mode (every output is one fp32 FMA chain over k ascending), used
// for parity within 1e-5 of the reference's fp32 CPU path.  It is a shared-memory tiled FFMA kernel:
// CTA tile 128 x 64, k-step 16, 256 threads, 8 x 4 outputs per thread, register-staged double buffer.
// The tensor-core (tcgen05) ranking path lives in rank_fused.cu.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int LDA = BK + 4;  // floats; keeps float4 alignment and spreads rows over banks

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == RS_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == RS_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}

__global__ void __launch_bounds__(NT) dense_tiled_kernel(const float* __restrict__ x, int64_t ldx,
                                                         const float* __restrict__ w, int64_t ldw,
                                                         const float* __restrict__ bias, float* __restrict__ y,
                                                         int64_t ldy, int64_t M, int N, int K, int act) {
  __shared__ __align__(16) float As[2][BM * LDA];
  __shared__ __align__(16) float Bs[2][BK * BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // global -> register staging coordinates
  const int a_row0 = tid >> 2, a_kq = tid & 3;  // rows a_row0 and a_row0 + 64
  const int b_k = tid >> 4, b_nq = tid & 15;
  float4 ra[2], rb;

  auto load_tile = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = m0 + a_row0 + 64 * h;
      ra[h] = (r < M) ? __ldg(reinterpret_cast<const float4*>(x + r * ldx + k0 + a_kq * 4))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    rb = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(k0 + b_k) * ldw + n0 + b_nq * 4));
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      *reinterpret_cast<float4*>(&As[buf][(a_row0 + 64 * h) * LDA + a_kq * 4]) = ra[h];
    *reinterpret_cast<float4*>(&Bs[buf][b_k * BN + b_nq * 4]) = rb;
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  load_tile(0);
  store_tile(0);
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      float4 a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(&As[buf][(ty * 8 + i) * LDA + kk]);
#pragma unroll
      for (int u = 0; u < 4; ++u) b[u] = *reinterpret_cast<const float4*>(&Bs[buf][(kk + u) * BN + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // k ascending inside the 4-wide step
          acc[i][0] = fmaf(av[u], b[u].x, acc[i][0]);
          acc[i][1] = fmaf(av[u], b[u].y, acc[i][1]);
          acc[i][2] = fmaf(av[u], b[u].z, acc[i][2]);
          acc[i][3] = fmaf(av[u], b[u].w, acc[i][3]);
        }
      }
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }

  const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + n0 + tx * 4))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = m0 + ty * 8 + i;
    if (r < M) {
      float4 o;
      o.x = apply_act(acc[i][0] + bv.x, act);
      o.y = apply_act(acc[i][1] + bv.y, act);
      o.z = apply_act(acc[i][2] + bv.z, act);
      o.w = apply_act(acc[i][3] + bv.w, act);
      *reinterpret_cast<float4*>(y + r * ldy + n0 + tx * 4) = o;
    }
  }
}

// shape-generic fallback (one output per thread); same arithmetic, used for odd widths only
__global__ void __launch_bounds__(256) dense_generic_kernel(const float* __restrict__ x, int64_t ldx,
                                                            const float* __restrict__ w, int64_t ldw,
                                                            const float* __restrict__ bias, float* __restrict__ y,
                                                            int64_t ldy, int64_t M, int N, int K, int act) {
  const int64_t total = M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / N;
    const int n = static_cast<int>(i - r * N);
    float acc = 0.0f;
    for (int k = 0; k < K; ++k) acc = fmaf(x[r * ldx + k], w[(int64_t)k * ldw + n], acc);
    y[r * ldy + n] = apply_act(acc + (bias ? bias[n] : 0.0f), act);
  }
}

}  // namespace

extern "C" int rs_dense_forward(const DLTensor* x, const DLTensor* kernel, const DLTensor* bias, int act,
                                DLTensor* y, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(x, "x", 2, kDLFloat, 32);
    require(kernel, "kernel", 2, kDLFloat, 32);
    require(y, "y", 2, kDLFloat, 32);
    RS_CHECK(on_device(x) && on_device(kernel) && on_device(y), "x, kernel, y must be CUDA tensors");
    RS_CHECK(act == RS_ACT_NONE || act == RS_ACT_RELU || act == RS_ACT_SIGMOID, "unknown activation code");
    const int64_t M = x->shape[0];
    const int64_t K = x->shape[1];
    const int64_t N = kernel->shape[1];
    RS_CHECK(kernel->shape[0] == K, "kernel rows must equal x columns");
    RS_CHECK(y->shape[0] == M && y->shape[1] == N, "y must be [B, out]");
    const float* bp = nullptr;
    if (bias) {
      require(bias, "bias", 1, kDLFloat, 32);
      RS_CHECK(on_device(bias) && bias->shape[0] == N, "bias must be a CUDA [out] tensor");
      RS_CHECK(N <= 1 || stride_of(bias, 0) == 1, "bias must be contiguous");
      bp = data_ptr<float>(bias);
    }
    const int64_t ldx = row_stride_2d(x, "x"), ldw = row_stride_2d(kernel, "kernel"), ldy = row_stride_2d(y, "y");
    if (M == 0) return;
    const float* xp = data_ptr<float>(x);
    const float* wp = data_ptr<float>(kernel);
    float* yp = data_ptr<float>(y);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool tiled = (N % BN == 0) && (K % BK == 0) && ldx % 4 == 0 && ldw % 4 == 0 && ldy % 4 == 0 &&
                       al16(xp) && al16(wp) && al16(yp) && (bp == nullptr || al16(bp));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (tiled) {
      dim3 grid(static_cast<unsigned>((M + BM - 1) / BM), static_cast<unsigned>(N / BN));
      dense_tiled_kernel<<<grid, NT, 0, st>>>(xp, ldx, wp, ldw, bp, yp, ldy, M, (int)N, (int)K, act);
    } else {
      const int64_t want = (M * N + 255) / 256;
      const int64_t cap = (int64_t)sm_count() * 16;
      dense_generic_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, st>>>(xp, ldx, wp, ldw, bp, yp,
                                                                                      ldy, M, (int)N, (int)K, act);
    }
    RS_LAUNCH_CHECK();
  });
}
