// dcn.cu -- K3 (fp32 mode): DCN-v1 cross stack and the two ranking heads.
// Replaces DeepCrossNetwork.call's cross loop (src/models.py:38-44) and
// rating_head / ctr_head (src/models.py:119-120,131).
//
// The reference's cross layer is rank-1: s = x_l . w_i is one scalar per row, then
// x_{l+1} = x0 * s + b_i + x_l.  That is a row-wise dot product, not a GEMM, so the kernel is a
// warp-per-row streaming kernel (HBM-bound: D*4 bytes in, D*4 bytes out per row): each lane keeps
// D/32 features of x0 and x_l in registers, the dot is a shuffle tree, and all layers run without
// touching memory in between.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kMaxLayers = 8;

// D = 128 * C features per row; lane owns float4 chunks lane, lane+32, ...
template <int C>
__global__ void __launch_bounds__(256) dcn_cross_kernel(const float* __restrict__ x0p, int64_t ldx,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        int n_layers, float* __restrict__ out, int64_t ldo,
                                                        int64_t B) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int D = 128 * C;
  for (int64_t r = warp; r < B; r += nwarps) {
    float4 x0[C], xl[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      x0[c] = __ldg(reinterpret_cast<const float4*>(x0p + r * ldx) + lane + 32 * c);
      xl[c] = x0[c];
    }
    for (int l = 0; l < n_layers; ++l) {
      float s = 0.0f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (int64_t)l * D) + lane + 32 * c);
        s = fmaf(xl[c].x, wv.x, s);
        s = fmaf(xl[c].y, wv.y, s);
        s = fmaf(xl[c].z, wv.z, s);
        s = fmaf(xl[c].w, wv.w, s);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(b + (int64_t)l * D) + lane + 32 * c);
        xl[c].x = (x0[c].x * s + bv.x) + xl[c].x;  // bias sits outside the product (src/models.py:44)
        xl[c].y = (x0[c].y * s + bv.y) + xl[c].y;
        xl[c].z = (x0[c].z * s + bv.z) + xl[c].z;
        xl[c].w = (x0[c].w * s + bv.w) + xl[c].w;
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) *(reinterpret_cast<float4*>(out + r * ldo) + lane + 32 * c) = xl[c];
  }
}

// DCN-v2 combine (EXTENSION, not in the reference -- see rs_cross_v2_combine in the header): out = x0 * u + xl with
// u = x_l W_l + b_l from the Dense kernels; one fused multiply-add per element, 128-bit accesses
__global__ void __launch_bounds__(256) cross_v2_combine_kernel(const float* __restrict__ x0, int64_t ld0, const float* __restrict__ u,
                                                               int64_t ldu, const float* __restrict__ xl, int64_t ldl,
                                                               float* __restrict__ out, int64_t ldo, int64_t B, int chunks) {
  const int64_t total = B * chunks;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / chunks;
    const int c = static_cast<int>(i - r * chunks);
    const float4 a = __ldg(reinterpret_cast<const float4*>(x0 + r * ld0) + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(u + r * ldu) + c);
    const float4 d = __ldg(reinterpret_cast<const float4*>(xl + r * ldl) + c);
    *(reinterpret_cast<float4*>(out + r * ldo) + c) =
        make_float4(fmaf(a.x, b.x, d.x), fmaf(a.y, b.y, d.y), fmaf(a.z, b.z, d.z), fmaf(a.w, b.w, d.w));
  }
}

// z: [B, W] (W % 4 == 0); one warp per row computes both head dots
__global__ void __launch_bounds__(256) heads_kernel(const float* __restrict__ z, int64_t ldz, int W,
                                                    const float* __restrict__ hw, const float* __restrict__ hb,
                                                    float* __restrict__ rating, int64_t ldr,
                                                    float* __restrict__ ctr, int64_t ldc, int64_t B) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nchunk = W >> 2;
  for (int64_t r = warp; r < B; r += nwarps) {
    float sr = 0.0f, sc = 0.0f;
    for (int c = lane; c < nchunk; c += 32) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(z + r * ldz) + c);
      const float4 wr = __ldg(reinterpret_cast<const float4*>(hw) + c);
      const float4 wc = __ldg(reinterpret_cast<const float4*>(hw + W) + c);
      sr = fmaf(v.x, wr.x, sr); sr = fmaf(v.y, wr.y, sr); sr = fmaf(v.z, wr.z, sr); sr = fmaf(v.w, wr.w, sr);
      sc = fmaf(v.x, wc.x, sc); sc = fmaf(v.y, wc.y, sc); sc = fmaf(v.z, wc.z, sc); sc = fmaf(v.w, wc.w, sc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_xor_sync(0xFFFFFFFFu, sr, o);
      sc += __shfl_xor_sync(0xFFFFFFFFu, sc, o);
    }
    if (lane == 0) {
      rating[r * ldr] = sr + hb[0];
      ctr[r * ldc] = 1.0f / (1.0f + expf(-(sc + hb[1])));
    }
  }
}

}  // namespace

extern "C" int rs_dcn_cross_forward(const DLTensor* x0, const DLTensor* cross_w, const DLTensor* cross_b,
                                    DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(x0, "x0", 2, kDLFloat, 32);
    require(cross_w, "cross_w", 2, kDLFloat, 32);
    require(cross_b, "cross_b", 2, kDLFloat, 32);
    require(out, "out", 2, kDLFloat, 32);
    RS_CHECK(on_device(x0) && on_device(cross_w) && on_device(cross_b) && on_device(out),
             "all tensors must be CUDA tensors");
    const int64_t B = x0->shape[0], D = x0->shape[1];
    const int64_t L = cross_w->shape[0];
    RS_CHECK(L >= 0 && L <= kMaxLayers, "cross layer count must be in [0, 8]");
    RS_CHECK(cross_w->shape[1] == D && cross_b->shape[0] == L && cross_b->shape[1] == D,
             "cross_w / cross_b must be [n_layers, D]");
    RS_CHECK(is_compact(cross_w) && is_compact(cross_b), "cross_w / cross_b must be contiguous");
    RS_CHECK(out->shape[0] == B && out->shape[1] == D, "out must be [B, D]");
    RS_CHECK(D % 128 == 0 && D <= 512, "DCN input width must be a multiple of 128 and <= 512");
    const int64_t ldx = row_stride_2d(x0, "x0"), ldo = row_stride_2d(out, "out");
    RS_CHECK(ldx % 4 == 0 && ldo % 4 == 0, "row strides must be multiples of 4 floats");
    const float* xp = data_ptr<float>(x0);
    float* op = data_ptr<float>(out);
    const float* wp = data_ptr<float>(cross_w);
    const float* bp = data_ptr<float>(cross_b);
    require_aligned16(xp, "x0");
    require_aligned16(op, "out");
    require_aligned16(wp, "cross_w");
    require_aligned16(bp, "cross_b");
    if (B == 0) return;
    const int threads = 256;
    const int64_t want = (B * 32 + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 8;
    const int blocks = static_cast<int>(want < cap ? want : cap);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (D / 128) {
      case 1: dcn_cross_kernel<1><<<blocks, threads, 0, st>>>(xp, ldx, wp, bp, (int)L, op, ldo, B); break;
      case 2: dcn_cross_kernel<2><<<blocks, threads, 0, st>>>(xp, ldx, wp, bp, (int)L, op, ldo, B); break;
      case 3: dcn_cross_kernel<3><<<blocks, threads, 0, st>>>(xp, ldx, wp, bp, (int)L, op, ldo, B); break;
      default: dcn_cross_kernel<4><<<blocks, threads, 0, st>>>(xp, ldx, wp, bp, (int)L, op, ldo, B); break;
    }
    RS_LAUNCH_CHECK();
  });
}

extern "C" int rs_heads_forward(const DLTensor* z, const DLTensor* head_w, const DLTensor* head_b,
                                DLTensor* rating, DLTensor* ctr, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(z, "z", 2, kDLFloat, 32);
    require(head_w, "head_w", 2, kDLFloat, 32);
    require(head_b, "head_b", 1, kDLFloat, 32);
    require(rating, "rating", 2, kDLFloat, 32);
    require(ctr, "ctr", 2, kDLFloat, 32);
    RS_CHECK(on_device(z) && on_device(head_w) && on_device(head_b) && on_device(rating) && on_device(ctr),
             "all tensors must be CUDA tensors");
    const int64_t B = z->shape[0], W = z->shape[1];
    RS_CHECK(head_w->shape[0] == 2 && head_w->shape[1] == W && is_compact(head_w), "head_w must be contiguous [2, W]");
    RS_CHECK(head_b->shape[0] == 2, "head_b must be [2]");
    RS_CHECK(rating->shape[0] == B && rating->shape[1] == 1 && ctr->shape[0] == B && ctr->shape[1] == 1,
             "rating / ctr must be [B, 1]");
    RS_CHECK(W % 4 == 0, "head input width must be a multiple of 4");
    const int64_t ldz = row_stride_2d(z, "z");
    RS_CHECK(ldz % 4 == 0, "z row stride must be a multiple of 4 floats");
    require_aligned16(data_ptr<float>(z), "z");
    require_aligned16(data_ptr<float>(head_w), "head_w");
    if (B == 0) return;
    const int64_t ldr = B <= 1 ? 1 : stride_of(rating, 0), ldc = B <= 1 ? 1 : stride_of(ctr, 0);
    const int threads = 256;
    const int64_t want = (B * 32 + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 8;
    heads_kernel<<<static_cast<int>(want < cap ? want : cap), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        data_ptr<float>(z), ldz, (int)W, data_ptr<float>(head_w), data_ptr<float>(head_b), data_ptr<float>(rating),
        ldr, data_ptr<float>(ctr), ldc, B);
    RS_LAUNCH_CHECK();
  });
}

extern "C" int rs_cross_v2_combine(const DLTensor* x0, const DLTensor* u, const DLTensor* xl, DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    const DLTensor* ts[4] = {x0, u, xl, out};
    const char* names[4] = {"x0", "u", "xl", "out"};
    int64_t ld[4];
    for (int i = 0; i < 4; ++i) {
      require(ts[i], names[i], 2, kDLFloat, 32);
      RS_CHECK(on_device(ts[i]) && ts[i]->shape[0] == x0->shape[0] && ts[i]->shape[1] == x0->shape[1], "x0, u, xl, out must be CUDA f32 [B, D] tensors");
      ld[i] = row_stride_2d(ts[i], names[i]);
      RS_CHECK(ld[i] % 4 == 0, "row strides must be multiples of 4 floats");
      require_aligned16(data_ptr<float>(ts[i]), names[i]);
    }
    const int64_t B = x0->shape[0], D = x0->shape[1];
    RS_CHECK(D % 4 == 0, "D must be a multiple of 4");
    if (B == 0 || D == 0) return;
    const int chunks = static_cast<int>(D / 4);
    const int blocks = static_cast<int>(std::min<int64_t>((B * chunks + 255) / 256, (int64_t)sm_count() * 16));
    cross_v2_combine_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(data_ptr<float>(x0), ld[0], data_ptr<float>(u), ld[1],
                                                                                  data_ptr<float>(xl), ld[2], data_ptr<float>(out), ld[3], B, chunks);
    RS_LAUNCH_CHECK();
  });
}
