// common.cuh -- shared host/device helpers for librecsys_b200 (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/recsys_b200.h"

// --------------------------------------------------------------------------------------
// host: errors, DLTensor validation
// --------------------------------------------------------------------------------------
namespace rs {

struct Error : public std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

void set_last_error(const std::string& m);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define RS_CHECK(cond, msg)                                                          \
  do {                                                                               \
    if (!(cond)) throw ::rs::Error(std::string(__func__) + ": " + (msg));            \
  } while (0)

#define RS_CUDA(expr)                                                                \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess)                                                           \
      throw ::rs::Error(std::string(__func__) + ": " #expr " -> " + cudaGetErrorString(_e)); \
  } while (0)

#define RS_LAUNCH_CHECK()                                                            \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess)                                                           \
      throw ::rs::Error(std::string(__func__) + ": kernel launch -> " + cudaGetErrorString(_e)); \
    ::rs::count_launch();                                                            \
  } while (0)

// Every extern "C" entry is `return guard([&]{ ... });`
template <class F>
inline int guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return 1;
  } catch (...) {
    set_last_error("unknown error");
    return 2;
  }
}

inline bool is_dtype(const DLTensor* t, uint8_t code, uint8_t bits) {
  return t->dtype.code == code && t->dtype.bits == bits && t->dtype.lanes == 1;
}
inline bool on_device(const DLTensor* t) { return t->device.device_type == kDLCUDA; }
inline bool on_host(const DLTensor* t) {
  return t->device.device_type == kDLCPU || t->device.device_type == kDLCUDAHost;
}
inline int64_t stride_of(const DLTensor* t, int dim) {
  if (t->strides) return t->strides[dim];
  int64_t s = 1;
  for (int i = t->ndim - 1; i > dim; --i) s *= t->shape[i];
  return s;
}
inline bool is_compact(const DLTensor* t) {
  int64_t s = 1;
  for (int i = t->ndim - 1; i >= 0; --i) {
    if (t->shape[i] != 1 && stride_of(t, i) != s) return false;
    s *= t->shape[i];
  }
  return true;
}
template <class T>
inline T* data_ptr(const DLTensor* t) {
  return reinterpret_cast<T*>(static_cast<char*>(t->data) + t->byte_offset);
}
inline void require(const DLTensor* t, const char* name, int ndim, uint8_t code, uint8_t bits) {
  RS_CHECK(t != nullptr, std::string(name) + " is NULL");
  RS_CHECK(t->ndim == ndim, std::string(name) + ": expected ndim " + std::to_string(ndim) + ", got " +
                                std::to_string(t->ndim));
  RS_CHECK(is_dtype(t, code, bits), std::string(name) + ": wrong dtype (code " +
                                        std::to_string(t->dtype.code) + ", bits " +
                                        std::to_string(t->dtype.bits) + ")");
}
// 2-D tensor whose last dim is contiguous; returns the row stride in elements
inline int64_t row_stride_2d(const DLTensor* t, const char* name) {
  RS_CHECK(t->shape[1] == 1 || stride_of(t, 1) == 1, std::string(name) + ": last dim must be contiguous");
  return t->shape[0] <= 1 ? t->shape[1] : stride_of(t, 0);
}
inline void require_aligned16(const void* p, const char* name) {
  RS_CHECK((reinterpret_cast<uintptr_t>(p) & 15) == 0, std::string(name) + ": data must be 16-byte aligned");
}

int sm_count();  // SMs of the current device (cached per device)

// Brackets a kernel family's launches with CUDA events when rs_prof_enable(1) is active.
struct ProfScope {
  int slot;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfScope(int slot, cudaStream_t st);
  ~ProfScope();
};

// cuTensorMapEncodeTiled fetched through the runtime (no -lcuda link)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode_tiled();

// 2-D row-major [rows, cols] tensor map with a [box_rows, box_cols] box and 128-byte swizzle
CUtensorMap make_tmap_2d(const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols,
                         uint64_t row_stride_bytes, uint32_t box_rows, uint32_t box_cols);
// the same with a 64-byte inner box and the 64-byte swizzle (16-byte chunk index ^= (row >> 1) & 3)
CUtensorMap make_tmap_2d_sw64(const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols,
                              uint64_t row_stride_bytes, uint32_t box_rows, uint32_t box_cols);

}  // namespace rs

// --------------------------------------------------------------------------------------
// device: PTX wrappers (mbarrier, TMA), packed top-k keys
// --------------------------------------------------------------------------------------
namespace rsd {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of burning issue slots that the MMA issuer and the working epilogue warps need
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails with "unspecified launch failure") instead of hanging the GPU.
// No printf and no out-of-line helper here on purpose: either is an ABI call, and a call inside a warp-specialised
// kernel makes the compiler spill every register that is live across the wait site (measured: the gather warps of
// tower_sp_kernel spilled 16 of their 64 row registers around each wait) -- and a spill is an L2 round trip in kernels
// that leave the SM no L1 (227 KB of shared memory).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();   // each failed try sleeps up to the hint, so this is seconds, not microseconds
  }
}

// Spinning wait (mbarrier.test_wait, no suspend): for strictly serial hand-offs where the waiter has nothing else to
// do and the wake-up latency of a suspended try_wait would sit on the critical path every time.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_test_wait(bar, parity)) {
    if (++spins > (1u << 28)) __trap();
  }
}

// L2 cache policies (the fixed encodings CUTLASS uses for createpolicy results)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0,
                                            int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// shared -> global bulk stores (async proxy).  The issuing thread makes its own and -- after a __syncwarp -- its warp's
// shared-memory writes visible with fence_proxy_async_smem() first; completion is tracked per thread in bulk groups.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups may still be READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// all of this thread's bulk groups have completed (writes performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// 256-bit global stores (STG.256, new on sm_100): one full 32-byte sector per thread and instruction
__device__ __forceinline__ void st_global_256(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f),
               "f"(g), "f"(h)
               : "memory");
}
__device__ __forceinline__ void st_global_256(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                              uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f),
               "r"(g), "r"(h)
               : "memory");
}

// ---- packed (score, row) keys: larger key == better hit under (score desc, row asc) ----
__device__ __forceinline__ uint32_t f32_to_ordered(float s) {
  s += 0.0f;  // -0 -> +0 so that equal floats give equal keys
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(uint32_t o) {
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t pack_key(float s, uint32_t row) {
  return (static_cast<uint64_t>(f32_to_ordered(s)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - row);
}
__device__ __forceinline__ float key_score(uint64_t k) { return ordered_to_f32(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ uint32_t key_row(uint64_t k) { return 0xFFFFFFFFu - static_cast<uint32_t>(k); }
constexpr uint64_t kEmptyKey = 0ull;  // below every real key (ordered(-inf) = 0x007FFFFF)

// CTA-wide bitonic sort, descending, of a[0..CAP) in shared memory; ends with __syncthreads().
template <int CAP, int NT>
__device__ __forceinline__ void block_sort_desc(uint64_t* a, int tid) {
#pragma unroll 1
  for (int size = 2; size <= CAP; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < CAP / 2; i += NT) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t x = a[lo], y = a[hi];
        if ((x < y) == desc) {
          a[lo] = y;
          a[hi] = x;
        }
      }
      __syncthreads();
    }
  }
}

// Block-wide radix select: the k-th largest (1-based, k <= n) of the n 32-bit values v(0..n-1).  Four 8-bit passes
// from the top byte; per pass a shared-memory histogram and a one-warp scan.  12 barriers in total (a bitonic
// sort of 1024 keys needs 55), which is what matters for these latency-bound per-query kernels.
template <int NT, class F>
__device__ __forceinline__ uint32_t block_kth_largest(F v, int n, int k, unsigned* hist /*[256]*/,
                                                      unsigned* state /*[2]*/, int tid) {
  uint32_t prefix = 0, mask = 0;
  unsigned kk = static_cast<unsigned>(k);
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += NT) hist[i] = 0;
    __syncthreads();
    // Warp-aggregated histogram: the values are scores of one query, so most of them share the top byte (and often the
    // second) -- plain atomics then all hit one or two shared-memory words and serialise 32 ways inside every warp and
    // again across the eight CTAs that share an SM.  One atomic per distinct bin and warp instead.
    // (The two low bytes are spread over many bins: there a match costs more than the atomics it saves.)
    for (int i0 = 0; i0 < n; i0 += NT) {
      const int i = i0 + tid;
      uint32_t bin = 0xFFFFFFFFu;                 // "no value": inactive lanes group among themselves
      if (i < n) {
        const uint32_t x = v(i);
        if ((x & mask) == prefix) bin = (x >> shift) & 255u;
      }
      if (shift >= 16) {
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, bin);
        if (bin != 0xFFFFFFFFu && (tid & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], static_cast<unsigned>(__popc(peers)));
      } else if (bin != 0xFFFFFFFFu) {
        atomicAdd(&hist[bin], 1u);
      }
    }
    __syncthreads();
    if (tid < 32) {   // lane L owns bins 255-8L .. 248-8L (descending), prefix-summed over lanes
      unsigned loc[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        loc[j] = hist[255 - (tid * 8 + j)];
        sum += loc[j];
      }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (tid >= o) incl += t;
      }
      const unsigned excl = incl - sum;
      if (excl < kk && kk <= incl) {
        unsigned run = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (run < kk && kk <= run + loc[j]) {
            state[0] = 255u - (tid * 8 + j);
            state[1] = kk - run;
          }
          run += loc[j];
        }
      }
    }
    __syncthreads();
    prefix |= state[0] << shift;
    mask |= 0xFFu << shift;
    kk = state[1];
  }
  return prefix;
}


// CTA-wide descending sort of a[0..n_pow2) for the smallest power of two >= n (entries past n must hold kEmptyKey up to it):
// the latency of these sorts is barriers, so short lists get short sorts
template <int CAP, int NT>
__device__ __forceinline__ void block_sort_desc_upto(uint64_t* a, int n, int tid) {
  if (CAP >= 4096 && n > 2048) block_sort_desc<(CAP >= 4096 ? 4096 : CAP), NT>(a, tid);
  else if (CAP >= 2048 && n > 1024) block_sort_desc<(CAP >= 2048 ? 2048 : CAP), NT>(a, tid);
  else if (CAP >= 1024 && n > 512) block_sort_desc<(CAP >= 1024 ? 1024 : CAP), NT>(a, tid);
  else if (CAP >= 512 && n > 256) block_sort_desc<(CAP >= 512 ? 512 : CAP), NT>(a, tid);
  else if (CAP >= 256 && n > 128) block_sort_desc<(CAP >= 256 ? 256 : CAP), NT>(a, tid);
  else block_sort_desc<(CAP >= 128 ? 128 : CAP), NT>(a, tid);
}

}  // namespace rsd
