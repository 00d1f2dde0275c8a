// common.cu -- error state, device queries, TMA tensor-map construction.
#include "common.cuh"

#include <mutex>
#include <vector>

namespace rs {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launches{0};

void set_last_error(const std::string& m) { t_last_error = m; }

int sm_count() {
  static std::mutex mu;
  static int cached[64];
  static bool have[64] = {false};
  int dev = 0;
  RS_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 64 && have[dev]) return cached[dev];
  int n = 0;
  RS_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  int major = 0;
  RS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  RS_CHECK(major == 10, "this library is built for sm_100a only; device has compute capability major " +
                            std::to_string(major));
  // keep freed workspace in the stream-ordered pool instead of handing it back to the OS at every sync
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  if (dev < 64) {
    cached[dev] = n;
    have[dev] = true;
  }
  return n;
}

static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pairs[RS_PROF_SLOTS];

ProfScope::ProfScope(int slot_, cudaStream_t st_) : slot(slot_), st(st_) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    e0 = e1 = nullptr;
    return;
  }
  cudaEventRecord(e0, st);
}
ProfScope::~ProfScope() {
  if (!e0) return;
  cudaEventRecord(e1, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_pairs[slot].emplace_back(e0, e1);
}

encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  RS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
  RS_CHECK(qres == cudaDriverEntryPointSuccess && p != nullptr, "cuTensorMapEncodeTiled not available");
  fn = reinterpret_cast<encode_tiled_fn>(p);
  return fn;
}

static CUtensorMap make_tmap_2d_impl(const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols,
                                     uint64_t row_stride_bytes, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle sw,
                                     uint32_t inner_bytes) {
  RS_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base must be 16-byte aligned");
  RS_CHECK(row_stride_bytes % 16 == 0, "tensor map row stride must be a multiple of 16 bytes");
  RS_CHECK(box_cols * (uint64_t)elem_bytes == inner_bytes, "the swizzle mode needs an inner box of exactly its width");
  RS_CHECK(box_rows <= 256, "box rows must be <= 256");
  CUtensorMap m;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_tiled()(&m, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return m;
}

CUtensorMap make_tmap_2d(const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols,
                         uint64_t row_stride_bytes, uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d_impl(base, dt, elem_bytes, rows, cols, row_stride_bytes, box_rows, box_cols, CU_TENSOR_MAP_SWIZZLE_128B, 128);
}

CUtensorMap make_tmap_2d_sw64(const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols,
                              uint64_t row_stride_bytes, uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d_impl(base, dt, elem_bytes, rows, cols, row_stride_bytes, box_rows, box_cols, CU_TENSOR_MAP_SWIZZLE_64B, 64);
}

}  // namespace rs

extern "C" {

int rs_version(void) { return RS_ABI_VERSION; }

const char* rs_last_error(void) { return rs::t_last_error.c_str(); }

int rs_device_sm_count(int* out) {
  return rs::guard([&] {
    RS_CHECK(out != nullptr, "out is NULL");
    *out = rs::sm_count();
  });
}

int64_t rs_launch_count(void) { return rs::g_launches.load(std::memory_order_relaxed); }

int rs_prof_enable(int on) {
  return rs::guard([&] {
    std::lock_guard<std::mutex> lk(rs::g_prof_mu);
    for (auto& v : rs::g_prof_pairs) {
      for (auto& pr : v) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
      }
      v.clear();
    }
    rs::g_prof_on.store(on != 0);
  });
}

int rs_prof_read(int slot, double* total_ms, int64_t* launches) {
  return rs::guard([&] {
    RS_CHECK(slot >= 0 && slot < RS_PROF_SLOTS && total_ms && launches, "bad arguments");
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pairs;
    {
      std::lock_guard<std::mutex> lk(rs::g_prof_mu);
      pairs.swap(rs::g_prof_pairs[slot]);
    }
    double sum = 0.0;
    for (auto& pr : pairs) {
      RS_CUDA(cudaEventSynchronize(pr.second));
      float ms = 0.f;
      RS_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
      sum += ms;
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    *total_ms = sum;
    *launches = static_cast<int64_t>(pairs.size());
  });
}

}  // extern "C"
