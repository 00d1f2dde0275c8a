// flat_index.cuh -- the catalog handle shared by the exact scan (flat_index.cu) and the
// tensor-core pre-filter (tc_filter.cu).
#pragma once

#include <mutex>

#include "common.cuh"

struct rs_peer;

struct rs_index {
  int d = 0;
  int device = 0;
  int64_t ntotal = 0;
  int64_t capacity = 0;
  float* xb = nullptr;           // fp32 catalog [capacity, d], row-major; rows >= ntotal unused
  CUtensorMap tmap_f32{};        // over [ntotal, d] fp32, box [128 rows, 32 floats], 128B swizzle
  // bf16 shadow catalog for the tcgen05 pre-filter (built lazily on the first tensor-mode search)
  uint16_t* xb_bf16 = nullptr;   // [capacity, d] bf16 bits
  float* row_norm = nullptr;     // [capacity] ||x_i||_2 (fp32), for the filter's error bound
  int64_t shadow_rows = 0;       // rows of the shadow that are current
  float max_norm = 0.0f;         // max_i ||x_i||_2 over the shadow rows (host copy)
  CUtensorMap tmap_bf16{};       // over [ntotal, d] bf16, box [256 rows, 64 bf16], 128B swizzle
  // queries of the last / all tensor-mode searches that were re-run on the exact scan; atomics: concurrent searches on one
  // handle from several host threads are allowed (include/recsys_b200.h)
  std::atomic<int64_t> last_fallbacks{0};
  std::atomic<int64_t> total_fallbacks{0};
  // AUTO mode on a small catalog: set when a tensor-path search had to repeat most of its queries on the exact scan (scores
  // packed inside the filter margin, e.g. item embeddings that all point the same way) -- AUTO then stays on the scan until
  // rows are added; an explicit RS_TENSOR_BF16 request still tries
  std::atomic<bool> small_tensor_off{false};
  std::mutex mu;                 // serialises add / shadow build
};

namespace rs {

// launches of the exact fp32 scan (flat_index.cu); q/D/I are device pointers
void search_exact_device(rs_index* h, const float* q, int64_t nq, int k, int64_t id_base, float* D, int64_t ldD,
                         int64_t* I, int64_t ldI, cudaStream_t st);
// tensor-core pre-filter + exact re-score (tc_filter.cu); same results as search_exact_device
// makes `dev` current for the lifetime of the object
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// shared-threshold search of one shard of a partitioned catalog (tc_filter.cu)
int64_t tensor_sample_rows(int64_t ntotal);
int tensor_threshold_rank(int k, double rho);
int64_t tensor_max_queries();
// publish_tops != nullptr: the list goes into slot `rank` of every rank's tops region instead of `top` (the kernel that
// builds it publishes it: no separate publish launch), and the call returns with every rank's list present in the local region
rs_tc_session* shared_begin(rs_index* h, const float* q, int64_t nq, int k, int R, float* top, int64_t ld_top, cudaStream_t st,
                            rs_peer* publish_tops = nullptr);
// tops_region != nullptr: `tops` is that peer-memory region's local data; the threshold kernel waits for every rank's
// message and acknowledges the step itself (no rs_peer_wait / rs_peer_ack around the call)
void shared_finish(rs_tc_session* s, const float* tops, int n_parts, int64_t part_ld, int R, int rank, int64_t id_base,
                   float* D, int64_t ldD, int64_t* I, int64_t ldI, int* need_exact, float* bound,
                   rs_peer* tops_region = nullptr, rs_peer* publish_lists = nullptr, int64_t off_I = 0, int64_t off_F = 0,
                   int64_t off_B = 0);
// publish_lists != nullptr: D / I / need_exact / bound are ignored; the re-scoring kernel writes this shard's lists message
// (scores f32 [nq, k] at 0 | rows int64 [nq, k] at off_I | need_exact int32 [nq] at off_F | bounds f32 [nq] at off_B) into
// slot `rank` of every rank's lists region and waits, in its last CTA, for every rank's message

// G-way merge of per-shard (score, global row) lists -- parts strided by part_ld_* elements, each a packed [nq, k] block --
// optionally fused with the steps around it in the sharded search (sharded.cu):
//   lists_region  the parts live in that peer-memory region: wait for every rank's message at entry, acknowledge at the end
//   need_exact    per query: OR of the shards' flags [G rows, stride ld_f]; with bounds [G rows, stride ld_b] also the
//                 shared-threshold verification (rs_topk_verify): merged list full and its k-th score >= every bound
void merge_lists_device(const float* Dparts, int64_t part_ld_D, const int64_t* Iparts, int64_t part_ld_I, int P, int64_t nq,
                        int k, float* D, int64_t ldD, int64_t* I, int64_t ldI, rs_peer* lists_region, const float* bounds,
                        int64_t ld_b, const int* flags, int64_t ld_f, int* need_exact, cudaStream_t st);
void tc_session_free(rs_tc_session* s);

// need_exact == nullptr: flagged queries are re-run on the exact scan before returning (one stream synchronise).
// need_exact != nullptr (device int32 [nq]): nothing is synchronised; need_exact[i] = 1 marks rows of D / I that are NOT
// final and that the caller must re-search with search_exact_device.
void search_tensor_device(rs_index* h, const float* q, int64_t nq, int k, int64_t id_base, float* D, int64_t ldD,
                          int64_t* I, int64_t ldI, cudaStream_t st, int* need_exact = nullptr);
// every low-precision score of the tensor path, out [nq, ntotal] (validation of the tcgen05 tiles)
void tensor_scores_debug(rs_index* h, const float* q, int64_t nq, float* out, cudaStream_t st);
// one or two queries: exact top-k through the bf16 shadow, two launches (flat_index.cu); same results as the exact scan
bool bf16_scan_applies(const rs_index* h, int64_t nq, int k);
void search_bf16_scan_device(rs_index* h, const float* q, int64_t nq, int k, int64_t id_base, float* D, int64_t ldD, int64_t* I,
                             int64_t ldI, cudaStream_t st, int* need_exact = nullptr);
// builds / refreshes the bf16 shadow and max row norm of the index (tc_filter.cu)
void ensure_tensor_shadow(rs_index* h, cudaStream_t st);
// true when the tensor path is applicable to this shape (d, k, sizes)
bool tensor_path_supported(const rs_index* h, int64_t nq, int k);

}  // namespace rs
