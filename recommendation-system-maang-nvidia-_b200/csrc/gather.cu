// gather.cu -- K1: embedding row gather fused with the feature concat.
// Replaces keras.layers.Embedding (src/models.py:71,74,85,89) + tf.concat (src/models.py:128).
//
// HBM-bound copy: one thread moves one 16-byte chunk of an output row with a 128-bit load and a
// 128-bit store; a warp covers 512 contiguous bytes of one row, so both sides are fully coalesced.
// Algorithmic bytes per sample: sum_t (E_t * 4) read + the same written + the ids.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kMaxTables = 8;

struct GatherParams {
  const float* table[kMaxTables];
  const void* ids[kMaxTables];
  int64_t rows[kMaxTables];      // V_t
  int64_t ld[kMaxTables];        // table row stride in floats
  int chunk_begin[kMaxTables + 1];  // prefix sums of E_t / 4
  int ids64[kMaxTables];
  int n_tables;
  int chunks_per_row;
  int64_t batch;
  float* out;
  int64_t ld_out;
};

__global__ void __launch_bounds__(256) gather_concat_kernel(const GatherParams p) {
  const int64_t total = p.batch * p.chunks_per_row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / p.chunks_per_row;
    const int c = static_cast<int>(i - b * p.chunks_per_row);
    int t = 0;
#pragma unroll
    for (int u = 1; u < kMaxTables; ++u)
      if (u < p.n_tables && c >= p.chunk_begin[u]) t = u;
    const int64_t row = p.ids64[t] ? static_cast<const int64_t*>(p.ids[t])[b]
                                   : static_cast<int64_t>(static_cast<const int32_t*>(p.ids[t])[b]);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row >= 0 && row < p.rows[t])
      v = __ldg(reinterpret_cast<const float4*>(p.table[t] + row * p.ld[t]) + (c - p.chunk_begin[t]));
    *reinterpret_cast<float4*>(p.out + b * p.ld_out + 4 * (int64_t)c) = v;
  }
}

}  // namespace

extern "C" int rs_gather_concat(int n_tables, const DLTensor* const* tables, const DLTensor* const* ids,
                                DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(n_tables >= 1 && n_tables <= kMaxTables, "n_tables must be in [1, 8]");
    require(out, "out", 2, kDLFloat, 32);
    RS_CHECK(on_device(out), "out must be a CUDA tensor");
    GatherParams p{};
    p.n_tables = n_tables;
    p.batch = out->shape[0];
    int chunks = 0;
    for (int t = 0; t < n_tables; ++t) {
      const DLTensor* tb = tables[t];
      const DLTensor* id = ids[t];
      require(tb, "table", 2, kDLFloat, 32);
      RS_CHECK(id != nullptr && id->ndim == 1, "ids must be 1-D");
      RS_CHECK(is_dtype(id, kDLInt, 32) || is_dtype(id, kDLInt, 64), "ids must be int32 or int64");
      RS_CHECK(on_device(tb) && on_device(id), "tables and ids must be CUDA tensors");
      RS_CHECK(id->shape[0] == p.batch, "ids length must equal out rows");
      RS_CHECK(stride_of(id, 0) == 1 || id->shape[0] <= 1, "ids must be contiguous");
      RS_CHECK(tb->shape[1] % 4 == 0, "embedding width must be a multiple of 4");
      p.table[t] = data_ptr<float>(tb);
      p.ld[t] = row_stride_2d(tb, "table");
      RS_CHECK(p.ld[t] % 4 == 0, "table row stride must be a multiple of 4 floats");
      require_aligned16(p.table[t], "table");
      p.rows[t] = tb->shape[0];
      p.ids[t] = data_ptr<void>(id);
      p.ids64[t] = is_dtype(id, kDLInt, 64) ? 1 : 0;
      p.chunk_begin[t] = chunks;
      chunks += static_cast<int>(tb->shape[1] / 4);
    }
    p.chunk_begin[n_tables] = chunks;
    p.chunks_per_row = chunks;
    RS_CHECK(out->shape[1] == 4 * (int64_t)chunks, "out width must equal the sum of the table widths");
    p.out = data_ptr<float>(out);
    p.ld_out = row_stride_2d(out, "out");
    RS_CHECK(p.ld_out % 4 == 0, "out row stride must be a multiple of 4 floats");
    require_aligned16(p.out, "out");
    if (p.batch == 0) return;
    const int64_t total = p.batch * chunks;
    const int threads = 256;
    const int64_t want = (total + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    const int blocks = static_cast<int>(want < cap ? want : cap);
    gather_concat_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(p);
    RS_LAUNCH_CHECK();
  });
}

// ------------------------------------------------------------------------------------------------
// K7b. Sharded embedding tables (BASELINE config 4): rows are sharded `row mod G` over the GPUs, the batch is
// data-parallel.  A lookup is: plan (bucket the local ids by owner) -> all-to-all of the local row numbers ->
// owner-side gather (gather_concat_kernel) -> all-to-all of the rows back -> un-permute (gather_concat_kernel
// with the permutation as ids).  The two kernels below are the plan: a per-owner histogram, then a scatter of
// (row div G) into per-owner send segments that records each id's slot so the original order can be restored.
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxShards = 64;

__global__ void __launch_bounds__(256) a2a_count_kernel(const int64_t* __restrict__ ids, int64_t B, int G,
                                                        unsigned long long* __restrict__ counts) {
  __shared__ unsigned int local[kMaxShards];
  if (threadIdx.x < kMaxShards) local[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = ids[i] < 0 ? 0 : ids[i];
    atomicAdd(&local[id % G], 1u);
  }
  __syncthreads();
  if (threadIdx.x < G && local[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)local[threadIdx.x]);
}

// send[base[g] + slot] = id div G for every id owned by g; perm[i] = that position
__global__ void __launch_bounds__(256) a2a_scatter_kernel(const int64_t* __restrict__ ids, int64_t B, int G,
                                                          const int64_t* __restrict__ base,
                                                          unsigned long long* __restrict__ cursor,
                                                          int64_t* __restrict__ send, int64_t* __restrict__ perm) {
  __shared__ unsigned int local[kMaxShards];
  __shared__ unsigned long long start[kMaxShards];
  if (threadIdx.x < kMaxShards) local[threadIdx.x] = 0;
  __syncthreads();
  // one chunk of blockDim.x ids per iteration: block-aggregate the per-owner slots, one global atomic per owner
  for (int64_t c0 = blockIdx.x * (int64_t)blockDim.x; c0 < B; c0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = c0 + threadIdx.x;
    int64_t id = 0;
    int g = -1;
    unsigned int slot = 0;
    if (i < B) {
      id = ids[i] < 0 ? 0 : ids[i];
      g = static_cast<int>(id % G);
      slot = atomicAdd(&local[g], 1u);
    }
    __syncthreads();
    if (threadIdx.x < G) start[threadIdx.x] = local[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (unsigned long long)local[threadIdx.x]) : 0ull;
    __syncthreads();
    if (i < B) {
      const int64_t pos = base[g] + static_cast<int64_t>(start[g]) + slot;
      send[pos] = id / G;
      perm[i] = pos;
    }
    __syncthreads();
    if (threadIdx.x < kMaxShards) local[threadIdx.x] = 0;
    __syncthreads();
  }
}

}  // namespace

extern "C" int rs_a2a_count(const DLTensor* ids, int n_shards, DLTensor* counts, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(ids, "ids", 1, kDLInt, 64);
    require(counts, "counts", 1, kDLInt, 64);
    RS_CHECK(n_shards >= 1 && n_shards <= kMaxShards, "n_shards must be in [1, 64]");
    RS_CHECK(on_device(ids) && on_device(counts) && counts->shape[0] == n_shards, "ids / counts must be CUDA tensors, counts [n_shards]");
    RS_CHECK(ids->shape[0] <= 1 || stride_of(ids, 0) == 1, "ids must be contiguous");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RS_CUDA(cudaMemsetAsync(data_ptr<void>(counts), 0, sizeof(int64_t) * n_shards, st));
    const int64_t B = ids->shape[0];
    if (B == 0) return;
    const int blocks = static_cast<int>(std::min<int64_t>((B + 255) / 256, (int64_t)sm_count() * 8));
    a2a_count_kernel<<<blocks, 256, 0, st>>>(data_ptr<int64_t>(ids), B, n_shards, data_ptr<unsigned long long>(counts));
    RS_LAUNCH_CHECK();
  });
}

extern "C" int rs_a2a_scatter(const DLTensor* ids, int n_shards, const DLTensor* base, DLTensor* cursor, DLTensor* send,
                              DLTensor* perm, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(ids, "ids", 1, kDLInt, 64);
    require(base, "base", 1, kDLInt, 64);
    require(cursor, "cursor", 1, kDLInt, 64);
    require(send, "send", 1, kDLInt, 64);
    require(perm, "perm", 1, kDLInt, 64);
    RS_CHECK(n_shards >= 1 && n_shards <= kMaxShards, "n_shards must be in [1, 64]");
    const int64_t B = ids->shape[0];
    RS_CHECK(on_device(ids) && on_device(base) && on_device(cursor) && on_device(send) && on_device(perm), "all tensors must be CUDA tensors");
    RS_CHECK(base->shape[0] == n_shards && cursor->shape[0] == n_shards && send->shape[0] == B && perm->shape[0] == B, "shape mismatch");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RS_CUDA(cudaMemsetAsync(data_ptr<void>(cursor), 0, sizeof(int64_t) * n_shards, st));
    if (B == 0) return;
    const int blocks = static_cast<int>(std::min<int64_t>((B + 255) / 256, (int64_t)sm_count() * 8));
    a2a_scatter_kernel<<<blocks, 256, 0, st>>>(data_ptr<int64_t>(ids), B, n_shards, data_ptr<int64_t>(base),
                                               data_ptr<unsigned long long>(cursor), data_ptr<int64_t>(send),
                                               data_ptr<int64_t>(perm));
    RS_LAUNCH_CHECK();
  });
}

// Synthetic table rows, replayable on the host: T[r, c] = (u32(mix(seed, r * dim + c)) * 2^-32 - 0.5) * 0.1
namespace {
__host__ __device__ inline uint32_t mix32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return static_cast<uint32_t>(x);
}
__global__ void __launch_bounds__(256) hash_table_kernel(float* __restrict__ out, int64_t n_local, int dim, int64_t row0,
                                                         int64_t row_step, uint64_t seed) {
  const int64_t total = n_local * dim;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lr = i / dim;
    const int c = static_cast<int>(i - lr * dim);
    const uint64_t g = static_cast<uint64_t>(row0 + lr * row_step) * dim + c;
    out[i] = (static_cast<float>(mix32(seed * 0x9E3779B97F4A7C15ULL + g)) * 2.3283064365386963e-10f - 0.5f) * 0.1f;
  }
}
}  // namespace

// rows row0, row0 + row_step, ... of the synthetic table into out [n_local, dim] (the shard `row mod G == rank`)
extern "C" int rs_hash_table_fill(DLTensor* out, int64_t row0, int64_t row_step, uint64_t seed, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(out, "out", 2, kDLFloat, 32);
    RS_CHECK(on_device(out) && is_compact(out), "out must be a contiguous CUDA tensor");
    const int64_t n = out->shape[0];
    if (n == 0) return;
    const int64_t total = n * out->shape[1];
    const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 32));
    hash_table_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(data_ptr<float>(out), n, (int)out->shape[1], row0,
                                                                             row_step, seed);
    RS_LAUNCH_CHECK();
  });
}
