// gather.cu -- K1: embedding row gather fused with the feature concat.
// Replaces keras.layers.Embedding (src/models.py:71,74,85,89) + tf.concat (src/models.py:128).
//
// HBM-bound copy: one thread moves one 16-byte chunk of an output row with a 128-bit load and a
// 128-bit store; a warp covers 512 contiguous bytes of one row, so both sides are fully coalesced.
// Algorithmic bytes per sample: sum_t (E_t * 4) read + the same written + the ids.
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
