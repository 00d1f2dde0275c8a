// string_lookup.cu -- K0: keras.layers.StringLookup(vocabulary=v, mask_token=None) on the device
// (src/models.py:70,73, used at :85,89): string id -> row number, v[i] -> i + 1, anything else -> 0.
//
// Ids arrive as a fixed-width byte matrix uint8 [B, W] (NUL-padded rows: numpy's `S` dtype, an Arrow fixed-size binary
// column) -- byte and integer work only.  The vocabulary is an open-addressing table in HBM built once on the host:
//   slots uint32 [cap]      cap = power of two >= 2 * n_words; 0 = empty, else word number + 1
//   pool  uint8  [n, Wv]    the words themselves, NUL-padded, Wv a multiple of 16 (16-byte loads)
// One thread per id: the row is read 16 bytes at a time, everything from the first NUL on is masked off with word-wide bit
// tricks (registers only), the masked chunks are hashed by a 64-bit multiply-xorshift chain, then linear probing from
// hash & (cap - 1) -- and every occupied slot on the way is settled by comparing the WORD BYTES (16 at a time), never
// the hash alone, so a hash collision can cost a probe but cannot return a wrong row.  Load factor <= 0.5 -> ~1.5 probes.
// Algorithmic bytes per id: W read + 8 written + ~1.5 x (4 + Wv) of table (L2-resident up to ~2M words; beyond that
// each probe is a random 32-byte HBM sector pair).  A warp reads 32 consecutive rows of the id matrix: coalesced.
#include <cstring>
#include <vector>

#include "common.cuh"

#ifndef RS_STRING_LOOKUP_HOST_CHECK
struct rs_vocab {
  int device = 0;
  int64_t n_words = 0;
  int64_t width = 0;      // Wv, multiple of 16
  uint32_t mask = 0;      // cap - 1
  uint32_t* slots = nullptr;
  uint8_t* pool = nullptr;
};
#endif

namespace {

constexpr uint64_t kSeed = 0xCBF29CE484222325ull;
constexpr uint64_t kMulLo = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kMulHi = 0xC2B2AE3D27D4EB4Full;

__host__ __device__ __forceinline__ uint32_t fold(uint64_t h) { return static_cast<uint32_t>(h ^ (h >> 32)); }

// The per-id routine is __host__ __device__ so that tests/native/string_lookup_check.cu can run the very same masking,
// hashing, probing and comparison code on the host (no GPU in the build container); the product only launches the kernel.
__host__ __device__ __forceinline__ uint4 load16(const uint8_t* row, int c) {
#ifdef __CUDA_ARCH__
  return __ldg(reinterpret_cast<const uint4*>(row) + c);
#else
  uint4 v;
  std::memcpy(&v, row + 16 * c, 16);
  return v;
#endif
}

__host__ __device__ __forceinline__ int first_set_bit(uint32_t x) {   // index of the lowest set bit, x != 0
#ifdef __CUDA_ARCH__
  return __ffs(static_cast<int>(x)) - 1;
#else
  return __builtin_ctz(x);
#endif
}

// One 32-bit word of a row (4 bytes, little endian).  Keeps the bytes before the first NUL, zeroes the rest; `ended`
// carries "a NUL was seen" from word to word.  The zero-byte detector (w - 0x01010101) & ~w & 0x80808080 can flag
// bytes ABOVE the first zero byte by borrow, never below it, so its lowest set bit is exact.
__host__ __device__ __forceinline__ uint32_t keep_before_nul(uint32_t w, bool& ended) {
  if (ended) return 0u;
  const uint32_t z = (w - 0x01010101u) & ~w & 0x80808080u;
  if (z == 0u) return w;
  ended = true;
  const int byte = first_set_bit(z) >> 3;               // 0..3
  return byte == 0 ? 0u : (w & (0xFFFFFFFFu >> (8 * (4 - byte))));
}

__host__ __device__ __forceinline__ uint4 keep_before_nul(uint4 v, bool& ended) {
  v.x = keep_before_nul(v.x, ended);
  v.y = keep_before_nul(v.y, ended);
  v.z = keep_before_nul(v.z, ended);
  v.w = keep_before_nul(v.w, ended);
  return v;
}

// Hash of a string = a multiply-xorshift chain over its NUL-masked 16-byte chunks (two 64-bit lanes per chunk).  Only
// chunks that hold string bytes take part, so the value does not depend on how wide the row around the string is.
__host__ __device__ __forceinline__ uint64_t mix_chunk(uint64_t h, const uint4& v) {
  const uint64_t lo = static_cast<uint64_t>(v.x) | (static_cast<uint64_t>(v.y) << 32);
  const uint64_t hi = static_cast<uint64_t>(v.z) | (static_cast<uint64_t>(v.w) << 32);
  h = (h ^ lo) * kMulLo;
  h ^= h >> 32;
  h = (h ^ hi) * kMulHi;
  h ^= h >> 29;
  return h;
}

struct LookupParams {
  const uint8_t* ids;     // [B, W], row stride ld
  int64_t ld;
  int64_t B;
  int W;                  // bytes per id row
  int id_chunks;          // ceil(W / 16)
  int vec;                // rows are 16-byte aligned and W % 16 == 0: 128-bit loads
  const uint32_t* slots;
  const uint8_t* pool;
  int word_chunks;        // Wv / 16
  uint32_t mask;
  int64_t* rows;
};

__host__ __device__ __forceinline__ uint4 id_chunk(const LookupParams& p, const uint8_t* row, int c) {
  if (p.vec) return load16(row, c);
  uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;    // unaligned / odd-width id matrices: byte loads, zero past the row
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    const int at = 16 * c + b;
    const uint32_t byte = at < p.W ? static_cast<uint32_t>(row[at]) << (8 * (b & 3)) : 0u;
    if ((b >> 2) == 0) w0 |= byte;
    else if ((b >> 2) == 1) w1 |= byte;
    else if ((b >> 2) == 2) w2 |= byte;
    else w3 |= byte;
  }
  return make_uint4(w0, w1, w2, w3);
}

__host__ __device__ __forceinline__ bool any_byte(const uint4& v) { return (v.x | v.y | v.z | v.w) != 0u; }

// hash and chunk count of a NUL-padded row of n_chunks aligned chunks (the table build hashes the pool rows with it)
__host__ __device__ __forceinline__ uint64_t hash_row(const uint8_t* row, int n_chunks) {
  uint64_t h = kSeed;
  bool ended = false;
  for (int c = 0; c < n_chunks && !ended; ++c) {
    const uint4 v = keep_before_nul(load16(row, c), ended);
    if (!any_byte(v)) break;
    h = mix_chunk(h, v);
  }
  return h;
}

__host__ __device__ __forceinline__ int64_t lookup_one(const LookupParams& p, int64_t i) {
  const uint8_t* row = p.ids + i * p.ld;
  // pass 1: hash and length in chunks (bytes from the first NUL on do not belong to the string)
  uint64_t h = kSeed;
  int n_used = 0;                                   // chunks holding string bytes
  bool ended = false;
  uint4 first = make_uint4(0u, 0u, 0u, 0u);         // chunk 0, masked: most ids end inside it
  for (int c = 0; c < p.id_chunks && !ended; ++c) {
    const uint4 v = keep_before_nul(id_chunk(p, row, c), ended);
    if (!any_byte(v)) break;
    h = mix_chunk(h, v);
    n_used = c + 1;
    if (c == 0) first = v;
  }
  if (n_used > p.word_chunks) return 0;             // longer than every word: out of vocabulary
  uint32_t slot = fold(h) & p.mask;
  for (uint32_t probe = 0; probe <= p.mask; ++probe, slot = (slot + 1) & p.mask) {
#ifdef __CUDA_ARCH__
    const uint32_t e = __ldg(p.slots + slot);
#else
    const uint32_t e = p.slots[slot];
#endif
    if (e == 0u) return 0;                          // an empty slot ends the probe sequence: not in the vocabulary
    const uint8_t* word = p.pool + static_cast<int64_t>(e - 1) * (16 * (int64_t)p.word_chunks);
    bool same = true;
    bool tail_ended = false;
    for (int c = 0; c < p.word_chunks && same; ++c) {
      const uint4 a = load16(word, c);
      uint4 b = make_uint4(0u, 0u, 0u, 0u);
      if (c == 0) b = first;
      else if (c < n_used) b = keep_before_nul(id_chunk(p, row, c), tail_ended);
      same = a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
    }
    if (same) return static_cast<int64_t>(e);
  }
  return 0;
}

__global__ void __launch_bounds__(256) string_lookup_kernel(const LookupParams p) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < p.B; i += (int64_t)gridDim.x * blockDim.x)
    p.rows[i] = lookup_one(p, i);
}

// Host-side table build: slots (cap entries) and the NUL-padded word pool (n_words x wv bytes).
void build_table(const uint8_t* words, int64_t n_words, int64_t width, int64_t wv, std::vector<uint32_t>& slots,
                 std::vector<uint8_t>& pool) {
  uint64_t cap = 16;
  while (cap < 2 * static_cast<uint64_t>(n_words)) cap <<= 1;
  slots.assign(cap, 0u);
  pool.assign(static_cast<size_t>(n_words * wv), 0);
  const uint32_t mask = static_cast<uint32_t>(cap - 1);
  for (int64_t i = 0; i < n_words; ++i) {
    const uint8_t* s = words + i * width;
    int64_t len = 0;
    while (len < width && s[len]) ++len;
    uint8_t* dst = pool.data() + i * wv;
    std::memcpy(dst, s, static_cast<size_t>(len));                    // bytes after the first NUL stay zero
    uint32_t slot = fold(hash_row(dst, static_cast<int>(wv / 16))) & mask;
    for (;; slot = (slot + 1) & mask) {
      if (slots[slot] == 0u) { slots[slot] = static_cast<uint32_t>(i + 1); break; }
      // a repeated word keeps its LAST position, as the host dict {word: i + 1} does
      if (std::memcmp(pool.data() + static_cast<int64_t>(slots[slot] - 1) * wv, dst, static_cast<size_t>(wv)) == 0) {
        slots[slot] = static_cast<uint32_t>(i + 1);
        break;
      }
    }
  }
}

}  // namespace

#ifndef RS_STRING_LOOKUP_HOST_CHECK
extern "C" int rs_vocab_create(const uint8_t* words, int64_t n_words, int64_t width, rs_vocab_t** out) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(out != nullptr, "out is NULL");
    RS_CHECK(n_words >= 0 && n_words < 0x7FFFFFFFll, "n_words must be in [0, 2^31)");
    RS_CHECK(width >= 1 && width <= 4096, "width must be in [1, 4096] bytes");
    RS_CHECK(words != nullptr || n_words == 0, "words is NULL");
    const int64_t wv = (width + 15) / 16 * 16;
    std::vector<uint32_t> slots;
    std::vector<uint8_t> pool;
    build_table(words, n_words, width, wv, slots, pool);
    const uint64_t cap = slots.size();
    const uint32_t mask = static_cast<uint32_t>(cap - 1);
    auto v = new rs_vocab();
    try {
      RS_CUDA(cudaGetDevice(&v->device));
      v->n_words = n_words;
      v->width = wv;
      v->mask = mask;
      RS_CUDA(cudaMalloc(&v->slots, cap * sizeof(uint32_t)));
      RS_CUDA(cudaMalloc(&v->pool, pool.size() ? pool.size() : 16));
      RS_CUDA(cudaMemcpy(v->slots, slots.data(), cap * sizeof(uint32_t), cudaMemcpyHostToDevice));
      if (!pool.empty()) RS_CUDA(cudaMemcpy(v->pool, pool.data(), pool.size(), cudaMemcpyHostToDevice));
    } catch (...) {
      cudaFree(v->slots);
      cudaFree(v->pool);
      delete v;
      throw;
    }
    *out = v;
  });
}

extern "C" int rs_vocab_free(rs_vocab_t* v) {
  return rs::guard([&] {
    if (!v) return;
    cudaFree(v->slots);
    cudaFree(v->pool);
    delete v;
  });
}

extern "C" int rs_vocab_size(const rs_vocab_t* v, int64_t* n_words) {
  return rs::guard([&] {
    RS_CHECK(v != nullptr && n_words != nullptr, "NULL argument");
    *n_words = v->n_words;
  });
}

extern "C" int rs_vocab_lookup(const rs_vocab_t* v, const DLTensor* ids, DLTensor* rows, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(v != nullptr, "vocabulary handle is NULL");
    require(ids, "ids", 2, kDLUInt, 8);
    require(rows, "rows", 1, kDLInt, 64);
    RS_CHECK(on_device(ids) && on_device(rows), "ids and rows must be CUDA tensors");
    RS_CHECK(ids->device.device_id == v->device && rows->device.device_id == v->device,
             "ids / rows live on another device than the vocabulary");
    RS_CHECK(rows->shape[0] == ids->shape[0], "rows length must equal the number of ids");
    if (ids->shape[0] == 0) return;                 // nothing to do (an empty tensor's strides carry no meaning)
    RS_CHECK(rows->shape[0] <= 1 || stride_of(rows, 0) == 1, "rows must be contiguous");
    RS_CHECK(ids->shape[1] >= 1 && ids->shape[1] <= 4096, "id width must be in [1, 4096] bytes");
    LookupParams p{};
    p.ids = data_ptr<uint8_t>(ids);
    p.ld = row_stride_2d(ids, "ids");
    p.B = ids->shape[0];
    p.W = static_cast<int>(ids->shape[1]);
    p.id_chunks = (p.W + 15) / 16;
    p.vec = (p.W % 16 == 0) && (p.ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.ids) & 15) == 0);
    p.slots = v->slots;
    p.pool = v->pool;
    p.word_chunks = static_cast<int>(v->width / 16);
    p.mask = v->mask;
    p.rows = data_ptr<int64_t>(rows);
    if (p.B == 0) return;
    const int threads = 256;
    const int64_t want = (p.B + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 8;
    string_lookup_kernel<<<static_cast<int>(want < cap ? want : cap), threads, 0, static_cast<cudaStream_t>(stream)>>>(p);
    RS_LAUNCH_CHECK();
  });
}
#endif  // RS_STRING_LOOKUP_HOST_CHECK
