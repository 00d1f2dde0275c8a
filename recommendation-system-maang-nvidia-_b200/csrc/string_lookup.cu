// string_lookup.cu -- K0: keras.layers.StringLookup(vocabulary=v, mask_token=None) on the device
// (src/models.py:70,73, used at :85,89): string id -> row number, v[i] -> i + 1, anything else -> 0.
//
// Ids arrive as a fixed-width byte matrix uint8 [B, W] (NUL-padded rows: numpy's `S` dtype, an Arrow fixed-size binary
// column) -- byte and integer work only.  The vocabulary is a bucketed open-addressing table in HBM built once on the host:
//   slots uint32 [cap][2]   (word number + 1, 32-bit tag of the hash); 0 = empty.  cap = power of two >= 4 * n_words,
//                           grouped into buckets of 2 slots = 16 bytes (one 128-bit load); a bucket fills left to right
//   pool  uint8  [n, Wv]    the words themselves, NUL-padded, Wv a multiple of 16 (16-byte loads)
// Per id: the row is read 16 bytes at a time, everything from the first NUL on is masked off with word-wide bit tricks
// (registers only), the masked chunks are hashed by a 64-bit multiply-xorshift chain; ONE 16-byte load fetches the id's
// home bucket, slots whose tag matches are settled by comparing the WORD BYTES (16 at a time) -- never the hash alone, so a
// collision can cost a load but cannot return a wrong row -- and the first empty slot ends the search.  At load <= 0.25 a
// bucket overflows into the next one for 1.4 % of the words.  A thread works on two ids at once (their row loads and
// their bucket loads are issued back to back): the kernel is bound by the latency of the dependent loads
// (row -> bucket -> word), not by bytes.
// Algorithmic bytes per id: W read + 8 written; + 16 B of bucket and Wv of word, L2-resident up to ~2M words.
#include <cstring>
#include <vector>

#include "common.cuh"

#ifndef RS_STRING_LOOKUP_HOST_CHECK
struct rs_vocab {
  int device = 0;
  int64_t n_words = 0;
  int64_t width = 0;      // Wv, multiple of 16
  uint32_t mask = 0;      // number of buckets - 1
  uint32_t* slots = nullptr;   // [cap][2]
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
  uint32_t mask;          // number of buckets - 1
  int64_t* rows;
};

constexpr int kBucketSlots = 2;                        // 2 x (row + 1, tag) = 16 bytes = one 128-bit load

struct Bucket {
  uint4 v;                                             // x, y = slot 0 (row + 1, tag); z, w = slot 1
};

__host__ __device__ __forceinline__ Bucket load_bucket(const LookupParams& p, uint32_t b) {
  Bucket k;
  k.v = load16(reinterpret_cast<const uint8_t*>(p.slots) + static_cast<int64_t>(b) * 16, 0);
  return k;
}

// RS_STRING_LOOKUP_TAG_MASK narrows the tag in the HOST CHECK builds only (tests/native: 2-bit tags make tag collisions
// between different words common, which drives the byte comparison's fall-back path); the product keeps all 32 bits.
#ifndef RS_STRING_LOOKUP_TAG_MASK
#define RS_STRING_LOOKUP_TAG_MASK 0xFFFFFFFFu
#endif
__host__ __device__ __forceinline__ uint32_t tag_of(uint64_t h) {
  return static_cast<uint32_t>(h >> 32) & RS_STRING_LOOKUP_TAG_MASK;
}

// unaligned / odd-width id matrices: byte loads, zero past the row (kept out of line: the aligned path is the fast one)
__host__ __device__ __noinline__ uint4 id_chunk_bytes(const uint8_t* row, int c, int W) {
  uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    const int at = 16 * c + b;
    const uint32_t byte = at < W ? static_cast<uint32_t>(row[at]) << (8 * (b & 3)) : 0u;
    if ((b >> 2) == 0) w0 |= byte;
    else if ((b >> 2) == 1) w1 |= byte;
    else if ((b >> 2) == 2) w2 |= byte;
    else w3 |= byte;
  }
  return make_uint4(w0, w1, w2, w3);
}

__host__ __device__ __forceinline__ uint4 id_chunk(const LookupParams& p, const uint8_t* row, int c) {
  return p.vec ? load16(row, c) : id_chunk_bytes(row, c, p.W);
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

// What pass 1 leaves of an id: its hash, the number of 16-byte chunks that hold string bytes, and chunk 0 (masked).
struct Prepared {
  uint64_t h;
  int n_used;
  uint4 first;
};

// pass 1, given the id's first chunk (already loaded): hash and length in chunks
__host__ __device__ __forceinline__ Prepared prepare(const LookupParams& p, const uint8_t* row, uint4 chunk0) {
  Prepared r;
  r.h = kSeed;
  r.n_used = 0;
  r.first = make_uint4(0u, 0u, 0u, 0u);
  bool ended = false;
  for (int c = 0; c < p.id_chunks && !ended; ++c) {
    const uint4 v = keep_before_nul(c == 0 ? chunk0 : id_chunk(p, row, c), ended);
    if (!any_byte(v)) break;
    r.h = mix_chunk(r.h, v);
    r.n_used = c + 1;
    if (c == 0) r.first = v;
  }
  return r;
}

__host__ __device__ __forceinline__ const uint8_t* word_at(const LookupParams& p, uint32_t e) {
  return p.pool + static_cast<int64_t>(e - 1) * (16 * (int64_t)p.word_chunks);
}

__host__ __device__ __forceinline__ bool same16(const uint4& a, const uint4& b) {
  return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}

// chunks 1.. of word e against the id (chunk 0 is compared by the caller)
__host__ __device__ __forceinline__ bool tail_equals(const LookupParams& p, const uint8_t* row, const Prepared& q, uint32_t e) {
  const uint8_t* word = word_at(p, e);
  bool same = true;
  bool tail_ended = false;
  for (int c = 1; c < p.word_chunks && same; ++c) {
    uint4 b = make_uint4(0u, 0u, 0u, 0u);
    if (c < q.n_used) b = keep_before_nul(id_chunk(p, row, c), tail_ended);
    same = same16(load16(word, c), b);
  }
  return same;
}

__host__ __device__ __forceinline__ bool word_equals(const LookupParams& p, const uint8_t* row, const Prepared& q, uint32_t e) {
  return same16(load16(word_at(p, e), 0), q.first) && tail_equals(p, row, q, e);
}

// Sequential walk from the id's home bucket: every slot whose tag matches is settled on the bytes; the first empty slot
// ends the search.  Complete on its own -- the kernel only comes here when the one-shot pick below does not settle an id
// (home bucket full without a match: 1.4 % of the words; a tag that matched another word: ~1e-9).  Inlined on purpose: as
// a call it forced the prepared id through a stack frame, and that local-memory traffic cost more than the walk saves.
__host__ __device__ __forceinline__ int64_t finish_seq(const LookupParams& p, const uint8_t* row, const Prepared& q) {
  const uint32_t tag = tag_of(q.h);
  uint32_t b = fold(q.h) & p.mask;
  for (uint32_t probe = 0; probe <= p.mask; ++probe, b = (b + 1) & p.mask) {
    const Bucket k = load_bucket(p, b);
    const uint32_t e[kBucketSlots] = {k.v.x, k.v.z};
    const uint32_t t[kBucketSlots] = {k.v.y, k.v.w};
#pragma unroll
    for (int j = 0; j < kBucketSlots; ++j) {
      if (e[j] == 0u) return 0;                     // slots fill left to right: nothing further on
      if (t[j] == tag && word_equals(p, row, q, e[j])) return static_cast<int64_t>(e[j]);
    }
  }
  return 0;
}

// One look at the home bucket, no dependent load: the first slot (before any empty one) whose tag matches.
//   state 0: an empty slot came first -> not in the vocabulary;  1: `cand` is the candidate;  2: bucket full, no match
struct Pick {
  uint32_t cand;
  int state;
};

__host__ __device__ __forceinline__ Pick pick(const Bucket& k, uint32_t tag) {
  Pick r;
  r.cand = 0u;
  r.state = 2;
  const uint32_t e[kBucketSlots] = {k.v.x, k.v.z};
  const uint32_t t[kBucketSlots] = {k.v.y, k.v.w};
#pragma unroll
  for (int j = kBucketSlots - 1; j >= 0; --j) {     // right to left, so the leftmost event wins
    if (e[j] == 0u) { r.state = 0; r.cand = 0u; }
    else if (t[j] == tag) { r.state = 1; r.cand = e[j]; }
  }
  return r;
}

// settle one id from its pick and the candidate's first 16 word bytes (already loaded)
__host__ __device__ __forceinline__ int64_t conclude(const LookupParams& p, const uint8_t* row, const Prepared& q,
                                                     const Pick& k, const uint4& word0) {
  if (q.n_used > p.word_chunks) return 0;           // longer than every word: out of vocabulary
  if (k.state == 0) return 0;
  if (k.state == 1 && same16(word0, q.first) && tail_equals(p, row, q, k.cand)) return static_cast<int64_t>(k.cand);
  return finish_seq(p, row, q);
}

__host__ __device__ __forceinline__ int64_t lookup_one(const LookupParams& p, int64_t i) {
  const uint8_t* row = p.ids + i * p.ld;
  const Prepared q = prepare(p, row, id_chunk(p, row, 0));
  const Pick k = pick(load_bucket(p, fold(q.h) & p.mask), tag_of(q.h));
  const uint4 w = k.state == 1 ? load16(word_at(p, k.cand), 0) : make_uint4(0u, 0u, 0u, 0u);
  return conclude(p, row, q, k, w);
}

// Two ids per thread: the two row loads are in flight together, then the two bucket loads, then the two word loads --
// three dependent round trips per pair.
__host__ __device__ __forceinline__ void lookup_two(const LookupParams& p, int64_t i0, int64_t i1) {
  const uint8_t* row0 = p.ids + i0 * p.ld;
  const uint8_t* row1 = p.ids + i1 * p.ld;
  const uint4 c0 = id_chunk(p, row0, 0);
  const uint4 c1 = id_chunk(p, row1, 0);
  const Prepared q0 = prepare(p, row0, c0);
  const Prepared q1 = prepare(p, row1, c1);
  const Bucket b0 = load_bucket(p, fold(q0.h) & p.mask);
  const Bucket b1 = load_bucket(p, fold(q1.h) & p.mask);
  const Pick k0 = pick(b0, tag_of(q0.h));
  const Pick k1 = pick(b1, tag_of(q1.h));
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const uint4 w0 = k0.state == 1 ? load16(word_at(p, k0.cand), 0) : zero;
  const uint4 w1 = k1.state == 1 ? load16(word_at(p, k1.cand), 0) : zero;
  p.rows[i0] = conclude(p, row0, q0, k0, w0);
  p.rows[i1] = conclude(p, row1, q1, k1, w1);
}

__global__ void __launch_bounds__(256) string_lookup_kernel(const LookupParams p) {
  const int64_t n_threads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < p.B; i += 2 * n_threads) {
    const int64_t i1 = i + n_threads;               // a warp's two sets of ids are each 32 consecutive rows
    if (i1 < p.B) lookup_two(p, i, i1);
    else p.rows[i] = lookup_one(p, i);
  }
}

// Host-side table build: slots ([cap][2] uint32, buckets of 4) and the NUL-padded word pool (n_words x wv bytes).
// Returns the bucket mask (number of buckets - 1).
uint32_t build_table(const uint8_t* words, int64_t n_words, int64_t width, int64_t wv, std::vector<uint32_t>& slots,
                     std::vector<uint8_t>& pool) {
  uint64_t cap = 16;
  while (cap < 4 * static_cast<uint64_t>(n_words)) cap <<= 1;
  slots.assign(2 * cap, 0u);
  pool.assign(static_cast<size_t>(n_words * wv), 0);
  const uint32_t mask = static_cast<uint32_t>(cap / kBucketSlots - 1);
  for (int64_t i = 0; i < n_words; ++i) {
    const uint8_t* s = words + i * width;
    int64_t len = 0;
    while (len < width && s[len]) ++len;
    uint8_t* dst = pool.data() + i * wv;
    std::memcpy(dst, s, static_cast<size_t>(len));                    // bytes after the first NUL stay zero
    const uint64_t h = hash_row(dst, static_cast<int>(wv / 16));
    const uint32_t tag = tag_of(h);
    bool placed = false;
    for (uint32_t b = fold(h) & mask; !placed; b = (b + 1) & mask) {
      for (int j = 0; j < kBucketSlots && !placed; ++j) {
        uint32_t* slot = slots.data() + (static_cast<size_t>(b) * kBucketSlots + j) * 2;
        if (slot[0] == 0u) {
          slot[0] = static_cast<uint32_t>(i + 1);
          slot[1] = tag;
          placed = true;
        } else if (slot[1] == tag &&
                   std::memcmp(pool.data() + static_cast<int64_t>(slot[0] - 1) * wv, dst, static_cast<size_t>(wv)) == 0) {
          slot[0] = static_cast<uint32_t>(i + 1);     // a repeated word keeps its LAST position, as the host dict does
          placed = true;
        }
      }
    }
  }
  return mask;
}

}  // namespace

#ifndef RS_STRING_LOOKUP_HOST_CHECK
extern "C" int rs_vocab_create(const uint8_t* words, int64_t n_words, int64_t width, rs_vocab_t** out) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(out != nullptr, "out is NULL");
    RS_CHECK(n_words >= 0 && n_words <= 0x10000000ll, "n_words must be in [0, 2^28]");
    RS_CHECK(width >= 1 && width <= 4096, "width must be in [1, 4096] bytes");
    RS_CHECK(words != nullptr || n_words == 0, "words is NULL");
    const int64_t wv = (width + 15) / 16 * 16;
    std::vector<uint32_t> slots;
    std::vector<uint8_t> pool;
    const uint32_t mask = build_table(words, n_words, width, wv, slots, pool);
    auto v = new rs_vocab();
    try {
      RS_CUDA(cudaGetDevice(&v->device));
      v->n_words = n_words;
      v->width = wv;
      v->mask = mask;
      RS_CUDA(cudaMalloc(&v->slots, slots.size() * sizeof(uint32_t)));
      RS_CUDA(cudaMalloc(&v->pool, pool.size() ? pool.size() : 16));
      RS_CUDA(cudaMemcpy(v->slots, slots.data(), slots.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
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
