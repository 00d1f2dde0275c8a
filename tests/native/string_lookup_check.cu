// string_lookup_check.cu -- TEST INFRASTRUCTURE.  Runs the product's per-id lookup routine (lookup_one) and its table
// build (build_table) from csrc/string_lookup.cu ON THE HOST, so the probing / byte-comparison logic can be checked in a
// container without a GPU (tests/test_string_lookup.py).  The product library never contains this entry point: the
// file below is compiled with RS_STRING_LOOKUP_HOST_CHECK, which leaves out the C ABI and keeps only the shared routines.
#define RS_STRING_LOOKUP_HOST_CHECK 1
#include "../../recommendation-system-maang-nvidia-_b200/csrc/string_lookup.cu"

extern "C" int string_lookup_host_check(const uint8_t* words, int64_t n_words, int64_t width, const uint8_t* ids,
                                        int64_t B, int64_t W, int64_t ld, int64_t* rows) {
  const int64_t wv = (width + 15) / 16 * 16;
  std::vector<uint32_t> slots;
  std::vector<uint8_t> pool;
  const uint32_t mask = build_table(words, n_words, width, wv, slots, pool);
  // 16-byte aligned copies, as cudaMalloc would hand out
  std::vector<uint4> slots_al((slots.size() * 4 + 15) / 16), pool_al((pool.size() + 15) / 16 + 1);
  std::memcpy(slots_al.data(), slots.data(), slots.size() * 4);
  if (!pool.empty()) std::memcpy(pool_al.data(), pool.data(), pool.size());
  LookupParams p{};
  p.ids = ids;
  p.ld = ld;
  p.B = B;
  p.W = static_cast<int>(W);
  p.id_chunks = (p.W + 15) / 16;
  p.vec = (p.W % 16 == 0) && (p.ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.ids) & 15) == 0);
  p.slots = reinterpret_cast<const uint32_t*>(slots_al.data());
  p.pool = reinterpret_cast<const uint8_t*>(pool_al.data());
  p.word_chunks = static_cast<int>(wv / 16);
  p.mask = mask;
  p.rows = rows;
  // the kernel's own pairing: ids i and i + n_threads together, a single id when the partner is past the end
  const int64_t n_threads = B > 7 ? (B + 2) / 3 : 1;
  for (int64_t base = 0; base < B; base += 2 * n_threads)
    for (int64_t t = 0; t < n_threads && base + t < B; ++t) {
      const int64_t i = base + t, i1 = i + n_threads;
      if (i1 < B) lookup_two(p, i, i1);
      else rows[i] = lookup_one(p, i);
    }
  return p.vec;
}
