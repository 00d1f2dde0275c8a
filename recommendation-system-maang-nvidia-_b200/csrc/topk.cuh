// topk.cuh -- CTA-level top-k candidate buffer in shared memory.
//
// A buffer holds up to CAP packed keys (common.cuh: larger key == better hit).  Threads append
// candidates that beat the current threshold key with a shared-memory atomic; when the buffer fills the
// whole CTA sorts it (bitonic, descending), keeps the best k and raises the threshold to the k-th key.
// Only candidates better than the running k-th ever reach the buffer, so after warm-up appends are
// rare (about k*ln(n/k) per n scanned rows) and the common per-tile cost is one comparison per score.
#pragma once

#include "common.cuh"

namespace rsd {

// Collective over the CTA.  Sorts keys[0..CAP), keeps the best k, returns the new threshold key
// (0 while fewer than k hits are held).  On return *cnt == min(valid, k) and all threads are synced.
template <int CAP, int NT>
__device__ __forceinline__ uint64_t prune_buffer(uint64_t* keys, int* cnt, int k, int tid) {
  const int n_valid = min(*cnt, CAP);
  for (int i = n_valid + tid; i < CAP; i += NT) keys[i] = kEmptyKey;
  __syncthreads();
  block_sort_desc<CAP, NT>(keys, tid);
  const int new_cnt = min(n_valid, k);
  const uint64_t thr = (new_cnt >= k) ? keys[k - 1] : kEmptyKey;
  __syncthreads();
  if (tid == 0) *cnt = new_cnt;
  __syncthreads();
  return thr;
}

// Collective: every thread offers at most one key to one buffer.  `thr` is the caller's register copy
// of the threshold key and is refreshed when a prune happens.  Returns true if any thread of the CTA
// offered a key in this call (callers use it for early termination).
template <int CAP, int NT>
__device__ __forceinline__ bool offer_one(uint64_t* keys, int* cnt, int k, bool want, uint64_t key,
                                          uint64_t& thr, int tid) {
  bool pend = want && key > thr;
  const bool any_offer = __syncthreads_or(pend) != 0;
  if (!any_offer) return false;
  for (;;) {
    bool saw_full = false;
    if (pend) {
      const int pos = atomicAdd(cnt, 1);
      if (pos < CAP) {
        keys[pos] = key;
        pend = false;
      }
      saw_full = pos >= CAP - 1;
    }
    if (!__syncthreads_or(pend || saw_full)) break;
    thr = prune_buffer<CAP, NT>(keys, cnt, k, tid);
    pend = pend && key > thr;
  }
  return true;
}

}  // namespace rsd
