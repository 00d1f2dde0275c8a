// peer.cuh -- the peer-memory region handle and the device-side protocol helpers shared by peer.cu (publish / wait / ack,
// sharded embedding lookups) and sharded.cu (the sharded catalog search sequenced in C++, whose consumer kernels wait and
// acknowledge by themselves).  See peer.cu for the protocol.
#pragma once

#include <cstdio>
#include <cstdlib>

#include "common.cuh"

struct rs_peer {
  int rank = 0, world = 0, device = 0;
  int64_t slot_bytes = 0;
  size_t total = 0;
  uint8_t* local = nullptr;
  void* bases[16] = {};          // mapped regions (bases[rank] == local)
  bool opened[16] = {};
  uint8_t** d_bases = nullptr;   // device copy of bases
  bool connected = false;
  ~rs_peer() {
    for (int g = 0; g < world; ++g)
      if (opened[g]) cudaIpcCloseMemHandle(bases[g]);
    cudaFree(d_bases);
    cudaFree(local);
  }
};

namespace rsp {

constexpr int kMaxWorld = 16;

// makes the region's device current for the lifetime of the object (launches must target the device that owns it)
struct PeerDeviceGuard {
  int prev = -1;
  explicit PeerDeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~PeerDeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct PeerLayout {
  int64_t slot_bytes;
  int64_t off_flags, off_acks, off_seq, off_done;
  int world, rank;
  unsigned long long spin_budget;   // iterations of ~100 ns a kernel may wait for a peer before it traps
};

// Spin budget in iterations of ~100 ns (ld.acquire.sys + nanosleep): default 2^27 (over a minute -- ranks can be seconds
// apart on a first step, after a GC pause or while one side runs the host-synchronised exact fallback); RS_PEER_TIMEOUT_S
// (seconds) overrides it.  A peer that never arrives still ends in a trap with a message rather than a hung device.
inline unsigned long long spin_budget_from_env() {
  static const unsigned long long budget = [] {
    if (const char* e = getenv("RS_PEER_TIMEOUT_S")) {
      const double sec = atof(e);
      if (sec > 0) return static_cast<unsigned long long>(sec * 1.0e7);
    }
    return 1ull << 27;
  }();
  return budget;
}

inline PeerLayout layout_of(const rs_peer* h) {
  PeerLayout L{};
  L.slot_bytes = h->slot_bytes;
  L.world = h->world;
  L.rank = h->rank;
  L.off_flags = h->slot_bytes * h->world;
  L.off_acks = L.off_flags + 8 * kMaxWorld;
  L.off_seq = L.off_acks + 8 * kMaxWorld;
  L.off_done = L.off_seq + 8;
  L.spin_budget = spin_budget_from_env();
  return L;
}

__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Signal words go out as RELAXED system-scope stores behind ONE system-scope fence of the signalling thread (fence + relaxed
// store is a release pattern).  A release store per peer made the signalling thread wait for each previous store's round
// trip before issuing the next one -- `world` NVLink round trips in a row on the critical path of every exchange.
__device__ __forceinline__ void st_relaxed_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// spins until *p >= want
__device__ __forceinline__ void spin_until(const uint64_t* p, uint64_t want, unsigned long long budget, const char* what) {
  for (unsigned long long it = 0; ld_acquire_sys(p) < want; ++it) {
    __nanosleep(64);
    if (it > budget) {
      printf("recsys_b200: peer exchange timed out waiting for %s (have %llu, want %llu)\n", what,
             (unsigned long long)ld_acquire_sys(p), (unsigned long long)want);
      __trap();
    }
  }
}

// Every CTA of a publishing kernel: the step number this launch will publish, after every peer has acknowledged the
// previous step (its slot in their regions may then be overwritten).
__device__ __forceinline__ uint64_t peer_begin(const uint8_t* local, const PeerLayout& L) {
  const uint64_t seq = *reinterpret_cast<const uint64_t*>(local + L.off_seq) + 1;
  if (threadIdx.x < L.world)
    spin_until(reinterpret_cast<const uint64_t*>(local + L.off_acks) + threadIdx.x, seq - 1, L.spin_budget, "a peer's ack");
  __syncthreads();
  return seq;
}
// Every CTA, after its last store into the peers' regions: the last CTA to get here raises this rank's flag in every
// region (system-scope release, behind a fence that covers all CTAs' stores) and advances the local step number.
// consumed / consumed_L / consumed_seq (optional): a second region whose messages this launch has finished reading;
// the same last CTA acknowledges them to every peer (what a separate peer_ack_kernel would do).
__device__ __forceinline__ void peer_finish(uint8_t* const* bases, const PeerLayout& L, uint64_t seq,
                                            uint8_t* const* consumed = nullptr, const PeerLayout* consumed_L = nullptr,
                                            uint64_t consumed_seq = 0, int* is_last = nullptr) {
  uint8_t* local = bases[L.rank];
  // One system-scope fence per CTA, by the thread that signals: the barrier orders every thread's stores before it and the
  // fence is cumulative (a fence by each of the 256 threads cost several microseconds per launch for the same guarantee).
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned int* done = reinterpret_cast<unsigned int*>(local + L.off_done);
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      *done = 0;
      __threadfence_system();        // every CTA's stores (ordered before their atomicAdd by their own fences) before the flags
      for (int g = 0; g < L.world; ++g) st_relaxed_sys(reinterpret_cast<uint64_t*>(bases[g] + L.off_flags) + L.rank, seq);
      *reinterpret_cast<uint64_t*>(local + L.off_seq) = seq;
      if (is_last) *is_last = 1;     // shared-memory word of the calling CTA
      if (consumed)
        for (int g = 0; g < consumed_L->world; ++g)
          st_relaxed_sys(reinterpret_cast<uint64_t*>(consumed[g] + consumed_L->off_acks) + consumed_L->rank, consumed_seq);
    }
  }
}
// Every CTA of a kernel that only CONSUMES the local region: wait until this step's message of every rank is there.
__device__ __forceinline__ uint64_t peer_arrived(const uint8_t* local, const PeerLayout& L) {
  const uint64_t seq = *reinterpret_cast<const uint64_t*>(local + L.off_seq);
  if (threadIdx.x < L.world)
    spin_until(reinterpret_cast<const uint64_t*>(local + L.off_flags) + threadIdx.x, seq, L.spin_budget, "a peer's message");
  __syncthreads();
  return seq;
}
// ... and, once every CTA has finished reading it, acknowledge the step to every peer (last CTA).
__device__ __forceinline__ void peer_consumed(uint8_t* const* bases, const PeerLayout& L, uint64_t seq) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int* done = reinterpret_cast<unsigned int*>(bases[L.rank] + L.off_done);
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      *done = 0;
      __threadfence_system();        // all CTAs' reads of the region (ordered before their atomicAdd) before the acknowledgements
      for (int g = 0; g < L.world; ++g) st_relaxed_sys(reinterpret_cast<uint64_t*>(bases[g] + L.off_acks) + L.rank, seq);
    }
  }
}

// host side (peer.cu): publish src into every rank's region and wait, in the same launch, for every rank's message
void publish_and_wait(rs_peer* h, const void* src, int64_t bytes, cudaStream_t st);

}  // namespace rsp
