// peer.cu -- the exchange step of the sharded catalog search over NVLink peer memory, without NCCL
// (BASELINE config 3; SURVEY.md section 8e "write partial lists straight into peers' buffers").
//
// One process per GPU.  Every rank owns one cudaMalloc'ed region
//     [ data: world slots of slot_bytes ][ flags: world x u64 ][ acks: world x u64 ][ seq: u64 ][ done: u32 ]
// exported through CUDA IPC; after the 64-byte handles have been exchanged (the caller's one-time all-gather) every
// rank has all regions mapped.  An all-gather of one message per rank is then three tiny launches per rank:
//
//   publish  grid-stride copy of the local message into slot `rank` of EVERY rank's region (128-bit stores, the remote
//            ones travel over NVLink), then -- last CTA, after a system-scope fence -- a release store of the step
//            number into flags[rank] of every region.  Before writing, the kernel waits until every peer has
//            acknowledged the previous step (acks), so one buffer is enough and its address never changes.
//   wait     one warp spins (acquire loads, local memory) until all `world` flags carry this step's number.  Kernels
//            queued behind it -- the merge reads the parts straight out of the region, strided -- see every message.
//   ack      after the consumers: release store of the step number into acks[rank] of every region.
//
// The step number lives in device memory and is advanced by publish itself, so the three launches can be captured in
// a CUDA graph and replayed.  Every spin has a ~10 s budget and traps with a message instead of hanging the device.
#include <cstdio>
#include <cstring>
#include <memory>

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

namespace {

constexpr int kMaxWorld = 16;

struct PeerLayout {
  int64_t slot_bytes;
  int64_t off_flags, off_acks, off_seq, off_done;
  int world, rank;
};

PeerLayout layout_of(const rs_peer* h) {
  PeerLayout L{};
  L.slot_bytes = h->slot_bytes;
  L.world = h->world;
  L.rank = h->rank;
  L.off_flags = h->slot_bytes * h->world;
  L.off_acks = L.off_flags + 8 * kMaxWorld;
  L.off_seq = L.off_acks + 8 * kMaxWorld;
  L.off_done = L.off_seq + 8;
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

// spins until *p >= want; ~10 s budget (ranks may be seconds apart on their first step)
__device__ __forceinline__ void spin_until(const uint64_t* p, uint64_t want, const char* what) {
  for (uint32_t it = 0; ld_acquire_sys(p) < want; ++it) {
    __nanosleep(64);
    if (it > (1u << 23)) {
      printf("recsys_b200: peer exchange timed out waiting for %s (have %llu, want %llu)\n", what,
             (unsigned long long)ld_acquire_sys(p), (unsigned long long)want);
      __trap();
    }
  }
}

__global__ void __launch_bounds__(256) peer_publish_kernel(uint8_t* const* __restrict__ bases, const PeerLayout L,
                                                            const uint8_t* __restrict__ src, int64_t bytes) {
  uint8_t* local = bases[L.rank];
  const uint64_t seq = *reinterpret_cast<const uint64_t*>(local + L.off_seq) + 1;
  // every peer must have consumed the previous message before its slot is overwritten
  if (threadIdx.x < L.world)
    spin_until(reinterpret_cast<const uint64_t*>(local + L.off_acks) + threadIdx.x, seq - 1, "a peer's ack");
  __syncthreads();
  const int64_t n16 = bytes >> 4;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = s4[i];
    for (int g = 0; g < L.world; ++g) reinterpret_cast<uint4*>(bases[g] + L.rank * L.slot_bytes)[i] = v;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n16 << 4) + threadIdx.x; i < bytes; i += blockDim.x)
      for (int g = 0; g < L.world; ++g) (bases[g] + L.rank * L.slot_bytes)[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* done = reinterpret_cast<unsigned int*>(local + L.off_done);
    if (atomicAdd(done, 1u) == gridDim.x - 1) {   // every CTA's stores are ordered before this point
      *done = 0;
      __threadfence_system();
      for (int g = 0; g < L.world; ++g) st_release_sys(reinterpret_cast<uint64_t*>(bases[g] + L.off_flags) + L.rank, seq);
      *reinterpret_cast<uint64_t*>(local + L.off_seq) = seq;
    }
  }
}

__global__ void peer_wait_kernel(const uint8_t* __restrict__ local, const PeerLayout L) {
  const uint64_t seq = *reinterpret_cast<const uint64_t*>(local + L.off_seq);
  if (threadIdx.x < L.world)
    spin_until(reinterpret_cast<const uint64_t*>(local + L.off_flags) + threadIdx.x, seq, "a peer's message");
}

__global__ void peer_ack_kernel(uint8_t* const* __restrict__ bases, const PeerLayout L) {
  const uint64_t seq = *reinterpret_cast<const uint64_t*>(bases[L.rank] + L.off_seq);
  if (threadIdx.x < L.world) st_release_sys(reinterpret_cast<uint64_t*>(bases[threadIdx.x] + L.off_acks) + L.rank, seq);
}

}  // namespace

extern "C" {

int rs_peer_create(int rank, int world, int64_t slot_bytes, rs_peer_t** out, void* handle_out) {
  return rs::guard([&] {
    RS_CHECK(out != nullptr && handle_out != nullptr, "out / handle_out is NULL");
    RS_CHECK(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "rank / world out of range (world <= 16)");
    RS_CHECK(slot_bytes > 0 && slot_bytes % 16 == 0, "slot_bytes must be a positive multiple of 16");
    auto h = std::unique_ptr<rs_peer>(new rs_peer());
    h->rank = rank;
    h->world = world;
    h->slot_bytes = slot_bytes;
    RS_CUDA(cudaGetDevice(&h->device));
    const PeerLayout L = layout_of(h.get());
    h->total = static_cast<size_t>(L.off_done + 16);
    RS_CUDA(cudaMalloc(&h->local, h->total));
    RS_CUDA(cudaMemset(h->local, 0, h->total));
    RS_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t ipc;
    RS_CUDA(cudaIpcGetMemHandle(&ipc, h->local));
    static_assert(sizeof(ipc) == 64, "CUDA IPC handles are 64 bytes");
    std::memcpy(handle_out, &ipc, sizeof(ipc));
    h->bases[rank] = h->local;
    *out = h.release();
  });
}

int rs_peer_connect(rs_peer_t* h, const void* handles) {
  return rs::guard([&] {
    RS_CHECK(h != nullptr && handles != nullptr, "peer / handles is NULL");
    RS_CHECK(!h->connected, "peer region is already connected");
    for (int g = 0; g < h->world; ++g) {
      if (g == h->rank) continue;
      cudaIpcMemHandle_t ipc;
      std::memcpy(&ipc, static_cast<const uint8_t*>(handles) + 64 * g, sizeof(ipc));
      RS_CUDA(cudaIpcOpenMemHandle(&h->bases[g], ipc, cudaIpcMemLazyEnablePeerAccess));
      h->opened[g] = true;
    }
    RS_CUDA(cudaMalloc(&h->d_bases, sizeof(uint8_t*) * kMaxWorld));
    RS_CUDA(cudaMemcpy(h->d_bases, h->bases, sizeof(uint8_t*) * kMaxWorld, cudaMemcpyHostToDevice));
    h->connected = true;
  });
}

int rs_peer_local(rs_peer_t* h, void** data, int64_t* bytes) {
  return rs::guard([&] {
    RS_CHECK(h != nullptr && data != nullptr && bytes != nullptr, "NULL argument");
    *data = h->local;
    *bytes = h->slot_bytes * h->world;
  });
}

int rs_peer_publish(rs_peer_t* h, const DLTensor* src, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr && h->connected, "peer region is not connected");
    require(src, "src", 1, kDLUInt, 8);
    RS_CHECK(on_device(src) && (src->shape[0] <= 1 || stride_of(src, 0) == 1), "src must be a contiguous CUDA uint8 tensor");
    const int64_t bytes = src->shape[0];
    RS_CHECK(bytes >= 0 && bytes <= h->slot_bytes, "message longer than the slot");
    require_aligned16(data_ptr<uint8_t>(src), "src");
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((bytes / 16 + 255) / 256, sm_count())));
    peer_publish_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(h->d_bases, layout_of(h), data_ptr<uint8_t>(src), bytes);
    RS_LAUNCH_CHECK();
  });
}

int rs_peer_wait(rs_peer_t* h, void* stream) {
  return rs::guard([&] {
    RS_CHECK(h != nullptr && h->connected, "peer region is not connected");
    peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(h->local, layout_of(h));
    RS_LAUNCH_CHECK();
  });
}

int rs_peer_ack(rs_peer_t* h, void* stream) {
  return rs::guard([&] {
    RS_CHECK(h != nullptr && h->connected, "peer region is not connected");
    peer_ack_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(h->d_bases, layout_of(h));
    RS_LAUNCH_CHECK();
  });
}

int rs_peer_free(rs_peer_t* h) {
  return rs::guard([&] {
    if (!h) return;
    cudaDeviceSynchronize();
    delete h;
  });
}

}  // extern "C"
