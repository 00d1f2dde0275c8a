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
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "common.cuh"

#include "peer.cuh"

namespace {

using namespace rsp;


// wait_arrival: the last CTA, after raising this rank's flags, also waits until every rank's message of this step has
// arrived here -- ONE warp spins, and the consumer kernels queued behind (1,024 CTAs each in the sharded search) find their
// data present.  Folding the wait into the consumers instead was measured slower on 8 GPUs (0.483 vs 0.453 ms per step): a
// thousand CTAs polling system-scope flags hammer the very L2 lines the peers' flag stores have to land in.  No circular
// wait: a rank's publish depends only on its own earlier kernels and on acknowledgements its peers' consumers sent a step ago.
__global__ void __launch_bounds__(256) peer_publish_kernel(uint8_t* const* __restrict__ bases, const PeerLayout L,
                                                            const uint8_t* __restrict__ src, int64_t bytes, int wait_arrival) {
  const uint64_t seq = peer_begin(bases[L.rank], L);
  const int64_t n16 = bytes >> 4;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = s4[i];
    for (int g = 0; g < L.world; ++g) reinterpret_cast<uint4*>(bases[g] + L.rank * L.slot_bytes)[i] = v;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n16 << 4) + threadIdx.x; i < bytes; i += blockDim.x)
      for (int g = 0; g < L.world; ++g) (bases[g] + L.rank * L.slot_bytes)[i] = src[i];
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = 0;
  __syncthreads();
  peer_finish(bases, L, seq, nullptr, nullptr, 0, wait_arrival ? &s_last : nullptr);
  if (wait_arrival) {
    __syncthreads();
    if (s_last && threadIdx.x < L.world)
      spin_until(reinterpret_cast<const uint64_t*>(bases[L.rank] + L.off_flags) + threadIdx.x, seq, L.spin_budget, "a peer's message");
  }
}

__global__ void peer_wait_kernel(const uint8_t* __restrict__ local, const PeerLayout L) {
  const uint64_t seq = *reinterpret_cast<const uint64_t*>(local + L.off_seq);
  if (threadIdx.x < L.world)
    spin_until(reinterpret_cast<const uint64_t*>(local + L.off_flags) + threadIdx.x, seq, L.spin_budget, "a peer's message");
}

// One warp that waits for what a machine-filling kernel queued behind it would otherwise spin on at entry.  The lookup's
// rows / restore kernels launch up to 8 CTAs per SM; with several tables looked up concurrently (side streams, parallel
// graph branches) such grids could fill every SM of rank A with restore(user) waiting for B's rows(user) while B is full
// of restore(item) waiting for A's rows(item) -- neither rows kernel would ever get an SM.  With the wait in a single
// warp ahead of them, the big grids only start once their data has arrived and never spin.
//   arrived != nullptr: this step's message of every rank is in region `arrived` (peer_arrived's condition)
//   begin   != nullptr: every peer has acknowledged the previous step of region `begin` (peer_begin's condition)
__global__ void peer_prewait_kernel(const uint8_t* __restrict__ arrived, const PeerLayout LA, const uint8_t* __restrict__ begin,
                                    const PeerLayout LB) {
  if (arrived) {
    const uint64_t seq = *reinterpret_cast<const uint64_t*>(arrived + LA.off_seq);
    if (threadIdx.x < LA.world)
      spin_until(reinterpret_cast<const uint64_t*>(arrived + LA.off_flags) + threadIdx.x, seq, LA.spin_budget, "a peer's message");
  }
  if (begin) {
    const uint64_t seq = *reinterpret_cast<const uint64_t*>(begin + LB.off_seq);
    if (threadIdx.x < LB.world)
      spin_until(reinterpret_cast<const uint64_t*>(begin + LB.off_acks) + threadIdx.x, seq, LB.spin_budget, "a peer's ack");
  }
}

__global__ void peer_ack_kernel(uint8_t* const* __restrict__ bases, const PeerLayout L) {
  const uint64_t seq = *reinterpret_cast<const uint64_t*>(bases[L.rank] + L.off_seq);
  if (threadIdx.x < L.world) st_release_sys(reinterpret_cast<uint64_t*>(bases[threadIdx.x] + L.off_acks) + L.rank, seq);
}

// ------------------------------------------------------------------------------------------------
// Sharded embedding lookups (BASELINE config 4) over the same regions: both all-to-alls of a lookup are stores into
// peer memory with DEVICE-known counts -- no NCCL, no host synchronisation, exact byte counts, and the owner-side
// row gather is fused with the way back (the gathered rows are written straight into the requester's region).
//   plan      (lookup_* kernels) optional de-duplication through an open-addressing hash table, per-owner lists of
//             local row numbers (row div G), inv[i] = where id i's row will land in this rank's rows region
//   ids out   a2a_ids_kernel: list for owner g -> slot `rank` of g's ids region, [count | pad | rows ...]
//   rows back a2a_rows_kernel (on the owner): waits for every requester's list, one warp per requested row,
//             table row -> slot `rank` of the requester's rows region; acknowledges the ids messages
//   restore   a2a_restore_kernel: waits for every owner's rows, copies them into batch order out of the local rows
//             region (in de-duplication mode the position comes from the hash table: no inverse pass), acknowledges
// Seven stream operations per lookup (two memsets, insert, assign, ids, rows, restore); the waits and acknowledgements
// ride inside the kernels that need them.
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kIdsHeader = 16;   // bytes in front of the row numbers of an ids message: int64 count + padding

__device__ __forceinline__ uint32_t hash_u64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return static_cast<uint32_t>(x);
}
__device__ __forceinline__ int64_t clamp_id(int64_t id) { return id < 0 ? 0 : id; }   // as rs_a2a_scatter does

// pass 1: slot_of[i] = hash-table slot of ids[i]; bit 31 set for the one thread that inserted the key
__global__ void __launch_bounds__(256) lookup_insert_kernel(const int64_t* __restrict__ ids, int64_t B,
                                                            unsigned long long* __restrict__ keys, uint32_t mask,
                                                            uint32_t* __restrict__ slot_of) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long id = static_cast<unsigned long long>(clamp_id(ids[i]));
    uint32_t h = hash_u64(id) & mask;
    for (;;) {
      const unsigned long long old = atomicCAS(&keys[h], kEmptyKey, id);
      if (old == kEmptyKey) { slot_of[i] = h | 0x80000000u; break; }
      if (old == id) { slot_of[i] = h; break; }
      h = (h + 1) & mask;
    }
  }
}

// pass 2: every first occurrence (all ids when slot_of == nullptr: no de-duplication) takes the next position of its
// owner's list; block-aggregated so that a block costs one global atomic per owner
__global__ void __launch_bounds__(256) lookup_assign_kernel(const int64_t* __restrict__ ids, int64_t B, int G, int64_t b_max,
                                                            const uint32_t* __restrict__ slot_of, int32_t* __restrict__ vals,
                                                            unsigned long long* __restrict__ counts,
                                                            int64_t* __restrict__ stage, int64_t* __restrict__ inv) {
  __shared__ unsigned int local[kMaxWorld];
  __shared__ unsigned long long start[kMaxWorld];
  if (threadIdx.x < kMaxWorld) local[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t c0 = blockIdx.x * (int64_t)blockDim.x; c0 < B; c0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = c0 + threadIdx.x;
    int64_t id = 0;
    int g = -1;
    unsigned int slot = 0;
    if (i < B && (!slot_of || (slot_of[i] & 0x80000000u))) {
      id = clamp_id(ids[i]);
      g = static_cast<int>(id % G);
      slot = atomicAdd(&local[g], 1u);
    }
    __syncthreads();
    if (threadIdx.x < G) start[threadIdx.x] = local[threadIdx.x] ? atomicAdd(&counts[threadIdx.x], (unsigned long long)local[threadIdx.x]) : 0ull;
    __syncthreads();
    if (g >= 0) {
      const int64_t pos = static_cast<int64_t>(start[g]) + slot;
      stage[g * b_max + pos] = id / G;
      if (slot_of) vals[slot_of[i] & 0x7FFFFFFFu] = static_cast<int32_t>(pos);
      else inv[i] = g * b_max + pos;
    }
    __syncthreads();
    if (threadIdx.x < kMaxWorld) local[threadIdx.x] = 0;
    __syncthreads();
  }
}

// pass 3 (de-duplication only): every id finds its key's position
__global__ void __launch_bounds__(256) lookup_inverse_kernel(const int64_t* __restrict__ ids, int64_t B, int G, int64_t b_max,
                                                             const uint32_t* __restrict__ slot_of, const int32_t* __restrict__ vals,
                                                             int64_t* __restrict__ inv) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x)
    inv[i] = (clamp_id(ids[i]) % G) * b_max + vals[slot_of[i] & 0x7FFFFFFFu];
}

__global__ void __launch_bounds__(256) a2a_ids_kernel(uint8_t* const* __restrict__ bases, const PeerLayout L,
                                                       const int64_t* __restrict__ stage, int64_t b_max,
                                                       const unsigned long long* __restrict__ counts) {
  const uint64_t seq = peer_begin(bases[L.rank], L);
  const int64_t total = (int64_t)L.world * b_max;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = static_cast<int>(i / b_max);
    const int64_t j = i - g * b_max;
    const int64_t n = static_cast<int64_t>(counts[g]);
    uint8_t* slot = bases[g] + L.rank * L.slot_bytes;
    if (j == 0) *reinterpret_cast<int64_t*>(slot) = n;
    if (j < n) reinterpret_cast<int64_t*>(slot + kIdsHeader)[j] = stage[i];
  }
  peer_finish(bases, L, seq);
}

// on the owner: requester r asked for LA.count(r) rows; row j of that list goes to slot `rank` of r's rows region
__global__ void __launch_bounds__(256) a2a_rows_kernel(uint8_t* const* __restrict__ ids_bases, const PeerLayout LA,
                                                        uint8_t* const* __restrict__ row_bases, const PeerLayout LB,
                                                        const float* __restrict__ table, int64_t n_local, int64_t ld, int dim) {
  const uint8_t* ids_local = ids_bases[LA.rank];
  const uint64_t seq_ids = peer_arrived(ids_local, LA);
  const uint64_t seq = peer_begin(row_bases[LB.rank], LB);
  __shared__ int64_t prefix[kMaxWorld + 1];
  if (threadIdx.x == 0) {
    int64_t acc = 0;
    for (int r = 0; r < LA.world; ++r) {
      prefix[r] = acc;
      acc += *reinterpret_cast<const int64_t*>(ids_local + r * LA.slot_bytes);
    }
    prefix[LA.world] = acc;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int chunks = dim >> 2;
  const int64_t row_bytes = (int64_t)dim * 4;
  // four rows per warp in flight: a row is a dependent chain (list entry -> random 512-byte row -> remote store) of ~1.5 us,
  // and one row at a time per warp left the kernel latency-bound (36 us for 25 k rows)
  constexpr int UR = 4;
  const int64_t n_req = prefix[LA.world];
  for (int64_t t0 = warp0 * UR; t0 < n_req; t0 += nwarps * UR) {
    float4* dst[UR];
    const float4* src[UR];
    bool ok[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int64_t t = t0 + u;
      dst[u] = nullptr;
      src[u] = nullptr;
      ok[u] = false;
      if (t < n_req) {
        int r = 0;
        while (t >= prefix[r + 1]) ++r;
        const int64_t j = t - prefix[r];
        const int64_t lr = reinterpret_cast<const int64_t*>(ids_local + r * LA.slot_bytes + kIdsHeader)[j];
        dst[u] = reinterpret_cast<float4*>(row_bases[r] + LB.rank * LB.slot_bytes + j * row_bytes);
        ok[u] = lr >= 0 && lr < n_local;
        src[u] = reinterpret_cast<const float4*>(table + (ok[u] ? lr : 0) * ld);
      }
    }
    for (int c = lane; c < chunks; c += 32) {
      float4 v[UR];
#pragma unroll
      for (int u = 0; u < UR; ++u) v[u] = (dst[u] && ok[u]) ? __ldg(src[u] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < UR; ++u)
        if (dst[u]) dst[u][c] = v[u];
    }
  }
  peer_finish(row_bases, LB, seq, ids_bases, &LA, seq_ids);
}

// on the requester: out[i] = the row of ids[i], wherever its owner put it in the local rows region
__global__ void __launch_bounds__(256) a2a_restore_kernel(uint8_t* const* __restrict__ row_bases, const PeerLayout LB,
                                                           const int64_t* __restrict__ ids, int64_t B, int G, int64_t b_max,
                                                           const uint32_t* __restrict__ slot_of, const int32_t* __restrict__ vals,
                                                           const int64_t* __restrict__ inv, float* __restrict__ out,
                                                           int64_t ld_out, int dim) {
  const uint8_t* local = row_bases[LB.rank];
  const uint64_t seq = peer_arrived(local, LB);
  const int chunks = dim >> 2;
  const int64_t total = B * chunks;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / chunks;
    const int c = static_cast<int>(t - i * chunks);
    const int64_t pos = slot_of ? (clamp_id(ids[i]) % G) * b_max + vals[slot_of[i] & 0x7FFFFFFFu] : inv[i];
    // plain loads: the region was written by other GPUs during this launch's lifetime (no read-only path)
    const float4 v = reinterpret_cast<const float4*>(local + pos * (int64_t)dim * 4)[c];
    *reinterpret_cast<float4*>(out + i * ld_out + 4 * c) = v;
  }
  peer_consumed(row_bases, LB, seq);
}

}  // namespace

namespace rsp {
// publish + wait for every rank's message of this step (see peer_publish_kernel); src: device bytes, 16-byte aligned
void publish_and_wait(rs_peer* h, const void* src, int64_t bytes, cudaStream_t st) {
  RS_CHECK(h != nullptr && h->connected, "peer region is not connected");
  RS_CHECK(bytes >= 0 && bytes <= h->slot_bytes, "message longer than the slot");
  rs::require_aligned16(src, "src");
  PeerDeviceGuard g(h->device);
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((bytes / 16 + 255) / 256, rs::sm_count())));
  rs::ProfScope prof(RS_PROF_EXCHANGE, st);
  peer_publish_kernel<<<blocks, 256, 0, st>>>(h->d_bases, layout_of(h), static_cast<const uint8_t*>(src), bytes, 1);
  RS_LAUNCH_CHECK();
}
}  // namespace rsp

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
    PeerDeviceGuard g(h->device);
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((bytes / 16 + 255) / 256, sm_count())));
    peer_publish_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(h->d_bases, layout_of(h), data_ptr<uint8_t>(src), bytes, 0);
    RS_LAUNCH_CHECK();
  });
}

int rs_peer_wait(rs_peer_t* h, void* stream) {
  return rs::guard([&] {
    RS_CHECK(h != nullptr && h->connected, "peer region is not connected");
    PeerDeviceGuard g(h->device);
    peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(h->local, layout_of(h));
    RS_LAUNCH_CHECK();
  });
}

int rs_peer_ack(rs_peer_t* h, void* stream) {
  return rs::guard([&] {
    RS_CHECK(h != nullptr && h->connected, "peer region is not connected");
    PeerDeviceGuard g(h->device);
    peer_ack_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(h->d_bases, layout_of(h));
    RS_LAUNCH_CHECK();
  });
}

int rs_a2a_plan_peer(const DLTensor* ids, int n_shards, int dedup, DLTensor* keys, DLTensor* vals, DLTensor* slot_of,
                     DLTensor* counts, DLTensor* stage, DLTensor* inv, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(ids, "ids", 1, kDLInt, 64);
    require(counts, "counts", 1, kDLInt, 64);
    require(stage, "stage", 2, kDLInt, 64);
    const int64_t B = ids->shape[0], G = n_shards, b_max = stage->shape[1];
    RS_CHECK(G >= 1 && G <= kMaxWorld, "n_shards must be in [1, 16]");
    RS_CHECK(inv != nullptr || dedup, "inv may only be NULL with de-duplication");
    if (inv) {
      require(inv, "inv", 1, kDLInt, 64);
      RS_CHECK(on_device(inv) && inv->shape[0] == B && (B <= 1 || stride_of(inv, 0) == 1), "inv must be a contiguous CUDA [B] tensor");
    }
    RS_CHECK(on_device(ids) && on_device(counts) && on_device(stage), "all tensors must be CUDA tensors");
    RS_CHECK(counts->shape[0] == G && stage->shape[0] == G && is_compact(stage) && B <= b_max,
             "counts [G], stage [G, b_max] contiguous, B <= b_max");
    RS_CHECK(B <= 1 || stride_of(ids, 0) == 1, "ids must be contiguous");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RS_CUDA(cudaMemsetAsync(data_ptr<void>(counts), 0, sizeof(int64_t) * G, st));
    if (B == 0) return;
    const int blocks = static_cast<int>(std::min<int64_t>((B + 255) / 256, (int64_t)sm_count() * 8));
    uint32_t* so = nullptr;
    int32_t* vp = nullptr;
    if (dedup) {
      require(keys, "keys", 1, kDLInt, 64);
      require(vals, "vals", 1, kDLInt, 32);
      require(slot_of, "slot_of", 1, kDLInt, 32);
      const int64_t H = keys->shape[0];
      RS_CHECK(on_device(keys) && on_device(vals) && on_device(slot_of) && vals->shape[0] == H && slot_of->shape[0] >= B &&
                   H >= 2 * B && (H & (H - 1)) == 0 && H <= (int64_t(1) << 30),
               "keys / vals [H] with H a power of two >= 2B, slot_of [>= B]");
      RS_CUDA(cudaMemsetAsync(data_ptr<void>(keys), 0xFF, sizeof(int64_t) * H, st));
      so = data_ptr<uint32_t>(slot_of);
      vp = data_ptr<int32_t>(vals);
      lookup_insert_kernel<<<blocks, 256, 0, st>>>(data_ptr<int64_t>(ids), B, data_ptr<unsigned long long>(keys),
                                                   static_cast<uint32_t>(H - 1), so);
      RS_LAUNCH_CHECK();
    }
    lookup_assign_kernel<<<blocks, 256, 0, st>>>(data_ptr<int64_t>(ids), B, static_cast<int>(G), b_max, so, vp,
                                                 data_ptr<unsigned long long>(counts), data_ptr<int64_t>(stage),
                                                 inv ? data_ptr<int64_t>(inv) : nullptr);
    RS_LAUNCH_CHECK();
    if (dedup && inv) {   // callers that restore with rs_peer_a2a_restore pass inv = NULL: it reads the hash table itself
      lookup_inverse_kernel<<<blocks, 256, 0, st>>>(data_ptr<int64_t>(ids), B, static_cast<int>(G), b_max, so, vp, data_ptr<int64_t>(inv));
      RS_LAUNCH_CHECK();
    }
  });
}

int rs_peer_a2a_ids(rs_peer_t* h, const DLTensor* stage, const DLTensor* counts, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr && h->connected, "peer region is not connected");
    require(stage, "stage", 2, kDLInt, 64);
    require(counts, "counts", 1, kDLInt, 64);
    const int64_t b_max = stage->shape[1];
    RS_CHECK(on_device(stage) && on_device(counts) && is_compact(stage) && stage->shape[0] == h->world && counts->shape[0] == h->world,
             "stage must be a contiguous CUDA [world, b_max] tensor, counts [world]");
    RS_CHECK(kIdsHeader + b_max * 8 <= h->slot_bytes, "ids region slots are too small for b_max row numbers");
    const int64_t total = (int64_t)h->world * b_max;
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, sm_count())));
    a2a_ids_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(h->d_bases, layout_of(h), data_ptr<int64_t>(stage), b_max,
                                                                         data_ptr<unsigned long long>(counts));
    RS_LAUNCH_CHECK();
  });
}

int rs_peer_a2a_rows(rs_peer_t* ids_region, rs_peer_t* rows_region, const DLTensor* table, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(ids_region != nullptr && ids_region->connected && rows_region != nullptr && rows_region->connected,
             "peer regions are not connected");
    RS_CHECK(ids_region->world == rows_region->world && ids_region->rank == rows_region->rank, "the two regions belong to different groups");
    require(table, "table", 2, kDLFloat, 32);
    RS_CHECK(on_device(table) && table->shape[1] % 4 == 0, "table must be a CUDA [n_local, dim] tensor, dim a multiple of 4");
    const int64_t ld = row_stride_2d(table, "table");
    RS_CHECK(ld % 4 == 0, "table row stride must be a multiple of 4 floats");
    require_aligned16(data_ptr<float>(table), "table");
    const int64_t b_max = (ids_region->slot_bytes - kIdsHeader) / 8;
    RS_CHECK(b_max * table->shape[1] * 4 <= rows_region->slot_bytes, "rows region slots are too small for b_max rows");
    PeerDeviceGuard g(rows_region->device);
    peer_prewait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(ids_region->local, layout_of(ids_region), rows_region->local,
                                                                          layout_of(rows_region));
    RS_LAUNCH_CHECK();
    a2a_rows_kernel<<<sm_count(), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        ids_region->d_bases, layout_of(ids_region), rows_region->d_bases, layout_of(rows_region), data_ptr<float>(table),
        table->shape[0], ld, static_cast<int>(table->shape[1]));
    RS_LAUNCH_CHECK();
  });
}

int rs_peer_a2a_restore(rs_peer_t* rows_region, const DLTensor* ids, const DLTensor* slot_of, const DLTensor* vals,
                        const DLTensor* inv, DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(rows_region != nullptr && rows_region->connected, "peer region is not connected");
    require(ids, "ids", 1, kDLInt, 64);
    require(out, "out", 2, kDLFloat, 32);
    const int64_t B = ids->shape[0], dim = out->shape[1];
    RS_CHECK(on_device(ids) && on_device(out) && out->shape[0] == B && dim % 4 == 0 && (B <= 1 || stride_of(ids, 0) == 1),
             "ids must be a contiguous CUDA [B] tensor, out a CUDA [B, dim] tensor with dim a multiple of 4");
    const int64_t ld_out = row_stride_2d(out, "out");
    RS_CHECK(ld_out % 4 == 0, "out row stride must be a multiple of 4 floats");
    require_aligned16(data_ptr<float>(out), "out");
    const int64_t b_max = rows_region->slot_bytes / (dim * 4);
    RS_CHECK(b_max * dim * 4 == rows_region->slot_bytes, "rows region slots must hold a whole number of rows of this width");
    const uint32_t* so = nullptr;
    const int32_t* vp = nullptr;
    const int64_t* ip = nullptr;
    if (slot_of) {
      require(slot_of, "slot_of", 1, kDLInt, 32);
      require(vals, "vals", 1, kDLInt, 32);
      RS_CHECK(on_device(slot_of) && on_device(vals) && slot_of->shape[0] >= B, "slot_of [>= B] / vals must be CUDA tensors");
      so = data_ptr<uint32_t>(slot_of);
      vp = data_ptr<int32_t>(vals);
    } else {
      require(inv, "inv", 1, kDLInt, 64);
      RS_CHECK(on_device(inv) && inv->shape[0] == B && (B <= 1 || stride_of(inv, 0) == 1), "inv must be a contiguous CUDA [B] tensor");
      ip = data_ptr<int64_t>(inv);
    }
    const int64_t total = std::max<int64_t>(1, B * (dim / 4));
    const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8));
    PeerDeviceGuard g(rows_region->device);
    peer_prewait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(rows_region->local, layout_of(rows_region), nullptr,
                                                                          layout_of(rows_region));
    RS_LAUNCH_CHECK();
    a2a_restore_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows_region->d_bases, layout_of(rows_region),
                                                                             data_ptr<int64_t>(ids), B, rows_region->world, b_max, so,
                                                                             vp, ip, data_ptr<float>(out), ld_out, static_cast<int>(dim));
    RS_LAUNCH_CHECK();
  });
}

int rs_peer_disconnect(rs_peer_t* h) {
  return rs::guard([&] {
    if (!h) return;
    RS_CUDA(cudaDeviceSynchronize());
    for (int g = 0; g < h->world; ++g)
      if (h->opened[g]) {
        cudaIpcCloseMemHandle(h->bases[g]);
        h->opened[g] = false;
        h->bases[g] = nullptr;
      }
    h->connected = false;
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
