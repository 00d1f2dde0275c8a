// flat_index.cu -- K4/K5/K6/K7 (exact fp32 mode): the FAISS IndexFlatIP replacement.
//
// Replaces faiss.IndexFlatIP.add/search (src/trainer.py:241-242, app/recommendation_service.py:72),
// faiss.normalize_L2 (src/trainer.py:240, app/recommendation_service.py:71) and
// tf.linalg.matvec (app/recommendation_service.py:91).
//
// scan_topk_kernel -- the catalog scan.  HBM-bound: the fp32 catalog ([N, d] row-major, the layout
// faiss keeps) is streamed once per group of QT queries through a 3-stage shared-memory ring filled by
// TMA (cp.async.bulk.tensor, 128-byte swizzle, one mbarrier per stage).  One thread scores one catalog
// row of the tile against the QT resident queries with the stated arithmetic -- a single fp32
// accumulator per (query, row), fused multiply-add, feature index ascending -- which the oracle
// (oracle/chain.c) reproduces bit for bit.  The swizzle makes the row-per-thread 128-bit shared loads
// bank-conflict free.  Scores never go to HBM: each is compared in registers with the CTA's running
// k-th key and only the rare survivors enter the shared-memory candidate buffer (topk.cuh).
// Algorithmic bytes: N*d*4 per query group (+ nq*d*4 queries, + nq*P*k*8 partial lists).
//
// merge_topk_kernel -- one CTA per query merges the P per-CTA lists (or the G per-shard lists after
// the all-gather) in rank-major order with early termination.
#include <algorithm>
#include <cfloat>

#include "flat_index.cuh"
#include "peer.cuh"
#include "topk.cuh"

namespace {

using namespace rsd;

constexpr int TILE_ROWS = 128;  // catalog rows per pipeline stage == compute threads
constexpr int SCAN_NT = 128;
constexpr int STAGES = 3;
constexpr int MERGE_NT = 256;

template <int D, int QT, int CAP>
struct ScanSmem {
  static constexpr int kStageBytes = TILE_ROWS * D * 4;
  static constexpr int kKeysOff = STAGES * kStageBytes;
  static constexpr int kQOff = kKeysOff + QT * CAP * 8;
  static constexpr int kBarOff = kQOff + QT * D * 4;
  static constexpr int kCntOff = kBarOff + STAGES * 8;
  static constexpr int kTotal = kCntOff + QT * 4;
  static constexpr int kDynamic = kTotal + 1024;  // slack for the 1024-byte alignment of the ring
};

template <int D, int QT, int CAP>
__global__ void __launch_bounds__(SCAN_NT, 1)
scan_topk_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ q, int64_t n, int nq, int k,
                 uint64_t* __restrict__ partial) {
  using L = ScanSmem<D, QT, CAP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* keys = reinterpret_cast<uint64_t*>(base + L::kKeysOff);
  float* qs = reinterpret_cast<float*>(base + L::kQOff);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + L::kBarOff);
  int* cnt = reinterpret_cast<int*>(base + L::kCntOff);

  const int tid = threadIdx.x;
  const int qbase = blockIdx.y * QT;
  const int64_t ntiles = (n + TILE_ROWS - 1) / TILE_ROWS;
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (tid == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < QT * D; i += SCAN_NT) {
    const int g = i / D;
    qs[i] = (qbase + g < nq) ? q[(int64_t)(qbase + g) * D + (i - g * D)] : 0.0f;
  }
  for (int i = tid; i < QT * CAP; i += SCAN_NT) keys[i] = kEmptyKey;
  if (tid < QT) cnt[tid] = 0;
  __syncthreads();

  auto issue = [&](int stage, int64_t it) {
    const int64_t tile = blockIdx.x + it * gridDim.x;
    mbar_arrive_expect_tx(&full[stage], L::kStageBytes);
#pragma unroll
    for (int j = 0; j < D / 32; ++j)
      tma_load_2d(base + stage * L::kStageBytes + j * (TILE_ROWS * 128), &tmap, &full[stage], j * 32,
                  static_cast<int>(tile * TILE_ROWS), kEvictFirst);
  };
  if (tid == 0)
    for (int s = 0; s < STAGES && s < my_tiles; ++s) issue(s, s);

  uint64_t thr[QT];
#pragma unroll
  for (int g = 0; g < QT; ++g) thr[g] = kEmptyKey;

  const uint32_t sw = static_cast<uint32_t>(tid & 7) << 4;
  for (int64_t it = 0; it < my_tiles; ++it) {
    const int stage = static_cast<int>(it % STAGES);
    mbar_wait(&full[stage], static_cast<uint32_t>((it / STAGES) & 1));

    // ---- score: row `tid` of the tile against QT queries, one FMA chain each ----
    float acc[QT];
#pragma unroll
    for (int g = 0; g < QT; ++g) acc[g] = 0.0f;
    const uint8_t* srow = base + stage * L::kStageBytes + tid * 128;
#pragma unroll
    for (int j = 0; j < D / 32; ++j) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 xv = *reinterpret_cast<const float4*>(srow + j * (TILE_ROWS * 128) + ((c << 4) ^ sw));
#pragma unroll
        for (int g = 0; g < QT; ++g) {
          const float4 qv = *reinterpret_cast<const float4*>(qs + g * D + j * 32 + c * 4);
          acc[g] = fmaf(qv.x, xv.x, acc[g]);
          acc[g] = fmaf(qv.y, xv.y, acc[g]);
          acc[g] = fmaf(qv.z, xv.z, acc[g]);
          acc[g] = fmaf(qv.w, xv.w, acc[g]);
        }
      }
    }

    // ---- select: threshold-reject in registers, survivors to the candidate buffers ----
    const int64_t row = (blockIdx.x + it * gridDim.x) * TILE_ROWS + tid;
    uint64_t key[QT];
    uint32_t pend = 0;
#pragma unroll
    for (int g = 0; g < QT; ++g) {
      key[g] = pack_key(acc[g], static_cast<uint32_t>(row));
      if (row < n && qbase + g < nq && key[g] > thr[g]) pend |= 1u << g;
    }
    for (;;) {
      bool saw_full = false;
#pragma unroll
      for (int g = 0; g < QT; ++g) {
        if (pend & (1u << g)) {
          const int pos = atomicAdd(&cnt[g], 1);
          if (pos < CAP) {
            keys[g * CAP + pos] = key[g];
            pend &= ~(1u << g);
          }
          saw_full |= pos >= CAP - 1;
        }
      }
      // this barrier also orders every thread's reads of the stage before the refill below
      if (!__syncthreads_or(pend != 0 || saw_full)) break;
#pragma unroll
      for (int g = 0; g < QT; ++g) {
        if (cnt[g] >= CAP) {  // uniform: nobody touches cnt between the barrier and the prune
          thr[g] = prune_buffer<CAP, SCAN_NT>(keys + g * CAP, &cnt[g], k, tid);
          if ((pend & (1u << g)) && !(key[g] > thr[g])) pend &= ~(1u << g);
        }
      }
      __syncthreads();
    }
    if (tid == 0 && it + STAGES < my_tiles) issue(stage, it + STAGES);
  }

  // ---- emit this CTA's sorted list for each query ----
  const int P = gridDim.x;
#pragma unroll 1
  for (int g = 0; g < QT; ++g) {
    if (qbase + g >= nq) break;
    prune_buffer<CAP, SCAN_NT>(keys + g * CAP, &cnt[g], k, tid);
    uint64_t* dst = partial + ((int64_t)(qbase + g) * P + blockIdx.x) * k;
    for (int j = tid; j < k; j += SCAN_NT) dst[j] = keys[g * CAP + j];
  }
}

// ------------------------------------------------------------------------------------------------
// merge of sorted lists.  Source A: packed keys [nq][P][k] (rows local to the index).
// Source B: (score, id) pairs [P][nq][k] (global ids, -1 = padding) after an all-gather.
// ------------------------------------------------------------------------------------------------
// after the merge of a shared-threshold search: a query needs the exact re-search iff a shard overflowed, the merged
// list is short of k, or its k-th score does not clear every shard's bound (tau + eps)
__global__ void verify_topk_kernel(const float* __restrict__ D, int64_t ldD, const int64_t* __restrict__ I, int64_t ldI, int k,
                                   const float* __restrict__ bounds, int64_t ld_b, const int* __restrict__ flags, int64_t ld_f,
                                   int G, int64_t nq, int* __restrict__ need_exact) {
  const int64_t qi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  float b = -INFINITY;
  int f = 0;
  for (int g = 0; g < G; ++g) {
    b = fmaxf(b, bounds[g * ld_b + qi]);
    f |= flags[g * ld_f + qi];
  }
  const bool ok = f == 0 && I[qi * ldI + (k - 1)] >= 0 && D[qi * ldD + (k - 1)] >= b;
  need_exact[qi] = ok ? 0 : 1;
}

struct KeySource {
  const uint64_t* keys;  // A
  const float* Dp;       // B
  const int64_t* Ip;     // B
  int64_t nq;
  int64_t part_ld_D, part_ld_I;   // B: elements between consecutive parts (nq * k when the parts are packed)
  __device__ __forceinline__ uint64_t load(int64_t qi, int p, int P, int r, int k) const {
    if (keys) return keys[((int64_t)qi * P + p) * k + r];
    const int64_t o = qi * k + r;
    const int64_t id = Ip[p * part_ld_I + o];
    return id < 0 ? kEmptyKey : pack_key(Dp[p * part_ld_D + o], static_cast<uint32_t>(id));
  }
};

// What the sharded search (sharded.cu) folds into the merge: the wait for / acknowledgement of the peer-memory region the
// parts live in, and the per-query verdict that rs_topk_verify computes (OR of the shards' flags; with bounds also the
// shared-threshold verification).  All pointers null: a plain merge.
struct MergeExtras {
  uint8_t* const* peer_bases = nullptr;
  rsp::PeerLayout PL{};
  const float* bounds = nullptr;
  int64_t ld_b = 0;
  const int* flags = nullptr;
  int64_t ld_f = 0;
  int* need_exact = nullptr;
};

template <int CAP>
__global__ void __launch_bounds__(MERGE_NT)
merge_topk_kernel(const KeySource src, int P, int k, int64_t id_base, float* __restrict__ D, int64_t ldD,
                  int64_t* __restrict__ I, int64_t ldI, const MergeExtras ex) {
  __shared__ uint64_t keys[CAP];
  __shared__ int cnt;
  const int tid = threadIdx.x;
  const int64_t qi = blockIdx.x;
  uint64_t seq = 0;
  if (ex.peer_bases) seq = rsp::peer_arrived(ex.peer_bases[ex.PL.rank], ex.PL);
  for (int i = tid; i < CAP; i += MERGE_NT) keys[i] = kEmptyKey;
  if (tid == 0) cnt = 0;
  __syncthreads();
  uint64_t thr = kEmptyKey;
  const int64_t total = (int64_t)P * k;
  int quiet = 0;  // consecutive elements (rank-major) in which no list offered a key
  for (int64_t e0 = 0; e0 < total; e0 += MERGE_NT) {
    const int64_t e = e0 + tid;
    uint64_t key = kEmptyKey;
    if (e < total) {
      const int r = static_cast<int>(e / P);
      const int p = static_cast<int>(e - (int64_t)r * P);
      key = src.load(qi, p, P, r, k);
    }
    const bool any = offer_one<CAP, MERGE_NT>(keys, &cnt, k, key != kEmptyKey, key, thr, tid);
    quiet = any ? 0 : quiet + MERGE_NT;
    if (quiet >= P) break;  // every list's next entry already lost; lists are sorted, so we are done
  }
  prune_buffer<CAP, MERGE_NT>(keys, &cnt, k, tid);
  for (int j = tid; j < k; j += MERGE_NT) {
    const uint64_t key = keys[j];
    if (key == kEmptyKey) {
      D[qi * ldD + j] = -FLT_MAX;
      I[qi * ldI + j] = -1;
    } else {
      D[qi * ldD + j] = key_score(key);
      I[qi * ldI + j] = id_base + static_cast<int64_t>(key_row(key));
    }
  }
  if (ex.need_exact && tid == 0) {   // verify_topk_kernel's verdict for this query
    float b = -INFINITY;
    int f = 0;
    for (int g = 0; g < P; ++g) {
      if (ex.bounds) b = fmaxf(b, ex.bounds[g * ex.ld_b + qi]);
      f |= ex.flags[g * ex.ld_f + qi];
    }
    const uint64_t kth = keys[k - 1];
    const bool ok = f == 0 && (!ex.bounds || (kth != kEmptyKey && key_score(kth) >= b));
    ex.need_exact[qi] = ok ? 0 : 1;
  }
  if (ex.peer_bases) rsp::peer_consumed(ex.peer_bases, ex.PL, seq);
}

__global__ void fill_pad_kernel(float* D, int64_t ldD, int64_t* I, int64_t ldI, int64_t nq, int k) {
  const int64_t total = nq * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k;
    const int j = static_cast<int>(i - r * k);
    D[r * ldD + j] = -FLT_MAX;
    I[r * ldI + j] = -1;
  }
}

// K4: one thread per row, the stated chain over x_j^2 then x_j *= 1/sqrt(nr); zero rows untouched
__global__ void __launch_bounds__(128) normalize_l2_kernel(float* __restrict__ x, int64_t ld, int64_t n, int d) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  float* row = x + r * ld;
  float nr = 0.0f;
  for (int j = 0; j < d; ++j) nr = fmaf(row[j], row[j], nr);
  if (nr > 0.0f) {
    const float inv = 1.0f / sqrtf(nr);
    for (int j = 0; j < d; ++j) row[j] *= inv;
  }
}

// K6: one thread per item row, the stated chain (n <= 100 on the /score path)
__global__ void __launch_bounds__(128) matvec_kernel(const float* __restrict__ items, int64_t ld,
                                                     const float* __restrict__ user, float* __restrict__ out,
                                                     int64_t n, int d) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float* row = items + r * ld;
  float acc = 0.0f;
  for (int j = 0; j < d; ++j) acc = fmaf(user[j], row[j], acc);
  out[r] = acc;
}

template <int CAP>
void launch_merge(const KeySource& src, int64_t nq, int P, int k, int64_t id_base, float* D, int64_t ldD, int64_t* I,
                  int64_t ldI, cudaStream_t st, const MergeExtras& ex) {
  rs::ProfScope prof(RS_PROF_MERGE, st);
  merge_topk_kernel<CAP><<<static_cast<unsigned>(nq), MERGE_NT, 0, st>>>(src, P, k, id_base, D, ldD, I, ldI, ex);
  RS_LAUNCH_CHECK();
}

void merge_dispatch(const KeySource& src, int64_t nq, int P, int k, int64_t id_base, float* D, int64_t ldD, int64_t* I,
                    int64_t ldI, cudaStream_t st, const MergeExtras& ex = MergeExtras{}) {
  RS_CHECK(k >= 1 && k <= 1024, "k must be in [1, 1024]");
  RS_CHECK(nq <= 0x7FFFFFFF, "too many queries");
  if (k <= 128) launch_merge<256>(src, nq, P, k, id_base, D, ldD, I, ldI, st, ex);
  else if (k <= 256) launch_merge<512>(src, nq, P, k, id_base, D, ldD, I, ldI, st, ex);
  else if (k <= 512) launch_merge<1024>(src, nq, P, k, id_base, D, ldD, I, ldI, st, ex);
  else launch_merge<2048>(src, nq, P, k, id_base, D, ldD, I, ldI, st, ex);
}

template <int D, int QT, int CAP>
void launch_scan(const rs_index* h, const float* q, int64_t nq, int k, int P, uint64_t* partial, cudaStream_t st) {
  using L = ScanSmem<D, QT, CAP>;
  static std::once_flag once[64];
  std::call_once(once[h->device & 63], [] {
    cudaFuncSetAttribute(scan_topk_kernel<D, QT, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
  });
  const int64_t groups = (nq + QT - 1) / QT;
  RS_CHECK(groups <= 65535, "too many queries for one search call (limit 524,280)");
  dim3 grid(static_cast<unsigned>(P), static_cast<unsigned>(groups));
  rs::ProfScope prof(RS_PROF_SCAN, st);
  scan_topk_kernel<D, QT, CAP><<<grid, SCAN_NT, L::kDynamic, st>>>(h->tmap_f32, q, h->ntotal, static_cast<int>(nq), k,
                                                                  partial);
  RS_LAUNCH_CHECK();
}

template <int D>
void scan_dispatch(const rs_index* h, const float* q, int64_t nq, int k, int P, uint64_t* partial, cudaStream_t st) {
  if (k > 128) return launch_scan<D, 1, 2048>(h, q, nq, k, P, partial, st);
  if (nq >= 8) return launch_scan<D, 8, 256>(h, q, nq, k, P, partial, st);
  if (nq > 2) return launch_scan<D, 4, 256>(h, q, nq, k, P, partial, st);
  if (nq == 2) return launch_scan<D, 2, 256>(h, q, nq, k, P, partial, st);
  return launch_scan<D, 1, 256>(h, q, nq, k, P, partial, st);
}

void rebuild_tmaps(rs_index* h) {
  if (h->ntotal == 0) return;
  h->tmap_f32 = rs::make_tmap_2d(h->xb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)h->ntotal, (uint64_t)h->d,
                                 (uint64_t)h->d * 4, TILE_ROWS, 32);
}

}  // namespace

namespace rs {

void merge_lists_device(const float* Dparts, int64_t part_ld_D, const int64_t* Iparts, int64_t part_ld_I, int P, int64_t nq,
                        int k, float* D, int64_t ldD, int64_t* I, int64_t ldI, rs_peer* lists_region, const float* bounds,
                        int64_t ld_b, const int* flags, int64_t ld_f, int* need_exact, cudaStream_t st) {
  RS_CHECK(P >= 1 && P <= 4096, "nparts must be in [1, 4096]");
  RS_CHECK(need_exact == nullptr || flags != nullptr, "need_exact needs the shards' flags");
  if (nq == 0) return;
  MergeExtras ex;
  if (lists_region) {
    ex.peer_bases = lists_region->d_bases;
    ex.PL = rsp::layout_of(lists_region);
  }
  ex.bounds = bounds;
  ex.ld_b = ld_b;
  ex.flags = flags;
  ex.ld_f = ld_f;
  ex.need_exact = need_exact;
  KeySource src{nullptr, Dparts, Iparts, nq, part_ld_D, part_ld_I};
  merge_dispatch(src, nq, P, k, 0, D, ldD, I, ldI, st, ex);
}

void search_exact_device(rs_index* h, const float* q, int64_t nq, int k, int64_t id_base, float* D, int64_t ldD,
                         int64_t* I, int64_t ldI, cudaStream_t st) {
  RS_CHECK(k >= 1 && k <= 1024, "k must be in [1, 1024]");
  if (nq == 0) return;
  if (h->ntotal == 0) {
    fill_pad_kernel<<<64, 256, 0, st>>>(D, ldD, I, ldI, nq, k);
    RS_LAUNCH_CHECK();
    return;
  }
  RS_CHECK(h->ntotal <= 0xFFFFFFFFll, "catalog shards above 2^32 rows are not supported");
  const int64_t ntiles = (h->ntotal + TILE_ROWS - 1) / TILE_ROWS;
  int P = static_cast<int>(std::min<int64_t>(ntiles, sm_count()));
  if (ntiles < sm_count()) {
    // Small catalog, many queries (BASELINE config 1: 4,096 queries x 3,706 rows): one CTA per (tile, query group) is
    // 14,848 CTAs that each pay the pipeline start-up for a single 128-row tile.  Give every query group just enough
    // row slices to fill the machine about twice; a CTA then walks several tiles and fewer lists need merging.
    const int qt = k > 128 ? 1 : nq >= 8 ? 8 : nq > 2 ? 4 : static_cast<int>(nq);   // as scan_dispatch picks it
    const int64_t groups = (nq + qt - 1) / qt;
    P = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(P, (2 * (int64_t)sm_count() + groups - 1) / groups)));
  }
  uint64_t* partial = nullptr;
  RS_CUDA(cudaMallocAsync(&partial, sizeof(uint64_t) * (size_t)nq * P * k, st));
  try {
    switch (h->d) {
      case 32: scan_dispatch<32>(h, q, nq, k, P, partial, st); break;
      case 64: scan_dispatch<64>(h, q, nq, k, P, partial, st); break;
      case 128: scan_dispatch<128>(h, q, nq, k, P, partial, st); break;
      default: RS_CHECK(false, "index dimension must be 32, 64 or 128 (got " + std::to_string(h->d) + ")");
    }
    KeySource src{partial, nullptr, nullptr, nq};
    merge_dispatch(src, nq, P, k, id_base, D, ldD, I, ldI, st);
  } catch (...) {
    cudaFreeAsync(partial, st);
    throw;
  }
  RS_CUDA(cudaFreeAsync(partial, st));
}

}  // namespace rs

extern "C" {

int rs_index_create(int d, rs_index_t** out) {
  return rs::guard([&] {
    RS_CHECK(out != nullptr, "out is NULL");
    RS_CHECK(d == 32 || d == 64 || d == 128, "index dimension must be 32, 64 or 128");
    rs::sm_count();  // fails loudly when there is no usable sm_100 device: no CPU fallback
    auto* h = new rs_index();
    h->d = d;
    RS_CUDA(cudaGetDevice(&h->device));
    *out = h;
  });
}

int rs_index_free(rs_index_t* h) {
  return rs::guard([&] {
    if (!h) return;
    {
      rs::DeviceGuard g(h->device);
      cudaFree(h->xb);
      cudaFree(h->xb_bf16);
      cudaFree(h->row_norm);
    }
    delete h;
  });
}

int rs_index_ntotal(const rs_index_t* h, int64_t* out) {
  return rs::guard([&] {
    RS_CHECK(h && out, "NULL argument");
    *out = h->ntotal;
  });
}

int rs_index_dim(const rs_index_t* h, int* out) {
  return rs::guard([&] {
    RS_CHECK(h && out, "NULL argument");
    *out = h->d;
  });
}

int rs_index_reconstruct(const rs_index_t* h, int64_t i0, DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "index is NULL");
    require(out, "out", 2, kDLFloat, 32);
    RS_CHECK(out->shape[1] == h->d && is_compact(out), "out must be a contiguous [n, d] tensor");
    const int64_t n = out->shape[0];
    RS_CHECK(i0 >= 0 && n >= 0 && i0 + n <= h->ntotal, "row range out of bounds");
    RS_CHECK(on_device(out) || on_host(out), "out must be a CUDA or host tensor");
    if (n == 0) return;
    DeviceGuard g(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RS_CUDA(cudaMemcpyAsync(data_ptr<float>(out), h->xb + i0 * h->d, sizeof(float) * (size_t)n * h->d,
                            on_device(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (on_host(out)) RS_CUDA(cudaStreamSynchronize(st));
  });
}

int rs_index_fallbacks(const rs_index_t* h, int64_t* last, int64_t* total) {
  return rs::guard([&] {
    RS_CHECK(h && last && total, "NULL argument");
    *last = h->last_fallbacks.load(std::memory_order_relaxed);
    *total = h->total_fallbacks.load(std::memory_order_relaxed);
  });
}

int rs_debug_tensor_scores(rs_index_t* h, const DLTensor* q, DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "index is NULL");
    require(q, "q", 2, kDLFloat, 32);
    require(out, "out", 2, kDLFloat, 32);
    RS_CHECK(on_device(q) && on_device(out) && is_compact(q) && is_compact(out), "q and out must be contiguous CUDA tensors");
    RS_CHECK(q->shape[1] == h->d && out->shape[0] == q->shape[0] && out->shape[1] == h->ntotal, "shape mismatch");
    DeviceGuard g(h->device);
    tensor_scores_debug(h, data_ptr<float>(q), q->shape[0], data_ptr<float>(out), static_cast<cudaStream_t>(stream));
  });
}

int rs_index_add(rs_index_t* h, const DLTensor* x, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "index is NULL");
    require(x, "x", 2, kDLFloat, 32);
    RS_CHECK(x->shape[1] == h->d, "x must be [n, d] with d == index dimension");
    RS_CHECK(on_device(x) || on_host(x), "x must be a CUDA or host tensor");
    const int64_t n = x->shape[0];
    if (n == 0) return;
    const int64_t ldx = row_stride_2d(x, "x");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t need = h->ntotal + n;
    RS_CHECK(need <= 0xFFFFFFFFll, "catalog shards above 2^32 rows are not supported");
    if (need > h->capacity) {
      int64_t cap = h->capacity == 0 ? need : std::max(need, h->capacity + h->capacity / 2);
      float* nb = nullptr;
      RS_CUDA(cudaMalloc(&nb, sizeof(float) * (size_t)cap * h->d));
      if (h->ntotal > 0)
        RS_CUDA(cudaMemcpyAsync(nb, h->xb, sizeof(float) * (size_t)h->ntotal * h->d, cudaMemcpyDeviceToDevice, st));
      RS_CUDA(cudaStreamSynchronize(st));
      cudaFree(h->xb);
      h->xb = nb;
      h->capacity = cap;
      cudaFree(h->xb_bf16);
      cudaFree(h->row_norm);
      h->xb_bf16 = nullptr;
      h->row_norm = nullptr;
      h->shadow_rows = 0;
    }
    RS_CUDA(cudaMemcpy2DAsync(h->xb + h->ntotal * h->d, sizeof(float) * h->d, data_ptr<float>(x), sizeof(float) * ldx,
                              sizeof(float) * h->d, (size_t)n, on_device(x) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                              st));
    if (on_host(x)) RS_CUDA(cudaStreamSynchronize(st));
    h->ntotal = need;
    h->small_tensor_off.store(false, std::memory_order_relaxed);
    rebuild_tmaps(h);
  });
}

int rs_index_search(rs_index_t* h, const DLTensor* q, int k, int64_t id_base, int mode, DLTensor* D, DLTensor* I,
                    void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "index is NULL");
    require(q, "q", 2, kDLFloat, 32);
    require(D, "D", 2, kDLFloat, 32);
    require(I, "I", 2, kDLInt, 64);
    RS_CHECK(q->shape[1] == h->d, "q must be [nq, d] with d == index dimension");
    RS_CHECK(k >= 1 && k <= 1024, "k must be in [1, 1024]");
    const int64_t nq = q->shape[0];
    RS_CHECK(D->shape[0] == nq && D->shape[1] == k && I->shape[0] == nq && I->shape[1] == k, "D and I must be [nq, k]");
    RS_CHECK(mode == RS_EXACT_FP32 || mode == RS_TENSOR_BF16 || mode == RS_AUTO, "unknown mode");
    const bool dev = on_device(q) && on_device(D) && on_device(I);
    const bool host = on_host(q) && on_host(D) && on_host(I);
    RS_CHECK(dev || host, "q, D, I must be all CUDA tensors or all host tensors");
    if (nq == 0) return;
    DeviceGuard g(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bool tensor = false;
    if (mode == RS_TENSOR_BF16) {
      RS_CHECK(tensor_path_supported(h, nq, k), "tensor mode is not applicable to this shape (needs d=128, k<=128 and ntotal>=65536, or 256<=ntotal<=32768 with k<=ntotal/128)");
      tensor = true;
    } else if (mode == RS_AUTO) {
      // large catalogs: faster than the fp32 scan at every batch size (bf16 shadow = half the bytes); small (L2-resident)
      // ones: from 128 queries -- below that the scan's two launches beat the tensor path's seven (measured, config 1 shape)
      tensor = tensor_path_supported(h, nq, k) &&
               (h->ntotal >= 65536 || (nq >= 128 && !h->small_tensor_off.load(std::memory_order_relaxed)));
    }
    auto run = [&](const float* qp, float* Dp, int64_t ldD, int64_t* Ip, int64_t ldI) {
      if (tensor) {
        search_tensor_device(h, qp, nq, k, id_base, Dp, ldD, Ip, ldI, st);
        if (mode == RS_AUTO && h->ntotal < 65536 && 2 * h->last_fallbacks.load(std::memory_order_relaxed) >= nq)
          h->small_tensor_off.store(true, std::memory_order_relaxed);
      } else {
        search_exact_device(h, qp, nq, k, id_base, Dp, ldD, Ip, ldI, st);
      }
    };
    if (dev) {
      RS_CHECK(is_compact(q), "q must be contiguous");
      RS_CHECK(q->device.device_id == h->device, "q is on a different device than the index");
      require_aligned16(data_ptr<float>(q), "q");
      run(data_ptr<float>(q), data_ptr<float>(D), row_stride_2d(D, "D"), data_ptr<int64_t>(I), row_stride_2d(I, "I"));
      return;
    }
    // host tensors: stage through the device on `st` and return once the results are back
    RS_CHECK(is_compact(q) && is_compact(D) && is_compact(I), "host q, D, I must be contiguous");
    float* dq = nullptr;
    float* dD = nullptr;
    int64_t* dI = nullptr;
    RS_CUDA(cudaMallocAsync(&dq, sizeof(float) * (size_t)nq * h->d, st));
    RS_CUDA(cudaMallocAsync(&dD, sizeof(float) * (size_t)nq * k, st));
    RS_CUDA(cudaMallocAsync(&dI, sizeof(int64_t) * (size_t)nq * k, st));
    try {
      RS_CUDA(cudaMemcpyAsync(dq, data_ptr<float>(q), sizeof(float) * (size_t)nq * h->d, cudaMemcpyHostToDevice, st));
      run(dq, dD, k, dI, k);
      RS_CUDA(cudaMemcpyAsync(data_ptr<float>(D), dD, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
      RS_CUDA(cudaMemcpyAsync(data_ptr<int64_t>(I), dI, sizeof(int64_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
    } catch (...) {
      cudaFreeAsync(dq, st);
      cudaFreeAsync(dD, st);
      cudaFreeAsync(dI, st);
      cudaStreamSynchronize(st);
      throw;
    }
    cudaFreeAsync(dq, st);
    cudaFreeAsync(dD, st);
    cudaFreeAsync(dI, st);
    RS_CUDA(cudaStreamSynchronize(st));
  });
}

int rs_index_search_deferred(rs_index_t* h, const DLTensor* q, int k, int64_t id_base, int mode, DLTensor* D, DLTensor* I,
                             DLTensor* need_exact, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    RS_CHECK(h != nullptr, "index is NULL");
    require(q, "q", 2, kDLFloat, 32);
    require(D, "D", 2, kDLFloat, 32);
    require(I, "I", 2, kDLInt, 64);
    require(need_exact, "need_exact", 1, kDLInt, 32);
    RS_CHECK(q->shape[1] == h->d, "q must be [nq, d] with d == index dimension");
    RS_CHECK(k >= 1 && k <= 1024, "k must be in [1, 1024]");
    const int64_t nq = q->shape[0];
    RS_CHECK(D->shape[0] == nq && D->shape[1] == k && I->shape[0] == nq && I->shape[1] == k, "D and I must be [nq, k]");
    RS_CHECK(need_exact->shape[0] == nq && (nq <= 1 || stride_of(need_exact, 0) == 1), "need_exact must be a contiguous [nq] tensor");
    RS_CHECK(mode == RS_EXACT_FP32 || mode == RS_TENSOR_BF16 || mode == RS_AUTO, "unknown mode");
    RS_CHECK(on_device(q) && on_device(D) && on_device(I) && on_device(need_exact), "the deferred search takes CUDA tensors only");
    if (nq == 0) return;
    DeviceGuard g(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RS_CHECK(is_compact(q), "q must be contiguous");
    RS_CHECK(q->device.device_id == h->device, "q is on a different device than the index");
    require_aligned16(data_ptr<float>(q), "q");
    const bool supported = tensor_path_supported(h, nq, k);
    RS_CHECK(mode != RS_TENSOR_BF16 || supported, "tensor mode is not applicable to this shape (needs d=128, k<=128 and ntotal>=65536, or 256<=ntotal<=32768 with k<=ntotal/128)");
    if (mode != RS_EXACT_FP32 && supported) {
      search_tensor_device(h, data_ptr<float>(q), nq, k, id_base, data_ptr<float>(D), row_stride_2d(D, "D"),
                           data_ptr<int64_t>(I), row_stride_2d(I, "I"), st, data_ptr<int>(need_exact));
    } else {
      search_exact_device(h, data_ptr<float>(q), nq, k, id_base, data_ptr<float>(D), row_stride_2d(D, "D"),
                          data_ptr<int64_t>(I), row_stride_2d(I, "I"), st);
      RS_CUDA(cudaMemsetAsync(data_ptr<int>(need_exact), 0, sizeof(int) * (size_t)nq, st));
    }
  });
}

int rs_normalize_l2(DLTensor* x, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(x, "x", 2, kDLFloat, 32);
    const int64_t n = x->shape[0];
    const int d = static_cast<int>(x->shape[1]);
    if (n == 0 || d == 0) return;
    const int64_t ld = row_stride_2d(x, "x");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sm_count();
    const unsigned blocks = static_cast<unsigned>((n + 127) / 128);
    if (on_device(x)) {
      normalize_l2_kernel<<<blocks, 128, 0, st>>>(data_ptr<float>(x), ld, n, d);
      RS_LAUNCH_CHECK();
      return;
    }
    RS_CHECK(on_host(x), "x must be a CUDA or host tensor");
    float* dx = nullptr;
    RS_CUDA(cudaMallocAsync(&dx, sizeof(float) * (size_t)n * d, st));
    try {
      RS_CUDA(cudaMemcpy2DAsync(dx, sizeof(float) * d, data_ptr<float>(x), sizeof(float) * ld, sizeof(float) * d, (size_t)n,
                                cudaMemcpyHostToDevice, st));
      normalize_l2_kernel<<<blocks, 128, 0, st>>>(dx, d, n, d);
      RS_LAUNCH_CHECK();
      RS_CUDA(cudaMemcpy2DAsync(data_ptr<float>(x), sizeof(float) * ld, dx, sizeof(float) * d, sizeof(float) * d, (size_t)n,
                                cudaMemcpyDeviceToHost, st));
    } catch (...) {
      cudaFreeAsync(dx, st);
      cudaStreamSynchronize(st);
      throw;
    }
    cudaFreeAsync(dx, st);
    RS_CUDA(cudaStreamSynchronize(st));
  });
}

int rs_matvec(const DLTensor* items, const DLTensor* user, DLTensor* out, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(items, "items", 2, kDLFloat, 32);
    require(user, "user", 1, kDLFloat, 32);
    require(out, "out", 1, kDLFloat, 32);
    RS_CHECK(on_device(items) && on_device(user) && on_device(out), "all tensors must be CUDA tensors");
    const int64_t n = items->shape[0];
    const int d = static_cast<int>(items->shape[1]);
    RS_CHECK(user->shape[0] == d && out->shape[0] == n, "shape mismatch: items [n, d], user [d], out [n]");
    RS_CHECK((d <= 1 || stride_of(user, 0) == 1) && (n <= 1 || stride_of(out, 0) == 1), "user and out must be contiguous");
    if (n == 0) return;
    sm_count();
    matvec_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        data_ptr<float>(items), row_stride_2d(items, "items"), data_ptr<float>(user), data_ptr<float>(out), n, d);
    RS_LAUNCH_CHECK();
  });
}

int rs_topk_verify(const DLTensor* D, const DLTensor* I, const DLTensor* bounds, const DLTensor* flags,
                   DLTensor* need_exact, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(D, "D", 2, kDLFloat, 32);
    require(I, "I", 2, kDLInt, 64);
    require(bounds, "bounds", 2, kDLFloat, 32);
    require(flags, "flags", 2, kDLInt, 32);
    require(need_exact, "need_exact", 1, kDLInt, 32);
    const int64_t nq = D->shape[0], k = D->shape[1], G = bounds->shape[0];
    RS_CHECK(I->shape[0] == nq && I->shape[1] == k && bounds->shape[1] == nq && flags->shape[0] == G && flags->shape[1] == nq &&
                 need_exact->shape[0] == nq, "shapes: D, I [nq, k]; bounds, flags [shards, nq]; need_exact [nq]");
    RS_CHECK(on_device(D) && on_device(I) && on_device(bounds) && on_device(flags) && on_device(need_exact), "all tensors must be CUDA tensors");
    RS_CHECK((nq <= 1 || (stride_of(bounds, 1) == 1 && stride_of(flags, 1) == 1 && stride_of(need_exact, 0) == 1)) && k >= 1,
             "bounds / flags rows and need_exact must have unit stride");
    if (nq == 0) return;
    sm_count();
    verify_topk_kernel<<<static_cast<unsigned>((nq + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        data_ptr<float>(D), row_stride_2d(D, "D"), data_ptr<int64_t>(I), row_stride_2d(I, "I"), static_cast<int>(k),
        data_ptr<float>(bounds), G > 1 ? stride_of(bounds, 0) : nq, data_ptr<int>(flags), G > 1 ? stride_of(flags, 0) : nq,
        static_cast<int>(G), nq, data_ptr<int>(need_exact));
    RS_LAUNCH_CHECK();
  });
}

int rs_topk_merge(const DLTensor* Dparts, const DLTensor* Iparts, DLTensor* D, DLTensor* I, void* stream) {
  return rs::guard([&] {
    using namespace rs;
    require(Dparts, "Dparts", 3, kDLFloat, 32);
    require(Iparts, "Iparts", 3, kDLInt, 64);
    require(D, "D", 2, kDLFloat, 32);
    require(I, "I", 2, kDLInt, 64);
    RS_CHECK(on_device(Dparts) && on_device(Iparts) && on_device(D) && on_device(I), "all tensors must be CUDA tensors");
    const int64_t P = Dparts->shape[0], nq = Dparts->shape[1], k = Dparts->shape[2];
    RS_CHECK(Iparts->shape[0] == P && Iparts->shape[1] == nq && Iparts->shape[2] == k, "Iparts must match Dparts");
    RS_CHECK(D->shape[0] == nq && D->shape[1] == k && I->shape[0] == nq && I->shape[1] == k, "D and I must be [nq, k]");
    // each part is a packed [nq, k] block; the parts themselves may be strided (slices of one exchange buffer)
    auto part_ld = [&](const DLTensor* t) -> int64_t {
      if (!t->strides) return nq * k;
      RS_CHECK((k <= 1 || t->strides[2] == 1) && (nq <= 1 || t->strides[1] == k) && (P <= 1 || t->strides[0] >= nq * k),
               "Dparts / Iparts: each part must be a contiguous [nq, k] block");
      return P <= 1 ? nq * k : t->strides[0];
    };
    RS_CHECK(P >= 1 && P <= 4096, "nparts must be in [1, 4096]");
    if (nq == 0) return;
    sm_count();
    KeySource src{nullptr, data_ptr<float>(Dparts), data_ptr<int64_t>(Iparts), nq, part_ld(Dparts), part_ld(Iparts)};
    merge_dispatch(src, nq, static_cast<int>(P), static_cast<int>(k), 0, data_ptr<float>(D), row_stride_2d(D, "D"),
                   data_ptr<int64_t>(I), row_stride_2d(I, "I"), static_cast<cudaStream_t>(stream));
  });
}

}  // extern "C"
