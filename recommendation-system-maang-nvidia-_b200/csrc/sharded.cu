// sharded.cu -- the multi-GPU paths behind plain C entry points, so that a host in any language can drive them
// (include/recsys_b200.h: rs_comm_*, rs_index_search_sharded, rs_sharded_table_*, rs_embedding_a2a_lookup).
//
// One process per GPU.  A communicator (rs_comm) is this rank's view of the group: rank, world and ONE host-side
// all-gather callback supplied by the caller (torch.distributed, MPI, a socket -- the library does not care).  The
// callback is used for set-up and teardown only: it carries the 64-byte CUDA IPC handles of the peer-memory regions
// (csrc/peer.cu) and serves as the barrier around mapping / unmapping them.  Every per-step byte travels through this
// repo's own kernels over NVLink peer memory; no NCCL, no host synchronisation inside a step.
//
// Sharded catalog search (BASELINE config 3; replaces faiss IndexFlatIP.search of app/recommendation_service.py:72 over a
// catalog too large or too slow for one GPU).  Rank g holds rows [N*g/G, N*(g+1)/G).  One step, all on the caller's stream:
//   sample pass + top list        rs::shared_begin; sample_top_kernel stores each query's list into every rank's tops region
//                                 (tops exchange), its last CTA waits for all ranks' tops
//   threshold (acks)              threshold_shared_kernel reads the G lists in place
//   filter, re-score              rs::shared_finish; the filter appends survivors to the per-query lists itself,
//                                 select_rescore_kernel stores [scores | rows | flags | bounds] of each query into every
//                                 rank's lists region (lists exchange), its last CTA waits for all ranks' lists
//   merge (verifies, acks)        merge_topk_kernel reads the G parts in place and writes D, I, need_exact
// 7 kernel launches and 2 memsets per step (round 2 started at 17; 11 with separate publish / bucket / flag launches).  Shapes outside the tensor path (or RS_EXACT_FP32) run every shard's
// own search (per-shard thresholds or the exact fp32 scan) and share the lists exchange + merge.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "flat_index.cuh"
#include "peer.cuh"

struct rs_sharded_table;

namespace {

int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// byte offsets of one shard's message in the lists exchange (the same layout as sharded.exchange_layout in Python):
// scores f32 [nq, k] | rows int64 [nq, k] | need_exact int32 [nq] | bound f32 [nq]
struct ListsLayout {
  int64_t off_I, off_F, off_B, total;
  ListsLayout(int64_t nq, int k) {
    off_I = round_up(nq * k * 4, 16);
    off_F = off_I + nq * k * 8;
    off_B = off_F + nq * 4;
    total = round_up(off_F + nq * 8, 16);
  }
};

struct SearchPlan {
  int64_t nq_cap = 0;
  int k = 0;
  int rank = 0, rank_pad = 0;       // shared-threshold rank for this k (0: not applicable) and its 4-float padding
  rs_peer* tops = nullptr;          // slots of nq_cap * rank_pad floats
  rs_peer* lists = nullptr;         // slots of ListsLayout(nq_cap, k).total bytes
  float* top = nullptr;             // [nq_cap, rank_pad] this shard's contribution
  uint8_t* send = nullptr;          // this shard's lists message
};

}  // namespace

struct rs_comm {
  int rank = 0, world = 1, device = 0;
  rs_allgather_fn allgather = nullptr;
  void* ctx = nullptr;
  std::mutex mu;
  std::vector<std::unique_ptr<SearchPlan>> plans;
  int live_tables = 0;
};

struct rs_sharded_table {
  rs_comm* comm = nullptr;
  const float* table = nullptr;     // borrowed [n_local, dim]
  int64_t n_local = 0, ld = 0, b_max = 0, H = 0;
  int dim = 0;
  bool dedup = false;
  rs_peer *ids = nullptr, *rows = nullptr;
  int64_t *counts = nullptr, *stage = nullptr, *keys = nullptr, *inv = nullptr;
  int32_t *vals = nullptr, *slot_of = nullptr;
};

namespace {

using namespace rs;

// host all-gather of `bytes` per rank through the caller's callback (identity for a single rank)
void all_gather(rs_comm* c, const void* send, void* recv, int64_t bytes) {
  if (c->world == 1) {
    std::memcpy(recv, send, static_cast<size_t>(bytes));
    return;
  }
  RS_CHECK(c->allgather(c->ctx, send, recv, bytes) == 0, "the caller's all-gather callback failed");
}
void barrier(rs_comm* c) {
  uint8_t one = 1;
  std::vector<uint8_t> all(static_cast<size_t>(c->world));
  all_gather(c, &one, all.data(), 1);
}
// every rank reports ok / not ok; true iff all are ok (so that a failure anywhere is a failure everywhere)
bool agree(rs_comm* c, bool ok) {
  const uint8_t mine = ok ? 1 : 0;
  std::vector<uint8_t> all(static_cast<size_t>(c->world));
  all_gather(c, &mine, all.data(), 1);
  return std::all_of(all.begin(), all.end(), [](uint8_t v) { return v != 0; });
}

// Collective: one peer-memory region of `world` slots, mapped on every rank.  Every rank walks through the same
// all-gathers whether or not its own CUDA calls succeed.
rs_peer* make_region(rs_comm* c, int64_t slot_bytes) {
  rs_peer* h = nullptr;
  uint8_t handle[64] = {};
  std::string err;
  bool ok = rs_peer_create(c->rank, c->world, round_up(slot_bytes, 16), &h, handle) == 0;
  if (!ok) err = rs_last_error();
  std::vector<uint8_t> all(static_cast<size_t>(c->world) * 64);
  all_gather(c, handle, all.data(), 64);
  if (ok) {
    ok = rs_peer_connect(h, all.data()) == 0;
    if (!ok) err = rs_last_error();
  }
  if (!agree(c, ok)) {
    if (h) {
      rs_peer_disconnect(h);
      barrier(c);
      rs_peer_free(h);
    } else {
      barrier(c);
    }
    throw Error("peer-memory region could not be set up on every rank" + (err.empty() ? std::string() : ": " + err));
  }
  return h;   // agree() was also the barrier: every rank has every region mapped before anybody publishes
}

// Collective teardown of regions: nobody frees a region a peer may still write into or still has mapped.
void free_regions(rs_comm* c, std::vector<rs_peer*> regions) {
  cudaDeviceSynchronize();
  barrier(c);
  for (rs_peer* r : regions)
    if (r) rs_peer_disconnect(r);
  barrier(c);
  for (rs_peer* r : regions)
    if (r) rs_peer_free(r);
}

int shared_rank(const rs_comm* c, int d, int64_t n_total, int64_t nq, int k) {
  int64_t sampled = 0;
  for (int g = 0; g < c->world; ++g) {
    const int64_t rows = n_total * (g + 1) / c->world - n_total * g / c->world;
    int64_t s = 0;
    int ok = 0;
    if (rs_tensor_plan(d, rows, nq, k, &s, &ok) != 0 || !ok) return 0;
    sampled += s;
  }
  const int r = tensor_threshold_rank(k, double(sampled) / double(std::max<int64_t>(1, n_total)));
  // every shard sends its r largest sample maxima (its sample has at least 256 of them); the union must fit the kernel
  return (r <= 128 && (int64_t)round_up(r, 4) * c->world <= 1024) ? r : 0;
}

SearchPlan* plan_for(rs_comm* c, int d, int64_t n_total, int64_t nq, int k) {
  const int rank = shared_rank(c, d, n_total, nq, k);
  for (auto& p : c->plans)
    if (p->k == k && p->nq_cap >= nq && (rank == 0 || p->rank_pad >= round_up(rank, 4))) return p.get();
  // collective creation: all ranks reach this point for the same (nq, k) because they issue the same calls
  auto p = std::make_unique<SearchPlan>();
  p->nq_cap = nq;
  p->k = k;
  p->rank = rank;
  p->rank_pad = static_cast<int>(round_up(rank, 4));
  const ListsLayout L(nq, k);
  if (p->rank_pad > 0) p->tops = make_region(c, nq * p->rank_pad * 4);
  try {
    p->lists = make_region(c, L.total);
  } catch (...) {
    if (p->tops) free_regions(c, {p->tops});
    throw;
  }
  RS_CUDA(cudaMalloc(&p->top, sizeof(float) * static_cast<size_t>(nq) * std::max(4, p->rank_pad)));
  RS_CUDA(cudaMalloc(&p->send, static_cast<size_t>(L.total)));
  c->plans.push_back(std::move(p));
  return c->plans.back().get();
}

struct DL {
  DLTensor t{};
  int64_t shape[3] = {0, 0, 0};
  DL(void* p, int device, uint8_t code, uint8_t bits, int64_t s0, int64_t s1 = -1, int64_t s2 = -1) {
    shape[0] = s0; shape[1] = s1; shape[2] = s2;
    t.data = p;
    t.device = DLDevice{kDLCUDA, device};
    t.ndim = s2 >= 0 ? 3 : s1 >= 0 ? 2 : 1;
    t.dtype = DLDataType{code, bits, 1};
    t.shape = shape;
    t.strides = nullptr;
    t.byte_offset = 0;
  }
  DL(const DL&) = delete;
  operator const DLTensor*() const { return &t; }
  operator DLTensor*() { return &t; }
};

// RS_PEER_WAIT_IN_CONSUMER=1 (A/B knob, read per call): publish without waiting; the consumer kernels then spin themselves
void publish(rs_peer* region, const void* src, int64_t bytes, cudaStream_t st) {
  const char* e = getenv("RS_PEER_WAIT_IN_CONSUMER");
  if (e && e[0] == '1') {
    DL m(const_cast<void*>(src), region->device, kDLUInt, 8, bytes);
    RS_CHECK(rs_peer_publish(region, m, st) == 0, rs_last_error());
  } else {
    rsp::publish_and_wait(region, src, bytes, st);
  }
}

// one deferred step over `nq` queries (device pointers); need = int32 [nq] device or nullptr (exact mode: nothing to flag)
void search_step(rs_index* shard, rs_comm* c, SearchPlan* p, int64_t n_total, const float* q, int64_t nq, int k, int mode,
                 float* D, int64_t ldD, int64_t* I, int64_t ldI, int* need, cudaStream_t st) {
  const int G = c->world;
  const int64_t id_base = n_total * c->rank / G;
  const ListsLayout L(nq, k);
  float* sD = reinterpret_cast<float*>(p->send);
  int64_t* sI = reinterpret_cast<int64_t*>(p->send + L.off_I);
  int* sF = reinterpret_cast<int*>(p->send + L.off_F);
  float* sB = reinterpret_cast<float*>(p->send + L.off_B);
  const bool tensor_ok = mode != RS_EXACT_FP32 && tensor_path_supported(shard, nq, k) && nq <= tensor_max_queries();
  // one or two queries: every shard's own two-launch scan through its bf16 shadow beats the shared-threshold protocol (one
  // exchange and five launches fewer); the lists exchange + merge below is the same
  const bool tiny = tensor_ok && bf16_scan_applies(shard, nq, k);
  const int rank = (mode != RS_EXACT_FP32 && !tiny) ? shared_rank(c, shard->d, n_total, nq, k) : 0;
  const bool shared = rank > 0 && p->tops != nullptr && p->rank_pad >= round_up(rank, 4);
  bool lists_published = false;   // the re-scoring kernel wrote the lists message into the peers itself
  if (shared) {
    const int R = static_cast<int>(round_up(rank, 4));
    // RS_PEER_FUSED_PUBLISH=0 (A/B knob, read per call): the round-2 sequence with separate publish launches
    const char* e = getenv("RS_PEER_FUSED_PUBLISH");
    const bool fused = !(e && e[0] == '0');
    rs_tc_session* sess = shared_begin(shard, q, nq, k, R, p->top, R, st, fused ? p->tops : nullptr);
    if (!fused) {
      try {
        publish(p->tops, p->top, nq * R * 4, st);
      } catch (...) {
        tc_session_free(sess);
        throw;
      }
    }
    shared_finish(sess, reinterpret_cast<const float*>(p->tops->local), G, p->tops->slot_bytes / 4, R, rank, id_base, sD, k,
                  sI, k, sF, sB, p->tops, fused ? p->lists : nullptr, L.off_I, L.off_F, L.off_B);
    lists_published = fused;
  } else if (tensor_ok) {
    search_tensor_device(shard, q, nq, k, id_base, sD, k, sI, k, st, sF);
  } else {
    search_exact_device(shard, q, nq, k, id_base, sD, k, sI, k, st);
    RS_CUDA(cudaMemsetAsync(sF, 0, sizeof(int) * static_cast<size_t>(nq), st));
  }
  if (!lists_published) publish(p->lists, p->send, L.total, st);
  const uint8_t* R0 = p->lists->local;
  const int64_t slot = p->lists->slot_bytes;
  merge_lists_device(reinterpret_cast<const float*>(R0), slot / 4, reinterpret_cast<const int64_t*>(R0 + L.off_I), slot / 8, G,
                     nq, k, D, ldD, I, ldI, p->lists, shared ? reinterpret_cast<const float*>(R0 + L.off_B) : nullptr, slot / 4,
                     reinterpret_cast<const int*>(R0 + L.off_F), slot / 4, need, st);
}

}  // namespace

extern "C" {

int rs_comm_init(int rank, int world, rs_allgather_fn allgather, void* ctx, rs_comm_t** out) {
  return rs::guard([&] {
    RS_CHECK(out != nullptr, "out is NULL");
    RS_CHECK(world >= 1 && world <= rsp::kMaxWorld && rank >= 0 && rank < world, "rank / world out of range (world <= 16)");
    RS_CHECK(world == 1 || allgather != nullptr, "an all-gather callback is required for world > 1");
    auto c = std::make_unique<rs_comm>();
    c->rank = rank;
    c->world = world;
    c->allgather = allgather;
    c->ctx = ctx;
    RS_CUDA(cudaGetDevice(&c->device));
    // the callback must work and the ranks must agree on the geometry before anything is built on it
    std::vector<int32_t> all(static_cast<size_t>(world) * 2);
    const int32_t mine[2] = {rank, world};
    all_gather(c.get(), mine, all.data(), sizeof(mine));
    for (int g = 0; g < world; ++g)
      RS_CHECK(all[2 * g] == g && all[2 * g + 1] == world, "all-gather callback returned inconsistent ranks (slot g must hold rank g's bytes)");
    *out = c.release();
  });
}

int rs_comm_rank(const rs_comm_t* c, int* rank, int* world) {
  return rs::guard([&] {
    RS_CHECK(c != nullptr && rank != nullptr && world != nullptr, "NULL argument");
    *rank = c->rank;
    *world = c->world;
  });
}

int rs_comm_free(rs_comm_t* c) {
  return rs::guard([&] {
    if (!c) return;
    RS_CHECK(c->live_tables == 0, "free the communicator's sharded tables first");
    std::unique_ptr<rs_comm> own(c);
    rsp::PeerDeviceGuard g(c->device);
    std::vector<rs_peer*> regions;
    for (auto& p : c->plans) {
      regions.push_back(p->tops);
      regions.push_back(p->lists);
    }
    free_regions(c, regions);   // collective, also with no regions: the ranks leave together
    for (auto& p : c->plans) {
      cudaFree(p->top);
      cudaFree(p->send);
    }
  });
}

int rs_index_search_sharded(rs_index_t* shard, rs_comm_t* c, int64_t n_total, const DLTensor* q, int k, int mode, DLTensor* D,
                            DLTensor* I, DLTensor* need_exact, void* stream) {
  return rs::guard([&] {
    RS_CHECK(shard != nullptr && c != nullptr, "shard / comm is NULL");
    require(q, "q", 2, kDLFloat, 32);
    require(D, "D", 2, kDLFloat, 32);
    require(I, "I", 2, kDLInt, 64);
    const int64_t nq = q->shape[0];
    RS_CHECK(q->shape[1] == shard->d, "q must be [nq, d] with d == index dimension");
    RS_CHECK(k >= 1 && k <= 1024, "k must be in [1, 1024]");
    RS_CHECK(D->shape[0] == nq && D->shape[1] == k && I->shape[0] == nq && I->shape[1] == k, "D and I must be [nq, k]");
    RS_CHECK(mode == RS_EXACT_FP32 || mode == RS_TENSOR_BF16 || mode == RS_AUTO, "unknown mode");
    RS_CHECK(on_device(q) && on_device(D) && on_device(I) && is_compact(q), "q, D, I must be CUDA tensors, q contiguous");
    RS_CHECK(q->device.device_id == shard->device && c->device == shard->device, "q, the shard and the communicator must share one device");
    RS_CHECK(n_total >= 0 && n_total <= 0xFFFFFFFFll, "sharded catalogs above 2^32 rows are not supported (rows are merged as 32-bit keys)");
    const int64_t lo = n_total * c->rank / c->world, hi = n_total * (c->rank + 1) / c->world;
    RS_CHECK(shard->ntotal == hi - lo, "this rank's shard must hold rows [N*rank/world, N*(rank+1)/world) of the catalog");
    int* need = nullptr;
    if (need_exact) {
      require(need_exact, "need_exact", 1, kDLInt, 32);
      RS_CHECK(on_device(need_exact) && need_exact->shape[0] == nq && (nq <= 1 || stride_of(need_exact, 0) == 1),
               "need_exact must be a contiguous CUDA int32 [nq] tensor");
      need = data_ptr<int>(need_exact);
    }
    if (nq == 0) return;
    require_aligned16(data_ptr<float>(q), "q");
    DeviceGuard g(shard->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::lock_guard<std::mutex> lk(c->mu);   // one step at a time per communicator: the regions carry one message per rank
    SearchPlan* p = plan_for(c, shard->d, n_total, nq, k);
    float* Dp = data_ptr<float>(D);
    int64_t* Ip = data_ptr<int64_t>(I);
    const int64_t ldD = row_stride_2d(D, "D"), ldI = row_stride_2d(I, "I");
    if (need || mode == RS_EXACT_FP32) {
      search_step(shard, c, p, n_total, data_ptr<float>(q), nq, k, mode, Dp, ldD, Ip, ldI, need, st);   // exact lists carry no flags
      return;
    }
    // need_exact == NULL: resolve flagged queries here, at the price of one stream synchronise.  The flags are the same on
    // every rank (the verdict is a function of the merged data), so all ranks take the same branch.
    int* d_need = nullptr;
    RS_CUDA(cudaMallocAsync(&d_need, sizeof(int) * static_cast<size_t>(nq), st));
    std::vector<int> h_need(static_cast<size_t>(nq));
    try {
      search_step(shard, c, p, n_total, data_ptr<float>(q), nq, k, mode, Dp, ldD, Ip, ldI, d_need, st);
      RS_CUDA(cudaMemcpyAsync(h_need.data(), d_need, sizeof(int) * static_cast<size_t>(nq), cudaMemcpyDeviceToHost, st));
      RS_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
      cudaFreeAsync(d_need, st);
      throw;
    }
    cudaFreeAsync(d_need, st);
    if (const char* e = getenv("RS_DEBUG_FLAG_QUERIES")) {   // test hook: comma-separated query numbers treated as flagged
      for (const char* s = e; *s;) {
        char* end = nullptr;
        const long v = strtol(s, &end, 10);
        if (end == s) break;
        if (v >= 0 && v < nq) h_need[static_cast<size_t>(v)] = 1;
        s = (*end == ',') ? end + 1 : end;
      }
    }
    std::vector<int64_t> idx;
    for (int64_t i = 0; i < nq; ++i)
      if (h_need[static_cast<size_t>(i)]) idx.push_back(i);
    if (idx.empty()) return;
    const int64_t ns = static_cast<int64_t>(idx.size());
    const int d = shard->d;
    float *qs = nullptr, *Ds = nullptr;
    int64_t* Is = nullptr;
    RS_CUDA(cudaMallocAsync(&qs, sizeof(float) * static_cast<size_t>(ns) * d, st));
    RS_CUDA(cudaMallocAsync(&Ds, sizeof(float) * static_cast<size_t>(ns) * k, st));
    RS_CUDA(cudaMallocAsync(&Is, sizeof(int64_t) * static_cast<size_t>(ns) * k, st));
    try {
      for (int64_t j = 0; j < ns; ++j)
        RS_CUDA(cudaMemcpyAsync(qs + j * d, data_ptr<float>(q) + idx[j] * d, sizeof(float) * d, cudaMemcpyDeviceToDevice, st));
      search_step(shard, c, p, n_total, qs, ns, k, RS_EXACT_FP32, Ds, k, Is, k, nullptr, st);
      for (int64_t j = 0; j < ns; ++j) {
        RS_CUDA(cudaMemcpyAsync(Dp + idx[j] * ldD, Ds + j * k, sizeof(float) * k, cudaMemcpyDeviceToDevice, st));
        RS_CUDA(cudaMemcpyAsync(Ip + idx[j] * ldI, Is + j * k, sizeof(int64_t) * k, cudaMemcpyDeviceToDevice, st));
      }
    } catch (...) {
      cudaFreeAsync(qs, st);
      cudaFreeAsync(Ds, st);
      cudaFreeAsync(Is, st);
      throw;
    }
    cudaFreeAsync(qs, st);
    cudaFreeAsync(Ds, st);
    cudaFreeAsync(Is, st);
    shard->last_fallbacks.store(ns, std::memory_order_relaxed);
    shard->total_fallbacks.fetch_add(ns, std::memory_order_relaxed);
  });
}

int rs_sharded_table_create(rs_comm_t* c, const DLTensor* local_rows, int64_t max_batch, int dedup, rs_sharded_table_t** out) {
  return rs::guard([&] {
    RS_CHECK(c != nullptr && out != nullptr, "comm / out is NULL");
    require(local_rows, "local_rows", 2, kDLFloat, 32);
    RS_CHECK(on_device(local_rows) && local_rows->device.device_id == c->device, "local_rows must be a CUDA tensor on the communicator's device");
    RS_CHECK(local_rows->shape[1] >= 4 && local_rows->shape[1] % 4 == 0, "row width must be a positive multiple of 4 floats");
    RS_CHECK(max_batch >= 1 && max_batch <= (int64_t(1) << 28), "max_batch out of range");
    auto t = std::make_unique<rs_sharded_table>();
    t->comm = c;
    t->table = data_ptr<float>(local_rows);
    t->n_local = local_rows->shape[0];
    t->dim = static_cast<int>(local_rows->shape[1]);
    t->ld = row_stride_2d(local_rows, "local_rows");
    RS_CHECK(t->ld % 4 == 0, "row stride must be a multiple of 4 floats");
    require_aligned16(t->table, "local_rows");
    t->b_max = max_batch;
    t->dedup = dedup != 0;
    rsp::PeerDeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    t->ids = make_region(c, 16 + 8 * max_batch);
    try {
      t->rows = make_region(c, max_batch * t->dim * 4);
    } catch (...) {
      free_regions(c, {t->ids});
      throw;
    }
    const int G = c->world;
    RS_CUDA(cudaMalloc(&t->counts, sizeof(int64_t) * G));
    RS_CUDA(cudaMalloc(&t->stage, sizeof(int64_t) * static_cast<size_t>(G) * max_batch));
    RS_CUDA(cudaMemset(t->stage, 0, sizeof(int64_t) * static_cast<size_t>(G) * max_batch));
    if (t->dedup) {
      int64_t H = 16;
      while (H < 2 * max_batch) H <<= 1;
      t->H = H;
      RS_CUDA(cudaMalloc(&t->keys, sizeof(int64_t) * static_cast<size_t>(H)));
      RS_CUDA(cudaMalloc(&t->vals, sizeof(int32_t) * static_cast<size_t>(H)));
      RS_CUDA(cudaMalloc(&t->slot_of, sizeof(int32_t) * static_cast<size_t>(max_batch)));
    } else {
      RS_CUDA(cudaMalloc(&t->inv, sizeof(int64_t) * static_cast<size_t>(max_batch)));
    }
    RS_CUDA(cudaDeviceSynchronize());
    ++c->live_tables;
    *out = t.release();
  });
}

int rs_sharded_table_free(rs_sharded_table_t* t) {
  return rs::guard([&] {
    if (!t) return;
    std::unique_ptr<rs_sharded_table> own(t);
    rs_comm* c = t->comm;
    rsp::PeerDeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    free_regions(c, {t->ids, t->rows});
    cudaFree(t->counts);
    cudaFree(t->stage);
    cudaFree(t->keys);
    cudaFree(t->vals);
    cudaFree(t->slot_of);
    cudaFree(t->inv);
    --c->live_tables;
  });
}

int rs_sharded_table_last_counts(rs_sharded_table_t* t, int64_t* counts_out, void* stream) {
  return rs::guard([&] {
    RS_CHECK(t != nullptr && counts_out != nullptr, "NULL argument");
    rsp::PeerDeviceGuard g(t->comm->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RS_CUDA(cudaMemcpyAsync(counts_out, t->counts, sizeof(int64_t) * t->comm->world, cudaMemcpyDeviceToHost, st));
    RS_CUDA(cudaStreamSynchronize(st));
  });
}

int rs_embedding_a2a_lookup(rs_sharded_table_t* t, const DLTensor* ids, DLTensor* out, void* stream) {
  return rs::guard([&] {
    RS_CHECK(t != nullptr, "table is NULL");
    require(ids, "ids", 1, kDLInt, 64);
    require(out, "out", 2, kDLFloat, 32);
    const int64_t B = ids->shape[0];
    RS_CHECK(on_device(ids) && on_device(out) && out->shape[0] == B && out->shape[1] == t->dim, "ids [B] and out [B, dim] must be CUDA tensors");
    RS_CHECK(B <= t->b_max, "batch exceeds the max_batch the table's peer regions were sized for");
    const int dev = t->comm->device, G = t->comm->world;
    rsp::PeerDeviceGuard g(dev);
    DL counts(t->counts, dev, kDLInt, 64, G), stage(t->stage, dev, kDLInt, 64, G, t->b_max);
    DL table(const_cast<float*>(t->table), dev, kDLFloat, 32, t->n_local, t->dim);
    int64_t tstrides[2] = {t->ld, 1};
    table.t.strides = tstrides;
    auto ck = [](int rc) { RS_CHECK(rc == 0, rs_last_error()); };
    if (t->dedup) {
      DL keys(t->keys, dev, kDLInt, 64, t->H), vals(t->vals, dev, kDLInt, 32, t->H), slot_of(t->slot_of, dev, kDLInt, 32, t->b_max);
      ck(rs_a2a_plan_peer(ids, G, 1, keys, vals, slot_of, counts, stage, nullptr, stream));
      ck(rs_peer_a2a_ids(t->ids, stage, counts, stream));
      ck(rs_peer_a2a_rows(t->ids, t->rows, table, stream));
      ck(rs_peer_a2a_restore(t->rows, ids, slot_of, vals, nullptr, out, stream));
    } else {
      DL inv(t->inv, dev, kDLInt, 64, B);
      ck(rs_a2a_plan_peer(ids, G, 0, nullptr, nullptr, nullptr, counts, stage, inv, stream));
      ck(rs_peer_a2a_ids(t->ids, stage, counts, stream));
      ck(rs_peer_a2a_rows(t->ids, t->rows, table, stream));
      ck(rs_peer_a2a_restore(t->rows, ids, nullptr, nullptr, inv, out, stream));
    }
  });
}

}  // extern "C"
