/*
 * recsys_b200.h -- C ABI of librecsys_b200.so, the B200 (sm_100a) drop-in for the
 * serving/scoring hot path of OnlyAhad13/Recommendation-System-MAANG-NVIDIA-.
 *
 * The reference has no FFI of its own (it is pure Python over TensorFlow and
 * faiss-cpu); each entry point below names the reference call site it replaces
 * (paths relative to the reference root).  Conventions:
 *   - every function returns 0 on success and non-zero on failure; the message is
 *     then available from rs_last_error() (thread-local, valid until the next call
 *     on that thread).  There is no CPU fallback: a missing GPU is an error.
 *   - tensors cross the boundary as borrowed DLPack `DLTensor*` (the struct below is
 *     ABI-identical to dlpack.h v0.8 / v1.0 `DLTensor`).  The library never frees or
 *     keeps caller memory; the only memory it owns sits behind opaque handles.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *     Calls with device tensors are asynchronous with respect to the host.  Calls
 *     that are handed host tensors (device_type kDLCPU / kDLCUDAHost) stage them
 *     through the device on `stream` and return after the results are back in the
 *     host buffers (this is the faiss-like numpy-in / numpy-out path).
 *   - handles are immutable between rs_index_add calls, so concurrent searches on
 *     one handle from several host threads are allowed.
 */
#ifndef RECSYS_B200_H_
#define RECSYS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- minimal DLPack tensor ABI (matches dlpack.h; include that instead if you have it) ---- */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U, kDLBfloat = 4U } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides; /* in elements; NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;
#endif

#define RS_ABI_VERSION 1

/* activation codes for rs_dense_forward */
#define RS_ACT_NONE 0
#define RS_ACT_RELU 1
#define RS_ACT_SIGMOID 2

/* arithmetic modes of the search and ranking entries */
#define RS_EXACT_FP32 0 /* fp32 FMA chain; index-exact and bit-exact vs the oracle        */
#define RS_TENSOR_BF16 1 /* tcgen05 bf16 pre-filter + exact fp32 re-score (same results)   */
#define RS_AUTO 2        /* library picks by query count and catalog size                 */

int rs_version(void);
const char* rs_last_error(void);
/* number of SMs of the current device; fails (no fallback) when no CUDA device is usable */
int rs_device_sm_count(int* out);

/*
 * K0. StringLookup on the device (the default models keep the lookup in a host dict).
 * Replaces keras.layers.StringLookup(vocabulary=v, mask_token=None) (src/models.py:70,73; called at :85,89):
 * v[i] -> i + 1, every other string -> 0 (the single out-of-vocabulary bucket), int64.
 *   rs_vocab_create: words = HOST bytes [n_words, width], each row one word, NUL-padded (numpy `S<width>` layout;
 *                    a row is the string before its first NUL, so words cannot contain NUL).  Builds the hash table
 *                    on the current device.  A repeated word keeps its last position (as a dict built in order does).
 *                    Limits: n_words <= 2^28, 1 <= width <= 4096.  Table memory: 32 B per word (tagged slots at load
 *                    <= 0.25) + the words padded to a multiple of 16 bytes.
 *   rs_vocab_lookup: ids uint8 [B, W] device (same row layout, any W; rows longer than every word are out of
 *                    vocabulary), rows int64 [B] device.  Matches are decided on the bytes, never on the hash alone.
 */
typedef struct rs_vocab rs_vocab_t;
int rs_vocab_create(const uint8_t* words, int64_t n_words, int64_t width, rs_vocab_t** out);
int rs_vocab_free(rs_vocab_t* v);
int rs_vocab_size(const rs_vocab_t* v, int64_t* n_words);
int rs_vocab_lookup(const rs_vocab_t* v, const DLTensor* ids, DLTensor* rows, void* stream);

/*
 * K1. Embedding row gather fused with the concat.
 * Replaces keras.layers.Embedding lookups (src/models.py:71,74,85,89) and, when several
 * tables are given, the tf.concat that follows (src/models.py:128).
 *   tables[t] : float32 [V_t, E_t] device, row-major
 *   ids[t]    : int32 or int64 [B] device (row numbers, i.e. StringLookup output)
 *   out       : float32 [B, sum_t E_t] device; row stride may exceed the width
 * Rows outside [0, V_t) produce zeros (TensorFlow's GPU gather behaviour).
 */
int rs_gather_concat(int n_tables, const DLTensor* const* tables, const DLTensor* const* ids,
                     DLTensor* out, void* stream);

/*
 * K7b. Sharded embedding tables (BASELINE config 4; new in this build): rows sharded `row mod n_shards`, batch
 * data-parallel.  The plan kernels bucket a rank's ids by owner; the exchanges themselves are NCCL all-to-alls
 * issued by the host layer (sharded.py), the owner-side gather and the final un-permute are rs_gather_concat.
 *   rs_a2a_count:   counts[g] = #ids owned by shard g                      (ids int64 [B], counts int64 [G])
 *   rs_a2a_scatter: send[base[g] + slot] = id div G, perm[i] = that position (base/cursor int64 [G], send/perm int64 [B])
 *   rs_hash_table_fill: synthetic, host-replayable table rows row0, row0+row_step, ... into out f32 [n, dim]
 */
int rs_a2a_count(const DLTensor* ids, int n_shards, DLTensor* counts, void* stream);
int rs_a2a_scatter(const DLTensor* ids, int n_shards, const DLTensor* base, DLTensor* cursor, DLTensor* send,
                   DLTensor* perm, void* stream);
int rs_hash_table_fill(DLTensor* out, int64_t row0, int64_t row_step, uint64_t seed, void* stream);

/*
 * K2. One Keras Dense layer: y = act(x @ kernel + bias), kernel stored (in, out).
 * Replaces keras.layers.Dense inside user_tower / item_tower / deep_nets
 * (src/models.py:26-29,76-77).  x: f32 [B, in], y: f32 [B, out]; both may be column
 * slices of wider buffers (row stride >= width), which is how a tower writes straight
 * into the concatenated DCN input.  fp32 FMA arithmetic, k ascending.
 */
int rs_dense_forward(const DLTensor* x, const DLTensor* kernel, const DLTensor* bias, int act,
                     DLTensor* y, void* stream);

/*
 * K2 (bf16 tensor mode). The same Dense layer on tcgen05 tensor cores: x bf16 [B, in], wt = the Keras
 * kernel TRANSPOSED to bf16 [out, in], bias f32 [out] or NULL, fp32 accumulation.  Outputs: y_bf16 [B, out]
 * (the next layer's input) and/or y_f32 [B, out] (may be a column slice of a wider buffer); either may be
 * NULL.  (in, out) in {64,128,256}^2.  Tolerance vs the fp32 path is stated in tests/test_gpu_ranking_bf16.py.
 */
int rs_dense_forward_bf16(const DLTensor* x, const DLTensor* wt, const DLTensor* bias, int act,
                          DLTensor* y_bf16, DLTensor* y_f32, void* stream);
/*
 * K2 (fp32 tensor mode). The same Dense layer at fp32 accuracy on the tensor cores: every fp32 operand is split into two
 * fp16 numbers (v = hi + lo * 2^-11) and x.w is taken as hi.hi + 2^-11 (hi.lo + lo.hi) with fp32 accumulation (three
 * tcgen05.mma per K step; csrc/dense_x3.cu).  Tolerance vs the fp64 oracle: 1e-5 of tensor scale, the same as the FFMA
 * kernel behind rs_dense_forward (stated and measured in tests/test_gpu_dense_x3.py).  |values| < 65504.
 *   rs_dense_x3_prepare    kernel f32 [in, out] -> wt_hi, wt_lo fp16 [out, in] (once per weight change)
 *   rs_dense_x3_supported  1 when (in, out, act) is served: in, out in {64,128,256}, act RS_ACT_NONE / RS_ACT_RELU
 *   rs_dense_forward_x3    one layer, or two independent layers of the same shape in one launch (the two towers).
 * A layer's input is EITHER x f32 [B, in] (row stride % 4 == 0; with ids int32 / int64 [B]: x is an embedding table [V, in]
 * and input row b is x[ids[b]], zeros outside [0, V) -- the Embedding gather of src/models.py:85,89 fused into the tower's
 * first layer) OR the split planes x_hi, x_lo fp16 [B, in] a previous layer wrote.  Its output is y f32 [B, out] (may be a
 * column slice of a wider buffer; 32-byte aligned, row stride % 8 == 0) and / or the split planes y_hi, y_lo fp16 [B, out]
 * (32-byte aligned, row stride % 16 == 0) for the next layer -- same bytes as fp32, but the next layer then loads its
 * operands by TMA instead of converting them.  bias f32 [out] or NULL.  Unused pointers are NULL.
 */
typedef struct {
  const DLTensor* x;
  const DLTensor* ids;
  const DLTensor* x_hi;
  const DLTensor* x_lo;
  const DLTensor* wt_hi;
  const DLTensor* wt_lo;
  const DLTensor* bias;
  DLTensor* y;
  DLTensor* y_hi;
  DLTensor* y_lo;
} rs_x3_layer_t;
int rs_dense_x3_prepare(const DLTensor* kernel, DLTensor* wt_hi, DLTensor* wt_lo, void* stream);
int rs_dense_x3_supported(int64_t in, int64_t out, int act);
int rs_dense_forward_x3(int n_layers, const rs_x3_layer_t* layers, int act, void* stream);
/*
 * K1+K2 / K3 fused (bf16 tensor mode): a whole Dense chain in one tcgen05 kernel, activations kept in tensor
 * memory between layers.  n_layers = 4: tower chain ids -> Embedding -> 128-256-128-64-128 (tables[p] f32
 * [V,128], ids[p]); n_layers = 2: deep chain 256-256-128 on x[p] bf16 [B,256].  n_prob = 1 or 2 problems of the
 * same shape in one launch (the two towers).  wts[p*n_layers+l] bf16 [out,in]; biases likewise (f32, may be NULL);
 * acts[l] in {RS_ACT_NONE, RS_ACT_RELU}; outputs of the last layer: y_f32[p] and/or y_bf16[p], [B, out], strided.
 */
int rs_chain_forward_bf16(int n_prob, int n_layers, const DLTensor* const* wts, const DLTensor* const* biases,
                          const int* acts, const DLTensor* const* tables, const DLTensor* const* ids,
                          const DLTensor* const* x, DLTensor* const* y_f32, DLTensor* const* y_bf16, void* stream);
/*
 * MultiTaskModel.call (src/models.py:125-131) in bf16 tensor mode behind one handle and one call per batch:
 * both towers (gather fused) -> deep net + DCN cross + heads (two kernel launches).  The handle COPIES the kernels
 * (bf16 [out,in]; towers 128-256-128-64-128, index s*4+l with s = 0 user / 1 item; deep net 256-256-128), biases (may
 * be NULL), cross_w / cross_b f32 [L,256], head_w f32 [2,384] (row 0 rating, row 1 ctr), head_b f32 [2]; it BORROWS
 * the two f32 embedding tables [V,128], which must outlive it, and for tables of up to 2^20 rows also keeps a bf16 copy
 * (what the tower kernel gathers from): create a new handle after the tables or weights change.
 * forward: user_rows/item_rows int32|int64 [B] -> combined f32 [B,256] (user embedding | item embedding),
 * rating f32 [B,1], ctr f32 [B,1].
 */
typedef struct rs_ranker rs_ranker_t;
int rs_ranker_create(const DLTensor* user_table, const DLTensor* item_table, const DLTensor* const* tower_wt,
                     const DLTensor* const* tower_b, const DLTensor* const* deep_wt, const DLTensor* const* deep_b,
                     const DLTensor* cross_w, const DLTensor* cross_b, const DLTensor* head_w, const DLTensor* head_b,
                     rs_ranker_t** out);
int rs_ranker_free(rs_ranker_t* h);
int rs_ranker_forward(rs_ranker_t* h, const DLTensor* user_rows, const DLTensor* item_rows, DLTensor* combined,
                      DLTensor* rating, DLTensor* ctr, void* stream);
/* the same forward from already-looked-up embedding rows f32 [B,128] (what a sharded-table lookup returns) */
int rs_ranker_forward_rows(rs_ranker_t* h, const DLTensor* user_emb, const DLTensor* item_emb, DLTensor* combined,
                           DLTensor* rating, DLTensor* ctr, void* stream);
/* K1 (bf16 tensor mode). Embedding rows gathered from the fp32 table straight into bf16 activations. */
int rs_gather_bf16(const DLTensor* table, const DLTensor* ids, DLTensor* out, void* stream);

/*
 * K3a. DCN-v1 cross stack (src/models.py:38-44):
 *   for i < n_layers:  s = x_l . w_i  (per row);  x_{l+1} = x0 * s + b_i + x_l
 *   x0: f32 [B, D]; cross_w: f32 [n_layers, D]; cross_b: f32 [n_layers, D];
 *   out: f32 [B, D] (may be a column slice of the [B, D + deep] concat, src/models.py:50)
 */
int rs_dcn_cross_forward(const DLTensor* x0, const DLTensor* cross_w, const DLTensor* cross_b,
                         DLTensor* out, void* stream);

/*
 * K3a' (EXTENSION -- named by the north-star, absent from the reference, so there is no reference oracle for it; the
 * restatement it is tested against is oracle/extensions.py).  DCN-v2 matrix cross layer x_{l+1} = x0 * (x_l W_l + b_l) + x_l:
 * the d x d product is a Dense layer (rs_dense_forward, or rs_dense_forward_x3 on the tensor cores at ranking batches);
 * this entry is the elementwise rest, out = x0 * u + xl (one fma per element).  All f32 [B, D], rows may be strided.
 */
int rs_cross_v2_combine(const DLTensor* x0, const DLTensor* u, const DLTensor* xl, DLTensor* out, void* stream);

/*
 * K3b. The two ranking heads on the DCN output (src/models.py:119-120,131):
 *   rating = z . w_r + b_r ;  ctr = sigmoid(z . w_c + b_c)
 *   z: f32 [B, W]; head_w: f32 [2, W] (row 0 = rating, row 1 = ctr); head_b: f32 [2];
 *   rating, ctr: f32 [B, 1]
 */
int rs_heads_forward(const DLTensor* z, const DLTensor* head_w, const DLTensor* head_b,
                     DLTensor* rating, DLTensor* ctr, void* stream);

/*
 * K3a + K3b in one pass, without materialising the [B, 384] concat of src/models.py:50: the cross stack on x0 f32 [B, 256]
 * and both heads on (cross output | deep f32 [B, 128]); cross_w / cross_b f32 [L <= 4, 256], head_w f32 [2, 384], head_b
 * f32 [2]; rating, ctr f32 [B, 1].  fp32 FMA arithmetic in the reference's order.  (What MultiTaskModel.call uses in fp32
 * mode at ranking batches: 96 B read per feature row instead of three passes over the concat.)
 */
int rs_cross_heads_forward(const DLTensor* x0, const DLTensor* deep, const DLTensor* cross_w, const DLTensor* cross_b,
                           const DLTensor* head_w, const DLTensor* head_b, DLTensor* rating, DLTensor* ctr, void* stream);

/*
 * K4. faiss.normalize_L2 (src/trainer.py:240, app/recommendation_service.py:71):
 * in place x_i *= 1/sqrt(sum_j x_ij^2), all-zero rows untouched.  x: f32 [n, d],
 * device or host.
 */
int rs_normalize_l2(DLTensor* x, void* stream);

/* K5. faiss.IndexFlatIP (src/trainer.py:241-242, app/recommendation_service.py:47-50,72). */
typedef struct rs_index rs_index_t;
int rs_index_create(int d, rs_index_t** out);          /* faiss.IndexFlatIP(d)                 */
int rs_index_free(rs_index_t* h);
int rs_index_add(rs_index_t* h, const DLTensor* x, void* stream);   /* index.add(x); copies x   */
int rs_index_ntotal(const rs_index_t* h, int64_t* out);             /* index.ntotal             */
int rs_index_dim(const rs_index_t* h, int* out);                    /* index.d                  */
/* copy stored rows [i0, i0 + out.shape[0]) into out (f32 [n, d], host or device); faiss reconstruct_n,
 * used by write_index (src/trainer.py:243) */
int rs_index_reconstruct(const rs_index_t* h, int64_t i0, DLTensor* out, void* stream);
/*
 * index.search(q, k) -> (D, I): for every query the k largest q.x, score descending then
 * row ascending; rows short of k are padded with (D = -FLT_MAX, I = -1).
 *   q: f32 [nq, d]; D: f32 [nq, k]; I: int64 [nq, k]; all device or all host.
 *   id_base is added to every returned row number (shard offset of a partitioned catalog).
 *   mode: RS_EXACT_FP32 | RS_TENSOR_BF16 | RS_AUTO -- all three return identical results.
 */
int rs_index_search(rs_index_t* h, const DLTensor* q, int k, int64_t id_base, int mode, DLTensor* D,
                    DLTensor* I, void* stream);
/*
 * The same search without any host synchronisation, for callers that have more device work to queue behind it
 * (the sharded catalog's exchange + merge, BASELINE config 3).  Device tensors only.  need_exact: int32 [nq];
 * need_exact[i] = 1 marks a query whose tensor-mode candidate set could not be verified: row i of D / I is
 * then NOT final and the caller must search that query again with mode RS_EXACT_FP32 (rs_index_search does
 * this itself, at the price of a stream synchronise).  Exact mode never sets a flag.
 */
int rs_index_search_deferred(rs_index_t* h, const DLTensor* q, int k, int64_t id_base, int mode,
                             DLTensor* D, DLTensor* I, DLTensor* need_exact, void* stream);

/*
 * Shared-threshold search of ONE SHARD of a catalog partitioned over several indexes / GPUs (BASELINE config 3).
 * The tensor path filters with a per-query threshold guessed from a sample.  Guessed per shard, every shard keeps
 * ~6.6 k candidates per query for k = 100; guessed once from the union of all shards' samples, the shards keep
 * that many TOGETHER, and the price is one small extra exchange.  Protocol, per batch of queries, on every shard:
 *   rs_index_shared_begin   sample pass; top[nq, R] = the R largest sample maxima of each query (any order)
 *   (caller)                all-gather the tops -> tops[shards, nq, R]
 *   rs_index_shared_finish  tau_q = the rank-th largest of the gathered values; filter; exact re-score.  D / I
 *                           receive the shard's exact top-min(candidates, k) (padded with D = -FLT_MAX, I = -1),
 *                           bound[q] = tau_q + eps_q of this shard, need_exact[q] = 1 on candidate-list overflow
 *   (caller)                all-gather D, I, bound, need_exact; rs_topk_merge; rs_topk_verify
 *   rs_topk_verify          need_exact[q] |= the merged list is short of k or its k-th score < max over shards of
 *                           bound[q]: only then could a filtered-out row belong to the top-k, and the caller
 *                           re-searches that query with RS_EXACT_FP32 on every shard.  Otherwise the merged list IS
 *                           the exact top-k, bit-identical to a single-index search.
 * rs_tensor_plan: for an index of `ntotal` rows of dimension d, whether the tensor path applies to (nq, k) and
 * how many rows its sample covers; rs_tensor_threshold_rank: the rank to use for a sample covering the fraction
 * rho of ALL rows (callers pass rank to finish and R >= min(rank, ...) to begin; R = rank is always enough).
 * A session belongs to the stream it was begun on; finish (or rs_tc_session_free) releases it.  Device tensors.
 */
typedef struct rs_tc_session rs_tc_session_t;
int rs_tensor_plan(int d, int64_t ntotal, int64_t nq, int k, int64_t* sampled_rows, int* supported);
int rs_tensor_threshold_rank(int k, double rho, int* rank);
int rs_index_shared_begin(rs_index_t* h, const DLTensor* q, int k, DLTensor* top, rs_tc_session_t** out,
                          void* stream);
int rs_index_shared_finish(rs_tc_session_t* s, const DLTensor* tops, int rank, int64_t id_base, DLTensor* D,
                           DLTensor* I, DLTensor* need_exact, DLTensor* bound, void* stream);
int rs_tc_session_free(rs_tc_session_t* s);
/* D: f32 [nq, k], I: int64 [nq, k] merged lists; bounds: f32 [shards, nq]; flags: int32 [shards, nq] (rows of
 * both may be strided slices of the exchange buffer); need_exact: int32 [nq] out */
int rs_topk_verify(const DLTensor* D, const DLTensor* I, const DLTensor* bounds, const DLTensor* flags,
                   DLTensor* need_exact, void* stream);

/* tensor-mode bookkeeping: queries of the last / all tensor-mode searches that failed the filter's
 * verification and were re-run on the exact scan (results are identical either way) */
int rs_index_fallbacks(const rs_index_t* h, int64_t* last, int64_t* total);
/* validation aid: the raw low-precision (bf16 x bf16 -> f32, tcgen05) scores of the tensor path,
 * out f32 [nq, ntotal]; device tensors; small catalogs only */
int rs_debug_tensor_scores(rs_index_t* h, const DLTensor* q, DLTensor* out, void* stream);

/*
 * K6. tf.linalg.matvec(item_embeddings, user_embedding[0]) (app/recommendation_service.py:91):
 * raw, un-normalised dots.  items: f32 [n, d]; user: f32 [d]; out: f32 [n]; device.
 */
int rs_matvec(const DLTensor* items, const DLTensor* user, DLTensor* out, void* stream);

/*
 * K7. Merge of per-shard top-k lists after the all-gather (new in this build; BASELINE
 * config 3).  Dparts: f32 [nparts, nq, k]; Iparts: int64 [nparts, nq, k] (global ids,
 * -1 = padding; ids must be below 2^32 -- a (score, id) pair is merged as one 64-bit key with a 32-bit id field, so
 * catalogs above 2^32 rows are refused by the sharded entry points); D: f32 [nq, k]; I: int64 [nq, k]; device.
 * Order (score desc, id asc).
 * Each part is a contiguous [nq, k] block; the stride between parts is free, so the parts can be
 * slices of one exchange buffer that carries scores, ids and flags of a shard side by side.
 */
int rs_topk_merge(const DLTensor* Dparts, const DLTensor* Iparts, DLTensor* D, DLTensor* I,
                  void* stream);

/*
 * K7p. The exchange step of the sharded search over NVLink peer memory instead of an NCCL all-gather (new in this
 * build; SURVEY.md section 8e "write partial lists straight into peers' buffers").  One process per GPU.  Every rank
 * creates a region of `world` slots of slot_bytes (multiple of 16) and gets a 64-byte CUDA IPC handle for it; the
 * caller exchanges the handles once (any transport; the Python side uses one torch.distributed all-gather) and
 * passes all of them, rank-major, to rs_peer_connect.  Per step and rank, stream-ordered and graph-capturable:
 *   rs_peer_publish  copy src (uint8 [bytes <= slot_bytes], 16-byte aligned) into slot `rank` of EVERY rank's region
 *                    and release this step's number into every rank's flags; waits first until all peers have
 *                    acknowledged the previous step, so one region is enough and its address never changes
 *   rs_peer_wait     returns (on the stream) once the local region holds this step's message of all `world` ranks;
 *                    rs_peer_local gives the region: message g at data + g * slot_bytes
 *   rs_peer_ack      after the kernels that read the region: tell every rank this step has been consumed
 * All ranks must issue the same sequence.  A peer that never arrives trips a ~10 s device-side timeout (trap) rather
 * than a hang.  world <= 16.
 */
typedef struct rs_peer rs_peer_t;
int rs_peer_create(int rank, int world, int64_t slot_bytes, rs_peer_t** out, void* handle_out /* 64 bytes, host */);
int rs_peer_connect(rs_peer_t* h, const void* handles /* world x 64 bytes, host */);
int rs_peer_local(rs_peer_t* h, void** data, int64_t* bytes);
int rs_peer_publish(rs_peer_t* h, const DLTensor* src, void* stream);
int rs_peer_wait(rs_peer_t* h, void* stream);
int rs_peer_ack(rs_peer_t* h, void* stream);
/* Teardown across ranks: an owner must not free a region a peer still has mapped, and nobody may still be writing
 * into it.  Order: every rank's work finished (caller's barrier) -> rs_peer_disconnect on every rank (unmaps the
 * peers' regions) -> barrier -> rs_peer_free.  rs_peer_free alone is enough for a single rank. */
int rs_peer_disconnect(rs_peer_t* h);
int rs_peer_free(rs_peer_t* h);
/*
 * K7q. Sharded embedding lookups (BASELINE config 4) over the same peer regions: both all-to-alls of a lookup are
 * stores into peer memory with device-known counts (no NCCL, no host synchronisation, exact byte counts), and the
 * owner-side row gather writes straight into the requester's region.  Rows are sharded `row mod G`.
 *   rs_a2a_plan_peer  ids int64 [B] -> counts [G], stage [G, b_max] (local row numbers `row div G` per owner) and
 *                     inv [B] (row of this rank's rows region [G * b_max, dim] where each id's embedding will land).
 *                     dedup != 0: each distinct id is requested once (open-addressing hash table in keys [H] int64 /
 *                     vals [H] int32, H a power of two >= 2B; slot_of [B] int32 scratch); dedup == 0: keys / vals /
 *                     slot_of may be NULL.  Negative ids are clamped to 0 (as rs_a2a_scatter does).
 *   rs_peer_a2a_ids   ids region (slots of >= 16 + 8 * b_max bytes): list for owner g -> slot `rank` of g's region
 *   rs_peer_a2a_rows  on the owner: waits for every requester's list, then table f32 [n_local, dim] rows named by
 *                     the lists -> slot `rank` of that requester's rows region (slots of exactly b_max * dim * 4
 *                     bytes); out-of-range rows are zeros; acknowledges the ids messages itself
 *   rs_peer_a2a_restore  waits for every owner's rows, then out[i] = row of ids[i] out of the local rows region --
 *                     position from the hash table (slot_of / vals, de-duplication; inv may then be NULL in both
 *                     calls) or from inv (slot_of = vals = NULL); acknowledges the rows messages itself
 * No rs_peer_wait / rs_peer_ack calls are needed around these three.
 */
int rs_a2a_plan_peer(const DLTensor* ids, int n_shards, int dedup, DLTensor* keys, DLTensor* vals, DLTensor* slot_of,
                     DLTensor* counts, DLTensor* stage, DLTensor* inv, void* stream);
int rs_peer_a2a_ids(rs_peer_t* ids_region, const DLTensor* stage, const DLTensor* counts, void* stream);
int rs_peer_a2a_rows(rs_peer_t* ids_region, rs_peer_t* rows_region, const DLTensor* table, void* stream);
int rs_peer_a2a_restore(rs_peer_t* rows_region, const DLTensor* ids, const DLTensor* slot_of, const DLTensor* vals,
                        const DLTensor* inv, DLTensor* out, void* stream);

/*
 * K7s. The multi-GPU paths behind one call each, for hosts in any language (new in this build).
 * One process per GPU.  A communicator is this rank's view of the group plus ONE host-side all-gather callback the caller
 * supplies (torch.distributed, MPI, a socket ...): recv[g * bytes .. (g+1) * bytes) = rank g's `bytes` bytes of send, host
 * memory, 0 on success.  The library calls it during set-up and teardown only -- it carries the CUDA IPC handles of the
 * peer-memory regions and serves as the barrier around mapping / unmapping them; every per-step byte moves through this
 * library's kernels over NVLink peer memory (no NCCL, no host synchronisation inside a step).  All ranks must issue the
 * same sequence of rs_comm_* / rs_index_search_sharded / rs_sharded_table_* / rs_embedding_a2a_lookup calls; creation and
 * destruction calls are collective.  world <= 16.
 *
 * rs_index_search_sharded: index.search(q, k) (app/recommendation_service.py:72) over a catalog of n_total rows
 *   partitioned by contiguous ranges -- rank g's `shard` holds rows [n_total*g/world, n_total*(g+1)/world).  q f32 [nq, d]
 *   replicated on every rank; D f32 [nq, k] / I int64 [nq, k] (global row numbers) come out identical on every rank and
 *   bit-identical to one index over the whole catalog.  Device tensors.  The first call for a new (nq, k) builds the
 *   peer regions (collective, not capturable); later calls are pure stream work and can be captured in a CUDA graph.
 *   need_exact != NULL (int32 [nq]): nothing is synchronised; rows of flagged queries are not final -- call again for
 *   those queries with mode RS_EXACT_FP32 (every rank sees the same flags).  need_exact == NULL: the library does that
 *   itself (one stream synchronise).  mode: RS_AUTO / RS_TENSOR_BF16 (shared-threshold tensor path where it applies) or
 *   RS_EXACT_FP32 (fp32 scan of every shard); same results.
 *
 * rs_sharded_table_*: an embedding table sharded `row mod world` (global row r = local row r div world of rank r mod world;
 *   replaces keras.layers.Embedding of src/models.py:71,74 for tables that do not fit one GPU, BASELINE config 4).
 *   local_rows f32 [n_local, dim] is BORROWED and must outlive the handle.  rs_embedding_a2a_lookup: ids int64 [B <=
 *   max_batch] global rows of this rank's part of the batch -> out f32 [B, dim] in batch order; both all-to-alls are stores
 *   into peer memory with device-known counts; dedup != 0 requests each distinct id once per batch.  Stream-ordered,
 *   capturable in a CUDA graph.
 */
typedef int (*rs_allgather_fn)(void* ctx, const void* send, void* recv, int64_t bytes);
typedef struct rs_comm rs_comm_t;
typedef struct rs_sharded_table rs_sharded_table_t;
int rs_comm_init(int rank, int world, rs_allgather_fn allgather, void* ctx, rs_comm_t** out);
int rs_comm_rank(const rs_comm_t* c, int* rank, int* world);
int rs_comm_free(rs_comm_t* c);
int rs_index_search_sharded(rs_index_t* shard, rs_comm_t* c, int64_t n_total, const DLTensor* q, int k, int mode, DLTensor* D,
                            DLTensor* I, DLTensor* need_exact, void* stream);
int rs_sharded_table_create(rs_comm_t* c, const DLTensor* local_rows, int64_t max_batch, int dedup, rs_sharded_table_t** out);
int rs_sharded_table_free(rs_sharded_table_t* t);
int rs_embedding_a2a_lookup(rs_sharded_table_t* t, const DLTensor* ids, DLTensor* out, void* stream);
/* rows this rank requested from each owner in the last lookup queued on `stream` (synchronises it): counts_out int64
 * [world], host.  With de-duplication their sum is the number of distinct ids of the batch. */
int rs_sharded_table_last_counts(rs_sharded_table_t* t, int64_t* counts_out, void* stream);

/* counters for the bench's gpu_launches claim: kernels launched by this library so far */
int64_t rs_launch_count(void);

/*
 * Per-kernel device timing for the bench's roofline line: when enabled, the library brackets the
 * launches of each kernel family with CUDA events on the launching stream.  rs_prof_read waits for the
 * recorded events of one slot and returns their summed duration and count, then clears the slot.
 */
#define RS_PROF_SCAN 0     /* scan_topk_kernel (exact fp32 catalog scan)          */
#define RS_PROF_MERGE 1    /* merge_topk_kernel                                   */
#define RS_PROF_FILTER 2   /* tcgen05 bf16 pre-filter                             */
#define RS_PROF_RESCORE 3  /* candidate select + exact re-score (in a sharded step: + publishing the lists and the wait
                              for every rank's lists in the kernel's last CTA)                                        */
#define RS_PROF_RANK 4     /* fused ranking forward: towers (and unfused chains)  */
#define RS_PROF_SAMPLE 5   /* tcgen05 pre-filter, threshold-sampling pass         */
#define RS_PROF_DEEP 6     /* fused ranking forward: deep net + cross + heads     */
#define RS_PROF_THRESHOLD 7 /* query preparation, per-query threshold kernels (in a sharded step: + the tops exchange) */
#define RS_PROF_EXCHANGE 8 /* separate publish launches of the peer-memory exchange (copy + wait for every rank)  */
#define RS_PROF_TOPS 9     /* sharded step: a shard's list of largest sample maxima (+ publishing it, + the wait)     */
#define RS_PROF_SLOTS 16
int rs_prof_enable(int on);
int rs_prof_read(int slot, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* RECSYS_B200_H_ */
