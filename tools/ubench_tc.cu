// ubench_tc.cu -- hardware questions behind the round-2 redesign of the ranking kernels (tools only, not product code):
//   A. tcgen05.mma issue/execute rate for small N (TS and SS form): cycles per K=16 MMA at N = 16..256
//   B. the cost of one chain hand-off: MMA -> commit -> epilogue warps (tcgen05.ld, convert, tcgen05.st) -> arrive -> MMA
//   C. how fast one SM can push stores out (coalesced STG.128, per-row STG.256, bulk shared->global copies of 2 KB / 16 KB),
//      with 1, 32 and 148 CTAs active
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/ubench_tc.cu -o tools/_bin/ubench_tc
// Output: one JSON object per line on stdout.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../recommendation-system-maang-nvidia-_b200/csrc/tcgen05.cuh"

using namespace rsd;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

// ---------------------------------------------------------------- A: MMA rate
// one thread issues `reps` groups of 8 MMAs (K = 128) into the same accumulator, commits once, waits; operands are
// whatever the (zeroed) shared memory / tensor memory holds -- only the timing matters
template <bool TS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tslot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = tslot;
  if (threadIdx.x == 0) {
    const uint64_t da = umma_desc_k_sw128(smem_u32(smem));              // A (SS): [128][64] bf16 tiles, 16 KB each
    const uint64_t db = umma_desc_k_sw128(smem_u32(smem + 32 * 1024));  // B: [N][64] bf16 tiles
    const uint32_t idesc = umma_idesc_bf16_f32(128, N);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t off = static_cast<uint64_t>(((k >> 2) * (N * 128) + (k & 3) * 32) >> 4);
        if (TS) umma_bf16_ts(tm + 256, tm + k * 8, db + off, idesc, (r | k) ? 1u : 0u);
        else umma_bf16_ss(tm + 256, da + static_cast<uint64_t>(((k >> 2) * 16384 + (k & 3) * 32) >> 4), db + off, idesc, (r | k) ? 1u : 0u);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------- B: chain hand-off
// issuer: wait a_ready -> `nmma` MMAs (TS, N) -> commit acc_ready.  `teams` epilogue teams of 4 warps take the steps in
// rotation (team t handles steps t, t + teams, ...), each with its own barrier pair and its own TMEM columns, so up to
// `teams` steps are in flight: teams = 1 measures the pure serial latency, teams >= 2 the pipelined rate.
// epilogue per step: wait acc -> tcgen05.ld COLS columns -> relu/pack -> tcgen05.st COLS/2 -> fence -> arrive.
template <int COLS>
__global__ void __launch_bounds__(128 + 16 * 32, 1) handoff_kernel(int N, int nmma, int teams, int steps, int spin, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t a_ready[4], acc_ready[4];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int t = 0; t < 4; ++t) {
      mbar_init(&a_ready[t], 4);
      mbar_init(&acc_ready[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tslot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = tslot;
  auto wait = [&](uint64_t* b, uint32_t ph) {
    if (spin) mbar_wait_spin(b, ph);
    else mbar_wait(b, ph);
  };
  if (warp == 1 && lane == 0) {
    const uint64_t db = umma_desc_k_sw128(smem_u32(smem));
    const uint32_t idesc = umma_idesc_bf16_f32(128, N);
    const long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
      const int t = s % teams;
      const uint32_t use = s / teams;
      if (use > 0) {   // the team's previous step has been consumed and its next A operand written
        wait(&a_ready[t], (use - 1) & 1u);
        tc_fence_after_sync();
      }
      for (int k = 0; k < nmma; ++k)
        umma_bf16_ts(tm + 256 + t * 64, tm + t * 32 + (k & 3) * 8, db + static_cast<uint64_t>(((k & 3) * 32) >> 4), idesc, k ? 1u : 0u);
      umma_commit(&acc_ready[t]);
    }
    for (int t = 0; t < teams; ++t) {
      const int n_t = (steps - t + teams - 1) / teams;
      if (n_t > 0) wait(&a_ready[t], (n_t - 1) & 1u);
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp >= 4) {
    const int team = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const uint32_t lsel = static_cast<uint32_t>(quarter * 32) << 16;
    if (team < teams) {
      const int n_t = (steps - team + teams - 1) / teams;
      for (int u = 0; u < n_t; ++u) {
        wait(&acc_ready[team], u & 1u);
        tc_fence_after_sync();
        float v[COLS];
        tmem_ld_cols(tm + 256 + team * 64 + lsel, v);
        uint32_t r[COLS / 2];
#pragma unroll
        for (int e = 0; e < COLS / 2; ++e) r[e] = pack_bf16x2_relu(v[2 * e], v[2 * e + 1]);
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) {
          uint32_t (&rr)[16] = reinterpret_cast<uint32_t (&)[16]>(r[16 * c]);
          tmem_st_32x32_x16(tm + team * 32 + c * 16 + lsel, rr);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[team]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 512);
}


// ---------------------------------------------------------------- B2: the same hand-off with clock64 stamps
// out[0..]: issuer stamps [step][4] = {a_ready seen, MMAs issued, commit issued, -}; epilogue warp of team 0 / team 1
// (lane 0 of the quarter-0 warp) stamps [step][6] = {acc_ready seen, after fence, ld done, packed, st done, arrived}
// variant bit 0: no tcgen05.st (the packed values are dropped); bit 1: SS-form MMA (A from shared memory)
__global__ void __launch_bounds__(128 + 16 * 32, 1) handoff_trace_kernel(int N, int nmma, int teams, int steps, int variant, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t a_ready[4], acc_ready[4];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int t = 0; t < 4; ++t) {
      mbar_init(&a_ready[t], 4);
      mbar_init(&acc_ready[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tslot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = tslot;
  const bool rec = blockIdx.x == 0;
  long long* iss = out;                 // [steps][4]
  long long* ep0 = out + steps * 4;     // [steps][6]
  long long* ep1 = ep0 + steps * 6;
  if (warp == 1 && lane == 0) {
    const uint64_t db = umma_desc_k_sw128(smem_u32(smem));
    const uint64_t da = umma_desc_k_sw128(smem_u32(smem + 32 * 1024));
    const uint32_t idesc = umma_idesc_bf16_f32(128, N);
    for (int s = 0; s < steps; ++s) {
      const int t = s % teams;
      const uint32_t use = s / teams;
      if (use > 0) {
        mbar_wait(&a_ready[t], (use - 1) & 1u);
        tc_fence_after_sync();
      }
      if (rec) iss[s * 4 + 0] = clock64();
      for (int k = 0; k < nmma; ++k) {
        if (variant & 2) umma_bf16_ss(tm + 256 + t * 64, da + static_cast<uint64_t>(((k & 3) * 32) >> 4), db + static_cast<uint64_t>(((k & 3) * 32) >> 4), idesc, k ? 1u : 0u);
        else umma_bf16_ts(tm + 256 + t * 64, tm + t * 32 + (k & 3) * 8, db + static_cast<uint64_t>(((k & 3) * 32) >> 4), idesc, k ? 1u : 0u);
      }
      if (rec) iss[s * 4 + 1] = clock64();
      umma_commit(&acc_ready[t]);
      if (rec) iss[s * 4 + 2] = clock64();
    }
    for (int t = 0; t < teams; ++t) {
      const int n_t = (steps - t + teams - 1) / teams;
      if (n_t > 0) mbar_wait(&a_ready[t], (n_t - 1) & 1u);
    }
  } else if (warp >= 4) {
    const int team = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const uint32_t lsel = static_cast<uint32_t>(quarter * 32) << 16;
    long long* ep = team == 0 ? ep0 : ep1;
    const bool me = rec && team < 2 && quarter == 0 && lane == 0;
    if (team < teams) {
      const int n_t = (steps - team + teams - 1) / teams;
      for (int u = 0; u < n_t; ++u) {
        const int s = u * teams + team;
        mbar_wait(&acc_ready[team], u & 1u);
        if (me) ep[s * 6 + 0] = clock64();
        tc_fence_after_sync();
        if (me) ep[s * 6 + 1] = clock64();
        float v[32];
        tmem_ld_cols(tm + 256 + team * 64 + lsel, v);
        if (me) ep[s * 6 + 2] = clock64();
        uint32_t r[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = pack_bf16x2_relu(v[2 * e], v[2 * e + 1]);
        if (me) ep[s * 6 + 3] = clock64() + (r[3] == 0x12345u ? 1 : 0);
        if (!(variant & 1)) tmem_st_32x32_x16(tm + team * 32 + lsel, r);
        else if (r[5] == 0x7777u) out[0] = 1;
        if (me) ep[s * 6 + 4] = clock64();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[team]);
        if (me) ep[s * 6 + 5] = clock64();
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------- C: store rate of an SM
// every CTA writes `tiles` chunks of 96 KB to its own region.  mode 0: coalesced STG.128 (a warp writes 512 contiguous
// bytes); 1: one 512-byte row per thread, 32 bytes (STG.256) at a time; 2: bulk shared->global copies of 2 KB, one
// issuing lane per warp, two in flight per warp; 3: the same with 16 KB copies issued by 4 warps
__global__ void __launch_bounds__(512, 1) store_rate_kernel(uint8_t* dst, int tiles, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, 1, 2, 3);
  fence_proxy_async_smem();
  __syncthreads();
  uint8_t* base = dst + static_cast<size_t>(blockIdx.x) * tiles * 98304;
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    uint8_t* p = base + static_cast<size_t>(t) * 98304;
    if (mode == 0) {
      for (int i = threadIdx.x; i < 98304 / 16; i += 512) reinterpret_cast<uint4*>(p)[i] = make_uint4(i, t, 2, 3);
    } else if (mode == 1) {
      if (threadIdx.x < 192) {   // 192 rows of 512 bytes
        float* row = reinterpret_cast<float*>(p + threadIdx.x * 512);
        for (int e = 0; e < 128; e += 8) st_global_256(row + e, 1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f, (float)t);
      }
    } else if (mode == 2) {
      if (lane == 0) {
        for (int c = 0; c < 3; ++c) {   // 16 warps x 3 x 2 KB = 96 KB
          bulk_wait_read<1>();
          bulk_store_1d(p + (warp * 3 + c) * 2048, smem + (warp * 3 + c) * 2048, 2048);
          bulk_commit();
        }
      }
    } else {
      if (lane == 0 && warp < 6) {      // 6 x 16 KB
        bulk_wait_read<1>();
        bulk_store_1d(p + warp * 16384, smem + warp * 16384, 16384);
        bulk_commit();
      }
    }
  }
  if (mode >= 2 && lane == 0) bulk_wait_all();
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  long long* d_out;
  CK(cudaMalloc(&d_out, 1024 * sizeof(long long)));
  long long h[1024];
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));


  // ---- B2: stamped hand-off (steady-state steps 40..47)
  if (getenv("UB_TRACE")) {
    CK(cudaFuncSetAttribute(handoff_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int steps = 64;
    long long* d_tr;
    CK(cudaMalloc(&d_tr, steps * 16 * sizeof(long long)));
    std::vector<long long> tr(steps * 16);
    for (int variant : {0, 1, 2})
      for (int teams : {1, 2, 4})
        for (int nmma : {1, 8}) {
          CK(cudaMemset(d_tr, 0, steps * 16 * sizeof(long long)));
          handoff_trace_kernel<<<sms, 640, 64 * 1024>>>(64, nmma, teams, steps, variant, d_tr);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(tr.data(), d_tr, steps * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
          const long long base = tr[40 * 4];
          printf("{\"bench\": \"handoff_trace\", \"variant\": %d, \"teams\": %d, \"mma_per_step\": %d, \"steps\": [", variant, teams, nmma);
          for (int s = 40; s < 48; ++s) {
            const long long* ep = (s % teams == 0) ? &tr[steps * 4 + s * 6] : (s % teams == 1) ? &tr[steps * 10 + s * 6] : nullptr;
            printf("%s{\"s\": %d, \"iss\": [%lld, %lld, %lld]", s > 40 ? ", " : "", s, tr[s * 4] - base, tr[s * 4 + 1] - base, tr[s * 4 + 2] - base);
            if (ep) printf(", \"epi\": [%lld, %lld, %lld, %lld, %lld, %lld]", ep[0] - base, ep[1] - base, ep[2] - base, ep[3] - base, ep[4] - base, ep[5] - base);
            printf("}");
          }
          printf("]}\n");
        }
    return 0;
  }

  // ---- A
  CK(cudaFuncSetAttribute(mma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  CK(cudaFuncSetAttribute(mma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  for (int ts = 1; ts >= 0; --ts)
    for (int N : {16, 32, 64, 128, 256})
      for (int grid : {1, sms}) {
        const int reps = 256;
        for (int it = 0; it < 2; ++it) {
          if (ts) mma_rate_kernel<true><<<grid, 128, 96 * 1024>>>(N, reps, d_out);
          else mma_rate_kernel<false><<<grid, 128, 96 * 1024>>>(N, reps, d_out);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
        printf("{\"bench\": \"mma_rate\", \"form\": \"%s\", \"N\": %d, \"grid\": %d, \"issue_cyc_per_mma\": %.1f, \"exec_cyc_per_mma\": %.1f, "
               "\"floor_128N_over_256\": %.1f}\n",
               ts ? "TS" : "SS", N, grid, h[0] / (reps * 8.0), h[1] / (reps * 8.0), 128.0 * N / 256.0);
      }

  // ---- B
  CK(cudaFuncSetAttribute(handoff_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(handoff_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int cols : {32, 64})
    for (int spin : {0, 1})
      for (int teams : {1, 2, 3, 4})
        for (int nmma : {1, 4, 8}) {
          const int N = 64, steps = 1200;
          for (int it = 0; it < 2; ++it) {
            if (cols == 32) handoff_kernel<32><<<sms, 640, 64 * 1024>>>(N, nmma, teams, steps, spin, d_out);
            else handoff_kernel<64><<<sms, 640, 64 * 1024>>>(N, nmma, teams, steps, spin, d_out);
            CK(cudaDeviceSynchronize());
          }
          CK(cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost));
          printf("{\"bench\": \"handoff\", \"ld_cols\": %d, \"spin\": %d, \"teams\": %d, \"mma_per_step\": %d, \"N\": %d, "
                 "\"cyc_per_step\": %.1f, \"mma_exec_floor\": %.1f}\n",
                 cols, spin, teams, nmma, N, h[0] / (double)steps, nmma * 128.0 * N / 256.0);
        }

  // ---- C
  CK(cudaFuncSetAttribute(store_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  uint8_t* d_dst;
  const int tiles_max = 64;
  CK(cudaMalloc(&d_dst, static_cast<size_t>(sms) * tiles_max * 98304));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int tiles : {7, 64})
    for (int mode = 0; mode < 4; ++mode)
      for (int grid : {1, 32, sms}) {
        float ms = 0.f;
        for (int it = 0; it < 3; ++it) {
          CK(cudaEventRecord(e0));
          store_rate_kernel<<<grid, 512, 96 * 1024>>>(d_dst, tiles, mode, d_out);
          CK(cudaEventRecord(e1));
          CK(cudaDeviceSynchronize());
          CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        CK(cudaMemcpy(h, d_out, grid * 8, cudaMemcpyDeviceToHost));
        double mx = 0, sum = 0;
        for (int i = 0; i < grid; ++i) {
          mx = h[i] > mx ? h[i] : mx;
          sum += h[i];
        }
        const double bytes = (double)tiles * 98304;
        printf("{\"bench\": \"store_rate\", \"mode\": \"%s\", \"tiles_per_cta\": %d, \"grid\": %d, \"bytes_per_clk_per_sm_mean\": %.1f, "
               "\"bytes_per_clk_per_sm_slowest\": %.1f, \"kernel_ms\": %.4f, \"chip_GBps\": %.0f}\n",
               mode == 0 ? "stg128_coalesced" : mode == 1 ? "stg256_row_per_thread" : mode == 2 ? "bulk_2KB" : "bulk_16KB", tiles, grid,
               bytes / (sum / grid), bytes / mx, ms, bytes * grid / (ms * 1e-3) / 1e9);
      }
  return 0;
}
