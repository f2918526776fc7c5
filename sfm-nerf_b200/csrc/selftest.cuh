// selftest.cuh -- device self tests of the hand-built UMMA descriptors (tc.cuh) on a tiny GEMM, so
// a descriptor-encoding mistake is reported as such instead of as a parity failure deep in K4.
//   variant 0: K-major operands  (forward / dgrad form)   D[128x64] = A[128x32] . B[64x32]^T
//   variant 1: MN-major operands (wgrad form)             D[128x64] = sum_{r<64} dY[r][m] X[r][n]
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace nerf {

__device__ __forceinline__ float st_val(int i, int j, int salt) {      // small integers: exact in bf16
  return (float)(((i * 7 + j * 13 + salt * 5) % 9) - 4);
}

__global__ void __launch_bounds__(128) selftest_umma_kernel(int variant, float* result) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t s_a = tc::smem_u32(smem), s_b = s_a + 16384;
  const int t = threadIdx.x, warp = t >> 5;
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 16384);
  if (variant == 0) {
    // A[m][k]: [kcore = k/8][m (128)][8];  B[n][k]: [kstep = k/16][kc][n (64)][8]
    for (int i = t; i < 128 * 32; i += 128) { const int m = i / 32, k = i % 32; A[(k / 8) * 1024 + m * 8 + (k % 8)] = __float2bfloat16(st_val(m, k, 1)); }
    for (int i = t; i < 64 * 32; i += 128) { const int n = i / 32, k = i % 32; B[(k / 8) * 512 + n * 8 + (k % 8)] = __float2bfloat16(st_val(n, k, 2)); }
  } else {
    // dY[r][m]: [kcore = m/8 (16)][r (64)][8];  X[r][n]: [kcore = n/8 (8)][r (64)][8]
    for (int i = t; i < 64 * 128; i += 128) { const int r = i / 128, m = i % 128; A[(m / 8) * 512 + r * 8 + (m % 8)] = __float2bfloat16(st_val(r, m, 3)); }
    for (int i = t; i < 64 * 64; i += 128) { const int r = i / 64, n = i % 64; B[(n / 8) * 512 + r * 8 + (n % 8)] = __float2bfloat16(st_val(r, n, 4)); }
  }
  if (t == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_slot), 64);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    if (variant == 0) {
      const uint32_t idesc = tc::umma_idesc_bf16(128, 64, 0, 0);
      for (int ks = 0; ks < 2; ++ks)
        tc::umma_bf16(tmem, tc::umma_smem_desc(s_a + ks * 4096, 2048, 128), tc::umma_smem_desc(s_b + ks * 2048, 1024, 128), idesc, ks > 0);
    } else {
      const uint32_t idesc = tc::umma_idesc_bf16(128, 64, 1, 1);
      for (int ks = 0; ks < 4; ++ks)
        tc::umma_bf16(tmem, tc::umma_smem_desc(s_a + ks * 256, 128, 1024), tc::umma_smem_desc(s_b + ks * 256, 128, 1024), idesc, ks > 0);
    }
    tc::umma_commit(tc::smem_u32(&bar));
  }
  tc::mbar_wait(tc::smem_u32(&bar), 0);
  tc::tc_fence_after();
  float max_err = 0.0f, checksum = 0.0f;
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
    tc::tmem_wait_ld();
    for (int j = 0; j < 32; ++j) {
      const int n = ch * 32 + j;
      float ref = 0.0f;
      if (variant == 0) for (int k = 0; k < 32; ++k) ref += st_val(t, k, 1) * st_val(n, k, 2);
      else for (int r = 0; r < 64; ++r) ref += st_val(r, t, 3) * st_val(r, n, 4);
      float e = fabsf(ref - __uint_as_float(v[j]));
      if (!(e < 1e30f)) e = 1e30f;                 // NaN/Inf must not hide behind fmaxf
      max_err = fmaxf(max_err, e);
      checksum += fabsf(ref);
    }
  }
  atomicMax(reinterpret_cast<int*>(result), __float_as_int(max_err));   // non-negative floats order like ints
  atomicAdd(result + 1, checksum);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

// variant 2: CTA-pair form (cta_group::2), K-major operands: D[256x64] = A[256x32] . B[64x32]^T with rows
// 128r..128r+127 of A and B rows 32r..32r+31 in CTA r; the leader issues, the commit is multicast to both.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) selftest_umma_pair_kernel(float* result) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t s_a = tc::smem_u32(smem), s_b = s_a + 16384;
  const int t = threadIdx.x, warp = t >> 5;
  const int rank = (int)tc::cluster_ctarank();
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 16384);
  for (int i = t; i < 128 * 32; i += 128) { const int m = i / 32, k = i % 32; A[(k / 8) * 1024 + m * 8 + (k % 8)] = __float2bfloat16(st_val(rank * 128 + m, k, 1)); }
  for (int i = t; i < 32 * 32; i += 128) { const int n = i / 32, k = i % 32; B[(k / 8) * 256 + n * 8 + (k % 8)] = __float2bfloat16(st_val(rank * 32 + n, k, 2)); }
  if (t == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::fence_mbar_init(); tmem_slot = 0xdeadbeefu; }
  __syncthreads();
  if (warp == 0) tc::tmem_alloc_pair(tc::smem_u32(&tmem_slot), 64);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (t == 0) result[2 + rank] = __uint_as_float(tmem);        // diagnostics: the address each CTA of the pair received
  if (rank == 0 && t == 0) {
    const uint32_t idesc = tc::umma_idesc_bf16(256, 64, 0, 0);
    for (int ks = 0; ks < 2; ++ks)
      tc::umma_bf16_pair(tmem, tc::umma_smem_desc(s_a + ks * 4096, 2048, 128), tc::umma_smem_desc(s_b + ks * 1024, 512, 128), idesc, ks > 0);
    tc::umma_commit_pair(tc::smem_u32(&bar), 3);
  }
  tc::mbar_wait(tc::smem_u32(&bar), 0);
  tc::tc_fence_after();
  float max_err = 0.0f, checksum = 0.0f;
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
    tc::tmem_wait_ld();
    for (int j = 0; j < 32; ++j) {
      const int n = ch * 32 + j;
      float ref = 0.0f;
      for (int k = 0; k < 32; ++k) ref += st_val(rank * 128 + t, k, 1) * st_val(n, k, 2);
      float e = fabsf(ref - __uint_as_float(v[j]));
      if (!(e < 1e30f)) e = 1e30f;
      max_err = fmaxf(max_err, e);
      checksum += fabsf(ref);
    }
  }
  atomicMax(reinterpret_cast<int*>(result), __float_as_int(max_err));
  atomicAdd(result + 1, checksum);
  tc::tc_fence_before();
  tc::cluster_sync();
  if (warp == 0) tc::tmem_dealloc_pair(tmem, 64);
}

#ifdef NERF_DIAGNOSTICS
// Microbenchmark (diagnostics): issue rate of 128x256x16 bf16 UMMAs on one SM.
//   variant 0: one thread, one accumulator          1: one thread, two accumulators alternating
//   variant 2: two threads (warps 0 and 1), one accumulator each
//   variant 3: one thread, N=128 MMAs into two accumulators alternating (same work per pair as one N=256)
//   variant 4: two threads, SAME accumulator (accumulate only)
// result[0] = cycles from first issue to the completion of all `n_mma` MMAs, result[1] = n_mma.
__global__ void __launch_bounds__(128) umma_rate_kernel(int variant, int n_mma, long long* result) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t s_a = tc::smem_u32(smem), s_b = s_a + 65536;     // A: 128 rows x 256 ch (64 KB); B: 16 k-steps x 8 KB
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < (65536 + 131072) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 7);
  if (t == 0) { tc::mbar_init(tc::smem_u32(&bar[0]), 1); tc::mbar_init(tc::smem_u32(&bar[1]), 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_slot), 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const bool two_threads = (variant == 2 || variant == 4);
  long long t0 = 0;
  if ((t == 0) || (two_threads && t == 32)) {
    const int me = t >> 5;
    const int n = (variant == 3) ? 128 : 256;
    const uint32_t idesc = tc::umma_idesc_bf16(128, n, 0, 0);
    const int mine = two_threads ? n_mma / 2 : n_mma;
    t0 = clock64();
    for (int i = 0; i < mine; ++i) {
      const int ks = i & 15;
      uint32_t d = tmem;
      if (variant == 1) d += (i & 1) * 256;
      if (variant == 2) d += me * 256;
      if (variant == 3) d += (i & 1) * 128;
      tc::umma_bf16(d, tc::umma_smem_desc(s_a + ks * 4096, 2048, 128), tc::umma_smem_desc(s_b + ks * 8192, n * 16, 128), idesc,
                    (variant == 4) ? 1u : (i > 1 ? 1u : 0u));
    }
    tc::umma_commit(tc::smem_u32(&bar[me]));
    tc::mbar_wait(tc::smem_u32(&bar[me]), 0);
    const long long t1 = clock64();
    atomicMax(reinterpret_cast<unsigned long long*>(result), (unsigned long long)(t1 - t0));
    result[1] = n_mma;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int F>
__device__ __forceinline__ long long pair_stream_loop(int n_mma, uint32_t tmem, uint32_t s_a, uint32_t s_b, uint64_t* bar, uint32_t done_bar) {
  const uint32_t idesc = tc::umma_idesc_bf16(256, 256, 0, 0);
  const uint32_t commit_bar = tc::smem_u32(&bar[2]), final_bar = tc::smem_u32(&bar[0]);
  const long long t0 = clock64();
  for (int i = 0; i < n_mma; ++i) {
    const int ks = i & 15;
    if ((F & 1) && !(F & 32)) tc::mbar_wait(done_bar, 0);
    if (F & 8) tc::tc_fence_after();
    if (F & 16) {                                    // poll a plain shared-memory word instead of an mbarrier
      uint32_t v;
      do { asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(done_bar + 8) : "memory"); } while (v > 0x7fffffffu);
    }
    if (F & 32) {                                    // order: MMA, mbarrier wait, commit
      if (tc::elect_one())
        tc::umma_bf16_pair(tmem, tc::umma_smem_desc(s_a + ks * 4096, 2048, 128), tc::umma_smem_desc(s_b + ks * 4096, 2048, 128), idesc, i > 1 ? 1u : 0u);
      __syncwarp();
      tc::mbar_wait(done_bar, 0);
      if (tc::elect_one()) tc::umma_commit_pair(commit_bar, 3);
    } else if (tc::elect_one()) {
      tc::umma_bf16_pair(tmem, tc::umma_smem_desc(s_a + ks * 4096, 2048, 128), tc::umma_smem_desc(s_b + ks * 4096, 2048, 128), idesc, i > 1 ? 1u : 0u);
      tc::umma_commit_pair(commit_bar, 3);
    }
    __syncwarp();
  }
  if (tc::elect_one()) tc::umma_commit_pair(final_bar, 3);
  __syncwarp();
  tc::mbar_wait(final_bar, 0);
  return clock64() - t0;
}

// Same for CTA-pair MMAs (cta_group::2, 256x256x16; each CTA holds 128 rows of A and 128 rows of B):
//   variant 0: one thread, one accumulator   1: one thread, two accumulators alternating
//   variant 2: two threads (warps 0 and 1 of the leader), one accumulator each
//   variant 3: like 0 with a multicast commit after every MMA      4: like 2 with a commit after every MMA
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) umma_pair_rate_kernel(int variant, int n_mma, long long* result) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[5];     // bar[4] doubles as a plain flag word (zero)
  __shared__ uint32_t tmem_slot;
  const uint32_t s_a = tc::smem_u32(smem), s_b = s_a + 65536;     // A: 128 rows x 256 ch (64 KB); B: 16 k-steps x 4 KB
  if (threadIdx.x == 0) bar[4] = 0;
  const int t = threadIdx.x, warp = t >> 5;
  const int rank = (int)tc::cluster_ctarank();
  for (int i = t; i < (65536 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 7);
  if (t == 0) { for (int i = 0; i < 4; ++i) tc::mbar_init(tc::smem_u32(&bar[i]), 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc_pair(tc::smem_u32(&tmem_slot), 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (variant >= 5) {
    // variants 5 + flags: one stream issued the way the fused MLP kernel does it (whole warp runs the loop, elect.sync
    // lane issues MMA + multicast commit, __syncwarp per trip), plus bit0: mbarrier wait on a completed barrier per
    // trip; bit2: warps 1-3 of both CTAs poll a barrier that completes at the end; bit3: tcgen05.fence per trip
    const int flags = variant - 5;
    if (rank == 0 && warp == 0) {
      const uint32_t done_bar = tc::smem_u32(&bar[3]);
      if (t == 0) tc::mbar_arrive(done_bar);                 // phase 0 of bar[3] is complete: a wait on parity 0 succeeds at once
      __syncwarp();
      const long long cyc = (flags & 16) ? pair_stream_loop<16>(n_mma, tmem, s_a, s_b, bar, done_bar) : (flags & 32) ? pair_stream_loop<33>(n_mma, tmem, s_a, s_b, bar, done_bar) : (flags & 8) ? ((flags & 1) ? pair_stream_loop<9>(n_mma, tmem, s_a, s_b, bar, done_bar) : pair_stream_loop<8>(n_mma, tmem, s_a, s_b, bar, done_bar))
                                        : ((flags & 1) ? pair_stream_loop<1>(n_mma, tmem, s_a, s_b, bar, done_bar) : pair_stream_loop<0>(n_mma, tmem, s_a, s_b, bar, done_bar));
      if (t == 0) { atomicMax(reinterpret_cast<unsigned long long*>(result), (unsigned long long)cyc); result[1] = n_mma; }
    } else if ((flags & 4) || (rank == 1 && t == 0)) {
      tc::mbar_wait(tc::smem_u32(&bar[0]), 0);               // pollers: spin until the final multicast commit lands
    }
  } else {
  const bool two_threads = (variant == 2 || variant == 4);
  const bool commit_each = (variant == 3 || variant == 4);
  if (rank == 0 && ((t == 0) || (two_threads && t == 32))) {
    const int me = t >> 5;
    const uint32_t idesc = tc::umma_idesc_bf16(256, 256, 0, 0);
    const int mine = two_threads ? n_mma / 2 : n_mma;
    const long long t0 = clock64();
    for (int i = 0; i < mine; ++i) {
      const int ks = i & 15;
      uint32_t d = tmem;
      if (variant == 1) d += (i & 1) * 256;
      if (two_threads) d += me * 256;
      tc::umma_bf16_pair(d, tc::umma_smem_desc(s_a + ks * 4096, 2048, 128), tc::umma_smem_desc(s_b + ks * 4096, 2048, 128), idesc, i > 1 ? 1u : 0u);
      if (commit_each) tc::umma_commit_pair(tc::smem_u32(&bar[2 + me]), 3);   // nobody waits on these (arrivals just accumulate phases)
    }
    tc::umma_commit_pair(tc::smem_u32(&bar[me]), 3);
    tc::mbar_wait(tc::smem_u32(&bar[me]), 0);
    const long long t1 = clock64();
    atomicMax(reinterpret_cast<unsigned long long*>(result), (unsigned long long)(t1 - t0));
    result[1] = n_mma;
  }
  if (rank == 1 && (t == 0 || (two_threads && t == 32))) tc::mbar_wait(tc::smem_u32(&bar[t >> 5]), 0);
  }
  tc::tc_fence_before();
  tc::cluster_sync();
  if (warp == 0) tc::tmem_dealloc_pair(tmem, 512);
}

// Store-rate microbenchmark: one CTA per SM, `warps` warps; every warp instruction stores 512 contiguous bytes (16 B per lane,
// like the stash / dstash stores of the MLP kernels) of a [row][16 B] piece read from shared memory.  mode 0: st.global.cs,
// 1: st.global (default policy), 2: st.global.cs of a register value (no shared-memory read).  Each CTA streams through its own
// region of `buf`.  result[0] = max cycles over CTAs, result[1] = bytes per CTA.
__global__ void store_rate_kernel(uint4* __restrict__ buf, int iters, int mode, long long* __restrict__ result) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4* sm = reinterpret_cast<uint4*>(smem);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
  __syncthreads();
  uint4* base = buf + (size_t)blockIdx.x * (size_t)iters * warps * 32;
  const long long t0 = clock64();
  uint4 v = sm[threadIdx.x & 4095];
  // modes 0-2: every CTA streams through its own contiguous region.  mode 3: 64 KB blocks interleaved over the CTAs (block j of
  // CTA b at (j * gridDim + b) * 64 KB: what a [layer][tile] stash layout would give); mode 4: 2 KB pieces interleaved over the CTAs.
  const int per_block = (65536 / 512) / warps;          // iterations per 64 KB block
  for (int it = 0; it < iters; ++it) {
    if (mode != 2) v = sm[(it * 128 + warp * 32 + lane) & 4095];
    uint4* dst = base + ((size_t)it * warps + warp) * 32 + lane;
    if (mode == 3 && per_block > 0) {
      const size_t blk = (size_t)(it / per_block) * gridDim.x + blockIdx.x;
      dst = buf + blk * 4096 + ((size_t)(it % per_block) * warps + warp) * 32 + lane;
    } else if (mode == 4) {
      dst = buf + ((size_t)it * gridDim.x + blockIdx.x) * (size_t)warps * 32 + warp * 32 + lane;
    }
    if (mode == 1) *dst = v; else __stcs(dst, v);
  }
  __syncthreads();
  const long long dt = clock64() - t0;
  if (threadIdx.x == 0) {
    atomicMax(reinterpret_cast<unsigned long long*>(result), (unsigned long long)dt);
    result[1] = (long long)iters * warps * 512;
  }
}
#endif  // NERF_DIAGNOSTICS

}  // namespace nerf
