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
  if (t == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc_pair(tc::smem_u32(&tmem_slot), 64);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
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

}  // namespace nerf
