// common.cuh -- shared helpers for libnerf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/nerf_b200.h"

#define NERF_CUDA_TRY(expr)                     \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

#define NERF_LAUNCH_CHECK()                     \
  do {                                          \
    cudaError_t _e = cudaGetLastError();        \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

static inline bool nerf_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

namespace nerf {

constexpr int kHidden = 256;      // NeRFModel.py: num_channels (fixed-function kernels)
constexpr int kRgbHidden = 128;   // NeRFModel.py:139 rgb_layer1 out_features
constexpr int kTileRows = 128;    // UMMA M

// ---- counter-based RNG: Philox4x32-10 ------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint64_t seed, uint64_t ctr, uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x6e657266u, c3 = 0x62323030u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in [0,1) with 24 random bits, like torch.rand for float32
__host__ __device__ inline float u32_to_unit_float(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

// ---- positional-encoding angle helper (NeRFModel.py:151-156) --------------------------------
// The reference evaluates sin/cos of a_i = fl32(fl32(2^i*pi) * x) == 2^i * fl32(fl32(pi)*x)
// exactly (power-of-two scaling commutes with rounding).  We carry t = a_0 / (2*pi) in fp64,
// scale by 2^i exactly and reduce to r in [-0.5, 0.5] turns, so no large-argument reduction
// is ever needed.
__device__ __forceinline__ double pe_turns(float x) {
  const float a0 = __fmul_rn(3.14159274101257324f, x);   // fl32(pi) * x, rounded like torch
  return (double)a0 * 0.15915494309189533576888;         // / (2*pi)
}
__device__ __forceinline__ float pe_reduced(double turns, int level) {
  double t = turns * (double)(1u << level);
  t -= rint(t);
  return (float)t;                                       // in [-0.5, 0.5]
}

}  // namespace nerf
