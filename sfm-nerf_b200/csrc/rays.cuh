// rays.cuh -- K1 ray generation, K2 stratified depths, K3 positional encoding (HBM-bound).
#pragma once
#include "common.cuh"

namespace nerf {

// train.py:68-102.  fp64 like the reference, one rounding to fp32 at the end (train.py:127).
__device__ __forceinline__ void pixel_to_ray(const double* __restrict__ pose, double fx, double fy,
                                             double cx, double cy, int px, int py,
                                             float o[3], float d[3]) {
  const double x = ((double)px - cx) / fx;
  const double y = ((double)py - cy) / fy;
  double v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = pose[4 * i + 0] * x + pose[4 * i + 1] * y + pose[4 * i + 2];
  const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    d[i] = (float)(v[i] / n);
    o[i] = (float)pose[4 * i + 3];
  }
}

__global__ void raygen_image_kernel(const double* __restrict__ poses, int pose_index, double fx,
                                    double fy, double cx, double cy, int width, int64_t ray_begin,
                                    int64_t ray_count, float* __restrict__ rays_o,
                                    float* __restrict__ rays_d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ray_count) return;
  const int64_t id = ray_begin + i;                 // ray id = y*W + x (train.py:735-739)
  float o[3], d[3];
  pixel_to_ray(poses + 16 * (int64_t)pose_index, fx, fy, cx, cy, (int)(id % width), (int)(id / width), o, d);
#pragma unroll
  for (int k = 0; k < 3; ++k) { rays_o[3 * i + k] = o[k]; rays_d[3 * i + k] = d[k]; }
}

__global__ void raygen_batch_kernel(const double* __restrict__ poses, double fx, double fy, double cx,
                                    double cy, int width, int height,
                                    const int32_t* __restrict__ img_idx, const int32_t* __restrict__ px,
                                    const int32_t* __restrict__ py, const float* __restrict__ images,
                                    int64_t n_rays, float* __restrict__ rays_o, float* __restrict__ rays_d,
                                    float* __restrict__ target) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  const int k = img_idx[i], x = px[i], y = py[i];
  float o[3], d[3];
  pixel_to_ray(poses + 16 * (int64_t)k, fx, fy, cx, cy, x, y, o, d);
#pragma unroll
  for (int c = 0; c < 3; ++c) { rays_o[3 * i + c] = o[c]; rays_d[3 * i + c] = d[c]; }
  if (images != nullptr && target != nullptr) {     // train.py:123 rgb = images[k][y, x]
    const float* p = images + (((int64_t)k * height + y) * width + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) target[3 * i + c] = __ldg(p + c);
  }
}

__global__ void draw_pixels_kernel(uint64_t seed, uint64_t offset, int n_images, int width, int height,
                                   int64_t n_rays, int32_t* __restrict__ img_idx,
                                   int32_t* __restrict__ px, int32_t* __restrict__ py) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  uint32_t r[4];
  philox4x32_10(seed, offset + (uint64_t)i, r);
  img_idx[i] = (int32_t)(((uint64_t)r[0] * (uint32_t)n_images) >> 32);
  px[i] = (int32_t)(((uint64_t)r[1] * (uint32_t)width) >> 32);
  py[i] = (int32_t)(((uint64_t)r[2] * (uint32_t)height) >> 32);
}

// train.py:205-217.  torch.linspace (CPU) = start + step*i below the midpoint and a single-
// rounding end - step*(S-1-i) above it; every later op rounds separately like eager torch.
__device__ __forceinline__ float linspace01(int i, int S) {
  if (S == 1) return 0.0f;
  const float step = __fdiv_rn(1.0f, (float)(S - 1));
  return (i < S / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(S - 1 - i), 1.0f);
}

__device__ __forceinline__ float stratified_t(int s, int S, float u) {
  const float t = linspace01(s, S);
  const float t_prev = linspace01(s > 0 ? s - 1 : 0, S);
  const float t_next = linspace01(s < S - 1 ? s + 1 : S - 1, S);
  const float lower = (s == 0) ? t : __fmul_rn(0.5f, __fadd_rn(t, t_prev));
  const float upper = (s == S - 1) ? t : __fmul_rn(0.5f, __fadd_rn(t_next, t));
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
}

// Four consecutive depths per thread: one Philox call yields the four uniforms of elements 4k .. 4k+3 (the mapping the
// one-element form used: u(i) = word i & 3 of philox(seed, offset + (i >> 2))), one 16-byte store.  Same roundings as before.
__device__ __forceinline__ void sample_z_quad(int64_t q, int64_t total, int S, float near, float far, int perturb,
                                              const float* __restrict__ t_rand, uint64_t seed, uint64_t offset,
                                              float* __restrict__ z) {
  const int64_t i0 = 4 * q;
  if (i0 >= total) return;
  uint32_t r[4] = {0u, 0u, 0u, 0u};
  if (perturb && t_rand == nullptr) philox4x32_10(seed, offset + (uint64_t)q, r);
  const int s0 = (int)(i0 % S);
  const float range = __fsub_rn(far, near);
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int s = (s0 + k) % S;
    float t;
    if (!perturb) {
      t = linspace01(s, S);
    } else {
      const float u = (t_rand != nullptr) ? ((i0 + k < total) ? t_rand[i0 + k] : 0.0f) : u32_to_unit_float(r[k]);
      t = stratified_t(s, S, u);
    }
    v[k] = __fadd_rn(near, __fmul_rn(range, t));   // train.py:217
  }
  if (i0 + 3 < total && (reinterpret_cast<uintptr_t>(z) & 15) == 0) {
    reinterpret_cast<float4*>(z)[q] = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i0 + k < total) z[i0 + k] = v[k];
  }
}

__global__ void sample_z_kernel(int64_t total, int S, float near, float far, int perturb,
                                const float* __restrict__ t_rand, uint64_t seed, uint64_t offset,
                                float* __restrict__ z) {
  sample_z_quad((int64_t)blockIdx.x * blockDim.x + threadIdx.x, total, S, near, far, perturb, t_rand, seed, offset, z);
}

// generateBatch (train.py:104-127) + the depths of render (train.py:205-217) for one train step in ONE launch: the first
// `ray_blocks` blocks draw (image, x, y) per ray (draw_pixels_kernel), generate the ray and gather its target colour
// (raygen_batch_kernel); the remaining blocks write z (sample_z_kernel).  Bit-identical to the three separate kernels.
// With `batch9` != nullptr the rays are not drawn: batch9 is a generateBatch() result [n_rays, 9] = o | d | rgb
// (train.py:127, e.g. just uploaded from the host) and the ray blocks only split it into the three [n_rays, 3] arrays.
// With `iter_ptr` != nullptr the RNG counters come from the device-side iteration counter (graph-replayable step):
// batch index c = iteration * world + rank, ray_offset = c * n_rays, z_offset = c * ceil(n_rays * S / 4).
__global__ void train_batch_kernel(const unsigned long long* __restrict__ iter_ptr, unsigned world, unsigned rank,
                                   const float* __restrict__ batch9, const double* __restrict__ poses, double fx, double fy, double cx, double cy, int width, int height,
                                   int n_images, const float* __restrict__ images, uint64_t seed, uint64_t ray_offset, int64_t n_rays,
                                   float* __restrict__ rays_o, float* __restrict__ rays_d, float* __restrict__ target, int S, float near,
                                   float far, int perturb, uint64_t z_seed, uint64_t z_offset, float* __restrict__ z,
                                   unsigned ray_blocks) {
  if (iter_ptr != nullptr) {
    const uint64_t c = (uint64_t)(*iter_ptr) * world + rank;
    ray_offset = c * (uint64_t)n_rays;
    z_offset = c * (uint64_t)((n_rays * S + 3) / 4);
  }
  if (blockIdx.x < ray_blocks) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    if (batch9 != nullptr) {
      const float* b = batch9 + 9 * i;
#pragma unroll
      for (int c = 0; c < 3; ++c) { rays_o[3 * i + c] = __ldg(b + c); rays_d[3 * i + c] = __ldg(b + 3 + c); target[3 * i + c] = __ldg(b + 6 + c); }
      return;
    }
    uint32_t r[4];
    philox4x32_10(seed, ray_offset + (uint64_t)i, r);
    const int k = (int)(((uint64_t)r[0] * (uint32_t)n_images) >> 32);
    const int x = (int)(((uint64_t)r[1] * (uint32_t)width) >> 32);
    const int y = (int)(((uint64_t)r[2] * (uint32_t)height) >> 32);
    float o[3], d[3];
    pixel_to_ray(poses + 16 * (int64_t)k, fx, fy, cx, cy, x, y, o, d);
    const float* p = images + (((int64_t)k * height + y) * width + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) { rays_o[3 * i + c] = o[c]; rays_d[3 * i + c] = d[c]; target[3 * i + c] = __ldg(p + c); }
  } else {
    sample_z_quad((int64_t)(blockIdx.x - ray_blocks) * blockDim.x + threadIdx.x, n_rays * S, S, near, far, perturb, nullptr, z_seed,
                  z_offset, z);
  }
}

// NeRFModel.py:142-156.  One thread per output element (coalesced stores).
__global__ void encode_kernel(const float* __restrict__ x, int64_t n, int L, float* __restrict__ out) {
  const int C = 3 * (2 * L + 1);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * C) return;
  const int64_t row = i / C;
  const int ch = (int)(i % C);
  if (ch < 3) { out[i] = x[3 * row + ch]; return; }
  const int level = (ch - 3) / 6, rem = (ch - 3) % 6, coord = rem % 3;
  const float xv = x[3 * row + coord];
  float r;
  if (level < 31 && fabsf(xv) < 1.0e6f) {
    r = pe_reduced(pe_turns(xv), level);
    float s, c;
    sincospif(2.0f * r, &s, &c);
    out[i] = (rem < 3) ? s : c;
  } else {                                     // out-of-domain inputs: plain libm path
    const float a = __fmul_rn(__fmul_rn(3.14159274101257324f, exp2f((float)level)), xv);
    out[i] = (rem < 3) ? sinf(a) : cosf(a);
  }
}

}  // namespace nerf
