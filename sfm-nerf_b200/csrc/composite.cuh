// composite.cuh -- K5 volume compositing forward/backward (train.py:248-269), MSE, Adam.
// One warp per ray; lane l handles samples l, l+32, ... so every load is a coalesced float4 /
// float row segment; the transmittance product is a warp scan with a running carry.
#pragma once
#include "common.cuh"

namespace nerf {

constexpr int kCompWarpsPerBlock = 8;
constexpr int kCompMaxChunks = 8;          // backward keeps per-lane state for S <= 256

__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float warp_incl_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sums of four per-lane values over the warp with 6 shuffles instead of 20: after the xor-16 / xor-8 steps every lane keeps
// one of the four partial sums; lanes 0-7 end with the total of a, 8-15 of b, 16-23 of c, 24-31 of d.
__device__ __forceinline__ float warp_sum4_scatter(float a, float b, float c, float d, int lane) {
  const bool up16 = (lane & 16) != 0;
  float k0 = up16 ? c : a, k1 = up16 ? d : b;
  k0 += __shfl_xor_sync(0xffffffffu, up16 ? a : c, 16);
  k1 += __shfl_xor_sync(0xffffffffu, up16 ? b : d, 16);
  const bool up8 = (lane & 8) != 0;
  float k = up8 ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}

// alpha_i = 1 - exp(-sigma_i * delta_i); delta_{S-1} = 1e-3 (train.py:248-254).  z / z_next are the depths of samples s, s+1.
__device__ __forceinline__ void sample_alpha(float z, float z_next, int s, int S, float sigma, float& delta, float& alpha) {
  // last delta = 1e-3 (train.py:250); with S == 1 the reference's `dists` is EMPTY (its ones_like of an
  // empty slice) and every output is 0 -- reproduced by a zero delta.
  delta = (s < S - 1) ? __fsub_rn(z_next, z) : (S == 1 ? 0.0f : 1e-3f);
  alpha = __fsub_rn(1.0f, expf(-__fmul_rn(sigma, delta)));
}

// One 32-sample chunk of a ray as its warp holds it: lane l = sample c*32 + l (zeros past the end of the ray).  All chunks
// of a ray (or of a 128-sample group) are loaded before any of them is used, so every warp keeps several 512-byte rows in
// flight instead of one load -> scan -> load chain.
struct CompChunk { float4 v; float z, zn; };
__device__ __forceinline__ CompChunk load_chunk(const float4* __restrict__ sr, const float* __restrict__ zr, int s, int S) {
  CompChunk c;
  c.v = make_float4(0.f, 0.f, 0.f, 0.f); c.z = 0.f; c.zn = 0.f;
  if (s < S) {
    c.v = __ldg(sr + s);
    c.z = __ldg(zr + s);
    if (s < S - 1) c.zn = __ldg(zr + s + 1);
  }
  return c;
}

// Forward, any S: one warp per ray, groups of four chunks with the loads first.
__global__ void __launch_bounds__(kCompWarpsPerBlock * 32)
composite_forward_kernel(const float4* __restrict__ sigma_rgb, const float* __restrict__ z,
                         int64_t n_rays, int S, float* __restrict__ rgb, float* __restrict__ depth,
                         float* __restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kCompWarpsPerBlock + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const float4* sr = sigma_rgb + ray * S;
  const float* zr = z + ray * S;
  float carry = 1.0f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
  for (int base = 0; base < S; base += 128) {
    CompChunk ck[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ck[k] = load_chunk(sr, zr, base + 32 * k + lane, S);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (base + 32 * k < S) {                                   // warp-uniform
        const int s = base + 32 * k + lane;
        float om = 1.0f, alpha = 0.0f;
        if (s < S) {
          float delta;
          sample_alpha(ck[k].z, ck[k].zn, s, S, ck[k].v.x, delta, alpha);
          om = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);        // train.py:262
        }
        const float incl = warp_incl_prod(om, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float T = carry * excl;                             // exclusive cumprod (261-263)
        const float w = alpha * T;                                // train.py:266
        ar += w * ck[k].v.y; ag += w * ck[k].v.z; ab += w * ck[k].v.w; ad += w * ck[k].z; aa += w;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
      }
    }
  }
  const float r = warp_sum4_scatter(ar, ag, ab, ad, lane);
  aa = warp_sum(aa);
  if ((lane & 7) == 0) {
    if (lane < 24) rgb[3 * ray + (lane >> 3)] = r;               // train.py:269
    else if (depth) depth[ray] = r;
    if (lane == 0 && acc) acc[ray] = aa;
  }
}

// Forward for S == 32 * NC (64, 128, 192, 256 samples per ray: every chunk is full).  ncu showed the generic kernels
// ISSUE-bound, not memory-bound: 282 warp instructions per 64-sample ray of which ~165 were index arithmetic, bounds predicates
// and branches (smsp issue active 62-80 % at 3.4 TB/s).  With the ray length a compile-time constant the loads are NC
// unguarded 16-byte + 2 NC 4-byte loads from one base pointer and the loop body is straight-line code.
template <int NC>
__global__ void __launch_bounds__(kCompWarpsPerBlock * 32)
composite_forward_exact_kernel(const float4* __restrict__ sigma_rgb, const float* __restrict__ z, int64_t n_rays,
                               float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ acc) {
  constexpr int S = 32 * NC;
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kCompWarpsPerBlock + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const float4* sr = sigma_rgb + ray * S + lane;
  const float* zr = z + ray * S + lane;
  float4 v[NC];
  float zz[NC], zn[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) { v[k] = __ldg(sr + 32 * k); zz[k] = __ldg(zr + 32 * k); }
#pragma unroll
  for (int k = 0; k < NC - 1; ++k) zn[k] = __ldg(zr + 32 * k + 1);
  zn[NC - 1] = __ldg(zr + 32 * (NC - 1) + (lane < 31 ? 1 : 0));   // the last sample has no successor: its delta is 1e-3
  float carry = 1.0f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    float delta = __fsub_rn(zn[k], zz[k]);
    if (k == NC - 1 && lane == 31) delta = 1e-3f;                 // train.py:250
    const float alpha = __fsub_rn(1.0f, expf(-__fmul_rn(v[k].x, delta)));
    const float om = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);   // train.py:262
    const float incl = warp_incl_prod(om, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float w = alpha * (carry * excl);                       // train.py:261-266
    ar += w * v[k].y; ag += w * v[k].z; ab += w * v[k].w; ad += w * zz[k]; aa += w;
    if (k < NC - 1) carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
  const float r = warp_sum4_scatter(ar, ag, ab, ad, lane);
  aa = warp_sum(aa);
  if ((lane & 7) == 0) {
    if (lane < 24) rgb[3 * ray + (lane >> 3)] = r;               // train.py:269
    else if (depth) depth[ray] = r;
    if (lane == 0 && acc) acc[ray] = aa;
  }
}

// Backward: with gc_i = g . c_i and suffix_i = sum_{j>i} w_j gc_j
//   dL/dc_i = w_i g;  dL/dalpha_i = T_i gc_i - suffix_i / om_i;  dL/dsigma_i = delta_i (1-alpha_i) dL/dalpha_i
// NC = chunks per ray this instance is unrolled for (S <= 32 * NC).
struct CompState { float T, w, f, om; };
template <int NC>
__device__ __forceinline__ void composite_ray_forward(const CompChunk (&ck)[NC], int n_chunks, int lane, int S, CompState (&st)[NC],
                                                      float& ar, float& ag, float& ab) {
  float carry = 1.0f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < n_chunks) {
      const int s = c * 32 + lane;
      float om = 1.0f, alpha = 0.0f, delta = 0.0f;
      if (s < S) {
        sample_alpha(ck[c].z, ck[c].zn, s, S, ck[c].v.x, delta, alpha);
        om = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      }
      const float incl = warp_incl_prod(om, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      st[c].T = carry * excl;
      st[c].w = alpha * st[c].T;
      st[c].f = delta * (1.0f - alpha);
      st[c].om = om;
      ar += st[c].w * ck[c].v.y; ag += st[c].w * ck[c].v.z; ab += st[c].w * ck[c].v.w;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}
template <int NC>
__device__ __forceinline__ void composite_ray_backward(const CompChunk (&ck)[NC], const CompState (&st)[NC], int n_chunks, int lane, int S,
                                                       float g0, float g1, float g2, float4* __restrict__ out) {
  float tail = 0.0f;                                            // reverse pass: suffix sums of w*gc
#pragma unroll
  for (int c = NC - 1; c >= 0; --c) {
    if (c < n_chunks) {
      const int s = c * 32 + lane;
      const float gc = g0 * ck[c].v.y + g1 * ck[c].v.z + g2 * ck[c].v.w;     // 0 past the end of the ray
      const float incl = warp_incl_sum(st[c].w * gc, lane);
      const float total = __shfl_sync(0xffffffffu, incl, 31);
      const float suffix = tail + (total - incl);               // sum over j > i
      if (s < S) {
        const float dalpha = st[c].T * gc - suffix / st[c].om;
        out[s] = make_float4(st[c].f * dalpha, st[c].w * g0, st[c].w * g1, st[c].w * g2);
      }
      tail += total;
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(kCompWarpsPerBlock * 32)
composite_backward_kernel(const float4* __restrict__ sigma_rgb, const float* __restrict__ z,
                          const float* __restrict__ g_rgb, int64_t n_rays, int S,
                          float4* __restrict__ d_sigma_rgb) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kCompWarpsPerBlock + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const float4* sr = sigma_rgb + ray * S;
  const float* zr = z + ray * S;
  const int n_chunks = (S + 31) >> 5;
  CompChunk ck[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) ck[c] = load_chunk(sr, zr, c * 32 + lane, S);
  const float g0 = g_rgb[3 * ray + 0], g1 = g_rgb[3 * ray + 1], g2 = g_rgb[3 * ray + 2];
  CompState st[NC];
  float ar = 0.f, ag = 0.f, ab = 0.f;
  composite_ray_forward<NC>(ck, n_chunks, lane, S, st, ar, ag, ab);
  composite_ray_backward<NC>(ck, st, n_chunks, lane, S, g0, g1, g2, d_sigma_rgb + ray * S);
}

// Training form: compositing forward + MSE (train.py:460) + its gradient + compositing backward in ONE launch.  One warp
// per ray loads the ray once, runs the forward pass, forms g = dL/d rgb of its ray, then the backward pass from the same
// registers.  The loss is reduced deterministically: per-block partial sums go to `partial`, the last block to
// finish (ticket counter behind `partial`) adds them in index order and resets the counter.
template <int NC>
__global__ void __launch_bounds__(kCompWarpsPerBlock * 32)
composite_train_kernel(const float4* __restrict__ sigma_rgb, const float* __restrict__ z, const float* __restrict__ target,
                       int64_t n_rays, int S, float grad_scale, float* __restrict__ loss, float* __restrict__ partial,
                       unsigned int* __restrict__ ticket, float4* __restrict__ d_sigma_rgb) {
  __shared__ float red[kCompWarpsPerBlock];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * kCompWarpsPerBlock + wib;
  const int n_chunks = (S + 31) >> 5;
  float sq = 0.0f;
  if (ray < n_rays) {
    const float4* sr = sigma_rgb + ray * S;
    const float* zr = z + ray * S;
    CompChunk ck[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) ck[c] = load_chunk(sr, zr, c * 32 + lane, S);
    const float t0 = target[3 * ray + 0], t1 = target[3 * ray + 1], t2 = target[3 * ray + 2];
    CompState st[NC];
    float ar = 0.f, ag = 0.f, ab = 0.f;
    composite_ray_forward<NC>(ck, n_chunks, lane, S, st, ar, ag, ab);
    const float r = warp_sum4_scatter(ar, ag, ab, 0.0f, lane);
    ar = __shfl_sync(0xffffffffu, r, 0); ag = __shfl_sync(0xffffffffu, r, 8); ab = __shfl_sync(0xffffffffu, r, 16);
    const float d0 = ar - t0, d1 = ag - t1, d2 = ab - t2;
    sq = d0 * d0 + d1 * d1 + d2 * d2;
    const float gs = grad_scale * 2.0f / (float)(3 * n_rays);
    composite_ray_backward<NC>(ck, st, n_chunks, lane, S, gs * d0, gs * d1, gs * d2, d_sigma_rgb + ray * S);
  }
  if (lane == 0) red[wib] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = 0.0f;
    for (int w = 0; w < kCompWarpsPerBlock; ++w) b += red[w];
    partial[blockIdx.x] = b;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && wib == 0) {
    __threadfence();
    float acc = 0.0f;
    for (unsigned int i = lane; i < gridDim.x; i += 32) acc += __ldcg(partial + i);     // fixed order per lane, fixed tree below
    acc = warp_sum(acc);
    if (lane == 0) { loss[0] = acc / (float)(3 * n_rays); *ticket = 0u; }
  }
}

// Any ray length (S > 32 * kCompMaxChunks, e.g. --n_sample 512): two passes over the ray instead of per-lane state for every
// chunk.  Pass 1 is the forward pass (rgb of the ray).  The backward needs suffix_i = sum_{j>i} w_j (g . c_j); since
// sum_j w_j (g . c_j) = g . rgb, it is total - inclusive prefix, so pass 2 re-reads the ray front to back, recomputes
// T / alpha / w per chunk and carries the prefix.  TRAIN: g = dL/d rgb of the MSE (train.py:460) and the loss partial sums, as in
// composite_train_kernel; otherwise g comes from g_rgb.
__device__ __forceinline__ float comp_chunk_forward(const CompChunk& ck, int s, int S, int lane, float& carry, float& alpha, float& om, float& delta) {
  om = 1.0f; alpha = 0.0f; delta = 0.0f;
  if (s < S) {
    sample_alpha(ck.z, ck.zn, s, S, ck.v.x, delta, alpha);
    om = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
  }
  const float incl = warp_incl_prod(om, lane);
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.0f;
  const float T = carry * excl;
  carry *= __shfl_sync(0xffffffffu, incl, 31);
  return T;
}

template <bool TRAIN>
__global__ void __launch_bounds__(kCompWarpsPerBlock * 32)
composite_backward_long_kernel(const float4* __restrict__ sigma_rgb, const float* __restrict__ z, const float* __restrict__ g_or_target,
                               int64_t n_rays, int S, float grad_scale, float* __restrict__ loss, float* __restrict__ partial,
                               unsigned int* __restrict__ ticket, float4* __restrict__ d_sigma_rgb) {
  __shared__ float red[kCompWarpsPerBlock];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * kCompWarpsPerBlock + wib;
  float sq = 0.0f;
  if (ray < n_rays) {
    const float4* sr = sigma_rgb + ray * S;
    const float* zr = z + ray * S;
    float4* out = d_sigma_rgb + ray * S;
    float g0 = g_or_target[3 * ray + 0], g1 = g_or_target[3 * ray + 1], g2 = g_or_target[3 * ray + 2];
    float carry = 1.0f, ar = 0.f, ag = 0.f, ab = 0.f;
    for (int base = 0; base < S; base += 32) {                    // pass 1: forward
      const int s = base + lane;
      const CompChunk ck = load_chunk(sr, zr, s, S);
      float alpha, om, delta;
      const float T = comp_chunk_forward(ck, s, S, lane, carry, alpha, om, delta);
      const float w = alpha * T;
      ar += w * ck.v.y; ag += w * ck.v.z; ab += w * ck.v.w;
    }
    const float r = warp_sum4_scatter(ar, ag, ab, 0.0f, lane);
    ar = __shfl_sync(0xffffffffu, r, 0); ag = __shfl_sync(0xffffffffu, r, 8); ab = __shfl_sync(0xffffffffu, r, 16);
    if (TRAIN) {
      const float d0 = ar - g0, d1 = ag - g1, d2 = ab - g2;       // g_or_target holds the target colour
      sq = d0 * d0 + d1 * d1 + d2 * d2;
      const float gs = grad_scale * 2.0f / (float)(3 * n_rays);
      g0 = gs * d0; g1 = gs * d1; g2 = gs * d2;
    }
    const float total = g0 * ar + g1 * ag + g2 * ab;              // = sum_j w_j (g . c_j)
    float prefix = 0.0f;
    carry = 1.0f;
    for (int base = 0; base < S; base += 32) {                    // pass 2: backward, front to back
      const int s = base + lane;
      const CompChunk ck = load_chunk(sr, zr, s, S);
      float alpha, om, delta;
      const float T = comp_chunk_forward(ck, s, S, lane, carry, alpha, om, delta);
      const float w = alpha * T;
      const float gc = g0 * ck.v.y + g1 * ck.v.z + g2 * ck.v.w;
      const float incl = warp_incl_sum(w * gc, lane);
      const float suffix = total - (prefix + incl);               // sum over j > i
      if (s < S) {
        const float dalpha = T * gc - suffix / om;
        out[s] = make_float4(delta * (1.0f - alpha) * dalpha, w * g0, w * g1, w * g2);
      }
      prefix += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  if (!TRAIN) return;
  if (lane == 0) red[wib] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = 0.0f;
    for (int w = 0; w < kCompWarpsPerBlock; ++w) b += red[w];
    partial[blockIdx.x] = b;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && wib == 0) {
    __threadfence();
    float acc = 0.0f;
    for (unsigned int i = lane; i < gridDim.x; i += 32) acc += __ldcg(partial + i);
    acc = warp_sum(acc);
    if (lane == 0) { loss[0] = acc / (float)(3 * n_rays); *ticket = 0u; }
  }
}

// train.py:460 nn.MSELoss (mean over n*3) and its gradient; single block.
__global__ void __launch_bounds__(1024)
mse_loss_kernel(const float* __restrict__ rgb, const float* __restrict__ target, int64_t n3,
                float grad_scale, float* __restrict__ loss, float* __restrict__ g_rgb) {
  __shared__ float red[32];
  float acc = 0.0f;
  const float gs = grad_scale * 2.0f / (float)n3;
  for (int64_t i = threadIdx.x; i < n3; i += blockDim.x) {
    const float d = rgb[i] - target[i];
    acc += d * d;
    if (g_rgb) g_rgb[i] = gs * d;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
    v = warp_sum(v);
    if (threadIdx.x == 0) loss[0] = v / (float)n3;
  }
}

// torch.optim.Adam's per-step scalars (torch/optim/adam.py, _single_tensor_adam): formed in fp64 and rounded to fp32 once,
// exactly where torch rounds them -- lerp weight 1 - beta1, beta2 and 1 - beta2 of the second-moment update,
// step_size = lr / (1 - beta1^t), sqrt(1 - beta2^t), eps.
struct AdamCoef { float step_size, b2, w1, w2, eps, bc2_sqrt; };
__host__ __device__ inline AdamCoef adam_coef(double lr, double beta1, double beta2, double eps, double step) {
  AdamCoef c;
  c.step_size = (float)(lr / (1.0 - pow(beta1, step)));
  c.b2 = (float)beta2; c.w1 = (float)(1.0 - beta1); c.w2 = (float)(1.0 - beta2);
  c.eps = (float)eps;
  c.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, step));
  return c;
}

// Device-side schedule of a graph-replayable step (nerf_train_step): iteration counter in device memory; the step size
// follows StepLR (train.py:403-404, lr0 * gamma^(iteration / decay_steps)) and the bias corrections follow torch.optim.Adam,
// evaluated in fp64 like the host path of nerf_adam_step and rounded to fp32 once.
struct AdamSchedule { const unsigned long long* iter_ptr; double lr0, gamma, beta1, beta2, eps; long long decay_steps; };

__device__ __forceinline__ void adam_device_schedule(const AdamSchedule& sched, AdamCoef& c) {
  if (sched.iter_ptr == nullptr) return;                 // block-uniform
  __shared__ AdamCoef sh;
  if (threadIdx.x == 0) {
    const unsigned long long it = *sched.iter_ptr;
    const double lr = sched.lr0 * pow(sched.gamma, (double)(it / (unsigned long long)sched.decay_steps));
    sh = adam_coef(lr, sched.beta1, sched.beta2, sched.eps, (double)(it + 1ull));
  }
  __syncthreads();
  c = sh;
}

// one element of torch.optim.Adam (train.py:401), in torch's operation order:
//   m = lerp(m, g, 1 - b1);  v = v * b2 + (1 - b2) * (g * g);  p += (-step_size * m) / (sqrt(v) / sqrt(bc2) + eps)
__device__ __forceinline__ void adam_element(float& p, float g, float& m, float& v, const AdamCoef& c) {
  m = m + c.w1 * (g - m);
  v = __fmaf_rn(c.w2, __fmul_rn(g, g), __fmul_rn(v, c.b2));
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), c.bc2_sqrt), c.eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(-c.step_size, m), denom));
}

// torch.optim.Adam over the flat parameter vector, four elements per thread (16-byte loads / stores; the
// flat vectors are 16-byte aligned and both parameter counts are multiples of 4 -- the API falls back to VEC = 1 otherwise).
template <int VEC>
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, AdamCoef c, float grad_scale, AdamSchedule sched) {
  adam_device_schedule(sched, c);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i * VEC >= n) return;
  float pi[VEC], gi[VEC], mi[VEC], vi[VEC];
  if (VEC == 4) {
    *reinterpret_cast<float4*>(pi) = reinterpret_cast<const float4*>(p)[i];
    *reinterpret_cast<float4*>(gi) = reinterpret_cast<const float4*>(g)[i];
    *reinterpret_cast<float4*>(mi) = reinterpret_cast<const float4*>(m)[i];
    *reinterpret_cast<float4*>(vi) = reinterpret_cast<const float4*>(v)[i];
  } else {
    pi[0] = p[i]; gi[0] = g[i]; mi[0] = m[i]; vi[0] = v[i];
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) adam_element(pi[k], gi[k] * grad_scale, mi[k], vi[k], c);
  if (VEC == 4) {
    reinterpret_cast<float4*>(p)[i] = *reinterpret_cast<float4*>(pi);
    reinterpret_cast<float4*>(m)[i] = *reinterpret_cast<float4*>(mi);
    reinterpret_cast<float4*>(v)[i] = *reinterpret_cast<float4*>(vi);
  } else {
    p[i] = pi[0]; m[i] = mi[0]; v[i] = vi[0];
  }
}

// Data-parallel Adam with the gradient exchange FUSED into the optimiser kernel: every rank's wgrad has written its local
// fp32 gradient into a symmetric (peer-mapped) buffer; this kernel sums them while it reads them --
//   * NVLS: one multimem.ld_reduce on the multicast address of the buffer, the NVSwitch adds the ranks' values in flight;
//   * otherwise: loads from every rank's buffer over NVLink, summed in rank order
// -- so there is no allreduce kernel and no reduced copy in HBM.  Both forms give every rank the same bits (the same
// operation on the same operands in the same order), so the replicated weights stay bit-identical.  4 elements per thread.
struct PeerGrads { const float* ptr[8]; };
// Cross-rank barrier fused into the kernel: flags.ptr[r] = rank r's array of 8 epoch words (symmetric memory, mapped here);
// word i of rank r's array is written by rank i.  Every rank announces "my gradient of exchange number `epoch` is complete"
// to all ranks, then every block waits until all ranks have announced it to this rank.  Epochs only grow, so nothing is reset.
struct PeerFlags { unsigned int* ptr[8]; };
__global__ void adam_reduce_kernel(float* __restrict__ p, PeerGrads peers, int n_ranks, const float* mc_grads, float* __restrict__ m,
                                   float* __restrict__ v, int64_t n4, AdamCoef c, float grad_scale, AdamSchedule sched,
                                   PeerFlags flags, int rank, const unsigned long long* __restrict__ epoch_ptr) {
  adam_device_schedule(sched, c);
  // this rank's own operands do not depend on the other ranks: they are on their way while the kernel waits at the barrier
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), mi = pi, vi = pi;
  if (i < n4) { pi = reinterpret_cast<float4*>(p)[i]; mi = reinterpret_cast<float4*>(m)[i]; vi = reinterpret_cast<float4*>(v)[i]; }
  if (epoch_ptr != nullptr) {                            // block-uniform
    const unsigned int epoch = (unsigned int)(*epoch_ptr) + 1u;
    if (blockIdx.x == 0 && (int)threadIdx.x < n_ranks) {
      // this kernel runs after wgrad in stream order: the local gradient is complete; make it visible system-wide, then signal
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.ptr[threadIdx.x] + rank), "r"(epoch) : "memory");
    }
    if ((int)threadIdx.x < n_ranks) {
      const unsigned int* mine = flags.ptr[rank] + threadIdx.x;
      unsigned int seen;
      const long long t0 = clock64();
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if ((int)(seen - epoch) < 0 && (clock64() - t0) > 20000000000LL) {      // ~10 s: a rank is missing -- fail loudly, do not hang the box
          printf("nerf_b200: gradient-exchange barrier timeout, rank %d waits for rank %d (epoch %u, seen %u)\n", rank, (int)threadIdx.x, epoch, seen);
          __trap();
        }
      } while ((int)(seen - epoch) < 0);
    }
    __syncthreads();
  }
  if (i >= n4) return;
  float4 g;
  if (mc_grads != nullptr) {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w) : "l"(reinterpret_cast<const float4*>(mc_grads) + i) : "memory");
  } else {
    g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < n_ranks; ++r) {
      const float4 t = __ldcv(reinterpret_cast<const float4*>(peers.ptr[r]) + i);     // volatile: written by another GPU
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
  }
  const float gg[4] = {g.x * grad_scale, g.y * grad_scale, g.z * grad_scale, g.w * grad_scale};
  float* pp = &pi.x; float* mm = &mi.x; float* vv = &vi.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) adam_element(pp[k], gg[k], mm[k], vv[k], c);
  reinterpret_cast<float4*>(p)[i] = pi; reinterpret_cast<float4*>(m)[i] = mi; reinterpret_cast<float4*>(v)[i] = vi;
}

}  // namespace nerf
