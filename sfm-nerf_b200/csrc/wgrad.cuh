// wgrad.cuh -- K4b weight gradients: dW_l[o][i] = sum_rows dY_l[row][o] * X_l[row][i] on tcgen05.
// The contraction axis is the SAMPLE axis, so both operands are read MN-major straight from the
// blocked stash/dstash tiles ([k_core][row][8]); the fp32 accumulator (up to 256x256) lives in the
// whole 512-column TMEM of one SM for the CTA's entire row range and is added into the flat
// gradient vector once at the end (split-K over CTAs, red.global.add.f32).
//   warp 0    : loader -- 128-row x 128-channel sub-blocks of dY / X, ONE bulk copy (<= 32 KB) each, ring of 6 slots
//   warp 1    : MMA issuer (one thread)
//   warps 2-5 : bias gradients (column sums of dY from the staged tile) + final TMEM -> HBM epilogue
// Head gradients (density_layer, rgb_layer2: N < 16, not UMMA shapes) run on CUDA cores.
#pragma once
#include "common.cuh"
#include "mlp_layout.cuh"
#include "tc.cuh"

namespace nerf {

constexpr int kWgSlots = 6;                // ring of operand sub-blocks
constexpr int kWgSlotBytes = 32768;        // one 128-row block of 128 channels (dY M-half or X N-half)
constexpr int kWgSmemBytes = kWgSlots * kWgSlotBytes + 128;

struct WgradArgs {
  const uint8_t* stash;
  const uint8_t* dstash;
  float* grads;
  int32_t stash_tile_bytes, dstash_tile_bytes;
  int32_t n_tiles, n_items;
  int32_t cta_begin[kMaxWgradItems + 1];   // CTAs [cta_begin[i], cta_begin[i+1]) work on item i
  WgradItem items[kMaxWgradItems];
};

// Per 128-row tile an item streams its operands as sub-blocks in the order dY0, X0, [X1], [dY1]
// (dY_i = M-half i of dY, 16 k-cores; X_j = N-part j of the layer input, <= 16 k-cores) and issues the
// UMMA groups (dY0,X0) (dY0,X1) (dY1,X0) (dY1,X1).  Each sub-block's ring slot is released right after
// the last group that reads it, so the loader always has 2-3 sub-blocks (64-96 KB) in flight.
struct WgPlan {
  int m_halves, x_parts, n_loads;
  int dy_load[2], x_load[2];               // position of each sub-block in the per-tile load order
  uint32_t dy_bytes, x_bytes[2];
  int x_n[2];                              // UMMA N of each X part
};
__device__ __forceinline__ WgPlan make_plan(const WgradItem& W) {
  WgPlan p;
  p.m_halves = W.m_kcores / 16;
  p.x_parts = W.n > 128 ? 2 : 1;
  p.dy_bytes = 16 * kKcoreBytes;
  p.x_n[0] = W.n > 128 ? 128 : W.n; p.x_n[1] = W.n - p.x_n[0];
  p.x_bytes[0] = (uint32_t)(p.x_n[0] / 8) * kKcoreBytes; p.x_bytes[1] = (uint32_t)(p.x_n[1] / 8) * kKcoreBytes;
  p.dy_load[0] = 0; p.x_load[0] = 1; p.x_load[1] = 2;
  p.dy_load[1] = 1 + p.x_parts;
  p.n_loads = p.m_halves + p.x_parts;
  return p;
}

__global__ void __launch_bounds__(192, 1) wgrad_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = tc::smem_u32(smem);
  const uint32_t s_bar = s_base + kWgSlots * kWgSlotBytes;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 48, bar_acc = s_bar + 96;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + kWgSlots * kWgSlotBytes + 104);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgSlots; ++i) { tc::mbar_init(bar_full + 8 * i, 1); tc::mbar_init(bar_empty + 8 * i, 5); }
    tc::mbar_init(bar_acc, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // which item / row range is mine
  int item = 0;
  while (item + 1 < a.n_items && (int)blockIdx.x >= a.cta_begin[item + 1]) ++item;
  const WgradItem& W = a.items[item];
  const int splits = a.cta_begin[item + 1] - a.cta_begin[item];
  const int split = (int)blockIdx.x - a.cta_begin[item];
  const int t0 = (int)(((int64_t)a.n_tiles * split) / splits);
  const int t1 = (int)(((int64_t)a.n_tiles * (split + 1)) / splits);
  const int n_my_tiles = t1 - t0;
  const WgPlan pl = make_plan(W);
  const int m_halves = pl.m_halves;

  if (warp == 0) {
    // ===================== loader: one bulk copy per operand sub-block =====================
    if (lane == 0) {
      uint32_t L = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        const uint8_t* dy = a.dstash + (size_t)(t0 + t) * a.dstash_tile_bytes + W.dy_off;
        const uint8_t* x = a.stash + (size_t)(t0 + t) * a.stash_tile_bytes + W.x_off;
        for (int l = 0; l < pl.n_loads; ++l, ++L) {
          const uint8_t* src; uint32_t bytes;
          if (l == 0) { src = dy; bytes = pl.dy_bytes; }
          else if (l <= pl.x_parts) { src = x + (size_t)(l - 1) * 16 * kKcoreBytes; bytes = pl.x_bytes[l - 1]; }
          else { src = dy + 16 * kKcoreBytes; bytes = pl.dy_bytes; }
          const uint32_t slot = L % kWgSlots, ph = (L / kWgSlots) & 1u;
          tc::mbar_wait(bar_empty + 8 * slot, ph ^ 1u);
          tc::mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
          tc::bulk_g2s(s_base + slot * kWgSlotBytes, src, bytes, bar_full + 8 * slot);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    uint32_t L0 = 0;                                   // index of this tile's first load
    for (int t = 0; t < n_my_tiles; ++t, L0 += (uint32_t)pl.n_loads) {
      for (int i = 0; i < m_halves; ++i) {
        const uint32_t La = L0 + (uint32_t)pl.dy_load[i];
        const uint32_t slot_a = La % kWgSlots;
        tc::mbar_wait(bar_full + 8 * slot_a, (La / kWgSlots) & 1u);
        for (int j = 0; j < pl.x_parts; ++j) {
          const uint32_t Lb = L0 + (uint32_t)pl.x_load[j];
          const uint32_t slot_b = Lb % kWgSlots;
          tc::mbar_wait(bar_full + 8 * slot_b, (Lb / kWgSlots) & 1u);
          tc::tc_fence_after();
          const uint32_t sa = s_base + slot_a * kWgSlotBytes, sb = s_base + slot_b * kWgSlotBytes;
          const uint32_t idesc = tc::umma_idesc_bf16(128, pl.x_n[j], 1, 1);
          const uint32_t d_col = tmem_base + (uint32_t)(i * W.n + j * 128);
          if (tc::elect_one()) {
            // 16 rows (the contraction axis) per UMMA; the eight UMMAs of a 128-row block go out from one asm block with
            // sixteen distinct descriptor registers: issued one by one they reuse the same uniform registers and each
            // waits ~225 cycles on the previous one's scoreboard, which made this kernel issue-bound (232 cycles per MMA
            // measured) instead of HBM-bound.
            uint64_t da[8], db[8];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              da[ks] = tc::umma_smem_desc(sa + ks * 256, 128, kKcoreBytes);
              db[ks] = tc::umma_smem_desc(sb + ks * 256, 128, kKcoreBytes);
            }
            tc::umma_bf16_x8(d_col, da, db, idesc, t > 0 ? 1u : 0u);
            if (i == m_halves - 1) tc::umma_commit(bar_empty + 8 * slot_b);   // last reader of X_j
            if (j == pl.x_parts - 1) tc::umma_commit(bar_empty + 8 * slot_a); // last reader of dY_i
          }
          __syncwarp();
        }
      }
    }
    if (tc::elect_one()) tc::umma_commit(bar_acc);
    __syncwarp();
  } else {
    // ===================== bias sums + epilogue =====================
    const int q = warp - 2;                       // k_cores q, q+4, q+8, q+12 of each dY half
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    uint32_t L0 = 0;
    for (int t = 0; t < n_my_tiles; ++t, L0 += (uint32_t)pl.n_loads) {
      for (int l = 0; l < pl.n_loads; ++l) {
        const uint32_t L = L0 + (uint32_t)l;
        const uint32_t slot = L % kWgSlots;
        tc::mbar_wait(bar_full + 8 * slot, (L / kWgSlots) & 1u);
        const bool is_dy = (l == 0) || (l > pl.x_parts);
        if (is_dy && W.b_off >= 0) {
          const uint32_t sa = s_base + slot * kWgSlotBytes;
          const int half = (l == 0) ? 0 : 1;
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int kc = q + 4 * ii;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              uint32_t u0, u1, u2, u3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3) : "r"(sa + kc * kKcoreBytes + (lane + 32 * rr) * 16));
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                if (hh == half) {                    // compile-time accumulator index, runtime half
                  acc[hh * 4 + ii][0] += __uint_as_float(u0 << 16); acc[hh * 4 + ii][1] += __uint_as_float(u0 & 0xffff0000u);
                  acc[hh * 4 + ii][2] += __uint_as_float(u1 << 16); acc[hh * 4 + ii][3] += __uint_as_float(u1 & 0xffff0000u);
                  acc[hh * 4 + ii][4] += __uint_as_float(u2 << 16); acc[hh * 4 + ii][5] += __uint_as_float(u2 & 0xffff0000u);
                  acc[hh * 4 + ii][6] += __uint_as_float(u3 << 16); acc[hh * 4 + ii][7] += __uint_as_float(u3 & 0xffff0000u);
                }
              }
            }
          }
        }
        __syncwarp();
        // every slot's "empty" barrier counts this warp: arrive only after the slot has been filled for
        // THIS load, so the arrival cannot land in an earlier phase
        if (lane == 0) tc::mbar_arrive(bar_empty + 8 * slot);
      }
    }
    if (W.b_off >= 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if ((i >> 2) < m_halves) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v = acc[i][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(a.grads + W.b_off + ((i >> 2) * 16 + q + 4 * (i & 3)) * 8 + j, v);
          }
        }
      }
    }
    // final epilogue: TMEM accumulator -> gradient vector
    tc::mbar_wait(bar_acc, 0);
    tc::tc_fence_after();
    const int lq = warp & 3;                      // TMEM lane quarter this warp may read
    const int n_chunks = (W.n + 31) >> 5;
    if (n_my_tiles > 0) {
      for (int mh = 0; mh < m_halves; ++mh) {
        const int o = mh * 128 + lq * 32 + lane;
        float* dst = a.grads + W.w_off + (int64_t)o * W.in_f + W.col0;
        for (int ch = 0; ch < n_chunks; ++ch) {
          uint32_t v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + mh * W.n + ch * 32, v);
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ch * 32 + j < W.n_valid) atomicAdd(dst + ch * 32 + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// Head gradients on CUDA cores (train: dL/d density_layer, dL/d rgb_layer2).
//   d w_d[j]   = sum_rows dsp[row] * h8[row][j]      d b_d    = sum_rows dsp[row]
//   d W2[c][j] = sum_rows drgb_pre[c][row] * r1[row][j]   d b2[c] = sum_rows drgb_pre[c][row]
// Threads 0-255: the 32 k-cores of h8 (density_layer); 256-383: the 16 k-cores of r1 (rgb_layer2).
struct HeadGradArgs {
  const uint8_t* stash;
  const uint8_t* dstash;
  float* grads;
  int32_t stash_tile_bytes, dstash_tile_bytes, n_tiles;
  int32_t st_h8, st_r1, ds_sr;
  int32_t g_wd, g_bd, g_w2, g_b2;     // float offsets in the flat gradient vector
};
// Every thread owns one k-core (8 channels) and every 8th row of a tile: 16-byte loads, eight neighbouring threads read 128
// contiguous bytes.  Partial sums are reduced over the 8 row-phases with shuffles at the end.
__global__ void __launch_bounds__(384) head_grad_kernel(const HeadGradArgs a) {
  const int t = threadIdx.x;
  const bool dens = t < 256;
  const int kc = dens ? (t >> 3) : ((t - 256) >> 3);       // k-core of h8 (32) / r1 (16)
  const int rp = t & 7;                                     // row phase
  float acc[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
  float accb0 = 0.f, accb1 = 0.f, accb2 = 0.f, accbd = 0.f;
  // no block-level staging or barriers: the per-row cotangents (16 B) are read through L1 by the 48 threads that need them,
  // so consecutive tiles' loads overlap freely
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const uint8_t* st = a.stash + (size_t)tile * a.stash_tile_bytes;
    const float4* dsr = reinterpret_cast<const float4*>(a.dstash + (size_t)tile * a.dstash_tile_bytes + a.ds_sr);
    const uint4* src = reinterpret_cast<const uint4*>(st + (dens ? a.st_h8 : a.st_r1) + kc * kKcoreBytes);
#pragma unroll 8
    for (int i = 0; i < 16; ++i) {
      const int r = rp + 8 * i;
      const uint4 u = __ldg(src + r);
      const float4 d = __ldg(dsr + r);
      float v[8] = {__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u),
                    __uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u), __uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u)};
      if (dens) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[0][j] = fmaf(d.x, v[j], acc[0][j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[0][j] = fmaf(d.y, v[j], acc[0][j]); acc[1][j] = fmaf(d.z, v[j], acc[1][j]); acc[2][j] = fmaf(d.w, v[j], acc[2][j]); }
      }
    }
    if (t < 128) { const float4 d = __ldg(dsr + t); accbd += d.x; accb0 += d.y; accb1 += d.z; accb2 += d.w; }   // bias sums: thread = row
  }
  // reduce over the 8 row phases (lanes differing in bits 0-2)
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) acc[c][j] += __shfl_xor_sync(0xffffffffu, acc[c][j], o);
  if (rp == 0) {
    if (dens) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(a.grads + a.g_wd + kc * 8 + j, acc[0][j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(a.grads + a.g_w2 + kc * 8 + j, acc[0][j]);
        atomicAdd(a.grads + a.g_w2 + 128 + kc * 8 + j, acc[1][j]);
        atomicAdd(a.grads + a.g_w2 + 256 + kc * 8 + j, acc[2][j]);
      }
    }
  }
  if (t < 128) {                                            // bias gradients: warp sums, one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      accbd += __shfl_xor_sync(0xffffffffu, accbd, o); accb0 += __shfl_xor_sync(0xffffffffu, accb0, o);
      accb1 += __shfl_xor_sync(0xffffffffu, accb1, o); accb2 += __shfl_xor_sync(0xffffffffu, accb2, o);
    }
    if ((t & 31) == 0) {
      atomicAdd(a.grads + a.g_bd, accbd);
      atomicAdd(a.grads + a.g_b2, accb0); atomicAdd(a.grads + a.g_b2 + 1, accb1); atomicAdd(a.grads + a.g_b2 + 2, accb2);
    }
  }
}

}  // namespace nerf
