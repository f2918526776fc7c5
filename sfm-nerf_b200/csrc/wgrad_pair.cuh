// wgrad_pair.cuh -- K4b weight gradients on CTA PAIRS (tcgen05 cta_group::2): dW_l[o][i] = sum_rows dY_l[row][o] * X_l[row][i],
// every operand block of the stash / dstash read from HBM exactly ONCE per step.
//
// Why pairs.  The single-CTA kernel (wgrad.cuh) keeps a 256 x N fp32 block in the 512 TMEM columns of one SM, so skip_layer
// (N = 256 + 64) and rgb_layer1 (N = 256 + 32) had to be split into two items that each re-read dY, the encoding tile was read by
// two items, and the two CUDA-core heads ran as a second kernel re-reading h8: 12,368 B per sample against 10,960 B of distinct
// operand bytes.  With M = 256 spread over a pair, each CTA holds 128 output rows x ALL N columns, so one pair accumulates
//     skip_layer:  dY5 x [h4 | PE]  and, from the same staged PE, fc1: dY0 x PE                      (416 columns)
//     rgb_layer1:  dY10 x [h9 | DE] with rgb_layer2's head riding on the peer's rows (r1 x cotangents) (320 columns)
//     feature_layer: dY9 x h8, and density_layer's head h8 x cotangents from the SAME staged h8        (288 columns)
// and each CTA loads only ITS half of every operand (its 128 dY channels, its N/2 input channels).  Bias gradients come from
// the tensor core too (dY x a constant-ones operand), so no CUDA-core pass reads the staged tiles.
//
// The heads' cotangents (d sigma_pre, d rgb_pre[3]) are written by dgrad as a bf16 hi + lo pair ([2 k-cores][128 rows][8], hi in
// k-core 0 / lo in k-core 1: ~16 mantissa bits); they form a 16-wide operand whose two halves the two CTAs hold, and the
// epilogue adds column c and column c + 8.
//
//   warp 0    : loader -- per 128-row tile the item's 2-3 operand loads (up to 3 bulk copies each) into a ring of tile STAGES
//   warp 1    : leader: MMA issuer (one thread);  peer: relays "my half of load L has landed" to the leader
//   warp 2    : head-bias sums from the staged cotangents (items that stage them) + the per-slot consumer arrival
//   warps 4-7 : epilogue -- TMEM -> red.global.add.f32 into the flat gradient (split-K over pairs), transposed through
//               shared memory so that one instruction adds 2 x 16 CONSECUTIVE floats (4 sectors instead of 32)
//
// Ring.  What limits a stream like this is how far ahead of the MMAs the loads are issued: a load takes 2,400-3,000 cycles
// under load, a tile of MMAs 1,400.  The first version had 5 slots of 40 KB -- 2.5 tiles, i.e. the second operand of a tile
// was requested only one tile ahead: 3,640 cycles per tile, 58 % of the DRAM peak (ncu, profiles/r2_ncu_wgrad.csv).  Now every
// item has its own geometry: `depth` stages (3 for the 64-70 KB tiles, 2 for skip_layer + fc1's 104 KB) holding all loads of
// one tile at fixed offsets, so every load is issued two tiles ahead.
//
// Work split.  The items cost between 64 and 104 KB per tile and CTA; giving each item a whole number of the 74 pairs leaves
// the slowest pair 7.6 % above the mean.  Instead all (item, tile) work is laid on one line in item order and cut into
// equal pieces: a pair works on at most two items (UNITS), flushing its accumulators in between.
#pragma once
#include "common.cuh"
#include "mlp_layout.cuh"
#include "tc.cuh"

namespace nerf {

constexpr int kW2MaxSlots = 9;                               // depth (<= 3) x loads per tile (<= 3)
constexpr int kW2Ones = kW2RingBytes;                        // 2 KB: one k-core of bf16 1.0 (bias gradients through the MMA)
constexpr int kW2Scratch = kW2Ones + 2048;                   // 4 epilogue warps x 32 rows x 17 floats
constexpr int kW2ScratchWarp = 32 * 17 * 4;
constexpr int kW2Bars = kW2Scratch + 4 * kW2ScratchWarp;     // full[9] @0, empty[9] @72, acc @144, tmem_free @152; tmem ptr @160
constexpr int kW2SmemBytes = kW2Bars + 192;
constexpr int kW2Threads = 256;
static_assert(kW2SmemBytes <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

struct W2PairArgs {
  const uint8_t* stash;
  const uint8_t* dstash;
  float* grads;
  int32_t stash_tile_bytes, dstash_tile_bytes;
  int32_t n_items, n_pairs;
  int32_t n_units[kW2MaxPairs];
  W2Unit unit[kW2MaxPairs][kW2MaxUnits];
  W2Item items[kMaxW2Items];
};

__device__ __forceinline__ void umma_bf16_pair_x8(uint32_t d_tmem, const uint64_t (&a)[8], const uint64_t (&b)[8], uint32_t idesc,
                                                  uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t"
      "setp.ne.b32 p, %18, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %9, %17, p;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %2, %10, %17, t;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %3, %11, %17, t;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %4, %12, %17, t;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %5, %13, %17, t;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %6, %14, %17, t;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %7, %15, %17, t;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %8, %16, %17, t;\n\t}"
      ::"r"(d_tmem), "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "l"(a[4]), "l"(a[5]), "l"(a[6]), "l"(a[7]),
        "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]), "l"(b[4]), "l"(b[5]), "l"(b[6]), "l"(b[7]), "r"(idesc), "r"(accumulate_first) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kW2Threads, 1) wgrad_pair_kernel(const __grid_constant__ W2PairArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = tc::smem_u32(smem);
  const uint32_t s_ones = s_base + kW2Ones;
  const uint32_t s_bar = s_base + kW2Bars;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 72, bar_acc = s_bar + 144, bar_free = s_bar + 152;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + kW2Bars + 160);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int i = 0; i < kW2MaxSlots; ++i) {
      tc::mbar_init(bar_full + 8 * i, rank == 0 ? 2 : 1);      // leader: its own expect_tx arrival + the peer's relayed arrival
      tc::mbar_init(bar_empty + 8 * i, 2);                     // the MMA warp's multicast commit + this CTA's warp 2
    }
    tc::mbar_init(bar_acc, 1);
    tc::mbar_init(bar_free, 8);                                // leader only: 4 epilogue warps of each CTA have drained TMEM
    tc::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + kW2Ones)[i] = 0x3F803F80u;   // bf16 1.0 x 2
  tc::fence_proxy_async_smem();
  if (warp == 0) tc::tmem_alloc_pair(tc::smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
  tc::tc_fence_before();
  tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int pair = (int)(blockIdx.x >> 1);
  const int n_units = a.n_units[pair];
  // Every role walks the same sequence of fills: unit by unit, tile by tile, load by load; fill number j of a unit's tile
  // sequence goes to slot (tile % depth) * n_loads + load.  Each role tracks the phase parity of every slot in a bit mask.

  if (warp == 0) {
    // ===================== loader: this CTA's halves of the operand loads =====================
    if (lane == 0) {
      uint32_t par = 0;                                        // bit s: parity of the NEXT fill of slot s
      for (int u = 0; u < n_units; ++u) {
        const W2Item& W = a.items[a.unit[pair][u].item];
        const int t0 = a.unit[pair][u].t0, t1 = a.unit[pair][u].t1;
        for (int t = t0; t < t1; ++t) {
          const uint8_t* st_tile = a.stash + (size_t)t * a.stash_tile_bytes;
          const uint8_t* ds_tile = a.dstash + (size_t)t * a.dstash_tile_bytes;
          const int stage = (t - t0) % W.depth;
          for (int l = 0; l < W.n_loads; ++l) {
            const W2Load& ld = W.load[l];
            const uint32_t slot = (uint32_t)(stage * W.n_loads + l);
            tc::mbar_wait(bar_empty + 8 * slot, ((par >> slot) & 1u) ^ 1u);
            par ^= 1u << slot;
            uint32_t total = 0;
            for (int sg = 0; sg < ld.n_segs; ++sg) total += (uint32_t)ld.seg[sg].bytes;
            tc::mbar_arrive_expect_tx(bar_full + 8 * slot, total);
            uint32_t dst = s_base + (uint32_t)(stage * W.stage_bytes + W.load_off[l]);
            for (int sg = 0; sg < ld.n_segs; ++sg) {
              const W2Seg& sgm = ld.seg[sg];
              const uint8_t* src = (sgm.from_dstash[rank] ? ds_tile : st_tile) + sgm.off[rank];
              tc::bulk_g2s(dst, src, (uint32_t)sgm.bytes, bar_full + 8 * slot);
              dst += (uint32_t)sgm.bytes;
            }
          }
        }
        // Unit boundary (the next unit lays its stages out differently) and kernel end (no multicast commit may still be on
        // its way to this CTA's barriers when it exits): wait for the release of the last fill of every slot.
        for (uint32_t sl = 0; sl < (uint32_t)kW2MaxSlots; ++sl) tc::mbar_wait(bar_empty + 8 * sl, ((par >> sl) & 1u) ^ 1u);
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank != 0) {
    // ===================== peer: relay "my half of this load has landed" to the leader's full barrier =====================
    if (lane == 0) {
      const uint32_t leader_full = tc::mapa(bar_full, 0);
      uint32_t par = 0;
      for (int u = 0; u < n_units; ++u) {
        const W2Item& W = a.items[a.unit[pair][u].item];
        const int n_t = a.unit[pair][u].t1 - a.unit[pair][u].t0;
        for (int t = 0; t < n_t; ++t)
          for (int l = 0; l < W.n_loads; ++l) {
            const uint32_t slot = (uint32_t)((t % W.depth) * W.n_loads + l);
            tc::mbar_wait(bar_full + 8 * slot, (par >> slot) & 1u);
            par ^= 1u << slot;
            tc::mbar_arrive_cluster(leader_full + 8 * slot);
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== leader: MMA issuer (warp-uniform loop, one elected lane issues) =====================
    uint32_t par = 0;
    for (int u = 0; u < n_units; ++u) {
      const W2Item& W = a.items[a.unit[pair][u].item];
      const int n_t = a.unit[pair][u].t1 - a.unit[pair][u].t0;
      if (u > 0) {                                              // the previous unit's accumulators have been read out in both CTAs
        tc::mbar_wait_cluster(bar_free, (uint32_t)(u - 1) & 1u);
        tc::tc_fence_after();
      }
      for (int t = 0; t < n_t; ++t) {
        const int stage = t % W.depth;
        for (int l = 0; l < W.n_loads; ++l) {                    // every load of the tile has landed in both CTAs
          const uint32_t slot = (uint32_t)(stage * W.n_loads + l);
          tc::mbar_wait_cluster(bar_full + 8 * slot, (par >> slot) & 1u);
          par ^= 1u << slot;
        }
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t s_stage = s_base + (uint32_t)(stage * W.stage_bytes);
          for (int q = 0; q < W.n_mmas; ++q) {
            const W2Mma& mm = W.mma[q];
            const uint32_t sa = s_stage + (uint32_t)W.load_off[mm.a_load] + (uint32_t)mm.a_kc * kKcoreBytes;
            const uint32_t sb = mm.b_load < 0 ? s_ones : s_stage + (uint32_t)W.load_off[mm.b_load] + (uint32_t)mm.b_kc * kKcoreBytes;
            // 16 rows (the contraction axis) per UMMA, MN-major operands straight from the [k_core][row][8] tiles; the ones
            // operand is the same 16 rows for every k-step
            const uint32_t b_step = mm.b_load < 0 ? 0u : 256u;
            uint64_t da[8], db[8];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              da[ks] = tc::umma_smem_desc(sa + ks * 256, 128, kKcoreBytes);
              db[ks] = tc::umma_smem_desc(sb + ks * b_step, 128, kKcoreBytes);
            }
            umma_bf16_pair_x8(tmem_base + (uint32_t)mm.d_col, da, db, tc::umma_idesc_bf16(256, mm.n, 1, 1), t > 0 ? 1u : 0u);
          }
          for (int l = 0; l < W.n_loads; ++l) tc::umma_commit_pair(bar_empty + 8 * (uint32_t)(stage * W.n_loads + l), 3);
        }
        __syncwarp();
      }
      if (tc::elect_one()) tc::umma_commit_pair(bar_acc, 3);
      __syncwarp();
    }
  } else if (warp == 2) {
    // ===================== consumer arrival per slot + head-bias sums from the staged cotangents =====================
    // The cotangent k-core of this CTA ([128 rows][8] bf16: hi parts in the leader, lo parts in the peer; channels 0-3 =
    // d sigma_pre, d rgb_pre[0..2]) sums over rows to this CTA's share of d b_density / d b_rgb2.
    uint32_t par = 0;
    for (int u = 0; u < n_units; ++u) {
      const W2Item& W = a.items[a.unit[pair][u].item];
      const int n_t = a.unit[pair][u].t1 - a.unit[pair][u].t0;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int t = 0; t < n_t; ++t) {
        const int stage = t % W.depth;
        for (int l = 0; l < W.n_loads; ++l) {
          const uint32_t slot = (uint32_t)(stage * W.n_loads + l);
          tc::mbar_wait(bar_full + 8 * slot, (par >> slot) & 1u);
          par ^= 1u << slot;
          if (l == W.cot_load) {
            const uint32_t sc = s_base + (uint32_t)(stage * W.stage_bytes + W.load_off[l]) + (uint32_t)W.cot_kc * kKcoreBytes;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              uint32_t u0, u1;
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(u0), "=r"(u1) : "r"(sc + (lane + 32 * rr) * 16));
              s0 += __uint_as_float(u0 << 16); s1 += __uint_as_float(u0 & 0xffff0000u);
              s2 += __uint_as_float(u1 << 16); s3 += __uint_as_float(u1 & 0xffff0000u);
            }
          }
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(bar_empty + 8 * slot);   // after the slot was filled for THIS load: never an earlier phase
        }
      }
      if (W.cot_load >= 0 && n_t > 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
        }
        if (lane == 0) {
          if (W.g_bd >= 0) atomicAdd(a.grads + W.g_bd, s0);
          if (W.g_b2 >= 0) { atomicAdd(a.grads + W.g_b2, s1); atomicAdd(a.grads + W.g_b2 + 1, s2); atomicAdd(a.grads + W.g_b2 + 2, s3); }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM accumulators -> flat gradient vector =====================
    const int lq = warp & 3;                       // TMEM lane quarter this warp may read
    const int m_local = lq * 32 + lane;            // this CTA's row of the M = 256 block
    const uint32_t t_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
    float* scratch = reinterpret_cast<float*>(smem + kW2Scratch + lq * kW2ScratchWarp);
    const uint32_t leader_free = tc::mapa(bar_free, 0);
    for (int u = 0; u < n_units; ++u) {
      const W2Item& W = a.items[a.unit[pair][u].item];
      tc::mbar_wait(bar_acc, (uint32_t)u & 1u);
      tc::tc_fence_after();
      for (int oi = 0; oi < W.n_outs; ++oi) {
        const W2Out& o = W.out[oi];
        if (!((o.rank_mask >> rank) & 1)) continue;
        if (o.kind == W2_OUT_WEIGHT) {
          // rows lq*32 .. +31 of the block, 16 columns at a time: registers (lane = row) -> scratch -> (2 rows x 16 columns) per add
          float* dst0 = a.grads + o.dst + (int64_t)(o.row_base[rank] + lq * 32) * o.ld + o.col0;
          for (int ch = 0; ch * 32 < o.n_valid; ++ch) {
            uint32_t v[32];
            tc::tmem_ld32(t_lane + (uint32_t)(o.d_col + ch * 32), v);
            tc::tmem_wait_ld();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int cbase = ch * 32 + h * 16;
              if (cbase < o.n_valid) {                           // warp-uniform
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 16; ++j) scratch[lane * 17 + j] = __uint_as_float(v[h * 16 + j]);
                __syncwarp();
                const int c = lane & 15;
                if (cbase + c < o.n_valid) {
#pragma unroll 4
                  for (int it = 0; it < 16; ++it) {
                    const int r = 2 * it + (lane >> 4);
                    atomicAdd(dst0 + (int64_t)r * o.ld + cbase + c, scratch[r * 17 + c]);
                  }
                }
              }
            }
          }
        } else {
          const int row = o.row_base[rank] + m_local;
          uint32_t v[32];
          tc::tmem_ld32(t_lane + (uint32_t)o.d_col, v);          // a 16-column block (the upper 16 columns read are ignored)
          tc::tmem_wait_ld();
          if (o.kind == W2_OUT_BIAS) {                           // dY x ones: every column holds sum_rows dY[row][m]
            atomicAdd(a.grads + o.dst + row, __uint_as_float(v[0]));
          } else {                                               // head: column c (hi cotangents) + column c + 8 (lo cotangents)
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c >= o.col0 && c < o.col0 + o.n_valid)
                atomicAdd(a.grads + o.dst + (int64_t)(c - o.col0) * o.ld + row, __uint_as_float(v[c]) + __uint_as_float(v[c + 8]));
          }
        }
      }
      // this warp's TMEM reads are done: the next unit's first MMAs may overwrite the accumulators
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0 && u + 1 < n_units) { if (rank == 0) tc::mbar_arrive(bar_free); else tc::mbar_arrive_cluster(leader_free); }
    }
  }
  tc::tc_fence_before();
  tc::cluster_sync();                              // no remote arrive / multicast commit may target a CTA that has exited
  if (warp == 0) tc::tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace nerf
