// mlp_fused.cuh -- K4: the fused NeRF MLP (NeRFModel.py:158-202) as ONE persistent tcgen05 kernel:
//   MODE_FWD : positions/directions (or rays + depths) -> encoding -> 11 chained UMMA layers ->
//              density/rgb heads; activations never leave the SM (training additionally streams
//              a bf16 stash of every layer input + ReLU bit masks for the backward kernels).
//   MODE_BWD : the data-gradient chain dY_l = (dY_{l+1} . W_{l+1}) * relu'(h_l), same machinery
//              with transposed weight images; emits every dY_l to the dstash.
//
// CTA PAIR (cluster of 2 = the two SMs of a TPC, tcgen05 cta_group::2).  Every CTA keeps TWO 128-row
// tiles resident (2 x 64 KB activation tiles in shared memory, 2 x 256 TMEM columns); tile slot s of
// the leader (rank 0) and of the peer (rank 1) form one M = 256 UMMA, so a weight slab (one K=16 step)
// is split between the two CTAs: each loads and holds only its N/2 rows (4 KB), which halves the
// B-operand shared-memory reads per SM.
// SCHEDULE: one warp issues every MMA of the pair in the fixed order
//     slot 0 layer L (all k-steps) | slot 1 layer L | slot 0 layer L+1 | slot 1 layer L+1 | ...
// and the tensor pipe executes in issue order, so slot 0's epilogue of layer L runs exactly under slot 1's MMAs of
// layer L and vice versa: the pipe only idles if an epilogue takes longer than a layer of MMAs.  The loader streams the
// slabs in that same order through one ring of 8 slab pairs (each layer's slabs twice, once per slot: 32 B/clk/SM from L2).
// Sharing one pass of slabs between the slots was measured too: it ties the slots to within one ring of each other,
// too close to hide an epilogue (also with the lane-parallel loaders); a private 8-slot ring per slot covers less
// than the load latency.
//   warps 0-3 / 4-7 : epilogue of tile slot 0 / 1 (thread = row = TMEM lane), both CTAs:
//                     TMEM -> regs -> ReLU+bf16 pack (one cvt) -> next layer's A tile, in place
//   warps 8, 10 (+ 11 in the inference kernel) : weight loaders (tensor-map copies of this CTA's half slabs into the ring, completion signalled on the
//                     LEADER's barriers from both CTAs), both CTAs
//   warp 9          : leader: MMA issuer (one elected lane), tcgen05.mma cta_group::2 kind::f16 256xNx16,
//                     completion multicast to both CTAs' barriers
// Biases ride on the MMA: channel 63 of the encoding tile is the constant 1.0 and the packed weight
// image carries the bias in that K column, so the epilogue does no bias loads or adds.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "composite.cuh"
#include "mlp_layout.cuh"
#include "tc.cuh"

namespace nerf {

constexpr int MODE_FWD = 0;
constexpr int MODE_BWD = 1;
// 8 epilogue warps + 2 loaders + MMA issuer (leader; warp 11 and the peer's warp 9 have no role): the CTA must be a whole number of
// 128-thread warpgroups.  tcgen05.ld lets warp w touch TMEM lanes 32*(w%4).., and only for warpgroup-aligned CTAs is a
// warp's index in the CTA tied to the SM sub-partition it runs on: with 320 threads the epilogue warps were seen in
// rotated hardware slots (cuda-gdb) and tcgen05.ld raised "misaligned address", depending on what ran on the SM before.
constexpr int kMlpThreads = 384;
// How a layer's output tile reaches the training stash (forward) / the gradient stash (dgrad): the forward's epilogue threads read
// their rows of the tile back after the release and stream them with 16-byte stores; dgrad stores the packed words straight
// from the accumulator read-out loop.  Round 2 built and A/B-timed every alternative -- stores from the read-out registers in
// the forward, one bulk (TMA) store per tile, bulk stores in 2-8 KB pieces with 0-2 in flight, read-back for dgrad, read-back
// stores that yield to the other slot's read-out -- and all were equal or slower: both kernels are bound by how fast an SM
// streams these stores (22 B/clk = 6.3 TB/s chip-wide, tools/store_rate.py), not by who issues them.  The table is
// profiles/r2_stash_store_variants.md; the code of the variants is in the history (commits 74f67cc .. a0e94fb).
#ifndef NERF_WGRAD_PAIR
#define NERF_WGRAD_PAIR 1           // 1: wgrad_pair.cuh (CTA pairs, heads on the tensor core); 0: wgrad.cuh + head_grad_kernel
#endif
#ifndef NERF_READBACK_BATCH
#define NERF_READBACK_BATCH 8       // k-cores (16 B per row each) a read-back loop loads before it stores the first
#endif

struct MlpArgs {
  CUtensorMap tm4, tm2;      // the weight images as [rows][256] uint32 (1 KB rows), boxes of 4 / 2 rows = this CTA's half of an N = 256 / 128 slab
  const uint8_t* packed;     // weight images
  const float* small;        // heads (fp32)
  // forward inputs: per-row (pos, dir) or per-ray (rays_o, rays_d, z) with n_sample > 0
  const float* pos;
  const float* dir;
  const float* z;
  int32_t n_sample;
  int32_t pos_L, dir_L;
  int32_t density_relu;      // density head activation: 1 = ReLU (legacy checkpoints), 0 = softplus
  int32_t c_kcores;          // hidden width / 8: k-cores of an activation tile that hold data (32 or 16)
  int32_t sm_wd, sm_bd, sm_w2, sm_b2;
  int64_t n_rows;
  int32_t n_tiles;
  float* sigma_rgb;          // fwd out [n_rows,4]; bwd in
  const float* d_sigma_rgb;  // bwd in [n_rows,4]
  uint8_t* stash;            // fwd: written when non-null; bwd: read
  uint8_t* dstash;           // bwd: written
  int32_t stash_tile_bytes, dstash_tile_bytes;
  int32_t st_pe, st_de, st_r1, st_mask, ds_y10, ds_sr, ds_cot;
  long long* trace;          // diagnostics: per-(pair,step,slot) clock64 stamps of CTA 0 (or null)
  // fused render (inference, n_sample == 64): per-RAY outputs composited in the epilogue (train.py:248-269); sigma_rgb is then not written
  float* comp_rgb; float* comp_depth; float* comp_acc;
  Program prog;
};

struct MlpSmem {
  static constexpr int kAct = 0;                                   // 2 x 64 KB
  static constexpr int kEnc = kAct + 2 * kActBytes;                // 2 x 16 KB
  static constexpr int kRing = kEnc + 2 * kEncBytes;               // 8 x 8 KB (slab pairs: this CTA's half of two k-steps each)
  static constexpr int kBars = kRing + kRingSlots * kRingSlotBytes;
  // barriers (8 B each): full[8] @0, empty[8] @128, acc_full[2] @256, act_ready[2] @272, drain @288; tmem ptr @320
  static constexpr int kTmemPtr = kBars + 320;
  static constexpr int kHead = kBars + 384;                        // fp32 head tables (w_d, b_d, W2, b2): 648 floats
  static constexpr int kHeadFloats = 648;
  static constexpr int kTotal = kHead + kHeadFloats * 4;
};
static_assert(MlpSmem::kTotal <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
static_assert(MlpSmem::kBars + 296 <= MlpSmem::kTmemPtr && MlpSmem::kTmemPtr + 4 <= MlpSmem::kHead, "barriers / TMEM pointer / head tables overlap");

// sin/cos positional encoding of one scalar for level `i` from the exact decomposition
// a_i = 2^i * fl32(fl32(pi) * x) (NeRFModel.py:151-156): the angle is carried in
// turns as a two-term fp32 number, scaled by 2^i exactly, reduced to [-0.5,0.5] exactly, and only
// then handed to the SFU (abs error ~5e-7, far below the bf16 operand rounding that follows).
struct TurnsF { float hi, lo; };
__device__ __forceinline__ TurnsF pe_turns_f(float x) {
  const float a0 = __fmul_rn(3.14159274101257324f, x);
  const float c_hi = 0.15915493667125702f;          // fl32(1/(2*pi))
  const float c_lo = 6.4206383e-09f;                // 1/(2*pi) - c_hi
  TurnsF t;
  t.hi = __fmul_rn(a0, c_hi);
  t.lo = __fmaf_rn(a0, c_hi, -t.hi) + a0 * c_lo;
  return t;
}
__device__ __forceinline__ void pe_sincos(const TurnsF& t, int level, float& s, float& c) {
  const float sc = __uint_as_float((uint32_t)(127 + level) << 23);   // 2^level
  const float h = t.hi * sc;
  const float r = (h - rintf(h)) + t.lo * sc;
  const float a = 6.28318530717958647692f * r;
  s = __sinf(a);
  c = __cosf(a);
}

// Encodes v[3] with L levels into bf16 channels [v, sin0(v), cos0(v), ...] padded with zeros to
// 8*NKC channels, stored at enc + kcore*2048 + row*16.  ONE: the last channel is the constant 1.0
// (the K column the packed weight images pair with the bias).
// MAXL: the levels this instance is unrolled for; `L` (<= MAXL, warp-uniform) levels are evaluated, the rest stay zero.
template <int MAXL, int NKC, bool ONE>
__device__ __forceinline__ void write_encoding(uint32_t enc_smem, int row, const float v[3], uint8_t* gdst, int L) {
  float e[NKC * 8];
#pragma unroll
  for (int i = 0; i < NKC * 8; ++i) e[i] = 0.0f;
  e[0] = v[0]; e[1] = v[1]; e[2] = v[2];
  if (MAXL > 0) {
    TurnsF t[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = pe_turns_f(v[k]);
#pragma unroll
    for (int i = 0; i < MAXL; ++i) {
      if (i < L) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float s, c;
          pe_sincos(t[k], i, s, c);
          e[3 + 6 * i + k] = s;
          e[6 + 6 * i + k] = c;
        }
      }
    }
  }
  if (ONE) e[NKC * 8 - 1] = 1.0f;
#pragma unroll
  for (int kc = 0; kc < NKC; ++kc) {
    const uint4 p = make_uint4(tc::pack_bf16x2(e[8 * kc + 0], e[8 * kc + 1]), tc::pack_bf16x2(e[8 * kc + 2], e[8 * kc + 3]),
                               tc::pack_bf16x2(e[8 * kc + 4], e[8 * kc + 5]), tc::pack_bf16x2(e[8 * kc + 6], e[8 * kc + 7]));
    tc::st_shared_v4(enc_smem + kc * kKcoreBytes + row * 16, p.x, p.y, p.z, p.w);
    if (gdst != nullptr) __stcs(reinterpret_cast<uint4*>(gdst + kc * kKcoreBytes) + row, p);   // training stash copy (streaming store)
  }
}

// word `i` (runtime) of eight mask words held in registers, without dynamic register indexing
__device__ __forceinline__ uint32_t pick8(const uint4& a, const uint4& b, int i) {
  uint32_t r = a.x;
  r = (i == 1) ? a.y : r; r = (i == 2) ? a.z : r; r = (i == 3) ? a.w : r;
  r = (i == 4) ? b.x : r; r = (i == 5) ? b.y : r; r = (i == 6) ? b.z : r; r = (i == 7) ? b.w : r;
  return r;
}
__device__ __forceinline__ void put8(uint4& a, uint4& b, int i, uint32_t v) {
  a.x = (i == 0) ? v : a.x; a.y = (i == 1) ? v : a.y; a.z = (i == 2) ? v : a.z; a.w = (i == 3) ? v : a.w;
  b.x = (i == 4) ? v : b.x; b.y = (i == 5) ? v : b.y; b.z = (i == 6) ? v : b.z; b.w = (i == 7) ? v : b.w;
}

__device__ __forceinline__ float softplus_f(float x) {        // torch default beta=1, threshold=20
  return x > 20.0f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int MODE, bool TRAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1) mlp_fused_kernel(const __grid_constant__ MlpArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = tc::smem_u32(smem);
  const uint32_t s_ring = s_base + MlpSmem::kRing;
  const uint32_t s_bar = s_base + MlpSmem::kBars;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 128, bar_acc = s_bar + 256, bar_act = s_bar + 272;
  const uint32_t bar_drain = s_bar + 288;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + MlpSmem::kTmemPtr);
  const uint32_t rank = tc::cluster_ctarank();         // 0 = leader (issues the pair MMAs), 1 = peer
  // head weights live in shared memory: with 224 KB of it carved out there is no L1 left to cache them
  float* sm = reinterpret_cast<float*>(smem + MlpSmem::kHead);
  for (int i = threadIdx.x; i < MlpSmem::kHeadFloats; i += blockDim.x) sm[i] = a.small[i];

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops
  // (barrier addresses, descriptors, counters) on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    // only the LEADER's full barriers are used: its loader's expect_tx arrival, transaction bytes from both CTAs' copies
    for (int i = 0; i < kRingSlots; ++i) { tc::mbar_init(bar_full + 8 * i, 1); tc::mbar_init(bar_empty + 8 * i, 1); }
    // act_ready lives in the leader: 128 epilogue threads of each CTA arrive on it
    for (int i = 0; i < 2; ++i) { tc::mbar_init(bar_acc + 8 * i, 1); tc::mbar_init(bar_act + 8 * i, 256); }
    tc::mbar_init(bar_drain, 1);                       // the MMA warp: "all my MMAs have completed" (multicast to both CTAs)
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc_pair(tc::smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
  tc::tc_fence_before();
  tc::cluster_sync();                                  // both CTAs' barriers exist before any remote arrive / multicast commit
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const Program& P = a.prog;
  const int n_quads = (a.n_tiles + 3) >> 2;            // 4 tiles per cluster iteration: tile = 4*quad + 2*slot + rank
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  // The MMA issuer and the loaders walk ONE sequence of passes: pass p = layer p % n_steps of this cluster's quad
  // p / n_steps, tile slot 0's pass p and then tile slot 1's pass p.  The ring is addressed in SLAB PAIRS (two consecutive
  // k-steps of a pass, 2 x 4 KB per CTA, one full and one empty barrier per pair): every pass starts on a fresh pair, a pass with
  // an odd number of k-steps (16 activation k-steps + the bias k-step) leaves the second half of its last pair unused.  Before
  // that (one barrier per slab, released in pairs) slot 1's passes began on an odd slab: the loader lane of the even slab then
  // waited for a release that came a whole MMA group later, and those passes issued at ~170 instead of ~135 cycles per MMA
  // (tools/trace_mlp.py: 2,900 against 2,300 cycles per 17-k-step pass, except at the 20-k-step skip layer).
  const int n_local = cluster_id < n_quads ? (n_quads - cluster_id + n_clusters - 1) / n_clusters : 0;   // quads of this cluster
  const int n_pass = n_local * P.n_steps;

  // Loader warps: 8 and 10, and in the inference kernel also warp 11.  Inference is bound by the slab supply (no stash stores: the
  // tensor pipe waits for weights), and a third warp taking every third slab pair shortens the chain wait(empty) -> copies ->
  // expect_tx per pair: 0.271 -> 0.248 ms at 4096 x 64 (1.39 PFLOP/s).  The training forward and dgrad are bound by their stores:
  // three loaders measured 1 % slower there; four in the peer (its warp 9 has no role) and three in the leader: 0.31 ms.
  constexpr bool kInfer = (MODE == MODE_FWD) && !TRAIN;
  const int n_loaders = kInfer ? 3 : 2;
  // ... and there every loader warp takes TWO consecutive slab pairs per trip, one per lane (the lanes wait for their slots side by
  // side): 0.248 -> 0.225 ms.  Six pairs in flight is the limit: with nine or twelve (3-4 lanes) a lane can be a whole ring round
  // ahead of an unloaded slot, its parity wait aliases, and the kernel deadlocks (measured: mbarrier time-outs).
  const int n_lanes = kInfer ? 2 : 1;
  const int loader_idx = warp == 8 ? 0 : warp == 10 ? 1 : warp == 11 ? 2 : -1;
  if (loader_idx >= 0 && loader_idx < n_loaders) {
    // ===================== weight loaders (both CTAs: this CTA's N/2 rows of every slab, every layer twice) ==========
    // One slab pair per trip and lane, the loader warps taking the trips in turn (a single thread doing wait(empty) -> copies ->
    // arrive.expect_tx pair after pair is a chain of ~130-cycle barrier operations; several warps overlap their trips).  The copies
    // are 2-D tensor-map (TMA) loads in the cta_group::2 form, whose completion may be signalled on the OTHER CTA's mbarrier: both
    // CTAs' halves of a pair report their bytes to the leader's full barrier, which the leader's loader arms with one
    // expect_tx for all of them (the copies are issued before that arrival: the phase cannot complete while it is pending).
    // Round 1 / early round 2 used 1-D bulk copies onto the CTA's own barrier and two relay warps in the peer that forwarded
    // "landed" to the leader (remote mbarrier arrive): one hop more in the ring's round trip, ~1 % on every MLP kernel.
    const int which = loader_idx;
    const uint64_t keep_in_l2 = tc::l2_policy_evict_last();    // the 2.6 MB of weight images are re-read by every CTA, all the time
    const uint32_t leader_full = tc::mapa(bar_full, 0);
    uint32_t pj0 = 0, trip = 0;                          // slab pairs issued before the current pass; trip counter
    int si = 0;
    for (int ps = 0; ps < n_pass; ++ps) {
      const Step& st = P.s[si];
      si = (si + 1 == P.n_steps) ? 0 : si + 1;
      const uint32_t bytes = (uint32_t)st.n * 16u;
      const int total = st.act_ksteps + st.enc_ksteps + st.bias_kstep;
      const int n_pairs = (total + 1) >> 1;
      const uint64_t tmap = reinterpret_cast<uint64_t>(bytes == 4096u ? &a.tm4 : &a.tm2);
      for (int slot = 0; slot < 2; ++slot) {             // slot 0's pass, then slot 1's
        for (int j0 = 0; j0 < n_pairs; j0 += n_lanes, ++trip) {
          if ((int)(trip % (uint32_t)n_loaders) != which) continue;
          const int j = j0 + lane;
          if (lane < n_lanes && j < n_pairs) {
            const uint32_t it = pj0 + (uint32_t)j;
            const uint32_t rs = it % kRingSlots, ph = (it / kRingSlots) & 1u;
            const bool two = 2 * j + 1 < total;
            tc::mbar_wait(bar_empty + 8 * rs, ph ^ 1u);
            // row (1 KB) of this CTA's half of k-step 2j inside the weight buffer; both CTAs' copies report to the LEADER's barrier
            const int32_t row0 = (int32_t)(((uint32_t)st.w_off + (uint32_t)(2 * j) * (2u * bytes) + rank * bytes) >> 10);
            const int32_t step_rows = (int32_t)((2u * bytes) >> 10);
            tc::tma_g2s_2d_pair(s_ring + rs * kRingSlotBytes, tmap, 0, row0, leader_full + 8 * rs, keep_in_l2);
            if (two) tc::tma_g2s_2d_pair(s_ring + rs * kRingSlotBytes + kSlabBytes, tmap, 0, row0 + step_rows, leader_full + 8 * rs, keep_in_l2);
            if (rank == 0) tc::mbar_arrive_expect_tx(bar_full + 8 * rs, 2u * (two ? 2u * bytes : bytes));
          }
          __syncwarp();
        }
        pj0 += (uint32_t)n_pairs;
      }
    }
    if (warp == 8 && lane == 0) {
      // tail: every multicast commit onto this CTA's barriers has landed before it may exit -- the releases of the last
      // ring-ful of pairs (each is the phase a further load into that slot would wait for), then the drain barrier
      const uint32_t n_all = pj0;
      for (uint32_t it = n_all > (uint32_t)kRingSlots ? n_all - kRingSlots : 0u; it < n_all; ++it)
        tc::mbar_wait(bar_empty + 8 * (it % kRingSlots), (it / kRingSlots) & 1u);
      tc::mbar_wait(bar_drain, 0);
    }
    __syncwarp();
  } else if (warp == 9 && rank == 0) {
    // ===================== leader MMA issuer =====================
    // The warp runs with warp-uniform control flow (all lanes poll the barriers) and one elected lane issues the
    // cta_group::2 instructions (M = 256: this CTA's tile + the peer's tile of the same slot).  What shaped this loop
    // (all measured, tools/umma_rate.py + tools/trace_mlp.py):
    //  * consecutive UTCHMMAs that reuse the same uniform registers for their descriptors serialise on the previous
    //    one's scoreboard (~225 cycles per MMA): k-steps are issued four at a time from ONE asm block whose eight
    //    descriptors are live inputs, with the two ring releases (tcgen05.commit, multicast to both CTAs) inside the block;
    //  * any mbarrier operation costs the warp ~130 cycles: the full barriers of a trip (8 k-steps = 4 slab pairs) are probed
    //    back to back (independent try_waits overlap) and blocked on one by one only when a pair has not landed yet;
    //  * the next pass's "A tile ready" barrier is probed together with the last trip's full barriers, so the hand-over
    //    from one tile slot to the other costs no serial barrier round trip (the pipe ran dry for 400-900 cycles there);
    //  * k-steps are grouped from the start of the pass, across the activation / encoding / bias runs of the A operand.
    const uint64_t desc_hi = (uint64_t)(((128u >> 4) & 0x3FFFu) | (1u << 14)) << 32;      // SBO = 128 B, version 1
    const uint32_t a_lbo = ((uint32_t)kKcoreBytes >> 4) << 16;                              // LBO = 2048 B
    const uint32_t ring_lo = (s_ring & 0x3FFFFu) >> 4;
    constexpr uint32_t kStepLo = (2 * kKcoreBytes) >> 4;                                    // A descriptor advance per k-step
    constexpr uint32_t kSlotLo = kRingSlotBytes >> 4, kSlabLo = kSlabBytes >> 4;
    uint32_t pj = 0;                                               // slab pairs consumed so far: ring slot pj % 8, parity (pj / 8) & 1
    bool act_ok = false;                                           // the coming pass's A tile was already seen ready
    int si = 0;
    for (int ps = 0; ps < n_pass; ++ps) {
      const int n = P.s[si].n, act_ks = P.s[si].act_ksteps, enc_ks = P.s[si].enc_ksteps;
      const int total = act_ks + enc_ks + P.s[si].bias_kstep;
      const int si_now = si; (void)si_now;
      si = (si + 1 == P.n_steps) ? 0 : si + 1;
      const uint32_t idesc = tc::umma_idesc_bf16(256, n, 0, 0);
      const uint32_t b_lo = ((((uint32_t)n * 8u) >> 4) << 16) | ring_lo;        // LBO = (N/2 rows) * 16 B
#pragma unroll 1
      for (int slot = 0; slot < 2; ++slot) {
        const uint32_t act_lo = a_lbo | (((s_base + MlpSmem::kAct + slot * kActBytes) & 0x3FFFFu) >> 4);
        const uint32_t enc_lo = a_lbo | (((s_base + MlpSmem::kEnc + slot * kEncBytes) & 0x3FFFFu) >> 4);
        const uint32_t d_tmem = tmem_base + 256 * slot;
        // the pass after this one: the other slot's pass ps (slot 0 -> 1) or slot 0's pass ps + 1
        const bool has_next = slot == 0 || ps + 1 < n_pass;
        const uint32_t next_bar = bar_act + 8 * (slot ^ 1), next_par = (uint32_t)(ps + slot) & 1u;
        // A descriptor of k-step k of the pass: activation k-steps, then encoding k-steps, then the bias k-step (ENC k-step 3)
        auto a_desc = [&](int k) -> uint64_t {
          const int e = k - act_ks;
          return desc_hi | ((e < 0) ? act_lo + (uint32_t)k * kStepLo : enc_lo + (uint32_t)(e < enc_ks ? e : 3) * kStepLo);
        };
        auto b_desc = [&](uint32_t pair, uint32_t half) -> uint64_t {
          return desc_hi | (b_lo + (pair % kRingSlots) * kSlotLo + half * kSlabLo);
        };
        auto full_bar = [&](uint32_t pair) -> uint32_t { return bar_full + 8 * (pair % kRingSlots); };
        auto empty_bar = [&](uint32_t pair) -> uint32_t { return bar_empty + 8 * (pair % kRingSlots); };
        auto par_of = [&](uint32_t pair) -> uint32_t { return (pair / kRingSlots) & 1u; };

        if (!act_ok) tc::mbar_wait_cluster(bar_act + 8 * slot, (uint32_t)ps & 1u);   // this slot's A tile of the layer is ready
        act_ok = false;
        tc::tc_fence_after();
#ifdef NERF_DIAGNOSTICS
        long long* tr = (a.trace != nullptr && blockIdx.x == 0 && lane == 0)
                            ? a.trace + ((size_t)((ps / P.n_steps) * kMaxSteps + si_now) * 2 + slot) * 4 : nullptr;
        if (tr) tr[0] = clock64();
#endif
        int k = 0;
#pragma unroll 1
        while (total - k >= 8) {                         // a trip: 8 k-steps = 4 slab pairs, two 4-MMA blocks
          const bool last = (total - k == 8);
          const uint64_t a0 = a_desc(k), a1 = a_desc(k + 1), a2 = a_desc(k + 2), a3 = a_desc(k + 3),
                         a4 = a_desc(k + 4), a5 = a_desc(k + 5), a6 = a_desc(k + 6), a7 = a_desc(k + 7);
          const uint64_t b0 = b_desc(pj, 0), b1 = b_desc(pj, 1), b2 = b_desc(pj + 1, 0), b3 = b_desc(pj + 1, 1),
                         b4 = b_desc(pj + 2, 0), b5 = b_desc(pj + 2, 1), b6 = b_desc(pj + 3, 0), b7 = b_desc(pj + 3, 1);
          const uint32_t r01 = (uint32_t)tc::mbar_try_wait_cluster(full_bar(pj), par_of(pj)) & (uint32_t)tc::mbar_try_wait_cluster(full_bar(pj + 1), par_of(pj + 1));
          const uint32_t r23 = (uint32_t)tc::mbar_test_wait(full_bar(pj + 2), par_of(pj + 2)) & (uint32_t)tc::mbar_test_wait(full_bar(pj + 3), par_of(pj + 3));
          if (last && has_next) act_ok = tc::mbar_test_wait(next_bar, next_par);
          if (!r01) { tc::mbar_wait_cluster(full_bar(pj), par_of(pj)); tc::mbar_wait_cluster(full_bar(pj + 1), par_of(pj + 1)); }
          if (tc::elect_one())
            tc::umma_bf16_pair_x4_release2(d_tmem, a0, a1, a2, a3, b0, b1, b2, b3, idesc, k > 0 ? 1u : 0u, empty_bar(pj), empty_bar(pj + 1));
          __syncwarp();
          if (!r23) { tc::mbar_wait_cluster(full_bar(pj + 2), par_of(pj + 2)); tc::mbar_wait_cluster(full_bar(pj + 3), par_of(pj + 3)); }
          if (tc::elect_one())
            tc::umma_bf16_pair_x4_release2(d_tmem, a4, a5, a6, a7, b4, b5, b6, b7, idesc, 1u, empty_bar(pj + 2), empty_bar(pj + 3));
          __syncwarp();
          pj += 4; k += 8;
        }
#pragma unroll 1
        while (k < total) {                              // the remainder of the pass: up to four k-steps (two pairs) at a time
          const int kn = (total - k) < 4 ? (total - k) : 4;
          const bool last = (k + kn == total);
          const uint32_t r0 = (uint32_t)tc::mbar_try_wait_cluster(full_bar(pj), par_of(pj));
          const uint32_t r1 = kn > 2 ? (uint32_t)tc::mbar_test_wait(full_bar(pj + 1), par_of(pj + 1)) : 1u;
          if (last && has_next) act_ok = tc::mbar_test_wait(next_bar, next_par);
          if (!r0) tc::mbar_wait_cluster(full_bar(pj), par_of(pj));
          if (!r1) tc::mbar_wait_cluster(full_bar(pj + 1), par_of(pj + 1));
          if (tc::elect_one()) {
            if (kn == 4) {
              tc::umma_bf16_pair_x4_release2(d_tmem, a_desc(k), a_desc(k + 1), a_desc(k + 2), a_desc(k + 3), b_desc(pj, 0), b_desc(pj, 1),
                                             b_desc(pj + 1, 0), b_desc(pj + 1, 1), idesc, k > 0 ? 1u : 0u, empty_bar(pj), empty_bar(pj + 1));
            } else {
              for (int q = 0; q < kn; ++q) {
                tc::umma_bf16_pair(d_tmem, a_desc(k + q), b_desc(pj + (uint32_t)(q >> 1), (uint32_t)(q & 1)), idesc, (k + q) > 0 ? 1u : 0u);
                if ((q & 1) || q == kn - 1) tc::umma_commit_pair(empty_bar(pj + (uint32_t)(q >> 1)), 3);
              }
            }
          }
          __syncwarp();
          pj += (uint32_t)((kn + 1) >> 1); k += kn;
        }
        if (tc::elect_one()) tc::umma_commit_pair(bar_acc + 8 * slot, 3);
        __syncwarp();
#ifdef NERF_DIAGNOSTICS
        if (tr) tr[1] = clock64();
#endif
      }
    }
    // every earlier commit is covered by the drain barrier (both CTAs' loaders wait for it before they exit)
    if (tc::elect_one()) tc::umma_commit_pair(bar_drain, 3);
    __syncwarp();
  } else if (warp < 8) {
    // ===================== epilogue warpgroups (thread = row) =====================
    const int slot = warp >> 2;                        // tile slot of this warpgroup
    const int row = threadIdx.x & 127;                 // 0..127 == TMEM lane
    const uint32_t s_act = s_base + MlpSmem::kAct + slot * kActBytes;
    const uint32_t s_enc = s_base + MlpSmem::kEnc + slot * kEncBytes;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256 * slot;
    const uint32_t my_acc = bar_acc + 8 * slot;
    const uint32_t leader_act = tc::mapa(bar_act + 8 * slot, 0);   // "A tile of this slot is ready" lives in the leader
    constexpr bool train = TRAIN || (MODE == MODE_BWD);   // the stash exists: written by the training forward, read (masks) by dgrad
    uint32_t acc_phase = 0;
    for (int quad = cluster_id; quad < n_quads; quad += n_clusters) {
      const int tile = 4 * quad + 2 * slot + (int)rank;
      const bool tile_ok = tile < a.n_tiles;           // tile count not a multiple of 4: the last quad has phantom tiles
      const int64_t grow = (int64_t)tile * 128 + row;
      const bool valid = tile_ok && grow < a.n_rows;
      uint8_t* stash_tile = (train && tile_ok) ? a.stash + (size_t)tile * a.stash_tile_bytes : nullptr;
      uint8_t* dstash_tile = (MODE == MODE_BWD && tile_ok) ? a.dstash + (size_t)tile * a.dstash_tile_bytes : nullptr;
      const bool store_out = tile_ok && ((MODE == MODE_BWD) || train);
      float dirv[3] = {0.f, 0.f, 0.f};
      float sigma_keep = 0.0f, dsp = 0.0f;
      (void)sigma_keep; (void)dsp;

      // ---------------- prologue ----------------
      if (MODE == MODE_FWD) {
        float p[3] = {0.f, 0.f, 0.f};
        if (valid) {
          if (a.n_sample > 0) {                        // train.py:220-228: pts = o + d * z
            const int64_t ray = grow / a.n_sample;
            const float zz = a.z[grow];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              dirv[k] = a.dir[3 * ray + k];
              p[k] = __fadd_rn(a.pos[3 * ray + k], __fmul_rn(dirv[k], zz));
            }
          } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) { p[k] = a.pos[3 * grow + k]; dirv[k] = a.dir[3 * grow + k]; }
          }
        }
        uint8_t* g_pe = store_out ? stash_tile + a.st_pe : nullptr;
        if (a.pos_L > 0) write_encoding<10, 8, true>(s_enc, row, p, g_pe, a.pos_L);
        else write_encoding<0, 8, true>(s_enc, row, p, g_pe, 0);
      } else {
        // backward prologue: d(rgb2 pre-act), d(density pre-act), dR1 = d_rgb_pre . W2 masked by r1 > 0
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f), o = make_float4(0.f, 0.5f, 0.5f, 0.5f);
        uint4 r1m = make_uint4(0u, 0u, 0u, 0u);
        if (valid) {
          d = reinterpret_cast<const float4*>(a.d_sigma_rgb)[grow];
          o = reinterpret_cast<const float4*>(a.sigma_rgb)[grow];
        }
        if (tile_ok) r1m = __ldg(reinterpret_cast<const uint4*>(stash_tile + a.st_mask + 10 * 4096) + row * 2);
        dsp = a.density_relu ? (o.x > 0.0f ? d.x : 0.0f)                 // relu'
                             : d.x * (1.0f - expf(-o.x));              // softplus'(x) = 1 - exp(-softplus(x))
        const float d0 = d.y * o.y * (1.0f - o.y), d1 = d.z * o.z * (1.0f - o.z), d2 = d.w * o.w * (1.0f - o.w);
        if (tile_ok) {
          // head cotangents for the weight-gradient kernel: a bf16 hi + lo pair per value (k-core 0: hi, k-core 1: lo; channels
          // 0-3 of 8), so the heads' gradients ride on the tensor core with ~16 mantissa bits
          const float c4[4] = {dsp, d0, d1, d2};
          uint32_t hi[2], lo[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float h0 = __bfloat162float(__float2bfloat16_rn(c4[2 * j])), h1 = __bfloat162float(__float2bfloat16_rn(c4[2 * j + 1]));
            hi[j] = tc::pack_bf16x2(h0, h1);
            lo[j] = tc::pack_bf16x2(c4[2 * j] - h0, c4[2 * j + 1] - h1);
          }
          uint4* cot = reinterpret_cast<uint4*>(dstash_tile + a.ds_cot);
          cot[row] = make_uint4(hi[0], hi[1], 0u, 0u);
          cot[128 + row] = make_uint4(lo[0], lo[1], 0u, 0u);
#if NERF_WGRAD_PAIR == 0
          reinterpret_cast<float4*>(dstash_tile + a.ds_sr)[row] = make_float4(dsp, d0, d1, d2);      // single-CTA wgrad + CUDA-core head kernel
#endif
        }
        const float4* w2 = reinterpret_cast<const float4*>(sm + a.sm_w2);     // rgb_layer2.weight [3][128]: 16-byte loads (96 instead of 384)
        const uint32_t r1w[4] = {r1m.x, r1m.y, r1m.z, r1m.w};
#pragma unroll
        for (int kc = 0; kc < 16; ++kc) {
          float g[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 wa = w2[2 * kc + h], wb = w2[32 + 2 * kc + h], wc = w2[64 + 2 * kc + h];
            const float wa4[4] = {wa.x, wa.y, wa.z, wa.w}, wb4[4] = {wb.x, wb.y, wb.z, wb.w}, wc4[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int j = 4 * h + q;
              const float v = d0 * wa4[q] + d1 * wb4[q] + d2 * wc4[q];
              g[j] = ((r1w[kc >> 2] >> (((((kc & 3) * 8 + j)) >> 1) + 16 * (j & 1))) & 1u) ? v : 0.0f;
            }
          }
          const uint4 pk = make_uint4(tc::pack_bf16x2(g[0], g[1]), tc::pack_bf16x2(g[2], g[3]),
                                      tc::pack_bf16x2(g[4], g[5]), tc::pack_bf16x2(g[6], g[7]));
          tc::st_shared_v4(s_act + kc * kKcoreBytes + row * 16, pk.x, pk.y, pk.z, pk.w);
          if (store_out) __stcs(reinterpret_cast<uint4*>(dstash_tile + a.ds_y10 + kc * kKcoreBytes) + row, pk);
        }
      }
      tc::fence_proxy_async_smem();
      if (rank == 0) tc::mbar_arrive(bar_act + 8 * slot); else tc::mbar_arrive_cluster(leader_act);

      // ---------------- layer steps ----------------
      for (int si = 0; si < P.n_steps; ++si) {
        const Step& st = P.s[si];
        uint4 mk_a = make_uint4(0u, 0u, 0u, 0u), mk_b = make_uint4(0u, 0u, 0u, 0u);   // this row's ReLU mask words
        if (MODE == MODE_BWD && tile_ok) {             // fetched while the MMAs of this step still run
          const uint4* mp = reinterpret_cast<const uint4*>(stash_tile + st.mask_off) + row * 2;
          mk_a = __ldg(mp); mk_b = __ldg(mp + 1);
        }
        tc::mbar_wait(my_acc, acc_phase & 1u); ++acc_phase;
        tc::tc_fence_after();
#ifdef NERF_DIAGNOSTICS
        long long* tr = (a.trace != nullptr && blockIdx.x == 0 && row == 0)
                            ? a.trace + ((size_t)((quad / n_clusters) * kMaxSteps + si) * 2 + slot) * 4 : nullptr;
        if (tr) tr[2] = clock64();
#endif
        const int n_units = st.n >> 6;                 // 64 accumulator columns per TMEM load (4 or 2 units, even)
        const bool is_rgb = (MODE == MODE_FWD) && (st.flags & EPI_RGB) != 0;
        const bool is_dens = (MODE == MODE_FWD) && (st.flags & EPI_DENSITY) != 0;
        const bool add_sigma = (MODE == MODE_BWD) && (st.flags & EPI_BWD_SIGMA) != 0;

        uint4* g_out = (MODE == MODE_BWD && tile_ok) ? reinterpret_cast<uint4*>(dstash_tile + st.stash_off) + row : nullptr;
        (void)g_out;
        // one 64-column unit: activation / mask, bf16 pack, store as the next A tile
        auto process = [&](uint32_t (&v)[64], int u) {
          if (MODE == MODE_FWD) {
            {                                          // ReLU + bf16 pack in one cvt; the bias came through the MMA
#pragma unroll
              for (int q = 0; q < 8; ++q)
                tc::st_shared_v4(s_act + (u * 8 + q) * kKcoreBytes + row * 16,
                                 tc::pack_bf16x2_relu(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1])),
                                 tc::pack_bf16x2_relu(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])),
                                 tc::pack_bf16x2_relu(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])),
                                 tc::pack_bf16x2_relu(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
            }
          } else {
            const uint32_t w_lo = pick8(mk_a, mk_b, 2 * u), w_hi = pick8(mk_a, mk_b, 2 * u + 1);
            if (add_sigma) {                           // density head: d h8 += d_sigma_pre * w_d
              const float4* wd = reinterpret_cast<const float4*>(sm + a.sm_wd + u * 64);
#pragma unroll
              for (int j = 0; j < 64; j += 4) {
                const float4 w4 = wd[j >> 2];
                v[j] = __float_as_uint(fmaf(dsp, w4.x, __uint_as_float(v[j])));
                v[j + 1] = __float_as_uint(fmaf(dsp, w4.y, __uint_as_float(v[j + 1])));
                v[j + 2] = __float_as_uint(fmaf(dsp, w4.z, __uint_as_float(v[j + 2])));
                v[j + 3] = __float_as_uint(fmaf(dsp, w4.w, __uint_as_float(v[j + 3])));
              }
            }
            // relu'(h) in the packed domain: the mask bits of columns 2i / 2i+1 of a 32-column chunk sit at i / i+16 of its
            // mask word, so ((w >> i) & 0x00010001) * 0xFFFF is the AND mask of the bf16x2 word holding those two columns.
            // The packed words go to the next A tile AND straight to the dstash from registers (a warp writes 512
            // contiguous bytes per instruction): no shared-memory read-back after the layer is released.
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint32_t p[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int i = 4 * q + t;
                const uint32_t w = (i < 16) ? w_lo : w_hi;
                p[t] = tc::pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])) & (((w >> (i & 15)) & 0x00010001u) * 0xFFFFu);
              }
              tc::st_shared_v4(s_act + (u * 8 + q) * kKcoreBytes + row * 16, p[0], p[1], p[2], p[3]);
              if (g_out != nullptr) __stcs(g_out + (u * 8 + q) * 128, make_uint4(p[0], p[1], p[2], p[3]));   // streaming: keep the weights in L2
            }
          }
        };

        {   // double-buffered accumulator read-out: the next unit's TMEM load flies while this one is packed
          uint32_t va[64], vb[64];
          tc::tmem_ld64(t_lane, va);
          tc::tmem_wait_ld();
          for (int u = 0; u < n_units; u += 2) {
            tc::tmem_ld64(t_lane + (u + 1) * 64, vb);
            process(va, u);
            tc::tmem_wait_ld();
            if (u + 2 < n_units) tc::tmem_ld64(t_lane + (u + 2) * 64, va);
            process(vb, u + 1);
            tc::tmem_wait_ld();
          }
        }
        tc::tc_fence_before();                         // TMEM reads retired before the next MMA overwrites
        tc::fence_proxy_async_smem();                  // ACT writes visible to UMMA
        if (!(st.flags & EPI_LAST)) {                  // the next layer's MMAs may start NOW
          if (rank == 0) tc::mbar_arrive(bar_act + 8 * slot); else tc::mbar_arrive_cluster(leader_act);
        }
#ifdef NERF_DIAGNOSTICS
        if (tr) tr[3] = clock64();
#endif

        // ---- everything below is off the MMA -> epilogue -> MMA critical path ----
        if (MODE == MODE_FWD) {
          if (st.flags & EPI_WRITE_DIR) {              // ENC k-cores 0-3 <- direction encoding (used by rgb_layer1)
            uint8_t* g_de = store_out ? stash_tile + a.st_de : nullptr;
            // 27 channels (4 levels, the reference default) or fewer: k-cores 0-3, the constant-1 channel 63 of the position tile
            // stays; more levels: the whole 64-channel tile with its own constant 1
            if (a.dir_L > 4) write_encoding<10, 8, true>(s_enc, row, dirv, g_de, a.dir_L);
            else if (a.dir_L > 0) write_encoding<4, 4, false>(s_enc, row, dirv, g_de, a.dir_L);
            else write_encoding<0, 4, false>(s_enc, row, dirv, g_de, 0);
            tc::fence_proxy_async_smem();
          }
          // The two CUDA-core heads read this thread's row of the bf16 tile it has just written, AFTER the layer loop:
          // inside it the 128 accumulator registers of the double-buffered read-out leave no room to prefetch head
          // weights, and the exposed shared-memory latency made these two epilogues 3x longer than a layer of MMAs.
          if (is_dens) {                               // density_layer + softplus on h8 (NeRFModel.py:192)
            const float4* wd = reinterpret_cast<const float4*>(sm + a.sm_wd);
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll 8
            for (int kc = 0; kc < a.c_kcores; ++kc) {       // the hidden channels only: k-cores beyond them hold stale shared memory
              uint32_t x0, x1, x2, x3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(s_act + kc * kKcoreBytes + row * 16));
              const float4 wa = wd[2 * kc], wb = wd[2 * kc + 1];
              d0 = fmaf(__uint_as_float(x0 << 16), wa.x, d0); d1 = fmaf(__uint_as_float(x0 & 0xffff0000u), wa.y, d1);
              d2 = fmaf(__uint_as_float(x1 << 16), wa.z, d2); d3 = fmaf(__uint_as_float(x1 & 0xffff0000u), wa.w, d3);
              d0 = fmaf(__uint_as_float(x2 << 16), wb.x, d0); d1 = fmaf(__uint_as_float(x2 & 0xffff0000u), wb.y, d1);
              d2 = fmaf(__uint_as_float(x3 << 16), wb.z, d2); d3 = fmaf(__uint_as_float(x3 & 0xffff0000u), wb.w, d3);
            }
            const float pre = (d0 + d1) + (d2 + d3) + sm[a.sm_bd];
            sigma_keep = a.density_relu ? fmaxf(pre, 0.0f) : softplus_f(pre);
          }
          if (is_rgb) {                                // rgb_layer2 + sigmoid on r1 (NeRFModel.py:199-200)
            const float4* w2 = reinterpret_cast<const float4*>(sm + a.sm_w2);
            float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll 4
            for (int kc = 0; kc < 16; ++kc) {
              uint32_t x[4];
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(s_act + kc * kKcoreBytes + row * 16));
              float y[8];
#pragma unroll
              for (int q = 0; q < 4; ++q) { y[2 * q] = __uint_as_float(x[q] << 16); y[2 * q + 1] = __uint_as_float(x[q] & 0xffff0000u); }
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float4 wa = w2[2 * kc + h], wb = w2[32 + 2 * kc + h], wc = w2[64 + 2 * kc + h];
                c0 = fmaf(y[4 * h], wa.x, c0); c1 = fmaf(y[4 * h], wb.x, c1); c2 = fmaf(y[4 * h], wc.x, c2);
                c0 = fmaf(y[4 * h + 1], wa.y, c0); c1 = fmaf(y[4 * h + 1], wb.y, c1); c2 = fmaf(y[4 * h + 1], wc.y, c2);
                c0 = fmaf(y[4 * h + 2], wa.z, c0); c1 = fmaf(y[4 * h + 2], wb.z, c1); c2 = fmaf(y[4 * h + 2], wc.z, c2);
                c0 = fmaf(y[4 * h + 3], wa.w, c0); c1 = fmaf(y[4 * h + 3], wb.w, c1); c2 = fmaf(y[4 * h + 3], wc.w, c2);
              }
            }
            const float* b2 = sm + a.sm_b2;
            const float cr = sigmoid_f(c0 + b2[0]), cg = sigmoid_f(c1 + b2[1]), cb = sigmoid_f(c2 + b2[2]);
            bool composited = false;
            if constexpr (!TRAIN) {
              if (a.comp_rgb != nullptr) {                     // uniform: fused render, 64 or 128 samples per ray
                // A 128-row tile is two rays of 64 samples or one of 128: warp q of the warpgroup holds chunk q % NC (32 samples) of
                // local ray q / NC.  The arithmetic is composite_forward_exact_kernel<NC>'s, operation for operation (lane l of that
                // kernel's warp handles samples l, l + 32, ...), so the result has the same bits: a warp needs the transmittance
                // product of the chunks before it, and the first warp of a ray then folds the other chunks' (weight, colour, depth)
                // into its own per-lane sums in chunk order.
                composited = true;
                const int nc = a.n_sample >> 5;                // chunks per ray: 2 or 4
                const int wq = warp & 3, chunk = wq & (nc - 1), lr = wq / nc;
                const int64_t ray = (int64_t)tile * (4 / nc) + lr;
                float* sc = reinterpret_cast<float*>(smem + MlpSmem::kAct + slot * kActBytes + 16 * kKcoreBytes) + lr * 512;   // free k-cores of the tile
                const bool last_sample = chunk == nc - 1 && lane == 31;
                const float zz = valid ? a.z[grow] : 0.0f;
                const float zn = (valid && !last_sample) ? a.z[grow + 1] : zz;
                float delta = __fsub_rn(zn, zz);
                if (last_sample) delta = 1e-3f;                 // train.py:250
                const float alpha = __fsub_rn(1.0f, expf(-__fmul_rn(sigma_keep, delta)));
                const float om = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);   // train.py:262
                const float incl = warp_incl_prod(om, lane);
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.0f;
                if (lane == 31) sc[480 + chunk] = incl;
                tc::named_bar_sync(1 + slot, 128);
                float carry = 1.0f;
                for (int j = 0; j < chunk; ++j) carry *= sc[480 + j];
                const float w = alpha * (carry * excl);         // train.py:261-266
                if (chunk > 0) {
                  float* q = sc + (chunk - 1) * 160 + lane;
                  q[0] = w; q[32] = cr; q[64] = cg; q[96] = cb; q[128] = zz;
                }
                tc::named_bar_sync(1 + slot, 128);
                if (chunk == 0) {
                  float ar = w * cr, ag = w * cg, ab = w * cb, ad = w * zz, aa = w;
                  for (int k = 1; k < nc; ++k) {
                    const float* q = sc + (k - 1) * 160 + lane;
                    const float wk = q[0];
                    ar = fmaf(wk, q[32], ar); ag = fmaf(wk, q[64], ag); ab = fmaf(wk, q[96], ab); ad = fmaf(wk, q[128], ad); aa += wk;
                  }
                  const float r = warp_sum4_scatter(ar, ag, ab, ad, lane);
                  aa = warp_sum(aa);
                  if ((lane & 7) == 0 && tile_ok && ray * a.n_sample < a.n_rows) {
                    if (lane < 24) a.comp_rgb[3 * ray + (lane >> 3)] = r;               // train.py:269
                    else if (a.comp_depth) a.comp_depth[ray] = r;
                    if (lane == 0 && a.comp_acc) a.comp_acc[ray] = aa;
                  }
                }
              }
            }
            if (!composited && valid) reinterpret_cast<float4*>(a.sigma_rgb)[grow] = make_float4(sigma_keep, cr, cg, cb);
          }
        }
        if (MODE == MODE_FWD && TRAIN && store_out) {
          // Training stash: this thread reads back the 16-byte pieces of ITS OWN row of the bf16 tile it
          // just wrote and streams them to HBM (a warp writes 512 contiguous bytes per instruction), and
          // folds them into the ReLU bit masks (bit of column c in a 32-column chunk: (c>>1) + 16*(c&1)).
          // This runs under the other slot's MMAs.  (Storing from registers inside the read-out loop, as dgrad
          // does, was measured too: no gain for training and a slower inference loop.)  A bulk (TMA) store of
          // the tile would be one instruction, but its shared-memory reads halve the UMMA issue rate (measured).
          uint4* gdst = reinterpret_cast<uint4*>(stash_tile + st.stash_off) + row;
          uint32_t mw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          // NERF_READBACK_BATCH k-cores are loaded before the first is stored: ncu showed this loop waiting on one LDS at a time
          // (short-scoreboard stalls on the first use of every load; ~100 cycles each under the UMMA's operand traffic).
#pragma unroll
          for (int g = 0; g < 32 / NERF_READBACK_BATCH; ++g) {
            if (g * NERF_READBACK_BATCH * 8 < st.n) {
              uint4 pk[NERF_READBACK_BATCH];
#pragma unroll
              for (int q = 0; q < NERF_READBACK_BATCH; ++q)
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(pk[q].x), "=r"(pk[q].y), "=r"(pk[q].z), "=r"(pk[q].w)
                             : "r"(s_act + (g * NERF_READBACK_BATCH + q) * kKcoreBytes + row * 16));
#pragma unroll
              for (int q = 0; q < NERF_READBACK_BATCH; ++q) {
                const int kc = g * NERF_READBACK_BATCH + q;        // k-core: 8 channels; 4 k-cores per 32-column mask chunk
                __stcs(gdst + kc * 128, pk[q]);                    // streaming (evict-first) store: the hot weight images stay in L2
                mw[kc >> 2] |= ((pk[q].x + 0x7fff7fffu) & 0x80008000u) >> (15 - (4 * (kc & 3) + 0));
                mw[kc >> 2] |= ((pk[q].y + 0x7fff7fffu) & 0x80008000u) >> (15 - (4 * (kc & 3) + 1));
                mw[kc >> 2] |= ((pk[q].z + 0x7fff7fffu) & 0x80008000u) >> (15 - (4 * (kc & 3) + 2));
                mw[kc >> 2] |= ((pk[q].w + 0x7fff7fffu) & 0x80008000u) >> (15 - (4 * (kc & 3) + 3));
              }
            }
          }
          uint4* mp = reinterpret_cast<uint4*>(stash_tile + st.mask_off) + row * 2;
          __stcs(mp, make_uint4(mw[0], mw[1], mw[2], mw[3]));
          __stcs(mp + 1, make_uint4(mw[4], mw[5], mw[6], mw[7]));
        }
      }
    }
  }
  tc::tc_fence_before();
  tc::cluster_sync();                                  // no remote arrive / multicast commit may target a CTA that has exited
  if (warp == 8) tc::tmem_dealloc_pair(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// Weight packing: flat fp32 parameters -> bf16 UMMA images (fwd + transposed dgrad) + fp32 table.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_image_step(const float* __restrict__ params, const Step& st, uint8_t* __restrict__ packed) {
  const int total_k = (st.act_ksteps + st.enc_ksteps + st.bias_kstep) * 16;
  const int n_elems = total_k * st.n;
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(packed + st.w_off);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_elems; idx += gridDim.x * blockDim.x) {
    // idx enumerates the destination linearly: [kstep][half][kc][n_local][8] (half h = rows h*N/2.. of CTA rank h)
    const int j = idx & 7;
    const int nh = st.n >> 1;
    const int nl = (idx >> 3) % nh;
    const int r2 = (idx >> 3) / nh;                    // = (kstep*2 + half)*2 + kc
    const int n = ((r2 >> 1) & 1) * nh + nl;
    const int kk = ((r2 >> 2) * 2 + (r2 & 1)) * 8 + j;
    float v = 0.0f;
    if (n < st.p_rows) {
      if (st.p_transpose) {
        if (kk < st.act_cols) v = params[st.p_w + (int64_t)kk * st.p_in + n];
      } else {
        const int act_k = st.act_ksteps * 16, enc_k = st.enc_ksteps * 16;
        if (kk < act_k) { if (kk < st.act_cols) v = params[st.p_w + (int64_t)n * st.p_in + kk]; }
        else if (kk < act_k + enc_k) {
          const int e = kk - act_k;
          if (e < st.enc_cols) v = params[st.p_w + (int64_t)n * st.p_in + st.enc_col0 + e];
          else if (st.bias_in_enc && e == 63) v = params[st.p_bias + n];       // pairs with the constant-1 channel
        } else if ((kk - act_k - enc_k) == 15 && st.p_bias >= 0) {
          v = params[st.p_bias + n];                                           // bias k-step: ENC channel 63
        }
      }
    }
    img[idx] = __float2bfloat16_rn(v);
  }
}

struct PackSmallArgs {
  int32_t hidden;
  int32_t wd_src, bd_src, w2_src, b2_src;
  int32_t sm_wd, sm_bd, sm_w2, sm_b2;
};
__device__ __forceinline__ void pack_small(const float* __restrict__ params, const PackSmallArgs& p, float* __restrict__ small) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 256) small[p.sm_wd + i] = i < p.hidden ? params[p.wd_src + i] : 0.0f;
  if (i < 384) small[p.sm_w2 + i] = params[p.w2_src + i];
  if (i < 3) small[p.sm_b2 + i] = params[p.b2_src + i];
  if (i == 0) { small[p.sm_bd] = params[p.bd_src]; small[p.sm_bd + 1] = 0.f; small[p.sm_bd + 2] = 0.f; small[p.sm_bd + 3] = 0.f; small[p.sm_b2 + 3] = 0.f; }
}

// Everything the MLP kernels read, from the flat fp32 parameters, in one launch: blockIdx.y selects a forward step, a dgrad
// step, or (last row) the fp32 head table.  `bump` (optional) is the device-side iteration counter of nerf_train_step: this
// is the last launch of a step and does not read it, so one thread advances it here for the next step.
__global__ void pack_all_kernel(const float* __restrict__ params, const __grid_constant__ Program fwd, const __grid_constant__ Program bwd,
                                const PackSmallArgs small_args, uint8_t* __restrict__ packed, float* __restrict__ small,
                                unsigned long long* __restrict__ bump, unsigned long long* __restrict__ bump2) {
  const int y = blockIdx.y;
  if (bump != nullptr && blockIdx.x == 0 && y == 0 && threadIdx.x == 0) *bump += 1ull;
  if (bump2 != nullptr && blockIdx.x == 0 && y == 0 && threadIdx.x == 0) *bump2 += 1ull;       // gradient-exchange epoch (data parallel)
  if (y < fwd.n_steps) pack_image_step(params, fwd.s[y], packed);
  else if (y < fwd.n_steps + bwd.n_steps) pack_image_step(params, bwd.s[y - fwd.n_steps], packed);
  else if (blockIdx.x < 2) pack_small(params, small_args, small);
}

}  // namespace nerf
