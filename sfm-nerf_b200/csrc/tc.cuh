// tc.cuh -- hand-written PTX wrappers for sm_100a: mbarrier, 1-D bulk (TMA) copies,
// tcgen05 (UMMA) MMA / TMEM alloc / TMEM load, proxy fences.  No CUTLASS/CuTe.
#pragma once
#include "common.cuh"

namespace nerf {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---------------------------------------------------------------- mbarrier ----------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (the launch then reports an error) instead of hanging
// the GPU box.  ~4e9 cycles is seconds; real waits are microseconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && (clock64() - t0) > 4000000000LL) {
      printf("nerf_b200: mbarrier timeout block %d thread %d bar %u parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------- CTA pair (cluster of 2) --
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {        // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> the same offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on an mbarrier of another CTA of the cluster (address from mapa).  Default semantics (release at CTA
// scope), as CUTLASS's ClusterBarrier does: the .release.cluster form carries a cluster-wide fence that costs
// ~600 cycles per arrive (measured: it throttled the slab relay to one k-step per 635 cycles).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// NON-BLOCKING probe (mbarrier.test_wait): try_wait may suspend the thread up to a system time limit when the phase is not
// complete (measured: an early try_wait on a barrier thousands of cycles from completing held the MMA warp for ~3,000 cycles),
// so anything probed AHEAD of need -- the next pass's A tile, the second half of a trip -- must use this form.
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// wait on a local mbarrier that CTAs of the whole cluster arrive on (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && (clock64() - t0) > 4000000000LL) {
      printf("nerf_b200: cluster mbarrier timeout block %d thread %d bar %u parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------- bulk async copies --------
// global -> shared, completion signalled on an mbarrier (complete_tx::bytes).  16 B aligned.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy): evict_last keeps re-read data (the weight images) resident while
// GBs of write-once stash traffic stream through the same L2
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// 2-D tiled (tensor-map) copy global -> this CTA's shared memory whose completion is signalled on an mbarrier of EITHER CTA of the
// pair (cta_group::2; `bar_cluster` = shared::cluster address, e.g. mapa(bar, 0) for the leader's): the peer's half of a weight slab
// then reports straight to the leader's full barrier, with no relay warp in between.  `tmap` = generic address of the CUtensorMap.
__device__ __forceinline__ void tma_g2s_2d_pair(uint32_t dst_smem, uint64_t tmap, int32_t c0, int32_t c1, uint32_t bar_cluster, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst_smem), "l"(tmap), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// true in exactly one (converged) lane of the warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// ---------------------------------------------------------------- TMEM ---------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t n_cols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(n_cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t n_cols) {    // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(n_cols) : "memory");
}
// CTA-pair allocation: one warp of EACH CTA of the pair executes it (and later the dealloc)
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t n_cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(n_cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t n_cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(n_cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of fp32: thread t of the warp receives lane (32*(warp%4)+t), columns
// col..col+31.  taddr = (lane << 16) | column.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}

// 32 lanes x 64 columns (two accumulator chunks per instruction: half as many waits per layer)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr) : "memory");
}

// ---------------------------------------------------------------- UMMA ---------------------
// Shared-memory matrix descriptor, no-swizzle ("interleave") canonical layouts, 16-bit types:
//   K-major : 8 rows x 16 B core matrices; LBO = byte stride between the two K-adjacent core
//             matrices of one K=16 step, SBO = byte stride between 8-row groups along M/N.
//   MN-major: core matrix = 8 k-rows x 16 B (8 MN-contiguous elements); LBO = byte stride
//             between groups of 8 k, SBO = byte stride between 8-element MN chunks.
// Bits: [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) swizzle=0.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32 accumulate.
//   [4,6) D fmt (1=f32), [7,10) A fmt (1=bf16), [10,13) B fmt, 15 A major (1=MN), 16 B major,
//   [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Eight consecutive contraction steps into one accumulator from ONE asm block (single-CTA form): the sixteen descriptors
// are live inputs in distinct registers, so the UTCHMMAs do not serialise on each other's descriptor-register scoreboard
// (see umma_bf16_pair_x4_release2).  a[i] / b[i] are full 64-bit shared-memory descriptors.
__device__ __forceinline__ void umma_bf16_x8(uint32_t d_tmem, const uint64_t (&a)[8], const uint64_t (&b)[8], uint32_t idesc,
                                             uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t"
      "setp.ne.b32 p, %18, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %9, %17, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %10, %17, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %11, %17, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %4, %12, %17, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %5, %13, %17, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %6, %14, %17, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %7, %15, %17, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %8, %16, %17, t;\n\t}"
      ::"r"(d_tmem), "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "l"(a[4]), "l"(a[5]), "l"(a[6]), "l"(a[7]),
        "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]), "l"(b[4]), "l"(b[5]), "l"(b[6]), "l"(b[7]), "r"(idesc), "r"(accumulate_first) : "memory");
}
// mbarrier arrive when all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// CTA-pair forms (issued by ONE thread of the leader CTA, rank 0): D[256 x N] = A[256 x 16] . B[N x 16]^T with
// rows 0-127 of A and D in the leader, rows 128-255 in the peer, B rows 0..N/2-1 in the leader's shared memory
// and N/2..N-1 in the peer's -- all at the SAME shared-memory / TMEM offsets in both CTAs.
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Four consecutive k-steps into one accumulator from ONE asm block: all eight descriptors are live inputs, so they sit
// in distinct (uniform) registers.  Issued one by one, consecutive MMAs reuse the same uniform registers and each
// waits on the previous UTCHMMA's scoreboard, which is released only about when that MMA completes (~225 cycles per
// MMA instead of 128, measured).
// The two ring releases are inside the block: the commit for the first slab pair goes out right behind the second MMA (its slot
// returns to the loaders two MMA-times earlier than with both commits behind the fourth), the commit for the second pair behind the
// fourth.  Both are multicast to the two CTAs of the pair (mask 3).
__device__ __forceinline__ void umma_bf16_pair_x4_release2(uint32_t d_tmem, uint64_t a0, uint64_t a1, uint64_t a2, uint64_t a3, uint64_t b0,
                                                           uint64_t b1, uint64_t b2, uint64_t b3, uint32_t idesc, uint32_t accumulate_first,
                                                           uint32_t bar0, uint32_t bar1) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t.reg .b16 m;\n\t"
      "setp.ne.b32 p, %10, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "mov.b16 m, 3;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %5, %9, p;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %2, %6, %9, t;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%11], m;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %3, %7, %9, t;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %4, %8, %9, t;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%12], m;\n\t}"
      ::"r"(d_tmem), "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(accumulate_first),
        "r"(bar0), "r"(bar1) : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `cta_mask` once all previously issued pair MMAs completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}

// ---------------------------------------------------------------- misc ---------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {   // lo -> bits [0,16)
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

}  // namespace tc
}  // namespace nerf
