// dy_ptx.cuh — inline-PTX wrappers for the Blackwell pieces the conv kernel uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences).
#pragma once
#include <cstdint>
#include <cuda.h>

namespace dy { namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kills the context) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("dy: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// Same operations on raw 32-bit shared-window addresses: lets the single-thread producer / MMA loops keep every
// barrier and stage address in a register instead of re-deriving it per iteration.
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_a(bar, parity)) {
    if (++spins > (1u << 26)) __trap();      // protocol bug: kill the context instead of hanging the GPU box
  }
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d_a(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_a(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- lane-predicated forms ---------------------------------------------------------------------
// The producer / MMA loops run WARP-UNIFORM (all 32 lanes execute the loop, so ptxas keeps stage counters, barrier
// addresses and descriptors in uniform registers instead of converting per-thread values with R2UR on every use); only
// the instructions with side effects are predicated on the elected lane.
__device__ __forceinline__ void mbar_arrive_expect_tx_p(uint32_t lead, uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred L;\n\tsetp.ne.b32 L, %0, 0;\n\t@L mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %2;\n\t}"
               ::"r"(lead), "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d_p(uint32_t lead, uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n\t.reg .pred L;\n\tsetp.ne.b32 L, %0, 0;\n\t"
      "@L cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%2, {%4, %5, %6, %7}], [%3];\n\t}"
      ::"r"(lead), "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_p(uint32_t lead, uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "{\n\t.reg .pred L;\n\tsetp.ne.b32 L, %0, 0;\n\t"
      "@L cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%2, {%4, %5, %6}], [%3];\n\t}"
      ::"r"(lead), "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint32_t lead, uint32_t bar) {
  asm volatile("{\n\t.reg .pred L;\n\tsetp.ne.b32 L, %0, 0;\n\t"
               "@L tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t}" ::"r"(lead), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_p(uint32_t lead, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred L, p;\n\tsetp.ne.b32 L, %0, 0;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "@L tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t}"
      ::"r"(lead), "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// One 64-channel k-block (4 x K16), elected lane only, with the NEXT barrier's try_wait (all lanes) issued first: its
// ~200-cycle latency overlaps the MMA issue.  Returns whether that barrier phase had already completed.
__device__ __forceinline__ bool umma_bf16_ss_x4_waitahead_p(uint32_t lead, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                            uint32_t accumulate_first, uint32_t next_bar, uint32_t next_parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred L, p, q, t;\n\t.reg .b64 a, b;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
      "setp.ne.b32 L, %8, 0;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.ne.b32 t, %4, 0;\n\t"                       // always true (the instruction descriptor is never 0)
      "@L tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
      "add.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
      "@L tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "add.u64 a, %2, 4;\n\tadd.u64 b, %3, 4;\n\t"
      "@L tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "add.u64 a, %2, 6;\n\tadd.u64 b, %3, 6;\n\t"
      "@L tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(ok) : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(next_bar), "r"(next_parity), "r"(lead) : "memory");
  return ok != 0;
}

// ---- TMA ----------------------------------------------------------------------------------------
// 1-D bulk copy global -> shared (src, dst and bytes multiples of 16), completion counted on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// make generic-proxy smem writes visible to the async proxy (TMA) before a bulk store reads them
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 16-byte shared-memory accesses on 32-bit shared-window addresses (generic pointers make ptxas emit LD.E/ST.E)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tma_store_4d_a(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// Programmatic dependent launch: wait for the upstream grid's memory / let the downstream grid start its prologue
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- tcgen05 ------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// One 64-channel k-block (4 x K16) with the NEXT barrier's try_wait issued first: its ~200-cycle latency overlaps the MMA
// issue instead of stalling the (in-order) issuing thread.  Returns whether that barrier phase had already completed.
__device__ __forceinline__ bool umma_bf16_ss_x4_waitahead(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                          uint32_t accumulate_first, uint32_t next_bar, uint32_t next_parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.ne.b32 t, %4, 0;\n\t"                       // always true (the instruction descriptor is never 0)
      "tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
      "add.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "add.u64 a, %2, 4;\n\tadd.u64 b, %3, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "add.u64 a, %2, 6;\n\tadd.u64 b, %3, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(ok) : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(next_bar), "r"(next_parity) : "memory");
  return ok != 0;
}
// Two k-blocks (8 x K16) of one pipeline stage: A blocks 16 KB apart, B blocks `bstep` (16-byte units) apart.
__device__ __forceinline__ bool umma_bf16_ss_x8_waitahead(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t bstep, uint32_t idesc,
                                                          uint32_t accumulate_first, uint32_t next_bar, uint32_t next_parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b, b1, st;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.ne.b32 t, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
      "add.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "add.u64 a, %2, 4;\n\tadd.u64 b, %3, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "add.u64 a, %2, 6;\n\tadd.u64 b, %3, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "cvt.u64.u32 st, %8;\n\tadd.u64 b1, %3, st;\n\t"
      "add.u64 a, %2, 1024;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b1, %4, t;\n\t"
      "add.u64 a, %2, 1026;\n\tadd.u64 b, b1, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "add.u64 a, %2, 1028;\n\tadd.u64 b, b1, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "add.u64 a, %2, 1030;\n\tadd.u64 b, b1, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(ok) : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(next_bar), "r"(next_parity), "r"(bstep) : "memory");
  return ok != 0;
}
// Three 32-channel k-blocks (64-byte rows: 2 x K16 each) of one pipeline stage: A blocks 8 KB apart, B blocks `bstep` apart.
__device__ __forceinline__ bool umma_bf16_ss_k32x3_waitahead(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t bstep, uint32_t idesc,
                                                             uint32_t accumulate_first, uint32_t next_bar, uint32_t next_parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b, b1, b2, st;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.ne.b32 t, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
      "add.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "cvt.u64.u32 st, %8;\n\tadd.u64 b1, %3, st;\n\tadd.u64 b2, b1, st;\n\t"
      "add.u64 a, %2, 512;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b1, %4, t;\n\t"
      "add.u64 a, %2, 514;\n\tadd.u64 b, b1, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "add.u64 a, %2, 1024;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b2, %4, t;\n\t"
      "add.u64 a, %2, 1026;\n\tadd.u64 b, b2, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(ok) : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "r"(next_bar), "r"(next_parity), "r"(bstep) : "memory");
  return ok != 0;
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane+i), columns c..c+15.
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, 128B swizzle, rows of 128 bytes
// (64 bf16), 8-row groups `sbo_bytes` apart.  Bit layout (cute/arch/mma_sm100_desc.hpp):
// [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [49,52) base offset | [61,64) layout (2 = SW128)
// layout: 2 = 128B swizzle (128-byte rows), 4 = 64B swizzle (64-byte rows)
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout & 7u) << 61;
  return d;
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  return umma_desc_kmajor(smem_addr, sbo_bytes, 2u);
}
// Instruction descriptor for kind::f16, A/B = bf16 K-major, D = fp32, M x N tile.
// [4,6) c_format=1 (f32) | [7,10) a_format=1 (bf16) | [10,13) b_format=1 | [15] a_major=0 | [16] b_major=0
// | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ inline uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

}}  // namespace dy::ptx
