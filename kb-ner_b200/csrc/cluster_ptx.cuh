// Inline-PTX wrappers for the CTA-pair (cta_group::2) and thread-block-cluster side of the tcgen05 datapath:
// cluster rank / barrier / mapa, remote mbarrier arrives, 2-SM TMA loads, 2-SM TMEM allocation, tcgen05.mma.cta_group::2
// and its multicast commit.  Shared by gemm_tcgen05.cu and gemm_ln_tcgen05.cu.
#pragma once
#include "tc_ptx.cuh"

namespace kbner {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
// Remote arrive WITHOUT release semantics: what it orders is the drained accumulator, and those tcgen05.ld's have
// already completed (tcgen05.wait::ld) and are fenced by tcgen05.fence::before_thread_sync.  The default .release form
// compiles to MEMBAR.ALL.CTA + ERRBAR, which waited for every outstanding global load of the epilogue warp
// (8 % of the stall samples in profiles/r01/gemm_attnout_ncu.txt).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA load: data lands in THIS CTA's smem, transaction bytes complete on the barrier at `bar_cluster_addr`
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap *map, uint32_t bar_cluster_addr,
                                                int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
template <uint32_t kTmemCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kTmemCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kTmemCols) : "memory");
}
__device__ __forceinline__ void mma_f16_ss_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once) on the barrier at the same smem offset in every CTA of `mask` when all prior MMAs have retired
__device__ __forceinline__ void mma_commit_mc(uint32_t bar_addr, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar_addr), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint64_t pack_desc(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}

}  // namespace kbner
