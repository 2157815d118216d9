// bf16 "TN" GEMM, CTA-pair version:  C[M,N] = epilogue(A[M,K] . B[N,K]^T)  on tcgen05.mma.cta_group::2.
//
// Why pairs: the round-1 profile of the single-CTA 128x256 kernel (profiles/r01) showed the L2->SM fabric at its
// ~12 TB/s ceiling with the tensor pipe only ~60 % busy: 48 KB of operands per 128x256x64 MMA block is 85 FLOP/B.
// A CTA pair computes a 256x256 tile with UMMA M=256: each CTA stages its own 128 rows of A and HALF of B
// (32 KB per CTA per k-block, 131 FLOP/B), the tensor cores of both SMs read B from both shared memories.
//
// Cluster (2,1,1), persistent: cluster c walks tiles c, c+C, ... in n-fastest order (concurrent clusters share one
// A row-panel; the weight matrix B stays L2-resident).
//   warp 0      TMA producer (both CTAs): A[128x64] + B[128x64] per stage, SWIZZLE_128B, 6-stage ring; the
//               transaction bytes of BOTH CTAs complete on the leader's `full` barrier
//   warp 1      leader CTA only: one lane issues tcgen05.mma.cta_group::2 (M256 N256 K16 x4 per stage);
//               tcgen05.commit ... multicast::cluster frees the stage in both CTAs / publishes the accumulator
//   warps 2..9  epilogue (both CTAs, own 128 accumulator rows): residual rows prefetched before the accumulator is
//               ready, bias loads overlapped with tcgen05.ld, TMEM double-buffered (2 x 256 columns)
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace kbner {
namespace g2 {

constexpr int BM = 256, BN = 256, BK = 64, kStages = 6;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr uint32_t kABytes = 128 * BK * 2, kBBytes = 128 * BK * 2;   // per CTA per stage
constexpr uint32_t kTmemCols = 512;

struct Smem {
    uint8_t a[kStages][kABytes];
    uint8_t b[kStages][kBBytes];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

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
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA load: data lands in THIS CTA's smem, transaction bytes complete on the barrier at `bar_cluster_addr`
__device__ __forceinline__ void tma_load_2d_2sm(void *smem_dst, const CUtensorMap *map, uint32_t bar_cluster_addr,
                                                int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
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
__device__ __forceinline__ void mma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(ptx::smem_u32(bar)), "h"(mask)
                 : "memory");
}

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const float *__restrict__ bias, const uint16_t *__restrict__ resid, void *__restrict__ Cv,
                     int M, int N, int K, int ldc) {
    extern __shared__ uint8_t smem_raw[];
    Smem &s = *reinterpret_cast<Smem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        for (int i = 0; i < kStages; ++i) {
            ptx::mbar_init(&s.full[i], 1);
            ptx::mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&s.tmem_full[i], 1);
            ptx::mbar_init(&s.tmem_empty[i], 2 * kEpiWarps);   // epilogue warps of BOTH CTAs arrive on the leader's
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(&s.tmem_base);
    ptx::tc_fence_before();
    cluster_sync();            // barriers of the peer are initialised, TMEM allocated in both CTAs
    ptx::tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m_blk = tile / num_n, n_blk = tile % num_n;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&s.empty[stage], phase ^ 1);
                    if (leader) ptx::mbar_expect_tx(&s.full[stage], 2 * (kABytes + kBBytes));
                    const uint32_t full_leader = mapa(ptx::smem_u32(&s.full[stage]), 0);
                    tma_load_2d_2sm(s.a[stage], &tmA, full_leader, kb * BK, m_blk * BM + (int)rank * 128);
                    tma_load_2d_2sm(s.b[stage], &tmB, full_leader, kb * BK, n_blk * BN + (int)rank * 128);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA, one lane) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(&s.tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&s.full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(s.a[stage]);
                    const uint32_t b_addr = ptx::smem_u32(s.b[stage]);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = ptx::make_sw128_desc(a_addr + k * 32, 16, 1024);
                        const uint64_t db = ptx::make_sw128_desc(b_addr + k * 32, 16, 1024);
                        mma_f16_ss_2sm(d_tmem, da, db, idesc, (kb | k) != 0);
                    }
                    mma_commit_mc(&s.empty[stage], 0b11);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                mma_commit_mc(&s.tmem_full[acc], 0b11);
            }
        }
    } else {
        // ===================== epilogue (both CTAs; own 128 rows of the 256-row tile) =====================
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int row = m_blk * BM + (int)rank * 128 + quarter * 32 + lane;
            const bool row_ok = row < M;
            const int colbase = n_blk * BN + half * (BN / 2);
            // residual rows do not depend on the accumulator: fetch them while the main loop still runs
            uint4 rres[16];
            if (EPI == KBNER_EPI_BIAS_RESID_F32) {
                const uint16_t *rrow = resid + (size_t)(row_ok ? row : 0) * ldc + colbase;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    rres[i] = (row_ok && colbase + i * 8 < N) ? ld_nc_v4(rrow + i * 8) : make_uint4(0, 0, 0, 0);
            }
            ptx::mbar_wait(&s.tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
#pragma unroll
            for (int c = 0; c < (BN / 2) / 32; ++c) {
                const int col0 = colbase + c * 32;
                uint32_t r[32];
                const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN + half * (BN / 2) + c * 32;
                ptx::tmem_ld_32x32b_x32(taddr, r);
                float bv[32];
                if (EPI != KBNER_EPI_NONE_F32) {       // bias loads overlap the TMEM load
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (col0 + i < N) t = __ldg(reinterpret_cast<const float4 *>(bias + col0 + i));
                        bv[i] = t.x; bv[i + 1] = t.y; bv[i + 2] = t.z; bv[i + 3] = t.w;
                    }
                }
                ptx::tmem_ld_wait();
                if (col0 < N) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    if (EPI != KBNER_EPI_NONE_F32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += bv[i];
                    }
                    if (EPI == KBNER_EPI_BIAS_GELU) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
                    }
                    if (row_ok) {
                        if (EPI == KBNER_EPI_BIAS || EPI == KBNER_EPI_BIAS_GELU) {
                            uint16_t *crow = reinterpret_cast<uint16_t *>(Cv) + (size_t)row * ldc + col0;
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                if (col0 + i < N) {
                                    uint4 o;
                                    o.x = pack_bf16x2(v[i], v[i + 1]);
                                    o.y = pack_bf16x2(v[i + 2], v[i + 3]);
                                    o.z = pack_bf16x2(v[i + 4], v[i + 5]);
                                    o.w = pack_bf16x2(v[i + 6], v[i + 7]);
                                    *reinterpret_cast<uint4 *>(crow + i) = o;
                                }
                            }
                        } else {
                            if (EPI == KBNER_EPI_BIAS_RESID_F32) {
#pragma unroll
                                for (int i = 0; i < 32; i += 8) {
                                    const uint4 rv = rres[c * 4 + i / 8];
                                    float a0, a1;
                                    unpack_bf16x2(rv.x, a0, a1); v[i] += a0; v[i + 1] += a1;
                                    unpack_bf16x2(rv.y, a0, a1); v[i + 2] += a0; v[i + 3] += a1;
                                    unpack_bf16x2(rv.z, a0, a1); v[i + 4] += a0; v[i + 5] += a1;
                                    unpack_bf16x2(rv.w, a0, a1); v[i + 6] += a0; v[i + 7] += a1;
                                }
                            }
                            float *crow = reinterpret_cast<float *>(Cv) + (size_t)row * ldc + col0;
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                if (col0 + i < N)
                                    *reinterpret_cast<float4 *>(crow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                            }
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa(ptx::smem_u32(&s.tmem_empty[acc]), 0));
        }
    }
    ptx::tc_fence_before();
    cluster_sync();            // nobody leaves while the peer may still touch this CTA's smem / barriers / TMEM
    if (warp == 1) {
        ptx::tc_fence_after();
        tmem_dealloc_2sm(tmem_base);
    }
}

template <int EPI>
static int launch(const CUtensorMap &tmA, const CUtensorMap &tmB, const float *bias, const uint16_t *resid, void *C,
                  int M, int N, int K, int ldc, cudaStream_t st) {
    const size_t smem = sizeof(Smem) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_tn_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) {
            set_error("gemm2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return KBNER_ECUDA;
        }
        configured = true;
    }
    const int num_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int clusters = num_tiles < kNumSMs / 2 ? num_tiles : kNumSMs / 2;
    gemm2_bf16_tn_kernel<EPI><<<clusters * 2, kThreads, smem, st>>>(tmA, tmB, bias, resid, C, M, N, K, ldc);
    KBNER_CHECK_LAUNCH("gemm2_bf16_tn");
    return KBNER_OK;
}

}  // namespace g2

int gemm2_dispatch(const uint16_t *A, const uint16_t *B, const float *bias, const uint16_t *residual, void *C, int M,
                   int N, int K, int lda, int ldb, int ldc, int epilogue, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    int rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, g2::BK);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 128, g2::BK);
    if (rc) return rc;
    switch (epilogue) {
        case KBNER_EPI_BIAS: return g2::launch<KBNER_EPI_BIAS>(tmA, tmB, bias, residual, C, M, N, K, ldc, st);
        case KBNER_EPI_BIAS_GELU: return g2::launch<KBNER_EPI_BIAS_GELU>(tmA, tmB, bias, residual, C, M, N, K, ldc, st);
        case KBNER_EPI_BIAS_RESID_F32:
            return g2::launch<KBNER_EPI_BIAS_RESID_F32>(tmA, tmB, bias, residual, C, M, N, K, ldc, st);
        case KBNER_EPI_NONE_F32: return g2::launch<KBNER_EPI_NONE_F32>(tmA, tmB, bias, residual, C, M, N, K, ldc, st);
        default: set_error("gemm: unknown epilogue %d", epilogue); return KBNER_EINVAL;
    }
}

}  // namespace kbner
