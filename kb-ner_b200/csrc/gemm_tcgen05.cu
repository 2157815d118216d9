// bf16 "TN" GEMM on the 5th-gen tensor cores:  C[M,N] = epilogue(A[M,K] . B[N,K]^T).
// This is every dense projection of the XLM-R encoder the reference runs through
// transformers' torch.nn.Linear (call site /root/reference/flair/embeddings.py:3269;
// SURVEY.md E2 QKV, E4 attention-out, E5 FFN-up + GELU, E6 FFN-down).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer : cp.async.bulk.tensor 128x64 (A) + 256x64 (B) bf16 tiles,
//                              SWIZZLE_128B, 4-stage mbarrier ring
//   warp 1      MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                              (M=128, N=256, K=16) x4 per stage, accumulators in TMEM,
//                              2 accumulator buffers (2 x 256 columns) so the epilogue of
//                              tile i overlaps the main loop of tile i+1
//   warps 2..9  epilogue     : tcgen05.ld 32 lanes x 32 columns -> registers -> bias /
//                              GELU(erf) / residual -> bf16 or fp32 -> global
// Tiles are walked m-fastest so the 148 concurrently running CTAs share one B (weight) tile.
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace kbner {

constexpr int BM = 128, BN = 256, BK = 64, kStages = 4;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr uint32_t kABytes = BM * BK * 2, kBBytes = BN * BK * 2;
constexpr uint32_t kTmemCols = 512;

struct GemmSmem {
    // tiles first: SWIZZLE_128B needs 1024-B alignment (every tile size is a multiple of 1024)
    uint8_t a[kStages][kABytes];
    uint8_t b[kStages][kBBytes];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float gelu_erf(float x) {
    // HF "gelu": x * 0.5 * (1 + erf(x / sqrt(2)))
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const float *__restrict__ bias, const uint16_t *__restrict__ resid, void *__restrict__ Cv,
                    int M, int N, int K, int ldc) {
    extern __shared__ uint8_t smem_raw[];
    GemmSmem &s = *reinterpret_cast<GemmSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
            ptx::mbar_init(&s.tmem_empty[i], kEpiWarps);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<kTmemCols>(&s.tmem_base);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % num_m, n_blk = tile / num_m;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&s.empty[stage], phase ^ 1);
                    ptx::mbar_expect_tx(&s.full[stage], kABytes + kBBytes);
                    ptx::tma_load_2d(s.a[stage], &tmA, &s.full[stage], kb * BK, m_blk * BM);
                    ptx::tma_load_2d(s.b[stage], &tmB, &s.full[stage], kb * BK, n_blk * BN);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
                        ptx::mma_f16_ss(d_tmem, da, db, idesc, (kb | k) != 0);
                    }
                    ptx::mma_commit(&s.empty[stage]);      // smem slot free once these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                ptx::mma_commit(&s.tmem_full[acc]);         // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue =====================
        const int ew = warp - 2;                 // 0..7
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int half = ew >> 2;                // which 128 of the 256 accumulator columns
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            ptx::mbar_wait(&s.tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const int row = m_blk * BM + quarter * 32 + lane;
            const bool row_ok = row < M;
#pragma unroll 1
            for (int c = 0; c < (BN / 2) / 32; ++c) {
                const int col0 = n_blk * BN + half * (BN / 2) + c * 32;
                uint32_t r[32];
                const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN + half * (BN / 2) + c * 32;
                ptx::tmem_ld_32x32b_x32(taddr, r);
                ptx::tmem_ld_wait();
                if (col0 < N) {      // warp-uniform
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    if (EPI != KBNER_EPI_NONE_F32) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            if (col0 + i < N) {
                                const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + col0 + i));
                                v[i] += bv.x; v[i + 1] += bv.y; v[i + 2] += bv.z; v[i + 3] += bv.w;
                            }
                        }
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
                                const uint16_t *rrow = resid + (size_t)row * ldc + col0;
#pragma unroll
                                for (int i = 0; i < 32; i += 8) {
                                    if (col0 + i < N) {
                                        const uint4 rv = *reinterpret_cast<const uint4 *>(rrow + i);
                                        float a0, a1;
                                        unpack_bf16x2(rv.x, a0, a1); v[i] += a0; v[i + 1] += a1;
                                        unpack_bf16x2(rv.y, a0, a1); v[i + 2] += a0; v[i + 3] += a1;
                                        unpack_bf16x2(rv.z, a0, a1); v[i + 4] += a0; v[i + 5] += a1;
                                        unpack_bf16x2(rv.w, a0, a1); v[i + 6] += a0; v[i + 7] += a1;
                                    }
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
            // this warp has drained its part of the accumulator
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&s.tmem_empty[acc]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kTmemCols>(tmem_base);
    }
}

template <int EPI>
static int launch_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const float *bias, const uint16_t *resid,
                       void *C, int M, int N, int K, int ldc, cudaStream_t st) {
    const size_t smem = sizeof(GemmSmem) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) {
            set_error("gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return KBNER_ECUDA;
        }
        configured = true;
    }
    const int num_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
    gemm_bf16_tn_kernel<EPI><<<grid, kGemmThreads, smem, st>>>(tmA, tmB, bias, resid, C, M, N, K, ldc);
    KBNER_CHECK_LAUNCH("gemm_bf16_tn");
    return KBNER_OK;
}

int gemm2_dispatch(const uint16_t *A, const uint16_t *B, const float *bias, const uint16_t *residual, void *C, int M,
                   int N, int K, int lda, int ldb, int ldc, int epilogue, cudaStream_t st);   // gemm2_tcgen05.cu

}  // namespace kbner

using namespace kbner;

static int gemm_impl_choice() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("KBNER_GEMM");
        v = (e && e[0] == '2') ? 2 : 1;
    }
    return v;
}

extern "C" int kbner_gemm_bf16_tn(const uint16_t *A, const uint16_t *B, const float *bias,
                                  const uint16_t *residual, void *C, int M, int N, int K, int lda, int ldb,
                                  int ldc, int epilogue, void *stream) {
    KBNER_CHECK_ARG(A && B && C, "gemm: null pointer");
    KBNER_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    KBNER_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0,
                    "gemm: N, K and leading dimensions must be multiples of 8 (N=%d K=%d lda=%d ldb=%d ldc=%d)", N, K,
                    lda, ldb, ldc);
    KBNER_CHECK_ARG(epilogue == KBNER_EPI_NONE_F32 || bias, "gemm: epilogue %d needs a bias", epilogue);
    KBNER_CHECK_ARG(epilogue != KBNER_EPI_BIAS_RESID_F32 || residual, "gemm: residual epilogue without residual");
    KBNER_CHECK_ARG(((uintptr_t)C & 15u) == 0 && (!bias || ((uintptr_t)bias & 15u) == 0) &&
                        (!residual || ((uintptr_t)residual & 15u) == 0),
                    "gemm: C / bias / residual must be 16-byte aligned");
    if (gemm_impl_choice() == 2)
        return gemm2_dispatch(A, B, bias, residual, C, M, N, K, lda, ldb, ldc, epilogue, (cudaStream_t)stream);
    CUtensorMap tmA, tmB;
    int rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN, BK);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (epilogue) {
        case KBNER_EPI_BIAS: return launch_gemm<KBNER_EPI_BIAS>(tmA, tmB, bias, residual, C, M, N, K, ldc, st);
        case KBNER_EPI_BIAS_GELU: return launch_gemm<KBNER_EPI_BIAS_GELU>(tmA, tmB, bias, residual, C, M, N, K, ldc, st);
        case KBNER_EPI_BIAS_RESID_F32:
            return launch_gemm<KBNER_EPI_BIAS_RESID_F32>(tmA, tmB, bias, residual, C, M, N, K, ldc, st);
        case KBNER_EPI_NONE_F32: return launch_gemm<KBNER_EPI_NONE_F32>(tmA, tmB, bias, residual, C, M, N, K, ldc, st);
        default: set_error("gemm: unknown epilogue %d", epilogue); return KBNER_EINVAL;
    }
}
