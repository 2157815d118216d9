// bf16 GEMM on the 5th-gen tensor cores, CTA-pair version (tcgen05.mma.cta_group::2, UMMA M=256 N=256 K=16):
//     C[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
// This is every dense projection of the XLM-R encoder the reference runs through transformers' torch.nn.Linear
// (call site /root/reference/flair/embeddings.py:3269; SURVEY.md E2 QKV, E4 attention-out, E5 FFN-up + GELU,
// E6 FFN-down) and, with the operand-layout flags, their dgrad / wgrad in the fine-tuning step.
//
// Operand layouts (runtime flags; the tensor core consumes both straight from row-major global tensors):
//   K-major  : operand stored [rows = M or N][cols = K]   (activations, torch Linear weights in the forward)
//   MN-major : operand stored [rows = K][cols = M or N]   (dY / X in wgrad, W in dgrad: no transposes are materialised)
//
// Why CTA pairs: the single-CTA 128x256 kernel of the first profile (profiles/r01) moved 48 KB of operands per
// 128x256x64 block (85 FLOP/B); a pair computes 256x256 with each CTA staging its own 128 rows of A and HALF of B
// (32 KB per CTA per k-block, 131 FLOP/B) and the tensor cores of both SMs read B from both shared memories.
//
// Cluster (2,1,1), persistent; cluster c walks tiles c, c+C, ... in n-fastest order (concurrent clusters share one A
// row panel, B stays L2-resident).  Warp roles -- every role branch is WARP-UNIFORM and only the instruction that
// must be issued once is under elect.sync (a lane-0 branch made ptxas wrap every UTCHMMA in an ELECT / R2UR loop,
// ~100 instructions per k-block, which capped the tensor pipe at ~63 %):
//   warp 0      TMA producer (both CTAs): A + B stage tiles, SWIZZLE_128B, 6-stage ring; transaction bytes of BOTH
//               CTAs complete on the leader's `full` barrier
//   warp 1      MMA issuer (leader CTA): 4 x tcgen05.mma per stage; tcgen05.commit multicast frees the stage in both
//               CTAs / publishes the accumulator
//   warps 2..9  epilogue (both CTAs, own 128 accumulator rows): operand rows (residual / GELU input) prefetched before
//               the accumulator is ready, bias loads overlapped with tcgen05.ld, TMEM double-buffered (2 x 256 cols)
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "cluster_ptx.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace kbner {

// Tile width TN (template): 256.  The 128-wide instantiation exists to measure whether finer tiles pay where 256-wide ones
// quantise badly onto the 74 CTA pairs (M = 4096 x N = 4096: 3.46 rounds -> 4 against 6.92 half-rounds -> 7): they do not
// (pick_tile_n below).
constexpr int BM = 256, BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr uint32_t kABytes = 128 * BK * 2;                           // per CTA per stage
constexpr uint32_t kTmemCols = 512;

template <int TN>
struct GemmSmemT {
    static constexpr int kStages = TN == 256 ? 5 : 6;
    static constexpr uint32_t kBBytes = (TN / 2) * BK * 2;           // per CTA per stage: half of the pair's B tile
    uint8_t a[kStages][kABytes];
    uint8_t b[kStages][kBBytes];
    uint8_t cstage[kEpiWarps][2][4096];     // per epilogue warp: 2 x (32 rows x 128 B) SWIZZLE_128B store staging
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint64_t aux_full[kEpiWarps][2];        // DGELU: the warp's aux (pre-activation) chunk has landed in cstage[warp][b]
    uint32_t tmem_base;
};

// HF "gelu" = x * Phi(x), Phi the standard normal CDF (transformers' erf GELU, not the tanh form).  Evaluated from the
// upper-tail probability Q(u) = Phi(-u) = 2^p(u), u = min(|x|, 6), p a degree-6 polynomial fitted to log2 Q on [0, 6]
// (relative error of Q <= 4.8e-5 with fp32 Horner; fit script: scripts/fit_gelu_tail.py):
//     gelu(x)  = max(x, 0) - u * Q(u)                       (|error| <= 6.7e-6 absolute, far below the bf16 output ulp)
//     gelu'(x) = Phi(x) + x * phi(x),  Phi(x) = x >= 0 ? 1 - Q(u) : Q(u),  phi(x) = 2^(-x^2 log2(e) / 2) / sqrt(2 pi)
// on PAIRS of columns: the polynomial runs as packed fp32 FMAs (FFMA2), one MUFU.EX2 per element (two for the gradient).
// The first version (Abramowitz-Stegun 7.1.26 on rcp.approx + ex2.approx, scalar) cost ~17 issue slots and 2 MUFU per
// element and made the FFN-up epilogue longer than its main loop: 33.2 us at 4096 x 4096 x 1024 against 26.4 us with the
// bias-only epilogue (profiles/r02/gemm_tile_width.json).
__device__ __forceinline__ void gelu_tail2(float x0, float x1, float &u0, float &u1, float &q0, float &q1) {
    u0 = fminf(fabsf(x0), 6.0f);
    u1 = fminf(fabsf(x1), 6.0f);
    const uint64_t u = f2_pack(u0, u1);
    uint64_t p = f2_fma(u, f2_pack(2.2990062444007588e-05f, 2.2990062444007588e-05f), f2_pack(-0.0006111000751876989f, -0.0006111000751876989f));
    p = f2_fma(p, u, f2_pack(0.007195565982293077f, 0.007195565982293077f));
    p = f2_fma(p, u, f2_pack(-0.05118533844324097f, -0.05118533844324097f));
    p = f2_fma(p, u, f2_pack(-0.46127192160306274f, -0.46127192160306274f));
    p = f2_fma(p, u, f2_pack(-1.150174258384186f, -1.150174258384186f));
    p = f2_fma(p, u, f2_pack(-1.0000647888776388f, -1.0000647888776388f));
    float p0, p1;
    f2_unpack(p, p0, p1);
    q0 = ex2_fast(p0);
    q1 = ex2_fast(p1);
}
__device__ __forceinline__ void gelu_erf2(float &x0, float &x1) {
    float u0, u1, q0, q1;
    gelu_tail2(x0, x1, u0, u1, q0, q1);
    x0 = fmaf(-u0, q0, fmaxf(x0, 0.0f));
    x1 = fmaf(-u1, q1, fmaxf(x1, 0.0f));
}
// (d0, d1) *= gelu'(x0), gelu'(x1)
__device__ __forceinline__ void gelu_grad_mul2(float &d0, float &d1, float x0, float x1) {
    float u0, u1, q0, q1;
    gelu_tail2(x0, x1, u0, u1, q0, q1);
    const uint64_t x = f2_pack(x0, x1);
    float t0, t1;
    f2_unpack(f2_mul(f2_mul(x, f2_pack(-0.7213475204444817f, -0.7213475204444817f)), x), t0, t1);
    const float e0 = ex2_fast(t0), e1 = ex2_fast(t1);
    const float c0 = x0 >= 0.0f ? 1.0f - q0 : q0, c1 = x1 >= 0.0f ? 1.0f - q1 : q1;
    float g0, g1;
    f2_unpack(f2_fma(f2_mul(x, f2_pack(0.3989422804014327f, 0.3989422804014327f)), f2_pack(e0, e1), f2_pack(c0, c1)), g0, g1);
    d0 *= g0;
    d1 *= g1;
}

struct GemmArgs {
    const float *bias;        // [N] or NULL
    const uint16_t *aux;      // bf16 [M,N]: residual (RESID), saved pre-activation (DGELU); NULL otherwise
    uint16_t *aux_out;        // bf16 [M,N]: pre-activation written by BIAS_GELU when non-NULL (training forward)
    void *C;
    int M, N, K, ldc;
    int a_mn, b_mn;           // operand layouts (0 = K-major, 1 = MN-major)
    int kb_per_split;         // k-blocks per work item (split-K, ACCUM epilogue only; otherwise all of K)
    int stream_units;         // > 0 (ACCUM epilogue only): stream-K -- cluster c owns k-block units [c * U, (c + 1) * U) of the
                              // linearised (tile, k-block) space, every intersection with a tile is one work item
};

// The work items of one cluster, in the order all three warp roles walk them.  Tile-per-item / split-K: item w =
// cluster_id + n * num_clusters -> (tile = w % num_tiles, split = w / num_tiles).  Stream-K: the cluster's unit range cut at
// tile boundaries.  `cursor` is the role's private position (items done, or units done).
struct GemmItem {
    int tile, kb0, kb1;
};
__device__ __forceinline__ bool gemm_next_item(const GemmArgs &g, int num_tiles, int num_kb, int cluster_id, int num_clusters,
                                               int &cursor, GemmItem &it) {
    if (g.stream_units > 0) {
        const int total = num_tiles * num_kb;
        const int begin = cluster_id * g.stream_units;
        const int end = min(begin + g.stream_units, total);
        const int pos = begin + cursor;
        if (pos >= end) return false;
        it.tile = pos / num_kb;
        it.kb0 = pos - it.tile * num_kb;
        it.kb1 = min(num_kb, it.kb0 + (end - pos));
        cursor += it.kb1 - it.kb0;
        return true;
    }
    const int kb_per = g.kb_per_split;
    const int num_work = num_tiles * ((num_kb + kb_per - 1) / kb_per);
    const int w = cluster_id + cursor * num_clusters;
    if (w >= num_work) return false;
    it.tile = w % num_tiles;
    it.kb0 = (w / num_tiles) * kb_per;
    it.kb1 = min(it.kb0 + kb_per, num_kb);
    ++cursor;
    return true;
}

template <int EPI, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux, const GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    using GemmSmem = GemmSmemT<BN>;
    constexpr int kStages = GemmSmem::kStages;
    constexpr uint32_t kBBytes = GemmSmem::kBBytes;
    GemmSmem &s = *reinterpret_cast<GemmSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int M = g.M, N = g.N, ldc = g.ldc;
    const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (g.K + BK - 1) / BK;
    // wgrad (few output tiles, long K) cuts K: every work item adds its partial product into C with the TMA reduce-add
    // epilogue (gemm_next_item)

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        if ((ptx::smem_u32(smem_raw) & 1023u) != 0) {
            printf("kbner gemm: dynamic shared memory is not 1024-byte aligned\n");
            __trap();
        }
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        ptx::prefetch_tensormap(&tmC);
        for (int i = 0; i < kStages; ++i) {
            ptx::mbar_init(&s.full[i], 1);
            ptx::mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&s.tmem_full[i], 1);
            ptx::mbar_init(&s.tmem_empty[i], 2 * kEpiWarps);   // epilogue warps of BOTH CTAs arrive on the leader's
        }
        if (EPI == KBNER_EPI_DGELU_BF16) {
            ptx::prefetch_tensormap(&tmAux);
            for (int i = 0; i < kEpiWarps; ++i) {
                ptx::mbar_init(&s.aux_full[i][0], 1);
                ptx::mbar_init(&s.aux_full[i][1], 1);
            }
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<kTmemCols>(&s.tmem_base);
    ptx::tc_fence_before();
    cluster_sync();            // barriers of the peer are initialised, TMEM allocated in both CTAs
    ptx::tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    pdl_wait();                // prologue done; from here on memory written by the preceding kernel is touched

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; warp-uniform, elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t a_smem0 = ptx::smem_u32(s.a[0]), b_smem0 = ptx::smem_u32(s.b[0]);
        const uint32_t full0_leader = mapa(ptx::smem_u32(&s.full[0]), 0);
        int cursor = 0;
        GemmItem wi;
        while (gemm_next_item(g, num_tiles, num_kb, cluster_id, num_clusters, cursor, wi)) {
            const int m_blk = wi.tile / num_n, n_blk = wi.tile % num_n;
            const int am0 = m_blk * BM + (int)rank * 128, bn0 = n_blk * BN + (int)rank * (BN / 2);
            for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                ptx::mbar_wait(&s.empty[stage], phase ^ 1);
                if (ptx::elect_one()) {
                    if (leader) ptx::mbar_expect_tx(&s.full[stage], 2 * (kABytes + kBBytes));
                    const uint32_t bar = full0_leader + stage * 8;
                    const uint32_t a_dst = a_smem0 + stage * kABytes, b_dst = b_smem0 + stage * kBBytes;
                    if (!g.a_mn) {
                        tma_load_2d_2sm(a_dst, &tmA, bar, kb * BK, am0);
                    } else {             // [K rows][M cols]: two 64-wide MN slabs of 64 k-rows each
                        tma_load_2d_2sm(a_dst, &tmA, bar, am0, kb * BK);
                        tma_load_2d_2sm(a_dst + 8192, &tmA, bar, am0 + 64, kb * BK);
                    }
                    if (!g.b_mn) {
                        tma_load_2d_2sm(b_dst, &tmB, bar, kb * BK, bn0);
                    } else {
                        tma_load_2d_2sm(b_dst, &tmB, bar, bn0, kb * BK);
                        if (BN == 256) tma_load_2d_2sm(b_dst + 8192, &tmB, bar, bn0 + 64, kb * BK);
                    }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA; warp-uniform, elected lane issues) =====================
        if (leader) {
            const uint32_t idesc = ptx::make_idesc_bf16(BM, BN, (uint32_t)g.a_mn, (uint32_t)g.b_mn);
            // descriptor: lo = start>>4 | LBO>>4 << 16 ; hi = SBO>>4 | version(1)<<14 | SWIZZLE_128B(2)<<29
            //   K-major : LBO unused (1), SBO = 1024 (8-row groups), k-step = +32 B
            //   MN-major: LBO = 8192 (next 64-wide MN slab), SBO = 1024 (8 k-rows), k-step = 16 rows * 128 B = +2048 B
            const uint32_t hi = 0x40004040u;
            const uint32_t a_lo0 = ((ptx::smem_u32(s.a[0]) >> 4) & 0x3FFFu) | ((g.a_mn ? 512u : 1u) << 16);
            const uint32_t b_lo0 = ((ptx::smem_u32(s.b[0]) >> 4) & 0x3FFFu) | ((g.b_mn ? 512u : 1u) << 16);
            const uint32_t a_kstep = g.a_mn ? 128u : 2u, b_kstep = g.b_mn ? 128u : 2u;
            const uint32_t empty0 = ptx::smem_u32(&s.empty[0]), tfull0 = ptx::smem_u32(&s.tmem_full[0]);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            int cursor = 0;
            GemmItem wi;
            for (; gemm_next_item(g, num_tiles, num_kb, cluster_id, num_clusters, cursor, wi); ++it) {
                const int kb0 = wi.kb0, kb1 = wi.kb1;
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(&s.tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&s.full[stage], phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint32_t a_lo = a_lo0 + stage * (kABytes >> 4), b_lo = b_lo0 + stage * (kBBytes >> 4);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            mma_f16_ss_2sm(d_tmem, pack_desc(a_lo + k * a_kstep, hi), pack_desc(b_lo + k * b_kstep, hi),
                                           idesc, (kb != kb0) || (k != 0));
                        mma_commit_mc(empty0 + stage * 8, 0b11);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (ptx::elect_one()) mma_commit_mc(tfull0 + acc * 8, 0b11);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (both CTAs; own 128 rows of the 256-row tile) =====================
        // Each warp owns 32 accumulator rows x 128 columns.  Results leave through a per-warp SWIZZLE_128B staging
        // tile (32 rows x 128 B, double-buffered) and a TMA store: the row-per-thread register layout that
        // tcgen05.ld produces would otherwise turn every 16-byte global store into 32 separate sectors (the LSU
        // was the bottleneck of the N=1024 GEMMs in profiles/r01).  ACCUM uses the TMA reduce-add (C += tile in L2).
        constexpr bool kBf16Out = (EPI == KBNER_EPI_BIAS || EPI == KBNER_EPI_BIAS_GELU || EPI == KBNER_EPI_DGELU_BF16);
        constexpr int CW = kBf16Out ? 64 : 32;          // columns per chunk = one 128-byte row segment
        constexpr int NCH = (BN / 2) / CW;
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        const uint32_t tempty_leader = mapa(ptx::smem_u32(&s.tmem_empty[0]), 0);
        const bool has_bias = g.bias != nullptr;
        uint8_t *stage_base = s.cstage[ew][0];
        const uint32_t stage_u32 = ptx::smem_u32(stage_base);
        uint32_t nstores = 0;                           // staging-buffer uses so far (lane 0 owns the bulk groups)
        uint32_t aux_phase = 0;                         // bit b: parity of the next completion of aux_full[ew][b]
        auto stage_and_store = [&](const uint4 (&q)[8], const CUtensorMap *map, int col0, int row_base, bool reduce_add) {
            const uint32_t buf = nstores & 1;
            if (nstores >= 2) {                         // the store that last read this buffer must have drained it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
            }
            uint8_t *dst = stage_base + buf * 4096 + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4 *>(dst + ((j ^ (lane & 7)) << 4)) = q[j];
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                const uint32_t src = stage_u32 + buf * 4096;
                if (reduce_add)
                    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(col0), "r"(row_base) : "memory");
                else
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(col0), "r"(row_base) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++nstores;
        };
        int it = 0;
        int cursor = 0;
        GemmItem wi;
        for (; gemm_next_item(g, num_tiles, num_kb, cluster_id, num_clusters, cursor, wi); ++it) {
            const int m_blk = wi.tile / num_n, n_blk = wi.tile % num_n;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int row_base = m_blk * BM + (int)rank * 128 + quarter * 32;
            const int row = row_base + lane;
            const bool row_ok = row < M;
            const int colbase = n_blk * BN + half * (BN / 2);
            // operand rows that do not depend on the accumulator: fetch them while the main loop still runs.
            // DGELU: the two 32-row x 64-column chunks of the saved pre-activation come in by TMA, into the very staging
            // buffers their products leave from (same box, same SWIZZLE_128B; a lane only ever touches its own 128-byte
            // row, so reading the operand and overwriting it with the result needs no cross-lane ordering).  The
            // row-per-lane ld.global this replaces was 32 sectors per request and made this epilogue (57 us at
            // 4096 x 4096 x 1024) twice as long as the main loop it is supposed to hide behind.
            const uint32_t nb = nstores;                 // chunk c of this tile uses staging buffer (nb + c) & 1
            if (EPI == KBNER_EPI_DGELU_BF16) {
                if (lane == 0) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // last tile's stores have read both buffers
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const int col0 = colbase + c * CW;
                        if (col0 < N) {
                            const uint32_t b = (nb + c) & 1;
                            ptx::mbar_expect_tx(&s.aux_full[ew][b], 4096);
                            ptx::tma_load_2d_cta(stage_base + b * 4096, &tmAux, &s.aux_full[ew][b], col0, row_base);
                        }
                    }
                }
                __syncwarp();
            }
            uint4 raux[EPI == KBNER_EPI_BIAS_RESID_F32 ? BN / 16 : 1];
            if (EPI == KBNER_EPI_BIAS_RESID_F32) {
                const uint16_t *rrow = g.aux + (size_t)(row_ok ? row : 0) * ldc + colbase;
#pragma unroll
                for (int i = 0; i < BN / 16; ++i)
                    raux[i] = (row_ok && colbase + i * 8 < N) ? ld_nc_v4(rrow + i * 8) : make_uint4(0, 0, 0, 0);
            }
            // this warp's 128 bias values: one coalesced float4 per lane, redistributed with shuffles in the chunk loop
            // (per-chunk __ldg's exposed an L2 round trip four times per tile)
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_bias && lane * 4 < BN / 2 && colbase + lane * 4 < N) bias4 = __ldg(reinterpret_cast<const float4 *>(g.bias + colbase) + lane);
            ptx::mbar_wait(&s.tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col0 = colbase + c * CW;
                float v[CW];
#pragma unroll
                for (int h2 = 0; h2 < CW / 32; ++h2) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN + half * (BN / 2) + c * CW + h2 * 32;
                    ptx::tmem_ld_32x32b_x32(taddr, r);
                    // bias of column (c*CW + h2*32 + i) sits in lane (that offset / 4), component (offset % 4)
                    float bv[32];
                    if (has_bias) {                    // warp-uniform
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int off = c * CW + h2 * 32 + i;
                            const float comp = (off & 3) == 0 ? bias4.x : (off & 3) == 1 ? bias4.y : (off & 3) == 2 ? bias4.z : bias4.w;
                            bv[i] = __shfl_sync(0xffffffffu, comp, off >> 2);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) bv[i] = 0.0f;
                    }
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[h2 * 32 + i] = __uint_as_float(r[i]) + bv[i];
                }
                if (c == NCH - 1) {                    // accumulator fully drained: let the MMA warp reuse it
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty_leader + acc * 8);
                }
                if (col0 >= N) continue;               // warp-uniform: whole chunk outside the matrix
                if (EPI == KBNER_EPI_BIAS_GELU) {
                    if (g.aux_out) {                   // training forward: keep the pre-activation for the backward pass
                        uint4 q[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            q[j].x = pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]); q[j].y = pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]);
                            q[j].z = pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]); q[j].w = pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]);
                        }
                        stage_and_store(q, &tmAux, col0, row_base, false);
                    }
#pragma unroll
                    for (int i = 0; i < CW; i += 2) gelu_erf2(v[i], v[i + 1]);
                }
                if (EPI == KBNER_EPI_DGELU_BF16) {
                    const uint32_t b = (nb + c) & 1;
                    ptx::mbar_wait(&s.aux_full[ew][b], (aux_phase >> b) & 1u);
                    aux_phase ^= 1u << b;
                }
                if (EPI == KBNER_EPI_BIAS_RESID_F32 || EPI == KBNER_EPI_DGELU_BF16) {
#pragma unroll
                    for (int i = 0; i < CW; i += 8) {
                        uint4 rv;
                        if (EPI == KBNER_EPI_DGELU_BF16)
                            rv = *reinterpret_cast<const uint4 *>(stage_base + ((nb + c) & 1) * 4096 + lane * 128 +
                                                                  (((i / 8) ^ (lane & 7)) << 4));
                        else
                            rv = raux[EPI == KBNER_EPI_BIAS_RESID_F32 ? (c * CW + i) / 8 : 0];
                        float a[8];
                        unpack_bf16x2(rv.x, a[0], a[1]); unpack_bf16x2(rv.y, a[2], a[3]);
                        unpack_bf16x2(rv.z, a[4], a[5]); unpack_bf16x2(rv.w, a[6], a[7]);
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            if (EPI == KBNER_EPI_BIAS_RESID_F32) { v[i + e] += a[e]; v[i + e + 1] += a[e + 1]; }
                            else gelu_grad_mul2(v[i + e], v[i + e + 1], a[e], a[e + 1]);
                        }
                    }
                }
                uint4 q[8];
                if (kBf16Out) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        q[j].x = pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]); q[j].y = pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]);
                        q[j].z = pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]); q[j].w = pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        q[j] = make_uint4(__float_as_uint(v[(j * 4 + 0) % CW]), __float_as_uint(v[(j * 4 + 1) % CW]),
                                          __float_as_uint(v[(j * 4 + 2) % CW]), __float_as_uint(v[(j * 4 + 3) % CW]));
                }
                stage_and_store(q, &tmC, col0, row_base, EPI == KBNER_EPI_ACCUM_F32);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the stores have READ the staging tiles (their global writes complete with the grid)
        __syncwarp();
    }
    ptx::tc_fence_before();
    cluster_sync();            // nobody leaves while the peer may still touch this CTA's smem / barriers / TMEM
    if (warp == 1) {
        ptx::tc_fence_after();
        tmem_dealloc_2sm<kTmemCols>(tmem_base);
    }
}

template <int EPI, int BN>
static int launch_gemm_t(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const CUtensorMap &tmAux,
                       const GemmArgs &g, cudaStream_t st) {
    const size_t smem = sizeof(GemmSmemT<BN>);
    static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<EPI, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return KBNER_ECUDA;
        }
        configured = true;
    }
    const int num_tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
    const int num_kb = (g.K + BK - 1) / BK;
    const int pairs = sm_budget() / 2;               // persistent: one CTA pair per two SMs of the budget
    int clusters;
    if (g.stream_units > 0) {
        clusters = (num_tiles * num_kb + g.stream_units - 1) / g.stream_units;
    } else {
        const int num_work = num_tiles * ((num_kb + g.kb_per_split - 1) / g.kb_per_split);
        clusters = num_work < pairs ? num_work : pairs;
    }
    cudaError_t le = launch_kernel(gemm_bf16_kernel<EPI, BN>, dim3(clusters * 2), dim3(kGemmThreads), smem, st, 0, true, tmA, tmB, tmC,
                                   tmAux, g);
    if (le != cudaSuccess) {
        set_error("gemm_bf16: launch failed: %s", cudaGetErrorString(le));
        return KBNER_ECUDA;
    }
    KBNER_CHECK_LAUNCH("gemm_bf16");
    return KBNER_OK;
}

template <int EPI>
static int launch_gemm(int TN, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const CUtensorMap &tmAux,
                       const GemmArgs &g, cudaStream_t st) {
    return TN == 128 ? launch_gemm_t<EPI, 128>(tmA, tmB, tmC, tmAux, g, st) : launch_gemm_t<EPI, 256>(tmA, tmB, tmC, tmAux, g, st);
}

// Tile width: 256.  KBNER_GEMM_TN=128 forces the 128-wide instantiation, kept as a measured negative result
// (profiles/r02/gemm_tile_width.json): its operand traffic from shared memory (128 rows of A + 64 of B per CTA per 64 MMA
// cycles = 128 B/clk) caps it at ~78 % of the 256-wide rate, which costs more than the rounds it saves on any shape here
// (16384 x 1024 x 4096: 119.8 vs 103.0 us; 4096 x 4096 x 1024: 32.9 vs 26.4 us).
static int pick_tile_n(int, int, int) {
    static const int forced = [] {
        const char *e = getenv("KBNER_GEMM_TN");
        return e ? atoi(e) : 0;
    }();
    return forced == 128 ? 128 : 256;
}

}  // namespace kbner

using namespace kbner;

extern "C" int kbner_gemm_bf16(const uint16_t *A, const uint16_t *B, const float *bias, const uint16_t *aux,
                               uint16_t *aux_out, void *C, int M, int N, int K, int lda, int ldb, int ldc,
                               int a_mn_major, int b_mn_major, int epilogue, void *stream) {
    KBNER_NVTX("kbner/gemm");
    KBNER_CHECK_ARG(A && B && C, "gemm: null pointer");
    KBNER_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    // TMA needs 16-byte row pitches; extents are free (ragged tile edges are zero-filled on load, clipped on store)
    KBNER_CHECK_ARG(N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0,
                    "gemm: N and the leading dimensions must be multiples of 8 (N=%d lda=%d ldb=%d ldc=%d)", N, lda, ldb, ldc);
    KBNER_CHECK_ARG((epilogue != KBNER_EPI_BIAS_RESID_F32 && epilogue != KBNER_EPI_DGELU_BF16) || aux,
                    "gemm: epilogue %d needs the aux operand", epilogue);
    KBNER_CHECK_ARG(((uintptr_t)C & 15u) == 0 && (!bias || ((uintptr_t)bias & 15u) == 0) &&
                        (!aux || ((uintptr_t)aux & 15u) == 0) && (!aux_out || ((uintptr_t)aux_out & 15u) == 0),
                    "gemm: C / bias / aux must be 16-byte aligned");
    const int TN = pick_tile_n(M, N, epilogue);
    CUtensorMap tmA, tmB;
    int rc;
    // K-major operand: tensor [rows = M|N][cols = K], box 128 (A) or TN/2 (B) rows x 64 k.  MN-major: [rows = K][cols = M|N], box 64 k x 64 mn.
    rc = a_mn_major ? make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, 64)
                    : make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, BK);
    if (rc) return rc;
    rc = b_mn_major ? make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, 64)
                    : make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)(TN / 2), BK);
    if (rc) return rc;
    const bool bf16_out = (epilogue == KBNER_EPI_BIAS || epilogue == KBNER_EPI_BIAS_GELU || epilogue == KBNER_EPI_DGELU_BF16);
    CUtensorMap tmC, tmAux;
    rc = make_tmap_2d(&tmC, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, bf16_out ? 64 : 32, bf16_out ? 2 : 4);
    if (rc) return rc;
    tmAux = tmC;
    if (aux_out || epilogue == KBNER_EPI_DGELU_BF16) {
        rc = make_tmap_2d(&tmAux, aux_out ? aux_out : aux, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, 64, 2);
        if (rc) return rc;
    }
    const int num_kb_h = (K + BK - 1) / BK;
    int kb_per = num_kb_h, stream_units = 0;
    if (epilogue == KBNER_EPI_ACCUM_F32) {
        // wgrad: few output tiles (64 at 1024 x 4096), long K (the 4096 tokens).  Stream-K: the tiles' k-blocks form one
        // line of tiles * num_kb units, cut into equal shares for the CTA pairs; a share that crosses a tile boundary is two
        // work items, and every item leaves through the reduce-add epilogue.  Shares are even, so no item is shorter than 2
        // k-blocks when num_kb is even.  Round 1's split-K (kb_per_split, KBNER_GEMM_STREAMK=0) gave every pair whole
        // splits: 64 tiles x 2 splits = 1.73 rounds of 32 k-blocks on 74 pairs = 64 k-blocks of time for 55.4 of work.
        static const bool streamk = [] {
            const char *e = getenv("KBNER_GEMM_STREAMK");
            return !(e && e[0] == '0');
        }();
        const int tiles = ((M + BM - 1) / BM) * ((N + TN - 1) / TN);
        const int pairs = sm_budget() / 2;
        if (streamk) {
            const int total = tiles * num_kb_h;
            int U = (total + pairs - 1) / pairs;
            U += U & 1;
            if (U < 4) U = 4;
            stream_units = U;
        } else {
            // fill ~2 work items per cluster, but keep >= 4 k-blocks per item so the pipeline prologue stays amortised
            int splits = (2 * pairs) / tiles;
            if (splits > num_kb_h / 4) splits = num_kb_h / 4;
            if (splits < 1) splits = 1;
            kb_per = (num_kb_h + splits - 1) / splits;
        }
    }
    GemmArgs g{bias, aux, aux_out, C, M, N, K, ldc, a_mn_major ? 1 : 0, b_mn_major ? 1 : 0, kb_per, stream_units};
    cudaStream_t st = (cudaStream_t)stream;
    switch (epilogue) {
        case KBNER_EPI_BIAS: return launch_gemm<KBNER_EPI_BIAS>(TN, tmA, tmB, tmC, tmAux, g, st);
        case KBNER_EPI_BIAS_GELU: return launch_gemm<KBNER_EPI_BIAS_GELU>(TN, tmA, tmB, tmC, tmAux, g, st);
        case KBNER_EPI_BIAS_RESID_F32: return launch_gemm<KBNER_EPI_BIAS_RESID_F32>(TN, tmA, tmB, tmC, tmAux, g, st);
        case KBNER_EPI_NONE_F32: return launch_gemm<KBNER_EPI_NONE_F32>(TN, tmA, tmB, tmC, tmAux, g, st);
        case KBNER_EPI_DGELU_BF16: return launch_gemm<KBNER_EPI_DGELU_BF16>(TN, tmA, tmB, tmC, tmAux, g, st);
        case KBNER_EPI_ACCUM_F32: return launch_gemm<KBNER_EPI_ACCUM_F32>(TN, tmA, tmB, tmC, tmAux, g, st);
        default: set_error("gemm: unknown epilogue %d", epilogue); return KBNER_EINVAL;
    }
}

// The forward "TN" call of round 1 (both operands K-major, torch.nn.Linear layout).
extern "C" int kbner_gemm_bf16_tn(const uint16_t *A, const uint16_t *B, const float *bias,
                                  const uint16_t *residual, void *C, int M, int N, int K, int lda, int ldb,
                                  int ldc, int epilogue, void *stream) {
    KBNER_NVTX("kbner/gemm");
    KBNER_CHECK_ARG(epilogue == KBNER_EPI_NONE_F32 || bias, "gemm: epilogue %d needs a bias", epilogue);
    return kbner_gemm_bf16(A, B, bias, residual, nullptr, C, M, N, K, lda, ldb, ldc, 0, 0, epilogue, stream);
}
