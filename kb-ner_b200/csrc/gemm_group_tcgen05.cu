// GROUPED weight-gradient GEMM of one encoder layer (fine-tuning backward):
//     dW_p[out_p, in_p] += dY_p^T . X_p        p = 0 .. count-1 (<= 4),  all over the same `tokens` rows
// i.e. the four torch.nn.Linear weight gradients autograd produces for a BertLayer (FFN-down, FFN-up, attention-output, fused
// Q|K|V) under /root/reference/flair/trainers/finetune_trainer.py:956-957 `loss.backward()`; SURVEY.md E2/E4/E5/E6 backward.
//
// Why one launch: as four launches of gemm_tcgen05.cu (stream-K, reduce-add epilogue) the layer's weight gradients took
// 94 us at 4096 tokens for 103 GFLOP (profiles/r02/wgrad_streamk.json) -- 1100 TF/s: every launch pays its ramp (TMEM
// allocation, barrier set-up, the first operand loads, ~3 us), the exposed epilogue of its last item and the imbalance of
// its own tile count over 74 CTA pairs (the attention-output gradient, 16 tiles, ran at 560 TF/s).  Here the k-blocks of ALL
// tiles of ALL problems form one line of units, cut into equal shares for the CTA pairs (stream-K across problems); a share
// that crosses a tile (or problem) boundary is two work items, every item leaves through the TMA reduce-add epilogue.
//
// Same machinery as gemm_tcgen05.cu: cluster (2,1,1) = CTA pair, tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16), both
// operands read MN-major in place (dY [tokens][out], X [tokens][in]: no transposes), 5-stage TMA ring, TMEM accumulators
// double-buffered, warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue.
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "cluster_ptx.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace kbner {

constexpr int GBM = 256, GBN = 256, GBK = 64, kGStages = 5;
constexpr int kGEpiWarps = 8;
constexpr int kGThreads = 64 + kGEpiWarps * 32;
constexpr uint32_t kGABytes = 128 * GBK * 2, kGBBytes = 128 * GBK * 2;     // per CTA per stage
constexpr uint32_t kGTmemCols = 512;
constexpr int kMaxGroup = 4;

struct GroupSmem {
    uint8_t a[kGStages][kGABytes];
    uint8_t b[kGStages][kGBBytes];
    uint8_t cstage[kGEpiWarps][2][4096];    // per epilogue warp: 2 x (32 rows x 128 B) SWIZZLE_128B reduce-add staging
    uint64_t full[kGStages];
    uint64_t empty[kGStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

struct WgradGroup {
    CUtensorMap tmA[kMaxGroup];             // dY_p  [tokens][out_p] bf16, box 64 tokens x 64 columns
    CUtensorMap tmB[kMaxGroup];             // X_p   [tokens][in_p]  bf16, box 64 x 64
    CUtensorMap tmC[kMaxGroup];             // dW_p  [out_p][in_p]   fp32, box 32 rows x 32 columns
    int unit_begin[kMaxGroup + 1];          // first k-block unit of problem p on the common line (tiles_p * num_kb each)
    int num_n[kMaxGroup];                   // 256-wide tiles along in_p
    int count, num_kb;
    int stream_units;                       // > 0: stream-K shares of this many k-block units per cluster
    int lanes;                              // > 0: whole tiles, cluster c takes tiles c, c + lanes, c + 2 lanes, ... of the common line
};

struct GroupItem {
    int p, m_blk, n_blk, kb0, kb1;
};
// The work items of one cluster in the order all three warp roles walk them: its share [c * U, (c + 1) * U) of the unit line,
// cut at tile boundaries.  `cursor` = units done.
__device__ __forceinline__ bool group_next_item(const WgradGroup &G, int cluster_id, int &cursor, GroupItem &it) {
    const int total = G.unit_begin[G.count];
    if (G.lanes > 0) {
        const int w = cluster_id + cursor * G.lanes;              // tile index on the common line
        const int pos = w * G.num_kb;
        if (pos >= total) return false;
        int p = 0;
        while (p + 1 < G.count && pos >= G.unit_begin[p + 1]) ++p;
        const int tile = (pos - G.unit_begin[p]) / G.num_kb;
        it.p = p;
        it.m_blk = tile / G.num_n[p];
        it.n_blk = tile - it.m_blk * G.num_n[p];
        it.kb0 = 0;
        it.kb1 = G.num_kb;
        ++cursor;
        return true;
    }
    const int begin = cluster_id * G.stream_units;
    const int end = min(begin + G.stream_units, total);
    const int pos = begin + cursor;
    if (pos >= end) return false;
    int p = 0;
    while (p + 1 < G.count && pos >= G.unit_begin[p + 1]) ++p;
    const int local = pos - G.unit_begin[p];
    const int tile = local / G.num_kb;
    it.p = p;
    it.m_blk = tile / G.num_n[p];
    it.n_blk = tile - it.m_blk * G.num_n[p];
    it.kb0 = local - tile * G.num_kb;
    it.kb1 = min(G.num_kb, it.kb0 + (end - pos));
    cursor += it.kb1 - it.kb0;
    return true;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGThreads, 1)
gemm_wgrad_group_kernel(const __grid_constant__ WgradGroup G) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    GroupSmem &s = *reinterpret_cast<GroupSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        if ((ptx::smem_u32(smem_raw) & 1023u) != 0) {
            printf("kbner gemm_group: dynamic shared memory is not 1024-byte aligned\n");
            __trap();
        }
        for (int p = 0; p < G.count; ++p) {
            ptx::prefetch_tensormap(&G.tmA[p]);
            ptx::prefetch_tensormap(&G.tmB[p]);
            ptx::prefetch_tensormap(&G.tmC[p]);
        }
        for (int i = 0; i < kGStages; ++i) {
            ptx::mbar_init(&s.full[i], 1);
            ptx::mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&s.tmem_full[i], 1);
            ptx::mbar_init(&s.tmem_empty[i], 2 * kGEpiWarps);   // epilogue warps of BOTH CTAs arrive on the leader's
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<kGTmemCols>(&s.tmem_base);
    ptx::tc_fence_before();
    cluster_sync();            // barriers of the peer are initialised, TMEM allocated in both CTAs
    ptx::tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    pdl_wait();                // prologue done; from here on memory written by the preceding kernels is touched

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; warp-uniform, elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t a_smem0 = ptx::smem_u32(s.a[0]), b_smem0 = ptx::smem_u32(s.b[0]);
        const uint32_t full0_leader = mapa(ptx::smem_u32(&s.full[0]), 0);
        int cursor = 0;
        GroupItem wi;
        while (group_next_item(G, cluster_id, cursor, wi)) {
            const CUtensorMap *tmA = &G.tmA[wi.p], *tmB = &G.tmB[wi.p];
            const int am0 = wi.m_blk * GBM + (int)rank * 128, bn0 = wi.n_blk * GBN + (int)rank * 128;
            for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                ptx::mbar_wait(&s.empty[stage], phase ^ 1);
                if (ptx::elect_one()) {
                    if (leader) ptx::mbar_expect_tx(&s.full[stage], 2 * (kGABytes + kGBBytes));
                    const uint32_t bar = full0_leader + stage * 8;
                    const uint32_t a_dst = a_smem0 + stage * kGABytes, b_dst = b_smem0 + stage * kGBBytes;
                    // [K rows = tokens][MN columns]: two 64-wide MN slabs of 64 token rows each, per operand
                    tma_load_2d_2sm(a_dst, tmA, bar, am0, kb * GBK);
                    tma_load_2d_2sm(a_dst + 8192, tmA, bar, am0 + 64, kb * GBK);
                    tma_load_2d_2sm(b_dst, tmB, bar, bn0, kb * GBK);
                    tma_load_2d_2sm(b_dst + 8192, tmB, bar, bn0 + 64, kb * GBK);
                }
                __syncwarp();
                if (++stage == kGStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA; warp-uniform, elected lane issues) =====================
        if (leader) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(GBM, GBN, 1, 1);          // both operands MN-major
            // descriptor: MN-major, LBO = 8192 (next 64-wide MN slab), SBO = 1024 (8 k-rows), k-step = 16 rows * 128 B = +2048 B
            const uint32_t hi = 0x40004040u;
            const uint32_t a_lo0 = ((ptx::smem_u32(s.a[0]) >> 4) & 0x3FFFu) | (512u << 16);
            const uint32_t b_lo0 = ((ptx::smem_u32(s.b[0]) >> 4) & 0x3FFFu) | (512u << 16);
            const uint32_t empty0 = ptx::smem_u32(&s.empty[0]), tfull0 = ptx::smem_u32(&s.tmem_full[0]);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            int cursor = 0;
            GroupItem wi;
            for (; group_next_item(G, cluster_id, cursor, wi); ++it) {
                const int acc = it & 1;
                ptx::mbar_wait(&s.tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * GBN;
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    ptx::mbar_wait(&s.full[stage], phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint32_t a_lo = a_lo0 + stage * (kGABytes >> 4), b_lo = b_lo0 + stage * (kGBBytes >> 4);
#pragma unroll
                        for (int k = 0; k < GBK / 16; ++k)
                            mma_f16_ss_2sm(d_tmem, pack_desc(a_lo + k * 128u, hi), pack_desc(b_lo + k * 128u, hi), idesc,
                                           (kb != wi.kb0) || (k != 0));
                        mma_commit_mc(empty0 + stage * 8, 0b11);
                    }
                    __syncwarp();
                    if (++stage == kGStages) { stage = 0; phase ^= 1; }
                }
                if (ptx::elect_one()) mma_commit_mc(tfull0 + acc * 8, 0b11);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (both CTAs; own 128 rows of the 256-row tile): dW += tile =====================
        // Each warp owns 32 accumulator rows x 128 columns, four chunks of 32 fp32 columns: registers -> the warp's
        // SWIZZLE_128B staging tile (double-buffered) -> TMA reduce-add into dW in L2.
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        const uint32_t tempty_leader = mapa(ptx::smem_u32(&s.tmem_empty[0]), 0);
        uint8_t *stage_base = s.cstage[ew][0];
        const uint32_t stage_u32 = ptx::smem_u32(stage_base);
        uint32_t nstores = 0;                           // staging-buffer uses so far (lane 0 owns the bulk groups)
        int it = 0;
        int cursor = 0;
        GroupItem wi;
        for (; group_next_item(G, cluster_id, cursor, wi); ++it) {
            const CUtensorMap *tmC = &G.tmC[wi.p];
            const int acc = it & 1;
            const int row_base = wi.m_blk * GBM + (int)rank * 128 + quarter * 32;
            const int colbase = wi.n_blk * GBN + half * 128;
            ptx::mbar_wait(&s.tmem_full[acc], (it >> 1) & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(tmem_base + (uint32_t(quarter * 32) << 16) + acc * GBN + half * 128 + c * 32, r);
                ptx::tmem_ld_wait();
                if (c == 3) {                          // accumulator fully drained: let the MMA warp reuse it
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty_leader + acc * 8);
                }
                const uint32_t buf = nstores & 1;
                if (nstores >= 2) {                    // the reduce that last read this buffer must have drained it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
                uint8_t *dst = stage_base + buf * 4096 + lane * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4 *>(dst + ((j ^ (lane & 7)) << 4)) = make_uint4(r[j * 4], r[j * 4 + 1], r[j * 4 + 2], r[j * 4 + 3]);
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(tmC)), "r"(stage_u32 + buf * 4096), "r"(colbase + c * 32), "r"(row_base)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                ++nstores;
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the reduces have read the staging tiles
        __syncwarp();
    }
    ptx::tc_fence_before();
    cluster_sync();            // nobody leaves while the peer may still touch this CTA's smem / barriers / TMEM
    if (warp == 1) {
        ptx::tc_fence_after();
        tmem_dealloc_2sm<kGTmemCols>(tmem_base);
    }
}

}  // namespace kbner

using namespace kbner;

extern "C" int kbner_gemm_wgrad_group(int count, const uint16_t *const *dY, const uint16_t *const *X, float *const *dW,
                                      const int *n_out, const int *n_in, const int *ld_dy, const int *ld_x, int tokens,
                                      void *stream) {
    KBNER_NVTX("kbner/gemm");
    KBNER_CHECK_ARG(count >= 1 && count <= kMaxGroup, "gemm_wgrad_group: 1..%d problems (count=%d)", kMaxGroup, count);
    KBNER_CHECK_ARG(dY && X && dW && n_out && n_in && ld_dy && ld_x && tokens > 0, "gemm_wgrad_group: null pointer / no tokens");
    WgradGroup G;
    memset(&G, 0, sizeof(G));
    G.count = count;
    G.num_kb = (tokens + GBK - 1) / GBK;
    int units = 0;
    for (int p = 0; p < count; ++p) {
        KBNER_CHECK_ARG(dY[p] && X[p] && dW[p] && n_out[p] > 0 && n_in[p] > 0, "gemm_wgrad_group: problem %d is empty", p);
        KBNER_CHECK_ARG(n_out[p] % 8 == 0 && n_in[p] % 8 == 0 && ld_dy[p] % 8 == 0 && ld_x[p] % 8 == 0 && ld_dy[p] >= n_out[p] &&
                            ld_x[p] >= n_in[p],
                        "gemm_wgrad_group: problem %d: extents and leading dimensions must be multiples of 8 (out=%d in=%d ld=%d/%d)", p,
                        n_out[p], n_in[p], ld_dy[p], ld_x[p]);
        KBNER_CHECK_ARG(((uintptr_t)dW[p] & 15u) == 0, "gemm_wgrad_group: dW[%d] must be 16-byte aligned", p);
        int rc = make_tmap_bf16_2d(&G.tmA[p], dY[p], (uint64_t)tokens, (uint64_t)n_out[p], (uint64_t)ld_dy[p], 64, 64);
        if (rc) return rc;
        rc = make_tmap_bf16_2d(&G.tmB[p], X[p], (uint64_t)tokens, (uint64_t)n_in[p], (uint64_t)ld_x[p], 64, 64);
        if (rc) return rc;
        rc = make_tmap_2d(&G.tmC[p], dW[p], (uint64_t)n_out[p], (uint64_t)n_in[p], (uint64_t)n_in[p], 32, 32, 4);
        if (rc) return rc;
        G.num_n[p] = (n_in[p] + GBN - 1) / GBN;
        G.unit_begin[p] = units;
        units += ((n_out[p] + GBM - 1) / GBM) * G.num_n[p] * G.num_kb;
    }
    for (int p = count; p <= kMaxGroup; ++p) G.unit_begin[p] = units;
    const int pairs = sm_budget() / 2;
    // Work assignment.  "tiles" (default): whole tiles, all K, round-robin over `lanes` clusters with lanes = tiles / rounds --
    // the clusters of a round walk the token axis in lockstep, so the operand panels the tiles of a row / column share are
    // read from DRAM once and served from L2.  "stream" (KBNER_WGRAD_GROUP_MODE=stream): equal shares of the unit line; every
    // pair is busy to the end, but shares start at arbitrary k-blocks, pairs that share a panel read different parts of it
    // at any moment, and the launch re-read its operands 2.5 times from DRAM (429 MB for 134 MB, 65 % of the HBM bandwidth:
    // profiles/r02/wgrad_group_ncu_stream.txt) -- 96 us for the four gradients of a layer, no better than four launches.
    static const bool stream_mode = [] {
        const char *e = getenv("KBNER_WGRAD_GROUP_MODE");
        return e && e[0] == 's';
    }();
    int clusters;
    if (stream_mode) {
        int U = (units + pairs - 1) / pairs;
        U += U & 1;                                   // even shares: no item shorter than 2 k-blocks when num_kb is even
        if (U < 4) U = 4;
        G.stream_units = U;
        clusters = (units + U - 1) / U;
    } else {
        const int tiles = units / G.num_kb;
        const int rounds = (tiles + pairs - 1) / pairs;
        G.lanes = (tiles + rounds - 1) / rounds;      // 192 tiles on 74 pairs: 3 rounds of 64
        clusters = G.lanes;
    }
    const size_t smem = sizeof(GroupSmem);
    static std::atomic<bool> configured{false};       // idempotent set-up: a race only repeats it
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("gemm_wgrad_group: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return KBNER_ECUDA;
        }
        configured = true;
    }
    cudaError_t le = launch_kernel(gemm_wgrad_group_kernel, dim3(clusters * 2), dim3(kGThreads), smem, (cudaStream_t)stream, 0, true, G);
    if (le != cudaSuccess) {
        set_error("gemm_wgrad_group: launch failed: %s", cudaGetErrorString(le));
        return KBNER_ECUDA;
    }
    KBNER_CHECK_LAUNCH("gemm_wgrad_group");
    return KBNER_OK;
}
