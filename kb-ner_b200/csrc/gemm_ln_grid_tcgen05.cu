// Fused   Y = LayerNorm(A . W^T + bias + resid) * gamma + beta      (bf16 in, fp32 accumulate / statistics, bf16 out)
// -- second generation of csrc/gemm_ln_tcgen05.cu: the same fusion (attention-output dense and FFN-down dense of an encoder
// layer with the BertSelfOutput / BertOutput tail, called under /root/reference/flair/embeddings.py:3269),
// but the per-row LayerNorm statistics are exchanged through GLOBAL memory between independent CTA pairs instead of
// through distributed shared memory inside one thread-block cluster.
//
// Why (profiles/r02/layer_full_prof1.txt): the cluster version needs all N = 1024 columns of a 256-row panel in the TMEM of
// ONE cluster of 4 CTAs (2 pairs x 2 tiles = all 512 columns of every CTA), so the tensor cores idle from the last MMA of a
// panel until pass 1 of the second tile, the DSMEM exchange and pass 2 of the first tile are over (tensor pipe 46 % at
// K = 1024, 67 % at K = 4096), and clusters of 4 fit on 132 of the 148 SMs only.  Here a work item is ONE 256 x 256 tile
// (panel, column tile), every CTA pair (cta_group::2) walks items exactly like gemm_tcgen05.cu -- accumulators double-buffered
// in TMEM, so the main loop of item i+1 runs under the whole epilogue of item i -- and the N/256 pairs that hold the
// column tiles of one panel meet through one 16-byte statistics slot per (tile, column half, row) in L2:
//   pass 1   z = acc + bias + resid back to TMEM (tcgen05.st), running (mean, M2) of the thread's 128 columns (packed
//            fp32-pair arithmetic); the NEXT item's residual tile is requested as soon as this one has been read
//   publish  ONE st.volatile.v4 {mean, tag, M2, tag} per thread (NCCL-LL idiom: each 8-byte half carries the launch tag, a
//            reader that sees both tags sees both values) -- no fence, no atomic, no per-warp serialisation
//   collect  every thread polls the 2 * N/256 slots of ITS row until all carry this launch's tag, then merges them in FIXED
//            order (Chan et al.) -> same bits in every consumer
//   pass 2   normalise z from TMEM, bf16, SWIZZLE_128B staging, TMA store; the accumulator buffer goes back to the MMA warp
// (first version, profiles/r02/gemm_ln_grid_timeline_v1.json: counter + __threadfence + atomicAdd + lane-0 poll cost 4300
// cycles per item, the residual fetch 2500, and the epilogue (14 400 cycles) was longer than the main loop at K = 1024.)
// The tag is epoch + 1; the last CTA to leave the kernel bumps the epoch in the workspace, so a captured CUDA graph can
// replay the launch on the same workspace and stale slots never match.  Deadlock freedom: the grid is persistent with at
// most one CTA per SM (all pairs co-resident), every pair walks its items in increasing order, and an item waits only on
// items of the same panel, which its peers reach in the same round (the pair count is a multiple of N/256).
//
// Warp roles per CTA: warp 0 TMA producer, warp 1 MMA issuer (leader CTA of the pair), warps 2..9 epilogue.
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

#include "cluster_ptx.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace kbner {

constexpr int QBM = 256, QBN = 256, QBK = 64, kQStages = 4;
constexpr int kQEpiWarps = 8;
constexpr int kQThreads = 64 + kQEpiWarps * 32;
constexpr uint32_t kQABytes = 128 * QBK * 2, kQBBytes = 128 * QBK * 2;
constexpr uint32_t kQTmemCols = 512;
constexpr int kQMaxTiles = 4;               // N <= 1024

struct GemmLnGridSmem {
    uint8_t a[kQStages][kQABytes];
    uint8_t b[kQStages][kQBBytes];
    uint8_t rstage[kQEpiWarps][2][4096];    // per warp, per 64-column half: residual tile of the current / next item (32 rows x 128 B,
                                            // 16-byte chunks XOR-swizzled); refilled right after pass 1, under the exchange + pass 2
    uint8_t ostage[kQEpiWarps][4096];       // per warp: bf16 output tile (32 rows x 64 columns, SWIZZLE_128B) of the TMA store
    uint64_t full[kQStages];
    uint64_t empty[kQStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

struct GemmLnGridArgs {
    const float *bias;        // [N] or NULL
    const uint16_t *resid;    // [M,N] bf16 or NULL
    const float *gamma, *beta;
    uint32_t *control;        // [0] epoch (launches completed on this workspace), [1] CTAs of the running launch that have left
    uint4 *stats;             // [panels][N/256][2][256]  {mean, tag, M2, tag} of 128 columns
    int M, N, K;
    float eps;
    unsigned long long *timeline;   // debug (kbner_debug_gemm_ln_timeline): clock64 stamps [block][warp 0..9][item 0..7][8]; NULL = off
};

#define GLN_STAMP(item, ev)                                                                                         \
    do {                                                                                                            \
        if (g.timeline && lane == 0 && (item) < 8)                                                                  \
            g.timeline[(((size_t)blockIdx.x * 10 + warp) * 8 + (item)) * 8 + (ev)] = (unsigned long long)clock64(); \
    } while (0)

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_v4(uint4 *p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kQThreads, 1)
gemm_ln_grid_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmY, const GemmLnGridArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    GemmLnGridSmem &s = *reinterpret_cast<GemmLnGridSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sub = cluster_ctarank();            // which 128 rows of the panel
    const bool leader = sub == 0;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int M = g.M, N = g.N;
    const int ntn = N / QBN;                           // column tiles per panel; num_pairs % ntn == 0 (host)
    const int num_panels = (M + QBM - 1) / QBM;
    const int num_work = num_panels * ntn;
    const int num_kb = (g.K + QBK - 1) / QBK;
    const int ct = pair_id % ntn;                      // this pair's column tile: the same for every item it walks
    const int col0 = ct * QBN;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        if ((ptx::smem_u32(smem_raw) & 1023u) != 0) {
            printf("kbner gemm_ln_grid: dynamic shared memory is not 1024-byte aligned\n");
            __trap();
        }
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        ptx::prefetch_tensormap(&tmY);
        for (int i = 0; i < kQStages; ++i) {
            ptx::mbar_init(&s.full[i], 1);
            ptx::mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&s.tmem_full[i], 1);
            ptx::mbar_init(&s.tmem_empty[i], 2 * kQEpiWarps);   // epilogue warps of BOTH CTAs arrive on the leader's
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<kQTmemCols>(&s.tmem_base);
    ptx::tc_fence_before();
    cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    pdl_wait();                // (the prologue only read parameters: bias / gamma / beta are not written inside a step)

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t a_smem0 = ptx::smem_u32(s.a[0]), b_smem0 = ptx::smem_u32(s.b[0]);
        const uint32_t full0_leader = mapa(ptx::smem_u32(&s.full[0]), 0);
        const int bn0 = col0 + (int)sub * 128;
        for (int w = pair_id; w < num_work; w += num_pairs) {
            const int am0 = (w / ntn) * QBM + (int)sub * 128;
            for (int kb = 0; kb < num_kb; ++kb) {
                ptx::mbar_wait(&s.empty[stage], phase ^ 1);
                if (ptx::elect_one()) {
                    if (leader) ptx::mbar_expect_tx(&s.full[stage], 2 * (kQABytes + kQBBytes));
                    const uint32_t bar = full0_leader + stage * 8;
                    tma_load_2d_2sm(a_smem0 + stage * kQABytes, &tmA, bar, kb * QBK, am0);
                    tma_load_2d_2sm(b_smem0 + stage * kQBBytes, &tmB, bar, kb * QBK, bn0);
                }
                __syncwarp();
                if (++stage == kQStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA of the pair) =====================
        if (leader) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(QBM, QBN, 0, 0);
            const uint32_t hi = 0x40004040u;          // SBO = 1024, version 1, SWIZZLE_128B (see gemm_tcgen05.cu)
            const uint32_t a_lo0 = ((ptx::smem_u32(s.a[0]) >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t b_lo0 = ((ptx::smem_u32(s.b[0]) >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t empty0 = ptx::smem_u32(&s.empty[0]), tfull0 = ptx::smem_u32(&s.tmem_full[0]);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int w = pair_id; w < num_work; w += num_pairs, ++it) {
                const int acc = it & 1;
                GLN_STAMP(it, 0);
                ptx::mbar_wait(&s.tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                GLN_STAMP(it, 1);
                const uint32_t d_tmem = tmem_base + acc * QBN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&s.full[stage], phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint32_t a_lo = a_lo0 + stage * (kQABytes >> 4), b_lo = b_lo0 + stage * (kQBBytes >> 4);
#pragma unroll
                        for (int k = 0; k < QBK / 16; ++k)
                            mma_f16_ss_2sm(d_tmem, pack_desc(a_lo + k * 2u, hi), pack_desc(b_lo + k * 2u, hi), idesc,
                                           (kb != 0) || (k != 0));
                        mma_commit_mc(empty0 + stage * 8, 0b11);
                    }
                    __syncwarp();
                    if (++stage == kQStages) { stage = 0; phase ^= 1; }
                }
                if (ptx::elect_one()) mma_commit_mc(tfull0 + acc * 8, 0b11);
                __syncwarp();
                GLN_STAMP(it, 2);
            }
        }
    } else {
        // ===================== epilogue: bias + residual + LayerNorm across the pairs of a panel =====================
        const int ew = warp - 2;
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
        const int half = ew >> 2;                     // which 128 of the tile's 256 columns
        const int row_l = (int)sub * 128 + quarter * 32 + lane;      // row inside the 256-row panel
        const uint32_t tempty_leader = mapa(ptx::smem_u32(&s.tmem_empty[0]), 0);
        uint8_t *rbuf = s.rstage[ew][0];
        uint8_t *obuf = s.ostage[ew];
        const uint32_t rbuf_u32 = ptx::smem_u32(rbuf), obuf_u32 = ptx::smem_u32(obuf);
        const bool has_resid = g.resid != nullptr, has_bias = g.bias != nullptr;
        const float inv_n = 1.0f / (float)N;
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const int colw = col0 + half * 128;           // first global column of this warp
        const float4 *bias4 = reinterpret_cast<const float4 *>(g.bias + (has_bias ? colw : 0));
        const float4 *gamma4 = reinterpret_cast<const float4 *>(g.gamma + colw), *beta4 = reinterpret_cast<const float4 *>(g.beta + colw);
        const int nsl = 2 * ntn;                      // partials behind one row: 2 column halves of every tile
        // launch tag of the statistics slots: epoch + 1 (the last CTA to leave bumps the epoch, so the slots of the previous
        // launch never match); read after griddepcontrol.wait, i.e. after the previous launch has completely finished
        const uint32_t tag = ld_volatile_u32(g.control) + 1u;
        // residual tile [32 rows x 128 bf16] -> shared memory with cp.async: the global side is coalesced (4 rows x 128 B
        // per request), the shared side lands row-per-lane ready (row r at r * 128 B, 16-byte chunks XOR-swizzled)
        auto issue_resid = [&](int w_item) {
            const int row0 = (w_item / ntn) * QBM + (int)sub * 128 + quarter * 32;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = 4 * i + (lane >> 3);
                    const int grow = row0 + r;
                    const uint32_t dst = rbuf_u32 + hh * 4096 + r * 128 + (((lane & 7) ^ (r & 7)) << 4);
                    if (grow < M) {
                        const uint16_t *src = g.resid + (size_t)grow * N + colw + hh * 64 + (lane & 7) * 8;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    } else {
                        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (has_resid && pair_id < num_work) issue_resid(pair_id);
        if (g.timeline && lane == 0) {                 // common time base of the CTAs: (globaltimer, clock64) pair in item slot 7
            g.timeline[(((size_t)blockIdx.x * 10 + warp) * 8 + 7) * 8 + 0] = ptx::global_timer_ns();
            g.timeline[(((size_t)blockIdx.x * 10 + warp) * 8 + 7) * 8 + 1] = (unsigned long long)clock64();
        }
        int it = 0;
        for (int w = pair_id; w < num_work; w += num_pairs, ++it) {
            const int panel = w / ntn;
            const int acc = it & 1;
            const int row_w0 = panel * QBM + (int)sub * 128 + quarter * 32;     // first global row of this warp
            GLN_STAMP(it, 0);
            ptx::mbar_wait(&s.tmem_full[acc], (it >> 1) & 1);
            ptx::tc_fence_after();
            GLN_STAMP(it, 1);
            if (has_resid) asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();                              // every lane's part of the residual tile is in shared memory
            GLN_STAMP(it, 2);
            const uint32_t taddr0 = tmem_base + lane_addr + acc * QBN + half * 128;
            // ---- pass 1: z = acc + bias + resid back to TMEM, (mean, M2) of this thread's 128 columns (packed fp32 pairs:
            //      FADD2 / FFMA2 halve the issue slots; two accumulator pairs per sum keep four dependency chains in flight)
            float n_a = 0.0f, mean_a = 0.0f, m2_a = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(taddr0 + c * 32, r);
                float4 bv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) bv[q] = has_bias ? __ldg(bias4 + c * 8 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                uint4 rv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = (c & 1) * 4 + q;                    // 16-byte chunk inside the 128-byte row of half c >> 1
                    rv[q] = has_resid ? *reinterpret_cast<const uint4 *>(rbuf + (c >> 1) * 4096 + lane * 128 + ((chunk ^ (lane & 7)) << 4))
                                      : make_uint4(0, 0, 0, 0);
                }
                ptx::tmem_ld_wait();
                uint64_t z2[16];
                uint64_t s0 = 0ull, s1 = 0ull;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t rw[4] = {rv[q].x, rv[q].y, rv[q].z, rv[q].w};
                    const float4 b0 = bv[2 * q], b1 = bv[2 * q + 1];
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = q * 8 + e * 2;
                        float a0, a1;
                        unpack_bf16x2(rw[e], a0, a1);
                        // (acc + bias) + resid: the order of the unfused path
                        const uint64_t z = f2_add(f2_add(f2_pack(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), f2_pack(bb[e * 2], bb[e * 2 + 1])),
                                                  f2_pack(a0, a1));
                        z2[i >> 1] = z;
                        if (e & 1) s1 = f2_add(s1, z); else s0 = f2_add(s0, z);
                        float zx, zy;
                        f2_unpack(z, zx, zy);
                        r[i] = __float_as_uint(zx);
                        r[i + 1] = __float_as_uint(zy);
                    }
                }
                ptx::tmem_st_32x32b_x32(taddr0 + c * 32, r);
                float sx, sy;
                f2_unpack(f2_add(s0, s1), sx, sy);
                const float cm = (sx + sy) * (1.0f / 32.0f);
                const uint64_t ncm = f2_pack(-cm, -cm);
                uint64_t q0 = 0ull, q1 = 0ull;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const uint64_t d0 = f2_add(z2[i], ncm), d1 = f2_add(z2[i + 1], ncm);
                    q0 = f2_fma(d0, d0, q0);
                    q1 = f2_fma(d1, d1, q1);
                }
                float qx, qy;
                f2_unpack(f2_add(q0, q1), qx, qy);
                const float sq = qx + qy;
                if (c == 0) { n_a = 32.0f; mean_a = cm; m2_a = sq; }
                else chan_merge(n_a, mean_a, m2_a, 32.0f, cm, sq);
            }
            __syncwarp();                              // every lane has read its residual rows: the buffer may be refilled
            if (has_resid && w + num_pairs < num_work) issue_resid(w + num_pairs);      // lands under the exchange + pass 2
            GLN_STAMP(it, 3);
            // ---- publish the partial: ONE 16-byte store {mean, tag, M2, tag} (each 8-byte half carries the tag, the NCCL-LL
            //      idiom: a reader that sees both tags sees both values) -- no fence, no atomic, no per-warp serialisation
            uint4 *slab = g.stats + (size_t)panel * nsl * QBM;                         // [tile][half][row]
            st_volatile_v4(&slab[(size_t)(ct * 2 + half) * QBM + row_l],
                           make_uint4(__float_as_uint(mean_a), tag, __float_as_uint(m2_a), tag));
            ptx::tmem_st_wait();                       // (z must be in TMEM before pass 2 reads it back; overlaps the exchange)
            GLN_STAMP(it, 4);
            // ---- collect the 2 * ntn partials of this row: poll the slots themselves
            float2 part[2 * kQMaxTiles];
            {
                uint32_t pending = (1u << nsl) - 1u;
                uint64_t t0 = 0;
                uint32_t spins = 0;
                while (pending) {
                    uint4 v[2 * kQMaxTiles];            // all loads of a round in flight together (a load per check serialised
#pragma unroll                                          // 8 L2 round trips: 5000-7000 cycles per item in the first timeline)
                    for (int sl = 0; sl < 2 * kQMaxTiles; ++sl)
                        if (sl < nsl) v[sl] = ld_volatile_v4(&slab[(size_t)sl * QBM + row_l]);
#pragma unroll
                    for (int sl = 0; sl < 2 * kQMaxTiles; ++sl) {
                        if (sl < nsl && v[sl].y == tag && v[sl].w == tag) {
                            part[sl] = make_float2(__uint_as_float(v[sl].x), __uint_as_float(v[sl].z));
                            pending &= ~(1u << sl);
                        }
                    }
                    if (pending && (++spins & 0x3Fu) == 0) {
                        const uint64_t now = (uint64_t)clock64();
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > 8000000000ull) {
                            printf("kbner gemm_ln_grid: statistics of panel %d never completed (block %d warp %d lane %d: missing 0x%x)\n",
                                   panel, blockIdx.x, warp, lane, pending);
                            __trap();
                        }
                    }
                }
            }
            __syncwarp();
            GLN_STAMP(it, 5);
            // Chan et al. with equal counts (128 columns per partial): after sl partials n_a = 128 sl, so n_b / n = 1 / (sl + 1)
            // and n_a n_b / n = 128 sl / (sl + 1) are compile-time constants.  Fixed order: same bits in every consumer.
            float mean = part[0].x, m2 = part[0].y;
#pragma unroll
            for (int sl = 1; sl < 2 * kQMaxTiles; ++sl) {
                if (sl < nsl) {
                    const float delta = part[sl].x - mean;
                    mean = fmaf(delta, 1.0f / (float)(sl + 1), mean);
                    m2 = (m2 + part[sl].y) + delta * delta * (128.0f * (float)sl / (float)(sl + 1));
                }
            }
            const float rstd = rsqrtf(m2 * inv_n + g.eps);
            const uint64_t rstd2 = f2_pack(rstd, rstd), nmr2 = f2_pack(-mean * rstd, -mean * rstd);
            GLN_STAMP(it, 6);
            // ---- pass 2: normalise z from TMEM, bf16, SWIZZLE_128B staging, TMA store (chunks of 64 columns)
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                uint4 ov[8];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32b_x32(taddr0 + (c2 * 2 + h2) * 32, r);
                    float4 gv[8], bt[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        gv[q] = __ldg(gamma4 + (c2 * 2 + h2) * 8 + q);
                        bt[q] = __ldg(beta4 + (c2 * 2 + h2) * 8 + q);
                    }
                    ptx::tmem_ld_wait();
                    uint32_t o[16];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        // ((z * rstd) - mean * rstd) * gamma + beta on pairs
                        const uint64_t t0 = f2_fma(f2_pack(__uint_as_float(r[q * 4 + 0]), __uint_as_float(r[q * 4 + 1])), rstd2, nmr2);
                        const uint64_t t1 = f2_fma(f2_pack(__uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3])), rstd2, nmr2);
                        float y0, y1, y2, y3;
                        f2_unpack(f2_fma(t0, f2_pack(gv[q].x, gv[q].y), f2_pack(bt[q].x, bt[q].y)), y0, y1);
                        f2_unpack(f2_fma(t1, f2_pack(gv[q].z, gv[q].w), f2_pack(bt[q].z, bt[q].w)), y2, y3);
                        o[q * 2] = pack_bf16x2(y0, y1);
                        o[q * 2 + 1] = pack_bf16x2(y2, y3);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) ov[h2 * 4 + q] = make_uint4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
                }
                if (c2 == 1) {                         // accumulator drained for good: the MMA warp may reuse it
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty_leader + acc * 8);
                }
                // one staging tile per warp: the store that last read it (64 columns ago) must be done with it -- it was
                // issued before the 64 columns above were computed
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                uint8_t *dst = obuf + lane * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4 *>(dst + ((q ^ (lane & 7)) << 4)) = ov[q];
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&tmY)), "r"(obuf_u32), "r"(colw + c2 * 64), "r"(row_w0)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            GLN_STAMP(it, 7);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
    ptx::tc_fence_before();
    cluster_sync();            // nobody leaves while the peer may still touch this CTA's smem / barriers / TMEM
    if (warp == 1) {
        ptx::tc_fence_after();
        tmem_dealloc_2sm<kQTmemCols>(tmem_base);
    }
    if (threadIdx.x == 0) {
        // every warp of this CTA has finished polling.  The last CTA of the grid to get here bumps the epoch: the next launch
        // (which reads it after its griddepcontrol.wait, i.e. after this grid has completed) tags its slots differently.
        __threadfence();
        const uint32_t fin = atomicAdd(g.control + 1, 1u);
        if (fin == gridDim.x - 1) {
            g.control[1] = 0u;
            g.control[0] = g.control[0] + 1u;
        }
    }
}

}  // namespace kbner

using namespace kbner;

// Workspace of the statistics exchange for an M x N problem: 16 control bytes (epoch, CTAs that have left), then one
// 16-byte slot per (panel, tile, column half, row).  The caller allocates it ZEROED once.
extern "C" size_t kbner_gemm_ln_workspace_bytes(int M, int N) {
    if (M <= 0 || N <= 0 || N % QBN != 0) return 0;
    const size_t panels = (size_t)(M + QBM - 1) / QBM;
    return 16 + panels * (size_t)(N / QBN) * 2 * QBM * sizeof(uint4);
}

// Debug hook (scripts/gemm_ln_timeline.py): `buf` = 148 * 10 * 8 * 8 uint64 of device memory receives clock64 stamps of the
// next launches (NULL turns it off again).  Not part of the product path.
static unsigned long long *g_gln_timeline = nullptr;
extern "C" int kbner_debug_gemm_ln_timeline(void *buf) {
    g_gln_timeline = reinterpret_cast<unsigned long long *>(buf);
    return KBNER_OK;
}
extern "C" int kbner_gemm_ln_grid(const uint16_t *A, const uint16_t *W, const float *bias, const uint16_t *resid,
                                  const float *gamma, const float *beta, float eps, uint16_t *Y, int M, int N, int K, int lda,
                                  int ldw, void *workspace, size_t workspace_bytes, void *stream) {
    KBNER_NVTX("kbner/gemm_ln");
    KBNER_CHECK_ARG(A && W && gamma && beta && Y && workspace, "gemm_ln_grid: null pointer");
    KBNER_CHECK_ARG(M > 0 && K > 0, "gemm_ln_grid: empty problem M=%d K=%d", M, K);
    KBNER_CHECK_ARG(N % QBN == 0 && N >= QBN && N <= QBN * kQMaxTiles,
                    "gemm_ln_grid: the fused LayerNorm epilogue needs N in {256, 512, 768, 1024} (N=%d)", N);
    KBNER_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0, "gemm_ln_grid: leading dimensions must be multiples of 8");
    KBNER_CHECK_ARG((((uintptr_t)Y | (uintptr_t)resid | (uintptr_t)bias | (uintptr_t)gamma | (uintptr_t)beta |
                      (uintptr_t)workspace) & 15u) == 0,
                    "gemm_ln_grid: operands must be 16-byte aligned");
    KBNER_CHECK_ARG(workspace_bytes >= kbner_gemm_ln_workspace_bytes(M, N),
                    "gemm_ln_grid: workspace of %zu bytes, kbner_gemm_ln_workspace_bytes(%d, %d) = %zu", workspace_bytes, M, N,
                    kbner_gemm_ln_workspace_bytes(M, N));
    const int ntn = N / QBN;
    const int pairs = (sm_budget() / 2) / ntn * ntn;   // the pairs of one panel must run in the same round
    KBNER_CHECK_ARG(pairs >= ntn, "gemm_ln_grid: an SM budget of %d cannot hold the %d CTA pairs of one row panel", sm_budget(), ntn);
    CUtensorMap tmA, tmB, tmY;
    int rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, QBK);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 128, QBK);
    if (rc) return rc;
    rc = make_tmap_2d(&tmY, Y, (uint64_t)M, (uint64_t)N, (uint64_t)N, 32, 64, 2);
    if (rc) return rc;
    const size_t smem = sizeof(GemmLnGridSmem);
    static std::atomic<bool> configured{false};       // idempotent set-up: a race only repeats it
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_ln_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("gemm_ln_grid: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return KBNER_ECUDA;
        }
        configured = true;
    }
    const int panels = (M + QBM - 1) / QBM;
    const int num_work = panels * ntn;
    const int clusters = num_work < pairs ? num_work : pairs;
    uint32_t *control = reinterpret_cast<uint32_t *>(workspace);
    uint4 *stats = reinterpret_cast<uint4 *>(control + 4);
    GemmLnGridArgs g{bias, resid, gamma, beta, control, stats, M, N, K, eps, g_gln_timeline};
    cudaError_t le = launch_kernel(gemm_ln_grid_kernel, dim3(clusters * 2), dim3(kQThreads), smem, (cudaStream_t)stream, 0, true,
                                   tmA, tmB, tmY, g);
    if (le != cudaSuccess) {
        set_error("gemm_ln_grid: launch failed: %s", cudaGetErrorString(le));
        return KBNER_ECUDA;
    }
    KBNER_CHECK_LAUNCH("gemm_ln_grid");
    return KBNER_OK;
}

extern "C" int kbner_gemm_bias_resid_layernorm_ws(const uint16_t *A, const uint16_t *W, const float *bias,
                                                  const uint16_t *resid, const float *gamma, const float *beta, float eps,
                                                  uint16_t *Y, int M, int N, int K, int lda, int ldw, void *workspace,
                                                  size_t workspace_bytes, void *stream) {
    KBNER_CHECK_ARG(M > 0 && N % QBN == 0 && N >= QBN && N <= QBN * kQMaxTiles,
                    "gemm_ln: the fused LayerNorm epilogue needs M > 0 and N in {256, 512, 768, 1024} (M=%d N=%d)", M, N);
    const int ntn = N / QBN;
    const int pairs = (sm_budget() / 2) / ntn * ntn;
    // Which kernel (profiles/r02/gemm_ln_grid_vs_cluster.json): with ONE round of items (M <= 4608 at N = 1024) nothing waits
    // for a TMEM buffer and every SM works -- the grid version wins (16.1 vs 18.7 us at 4096 x 1024 x 1024).  With several
    // rounds both versions serialise pass 2 of an item with the main loop of the next-but-one (TMEM holds two accumulators),
    // and the L2 round trips + pair skew of the global exchange (~7000 cycles per item, gemm_ln_grid_timeline_v3.json) cost
    // more than the cluster's DSMEM exchange saves by idling 16 SMs (44.4 vs 39.6 us at 16384 x 1024 x 1024; 112 vs 113 us at
    // K = 4096).  KBNER_GEMM_LN=grid|cluster forces one.
    static const int forced = [] {
        const char *e = getenv("KBNER_GEMM_LN");
        return !e ? 0 : (e[0] == 'g' ? 1 : (e[0] == 'c' ? 2 : 0));
    }();
    const int panels_h = (M + QBM - 1) / QBM;
    if (forced == 2 || (forced == 0 && panels_h * ntn > pairs))
        return kbner_gemm_bias_resid_layernorm(A, W, bias, resid, gamma, beta, eps, Y, M, N, K, lda, ldw, stream);
    return kbner_gemm_ln_grid(A, W, bias, resid, gamma, beta, eps, Y, M, N, K, lda, ldw, workspace, workspace_bytes, stream);
}
