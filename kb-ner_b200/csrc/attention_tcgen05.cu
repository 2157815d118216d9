// Fused multi-head self-attention forward for one encoder layer (head dim 64):
//   O = softmax(Q.K^T / 8 + key-padding mask) . V       per (window, head),
// flash-style: the [S,S] score matrix the reference's transformers-3.0.0 path materialises in
// fp32 (call site /root/reference/flair/embeddings.py:3269) never leaves the SM.
//
// PERSISTENT CTAs, two resident per SM (97 KB of shared memory, 256 TMEM columns, 96 registers each), each walking the work
// items (window r, head h, block of 128 query rows) blockIdx.x, blockIdx.x + gridDim.x, ...; every barrier, ring slot and
// TMEM buffer is indexed by g, the running count of 64-key blocks the CTA has gone through, so the pipeline never drains
// between items.  (One CTA per item spent 41 % of its life outside the key-block loop -- profiles/README.md, steps 17-19.)
//   warp 9        TMA producer: K and V in separate 3-slot rings, the Q tile of the next item; per-item address arithmetic
//   warp 8        MMA issuer (warp-uniform, elect.sync around the issue): S_g = Q.K_g^T -> TMEM (2 buffers x 64 cols) two
//                 blocks ahead of the softmax, O += P_g.V_g -> TMEM (2 buffers x 64 cols, one per item in flight);
//                 tcgen05.mma kind::f16, V consumed MN-major straight from its row-major [key][d] tile
//   warps 0..7    softmax, two threads per query row (= TMEM lane; warps w and w+4 share a lane quarter and split the
//                 64 key columns): tcgen05.ld the scores, online max / ex2 / sum in fp32 registers, P_g -> bf16 ->
//                 shared memory in the SWIZZLE_128B K-major layout the MMA reads; O stays in TMEM and is rescaled lazily
//                 (only when a row maximum grows by more than 2^8); an item's read-out is deferred past the next item's
//                 first key block.
// Training (DROP): attention-probability dropout (transformers BertSelfAttention.dropout, p = 0.1) is applied to P_g
// after the row sum has been taken, with the stateless counter-hash mask of common.cuh indexed by
// (window, head, query, key); the 1/(1-p) scale is folded into the final normalisation.
// Bound: MUFU (one ex2 per score: 16/clk/SM => 512 clk per 128x64 block) + issue + the per-block latency chain; the
// tensor pipe is ~21 % busy.
#include <math_constants.h>

#include <atomic>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace kbner {

constexpr int kAttnD = 64;
constexpr int kBQ = 128, kBKV = 64, kKVStages = 3, kMaxS = 512;
constexpr int kAttnThreads = 320;      // 8 softmax warps (2 threads per query row) + MMA warp + TMA producer warp
constexpr uint32_t kQBytes = 128 * 64 * 2;     // [128 rows][64 bf16], SWIZZLE_128B
constexpr uint32_t kKVBytes = 64 * 64 * 2;     // [64 keys][64 bf16]
constexpr uint32_t kAttnTmemCols = 256;

struct AttnSmem {
    uint8_t q[kQBytes];
    uint8_t k[kKVStages][kKVBytes];
    uint8_t v[kKVStages][kKVBytes];
    uint8_t p[2][kQBytes];             // [buffer][128 rows x 64 keys]
    uint64_t bar_q;                    // Q tile of the next work item has landed
    uint64_t q_free;                   // the last S = Q.K^T of an item has retired: the Q tile may be replaced
    uint32_t blk_flags[8];             // per key block g (slot g & 7), written by the producer: bit 0 = first of its item, bit 1 = last
    uint64_t k_full[kKVStages], v_full[kKVStages];
    uint64_t k_free[kKVStages];        // S_j = Q.K_j^T retired: the K slot may be refilled (two blocks before its V slot)
    uint64_t v_free[kKVStages];        // P_j.V_j retired: the V slot may be refilled
    uint64_t bar_s[2];                 // S_j ready in TMEM buffer j&1
    uint64_t bar_p[2];                 // P_j written to smem buffer j&1 (128 arrivals)
    uint64_t bar_o[2];                 // O_j ready in TMEM buffer j&1
    float xchg[2][2][128];             // [parity][column half][row]: block row-max exchange between the two halves
    float xsum[2][2][128];             // [item parity][column half][row]: final row-sum exchange
    uint32_t tmem_base;
};

// registers -> TMEM: this warp's 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#ifdef KBNER_ATTN_DEBUG
// timeline probe (debug builds only, scripts/attn_timeline.py): clock64 stamps of CTA 0 / CTA 150,
// [cta][role 0 = MMA warp, 1 = softmax warp 0, 2 = softmax warp 7][key block g < 24][stamp]
__device__ unsigned long long g_attn_dbg[2 * 3 * 24 * 4];
#define ATTN_STAMP(role, g, k)                                                                                      \
    do {                                                                                                            \
        if (dbg_cta >= 0 && (role) >= 0 && lane == 0 && (g) < 24)                                                   \
            g_attn_dbg[((dbg_cta * 3 + (role)) * 24 + (g)) * 4 + (k)] = clock64();                                   \
    } while (0)
#else
#define ATTN_STAMP(role, g, k) do { } while (0)
#endif

// Position in a CTA's stream of key blocks: work item w = ((window * heads + head) * nqb + query block), key block j.
struct BlockCursor {
    int w, j, nkb, row0, h, qb;
};

template <bool DROP>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     const int32_t *__restrict__ key_len, int R, int S, int H, int heads, uint16_t *__restrict__ out,
                     int ldo, uint16_t *__restrict__ out_lo, uint16_t *__restrict__ out_hi2, float *__restrict__ lse_out,
                     const Dropout drop) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    AttnSmem &s = *reinterpret_cast<AttnSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqb = (S + kBQ - 1) / kBQ;
    const int total = R * heads * nqb;               // work items; this CTA takes blockIdx.x, blockIdx.x + gridDim.x, ...

#ifdef KBNER_ATTN_DEBUG
    const int dbg_cta = blockIdx.x == 0 ? 0 : (blockIdx.x == 150 ? 1 : -1);
#endif
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        if ((ptx::smem_u32(smem_raw) & 1023u) != 0) {
            printf("kbner attention: dynamic shared memory is not 1024-byte aligned\n");
            __trap();
        }
        ptx::prefetch_tensormap(&tmQ);
        ptx::prefetch_tensormap(&tmKV);
        ptx::mbar_init(&s.bar_q, 1);
        ptx::mbar_init(&s.q_free, 1);
        for (int i = 0; i < kKVStages; ++i) {
            ptx::mbar_init(&s.k_full[i], 1);
            ptx::mbar_init(&s.v_full[i], 1);
            ptx::mbar_init(&s.k_free[i], 1);
            ptx::mbar_init(&s.v_free[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&s.bar_s[i], 1);
            ptx::mbar_init(&s.bar_p[i], 256);
            ptx::mbar_init(&s.bar_o[i], 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 8) ptx::tmem_alloc<kAttnTmemCols>(&s.tmem_base);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    const uint32_t tmem_s = tmem_base;            // 2 x 64 columns: S of key block g in buffer g & 1
    const uint32_t tmem_o = tmem_base + 128;      // 2 x 64 columns: O of the CTA's n-th work item in buffer n & 1
    pdl_wait();                // Q / K / V come from the preceding GEMM (key_len is a step input, not a kernel output)

    // every barrier and ring slot is indexed by g, the running count of key blocks this CTA has gone through (all items)
    auto open_item = [&](BlockCursor &c) {           // decode c.w, skipping windows without a valid key; c.nkb == 0 at the end
        c.nkb = 0;
        while (c.w < total) {
            const int rh = c.w / nqb;
            const int r = rh / heads;
            const int nkb = (min(__ldg(key_len + r), S) + kBKV - 1) / kBKV;
            if (nkb > 0) {
                c.qb = c.w - rh * nqb;
                c.h = rh - r * heads;
                c.row0 = r * S;
                c.nkb = nkb;
                return;
            }
            c.w += gridDim.x;
        }
    };
    auto advance = [&](BlockCursor &c) {
        if (++c.j == c.nkb) {
            c.w += gridDim.x;
            c.j = 0;
            open_item(c);
        }
    };

    if (warp == 9) {
        // ===================== TMA producer (warp-uniform; the copy itself under elect.sync) =====================
        // Walks this CTA's key-block stream once for K (and the Q tile at each item's first block) and, two blocks behind,
        // once for V.  K_t goes into slot t % 3 as soon as S_{t-3} has retired, V_t after P.V of block t-3; the Q tile of the
        // next item replaces the current one when the item's last S has retired.  All the per-item address arithmetic
        // (integer divisions, the key_len load) lives here, off the MMA warp's serial instruction stream.
        BlockCursor kc{(int)blockIdx.x, 0, 0, 0, 0, 0};
        open_item(kc);
        BlockCursor vc = kc;
        uint32_t t = 0, kslot = 0, kuse = 0, vslot = 0, vuse = 0, nq = 0;     // use = how often the slot has been filled
        while (kc.nkb != 0 || vc.nkb != 0) {
            if (kc.nkb != 0) {
                if (kuse > 0) ptx::mbar_wait(&s.k_free[kslot], (kuse - 1) & 1);
                const bool first = kc.j == 0;
                if (ptx::elect_one()) {
                    s.blk_flags[t & 7] = (first ? 1u : 0u) | (kc.j == kc.nkb - 1 ? 2u : 0u);
                    ptx::mbar_expect_tx(&s.k_full[kslot], kKVBytes);
                    ptx::tma_load_2d(s.k[kslot], &tmKV, &s.k_full[kslot], H + kc.h * kAttnD, kc.row0 + kc.j * kBKV);
                }
                __syncwarp();
                if (++kslot == kKVStages) { kslot = 0; ++kuse; }
            }
            if (t >= 2 && vc.nkb != 0) {
                if (vuse > 0) ptx::mbar_wait(&s.v_free[vslot], (vuse - 1) & 1);
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(&s.v_full[vslot], kKVBytes);
                    ptx::tma_load_2d(s.v[vslot], &tmKV, &s.v_full[vslot], 2 * H + vc.h * kAttnD, vc.row0 + vc.j * kBKV);
                }
                __syncwarp();
                if (++vslot == kKVStages) { vslot = 0; ++vuse; }
                advance(vc);
            }
            if (kc.nkb != 0) {
                if (kc.j == 0) {                     // Q of this item (after the previous item's last S)
                    if (nq > 0) ptx::mbar_wait(&s.q_free, (nq - 1) & 1);
                    if (ptx::elect_one()) {
                        ptx::mbar_expect_tx(&s.bar_q, kQBytes);
                        ptx::tma_load_2d(s.q, &tmQ, &s.bar_q, kc.h * kAttnD, kc.row0 + kc.qb * kBQ);
                    }
                    __syncwarp();
                    ++nq;
                }
                advance(kc);
            }
            ++t;
        }
        // Drain: every tcgen05.commit of the stream's tail arrives on a k_free / v_free / q_free barrier that no refill waits
        // for any more.  Wait for them here, so that no asynchronous arrive is still in flight towards this CTA's shared
        // memory when it exits (compute-sanitizer synccheck: "Missing wait", profiles/r02/sanitizer_san1.txt).
        for (uint32_t sl = 0; sl < (uint32_t)kKVStages; ++sl) {
            const uint32_t kf = kuse + (sl < kslot ? 1u : 0u), vf = vuse + (sl < vslot ? 1u : 0u);   // fills of slot sl
            if (kf > 0) ptx::mbar_wait(&s.k_free[sl], (kf - 1) & 1);
            if (vf > 0) ptx::mbar_wait(&s.v_free[sl], (vf - 1) & 1);
        }
        if (nq > 0) ptx::mbar_wait(&s.q_free, (nq - 1) & 1);
    } else if (warp == 8) {
        // ===================== MMA issuer: WARP-UNIFORM, only the issuing instructions sit under elect.sync ===============
        // (Under `if (lane == 0)` ptxas wraps every UTCHMMA / UTCBAR in an ELECT / PLOP3 / BRA.U.ANY loop.)  This warp's
        // instruction stream is serial and shares a scheduler with four softmax warps -- measured ~10 cycles per
        // instruction -- so it carries nothing but waits, 8 MMAs and their commits per key block: S = Q.K^T runs two blocks
        // ahead of the softmax, P.V follows it; item boundaries arrive as flags from the producer.
        int G = 0;                                   // key blocks in this CTA's stream
        for (int w = blockIdx.x + lane * gridDim.x; w < total; w += 32 * gridDim.x)
            G += (min(__ldg(key_len + w / (nqb * heads)), S) + kBKV - 1) / kBKV;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) G += __shfl_xor_sync(0xffffffffu, G, o);
        constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, 64, 0, 0);   // S = Q.K^T : both K-major
        constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, 64, 0, 1);   // O = P.V   : V is MN-major
        const uint32_t q_addr = ptx::smem_u32(s.q);
        const uint32_t k_addr0 = ptx::smem_u32(s.k[0]), v_addr0 = ptx::smem_u32(s.v[0]), p_addr0 = ptx::smem_u32(s.p[0]);
        uint32_t kslot = 0, kuse = 0, vslot = 0, vuse = 0, nq = 0, obuf = 0;
        auto issue_s = [&](int m) {                  // S of block m into TMEM buffer m & 1
            ptx::mbar_wait(&s.k_full[kslot], kuse & 1);
            const uint32_t f = s.blk_flags[m & 7];
            if (f & 1u) {
                ptx::mbar_wait(&s.bar_q, nq & 1);
                ++nq;
            }
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t k_addr = k_addr0 + kslot * kKVBytes;
#pragma unroll
                for (int kk = 0; kk < kAttnD / 16; ++kk) {
                    const uint64_t da = ptx::make_sw128_desc(q_addr + kk * 32, 16, 1024);
                    const uint64_t db = ptx::make_sw128_desc(k_addr + kk * 32, 16, 1024);
                    ptx::mma_f16_ss(tmem_s + (m & 1) * 64, da, db, idesc_s, kk != 0);
                }
                ptx::mma_commit(&s.bar_s[m & 1]);
                ptx::mma_commit(&s.k_free[kslot]);
                if (f & 2u) ptx::mma_commit(&s.q_free);
            }
            __syncwarp();
            if (++kslot == kKVStages) { kslot = 0; ++kuse; }
        };
        if (G > 0) issue_s(0);
        if (G > 1) issue_s(1);
        for (int g = 0; g < G; ++g) {
            // P_g in smem (the softmax warps executed fence.proxy.async before arriving)
            ptx::mbar_wait(&s.bar_p[g & 1], (g >> 1) & 1);
            ATTN_STAMP(0, g, 0);
            ptx::mbar_wait(&s.v_full[vslot], vuse & 1);
            ATTN_STAMP(0, g, 1);
            const uint32_t f = s.blk_flags[g & 7];
            if ((f & 1u) && g > 0) ++obuf;           // a new item accumulates into the other O buffer (obuf = item index)
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t v_addr = v_addr0 + vslot * kKVBytes;
                const uint32_t p_addr = p_addr0 + (g & 1) * kQBytes;
#pragma unroll
                for (int kk = 0; kk < kBKV / 16; ++kk) {
                    const uint64_t da = ptx::make_sw128_desc(p_addr + kk * 32, 16, 1024);
                    // V tile [key][d]: 16 keys per MMA = two 8-row swizzle atoms, 1024 B apart (SBO)
                    const uint64_t db = ptx::make_sw128_desc(v_addr + kk * 16 * 128, 16, 1024);
                    ptx::mma_f16_ss(tmem_o + (obuf & 1u) * 64, da, db, idesc_o, !(f & 1u) || (kk != 0));   // O accumulates in TMEM
                }
                // bar_o[item & 1] completes ONCE per item, with its last P.V: every phase of it is waited for by the
                // read-out (a per-block commit left phases nobody observed -- compute-sanitizer synccheck "Missing wait");
                // the lazy rescale of the running O synchronises on v_free instead, which covers the same MMAs
                if (f & 2u) ptx::mma_commit(&s.bar_o[obuf & 1u]);
                ptx::mma_commit(&s.v_free[vslot]);
            }
            __syncwarp();
            if (++vslot == kKVStages) { vslot = 0; ++vuse; }
            ATTN_STAMP(0, g, 2);
            if (g + 2 < G) issue_s(g + 2);           // its S buffer was drained before P_g was published
            ATTN_STAMP(0, g, 3);
        }
    } else {
        // ===================== softmax warps: two threads per query row =====================
        // warps w and w+4 share TMEM lane quarter w (rows 32w..32w+31); `half` selects which 32 of the block's 64 key
        // columns -- and which 32 of the 64 output columns -- the thread owns.  Four warps per scheduler (with the two
        // resident CTAs) instead of two: the round-1 kernel was latency-bound with MUFU and issue both at ~40 %.
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;              // TMEM lane == row in the query block
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const float scale_log2 = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
        const uint32_t dkey = DROP ? drop_key(drop) : 0u;
        uint32_t g = 0, n_item = 0, n_epi = 0;
        // The read-out of an item (wait for its last P.V, O from TMEM, normalise, store) is DEFERRED until this warp has
        // finished the first key block of the NEXT item: O is double-buffered in TMEM and the next S is already waiting, so
        // the all-warps-arrive -> last P.V -> read-out chain (3.9-4.5 k cycles per item in the clock64 timeline) leaves the
        // critical path.  `pend_*` is the item whose read-out is outstanding.
        bool pend = false;
        float pend_m = 0.0f, pend_l = 0.0f;
        int pend_qrow = 0, pend_row0 = 0, pend_h = 0, pend_r = 0;
        uint32_t pend_o = 0, pend_g = 0;
        // The output tile leaves through shared memory: a thread owns one query ROW (TMEM lane), so storing its 64 bytes
        // directly made every 16-byte store instruction touch 32 rows = 32 half-written sectors per request (round-1 ncu:
        // 2 097 152 sectors / 65 536 requests).  Instead the two warps of a quarter write their halves of the 32 rows into
        // the P buffer of the item's LAST key block -- free once bar_o has confirmed that P.V retired, in the same
        // SWIZZLE_128B pattern P uses (conflict-free) -- and read it back transposed: 8 lanes per row, 4 rows per store
        // instruction = 16 fully written sectors per request.  Only this quarter's rows of the buffer are touched and
        // the next P write to it sits behind the row-max pair barrier of the next key block.
        auto stage_store = [&](uint8_t *stage, const uint32_t (&pk)[16], uint16_t *gbase, int qrow0, uint16_t *gbase2) {
            uint8_t *srow = stage + row * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4 *>(srow + (((half * 4 + c) ^ (row & 7)) << 4)) =
                    make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rq = quarter * 32 + half * 16 + i * 4 + (lane >> 3);      // row inside the 128-query block
                const int chunk = lane & 7;
                const uint4 v = *reinterpret_cast<const uint4 *>(stage + rq * 128 + ((chunk ^ (rq & 7)) << 4));
                if (qrow0 + rq < S) {
                    const size_t go = (size_t)rq * ldo + chunk * 8;
                    *reinterpret_cast<uint4 *>(gbase + go) = v;
                    if (gbase2) *reinterpret_cast<uint4 *>(gbase2 + go) = v;
                }
            }
        };
        auto read_out = [&](bool from_tmem, uint32_t o_addr, uint32_t item, uint8_t *stage, float m_run, float l_run, int qrow,
                            int row0, int h, int r) {
            const int qrow0 = qrow - row;                       // first query row of the item's 128-row block
            float o_acc[32];
            if (from_tmem) {
                ptx::mbar_wait(&s.bar_o[item & 1], (item >> 1) & 1);       // the item's last P.V has retired
                ptx::tc_fence_after();
                uint32_t ro[32];
                ptx::tmem_ld_32x32b_x32(o_addr, ro);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[i] = __uint_as_float(ro[i]);
                ptx::tc_fence_before();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[i] = 0.0f;
            }
            // total row sum = both halves.  The exchange slot alternates per read-out: the partner warp reads slot p after
            // this pair barrier and cannot see the slot rewritten before it has passed the NEXT pair barrier.
            s.xsum[n_epi & 1][half][row] = l_run;
            asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
            const float l_tot = l_run + s.xsum[n_epi & 1][half ^ 1][row];
            ++n_epi;
            const float inv = (l_tot > 0.0f) ? (DROP ? drop.scale : 1.0f) / l_tot : 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) o_acc[i] *= inv;
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(o_acc[2 * i], o_acc[2 * i + 1]);
            const size_t gofs = (size_t)(row0 + qrow0) * ldo + h * kAttnD;
            // out_lo: the rounding residual lo = bf16(o - hi), hi + lo = the fp32 value to 2^-17 (the "bf16x3" operand pair of
            // the attention-output GEMM, and what the backward's D = rowsum(dO * O) is taken from); out_hi2: a second copy of
            // hi (the K-concatenated operand row is [ hi | lo | hi ])
            stage_store(stage, pk, out + gofs, qrow0, out_hi2 ? out_hi2 + gofs : nullptr);
            if (out_lo) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float h0, h1;
                    unpack_bf16x2(pk[i], h0, h1);
                    pk[i] = pack_bf16x2(o_acc[2 * i] - h0, o_acc[2 * i + 1] - h1);
                }
                asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");     // the partner has read the hi tile back
                stage_store(stage, pk, out_lo + gofs, qrow0, nullptr);
            }
            if (qrow < S && lse_out && half == 0)   // natural-log LSE of the scaled scores (for the backward pass)
                lse_out[((size_t)r * heads + h) * S + qrow] =
                    (l_tot > 0.0f) ? (m_run + log2f(l_tot)) * 0.6931471805599453f : -CUDART_INF_F;
        };
        int klen_next = (int)blockIdx.x < total ? __ldg(key_len + blockIdx.x / (nqb * heads)) : 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            const int rh = w / nqb;
            const int qb = w - rh * nqb, r = rh / heads;
            const int h = rh - r * heads;
            const int klen = min(klen_next, S);
            if (w + (int)gridDim.x < total) klen_next = __ldg(key_len + (w + gridDim.x) / (nqb * heads));   // lands during this item
            const int nkb = (klen + kBKV - 1) / kBKV;        // key blocks that hold at least one valid key
            const int row0 = r * S;                          // first row of this window in the [R*S, 3H] matrix
            const int qrow = qb * kBQ + row;                  // sub-token index inside the window
            float m_run = -CUDART_INF_F, l_run = 0.0f;
            // dropout counter of (window, head, query): 256 key PAIRS per row (kMaxS / 2), the same in the backward kernel
            const uint32_t drow = (((uint32_t)(r * heads + h) * (uint32_t)kMaxS) + (uint32_t)qrow) * (uint32_t)(kMaxS / 2);
            const uint32_t o_addr = tmem_o + (n_item & 1) * 64 + lane_addr + half * 32;     // this thread's 32 of the 64 output columns

            for (int j = 0; j < nkb; ++j, ++g) {
                ATTN_STAMP((warp == 0 ? 1 : (warp == 7 ? 2 : -1)), g, 0);
                ptx::mbar_wait(&s.bar_s[g & 1], (g >> 1) & 1);
                ATTN_STAMP((warp == 0 ? 1 : (warp == 7 ? 2 : -1)), g, 1);
                ptx::tc_fence_after();
                float sc[32];
                {
                    uint32_t rs[32];
                    ptx::tmem_ld_32x32b_x32(tmem_s + lane_addr + (g & 1) * 64 + half * 32, rs);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) sc[i] = __uint_as_float(rs[i]);
                }
                ptx::tc_fence_before();
                const int kbase = j * kBKV + half * 32;
                if (kbase + 32 > klen) {         // only the last block of the window is ragged
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (kbase + i >= klen) sc[i] = -CUDART_INF_F;
                }
                float m_half = sc[0];
#pragma unroll
                for (int i = 1; i < 31; i += 2) m_half = fmax3(m_half, sc[i], sc[i + 1]);      // FMNMX3: 16 instead of 31 issue slots
                m_half = fmaxf(m_half, sc[31]);
                // row max over both halves: exchange through shared memory, pair barrier = the two warps of this quarter
                s.xchg[g & 1][half][row] = m_half;
                asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
                const float m_blk = fmaxf(m_half, s.xchg[g & 1][half ^ 1][row]) * scale_log2;   // finite: the block has >= 1 valid key
                // LAZY rescaling: the running O lives in TMEM and is only touched when the row maximum grows by more than 2^8;
                // otherwise the stale maximum stays the reference (probabilities up to 256 are exact enough in bf16 / fp32 and the
                // final division by the row sum cancels the common factor).  Both threads of a row take the same decision (same
                // m_blk, same m_run).  tcgen05.ld / .st are warp-collective: the decision is taken per WARP (any row of the warp
                // over the threshold -> all 32 rows do the exact online-softmax update); warps w and w+4 see identical per-row
                // values, so both halves of a row agree.
                if (j == 0) {
                    m_run = m_blk;
                } else if (__any_sync(0xffffffffu, m_blk > m_run + 8.0f)) {
                    const float m_new = fmaxf(m_run, m_blk);
                    const float alpha = ex2_approx(m_run - m_new);
                    m_run = m_new;
                    l_run *= alpha;
                    ptx::mbar_wait(&s.v_free[(g - 1) % kKVStages], ((g - 1) / kKVStages) & 1);    // P.V of the previous block has retired
                    ptx::tc_fence_after();
                    uint32_t ro[32];
                    ptx::tmem_ld_32x32b_x32(o_addr, ro);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
                    tmem_st_32x32b_x32(o_addr, ro);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    ptx::tc_fence_before();
                }
                ATTN_STAMP((warp == 0 ? 1 : (warp == 7 ? 2 : -1)), g, 2);
                const float neg_m = -m_run;
                float l_blk = 0.0f, l_blk1 = 0.0f;
                uint8_t *prow = s.p[g & 1] + row * 128;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    float pv[8];
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {          // packed fp32: one FFMA2 + one FADD2 per two scores
                        float x0, x1;
                        ffma2(x0, x1, sc[cc * 8 + e], sc[cc * 8 + e + 1], scale_log2, neg_m);
                        pv[e] = ex2_approx(x0);
                        pv[e + 1] = ex2_approx(x1);
                        fadd2(l_blk, l_blk1, pv[e], pv[e + 1]);
                    }
                    if (DROP) {                  // the softmax denominator is of the un-dropped row; only P.V sees the mask
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            const uint32_t bits = drop_bits(dkey, drow + (uint32_t)((kbase + cc * 8) >> 1) + e2);
                            if (!drop_keep_lo(bits, drop.thresh)) pv[2 * e2] = 0.0f;
                            if (!drop_keep_hi(bits, drop.thresh)) pv[2 * e2 + 1] = 0.0f;
                        }
                    }
                    uint4 pk;
                    pk.x = pack_bf16x2(pv[0], pv[1]);
                    pk.y = pack_bf16x2(pv[2], pv[3]);
                    pk.z = pack_bf16x2(pv[4], pv[5]);
                    pk.w = pack_bf16x2(pv[6], pv[7]);
                    *reinterpret_cast<uint4 *>(prow + (((half * 4 + cc) ^ (row & 7)) << 4)) = pk;
                }
                l_run += l_blk + l_blk1;            // partial row sum over this thread's columns (both halves share m_run)
                ptx::fence_proxy_async_smem();      // generic-proxy writes -> async proxy (tensor core)
                ptx::mbar_arrive(&s.bar_p[g & 1]);
                ATTN_STAMP((warp == 0 ? 1 : (warp == 7 ? 2 : -1)), g, 3);
                if (j == 0 && pend) {            // the previous item's read-out, now that this item's first block is under way
                    read_out(true, pend_o, n_item - 1, s.p[pend_g & 1], pend_m, pend_l, pend_qrow, pend_row0, pend_h, pend_r);
                    pend = false;
                }
            }
            if (nkb > 0) {
                pend = true;
                pend_m = m_run; pend_l = l_run; pend_qrow = qrow; pend_row0 = row0; pend_h = h; pend_r = r;
                pend_o = o_addr; pend_g = g - 1;
                ++n_item;
            } else {                             // window without a valid key: zeros (flush the outstanding read-out first)
                if (pend) {
                    read_out(true, pend_o, n_item - 1, s.p[pend_g & 1], pend_m, pend_l, pend_qrow, pend_row0, pend_h, pend_r);
                    pend = false;
                }
                read_out(false, 0u, 0u, s.p[g & 1], m_run, l_run, qrow, row0, h, r);   // every earlier P.V has retired
            }
        }
        if (pend) read_out(true, pend_o, n_item - 1, s.p[pend_g & 1], pend_m, pend_l, pend_qrow, pend_row0, pend_h, pend_r);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kAttnTmemCols>(tmem_base);
    }
}

}  // namespace kbner

using namespace kbner;

extern "C" int kbner_attention_fwd_ex(const uint16_t *qkv, const int32_t *key_len, int R, int S, int heads,
                                      uint16_t *out, int ldo, uint16_t *out_lo, uint16_t *out_hi2, float *lse,
                                      const uint32_t *drop_seed, uint32_t drop_site, float drop_p, void *stream) {
    KBNER_NVTX("kbner/attention");
    KBNER_CHECK_ARG(qkv && key_len && out, "attention_fwd: null pointer");
    KBNER_CHECK_ARG(ldo >= heads * kAttnD && ldo % 8 == 0 && (((uintptr_t)out | (uintptr_t)out_lo | (uintptr_t)out_hi2) & 15u) == 0,
                    "attention_fwd: output row stride %d / 16-byte alignment of out, out_lo, out_hi2", ldo);
    KBNER_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f, "attention_fwd: dropout probability %f", (double)drop_p);
    KBNER_CHECK_ARG(!(drop_seed && drop_p > 0.0f) || (uint64_t)R * heads * kMaxS * (kMaxS / 2) < (1ull << 32),
                    "attention_fwd: R*heads exceeds the 32-bit dropout counter");
    KBNER_CHECK_ARG(R > 0 && S > 0 && heads > 0, "attention_fwd: empty problem");
    KBNER_CHECK_ARG(S <= kMaxS, "attention_fwd: S=%d exceeds the %d-sub-token window of XLM-R", S, kMaxS);
    const int H = heads * kAttnD;
    CUtensorMap tmQ, tmKV;
    int rc = make_tmap_bf16_2d(&tmQ, qkv, (uint64_t)R * S, (uint64_t)3 * H, (uint64_t)3 * H, kBQ, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmKV, qkv, (uint64_t)R * S, (uint64_t)3 * H, (uint64_t)3 * H, kBKV, 64);
    if (rc) return rc;
    const size_t smem = sizeof(AttnSmem);
    const Dropout drop = make_dropout(drop_seed, drop_site, drop_p);
    static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attention_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("attention_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return KBNER_ECUDA;
        }
        configured = true;
    }
    // persistent: two CTAs per SM, each walking work items (window, head, query block) blockIdx.x, blockIdx.x + grid, ...
    // One CTA per item spent 41 % of its life outside the key-block loop (3.9 k cycles until the first S was ready, ~5 k
    // for the last P.V, the read-out, teardown and the next launch) -- measured with clock64 stamps, see profiles/README.md.
    const long long total = (long long)R * heads * ((S + kBQ - 1) / kBQ);
    KBNER_CHECK_ARG(total < (1ll << 30), "attention_fwd: too many work items");
    const long long resident = 2ll * sm_budget();
    dim3 grid((unsigned)(total < resident ? total : resident));
    cudaError_t le;
    if (drop.thresh)
        le = launch_kernel(attention_fwd_kernel<true>, grid, dim3(kAttnThreads), smem, (cudaStream_t)stream, 0, true, tmQ, tmKV,
                           key_len, R, S, H, heads, out, ldo, out_lo, out_hi2, lse, drop);
    else
        le = launch_kernel(attention_fwd_kernel<false>, grid, dim3(kAttnThreads), smem, (cudaStream_t)stream, 0, true, tmQ, tmKV,
                           key_len, R, S, H, heads, out, ldo, out_lo, out_hi2, lse, drop);
    if (le != cudaSuccess) {
        set_error("attention_fwd: launch failed: %s", cudaGetErrorString(le));
        return KBNER_ECUDA;
    }
    KBNER_CHECK_LAUNCH("attention_fwd");
    return KBNER_OK;
}

#ifdef KBNER_ATTN_DEBUG
extern "C" int kbner_attention_debug_read(unsigned long long *host, int n) {
    KBNER_NVTX("kbner/attention");
    return (int)cudaMemcpyFromSymbol(host, g_attn_dbg, sizeof(unsigned long long) * (size_t)n);
}
#endif

extern "C" int kbner_attention_fwd_dropout(const uint16_t *qkv, const int32_t *key_len, int R, int S, int heads,
                                           uint16_t *out, float *lse, const uint32_t *drop_seed, uint32_t drop_site,
                                           float drop_p, void *stream) {
    KBNER_NVTX("kbner/attention");
    return kbner_attention_fwd_ex(qkv, key_len, R, S, heads, out, heads * kAttnD, nullptr, nullptr, lse, drop_seed, drop_site, drop_p,
                                  stream);
}

extern "C" int kbner_attention_fwd(const uint16_t *qkv, const int32_t *key_len, int R, int S, int heads,
                                   uint16_t *out, float *lse, void *stream) {
    KBNER_NVTX("kbner/attention");
    return kbner_attention_fwd_ex(qkv, key_len, R, S, heads, out, heads * kAttnD, nullptr, nullptr, lse, nullptr, 0u, 0.0f, stream);
}
