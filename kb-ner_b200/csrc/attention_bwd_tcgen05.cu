// Fused multi-head self-attention BACKWARD for one encoder layer (head dim 64), flash-style: the [S,S] probability
// matrix is recomputed per tile from Q, K and the saved row log-sum-exp; it never reaches HBM.  The reference gets
// this from autograd through transformers' eager attention (call site /root/reference/flair/embeddings.py:3269 in
// training mode; flair/trainers/finetune_trainer.py:956-957 `loss.backward()`).
//
//   P  = exp(Q.K^T / 8 - LSE)            dV = P^T . dO
//   dP = dO . V^T                        dS = P * (dP - D),   D = rowsum(dO * O)
//   dQ = dS . K / 8                      dK = dS^T . Q / 8
//
// One CTA = one (window r, head h, block j of 128 keys); it loops over the query blocks i of 128 rows.
//   warp 8       TMA + MMA issuer (elected lane).  Per (i, j):
//                  S^T  = K_j . Q_i^T     (M = keys, N = queries; both operands K-major)            -> TMEM
//                  dP^T = V_j . dO_i^T                                                              -> TMEM
//                  ... threads turn them into P^T and dS^T (bf16, shared memory) ...
//                  dV_j += P^T  . dO_i    (A = P^T  K-major from smem,  B = dO_i MN-major in place)  -> TMEM, kept over i
//                  dK_j += dS^T . Q_i     (A = dS^T K-major,            B = Q_i  MN-major in place)  -> TMEM, kept over i
//                  dQ_i  = dS   . K_j     (A = the SAME dS^T tile read MN-major, B = K_j MN-major)   -> TMEM
//   warps 0..7   two threads per key row (= TMEM lane) build P^T / dS^T; two threads per query row read dQ_i out, which
//                is added to the fp32 dQ accumulator in global memory with red.global.add (other key blocks add theirs).
// dK_j / dV_j are written once, as bf16, into the K | V column blocks of the fused dqkv matrix; dQ is converted from
// the fp32 accumulator by `attn_bwd_dq_kernel`.  D is produced by `attn_bwd_prep_kernel`.
#include <math_constants.h>

#include <atomic>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace kbner {

constexpr int kBwdThreads = 288;        // 8 compute warps (two threads per key / query row) + 1 TMA/MMA warp
constexpr uint32_t kT128 = 128 * 64 * 2;        // [128 rows][64 bf16] SWIZZLE_128B tile = 16 KB

struct AttnBwdSmem {
    uint8_t k[kT128];
    uint8_t v[kT128];
    uint8_t q[2][kT128];                 // query-block ring
    uint8_t dO[2][kT128];
    uint8_t pt[2][2][kT128];             // [query-block parity] P^T  [128 keys][128 queries] as two 64-query sub-tiles
    uint8_t dst[2][2][kT128];            // [query-block parity] dS^T, same layout (scaled by 1/8)
                                         // Once the MMAs of block i have retired, pt[i & 1] doubles as the eight per-warp staging tiles
                                         // (32 rows x 128 B, SWIZZLE_128B) of the dQ_i reduce-add, and at the end of dK_j / dV_j.
    alignas(16) float lse2[2][128];      // [query-block parity] -LSE * log2(e) of the block's query rows (-inf beyond the window)
    alignas(16) float dsum[2][128];      // -D of the block's query rows
    uint64_t bar_kv;
    uint64_t qdo_full[2];
    uint64_t bar_sdp;                    // S^T and dP^T ready in TMEM
    uint64_t bar_drain;                  // every thread has loaded its S^T / dP^T columns (256 arrivals): the next block's may be issued
    uint64_t bar_pd;                     // P^T and dS^T written to smem (256 arrivals)
    uint64_t bar_out[2];                 // [query-block parity] dV/dK/dQ MMAs of the block retired (also: ring stage free)
    uint32_t tmem_base;
};

// D[r][h][s] = sum_d dO[s][h*64+d] * O[s][h*64+d]     (one warp per sub-token row; lane pair = one head)
// O_lo (optional): the forward's rounding residual bf16(o - O).  dS = P * (dP - D) subtracts two nearly equal numbers when
// the attention is close to uniform (dP_ij ~ dO_i . mean(V) for every j): D taken from the bf16-rounded O alone is off by
// 2^-9 |dO||O| there, which measured as 11 % relative error on the top layer's dW_q / dW_k at 24 layers, 8 x 512, against
// autograd through the fp32 oracle (tests/test_precision_gpu.py); with O = O_hi + O_lo the residual is 2^-17.
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const uint16_t *__restrict__ O, const uint16_t *__restrict__ O_lo, const uint16_t *__restrict__ dO, int R,
                     int S, int heads, float *__restrict__ D) {
    const int H = heads * 64;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= R * S) return;
    float acc = 0.0f;
    if (lane * 32 < H) {
        const uint16_t *o = O + (size_t)row * H + lane * 32, *d = dO + (size_t)row * H + lane * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 a = ld_nc_v4(o + c * 8), b = ld_nc_v4(d + c * 8);
            float x0, x1, y0, y1;
            unpack_bf16x2(a.x, x0, x1); unpack_bf16x2(b.x, y0, y1); acc = fmaf(x0, y0, fmaf(x1, y1, acc));
            unpack_bf16x2(a.y, x0, x1); unpack_bf16x2(b.y, y0, y1); acc = fmaf(x0, y0, fmaf(x1, y1, acc));
            unpack_bf16x2(a.z, x0, x1); unpack_bf16x2(b.z, y0, y1); acc = fmaf(x0, y0, fmaf(x1, y1, acc));
            unpack_bf16x2(a.w, x0, x1); unpack_bf16x2(b.w, y0, y1); acc = fmaf(x0, y0, fmaf(x1, y1, acc));
            if (O_lo) {
                const uint4 l = ld_nc_v4(O_lo + (size_t)row * H + lane * 32 + c * 8);
                unpack_bf16x2(l.x, x0, x1); unpack_bf16x2(b.x, y0, y1); acc = fmaf(x0, y0, fmaf(x1, y1, acc));
                unpack_bf16x2(l.y, x0, x1); unpack_bf16x2(b.y, y0, y1); acc = fmaf(x0, y0, fmaf(x1, y1, acc));
                unpack_bf16x2(l.z, x0, x1); unpack_bf16x2(b.z, y0, y1); acc = fmaf(x0, y0, fmaf(x1, y1, acc));
                unpack_bf16x2(l.w, x0, x1); unpack_bf16x2(b.w, y0, y1); acc = fmaf(x0, y0, fmaf(x1, y1, acc));
            }
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    const int head = lane >> 1;
    if ((lane & 1) == 0 && head < heads) {
        const int r = row / S, s = row - r * S;
        D[((size_t)r * heads + head) * S + s] = acc;
    }
}

// dqkv[:, 0:H] = bf16(dQ_acc)   (dQ_acc already carries the 1/8 scale)
__global__ void __launch_bounds__(256)
attn_bwd_dq_kernel(const float *__restrict__ dq_acc, int M, int H, uint16_t *__restrict__ dqkv) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= (size_t)M * H) return;
    const size_t row = i / H, col = i - row * H;
    const float4 a = *reinterpret_cast<const float4 *>(dq_acc + i), b = *reinterpret_cast<const float4 *>(dq_acc + i + 4);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w); o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4 *>(dqkv + row * 3 * H + col) = o;
}

// Debug build (KBNER_EXTRA_NVCC_FLAGS=-DKBNER_ATTN_BWD_DEBUG, scripts/attn_bwd_timeline.py): clock64 stamps of the MMA warp and
// of compute warps 0 / 7 of CTAs 0 and 300, per query block.
#ifdef KBNER_ATTN_BWD_DEBUG
__device__ unsigned long long g_attn_bwd_dbg[2 * 3 * 6 * 6];
#define BWD_STAMP(role, blk, ev)                                                                          \
    do {                                                                                                  \
        if (dbg_cta >= 0 && lane == 0 && (blk) < 6)                                                       \
            g_attn_bwd_dbg[((dbg_cta * 3 + (role)) * 6 + (blk)) * 6 + (ev)] = (unsigned long long)clock64(); \
    } while (0)
#else
#define BWD_STAMP(role, blk, ev) do { } while (0)
#endif

// DROP: attention-probability dropout.  The forward multiplied P by mask / (1 - p) before P.V, so
//   dV = (P * mask / (1-p))^T . dO,   dP = (dO . V^T) * mask / (1-p),   dS = P * (dP - D)   with D = rowsum(dO * O)
// (D needs no change: O already is the dropped product).  The mask bits are regenerated from (window, head, query, key).
template <bool DROP>
__global__ void __launch_bounds__(kBwdThreads, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDKV,
                     const int32_t *__restrict__ key_len, const float *__restrict__ lse, const float *__restrict__ Dsum,
                     int S, int H, int heads, float *__restrict__ dq_acc, uint16_t *__restrict__ dqkv, const Dropout drop) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    AttnBwdSmem &s = *reinterpret_cast<AttnBwdSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jb = blockIdx.x, h = blockIdx.y, r = blockIdx.z;
    const int klen = min(key_len[r], S);
    const int row0 = r * S;
    const int nqb = (S + 127) / 128;
    const bool active = jb * 128 < klen;          // key block with at least one valid key
#ifdef KBNER_ATTN_BWD_DEBUG
    const int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int dbg_cta = lin == 0 ? 0 : (lin == 300 ? 1 : -1);
#endif

    if (threadIdx.x == 0) {
        if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();
        ptx::prefetch_tensormap(&tmQKV);
        ptx::prefetch_tensormap(&tmDO);
        ptx::prefetch_tensormap(&tmDQ);
        ptx::prefetch_tensormap(&tmDKV);
        ptx::mbar_init(&s.bar_kv, 1);
        ptx::mbar_init(&s.qdo_full[0], 1);
        ptx::mbar_init(&s.qdo_full[1], 1);
        ptx::mbar_init(&s.bar_sdp, 1);
        ptx::mbar_init(&s.bar_pd, 256);
        ptx::mbar_init(&s.bar_drain, 256);
        ptx::mbar_init(&s.bar_out[0], 1);
        ptx::mbar_init(&s.bar_out[1], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 8) ptx::tmem_alloc<512>(&s.tmem_base);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    const uint32_t t_st = tmem_base, t_dpt = tmem_base + 128, t_dk = tmem_base + 256, t_dv = tmem_base + 320,
                   t_dq0 = tmem_base + 384;        // dQ: two buffers (+64), block i accumulates into t_dq0 + (i & 1) * 64

    if (!active) {
        // every key of this block is padding: dK = dV = 0 for its rows (rows inside the window), no dQ contribution
        if (warp < 4) {
            const int krow = jb * 128 + warp * 32 + lane;
            if (krow < S) {
                uint16_t *o = dqkv + (size_t)(row0 + krow) * 3 * H + h * 64;
#pragma unroll
                for (int i = 0; i < 64; i += 8) {
                    *reinterpret_cast<uint4 *>(o + H + i) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4 *>(o + 2 * H + i) = make_uint4(0, 0, 0, 0);
                }
            }
        }
    } else if (warp == 8) {
        // ===================== TMA + MMA issuer (warp-uniform; elected lane issues) =====================
        constexpr uint32_t idesc_kk = ptx::make_idesc_bf16(128, 128, 0, 0);   // S^T, dP^T : both K-major
        constexpr uint32_t idesc_kmn = ptx::make_idesc_bf16(128, 64, 0, 1);   // dV, dK    : A K-major, B MN-major
        constexpr uint32_t idesc_mnmn = ptx::make_idesc_bf16(128, 64, 1, 1);  // dQ        : both MN-major
        // Shared-memory descriptors: hi word is the same for every operand (SBO 1024 B, version 1, SWIZZLE_128B); the lo
        // word is (address >> 4) | LBO field, so a k-step is an ADD of a constant.  (Building each descriptor with 64-bit
        // shifts and ors made the elected lane spend 1300 cycles issuing the 24 MMAs of a query block --
        // profiles/r02/attn_bwd_timeline_t1.json -- for 768 cycles of tensor work.)
        const uint32_t dhi = 0x40004040u;
        const uint32_t k_lo = ((ptx::smem_u32(s.k) >> 4) & 0x3FFFu) | (1u << 16), v_lo = ((ptx::smem_u32(s.v) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t pt_lo0 = ((ptx::smem_u32(s.pt[0][0]) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t dst_lo0 = ((ptx::smem_u32(s.dst[0][0]) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t dst_mn_lo0 = ((ptx::smem_u32(s.dst[0][0]) >> 4) & 0x3FFFu) | ((kT128 >> 4) << 16);   // MN-major A: 2 query slabs, LBO 16 KB
        auto desc = [&](uint32_t lo, uint32_t byte_off) {
            uint64_t d;
            asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo + (byte_off >> 4)), "r"(dhi));
            return d;
        };
        auto load_qdo = [&](int i) {
            const int st = i & 1;
            ptx::mbar_expect_tx(&s.qdo_full[st], 2 * kT128);
            ptx::tma_load_2d(s.q[st], &tmQKV, &s.qdo_full[st], h * 64, row0 + i * 128);
            ptx::tma_load_2d(s.dO[st], &tmDO, &s.qdo_full[st], h * 64, row0 + i * 128);
        };
        auto issue_sdp = [&](int i) {     // S^T and dP^T of query block i
            const int st = i & 1;
            const uint32_t q_lo = ((ptx::smem_u32(s.q[st]) >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t do_lo = ((ptx::smem_u32(s.dO[st]) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) ptx::mma_f16_ss(t_st, desc(k_lo, kk * 32), desc(q_lo, kk * 32), idesc_kk, kk != 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) ptx::mma_f16_ss(t_dpt, desc(v_lo, kk * 32), desc(do_lo, kk * 32), idesc_kk, kk != 0);
            ptx::mma_commit(&s.bar_sdp);
        };
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&s.bar_kv, 2 * kT128);
            ptx::tma_load_2d(s.k, &tmQKV, &s.bar_kv, H + h * 64, row0 + jb * 128);
            ptx::tma_load_2d(s.v, &tmQKV, &s.bar_kv, 2 * H + h * 64, row0 + jb * 128);
            load_qdo(0);
            if (nqb > 1) load_qdo(1);
        }
        __syncwarp();
        ptx::mbar_wait(&s.bar_kv, 0);
        ptx::mbar_wait(&s.qdo_full[0], 0);
        ptx::tc_fence_after();
        if (ptx::elect_one()) issue_sdp(0);
        __syncwarp();
        // Order inside the tensor pipe per query block i: S^T / dP^T of block i+1 FIRST (their TMEM was drained before
        // bar_pd(i)), then dV / dK / dQ of block i -- so the threads can start on block i+1 while block i's products are
        // still running; P^T / dS^T and dQ are double-buffered by block parity for that.  (With one buffer the threads sat
        // out the 1600 cycles of a block's dV / dK / dQ plus the next S^T / dP^T: profiles/r02/attn_bwd_timeline_t3.json.)
        for (int i = 0; i < nqb; ++i) {
            const int st = i & 1;
            if (i + 1 < nqb) {
                // S^T / dP^T of block i are in the threads' registers (half-way through their element-wise phase): the next
                // block's go into the tensor pipe now, 8 MMAs that shared-memory bandwidth holds at ~660 cycles, instead of
                // after bar_pd(i), where the threads waited for them (850 cycles per block, attn_bwd_timeline_t5.json)
                ptx::mbar_wait(&s.bar_drain, i & 1);
                ptx::mbar_wait(&s.qdo_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
                ptx::tc_fence_after();
                if (ptx::elect_one()) issue_sdp(i + 1);
                __syncwarp();
                BWD_STAMP(0, i, 1);
            }
            ptx::mbar_wait(&s.bar_pd, i & 1);              // P^T / dS^T of block i are in smem
            ptx::tc_fence_after();
            BWD_STAMP(0, i, 0);
            if (ptx::elect_one()) {
                const uint32_t q_lo = ((ptx::smem_u32(s.q[st]) >> 4) & 0x3FFFu) | (1u << 16);
                const uint32_t do_lo = ((ptx::smem_u32(s.dO[st]) >> 4) & 0x3FFFu) | (1u << 16);
                const uint32_t boff = (uint32_t)st * ((2 * kT128) >> 4);              // this block's P^T / dS^T buffer
                const uint32_t pt_lo = pt_lo0 + boff, dst_lo = dst_lo0 + boff, dst_mn_lo = dst_mn_lo0 + boff;
                const uint32_t t_dq = t_dq0 + (uint32_t)st * 64;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {           // K = 128 queries: 8 steps of 16 over the two sub-tiles
                    const uint32_t a_off = (kk >> 2) * kT128 + (kk & 3) * 32;
                    const uint32_t b_off = kk * 16 * 128;  // 16 query rows of the [query][d] tile (MN-major B)
                    ptx::mma_f16_ss(t_dv, desc(pt_lo, a_off), desc(do_lo, b_off), idesc_kmn, (i | kk) != 0);
                    ptx::mma_f16_ss(t_dk, desc(dst_lo, a_off), desc(q_lo, b_off), idesc_kmn, (i | kk) != 0);
                }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {           // dQ_i = dS . K_j : K = 128 keys
                    const uint32_t off = kk * 16 * 128;    // 16 key rows of dS^T (MN-major A: 2 query slabs, LBO 16 KB) / K_j
                    ptx::mma_f16_ss(t_dq, desc(dst_mn_lo, off), desc(k_lo, off), idesc_mnmn, kk != 0);
                }
                ptx::mma_commit(&s.bar_out[st]);
            }
            __syncwarp();
            BWD_STAMP(0, i, 2);
            if (i + 2 < nqb) {                              // refill this ring stage once block i's MMAs retired
                ptx::mbar_wait(&s.bar_out[st], (i >> 1) & 1);
                if (ptx::elect_one()) load_qdo(i + 2);
                __syncwarp();
            }
        }
    } else {
        // ===================== compute warps: two threads per row =====================
        // warps w and w+4 share TMEM lane quarter w; `half` picks which 64 of the 128 query columns of the P^T / dS^T
        // tile (and which 32 of the 64 output columns of dQ / dK / dV) the thread owns.
        const int quarter = warp & 3, half = warp >> 2;
        const int t = quarter * 32 + lane;                // key row of this block (P^T / dS^T) and query row (dQ read-out)
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const bool key_ok = jb * 128 + t < klen;
        const float scale_log2 = 0.125f * 1.4426950408889634f;
        const uint32_t dkey = DROP ? drop_key(drop) : 0u;
        const uint32_t kpair = (uint32_t)(jb * 128 + t) >> 1;          // this thread's key: pair index and half
        const uint32_t dshift = ((jb * 128 + t) & 1) ? 0u : 16u;       // the key's 16-bit half of the mask word, moved to the top
        const uint32_t dthr = drop.thresh << 16;
        const uint32_t dwin = (uint32_t)(r * heads + h) * 512u;        // same counter layout as attention_fwd_kernel
#ifdef KBNER_ATTN_BWD_DEBUG
        const int drole = warp == 0 ? 1 : (warp == 7 ? 2 : 0);
        const int dbg_cta_w = drole ? dbg_cta : -1;
#define BWD_STAMP_C(blk, ev) do { if (dbg_cta_w >= 0 && lane == 0 && (blk) < 6) g_attn_bwd_dbg[((dbg_cta_w * 3 + drole) * 6 + (blk)) * 6 + (ev)] = (unsigned long long)clock64(); } while (0)
#else
#define BWD_STAMP_C(blk, ev) do { } while (0)
#endif
        // -LSE / -D (negated: FMA addends) of a query block: 256 threads = 128 + 128 values.  Block i+1's are requested at the
        // top of block i and parked in shared memory after its element-wise phase, so no block waits for DRAM (staging them
        // at the top of their own block was 23 % of the stall samples of profiles/r01/attn_bwd_ncu_r38.txt).
        const int sidx = threadIdx.x & 127;
        const bool s_is_lse = threadIdx.x < 128;
        const float *stat_src = (s_is_lse ? lse : Dsum) + ((size_t)r * heads + h) * S;
        auto load_stat = [&](int blk) -> float {          // the raw value: nothing may depend on it before store_stat
            const int qi = blk * 128 + sidx;
            return (qi < S) ? __ldg(stat_src + qi) : 0.0f;
        };
        auto store_stat = [&](int blk, float raw) {
            const int qi = blk * 128 + sidx;
            if (s_is_lse) s.lse2[blk & 1][sidx] = (qi < S) ? -raw * 1.4426950408889634f : -CUDART_INF_F;
            else s.dsum[blk & 1][sidx] = -raw;
        };
        store_stat(0, load_stat(0));
        asm volatile("bar.sync 1, 256;" ::: "memory");   // compute warps only
        // dQ_i partial of this key block -> fp32 accumulator: the warp's 32 rows x 32 columns go through a staging tile (row =
        // 128 B, 16-byte chunks XOR-swizzled as SWIZZLE_128B wants them) and leave as ONE TMA reduce-add.  The red.global.add.v4
        // per thread this replaces was 32 separate 16-byte row segments per request and took 2000 cycles per query block
        // (profiles/r02/attn_bwd_timeline_t1.json).  Rows past the window end add zeros.  The tile lives in pt[i & 1]: free
        // since bar_out(i), written again by block i+2 after the wait_group.read + bar.sync at its top.
        auto read_out_dq = [&](int i) {
            const int b = i & 1;
            ptx::mbar_wait(&s.bar_out[b], (i >> 1) & 1);
            ptx::tc_fence_after();
            BWD_STAMP_C(i, 3);
            uint32_t rq[32];
            ptx::tmem_ld_32x32b_x32(t_dq0 + (uint32_t)b * 64 + lane_addr + half * 32, rq);
            ptx::tmem_ld_wait();
            uint8_t *tile = &s.pt[b][0][0] + warp * 4096;
            uint8_t *dstq = tile + lane * 128;
            const bool row_in = i * 128 + t < S;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint4 val = row_in ? make_uint4(rq[e * 4], rq[e * 4 + 1], rq[e * 4 + 2], rq[e * 4 + 3]) : make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(dstq + ((e ^ (lane & 7)) << 4)) = val;
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmDQ)), "r"(ptx::smem_u32(tile)), "r"(h * 64 + half * 32),
                               "r"(row0 + i * 128 + quarter * 32)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ptx::tc_fence_before();
            BWD_STAMP_C(i, 4);
        };
        for (int i = 0; i < nqb; ++i) {
            const int b = i & 1;
            BWD_STAMP_C(i, 0);
            const float stat_next = (i + 1 < nqb) ? load_stat(i + 1) : 0.0f;
            ptx::mbar_wait(&s.bar_sdp, i & 1);
            ptx::tc_fence_after();
            BWD_STAMP_C(i, 1);
            // The 64 elements per thread of this block cost 24 issue slots each in the first version and the eight warps
            // were issue-bound on them (3700 of a block's 6170 cycles, profiles/r02/attn_bwd_timeline_t2.json).  Now:
            //  * packed fp32 pairs for the arithmetic: arg = S scale - LSE (FFMA2), P = p m (FMUL2), t = dP m - D (FFMA2),
            //    dS = (p / 8) t (2 FMUL2) -- -LSE and -D are staged negated, m = keep ? 1/(1-p) : 0;
            //  * the mask bits of a (query, key pair) serve BOTH keys of the pair, i.e. this lane and lane ^ 1: each lane
            //    hashes the query columns of its own parity and fetches the other half with one shuffle;
            //  * key padding zeroes the packed words, not every element.
            const float *nlse2 = s.lse2[b], *ndsum = s.dsum[b];
            const uint64_t scale2 = f2_pack(scale_log2, scale_log2), eighth2 = f2_pack(0.125f, 0.125f);
            const uint32_t odd = lane & 1u;
            // hash input of query column q: ((dwin + i*128 + q) * 256 + kpair) * 0x9E3779B1 + dkey
            const uint32_t hbase = DROP ? ((dwin + (uint32_t)(i * 128) + odd) * 256u + kpair) * 0x9E3779B1u + dkey : 0u;
#pragma unroll
            for (int cl = 0; cl < 2; ++cl) {              // this thread's 2 chunks of 32 query columns
                const int c = half * 2 + cl;
                uint32_t rs[32], rd[32];
                ptx::tmem_ld_32x32b_x32(t_st + lane_addr + c * 32, rs);
                ptx::tmem_ld_32x32b_x32(t_dpt + lane_addr + c * 32, rd);
                if (cl == 0 && i > 0) {
                    // pt[b] is about to be rewritten: the dQ reduce of block i-2 (issued in the middle of block i-1) staged this
                    // warp's tile exactly where the warp now writes its 32 rows of the sub-tile; lse2[b] / dsum[b] were stored
                    // by all threads in block i-1 (hence the barrier).
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                ptx::tmem_ld_wait();
                if (cl == 1) {                            // this thread's last S^T / dP^T columns are in registers
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&s.bar_drain);
                }
                uint8_t *prow = s.pt[b][half] + t * 128, *drow = s.dst[b][half] + t * 128;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {          // 16-byte chunks of 8 queries
                    uint32_t pw[4], dw[4];
                    const float4 nl0 = *reinterpret_cast<const float4 *>(nlse2 + c * 32 + cc * 8),
                                 nl1 = *reinterpret_cast<const float4 *>(nlse2 + c * 32 + cc * 8 + 4);
                    const float4 nd0 = *reinterpret_cast<const float4 *>(ndsum + c * 32 + cc * 8),
                                 nd1 = *reinterpret_cast<const float4 *>(ndsum + c * 32 + cc * 8 + 4);
                    const float nl[8] = {nl0.x, nl0.y, nl0.z, nl0.w, nl1.x, nl1.y, nl1.z, nl1.w};
                    const float nd[8] = {nd0.x, nd0.y, nd0.z, nd0.w, nd1.x, nd1.y, nd1.z, nd1.w};
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        const int col = c * 32 + cc * 8 + e;                 // even query column of the pair (col, col + 1)
                        float a0, a1;
                        f2_unpack(f2_fma(f2_pack(__uint_as_float(rs[cc * 8 + e]), __uint_as_float(rs[cc * 8 + e + 1])), scale2,
                                         f2_pack(nl[e], nl[e + 1])), a0, a1);
                        const uint64_t pe2 = f2_pack(ex2_fast(a0), ex2_fast(a1));
                        uint64_t dp2 = f2_pack(__uint_as_float(rd[cc * 8 + e]), __uint_as_float(rd[cc * 8 + e + 1]));
                        uint64_t p2 = pe2, t2;
                        if (DROP) {
                            // this lane hashes column col + odd, its partner column col + 1 - odd
                            const uint32_t mine = fmix32(hbase + (uint32_t)col * (256u * 0x9E3779B1u));
                            const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
                            const uint32_t b0 = odd ? other : mine, b1 = odd ? mine : other;       // bits of column col, col + 1
                            // 16-bit half of THIS key: (bits >> 16) >= thresh  <=>  bits >= thresh << 16; the low half is shifted up first
                            const float m0 = ((b0 << dshift) >= dthr) ? drop.scale : 0.0f;
                            const float m1 = ((b1 << dshift) >= dthr) ? drop.scale : 0.0f;
                            const uint64_t m2 = f2_pack(m0, m1);
                            p2 = f2_mul(pe2, m2);                                  // P after dropout: the A operand of dV
                            t2 = f2_fma(dp2, m2, f2_pack(nd[e], nd[e + 1]));
                        } else {
                            t2 = f2_add(dp2, f2_pack(nd[e], nd[e + 1]));
                        }
                        float p0, p1, d0, d1;
                        f2_unpack(p2, p0, p1);
                        f2_unpack(f2_mul(f2_mul(pe2, eighth2), t2), d0, d1);
                        pw[e >> 1] = pack_bf16x2(p0, p1);
                        dw[e >> 1] = pack_bf16x2(d0, d1);
                    }
                    const uint4 pk = key_ok ? make_uint4(pw[0], pw[1], pw[2], pw[3]) : make_uint4(0, 0, 0, 0);
                    const uint4 dk = key_ok ? make_uint4(dw[0], dw[1], dw[2], dw[3]) : make_uint4(0, 0, 0, 0);
                    const int chunk = cl * 4 + cc;        // 16-byte chunk inside the 128-byte row of sub-tile `half`
                    const int phys = (chunk ^ (t & 7)) << 4;
                    *reinterpret_cast<uint4 *>(prow + phys) = pk;
                    *reinterpret_cast<uint4 *>(drow + phys) = dk;
                }
                // dQ read-out of the PREVIOUS block (thread = query row, 32-column half), between this block's two column
                // chunks: by now its MMAs have retired (they were issued behind this block's S^T / dP^T), and the reduce-add
                // has the second chunk, the wait for the next S^T / dP^T and the next block's TMEM loads to read its staging
                // tile before pt[b ^ 1] is written again.  (At the end of the block it had ~1000 cycles, and the wait for it
                // grew the element-wise phase from 2800 to 4400 cycles: profiles/r02/attn_bwd_timeline_t4.json.)
                if (cl == 0 && i > 0) read_out_dq(i - 1);
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();
            BWD_STAMP_C(i, 2);
            ptx::mbar_arrive(&s.bar_pd);
            if (i + 1 < nqb) store_stat(i + 1, stat_next);     // read after the next bar.sync
        }
        read_out_dq(nqb - 1);
        // dK_j, dV_j (accumulated over all query blocks; the last bar_out covered them): warps 0..3 take dK, warps 4..7 dV,
        // each thread its key row's 64 columns -> bf16 -> the warp's staging tile -> one TMA store of 32 rows x 64 columns
        // into the K | V column block of dqkv.  (Row-per-thread 16-byte global stores, 32 sectors per request, made this
        // write-out 3300 cycles per CTA.)  A warp whose 32 rows reach past the window end stores row by row instead: the TMA
        // box would overwrite rows of the next window.
        // tcgen05.ld is warp-collective: every lane executes it, only the stores are guarded.
        {
            const int krow = jb * 128 + t;
            const int which = half;                       // 0: dK, 1: dV
            uint32_t rr[64];
            ptx::tmem_ld_32x32b_x32((which ? t_dv : t_dk) + lane_addr, *reinterpret_cast<uint32_t (*)[32]>(&rr[0]));
            ptx::tmem_ld_32x32b_x32((which ? t_dv : t_dk) + lane_addr + 32, *reinterpret_cast<uint32_t (*)[32]>(&rr[32]));
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // this warp's dQ reduces have read their tiles
            __syncwarp();
            ptx::tmem_ld_wait();
            uint8_t *tile = &s.pt[nqb & 1][0][0] + warp * 4096;     // (every MMA has retired: both P^T buffers are free)
            uint4 ov[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                ov[e].x = pack_bf16x2(__uint_as_float(rr[e * 8]), __uint_as_float(rr[e * 8 + 1]));
                ov[e].y = pack_bf16x2(__uint_as_float(rr[e * 8 + 2]), __uint_as_float(rr[e * 8 + 3]));
                ov[e].z = pack_bf16x2(__uint_as_float(rr[e * 8 + 4]), __uint_as_float(rr[e * 8 + 5]));
                ov[e].w = pack_bf16x2(__uint_as_float(rr[e * 8 + 6]), __uint_as_float(rr[e * 8 + 7]));
            }
            const int wrow0 = jb * 128 + quarter * 32;    // first key row of this warp
            if (wrow0 + 32 <= S) {
                uint8_t *dsto = tile + lane * 128;
#pragma unroll
                for (int e = 0; e < 8; ++e) *reinterpret_cast<uint4 *>(dsto + ((e ^ (lane & 7)) << 4)) = ov[e];
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&tmDKV)), "r"(ptx::smem_u32(tile)),
                                   "r"((which ? 2 * H : H) + h * 64), "r"(row0 + wrow0)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else if (krow < S) {
                uint16_t *o = dqkv + (size_t)(row0 + krow) * 3 * H + (which ? 2 * H : H) + h * 64;
#pragma unroll
                for (int e = 0; e < 8; ++e) *reinterpret_cast<uint4 *>(o + e * 8) = ov[e];
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");            // the stores have read their tiles
            __syncwarp();
        }
        BWD_STAMP_C(nqb, 0);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace kbner

using namespace kbner;

extern "C" int kbner_attention_bwd_ex(const uint16_t *qkv, const uint16_t *out, const uint16_t *out_lo, const uint16_t *d_out,
                                      const float *lse, const int32_t *key_len, int R, int S, int heads,
                                      float *d_scratch, float *dq_acc, uint16_t *dqkv, const uint32_t *drop_seed,
                                      uint32_t drop_site, float drop_p, void *stream) {
    KBNER_NVTX("kbner/attention_bwd");
    KBNER_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f, "attention_bwd: dropout probability %f", (double)drop_p);
    KBNER_CHECK_ARG(!(drop_seed && drop_p > 0.0f) || (uint64_t)R * heads * 512 * 256 < (1ull << 32),
                    "attention_bwd: R*heads exceeds the 32-bit dropout counter");
    const Dropout drop = make_dropout(drop_seed, drop_site, drop_p);
    KBNER_CHECK_ARG(qkv && out && d_out && lse && key_len && d_scratch && dq_acc && dqkv, "attention_bwd: null pointer");
    KBNER_CHECK_ARG(R > 0 && S > 0 && heads > 0 && S <= 512, "attention_bwd: bad shape R=%d S=%d heads=%d", R, S, heads);
    const int H = heads * 64;
    KBNER_CHECK_ARG(H % 32 == 0 && H / 32 <= 32, "attention_bwd: hidden size %d not supported", H);
    cudaStream_t st = (cudaStream_t)stream;
    const int M = R * S;
    cudaError_t e = cudaMemsetAsync(dq_acc, 0, sizeof(float) * (size_t)M * H, st);
    if (e != cudaSuccess) {
        set_error("attention_bwd: memset: %s", cudaGetErrorString(e));
        return KBNER_ECUDA;
    }
    attn_bwd_prep_kernel<<<(M + 7) / 8, 256, 0, st>>>(out, out_lo, d_out, R, S, heads, d_scratch);
    KBNER_CHECK_LAUNCH("attn_bwd_prep");
    CUtensorMap tmQKV, tmDO;
    int rc = make_tmap_bf16_2d(&tmQKV, qkv, (uint64_t)M, (uint64_t)3 * H, (uint64_t)3 * H, 128, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmDO, d_out, (uint64_t)M, (uint64_t)H, (uint64_t)H, 128, 64);
    if (rc) return rc;
    CUtensorMap tmDQ, tmDKV;
    rc = make_tmap_2d(&tmDQ, dq_acc, (uint64_t)M, (uint64_t)H, (uint64_t)H, 32, 32, 4);
    if (rc) return rc;
    rc = make_tmap_2d(&tmDKV, dqkv, (uint64_t)M, (uint64_t)3 * H, (uint64_t)3 * H, 32, 64, 2);
    if (rc) return rc;
    const size_t smem = sizeof(AttnBwdSmem);
    static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
    if (!configured) {
        e = cudaFuncSetAttribute(attention_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attention_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return KBNER_ECUDA;
        }
        configured = true;
    }
    dim3 grid((S + 127) / 128, heads, R);
    if (drop.thresh)
        attention_bwd_kernel<true><<<grid, kBwdThreads, smem, st>>>(tmQKV, tmDO, tmDQ, tmDKV, key_len, lse, d_scratch, S, H, heads, dq_acc, dqkv, drop);
    else
        attention_bwd_kernel<false><<<grid, kBwdThreads, smem, st>>>(tmQKV, tmDO, tmDQ, tmDKV, key_len, lse, d_scratch, S, H, heads, dq_acc, dqkv, drop);
    KBNER_CHECK_LAUNCH("attention_bwd");
    const size_t n8 = (size_t)M * H / 8;
    attn_bwd_dq_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(dq_acc, M, H, dqkv);
    KBNER_CHECK_LAUNCH("attn_bwd_dq");
    return KBNER_OK;
}

#ifdef KBNER_ATTN_BWD_DEBUG
extern "C" int kbner_attention_bwd_debug_read(unsigned long long *host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_attn_bwd_dbg, sizeof(unsigned long long) * (size_t)n);
}
#endif

extern "C" int kbner_attention_bwd_dropout(const uint16_t *qkv, const uint16_t *out, const uint16_t *d_out,
                                           const float *lse, const int32_t *key_len, int R, int S, int heads,
                                           float *d_scratch, float *dq_acc, uint16_t *dqkv, const uint32_t *drop_seed,
                                           uint32_t drop_site, float drop_p, void *stream) {
    KBNER_NVTX("kbner/attention_bwd");
    return kbner_attention_bwd_ex(qkv, out, nullptr, d_out, lse, key_len, R, S, heads, d_scratch, dq_acc, dqkv, drop_seed, drop_site,
                                  drop_p, stream);
}

extern "C" int kbner_attention_bwd(const uint16_t *qkv, const uint16_t *out, const uint16_t *d_out,
                                   const float *lse, const int32_t *key_len, int R, int S, int heads,
                                   float *d_scratch, float *dq_acc, uint16_t *dqkv, void *stream) {
    KBNER_NVTX("kbner/attention_bwd");
    return kbner_attention_bwd_dropout(qkv, out, d_out, lse, key_len, R, S, heads, d_scratch, dq_acc, dqkv, nullptr, 0u, 0.0f,
                                       stream);
}
