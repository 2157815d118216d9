// Viterbi decode (+ per-step confidence, + S-X padding) of the linear-chain CRF on sm_100a.
// Replaces SequenceTagger._viterbi_decode (/root/reference/flair/models/sequence_tagger_model.py
// :1248-1304) as driven by _obtain_labels (:1193-1210).
//
// Bit-exact with the reference by construction: fp32, c[j][k] = v[k] + A[j][k] (one add), FIRST
// maximal k, then + e[j] (second add); terminal v + A[STOP] with STOP / START forced to -1e12.
//
// Mapping (second design).  The first one -- one lane per tag, 16 shuffles + a (value, index) tournament per step,
// emissions prefetched into registers -- measured 1050 cycles per time step for ONE warp, 173 issued instructions per
// warp-step (2 sentences) and 39 % long-scoreboard stalls (profiles/r01/crf_viterbi_v1_ncu_b4096.txt).  Now:
//   * one warp per block, Q lanes per sentence (32/Q sentences per warp), each lane owns JL consecutive tags
//     (Q = 16, JL = 1 for L <= 16; Q = 32 for L <= 32; two-tags-per-lane variants are kept behind an env knob);
//   * the previous state is all-gathered through a shared-memory staging tile (one STS per lane, K/4 16-byte broadcast
//     LDS) into ABSOLUTE tag order -- the tile is also what the confidences are computed from;
//   * the lane's transition rows sit in registers; c = v + A with packed FADD2, max by a 3-input FMNMX3 tree, the first
//     maximal index by an equality scan that is off the critical path of the recurrence (predicated IMADs: FMA pipe);
//   * emissions stream through a 3-stage cp.async ring (16 steps per stage; 16-byte copies when the rows allow it, else
//     4-byte copies that also follow the remove-X index list): no global load sits on the dependent chain;
//   * the step loop has no data-dependent branch: sentences shorter than the longest one of the warp are masked by
//     selects, the trip count is block-uniform (read back from shared memory) so the warp stays converged;
//   * confidences are finished FP steps at a time, branch-free, at the top of the NEXT block of steps (two staging tiles),
//     so their MUFU latency overlaps the recurrence;
//   * back-pointers are packed (4 or 8 bits) into one 32-bit word per lane per 2..8 steps in shared memory (odd row
//     stride); the back-trace is done in three phases by all Q lanes (segment maps, stitch, replay);
//   * every per-sentence shared region is padded so the sentences of a warp tile the 32 banks.
// Measured (profiles/README.md): 512 x 13, B = 32..512: 274 -> 53 us; B = 4096: 354 -> 145 us; B = 16384: 1076 -> 439 us.
// The kernel is issue/latency bound (about 30 instructions per sentence-step), not HBM bound.
#include <math_constants.h>
#include <stdlib.h>

#include <mutex>
#include <utility>

#include "crf_common.cuh"

namespace kbner {

namespace {

constexpr float kNegV = -1e12f;   // the reference's sentinel (sequence_tagger_model.py:402-410,1252)
constexpr int kVitChunk = 16;     // steps per ring stage
constexpr int kVitStages = 2;    // chunk c+1 lands while chunk c (16 steps x ~200 cycles) is consumed: one stage ahead is enough

// First index k with cc[k] == m, pre-shifted by SH: a descending chain of "if (cc[k] == m) idx = k << SH".  The move is
// written as a predicated IMAD (z is an opaque zero) so that it issues on the FMA pipe: FSETP + SEL + FMNMX3 all sit on
// the half-rate ALU pipe, which bounded the step loop (ALU pipe 42 % vs FMA 14 % busy in the capture).
template <int K, int SH, int k = K - 2>
__device__ __forceinline__ void first_max_scan(uint32_t &idx, const float (&cc)[K], float m, uint32_t z) {
    if constexpr (k >= 0) {
        asm("{.reg .pred p;\n\tsetp.eq.f32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %3, %4;}"
            : "+r"(idx) : "f"(cc[k]), "f"(m), "r"(z), "n"((uint32_t)k << SH));
        first_max_scan<K, SH, k - 1>(idx, cc, m, z);
    }
}

template <int Q, int JL, int K>
struct VitCfg {
    static constexpr int SPW = 32 / Q;                 // sentences per warp
    static constexpr int BITS = (K <= 16) ? 4 : 8;     // bits per back-pointer
    static constexpr int SPWD = 32 / (JL * BITS);      // steps per back-pointer word
    // steps per confidence flush (= unrolled steps).  What a block keeps live decides how many blocks share an SM: with 16
    // unrolled steps the L = 13 kernel needed 168 registers and the L = 29 kernel 254 (8 blocks per SM: 4096 sentences =
    // 3.5 waves); 8 steps (4 for L > 16) bring them to 67 / 86.
    static constexpr int FP = (K > 16) ? 4 : ((Q < 8) ? Q : 8);
    static constexpr int KP = (K + 3) / 4 * 4;         // staging row stride (16-byte rows)
    static_assert(Q * JL >= K && JL <= 2, "every tag needs an owner; a lane stores its JL states with one STS");
    static_assert(kVitChunk % FP == 0 && FP % SPWD == 0, "flush period must tile the chunk and the back-pointer words");
    static_assert(SPWD >= 1 && 8 % SPWD == 0, "a word must hold a whole number of steps");
    static __host__ __device__ int tpad(int T) { return (T + kVitChunk - 1) / kVitChunk * kVitChunk; }
    // Per-sentence regions are padded so that consecutive sentences of a warp start 128/SPW bytes apart modulo the
    // 128-byte bank window: the same-shaped accesses of the SPW sentences then tile the 32 banks instead of piling onto
    // the same ones (13.7 M bank conflicts and an LSU pipe 48 % busy in profiles/r01/crf_viterbi_v3_ncu_b4096.txt).
    static constexpr int PADB = (SPW > 1) ? 128 / SPW : 0;
    static __host__ __device__ int padded(int bytes) { return bytes + (PADB - bytes % 128 + 128) % 128; }
    static __host__ __device__ int chunk_stride(int L) { return padded(kVitChunk * L * 4); }          // ring, per sentence
    static __host__ __device__ int bp_stride(int T) { return padded(tpad(T) / SPWD * RS * 4); }      // per sentence
    static __host__ __device__ int stg_stride() { return padded(2 * FP * KP * 4); }                  // per sentence
    static __host__ __device__ size_t ring_bytes(int L) { return (size_t)kVitStages * SPW * chunk_stride(L); }
    static constexpr int RS = Q + 1;                      // back-pointer row stride in words (odd: see the back-trace)
    static __host__ __device__ size_t bp_bytes(int T) { return (size_t)SPW * bp_stride(T); }
    static __host__ __device__ size_t stg_bytes() { return (size_t)SPW * stg_stride(); }
    static __host__ __device__ size_t path_bytes(int T) { return (size_t)SPW * tpad(T); }
    // The decoded path and the back-trace maps reuse the emission ring (dead once the recurrence is over): what a block
    // keeps resident decides how many blocks share an SM, and 4096 x 512 x 13 must fit in ONE wave (2048 blocks over 148
    // SMs = 13.8 per SM; the first layout needed 19.6 KB per block = 11 per SM = 1.26 waves, profiles/r01/crf_ncu_r43.txt).
    static __host__ __device__ size_t tail_bytes(int T) { return path_bytes(T) + 32 * K; }
    static __host__ __device__ size_t ring_region(int T, int L) {
        const size_t r = ring_bytes(L), t = tail_bytes(T);
        return (r > t ? r : t) + 15 & ~(size_t)15;
    }
    static __host__ __device__ size_t smem_bytes(int T, int L) { return ring_region(T, L) + bp_bytes(T) + stg_bytes() + 64; }
};

template <int Q, int JL, int K>
__global__ void __launch_bounds__(32)
crf_viterbi_kernel(const float *__restrict__ emis, const int32_t *__restrict__ pos,
                   const int32_t *__restrict__ klen, const int32_t *__restrict__ slen,
                   const float *__restrict__ trans, int B, int T, int L, int start, int stop,
                   int x_idx, int vec16, int32_t *__restrict__ tags_out, float *__restrict__ conf_out) {
    using C = VitCfg<Q, JL, K>;
    constexpr int SPW = C::SPW, BITS = C::BITS, SPWD = C::SPWD, FP = C::FP, KP = C::KP, CH = kVitChunk, NST = kVitStages;
    extern __shared__ __align__(16) uint8_t smem[];
    float *ring = reinterpret_cast<float *>(smem);
    uint32_t *bp_all = reinterpret_cast<uint32_t *>(smem + C::ring_region(T, L));
    float *stg_all = reinterpret_cast<float *>(smem + C::ring_region(T, L) + C::bp_bytes(T));
    uint8_t *path_all = smem;                                         // aliases the ring: used after the recurrence only
    uint8_t *maps_all = path_all + C::path_bytes(T);                  // 32 lanes x K bytes
    int *s_ctl = reinterpret_cast<int *>(smem + C::ring_region(T, L) + C::bp_bytes(T) + C::stg_bytes());   // [0] = nmax, [1..SPW] = klen

    const int lane = threadIdx.x;
    const int sub = lane / Q, q = lane % Q;
    const int b = blockIdx.x * SPW + sub;
    const bool valid = b < B;
    const int n = valid ? klen[b] : 0;
    const int ns = valid ? slen[b] : 0;
    const size_t rowbase = (size_t)(valid ? b : 0) * T;
    const int tp = C::tpad(T);
    uint32_t *bp = bp_all + (size_t)sub * (C::bp_stride(T) / 4);
    float *stg = stg_all + sub * (C::stg_stride() / 4);
    uint8_t *path = path_all + (size_t)sub * tp;

    {
        int nm = n;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nm = max(nm, __shfl_xor_sync(0xffffffffu, nm, o));
        if (lane == 0) s_ctl[0] = nm;
        if (q == 0) s_ctl[1 + sub] = n;
    }
    __syncwarp();
    const int nmax = s_ctl[0];                     // block-uniform: the loops below stay convergent
    const int nchunks = (nmax + CH - 1) / CH;

    // default fill (S-X padding of _obtain_labels :1202-1208, -1 beyond the sentence)
    if (valid) {
        for (int t = q; t < T; t += Q) {
            tags_out[rowbase + t] = (t < ns) ? x_idx : -1;
            conf_out[rowbase + t] = (t < ns) ? 1.0f : 0.0f;
        }
    }

    // emission ring: stage (c % NST) holds steps [c*CH, (c+1)*CH) of the SPW sentences, row stride L.  The Q lanes of a
    // sentence copy that sentence's chunk (no index arithmetic beyond a running offset).
    const uint32_t chunk_b = C::chunk_stride(L), stage_b = SPW * chunk_b;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t my_chunk_s = ring_s + sub * chunk_b;
    const float *my_row = emis + rowbase * L;
    auto issue = [&](int c) {
        if (c < nchunks) {
            const uint32_t dst = my_chunk_s + (c % NST) * stage_b;
            const int i0 = c * CH;
            if (vec16) {
                const int lim = n * L - i0 * L;                // floats of this sentence left from the chunk start
                for (int r = q * 4; r < CH * L && r < lim; r += Q * 4) cp_async16(dst + r * 4, my_row + i0 * L + r);
            } else {
                for (int u = 0; u < CH && i0 + u < n; ++u) {
                    const int t = pos ? __ldg(pos + rowbase + i0 + u) : i0 + u;
                    for (int jj = q; jj < L; jj += Q) cp_async4(dst + (u * L + jj) * 4, my_row + (size_t)t * L + jj);
                }
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int c = 0; c < NST - 1; ++c) issue(c);

    // transition rows of this lane; padding rows / columns are -inf so they never win
    float A[JL][K];
#pragma unroll
    for (int jl = 0; jl < JL; ++jl) {
        const int j = q * JL + jl;
#pragma unroll
        for (int k = 0; k < K; ++k) A[jl][k] = (j < L && k < L) ? trans[j * L + k] : -CUDART_INF_F;
    }
    float v[JL];
#pragma unroll
    for (int jl = 0; jl < JL; ++jl) {
        const int j = q * JL + jl;
        v[jl] = (j < L) ? ((j == start) ? 0.0f : kNegV) : -CUDART_INF_F;
    }
    // shared-window byte addresses: this lane's emission inside a ring row (clamped for padding tags: their state is
    // -inf whatever is added), its staging slots, its back-pointer word
    const uint32_t stg_s = (uint32_t)__cvta_generic_to_shared(stg);
    const uint32_t bp_s = (uint32_t)__cvta_generic_to_shared(bp) + q * 4;
    uint32_t eoff[JL];
#pragma unroll
    for (int jl = 0; jl < JL; ++jl) eoff[jl] = sub * chunk_b + min(q * JL + jl, L - 1) * 4;
    const uint32_t L4 = L * 4;

    // staging tiles (two, used alternately by consecutive blocks of FP steps): everything -inf (tags >= L stay -inf for
    // good), then the initial state in the row "before" step 0 = last row of tile 1
    constexpr uint32_t kTile = FP * KP * 4;
    for (int x = q; x < 2 * FP * KP; x += Q) sts_b32(stg_s + x * 4, __float_as_uint(-CUDART_INF_F));
    __syncwarp();
    if (q * JL < KP) {
        if constexpr (JL == 1) sts_b32(stg_s + kTile + ((FP - 1) * KP + q) * 4, __float_as_uint(v[0]));
        else sts_v2(stg_s + kTile + ((FP - 1) * KP + q * 2) * 4, v[0], v[1]);
    }
    __syncwarp();

    // confidence_t = softmax(v_t)[argmax v_t] = 1 / sum_k exp(v_t[k] - max_k v_t[k])  (:1295-1300): lane q finishes step
    // first + q from a staging tile.  Branch-free apart from the store, and called at the TOP of the next block of steps,
    // so that its loads / MUFUs interleave with the recurrence instead of stalling it.
    auto conf_flush = [&](uint32_t tile_s, int first) {
        float r[KP];
        const int qq = (Q > FP) ? (q % FP) : q;
#pragma unroll
        for (int k4 = 0; k4 < KP / 4; ++k4) {
            const float4 x = lds_v4(tile_s + (qq * KP + k4 * 4) * 4);
            r[k4 * 4 + 0] = x.x; r[k4 * 4 + 1] = x.y; r[k4 * 4 + 2] = x.z; r[k4 * 4 + 3] = x.w;
        }
        float rr[K];
#pragma unroll
        for (int k = 0; k < K; ++k) rr[k] = r[k];
        const float m = max_tree<K>(rr);
        float p[K];
#pragma unroll
        for (int k = 0; k < K; ++k) p[k] = ex2_fast((rr[k] - m) * 1.4426950408889634f);
#pragma unroll
        for (int w = 1; w < K; w *= 2)          // pairwise sum: short dependent chain
#pragma unroll
            for (int k = 0; k + w < K; k += 2 * w) p[k] += p[k + w];
        const int sidx = first + q;
        const bool ok = q < FP && sidx >= 0 && sidx < n;
        int t = sidx;
        if (pos && ok) t = __ldg(pos + rowbase + sidx);
        if (ok) conf_out[rowbase + t] = rcp_fast(p[0]);      // sum in [1, K]: rcp.approx is within 1 ulp
    };

    uint32_t bpw = 0;
    const uint32_t zero = (uint32_t)vec16 >> 8;      // 0 at run time, opaque to the compiler (first_max_scan)
    uint32_t cur_s = stg_s, prev_s = stg_s + kTile;  // tile written by this block / by the previous one
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        issue(c + NST - 1);
        cp_async_wait<NST - 1>();
        __syncwarp();
        const uint32_t st_s = ring_s + (c % NST) * stage_b;
#pragma unroll 1
        for (int h = 0; h < CH / FP; ++h) {
            const int i0 = c * CH + h * FP;
            conf_flush(prev_s, i0 - FP);
            static_for<FP>([&](auto U) {
                constexpr int u = decltype(U)::value;
                const int i = i0 + u;
                const bool act = i < n;
                float e[JL];
#pragma unroll
                for (int jl = 0; jl < JL; ++jl) e[jl] = lds_f32(st_s + eoff[jl] + (h * FP + u) * L4);
                // all-gather of the previous state through the staging tile (row u-1; last row of the previous block's
                // tile for u = 0): KP/4 16-byte broadcast reads instead of K shuffles
                float vk[KP];
#pragma unroll
                for (int k4 = 0; k4 < KP / 4; ++k4) {
                    const float4 x = lds_v4((u == 0 ? prev_s + (FP - 1) * KP * 4 : cur_s + (u - 1) * KP * 4) + k4 * 16);
                    vk[k4 * 4 + 0] = x.x; vk[k4 * 4 + 1] = x.y; vk[k4 * 4 + 2] = x.z; vk[k4 * 4 + 3] = x.w;
                }
                static_for<JL>([&](auto JJ) {
                    constexpr int jl = decltype(JJ)::value;
                    float cc[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) cc[k] = vk[k];
#pragma unroll
                    for (int k = 0; k + 1 < K; k += 2) fadd2(cc[k], cc[k + 1], A[jl][k], A[jl][k + 1]);   // one FADD2 per pair
                    if constexpr (K % 2 == 1) cc[K - 1] += A[jl][K - 1];
                    const float m = max_tree<K>(cc);
                    constexpr int SH = ((u % SPWD) * JL + jl) * BITS;
                    uint32_t idx = (uint32_t)(K - 1) << SH;
                    first_max_scan<K, SH>(idx, cc, m, zero);
                    const float nv = m + e[jl];
                    v[jl] = act ? nv : v[jl];
                    bpw |= idx;
                });
                if (q * JL < KP) {                       // JL consecutive floats, 4*JL-byte aligned
                    if constexpr (JL == 1) sts_b32(cur_s + (u * KP + q) * 4, __float_as_uint(v[0]));
                    else sts_v2(cur_s + (u * KP + q * 2) * 4, v[0], v[1]);
                }
                if constexpr (u % SPWD == SPWD - 1) {
                    sts_b32(bp_s + (i / SPWD) * (C::RS * 4), bpw);
                    bpw = 0;
                }
                __syncwarp();
            });
            const uint32_t t = cur_s; cur_s = prev_s; prev_s = t;
        }
    }
    conf_flush(prev_s, nchunks * CH - FP);
    cp_async_wait<0>();
    __syncwarp();              // the ring is dead from here on: path / maps reuse it

    // terminal (:1279-1287): first max of v + A[STOP], with STOP / START forced to -1e12
    float term = -CUDART_INF_F;
    int idx = 0;
#pragma unroll
    for (int jl = 0; jl < JL; ++jl) {
        const int j = q * JL + jl;
        float x = -CUDART_INF_F;
        if (j < L) {
            x = v[jl] + trans[stop * L + j];
            if (j == stop || j == start) x = kNegV;
        }
        if (jl == 0 || x > term) { term = x; idx = j; }
    }
#pragma unroll
    for (int o = Q / 2; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, term, o, Q);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o, Q);
        if (ov > term || (ov == term && oi < idx)) { term = ov; idx = oi; }
    }
    __syncwarp();

    // Back-trace.  bp_i maps the tag at step i to the tag at step i-1.  Lane q owns the segment of steps
    // [q*seg, (q+1)*seg): phase 1 composes its maps for every possible tag at the segment's upper end (K independent
    // pointer chases: latency of one, throughput of K), phase 2 stitches the Q segment maps serially from the terminal
    // tag, phase 3 replays the segment from its now known end tag and records the path.
    auto bp_at = [&](int i, int cur) -> int {
        const uint32_t w = bp[(size_t)(i / SPWD) * C::RS + cur / JL];
        return (w >> (((i % SPWD) * JL + cur % JL) * BITS)) & ((1u << BITS) - 1u);
    };
    // Segment length: a multiple of SPWD with an ODD number of back-pointer rows, so that with the odd row stride the Q
    // lanes (each walking its own segment in lockstep) hit different banks -- with seg = n/Q and stride Q every lane
    // of the warp read the same bank (a 32-way conflict per load; 24 % of the kernel in the first capture).
    int seg = ((n + Q - 1) / Q + SPWD - 1) / SPWD;
    seg = (seg | 1) * SPWD;
    const int lo = min(q * seg, n), hi = min(lo + seg, n);        // this lane's steps [lo, hi)
    uint8_t *maps = maps_all + (size_t)(sub * Q + q) * K;
    {
        int cur[K];
#pragma unroll
        for (int k = 0; k < K; ++k) cur[k] = k;
        for (int i = hi - 1; i >= lo; --i) {
#pragma unroll
            for (int k = 0; k < K; ++k) cur[k] = bp_at(i, cur[k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) maps[k] = (uint8_t)cur[k];
    }
    __syncwarp();
    int endtag = idx;                                             // tag at step hi-1 for this lane's segment
    for (int qq = Q - 1; qq > q; --qq) {
        const int lo2 = min(qq * seg, n), hi2 = min(lo2 + seg, n);
        if (hi2 > lo2) endtag = maps_all[(size_t)(sub * Q + qq) * K + endtag];
    }
    {
        int cur = endtag;
        for (int i = hi - 1; i >= lo; --i) {
            path[i] = (uint8_t)cur;
            cur = bp_at(i, cur);
        }
        // for q == 0, cur is START here for every well-formed transition matrix (reference assert :1303)
    }
    __syncwarp();
    if (valid) {
        for (int i = q; i < n; i += Q) {
            const int t = pos ? __ldg(pos + rowbase + i) : i;
            tags_out[rowbase + t] = path[i];
        }
    }
}

template <int Q, int JL, int K>
int launch_viterbi(const float *emis, const int32_t *pos, const int32_t *klen, const int32_t *slen,
                   const float *trans, int B, int T, int L, int start, int stop, int x_idx,
                   int32_t *tags_out, float *conf_out, cudaStream_t st) {
    using C = VitCfg<Q, JL, K>;
    const size_t smem = C::smem_bytes(T, L);
    if (smem > 200 * 1024) {
        set_error("crf_viterbi: T=%d too long for the shared-memory back-pointer table", T);
        return KBNER_EUNSUPPORTED;
    }
    static std::once_flag once;
    static cudaError_t cfg_err = cudaSuccess;
    std::call_once(once, [] {
        cfg_err = cudaFuncSetAttribute(crf_viterbi_kernel<Q, JL, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
        if (cfg_err == cudaSuccess)      // all of the SM's shared memory to the blocks: residency is what bounds this kernel
            cfg_err = cudaFuncSetAttribute(crf_viterbi_kernel<Q, JL, K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    });
    if (cfg_err != cudaSuccess) {
        set_error("crf_viterbi: cudaFuncSetAttribute: %s", cudaGetErrorString(cfg_err));
        return KBNER_ECUDA;
    }
    // 16-byte copies need: no index list, 16-byte aligned rows (T*L % 4 == 0) and base pointer
    const int vec16 = (pos == nullptr && ((size_t)T * L) % 4 == 0 && (reinterpret_cast<uintptr_t>(emis) & 15) == 0) ? 1 : 0;
    const int blocks = (B + C::SPW - 1) / C::SPW;
    crf_viterbi_kernel<Q, JL, K><<<blocks, 32, smem, st>>>(emis, pos, klen, slen, trans, B, T, L, start, stop, x_idx,
                                                            vec16, tags_out, conf_out);
    KBNER_CHECK_LAUNCH("crf_viterbi");
    return KBNER_OK;
}

}  // namespace
}  // namespace kbner

using namespace kbner;

extern "C" int kbner_crf_viterbi(const float *emis, const int32_t *pos, const int32_t *klen,
                                 const int32_t *slen, const float *trans, int B, int T, int L,
                                 int start_idx, int stop_idx, int x_idx, int32_t *tags_out,
                                 float *conf_out, void *stream) {
    KBNER_NVTX("kbner/crf");
    KBNER_CHECK_ARG(emis && klen && slen && trans && tags_out && conf_out, "crf_viterbi: null pointer");
    KBNER_CHECK_ARG(B >= 0 && T > 0 && L >= 2 && L <= 32, "crf_viterbi: need L in [2,32], got L=%d T=%d", L, T);
    KBNER_CHECK_ARG(start_idx >= 0 && start_idx < L && stop_idx >= 0 && stop_idx < L,
                    "crf_viterbi: start/stop index out of range");
    if (B == 0) return KBNER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // One tag per lane ("wide": Q = 16 or 32) measured faster than two tags per lane at every batch size of the sweep
    // (4096 x 512 x 13: 145 vs 156 us; 16384: 439 vs 462 us); KBNER_VIT_WIDE_MAX=<B> selects the two-tags-per-lane
    // variants above B sentences for experiments.
    static const int wide_max = [] { const char *e = getenv("KBNER_VIT_WIDE_MAX"); return e ? atoi(e) : (1 << 30); }();
    const bool wide = B <= wide_max;
#define KBNER_VIT(Q, JL, K) \
    return launch_viterbi<Q, JL, K>(emis, pos, klen, slen, trans, B, T, L, start_idx, stop_idx, x_idx, tags_out, conf_out, st)
    if (L == 13) { if (wide) KBNER_VIT(16, 1, 13); else KBNER_VIT(8, 2, 13); }
    if (L <= 16) { if (wide) KBNER_VIT(16, 1, 16); else KBNER_VIT(8, 2, 16); }
    if (L == 29) { if (wide) KBNER_VIT(32, 1, 29); else KBNER_VIT(16, 2, 29); }
    if (wide) KBNER_VIT(32, 1, 32); else KBNER_VIT(16, 2, 32);
#undef KBNER_VIT
}
