// Linear-chain CRF on sm_100a: remove-X compaction, log-partition / gold score, and their
// gradient (Viterbi: crf_viterbi.cu).  Replaces the per-token Python loops of
// /root/reference/flair/models/sequence_tagger_model.py (_viterbi_decode :1248-1304,
// _forward_alg :1329-1394, _score_sentence :2544-2591, _calculate_loss :2448-2506,
// _obtain_labels :1193-1210).
//
// Mapping: one warp per block, Q lanes per sentence (Q = 16 when L <= 16: two sentences per warp; else 32), lane j owns
// tag j, its transition row in registers.  A step is ONE exchange: the lane publishes its state in a two-row shared-memory
// tile, every lane gathers the K states with broadcast loads and reduces them itself; what a step reads from global
// memory streams through a cp.async ring.  These kernels are bound by the per-step instruction count / the dependent
// chain, not by HBM: their traffic is the emissions (and, for the gradient, the stored alpha) read once ("CRF").
#include <math_constants.h>

#include <atomic>

#include "crf_common.cuh"

namespace kbner {

constexpr float kNeg = -1e12f;  // the reference's sentinel (sequence_tagger_model.py:402-410,1252)

// ------------------------------------------------------------------------------------------
// compaction: one warp per sentence, ballot + popc prefix sum
// ------------------------------------------------------------------------------------------
__global__ void crf_compact_kernel(const uint8_t *__restrict__ keep, int B, int T,
                                   int32_t *__restrict__ pos, int32_t *__restrict__ klen) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const uint8_t *kr = keep + (size_t)warp * T;
    int32_t *pr = pos + (size_t)warp * T;
    int n = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const bool k = (t < T) && kr[t] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (k) pr[n + __popc(m & ((1u << lane) - 1u))] = t;
        n += __popc(m);
    }
    for (int t = n + lane; t < T; t += 32) pr[t] = -1;
    if (lane == 0) klen[warp] = n;
}

template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, G));
    return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
    return v;
}
template <int G>
__device__ __forceinline__ int group_max_int(int v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o, G));
    return v;
}
__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------
// log Z and gold score.  The reference evaluates alpha'[j] = max_k x + log sum_k exp(x - max),
// x = (e[j] + A[j][k]) + alpha[k]  (L^2 exps per step).  Here the same quantity is carried as
//   alpha_t[j] = S_t + a_t[j],   S_t accumulated in fp64,
//   m = max_k a_t[k],  p[k] = exp(a_t[k]),  E[j][k] = exp(A[j][k] - rmax_j)
//   a_{t+1}[j] = (e[j] + rmax_j) + log sum_k E[j][k] * p[k] - m,      S_{t+1} = S_t + m
// (a_{t+1} <= e + rmax + log L: bounded by ONE step's growth, so p stays finite for |e + rmax| < 80).
// Keeping the O(T) magnitude in a separate fp64 scalar is what keeps the marginals of the backward pass accurate at
// T = 512 (alpha ~ 2000 would otherwise carry ~1e-4 absolute error per step into exp(alpha + beta - logZ)).
//
// Mapping (second design, same as the Viterbi kernel): one warp per block, Q lanes per sentence, lane j owns tag j; the
// lane publishes the pair (a[j], exp(a[j])) in a two-row shared-memory tile (one STS.64, K/2 broadcast LDS.128 per
// step) and EVERY lane takes the max of the gathered a itself, so a step needs one exchange, one exp, one log and no
// cross-lane reduction; emissions stream through a cp.async ring.  The first design (16 shuffles + a 4-level shuffle max per step,
// register-prefetched emissions) took 1075 cycles per step for one warp.
// The normalised pair the backward pass consumes (alpha = a - max a, scale = S + max a) falls out one step later, when
// the max of the stored vector has been computed anyway.
// ------------------------------------------------------------------------------------------
constexpr int kNllChunk = 16;     // steps per ring stage
constexpr int kNllStages = 3;

template <int Q, int K>
struct NllCfg {
    static constexpr int SPW = 32 / Q;
    static constexpr int KP = (K + 3) / 4 * 4;
    static constexpr int PADB = (SPW > 1) ? 128 / SPW : 0;     // see VitCfg: the sentences of a warp tile the banks
    static __host__ __device__ int padded(int bytes) { return bytes + (PADB - bytes % 128 + 128) % 128; }
    static __host__ __device__ int chunk_stride(int L) { return padded(kNllChunk * L * 4); }
    static __host__ __device__ int tile_stride() { return padded(2 * 2 * KP * 4); }   // 2 rows of KP (a, p) pairs
    static __host__ __device__ size_t ring_bytes(int L) { return (size_t)kNllStages * SPW * chunk_stride(L); }
    static __host__ __device__ size_t smem_bytes(int L) { return ring_bytes(L) + (size_t)SPW * tile_stride() + 64; }
};

template <int Q, int K>
__global__ void __launch_bounds__(32)
crf_nll_fwd_kernel(const float *__restrict__ emis, const int32_t *__restrict__ tags,
                   const int32_t *__restrict__ pos, const int32_t *__restrict__ klen,
                   const float *__restrict__ trans, int B, int T, int L, int start, int stop, int vec16,
                   float *__restrict__ logz, float *__restrict__ gold, float *__restrict__ alpha_out,
                   double *__restrict__ ascale_out) {
    using C = NllCfg<Q, K>;
    constexpr int SPW = C::SPW, KP = C::KP, CH = kNllChunk, NST = kNllStages, UN = 8;
    extern __shared__ __align__(16) uint8_t smem[];
    int *s_ctl = reinterpret_cast<int *>(smem + C::ring_bytes(L) + (size_t)SPW * C::tile_stride());
    const int lane = threadIdx.x;
    const int sub = lane / Q, j = lane % Q;
    const int b = blockIdx.x * SPW + sub;
    const bool valid = b < B;
    const int n = valid ? klen[b] : 0;
    const size_t rowbase = (size_t)(valid ? b : 0) * T;
    {
        const int nm = warp_max_int(n);
        if (lane == 0) s_ctl[0] = nm;
    }
    __syncwarp();
    const int nmax = s_ctl[0];                     // block-uniform trip count: the warp stays converged
    const int nchunks = (nmax + CH - 1) / CH;

    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t chunk_b = C::chunk_stride(L), stage_b = SPW * chunk_b;
    const uint32_t my_chunk_s = ring_s + sub * chunk_b;
    const uint32_t tile_s = ring_s + (uint32_t)C::ring_bytes(L) + sub * C::tile_stride();
    const float *my_row = emis + rowbase * L;
    auto issue = [&](int c) {
        if (c < nchunks) {
            const uint32_t dst = my_chunk_s + (c % NST) * stage_b;
            const int i0 = c * CH;
            if (vec16) {
                const int lim = n * L - i0 * L;
                for (int r = j * 4; r < CH * L && r < lim; r += Q * 4) cp_async16(dst + r * 4, my_row + i0 * L + r);
            } else {
                for (int u = 0; u < CH && i0 + u < n; ++u) {
                    const int t = pos ? __ldg(pos + rowbase + i0 + u) : i0 + u;
                    for (int jj = j; jj < L; jj += Q) cp_async4(dst + (u * L + jj) * 4, my_row + (size_t)t * L + jj);
                }
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int c = 0; c < NST - 1; ++c) issue(c);

    float rmax = -CUDART_INF_F;
    if (j < L)
        for (int k = 0; k < L; ++k) rmax = fmaxf(rmax, trans[j * L + k]);
    float E[K];
#pragma unroll
    for (int k = 0; k < K; ++k) E[k] = (j < L && k < L) ? expf(trans[j * L + k] - rmax) : 0.0f;
    if (j >= L) rmax = 0.0f;

    float a = (j < L) ? ((j == start) ? 0.0f : kNeg) : -CUDART_INF_F;
    double S = 0.0;
    // tile: two rows of KP (a, exp(a)) pairs, (-inf, 0) in the padding columns; step i writes row i & 1, reads the other
    constexpr uint32_t kRow = 2 * KP * 4;
    for (int x = j; x < 2 * KP; x += Q) sts_v2(tile_s + x * 8, -CUDART_INF_F, 0.0f);
    __syncwarp();
    if (j < KP) sts_v2(tile_s + kRow + j * 8, a, (j < L) ? expf(a) : 0.0f);
    __syncwarp();
    const uint32_t eoff = sub * chunk_b + min(j, L - 1) * 4;
    const uint32_t L4 = L * 4;

    // top of step i: gather (alpha_i, exp alpha_i), the max, and -- one step late -- the normalised record of step i - 1
    auto gather = [&](int row, int i, float (&pk)[K]) -> float {
        float ak[K];
#pragma unroll
        for (int k2 = 0; k2 < (K + 1) / 2; ++k2) {
            const float4 x = lds_v4(tile_s + row * kRow + k2 * 16);
            ak[2 * k2] = x.x; pk[2 * k2] = x.y;
            if (2 * k2 + 1 < K) { ak[2 * k2 + 1] = x.z; pk[2 * k2 + 1] = x.w; }
        }
        const float m = max_tree<K>(ak);
        if (alpha_out && i >= 1 && i - 1 < n) {
            if (j < L) alpha_out[(rowbase + i - 1) * L + j] = a - m;
            if (j == 0) ascale_out[rowbase + i - 1] = S + (double)m;
        }
        return m;
    };

#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        issue(c + NST - 1);
        cp_async_wait<NST - 1>();
        __syncwarp();
        const uint32_t st_s = ring_s + (c % NST) * stage_b;
#pragma unroll 1
        for (int h = 0; h < CH / UN; ++h) {
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = c * CH + h * UN + u;
                const float e = lds_f32(st_s + eoff + (h * UN + u) * L4);
                float pk[K];
                const float m = gather((u & 1) ^ 1, i, pk);
                float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int k = 0; k < K; ++k) s4[k & 3] = fmaf(E[k], pk[k], s4[k & 3]);
                const float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                const float r = (j < L) ? ((e + rmax) + __logf(s)) - m : -CUDART_INF_F;   // lg2.approx: |err| < 2e-7
                if (i < n) {
                    a = r;
                    S += (double)m;
                }
                if (j < KP) sts_v2(tile_s + (u & 1) * kRow + j * 8, a, ex2_fast(a * 1.4426950408889634f));
                __syncwarp();
            }
        }
    }
    cp_async_wait<0>();
    {
        float pk[K];
        gather(1, nchunks * CH, pk);        // record of the last step of full-length sentences (row 1: CH is even)
    }

    // terminal: log_sum_exp_batch(alpha_len + A[STOP])  (:1381-1392)
    const float x = (j < L) ? a + trans[stop * L + j] : -CUDART_INF_F;
    const float M2 = group_max<Q>(x);
    const float s2 = group_sum<Q>(expf(x - M2));
    // gold score (:2544-2591): lanes stride over the kept tokens
    float g = 0.0f;
    for (int i = j; i < n; i += Q) {
        const int t = pos ? pos[rowbase + i] : i;
        const int y = tags[rowbase + t];
        int prev = start;
        if (i > 0) {
            const int tp = pos ? pos[rowbase + i - 1] : i - 1;
            prev = tags[rowbase + tp];
        }
        g += emis[(rowbase + t) * L + y] + trans[y * L + prev];
    }
    g = group_sum<Q>(g);
    if (valid && j == 0) {
        int last = start;
        if (n > 0) {
            const int tl = pos ? pos[rowbase + n - 1] : n - 1;
            last = tags[rowbase + tl];
        }
        logz[b] = (float)(S + (double)M2 + (double)logf(s2));
        gold[b] = g + trans[stop * L + last];
    }
}

// ------------------------------------------------------------------------------------------
// gradient of sum_b w[b] (logZ_b - gold_b).  Reverse (beta) recursion in the same form as the forward pass:
//   beta_i[k] = Sb_i + b_i[k],  u[j] = e_i[j] + b_{i+1}[j] + rmax_j,  M = max_j u[j],
//   b_i[k] = log sum_j E[j][k] * exp(u[j]) - M,   Sb_i = Sb_{i+1} + M          (Sb in fp64)
// unary marginals -> d_emis; pairwise marginals accumulated per lane as  acc[j][k] += c * exp(u[j]) * exp(ahat_i[k])
// (dT[j][k] = E[j][k] * acc[j][k], c = w * exp(Sa_i + Sb_{i+1} - logZ)), persistent over the sentences of a block.
//
// Mapping (second design): one warp per block, Q lanes per sentence, lane j owns tag j and publishes the triple
// (u[j], exp u[j], exp ahat_i[j]) in a two-row shared-memory tile (one STS.128, K broadcast LDS.128 per step); every
// lane takes the max itself, so a step is ONE exchange.  Everything a step reads from global memory -- the emission row
// (through the remove-X index list), the stored alpha row and scale, the gold tags -- is staged by cp.async into a
// 3-stage ring, 16 steps per stage, walked backwards; lane q of a sentence fetches entry q of a chunk and the index-list
// entries it needs are prefetched one chunk ahead.  The first design loaded all of this inside the step (a global
// round trip on the dependent chain of every step) and reduced twice per step across lanes with shuffles.
// ------------------------------------------------------------------------------------------
template <int Q, int K>
struct NllBwdCfg {
    static constexpr int SPW = 32 / Q;
    static constexpr int CH = 16, NST = 3;
    static constexpr int PADB = (SPW > 1) ? 128 / SPW : 0;
    static __host__ __device__ int padded(int bytes) { return bytes + (PADB - bytes % 128 + 128) % 128; }
    // per sentence and stage: SAP[CH] f64 | E[CH][L] f32 | AP[CH][L] f32 | YP[CH] i32 | TL[CH] i32
    static __host__ __device__ int off_e() { return CH * 8; }
    static __host__ __device__ int off_ap(int L) { return off_e() + CH * L * 4; }
    static __host__ __device__ int off_yp(int L) { return off_ap(L) + CH * L * 4; }
    static __host__ __device__ int off_tl(int L) { return off_yp(L) + CH * 4; }
    static __host__ __device__ int chunk_stride(int L) { return padded(off_tl(L) + CH * 4); }
    static __host__ __device__ int tile_stride() { return padded(2 * K * 16); }
    static __host__ __device__ size_t ring_bytes(int L) { return (size_t)NST * SPW * chunk_stride(L); }
    static __host__ __device__ size_t dt_off(int L) { return ring_bytes(L) + (size_t)SPW * tile_stride(); }
    // dT partials [L][L] | control word (64 B) | gold transition counts [SPW][L][L]
    static __host__ __device__ size_t smem_bytes(int L) { return dt_off(L) + (size_t)(1 + SPW) * L * L * 4 + 64; }
};

__device__ __forceinline__ void cp_async8(uint32_t dst_s, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst_s), "l"(src) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

template <int Q, int K>
__global__ void __launch_bounds__(32)
crf_nll_bwd_kernel(const float *__restrict__ emis, const int32_t *__restrict__ tags,
                   const int32_t *__restrict__ pos, const int32_t *__restrict__ klen,
                   const float *__restrict__ trans, const float *__restrict__ alpha,
                   const double *__restrict__ ascale, const float *__restrict__ w, int B, int T, int L,
                   int start, int stop, float *__restrict__ d_emis, float *__restrict__ d_trans) {
    using C = NllBwdCfg<Q, K>;
    constexpr int SPW = C::SPW, CH = C::CH, NST = C::NST;
    extern __shared__ __align__(16) uint8_t smem[];
    float *s_dt = reinterpret_cast<float *>(smem + C::dt_off(L));
    int *s_ctl = reinterpret_cast<int *>(smem + C::dt_off(L) + (size_t)L * L * 4);
    const int lane = threadIdx.x;
    const int sub = lane / Q, j = lane % Q;
    // gold transition counts: lane 0 of a sentence owns a private [L][L] table and updates it with a plain load / add /
    // store per step (an fp32 shared-memory atomicAdd is a compare-and-swap loop, and it sat on every step of the chain)
    float *s_gold = reinterpret_cast<float *>(smem + C::dt_off(L) + (size_t)L * L * 4 + 64) + (size_t)sub * L * L;
    for (int i = lane; i < (1 + SPW) * L * L + 16; i += 32) s_dt[i] = 0.0f;

    float rmax = -CUDART_INF_F;
    if (j < L)
        for (int k = 0; k < L; ++k) rmax = fmaxf(rmax, trans[j * L + k]);
    // Ecol[jj] = E[jj][j]: column j of the row-normalised exp(A)
    float Ecol[K];
#pragma unroll
    for (int jj = 0; jj < K; ++jj) {
        float v = 0.0f;
        if (j < L && jj < L) {
            float rm = -CUDART_INF_F;
            for (int k = 0; k < L; ++k) rm = fmaxf(rm, trans[jj * L + k]);
            v = expf(trans[jj * L + j] - rm);
        }
        Ecol[jj] = v;
    }
    if (j >= L) rmax = 0.0f;
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0f;

    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t chunk_b = C::chunk_stride(L), stage_b = SPW * chunk_b;
    const uint32_t my_chunk_s = ring_s + sub * chunk_b;
    const uint32_t tile_s = ring_s + (uint32_t)C::ring_bytes(L) + sub * C::tile_stride();
    const uint32_t oE = C::off_e(), oAP = C::off_ap(L), oYP = C::off_yp(L), oTL = C::off_tl(L);
    const uint32_t jc4 = min(j, L - 1) * 4, L4 = L * 4;
    constexpr uint32_t kRow = K * 16;

    const int groups_total = gridDim.x * SPW;
    const int iters = (B + groups_total - 1) / groups_total;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const int b = (blockIdx.x + it * gridDim.x) * SPW + sub;
        const bool valid = b < B;
        const int n = valid ? klen[b] : 0;
        const size_t rowbase = (size_t)(valid ? b : 0) * T;
        const float wb = valid ? w[b] : 0.0f;
        __syncwarp();
        {
            const int nm = warp_max_int(n);
            if (lane == 0) s_ctl[0] = nm;
        }
        __syncwarp();
        const int nmax = s_ctl[0];               // block-uniform
        if (nmax == 0) continue;
        const int nchunks = (nmax + CH - 1) / CH;

        // beta_n[k] = A[STOP][k], normalised; logZ re-derived in fp64 from the stored alpha
        const float astop = (j < L) ? trans[stop * L + j] : -CUDART_INF_F;
        const float mb0 = group_max<Q>(astop);
        float beta = astop - mb0;
        double Sb = (double)mb0;
        float anext = (n > 0 && j < L) ? alpha[(rowbase + n - 1) * L + j] : -CUDART_INF_F;     // ahat_{i+1} of the step at hand
        double Sa_next = (n > 0) ? ascale[rowbase + n - 1] : 0.0;
        const float xt = (j < L) ? anext + astop : -CUDART_INF_F;
        const float Mt = group_max<Q>(xt);
        const float st = group_sum<Q>(expf(xt - Mt));
        const double lz = Sa_next + (double)Mt + (double)logf(st);
        int ycur = start;
        if (n > 0) {
            const int tl = pos ? pos[rowbase + n - 1] : n - 1;
            ycur = tags[rowbase + tl];
            // terminal pairwise term dT[STOP][k] += w * exp(alpha_n[k] + A[STOP][k] - logZ), and its gold count
            if (j < L) atomicAdd(&s_dt[stop * L + j], wb * expf(xt - Mt) / st);
            if (j == 0) atomicAdd(&s_dt[stop * L + ycur], -wb);
        }

        // ring: chunk c = steps [c*CH, (c+1)*CH), walked from the last chunk down.  Lane q < CH of a sentence fetches
        // entry q; (pf_t, pf_tm1) are the index-list entries of the NEXT chunk to issue, prefetched a chunk ahead.
        auto index_of = [&](int i) -> int { return (i >= 0 && i < n) ? (pos ? __ldg(pos + rowbase + i) : i) : 0; };
        int pf_t = 0, pf_tm1 = 0;
        auto prefetch_idx = [&](int c) {
            if (c >= 0 && j < CH) { pf_t = index_of(c * CH + j); pf_tm1 = index_of(c * CH + j - 1); }
        };
        auto issue = [&](int cc) {               // cc counts issued chunks; chunk number c = nchunks - 1 - cc
            const int c = nchunks - 1 - cc;
            if (c >= 0 && j < CH) {
                const uint32_t dst = my_chunk_s + (cc % NST) * stage_b;
                const int i = c * CH + j;
                if (i < n) {
                    const float *er = emis + (rowbase + pf_t) * L;
                    for (int jj = 0; jj < L; ++jj) cp_async4(dst + oE + (j * L + jj) * 4, er + jj);
                    sts_b32(dst + oTL + j * 4, (uint32_t)pf_t);
                    if (i >= 1) {
                        const float *ar = alpha + (rowbase + i - 1) * L;
                        for (int jj = 0; jj < L; ++jj) cp_async4(dst + oAP + (j * L + jj) * 4, ar + jj);
                        cp_async8(dst + j * 8, ascale + rowbase + i - 1);
                        cp_async4(dst + oYP + j * 4, tags + rowbase + pf_tm1);
                    }
                }
            }
            cp_async_commit();
            prefetch_idx(c - 1);
        };
        prefetch_idx(nchunks - 1);
#pragma unroll
        for (int cc = 0; cc < NST - 1; ++cc) issue(cc);

        // tile: two rows of K (u, exp u, exp ahat, -) entries; row parity alternates per step
        int par = 0;
#pragma unroll 1
        for (int cc = 0; cc < nchunks; ++cc) {
            const int c = nchunks - 1 - cc;
            issue(cc + NST - 1);
            cp_async_wait<NST - 1>();
            __syncwarp();
            const uint32_t stg = my_chunk_s + (cc % NST) * stage_b;
#pragma unroll 4
            for (int uu = 0; uu < CH; ++uu) {
                const int u = CH - 1 - uu;
                const int i = c * CH + u;
                const bool act = i < n;
                // staged inputs of step i (sanitised when the step is not part of this sentence)
                float e = lds_f32(stg + oE + u * L4 + jc4);
                float ap = lds_f32(stg + oAP + u * L4 + jc4);
                double Sa_prev = lds_f64(stg + u * 8);
                int yprev = lds_s32(stg + oYP + u * 4);
                const int t = lds_s32(stg + oTL + u * 4);
                if (i == 0) {
                    ap = (j == start) ? 0.0f : kNeg;
                    Sa_prev = 0.0;
                    yprev = start;
                }
                if (!act) e = 0.0f;
                if (!act || j >= L) ap = -CUDART_INF_F;
                // unary marginal: exp(ahat_{i+1}[j] + b_{i+1}[j] + (Sa_{i+1} + Sb_{i+1} - logZ))
                if (act && j < L) {
                    const float off = (float)(Sa_next + Sb - lz);
                    const float pj = expf(anext + beta + off);
                    d_emis[(rowbase + t) * L + j] = wb * (pj - ((j == ycur) ? 1.0f : 0.0f));
                }
                const float uj = e + beta + rmax;                                   // -inf for j >= L
                const float eu = ex2_fast(uj * 1.4426950408889634f);
                const float av = ex2_fast(ap * 1.4426950408889634f);
                if (j < K) sts_v4(tile_s + par * kRow + j * 16, uj, eu, av, 0.0f);
                const float cq = act ? wb * expf(fminf((float)(Sa_prev + Sb - lz), 80.0f)) * eu : 0.0f;
                __syncwarp();
                float uk[K], s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float4 x = lds_v4(tile_s + par * kRow + k * 16);
                    uk[k] = x.x;
                    s4[k & 3] = fmaf(Ecol[k], x.y, s4[k & 3]);
                    acc[k] = fmaf(cq, x.z, acc[k]);
                }
                const float M = max_tree<K>(uk);
                const float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                if (act) {
                    beta = (j < L) ? __logf(s) - M : -CUDART_INF_F;
                    Sb += (double)M;
                    if (j == 0) s_gold[ycur * L + yprev] -= wb;                   // gold transition count
                    anext = ap;      // for j >= L this stays -inf; (ap was sanitised only when !act)
                    Sa_next = Sa_prev;
                    ycur = yprev;
                }
                par ^= 1;
            }
            __syncwarp();
        }
        cp_async_wait<0>();
    }
    // dT[j][k] += E[j][k] * acc[k]
    __syncwarp();
    if (j < L) {
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (k < L && acc[k] != 0.0f) atomicAdd(&s_dt[j * L + k], expf(trans[j * L + k] - rmax) * acc[k]);
    }
    __syncwarp();
    for (int i = lane; i < L * L; i += 32) {
        float v = s_dt[i];
#pragma unroll
        for (int q = 0; q < SPW; ++q) v += s_dt[(1 + q) * L * L + 16 + i];
        if (v != 0.0f) atomicAdd(&d_trans[i], v);
    }
}

}  // namespace kbner

using namespace kbner;

extern "C" int kbner_crf_compact(const uint8_t *keep, int B, int T, int32_t *pos, int32_t *klen,
                                 void *stream) {
    KBNER_NVTX("kbner/crf");
    KBNER_CHECK_ARG(keep && pos && klen && B >= 0 && T > 0, "crf_compact: bad arguments");
    if (B == 0) return KBNER_OK;
    const int threads = 128;
    const int blocks = (B * 32 + threads - 1) / threads;
    crf_compact_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(keep, B, T, pos, klen);
    KBNER_CHECK_LAUNCH("crf_compact");
    return KBNER_OK;
}

extern "C" int kbner_crf_nll_fwd(const float *emis, const int32_t *tags, const int32_t *pos,
                                 const int32_t *klen, const float *trans, int B, int T, int L,
                                 int start_idx, int stop_idx, float *logz, float *gold, float *alpha,
                                 double *alpha_scale, void *stream) {
    KBNER_NVTX("kbner/crf");
    KBNER_CHECK_ARG(emis && tags && klen && trans && logz && gold, "crf_nll_fwd: null pointer");
    KBNER_CHECK_ARG(B >= 0 && T > 0 && L >= 2 && L <= 32, "crf_nll_fwd: need L in [2,32], got %d", L);
    KBNER_CHECK_ARG(start_idx >= 0 && start_idx < L && stop_idx >= 0 && stop_idx < L,
                    "crf_nll_fwd: start/stop index out of range");
    KBNER_CHECK_ARG((alpha == nullptr) == (alpha_scale == nullptr), "crf_nll_fwd: alpha and alpha_scale go together");
    if (B == 0) return KBNER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int vec16 = (pos == nullptr && ((size_t)T * L) % 4 == 0 && (reinterpret_cast<uintptr_t>(emis) & 15) == 0) ? 1 : 0;
#define KBNER_NLL(Q, K)                                                                                             \
    crf_nll_fwd_kernel<Q, K><<<(B + NllCfg<Q, K>::SPW - 1) / NllCfg<Q, K>::SPW, 32, NllCfg<Q, K>::smem_bytes(L), st>>>( \
        emis, tags, pos, klen, trans, B, T, L, start_idx, stop_idx, vec16, logz, gold, alpha, alpha_scale)
    if (L == 13) KBNER_NLL(16, 13);
    else if (L <= 16) KBNER_NLL(16, 16);
    else if (L == 29) KBNER_NLL(32, 29);
    else KBNER_NLL(32, 32);
#undef KBNER_NLL
    KBNER_CHECK_LAUNCH("crf_nll_fwd");
    return KBNER_OK;
}

extern "C" int kbner_crf_nll_bwd(const float *emis, const int32_t *tags, const int32_t *pos,
                                 const int32_t *klen, const float *trans, const float *alpha,
                                 const double *alpha_scale, const float *w, int B, int T, int L,
                                 int start_idx, int stop_idx, float *d_emis, float *d_trans,
                                 void *stream) {
    KBNER_NVTX("kbner/crf");
    KBNER_CHECK_ARG(emis && tags && klen && trans && alpha && alpha_scale && w && d_emis && d_trans,
                    "crf_nll_bwd: null pointer");
    KBNER_CHECK_ARG(B >= 0 && T > 0 && L >= 2 && L <= 32, "crf_nll_bwd: need L in [2,32], got %d", L);
    if (B == 0) return KBNER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_emis, 0, sizeof(float) * (size_t)B * T * L, st);
    if (e != cudaSuccess) {
        set_error("crf_nll_bwd: memset: %s", cudaGetErrorString(e));
        return KBNER_ECUDA;
    }
#define KBNER_NLLB(Q, K)                                                                                      \
    do {                                                                                                      \
        using C = NllBwdCfg<Q, K>;                                                                            \
        int blocks = (B + C::SPW - 1) / C::SPW;                                                               \
        /* One block per sentence group while a single wave holds them all; persistent over sentences beyond that.   \
           (The first version capped the grid at 8 blocks per SM: 4096 sentences = 2048 groups then ran as 1184 blocks \
           of which 864 walked two groups back to back -- twice the chain -- with 6.7 warps per SM resident: 528 us,   \
           issue slots 32 % busy, profiles/r02/crf_ncu_p1.txt.) */                                                  \
        static std::atomic<int> per_sm{0};                                                                    \
        if (per_sm == 0) {                                                                                    \
            cudaFuncSetAttribute(crf_nll_bwd_kernel<Q, K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);   \
            int n_ = 0;                                                                                       \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n_, crf_nll_bwd_kernel<Q, K>, 32, C::smem_bytes(L)) != cudaSuccess || n_ < 1) n_ = 8; \
            per_sm = n_;                                                                                      \
        }                                                                                                     \
        if (blocks > per_sm * num_sms()) blocks = per_sm * num_sms();                                         \
        crf_nll_bwd_kernel<Q, K><<<blocks, 32, C::smem_bytes(L), st>>>(emis, tags, pos, klen, trans, alpha,    \
                                                                      alpha_scale, w, B, T, L, start_idx,     \
                                                                      stop_idx, d_emis, d_trans);             \
    } while (0)
    if (L == 13) KBNER_NLLB(16, 13);
    else if (L <= 16) KBNER_NLLB(16, 16);
    else if (L == 29) KBNER_NLLB(32, 29);
    else KBNER_NLLB(32, 32);
#undef KBNER_NLLB
    KBNER_CHECK_LAUNCH("crf_nll_bwd");
    return KBNER_OK;
}
