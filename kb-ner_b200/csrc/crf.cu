// Linear-chain CRF on sm_100a: remove-X compaction, log-partition / gold score, and their
// gradient (Viterbi: crf_viterbi.cu).  Replaces the per-token Python loops of
// /root/reference/flair/models/sequence_tagger_model.py (_viterbi_decode :1248-1304,
// _forward_alg :1329-1394, _score_sentence :2544-2591, _calculate_loss :2448-2506,
// _obtain_labels :1193-1210).
//
// Mapping: one group of G lanes per sentence (G = 16 when L <= 16, two sentences per warp;
// else G = 32), lane j owns tag j.  The recurrence state lives in registers, the transition
// row of the lane in registers, the all-to-all exchange of the previous state is G width-G
// shuffles.  These kernels are bound by the dependent chain / issue rate, their HBM traffic
// is the emissions read once ("CRF kernels").
#include <math_constants.h>

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
// gradient of sum_b w[b] (logZ_b - gold_b).  Reverse (beta) recursion in the same normalised form
// (beta_t = Sb_t + bhat_t, Sb in fp64); unary marginals -> d_emis, pairwise marginals accumulated
// per lane as  acc[j][k] += c * q_j * a_k   (dT[j][k] = E[j][k] * acc[j][k]), persistent over
// sentences, reduced through shared memory, one global atomic set per block.
// ------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(128)
crf_nll_bwd_kernel(const float *__restrict__ emis, const int32_t *__restrict__ tags,
                   const int32_t *__restrict__ pos, const int32_t *__restrict__ klen,
                   const float *__restrict__ trans, const float *__restrict__ alpha,
                   const double *__restrict__ ascale, const float *__restrict__ w, int B, int T, int L,
                   int start, int stop, float *__restrict__ d_emis, float *__restrict__ d_trans) {
    __shared__ float s_dt[32 * 32];
    constexpr int SPW = 32 / G;
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / G, j = lane % G;
    for (int i = threadIdx.x; i < L * L; i += blockDim.x) s_dt[i] = 0.0f;
    __syncthreads();

    float rmax = -CUDART_INF_F;
    if (j < L)
        for (int k = 0; k < L; ++k) rmax = fmaxf(rmax, trans[j * L + k]);
    // Ecol[jj] = E[jj][j]: column j of the row-normalised exp(A)
    float Ecol[G];
#pragma unroll
    for (int jj = 0; jj < G; ++jj) {
        float v = 0.0f;
        if (j < L && jj < L) {
            float rm = -CUDART_INF_F;
            for (int k = 0; k < L; ++k) rm = fmaxf(rm, trans[jj * L + k]);
            v = expf(trans[jj * L + j] - rm);
        }
        Ecol[jj] = v;
    }
    if (j >= L) rmax = 0.0f;
    float acc[G];
#pragma unroll
    for (int k = 0; k < G; ++k) acc[k] = 0.0f;

    const int groups_total = gridDim.x * W * SPW;
    const int g0 = (blockIdx.x * W + warp) * SPW + sub;
    const int iters = (B + groups_total - 1) / groups_total;
    for (int it = 0; it < iters; ++it) {
        const int b = g0 + it * groups_total;
        const bool valid = b < B;
        const int n = valid ? klen[b] : 0;
        const int nmax = warp_max_int(n);
        const size_t rowbase = (size_t)(valid ? b : 0) * T;
        const float wb = valid ? w[b] : 0.0f;
        if (nmax == 0) continue;   // warp-uniform

        // beta_n[k] = A[STOP][k], normalised; logZ re-derived in fp64 from the stored alpha
        const float astop = (j < L) ? trans[stop * L + j] : -CUDART_INF_F;
        const float mb0 = group_max<G>(astop);
        float beta = astop - mb0;                                   // bhat_n
        double Sb = (double)mb0;
        const float a_n = (n > 0 && j < L) ? alpha[(rowbase + n - 1) * L + j] : -CUDART_INF_F;
        const double Sa_n = (n > 0) ? ascale[rowbase + n - 1] : 0.0;
        const float xt = (j < L) ? a_n + astop : -CUDART_INF_F;
        const float Mt = group_max<G>(xt);
        const float st = group_sum<G>(expf(xt - Mt));
        const double lz = Sa_n + (double)Mt + (double)logf(st);
        // terminal pairwise term: dT[STOP][k] += w * exp(alpha_n[k] + A[STOP][k] - logZ)
        if (n > 0 && j < L) atomicAdd(&s_dt[stop * L + j], wb * expf(xt - Mt) / st);
        for (int i = nmax - 1; i >= 0; --i) {
            const bool act = i < n;
            float e = 0.0f, anext = -CUDART_INF_F, aprev = -CUDART_INF_F;
            double Sa_next = 0.0, Sa_prev = 0.0;
            int t = 0, y = -1, yprev = start;
            if (act) {
                t = pos ? pos[rowbase + i] : i;
                y = tags[rowbase + t];
                Sa_next = ascale[rowbase + i];
                if (i > 0) {
                    const int tp = pos ? pos[rowbase + i - 1] : i - 1;
                    yprev = tags[rowbase + tp];
                    Sa_prev = ascale[rowbase + i - 1];
                }
                if (j < L) {
                    e = emis[(rowbase + t) * L + j];
                    anext = alpha[(rowbase + i) * L + j];
                    aprev = (i > 0) ? alpha[(rowbase + i - 1) * L + j] : ((j == start) ? 0.0f : kNeg);
                }
            }
            // unary marginal: exp(ahat_{i+1}[j] + bhat_{i+1}[j] + (Sa_{i+1} + Sb_{i+1} - logZ))
            if (act && j < L) {
                const float off = (float)(Sa_next + Sb - lz);
                const float pj = expf(anext + beta + off);
                d_emis[(rowbase + t) * L + j] = wb * (pj - ((j == y) ? 1.0f : 0.0f));
            }
            const float u = (j < L && act) ? (e + beta + rmax) : -CUDART_INF_F;
            const float Mb = group_max<G>(u);
            const float q = (act && j < L) ? expf(u - Mb) : 0.0f;
            const float av = (act && j < L) ? expf(aprev) : 0.0f;      // ahat is normalised: max = 0
            const float c = act ? wb * expf(fminf((float)(Sa_prev + Sb + (double)Mb - lz), 80.0f)) : 0.0f;
            const float cq = c * q;
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < G; ++k) {
                acc[k] = fmaf(cq, __shfl_sync(0xffffffffu, av, k, G), acc[k]);
                s = fmaf(Ecol[k], __shfl_sync(0xffffffffu, q, k, G), s);
            }
            const float tk = (j < L) ? logf(s) : -CUDART_INF_F;
            const float mt = group_max<G>(tk);
            if (act) {
                beta = tk - mt;
                Sb += (double)Mb + (double)mt;
            }
            // gold transition count
            if (act && j == 0) atomicAdd(&s_dt[y * L + yprev], -wb);
        }
        if (n > 0 && j == 0) {
            const int tl = pos ? pos[rowbase + n - 1] : n - 1;
            atomicAdd(&s_dt[stop * L + tags[rowbase + tl]], -wb);
        }
    }
    // dT[j][k] += E[j][k] * acc[k]
    if (j < L) {
#pragma unroll
        for (int k = 0; k < G; ++k)
            if (k < L && acc[k] != 0.0f) atomicAdd(&s_dt[j * L + k], expf(trans[j * L + k] - rmax) * acc[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L * L; i += blockDim.x)
        if (s_dt[i] != 0.0f) atomicAdd(&d_trans[i], s_dt[i]);
}

}  // namespace kbner

using namespace kbner;

extern "C" int kbner_crf_compact(const uint8_t *keep, int B, int T, int32_t *pos, int32_t *klen,
                                 void *stream) {
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
    const int W = 4;
    const int spw = (L <= 16) ? 2 : 1;
    int blocks = (B + W * spw - 1) / (W * spw);
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;   // persistent over sentences beyond that
    if (L <= 16)
        crf_nll_bwd_kernel<16><<<blocks, W * 32, 0, st>>>(emis, tags, pos, klen, trans, alpha, alpha_scale, w,
                                                          B, T, L, start_idx, stop_idx, d_emis, d_trans);
    else
        crf_nll_bwd_kernel<32><<<blocks, W * 32, 0, st>>>(emis, tags, pos, klen, trans, alpha, alpha_scale, w,
                                                          B, T, L, start_idx, stop_idx, d_emis, d_trans);
    KBNER_CHECK_LAUNCH("crf_nll_bwd");
    return KBNER_OK;
}
