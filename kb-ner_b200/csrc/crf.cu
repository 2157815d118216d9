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

#include "common.cuh"

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
// x = (e[j] + A[j][k]) + alpha[k]  (L^2 exps per step).  Here the same quantity is carried in a
// normalised form  alpha_t[j] = S_t + ahat_t[j],  max_j ahat_t[j] = 0,  S_t accumulated in fp64:
//   r[j]        = e[j] + rmax_j + log sum_k E[j][k] * exp(ahat[k]),   E[j][k] = exp(A[j][k] - rmax_j)
//   ahat'[j]    = r[j] - max_j r[j],      S' = S + max_j r[j]
// one exp per lane per step and an FMA per pair.  Keeping the O(T) magnitude in a separate fp64
// scalar is what keeps the marginals of the backward pass accurate at T = 512 (alpha ~ 2000 would
// otherwise carry ~1e-4 absolute error per step into exp(alpha + beta - logZ)).
// ------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(128)
crf_nll_fwd_kernel(const float *__restrict__ emis, const int32_t *__restrict__ tags,
                   const int32_t *__restrict__ pos, const int32_t *__restrict__ klen,
                   const float *__restrict__ trans, int B, int T, int L, int start, int stop,
                   float *__restrict__ logz, float *__restrict__ gold, float *__restrict__ alpha_out,
                   double *__restrict__ ascale_out) {
    constexpr int SPW = 32 / G;
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / G, j = lane % G;
    const int b = (blockIdx.x * W + warp) * SPW + sub;
    const bool valid = b < B;
    const int n = valid ? klen[b] : 0;
    const int nmax = warp_max_int(n);
    const size_t rowbase = (size_t)(valid ? b : 0) * T;

    float rmax = -CUDART_INF_F;
    if (j < L)
        for (int k = 0; k < L; ++k) rmax = fmaxf(rmax, trans[j * L + k]);
    float E[G];
#pragma unroll
    for (int k = 0; k < G; ++k) E[k] = (j < L && k < L) ? expf(trans[j * L + k] - rmax) : 0.0f;
    if (j >= L) rmax = 0.0f;

    float a = (j < L) ? ((j == start) ? 0.0f : kNeg) : -CUDART_INF_F;   // ahat_0 (max = 0)
    double S = 0.0;
    constexpr int U = 4;
    float e_cur[U], e_nxt[U];
    auto load_block = [&](int i0, float (&dst)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u;
            float ev = 0.0f;
            if (i < n && j < L) {
                const int t = pos ? __ldg(pos + rowbase + i) : i;
                ev = __ldg(emis + (rowbase + t) * L + j);
            }
            dst[u] = ev;
        }
    };
    load_block(0, e_cur);
    for (int i0 = 0; i0 < nmax; i0 += U) {
        load_block(i0 + U, e_nxt);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u;
            if (i < nmax) {
                const float p = expf(a);
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < G; ++k) s = fmaf(E[k], __shfl_sync(0xffffffffu, p, k, G), s);
                const float r = (j < L) ? (e_cur[u] + rmax) + logf(s) : -CUDART_INF_F;
                const float m = group_max<G>(r);
                if (i < n) {
                    a = r - m;
                    S += (double)m;
                    if (alpha_out && j < L) alpha_out[(rowbase + i) * L + j] = a;
                    if (ascale_out && j == 0) ascale_out[rowbase + i] = S;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) e_cur[u] = e_nxt[u];
    }
    // terminal: log_sum_exp_batch(alpha_len + A[STOP])  (:1381-1392)
    const float x = (j < L) ? a + trans[stop * L + j] : -CUDART_INF_F;
    const float M2 = group_max<G>(x);
    const float s2 = group_sum<G>(expf(x - M2));
    // gold score (:2544-2591): lanes stride over the kept tokens
    float g = 0.0f;
    for (int i = j; i < n; i += G) {
        const int t = pos ? pos[rowbase + i] : i;
        const int y = tags[rowbase + t];
        int prev = start;
        if (i > 0) {
            const int tp = pos ? pos[rowbase + i - 1] : i - 1;
            prev = tags[rowbase + tp];
        }
        g += emis[(rowbase + t) * L + y] + trans[y * L + prev];
    }
    g = group_sum<G>(g);
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
    const int W = 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (L <= 16) {
        const int per_block = W * 2;
        crf_nll_fwd_kernel<16><<<(B + per_block - 1) / per_block, W * 32, 0, st>>>(
            emis, tags, pos, klen, trans, B, T, L, start_idx, stop_idx, logz, gold, alpha, alpha_scale);
    } else {
        crf_nll_fwd_kernel<32><<<(B + W - 1) / W, W * 32, 0, st>>>(
            emis, tags, pos, klen, trans, B, T, L, start_idx, stop_idx, logz, gold, alpha, alpha_scale);
    }
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
