// HBM-bound kernels of the encoder / tagger head: embedding gather + LayerNorm, LayerNorm,
// first-sub-token gather + word dropout + tag projection.  One warp per row, 16-byte loads,
// fp32 statistics (two-pass, from registers), bf16 activations out.
#include <atomic>

#include "common.cuh"

namespace kbner {

constexpr int kLnWarps = 4;

// ------------------------------------------------------------------------------------------
// LayerNorm over fp32 rows (the GEMM epilogue already produced x = A.B^T + bias + residual in
// fp32), bf16 out.  HF BertSelfOutput / BertOutput LayerNorm as called from
// /root/reference/flair/embeddings.py:3269.  VPL = float4 per lane.
// ------------------------------------------------------------------------------------------
template <int VPL, bool FUSED>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_kernel(const float *__restrict__ x, const float *__restrict__ bias, const uint16_t *__restrict__ resid,
                     const float *__restrict__ gamma, const float *__restrict__ beta, float eps, int M,
                     uint16_t *__restrict__ y, float *__restrict__ mean_out, float *__restrict__ rstd_out,
                     const Dropout drop) {
    constexpr int H = VPL * 128;
    const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float4 *xr = reinterpret_cast<const float4 *>(x + (size_t)row * H);
    float4 v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const uint4 u = ld_nc_v4(xr + i * 32 + lane);
        v[i] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
    }
    if (FUSED) {
        // z = dropout(x + bias) + resid : the Linear's bias, the hidden-state dropout and the residual connection of
        // BertSelfOutput / BertOutput, taken out of the GEMM epilogue (where the row-per-thread accumulator layout made
        // the residual an uncoalesced 32-sector-per-request load on the critical path; profiles/r01/gemm_attnout_ncu.txt)
        const uint32_t key = drop.thresh ? drop_key(drop) : 0u;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c4 = i * 32 + lane;           // float4 index inside the row
            if (bias) {
                const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
                v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
            }
            if (drop.thresh) {
                const uint32_t pair = (uint32_t)row * (H / 2) + (uint32_t)c4 * 2u;
                const uint32_t b0 = drop_bits(key, pair), b1 = drop_bits(key, pair + 1u);
                v[i].x = drop_keep_lo(b0, drop.thresh) ? v[i].x * drop.scale : 0.0f;
                v[i].y = drop_keep_hi(b0, drop.thresh) ? v[i].y * drop.scale : 0.0f;
                v[i].z = drop_keep_lo(b1, drop.thresh) ? v[i].z * drop.scale : 0.0f;
                v[i].w = drop_keep_hi(b1, drop.thresh) ? v[i].w * drop.scale : 0.0f;
            }
            if (resid) {
                const uint2 r = __ldg(reinterpret_cast<const uint2 *>(resid + (size_t)row * H) + c4);
                float r0, r1, r2, r3;
                unpack_bf16x2(r.x, r0, r1);
                unpack_bf16x2(r.y, r2, r3);
                v[i].x += r0; v[i].y += r1; v[i].z += r2; v[i].w += r3;
            }
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(sum) * (1.0f / H);
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / H) + eps);
    uint16_t *yr = y + (size_t)row * H;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma) + i * 32 + lane);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(beta) + i * 32 + lane);
        uint2 o;
        o.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
        o.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        *reinterpret_cast<uint2 *>(yr + (i * 32 + lane) * 4) = o;
    }
    if (mean_out && lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}


// ------------------------------------------------------------------------------------------
// Precision modes of the inference forward (kb-ner_b200/encoder.py, `precision`):
//   "bf16-res32": the residual stream stays fp32 -- LayerNorm reads an fp32 residual and writes its output twice, fp32
//                 for the next residual add and bf16 as the next GEMM operand;
//   "bf16x3"    : additionally every GEMM operand is carried as a bf16 PAIR hi + lo (lo = bf16(v - hi), |v - hi - lo| <=
//                 2^-17 |v|) and a product x.W is evaluated as x_hi.W_hi + x_lo.W_hi + x_hi.W_lo by ONE tensor-core
//                 GEMM over the K-concatenated operands [x_hi | x_lo | x_hi] . [W_hi | W_hi | W_lo]^T with fp32
//                 accumulation in TMEM: the activation row is written as [ hi | lo | hi ], 3H wide.
// Both exist to meet BASELINE.json's "logits within 1e-3 relative" against the reference's fp32 arithmetic
// (flair/embeddings.py:3269): bf16 weights alone put the 24-layer hidden state 6.4e-3 away (scripts/bf16_ablation.py).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_store4(uint16_t *row, int c4, int H, bool split, const float4 o) {
    uint2 hi;
    hi.x = pack_bf16x2(o.x, o.y);
    hi.y = pack_bf16x2(o.z, o.w);
    *reinterpret_cast<uint2 *>(row + c4 * 4) = hi;
    if (split) {
        float h0, h1, h2, h3;
        unpack_bf16x2(hi.x, h0, h1);
        unpack_bf16x2(hi.y, h2, h3);
        uint2 lo;
        lo.x = pack_bf16x2(o.x - h0, o.y - h1);
        lo.y = pack_bf16x2(o.z - h2, o.w - h3);
        *reinterpret_cast<uint2 *>(row + H + c4 * 4) = lo;
        *reinterpret_cast<uint2 *>(row + 2 * H + c4 * 4) = hi;
    }
}

template <int VPL>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_res32_kernel(const float *__restrict__ x, const float *__restrict__ bias, const float *__restrict__ resid,
                           const float *__restrict__ gamma, const float *__restrict__ beta, float eps, int M,
                           float *__restrict__ y32, uint16_t *__restrict__ y, int split) {
    constexpr int H = VPL * 128;
    const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float4 *xr = reinterpret_cast<const float4 *>(x + (size_t)row * H);
    float4 v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const uint4 u = ld_nc_v4(xr + i * 32 + lane);
        v[i] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
        const int c4 = i * 32 + lane;
        if (bias) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
            v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
        }
        if (resid) {
            const uint4 r = ld_nc_v4(reinterpret_cast<const float4 *>(resid + (size_t)row * H) + c4);
            v[i].x += __uint_as_float(r.x); v[i].y += __uint_as_float(r.y);
            v[i].z += __uint_as_float(r.z); v[i].w += __uint_as_float(r.w);
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(sum) * (1.0f / H);
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
    // 1/sqrt in full precision: this path is the one held to 1e-3 against fp32 arithmetic
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / H) + eps);
    uint16_t *yr = y + (size_t)row * (split ? 3 * H : H);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c4 = i * 32 + lane;
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(beta) + c4);
        const float4 o = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                     (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        if (y32) *reinterpret_cast<float4 *>(y32 + (size_t)row * H + c4 * 4) = o;
        split_store4(yr, c4, H, split != 0, o);
    }
}

// h = gelu_erf(x + bias) written as [ hi | lo | hi ] (FFN-up output of the "bf16x3" mode; erff, not the epilogue's
// polynomial: this mode is the one held to fp32 arithmetic).  x fp32 [M,F]; out bf16 [M,3F].
__global__ void __launch_bounds__(256)
bias_gelu_split_kernel(const float *__restrict__ x, const float *__restrict__ bias, size_t M, int F, uint16_t *__restrict__ out) {
    const int f4 = F / 4;
    const size_t total = M * (size_t)f4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / f4;
        const int c4 = (int)(i - r * f4);
        const uint4 u = ld_nc_v4(reinterpret_cast<const float4 *>(x) + i);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
        float4 o;
        float t;
        t = __uint_as_float(u.x) + b.x; o.x = 0.5f * t * (1.0f + erff(t * 0.70710678118654752f));
        t = __uint_as_float(u.y) + b.y; o.y = 0.5f * t * (1.0f + erff(t * 0.70710678118654752f));
        t = __uint_as_float(u.z) + b.z; o.z = 0.5f * t * (1.0f + erff(t * 0.70710678118654752f));
        t = __uint_as_float(u.w) + b.w; o.w = 0.5f * t * (1.0f + erff(t * 0.70710678118654752f));
        split_store4(out + r * (size_t)(3 * F), c4, F, true, o);
    }
}

// ------------------------------------------------------------------------------------------
// Embedding gather + LayerNorm.  HF (XLM-)RobertaEmbeddings: word[ids] + type[0] + pos[p],
// p = cumsum(ids != pad) * (ids != pad) + pad  (padding_idx = 1), LayerNorm(eps), SURVEY E1.
// Block = 4 warps = 16 consecutive sub-tokens of one window; the non-pad prefix count up to the
// chunk is a block-wide count, inside the chunk a ballot.
// ------------------------------------------------------------------------------------------
constexpr int kEmbTok = 16;

template <int VPL>
__global__ void __launch_bounds__(128)
embed_ln_fwd_kernel(const int32_t *__restrict__ ids, const float *__restrict__ word_emb,
                    const float *__restrict__ pos_emb, const float *__restrict__ type_emb,
                    const float *__restrict__ gamma, const float *__restrict__ beta, float eps, int pad_id, int S,
                    int V, int P, uint16_t *__restrict__ out, float *__restrict__ out32, int split) {
    constexpr int H = VPL * 128;
    const int r = blockIdx.y;
    const int s0 = blockIdx.x * kEmbTok;
    const int32_t *idr = ids + (size_t)r * S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // non-pad tokens strictly before the chunk
    int before = 0;
    for (int base = 0; base < s0; base += 128) {
        const int t = base + threadIdx.x;
        before += __syncthreads_count(t < s0 && idr[t] != pad_id);
    }
    const int tl = s0 + (lane & (kEmbTok - 1));
    const int my_id = (tl < S) ? idr[tl] : pad_id;
    const unsigned nonpad = __ballot_sync(0xffffffffu, my_id != pad_id) & 0xffffu;
#pragma unroll
    for (int q = 0; q < kEmbTok / 4; ++q) {
        const int local = warp * (kEmbTok / 4) + q;
        const int sidx = s0 + local;
        if (sidx >= S) break;
        const int id = __shfl_sync(0xffffffffu, my_id, local);
        int p = pad_id;
        if (id != pad_id) p = before + __popc(nonpad & ((2u << local) - 1u)) + pad_id;
        if (id < 0 || id >= V || p >= P) {      // corrupt input: fail loudly instead of reading out of bounds
            if (lane == 0) printf("kbner embed_ln: id %d / position %d out of range (V=%d P=%d)\n", id, p, V, P);
            __trap();
        }
        const float4 *wr = reinterpret_cast<const float4 *>(word_emb + (size_t)id * H);
        const float4 *pr = reinterpret_cast<const float4 *>(pos_emb + (size_t)p * H);
        const float4 *tr = reinterpret_cast<const float4 *>(type_emb);
        float4 v[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 a = __ldg(wr + i * 32 + lane), b = __ldg(tr + i * 32 + lane), c = __ldg(pr + i * 32 + lane);
            v[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
        }
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float mean = warp_sum(sum) * (1.0f / H);
        float sq = 0.0f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            sq += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(warp_sum(sq) * (1.0f / H) + eps);
        uint16_t *yr = out + ((size_t)r * S + sidx) * (split ? 3 * H : H);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma) + i * 32 + lane);
            const float4 b = __ldg(reinterpret_cast<const float4 *>(beta) + i * 32 + lane);
            const float4 o = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                         (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
            if (out32) *reinterpret_cast<float4 *>(out32 + ((size_t)r * S + sidx) * H + (i * 32 + lane) * 4) = o;   // precision modes
            split_store4(yr, i * 32 + lane, H, split != 0, o);
        }
    }
}

// ------------------------------------------------------------------------------------------
// First-sub-token gather + word dropout + Linear(H -> L), fp32 logits.
// Replaces /root/reference/flair/embeddings.py:3288-3345 + :108-124 (pooling, .cpu() round trip),
// flair/nn.py:176-183 (WordDropout) and sequence_tagger_model.py:1027 (self.linear).
// W (fp32, [L,H]) is staged once per block in shared memory; one warp per word.
// ------------------------------------------------------------------------------------------
// Two words per warp iteration share every W read from shared memory (the first version read the whole 53 KB of W per
// WORD: 868 MB of shared-memory traffic at 16320 words, 74 us), and the 2 x L partial sums are reduced through a
// conflict-free [2L][33] shared tile by 2L lanes instead of 2L five-step shuffle reductions.
constexpr int kTagprojWarps = 8;
template <int CPL, bool F32>   // 8-element chunks per lane: H = CPL * 256; F32: the hidden state is fp32 (precision modes)
__global__ void __launch_bounds__(kTagprojWarps * 32, 2)
gather_tagproj_fwd_kernel(const void *__restrict__ hidden_v, const int32_t *__restrict__ row_of,
                          const int32_t *__restrict__ first_idx, const uint8_t *__restrict__ drop_keep,
                          const float *__restrict__ W, const float *__restrict__ bias, int B, int T, int S, int L,
                          float *__restrict__ logits) {
    constexpr int H = CPL * 256;
    const uint16_t *hidden = reinterpret_cast<const uint16_t *>(hidden_v);
    extern __shared__ __align__(16) float w_s[];     // [L][H], then per warp a [2L][33] reduction tile
    for (int i = threadIdx.x; i < L * H / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(w_s)[i] = __ldg(reinterpret_cast<const float4 *>(W) + i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *red = w_s + (size_t)L * H + (size_t)warp * 2 * L * 33;
    const int words = B * T;
    const int pairs = (words + 1) / 2;
    const int stride = gridDim.x * kTagprojWarps;
    for (int p = blockIdx.x * kTagprojWarps + warp; p < pairs; p += stride) {
        float x[2][CPL * 8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int w = 2 * p + h;
            bool live = false;
            size_t hoff = 0;
            if (w < words) {
                const int b = w / T, t = w - b * T;
                const int fi = first_idx[w];
                live = fi >= 0 && (!drop_keep || drop_keep[t] != 0);
                if (live) hoff = ((size_t)row_of[b] * S + fi) * H;
            }
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                if (F32) {
                    uint4 u0 = make_uint4(0u, 0u, 0u, 0u), u1 = u0;
                    if (live) {
                        const float *hr = reinterpret_cast<const float *>(hidden_v) + hoff + c * 256 + lane * 8;
                        u0 = ld_nc_v4(hr);
                        u1 = ld_nc_v4(hr + 4);
                    }
                    x[h][c * 8 + 0] = __uint_as_float(u0.x); x[h][c * 8 + 1] = __uint_as_float(u0.y);
                    x[h][c * 8 + 2] = __uint_as_float(u0.z); x[h][c * 8 + 3] = __uint_as_float(u0.w);
                    x[h][c * 8 + 4] = __uint_as_float(u1.x); x[h][c * 8 + 5] = __uint_as_float(u1.y);
                    x[h][c * 8 + 6] = __uint_as_float(u1.z); x[h][c * 8 + 7] = __uint_as_float(u1.w);
                } else {
                    uint4 u = make_uint4(0u, 0u, 0u, 0u);      // zero vector (0 sub-tokens / dropped word) => bias only
                    if (live) u = ld_nc_v4(hidden + hoff + c * 256 + lane * 8);
                    unpack_bf16x2(u.x, x[h][c * 8 + 0], x[h][c * 8 + 1]);
                    unpack_bf16x2(u.y, x[h][c * 8 + 2], x[h][c * 8 + 3]);
                    unpack_bf16x2(u.z, x[h][c * 8 + 4], x[h][c * 8 + 5]);
                    unpack_bf16x2(u.w, x[h][c * 8 + 6], x[h][c * 8 + 7]);
                }
            }
        }
        for (int l = 0; l < L; ++l) {
            const float *wl = w_s + (size_t)l * H;
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const float4 w0 = *reinterpret_cast<const float4 *>(wl + c * 256 + lane * 8);
                const float4 w1 = *reinterpret_cast<const float4 *>(wl + c * 256 + lane * 8 + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    a0 = fmaf(x[0][c * 8 + i], wv[i], a0);
                    a1 = fmaf(x[1][c * 8 + i], wv[i], a1);
                }
            }
            red[l * 33 + lane] = a0;                     // bank (l + lane) % 32: conflict-free
            red[(L + l) * 33 + lane] = a1;
        }
        __syncwarp();
        for (int v = lane; v < 2 * L; v += 32) {         // lane sums row v: bank (v + i) % 32
            const float *r = red + v * 33;
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) { s0 += r[i]; s1 += r[i + 1]; s2 += r[i + 2]; s3 += r[i + 3]; }
            const int h = (v >= L) ? 1 : 0, l = v - h * L;
            const int w = 2 * p + h;
            if (w < words) logits[(size_t)w * L + l] = __ldg(bias + l) + ((s0 + s1) + (s2 + s3));
        }
        __syncwarp();
    }
}

}  // namespace kbner

using namespace kbner;

#define DISPATCH_VPL(H, CALL)                                                                   \
    switch ((H) / 128) {                                                                        \
        case 2: { constexpr int VPL = 2; CALL; } break;                                         \
        case 4: { constexpr int VPL = 4; CALL; } break;                                         \
        case 6: { constexpr int VPL = 6; CALL; } break;                                         \
        case 8: { constexpr int VPL = 8; CALL; } break;                                         \
        case 16: { constexpr int VPL = 16; CALL; } break;                                       \
        default: set_error("hidden size %d not built (supported: 256, 512, 768, 1024, 2048)", (H)); \
                 return KBNER_EUNSUPPORTED;                                                     \
    }

extern "C" int kbner_add_layernorm_fwd(const float *x, const float *bias, const uint16_t *resid, const float *gamma,
                                       const float *beta, float eps, int M, int H, uint16_t *y, float *mean, float *rstd,
                                       const uint32_t *drop_seed, uint32_t drop_site, float drop_p, void *stream) {
    KBNER_NVTX("kbner/elementwise");
    KBNER_CHECK_ARG(x && gamma && beta && y, "layernorm_fwd: null pointer");
    KBNER_CHECK_ARG(M >= 0 && H > 0 && H % 128 == 0, "layernorm_fwd: H=%d must be a multiple of 128", H);
    KBNER_CHECK_ARG((mean == nullptr) == (rstd == nullptr), "layernorm_fwd: mean and rstd go together");
    KBNER_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f, "layernorm_fwd: dropout probability %f", (double)drop_p);
    KBNER_CHECK_ARG((uint64_t)M * (uint64_t)(H / 2) < (1ull << 32), "layernorm_fwd: M*H/2 exceeds the 32-bit dropout counter");
    if (M == 0) return KBNER_OK;
    const int blocks = (M + kLnWarps - 1) / kLnWarps;
    cudaStream_t st = (cudaStream_t)stream;
    const Dropout drop = make_dropout(drop_seed, drop_site, drop_p);
    if (bias || resid || drop.thresh) {
        DISPATCH_VPL(H, (layernorm_fwd_kernel<VPL, true><<<blocks, kLnWarps * 32, 0, st>>>(x, bias, resid, gamma, beta, eps, M,
                                                                                         y, mean, rstd, drop)));
    } else {
        DISPATCH_VPL(H, (layernorm_fwd_kernel<VPL, false><<<blocks, kLnWarps * 32, 0, st>>>(x, bias, resid, gamma, beta, eps,
                                                                                          M, y, mean, rstd, drop)));
    }
    KBNER_CHECK_LAUNCH("layernorm_fwd");
    return KBNER_OK;
}

extern "C" int kbner_layernorm_fwd(const float *x, const float *gamma, const float *beta, float eps, int M, int H,
                                   uint16_t *y, float *mean, float *rstd, void *stream) {
    KBNER_NVTX("kbner/elementwise");
    return kbner_add_layernorm_fwd(x, nullptr, nullptr, gamma, beta, eps, M, H, y, mean, rstd, nullptr, 0u, 0.0f, stream);
}

extern "C" int kbner_embed_ln_fwd_ex(const int32_t *ids, const float *word_emb, const float *pos_emb,
                                     const float *type_emb, const float *gamma, const float *beta, float eps,
                                     int pad_id, int R, int S, int H, int V, int P, uint16_t *out, float *out32, int split,
                                     void *stream) {
    KBNER_NVTX("kbner/elementwise");
    KBNER_CHECK_ARG(ids && word_emb && pos_emb && type_emb && gamma && beta && out, "embed_ln_fwd: null pointer");
    KBNER_CHECK_ARG(R >= 0 && S > 0 && H % 128 == 0 && V > 0 && P > 0, "embed_ln_fwd: bad shape");
    if (R == 0) return KBNER_OK;
    dim3 grid((S + kEmbTok - 1) / kEmbTok, R);
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_VPL(H, (embed_ln_fwd_kernel<VPL><<<grid, 128, 0, st>>>(ids, word_emb, pos_emb, type_emb, gamma, beta, eps,
                                                                     pad_id, S, V, P, out, out32, split)));
    KBNER_CHECK_LAUNCH("embed_ln_fwd");
    return KBNER_OK;
}

extern "C" int kbner_embed_ln_fwd(const int32_t *ids, const float *word_emb, const float *pos_emb,
                                  const float *type_emb, const float *gamma, const float *beta, float eps,
                                  int pad_id, int R, int S, int H, int V, int P, uint16_t *out, void *stream) {
    KBNER_NVTX("kbner/elementwise");
    return kbner_embed_ln_fwd_ex(ids, word_emb, pos_emb, type_emb, gamma, beta, eps, pad_id, R, S, H, V, P, out, nullptr, 0,
                                 stream);
}

extern "C" int kbner_add_layernorm_fwd_res32(const float *x, const float *bias, const float *resid, const float *gamma,
                                             const float *beta, float eps, int M, int H, float *y32, uint16_t *y,
                                             int split, void *stream) {
    KBNER_NVTX("kbner/elementwise");
    KBNER_CHECK_ARG(x && gamma && beta && y, "layernorm_fwd_res32: null pointer");
    KBNER_CHECK_ARG(M >= 0 && H > 0 && H % 128 == 0, "layernorm_fwd_res32: H=%d must be a multiple of 128", H);
    if (M == 0) return KBNER_OK;
    const int blocks = (M + kLnWarps - 1) / kLnWarps;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_VPL(H, (layernorm_fwd_res32_kernel<VPL><<<blocks, kLnWarps * 32, 0, st>>>(x, bias, resid, gamma, beta, eps, M, y32,
                                                                                       y, split)));
    KBNER_CHECK_LAUNCH("layernorm_fwd_res32");
    return KBNER_OK;
}

extern "C" int kbner_bias_gelu_split(const float *x, const float *bias, int M, int F, uint16_t *out3, void *stream) {
    KBNER_NVTX("kbner/elementwise");
    KBNER_CHECK_ARG(x && bias && out3, "bias_gelu_split: null pointer");
    KBNER_CHECK_ARG(M >= 0 && F > 0 && F % 4 == 0, "bias_gelu_split: F=%d must be a multiple of 4", F);
    if (M == 0) return KBNER_OK;
    const size_t total = (size_t)M * (F / 4);
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
    bias_gelu_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, bias, (size_t)M, F, out3);
    KBNER_CHECK_LAUNCH("bias_gelu_split");
    return KBNER_OK;
}

template <int CPL, bool F32>
static int launch_tagproj(const void *hidden, const int32_t *row_of, const int32_t *first_idx,
                          const uint8_t *drop_keep, const float *W, const float *bias, int B, int T, int S, int L,
                          float *logits, cudaStream_t st) {
    const size_t smem = ((size_t)L * CPL * 256 + (size_t)kTagprojWarps * 2 * L * 33) * sizeof(float);
    static std::atomic<size_t> configured{0};     // idempotent set-up: a race only repeats it
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(gather_tagproj_fwd_kernel<CPL, F32>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("gather_tagproj: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
            return KBNER_ECUDA;
        }
        configured = smem;
    }
    const int pairs = (B * T + 1) / 2;
    int blocks = (pairs + kTagprojWarps - 1) / kTagprojWarps;
    if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
    gather_tagproj_fwd_kernel<CPL, F32><<<blocks, kTagprojWarps * 32, smem, st>>>(hidden, row_of, first_idx, drop_keep, W, bias, B, T, S, L,
                                                              logits);
    KBNER_CHECK_LAUNCH("gather_tagproj_fwd");
    return KBNER_OK;
}

template <bool F32>
static int tagproj_dispatch(const void *hidden, const int32_t *row_of, const int32_t *first_idx,
                            const uint8_t *drop_keep, const float *W, const float *bias, int B, int T,
                            int S, int H, int L, float *logits, void *stream) {
    KBNER_CHECK_ARG(hidden && row_of && first_idx && W && bias && logits, "gather_tagproj_fwd: null pointer");
    KBNER_CHECK_ARG(B >= 0 && T > 0 && S > 0 && L >= 1 && L <= 32, "gather_tagproj_fwd: need 1 <= L <= 32 (L=%d)", L);
    KBNER_CHECK_ARG(H % 256 == 0 && ((size_t)L * H + (size_t)kTagprojWarps * 2 * L * 33) * 4 <= 220 * 1024,
                    "gather_tagproj_fwd: H=%d must be a multiple of 256 with L*H*4 <= 200 KB", H);
    if (B == 0) return KBNER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (H / 256) {
        case 1: return launch_tagproj<1, F32>(hidden, row_of, first_idx, drop_keep, W, bias, B, T, S, L, logits, st);
        case 2: return launch_tagproj<2, F32>(hidden, row_of, first_idx, drop_keep, W, bias, B, T, S, L, logits, st);
        case 3: return launch_tagproj<3, F32>(hidden, row_of, first_idx, drop_keep, W, bias, B, T, S, L, logits, st);
        case 4: return launch_tagproj<4, F32>(hidden, row_of, first_idx, drop_keep, W, bias, B, T, S, L, logits, st);
        default: set_error("gather_tagproj_fwd: hidden size %d not built", H); return KBNER_EUNSUPPORTED;
    }
}

extern "C" int kbner_gather_tagproj_fwd(const uint16_t *hidden, const int32_t *row_of, const int32_t *first_idx,
                                        const uint8_t *drop_keep, const float *W, const float *bias, int B, int T,
                                        int S, int H, int L, float *logits, void *stream) {
    KBNER_NVTX("kbner/elementwise");
    return tagproj_dispatch<false>(hidden, row_of, first_idx, drop_keep, W, bias, B, T, S, H, L, logits, stream);
}

extern "C" int kbner_gather_tagproj_fwd_f32(const float *hidden, const int32_t *row_of, const int32_t *first_idx,
                                            const uint8_t *drop_keep, const float *W, const float *bias, int B, int T,
                                            int S, int H, int L, float *logits, void *stream) {
    KBNER_NVTX("kbner/elementwise");
    return tagproj_dispatch<true>(hidden, row_of, first_idx, drop_keep, W, bias, B, T, S, H, L, logits, stream);
}
