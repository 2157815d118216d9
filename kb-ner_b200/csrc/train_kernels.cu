// HBM-bound kernels of the fine-tuning step (backward of LayerNorm / embedding / tag projection, bias gradients,
// gradient norm, fused AdamW).  The reference gets all of these from autograd + transformers.AdamW
// (/root/reference/flair/trainers/finetune_trainer.py:939-957 backward, :1007-1023 clip / step / zero_grad);
// here each is one coalesced pass.
#include <atomic>

#include "common.cuh"

namespace kbner {

__device__ __forceinline__ void red_add_v4(float *p, const float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void f4_add(float4 &a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// ------------------------------------------------------------------------------------------
// LayerNorm backward.  x = saved fp32 pre-LN sum, dout = grad w.r.t. the LN output (fp32),
//   xhat = (x - mean) * rstd,  g = dout * gamma,
//   dx   = rstd * (g - mean_H(g) - xhat * mean_H(g * xhat))          -> bf16 (operand of the dgrad / wgrad GEMMs)
//   dgamma += sum_rows dout * xhat,  dbeta += sum_rows dout          (fp32, accumulated)
// Mapping (second design): H / 256 warps per row, a lane owns 8 columns (two float4), a block works on two rows at a
// time.  The row sums cross the warps of a row through shared memory and a named barrier.  The first design (one warp per
// row, 32 columns per lane) needed 255 registers -- 8 warps per SM, 3.5 sequential rows each at M = 4096 -- and reduced
// its column partials with fp32 shared-memory atomics (CAS loops): 24 us for 59 MB.  Here ~32 warps per SM keep 8 rows in
// flight, the column partials of the two row slots are combined by plain shared-memory read-modify-writes and leave as
// 16-byte red.global.add.v4.f32.
// ------------------------------------------------------------------------------------------
template <int WPR, bool FUSED>
__global__ void __launch_bounds__(WPR * 64, 2)
layernorm_bwd_kernel(const float *__restrict__ x, const float *__restrict__ bias, const uint16_t *__restrict__ resid,
                     const float *__restrict__ dout, const uint16_t *__restrict__ dres, const float *__restrict__ gamma,
                     const float *__restrict__ mean, const float *__restrict__ rstd, int M,
                     uint16_t *__restrict__ dx, uint16_t *__restrict__ dxm, float *__restrict__ dgamma,
                     float *__restrict__ dbeta, float *__restrict__ dxsum, const Dropout drop) {
    constexpr int H = WPR * 256, NT = WPR * 64;
    __shared__ float2 s_part[2][2][WPR];        // [row slot][iteration parity][warp of the row]
    __shared__ float4 s_col[3][H / 4];          // column partials of row slot 0 (dgamma, dbeta, sum of dx)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp / WPR, wr = warp - slot * WPR;
    float4 gm[2], ag[2], ab[2], ax[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        gm[i] = __ldg(reinterpret_cast<const float4 *>(gamma) + wr * 64 + i * 32 + lane);
        ag[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); ax[i] = make_float4(0, 0, 0, 0);
    }
    const uint32_t key = (FUSED && drop.thresh) ? drop_key(drop) : 0u;
    // raw operands of one row for this lane; the NEXT row's are requested before the current row is reduced, so the DRAM
    // round trip of row k+1 overlaps the arithmetic, the barrier and the stores of row k (without it the capture in
    // profiles/r01/trainhbm_ncu_r38.txt shows 33 % DRAM and 42 % issue utilisation: latency-bound)
    struct RowRaw { uint4 x[2], d[2]; uint2 r[2], dr[2]; float mu, rs; };
    auto load_row = [&](int row, RowRaw &q) {
        q.mu = __ldg(mean + row);
        q.rs = __ldg(rstd + row);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c4 = wr * 64 + i * 32 + lane;
            q.x[i] = ld_nc_v4(reinterpret_cast<const float4 *>(x + (size_t)row * H) + c4);
            q.d[i] = ld_nc_v4(reinterpret_cast<const float4 *>(dout + (size_t)row * H) + c4);
            q.r[i] = make_uint2(0u, 0u);
            q.dr[i] = make_uint2(0u, 0u);
            if (FUSED && resid) q.r[i] = __ldg(reinterpret_cast<const uint2 *>(resid + (size_t)row * H) + c4);
            if (FUSED && dres) q.dr[i] = __ldg(reinterpret_cast<const uint2 *>(dres + (size_t)row * H) + c4);
        }
    };
    float4 bs[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
    if (FUSED && bias) {
#pragma unroll
        for (int i = 0; i < 2; ++i) bs[i] = __ldg(reinterpret_cast<const float4 *>(bias) + wr * 64 + i * 32 + lane);
    }
    int it = 0;
    const int row_first = blockIdx.x * 2 + slot, row_step = gridDim.x * 2;
    RowRaw cur, nxt;
    if (row_first < M) load_row(row_first, cur);
    for (int row = row_first; row < M; row += row_step, ++it) {
        if (row + row_step < M) load_row(row + row_step, nxt);
        const float mu = cur.mu, rs = cur.rs;
        float4 xh[2], g[2];
        uint32_t keep[2];                 // 4 keep bits per float4 (FUSED + dropout only)
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c4 = wr * 64 + i * 32 + lane;
            const uint4 ux = cur.x[i], ud = cur.d[i];
            float4 xv = make_float4(__uint_as_float(ux.x), __uint_as_float(ux.y), __uint_as_float(ux.z), __uint_as_float(ux.w));
            float4 dv = make_float4(__uint_as_float(ud.x), __uint_as_float(ud.y), __uint_as_float(ud.z), __uint_as_float(ud.w));
            keep[i] = 0xfu;
            if (FUSED) {
                // the LayerNorm input is recomputed exactly as the forward built it: z = dropout(x + bias) + resid
                if (bias) { xv.x += bs[i].x; xv.y += bs[i].y; xv.z += bs[i].z; xv.w += bs[i].w; }
                if (drop.thresh) {
                    const uint32_t pair = (uint32_t)row * (H / 2) + (uint32_t)c4 * 2u;
                    const uint32_t b0 = drop_bits(key, pair), b1 = drop_bits(key, pair + 1u);
                    keep[i] = (drop_keep_lo(b0, drop.thresh) ? 1u : 0u) | (drop_keep_hi(b0, drop.thresh) ? 2u : 0u) |
                              (drop_keep_lo(b1, drop.thresh) ? 4u : 0u) | (drop_keep_hi(b1, drop.thresh) ? 8u : 0u);
                    xv.x = (keep[i] & 1u) ? xv.x * drop.scale : 0.0f;
                    xv.y = (keep[i] & 2u) ? xv.y * drop.scale : 0.0f;
                    xv.z = (keep[i] & 4u) ? xv.z * drop.scale : 0.0f;
                    xv.w = (keep[i] & 8u) ? xv.w * drop.scale : 0.0f;
                }
                if (resid) {
                    float r0, r1, r2, r3;
                    unpack_bf16x2(cur.r[i].x, r0, r1);
                    unpack_bf16x2(cur.r[i].y, r2, r3);
                    xv.x += r0; xv.y += r1; xv.z += r2; xv.w += r3;
                }
                if (dres) {                 // gradient arriving over the residual connection (bf16) joins the GEMM's fp32 dgrad
                    float r0, r1, r2, r3;
                    unpack_bf16x2(cur.dr[i].x, r0, r1);
                    unpack_bf16x2(cur.dr[i].y, r2, r3);
                    dv.x += r0; dv.y += r1; dv.z += r2; dv.w += r3;
                }
            }
            xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
            g[i] = make_float4(dv.x * gm[i].x, dv.y * gm[i].y, dv.z * gm[i].z, dv.w * gm[i].w);
            s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
            ag[i].x += dv.x * xh[i].x; ag[i].y += dv.y * xh[i].y; ag[i].z += dv.z * xh[i].z; ag[i].w += dv.w * xh[i].w;
            ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (WPR > 1) {
            // the warps of a row slot run the same trip count; parity double-buffers the exchange (a warp can be at most
            // one barrier ahead of the slowest reader)
            if (lane == 0) s_part[slot][it & 1][wr] = make_float2(s1, s2);
            if (slot == 0) asm volatile("bar.sync 1, %0;" ::"n"(WPR * 32) : "memory");
            else asm volatile("bar.sync 2, %0;" ::"n"(WPR * 32) : "memory");
            s1 = 0.0f; s2 = 0.0f;
#pragma unroll
            for (int k = 0; k < WPR; ++k) { const float2 t = s_part[slot][it & 1][k]; s1 += t.x; s2 += t.y; }
        }
        const float m1 = s1 * (1.0f / H), m2 = s2 * (1.0f / H);
        uint16_t *o = dx + (size_t)row * H;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c4 = wr * 64 + i * 32 + lane;
            float d0 = rs * (g[i].x - m1 - xh[i].x * m2), d1 = rs * (g[i].y - m1 - xh[i].y * m2);
            float d2 = rs * (g[i].z - m1 - xh[i].z * m2), d3 = rs * (g[i].w - m1 - xh[i].w * m2);
            uint2 p;
            p.x = pack_bf16x2(d0, d1);
            p.y = pack_bf16x2(d2, d3);
            *reinterpret_cast<uint2 *>(o + c4 * 4) = p;                          // grad w.r.t. z: the residual path
            if (FUSED && drop.thresh) {
                // grad w.r.t. the Linear's output: through the dropout mask (what dgrad / wgrad / the bias gradient consume)
                d0 = (keep[i] & 1u) ? d0 * drop.scale : 0.0f;
                d1 = (keep[i] & 2u) ? d1 * drop.scale : 0.0f;
                d2 = (keep[i] & 4u) ? d2 * drop.scale : 0.0f;
                d3 = (keep[i] & 8u) ? d3 * drop.scale : 0.0f;
                p.x = pack_bf16x2(d0, d1);
                p.y = pack_bf16x2(d2, d3);
                *reinterpret_cast<uint2 *>(dxm + (size_t)row * H + c4 * 4) = p;
            }
            ax[i].x += d0; ax[i].y += d1; ax[i].z += d2; ax[i].w += d3;
        }
        cur = nxt;
    }
    // column partials: slot 0 publishes, slot 1 adds its own on top and flushes
    if (slot == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c4 = wr * 64 + i * 32 + lane;
            s_col[0][c4] = ag[i]; s_col[1][c4] = ab[i]; s_col[2][c4] = ax[i];
        }
    }
    __syncthreads();
    if (slot == 1) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c4 = wr * 64 + i * 32 + lane;
            f4_add(ag[i], s_col[0][c4]); f4_add(ab[i], s_col[1][c4]); f4_add(ax[i], s_col[2][c4]);
            red_add_v4(dgamma + c4 * 4, ag[i]);
            red_add_v4(dbeta + c4 * 4, ab[i]);
            if (dxsum) red_add_v4(dxsum + c4 * 4, ax[i]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// bias gradient: db[n] += sum_m dY[m][n]   (dY bf16).  Block = 32 x 8 threads over a 256-column strip; a thread walks its
// rows eight at a time (eight independent 16-byte loads in flight) and the eight row groups of a block are combined in
// shared memory before ONE 16-byte red per 4 columns.  The first version used 96 row groups per strip and 8 scalar
// atomics per thread: 393 K same-address atomics for a 4096 x 4096 matrix, 15 us for 33 MB.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const uint16_t *__restrict__ dY, int M, int N, float *__restrict__ db) {
    __shared__ float4 s[8][32][2];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = (blockIdx.x * 32 + tx) * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (col < N) {
        const int step = gridDim.y * 8;
        int r = blockIdx.y * 8 + ty;
        for (; r + 7 * step < M; r += 8 * step) {
            uint4 u[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] = ld_nc_v4(dY + (size_t)(r + k * step) * N + col);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float a, b;
                unpack_bf16x2(u[k].x, a, b); acc[0] += a; acc[1] += b;
                unpack_bf16x2(u[k].y, a, b); acc[2] += a; acc[3] += b;
                unpack_bf16x2(u[k].z, a, b); acc[4] += a; acc[5] += b;
                unpack_bf16x2(u[k].w, a, b); acc[6] += a; acc[7] += b;
            }
        }
        for (; r < M; r += step) {
            const uint4 u = ld_nc_v4(dY + (size_t)r * N + col);
            float a, b;
            unpack_bf16x2(u.x, a, b); acc[0] += a; acc[1] += b;
            unpack_bf16x2(u.y, a, b); acc[2] += a; acc[3] += b;
            unpack_bf16x2(u.z, a, b); acc[4] += a; acc[5] += b;
            unpack_bf16x2(u.w, a, b); acc[6] += a; acc[7] += b;
        }
    }
    s[ty][tx][0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    s[ty][tx][1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    __syncthreads();
    if (ty < 2 && col < N) {              // warp 0 sums the low four columns of every lane, warp 1 the high four
        float4 t = s[0][tx][ty];
#pragma unroll
        for (int y = 1; y < 8; ++y) f4_add(t, s[y][tx][ty]);
        red_add_v4(db + col + ty * 4, t);
    }
}

// ------------------------------------------------------------------------------------------
// Embedding + LayerNorm backward: recompute x = word[id] + type[0] + pos[p] and its statistics, LayerNorm backward,
// scatter-add dx into the word / position rows (fp32 atomics), type row and dgamma/dbeta reduced per block.
// One warp per sub-token; position ids recomputed exactly as in the forward kernel.
// ------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(128)
embed_ln_bwd_kernel(const int32_t *__restrict__ ids, const float *__restrict__ word_emb,
                    const float *__restrict__ pos_emb, const float *__restrict__ type_emb,
                    const float *__restrict__ gamma, float eps, int pad_id, int S, const float *__restrict__ dout,
                    float *__restrict__ d_word, float *__restrict__ d_pos, float *__restrict__ d_type,
                    float *__restrict__ dgamma, float *__restrict__ dbeta) {
    constexpr int H = VPL * 128;
    constexpr int TOK = 16;
    __shared__ float s_red[3][H];
    const int r = blockIdx.y, s0 = blockIdx.x * TOK;
    const int32_t *idr = ids + (size_t)r * S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) (&s_red[0][0])[i] = 0.0f;
    int before = 0;
    for (int base = 0; base < s0; base += 128) {
        const int t = base + threadIdx.x;
        before += __syncthreads_count(t < s0 && idr[t] != pad_id);
    }
    __syncthreads();
    const int tl = s0 + (lane & (TOK - 1));
    const int my_id = (tl < S) ? idr[tl] : pad_id;
    const unsigned nonpad = __ballot_sync(0xffffffffu, my_id != pad_id) & 0xffffu;
#pragma unroll 1
    for (int q = 0; q < TOK / 4; ++q) {
        const int local = warp * (TOK / 4) + q;
        const int sidx = s0 + local;
        if (sidx >= S) break;
        const int id = __shfl_sync(0xffffffffu, my_id, local);
        int p = pad_id;
        if (id != pad_id) p = before + __popc(nonpad & ((2u << local) - 1u)) + pad_id;
        const float4 *wr = reinterpret_cast<const float4 *>(word_emb + (size_t)id * H);
        const float4 *pr = reinterpret_cast<const float4 *>(pos_emb + (size_t)p * H);
        const float4 *tr = reinterpret_cast<const float4 *>(type_emb);
        const float4 *dr = reinterpret_cast<const float4 *>(dout + ((size_t)r * S + sidx) * H);
        float4 v[VPL];
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 a = __ldg(wr + i * 32 + lane), b = __ldg(tr + i * 32 + lane), c = __ldg(pr + i * 32 + lane);
            v[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(sum) * (1.0f / H);
        float sq = 0.0f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            sq += (a * a + b * b) + (c * c + d * d);
        }
        const float rs = rsqrtf(warp_sum(sq) * (1.0f / H) + eps);
        float4 g[VPL];
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 dv = __ldg(dr + i * 32 + lane);
            const float4 gm = __ldg(reinterpret_cast<const float4 *>(gamma) + i * 32 + lane);
            v[i] = make_float4((v[i].x - mean) * rs, (v[i].y - mean) * rs, (v[i].z - mean) * rs, (v[i].w - mean) * rs);
            g[i] = make_float4(dv.x * gm.x, dv.y * gm.y, dv.z * gm.z, dv.w * gm.w);
            s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            s2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
            const int c = (i * 32 + lane) * 4;
            atomicAdd(&s_red[0][c + 0], dv.x * v[i].x); atomicAdd(&s_red[0][c + 1], dv.y * v[i].y);
            atomicAdd(&s_red[0][c + 2], dv.z * v[i].z); atomicAdd(&s_red[0][c + 3], dv.w * v[i].w);
            atomicAdd(&s_red[1][c + 0], dv.x); atomicAdd(&s_red[1][c + 1], dv.y);
            atomicAdd(&s_red[1][c + 2], dv.z); atomicAdd(&s_red[1][c + 3], dv.w);
        }
        const float m1 = warp_sum(s1) * (1.0f / H), m2 = warp_sum(s2) * (1.0f / H);
        float *dw = d_word + (size_t)id * H, *dp = d_pos + (size_t)p * H;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = (i * 32 + lane) * 4;
            const float e0 = rs * (g[i].x - m1 - v[i].x * m2), e1 = rs * (g[i].y - m1 - v[i].y * m2);
            const float e2 = rs * (g[i].z - m1 - v[i].z * m2), e3 = rs * (g[i].w - m1 - v[i].w * m2);
            atomicAdd(dw + c + 0, e0); atomicAdd(dw + c + 1, e1); atomicAdd(dw + c + 2, e2); atomicAdd(dw + c + 3, e3);
            atomicAdd(dp + c + 0, e0); atomicAdd(dp + c + 1, e1); atomicAdd(dp + c + 2, e2); atomicAdd(dp + c + 3, e3);
            atomicAdd(&s_red[2][c + 0], e0); atomicAdd(&s_red[2][c + 1], e1);
            atomicAdd(&s_red[2][c + 2], e2); atomicAdd(&s_red[2][c + 3], e3);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        atomicAdd(&dgamma[i], s_red[0][i]);
        atomicAdd(&dbeta[i], s_red[1][i]);
        atomicAdd(&d_type[i], s_red[2][i]);
    }
}

// ------------------------------------------------------------------------------------------
// Tag-projection backward (gather + word dropout + Linear):
//   d_hidden[row(b,t)] = keep * sum_l dlogits[b,t,l] * W[l]      (fp32 rows; the buffer is pre-zeroed, rows are unique)
//   dW[l] += sum_{b,t} dlogits[b,t,l] * keep * x[row(b,t)],  db[l] += sum dlogits[b,t,l]
// ------------------------------------------------------------------------------------------
// Mapping (second design): block = H / 8 threads, thread c owns columns [8c, 8c+8) of EVERY word the block handles and
// keeps its slice of dW in registers (LG tags x 8 columns; tags beyond LG take another pass over the words), so the
// weight gradient needs no atomics inside the word loop and leaves as one red.global.add.v4.f32 per 4 columns per block.
// dlogits and the gather rows of up to kTpbWords words are staged in shared memory per round; the next word's hidden row is
// fetched while the current one is multiplied.  The first design gave a warp one word and accumulated dW with L x 32
// fp32 shared-memory atomics (CAS loops) per lane per word: 235 us for 4080 words.
constexpr int kTpbWords = 32;
template <int CPL, int LG>
__global__ void __launch_bounds__(CPL * 32)
gather_tagproj_bwd_kernel(const uint16_t *__restrict__ hidden, const int32_t *__restrict__ row_of,
                          const int32_t *__restrict__ first_idx, const uint8_t *__restrict__ drop_keep,
                          const float *__restrict__ W, const float *__restrict__ dlogits, int B, int T, int S, int L,
                          float *__restrict__ d_hidden, float *__restrict__ dW, float *__restrict__ db) {
    constexpr int H = CPL * 256, NT = CPL * 32;
    extern __shared__ __align__(16) float sm[];     // W [L][H], then dlogits of the staged words [kTpbWords][L]
    float *w_s = sm, *dl_s = sm + (size_t)L * H;
    __shared__ long long row_s[kTpbWords];           // gather row of a staged word, -1 = no sub-token / dropped / past the end
    for (int i = threadIdx.x; i < L * H / 4; i += NT)
        reinterpret_cast<float4 *>(w_s)[i] = __ldg(reinterpret_cast<const float4 *>(W) + i);
    const int words = B * T;
    const int col = threadIdx.x * 8;
    float dbacc = 0.0f;                              // thread l < L: db[l] over this block's words
    for (int l0 = 0; l0 < L; l0 += LG) {
        float acc[LG][8];
#pragma unroll
        for (int l = 0; l < LG; ++l)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[l][e] = 0.0f;
        for (int base = blockIdx.x; base < words; base += gridDim.x * kTpbWords) {
            __syncthreads();                         // previous round consumed (and, first round, W staged)
            for (int idx = threadIdx.x; idx < kTpbWords * L; idx += NT) {
                const int i = idx / L, l = idx - i * L;
                const int w = base + i * gridDim.x;
                dl_s[idx] = (w < words) ? __ldg(dlogits + (size_t)w * L + l) : 0.0f;
            }
            if (threadIdx.x < kTpbWords) {
                const int w = base + threadIdx.x * gridDim.x;
                long long r = -1;
                if (w < words) {
                    const int b = w / T, t = w - b * T;
                    const int fi = first_idx[w];
                    if (fi >= 0 && (!drop_keep || drop_keep[t] != 0)) r = (long long)row_of[b] * S + fi;
                }
                row_s[threadIdx.x] = r;
            }
            __syncthreads();
            if (l0 == 0 && threadIdx.x < L) {
#pragma unroll 4
                for (int i = 0; i < kTpbWords; ++i) dbacc += dl_s[i * L + threadIdx.x];
            }
            uint4 unext = make_uint4(0u, 0u, 0u, 0u);
            if (row_s[0] >= 0) unext = ld_nc_v4(hidden + (size_t)row_s[0] * H + col);
#pragma unroll 1
            for (int i = 0; i < kTpbWords; ++i) {
                const long long r = row_s[i];
                const uint4 u = unext;
                if (i + 1 < kTpbWords && row_s[i + 1] >= 0) unext = ld_nc_v4(hidden + (size_t)row_s[i + 1] * H + col);
                if (r < 0) continue;                 // block-uniform
                float x[8], dh[8];
                unpack_bf16x2(u.x, x[0], x[1]); unpack_bf16x2(u.y, x[2], x[3]);
                unpack_bf16x2(u.z, x[4], x[5]); unpack_bf16x2(u.w, x[6], x[7]);
#pragma unroll
                for (int e = 0; e < 8; ++e) dh[e] = 0.0f;
#pragma unroll
                for (int l = 0; l < LG; ++l) {
                    if (l0 + l < L) {                // block-uniform
                        const float d = dl_s[i * L + l0 + l];
                        const float4 w0 = *reinterpret_cast<const float4 *>(w_s + (size_t)(l0 + l) * H + col);
                        const float4 w1 = *reinterpret_cast<const float4 *>(w_s + (size_t)(l0 + l) * H + col + 4);
                        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            dh[e] = fmaf(d, wv[e], dh[e]);
                            acc[l][e] = fmaf(d, x[e], acc[l][e]);
                        }
                    }
                }
                float4 *o = reinterpret_cast<float4 *>(d_hidden + (size_t)r * H + col);
                if (l0 > 0) {                        // later tag groups add to what the first pass wrote (same thread)
                    const float4 p0 = o[0], p1 = o[1];
                    dh[0] += p0.x; dh[1] += p0.y; dh[2] += p0.z; dh[3] += p0.w;
                    dh[4] += p1.x; dh[5] += p1.y; dh[6] += p1.z; dh[7] += p1.w;
                }
                o[0] = make_float4(dh[0], dh[1], dh[2], dh[3]);
                o[1] = make_float4(dh[4], dh[5], dh[6], dh[7]);
            }
        }
#pragma unroll
        for (int l = 0; l < LG; ++l) {
            if (l0 + l < L) {
                red_add_v4(dW + (size_t)(l0 + l) * H + col, make_float4(acc[l][0], acc[l][1], acc[l][2], acc[l][3]));
                red_add_v4(dW + (size_t)(l0 + l) * H + col + 4, make_float4(acc[l][4], acc[l][5], acc[l][6], acc[l][7]));
            }
        }
    }
    if (threadIdx.x < L && dbacc != 0.0f) atomicAdd(&db[threadIdx.x], dbacc);
}

// ------------------------------------------------------------------------------------------
// Gradient norm (sum of squares, accumulated into out[0]) and fused AdamW over a flat parameter arena.
// AdamW = transformers-3.0.0 `AdamW` as constructed at finetune_trainer.py:552-571 (betas 0.9/0.999, eps 1e-6,
// correct_bias=True, weight_decay 0 applied decoupled AFTER the Adam update):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps);  p -= lr*wd*p
// g is scaled by `gscale` first (1/accumulation x clip coefficient, :939-946, :1010).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float *__restrict__ g, size_t n, float *__restrict__ out) {
    float acc = 0.0f;
    const size_t n4 = n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(g) + i);
        acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float t = g[n4 * 4 + threadIdx.x]; acc += t * t; }
    acc = warp_sum(acc);
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = s[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
        if (threadIdx.x == 0) atomicAdd(out, t);
    }
}

__global__ void __launch_bounds__(256)
adamw_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, size_t n,
             float lr, float b1, float b2, float eps, float wd, float step_size, const float *__restrict__ gscale_ptr,
             float gscale_host) {
    // clip coefficient may live on the device (computed from the norm without a host sync)
    const float gs = gscale_host * (gscale_ptr ? *gscale_ptr : 1.0f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * gs;
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float pi = p[i] - step_size * mi / (sqrtf(vi) + eps);
        if (wd != 0.0f) pi -= lr * wd * pi;
        p[i] = pi;
    }
}

// The arena form: 16-byte accesses (28 B / parameter of HBM traffic is the whole cost of the step), the gradient either
// fp32 (single GPU) or the bf16 buffer the NCCL all-reduce left behind (GBF16), and -- for the first n_shadow4 groups,
// the encoder layers -- the updated parameter written a second time as bf16 into the shadow arena the tensor-core GEMMs
// read: the 96 copy kernels that refreshed the bf16 weights after every step (a 1.8 GB pass) are gone.
template <bool GBF16>
__global__ void __launch_bounds__(256)
adamw_vec_kernel(float4 *__restrict__ p, const void *__restrict__ g, float4 *__restrict__ m, float4 *__restrict__ v, size_t n4,
                 float lr, float b1, float b2, float eps, float wd, float step_size, const float *__restrict__ gscale_ptr,
                 float gscale_host, uint2 *__restrict__ shadow, size_t n_shadow4) {
    const float gs = gscale_host * (gscale_ptr ? *gscale_ptr : 1.0f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float gi[4];
        if (GBF16) {
            const uint2 u = __ldg(reinterpret_cast<const uint2 *>(g) + i);
            unpack_bf16x2(u.x, gi[0], gi[1]);
            unpack_bf16x2(u.y, gi[2], gi[3]);
        } else {
            const uint4 u = ld_nc_v4(reinterpret_cast<const float4 *>(g) + i);
            gi[0] = __uint_as_float(u.x); gi[1] = __uint_as_float(u.y); gi[2] = __uint_as_float(u.z); gi[3] = __uint_as_float(u.w);
        }
        const float4 pm = p[i], mm = m[i], vm = v[i];
        float pi[4] = {pm.x, pm.y, pm.z, pm.w}, mi[4] = {mm.x, mm.y, mm.z, mm.w}, vi[4] = {vm.x, vm.y, vm.z, vm.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gi[k] * gs;
            mi[k] = b1 * mi[k] + (1.0f - b1) * gk;
            vi[k] = b2 * vi[k] + (1.0f - b2) * gk * gk;
            pi[k] = pi[k] - step_size * mi[k] / (sqrtf(vi[k]) + eps);
            if (wd != 0.0f) pi[k] -= lr * wd * pi[k];
        }
        m[i] = make_float4(mi[0], mi[1], mi[2], mi[3]);
        v[i] = make_float4(vi[0], vi[1], vi[2], vi[3]);
        p[i] = make_float4(pi[0], pi[1], pi[2], pi[3]);
        if (i < n_shadow4) shadow[i] = make_uint2(pack_bf16x2(pi[0], pi[1]), pack_bf16x2(pi[2], pi[3]));
    }
}

// Sparse exchange of the word-embedding gradient (data-parallel fine-tuning): an optimizer step touches at most
// tokens-per-step rows of the [250002, 1024] table, so the ranks exchange the touched ROWS (all-gather of bf16 rows + ids)
// instead of all-reducing 0.5 GB of mostly zeros.  rows_gather: rows[i] = bf16(src[ids[i]]) (zeros for ids[i] < 0) and,
// with zero_src, the source row is cleared so that rows_scatter_add can rebuild it as the sum over ranks IN RANK ORDER --
// the same association on every rank, replicas stay bit-identical.  ids are unique within a call (sorted, duplicates
// replaced by -1 by the caller): no atomics.  One warp per row, 16-byte accesses.
__global__ void __launch_bounds__(256)
rows_gather_bf16_kernel(float *__restrict__ src, const int32_t *__restrict__ ids, int n, int H, uint16_t *__restrict__ rows,
                        int zero_src) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const int id = __ldg(ids + i);
    uint2 *out = reinterpret_cast<uint2 *>(rows + (size_t)i * H);
    float4 *row = (id >= 0) ? reinterpret_cast<float4 *>(src + (size_t)id * H) : nullptr;
    for (int c = lane; c < H / 4; c += 32) {
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (row) {
            v = row[c];
            if (zero_src) row[c] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        out[c] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

__global__ void __launch_bounds__(256)
rows_scatter_add_bf16_kernel(const uint16_t *__restrict__ rows, const int32_t *__restrict__ ids, int n, int H,
                             float *__restrict__ dst) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const int id = __ldg(ids + i);
    if (id < 0) return;
    const uint2 *in = reinterpret_cast<const uint2 *>(rows + (size_t)i * H);
    float4 *row = reinterpret_cast<float4 *>(dst + (size_t)id * H);
    for (int c = lane; c < H / 4; c += 32) {
        const uint2 u = __ldg(in + c);
        float a, b, e, f;
        unpack_bf16x2(u.x, a, b);
        unpack_bf16x2(u.y, e, f);
        float4 v = row[c];
        v.x += a; v.y += b; v.z += e; v.w += f;
        row[c] = v;
    }
}

// Row-sparse optimizer work on the word-embedding table (250 002 x 1024 = 46 % of XLM-R-large's parameters): a row no
// sentence has ever touched has g = m = v = 0, so AdamW leaves it where it is (p -= step * 0 / (0 + eps); weight decay is 0
// in the reference's groups, finetune_trainer.py:552-571) -- reading and rewriting it (28 B per parameter) every optimizer
// step is pure HBM traffic.  `touched[row]` is set the first time a sub-token id is embedded (mark_rows) and never cleared:
// rows with momentum keep being updated.  The three passes over the table (clip norm, AdamW, zero_grad) visit flagged rows
// only; per element the arithmetic is the dense kernels', so the parameters are bit-identical to the dense path.
__global__ void __launch_bounds__(256) mark_rows_kernel(const int32_t *__restrict__ ids, size_t n, int V, uint8_t *__restrict__ touched) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int id = ids[i];
        if (id >= 0 && id < V) touched[id] = 1;
    }
}

__global__ void __launch_bounds__(256)
adamw_rows_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                  const uint8_t *__restrict__ touched, int V, int H, float lr, float b1, float b2, float eps, float wd,
                  float step_size, const float *__restrict__ gscale_ptr, float gscale_host) {
    const float gs = gscale_host * (gscale_ptr ? *gscale_ptr : 1.0f);
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < V; row += gridDim.x * 8) {
        if (!touched[row]) continue;
        const size_t base = (size_t)row * H / 4;
        for (int c = lane; c < H / 4; c += 32) {
            const float4 gq = reinterpret_cast<const float4 *>(g)[base + c];
            float4 pq = reinterpret_cast<float4 *>(p)[base + c], mq = reinterpret_cast<float4 *>(m)[base + c],
                   vq = reinterpret_cast<float4 *>(v)[base + c];
            float gi[4] = {gq.x, gq.y, gq.z, gq.w}, pi[4] = {pq.x, pq.y, pq.z, pq.w}, mi[4] = {mq.x, mq.y, mq.z, mq.w},
                  vi[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gk = gi[k] * gs;
                mi[k] = b1 * mi[k] + (1.0f - b1) * gk;
                vi[k] = b2 * vi[k] + (1.0f - b2) * gk * gk;
                pi[k] = pi[k] - step_size * mi[k] / (sqrtf(vi[k]) + eps);
                if (wd != 0.0f) pi[k] -= lr * wd * pi[k];
            }
            reinterpret_cast<float4 *>(m)[base + c] = make_float4(mi[0], mi[1], mi[2], mi[3]);
            reinterpret_cast<float4 *>(v)[base + c] = make_float4(vi[0], vi[1], vi[2], vi[3]);
            reinterpret_cast<float4 *>(p)[base + c] = make_float4(pi[0], pi[1], pi[2], pi[3]);
        }
    }
}

__global__ void __launch_bounds__(256)
sumsq_rows_partial_kernel(const float *__restrict__ g, const uint8_t *__restrict__ touched, int V, int H, float *__restrict__ partials) {
    float acc = 0.0f;
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < V; row += gridDim.x * 8) {
        if (!touched[row]) continue;
        const size_t base = (size_t)row * H / 4;
        for (int c = lane; c < H / 4; c += 32) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(g) + base + c);
            acc += (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);
        }
    }
    acc = warp_sum(acc);
    __shared__ float s[8];
    if (lane == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) partials[blockIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

__global__ void __launch_bounds__(256) zero_rows_kernel(float *__restrict__ g, const uint8_t *__restrict__ touched, int V, int H) {
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < V; row += gridDim.x * 8) {
        if (!touched[row]) continue;
        float4 *r = reinterpret_cast<float4 *>(g) + (size_t)row * H / 4;
        for (int c = lane; c < H / 4; c += 32) r[c] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// Deterministic sum of squares.  The atomicAdd version above adds the block partials in arrival order: the clip
// coefficient then differs in its last bits from run to run AND from rank to rank -- with data-parallel fine-tuning every
// rank clips the same all-reduced gradient, and replicas whose coefficients differ by an ulp drift apart
// (tests/test_ddp_gpu.py caught it: `replicas diverged at transitions`).  Here every block writes its partial to a
// caller-owned slot and one warp adds the slots in index order, in fp64.
template <bool GBF16>
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const void *__restrict__ g, size_t n4, float *__restrict__ partials) {
    float acc = 0.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float a, b, c, d;
        if (GBF16) {
            const uint2 u = __ldg(reinterpret_cast<const uint2 *>(g) + i);
            unpack_bf16x2(u.x, a, b);
            unpack_bf16x2(u.y, c, d);
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(g) + i);
            a = v.x; b = v.y; c = v.z; d = v.w;
        }
        acc += (a * a + b * b) + (c * c + d * d);
    }
    acc = warp_sum(acc);                       // xor-shuffle tree: a fixed order
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) partials[blockIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

__global__ void __launch_bounds__(32) sumsq_final_kernel(const float *__restrict__ partials, int n, float *__restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) acc += (double)partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) out[0] = (float)((double)out[0] + acc);
}

// fp32 -> bf16 (round to nearest even) of a flat buffer: the gradient arena packed for the NCCL all-reduce, and the first
// fill of the bf16 shadow arena.  `scale` multiplies first (1.0 for plain conversion).
__global__ void __launch_bounds__(256)
pack_bf16_kernel(const float4 *__restrict__ src, uint2 *__restrict__ dst, size_t n4, float scale) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 u = ld_nc_v4(src + i);
        dst[i] = make_uint2(pack_bf16x2(__uint_as_float(u.x) * scale, __uint_as_float(u.y) * scale),
                            pack_bf16x2(__uint_as_float(u.z) * scale, __uint_as_float(u.w) * scale));
    }
}

__global__ void __launch_bounds__(256) sumsq_bf16_kernel(const uint2 *__restrict__ g, size_t n4, float *__restrict__ out) {
    float acc = 0.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint2 u = __ldg(g + i);
        float a, b, c, d;
        unpack_bf16x2(u.x, a, b);
        unpack_bf16x2(u.y, c, d);
        acc += (a * a + b * b) + (c * c + d * d);
    }
    acc = warp_sum(acc);
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = s[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
        if (threadIdx.x == 0) atomicAdd(out, t);
    }
}

// clip coefficient on the device: coef = min(1, max_norm / (sqrt(sumsq) * pre + 1e-6))  (torch clip_grad_norm_)
__global__ void clip_coef_kernel(const float *__restrict__ sumsq, float pre, float max_norm, float *__restrict__ coef) {
    const float nrm = sqrtf(*sumsq) * pre;
    *coef = fminf(1.0f, max_norm / (nrm + 1e-6f));
}

// ------------------------------------------------------------------------------------------
// Element-wise dropout (in place) with the counter-hash mask of common.cuh: the embedding dropout of
// XLMRobertaEmbeddings (applied to the bf16 LayerNorm output in the forward, to the fp32 gradient in the backward).
// The per-layer dropouts are fused into the LayerNorm and attention kernels instead.
// ------------------------------------------------------------------------------------------
template <bool F32>
__global__ void __launch_bounds__(256)
dropout_apply_kernel(void *__restrict__ x, size_t n_pairs, const Dropout drop) {
    const uint32_t key = drop_key(drop);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n_pairs; q += stride) {   // 4 pairs = 8 elements
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t b = drop_bits(key, (uint32_t)(q * 4 + j));
            keep |= (drop_keep_lo(b, drop.thresh) ? 1u : 0u) << (2 * j);
            keep |= (drop_keep_hi(b, drop.thresh) ? 1u : 0u) << (2 * j + 1);
        }
        if (F32) {
            float4 *p = reinterpret_cast<float4 *>(x) + q * 2;
            float4 a = p[0], b = p[1];
            a.x = (keep & 1u) ? a.x * drop.scale : 0.0f;   a.y = (keep & 2u) ? a.y * drop.scale : 0.0f;
            a.z = (keep & 4u) ? a.z * drop.scale : 0.0f;   a.w = (keep & 8u) ? a.w * drop.scale : 0.0f;
            b.x = (keep & 16u) ? b.x * drop.scale : 0.0f;  b.y = (keep & 32u) ? b.y * drop.scale : 0.0f;
            b.z = (keep & 64u) ? b.z * drop.scale : 0.0f;  b.w = (keep & 128u) ? b.w * drop.scale : 0.0f;
            p[0] = a; p[1] = b;
        } else {
            uint4 *p = reinterpret_cast<uint4 *>(x) + q;
            uint4 u = *p;
            uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a, b;
                unpack_bf16x2(w[j], a, b);
                a = ((keep >> (2 * j)) & 1u) ? a * drop.scale : 0.0f;
                b = ((keep >> (2 * j + 1)) & 1u) ? b * drop.scale : 0.0f;
                w[j] = pack_bf16x2(a, b);
            }
            *p = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

}  // namespace kbner

using namespace kbner;

#define DISPATCH_VPL_T(H, CALL)                                                                   \
    switch ((H) / 128) {                                                                          \
        case 2: { constexpr int VPL = 2; CALL; } break;                                           \
        case 4: { constexpr int VPL = 4; CALL; } break;                                           \
        case 6: { constexpr int VPL = 6; CALL; } break;                                           \
        case 8: { constexpr int VPL = 8; CALL; } break;                                           \
        default: set_error("hidden size %d not built (supported: 256, 512, 768, 1024)", (H));     \
                 return KBNER_EUNSUPPORTED;                                                       \
    }

extern "C" int kbner_add_layernorm_bwd(const float *x, const float *bias, const uint16_t *resid, const float *dout,
                                       const uint16_t *dres, const float *gamma, const float *mean, const float *rstd,
                                       int M, int H, uint16_t *dx, uint16_t *dx_masked, float *dgamma, float *dbeta,
                                       float *dxsum, const uint32_t *drop_seed, uint32_t drop_site, float drop_p,
                                       void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(x && dout && gamma && mean && rstd && dx && dgamma && dbeta, "layernorm_bwd: null pointer");
    KBNER_CHECK_ARG(M >= 0 && H % 128 == 0, "layernorm_bwd: bad shape");
    KBNER_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f, "layernorm_bwd: dropout probability %f", (double)drop_p);
    const Dropout drop = make_dropout(drop_seed, drop_site, drop_p);
    KBNER_CHECK_ARG(!drop.thresh || dx_masked, "layernorm_bwd: dropout needs the dx_masked output");
    KBNER_CHECK_ARG((uint64_t)M * (uint64_t)(H / 2) < (1ull << 32), "layernorm_bwd: M*H/2 exceeds the 32-bit dropout counter");
    if (M == 0) return KBNER_OK;
    KBNER_CHECK_ARG(H % 256 == 0, "layernorm_bwd: hidden size %d must be a multiple of 256", H);
    int blocks = (M + 1) / 2;                     // two rows per block pass
    if (blocks > 2 * num_sms()) blocks = 2 * num_sms();   // <= 128 registers x 256 threads: two blocks per SM, one row prefetched each
    cudaStream_t st = (cudaStream_t)stream;
    if (bias || resid || dres || drop.thresh) {
        DISPATCH_VPL_T(H, (layernorm_bwd_kernel<VPL / 2, true><<<blocks, VPL * 32, 0, st>>>(x, bias, resid, dout, dres, gamma, mean, rstd, M,
                                                                                          dx, dx_masked, dgamma, dbeta, dxsum, drop)));
    } else {
        DISPATCH_VPL_T(H, (layernorm_bwd_kernel<VPL / 2, false><<<blocks, VPL * 32, 0, st>>>(x, bias, resid, dout, dres, gamma, mean, rstd, M,
                                                                                           dx, dx_masked, dgamma, dbeta, dxsum, drop)));
    }
    KBNER_CHECK_LAUNCH("layernorm_bwd");
    return KBNER_OK;
}

extern "C" int kbner_layernorm_bwd(const float *x, const float *dout, const float *gamma, const float *mean,
                                   const float *rstd, int M, int H, uint16_t *dx, float *dgamma, float *dbeta,
                                   float *dxsum, void *stream) {
    KBNER_NVTX("kbner/train");
    return kbner_add_layernorm_bwd(x, nullptr, nullptr, dout, nullptr, gamma, mean, rstd, M, H, dx, nullptr, dgamma, dbeta,
                                   dxsum, nullptr, 0u, 0.0f, stream);
}

extern "C" int kbner_colsum_bf16(const uint16_t *dY, int M, int N, float *db, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(dY && db && M >= 0 && N > 0 && N % 8 == 0, "colsum_bf16: bad arguments");
    if (M == 0) return KBNER_OK;
    dim3 grid((N + 255) / 256, 1);
    grid.y = (4 * num_sms() + grid.x - 1) / grid.x;             // ~4 blocks per SM in total, 8 loads in flight per thread
    if ((int)grid.y * 8 > M) grid.y = (M + 7) / 8;
    colsum_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dY, M, N, db);
    KBNER_CHECK_LAUNCH("colsum_bf16");
    return KBNER_OK;
}

extern "C" int kbner_embed_ln_bwd(const int32_t *ids, const float *word_emb, const float *pos_emb,
                                  const float *type_emb, const float *gamma, float eps, int pad_id, int R, int S,
                                  int H, const float *dout, float *d_word, float *d_pos, float *d_type,
                                  float *dgamma, float *dbeta, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(ids && word_emb && pos_emb && type_emb && gamma && dout && d_word && d_pos && d_type && dgamma && dbeta,
                    "embed_ln_bwd: null pointer");
    if (R == 0) return KBNER_OK;
    dim3 grid((S + 15) / 16, R);
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_VPL_T(H, (embed_ln_bwd_kernel<VPL><<<grid, 128, 0, st>>>(ids, word_emb, pos_emb, type_emb, gamma, eps, pad_id, S,
                                                                       dout, d_word, d_pos, d_type, dgamma, dbeta)));
    KBNER_CHECK_LAUNCH("embed_ln_bwd");
    return KBNER_OK;
}

template <int CPL>
static int launch_tagproj_bwd(const uint16_t *hidden, const int32_t *row_of, const int32_t *first_idx,
                              const uint8_t *drop_keep, const float *W, const float *dlogits, int B, int T, int S, int L,
                              float *d_hidden, float *dW, float *db, cudaStream_t st) {
    constexpr int LG = 16;
    const size_t smem = ((size_t)L * CPL * 256 + (size_t)kTpbWords * L) * sizeof(float);
    static std::atomic<size_t> configured{0};     // idempotent set-up: a race only repeats it
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(gather_tagproj_bwd_kernel<CPL, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("gather_tagproj_bwd: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
            return KBNER_ECUDA;
        }
        configured = smem;
    }
    int blocks = B * T;
    if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
    gather_tagproj_bwd_kernel<CPL, LG><<<blocks, CPL * 32, smem, st>>>(hidden, row_of, first_idx, drop_keep, W, dlogits, B, T, S, L,
                                                                      d_hidden, dW, db);
    KBNER_CHECK_LAUNCH("gather_tagproj_bwd");
    return KBNER_OK;
}

extern "C" int kbner_gather_tagproj_bwd(const uint16_t *hidden, const int32_t *row_of, const int32_t *first_idx,
                                        const uint8_t *drop_keep, const float *W, const float *dlogits, int B, int T,
                                        int S, int H, int L, float *d_hidden, float *dW, float *db, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(hidden && row_of && first_idx && W && dlogits && d_hidden && dW && db, "gather_tagproj_bwd: null pointer");
    KBNER_CHECK_ARG(L >= 1 && L <= 32 && H % 256 == 0 && (size_t)L * H * 4 <= 200 * 1024,
                    "gather_tagproj_bwd: needs L <= 32 and L*H*4 <= 200 KB (L=%d H=%d)", L, H);
    if (B == 0) return KBNER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (H / 256) {
        case 1: return launch_tagproj_bwd<1>(hidden, row_of, first_idx, drop_keep, W, dlogits, B, T, S, L, d_hidden, dW, db, st);
        case 2: return launch_tagproj_bwd<2>(hidden, row_of, first_idx, drop_keep, W, dlogits, B, T, S, L, d_hidden, dW, db, st);
        case 3: return launch_tagproj_bwd<3>(hidden, row_of, first_idx, drop_keep, W, dlogits, B, T, S, L, d_hidden, dW, db, st);
        case 4: return launch_tagproj_bwd<4>(hidden, row_of, first_idx, drop_keep, W, dlogits, B, T, S, L, d_hidden, dW, db, st);
        default: set_error("gather_tagproj_bwd: hidden size %d not built", H); return KBNER_EUNSUPPORTED;
    }
}

extern "C" int kbner_sumsq_f32(const float *g, size_t n, float *out, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(g && out, "sumsq: null pointer");
    KBNER_CHECK_ARG(((uintptr_t)g & 15u) == 0, "sumsq: buffer must be 16-byte aligned");
    if (n == 0) return KBNER_OK;
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks > 8 * (size_t)num_sms()) blocks = 8 * num_sms();
    if (blocks == 0) blocks = 1;
    sumsq_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, n, out);
    KBNER_CHECK_LAUNCH("sumsq");
    return KBNER_OK;
}

extern "C" int kbner_clip_coef(const float *sumsq, float pre_scale, float max_norm, float *coef, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(sumsq && coef, "clip_coef: null pointer");
    clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, pre_scale, max_norm, coef);
    KBNER_CHECK_LAUNCH("clip_coef");
    return KBNER_OK;
}

extern "C" int kbner_adamw_step_ex(float *p, const float *g, const uint16_t *g_bf16, float *m, float *v, size_t n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, int step,
                                   const float *gscale_dev, float gscale_host, uint16_t *shadow, size_t n_shadow,
                                   void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(p && (g != nullptr) != (g_bf16 != nullptr) && m && v && step >= 1,
                    "adamw_step: bad arguments (exactly one of g / g_bf16)");
    KBNER_CHECK_ARG(!shadow || (n_shadow <= n && n_shadow % 4 == 0), "adamw_step: shadow length %zu", n_shadow);
    if (n == 0) return KBNER_OK;
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = (((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g) & 15u) == 0 &&
                         ((uintptr_t)g_bf16 & 7u) == 0 && ((uintptr_t)shadow & 7u) == 0;
    const size_t n4 = aligned ? n / 4 : 0;
    KBNER_CHECK_ARG(aligned || (!g_bf16 && !shadow), "adamw_step: bf16 gradient / shadow need 16-byte aligned arenas");
    if (n4) {
        size_t blocks = (n4 + 255) / 256;
        if (blocks > 16 * (size_t)num_sms()) blocks = 16 * num_sms();
        const size_t ns4 = shadow ? n_shadow / 4 : 0;
        if (g_bf16)
            adamw_vec_kernel<true><<<(int)blocks, 256, 0, st>>>((float4 *)p, g_bf16, (float4 *)m, (float4 *)v, n4, lr, beta1, beta2, eps,
                                                              weight_decay, step_size, gscale_dev, gscale_host, (uint2 *)shadow, ns4);
        else
            adamw_vec_kernel<false><<<(int)blocks, 256, 0, st>>>((float4 *)p, g, (float4 *)m, (float4 *)v, n4, lr, beta1, beta2, eps,
                                                               weight_decay, step_size, gscale_dev, gscale_host, (uint2 *)shadow, ns4);
        KBNER_CHECK_LAUNCH("adamw_step");
    }
    const size_t done = n4 * 4;
    if (done < n) {
        KBNER_CHECK_ARG(!g_bf16, "adamw_step: a bf16 gradient buffer must have a multiple of 4 elements");
        size_t blocks = (n - done + 255) / 256;
        if (blocks > 16 * (size_t)num_sms()) blocks = 16 * num_sms();
        adamw_kernel<<<(int)blocks, 256, 0, st>>>(p + done, g + done, m + done, v + done, n - done, lr, beta1, beta2, eps,
                                                  weight_decay, step_size, gscale_dev, gscale_host);
        KBNER_CHECK_LAUNCH("adamw_step");
    }
    return KBNER_OK;
}

extern "C" int kbner_adamw_step(float *p, const float *g, float *m, float *v, size_t n, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int step, const float *gscale_dev,
                                float gscale_host, void *stream) {
    KBNER_NVTX("kbner/train");
    return kbner_adamw_step_ex(p, g, nullptr, m, v, n, lr, beta1, beta2, eps, weight_decay, step, gscale_dev, gscale_host,
                               nullptr, 0, stream);
}

extern "C" int kbner_pack_bf16(const float *src, uint16_t *dst, size_t n, float scale, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(src && dst, "pack_bf16: null pointer");
    KBNER_CHECK_ARG(n % 4 == 0 && ((uintptr_t)src & 15u) == 0 && ((uintptr_t)dst & 7u) == 0,
                    "pack_bf16: n must be a multiple of 4 and the buffers 16- / 8-byte aligned");
    if (n == 0) return KBNER_OK;
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks > 16 * (size_t)num_sms()) blocks = 16 * num_sms();
    pack_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)src, (uint2 *)dst, n / 4, scale);
    KBNER_CHECK_LAUNCH("pack_bf16");
    return KBNER_OK;
}

extern "C" int kbner_mark_rows(const int32_t *ids, size_t n, int V, uint8_t *touched, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(ids && touched && V > 0, "mark_rows: null pointer");
    if (n == 0) return KBNER_OK;
    size_t blocks = (n + 255) / 256;
    if (blocks > 4 * (size_t)num_sms()) blocks = 4 * num_sms();
    mark_rows_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(ids, n, V, touched);
    KBNER_CHECK_LAUNCH("mark_rows");
    return KBNER_OK;
}

static int rows_grid(int V) {
    int blocks = (V + 7) / 8;
    const int cap = 16 * num_sms();
    return blocks < cap ? blocks : cap;
}

extern "C" int kbner_adamw_rows(float *p, const float *g, float *m, float *v, const uint8_t *touched, int V, int H, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int step, const float *gscale_dev,
                                float gscale_host, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(p && g && m && v && touched && step >= 1, "adamw_rows: bad arguments");
    KBNER_CHECK_ARG(V > 0 && H > 0 && H % 4 == 0 && (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15u) == 0,
                    "adamw_rows: H=%d must be a multiple of 4 and the buffers 16-byte aligned", H);
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
    adamw_rows_kernel<<<rows_grid(V), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, touched, V, H, lr, beta1, beta2, eps, weight_decay,
                                                                      step_size, gscale_dev, gscale_host);
    KBNER_CHECK_LAUNCH("adamw_rows");
    return KBNER_OK;
}

extern "C" int kbner_sumsq_rows_det(const float *g, const uint8_t *touched, int V, int H, float *partials, int n_partials,
                                    float *out, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(g && touched && partials && out && n_partials >= 32, "sumsq_rows_det: null pointer / fewer than 32 partial slots");
    KBNER_CHECK_ARG(V > 0 && H > 0 && H % 4 == 0 && ((uintptr_t)g & 15u) == 0, "sumsq_rows_det: H=%d / alignment", H);
    int blocks = rows_grid(V);
    if (blocks > n_partials) blocks = n_partials;
    cudaStream_t st = (cudaStream_t)stream;
    sumsq_rows_partial_kernel<<<blocks, 256, 0, st>>>(g, touched, V, H, partials);
    KBNER_CHECK_LAUNCH("sumsq_rows_partial");
    sumsq_final_kernel<<<1, 32, 0, st>>>(partials, blocks, out);
    KBNER_CHECK_LAUNCH("sumsq_final");
    return KBNER_OK;
}

extern "C" int kbner_zero_rows(float *g, const uint8_t *touched, int V, int H, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(g && touched && V > 0 && H > 0 && H % 4 == 0 && ((uintptr_t)g & 15u) == 0, "zero_rows: bad arguments");
    zero_rows_kernel<<<rows_grid(V), 256, 0, (cudaStream_t)stream>>>(g, touched, V, H);
    KBNER_CHECK_LAUNCH("zero_rows");
    return KBNER_OK;
}

extern "C" int kbner_rows_gather_bf16(float *src, const int32_t *ids, int n, int V, int H, uint16_t *rows, int zero_src,
                                      void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(src && ids && rows, "rows_gather: null pointer");
    KBNER_CHECK_ARG(n >= 0 && V > 0 && H > 0 && H % 4 == 0 && (((uintptr_t)src | (uintptr_t)rows) & 15u) == 0,
                    "rows_gather: H=%d must be a multiple of 4 and the buffers 16-byte aligned", H);
    if (n == 0) return KBNER_OK;
    rows_gather_bf16_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(src, ids, n, H, rows, zero_src);
    KBNER_CHECK_LAUNCH("rows_gather");
    return KBNER_OK;
}

extern "C" int kbner_rows_scatter_add_bf16(const uint16_t *rows, const int32_t *ids, int n, int V, int H, float *dst,
                                           void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(rows && ids && dst, "rows_scatter_add: null pointer");
    KBNER_CHECK_ARG(n >= 0 && V > 0 && H > 0 && H % 4 == 0 && (((uintptr_t)dst | (uintptr_t)rows) & 15u) == 0,
                    "rows_scatter_add: H=%d must be a multiple of 4 and the buffers 16-byte aligned", H);
    if (n == 0) return KBNER_OK;
    rows_scatter_add_bf16_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(rows, ids, n, H, dst);
    KBNER_CHECK_LAUNCH("rows_scatter_add");
    return KBNER_OK;
}

extern "C" int kbner_sumsq_det(const void *g, size_t n, int is_bf16, float *partials, int n_partials, float *out, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(g && partials && out && n_partials >= 32, "sumsq_det: null pointer / fewer than 32 partial slots");
    KBNER_CHECK_ARG(n % 4 == 0 && ((uintptr_t)g & (is_bf16 ? 7u : 15u)) == 0, "sumsq_det: n must be a multiple of 4, buffer aligned");
    if (n == 0) return KBNER_OK;
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks > 8 * (size_t)num_sms()) blocks = 8 * num_sms();
    if (blocks > (size_t)n_partials) blocks = (size_t)n_partials;
    cudaStream_t st = (cudaStream_t)stream;
    if (is_bf16) sumsq_partial_kernel<true><<<(int)blocks, 256, 0, st>>>(g, n / 4, partials);
    else sumsq_partial_kernel<false><<<(int)blocks, 256, 0, st>>>(g, n / 4, partials);
    KBNER_CHECK_LAUNCH("sumsq_partial");
    sumsq_final_kernel<<<1, 32, 0, st>>>(partials, (int)blocks, out);
    KBNER_CHECK_LAUNCH("sumsq_final");
    return KBNER_OK;
}

extern "C" int kbner_sumsq_bf16(const uint16_t *g, size_t n, float *out, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(g && out, "sumsq_bf16: null pointer");
    KBNER_CHECK_ARG(n % 4 == 0 && ((uintptr_t)g & 7u) == 0, "sumsq_bf16: n must be a multiple of 4, buffer 8-byte aligned");
    if (n == 0) return KBNER_OK;
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks > 8 * (size_t)num_sms()) blocks = 8 * num_sms();
    sumsq_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const uint2 *)g, n / 4, out);
    KBNER_CHECK_LAUNCH("sumsq_bf16");
    return KBNER_OK;
}

extern "C" int kbner_dropout_apply(void *x, int is_f32, int M, int H, const uint32_t *drop_seed, uint32_t drop_site,
                                   float drop_p, void *stream) {
    KBNER_NVTX("kbner/train");
    KBNER_CHECK_ARG(x && drop_seed, "dropout_apply: null pointer");
    KBNER_CHECK_ARG(M >= 0 && H > 0 && H % 8 == 0, "dropout_apply: H=%d must be a multiple of 8", H);
    KBNER_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f, "dropout_apply: dropout probability %f", (double)drop_p);
    KBNER_CHECK_ARG((uint64_t)M * (uint64_t)(H / 2) < (1ull << 32), "dropout_apply: M*H/2 exceeds the 32-bit dropout counter");
    const Dropout drop = make_dropout(drop_seed, drop_site, drop_p);
    if (M == 0 || !drop.thresh) return KBNER_OK;
    const size_t n_pairs = (size_t)M * H / 2;
    size_t blocks = (n_pairs / 4 + 255) / 256;
    if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
    cudaStream_t st = (cudaStream_t)stream;
    if (is_f32) dropout_apply_kernel<true><<<(int)blocks, 256, 0, st>>>(x, n_pairs, drop);
    else dropout_apply_kernel<false><<<(int)blocks, 256, 0, st>>>(x, n_pairs, drop);
    KBNER_CHECK_LAUNCH("dropout_apply");
    return KBNER_OK;
}
