// Shared helpers for the kbner_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kbner_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "kbner_b200 kernels are written for sm_100a only"
#endif

namespace kbner {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
bool pdl_enabled();   // programmatic dependent launch between the hot kernels (KBNER_PDL=0 turns it off)
int num_sms();        // SM count of the current device (148 on B200), queried once per device
int sm_budget();      // SMs the persistent tensor-core grids may occupy: num_sms() unless kbner_set_sm_budget() lowered it
// NVTX range around every C-ABI entry point, named after the kernel family ("kbner/gemm", "kbner/attention", ...), so
// that a timeline or `ncu --nvtx --nvtx-include "kbner/crf/"` groups launches by family.  Off unless KBNER_NVTX=1
// (one relaxed load per call otherwise).
void nvtx_push(const char *name);
void nvtx_pop();
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtx_push(name); }
    ~NvtxRange() { nvtx_pop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
#define KBNER_NVTX(name) ::kbner::NvtxRange nvtx_range__(name)

#define KBNER_CHECK_ARG(cond, ...)                       \
    do {                                                 \
        if (!(cond)) {                                   \
            ::kbner::set_error(__VA_ARGS__);             \
            return KBNER_EINVAL;                         \
        }                                                \
    } while (0)

#define KBNER_CHECK_LAUNCH(name)                                                           \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            ::kbner::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
            return KBNER_ECUDA;                                                            \
        }                                                                                  \
        ::kbner::count_launch();                                                           \
    } while (0)

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t lo16) { return __uint_as_float(lo16 << 16); }
__device__ __forceinline__ void unpack_bf16x2(uint32_t p, float &a, float &b) {
    a = __uint_as_float(p << 16);
    b = __uint_as_float(p & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);   // .x = a (low half), .y = b
    return *reinterpret_cast<uint32_t *>(&h);
}

// fast special-function-unit ops (1 MUFU each, ~2 ulp): enough for bf16-rounded results
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 16-byte streaming loads/stores (data touched once: do not pollute L1)
__device__ __forceinline__ uint4 ld_nc_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_na_v4(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- Blackwell packed-fp32 and 3-input ALU ops (SASS FFMA2 / FADD2 / FMNMX3): half the issue slots of the scalar forms ----
__device__ __forceinline__ void ffma2(float &d0, float &d1, float a0, float a1, float b, float c) {   // d = a * b + c on a pair
    asm("{.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fadd2(float &s0, float &s1, float a0, float a1) {                    // s += a on a pair
    asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "+f"(s0), "+f"(s1) : "f"(a0), "f"(a1));
}
// the same packed ops on values that STAY pairs (64-bit registers) across several instructions
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t p, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// ---- dropout: stateless keep mask from a counter hash ---------------------------------------------------------------
// The reference trains with torch.nn.Dropout inside transformers (hidden 0.1, attention probabilities 0.1).  Here a mask
// is never stored: element e of dropout site `site` is kept iff the 16-bit half (e & 1) of
//     fmix32((e >> 1) * 0x9E3779B1 + key),   key = f(seed[0], seed[1], site)
// is >= thresh = round(p * 65536); kept values are scaled by 1 / (1 - thresh / 65536).  The backward kernels regenerate
// the same bits.  `seed` lives in DEVICE memory so that a captured CUDA graph sees a fresh seed at every replay.
struct Dropout {
    const uint32_t *seed;   // device, 2 words; NULL = dropout off
    uint32_t site;
    uint32_t thresh;        // 0 = off
    float scale;
};
__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t drop_key(const Dropout &d) {
    return fmix32(__ldg(d.seed) + 0x9E3779B9u * (d.site + 1u)) ^ __ldg(d.seed + 1);
}
__device__ __forceinline__ uint32_t drop_bits(uint32_t key, uint32_t pair) { return fmix32(pair * 0x9E3779B1u + key); }
// keep flags of the two elements of a pair
__device__ __forceinline__ bool drop_keep_lo(uint32_t bits, uint32_t thresh) { return (bits & 0xffffu) >= thresh; }
__device__ __forceinline__ bool drop_keep_hi(uint32_t bits, uint32_t thresh) { return (bits >> 16) >= thresh; }
static inline Dropout make_dropout(const uint32_t *seed, uint32_t site, float p) {
    Dropout d{seed, site, 0u, 1.0f};
    if (seed && p > 0.0f) {
        uint32_t t = (uint32_t)(p * 65536.0f + 0.5f);
        if (t > 65535u) t = 65535u;
        d.thresh = t;
        d.scale = 65536.0f / (float)(65536u - t);
    } else {
        d.seed = nullptr;
    }
    return d;
}

// Chan et al.: merge (n_b, mean_b, M2_b) into (n_a, mean_a, M2_a)
__device__ __forceinline__ void chan_merge(float &n_a, float &mean_a, float &m2_a, float n_b, float mean_b, float m2_b) {
    const float n = n_a + n_b;
    const float delta = mean_b - mean_a;
    const float f = n_b / n;
    mean_a = fmaf(delta, f, mean_a);
    m2_a = m2_a + m2_b + delta * delta * n_a * f;
    n_a = n;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// Inside the replayed CUDA graph every kernel waited for its predecessor to drain completely and then paid its own launch
// latency + prologue (barrier init, TMEM allocation, tensor-map prefetch): ~1 ms of the 11.2 ms inference step over 123
// launches.  The hot kernels call pdl_launch_dependents() first thing (the next kernel may be scheduled onto SMs as they free
// up) and pdl_wait() after their prologue, before the first access to memory the predecessor wrote; both are no-ops when
// the launch did not carry the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        unsigned cluster_x, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster_x > 0) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl && pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace kbner
