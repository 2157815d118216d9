// Helpers shared by the CRF kernels (crf.cu, crf_viterbi.cu): cp.async ring copies, explicit shared-window loads /
// stores, compile-time loops, 3-input max trees.
#pragma once
#include <utility>

#include "common.cuh"

namespace kbner {

__device__ __forceinline__ void cp_async16(uint32_t dst_s, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst_s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst_s, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst_s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// explicit shared-window addresses: keeps the per-step loads / stores to one LDS / STS with an immediate offset
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts_b32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) {      // compile-time loop: the index is a constant expression
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

template <int N>
__device__ __forceinline__ float max_tree(const float (&c)[N]) {
    float t[(N + 2) / 3];
#pragma unroll
    for (int i = 0; i < (N + 2) / 3; ++i) {
        const int a = 3 * i, b = (3 * i + 1 < N) ? 3 * i + 1 : a, d = (3 * i + 2 < N) ? 3 * i + 2 : a;
        t[i] = (b == a) ? c[a] : ((d == a) ? fmaxf(c[a], c[b]) : fmax3(c[a], c[b], c[d]));
    }
    if constexpr ((N + 2) / 3 == 1) return t[0];
    else return max_tree<(N + 2) / 3>(t);
}

}  // namespace kbner
