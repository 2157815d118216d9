// Host-side TMA tensor-map construction (driver entry point resolved at run time: the library
// does not link against libcuda) with a small cache keyed by (pointer, shape, box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace kbner {

// 2-D row-major tensor [rows][cols] of `elem_bytes`-byte elements (2 = bf16, 4 = fp32) with leading dimension ld
// (elements); box = [box_rows][box_cols] with box_cols * elem_bytes == 128 (one SWIZZLE_128B span).  0 on success.
int make_tmap_2d(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols, uint32_t elem_bytes);

inline int make_tmap_bf16_2d(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_rows, uint32_t box_cols) {
    return make_tmap_2d(out, base, rows, cols, ld, box_rows, box_cols, 2);
}

}  // namespace kbner
