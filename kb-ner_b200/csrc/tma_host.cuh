// Host-side TMA tensor-map construction (driver entry point resolved at run time: the library
// does not link against libcuda) with a small cache keyed by (pointer, shape, box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace kbner {

// 2-D row-major bf16 tensor [rows][cols] with leading dimension ld (elements); box = [box_rows][box_cols]
// with box_cols * 2 bytes == 128 (one SWIZZLE_128B span).  Returns 0 on success.
int make_tmap_bf16_2d(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);

}  // namespace kbner
