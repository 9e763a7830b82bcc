// Host-side helpers shared by the C-ABI entry points (not part of the public header).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// status codes returned by every gridmm_* entry point (0 = ok; positive values are cudaError_t)
#define GRIDMM_ERR_SHAPE (-1)     // unsupported size / alignment
#define GRIDMM_ERR_DRIVER (-2)    // cuTensorMapEncodeTiled unavailable or failed
#define GRIDMM_ERR_ARG (-3)       // null pointer / inconsistent arguments

// fp16 row-major 2D tensor map: dims {inner, outer}, SWIZZLE_128B, box {box_inner, box_outer}.
int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_pitch_bytes,
                     uint32_t box_inner, uint32_t box_outer);

// counts kernel launches issued through the C ABI (bench.py reports it as gpu_launches)
void gridmm_count_launch(int n);
