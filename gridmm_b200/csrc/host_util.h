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

// SM count of the CURRENT device (cached per device ordinal); <= 0 on error
int gridmm_sm_count();

// counts kernel launches issued through the C ABI (bench.py reports it as gpu_launches)
void gridmm_count_launch(int n);

// Programmatic dependent launch (PDL): every kernel of this library calls griddepcontrol.wait before its first global-memory
// access and griddepcontrol.launch_dependents once its loads are issued (the GEMMs) or at its end (everything else), and is
// launched with programmatic stream serialization, so the launch latency and prologue (barrier init, TMEM allocation,
// descriptor prefetch) of kernel N+1 overlap the tail of kernel N -- inside the step's CUDA graph too.  On by default
// (measured on the B = 32, T = 8 navigation step: 1.121 -> 1.088 ms); GRIDMM_PDL=0 launches without the attribute, the
// device-side instructions are then no-ops.  (Round 1 measured PDL 4 % slower: every kernel then released its dependents at
// its START, and early dependents competed for SM slots.)
bool gridmm_use_pdl();

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gridmm_use_pdl() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif
