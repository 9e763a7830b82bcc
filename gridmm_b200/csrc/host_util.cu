#include "host_util.h"

#include <atomic>
#include <cstdlib>

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        // resolved through the runtime so the library has no link-time dependency on libcuda
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

std::atomic<long long> g_launches{0};
}  // namespace

int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_pitch_bytes,
                     uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return GRIDMM_ERR_DRIVER;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_pitch_bytes & 15)) return GRIDMM_ERR_SHAPE;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : GRIDMM_ERR_DRIVER;
}

bool gridmm_use_pdl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GRIDMM_PDL");      // on by default (round 2: 1.121 -> 1.088 ms per navigation step); GRIDMM_PDL=0 turns it off
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

int gridmm_sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
    if (dev < 64) {
        const int c = cache[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (dev < 64) cache[dev].store(sms, std::memory_order_relaxed);
    return sms;
}

void gridmm_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" long long gridmm_launch_count() { return g_launches.load(std::memory_order_relaxed); }
extern "C" void gridmm_launch_count_reset() { g_launches.store(0, std::memory_order_relaxed); }
extern "C" int gridmm_abi_version() { return 4; }
