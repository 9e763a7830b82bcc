// Action heads of the navigation step (map_nav_src/models/vilmodel.py:859-907) in three launches:
//   head_rows_kernel   fp32 rows of gmap' / vp / map -> one [hi | lo | hi] fp16 operand matrix for ALL ClsPrediction heads
//   gridmm_cls_heads_f16 (gemm_tc.cu)  grouped tcgen05 GEMM, epilogue reduces ReLU(xW+b) to three sums per row and 64 columns
//                      (sap_fuse_linear's first layer rides along as two raw-product row tiles: [gmap'_0 ; vp_0] . [Wg | Wv]^T)
//   nav_logits2_kernel finishes every head (LayerNorm + Linear(768 -> 1) from the sums) and fuses the logits
// ClsPrediction (vilmodel.py:663-674) = Linear, ReLU, LayerNorm(eps 1e-12), Linear(768 -> 1):
//   logit = sum_n ((r_n - mean) * rstd * gamma_n + beta_n) * w2_n + b2 = rstd * (S3 - mean * c1) + c0
//   with r = ReLU(xW + b), S1 = sum r, S2 = sum r^2, S3 = sum r * gamma * w2, c1 = sum gamma * w2, c0 = sum beta * w2 + b2.
#include "common.cuh"
#include "host_util.h"

namespace gmm {

constexpr int HD = 768;

struct HeadSeg {
    const float* x; int ldx, in_rows_per_b, in_off, rows_per_b, rows, out_row0;
    const int* row_off;      // optional [batch] first input row of every episode (ragged / packed source), replaces b * in_rows_per_b
};
struct HeadRowsParams {
    HeadSeg seg[6];
    int nseg, total;
    __half* out; int ld_out;      // [rows, 3 * 768]
};

__global__ void __launch_bounds__(256) head_rows_kernel(HeadRowsParams p) {
    pdl_wait();
    int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= p.total) return;
    int si = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i)
        if (si < p.nseg - 1 && row >= p.seg[si].rows) { row -= p.seg[si].rows; ++si; }
    const HeadSeg sg = p.seg[si];
    const int b = row / sg.rows_per_b, r = row - b * sg.rows_per_b;
    const size_t irow = (sg.row_off ? static_cast<size_t>(sg.row_off[b]) : static_cast<size_t>(b) * sg.in_rows_per_b) + sg.in_off + r;
    __half* o = p.out + static_cast<size_t>(sg.out_row0 + row) * p.ld_out;
#pragma unroll
    for (int i = 0; i < HD / 128; ++i) {
        const int col = (i * 32 + lane) * 4;
        const float4 v = *reinterpret_cast<const float4*>(sg.x + irow * sg.ldx + col);
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
        uint2 hi, lo;
        hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
        lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(o + col) = hi;
        *reinterpret_cast<uint2*>(o + HD + col) = lo;
        *reinterpret_cast<uint2*>(o + 2 * HD + col) = hi;
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// vilmodel.py:859-907.  One CTA per episode; every head's logit is finished from its partial sums first.
struct Logit2Params {
    const float* part;         // [rows][12][3] grouped-GEMM partial sums
    const float* fuse_raw;     // [2][B_pad][768]: gmap'_0 . Wg^T and vp_0 . Wv^T (rows row_fuse_g + b, row_fuse_v + b of raw), or null (fuse weight 0.5)
    const float* fuse_bias; const float* fuse_gw2;    // sap_fuse_linear.net[0].bias, gamma * w2
    int row_fuse_g, row_fuse_v;
    const float* consts;       // [5][2] (c1, c0) of global, local, grid, obj, fuse
    int row_global, row_local, row_grid, row_obj;      // first partial row of each head (row_obj < 0: no object head)
    const uint8_t* gmap_masks; const uint8_t* gmap_visited;   // [B, G]
    const uint8_t* vp_nav_masks; const uint8_t* vp_obj_masks; // [B, V]
    const int* fuse_src;       // [B, G]  >=0: add local[src]; -2: add the back-track sum; -1: nothing
    const uint8_t* bw_mask;    // [B, V]  candidates that are already visited (their local logits are summed); unused with cand_node
    // mask-free form of the same tables (the host then never has to read gmap_visited back from the device): fuse_src[b, g] is
    // computed from the vpid strings alone (>= 0: LAST candidate slot holding node g's viewpoint, -2: none, -1: g is [stop] or
    // padding) and only applies to nodes that are not visited; cand_node[b, v] = gmap slot of candidate v's viewpoint (-1: none,
    // [stop], padding) -- a candidate is "already visited" when that node's gmap_visited flag is set (vilmodel.py:884-891)
    const int* cand_node;      // [B, V] or null
    float* global_logits; float* grid_logits; float* local_logits; float* fused_logits; float* obj_logits;
    int G, V;
};

__device__ __forceinline__ float finish_head(const float* pr, int nparts, float c1, float c0) {
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int i = 0; i < nparts; ++i) { s1 += pr[3 * i]; s2 += pr[3 * i + 1]; s3 += pr[3 * i + 2]; }
    const float mean = s1 * (1.0f / HD);
    const float var = fmaxf(s2 * (1.0f / HD) - mean * mean, 0.0f);
    return rsqrtf(var + 1e-12f) * (s3 - mean * c1) + c0;
}

__global__ void __launch_bounds__(128) nav_logits2_kernel(Logit2Params p) {
    extern __shared__ float s_local[];
    __shared__ float s_bw, s_fw;
    pdl_wait();
    const int b = blockIdx.x, tid = threadIdx.x;
    const float ninf = -INFINITY;
    if (p.fuse_raw) {
        // sap_fuse_linear (vilmodel.py:859-862): r = ReLU(gmap'_0 Wg^T + vp_0 Wv^T + b) over 768 columns -> the three sums -> sigmoid
        const float* hg = p.fuse_raw + static_cast<size_t>(p.row_fuse_g + b) * HD;
        const float* hv = p.fuse_raw + static_cast<size_t>(p.row_fuse_v + b) * HD;
        float s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int n = tid; n < HD; n += blockDim.x) {
            const float r = fmaxf(hg[n] + hv[n] + p.fuse_bias[n], 0.0f);
            s1 += r; s2 = fmaf(r, r, s2); s3 = fmaf(r, p.fuse_gw2[n], s3);
        }
        s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
        __shared__ float s_red[3][4];
        if ((tid & 31) == 0) { s_red[0][tid >> 5] = s1; s_red[1][tid >> 5] = s2; s_red[2][tid >> 5] = s3; }
        __syncthreads();
        if (tid == 0) {
            float pr[3] = {0.f, 0.f, 0.f};
            for (int w = 0; w < 4; ++w) { pr[0] += s_red[0][w]; pr[1] += s_red[1][w]; pr[2] += s_red[2][w]; }
            s_fw = 1.0f / (1.0f + expf(-finish_head(pr, 1, p.consts[8], p.consts[9])));
        }
    } else if (tid == 0) {
        s_fw = 0.5f;
    }
    __syncthreads();
    const float fw = s_fw;
    for (int v = tid; v < p.V; v += blockDim.x) {
        const float raw = finish_head(p.part + static_cast<size_t>(p.row_local + b * p.V + v) * 36, 12, p.consts[2], p.consts[3]);
        float l = raw * (1.0f - fw);
        if (!p.vp_nav_masks[b * p.V + v]) l = ninf;
        s_local[v] = l;
        p.local_logits[b * p.V + v] = l;
        if (p.row_obj >= 0) {
            float o = finish_head(p.part + static_cast<size_t>(p.row_obj + b * p.V + v) * 36, 12, p.consts[6], p.consts[7]);
            if (!p.vp_obj_masks[b * p.V + v]) o = ninf;
            p.obj_logits[b * p.V + v] = o;
        }
    }
    __syncthreads();
    if (tid == 0) {
        float bw = 0.0f;   // sequential, candidate order (the reference accumulates with `+=` in a Python loop)
        for (int v = 1; v < p.V; ++v) {
            bool back;
            if (p.cand_node) {
                const int node = p.cand_node[b * p.V + v];
                back = node >= 0 && p.gmap_visited[b * p.G + node] != 0;
            } else {
                back = p.bw_mask[b * p.V + v] != 0;
            }
            if (back) bw += s_local[v];
        }
        s_bw = bw;
    }
    __syncthreads();
    for (int g = tid; g < p.G; g += blockDim.x) {
        const bool visited = p.gmap_visited[b * p.G + g] != 0;
        const bool masked = visited || !p.gmap_masks[b * p.G + g];
        float gl = finish_head(p.part + static_cast<size_t>(p.row_global + b * p.G + g) * 36, 12, p.consts[0], p.consts[1]) * fw;
        float gr = finish_head(p.part + static_cast<size_t>(p.row_grid + b * p.G + g) * 36, 12, p.consts[4], p.consts[5]);
        if (masked) { gl = ninf; gr = ninf; }
        p.global_logits[b * p.G + g] = gl;
        p.grid_logits[b * p.G + g] = gr;
        float f = gl;
        if (g == 0) f += s_local[0];
        else {
            int src = p.fuse_src[b * p.G + g];
            if (p.cand_node && visited) src = -1;           // `vp not in visited_nodes` (vilmodel.py:893)
            if (src >= 0) f += s_local[src];
            else if (src == -2) f += s_bw;
        }
        p.fused_logits[b * p.G + g] = f;
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

}  // namespace gmm

// segs: HOST array of nseg x 7 ints is awkward across a C ABI with pointers inside, so the (at most 4) segments are passed flat:
// x[i], ldx[i], in_rows_per_b[i], in_off[i], rows_per_b[i], out_row0[i]; rows = rows_per_b * batch.
extern "C" int gridmm_head_rows(int nseg, const float* const* x, const int* ldx, const int* in_rows_per_b, const int* in_off,
                                const int* rows_per_b, const int* out_row0, const int* const* row_off, int batch, void* out_f16,
                                int ld_f16, int hidden, cudaStream_t stream) {
    using namespace gmm;
    if (nseg < 1 || nseg > 6 || batch <= 0) return GRIDMM_ERR_SHAPE;
    if (hidden != HD || (ld_f16 % 4) || ld_f16 < 3 * HD || !out_f16) return GRIDMM_ERR_SHAPE;
    HeadRowsParams p;
    p.nseg = nseg; p.total = 0; p.out = reinterpret_cast<__half*>(out_f16); p.ld_out = ld_f16;
    for (int i = 0; i < nseg; ++i) {
        if (!x[i] || (ldx[i] % 4)) return GRIDMM_ERR_ARG;
        p.seg[i] = HeadSeg{x[i], ldx[i], in_rows_per_b[i], in_off[i], rows_per_b[i], rows_per_b[i] * batch, out_row0[i],
                           row_off ? row_off[i] : nullptr};
        p.total += rows_per_b[i] * batch;
    }
    for (int i = nseg; i < 6; ++i) p.seg[i] = p.seg[nseg - 1];
    if (p.total <= 0) return 0;
    GMM_CUDA_CHECK(launch_pdl(head_rows_kernel, dim3((p.total + 7) / 8), dim3(256), 0, stream, p));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_nav_logits2(const float* part, const float* fuse_raw, const float* fuse_bias, const float* fuse_gw2,
                                  int row_fuse_g, int row_fuse_v, const float* consts, int row_global, int row_local,
                                  int row_grid, int row_obj, const unsigned char* gmap_masks, const unsigned char* gmap_visited,
                                  const unsigned char* vp_nav_masks, const unsigned char* vp_obj_masks, const int* fuse_src,
                                  const unsigned char* bw_mask, const int* cand_node, float* global_logits, float* grid_logits,
                                  float* local_logits, float* fused_logits, float* obj_logits, int batch, int G, int V,
                                  cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (!part || !consts || !gmap_masks || !gmap_visited || !vp_nav_masks || !fuse_src || (!bw_mask && !cand_node) || !global_logits ||
        !grid_logits || !local_logits || !fused_logits || (row_obj >= 0 && (!vp_obj_masks || !obj_logits)))
        return GRIDMM_ERR_ARG;
    if (fuse_raw && (!fuse_bias || !fuse_gw2)) return GRIDMM_ERR_ARG;
    Logit2Params p{part, fuse_raw, fuse_bias, fuse_gw2, row_fuse_g, row_fuse_v, consts, row_global, row_local, row_grid, row_obj, gmap_masks, gmap_visited, vp_nav_masks,
                   vp_obj_masks, fuse_src, bw_mask, cand_node, global_logits, grid_logits, local_logits, fused_logits, obj_logits, G, V};
    GMM_CUDA_CHECK(launch_pdl(nav_logits2_kernel, dim3(batch), dim3(128), V * sizeof(float), stream, p));
    gridmm_count_launch(1);
    return 0;
}
