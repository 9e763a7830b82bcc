// Row-wise kernels around the GEMMs of the navigation step (all hidden size 768, one warp per row):
//   layer norm (BertLayerNorm eps 1e-12 / nn.LayerNorm eps 1e-5), the small position-feature embeddings
//   (Linear(5|7|14 -> 768) + LayerNorm, map_nav_src/models/vilmodel.py:697-700, 563-566, 534-537),
//   grid-cell assembly incl. the reference's mask-compaction quirk (vilmodel.py:813-823) together with the gmap tokens and
//   grid_encoder's first pre-norm LayerNorm (gridmm_map_inputs, vilmodel.py:828-838), the packed fusion context
//   (gridmm_kv_index + gridmm_fusion_inputs: [map; txt] without its masked rows, queries [gmap'; vp], vilmodel.py:843-850),
//   text / panorama input embeddings, ClsPrediction tails and the action-logit fusion (vilmodel.py:859-907).
#include "common.cuh"
#include "host_util.h"

namespace gmm {

constexpr int HID = 768;
constexpr int HV = HID / 128;   // float4 per lane

__device__ __forceinline__ void ln_row(float4 (&x)[HV], const float* gamma, const float* beta, float eps, int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HV; ++i) s += x[i].x + x[i].y + x[i].z + x[i].w;
    const float mean = warp_sum(s) * (1.0f / HID);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
        v += a * a + b * b + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(v) * (1.0f / HID) + eps);
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const int col = (i * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(gamma + col);
        const float4 bb = *reinterpret_cast<const float4*>(beta + col);
        x[i].x = (x[i].x - mean) * rstd * g.x + bb.x;
        x[i].y = (x[i].y - mean) * rstd * g.y + bb.y;
        x[i].z = (x[i].z - mean) * rstd * g.z + bb.z;
        x[i].w = (x[i].w - mean) * rstd * g.w + bb.w;
    }
}

__device__ __forceinline__ void store_row(float4 (&x)[HV], float* o32, __half* o16, int lane) {
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const int col = (i * 32 + lane) * 4;
        if (o32) *reinterpret_cast<float4*>(o32 + col) = x[i];
        if (o16) {
            const __half2 a = __floats2half2_rn(x[i].x, x[i].y), b = __floats2half2_rn(x[i].z, x[i].w);
            uint2 u;
            u.x = *reinterpret_cast<const uint32_t*>(&a);
            u.y = *reinterpret_cast<const uint32_t*>(&b);
            *reinterpret_cast<uint2*>(o16 + col) = u;
        }
    }
}

// v = bias + f[0..kin) . W^T[kin, 768] for this lane's 6 float4 column groups (position-feature embeddings: kin <= 16).  The
// weight loads of TWO column groups (2 x 16 predicated LDG.128) are issued before their FMAs, so a row costs three global round
// trips instead of six (one warp per row: these kernels are pure latency).
__device__ __forceinline__ void small_linear(float4 (&v)[HV], const float (&f)[16], int kin, const float* w, const float* bias, int lane) {
#pragma unroll
    for (int i0 = 0; i0 < HV; i0 += 2) {
        float4 w4[2][16];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = ((i0 + j) * 32 + lane) * 4;
#pragma unroll
            for (int k = 0; k < 16; ++k)
                w4[j][k] = (k < kin) ? __ldg(reinterpret_cast<const float4*>(w + static_cast<size_t>(k) * HID + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float4 a = *reinterpret_cast<const float4*>(bias + ((i0 + j) * 32 + lane) * 4);
#pragma unroll
            for (int k = 0; k < 16; ++k) {         // f[k] = 0 beyond kin
                a.x = fmaf(f[k], w4[j][k].x, a.x); a.y = fmaf(f[k], w4[j][k].y, a.y); a.z = fmaf(f[k], w4[j][k].z, a.z); a.w = fmaf(f[k], w4[j][k].w, a.w);
            }
            v[i0 + j] = a;
        }
    }
}

// ------------------------------------------------------------------------------------------- layer norm
__global__ void __launch_bounds__(256) layernorm_kernel(const float* x, int ldx, const float* gamma, const float* beta,
                                                        float eps, float* o32, int ld32, __half* o16, int ld16, int rows) {
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4 v[HV];
#pragma unroll
    for (int i = 0; i < HV; ++i) v[i] = *reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * ldx + (i * 32 + lane) * 4);
    ln_row(v, gamma, beta, eps, lane);
    store_row(v, o32 ? o32 + static_cast<size_t>(row) * ld32 : nullptr, o16 ? o16 + static_cast<size_t>(row) * ld16 : nullptr, lane);
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// ------------------------------------------------------------------------------------------- row copies / casts
// out[b, out_off + r] = in[b, in_off + r] for r < rows_per_b  (fp32 in; fp32 and/or fp16 out): builds the
// [map; txt] and [gmap; vp] sequences (vilmodel.py:843-850) and the head inputs without torch.cat.
__global__ void __launch_bounds__(256) copy_rows_kernel(const float* x, int ldx, int in_rows_per_b, int in_off, float* o32,
                                                        int ld32, __half* o16, int ld16, int out_rows_per_b, int out_off,
                                                        int rows_per_b, int rows) {
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = row / rows_per_b, r = row - b * rows_per_b;
    const size_t irow = static_cast<size_t>(b) * in_rows_per_b + in_off + r;
    const size_t orow = static_cast<size_t>(b) * out_rows_per_b + out_off + r;
    float4 v[HV];
#pragma unroll
    for (int i = 0; i < HV; ++i) v[i] = *reinterpret_cast<const float4*>(x + irow * ldx + (i * 32 + lane) * 4);
    store_row(v, o32 ? o32 + orow * ld32 : nullptr, o16 ? o16 + orow * ld16 : nullptr, lane);
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// ------------------------------------------------------------------------------------------- fusion-encoder inputs
// vilmodel.py:843-850 in one launch: queries x = [gmap' ; vp] get their gmap' rows from the map sequence (fp32 + fp16), the
// context kv = [map ; txt] is assembled as fp16, and both masks are concatenated.  One warp per row of (kv rows, then gmap' rows).
struct FusionInParams {
    const float* map32; const float* txt32;     // [B, S, 768], [B, L, 768]
    const uint8_t* map_mask; const uint8_t* txt_mask; const uint8_t* gmap_mask; const uint8_t* vp_mask;   // [B,S] [B,L] [B,G] [B,V]
    float* x32; __half* x16; __half* kv16;      // [B, Q, 768] x2, [B, KC, 768]
    uint8_t* kv_mask; uint8_t* q_mask;          // [B, KC], [B, Q]
    int B, S, L, G, V;
    const int* kv_pos;                          // [B, KC] packed row of every VALID context row (-1: masked), or null = unpacked
    // optional: the vp tokens of x (rows G.. of every episode) = vp_img + LN(Linear(vp_pos)) (vilmodel.py:832-833), else they must
    // already be in x32 / x16
    const float* v_feat; int v_kin; const float* v_w; const float* v_bias; const float* v_gamma; const float* v_beta; const float* v_base;
};

// Packed ("ragged") context index for the fusion encoder: masked context rows (empty grid-cell slots, padded text) get no K/V
// projection and no attention work.  kv_off[b] = number of valid rows of the episodes before b,
// kv_cnt[b] = valid rows of b, kv_pos[b, r] = kv_off[b] + rank of row r among b's valid rows (or -1), kv_off[B] = total.
__global__ void __launch_bounds__(1024) kv_index_kernel(const uint8_t* map_mask, const uint8_t* txt_mask, int S, int L, int B,
                                                        int* kv_pos, int* kv_off, int* kv_cnt) {
    // ONE CTA, one warp per episode (round-robin): count with ballots, scan the counts, then write the packed positions
    extern __shared__ int s_cnt[];          // [B + 1]
    pdl_wait();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int KC = S + L;
    for (int b = warp; b < B; b += nwarps) {
        int cnt = 0;
        for (int r0 = 0; r0 < KC; r0 += 32) {
            const int r = r0 + lane;
            const bool valid = r < KC && ((r < S) ? map_mask[b * S + r] : txt_mask[b * L + (r - S)]) != 0;
            cnt += __popc(__ballot_sync(0xffffffffu, valid));
        }
        if (lane == 0) s_cnt[b] = cnt;
    }
    __syncthreads();
    if (warp == 0) {                        // exclusive scan of the counts (B <= a few hundred)
        int run = 0;
        for (int b0 = 0; b0 < B; b0 += 32) {
            const int b = b0 + lane;
            const int c = b < B ? s_cnt[b] : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (b < B) { kv_off[b] = run + incl - c; kv_cnt[b] = c; s_cnt[b] = run + incl - c; }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) kv_off[B] = run;
    }
    __syncthreads();
    for (int b = warp; b < B; b += nwarps) {
        int rank = s_cnt[b];
        for (int r0 = 0; r0 < KC; r0 += 32) {
            const int r = r0 + lane;
            const bool valid = r < KC && ((r < S) ? map_mask[b * S + r] : txt_mask[b * L + (r - S)]) != 0;
            const unsigned bal = __ballot_sync(0xffffffffu, valid);
            if (r < KC) kv_pos[b * KC + r] = valid ? rank + __popc(bal & ((1u << lane) - 1u)) : -1;
            rank += __popc(bal);
        }
    }
    pdl_launch_dependents();
}

__global__ void __launch_bounds__(256) fusion_inputs_kernel(FusionInParams p) {
    pdl_wait();
    const int KC = p.S + p.L, Q = p.G + p.V;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= p.B * (KC + p.G + (p.v_feat ? p.V : 0))) return;
    float4 v[HV];
    if (row >= p.B * (KC + p.G)) {
        // vp token j of episode b
        const int vr = row - p.B * (KC + p.G);
        const int b = vr / p.V, j = vr - b * p.V;
        float f[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) f[k] = (k < p.v_kin) ? p.v_feat[static_cast<size_t>(vr) * p.v_kin + k] : 0.0f;
        small_linear(v, f, p.v_kin, p.v_w, p.v_bias, lane);
        ln_row(v, p.v_gamma, p.v_beta, 1e-12f, lane);
#pragma unroll
        for (int i = 0; i < HV; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(p.v_base + static_cast<size_t>(vr) * HID + (i * 32 + lane) * 4);
            v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
        }
        const size_t orow = static_cast<size_t>(b) * Q + p.G + j;
        store_row(v, p.x32 + orow * HID, p.x16 + orow * HID, lane);
    } else if (row < p.B * KC) {
        const int b = row / KC, r = row - b * KC;
        const float* src = (r < p.S) ? p.map32 + (static_cast<size_t>(b) * p.S + r) * HID
                                     : p.txt32 + (static_cast<size_t>(b) * p.L + (r - p.S)) * HID;
#pragma unroll
        for (int i = 0; i < HV; ++i) v[i] = *reinterpret_cast<const float4*>(src + (i * 32 + lane) * 4);
        const int prow = p.kv_pos ? p.kv_pos[row] : row;
        if (prow >= 0) store_row(v, nullptr, p.kv16 + static_cast<size_t>(prow) * HID, lane);
        if (lane == 0) {
            p.kv_mask[row] = (r < p.S) ? p.map_mask[b * p.S + r] : p.txt_mask[b * p.L + (r - p.S)];
            for (int j = r; j < p.V; j += KC) p.q_mask[b * Q + p.G + j] = p.vp_mask[b * p.V + j];
        }
    } else {
        const int rr = row - p.B * KC;
        const int b = rr / p.G, g = rr - b * p.G;
        const float* src = p.map32 + (static_cast<size_t>(b) * p.S + (p.S - p.G) + g) * HID;
#pragma unroll
        for (int i = 0; i < HV; ++i) v[i] = *reinterpret_cast<const float4*>(src + (i * 32 + lane) * 4);
        const size_t orow = static_cast<size_t>(b) * Q + g;
        store_row(v, p.x32 + orow * HID, p.x16 + orow * HID, lane);
        if (lane == 0) p.q_mask[b * Q + g] = p.gmap_mask[b * p.G + g];
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// ------------------------------------------------------------------------------------------- batched copies
// One launch that copies up to 24 independent byte ranges (device or pinned-host sources -> device destinations): the
// per-step inputs of forward('navigation') arrive as ~14 freshly allocated tensors (r2r/agent.py:163-205) and have to land in
// the static buffers a CUDA graph replays from; 14 cudaMemcpyAsync / copy kernels cost more host and device time than the data.
constexpr int SEG_MAX = 24;
struct CopySegs {
    const uint8_t* src[SEG_MAX];
    uint8_t* dst[SEG_MAX];
    long long units[SEG_MAX + 1];      // exclusive prefix of ceil(bytes / 16)
    long long bytes[SEG_MAX];
    int n;
};

__global__ void __launch_bounds__(256) copy_segments_kernel(CopySegs p) {
    pdl_wait();
    const long long total = p.units[p.n];
    for (long long u = blockIdx.x * 256LL + threadIdx.x; u < total; u += static_cast<long long>(gridDim.x) * 256) {
        int s = 0;
        while (u >= p.units[s + 1]) ++s;
        const long long off = (u - p.units[s]) * 16;
        const uint8_t* src = p.src[s] + off;
        uint8_t* dst = p.dst[s] + off;
        const long long left = p.bytes[s] - off;
        if (left >= 16 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
        } else {
            const int nb = left < 16 ? static_cast<int>(left) : 16;
            for (int i = 0; i < nb; ++i) dst[i] = src[i];
        }
    }
    pdl_launch_dependents();
}

// ------------------------------------------------------------------------------------------- fp32 -> (hi, lo) fp16 split
// out[b, r] = [ hi | lo | hi ] at column blocks 0, k_total, 2*k_total (each 768 wide at the given column offset), where
// hi = fp16(x), lo = fp16(x - hi).  Against weights laid out [Wh | Wh | Wl] a single K-concatenated GEMM then computes
// xh.Wh + xl.Wh + xh.Wl = x.W to ~2^-22 -- used for the tiny ClsPrediction GEMMs, whose fp16 rounding would otherwise
// dominate the logit error (numerics).
__global__ void __launch_bounds__(256) split_rows_kernel(const float* x, int ldx, int in_rows_per_b, int in_off, __half* o16,
                                                         int ld16, int k_total, int rows_per_b, int rows) {
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = row / rows_per_b, r = row - b * rows_per_b;
    const size_t irow = static_cast<size_t>(b) * in_rows_per_b + in_off + r;
    __half* o = o16 + static_cast<size_t>(row) * ld16;
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const int col = (i * 32 + lane) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + irow * ldx + col);
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
        uint2 hi, lo;
        hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
        lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(o + col) = hi;
        *reinterpret_cast<uint2*>(o + k_total + col) = lo;
        *reinterpret_cast<uint2*>(o + 2 * k_total + col) = hi;
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// ------------------------------------------------------------------------------------------- position embeddings
// out[b, off + r] = base[b, r] + table[idx[b, r]] + LN(W f[b, r] + bias)      (base / table optional)
struct EmbedParams {
    const float* feat; int kin;                 // [rows, kin]
    const float* w; const float* bias;          // TRANSPOSED weight [kin, 768], [768]
    const float* gamma; const float* beta; float eps;
    const float* base;                          // [rows, 768] or null
    const float* table; const long long* idx;   // [n, 768], [rows] or null
    float* o32; __half* o16;                    // [B, out_rows_per_b, 768]
    int in_rows_per_b, out_rows_per_b, out_row_off, rows;
};

__global__ void __launch_bounds__(256) embed_kernel(EmbedParams p) {
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= p.rows) return;
    float f[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) f[k] = (k < p.kin) ? p.feat[static_cast<size_t>(row) * p.kin + k] : 0.0f;
    float4 v[HV];
    small_linear(v, f, p.kin, p.w, p.bias, lane);      // transposed weight [kin, 768]: coalesced float4 per lane
    ln_row(v, p.gamma, p.beta, p.eps, lane);
    if (p.base) {
#pragma unroll
        for (int i = 0; i < HV; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(p.base + static_cast<size_t>(row) * HID + (i * 32 + lane) * 4);
            v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
        }
    }
    if (p.table) {
        const float* tr = p.table + static_cast<size_t>(p.idx[row]) * HID;
#pragma unroll
        for (int i = 0; i < HV; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(tr + (i * 32 + lane) * 4);
            v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
        }
    }
    const int b = row / p.in_rows_per_b, r = row - b * p.in_rows_per_b;
    const size_t orow = static_cast<size_t>(b) * p.out_rows_per_b + p.out_row_off + r;
    store_row(v, p.o32 ? p.o32 + orow * HID : nullptr, p.o16 ? p.o16 + orow * HID : nullptr, lane);
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// ------------------------------------------------------------------------------------------- BERT text embeddings
// LayerNorm(word[ids] + position[l] + token_type[0])   (BertEmbeddings.forward, vilmodel.py:77-93; one warp per token)
__global__ void __launch_bounds__(256) text_embed_kernel(const long long* ids, const float* word, const float* pos,
                                                         const float* type0, const float* gamma, const float* beta, float eps,
                                                         float* o32, __half* o16, int L, int rows) {
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* w = word + static_cast<size_t>(ids[row]) * HID;
    const float* pe = pos + static_cast<size_t>(row % L) * HID;
    float4 v[HV];
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const int col = (i * 32 + lane) * 4;
        const float4 a = *reinterpret_cast<const float4*>(w + col), b = *reinterpret_cast<const float4*>(pe + col);
        const float4 c = *reinterpret_cast<const float4*>(type0 + col);
        v[i] = make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
    }
    ln_row(v, gamma, beta, eps, lane);
    store_row(v, o32 ? o32 + static_cast<size_t>(row) * HID : nullptr, o16 ? o16 + static_cast<size_t>(row) * HID : nullptr, lane);
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// ------------------------------------------------------------------------------------------- grid-cell assembly
// Rows [0, n_cells) of every episode's map sequence (vilmodel.py:813-823):
//   r <  k_b : grid_proj(pooled)[b, r] + grid_pos_embeddings(pos_fts[b, cell of rank r])
//   r >= k_b : 0
// and the validity mask WITH the reference's aliasing quirk: grid_mask is a view of grid_masks[b], so after
// `grid_masks[b,:k] = 1` the second `.sum()` is re-read: k' = k + |S n [k, n_cells)| (S = non-empty cells) and the
// row ends as [0,k) u (S n [k,k')), finally truncated to C = max_b k_b columns.
struct AssembleParams {
    const float* proj;        // [B, n_cells, 768] grid_proj output in rank order (rows >= k_b undefined)
    const float* pos_fts;     // [B, n_cells, 5]
    const int* cell_rank;     // [B, n_cells]
    const int* n_nonempty;    // [B]
    const float* w; const float* bias; const float* gamma; const float* beta;   // grid_pos_embeddings (w TRANSPOSED: [5, 768])
    float* map32;             // [B, seq, 768]
    uint8_t* map_mask;        // [B, seq]
    int batch, n_cells, seq;
    // optional (gridmm_map_inputs): rows [n_cells, seq) = gmap tokens (vilmodel.py:828-831: img + step embedding +
    // LN(Linear(pos))) and the first pre-norm LayerNorm of grid_encoder over every row -> map16
    const float* g_feat; int g_kin; const float* g_w; const float* g_bias; const float* g_gamma; const float* g_beta;
    const float* g_base; const float* g_table; const long long* g_idx; const uint8_t* g_mask;
    const float* n_gamma; const float* n_beta; float n_eps; __half* map16;
};

__global__ void __launch_bounds__(256) grid_assemble_kernel(AssembleParams p) {
    __shared__ int s_inv[256];
    __shared__ int s_red[8];
    __shared__ int s_c, s_k2;
    pdl_wait();
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = p.n_nonempty[b];
    // C = max over the batch; k' for this episode
    int mx = 0;
    for (int i = tid; i < p.batch; i += 256) mx = max(mx, p.n_nonempty[i]);
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[warp] = mx;
    int above = 0;
    for (int c = tid; c < p.n_cells; c += 256) {
        const int r = p.cell_rank[b * p.n_cells + c];
        if (r >= 0) { s_inv[r] = c; above += (c >= k) ? 1 : 0; }
    }
    // block-sum of `above`
    for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
    __shared__ int s_sum[8];
    if (lane == 0) s_sum[warp] = above;
    __syncthreads();
    if (tid == 0) {
        int m = 0, s = 0;
        for (int i = 0; i < 8; ++i) { m = max(m, s_red[i]); s += s_sum[i]; }
        s_c = m; s_k2 = k + s;
    }
    __syncthreads();
    const int C = s_c, k2 = s_k2;
    const int n_rows = p.g_feat ? p.seq : p.n_cells;
    const int rows_per_cta = (n_rows + gridDim.x - 1) / gridDim.x;
    const int r_lo = blockIdx.x * rows_per_cta, r_hi = min(r_lo + rows_per_cta, n_rows);
    for (int r = r_lo + warp; r < r_hi; r += 8) {
        float4 v[HV];
        if (r >= p.n_cells) {
            // gmap token r - n_cells of this episode
            const int G = p.seq - p.n_cells;
            const size_t gr = static_cast<size_t>(b) * G + (r - p.n_cells);
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = (j < p.g_kin) ? p.g_feat[gr * p.g_kin + j] : 0.0f;
            small_linear(v, f, p.g_kin, p.g_w, p.g_bias, lane);
            ln_row(v, p.g_gamma, p.g_beta, 1e-12f, lane);
            const float* tr = p.g_table + static_cast<size_t>(p.g_idx[gr]) * HID;
#pragma unroll
            for (int i = 0; i < HV; ++i) {
                const float4 t1 = *reinterpret_cast<const float4*>(p.g_base + gr * HID + (i * 32 + lane) * 4);
                const float4 t2 = *reinterpret_cast<const float4*>(tr + (i * 32 + lane) * 4);
                v[i].x += t1.x + t2.x; v[i].y += t1.y + t2.y; v[i].z += t1.z + t2.z; v[i].w += t1.w + t2.w;
            }
            store_row(v, p.map32 + (static_cast<size_t>(b) * p.seq + r) * HID, nullptr, lane);
            if (lane == 0) p.map_mask[static_cast<size_t>(b) * p.seq + r] = p.g_mask[gr];
            if (p.map16) {
                ln_row(v, p.n_gamma, p.n_beta, p.n_eps, lane);
                store_row(v, nullptr, p.map16 + (static_cast<size_t>(b) * p.seq + r) * HID, lane);
            }
            continue;
        }
        if (r < k) {
            const int cellid = s_inv[r];
            float f[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) f[j] = p.pos_fts[(static_cast<size_t>(b) * p.n_cells + cellid) * 5 + j];
#pragma unroll
            for (int i = 0; i < HV; ++i) {
                const int col = (i * 32 + lane) * 4;
                float4 a = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
                for (int j = 0; j < 5; ++j) {      // transposed weight [5, 768]
                    const float4 w4 = *reinterpret_cast<const float4*>(p.w + j * HID + col);
                    a.x = fmaf(f[j], w4.x, a.x); a.y = fmaf(f[j], w4.y, a.y); a.z = fmaf(f[j], w4.z, a.z); a.w = fmaf(f[j], w4.w, a.w);
                }
                v[i] = a;
            }
            ln_row(v, p.gamma, p.beta, 1e-12f, lane);
#pragma unroll
            for (int i = 0; i < HV; ++i) {
                const float4 t = *reinterpret_cast<const float4*>(p.proj + (static_cast<size_t>(b) * p.n_cells + r) * HID + (i * 32 + lane) * 4);
                v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < HV; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        store_row(v, p.map32 + (static_cast<size_t>(b) * p.seq + r) * HID, nullptr, lane);
        if (lane == 0) {
            const bool in_s = p.cell_rank[b * p.n_cells + r] >= 0;
            const bool valid = (r < C) && ((r < k) || (r < k2 && in_s));
            p.map_mask[static_cast<size_t>(b) * p.seq + r] = valid ? 1 : 0;
        }
        if (p.map16) {
            ln_row(v, p.n_gamma, p.n_beta, p.n_eps, lane);
            store_row(v, nullptr, p.map16 + (static_cast<size_t>(b) * p.seq + r) * HID, lane);
        }
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// ------------------------------------------------------------------------------------------- ragged ("packed") map sequence
// The reference pads every episode's map sequence [grid cells ; gmap nodes] to the batch maximum (vilmodel.py:813-838); here the
// rows that matter are PACKED back to back, so that every map-sized GEMM / attention / LayerNorm launch only touches them:
//   episode b owns rows m_off[b] .. m_off[b+1]-1 =  [ k_b non-empty cells (rank order) | q_b | G gmap nodes ]
// q_b (0 or 1) stands for the z_b zero-vector slots that the reference's mask-aliasing quirk flags valid (grid_assemble above:
// valid = [0,k) u (S n [k,k')) truncated to C).  Those z_b rows are IDENTICAL at every layer (same input, permutation-equivariant
// blocks), so one representative row is computed and, wherever it acts as an attention KEY, its score gets + log(z_b):
// sum_j exp(s_j) v_j over z identical keys = exp(s + log z) v -- exact.  Masked (invalid) cell slots are dropped: as keys they have
// weight 0, as queries nobody reads them.  All G gmap rows stay (the reference returns their outputs, masked or not).
__global__ void __launch_bounds__(1024) map_index_kernel(const int* cell_rank, const int* n_nonempty, int B, int NC, int G,
                                                         int* m_off, int* m_info, float* m_logz, int* cell_of_rank, int* m_goff) {
    extern __shared__ int s_n[];            // [B + 1] rows per episode, then their exclusive prefix
    __shared__ int s_red[32];
    __shared__ int s_c;
    pdl_wait();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    int mx = 0;
    for (int i = tid; i < B; i += blockDim.x) mx = max(mx, n_nonempty[i]);
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int i = 0; i < nwarps; ++i) m = max(m, s_red[i]);
        s_c = m;                            // C = max_b k_b: the reference truncates the cell slots to C columns
    }
    __syncthreads();
    const int C = s_c;
    for (int b = warp; b < B; b += nwarps) {
        const int k = n_nonempty[b];
        int above = 0;
        for (int c = lane; c < NC; c += 32) {
            const int r = cell_rank[b * NC + c];
            if (r >= 0) { cell_of_rank[b * NC + r] = c; above += (c >= k) ? 1 : 0; }
        }
        above = __reduce_add_sync(0xffffffffu, above);
        const int k2 = min(k + above, C);
        int z = 0;
        for (int c = k + lane; c < k2; c += 32) z += (cell_rank[b * NC + c] >= 0) ? 1 : 0;
        z = __reduce_add_sync(0xffffffffu, z);
        if (lane == 0) {
            const int v = k + (z > 0 ? 1 : 0);
            m_info[b] = k; m_info[B + b] = v; m_info[2 * B + b] = v + G; m_info[3 * B + b] = z;      // [4][B]: each row is a contiguous per-episode array
            m_logz[b] = z > 0 ? logf(static_cast<float>(z)) : 0.0f;
            s_n[b] = v + G;
        }
    }
    __syncthreads();
    if (warp == 0) {
        int run = 0;
        for (int b0 = 0; b0 < B; b0 += 32) {
            const int b = b0 + lane;
            const int c = b < B ? s_n[b] : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (b < B) { m_off[b] = run + incl - c; m_goff[b] = run + incl - G; }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) m_off[B] = run;
    }
    pdl_launch_dependents();
}

struct MapPackedParams {
    const float* proj;        // [B, n_cells, 768] grid_proj output in rank order
    const float* pos_fts;     // [B, n_cells, 5]
    const int* cell_of_rank;  // [B, n_cells]
    const int* m_off; const int* m_info; const float* m_logz;
    const float* w; const float* bias; const float* gamma; const float* beta;   // grid_pos_embeddings (w TRANSPOSED: [5, 768])
    const float* g_feat; int g_kin; const float* g_w; const float* g_bias; const float* g_gamma; const float* g_beta;
    const float* g_base; const float* g_table; const long long* g_idx; const uint8_t* g_mask;
    const float* n_gamma; const float* n_beta; float n_eps;
    float* map32; __half* map16;          // packed rows
    uint8_t* kvalid; float* kbias;        // per packed row: valid as a key, additive score bias (log multiplicity)
    int batch, n_cells, G;
};

__global__ void __launch_bounds__(256) map_inputs_packed_kernel(MapPackedParams p) {
    extern __shared__ int s_off[];         // [B + 1]
    pdl_wait();
    for (int i = threadIdx.x; i <= p.batch; i += 256) s_off[i] = p.m_off[i];
    __syncthreads();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= s_off[p.batch]) return;
    int lo = 0, hi = p.batch;              // largest b with m_off[b] <= row
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= row) lo = mid; else hi = mid;
    }
    const int b = lo, r = row - s_off[b];
    const int k = p.m_info[b], vcells = p.m_info[p.batch + b];
    float4 v[HV];
    uint8_t valid = 1;
    float kb = 0.0f;
    if (r >= vcells) {
        // gmap token (vilmodel.py:828-831): gmap_img + step embedding + LN(Linear(gmap_pos))
        const size_t gr = static_cast<size_t>(b) * p.G + (r - vcells);
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = (j < p.g_kin) ? p.g_feat[gr * p.g_kin + j] : 0.0f;
        small_linear(v, f, p.g_kin, p.g_w, p.g_bias, lane);
        ln_row(v, p.g_gamma, p.g_beta, 1e-12f, lane);
        const float* tr = p.g_table + static_cast<size_t>(p.g_idx[gr]) * HID;
#pragma unroll
        for (int i = 0; i < HV; ++i) {
            const float4 t1 = *reinterpret_cast<const float4*>(p.g_base + gr * HID + (i * 32 + lane) * 4);
            const float4 t2 = *reinterpret_cast<const float4*>(tr + (i * 32 + lane) * 4);
            v[i].x += t1.x + t2.x; v[i].y += t1.y + t2.y; v[i].z += t1.z + t2.z; v[i].w += t1.w + t2.w;
        }
        valid = p.g_mask[gr];
    } else if (r < k) {
        // non-empty cell of rank r: grid_proj(pooled) + grid_pos_embeddings(cell-centre features) (vilmodel.py:813-816)
        const int cellid = p.cell_of_rank[b * p.n_cells + r];
        float f[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) f[j] = p.pos_fts[(static_cast<size_t>(b) * p.n_cells + cellid) * 5 + j];
#pragma unroll
        for (int i = 0; i < HV; ++i) {
            const int col = (i * 32 + lane) * 4;
            float4 a = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float4 w4 = *reinterpret_cast<const float4*>(p.w + j * HID + col);
                a.x = fmaf(f[j], w4.x, a.x); a.y = fmaf(f[j], w4.y, a.y); a.z = fmaf(f[j], w4.z, a.z); a.w = fmaf(f[j], w4.w, a.w);
            }
            v[i] = a;
        }
        ln_row(v, p.gamma, p.beta, 1e-12f, lane);
#pragma unroll
        for (int i = 0; i < HV; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(p.proj + (static_cast<size_t>(b) * p.n_cells + r) * HID + (i * 32 + lane) * 4);
            v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
        }
    } else {
        // the representative of the z_b zero-vector slots the compaction quirk flags valid
#pragma unroll
        for (int i = 0; i < HV; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        kb = p.m_logz[b];
    }
    store_row(v, p.map32 + static_cast<size_t>(row) * HID, nullptr, lane);
    if (lane == 0) { p.kvalid[row] = valid; p.kbias[row] = kb; }
    ln_row(v, p.n_gamma, p.n_beta, p.n_eps, lane);
    store_row(v, nullptr, p.map16 + static_cast<size_t>(row) * HID, lane);
    pdl_launch_dependents();
}

// Packed context of the fusion encoder over the PACKED map: context of episode b = its valid map rows + its valid text rows.
//   kv_src[r]  source of packed context row r: >= 0 packed map row, < 0: -1 - (b * L + l) text row
//   kv_bias[r] additive score bias of that key (log multiplicity of the quirk representative, else 0)
__global__ void __launch_bounds__(1024) kv_index_packed_kernel(const int* m_off, const uint8_t* kvalid, const float* kbias,
                                                               const uint8_t* txt_mask, int B, int L, int* kv_src, float* kv_bias,
                                                               int* kv_off, int* kv_cnt) {
    extern __shared__ int s_cnt[];          // [B + 1]
    pdl_wait();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int b = warp; b < B; b += nwarps) {
        const int r0 = m_off[b], n = m_off[b + 1] - r0;
        int cnt = 0;
        for (int i0 = 0; i0 < n + L; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < n + L && ((i < n) ? kvalid[r0 + i] : txt_mask[b * L + (i - n)]) != 0;
            cnt += __popc(__ballot_sync(0xffffffffu, valid));
        }
        if (lane == 0) s_cnt[b] = cnt;
    }
    __syncthreads();
    if (warp == 0) {
        int run = 0;
        for (int b0 = 0; b0 < B; b0 += 32) {
            const int b = b0 + lane;
            const int c = b < B ? s_cnt[b] : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (b < B) { kv_off[b] = run + incl - c; kv_cnt[b] = c; s_cnt[b] = run + incl - c; }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) kv_off[B] = run;
    }
    __syncthreads();
    for (int b = warp; b < B; b += nwarps) {
        const int r0 = m_off[b], n = m_off[b + 1] - r0;
        int rank = s_cnt[b];
        for (int i0 = 0; i0 < n + L; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < n + L && ((i < n) ? kvalid[r0 + i] : txt_mask[b * L + (i - n)]) != 0;
            const unsigned bal = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                const int dst = rank + __popc(bal & ((1u << lane) - 1u));
                kv_src[dst] = (i < n) ? (r0 + i) : (-1 - (b * L + (i - n)));
                kv_bias[dst] = (i < n) ? kbias[r0 + i] : 0.0f;
            }
            rank += __popc(bal);
        }
    }
    pdl_launch_dependents();
}

struct FusionPackedParams {
    const float* map32; const float* txt32;     // packed map rows, [B, L, 768]
    const int* kv_src; const int* kv_off;       // packed context index (kv_off[B] = number of context rows)
    const int* m_goff;                          // [B] first gmap row of every episode in the packed map
    const uint8_t* gmap_mask; const uint8_t* vp_mask;
    float* x32; __half* x16; __half* kv16; uint8_t* q_mask;
    int B, L, G, V, kv_rows_max;
    const float* v_feat; int v_kin; const float* v_w; const float* v_bias; const float* v_gamma; const float* v_beta; const float* v_base;
};

__global__ void __launch_bounds__(256, 4) fusion_inputs_packed_kernel(FusionPackedParams p) {
    pdl_wait();
    const int Q = p.G + p.V;
    // warp per output row; the grid is ordered [vp tokens | gmap tokens | context rows]: the vp rows (a 768 x K linear + LayerNorm
    // each) are the expensive ones and must not be the tail of the launch (as its last blocks they ran alone, one block per SM)
    const int wrow = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (wrow >= p.kv_rows_max + p.B * Q) return;
    const int row = wrow < p.B * p.V ? p.kv_rows_max + p.B * p.G + wrow
                                     : (wrow < p.B * Q ? p.kv_rows_max + (wrow - p.B * p.V) : wrow - p.B * Q);
    float4 v[HV];
    if (row < p.kv_rows_max) {
        const int n_kv = p.kv_off[p.B], src = p.kv_src[row];      // both requested at once (kv_src has kv_rows_max entries)
        if (row >= n_kv) return;
        const float* sp = (src >= 0) ? p.map32 + static_cast<size_t>(src) * HID : p.txt32 + static_cast<size_t>(-1 - src) * HID;
#pragma unroll
        for (int i = 0; i < HV; ++i) v[i] = *reinterpret_cast<const float4*>(sp + (i * 32 + lane) * 4);
        store_row(v, nullptr, p.kv16 + static_cast<size_t>(row) * HID, lane);
    } else if (row < p.kv_rows_max + p.B * p.G) {
        const int rr = row - p.kv_rows_max;
        const int b = rr / p.G, g = rr - b * p.G;
        const float* src = p.map32 + (static_cast<size_t>(p.m_goff[b]) + g) * HID;
#pragma unroll
        for (int i = 0; i < HV; ++i) v[i] = *reinterpret_cast<const float4*>(src + (i * 32 + lane) * 4);
        const size_t orow = static_cast<size_t>(b) * Q + g;
        store_row(v, p.x32 + orow * HID, p.x16 + orow * HID, lane);
        if (lane == 0) p.q_mask[b * Q + g] = p.gmap_mask[b * p.G + g];
    } else {
        // vp token j of episode b: vp_img + LN(Linear(vp_pos)) (vilmodel.py:832-833)
        const int vr = row - p.kv_rows_max - p.B * p.G;
        const int b = vr / p.V, j = vr - b * p.V;
        float f[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) f[k] = (k < p.v_kin) ? p.v_feat[static_cast<size_t>(vr) * p.v_kin + k] : 0.0f;
        small_linear(v, f, p.v_kin, p.v_w, p.v_bias, lane);
        ln_row(v, p.v_gamma, p.v_beta, 1e-12f, lane);
#pragma unroll
        for (int i = 0; i < HV; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(p.v_base + static_cast<size_t>(vr) * HID + (i * 32 + lane) * 4);
            v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
        }
        const size_t orow = static_cast<size_t>(b) * Q + p.G + j;
        store_row(v, p.x32 + orow * HID, p.x16 + orow * HID, lane);
        if (lane == 0) p.q_mask[b * Q + p.G + j] = p.vp_mask[b * p.V + j];
    }
    pdl_launch_dependents();
}

// ------------------------------------------------------------------------------------------- ClsPrediction tail
// logit[row] = w2 . LN(h[row]) + b2      (h = ReLU(Linear(x)) comes from the GEMM epilogue)
__global__ void __launch_bounds__(256) cls_tail_kernel(const float* h, const float* gamma, const float* beta, const float* w2,
                                                       const float* b2, float* logit, int rows) {
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4 v[HV];
#pragma unroll
    for (int i = 0; i < HV; ++i) v[i] = *reinterpret_cast<const float4*>(h + static_cast<size_t>(row) * HID + (i * 32 + lane) * 4);
    ln_row(v, gamma, beta, 1e-12f, lane);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(w2 + (i * 32 + lane) * 4);
        s += v[i].x * w.x + v[i].y * w.y + v[i].z * w.z + v[i].w * w.w;
    }
    s = warp_sum(s);
    if (lane == 0) logit[row] = s + b2[0];
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// ------------------------------------------------------------------------------------------- action logits
// vilmodel.py:859-907.  One CTA per episode.
struct LogitParams {
    const float* raw_global;   // [B, G]
    const float* raw_grid;     // [B, G]
    const float* raw_local;    // [B, V]
    const float* raw_obj;      // [B, V] or null
    const float* raw_fuse;     // [B] or null (fuse weight 0.5)
    const uint8_t* gmap_masks; const uint8_t* gmap_visited;   // [B, G]
    const uint8_t* vp_nav_masks; const uint8_t* vp_obj_masks; // [B, V]
    const int* fuse_src;       // [B, G]  >=0: add local[src]; -2: add the back-track sum; -1: nothing
    const uint8_t* bw_mask;    // [B, V]  candidates that are already visited (their local logits are summed)
    float* global_logits; float* grid_logits; float* local_logits; float* fused_logits; float* obj_logits;
    int G, V;
};

__global__ void __launch_bounds__(128) nav_logits_kernel(LogitParams p) {
    extern __shared__ float s_local[];
    __shared__ float s_bw;
    pdl_wait();
    const int b = blockIdx.x, tid = threadIdx.x;
    const float ninf = -INFINITY;
    float fw = 0.5f;
    if (p.raw_fuse) fw = 1.0f / (1.0f + expf(-p.raw_fuse[b]));
    for (int v = tid; v < p.V; v += blockDim.x) {
        float l = p.raw_local[b * p.V + v] * (1.0f - fw);
        if (!p.vp_nav_masks[b * p.V + v]) l = ninf;
        s_local[v] = l;
        p.local_logits[b * p.V + v] = l;
        if (p.raw_obj) {
            float o = p.raw_obj[b * p.V + v];
            if (!p.vp_obj_masks[b * p.V + v]) o = ninf;
            p.obj_logits[b * p.V + v] = o;
        }
    }
    __syncthreads();
    if (tid == 0) {
        float bw = 0.0f;   // sequential, candidate order (the reference accumulates with `+=` in a Python loop)
        for (int v = 1; v < p.V; ++v)
            if (p.bw_mask[b * p.V + v]) bw += s_local[v];
        s_bw = bw;
    }
    __syncthreads();
    for (int g = tid; g < p.G; g += blockDim.x) {
        const bool masked = p.gmap_visited[b * p.G + g] || !p.gmap_masks[b * p.G + g];
        float gl = p.raw_global[b * p.G + g] * fw;
        float gr = p.raw_grid[b * p.G + g];
        if (masked) { gl = ninf; gr = ninf; }
        p.global_logits[b * p.G + g] = gl;
        p.grid_logits[b * p.G + g] = gr;
        float f = gl;
        if (g == 0) f += s_local[0];
        else {
            const int src = p.fuse_src[b * p.G + g];
            if (src >= 0) f += s_local[src];
            else if (src == -2) f += s_bw;
        }
        p.fused_logits[b * p.G + g] = f;
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// ------------------------------------------------------------------------------------------- continuous-env action logits
// VLN_CE/vlnce_baselines/models/gridmap/vilmodel.py:791-800: w = sigmoid(fuse); logits[b, j] = global[b, j] * w +
// local[b, j] * (1 - w) for j < max(candidate_lengths), -inf where vp_nav_masks is false.
__global__ void __launch_bounds__(128) ce_logits_kernel(const float* raw_global, const float* raw_local, const float* raw_fuse,
                                                        const uint8_t* vp_nav_masks, float* fused, int G, int V, int maxc) {
    pdl_wait();
    const int b = blockIdx.x;
    const float fw = 1.0f / (1.0f + expf(-raw_fuse[b]));
    for (int j = threadIdx.x; j < maxc; j += blockDim.x) {
        const bool nav = vp_nav_masks[b * V + j] != 0;
        // both terms are masked with -inf in the reference, so a masked slot is -inf + -inf = -inf
        fused[b * maxc + j] = nav ? (raw_global[b * G + j] * fw + raw_local[b * V + j] * (1.0f - fw)) : -INFINITY;
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

}  // namespace gmm

// ----------------------------------------------------------------------------- C ABI
extern "C" int gridmm_layernorm(const float* x, int ldx, const float* gamma, const float* beta, float eps, float* out_f32,
                                int ld_f32, void* out_f16, int ld_f16, int rows, int hidden, cudaStream_t stream) {
    using namespace gmm;
    if (rows <= 0) return 0;
    if (hidden != HID || (ldx % 4) || (out_f32 && (ld_f32 % 4)) || (out_f16 && (ld_f16 % 4))) return GRIDMM_ERR_SHAPE;
    if (!x || !gamma || !beta) return GRIDMM_ERR_ARG;
    GMM_CUDA_CHECK(launch_pdl(layernorm_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, x, ldx, gamma, beta, eps, out_f32, ld_f32,
                                                         reinterpret_cast<__half*>(out_f16), ld_f16, rows));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_copy_rows(const float* x, int ldx, int in_rows_per_b, int in_off, float* out_f32, int ld_f32,
                                void* out_f16, int ld_f16, int out_rows_per_b, int out_off, int rows_per_b, int batch,
                                int hidden, cudaStream_t stream) {
    using namespace gmm;
    const int rows = rows_per_b * batch;
    if (rows <= 0) return 0;
    if (hidden != HID || (ldx % 4) || (out_f32 && (ld_f32 % 4)) || (out_f16 && (ld_f16 % 4))) return GRIDMM_ERR_SHAPE;
    if (!x || (!out_f32 && !out_f16)) return GRIDMM_ERR_ARG;
    GMM_CUDA_CHECK(launch_pdl(copy_rows_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, x, ldx, in_rows_per_b, in_off, out_f32, ld_f32,
                                                         reinterpret_cast<__half*>(out_f16), ld_f16, out_rows_per_b, out_off,
                                                         rows_per_b, rows));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_kv_index(const unsigned char* map_mask, const unsigned char* txt_mask, int batch, int S, int L, int* kv_pos,
                               int* kv_off, int* kv_cnt, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (S < 1 || L < 1 || S + L > 1024) return GRIDMM_ERR_SHAPE;
    if (!map_mask || !txt_mask || !kv_pos || !kv_off || !kv_cnt) return GRIDMM_ERR_ARG;
    if (batch > 8192) return GRIDMM_ERR_SHAPE;
    GMM_CUDA_CHECK(launch_pdl(kv_index_kernel, dim3(1), dim3(1024), (batch + 1) * sizeof(int), stream, map_mask, txt_mask, S, L, batch, kv_pos,
                              kv_off, kv_cnt));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_fusion_inputs(const float* map32, const float* txt32, const unsigned char* map_mask,
                                    const unsigned char* txt_mask, const unsigned char* gmap_mask, const unsigned char* vp_mask,
                                    float* x32, void* x16, void* kv16, unsigned char* kv_mask, unsigned char* q_mask,
                                    const int* kv_pos, const float* vp_pos, int vp_kin, const float* vp_w, const float* vp_bias,
                                    const float* vp_gamma, const float* vp_beta, const float* vp_img, int batch, int S, int L, int G,
                                    int V, int hidden, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (hidden != HID || S < G || L < 1 || G < 1 || V < 1) return GRIDMM_ERR_SHAPE;
    if (!map32 || !txt32 || !map_mask || !txt_mask || !gmap_mask || !vp_mask || !x32 || !x16 || !kv16 || !kv_mask || !q_mask)
        return GRIDMM_ERR_ARG;
    FusionInParams p{map32, txt32, map_mask, txt_mask, gmap_mask, vp_mask, x32, reinterpret_cast<__half*>(x16),
                     reinterpret_cast<__half*>(kv16), kv_mask, q_mask, batch, S, L, G, V, kv_pos,
                     vp_pos, vp_kin, vp_w, vp_bias, vp_gamma, vp_beta, vp_img};
    if (vp_pos && (!vp_w || !vp_bias || !vp_gamma || !vp_beta || !vp_img || vp_kin < 1 || vp_kin > 16)) return GRIDMM_ERR_ARG;
    const int rows = batch * (S + L + G + (vp_pos ? V : 0));
    GMM_CUDA_CHECK(launch_pdl(fusion_inputs_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, p));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_split_rows(const float* x, int ldx, int in_rows_per_b, int in_off, void* out_f16, int ld_f16,
                                 int k_total, int rows_per_b, int batch, int hidden, cudaStream_t stream) {
    using namespace gmm;
    const int rows = rows_per_b * batch;
    if (rows <= 0) return 0;
    if (hidden != HID || (ldx % 4) || (ld_f16 % 4) || (k_total % 4) || ld_f16 < 3 * k_total) return GRIDMM_ERR_SHAPE;
    if (!x || !out_f16) return GRIDMM_ERR_ARG;
    GMM_CUDA_CHECK(launch_pdl(split_rows_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, x, ldx, in_rows_per_b, in_off, reinterpret_cast<__half*>(out_f16),
                                                          ld_f16, k_total, rows_per_b, rows));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_pos_embed(const float* feat, int kin, const float* w, const float* bias, const float* gamma,
                                const float* beta, float eps, const float* base, const float* table, const long long* idx,
                                float* out_f32, void* out_f16, int in_rows_per_b, int out_rows_per_b, int out_row_off,
                                int rows, int hidden, cudaStream_t stream) {
    using namespace gmm;
    if (rows <= 0) return 0;
    if (hidden != HID || kin < 1 || kin > 16 || in_rows_per_b <= 0) return GRIDMM_ERR_SHAPE;
    if (!feat || !w || !bias || !gamma || !beta || (table && !idx) || (!out_f32 && !out_f16)) return GRIDMM_ERR_ARG;
    EmbedParams p{feat, kin, w, bias, gamma, beta, eps, base, table, idx, out_f32, reinterpret_cast<__half*>(out_f16),
                  in_rows_per_b, out_rows_per_b, out_row_off, rows};
    GMM_CUDA_CHECK(launch_pdl(embed_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, p));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_text_embed(const long long* ids, const float* word, const float* pos, const float* type0,
                                 const float* gamma, const float* beta, float eps, float* out_f32, void* out_f16, int batch,
                                 int L, int hidden, cudaStream_t stream) {
    using namespace gmm;
    const int rows = batch * L;
    if (rows <= 0) return 0;
    if (hidden != HID) return GRIDMM_ERR_SHAPE;
    if (!ids || !word || !pos || !type0 || !gamma || !beta || (!out_f32 && !out_f16)) return GRIDMM_ERR_ARG;
    GMM_CUDA_CHECK(launch_pdl(text_embed_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, ids, word, pos, type0, gamma, beta,
                              eps, out_f32, reinterpret_cast<__half*>(out_f16), L, rows));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_grid_assemble(const float* proj, const float* pos_fts, const int* cell_rank, const int* n_nonempty,
                                    const float* w, const float* bias, const float* gamma, const float* beta, float* map_f32,
                                    unsigned char* map_mask, int batch, int n_cells, int seq, int hidden,
                                    cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (hidden != HID || n_cells > 256 || seq < n_cells) return GRIDMM_ERR_SHAPE;
    if (!proj || !pos_fts || !cell_rank || !n_nonempty || !w || !bias || !gamma || !beta || !map_f32 || !map_mask)
        return GRIDMM_ERR_ARG;
    AssembleParams p{proj, pos_fts, cell_rank, n_nonempty, w, bias, gamma, beta, map_f32, map_mask, batch, n_cells, seq,
                     nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr};
    GMM_CUDA_CHECK(launch_pdl(grid_assemble_kernel, dim3(dim3(4, batch)), dim3(256), 0, stream, p));
    gridmm_count_launch(1);
    return 0;
}

// grid_assemble + the gmap tokens (rows [n_cells, seq): gmap_img + step embedding + LN(Linear(gmap_pos)), vilmodel.py:828-831) +
// the first pre-norm LayerNorm of grid_encoder (transformer.py:170-172) in one launch: map_f32 [B, seq, 768] and its mask are the
// encoder input, map_f16 = LayerNorm(map_f32; norm_gamma, norm_beta, norm_eps) is the operand of the first QKV projection.
extern "C" int gridmm_map_inputs(const float* proj, const float* pos_fts, const int* cell_rank, const int* n_nonempty,
                                 const float* w, const float* bias, const float* gamma, const float* beta, const float* gmap_pos,
                                 int gmap_kin, const float* gw, const float* gbias, const float* ggamma, const float* gbeta,
                                 const float* gmap_img, const float* step_table, const long long* step_ids,
                                 const unsigned char* gmap_mask, const float* norm_gamma, const float* norm_beta, float norm_eps,
                                 float* map_f32, void* map_f16, unsigned char* map_mask, int batch, int n_cells, int seq,
                                 int hidden, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (hidden != HID || n_cells > 256 || seq <= n_cells || gmap_kin < 1 || gmap_kin > 16) return GRIDMM_ERR_SHAPE;
    if (!proj || !pos_fts || !cell_rank || !n_nonempty || !w || !bias || !gamma || !beta || !gmap_pos || !gw || !gbias || !ggamma ||
        !gbeta || !gmap_img || !step_table || !step_ids || !gmap_mask || !norm_gamma || !norm_beta || !map_f32 || !map_f16 || !map_mask)
        return GRIDMM_ERR_ARG;
    AssembleParams p{proj, pos_fts, cell_rank, n_nonempty, w, bias, gamma, beta, map_f32, map_mask, batch, n_cells, seq,
                     gmap_pos, gmap_kin, gw, gbias, ggamma, gbeta, gmap_img, step_table, step_ids, gmap_mask, norm_gamma, norm_beta,
                     norm_eps, reinterpret_cast<__half*>(map_f16)};
    GMM_CUDA_CHECK(launch_pdl(grid_assemble_kernel, dim3(dim3(5, batch)), dim3(256), 0, stream, p));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_cls_tail(const float* h, const float* gamma, const float* beta, const float* w2, const float* b2,
                               float* logit, int rows, int hidden, cudaStream_t stream) {
    using namespace gmm;
    if (rows <= 0) return 0;
    if (hidden != HID) return GRIDMM_ERR_SHAPE;
    if (!h || !gamma || !beta || !w2 || !b2 || !logit) return GRIDMM_ERR_ARG;
    GMM_CUDA_CHECK(launch_pdl(cls_tail_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, h, gamma, beta, w2, b2, logit, rows));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_nav_logits(const float* raw_global, const float* raw_grid, const float* raw_local, const float* raw_obj,
                                 const float* raw_fuse, const unsigned char* gmap_masks, const unsigned char* gmap_visited,
                                 const unsigned char* vp_nav_masks, const unsigned char* vp_obj_masks, const int* fuse_src,
                                 const unsigned char* bw_mask, float* global_logits, float* grid_logits, float* local_logits,
                                 float* fused_logits, float* obj_logits, int batch, int G, int V, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (!raw_global || !raw_grid || !raw_local || !gmap_masks || !gmap_visited || !vp_nav_masks || !fuse_src || !bw_mask ||
        !global_logits || !grid_logits || !local_logits || !fused_logits || (raw_obj && (!vp_obj_masks || !obj_logits)))
        return GRIDMM_ERR_ARG;
    LogitParams p{raw_global, raw_grid, raw_local, raw_obj, raw_fuse, gmap_masks, gmap_visited, vp_nav_masks, vp_obj_masks,
                  fuse_src, bw_mask, global_logits, grid_logits, local_logits, fused_logits, obj_logits, G, V};
    GMM_CUDA_CHECK(launch_pdl(nav_logits_kernel, dim3(batch), dim3(128), V * sizeof(float), stream, p));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_ce_logits(const float* raw_global, const float* raw_local, const float* raw_fuse,
                                const unsigned char* vp_nav_masks, float* fused, int batch, int G, int V, int maxc,
                                cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (maxc < 1 || maxc > G || maxc > V) return GRIDMM_ERR_SHAPE;
    if (!raw_global || !raw_local || !raw_fuse || !vp_nav_masks || !fused) return GRIDMM_ERR_ARG;
    GMM_CUDA_CHECK(launch_pdl(ce_logits_kernel, dim3(batch), dim3(128), 0, stream, raw_global, raw_local, raw_fuse, vp_nav_masks,
                              fused, G, V, maxc));
    gridmm_count_launch(1);
    return 0;
}

// dst[i][0 .. nbytes[i]) = src[i][0 .. nbytes[i]) for i < n (n <= 24), one launch.  src / dst / nbytes are HOST arrays; the
// pointers in them must be device-accessible (device memory, or pinned host memory for sources).
extern "C" int gridmm_copy_segments(int n, const void* const* src, void* const* dst, const long long* nbytes, cudaStream_t stream) {
    using namespace gmm;
    if (n <= 0) return 0;
    if (n > SEG_MAX || !src || !dst || !nbytes) return GRIDMM_ERR_ARG;
    CopySegs p;
    p.n = 0;
    p.units[0] = 0;
    for (int i = 0; i < n; ++i) {
        if (nbytes[i] < 0 || (nbytes[i] > 0 && (!src[i] || !dst[i]))) return GRIDMM_ERR_ARG;
        if (nbytes[i] == 0) continue;
        const int k = p.n++;
        p.src[k] = reinterpret_cast<const uint8_t*>(src[i]);
        p.dst[k] = reinterpret_cast<uint8_t*>(dst[i]);
        p.bytes[k] = nbytes[i];
        p.units[k + 1] = p.units[k] + (nbytes[i] + 15) / 16;
    }
    if (p.n == 0) return 0;
    for (int i = p.n; i < SEG_MAX; ++i) { p.src[i] = nullptr; p.dst[i] = nullptr; p.bytes[i] = 0; p.units[i + 1] = p.units[p.n]; }
    const long long total = p.units[p.n];
    long long blocks = (total + 255) / 256;
    const int sms = gridmm_sm_count();
    const long long cap = static_cast<long long>(sms > 0 ? sms : 148) * 8;
    if (blocks > cap) blocks = cap;
    GMM_CUDA_CHECK(launch_pdl(copy_segments_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, p));
    gridmm_count_launch(1);
    return 0;
}

// ---- ragged ("packed") map sequence: see the comment above map_index_kernel
extern "C" int gridmm_map_index(const int* cell_rank, const int* n_nonempty, int batch, int n_cells, int G, int* m_off, int* m_info,
                                float* m_logz, int* cell_of_rank, int* m_goff, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (batch > 8192 || n_cells < 1 || n_cells > 256 || G < 1) return GRIDMM_ERR_SHAPE;
    if (!cell_rank || !n_nonempty || !m_off || !m_info || !m_logz || !cell_of_rank || !m_goff) return GRIDMM_ERR_ARG;
    GMM_CUDA_CHECK(launch_pdl(map_index_kernel, dim3(1), dim3(1024), (batch + 1) * sizeof(int), stream, cell_rank, n_nonempty, batch, n_cells,
                              G, m_off, m_info, m_logz, cell_of_rank, m_goff));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_map_inputs_packed(const float* proj, const float* pos_fts, const int* cell_of_rank, const int* m_off,
                                        const int* m_info, const float* m_logz, const float* w, const float* bias, const float* gamma,
                                        const float* beta, const float* gmap_pos, int gmap_kin, const float* gw, const float* gbias,
                                        const float* ggamma, const float* gbeta, const float* gmap_img, const float* step_table,
                                        const long long* step_ids, const unsigned char* gmap_mask, const float* norm_gamma,
                                        const float* norm_beta, float norm_eps, float* map_f32, void* map_f16, unsigned char* kvalid,
                                        float* kbias, int batch, int n_cells, int G, int hidden, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (hidden != HID || n_cells > 256 || G < 1 || gmap_kin < 1 || gmap_kin > 16 || batch > 8192) return GRIDMM_ERR_SHAPE;
    if (!proj || !pos_fts || !cell_of_rank || !m_off || !m_info || !m_logz || !w || !bias || !gamma || !beta || !gmap_pos || !gw ||
        !gbias || !ggamma || !gbeta || !gmap_img || !step_table || !step_ids || !gmap_mask || !norm_gamma || !norm_beta || !map_f32 ||
        !map_f16 || !kvalid || !kbias)
        return GRIDMM_ERR_ARG;
    MapPackedParams p{proj, pos_fts, cell_of_rank, m_off, m_info, m_logz, w, bias, gamma, beta, gmap_pos, gmap_kin, gw, gbias, ggamma,
                      gbeta, gmap_img, step_table, step_ids, gmap_mask, norm_gamma, norm_beta, norm_eps, map_f32,
                      reinterpret_cast<__half*>(map_f16), kvalid, kbias, batch, n_cells, G};
    const int rows = batch * (n_cells + G);      // upper bound; warps past m_off[batch] exit
    GMM_CUDA_CHECK(launch_pdl(map_inputs_packed_kernel, dim3((rows + 7) / 8), dim3(256), (batch + 1) * sizeof(int), stream, p));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_kv_index_packed(const int* m_off, const unsigned char* kvalid, const float* kbias, const unsigned char* txt_mask,
                                      int batch, int L, int* kv_src, float* kv_bias, int* kv_off, int* kv_cnt, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (L < 1 || batch > 8192) return GRIDMM_ERR_SHAPE;
    if (!m_off || !kvalid || !kbias || !txt_mask || !kv_src || !kv_bias || !kv_off || !kv_cnt) return GRIDMM_ERR_ARG;
    GMM_CUDA_CHECK(launch_pdl(kv_index_packed_kernel, dim3(1), dim3(1024), (batch + 1) * sizeof(int), stream, m_off, kvalid, kbias, txt_mask,
                              batch, L, kv_src, kv_bias, kv_off, kv_cnt));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_fusion_inputs_packed(const float* map32, const float* txt32, const int* kv_src, const int* kv_off, const int* m_goff,
                                           const unsigned char* gmap_mask, const unsigned char* vp_mask, float* x32, void* x16, void* kv16,
                                           unsigned char* q_mask, const float* vp_pos, int vp_kin, const float* vp_w, const float* vp_bias,
                                           const float* vp_gamma, const float* vp_beta, const float* vp_img, int batch, int L, int G, int V,
                                           int kv_rows_max, int hidden, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (hidden != HID || L < 1 || G < 1 || V < 1 || kv_rows_max < 1 || vp_kin < 1 || vp_kin > 16) return GRIDMM_ERR_SHAPE;
    if (!map32 || !txt32 || !kv_src || !kv_off || !m_goff || !gmap_mask || !vp_mask || !x32 || !x16 || !kv16 || !q_mask || !vp_pos ||
        !vp_w || !vp_bias || !vp_gamma || !vp_beta || !vp_img)
        return GRIDMM_ERR_ARG;
    FusionPackedParams p{map32, txt32, kv_src, kv_off, m_goff, gmap_mask, vp_mask, x32, reinterpret_cast<__half*>(x16),
                         reinterpret_cast<__half*>(kv16), q_mask, batch, L, G, V, kv_rows_max, vp_pos, vp_kin, vp_w, vp_bias, vp_gamma,
                         vp_beta, vp_img};
    const int rows = kv_rows_max + batch * (G + V);
    GMM_CUDA_CHECK(launch_pdl(fusion_inputs_packed_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, p));
    gridmm_count_launch(1);
    return 0;
}
