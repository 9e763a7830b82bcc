// Instruction-relevance pooling: the reference's per-episode loop
//   grid_fts_weight = (grid_fts @ text_fts).max(-1)            map_nav_src/models/vilmodel.py:797-798
//   for i in range(196): softmax over the points of cell i, weighted sum   vilmodel.py:801-807
// as ONE persistent kernel that reads every valid patch-feature row from HBM exactly once.
//
//   * rows are streamed in cell-sorted order (gridmm_grid_update produced `perm`), 64 rows per tile,
//     gathered with 16-byte cp.async into a SWIZZLE_128B K-major tile (12 chunks of 64 x 128 B);
//     the next tile's rows are pulled into L2 with cp.async.bulk.prefetch (one 1536-byte request per row);
//   * relevance  S[64, L] = X_tile . text_fts^T  on tcgen05 (UMMA M=64, N=L, K=16 x 48), text_fts of the
//     current episode resident in shared memory (TMA, 12 chunks of L x 128 B), accumulator in TMEM
//     (double buffered), w = max_l S  (over ALL L positions, padding included -- vilmodel.py:798);
//   * per-cell softmax + weighted sum on CUDA cores straight from the resident tile: cells are contiguous
//     row segments; a cell that straddles tiles is carried in registers with the usual online-softmax
//     rescale.  CTA ranges are cut at cell boundaries, so no atomics and no cross-CTA merge exist.
//   * grid_proj is applied AFTER pooling by the GEMM kernel (sum_j p_j (W x_j + b) = W (sum_j p_j x_j) + b),
//     so this kernel emits the pooled raw feature per non-empty cell, compacted in ascending cell order
//     (the order vilmodel.py:819 gathers them in), as fp16 GEMM input.
//
// HBM roofline: algorithmic bytes = valid_rows * D * 2 (+ L*D*2 per episode + outputs); see DESIGN.md.
#include "common.cuh"
#include "host_util.h"

namespace gmm {

constexpr int POOL_ROWS = 64;
constexpr int POOL_GATHER_WARPS = 4;
constexpr int POOL_MMA_WARP = 4;
constexpr int POOL_TMA_WARP = 5;
constexpr int POOL_EPI_WARP0 = 6;      // warps 6..9  (TMEM lane quadrant = warp & 3)
constexpr int POOL_POOL_WARP0 = 10;    // warps 10..  (D / 128 of them)
constexpr int POOL_MAX_BATCH = 1024;
constexpr int POOL_MAX_CELLS = 256;
constexpr int POOL_LAG = 4;            // cp.async groups in flight per gather thread

struct PoolParams {
    const __half* fts;       // feature slab; row r at fts + r * D
    const int* slots;        // [B, t_cap]   slab slot of (episode, step)
    const int* perm;         // [B, cap]     valid point indices sorted by cell
    const int* cell_start;   // [B, n_cells + 1]
    const int* cell_rank;    // [B, n_cells]
    __half* pooled;          // [B, n_cells, D]  compacted by cell rank
    float* w_out;            // [B, cap] relevance weight per sorted position, or null (tests)
    int batch, t_cap, cap, n_cells;
    int l_pad;               // text positions (multiple of 8, <= 80 for D = 768)
    int slot_rows, view_rows, tok_off;   // row = slot*slot_rows + view*view_rows + tok_off + patch
};

struct Tile {
    int b, pos, nrows;
};

struct Walker {
    const int* vbase;
    int b, pos, g, g_end;
    __device__ void init(const int* vb, int batch, int g0, int g1) {
        vbase = vb; g = g0; g_end = g1; b = 0; pos = 0;
        if (g0 < g1) {
            int lo = 0, hi = batch;   // largest b with vbase[b] <= g0
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (vbase[mid] <= g0) lo = mid; else hi = mid;
            }
            b = lo; pos = g0 - vbase[lo];
        }
    }
    __device__ bool next(Tile& t) {
        if (g >= g_end) return false;
        while (pos >= vbase[b + 1] - vbase[b]) { ++b; pos = 0; }
        const int nv = vbase[b + 1] - vbase[b];
        const int nrows = min(min(POOL_ROWS, nv - pos), g_end - g);
        t.b = b; t.pos = pos; t.nrows = nrows;
        pos += nrows; g += nrows;
        return true;
    }
};

__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        default: cp_async_wait<4>(); break;
    }
}

__device__ __forceinline__ void l2_prefetch_bulk(const void* g, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int D>
struct PoolSmem {
    static constexpr int CH = D / 64;
    static constexpr int A_CHUNK = POOL_ROWS * 128;
    static constexpr int A_BYTES = CH * A_CHUNK;
    static constexpr int b_bytes(int l_pad) { return CH * l_pad * 128; }
    // after A and B: barriers and small arrays
    static constexpr int MISC_BYTES = 64 * 8                          // barriers
                                      + 2 * POOL_ROWS * 4 * 4         // w, p, fin, rank (double buffered)
                                      + 64                            // scalars
                                      + (POOL_MAX_CELLS + 1) * 4 + POOL_MAX_CELLS * 4   // cell_start, cell_rank of current episode
                                      + (POOL_MAX_BATCH + 1) * 4;     // vbase
    static constexpr int total(int l_pad) { return 1024 + b_bytes(l_pad) + A_BYTES + MISC_BYTES; }
};

template <int D>
__global__ void __launch_bounds__(320 + D / 4, 1)
pool_kernel(const __grid_constant__ CUtensorMap tmT, PoolParams p) {
    using L = PoolSmem<D>;
    constexpr int CH = L::CH;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_chunk = p.l_pad * 128;
    uint8_t* sB = smem;
    uint8_t* sA = smem + CH * b_chunk;
    uint8_t* misc = sA + L::A_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(misc);        // [CH]
    uint64_t* a_empty = a_full + 12;                               // [CH]
    uint64_t* d_full = a_empty + 12;                               // [2]
    uint64_t* d_empty = d_full + 2;                                // [2]
    uint64_t* p_full = d_empty + 2;                                // [2]
    uint64_t* b_full = p_full + 2;                                 // [1]
    uint64_t* b_empty = b_full + 1;                                // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + 1);
    float* s_w = reinterpret_cast<float*>(misc + 64 * 8);          // [2][64]
    float* s_p = s_w + 2 * POOL_ROWS;                              // [2][64]
    float* s_fin = s_p + 2 * POOL_ROWS;                            // [2][64]  1/sum at the last row of a finished cell, else 0
    int* s_rank = reinterpret_cast<int*>(s_fin + 2 * POOL_ROWS);   // [2][64]
    float* s_scal = reinterpret_cast<float*>(s_rank + 2 * POOL_ROWS);   // [0..1] carry scale per buffer, [2] m_carry, [3] s_carry
    int* s_range = reinterpret_cast<int*>(s_scal + 8);             // [0] g_start, [1] g_end
    int* s_cs = s_range + 8;                                       // [n_cells + 1]
    int* s_cr = s_cs + POOL_MAX_CELLS + 1;                         // [n_cells]
    int* s_vbase = s_cr + POOL_MAX_CELLS;                          // [batch + 1]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n_cells = p.n_cells;

    // ---------------------------------------------------------------- setup: barriers, TMEM, schedule
    if (tid == 0) {
        for (int k = 0; k < CH; ++k) {
            mbar_init(&a_full[k], POOL_GATHER_WARPS * 32);
            mbar_init(&a_empty[k], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], 4);
            mbar_init(&p_full[i], 128);
        }
        mbar_init(b_full, 1);
        mbar_init(b_empty, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmT);
    }
    if (warp == POOL_MMA_WARP) tmem_alloc(tmem_slot, 256);
    // exclusive prefix of the valid-point counts: vbase[b] = sum_{b' < b} cell_start[b'][n_cells]
    for (int i = tid; i < p.batch; i += blockDim.x) s_vbase[i + 1] = p.cell_start[i * (n_cells + 1) + n_cells];
    if (tid == 0) s_vbase[0] = 0;
    __syncthreads();
    if (warp == 0) {
        const int per = (p.batch + 31) / 32;
        const int lo = min(lane * per, p.batch), hi = min(lo + per, p.batch);
        int sum = 0;
        for (int i = lo; i < hi; ++i) sum += s_vbase[i + 1];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - sum;
        for (int i = lo; i < hi; ++i) { run += s_vbase[i + 1]; s_vbase[i + 1] = run; }
    }
    __syncthreads();
    if (tid < 2) {
        // CTA range [g0, g1) in global sorted-valid coordinates, snapped up to a cell boundary
        const int total = s_vbase[p.batch];
        const long long tgt = (static_cast<long long>(blockIdx.x + tid) * total) / gridDim.x;
        int g = static_cast<int>(tgt);
        if (g >= total) g = total;
        else if (g > 0) {
            int lo = 0, hi = p.batch;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_vbase[mid] <= g) lo = mid; else hi = mid;
            }
            const int local = g - s_vbase[lo];
            const int* cs = p.cell_start + lo * (n_cells + 1);
            int a = 0, c = n_cells;    // first boundary index with cs[idx] >= local
            while (a < c) {
                const int mid = (a + c) >> 1;
                if (cs[mid] >= local) c = mid; else a = mid + 1;
            }
            g = s_vbase[lo] + cs[a];
        }
        s_range[tid] = g;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int g0 = s_range[0], g1 = s_range[1];

    Walker wk;
    wk.init(s_vbase, p.batch, g0, g1);
    Tile t;

    if (warp < POOL_GATHER_WARPS) {
        // ------------------------------------------------------------ gather producers
        const int u = tid & 7;            // 16-byte unit inside the 128-byte chunk row
        const int r0 = tid >> 3;          // rows r0 + 16*i
        Walker ahead = wk;
        Tile tn;
        bool have_next = ahead.next(tn);
        int it = 0;
        while (wk.next(t)) {
            have_next = ahead.next(tn);   // `ahead` runs one tile in front of `wk`
            const int* perm_b = p.perm + static_cast<size_t>(t.b) * p.cap + t.pos;
            const __half* src[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + 16 * i;
                src[i] = nullptr;
                if (r < t.nrows) {
                    const int j = perm_b[r];
                    const int step = j / 588, q = j - step * 588;
                    const int v = q / 49, k = q - v * 49;
                    const long long row = static_cast<long long>(p.slots[t.b * p.t_cap + step]) * p.slot_rows +
                                          v * p.view_rows + p.tok_off + k;
                    src[i] = p.fts + row * D + u * 8;
                }
            }
            // pull the NEXT tile's rows into L2 with one bulk request per row
            if (have_next && tid < tn.nrows) {
                const int j = p.perm[static_cast<size_t>(tn.b) * p.cap + tn.pos + tid];
                const int step = j / 588, q = j - step * 588;
                const int v = q / 49, k = q - v * 49;
                const long long row = static_cast<long long>(p.slots[tn.b * p.t_cap + step]) * p.slot_rows +
                                      v * p.view_rows + p.tok_off + k;
                l2_prefetch_bulk(p.fts + row * D, D * 2);
            }
            const uint32_t ph = it & 1;
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                mbar_wait(&a_empty[k], ph ^ 1);
                const uint32_t dst = smem_u32(sA + k * L::A_CHUNK);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (src[i]) cp_async_16(dst + sw128_offset(r0 + 16 * i, u), src[i] + k * 64);
                cp_async_commit();
                if (k >= POOL_LAG) {
                    cp_async_wait<POOL_LAG>();
                    fence_proxy_async_smem();
                    mbar_arrive(&a_full[k - POOL_LAG]);
                }
            }
#pragma unroll
            for (int k = CH - POOL_LAG; k < CH; ++k) {
                cp_async_wait_dyn(CH - 1 - k);
                fence_proxy_async_smem();
                mbar_arrive(&a_full[k]);
            }
            ++it;
        }
    } else if (warp == POOL_MMA_WARP) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(POOL_ROWS, p.l_pad);
            int it = 0, cur_b = -1, visits = 0;
            bool more = wk.next(t);
            while (more) {
                Tile tnext;
                const bool more_next = wk.next(tnext);
                const int buf = it & 1;
                if (t.b != cur_b) {
                    mbar_wait(b_full, visits & 1);
                    cur_b = t.b;
                    ++visits;
                }
                mbar_wait(&d_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 128;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    mbar_wait(&a_full[k], it & 1);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128_kmajor(smem_u32(sA + k * L::A_CHUNK));
                    const uint64_t db = umma_desc_sw128_kmajor(smem_u32(sB + k * b_chunk));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_f16_ss(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (k | kk) ? 1u : 0u);
                }
                umma_commit(&d_full[buf]);
                if (!more_next || tnext.b != cur_b) umma_commit(b_empty);   // text_fts of this episode may be replaced
                t = tnext;
                more = more_next;
                ++it;
            }
        }
    } else if (warp == POOL_TMA_WARP) {
        // ------------------------------------------------------------ text_fts loader (one episode resident)
        if (lane == 0) {
            int cur_b = -1, visits = 0;
            while (wk.next(t)) {
                if (t.b == cur_b) continue;
                cur_b = t.b;
                mbar_wait(b_empty, (visits & 1) ^ 1);
                mbar_arrive_expect_tx(b_full, CH * b_chunk);
                for (int k = 0; k < CH; ++k) tma_load_2d(sB + k * b_chunk, &tmT, k * 64, t.b * p.l_pad, b_full);
                ++visits;
            }
        }
    } else if (warp < POOL_POOL_WARP0) {
        // ------------------------------------------------------------ relevance max + per-cell softmax weights
        const int q = warp & 3;
        const int e = tid - POOL_EPI_WARP0 * 32;       // 0..127
        int it = 0, cur_b = -1;
        if (e == 0) { s_scal[2] = 0.0f; s_scal[3] = 0.0f; }
        while (wk.next(t)) {
            const int buf = it & 1;
            if (t.b != cur_b) {
                // all 128 threads are past the previous tile's reads of s_cs/s_cr (barrier 2 below)
                cur_b = t.b;
                for (int i = e; i <= n_cells; i += 128) s_cs[i] = p.cell_start[t.b * (n_cells + 1) + i];
                for (int i = e; i < n_cells; i += 128) s_cr[i] = p.cell_rank[t.b * n_cells + i];
            }
            mbar_wait(&d_full[buf], (it >> 1) & 1);
            tc_fence_after();
            float mx = -INFINITY;
            for (int c = 0; c < p.l_pad; c += 16) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 128 + c, v);
                tmem_ld_wait();
                const int lim = min(16, p.l_pad - c);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[buf]);
            // UMMA M=64: accumulator row 16*q + i lives in TMEM lane 32*q + i (i < 16)
            if (lane < 16) s_w[buf * POOL_ROWS + q * 16 + lane] = mx;
            named_bar_sync(1, 128);
            // --- softmax weights of this tile's rows; one thread per row
            float m_carry = s_scal[2], s_carry = s_scal[3];
            float my_p = 0.0f, my_fin = 0.0f, new_m = 0.0f, new_s = 0.0f, cscale = 1.0f;
            int my_rank = 0;
            bool writes_carry = false;
            if (e < t.nrows) {
                const int P = t.pos + e;
                int a = 0, c = n_cells;        // last cell with cs[cell] <= P
                while (c - a > 1) {
                    const int mid = (a + c) >> 1;
                    if (s_cs[mid] <= P) a = mid; else c = mid;
                }
                const int cellid = a;
                const int c_lo = s_cs[cellid], c_hi = s_cs[cellid + 1];
                const int seg_lo = max(c_lo, t.pos) - t.pos;
                const int seg_hi = min(c_hi, t.pos + t.nrows) - t.pos;
                const float* wrow = s_w + buf * POOL_ROWS;
                float m = -INFINITY;
                for (int j = seg_lo; j < seg_hi; ++j) m = fmaxf(m, wrow[j]);
                float s0 = 0.0f;
                const bool continues = (c_lo < t.pos);
                if (continues) {
                    const float mn = fmaxf(m, m_carry);
                    cscale = expf(m_carry - mn);
                    s0 = s_carry * cscale;
                    m = mn;
                }
                float s = s0;
                for (int j = seg_lo; j < seg_hi; ++j) s += expf(wrow[j] - m);
                my_p = expf(wrow[e] - m);
                if (e == seg_hi - 1) {
                    if (c_hi <= t.pos + t.nrows) { my_fin = 1.0f / s; my_rank = s_cr[cellid]; }
                    else { writes_carry = true; new_m = m; new_s = s; }
                }
                if (e == 0) s_scal[buf] = continues ? cscale : 1.0f;
                if (p.w_out) p.w_out[static_cast<size_t>(t.b) * p.cap + P] = wrow[e];
            }
            if (e < POOL_ROWS) {
                s_p[buf * POOL_ROWS + e] = my_p;
                s_fin[buf * POOL_ROWS + e] = my_fin;
                s_rank[buf * POOL_ROWS + e] = my_rank;
            }
            named_bar_sync(2, 128);             // everyone has read the old carry / s_cs
            if (writes_carry) { s_scal[2] = new_m; s_scal[3] = new_s; }
            mbar_arrive(&p_full[buf]);          // release: s_p / s_fin / s_rank / s_scal[buf] are visible to the pooling warps
            ++it;
        }
    } else {
        // ------------------------------------------------------------ weighted sums from the resident tile
        const int pt = tid - POOL_POOL_WARP0 * 32;     // owns columns 4*pt .. 4*pt+3
        const int pw = pt >> 5;                        // pooling warp: chunks 2*pw, 2*pw+1
        const int k = pt >> 4;
        const int u = (pt & 15) >> 1;
        const int sub = (pt & 1) * 8;
        const uint8_t* chunk = sA + k * L::A_CHUNK;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        int it = 0;
        while (wk.next(t)) {
            const int buf = it & 1;
            mbar_wait(&p_full[buf], (it >> 1) & 1);
            mbar_wait(&a_full[k], it & 1);
            const float cs = s_scal[buf];
            acc0 *= cs; acc1 *= cs; acc2 *= cs; acc3 *= cs;
            const float* pp = s_p + buf * POOL_ROWS;
            const float* pf = s_fin + buf * POOL_ROWS;
            const int* pr = s_rank + buf * POOL_ROWS;
            __half* out_b = p.pooled + static_cast<size_t>(t.b) * n_cells * D + pt * 4;
            for (int r = 0; r < t.nrows; ++r) {
                const uint2 raw = *reinterpret_cast<const uint2*>(chunk + sw128_offset(r, u) + sub);
                const float2 x01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                const float2 x23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                const float w = pp[r];
                acc0 = fmaf(w, x01.x, acc0); acc1 = fmaf(w, x01.y, acc1);
                acc2 = fmaf(w, x23.x, acc2); acc3 = fmaf(w, x23.y, acc3);
                const float fin = pf[r];
                if (fin != 0.0f) {
                    const __half2 h01 = __floats2half2_rn(acc0 * fin, acc1 * fin);
                    const __half2 h23 = __floats2half2_rn(acc2 * fin, acc3 * fin);
                    uint2 o;
                    o.x = *reinterpret_cast<const uint32_t*>(&h01);
                    o.y = *reinterpret_cast<const uint32_t*>(&h23);
                    *reinterpret_cast<uint2*>(out_b + static_cast<size_t>(pr[r]) * D) = o;
                    acc0 = acc1 = acc2 = acc3 = 0.f;
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a_empty[2 * pw]);
                mbar_arrive(&a_empty[2 * pw + 1]);
            }
            ++it;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == POOL_MMA_WARP) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace gmm

extern "C" int gridmm_pool(const void* fts, int feat_dim, const int* slots, int t_cap, int slot_rows, int view_rows,
                           int tok_off, const int* perm, int cap, const int* cell_start, const int* cell_rank, int n_cells,
                           const void* text_fts, int l_pad, int batch, void* pooled, float* w_out, int num_ctas,
                           cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (!fts || !slots || !perm || !cell_start || !cell_rank || !text_fts || !pooled) return GRIDMM_ERR_ARG;
    if (batch > POOL_MAX_BATCH || n_cells > POOL_MAX_CELLS || l_pad < 8 || (l_pad % 8) || l_pad > 128) return GRIDMM_ERR_SHAPE;
    if (feat_dim != 768 && feat_dim != 512) return GRIDMM_ERR_SHAPE;
    CUtensorMap tmT;
    int rc = make_tmap_f16_2d(&tmT, text_fts, static_cast<uint64_t>(feat_dim), static_cast<uint64_t>(batch) * l_pad,
                              static_cast<uint64_t>(feat_dim) * 2, 64, static_cast<uint32_t>(l_pad));
    if (rc) return rc;
    PoolParams p;
    p.fts = reinterpret_cast<const __half*>(fts); p.slots = slots; p.perm = perm; p.cell_start = cell_start;
    p.cell_rank = cell_rank; p.pooled = reinterpret_cast<__half*>(pooled); p.w_out = w_out;
    p.batch = batch; p.t_cap = t_cap; p.cap = cap; p.n_cells = n_cells; p.l_pad = l_pad;
    p.slot_rows = slot_rows; p.view_rows = view_rows; p.tok_off = tok_off;
    int dev = 0, sms = 0, max_smem = 0;
    GMM_CUDA_CHECK(cudaGetDevice(&dev));
    GMM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GMM_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int grid = num_ctas > 0 ? num_ctas : sms;
    if (feat_dim == 768) {
        const int smem = PoolSmem<768>::total(l_pad);
        if (smem > max_smem) return GRIDMM_ERR_SHAPE;
        GMM_CUDA_CHECK(cudaFuncSetAttribute(pool_kernel<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        pool_kernel<768><<<grid, 320 + 768 / 4, smem, stream>>>(tmT, p);
    } else {
        const int smem = PoolSmem<512>::total(l_pad);
        if (smem > max_smem) return GRIDMM_ERR_SHAPE;
        GMM_CUDA_CHECK(cudaFuncSetAttribute(pool_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        pool_kernel<512><<<grid, 320 + 512 / 4, smem, stream>>>(tmT, p);
    }
    gridmm_count_launch(1);
    return static_cast<int>(cudaGetLastError());
}
