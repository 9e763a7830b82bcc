// Instruction-relevance pooling: the reference's per-episode loop
//   grid_fts_weight = (grid_fts @ text_fts).max(-1)            map_nav_src/models/vilmodel.py:797-798
//   for i in range(196): softmax over the points of cell i, weighted sum   vilmodel.py:801-807
// as ONE persistent kernel that reads every valid patch-feature row from HBM exactly once.
//
//   * rows are streamed in cell-sorted order (gridmm_grid_update produced `perm`), 64 rows per tile, gathered with
//     16-byte cp.async into a SWIZZLE_128B K-major tile (D/64 chunks of 64 x 128 B); two tile buffers, so tile i+1 is
//     in flight while tile i is multiplied, reduced and pooled; rows of tile i+2 are pulled into L2 with
//     cp.async.bulk.prefetch (one request per 1536-byte row) so DRAM sees whole rows;
//   * relevance  S^T[L, 64] = text_fts[L, D] . X_tile^T  on tcgen05 with the A operand (text_fts of the current
//     episode, up to 128 positions) held in TENSOR MEMORY for the whole episode (tcgen05.st once per episode, 3 KB per
//     lane), B operand = the feature tile in shared memory, accumulator in TMEM (double buffered);
//     w = max over ALL text positions (padding included -- vilmodel.py:798) = max over TMEM lanes, taken with a
//     warp butterfly (62 shuffles per thread) + one shared-memory hop across the four lane quadrants;
//   * per-cell softmax + weighted sum on CUDA cores straight from the resident tile: cells are contiguous row segments;
//     a cell that straddles tiles is carried in registers with the usual online-softmax rescale.  CTA ranges are cut at
//     cell boundaries, so no atomics and no cross-CTA merge exist and the result is deterministic;
//   * grid_proj is applied AFTER pooling by the GEMM kernel (sum_j p_j (W x_j + b) = W (sum_j p_j x_j) + b), so this
//     kernel emits the pooled raw feature per non-empty cell, compacted in ascending cell order (the order
//     vilmodel.py:819 gathers them in), as fp16 GEMM input.
//
// HBM roofline: algorithmic bytes = valid_rows * D * 2 (+ L*D*2 per episode + outputs); see DESIGN.md.
#include "common.cuh"
#include "host_util.h"

namespace gmm {

constexpr int POOL_ROWS = 64;
constexpr int POOL_GATHER_WARP0 = 4;   // warps 0..3: epilogue (TMEM lane quadrant = warp); warps 4..7: gather
constexpr int POOL_MMA_WARP = 8;
constexpr int POOL_POOL_WARP0 = 9;     // warps 9..  (D / 128 of them)
constexpr int POOL_FIXED_THREADS = 9 * 32;
constexpr int POOL_MAX_BATCH = 1024;
constexpr int POOL_MAX_CELLS = 256;
constexpr int POOL_TMEM_COLS = 512;

struct PoolParams {
    const __half* fts;       // feature slab; row r at fts + r * D
    const int* slots;        // [B, t_cap]   slab slot of (episode, step)
    const int* perm;         // [B, cap]     valid point indices sorted by cell
    const int* cell_start;   // [B, n_cells + 1]
    const int* cell_rank;    // [B, n_cells]
    const __half* text;      // [B, l_pad, D] text_fts
    const uint4* text_ws;    // [B, D/8, 128] lane-major copy of text_fts (written by text_to_lanes_kernel)
    __half* pooled;          // [B, n_cells, D]  compacted by cell rank
    float* w_out;            // [B, cap] relevance weight per sorted position, or null (tests)
    int batch, t_cap, cap, n_cells;
    int l_pad;               // text positions (<= 128)
    int slot_rows, view_rows, tok_off;   // row = slot*slot_rows + view*view_rows + tok_off + patch
    long long* dbg;          // optional [grid][16] cycle counters (tools/microbench.py), null in production
    int mode;                // debug experiments (tools/microbench.py): bit0 skip the pooling loop, bit1 skip softmax weights
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
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        case 6: cp_async_wait<6>(); break;
        case 7: cp_async_wait<7>(); break;
        case 8: cp_async_wait<8>(); break;
        case 9: cp_async_wait<9>(); break;
        case 10: cp_async_wait<10>(); break;
        default: cp_async_wait<11>(); break;
    }
}

__device__ __forceinline__ void l2_prefetch_bulk(const void* g, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One butterfly level of the lane-max: every lane keeps the half of its N columns selected by `bit` of its lane id and
// merges in the partner's copy of that half.  After the 64 -> 2 levels lane l holds the warp-wide max of columns 2l, 2l+1.
template <int N>
__device__ __forceinline__ void lane_max_level(float (&v)[64], int lane, int bit) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
        const float keep = up ? v[N / 2 + j] : v[j];
        const float send = up ? v[j] : v[N / 2 + j];
        v[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, bit));
    }
}

template <int D>
struct PoolSmem {
    static constexpr int CH = D / 64;
    static constexpr int A_CHUNK = POOL_ROWS * 128;
    static constexpr int A_BYTES = CH * A_CHUNK;                  // one feature tile
    static constexpr int MISC_BYTES = 64 * 8                      // barriers + tmem slot
                                      + 2 * POOL_ROWS * 4 * 5     // w, p, seg_end, seg_fin, seg_rank (double buffered)
                                      + 4 * POOL_ROWS * 4         // per-quadrant partial maxima
                                      + 64                        // scalars
                                      + (POOL_MAX_CELLS + 1) * 4 + POOL_MAX_CELLS * 4   // cell_start, cell_rank of the episode
                                      + (POOL_MAX_BATCH + 1) * 4; // vbase
    static constexpr int TOTAL = 1024 + 2 * A_BYTES + MISC_BYTES;
};

// text_fts [B, l_pad, D] -> lane-major copy [B, D/8, 128] of 16-byte units: unit c of text position t sits at
// ((b * D/8 + c) * 128 + t) * 16 bytes, so the 32 lanes of a warp (32 consecutive positions) read 512 contiguous bytes
// when the operand is moved into tensor memory.  Positions >= l_pad replicate position 0.
template <int D>
__global__ void __launch_bounds__(128) text_to_lanes_kernel(const __half* text, uint4* ws, int l_pad) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, c = blockIdx.x, t = threadIdx.x;
    const int tok = (t < l_pad) ? t : 0;
    const uint4 v = *reinterpret_cast<const uint4*>(text + (static_cast<size_t>(b) * l_pad + tok) * D + c * 8);
    ws[(static_cast<size_t>(b) * (D / 8) + c) * 128 + t] = v;
}

template <int D>
__global__ void __launch_bounds__(POOL_FIXED_THREADS + D / 4, 1)
pool_kernel(PoolParams p) {
    using L = PoolSmem<D>;
    constexpr int CH = L::CH;
    constexpr int NPW = D / 128;                  // pooling warps
    constexpr int A_COLS = D / 2;                 // TMEM columns of the text operand (two fp16 per column)
    constexpr int D_COL0 = A_COLS;                // accumulators behind it: 2 x 64 columns
    static_assert(A_COLS + 2 * POOL_ROWS <= POOL_TMEM_COLS, "text operand + accumulators must fit in tensor memory");
    extern __shared__ uint8_t smem_raw[];
    // align by pointer arithmetic on the __shared__ array (an integer round-trip would demote every access below to a
    // generic LD/ST instead of LDS/STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                            // [2][CH][64 x 128 B]
    uint8_t* misc = sA + 2 * L::A_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(misc);          // [2] tile buffer filled
    uint64_t* a_empty = a_full + 2;                                // [2] tile buffer drained by the pooling warps
    uint64_t* d_full = a_empty + 2;                                // [2]
    uint64_t* d_empty = d_full + 2;                                // [2]
    uint64_t* p_full = d_empty + 2;                                // [2]
    uint64_t* t_ready = p_full + 2;                                // [1] text operand of the episode is in TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 60);
    float* s_w = reinterpret_cast<float*>(misc + 64 * 8);          // [2][64] relevance weight per row
    float* s_p = s_w + 2 * POOL_ROWS;                              // [2][64] exp(w - m) per row
    int* s_send = reinterpret_cast<int*>(s_p + 2 * POOL_ROWS);     // [2][64] exclusive end row of the k-th cell segment
    float* s_sfin = reinterpret_cast<float*>(s_send + 2 * POOL_ROWS);   // [2][64] 1/sum if the segment closes its cell, else 0
    int* s_srank = reinterpret_cast<int*>(s_sfin + 2 * POOL_ROWS); // [2][64] compact output slot of the segment's cell
    float* s_part = reinterpret_cast<float*>(s_srank + 2 * POOL_ROWS);  // [4][64]
    float* s_scal = s_part + 4 * POOL_ROWS;                        // [0..1] carry scale per buffer, [2] m_carry, [3] s_carry
    int* s_nseg = reinterpret_cast<int*>(s_scal + 4);              // [2] segments in the tile
    int* s_range = s_nseg + 4;                                     // [0] g_start, [1] g_end
    int* s_cs = s_range + 8;                                       // [n_cells + 1]
    int* s_cr = s_cs + POOL_MAX_CELLS + 1;                         // [n_cells]
    int* s_vbase = s_cr + POOL_MAX_CELLS;                          // [batch + 1]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n_cells = p.n_cells;

    // ---------------------------------------------------------------- setup: barriers, TMEM, schedule
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 64);
            mbar_init(&a_empty[i], NPW);
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], 4);
            mbar_init(&p_full[i], 128);
        }
        mbar_init(t_ready, 128);
        fence_mbar_init();
    }
    pdl_launch_dependents();
    if (warp == POOL_MMA_WARP) tmem_alloc(tmem_slot, POOL_TMEM_COLS);
    pdl_wait();      // barrier init / TMEM allocation above overlap the previous kernel's tail
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

    if (warp >= POOL_GATHER_WARP0 && warp < POOL_MMA_WARP) {
        // ------------------------------------------------------------ gather producers
        // Two independent groups of 64 threads: group g fills tile buffer g with the tiles of parity g.  The row
        // addresses of a group's NEXT tile are resolved (perm -> slot -> row, two dependent global loads) and its rows
        // pulled into L2 while the group still waits for its buffer; once the buffer is free the WHOLE tile is issued
        // (96 x 16 B per thread) before anything is waited for, so ~96 KB per buffer are in flight.
        const int grp = (warp - POOL_GATHER_WARP0) >> 1;
        const int gt = (tid - POOL_GATHER_WARP0 * 32) & 63;   // 0..63 within the group
        const int u = gt & 7;             // 16-byte unit inside the 128-byte chunk row
        const int r0 = gt >> 3;           // rows r0 + 8*i
        uint8_t* tile_base = sA + grp * L::A_BYTES;
        auto resolve = [&](const Tile& tt, const __half* (&src)[8]) {
            const int* perm_b = p.perm + static_cast<size_t>(tt.b) * p.cap + tt.pos;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + 8 * i;
                src[i] = nullptr;
                if (r < tt.nrows) {
                    const int j = perm_b[r];
                    const int step = j / 588, q = j - step * 588;
                    const int v = q / 49, k = q - v * 49;
                    const long long row = static_cast<long long>(p.slots[tt.b * p.t_cap + step]) * p.slot_rows +
                                          v * p.view_rows + p.tok_off + k;
                    src[i] = p.fts + row * D + u * 8;
                }
            }
        };
        bool have = wk.next(t);
        if (grp == 1 && have) have = wk.next(t);           // group 1 starts at tile 1
        const __half* src[8];
        if (have) resolve(t, src);
        int n = 0;                                          // tiles this group has filled
        long long w_empty = 0, w_land = 0;
        const long long t_begin = clock64();
        while (have) {
            const long long c0 = clock64();
            mbar_wait(&a_empty[grp], (n & 1) ^ 1);
            w_empty += clock64() - c0;
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const uint32_t dst = smem_u32(tile_base + k * L::A_CHUNK);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (src[i]) cp_async_16(dst + sw128_offset(r0 + 8 * i, u), src[i] + k * 64);
            }
            cp_async_commit();
            const long long c1 = clock64();
            cp_async_wait<0>();
            fence_proxy_async_smem();
            mbar_arrive(&a_full[grp]);
            w_land += clock64() - c1;
            // this group's next tile (two tiles ahead in the CTA's sequence): resolve its rows and prefetch them into L2
            // while the consumers work on the tile just published
            Tile tn;
            have = wk.next(tn) && wk.next(tn);
            const __half* nsrc[8];
            if (have) {
                resolve(tn, nsrc);
                if (u == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (nsrc[i]) l2_prefetch_bulk(nsrc[i], D * 2);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) src[i] = nsrc[i];
            t = tn;
            ++n;
        }
        if (p.dbg && gt == 0) {
            long long* d = p.dbg + blockIdx.x * 16 + grp * 3;
            d[0] = clock64() - t_begin; d[1] = w_empty; d[2] = w_land;
        }
    } else if (warp == POOL_MMA_WARP) {
        // ------------------------------------------------------------ MMA issuer: S^T = text (TMEM) x tile^T (smem)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, POOL_ROWS);
            int it = 0, cur_b = -1, visits = 0;
            long long w_afull = 0, w_dempty = 0;
            const long long t_begin = clock64();
            while (wk.next(t)) {
                const int buf = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                if (t.b != cur_b) {
                    mbar_wait(t_ready, visits & 1);
                    cur_b = t.b;
                    ++visits;
                }
                const long long c0 = clock64();
                mbar_wait(&d_empty[buf], ph ^ 1);
                const long long c1 = clock64();
                mbar_wait(&a_full[buf], ph);
                w_dempty += c1 - c0;
                w_afull += clock64() - c1;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + D_COL0 + buf * POOL_ROWS;
                const uint32_t tile_s = smem_u32(sA + buf * L::A_BYTES);
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const uint64_t db = umma_desc_sw128_kmajor(tile_s + k * L::A_CHUNK);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)   // K = 16 per instruction = 8 TMEM columns of A, 32 bytes of B
                        umma_f16_ts(d_tmem, tmem_base + k * 32 + kk * 8, db + 2 * kk, idesc, (k | kk) ? 1u : 0u);
                }
                umma_commit(&d_full[buf]);
                ++it;
            }
            if (p.dbg) {
                long long* d = p.dbg + blockIdx.x * 16 + 6;
                d[0] = clock64() - t_begin; d[1] = w_afull; d[2] = w_dempty;
            }
        }
    } else if (warp < POOL_GATHER_WARP0) {
        // ------------------------------------------------------------ text operand -> TMEM, relevance max, softmax weights
        const int q = warp;                         // TMEM lane quadrant
        const int e = tid;                          // 0..127 = TMEM lane = text position
        int it = 0, cur_b = -1;
        if (e == 0) { s_scal[2] = 0.0f; s_scal[3] = 0.0f; }
        long long w_dfull = 0, c_text = 0, c_red = 0;
        const long long t_begin = clock64();
        while (wk.next(t)) {
            const int buf = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const long long c0 = clock64();
            if (t.b != cur_b) {
                // Every MMA that read the previous episode's text has retired: this warp waited on d_full of the previous
                // tile below.  Lane-major workspace: unit c of this lane's text position at ((b*D/8 + c)*128 + lane)*16.
                cur_b = t.b;
                const uint4* src = p.text_ws + static_cast<size_t>(t.b) * (D / 8) * 128 + e;
                const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
                for (int c0u = 0; c0u < D / 8; c0u += 16) {     // 16 coalesced 16-byte loads in flight, then 8 stores of 32 B
                    uint4 vv[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) vv[c] = __ldg(src + static_cast<size_t>(c0u + c) * 128);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t v8[8] = {vv[2 * c].x, vv[2 * c].y, vv[2 * c].z, vv[2 * c].w,
                                                vv[2 * c + 1].x, vv[2 * c + 1].y, vv[2 * c + 1].z, vv[2 * c + 1].w};
                        tmem_st_32x32b_x8(ta + (c0u / 2 + c) * 8, v8);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(t_ready);
                for (int i = e; i <= n_cells; i += 128) s_cs[i] = p.cell_start[t.b * (n_cells + 1) + i];
                for (int i = e; i < n_cells; i += 128) s_cr[i] = p.cell_rank[t.b * n_cells + i];
            }
            const long long c1 = clock64();
            c_text += c1 - c0;
            mbar_wait(&d_full[buf], ph);
            const long long c2 = clock64();
            w_dfull += c2 - c1;
            tc_fence_after();
            float v[64];
            {
                const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + D_COL0 + buf * POOL_ROWS;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(ta + c * 16, r);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[c * 16 + j] = __uint_as_float(r[j]);
                }
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[buf]);
            lane_max_level<64>(v, lane, 16);
            lane_max_level<32>(v, lane, 8);
            lane_max_level<16>(v, lane, 4);
            lane_max_level<8>(v, lane, 2);
            lane_max_level<4>(v, lane, 1);
            s_part[q * POOL_ROWS + 2 * lane] = v[0];
            s_part[q * POOL_ROWS + 2 * lane + 1] = v[1];
            named_bar_sync(1, 128);
            if (e < POOL_ROWS)
                s_w[buf * POOL_ROWS + e] = fmaxf(fmaxf(s_part[e], s_part[POOL_ROWS + e]),
                                                 fmaxf(s_part[2 * POOL_ROWS + e], s_part[3 * POOL_ROWS + e]));
            named_bar_sync(2, 128);
            c_red += clock64() - c2;
            // --- softmax weights of this tile's rows: one thread per row (warps 0 and 1), per-cell max and sum by
            //     warp-segmented scans (cells are contiguous row segments), one shared-memory hop for a cell that
            //     straddles rows 31|32; the cells of the tile are emitted as a compact segment list for the pooling warps
            float m_carry = s_scal[2], s_carry = s_scal[3];
            float new_m = 0.0f, new_s = 0.0f;
            bool writes_carry = false;
            if (warp < 2 && !(p.mode & 2)) {
                const bool valid = e < t.nrows;
                const float* wrow = s_w + buf * POOL_ROWS;
                int c_lo = 0, c_hi = 0, cellid = 0, seg_lo = e, seg_hi = e + 1;
                if (valid) {
                    const int P = t.pos + e;
                    int a = 0, c = n_cells;        // last cell with cs[cell] <= P
                    while (c - a > 1) {
                        const int mid = (a + c) >> 1;
                        if (s_cs[mid] <= P) a = mid; else c = mid;
                    }
                    cellid = a;
                    c_lo = s_cs[cellid]; c_hi = s_cs[cellid + 1];
                    seg_lo = max(c_lo, t.pos) - t.pos;
                    seg_hi = min(c_hi, t.pos + t.nrows) - t.pos;
                }
                const int w0 = warp * 32;
                const int lo_w = max(seg_lo, w0) - w0, hi_w = min(seg_hi, w0 + 32) - w0;     // this warp's part, in lanes
                const bool spans = valid && seg_lo < 32 && seg_hi > 32;
                const float wr = valid ? wrow[e] : -INFINITY;
                // segment max
                float m = wr;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float v2 = __shfl_up_sync(0xffffffffu, m, o);
                    if (lane - o >= lo_w) m = fmaxf(m, v2);
                }
                m = __shfl_sync(0xffffffffu, m, hi_w - 1);
                // index of this row's segment inside the tile = number of segment heads before it
                const bool head = valid && (e == seg_lo);
                const unsigned heads = __ballot_sync(0xffffffffu, head);
                if (lane == (warp == 0 ? 31 : 0)) s_part[warp] = m;          // s_part is free again after barrier 2
                if (lane == 0) reinterpret_cast<int*>(s_part)[8 + warp] = __popc(heads);
                named_bar_sync(4, 64);
                if (spans) m = fmaxf(s_part[0], s_part[1]);
                // (rows of a cell that started in warp 0 see no head before them in warp 1 and land on warp 0's last segment)
                const int seg_idx = __popc(heads & ((2u << lane) - 1u)) - 1 + (warp == 1 ? reinterpret_cast<int*>(s_part)[8] : 0);
                const bool continues = valid && (c_lo < t.pos);
                float cscale = 1.0f, s0 = 0.0f;
                if (continues) {
                    const float mn = fmaxf(m, m_carry);
                    cscale = expf(m_carry - mn);
                    s0 = s_carry * cscale;
                    m = mn;
                }
                const float my_p = valid ? expf(wr - m) : 0.0f;
                // segment sum
                float sm = my_p;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float v2 = __shfl_up_sync(0xffffffffu, sm, o);
                    if (lane - o >= lo_w) sm += v2;
                }
                sm = __shfl_sync(0xffffffffu, sm, hi_w - 1);
                if (lane == (warp == 0 ? 31 : 0)) s_part[2 + warp] = sm;
                named_bar_sync(5, 64);
                if (spans) sm = s_part[2] + s_part[3];
                const float ssum = s0 + sm;
                if (valid) {
                    if (e == seg_hi - 1) {       // last row of the segment publishes it
                        const int k = seg_idx;
                        float fin = 0.0f;
                        int rank = 0;
                        if (c_hi <= t.pos + t.nrows) { fin = 1.0f / ssum; rank = s_cr[cellid]; }
                        else { writes_carry = true; new_m = m; new_s = ssum; }
                        s_send[buf * POOL_ROWS + k] = seg_hi;
                        s_sfin[buf * POOL_ROWS + k] = fin;
                        s_srank[buf * POOL_ROWS + k] = rank;
                        if (seg_hi == t.nrows) s_nseg[buf] = k + 1;
                    }
                    if (e == 0) s_scal[buf] = continues ? cscale : 1.0f;
                    if (p.w_out) p.w_out[static_cast<size_t>(t.b) * p.cap + t.pos + e] = wr;
                }
                s_p[buf * POOL_ROWS + e] = my_p;
            } else if (warp < 2) {
                if (e == 0) { s_nseg[buf] = 0; s_scal[buf] = 1.0f; }
            }
            named_bar_sync(3, 128);             // everyone has read the old carry / s_cs / s_part
            if (writes_carry) { s_scal[2] = new_m; s_scal[3] = new_s; }
            mbar_arrive(&p_full[buf]);          // release: s_p / segment list / s_scal[buf] are visible to the pooling warps
            ++it;
        }
        if (p.dbg && e == 0) {
            long long* d = p.dbg + blockIdx.x * 16 + 10;
            d[0] = clock64() - t_begin; d[1] = w_dfull; d[2] = c_text; p.dbg[blockIdx.x * 16 + 9] = c_red;
        }
    } else {
        // ------------------------------------------------------------ weighted sums from the resident tile
        const int pt = tid - POOL_POOL_WARP0 * 32;     // owns columns 4*pt .. 4*pt+3
        const int k = pt >> 4;                         // chunk of those columns
        const int u = (pt & 15) >> 1;
        const int sub = (pt & 1) * 8;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        int it = 0;
        long long w_pfull = 0, c_loop = 0;
        const long long t_begin = clock64();
        while (wk.next(t)) {
            const int buf = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const long long c0 = clock64();
            mbar_wait(&p_full[buf], ph);
            w_pfull += clock64() - c0;
            mbar_wait(&a_full[buf], ph);               // already complete; orders the cp.async writes before our reads
            const uint32_t chunk_s = smem_u32(sA + buf * L::A_BYTES + k * L::A_CHUNK) + sub;
            const float cs = s_scal[buf];
            acc0 *= cs; acc1 *= cs; acc2 *= cs; acc3 *= cs;
            const float* pp = s_p + buf * POOL_ROWS;
            __half* out_b = p.pooled + static_cast<size_t>(t.b) * n_cells * D + pt * 4;
            const long long c_loop0 = clock64();
            const int nseg = (p.mode & 1) ? 0 : s_nseg[buf];
            int r = 0;
            for (int sg = 0; sg < nseg; ++sg) {
                const int end = s_send[buf * POOL_ROWS + sg];
                // rows of one cell: branch-free inner loop, 4 rows per round (loads first, then the FMA chain)
                for (; r + 4 <= end; r += 4) {
                    uint2 raw[4];
                    float w4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(raw[i].x), "=r"(raw[i].y) : "r"(chunk_s + sw128_offset(r + i, u)));
                        w4[i] = pp[r + i];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 x01 = __half22float2(*reinterpret_cast<const __half2*>(&raw[i].x));
                        const float2 x23 = __half22float2(*reinterpret_cast<const __half2*>(&raw[i].y));
                        acc0 = fmaf(w4[i], x01.x, acc0); acc1 = fmaf(w4[i], x01.y, acc1);
                        acc2 = fmaf(w4[i], x23.x, acc2); acc3 = fmaf(w4[i], x23.y, acc3);
                    }
                }
                for (; r < end; ++r) {
                    uint2 raw;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "r"(chunk_s + sw128_offset(r, u)));
                    const float w = pp[r];
                    const float2 x01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                    const float2 x23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                    acc0 = fmaf(w, x01.x, acc0); acc1 = fmaf(w, x01.y, acc1);
                    acc2 = fmaf(w, x23.x, acc2); acc3 = fmaf(w, x23.y, acc3);
                }
                const float fin = s_sfin[buf * POOL_ROWS + sg];
                if (fin != 0.0f) {      // the cell is complete: normalise, store, reset (else it is carried into the next tile)
                    const __half2 h01 = __floats2half2_rn(acc0 * fin, acc1 * fin);
                    const __half2 h23 = __floats2half2_rn(acc2 * fin, acc3 * fin);
                    uint2 o;
                    o.x = *reinterpret_cast<const uint32_t*>(&h01);
                    o.y = *reinterpret_cast<const uint32_t*>(&h23);
                    *reinterpret_cast<uint2*>(out_b + static_cast<size_t>(s_srank[buf * POOL_ROWS + sg]) * D) = o;
                    acc0 = acc1 = acc2 = acc3 = 0.f;
                }
            }
            c_loop += clock64() - c_loop0;
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_empty[buf]);
            ++it;
        }
        if (p.dbg && pt == 0) {
            long long* d = p.dbg + blockIdx.x * 16 + 13;
            d[0] = clock64() - t_begin; d[1] = w_pfull; d[2] = c_loop;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == POOL_MMA_WARP) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, POOL_TMEM_COLS);
    }
}

}  // namespace gmm

static long long* g_pool_dbg = nullptr;
static int g_pool_mode = 0;
extern "C" void gridmm_debug_set_pool_mode(int mode) { g_pool_mode = mode; }
// Debug hook (tools/microbench.py): per-CTA cycle counters [grid][16] written by the next pool launches; null disables.
extern "C" void gridmm_debug_set_pool_counters(long long* dbg) { g_pool_dbg = dbg; }

extern "C" int gridmm_pool(const void* fts, int feat_dim, const int* slots, int t_cap, int slot_rows, int view_rows,
                           int tok_off, const int* perm, int cap, const int* cell_start, const int* cell_rank, int n_cells,
                           const void* text_fts, int l_pad, int batch, void* text_ws, void* pooled, float* w_out, int num_ctas,
                           cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (!fts || !slots || !perm || !cell_start || !cell_rank || !text_fts || !text_ws || !pooled) return GRIDMM_ERR_ARG;
    if (batch > POOL_MAX_BATCH || n_cells > POOL_MAX_CELLS || l_pad < 1 || l_pad > 128) return GRIDMM_ERR_SHAPE;
    if (feat_dim != 768 && feat_dim != 512) return GRIDMM_ERR_SHAPE;
    if ((reinterpret_cast<uintptr_t>(text_fts) & 15) || (reinterpret_cast<uintptr_t>(text_ws) & 15)) return GRIDMM_ERR_SHAPE;
    PoolParams p;
    p.fts = reinterpret_cast<const __half*>(fts); p.slots = slots; p.perm = perm; p.cell_start = cell_start;
    p.cell_rank = cell_rank; p.text = reinterpret_cast<const __half*>(text_fts);
    p.text_ws = reinterpret_cast<const uint4*>(text_ws);
    p.pooled = reinterpret_cast<__half*>(pooled); p.w_out = w_out;
    p.batch = batch; p.t_cap = t_cap; p.cap = cap; p.n_cells = n_cells; p.l_pad = l_pad;
    p.slot_rows = slot_rows; p.view_rows = view_rows; p.tok_off = tok_off; p.dbg = g_pool_dbg; p.mode = g_pool_mode;
    int dev = 0, sms = 0;
    GMM_CUDA_CHECK(cudaGetDevice(&dev));
    GMM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = num_ctas > 0 ? num_ctas : sms;
    if (feat_dim == 768) {
        constexpr int smem = PoolSmem<768>::TOTAL;
        GMM_CUDA_CHECK(cudaFuncSetAttribute(pool_kernel<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GMM_CUDA_CHECK(launch_pdl(text_to_lanes_kernel<768>, dim3(768 / 8, batch), dim3(128), 0, stream, p.text,
                                  reinterpret_cast<uint4*>(text_ws), l_pad));
        GMM_CUDA_CHECK(launch_pdl(pool_kernel<768>, dim3(grid), dim3(POOL_FIXED_THREADS + 768 / 4), smem, stream, p));
    } else {
        constexpr int smem = PoolSmem<512>::TOTAL;
        GMM_CUDA_CHECK(cudaFuncSetAttribute(pool_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GMM_CUDA_CHECK(launch_pdl(text_to_lanes_kernel<512>, dim3(512 / 8, batch), dim3(128), 0, stream, p.text,
                                  reinterpret_cast<uint4*>(text_ws), l_pad));
        GMM_CUDA_CHECK(launch_pdl(pool_kernel<512>, dim3(grid), dim3(POOL_FIXED_THREADS + 512 / 4), smem, stream, p));
    }
    gridmm_count_launch(2);
    return 0;
}
