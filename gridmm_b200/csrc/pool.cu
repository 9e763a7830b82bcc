// Instruction-relevance pooling: the reference's per-episode loop
//   grid_fts_weight = (grid_fts @ text_fts).max(-1)            map_nav_src/models/vilmodel.py:797-798
//   for i in range(196): softmax over the points of cell i, weighted sum   vilmodel.py:801-807
// as ONE persistent kernel that reads every valid patch-feature row from HBM exactly once.
//
//   * pool_plan_kernel (one small launch, off the critical path: it only needs the sorted cells) cuts the batch's sorted valid
//     rows into one contiguous range per CTA, equal in cost (a row = 1, an episode switch = POOL_EPISODE_COST).  A cut may fall
//     inside a large cell: such a cell is pooled in pieces, each piece leaves (max, weight sum, fp32 sums) in a workspace and the
//     last piece to arrive merges them in CTA order, per 128-column slice (no atomics on data, deterministic result);
//   * rows are streamed in cell-sorted order (gridmm_grid_update produced `perm`), 32 rows per tile, through a ring of
//     four 48 KB shared-memory tiles.  Four producer warps resolve the tile's slab rows (perm entry requested two tiles ahead,
//     the episode's slot table in registers) and fetch them with TMA tile::gather4 (cp.async.bulk.tensor.2d...tile::gather4:
//     four arbitrary rows x 64 fp16 columns per instruction, hardware SWIZZLE_128B, completion on an mbarrier) straight into a
//     K-major operand tile (D/64 chunks of 32 x 128 B).  Row indices are made warp-uniform with shuffles and one elected lane
//     issues the copies back to back (elect_one(), common.cuh: per-lane operands cost a ~100-cycle R2UR waterfall per gather4);
//   * relevance  S^T[L, 32] = text_fts[L, D] . X_tile^T  on tcgen05 with the A operand (text_fts of the current
//     episode, up to 128 positions) held in TENSOR MEMORY for the whole episode (tcgen05.st by 8 warps, 12-16 loads in flight
//     per thread), B operand = the feature tile in shared memory, one TMEM accumulator per ring slot;
//     w = max over ALL text positions (padding included -- vilmodel.py:798) = max over TMEM lanes, taken with
//     redux.sync.max.f32 (32 independent warp-wide maxima) by the reducer warps whose lane quadrant holds real text positions,
//     merged across quadrants through shared memory + an mbarrier by the reducer warp that owns the padding quadrant;
//   * that warp also turns w into per-cell softmax numerators exp(w - cell max) (binary search of the row's cell ahead of the
//     relevance wait, match.any + redux.max inside the warp, running max carried across tiles);
//   * a builder warp turns the numerators into the mma.sync B fragments of the pooling warps, keeps the running weight sum of
//     the open cells and publishes which cells complete and their normalisers (once, instead of once per pooling warp);
//   * weighted sums as warp-level HMMA from the resident tile: out[cell, :] = sum_r p[r] x[r, :] is the product
//     X^T[D, 32 rows] . W[32 rows, 8 cell slots] (slot = rank & 7; an open cell keeps its slot across tiles), A through
//     ldmatrix.trans, fp32 accumulators in registers across tiles, rescaled when a later tile raises the open cell's max;
//   * grid_proj is applied AFTER pooling by the GEMM kernel (sum_j p_j (W x_j + b) = W (sum_j p_j x_j) + b), so this
//     kernel emits the pooled raw feature per non-empty cell, compacted in ascending cell order (the order
//     vilmodel.py:819 gathers them in), as fp16 GEMM input.
//
// HBM roofline: algorithmic bytes = valid_rows * D * 2 (+ L*D*2 per episode + outputs); see DESIGN.md.
#include "common.cuh"
#include "host_util.h"

namespace gmm {

constexpr int POOL_ROWS = 32;          // rows per tile (= N of the relevance MMA)
constexpr int POOL_NBUF = 4;           // tile buffers / TMEM accumulators in flight
constexpr int POOL_PROD_WARPS = 4;     // warps 0..3: TMA gather producers (8 rows of every tile each) + text staging, first half
constexpr int POOL_RED_WARP0 = 4;      // warps 4..7: relevance max (TMEM lane quadrant = warp % 4), softmax weights, text staging
constexpr int POOL_MMA_WARP = 8;       // tcgen05.mma issuer, owns the TMEM allocation
constexpr int POOL_POOL_WARP0 = 9;     // warps 9 .. 9 + D/128 - 1: weighted sums (mma.sync), one warp per 128 feature columns;
                                       // warp 9 + D/128: builder of their weight operands
constexpr int POOL_FIXED_THREADS = (POOL_POOL_WARP0 + 1) * 32;
constexpr int POOL_MAXPASS = 4;        // a 32-row tile holds at most 32 cells = 4 passes of 8 cell slots
constexpr int POOL_SLICE = 136;        // floats of one 128-column slice of a partial: 128 sums, weight sum, max, padding
constexpr int POOL_MAX_BATCH = 1024;
constexpr int POOL_MAX_CTAS = 1024;
constexpr int POOL_MAX_CELLS = 256;
constexpr int POOL_TMEM_COLS = 512;
constexpr int POOL_PTS = 588, POOL_VIEW_PTS = 49;   // points per viewpoint / per view (r2r/env.py:279-289)
constexpr int POOL_EPISODE_COST = 200;              // rows' worth of time an episode switch costs a CTA (text staging, partial tiles)
constexpr int POOL_SNAP = 16;                       // a CTA boundary within this many rows of a cell boundary moves onto it

// Work plan of one launch (gridmm_pool_plan -> workspace, read by every CTA of pool_kernel): the sorted valid rows of the whole
// batch are cut into one contiguous range per CTA, equal in COST units (a row = 1, an episode start = `episode_cost`).  A cut
// either sits on a cell boundary or -- when the nearest boundary is more than POOL_SNAP rows away -- in the middle of a cell.  A
// cell cut that way is pooled in pieces: every piece leaves its un-normalised partial (max, weight sum, fp32 sums) in the
// workspace, the LAST piece to arrive (an atomic counter per cell chain and 128-column slice: every pooling warp runs the
// protocol on its own columns, no CTA-wide barrier) merges all of them in CTA order, so the result does not depend on the
// arrival order.  The CTA that holds the START of a cut cell reaches it at the end of its range, when the later pieces (first in
// their CTAs' ranges) have normally arrived long ago: it then merges straight from its registers.
struct PoolPlanCta {
    int g0, g1;              // rows [g0, g1) of the global sorted-valid order
    int head_first, head_last, head_n;   // chain of CTAs that share this CTA's FIRST cell (when its start cut is mid-cell)
    int tail_last, tail_n;   // chain [this CTA, tail_last] sharing its LAST cell (when its end cut is mid-cell)
    int flags;               // 1: first cell is a piece, 2: last cell is a piece, 4: the whole range lies inside one cell
};
// workspace layout (ints): PoolPlanCta[G] | vbase[batch + 1] | arrival counters[G][8] | (16-byte aligned)
//                           float partial[G][2 = head piece, tail piece][D / 128][POOL_SLICE]
__host__ __device__ inline size_t pool_ws_part_offset(int batch, int G) {
    return ((static_cast<size_t>(G) * 8 + batch + 1 + static_cast<size_t>(G) * 8) * 4 + 15) / 16 * 16;
}
__host__ __device__ inline size_t pool_ws_bytes(int batch, int D, int G) {
    return pool_ws_part_offset(batch, G) + static_cast<size_t>(G) * 2 * (D / 128) * POOL_SLICE * 4;
}

struct PoolParams {
    const int* slots;        // [B, t_cap]   slab slot of (episode, step)
    const int* perm;         // [B, cap]     valid point indices sorted by cell
    const int* cell_start;   // [B, n_cells + 1]
    const int* cell_rank;    // [B, n_cells]
    const uint4* text_ws;    // [B, D/8, 128] lane-major copy of text_fts (16-byte units; positions < l_pad are read)
    __half* pooled;          // [B, n_cells, D]  compacted by cell rank
    float* w_out;            // [B, cap] relevance weight per sorted position, or null (tests; the first pass of a long text)
    const float* w_in;       // [B, cap] row maxima over the text positions an EARLIER launch covered (merged into w), or null
    int max_only;            // 1: only w_out is produced (first pass over a text longer than 128 positions): no softmax, no sums
    int split_weights;       // mma.sync sums: 1 = every softmax weight enters as fp16 value + fp16 residual (two MMAs), 0 = fp16 value only
    int* plan;               // workspace written by pool_plan_kernel (layout above)
    int batch, t_cap, cap, n_cells;
    int l_pad;               // text positions of THIS launch (<= 128: one per tensor-memory lane)
    int slot_rows, view_rows, tok_off;   // row = slot*slot_rows + view*view_rows + tok_off + patch
    long long* dbg;          // optional [grid][16] cycle counters (tools/microbench2.py), null in production
    long long* trace;        // optional [4 CTAs][64 tiles][8] clock64 stamps of the first tiles' stage hand-overs (tools/pool_probe.py)
    int exp;                 // timing experiments (tools/pool_probe.py; results are garbage): 1 = fetch half of every row, 2 = always the same 32 rows (L2 hits), 3 / 4 = half / a twelfth of the relevance contraction
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

// TMA tile::gather4: rows r0..r3 (arbitrary), columns [c0, c0 + box) of a 2D tensor -> 4 consecutive 128-byte rows at
// smem_dst (hardware SWIZZLE_128B on the shared-memory address), completion counted on `bar`.
__device__ __forceinline__ void tma_gather4(uint32_t smem_dst, const CUtensorMap* tm, int c0, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}

// mbarrier wait that traps instead of hanging the GPU if a phase never completes (a wrong transaction count would
// otherwise spin forever): each failed try_wait parks the thread for up to the suspend-time hint
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1 << 22)) __trap();
    }
}

// same, with a sleep between polls: for waiters with slack (the gather producers run several tiles ahead), so that their
// polling does not take issue slots from the pooling warps
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(128);
        if (++spins > (1 << 22)) __trap();
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// warp-wide maximum of a float, returned to every lane (sm_100a: CREDUX.MAX.F32)
__device__ __forceinline__ float redux_max_f32(float x) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// order-preserving float <-> uint32 map (for redux.sync max over the rows of one cell)
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

template <int D>
struct PoolSmem {
    static constexpr int CH = D / 64;
    static constexpr int A_CHUNK = POOL_ROWS * 128;
    static constexpr int A_BYTES = CH * A_CHUNK;                  // one feature tile
    static constexpr int MISC_BYTES = 64 * 8                      // barriers + tmem slot
                                      + POOL_NBUF * POOL_ROWS * 4 // softmax numerator per row (per buffer)
                                      + POOL_NBUF * POOL_ROWS * 4 // cell id per row (per buffer)
                                      + POOL_NBUF * 4 * POOL_ROWS * 4   // per-quadrant partial maxima (per buffer)
                                      + 128                       // scalars
                                      + 2 * (POOL_MAX_CELLS + 1) * 4 + POOL_MAX_CELLS * 4   // cell_start (reducers / poolers), cell_rank
                                      + (POOL_MAX_BATCH + 1) * 4  // vbase
                                      + POOL_NBUF * POOL_MAXPASS * (8 + 1) * 4;    // normalisers per cell slot, pass word
    static constexpr int W_BYTES = POOL_NBUF * POOL_MAXPASS * 32 * 32;      // weight fragments of the pooling MMAs: 32 B per lane and pass
    static constexpr int TOTAL = 1024 + POOL_NBUF * A_BYTES + W_BYTES + MISC_BYTES + 64;
};

// text_fts [B, l_pad, D] -> lane-major copy [B, D/8, 128] of 16-byte units: unit c of text position t sits at
// ((b * D/8 + c) * 128 + t) * 16 bytes, so the 32 lanes of a warp (32 consecutive positions) read 512 contiguous bytes
// when the operand is moved into tensor memory.  (The navigation forward skips this kernel: the text_proj GEMM writes
// this layout directly from its epilogue, gridmm_linear_f16_lanes.)
template <int D>
__global__ void __launch_bounds__(128) text_to_lanes_kernel(const __half* text, uint4* ws, int l_pad) {
    pdl_wait();
    // positions 128 z .. 128 z + 127 go to the z-th [B, D/8, 128] block of the workspace (texts longer than 128 positions)
    const int b = blockIdx.y, c = blockIdx.x, t = blockIdx.z * 128 + threadIdx.x;
    if (t >= l_pad) return;
    const uint4 v = *reinterpret_cast<const uint4*>(text + (static_cast<size_t>(b) * l_pad + t) * D + c * 8);
    ws[(static_cast<size_t>(blockIdx.z) * gridDim.y + b) * (D / 8) * 128 + static_cast<size_t>(c) * 128 + threadIdx.x] = v;
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

// Move this thread's share of the episode's text operand into tensor memory: TMEM lane `tlane` (= text position, lanes
// >= l_pad replicate position 0 so that they never change the max), 16-byte units [u0, u0 + D / 16), 12 or 16 loads in flight
// per thread (the copy is bound by the L2 round trip, not by bandwidth: 8 batches of 6 loads took 6-8 k cycles per episode).
template <int D>
__device__ __forceinline__ void stage_text(const uint4* ws_b, int tlane, int l_pad, uint32_t taddr_lane, int u0) {
    constexpr int NU = D / 16, BU = (NU % 12 == 0) ? 12 : 16;
    static_assert(NU % BU == 0 && BU % 2 == 0, "half a text row is a whole number of batches");
    const uint4* src = ws_b + (tlane < l_pad ? tlane : 0);
#pragma unroll 1
    for (int ub = 0; ub < NU; ub += BU) {
        uint4 v[BU];
#pragma unroll
        for (int c = 0; c < BU; ++c) v[c] = __ldg(src + static_cast<size_t>(u0 + ub + c) * 128);
#pragma unroll
        for (int c = 0; c < BU / 2; ++c) {
            const uint32_t v8[8] = {v[2 * c].x, v[2 * c].y, v[2 * c].z, v[2 * c].w, v[2 * c + 1].x, v[2 * c + 1].y, v[2 * c + 1].z, v[2 * c + 1].w};
            // unit u holds fp16 elements 8u..8u+7 = TMEM columns 4u..4u+3
            tmem_st_32x32b_x8(taddr_lane + (u0 + ub + 2 * c) * 4, v8);
        }
    }
    tmem_st_wait();
}

// ---------------------------------------------------------------------------------------------------- work plan
// One CTA of 1024 threads: prefix of the valid-row counts, the G + 1 cuts (four lanes per cut, each loads a quarter of the
// episode's cell_start row), then per CTA the chains of its first / last cell (see PoolPlanCta).
__global__ void __launch_bounds__(1024) pool_plan_kernel(const int* __restrict__ cell_start, int n_cells, int batch, int G,
                                                         int episode_cost, int snap, int* __restrict__ ws) {
    __shared__ int s_vb[POOL_MAX_BATCH + 1];
    __shared__ int s_g[POOL_MAX_CTAS + 1], s_clo[POOL_MAX_CTAS + 1], s_chi[POOL_MAX_CTAS + 1];
    pdl_wait();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < batch; i += blockDim.x) s_vb[i + 1] = cell_start[i * (n_cells + 1) + n_cells];
    if (tid == 0) s_vb[0] = 0;
    __syncthreads();
    if (warp == 0) {
        const int per = (batch + 31) / 32;
        const int lo = min(lane * per, batch), hi = min(lo + per, batch);
        int sum = 0;
        for (int i = lo; i < hi; ++i) sum += s_vb[i + 1];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - sum;
        for (int i = lo; i < hi; ++i) { run += s_vb[i + 1]; s_vb[i + 1] = run; }
    }
    __syncthreads();
    const int total = s_vb[batch];
    const long long total_c = static_cast<long long>(total) + static_cast<long long>(batch) * episode_cost;
    // Four lanes per cut (256 cuts per pass of the block): every lane requests its quarter of the episode's cell_start row at once,
    // so a pass costs ONE round trip to L2 (a warp per cut took five dependent passes for 149 cuts, ~10 us in the step).
    for (int c0 = 0; c0 <= G; c0 += 256) {
        const int c = c0 + (tid >> 2), sub = tid & 3;
        // cut c in COST units: one per valid row plus `episode_cost` per episode start (moving a new text operand into tensor
        // memory and the partial tiles around an episode switch), so a CTA whose range crosses an episode boundary gets fewer rows
        const long long tgt = (static_cast<long long>(min(c, G)) * total_c) / G;
        const bool interior = c < G && tgt > 0;
        int lo = 0, hi = batch;            // largest b with vbase[b] + b * COST <= tgt
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (static_cast<long long>(s_vb[mid]) + static_cast<long long>(mid) * episode_cost <= tgt) lo = mid; else hi = mid;
        }
        const int nv = s_vb[lo + 1] - s_vb[lo];
        const long long lc = tgt - s_vb[lo] - static_cast<long long>(lo) * episode_cost - episode_cost;
        const int local = lc <= 0 ? 0 : (lc >= nv ? nv : static_cast<int>(lc));
        const int* cs = cell_start + lo * (n_cells + 1);
        int lb = 0, ub = nv;               // cell boundaries around the cut: lb = largest <= local, ub = smallest > local
        if (interior) {
            for (int i = sub; i <= n_cells; i += 4) {
                const int v = __ldg(cs + i);
                if (v <= local) lb = max(lb, v); else ub = min(ub, v);
            }
        }
#pragma unroll
        for (int o = 2; o > 0; o >>= 1) {
            lb = max(lb, __shfl_xor_sync(0xffffffffu, lb, o));
            ub = min(ub, __shfl_xor_sync(0xffffffffu, ub, o));
        }
        int g, clo = -1, chi = -1;
        if (c >= G) g = total;
        else if (tgt <= 0) g = 0;
        else {
            int cut = local;
            if (local >= nv) cut = nv;
            else if (local - lb <= snap && local - lb <= ub - local) cut = lb;
            else if (ub - local <= snap) cut = ub;
            if (cut > lb && cut < ub) { clo = s_vb[lo] + lb; chi = s_vb[lo] + ub; }      // mid-cell: the cell is rows [clo, chi)
            g = s_vb[lo] + cut;
        }
        if (sub == 0 && c <= G) { s_g[c] = g; s_clo[c] = clo; s_chi[c] = chi; }
    }
    __syncthreads();
    int* vb_out = ws + G * 8;
    int* cnt = vb_out + batch + 1;
    for (int i = tid; i <= batch; i += blockDim.x) vb_out[i] = s_vb[i];
    for (int c = tid; c < G; c += blockDim.x) {
        PoolPlanCta r;
        r.g0 = s_g[c]; r.g1 = s_g[c + 1];
        r.head_first = r.head_last = c; r.head_n = 0; r.tail_last = c; r.tail_n = 0; r.flags = 0;
        const bool nonempty = r.g1 > r.g0;
        const bool head = nonempty && s_clo[c] >= 0;
        const bool tail = nonempty && s_clo[c + 1] >= 0;
        const bool middle = head && tail && s_clo[c] == s_clo[c + 1];
        if (head) {
            const int cs_ = s_clo[c], ce_ = s_chi[c];
            int f = c - 1;
            while (s_g[f] > cs_) --f;                      // the CTA that holds the cell's first row (s_g[0] = 0 <= cs_)
            int l = c;
            while (s_g[l + 1] < ce_) ++l;                  // ... and its last row (s_g[G] = total >= ce_)
            int n = 0;
            for (int i = f; i <= l; ++i) n += s_g[i + 1] > s_g[i];
            r.head_first = f; r.head_last = l; r.head_n = n; r.flags |= 1;
        }
        if (tail && !middle) {
            const int ce_ = s_chi[c + 1];
            int l = c + 1;
            while (s_g[l + 1] < ce_) ++l;
            int n = 0;
            for (int i = c; i <= l; ++i) n += s_g[i + 1] > s_g[i];
            r.tail_last = l; r.tail_n = n; r.flags |= 2;
        }
        if (middle) r.flags |= 4;
        int4* o = reinterpret_cast<int4*>(ws + c * 8);
        o[0] = make_int4(r.g0, r.g1, r.head_first, r.head_last);
        o[1] = make_int4(r.head_n, r.tail_last, r.tail_n, r.flags);
#pragma unroll
        for (int i = 0; i < 8; ++i) cnt[c * 8 + i] = 0;
    }
    pdl_launch_dependents();
}

#define POOL_TRACE(k) do { if (p.trace && blockIdx.x < 4 && it < 64 && lane == 0) p.trace[(blockIdx.x * 64 + it) * 8 + (k)] = clock64(); } while (0)

template <int D>
__global__ void __launch_bounds__(POOL_FIXED_THREADS + D / 4, 1)
pool_kernel(const __grid_constant__ CUtensorMap tm_fts, PoolParams p) {
    using L = PoolSmem<D>;
    constexpr int CH = L::CH;
    constexpr int NPW = D / 128;                  // pooling warps = 128-dim blocks of a feature row
    constexpr int A_COLS = D / 2;                 // TMEM columns of the text operand (two fp16 per column)
    constexpr int D_COL0 = A_COLS;                // relevance accumulators behind it: 32 columns per tile buffer
    constexpr int UNITS = D / 8;                  // 16-byte units per text position
    constexpr float LOG2E = 1.4426950408889634f;
    static_assert(A_COLS + POOL_NBUF * POOL_ROWS <= POOL_TMEM_COLS, "text operand + accumulators must fit in tensor memory");
    static_assert(CH <= 32 && POOL_ROWS == 8 * POOL_PROD_WARPS && (D == 512 || D == 768), "unsupported feature width");
    extern __shared__ uint8_t smem_raw[];
    // align by pointer arithmetic on the __shared__ array (an integer round-trip would demote every access below to a
    // generic LD/ST instead of LDS/STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                            // [NBUF][CH][32 x 128 B]
    uint8_t* sW = sA + POOL_NBUF * L::A_BYTES;     // [NBUF][MAXPASS][32 lanes][2 x 16 B]: mma.sync B fragments (fp16 weights, residuals)
    uint8_t* misc = sW + L::W_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(misc);          // [NBUF] tile landed (producer arrivals + TMA transaction bytes)
    uint64_t* a_empty = a_full + POOL_NBUF;                        // [NBUF] tile buffer drained by the pooling warps
    uint64_t* d_full = a_empty + POOL_NBUF;                        // [NBUF] accumulator written (tcgen05.commit)
    uint64_t* d_empty = d_full + POOL_NBUF;                        // [NBUF] accumulator read back
    uint64_t* p_full = d_empty + POOL_NBUF;                        // [NBUF] softmax numerators of the tile are in s_p
    uint64_t* m_full = p_full + POOL_NBUF;                         // [NBUF] partial maxima of reducer warps 1..3 are in s_part
    uint64_t* t_ready = m_full + POOL_NBUF;                        // [1] text operand of the episode is in TMEM
    uint64_t* ep_done = t_ready + 1;                               // [1] every MMA of the previous episode has retired
    uint64_t* w_full = ep_done + 1;                                // [NBUF] weight fragments / pass words of the tile are in sW
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 60);
    float* s_p = reinterpret_cast<float*>(misc + 64 * 8);          // [NBUF][32] exp(w - cell max) per row
    int* s_cid = reinterpret_cast<int*>(s_p + POOL_NBUF * POOL_ROWS);      // [NBUF][32] compact cell rank (+ last-row flag) of every row
    float* s_part = reinterpret_cast<float*>(s_cid + POOL_NBUF * POOL_ROWS);   // [NBUF][4][32]
    float* s_scal = s_part + POOL_NBUF * 4 * POOL_ROWS;            // [NBUF] rescale of the open cell per buffer
    float* s_carry_m = s_scal + POOL_NBUF;         // [1] running max of the cell left open by the previous tile
    int* s_carry_c = reinterpret_cast<int*>(s_carry_m + 1);        // [1] its (episode << 16 | cell) key (-1: none)
    float* s_mfirst = reinterpret_cast<float*>(s_carry_c + 1);     // [NBUF] softmax max of the cell of the tile's FIRST row
    float* s_mlast = s_mfirst + POOL_NBUF;                         // [NBUF] ... of its LAST row (partials of cells cut between CTAs)
    float* s_hsum = s_mlast + POOL_NBUF;                           // [NBUF] weight sum of the range's first cell when it completes as a piece
    float* s_tsum = s_hsum + POOL_NBUF;                            // [NBUF] weight sum of the cell the tile leaves open
    int* s_meta = reinterpret_cast<int*>(s_tsum + POOL_NBUF);      // [NBUF][2] first compact cell rank of the tile, number of passes
    int* s_csr = reinterpret_cast<int*>(misc + 64 * 8 + 2 * POOL_NBUF * POOL_ROWS * 4 + POOL_NBUF * 4 * POOL_ROWS * 4 + 128);   // reducers' cell_start
    int* s_cs = s_csr + POOL_MAX_CELLS + 1;                        // (spare table slot)
    int* s_cr = s_cs + POOL_MAX_CELLS + 1;                         // reducers' cell_rank [n_cells]
    int* s_vbase = s_cr + POOL_MAX_CELLS;                          // [batch + 1]
    float* s_fin = reinterpret_cast<float*>(s_vbase + POOL_MAX_BATCH + 1);      // [NBUF][MAXPASS][8] 1 / weight sum of the cells completed in a pass
    int* s_pass = reinterpret_cast<int*>(s_fin + POOL_NBUF * POOL_MAXPASS * 8); // [NBUF][MAXPASS] completed slots (bits 0..7) | (piece slot + 1) << 8

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n_cells = p.n_cells;
    const long long t_entry = p.dbg ? clock64() : 0;
    if (p.dbg && tid == 0) p.dbg[blockIdx.x * 16 + 14] = static_cast<long long>(globaltimer_ns());

    // ---------------------------------------------------------------- setup: barriers, TMEM, zeroed tile ring, work plan
    if (tid == 0) {
        tma_prefetch_desc(&tm_fts);
        for (int i = 0; i < POOL_NBUF; ++i) {
            mbar_init(&a_full[i], POOL_PROD_WARPS);
            mbar_init(&a_empty[i], NPW);
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], 4);
            mbar_init(&p_full[i], 32);
            mbar_init(&m_full[i], 3);
        }
        for (int i = 0; i < POOL_NBUF; ++i) mbar_init(&w_full[i], 1);
        mbar_init(t_ready, 256);
        mbar_init(ep_done, 1);
        *s_carry_c = -1;
        fence_mbar_init();
    }
    if (warp == POOL_MMA_WARP) tmem_alloc(tmem_slot, POOL_TMEM_COLS);
    // the tile buffers start as zeros: rows past a partial tile's end are never fetched, only multiplied by weight 0
    for (int i = tid; i < POOL_NBUF * L::A_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    pdl_wait();      // everything above overlaps the previous kernel's tail
    // the plan (gridmm_pool_plan): this CTA's record and the prefix of the valid-row counts -- one round trip to L2
    const int4 pl0 = __ldg(reinterpret_cast<const int4*>(p.plan + blockIdx.x * 8));
    const int4 pl1 = __ldg(reinterpret_cast<const int4*>(p.plan + blockIdx.x * 8) + 1);
    {
        const int* vb = p.plan + gridDim.x * 8;
        for (int i = tid; i <= p.batch; i += blockDim.x) s_vbase[i] = __ldg(vb + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    const int g0 = pl0.x, g1 = pl0.y;

    Walker wk;
    wk.init(s_vbase, p.batch, g0, g1);
    Tile t;
    const long long t_begin = p.dbg ? clock64() : 0;
    if (p.dbg && tid == 0)      // setup cycles (20 bits) | g1 << 20 | g0 << 40
        p.dbg[blockIdx.x * 16 + 13] = ((t_begin - t_entry) & 0xfffff) | (static_cast<long long>(g1) << 20) | (static_cast<long long>(g0) << 40);
    long long w_a = 0, w_b = 0;

    if (warp < POOL_PROD_WARPS) {
        // ------------------------------------------------------------ TMA gather producers (+ first half of the text operand)
        // Warp w owns rows 8w .. 8w+7 of every tile (two gather4 row groups): lanes 0..7 resolve one slab row each (perm -> slot ->
        // row, two dependent global loads, done for the NEXT tile while this tile's buffer is still busy), then one gather4 copy
        // (4 rows x 128 B) per row group and 64-column chunk is issued.
        const int q = warp & 3;
        const int tlane = q * 32 + lane;
        const uint32_t taddr_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        // Row resolution runs two tiles ahead of the copies: the perm entry of tile i + 2 is requested while tile i is issued, and
        // the slot table of the episode (slots[b][0 .. t_cap)) sits in registers (lane l holds steps l, l + 32, l + 64, l + 96), so
        // that no global round trip is left between a buffer falling free and the copies of the next tile.
        int tab[4] = {0, 0, 0, 0}, tab_b = -1;
        auto load_perm = [&](const Tile& tt) -> int {
            const int r = min(8 * warp + (lane & 7), tt.nrows - 1);            // rows past the tile's end duplicate its last row
            return __ldg(p.perm + static_cast<size_t>(tt.b) * p.cap + tt.pos + r);
        };
        auto finish_row = [&](const Tile& tt, int j) -> int {
            if (tt.b != tab_b) {
                tab_b = tt.b;
#pragma unroll
                for (int i = 0; i < 4; ++i) tab[i] = (lane + 32 * i < p.t_cap) ? __ldg(p.slots + tt.b * p.t_cap + lane + 32 * i) : 0;
            }
            const int step = j / POOL_PTS, qq = j - step * POOL_PTS;
            const int v = qq / POOL_VIEW_PTS, k = qq - v * POOL_VIEW_PTS;
            int slot;
            if (p.t_cap <= 128) {
                const int s0 = __shfl_sync(0xffffffffu, tab[0], step & 31), s1 = __shfl_sync(0xffffffffu, tab[1], step & 31);
                const int s2 = __shfl_sync(0xffffffffu, tab[2], step & 31), s3 = __shfl_sync(0xffffffffu, tab[3], step & 31);
                slot = (step < 64) ? (step < 32 ? s0 : s1) : (step < 96 ? s2 : s3);
            } else {
                slot = __ldg(p.slots + tt.b * p.t_cap + step);
            }
            return slot * p.slot_rows + v * p.view_rows + p.tok_off + k;
        };
        Tile tn;
        bool have = wk.next(t), have_n = false;
        int myrow = 0, j_n = 0;
        int it = 0, cur_b = -1, visits = 0;
        long long c_text = 0;
        if (have) {
            const int j0 = load_perm(t);
            have_n = wk.next(tn);
            if (have_n) j_n = load_perm(tn);
            // the range's first episode: this warp's half of the text operand goes to tensor memory BEFORE the first copies are
            // issued (text loads queued behind 148 SMs' first tiles took three times as long)
            const long long c1 = p.dbg ? clock64() : 0;
            cur_b = t.b; ++visits;
            if (q * 32 < p.l_pad) stage_text<D>(p.text_ws + static_cast<size_t>(t.b) * UNITS * 128, tlane, p.l_pad, taddr_lane, 0);
            tc_fence_before();
            mbar_arrive(t_ready);
            if (p.dbg) c_text += clock64() - c1;
            myrow = finish_row(t, j0);
        }
        while (have) {
            const int buf = it % POOL_NBUF;
            const uint32_t ph = (it / POOL_NBUF) & 1;
            const long long c0 = p.dbg ? clock64() : 0;
            mbar_wait_relaxed(&a_empty[buf], ph ^ 1);
            if (p.dbg) w_a += clock64() - c0;
            if (warp == 0) POOL_TRACE(0);
            // this warp's row group is present if it holds at least one row of the tile.  Everything the copies need is made
            // warp-uniform (shuffle broadcasts) and ONE elected lane issues the CH copies back to back (elect_one(), common.cuh).
            const int ngroups = min(2, max(0, ((uniform_i32(t.nrows) + 3) >> 2) - 2 * warp));
            int rr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) rr[i] = __shfl_sync(0xffffffffu, p.exp == 2 ? 8 * warp + (lane & 7) : myrow, i);
            const int nch = p.exp == 1 ? CH / 2 : CH;
            const int ubuf = uniform_i32(buf);
            const uint32_t dst = smem_u32(sA) + ubuf * L::A_BYTES + warp * 1024;
            if (elect_one()) {
                mbar_arrive_expect_tx(&a_full[ubuf], static_cast<uint32_t>(ngroups) * 4u * (p.exp == 1 ? D : D * 2u));
                if (ngroups > 0) {
#pragma unroll
                    for (int ck = 0; ck < CH; ++ck)
                        if (ck < nch) tma_gather4(dst + ck * L::A_CHUNK, &tm_fts, ck * 64, rr[0], rr[1], rr[2], rr[3], &a_full[ubuf]);
                }
                if (ngroups > 1) {
#pragma unroll
                    for (int ck = 0; ck < CH; ++ck)
                        if (ck < nch) tma_gather4(dst + 512 + ck * L::A_CHUNK, &tm_fts, ck * 64, rr[4], rr[5], rr[6], rr[7], &a_full[ubuf]);
                }
            }
            __syncwarp();
            if (warp == 0) POOL_TRACE(1);
            if (t.b != cur_b) {
                // first tile of a later episode: once the tensor core is done with the previous episode, move this warp's half of
                // the new text operand into TMEM (the tile just issued lands meanwhile)
                const long long c1 = p.dbg ? clock64() : 0;
                mbar_wait_guard(ep_done, (visits - 1) & 1);
                tc_fence_after();
                cur_b = t.b; ++visits;
                // (a quadrant without any real text position is never read back: its lanes keep whatever tensor memory held)
                if (q * 32 < p.l_pad) stage_text<D>(p.text_ws + static_cast<size_t>(t.b) * UNITS * 128, tlane, p.l_pad, taddr_lane, 0);
                tc_fence_before();
                mbar_arrive(t_ready);
                if (p.dbg) c_text += clock64() - c1;
            }
            // advance: the next tile's row (its perm entry was requested one tile ago), and the request for the tile after it
            t = tn; have = have_n;
            if (have) {
                myrow = finish_row(t, j_n);
                have_n = wk.next(tn);
                if (have_n) j_n = load_perm(tn);
            }
            ++it;
        }
        if (p.dbg && tid == 0) { p.dbg[blockIdx.x * 16 + 0] = clock64() - t_begin; p.dbg[blockIdx.x * 16 + 1] = w_a; p.dbg[blockIdx.x * 16 + 2] = c_text; }
    } else if (warp == POOL_MMA_WARP) {
        // ------------------------------------------------------------ MMA issuer: S^T = text (TMEM) x tile^T (smem)
        {   // the whole (converged) warp walks the tiles and waits; one elected lane issues (elect_one(), common.cuh)
            constexpr uint32_t idesc = umma_idesc_f16(128, POOL_ROWS);
            int it = 0, cur_b = -1, visits = 0;
            while (wk.next(t)) {
                const int buf = uniform_i32(it % POOL_NBUF);
                const uint32_t ph = (it / POOL_NBUF) & 1;
                if (t.b != cur_b) {
                    if (visits > 0 && elect_one()) umma_commit(ep_done);      // arrives when every MMA issued so far has retired
                    __syncwarp();
                    mbar_wait_guard(t_ready, visits & 1);
                    cur_b = t.b;
                    ++visits;
                }
                const int dbuf = buf;
                const uint32_t dph = ph;
                const long long c0 = p.dbg ? clock64() : 0;
                mbar_wait_guard(&d_empty[dbuf], dph ^ 1);
                const long long c1 = p.dbg ? clock64() : 0;
                mbar_wait_guard(&a_full[buf], ph);
                if (p.dbg) { w_b += c1 - c0; w_a += clock64() - c1; }
                POOL_TRACE(2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + D_COL0 + dbuf * POOL_ROWS;
                const uint32_t tile_s = smem_u32(sA) + buf * L::A_BYTES;
                if (elect_one()) {
                    // one descriptor per tile, advanced by plain adds (the start-address field counts 16-byte units): a fully
                    // unrolled loop moved 48 operand addresses from vector to uniform registers, ~20 cycles per instruction
                    uint64_t db = umma_desc_sw128_kmajor(tile_s);
                    uint32_t ta = tmem_base;
                    const int kend = p.exp == 3 ? CH / 2 : (p.exp == 4 ? 1 : CH);      // (timing experiments: half / a twelfth of the contraction)
#pragma unroll 1
                    for (int k = 0; k < kend; ++k) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)   // K = 16 per instruction = 8 TMEM columns of A, 32 bytes of B
                            umma_f16_ts(d_tmem, ta + kk * 8, db + 2 * kk, idesc, (k | kk) ? 1u : 0u);
                        db += L::A_CHUNK >> 4;
                        ta += 32;
                    }
                    umma_commit(&d_full[dbuf]);
                }
                __syncwarp();
                POOL_TRACE(3);
                ++it;
            }
            if (p.dbg && lane == 0) {
                long long* d = p.dbg + blockIdx.x * 16 + 3;
                d[0] = clock64() - t_begin; d[1] = w_a; d[2] = w_b;
            }
        }
    } else if (warp < POOL_MMA_WARP) {
        // ------------------------------------------------------------ relevance max + softmax numerators (+ second half of the text)
        const int q = warp & 3;                     // TMEM lane quadrant this warp may access
        const int tlane = q * 32 + lane;            // TMEM lane = text position
        const uint32_t taddr_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int e = tid - POOL_RED_WARP0 * 32;    // 0..127; threads 0..63 own one tile row each in the softmax part
        int it = 0, cur_b = -1;
        long long c_text = 0, c_soft = 0;
        while (wk.next(t)) {
            const int buf = it % POOL_NBUF;
            const uint32_t ph = (it / POOL_NBUF) & 1;
            if (t.b != cur_b) {
                // every MMA that read the previous episode's text has retired: this warp waited on d_full of the previous tile
                const long long c0 = p.dbg ? clock64() : 0;
                cur_b = t.b;
                named_bar_sync(4, 128);             // every reducer is done with the previous episode's table
                for (int i = e; i <= n_cells; i += 128) s_csr[i] = p.cell_start[t.b * (n_cells + 1) + i];
                for (int i = e; i < n_cells; i += 128) s_cr[i] = p.cell_rank[t.b * n_cells + i];
                if (q * 32 < p.l_pad) stage_text<D>(p.text_ws + static_cast<size_t>(t.b) * UNITS * 128, tlane, p.l_pad, taddr_lane, UNITS / 2);
                tc_fence_before();
                mbar_arrive(t_ready);
                named_bar_sync(4, 128);             // the new tables are visible to every reducer warp
                if (p.dbg) c_text += clock64() - c0;
            }
            // The softmax warp resolves the cell of its row (binary search over the episode's cell_start, ~300 cycles) BEFORE it
            // waits for the tile's relevance: this lookup only depends on the tile's position, and the wait -> numerators chain of
            // this single warp is what paces the whole kernel.
            const bool sm_valid = lane < t.nrows;
            int sm_cid = -1, sm_rank = -1, sm_last = 0;
            if (warp == POOL_RED_WARP0 + 3 && sm_valid && !p.max_only) {
                const int P = t.pos + lane;
                int a = 0, c = n_cells;             // last cell with cs[cell] <= P
                while (c - a > 1) {
                    const int mid = (a + c) >> 1;
                    if (s_csr[mid] <= P) a = mid; else c = mid;
                }
                sm_cid = a;
                sm_rank = s_cr[a];
                sm_last = (P + 1 == s_csr[a + 1]) ? 0x10000 : 0;
            }
            const int dbuf = buf;
            const uint32_t dph = ph;
            const long long c1 = p.dbg ? clock64() : 0;
            mbar_wait_guard(&d_full[dbuf], dph);
            const long long c2 = p.dbg ? clock64() : 0;
            if (warp == POOL_RED_WARP0 + 3) POOL_TRACE(4);
            tc_fence_after();
            float* part = s_part + buf * 4 * POOL_ROWS;
            float wmax = -INFINITY;                 // this warp's quadrant: max over its 32 text positions for tile row `lane`
            const bool has_text = q * 32 < p.l_pad; // a quadrant of padding lanes only (replicas of position 0) adds nothing to the max
            if (!has_text) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[dbuf]);
            } else {
                float v[32];
                const uint32_t ta = taddr_lane + D_COL0 + dbuf * POOL_ROWS;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t r0[16];
                    tmem_ld_32x32b_x16(ta + c * 16, r0);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[c * 16 + j] = __uint_as_float(r0[j]);
                }
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[dbuf]);     // the accumulator may be overwritten
                // max over the 32 lanes (= text positions) of every column (= tile row): one warp-wide redux.sync.max.f32 per column
                // (CREDUX, result in a uniform register; 32 independent instructions instead of a 5-level dependent shuffle
                // butterfly), lane c keeps column c
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float r = redux_max_f32(v[c]);
                    if (lane == c) wmax = r;
                }
            }
            if (warp != POOL_RED_WARP0 + 3) {
                // reducer warps 0..2 hand their partial maxima to warp 3 through shared memory + an mbarrier and move on to the next
                // tile; warp 3 owns TMEM lanes 96..127 (padding whenever the instruction has <= 96 tokens), so the softmax part
                // below runs in parallel with the others' butterflies
                part[(warp - POOL_RED_WARP0) * POOL_ROWS + lane] = wmax;
                __syncwarp();
                if (lane == 0) mbar_arrive(&m_full[buf]);
            } else {
                mbar_wait_guard(&m_full[buf], ph);  // (at an episode switch this also publishes the other warps' s_csr / s_cr loads)
                // ---- softmax numerator of row e (one warp: 32 rows): exp(w - max over its cell), the cell max taken with
                //      match.any + redux.max, merged with the carried max of a cell that an earlier tile opened (the pooling
                //      warps rescale their accumulators by s_scal)
                float w = fmaxf(fmaxf(wmax, part[lane]), fmaxf(part[POOL_ROWS + lane], part[2 * POOL_ROWS + lane]));
                const bool valid = lane < t.nrows;
                // text longer than 128 positions: the maxima over the positions of the earlier launch are merged in here
                if (valid && p.w_in) w = fmaxf(w, __ldg(p.w_in + static_cast<size_t>(t.b) * p.cap + t.pos + lane));
                if (p.max_only) {
                    if (valid) p.w_out[static_cast<size_t>(t.b) * p.cap + t.pos + lane] = w;
                    mbar_arrive(&p_full[buf]);
                    if (p.dbg) { w_a += c2 - c1; c_soft += clock64() - c2; }
                    ++it;
                    continue;
                }
                const int cid = sm_cid;
                if (valid && p.w_out) p.w_out[static_cast<size_t>(t.b) * p.cap + t.pos + lane] = w;
                const unsigned same = __match_any_sync(0xffffffffu, cid);
                float m = ord2f(__reduce_max_sync(same, f2ord(w)));
                const float carry_m = *s_carry_m;
                const int carry_c = *s_carry_c;     // (episode << 16 | cell) of the cell the previous tile left open
                const int key = (t.b << 16) | (cid & 0xffff);
                __syncwarp();                       // every lane has read the carry before the last row's lane replaces it
                const bool cont = valid && (key == carry_c);
                if (cont) m = fmaxf(m, carry_m);
                const float pnum = valid ? ex2_approx((w - m) * LOG2E) : 0.0f;
                s_p[buf * POOL_ROWS + lane] = pnum;
                // compact rank of the row's cell (= its row in `pooled`), flagged when this is the cell's last row
                const int rank = sm_rank;
                s_cid[buf * POOL_ROWS + lane] = valid ? (rank | sm_last) : -1;
                if (lane == 0) s_scal[buf] = cont ? ex2_approx((carry_m - m) * LOG2E) : 1.0f;
                if (lane == 0) s_mfirst[buf] = m;
                if (lane == t.nrows - 1) { *s_carry_m = m; *s_carry_c = key; s_mlast[buf] = m; }
                mbar_arrive(&p_full[buf]);          // release: s_p[buf] / s_scal[buf] are visible to the pooling warps
                POOL_TRACE(5);
            }
            if (p.dbg) { w_a += c2 - c1; c_soft += clock64() - c2; }
            ++it;
        }
        if (p.dbg && e == 96) {
            long long* d = p.dbg + blockIdx.x * 16 + 6;
            d[0] = clock64() - t_begin; d[1] = w_a; d[2] = c_text; d[3] = c_soft;
        }
    } else if (warp == POOL_POOL_WARP0 + NPW) {
        // ------------------------------------------------------------ builder of the pooling warps' weight operands
        // out[cell, :] = sum_r p[r] x[r, :] is a matrix product  X^T[D, 32 rows] . W[32 rows, 8 cell slots]  with W[r, slot] = p[r]
        // when row r belongs to the cell in that slot (slot = compact cell rank & 7; ranks are consecutive along the sorted
        // rows, so an open cell keeps its slot from tile to tile).  This warp turns the softmax numerators of a tile into the
        // mma.sync B fragments of every lane (fp16 value + fp16 rounding residual: with one fp16 weight the action logits drifted
        // to 1.08e-3), keeps the running weight sum of the (up to 8) open cells and publishes, per pass of 8 consecutive cells,
        // which slots complete in this tile and their normalisers -- once, instead of once per pooling warp.
        const int g = lane >> 2, tq = lane & 3;
        float s_slot = 0.f;                            // sum of weights of the cell in slot g (replicated over tq)
        bool head_pending = !p.max_only && (pl1.w & 1);     // the first cell of the range started in an earlier CTA: it is a piece
        int it = 0;
        long long c_build = 0;
        while (wk.next(t)) {
            const int buf = it % POOL_NBUF;
            const uint32_t ph = (it / POOL_NBUF) & 1;
            mbar_wait_guard(&p_full[buf], ph);
            const long long c1 = p.dbg ? clock64() : 0;
            if (p.max_only) {                          // first pass of a long text: the pooling warps only drain the ring
                __syncwarp();
                if (lane == 0) mbar_arrive(&w_full[buf]);
                ++it;
                continue;
            }
            const float* pp = s_p + buf * POOL_ROWS;
            const int* rk = s_cid + buf * POOL_ROWS;    // compact cell rank | (last row of its cell ? 0x10000 : 0); -1 = no row
            const int my_rk = rk[lane];
            const int rank0 = __shfl_sync(0xffffffffu, my_rk, 0) & 0xffff;
            const int rank_last = __shfl_sync(0xffffffffu, my_rk, t.nrows - 1) & 0xffff;
            {
                const float sc = s_scal[buf];          // a later tile raised the open cell's max (1 otherwise): its slot is rank0 & 7
                if (sc != 1.0f && g == (rank0 & 7)) s_slot *= sc;
            }
            // rows this lane feeds into the B fragments: row step ks covers rows 16 ks + {2tq, 2tq+1, 2tq+8, 2tq+9}
            float2 pv[2][2];
            int2 rv[2][2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    pv[ks][h] = *reinterpret_cast<const float2*>(pp + 16 * ks + 8 * h + 2 * tq);
                    rv[ks][h] = *reinterpret_cast<const int2*>(rk + 16 * ks + 8 * h + 2 * tq);
                }
            int ps = 0;
            for (int base = rank0; base <= rank_last; base += 8, ++ps) {       // one pass per 8 consecutive cells (normally one)
                uint32_t bh[2][2], bl[2][2];           // fp16 weights of slot g and their fp16 rounding residuals
                float add = 0.f;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int ra = (rv[ks][h].x & 0xffff) - base, rb = (rv[ks][h].y & 0xffff) - base;
                        const bool va = rv[ks][h].x >= 0 && ra >= 0 && ra < 8 && (((ra + base) & 7) == g);
                        const bool vb = rv[ks][h].y >= 0 && rb >= 0 && rb < 8 && (((rb + base) & 7) == g);
                        const float2 w = make_float2(va ? pv[ks][h].x : 0.f, vb ? pv[ks][h].y : 0.f);
                        const __half2 hi = __floats2half2_rn(w.x, w.y);
                        const float2 fhi = __half22float2(hi);
                        const __half2 lo = __floats2half2_rn(w.x - fhi.x, w.y - fhi.y);
                        bh[ks][h] = *reinterpret_cast<const uint32_t*>(&hi);
                        bl[ks][h] = *reinterpret_cast<const uint32_t*>(&lo);
                        // the normaliser is the sum of the weights the MMAs actually apply: with single fp16 weights the
                        // result is an exactly normalised convex combination with weights perturbed by <= 2^-12 relative
                        add += p.split_weights ? (w.x + w.y) : (fhi.x + fhi.y);
                    }
                uint4* fr = reinterpret_cast<uint4*>(sW) + ((buf * POOL_MAXPASS + ps) * 32 + lane) * 2;
                fr[0] = make_uint4(bh[0][0], bh[0][1], bh[1][0], bh[1][1]);
                fr[1] = make_uint4(bl[0][0], bl[0][1], bl[1][0], bl[1][1]);
                add += __shfl_xor_sync(0xffffffffu, add, 1);
                add += __shfl_xor_sync(0xffffffffu, add, 2);
                s_slot += add;
                // cells of this pass whose last row lies in this tile
                unsigned done = 0;
                {
                    const int rel = (my_rk & 0xffff) - base;
                    if (my_rk >= 0 && (my_rk & 0x10000) && rel >= 0 && rel < 8) done = 1u << ((rel + base) & 7);
                    done = __reduce_or_sync(0xffffffffu, done);
                }
                int piece = 0;
                if (done && head_pending) {
                    // cells complete in rank order: the first completion of the range is its first cell, i.e. the piece
                    int best = 8;
#pragma unroll
                    for (int s_ = 0; s_ < 8; ++s_)
                        if ((done >> s_) & 1u) best = min(best, (s_ - base) & 7);
                    const int hs = (base + best) & 7;
                    const float hsum = __shfl_sync(0xffffffffu, s_slot, hs * 4);
                    if (lane == 0) s_hsum[buf] = hsum;
                    if (g == hs) s_slot = 0.f;
                    done &= ~(1u << hs);
                    piece = hs + 1;
                    head_pending = false;
                }
                if ((done >> g) & 1u) {
                    if (tq == 0) s_fin[(buf * POOL_MAXPASS + ps) * 8 + g] = 1.0f / s_slot;
                    s_slot = 0.f;
                }
                if (lane == 0) s_pass[buf * POOL_MAXPASS + ps] = static_cast<int>(done) | (piece << 8);
            }
            {
                const float ts = __shfl_sync(0xffffffffu, s_slot, (rank_last & 7) * 4);     // weight sum of the cell left open (if any)
                if (lane == 0) { s_tsum[buf] = ts; s_meta[buf * 2] = rank0; s_meta[buf * 2 + 1] = ps; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&w_full[buf]);
            POOL_TRACE(6);
            if (p.dbg) c_build += clock64() - c1;
            ++it;
        }
        (void)c_build;
    } else {
        // ------------------------------------------------------------ weighted sums from the resident tile (warp-level HMMA)
        // Each warp owns 128 columns: per pass 8 column tiles x 2 row steps of mma.sync.m16n8k16 (A = the resident tile through
        // ldmatrix.trans, B = the builder warp's fragments, fp32 accumulators in registers across tiles, rescaled when a later
        // tile raises the open cell's max).  Cells cut between CTAs (PoolPlanCta) are exchanged per warp, i.e. per 128-column slice.
        const int pw = warp - POOL_POOL_WARP0;         // columns [128 pw, 128 pw + 128)
        const int g = lane >> 2, tq = lane & 3;
        float cacc[8][4];                              // [column tile][(col g, slot 2tq), (g, 2tq+1), (g+8, 2tq), (g+8, 2tq+1)]
#pragma unroll
        for (int i = 0; i < 8; ++i) cacc[i][0] = cacc[i][1] = cacc[i][2] = cacc[i][3] = 0.f;
        const int pflags = p.max_only ? 0 : pl1.w;
        // slice `pw` of the partial of CTA c (which = 0: its first cell, 1: its last cell)
        auto part_of = [&](int c, int which) -> float* {
            return reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.plan) + pool_ws_part_offset(p.batch, gridDim.x)) +
                   (static_cast<size_t>(c * 2 + which) * NPW + pw) * POOL_SLICE;
        };
        // this warp's 128 columns of cell slot s_ -> workspace, un-normalised, with the piece's weight sum and max; resets the slot
        auto flush_raw = [&](int s_, float* part, float wsum, float mval) {
            if (tq == (s_ >> 1)) {
                float* dst = part + g;
                if (s_ & 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { dst[i * 16] = cacc[i][1]; dst[i * 16 + 8] = cacc[i][3]; cacc[i][1] = cacc[i][3] = 0.f; }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { dst[i * 16] = cacc[i][0]; dst[i * 16 + 8] = cacc[i][2]; cacc[i][0] = cacc[i][2] = 0.f; }
                }
            }
            if (lane == 0) { part[128] = wsum; part[129] = mval; }
        };
        // arrival of this slice's piece on the chain [first, last] of n pieces; the last one to arrive merges the slices of all
        // pieces in CTA order (the first CTA of a chain holds the cell's start in its tail partial, every other piece is a head
        // partial):  out = sum_i e^(m_i - M) acc_i / sum_i e^(m_i - M) s_i.  atomicInc wraps the counter back to 0 for the next launch.
        auto arrive_merge = [&](int first, int last, int n, int b, int rank) {
            __threadfence();                           // this lane's part of the slice is visible device-wide before the arrival
            __syncwarp();
            int old = 0;
            if (lane == 0) old = static_cast<int>(atomicInc(reinterpret_cast<unsigned*>(p.plan + gridDim.x * 8 + p.batch + 1) + first * 8 + pw,
                                                            static_cast<unsigned>(n - 1)));
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old != n - 1) return;
            __threadfence();
            float M = -INFINITY;
            for (int c = first; c <= last; ++c) {
                if (n != last - first + 1) {
                    const int2 gg = __ldg(reinterpret_cast<const int2*>(p.plan + c * 8));
                    if (gg.y <= gg.x) continue;        // a CTA without rows holds no piece
                }
                M = fmaxf(M, __ldcg(part_of(c, c == first ? 1 : 0) + 129));
            }
            float S = 0.f;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = first; c <= last; ++c) {
                if (n != last - first + 1) {
                    const int2 gg = __ldg(reinterpret_cast<const int2*>(p.plan + c * 8));
                    if (gg.y <= gg.x) continue;
                }
                const float* part = part_of(c, c == first ? 1 : 0);
                const float f = ex2_approx((__ldcg(part + 129) - M) * LOG2E);
                S += f * __ldcg(part + 128);
                const float4 v = __ldcg(reinterpret_cast<const float4*>(part) + lane);
                acc.x += f * v.x; acc.y += f * v.y; acc.z += f * v.z; acc.w += f * v.w;
            }
            const float fin = 1.0f / S;
            __half2 h0 = __floats2half2_rn(acc.x * fin, acc.y * fin), h1 = __floats2half2_rn(acc.z * fin, acc.w * fin);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(p.pooled + (static_cast<size_t>(b) * n_cells + rank) * D + pw * 128 + lane * 4) = o;
        };
        int it = 0, pre_cnt = -1;
        long long c_loop = 0;
        while (wk.next(t)) {
            const int buf = it % POOL_NBUF;
            const uint32_t ph = (it / POOL_NBUF) & 1;
            const long long c0 = p.dbg ? clock64() : 0;
            if ((pflags & 2) && wk.g >= wk.g_end && lane == 0) {
                // last tile of a range that ends inside a cell: how many of the cell's later pieces have arrived (acquire: their
                // partials are visible to the loads that follow the loop)
                const unsigned* cp = reinterpret_cast<const unsigned*>(p.plan + gridDim.x * 8 + p.batch + 1) + blockIdx.x * 8 + pw;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(pre_cnt) : "l"(cp) : "memory");
            }
            mbar_wait_guard(&w_full[buf], ph);
            mbar_wait_guard(&a_full[buf], ph);         // already complete; makes the TMA-written tile visible to this thread
            const long long c1 = p.dbg ? clock64() : 0;
            if (p.max_only) {                          // first pass of a long text: only drain the ring
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_empty[buf]);
                ++it;
                continue;
            }
            const uint32_t tile_s = smem_u32(sA) + buf * L::A_BYTES;
            const int2 meta = *reinterpret_cast<const int2*>(s_meta + buf * 2);      // first compact cell rank, passes
            {
                const float sc = s_scal[buf];          // a later tile raised the open cell's max (1 otherwise): its slot is rank0 & 7
                if (sc != 1.0f) {
                    const int s0 = meta.x & 7;
                    if (2 * tq == s0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { cacc[i][0] *= sc; cacc[i][2] *= sc; }
                    } else if (2 * tq + 1 == s0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { cacc[i][1] *= sc; cacc[i][3] *= sc; }
                    }
                }
            }
            for (int ps = 0; ps < meta.y; ++ps) {
                const int base = meta.x + 8 * ps;
                const uint4* fr = reinterpret_cast<const uint4*>(sW) + ((buf * POOL_MAXPASS + ps) * 32 + lane) * 2;
                const uint4 fh = fr[0], fl = fr[1];
                {
                    const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const int r = 16 * ks + (mi >> 1) * 8 + rr;
                        const uint32_t bh0 = ks ? fh.z : fh.x, bh1 = ks ? fh.w : fh.y, bl0 = ks ? fl.z : fl.x, bl1 = ks ? fl.w : fl.y;
#pragma unroll
                        for (int ct = 0; ct < 8; ++ct) {       // 16 columns per tile: units 2ct, 2ct+1 of the warp's 16
                            const int chunk = 2 * pw + (ct >> 2), u = (ct & 3) * 2 + (mi & 1);
                            uint32_t a0, a1, a2, a3;
                            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                                         : "r"(tile_s + chunk * L::A_CHUNK + r * 128 + ((u ^ rr) << 4)));
                            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                         : "+f"(cacc[ct][0]), "+f"(cacc[ct][1]), "+f"(cacc[ct][2]), "+f"(cacc[ct][3])
                                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bh0), "r"(bh1));
                            if (p.split_weights)
                                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                             : "+f"(cacc[ct][0]), "+f"(cacc[ct][1]), "+f"(cacc[ct][2]), "+f"(cacc[ct][3])
                                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bl0), "r"(bl1));
                        }
                    }
                }
                const int dm = s_pass[buf * POOL_MAXPASS + ps];
                if (dm) {
                    if (dm >> 8) {
                        // the range's first cell began in an earlier CTA: its piece ends here
                        const int hs = (dm >> 8) - 1;
                        const int4 pa = __ldg(reinterpret_cast<const int4*>(p.plan + blockIdx.x * 8));
                        const int4 pb = __ldg(reinterpret_cast<const int4*>(p.plan + blockIdx.x * 8) + 1);
                        flush_raw(hs, part_of(blockIdx.x, 0), s_hsum[buf], s_mfirst[buf]);
                        arrive_merge(pa.z, pa.w, pb.x, t.b, base + ((hs - base) & 7));
                    }
                    const unsigned done = dm & 0xff;
                    if (done) {
                        // normalise, store (rows of `pooled` are the compact ranks), reset
                        const float2 fin = *reinterpret_cast<const float2*>(s_fin + (buf * POOL_MAXPASS + ps) * 8 + 2 * tq);
                        __half* out_b = p.pooled + static_cast<size_t>(t.b) * n_cells * D + pw * 128 + g;
                        if ((done >> (2 * tq)) & 1u) {
                            __half* orow = out_b + static_cast<size_t>(base + ((2 * tq - base) & 7)) * D;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                orow[i * 16] = __float2half_rn(cacc[i][0] * fin.x);
                                orow[i * 16 + 8] = __float2half_rn(cacc[i][2] * fin.x);
                                cacc[i][0] = cacc[i][2] = 0.f;
                            }
                        }
                        if ((done >> (2 * tq + 1)) & 1u) {
                            __half* orow = out_b + static_cast<size_t>(base + ((2 * tq + 1 - base) & 7)) * D;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                orow[i * 16] = __float2half_rn(cacc[i][1] * fin.y);
                                orow[i * 16 + 8] = __float2half_rn(cacc[i][3] * fin.y);
                                cacc[i][1] = cacc[i][3] = 0.f;
                            }
                        }
                    }
                }
            }
            if (p.dbg) { w_a += c1 - c0; c_loop += clock64() - c1; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_empty[buf]);
            if (pw == 0) POOL_TRACE(7);
            ++it;
        }
        if (it > 0 && (pflags & 6)) {
            // the cell left open by the last tile (t still holds it) continues in the next CTA(s), or the whole range lies inside one cell
            const int lbuf = (it - 1) % POOL_NBUF;
            const int rank = s_cid[lbuf * POOL_ROWS + t.nrows - 1] & 0xffff;
            const float wsum = s_tsum[lbuf], mval = s_mlast[lbuf];
            const int4 pa = __ldg(reinterpret_cast<const int4*>(p.plan + blockIdx.x * 8));
            const int4 pb = __ldg(reinterpret_cast<const int4*>(p.plan + blockIdx.x * 8) + 1);
            if (pflags & 4) {                          // one more piece of the chain its first cell belongs to
                flush_raw(rank & 7, part_of(blockIdx.x, 0), wsum, mval);
                arrive_merge(pa.z, pa.w, pb.x, t.b, rank);
            } else {
                // this CTA heads the chain of its last cell.  The other pieces are the FIRST cells of the following CTAs' ranges and
                // have normally arrived long ago (the counter was read when the last tile began): merge them into the registers and
                // store the finished row -- no round trip of this CTA's own piece through the workspace
                const int last = pb.y, n = pb.z, s_ = rank & 7;
                const bool fast = __shfl_sync(0xffffffffu, pre_cnt, 0) == n - 1 && n == last - static_cast<int>(blockIdx.x) + 1 && n <= 3;
                if (fast) {
                    const bool act = tq == (s_ >> 1);
                    float pm[2], psum[2], pv[2][16];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        if (k < n - 1) {
                            const float* part = part_of(blockIdx.x + 1 + k, 0);
                            pm[k] = __ldcg(part + 129); psum[k] = __ldcg(part + 128);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                pv[k][2 * i] = act ? __ldcg(part + i * 16 + g) : 0.f;
                                pv[k][2 * i + 1] = act ? __ldcg(part + i * 16 + g + 8) : 0.f;
                            }
                        }
                    }
                    float M = mval;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        if (k < n - 1) M = fmaxf(M, pm[k]);
                    const float f0 = ex2_approx((mval - M) * LOG2E);
                    float S = f0 * wsum;
                    float fk[2] = {0.f, 0.f};
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        if (k < n - 1) { fk[k] = ex2_approx((pm[k] - M) * LOG2E); S += fk[k] * psum[k]; }
                    const float fin = 1.0f / S;
                    if (act) {
                        __half* orow = p.pooled + (static_cast<size_t>(t.b) * n_cells + rank) * D + pw * 128 + g;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float v0 = f0 * ((s_ & 1) ? cacc[i][1] : cacc[i][0]), v1 = f0 * ((s_ & 1) ? cacc[i][3] : cacc[i][2]);
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                if (k < n - 1) { v0 += fk[k] * pv[k][2 * i]; v1 += fk[k] * pv[k][2 * i + 1]; }
                            orow[i * 16] = __float2half_rn(v0 * fin);
                            orow[i * 16 + 8] = __float2half_rn(v1 * fin);
                        }
                    }
                    // this slice's arrival: the counter wraps back to 0 for the next launch (nobody merges: the row is complete)
                    if (lane == 0) atomicInc(reinterpret_cast<unsigned*>(p.plan + gridDim.x * 8 + p.batch + 1) + blockIdx.x * 8 + pw,
                                             static_cast<unsigned>(n - 1));
                } else {
                    flush_raw(s_, part_of(blockIdx.x, 1), wsum, mval);
                    arrive_merge(blockIdx.x, last, n, t.b, rank);
                }
            }
        }
        if (p.dbg && tid == POOL_FIXED_THREADS) {
            long long* d = p.dbg + blockIdx.x * 16 + 10;
            d[0] = clock64() - t_begin; d[1] = w_a; d[2] = c_loop;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (p.dbg && tid == 0) p.dbg[blockIdx.x * 16 + 15] = static_cast<long long>(globaltimer_ns());
    if (warp == POOL_MMA_WARP) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, POOL_TMEM_COLS);
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

template <int D>
static int launch_pool(const void* fts, long long fts_rows, const void* text_fts, void* text_ws, int text_ws_ready, int grid,
                       PoolParams& p, float* w_scratch, cudaStream_t stream) {
    // gather4 tensor map: the slab as [fts_rows, D] fp16, box = 64 columns x 1 row (the instruction names 4 rows)
    CUtensorMap tm;
    const int rc = make_tmap_f16_2d(&tm, fts, static_cast<uint64_t>(D), static_cast<uint64_t>(fts_rows), static_cast<uint64_t>(D) * 2, 64, 1);
    if (rc) return rc;
    constexpr int smem = PoolSmem<D>::TOTAL;
    constexpr int threads = POOL_FIXED_THREADS + D / 4;
    GMM_CUDA_CHECK(cudaFuncSetAttribute(pool_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int l_total = p.l_pad;
    if (!text_ws_ready) {
        GMM_CUDA_CHECK(launch_pdl(text_to_lanes_kernel<D>, dim3(D / 8, p.batch, (l_total + 127) / 128), dim3(128), 0, stream,
                                  reinterpret_cast<const __half*>(text_fts), reinterpret_cast<uint4*>(text_ws), l_total));
        gridmm_count_launch(1);
    }
    if (l_total > 128) {
        // A text of 129..256 positions does not fit the 128 tensor-memory lanes: a first pass over positions 128.. (second block of
        // the workspace) only produces the row maxima; the main pass over positions 0..127 merges them before the softmax
        // (vilmodel.py:798 takes the max over ALL positions).  The features are streamed twice in this case.
        PoolParams p1 = p;
        p1.text_ws = p.text_ws + static_cast<size_t>(p.batch) * (D / 8) * 128;
        p1.l_pad = l_total - 128; p1.max_only = 1; p1.w_out = w_scratch; p1.w_in = nullptr;
        GMM_CUDA_CHECK(launch_pdl(pool_kernel<D>, dim3(grid), dim3(threads), smem, stream, tm, p1));
        gridmm_count_launch(1);
        p.l_pad = 128; p.w_in = w_scratch;
    }
    GMM_CUDA_CHECK(launch_pdl(pool_kernel<D>, dim3(grid), dim3(threads), smem, stream, tm, p));
    gridmm_count_launch(1);
    return 0;
}

}  // namespace gmm

static long long* g_pool_dbg = nullptr;
static int g_pool_split = 0;
// Debug hook: 0 = single fp16 softmax weights in the mma.sync weighted sums (half the HMMA count), 1 = value + residual.
extern "C" void gridmm_debug_set_pool_split(int on) { g_pool_split = on; }
static int g_pool_cost = gmm::POOL_EPISODE_COST, g_pool_snap = gmm::POOL_SNAP;
// Debug hook (tools/pool_probe.py): the plan's cost of an episode start in rows and the snap distance of a cut (< 0 keeps a value).
extern "C" void gridmm_debug_set_pool_plan(int episode_cost, int snap) {
    if (episode_cost >= 0) g_pool_cost = episode_cost;
    if (snap >= 0) g_pool_snap = snap;
}
static long long* g_pool_trace = nullptr;
// Debug hook (tools/pool_probe.py): [4][64][8] clock64 stamps of the stage hand-overs of the first 64 tiles of CTAs 0..3; null disables.
extern "C" void gridmm_debug_set_pool_trace(long long* t) { g_pool_trace = t; }
static int g_pool_exp = 0;
// Debug hook (tools/pool_probe.py): timing experiments of the producer stage (PoolParams::exp); results are garbage when != 0.
extern "C" void gridmm_debug_set_pool_exp(int e) { g_pool_exp = e; }
// Debug hook (tools/microbench2.py): per-CTA cycle counters [grid][16] written by the next pool launches; null disables.
extern "C" void gridmm_debug_set_pool_counters(long long* dbg) { g_pool_dbg = dbg; }

static int pool_grid(int num_ctas) {
    const int sms = gridmm_sm_count();
    if (sms <= 0) return 0;
    return num_ctas > 0 ? num_ctas : sms;
}

extern "C" long long gridmm_pool_ws_bytes(int batch, int feat_dim, int num_ctas) {
    const int grid = pool_grid(num_ctas);
    if (grid <= 0 || grid > gmm::POOL_MAX_CTAS || batch <= 0 || batch > gmm::POOL_MAX_BATCH) return -1;
    return static_cast<long long>(gmm::pool_ws_bytes(batch, feat_dim, grid));
}

extern "C" int gridmm_pool_plan(const int* cell_start, int n_cells, int batch, int feat_dim, int num_ctas, void* pool_ws,
                                cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (!cell_start || !pool_ws) return GRIDMM_ERR_ARG;
    if (batch > POOL_MAX_BATCH || n_cells > POOL_MAX_CELLS || (feat_dim != 768 && feat_dim != 512)) return GRIDMM_ERR_SHAPE;
    if (reinterpret_cast<uintptr_t>(pool_ws) & 15) return GRIDMM_ERR_SHAPE;
    const int grid = pool_grid(num_ctas);
    if (grid <= 0) return GRIDMM_ERR_DRIVER;
    if (grid > POOL_MAX_CTAS) return GRIDMM_ERR_SHAPE;
    GMM_CUDA_CHECK(launch_pdl(pool_plan_kernel, dim3(1), dim3(1024), 0, stream, cell_start, n_cells, batch, grid, g_pool_cost, g_pool_snap,
                              reinterpret_cast<int*>(pool_ws)));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_pool(const void* fts, long long fts_rows, int feat_dim, const int* slots, int t_cap, int slot_rows,
                           int view_rows, int tok_off, const int* perm, int cap, const int* cell_start, const int* cell_rank,
                           int n_cells, const void* text_fts, int l_pad, int batch, void* text_ws, int text_ws_ready,
                           void* pooled, float* w_out, float* w_scratch, void* pool_ws, int plan_ready, int num_ctas,
                           cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (!fts || !slots || !perm || !cell_start || !cell_rank || !text_ws || !pooled || !pool_ws) return GRIDMM_ERR_ARG;
    if (!text_ws_ready && !text_fts) return GRIDMM_ERR_ARG;
    if (batch > POOL_MAX_BATCH || n_cells > POOL_MAX_CELLS || l_pad < 1 || l_pad > 256 || fts_rows <= 0) return GRIDMM_ERR_SHAPE;
    if (l_pad > 128 && !w_scratch) return GRIDMM_ERR_ARG;
    if (feat_dim != 768 && feat_dim != 512) return GRIDMM_ERR_SHAPE;
    if ((reinterpret_cast<uintptr_t>(text_fts) & 15) || (reinterpret_cast<uintptr_t>(text_ws) & 15) ||
        (reinterpret_cast<uintptr_t>(pool_ws) & 15)) return GRIDMM_ERR_SHAPE;
    const int grid = pool_grid(num_ctas);
    if (grid <= 0) return GRIDMM_ERR_DRIVER;
    if (grid > POOL_MAX_CTAS) return GRIDMM_ERR_SHAPE;
    if (!plan_ready) {
        const int rc = gridmm_pool_plan(cell_start, n_cells, batch, feat_dim, num_ctas, pool_ws, stream);
        if (rc) return rc;
    }
    PoolParams p;
    p.slots = slots; p.perm = perm; p.cell_start = cell_start; p.cell_rank = cell_rank;
    p.text_ws = reinterpret_cast<const uint4*>(text_ws);
    p.pooled = reinterpret_cast<__half*>(pooled); p.w_out = w_out; p.w_in = nullptr; p.max_only = 0; p.split_weights = g_pool_split;
    p.plan = reinterpret_cast<int*>(pool_ws);
    p.batch = batch; p.t_cap = t_cap; p.cap = cap; p.n_cells = n_cells; p.l_pad = l_pad;
    p.slot_rows = slot_rows; p.view_rows = view_rows; p.tok_off = tok_off; p.dbg = g_pool_dbg; p.exp = g_pool_exp; p.trace = g_pool_trace;
    if (feat_dim == 768) return launch_pool<768>(fts, fts_rows, text_fts, text_ws, text_ws_ready, grid, p, w_scratch, stream);
    return launch_pool<512>(fts, fts_rows, text_fts, text_ws, text_ws_ready, grid, p, w_scratch, stream);
}
