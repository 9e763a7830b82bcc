// Instruction-relevance pooling: the reference's per-episode loop
//   grid_fts_weight = (grid_fts @ text_fts).max(-1)            map_nav_src/models/vilmodel.py:797-798
//   for i in range(196): softmax over the points of cell i, weighted sum   vilmodel.py:801-807
// as ONE persistent kernel that reads every valid patch-feature row from HBM exactly once.
//
//   * rows are streamed in cell-sorted order (gridmm_grid_update produced `perm`), 32 rows per tile, through a ring of
//     four 48 KB shared-memory tiles.  Four producer warps resolve the tile's slab rows (perm -> slot -> row) and fetch
//     them with TMA tile::gather4 (cp.async.bulk.tensor.2d...tile::gather4: four arbitrary rows x 64 fp16 columns per
//     instruction, hardware SWIZZLE_128B, completion on an mbarrier) straight into a K-major operand tile (D/64 chunks
//     of 32 x 128 B).  Row indices are made warp-uniform with shuffles and one elected lane issues the copies back to back
//     (elect_one(), common.cuh: per-lane operands cost a ~100-cycle R2UR waterfall per gather4);
//   * relevance  S^T[L, 32] = text_fts[L, D] . X_tile^T  on tcgen05 with the A operand (text_fts of the current
//     episode, up to 128 positions) held in TENSOR MEMORY for the whole episode (tcgen05.st by 8 warps, software
//     pipelined), B operand = the feature tile in shared memory, one TMEM accumulator per ring slot;
//     w = max over ALL text positions (padding included -- vilmodel.py:798) = max over TMEM lanes, taken with a
//     warp butterfly (31 shuffles per thread) by the reducer warps whose lane quadrant holds real text positions, merged
//     across quadrants through shared memory + an mbarrier by the reducer warp that owns the padding quadrant;
//   * that warp also turns w into per-cell softmax numerators exp(w - cell max) (binary search of the row's cell,
//     match.any + redux.max inside the warp, running max carried across tiles) and publishes each row's compact cell rank;
//   * weighted sums as warp-level HMMA from the resident tile: out[cell, :] = sum_r p[r] x[r, :] is the product
//     X^T[D, 32 rows] . W[32 rows, 8 cell slots] (slot = rank & 7; an open cell keeps its slot across tiles), A through
//     ldmatrix.trans, weights as fp16 value + fp16 residual (two MMAs), fp32 accumulators in registers across tiles,
//     rescaled when a later tile raises the open cell's max.  CTA ranges are cut at cell boundaries, so no atomics and no
//     cross-CTA merge exist and the result is deterministic;
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
constexpr int POOL_POOL_WARP0 = 9;     // warps 9.. : weighted sums.  TC mode: warp 9 issues the pooling MMAs, warps 10..13 are the
                                       // accumulator epilogue (one per tensor-memory lane quadrant); HMMA mode: D / 128 mma.sync warps
constexpr int POOL_FIXED_THREADS = POOL_POOL_WARP0 * 32;
constexpr int POOL_TC_THREADS = (POOL_POOL_WARP0 + 5) * 32;
constexpr int POOL_MAXPASS = 4;        // a 32-row tile holds at most 32 cells = 4 passes of 8 cell slots
constexpr int POOL_W_BYTES = 16 * 32 * 2;      // one pass of the weight operand: [16 = 8 slots x (hi, lo)] x [32 rows] fp16
constexpr int POOL_MAX_BATCH = 1024;
constexpr int POOL_MAX_CELLS = 256;
constexpr int POOL_TMEM_COLS = 512;
constexpr int POOL_PTS = 588, POOL_VIEW_PTS = 49;   // points per viewpoint / per view (r2r/env.py:279-289)
constexpr int POOL_EPISODE_COST = 96;               // rows' worth of time an episode switch costs a CTA (text staging, partial tiles)

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
    int batch, t_cap, cap, n_cells;
    int l_pad;               // text positions of THIS launch (<= 128: one per tensor-memory lane)
    int slot_rows, view_rows, tok_off;   // row = slot*slot_rows + view*view_rows + tok_off + patch
    long long* dbg;          // optional [grid][16] cycle counters (tools/microbench2.py), null in production
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

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One butterfly level of the lane-max: every lane keeps the half of its N columns selected by `bit` of its lane id and
// merges in the partner's copy of that half.  After the 32 -> 1 levels lane l holds the warp-wide max of column l.
template <int N>
__device__ __forceinline__ void lane_max_level(float (&v)[32], int lane, int bit) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
        const float keep = up ? v[N / 2 + j] : v[j];
        const float send = up ? v[j] : v[N / 2 + j];
        v[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, bit));
    }
}

// order-preserving float <-> uint32 map (for redux.sync max over the rows of one cell)
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// MN-major shared-memory operand of the pooling MMA: A[M = 128 feature dims, K = 16 tile rows] read straight out of the K-major
// feature tile the relevance MMA uses (chunk = 64 dims: 32 rows x 128 B, SWIZZLE_128B).  Seen MN-major, a swizzle atom is
// 64 dims (128 B) x 8 rows; atoms repeat every `chunk_bytes` along M (LBO) and every 1024 B along K (SBO).
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
    return d;
}
// K-major operand WITHOUT swizzle: 8 x 8 core matrices of 128 contiguous bytes (row r of a core matrix = 16 B at r * 16);
// lbo = byte distance between core matrices adjacent in K, sbo = between core matrices adjacent in M/N.
__device__ __forceinline__ uint64_t umma_desc_noswizzle_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}
// kind::f16 instruction descriptor with an MN-major A operand (bit 15), K-major B, fp32 accumulate
__host__ __device__ constexpr uint32_t umma_idesc_f16_amn(int m, int n) {
    return (1u << 4) | (1u << 15) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
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
                                      + POOL_NBUF * POOL_MAXPASS * 8 * 4;   // weight sum per cell slot (TC mode)
    static constexpr int W_BYTES = POOL_NBUF * POOL_MAXPASS * POOL_W_BYTES;      // weight operands of the pooling MMAs (TC mode)
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
// >= l_pad replicate position 0 so that they never change the max), 16-byte units [u0, u0 + 8 * BU), software pipelined
// two batches deep.
template <int D, int BU>
__device__ __forceinline__ void stage_text(const uint4* ws_b, int tlane, int l_pad, uint32_t taddr_lane, int u0) {
    const uint4* src = ws_b + (tlane < l_pad ? tlane : 0);
    uint4 cur[BU], nxt[BU];
#pragma unroll
    for (int c = 0; c < BU; ++c) cur[c] = __ldg(src + static_cast<size_t>(u0 + c) * 128);
#pragma unroll
    for (int bi = 0; bi < 8; ++bi) {
        if (bi + 1 < 8) {
#pragma unroll
            for (int c = 0; c < BU; ++c) nxt[c] = __ldg(src + static_cast<size_t>(u0 + (bi + 1) * BU + c) * 128);
        }
#pragma unroll
        for (int c = 0; c < BU / 2; ++c) {
            const uint32_t v8[8] = {cur[2 * c].x, cur[2 * c].y, cur[2 * c].z, cur[2 * c].w,
                                    cur[2 * c + 1].x, cur[2 * c + 1].y, cur[2 * c + 1].z, cur[2 * c + 1].w};
            // unit u holds fp16 elements 8u..8u+7 = TMEM columns 4u..4u+3
            tmem_st_32x32b_x8(taddr_lane + (u0 + bi * BU + 2 * c) * 4, v8);
        }
#pragma unroll
        for (int c = 0; c < BU; ++c) cur[c] = nxt[c];
    }
    tmem_st_wait();
}

template <int D, bool TC>
__global__ void __launch_bounds__(TC ? POOL_TC_THREADS : POOL_FIXED_THREADS + D / 4, 1)
pool_kernel(const __grid_constant__ CUtensorMap tm_fts, PoolParams p) {
    using L = PoolSmem<D>;
    constexpr int CH = L::CH;
    constexpr int NPW = D / 128;                  // pooling warps (HMMA mode) = 128-dim blocks of a feature row
    constexpr int A_COLS = D / 2;                 // TMEM columns of the text operand (two fp16 per column)
    constexpr int D_COL0 = A_COLS;                // relevance accumulators behind it: 32 columns per tile buffer (TC mode: ONE buffer --
                                                  // the pooling accumulators take the rest of tensor memory; the reducers read a
                                                  // relevance tile back within ~300 cycles, well inside the HBM time of a tile)
    constexpr int NDBUF = TC ? 1 : POOL_NBUF;     // relevance accumulators in flight
    constexpr int P_COL0 = D_COL0 + NDBUF * POOL_ROWS;   // TC mode: pooling accumulators, 16 columns (8 slots x (hi, lo)) per 128-dim block
    static_assert(!TC || P_COL0 + NPW * 16 <= POOL_TMEM_COLS, "text operand + accumulators must fit in tensor memory");
    constexpr int UNITS = D / 8;                  // 16-byte units per text position
    constexpr int BU = UNITS / 16;                // units per staging batch (16 batches: 8 per half)
    constexpr float LOG2E = 1.4426950408889634f;
    static_assert(A_COLS + POOL_NBUF * POOL_ROWS <= POOL_TMEM_COLS, "text operand + accumulators must fit in tensor memory");
    static_assert(BU % 2 == 0 && CH <= 32 && POOL_ROWS == 8 * POOL_PROD_WARPS && (D == 512 || D == 768), "unsupported feature width");
    extern __shared__ uint8_t smem_raw[];
    // align by pointer arithmetic on the __shared__ array (an integer round-trip would demote every access below to a
    // generic LD/ST instead of LDS/STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                            // [NBUF][CH][32 x 128 B]
    uint8_t* sW = sA + POOL_NBUF * L::A_BYTES;     // [NBUF][MAXPASS][2 n-cores][4 k-cores][8 x 16 B]: weights of the pooling MMAs (TC mode)
    uint8_t* misc = sW + L::W_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(misc);          // [NBUF] tile landed (producer arrivals + TMA transaction bytes)
    uint64_t* a_empty = a_full + POOL_NBUF;                        // [NBUF] tile buffer drained by the pooling warps
    uint64_t* d_full = a_empty + POOL_NBUF;                        // [NBUF] accumulator written (tcgen05.commit)
    uint64_t* d_empty = d_full + POOL_NBUF;                        // [NBUF] accumulator read back
    uint64_t* p_full = d_empty + POOL_NBUF;                        // [NBUF] softmax numerators of the tile are in s_p
    uint64_t* m_full = p_full + POOL_NBUF;                         // [NBUF] partial maxima of reducer warps 1..3 are in s_part
    uint64_t* t_ready = m_full + POOL_NBUF;                              // [1] text operand of the episode is in TMEM
    uint64_t* ep_done = t_ready + 1;                               // [1] every MMA of the previous episode has retired
    uint64_t* pacc_full = ep_done + 1;                             // [1] TC mode: the pooling MMAs of a pass have retired
    uint64_t* pacc_empty = pacc_full + 1;                          // [1] TC mode: the epilogue warps have read the pooling accumulators
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 60);
    float* s_p = reinterpret_cast<float*>(misc + 64 * 8);          // [NBUF][32] exp(w - cell max) per row
    int* s_cid = reinterpret_cast<int*>(s_p + POOL_NBUF * POOL_ROWS);      // [NBUF][32] compact cell rank (+ last-row flag) of every row
    float* s_part = reinterpret_cast<float*>(s_cid + POOL_NBUF * POOL_ROWS);   // [2][4][32]
    float* s_scal = s_part + POOL_NBUF * 4 * POOL_ROWS;            // [NBUF] rescale of the open cell per buffer
    float* s_carry_m = s_scal + POOL_NBUF;         // [1] running max of the cell left open by the previous tile
    int* s_carry_c = reinterpret_cast<int*>(s_carry_m + 1);        // [1] its (episode << 16 | cell) key (-1: none)
    int* s_range = s_carry_c + 1;                                  // [0] g_start, [1] g_end
    int* s_meta = s_range + 2;                                     // [NBUF][2] first / last compact cell rank of the tile (TC mode)
    int* s_csr = reinterpret_cast<int*>(misc + 64 * 8 + 2 * POOL_NBUF * POOL_ROWS * 4 + POOL_NBUF * 4 * POOL_ROWS * 4 + 128);   // reducers' cell_start
    int* s_cs = s_csr + POOL_MAX_CELLS + 1;                        // (spare table slot)
    int* s_cr = s_cs + POOL_MAX_CELLS + 1;                         // reducers' cell_rank [n_cells]
    int* s_vbase = s_cr + POOL_MAX_CELLS;                          // [batch + 1]
    float* s_wsum = reinterpret_cast<float*>(s_vbase + POOL_MAX_BATCH + 1);     // [NBUF][MAXPASS][8] weight sum per cell slot (TC mode)

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n_cells = p.n_cells;

    // ---------------------------------------------------------------- setup: barriers, TMEM, schedule
    if (tid == 0) {
        tma_prefetch_desc(&tm_fts);
        for (int i = 0; i < POOL_NBUF; ++i) {
            mbar_init(&a_full[i], POOL_PROD_WARPS);
            mbar_init(&a_empty[i], TC ? 1 : NPW);
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], 4);
            mbar_init(&p_full[i], 32);
            mbar_init(&m_full[i], 3);
        }
        mbar_init(t_ready, 256);
        mbar_init(ep_done, 1);
        mbar_init(pacc_full, 1);
        mbar_init(pacc_empty, 4);
        *s_carry_c = -1;
        fence_mbar_init();
    }
    if (warp == POOL_MMA_WARP) tmem_alloc(tmem_slot, POOL_TMEM_COLS);
    pdl_wait();      // barrier init / TMEM allocation above overlap the previous kernel's tail
    // the tile buffers start as zeros: rows past a partial tile's end are never fetched, only multiplied by weight 0
    for (int i = tid; i < POOL_NBUF * L::A_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
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
    if (warp < 2) {
        // CTA range [g0, g1) in global sorted-valid coordinates, snapped up to a cell boundary: warp w resolves boundary
        // blockIdx.x + w with ONE round trip to global memory (every lane loads a slice of the episode's cell_start row)
        // Work is split in COST units: one per valid row plus POOL_EPISODE_COST per episode start (moving a new text operand into
        // tensor memory and the partial tiles around an episode switch cost about three tiles), so a CTA whose range crosses an
        // episode boundary gets fewer rows: equal rows per CTA left the slowest CTA 20 % behind the mean.
        const int total = s_vbase[p.batch];
        const long long total_c = static_cast<long long>(total) + static_cast<long long>(p.batch) * POOL_EPISODE_COST;
        const long long tgt = (static_cast<long long>(blockIdx.x + warp) * total_c) / gridDim.x;
        int g;
        if (blockIdx.x + warp >= gridDim.x) g = total;
        else if (tgt <= 0) g = 0;
        else {
            int lo = 0, hi = p.batch;      // largest b with vbase[b] + b * COST <= tgt
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (static_cast<long long>(s_vbase[mid]) + static_cast<long long>(mid) * POOL_EPISODE_COST <= tgt) lo = mid; else hi = mid;
            }
            const int nv = s_vbase[lo + 1] - s_vbase[lo];
            const long long lc = tgt - s_vbase[lo] - static_cast<long long>(lo) * POOL_EPISODE_COST - POOL_EPISODE_COST;
            const int local = lc <= 0 ? 0 : (lc >= nv ? nv : static_cast<int>(lc));
            const int* cs = p.cell_start + lo * (n_cells + 1);
            int best = nv;                 // smallest cell boundary >= local (cell_start is non-decreasing, cs[n_cells] = nv)
            for (int i = lane; i <= n_cells; i += 32) {
                const int v = cs[i];
                if (v >= local) best = min(best, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            g = s_vbase[lo] + best;
        }
        if (lane == 0) s_range[warp] = g;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    const int g0 = s_range[0], g1 = s_range[1];

    Walker wk;
    wk.init(s_vbase, p.batch, g0, g1);
    Tile t;
    const long long t_begin = p.dbg ? clock64() : 0;
    long long w_a = 0, w_b = 0;

    if (warp < POOL_PROD_WARPS) {
        // ------------------------------------------------------------ TMA gather producers (+ first half of the text operand)
        // Warp w owns rows 8w .. 8w+7 of every tile (two gather4 row groups): lanes 0..7 resolve one slab row each (perm -> slot ->
        // row, two dependent global loads, done for the NEXT tile while this tile's buffer is still busy), then one gather4 copy
        // (4 rows x 128 B) per row group and 64-column chunk is issued.
        const int q = warp & 3;
        const int tlane = q * 32 + lane;
        const uint32_t taddr_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        int myrow = 0;
        auto resolve = [&](const Tile& tt) {
            const int r = min(8 * warp + (lane & 7), tt.nrows - 1);            // rows past the tile's end duplicate its last row
            const int j = p.perm[static_cast<size_t>(tt.b) * p.cap + tt.pos + r];
            const int step = j / POOL_PTS, qq = j - step * POOL_PTS;
            const int v = qq / POOL_VIEW_PTS, k = qq - v * POOL_VIEW_PTS;
            myrow = p.slots[tt.b * p.t_cap + step] * p.slot_rows + v * p.view_rows + p.tok_off + k;
        };
        bool have = wk.next(t);
        if (have) resolve(t);
        int it = 0, cur_b = -1, visits = 0;
        long long c_text = 0;
        while (have) {
            const int buf = it % POOL_NBUF;
            const uint32_t ph = (it / POOL_NBUF) & 1;
            const long long c0 = p.dbg ? clock64() : 0;
            mbar_wait_relaxed(&a_empty[buf], ph ^ 1);
            if (p.dbg) w_a += clock64() - c0;
            // this warp's row group is present if it holds at least one row of the tile.  Everything the copies need is made
            // warp-uniform (shuffle broadcasts) and ONE elected lane issues the CH copies back to back (elect_one(), common.cuh).
            const int ngroups = min(2, max(0, ((uniform_i32(t.nrows) + 3) >> 2) - 2 * warp));
            int rr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) rr[i] = __shfl_sync(0xffffffffu, myrow, i);
            const int ubuf = uniform_i32(buf);
            const uint32_t dst = smem_u32(sA) + ubuf * L::A_BYTES + warp * 1024;
            if (elect_one()) {
                mbar_arrive_expect_tx(&a_full[ubuf], static_cast<uint32_t>(ngroups) * 4u * D * 2u);
                if (ngroups > 0) {
#pragma unroll
                    for (int ck = 0; ck < CH; ++ck)
                        tma_gather4(dst + ck * L::A_CHUNK, &tm_fts, ck * 64, rr[0], rr[1], rr[2], rr[3], &a_full[ubuf]);
                }
                if (ngroups > 1) {
#pragma unroll
                    for (int ck = 0; ck < CH; ++ck)
                        tma_gather4(dst + 512 + ck * L::A_CHUNK, &tm_fts, ck * 64, rr[4], rr[5], rr[6], rr[7], &a_full[ubuf]);
                }
            }
            __syncwarp();
            if (t.b != cur_b) {
                // first tile of an episode: once the tensor core is done with the previous episode, move this warp's half of
                // the new text operand into TMEM (the tile just issued lands meanwhile)
                const long long c1 = p.dbg ? clock64() : 0;
                if (visits > 0) mbar_wait_guard(ep_done, (visits - 1) & 1);
                tc_fence_after();
                cur_b = t.b; ++visits;
                stage_text<D, BU>(p.text_ws + static_cast<size_t>(t.b) * UNITS * 128, tlane, p.l_pad, taddr_lane, 0);
                tc_fence_before();
                mbar_arrive(t_ready);
                if (p.dbg) c_text += clock64() - c1;
            }
            have = wk.next(t);
            if (have) resolve(t);
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
                const int dbuf = TC ? 0 : buf;                    // TC mode: one relevance accumulator, phases advance per tile
                const uint32_t dph = TC ? (it & 1) : ph;
                const long long c0 = p.dbg ? clock64() : 0;
                mbar_wait_guard(&d_empty[dbuf], dph ^ 1);
                const long long c1 = p.dbg ? clock64() : 0;
                mbar_wait_relaxed(&a_full[buf], ph);
                if (p.dbg) { w_b += c1 - c0; w_a += clock64() - c1; }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + D_COL0 + dbuf * POOL_ROWS;
                const uint32_t tile_s = smem_u32(sA) + buf * L::A_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < CH; ++k) {
                        const uint64_t db = umma_desc_sw128_kmajor(tile_s + k * L::A_CHUNK);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)   // K = 16 per instruction = 8 TMEM columns of A, 32 bytes of B
                            umma_f16_ts(d_tmem, tmem_base + k * 32 + kk * 8, db + 2 * kk, idesc, (k | kk) ? 1u : 0u);
                    }
                    umma_commit(&d_full[dbuf]);
                }
                __syncwarp();
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
                stage_text<D, BU>(p.text_ws + static_cast<size_t>(t.b) * UNITS * 128, tlane, p.l_pad, taddr_lane, UNITS / 2);
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
            const int dbuf = TC ? 0 : buf;
            const uint32_t dph = TC ? (it & 1) : ph;
            const long long c1 = p.dbg ? clock64() : 0;
            mbar_wait_guard(&d_full[dbuf], dph);
            const long long c2 = p.dbg ? clock64() : 0;
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
                if (lane == t.nrows - 1) { *s_carry_m = m; *s_carry_c = key; }
                if constexpr (TC) {
                    // B operand of the pooling MMAs: W^T[16 = 8 cell slots x (hi, lo)][32 rows] per pass of 8 consecutive cell ranks,
                    // row r of the tile has weight p_r in the slot of its cell (slot = rank & 7, so an open cell keeps its slot
                    // from tile to tile) as fp16 value (n-core 0) + fp16 rounding residual (n-core 1): K-major without swizzle,
                    // 8 x 8 core matrices [n-core][k-core][slot][row & 7]
                    const int rank0 = __shfl_sync(0xffffffffu, rank, 0);
                    const int rank_last = __shfl_sync(0xffffffffu, rank, t.nrows - 1);
                    const int npass = ((rank_last - rank0) >> 3) + 1;
                    uint8_t* wb = sW + buf * (POOL_MAXPASS * POOL_W_BYTES);
                    for (int i = lane; i < npass * (POOL_W_BYTES / 16); i += 32) reinterpret_cast<uint4*>(wb)[i] = make_uint4(0, 0, 0, 0);
                    __syncwarp();
                    if (valid) {
                        const __half hi = __float2half_rn(pnum);
                        const __half lo = __float2half_rn(pnum - __half2float(hi));
                        __half* dst = reinterpret_cast<__half*>(wb + ((rank - rank0) >> 3) * POOL_W_BYTES + (lane >> 3) * 128 + (rank & 7) * 16) + (lane & 7);
                        dst[0] = hi;
                        dst[256] = lo;                  // n-core 1 (512 B further)
                    }
                    if (lane == 0) { s_meta[buf * 2] = rank0; s_meta[buf * 2 + 1] = rank_last; }
                    // sum of the weights of every cell slot of every pass (the epilogue warps normalise with it): one integer
                    // redux over the lanes of a cell (24-bit fixed point: p <= 1, at most 32 rows -> exact to 2^-24)
                    float* ws = s_wsum + buf * (POOL_MAXPASS * 8);
                    ws[lane] = 0.0f;
                    __syncwarp();
                    const int qsum = __reduce_add_sync(same, __float2int_rn(pnum * 16777216.0f));
                    if (valid && lane == __ffs(same) - 1) ws[((rank - rank0) >> 3) * 8 + (rank & 7)] = static_cast<float>(qsum) * (1.0f / 16777216.0f);
                    fence_proxy_async_smem();           // generic-proxy stores -> visible to the tensor core's operand reads
                }
                mbar_arrive(&p_full[buf]);          // release: s_p[buf] / s_scal[buf] are visible to the pooling warps
            }
            if (p.dbg) { w_a += c2 - c1; c_soft += clock64() - c2; }
            ++it;
        }
        if (p.dbg && e == 96) {
            long long* d = p.dbg + blockIdx.x * 16 + 6;
            d[0] = clock64() - t_begin; d[1] = w_a; d[2] = c_text; d[3] = c_soft;
        }
    } else if constexpr (!TC) {
        // ------------------------------------------------------------ weighted sums from the resident tile (warp-level HMMA)
        // out[cell, :] = sum_r p[r] x[r, :] is a matrix product  X^T[D, 32 rows] . W[32 rows, 8 cell slots]  with W[r, slot] = p[r]
        // when row r belongs to the cell in that slot (slot = compact cell rank & 7; ranks are consecutive along the sorted
        // rows, so an open cell keeps its slot from tile to tile).  Each warp owns 128 columns: 8 column tiles x 2 row steps of
        // mma.sync.m16n8k16 per tile (A = the resident tile through ldmatrix.trans, B = weights built in registers from
        // s_p / s_cid, fp32 accumulators in registers across tiles) instead of ~570 scalar instructions.  Each weight enters
        // as fp16 value + fp16 rounding residual (two MMAs on the same A fragment), i.e. to ~2^-22: with one fp16 weight the
        // action logits drifted to 1.08e-3.
        const int pw = warp - POOL_POOL_WARP0;         // columns [128 pw, 128 pw + 128)
        const int g = lane >> 2, tq = lane & 3;
        float cacc[8][4];                              // [column tile][(col g, slot 2tq), (g, 2tq+1), (g+8, 2tq), (g+8, 2tq+1)]
#pragma unroll
        for (int i = 0; i < 8; ++i) cacc[i][0] = cacc[i][1] = cacc[i][2] = cacc[i][3] = 0.f;
        float s_slot = 0.f;                            // sum of weights of the cell in slot g (replicated over tq)
        int it = 0;
        long long c_loop = 0;
        while (wk.next(t)) {
            const int buf = it % POOL_NBUF;
            const uint32_t ph = (it / POOL_NBUF) & 1;
            const long long c0 = p.dbg ? clock64() : 0;
            mbar_wait_guard(&p_full[buf], ph);
            mbar_wait_guard(&a_full[buf], ph);         // already complete; makes the TMA-written tile visible to this thread
            const long long c1 = p.dbg ? clock64() : 0;
            if (p.max_only) {                          // first pass of a long text: only drain the ring
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_empty[buf]);
                ++it;
                continue;
            }
            const uint32_t tile_s = smem_u32(sA) + buf * L::A_BYTES;
            const float* pp = s_p + buf * POOL_ROWS;
            const int* rk = s_cid + buf * POOL_ROWS;    // compact cell rank | (last row of its cell ? 0x10000 : 0); -1 = no row
            const int my_rk = rk[lane];
            const int rank0 = __shfl_sync(0xffffffffu, my_rk, 0) & 0xffff;
            const int rank_last = __shfl_sync(0xffffffffu, my_rk, t.nrows - 1) & 0xffff;
            {
                const float sc = s_scal[buf];          // a later tile raised the open cell's max (1 otherwise): its slot is rank0 & 7
                if (sc != 1.0f) {
                    const int s0 = rank0 & 7;
                    if (g == s0) s_slot *= sc;
                    if (2 * tq == s0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { cacc[i][0] *= sc; cacc[i][2] *= sc; }
                    } else if (2 * tq + 1 == s0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { cacc[i][1] *= sc; cacc[i][3] *= sc; }
                    }
                }
            }
            // rows this thread feeds into the B fragments: row step ks covers rows 16 ks + {2tq, 2tq+1, 2tq+8, 2tq+9}
            float2 pv[2][2];
            int2 rv[2][2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    pv[ks][h] = *reinterpret_cast<const float2*>(pp + 16 * ks + 8 * h + 2 * tq);
                    rv[ks][h] = *reinterpret_cast<const int2*>(rk + 16 * ks + 8 * h + 2 * tq);
                }
            for (int base = rank0; base <= rank_last; base += 8) {       // one pass per 8 consecutive cells (normally one)
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
                add += __shfl_xor_sync(0xffffffffu, add, 1);
                add += __shfl_xor_sync(0xffffffffu, add, 2);
                s_slot += add;
                {
                    const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const int r = 16 * ks + (mi >> 1) * 8 + rr;
#pragma unroll
                        for (int ct = 0; ct < 8; ++ct) {       // 16 columns per tile: units 2ct, 2ct+1 of the warp's 16
                            const int chunk = 2 * pw + (ct >> 2), u = (ct & 3) * 2 + (mi & 1);
                            uint32_t a0, a1, a2, a3;
                            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                                         : "r"(tile_s + chunk * L::A_CHUNK + r * 128 + ((u ^ rr) << 4)));
                            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                         : "+f"(cacc[ct][0]), "+f"(cacc[ct][1]), "+f"(cacc[ct][2]), "+f"(cacc[ct][3])
                                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bh[ks][0]), "r"(bh[ks][1]));
                            if (p.split_weights)
                                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                             : "+f"(cacc[ct][0]), "+f"(cacc[ct][1]), "+f"(cacc[ct][2]), "+f"(cacc[ct][3])
                                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bl[ks][0]), "r"(bl[ks][1]));
                        }
                    }
                }
                // cells of this pass whose last row lies in this tile: normalise, store (rows of `pooled` are the compact ranks), reset
                unsigned done = 0;
                {
                    const int rel = (my_rk & 0xffff) - base;
                    if (my_rk >= 0 && (my_rk & 0x10000) && rel >= 0 && rel < 8) done = 1u << ((rel + base) & 7);
                    done = __reduce_or_sync(0xffffffffu, done);
                }
                if (done) {
                    // weight sums of slots 2tq and 2tq+1 live in the lanes with g = 2tq, 2tq+1
                    const float sa = __shfl_sync(0xffffffffu, s_slot, (2 * tq) * 4), sb = __shfl_sync(0xffffffffu, s_slot, (2 * tq + 1) * 4);
                    __half* out_b = p.pooled + static_cast<size_t>(t.b) * n_cells * D + pw * 128 + g;
                    if ((done >> (2 * tq)) & 1u) {
                        const float fin = 1.0f / sa;
                        __half* orow = out_b + static_cast<size_t>(base + ((2 * tq - base) & 7)) * D;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            orow[i * 16] = __float2half_rn(cacc[i][0] * fin);
                            orow[i * 16 + 8] = __float2half_rn(cacc[i][2] * fin);
                            cacc[i][0] = cacc[i][2] = 0.f;
                        }
                    }
                    if ((done >> (2 * tq + 1)) & 1u) {
                        const float fin = 1.0f / sb;
                        __half* orow = out_b + static_cast<size_t>(base + ((2 * tq + 1 - base) & 7)) * D;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            orow[i * 16] = __float2half_rn(cacc[i][1] * fin);
                            orow[i * 16 + 8] = __float2half_rn(cacc[i][3] * fin);
                            cacc[i][1] = cacc[i][3] = 0.f;
                        }
                    }
                    if ((done >> g) & 1u) s_slot = 0.f;
                }
            }
            if (p.dbg) { w_a += c1 - c0; c_loop += clock64() - c1; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_empty[buf]);
            ++it;
        }
        if (p.dbg && tid == POOL_FIXED_THREADS) {
            long long* d = p.dbg + blockIdx.x * 16 + 10;
            d[0] = clock64() - t_begin; d[1] = w_a; d[2] = c_loop;
        }
    } else if (warp == POOL_POOL_WARP0) {
        // ------------------------------------------------------------ TC mode: issuer of the pooling MMAs
        // out[cell, :] = sum_r p[r] x[r, :] as  D[128 dims, 16] (+)= X^T[128 dims, 16 rows] . W[16 rows, 16 = 8 slots x (hi, lo)]
        // per 128-dim block and 16-row half of the tile: the A operand is the resident feature tile read MN-major (no transpose,
        // no second copy), the accumulators (NPW x 16 tensor-memory columns) are read back by the epilogue warps after every pass.
        constexpr uint32_t idesc_p = umma_idesc_f16_amn(128, 16);
        int it = 0;
        uint32_t pc = 0;                                  // passes issued so far (phase of pacc_empty / pacc_full)
        while (wk.next(t)) {
            const int buf = uniform_i32(it % POOL_NBUF);
            const uint32_t ph = (it / POOL_NBUF) & 1;
            mbar_wait_guard(&p_full[buf], ph);
            if (p.max_only) {                              // first pass of a long text: only drain the ring
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_empty[buf]);
                ++it;
                continue;
            }
            const int npass = uniform_i32(((s_meta[buf * 2 + 1] - s_meta[buf * 2]) >> 3) + 1);
            const uint32_t tile_s = smem_u32(sA) + buf * L::A_BYTES;
            const uint32_t w_s = smem_u32(sW) + buf * (POOL_MAXPASS * POOL_W_BYTES);
            for (int ps = 0; ps < npass; ++ps, ++pc) {     // pass = 8 consecutive cell ranks (normally one per tile)
                mbar_wait_guard(pacc_empty, (pc & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {       // 16 tile rows per instruction
                        const uint64_t db = umma_desc_noswizzle_kmajor(w_s + ps * POOL_W_BYTES + kk * 256, 128, 512);
#pragma unroll
                        for (int mb = 0; mb < NPW; ++mb) {
                            const uint64_t da = umma_desc_sw128_mnmajor_lbo(tile_s + mb * 2 * L::A_CHUNK + kk * 2048, L::A_CHUNK);
                            umma_f16_ss(tmem_base + P_COL0 + mb * 16, da, db, idesc_p, kk ? 1u : 0u);
                        }
                    }
                    umma_commit(pacc_full);
                    if (ps == npass - 1) umma_commit(&a_empty[buf]);     // the tile buffer is free once these MMAs have read it
                }
                __syncwarp();
            }
            ++it;
        }
    } else if (warp > POOL_POOL_WARP0 && warp <= POOL_POOL_WARP0 + 4) {
        // ------------------------------------------------------------ TC mode: accumulator epilogue, one warp per TMEM lane quadrant
        // thread (quadrant q, lane) owns feature dims mb * 128 + q * 32 + lane of every 128-dim block; the fp32 sums of the (up to
        // 8) open cell slots stay in registers across tiles, rescaled when a later tile raises the open cell's max
        const int q = warp & 3;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + P_COL0;
        float acc[NPW][8];
        float ssum[8];                                     // sum of the weights of every slot (replicated in all threads)
#pragma unroll
        for (int s_ = 0; s_ < 8; ++s_) {
            ssum[s_] = 0.f;
#pragma unroll
            for (int mb = 0; mb < NPW; ++mb) acc[mb][s_] = 0.f;
        }
        int it = 0;
        uint32_t pc = 0;
        long long c_loop = 0;
        while (!p.max_only && wk.next(t)) {
            const int buf = it % POOL_NBUF;
            const uint32_t ph = (it / POOL_NBUF) & 1;
            const long long c0 = p.dbg ? clock64() : 0;
            mbar_wait_guard(&p_full[buf], ph);             // s_p / s_cid / s_scal of the tile are visible
            const long long c1 = p.dbg ? clock64() : 0;
            const int my_rk = s_cid[buf * POOL_ROWS + lane];      // compact cell rank | (last row of its cell ? 0x10000 : 0); -1 = no row
            const int rank0 = __shfl_sync(0xffffffffu, my_rk, 0) & 0xffff;
            const int rank_last = __shfl_sync(0xffffffffu, my_rk, t.nrows - 1) & 0xffff;
            {
                const float sc = s_scal[buf];              // a later tile raised the open cell's max (1 otherwise): its slot is rank0 & 7
                if (sc != 1.0f) {
                    const int s0 = rank0 & 7;
#pragma unroll
                    for (int s_ = 0; s_ < 8; ++s_)
                        if (s_ == s0) {
                            ssum[s_] *= sc;
#pragma unroll
                            for (int mb = 0; mb < NPW; ++mb) acc[mb][s_] *= sc;
                        }
                }
            }
            const float* wsum = s_wsum + buf * (POOL_MAXPASS * 8);
            int ps = 0;
            for (int base = rank0; base <= rank_last; base += 8, ++ps, ++pc) {
                const int rel = (my_rk & 0xffff) - base;
                const bool in_pass = my_rk >= 0 && rel >= 0 && rel < 8;
                const int my_slot = my_rk & 7;
                const unsigned done = __reduce_or_sync(0xffffffffu, (in_pass && (my_rk & 0x10000)) ? (1u << my_slot) : 0u);
                mbar_wait_guard(pacc_full, pc & 1);
                tc_fence_after();
                constexpr int HB = NPW / 2;                 // two batches of 128-dim blocks: 48 registers in flight instead of 96
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    uint32_t raw[HB][16];
#pragma unroll
                    for (int mb = 0; mb < HB; ++mb) tmem_ld_32x32b_x16(t_lane + (hb * HB + mb) * 16, raw[mb]);
                    tmem_ld_wait();
                    if (hb == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(pacc_empty);     // the accumulators may be overwritten by the next pass
                    }
#pragma unroll
                    for (int mb = 0; mb < HB; ++mb)
#pragma unroll
                        for (int s_ = 0; s_ < 8; ++s_)
                            acc[hb * HB + mb][s_] += __uint_as_float(raw[mb][s_]) + __uint_as_float(raw[mb][8 + s_]);
                }
#pragma unroll
                for (int s_ = 0; s_ < 8; ++s_) ssum[s_] += wsum[ps * 8 + s_];      // weight sums per slot, from the softmax warp
                if (done) {
                    // cells of this pass whose last row lies in this tile: normalise, store (rows of `pooled` are the compact ranks), reset
                    __half* out_b = p.pooled + static_cast<size_t>(t.b) * n_cells * D + q * 32 + lane;
#pragma unroll
                    for (int s_ = 0; s_ < 8; ++s_)
                        if ((done >> s_) & 1u) {
                            const float fin = 1.0f / ssum[s_];
                            __half* orow = out_b + static_cast<size_t>(base + ((s_ - base) & 7)) * D;
#pragma unroll
                            for (int mb = 0; mb < NPW; ++mb) {
                                orow[mb * 128] = __float2half_rn(acc[mb][s_] * fin);
                                acc[mb][s_] = 0.f;
                            }
                            ssum[s_] = 0.f;
                        }
                }
            }
            if (p.dbg) { w_a += c1 - c0; c_loop += clock64() - c1; }
            ++it;
        }
        if (p.dbg && warp == POOL_POOL_WARP0 + 1 && lane == 0) {
            long long* d = p.dbg + blockIdx.x * 16 + 10;
            d[0] = clock64() - t_begin; d[1] = w_a; d[2] = c_loop;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == POOL_MMA_WARP) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, POOL_TMEM_COLS);
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

template <int D, bool TC>
static int launch_pool(const void* fts, long long fts_rows, const void* text_fts, void* text_ws, int text_ws_ready, int grid,
                       PoolParams& p, float* w_scratch, cudaStream_t stream) {
    // gather4 tensor map: the slab as [fts_rows, D] fp16, box = 64 columns x 1 row (the instruction names 4 rows)
    CUtensorMap tm;
    const int rc = make_tmap_f16_2d(&tm, fts, static_cast<uint64_t>(D), static_cast<uint64_t>(fts_rows), static_cast<uint64_t>(D) * 2, 64, 1);
    if (rc) return rc;
    constexpr int smem = PoolSmem<D>::TOTAL;
    constexpr int threads = TC ? POOL_TC_THREADS : POOL_FIXED_THREADS + D / 4;
    GMM_CUDA_CHECK(cudaFuncSetAttribute(pool_kernel<D, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
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
        GMM_CUDA_CHECK(launch_pdl(pool_kernel<D, TC>, dim3(grid), dim3(threads), smem, stream, tm, p1));
        gridmm_count_launch(1);
        p.l_pad = 128; p.w_in = w_scratch;
    }
    GMM_CUDA_CHECK(launch_pdl(pool_kernel<D, TC>, dim3(grid), dim3(threads), smem, stream, tm, p));
    gridmm_count_launch(1);
    return 0;
}

}  // namespace gmm

static long long* g_pool_dbg = nullptr;
static int g_pool_split = 1;
// Debug hook: 0 = single fp16 softmax weights in the mma.sync weighted sums (half the HMMA count), 1 = value + residual.
extern "C" void gridmm_debug_set_pool_split(int on) { g_pool_split = on; }
static int g_pool_hmma = 1;
// Weighted-sum stage: 1 (default) = warp-level mma.sync from the resident tile, accumulators in registers; 0 = tcgen05 (MN-major A
// operand out of the same tile, weight operand built by the softmax warp, accumulators in tensor memory read back by four epilogue
// warps).  Both are parity-tested.  Measured at B = 32, T = 8 (profiles/r2_pool_modes.md): mma.sync 67.8-68.4 us; tcgen05 72.0-77.2 us
// in three layouts -- next to the 384 columns of the text operand tensor memory only holds EITHER all six pooling accumulators and
// one relevance accumulator (the relevance MMA then waits for the reducers 35 % of the time) OR two relevance accumulators and half
// of the pooling blocks (two MMA -> epilogue round trips per tile), and the per-tile weight operand adds ~500 cycles to the softmax
// warp, which paces the kernel.
extern "C" void gridmm_debug_set_pool_hmma(int on) { g_pool_hmma = on; }
// Debug hook (tools/microbench2.py): per-CTA cycle counters [grid][16] written by the next pool launches; null disables.
extern "C" void gridmm_debug_set_pool_counters(long long* dbg) { g_pool_dbg = dbg; }

extern "C" int gridmm_pool(const void* fts, long long fts_rows, int feat_dim, const int* slots, int t_cap, int slot_rows,
                           int view_rows, int tok_off, const int* perm, int cap, const int* cell_start, const int* cell_rank,
                           int n_cells, const void* text_fts, int l_pad, int batch, void* text_ws, int text_ws_ready,
                           void* pooled, float* w_out, float* w_scratch, int num_ctas, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (!fts || !slots || !perm || !cell_start || !cell_rank || !text_ws || !pooled) return GRIDMM_ERR_ARG;
    if (!text_ws_ready && !text_fts) return GRIDMM_ERR_ARG;
    if (batch > POOL_MAX_BATCH || n_cells > POOL_MAX_CELLS || l_pad < 1 || l_pad > 256 || fts_rows <= 0) return GRIDMM_ERR_SHAPE;
    if (l_pad > 128 && !w_scratch) return GRIDMM_ERR_ARG;
    if (feat_dim != 768 && feat_dim != 512) return GRIDMM_ERR_SHAPE;
    if ((reinterpret_cast<uintptr_t>(text_fts) & 15) || (reinterpret_cast<uintptr_t>(text_ws) & 15)) return GRIDMM_ERR_SHAPE;
    PoolParams p;
    p.slots = slots; p.perm = perm; p.cell_start = cell_start; p.cell_rank = cell_rank;
    p.text_ws = reinterpret_cast<const uint4*>(text_ws);
    p.pooled = reinterpret_cast<__half*>(pooled); p.w_out = w_out; p.w_in = nullptr; p.max_only = 0; p.split_weights = g_pool_split;
    p.batch = batch; p.t_cap = t_cap; p.cap = cap; p.n_cells = n_cells; p.l_pad = l_pad;
    p.slot_rows = slot_rows; p.view_rows = view_rows; p.tok_off = tok_off; p.dbg = g_pool_dbg;
    const int sms = gridmm_sm_count();
    if (sms <= 0) return GRIDMM_ERR_DRIVER;
    const int grid = num_ctas > 0 ? num_ctas : sms;
    if (g_pool_hmma) {
        if (feat_dim == 768) return launch_pool<768, false>(fts, fts_rows, text_fts, text_ws, text_ws_ready, grid, p, w_scratch, stream);
        return launch_pool<512, false>(fts, fts_rows, text_fts, text_ws, text_ws_ready, grid, p, w_scratch, stream);
    }
    if (feat_dim == 768) return launch_pool<768, true>(fts, fts_rows, text_fts, text_ws, text_ws_ready, grid, p, w_scratch, stream);
    return launch_pool<512, true>(fts, fts_rows, text_fts, text_ws, text_ws_ready, grid, p, w_scratch, stream);
}
