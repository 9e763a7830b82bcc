// Grid build: depth back-projection of the 588 new patch points of every episode,
// running bounds, egocentric window, cell-index assignment of ALL accumulated points, then a stable
// counting sort of the valid points by cell (the layout the pooling kernel streams) and the
// cell-centre polar features.
//
// Replaces, per episode and per step,
//   get_rel_position                 map_nav_src/r2r/env.py:115-121
//   EnvBatch.getGlobalMap            map_nav_src/r2r/env.py:267-374   (the 196-iteration compare loop :366-369)
//   EnvBatch.get_gridmap_pos_fts     map_nav_src/r2r/env.py:242-265
// and the continuous-env variant VLN_CE/vlnce_baselines/models/Policy_ViewSelection_GridMap.py:632-641,689-825.
//
// Arithmetic contract (bit-exact cell ids): every fp32 operation is individually rounded
// (__fadd_rn/__fmul_rn/__fdiv_rn keep the compiler from contracting to FMA), trigonometry arrives from
// the host already evaluated in double and rounded to fp32, int conversion truncates toward zero.
//
// One CTA (1024 threads) per episode; all episodes of the batch in one launch.  Integer/fp32 work over
// 11 bytes per point -- launch-latency bound, so it is batched rather than tiled.
#include "common.cuh"
#include "host_util.h"

namespace gmm {

constexpr int GRID_THREADS = 1024;
constexpr int PTS_PER_VP = 588;   // 12 views x 7x7 patch centres
constexpr int MAX_CELLS = 256;

struct GridParams {
    // per-step inputs
    const void* depth;         // [B, 588] uint16 (depth_is_f32 = 0) or float32 metres (= 1)
    const float* pose;         // [B, 4]  px, py, cos(map angle), sin(map angle)      (fp32-rounded on host)
    const float* view_cs;      // [B, 12, 2] cos/sin of every view's heading            (fp32-rounded on host)
    const uint8_t* active;     // [B] or null: 0 = episode receives no new viewpoint this step (cells are still re-assigned)
    const int* new_slot;       // [B] or null: feature-slab slot of this step's viewpoint (device-resident feature DB): written
    int* slots; int t_cap;     //   into slots[b, n_pts / 588], the table gridmm_pool resolves rows through
    // persistent per-episode state
    float* wx;                 // [B, cap]
    float* wy;                 // [B, cap]
    uint8_t* valid;            // [B, cap]
    float* bounds;             // [B, 4] max_x, min_x, max_y, min_y
    int* n_pts;                // [B] points accumulated so far (updated here)
    // outputs
    int16_t* cell;             // [B, cap]  -1 = masked
    float* half_len;           // [B]
    int* perm;                 // [B, cap]  point indices of the valid points, sorted by cell (stable)
    int* cell_start;           // [B, n_cells + 1]
    int* cell_rank;            // [B, n_cells]  rank among non-empty cells, -1 if empty
    int* n_nonempty;           // [B]
    float* pos_fts;            // [B, n_cells, 5]
    int cap;                   // capacity in points per episode
    int grid_w;                // 14
    int depth_is_f32;
    float depth_scale;         // 4000 (R2R uint16 0.25 mm units) or 1 (CE metres)
    float off[7];              // f32(o_k) * f32(tan(hfov/2)), o = -6/7 .. 6/7
    int flip_y;                // CE: world_y = -rel_y + py
    int negate_map_x;          // CE: map_x = -(...)
    int sort_only;             // 1: `cell` and `n_pts` are inputs (reference-format grid_map); only steps 3-4 run
    int pos_mode;              // 0: discrete-env polar features (env.py:242-265); 1: the CE code's (x, z, y) convention
    float max_dist;            // distance normaliser: 30 (env.py:47), 25 / 40 (R2R-CE / RxR-CE)
};

__device__ __forceinline__ float block_reduce_max(float v, float* red, int tid) {
    v = warp_max(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        float x = red[tid];
        x = warp_max(x);
        if (tid == 0) red[0] = x;
    }
    __syncthreads();
    const float r = red[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(GRID_THREADS, 1) grid_update_kernel(GridParams p) {
    __shared__ float red[32];
    __shared__ float s_scalar[8];
    __shared__ uint16_t warp_hist[32][MAX_CELLS];   // per-warp counts, then per-warp write cursors
    __shared__ int s_cell_start[MAX_CELLS + 1];

    pdl_wait();
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n_cells = p.grid_w * p.grid_w;
    float* wx = p.wx + static_cast<size_t>(b) * p.cap;
    float* wy = p.wy + static_cast<size_t>(b) * p.cap;
    uint8_t* valid = p.valid + static_cast<size_t>(b) * p.cap;
    int16_t* cell = p.cell + static_cast<size_t>(b) * p.cap;
    int* perm = p.perm + static_cast<size_t>(b) * p.cap;

    const bool full = (p.sort_only == 0);
    const float px = full ? p.pose[b * 4 + 0] : 0.f, py = full ? p.pose[b * 4 + 1] : 0.f;
    const float mc = full ? p.pose[b * 4 + 2] : 0.f, ms = full ? p.pose[b * 4 + 3] : 0.f;
    const int n_old = p.n_pts[b];
    const bool is_active = full && ((p.active == nullptr) || (p.active[b] != 0));
    const int n_new = is_active ? PTS_PER_VP : 0;
    const int n = n_old + n_new;

    // ---- 1. back-project the new viewpoint (env.py:115-121, 289-294) and reduce its bounds
    float lx_max = -3.0e38f, lx_min = 3.0e38f, ly_max = -3.0e38f, ly_min = 3.0e38f;
    if (tid < n_new && n <= p.cap) {
        const int v = tid / 49, k = tid % 49;
        float d;
        bool ok;
        if (p.depth_is_f32) {
            d = reinterpret_cast<const float*>(p.depth)[b * PTS_PER_VP + tid];
            ok = (d != 0.0f);
        } else {
            const uint16_t u = reinterpret_cast<const uint16_t*>(p.depth)[b * PTS_PER_VP + tid];
            d = static_cast<float>(u);
            ok = (u != 0);
        }
        const float dy = (p.depth_scale != 1.0f) ? __fdiv_rn(d, p.depth_scale) : d;
        const float dx = __fmul_rn(dy, p.off[k % 7]);
        const float ca = p.view_cs[(b * 12 + v) * 2 + 0], sa = p.view_cs[(b * 12 + v) * 2 + 1];
        const float rel_x = __fadd_rn(__fmul_rn(dx, ca), __fmul_rn(dy, sa));
        const float rel_y = __fsub_rn(__fmul_rn(dy, ca), __fmul_rn(dx, sa));
        const float gx = __fadd_rn(rel_x, px);
        const float gy = p.flip_y ? __fadd_rn(-rel_y, py) : __fadd_rn(rel_y, py);
        wx[n_old + tid] = gx;
        wy[n_old + tid] = gy;
        valid[n_old + tid] = ok ? 1 : 0;
        lx_max = lx_min = gx;   // bounds include masked points (env.py:307-319)
        ly_max = ly_min = gy;
    }
    lx_max = block_reduce_max(lx_max, red, tid);
    lx_min = -block_reduce_max(-lx_min, red, tid);
    ly_max = block_reduce_max(ly_max, red, tid);
    ly_min = -block_reduce_max(-ly_min, red, tid);

    if (tid == 0 && full) {
        float max_x = p.bounds[b * 4 + 0], min_x = p.bounds[b * 4 + 1];
        float max_y = p.bounds[b * 4 + 2], min_y = p.bounds[b * 4 + 3];
        if (n_new > 0 && n <= p.cap) {
            if (lx_max > max_x) max_x = lx_max;
            if (lx_min < min_x) min_x = lx_min;
            if (ly_max > max_y) max_y = ly_max;
            if (ly_min < min_y) min_y = ly_min;
            p.bounds[b * 4 + 0] = max_x; p.bounds[b * 4 + 1] = min_x;
            p.bounds[b * 4 + 2] = max_y; p.bounds[b * 4 + 3] = min_y;
            p.n_pts[b] = n;
            if (p.new_slot && n_old / PTS_PER_VP < p.t_cap) p.slots[b * p.t_cap + n_old / PTS_PER_VP] = p.new_slot[b];
        }
        // window (env.py:322-331)
        const float ax = __fsub_rn(px, min_x), bx = __fsub_rn(max_x, px);
        const float xh = (ax > bx) ? ax : bx;
        const float ay = __fsub_rn(py, min_y), by = __fsub_rn(max_y, py);
        const float yh = (ay > by) ? ay : by;
        float half = (xh > yh) ? xh : yh;
        half = __fdiv_rn(__fmul_rn(half, 2.0f), 3.0f);
        s_scalar[0] = half;
        p.half_len[b] = half;
    }
    for (int i = tid; i < 32 * MAX_CELLS; i += GRID_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const int n_eff = min((n <= p.cap) ? n : n_old, p.cap);   // the host wrapper grows the buffers before they overflow
    const float half = full ? s_scalar[0] : 1.0f;
    const float two_half = __fmul_rn(2.0f, half);
    const float scale = static_cast<float>(p.grid_w - 1);

    // ---- 2. cell assignment of every accumulated point (env.py:337-369); each warp owns a contiguous range
    const int per_warp = (n_eff + 31) / 32;
    const int w_lo = min(warp * per_warp, n_eff), w_hi = min(w_lo + per_warp, n_eff);
    for (int j0 = w_lo; j0 < w_hi; j0 += 32) {
        const int j = j0 + lane;
        int c = -1;
        if (j < w_hi && !full) c = cell[j];
        if (j < w_hi && full) {
            const float tx = __fsub_rn(wx[j], px), ty = __fsub_rn(wy[j], py);
            float mx = __fadd_rn(__fmul_rn(tx, mc), __fmul_rn(ty, ms));
            const float my = __fsub_rn(__fmul_rn(ty, mc), __fmul_rn(tx, ms));
            if (p.negate_map_x) mx = -mx;
            int ix = __float2int_rz(__fmul_rn(__fdiv_rn(__fadd_rn(mx, half), two_half), scale));
            int iy = __float2int_rz(__fmul_rn(__fdiv_rn(__fadd_rn(my, half), two_half), scale));
            ix = min(max(ix, 0), p.grid_w - 1);
            iy = min(max(iy, 0), p.grid_w - 1);
            c = valid[j] ? (ix * p.grid_w + iy) : -1;
            cell[j] = static_cast<int16_t>(c);
        }
        // per-warp histogram; lanes of a warp that share a cell are combined first (deterministic counts)
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (c >= 0 && lane == (__ffs(peers) - 1)) warp_hist[warp][c] += static_cast<uint16_t>(__popc(peers));
        __syncwarp();
    }
    __syncthreads();

    // ---- 3. exclusive scan over (cell, warp): cell-major so the sort is stable in point order
    if (tid < MAX_CELLS) {
        int tot = 0;
        if (tid < n_cells)
            for (int w = 0; w < 32; ++w) tot += warp_hist[w][tid];
        s_cell_start[tid] = tot;    // counts for now
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0, rank = 0;
        for (int c = 0; c < n_cells; ++c) {
            const int cnt = s_cell_start[c];
            s_cell_start[c] = run;
            p.cell_rank[b * n_cells + c] = cnt > 0 ? rank : -1;
            rank += cnt > 0 ? 1 : 0;
            run += cnt;
        }
        s_cell_start[n_cells] = run;
        p.n_nonempty[b] = rank;
    }
    __syncthreads();
    if (tid <= n_cells) p.cell_start[b * (n_cells + 1) + tid] = s_cell_start[tid];
    if (tid < n_cells) {
        int run = s_cell_start[tid];
        for (int w = 0; w < 32; ++w) {
            const int cnt = warp_hist[w][tid];
            warp_hist[w][tid] = static_cast<uint16_t>(run);   // cap <= 65535 points per episode
            run += cnt;
        }
    }
    __syncthreads();

    // ---- 4. stable scatter: each warp walks its range in order, 32 points at a time
    for (int j0 = w_lo; j0 < w_hi; j0 += 32) {
        const int j = j0 + lane;
        const int c = (j < w_hi) ? static_cast<int>(cell[j]) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (c >= 0) {
            const int leader = __ffs(peers) - 1;
            const int before = __popc(peers & ((1u << lane) - 1u));
            const int base = warp_hist[warp][c];
            perm[base + before] = j;
            __syncwarp(peers);
            if (lane == leader) warp_hist[warp][c] = static_cast<uint16_t>(base + __popc(peers));
        }
        __syncwarp();
    }

    // ---- 5. cell-centre polar features (env.py:242-265, 60-77, 52-58): [sin h, cos h, sin e, cos e, dist/30]
    if (tid < n_cells && full) {
        const int i = tid / p.grid_w, jj = tid % p.grid_w;
        const float cell_len = __fdiv_rn(__fmul_rn(half, 2.0f), static_cast<float>(p.grid_w));
        const float hc = __fdiv_rn(cell_len, 2.0f);
        const float x = __fadd_rn(__fsub_rn(__fmul_rn(static_cast<float>(i), cell_len), half), hc);
        const float y = __fadd_rn(__fsub_rn(__fmul_rn(static_cast<float>(jj), cell_len), half), hc);
        const float r = fmaxf(sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))), 1e-8f);
        float* o = p.pos_fts + (static_cast<size_t>(b) * n_cells + tid) * 5;
        if (p.pos_mode == 0) {
            float h = asinf(__fdiv_rn(x, r));
            if (y < 0.0f) h = 3.14159265358979323846f - h;
            o[0] = sinf(h);
            o[1] = cosf(h);
            o[2] = 0.0f;   // elevation = asin(0 / r) = 0
            o[3] = 1.0f;
        } else {
            // VLN_CE/vlnce_baselines/models/utils.py:125-144 reads (x, z, y): the cell's second coordinate is taken as the
            // HEIGHT, so heading = asin(x / |x|) = +-pi/2 and elevation = asin(y / r)  (Policy_ViewSelection_GridMap.py:661-684)
            const float h = asinf(__fdiv_rn(x, fmaxf(fabsf(x), 1e-8f)));
            const float e = asinf(__fdiv_rn(y, r));
            o[0] = sinf(h);
            o[1] = cosf(h);
            o[2] = sinf(e);
            o[3] = cosf(e);
        }
        o[4] = r / p.max_dist;
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

}  // namespace gmm

extern "C" int gridmm_grid_update(int batch, const void* depth, int depth_is_f32, float depth_scale, const float* pose,
                                  const float* view_cs, const unsigned char* active, const float* off7, int flip_y,
                                  int negate_map_x, int pos_mode, float max_dist, int grid_w, int cap, float* wx, float* wy, unsigned char* valid,
                                  float* bounds, int* n_pts, short* cell, float* half_len, int* perm, int* cell_start,
                                  int* cell_rank, int* n_nonempty, float* pos_fts, const int* new_slot, int* slots, int t_cap,
                                  cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (grid_w < 2 || grid_w * grid_w > MAX_CELLS || cap < PTS_PER_VP || cap > 65535) return GRIDMM_ERR_SHAPE;
    if (new_slot && (!slots || t_cap < 1)) return GRIDMM_ERR_ARG;
    if (!depth || !pose || !view_cs || !off7 || !wx || !wy || !valid || !bounds || !n_pts || !cell || !half_len || !perm ||
        !cell_start || !cell_rank || !n_nonempty || !pos_fts)
        return GRIDMM_ERR_ARG;
    GridParams p;
    p.depth = depth; p.pose = pose; p.view_cs = view_cs; p.active = active;
    p.new_slot = new_slot; p.slots = slots; p.t_cap = t_cap;
    p.wx = wx; p.wy = wy; p.valid = valid; p.bounds = bounds; p.n_pts = n_pts;
    p.cell = cell; p.half_len = half_len; p.perm = perm; p.cell_start = cell_start; p.cell_rank = cell_rank;
    p.n_nonempty = n_nonempty; p.pos_fts = pos_fts;
    p.cap = cap; p.grid_w = grid_w; p.depth_is_f32 = depth_is_f32; p.depth_scale = depth_scale;
    for (int i = 0; i < 7; ++i) p.off[i] = off7[i];
    p.flip_y = flip_y; p.negate_map_x = negate_map_x; p.sort_only = 0; p.pos_mode = pos_mode; p.max_dist = max_dist;
    GMM_CUDA_CHECK(launch_pdl(grid_update_kernel, dim3(batch), dim3(GRID_THREADS), 0, stream, p));
    gridmm_count_launch(1);
    return 0;
}

// Sort-only entry for callers that already hold the reference's `grid_map` (cell id per point, -1 = masked):
// produces the same perm / cell_start / cell_rank / n_nonempty as gridmm_grid_update.
extern "C" int gridmm_cell_sort(int batch, const short* cell, const int* n_pts, int grid_w, int cap, int* perm,
                                int* cell_start, int* cell_rank, int* n_nonempty, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (grid_w < 2 || grid_w * grid_w > MAX_CELLS || cap < 1 || cap > 65535) return GRIDMM_ERR_SHAPE;
    if (!cell || !n_pts || !perm || !cell_start || !cell_rank || !n_nonempty) return GRIDMM_ERR_ARG;
    GridParams p = {};
    p.n_pts = const_cast<int*>(n_pts);
    p.cell = const_cast<short*>(cell);
    p.perm = perm; p.cell_start = cell_start; p.cell_rank = cell_rank; p.n_nonempty = n_nonempty;
    p.cap = cap; p.grid_w = grid_w; p.depth_scale = 1.0f; p.sort_only = 1;
    GMM_CUDA_CHECK(launch_pdl(grid_update_kernel, dim3(batch), dim3(GRID_THREADS), 0, stream, p));
    gridmm_count_launch(1);
    return 0;
}
