// tcgen05 GEMM for the cross-modal encoder blocks:
//   C[M,N] = epilogue( A[M,K] (fp16, K contiguous) x W[N,K]^T (fp16, K contiguous) ),  fp32 accumulation in TMEM.
// This is nn.Linear: the reference's BertSelfAttention/BertOutAttention projections
// (map_nav_src/models/vilmodel.py:95-153, 317-368), BertIntermediate/BertOutput (:184-209),
// nn.MultiheadAttention in/out projections + linear1/linear2 (map_nav_src/models/transformer.py:133-182),
// ClsPrediction first layer (vilmodel.py:663-674), text_proj/grid_proj (:702-703).
//
// Structure: persistent CTAs (one per SM), 128 x BN output tiles (BN = 128 or 256) handed out round-robin -- or, for large
// problems, CTA pairs (tcgen05 cta_group::2) on 256 x 256 tiles, each CTA loading half of the W tile -- 10 warps:
//   warp 0      TMA producer: 128 x BK A tile + (BN / CG) x BK W tile per stage (BK = 64 or 128), SWIZZLE_128B, mbarrier
//               complete_tx; the stage ring runs straight across tile boundaries.  The loop runs warp-converged and issues under
//               elect.sync (common.cuh: a divergent single-thread loop costs ~100 cycles per TMA / MMA instruction)
//   warp 1      allocates TMEM (2 accumulators of BN columns), issues tcgen05.mma (one elected lane), commits smem stages
//               back to the producer and finished accumulators to the epilogue
//   warps 2..9  epilogue (two warps per TMEM lane quadrant, half of the columns each): the tile's bias slice is staged in
//               shared memory and the first residual chunks are already in flight BEFORE the accumulator is ready;
//               tcgen05.ld 32 lanes x 16 columns, bias / GELU / ReLU / residual (register-pipelined 4 chunks deep),
//               fp32 and/or fp16 stores; accumulator t is drained while the MMA warp is already filling t+1
#include "common.cuh"
#include "host_util.h"

namespace gmm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_KATOM = 64;         // one SWIZZLE_128B atom = 64 fp16 along K
constexpr int GEMM_EPI_WARPS = 8;       // 2 per TMEM lane quadrant, each takes half of the tile's columns
constexpr int GEMM_THREADS = 64 + GEMM_EPI_WARPS * 32;
constexpr int GEMM_RES_PREFETCH = 4;    // residual chunks (16 columns each) kept in flight per epilogue thread

struct GemmEpilogue {
    const float* bias;       // [N] or null
    const float* residual;   // [M, ld_res] fp32 or null (added after the activation)
    float* out_f32;          // [M, ld_f32] or null
    __half* out_f16;         // [M, ld_f16] or null
    int ld_res, ld_f32, ld_f16;
    int act;                 // 0 none, 1 GELU(erf), 2 ReLU
    __half* out_lanes;       // optional lane-major fp16 output [lanes_rows / 128 blocks][M / lanes_rows, N / 8, 128] of 16-byte units (gridmm_pool's
    int lanes_rows;          // text operand layout): row = b * lanes_rows + t  ->  unit u of it at ((b * N/8 + u) * 128 + t)
    // grouped ClsPrediction mode (gridmm_cls_heads_f16): per 128-row tile {first A row, first W / bias row, first output row};
    // the epilogue keeps only three sums per row and 64-column slice (see cls_part) instead of storing the activations
    const int* grp;          // [tiles_m][4] or null: first A row, first W row, first output row, mode (0: ReLU + sums, 1: raw fp32)
    const float* gw2;        // [groups * N] gamma * w2 of every head (indexed like bias), cls mode only
    float* cls_part;         // [rows][N / 64][3]: sum r, sum r^2, sum r * gw2 with r = act(acc + bias), or null
    float* cls_raw;          // [rows][N] plain products of the mode-1 tiles (K-split halves of sap_fuse_linear), or null
    const int* m_dev;        // optional device-side row count (<= M): packed / ragged operands whose size only the GPU knows
    long long* dbg;          // optional [grid][8] cycle counters (tools/microbench2.py), null in production
};

template <int BN, int STAGES, int BK, int CG = 1>
struct GemmSmem {
    static constexpr int A_ATOM = GEMM_BM * GEMM_KATOM * 2;     // 128 rows x 128 B
    static constexpr int B_ATOM = (BN / CG) * GEMM_KATOM * 2;   // a CTA pair splits the W tile: BN / 2 rows each
    static constexpr int A_BYTES = A_ATOM * (BK / GEMM_KATOM);
    static constexpr int B_BYTES = B_ATOM * (BK / GEMM_KATOM);
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;   // [2][BN] floats
    static constexpr int GW2_OFFSET = BIAS_OFFSET + 2 * BN * 4;    // [2][BN] floats (cls mode)
    static constexpr int STG_OFFSET = GW2_OFFSET + 2 * BN * 4;      // per epilogue warp: 32 rows x 20 floats (16 + pad)
    static constexpr int STG_WARP_BYTES = 32 * 20 * 4;
    static constexpr int TOTAL = STG_OFFSET + GEMM_EPI_WARPS * STG_WARP_BYTES + 1024;
};

// erf-GELU (BertIntermediate / F.gelu: 0.5 x (1 + erf(x / sqrt 2))).  erf by Abramowitz & Stegun 7.1.28,
//   erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16,  |error| <= 3e-7,
// i.e. 6 FMA + ONE MUFU (rcp) + 4 squarings per element.  The epilogue of the FFN1 GEMM applies it to every element and is
// MUFU-bound: a 128 x 256 tile is 32 k elements against 16 MUFU results per clock per SM, so the earlier 7.1.26 form
// (rcp + ex2 = 2 MUFU, 4 k cycles per tile) cost ~6 us per FFN1 launch more than the mainloop could hide.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float p = fmaf(0.0000430638f, z, 0.0002765672f);
    p = fmaf(p, z, 0.0001520143f);
    p = fmaf(p, z, 0.0092705272f);
    p = fmaf(p, z, 0.0422820123f);
    p = fmaf(p, z, 0.0705230784f);
    p = fmaf(p, z, 1.0f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));      // p >= 1; overflow of p -> r = 0 -> erf = 1
    r *= r; r *= r; r *= r; r *= r;                              // p^-16 = 1 - erf(z)
    const float erf_abs = 1.0f - r;
    return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// fp16 outputs SATURATE at +-65504 instead of overflowing to inf: the reference computes these activations in fp32, where a
// large QKV / FFN1 value is harmless; as an fp16 GEMM operand an inf would turn the next layer into NaN.  (With trained
// BERT-size weights |QKV| and |GELU(FFN1)| stay below a few hundred; tests/test_gpu_nav.py::test_trained_scale_activations.)
__device__ __forceinline__ float sat_f16(float x) { return fminf(fmaxf(x, -65504.0f), 65504.0f); }
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(sat_f16(a), sat_f16(b));
    return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void named_bar_sync_epi() {
    asm volatile("bar.sync 1, %0;" ::"n"(GEMM_EPI_WARPS * 32) : "memory");
}

template <int BN, int STAGES, int BK, bool DBG, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int M, int N, int K,
                   GemmEpilogue ep) {
    using L = GemmSmem<BN, STAGES, BK, CG>;
    constexpr int ATOMS = BK / GEMM_KATOM;
    constexpr int TILE_M = GEMM_BM * CG;          // CG = 2: a CTA pair owns a 256 x BN tile (tcgen05 cta_group::2)
    // BN = 384 (pairs only): 256 x 384 tiles as two N = 192 MMAs per K step, ONE accumulator of 384 columns (no double buffering:
    // used where the whole problem is a single round of tiles, so there is no next main loop to overlap the epilogue with).
    // Per CTA and k-block 16 KB of A + 24 KB of W arrive for 128 x 384 outputs: 153 flop per delivered byte against 128 for
    // 256 x 256 pair tiles, and M = 1824 x N = 3072 becomes 8 x 8 = 64 pair tiles = one round instead of 96 = two.
    constexpr int NMMA = (BN > 256) ? 2 : 1;
    constexpr int MMA_N = BN / NMMA;
    constexpr int NACC = (BN > 256) ? 1 : 2;
    static_assert(BN <= 256 || CG == 2, "384-wide tiles exist in pair mode only");
    extern __shared__ uint8_t smem_raw[];
    // align by pointer arithmetic on the __shared__ array (an integer round-trip would demote every access below to a
    // generic LD/ST instead of LDS/STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = K / BK;
    const int tiles_n = N / BN;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;     // 0 = leader: owns the full barriers and issues the MMAs
    const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], GEMM_EPI_WARPS * CG);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_2sm(tmem_slot, NACC * BN > 256 ? 512 : NACC * BN); else tmem_alloc(tmem_slot, NACC * BN);
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();     // the peer's barriers must be initialised before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    pdl_wait();      // everything above overlapped the previous kernel's tail; global memory is touched only below
    if (ep.m_dev) M = min(M, __ldg(ep.m_dev));      // packed operand: the row count lives on the device
    const int tiles_m = (M + TILE_M - 1) / TILE_M;
    const int num_tiles = tiles_m * tiles_n;

    if (warp == 0) {
        {   // the whole (converged) warp runs the schedule, one elected lane issues: see elect_one() in common.cuh
            int it = 0;
            long long w_empty = 0;
            const long long t_begin = DBG ? clock64() : 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                // pair mode: this CTA loads its own 128 rows of A and its half of the W tile
                int m0 = (tile / tiles_n) * TILE_M + static_cast<int>(rank) * GEMM_BM;
                int n0 = (tile % tiles_n) * BN + static_cast<int>(rank) * (BN / CG);
                if (ep.grp) {      // grouped heads: this row tile reads its own A rows and its head's weight rows
                    const int* gt = ep.grp + (tile / tiles_n) * 4;
                    m0 = __ldg(gt + 0);
                    n0 += __ldg(gt + 1);
                }
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const long long c0 = DBG ? clock64() : 0;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    if (DBG) w_empty += clock64() - c0;
                    uint8_t* a_dst = smem + s * L::STAGE_BYTES;
                    uint8_t* b_dst = a_dst + L::A_BYTES;
                    if (!elect_one()) continue;
                    if (CG == 1) {
                        mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
#pragma unroll
                        for (int a = 0; a < ATOMS; ++a) {
                            tma_load_2d(a_dst + a * L::A_ATOM, &tmA, kb * BK + a * GEMM_KATOM, m0, &full_bar[s]);
                            tma_load_2d(b_dst + a * L::B_ATOM, &tmW, kb * BK + a * GEMM_KATOM, n0, &full_bar[s]);
                        }
                    } else {
                        // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the pair
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);
                        const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[s]), 0);
#pragma unroll
                        for (int a = 0; a < ATOMS; ++a) {
                            tma_load_2d_2sm(a_dst + a * L::A_ATOM, &tmA, kb * BK + a * GEMM_KATOM, m0, lead_bar);
                            if (NMMA == 1) {
                                tma_load_2d_2sm(b_dst + a * L::B_ATOM, &tmW, kb * BK + a * GEMM_KATOM, n0, lead_bar);
                            } else {
                                // this CTA's half (MMA_N / 2 rows) of each of the two MMAs' W rows
                                const int nt = (tile % tiles_n) * BN;
#pragma unroll
                                for (int j = 0; j < NMMA; ++j)
                                    tma_load_2d_2sm(b_dst + a * L::B_ATOM + j * (MMA_N / 2) * 128, &tmW, kb * BK + a * GEMM_KATOM,
                                                    nt + j * MMA_N + static_cast<int>(rank) * (MMA_N / 2), lead_bar);
                            }
                        }
                    }
                }
            }
            // every load of this CTA is in flight: let the next kernel's CTAs be scheduled (their prologue overlaps our tail; they
            // still wait for this grid to complete in griddepcontrol.wait before touching global memory)
            pdl_launch_dependents();
            if (DBG && ep.dbg && lane == 0) { ep.dbg[blockIdx.x * 8 + 0] = clock64() - t_begin; ep.dbg[blockIdx.x * 8 + 1] = w_empty; }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(TILE_M, MMA_N);
            int it = 0, t = 0;
            long long w_full = 0, w_acc = 0;
            const long long t_begin = DBG ? clock64() : 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step, ++t) {
                const int acc = (NACC == 2) ? (t & 1) : 0;
                const long long c1 = DBG ? clock64() : 0;
                mbar_wait(&tmem_empty_bar[acc], (((NACC == 2) ? (t >> 1) : t) & 1) ^ 1);   // epilogue has drained this accumulator
                if (DBG) w_acc += clock64() - c1;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const long long c0 = DBG ? clock64() : 0;
                    mbar_wait(&full_bar[s], ph);
                    if (DBG) w_full += clock64() - c0;
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
                    if (elect_one()) {
#pragma unroll
                    for (int a = 0; a < ATOMS; ++a) {
                        const uint64_t da = umma_desc_sw128_kmajor(a_addr + a * L::A_ATOM);
                        const uint64_t db = umma_desc_sw128_kmajor(a_addr + L::A_BYTES + a * L::B_ATOM);
#pragma unroll
                        for (int j = 0; j < NMMA; ++j) {
                            const uint64_t dbj = db + static_cast<uint64_t>((j * (MMA_N / 2) * 128) >> 4);
#pragma unroll
                            for (int k = 0; k < GEMM_KATOM / 16; ++k) {
                                // +32 bytes per K=16 step inside the 128-byte swizzled row (start-address field is >> 4)
                                if (CG == 2) umma_f16_ss_2sm(d_tmem + j * MMA_N, da + 2 * k, dbj + 2 * k, idesc, (kb | a | k) ? 1u : 0u);
                                else umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | a | k) ? 1u : 0u);
                            }
                        }
                    }
                    if (CG == 2) umma_commit_2sm(&empty_bar[s]); else umma_commit(&empty_bar[s]);     // frees the stage in both CTAs
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    if (CG == 2) umma_commit_2sm(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
                }
                __syncwarp();
            }
            if (DBG && ep.dbg && lane == 0) {
                ep.dbg[blockIdx.x * 8 + 2] = clock64() - t_begin; ep.dbg[blockIdx.x * 8 + 3] = w_full;
                ep.dbg[blockIdx.x * 8 + 4] = w_acc;
            }
        }
    } else {
        // epilogue warps: TMEM lane quadrant is fixed by warp id % 4; warps 2..5 take columns [0, BN/2), 6..9 the rest
        constexpr int HALF = BN / 2, NCH = HALF / 16, PD = GEMM_RES_PREFETCH;
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int et = threadIdx.x - 64;                  // 0..255
        float* s_bias = reinterpret_cast<float*>(smem + L::BIAS_OFFSET);
        int t = 0;
        long long w_tfull = 0;
        const long long t_begin = DBG ? clock64() : 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step, ++t) {
            const int acc = (NACC == 2) ? (t & 1) : 0;
            int m0 = (tile / tiles_n) * TILE_M + static_cast<int>(rank) * GEMM_BM;
            const int n0 = (tile % tiles_n) * BN;
            int wrow0 = n0;                                // row of W / index of bias for column n0 of this tile
            bool raw_tile = false;
            if (ep.grp) {
                const int* gt = ep.grp + (tile / tiles_n) * 4;
                wrow0 += __ldg(gt + 1);
                m0 = __ldg(gt + 2);                        // output rows of this tile
                raw_tile = __ldg(gt + 3) != 0;
            }
            // stage this tile's bias slice (the slot of tile t-2 is free: all warps passed the barrier of tile t-1)
            float* sb = s_bias + (t & 1) * BN;
            float* sg = reinterpret_cast<float*>(smem + L::GW2_OFFSET) + (t & 1) * BN;
            for (int i = et; i < BN; i += GEMM_EPI_WARPS * 32) {
                sb[i] = ep.bias ? __ldg(ep.bias + wrow0 + i) : 0.0f;
                if (ep.cls_part) sg[i] = __ldg(ep.gw2 + wrow0 + i);
            }
            // Global accesses use a coalesced mapping: in pass i a lane owns row 8*i + lane/4 of the warp's 32 rows and 4
            // of the chunk's 16 columns, so one warp instruction covers 8 rows x 64 contiguous bytes (the TMEM layout --
            // one row per lane -- would touch 32 different lines per instruction).  Accumulator chunks cross over through
            // a per-warp shared-memory tile.
            const int col0 = n0 + half * HALF;
            const int rsub = lane >> 2, csub = (lane & 3) * 4;
            const int row_base = m0 + q * 32 + rsub;
            float4 res[PD][4];
            if (ep.residual) {
#pragma unroll
                for (int c = 0; c < PD && c < NCH; ++c)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rg = row_base + 8 * i;
                        res[c][i] = *reinterpret_cast<const float4*>(ep.residual + static_cast<size_t>(rg < M ? rg : 0) * ep.ld_res +
                                                                     col0 + c * 16 + csub);
                    }
            }
            named_bar_sync_epi();
            const long long c0 = DBG ? clock64() : 0;
            mbar_wait(&tmem_full_bar[acc], ((NACC == 2) ? (t >> 1) : t) & 1);
            if (DBG) w_tfull += clock64() - c0;
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * HALF;
            float* stg = reinterpret_cast<float*>(smem + L::STG_OFFSET + (warp - 2) * L::STG_WARP_BYTES);
            float cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;          // cls mode: this row's sums over the warp's column half
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(t_addr + c * 16, v);
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sb + half * HALF + c * 16 + j);
                    f[j] = __uint_as_float(v[j]) + b4.x; f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
                    f[j + 2] = __uint_as_float(v[j + 2]) + b4.z; f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
                }
                if (raw_tile) {
                    // plain product rows (no activation): 64 contiguous bytes per lane, a handful of rows per launch
                    const int rg = m0 + q * 32 + lane;
                    float4* dst = reinterpret_cast<float4*>(ep.cls_raw + static_cast<size_t>(rg) * N + col0 + c * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    continue;
                }
                if (ep.act == 1) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = gelu_erf(f[j]);
                } else if (ep.act == 2) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
                }
                if (ep.cls_part) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 g4 = *reinterpret_cast<const float4*>(sg + half * HALF + c * 16 + j);
                        cs1 += (f[j] + f[j + 1]) + (f[j + 2] + f[j + 3]);
                        cs2 = fmaf(f[j], f[j], cs2); cs2 = fmaf(f[j + 1], f[j + 1], cs2);
                        cs2 = fmaf(f[j + 2], f[j + 2], cs2); cs2 = fmaf(f[j + 3], f[j + 3], cs2);
                        cs3 = fmaf(f[j], g4.x, cs3); cs3 = fmaf(f[j + 1], g4.y, cs3);
                        cs3 = fmaf(f[j + 2], g4.z, cs3); cs3 = fmaf(f[j + 3], g4.w, cs3);
                    }
                    continue;
                }
                if (ep.out_lanes) {
                    // the TMEM layout (one row per lane) IS the coalesced mapping for the lane-major layout: consecutive
                    // lanes = consecutive text positions = consecutive 16-byte units
                    const int rg = m0 + q * 32 + lane;
                    if (rg < M) {
                        const int bb = rg / ep.lanes_rows, tt = rg - bb * ep.lanes_rows;
                        const int unit = (col0 + c * 16) >> 3;
                        // positions >= 128 of a long text go to the second [batch][N/8][128] block
                        const size_t blk = static_cast<size_t>(tt >> 7) * (M / ep.lanes_rows);
                        uint4* dst = reinterpret_cast<uint4*>(ep.out_lanes) + ((blk + bb) * (N >> 3) + unit) * 128 + (tt & 127);
                        uint4 o0, o1;
                        o0.x = pack_half2(f[0], f[1]); o0.y = pack_half2(f[2], f[3]); o0.z = pack_half2(f[4], f[5]); o0.w = pack_half2(f[6], f[7]);
                        o1.x = pack_half2(f[8], f[9]); o1.y = pack_half2(f[10], f[11]); o1.z = pack_half2(f[12], f[13]); o1.w = pack_half2(f[14], f[15]);
                        dst[0] = o0;
                        dst[128] = o1;
                    }
                    continue;
                }
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * 20 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                __syncwarp();
                const int col = col0 + c * 16 + csub;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 x = *reinterpret_cast<const float4*>(stg + (rsub + 8 * i) * 20 + csub);
                    const int rg = row_base + 8 * i;
                    if (ep.residual) {
                        const float4 r4 = res[c % PD][i];
                        x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
                        if (c + PD < NCH)
                            res[c % PD][i] = *reinterpret_cast<const float4*>(ep.residual + static_cast<size_t>(rg < M ? rg : 0) * ep.ld_res +
                                                                              col + PD * 16);
                    }
                    if (rg < M) {
                        if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + static_cast<size_t>(rg) * ep.ld_f32 + col) = x;
                        if (ep.out_f16) {
                            uint2 o;
                            o.x = pack_half2(x.x, x.y);
                            o.y = pack_half2(x.z, x.w);
                            *reinterpret_cast<uint2*>(ep.out_f16 + static_cast<size_t>(rg) * ep.ld_f16 + col) = o;
                        }
                    }
                }
                __syncwarp();
            }
            if (ep.cls_part && !raw_tile) {
                const int rg = m0 + q * 32 + lane;
                if (rg < M) {
                    float* dst = ep.cls_part + (static_cast<size_t>(rg) * (N / 64) + (n0 + half * HALF) / 64) * 3;
                    dst[0] = cs1; dst[1] = cs2; dst[2] = cs3;
                }
            }
            // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above): hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2 && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));   // leader's barrier
                else mbar_arrive(&tmem_empty_bar[acc]);
            }
        }
        if (DBG && ep.dbg && et == 0) { ep.dbg[blockIdx.x * 8 + 5] = clock64() - t_begin; ep.dbg[blockIdx.x * 8 + 6] = w_tfull; }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();     // neither CTA may exit / free TMEM while its peer can still signal it
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if (CG == 2) tmem_dealloc_2sm(tmem_base, NACC * BN > 256 ? 512 : NACC * BN); else tmem_dealloc(tmem_base, NACC * BN);
    }
}

template <int BN, int STAGES, int BK, bool DBG, int CG>
static int launch_gemm(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const GemmEpilogue& ep, int sms,
                       cudaStream_t stream, long long a_rows = 0, long long w_rows = 0) {
    CUtensorMap tmA, tmW;
    int rc = make_tmap_f16_2d(&tmA, a, static_cast<uint64_t>(K), static_cast<uint64_t>(a_rows > 0 ? a_rows : M),
                              static_cast<uint64_t>(lda) * 2, GEMM_KATOM, GEMM_BM);
    if (rc) return rc;
    rc = make_tmap_f16_2d(&tmW, w, static_cast<uint64_t>(K), static_cast<uint64_t>(w_rows > 0 ? w_rows : N),
                          static_cast<uint64_t>(ldw) * 2, GEMM_KATOM, BN > 256 ? BN / 4 : BN / CG);
    if (rc) return rc;
    auto kern = gemm_f16_tn_kernel<BN, STAGES, BK, DBG, CG>;
    constexpr int smem = GemmSmem<BN, STAGES, BK, CG>::TOTAL;
    GMM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int tiles = ((M + GEMM_BM * CG - 1) / (GEMM_BM * CG)) * (N / BN);
    int ctas = tiles * CG < sms ? tiles * CG : (sms / CG) * CG;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CG == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (gridmm_use_pdl()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return static_cast<int>(cudaLaunchKernelEx(&cfg, kern, tmA, tmW, M, N, K, ep));
}

}  // namespace gmm

// ----------------------------------------------------------------------------- C ABI
static long long* g_gemm_dbg = nullptr;
static int g_gemm_pairs = 1;
static int g_gemm_384 = 0;
// 256 x 384 pair tiles are OFF by default: measured on the navigation step (B = 32, the 57-query GEMMs N = 3072 / 2304) they were
// ~8 us per launch SLOWER than the two-round 256 x 256 schedule (1.197 vs 1.131 ms per step) -- a single accumulator serialises the
// GELU epilogue behind the main loop and the 4-stage ring is shallower.  Kept behind this hook (1 = on) with its parity tests.
extern "C" void gridmm_debug_set_gemm_384(int on) { g_gemm_384 = on; }
// Debug hook: 0 disables the CTA-pair (cta_group::2) path (A/B comparisons in tools/microbench.py).
extern "C" void gridmm_debug_set_gemm_pairs(int on) { g_gemm_pairs = on; }
// Debug hook (tools/microbench.py): per-CTA cycle counters [grid][8] written by the next GEMM launches; null disables.
extern "C" void gridmm_debug_set_gemm_counters(long long* dbg) { g_gemm_dbg = dbg; }

static int gemm_dispatch(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias,
                         const float* residual, int ld_res, float* out_f32, int ld_f32, void* out_f16, int ld_f16,
                         int act, void* out_lanes, int lanes_rows, cudaStream_t stream, const int* grp = nullptr,
                         const float* gw2 = nullptr, float* cls_part = nullptr, long long a_rows = 0, long long w_rows = 0,
                         float* cls_raw = nullptr, const int* m_dev = nullptr) {
    using namespace gmm;
    if (M <= 0) return 0;
    if (N % 128 != 0 || K % GEMM_KATOM != 0 || (lda % 8) || (ldw % 8)) return GRIDMM_ERR_SHAPE;
    if ((out_f32 && (ld_f32 % 4)) || (out_f16 && (ld_f16 % 8)) || (residual && (ld_res % 4))) return GRIDMM_ERR_SHAPE;
    if (!a || !w || (!out_f32 && !out_f16 && !out_lanes && !cls_part)) return GRIDMM_ERR_ARG;
    const int sms = gridmm_sm_count();
    if (sms <= 0) return GRIDMM_ERR_DRIVER;
    GemmEpilogue ep{bias, residual, out_f32, reinterpret_cast<__half*>(out_f16), ld_res, ld_f32, ld_f16, act,
                    reinterpret_cast<__half*>(out_lanes), lanes_rows, grp, gw2, cls_part, cls_raw, m_dev, g_gemm_dbg};
    if (grp) {
        // grouped heads: 128 x 128 tiles, one CTA each (row tiles address A / W / outputs through the table; the tensor maps
        // span all A rows and all stacked weight rows)
        if (K % 128 != 0 || (N % 128 != 0)) return GRIDMM_ERR_SHAPE;
        gridmm_count_launch(1);
        return launch_gemm<128, 3, 128, false, 1>(a, lda, w, ldw, M, N, K, ep, sms, stream, a_rows, w_rows);
    }
    // tile width: a 128x256 tile does twice the work of a 128x128 one in ~1.45x the time (shared-memory bandwidth: both
    // re-read their operands for every MMA, the wide tile reads 96 B/clk + fills 96 B/clk against 128 + 128 for the narrow
    // one), the narrow tile quantises better over the SMs; pick the cheaper schedule
    const int tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
    bool wide = false;
    if (N % 256 == 0) {
        const long long r256 = (static_cast<long long>(tiles_m) * (N / 256) + sms - 1) / sms;
        const long long r128 = (static_cast<long long>(tiles_m) * (N / 128) + sms - 1) / sms;
        wide = r256 * 145 <= r128 * 100;
    }
    // Large problems run on CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles): each SM then reads only half of the W tile
    // per MMA and fills half per stage, which is what lifts the shared-memory-bandwidth cap of the single-CTA tiles.
    // 128-wide single-CTA tiles take 128 K-columns per stage (8 MMAs per barrier round trip) when K allows.
    const bool pair = g_gemm_pairs && (N % 256 == 0) && (static_cast<long long>(tiles_m / 2) * (N / 256) * 2 >= sms);
    const bool deep = (K % 128 == 0);
    // (opt-in, see g_gemm_384) 256 x 384 pair tiles when they cover the whole problem in ONE round and that round would be cheaper
    // than the schedule picked above by the ingest model (cost ~ rounds x bytes a CTA ingests per k-block: 32 KB for 128 x 128 and
    // for 256 x 256 pair tiles, 48 KB for 128 x 256, 40 KB for 256 x 384 pair tiles)
    bool wide384 = false;
    if (g_gemm_pairs && g_gemm_384 && N % 384 == 0 && !lanes_rows && !cls_part) {
        const long long pairs384 = static_cast<long long>((tiles_m + 1) / 2) * (N / 384);
        if (pairs384 * 2 <= sms) {
            long long cost;
            if (pair) cost = ((static_cast<long long>((tiles_m + 1) / 2) * (N / 256) + sms / 2 - 1) / (sms / 2)) * 32;
            else if (wide) cost = ((static_cast<long long>(tiles_m) * (N / 256) + sms - 1) / sms) * 48;
            else cost = ((static_cast<long long>(tiles_m) * (N / 128) + sms - 1) / sms) * 32;
            wide384 = 40 < cost;
        }
    }
    int rc;
    if (g_gemm_dbg)
        rc = wide384 ? launch_gemm<384, 4, 64, true, 2>(a, lda, w, ldw, M, N, K, ep, sms, stream)
           : pair ? launch_gemm<256, 6, 64, true, 2>(a, lda, w, ldw, M, N, K, ep, sms, stream)
           : wide ? launch_gemm<256, 4, 64, true, 1>(a, lda, w, ldw, M, N, K, ep, sms, stream)
           : deep ? launch_gemm<128, 3, 128, true, 1>(a, lda, w, ldw, M, N, K, ep, sms, stream)
                  : launch_gemm<128, 6, 64, true, 1>(a, lda, w, ldw, M, N, K, ep, sms, stream);
    else
        rc = wide384 ? launch_gemm<384, 4, 64, false, 2>(a, lda, w, ldw, M, N, K, ep, sms, stream)
           : pair ? launch_gemm<256, 6, 64, false, 2>(a, lda, w, ldw, M, N, K, ep, sms, stream)
           : wide ? launch_gemm<256, 4, 64, false, 1>(a, lda, w, ldw, M, N, K, ep, sms, stream)
           : deep ? launch_gemm<128, 3, 128, false, 1>(a, lda, w, ldw, M, N, K, ep, sms, stream)
                  : launch_gemm<128, 6, 64, false, 1>(a, lda, w, ldw, M, N, K, ep, sms, stream);
    gridmm_count_launch(1);
    return rc;
}

extern "C" int gridmm_linear_f16(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias,
                                 const float* residual, int ld_res, float* out_f32, int ld_f32, void* out_f16, int ld_f16,
                                 int act, const int* m_dev, cudaStream_t stream) {
    return gemm_dispatch(a, lda, w, ldw, M, N, K, bias, residual, ld_res, out_f32, ld_f32, out_f16, ld_f16, act, nullptr, 0, stream,
                         nullptr, nullptr, nullptr, 0, 0, nullptr, m_dev);
}

// text_proj for gridmm_pool: out_lanes[b, N/8, 128] (16-byte units, lane-major) = a[M = batch*rows_per_b, K] . w[N, K]^T + bias
extern "C" int gridmm_linear_f16_lanes(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias,
                                       void* out_lanes, int rows_per_b, cudaStream_t stream) {
    if (!out_lanes || rows_per_b < 1 || rows_per_b > 256 || (M % rows_per_b)) return GRIDMM_ERR_SHAPE;
    return gemm_dispatch(a, lda, w, ldw, M, N, K, bias, nullptr, 0, nullptr, 0, nullptr, 0, 0, out_lanes, rows_per_b, stream);
}

// All ClsPrediction heads of a navigation step in ONE grouped launch (vilmodel.py:663-674, 859-878, 903-907):
//   r = ReLU(x . W_h^T + b_h) for every row of every head h, reduced on the fly to the three sums the head's tail needs
//   (LayerNorm + Linear(768 -> 1) = rstd * (sum r*gamma*w2 - mean * sum gamma*w2) + const): cls_part[row][N/64][3].
//   a         fp16 [a_rows, K] (K = 3 x 768: the [hi | lo | hi] split of the fp32 inputs, see gridmm_split_rows)
//   w         fp16 [groups * N, K] stacked [Wh | Wh | Wl] weights; bias, gw2 fp32 [groups * N]
//   grp       device int [tiles_m][4]: first A row, first W row (group * N), first output row of every 128-row tile, mode
//   cls_part  fp32 [out_rows, N / 64, 3]
//   cls_raw   fp32 [out_rows, N]: mode-1 tiles store x . W^T itself (bias-free, no activation) -- the two K halves of
//             sap_fuse_linear's first layer ([gmap'_0 ; vp_0] . [Wg | Wv]^T), which gridmm_nav_logits2 adds and finishes
extern "C" int gridmm_cls_heads_f16(const void* a, int lda, long long a_rows, const void* w, int ldw, int groups, int tiles_m,
                                    int N, int K, const float* bias, const float* gw2, const int* grp, float* cls_part,
                                    float* cls_raw, cudaStream_t stream) {
    if (tiles_m <= 0) return 0;
    if (!grp || !gw2 || !cls_part || !cls_raw || groups < 1) return GRIDMM_ERR_ARG;
    return gemm_dispatch(a, lda, w, ldw, tiles_m * 128, N, K, bias, nullptr, 0, nullptr, 0, nullptr, 0, 2, nullptr, 0, stream, grp, gw2,
                         cls_part, a_rows, static_cast<long long>(groups) * N, cls_raw);
}

// gridmm_linear_f16 over the first *m_dev rows only (m_dev: device int, <= M): for packed operands whose row count is computed on
// the GPU (the fusion encoder's context without its masked rows, gridmm_kv_index).  The launch is sized for M; CTAs without
// tiles exit at once.
extern "C" int gridmm_linear_f16_rows(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias,
                                      void* out_f16, int ld_f16, const int* m_dev, cudaStream_t stream) {
    if (!m_dev) return GRIDMM_ERR_ARG;
    return gemm_dispatch(a, lda, w, ldw, M, N, K, bias, nullptr, 0, nullptr, 0, out_f16, ld_f16, 0, nullptr, 0, stream, nullptr, nullptr,
                         nullptr, 0, 0, nullptr, m_dev);
}
