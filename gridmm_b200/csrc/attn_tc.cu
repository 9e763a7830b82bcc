// Multi-head attention core on tcgen05: O = softmax(Q K^T / sqrt(64) + key_mask) V, 12 heads x 64.
// Reference: BertSelfAttention / BertOutAttention (map_nav_src/models/vilmodel.py:95-153, 317-368, additive -10000 mask,
// models/ops.py:25-34) and nn.MultiheadAttention with key_padding_mask (-inf) in TransformerEncoderLayer.forward_pre
// (models/transformer.py:170-182).
//
// One CTA per (128-query tile, head, episode), 5 warps:
//   warp 4      TMA (Q tile, K and V of the (episode, head): 64-column boxes of the fused projection buffers, SWIZZLE_128B) and
//               the two tcgen05.mma phases, issued from a converged warp through elect_one()
//                 S[128, Sk]  = Q[128, 64] . K[Sk, 64]^T        (both operands K-major from shared memory, N = Sk in one or two MMAs)
//                 O[128, 64]  = P[128, Sk] . V[Sk, 64]          (A = P from TENSOR MEMORY, B = V as an MN-major shared-memory operand:
//                                                                 V stays [key][dim] exactly as the projection wrote it)
//   warps 0..3  softmax: thread = query row = TMEM lane.  Two passes over the row of S in TMEM (max, then exp2 and sum), the
//               probabilities go back into tensor memory as packed fp16 over the dead part of S; after the second MMA the
//               same threads scale O by 1/sum and store 128 contiguous bytes per row.
// Sk <= 320 (S + O must fit 512 TMEM columns); larger Sk falls back to the mma.sync kernel of attn.cu.
#include "common.cuh"
#include "host_util.h"

namespace gmm {

struct AttnTcParams {
    __half* o; int ldo;
    int q_rows, k_rows;                // rows per episode in the q / kv buffers (batch stride)
    const uint8_t* kmask;              // [B, Sk] 1 = valid key
    float mask_neg;                    // -10000 (BERT additive) or -inf (key_padding_mask)
    int sq, sk, nkey;                  // nkey = sk rounded up to 16
    int o_col, tmem_cols;              // TMEM column of O, allocation size (power of two)
    float scale;
    // ragged ("packed") sequences: rows of episode b start at q_off[b] / k_off[b] and number q_cnt[b] / k_cnt[b] (<= sq / sk, which
    // then only size the launch and the shared memory); kmask / kbias are indexed by packed key row when k_off is given
    const int* q_off; const int* q_cnt; const int* k_off; const int* k_cnt;
    const float* kbias;                // optional additive score bias per VALID key (log of a key's multiplicity), or null
};

// MN-major shared-memory operand (V as [key][dim]: 64 dims = 128 contiguous bytes per key, SWIZZLE_128B, 8-key groups 1024 B
// apart): canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- n = 1 here, SBO = 1024 B.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;             // LBO: one 64-element block along N only
    d |= static_cast<uint64_t>(1024 >> 4) << 32;     // SBO: 8 keys x 128 B
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint32_t umma_idesc_f16_rt(int m, int n, int b_mn_major) {
    return (1u << 4) | (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_alloc_rt(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(160) attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                      const __grid_constant__ CUtensorMap tmV, AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;                                  // 128 x 128 B
    uint8_t* sK = sQ + 128 * 128;                        // nkey x 128 B
    uint8_t* sV = sK + p.nkey * 128;                     // nkey x 128 B
    float* sM = reinterpret_cast<float*>(sV + p.nkey * 128);          // [nkey] additive mask
    uint64_t* bars = reinterpret_cast<uint64_t*>(sM + p.nkey);        // [0] loads, [1] S ready, [2] P ready, [3] O ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    pdl_wait();
    const int q_base = p.q_off ? __ldg(p.q_off + b) : b * p.q_rows;
    const int q_n = p.q_cnt ? __ldg(p.q_cnt + b) : p.sq;
    const int k_base = p.k_off ? __ldg(p.k_off + b) : b * p.k_rows;
    const int k_n = p.k_cnt ? __ldg(p.k_cnt + b) : p.sk;
    if (q0 >= q_n) return;                              // ragged batch: this query tile lies past the episode's rows (whole CTA)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 128); mbar_init(&bars[3], 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc_rt(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    const int nkey = max((k_n + 15) & ~15, 16);          // keys of THIS episode, rounded to the MMA's K step (<= p.nkey)
    const size_t m_base = p.k_off ? static_cast<size_t>(k_base) : static_cast<size_t>(b) * p.sk;
    for (int j = threadIdx.x; j < nkey; j += blockDim.x) {
        float m = -INFINITY;                               // keys past Sk never contribute
        if (j < k_n) {
            m = p.kmask[m_base + j] ? 0.0f : p.mask_neg;
            if (p.kbias && m == 0.0f) m = p.kbias[m_base + j];
        }
        sM[j] = m;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    if (warp == 4) {
        // ---------------------------------------------------------------- loads + both MMA phases
        if (elect_one()) {
            mbar_arrive_expect_tx(&bars[0], static_cast<uint32_t>(128 * 128 + 2 * nkey * 128));
            tma_load_2d(sQ, &tmQ, h * 64, q_base + q0, &bars[0]);
            for (int j = 0; j < nkey; j += 16) {
                tma_load_2d(sK + j * 128, &tmK, h * 64, k_base + j, &bars[0]);
                tma_load_2d(sV + j * 128, &tmV, h * 64, k_base + j, &bars[0]);
            }
        }
        __syncwarp();
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        if (elect_one()) {
            // S = Q K^T: N = nkey in one MMA (<= 256) or 256 + the rest
            const int n_a = nkey <= 256 ? nkey : 256, n_b = nkey - n_a;
            const uint32_t id_a = umma_idesc_f16_rt(128, n_a, 0);
            const uint64_t dq = umma_desc_sw128_kmajor(smem_u32(sQ));
            const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(sK));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, dq + 2 * k, dk + 2 * k, id_a, k ? 1u : 0u);
            if (n_b > 0) {
                const uint32_t id_b = umma_idesc_f16_rt(128, n_b, 0);
                const uint64_t dk2 = umma_desc_sw128_kmajor(smem_u32(sK) + 256 * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + 256, dq + 2 * k, dk2 + 2 * k, id_b, k ? 1u : 0u);
            }
            umma_commit(&bars[1]);
        }
        __syncwarp();
        mbar_wait(&bars[2], 0);                            // probabilities are in tensor memory
        tc_fence_after();
        if (elect_one()) {
            const uint32_t id_o = umma_idesc_f16_rt(128, 64, 1);
            const uint32_t v_addr = smem_u32(sV);
            for (int j = 0; j < nkey / 16; ++j)          // K = 16 keys per instruction: 8 TMEM columns of P, two 8-key groups of V
                umma_f16_ts(tmem_base + p.o_col, tmem_base + j * 8, umma_desc_sw128_mnmajor(v_addr + j * 2048), id_o, j ? 1u : 0u);
            umma_commit(&bars[3]);
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- softmax + epilogue: thread = query row
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        constexpr float LOG2E = 1.4426950408889634f;
        const float sc = p.scale * LOG2E;                  // scores are compared / exponentiated in the log2 domain
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        // a warp whose 32 rows all lie past the episode's queries (the ragged tail tile) skips both passes: its lanes of P stay
        // whatever they were, rows are independent and its output rows are never stored
        const int nk_w = (q0 + warp * 32 < q_n) ? nkey : 0;
        float mx = -INFINITY;
        for (int c = 0; c < nk_w; c += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_row + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float mj = sM[c + j];
                if (mj != -INFINITY) mx = fmaxf(mx, fmaf(__uint_as_float(v[j]), sc, mj * LOG2E));
            }
        }
        const float m_use = (mx == -INFINITY) ? 0.0f : mx;
        float l = 0.f;
        for (int c = 0; c < nk_w; c += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_row + c, v);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                // keys past the episode's end / -inf masked keys get probability exactly 0 whatever the (stale) score is
                const float m0 = sM[c + j], m1 = sM[c + j + 1];
                const float p0 = (m0 == -INFINITY) ? 0.0f : ex2f(fmaf(__uint_as_float(v[j]), sc, m0 * LOG2E) - m_use);
                const float p1 = (m1 == -INFINITY) ? 0.0f : ex2f(fmaf(__uint_as_float(v[j + 1]), sc, m1 * LOG2E) - m_use);
                const __half2 hp = __floats2half2_rn(p0, p1);
                const float2 fp = __half22float2(hp);      // the normaliser uses the rounded probabilities the MMA will see
                l += fp.x + fp.y;
                pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
            }
            tmem_st_32x32b_x8(t_row + (c >> 1), pk);     // 16 probabilities = 8 columns, over the part of S already consumed
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[2]);
        mbar_wait(&bars[3], 0);
        tc_fence_after();
        const float inv = l > 0.f ? 1.0f / l : 0.0f;
        const int row = q0 + warp * 32 + lane;
        uint4 ov[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_row + p.o_col + c * 16, v);
            tmem_ld_wait();
            uint32_t hh[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const __half2 h2 = __floats2half2_rn(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
                hh[j] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            ov[2 * c] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            ov[2 * c + 1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
        }
        if (row < q_n) {
            uint4* dst = reinterpret_cast<uint4*>(p.o + (static_cast<size_t>(q_base) + row) * p.ldo + h * 64);
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c] = ov[c];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}


// ---------------------------------------------------------------------------------------------------------------------------------
// Head-PAIR variant for short query sequences (<= 64 rows per (episode, head): the 57-query fusion-encoder shapes).  One CTA per
// (two heads, episode): the 128 TMEM lanes hold the queries of head h0 (lanes 0..63) and of head h0+1 (lanes 64..127), so every
// lane owns a real row instead of half of them idling on padding.
//   S1[128, Sk] = Qpair . K(h0)^T   (lanes 0..63 meaningful)      S2[128, Sk] = Qpair . K(h0+1)^T   (lanes 64..127 meaningful)
//   P (fp16) written by each lane over consumed S1 columns of ITS OWN lane (lanes 64..127 read S2, their S1 columns are dead)
//   O1[128, 64] = P . V(h0), O2[128, 64] = P . V(h0+1) over the dead S2 columns; lanes 0..63 store O1, lanes 64..127 store O2.
// S1 + S2 need up to 512 TMEM columns, i.e. one CTA per SM, so the row softmax (the bulk of this kernel) is spread over EIGHT warps:
// warps w and w + 4 share TMEM lane quadrant w (a warp may only touch lanes 32 (warp % 4) ..) and split the row's keys in two
// column ranges; row maxima and sums meet in shared memory.  The tensor pipe does twice the useful flops (it is ~10 % busy here).
// Sk <= 256.
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(288) attn_tc_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                           const __grid_constant__ CUtensorMap tmV, AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;                                  // 2 x 64 rows x 128 B
    uint8_t* sK0 = sQ + 128 * 128;                       // nkey x 128 B each
    uint8_t* sK1 = sK0 + p.nkey * 128;
    uint8_t* sV0 = sK1 + p.nkey * 128;
    uint8_t* sV1 = sV0 + p.nkey * 128;
    float* sM = reinterpret_cast<float*>(sV1 + p.nkey * 128);         // [nkey] additive mask, log2 domain
    float* sX = sM + p.nkey;                                          // [2][128] partial row maxima, then partial row sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 256);           // [0] loads, [1] S ready, [2] P ready, [3] O ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h0 = blockIdx.x * 2, b = blockIdx.y;
    pdl_wait();
    const int q_base = p.q_off ? __ldg(p.q_off + b) : b * p.q_rows;
    const int q_n = p.q_cnt ? __ldg(p.q_cnt + b) : p.sq;
    const int k_base = p.k_off ? __ldg(p.k_off + b) : b * p.k_rows;
    const int k_n = p.k_cnt ? __ldg(p.k_cnt + b) : p.sk;
    constexpr float LOG2E = 1.4426950408889634f;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 256); mbar_init(&bars[3], 1);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc_rt(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    const int nkey = max((k_n + 15) & ~15, 16);
    const size_t m_base = p.k_off ? static_cast<size_t>(k_base) : static_cast<size_t>(b) * p.sk;
    for (int j = threadIdx.x; j < nkey; j += blockDim.x) {
        float m = -INFINITY;
        if (j < k_n) {
            m = (!p.kmask || p.kmask[m_base + j]) ? 0.0f : p.mask_neg;
            if (p.kbias && m == 0.0f) m = p.kbias[m_base + j];
        }
        sM[j] = m * LOG2E;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    const uint32_t oc = static_cast<uint32_t>(p.o_col);      // column of S2, later of O1 | O2
    const int n_chunks = nkey >> 4;                          // 16 keys per chunk
    const int c_split = (n_chunks + 1) >> 1;                 // part 0: chunks [0, c_split), part 1: [c_split, n_chunks)

    if (warp == 8) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&bars[0], static_cast<uint32_t>(128 * 128 + 4 * nkey * 128));
            tma_load_2d(sQ, &tmQ, h0 * 64, q_base, &bars[0]);
            tma_load_2d(sQ + 64 * 128, &tmQ, (h0 + 1) * 64, q_base, &bars[0]);
            for (int j = 0; j < nkey; j += 16) {
                tma_load_2d(sK0 + j * 128, &tmK, h0 * 64, k_base + j, &bars[0]);
                tma_load_2d(sK1 + j * 128, &tmK, (h0 + 1) * 64, k_base + j, &bars[0]);
            }
            for (int j = 0; j < nkey; j += 16) {
                tma_load_2d(sV0 + j * 128, &tmV, h0 * 64, k_base + j, &bars[0]);
                tma_load_2d(sV1 + j * 128, &tmV, (h0 + 1) * 64, k_base + j, &bars[0]);
            }
        }
        __syncwarp();
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t id_s = umma_idesc_f16_rt(128, nkey, 0);
            const uint64_t dq = umma_desc_sw128_kmajor(smem_u32(sQ));
            const uint64_t dk0 = umma_desc_sw128_kmajor(smem_u32(sK0));
            const uint64_t dk1 = umma_desc_sw128_kmajor(smem_u32(sK1));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, dq + 2 * k, dk0 + 2 * k, id_s, k ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + oc, dq + 2 * k, dk1 + 2 * k, id_s, k ? 1u : 0u);
            umma_commit(&bars[1]);
        }
        __syncwarp();
        mbar_wait(&bars[2], 0);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t id_o = umma_idesc_f16_rt(128, 64, 1);
            const uint32_t v0 = smem_u32(sV0), v1 = smem_u32(sV1);
            // chunk j of P lives at column 8 j (part 0) or 16 c_split + 8 (j - c_split) (part 1: over that part's own S columns)
            for (int j = 0; j < n_chunks; ++j) {
                const uint32_t pc = j < c_split ? j * 8 : c_split * 16 + (j - c_split) * 8;
                umma_f16_ts(tmem_base + oc, tmem_base + pc, umma_desc_sw128_mnmajor(v0 + j * 2048), id_o, j ? 1u : 0u);
            }
            for (int j = 0; j < n_chunks; ++j) {
                const uint32_t pc = j < c_split ? j * 8 : c_split * 16 + (j - c_split) * 8;
                umma_f16_ts(tmem_base + oc + 64, tmem_base + pc, umma_desc_sw128_mnmajor(v1 + j * 2048), id_o, j ? 1u : 0u);
            }
            umma_commit(&bars[3]);
        }
        __syncwarp();
    } else {
        const int quad = warp & 3, part = warp >> 2;        // TMEM lane quadrant, column half
        const int half = quad >> 1;                         // 0: head h0 (lanes 0..63), 1: head h0 + 1 (lanes 64..127)
        const int row128 = quad * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const uint32_t t_s = t_lane + (half ? oc : 0u);
        const float sc = p.scale * LOG2E;
        const int c_lo = part ? c_split : 0, c_hi = part ? n_chunks : c_split;
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        float mx = -INFINITY;
        for (int c = c_lo; c < c_hi; c += 2) {
            uint32_t v[2][16];
            const bool two = c + 1 < c_hi;
            tmem_ld_32x32b_x16(t_s + c * 16, v[0]);
            if (two) tmem_ld_32x32b_x16(t_s + c * 16 + 16, v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
#pragma unroll
                for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fmaf(__uint_as_float(v[u][j]), sc, sM[(c + u) * 16 + j]));
            }
        }
        sX[part * 128 + row128] = mx;
        bar_sync_named(1 + quad, 64);
        mx = fmaxf(mx, sX[(part ^ 1) * 128 + row128]);
        const float m_use = (mx == -INFINITY) ? 0.0f : mx;
        bar_sync_named(1 + quad, 64);                      // both partial maxima are read before the slots are reused for the sums
        float l = 0.f;
        for (int c = c_lo; c < c_hi; c += 2) {
            uint32_t v[2][16];
            const bool two = c + 1 < c_hi;
            tmem_ld_32x32b_x16(t_s + c * 16, v[0]);
            if (two) tmem_ld_32x32b_x16(t_s + c * 16 + 16, v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    // -inf masks (keys past the episode's end, key_padding_mask) give exactly 0 through ex2(-inf); the guard keeps a
                    // stale non-finite score out
                    const float m0 = sM[(c + u) * 16 + j], m1 = sM[(c + u) * 16 + j + 1];
                    const float p0 = (m0 == -INFINITY) ? 0.0f : ex2f(fmaf(__uint_as_float(v[u][j]), sc, m0) - m_use);
                    const float p1 = (m1 == -INFINITY) ? 0.0f : ex2f(fmaf(__uint_as_float(v[u][j + 1]), sc, m1) - m_use);
                    const __half2 hp = __floats2half2_rn(p0, p1);
                    const float2 fp = __half22float2(hp);
                    l += fp.x + fp.y;
                    pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
                }
                const int cc = c + u;
                tmem_st_32x32b_x8(t_lane + (part ? c_split * 16 + (cc - c_split) * 8 : cc * 8), pk);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[2]);
        sX[part * 128 + row128] = l;
        bar_sync_named(1 + quad, 64);
        l += sX[(part ^ 1) * 128 + row128];
        mbar_wait(&bars[3], 0);
        tc_fence_after();
        const float inv = l > 0.f ? 1.0f / l : 0.0f;
        const int row = (quad & 1) * 32 + lane;
        // each of the row's two threads stores 32 of the 64 output columns
        uint4 ov[4];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_lane + oc + half * 64 + part * 32 + c * 16, v);
            tmem_ld_wait();
            uint32_t hh[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const __half2 h2 = __floats2half2_rn(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
                hh[j] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            ov[2 * c] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            ov[2 * c + 1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
        }
        if (row < q_n) {
            uint4* dst = reinterpret_cast<uint4*>(p.o + (static_cast<size_t>(q_base) + row) * p.ldo + (h0 + half) * 64 + part * 32);
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[c] = ov[c];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    }
    pdl_launch_dependents();
}

}  // namespace gmm

// returns GRIDMM_ERR_SHAPE when the shape is outside this kernel's range (the caller then uses the mma.sync kernel)
int gridmm_attention_tc(const void* q, int ldq, int q_rows, const void* k, int ldk, const void* v, int ldv, int k_rows, void* o,
                        int ldo, const unsigned char* kmask, float mask_neg, int batch, int heads, int sq, int sk, float scale,
                        cudaStream_t stream, const int* q_off, const int* q_cnt, const int* k_off, const int* k_cnt,
                        const float* kbias, long long q_total, long long k_total) {
    using namespace gmm;
    const int nkey = (sk + 15) & ~15;
    if (nkey > 320 || (ldo % 8) || (reinterpret_cast<uintptr_t>(o) & 15)) return GRIDMM_ERR_SHAPE;
    AttnTcParams p;
    p.o = reinterpret_cast<__half*>(o); p.ldo = ldo; p.q_rows = q_rows; p.k_rows = k_rows; p.kmask = kmask; p.mask_neg = mask_neg;
    p.sq = sq; p.sk = sk; p.nkey = nkey; p.scale = scale;
    p.q_off = q_off; p.q_cnt = q_cnt; p.k_off = k_off; p.k_cnt = k_cnt; p.kbias = kbias;
    p.o_col = ((nkey / 2) + 31) & ~31;
    const int need = nkey > p.o_col + 64 ? nkey : p.o_col + 64;
    p.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    const uint64_t q_outer = q_total > 0 ? static_cast<uint64_t>(q_total) : static_cast<uint64_t>(batch) * q_rows;
    const uint64_t k_outer = k_total > 0 ? static_cast<uint64_t>(k_total) : static_cast<uint64_t>(batch) * k_rows;
    CUtensorMap tmQ, tmK, tmV;
    int rc = make_tmap_f16_2d(&tmQ, q, static_cast<uint64_t>(heads) * 64, q_outer, static_cast<uint64_t>(ldq) * 2, 64, 128);
    if (rc) return rc;
    rc = make_tmap_f16_2d(&tmK, k, static_cast<uint64_t>(heads) * 64, k_outer, static_cast<uint64_t>(ldk) * 2, 64, 16);
    if (rc) return rc;
    rc = make_tmap_f16_2d(&tmV, v, static_cast<uint64_t>(heads) * 64, k_outer, static_cast<uint64_t>(ldv) * 2, 64, 16);
    if (rc) return rc;
    const int smem = 1024 + 128 * 128 + 2 * nkey * 128 + nkey * 4 + 64;
    GMM_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid((sq + 127) / 128, heads, batch);
    GMM_CUDA_CHECK(launch_pdl(attn_tc_kernel, grid, dim3(160), smem, stream, tmQ, tmK, tmV, p));
    return 0;
}

// Head-pair kernel for <= 64 queries per (episode, head) and <= 256 keys; kmask may be NULL (every key of the k_cnt[b] rows valid).
// Returns GRIDMM_ERR_SHAPE outside its range (the caller falls back to the mma.sync kernel).
int gridmm_attention_tc_pair(const void* q, int ldq, int q_rows, const void* k, int ldk, const void* v, int ldv, int k_rows, void* o,
                             int ldo, const unsigned char* kmask, float mask_neg, int batch, int heads, int sq, int sk, float scale,
                             cudaStream_t stream, const int* q_off, const int* q_cnt, const int* k_off, const int* k_cnt,
                             const float* kbias, long long q_total, long long k_total) {
    using namespace gmm;
    const int nkey = max((sk + 15) & ~15, 16);
    if (sq > 64 || nkey > 256 || (heads & 1) || (ldo % 8) || (reinterpret_cast<uintptr_t>(o) & 15)) return GRIDMM_ERR_SHAPE;
    AttnTcParams p;
    p.o = reinterpret_cast<__half*>(o); p.ldo = ldo; p.q_rows = q_rows; p.k_rows = k_rows; p.kmask = kmask; p.mask_neg = mask_neg;
    p.sq = sq; p.sk = sk; p.nkey = nkey; p.scale = scale;
    p.q_off = q_off; p.q_cnt = q_cnt; p.k_off = k_off; p.k_cnt = k_cnt; p.kbias = kbias;
    p.o_col = (nkey + 31) & ~31;
    const int need = p.o_col + (nkey > 128 ? nkey : 128);
    p.tmem_cols = need <= 256 ? 256 : 512;
    const uint64_t q_outer = q_total > 0 ? static_cast<uint64_t>(q_total) : static_cast<uint64_t>(batch) * q_rows;
    const uint64_t k_outer = k_total > 0 ? static_cast<uint64_t>(k_total) : static_cast<uint64_t>(batch) * k_rows;
    CUtensorMap tmQ, tmK, tmV;
    int rc = make_tmap_f16_2d(&tmQ, q, static_cast<uint64_t>(heads) * 64, q_outer, static_cast<uint64_t>(ldq) * 2, 64, 64);
    if (rc) return rc;
    rc = make_tmap_f16_2d(&tmK, k, static_cast<uint64_t>(heads) * 64, k_outer, static_cast<uint64_t>(ldk) * 2, 64, 16);
    if (rc) return rc;
    rc = make_tmap_f16_2d(&tmV, v, static_cast<uint64_t>(heads) * 64, k_outer, static_cast<uint64_t>(ldv) * 2, 64, 16);
    if (rc) return rc;
    const int smem = 1024 + 128 * 128 + 4 * nkey * 128 + nkey * 4 + 256 * 4 + 64;
    GMM_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid(heads / 2, batch);
    GMM_CUDA_CHECK(launch_pdl(attn_tc_pair_kernel, grid, dim3(288), smem, stream, tmQ, tmK, tmV, p));
    return 0;
}
