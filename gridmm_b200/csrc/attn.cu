// Multi-head attention core for the small sequences of the navigation step:
//   O = softmax(Q K^T / sqrt(64) + key_mask) V,   12 heads x 64, Sq <= ~300, Sk <= ~700.
// Reference: BertSelfAttention / BertOutAttention (map_nav_src/models/vilmodel.py:95-153, 317-368, additive
// -10000 mask from extend_neg_masks, models/ops.py:25-34) and nn.MultiheadAttention with key_padding_mask
// (-inf) inside TransformerEncoderLayer.forward_pre (models/transformer.py:170-182).
//
// One CTA per (query tile of 64 or 128 rows, head, episode), one warp per 16 query rows: K, V of that (episode, head)
// stream through a ring of three 64-key shared-memory tiles (cp.async groups; 64 KB per CTA, so three CTAs share an SM and
// the 384 CTAs of a 57-query launch run in one wave) and the flash-style online softmax (mma.sync.m16n8k16, fp16 in,
// fp32 accumulate) works on one tile while the next ones are in flight.  The projections around this core -- where the FLOPs are --
// run on tcgen05 (gemm_tc.cu); this core is softmax/latency bound at these sizes.
#include "common.cuh"
#include "host_util.h"

namespace gmm {

constexpr int ATT_DH = 64;
constexpr int ATT_LD = 72;          // padded smem row (halves): 144 B, 16-byte aligned, conflict-free ldmatrix
constexpr int ATT_KT = 64;           // keys per pipeline step (one cp.async group)
constexpr int ATT_NSTG = 3;          // key tiles resident at a time (ring): 64 KB of shared memory per CTA -> 3 CTAs per SM

struct AttnParams {
    const __half* q; const __half* k; const __half* v; __half* o;
    int ldq, ldk, ldv, ldo;            // row pitches in halves
    int q_rows, k_rows;                // rows per episode in the q / kv buffers (batch stride)
    const uint8_t* kmask;              // [B, Sk] 1 = valid key
    float mask_neg;                    // -10000 (BERT additive) or -inf (key_padding_mask)
    int sq, sk;
    float scale;
    const int* k_off;                  // optional packed context: keys of episode b are rows k_off[b] .. k_off[b] + k_cnt[b] - 1,
    const int* k_cnt;                  // all valid (kmask unused); sk is then the maximum over the batch (shared-memory sizing)
    const float* k_bias;               // optional additive score bias per packed key row (log of a key's multiplicity), or null
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// wait until at most `n` of this thread's cp.async groups are still pending (n is clamped to 7: older groups first)
__device__ __forceinline__ void cp_async_wait_upto(int n) {
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        case 6: cp_async_wait<6>(); break;
        default: cp_async_wait<7>(); break;
    }
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) attn_kernel(AttnParams p) {
    constexpr int ATT_QT = NW * 16;
    constexpr int ATT_THREADS = NW * 32;
    extern __shared__ __align__(16) uint8_t att_smem[];
    const int b = blockIdx.z;
    __half* sK = reinterpret_cast<__half*>(att_smem);
    __half* sV = sK + static_cast<size_t>(ATT_NSTG * ATT_KT) * ATT_LD;
    __half* sQ = sV + static_cast<size_t>(ATT_NSTG * ATT_KT) * ATT_LD;
    float* sM = reinterpret_cast<float*>(sQ + ATT_QT * ATT_LD);

    pdl_wait();
    const int sk = p.k_cnt ? p.k_cnt[b] : p.sk;               // keys of this episode
    const int sk_pad = (sk + 63) & ~63;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * ATT_QT, h = blockIdx.y;
    const __half* gq = p.q + (static_cast<size_t>(b) * p.q_rows) * p.ldq + h * ATT_DH;
    const size_t krow0 = p.k_off ? static_cast<size_t>(p.k_off[b]) : static_cast<size_t>(b) * p.k_rows;
    const __half* gk = p.k + krow0 * p.ldk + h * ATT_DH;
    const __half* gv = p.v + krow0 * p.ldv + h * ATT_DH;

    // stage the Q tile with the first key tile, then K, V in 64-key cp.async groups; rows past the end are zero-filled
    for (int i = tid; i < ATT_QT * 8; i += ATT_THREADS) {
        const int r = i >> 3, u = i & 7;
        if (q0 + r < p.sq) cp_async_16(smem_u32(sQ + r * ATT_LD + u * 8), gq + static_cast<size_t>(q0 + r) * p.ldq + u * 8);
        else *reinterpret_cast<uint4*>(sQ + r * ATT_LD + u * 8) = make_uint4(0, 0, 0, 0);
    }
    const int n_kt = sk_pad / ATT_KT;
    // key tile kt lives in ring slot kt % ATT_NSTG; one cp.async group per tile
    auto issue_tile = [&](int kt) {
        const int slot = kt % ATT_NSTG;
        for (int i = tid; i < ATT_KT * 8; i += ATT_THREADS) {
            const int r = kt * ATT_KT + (i >> 3), u = i & 7;
            __half* dk = sK + (slot * ATT_KT + (i >> 3)) * ATT_LD + u * 8;
            __half* dv = sV + (slot * ATT_KT + (i >> 3)) * ATT_LD + u * 8;
            if (r < sk) {
                cp_async_16(smem_u32(dk), gk + static_cast<size_t>(r) * p.ldk + u * 8);
                cp_async_16(smem_u32(dv), gv + static_cast<size_t>(r) * p.ldv + u * 8);
            } else {
                *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
            }
        }
        cp_async_commit();
    };
    for (int kt = 0; kt < n_kt && kt < ATT_NSTG; ++kt) issue_tile(kt);
    for (int j = tid; j < sk_pad; j += ATT_THREADS) {
        float m = -INFINITY;                                   // keys past Sk never contribute
        if (j < sk) m = p.k_off ? (p.k_bias ? p.k_bias[krow0 + j] : 0.0f) : (p.kmask[static_cast<size_t>(b) * p.sk + j] ? 0.0f : p.mask_neg);
        sM[j] = m;
    }
    // first group (Q + key tile 0) must have landed before the Q fragments are read
    cp_async_wait_upto(min(ATT_NSTG, n_kt) - 1);
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    // Q fragments: 16 rows x 64 dims = 4 k-steps
    uint32_t qf[4][4];
    {
        const int m = lane >> 3, rr = lane & 7;
        const int row = warp * 16 + (m & 1) * 8 + rr;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            ldsm_x4(smem_u32(sQ + row * ATT_LD + ks * 16 + (m >> 1) * 8), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.0f, 0.0f};
    constexpr float LOG2E = 1.4426950408889634f;

    for (int kt = 0; kt < sk_pad; kt += 64) {
        const int ti = kt / 64;
        const int soff = (ti % ATT_NSTG) * ATT_KT - kt;      // ring-slot row of key `kt + j` is soff + kt + j
        if (kt > 0) {      // the groups issued after tile ti's are those of tiles ti+1 .. min(ti + NSTG - 1, n_kt - 1)
            cp_async_wait_upto(min(ti + ATT_NSTG - 1, n_kt - 1) - ti);
            __syncthreads();
        }
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
            const int key = soff + kt + nt * 8 + (lane & 7);
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {   // two k-steps per ldmatrix.x4
                uint32_t b0, b1, b2, b3;
                ldsm_x4(smem_u32(sK + key * ATT_LD + kp * 32 + (lane >> 3) * 8), b0, b1, b2, b3);
                mma_16816(s[nt], qf[2 * kp], b0, b1);
                mma_16816(s[nt], qf[2 * kp + 1], b2, b3);
            }
        }
        // scale + mask, row maxima (rows g and g+8)
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float ma = sM[kt + nt * 8 + 2 * t], mb = sM[kt + nt * 8 + 2 * t + 1];
            s[nt][0] = s[nt][0] * p.scale + ma; s[nt][1] = s[nt][1] * p.scale + mb;
            s[nt][2] = s[nt][2] * p.scale + ma; s[nt][3] = s[nt][3] * p.scale + mb;
            mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
        float corr[2], m_use[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            m_use[r] = (m_new == -INFINITY) ? 0.0f : m_new;
            corr[r] = exp2f((m_run[r] - m_use[r]) * LOG2E);      // m_run = -inf -> 0
            m_run[r] = m_new;
            l_run[r] *= corr[r];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = exp2f((s[nt][0] - m_use[0]) * LOG2E), p1 = exp2f((s[nt][1] - m_use[0]) * LOG2E);
            const float p2 = exp2f((s[nt][2] - m_use[1]) * LOG2E), p3 = exp2f((s[nt][3] - m_use[1]) * LOG2E);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            // C fragments of n-tiles 2j, 2j+1 form the A fragment of key step j
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_h2(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_h2(p2, p3);
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int m = lane >> 3, rr = lane & 7;
            const int key = soff + kt + ks * 16 + (m & 1) * 8 + rr;
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {   // two 8-wide dim tiles per ldmatrix.x4.trans
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(smem_u32(sV + key * ATT_LD + (dp * 2 + (m >> 1)) * 8), b0, b1, b2, b3);
                mma_16816(o[dp * 2], pf[ks], b0, b1);
                mma_16816(o[dp * 2 + 1], pf[ks], b2, b3);
            }
        }
        if (ti + ATT_NSTG < n_kt) {      // every warp is done with this ring slot: refill it with tile ti + NSTG
            __syncthreads();
            issue_tile(ti + ATT_NSTG);
        }
    }
    // finalize: quad-reduce the row sums, normalise, store fp16
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = l_run[0] > 0.0f ? 1.0f / l_run[0] : 0.0f;
    const float inv1 = l_run[1] > 0.0f ? 1.0f / l_run[1] : 0.0f;
    const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
    __half* go = p.o + (static_cast<size_t>(b) * p.q_rows) * p.ldo + h * ATT_DH;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (row0 < p.sq)
            *reinterpret_cast<uint32_t*>(go + static_cast<size_t>(row0) * p.ldo + i * 8 + 2 * t) = pack_h2(o[i][0] * inv0, o[i][1] * inv0);
        if (row1 < p.sq)
            *reinterpret_cast<uint32_t*>(go + static_cast<size_t>(row1) * p.ldo + i * 8 + 2 * t) = pack_h2(o[i][2] * inv1, o[i][3] * inv1);
    }
    pdl_launch_dependents();      // at the very end: only the next launch's latency and prologue overlap this kernel's tail
}

}  // namespace gmm

int gridmm_attention_tc(const void* q, int ldq, int q_rows, const void* k, int ldk, const void* v, int ldv, int k_rows, void* o,
                        int ldo, const unsigned char* kmask, float mask_neg, int batch, int heads, int sq, int sk, float scale,
                        cudaStream_t stream, const int* q_off, const int* q_cnt, const int* k_off, const int* k_cnt,
                        const float* kbias, long long q_total, long long k_total);      // attn_tc.cu
int gridmm_attention_tc_pair(const void* q, int ldq, int q_rows, const void* k, int ldk, const void* v, int ldv, int k_rows, void* o,
                             int ldo, const unsigned char* kmask, float mask_neg, int batch, int heads, int sq, int sk, float scale,
                             cudaStream_t stream, const int* q_off, const int* q_cnt, const int* k_off, const int* k_cnt,
                             const float* kbias, long long q_total, long long k_total);      // attn_tc.cu
static int g_attn_legacy = 0;
// Debug hook: 1 forces the mma.sync kernel, 2 the tcgen05 kernel (A/B timing and parity of the two paths), 3 = by shape but with the
// tcgen05 head-pair kernel for <= 64 queries; 0 = by shape with the measured-fastest kernel per shape.
extern "C" void gridmm_debug_set_attn_legacy(int on) { g_attn_legacy = on; }

extern "C" int gridmm_attention_f16(const void* q, int ldq, int q_rows, const void* k, int ldk, const void* v, int ldv,
                                    int k_rows, void* o, int ldo, const unsigned char* kmask, float mask_neg, int batch,
                                    int heads, int sq, int sk, float scale, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0 || sq <= 0) return 0;
    if (!q || !k || !v || !o || !kmask) return GRIDMM_ERR_ARG;
    if (sk <= 0 || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 2)) return GRIDMM_ERR_SHAPE;
    // Query tiles are 128 rows on the tcgen05 path: with <= 64 queries per (episode, head) half of every MMA and of the softmax
    // threads is padding and the mma.sync kernel (64-row tiles, 4 warps) is faster -- measured at B=32: Sq=57/Sk=296 18.4 vs
    // 22.5 us, Sq=57/Sk=57 5.3 vs 6.8 us; Sq=216/Sk=216 35.8 vs 24.6 us, Sq=216/Sk=80 21.6 vs 13.6 us.  g_attn_legacy: 1 forces
    // mma.sync, 2 forces tcgen05 (tests).
    if (g_attn_legacy == 3 && sq <= 64 && heads * 64 <= ldq && heads * 64 <= ldk && heads * 64 <= ldv) {
        // <= 64 queries: two heads per CTA on the tcgen05 head-pair kernel (Sk <= 256).  Opt-in: measured at B=32 against the
        // mma.sync kernel below, Sq=57/Sk=208 14.8 vs 12.0 us, Sq=57/Sk=57 7.0 vs 5.5 us (profiles/r2_microbench.txt): both are at
        // the latency floor of one (load -> S -> softmax -> O -> store) chain per CTA and the tcgen05 chain has more hand-offs
        const int rc = gridmm_attention_tc_pair(q, ldq, q_rows, k, ldk, v, ldv, k_rows, o, ldo, kmask, mask_neg, batch, heads, sq, sk, scale,
                                                stream, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0);
        if (rc != GRIDMM_ERR_SHAPE) {
            if (rc == 0) gridmm_count_launch(1);
            return rc;
        }
    }
    if (g_attn_legacy != 1 && (sq > 64 || g_attn_legacy == 2) && heads * 64 <= ldq && heads * 64 <= ldk && heads * 64 <= ldv) {
        // tcgen05 path (attn_tc.cu); shapes it does not cover (Sk > 320, unaligned output) fall through to mma.sync
        const int rc = gridmm_attention_tc(q, ldq, q_rows, k, ldk, v, ldv, k_rows, o, ldo, kmask, mask_neg, batch, heads, sq, sk, scale, stream,
                                           nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0);
        if (rc != GRIDMM_ERR_SHAPE) {
            if (rc == 0) gridmm_count_launch(1);
            return rc;
        }
    }
    AttnParams p;
    p.q = reinterpret_cast<const __half*>(q); p.k = reinterpret_cast<const __half*>(k);
    p.v = reinterpret_cast<const __half*>(v); p.o = reinterpret_cast<__half*>(o);
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.q_rows = q_rows; p.k_rows = k_rows;
    p.kmask = kmask; p.mask_neg = mask_neg; p.sq = sq; p.sk = sk; p.scale = scale; p.k_off = nullptr; p.k_cnt = nullptr;
    p.k_bias = nullptr;
    const int sk_pad = (sk + 63) & ~63;
    // 8 warps (128 query rows per CTA) halve the K/V re-reads of long query sequences; 4 warps otherwise
    const int nw = (sq > 64) ? 8 : 4;
    const int qt = nw * 16;
    const int smem = (2 * ATT_NSTG * ATT_KT + qt) * ATT_LD * 2 + sk_pad * 4;
    if (smem > 227 * 1024) return GRIDMM_ERR_SHAPE;
    dim3 grid((sq + qt - 1) / qt, heads, batch);
    if (nw == 8) {
        GMM_CUDA_CHECK(cudaFuncSetAttribute(attn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GMM_CUDA_CHECK(launch_pdl(attn_kernel<8>, grid, dim3(256), smem, stream, p));
    } else {
        GMM_CUDA_CHECK(cudaFuncSetAttribute(attn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GMM_CUDA_CHECK(launch_pdl(attn_kernel<4>, grid, dim3(128), smem, stream, p));
    }
    gridmm_count_launch(1);
    return 0;
}

// Attention over a PACKED context (gridmm_kv_index): the keys / values of episode b are rows k_off[b] .. k_off[b] + k_cnt[b] - 1 of
// k / v and all of them are valid, so no key mask is applied (a masked key contributes exp(-10000) = 0 in fp32: dropping it is
// exact).  max_sk >= max_b k_cnt[b] sizes the shared memory.  Always the mma.sync kernel (query tiles of 64 / 128 rows).
extern "C" int gridmm_attention_varlen_f16(const void* q, int ldq, int q_rows, const void* k, int ldk, const void* v, int ldv,
                                           const int* k_off, const int* k_cnt, int max_sk, long long k_total, const float* k_bias,
                                           void* o, int ldo, int batch, int heads, int sq, float scale, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0 || sq <= 0) return 0;
    if (!q || !k || !v || !o || !k_off || !k_cnt) return GRIDMM_ERR_ARG;
    if (max_sk <= 0 || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 2)) return GRIDMM_ERR_SHAPE;
    if (g_attn_legacy == 3 && sq <= 64 && k_total > 0 && heads * 64 <= ldq && heads * 64 <= ldk && heads * 64 <= ldv) {
        const int rc = gridmm_attention_tc_pair(q, ldq, q_rows, k, ldk, v, ldv, 0, o, ldo, nullptr, 0.0f, batch, heads, sq, max_sk, scale, stream,
                                                nullptr, nullptr, k_off, k_cnt, k_bias, 0, k_total);
        if (rc != GRIDMM_ERR_SHAPE) {
            if (rc == 0) gridmm_count_launch(1);
            return rc;
        }
    }
    AttnParams p;
    p.q = reinterpret_cast<const __half*>(q); p.k = reinterpret_cast<const __half*>(k);
    p.v = reinterpret_cast<const __half*>(v); p.o = reinterpret_cast<__half*>(o);
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.q_rows = q_rows; p.k_rows = 0;
    p.kmask = nullptr; p.mask_neg = 0.0f; p.sq = sq; p.sk = max_sk; p.scale = scale; p.k_off = k_off; p.k_cnt = k_cnt;
    p.k_bias = k_bias;
    const int sk_pad = (max_sk + 63) & ~63;
    const int nw = (sq > 64) ? 8 : 4;
    const int qt = nw * 16;
    const int smem = (2 * ATT_NSTG * ATT_KT + qt) * ATT_LD * 2 + sk_pad * 4;
    dim3 grid((sq + qt - 1) / qt, heads, batch);
    if (nw == 8) {
        GMM_CUDA_CHECK(cudaFuncSetAttribute(attn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GMM_CUDA_CHECK(launch_pdl(attn_kernel<8>, grid, dim3(256), smem, stream, p));
    } else {
        GMM_CUDA_CHECK(cudaFuncSetAttribute(attn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GMM_CUDA_CHECK(launch_pdl(attn_kernel<4>, grid, dim3(128), smem, stream, p));
    }
    gridmm_count_launch(1);
    return 0;
}

// Attention over RAGGED (packed) query sequences on the tcgen05 kernel: the queries of episode b are rows q_off[b] .. q_off[b] +
// q_cnt[b] - 1 of q (and of o); the keys are either ragged too (k_off / k_cnt given: rows of k / v, with kmask / kbias indexed by
// packed key row) or regular (k_off = NULL: rows b * k_rows .. + sk - 1, kmask [batch, sk]).  max_sq / max_sk bound the per-episode
// counts (launch and shared-memory sizing), q_total / k_total are the row counts of the buffers (tensor-map extents).  Valid keys get
// kbias added to their score when kbias != NULL (log-multiplicity of de-duplicated keys), masked keys mask_neg.
extern "C" int gridmm_attention_ragged_f16(const void* q, int ldq, const int* q_off, const int* q_cnt, int max_sq, long long q_total,
                                           const void* k, int ldk, const void* v, int ldv, const int* k_off, const int* k_cnt,
                                           int k_rows, int max_sk, long long k_total, const unsigned char* kmask, const float* kbias,
                                           float mask_neg, void* o, int ldo, int batch, int heads, float scale, cudaStream_t stream) {
    if (batch <= 0 || max_sq <= 0) return 0;
    if (!q || !k || !v || !o || !kmask || !q_off || !q_cnt || ((k_off == nullptr) != (k_cnt == nullptr))) return GRIDMM_ERR_ARG;
    if (max_sk <= 0 || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8) || heads * 64 > ldq || heads * 64 > ldk || heads * 64 > ldv)
        return GRIDMM_ERR_SHAPE;
    const int rc = gridmm_attention_tc(q, ldq, 0, k, ldk, v, ldv, k_rows, o, ldo, kmask, mask_neg, batch, heads, max_sq, max_sk, scale,
                                       stream, q_off, q_cnt, k_off, k_cnt, kbias, q_total, k_total);
    if (rc == 0) gridmm_count_launch(1);
    return rc;
}
