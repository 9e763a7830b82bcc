// Helpers of the training path (BASELINE config 5): the dgrad / wgrad GEMMs of every nn.Linear run on the same tcgen05 kernel as
// the forward (gridmm_linear_f16 computes A[M,K] . W[N,K]^T with both operands K-contiguous), which needs the operands of
//   dx[M,K] = dy[M,N] . W[N,K]            ->  A = dy,   "W" = W^T  [K,N]
//   dW[N,K] = dy^T[N,M] . x[M,K]          ->  A = dy^T, "W" = x^T  [K,M]   (the contraction runs over the M rows, padded to 64)
// in transposed fp16 form.  These kernels produce them: fp32 / fp16 source -> fp16 copy and / or fp16 transpose (zero padded), and
// the bias gradient (column sums of dy).  Reference: loss.backward() through nn.Linear, pretrain_src/train_r2r.py:258.
#include "common.cuh"
#include "host_util.h"

namespace gmm {

// src [R, C] (fp32 or fp16, row pitch lds) -> dst [R, C] fp16 (optional, pitch ldd) and dst_t [C, r_pad] fp16 (optional, pitch ldt,
// columns R .. r_pad-1 zero-filled).  32 x 32 tiles through shared memory, 256 threads.
// colsum (optional, fp32 [C], zeroed by the caller): += the column sums of src (the bias gradient when src = dy), one atomicAdd per
// column and 32-row tile.
template <typename T>
__global__ void __launch_bounds__(256) cast_transpose_kernel(const T* __restrict__ src, long long lds, int R, int C, __half* dst,
                                                             long long ldd, __half* dst_t, long long ldt, int r_pad, float* colsum) {
    __shared__ __half tile[32][33];
    __shared__ float csum[8][32];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        __half h = __float2half_rn(0.0f);
        if (r < R && c < C) {
            const float v = static_cast<float>(src[static_cast<long long>(r) * lds + c]);
            acc += v;
            h = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
            if (dst) dst[static_cast<long long>(r) * ldd + c] = h;
        }
        tile[ty + 8 * i][tx] = h;
    }
    if (colsum) csum[ty][tx] = acc;
    if (!dst_t && !colsum) return;
    __syncthreads();
    if (colsum && ty == 0 && c0 + tx < C && r0 < R) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += csum[k][tx];
        atomicAdd(colsum + c0 + tx, t);
    }
    if (!dst_t) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;               // dst_t[c][r]
        if (c < C && r < r_pad) dst_t[static_cast<long long>(c) * ldt + r] = tile[tx][ty + 8 * i];
    }
}

// out[n] (+)= sum_m dy[m, n]   (fp32 source): one CTA per 128 columns and row slab, atomicAdd of the slab sums
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, long long ld, int M, int N, float* out, int rows_per_cta) {
    const int c = blockIdx.x * 128 + (threadIdx.x & 127);
    const int half = threadIdx.x >> 7;                              // two row phases per CTA
    const int m0 = blockIdx.y * rows_per_cta, m1 = min(m0 + rows_per_cta, M);
    float s = 0.f;
    if (c < N)
        for (int m = m0 + half; m < m1; m += 2) s += dy[static_cast<long long>(m) * ld + c];
    __shared__ float red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    if (half == 0 && c < N) atomicAdd(out + c, red[threadIdx.x] + red[threadIdx.x + 128]);
}

}  // namespace gmm

// src [R, C] fp32 (src_is_f16 = 0) or fp16 -> dst [R, C] fp16 (may be NULL) and dst_t [C, r_pad] fp16 (may be NULL; r_pad >= R,
// the padding columns are written as zeros); pitches in elements.
extern "C" int gridmm_cast_transpose_f16(const void* src, int src_is_f16, long long lds, int R, int C, void* dst, long long ldd,
                                         void* dst_t, long long ldt, int r_pad, cudaStream_t stream) {
    using namespace gmm;
    if (R <= 0 || C <= 0) return 0;
    if (!src || (!dst && !dst_t) || (dst_t && r_pad < R)) return GRIDMM_ERR_ARG;
    const int rows = dst_t ? r_pad : R;
    dim3 grid((C + 31) / 32, (rows + 31) / 32);
    if (src_is_f16)
        cast_transpose_kernel<__half><<<grid, 256, 0, stream>>>(reinterpret_cast<const __half*>(src), lds, R, C, reinterpret_cast<__half*>(dst),
                                                                 ldd, reinterpret_cast<__half*>(dst_t), ldt, r_pad, nullptr);
    else
        cast_transpose_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(src), lds, R, C, reinterpret_cast<__half*>(dst), ldd,
                                                                reinterpret_cast<__half*>(dst_t), ldt, r_pad, nullptr);
    gridmm_count_launch(1);
    return static_cast<int>(cudaGetLastError());
}

// out[n] += sum_m dy[m, n] for fp32 dy [M, N] (pitch ld): the bias gradient of nn.Linear.  The caller zeroes `out` (or accumulates).
extern "C" int gridmm_colsum_f32(const float* dy, long long ld, int M, int N, float* out, cudaStream_t stream) {
    using namespace gmm;
    if (M <= 0 || N <= 0) return 0;
    if (!dy || !out) return GRIDMM_ERR_ARG;
    const int slabs = M >= 4096 ? 16 : (M >= 512 ? 4 : 1);
    const int rows_per_cta = (M + slabs - 1) / slabs;
    colsum_kernel<<<dim3((N + 127) / 128, slabs), 256, 0, stream>>>(dy, ld, M, N, out, rows_per_cta);
    gridmm_count_launch(1);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int gridmm_linear_f16(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias,
                                 const float* residual, int ld_res, float* out_f32, int ld_f32, void* out_f16, int ld_f16, int act,
                                 const int* m_dev, cudaStream_t stream);

// Forward of a trainable nn.Linear in one call: x fp32 [M, K] (pitch ldx) -> x16 [M, K] (scratch) and x16t [K, m_pad] (kept by the
// caller for the weight gradient; m_pad = M rounded up to 64), then y[M, N] = x16 . w16[N, K]^T + bias (fp32).
extern "C" int gridmm_linear_train_fwd(const float* x, long long ldx, int M, int K, const void* w16, int N, const float* bias, float* y,
                                       void* x16, void* x16t, int m_pad, cudaStream_t stream) {
    using namespace gmm;
    if (M <= 0) return 0;
    if (!x || !w16 || !y || !x16 || !x16t || m_pad < M || (m_pad & 63) || (K & 127) || (N & 127)) return GRIDMM_ERR_ARG;
    dim3 grid((K + 31) / 32, (m_pad + 31) / 32);
    cast_transpose_kernel<float><<<grid, 256, 0, stream>>>(x, ldx, M, K, reinterpret_cast<__half*>(x16), K, reinterpret_cast<__half*>(x16t),
                                                            m_pad, m_pad, nullptr);
    gridmm_count_launch(1);
    return gridmm_linear_f16(x16, K, w16, K, M, N, K, bias, nullptr, 0, y, N, nullptr, 0, 0, nullptr, stream);
}

// Backward of the same layer in one call: dy fp32 [M, N] (pitch lddy) -> dy16 [M, N], dy16t [N, m_pad] (both scratch) and, fused in
// the cast, db[N] = column sums of dy (overwritten; may be NULL); dx[M, K] = dy16 . w16t[K, N]^T (NULL = skip);
// dw[N, K] = dy16t . x16t[K, m_pad]^T (NULL = skip).
extern "C" int gridmm_linear_train_bwd(const float* dy, long long lddy, int M, int N, int K, const void* w16t, const void* x16t, int m_pad,
                                       void* dy16, void* dy16t, float* dx, float* dw, float* db, cudaStream_t stream) {
    using namespace gmm;
    if (M <= 0) return 0;
    if (!dy || !dy16 || !dy16t || m_pad < M || (m_pad & 63) || (K & 127) || (N & 127) || (dx && !w16t) || (dw && !x16t)) return GRIDMM_ERR_ARG;
    if (db) {
        cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * N, stream);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    dim3 grid((N + 31) / 32, (m_pad + 31) / 32);
    cast_transpose_kernel<float><<<grid, 256, 0, stream>>>(dy, lddy, M, N, reinterpret_cast<__half*>(dy16), N, reinterpret_cast<__half*>(dy16t),
                                                            m_pad, m_pad, db);
    gridmm_count_launch(1);
    int rc = 0;
    if (dx) rc = gridmm_linear_f16(dy16, N, w16t, N, M, K, N, nullptr, nullptr, 0, dx, K, nullptr, 0, 0, nullptr, stream);
    if (rc == 0 && dw) rc = gridmm_linear_f16(dy16t, m_pad, x16t, m_pad, N, K, m_pad, nullptr, nullptr, 0, dw, K, nullptr, 0, 0, nullptr, stream);
    return rc;
}
