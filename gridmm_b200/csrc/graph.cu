// Node embeddings of the episodes' topological maps, kept on the device (second half).
// The reference agent keeps, per episode, a Python dict  viewpoint -> [sum of embeddings, count]  (map_nav_src/models/graph_utils.py:
// 114-125) and updates it every step with per-episode Python loops over GPU tensors (map_nav_src/r2r/agent.py:306-320):
//     avg_pano = sum_v pano_embeds[v] * mask[v] / sum_v mask[v]
//     node[current] = [avg_pano, 1]                                  (rewrite)
//     for each candidate j that is not visited: node[cand_j] += [pano_embeds[j], 1]
// and reads it back as sum / count when it stacks gmap_img_embeds (agent.py:126-129: a zero row for the stop node first).
// Here the sums and counts live in [B, N, D] / [B, N] device arrays, the host only keeps the viewpoint -> slot maps and sends
// two small index arrays per step; one launch updates every episode, one launch gathers gmap_img_embeds.
#include "common.cuh"
#include "host_util.h"

namespace gmm {

__global__ void __launch_bounds__(256) gmap_update_kernel(const float* __restrict__ pano, const unsigned char* __restrict__ mask, int V, int D,
                                                          const int* __restrict__ cur_slot, const int* __restrict__ cand_slot,
                                                          float* __restrict__ node_sum, float* __restrict__ node_cnt, int N) {
    pdl_wait();
    const int b = blockIdx.y, d = blockIdx.x * blockDim.x + threadIdx.x;
    const int cur = cur_slot[b];
    if (cur < 0 || d >= D) return;                     // episode has ended: its map is frozen
    const float* pb = pano + static_cast<size_t>(b) * V * D + d;
    float acc = 0.f, n = 0.f;
    for (int v = 0; v < V; ++v)
        if (mask[b * V + v]) { acc += pb[static_cast<size_t>(v) * D]; n += 1.f; }
    float* sb = node_sum + static_cast<size_t>(b) * N * D + d;
    sb[static_cast<size_t>(cur) * D] = acc / n;
    if (d == 0) node_cnt[b * N + cur] = 1.f;
    for (int j = 0; j < V; ++j) {
        const int s = cand_slot[b * V + j];
        if (s < 0) continue;
        sb[static_cast<size_t>(s) * D] += pb[static_cast<size_t>(j) * D];
        if (d == 0) node_cnt[b * N + s] += 1.f;
    }
    pdl_launch_dependents();
}

__global__ void __launch_bounds__(256) gmap_gather_kernel(const float* __restrict__ node_sum, const float* __restrict__ node_cnt, int N, int D,
                                                          const int* __restrict__ slots, int G, float* __restrict__ out) {
    pdl_wait();
    const int b = blockIdx.z, g = blockIdx.y, d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const int s = slots[b * G + g];
    float v = 0.f;                                     // the stop node and the padding rows are zero vectors
    if (s >= 0) v = node_sum[(static_cast<size_t>(b) * N + s) * D + d] / node_cnt[b * N + s];
    out[(static_cast<size_t>(b) * G + g) * D + d] = v;
    pdl_launch_dependents();
}

}  // namespace gmm

extern "C" int gridmm_gmap_update(const float* pano_embeds, const unsigned char* pano_masks, int n_views, int dim, const int* cur_slot,
                                  const int* cand_slot, float* node_sum, float* node_cnt, int n_nodes, int batch, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0) return 0;
    if (!pano_embeds || !pano_masks || !cur_slot || !cand_slot || !node_sum || !node_cnt) return GRIDMM_ERR_ARG;
    if (n_views <= 0 || dim <= 0 || n_nodes <= 0) return GRIDMM_ERR_SHAPE;
    GMM_CUDA_CHECK(launch_pdl(gmap_update_kernel, dim3((dim + 255) / 256, batch), dim3(256), 0, stream, pano_embeds, pano_masks, n_views, dim,
                              cur_slot, cand_slot, node_sum, node_cnt, n_nodes));
    gridmm_count_launch(1);
    return 0;
}

extern "C" int gridmm_gmap_gather(const float* node_sum, const float* node_cnt, int n_nodes, int dim, const int* slots, int gmap_len,
                                  int batch, float* out, cudaStream_t stream) {
    using namespace gmm;
    if (batch <= 0 || gmap_len <= 0) return 0;
    if (!node_sum || !node_cnt || !slots || !out) return GRIDMM_ERR_ARG;
    if (dim <= 0 || n_nodes <= 0 || gmap_len > 65535) return GRIDMM_ERR_SHAPE;
    GMM_CUDA_CHECK(launch_pdl(gmap_gather_kernel, dim3((dim + 255) / 256, gmap_len, batch), dim3(256), 0, stream, node_sum, node_cnt, n_nodes,
                              dim, slots, gmap_len, out));
    gridmm_count_launch(1);
    return 0;
}
