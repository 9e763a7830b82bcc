/* gridmm_b200 -- C ABI of the B200 (sm_100a) implementation of GridMM's per-navigation-step hot path.
 *
 * The reference (MrZihan/GridMM) is pure Python/PyTorch: there is no FFI to bind.  These entry points are
 * what a ctypes/cffi stub on the reference side binds instead of the Python code cited at each function;
 * gridmm_b200/_lib.py is that stub, INTEGRATION.md shows the three call sites a maintainer changes.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the comment says "host";
 *   - every function is asynchronous on `stream` (pass torch's current stream), re-entrant per device,
 *     keeps no global state besides a launch counter, allocates nothing;
 *   - return value: 0 = ok, > 0 = cudaError_t of the failed runtime call / launch,
 *     GRIDMM_ERR_* (< 0) = rejected arguments.  No entry point has a CPU fallback.
 *   - fp16 = IEEE binary16 ("half"); matrices are row-major with an explicit pitch in ELEMENTS.
 */
#ifndef GRIDMM_B200_H
#define GRIDMM_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRIDMM_OK 0
#define GRIDMM_ERR_SHAPE (-1)  /* unsupported size or alignment */
#define GRIDMM_ERR_DRIVER (-2) /* cuTensorMapEncodeTiled unavailable or failed */
#define GRIDMM_ERR_ARG (-3)    /* null pointer / inconsistent arguments */

int gridmm_abi_version(void);
long long gridmm_launch_count(void);      /* kernels launched through this library since the last reset */
void gridmm_launch_count_reset(void);

/* ---- stage 1: grid build ------------------------------------------------------------------------------
 * Replaces EnvBatch.getGlobalMap + get_rel_position + get_gridmap_pos_fts
 * (map_nav_src/r2r/env.py:115-121, 242-265, 267-374; same code in reverie/env.py, rxr/env.py,
 * pretrain_src/data/dataset.py:351-473; CE geometry VLN_CE/.../Policy_ViewSelection_GridMap.py:632-641, 689-825)
 * for ALL `batch` episodes in one launch.  Appends the 588 points (12 horizon views x 7x7 patch centres) of the
 * current viewpoint to the per-episode state, then re-assigns every accumulated point to a cell of the
 * egocentric grid_w x grid_w window (bit-exact with the reference's fp32 numpy arithmetic), and sorts the valid
 * points by cell (stable) for gridmm_pool.
 *   depth        [batch,588] uint16 (depth_is_f32=0; value/depth_scale = metres) or float32 (depth_is_f32=1)
 *   pose         [batch,4]   px, py, cos(angle), sin(angle); angle = -heading (+pi for CE), evaluated in double on the
 *                            host and rounded to fp32 exactly like `python float * np.float32 array`
 *   view_cs      [batch,12,2] cos/sin of each view's angle (R2R: v*pi/6; CE: v*pi/6 - heading), host double -> fp32
 *   active       [batch] or NULL; 0 = no new viewpoint for this episode (NULL = all active, the reference's behaviour)
 *   off7         HOST pointer, 7 floats: f32(o_k) * f32(tan(hfov/2)), o = -6/7..6/7 (env.py:118)
 *   pos_mode     0: discrete-env cell features [sin h, cos h, 0, 1, dist/max_dist] (env.py:242-265); 1: the CE code's
 *                (x, z, y) reading of the same call (VLN_CE/.../Policy_ViewSelection_GridMap.py:661-684, models/utils.py:125-144);
 *                max_dist = 30 (env.py:47), 25 (R2R-CE), 40 (RxR-CE)
 *   state        wx, wy [batch,cap] f32; valid [batch,cap] u8; bounds [batch,4] = max_x,min_x,max_y,min_y
 *                (initialise to -10000,10000,-10000,10000: env.py:187-190); n_pts [batch] int32 (initialise to 0)
 *   outputs      cell [batch,cap] int16 (-1 = masked: env.py:306,366-369); half_len [batch];
 *                perm [batch,cap] int32; cell_start [batch,grid_w^2+1]; cell_rank [batch,grid_w^2]; n_nonempty [batch];
 *                pos_fts [batch,grid_w^2,5]
 * An episode whose n_pts + 588 would exceed cap is left unchanged (the host wrapper grows the buffers first). */
int gridmm_grid_update(int batch, const void* depth, int depth_is_f32, float depth_scale, const float* pose,
                       const float* view_cs, const unsigned char* active, const float* off7, int flip_y, int negate_map_x,
                       int pos_mode, float max_dist, int grid_w, int cap, float* wx, float* wy, unsigned char* valid, float* bounds, int* n_pts,
                       short* cell, float* half_len, int* perm, int* cell_start, int* cell_rank, int* n_nonempty,
                       float* pos_fts, const int* new_slot, int* slots, int t_cap, cudaStream_t stream);
/* new_slot (optional, device int[batch]) + slots (device int[batch, t_cap]): device-resident feature DB -- the CLIP tokens of every
 * viewpoint already live in HBM (the whole Matterport feature DB is ~10 GB of 180), the step only names the slab slot of each
 * episode's new viewpoint; the kernel records it as slots[b, (viewpoints so far)], the table gridmm_pool resolves rows through. */

/* Sort-only variant for callers that already hold the reference's `grid_map` tensors (cell id per point, -1 = masked,
 * r2r/env.py:611): cell [batch,cap] int16 and n_pts [batch] are INPUTS; outputs as above. */
int gridmm_cell_sort(int batch, const short* cell, const int* n_pts, int grid_w, int cap, int* perm, int* cell_start,
                     int* cell_rank, int* n_nonempty, cudaStream_t stream);

/* ---- stage 2: instruction-relevance pooling --------------------------------------------------------------
 * Replaces the B x 196 Python loop of GlocalTextPathNavCMT.forward_navigation_per_step
 * (map_nav_src/models/vilmodel.py:796-807; pretrain_src/model/vilmodel.py:688-700;
 * VLN_CE/.../gridmap/vilmodel.py:720-735).  For every episode b and non-empty cell c:
 *     w_j = max_l <x_j, text_fts[b,l]>      over ALL l in [0,l_pad)  (padding positions included, vilmodel.py:798)
 *     pooled[b, rank(c)] = sum_{j in c} softmax_c(w)_j * x_j          (fp32 accumulate, fp16 result)
 * grid_proj is applied afterwards with gridmm_linear_f16 (it commutes with the convex combination).
 *   fts          fp16 feature slab [fts_rows, feat_dim], row r at fts + r*feat_dim; point (step t, view v, patch k) of episode
 *                b is row slots[b*t_cap+t]*slot_rows + v*view_rows + tok_off + k (CLS token skipped via tok_off,
 *                env.py:299); rows are fetched with TMA tile::gather4 through a tensor map over the whole slab
 *   text_fts     fp16 [batch, l_pad, feat_dim] = text_proj(txt_embeds), 16-byte aligned; l_pad <= 256 (the reference's launch
 *                scripts use --max_instr_len 200, 250 for RxR: scripts/run_r2r.sh:38, run_rxr.sh:38).  The operand lives in
 *                tensor memory, one text position per TMEM lane (unused lanes replicate a real position): up to 128 positions
 *                are ONE pass over the features; 129..256 positions take two (the first only produces the row maxima over
 *                positions 128.. into w_scratch, the second merges them before the softmax).  May be NULL when
 *                text_ws_ready != 0
 *   text_ws      workspace, ceil(l_pad/128) * batch * 128 * feat_dim * 2 bytes, 16-byte aligned: lane-major copy of text_fts
 *                as [ceil(l_pad/128)][batch][feat_dim/8][128] 16-byte units -- written here (text_ws_ready = 0) or already
 *                produced by gridmm_linear_f16_lanes (text_ws_ready = 1)
 *   pooled       fp16 [batch, n_cells, feat_dim], rows >= n_nonempty[b] are not written
 *   w_out        optional f32 [batch,cap]: w per sorted position (tests), or NULL
 *   w_scratch    f32 [batch,cap], required when l_pad > 128 (else may be NULL)
 *   pool_ws      workspace of gridmm_pool_ws_bytes(batch, feat_dim, num_ctas) bytes, 16-byte aligned: the work plan (one
 *                contiguous range of the batch's sorted valid rows per CTA, equal in cost; a range may end in the middle of a
 *                large cell) and the un-normalised partials of cells that are pooled in pieces by several CTAs (merged by the
 *                last piece to arrive, in CTA order: the result is deterministic).  plan_ready = 0: the plan is computed here
 *                (one extra small launch); 1: gridmm_pool_plan already ran for this cell_start / num_ctas (e.g. on the stream
 *                of gridmm_grid_update, concurrently with the text branch)
 *   num_ctas     0 = one CTA per SM */
int gridmm_pool(const void* fts, long long fts_rows, int feat_dim, const int* slots, int t_cap, int slot_rows, int view_rows,
                int tok_off, const int* perm, int cap, const int* cell_start, const int* cell_rank, int n_cells,
                const void* text_fts, int l_pad, int batch, void* text_ws, int text_ws_ready, void* pooled, float* w_out,
                float* w_scratch, void* pool_ws, int plan_ready, int num_ctas, cudaStream_t stream);
/* Work plan of gridmm_pool (replaces nothing in the reference: the reference loops over episodes and cells serially,
 * vilmodel.py:796-807; this is the load balancing of that loop over the SMs).  Returns the byte size / fills the workspace. */
long long gridmm_pool_ws_bytes(int batch, int feat_dim, int num_ctas);
int gridmm_pool_plan(const int* cell_start, int n_cells, int batch, int feat_dim, int num_ctas, void* pool_ws, cudaStream_t stream);

/* ---- caller side of the step: node embeddings of the episodes' topological maps ------------------------
 * Replaces GraphMap.update_node_embed / get_node_embed (map_nav_src/models/graph_utils.py:114-125) and the per-episode loops
 * around them (map_nav_src/r2r/agent.py:306-320, 126-129): sums [batch, n_nodes, dim] and counts [batch, n_nodes] stay on the
 * device (zero-initialised by the caller at the start of an episode batch), the host keeps the viewpoint -> slot maps.
 *   gridmm_gmap_update: avg = masked mean of pano_embeds[b] over its n_views tokens; node cur_slot[b] = (avg, 1) (rewrite);
 *                       for every token j with cand_slot[b, j] >= 0: node cand_slot[b, j] += (pano_embeds[b, j], 1), in token
 *                       order.  cur_slot[b] < 0: the episode has ended, nothing changes.
 *   gridmm_gmap_gather: out[b, g] = sum / count of node slots[b, g], or a zero row for slots[b, g] < 0 (stop node, padding). */
int gridmm_gmap_update(const float* pano_embeds, const unsigned char* pano_masks, int n_views, int dim, const int* cur_slot,
                       const int* cand_slot, float* node_sum, float* node_cnt, int n_nodes, int batch, cudaStream_t stream);
int gridmm_gmap_gather(const float* node_sum, const float* node_cnt, int n_nodes, int dim, const int* slots, int gmap_len,
                       int batch, float* out, cudaStream_t stream);

/* ---- stage 3: cross-modal encoder blocks ------------------------------------------------------------------
 * nn.Linear on tcgen05: out = act(a[M,K] . w[N,K]^T + bias) + residual; fp16 operands, fp32 accumulate.
 * Replaces every nn.Linear of vilmodel.py:95-209, 317-379, 663-674, 702-703 and transformer.py:133-182.
 * N % 128 == 0, K % 64 == 0; act: 0 none, 1 GELU(erf), 2 ReLU; either output may be NULL. */
int gridmm_linear_f16(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias,
                      const float* residual, int ld_res, float* out_f32, int ld_f32, void* out_f16, int ld_f16, int act,
                      const int* m_dev, cudaStream_t stream);
/* m_dev (here and in gridmm_linear_ln_f16): optional DEVICE int, the number of rows to process (<= M) when only the GPU knows it
 * (packed / ragged operands); the launch is sized for M, row tiles past *m_dev do nothing.  NULL = M rows. */

/* nn.Linear(K -> 768) + residual + LayerNorm in one kernel (a cluster of 2 or 6 CTAs per 128-row tile, row statistics merged
 * through distributed shared memory):  v = a . w^T + bias + residual;  y = LayerNorm(v; gamma, beta, eps).
 * Replaces BertSelfOutput / BertOutput / BertOutAttention.output (vilmodel.py:155-170, 196-209, 370-379: out_f32 = out_f16 = y)
 * and `x = x + out_proj(a); norm2(x)` of TransformerEncoderLayer.forward_pre (transformer.py:170-182: f32_raw = 1, out_f32 = v,
 * out_f16 = y).  N must be 768, K % 64 == 0; out_f32 may alias residual. */
int gridmm_linear_ln_f16(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias,
                         const float* residual, int ld_res, const float* gamma, const float* beta, float eps, float* out_f32,
                         int ld_f32, void* out_f16, int ld_f16, int f32_raw, const int* m_dev, cudaStream_t stream);

/* gridmm_grid_assemble + the gmap tokens (rows [n_cells, seq) of the map sequence: gmap_img + step_table[step_ids] +
 * LN(Linear(gmap_pos)), vilmodel.py:828-831; mask from gmap_mask) + the first pre-norm LayerNorm of grid_encoder
 * (transformer.py:170-172: map_f16 = LN(map_f32; norm_gamma, norm_beta, norm_eps)) in one launch.  gw is the TRANSPOSED weight
 * [gmap_kin, 768] like w. */
int gridmm_map_inputs(const float* proj, const float* pos_fts, const int* cell_rank, const int* n_nonempty, const float* w,
                      const float* bias, const float* gamma, const float* beta, const float* gmap_pos, int gmap_kin,
                      const float* gw, const float* gbias, const float* ggamma, const float* gbeta, const float* gmap_img,
                      const float* step_table, const long long* step_ids, const unsigned char* gmap_mask,
                      const float* norm_gamma, const float* norm_beta, float norm_eps, float* map_f32, void* map_f16,
                      unsigned char* map_mask, int batch, int n_cells, int seq, int hidden, cudaStream_t stream);

/* ---- packed ("ragged") fusion-encoder context: masked context rows (empty grid-cell slots, padded text) are dropped before the
 * K/V projection of the 4 fusion layers and before their cross-attention (vilmodel.py:843-853; a masked key has weight
 * exp(-10000) = 0 in fp32, so the result is the same).
 * gridmm_kv_index: kv_off[b] = valid rows of episodes < b (kv_off[batch] = total), kv_cnt[b], kv_pos[b, r] = packed row or -1,
 *   for the context [map (S rows, map_mask) ; txt (L rows, txt_mask)]; S + L <= 1024.
 * gridmm_linear_f16_rows: gridmm_linear_f16 (fp16 output) over the first *m_dev rows only (m_dev on the device, <= M).
 * gridmm_attention_varlen_f16: attention whose keys / values of episode b are rows k_off[b] .. + k_cnt[b] of k / v, all valid. */
int gridmm_kv_index(const unsigned char* map_mask, const unsigned char* txt_mask, int batch, int S, int L, int* kv_pos, int* kv_off,
                    int* kv_cnt, cudaStream_t stream);
int gridmm_linear_f16_rows(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias, void* out_f16,
                           int ld_f16, const int* m_dev, cudaStream_t stream);
int gridmm_attention_varlen_f16(const void* q, int ldq, int q_rows, const void* k, int ldk, const void* v, int ldv,
                                const int* k_off, const int* k_cnt, int max_sk, long long k_total, const float* k_bias, void* o,
                                int ldo, int batch, int heads, int sq, float scale, cudaStream_t stream);
/* k_bias: optional f32 per packed key row, added to that key's scores (log multiplicity of a de-duplicated key), or NULL.
 * k_total: rows of the k / v buffers (extent of the TMA tensor maps of the tcgen05 head-pair kernel, which serves sq <= 64 and
 * max_sk <= 256: two heads of one episode share the 128 TMEM lanes); 0 = unknown (mma.sync kernel). */

/* ---- packed ("ragged") MAP sequence ---------------------------------------------------------------------------
 * The reference pads every episode's map sequence [grid cells ; gmap nodes] to the batch maximum (vilmodel.py:813-838: on
 * BASELINE config 2 about 40 % of the padded rows are masked).  Here the rows that matter are packed back to back and every
 * map-sized launch (grid_encoder, grid_txt_encoder: transformer.py:170-182, vilmodel.py:399-414) runs over them only:
 *   episode b owns packed rows m_off[b] .. m_off[b+1]-1 = [ k_b non-empty cells (rank order) | q_b | G gmap nodes ]
 * q_b in {0,1} is ONE representative of the z_b zero-vector slots that the reference's mask-aliasing quirk flags valid
 * (gridmm_grid_assemble); wherever it acts as an attention key its score gets + log(z_b) (z identical keys: exact).
 * gridmm_map_index (one CTA): m_off [batch+1], m_info [4][batch] = rows of k_b, k_b + q_b, rows per episode, z_b, m_logz [batch],
 *   cell_of_rank [batch][n_cells] (inverse of cell_rank), m_goff [batch] = first gmap row of every episode.
 * gridmm_map_inputs_packed: gridmm_map_inputs over the packed rows; kvalid [rows] u8 (1 = valid key: cells, valid gmap nodes),
 *   kbias [rows] f32 (log z_b on the representative row, else 0).
 * gridmm_attention_ragged_f16: tcgen05 attention with per-episode query rows q_off[b] .. + q_cnt[b] (also the rows of o); keys
 *   ragged too (k_off / k_cnt; kmask / kbias indexed by packed key row) or regular (k_off = k_cnt = NULL: rows b * k_rows .. + max_sk,
 *   kmask [batch, max_sk]); max_sq / max_sk bound the per-episode counts (<= 320 keys), q_total / k_total = rows of the buffers.
 * gridmm_kv_index_packed / gridmm_fusion_inputs_packed: the fusion encoder's packed context [valid map rows ; valid text rows] and
 *   its queries [gmap' ; vp] read from the packed map (kv_src[r] >= 0: packed map row, < 0: text row -1 - (b * L + l);
 *   kv_bias[r] = that key's score bias). */
int gridmm_map_index(const int* cell_rank, const int* n_nonempty, int batch, int n_cells, int G, int* m_off, int* m_info,
                     float* m_logz, int* cell_of_rank, int* m_goff, cudaStream_t stream);
int gridmm_map_inputs_packed(const float* proj, const float* pos_fts, const int* cell_of_rank, const int* m_off, const int* m_info,
                             const float* m_logz, const float* w, const float* bias, const float* gamma, const float* beta,
                             const float* gmap_pos, int gmap_kin, const float* gw, const float* gbias, const float* ggamma,
                             const float* gbeta, const float* gmap_img, const float* step_table, const long long* step_ids,
                             const unsigned char* gmap_mask, const float* norm_gamma, const float* norm_beta, float norm_eps,
                             float* map_f32, void* map_f16, unsigned char* kvalid, float* kbias, int batch, int n_cells, int G,
                             int hidden, cudaStream_t stream);
int gridmm_attention_ragged_f16(const void* q, int ldq, const int* q_off, const int* q_cnt, int max_sq, long long q_total,
                                const void* k, int ldk, const void* v, int ldv, const int* k_off, const int* k_cnt, int k_rows,
                                int max_sk, long long k_total, const unsigned char* kmask, const float* kbias, float mask_neg,
                                void* o, int ldo, int batch, int heads, float scale, cudaStream_t stream);
int gridmm_kv_index_packed(const int* m_off, const unsigned char* kvalid, const float* kbias, const unsigned char* txt_mask, int batch,
                           int L, int* kv_src, float* kv_bias, int* kv_off, int* kv_cnt, cudaStream_t stream);
int gridmm_fusion_inputs_packed(const float* map32, const float* txt32, const int* kv_src, const int* kv_off, const int* m_goff,
                                const unsigned char* gmap_mask, const unsigned char* vp_mask, float* x32, void* x16, void* kv16,
                                unsigned char* q_mask, const float* vp_pos, int vp_kin, const float* vp_w, const float* vp_bias,
                                const float* vp_gamma, const float* vp_beta, const float* vp_img, int batch, int L, int G, int V,
                                int kv_rows_max, int hidden, cudaStream_t stream);

/* Inputs of the fusion encoder (vilmodel.py:828-833, 843-850) in one launch: x[b, :G] = map[b, S-G:] (fp32 + fp16),
 * x[b, G:] = vp_img + LN(Linear(vp_pos)) when vp_pos is given (vp_w = TRANSPOSED weight [vp_kin, 768]; with vp_pos NULL rows G..
 * of x must hold the vp tokens already), kv16[b] = fp16([map[b] ; txt[b]]) -- or, with kv_pos (gridmm_kv_index), only the valid
 * rows at their packed positions --, kv_mask = [map_mask ; txt_mask], q_mask = [gmap_mask ; vp_mask]. */
int gridmm_fusion_inputs(const float* map32, const float* txt32, const unsigned char* map_mask, const unsigned char* txt_mask,
                         const unsigned char* gmap_mask, const unsigned char* vp_mask, float* x32, void* x16, void* kv16,
                         unsigned char* kv_mask, unsigned char* q_mask, const int* kv_pos, const float* vp_pos, int vp_kin,
                         const float* vp_w, const float* vp_bias, const float* vp_gamma, const float* vp_beta, const float* vp_img,
                         int batch, int S, int L, int G, int V, int hidden, cudaStream_t stream);

/* ---- action heads (vilmodel.py:663-674, 859-907) in three launches ------------------------------------------
 * ClsPrediction = Linear, ReLU, LayerNorm(1e-12), Linear(768 -> 1):  logit = rstd * (S3 - mean * c1) + c0 with r = ReLU(xW + b),
 * S1 = sum r, S2 = sum r^2, S3 = sum r * gamma * w2, c1 = sum gamma * w2, c0 = sum beta * w2 + b2.
 * gridmm_head_rows: up to 6 row segments of fp32 matrices (host arrays of device pointers / ints) -> one [hi | lo | hi] fp16
 *   operand matrix (row = out_row0[i] + b * rows_per_b[i] + r  <-  x[i][b * in_rows_per_b[i] + in_off[i] + r]).
 * gridmm_cls_heads_f16: grouped tcgen05 GEMM over `tiles_m` 128-row tiles; grp (device) = per tile {first A row, first W row,
 *   first output row, mode}; w = stacked [groups * N, K] split weights; mode 0: cls_part[out_row][N / 64][3] = (S1, S2, S3) per
 *   64 columns; mode 1: cls_raw[out_row][N] = x . W^T (the two K halves of sap_fuse_linear's first layer, vilmodel.py:859-862).
 * gridmm_nav_logits2: finishes every head from its sums (consts [5][2] = (c1, c0) of global, local, grid, obj, fuse; the fuse
 *   head from fuse_raw rows row_fuse_g + b, row_fuse_v + b) and fuses the logits exactly like gridmm_nav_logits;
 *   row_obj < 0 / fuse_raw NULL disable the object head / dynamic fusion. */
int gridmm_head_rows(int nseg, const float* const* x, const int* ldx, const int* in_rows_per_b, const int* in_off,
                     const int* rows_per_b, const int* out_row0, const int* const* row_off, int batch, void* out_f16, int ld_f16,
                     int hidden, cudaStream_t stream);
/* row_off: NULL, or a host array of nseg device pointers (each NULL or int[batch]): first input row of every episode of that
 * segment (a packed / ragged source: the gmap rows of the packed map sequence), replacing b * in_rows_per_b[i]. */
int gridmm_cls_heads_f16(const void* a, int lda, long long a_rows, const void* w, int ldw, int groups, int tiles_m, int N, int K,
                         const float* bias, const float* gw2, const int* grp, float* cls_part, float* cls_raw,
                         cudaStream_t stream);
int gridmm_nav_logits2(const float* part, const float* fuse_raw, const float* fuse_bias, const float* fuse_gw2, int row_fuse_g,
                       int row_fuse_v, const float* consts, int row_global, int row_local, int row_grid, int row_obj,
                       const unsigned char* gmap_masks, const unsigned char* gmap_visited, const unsigned char* vp_nav_masks,
                       const unsigned char* vp_obj_masks, const int* fuse_src, const unsigned char* bw_mask, const int* cand_node,
                       float* global_logits, float* grid_logits, float* local_logits, float* fused_logits, float* obj_logits,
                       int batch, int G, int V, cudaStream_t stream);
/* cand_node != NULL selects the mask-free form of the vpid tables of vilmodel.py:881-899 (built from the vpid strings alone, so
 * the host never reads gmap_visited back): fuse_src[b,g] = last candidate slot holding node g's viewpoint (-2 none, -1 [stop] /
 * padding), applied only to unvisited nodes; cand_node[b,v] = gmap slot of candidate v's viewpoint (-1 none): candidate v
 * counts as "already visited" when gmap_visited[b, cand_node[b,v]] is set.  bw_mask is then unused (may be NULL). */

/* dst[i][0 .. nbytes[i]) = src[i][0 .. nbytes[i]) for i < n <= 24 in ONE launch (src / dst / nbytes: host arrays of device-
 * accessible pointers; sources may be pinned host memory).  Lands the ~14 per-step input tensors of forward('navigation')
 * (r2r/agent.py:163-205) in the static buffers the captured graph reads. */
int gridmm_copy_segments(int n, const void* const* src, void* const* dst, const long long* nbytes, cudaStream_t stream);

/* text_proj (vilmodel.py:702, 793-795) written straight into gridmm_pool's lane-major operand layout:
 * out_lanes[t / 128][b][u][t % 128] (16-byte units, 128 slots per unit row) = (a[b*rows_per_b + t, :] . w^T + bias)[8u .. 8u+7];
 * M = batch*rows_per_b, rows_per_b <= 256 (a second [batch][N/8][128] block holds positions 128..). */
int gridmm_linear_f16_lanes(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias,
                            void* out_lanes, int rows_per_b, cudaStream_t stream);

/* softmax(q k^T * scale + mask) v per (episode, head); head dim 64; q/k/v/o fp16 with pitches; kmask [batch,sk] u8,
 * masked keys get `mask_neg` added (-10000: models/ops.py:25-34; -inf: key_padding_mask, transformer.py:176). */
int gridmm_attention_f16(const void* q, int ldq, int q_rows, const void* k, int ldk, const void* v, int ldv, int k_rows,
                         void* o, int ldo, const unsigned char* kmask, float mask_neg, int batch, int heads, int sq, int sk,
                         float scale, cudaStream_t stream);

/* LayerNorm over 768 (BertLayerNorm = torch.nn.LayerNorm, eps 1e-12: map_nav_src/models/vilmodel.py:40-44, 163-169;
 * nn.LayerNorm eps 1e-5 inside the pre-norm layers: models/transformer.py:144-145, 170-180); fp32 and/or fp16 output. */
int gridmm_layernorm(const float* x, int ldx, const float* gamma, const float* beta, float eps, float* out_f32, int ld_f32,
                     void* out_f16, int ld_f16, int rows, int hidden, cudaStream_t stream);

/* out[b, out_off + r] = in[b, in_off + r], r < rows_per_b: fp32 rows to fp32 and/or fp16 rows of another per-episode
 * sequence (builds [map; txt], [gmap; vp] and the head inputs, vilmodel.py:843-850, 855-856, 863). */
int gridmm_copy_rows(const float* x, int ldx, int in_rows_per_b, int in_off, float* out_f32, int ld_f32, void* out_f16,
                     int ld_f16, int out_rows_per_b, int out_off, int rows_per_b, int batch, int hidden, cudaStream_t stream);

/* out[b*rows_per_b + r] = [hi | lo | hi] of in[b, in_off + r] at column blocks 0, k_total, 2*k_total (hi = fp16(x),
 * lo = fp16(x - hi)): with weights [Wh | Wh | Wl] one K-concatenated gridmm_linear_f16 evaluates x.W to ~2^-22.
 * Used for the ClsPrediction GEMMs (vilmodel.py:663-674), where plain fp16 rounding dominates the logit error. */
int gridmm_split_rows(const float* x, int ldx, int in_rows_per_b, int in_off, void* out_f16, int ld_f16, int k_total,
                      int rows_per_b, int batch, int hidden, cudaStream_t stream);

/* out[b, off + r] = base + table[idx] + LayerNorm(Linear(kin -> 768)(feat))   (vilmodel.py:828-833)
 * w is the TRANSPOSED nn.Linear weight, [kin, 768] (coalesced reads). */
int gridmm_pos_embed(const float* feat, int kin, const float* w, const float* bias, const float* gamma, const float* beta,
                     float eps, const float* base, const float* table, const long long* idx, float* out_f32, void* out_f16,
                     int in_rows_per_b, int out_rows_per_b, int out_row_off, int rows, int hidden, cudaStream_t stream);

/* BERT text embeddings: LayerNorm(word[ids] + position[0..L) + token_type[0]), eps = config.layer_norm_eps (1e-12 for
 * bert-base, 1e-5 for xlm-roberta-base: BertEmbeddings.forward, vilmodel.py:75-93); ids int64 [batch, L]. */
int gridmm_text_embed(const long long* ids, const float* word, const float* pos, const float* type0, const float* gamma,
                      const float* beta, float eps, float* out_f32, void* out_f16, int batch, int L, int hidden,
                      cudaStream_t stream);

/* grid cells of the map sequence + validity mask incl. the reference's compaction quirk (vilmodel.py:813-823);
 * w is the TRANSPOSED grid_pos_embeddings.0 weight, [5, 768]. */
int gridmm_grid_assemble(const float* proj, const float* pos_fts, const int* cell_rank, const int* n_nonempty, const float* w,
                         const float* bias, const float* gamma, const float* beta, float* map_f32, unsigned char* map_mask,
                         int batch, int n_cells, int seq, int hidden, cudaStream_t stream);

/* ClsPrediction tail: logit = w2 . LayerNorm(h) + b2 (vilmodel.py:663-674; h = ReLU(Linear(x)) from gridmm_linear_f16) */
int gridmm_cls_tail(const float* h, const float* gamma, const float* beta, const float* w2, const float* b2, float* logit,
                    int rows, int hidden, cudaStream_t stream);

/* fuse weight, masking and global/local logit fusion (vilmodel.py:859-907); fuse_src/bw_mask are the integer form of the
 * reference's vpid-string loops (built on the host by gridmm_b200.model.build_fuse_index). */
int gridmm_nav_logits(const float* raw_global, const float* raw_grid, const float* raw_local, const float* raw_obj,
                      const float* raw_fuse, const unsigned char* gmap_masks, const unsigned char* gmap_visited,
                      const unsigned char* vp_nav_masks, const unsigned char* vp_obj_masks, const int* fuse_src,
                      const unsigned char* bw_mask, float* global_logits, float* grid_logits, float* local_logits,
                      float* fused_logits, float* obj_logits, int batch, int G, int V, cudaStream_t stream);

/* continuous-env action logits (VLN_CE/vlnce_baselines/models/gridmap/vilmodel.py:791-800):
 * fused[b, j] = global[b, j] * w + local[b, j] * (1 - w), w = sigmoid(raw_fuse[b]), j < maxc; -inf where vp_nav_masks = 0. */
int gridmm_ce_logits(const float* raw_global, const float* raw_local, const float* raw_fuse, const unsigned char* vp_nav_masks,
                     float* fused, int batch, int G, int V, int maxc, cudaStream_t stream);

/* ---- optimizer half of the pretraining step (BASELINE config 5) -------------------------------------------
 * After the NCCL all-reduce of the flat gradient buffer (gridmm_b200/train.py) every rank applies the same update:
 * torch.nn.utils.clip_grad_norm_ (pretrain_src/train_r2r.py:281-285) + AdamW with decoupled weight decay and bias correction
 * (pretrain_src/optim/adamw.py:57-104), one flat fp32 range per parameter group (weight decay 0.01 / 0: optim/misc.py:12-22).
 * gridmm_grad_sumsq: out[0] += sum g[i]^2 (zero `out` first; ranges may be accumulated).
 * gridmm_adamw_step: p, m, v updated in place from g * grad_scale (1 / world size, 1 / loss scale); sumsq != NULL clips by the
 *   global norm sqrt(*sumsq) * grad_scale against max_norm on the device (no host synchronisation); step >= 1. */
int gridmm_grad_sumsq(const float* g, long long n, float* out, cudaStream_t stream);
int gridmm_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int step, float grad_scale, const float* sumsq, float max_norm, cudaStream_t stream);

/* ---- training path helpers (BASELINE config 5): operands of the dgrad / wgrad GEMMs of nn.Linear ------------------------------
 * dx = dy . W and dW = dy^T . x run on gridmm_linear_f16 (A[M,K] . W[N,K]^T, both K-contiguous) with transposed fp16 operands:
 * gridmm_cast_transpose_f16: src [R, C] fp32 / fp16 -> dst [R, C] fp16 and / or dst_t [C, r_pad] fp16 (zero padded beyond R: the wgrad
 *   contraction runs over the rows, padded to the GEMM's K granularity of 64); pitches in elements.
 * gridmm_colsum_f32: out[n] += sum_m dy[m, n], the bias gradient.  (loss.backward() through nn.Linear, pretrain_src/train_r2r.py:258) */
int gridmm_cast_transpose_f16(const void* src, int src_is_f16, long long lds, int R, int C, void* dst, long long ldd, void* dst_t,
                              long long ldt, int r_pad, cudaStream_t stream);
int gridmm_colsum_f32(const float* dy, long long ld, int M, int N, float* out, cudaStream_t stream);

/* A trainable nn.Linear(K -> N) (K, N multiples of 128) in one call per direction, what torch.nn.functional.linear and its autograd
 * node do in the reference's training loop (pretrain_src/train_r2r.py:244-258):
 * forward:  x fp32 [M, K] (pitch ldx) -> x16 [M, K] (scratch) and x16t [K, m_pad] (the caller keeps it for backward; m_pad = M rounded
 *           up to 64), y[M, N] = x16 . w16[N, K]^T + bias (fp32 out, bias may be NULL);
 * backward: dy fp32 [M, N] (pitch lddy) -> dy16 [M, N] and dy16t [N, m_pad] (scratch); db[N] = column sums of dy, fused into the cast
 *           (overwritten; NULL = skip); dx[M, K] = dy16 . w16t[K, N]^T (NULL = skip); dw[N, K] = dy16t . x16t^T (NULL = skip). */
int gridmm_linear_train_fwd(const float* x, long long ldx, int M, int K, const void* w16, int N, const float* bias, float* y,
                            void* x16, void* x16t, int m_pad, cudaStream_t stream);
int gridmm_linear_train_bwd(const float* dy, long long lddy, int M, int N, int K, const void* w16t, const void* x16t, int m_pad,
                            void* dy16, void* dy16t, float* dx, float* dw, float* db, cudaStream_t stream);

/* Debug hooks (tools/microbench.py only): per-CTA clock64 counters written by the following launches ([grid][8] for the
 * GEMM, [grid][16] for the pooling kernel: role totals and time spent waiting on each mbarrier).  NULL disables. */
void gridmm_debug_set_gemm_counters(long long* dbg);
void gridmm_debug_set_pool_counters(long long* dbg);
void gridmm_debug_set_attn_legacy(int on);    /* gridmm_attention_f16: 1 forces the mma.sync kernel, 2 the tcgen05 one, 0 by shape */
void gridmm_debug_set_ln_cluster(int cl);     /* force the cluster size (2 / 6) of gridmm_linear_ln_f16; 0 = automatic */
void gridmm_debug_set_gemm_pairs(int on);    /* 0: disable the cta_group::2 GEMM path (A/B timing) */
void gridmm_debug_set_pool_trace(long long* t);   /* gridmm_pool: [4][64][8] clock64 stamps of the first tiles' stage hand-overs (CTAs 0..3) */
void gridmm_debug_set_pool_exp(int e);        /* gridmm_pool timing experiments (1: half rows, 2: L2-resident rows); results are garbage when != 0 */
void gridmm_debug_set_pool_plan(int episode_cost, int snap);   /* gridmm_pool_plan: rows an episode start costs, snap distance of a cut (< 0 keeps) */
void gridmm_debug_set_pool_split(int on);    /* gridmm_pool (mma.sync sums): 0 = single fp16 softmax weights, 1 = value + residual */
void gridmm_debug_set_gemm_384(int on);      /* 1: enable the 256 x 384 pair tiles of gridmm_linear_f16 (off by default: measured slower) */

#ifdef __cplusplus
}
#endif
#endif /* GRIDMM_B200_H */
