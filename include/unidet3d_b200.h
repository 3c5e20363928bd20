/*
 * unidet3d_b200 -- C-ABI of the B200-native UniDet3D forward hot path.
 *
 * The reference (filaPro/unidet3d) has no native code and no FFI: its hot path calls
 * Python APIs of un-vendored CUDA wheels (spconv, MinkowskiEngine, torch_scatter, mmcv,
 * torch).  Each entry point below therefore cites the reference *Python call site* whose
 * native work it replaces (paths relative to the reference checkout).  INTEGRATION.md
 * shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative UD3D_E* code on failure;
 *     ud3d_last_error() returns a thread-local message for the last failure;
 *   - all pointers are DEVICE pointers unless the name ends in _host; buffers are
 *     caller-owned, the library never allocates or frees caller-visible memory;
 *   - scratch memory comes from a caller-supplied workspace (pointer + size, with a
 *     *_workspace_bytes() query);
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises the host;
 *   - there is no CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef UNIDET3D_B200_H_
#define UNIDET3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UD3D_OK 0
#define UD3D_EINVAL (-1)    /* bad argument (shape / alignment / NULL) */
#define UD3D_ECUDA (-2)     /* CUDA runtime error (message has the CUDA string) */
#define UD3D_EWORKSPACE (-3)/* workspace too small */
#define UD3D_EUNSUPPORTED (-4)

#define UD3D_TILE_M 128     /* output rows per CTA tile of the gather-GEMM (tile_mask granularity) */

int ud3d_version(void);
const char* ud3d_last_error(void);
/* number of this library's kernels launched by the process (all host threads) since the last reset */
int64_t ud3d_launch_count(int reset);
/* Per-device context: the only state the library keeps -- per device, the SM count and which function
 * attributes (dynamic shared-memory size, carve-out) its kernels were configured with.  Created lazily for the calling
 * thread's CURRENT device by the first entry point that needs it, mutex-guarded (entry points may be called from
 * several host threads; buffers, workspaces and streams are the caller's).  The handle is only for inspection. */
typedef struct ud3d_ctx ud3d_ctx;
const ud3d_ctx* ud3d_ctx_current(void);          /* NULL (+ ud3d_last_error) when no CUDA device is current */
int ud3d_ctx_device(const ud3d_ctx* ctx);
int ud3d_ctx_sm_count(const ud3d_ctx* ctx);

/* ------------------------------------------------------------------ voxelisation
 * reference: unidet3d/unidet3d.py:136-176 (UniDet3D.collate -> ME.utils.batch_sparse_collate,
 * ME.TensorField(...).sparse(), field.inverse_mapping).
 *
 * ud3d_point_coords: per scene b (points [scene_offsets[b], scene_offsets[b+1])):
 *   coords[p] = (b, floor((xyz - min_b) / voxel_size))  int32, IEEE fp32 division
 *   feats[p]  = (rgb, xyz - mean_b)                      fp32 [n,6]
 *   max_coord[3] = max over the batch of coords[:,1:4]   (spatial_shape = clip(max+1, 128))
 * points fp32 [n,6] (x,y,z,r,g,b); scene_offsets int32 [B+1]; stats fp32 [B,6] out (min, mean).
 * ws: >= ud3d_point_coords_workspace_bytes(B). */
size_t ud3d_point_coords_workspace_bytes(int B);
int ud3d_point_coords(const float* points, int n, const int32_t* scene_offsets, int B, float voxel_size,
                      int32_t* coords, float* feats, float* stats, int32_t* max_coord,
                      void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ occupancy grid ("hash-grid")
 * A collision-free spatial hash of the active voxel set: one bit per cell of the dense
 * (B, X, Y, Z) box (z fastest, Z padded to 32) + an exclusive popcount prefix per 32-bit
 * word.  rank(b,x,y,z) = prefix[word] + popc(word & below(z)) is the row of the voxel in
 * canonical ascending (b,x,y,z) order -- no probing, no sort, deterministic.
 * It replaces MinkowskiEngine's coordinate hash map (unidet3d.py:171-174) and spconv's
 * indice-pair hash (every SubMConv3d/SparseConv3d with an indice_key: spconv_unet.py:43-56,
 * 148-154; unidet3d.py:96-103).
 * dims_host = {B, X, Y, Z} (host array). */
size_t ud3d_grid_workspace_bytes(const int32_t dims_host[4]);
/* set bits for coords [n,4] (rows with coords[i][0] < 0 are skipped), build prefix;
 * *n_unique (device int32) = number of distinct cells. */
int ud3d_grid_build(const int32_t* coords, int n, const int32_t dims_host[4], void* ws, size_t ws_bytes,
                    int32_t* n_unique, void* stream);
/* Occupancy grid of the NEXT (k=2, s=2 down-sampled) level straight from the finer level's bitmap: cell c -> c / 2,
 * cells whose coarse coordinate falls outside `dims_host` are dropped (spconv's odd-extent rule, see
 * ud3d_down2_parents).  Same result as ud3d_grid_build on ud3d_down_ancestors(coords, 1), without touching the points
 * or voxels again (a pass over the fine bitmap words). */
int ud3d_grid_build_coarser(const int32_t fine_dims_host[4], const void* fine_ws, const int32_t dims_host[4], void* ws,
                            size_t ws_bytes, int32_t* n_unique, void* stream);
/* rank_out[i] = canonical row of coords[i] or -1 (absent / outside the grid / coords[i][0] < 0) */
int ud3d_grid_rank(const int32_t* coords, int n, const int32_t dims_host[4], const void* ws,
                   int32_t* rank_out, void* stream);
/* coords_out [n_unique,4]: the distinct cells in canonical order */
int ud3d_grid_coords(const int32_t dims_host[4], const void* ws, int32_t* coords_out, int n_unique,
                     void* stream);

/* per-voxel unweighted mean of point features (ME TensorField UNWEIGHTED_AVERAGE, unidet3d.py:171-173)
 * feats_pts [n,C], rank [n] (from ud3d_grid_rank), out [n_vox,C]; ws >= n_vox*4 bytes. */
int ud3d_voxel_mean(const float* feats_pts, const int32_t* rank, int n, int C, int n_vox, float* out,
                    void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ rulebooks (integer, bit-exact)
 * SubMConv3d(k=3,pad=1) neighbour table (spconv_unet.py:43-56 `subm{l}`, unidet3d.py:96-103):
 *   table[k][i] = row of voxel coords[i] + (kx-1,ky-1,kz-1), k = kx*9+ky*3+kz, or -1.
 * The grid must have been built from `coords`.  row_of_rank: NULL when coords are already in
 * canonical order (row == rank), else an int32 [n] scratch the call fills (rank -> input row).
 * tile_mask [ceil(n/UD3D_TILE_M)] (optional, may be NULL): bit k set iff some row of the tile has
 * an input for offset k. */
int ud3d_rulebook_subm3(const int32_t* coords, int n, const int32_t dims_host[4], const void* ws,
                        int32_t* row_of_rank, int32_t* table, uint32_t* tile_mask, void* stream);
/* Tile order for the SubM3 convolutions of one level (no reference counterpart: spconv's implicit-GEMM kernels carry
 * their own per-tile masks, `pair_mask_fwd_splits`, produced inside get_indice_pairs_implicit_gemm).  ud3d_gemm_fwd skips
 * kernel offset k for a 128-row tile only when no row of the tile has that neighbour; rows are regrouped by a 16-bit
 * key of their rarest neighbours so that tiles hold rows with similar neighbourhoods (26 -> 17..20 active offsets per
 * tile on indoor scans):
 *   perm [n]              a permutation of 0..n-1 (position -> row); order inside a key bucket is unspecified
 *   table_p [27, n]       table_p[k][i] = table[k][perm[i]]
 *   tile_mask_p [tiles]   tile mask of table_p
 * Pass table_p / tile_mask_p / perm as table / tile_mask / row_perm of ud3d_gemm_args: results are bit-identical to the
 * unpermuted call.  ws: ud3d_subm3_tile_order_workspace_bytes(n), 16-byte aligned. */
size_t ud3d_subm3_tile_order_workspace_bytes(int n);
int ud3d_subm3_tile_order(const int32_t* table, int n, int32_t* perm, int32_t* table_p, uint32_t* tile_mask_p,
                          void* ws, size_t ws_bytes, void* stream);

/* SparseConv3d(k=2,s=2) (spconv_unet.py:148-154 `spconv{l}`), phase 1:
 *   parents[i] = (b, x/2, y/2, z/2), or b=-1 when x/2 >= out_shape (odd extent, last index dropped),
 *   out_shape = (in_shape-2)/2+1.  Then the caller builds a grid on `parents`
 *   (ud3d_grid_build) and reads back n_coarse. */
int ud3d_down2_parents(const int32_t* coords, int n, const int32_t in_shape_host[3], int32_t* parents,
                       void* stream);
/* ancestors[i] = coordinates of voxel i after `levels` successive k=2,s=2 down-samplings (b = -1 when the voxel is
 * dropped at any of them).  The coarse grids of every level can thus be built from the finest coordinates (or even the
 * per-point coordinates) back to back, and all voxel counts read with a single host synchronisation. */
int ud3d_down_ancestors(const int32_t* coords, int n, const int32_t in_shape_host[3], int levels, int32_t* ancestors,
                        void* stream);
/* phase 2: child[s][p] = fine row feeding coarse row p through slot s = (x&1)*4+(y&1)*2+(z&1);
 * up[s][i] = coarse row of fine row i (slot s), both -1-filled elsewhere.  The pair list is shared by
 * SparseInverseConv3d (spconv_unet.py:178-183).  coarse_ws = grid built on `parents`.
 * child_mask/up_mask: optional tile masks like ud3d_rulebook_subm3. */
int ud3d_rulebook_down2(const int32_t* coords, const int32_t* parents, int n_fine, int n_coarse,
                        const int32_t coarse_dims_host[4], const void* coarse_ws,
                        int32_t* child, int32_t* up, uint32_t* child_mask, uint32_t* up_mask, void* stream);

/* ------------------------------------------------------------------ gather-GEMM (sparse conv + linear)
 * reference: every spconv.SubMConv3d / SparseConv3d / SparseInverseConv3d forward
 * (spconv_unet.py:37-72,148-191; unidet3d.py:96-103) with the preceding BatchNorm(eval)+ReLU
 * (spconv_unet.py:42,49,147,177) folded into the operand load and the residual add / concat
 * (spconv_unet.py:88-89,229-230) folded into the epilogue; also every torch.nn.Linear of the
 * encoder (encoder.py:19-22,55-61,138-163).
 *
 *   out[o, :] = act( sum_k  pre(in[table[k][o], :]) @ W_k  + bias ) + residual[o, :]
 *   pre(x) = relu?( x * in_scale + in_shift )   (per input channel; skipped for missing rows)
 *
 * W is given in the reference parameter layout [C_out, K, C_in] (spconv [C_out,k0,k1,k2,C_in]
 * flattened; nn.Linear [C_out, C_in] with K = 1) and packed once by ud3d_gemm_pack_weight into
 * the kernel's shared-memory image (bf16 hi/lo split, 128B-swizzled K-major tiles).
 * Numerics: each fp32 operand is split into bf16 hi+lo, three tcgen05 MMAs (hi*hi + lo*hi + hi*lo)
 * accumulate in fp32 in TMEM: ~2e-5 relative to an fp32 reference over the whole backbone. */
typedef struct {
  const float* in; int32_t ld_in; int32_t c_in;
  const int32_t* table;        /* [K, n_out]; NULL => identity gather (requires K == 1) */
  const uint32_t* tile_mask;   /* optional [ceil(n_out/128)] */
  int32_t K; int32_t n_out;
  const void* w_packed;        /* from ud3d_gemm_pack_weight */
  float* out; int32_t ld_out; int32_t c_out;   /* always a valid buffer (scratch when no_raw) */
  const float* in_scale; const float* in_shift; int32_t in_relu;  /* in_scale NULL => no affine */
  const float* bias;           /* [c_out] or NULL */
  int32_t act;                 /* 0 none, 1 relu, 2 gelu(erf) */
  const float* residual; int32_t ld_res;
  /* operand-form feature maps (see ud3d_act_split):
   * in_split != 0: `in` holds the split-bf16 operand form of the ALREADY ACTIVATED input (ld_in, c_in in
   * channels = 4-byte units as for fp32; c_in % 32 == 0); it is gathered with cp.async straight into the
   * swizzled tile, no per-use conversion.  in_scale must be NULL.
   * in_split == 2: the same data in the INTERLEAVED operand form (per 32-channel chunk the 16-byte pieces in the order
   * hi(ch 0-7), lo(ch 0-7), hi(ch 8-15), lo(ch 8-15), ...): the input form of the experimental kernel that gathers
   * the A operand through registers into TMEM (needs w_packed_ts; not used by the product path).
   * out_act[i] != NULL: additionally store relu(result * act_scale[i] + act_shift[i]) in operand form
   * (the consumer conv's folded BatchNorm+ReLU, applied ONCE per element instead of once per use);
   * c_out % 32 == 0.  no_raw != 0: skip the fp32 store to `out` (still used as scratch by split-K). */
  int32_t in_split; int32_t no_raw;
  float* out_act[2]; int32_t ld_act[2]; const float* act_scale[2]; const float* act_shift[2];
  int32_t act_norelu;          /* bit i set: out_act[i] = split(result * scale + shift) without the ReLU;
                                  act_scale[i] == NULL means identity (scale 1, shift 0) */
  const int32_t* row_perm;     /* optional [n_out] (requires table): position i of the table / tile_mask describes output
                                  row row_perm[i], i.e. out / out_act / residual are addressed at row_perm[i]
                                  (see ud3d_subm3_tile_order); NULL = identity */
  const void* w_packed_ts;     /* optional, from ud3d_gemm_pack_weight_ts: the same weight in the K order of the kernel
                                  variant that gathers the A operand through registers into TMEM (operand-form inputs
                                  of launches that fill the GPU); NULL => the shared-memory-operand kernel is used */
} ud3d_gemm_args;
size_t ud3d_gemm_packed_weight_bytes(int K, int c_in, int c_out);
int ud3d_gemm_pack_weight(const float* w, int K, int c_in, int c_out, void* packed, void* stream);
/* second packing of the same weight (same size), see ud3d_gemm_args.w_packed_ts */
int ud3d_gemm_pack_weight_ts(const float* w, int K, int c_in, int c_out, void* packed, void* stream);
int ud3d_gemm_fwd(const ud3d_gemm_args* args, void* stream);
/* operand form of a feature map: per row, per 32-channel chunk, 64 bytes of bf16 "hi" followed by 64 bytes of
 * bf16 "lo" (x ~= hi + lo), i.e. the same 4 bytes per element as fp32 and exactly the 128-byte row of the
 * kernel's shared-memory tile.  out_split[r, c] = split(relu?(raw[r, c] * scale[c] + shift[c])); scale may be
 * NULL (identity).  ld in 4-byte units; when c is not a multiple of 32 the last chunk is zero-padded
 * (ld_out >= roundup(c, 32)). */
int ud3d_act_split(const float* raw, int ld_raw, int n, int c, const float* scale, const float* shift, int relu,
                   float* out_split, int ld_out, void* stream);
/* same contract, plain fp32 CUDA-core kernel on the unpacked weight w [C_out,K,C_in]
 * (diagnostic cross-check of the tensor-core path; not used by the product path) */
int ud3d_gemm_fwd_simt(const ud3d_gemm_args* args, const float* w, void* stream);

/* ------------------------------------------------------------------ training-pipeline transforms in front of the path
 * ElasticTransfrom (unidet3d/transforms_3d.py:12-83): `elastic_coords = elastic(elastic(points/voxel_size, gran0, mag0),
 * gran1, mag1)`, elastic(x) = x + trilinear(blurred noise grids)(x) * mag.  The caller draws the three float32 noise
 * grids [3, X, Y, Z] (dims = |x|.max(0).astype(int32) // gran + 3, reference order of np.random calls) and keeps the
 * coordinates in double like the reference (scipy returns float64):
 *   ud3d_points_to_voxel_units : out[i, k] = double(float32(points[i, k]) / float32(voxel_size))
 *   ud3d_elastic_blur          : the six 3-tap box blurs (scipy.ndimage.convolve, mode='constant'), in place
 *   ud3d_elastic_apply         : out = x + interp(x) * mag  (RegularGridInterpolator: linear, fill_value 0 outside) */
size_t ud3d_elastic_workspace_bytes(const int32_t dims_host[3]);
int ud3d_elastic_blur(float* noise, const int32_t dims_host[3], void* ws, size_t ws_bytes, void* stream);
int ud3d_points_to_voxel_units(const float* points, int ld, int n, float voxel_size, double* out, void* stream);
int ud3d_elastic_apply(const double* x, int n, const float* noise, const int32_t dims_host[3], double gran, double mag,
                       double* out, void* stream);
/* Voxel coordinates of elastic coordinates (unidet3d.py:162-166): coords[p] = (b, floor(el[p] - min over scene b of el)),
 * computed in double like the reference; max_coord int32 [3] = per-axis maximum over the batch (spatial extents). */
size_t ud3d_elastic_voxel_coords_workspace_bytes(int B);
int ud3d_elastic_voxel_coords(const double* elastic, int n, const int32_t* scene_offsets, int B, int32_t* coords,
                              int32_t* max_coord, void* ws, size_t ws_bytes, void* stream);
/* PointSample_ (unidet3d/transforms_3d.py:233-295) after the row gather: ids -> dense ranks of the values present
 * (np.unique(ids, return_inverse=True)[1]); negative ids (the -1 "no instance" label) stay -1, which is what the
 * reference's `mapping` does for pts_instance_mask.  ids in [-inf, max_id]; n_unique = number of distinct ids >= 0. */
size_t ud3d_compact_ids_workspace_bytes(int64_t max_id);
int ud3d_compact_ids(const int64_t* ids, int n, int64_t max_id, int64_t* out, int32_t* n_unique, void* ws, size_t ws_bytes,
                     void* stream);

/* ------------------------------------------------------------------ evaluator (the step after the path)
 * reference: unidet3d/indoor_eval.py:56-202 (eval_det_cls + average_precision 'area' + eval_map_recall) with the 3-D IoU of
 * mmdet3d's BaseInstance3DBoxes.overlaps.  All classes and IoU thresholds of a result set in one call:
 *   detections: det_boxes [D,7] (cx,cy,cz,dx,dy,dz,yaw; gravity centre), det_labels / det_img int32 [D];
 *     order int32 [D] = detection indices sorted by (label ascending, score descending); class_offsets int32 [n_cls+1]
 *     = segment of each class inside `order`;
 *   ground truth: gt_boxes [G,7], gt_labels int32 [G], grouped by image: gt_img_offsets int32 [n_img+1];
 *   thr_host: n_thr <= 16 IoU thresholds (host array).
 * Outputs (device): ap float32 [n_cls, n_thr] (nan for a class with detections but no ground truth, like the reference),
 * rec double [n_cls, n_thr] = final recall tp / npos, npos int32 [n_cls].  A class without detections yields 0 / 0. */
size_t ud3d_eval_workspace_bytes(int D, int G, int n_thr);
int ud3d_eval_detections(const float* det_boxes, const int32_t* det_labels, const int32_t* det_img, int D,
                         const int32_t* order, const int32_t* class_offsets, int n_cls, const float* gt_boxes,
                         const int32_t* gt_labels, const int32_t* gt_img_offsets, int G, int n_img, const float* thr_host,
                         int n_thr, float* ap, double* rec, int32_t* npos, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ stage plan: the whole U-Net in one call
 * reference: SpConvUNet.forward, unidet3d/spconv_unet.py:117-240 (eval mode; channel counts multiples of 32).
 * A plan holds, per level, the packed weights (ud3d_gemm_pack_weight) and the folded eval-mode BatchNorms
 * (scale = gamma / sqrt(var + eps), shift = beta - mean * scale) of the parameter tree
 *   blocks.block{i}.conv_branch.{0,2,3,5}, conv.{0,2}, deconv.{0,2}, blocks_tail.block{i}.{i_branch.0, conv_branch.*}
 * and ud3d_unet_forward issues every convolution of the recursion through ud3d_gemm_fwd with the fusion described there
 * (operand-form maps under the consumer's BatchNorm, residuals in the epilogue, concat never materialised).  It exists to
 * take the host off the critical path: ~50 launches marshalled in C instead of through the caller's language runtime.
 *   x_raw [n_0, c_0] fp32 input features, x_act the same under level[0].blocks[0].bn0 in operand form (ud3d_act_split or a
 *   producer epilogue); out_raw [n_0, c_0]; level_out: NULL or n_levels pointers ([l] NULL or [n_l, c_l]: the fp32 output
 *   of level l, the `previous_outputs` of return_blocks=True; [0] ignored).  ws: ud3d_unet_workspace_bytes, 256-byte aligned. */
#define UD3D_UNET_MAX_LEVELS 8
#define UD3D_UNET_MAX_REPS 4
typedef struct {
  const void* w0; const void* w1;      /* SubM3 c -> c (tail block 0: w0 is 2c -> c) */
  const void* wi;                      /* SubM1 i_branch (tail block 0 only), else NULL */
  const float* bn0_scale; const float* bn0_shift;   /* conv_branch.0 (tail block 0: 2c channels) */
  const float* bn1_scale; const float* bn1_shift;   /* conv_branch.3 */
} ud3d_unet_block;
typedef struct {
  int32_t c;
  ud3d_unet_block blocks[UD3D_UNET_MAX_REPS];
  ud3d_unet_block tail[UD3D_UNET_MAX_REPS];         /* unused at the deepest level */
  const void* down_w; const void* up_w;             /* conv.2 (k2 s2, c -> c_next), deconv.2 (inverse, c_next -> c) */
  const float* down_scale; const float* down_shift; /* conv.0   [c] */
  const float* up_scale; const float* up_shift;     /* deconv.0 [c_next] */
} ud3d_unet_level;
typedef struct {
  int32_t n_levels; int32_t block_reps;
  ud3d_unet_level level[UD3D_UNET_MAX_LEVELS];
} ud3d_unet_plan;
typedef struct {                                    /* rulebooks of one level (ud3d_rulebook_subm3 / _down2 / tile order) */
  int32_t n;
  const int32_t* subm; const uint32_t* subm_mask; const int32_t* row_perm;   /* row_perm NULL = canonical order */
  const int32_t* child; const uint32_t* child_mask;                           /* [8, n_next]; NULL at the deepest level */
  const int32_t* up; const uint32_t* up_mask;                                 /* [8, n] */
} ud3d_unet_tables;
size_t ud3d_unet_workspace_bytes(const ud3d_unet_plan* plan, const ud3d_unet_tables* levels);
int ud3d_unet_forward(const ud3d_unet_plan* plan, const ud3d_unet_tables* levels, const float* x_raw, const float* x_act,
                      float* out_raw, float* const* level_out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ stage plan: the encoder in one call
 * reference: UniDet3DEncoder.forward, unidet3d/encoder.py:203-239 with the last head (:165-201), eval mode, no auxiliary
 * heads.  x fp32 [n, in_channels] packed over the B scenes (cu_seqlens int32 [B+1], max_T = longest scene).
 * Outputs: logits [n, n_union] (union of the datasets' classes + no_obj; the per-dataset column gather stays with the
 * caller), raw_boxes [n, 8] (PredBBox linear output, before ud3d_bbox_decode), h_out NULL or [n, d_model] (the queries
 * after the last layer).  Weights packed by ud3d_gemm_pack_weight (K = 1).  activation: 1 relu, 2 gelu. */
#define UD3D_ENCODER_MAX_LAYERS 12
typedef struct { const void* w; const float* bias; } ud3d_linear;
typedef struct {
  ud3d_linear qkv, out, f1, f2;                    /* attn.in_proj, attn.out_proj, ffn.net.0, ffn.net.3 */
  const float* n1_gamma; const float* n1_beta; float n1_eps;
  const float* n2_gamma; const float* n2_beta; float n2_eps;
} ud3d_encoder_layer;
typedef struct {
  int32_t num_layers, in_channels, d_model, num_heads, hidden, n_union, activation;
  ud3d_linear ip0, ip2;                            /* input_proj.0, input_proj.2 */
  ud3d_encoder_layer layer[UD3D_ENCODER_MAX_LAYERS];
  const float* on_gamma; const float* on_beta; float on_eps;   /* out_norm */
  ud3d_linear c0, c2, bb;                          /* outs_cls.0, outs_cls.2, out_bboxes.linear */
} ud3d_encoder_plan;
size_t ud3d_encoder_workspace_bytes(const ud3d_encoder_plan* plan, int n);
int ud3d_encoder_forward(const ud3d_encoder_plan* plan, const float* x, int n, const int32_t* cu_seqlens, int B, int max_T,
                         float* logits, float* raw_boxes, float* h_out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ training side of the backbone
 * Train-mode (Sync)BatchNorm of the reference (spconv_unet.py:119-124, unidet3d.py:104-107; torch.nn.SyncBatchNorm,
 * eps 1e-4, momentum 0.1): statistics over ALL active voxels of the (global) batch.
 *   ud3d_bn_batch_sums : sums[0..C) = sum_r x[r,c], sums[C..2C) = sum_r x[r,c]^2 (fp64, deterministic order).
 *     For SyncBatchNorm the caller all-reduces `sums` (and the row count) across ranks before the fold.
 *   ud3d_bn_train_fold : mean / biased variance -> scale = gamma / sqrt(var + eps), shift = beta - mean * scale (the
 *     form ud3d_gemm_args.in_scale / in_shift and ud3d_act_split consume), running_mean / running_var updated in place
 *     (unbiased variance, like torch); save_mean / save_invstd optional (for the backward pass).
 *   count_dev (both fold and backward_apply; may be NULL): the row count as a DEVICE double -- under SyncBatchNorm it comes out
 *     of the same all-reduce as the sums, and reading it on the host would cost one synchronisation per BatchNorm. */
size_t ud3d_bn_batch_sums_workspace_bytes(int n, int C);
int ud3d_bn_batch_sums(const float* x, int ld, int n, int C, double* sums, void* ws, size_t ws_bytes, void* stream);
int ud3d_bn_train_fold(const double* sums, double count, int C, const float* gamma, const float* beta, float eps,
                       float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                       float* save_mean, float* save_invstd, const double* count_dev, void* stream);
/* Backward of the fused train-mode BatchNorm + ReLU in front of a conv (a = relu(x * scale + shift)), given dA = the conv's
 * input gradient:  g = dA * [a > 0];  sums[0..C) = sum_r g (= dbeta), sums[C..2C) = sum_r g * xhat (= dgamma), fp64, fixed
 * order (all-reduced across ranks for SyncBatchNorm);  dx (+)= scale * (g - dbeta / count - xhat * dgamma / count).
 * mean / invstd: the save_mean / save_invstd of ud3d_bn_train_fold.  ws: ud3d_bn_batch_sums_workspace_bytes(n, C). */
int ud3d_bn_backward_sums(const float* x, int ld_x, const float* da, int ld_da, int n, int C, const float* scale,
                          const float* shift, const float* mean, const float* invstd, int relu, double* sums, void* ws,
                          size_t ws_bytes, void* stream);
int ud3d_bn_backward_apply(const float* x, int ld_x, const float* da, int ld_da, int n, int C, const float* scale,
                           const float* shift, const float* mean, const float* invstd, int relu, const double* sums,
                           double count, float* dx, int ld_dx, int accumulate, const double* count_dev, void* stream);
/* out = relu?(x * scale + shift) as an fp32 map: the X operand of ud3d_conv_wgrad (recomputed, not stored, in forward) */
int ud3d_bn_relu_apply(const float* x, int ld_x, int n, int C, const float* scale, const float* shift, int relu, float* out,
                       int ld_out, void* stream);
/* Backward of the attention core (encoder.py:36-37; head_dim 32, no mask): qkv fp32 [T_total, 3 d] (q | k | v), out / d_out
 * [T_total, d] = the forward result and its gradient -> dqkv [T_total, 3 d].  fp32 on the CUDA cores, deterministic.
 * ws: ud3d_attention_bwd_workspace_bytes (row log-sum-exp and D = rowsum(dO o O)). */
size_t ud3d_attention_bwd_workspace_bytes(int total_T, int num_heads);
int ud3d_attention_bwd(const float* qkv, const int32_t* cu_seqlens, int B, int total_T, int num_heads, const float* out,
                       const float* d_out, float* dqkv, void* ws, size_t ws_bytes, void* stream);
/* Same contract and results (up to fp32 summation order); one thread per query / key with register-resident rows and
 * accumulators, the other side staged through shared memory: the variant the training step uses. */
int ud3d_attention_bwd_reg(const float* qkv, const int32_t* cu_seqlens, int B, int total_T, int num_heads, const float* out,
                       const float* d_out, float* dqkv, void* ws, size_t ws_bytes, void* stream);
/* LayerNorm backward (encoder.py:38-39,77-78,189): dx [rows, C]; dgamma_dbeta fp64 [2, C] = (sum_r dy * xhat, sum_r dy).
 * Statistics are recomputed from x (= the LayerNorm's input, residual already added). */
size_t ud3d_layernorm_backward_workspace_bytes(int rows, int C);
int ud3d_layernorm_backward(const float* x, const float* dy, const float* gamma, int rows, int C, float eps, float* dx,
                            double* dgamma_dbeta, void* ws, size_t ws_bytes, void* stream);
/* dx = dy * act'(pre) element-wise; act 1 = relu, 2 = gelu(erf) (the activation codes of ud3d_gemm_args.act) */
int ud3d_activation_backward(const float* pre, const float* dy, long long total, int act, float* dx, void* stream);
/* out = act(pre) as its own pass: the training forward keeps the pre-activation values for the backward pass (the
 * inference path applies the activation in the GEMM epilogue) */
int ud3d_activation_forward(const float* pre, long long total, int act, float* out, void* stream);
/* Backward of ud3d_segmented_mean without its affine (apply ud3d_bn_backward_* on the result for the fused output
 * BatchNorm): d_src[gather ? gather[p] : p, :] += d_pooled[seg[p], :] / count[seg[p]];  d_src [n_rows, C] is overwritten.
 * Deterministic (64-bit fixed-point atomics, 2^-32).  ws 8-byte aligned. */
size_t ud3d_segmented_mean_backward_workspace_bytes(int n_rows, int n_seg, int C);
int ud3d_segmented_mean_backward(const float* d_pooled, int C, const int32_t* gather, const int64_t* seg, int n, int n_seg,
                                 int n_rows, float* d_src, void* ws, size_t ws_bytes, void* stream);
/* Weight gradient of a sparse convolution / linear layer (autograd of spconv's conv forward, spconv_unet.py:37-72):
 *   dw[co][k][ci] (+)= sum_o dy[o][co] * x[table[k][o]][ci]      (layout of the reference parameter [C_out, K, C_in])
 * x = the conv's input as it entered the contraction (after its BatchNorm + ReLU), fp32; deterministic.  The input
 * gradient needs no kernel of its own: it is ud3d_gemm_fwd on dy with the transposed weight and the transposed
 * rulebook (SubM3: the same table with the kernel offsets reversed; k2s2 <-> its inverse conv's table). */
size_t ud3d_conv_wgrad_workspace_bytes(int n_out, int K, int c_in, int c_out);
int ud3d_conv_wgrad(const float* x, int ld_x, int c_in, const float* dy, int ld_dy, int c_out, const int32_t* table,
                    int n_out, int K, float* dw, int accumulate, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ superpoint pooling
 * reference: torch_scatter.scatter_mean at unidet3d.py:130 (features, with the output
 * BatchNorm+ReLU of unidet3d.py:104-111,129 and the x.features[inverse_mapping] gather fused)
 * and unidet3d.py:446-447 (superpoint centres).
 *   out[s,:] = mean_{p: seg[p]==s} pre(src[gather ? gather[p] : p, :]),  empty s -> 0.
 * seg int64 [n] (reference loader dtype), out [n_seg,C]; ws >= ud3d_segmented_mean_workspace_bytes (8-byte aligned).
 * Deterministic (bit-identical from run to run): partial sums are accumulated as 64-bit fixed point (2^-24), whose
 * addition is associative; valid for |sum of a segment| < 5e11. */
size_t ud3d_segmented_mean_workspace_bytes(int n_seg, int C);
int ud3d_segmented_mean(const float* src, int ld_src, int C, const int32_t* gather, const int64_t* seg,
                        int n, int n_seg, const float* scale, const float* shift, int relu,
                        float* out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ encoder pieces
 * LayerNorm(eps) of (x + residual) per row (encoder.py:38-39,77-78,189); C % 32 == 0, C <= 1024 */
int ud3d_layernorm(const float* x, const float* residual, const float* gamma, const float* beta,
                   float* out, int rows, int C, float eps, void* stream);
/* same, additionally (out_split != NULL) storing the result in operand form (ud3d_act_split layout) for the
 * consumer GEMMs; out may be NULL when only the operand form is needed.  C in {128, 256}. */
int ud3d_layernorm_split(const float* x, const float* residual, const float* gamma, const float* beta,
                         float* out, float* out_split, int rows, int C, float eps, void* stream);
/* multi-head self-attention core of nn.MultiheadAttention (encoder.py:37) on packed projections:
 * qkv [T_total, 3*d] (q|k|v, heads contiguous, head_dim = 32), cu_seqlens int32 [B+1] scene
 * boundaries (no cross-scene attention, no mask/padding), out [T_total, d] = softmax(QK^T/sqrt(32))V */
int ud3d_attention_fwd(const float* qkv, const int32_t* cu_seqlens, int B, int max_T, int num_heads,
                       float* out, void* stream);
/* same with the output stored in operand form (one head = one 32-channel chunk) for the out-projection GEMM */
int ud3d_attention_fwd_split(const float* qkv, const int32_t* cu_seqlens, int B, int max_T, int num_heads,
                             float* out_split, void* stream);
/* both the packed q|k|v projection and the output in operand form (the QKV GEMM writes it with out_act):
 * K/V tiles stream through cp.async as raw 128-byte rows, no per-tile conversion */
int ud3d_attention_fwd_opform(const float* qkv_split, const int32_t* cu_seqlens, int B, int max_T, int num_heads,
                              float* out_split, void* stream);
/* same contract on the 5th-gen tensor cores (csrc/attention_tc.cu): S = Q K^T and O' = P V' accumulate in TMEM, P goes
 * softmax -> TMEM -> second GEMM without touching shared memory, row sums on the tensor core, Q / K / V tiles by TMA
 * tensor-map copies.  total_T = number of token rows of qkv_split (= cu_seqlens[B]), needed for the tensor map. */
int ud3d_attention_fwd_tc(const float* qkv_split, const int32_t* cu_seqlens, int B, int max_T, int total_T, int num_heads,
                          float* out_split, void* stream);
/* PredBBox exp + _bbox_pred_to_bbox (encoder.py:109-111,241-283): raw [T,ld_raw>=8], centres [T,3]
 * -> out [T, with_angle ? 7 : 6] */
int ud3d_bbox_decode(const float* raw, int ld_raw, const float* centers, int T, int with_angle,
                     float* out, void* stream);
/* out[t, j] = src[t, cols[j]]  (per-dataset class column gather, encoder.py:191-194) */
int ud3d_gather_columns(const float* src, int ld_src, const int32_t* cols, int n_cols, int T,
                        float* out, void* stream);

/* ------------------------------------------------------------------ post-processing
 * softmax over classes, drop the last (no_obj) column, top-k over the flattened [T*C] scores,
 * sorted descending (unidet3d.py:504-515).  logits [T, C+1]; k <= 1024 and k <= T*C.
 * scores_out[k], labels_out[k] (= idx % C), query_out[k] (= idx / C); ws >= T*C*4 bytes. */
int ud3d_topk_scores(const float* logits, int T, int C_plus1, int k, float* scores_out,
                     int32_t* labels_out, int32_t* query_out, void* ws, size_t ws_bytes, void* stream);
/* multi-class 3D NMS (unidet3d.py:595-650): boxes [n,box_dim] (6 or 7), n <= 1024, already in
 * descending score order.  mode 0: mmcv nms3d (rotated BEV IoU), 1: mmcv nms3d_normal (axis-aligned
 * BEV), 2: mmdet3d aligned_3d_nms (3-D IoU).  keep_out[n]: kept input indices, ascending class then
 * descending score; *n_keep device int32.  ws >= ud3d_nms_workspace_bytes(n). */
size_t ud3d_nms_workspace_bytes(int n);
int ud3d_nms_multiclass(const float* boxes, int box_dim, const float* scores, const int32_t* labels, int n,
                        int mode, float iou_thr, float score_thr, int32_t* keep_out, int32_t* n_keep,
                        void* ws, size_t ws_bytes, void* stream);
/* trim_bboxes_by_superpoints (unidet3d.py:540-593, get_face_distances :652-677):
 * points [n_pts,ld_pts>=3], sp int64 [n_pts], boxes [m,box_dim]; box_index (optional int32 [m]) selects
 * rows of `boxes`.  out [m,6] = (centre, size) of the tight AABB of the voted points.
 * m_dev (optional device int32): only the first min(m, *m_dev) boxes are processed (lets the NMS
 * count stay on the device).  ws >= ud3d_trim_workspace_bytes(n_sp, n_pts, m). */
size_t ud3d_trim_workspace_bytes(int n_sp, int n_pts, int m);
int ud3d_trim_boxes(const float* points, int ld_pts, const int64_t* sp, int n_pts, int n_sp,
                    const float* boxes, int box_dim, const int32_t* box_index, int m, const int32_t* m_dev,
                    float low_thr, float up_thr, float* out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ fused per-scene post-processing
 * predict_by_feat for one scene (unidet3d.py:475-538) as ONE host call: softmax/top-k -> candidate box gather ->
 * multi-class NMS -> (optional) superpoint trimming, ~13 kernel launches issued back to back from C++ on `stream`
 * (lets several scenes run concurrently on different streams without per-launch interpreter overhead).
 * Outputs (all device, caller-owned): scores[k], labels[k], cand[k,box_dim] (candidate boxes in score order),
 * keep[k] + n_keep[1] (kept candidate indices: ascending class, descending score), trimmed[k,6] (row j = trimmed box
 * of candidate keep[j]; only when use_trim). */
typedef struct {
  const float* logits; int32_t ld_logits; int32_t T; int32_t C1;
  const float* boxes; int32_t box_dim;
  int32_t k; int32_t nms_mode; float iou_thr; float score_thr;
  int32_t use_trim; const float* points; int32_t ld_pts; const int64_t* sp; int32_t n_pts; int32_t n_sp;
  float low_thr; float up_thr;
  float* scores; int32_t* labels; float* cand; int32_t* keep; int32_t* n_keep; float* trimmed;
} ud3d_post_args;
size_t ud3d_postprocess_workspace_bytes(const ud3d_post_args* args);
int ud3d_postprocess_scene(const ud3d_post_args* args, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ training-side targets, matcher, loss values
 * (forward values here, gradients in ud3d_criterion_layer_grad below)
 *
 * get_bboxes_by_masks (unidet3d.py:220-275): inst int64 [n] (-1 = no instance), points [n, >=3] ->
 * out [n_inst, 6] = (centre, size) of the tight AABB of every instance's points.  ws >= n_inst*6*4 bytes. */
int ud3d_boxes_by_instance(const float* points, int ld_pts, const int64_t* inst, int n, int n_inst, float* out,
                           void* ws, size_t ws_bytes, void* stream);
/* get_targets (unidet3d.py:371-409): superpoint centres [S,3], gt_boxes [G, box_dim] (gravity centre first) ->
 * masks uint8 [G, S]: every box keeps the centres strictly nearer than its (topk+1)-th nearest, every centre goes to
 * the nearest box that kept it.  ws >= G*4 bytes. */
int ud3d_targets_by_distance(const float* centers, int S, const float* gt_boxes, int box_dim, int G, int topk,
                             uint8_t* masks, void* ws, size_t ws_bytes, void* stream);
/* One (decoder layer, scene) of UniDet3DCriterion.get_layer_loss (criterion.py:44-142) with its UniMatcher
 * (criterion.py:286-320; costs QueryClassificationCost * w_cls + BboxCostJointTraining * w_box, criterion.py:200-284):
 *   cost[q,g] = -w_cls * softmax(logits[q])[labels[g]] + w_box * DIoU_loss(boxes[q], gt_boxes[g]),  1e8 where
 *               !query_masks[g,q];   thr[g] = (topk+1)-th smallest cost of column g;   match[q,g] = cost < thr[g]
 *   target[q] = labels[largest matched g] or C (no object);  class weights 1 (objects) / non_object_weight
 *   sums[0] = sum_q w[target] * -log_softmax(logits[q])[target],  sums[1] = sum_q w[target]   (weighted CE = s0/s1)
 *   sums[2] = sum over matched pairs of DIoU_loss(boxes[q], gt_boxes[g]),  sums[3] = number of matched pairs
 * box_dim 6: (centre, size), axis-aligned DIoU (axis_aligned_iou_loss.py:14-53; as a matching cost the reference adds
 * the centre-distance penalty of GT 0 to every column, line 51 -- kept); box_dim 7: (x,y,z,w,h,l,alpha), rotated DIoU
 * (rotated_iou_loss.py:14-82, exact BEV intersection).  G == 0: no match, every query is "no object".
 * Requires T >= topk+1 (torch.topk raises otherwise).  Deterministic (fixed-order reductions). */
typedef struct {
  const float* logits; int32_t ld_logits; int32_t T; int32_t C1;     /* [T, C+1] */
  const float* boxes; int32_t box_dim;                               /* [T, box_dim] predicted boxes */
  const float* gt_boxes; const int64_t* gt_labels; int32_t G;        /* [G, box_dim], [G] */
  const uint8_t* query_masks;                                        /* bool [G, T] */
  int32_t topk; float w_cls; float w_box; float non_object_weight;
  uint8_t* match;                                                    /* out [T, G] */
  float* sums;                                                       /* out [4] */
} ud3d_criterion_args;
size_t ud3d_criterion_workspace_bytes(int T, int G);
int ud3d_criterion_layer(const ud3d_criterion_args* args, void* ws, size_t ws_bytes, void* stream);
/* Gradients of the same (decoder layer, scene) -- what torch.autograd computes through criterion.py:86-142 in the
 * reference (F.cross_entropy with class weights; bbox_loss_simple / bbox_loss_rotated of the matched pairs,
 * axis_aligned_iou_loss.py:14-53, rotated_iou_loss.py:14-82 over mmcv's differentiable rotated intersection):
 *   d_logits[q,c] = scales[0] * w[target[q]] / sums[1] * (softmax(logits[q])[c] - [c == target[q]])
 *   d_boxes[q,:]  = scales[1] / sums[3] * sum over matched g of d DIoU_loss(boxes[q], gt_boxes[g]) / d boxes[q]   (0 if no pair)
 * match / sums: the outputs of ud3d_criterion_layer for these inputs.  scales: DEVICE float[2] = d det_loss / d (this
 * scene's weighted CE) and d det_loss / d (this scene's mean box loss) -- loss weights, dataset weight and the means over
 * scenes (criterion.py:111,136-142), kept on the device because the second depends on how many scenes matched at all.
 * The DIoU derivatives are forward-mode (dual numbers through the same polygon clipping that evaluates the loss,
 * csrc/box_loss.cuh).  Deterministic. */
typedef struct {
  const float* logits; int32_t ld_logits; int32_t T; int32_t C1;     /* [T, C+1] */
  const float* boxes; int32_t box_dim;                               /* [T, box_dim] */
  const float* gt_boxes; const int64_t* gt_labels; int32_t G;        /* [G, box_dim], [G] */
  const uint8_t* match;                                              /* [T, G] */
  const float* sums;                                                 /* [4] */
  const float* scales;                                               /* [2] */
  float non_object_weight;
  float* d_logits; int32_t ld_dlogits;                               /* out [T, C+1] */
  float* d_boxes;                                                    /* out [T, box_dim] */
} ud3d_criterion_grad_args;
int ud3d_criterion_layer_grad(const ud3d_criterion_grad_args* args, void* stream);
/* Backward of one scene's head outputs (what torch.autograd does through encoder.py:165-201,241-283): the per-dataset
 * class column gather and PredBBox's exp + _bbox_pred_to_bbox.  raw [T, ld_raw >= 8] (the box Linear's output),
 * d_box [T, with_angle ? 7 : 6] or NULL, d_cls [T, n_cols] or NULL with cols int32 [n_cols] (distinct columns) ->
 * d_raw [T, 8] = J^T d_box,  d_logits [T, n_union] = scatter of d_cls (every element is written: zeros elsewhere). */
int ud3d_head_backward(const float* raw, int ld_raw, const float* d_box, int with_angle, const float* d_cls, const int32_t* cols,
                       int n_cols, int T, float* d_raw, int ld_draw, float* d_logits, int ld_dlogits, int n_union, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIDET3D_B200_H_ */
