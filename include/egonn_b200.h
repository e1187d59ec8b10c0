/*
 * egonn_b200 - C ABI of the B200-native sparse-voxel descriptor-extraction engine.
 *
 * The reference (jac99/Egonn, /root/reference) has no FFI of its own: its forward path calls the
 * third-party MinkowskiEngine Python API.  Each entry point below names the reference call site
 * (file:line under /root/reference) whose work it replaces; MinkowskiEngine's own analogue would be
 * its pybind module MinkowskiEngineBackend._C (quantize_th, CoordinateMapManagerGPU_c10,
 * ConvolutionForwardGPU, ConvolutionTransposeForwardGPU, GlobalPoolingForwardGPU, BroadcastForwardGPU).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.  All data pointers are DEVICE pointers unless
 *     the parameter is documented "host".  The caller owns every buffer it passes; the engine never frees
 *     caller memory.  Engine-owned scratch lives in the egn_ctx and is borrowed until the next call.
 *   - every function returns 0 on success or a negative egn_status; egn_last_error() gives the text.
 *   - work is enqueued on the cudaStream_t passed as `stream` (void* to keep this header CUDA-free).
 *     egn_coords_build / egn_quantize synchronise that stream once, because row counts must reach the host.
 *   - one egn_ctx per (device, stream); contexts are independent (no global mutable state besides the
 *     thread-local error string).
 *   - coordinates are int32 [batch, x, y, z]; spatial range [-2^17, 2^17), batch index < 1023.
 *   - rows of every coordinate map are kept in the engine's canonical order: ascending
 *     (batch, Morton(z,y,x)) - MinkowskiEngine's row order is not a contract.
 *   - there is NO CPU fallback: every entry point needs a CUDA device.
 *   - names vs SURVEY.md 8b.3's minimum export set.  "egn_quantize_cartesian": here egn_quantize, which also does polar.
 *     "egn_build_pyramid" + "egn_build_kernel_map": here ONE call, egn_coords_build, because the kernel maps are derived
 *     top-down from the pyramid.  "egn_conv_fwd": here egn_conv and egn_conv_tc.  "egn_eca_gate": its building blocks
 *     are egn_global_pool and egn_broadcast_mul, the fused gate lives inside egn_forward.  "egn_head_global" /
 *     "egn_head_local": the global_out / desc_out arguments of egn_forward select the heads.  "egn_load_weights": the
 *     egn_net struct + weight blob passed to egn_forward.  egn_allgather_global is exported under that name.
 *   - polar quantisation: atan2f / sqrtf on the device may differ from the host libm by one ulp, so a point lying
 *     exactly on a sector / ring boundary can fall into the neighbouring voxel; the policy (tests/test_gpu_parity.py)
 *     is "identical voxel SET up to 1e-3 of the voxels"; cartesian quantisation is bit-exact.
 */
#ifndef EGONN_B200_H
#define EGONN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGN_MAX_LEVELS 8   /* network levels 0..7 (tensor stride 1..128) */
#define EGN_PYR_LEVELS 10  /* coordinate pyramid levels kept per context: 0..9 */
#define EGN_MAX_HEAD_LEVELS 4
#define EGN_MAX_EXTRA_BLOCKS 3 /* residual blocks per level beyond the first (layers[L] <= 4) */

typedef struct egn_ctx egn_ctx;
typedef void *egn_stream_t; /* cudaStream_t */

typedef enum {
  EGN_OK = 0,
  EGN_ERR_INVALID = -1,  /* bad argument */
  EGN_ERR_CUDA = -2,     /* CUDA runtime error (text in egn_last_error) */
  EGN_ERR_RANGE = -3,    /* coordinate / batch index outside the supported range */
  EGN_ERR_STATE = -4,    /* call order (e.g. forward before coords_build) */
  EGN_ERR_CAPACITY = -5  /* an output buffer is too small */
} egn_status;

const char *egn_last_error(void);
int egn_version(void);

/* ---- context = coordinate manager + scratch arena ------------------------------------------------
 * Replaces: the MinkowskiEngine CoordinateManager created by ME.SparseTensor(features, coordinates=...)
 * at models/minkgl.py:269 (and again by layers/pooling.py:84). */
int egn_ctx_create(egn_ctx **out, int device);
int egn_ctx_destroy(egn_ctx *ctx);

/* ---- voxel quantisation ---------------------------------------------------------------------------
 * Replaces: CartesianQuantizer.__call__ datasets/quantization.py:79-85 and PolarQuantizer.__call__
 * :29-44, i.e. ME.utils.sparse_quantize(pc, quantization_size=q, return_index=True):
 * floor(f32 divide), int32, first-occurrence-wins de-duplication, survivors in input order.
 *   points      (n,3) f32        coords_out (n,3) int32 [capacity n]     index_out (n) int64 [capacity n]
 *   polar != 0: (x,y,z) -> (180 + atan2(y,x)*180/pi, sqrt(x^2+y^2), z) / step[0..2] first (step = deg, m, m);
 *   polar == 0: all three axes divided by step[0].
 *   n_out (host) receives the number of voxels.  Synchronises `stream`. */
int egn_quantize(egn_ctx *ctx, const float *points, int64_t n, const float step[3], int polar,
                 int32_t *coords_out, int64_t *index_out, int64_t *n_out, egn_stream_t stream);

/* ---- coordinate pyramid + kernel maps ---------------------------------------------------------------
 * Replaces: the coordinate hash-table build of ME.SparseTensor (models/minkgl.py:269), the strided maps
 * created by the kernel-2 stride-2 convolutions (models/minkgl.py:104-105,145-146) and the 3x3x3 / 5x5x5 /
 * 2x2x2 kernel maps MinkowskiEngine generates and caches for models/minkgl.py:140-151 and
 * layers/eca_block.py:59-64. */
typedef struct {
  int32_t n_batches;                 /* max batch index + 1 */
  int32_t n_input;                   /* rows passed in */
  int32_t n_rows[EGN_PYR_LEVELS];    /* rows of the level-L map (n_rows[0] < n_input if duplicates were dropped) */
  int32_t status;                    /* egn_status of the device-side validation */
} egn_coords_info;

/* coords (n,4) int32 [b,x,y,z], any order; duplicates keep the first occurrence (ME RANDOM_SUBSAMPLE
 * keeps one arbitrary row).  Builds levels 0..9, child/parent links and the 27-neighbour tables.
 * info (host) is filled; synchronises `stream` once. */
int egn_coords_build(egn_ctx *ctx, const int32_t *coords, int64_t n, egn_coords_info *info, egn_stream_t stream);

/* Fused ingest: raw points -> pyramid in ONE sort.  Replaces the caller sequence eval/evaluate.py:331-335
 * (quantizer(pc) per cloud, ME.utils.batched_coordinates, ones features) + the ME.SparseTensor build of
 * models/minkgl.py:269.  points (n,3) f32 = n_clouds clouds concatenated; cloud_offsets (n_clouds+1) int32 DEVICE
 * array of first-point indices (offsets[n_clouds] == n); step/polar as in egn_quantize.  The batch index of a voxel
 * is its cloud; info->n_input = n points; the "input row" of a level-0 voxel is its first point.  egn_forward may
 * then be called with features == NULL (all-ones occupancy features, what every reference caller feeds). */
int egn_coords_build_points(egn_ctx *ctx, const float *points, int64_t n, const int32_t *cloud_offsets, int n_clouds,
                            const float step[3], int polar, egn_coords_info *info, egn_stream_t stream);

/* Copy the coordinates of level L (tensor stride 2^L) into out (n_rows[L],4) int32, canonical order. */
int egn_coords_get(egn_ctx *ctx, int level, int32_t *out, egn_stream_t stream);
/* Input row that became canonical L0 row r: out (n_rows[0]) int32. */
int egn_coords_input_rows(egn_ctx *ctx, int32_t *out, egn_stream_t stream);
/* First row of each batch index at level L: out (n_batches+1) int32. */
int egn_coords_batch_offsets(egn_ctx *ctx, int level, int32_t *out, egn_stream_t stream);
/* 27-neighbour table of level L >= 1 (k = kx + 3(ky + 3kz)): out (n_rows[L],27) int32, -1 = absent. */
int egn_coords_neighbors(egn_ctx *ctx, int level, int32_t *out, egn_stream_t stream);

/* ---- network description (weights live in one caller-owned f32 device blob; fields are offsets in floats,
 *      -1 = absent).  Filled by egonn_b200/weights.py from the reference state_dict. */
typedef struct {
  int32_t cin, cout;
  int64_t w;      /* (K,cin,cout) kernel, K = 125/27/8/1 */
  int64_t scale;  /* (cout) folded eval-mode BatchNorm scale gamma/sqrt(var+eps), or -1 */
  int64_t shift;  /* (cout) folded BatchNorm shift / Linear bias, or -1 */
  int64_t wtc;    /* tensor-core image of the kernel (egn_conv_tc layout, bf16 hi/lo, offset in floats), or -1 */
} egn_layer;

typedef struct {
  int32_t n_levels;                       /* number of input levels, 0 = head absent */
  int32_t levels[EGN_MAX_HEAD_LEVELS];    /* ascending trunk levels feeding the head */
  int32_t out_channels;
  egn_layer conv1x1[EGN_MAX_LEVELS];      /* indexed by trunk level */
  egn_layer tconv[EGN_MAX_LEVELS];        /* tconv[L]: level L -> L-1 */
} egn_head;

typedef struct {
  int32_t n_levels;                       /* trunk levels 1..n_levels */
  int32_t conv0_ksize;                    /* 5 (or 3) */
  egn_layer conv0;                        /* + BN + ReLU    models/minkgl.py:100-102,140-142 */
  egn_layer down[EGN_MAX_LEVELS];         /* [L] 2x2x2 stride-2 + BN + ReLU   :104-107,145-148 */
  egn_layer conv1[EGN_MAX_LEVELS];        /* [L] block conv1 + norm1 + ReLU   layers/eca_block.py:59-61 */
  egn_layer conv2[EGN_MAX_LEVELS];        /* [L] block conv2 + norm2          :63-64 */
  egn_layer res[EGN_MAX_LEVELS];          /* [L] downsample 1x1 + BN (cin == 0: identity)  models/minkgl.py:121-126 */
  int32_t eca_k[EGN_MAX_LEVELS];          /* ECA Conv1d kernel size (0 = plain BasicBlock)  layers/eca_block.py:14-17 */
  int64_t eca_w[EGN_MAX_LEVELS];
  egn_head global_head;                   /* models/minkgl.py:46-60 */
  egn_head local_head;
  egn_layer global_mlp[2];                /* DescriptorDecoder (cin == 0: absent)  models/minkgl.py:207-225 */
  int32_t pool_method;                    /* 0 GeM, 1 SPoC (mean), 2 MAC (max)     layers/pooling.py:46-86 */
  float gem_p, gem_eps;
  egn_layer desc_mlp[2];                  /* local DescriptorDecoder + L2 normalise */
  egn_layer kp_mlp[2];                    /* KeypointRegressor + tanh      models/minkgl.py:175-185 */
  egn_layer sigma_mlp[2];                 /* SigmaRegressor + softplus     models/minkgl.py:188-204 */
  int32_t polar;                          /* keypoint_position: datasets/quantization.py:60-72 / :93-103 */
  float quant_step[3];
  int32_t ignore_keypoint_regressor;      /* models/minkgl.py:296-299 */
  /* Optional fused form of kp_mlp + sigma_mlp (cin == 0: absent, the separate layers are used): both regressors read the same
   * local map, so their first Linear layers run as ONE 64 -> 32+32 layer and their second layers as ONE block-diagonal
   * (32+32) -> 3+1 layer; identical arithmetic per output (the extra weights are exact zeros).  models/minkgl.py:175-204 */
  egn_layer kpsig_mlp[2];
  /* Blocks 1.. of a level when layers[L] > 1 (MinkTrunk._make_layer models/minkgl.py:121-134 / ResNetBase._make_layer
   * models/resnet.py:81-97: the blocks after the first keep the channel count and have an identity residual).  Block j+1
   * of level L: xconv1[L][j] + norm1 + ReLU, xconv2[L][j] + norm2, ECA gate xeca_*[L][j] (k == 0: plain BasicBlock). */
  int32_t n_extra[EGN_MAX_LEVELS];
  egn_layer xconv1[EGN_MAX_LEVELS][EGN_MAX_EXTRA_BLOCKS];
  egn_layer xconv2[EGN_MAX_LEVELS][EGN_MAX_EXTRA_BLOCKS];
  int32_t xeca_k[EGN_MAX_LEVELS][EGN_MAX_EXTRA_BLOCKS];
  int64_t xeca_w[EGN_MAX_LEVELS][EGN_MAX_EXTRA_BLOCKS];
} egn_net;

/* Keep the weight blob resident in L2 across forwards: reserves a persisting-L2 carve-out
 * (cudaLimitPersistingL2CacheSize, clamped to the device maximum) and makes egn_forward tag accesses to
 * [weights, weights+bytes) as persisting on its streams (cudaStreamAttributeAccessPolicyWindow).  The 37 MB of
 * fp32 + bf16 kernel images are re-read by every CTA of every layer of every batch; the ~1 GB of activations that
 * stream through the 126 MB L2 per batch would otherwise evict them.  bytes == 0 removes the window. */
int egn_weights_resident(egn_ctx *ctx, const void *weights, size_t bytes);

/* ---- whole forward ----------------------------------------------------------------------------------
 * Replaces: MinkGL.forward models/minkgl.py:267-315 (MinkTrunk.forward :136-153, MinkHead.forward :46-60,
 * ECABasicBlock.forward layers/eca_block.py:56-73, GeM.forward layers/pooling.py:82-86, the three
 * regressors/decoders and Quantizer.keypoint_position) and MinkLoc.forward models/minkloc.py:44-61.
 * Requires egn_coords_build on the same ctx.  features (n_input) f32 in INPUT row order, or NULL = all ones.
 *   global_out       (n_batches, global dim) f32, or NULL to skip the global head
 *   desc_out         (n_rows[Llocal], desc dim) f32 |
 *   keypoints_out    (n_rows[Llocal], 3) f32        |  all NULL to skip the local head
 *   sigma_out        (n_rows[Llocal], 1) f32        |
 * Local rows are in canonical order of level Llocal = min(local_head.levels); use egn_coords_get /
 * egn_coords_batch_offsets to split them per cloud.  Asynchronous on `stream`. */
int egn_forward(egn_ctx *ctx, const egn_net *net, const float *weights, const float *features,
                float *global_out, float *desc_out, float *keypoints_out, float *sigma_out,
                egn_stream_t stream);

/* Debug / parity taps: copy an intermediate feature map of the last egn_forward.
 * which: 0 conv0 output (L0), 1 down-conv output of `level`, 2 block output of `level`,
 *        3 global head map, 4 local head map.  out must hold rows*channels floats. */
int egn_forward_tap(egn_ctx *ctx, int which, int level, float *out, egn_stream_t stream);

/* ---- single operators on the context's coordinate maps (the MinkowskiEngine-shaped shim calls these) ----
 * Replaces: ME.MinkowskiConvolution / MinkowskiConvolutionTranspose forward.
 * ksize 1|3 (stride 1), 2 (stride 2: level_in -> level_in+1), 5 (level 0 only, cin == 1);
 * transposed != 0 with ksize 2: level_in -> level_in-1 on the existing map.
 * out = act((conv) * scale + shift) [+ out if accumulate]; scale/shift may be NULL; relu 0|1. */
int egn_conv(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout,
             const float *in, const float *w, const float *scale, const float *shift, int relu,
             int accumulate, float *out, egn_stream_t stream);
/* Tensor-core (tcgen05) variant of egn_conv for ksize 3 (cin->cout in {32->32, 32->64, 64->64, 64->128, 128->128})
 * ksize 2 stride 2, plain or transposed (cin == cout in {32, 64, 128}) and ksize 1 ({32->64, 64->64, 64->128,
 * 128->64, 128->128}).  wpack is the kernel pre-split into bf16 hi/lo and laid out
 * as the 128-byte-swizzled shared-memory images the kernel loads with one bulk copy per 64-element reduction chunk:
 * [ceil(K*cin/64)][hi|lo][cout][64] bf16 with 16-byte group g of row n stored at group g ^ (n & 7)
 * (egonn_b200/weights.py:pack_tc).  Results are fp32-class (bf16x3 split products, FP32 accumulation in TMEM). */
int egn_conv_tc(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout, const float *in, const void *wpack,
                const float *scale, const float *shift, int relu, float *out, egn_stream_t stream);
/* 1 (default): egn_forward runs layers that carry a tensor-core image on the tcgen05 path; 0: FP32 CUDA cores only. */
int egn_set_tensor_cores(egn_ctx *ctx, int enable);

/* Replaces: ME.MinkowskiGlobalPooling / GlobalAvgPooling / GlobalMaxPooling: out (n_batches, c). */
int egn_global_pool(egn_ctx *ctx, int level, int c, const float *in, int is_max, float *out, egn_stream_t stream);
/* Replaces: ME.MinkowskiBroadcastMultiplication: out[r] = in[r] * g[batch(r)]. */
int egn_broadcast_mul(egn_ctx *ctx, int level, int c, const float *in, const float *g, float *out, egn_stream_t stream);

/* ---- "next" rows: keypoint selection on device -------------------------------------------
 * Replaces: get_keypoints_idxes eval/evaluate.py:352-361 (torch.topk(sigma, k, largest=False)) per cloud.
 * sigma (n) f32, offsets (n_batches+1) int32 device; idx_out (n_batches,k) int32 rows (cloud-relative),
 * ascending sigma, ties by lower row; -1 padding when a cloud has fewer than k rows. */
int egn_topk_smallest(const float *sigma, const int32_t *offsets, int n_batches, int k, int32_t *idx_out,
                      egn_stream_t stream);

/* Replaces: the per-cloud selection that follows it, eval/evaluate.py:339-350 (descriptors[ndx], keypoints[ndx] next to the
 * cloud's global descriptor), for ALL clouds of a batch in one launch, packed for ONE device-to-host copy.
 * idx (n_batches,k) from egn_topk_smallest, offsets (n_batches+1), keypoints (n,3), descriptors (n,desc_dim), global
 * (n_batches, global_dim) or NULL with global_dim 0.  out (n_batches, global_dim + k*3 + k*desc_dim) f32 per cloud:
 * [global | keypoints (k,3) | descriptors (k,desc_dim)], zeros where idx == -1. */
int egn_pack_topk(const int32_t *idx, const int32_t *offsets, int n_batches, int k, const float *keypoints, const float *descriptors,
                  int desc_dim, const float *global, int global_dim, float *out, egn_stream_t stream);

/* Replaces: the global-descriptor nearest-neighbour search of eval/evaluate.py:173-176
 * (embed_dist = np.linalg.norm(map_embeddings - query_embedding, axis=1); nn_ndx = np.argsort(embed_dist)[:k]).
 * query (n_query, dim), map (n_map, dim) f32; idx_out (n_query, k) int32 map rows by ascending distance (ties: lower
 * row first, -1 padding if n_map < k); dist_out (n_query, n_map) f32 receives every distance (caller-owned scratch). */
int egn_knn_l2(egn_ctx *ctx, const float *query, const float *map, int n_query, int n_map, int dim, int k, int32_t *idx_out,
               float *dist_out, egn_stream_t stream);

/* ---- "next" rows: correspondences between the local descriptors of two clouds ------------------------
 * Replaces: the feature-matching step inside eval/evaluate.py:381-399 (Open3D
 * registration_ransac_based_on_feature_matching(..., mutual_filter=True): nearest neighbour in descriptor space, kept
 * when mutual).  desc_a (n_a, dim), desc_b (n_b, dim) f32; idx_out (n_a) int32 = row of b matched to every row of a
 * (-1: not mutual); dist_out (n_a) f32 Euclidean descriptor distance of the nearest neighbour, or NULL.  mutual = 0
 * returns the plain nearest neighbour.  Ties: lower row.  RANSAC itself stays out of scope. */
int egn_match_mutual(egn_ctx *ctx, const float *desc_a, const float *desc_b, int n_a, int n_b, int dim, int mutual, int32_t *idx_out,
                     float *dist_out, egn_stream_t stream);

/* ---- "next" rows: raw-scan ingest --------------------------------------------------------------------------
 * Replaces: PointCloudLoader.__call__ misc/point_clouds.py:95-111 after read_pc (datasets/kitti/kitti_raw.py:16-22,
 * datasets/mulran/mulran_raw.py:19-25: np.fromfile(...).reshape(-1, 4)[:, :3]).  records (n, stride) f32 with x, y, z in
 * the first three floats (stride 4 for the .bin files, 3 for xyz arrays); remove_zero drops points with all |v| <= 1e-8
 * (np.isclose(pc, 0)), remove_ground drops z <= ground_level (-1.5 KITTI, -0.9 MulRan); survivors keep their order.
 * points_out (n,3) f32 caller-owned, *n_out = number of survivors (synchronises the stream once). */
int egn_filter_points(egn_ctx *ctx, const float *records, int64_t n, int stride, int remove_zero, int remove_ground,
                      float ground_level, float *points_out, int64_t *n_out, egn_stream_t stream);

/* ---- multi-GPU: ONE all-gather of the global descriptors -----------------------------------------------------
 * Replaces: nothing in the reference - its evaluation loop is single-device (eval/evaluate.py:454-466).  When the clouds
 * of a batch are sharded over the GPUs of a node (one process per GPU, egonn_b200/parallel.py), every rank extracts its
 * own clouds and the ranks exchange the (clouds_per_rank, 256) f32 global descriptors with one NCCL all-gather over
 * NVLink; local descriptors and keypoints stay rank-local, and there is no collective inside the network.
 * NCCL is resolved at run time (dlopen of libnccl.so.2, or the path in EGN_NCCL_LIB); the library does not link it.
 *   egn_comm_unique_id : rank 0 fills `id_out` (host, EGN_COMM_ID_BYTES) and hands the bytes to the other ranks by any
 *                        out-of-band channel (torch.distributed store, MPI, a file).
 *   egn_comm_create    : collective over all ranks (ncclCommInitRank) on `device`.
 *   egn_allgather_global: recv (world * floats_per_rank) <- send (floats_per_rank) of every rank, in rank order;
 *                        asynchronous on `stream`; every rank must pass the same floats_per_rank (pad uneven shards). */
#define EGN_COMM_ID_BYTES 128
typedef struct egn_comm egn_comm;
int egn_comm_unique_id(void *id_out);
int egn_comm_create(egn_comm **out, int device, int rank, int world, const void *id);
int egn_comm_destroy(egn_comm *comm);
int egn_allgather_global(egn_comm *comm, const float *send, float *recv, int64_t floats_per_rank, egn_stream_t stream);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------------
 * egn_profile_enable(ctx, 1): every kernel class launched by this context is bracketed by CUDA events on its
 * stream and the pair counts needed for the algorithmic-byte model are computed at coords_build.
 * egn_profile_read drains the finished events into per-class totals (synchronises the events, not the device).
 * alg_bytes follows SURVEY.md 8(d): conv P*(Cin+Cout)*4 + P*8 + K*Cin*Cout*4; row-wise R*(Cin+Cout)*4;
 * kernel-map build N*8 + N*K*4.  egn_launch_count: kernels of THIS library launched so far by the context. */
typedef struct {
  char name[48];
  int64_t launches;
  double ms;         /* summed device time between the class's start/stop events */
  double alg_bytes;  /* summed algorithmic bytes */
  double flops;      /* summed useful FLOPs (2*P*Cin*Cout for convolutions) */
} egn_profile_entry;
int egn_profile_enable(egn_ctx *ctx, int enable);
int egn_profile_read(egn_ctx *ctx, egn_profile_entry *out, int capacity, int *n_out, int reset);
int64_t egn_launch_count(egn_ctx *ctx);
/* debug only (EGN_TRACE=1): clock64 timeline of CTA 0 of the last tensor-core convolution, 64 chunks x 8 stamps (host) */
int egn_debug_trace(egn_ctx *ctx, long long *host_out);

#ifdef __cplusplus
}
#endif
#endif /* EGONN_B200_H */
