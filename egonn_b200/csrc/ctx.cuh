// Engine context: coordinate pyramid ("coordinate manager") + scratch arenas.
#pragma once
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace egn {

constexpr int P = EGN_PYR_LEVELS;  // pyramid levels 0..P-1
constexpr int kPoolCounters = 1024; // clouds per batch that the single-launch poolings (ops.cu: PoolTail) can count

struct Pyramid {
  bool valid = false;
  int n_input = 0, n_batches = 0;
  int n[P] = {0};
  // per level L (device pointers into ctx->coords arena)
  uint64_t *keys[P] = {nullptr};   // sorted level-L keys (level-0 key >> 3L)
  int *up[P] = {nullptr};          // row at L -> parent row at L+1           (L < P-1)
  int *cstart[P] = {nullptr};      // row at L -> first child row at L-1      (L >= 1)
  uint32_t *cmask[P] = {nullptr};  // row at L -> 8-bit child occupancy       (L >= 1)
  int *nbr[P] = {nullptr};         // (n[L],27) neighbour rows at L, -1 absent (L >= 1)
  uint32_t *mask27[P] = {nullptr}; // row at L -> 27-bit presence mask of its 3x3x3 neighbourhood (L >= 1)
  // tile row orders (coords.cu "tile row orders"): permutations of the level's rows used ONLY to form the 128-row tiles of
  // the tensor-core convolutions - rows with similar offset sets adjacent, so that whole (tile, chunk) pairs vanish
  int *ord27[P] = {nullptr};       // 3x3x3 convolutions at L            (L >= 1)
  int *ordc[P] = {nullptr};        // 2x2x2 stride-2 convolution INTO L  (L >= 1; keyed by the child mask)
  int *ordt[P] = {nullptr};        // transposed 2x2x2 convolution INTO L (L < P-1; keyed by the row's own child code)
  bool ordered = false;
  int *boff[P] = {nullptr};        // (n_batches+1) first row of each batch
  int *perm0 = nullptr;            // canonical L0 row -> input row
  uint64_t *mask64 = nullptr;      // per L2 cell: occupancy of its 4x4x4 level-0 voxels (bit = key0 & 63)
  int *first0 = nullptr;           // per L2 cell: first level-0 row
  long long pairs27[P] = {0};      // present (out,in) pairs of the 3^3 kernel map (profile mode only)
  long long pairs_conv0 = 0;       // present pairs of the conv0 window (profile mode only)
};

struct HostCounts {  // pinned
  int totals[P];
  int n_batches;
  int status;
  int n_out;
  int flags;   // [P+3] bit 0: some input feature != 1.0f (set by the forward's feature gather)
};

// per-kernel-class event timing (bench.py's live roofline numbers)
struct Prof {
  bool on = false;
  long long launches = 0;
  struct Pending { int entry; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  std::vector<egn_profile_entry> entries;
  std::vector<cudaEvent_t> pool;
  int begin(const char *name, double bytes, double flops, cudaStream_t s);
  void end(cudaStream_t s);
  int drain();
  int cur = -1;
  cudaEvent_t cur_a = nullptr;
};

struct Taps {  // device pointers of the last forward's feature maps (feature arena)
  float *conv0 = nullptr;
  float *down[EGN_MAX_LEVELS] = {nullptr};
  float *block[EGN_MAX_LEVELS] = {nullptr};
  int c_down[EGN_MAX_LEVELS] = {0}, c_block[EGN_MAX_LEVELS] = {0};
  int c0 = 0;
  bool conv0_split = false, down_split[EGN_MAX_LEVELS] = {false}, block_split[EGN_MAX_LEVELS] = {false};   // pre-split maps (tc_ptx.cuh)
  float *gmap = nullptr, *lmap = nullptr;
  int c_g = 0, c_l = 0, lvl_g = 0, lvl_l = 0;
};

}  // namespace egn

struct egn_ctx {
  int device = 0;
  egn::Arena scratch;   // sort double-buffers, CUB temp, tile counts (dead after coords_build / quantize)
  egn::Arena coords;    // pyramid
  egn::Arena feats;     // feature maps of the running forward
  egn::Pyramid pyr;
  egn::HostCounts *host = nullptr;  // pinned
  int *dev_counts = nullptr;        // device mirror of HostCounts
  int *pool_counters = nullptr;     // [kPoolCounters] "slices of cloud b finished" counters of the single-launch poolings (zero between launches)
  egn::Taps taps;
  egn::Prof prof;
  bool use_tc = true;
  bool light_ctas = true;           // short-tile convolutions as four light CTAs per SM (EGN_LIGHT=0 disables)
  bool use_order = true;            // mask-sorted tile row orders for the tensor-core convolutions (EGN_ORDER=0 disables)
  int order_window = 8192;          // rows are re-grouped inside windows of this many canonical rows (EGN_ORDER_WINDOW = 2048 | 4096 | 8192)
  // kernels already opted in to > 48 KB dynamic shared memory ON THIS CONTEXT'S DEVICE: the attribute is per device, a
  // context belongs to one device, so the bookkeeping lives here and not in process-wide statics (egn_smem_optin below)
  std::unordered_set<const void *> smem_optin;
  cudaStream_t aux = nullptr;       // second stream: the local head overlaps the upper trunk levels
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  const void *win_ptr = nullptr;    // persisting-L2 window over the weight blob (egn_weights_resident)
  size_t win_bytes = 0;
  unsigned hint_producer = 20000, hint_single = 20000;   // mbarrier try_wait suspend hints (EGN_HINT_P / EGN_HINT_S)
  void *trace = nullptr;            // debug timeline buffer for k_sconv_tc (EGN_TRACE=1 allocates 64*8 int64)
  int nsplit_max = 74;              // N-split 128-channel convolutions up to this many row tiles (EGN_NSPLIT_MAX)
  bool ksplit = false;              // K-split of small 128-channel levels (EGN_KSPLIT=1)
  bool in_forward = false;          // inside egn_forward: temporaries come from the planned feature arena
};

// opt a kernel in to `bytes` of dynamic shared memory once per context (the current device must be ctx->device)
#define EGN_SMEM_OPTIN(ctx, kernel, bytes)                                                                              \
  do {                                                                                                                   \
    const void *egn_f_ = (const void *)(kernel);                                                                         \
    if (!(ctx)->smem_optin.count(egn_f_)) {                                                                              \
      EGN_CUDA(cudaFuncSetAttribute(egn_f_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));                 \
      (ctx)->smem_optin.insert(egn_f_);                                                                                  \
    }                                                                                                                    \
  } while (0)

// bracket one kernel class: counts the launch, and in profile mode records start/stop events
#define EGN_LAUNCH(ctx, name, bytes, flops, stream, ...)            \
  do {                                                               \
    (ctx)->prof.launches++;                                          \
    if ((ctx)->prof.on) (ctx)->prof.begin(name, bytes, flops, stream); \
    __VA_ARGS__;                                                     \
    if ((ctx)->prof.on) (ctx)->prof.end(stream);                     \
  } while (0)

namespace egn {
// coords.cu
int coords_build(egn_ctx *ctx, const int32_t *coords, int64_t n, egn_coords_info *info, cudaStream_t s);
int coords_build_points(egn_ctx *ctx, const float *points, int64_t n, const int32_t *cloud_offsets, int n_clouds,
                        const float step[3], int polar, egn_coords_info *info, cudaStream_t s);
int coords_get(egn_ctx *ctx, int level, int32_t *out, cudaStream_t s);
int quantize(egn_ctx *ctx, const float *points, int64_t n, const float step[3], int polar, int32_t *coords_out,
             int64_t *index_out, int64_t *n_out, cudaStream_t s);
int filter_points(egn_ctx *ctx, const float *records, int64_t n, int stride, int remove_zero, int remove_ground, float ground_level,
                  float *points_out, int64_t *n_out, cudaStream_t s);
// ops.cu
int op_conv(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout, const float *in, const float *w,
            const float *scale, const float *shift, int relu, int accumulate, float *out, cudaStream_t s);
int op_global_pool(egn_ctx *ctx, int level, int c, const float *in, int mode, float p, float eps, float *out,
                   cudaStream_t s);
int op_broadcast_mul(egn_ctx *ctx, int level, int c, const float *in, const float *g, float *out, cudaStream_t s);
int op_knn_l2(egn_ctx *ctx, const float *query, const float *map, int Q, int M, int D, int k, int32_t *idx_out, float *dist_out,
              cudaStream_t s);
int op_match_mutual(egn_ctx *ctx, const float *a, const float *b, int na, int nb, int d, int mutual, int32_t *idx_out, float *dist_out,
                    cudaStream_t s);
int op_topk(const float *sigma, const int32_t *offsets, int n_batches, int k, int32_t *idx_out, cudaStream_t s);
int op_pack_topk(const int32_t *idx, const int32_t *offsets, int n_batches, int k, const float *kp, const float *desc, int D,
                 const float *glob, int G, float *out, cudaStream_t s);
// sconv_tc.cu
bool sconv_tc_supported(int ksize, int transposed, int cin, int cout);
int run_conv_tc(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout, const float *in, const void *wpack,
                const float *scale, const float *shift, int relu, int accumulate, float *out, cudaStream_t s, int in_split = 0,
                int out_split = 0);
// forward.cu
int forward(egn_ctx *ctx, const egn_net *net, const float *weights, const float *features, float *global_out,
            float *desc_out, float *kp_out, float *sigma_out, cudaStream_t s);
int forward_tap(egn_ctx *ctx, int which, int level, float *out, cudaStream_t s);
}  // namespace egn
