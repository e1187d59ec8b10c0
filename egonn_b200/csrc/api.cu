// extern "C" surface of libegonn_b200.so (declared in include/egonn_b200.h) + context / arena plumbing.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "ctx.cuh"

namespace egn {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int Arena::reserve(size_t bytes, cudaStream_t stream) {
  used = 0;
  if (bytes <= cap) return EGN_OK;
  // grow: kernels enqueued earlier may still read the old block
  EGN_CUDA(cudaStreamSynchronize(stream));
  if (base) EGN_CUDA(cudaFree(base));
  base = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 4 + (1 << 20);
  EGN_CUDA(cudaMalloc((void **)&base, want));
  cap = want;
  return EGN_OK;
}
void Arena::release() {
  if (base) cudaFree(base);
  base = nullptr;
  cap = used = 0;
}

int Prof::begin(const char *name, double bytes, double flops, cudaStream_t s) {
  int e = -1;
  for (size_t i = 0; i < entries.size(); ++i)
    if (strncmp(entries[i].name, name, sizeof(entries[i].name)) == 0) { e = (int)i; break; }
  if (e < 0) {
    egn_profile_entry pe;
    memset(&pe, 0, sizeof(pe));
    strncpy(pe.name, name, sizeof(pe.name) - 1);
    entries.push_back(pe);
    e = (int)entries.size() - 1;
  }
  entries[e].launches++;
  entries[e].alg_bytes += bytes;
  entries[e].flops += flops;
  cudaEvent_t a = nullptr;
  if (!pool.empty()) { a = pool.back(); pool.pop_back(); } else cudaEventCreate(&a);
  cudaEventRecord(a, s);
  cur = e;
  cur_a = a;
  return e;
}
void Prof::end(cudaStream_t s) {
  if (cur < 0) return;
  cudaEvent_t b = nullptr;
  if (!pool.empty()) { b = pool.back(); pool.pop_back(); } else cudaEventCreate(&b);
  cudaEventRecord(b, s);
  pending.push_back({cur, cur_a, b});
  cur = -1;
  cur_a = nullptr;
}
int Prof::drain() {
  for (auto &p : pending) {
    cudaEventSynchronize(p.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) entries[p.entry].ms += ms;
    pool.push_back(p.a);
    pool.push_back(p.b);
  }
  pending.clear();
  return EGN_OK;
}

}  // namespace egn

using namespace egn;

extern "C" {

const char *egn_last_error(void) { return g_err; }
int egn_version(void) { return 100; }

int egn_ctx_create(egn_ctx **out, int device) {
  EGN_CHECK(out != nullptr, EGN_ERR_INVALID, "ctx_create: null out");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  EGN_CHECK(e == cudaSuccess && count > 0, EGN_ERR_CUDA, "ctx_create: no CUDA device (%s) - this engine has no CPU path",
            cudaGetErrorString(e));
  EGN_CHECK(device >= 0 && device < count, EGN_ERR_INVALID, "ctx_create: device %d of %d", device, count);
  DeviceGuard g(device);
  cudaDeviceProp prop;
  EGN_CUDA(cudaGetDeviceProperties(&prop, device));
  EGN_CHECK(prop.major == 10, EGN_ERR_CUDA, "ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
            prop.major, prop.minor);
  egn_ctx *ctx = new egn_ctx();
  ctx->device = device;
  cudaError_t e1 = cudaMallocHost((void **)&ctx->host, sizeof(HostCounts));
  cudaError_t e2 = cudaMalloc((void **)&ctx->dev_counts, sizeof(HostCounts));
  if (e2 == cudaSuccess) e2 = cudaMalloc((void **)&ctx->pool_counters, kPoolCounters * sizeof(int));
  if (e2 == cudaSuccess) e2 = cudaMemset(ctx->pool_counters, 0, kPoolCounters * sizeof(int));
  if (e2 == cudaSuccess) e2 = cudaDeviceSynchronize();     // the counters are zero before any (non-blocking) stream can use them
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    set_error("ctx_create: allocation failed");
    delete ctx;
    return EGN_ERR_CUDA;
  }
  cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking);
  const char *ks = getenv("EGN_KSPLIT");
  ctx->ksplit = ks && ks[0] == '1';
  if (const char *t = getenv("EGN_TRACE")) if (t[0] == '1') { cudaMalloc(&ctx->trace, 64 * 8 * 8); cudaMemset(ctx->trace, 0, 64 * 8 * 8); }
  if (const char *v = getenv("EGN_NSPLIT_MAX")) ctx->nsplit_max = atoi(v);
  if (const char *v = getenv("EGN_ORDER")) ctx->use_order = v[0] != '0';
  if (const char *v = getenv("EGN_LIGHT")) ctx->light_ctas = v[0] != '0';
  if (const char *v = getenv("EGN_ORDER_WINDOW")) {
    const int w = atoi(v);
    if (w == 2048 || w == 4096 || w == 8192) ctx->order_window = w;
  }
  if (const char *h = getenv("EGN_HINT_P")) ctx->hint_producer = (unsigned)atoi(h);
  if (const char *h = getenv("EGN_HINT_S")) ctx->hint_single = (unsigned)atoi(h);
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  *out = ctx;
  return EGN_OK;
}

int egn_ctx_destroy(egn_ctx *ctx) {
  if (!ctx) return EGN_OK;
  DeviceGuard g(ctx->device);
  cudaDeviceSynchronize();
  ctx->scratch.release();
  ctx->coords.release();
  ctx->feats.release();
  ctx->prof.drain();
  for (auto e : ctx->prof.pool) cudaEventDestroy(e);
  if (ctx->aux) cudaStreamDestroy(ctx->aux);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->host) cudaFreeHost(ctx->host);
  if (ctx->dev_counts) cudaFree(ctx->dev_counts);
  if (ctx->pool_counters) cudaFree(ctx->pool_counters);
  delete ctx;
  return EGN_OK;
}

int egn_quantize(egn_ctx *ctx, const float *points, int64_t n, const float step[3], int polar, int32_t *coords_out,
                 int64_t *index_out, int64_t *n_out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return quantize(ctx, points, n, step, polar, coords_out, index_out, n_out, (cudaStream_t)stream);
}

int egn_coords_build(egn_ctx *ctx, const int32_t *coords, int64_t n, egn_coords_info *info, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return coords_build(ctx, coords, n, info, (cudaStream_t)stream);
}

int egn_coords_build_points(egn_ctx *ctx, const float *points, int64_t n, const int32_t *cloud_offsets, int n_clouds,
                            const float step[3], int polar, egn_coords_info *info, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return coords_build_points(ctx, points, n, cloud_offsets, n_clouds, step, polar, info, (cudaStream_t)stream);
}

int egn_coords_get(egn_ctx *ctx, int level, int32_t *out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return coords_get(ctx, level, out, (cudaStream_t)stream);
}

int egn_coords_input_rows(egn_ctx *ctx, int32_t *out, egn_stream_t stream) {
  EGN_CHECK(ctx && ctx->pyr.valid && out, EGN_ERR_STATE, "coords_input_rows before coords_build");
  DeviceGuard g(ctx->device);
  EGN_CUDA(cudaMemcpyAsync(out, ctx->pyr.perm0, (size_t)ctx->pyr.n[0] * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return EGN_OK;
}

int egn_coords_batch_offsets(egn_ctx *ctx, int level, int32_t *out, egn_stream_t stream) {
  EGN_CHECK(ctx && ctx->pyr.valid && out, EGN_ERR_STATE, "coords_batch_offsets before coords_build");
  EGN_CHECK(level >= 0 && level < P, EGN_ERR_INVALID, "bad level");
  DeviceGuard g(ctx->device);
  EGN_CUDA(cudaMemcpyAsync(out, ctx->pyr.boff[level], (size_t)(ctx->pyr.n_batches + 1) * 4, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return EGN_OK;
}

int egn_coords_neighbors(egn_ctx *ctx, int level, int32_t *out, egn_stream_t stream) {
  EGN_CHECK(ctx && ctx->pyr.valid && out, EGN_ERR_STATE, "coords_neighbors before coords_build");
  EGN_CHECK(level >= 1 && level < P, EGN_ERR_INVALID, "neighbour tables exist for levels 1..%d", P - 1);
  DeviceGuard g(ctx->device);
  EGN_CUDA(cudaMemcpyAsync(out, ctx->pyr.nbr[level], (size_t)ctx->pyr.n[level] * 27 * 4, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return EGN_OK;
}

static void apply_window(egn_ctx *ctx, cudaStream_t s) {
  cudaStreamAttrValue v;
  memset(&v, 0, sizeof(v));
  v.accessPolicyWindow.base_ptr = const_cast<void *>(ctx->win_ptr);
  v.accessPolicyWindow.num_bytes = ctx->win_bytes;
  v.accessPolicyWindow.hitRatio = 1.0f;
  v.accessPolicyWindow.hitProp = ctx->win_bytes ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v);   // best effort
  cudaGetLastError();
}

int egn_weights_resident(egn_ctx *ctx, const void *weights, size_t bytes) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  cudaDeviceProp prop;
  EGN_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
  size_t want = bytes;
  if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
  if (want > (size_t)prop.accessPolicyMaxWindowSize) want = (size_t)prop.accessPolicyMaxWindowSize;
  if (bytes) EGN_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
  ctx->win_ptr = bytes ? weights : nullptr;
  ctx->win_bytes = bytes ? want : 0;
  if (!bytes) cudaCtxResetPersistingL2Cache();
  return EGN_OK;
}

int egn_forward(egn_ctx *ctx, const egn_net *net, const float *weights, const float *features, float *global_out,
                float *desc_out, float *keypoints_out, float *sigma_out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  if (ctx->win_ptr) {
    apply_window(ctx, (cudaStream_t)stream);
    if (ctx->aux) apply_window(ctx, ctx->aux);
  }
  return forward(ctx, net, weights, features, global_out, desc_out, keypoints_out, sigma_out, (cudaStream_t)stream);
}

int egn_forward_tap(egn_ctx *ctx, int which, int level, float *out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return forward_tap(ctx, which, level, out, (cudaStream_t)stream);
}

int egn_conv(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout, const float *in, const float *w,
             const float *scale, const float *shift, int relu, int accumulate, float *out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return op_conv(ctx, level_in, ksize, transposed, cin, cout, in, w, scale, shift, relu, accumulate, out, (cudaStream_t)stream);
}

int egn_conv_tc(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout, const float *in, const void *wpack,
                const float *scale, const float *shift, int relu, float *out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr && in && wpack && out, EGN_ERR_INVALID, "conv_tc: null argument");
  DeviceGuard g(ctx->device);
  return run_conv_tc(ctx, level_in, ksize, transposed, cin, cout, in, wpack, scale, shift, relu, 0, out, (cudaStream_t)stream);
}

int egn_set_tensor_cores(egn_ctx *ctx, int enable) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  ctx->use_tc = enable != 0;
  return EGN_OK;
}

int egn_global_pool(egn_ctx *ctx, int level, int c, const float *in, int is_max, float *out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return op_global_pool(ctx, level, c, in, is_max ? 2 : 0, 1.f, 0.f, out, (cudaStream_t)stream);
}

int egn_broadcast_mul(egn_ctx *ctx, int level, int c, const float *in, const float *gate, float *out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return op_broadcast_mul(ctx, level, c, in, gate, out, (cudaStream_t)stream);
}

int egn_knn_l2(egn_ctx *ctx, const float *query, const float *map, int n_query, int n_map, int dim, int k, int32_t *idx_out,
               float *dist_out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return op_knn_l2(ctx, query, map, n_query, n_map, dim, k, idx_out, dist_out, (cudaStream_t)stream);
}

int egn_match_mutual(egn_ctx *ctx, const float *desc_a, const float *desc_b, int n_a, int n_b, int dim, int mutual, int32_t *idx_out,
                     float *dist_out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return op_match_mutual(ctx, desc_a, desc_b, n_a, n_b, dim, mutual, idx_out, dist_out, (cudaStream_t)stream);
}

int egn_filter_points(egn_ctx *ctx, const float *records, int64_t n, int stride, int remove_zero, int remove_ground, float ground_level,
                      float *points_out, int64_t *n_out, egn_stream_t stream) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  DeviceGuard g(ctx->device);
  return filter_points(ctx, records, n, stride, remove_zero, remove_ground, ground_level, points_out, n_out, (cudaStream_t)stream);
}

int egn_profile_enable(egn_ctx *ctx, int enable) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  ctx->prof.on = enable != 0;
  return EGN_OK;
}

int egn_profile_read(egn_ctx *ctx, egn_profile_entry *out, int capacity, int *n_out, int reset) {
  EGN_CHECK(ctx && out && n_out, EGN_ERR_INVALID, "profile_read: null argument");
  DeviceGuard g(ctx->device);
  ctx->prof.drain();
  const int n = (int)ctx->prof.entries.size();
  EGN_CHECK(n <= capacity, EGN_ERR_CAPACITY, "profile_read: %d entries, capacity %d", n, capacity);
  for (int i = 0; i < n; ++i) out[i] = ctx->prof.entries[i];
  *n_out = n;
  if (reset) ctx->prof.entries.clear();
  return EGN_OK;
}

int64_t egn_launch_count(egn_ctx *ctx) { return ctx ? ctx->prof.launches : 0; }

/* debug: copy the k_sconv_tc timeline (EGN_TRACE=1) to host memory: 64 chunks x 8 clock64 stamps */
int egn_debug_trace(egn_ctx *ctx, long long *host_out) {
  EGN_CHECK(ctx && ctx->trace && host_out, EGN_ERR_STATE, "trace buffer not allocated (EGN_TRACE=1)");
  EGN_CUDA(cudaMemcpy(host_out, ctx->trace, 64 * 8 * 8, cudaMemcpyDeviceToHost));
  return EGN_OK;
}

int egn_topk_smallest(const float *sigma, const int32_t *offsets, int n_batches, int k, int32_t *idx_out, egn_stream_t stream) {
  return op_topk(sigma, offsets, n_batches, k, idx_out, (cudaStream_t)stream);
}
int egn_pack_topk(const int32_t *idx, const int32_t *offsets, int n_batches, int k, const float *keypoints, const float *descriptors,
                  int desc_dim, const float *global, int global_dim, float *out, egn_stream_t stream) {
  return op_pack_topk(idx, offsets, n_batches, k, keypoints, descriptors, desc_dim, global, global_dim, out, (cudaStream_t)stream);
}

}  // extern "C"
