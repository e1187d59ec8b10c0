// Whole-network scheduler: MinkGL.forward (models/minkgl.py:267-315) / MinkLoc.forward (models/minkloc.py:44-61)
// as one stream of kernels on the context's coordinate pyramid.  Feature maps live in the context's feature
// arena (exact sizes - the level row counts are known on the host after coords_build).
#include "ctx.cuh"

namespace egn {

// ops.cu
int run_conv0(egn_ctx *ctx, int ksize, const float *f0, const float *w, const float *scale, const float *shift, int cout,
              int relu, const int *not_ones, int out_split, float *out, cudaStream_t s, const void *wtc);
int run_presplit_to_f32(egn_ctx *ctx, const float *in, int64_t floats, float *out, cudaStream_t s);
int run_conv(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout, const float *in, const float *w,
             const float *scale, const float *shift, int relu, int accumulate, float *out, cudaStream_t s);
int run_pool(egn_ctx *ctx, int level, int c, const float *in, int mode, float p, float eps, float *part, int slices,
             float *out, cudaStream_t s);
int run_gather_rows1(egn_ctx *ctx, const float *in, const int *perm, int n, float *out, int *not_ones, cudaStream_t s);
int run_eca_gate(egn_ctx *ctx, int level, int c, const float *t, const float *wk, int k, float *part, int slices, float *gate,
                 cudaStream_t s);
int run_eca_apply(egn_ctx *ctx, int level, int c, const float *t, const float *res, const float *gate, int relu, int res_split,
                  int out_split, float *out, cudaStream_t s);
int run_l2norm(egn_ctx *ctx, const float *x, int n, int c, float *out, cudaStream_t s);
int run_kp_sigma(egn_ctx *ctx, int level, const float *kp_raw, int kp_stride, const float *sg_raw, int sg_stride, int polar,
                 const float q[3], int ignore_offset, float *kp_out, float *sg_out, cudaStream_t s);
int pool_slices_for(egn_ctx *ctx, int level);

namespace {

// A feature map of the running forward.  split: stored in the pre-split format of tc_ptx.cuh (bf16 hi/lo pairs + a zero
// row at index n) because its main consumer is a tensor-core convolution that gathers from it.
struct Map {
  float *p = nullptr;
  bool split = false;
};

struct Fwd {
  egn_ctx *ctx;
  const egn_net *net;
  const float *wb;
  cudaStream_t s;
  const float *W(int64_t off) const { return off < 0 ? nullptr : wb + off; }
  float *alloc(size_t floats) { return (float *)ctx->feats.take(floats * 4); }
  // (n rows + the zero row of a pre-split map) x c
  float *alloc_map(size_t n, int c) { return alloc((n + 1) * (size_t)c); }

  bool tc_ok(const egn_layer &l, int ksize, int transposed) const {
    return ctx->use_tc && l.wtc >= 0 && sconv_tc_supported(ksize, transposed, l.cin, l.cout);
  }
  // may a map whose main consumer is layer l be stored pre-split?
  bool want_split(const egn_layer &l, int ksize, int transposed) const { return tc_ok(l, ksize, transposed); }

  // conv described by an egn_layer on the pyramid
  int layer(const egn_layer &l, int level_in, int ksize, int transposed, Map in, int relu, int accumulate, Map out) {
    const Pyramid &py = ctx->pyr;
    if (tc_ok(l, ksize, transposed))
      return run_conv_tc(ctx, level_in, ksize, transposed, l.cin, l.cout, in.p, wb + l.wtc, W(l.scale), W(l.shift), relu, accumulate, out.p, s,
                         in.split, out.split);
    EGN_CHECK(!out.split, EGN_ERR_STATE, "FP32 convolution asked to write a pre-split map");
    const float *src = in.p;
    if (in.split) {                                           // the FP32 CUDA-core kernels read fp32 only
      const size_t floats = (size_t)py.n[level_in] * l.cin;
      float *tmp = alloc(floats);
      EGN_CHECK(tmp != nullptr, EGN_ERR_STATE, "feature arena exhausted (format conversion)");
      EGN_TRY(run_presplit_to_f32(ctx, in.p, (int64_t)floats, tmp, s));
      src = tmp;
    }
    return run_conv(ctx, level_in, ksize, transposed, l.cin, l.cout, src, W(l.w), W(l.scale), W(l.shift), relu, accumulate, out.p, s);
  }
};

size_t head_floats(const Pyramid &py, const egn_head &h) {
  if (h.n_levels == 0) return 0;
  size_t f = 0;
  for (int L = h.levels[0]; L <= h.levels[h.n_levels - 1]; ++L) f += (size_t)py.n[L] * h.out_channels + 64;
  return f;
}

size_t plan_floats(const Pyramid &py, const egn_net &net) {
  size_t f = (size_t)py.n[0] + (size_t)(py.n[0] + 1) * net.conv0.cout + 256;
  for (int L = 1; L <= net.n_levels; ++L) {
    const size_t n = py.n[L] + 1;                          // + the zero row of pre-split maps
    f += n * net.down[L].cout + 3 * n * net.conv2[L].cout + (net.res[L].cin ? n * net.res[L].cout : 0) + 1024;
    f += (size_t)net.n_extra[L] * (3 * n * net.conv2[L].cout + ((size_t)kNumSMs * 8 + 2 * (size_t)py.n_batches) * net.conv2[L].cout + 256);   // blocks 1..
    f += 2 * n * net.conv2[L].cout;                        // fp32 copies for consumers that cannot read a pre-split map (rare)
    f += ((size_t)kNumSMs * 8 + 2 * (size_t)py.n_batches) * net.conv2[L].cout + 128;   // pooling partials (<= 8 CTAs per SM) + gates
  }
  for (int L = 1; L <= net.n_levels; ++L)                   // K-split partial tiles (3 splits, levels of <= 37 row tiles, EGN_KSPLIT=1)
    if (py.n[L] <= 37 * 128) f += 2 * 3 * (size_t)py.n[L] * net.conv2[L].cout + 64;
  f += head_floats(py, net.global_head) + head_floats(py, net.local_head);
  if (net.global_head.n_levels) {
    const size_t n = py.n[net.global_head.levels[0]];
    f += n * (net.global_mlp[0].cin ? net.global_mlp[0].cout + net.global_mlp[1].cout : 0) + 256;
    f += ((size_t)kNumSMs * 8 + (size_t)py.n_batches) * 512 + 256;
  }
  if (net.local_head.n_levels) {
    const size_t n = py.n[net.local_head.levels[0]];
    f += n * (net.desc_mlp[0].cout + net.desc_mlp[1].cout + net.kp_mlp[0].cout + net.kp_mlp[1].cout + net.sigma_mlp[0].cout +
              net.sigma_mlp[1].cout + net.kpsig_mlp[0].cout + net.kpsig_mlp[1].cout) + 2048;
  }
  return f + 4096;
}

// MinkHead.forward (models/minkgl.py:46-60): y = conv1x1[hi](x[hi]); for level hi-1..lo: y = tconv[level+1](y) (+ conv1x1[level](x[level]))
int run_head(Fwd &F, const egn_head &h, const Map x[], float **out_map) {
  const Pyramid &py = F.ctx->pyr;
  const int lo = h.levels[0], hi = h.levels[h.n_levels - 1];
  float *y = F.alloc((size_t)py.n[hi] * h.out_channels);
  EGN_CHECK(y != nullptr, EGN_ERR_STATE, "feature arena exhausted (head)");
  EGN_TRY(F.layer(h.conv1x1[hi], hi, 1, 0, x[hi], 0, 0, Map{y, false}));
  for (int level = hi - 1; level >= lo; --level) {
    float *y2 = F.alloc((size_t)py.n[level] * h.out_channels);
    EGN_CHECK(y2 != nullptr, EGN_ERR_STATE, "feature arena exhausted (head)");
    EGN_TRY(F.layer(h.tconv[level + 1], level + 1, 2, 1, Map{y, false}, 0, 0, Map{y2, false}));
    bool lateral = false;
    for (int i = 0; i < h.n_levels; ++i) lateral |= (h.levels[i] == level);
    if (lateral) EGN_TRY(F.layer(h.conv1x1[level], level, 1, 0, x[level], 0, 1, Map{y2, false}));
    y = y2;
  }
  *out_map = y;
  return EGN_OK;
}

}  // namespace

int forward(egn_ctx *ctx, const egn_net *net, const float *weights, const float *features, float *global_out,
            float *desc_out, float *kp_out, float *sigma_out, cudaStream_t s) {
  EGN_CHECK(ctx && net && weights, EGN_ERR_INVALID, "forward: null argument");
  Pyramid &py = ctx->pyr;
  EGN_CHECK(py.valid, EGN_ERR_STATE, "forward before coords_build");
  EGN_CHECK(net->n_levels >= 1 && net->n_levels < EGN_MAX_LEVELS && net->n_levels + 2 < P, EGN_ERR_INVALID, "forward: n_levels out of range");
  EGN_CHECK(net->conv0.cin == 1, EGN_ERR_INVALID, "forward: conv0 expects one input channel (models/model_factory.py:13)");
  const bool do_global = global_out != nullptr, do_local = desc_out || kp_out || sigma_out;
  EGN_CHECK(!do_global || net->global_head.n_levels > 0, EGN_ERR_INVALID, "forward: model has no global head");
  EGN_CHECK(!do_local || (net->local_head.n_levels > 0 && desc_out && kp_out && sigma_out), EGN_ERR_INVALID,
            "forward: local outputs need a local head and all three buffers");

  EGN_TRY(ctx->feats.reserve(plan_floats(py, *net) * 4, s));
  struct InForward {                                        // run_conv_tc takes its K-split temporaries from the feature arena
    egn_ctx *c;
    explicit InForward(egn_ctx *c_) : c(c_) { c->in_forward = true; }
    ~InForward() { c->in_forward = false; }
  } in_forward_guard(ctx);
  Fwd F{ctx, net, weights, s};
  Taps &tp = ctx->taps;
  tp = Taps();

  // ---- trunk (models/minkgl.py:136-153) ----
  float *f0 = F.alloc(py.n[0]);
  Map x0;
  x0.p = F.alloc_map(py.n[0], net->conv0.cout);
  x0.split = F.want_split(net->down[1], 2, 0);
  EGN_CHECK(f0 && x0.p, EGN_ERR_STATE, "feature arena exhausted (conv0)");
  int *not_ones = ctx->dev_counts + P + 3;
  EGN_TRY(run_gather_rows1(ctx, features, py.perm0, py.n[0], f0, not_ones, s));
  EGN_TRY(run_conv0(ctx, net->conv0_ksize, f0, F.W(net->conv0.w), F.W(net->conv0.scale), F.W(net->conv0.shift),
                    net->conv0.cout, 1, not_ones, x0.split, x0.p, s,
                    (ctx->use_tc && net->conv0.wtc >= 0 && net->conv0_ksize == 5) ? (const void *)(weights + net->conv0.wtc) : nullptr));
  tp.conv0 = x0.p;
  tp.conv0_split = x0.split;
  tp.c0 = net->conv0.cout;

  Map x[EGN_MAX_LEVELS];
  x[0] = x0;
  Map cur = x0;
  // ---- local head (models/minkgl.py:288-308): needs only trunk levels <= max(local levels), so it is enqueued on the
  //      context's second stream as soon as that level is done and overlaps the small upper trunk levels ----
  auto run_local = [&](cudaStream_t ls) -> int {
    Fwd F{ctx, net, weights, ls};
    const egn_head &h = net->local_head;
    float *lm = nullptr;
    EGN_TRY(run_head(F, h, x, &lm));
    const int lvl = h.levels[0];
    const size_t n = py.n[lvl];
    tp.lmap = lm; tp.c_l = h.out_channels; tp.lvl_l = lvl;
    float *d1 = F.alloc(n * net->desc_mlp[0].cout), *d2 = F.alloc(n * net->desc_mlp[1].cout);
    EGN_CHECK(d1 && d2, EGN_ERR_STATE, "feature arena exhausted (local mlps)");
    EGN_TRY(F.layer(net->desc_mlp[0], lvl, 1, 0, Map{lm, false}, 1, 0, Map{d1, false}));
    EGN_TRY(F.layer(net->desc_mlp[1], lvl, 1, 0, Map{d1, false}, 0, 0, Map{d2, false}));
    EGN_TRY(run_l2norm(ctx, d2, (int)n, net->desc_mlp[1].cout, desc_out, ls));
    if (net->kpsig_mlp[0].cin) {                            // fused regressors: one 64 -> 32+32 layer, one (32+32) -> 3+1 layer
      const int so = net->kpsig_mlp[1].cout;                // 3 + 1 outputs, zero-padded to a tensor-core width (32) by weights.py
      EGN_CHECK(so >= 4, EGN_ERR_INVALID, "fused keypoint/sigma regressor must end in 3+1 outputs");
      float *h1 = F.alloc(n * net->kpsig_mlp[0].cout), *o4 = F.alloc(n * (size_t)so);
      EGN_CHECK(h1 && o4, EGN_ERR_STATE, "feature arena exhausted (local mlps)");
      EGN_TRY(F.layer(net->kpsig_mlp[0], lvl, 1, 0, Map{lm, false}, 1, 0, Map{h1, false}));
      EGN_TRY(F.layer(net->kpsig_mlp[1], lvl, 1, 0, Map{h1, false}, 0, 0, Map{o4, false}));
      EGN_TRY(run_kp_sigma(ctx, lvl, o4, so, o4 + 3, so, net->polar, net->quant_step, net->ignore_keypoint_regressor, kp_out, sigma_out, ls));
    } else {
      float *k1 = F.alloc(n * net->kp_mlp[0].cout), *k2 = F.alloc(n * net->kp_mlp[1].cout);
      float *s1 = F.alloc(n * net->sigma_mlp[0].cout), *s2 = F.alloc(n * net->sigma_mlp[1].cout);
      EGN_CHECK(k1 && k2 && s1 && s2, EGN_ERR_STATE, "feature arena exhausted (local mlps)");
      EGN_CHECK(net->kp_mlp[1].cout == 3 && net->sigma_mlp[1].cout == 1, EGN_ERR_INVALID, "keypoint/sigma regressor shapes");
      EGN_TRY(F.layer(net->kp_mlp[0], lvl, 1, 0, Map{lm, false}, 1, 0, Map{k1, false}));
      EGN_TRY(F.layer(net->kp_mlp[1], lvl, 1, 0, Map{k1, false}, 0, 0, Map{k2, false}));
      EGN_TRY(F.layer(net->sigma_mlp[0], lvl, 1, 0, Map{lm, false}, 1, 0, Map{s1, false}));
      EGN_TRY(F.layer(net->sigma_mlp[1], lvl, 1, 0, Map{s1, false}, 0, 0, Map{s2, false}));
      EGN_TRY(run_kp_sigma(ctx, lvl, k2, 3, s2, 1, net->polar, net->quant_step, net->ignore_keypoint_regressor, kp_out, sigma_out, ls));
    }
    return EGN_OK;
  };
  bool local_forked = false;
  const int local_top = do_local ? net->local_head.levels[net->local_head.n_levels - 1] : -1;
  for (int L = 1; L <= net->n_levels; ++L) {
    const size_t n = py.n[L];
    const int c = net->conv2[L].cout;
    Map d, t1, t2, xo;
    d.p = F.alloc_map(n, net->down[L].cout); t1.p = F.alloc_map(n, c); t2.p = F.alloc_map(n, c); xo.p = F.alloc_map(n, c);
    EGN_CHECK(d.p && t1.p && t2.p && xo.p, EGN_ERR_STATE, "feature arena exhausted (level %d)", L);
    // pre-split where the producer can write it (a tensor-core convolution / the fused residual pass) and the main consumer gathers it
    d.split = F.tc_ok(net->down[L], 2, 0) && F.want_split(net->conv1[L], 3, 0);
    t1.split = F.tc_ok(net->conv1[L], 3, 0) && F.want_split(net->conv2[L], 3, 0);
    const int n_extra = net->n_extra[L];
    EGN_CHECK(n_extra >= 0 && n_extra <= EGN_MAX_EXTRA_BLOCKS, EGN_ERR_INVALID, "forward: %d extra blocks at level %d", n_extra, L);
    // the block output feeds the next block's 3x3x3 convolution (layers[L] > 1) or the next level's stride-2 convolution
    xo.split = n_extra > 0 ? F.want_split(net->xconv1[L][0], 3, 0) : (L < net->n_levels && F.want_split(net->down[L + 1], 2, 0));
    EGN_TRY(F.layer(net->down[L], L - 1, 2, 0, cur, 1, 0, d));              // convs[L] + bn[L] + relu
    EGN_TRY(F.layer(net->conv1[L], L, 3, 0, d, 1, 0, t1));                  // conv1 + norm1 + relu
    EGN_TRY(F.layer(net->conv2[L], L, 3, 0, t1, 0, 0, t2));                 // conv2 + norm2
    Map res = d;
    if (net->res[L].cin) {                                                  // downsample: 1x1 + bn
      Map r;
      r.p = F.alloc_map(n, c);
      EGN_CHECK(r.p != nullptr, EGN_ERR_STATE, "feature arena exhausted (residual)");
      EGN_TRY(F.layer(net->res[L], L, 1, 0, d, 0, 0, r));
      res = r;
    } else {
      EGN_CHECK(net->down[L].cout == c, EGN_ERR_INVALID, "identity residual with a channel change at level %d", L);
    }
    const float *gate = nullptr;
    if (net->eca_k[L] > 0) {                                                // ECALayer
      const int slices = pool_slices_for(ctx, L);
      float *part = F.alloc((size_t)py.n_batches * slices * c), *g = F.alloc((size_t)py.n_batches * c);
      EGN_CHECK(part && g, EGN_ERR_STATE, "feature arena exhausted (eca)");
      EGN_TRY(run_eca_gate(ctx, L, c, t2.p, F.W(net->eca_w[L]), net->eca_k[L], part, slices, g, s));
      gate = g;
    }
    EGN_TRY(run_eca_apply(ctx, L, c, t2.p, res.p, gate, 1, res.split, xo.split, xo.p, s));   // out = relu(eca(out) + residual)
    for (int j = 0; j < n_extra; ++j) {                                     // blocks 1.. : same channels, identity residual
      const egn_layer &c1 = net->xconv1[L][j], &c2 = net->xconv2[L][j];
      EGN_CHECK(c1.cin == c && c1.cout == c && c2.cin == c && c2.cout == c, EGN_ERR_INVALID, "extra block %d of level %d must keep %d channels", j + 1, L, c);
      Map u1, u2, xn;
      u1.p = F.alloc_map(n, c); u2.p = F.alloc_map(n, c); xn.p = F.alloc_map(n, c);
      EGN_CHECK(u1.p && u2.p && xn.p, EGN_ERR_STATE, "feature arena exhausted (level %d block %d)", L, j + 1);
      u1.split = F.tc_ok(c1, 3, 0) && F.want_split(c2, 3, 0);
      xn.split = j + 1 < n_extra ? F.want_split(net->xconv1[L][j + 1], 3, 0) : (L < net->n_levels && F.want_split(net->down[L + 1], 2, 0));
      EGN_TRY(F.layer(c1, L, 3, 0, xo, 1, 0, u1));
      EGN_TRY(F.layer(c2, L, 3, 0, u1, 0, 0, u2));
      const float *g2 = nullptr;
      if (net->xeca_k[L][j] > 0) {
        const int slices = pool_slices_for(ctx, L);
        float *part = F.alloc((size_t)py.n_batches * slices * c), *g = F.alloc((size_t)py.n_batches * c);
        EGN_CHECK(part && g, EGN_ERR_STATE, "feature arena exhausted (eca)");
        EGN_TRY(run_eca_gate(ctx, L, c, u2.p, F.W(net->xeca_w[L][j]), net->xeca_k[L][j], part, slices, g, s));
        g2 = g;
      }
      EGN_TRY(run_eca_apply(ctx, L, c, u2.p, xo.p, g2, 1, xo.split, xn.split, xn.p, s));
      xo = xn;
    }
    x[L] = xo;
    cur = xo;
    tp.down[L] = d.p; tp.c_down[L] = net->down[L].cout; tp.down_split[L] = d.split;
    tp.block[L] = xo.p; tp.c_block[L] = c; tp.block_split[L] = xo.split;
    if (L == local_top) {
      if (L < net->n_levels && ctx->aux && !ctx->prof.on) {               // fork: local head || levels L+1..n + global head
        EGN_CUDA(cudaEventRecord(ctx->ev_fork, s));
        EGN_CUDA(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
        EGN_TRY(run_local(ctx->aux));
        EGN_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux));
        local_forked = true;
      } else {
        EGN_TRY(run_local(s));
      }
    }
  }

  // ---- global head -> decoder -> pooling (models/minkgl.py:273-286) ----
  if (do_global) {
    const egn_head &h = net->global_head;
    float *gm = nullptr;
    EGN_TRY(run_head(F, h, x, &gm));
    const int lvl = h.levels[0];
    int c = h.out_channels;
    tp.gmap = gm; tp.c_g = c; tp.lvl_g = lvl;
    const float *pin = gm;
    if (net->global_mlp[0].cin) {
      float *m1 = F.alloc((size_t)py.n[lvl] * net->global_mlp[0].cout), *m2 = F.alloc((size_t)py.n[lvl] * net->global_mlp[1].cout);
      EGN_CHECK(m1 && m2, EGN_ERR_STATE, "feature arena exhausted (global mlp)");
      EGN_TRY(F.layer(net->global_mlp[0], lvl, 1, 0, Map{gm, false}, 1, 0, Map{m1, false}));
      EGN_TRY(F.layer(net->global_mlp[1], lvl, 1, 0, Map{m1, false}, 0, 0, Map{m2, false}));
      pin = m2;
      c = net->global_mlp[1].cout;
    }
    const int slices = pool_slices_for(ctx, lvl);
    float *part = F.alloc((size_t)py.n_batches * slices * c);
    EGN_CHECK(part != nullptr, EGN_ERR_STATE, "feature arena exhausted (pool)");
    const int mode = net->pool_method == 0 ? 1 : (net->pool_method == 1 ? 0 : 2);
    EGN_TRY(run_pool(ctx, lvl, c, pin, mode, net->gem_p, net->gem_eps, part, slices, global_out, s));
  }

  if (local_forked) EGN_CUDA(cudaStreamWaitEvent(s, ctx->ev_join, 0));   // join the local-head stream
  return EGN_OK;
}

int forward_tap(egn_ctx *ctx, int which, int level, float *out, cudaStream_t s) {
  EGN_CHECK(ctx && out && ctx->pyr.valid, EGN_ERR_STATE, "forward_tap: no forward to tap");
  const Pyramid &py = ctx->pyr;
  const Taps &tp = ctx->taps;
  const float *src = nullptr;
  size_t floats = 0;
  bool split = false;
  if (which == 0) { src = tp.conv0; floats = (size_t)py.n[0] * tp.c0; split = tp.conv0_split; }
  else if (which == 1 && level >= 1 && level < EGN_MAX_LEVELS) { src = tp.down[level]; floats = (size_t)py.n[level] * tp.c_down[level]; split = tp.down_split[level]; }
  else if (which == 2 && level >= 1 && level < EGN_MAX_LEVELS) { src = tp.block[level]; floats = (size_t)py.n[level] * tp.c_block[level]; split = tp.block_split[level]; }
  else if (which == 3) { src = tp.gmap; floats = (size_t)py.n[tp.lvl_g] * tp.c_g; }
  else if (which == 4) { src = tp.lmap; floats = (size_t)py.n[tp.lvl_l] * tp.c_l; }
  EGN_CHECK(src != nullptr, EGN_ERR_STATE, "forward_tap: tap %d/%d not available", which, level);
  if (split) return run_presplit_to_f32(ctx, src, (int64_t)floats, out, s);    // taps are always fp32
  EGN_CUDA(cudaMemcpyAsync(out, src, floats * 4, cudaMemcpyDeviceToDevice, s));
  return EGN_OK;
}

}  // namespace egn
