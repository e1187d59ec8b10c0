// Tensor-core sparse convolution: shape table, dispatch and the K-split finish kernel.
//
// The convolution kernel itself is k_sconv_ts (sconv_ts.cu: gathered A operand in tensor memory).  Its first generation,
// k_sconv_tc (A staged in shared memory as SWIZZLE_128B bf16 hi/lo images, SS-form tcgen05.mma), lived in this file until the
// weight image took the 32-channel K permutation of the 256-bit gathers; its measured limits are documented in
// profiles/r01_tc_kernel_stalls.md and profiles/r01_ts_kernel_analysis.md, its code is in the git history.
#include "ctx.cuh"
#include "tc_ptx.cuh"

namespace egn {

namespace tc {

// K-split finish: out = epilogue(sum_z partial[z]) in fixed z order (deterministic)
__global__ void k_splitk_finish(const float4 *__restrict__ part, int splits, int64_t n4 /* rows*C/4 */, int c4, const float *__restrict__ scale,
                                const float *__restrict__ shift, int relu, float4 *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = part[i];
    for (int z = 1; z < splits; ++z) { const float4 p = part[i + z * n4]; v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
    const int c = (int)(i % c4) * 4;
    if (scale) { v.x *= scale[c]; v.y *= scale[c + 1]; v.z *= scale[c + 2]; v.w *= scale[c + 3]; }
    if (shift) { v.x += shift[c]; v.y += shift[c + 1]; v.z += shift[c + 2]; v.w += shift[c + 3]; }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    out[i] = v;
  }
}

}  // namespace tc

bool sconv_tc_supported(int ksize, int transposed, int cin, int cout) {
  if (ksize == 1) return !transposed && ((cin == 32 && cout == 64) || (cin == 64 && (cout == 32 || cout == 64 || cout == 128)) ||
                                         (cin == 128 && (cout == 64 || cout == 128 || cout == 256)) || (cin == 256 && cout == 256));
  if (transposed) return ksize == 2 && cin == cout && (cin == 32 || cin == 64 || cin == 128);
  if (ksize == 3) return (cin == 32 && (cout == 32 || cout == 64)) || (cin == 64 && (cout == 64 || cout == 128)) || (cin == 128 && cout == 128);
  if (ksize == 2) return cin == cout && (cin == 32 || cin == 64 || cin == 128);
  return false;
}

// packed-weight bytes for a (ksize, cin, cout) convolution: n_chunks * 2 images * cout * 128
size_t sconv_tc_wpack_bytes(int ksize, int cin, int cout) {
  const int koff = ksize == 3 ? 27 : (ksize == 2 ? 8 : 1);
  return (size_t)((koff * cin + 63) / 64) * 2 * cout * 128;     // (a 256-channel row-wise layer: koff 1 x cin 256 == 2 x 128)
}

// sconv_ts.cu: the TMEM-resident-A kernels
int launch_conv_ts(egn_ctx *ctx, int koff, int cin, int cout_cta, const tcx::Args &a, const char *name, double bytes, double flops,
                   cudaStream_t s);

int run_conv_tc(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout, const float *in, const void *wpack,
                const float *scale, const float *shift, int relu, int accumulate, float *out, cudaStream_t s, int in_split,
                int out_split) {
  const Pyramid &py = ctx->pyr;
  EGN_CHECK(!(out_split && accumulate), EGN_ERR_INVALID, "accumulate into a pre-split map is not supported");
  EGN_CHECK(py.valid, EGN_ERR_STATE, "conv before coords_build");
  EGN_CHECK(sconv_tc_supported(ksize, transposed, cin, cout), EGN_ERR_INVALID, "tensor-core conv: unsupported shape k=%d %d->%d", ksize, cin, cout);
  EGN_CHECK(((uintptr_t)wpack & 15) == 0 && ((uintptr_t)in & 31) == 0 && ((uintptr_t)out & 15) == 0, EGN_ERR_INVALID,
            "tensor-core conv: the input map must be 32-byte aligned (256-bit gathers), weights and output 16-byte aligned");
  tcx::Args a = {};
  a.in = in; a.out = out; a.wpack = (const uint8_t *)wpack; a.scale = scale; a.shift = shift; a.relu = relu; a.accumulate = accumulate; a.cout_total = cout; a.ksplit = 1; a.hint_producer = ctx->hint_producer; a.hint_single = ctx->hint_single; a.trace = (long long *)ctx->trace; a.in_split = in_split; a.out_split = out_split; a.out_zero_row = out_split;
  long long pairs;
  char name[48];
  if (ksize == 1) {
    EGN_CHECK(level_in >= 0 && level_in < P, EGN_ERR_INVALID, "conv k=1: bad level");
    a.mode = 0; a.n_out = py.n[level_in]; a.zero_row = py.n[level_in];
    pairs = a.n_out;
    snprintf(name, sizeof(name), "tc_rowmm_c%d_%d", cin, cout);
  } else if (ksize == 3) {
    EGN_CHECK(level_in >= 1 && level_in < P, EGN_ERR_INVALID, "conv k=3: bad level");
    a.mode = 1; a.n_out = py.n[level_in]; a.nbr = py.nbr[level_in]; a.zero_row = py.n[level_in];
    if (py.ordered) a.order = py.ord27[level_in];
    pairs = py.pairs27[level_in];
    snprintf(name, sizeof(name), "tc_conv3x3x3_c%d_%d", cin, cout);
  } else if (transposed) {
    EGN_CHECK(level_in >= 1 && level_in < P, EGN_ERR_INVALID, "transposed conv: bad level");
    a.mode = 3; a.n_out = py.n[level_in - 1]; a.up = py.up[level_in - 1]; a.keys = py.keys[level_in - 1]; a.zero_row = py.n[level_in];
    if (py.ordered) a.order = py.ordt[level_in - 1];
    pairs = a.n_out;
    snprintf(name, sizeof(name), "tc_tconv2x2x2s2_c%d_%d", cin, cout);
  } else {
    EGN_CHECK(level_in >= 0 && level_in + 1 < P, EGN_ERR_INVALID, "conv k=2: bad level");
    a.mode = 2; a.n_out = py.n[level_in + 1]; a.cstart = py.cstart[level_in + 1]; a.cmask = py.cmask[level_in + 1]; a.zero_row = py.n[level_in];
    if (py.ordered) a.order = py.ordc[level_in + 1];
    pairs = py.n[level_in];
    snprintf(name, sizeof(name), "tc_conv2x2x2s2_c%d_%d", cin, cout);
  }
  if (a.n_out == 0) return EGN_OK;
  const int K = ksize == 3 ? 27 : (ksize == 2 ? 8 : 1);
  const double bytes = ksize == 1 ? (double)pairs * (cin + cout) * 4
                                  : (double)pairs * (cin + cout) * 4 + (double)pairs * 8 + (double)K * cin * cout * 4;
  const double flops = 2.0 * pairs * cin * cout;
  // small 128-channel levels: split the output channels over 2 or 4 CTAs per row tile (each streams only its share of
  // the weight chunks) so that a 12..148-tile level spreads over the whole GPU
  // and the 27 offsets over 3 CTAs (K-split): the 54-chunk serial chain per CTA becomes 18.  The raw partials go to the
  // scratch arena and k_splitk_finish adds them in fixed order and applies the BatchNorm/ReLU epilogue (deterministic).
  if (cin == 128 && cout == 128 && (ksize == 3 || ksize == 2) && !accumulate) {
    const int tiles = (int)div_up(a.n_out, tcx::kRows);
    if (tiles <= ctx->nsplit_max) {   // N = 32 -> <= 148 CTAs for <= 37 tiles; N = 64 -> <= 148 CTAs for <= 74 tiles (2 CTAs/SM beyond)
      const int splits = (!ctx->ksplit || out_split) ? 1 : (ksize == 3 ? (tiles <= 37 ? 3 : 1) : 1);
      tcx::Args b = a;
      float *part = nullptr;
      if (splits > 1) {
        const size_t bytes_part = (size_t)splits * a.n_out * cout * 4;
        // raw partial tiles: from the feature arena inside a forward (planned by forward.cu), from the scratch arena for a
        // single-operator call (dead sort buffers; reuse on the same stream is stream-ordered) - no cudaMalloc on this path
        if (ctx->in_forward) {
          part = (float *)ctx->feats.take(bytes_part);
        } else {
          EGN_TRY(ctx->scratch.reserve(bytes_part + 4096, s));
          part = (float *)ctx->scratch.take(bytes_part);
        }
        EGN_CHECK(part != nullptr, EGN_ERR_STATE, "arena exhausted (K-split partial tiles)");
        b.out = part; b.scale = nullptr; b.shift = nullptr; b.relu = 0; b.ksplit = splits;
      }
      EGN_TRY(launch_conv_ts(ctx, K, 128, tiles <= 37 ? 32 : 64, b, name, bytes, flops, s));
      if (splits > 1) {
        const int64_t n4 = (int64_t)a.n_out * cout / 4;
        EGN_LAUNCH(ctx, "splitk_finish", (double)(splits + 1) * a.n_out * cout * 4, 0, s,
                   tc::k_splitk_finish<<<grid_for(n4, 256), 256, 0, s>>>((const float4 *)part, splits, n4, cout / 4, scale, shift, relu,
                                                                          (float4 *)out));
        EGN_CUDA(cudaGetLastError());
      }
      return EGN_OK;
    }
  }
  if (ksize == 1 && (cin > 128 || cout > 128)) {
    // wide row-wise layers (the global descriptor decoder 128 -> 192 -> 256, padded to 256): output channels N-split over
    // grid.y CTAs of 128, a 256-channel input as two 128-channel "offsets" of the same row
    EGN_CHECK(!in_split && !out_split, EGN_ERR_INVALID, "wide row-wise layers read and write fp32 maps");
    return launch_conv_ts(ctx, cin / 128, 128, 128, a, name, bytes, flops, s);
  }
  return launch_conv_ts(ctx, K, cin, cout, a, name, bytes, flops, s);
}

}  // namespace egn
