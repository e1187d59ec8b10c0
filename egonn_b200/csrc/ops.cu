// Feature operators of the engine on the context's coordinate maps (FP32 CUDA-core path).
//
// Replaces (reference file:line under /root/reference; MinkowskiEngine semantics per SURVEY.md Appendix A):
//   ME.MinkowskiConvolution k=5 (conv0)        models/minkgl.py:100-102,140-142   -> k_conv0
//   ME.MinkowskiConvolution k=3 / k=2 s2 / k=1 models/minkgl.py:104-107,124-126, layers/eca_block.py:59-64 -> k_sconv
//   ME.MinkowskiConvolutionTranspose k=2 s2    models/minkgl.py:39,52-53          -> k_sconv (parent gather)
//   ME.MinkowskiBatchNorm (eval) / ReLU / Linear bias                              -> fused epilogue scale/shift/relu
//   ME.MinkowskiGlobalPooling, BroadcastMultiplication, ECALayer  layers/eca_block.py:21-36 -> k_pool_partial, k_eca_gate, k_eca_apply
//   GeM / SPoC / MAC                           layers/pooling.py:46-86            -> k_pool_partial + k_pool_final
//
// All convolutions are OUTPUT-STATIONARY: one CTA owns a tile of output rows, walks the kernel offsets,
// gathers the input rows named by the neighbour table (absent -> zeros) and accumulates in registers; the
// output row is written exactly once with the BatchNorm/ReLU/residual epilogue applied - no atomics, no
// separate scatter pass, deterministic.
#include <algorithm>
#include <type_traits>

#include "ctx.cuh"
#include "tc_ptx.cuh"

namespace egn {

using tcx::presplit_pack;
using tcx::presplit_unpack;

// ------------------------------------------------------------------------------------------------------
// generic gathered convolution
// ------------------------------------------------------------------------------------------------------
enum GatherMode { G_IDENTITY = 0, G_NBR27 = 1, G_CHILD8 = 2, G_PARENT = 3 };

struct ConvArgs {
  const float *in;
  float *out;
  const float *w;      // (K, cin, cout)
  const float *scale;  // (cout) or null
  const float *shift;  // (cout) or null
  int n_out, cin, cout, K, relu, accumulate, mode;
  const int *nbr;          // G_NBR27: (n_out,27)
  const int *cstart;       // G_CHILD8: per output row, first child (input row)
  const uint32_t *cmask;   // G_CHILD8: per output row, child occupancy
  const int *up;           // G_PARENT: per output (fine) row, parent (input) row
  const uint64_t *keys;    // G_PARENT: per output row key (code = key & 7)
};

constexpr int TM = 64;   // output rows per CTA
constexpr int KC = 32;   // input-channel chunk
constexpr int APAD = 4;

template <int TN>
__global__ void __launch_bounds__(256) k_sconv(ConvArgs a) {
  constexpr int RN = TN / 16;
  __shared__ int s_idx[TM * 27];
  __shared__ __align__(16) float s_a[TM][KC + APAD];
  __shared__ __align__(16) float s_w[KC][TN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int K = a.K;

  // resolve the input row of every (tile row, offset)
  for (int t = tid; t < TM * K; t += 256) {
    const int r = t / K, k = t % K, row = row0 + r;
    int src = -1;
    if (row < a.n_out) {
      if (a.mode == G_IDENTITY) src = row;
      else if (a.mode == G_NBR27) src = a.nbr[(int64_t)row * 27 + k];
      else if (a.mode == G_CHILD8) {
        const uint32_t m = a.cmask[row];
        if ((m >> k) & 1u) src = a.cstart[row] + __popc(m & ((1u << k) - 1u));
      } else {
        if ((int)(a.keys[row] & 7ull) == k) src = a.up[row];
      }
    }
    s_idx[r * K + k] = src;
  }
  __syncthreads();

  float acc[4][RN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

  const bool vec_a = (a.cin & 3) == 0;
  const bool vec_w = (a.cout & 3) == 0 && n0 + TN <= a.cout;
  for (int k = 0; k < K; ++k) {
    const int present = (tid < TM) && (s_idx[tid * K + k] >= 0);
    if (!__syncthreads_or(present)) continue;
    const float *wk = a.w + (size_t)k * a.cin * a.cout;
    for (int kc0 = 0; kc0 < a.cin; kc0 += KC) {
      // gather A chunk: 8 threads per row, 4 floats each
#pragma unroll
      for (int pass = 0; pass < TM / 32; ++pass) {
        const int r = pass * 32 + (tid >> 3), c = (tid & 7) * 4;
        const int src = s_idx[r * K + k];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src >= 0) {
          const float *p = a.in + (size_t)src * a.cin + kc0 + c;
          if (vec_a && kc0 + c + 3 < a.cin) v = *(const float4 *)p;
          else {
            if (kc0 + c < a.cin) v.x = p[0];
            if (kc0 + c + 1 < a.cin) v.y = p[1];
            if (kc0 + c + 2 < a.cin) v.z = p[2];
            if (kc0 + c + 3 < a.cin) v.w = p[3];
          }
        }
        *(float4 *)&s_a[r][c] = v;
      }
      // W chunk (KC x TN)
      for (int t = tid; t < KC * TN / 4; t += 256) {
        const int kk = t / (TN / 4), c = (t % (TN / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kc0 + kk < a.cin) {
          const float *p = wk + (size_t)(kc0 + kk) * a.cout + n0 + c;
          if (vec_w) v = *(const float4 *)p;
          else {
            if (n0 + c < a.cout) v.x = p[0];
            if (n0 + c + 1 < a.cout) v.y = p[1];
            if (n0 + c + 2 < a.cout) v.z = p[2];
            if (n0 + c + 3 < a.cout) v.w = p[3];
          }
        }
        *(float4 *)&s_w[kk][c] = v;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < KC; ++kk) {
        float av[4], wv[RN];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = s_a[ty + 16 * i][kk];
#pragma unroll
        for (int j = 0; j < RN; ++j) wv[j] = s_w[kk][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int j = 0; j < RN; ++j) {
    const int col = n0 + tx + 16 * j;
    if (col >= a.cout) continue;
    const float sc = a.scale ? a.scale[col] : 1.f, sh = a.shift ? a.shift[col] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = row0 + ty + 16 * i;
      if (row >= a.n_out) continue;
      float v = acc[i][j] * sc + sh;
      if (a.relu) v = fmaxf(v, 0.f);
      float *o = a.out + (size_t)row * a.cout + col;
      if (a.accumulate) v += *o;
      *o = v;
    }
  }
}

// pairs = number of (in,out) row pairs the convolution touches (SURVEY 8d byte model)
static int launch_sconv(egn_ctx *ctx, const ConvArgs &a, long long pairs, cudaStream_t s) {
  if (a.n_out <= 0) return EGN_OK;
  const int gx = (int)div_up(a.n_out, TM);
  char name[48];
  static const char *kind[] = {"rowmm", "conv3x3x3", "conv2x2x2s2", "tconv2x2x2s2"};
  snprintf(name, sizeof(name), "%s_c%d_%d", kind[a.mode], a.cin, a.cout);
  const double bytes = a.mode == G_IDENTITY ? (double)pairs * (a.cin + a.cout) * 4
                                            : (double)pairs * (a.cin + a.cout) * 4 + (double)pairs * 8 + (double)a.K * a.cin * a.cout * 4;
  const double flops = 2.0 * pairs * a.cin * a.cout;
  if (a.cout <= 32) EGN_LAUNCH(ctx, name, bytes, flops, s, k_sconv<32><<<dim3(gx, 1), 256, 0, s>>>(a));
  else if (a.cout <= 64) EGN_LAUNCH(ctx, name, bytes, flops, s, k_sconv<64><<<dim3(gx, 1), 256, 0, s>>>(a));
  else EGN_LAUNCH(ctx, name, bytes, flops, s, k_sconv<128><<<dim3(gx, (int)div_up(a.cout, 128)), 256, 0, s>>>(a));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

// ------------------------------------------------------------------------------------------------------
// conv0: 5x5x5, Cin = 1 -> out[o] = sum_k f[o + d_k] * W[k,:]: a masked sum of kernel rows, not a GEMM.
// One warp per level-0 row.  The 5^3 window of a voxel lies inside the 3x3x3 block of level-2 cells (4^3
// voxels each) around its own cell; presence = one bit of that cell's 64-bit occupancy word, row = first
// row of the cell + popcount of the lower bits.  No hash probes: 27 coalesced table reads per row.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t morton6(int x, int y, int z) {  // x,y,z in [0,4)
  return (uint32_t)((x & 1) | ((y & 1) << 1) | ((z & 1) << 2) | ((x & 2) << 2) | ((y & 2) << 3) | ((z & 2) << 4));
}

// Lane = output row.  Phase 1 (divergent, light): every lane walks the <= 8 level-2 cells its window touches,
// ANDs the cell occupancy with the window box (64-bit masks from a small LUT) and appends (offset index t,
// feature row) of every present neighbour to its column of a shared-memory list.  Phase 2 (converged, heavy):
// the warp loops to the longest list; each lane does 32 FMAs per neighbour with the kernel row W[t,:] read as
// 8 conflict-free LDS.128 (row stride 36 floats).  Work is proportional to the PRESENT pairs (~18 of 125).
constexpr int kC0WStride = 36;
// ONES = every input feature is 1.0f (what every EgoNN caller feeds): the list holds only the 7-bit offset index
// (1 byte per pair -> 8 warps per CTA, 4 CTAs per SM) and the FMA degenerates to an add.  Both variants are launched;
// the device-side flag written by the feature gather decides which one does the work (the other exits at once).
template <bool ONES> struct C0 { static constexpr int kWarps = ONES ? 8 : 4; static constexpr int kCtas = ONES ? 4 : 2; };

template <int KS, bool ONES>
__global__ void __launch_bounds__(C0<ONES>::kWarps * 32, C0<ONES>::kCtas) k_conv0(const float *__restrict__ f0 /* (n0) canonical order */,
                                                            const uint64_t *__restrict__ keys0, const int *__restrict__ up0,
                                                            const int *__restrict__ up1, const int *__restrict__ nbr2,
                                                            const uint64_t *__restrict__ mask64, const int *__restrict__ first0, int n0,
                                                            const float *__restrict__ w /* (KS^3,1,32) */, const float *__restrict__ scale,
                                                            const float *__restrict__ shift, int relu, const int *__restrict__ not_ones,
                                                            int out_split, float *__restrict__ out) {
  constexpr int KV = KS * KS * KS, R = KS / 2, COUT = 32, kC0Warps = C0<ONES>::kWarps;
  using entry_t = typename std::conditional<ONES, uint8_t, uint32_t>::type;
  const bool all_ones = not_ones != nullptr && *not_ones == 0;
  if (all_ones != ONES) return;
  if (out_split && blockIdx.x == 0 && threadIdx.x < COUT) out[(size_t)n0 * COUT + threadIdx.x] = 0.f;   // the zero row of a pre-split map
  extern __shared__ __align__(16) uint8_t s_raw[];
  float *s_w = (float *)s_raw;                                           // [KV][36]
  unsigned long long *s_box = (unsigned long long *)(s_w + KV * kC0WStride + 4);   // [axis 3][l 4][delta 3]
  entry_t *s_list = (entry_t *)(s_box + 36);                             // [warp][KV][32]
  for (int t = threadIdx.x; t < KV * COUT; t += blockDim.x) s_w[(t / COUT) * kC0WStride + (t % COUT)] = w[t];
  if (threadIdx.x < 36) {
    // 64-bit set of Morton-6 codes whose `axis` coordinate lies in window(l) /\ cell(delta)
    const int axis = threadIdx.x / 12, l = (threadIdx.x / 3) % 4, d = threadIdx.x % 3 - 1;
    const unsigned long long b0[3] = {0xAAAAAAAAAAAAAAAAull, 0xCCCCCCCCCCCCCCCCull, 0xF0F0F0F0F0F0F0F0ull};   // low coordinate bit = 1
    const unsigned long long b1[3] = {0xFF00FF00FF00FF00ull, 0xFFFF0000FFFF0000ull, 0xFFFFFFFF00000000ull};   // high coordinate bit = 1
    unsigned long long m = 0ull;
    for (int p = 0; p < 4; ++p) {
      const int g = p + 4 * d;                                           // coordinate relative to the own cell origin
      if (g >= l - R && g <= l + R)
        m |= ((p & 1) ? b0[axis] : ~b0[axis]) & ((p & 2) ? b1[axis] : ~b1[axis]);
    }
    s_box[threadIdx.x] = m;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  entry_t *list = s_list + (size_t)warp * KV * 32;
  const int nwarps = gridDim.x * kC0Warps;
  for (int base = (blockIdx.x * kC0Warps + warp) * 32; base < n0; base += nwarps * 32) {
    const int r = base + lane;
    int cnt = 0;
    if (r < n0) {
      const uint32_t m = (uint32_t)(keys0[r] & 63ull);
      const int lx = (m & 1) | ((m >> 2) & 2), ly = ((m >> 1) & 1) | ((m >> 3) & 2), lz = ((m >> 2) & 1) | ((m >> 4) & 2);
      const int cell = up1[up0[r]];
      const int *nb = nbr2 + (int64_t)cell * 27;
      // the window touches at most two cells per axis: delta d0 = floor((l-R)/4) and d1 = floor((l+R)/4)
      const int dx0 = (lx - R) >> 2, dx1 = (lx + R) >> 2, dy0 = (ly - R) >> 2, dy1 = (ly + R) >> 2, dz0 = (lz - R) >> 2,
                dz1 = (lz + R) >> 2;
      int q[8];
      unsigned long long box[8], occ[8];
      int fb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {          // 8 independent loads, then 16 more: two latencies instead of 24
        const int dx = (i & 1) ? dx1 : dx0, dy = (i & 2) ? dy1 : dy0, dz = (i & 4) ? dz1 : dz0;
        const bool dup = ((i & 1) && dx1 == dx0) || ((i & 2) && dy1 == dy0) || ((i & 4) && dz1 == dz0);
        box[i] = dup ? 0ull : (s_box[lx * 3 + dx + 1] & s_box[12 + ly * 3 + dy + 1] & s_box[24 + lz * 3 + dz + 1]);
        q[i] = box[i] ? nb[(dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)] : -1;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        occ[i] = q[i] >= 0 ? mask64[q[i]] : 0ull;
        fb[i] = q[i] >= 0 ? first0[q[i]] : 0;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int dx = (i & 1) ? dx1 : dx0, dy = (i & 2) ? dy1 : dy0, dz = (i & 4) ? dz1 : dz0;
        unsigned long long pm = occ[i] & box[i];
        while (pm) {
          const int b = __ffsll((long long)pm) - 1;
          pm &= pm - 1;
          const int x = (b & 1) | ((b >> 2) & 2), y = ((b >> 1) & 1) | ((b >> 3) & 2), z = ((b >> 2) & 1) | ((b >> 4) & 2);
          const int t = (x + 4 * dx - lx + R) + KS * ((y + 4 * dy - ly + R) + KS * (z + 4 * dz - lz + R));
          const int frow = fb[i] + __popcll(occ[i] & ((1ull << b) - 1ull));
          list[cnt * 32 + lane] = ONES ? (entry_t)t : (entry_t)((uint32_t)t | ((uint32_t)frow << 7));
          ++cnt;
        }
      }
    }
    __syncwarp();
    const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
    for (int i = 0; i < maxcnt; ++i) {
      if (i < cnt) {
        const uint32_t e = list[i * 32 + lane];
        const float f = ONES ? 1.f : f0[e >> 7];
        const float4 *wr = (const float4 *)(s_w + (e & 127u) * kC0WStride);
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
          const float4 ww = wr[c4];
          acc[c4 * 4 + 0] = fmaf(f, ww.x, acc[c4 * 4 + 0]);
          acc[c4 * 4 + 1] = fmaf(f, ww.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(f, ww.z, acc[c4 * 4 + 2]);
          acc[c4 * 4 + 3] = fmaf(f, ww.w, acc[c4 * 4 + 3]);
        }
      }
    }
    if (r < n0) {
      float4 *o = (float4 *)(out + (size_t)r * COUT);
#pragma unroll
      for (int c4 = 0; c4 < COUT / 4; ++c4) {
        float4 y;
        float *yy = (float *)&y;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c4 * 4 + e;
          float v = acc[c] * (scale ? __ldg(scale + c) : 1.f) + (shift ? __ldg(shift + c) : 0.f);
          if (relu) v = fmaxf(v, 0.f);
          yy[e] = v;
        }
        if (out_split) ((uint4 *)o)[c4] = presplit_pack(y); else o[c4] = y;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------
// row-wise / per-cloud operators
// ------------------------------------------------------------------------------------------------------
// canonical-order copy of the (n,1) input features; also records whether any feature differs from 1.0f
__global__ void k_gather_rows1(const float *__restrict__ in, const int *__restrict__ perm, int n, float *__restrict__ out,
                               int *__restrict__ not_ones) {
  int bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = in[perm[i]];
    out[i] = v;
    bad |= (v != 1.0f);
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicOr(not_ones, 1);
}

// per-cloud column reduction, deterministic two-stage: block (b, s) reduces slice s of cloud b's rows.
// mode 0: sum x ; 1: sum clamp(x, eps)^p (GeM) ; 2: max x
// CTA (b, s): 256 threads = (256 / c4) row lanes x c4 float4 column groups; every row lane walks its rows of the slice
// with 16-byte loads, the row lanes are then added in fixed order through shared memory (deterministic).
// Tail (single-launch pooling): the LAST CTA of cloud b to finish (device counter, self-resetting) adds the cloud's slice
// partials in fixed order and applies the final operator - tail 1: mean / GeM root / max -> out (B,c); tail 2: the ECA gate
// sigmoid(Conv1d_k(mean)) -> out (B,c).  Which CTA is last varies, what it computes does not: deterministic.
struct PoolTail {
  int kind;            // 0 none (partials only), 1 pooled value, 2 ECA gate
  int *counter;        // [B], zero between launches
  float *out;          // (B, c)
  const float *wk;     // ECA Conv1d taps
  int k;
};
__device__ __forceinline__ void reduce_slices(const float *__restrict__ part, int b, int c, int slices, bool is_max, float *s_tot);

__global__ void __launch_bounds__(256) k_pool_partial(const float *__restrict__ x, const int *__restrict__ boff, int c, int slices, int mode,
                                                      float p, float eps, float *__restrict__ part /* (B, slices, c) */, PoolTail tail) {
  __shared__ float4 s_acc[256];
  __shared__ float s_mean[256];
  __shared__ int s_last;
  const int b = blockIdx.x, s = blockIdx.y;
  const int r0 = boff[b], r1 = boff[b + 1];
  const int len = r1 - r0;
  const int a0 = r0 + (int)(((int64_t)len * s) / slices), a1 = r0 + (int)(((int64_t)len * (s + 1)) / slices);
  const int c4 = c >> 2;                                    // c is a multiple of 4, c4 <= 256 (checked by the launcher)
  const int lanes = 256 / c4, q = threadIdx.x % c4, rl = threadIdx.x / c4;
  const float init = mode == 2 ? -INFINITY : 0.f;
  float4 acc = make_float4(init, init, init, init);
  if (rl < lanes) {
    auto fold = [&](const float4 v) {
      if (mode == 0) { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      else if (mode == 1) {
        acc.x += powf(fmaxf(v.x, eps), p); acc.y += powf(fmaxf(v.y, eps), p);
        acc.z += powf(fmaxf(v.z, eps), p); acc.w += powf(fmaxf(v.w, eps), p);
      } else { acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w); }
    };
    int r = a0 + rl;
    for (; r + 3 * lanes < a1; r += 4 * lanes) {           // four rows in flight per thread (same summation order as one by one)
      const float4 v0 = *(const float4 *)(x + (size_t)r * c + 4 * q), v1 = *(const float4 *)(x + (size_t)(r + lanes) * c + 4 * q);
      const float4 v2 = *(const float4 *)(x + (size_t)(r + 2 * lanes) * c + 4 * q), v3 = *(const float4 *)(x + (size_t)(r + 3 * lanes) * c + 4 * q);
      fold(v0); fold(v1); fold(v2); fold(v3);
    }
    for (; r < a1; r += lanes) fold(*(const float4 *)(x + (size_t)r * c + 4 * q));
  }
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < c4) {
    float4 t = s_acc[threadIdx.x];
    for (int l = 1; l < lanes; ++l) {
      const float4 v = s_acc[l * c4 + threadIdx.x];
      if (mode == 2) { t.x = fmaxf(t.x, v.x); t.y = fmaxf(t.y, v.y); t.z = fmaxf(t.z, v.z); t.w = fmaxf(t.w, v.w); }
      else { t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    }
    *(float4 *)(part + ((size_t)b * slices + s) * c + 4 * threadIdx.x) = t;
  }
  if (tail.kind == 0) return;
  __threadfence();                                          // this CTA's partial is visible device-wide ...
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(tail.counter + b, 1);        // ... before it is counted
    s_last = prev == slices - 1;
    if (s_last) tail.counter[b] = 0;                        // self-resetting: zero again for the next launch on this stream
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float *s_tot = (float *)s_acc;                            // [256] (s_acc is dead)
  reduce_slices(part, b, c, slices, mode == 2, s_tot);
  if (tail.kind == 1) {
    if ((int)threadIdx.x < c) {
      const float acc = s_tot[threadIdx.x];
      float r;
      if (len == 0) r = 0.f;
      else if (mode == 0) r = acc / (float)len;
      else if (mode == 1) r = powf(acc / (float)len, 1.0f / p);
      else r = acc;
      tail.out[(size_t)b * c + threadIdx.x] = r;
    }
    return;
  }
  // ECA gate (layers/eca_block.py:21-31): mean -> Conv1d(1,1,k, zero pad, no bias) over channels -> sigmoid
  if ((int)threadIdx.x < c) s_mean[threadIdx.x] = len ? s_tot[threadIdx.x] / (float)len : 0.f;
  __syncthreads();
  const int pad = (tail.k - 1) / 2;
  if ((int)threadIdx.x < c) {
    const int ch = threadIdx.x;
    float y = 0.f;
    for (int j = 0; j < tail.k; ++j) {
      const int cc = ch + j - pad;
      if (cc >= 0 && cc < c) y = fmaf(tail.wk[j], s_mean[cc], y);
    }
    tail.out[(size_t)b * c + ch] = 1.f / (1.f + expf(-y));
  }
}
// sum (or max) of the per-slice partials of cloud b (c <= 256 channels): 256 threads = (256 / c) slice lanes x c channels, the
// lanes are then combined in fixed order through shared memory (deterministic); result in s_tot[ch]
__device__ __forceinline__ void reduce_slices(const float *__restrict__ part, int b, int c, int slices, bool is_max, float *s_tot /* [256] */) {
  const int lanes = 256 / c;
  const int ch = (int)threadIdx.x % c, l = (int)threadIdx.x / c;
  float acc = is_max ? -INFINITY : 0.f;
  if (l < lanes)
    for (int s = l; s < slices; s += lanes) {
      const float v = __ldcg(part + ((size_t)b * slices + s) * c + ch);   // L2: other CTAs of this launch wrote it (fused tail)
      acc = is_max ? fmaxf(acc, v) : acc + v;
    }
  s_tot[threadIdx.x] = acc;
  __syncthreads();
  if (l == 0) {
    for (int j = 1; j < lanes; ++j) {
      const float v = s_tot[j * c + ch];
      acc = is_max ? fmaxf(acc, v) : acc + v;
    }
  }
  __syncthreads();
  if (l == 0) s_tot[ch] = acc;
  __syncthreads();
}
__global__ void __launch_bounds__(256) k_pool_final(const float *__restrict__ part, const int *__restrict__ boff, int c, int slices, int mode,
                                                    float p, float *__restrict__ out /* (B,c) */) {
  __shared__ float s_tot[256];
  const int b = blockIdx.x;
  const int len = boff[b + 1] - boff[b];
  if (c <= 256) {
    reduce_slices(part, b, c, slices, mode == 2, s_tot);
    if ((int)threadIdx.x < c) {
      const float acc = s_tot[threadIdx.x];
      float r;
      if (len == 0) r = 0.f;
      else if (mode == 0) r = acc / (float)len;
      else if (mode == 1) r = powf(acc / (float)len, 1.0f / p);
      else r = acc;
      out[(size_t)b * c + threadIdx.x] = r;
    }
    return;
  }
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float acc = mode == 2 ? -INFINITY : 0.f;
    for (int s = 0; s < slices; ++s) {
      const float v = part[((size_t)b * slices + s) * c + ch];
      acc = mode == 2 ? fmaxf(acc, v) : acc + v;
    }
    float r;
    if (len == 0) r = 0.f;
    else if (mode == 0) r = acc / (float)len;
    else if (mode == 1) r = powf(acc / (float)len, 1.0f / p);
    else r = acc;
    out[(size_t)b * c + ch] = r;
  }
}
// ECA gate (layers/eca_block.py:21-31): mean -> Conv1d(1,1,k, zero pad, no bias) over channels -> sigmoid   (c <= 256)
__global__ void __launch_bounds__(256) k_eca_gate(const float *__restrict__ part, const int *__restrict__ boff, int c, int slices,
                                                  const float *__restrict__ wk, int k, float *__restrict__ gate /* (B,c) */) {
  __shared__ float s_tot[256];
  __shared__ float s_mean[256];
  const int b = blockIdx.x;
  const int len = boff[b + 1] - boff[b];
  reduce_slices(part, b, c, slices, false, s_tot);
  if ((int)threadIdx.x < c) s_mean[threadIdx.x] = len ? s_tot[threadIdx.x] / (float)len : 0.f;
  __syncthreads();
  const int pad = (k - 1) / 2;
  if ((int)threadIdx.x < c) {
    const int ch = threadIdx.x;
    float y = 0.f;
    for (int j = 0; j < k; ++j) {
      const int cc = ch + j - pad;
      if (cc >= 0 && cc < c) y = fmaf(wk[j], s_mean[cc], y);
    }
    gate[(size_t)b * c + ch] = 1.f / (1.f + expf(-y));
  }
}
// out = relu(t * gate[batch(row)] + res)   (layers/eca_block.py:36,70-71); gate == null -> plain BasicBlock.
// res_split / out_split: the residual / the output is a pre-split map (tc_ptx.cuh); out_split also writes the zero row.
__global__ void k_eca_apply(const float4 *__restrict__ t, const float4 *__restrict__ res, const float *__restrict__ gate,
                            const uint64_t *__restrict__ keys, int batch_shift, int n, int c4, int relu, int res_split, int out_split,
                            float4 *__restrict__ out) {
  const int64_t total = (int64_t)n * c4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;                                      // four 16-byte elements in flight per thread (pure streaming pass)
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
    float4 v[U], rr[U];
    uint64_t key[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) {
        v[u] = t[i];
        if (gate) key[u] = keys[(int)(i / c4)];
        if (res) rr[u] = res_split ? presplit_unpack(((const uint4 *)res)[i]) : res[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= total) continue;
      float4 x = v[u];
      if (gate) {
        const int b = (int)(key[u] >> batch_shift), q = (int)(i % c4);
        const float4 g = *(const float4 *)(gate + ((size_t)b * c4 + q) * 4);
        x.x *= g.x; x.y *= g.y; x.z *= g.z; x.w *= g.w;
      }
      if (res) { x.x += rr[u].x; x.y += rr[u].y; x.z += rr[u].z; x.w += rr[u].w; }
      if (relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
      if (out_split) ((uint4 *)out)[i] = presplit_pack(x); else out[i] = x;
    }
  }
  if (out_split && blockIdx.x == 0 && (int)threadIdx.x < c4) out[total + threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
}
// pre-split map -> fp32 (taps, and consumers that only read fp32)
__global__ void k_presplit_to_f32(const uint4 *__restrict__ in, int64_t n4, float4 *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) out[i] = presplit_unpack(in[i]);
}
__global__ void k_bcast_mul(const float *__restrict__ x, const float *__restrict__ g, const uint64_t *__restrict__ keys,
                            int batch_shift, int n, int c, float *__restrict__ out) {
  const int64_t total = (int64_t)n * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c), ch = (int)(i % c);
    out[i] = x[i] * g[(size_t)(keys[r] >> batch_shift) * c + ch];
  }
}
// F.normalize(x, p=2, dim=1, eps=1e-12): one warp per row
__global__ void k_l2norm_rows(const float *__restrict__ x, int n, int c, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    float ss = 0.f;
    for (int ch = lane; ch < c; ch += 32) { const float v = x[(size_t)r * c + ch]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    for (int ch = lane; ch < c; ch += 32) out[(size_t)r * c + ch] = x[(size_t)r * c + ch] * inv;
  }
}
// keypoint offset tanh + Quantizer.keypoint_position (datasets/quantization.py:60-72, 93-103); sigma softplus
__global__ void k_kp_sigma(const float *__restrict__ kp_raw /* (n,>=3), row stride kp_stride */, int kp_stride,
                           const float *__restrict__ sg_raw /* (n,>=1), row stride sg_stride */, int sg_stride,
                           const uint64_t *__restrict__ keys, int level, int n, int polar, float q0, float q1, float q2,
                           int ignore_offset, float *__restrict__ kp_out, float *__restrict__ sg_out) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    if (kp_out) {
      uint32_t b, vx, vy, vz;
      split_key(level, keys[r], b, vx, vy, vz);
      const float cx = (float)((int)(vx << level) - kAxisBias), cy = (float)((int)(vy << level) - kAxisBias),
                  cz = (float)((int)(vz << level) - kAxisBias);
      const float st = (float)(1 << level);
      float ox = 0.f, oy = 0.f, oz = 0.f;
      if (!ignore_offset) {
        const float *kr = kp_raw + (size_t)r * kp_stride;
        ox = tanhf(kr[0]); oy = tanhf(kr[1]); oz = tanhf(kr[2]);
      }
      const float qy = polar ? q1 : q0, qz = polar ? q2 : q0;
      // (c + 0.5) * q + off * (stride * q) / 2
      const float kx = __fadd_rn(__fmul_rn(cx + 0.5f, q0), __fmul_rn(__fmul_rn(ox, __fmul_rn(st, q0)), 0.5f));
      const float ky = __fadd_rn(__fmul_rn(cy + 0.5f, qy), __fmul_rn(__fmul_rn(oy, __fmul_rn(st, qy)), 0.5f));
      const float kz = __fadd_rn(__fmul_rn(cz + 0.5f, qz), __fmul_rn(__fmul_rn(oz, __fmul_rn(st, qz)), 0.5f));
      if (polar) {
        const float th = __fdiv_rn(__fmul_rn(3.14159265358979323846f, kx - 180.f), 180.f);
        kp_out[3 * (size_t)r] = cosf(th) * ky;
        kp_out[3 * (size_t)r + 1] = sinf(th) * ky;
        kp_out[3 * (size_t)r + 2] = kz;
      } else {
        kp_out[3 * (size_t)r] = kx; kp_out[3 * (size_t)r + 1] = ky; kp_out[3 * (size_t)r + 2] = kz;
      }
    }
    if (sg_out) {
      const float x = sg_raw[(size_t)r * sg_stride];
      sg_out[r] = x > 20.f ? x : log1pf(expf(x));
    }
  }
}

// per-cloud k smallest sigma (eval/evaluate.py:352-361: torch.topk(sigma, k, largest=False), sorted ascending).
// One CTA per cloud: 4-pass MSB radix select (8-bit digits, shared-memory histograms) finds the k-th smallest key,
// everything below it is collected, ties on the threshold are taken in row order, and the <= 256 survivors are
// bitonic-sorted by (value, row).  O(n) per cloud instead of k passes over the data.
constexpr int kTopkThreads = 1024;
constexpr int kTopkMaxK = 1024;

__device__ __forceinline__ uint32_t float_order_key(float v) {   // monotone float -> uint32 (NaN sorts last)
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(kTopkThreads) k_topk_smallest(const float *__restrict__ sigma, const int *__restrict__ off, int k,
                                                                int *__restrict__ idx_out) {
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned long long s_cand[kTopkMaxK];
  __shared__ unsigned int s_prefix, s_remaining, s_ncand, s_tie_base;
  __shared__ unsigned int s_warp[32];
  const int b = blockIdx.x, r0 = off[b], n = off[b + 1] - r0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kk = min(k, n);
  int *out = idx_out + (size_t)b * k;
  if (kk == 0) {
    for (int i = tid; i < k; i += kTopkThreads) out[i] = -1;
    return;
  }
  if (tid == 0) { s_prefix = 0u; s_remaining = (unsigned)kk; s_ncand = 0u; s_tie_base = 0u; }
  __syncthreads();
  // ---- radix select: after pass p the top 8*(p+1) bits of the k-th smallest key are known ----
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += kTopkThreads) s_hist[i] = 0u;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const unsigned int himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = tid; i < n; i += kTopkThreads) {
      const uint32_t key = float_order_key(sigma[r0 + i]);
      if ((key & himask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (warp == 0) {                                      // find the digit holding the `remaining`-th element
      unsigned int c[8], sum = 0u;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = s_hist[lane * 8 + j]; sum += c[j]; }
      unsigned int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
      unsigned int before = incl - sum;
      const unsigned int rem = s_remaining;
      if (before < rem && rem <= incl) {                  // exactly one lane
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (rem <= before + c[j]) { s_prefix = prefix | ((unsigned)(lane * 8 + j) << shift); s_remaining = rem - before; break; }
          before += c[j];
        }
      }
    }
    __syncthreads();
  }
  const unsigned int T = s_prefix;           // key of the k-th smallest element
  const unsigned int need_ties = s_remaining; // how many elements equal to T belong to the result (lowest rows first)
  // ---- collect: keys < T unordered, keys == T in row order ----
  for (int base = 0; base < n; base += kTopkThreads) {
    const int i = base + tid;
    uint32_t key = 0xFFFFFFFFu;
    bool lt = false, eq = false;
    if (i < n) { key = float_order_key(sigma[r0 + i]); lt = key < T; eq = key == T; }
    if (lt) s_cand[atomicAdd(&s_ncand, 1u)] = ((unsigned long long)key << 32) | (unsigned)i;
    const unsigned int bal = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    unsigned int wbase = s_tie_base;
    for (int w = 0; w < warp; ++w) wbase += s_warp[w];
    const unsigned int rank = wbase + __popc(bal & ((1u << lane) - 1u));
    if (eq && rank < need_ties) s_cand[atomicAdd(&s_ncand, 1u)] = ((unsigned long long)key << 32) | (unsigned)i;
    __syncthreads();
    if (tid == 0) { unsigned int t = 0; for (int w = 0; w < 32; ++w) t += s_warp[w]; s_tie_base += t; }
    __syncthreads();
  }
  // ---- bitonic sort of the kk survivors by (key, row) ----
  int m = 1;
  while (m < kk) m <<= 1;
  for (int i = kk + tid; i < m; i += kTopkThreads) s_cand[i] = ~0ull;
  __syncthreads();
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < m / 2; t += kTopkThreads) {
        const int lo = (t / stride) * stride * 2 + (t % stride), hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long x = s_cand[lo], y = s_cand[hi];
        if ((x > y) == up) { s_cand[lo] = y; s_cand[hi] = x; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += kTopkThreads) out[i] = i < kk ? (int)(unsigned)(s_cand[i] & 0xFFFFFFFFull) : -1;
}

// ------------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------------
// conv0_tc.cu
int run_conv0_tc(egn_ctx *ctx, const void *wpack, const float *scale, const float *shift, int relu, const int *not_ones, int out_split,
                 float *out, double bytes, double flops, cudaStream_t s);

int run_conv0(egn_ctx *ctx, int ksize, const float *f0, const float *w, const float *scale, const float *shift, int cout,
              int relu, const int *not_ones, int out_split, float *out, cudaStream_t s, const void *wtc) {
  const Pyramid &py = ctx->pyr;
  EGN_CHECK(ksize == 5 || ksize == 3, EGN_ERR_INVALID, "conv0: kernel size %d not supported (3 or 5)", ksize);
  EGN_CHECK(cout == 32, EGN_ERR_INVALID, "conv0: %d output channels not supported (the egonn / MinkLoc3D stems use 32)", cout);
  const int n0 = py.n[0];
  EGN_CHECK(n0 <= (1 << 25), EGN_ERR_INVALID, "conv0: more than 2^25 voxels");
  const int kv = ksize * ksize * ksize;
  const double pairs = (double)py.pairs_conv0;  // profile mode only (0 otherwise)
  const double bytes = pairs * (1 + cout) * 4 + pairs * 8 + (double)kv * cout * 4, flops = 2.0 * pairs * cout;
  const size_t fixed = (size_t)(kv * kC0WStride + 4) * 4 + 36 * 8;
  const size_t smem1 = fixed + (size_t)C0<true>::kWarps * kv * 32, smem0 = fixed + (size_t)C0<false>::kWarps * kv * 32 * 4;
  const int blocks1 = (int)std::min<int64_t>(div_up(n0, C0<true>::kWarps * 32), (int64_t)kNumSMs * C0<true>::kCtas * 4);
  const int blocks0 = (int)std::min<int64_t>(div_up(n0, C0<false>::kWarps * 32), (int64_t)kNumSMs * C0<false>::kCtas * 4);
#define EGN_C0(KS, NAME)                                                                                                    \
  {                                                                                                                         \
    EGN_SMEM_OPTIN(ctx, (k_conv0<KS, true>), smem1);                                                                        \
    EGN_SMEM_OPTIN(ctx, (k_conv0<KS, false>), smem0);                                                                       \
    if (not_ones && wtc && KS == 5)                                                                                         \
      EGN_TRY(run_conv0_tc(ctx, wtc, scale, shift, relu, not_ones, out_split, out, bytes, flops, s));                       \
    else if (not_ones)                                                                                                      \
      EGN_LAUNCH(ctx, NAME, bytes, flops, s,                                                                                \
                 k_conv0<KS, true><<<blocks1, C0<true>::kWarps * 32, smem1, s>>>(f0, py.keys[0], py.up[0], py.up[1], py.nbr[2], \
                                                                                 py.mask64, py.first0, n0, w, scale, shift, relu, not_ones, out_split, out)); \
    EGN_LAUNCH(ctx, not_ones ? NAME "(general variant)" : NAME, not_ones ? 0.0 : bytes, not_ones ? 0.0 : flops, s,          \
               k_conv0<KS, false><<<blocks0, C0<false>::kWarps * 32, smem0, s>>>(f0, py.keys[0], py.up[0], py.up[1], py.nbr[2], \
                                                                                py.mask64, py.first0, n0, w, scale, shift, relu, not_ones, out_split, out)); \
  }
  if (ksize == 5) EGN_C0(5, "conv0_5x5x5") else EGN_C0(3, "conv0_3x3x3")
#undef EGN_C0
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

// generic conv on the pyramid (cin >= 1); see egn_conv in the public header
int run_conv(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout, const float *in, const float *w,
             const float *scale, const float *shift, int relu, int accumulate, float *out, cudaStream_t s) {
  const Pyramid &py = ctx->pyr;
  EGN_CHECK(py.valid, EGN_ERR_STATE, "conv before coords_build");
  EGN_CHECK(cin >= 1 && cout >= 1 && in && w && out, EGN_ERR_INVALID, "conv: bad argument");
  ConvArgs a = {};
  a.in = in; a.out = out; a.w = w; a.scale = scale; a.shift = shift;
  a.cin = cin; a.cout = cout; a.relu = relu; a.accumulate = accumulate;
  long long pairs = 0;
  if (ksize == 1) {
    EGN_CHECK(level_in >= 0 && level_in < P, EGN_ERR_INVALID, "conv: bad level");
    a.mode = G_IDENTITY; a.K = 1; a.n_out = py.n[level_in];
    pairs = a.n_out;
  } else if (ksize == 3) {
    EGN_CHECK(level_in >= 1 && level_in < P, EGN_ERR_INVALID, "conv k=3: level must be 1..%d (level 0 has no neighbour table)", P - 1);
    a.mode = G_NBR27; a.K = 27; a.n_out = py.n[level_in]; a.nbr = py.nbr[level_in];
    pairs = py.pairs27[level_in];
  } else if (ksize == 2 && !transposed) {
    EGN_CHECK(level_in >= 0 && level_in + 1 < P, EGN_ERR_INVALID, "conv k=2 s=2: bad level");
    a.mode = G_CHILD8; a.K = 8; a.n_out = py.n[level_in + 1];
    a.cstart = py.cstart[level_in + 1]; a.cmask = py.cmask[level_in + 1];
    pairs = py.n[level_in];
  } else if (ksize == 2 && transposed) {
    EGN_CHECK(level_in >= 1 && level_in < P, EGN_ERR_INVALID, "transposed conv: bad level");
    a.mode = G_PARENT; a.K = 8; a.n_out = py.n[level_in - 1];
    a.up = py.up[level_in - 1]; a.keys = py.keys[level_in - 1];
    pairs = a.n_out;
  } else if (ksize == 5) {
    EGN_CHECK(level_in == 0 && cin == 1 && !accumulate, EGN_ERR_INVALID, "conv k=5 is supported at level 0 with cin=1 only");
    return run_conv0(ctx, 5, in, w, scale, shift, cout, relu, nullptr, 0, out, s, nullptr);
  } else {
    EGN_CHECK(false, EGN_ERR_INVALID, "conv: unsupported kernel size %d", ksize);
  }
  return launch_sconv(ctx, a, pairs, s);
}

int op_conv(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout, const float *in, const float *w,
            const float *scale, const float *shift, int relu, int accumulate, float *out, cudaStream_t s) {
  EGN_CHECK(ctx != nullptr, EGN_ERR_INVALID, "null ctx");
  return run_conv(ctx, level_in, ksize, transposed, cin, cout, in, w, scale, shift, relu, accumulate, out, s);
}

// slices per cloud for the two-stage per-cloud reductions: ~48 rows per CTA, at most ~8 CTAs per SM in total - a single
// 1M-point cloud (BASELINE config 5) gets as many CTAs as a batch of 16 scans
static int pool_slices(int n_rows, int n_batches) {
  const int cap = std::max(1, std::min(256, (kNumSMs * 8) / std::max(1, n_batches)));   // <= 256: the second stage walks them per cloud
  int s = (int)div_up(n_rows, (int64_t)std::max(1, n_batches) * 48);
  return s < 1 ? 1 : (s > cap ? cap : s);
}

// per-cloud pooling; part must hold n_batches*slices*c floats (taken from the feature arena by the caller)
int run_pool(egn_ctx *ctx, int level, int c, const float *in, int mode, float p, float eps, float *part, int slices,
             float *out, cudaStream_t s) {
  const Pyramid &py = ctx->pyr;
  const int B = py.n_batches;
  const int threads = c >= 256 ? 256 : (c < 32 ? 32 : c);
  EGN_CHECK((c & 3) == 0 && c <= 1024 && ((uintptr_t)in & 15) == 0, EGN_ERR_INVALID, "pooling: channels must be a multiple of 4 (<= 1024), 16-byte aligned rows");
  if (c <= 256 && B <= kPoolCounters) {                     // one launch: the last CTA of every cloud finishes it
    PoolTail tail = {1, ctx->pool_counters, out, nullptr, 0};
    EGN_LAUNCH(ctx, "global_pool", (double)py.n[level] * c * 4 + (double)B * (slices + 1) * c * 4, 0, s,
               k_pool_partial<<<dim3(B, slices), 256, 0, s>>>(in, py.boff[level], c, slices, mode, p, eps, part, tail));
    EGN_CUDA(cudaGetLastError());
    return EGN_OK;
  }
  EGN_LAUNCH(ctx, "global_pool", (double)py.n[level] * c * 4, 0, s,
             k_pool_partial<<<dim3(B, slices), 256, 0, s>>>(in, py.boff[level], c, slices, mode, p, eps, part, PoolTail{0, nullptr, nullptr, nullptr, 0}));
  EGN_LAUNCH(ctx, "global_pool", (double)B * (slices + 1) * c * 4, 0, s,
             k_pool_final<<<B, 256, 0, s>>>(part, py.boff[level], c, slices, mode, p, out));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

int op_global_pool(egn_ctx *ctx, int level, int c, const float *in, int mode, float p, float eps, float *out, cudaStream_t s) {
  EGN_CHECK(ctx && ctx->pyr.valid, EGN_ERR_STATE, "global_pool before coords_build");
  EGN_CHECK(level >= 0 && level < P && c >= 1 && in && out, EGN_ERR_INVALID, "global_pool: bad argument");
  const Pyramid &py = ctx->pyr;
  const int slices = pool_slices(py.n[level], py.n_batches);
  // the sort scratch is dead after coords_build; reuse on the same stream is stream-ordered
  EGN_TRY(ctx->scratch.reserve((size_t)py.n_batches * slices * c * 4 + 4096, s));
  float *part = (float *)ctx->scratch.take((size_t)py.n_batches * slices * c * 4);
  EGN_CHECK(part != nullptr, EGN_ERR_STATE, "scratch arena exhausted");
  return run_pool(ctx, level, c, in, mode, p, eps, part, slices, out, s);
}

int op_broadcast_mul(egn_ctx *ctx, int level, int c, const float *in, const float *g, float *out, cudaStream_t s) {
  EGN_CHECK(ctx && ctx->pyr.valid, EGN_ERR_STATE, "broadcast_mul before coords_build");
  EGN_CHECK(level >= 0 && level < P && c >= 1 && in && g && out, EGN_ERR_INVALID, "broadcast_mul: bad argument");
  const Pyramid &py = ctx->pyr;
  const int n = py.n[level];
  if (n == 0) return EGN_OK;
  EGN_LAUNCH(ctx, "broadcast_mul", (double)n * c * 8, 0, s,
             k_bcast_mul<<<grid_for((int64_t)n * c, 256), 256, 0, s>>>(in, g, py.keys[level], kMortonBits - 3 * level, n, c, out));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

// selected keypoints of every cloud, packed for ONE device-to-host copy (eval/evaluate.py:339-350: descriptors[ndx], keypoints[ndx],
// global descriptor, per cloud): out[b] = [global (G) | keypoints (k,3) | descriptors (k,D)], zeros where idx == -1
__global__ void k_pack_topk(const int *__restrict__ idx, const int *__restrict__ offsets, int k, const float *__restrict__ kp,
                            const float *__restrict__ desc, int D, const float *__restrict__ glob, int G, float *__restrict__ out) {
  const int b = blockIdx.y;
  const int per = G + k * (3 + D);
  float *ob = out + (size_t)b * per;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < G; i += blockDim.x) ob[i] = glob[(size_t)b * G + i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = blockIdx.x * nw + warp; j < k; j += gridDim.x * nw) {
    const int sel = idx[(size_t)b * k + j];
    const int row = sel >= 0 ? offsets[b] + sel : -1;
    if (lane < 3) ob[G + j * 3 + lane] = row >= 0 ? kp[(size_t)row * 3 + lane] : 0.f;
    float *od = ob + G + k * 3 + (size_t)j * D;
    for (int c = lane; c < D; c += 32) od[c] = row >= 0 ? desc[(size_t)row * D + c] : 0.f;
  }
}
int op_pack_topk(const int32_t *idx, const int32_t *offsets, int n_batches, int k, const float *kp, const float *desc, int D,
                 const float *glob, int G, float *out, cudaStream_t s) {
  EGN_CHECK(idx && offsets && kp && desc && out && n_batches >= 1 && k >= 1 && D >= 1 && (G == 0 || glob), EGN_ERR_INVALID, "pack_topk: bad argument");
  k_pack_topk<<<dim3((unsigned)std::min(16, (k + 7) / 8), (unsigned)n_batches), 256, 0, s>>>(idx, offsets, k, kp, desc, D, glob, G, out);
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

int op_topk(const float *sigma, const int32_t *offsets, int n_batches, int k, int32_t *idx_out, cudaStream_t s) {
  EGN_CHECK(sigma && offsets && idx_out && n_batches >= 1 && k >= 1, EGN_ERR_INVALID, "topk: bad argument");
  EGN_CHECK(k <= kTopkMaxK, EGN_ERR_INVALID, "topk: k=%d exceeds %d", k, kTopkMaxK);
  k_topk_smallest<<<n_batches, kTopkThreads, 0, s>>>(sigma, offsets, k, idx_out);
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

// ---- global-descriptor retrieval (eval/evaluate.py:173-176) ------------------------------------------------
// embed_dist = ||map - query||_2 computed on the difference (like np.linalg.norm(map - q, axis=1)), one warp per
// (query, map row); nn = the k smallest per query by the same radix select + bitonic sort as the keypoint top-k.
__global__ void k_l2_dist(const float *__restrict__ query, const float *__restrict__ map, int Q, int M, int D, float *__restrict__ dist) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5, total = (int64_t)Q * M;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
    const int q = (int)(w / M), m = (int)(w % M);
    const float *a = query + (size_t)q * D, *b = map + (size_t)m * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float t = b[d] - a[d]; ss = fmaf(t, t, ss); }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) dist[w] = sqrtf(ss);
  }
}
__global__ void k_row_offsets(int *__restrict__ off, int Q, int M) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= Q; i += gridDim.x * blockDim.x) off[i] = i * M;
}

int op_knn_l2(egn_ctx *ctx, const float *query, const float *map, int Q, int M, int D, int k, int32_t *idx_out, float *dist_out,
              cudaStream_t s) {
  EGN_CHECK(ctx && query && map && idx_out && dist_out, EGN_ERR_INVALID, "knn: null argument");
  EGN_CHECK(Q >= 1 && M >= 1 && D >= 1 && k >= 1 && k <= kTopkMaxK && (int64_t)Q * M < ((int64_t)1 << 31), EGN_ERR_INVALID,
            "knn: bad sizes (Q=%d M=%d D=%d k=%d)", Q, M, D, k);
  EGN_TRY(ctx->scratch.reserve(pad256((size_t)(Q + 1) * 4) + 4096, s));
  int *off = (int *)ctx->scratch.take((size_t)(Q + 1) * 4);
  EGN_CHECK(off != nullptr, EGN_ERR_STATE, "scratch arena exhausted");
  EGN_LAUNCH(ctx, "retrieval_l2_dist", (double)Q * M * (D * 4.0 + 4), 3.0 * Q * M * D, s,
             k_l2_dist<<<grid_for((int64_t)Q * M * 32, 256, 16), 256, 0, s>>>(query, map, Q, M, D, dist_out));
  k_row_offsets<<<(Q + 256) / 256, 256, 0, s>>>(off, Q, M);
  EGN_LAUNCH(ctx, "retrieval_topk", (double)Q * M * 4, 0, s, k_topk_smallest<<<Q, kTopkThreads, 0, s>>>(dist_out, off, k, idx_out));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

// ---- local-descriptor matching ------------------------------------------------------------------------------
// The correspondence step of eval/evaluate.py:381-399 (Open3D registration_ransac_based_on_feature_matching with
// mutual_filter=True): nearest neighbour in descriptor space from every row of A in B, kept when it is mutual.
// One CTA per row of `a`: the row sits in shared memory, thread t scans rows t, t+256, .. of `b` (squared L2 on the
// difference), then a fixed-order argmin (ties: lower row) - deterministic.
__global__ void __launch_bounds__(256) k_nn_rows(const float *__restrict__ a, const float *__restrict__ b, int nb, int d,
                                                 int *__restrict__ idx, float *__restrict__ dist) {
  extern __shared__ float s_row[];
  __shared__ float s_best[256];
  __shared__ int s_arg[256];
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i < d; i += 256) s_row[i] = a[(size_t)r * d + i];
  __syncthreads();
  float best = INFINITY;
  int arg = -1;
  for (int j = threadIdx.x; j < nb; j += 256) {
    const float *q = b + (size_t)j * d;
    float ss = 0.f;
    for (int i = 0; i < d; ++i) { const float t = q[i] - s_row[i]; ss = fmaf(t, t, ss); }
    if (ss < best) { best = ss; arg = j; }
  }
  s_best[threadIdx.x] = best;
  s_arg[threadIdx.x] = arg;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float v = s_best[threadIdx.x + o];
      const int g = s_arg[threadIdx.x + o];
      if (g >= 0 && (v < s_best[threadIdx.x] || (v == s_best[threadIdx.x] && g < s_arg[threadIdx.x]) || s_arg[threadIdx.x] < 0)) {
        s_best[threadIdx.x] = v;
        s_arg[threadIdx.x] = g;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { idx[r] = s_arg[0]; if (dist) dist[r] = s_arg[0] >= 0 ? sqrtf(s_best[0]) : INFINITY; }
}
__global__ void k_mutual_filter(int *__restrict__ idx_ab, const int *__restrict__ idx_ba, int na) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < na; i += gridDim.x * blockDim.x) {
    const int j = idx_ab[i];
    if (j >= 0 && idx_ba[j] != i) idx_ab[i] = -1;
  }
}

int op_match_mutual(egn_ctx *ctx, const float *a, const float *b, int na, int nb, int d, int mutual, int32_t *idx_out, float *dist_out,
                    cudaStream_t s) {
  EGN_CHECK(ctx && a && b && idx_out, EGN_ERR_INVALID, "match: null argument");
  EGN_CHECK(na >= 1 && nb >= 1 && d >= 1 && d <= 4096, EGN_ERR_INVALID, "match: bad sizes (na=%d nb=%d dim=%d)", na, nb, d);
  EGN_LAUNCH(ctx, "match_nearest", (double)na * nb * d * 4.0, 3.0 * na * nb * d, s,
             k_nn_rows<<<na, 256, (size_t)d * 4, s>>>(a, b, nb, d, idx_out, dist_out));
  if (mutual) {
    EGN_TRY(ctx->scratch.reserve(pad256((size_t)nb * 4) + 4096, s));
    int *back = (int *)ctx->scratch.take((size_t)nb * 4);
    EGN_CHECK(back != nullptr, EGN_ERR_STATE, "scratch arena exhausted");
    EGN_LAUNCH(ctx, "match_nearest", (double)na * nb * d * 4.0, 3.0 * na * nb * d, s,
               k_nn_rows<<<nb, 256, (size_t)d * 4, s>>>(b, a, na, d, back, nullptr));
    EGN_LAUNCH(ctx, "match_mutual_filter", (double)na * 12, 0, s, k_mutual_filter<<<grid_for(na, 256), 256, 0, s>>>(idx_out, back, na));
  }
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

// exported to forward.cu
__global__ void k_fill_ones(float *__restrict__ out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = 1.0f;
}
int run_gather_rows1(egn_ctx *ctx, const float *in, const int *perm, int n, float *out, int *not_ones, cudaStream_t s) {
  EGN_CUDA(cudaMemsetAsync(not_ones, 0, sizeof(int), s));
  if (in == nullptr) {   // implicit all-ones occupancy features
    EGN_LAUNCH(ctx, "gather_input_features", (double)n * 4, 0, s, k_fill_ones<<<grid_for(n, 256), 256, 0, s>>>(out, n));
    EGN_CUDA(cudaGetLastError());
    return EGN_OK;
  }
  EGN_LAUNCH(ctx, "gather_input_features", (double)n * 12, 0, s,
             k_gather_rows1<<<grid_for(n, 256), 256, 0, s>>>(in, perm, n, out, not_ones));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}
int run_eca_gate(egn_ctx *ctx, int level, int c, const float *t, const float *wk, int k, float *part, int slices, float *gate,
                 cudaStream_t s) {
  const Pyramid &py = ctx->pyr;
  const int threads = c >= 256 ? 256 : (c < 32 ? 32 : c);
  EGN_CHECK((c & 3) == 0 && c <= 1024, EGN_ERR_INVALID, "eca: channels must be a multiple of 4 (<= 1024)");
  EGN_CHECK(c <= 256, EGN_ERR_INVALID, "eca: at most 256 channels");
  if (py.n_batches <= kPoolCounters) {                      // pooling + gate in ONE launch (the last CTA of a cloud computes its gate)
    PoolTail tail = {2, ctx->pool_counters, gate, wk, k};
    EGN_LAUNCH(ctx, "eca_pool_gate", (double)py.n[level] * c * 4 + (double)py.n_batches * (slices + 1) * c * 4, 0, s,
               k_pool_partial<<<dim3(py.n_batches, slices), 256, 0, s>>>(t, py.boff[level], c, slices, 0, 1.f, 0.f, part, tail));
    EGN_CUDA(cudaGetLastError());
    return EGN_OK;
  }
  EGN_LAUNCH(ctx, "eca_pool", (double)py.n[level] * c * 4, 0, s,
             k_pool_partial<<<dim3(py.n_batches, slices), 256, 0, s>>>(t, py.boff[level], c, slices, 0, 1.f, 0.f, part, PoolTail{0, nullptr, nullptr, nullptr, 0}));
  EGN_LAUNCH(ctx, "eca_gate", (double)py.n_batches * (slices + 1) * c * 4, 0, s,
             k_eca_gate<<<py.n_batches, 256, 0, s>>>(part, py.boff[level], c, slices, wk, k, gate));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}
int run_presplit_to_f32(egn_ctx *ctx, const float *in, int64_t floats, float *out, cudaStream_t s) {
  if (floats <= 0) return EGN_OK;
  EGN_LAUNCH(ctx, "presplit_to_f32", (double)floats * 8, 0, s,
             k_presplit_to_f32<<<grid_for(floats / 4, 256), 256, 0, s>>>((const uint4 *)in, floats / 4, (float4 *)out));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}
int run_eca_apply(egn_ctx *ctx, int level, int c, const float *t, const float *res, const float *gate, int relu, int res_split,
                  int out_split, float *out, cudaStream_t s) {
  const Pyramid &py = ctx->pyr;
  const int n = py.n[level];
  EGN_CHECK((c & 3) == 0, EGN_ERR_INVALID, "block channels must be a multiple of 4");
  if (n == 0) return EGN_OK;
  EGN_LAUNCH(ctx, "eca_apply_residual_relu", (double)n * c * 12, 0, s,
             k_eca_apply<<<grid_for((int64_t)n * (c / 4), 256), 256, 0, s>>>((const float4 *)t, (const float4 *)res, gate, py.keys[level],
                                                                               kMortonBits - 3 * level, n, c / 4, relu, res_split, out_split, (float4 *)out));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}
int run_l2norm(egn_ctx *ctx, const float *x, int n, int c, float *out, cudaStream_t s) {
  if (n == 0) return EGN_OK;
  EGN_LAUNCH(ctx, "l2_normalize_rows", (double)n * c * 8, 0, s, k_l2norm_rows<<<grid_for((int64_t)n * 32, 256), 256, 0, s>>>(x, n, c, out));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}
int run_kp_sigma(egn_ctx *ctx, int level, const float *kp_raw, int kp_stride, const float *sg_raw, int sg_stride, int polar,
                 const float q[3], int ignore_offset, float *kp_out, float *sg_out, cudaStream_t s) {
  const Pyramid &py = ctx->pyr;
  const int n = py.n[level];
  if (n == 0) return EGN_OK;
  EGN_LAUNCH(ctx, "keypoint_position_sigma", (double)n * 40, 0, s,
             k_kp_sigma<<<grid_for(n, 128), 128, 0, s>>>(kp_raw, kp_stride, sg_raw, sg_stride, py.keys[level], level, n, polar, q[0], q[1], q[2], ignore_offset,
                                                         kp_out, sg_out));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}
int pool_slices_for(egn_ctx *ctx, int level) { return pool_slices(ctx->pyr.n[level], ctx->pyr.n_batches); }

}  // namespace egn
