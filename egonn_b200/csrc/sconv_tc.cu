// Tensor-core sparse convolution for sm_100a: output-stationary implicit GEMM with a GATHERED A operand.
//
//   out[o, :] = epilogue( sum_k  in[nbr(o,k), :] @ W[k] )          (MinkowskiConvolution forward)
//
// One CTA owns 128 consecutive output rows (UMMA M = 128, cta_group::1).  The reduction dimension
// (kernel offset k) x (input channel) is cut into chunks of 64 elements = one 128-byte swizzle row:
//   * A chunk  : 128 rows x 64 bf16, K-major, SWIZZLE_128B - written by 4 producer warps that gather the
//                fp32 input rows named by the neighbour table (absent -> zeros) with 16-byte loads, split
//                every value into bf16 hi + bf16 lo and store both images (st.shared.v2 at swizzled offsets).
//   * B chunk  : COUT rows x 64 bf16 (hi image, lo image), pre-swizzled on the host into exactly this shared
//                memory image, fetched by ONE cp.async.bulk (TMA) per chunk with an mbarrier transaction count.
//   * MMA      : one elected thread issues tcgen05.mma kind::f16, 4 K-steps x {hi*hi, lo*hi, hi*lo} per chunk,
//                FP32 accumulators in TMEM (COUT columns); tcgen05.commit releases the stage.
//   * epilogue : the 4 producer warps read their TMEM lane quarter with tcgen05.ld, apply the folded
//                BatchNorm scale/shift (+ReLU) and write the output row once.
// hi/lo split: a = hi + lo with |a - hi - lo| <= 2^-17 |a|; dropping lo*lo leaves a relative error of ~2^-16 per
// product, i.e. fp32-class results (the 1e-3 end-to-end budget of the north star needs better than TF32).
// Chunks in which no row of the tile has a neighbour are skipped entirely (no gather, no TMA, no MMA).
#include "ctx.cuh"
#include "tc_ptx.cuh"

namespace egn {

namespace tc {

using namespace tcx;


// Shared-memory plan.  A stages (gathered activations, 32 KB each) and the B ring (weight chunks, 256*COUT bytes each)
// are separate rings: weight chunks do not depend on anything but a free slot, so the TMA thread runs kBSlots chunks
// ahead and the ~2.5 us bulk-copy latency leaves the per-chunk critical path.
//   COUT == 128          : 1 CTA/SM, 16 gather warps, 3 A stages + 3 B slots (32 KB each)
//   CIN == 128, COUT < 128 (N-split of small levels): 1 CTA/SM, 16 gather warps, 3 A stages + deep B ring
//   otherwise            : 2 CTAs/SM, 8 gather warps, 2 A stages + 2..4 B slots
template <int CIN, int COUT>
struct Cfg {
  static constexpr bool kBig = COUT == 128 || CIN == 128;
  static constexpr int kStages = kBig ? 3 : 2;
  static constexpr int kProducerWarps = kBig ? 16 : 8;
  static constexpr int kCtasPerSm = kBig ? 1 : 2;
  static constexpr int kPrefetch = 1;                               // chunks gathered ahead in registers
  static constexpr int kThreads = (kProducerWarps + 2) * 32;
  static constexpr int kBBytes = 2 * COUT * 128;                    // hi + lo image of one weight chunk
  static constexpr int kBSlots = kStages;                          // weight chunk i lives in slot i % kStages, same barrier as the A stage
  static constexpr int kABytesAll = kStages * 2 * kABytes;
  static constexpr int kSmemBytes = kABytesAll + kBSlots * kBBytes + kRows * 27 * 4 + 2 * COUT * 4 + 768;
};

template <int CIN, int COUT, int KOFF>
__global__ void __launch_bounds__(Cfg<CIN, COUT>::kThreads, Cfg<CIN, COUT>::kCtasPerSm) k_sconv_tc(Args a) {
  using C = Cfg<CIN, COUT>;
  constexpr int kStages = C::kStages;
  constexpr int NPW = C::kProducerWarps, PT = NPW * 32, NT = C::kThreads;
  constexpr int RSTEP = PT / 8;                                   // 8 threads per row, 8 elements (one 16-byte bf16 group) each
  constexpr int F = kRows / RSTEP;                                // row slots per producer thread per chunk
  constexpr int NCH = (KOFF * CIN + kChunk - 1) / kChunk;         // chunks if nothing is skipped
  constexpr int NBR_ITERS = (kRows * KOFF + NT - 1) / NT;
  // dynamic shared memory starts 1024-byte aligned (checked below): SWIZZLE_128B tiles need it
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int kBSlots = C::kBSlots;
  uint8_t *tiles = smem;                                          // A stages: [kStages][hi 16 KB | lo 16 KB]
  uint8_t *btiles = smem + C::kABytesAll;                         // B ring:   [kBSlots][hi | lo] (COUT*128 bytes each image)
  int *s_nbr = (int *)(btiles + kBSlots * C::kBBytes);            // [kRows][KOFF]
  float *s_scale = (float *)(s_nbr + kRows * 27);                 // [COUT]
  float *s_shift = s_scale + COUT;                                // [COUT]
  uint64_t *full = (uint64_t *)(s_shift + COUT);                  // [kStages]  A stage written by the gather warps
  uint64_t *empty = full + kStages;                               // [kStages]  A stage consumed by the tensor core
  uint64_t *bfull = empty + kStages;                              // [kBSlots]  weight chunk landed (TMA transaction bytes)
  uint64_t *bempty = bfull + kBSlots;                             // [kBSlots]  weight chunk consumed
  uint64_t *accum = bempty + kBSlots;                             // [1]
  uint32_t *s_tmem = (uint32_t *)(accum + 1);
  int *s_nlist = (int *)(s_tmem + 1);
  uint32_t *s_present = (uint32_t *)(s_nlist + 1);                // [2] bit j: chunk j has at least one present row
  int *s_list = (int *)(s_present + 2);                           // [NCH] compacted chunk ids

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kRows;
  const int col0 = blockIdx.y * COUT;     // N-split: this CTA computes output channels [col0, col0 + COUT)

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], NPW + 1);   // one arrive per gather warp + the TMA thread's arrive.expect_tx
      mbar_init(&empty[s], 1);        // one tcgen05.commit
    }
    for (int s = 0; s < kBSlots; ++s) {
      mbar_init(&bfull[s], 1);        // the TMA thread's arrive.expect_tx
      mbar_init(&bempty[s], 1);       // one tcgen05.commit
    }
    mbar_init(accum, 1);
    fence_barrier_init();
    s_present[0] = 0u;
    s_present[1] = 0u;
  }
  if (warp == NPW + 1) {              // TMEM: COUT fp32 accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)COUT));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int c = tid; c < COUT; c += NT) {
    s_scale[c] = a.scale ? a.scale[col0 + c] : 1.f;
    s_shift[c] = a.shift ? a.shift[col0 + c] : 0.f;
  }
  __syncthreads();                    // s_present zeroed before the atomics below
  // neighbour rows of the tile: all global loads first, then the shared stores (one latency, not NBR_ITERS)
  {
    int src[NBR_ITERS];
#pragma unroll
    for (int it = 0; it < NBR_ITERS; ++it) {
      const int t = tid + it * NT;
      const int r = t / KOFF, k = t - r * KOFF, row = row0 + r;
      src[it] = -1;
      if (t < kRows * KOFF && row < a.n_out) {
        if (a.mode == 0) src[it] = row;
        else if (a.mode == 1) src[it] = __ldg(a.nbr + (int64_t)row0 * 27 + t);
        else if (a.mode == 2) {
          const uint32_t m = __ldg(a.cmask + row);
          if ((m >> k) & 1u) src[it] = __ldg(a.cstart + row) + __popc(m & ((1u << k) - 1u));
        } else {
          if ((int)(__ldg(a.keys + row) & 7ull) == k) src[it] = __ldg(a.up + row);
        }
      }
    }
    uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
    for (int it = 0; it < NBR_ITERS; ++it) {
      const int t = tid + it * NT;
      if (t < kRows * KOFF) {
        s_nbr[t] = src[it];
        if (src[it] >= 0) {
          const int k = t % KOFF;
          const int j = CIN == 128 ? 2 * k : (CIN == 64 ? k : (k >> 1));
          const uint32_t bits = CIN == 128 ? 3u : 1u;
          if (j < 32) m0 |= bits << j; else m1 |= bits << (j - 32);
        }
      }
    }
    m0 = __reduce_or_sync(0xffffffffu, m0);
    m1 = __reduce_or_sync(0xffffffffu, m1);
    if (lane == 0) {
      if (m0) atomicOr(&s_present[0], m0);
      if (m1) atomicOr(&s_present[1], m1);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    // K-split: this CTA owns chunks [j_lo, j_hi) of the reduction
    const int j_lo = (int)(((int64_t)NCH * blockIdx.z) / a.ksplit), j_hi = (int)(((int64_t)NCH * (blockIdx.z + 1)) / a.ksplit);
    int n = 0;
    for (int j = j_lo; j < j_hi; ++j)
      if ((s_present[j >> 5] >> (j & 31)) & 1u) s_list[n++] = j;
    *s_nlist = n;
  }
  __syncthreads();
  const int nlist = *s_nlist;
  const uint32_t tmem_base = *s_tmem;

  if (warp < NPW) {
    // ===================== A producers: gather + bf16 split, next chunk prefetched in registers =====================
    // thread -> (row slot rs, 16-byte group g): rows r = p*RSTEP + rs share (r & 7) = (rs & 7), so the swizzled
    // byte offset inside a row is a per-thread constant and every row offset is an immediate.
    const int g = tid & 7, rs = tid >> 3;
    const int sw_off = rs * 128 + ((g ^ (rs & 7)) << 4);
    const int *my_nbr = s_nbr + rs * KOFF;
    auto issue = [&](int j, float4 (&da)[F], float4 (&db)[F]) {
      int koff, coff;                      // kernel offset and float offset inside the source row for this thread
      if (CIN == 32) { koff = 2 * j + (g >> 2); coff = (g & 3) * 8; }
      else if (CIN == 64) { koff = j; coff = g * 8; }
      else { koff = j >> 1; coff = (j & 1) * 64 + g * 8; }
      const bool kvalid = koff < KOFF;
      if (!kvalid) koff = 0;
#pragma unroll
      for (int p = 0; p < F; ++p) {
        const int src = my_nbr[p * RSTEP * KOFF + koff];
        const bool pr = kvalid && src >= 0;
        ldg8_pred(a.in + (size_t)(pr ? src : 0) * CIN + coff, pr, da[p], db[p]);
      }
    };
    // register ring of PF+1 chunk buffers: the gathers of chunks i+1..i+PF are in flight while chunk i is converted
    constexpr int PF = C::kPrefetch;
    float4 ra[PF + 1][F], rb[PF + 1][F];
#pragma unroll
    for (int d = 0; d < PF; ++d)
      if (d < nlist) issue(s_list[d], ra[d], rb[d]);
    for (int i0 = 0; i0 < nlist; i0 += PF + 1) {
#pragma unroll
      for (int u = 0; u <= PF; ++u) {                       // u is a compile-time ring index: buffers stay in registers
        const int i = i0 + u;
        if (i < nlist) {
          const int s = i % kStages;
          const uint32_t ph = (uint32_t)(i / kStages) & 1u;
          if (i + PF < nlist) issue(s_list[i + PF], ra[(u + PF) % (PF + 1)], rb[(u + PF) % (PF + 1)]);
          const bool tr = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0 && i < 64;
          if (tr) a.trace[i * 8 + 5] = clock64();
          mbar_wait(&empty[s], ph ^ 1u, a.hint_producer);
          if (tr) a.trace[i * 8 + 6] = clock64();
          uint8_t *a_hi = tiles + s * (2 * kABytes) + sw_off, *a_lo = a_hi + kABytes;
#pragma unroll
          for (int p = 0; p < F; ++p) {
            uint4 hi, lo;
            split2(ra[u][p].x, ra[u][p].y, hi.x, lo.x);
            split2(ra[u][p].z, ra[u][p].w, hi.y, lo.y);
            split2(rb[u][p].x, rb[u][p].y, hi.z, lo.z);
            split2(rb[u][p].z, rb[u][p].w, hi.w, lo.w);
            *(uint4 *)(a_hi + p * RSTEP * 128) = hi;
            *(uint4 *)(a_lo + p * RSTEP * 128) = lo;
          }
          fence_proxy_async();            // this thread's generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();                   // ... of all 32 lanes, then ONE mbarrier arrive per warp instead of 32
          if (lane == 0) mbar_arrive(&full[s]);
          if (tr) a.trace[i * 8 + 7] = clock64();
        }
      }
    }
    // ===================== epilogue: TMEM -> scale/shift/relu -> global =====================
    // warp w reads TMEM lane quarter (w & 3) and the column group (w >> 2)
    constexpr int CPW = (COUT / (NPW / 4)) >= 16 ? COUT / (NPW / 4) : 16;   // tcgen05.ld granularity: 16 columns
    if (nlist > 0) {
      mbar_wait(accum, 0u, a.hint_producer);
      tc_fence_after();
    }
    const int q = warp & 3, cg = warp >> 2;
    const int row = row0 + q * 32 + lane;
    float *obase = a.out + (size_t)blockIdx.z * a.n_out * a.cout_total;
#pragma unroll
    for (int cc = 0; cc < CPW; cc += 16) {
      const int c0 = cg * CPW + cc;
      if (c0 >= COUT) break;                                                 // more gather warps than column groups
      uint32_t r[16];
      if (nlist > 0) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = 0u;             // no chunk of this split touches the tile: partial = 0
      }
      if (row < a.n_out) {
        float *o = obase + (size_t)row * a.cout_total + col0 + c0;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          float4 y;
          float *yy = (float *)&y;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c0 + gg * 4 + e;
            float val = fmaf(__uint_as_float(r[gg * 4 + e]), s_scale[c], s_shift[c]);
            if (a.relu) val = fmaxf(val, 0.f);
            yy[e] = val;
          }
          if (a.accumulate) {
            const float4 prev = *(const float4 *)(o + gg * 4);
            y.x += prev.x; y.y += prev.y; y.z += prev.z; y.w += prev.w;
          }
          *(float4 *)(o + gg * 4) = y;
        }
      }
    }
    tc_fence_before();
  } else if (warp == NPW) {
    // ===================== B loader: runs kBSlots chunks ahead of the tensor core =====================
    if (lane == 0) {
      for (int i = 0; i < nlist; ++i) {
        const int j = s_list[i];
        const int sb = i % kBSlots;
        const uint32_t ph = (uint32_t)(i / kBSlots) & 1u;
        mbar_wait(&empty[sb], ph ^ 1u, a.hint_single);
        mbar_arrive_expect_tx(&full[sb], (uint32_t)C::kBBytes);
        // this CTA's COUT rows of the hi image and of the lo image of chunk j (contiguous when COUT == cout_total)
        const uint8_t *src = a.wpack + (size_t)j * 2 * a.cout_total * 128 + (size_t)col0 * 128;
        uint8_t *dst = btiles + sb * C::kBBytes;
        bulk_g2s(dst, src, (uint32_t)(COUT * 128), &full[sb]);
        bulk_g2s(dst + COUT * 128, src + (size_t)a.cout_total * 128, (uint32_t)(COUT * 128), &full[sb]);
      }
    }
  } else {
    // ===================== MMA issuer =====================
    // The WHOLE warp walks the loop converged, so barrier addresses and matrix descriptors are warp-uniform values
    // (uniform registers); only the tcgen05 instructions themselves are issued by one elected lane.  With the loop
    // under `if (lane == 0)` the compiler wrapped every UTCHMMA in an ELECT + 5x R2UR.BROADCAST + branch sequence:
    // ~80 cycles per MMA, which made this thread - not the gathers, not the weights - the bottleneck.
    constexpr uint32_t idesc = umma_idesc(COUT);
    for (int i = 0; i < nlist; ++i) {
      const int s = i % kStages, sb = i % kBSlots;
      const bool tr = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && i < 64 && lane == 0;
      if (tr) a.trace[i * 8 + 0] = clock64();
      if (tr) a.trace[i * 8 + 1] = clock64();
      mbar_wait(&full[s], (uint32_t)(i / kStages) & 1u, a.hint_single);   // gathered rows AND the weight chunk have landed
      if (tr) a.trace[i * 8 + 2] = clock64();
      tc_fence_after();
      const uint32_t sa = smem_u32(tiles + s * (2 * kABytes));
      const uint32_t sbm = smem_u32(btiles + sb * C::kBBytes);
      const uint32_t first = i == 0 ? 0u : 1u;
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < kChunk / 16; ++ks) {
          const uint64_t ahi = umma_desc(sa + ks * 32), alo = umma_desc(sa + kABytes + ks * 32);
          const uint64_t bhi = umma_desc(sbm + ks * 32), blo = umma_desc(sbm + COUT * 128 + ks * 32);
          umma_f16(tmem_base, ahi, bhi, idesc, ks == 0 ? first : 1u);
          umma_f16(tmem_base, alo, bhi, idesc, 1u);
          umma_f16(tmem_base, ahi, blo, idesc, 1u);
        }
        umma_commit(&empty[s]);          // A stage and weight slot reusable once these MMAs have read them
      }
      __syncwarp();
      if (tr) a.trace[i * 8 + 4] = clock64();
    }
    if (nlist > 0 && elect_one_sync()) umma_commit(accum);   // accumulator complete
    __syncwarp();
  }
  __syncthreads();
  if (warp == NPW + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)COUT));
  }
}

template <int CIN, int COUT, int KOFF>
static int launch(egn_ctx *ctx, const Args &a, const char *name, double bytes, double flops, cudaStream_t s) {
  using C = Cfg<CIN, COUT>;
  static bool attr_done = false;
  if (!attr_done) {
    EGN_CUDA(cudaFuncSetAttribute(k_sconv_tc<CIN, COUT, KOFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_done = true;
  }
  const dim3 grid((unsigned)div_up(a.n_out, kRows), (unsigned)(a.cout_total / COUT), (unsigned)a.ksplit);
  EGN_LAUNCH(ctx, name, bytes, flops, s, k_sconv_tc<CIN, COUT, KOFF><<<grid, C::kThreads, C::kSmemBytes, s>>>(a));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

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
  if (ksize == 1) return !transposed && ((cin == 32 && cout == 64) || (cin == 64 && (cout == 64 || cout == 128)) ||
                                         (cin == 128 && (cout == 64 || cout == 128)));
  if (transposed) return ksize == 2 && cin == cout && (cin == 32 || cin == 64 || cin == 128);
  if (ksize == 3) return (cin == 32 && (cout == 32 || cout == 64)) || (cin == 64 && (cout == 64 || cout == 128)) || (cin == 128 && cout == 128);
  if (ksize == 2) return cin == cout && (cin == 32 || cin == 64 || cin == 128);
  return false;
}

// packed-weight bytes for a (ksize, cin, cout) convolution: n_chunks * 2 images * cout * 128
size_t sconv_tc_wpack_bytes(int ksize, int cin, int cout) {
  const int koff = ksize == 3 ? 27 : (ksize == 2 ? 8 : 1);
  return (size_t)((koff * cin + 63) / 64) * 2 * cout * 128;
}

// sconv_ts.cu: the TMEM-resident-A kernels (ctx->tc_variant == 1)
int launch_conv_ts(egn_ctx *ctx, int koff, int cin, int cout_cta, const tcx::Args &a, const char *name, double bytes, double flops,
                   cudaStream_t s);

int run_conv_tc(egn_ctx *ctx, int level_in, int ksize, int transposed, int cin, int cout, const float *in, const void *wpack,
                const float *scale, const float *shift, int relu, int accumulate, float *out, cudaStream_t s, int in_split,
                int out_split) {
  const Pyramid &py = ctx->pyr;
  EGN_CHECK((!in_split && !out_split) || ctx->tc_variant == 1, EGN_ERR_INVALID, "pre-split feature maps need the TMEM-A kernels");
  EGN_CHECK(!(out_split && accumulate), EGN_ERR_INVALID, "accumulate into a pre-split map is not supported");
  EGN_CHECK(py.valid, EGN_ERR_STATE, "conv before coords_build");
  EGN_CHECK(sconv_tc_supported(ksize, transposed, cin, cout), EGN_ERR_INVALID, "tensor-core conv: unsupported shape k=%d %d->%d", ksize, cin, cout);
  EGN_CHECK(((uintptr_t)wpack & 15) == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, EGN_ERR_INVALID,
            "tensor-core conv: pointers must be 16-byte aligned");
  tc::Args a = {};
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
    pairs = py.pairs27[level_in];
    snprintf(name, sizeof(name), "tc_conv3x3x3_c%d_%d", cin, cout);
  } else if (transposed) {
    EGN_CHECK(level_in >= 1 && level_in < P, EGN_ERR_INVALID, "transposed conv: bad level");
    a.mode = 3; a.n_out = py.n[level_in - 1]; a.up = py.up[level_in - 1]; a.keys = py.keys[level_in - 1]; a.zero_row = py.n[level_in];
    pairs = a.n_out;
    snprintf(name, sizeof(name), "tc_tconv2x2x2s2_c%d_%d", cin, cout);
  } else {
    EGN_CHECK(level_in >= 0 && level_in + 1 < P, EGN_ERR_INVALID, "conv k=2: bad level");
    a.mode = 2; a.n_out = py.n[level_in + 1]; a.cstart = py.cstart[level_in + 1]; a.cmask = py.cmask[level_in + 1]; a.zero_row = py.n[level_in];
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
    const int tiles = (int)div_up(a.n_out, tc::kRows);
    if (tiles <= ctx->nsplit_max) {   // N = 32 -> <= 148 CTAs for <= 37 tiles; N = 64 -> <= 148 CTAs for <= 74 tiles (2 CTAs/SM beyond)
      const int splits = (!ctx->ksplit || out_split) ? 1 : (ksize == 3 ? (tiles <= 37 ? 3 : 1) : 1);
      tc::Args b = a;
      float *part = nullptr;
      if (splits > 1) {
        const size_t bytes_part = (size_t)splits * a.n_out * cout * 4;
        if (ctx->splitk_cap < bytes_part) {
          EGN_CUDA(cudaStreamSynchronize(s));
          if (ctx->splitk_buf) EGN_CUDA(cudaFree(ctx->splitk_buf));
          ctx->splitk_buf = nullptr; ctx->splitk_cap = 0;
          EGN_CUDA(cudaMalloc((void **)&ctx->splitk_buf, bytes_part * 2));
          ctx->splitk_cap = bytes_part * 2;
        }
        part = (float *)ctx->splitk_buf;
        b.out = part; b.scale = nullptr; b.shift = nullptr; b.relu = 0; b.ksplit = splits;
      }
      int st;
      if (ctx->tc_variant == 1) st = launch_conv_ts(ctx, K, 128, tiles <= 37 ? 32 : 64, b, name, bytes, flops, s);
      else if (tiles <= 37) st = ksize == 3 ? tc::launch<128, 32, 27>(ctx, b, name, bytes, flops, s) : tc::launch<128, 32, 8>(ctx, b, name, bytes, flops, s);
      else st = ksize == 3 ? tc::launch<128, 64, 27>(ctx, b, name, bytes, flops, s) : tc::launch<128, 64, 8>(ctx, b, name, bytes, flops, s);
      EGN_TRY(st);
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
  if (ctx->tc_variant == 1) return launch_conv_ts(ctx, K, cin, cout, a, name, bytes, flops, s);
#define EGN_TC_CASE(KS, KO, CI, CO) \
  if (ksize == KS && cin == CI && cout == CO) return tc::launch<CI, CO, KO>(ctx, a, name, bytes, flops, s);
  EGN_TC_CASE(3, 27, 32, 32)
  EGN_TC_CASE(3, 27, 32, 64)
  EGN_TC_CASE(3, 27, 64, 64)
  EGN_TC_CASE(3, 27, 64, 128)
  EGN_TC_CASE(3, 27, 128, 128)
  EGN_TC_CASE(2, 8, 32, 32)
  EGN_TC_CASE(2, 8, 64, 64)
  EGN_TC_CASE(2, 8, 128, 128)
  EGN_TC_CASE(1, 1, 32, 64)
  EGN_TC_CASE(1, 1, 64, 64)
  EGN_TC_CASE(1, 1, 64, 128)
  EGN_TC_CASE(1, 1, 128, 64)
  EGN_TC_CASE(1, 1, 128, 128)
#undef EGN_TC_CASE
  EGN_CHECK(false, EGN_ERR_INVALID, "tensor-core conv: no kernel instance");
}

}  // namespace egn
