// Tensor-core sparse convolution, second generation: the gathered A operand lives in TENSOR MEMORY, not shared memory.
//
//   out[o, :] = epilogue( sum_k  in[nbr(o,k), :] @ W[k] )          (MinkowskiConvolution forward)
//
// Why: in the first-generation kernel (k_sconv_tc, see sconv_tc.cu's header) every 64-element chunk of the gathered operand is written to shared memory as
// two bf16 images (32 KB of STS) and then read three times by tcgen05.mma (hi*hi, lo*hi, hi*lo: 48 KB) - the
// shared-memory port and the instructions that feed it, not HBM and not the tensor pipe, bound the kernel
// (profiles/r01_tc_kernel_stalls.md).  tcgen05.mma can take A from tensor memory ("TS" form), and tcgen05.st writes
// registers straight into it, so here
//   * tiles     : 128 output rows taken from the level's TILE ROW ORDER (coords.cu: rows with similar offset sets are
//                 adjacent, so whole (tile, chunk) pairs vanish); the prologue fetches slot -> row, then all table entries,
//                 in two branch-free load batches, and compacts the list of non-empty chunks.
//   * producers : gather rows with 256-bit loads in the tcgen05.st.16x256b fragment layout (4 threads per row,
//                 8 rows + 8 rows per warp instruction), pre-split bf16 hi/lo maps move bits only (fp32 maps are split in
//                 registers), tcgen05.st both images into a TMEM stage (64 columns: 32 hi + 32 lo).  No shared-memory
//                 traffic for A at all.
//                 A warp may only touch its own TMEM lane quarter, so a GROUP of 4 warps builds one chunk
//                 (warp q -> rows 32q..32q+31); the 2 (or 4) groups of a CTA work on alternate chunks.
//   * weights   : unchanged - pre-swizzled bf16 hi/lo images, one cp.async.bulk (TMA) pair per chunk into a ring.
//   * MMA       : per chunk 4 K-steps x {hi*hi, lo*hi, hi*lo}, A = TMEM columns, B = shared-memory descriptor,
//                 FP32 accumulators in TMEM columns [0, COUT); tcgen05.commit frees the stage.
//   * epilogue  : tcgen05.ld -> the tile transposed through the idle weight ring in shared memory -> folded BatchNorm
//                 scale/shift (+ReLU, +accumulate) -> whole output rows (512 contiguous bytes per warp instruction),
//                 written once as fp32 or pre-split.
// TMEM plan: COUT == 128: 1 CTA/SM, 512 columns = 128 accumulator + 6 stages x 64; otherwise 2 CTAs/SM, 256 columns
// each = 64 accumulator + 3 stages x 64.  bf16x3 split (hi*hi + lo*hi + hi*lo), FP32 accumulation.
// Gathers are 256-bit loads: lane j4 of a row reads channels 8*j4 .. 8*j4+7 of each 32-channel block, one L1 wavefront per
// 128-byte line; the weight image carries the matching K permutation (weights.py: _perm32).
#include "ctx.cuh"
#include "tc_ptx.cuh"

namespace egn {

namespace ts {

using namespace tcx;

// LIGHT: the row-wise layers (1x1x1: 1-2 chunks per tile, <= 64 output channels per CTA) spend most of a tile's life in the
// prologue / first gather / epilogue chain, not in the chunk loop; they run FOUR small CTAs per SM (one producer group, one
// TMEM stage, 128 TMEM columns, 192 threads) so that four of those chains overlap instead of two.
template <int CIN, int COUT, int KOFF, bool LIGHT>
struct Cfg {
  static constexpr bool kBig = COUT == 128;
  static constexpr bool kFold = COUT == 32;                // 2-instruction form: D[:, 0:64] += Ahi*[Bhi;Blo]^T, D[:, 0:32] += Alo*Bhi^T
  static constexpr int kGroups = LIGHT ? 1 : (kBig ? 4 : 2);   // producer groups; one warp per TMEM lane quarter in each
  static constexpr int kProducerWarps = 4 * kGroups;
  static constexpr int kStages = LIGHT ? 1 : (kBig ? 6 : 3);
  static constexpr int kCtasPerSm = LIGHT ? 4 : (kBig ? 1 : 2);
  static constexpr int kTmemCols = LIGHT ? 128 : (kBig ? 512 : 256);
  static constexpr int kAccCols = kBig ? 128 : 64;         // A stages start here; accumulator = columns [0, COUT)
  static constexpr int kThreads = (kProducerWarps + 2) * 32;
  static constexpr int kBBytes = 2 * COUT * 128;           // hi + lo image of one weight chunk
  static constexpr int kTileBytes = kRows * COUT * 4;      // the fp32 output tile staged by the epilogue
  static constexpr int kRingBytes = kStages * kBBytes > kTileBytes ? kStages * kBBytes : kTileBytes;
  static constexpr int kSmemBytes = kRingBytes + kRows * KOFF * 4 + 1536;
  static_assert(!LIGHT || COUT <= 64, "light CTAs: at most 64 output channels");
};

// SPLIT_IN: the input feature map is in the engine's pre-split format (common.cuh: every 4 channels = 16 bytes
// [hi0 hi1 hi2 hi3 | lo0 lo1 lo2 lo3] bf16 - what the producing kernel's epilogue wrote) AND has an all-zero row at
// index a.zero_row that absent neighbours point to: the gather is then 16-byte loads straight into the tcgen05.st
// registers - no conversion, no predication.  Otherwise: fp32 rows, predicated loads, bf16 hi/lo split in registers.
// MODE (compile time - a run-time switch in the prologue made the compiler emit a jump table per table entry and the
// dependent loads of consecutive entries serialised: 5-10 k cycles per tile): 0 identity rows (1x1x1), 1 27-neighbour
// table, 2 2x2x2 stride-2 children, 3 transposed 2x2x2 (parent row, kernel slice = the row's own child code).
template <int CIN, int COUT, int KOFF, bool SPLIT_IN, int MODE, bool LIGHT>
__global__ void __launch_bounds__(Cfg<CIN, COUT, KOFF, LIGHT>::kThreads, Cfg<CIN, COUT, KOFF, LIGHT>::kCtasPerSm) k_sconv_ts(Args a) {
  // MODE 0 with KOFF == 2: a row-wise layer with 2 * CIN input channels - "offset" k reads channel block k of the same row
  // (the (n, 2 CIN) map viewed as (2 n, CIN): source row 2 r + k).
  static_assert((MODE == 0 && (KOFF == 1 || KOFF == 2)) || (MODE == 1 && KOFF == 27) || ((MODE == 2 || MODE == 3) && KOFF == 8), "mode / offsets");
  using C = Cfg<CIN, COUT, KOFF, LIGHT>;
  constexpr int kStages = C::kStages, NG = C::kGroups;
  constexpr int NPW = C::kProducerWarps, NT = C::kThreads;
  constexpr int NCH = (KOFF * CIN + kChunk - 1) / kChunk;         // chunks if nothing is skipped
  constexpr int NBR_ITERS = (kRows * KOFF + NT - 1) / NT;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *btiles = smem;                                         // B ring: [kStages][hi | lo] (COUT*128 bytes each image)
  int *s_nbr = (int *)(btiles + C::kRingBytes);                   // [kRows][KOFF]
  uint64_t *full = (uint64_t *)(s_nbr + kRows * KOFF);            // [kStages]  A stage stored + weight chunk landed
  uint64_t *empty = full + kStages;                               // [kStages]  stage consumed by the tensor core
  uint64_t *accum = empty + kStages;                              // [1]
  uint32_t *s_tmem = (uint32_t *)(accum + 1);
  int *s_nlist = (int *)(s_tmem + 1);
  uint32_t *s_present = (uint32_t *)(s_nlist + 1);                // [2] bit j: chunk j has at least one present row
  int *s_list = (int *)(s_present + 2);                           // [NCH] compacted chunk ids
  int *s_rows = s_list + 56;                                      // [kRows] output row of every tile slot (tile row order, ctx.cuh)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kRows;
  const int col0 = blockIdx.y * COUT;     // N-split: this CTA computes output channels [col0, col0 + COUT)
  // debug timeline (EGN_TRACE=1, tools/trace_conv.py): clock64 stamps of one mid-grid CTA; a.trace is null in production
  // The stamps cost ~12 instructions per chunk in the producers' loop, so they exist only in a -DEGN_TRACE_BUILD library
  // (EGN_TRACE_BUILD=1 python -m egonn_b200.build --force).
#ifdef EGN_TRACE_BUILD
  const bool trc = a.trace != nullptr && blockIdx.x == (gridDim.x >> 1) && blockIdx.y == 0 && blockIdx.z == 0;
#else
  constexpr bool trc = false;
#endif
  if (trc && tid == 0) a.trace[60 * 8 + 0] = clock64();

  // neighbour rows of the tile: the global loads go out FIRST (their latency overlaps barrier setup and TMEM allocation),
  // in two branch-free batches - every thread first fetches the output rows of its table entries (tile slot -> row through
  // the level's tile row order), then the entries themselves: two memory latencies per tile, whatever the entry count.
  int src[NBR_ITERS];
  {
    int rowv[NBR_ITERS];
#pragma unroll
    for (int it = 0; it < NBR_ITERS; ++it) {
      const int t = tid + it * NT;
      const int r = t / KOFF;
      const bool ok = t < kRows * KOFF && row0 + r < a.n_out;
      rowv[it] = ok ? row0 + r : -1;
    }
    if (a.order != nullptr) {
#pragma unroll
      for (int it = 0; it < NBR_ITERS; ++it) {
        const int v = __ldg(a.order + (rowv[it] >= 0 ? rowv[it] : 0));
        rowv[it] = rowv[it] >= 0 ? v : -1;
      }
    }
    if constexpr (MODE == 0) {
#pragma unroll
      for (int it = 0; it < NBR_ITERS; ++it) {
        const int t = tid + it * NT;
        src[it] = rowv[it] >= 0 ? rowv[it] * KOFF + (t - (t / KOFF) * KOFF) : -1;
      }
    } else if constexpr (MODE == 1) {
#pragma unroll
      for (int it = 0; it < NBR_ITERS; ++it) {
        const int t = tid + it * NT;
        const int k = t - (t / KOFF) * KOFF;
        const int v = __ldg(a.nbr + (int64_t)(rowv[it] >= 0 ? rowv[it] : 0) * 27 + k);
        src[it] = rowv[it] >= 0 ? v : -1;
      }
    } else if constexpr (MODE == 2) {
      uint32_t m[NBR_ITERS];
      int cs[NBR_ITERS];
#pragma unroll
      for (int it = 0; it < NBR_ITERS; ++it) {
        const int rr = rowv[it] >= 0 ? rowv[it] : 0;
        m[it] = __ldg(a.cmask + rr);
        cs[it] = __ldg(a.cstart + rr);
      }
#pragma unroll
      for (int it = 0; it < NBR_ITERS; ++it) {
        const int k = (tid + it * NT) & 7;
        src[it] = (rowv[it] >= 0 && ((m[it] >> k) & 1u)) ? cs[it] + __popc(m[it] & ((1u << k) - 1u)) : -1;
      }
    } else {
      uint32_t code[NBR_ITERS];
      int par[NBR_ITERS];
#pragma unroll
      for (int it = 0; it < NBR_ITERS; ++it) {
        const int rr = rowv[it] >= 0 ? rowv[it] : 0;
        code[it] = (uint32_t)(__ldg(a.keys + rr) & 7ull);
        par[it] = __ldg(a.up + rr);
      }
#pragma unroll
      for (int it = 0; it < NBR_ITERS; ++it) {
        const int k = (tid + it * NT) & 7;
        src[it] = (rowv[it] >= 0 && (int)code[it] == k) ? par[it] : -1;
      }
    }
#pragma unroll
    for (int it = 0; it < NBR_ITERS; ++it) {
      const int t = tid + it * NT;
      const int r = t / KOFF;
      if (t < kRows * KOFF && t - r * KOFF == 0) s_rows[r] = rowv[it];     // -1 beyond the last row of the map
    }
  }
  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 4 + 1);     // the 4 warps of the producing group + the TMA thread's arrive.expect_tx
      mbar_init(&empty[s], 1);        // one tcgen05.commit
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == NPW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)C::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid < 2) s_present[tid] = 0u;
  __syncthreads();                    // s_present zeroed before the atomics below
  {
    uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
    for (int it = 0; it < NBR_ITERS; ++it) {
      const int t = tid + it * NT;
      if (t < kRows * KOFF) {
        s_nbr[t] = (SPLIT_IN && src[it] < 0) ? a.zero_row : src[it];
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
  if (warp == 0) {
    // compact the ids of the non-empty chunks inside this CTA's K-split range: lane handles chunks lane and lane + 32
    const int j_lo = (int)(((int64_t)NCH * blockIdx.z) / a.ksplit), j_hi = (int)(((int64_t)NCH * (blockIdx.z + 1)) / a.ksplit);
    const uint32_t pm0 = s_present[0], pm1 = s_present[1];
    const bool p0 = lane >= j_lo && lane < j_hi && ((pm0 >> lane) & 1u);
    const bool p1 = lane + 32 >= j_lo && lane + 32 < j_hi && ((pm1 >> lane) & 1u);
    const uint32_t b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
    const uint32_t below = (1u << lane) - 1u;
    if (p0) s_list[__popc(b0 & below)] = lane;
    if (p1) s_list[__popc(b0) + __popc(b1 & below)] = lane + 32;
    if (lane == 0) *s_nlist = __popc(b0) + __popc(b1);
  }
  __syncthreads();
  const int nlist = *s_nlist;
  const uint32_t tmem_base = *s_tmem;
  if (trc && tid == 0) { a.trace[60 * 8 + 1] = clock64(); a.trace[61 * 8 + 0] = nlist; }

  if (warp < NPW) {
    // ===================== A producers: gather -> bf16 hi/lo split -> tcgen05.st =====================
    // warp = 4*g + q: group g builds chunks g, g+NG, ...; this warp owns TMEM lanes 32q..32q+31 (= tile rows).
    // lane = 4*i8 + j4: per 16-lane half hf the thread feeds rows 32q + 16hf + i8 (+8), K positions 16m + 4*j4 .. +3.
    const int q = warp & 3, g = warp >> 2, i8 = lane >> 2, j4 = lane & 3;
    const int *nb_base = s_nbr + (32 * q + i8) * KOFF;
    for (int i = g; i < nlist; i += NG) {
      const int jc = s_list[i];
      const int s = i % kStages;
      const uint32_t ph = (uint32_t)(i / kStages) & 1u;
      const uint32_t tstage = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(C::kAccCols + 64 * s);
      const bool tr = trc && q == 0 && lane == 0 && i < 60;
      if (tr) a.trace[i * 8 + 3] = clock64();
      if constexpr (SPLIT_IN) {
        uint4 v[2][2][4];                                    // [half][row select][m]: (hi01, hi23, lo01, lo23) of 4 channels
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int rs = 0; rs < 2; ++rs) {
            const int *nb = nb_base + (16 * hf + 8 * rs) * KOFF;
            // lane j4 reads channels 8*j4 .. 8*j4+7 of each 32-channel block with one 256-bit load (weights.py: _perm32)
            if (CIN == 32) {
              const int k0 = 2 * jc, k1 = 2 * jc + 1;
              const int s0 = nb[k0], s1 = k1 < KOFF ? nb[k1] : a.zero_row;
              ldg256(a.in + (size_t)s0 * CIN + 8 * j4, v[hf][rs][0], v[hf][rs][1]);
              ldg256(a.in + (size_t)s1 * CIN + 8 * j4, v[hf][rs][2], v[hf][rs][3]);
            } else {
              const int k = CIN == 64 ? jc : (jc >> 1);
              const float *p0 = a.in + (size_t)nb[k] * CIN + (CIN == 128 ? (jc & 1) * 64 : 0) + 8 * j4;
              ldg256(p0, v[hf][rs][0], v[hf][rs][1]);
              ldg256(p0 + 32, v[hf][rs][2], v[hf][rs][3]);
            }
          }
        if (tr) a.trace[i * 8 + 4] = clock64();
        mbar_wait(&empty[s], ph ^ 1u, a.hint_producer);
        tc_fence_after();
        if (tr) a.trace[i * 8 + 5] = clock64();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int rs = 0; rs < 2; ++rs) {
              hi[4 * m + 2 * rs] = v[hf][rs][m].x; hi[4 * m + 2 * rs + 1] = v[hf][rs][m].y;
              lo[4 * m + 2 * rs] = v[hf][rs][m].z; lo[4 * m + 2 * rs + 1] = v[hf][rs][m].w;
            }
          tmem_st_16x256b_x4(tstage + ((uint32_t)(16 * hf) << 16), hi);
          tmem_st_16x256b_x4(tstage + ((uint32_t)(16 * hf) << 16) + 32, lo);
        }
      } else {
        float4 v[2][2][4];                                   // [half][row select][m]
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int rs = 0; rs < 2; ++rs) {
            const int *nb = nb_base + (16 * hf + 8 * rs) * KOFF;
            if (CIN == 32) {
              const int k0 = 2 * jc, k1 = 2 * jc + 1;
              const int s0 = nb[k0], s1 = k1 < KOFF ? nb[k1] : -1;
              ldg256_pred(a.in + (size_t)(s0 >= 0 ? s0 : 0) * CIN + 8 * j4, s0 >= 0, v[hf][rs][0], v[hf][rs][1]);
              ldg256_pred(a.in + (size_t)(s1 >= 0 ? s1 : 0) * CIN + 8 * j4, s1 >= 0, v[hf][rs][2], v[hf][rs][3]);
            } else {
              const int k = CIN == 64 ? jc : (jc >> 1);
              const int s0 = nb[k];
              const float *p0 = a.in + (size_t)(s0 >= 0 ? s0 : 0) * CIN + (CIN == 128 ? (jc & 1) * 64 : 0) + 8 * j4;
              ldg256_pred(p0, s0 >= 0, v[hf][rs][0], v[hf][rs][1]);
              ldg256_pred(p0 + 32, s0 >= 0, v[hf][rs][2], v[hf][rs][3]);
            }
          }
        mbar_wait(&empty[s], ph ^ 1u, a.hint_producer);
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int rs = 0; rs < 2; ++rs) {
              split2(v[hf][rs][m].x, v[hf][rs][m].y, hi[4 * m + 2 * rs], lo[4 * m + 2 * rs]);
              split2(v[hf][rs][m].z, v[hf][rs][m].w, hi[4 * m + 2 * rs + 1], lo[4 * m + 2 * rs + 1]);
            }
          tmem_st_16x256b_x4(tstage + ((uint32_t)(16 * hf) << 16), hi);
          tmem_st_16x256b_x4(tstage + ((uint32_t)(16 * hf) << 16) + 32, lo);
        }
      }
      if (tr) a.trace[i * 8 + 6] = clock64();
      tmem_st_wait();                   // this thread's tensor-memory stores have completed ...
      tc_fence_before();                // ... and are ordered before the arrive the MMA warp synchronises on
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      if (tr) a.trace[i * 8 + 7] = clock64();
    }
    if (trc && warp == 0 && lane == 0) a.trace[60 * 8 + 2] = clock64();
    // ===================== epilogue: TMEM -> shared memory (the idle weight ring) -> scale/shift/relu -> global ==============
    // A thread owns one accumulator ROW (TMEM lane), so storing straight from the tcgen05.ld registers touches 32
    // different output rows per instruction (one 16-byte sector each).  The tile is transposed through shared memory
    // instead - the weight ring is idle once the accumulator barrier has fired - and leaves as whole rows: a warp
    // instruction writes 512 contiguous bytes of 1, 2 or 4 output rows.  Row r's 16-byte unit u sits at unit u ^ (r & 7):
    // conflict-free both ways.
    constexpr int CPW = COUT / NG;                                           // accumulator columns per warp (>= 16)
    static_assert(CPW >= 16 && CPW % 16 == 0, "tcgen05.ld granularity: 16 columns");
    static_assert(C::kRingBytes >= kRows * COUT * 4, "the weight ring region must hold one fp32 output tile");
    if (nlist > 0) {
      mbar_wait(accum, 0u, a.hint_producer);
      tc_fence_after();
    }
    if (trc && warp == 0 && lane == 0) a.trace[60 * 8 + 3] = clock64();
    constexpr int UPR = COUT / 4;                                            // 16-byte units per output row
    float4 *stage = (float4 *)btiles;
    {
      const int r = q * 32 + lane;
#pragma unroll
      for (int cc = 0; cc < CPW; cc += 16) {
        const int c0 = g * CPW + cc;
        uint32_t v[16];
        if (nlist > 0) {
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
          if constexpr (C::kFold) {                            // + the hi*lo half of the folded accumulator
            uint32_t v2[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(COUT + c0), v2);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0u;             // no chunk of this split touches the tile: partial = 0
        }
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          const int u = (c0 >> 2) + gg;
          stage[r * UPR + (u ^ (r & 7))] = make_float4(__uint_as_float(v[gg * 4]), __uint_as_float(v[gg * 4 + 1]), __uint_as_float(v[gg * 4 + 2]),
                                                       __uint_as_float(v[gg * 4 + 3]));
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NPW * 32) : "memory");              // producer warps only: the tile is staged
    {
      constexpr int RPI = 32 / UPR > 0 ? 32 / UPR : 1;                       // rows per warp instruction (4, 2 or 1)
      constexpr int UPL = UPR / 32 > 0 ? UPR / 32 : 1;                       // units per lane per row (1)
      static_assert(UPL == 1, "at most 128 output channels per CTA");
      const int u = lane % UPR, rsub = lane / UPR;
      const int c = col0 + 4 * u;
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.scale) sc = __ldg((const float4 *)(a.scale + c));
      if (a.shift) sh = __ldg((const float4 *)(a.shift + c));
      float *obase = a.out + (size_t)blockIdx.z * a.n_out * a.cout_total;
#pragma unroll 4
      for (int rb = warp * RPI; rb < kRows; rb += NPW * RPI) {
        const int r = rb + rsub;
        const int row = s_rows[r];
        if (row < 0) continue;
        float4 y = stage[r * UPR + (u ^ (r & 7))];
        y.x = y.x * sc.x + sh.x; y.y = y.y * sc.y + sh.y; y.z = y.z * sc.z + sh.z; y.w = y.w * sc.w + sh.w;
        if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
        float *o = obase + (size_t)row * a.cout_total + c;
        if (a.accumulate) {
          const float4 prev = *(const float4 *)o;
          y.x += prev.x; y.y += prev.y; y.z += prev.z; y.w += prev.w;
        }
        if (a.out_split) *(uint4 *)o = presplit_pack(y);       // pre-split output: [hi0..3 | lo0..3] bf16 in the same 16 bytes
        else *(float4 *)o = y;
      }
    }
    // the all-zero row behind the last output row (what absent neighbours of the NEXT convolution point to)
    if (a.out_zero_row && blockIdx.x == 0 && blockIdx.z == 0 && warp == 0)
      for (int c = lane; c < COUT; c += 32) a.out[(size_t)a.n_out * a.cout_total + col0 + c] = 0.f;
    if (trc && warp == 0 && lane == 0) a.trace[60 * 8 + 4] = clock64();
    tc_fence_before();
  } else if (warp == NPW) {
    // ===================== B loader =====================
    if (lane == 0) {
      for (int i = 0; i < nlist; ++i) {
        const int j = s_list[i];
        const int sb = i % kStages;
        const uint32_t ph = (uint32_t)(i / kStages) & 1u;
        mbar_wait(&empty[sb], ph ^ 1u, a.hint_single);
        mbar_arrive_expect_tx(&full[sb], (uint32_t)C::kBBytes);
        const uint8_t *src = a.wpack + (size_t)j * 2 * a.cout_total * 128 + (size_t)col0 * 128;
        uint8_t *dst = btiles + sb * C::kBBytes;
        bulk_g2s(dst, src, (uint32_t)(COUT * 128), &full[sb]);
        bulk_g2s(dst + COUT * 128, src + (size_t)a.cout_total * 128, (uint32_t)(COUT * 128), &full[sb]);
      }
    }
  } else {
    // ===================== MMA issuer (whole warp converged, tcgen05 instructions by one elected lane) =====================
    // COUT == 32 folds hi*[Bhi;Blo] into ONE N=64 instruction (the two weight images are adjacent rows of the same
    // shared-memory tile); the epilogue adds the two 32-column halves: 8 instead of 12 instructions per chunk where the
    // issue rate (~30 cycles per tcgen05.mma), not the tensor pipe (16 cycles at N=32), is the limit.
    constexpr uint32_t idesc = umma_idesc(COUT), idesc2 = umma_idesc(2 * COUT);
    for (int i = 0; i < nlist; ++i) {
      const int s = i % kStages;
      const bool tr = trc && lane == 0 && i < 60;
      if (tr) a.trace[i * 8 + 0] = clock64();
      mbar_wait(&full[s], (uint32_t)(i / kStages) & 1u, a.hint_single);   // A stage in tensor memory AND the weight chunk have landed
      tc_fence_after();
      if (tr) a.trace[i * 8 + 1] = clock64();
      const uint32_t ta = tmem_base + (uint32_t)(C::kAccCols + 64 * s);
      const uint32_t sbm = smem_u32(btiles + s * C::kBBytes);
      const uint32_t first = i == 0 ? 0u : 1u;
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < kChunk / 16; ++ks) {
          const uint64_t bhi = umma_desc(sbm + ks * 32);
          if constexpr (C::kFold) {
            umma_f16_ts(tmem_base, ta + 8 * ks, bhi, idesc2, ks == 0 ? first : 1u);      // Ahi * [Bhi ; Blo]^T -> columns [0, 2 COUT)
            umma_f16_ts(tmem_base, ta + 32 + 8 * ks, bhi, idesc, 1u);                    // Alo * Bhi^T         -> columns [0, COUT)
          } else {
            const uint64_t blo = umma_desc(sbm + COUT * 128 + ks * 32);
            umma_f16_ts(tmem_base, ta + 8 * ks, bhi, idesc, ks == 0 ? first : 1u);
            umma_f16_ts(tmem_base, ta + 32 + 8 * ks, bhi, idesc, 1u);
            umma_f16_ts(tmem_base, ta + 8 * ks, blo, idesc, 1u);
          }
        }
        umma_commit(&empty[s]);          // TMEM stage and weight slot reusable once these MMAs have read them
      }
      __syncwarp();
      if (tr) a.trace[i * 8 + 2] = clock64();
    }
    if (nlist > 0 && elect_one_sync()) umma_commit(accum);   // accumulator complete
    __syncwarp();
  }
  __syncthreads();
  if (trc && tid == 0) a.trace[60 * 8 + 5] = clock64();
  if (warp == NPW + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols));
  }
}

template <int CIN, int COUT, int KOFF, bool SPLIT_IN, int MODE, bool LIGHT>
static int launch2(egn_ctx *ctx, const Args &a, const char *name, double bytes, double flops, cudaStream_t s) {
  using C = Cfg<CIN, COUT, KOFF, LIGHT>;
  EGN_SMEM_OPTIN(ctx, (k_sconv_ts<CIN, COUT, KOFF, SPLIT_IN, MODE, LIGHT>), C::kSmemBytes);
  const dim3 grid((unsigned)div_up(a.n_out, kRows), (unsigned)(a.cout_total / COUT), (unsigned)a.ksplit);
  EGN_LAUNCH(ctx, name, bytes, flops, s, k_sconv_ts<CIN, COUT, KOFF, SPLIT_IN, MODE, LIGHT><<<grid, C::kThreads, C::kSmemBytes, s>>>(a));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

template <int CIN, int COUT, int KOFF, bool SPLIT_IN, int MODE>
static int launch1(egn_ctx *ctx, const Args &a, const char *name, double bytes, double flops, cudaStream_t s) {
  // row-wise layers (1-2 chunks per tile): four light CTAs per SM, measured -5 % on the class (EGN_LIGHT=0: the standard shape);
  // the stride-2 / transposed convolutions (2-4 chunks) lose 6-8 % with a single producer group and stay standard
  if constexpr (MODE == 0 && COUT <= 64) {
    if (ctx->light_ctas) return launch2<CIN, COUT, KOFF, SPLIT_IN, MODE, true>(ctx, a, name, bytes, flops, s);
  }
  return launch2<CIN, COUT, KOFF, SPLIT_IN, MODE, false>(ctx, a, name, bytes, flops, s);
}

template <int CIN, int COUT, int KOFF, int MODE>
static int launch(egn_ctx *ctx, const Args &a, const char *name, double bytes, double flops, cudaStream_t s) {
  return a.in_split ? launch1<CIN, COUT, KOFF, true, MODE>(ctx, a, name, bytes, flops, s)
                    : launch1<CIN, COUT, KOFF, false, MODE>(ctx, a, name, bytes, flops, s);
}

}  // namespace ts

// dispatch of one (CIN, COUT-per-CTA, KOFF) instance; returns EGN_ERR_INVALID when there is no such instance
int launch_conv_ts(egn_ctx *ctx, int koff, int cin, int cout_cta, const tcx::Args &a, const char *name, double bytes, double flops,
                   cudaStream_t s) {
#define EGN_TS_CASE(KO, CI, CO)                                                                                         \
  if (koff == KO && cin == CI && cout_cta == CO) {                                                                      \
    if (KO == 8 && a.mode == 3) return ts::launch<CI, CO, KO, KO == 8 ? 3 : (KO == 27 ? 1 : 0)>(ctx, a, name, bytes, flops, s); \
    return ts::launch<CI, CO, KO, KO == 27 ? 1 : (KO == 8 ? 2 : 0)>(ctx, a, name, bytes, flops, s);                     \
  }
  EGN_TS_CASE(27, 32, 32)
  EGN_TS_CASE(27, 32, 64)
  EGN_TS_CASE(27, 64, 64)
  EGN_TS_CASE(27, 64, 128)
  EGN_TS_CASE(27, 128, 128)
  EGN_TS_CASE(27, 128, 64)
  EGN_TS_CASE(27, 128, 32)
  EGN_TS_CASE(8, 32, 32)
  EGN_TS_CASE(8, 64, 64)
  EGN_TS_CASE(8, 128, 128)
  EGN_TS_CASE(8, 128, 64)
  EGN_TS_CASE(8, 128, 32)
  EGN_TS_CASE(2, 128, 128)
  EGN_TS_CASE(1, 64, 32)
  EGN_TS_CASE(1, 32, 64)
  EGN_TS_CASE(1, 64, 64)
  EGN_TS_CASE(1, 64, 128)
  EGN_TS_CASE(1, 128, 64)
  EGN_TS_CASE(1, 128, 128)
#undef EGN_TS_CASE
  EGN_CHECK(false, EGN_ERR_INVALID, "tensor-core conv (TMEM-A): no kernel instance k=%d %d->%d", koff, cin, cout_cta);
}

}  // namespace egn
