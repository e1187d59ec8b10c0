// conv0 (5x5x5, Cin = 1 -> 32, + BatchNorm + ReLU; models/minkgl.py:100-102,140-142) on the tensor cores, for the case every
// EgoNN caller feeds: all input features are 1.0 (eval/evaluate.py:334, datasets/dataset_utils.py:80).
//
//   out[o, :] = sum_k present(o + d_k) * W[k, 0, :]      =      M[o, 0:125] @ W[0:125, 0:32],   M in {0, 1}
//
// The FP32 kernel (ops.cu: k_conv0) spends its time in shared memory: every present neighbour costs each lane a 128-byte
// kernel row (8 x LDS.128, different row per lane - no broadcast): ncu shows the LSU data pipe at 71 % with DRAM at 5 %.
// Here the presence matrix is the A operand of a GEMM:
//   * lane = output row.  Phase 1: the 125-bit presence mask of the row's window from the occupancy words of the <= 8
//     level-2 cells it touches - 25 window rows of 5 bits each, extracted with shifts (no per-neighbour loop: k_conv0's
//     bit loop runs to the longest lane and was 3200 of its 3700 instructions per 32 rows).
//   * the mask expands to bf16 {0, 1} - EXACT, so A needs no hi/lo split - and goes to tensor memory with two
//     tcgen05.st.32x32b.x32 (lane = row = TMEM lane, registers = consecutive K columns: no transposition needed).
//   * B = the kernel as the usual pre-swizzled bf16 hi/lo image (weights.py: pack_tc, 2 chunks of 64 offsets); hi and lo
//     are adjacent rows of the same shared-memory tile, so ONE N = 64 tcgen05.mma per K-step computes M*Whi | M*Wlo and the
//     epilogue adds the halves: 8 instructions per 128-row tile, FP32 accumulation, error <= 2^-17 per weight.
//   * epilogue: tcgen05.ld -> BatchNorm scale/shift -> ReLU -> the output row (fp32 or pre-split), written once.
// One 128-row tile per CTA, 6 warps (4 row warps, 1 MMA issuer, 1 spare for the TMEM allocation), 128 TMEM columns and
// ~17 KB of shared memory per CTA: 4 CTAs per SM.
#include "ctx.cuh"
#include "tc_ptx.cuh"

namespace egn {

namespace c0tc {

using namespace tcx;

constexpr int kThreads = 192;
constexpr int kBBytes = 2 * (2 * 32 * 128);        // 2 chunks x (hi image + lo image) x 32 rows x 128 bytes
constexpr int kSmemBytes = kBBytes + 64 + 1024;

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

// occupancy word of a 4x4x4 cell: bit index x0 | y0<<1 | z0<<2 | x1<<3 | y1<<4 | z1<<5 (Morton) -> x | y<<2 | z<<4.
// An index-bit transposition (i, j) is one delta swap: exchange the positions with bit i = 1, bit j = 0 with their partners.
__host__ __device__ constexpr unsigned long long swap_mask(int i, int j) {
  unsigned long long m = 0ull;
  for (int p = 0; p < 64; ++p)
    if (((p >> i) & 1) && !((p >> j) & 1)) m |= 1ull << p;
  return m;
}
template <int I, int J>
__device__ __forceinline__ unsigned long long index_bit_swap(unsigned long long w) {
  constexpr unsigned long long M = swap_mask(I, J);
  constexpr int D = (1 << J) - (1 << I);
  const unsigned long long t = ((w >> D) ^ w) & M;
  return w ^ t ^ (t << D);
}
__device__ __forceinline__ unsigned long long demorton64(unsigned long long w) {
  // index bits [x0 y0 z0 x1 y1 z1] -> (1,3) [x0 x1 z0 y0 y1 z1] -> (2,3) [x0 x1 y0 z0 y1 z1] -> (3,4) [x0 x1 y0 y1 z0 z1]
  return index_bit_swap<3, 4>(index_bit_swap<2, 3>(index_bit_swap<1, 3>(w)));
}

// 64 presence bits -> 32 registers of two bf16 {0.0, 1.0} each (low half = even K position)
__device__ __forceinline__ void expand_bits(unsigned long long w, uint32_t (&r)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const uint32_t two = (uint32_t)(w >> (2 * j)) & 3u;
    r[j] = (two & 1u) * 0x3F80u + (two >> 1) * 0x3F800000u;
  }
}

__global__ void __launch_bounds__(kThreads, 4)
    k_conv0_tc(const uint64_t *__restrict__ keys0, const int *__restrict__ up0, const int *__restrict__ up1, const int *__restrict__ nbr2,
               const uint64_t *__restrict__ mask64, int n0, const uint8_t *__restrict__ wpack, const float *__restrict__ scale,
               const float *__restrict__ shift, int relu, const int *__restrict__ not_ones, int out_split, float *__restrict__ out) {
  constexpr int KS = 5, R = 2, COUT = 32;
  if (not_ones != nullptr && *not_ones != 0) return;            // some feature != 1: the general FP32 variant does the work
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *btiles = smem;                                                         // [chunk 2][hi 4 KB | lo 4 KB]
  uint64_t *bars = (uint64_t *)(btiles + kBBytes);                                 // b_full, a_full, accum
  uint32_t *s_tmem = (uint32_t *)(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kRows;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 4);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(&bars[0], (uint32_t)kBBytes);
    bulk_g2s(btiles, wpack, (uint32_t)kBBytes, &bars[0]);          // both chunks are contiguous in the packed image
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (out_split && blockIdx.x == 0 && tid < COUT) out[(size_t)n0 * COUT + tid] = 0.f;      // the zero row of a pre-split map
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < 4) {
    // ---------------- phase 1: 125-bit presence mask of the row's 5x5x5 window (no per-bit loop, no divergence) ----------------
    // The window spans at most 2 cells per axis.  Each cell's 64-bit occupancy word is re-ordered once from Morton order to
    // x + 4y + 16z (three delta swaps); a window row (fixed ty, tz) is then the 4 x-bits of cell dx0 and the 4 of cell dx1
    // at (y, z), shifted by the row's x phase: 5 bits that land at the compile-time position 5 ty + 25 tz of the mask.
    const int r = row0 + warp * 32 + lane;
    unsigned long long m_lo = 0ull, m_hi = 0ull;
    if (r < n0) {
      const uint32_t m = (uint32_t)(keys0[r] & 63ull);
      const int lx = (m & 1) | ((m >> 2) & 2), ly = ((m >> 1) & 1) | ((m >> 3) & 2), lz = ((m >> 2) & 1) | ((m >> 4) & 2);
      const int cell = up1[up0[r]];
      const int *nb = nbr2 + (int64_t)cell * 27;
      const int dx0 = (lx - R) >> 2, dx1 = (lx + R) >> 2, dy0 = (ly - R) >> 2, dy1 = (ly + R) >> 2, dz0 = (lz - R) >> 2,
                dz1 = (lz + R) >> 2;
      unsigned long long lin[8];                              // [cz][cy][cx] occupancy in x + 4y + 16z order (0: absent / duplicate cell)
      int q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int dx = (i & 1) ? dx1 : dx0, dy = (i & 2) ? dy1 : dy0, dz = (i & 4) ? dz1 : dz0;
        const bool dup = ((i & 1) && dx1 == dx0) || ((i & 2) && dy1 == dy0) || ((i & 4) && dz1 == dz0);
        q[i] = dup ? -1 : nb[(dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) lin[i] = q[i] >= 0 ? demorton64(mask64[q[i]]) : 0ull;
      const int sx = lx - R - 4 * dx0;                        // x phase of the window inside the 8-wide cell pair: 0..3
#pragma unroll
      for (int tz = 0; tz < KS; ++tz) {
        const int z = lz - R + tz, cz = (z >> 2) - dz0, zz = z & 3;             // cz in {0, 1}
#pragma unroll
        for (int ty = 0; ty < KS; ++ty) {
          const int y = ly - R + ty, cy = (y >> 2) - dy0, yy = y & 3;
          const unsigned long long w0 = cz ? (cy ? lin[6] : lin[4]) : (cy ? lin[2] : lin[0]);
          const unsigned long long w1 = cz ? (cy ? lin[7] : lin[5]) : (cy ? lin[3] : lin[1]);
          const int sh = 4 * yy + 16 * zz;
          const uint32_t bits8 = ((uint32_t)(w0 >> sh) & 0xFu) | (((uint32_t)(w1 >> sh) & 0xFu) << 4);
          const unsigned long long row5 = (unsigned long long)((bits8 >> sx) & 0x1Fu);
          constexpr int kDummy = 0; (void)kDummy;
          const int t0 = KS * ty + KS * KS * tz;              // compile-time after unrolling
          if (t0 < 64) m_lo |= row5 << t0;
          if (t0 + KS > 64) m_hi |= t0 >= 64 ? row5 << (t0 - 64) : row5 >> (64 - t0);
        }
      }
    }
    // ---------------- A operand: two chunks of 64 K positions, bf16 {0,1}, lane = row = TMEM lane ----------------
    const uint32_t ta = tmem_base + ((uint32_t)(32 * warp) << 16) + 64u;
    {
      uint32_t rr[32];
      expand_bits(m_lo, rr);
      tmem_st_32x32b_x32(ta, rr);
      expand_bits(m_hi, rr);
      tmem_st_32x32b_x32(ta + 32, rr);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[1]);
    // ---------------- epilogue ----------------
    mbar_wait(&bars[2], 0u, 0u);
    tc_fence_after();
    uint32_t acc[32], acc2[32];
    tmem_ld32(tmem_base + ((uint32_t)(32 * warp) << 16), acc);          // M * Whi
    tmem_ld32(tmem_base + ((uint32_t)(32 * warp) << 16) + 32u, acc2);   // M * Wlo
    if (r < n0) {
      float4 *o = (float4 *)(out + (size_t)r * COUT);
#pragma unroll
      for (int c4 = 0; c4 < COUT / 4; ++c4) {
        float4 y;
        float *yy = (float *)&y;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c4 * 4 + e;
          float v = __uint_as_float(acc[c]) + __uint_as_float(acc2[c]);
          v = v * (scale ? __ldg(scale + c) : 1.f) + (shift ? __ldg(shift + c) : 0.f);
          if (relu) v = fmaxf(v, 0.f);
          yy[e] = v;
        }
        if (out_split) ((uint4 *)o)[c4] = presplit_pack(y); else o[c4] = y;
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ---------------- MMA issuer: D[:, 0:64] = M * [Whi ; Wlo]^T over 2 chunks x 4 K-steps ----------------
    mbar_wait(&bars[0], 0u, 0u);
    mbar_wait(&bars[1], 0u, 0u);
    tc_fence_after();
    constexpr uint32_t idesc2 = umma_idesc(2 * COUT);
    if (elect_one_sync()) {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bd = umma_desc(smem_u32(btiles + ch * (kBBytes / 2)) + ks * 32);
          umma_f16_ts(tmem_base, tmem_base + 64u + 32u * ch + 8u * ks, bd, idesc2, (ch | ks) ? 1u : 0u);
        }
      umma_commit(&bars[2]);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
  }
}

}  // namespace c0tc

// tensor-core conv0 for all-ones features (5x5x5, 32 output channels); exits at once when *not_ones != 0
int run_conv0_tc(egn_ctx *ctx, const void *wpack, const float *scale, const float *shift, int relu, const int *not_ones, int out_split,
                 float *out, double bytes, double flops, cudaStream_t s) {
  const Pyramid &py = ctx->pyr;
  EGN_SMEM_OPTIN(ctx, c0tc::k_conv0_tc, c0tc::kSmemBytes);
  const int n0 = py.n[0];
  EGN_LAUNCH(ctx, "conv0_5x5x5", bytes, flops, s,
             c0tc::k_conv0_tc<<<(unsigned)div_up(n0, tcx::kRows), c0tc::kThreads, c0tc::kSmemBytes, s>>>(
                 py.keys[0], py.up[0], py.up[1], py.nbr[2], py.mask64, n0, (const uint8_t *)wpack, scale, shift, relu, not_ones, out_split, out));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

}  // namespace egn
