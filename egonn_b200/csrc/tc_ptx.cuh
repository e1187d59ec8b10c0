// PTX wrappers shared by the tcgen05 convolution kernels (sm_100a): mbarrier, TMA bulk copy, UMMA descriptors,
// tcgen05.mma / ld / st, bf16 hi/lo split, predicated gathers.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace egn {
namespace tcx {

constexpr int kRows = 128;           // tile rows (UMMA M)
constexpr int kChunk = 64;           // K elements per chunk (128 bytes of bf16)
constexpr int kABytes = kRows * 128; // one A image (hi or lo) of a shared-memory stage

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// hint_ns > 0: try_wait may suspend the thread in hardware up to hint_ns; 0: plain polling try_wait
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t done;
  if (hint_ns) {
    do {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
          : "memory");
    } while (!done);
  } else {
    do {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
    } while (!done);
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// LBO = 1 (ignored for swizzled K-major), SBO = 1024 B (8 rows x 128 B), version 1, layout type 2.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a fully converged warp (the compiler keeps operands of the guarded instruction in uniform registers)
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// bf16 hi/lo split: cvt.rn.bf16x2.f32 converts two values at once; lo = bf16(x - float(hi))
__device__ __forceinline__ void split2(float x, float y, uint32_t &hi, uint32_t &lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);                 // .x (low half) = x
  hi = *reinterpret_cast<uint32_t *>(&h);
  const float fx = __uint_as_float(hi << 16), fy = __uint_as_float(hi & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(x - fx, y - fy);
  lo = *reinterpret_cast<uint32_t *>(&l);
}
// 32-byte gather of 8 consecutive floats, predicated (no branch): zeros when the neighbour is absent
__device__ __forceinline__ void ldg8_pred(const float *p, bool pred, float4 &a, float4 &b) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %9, 0;\n\t"
      "mov.b32 %0, 0; mov.b32 %1, 0; mov.b32 %2, 0; mov.b32 %3, 0;\n\t"
      "mov.b32 %4, 0; mov.b32 %5, 0; mov.b32 %6, 0; mov.b32 %7, 0;\n\t"
      "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%8];\n\t"
      "@q ld.global.nc.v4.f32 {%4, %5, %6, %7}, [%8+16];\n\t"
      "}"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p), "r"((uint32_t)pred));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


// ---- A operand in tensor memory ("TS" form of tcgen05.mma) ------------------------------------------------------------
// A(m,k) of an M=128 tile lives at TMEM lane m, 16-bit column k: two bf16 per 32-bit column, low half = even k
// (cute::UMMA::tmem_frg_1sm<bf16,bf16>; checked on the device by tools/tmem_probe.cu).  One K=16 step = 8 columns.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.st.16x256b: the warp writes 16 lanes x (8*N columns).  Thread t = 4*i + j supplies, for n = 0..N-1:
//   r[4n+0], r[4n+1] -> lane base+i,   columns 8n + 2j, 8n + 2j + 1
//   r[4n+2], r[4n+3] -> lane base+i+8, columns 8n + 2j, 8n + 2j + 1          (probe: tools/tmem_probe.cu test 1)
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 16-byte gather, predicated (no branch): zeros when the neighbour is absent
__device__ __forceinline__ void ldg4_pred(const float *p, bool pred, float4 &a) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "mov.b32 %0, 0; mov.b32 %1, 0; mov.b32 %2, 0; mov.b32 %3, 0;\n\t"
      "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t"
      "}"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
      : "l"(p), "r"((uint32_t)pred));
}

// 256-bit global loads (sm_100: LDG.E.256): 32 contiguous bytes per lane -> the four lanes of a row cover a whole 128-byte
// line with ONE wavefront.  The address must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const void *p, uint4 &a, uint4 &b) {
  asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void ldg256_pred(const float *p, bool pred, float4 &a, float4 &b) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %9, 0;\n\t"
      "mov.b32 %0, 0; mov.b32 %1, 0; mov.b32 %2, 0; mov.b32 %3, 0;\n\t"
      "mov.b32 %4, 0; mov.b32 %5, 0; mov.b32 %6, 0; mov.b32 %7, 0;\n\t"
      "@q ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
      "}"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p), "r"((uint32_t)pred));
}

// ---- pre-split feature maps ---------------------------------------------------------------------------------------------
// A feature map that only tensor-core convolutions gather from is stored PRE-SPLIT: same (n, C) x 4-byte footprint as
// fp32, but every group of 4 channels is the 16 bytes [hi0 hi1 hi2 hi3 | lo0 lo1 lo2 lo3] (bf16; x = hi + lo to 2^-17
// relative - exactly the rounding the tensor-core path applies to its A operand anyway).  The producing kernel's
// epilogue splits each value ONCE; the ~7 gathers per value of a 3x3x3 convolution then move bits only.  Such a map
// also carries one all-zero row at index n, which absent neighbours point to (no predication in the gather).
__device__ __forceinline__ uint4 presplit_pack(float4 y) {
  uint4 u;
  split2(y.x, y.y, u.x, u.z);
  split2(y.z, y.w, u.y, u.w);
  return u;
}
__device__ __forceinline__ float4 presplit_unpack(uint4 u) {
  float4 y;
  y.x = __uint_as_float(u.x << 16) + __uint_as_float(u.z << 16);
  y.y = __uint_as_float(u.x & 0xffff0000u) + __uint_as_float(u.z & 0xffff0000u);
  y.z = __uint_as_float(u.y << 16) + __uint_as_float(u.w << 16);
  y.w = __uint_as_float(u.y & 0xffff0000u) + __uint_as_float(u.w & 0xffff0000u);
  return y;
}

// launch arguments of the tensor-core convolution kernels (k_sconv_tc: A staged in shared memory, k_sconv_ts: A in tensor memory)
struct Args {
  const float *in;
  float *out;
  const uint8_t *wpack;  // [n_chunks][hi|lo][COUT][64] bf16, swizzled shared-memory images
  const float *scale, *shift;
  long long *trace;                      // debug: clock64 timeline of CTA 0 (null in production)
  uint32_t hint_producer, hint_single;   // try_wait suspend hints (ns) for the gather warps / the TMA and MMA threads
  int n_out, relu, accumulate, cout_total, ksplit, mode;   // ksplit > 1: grid.z partitions the chunks, raw partials to out + z*n_out*cout_total  // mode 0: identity rows (1x1x1 convolution), 1: 27-neighbour table, 2: 2x2x2 stride-2 children, 3: transposed 2x2x2 (parent, slice = own code)
  const int *order;        // tile slot -> output row (null: identity); coords.cu "tile row orders"
  const int *nbr;
  const int *cstart;
  const uint32_t *cmask;
  const int *up;           // mode 3: parent row of every output (fine) row
  const uint64_t *keys;    // mode 3: key of every output row (kernel slice = key & 7)
  // k_sconv_ts only: pre-split feature maps (common.cuh, "pre-split format")
  int in_split;            // input rows are pre-split AND row zero_row of the input is all zeros
  int zero_row;
  int out_split;           // write the output pre-split
  int out_zero_row;        // also write an all-zero row at index n_out of the output
};

}  // namespace tcx
}  // namespace egn
