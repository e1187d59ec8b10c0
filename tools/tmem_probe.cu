// Probe of two tcgen05 facts the TMEM-resident-A convolution relies on (no public docs in this image, so measure):
//   1. the thread <-> (lane, column) mapping of tcgen05.st.16x256b.x4 (checked by reading back with tcgen05.ld.32x32b)
//   2. tcgen05.mma kind::f16 with the A operand in tensor memory ("TS" form): A(m,k) lives at lane m, 16-bit
//      column k (two bf16 per 32-bit column, low half = even k); D = A * B^T against a host reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmem_probe tools/tmem_probe.cu ; run on a B200.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

constexpr int N = 32, K = 64;

// out_map[lane][col] (128 x 32): value read back after the 16x256b.x4 stores; out_d[128][N]: MMA result
__global__ void __launch_bounds__(192) k_probe(const float *__restrict__ a /* [128][K] bf16-representable */,
                                               const uint8_t *__restrict__ bimg /* swizzled [N][64] bf16 */, uint32_t *out_map,
                                               float *out_d) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < N * 128; i += blockDim.x) smem[i] = bimg[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = s_tmem;
  const uint32_t colA = 64;   // A image at columns [64, 96), D at [0, 32)

  if (warp < 4) {
    // ---- test 1 + A staging: thread (i = lane/4, j = lane%4), lane group lg: rows 32*warp + lg + i and + 8 ----
    const int i = lane >> 2, j = lane & 3;
#pragma unroll
    for (int lg = 0; lg < 32; lg += 16) {
      uint32_t r[16];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          const int row = 32 * warp + lg + i + 8 * rs;
          const float *p = a + row * K + 16 * m + 4 * j;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(p[0], p[1]), h1 = __floats2bfloat162_rn(p[2], p[3]);
          r[4 * m + 2 * rs + 0] = *reinterpret_cast<uint32_t *>(&h0);
          r[4 * m + 2 * rs + 1] = *reinterpret_cast<uint32_t *>(&h1);
        }
      const uint32_t taddr = tb + ((uint32_t)(32 * warp + lg) << 16) + colA;
      asm volatile(
          "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    // read back with 32x32b.x32: thread = lane (32*warp + lane), register c = column colA + c
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(tb + ((uint32_t)(32 * warp) << 16) + colA));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 32; ++c) out_map[(32 * warp + lane) * 32 + c] = v[c];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == 5) {
    // ---- test 2: D[128][N] = A(TMEM) * B(smem)^T, K = 64 in four K=16 steps ----
    uint32_t elected;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(elected));
    if (elected) {
      const uint32_t sb = smem_u32(smem);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t bd = umma_desc(sb + ks * 32);
        const uint32_t at = tb + colA + ks * 8;
        const uint32_t acc = ks ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tb),
            "r"(at), "l"(bd), "r"(umma_idesc(N)), "r"(acc)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
  }
  if (warp < 4) {
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(&bar)), "r"(0u)
                   : "memory");
    } while (!done);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(tb + ((uint32_t)(32 * warp) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < N; ++c) out_d[(32 * warp + lane) * N + c] = __uint_as_float(v[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128u));
  }
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<float> a(128 * K), b(N * K);
  srand(1);
  for (auto &v : a) v = bf16r((float)(rand() % 2001 - 1000) / 500.f);
  for (auto &v : b) v = bf16r((float)(rand() % 2001 - 1000) / 500.f);
  // B image: [N rows][64 bf16], 16-byte group g of row n stored at group g ^ (n & 7)
  std::vector<uint8_t> bimg(N * 128);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 h = __float2bfloat16(b[n * K + k]);
      const int g = k / 8, e = k % 8, gs = g ^ (n & 7);
      memcpy(&bimg[n * 128 + gs * 16 + e * 2], &h, 2);
    }
  float *da, *dd;
  uint8_t *db;
  uint32_t *dm;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, bimg.size()); cudaMalloc(&dm, 128 * 32 * 4); cudaMalloc(&dd, 128 * N * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, bimg.data(), bimg.size(), cudaMemcpyHostToDevice);
  cudaMemset(dm, 0xff, 128 * 32 * 4); cudaMemset(dd, 0, 128 * N * 4);
  k_probe<<<1, 192, N * 128 + 1024>>>(da, db, dm, dd);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint32_t> m(128 * 32);
  std::vector<float> d(128 * N);
  cudaMemcpy(m.data(), dm, m.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
  // test 1: column c of lane r must hold bf16 pair (a[r][2c], a[r][2c+1]), low half = even k
  int bad1 = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 32; ++c) {
      __nv_bfloat16 lo = __float2bfloat16(a[r * K + 2 * c]), hi = __float2bfloat16(a[r * K + 2 * c + 1]);
      uint16_t ulo, uhi;
      memcpy(&ulo, &lo, 2); memcpy(&uhi, &hi, 2);
      const uint32_t exp = (uint32_t)ulo | ((uint32_t)uhi << 16);
      if (m[r * 32 + c] != exp) {
        if (bad1 < 8) printf("  map mismatch lane %d col %d: got %08x expected %08x\n", r, c, m[r * 32 + c], exp);
        ++bad1;
      }
    }
  printf("test1 (tcgen05.st.16x256b.x4 fragment -> lane/column): %d mismatches of %d\n", bad1, 128 * 32);
  // test 2
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)a[r * K + k] * b[n * K + k];
      maxerr = fmax(maxerr, fabs(s - d[r * N + n]));
      maxref = fmax(maxref, fabs(s));
    }
  printf("test2 (tcgen05.mma A from TMEM): max|err| = %.3e, max|ref| = %.3e -> %s\n", maxerr, maxref,
         maxerr <= 1e-4 * maxref ? "OK" : "MISMATCH");
  return (bad1 == 0 && maxerr <= 1e-4 * maxref) ? 0 : 2;
}
