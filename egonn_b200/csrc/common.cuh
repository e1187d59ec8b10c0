// Shared helpers for the egonn_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/egonn_b200.h"

namespace egn {

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define EGN_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::egn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));      \
      return EGN_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

#define EGN_CHECK(cond, code, ...)                                                                 \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      ::egn::set_error(__VA_ARGS__);                                                               \
      return (code);                                                                               \
    }                                                                                              \
  } while (0)

#define EGN_TRY(expr)                                                                              \
  do {                                                                                             \
    int _s = (expr);                                                                               \
    if (_s != EGN_OK) return _s;                                                                   \
  } while (0)

// make `dev` the current device for the scope (contexts and communicators belong to one device)
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ---- key layout --------------------------------------------------------------------------------------
// level-0 key = (batch << 54) | morton54(ux,uy,uz), u = c + 2^17 in [0, 2^18); x owns the lowest bit of
// every 3-bit group, so the 3 low bits of a level-L key are exactly MinkowskiEngine's kernel index of
// the 2x2x2 stride-2 region (k = dx + 2dy + 4dz) and key >> 3 is the parent voxel.
constexpr int kAxisBits = 18;
constexpr int kAxisBias = 1 << (kAxisBits - 1);
constexpr int kMortonBits = 3 * kAxisBits;  // 54
constexpr int kMaxBatch = 1023;             // batch index must be < 1023 (key 0xFFFF.. is never produced)
constexpr int kNumSMs = 148;

__host__ __device__ __forceinline__ uint64_t spread3(uint32_t v) {
  uint64_t x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__host__ __device__ __forceinline__ uint32_t compact3(uint64_t x) {
  x &= 0x1249249249249249ull;
  x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
  x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
  x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
  x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
  x = (x ^ (x >> 32)) & 0x1fffffull;
  return (uint32_t)x;
}
// v* are level-L voxel indices (u >> L), each < 2^(18-L)
__host__ __device__ __forceinline__ uint64_t make_key(int level, uint32_t b, uint32_t vx, uint32_t vy, uint32_t vz) {
  return ((uint64_t)b << (kMortonBits - 3 * level)) | spread3(vx) | (spread3(vy) << 1) | (spread3(vz) << 2);
}
__host__ __device__ __forceinline__ void split_key(int level, uint64_t key, uint32_t &b, uint32_t &vx, uint32_t &vy,
                                                   uint32_t &vz) {
  const int mb = kMortonBits - 3 * level;
  b = (uint32_t)(key >> mb);
  const uint64_t m = key & ((1ull << mb) - 1ull);
  vx = compact3(m);
  vy = compact3(m >> 1);
  vz = compact3(m >> 2);
}

__host__ __device__ __forceinline__ int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int grid_for(int64_t work_items, int threads, int max_waves = 8) {
  int64_t g = div_up(work_items, threads);
  int64_t cap = (int64_t)kNumSMs * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- bump arena (engine-owned scratch, grows only) -----------------------------------------------------
struct Arena {
  char *base = nullptr;
  size_t cap = 0, used = 0;
  int reserve(size_t bytes, cudaStream_t stream);  // ensure cap >= bytes (may sync + realloc), reset used
  void *take(size_t bytes) {                        // 256-byte aligned carve; nullptr if exhausted
    size_t a = (used + 255) & ~size_t(255);
    if (a + bytes > cap) return nullptr;
    used = a + bytes;
    return base + a;
  }
  void release();
};
inline size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

}  // namespace egn
