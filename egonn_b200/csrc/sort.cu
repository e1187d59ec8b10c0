// In-tree stable LSD radix sort of (uint64 key, uint32 value) pairs - the ONE sort behind the coordinate pyramid
// (ME.SparseTensor's coordinate hash-table build, models/minkgl.py:269) and behind sparse_quantize's de-duplication
// (datasets/quantization.py:42,83).  Replaces cub::DeviceRadixSort.
//
// Sizes here are 0.75 M - 5 M pairs, i.e. 180 - 1200 tiles of 4096: at that size a decoupled look-back chain (onesweep) is
// one serial hop per tile, and a histogram + scan + scatter triple is three launches per digit.  This sort needs ONE
// launch per 8-bit digit and no spinning:
//   * the digit counts of every (tile, digit) of pass p+1 are accumulated BY PASS p while it scatters: an element that
//     lands at output position q belongs to tile q / 4096 of the next pass, so the scattering warp adds its elements to
//     hist[p+1][q / 4096][next digit] (integer atomics, aggregated over equal (tile, digit) pairs inside the warp with
//     __match_any_sync: order-independent, deterministic).  Pass 0's counts come from a counting kernel.  A second, 16x
//     coarser count matrix ("super-tiles") keeps the prefix walk short.
//   * a pass CTA first turns the count matrices into its 256 output bases (digit total before it + same-digit elements of
//     earlier tiles: <= n_super + 15 coalesced loads per thread, no waiting on other CTAs), then ranks its 4096 elements
//     stably (a warp owns 256 consecutive elements and ranks them in order with __match_any_sync; per-(digit, warp)
//     counters are scanned digit-major), stages the tile in digit order in shared memory and writes it out in runs
//     (consecutive lanes -> consecutive addresses).
//   * a digit position whose value is the same for all keys (known from the count matrix) costs a tile copy, not a pass.
// Stable, deterministic, no co-residency assumption (safe under concurrent streams), 24 bytes of traffic per pair and pass.
#include "ctx.cuh"

namespace egn {

namespace {

constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;       // 16
constexpr int kSortItems = 8;                       // per thread: 64 registers, two CTAs per SM
constexpr int kSortTile = kSortThreads * kSortItems;   // 4096
constexpr int kSortSuper = 16;                      // tiles per super-tile

// pass-0 counts: hist_tiles[tile][digit], hist_super[tile / 16][digit]
__global__ void __launch_bounds__(kSortThreads) k_sort_count(const uint64_t *__restrict__ keys, int n, int shift, int *__restrict__ hist_tiles,
                                                             int *__restrict__ hist_super) {
  __shared__ int cnt[256];
  const int tile = blockIdx.x;
  if (threadIdx.x < 256) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int base = tile * kSortTile;
  const int lane = threadIdx.x & 31;
  uint32_t d[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {                           // all loads first: one memory latency per tile
    const int e = base + r * kSortThreads + threadIdx.x;
    d[r] = e < n ? (uint32_t)((keys[e] >> shift) & 255ull) : 256u;
  }
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const uint32_t peers = __match_any_sync(0xffffffffu, d[r]);    // one shared-memory atomic per distinct digit of the warp
    if (d[r] < 256u && lane == __ffs(peers) - 1) atomicAdd(&cnt[d[r]], __popc(peers));
  }
  __syncthreads();
  if (threadIdx.x < 256) {
    const int v = cnt[threadIdx.x];
    if (v) {
      hist_tiles[tile * 256 + threadIdx.x] = v;
      atomicAdd(&hist_super[(tile / kSortSuper) * 256 + threadIdx.x], v);
    }
  }
}

__global__ void __launch_bounds__(kSortThreads, 2) k_sort_pass(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                            uint64_t *__restrict__ kout, uint32_t *__restrict__ vout, int n, int shift,
                                                            const int *__restrict__ hist_tiles, const int *__restrict__ hist_super,
                                                            int n_super, int *__restrict__ next_tiles, int *__restrict__ next_super,
                                                            int next_shift /* < 0: last pass */) {
  __shared__ uint16_t cnt[256 * kSortWarps];          // [digit][warp] -> exclusive prefix in (digit, warp) order
  __shared__ int s_base[256];                          // output position of the tile's first element with digit d
  __shared__ int s_warp[kSortWarps];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;

  // ---- 1. output bases from the count matrices (no dependence on other CTAs of this launch) ----
  int before = 0, total = 0;
  if (tid < 256) {
    const int my_super = tile / kSortSuper;
    int s = 0;
    for (; s + 8 <= n_super; s += 8) {                 // eight independent loads in flight
      int v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldg(hist_super + (s + j) * 256 + tid);
#pragma unroll
      for (int j = 0; j < 8; ++j) { total += v[j]; if (s + j < my_super) before += v[j]; }
    }
    for (; s < n_super; ++s) {
      const int v = __ldg(hist_super + s * 256 + tid);
      total += v;
      if (s < my_super) before += v;
    }
    int v[kSortSuper - 1];
#pragma unroll
    for (int j = 0; j < kSortSuper - 1; ++j) {
      const int t = my_super * kSortSuper + j;
      v[j] = t < tile ? __ldg(hist_tiles + t * 256 + tid) : 0;
    }
#pragma unroll
    for (int j = 0; j < kSortSuper - 1; ++j) before += v[j];
  }
  for (int i = tid; i < 256 * kSortWarps; i += kSortThreads) cnt[i] = 0;
  {
    int incl = total;                                  // exclusive scan of `total` over the 256 digits (threads 0..255 = 8 warps)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp && w < 8; ++w) wbase += s_warp[w];
    if (tid < 256) s_base[tid] = wbase + incl - total + before;
  }
  __syncthreads();

  // ---- 1b. a digit that is the same for ALL n keys (high bits of small coordinates / batch indices): the pass is the identity
  //          permutation - copy the tile, count its next digits, done (a fraction of the cost of a real pass) ----
  if (__syncthreads_or(tid < 256 && total == n)) {
    int *hcnt = (int *)cnt;                               // [256] next-digit counts of this tile
    if (tid < 256) hcnt[tid] = 0;
    __syncthreads();
    const int base = tile * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
      const int e = base + r * kSortThreads + tid;
      const bool ok = e < n;
      uint64_t k = 0;
      if (ok) {
        k = kin[e];
        kout[e] = k;
        vout[e] = vin[e];
      }
      if (next_shift >= 0) {
        const uint32_t d2 = ok ? (uint32_t)((k >> next_shift) & 255ull) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d2);
        if (ok && lane == __ffs(peers) - 1) atomicAdd(&hcnt[d2], __popc(peers));
      }
    }
    __syncthreads();
    if (next_shift >= 0 && tid < 256 && hcnt[tid]) {
      next_tiles[tile * 256 + tid] = hcnt[tid];
      atomicAdd(next_super + (tile / kSortSuper) * 256 + tid, hcnt[tid]);
    }
    return;
  }

  // ---- 2. stable rank inside the tile: warp w owns elements [w * 256, (w + 1) * 256), visited in order ----
  uint64_t key[kSortItems];
  uint32_t val[kSortItems];
  uint16_t rank[kSortItems];
  const int e0 = tile * kSortTile + warp * (kSortTile / kSortWarps) + lane;
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int e = e0 + r * 32;
    const bool ok = e < n;
    key[r] = ok ? kin[e] : ~0ull;
    val[r] = ok ? vin[e] : 0u;
  }
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const bool ok = e0 + r * 32 < n;
    const uint32_t d = ok ? (uint32_t)((key[r] >> shift) & 255ull) : 256u;       // 256: beyond the array, never counted
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint16_t old = 0;
    if (ok && lane == leader) {
      old = cnt[d * kSortWarps + warp];
      cnt[d * kSortWarps + warp] = (uint16_t)(old + __popc(peers));
    }
    old = (uint16_t)__shfl_sync(0xffffffffu, (int)old, leader);
    rank[r] = (uint16_t)(old + __popc(peers & ((1u << lane) - 1u)));
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan of the 4096 counters in (digit, warp) order: 8 per thread
  {
    int v[8], sum = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) { v[e] = cnt[tid * 8 + e]; sum += v[e]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += s_warp[w];
    int run = wbase + incl - sum;
#pragma unroll
    for (int e = 0; e < 8; ++e) { cnt[tid * 8 + e] = (uint16_t)run; run += v[e]; }
  }
  __syncthreads();

  // ---- 3. stage the tile in digit order, then write it out in runs (+ the next pass's digit counts at the destination) ----
  extern __shared__ __align__(16) uint8_t sort_smem[];
  uint64_t *skey = (uint64_t *)sort_smem;                                  // [kSortTile]
  uint32_t *sval = (uint32_t *)(skey + kSortTile);                         // [kSortTile]
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    if (e0 + r * 32 >= n) continue;
    const uint32_t d = (uint32_t)((key[r] >> shift) & 255ull);
    const int li = (int)cnt[d * kSortWarps + warp] + (int)rank[r];         // position inside the digit-sorted tile
    skey[li] = key[r];
    sval[li] = val[r];
  }
  __syncthreads();
  const int count = min(kSortTile, n - tile * kSortTile);
  for (int i0 = warp * 32; i0 < count; i0 += kSortThreads) {               // warp-uniform bound
    const int i = i0 + lane;
    const bool ok = i < count;
    uint64_t k = 0;
    int pos = 0;
    if (ok) {
      k = skey[i];
      const uint32_t d = (uint32_t)((k >> shift) & 255ull);
      pos = s_base[d] + i - (int)cnt[d * kSortWarps];
      kout[pos] = k;
      vout[pos] = sval[i];
    }
    if (next_shift >= 0) {
      const int t2 = pos / kSortTile;
      const uint32_t tag = ok ? (((uint32_t)t2 << 8) | (uint32_t)((k >> next_shift) & 255ull)) : 0xffffffffu;
      const uint32_t peers = __match_any_sync(0xffffffffu, tag);
      if (ok && lane == __ffs(peers) - 1) {
        const int c = __popc(peers), d2 = (int)(tag & 255u);
        atomicAdd(next_tiles + t2 * 256 + d2, c);
        atomicAdd(next_super + (t2 / kSortSuper) * 256 + d2, c);
      }
    }
  }
}

constexpr size_t kSortSmemBytes = (size_t)kSortTile * 12;

}  // namespace

size_t sort_scratch_ints(int64_t n, int end_bit) {
  const int64_t tiles = div_up(n, kSortTile), supers = div_up(tiles, kSortSuper);
  return (size_t)((end_bit + 7) / 8) * (size_t)(tiles + supers) * 256;
}

// Sorts n pairs by key bits [0, end_bit), stable.  *kin/*vin hold the input, *kout/*vout are same-size buffers; on return
// *kout/*vout point at the sorted pairs (the pointers are swapped when the number of passes is even).  `hist` must hold
// sort_scratch_ints(n, end_bit) ints.
int sort_pairs(egn_ctx *ctx, uint64_t **kin, uint64_t **kout, uint32_t **vin, uint32_t **vout, int n, int end_bit, int *hist,
               cudaStream_t s) {
  const int passes = (end_bit + 7) / 8;
  const int tiles = (int)div_up(n, kSortTile), supers = (int)div_up(tiles, kSortSuper);
  const size_t per_pass = (size_t)(tiles + supers) * 256;
  EGN_SMEM_OPTIN(ctx, k_sort_pass, kSortSmemBytes);
  EGN_CUDA(cudaMemsetAsync(hist, 0, per_pass * passes * sizeof(int), s));
  ctx->prof.launches += 1;
  k_sort_count<<<tiles, kSortThreads, 0, s>>>(*kin, n, 0, hist, hist + (size_t)tiles * 256);
  uint64_t *ka = *kin, *kb = *kout;
  uint32_t *va = *vin, *vb = *vout;
  static const bool count_kernel = getenv("EGN_SORT_COUNT") && getenv("EGN_SORT_COUNT")[0] == '1';   // experiment: counting kernel per pass instead of atomics
  for (int p = 0; p < passes; ++p) {
    int *ht = hist + per_pass * p, *hs = ht + (size_t)tiles * 256;
    int *nt = p + 1 < passes ? hist + per_pass * (p + 1) : nullptr, *ns = nt ? nt + (size_t)tiles * 256 : nullptr;
    ctx->prof.launches += 1;
    if (count_kernel) {
      if (p > 0) { ctx->prof.launches += 1; k_sort_count<<<tiles, kSortThreads, 0, s>>>(ka, n, 8 * p, ht, hs); }
      nt = ns = nullptr;
    }
    k_sort_pass<<<tiles, kSortThreads, kSortSmemBytes, s>>>(ka, va, kb, vb, n, 8 * p, ht, hs, supers, nt, ns, (nt && p + 1 < passes) ? 8 * (p + 1) : -1);
    uint64_t *tk = ka; ka = kb; kb = tk;
    uint32_t *tv = va; va = vb; vb = tv;
  }
  EGN_CUDA(cudaGetLastError());
  // the sorted pairs are in (ka, va) after the last swap
  *kout = ka; *vout = va; *kin = kb; *vin = vb;
  return EGN_OK;
}

}  // namespace egn
