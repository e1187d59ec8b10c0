// Coordinate pyramid ("coordinate manager") of the engine: voxel quantisation, Morton keys, one radix sort,
// all strided levels in one pass, parent/child links and 27-neighbour tables without any hash table.
//
// Replaces (reference file:line under /root/reference):
//   ME.utils.sparse_quantize           datasets/quantization.py:42,83      -> quantize()
//   ME.SparseTensor(coords) hash build models/minkgl.py:269                -> coords_build() level 0
//   stride-2 coordinate maps           models/minkgl.py:104-105,145-146    -> coords_build() levels 1..
//   3^3 / 2^3 kernel maps              layers/eca_block.py:59-64, models/minkgl.py:39,145 -> nbr tables, cstart/cmask
//
// Design: keys are (batch | Morton(z,y,x)); after ONE sort every coarser level is a run-length compaction of
// level 0 (parent key = key >> 3), so level L row indices are prefix counts of "new run at level L" flags and
// the child->parent map, the first-child pointer and the 8-bit child occupancy fall out of the same pass.
// Neighbour tables are built top-down: the neighbours of a voxel are children of its parent's neighbours.

#include "ctx.cuh"

namespace egn {

constexpr int kTileThreads = 256;
constexpr int kTileSteps = 8;                                  // 32-element steps per warp
constexpr int kTile = kTileThreads * kTileSteps;               // 2048 keys per block
constexpr int kWarpsPerTile = kTileThreads / 32;

// ------------------------------------------------------------------------------------------------------
// Narrow sort keys.  The radix sort is the one multi-pass step of the build and its cost is the number of 8-bit passes
// (CUB onesweep: ~15 us per pass at 0.75 M keys, look-back latency).  A level-0 key has 54 Morton bits because the
// coordinate bias is 2^17; real scans span a few thousand voxels.  When every |c| < 2^11 the key
//     compact = batch << 36 | morton36(c + 2^11)
// is ORDER-EQUIVALENT to the full key (adding 2^17 - 2^11 maps the top bit h of c + 2^11 to the constant patterns 0111111 /
// 1000000 in bits 11..17 and leaves the low 11 bits: the most significant differing bit of two keys moves from bit 11 to
// bit 17 on the same axes with the same sign), so sorting 36 + batch bits (5-6 passes instead of 8) gives the same
// permutation; k_expand_keys rebuilds the full keys afterwards.  The pack kernels flag (status bit 1) any coordinate
// outside the narrow range and the build is then redone with full-width keys.
// ------------------------------------------------------------------------------------------------------
constexpr int kNarrowBits = 12;                                  // per-axis bits of the compact key
constexpr int kNarrowBias = 1 << (kNarrowBits - 1);
__device__ __forceinline__ uint64_t make_narrow_key(uint32_t b, int cx, int cy, int cz) {
  return ((uint64_t)b << (3 * kNarrowBits)) | spread3((uint32_t)(cx + kNarrowBias)) | (spread3((uint32_t)(cy + kNarrowBias)) << 1) |
         (spread3((uint32_t)(cz + kNarrowBias)) << 2);
}
__device__ __forceinline__ bool narrow_ok(int cx, int cy, int cz) {
  return cx >= -kNarrowBias && cx < kNarrowBias && cy >= -kNarrowBias && cy < kNarrowBias && cz >= -kNarrowBias && cz < kNarrowBias;
}
__global__ void k_expand_keys(uint64_t *__restrict__ keys, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    const uint32_t b = (uint32_t)(k >> (3 * kNarrowBits));
    const uint64_t m = k & ((1ull << (3 * kNarrowBits)) - 1ull);
    const uint32_t d = (uint32_t)(kAxisBias - kNarrowBias);
    keys[i] = make_key(0, b, compact3(m) + d, compact3(m >> 1) + d, compact3(m >> 2) + d);
  }
}

// ------------------------------------------------------------------------------------------------------
// pack (b,x,y,z) -> key, validate
// ------------------------------------------------------------------------------------------------------
__global__ void k_pack_keys(const int4 *__restrict__ coords, int n, int narrow, uint64_t *__restrict__ keys,
                            uint32_t *__restrict__ vals, int *__restrict__ dev_counts /* [P]=n_batches, [P+1]=status */) {
  int bmax = 0, bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = coords[i];
    const bool ok = c.x >= 0 && c.x < kMaxBatch && c.y >= -kAxisBias && c.y < kAxisBias && c.z >= -kAxisBias &&
                    c.z < kAxisBias && c.w >= -kAxisBias && c.w < kAxisBias;
    if (!ok) {
      bad = 1;
      c = make_int4(0, 0, 0, 0);
    }
    bmax = max(bmax, c.x + 1);
    if (narrow) {
      if (!narrow_ok(c.y, c.z, c.w)) { bad |= 2; c.y = c.z = c.w = 0; }          // outside the narrow range: the build is redone
      keys[i] = make_narrow_key((uint32_t)c.x, c.y, c.z, c.w);
    } else {
      keys[i] = make_key(0, (uint32_t)c.x, (uint32_t)(c.y + kAxisBias), (uint32_t)(c.z + kAxisBias),
                         (uint32_t)(c.w + kAxisBias));
    }
    vals[i] = (uint32_t)i;
  }
  bmax = __reduce_max_sync(0xffffffffu, bmax);
  bad = __reduce_or_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    if (bmax) atomicMax(&dev_counts[P], bmax);
    if (bad) atomicOr(&dev_counts[P + 1], bad);
  }
}

// height of element i: the highest pyramid level at which it starts a new run (-1: duplicate of i-1)
__device__ __forceinline__ int key_height(const uint64_t *__restrict__ keys, int i, int n) {
  if (i >= n) return -1;
  if (i == 0) return P - 1;
  const uint64_t d = keys[i] ^ keys[i - 1];
  if (d == 0) return -1;
  const int h = (63 - __clzll((long long)d)) / 3;
  return h < P - 1 ? h : P - 1;
}

// counts[L * nblocks + blk] = number of level-L rows that start inside this tile
__global__ void __launch_bounds__(kTileThreads) k_level_count(const uint64_t *__restrict__ keys, int n, int nblocks,
                                                              int *__restrict__ counts) {
  __shared__ int s_cnt[P];
  if (threadIdx.x < P) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int base = blockIdx.x * kTile + warp * (32 * kTileSteps);
  int cnt[P];
#pragma unroll
  for (int L = 0; L < P; ++L) cnt[L] = 0;
#pragma unroll
  for (int s = 0; s < kTileSteps; ++s) {
    const int h = key_height(keys, base + s * 32 + lane, n);
#pragma unroll
    for (int L = 0; L < P; ++L) cnt[L] += __popc(__ballot_sync(0xffffffffu, h >= L));
  }
  if (lane == 0) {
#pragma unroll
    for (int L = 0; L < P; ++L) atomicAdd(&s_cnt[L], cnt[L]);
  }
  __syncthreads();
  if (threadIdx.x < P) counts[threadIdx.x * nblocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// exclusive scan of counts along blocks for every level (single block), totals -> dev_counts[0..P)
__global__ void __launch_bounds__(1024) k_level_scan(int *__restrict__ counts, int nblocks, int *__restrict__ dev_counts) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const int L = blockIdx.x;                               // one CTA per pyramid level
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    int *row = counts + (size_t)L * nblocks;
    for (int start = 0; start < nblocks; start += 1024) {
      const int i = start + threadIdx.x;
      const int v = i < nblocks ? row[i] : 0;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) s_warp[warp] = x;
      __syncthreads();
      if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += y;
        }
        s_warp[lane] = w;
      }
      __syncthreads();
      const int carry = s_carry;
      const int incl = x + (warp ? s_warp[warp - 1] : 0) + carry;
      if (i < nblocks) row[i] = incl - v;
      __syncthreads();
      if (threadIdx.x == 1023) s_carry = incl;
      __syncthreads();
    }
    if (threadIdx.x == 0) dev_counts[L] = s_carry;
    __syncthreads();
  }
}

struct LevelPtrs {
  uint64_t *keys[P];
  int *up[P];
  int *cstart[P];
  uint32_t *cmask[P];
  int *perm0;
  unsigned long long *mask64;
  int *first0;
};

__global__ void __launch_bounds__(kTileThreads) k_level_emit(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                             int n, int nblocks, const int *__restrict__ base, LevelPtrs out) {
  __shared__ int s_wc[kWarpsPerTile][P];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int first = blockIdx.x * kTile + warp * (32 * kTileSteps);
  const uint32_t lt = (1u << lane) - 1u;
  int h[kTileSteps];
  int run[P];
#pragma unroll
  for (int L = 0; L < P; ++L) run[L] = 0;
#pragma unroll
  for (int s = 0; s < kTileSteps; ++s) {
    h[s] = key_height(keys, first + s * 32 + lane, n);
#pragma unroll
    for (int L = 0; L < P; ++L) run[L] += __popc(__ballot_sync(0xffffffffu, h[s] >= L));
  }
  if (lane == 0) {
#pragma unroll
    for (int L = 0; L < P; ++L) s_wc[warp][L] = run[L];
  }
  __syncthreads();
#pragma unroll
  for (int L = 0; L < P; ++L) {
    int b = base[(size_t)L * nblocks + blockIdx.x];
    for (int w = 0; w < warp; ++w) b += s_wc[w][L];
    run[L] = b;
  }
#pragma unroll
  for (int s = 0; s < kTileSteps; ++s) {
    const int i = first + s * 32 + lane;
    const int hh = h[s];
    int e[P], c[P];
#pragma unroll
    for (int L = 0; L < P; ++L) {
      const uint32_t bal = __ballot_sync(0xffffffffu, hh >= L);
      e[L] = run[L] + __popc(bal & lt);
      c[L] = e[L] + (hh >= L ? 1 : 0);
      run[L] += __popc(bal);
    }
    if (hh >= 0) {
      const uint64_t key = keys[i];
      out.perm0[e[0]] = (int)vals[i];
#pragma unroll
      for (int L = 0; L < P; ++L) {
        if (hh >= L) {
          out.keys[L][e[L]] = key >> (3 * L);
          if (L + 1 < P) out.up[L][e[L]] = c[L + 1] - 1;
          if (L >= 1) out.cstart[L][e[L]] = e[L - 1];
        }
        if (L >= 1 && hh >= L - 1) atomicOr(&out.cmask[L][c[L] - 1], 1u << (uint32_t)((key >> (3 * (L - 1))) & 7ull));
      }
      atomicOr(&out.mask64[c[2] - 1], 1ull << (key & 63ull));
      if (hh >= 2) out.first0[e[2]] = e[0];
    }
  }
}

// boff[L][b] = first row of batch b at level L (b = 0..n_batches)
struct BoffArgs {
  const uint64_t *keys[P];
  int *boff[P];
  int n[P];
};
// first row of every batch index at every level: one WARP per (level, batch), 32-ary search (4 probe rounds for a million
// rows instead of 20 dependent loads)
__global__ void k_batch_offsets(BoffArgs a, int n_batches) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int L = t / (n_batches + 1), b = t % (n_batches + 1);
  if (L >= P) return;
  const uint64_t target = (uint64_t)b << (kMortonBits - 3 * L);
  int lo = 0, hi = a.n[L];                                  // answer in [lo, hi]: first index with key >= target
  const uint64_t *k = a.keys[L];
  while (hi - lo > 0) {
    const int span = hi - lo;
    const int step = (span + 31) / 32;                      // lane probes lo + lane * step
    const int i = lo + lane * step;
    const bool below = i < hi && k[i] < target;             // keys sorted: `below` is true for a prefix of the lanes
    const uint32_t m = __ballot_sync(0xffffffffu, below);
    const int c = __popc(m);                                // probes [0, c) are below the target
    if (c == 0) { hi = lo; break; }
    const int nlo = lo + (c - 1) * step + 1;                // the last probe below the target is excluded ...
    const int nhi = min(hi, lo + c * step);                 // ... the first probe not below it (or hi) bounds the answer
    lo = nlo; hi = nhi;
  }
  if (lane == 0) a.boff[L][b] = lo;
}

// 27-neighbour table of the coarsest level by binary search (a few hundred rows per cloud).  One warp per row: lane k < 27
// computes entry k, the warp's ballot is the row's 27-bit presence mask (bit k = neighbour k exists).
__global__ void k_nbr_top(const uint64_t *__restrict__ keys, int n, int level, int *__restrict__ nbr, uint32_t *__restrict__ mask27) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
  const int lim = 1 << (kAxisBits - level);
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    uint32_t b, vx, vy, vz;
    split_key(level, keys[r], b, vx, vy, vz);
    const int nx = (int)vx + dx, ny = (int)vy + dy, nz = (int)vz + dz;
    int res = -1;
    if (lane < 27 && nx >= 0 && nx < lim && ny >= 0 && ny < lim && nz >= 0 && nz < lim) {
      const uint64_t q = make_key(level, b, (uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
      int lo = 0, hi = n;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] < q) lo = mid + 1; else hi = mid;
      }
      if (lo < n && keys[lo] == q) res = lo;
    }
    if (lane < 27) nbr[(int64_t)r * 27 + lane] = res;
    const uint32_t m = __ballot_sync(0xffffffffu, res >= 0);
    if (lane == 0) mask27[r] = m;
  }
}

// level L from level L+1: neighbour (c + d) lives in the parent's neighbour (or the parent itself) as the
// child with code c' ; child row = cstart + popc(mask below c').  One warp per row (lane = offset k), 108-byte coalesced
// table rows, presence mask by ballot.  Every entry is a chain of three dependent loads (row -> parent -> parent's
// neighbour -> its child mask), so a warp walks kNbrRows rows at once: their chains overlap instead of queueing.
constexpr int kNbrRows = 8;
__global__ void __launch_bounds__(256) k_nbr_down(const uint64_t *__restrict__ keys, const int *__restrict__ up, int n,
                                                  const int *__restrict__ nbr_up, const int *__restrict__ cstart_up,
                                                  const uint32_t *__restrict__ cmask_up, int *__restrict__ nbr, uint32_t *__restrict__ mask27) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
  for (int r0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kNbrRows; r0 < n; r0 += warps * kNbrRows) {
    uint32_t code[kNbrRows], cc[kNbrRows], m[kNbrRows];
    int p[kNbrRows], q[kNbrRows], cs[kNbrRows], kk[kNbrRows];
#pragma unroll
    for (int j = 0; j < kNbrRows; ++j) {
      const int r = min(r0 + j, n - 1);
      code[j] = (uint32_t)(keys[r] & 7ull);
      p[j] = up[r];
    }
#pragma unroll
    for (int j = 0; j < kNbrRows; ++j) {
      const int px = (int)(code[j] & 1u) + dx, py = (int)((code[j] >> 1) & 1u) + dy, pz = (int)((code[j] >> 2) & 1u) + dz;
      kk[j] = ((px >> 1) + 1) + 3 * ((py >> 1) + 1) + 9 * ((pz >> 1) + 1);
      cc[j] = (uint32_t)(px & 1) | ((uint32_t)(py & 1) << 1) | ((uint32_t)(pz & 1) << 2);
      q[j] = (lane < 27 && kk[j] != 13) ? nbr_up[(int64_t)p[j] * 27 + kk[j]] : p[j];
    }
#pragma unroll
    for (int j = 0; j < kNbrRows; ++j) {
      const int qq = q[j] >= 0 ? q[j] : 0;
      m[j] = cmask_up[qq];
      cs[j] = cstart_up[qq];
    }
#pragma unroll
    for (int j = 0; j < kNbrRows; ++j) {
      const int r = r0 + j;
      int res = -1;
      if (lane < 27 && q[j] >= 0 && ((m[j] >> cc[j]) & 1u)) res = cs[j] + __popc(m[j] & ((1u << cc[j]) - 1u));
      const uint32_t mm = __ballot_sync(0xffffffffu, res >= 0);
      if (r < n) {
        if (lane < 27) nbr[(int64_t)r * 27 + lane] = res;
        if (lane == 0) mask27[r] = mm;
      }
    }
  }
}

// ---- tile row orders ---------------------------------------------------------------------------------------------------
// The tensor-core convolutions work on tiles of 128 output rows and skip a (tile, 64-element reduction chunk) pair only
// when NO row of the tile has the chunk's neighbour(s).  In canonical (Morton) order a tile of a LiDAR level fills its
// chunks to ~30 %: rows on differently oriented surfaces share a tile and the tile pays for the union of their offsets.
// So every level also gets ROW ORDERS for tiling - permutations of its rows, computed here once per coords_build and
// shared by all convolutions of the level - in which rows with similar offset sets are adjacent:
//   kind 0 (3x3x3):            rows sorted by their 27-bit presence mask with the bits ranked rare -> significant
//                              (corner offsets, then edges, then faces; vertical before horizontal) so that the rows
//                              owning a rare offset are isolated in few tiles;
//   kind 1 (2x2x2 stride 2):   output (parent) rows sorted by their 8-bit child mask;
//   kind 2 (transposed 2x2x2): output (fine) rows sorted by their own child code (each row uses exactly one kernel slice).
// Rows are only re-grouped inside windows of W (8192 by default) canonical rows - one CTA sorts one window in shared
// memory (stable radix sort, deterministic) - so a tile still gathers from a compact neighbourhood.  The maps themselves stay in canonical order: only the tile -> row assignment changes.
constexpr int kOrderThreads = 1024;
constexpr int kMaxOrderJobs = 3 * P;

struct OrderJob {
  const void *src;      // kind 0: uint32 mask27[n]; kind 1: uint32 cmask[n]; kind 2: uint64 keys[n]
  int *out;             // order[n]
  int n, kind, first_block;
};
struct OrderJobs {
  OrderJob job[kMaxOrderJobs];
  int n_jobs;
};

// sort-key position of offset k (13 = centre: always present, not part of the key); most common offsets -> low bits
__constant__ int8_t c_order_bitpos[27] = {18, 10, 19, 11, 4, 12, 20, 13, 21, 6, 0, 7, 1, -1, 2, 8, 3, 9, 22, 14, 23, 15, 5, 16, 24, 17, 25};

// One CTA sorts one window of W rows by (key, row) in shared memory: LSD radix sort with 8-bit digits (4 passes for the
// 26-bit 3x3x3 keys, 1 pass for the 8-bit child masks and 3-bit child codes).  A warp owns W/32 consecutive elements and
// ranks them in order with __match_any_sync (stable), per-(digit, warp) counters are scanned digit-major by the block.
template <int W>
__global__ void __launch_bounds__(kOrderThreads) k_order_windows(OrderJobs jobs) {
  constexpr int ROUNDS = W / kOrderThreads;
  extern __shared__ __align__(16) uint8_t order_smem[];
  uint32_t *key_a = (uint32_t *)order_smem, *key_b = key_a + W;
  uint16_t *idx_a = (uint16_t *)(key_b + W), *idx_b = idx_a + W;
  uint16_t *cnt = idx_b + W;                                   // [256 digits][32 warps]
  __shared__ int warp_sums[32];
  int j = 0;
  while (j + 1 < jobs.n_jobs && (int)blockIdx.x >= jobs.job[j + 1].first_block) ++j;
  const OrderJob &jb = jobs.job[j];
  const int base = ((int)blockIdx.x - jb.first_block) * W;
  const int count = min(W, jb.n - base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < W; i += kOrderThreads) {
    uint32_t k32 = 0xffffffffu;
    if (i < count) {
      if (jb.kind == 0) {
        const uint32_t m = ((const uint32_t *)jb.src)[base + i];
        k32 = 0;
#pragma unroll
        for (int k = 0; k < 27; ++k)
          if (k != 13) k32 |= ((m >> k) & 1u) << c_order_bitpos[k];
      } else if (jb.kind == 1) {
        k32 = ((const uint32_t *)jb.src)[base + i] & 0xffu;
      } else {
        k32 = (uint32_t)(((const uint64_t *)jb.src)[base + i] & 7ull);
      }
    }
    key_a[i] = k32;
    idx_a[i] = (uint16_t)i;
  }
  const int passes = jb.kind == 0 ? 4 : 1;
  for (int p = 0; p < passes; ++p) {
    for (int i = tid; i < 256 * 32; i += kOrderThreads) cnt[i] = 0;
    __syncthreads();
    // rank inside the warp's segment (elements warp*W/32 + round*32 + lane, visited in order)
    uint32_t key[ROUNDS];
    uint16_t idx[ROUNDS], rank[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const int e = warp * (W / 32) + r * 32 + lane;
      key[r] = key_a[e];
      idx[r] = idx_a[e];
      const uint32_t d = (key[r] >> (8 * p)) & 255u;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const int leader = __ffs(peers) - 1;
      uint16_t old = 0;
      if (lane == leader) {
        old = cnt[d * 32 + warp];
        cnt[d * 32 + warp] = (uint16_t)(old + __popc(peers));
      }
      old = (uint16_t)__shfl_sync(0xffffffffu, (int)old, leader);
      rank[r] = (uint16_t)(old + __popc(peers & ((1u << lane) - 1u)));
      __syncwarp();
    }
    __syncthreads();
    // exclusive scan of the 8192 counters in (digit, warp) order: 8 per thread
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
      if (lane == 31) warp_sums[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        int w = warp_sums[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, wi, o);
          if (lane >= o) wi += t;
        }
        warp_sums[lane] = wi - w;
      }
      __syncthreads();
      int run = warp_sums[warp] + incl - sum;
#pragma unroll
      for (int e = 0; e < 8; ++e) { cnt[tid * 8 + e] = (uint16_t)run; run += v[e]; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const uint32_t d = (key[r] >> (8 * p)) & 255u;
      const int pos = cnt[d * 32 + warp] + rank[r];
      key_b[pos] = key[r];
      idx_b[pos] = idx[r];
    }
    __syncthreads();
    uint32_t *tk = key_a; key_a = key_b; key_b = tk;
    uint16_t *ti = idx_a; idx_a = idx_b; idx_b = ti;
  }
  for (int i = tid; i < count; i += kOrderThreads) jb.out[base + i] = base + (int)idx_a[i];
}

static size_t order_smem_bytes(int w) { return (size_t)w * 12 + 256 * 32 * 2; }

__global__ void k_decode_coords(const uint64_t *__restrict__ keys, int n, int level, int4 *__restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t b, vx, vy, vz;
    split_key(level, keys[i], b, vx, vy, vz);
    out[i] = make_int4((int)b, (int)(vx << level) - kAxisBias, (int)(vy << level) - kAxisBias,
                       (int)(vz << level) - kAxisBias);
  }
}

// profile mode only: number of present (out,in) pairs of a neighbour table / of the conv0 window
__global__ void k_count_pairs27(const int *__restrict__ nbr, int64_t total, unsigned long long *__restrict__ out) {
  unsigned long long c = 0;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) c += nbr[t] >= 0;
  c = __reduce_add_sync(0xffffffffu, (unsigned)c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
// conv0 pairs: for every level-2 cell, every voxel of the 5^3-dilated... counted exactly by brute force on bits:
// pairs = sum over ordered voxel pairs (a,b) with |a-b|_inf <= R; one warp per level-0 row, same walk as k_conv0.
__global__ void k_count_pairs_conv0(const uint64_t *__restrict__ keys0, const int *__restrict__ up0, const int *__restrict__ up1,
                                    const int *__restrict__ nbr2, const uint64_t *__restrict__ mask64, int n0, int KS,
                                    unsigned long long *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int KV = KS * KS * KS, R = KS / 2;
  unsigned long long total = 0;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n0; r += warps) {
    const uint32_t m = (uint32_t)(keys0[r] & 63ull);
    const int lx = (m & 1) | ((m >> 2) & 2), ly = ((m >> 1) & 1) | ((m >> 3) & 2), lz = ((m >> 2) & 1) | ((m >> 4) & 2);
    const int cell = up1[up0[r]];
    unsigned long long occ = 0ull;
    if (lane < 27) {
      const int q = nbr2[(int64_t)cell * 27 + lane];
      if (q >= 0) occ = mask64[q];
    }
    const uint32_t occ_lo = (uint32_t)occ, occ_hi = (uint32_t)(occ >> 32);
    for (int t0 = 0; t0 < KV; t0 += 32) {
      const int t = t0 + lane, tt = t < KV ? t : 0;
      const int px = lx + (tt % KS) - R, py = ly + (tt / KS) % KS - R, pz = lz + tt / (KS * KS) - R;
      const int j = ((px >> 2) + 1) + 3 * ((py >> 2) + 1) + 9 * ((pz >> 2) + 1);
      const uint32_t bit = (uint32_t)((px & 1) | ((py & 1) << 1) | ((pz & 1) << 2) | ((px & 2) << 2) | ((py & 2) << 3) | ((pz & 2) << 4));
      const uint32_t lo = __shfl_sync(0xffffffffu, occ_lo, j), hi = __shfl_sync(0xffffffffu, occ_hi, j);
      const unsigned long long o = ((unsigned long long)hi << 32) | lo;
      total += __popc(__ballot_sync(0xffffffffu, t < KV && ((o >> bit) & 1ull)));
    }
  }
  if (lane == 0 && total) atomicAdd(out, total);
}

// ------------------------------------------------------------------------------------------------------
// sort.cu: the in-tree stable LSD radix sort (one launch per 8-bit digit)
size_t sort_scratch_ints(int64_t n, int end_bit);
int sort_pairs(egn_ctx *ctx, uint64_t **kin, uint64_t **kout, uint32_t **vin, uint32_t **vout, int n, int end_bit, int *hist, cudaStream_t s);

static size_t sort_scratch_bytes(int64_t n) {
  // keys in/out, vals in/out, the sort's count matrices, tile counts, profile counters
  const int64_t nblocks = div_up(n, kTile);
  return pad256(n * 8) * 2 + pad256(n * 4) * 2 + pad256(sort_scratch_ints(n, 64) * 4) + pad256(nblocks * P * 4) + (1 << 16);
}

__global__ void k_quant_pack_batch(const float *__restrict__ pts, int n, const int *__restrict__ cloud_off, int n_clouds, float q0,
                                   float q1, float q2, int polar, int narrow, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                   int *__restrict__ dev_counts);

struct BuildSource {
  const int32_t *coords = nullptr;   // (n,4) voxel coordinates, or
  const float *points = nullptr;     // (n,3) raw points of n_clouds concatenated clouds
  const int *cloud_off = nullptr;
  int n_clouds = 0;
  float step[3] = {1.f, 1.f, 1.f};
  int polar = 0;
};

static int coords_build_common(egn_ctx *ctx, const BuildSource &src, int64_t n64, egn_coords_info *info, cudaStream_t s, bool narrow = true);

int coords_build(egn_ctx *ctx, const int32_t *coords, int64_t n64, egn_coords_info *info, cudaStream_t s) {
  EGN_CHECK(ctx && coords && info, EGN_ERR_INVALID, "coords_build: null argument");
  EGN_CHECK(((uintptr_t)coords & 15) == 0, EGN_ERR_INVALID, "coords_build: coords must be 16-byte aligned");
  BuildSource src;
  src.coords = coords;
  return coords_build_common(ctx, src, n64, info, s);
}

int coords_build_points(egn_ctx *ctx, const float *points, int64_t n64, const int32_t *cloud_offsets, int n_clouds,
                        const float step[3], int polar, egn_coords_info *info, cudaStream_t s) {
  EGN_CHECK(ctx && points && cloud_offsets && step && info, EGN_ERR_INVALID, "coords_build_points: null argument");
  EGN_CHECK(n_clouds >= 1 && n_clouds < kMaxBatch, EGN_ERR_RANGE, "coords_build_points: %d clouds (1..%d)", n_clouds, kMaxBatch - 1);
  EGN_CHECK(step[0] > 0.f && (!polar || (step[1] > 0.f && step[2] > 0.f)), EGN_ERR_INVALID, "coords_build_points: step must be > 0");
  BuildSource src;
  src.points = points; src.cloud_off = cloud_offsets; src.n_clouds = n_clouds; src.polar = polar;
  src.step[0] = step[0]; src.step[1] = step[1]; src.step[2] = step[2];
  return coords_build_common(ctx, src, n64, info, s);
}

static int coords_build_common(egn_ctx *ctx, const BuildSource &src, int64_t n64, egn_coords_info *info, cudaStream_t s, bool narrow) {
  EGN_CHECK(n64 > 0 && n64 < (int64_t)1 << 26, EGN_ERR_INVALID, "coords_build: n=%lld out of range (1..2^26)", (long long)n64);
  const int n = (int)n64;
  Pyramid &py = ctx->pyr;
  py = Pyramid();
  const int nblocks = (int)div_up(n, kTile);

  Arena &sc = ctx->scratch;
  EGN_TRY(sc.reserve(sort_scratch_bytes(n), s));
  uint64_t *kin = (uint64_t *)sc.take((size_t)n * 8), *kout = (uint64_t *)sc.take((size_t)n * 8);
  uint32_t *vin = (uint32_t *)sc.take((size_t)n * 4), *vout = (uint32_t *)sc.take((size_t)n * 4);
  int *counts = (int *)sc.take((size_t)nblocks * P * 4);
  EGN_CHECK(kin && kout && vin && vout && counts, EGN_ERR_STATE, "scratch arena exhausted");

  EGN_CUDA(cudaMemsetAsync(ctx->dev_counts, 0, sizeof(HostCounts), s));
  if (src.coords)
    EGN_LAUNCH(ctx, "coords_pack_keys", (double)n * 28, 0, s,
               k_pack_keys<<<grid_for(n, 256), 256, 0, s>>>((const int4 *)src.coords, n, narrow ? 1 : 0, kin, vin, ctx->dev_counts));
  else
    EGN_LAUNCH(ctx, "quantize_pack_batch", (double)n * 24, 0, s,
               k_quant_pack_batch<<<grid_for(n, 256), 256, 0, s>>>(src.points, n, src.cloud_off, src.n_clouds, src.step[0], src.step[1],
                                                                  src.step[2], src.polar, narrow ? 1 : 0, kin, vin, ctx->dev_counts));
  // sorted bits: full key = 64; narrow key = 36 Morton bits + the batch bits (known on the host for the point path)
  int batch_bits = 10;
  if (src.points) { batch_bits = 1; while ((1 << batch_bits) < src.n_clouds) ++batch_bits; }
  const int end_bit = narrow ? 3 * kNarrowBits + batch_bits : 64;
  int *sort_hist = (int *)sc.take(sort_scratch_ints(n, end_bit) * 4);
  EGN_CHECK(sort_hist != nullptr, EGN_ERR_STATE, "scratch arena exhausted (sort)");
  if (ctx->prof.on) ctx->prof.begin("coords_radix_sort", (double)n * 24 * ((end_bit + 7) / 8), 0, s);
  EGN_TRY(sort_pairs(ctx, &kin, &kout, &vin, &vout, n, end_bit, sort_hist, s));      // kout / vout: the sorted pairs
  if (ctx->prof.on) ctx->prof.end(s);
  if (narrow) EGN_LAUNCH(ctx, "coords_expand_keys", (double)n * 16, 0, s, k_expand_keys<<<grid_for(n, 256), 256, 0, s>>>(kout, n));
  EGN_LAUNCH(ctx, "coords_level_count", (double)n * 8, 0, s, k_level_count<<<nblocks, kTileThreads, 0, s>>>(kout, n, nblocks, counts));
  EGN_LAUNCH(ctx, "coords_level_scan", (double)nblocks * P * 8, 0, s, k_level_scan<<<P, 1024, 0, s>>>(counts, nblocks, ctx->dev_counts));
  EGN_CUDA(cudaMemcpyAsync(ctx->host, ctx->dev_counts, sizeof(HostCounts), cudaMemcpyDeviceToHost, s));
  EGN_CUDA(cudaStreamSynchronize(s));

  const HostCounts &hc = *ctx->host;
  if (narrow && (hc.status & 2) && !(hc.status & 1))          // some coordinate outside the narrow-key range: full-width keys
    return coords_build_common(ctx, src, n64, info, s, false);
  info->n_input = n;
  info->n_batches = hc.n_batches;
  info->status = (hc.status & 1) ? EGN_ERR_RANGE : ((hc.status & 4) ? EGN_ERR_INVALID : EGN_OK);
  for (int L = 0; L < P; ++L) info->n_rows[L] = hc.totals[L];
  EGN_CHECK((hc.status & 4) == 0, EGN_ERR_INVALID,
            "coords_build_points: cloud_offsets must start at 0, end at n and never decrease (they label every point with its cloud)");
  EGN_CHECK((hc.status & 1) == 0, EGN_ERR_RANGE,
            "coords_build: coordinate outside [-2^17, 2^17) (or NaN point) or batch index outside [0, 1023)");

  // exact-size pyramid storage
  const int B = hc.n_batches;
  size_t need = 0;
  for (int L = 0; L < P; ++L) {
    const size_t m = (size_t)hc.totals[L];
    need += pad256(m * 8) + 7 * pad256(m * 4) + pad256(m * 27 * 4) + pad256((size_t)(B + 1) * 4);
  }
  need += pad256((size_t)hc.totals[0] * 4) + pad256((size_t)hc.totals[2] * 8) + pad256((size_t)hc.totals[2] * 4) + 4096;
  Arena &ca = ctx->coords;
  EGN_TRY(ca.reserve(need, s));
  LevelPtrs lp;
  for (int L = 0; L < P; ++L) {
    const size_t m = (size_t)hc.totals[L];
    py.n[L] = hc.totals[L];
    py.keys[L] = lp.keys[L] = (uint64_t *)ca.take(m * 8);
    py.up[L] = lp.up[L] = (int *)ca.take(m * 4);
    py.cstart[L] = lp.cstart[L] = (int *)ca.take(m * 4);
    py.cmask[L] = lp.cmask[L] = (uint32_t *)ca.take(m * 4);
    py.nbr[L] = L >= 1 ? (int *)ca.take(m * 27 * 4) : nullptr;
    py.mask27[L] = L >= 1 ? (uint32_t *)ca.take(m * 4) : nullptr;
    py.ord27[L] = L >= 1 ? (int *)ca.take(m * 4) : nullptr;
    py.ordc[L] = L >= 1 ? (int *)ca.take(m * 4) : nullptr;
    py.ordt[L] = L + 1 < P ? (int *)ca.take(m * 4) : nullptr;
    py.boff[L] = (int *)ca.take((size_t)(B + 1) * 4);
    EGN_CHECK(py.keys[L] && py.up[L] && py.cstart[L] && py.cmask[L] && py.boff[L] && (L == 0 || (py.nbr[L] && py.mask27[L] && py.ord27[L] && py.ordc[L])) &&
                  (L + 1 >= P || py.ordt[L]),
              EGN_ERR_STATE, "coords arena exhausted");
    if (L >= 1) EGN_CUDA(cudaMemsetAsync(py.cmask[L], 0, m * 4, s));
  }
  py.perm0 = lp.perm0 = (int *)ca.take((size_t)hc.totals[0] * 4);
  py.mask64 = (uint64_t *)ca.take((size_t)hc.totals[2] * 8);
  lp.mask64 = (unsigned long long *)py.mask64;
  py.first0 = lp.first0 = (int *)ca.take((size_t)hc.totals[2] * 4);
  EGN_CHECK(py.perm0 && py.mask64 && py.first0, EGN_ERR_STATE, "coords arena exhausted");
  EGN_CUDA(cudaMemsetAsync(py.mask64, 0, (size_t)hc.totals[2] * 8, s));

  {
    double b = (double)n * 12;
    for (int L = 0; L < P; ++L) b += (double)py.n[L] * 20;
    EGN_LAUNCH(ctx, "coords_level_emit", b, 0, s, k_level_emit<<<nblocks, kTileThreads, 0, s>>>(kout, vout, n, nblocks, counts, lp));
  }

  BoffArgs ba;
  for (int L = 0; L < P; ++L) {
    ba.keys[L] = py.keys[L];
    ba.boff[L] = py.boff[L];
    ba.n[L] = py.n[L];
  }
  EGN_LAUNCH(ctx, "coords_batch_offsets", (double)P * (B + 1) * 4, 0, s,
             k_batch_offsets<<<(int)div_up((int64_t)P * (B + 1) * 32, 128), 128, 0, s>>>(ba, B));

  // kernel-map build: algorithmic bytes N*8 (keys) + N*27*4 (table) per level
  const int T = P - 1;
  EGN_LAUNCH(ctx, "kernel_map_3x3x3", (double)py.n[T] * (8 + 108), 0, s,
             k_nbr_top<<<grid_for((int64_t)py.n[T] * 32, 256), 256, 0, s>>>(py.keys[T], py.n[T], T, py.nbr[T], py.mask27[T]));
  for (int L = T - 1; L >= 1; --L)
    EGN_LAUNCH(ctx, "kernel_map_3x3x3", (double)py.n[L] * (8 + 108), 0, s,
               k_nbr_down<<<grid_for((int64_t)div_up(py.n[L], kNbrRows) * 32, 256, 16), 256, 0, s>>>(py.keys[L], py.up[L], py.n[L], py.nbr[L + 1],
                                                                                 py.cstart[L + 1], py.cmask[L + 1], py.nbr[L], py.mask27[L]));
  EGN_CUDA(cudaGetLastError());
  if (ctx->use_order) {   // tile row orders of every level: ONE launch, one CTA per window of kOrderWindow rows
    const int kOrderWindow = (ctx->order_window == 2048 || ctx->order_window == 4096) ? ctx->order_window : 8192;
    OrderJobs oj;
    oj.n_jobs = 0;
    int blocks = 0;
    double rows = 0;
    auto add = [&](const void *src, int *out, int n_rows, int kind) {
      if (n_rows <= 0) return;
      OrderJob &jb = oj.job[oj.n_jobs++];
      jb.src = src; jb.out = out; jb.n = n_rows; jb.kind = kind; jb.first_block = blocks;
      blocks += (int)div_up(n_rows, kOrderWindow);
      rows += n_rows;
    };
    for (int L = 1; L < P; ++L) add(py.mask27[L], py.ord27[L], py.n[L], 0);
    for (int L = 1; L < P; ++L) add(py.cmask[L], py.ordc[L], py.n[L], 1);
    for (int L = 0; L + 1 < P; ++L) add(py.keys[L], py.ordt[L], py.n[L], 2);
    if (kOrderWindow == 2048) {
      EGN_SMEM_OPTIN(ctx, k_order_windows<2048>, order_smem_bytes(2048));
      EGN_LAUNCH(ctx, "tile_row_orders", rows * 8, 0, s, k_order_windows<2048><<<blocks, kOrderThreads, order_smem_bytes(2048), s>>>(oj));
    } else if (kOrderWindow == 4096) {
      EGN_SMEM_OPTIN(ctx, k_order_windows<4096>, order_smem_bytes(4096));
      EGN_LAUNCH(ctx, "tile_row_orders", rows * 8, 0, s, k_order_windows<4096><<<blocks, kOrderThreads, order_smem_bytes(4096), s>>>(oj));
    } else {
      EGN_SMEM_OPTIN(ctx, k_order_windows<8192>, order_smem_bytes(8192));
      EGN_LAUNCH(ctx, "tile_row_orders", rows * 8, 0, s, k_order_windows<8192><<<blocks, kOrderThreads, order_smem_bytes(8192), s>>>(oj));
    }
    EGN_CUDA(cudaGetLastError());
    py.ordered = true;
  }
  py.n_input = n;
  py.n_batches = B;
  py.valid = true;

  if (ctx->prof.on) {  // pair counts for the algorithmic-byte model (not part of the product path)
    unsigned long long *cnt = (unsigned long long *)sc.take((P + 1) * 8);
    EGN_CHECK(cnt != nullptr, EGN_ERR_STATE, "scratch arena exhausted");
    EGN_CUDA(cudaMemsetAsync(cnt, 0, (P + 1) * 8, s));
    for (int L = 1; L < P; ++L)
      if (py.n[L]) k_count_pairs27<<<grid_for((int64_t)py.n[L] * 27, 256), 256, 0, s>>>(py.nbr[L], (int64_t)py.n[L] * 27, cnt + L);
    k_count_pairs_conv0<<<grid_for((int64_t)py.n[0] * 32, 256, 16), 256, 0, s>>>(py.keys[0], py.up[0], py.up[1], py.nbr[2], py.mask64,
                                                                                  py.n[0], 5, cnt + P);
    unsigned long long h[P + 1];
    EGN_CUDA(cudaMemcpyAsync(h, cnt, (P + 1) * 8, cudaMemcpyDeviceToHost, s));
    EGN_CUDA(cudaStreamSynchronize(s));
    for (int L = 1; L < P; ++L) py.pairs27[L] = (long long)h[L];
    py.pairs_conv0 = (long long)h[P];
  }
  return EGN_OK;
}

int coords_get(egn_ctx *ctx, int level, int32_t *out, cudaStream_t s) {
  EGN_CHECK(ctx && ctx->pyr.valid, EGN_ERR_STATE, "coords_get before coords_build");
  EGN_CHECK(level >= 0 && level < P && out, EGN_ERR_INVALID, "coords_get: bad level/out");
  const int n = ctx->pyr.n[level];
  if (n == 0) return EGN_OK;
  EGN_LAUNCH(ctx, "coords_decode", (double)n * 24, 0, s, k_decode_coords<<<grid_for(n, 256), 256, 0, s>>>(ctx->pyr.keys[level], n, level, (int4 *)out));
  EGN_CUDA(cudaGetLastError());
  return EGN_OK;
}

// ------------------------------------------------------------------------------------------------------
// sparse_quantize: floor(p / q) -> int32, first occurrence wins, survivors in input order
// ------------------------------------------------------------------------------------------------------
__global__ void k_quant_pack(const float *__restrict__ pts, int n, float q0, float q1, float q2, int polar,
                             int *__restrict__ vox /* (n,3) */, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                             int *__restrict__ dev_counts) {
  int bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
    float a, b, c;
    if (polar) {
      // datasets/quantization.py:35-41, f32 throughout: 180. + (atan2(y,x) * 180.) / np.pi, evaluated left to right
      const float theta = __fadd_rn(180.f, __fdiv_rn(__fmul_rn(atan2f(y, x), 180.f), 3.14159265358979323846f));
      const float dist = sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
      a = __fdiv_rn(theta, q0);
      b = __fdiv_rn(dist, q1);
      c = __fdiv_rn(z, q2);
    } else if (q0 != 1.0f) {
      a = __fdiv_rn(x, q0);
      b = __fdiv_rn(y, q0);
      c = __fdiv_rn(z, q0);
    } else {
      a = x; b = y; c = z;
    }
    const float fa = floorf(a), fb = floorf(b), fc = floorf(c);
    const float lim = (float)kAxisBias;
    const bool ok = fa >= -lim && fa < lim && fb >= -lim && fb < lim && fc >= -lim && fc < lim;  // also rejects NaN
    int ia = 0, ib = 0, ic = 0;
    if (ok) { ia = (int)fa; ib = (int)fb; ic = (int)fc; } else bad = 1;
    vox[3 * (size_t)i] = ia; vox[3 * (size_t)i + 1] = ib; vox[3 * (size_t)i + 2] = ic;
    keys[i] = make_key(0, 0u, (uint32_t)(ia + kAxisBias), (uint32_t)(ib + kAxisBias), (uint32_t)(ic + kAxisBias));
    vals[i] = (uint32_t)i;
  }
  bad = __reduce_max_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicOr(&dev_counts[P + 1], 1);
}

// fused path: raw points of B concatenated clouds -> level-0 keys (batch | Morton(voxel)) in one pass; the cloud of a
// point comes from a binary search in the (B+1) point offsets.  Same arithmetic as k_quant_pack.
__global__ void k_quant_pack_batch(const float *__restrict__ pts, int n, const int *__restrict__ cloud_off, int n_clouds, float q0,
                                   float q1, float q2, int polar, int narrow, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                   int *__restrict__ dev_counts) {
  int bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int lo = 0, hi = n_clouds;                       // last cloud c with cloud_off[c] <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(cloud_off + mid) <= i) lo = mid; else hi = mid;
    }
    float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
    float a, b, c;
    if (polar) {
      const float theta = __fadd_rn(180.f, __fdiv_rn(__fmul_rn(atan2f(y, x), 180.f), 3.14159265358979323846f));
      const float dist = sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
      a = __fdiv_rn(theta, q0); b = __fdiv_rn(dist, q1); c = __fdiv_rn(z, q2);
    } else if (q0 != 1.0f) {
      a = __fdiv_rn(x, q0); b = __fdiv_rn(y, q0); c = __fdiv_rn(z, q0);
    } else {
      a = x; b = y; c = z;
    }
    const float fa = floorf(a), fb = floorf(b), fc = floorf(c);
    const float lim = (float)kAxisBias;
    const bool ok = fa >= -lim && fa < lim && fb >= -lim && fb < lim && fc >= -lim && fc < lim;
    int ia = 0, ib = 0, ic = 0;
    if (ok) { ia = (int)fa; ib = (int)fb; ic = (int)fc; } else bad |= 1;
    if (narrow) {
      if (!narrow_ok(ia, ib, ic)) { bad |= 2; ia = ib = ic = 0; }
      keys[i] = make_narrow_key((uint32_t)lo, ia, ib, ic);
    } else {
      keys[i] = make_key(0, (uint32_t)lo, (uint32_t)(ia + kAxisBias), (uint32_t)(ib + kAxisBias), (uint32_t)(ic + kAxisBias));
    }
    vals[i] = (uint32_t)i;
  }
  // the first-point offsets must partition [0, n): start at 0, end at n, never decrease (checked once, by block 0)
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c <= n_clouds; c += blockDim.x) {
      const int o = __ldg(cloud_off + c);
      if ((c == 0 && o != 0) || (c == n_clouds && o != n) || (c > 0 && o < __ldg(cloud_off + c - 1))) bad |= 4;
    }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicOr(&dev_counts[P + 1], bad);
  if (blockIdx.x == 0 && threadIdx.x == 0) dev_counts[P] = n_clouds;
}

// stable sort => the first element of every run of equal keys is the earliest input row
__global__ void k_mark_first(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, int n,
                             uint8_t *__restrict__ keep) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (i == 0 || keys[i] != keys[i - 1]) keep[vals[i]] = 1;
}
__global__ void __launch_bounds__(kTileThreads) k_keep_count(const uint8_t *__restrict__ keep, int n, int *__restrict__ counts) {
  int c = 0;
  const int base = blockIdx.x * kTile;
  for (int j = threadIdx.x; j < kTile; j += kTileThreads) c += (base + j < n) ? keep[base + j] : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) counts[blockIdx.x] = s;
}
__global__ void __launch_bounds__(1024) k_scan1(int *__restrict__ counts, int nblocks, int *__restrict__ total) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int start = 0; start < nblocks; start += 1024) {
    const int i = start + threadIdx.x;
    const int v = i < nblocks ? counts[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int incl = x + (warp ? s_warp[warp - 1] : 0) + s_carry;
    if (i < nblocks) counts[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}
__global__ void __launch_bounds__(kTileThreads) k_keep_emit(const uint8_t *__restrict__ keep, const int *__restrict__ vox, int n,
                                                            const int *__restrict__ base, int *__restrict__ coords_out,
                                                            long long *__restrict__ index_out) {
  __shared__ int s_wc[kWarpsPerTile];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int first = blockIdx.x * kTile + warp * (32 * kTileSteps);
  int f[kTileSteps], cnt = 0;
#pragma unroll
  for (int s = 0; s < kTileSteps; ++s) {
    const int i = first + s * 32 + lane;
    f[s] = i < n ? keep[i] : 0;
    cnt += __popc(__ballot_sync(0xffffffffu, f[s]));
  }
  if (lane == 0) s_wc[warp] = cnt;
  __syncthreads();
  int run = base[blockIdx.x];
  for (int w = 0; w < warp; ++w) run += s_wc[w];
#pragma unroll
  for (int s = 0; s < kTileSteps; ++s) {
    const int i = first + s * 32 + lane;
    const uint32_t bal = __ballot_sync(0xffffffffu, f[s]);
    if (f[s]) {
      const int o = run + __popc(bal & ((1u << lane) - 1u));
      coords_out[3 * (size_t)o] = vox[3 * (size_t)i];
      coords_out[3 * (size_t)o + 1] = vox[3 * (size_t)i + 1];
      coords_out[3 * (size_t)o + 2] = vox[3 * (size_t)i + 2];
      index_out[o] = i;
    }
    run += __popc(bal);
  }
}

int quantize(egn_ctx *ctx, const float *points, int64_t n64, const float step[3], int polar, int32_t *coords_out,
             int64_t *index_out, int64_t *n_out, cudaStream_t s) {
  EGN_CHECK(ctx && points && step && coords_out && index_out && n_out, EGN_ERR_INVALID, "quantize: null argument");
  EGN_CHECK(n64 > 0 && n64 < (int64_t)1 << 26, EGN_ERR_INVALID, "quantize: n out of range");
  EGN_CHECK(step[0] > 0.f && (!polar || (step[1] > 0.f && step[2] > 0.f)), EGN_ERR_INVALID, "quantize: step must be > 0");
  const int n = (int)n64;
  const int nblocks = (int)div_up(n, kTile);
  Arena &sc = ctx->scratch;
  EGN_TRY(sc.reserve(sort_scratch_bytes(n) + pad256((size_t)n * 12) + pad256(n), s));
  uint64_t *kin = (uint64_t *)sc.take((size_t)n * 8), *kout = (uint64_t *)sc.take((size_t)n * 8);
  uint32_t *vin = (uint32_t *)sc.take((size_t)n * 4), *vout = (uint32_t *)sc.take((size_t)n * 4);
  int *vox = (int *)sc.take((size_t)n * 12);
  uint8_t *keep = (uint8_t *)sc.take((size_t)n);
  int *counts = (int *)sc.take((size_t)nblocks * 4);
  EGN_CHECK(kin && kout && vin && vout && vox && keep && counts, EGN_ERR_STATE, "scratch arena exhausted");
  EGN_CUDA(cudaMemsetAsync(ctx->dev_counts, 0, sizeof(HostCounts), s));
  EGN_CUDA(cudaMemsetAsync(keep, 0, (size_t)n, s));
  EGN_LAUNCH(ctx, "quantize_pack", (double)n * 36, 0, s,
             k_quant_pack<<<grid_for(n, 256), 256, 0, s>>>(points, n, step[0], step[1], step[2], polar, vox, kin, vin, ctx->dev_counts));
  int *sort_hist = (int *)sc.take(sort_scratch_ints(n, kMortonBits) * 4);
  EGN_CHECK(sort_hist != nullptr, EGN_ERR_STATE, "scratch arena exhausted (sort)");
  if (ctx->prof.on) ctx->prof.begin("quantize_radix_sort", (double)n * 24 * 7, 0, s);
  EGN_TRY(sort_pairs(ctx, &kin, &kout, &vin, &vout, n, kMortonBits, sort_hist, s));  // kout / vout: the sorted pairs
  if (ctx->prof.on) ctx->prof.end(s);
  EGN_LAUNCH(ctx, "quantize_mark_first", (double)n * 13, 0, s, k_mark_first<<<grid_for(n, 256), 256, 0, s>>>(kout, vout, n, keep));
  EGN_LAUNCH(ctx, "quantize_compact", (double)n * 2, 0, s, k_keep_count<<<nblocks, kTileThreads, 0, s>>>(keep, n, counts));
  EGN_LAUNCH(ctx, "quantize_compact", 0, 0, s, k_scan1<<<1, 1024, 0, s>>>(counts, nblocks, ctx->dev_counts + P + 2));
  EGN_LAUNCH(ctx, "quantize_compact", (double)n * 13, 0, s,
             k_keep_emit<<<nblocks, kTileThreads, 0, s>>>(keep, vox, n, counts, coords_out, (long long *)index_out));
  EGN_CUDA(cudaMemcpyAsync(ctx->host, ctx->dev_counts, sizeof(HostCounts), cudaMemcpyDeviceToHost, s));
  EGN_CUDA(cudaStreamSynchronize(s));
  EGN_CHECK(ctx->host->status == 0, EGN_ERR_RANGE, "quantize: voxel coordinate outside [-2^17, 2^17) (or NaN input)");
  *n_out = ctx->host->n_out;
  return EGN_OK;
}

// ------------------------------------------------------------------------------------------------------
// raw-scan ingest: PointCloudLoader.__call__ (misc/point_clouds.py:95-111) on the device - take the
// xyz of every (x, y, z[, reflectance]) record, drop the all-zero points (np.isclose(pc, 0): |v| <= 1e-8) and the
// points at or below the ground plane (z <= level), keep the input order.
// ------------------------------------------------------------------------------------------------------
__global__ void k_point_keep(const float *__restrict__ pts, int n, int stride, int remove_zero, int remove_ground, float ground,
                             uint8_t *__restrict__ keep) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = pts[(size_t)i * stride], y = pts[(size_t)i * stride + 1], z = pts[(size_t)i * stride + 2];
    bool k = true;
    if (remove_zero && fabsf(x) <= 1e-8f && fabsf(y) <= 1e-8f && fabsf(z) <= 1e-8f) k = false;
    if (remove_ground && !(z > ground)) k = false;
    keep[i] = k ? 1 : 0;
  }
}
__global__ void __launch_bounds__(kTileThreads) k_point_emit(const uint8_t *__restrict__ keep, const float *__restrict__ pts, int n, int stride,
                                                             const int *__restrict__ base, float *__restrict__ out) {
  __shared__ int s_wc[kWarpsPerTile];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int first = blockIdx.x * kTile + warp * (32 * kTileSteps);
  int f[kTileSteps], cnt = 0;
#pragma unroll
  for (int s = 0; s < kTileSteps; ++s) {
    const int i = first + s * 32 + lane;
    f[s] = i < n ? keep[i] : 0;
    cnt += __popc(__ballot_sync(0xffffffffu, f[s]));
  }
  if (lane == 0) s_wc[warp] = cnt;
  __syncthreads();
  int run = base[blockIdx.x];
  for (int w = 0; w < warp; ++w) run += s_wc[w];
#pragma unroll
  for (int s = 0; s < kTileSteps; ++s) {
    const int i = first + s * 32 + lane;
    const uint32_t bal = __ballot_sync(0xffffffffu, f[s]);
    if (f[s]) {
      const int o = run + __popc(bal & ((1u << lane) - 1u));
      out[3 * (size_t)o] = pts[(size_t)i * stride];
      out[3 * (size_t)o + 1] = pts[(size_t)i * stride + 1];
      out[3 * (size_t)o + 2] = pts[(size_t)i * stride + 2];
    }
    run += __popc(bal);
  }
}

int filter_points(egn_ctx *ctx, const float *records, int64_t n64, int stride, int remove_zero, int remove_ground, float ground_level,
                  float *points_out, int64_t *n_out, cudaStream_t s) {
  EGN_CHECK(ctx && records && points_out && n_out, EGN_ERR_INVALID, "filter_points: null argument");
  EGN_CHECK(n64 > 0 && n64 < (int64_t)1 << 28, EGN_ERR_INVALID, "filter_points: n out of range");
  EGN_CHECK(stride >= 3, EGN_ERR_INVALID, "filter_points: a record has at least 3 floats (x, y, z)");
  const int n = (int)n64;
  const int nblocks = (int)div_up(n, kTile);
  Arena &sc = ctx->scratch;
  EGN_TRY(sc.reserve(pad256((size_t)n) + pad256((size_t)nblocks * 4) + 4096, s));
  uint8_t *keep = (uint8_t *)sc.take((size_t)n);
  int *counts = (int *)sc.take((size_t)nblocks * 4);
  EGN_CHECK(keep && counts, EGN_ERR_STATE, "scratch arena exhausted");
  EGN_CUDA(cudaMemsetAsync(ctx->dev_counts, 0, sizeof(HostCounts), s));
  EGN_LAUNCH(ctx, "ingest_point_filter", (double)n * (stride * 4 + 1), 0, s,
             k_point_keep<<<grid_for(n, 256), 256, 0, s>>>(records, n, stride, remove_zero, remove_ground, ground_level, keep));
  EGN_LAUNCH(ctx, "ingest_compact", (double)n, 0, s, k_keep_count<<<nblocks, kTileThreads, 0, s>>>(keep, n, counts));
  EGN_LAUNCH(ctx, "ingest_compact", 0, 0, s, k_scan1<<<1, 1024, 0, s>>>(counts, nblocks, ctx->dev_counts + P + 2));
  EGN_LAUNCH(ctx, "ingest_compact", (double)n * 25, 0, s, k_point_emit<<<nblocks, kTileThreads, 0, s>>>(keep, records, n, stride, counts, points_out));
  EGN_CUDA(cudaMemcpyAsync(ctx->host, ctx->dev_counts, sizeof(HostCounts), cudaMemcpyDeviceToHost, s));
  EGN_CUDA(cudaStreamSynchronize(s));
  *n_out = ctx->host->n_out;
  return EGN_OK;
}

}  // namespace egn
