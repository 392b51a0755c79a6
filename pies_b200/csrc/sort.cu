// sort.cu — stable LSD radix sort and exclusive scan for the cell tables.
//
// The reference keeps cell occupancy in a phmap::parallel_flat_hash_map filled by
// 16 threads (reference Include/Pies/SpatialHash.h:129-189).  Here the (cell key, element)
// pairs are radix-sorted instead; because the sort is stable and pairs are emitted in
// ascending element order, every cell's member list comes out in ascending element index
// — exactly the bucket order the reference produces.
//
// One pass = histogram (per 2048-element tile, shared-memory counters), exclusive scan of
// the digit-major tile histogram, and a stable scatter that ranks equal digits inside a
// tile with warp match/ballot (no atomics on the ranking path => deterministic output).
// Sorts of at most kSortSmallTiles tiles skip the scan launches: every scatter CTA sums the
// tile-major histogram (L2 resident, at most 160 KB) into its own digit bases.
#include "kernels.h"
#include "common.cuh"

namespace pies {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048
constexpr uint32_t kSortSmallTiles = 160;               // up to 327 680 elements: two launches per pass instead of four
constexpr int kScanItems = 4;
constexpr int kScanTile = kThreads * kScanItems;      // 1024

// ------------------------------------------------------------------ scan -----
__global__ void __launch_bounds__(kThreads) k_scan_tile(uint32_t* __restrict__ data, uint64_t n,
                                                        uint32_t* __restrict__ tileSums) {
  __shared__ uint32_t warpTotals[kThreads / 32];
  uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems], sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { v[k] = base + k < n ? data[base + k] : 0u; sum += v[k]; }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warpTotals[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t t = lane < kThreads / 32 ? warpTotals[lane] : 0u, s = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += u; }
    if (lane < kThreads / 32) warpTotals[lane] = s - t;
    if (lane == kThreads / 32 - 1 && tileSums) tileSums[blockIdx.x] = s;
  }
  __syncthreads();
  uint32_t run = warpTotals[warp] + inc - sum;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { if (base + k < n) data[base + k] = run; run += v[k]; }
}

// Second (and last) kernel of a multi-tile scan: every CTA sums the totals of the tiles before its own (a few thousand
// L2-resident words at most) and adds that offset to its tile.  No third launch for a scan of the tile totals.
__global__ void __launch_bounds__(kThreads) k_scan_add(uint32_t* __restrict__ data, uint64_t n,
                                                       const uint32_t* __restrict__ tileSums) {
  __shared__ uint32_t warpTotals[kThreads / 32];
  __shared__ uint32_t sAdd;
  uint32_t acc = 0;
  for (uint32_t i = threadIdx.x; i < blockIdx.x; i += kThreads) acc += tileSums[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) warpTotals[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += warpTotals[w];
    sAdd = t;
  }
  __syncthreads();
  const uint32_t add = sAdd;
  if (!add) return;
  uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) if (base + k < n) data[base + k] += add;
}

size_t scanScratchElems(uint64_t n) {
  return (size_t)((n + kScanTile - 1) / kScanTile) + 2;
}

// In-place exclusive scan of data[0..n); scratch holds scanScratchElems(n) uint32.
int launchExclusiveScan(cudaStream_t s, uint32_t* data, uint64_t n, uint32_t* scratch) {
  if (!n) return 0;
  uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  if (tiles == 1) {
    k_scan_tile<<<1, kThreads, 0, s>>>(data, n, nullptr);
    return 1;
  }
  k_scan_tile<<<(unsigned)tiles, kThreads, 0, s>>>(data, n, scratch);
  k_scan_add<<<(unsigned)tiles, kThreads, 0, s>>>(data, n, scratch);
  return 2;
}

// ------------------------------------------------------------------ sort -----
// TILE_MAJOR: hist[tile][digit] (small sorts: the scatter kernel turns it into its own bases, no scan launches between);
// otherwise hist[digit][tile], scanned in place by launchExclusiveScan.
template <bool TILE_MAJOR>
__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint64_t* __restrict__ keys, uint64_t n, int shift,
                                                            uint32_t* __restrict__ hist, uint32_t numTiles) {
  __shared__ uint32_t bins[256];
  bins[threadIdx.x] = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    uint64_t i = base + (uint64_t)k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&bins[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (TILE_MAJOR) hist[(uint64_t)blockIdx.x * 256u + threadIdx.x] = bins[threadIdx.x];
  else hist[(uint64_t)threadIdx.x * numTiles + blockIdx.x] = bins[threadIdx.x];
}

template <bool TILE_MAJOR>
__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint64_t* __restrict__ keys,
                                                               const uint32_t* __restrict__ vals, uint64_t n, int shift,
                                                               const uint32_t* __restrict__ hist, uint32_t numTiles,
                                                               uint64_t* __restrict__ outKeys,
                                                               uint32_t* __restrict__ outVals) {
  constexpr int kWarps = kSortThreads / 32;
  __shared__ uint32_t cnt[kWarps][256];
  __shared__ uint32_t digitBase[256];
  __shared__ uint32_t warpTotals[kWarps];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < kWarps * 256; d += kSortThreads) (&cnt[0][0])[d] = 0;
  if (TILE_MAJOR) {
    // this digit's count over all tiles and over the tiles before this one, then an exclusive scan over the digits
    uint32_t total = 0, before = 0;
    for (uint32_t t = 0; t < numTiles; ++t) {
      uint32_t v = hist[(uint64_t)t * 256u + threadIdx.x];
      total += v;
      before += t < blockIdx.x ? v : 0u;
    }
    uint32_t inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) warpTotals[warp] = inc;
    __syncthreads();
    uint32_t prior = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) prior += w < warp ? warpTotals[w] : 0u;
    digitBase[threadIdx.x] = prior + inc - total + before;
  } else {
    digitBase[threadIdx.x] = hist[(uint64_t)threadIdx.x * numTiles + blockIdx.x];
  }
  __syncthreads();
  // warp w owns elements [w*256, (w+1)*256) of the tile, visited in 8 rounds of 32 => index order
  uint64_t warpBase = (uint64_t)blockIdx.x * kSortTile + (uint64_t)warp * (32 * kSortItems);
  uint64_t key[kSortItems];
  uint32_t rank[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    uint64_t i = warpBase + r * 32 + lane;
    bool valid = i < n;
    key[r] = valid ? keys[i] : 0ull;
    uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & 255u) : 256u + lane;  // invalid lanes never match
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t before = __popc(peers & ((1u << lane) - 1u));
    uint32_t prior = valid ? cnt[warp][d] : 0u;
    __syncwarp();
    if (valid && before == 0) cnt[warp][d] = prior + __popc(peers);
    __syncwarp();
    rank[r] = prior + before;
  }
  __syncthreads();
  {  // exclusive prefix over warps for digit = threadIdx.x, plus the tile's global base
    uint32_t run = digitBase[threadIdx.x];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { uint32_t c = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    uint64_t i = warpBase + r * 32 + lane;
    if (i < n) {
      uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
      uint64_t dst = (uint64_t)cnt[warp][d] + rank[r];
      outKeys[dst] = key[r];
      outVals[dst] = vals[i];
    }
  }
}

size_t sortHistBytes(uint64_t n) {
  uint64_t tiles = (n + kSortTile - 1) / kSortTile;
  uint64_t h = 256 * tiles;
  return (h + 1 + scanScratchElems(h)) * sizeof(uint32_t);
}

int launchSortPairs(cudaStream_t s, uint64_t n, uint64_t* keys, uint32_t* vals, uint64_t* tmpKeys,
                    uint32_t* tmpVals, void* histScratch, int keyBits) {
  if (n < 2 || keyBits <= 0) return 0;
  uint32_t tiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
  uint64_t h = 256ull * tiles;
  uint32_t* hist = reinterpret_cast<uint32_t*>(histScratch);
  uint32_t* scanScratch = hist + h + 1;
  int launches = 0;
  uint64_t *srcK = keys, *dstK = tmpKeys;
  uint32_t *srcV = vals, *dstV = tmpVals;
  const bool small = tiles <= kSortSmallTiles;
  for (int shift = 0; shift < keyBits; shift += 8) {
    if (small) {
      k_sort_hist<true><<<tiles, kSortThreads, 0, s>>>(srcK, n, shift, hist, tiles);
      k_sort_scatter<true><<<tiles, kSortThreads, 0, s>>>(srcK, srcV, n, shift, hist, tiles, dstK, dstV);
      launches += 2;
    } else {
      k_sort_hist<false><<<tiles, kSortThreads, 0, s>>>(srcK, n, shift, hist, tiles);
      launches += 1 + launchExclusiveScan(s, hist, h, scanScratch);
      k_sort_scatter<false><<<tiles, kSortThreads, 0, s>>>(srcK, srcV, n, shift, hist, tiles, dstK, dstV);
      ++launches;
    }
    uint64_t* tk = srcK; srcK = dstK; dstK = tk;
    uint32_t* tv = srcV; srcV = dstV; dstV = tv;
  }
  if (srcK != keys) {
    cudaMemcpyAsync(keys, srcK, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(vals, srcV, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
  }
  return launches;
}

// Loads this file's kernels now (CUDA loads a kernel lazily at its first launch; for the collision kernels that
// would be the first contact tick of a run, ~1 ms each in the middle of the simulation).
void preloadSortKernels() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_scan_tile);
  cudaFuncGetAttributes(&a, k_scan_add);
  cudaFuncGetAttributes(&a, k_sort_hist<false>);
  cudaFuncGetAttributes(&a, k_sort_scatter<false>);
  cudaFuncGetAttributes(&a, k_sort_hist<true>);
  cudaFuncGetAttributes(&a, k_sort_scatter<true>);
}

}  // namespace pies
