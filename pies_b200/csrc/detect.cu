// detect.cu — once-per-substep triangle-triangle collision detection.
//
// Replaces reference Solver::_parallelPointTriangleCollisions (Src/Solver.cpp:680-875) and the
// phmap-backed SpatialHash<Triangle> (Include/Pies/SpatialHash.h) with:
//   1. per-triangle swept cell ranges (TriCompRange, exact float semantics, ccd.cuh);
//   2. (cell key, pair index) pairs -> stable radix sort -> cell-start table;
//   3. a self-join narrow phase: a triangle queries exactly the cells it was inserted into
//      (sweptTriRange == TriCompRange except for the 20-cell cap), visiting its cells in key
//      order (= the reference's (dx,dy,dz) order) and each cell's members in ascending index
//      (= bucket order), three corners per candidate;
//   4. two passes (count, scan in the reference's canonical thread-striped triangle order,
//      write) so the output lists are in the reference's order with its multiplicities
//       without atomics.
// A conservative swept-AABB cull (margin = threshold + slop) skips candidates that cannot
// pass pointTriangleCCD; it never changes the result.
#include "detect.h"

#include "ccd.cuh"

namespace pies {

static inline int gridFor(uint64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// ---- 1. ranges -----------------------------------------------------------------------------
// bbox[0..2] = min cell, bbox[3..5] = max cell (inclusive), bbox[6] = bad-input flag
__global__ void __launch_bounds__(kThreads) k_tri_ranges(uint32_t nTri, const uint32_t* __restrict__ tri,
                                                         const float4* __restrict__ q, const float4* __restrict__ prev,
                                                         int4* __restrict__ triMin, uint32_t* __restrict__ triLen,
                                                         uint32_t* __restrict__ cnt, float4* __restrict__ aabbLo,
                                                         float4* __restrict__ aabbHi, int* __restrict__ bbox) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nTri) return;
  uint32_t a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
  V3 p0 = v3(q[a]), p1 = v3(q[b]), p2 = v3(q[c]);
  V3 o0 = v3(prev[a]), o1 = v3(prev[b]), o2 = v3(prev[c]);
  int mx, my, mz;
  unsigned lx, ly, lz;
  bool bad;
  ex::triCellRange(p0, p1, p2, o0, o1, o2, mx, my, mz, lx, ly, lz, bad);
  if (bad) atomicExch(bbox + 6, 1);
  if (lx > 50u || ly > 50u || lz > 50u) lx = ly = lz = 0;  // TriCompRange cap: not inserted at all
  triMin[t] = make_int4(mx, my, mz, 0);
  triLen[t] = lx | (ly << 8) | (lz << 16);
  uint32_t cells = lx * ly * lz;
  cnt[t] = cells;
  if (cells) {
    atomicMin(bbox + 0, mx); atomicMin(bbox + 1, my); atomicMin(bbox + 2, mz);
    atomicMax(bbox + 3, mx + (int)lx - 1); atomicMax(bbox + 4, my + (int)ly - 1); atomicMax(bbox + 5, mz + (int)lz - 1);
  }
  aabbLo[t] = make_float4(fminf(fminf(fminf(p0.x, o0.x), fminf(p1.x, o1.x)), fminf(p2.x, o2.x)),
                          fminf(fminf(fminf(p0.y, o0.y), fminf(p1.y, o1.y)), fminf(p2.y, o2.y)),
                          fminf(fminf(fminf(p0.z, o0.z), fminf(p1.z, o1.z)), fminf(p2.z, o2.z)), 0.0f);
  aabbHi[t] = make_float4(fmaxf(fmaxf(fmaxf(p0.x, o0.x), fmaxf(p1.x, o1.x)), fmaxf(p2.x, o2.x)),
                          fmaxf(fmaxf(fmaxf(p0.y, o0.y), fmaxf(p1.y, o1.y)), fmaxf(p2.y, o2.y)),
                          fmaxf(fmaxf(fmaxf(p0.z, o0.z), fmaxf(p1.z, o1.z)), fmaxf(p2.z, o2.z)), 0.0f);
}

// ---- 2. pairs ------------------------------------------------------------------------------
struct KeyPack { int minX, minY, minZ; int bitsY, bitsZ; };

__global__ void __launch_bounds__(kThreads) k_emit_pairs(uint32_t nTri, const int4* __restrict__ triMin,
                                                         const uint32_t* __restrict__ triLen,
                                                         const uint32_t* __restrict__ off, KeyPack kp,
                                                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                         uint32_t* __restrict__ pairTri) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nTri) return;
  uint32_t len = triLen[t];
  uint32_t lx = len & 255u, ly = (len >> 8) & 255u, lz = (len >> 16) & 255u;
  if (!(lx * ly * lz)) return;
  int4 m = triMin[t];
  uint32_t i = off[t];
  for (uint32_t dx = 0; dx < lx; ++dx)
    for (uint32_t dy = 0; dy < ly; ++dy)
      for (uint32_t dz = 0; dz < lz; ++dz, ++i) {
        uint64_t kx = (uint64_t)(uint32_t)(m.x + (int)dx - kp.minX);
        uint64_t ky = (uint64_t)(uint32_t)(m.y + (int)dy - kp.minY);
        uint64_t kz = (uint64_t)(uint32_t)(m.z + (int)dz - kp.minZ);
        keys[i] = (kx << (kp.bitsY + kp.bitsZ)) | (ky << kp.bitsZ) | kz;
        vals[i] = i;
        pairTri[i] = t;
      }
}

// ---- 3. cell-start table -------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_mark_heads(uint64_t nPairs, const uint64_t* __restrict__ keys,
                                                         const uint32_t* __restrict__ vals,
                                                         const uint32_t* __restrict__ pairTri,
                                                         uint32_t* __restrict__ heads, uint32_t* __restrict__ memberTri,
                                                         uint32_t* __restrict__ posOf) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > nPairs) return;
  if (j == nPairs) { heads[j] = 0; return; }
  heads[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
  uint32_t i = vals[j];
  memberTri[j] = pairTri[i];
  posOf[i] = (uint32_t)j;
}

// cellIdx (in place of the scanned heads) and cellStart
__global__ void __launch_bounds__(kThreads) k_cell_starts(uint64_t nPairs, const uint64_t* __restrict__ keys,
                                                          uint32_t* __restrict__ headScan /* -> cellIdx */,
                                                          uint32_t* __restrict__ cellStart) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nPairs) return;
  bool head = (j == 0 || keys[j] != keys[j - 1]);
  uint32_t idx = headScan[j] + (head ? 1u : 0u) - 1u;
  if (head) cellStart[idx] = (uint32_t)j;
  if (j == nPairs - 1) cellStart[idx + 1] = (uint32_t)nPairs;
  headScan[j] = idx;
}

// ---- 4. narrow phase ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t canonicalRank(uint32_t t, uint32_t nTri, uint32_t T) {
  // thread (t % T) handles t, t+T, ...; per-thread lists are concatenated in thread order (Solver.cpp:714,852-873)
  uint32_t th = t % T, k = t / T;
  uint32_t full = nTri / T, rem = nTri % T;  // threads < rem own full+1 triangles
  return th * full + (th < rem ? th : rem) + k;
}

struct NarrowParams {
  uint32_t nTri, threadCount;
  float threshold, floorLimit;  // floorHeight + thickness
  float cullMargin;
};

template <bool WRITE>
__global__ void __launch_bounds__(128) k_narrow(NarrowParams np, const uint32_t* __restrict__ tri,
                                                const float4* __restrict__ q, const float4* __restrict__ prev,
                                                const uint32_t* __restrict__ triLen, const uint32_t* __restrict__ off,
                                                const uint32_t* __restrict__ posOf, const uint32_t* __restrict__ cellIdx,
                                                const uint32_t* __restrict__ cellStart,
                                                const uint32_t* __restrict__ memberTri,
                                                const float4* __restrict__ aabbLo, const float4* __restrict__ aabbHi,
                                                uint32_t* __restrict__ hitCount /* canonical order, scanned when WRITE */,
                                                uint32_t* __restrict__ floorCount, uint4* __restrict__ outTri,
                                                uint32_t* __restrict__ outFloor, int* __restrict__ failFlag) {
  uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (t >= np.nTri) return;
  uint32_t rank = canonicalRank(t, np.nTri, np.threadCount);
  uint32_t ia[3] = {tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]};
  V3 pa[3], oa[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { pa[i] = v3(q[ia[i]]); oa[i] = v3(prev[ia[i]]); }
  uint32_t len = triLen[t];
  uint32_t lx = len & 255u, ly = (len >> 8) & 255u, lz = (len >> 16) & 255u;
  uint32_t nCells = lx * ly * lz;
  if (lx > 20u || ly > 20u || lz > 20u) nCells = 0;  // sweptTriRange cap: inserted but queries nothing
  if (nCells > 1000u) { if (lane == 0) atomicExch(failFlag, 1); nCells = 0; }  // hang guard, Solver.cpp:741-745
  uint32_t base = off[t];
  uint32_t outPos = WRITE ? hitCount[rank] : 0u;
  uint32_t total = 0;
  // swept boxes of the three corners, padded by the cull margin
  float cLo[3][3], cHi[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    cLo[i][0] = fminf(pa[i].x, oa[i].x) - np.cullMargin; cHi[i][0] = fmaxf(pa[i].x, oa[i].x) + np.cullMargin;
    cLo[i][1] = fminf(pa[i].y, oa[i].y) - np.cullMargin; cHi[i][1] = fmaxf(pa[i].y, oa[i].y) + np.cullMargin;
    cLo[i][2] = fminf(pa[i].z, oa[i].z) - np.cullMargin; cHi[i][2] = fmaxf(pa[i].z, oa[i].z) + np.cullMargin;
  }
  for (uint32_t k = 0; k < nCells; ++k) {
    uint32_t j = posOf[base + k];
    uint32_t cidx = cellIdx[j];
    uint32_t s = cellStart[cidx], e = cellStart[cidx + 1];
    if (e - s > 1000u) { if (lane == 0) atomicExch(failFlag, 1); }  // hang guard, Solver.cpp:751-755
    for (uint32_t m0 = s; m0 < e; m0 += 32) {
      uint32_t m = m0 + lane;
      uint32_t mask = 0;
      uint32_t ib = 0, ic = 0, id = 0;
      if (m < e) {
        uint32_t o = memberTri[m];
        ib = tri[3 * o]; ic = tri[3 * o + 1]; id = tri[3 * o + 2];
        bool common = false;
#pragma unroll
        for (int i = 0; i < 3; ++i) common |= (ia[i] == ib) | (ia[i] == ic) | (ia[i] == id);
        if (!common) {
          float4 lo = aabbLo[o], hi = aabbHi[o];
          uint32_t maybe = 0;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            bool ov = cLo[i][0] <= hi.x && cHi[i][0] >= lo.x && cLo[i][1] <= hi.y && cHi[i][1] >= lo.y &&
                      cLo[i][2] <= hi.z && cHi[i][2] >= lo.z;
            maybe |= ov ? (1u << i) : 0u;
          }
          if (maybe) {
            V3 pb = v3(q[ib]), pc = v3(q[ic]), pd = v3(q[id]);
            V3 ob = v3(prev[ib]), oc = v3(prev[ic]), od = v3(prev[id]);
            V3 ab0 = ex::sub(oc, ob), ac0 = ex::sub(od, ob), ab1 = ex::sub(pc, pb), ac1 = ex::sub(pd, pb);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              if (!(maybe & (1u << i))) continue;
              float tt;
              if (ex::pointTriangleCCD(ex::sub(oa[i], ob), ab0, ac0, ex::sub(pa[i], pb), ab1, ac1, np.threshold, tt))
                mask |= 1u << i;
            }
          }
        }
      }
      uint32_t hits = __popc(mask);
      // ordered by member (lane) then corner
      uint32_t incl = hits;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      uint32_t chunkTotal = __shfl_sync(0xffffffffu, incl, 31);
      if (WRITE && hits) {
        uint32_t w = outPos + total + incl - hits;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (mask & (1u << i)) outTri[w++] = make_uint4(ia[i], ib, ic, id);
      }
      total += chunkTotal;
    }
  }
  if (lane == 0) {
    // floor test per corner (Solver.cpp:829-834)
    uint32_t f = 0;
    uint32_t fbase = WRITE ? floorCount[rank] : 0u;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (pa[i].y < np.floorLimit) { if (WRITE) outFloor[fbase + f] = ia[i]; ++f; }
    if (!WRITE) { hitCount[rank] = total; floorCount[rank] = f; }
  }
}

// ---- 5. node -> incident entries --------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_inc_emit(uint32_t nTriC, const uint4* __restrict__ entries,
                                                       uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                       uint32_t* __restrict__ incCount) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nTriC) return;
  uint4 v = entries[e];
  uint32_t ids[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    keys[4ull * e + s] = ids[s];
    vals[4ull * e + s] = 4u * e + s;
    atomicAdd(incCount + ids[s], 1u);
  }
}

__global__ void __launch_bounds__(kThreads) k_floor_mult(uint32_t nFloor, const uint32_t* __restrict__ nodes,
                                                         uint32_t* __restrict__ mult) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nFloor) atomicAdd(mult + nodes[i], 1u);
}

__global__ void __launch_bounds__(kThreads) k_floor_weight(uint32_t n, const uint32_t* __restrict__ mult,
                                                           float* __restrict__ w) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.0f;
  for (uint32_t k = 0; k < mult[i]; ++k) acc += 10000.0f;  // StaticCollisionConstraint::w, coeffRef += per duplicate
  w[i] = acc;
}

__global__ void k_init_bbox(int* bbox) {
  bbox[0] = bbox[1] = bbox[2] = 0x7fffffff;
  bbox[3] = bbox[4] = bbox[5] = (int)0x80000000;
  bbox[6] = 0; bbox[7] = 0;
}

static int bitsFor(int64_t span) {  // bits to hold values 0..span
  int b = 1;
  while ((int64_t(1) << b) <= span) ++b;
  return b;
}

#define DCHECK(expr)                                                    \
  do {                                                                  \
    cudaError_t _e = (expr);                                            \
    if (_e != cudaSuccess) { w.lastError = _e; return -1; }             \
  } while (0)

int detectTriangles(DetectWork& w, cudaStream_t s, const DetectInput& in, ContactLists& out, int* launches) {
  int L = 0;
  const uint32_t nTri = in.nTri, n = in.nNodes;
  out.nTri = out.nFloor = 0;
  w.nPairs = 0; w.nCells = 0; w.failed = false; w.badInput = false;
  DCHECK(w.incPtr.reserve(n + 2));
  DCHECK(w.floorMult.reserve(n + 1));
  DCHECK(w.floorW.reserve(n + 1));
  out.incPtr = (int*)w.incPtr.p; out.floorW = w.floorW.p; out.floorMult = w.floorMult.p;
  if (!nTri) return 0;
  DCHECK(w.triMin.reserve(nTri)); DCHECK(w.triLen.reserve(nTri)); DCHECK(w.cnt.reserve(nTri + 2));
  DCHECK(w.aabbLo.reserve(nTri)); DCHECK(w.aabbHi.reserve(nTri));
  DCHECK(w.hitCount.reserve(nTri + 2)); DCHECK(w.floorCount.reserve(nTri + 2));
  DCHECK(w.bbox.reserve(8));
  DCHECK(w.scanScratch.reserve(scanScratchElems(std::max<uint64_t>(nTri + 2, w.scanCap))));
  k_init_bbox<<<1, 1, 0, s>>>(w.bbox.p); ++L;
  DCHECK(cudaMemsetAsync(w.cnt.p, 0, (nTri + 2) * sizeof(uint32_t), s));
  k_tri_ranges<<<gridFor(nTri, kThreads), kThreads, 0, s>>>(nTri, in.tri, in.q, in.prev, w.triMin.p, w.triLen.p, w.cnt.p,
                                                           w.aabbLo.p, w.aabbHi.p, w.bbox.p); ++L;
  L += launchExclusiveScan(s, w.cnt.p, nTri + 1, w.scanScratch.p);
  DCHECK(cudaMemcpyAsync(w.host, w.bbox.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
  DCHECK(cudaMemcpyAsync(w.host + 8, w.cnt.p + nTri, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  DCHECK(cudaStreamSynchronize(s));
  if (w.host[6]) { w.badInput = true; return 0; }
  uint64_t nPairs = (uint32_t)w.host[8];
  w.nPairs = nPairs;
  NarrowParams np{nTri, in.threadCount ? in.threadCount : 1u, in.threshold, in.floorLimit, in.threshold + 1e-3f};
  if (nPairs) {
    KeyPack kp{w.host[0], w.host[1], w.host[2], 0, 0};
    int bx = bitsFor((int64_t)w.host[3] - w.host[0]), by = bitsFor((int64_t)w.host[4] - w.host[1]),
        bz = bitsFor((int64_t)w.host[5] - w.host[2]);
    kp.bitsY = by; kp.bitsZ = bz;
    if (bx + by + bz > 63) { w.badInput = true; return 0; }
    w.keyPack[0] = kp.minX; w.keyPack[1] = kp.minY; w.keyPack[2] = kp.minZ; w.keyPack[3] = by; w.keyPack[4] = bz;
    DCHECK(w.keys.reserve(nPairs)); DCHECK(w.tmpKeys.reserve(nPairs));
    DCHECK(w.vals.reserve(nPairs)); DCHECK(w.tmpVals.reserve(nPairs));
    DCHECK(w.pairTri.reserve(nPairs)); DCHECK(w.posOf.reserve(nPairs)); DCHECK(w.memberTri.reserve(nPairs));
    DCHECK(w.heads.reserve(nPairs + 2)); DCHECK(w.cellStart.reserve(nPairs + 2));
    DCHECK(w.sortHist.reserve(sortHistBytes(nPairs) / 4 + 4));
    w.scanCap = std::max<uint64_t>(w.scanCap, nPairs + 2);
    DCHECK(w.scanScratch.reserve(scanScratchElems(w.scanCap)));
    k_emit_pairs<<<gridFor(nTri, kThreads), kThreads, 0, s>>>(nTri, w.triMin.p, w.triLen.p, w.cnt.p, kp, w.keys.p, w.vals.p,
                                                             w.pairTri.p); ++L;
    L += launchSortPairs(s, nPairs, w.keys.p, w.vals.p, w.tmpKeys.p, w.tmpVals.p, w.sortHist.p, bx + by + bz);
    k_mark_heads<<<gridFor(nPairs + 1, kThreads), kThreads, 0, s>>>(nPairs, w.keys.p, w.vals.p, w.pairTri.p, w.heads.p,
                                                                   w.memberTri.p, w.posOf.p); ++L;
    L += launchExclusiveScan(s, w.heads.p, nPairs + 1, w.scanScratch.p);
    DCHECK(cudaMemcpyAsync(w.host + 9, w.heads.p + nPairs, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    k_cell_starts<<<gridFor(nPairs, kThreads), kThreads, 0, s>>>(nPairs, w.keys.p, w.heads.p, w.cellStart.p); ++L;
  }
  // count pass
  DCHECK(cudaMemsetAsync(w.hitCount.p, 0, (nTri + 2) * sizeof(uint32_t), s));
  DCHECK(cudaMemsetAsync(w.floorCount.p, 0, (nTri + 2) * sizeof(uint32_t), s));
  k_narrow<false><<<gridFor((uint64_t)nTri * 32, 128), 128, 0, s>>>(np, in.tri, in.q, in.prev, w.triLen.p, w.cnt.p, w.posOf.p,
                                                                    w.heads.p, w.cellStart.p, w.memberTri.p, w.aabbLo.p,
                                                                    w.aabbHi.p, w.hitCount.p, w.floorCount.p, nullptr, nullptr,
                                                                    w.bbox.p + 7); ++L;
  L += launchExclusiveScan(s, w.hitCount.p, nTri + 1, w.scanScratch.p);
  L += launchExclusiveScan(s, w.floorCount.p, nTri + 1, w.scanScratch.p);
  DCHECK(cudaMemcpyAsync(w.host + 10, w.hitCount.p + nTri, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  DCHECK(cudaMemcpyAsync(w.host + 11, w.floorCount.p + nTri, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  DCHECK(cudaMemcpyAsync(w.host + 12, w.bbox.p + 7, sizeof(int), cudaMemcpyDeviceToHost, s));
  DCHECK(cudaStreamSynchronize(s));
  w.nCells = nPairs ? (uint32_t)w.host[9] : 0;
  if (w.host[12]) { w.failed = true; if (launches) *launches += L; return 0; }  // reference latches _simFailed, lists stay empty
  uint32_t nHit = (uint32_t)w.host[10], nFloor = (uint32_t)w.host[11];
  DCHECK(w.triList.reserve(nHit + 1)); DCHECK(w.floorList.reserve(nFloor + 1));
  k_narrow<true><<<gridFor((uint64_t)nTri * 32, 128), 128, 0, s>>>(np, in.tri, in.q, in.prev, w.triLen.p, w.cnt.p, w.posOf.p,
                                                                   w.heads.p, w.cellStart.p, w.memberTri.p, w.aabbLo.p,
                                                                   w.aabbHi.p, w.hitCount.p, w.floorCount.p, w.triList.p,
                                                                   w.floorList.p, w.bbox.p + 7); ++L;
  out.tri = w.triList.p; out.floorNode = w.floorList.p; out.nTri = nHit; out.nFloor = nFloor;
  // incidence CSR + floor multiplicities
  DCHECK(cudaMemsetAsync(w.incPtr.p, 0, (n + 2) * sizeof(uint32_t), s));
  DCHECK(cudaMemsetAsync(w.floorMult.p, 0, (n + 1) * sizeof(uint32_t), s));
  if (nHit) {
    uint64_t nInc = 4ull * nHit;
    DCHECK(w.incKeys.reserve(nInc)); DCHECK(w.incTmpKeys.reserve(nInc));
    DCHECK(w.incVals.reserve(nInc)); DCHECK(w.incTmpVals.reserve(nInc));
    DCHECK(w.sortHist.reserve(sortHistBytes(nInc) / 4 + 4));
    w.scanCap = std::max<uint64_t>(w.scanCap, std::max<uint64_t>(nInc, n) + 2);
    DCHECK(w.scanScratch.reserve(scanScratchElems(w.scanCap)));
    k_inc_emit<<<gridFor(nHit, kThreads), kThreads, 0, s>>>(nHit, w.triList.p, w.incKeys.p, w.incVals.p, w.incPtr.p); ++L;
    L += launchExclusiveScan(s, w.incPtr.p, n + 1, w.scanScratch.p);
    L += launchSortPairs(s, nInc, w.incKeys.p, w.incVals.p, w.incTmpKeys.p, w.incTmpVals.p, w.sortHist.p, bitsFor(n));
    out.inc = w.incVals.p;
  }
  if (nFloor) { k_floor_mult<<<gridFor(nFloor, kThreads), kThreads, 0, s>>>(nFloor, w.floorList.p, w.floorMult.p); ++L; }
  k_floor_weight<<<gridFor(n, kThreads), kThreads, 0, s>>>(n, w.floorMult.p, w.floorW.p); ++L;
  if (launches) *launches += L;
  return 0;
}

}  // namespace pies
