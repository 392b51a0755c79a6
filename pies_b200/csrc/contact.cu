// contact.cu — collision response of the PD path: per-iteration projections of the
// point-triangle / floor constraints, collision stabilisation and the friction pass.
//
// Reference: Src/CollisionConstraint.cpp:86-194 (point-triangle), :439-463 (floor),
// Src/Solver.cpp:298-308,337-349 (local step / RHS), :367-383 (stabilisation), :431-484 (friction).
//
// Stabilisation and friction are sequential Gauss-Seidel sweeps over the collision list in the
// reference, and the result depends on that order.  They run here as dataflow sweeps
// that honour exactly that order per node (k_gs_dataflow below).
#include "contact.h"
#include "reblock.h"

#include "common.cuh"

namespace pies {

static inline int gridFor(uint64_t n, int threads) { return (int)((n + threads - 1) / threads); }

constexpr float kFloorW = 10000.0f;  // StaticCollisionConstraint::w (CollisionConstraint.h:78)

// ---- per-iteration projections (parallel, no ordering issue) ---------------------------------------
// contribC[4u + slot] = copies * w * (A^T A p)_slot for distinct contact u, with p = current positions and
// p_A pushed to `thickness` above the triangle plane when below it (CollisionConstraint.cpp:86-124, :176-194).
// The reference adds the identical term once per copy; the copies are folded into the weight here.
__global__ void __launch_bounds__(kThreads) k_pt_project(uint32_t nTri, const uint4* __restrict__ entries,
                                                         const float* __restrict__ weight,
                                                         const float4* __restrict__ q, float thickness,
                                                         float4* __restrict__ contribC) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nTri) return;
  uint4 id = entries[e];
  const float kPtW = weight[e];  // copies * PointTriangleCollisionConstraint::w
  V3 A = v3(q[id.x]), B = v3(q[id.y]), C = v3(q[id.z]), D = v3(q[id.w]);
  V3 nrm = normalize(cross(C - B, D - B));
  float nDotP = dot(nrm, A - B);
  V3 pA = A;
  if (nDotP < thickness) pA += (thickness - nDotP) * nrm;
  V3 c0 = 3.0f * pA + (-1.0f) * B + (-1.0f) * C + (-1.0f) * D;
  contribC[4ull * e + 0] = f4(kPtW * c0, 0.0f);
  contribC[4ull * e + 1] = f4(kPtW * ((-1.0f) * pA + B), 0.0f);
  contribC[4ull * e + 2] = f4(kPtW * ((-1.0f) * pA + C), 0.0f);
  contribC[4ull * e + 3] = f4(kPtW * ((-1.0f) * pA + D), 0.0f);
}

// projectedPosition of every floor contact (CollisionConstraint.cpp:447-455): position with y<0 -> 0.
// Duplicates of a node all carry the same value, so it is stored per node.
__global__ void __launch_bounds__(kThreads) k_floor_project(uint32_t nFloor, const uint32_t* __restrict__ nodes,
                                                            const float4* __restrict__ q, float4* __restrict__ snap) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nFloor) return;
  uint32_t node = nodes[i];
  float4 p = q[node];
  if (p.y < 0.0f) p.y = 0.0f;
  snap[node] = p;
}

int launchContactProject(cudaStream_t s, const ContactLists& c, const float4* q, float thickness, float4* contribC,
                         float4* snap) {
  int L = 0;
  if (c.nUnique) { k_pt_project<<<gridFor(c.nUnique, kThreads), kThreads, 0, s>>>(c.nUnique, c.uTri, c.uW, q, thickness, contribC); ++L; }
  if (c.nFloor) { k_floor_project<<<gridFor(c.nFloor, kThreads), kThreads, 0, s>>>(c.nFloor, c.floorNode, q, snap); ++L; }
  return L;
}

// rhs_i += sum of collision contributions (list order) + per floor duplicate w * projectedPosition
// (Solver.cpp:337-349).
__global__ void __launch_bounds__(kThreads) k_gather_contacts(uint32_t n, ContactLists c,
                                                              const float4* __restrict__ contribC,
                                                              const float4* __restrict__ snap, float4* __restrict__ rhs) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int beg = 0, end = 0;
  if (c.nTri) { beg = c.incPtr[i]; end = c.incPtr[i + 1]; }
  uint32_t mult = c.nFloor ? c.floorMult[i] : 0u;
  if (beg == end && !mult) return;
  float4 acc = rhs[i];
  for (int k = beg; k < end; ++k) {
    float4 v = contribC[c.inc[k]];
    acc.x += v.x; acc.y += v.y; acc.z += v.z;
  }
  if (mult) {
    float4 p = snap[i];
    for (uint32_t k = 0; k < mult; ++k) { acc.x += kFloorW * p.x; acc.y += kFloorW * p.y; acc.z += kFloorW * p.z; }
  }
  rhs[i] = acc;
}

int launchGatherContacts(cudaStream_t s, uint32_t n, const ContactLists& c, const float4* contribC, const float4* snap,
                         float4* rhs) {
  if (!c.nTri && !c.nFloor) return 0;
  k_gather_contacts<<<gridFor(n, kThreads), kThreads, 0, s>>>(n, c, contribC, snap, rhs);
  return 1;
}

// ---- ordered Gauss-Seidel executor (dataflow) ---------------------------------------------------------
// The reference sweeps the collision list sequentially; entry e reads and writes its four nodes, so the
// only constraint a parallel schedule has to respect is, per node, the list order of the entries that
// touch it.  Detection gives every (entry, node) its ticket = number of earlier entries touching that
// node.  Warps take chunks of 32 consecutive entries from a global counter (so every chunk below the
// newest one is owned by a running warp); an entry executes as soon as nodeDone[v] == ticket for its
// four nodes, then publishes nodeDone[v] = ticket + 1.  The earliest unfinished entry is always
// runnable, so the sweep cannot stall, and the result equals the sequential sweep up to rounding of
// the identical operations (same operand values, same order per node).
// (A chunk-level variant that runs the 32 entries of a chunk from shared memory was measured and dropped: at
// chunk granularity the false dependencies serialise a percolated pile, 29 ms per sweep instead of 1.1 ms.)
constexpr int kGsThreads = 256;

struct StabilizeOp {
  float4* q;
  float4* prev;
  float thickness;
  // PointTriangleCollisionConstraint::stabilizeCollisions (CollisionConstraint.cpp:126-162)
  __device__ __forceinline__ void operator()(uint4 id) const {
    float4 a4 = __ldcg(q + id.x), b4 = __ldcg(q + id.y), c4 = __ldcg(q + id.z), d4 = __ldcg(q + id.w);
    V3 A = v3(a4), B = v3(b4), C = v3(c4), D = v3(d4);
    V3 nrm = normalize(cross(C - B, D - B));
    float nDotP = dot(nrm, A - B);
    if (!(nDotP < thickness)) return;
    V3 disp = (thickness - nDotP) * nrm;
    float wTri = b4.w + c4.w + d4.w;
    float wSum = a4.w + wTri;
    V3 da = disp * a4.w / wSum, dt = disp * wTri / wSum;
    float4 pa = __ldcg(prev + id.x), pb = __ldcg(prev + id.y), pc = __ldcg(prev + id.z), pd = __ldcg(prev + id.w);
    __stcg(q + id.x, f4(A + da, a4.w)); __stcg(q + id.y, f4(B - dt, b4.w));
    __stcg(q + id.z, f4(C - dt, c4.w)); __stcg(q + id.w, f4(D - dt, d4.w));
    __stcg(prev + id.x, f4(v3(pa) + da, pa.w)); __stcg(prev + id.y, f4(v3(pb) - dt, pb.w));
    __stcg(prev + id.z, f4(v3(pc) - dt, pc.w)); __stcg(prev + id.w, f4(v3(pd) - dt, pd.w));
  }
};

struct FrictionOp {
  const float4* q;
  float4* vel;
  float friction, staticThreshold;
  // point-triangle friction / restitution (Solver.cpp:431-471); positions are not modified by this pass
  __device__ __forceinline__ void operator()(uint4 id) const {
    float4 a4 = q[id.x], b4 = q[id.y], c4 = q[id.z], d4 = q[id.w];
    V3 va = v3(__ldcg(vel + id.x)), vb = v3(__ldcg(vel + id.y)), vc = v3(__ldcg(vel + id.z)), vd = v3(__ldcg(vel + id.w));
    V3 avgTri = (vb + vc + vd) / 3.0f;
    V3 nrm = normalize(cross(v3(c4) - v3(b4), v3(d4) - v3(b4)));
    V3 rel = va - avgTri;
    float vDotN = dot(rel, nrm);
    V3 perp = rel - vDotN * nrm;
    float fr = friction;
    if (length(perp) < staticThreshold) fr = 1.0f;
    float triW = b4.w + c4.w + d4.w;
    float wSum = a4.w + triW;
    V3 dv = (-fr) * perp - (1.1f * fminf(vDotN, 0.0f)) * nrm;
    V3 dtv = (-dv) * triW / wSum;
    __stcg(vel + id.x, f4(va + dv * a4.w / wSum, 0.0f));
    __stcg(vel + id.y, f4(vb + dtv, 0.0f)); __stcg(vel + id.z, f4(vc + dtv, 0.0f)); __stcg(vel + id.w, f4(vd + dtv, 0.0f));
  }
};

template <typename Op>
__global__ void __launch_bounds__(kGsThreads) k_gs_dataflow(const uint4* __restrict__ entries,
                                                            const uint4* __restrict__ ticket, uint32_t nTri,
                                                            const uint8_t* __restrict__ gsClass,
                                                            uint32_t* __restrict__ nodeDone,
                                                            uint32_t* __restrict__ chunkCounter, Op op) {
  const int lane = threadIdx.x & 31;
  const uint32_t nChunks = (nTri + 31u) >> 5;
  while (true) {
    uint32_t chunk = 0;
    if (lane == 0) chunk = atomicAdd(chunkCounter, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk >= nChunks) break;
    uint32_t e = (chunk << 5) + (uint32_t)lane;
    bool pending = e < nTri;
    uint4 id = make_uint4(0, 0, 0, 0), tk = make_uint4(0, 0, 0, 0);
    if (pending) {
      id = entries[e];
      // entries of small clusters are swept inside one warp (k_gs_cluster_*); they share no node with the rest
      if (gsClass[id.x] != 2) pending = false; else tk = ticket[e];
    }
    while (__any_sync(0xffffffffu, pending)) {
      if (pending) {
        bool ready = ldAcquire(nodeDone + id.x) == tk.x && ldAcquire(nodeDone + id.y) == tk.y &&
                     ldAcquire(nodeDone + id.z) == tk.z && ldAcquire(nodeDone + id.w) == tk.w;
        if (ready) {
          op(id);
          __threadfence();
          stRelease(nodeDone + id.x, tk.x + 1u); stRelease(nodeDone + id.y, tk.y + 1u);
          stRelease(nodeDone + id.z, tk.z + 1u); stRelease(nodeDone + id.w, tk.w + 1u);
          pending = false;
        }
      }
    }
  }
}

// ---- ordered sweeps of small contact clusters inside one warp -------------------------------------------
// A contact cluster (connected component of the contact graph, reblock.cu) of <= 32 nodes never interacts
// with anything outside it during the sweeps, so one warp keeps its nodes in registers (lane = node),
// walks the cluster's entries in list order and exchanges operands with shuffles: the sequential order of
// the reference at register latency, all four stabilisation sweeps in one launch.
__global__ void __launch_bounds__(kThreads) k_entry_keys(uint32_t nTri, const uint4* __restrict__ entries,
                                                         const uint32_t* __restrict__ clusterOf,
                                                         const uint32_t* __restrict__ rankOf, uint64_t* __restrict__ keys,
                                                         uint32_t* __restrict__ lanes, uint32_t* __restrict__ entCount) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nTri) return;
  uint4 id = entries[e];
  uint32_t c = clusterOf[id.x];
  // The sort only looks at the cluster bits, so the high word of the key travels as a second payload:
  // ranks of the four nodes inside their cluster, 16 bits each (unused above 65535 nodes: dataflow sweeps).
  keys[e] = ((uint64_t)((rankOf[id.z] & 0xffffu) | (rankOf[id.w] << 16)) << 32) | c;
  lanes[e] = (rankOf[id.x] & 0xffffu) | (rankOf[id.y] << 16);
  atomicAdd(entCount + c, 1u);
}

__device__ __forceinline__ V3 shflV3(float4 v, int src) {
  return v3(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src), __shfl_sync(0xffffffffu, v.z, src));
}

__global__ void __launch_bounds__(kThreads) k_gs_cluster_stabilize(ClusterView cv, float4* __restrict__ q,
                                                                   float4* __restrict__ prev, const float4* __restrict__ snap,
                                                                   const uint32_t* __restrict__ floorMult, int haveFloor,
                                                                   float thickness, uint32_t sweeps) {
  const int lane = threadIdx.x & 31;
  const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
  const uint32_t nC = *cv.nClusters;
  for (uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < nC; c += warpsPerGrid) {
    const uint32_t nb = cv.start[c], size = cv.start[c + 1] - nb;
    if (size > 32u) continue;  // large cluster: dataflow sweeps
    const bool have = (uint32_t)lane < size;
    const uint32_t node = have ? cv.nodes[nb + lane] : 0u;
    float4 q4 = make_float4(0.0f, 0.0f, 0.0f, 1.0f), p4 = q4, s4 = q4;
    bool onFloor = false;
    if (have) {
      q4 = q[node]; p4 = prev[node];
      onFloor = haveFloor && floorMult[node] != 0u;
      if (onFloor) s4 = snap[node];
    }
    const uint32_t eb = cv.entStart[c], ee = cv.entStart[c + 1];
    for (uint32_t sweep = 0; sweep < sweeps; ++sweep) {
      for (uint32_t base = eb; base < ee; base += 32) {
        const uint32_t myW = base + lane < ee ? cv.lanes[base + lane] : 0u;
        const uint32_t myW2 = base + lane < ee ? (uint32_t)(cv.keys[base + lane] >> 32) : 0u;
        const int cnt = (int)min(32u, ee - base);
        for (int i = 0; i < cnt; ++i) {
          const uint32_t lw = __shfl_sync(0xffffffffu, myW, i), lw2 = __shfl_sync(0xffffffffu, myW2, i);
          const int la = lw & 31u, lb = (lw >> 16) & 31u, lc = lw2 & 31u, ld = (lw2 >> 16) & 31u;
          // PointTriangleCollisionConstraint::stabilizeCollisions (CollisionConstraint.cpp:126-162), as StabilizeOp
          V3 A = shflV3(q4, la), B = shflV3(q4, lb), C = shflV3(q4, lc), D = shflV3(q4, ld);
          float wa = __shfl_sync(0xffffffffu, q4.w, la), wb = __shfl_sync(0xffffffffu, q4.w, lb),
                wc = __shfl_sync(0xffffffffu, q4.w, lc), wd = __shfl_sync(0xffffffffu, q4.w, ld);
          V3 nrm = normalize(cross(C - B, D - B));
          float nDotP = dot(nrm, A - B);
          if (!(nDotP < thickness)) continue;
          V3 disp = (thickness - nDotP) * nrm;
          float wTri = wb + wc + wd;
          float wSum = wa + wTri;
          V3 da = disp * wa / wSum, dt = disp * wTri / wSum;
          if (lane == la) {
            q4.x += da.x; q4.y += da.y; q4.z += da.z; p4.x += da.x; p4.y += da.y; p4.z += da.z;
          } else if (lane == lb || lane == lc || lane == ld) {
            q4.x -= dt.x; q4.y -= dt.y; q4.z -= dt.z; p4.x -= dt.x; p4.y -= dt.y; p4.z -= dt.z;
          }
        }
      }
      if (onFloor) { q4.x = s4.x; q4.y = s4.y; q4.z = s4.z; }  // Solver.cpp:379-382, after the sweep's tri entries
    }
    if (have) { q[node] = q4; prev[node] = p4; }
  }
}

__global__ void __launch_bounds__(kThreads) k_gs_cluster_friction(ClusterView cv, const float4* __restrict__ q,
                                                                  float4* __restrict__ vel, float friction,
                                                                  float staticThreshold) {
  const int lane = threadIdx.x & 31;
  const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
  const uint32_t nC = *cv.nClusters;
  for (uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < nC; c += warpsPerGrid) {
    const uint32_t nb = cv.start[c], size = cv.start[c + 1] - nb;
    if (size > 32u) continue;
    const bool have = (uint32_t)lane < size;
    const uint32_t node = have ? cv.nodes[nb + lane] : 0u;
    float4 q4 = make_float4(0.0f, 0.0f, 0.0f, 1.0f), v4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (have) { q4 = q[node]; v4 = vel[node]; }
    const uint32_t eb = cv.entStart[c], ee = cv.entStart[c + 1];
    for (uint32_t base = eb; base < ee; base += 32) {
      const uint32_t myW = base + lane < ee ? cv.lanes[base + lane] : 0u;
      const uint32_t myW2 = base + lane < ee ? (uint32_t)(cv.keys[base + lane] >> 32) : 0u;
      const int cnt = (int)min(32u, ee - base);
      for (int i = 0; i < cnt; ++i) {
        const uint32_t lw = __shfl_sync(0xffffffffu, myW, i), lw2 = __shfl_sync(0xffffffffu, myW2, i);
        const int la = lw & 31u, lb = (lw >> 16) & 31u, lc = lw2 & 31u, ld = (lw2 >> 16) & 31u;
        // point-triangle friction / restitution (Solver.cpp:431-471), as FrictionOp
        V3 B = shflV3(q4, lb), C = shflV3(q4, lc), D = shflV3(q4, ld);
        float wa = __shfl_sync(0xffffffffu, q4.w, la), wb = __shfl_sync(0xffffffffu, q4.w, lb),
              wc = __shfl_sync(0xffffffffu, q4.w, lc), wd = __shfl_sync(0xffffffffu, q4.w, ld);
        V3 va = shflV3(v4, la), vb = shflV3(v4, lb), vc = shflV3(v4, lc), vd = shflV3(v4, ld);
        V3 avgTri = (vb + vc + vd) / 3.0f;
        V3 nrm = normalize(cross(C - B, D - B));
        V3 rel = va - avgTri;
        float vDotN = dot(rel, nrm);
        V3 perp = rel - vDotN * nrm;
        float fr = friction;
        if (length(perp) < staticThreshold) fr = 1.0f;
        float triW = wb + wc + wd;
        float wSum = wa + triW;
        V3 dv = (-fr) * perp - (1.1f * fminf(vDotN, 0.0f)) * nrm;
        V3 dtv = (-dv) * triW / wSum;
        if (lane == la) {
          V3 nv = va + dv * wa / wSum;
          v4.x = nv.x; v4.y = nv.y; v4.z = nv.z;
        } else if (lane == lb || lane == lc || lane == ld) {
          v4.x += dtv.x; v4.y += dtv.y; v4.z += dtv.z;
        }
      }
    }
    if (have) vel[node] = make_float4(v4.x, v4.y, v4.z, 0.0f);
  }
}

// ---- ordered sweeps of mid-size contact clusters: one warp, nodes staged in shared memory ----------------------
// A cluster of 33 .. kMidClusterMax nodes (a column of stacked bodies, say) has little parallelism inside: its
// entries chain through shared nodes, and the dataflow executor pays an L2 round trip per link of that chain.
// Here one warp copies the cluster's nodes into shared memory and walks the entries in list order, every lane
// evaluating the same entry from broadcast reads (lanes 0..3 write the four nodes back): the reference's
// sequential order at shared-memory latency, all sweeps and their floor snaps in one launch.  Clusters are
// independent, so which CTA takes which cluster does not matter.
constexpr int kMidWarpsPerSm = 6;

__global__ void __launch_bounds__(32) k_gs_mid_stabilize(ClusterView cv, const uint32_t* __restrict__ nMidPtr,
                                                         float4* __restrict__ q, float4* __restrict__ prev,
                                                         const float4* __restrict__ snap,
                                                         const uint32_t* __restrict__ floorMult, int haveFloor,
                                                         float thickness, uint32_t sweeps) {
  __shared__ float4 sq[kMidClusterMax], sp[kMidClusterMax];
  const int lane = threadIdx.x;
  const uint32_t nMid = *nMidPtr;
  for (uint32_t mi = blockIdx.x; mi < nMid; mi += gridDim.x) {
    const uint32_t c = cv.midList[mi];
    const uint32_t nb = cv.start[c], size = cv.start[c + 1] - nb;
    for (uint32_t k = lane; k < size; k += 32) {
      const uint32_t node = cv.nodes[nb + k];
      sq[k] = q[node]; sp[k] = prev[node];
    }
    __syncwarp();
    const uint32_t eb = cv.entStart[c], ee = cv.entStart[c + 1];
    for (uint32_t sweep = 0; sweep < sweeps; ++sweep) {
      for (uint32_t base = eb; base < ee; base += 32) {
        const uint32_t myW = base + lane < ee ? cv.lanes[base + lane] : 0u;
        const uint32_t myW2 = base + lane < ee ? (uint32_t)(cv.keys[base + lane] >> 32) : 0u;
        const int cnt = (int)min(32u, ee - base);
        for (int i = 0; i < cnt; ++i) {
          const uint32_t lw = __shfl_sync(0xffffffffu, myW, i), lw2 = __shfl_sync(0xffffffffu, myW2, i);
          const uint32_t ra = lw & 0xffffu, rb = lw >> 16, rc = lw2 & 0xffffu, rd = lw2 >> 16;
          // PointTriangleCollisionConstraint::stabilizeCollisions (CollisionConstraint.cpp:126-162), as StabilizeOp
          const float4 a4 = sq[ra], b4 = sq[rb], c4 = sq[rc], d4 = sq[rd];
          V3 A = v3(a4), B = v3(b4), C = v3(c4), D = v3(d4);
          V3 nrm = normalize(cross(C - B, D - B));
          float nDotP = dot(nrm, A - B);
          if (!(nDotP < thickness)) continue;
          V3 disp = (thickness - nDotP) * nrm;
          float wTri = b4.w + c4.w + d4.w;
          float wSum = a4.w + wTri;
          V3 da = disp * a4.w / wSum, dt = disp * wTri / wSum;
          if (lane < 4) {
            const uint32_t r = lane == 0 ? ra : lane == 1 ? rb : lane == 2 ? rc : rd;
            const float4 cur = lane == 0 ? a4 : lane == 1 ? b4 : lane == 2 ? c4 : d4;
            const V3 d = lane == 0 ? da : -dt;
            const float4 pv = sp[r];
            sq[r] = make_float4(cur.x + d.x, cur.y + d.y, cur.z + d.z, cur.w);
            sp[r] = make_float4(pv.x + d.x, pv.y + d.y, pv.z + d.z, pv.w);
          }
          __syncwarp();
        }
      }
      if (haveFloor) {  // Solver.cpp:379-382, after the sweep's point-triangle entries
        for (uint32_t k = lane; k < size; k += 32) {
          const uint32_t node = cv.nodes[nb + k];
          if (floorMult[node] != 0u) { const float4 s4 = snap[node]; sq[k] = make_float4(s4.x, s4.y, s4.z, sq[k].w); }
        }
        __syncwarp();
      }
    }
    for (uint32_t k = lane; k < size; k += 32) {
      const uint32_t node = cv.nodes[nb + k];
      q[node] = sq[k]; prev[node] = sp[k];
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32) k_gs_mid_friction(ClusterView cv, const uint32_t* __restrict__ nMidPtr,
                                                        const float4* __restrict__ q, float4* __restrict__ vel,
                                                        float friction, float staticThreshold) {
  __shared__ float4 sq[kMidClusterMax], sv[kMidClusterMax];
  const int lane = threadIdx.x;
  const uint32_t nMid = *nMidPtr;
  for (uint32_t mi = blockIdx.x; mi < nMid; mi += gridDim.x) {
    const uint32_t c = cv.midList[mi];
    const uint32_t nb = cv.start[c], size = cv.start[c + 1] - nb;
    for (uint32_t k = lane; k < size; k += 32) {
      const uint32_t node = cv.nodes[nb + k];
      sq[k] = q[node]; sv[k] = vel[node];
    }
    __syncwarp();
    const uint32_t eb = cv.entStart[c], ee = cv.entStart[c + 1];
    for (uint32_t base = eb; base < ee; base += 32) {
      const uint32_t myW = base + lane < ee ? cv.lanes[base + lane] : 0u;
      const uint32_t myW2 = base + lane < ee ? (uint32_t)(cv.keys[base + lane] >> 32) : 0u;
      const int cnt = (int)min(32u, ee - base);
      for (int i = 0; i < cnt; ++i) {
        const uint32_t lw = __shfl_sync(0xffffffffu, myW, i), lw2 = __shfl_sync(0xffffffffu, myW2, i);
        const uint32_t ra = lw & 0xffffu, rb = lw >> 16, rc = lw2 & 0xffffu, rd = lw2 >> 16;
        // point-triangle friction / restitution (Solver.cpp:431-471), as FrictionOp
        const float4 a4 = sq[ra], b4 = sq[rb], c4 = sq[rc], d4 = sq[rd];
        const V3 va = v3(sv[ra]), vb = v3(sv[rb]), vc = v3(sv[rc]), vd = v3(sv[rd]);
        V3 avgTri = (vb + vc + vd) / 3.0f;
        V3 nrm = normalize(cross(v3(c4) - v3(b4), v3(d4) - v3(b4)));
        V3 rel = va - avgTri;
        float vDotN = dot(rel, nrm);
        V3 perp = rel - vDotN * nrm;
        float fr = friction;
        if (length(perp) < staticThreshold) fr = 1.0f;
        float triW = b4.w + c4.w + d4.w;
        float wSum = a4.w + triW;
        V3 dv = (-fr) * perp - (1.1f * fminf(vDotN, 0.0f)) * nrm;
        V3 dtv = (-dv) * triW / wSum;
        if (lane < 4) {
          const uint32_t r = lane == 0 ? ra : lane == 1 ? rb : lane == 2 ? rc : rd;
          const V3 v0 = lane == 0 ? va : lane == 1 ? vb : lane == 2 ? vc : vd;
          const V3 nv = lane == 0 ? va + dv * a4.w / wSum : v0 + dtv;
          sv[r] = f4(nv, 0.0f);
        }
        __syncwarp();
      }
    }
    for (uint32_t k = lane; k < size; k += 32) vel[cv.nodes[nb + k]] = make_float4(sv[k].x, sv[k].y, sv[k].z, 0.0f);
    __syncwarp();
  }
}

// nodes on the floor go back to their projected position (Solver.cpp:379-382)
__global__ void __launch_bounds__(kThreads) k_floor_snap(uint32_t nFloor, const uint32_t* __restrict__ nodes,
                                                         const float4* __restrict__ snap, float4* __restrict__ q,
                                                         const uint8_t* __restrict__ gsClass, int cls) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nFloor) return;
  uint32_t node = nodes[i];
  if (gsClass && gsClass[node] != cls) return;  // nodes of contact clusters are snapped inside their sweeps
  float4 p = snap[node];
  float4 cur = q[node];
  q[node] = make_float4(p.x, p.y, p.z, cur.w);
}

// floor friction, once per duplicate in list order (Solver.cpp:473-484); duplicates only touch their own node
__global__ void __launch_bounds__(kThreads) k_floor_friction(uint32_t n, const uint32_t* __restrict__ mult,
                                                             float4* __restrict__ vel, float friction, float staticThreshold) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t m = mult[i];
  if (!m) return;
  float4 v = vel[i];
  for (uint32_t k = 0; k < m; ++k) {
    float fr = friction;
    if (sqrtf(v.x * v.x + 0.0f + v.z * v.z) < staticThreshold) fr = 1.0f;
    v.x += (-fr) * v.x; v.y += (-fr) * 0.0f; v.z += (-fr) * v.z;
  }
  vel[i] = v;
}

constexpr int kSweepCounters = 64;

int prepareContactSweeps(ContactWork& w, cudaStream_t s, const ContactLists& c) {
  w.nTri = c.nTri;
  w.sweepsUsed = 0;
  w.haveClusters = false;
  if (!c.nTri) return 0;
  if (w.sweepCounters.reserve(kSweepCounters) != cudaSuccess) return -1;
  if (cudaMemsetAsync(w.sweepCounters.p, 0, kSweepCounters * sizeof(uint32_t), s) != cudaSuccess) return -1;
  return 0;
}

// After reblock.cu has formed this substep's contact clusters: entries grouped by cluster (stable, so list
// order survives inside a cluster) with the lanes of their four nodes.
int prepareClusterSweeps(ContactWork& w, cudaStream_t s, const ContactLists& c, const ClusterTables& t) {
  if (!c.nTri || !t.nTouched) return 0;
  const uint32_t nTri = c.nTri, clusterBound = t.nTouched / 4 + 2;
  if (w.keys.reserve(nTri) != cudaSuccess || w.tmpKeys.reserve(nTri) != cudaSuccess || w.lanes.reserve(nTri) != cudaSuccess ||
      w.tmpVals.reserve(nTri) != cudaSuccess || w.entStart.reserve(clusterBound + 2) != cudaSuccess ||
      w.sortHist.reserve(sortHistBytes(nTri) / 4 + 4) != cudaSuccess ||
      w.scanScratch.reserve(scanScratchElems(std::max<uint64_t>(clusterBound + 2, 1024))) != cudaSuccess)
    return -1;
  int L = 0;
  cudaMemsetAsync(w.entStart.p, 0, (size_t)(clusterBound + 2) * sizeof(uint32_t), s);
  k_entry_keys<<<gridFor(nTri, kThreads), kThreads, 0, s>>>(nTri, c.tri, t.clusterOf, t.rankOf, w.keys.p, w.lanes.p, w.entStart.p); ++L;
  L += launchExclusiveScan(s, w.entStart.p, clusterBound + 1, w.scanScratch.p);
  int bits = 1;
  while ((1ull << bits) <= (uint64_t)clusterBound) ++bits;
  L += launchSortPairs(s, nTri, w.keys.p, w.lanes.p, w.tmpKeys.p, w.tmpVals.p, w.sortHist.p, bits);
  w.view = ClusterView{t.nClusters, t.start, t.nodes, w.entStart.p, w.lanes.p, w.keys.p, t.midList};
  w.gsClass = t.gsClass;
  w.hostCounts = t.hostCounts; w.countsReady = t.countsReady; w.nMidDev = t.nMidDev;
  w.clusterBound = clusterBound;
  w.haveClusters = true;
  return L;
}

template <typename Op>
static int launchSweep(cudaStream_t s, ContactWork& w, const ContactLists& c, uint32_t n, Op op) {
  if (w.sweepsUsed == kSweepCounters) {  // more sweeps than counters in one substep: recycle
    cudaMemsetAsync(w.sweepCounters.p, 0, kSweepCounters * sizeof(uint32_t), s);
    w.sweepsUsed = 0;
  }
  cudaMemsetAsync(c.nodeDone, 0, (size_t)n * sizeof(uint32_t), s);
  uint32_t chunks = (c.nTri + 31u) / 32u;
  int grid = (int)std::min<uint32_t>(kNumSMs * 8, (chunks + kGsThreads / 32 - 1) / (kGsThreads / 32));
  k_gs_dataflow<Op><<<grid, kGsThreads, 0, s>>>(c.tri, c.ticket, c.nTri, w.gsClass, c.nodeDone,
                                                 w.sweepCounters.p + w.sweepsUsed, op);
  ++w.sweepsUsed;
  return 1;
}

static int clusterGrid(uint32_t clusterBound) {
  return (int)std::min<uint32_t>(kNumSMs * 8, (clusterBound + kThreads / 32 - 1) / (kThreads / 32));
}

int launchStabilize(cudaStream_t s, ContactWork& w, const ContactLists& c, uint32_t n, float4* q, float4* prev,
                    const float4* snap, float thickness, uint32_t iterations) {
  int L = 0;
  if (!iterations) return 0;
  const uint8_t* cls = (c.nTri && w.haveClusters) ? w.gsClass : nullptr;
  if (cls) {
    // small clusters: all sweeps (and their floor snaps) in one launch
    k_gs_cluster_stabilize<<<clusterGrid(w.clusterBound), kThreads, 0, s>>>(w.view, q, prev, snap, c.floorMult, c.nFloor ? 1 : 0,
                                                                           thickness, iterations); ++L;
    // how many mid / large clusters this substep has was copied to the host right after they were formed
    if (w.countsReady) cudaEventSynchronize(w.countsReady);
    const uint32_t nMid = w.hostCounts ? w.hostCounts[0] : 0u, nLarge = w.hostCounts ? w.hostCounts[1] : 1u;
    if (nMid) {
      k_gs_mid_stabilize<<<(int)std::min<uint32_t>(nMid, kNumSMs * kMidWarpsPerSm), 32, 0, s>>>(w.view, w.nMidDev, q, prev, snap, c.floorMult,
                                                                                         c.nFloor ? 1 : 0, thickness, iterations); ++L;
    }
    // large clusters: dataflow sweeps, floor snap of their nodes after each one
    for (uint32_t it = 0; nLarge && it < iterations; ++it) {
      L += launchSweep(s, w, c, n, StabilizeOp{q, prev, thickness});
      if (c.nFloor) { k_floor_snap<<<gridFor(c.nFloor, kThreads), kThreads, 0, s>>>(c.nFloor, c.floorNode, snap, q, cls, 2); ++L; }
    }
  }
  // floor nodes outside every cluster: the snap is idempotent, once is the same as once per sweep
  if (c.nFloor) { k_floor_snap<<<gridFor(c.nFloor, kThreads), kThreads, 0, s>>>(c.nFloor, c.floorNode, snap, q, cls, 0); ++L; }
  return L;
}

int launchFriction(cudaStream_t s, ContactWork& w, const ContactLists& c, uint32_t n, const float4* q, float4* vel,
                   float friction, float staticThreshold) {
  int L = 0;
  if (c.nTri && w.haveClusters) {
    k_gs_cluster_friction<<<clusterGrid(w.clusterBound), kThreads, 0, s>>>(w.view, q, vel, friction, staticThreshold); ++L;
    if (w.countsReady) cudaEventSynchronize(w.countsReady);
    const uint32_t nMid = w.hostCounts ? w.hostCounts[0] : 0u, nLarge = w.hostCounts ? w.hostCounts[1] : 1u;
    if (nMid) {
      k_gs_mid_friction<<<(int)std::min<uint32_t>(nMid, kNumSMs * kMidWarpsPerSm), 32, 0, s>>>(w.view, w.nMidDev, q, vel, friction,
                                                                                        staticThreshold); ++L;
    }
    if (nLarge) L += launchSweep(s, w, c, n, FrictionOp{q, vel, friction, staticThreshold});
  }
  if (c.nFloor) { k_floor_friction<<<gridFor(n, kThreads), kThreads, 0, s>>>(n, c.floorMult, vel, friction, staticThreshold); ++L; }
  return L;
}

}  // namespace pies
