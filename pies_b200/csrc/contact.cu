// contact.cu — collision response of the PD path: per-iteration projections of the
// point-triangle / floor constraints, collision stabilisation and the friction pass.
//
// Reference: Src/CollisionConstraint.cpp:86-194 (point-triangle), :439-463 (floor),
// Src/Solver.cpp:298-308,337-349 (local step / RHS), :367-383 (stabilisation), :431-484 (friction).
//
// Stabilisation and friction are sequential Gauss-Seidel sweeps over the collision list in the
// reference, and the result depends on that order.  They are executed here in
// exactly that order, in parallel where the order allows it:
//   * the contact graph (nodes linked by collision entries) is split into connected components
//     with a min-id union-find; components never interact, so each is swept by one CTA;
//   * inside a component the CTA walks the entries in list order, a window at a time; in every
//     round the entries whose four nodes are not claimed by an earlier pending entry run
//     together (claims are taken with shared-memory atomicMin on the entry's position, so the
//     earliest pending entry touching a node always wins).  This is the sequential dependency
//     order, so results differ from the reference only by rounding.
#include "contact.h"

#include "common.cuh"

namespace pies {

static inline int gridFor(uint64_t n, int threads) { return (int)((n + threads - 1) / threads); }

constexpr float kPtW = 10000.0f;     // PointTriangleCollisionConstraint::w (CollisionConstraint.h:32)
constexpr float kFloorW = 10000.0f;  // StaticCollisionConstraint::w (CollisionConstraint.h:78)

// ---- per-iteration projections (parallel, no ordering issue) ---------------------------------------
// contribC[4e + slot] = w * (A^T A p)_slot with p = current positions, p_A pushed to `thickness`
// above the triangle plane when below it (CollisionConstraint.cpp:86-124, :176-194).
__global__ void __launch_bounds__(kThreads) k_pt_project(uint32_t nTri, const uint4* __restrict__ entries,
                                                         const float4* __restrict__ q, float thickness,
                                                         float4* __restrict__ contribC) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nTri) return;
  uint4 id = entries[e];
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
  if (c.nTri) { k_pt_project<<<gridFor(c.nTri, kThreads), kThreads, 0, s>>>(c.nTri, c.tri, q, thickness, contribC); ++L; }
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

// ---- contact-graph components (min-id union-find) --------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_uf_init(uint32_t n, uint32_t* __restrict__ parent) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = i;
}

__device__ __forceinline__ uint32_t ufFind(uint32_t* parent, uint32_t x) {
  uint32_t p = *(volatile uint32_t*)(parent + x);
  while (p != x) { x = p; p = *(volatile uint32_t*)(parent + x); }
  return x;
}

__device__ __forceinline__ void ufUnite(uint32_t* parent, uint32_t u, uint32_t v) {
  while (true) {
    u = ufFind(parent, u);
    v = ufFind(parent, v);
    if (u == v) return;
    if (u < v) { uint32_t t = u; u = v; v = t; }  // hook the larger root under the smaller one
    uint32_t old = atomicMin(parent + u, v);
    if (old == u) return;
    u = old;  // someone else re-parented u meanwhile: merge what it points to with v
  }
}

__global__ void __launch_bounds__(kThreads) k_uf_unite(uint32_t nTri, const uint4* __restrict__ entries,
                                                       uint32_t* __restrict__ parent) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nTri) return;
  uint4 id = entries[e];
  ufUnite(parent, id.x, id.y);
  ufUnite(parent, id.x, id.z);
  ufUnite(parent, id.x, id.w);
}

__global__ void __launch_bounds__(kThreads) k_uf_keys(uint32_t nTri, const uint4* __restrict__ entries,
                                                      uint32_t* __restrict__ parent, uint64_t* __restrict__ keys,
                                                      uint32_t* __restrict__ vals) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nTri) return;
  keys[e] = ufFind(parent, entries[e].x);  // min node id of the component: deterministic
  vals[e] = e;
}

__global__ void __launch_bounds__(kThreads) k_comp_heads(uint32_t nTri, const uint64_t* __restrict__ keys,
                                                         uint32_t* __restrict__ heads) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > nTri) return;
  heads[j] = (j < nTri && (j == 0 || keys[j] != keys[j - 1])) ? 1u : 0u;
}

__global__ void __launch_bounds__(kThreads) k_comp_starts(uint32_t nTri, const uint64_t* __restrict__ keys,
                                                          const uint32_t* __restrict__ headScan,
                                                          uint32_t* __restrict__ compStart, uint32_t* __restrict__ nComp) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nTri) return;
  bool head = (j == 0 || keys[j] != keys[j - 1]);
  uint32_t idx = headScan[j] + (head ? 1u : 0u) - 1u;
  if (head) compStart[idx] = j;
  if (j == nTri - 1) { compStart[idx + 1] = nTri; *nComp = idx + 1; }
}

int buildContactComponents(ContactWork& w, cudaStream_t s, uint32_t nNodes, const ContactLists& c) {
  int L = 0;
  w.nTri = c.nTri;
  if (!c.nTri) return 0;
  if (w.parent.reserve(nNodes + 1) != cudaSuccess || w.keys.reserve(c.nTri) != cudaSuccess ||
      w.tmpKeys.reserve(c.nTri) != cudaSuccess || w.perm.reserve(c.nTri) != cudaSuccess ||
      w.tmpVals.reserve(c.nTri) != cudaSuccess || w.heads.reserve(c.nTri + 2) != cudaSuccess ||
      w.compStart.reserve(c.nTri + 2) != cudaSuccess || w.nComp.reserve(4) != cudaSuccess ||
      w.sortHist.reserve(sortHistBytes(c.nTri) / 4 + 4) != cudaSuccess ||
      w.scanScratch.reserve(scanScratchElems(c.nTri + 2)) != cudaSuccess)
    return -1;
  k_uf_init<<<gridFor(nNodes, kThreads), kThreads, 0, s>>>(nNodes, w.parent.p); ++L;
  k_uf_unite<<<gridFor(c.nTri, kThreads), kThreads, 0, s>>>(c.nTri, c.tri, w.parent.p); ++L;
  k_uf_keys<<<gridFor(c.nTri, kThreads), kThreads, 0, s>>>(c.nTri, c.tri, w.parent.p, w.keys.p, w.perm.p); ++L;
  int bits = 1;
  while ((1ull << bits) < (uint64_t)nNodes) ++bits;
  L += launchSortPairs(s, c.nTri, w.keys.p, w.perm.p, w.tmpKeys.p, w.tmpVals.p, w.sortHist.p, bits);
  k_comp_heads<<<gridFor(c.nTri + 1, kThreads), kThreads, 0, s>>>(c.nTri, w.keys.p, w.heads.p); ++L;
  L += launchExclusiveScan(s, w.heads.p, c.nTri + 1, w.scanScratch.p);
  k_comp_starts<<<gridFor(c.nTri, kThreads), kThreads, 0, s>>>(c.nTri, w.keys.p, w.heads.p, w.compStart.p, w.nComp.p); ++L;
  return L;
}

// ---- ordered Gauss-Seidel executor --------------------------------------------------------------------
constexpr int kGsThreads = 256;
constexpr int kGsTable = 2048;

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
    q[id.x] = f4(A + da, a4.w); q[id.y] = f4(B - dt, b4.w); q[id.z] = f4(C - dt, c4.w); q[id.w] = f4(D - dt, d4.w);
    float4 pa = __ldcg(prev + id.x), pb = __ldcg(prev + id.y), pc = __ldcg(prev + id.z), pd = __ldcg(prev + id.w);
    prev[id.x] = f4(v3(pa) + da, pa.w); prev[id.y] = f4(v3(pb) - dt, pb.w);
    prev[id.z] = f4(v3(pc) - dt, pc.w); prev[id.w] = f4(v3(pd) - dt, pd.w);
  }
};

struct FrictionOp {
  const float4* q;
  float4* vel;
  float friction, staticThreshold;
  // point-triangle friction / restitution (Solver.cpp:431-471)
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
    vel[id.x] = f4(va + dv * a4.w / wSum, 0.0f);
    vel[id.y] = f4(vb + dtv, 0.0f); vel[id.z] = f4(vc + dtv, 0.0f); vel[id.w] = f4(vd + dtv, 0.0f);
  }
};

template <typename Op>
__global__ void __launch_bounds__(kGsThreads) k_gs_ordered(const uint4* __restrict__ entries,
                                                           const uint32_t* __restrict__ perm,
                                                           const uint32_t* __restrict__ compStart,
                                                           const uint32_t* __restrict__ nCompPtr, Op op) {
  __shared__ int owner[kGsTable];
  for (int i = threadIdx.x; i < kGsTable; i += kGsThreads) owner[i] = 0x7fffffff;
  __syncthreads();
  const uint32_t nComp = *nCompPtr;
  for (uint32_t comp = blockIdx.x; comp < nComp; comp += gridDim.x) {
    uint32_t s = compStart[comp], e = compStart[comp + 1];
    for (uint32_t base = s; base < e; base += kGsThreads) {
      uint32_t idx = base + threadIdx.x;
      bool pending = idx < e;
      uint4 id = make_uint4(0, 0, 0, 0);
      if (pending) id = entries[perm[idx]];
      uint32_t h0 = (id.x * 2654435761u) >> 21, h1 = (id.y * 2654435761u) >> 21, h2 = (id.z * 2654435761u) >> 21,
               h3 = (id.w * 2654435761u) >> 21;  // 11 bits -> kGsTable
      while (__syncthreads_or(pending)) {
        int me = (int)threadIdx.x;
        if (pending) { atomicMin(&owner[h0], me); atomicMin(&owner[h1], me); atomicMin(&owner[h2], me); atomicMin(&owner[h3], me); }
        __syncthreads();
        bool ready = pending && owner[h0] == me && owner[h1] == me && owner[h2] == me && owner[h3] == me;
        __syncthreads();
        if (pending) { owner[h0] = 0x7fffffff; owner[h1] = 0x7fffffff; owner[h2] = 0x7fffffff; owner[h3] = 0x7fffffff; }
        if (ready) { op(id); pending = false; }
        __threadfence_block();
      }
    }
  }
}

// nodes on the floor go back to their projected position (Solver.cpp:379-382)
__global__ void __launch_bounds__(kThreads) k_floor_snap(uint32_t nFloor, const uint32_t* __restrict__ nodes,
                                                         const float4* __restrict__ snap, float4* __restrict__ q) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nFloor) return;
  uint32_t node = nodes[i];
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

static int gsGrid(uint32_t nTri) { return (int)std::min<uint32_t>(kNumSMs * 8, (nTri + kGsThreads - 1) / kGsThreads * 4 + 1); }

int launchStabilize(cudaStream_t s, const ContactWork& w, const ContactLists& c, float4* q, float4* prev, const float4* snap,
                    float thickness, uint32_t iterations) {
  int L = 0;
  for (uint32_t it = 0; it < iterations; ++it) {
    if (c.nTri) {
      k_gs_ordered<StabilizeOp><<<gsGrid(c.nTri), kGsThreads, 0, s>>>(c.tri, w.perm.p, w.compStart.p, w.nComp.p,
                                                                     StabilizeOp{q, prev, thickness});
      ++L;
    }
    if (c.nFloor) { k_floor_snap<<<gridFor(c.nFloor, kThreads), kThreads, 0, s>>>(c.nFloor, c.floorNode, snap, q); ++L; }
  }
  return L;
}

int launchFriction(cudaStream_t s, const ContactWork& w, const ContactLists& c, uint32_t n, const float4* q, float4* vel,
                   float friction, float staticThreshold) {
  int L = 0;
  if (c.nTri) {
    k_gs_ordered<FrictionOp><<<gsGrid(c.nTri), kGsThreads, 0, s>>>(c.tri, w.perm.p, w.compStart.p, w.nComp.p,
                                                                  FrictionOp{q, vel, friction, staticThreshold});
    ++L;
  }
  if (c.nFloor) { k_floor_friction<<<gridFor(n, kThreads), kThreads, 0, s>>>(n, c.floorMult, vel, friction, staticThreshold); ++L; }
  return L;
}

}  // namespace pies
