// pcg.cu — global step of Projective Dynamics as a block-Jacobi preconditioned CG.
//
// Replaces the reference's per-substep sparse Cholesky re-factorisation and solve
// (reference Src/Solver.cpp:242-262 SimplicialLLT(S + C_t), :356 solve) with a
// matrix that is never re-assembled: A = S (CSR, built once per topology) + C_t applied
// matrix-free from the substep's collision lists (point-triangle blocks
// w [3 -1 -1 -1; -1 1 0 0; -1 0 1 0; -1 0 0 1], CollisionConstraint.cpp:74-83, and the
// floor diagonal, :442-445).  The three coordinate columns are solved together with
// separate CG scalars per column, warm-started at the current positions.
//
// Determinism: every dot product is reduced in a fixed order (per-CTA partials on a
// fixed grid, summed redundantly by each consumer CTA in the same order) — no atomics.
#include "kernels.h"
#include "common.cuh"

namespace pies {

constexpr int kPartialStride = 16;
// partial slots
constexpr int kPA = 0;    // p.Ap            (3)   written by spmv
constexpr int kBB = 3;    // b.b             (3)   written by residual
constexpr int kRZ0 = 6;   // r.z, even parity (3)
constexpr int kRZ1 = 9;   // r.z, odd parity  (3)
constexpr int kRR = 12;   // r.r             (3)


// Sum `partials[b*stride + slot + c]` over the fixed producer grid, same order in every CTA.
__device__ __forceinline__ void reducePartials3(const float* __restrict__ partials, int slot, float (&out)[3],
                                                float* smem /* >= 96 + 3 floats */) {
  float v[3] = {0.0f, 0.0f, 0.0f};
  for (int b = threadIdx.x; b < kReduceBlocks; b += blockDim.x) {
    const float* p = partials + b * kPartialStride + slot;
    v[0] += p[0]; v[1] += p[1]; v[2] += p[2];
  }
  blockSum<3>(v, smem);
  if (threadIdx.x == 0) { smem[96] = v[0]; smem[97] = v[1]; smem[98] = v[2]; }
  __syncthreads();
  out[0] = smem[96]; out[1] = smem[97]; out[2] = smem[98];
  __syncthreads();
}

// CTA-uniform early exit once the solve has converged (lets the host enqueue a fixed
// number of iterations without a sync; finished iterations cost one flag read).
__device__ __forceinline__ bool ctaConverged(const int* flag) {
  __shared__ int sflag;
  if (threadIdx.x == 0) sflag = *(volatile const int*)flag;
  __syncthreads();
  return sflag != 0;
}

// y_i = (A x)_i for one row: CSR part + matrix-free collision part.
__device__ __forceinline__ V3 applyRow(uint32_t i, const CsrMatrix& A, const ContactLists& c,
                                       const float4* __restrict__ x, V3 xi) {
  V3 y = v3(0.0f, 0.0f, 0.0f);
  int beg = A.rowPtr[i], end = A.rowPtr[i + 1];
  for (int k = beg; k < end; ++k) {
    float a = __ldg(A.val + k);
    float4 xv = __ldg(x + __ldg(A.col + k));
    y.x += a * xv.x; y.y += a * xv.y; y.z += a * xv.z;
  }
  if (c.nFloor) {
    float fw = c.floorW[i];
    y.x += fw * xi.x; y.y += fw * xi.y; y.z += fw * xi.z;
  }
  if (c.nTri) {
    int cb = c.incPtr[i], ce = c.incPtr[i + 1];
    for (int k = cb; k < ce; ++k) {
      uint32_t v = c.inc[k];
      uint4 e = __ldg(c.uTri + (v >> 2));
      float wgt = __ldg(c.uW + (v >> 2));
      uint32_t slot = v & 3u;
      V3 t;
      if (slot == 0) {
        V3 xb = v3(__ldg(x + e.y)), xc = v3(__ldg(x + e.z)), xd = v3(__ldg(x + e.w));
        t = 3.0f * xi - xb - xc - xd;
      } else {
        t = xi - v3(__ldg(x + e.x));
      }
      y += wgt * t;
    }
  }
  return y;
}

// Same row product accumulated in fp64 (float products are exact in double): used once per
// solve for the start residual, where b and A x agree to ~7 digits and the difference is what
// matters (velocities are position differences divided by h, so sub-ulp position errors count).
__device__ __forceinline__ void applyRowD(uint32_t i, const CsrMatrix& A, const ContactLists& c,
                                          const float4* __restrict__ x, V3 xi, double (&y)[3]) {
  y[0] = y[1] = y[2] = 0.0;
  int beg = A.rowPtr[i], end = A.rowPtr[i + 1];
  for (int k = beg; k < end; ++k) {
    double a = (double)__ldg(A.val + k);
    float4 xv = __ldg(x + __ldg(A.col + k));
    y[0] += a * (double)xv.x; y[1] += a * (double)xv.y; y[2] += a * (double)xv.z;
  }
  if (c.nFloor) {
    double fw = (double)c.floorW[i];
    y[0] += fw * (double)xi.x; y[1] += fw * (double)xi.y; y[2] += fw * (double)xi.z;
  }
  if (c.nTri) {
    int cb = c.incPtr[i], ce = c.incPtr[i + 1];
    for (int k = cb; k < ce; ++k) {
      uint32_t v = c.inc[k];
      uint4 e = __ldg(c.uTri + (v >> 2));
      double wgt = (double)__ldg(c.uW + (v >> 2));
      uint32_t slot = v & 3u;
      double t[3];
      if (slot == 0) {
        float4 xb = __ldg(x + e.y), xc = __ldg(x + e.z), xd = __ldg(x + e.w);
        t[0] = 3.0 * (double)xi.x - (double)xb.x - (double)xc.x - (double)xd.x;
        t[1] = 3.0 * (double)xi.y - (double)xb.y - (double)xc.y - (double)xd.y;
        t[2] = 3.0 * (double)xi.z - (double)xb.z - (double)xc.z - (double)xd.z;
      } else {
        float4 xa = __ldg(x + e.x);
        t[0] = (double)xi.x - (double)xa.x; t[1] = (double)xi.y - (double)xa.y; t[2] = (double)xi.z - (double)xa.z;
      }
      y[0] += wgt * t[0]; y[1] += wgt * t[1]; y[2] += wgt * t[2];
    }
  }
}

// r = b - A x (fp64 accumulation) ; delta = 0 ; partial b.b
__global__ void __launch_bounds__(kThreads) k_pcg_residual(CsrMatrix A, ContactLists c, const float4* __restrict__ b,
                                                           const float4* __restrict__ x, float4* __restrict__ r,
                                                           float4* __restrict__ delta, float* __restrict__ partials,
                                                           int* __restrict__ flag) {
  __shared__ float smem[128];
  float bb[3] = {0.0f, 0.0f, 0.0f};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
    float4 xi4 = x[i];
    double y[3];
    applyRowD(i, A, c, x, v3(xi4), y);
    float4 bi = b[i];
    r[i] = make_float4((float)((double)bi.x - y[0]), (float)((double)bi.y - y[1]), (float)((double)bi.z - y[2]), 0.0f);
    delta[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    bb[0] += bi.x * bi.x; bb[1] += bi.y * bi.y; bb[2] += bi.z * bi.z;
  }
  blockSum<3>(bb, smem);
  if (threadIdx.x == 0) {
    float* p = partials + blockIdx.x * kPartialStride + kBB;
    p[0] = bb[0]; p[1] = bb[1]; p[2] = bb[2];
    if (blockIdx.x == 0) { flag[0] = 0; flag[1] = 0; }
  }
}

// x += delta, once per solve (single rounding of the accumulated correction)
__global__ void __launch_bounds__(kThreads) k_pcg_finish(uint32_t n, float4* __restrict__ x, const float4* __restrict__ delta) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 xv = x[i], d = delta[i];
  x[i] = make_float4(xv.x + d.x, xv.y + d.y, xv.z + d.z, xv.w);
}

// Warp-per-block preconditioner application on registers: z_lane = sum_j Minv[j][lane] r_j.
__device__ __forceinline__ V3 applyBlockInv(const float* __restrict__ inv, V3 r, int lane) {
  V3 z = v3(0.0f, 0.0f, 0.0f);
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    float m = __ldg(inv + j * 32 + lane);
    z.x += m * __shfl_sync(0xffffffffu, r.x, j);
    z.y += m * __shfl_sync(0xffffffffu, r.y, j);
    z.z += m * __shfl_sync(0xffffffffu, r.z, j);
  }
  return z;
}

// z = Minv r ; p = z ; partial r.z (odd parity slot = "iteration -1") and r.r
__global__ void __launch_bounds__(kThreads) k_pcg_start(PcgWork w, float* __restrict__ partials) {
  __shared__ float smem[6 * 32];
  int lane = threadIdx.x & 31;
  uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
  float acc[6] = {0, 0, 0, 0, 0, 0};
  const uint32_t nBlocks = *w.nBlocksDev;
  for (uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < nBlocks; blk += warpsPerGrid) {
    int node = w.blockNodes[blk * 32 + lane];
    V3 r = v3(0.0f, 0.0f, 0.0f);
    if (node >= 0) r = v3(w.r[node]);
    V3 z = applyBlockInv(w.blockInv + (size_t)blk * 1024, r, lane);
    if (node >= 0) {
      w.z[node] = f4(z, 0.0f);
      w.p[node] = f4(z, 0.0f);
      acc[0] += r.x * z.x; acc[1] += r.y * z.y; acc[2] += r.z * z.z;
      acc[3] += r.x * r.x; acc[4] += r.y * r.y; acc[5] += r.z * r.z;
    }
  }
  blockSum<6>(acc, smem);
  if (threadIdx.x == 0) {
    float* p = partials + blockIdx.x * kPartialStride;
    p[kRZ1] = acc[0]; p[kRZ1 + 1] = acc[1]; p[kRZ1 + 2] = acc[2];
    p[kRR] = acc[3]; p[kRR + 1] = acc[4]; p[kRR + 2] = acc[5];
  }
}

// Convergence test on the start residual (so an already-solved system costs no iteration).
__global__ void __launch_bounds__(kThreads) k_pcg_check(const float* __restrict__ partials, float* __restrict__ scalars,
                                                        int* __restrict__ flag, float tol2) {
  __shared__ float smem[128];
  float bb[3], rr[3];
  reducePartials3(partials, kBB, bb, smem);
  reducePartials3(partials, kRR, rr, smem);
  if (threadIdx.x == 0) {
    bool conv = rr[0] <= tol2 * bb[0] && rr[1] <= tol2 * bb[1] && rr[2] <= tol2 * bb[2];
    float rel = 0.0f;
    for (int k = 0; k < 3; ++k) if (bb[k] > 0.0f) rel = fmaxf(rel, rr[k] / bb[k]);
    scalars[0] = sqrtf(rel);
    if (conv) flag[0] = 1;
  }
}

// (A) ap = A p ; partial p.ap
__global__ void __launch_bounds__(kThreads) k_pcg_spmv(CsrMatrix A, ContactLists c, PcgWork w,
                                                       float* __restrict__ partials) {
  __shared__ float smem[128];
  if (ctaConverged(w.flag)) return;
  float pap[3] = {0.0f, 0.0f, 0.0f};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
    V3 pi = v3(w.p[i]);
    V3 y = applyRow(i, A, c, w.p, pi);
    w.ap[i] = f4(y, 0.0f);
    pap[0] += pi.x * y.x; pap[1] += pi.y * y.y; pap[2] += pi.z * y.z;
  }
  blockSum<3>(pap, smem);
  if (threadIdx.x == 0) {
    float* p = partials + blockIdx.x * kPartialStride + kPA;
    p[0] = pap[0]; p[1] = pap[1]; p[2] = pap[2];
  }
}

// (B) alpha = rz / pAp ; delta += alpha p ; r -= alpha ap ; z = Minv r ; partial r.z (parity slot), r.r
__global__ void __launch_bounds__(kThreads) k_pcg_update(PcgWork w, float4* __restrict__ x,
                                                         float* __restrict__ partials, int parity) {
  __shared__ float smem[6 * 32];
  if (ctaConverged(w.flag)) return;
  float rz[3], pap[3], alpha[3];
  reducePartials3(partials, parity ? kRZ0 : kRZ1, rz, smem);  // r.z of the previous step
  reducePartials3(partials, kPA, pap, smem);
#pragma unroll
  for (int k = 0; k < 3; ++k) alpha[k] = pap[k] > 0.0f ? rz[k] / pap[k] : 0.0f;
  int lane = threadIdx.x & 31;
  uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
  float acc[6] = {0, 0, 0, 0, 0, 0};
  const uint32_t nBlocks = *w.nBlocksDev;
  for (uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < nBlocks; blk += warpsPerGrid) {
    int node = w.blockNodes[blk * 32 + lane];
    V3 r = v3(0.0f, 0.0f, 0.0f);
    if (node >= 0) {
      float4 xv = x[node];
      V3 p = v3(w.p[node]), ap = v3(w.ap[node]);
      r = v3(w.r[node]);
      xv.x += alpha[0] * p.x; xv.y += alpha[1] * p.y; xv.z += alpha[2] * p.z;
      r.x -= alpha[0] * ap.x; r.y -= alpha[1] * ap.y; r.z -= alpha[2] * ap.z;
      x[node] = xv;
      w.r[node] = f4(r, 0.0f);
    }
    V3 z = applyBlockInv(w.blockInv + (size_t)blk * 1024, r, lane);
    if (node >= 0) {
      w.z[node] = f4(z, 0.0f);
      acc[0] += r.x * z.x; acc[1] += r.y * z.y; acc[2] += r.z * z.z;
      acc[3] += r.x * r.x; acc[4] += r.y * r.y; acc[5] += r.z * r.z;
    }
  }
  blockSum<6>(acc, smem);
  if (threadIdx.x == 0) {
    float* p = partials + blockIdx.x * kPartialStride;
    int s = parity ? kRZ1 : kRZ0;
    p[s] = acc[0]; p[s + 1] = acc[1]; p[s + 2] = acc[2];
    p[kRR] = acc[3]; p[kRR + 1] = acc[4]; p[kRR + 2] = acc[5];
  }
}

// (C) beta = rz_new / rz_old ; p = z + beta p ; CTA 0 records convergence
__global__ void __launch_bounds__(kThreads) k_pcg_direction(PcgWork w, uint32_t n, float* __restrict__ partials,
                                                            int parity, float tol2) {
  __shared__ float smem[128];
  if (ctaConverged(w.flag)) return;
  float rzNew[3], rzOld[3], rr[3], bb[3], beta[3];
  reducePartials3(partials, parity ? kRZ1 : kRZ0, rzNew, smem);
  reducePartials3(partials, parity ? kRZ0 : kRZ1, rzOld, smem);
  reducePartials3(partials, kRR, rr, smem);
  reducePartials3(partials, kBB, bb, smem);
  bool conv = rr[0] <= tol2 * bb[0] && rr[1] <= tol2 * bb[1] && rr[2] <= tol2 * bb[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) beta[k] = rzOld[k] > 0.0f ? rzNew[k] / rzOld[k] : 0.0f;
  if (!conv) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      float4 z = w.z[i], p = w.p[i];
      w.p[i] = make_float4(z.x + beta[0] * p.x, z.y + beta[1] * p.y, z.z + beta[2] * p.z, 0.0f);
    }
  }
  // Late CTAs may already see the flipped flag and return early: harmless, p is dead once converged.
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float rel = 0.0f;
    for (int k = 0; k < 3; ++k) if (bb[k] > 0.0f) rel = fmaxf(rel, rr[k] / bb[k]);
    w.scalars[0] = sqrtf(rel);
    w.flag[1] += 1;
    if (conv) w.flag[0] = 1;
  }
}

int launchPcgInit(cudaStream_t s, const CsrMatrix& A, const ContactLists& c, const PcgWork& w, const float4* b,
                  const float4* x, float tol) {
  k_pcg_residual<<<kReduceBlocks, kThreads, 0, s>>>(A, c, b, x, w.r, w.delta, w.partials, w.flag);
  k_pcg_start<<<kReduceBlocks, kThreads, 0, s>>>(w, w.partials);
  k_pcg_check<<<1, kThreads, 0, s>>>(w.partials, w.scalars, w.flag, tol * tol);
  return 3;
}

int launchPcgIteration(cudaStream_t s, const CsrMatrix& A, const ContactLists& c, const PcgWork& w, float tol, int parity) {
  k_pcg_spmv<<<kReduceBlocks, kThreads, 0, s>>>(A, c, w, w.partials);
  k_pcg_update<<<kReduceBlocks, kThreads, 0, s>>>(w, w.delta, w.partials, parity);
  k_pcg_direction<<<kReduceBlocks, kThreads, 0, s>>>(w, A.n, w.partials, parity, tol * tol);
  return 3;
}

int launchPcgFinish(cudaStream_t s, const PcgWork& w, uint32_t n, float4* x) {
  k_pcg_finish<<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(n, x, w.delta);
  return 1;
}

}  // namespace pies
