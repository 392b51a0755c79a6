// ccd.cuh — swept point-triangle test and the cell-range functors, with the
// reference's float evaluation order.
//
// The collision lists are compared bit-exactly against the reference on identical
// inputs, and their membership is decided by float comparisons, so this file does not
// use the contracted `a*b+c` the rest of the code allows: every operation goes through
// the round-to-nearest intrinsics (__fmul_rn, __fadd_rn, ...) which nvcc never fuses
// into FMAs, in the association order of the reference's glm expressions compiled for
// baseline x86-64 (no FMA).  Reference: Src/CollisionDetection.cpp:14-302,
// Src/Solver.cpp:639-677, :877-979.
#pragma once

#include "common.cuh"

namespace pies {
namespace ex {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return V3{sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)}; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return V3{add(a.x, b.x), add(a.y, b.y), add(a.z, b.z)}; }
__device__ __forceinline__ V3 scale(V3 a, float s) { return V3{mul(a.x, s), mul(a.y, s), mul(a.z, s)}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }
__device__ __forceinline__ V3 cross3(V3 x, V3 y) {
  return V3{sub(mul(x.y, y.z), mul(y.y, x.z)), sub(mul(x.z, y.x), mul(y.z, x.x)), sub(mul(x.x, y.y), mul(y.x, x.y))};
}
__device__ __forceinline__ V3 norm3(V3 v) { return scale(v, div(1.0f, __fsqrt_rn(dot3(v, v)))); }

// First two components of inverse(mat3(c0, c1, c2)) * v (glm cofactor expansion, column-major).
__device__ __forceinline__ void barycentric(V3 c0, V3 c1, V3 c2, V3 v, float& bx, float& by) {
  float m00 = c0.x, m01 = c0.y, m02 = c0.z, m10 = c1.x, m11 = c1.y, m12 = c1.z, m20 = c2.x, m21 = c2.y, m22 = c2.z;
  float k0 = sub(mul(m11, m22), mul(m21, m12));
  float k1 = sub(mul(m01, m22), mul(m21, m02));
  float k2 = sub(mul(m01, m12), mul(m11, m02));
  float det = add(sub(mul(m00, k0), mul(m10, k1)), mul(m20, k2));
  float ood = div(1.0f, det);
  float i00 = mul(k0, ood);
  float i10 = mul(-sub(mul(m10, m22), mul(m20, m12)), ood);
  float i20 = mul(sub(mul(m10, m21), mul(m20, m11)), ood);
  float i01 = mul(-k1, ood);
  float i11 = mul(sub(mul(m00, m22), mul(m20, m02)), ood);
  float i21 = mul(-sub(mul(m00, m21), mul(m20, m01)), ood);
  bx = add(add(mul(i00, v.x), mul(i10, v.y)), mul(i20, v.z));
  by = add(add(mul(i01, v.x), mul(i11, v.y)), mul(i21, v.z));
}

__device__ __forceinline__ bool outsideTriangle(float bx, float by) {
  return (0.0f > bx) || (bx > 1.0f) || (0.0f > by) || (by > 1.0f) || (add(bx, by) > 1.0f);
}

struct Cubic { float c3, c2, c1, c0; };

// expandTerm (CollisionDetection.cpp:209-221)
__device__ __forceinline__ void expandTerm(float a0, float b0, float c0, float ad, float bd, float cd, Cubic& e) {
  e.c3 = add(e.c3, mul(mul(ad, bd), cd));
  e.c2 = add(e.c2, add(add(mul(mul(ad, bd), c0), mul(mul(a0, bd), cd)), mul(mul(ad, b0), cd)));
  e.c1 = add(e.c1, add(add(mul(mul(ad, b0), c0), mul(mul(a0, bd), c0)), mul(mul(a0, b0), cd)));
  e.c0 = add(e.c0, mul(mul(a0, b0), c0));
}

// Earliest real root in [0,1] (CollisionDetection.cpp:143-205).  The degenerate (linear /
// quadratic) branches follow the reference's float formulas exactly.  The genuine cubic is
// where the reference calls Eigen::PolynomialSolver<float,3> (companion-matrix eigenvalues,
// real if |imag| < 1e-7); that arithmetic lives in Eigen and is replaced here by a closed-form
// solve in fp64 polished by Newton — decisions can differ only for near-tangent roots.
__device__ __forceinline__ bool findRootInInterval(const Cubic& e, float& tOut) {
  if (e.c3 == 0.0f) {
    if (e.c2 == 0.0f) {
      if (e.c1 == 0.0f) {
        if (e.c0 == 0.0f) { tOut = 0.0f; return true; }
        return false;
      }
      float t = div(-e.c0, e.c1);
      if (t >= 0.0f && t <= 1.0f) { tOut = t; return true; }
      return false;
    }
    float disc = sub(mul(e.c1, e.c1), mul(mul(4.0f, e.c2), e.c0));
    if (disc < 0.0f) return false;
    float sq = __fsqrt_rn(disc);
    float t = div(sub(-e.c1, sq), mul(2.0f, e.c2));
    if (t > 1.0f) return false;
    if (t < 0.0f) t = div(add(-e.c1, sq), mul(2.0f, e.c2));
    if (t >= 0.0f && t <= 1.0f) { tOut = t; return true; }
    return false;
  }
  // monic: t^3 + a t^2 + b t + c
  double a = (double)e.c2 / (double)e.c3, b = (double)e.c1 / (double)e.c3, c = (double)e.c0 / (double)e.c3;
  double q = (a * a - 3.0 * b) / 9.0, r = (2.0 * a * a * a - 9.0 * a * b + 27.0 * c) / 54.0;
  double roots[3];
  int nr;
  double r2 = r * r, q3 = q * q * q;
  if (r2 < q3) {
    double th = acos(fmin(fmax(r / sqrt(q3), -1.0), 1.0));
    double sq = -2.0 * sqrt(q);
    roots[0] = sq * cos(th / 3.0) - a / 3.0;
    roots[1] = sq * cos((th + 6.283185307179586) / 3.0) - a / 3.0;
    roots[2] = sq * cos((th - 6.283185307179586) / 3.0) - a / 3.0;
    nr = 3;
  } else {
    double A = -copysign(cbrt(fabs(r) + sqrt(r2 - q3)), r);
    double B = A != 0.0 ? q / A : 0.0;
    roots[0] = (A + B) - a / 3.0;
    nr = 1;
    if (r2 == q3) { roots[1] = -0.5 * (A + B) - a / 3.0; nr = 2; }  // double root is real
  }
  bool found = false;
  float best = 0.0f;
  for (int i = 0; i < nr; ++i) {
    double t = roots[i];
    for (int it = 0; it < 2; ++it) {  // Newton polish on the original coefficients
      double f = (((double)e.c3 * t + (double)e.c2) * t + (double)e.c1) * t + (double)e.c0;
      double df = (3.0 * (double)e.c3 * t + 2.0 * (double)e.c2) * t + (double)e.c1;
      if (df != 0.0) t -= f / df;
    }
    float tf = (float)t;
    if (tf >= 0.0f && tf <= 1.0f && (!found || tf < best)) { best = tf; found = true; }
  }
  tOut = best;
  return found;
}

// pointTriangleCCD (CollisionDetection.cpp:227-302).  Arguments are relative to triangle
// vertex B at the start (0) and end (1) of the substep.
__device__ __forceinline__ bool pointTriangleCCD(V3 ap0, V3 ab0, V3 ac0, V3 ap1, V3 ab1, V3 ac1, float threshold,
                                                 float& tOut) {
  V3 n0 = norm3(cross3(ab0, ac0));
  V3 n1 = norm3(cross3(ab1, ac1));
  float nDotP0 = dot3(n0, ap0);
  float nDotP1 = dot3(n1, ap1);
  if (mul(nDotP0, nDotP1) >= 0.0f) {
    if (nDotP1 >= 0.0f && nDotP1 < threshold) {
      float bx, by;
      barycentric(ab1, ac1, n1, ap1, bx, by);
      if (outsideTriangle(bx, by)) return false;
      tOut = 0.0f;
      return true;
    }
    return false;
  }
  V3 apd = sub(ap1, ap0), abd = sub(ab1, ab0), acd = sub(ac1, ac0);
  Cubic e{0.0f, 0.0f, 0.0f, 0.0f};
  expandTerm(ap0.x, ab0.y, ac0.z, apd.x, abd.y, acd.z, e);
  expandTerm(-ap0.x, ac0.y, ab0.z, -apd.x, acd.y, abd.z, e);
  expandTerm(-ab0.x, ap0.y, ac0.z, -abd.x, apd.y, acd.z, e);
  expandTerm(ab0.x, ac0.y, ap0.z, abd.x, acd.y, apd.z, e);
  expandTerm(ac0.x, ap0.y, ab0.z, acd.x, apd.y, abd.z, e);
  expandTerm(-ac0.x, ab0.y, ap0.z, -acd.x, abd.y, apd.z, e);
  float t;
  if (!findRootInInterval(e, t)) return false;
  V3 apt = add(ap0, scale(apd, t)), abt = add(ab0, scale(abd, t)), act = add(ac0, scale(acd, t));
  V3 n = norm3(cross3(abt, act));
  float bx, by;
  barycentric(abt, act, n, apt, bx, by);
  if (outsideTriangle(bx, by)) return false;
  tOut = t;
  return true;
}

// edgeEdgeCCD (CollisionDetection.cpp:304-418): edge (a, b) against edge (c, d), everything relative to a, positions
// at the start (0) and end (1) of the substep.  The reference never emits edge contacts (its only call site is commented
// out, Solver.cpp:799-823); the restatement exists so that the narrow phase is complete and checked, and it
// reproduces the reference's behaviour INCLUDING its shadowing bug: in the non-parallel case the closest-point parameters
// u, v are assigned to block-local variables and the outer ones stay 0, so the static proximity test is |c1 - a1| < 0.5.
__device__ __forceinline__ float mix1(float x, float y, float a) { return add(mul(x, sub(1.0f, a)), mul(y, a)); }  // glm::mix
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

__device__ __forceinline__ bool edgeEdgeCCD(V3 ab0, V3 ac0, V3 ad0, V3 ab1, V3 ac1, V3 ad1, float& tOut) {
  {
    V3 cd1 = sub(ad1, ac1);
    float abMagSq = dot3(ab1, ab1), cdMagSq = dot3(cd1, cd1), abDotCd = dot3(ab1, cd1);
    float det = add(mul(abMagSq, -cdMagSq), mul(abDotCd, abDotCd));
    float u = 0.0f, v = 0.0f;
    if (det != 0.0f) {
      // u, v of this branch are shadowed in the reference (:326-328): the outer ones keep 0
    } else {
      float u0 = 0.0f, u1 = 1.0f;
      float v0 = dot3(ac1, ab1), v1 = dot3(ad1, ab1);
      bool flip0 = false, flip1 = false;   // u0 > u1 never holds
      if (v0 > v1) { float t = v0; v0 = v1; v1 = t; flip1 = true; }
      if (u0 >= v1) {
        u = flip0 ? 1.0f : 0.0f; v = flip1 ? 0.0f : 1.0f;
      } else if (v0 >= u1) {
        u = flip0 ? 0.0f : 1.0f; v = flip1 ? 1.0f : 0.0f;
      } else {
        float mid = (u0 > v0) ? mul(add(u0, v1), 0.5f) : mul(add(v0, u1), 0.5f);
        u = (u0 == u1) ? 0.5f : div(sub(mid, u0), sub(u1, u0));
        v = (v0 == v1) ? 0.5f : div(sub(mid, v0), sub(v1, v0));
      }
    }
    u = clamp01(u); v = clamp01(v);
    V3 q0 = V3{mix1(0.0f, ab1.x, u), mix1(0.0f, ab1.y, u), mix1(0.0f, ab1.z, u)};
    V3 q1 = V3{mix1(ac1.x, ad1.x, v), mix1(ac1.y, ad1.y, v), mix1(ac1.z, ad1.z, v)};
    V3 n = sub(q0, q1);
    float dist = __fsqrt_rn(dot3(n, n));
    if (dist < 0.5f) { tOut = 1.0f; return true; }
  }
  V3 abd = sub(ab1, ab0), acd = sub(ac1, ac0), add_ = sub(ad1, ad0);
  Cubic e{0.0f, 0.0f, 0.0f, 0.0f};
  expandTerm(ab0.x, ac0.y, ad0.z, abd.x, acd.y, add_.z, e);
  expandTerm(-ab0.x, ad0.y, ac0.z, -abd.x, add_.y, acd.z, e);
  expandTerm(-ac0.x, ab0.y, ad0.z, -acd.x, abd.y, add_.z, e);
  expandTerm(ac0.x, ad0.y, ab0.z, acd.x, add_.y, abd.z, e);
  expandTerm(ad0.x, ab0.y, ac0.z, add_.x, abd.y, acd.z, e);
  expandTerm(-ad0.x, ac0.y, ab0.z, -add_.x, acd.y, abd.z, e);
  float t;
  if (!findRootInInterval(e, t)) return false;
  V3 abt = add(ab0, scale(abd, t)), act = add(ac0, scale(acd, t)), adt = add(ad0, scale(add_, t));
  V3 cdt = sub(adt, act);
  V3 nt = norm3(cross3(abt, cdt));
  float ux, uy;
  barycentric(abt, V3{-cdt.x, -cdt.y, -cdt.z}, nt, act, ux, uy);   // inverse(mat3(abt, -cdt, nt)) * act
  if (ux < 0.0f || ux > 1.0f || uy < 0.0f || uy > 1.0f) return false;
  tOut = t;
  return true;
}

// TriCompRange / sweptTriRange (Solver.cpp:942-979, :639-677): swept AABB in WORLD units
// (gridSpacing is ignored by the reference); a side longer than `cap` cells gives
// the empty range.  Returns false if the range is empty.
__device__ __forceinline__ bool triCellRange(V3 p0, V3 p1, V3 p2, V3 q0, V3 q1, V3 q2, int& minX, int& minY, int& minZ,
                                             unsigned& lx, unsigned& ly, unsigned& lz, bool& bad) {
  float mxX = fmaxf(fmaxf(fmaxf(p0.x, q0.x), fmaxf(p1.x, q1.x)), fmaxf(p2.x, q2.x));
  float mxY = fmaxf(fmaxf(fmaxf(p0.y, q0.y), fmaxf(p1.y, q1.y)), fmaxf(p2.y, q2.y));
  float mxZ = fmaxf(fmaxf(fmaxf(p0.z, q0.z), fmaxf(p1.z, q1.z)), fmaxf(p2.z, q2.z));
  float mnX = fminf(fminf(fminf(p0.x, q0.x), fminf(p1.x, q1.x)), fminf(p2.x, q2.x));
  float mnY = fminf(fminf(fminf(p0.y, q0.y), fminf(p1.y, q1.y)), fminf(p2.y, q2.y));
  float mnZ = fminf(fminf(fminf(p0.z, q0.z), fminf(p1.z, q1.z)), fminf(p2.z, q2.z));
  float fx = floorf(mnX), fy = floorf(mnY), fz = floorf(mnZ);
  const float lim = 1073741824.0f;  // 2^30: cell coordinates are kept in int32 here
  bad = !(fabsf(fx) < lim && fabsf(fy) < lim && fabsf(fz) < lim && fabsf(mxX) < lim && fabsf(mxY) < lim && fabsf(mxZ) < lim);
  if (bad) { lx = ly = lz = 0; minX = minY = minZ = 0; return false; }
  minX = (int)fx; minY = (int)fy; minZ = (int)fz;
  // static_cast<uint32_t>(ceil(max) - float(min))
  lx = (unsigned)sub(ceilf(mxX), (float)minX);
  ly = (unsigned)sub(ceilf(mxY), (float)minY);
  lz = (unsigned)sub(ceilf(mxZ), (float)minZ);
  return true;
}

// NodeCompRange (Solver.cpp:877-901)
__device__ __forceinline__ void nodeCellRange(V3 p, float radius, float gridScale, int& minX, int& minY, int& minZ,
                                              unsigned& lx, unsigned& ly, unsigned& lz, bool& bad) {
  float r = div(add(radius, 0.5f), gridScale);
  float gx = sub(div(p.x, gridScale), r), gy = sub(div(p.y, gridScale), r), gz = sub(div(p.z, gridScale), r);
  float fx = floorf(gx), fy = floorf(gy), fz = floorf(gz);
  const float lim = 1073741824.0f;
  bad = !(fabsf(fx) < lim && fabsf(fy) < lim && fabsf(fz) < lim);
  if (bad) { lx = ly = lz = 0; minX = minY = minZ = 0; return; }
  minX = (int)fx; minY = (int)fy; minZ = (int)fz;
  float d = mul(2.0f, r);
  lx = (unsigned)ceilf(add(sub(gx, fx), d));  // ceil(fract(min) + 2 r)
  ly = (unsigned)ceilf(add(sub(gy, fy), d));
  lz = (unsigned)ceilf(add(sub(gz, fz), d));
  if (lx > 50u || ly > 50u || lz > 50u) { lx = ly = lz = 0; minX = minY = minZ = 0; }
}

}  // namespace ex
}  // namespace pies
