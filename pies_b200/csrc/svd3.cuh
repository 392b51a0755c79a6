// svd3.cuh — in-register 3x3 SVD and the tet strain / volume projections.
//
// Replaces the reference's Eigen::JacobiSVD<Matrix3f> call sites
// (reference Src/Constraints.cpp:97-112, :225-239).  Same algorithm class —
// two-sided (Kogbetliantz) Jacobi sweeps directly on F, so no squaring of the
// condition number — written for one thread per tet with everything in
// registers: A (9), U (9), V (9).  Conventions of the result match Eigen's:
// sigma >= 0 sorted descending, U and V orthogonal, det(U)det(V) = sign(det F).
#pragma once

#include "common.cuh"

namespace pies {

struct M3 { float m[3][3]; };  // m[row][col]

__device__ __forceinline__ float det3(const M3& a) {
  return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) -
         a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
         a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}

// One Kogbetliantz step on the (P,Q) plane: L * A * J is diagonal on that 2x2 block.
// 1/sqrt(x): SFU estimate (2 ulp) + one Newton step, ~1 ulp without the IEEE sqrt / divide sequences and their branches
__device__ __forceinline__ float rsqrtNewton(float x) {
  const float y = rsqrtf(x);
  return y * fmaf(-0.5f * x, y * y, 1.5f);
}

template <int P, int Q>
__device__ __forceinline__ bool jacobiPair(M3& A, M3& U, M3& V, float thresh) {
  float w = A.m[P][P], x = A.m[P][Q], y = A.m[Q][P], z = A.m[Q][Q];
  if (fabsf(x) <= thresh && fabsf(y) <= thresh) return false;
  // 1) rotation R = [c s; -s c] making R*M symmetric: tan = (y - x) / (w + z).
  // Reciprocal square roots from the SFU plus a Newton step instead of IEEE sqrt + divide sequences: the kernel is bound
  // by instruction issue (r02g); a rotation angle that is off by an ulp is corrected by the next sweep.
  float t = w + z, d = y - x;
  float hh = t * t + d * d;
  float ih = rsqrtNewton(hh);
  const bool okR = hh > 1e-37f;
  float c = okR ? t * ih : 1.0f, s = okR ? d * ih : 0.0f;
  float al = c * w + s * y;          // B = R*M = [al be; be ga]
  float be = c * x + s * z;
  float ga = -s * x + c * z;
  // 2) symmetric Jacobi rotation J = [cj sj; -sj cj] diagonalising B, angle in [-pi/4, pi/4]:
  //    tan 2a = 2 be / (ga - al); cos a = sqrt((1 + cos 2a) / 2), sin a = sin 2a / (2 cos a)
  float dx = ga - al, dy = 2.0f * be;
  float r2 = dx * dx + dy * dy;
  float ir = rsqrtNewton(r2);
  float c2 = fabsf(dx) * ir, s2 = (dx >= 0.0f ? dy : -dy) * ir;
  float v = 0.5f + 0.5f * c2;
  float icj = rsqrtNewton(v);
  const bool okJ = fabsf(be) > 1e-30f && r2 > 1e-37f;
  float cj = okJ ? v * icj : 1.0f, sj = okJ ? 0.5f * s2 * icj : 0.0f;
  // left rotation L = J^T R = [cl sl; -sl cl]
  float cl = cj * c + sj * s, sl = cj * s - sj * c;
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // rows P,Q of A ; columns P,Q of U (U <- U L^T)
    float ap = A.m[P][k], aq = A.m[Q][k];
    A.m[P][k] = cl * ap + sl * aq;
    A.m[Q][k] = -sl * ap + cl * aq;
    float up = U.m[k][P], uq = U.m[k][Q];
    U.m[k][P] = cl * up + sl * uq;
    U.m[k][Q] = -sl * up + cl * uq;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // columns P,Q of A and V (X <- X J)
    float ap = A.m[k][P], aq = A.m[k][Q];
    A.m[k][P] = cj * ap - sj * aq;
    A.m[k][Q] = sj * ap + cj * aq;
    float vp = V.m[k][P], vq = V.m[k][Q];
    V.m[k][P] = cj * vp - sj * vq;
    V.m[k][Q] = sj * vp + cj * vq;
  }
  return true;
}

template <int I, int J>
__device__ __forceinline__ void swapCols(M3& U, M3& V, float (&s)[3]) {
  float t = s[I]; s[I] = s[J]; s[J] = t;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    t = U.m[k][I]; U.m[k][I] = U.m[k][J]; U.m[k][J] = t;
    t = V.m[k][I]; V.m[k][I] = V.m[k][J]; V.m[k][J] = t;
  }
}

// Rotation matrix of a unit quaternion (x, y, z, w).
__device__ __forceinline__ M3 quatToMat(float4 q) {
  const float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
  const float xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z, wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
  M3 R;
  R.m[0][0] = 1.0f - 2.0f * (yy + zz); R.m[0][1] = 2.0f * (xy - wz);        R.m[0][2] = 2.0f * (xz + wy);
  R.m[1][0] = 2.0f * (xy + wz);        R.m[1][1] = 1.0f - 2.0f * (xx + zz); R.m[1][2] = 2.0f * (yz - wx);
  R.m[2][0] = 2.0f * (xz - wy);        R.m[2][1] = 2.0f * (yz + wx);        R.m[2][2] = 1.0f - 2.0f * (xx + yy);
  return R;
}

// Unit quaternion of a proper rotation matrix (Shepperd: the largest of w, x, y, z is computed from the diagonal, the
// others from the off-diagonal sums; normalised at the end, so orthogonality drift of the matrix does not accumulate).
__device__ __forceinline__ float4 matToQuat(const M3& R) {
  const float tr = R.m[0][0] + R.m[1][1] + R.m[2][2];
  float4 q;
  if (tr > 0.0f) {
    q = make_float4(R.m[2][1] - R.m[1][2], R.m[0][2] - R.m[2][0], R.m[1][0] - R.m[0][1], 1.0f + tr);
  } else if (R.m[0][0] >= R.m[1][1] && R.m[0][0] >= R.m[2][2]) {
    q = make_float4(1.0f + R.m[0][0] - R.m[1][1] - R.m[2][2], R.m[0][1] + R.m[1][0], R.m[0][2] + R.m[2][0], R.m[2][1] - R.m[1][2]);
  } else if (R.m[1][1] >= R.m[2][2]) {
    q = make_float4(R.m[0][1] + R.m[1][0], 1.0f + R.m[1][1] - R.m[0][0] - R.m[2][2], R.m[1][2] + R.m[2][1], R.m[0][2] - R.m[2][0]);
  } else {
    q = make_float4(R.m[0][2] + R.m[2][0], R.m[1][2] + R.m[2][1], 1.0f + R.m[2][2] - R.m[0][0] - R.m[1][1], R.m[1][0] - R.m[0][1]);
  }
  const float inv = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
}

// Jacobi sweeps on A = U^T F V for given orthogonal U, V, accumulating into U and V until A is diagonal.
__device__ __forceinline__ void jacobiSweeps(M3& A, M3& U, M3& V) {
  for (int sweep = 0; sweep < 12; ++sweep) {
    float maxDiag = fmaxf(fabsf(A.m[0][0]), fmaxf(fabsf(A.m[1][1]), fabsf(A.m[2][2])));
    float thresh = fmaxf(2.0f * 1.1920929e-7f * maxDiag, 1e-37f);  // Eigen: 2*eps*maxDiagEntry
    bool any = jacobiPair<0, 1>(A, U, V, thresh);
    any |= jacobiPair<0, 2>(A, U, V, thresh);
    any |= jacobiPair<1, 2>(A, U, V, thresh);
    if (!any) break;
  }
}

// Eigen's conventions on the converged factors: sigma >= 0, sorted descending.
__device__ __forceinline__ void svdFinish(const M3& A, M3& U, float (&s)[3], M3& V) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float d = A.m[i][i];
    if (d < 0.0f) {
      d = -d;
#pragma unroll
      for (int k = 0; k < 3; ++k) U.m[k][i] = -U.m[k][i];
    }
    s[i] = d;
  }
  if (s[0] < s[1]) swapCols<0, 1>(U, V, s);
  if (s[0] < s[2]) swapCols<0, 2>(U, V, s);
  if (s[1] < s[2]) swapCols<1, 2>(U, V, s);
}

// Warm-started SVD: qu, qv hold the rotations the previous call converged to (identity quaternions at first); on return
// they hold this call's.  F changes little between two PD iterations, so U0^T F V0 is nearly diagonal and one sweep
// (plus the sweep that finds nothing left to rotate) replaces the four or five a cold start needs.
__device__ __forceinline__ void svd3Warm(const M3& F, float4& qu, float4& qv, M3& U, float (&s)[3], M3& V) {
  U = quatToMat(qu);
  V = quatToMat(qv);
  M3 T, A;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T.m[i][j] = F.m[i][0] * V.m[0][j] + F.m[i][1] * V.m[1][j] + F.m[i][2] * V.m[2][j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A.m[i][j] = U.m[0][i] * T.m[0][j] + U.m[1][i] * T.m[1][j] + U.m[2][i] * T.m[2][j];
  jacobiSweeps(A, U, V);
  qu = matToQuat(U);   // before the sign flips / column swaps of svdFinish: both stay proper rotations
  qv = matToQuat(V);
  svdFinish(A, U, s, V);
}

// F = U diag(s) V^T.
__device__ __forceinline__ void svd3(const M3& F, M3& U, float (&s)[3], M3& V) {
  M3 A = F;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { U.m[i][j] = i == j ? 1.0f : 0.0f; V.m[i][j] = i == j ? 1.0f : 0.0f; }
  for (int sweep = 0; sweep < 12; ++sweep) {
    float maxDiag = fmaxf(fabsf(A.m[0][0]), fmaxf(fabsf(A.m[1][1]), fabsf(A.m[2][2])));
    float thresh = fmaxf(2.0f * 1.1920929e-7f * maxDiag, 1e-37f);  // Eigen: 2*eps*maxDiagEntry
    bool any = jacobiPair<0, 1>(A, U, V, thresh);
    any |= jacobiPair<0, 2>(A, U, V, thresh);
    any |= jacobiPair<1, 2>(A, U, V, thresh);
    if (!any) break;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float d = A.m[i][i];
    if (d < 0.0f) {
      d = -d;
#pragma unroll
      for (int k = 0; k < 3; ++k) U.m[k][i] = -U.m[k][i];
    }
    s[i] = d;
  }
  if (s[0] < s[1]) swapCols<0, 1>(U, V, s);
  if (s[0] < s[2]) swapCols<0, 2>(U, V, s);
  if (s[1] < s[2]) swapCols<1, 2>(U, V, s);
}

// U diag(d) V^T
__device__ __forceinline__ M3 recompose(const M3& U, const float (&d)[3], const M3& V) {
  M3 R;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      R.m[i][j] = U.m[i][0] * d[0] * V.m[j][0] + U.m[i][1] * d[1] * V.m[j][1] + U.m[i][2] * d[2] * V.m[j][2];
  return R;
}

// Deformation gradient F = P * Qinv with P = [x2-x1 | x3-x1 | x4-x1]
// (reference Constraints.cpp:85-91); qinv is glm column-major: qinv[3*c + r].
__device__ __forceinline__ M3 deformationGradient(V3 x1, V3 x2, V3 x3, V3 x4, const float (&qinv)[9]) {
  V3 e1 = x2 - x1, e2 = x3 - x1, e3 = x4 - x1;
  float P[3][3] = {{e1.x, e2.x, e3.x}, {e1.y, e2.y, e3.y}, {e1.z, e2.z, e3.z}};
  M3 F;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      F.m[r][c] = P[r][0] * qinv[3 * c + 0] + P[r][1] * qinv[3 * c + 1] + P[r][2] * qinv[3 * c + 2];
  return F;
}

// Strain limiting (reference Constraints.cpp:100-112): clamp sigma, un-invert.
__device__ __forceinline__ void strainSigma(const float (&s)[3], float detF, float lo, float hi, float (&o)[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = fminf(fmaxf(s[i], lo), hi);
  if (detF < 0.0f) o[2] = -o[2];
}

// Volume preservation (reference Constraints.cpp:186-203, computeD): exactly 10
// projection steps of sigma onto prod(sigma) in [lo, hi], starting from D = 0.
__device__ __forceinline__ void volumeSigma(const float (&s)[3], float lo, float hi, float (&o)[3]) {
  float dx = 0.0f, dy = 0.0f, dz = 0.0f;
#pragma unroll
  for (int it = 0; it < 10; ++it) {
    float sx = s[0] + dx, sy = s[1] + dy, sz = s[2] + dz;
    float product = sx * sy * sz;
    float omega = fminf(fmaxf(product, lo), hi);
    float C = product - omega;
    float gx = sy * sz, gy = sx * sz, gz = sx * sy;
    float k = __fdividef((gx * dx + gy * dy + gz * dz) - C, gx * gx + gy * gy + gz * gz);
    dx = k * gx; dy = k * gy; dz = k * gz;
  }
  o[0] = s[0] + dx; o[1] = s[1] + dy; o[2] = s[2] + dz;
}

}  // namespace pies
