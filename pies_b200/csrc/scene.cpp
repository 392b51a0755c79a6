// scene.cpp — host-side scene construction for the B200 solver.
// Reproduces the node ids, constraint order and rest data of the reference
// factories (reference Src/PrimitiveUtilities.cpp, Src/Constraints.cpp,
// Src/ShapeMatchingConstraint.cpp); cited per function.  Plain float math in the
// same association order as the reference's glm expressions so the rest data
// (Qinv, rest lengths, rest angles) agree to rounding.
#include "scene.h"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace pies {
namespace {

// Lattice addressing used by every primitive (PrimitiveUtilities.cpp:20-39):
// id = k + depth * (j + height * i) + offset.
struct Lattice {
  uint32_t nx, ny, nz;
  uint32_t id(size_t offset, uint32_t i, uint32_t j, uint32_t k) const {
    return k + nz * (j + ny * i) + (uint32_t)offset;
  }
};

inline void sub3(const float* a, const float* b, float* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
inline float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cross3(const float* x, const float* y, float* o) {
  o[0] = x[1] * y[2] - y[1] * x[2];
  o[1] = x[2] * y[0] - y[2] * x[0];
  o[2] = x[0] * y[1] - y[0] * x[1];
}
inline void normalize3(float* v) {
  float s = 1.0f / std::sqrt(dot3(v, v));
  v[0] *= s; v[1] *= s; v[2] *= s;
}

// Adjugate inverse of a column-major 3x3 (m[3*c + r]); same cofactor expansion
// as glm::inverse(mat3) so Qinv matches the reference to rounding.
void inverse3(const float* m, float* inv) {
#define M(c, r) m[3 * (c) + (r)]
  float c00 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
  float c10 = M(0, 1) * M(2, 2) - M(2, 1) * M(0, 2);
  float c20 = M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2);
  float invDet = 1.0f / (M(0, 0) * c00 - M(1, 0) * c10 + M(2, 0) * c20);
  inv[0] = c00 * invDet;
  inv[3] = -(M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2)) * invDet;
  inv[6] = (M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1)) * invDet;
  inv[1] = -c10 * invDet;
  inv[4] = (M(0, 0) * M(2, 2) - M(2, 0) * M(0, 2)) * invDet;
  inv[7] = -(M(0, 0) * M(2, 1) - M(2, 0) * M(0, 1)) * invDet;
  inv[2] = c20 * invDet;
  inv[5] = -(M(0, 0) * M(1, 2) - M(1, 0) * M(0, 2)) * invDet;
  inv[8] = (M(0, 0) * M(1, 1) - M(1, 0) * M(0, 1)) * invDet;
#undef M
}

void inverse3d(const double* m, double* inv) {
#define M(c, r) m[3 * (c) + (r)]
  double c00 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
  double c10 = M(0, 1) * M(2, 2) - M(2, 1) * M(0, 2);
  double c20 = M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2);
  double invDet = 1.0 / (M(0, 0) * c00 - M(1, 0) * c10 + M(2, 0) * c20);
  inv[0] = c00 * invDet;
  inv[3] = -(M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2)) * invDet;
  inv[6] = (M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1)) * invDet;
  inv[1] = -c10 * invDet;
  inv[4] = (M(0, 0) * M(2, 2) - M(2, 0) * M(0, 2)) * invDet;
  inv[7] = -(M(0, 0) * M(2, 1) - M(2, 0) * M(0, 1)) * invDet;
  inv[2] = c20 * invDet;
  inv[5] = -(M(0, 0) * M(1, 2) - M(1, 0) * M(0, 2)) * invDet;
  inv[8] = (M(0, 0) * M(1, 1) - M(1, 0) * M(0, 1)) * invDet;
#undef M
}

// General 4x4 inverse (Gauss-Jordan with partial pivoting, double accumulation).
Mat4 inverse4(const Mat4& a) {
  double w[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) { w[r][c] = a.m[4 * c + r]; w[r][4 + c] = r == c ? 1.0 : 0.0; }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    for (int r = col + 1; r < 4; ++r) if (std::fabs(w[r][col]) > std::fabs(w[piv][col])) piv = r;
    if (piv != col) for (int c = 0; c < 8; ++c) std::swap(w[piv][c], w[col][c]);
    double d = 1.0 / w[col][col];
    for (int c = 0; c < 8; ++c) w[col][c] *= d;
    for (int r = 0; r < 4; ++r) if (r != col) {
      double f = w[r][col];
      for (int c = 0; c < 8; ++c) w[r][c] -= f * w[col][c];
    }
  }
  Mat4 o;
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) o.m[4 * c + r] = (float)w[r][4 + c];
  return o;
}

// (m0*x + m1*y) + (m2*z + m3*1): association of glm's mat4*vec4.
inline void xformPoint(const Mat4& m, const float* p, float* o) {
  for (int r = 0; r < 3; ++r)
    o[r] = (m.m[r] * p[0] + m.m[4 + r] * p[1]) + (m.m[8 + r] * p[2] + m.m[12 + r] * 1.0f);
}
inline bool insideUnitCube(const float* l) {
  return -1.0f <= l[0] && l[0] <= 1.0f && -1.0f <= l[1] && l[1] <= 1.0f && -1.0f <= l[2] && l[2] <= 1.0f;
}

Mat4 mul4(const Mat4& a, const Mat4& b) {
  Mat4 o;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      o.m[4 * c + r] = a.m[r] * b.m[4 * c] + a.m[4 + r] * b.m[4 * c + 1] + a.m[8 + r] * b.m[4 * c + 2] + a.m[12 + r] * b.m[4 * c + 3];
  return o;
}

Mat4 identity4() {
  Mat4 o{};
  o.m[0] = o.m[5] = o.m[10] = o.m[15] = 1.0f;
  return o;
}

float randUnit() { return static_cast<float>(double(std::rand()) / RAND_MAX); }

}  // namespace

// Colour/roughness/metallic come from std::rand() in the reference
// (PrimitiveUtilities.cpp:10-12, e.g. :348-350); same call sequence here so a host
// that seeds std::srand sees the same cosmetics.  Excluded from parity.
HostScene::Cosmetics HostScene::rollCosmetics() {
  Cosmetics c;
  c.color[0] = randUnit(); c.color[1] = randUnit(); c.color[2] = randUnit();
  c.roughness = randUnit();
  c.metallic = static_cast<float>(std::rand() % 2);
  return c;
}

void HostScene::syncVertices(size_t first, const Cosmetics& c) {
  vertices.resize(nodeCount());
  for (size_t i = first; i < vertices.size(); ++i) {
    PiesB200Vertex& v = vertices[i];
    v.position[0] = pos[3 * i]; v.position[1] = pos[3 * i + 1]; v.position[2] = pos[3 * i + 2];
    v.radius = radius[i];
    v.baseColor[0] = c.color[0]; v.baseColor[1] = c.color[1]; v.baseColor[2] = c.color[2];
    v.roughness = c.roughness;
    v.metallic = c.metallic;
  }
  ++topologyVersion;
}

uint32_t HostScene::appendNode(const float p[3], const float v[3], float r, float im) {
  uint32_t id = nodeCount();
  for (int k = 0; k < 3; ++k) { pos.push_back(p[k]); prev.push_back(p[k]); vel.push_back(v[k]); }
  radius.push_back(r);
  invMass.push_back(im);
  ++topologyVersion;
  return id;
}

// createDistanceConstraint (Constraints.cpp:39-56): rest = |b - a| at creation.
void HostScene::appendDistance(uint32_t a, uint32_t b, float w) {
  float d[3];
  sub3(&pos[3 * b], &pos[3 * a], d);
  distId.push_back(a); distId.push_back(b);
  distRest.push_back(std::sqrt(dot3(d, d)));
  distW.push_back(w);
  ++constraintId; ++topologyVersion;
}

// createPositionConstraint (Constraints.cpp:65-74): target = position at creation.
void HostScene::appendPosition(uint32_t a, float w) {
  posId.push_back(a);
  for (int k = 0; k < 3; ++k) posTarget.push_back(pos[3 * a + k]);
  posW.push_back(w);
  ++constraintId; ++topologyVersion;
}

// Qinv = inverse(mat3(x2-x1, x3-x1, x4-x1)) (Constraints.cpp:151-155, :277-281).
void HostScene::tetQinvOf(const uint32_t ids[4], float out[9]) const {
  float q[9];
  for (int c = 0; c < 3; ++c) sub3(&pos[3 * ids[c + 1]], &pos[3 * ids[0]], q + 3 * c);
  inverse3(q, out);
}

void HostScene::appendTet(const uint32_t ids[4], float w, float minStrain, float maxStrain) {
  float qi[9];
  tetQinvOf(ids, qi);
  tetId.insert(tetId.end(), ids, ids + 4);
  tetQinv.insert(tetQinv.end(), qi, qi + 9);
  tetW.push_back(w); tetMin.push_back(minStrain); tetMax.push_back(maxStrain);
  ++constraintId; ++topologyVersion;
}

void HostScene::appendVolume(const uint32_t ids[4], float w, float compression, float stretching) {
  float qi[9];
  tetQinvOf(ids, qi);
  volId.insert(volId.end(), ids, ids + 4);
  volQinv.insert(volQinv.end(), qi, qi + 9);
  volW.push_back(w); volMin.push_back(compression); volMax.push_back(stretching);
  ++constraintId; ++topologyVersion;
}

// createBendConstraint (Constraints.cpp:368-394): rest dihedral angle; the
// reference's unqualified acos() resolves to the double overload.
void HostScene::appendBend(const uint32_t ids[4], float w) {
  float p2[3], p3[3], p4[3], n1[3], n2[3];
  sub3(&pos[3 * ids[1]], &pos[3 * ids[0]], p2);
  sub3(&pos[3 * ids[2]], &pos[3 * ids[0]], p3);
  sub3(&pos[3 * ids[3]], &pos[3 * ids[0]], p4);
  cross3(p2, p3, n1); normalize3(n1);
  cross3(p2, p4, n2); normalize3(n2);
  bendId.insert(bendId.end(), ids, ids + 4);
  bendAngle.push_back(static_cast<float>(std::acos(static_cast<double>(dot3(n1, n2)))));
  bendW.push_back(w);
  ++constraintId; ++topologyVersion;
}

// ShapeMatchingConstraint ctor (ShapeMatchingConstraint.cpp:6-48): fp32 centroid with
// equal weights, centred material coordinates widened to fp64, mass-weighted
// covariance accumulated in fp64 from fp32 outer products, Qinv = Q^-1.
void HostScene::appendShape(uint32_t n, const uint32_t* ids, const float* mat, float w) {
  float com[3] = {0, 0, 0};
  float weight = 1.0f / static_cast<float>(n);
  for (uint32_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) com[k] += weight * mat[3 * i + k];
  double Q[9] = {0};
  for (uint32_t i = 0; i < n; ++i) {
    float mc[3];
    sub3(mat + 3 * i, com, mc);
    shapeId.push_back(ids[i]);
    for (int k = 0; k < 3; ++k) shapeMat.push_back(mc[k]);
    float im = invMass[ids[i]];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Q[3 * c + r] += (double)((mc[r] * mc[c]) / im);
  }
  double Qi[9];
  inverse3d(Q, Qi);
  shapeQinv.insert(shapeQinv.end(), Qi, Qi + 9);
  shapeOff.push_back((uint32_t)shapeId.size());
  shapeW.push_back(w);
  const double ident[4] = {0.0, 0.0, 0.0, 1.0};
  shapeQuat.insert(shapeQuat.end(), ident, ident + 4);
  ++topologyVersion;
}

// GoalMatchingConstraint ctor (ShapeMatchingConstraint.cpp:124-137).
void HostScene::appendGoal(uint32_t n, const uint32_t* ids, float w) {
  for (uint32_t i = 0; i < n; ++i) {
    goalId.push_back(ids[i]);
    for (int k = 0; k < 3; ++k) goalMat.push_back(pos[3 * ids[i] + k]);
  }
  goalOff.push_back((uint32_t)goalId.size());
  goalXform.push_back(identity4());
  goalW.push_back(w);
  ++topologyVersion;
}

// Solver::addNodes (PrimitiveUtilities.cpp:42-75): mass 1, radius 0.5, at rest.
void HostScene::addNodes(uint32_t n, const float* xyz) {
  Cosmetics c = rollCosmetics();
  size_t first = nodeCount();
  const float zero[3] = {0, 0, 0};
  for (uint32_t i = 0; i < n; ++i) appendNode(xyz + 3 * i, zero, 0.5f, 1.0f / 1.0f);
  syncVertices(first, c);
}

namespace {
// The 12 boundary triangles per lattice cell face pair, in the reference's emission
// order (PrimitiveUtilities.cpp:524-606).  For each of the three axis pairs (u,v)
// the reference emits, per (u,v) cell, two triangles on the "low" face and two on
// the "high" face; corners are (du,dv) offsets.
struct FaceTri { uint8_t c[3][2]; };
const FaceTri kFaceIJ[4] = {  // faces k=0 (first two) and k=nz-1 (last two)
    {{{0, 0}, {1, 1}, {1, 0}}}, {{{0, 0}, {0, 1}, {1, 1}}}, {{{0, 0}, {1, 0}, {1, 1}}}, {{{0, 0}, {1, 1}, {0, 1}}}};
const FaceTri kFaceIK[4] = {  // faces j=0 and j=ny-1
    {{{0, 0}, {1, 0}, {1, 1}}}, {{{0, 0}, {1, 1}, {0, 1}}}, {{{0, 0}, {1, 1}, {1, 0}}}, {{{0, 0}, {0, 1}, {1, 1}}}};
const FaceTri kFaceJK[4] = {  // faces i=0 and i=nx-1
    {{{0, 0}, {1, 1}, {1, 0}}}, {{{0, 0}, {0, 1}, {1, 1}}}, {{{0, 0}, {1, 0}, {1, 1}}}, {{{0, 0}, {1, 1}, {0, 1}}}};

void emitHullTriangles(const Lattice& g, size_t off, std::vector<uint32_t>& tris) {
  for (uint32_t i = 0; i + 1 < g.nx; ++i)
    for (uint32_t j = 0; j + 1 < g.ny; ++j)
      for (int t = 0; t < 4; ++t) {
        uint32_t k = t < 2 ? 0 : g.nz - 1;
        for (int c = 0; c < 3; ++c) tris.push_back(g.id(off, i + kFaceIJ[t].c[c][0], j + kFaceIJ[t].c[c][1], k));
      }
  for (uint32_t i = 0; i + 1 < g.nx; ++i)
    for (uint32_t k = 0; k + 1 < g.nz; ++k)
      for (int t = 0; t < 4; ++t) {
        uint32_t j = t < 2 ? 0 : g.ny - 1;
        for (int c = 0; c < 3; ++c) tris.push_back(g.id(off, i + kFaceIK[t].c[c][0], j, k + kFaceIK[t].c[c][1]));
      }
  for (uint32_t j = 0; j + 1 < g.ny; ++j)
    for (uint32_t k = 0; k + 1 < g.nz; ++k)
      for (int t = 0; t < 4; ++t) {
        uint32_t i = t < 2 ? 0 : g.nx - 1;
        for (int c = 0; c < 3; ++c) tris.push_back(g.id(off, i, j + kFaceJK[t].c[c][0], k + kFaceJK[t].c[c][1]));
      }
}

// Two triangles per quad of a single-layer sheet (PrimitiveUtilities.cpp:933-945, :1254-1266).
void emitSheetTriangles(const Lattice& g, size_t off, std::vector<uint32_t>& tris) {
  for (uint32_t i = 0; i + 1 < g.nx; ++i)
    for (uint32_t j = 0; j + 1 < g.ny; ++j) {
      const uint32_t quad[6][2] = {{0, 0}, {1, 1}, {1, 0}, {0, 0}, {0, 1}, {1, 1}};
      for (auto& q : quad) tris.push_back(g.id(off, i + q[0], j + q[1], 0));
    }
}
}  // namespace

// Solver::createTetBox (PrimitiveUtilities.cpp:330-618): 3x3x3 lattice (10x2x10 when
// hinged), six tets per cell along the (000)-(111) diagonal, one strain and one
// volume constraint per tet in alternation, then the hull triangles.
void HostScene::createTetBox(const float t[3], float scale, const float v0[3], float w, float mass, bool hinged) {
  Lattice g = hinged ? Lattice{10, 2, 10} : Lattice{3, 3, 3};
  size_t off = nodeCount();
  Cosmetics cos = rollCosmetics();
  for (uint32_t i = 0; i < g.nx; ++i)
    for (uint32_t j = 0; j < g.ny; ++j)
      for (uint32_t k = 0; k < g.nz; ++k) {
        float p[3] = {scale * (float)i + t[0], scale * (float)j + t[1], scale * (float)k + t[2]};
        appendNode(p, v0, 0.95f * 0.5f * scale, 1.0f / mass);
      }
  // middle two corners of each of the six tets around the cell diagonal, as (di,dj,dk)
  static const uint8_t kMid[6][2][3] = {{{0, 0, 1}, {0, 1, 1}}, {{0, 1, 0}, {0, 1, 1}}, {{0, 0, 1}, {1, 0, 1}},
                                        {{1, 0, 0}, {1, 0, 1}}, {{0, 1, 0}, {1, 1, 0}}, {{1, 0, 0}, {1, 1, 0}}};
  for (uint32_t i = 0; i + 1 < g.nx; ++i)
    for (uint32_t j = 0; j + 1 < g.ny; ++j)
      for (uint32_t k = 0; k + 1 < g.nz; ++k)
        for (const auto& m : kMid) {
          uint32_t ids[4] = {g.id(off, i, j, k), g.id(off, i + m[0][0], j + m[0][1], k + m[0][2]),
                             g.id(off, i + m[1][0], j + m[1][1], k + m[1][2]), g.id(off, i + 1, j + 1, k + 1)};
          appendTet(ids, w, 0.8f, 1.0f);
          appendVolume(ids, w, 1.0f, 1.0f);
        }
  emitHullTriangles(g, off, triangles);
  syncVertices(off, cos);
}

// Solver::createBox (PrimitiveUtilities.cpp:620-847): 5x5x5 lattice of unit-mass nodes,
// axis springs plus the four body diagonals of every cell.
void HostScene::createBox(const float t[3], float scale, float w) {
  Lattice g{5, 5, 5};
  size_t off = nodeCount();
  size_t firstDist = distW.size();
  Cosmetics cos = rollCosmetics();
  const float zero[3] = {0, 0, 0};
  for (uint32_t i = 0; i < g.nx; ++i)
    for (uint32_t j = 0; j < g.ny; ++j)
      for (uint32_t k = 0; k < g.nz; ++k) {
        float p[3] = {scale * (float)i + t[0], scale * (float)j + t[1], scale * (float)k + t[2]};
        appendNode(p, zero, 0.5f * scale, 1.0f);
      }
  for (uint32_t i = 0; i < g.nx; ++i)
    for (uint32_t j = 0; j < g.ny; ++j)
      for (uint32_t k = 0; k < g.nz; ++k) {
        bool bi = i + 1 < g.nx, bj = j + 1 < g.ny, bk = k + 1 < g.nz;
        uint32_t n000 = g.id(off, i, j, k);
        if (bi) appendDistance(n000, g.id(off, i + 1, j, k), w);
        if (bj) appendDistance(n000, g.id(off, i, j + 1, k), w);
        if (bk) appendDistance(n000, g.id(off, i, j, k + 1), w);
        if (bi && bj && bk) {
          appendDistance(n000, g.id(off, i + 1, j + 1, k + 1), w);
          appendDistance(g.id(off, i + 1, j, k), g.id(off, i, j + 1, k + 1), w);
          appendDistance(g.id(off, i, j + 1, k), g.id(off, i + 1, j, k + 1), w);
          appendDistance(g.id(off, i, j, k + 1), g.id(off, i + 1, j + 1, k), w);
        }
      }
  emitHullTriangles(g, off, triangles);
  for (size_t c = firstDist; c < distW.size(); ++c) { lines.push_back(distId[2 * c]); lines.push_back(distId[2 * c + 1]); }
  syncVertices(off, cos);
}

// Solver::createSheet (PrimitiveUtilities.cpp:849-976): 20x20 sheet in the xz plane,
// border nodes pinned, axis + both diagonal springs.
void HostScene::createSheet(const float t[3], float scale, float mass, float w) {
  Lattice g{20, 20, 1};
  size_t off = nodeCount();
  size_t firstDist = distW.size();
  Cosmetics cos = rollCosmetics();
  const float zero[3] = {0, 0, 0};
  for (uint32_t i = 0; i < g.nx; ++i)
    for (uint32_t j = 0; j < g.ny; ++j) {
      float p[3] = {scale * (float)i + t[0], scale * 0.0f + t[1], scale * (float)j + t[2]};
      uint32_t id = appendNode(p, zero, 0.5f * scale, 1.0f / mass);
      if (i == 0 || i == g.nx - 1 || j == 0 || j == g.ny - 1) appendPosition(id, w);
    }
  for (uint32_t i = 0; i < g.nx; ++i)
    for (uint32_t j = 0; j < g.ny; ++j) {
      bool bi = i + 1 < g.nx, bj = j + 1 < g.ny;
      uint32_t n00 = g.id(off, i, j, 0);
      if (bi) appendDistance(n00, g.id(off, i + 1, j, 0), w);
      if (bj) appendDistance(n00, g.id(off, i, j + 1, 0), w);
      if (bi && bj) {
        appendDistance(n00, g.id(off, i + 1, j + 1, 0), w);
        appendDistance(g.id(off, i + 1, j, 0), g.id(off, i, j + 1, 0), w);
      }
    }
  emitSheetTriangles(g, off, triangles);
  for (size_t c = firstDist; c < distW.size(); ++c) { lines.push_back(distId[2 * c]); lines.push_back(distId[2 * c + 1]); }
  syncVertices(off, cos);
}

// Solver::createShapeMatchingBox (PrimitiveUtilities.cpp:985-1048): the factory overrides
// scale with 0.5, ignores the initial velocity, mass 10, one cluster over the whole body.
void HostScene::createShapeMatchingBox(const float t[3], uint32_t cx, uint32_t cy, uint32_t cz, float /*scale*/,
                                       const float* /*v0*/, float w) {
  Lattice g{cx, cy, cz};
  const float scale = 0.5f;
  size_t off = nodeCount();
  Cosmetics cos = rollCosmetics();
  const float zero[3] = {0, 0, 0};
  for (uint32_t i = 0; i < g.nx; ++i)
    for (uint32_t j = 0; j < g.ny; ++j)
      for (uint32_t k = 0; k < g.nz; ++k) {
        float p[3] = {scale * (float)i + t[0], scale * (float)j + t[1], scale * (float)k + t[2]};
        appendNode(p, zero, 0.5f * scale, 1.0f / 10.0f);
      }
  uint32_t n = g.nx * g.ny * g.nz;
  std::vector<uint32_t> ids(n);
  for (uint32_t i = 0; i < n; ++i) ids[i] = (uint32_t)off + i;
  appendShape(n, ids.data(), &pos[3 * off], w);
  syncVertices(off, cos);
}

// Solver::createShapeMatchingSheet (PrimitiveUtilities.cpp:1050-1125): 50x50 sheet in the xy
// plane, overlapping 3x3 patches.  The reference's patch index uses patchHeight (3) as the
// row pitch of a 16x16 patch table, so patches alias and most of the 256 clusters are empty;
// reproduced as is (empty clusters have no members and do nothing).
void HostScene::createShapeMatchingSheet(const float t[3], float scale, const float* /*v0*/, float w) {
  Lattice g{50, 50, 1};
  const uint32_t pw = 3, ph = 3;
  size_t off = nodeCount();
  Cosmetics cos = rollCosmetics();
  struct Patch { std::vector<uint32_t> ids; std::vector<float> mat; };
  std::vector<Patch> patches((g.nx / pw) * (g.ny / ph));
  const float zero[3] = {0, 0, 0};
  for (uint32_t i = 0; i < g.nx; ++i)
    for (uint32_t j = 0; j < g.ny; ++j) {
      float p[3] = {scale * (float)i + t[0], scale * (float)j + t[1], scale * 0.0f + t[2]};
      uint32_t id = appendNode(p, zero, 0.5f * scale, 1.0f);
      auto put = [&](uint32_t patch) {
        patches[patch].ids.push_back(id);
        patches[patch].mat.insert(patches[patch].mat.end(), p, p + 3);
      };
      put(i / pw * ph + j / ph);
      if (i % pw == pw - 1 && i < g.nx - 1) put((1 + i / pw) * ph + j / ph);
      if (j % ph == ph - 1 && j < g.ny - 1) put(i / pw * ph + j / ph + 1);
    }
  for (const Patch& p : patches) appendShape((uint32_t)p.ids.size(), p.ids.data(), p.mat.data(), w);
  syncVertices(off, cos);
}

// Solver::createBendSheet (PrimitiveUtilities.cpp:1127-1289): 10x10 sheet, first three rows
// pinned, axis + one diagonal spring, dihedral constraints across the diagonal and the two
// neighbouring quads.
void HostScene::createBendSheet(const float t[3], float scale, float w) {
  Lattice g{10, 10, 1};
  Cosmetics cos = rollCosmetics();
  size_t off = nodeCount();
  size_t firstDist = distW.size();
  const float zero[3] = {0, 0, 0};
  for (uint32_t i = 0; i < g.nx; ++i)
    for (uint32_t j = 0; j < g.ny; ++j) {
      float p[3] = {scale * (float)i + t[0], scale * 0.0f + t[1], scale * (float)j + t[2]};
      uint32_t id = appendNode(p, zero, 0.5f * scale, 1.0f);
      if (i < 3) appendPosition(id, w);
    }
  for (uint32_t i = 0; i < g.nx; ++i)
    for (uint32_t j = 0; j < g.ny; ++j) {
      bool bi = i + 1 < g.nx, bj = j + 1 < g.ny;
      uint32_t n00 = g.id(off, i, j, 0);
      if (bi) appendDistance(n00, g.id(off, i + 1, j, 0), w);
      if (bj) appendDistance(n00, g.id(off, i, j + 1, 0), w);
      if (bi && bj) appendDistance(n00, g.id(off, i + 1, j + 1, 0), w);
    }
  for (uint32_t i = 0; i < g.nx; ++i)
    for (uint32_t j = 0; j < g.ny; ++j) {
      uint32_t n00 = g.id(off, i, j, 0), n01 = g.id(off, i, j + 1, 0), n10 = g.id(off, i + 1, j, 0),
               n11 = g.id(off, i + 1, j + 1, 0);
      if (i + 1 < g.nx && j + 1 < g.ny) {
        uint32_t ids[4] = {n00, n11, n10, n01};
        appendBend(ids, w);
      }
      if (i + 2 < g.nx && j + 2 < g.ny) {
        uint32_t a[4] = {n10, n11, n00, g.id(off, i + 2, j + 1, 0)};
        appendBend(a, w);
        uint32_t b[4] = {n01, n11, n00, g.id(off, i + 1, j + 2, 0)};
        appendBend(b, w);
      }
    }
  emitSheetTriangles(g, off, triangles);
  for (size_t c = firstDist; c < distW.size(); ++c) { lines.push_back(distId[2 * c]); lines.push_back(distId[2 * c + 1]); }
  syncVertices(off, cos);
}

// The part of Solver::addTriMeshVolume after tetrahedralize() (PrimitiveUtilities.cpp:243-327):
// boundary triangles first, then nodes (density is used as the per-node mass, :176,280),
// then per tet an optional strain and an optional volume constraint.
void HostScene::addTetMeshVolume(uint32_t nPoints, const float* xyz, uint32_t nTets, const uint32_t* tetIdx,
                                 uint32_t nTris, const uint32_t* triIdx, const float v0[3], float density,
                                 float strainStiffness, float minStrain, float maxStrain, float volumeStiffness,
                                 float compression, float stretching) {
  Cosmetics cos = rollCosmetics();
  size_t off = nodeCount();
  for (uint32_t i = 0; i < 3 * nTris; ++i) triangles.push_back((uint32_t)off + triIdx[i]);
  for (uint32_t i = 0; i < nPoints; ++i) appendNode(xyz + 3 * i, v0, 0.5f, 1.0f / density);
  for (uint32_t e = 0; e < nTets; ++e) {
    uint32_t ids[4];
    for (int k = 0; k < 4; ++k) ids[k] = (uint32_t)off + tetIdx[4 * e + k];
    if (strainStiffness != 0.0f) appendTet(ids, strainStiffness, minStrain, maxStrain);
    if (volumeStiffness != 0.0f) appendVolume(ids, volumeStiffness, compression, stretching);
  }
  syncVertices(off, cos);
}

// Solver::addFixedRegions (PrimitiveUtilities.cpp:77-112): one goal-matching cluster per
// region over the nodes whose region-local coordinates lie in [-1,1]^3.
void HostScene::addFixedRegions(uint32_t n, const float* mats, float w) {
  for (uint32_t r = 0; r < n; ++r) {
    FixedRegion reg;
    std::memcpy(reg.initial.m, mats + 16 * r, sizeof(float) * 16);
    reg.invInitial = inverse4(reg.initial);
    reg.goal = (uint32_t)goalW.size();
    std::vector<uint32_t> inside;
    for (uint32_t i = 0; i < nodeCount(); ++i) {
      float l[3];
      xformPoint(reg.invInitial, &pos[3 * i], l);
      if (insideUnitCube(l)) inside.push_back(i);
    }
    appendGoal((uint32_t)inside.size(), inside.data(), w);
    fixedRegions.push_back(reg);
  }
}

// Solver::updateFixedRegions (PrimitiveUtilities.cpp:114-128).
bool HostScene::updateFixedRegions(uint32_t n, const float* mats) {
  if (n != fixedRegions.size()) return false;
  for (uint32_t r = 0; r < n; ++r) {
    Mat4 cur;
    std::memcpy(cur.m, mats + 16 * r, sizeof(float) * 16);
    goalXform[fixedRegions[r].goal] = mul4(cur, fixedRegions[r].invInitial);
  }
  goalXformDirty = true;
  return true;
}

// Solver::addLinkedRegions (PrimitiveUtilities.cpp:130-162): a shape-matching cluster per
// region with at least three nodes.
void HostScene::addLinkedRegions(uint32_t n, const float* mats, float w) {
  for (uint32_t r = 0; r < n; ++r) {
    Mat4 reg;
    std::memcpy(reg.m, mats + 16 * r, sizeof(float) * 16);
    Mat4 inv = inverse4(reg);
    std::vector<uint32_t> ids;
    std::vector<float> mat;
    for (uint32_t i = 0; i < nodeCount(); ++i) {
      float l[3];
      xformPoint(inv, &pos[3 * i], l);
      if (insideUnitCube(l)) { ids.push_back(i); mat.insert(mat.end(), &pos[3 * i], &pos[3 * i] + 3); }
    }
    if (ids.size() >= 3) appendShape((uint32_t)ids.size(), ids.data(), mat.data(), w);
  }
}

// Solver::clear (Solver.cpp:488-507).  Like the reference it keeps the fixed regions; unlike
// the reference it also drops the collision lists and forces a topology rebuild (F2 fixed on purpose).
void HostScene::clear() {
  lines.clear(); triangles.clear();
  pos.clear(); prev.clear(); vel.clear(); radius.clear(); invMass.clear();
  distId.clear(); distRest.clear(); distW.clear();
  tetId.clear(); tetQinv.clear(); tetW.clear(); tetMin.clear(); tetMax.clear();
  volId.clear(); volQinv.clear(); volW.clear(); volMin.clear(); volMax.clear();
  shapeOff.assign(1, 0); shapeId.clear(); shapeMat.clear(); shapeQinv.clear(); shapeW.clear(); shapeQuat.clear();
  goalOff.assign(1, 0); goalId.clear(); goalMat.clear(); goalXform.clear(); goalW.clear();
  bendId.clear(); bendAngle.clear(); bendW.clear();
  posId.clear(); posTarget.clear(); posW.clear();
  vertices.clear();
  constraintId = 0;
  ++topologyVersion;
}

}  // namespace pies
