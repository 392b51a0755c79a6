// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// White-box C driver around the UNMODIFIED reference (nithinp7/Pies @ 2e552ea,
// compiled from the sources where they lie under /root/reference by
// oracle/Makefile into oracle/_ref/libpies_ref.so).  It exposes, through a
// plain C ABI that tests/ and bench.py's cpu_baseline / --impl reference legs
// load with ctypes:
//   * the reference's own public factories and tick (Include/Pies/Solver.h:54-116),
//   * white-box readers for velocities, constraint tables and the collision
//     lists (Solver.h:147-193 are private; opened with the `#define private
//     public` trick after pre-including every std/third-party header so only
//     Pies' own classes are affected; class layout is unchanged),
//   * per-function probes (projections, CCD, cell ranges, hash occupancy) used
//     by the per-kernel parity tests.
// Nothing here re-implements reference arithmetic: every number comes from the
// reference's own functions.

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Dense>
#include <Eigen/Sparse>
#include <Eigen/SparseCholesky>
#include <Eigen/SparseCore>
#include <glm/glm.hpp>
#include <glm/gtc/matrix_inverse.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <parallel_hashmap/phmap.h>

#define private public
#define protected public
#include <Pies/CollisionDetection.h>
#include <Pies/Solver.h>
#undef private
#undef protected

using namespace Pies;

namespace {
struct RefOptions { // mirrors Pies::SolverOptions (Solver.h:23-38) field by field
  float fixedTimestepSize;
  uint32_t timeSubsteps;
  uint32_t iterations;
  uint32_t collisionStabilizationIterations;
  float collisionThresholdDistance;
  float collisionThickness;
  float gravity;
  float damping;
  float friction;
  float staticFrictionThreshold;
  float floorHeight;
  float gridSpacing;
  uint32_t threadCount;
  uint32_t solver; // 0 = PBD, 1 = PD
};

inline Solver* S(void* h) { return reinterpret_cast<Solver*>(h); }
inline glm::vec3 V3(const float* p) { return glm::vec3(p[0], p[1], p[2]); }
inline glm::mat4 M4(const float* m) { // 16 floats, column-major like glm
  glm::mat4 r;
  std::memcpy(&r[0][0], m, 16 * sizeof(float));
  return r;
}
} // namespace

extern "C" {

void pref_default_options(RefOptions* o) {
  SolverOptions d{};
  o->fixedTimestepSize = d.fixedTimestepSize;
  o->timeSubsteps = d.timeSubsteps;
  o->iterations = d.iterations;
  o->collisionStabilizationIterations = d.collisionStabilizationIterations;
  o->collisionThresholdDistance = d.collisionThresholdDistance;
  o->collisionThickness = d.collisionThickness;
  o->gravity = d.gravity;
  o->damping = d.damping;
  o->friction = d.friction;
  o->staticFrictionThreshold = d.staticFrictionThreshold;
  o->floorHeight = d.floorHeight;
  o->gridSpacing = d.gridSpacing;
  o->threadCount = d.threadCount;
  o->solver = d.solver == SolverName::PD ? 1u : 0u;
}

void* pref_create(const RefOptions* o) {
  SolverOptions d{};
  d.fixedTimestepSize = o->fixedTimestepSize;
  d.timeSubsteps = o->timeSubsteps;
  d.iterations = o->iterations;
  d.collisionStabilizationIterations = o->collisionStabilizationIterations;
  d.collisionThresholdDistance = o->collisionThresholdDistance;
  d.collisionThickness = o->collisionThickness;
  d.gravity = o->gravity;
  d.damping = o->damping;
  d.friction = o->friction;
  d.staticFrictionThreshold = o->staticFrictionThreshold;
  d.floorHeight = o->floorHeight;
  d.gridSpacing = o->gridSpacing;
  d.threadCount = o->threadCount;
  d.solver = o->solver ? SolverName::PD : SolverName::PBD;
  return new Solver(d);
}

void pref_destroy(void* h) { delete S(h); }
void pref_srand(uint32_t seed) { std::srand(seed); }

// ---- public factories, called exactly as a host app would -----------------
void pref_create_tet_box(void* h, const float* t, float scale, const float* v0,
                         float w, float mass, int hinged) {
  S(h)->createTetBox(V3(t), scale, V3(v0), w, mass, hinged != 0);
}
void pref_create_box(void* h, const float* t, float scale, float w) {
  S(h)->createBox(V3(t), scale, w);
}
void pref_create_sheet(void* h, const float* t, float scale, float mass, float k) {
  S(h)->createSheet(V3(t), scale, mass, k);
}
void pref_create_shape_matching_box(void* h, const float* t, uint32_t cx, uint32_t cy,
                                    uint32_t cz, float scale, const float* v0, float w) {
  S(h)->createShapeMatchingBox(V3(t), cx, cy, cz, scale, V3(v0), w);
}
void pref_create_shape_matching_sheet(void* h, const float* t, float scale,
                                      const float* v0, float w) {
  S(h)->createShapeMatchingSheet(V3(t), scale, V3(v0), w);
}
void pref_create_bend_sheet(void* h, const float* t, float scale, float w) {
  S(h)->createBendSheet(V3(t), scale, w);
}
void pref_add_nodes(void* h, uint32_t n, const float* xyz) {
  std::vector<glm::vec3> v(n);
  for (uint32_t i = 0; i < n; ++i) v[i] = V3(xyz + 3 * i);
  S(h)->addNodes(v);
}
void pref_add_tri_mesh_volume(void* h, uint32_t nv, const float* xyz, uint32_t nidx,
                              const uint32_t* idx, const float* v0, float density,
                              float strainStiffness, float minStrain, float maxStrain,
                              float volumeStiffness, float compression, float stretching) {
  std::vector<glm::vec3> v(nv);
  for (uint32_t i = 0; i < nv; ++i) v[i] = V3(xyz + 3 * i);
  std::vector<uint32_t> ix(idx, idx + nidx);
  S(h)->addTriMeshVolume(v, ix, V3(v0), density, strainStiffness, minStrain, maxStrain,
                         volumeStiffness, compression, stretching);
}
void pref_add_fixed_regions(void* h, uint32_t n, const float* mats, float w) {
  std::vector<glm::mat4> m(n);
  for (uint32_t i = 0; i < n; ++i) m[i] = M4(mats + 16 * i);
  S(h)->addFixedRegions(m, w);
}
void pref_update_fixed_regions(void* h, uint32_t n, const float* mats) {
  std::vector<glm::mat4> m(n);
  for (uint32_t i = 0; i < n; ++i) m[i] = M4(mats + 16 * i);
  S(h)->updateFixedRegions(m);
}
void pref_add_linked_regions(void* h, uint32_t n, const float* mats, float w) {
  std::vector<glm::mat4> m(n);
  for (uint32_t i = 0; i < n; ++i) m[i] = M4(mats + 16 * i);
  S(h)->addLinkedRegions(m, w);
}
void pref_clear(void* h) { S(h)->clear(); }
void pref_set_release_hinge(void* h, int v) { S(h)->releaseHinge = v != 0; }

// ---- white-box scene building (F14: no public API for ropes / bulk scenes) --
// Reserving up-front makes the factories' own exact-size reserve() calls no-ops,
// which removes the quadratic re-allocation without touching results.
void pref_reserve(void* h, size_t nodes, size_t dist, size_t tets, size_t vols, size_t tris) {
  Solver* s = S(h);
  s->_nodes.reserve(nodes);
  s->_vertices.reserve(nodes);
  s->_distanceConstraints.reserve(dist);
  s->_tetConstraints.reserve(tets);
  s->_volumeConstraints.reserve(vols);
  s->_tets.reserve(tets);
  s->_triangles.reserve(tris);
}
uint32_t pref_append_node(void* h, const float* pos, const float* vel, float radius, float invMass) {
  Solver* s = S(h);
  Node& n = s->_nodes.emplace_back();
  n.id = static_cast<uint32_t>(s->_nodes.size() - 1);
  n.position = V3(pos);
  n.prevPosition = n.position;
  n.velocity = V3(vel);
  n.radius = radius;
  n.invMass = invMass;
  s->_vertices.resize(s->_nodes.size());
  s->_vertices[n.id].position = n.position;
  s->_vertices[n.id].radius = radius;
  return n.id;
}
void pref_append_distance(void* h, uint32_t a, uint32_t b, float w) {
  Solver* s = S(h);
  s->_distanceConstraints.push_back(
      createDistanceConstraint(s->_constraintId++, s->_nodes[a], s->_nodes[b], w));
}
void pref_append_position(void* h, uint32_t a, float w) {
  Solver* s = S(h);
  s->_positionConstraints.push_back(createPositionConstraint(s->_constraintId++, s->_nodes[a], w));
}
void pref_append_tet(void* h, const uint32_t* ids, float w, float minStrain, float maxStrain) {
  Solver* s = S(h);
  s->_tetConstraints.push_back(createTetrahedralConstraint(
      s->_constraintId++, w, s->_nodes[ids[0]], s->_nodes[ids[1]], s->_nodes[ids[2]],
      s->_nodes[ids[3]], minStrain, maxStrain));
}
void pref_append_volume(void* h, const uint32_t* ids, float w, float compression, float stretching) {
  Solver* s = S(h);
  s->_volumeConstraints.push_back(createVolumeConstraint(
      s->_constraintId++, w, s->_nodes[ids[0]], s->_nodes[ids[1]], s->_nodes[ids[2]],
      s->_nodes[ids[3]], compression, stretching));
}
void pref_append_bend(void* h, const uint32_t* ids, float w) {
  Solver* s = S(h);
  s->_bendConstraints.push_back(createBendConstraint(
      s->_constraintId++, w, s->_nodes[ids[0]], s->_nodes[ids[1]], s->_nodes[ids[2]],
      s->_nodes[ids[3]]));
}
void pref_append_triangle(void* h, uint32_t a, uint32_t b, uint32_t c) {
  S(h)->_triangles.push_back(Triangle{{a, b, c}});
}
void pref_append_shape(void* h, uint32_t n, const uint32_t* ids, float w) {
  Solver* s = S(h);
  std::vector<uint32_t> idx(ids, ids + n);
  std::vector<glm::vec3> mc(n);
  for (uint32_t i = 0; i < n; ++i) mc[i] = s->_nodes[ids[i]].position;
  s->_shapeConstraints.emplace_back(s->_nodes, idx, mc, w);
}

// ---- stepping ---------------------------------------------------------------
void pref_tick(void* h, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i) S(h)->tick(0.0f);
}
int pref_sim_failed(void* h) { return S(h)->_simFailed ? 1 : 0; }
// Runs only the detection pass of tickPD (Solver.cpp:240) on the current state.
void pref_detect(void* h) { S(h)->_parallelPointTriangleCollisions(); }

// ---- state access -----------------------------------------------------------
uint32_t pref_node_count(void* h) { return (uint32_t)S(h)->_nodes.size(); }
uint32_t pref_triangle_count(void* h) { return (uint32_t)S(h)->_triangles.size(); }
uint32_t pref_line_index_count(void* h) { return (uint32_t)S(h)->_lines.size(); }
uint32_t pref_tet_count(void* h) { return (uint32_t)S(h)->_tetConstraints.size(); }
uint32_t pref_volume_count(void* h) { return (uint32_t)S(h)->_volumeConstraints.size(); }
uint32_t pref_distance_count(void* h) { return (uint32_t)S(h)->_distanceConstraints.size(); }
uint32_t pref_position_count(void* h) { return (uint32_t)S(h)->_positionConstraints.size(); }
uint32_t pref_bend_count(void* h) { return (uint32_t)S(h)->_bendConstraints.size(); }
uint32_t pref_shape_count(void* h) { return (uint32_t)S(h)->_shapeConstraints.size(); }
uint32_t pref_goal_count(void* h) { return (uint32_t)S(h)->_goalConstraints.size(); }

// which: 0 position, 1 prevPosition, 2 velocity
void pref_get_node_vec(void* h, int which, float* out) {
  Solver* s = S(h);
  for (size_t i = 0; i < s->_nodes.size(); ++i) {
    const Node& n = s->_nodes[i];
    const glm::vec3& v = which == 0 ? n.position : which == 1 ? n.prevPosition : n.velocity;
    out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
  }
}
void pref_set_node_vec(void* h, int which, const float* in) {
  Solver* s = S(h);
  for (size_t i = 0; i < s->_nodes.size(); ++i) {
    Node& n = s->_nodes[i];
    glm::vec3& v = which == 0 ? n.position : which == 1 ? n.prevPosition : n.velocity;
    v = V3(in + 3 * i);
  }
}
void pref_get_node_scalars(void* h, float* radius, float* invMass) {
  Solver* s = S(h);
  for (size_t i = 0; i < s->_nodes.size(); ++i) {
    radius[i] = s->_nodes[i].radius;
    invMass[i] = s->_nodes[i].invMass;
  }
}
// getVertices() mirror positions (Solver.h:65), the public readback
void pref_get_vertices(void* h, float* xyz) {
  const std::vector<Solver::Vertex>& v = S(h)->getVertices();
  for (size_t i = 0; i < v.size(); ++i) {
    xyz[3 * i] = v[i].position.x; xyz[3 * i + 1] = v[i].position.y; xyz[3 * i + 2] = v[i].position.z;
  }
}
void pref_get_triangles(void* h, uint32_t* out) {
  const std::vector<Triangle>& t = S(h)->getTriangles();
  for (size_t i = 0; i < t.size(); ++i)
    for (int k = 0; k < 3; ++k) out[3 * i + k] = t[i].nodeIds[k];
}
void pref_get_lines(void* h, uint32_t* out) {
  const std::vector<uint32_t>& l = S(h)->getLines();
  std::copy(l.begin(), l.end(), out);
}
// glm::mat3 is column-major: out[3*c + r] = m[c][r]
static void copyMat3(const glm::mat3& m, float* out) {
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) out[3 * c + r] = m[c][r];
}
void pref_get_tets(void* h, uint32_t* ids, float* qinv, float* w, float* minS, float* maxS) {
  Solver* s = S(h);
  for (size_t i = 0; i < s->_tetConstraints.size(); ++i) {
    TetrahedralConstraint& c = s->_tetConstraints[i];
    for (int k = 0; k < 4; ++k) ids[4 * i + k] = c._nodeIds[k];
    copyMat3(c._projection.Qinv, qinv + 9 * i);
    w[i] = c._w; minS[i] = c._projection.minStrain; maxS[i] = c._projection.maxStrain;
  }
}
void pref_get_volumes(void* h, uint32_t* ids, float* qinv, float* w, float* minO, float* maxO) {
  Solver* s = S(h);
  for (size_t i = 0; i < s->_volumeConstraints.size(); ++i) {
    VolumeConstraint& c = s->_volumeConstraints[i];
    for (int k = 0; k < 4; ++k) ids[4 * i + k] = c._nodeIds[k];
    copyMat3(c._projection.Qinv, qinv + 9 * i);
    w[i] = c._w; minO[i] = c._projection.minOmega; maxO[i] = c._projection.maxOmega;
  }
}
void pref_get_distances(void* h, uint32_t* ids, float* rest, float* w) {
  Solver* s = S(h);
  for (size_t i = 0; i < s->_distanceConstraints.size(); ++i) {
    DistanceConstraint& c = s->_distanceConstraints[i];
    ids[2 * i] = c._nodeIds[0]; ids[2 * i + 1] = c._nodeIds[1];
    rest[i] = c._projection.targetDistance; w[i] = c._w;
  }
}
void pref_get_positions_c(void* h, uint32_t* ids, float* target, float* w) {
  Solver* s = S(h);
  for (size_t i = 0; i < s->_positionConstraints.size(); ++i) {
    PositionConstraint& c = s->_positionConstraints[i];
    ids[i] = c._nodeIds[0];
    target[3 * i] = c._projection.fixedPosition.x;
    target[3 * i + 1] = c._projection.fixedPosition.y;
    target[3 * i + 2] = c._projection.fixedPosition.z;
    w[i] = c._w;
  }
}
void pref_get_bends(void* h, uint32_t* ids, float* angle, float* w) {
  Solver* s = S(h);
  for (size_t i = 0; i < s->_bendConstraints.size(); ++i) {
    BendConstraint& c = s->_bendConstraints[i];
    for (int k = 0; k < 4; ++k) ids[4 * i + k] = c._nodeIds[k];
    angle[i] = c._projection.initialAngle; w[i] = c._w;
  }
}
uint32_t pref_shape_size(void* h, uint32_t i) { return (uint32_t)S(h)->_shapeConstraints[i]._nodeIndices.size(); }
void pref_get_shape(void* h, uint32_t i, uint32_t* ids, double* material /*3*n col-major*/, double* qinv /*9 col-major*/, float* w) {
  ShapeMatchingConstraint& c = S(h)->_shapeConstraints[i];
  size_t n = c._nodeIndices.size();
  for (size_t k = 0; k < n; ++k) {
    ids[k] = c._nodeIndices[k];
    for (int r = 0; r < 3; ++r) material[3 * k + r] = c._materialCoords(r, k);
  }
  for (int cc = 0; cc < 3; ++cc) for (int r = 0; r < 3; ++r) qinv[3 * cc + r] = c._Qinv(r, cc);
  *w = c._w;
}
uint32_t pref_goal_size(void* h, uint32_t i) { return (uint32_t)S(h)->_goalConstraints[i]._nodeIndices.size(); }
void pref_get_goal(void* h, uint32_t i, uint32_t* ids, float* material, float* w) {
  GoalMatchingConstraint& c = S(h)->_goalConstraints[i];
  for (size_t k = 0; k < c._nodeIndices.size(); ++k) {
    ids[k] = c._nodeIndices[k];
    material[3 * k] = c._materialCoords[k].x; material[3 * k + 1] = c._materialCoords[k].y; material[3 * k + 2] = c._materialCoords[k].z;
  }
  *w = c._w;
}

// collision lists as left by the last detection (Solver.h:186-189)
uint32_t pref_tri_collision_count(void* h) { return (uint32_t)S(h)->_triCollisions.size(); }
uint32_t pref_static_collision_count(void* h) { return (uint32_t)S(h)->_staticCollisions.size(); }
void pref_get_tri_collisions(void* h, uint32_t* ids) {
  Solver* s = S(h);
  for (size_t i = 0; i < s->_triCollisions.size(); ++i)
    for (int k = 0; k < 4; ++k) ids[4 * i + k] = s->_triCollisions[i].nodeIds[k];
}
void pref_get_static_collisions(void* h, uint32_t* ids) {
  Solver* s = S(h);
  for (size_t i = 0; i < s->_staticCollisions.size(); ++i) ids[i] = s->_staticCollisions[i].nodeId;
}
int64_t pref_stiffness_nnz(void* h) { return (int64_t)S(h)->_stiffnessMatrix.nonZeros(); }
// The system of the last global step, white-box: S + C_t in triplet form, the last right-hand side
// (_forceVector, N x 3 column-major) and the solver's answer (_stateVector).  Used to measure the
// reference's own fp32 solve error against an fp64 solve of the same float system.
int64_t pref_system_nnz(void* h) { return (int64_t)S(h)->_stiffnessAndCollisionMatrix.nonZeros(); }
void pref_get_system(void* h, int32_t* rows, int32_t* cols, float* vals, float* rhs, float* state) {
  Solver* s = S(h);
  const Eigen::SparseMatrix<float>& m = s->_stiffnessAndCollisionMatrix;
  int64_t k = 0;
  for (int c = 0; c < m.outerSize(); ++c)
    for (Eigen::SparseMatrix<float>::InnerIterator it(m, c); it; ++it) { rows[k] = (int32_t)it.row(); cols[k] = (int32_t)it.col(); vals[k] = it.value(); ++k; }
  size_t n = s->_nodes.size();
  for (size_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) { rhs[3 * i + c] = s->_forceVector(i, c); state[3 * i + c] = s->_stateVector(i, c); }
}

// ---- per-function probes ----------------------------------------------------
namespace {
std::vector<Node> makeNodes(uint32_t n, const float* pos, const float* invMass) {
  std::vector<Node> nodes(n);
  for (uint32_t i = 0; i < n; ++i) {
    nodes[i].id = i;
    nodes[i].position = V3(pos + 3 * i);
    nodes[i].prevPosition = nodes[i].position;
    nodes[i].invMass = invMass ? invMass[i] : 1.0f;
  }
  return nodes;
}
glm::mat3 toMat3(const float* q) {
  glm::mat3 m;
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) m[c][r] = q[3 * c + r];
  return m;
}
} // namespace

// TetrahedralConstraintProjection::operator() (Constraints.cpp:76-128); batch of n tets,
// pos = n*12 floats, qinv = n*9 col-major, out = n*12 floats (projected[0..3]).
void pref_probe_tet(uint32_t n, const float* pos, const float* qinv, float minS, float maxS, float* out) {
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<Node> nodes = makeNodes(4, pos + 12 * i, nullptr);
    TetrahedralConstraintProjection p{glm::mat3(1.0f), toMat3(qinv + 9 * i), minS, maxS};
    std::array<glm::vec3, 4> proj;
    p(nodes, {0, 1, 2, 3}, proj);
    for (int k = 0; k < 4; ++k) { out[12 * i + 3 * k] = proj[k].x; out[12 * i + 3 * k + 1] = proj[k].y; out[12 * i + 3 * k + 2] = proj[k].z; }
  }
}
void pref_probe_volume(uint32_t n, const float* pos, const float* qinv, float minO, float maxO, float* out) {
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<Node> nodes = makeNodes(4, pos + 12 * i, nullptr);
    VolumeConstraintProjection p{toMat3(qinv + 9 * i), minO, maxO};
    std::array<glm::vec3, 4> proj;
    p(nodes, {0, 1, 2, 3}, proj);
    for (int k = 0; k < 4; ++k) { out[12 * i + 3 * k] = proj[k].x; out[12 * i + 3 * k + 1] = proj[k].y; out[12 * i + 3 * k + 2] = proj[k].z; }
  }
}
// createTetrahedralConstraint's Qinv (Constraints.cpp:151-155) for n rest tets
void pref_probe_qinv(uint32_t n, const float* pos, float* qinv) {
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<Node> nodes = makeNodes(4, pos + 12 * i, nullptr);
    TetrahedralConstraint c = createTetrahedralConstraint(0, 1.0f, nodes[0], nodes[1], nodes[2], nodes[3]);
    copyMat3(c._projection.Qinv, qinv + 9 * i);
  }
}
void pref_probe_bend(uint32_t n, const float* pos, const float* invMass, const float* angle, float* out) {
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<Node> nodes = makeNodes(4, pos + 12 * i, invMass + 4 * i);
    BendConstraintProjection p{angle[i]};
    std::array<glm::vec3, 4> proj;
    p(nodes, {0, 1, 2, 3}, proj);
    for (int k = 0; k < 4; ++k) { out[12 * i + 3 * k] = proj[k].x; out[12 * i + 3 * k + 1] = proj[k].y; out[12 * i + 3 * k + 2] = proj[k].z; }
  }
}
void pref_probe_bend_angle(uint32_t n, const float* pos, float* angle) {
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<Node> nodes = makeNodes(4, pos + 12 * i, nullptr);
    BendConstraint c = createBendConstraint(0, 1.0f, nodes[0], nodes[1], nodes[2], nodes[3]);
    angle[i] = c._projection.initialAngle;
  }
}
void pref_probe_distance(uint32_t n, const float* pos, const float* rest, float* out) {
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<Node> nodes = makeNodes(2, pos + 6 * i, nullptr);
    DistanceConstraintProjection p{rest[i]};
    std::array<glm::vec3, 2> proj;
    p(nodes, {0, 1}, proj);
    for (int k = 0; k < 2; ++k) { out[6 * i + 3 * k] = proj[k].x; out[6 * i + 3 * k + 1] = proj[k].y; out[6 * i + 3 * k + 2] = proj[k].z; }
  }
}
// pointTriangleCCD (CollisionDetection.cpp:227-302); in = n*18 floats (ap0 ab0 ac0 ap1 ab1 ac1)
void pref_probe_ccd(uint32_t n, const float* in, float threshold, int32_t* hit, float* t) {
  for (uint32_t i = 0; i < n; ++i) {
    const float* p = in + 18 * i;
    std::optional<float> r = CollisionDetection::pointTriangleCCD(
        V3(p), V3(p + 3), V3(p + 6), V3(p + 9), V3(p + 12), V3(p + 15), threshold);
    hit[i] = r ? 1 : 0;
    t[i] = r ? *r : -1.0f;
  }
}
// edgeEdgeCCD (CollisionDetection.cpp:304-418); in = n*18 floats (ab0 ac0 ad0 ab1 ac1 ad1)
void pref_probe_edge_ccd(uint32_t n, const float* in, int32_t* hit, float* t) {
  for (uint32_t i = 0; i < n; ++i) {
    const float* p = in + 18 * i;
    std::optional<float> r = CollisionDetection::edgeEdgeCCD(V3(p), V3(p + 3), V3(p + 6), V3(p + 9), V3(p + 12), V3(p + 15));
    hit[i] = r ? 1 : 0;
    t[i] = r ? *r : -1.0f;
  }
}
// Cell ranges: NodeCompRange (Solver.cpp:877-901) / TriCompRange (:942-979) / sweptTriRange
// (:639-677, re-derived here through TriCompRange + the 20-cell cap because it is file-static).
void pref_probe_node_range(uint32_t n, const float* pos, const float* radius, float gridScale,
                           int64_t* mins, uint32_t* lens) {
  SpatialHashGrid grid{gridScale};
  Solver::NodeCompRange f;
  for (uint32_t i = 0; i < n; ++i) {
    Node node; node.position = V3(pos + 3 * i); node.radius = radius[i];
    SpatialHashGridCellRange r = f(node, grid);
    mins[3 * i] = r.minX; mins[3 * i + 1] = r.minY; mins[3 * i + 2] = r.minZ;
    lens[3 * i] = r.lengthX; lens[3 * i + 1] = r.lengthY; lens[3 * i + 2] = r.lengthZ;
  }
}
void pref_probe_tri_range(uint32_t n, const float* pos /*9 per tri*/, const float* prev /*9 per tri*/,
                          int64_t* mins, uint32_t* lens) {
  SpatialHashGrid grid{2.0f};
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<Node> nodes(3);
    for (int k = 0; k < 3; ++k) { nodes[k].position = V3(pos + 9 * i + 3 * k); nodes[k].prevPosition = V3(prev + 9 * i + 3 * k); }
    Solver::TriCompRange f{nodes};
    Triangle tri{{0, 1, 2}};
    SpatialHashGridCellRange r = f(tri, grid);
    mins[3 * i] = r.minX; mins[3 * i + 1] = r.minY; mins[3 * i + 2] = r.minZ;
    lens[3 * i] = r.lengthX; lens[3 * i + 1] = r.lengthY; lens[3 * i + 2] = r.lengthZ;
  }
}

// Triangle-hash occupancy of the solver's current state, through a private
// SpatialHash instance (the solver's own is being cleared asynchronously,
// Solver.cpp:848-849).  Two calls: count, then fill.  Cells are returned sorted
// by (x,y,z); members in bucket order (ascending triangle index).
namespace {
struct Occ { int64_t x, y, z; std::vector<uint32_t> members; };
std::vector<Occ> g_occ;
}
uint64_t pref_tri_occupancy_build(void* h, uint64_t* totalMembers) {
  Solver* s = S(h);
  SpatialHash<Triangle, Solver::TriCompRange> hash(s->_options.gridSpacing);
  hash.parallelBulkInsert(s->_triangles, {s->_nodes});
  g_occ.clear();
  uint64_t total = 0;
  for (auto& kv : hash._hashMap) {
    Occ o{kv.first.x, kv.first.y, kv.first.z, {}};
    for (Triangle* t : kv.second.values) o.members.push_back((uint32_t)(t - s->_triangles.data()));
    total += o.members.size();
    g_occ.push_back(std::move(o));
  }
  std::sort(g_occ.begin(), g_occ.end(), [](const Occ& a, const Occ& b) {
    return std::tie(a.x, a.y, a.z) < std::tie(b.x, b.y, b.z);
  });
  *totalMembers = total;
  return g_occ.size();
}
void pref_tri_occupancy_get(int64_t* cells /*3 per cell*/, uint32_t* counts, uint32_t* members) {
  size_t m = 0;
  for (size_t i = 0; i < g_occ.size(); ++i) {
    cells[3 * i] = g_occ[i].x; cells[3 * i + 1] = g_occ[i].y; cells[3 * i + 2] = g_occ[i].z;
    counts[i] = (uint32_t)g_occ[i].members.size();
    for (uint32_t v : g_occ[i].members) members[m++] = v;
  }
}
uint64_t pref_node_occupancy_build(void* h, uint64_t* totalMembers) {
  Solver* s = S(h);
  SpatialHash<Node, Solver::NodeCompRange> hash(s->_options.gridSpacing);
  hash.parallelBulkInsert(s->_nodes, {});
  g_occ.clear();
  uint64_t total = 0;
  for (auto& kv : hash._hashMap) {
    Occ o{kv.first.x, kv.first.y, kv.first.z, {}};
    for (Node* t : kv.second.values) o.members.push_back((uint32_t)(t - s->_nodes.data()));
    total += o.members.size();
    g_occ.push_back(std::move(o));
  }
  std::sort(g_occ.begin(), g_occ.end(), [](const Occ& a, const Occ& b) {
    return std::tie(a.x, a.y, a.z) < std::tie(b.x, b.y, b.z);
  });
  *totalMembers = total;
  return g_occ.size();
}

} // extern "C"
