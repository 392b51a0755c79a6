// Include/Pies/Solver.h — drop-in replacement of the reference's public header
// (nithinp7/Pies Include/Pies/Solver.h:21-199): same namespace, type names, member
// names, signatures and defaults, so a host application written against the reference
// (its Vulkan viewer / Maya plugin) recompiles unchanged and links libpies_b200.so.
//
// The class is header-only and owns nothing but an opaque C-ABI handle
// (include/pies_b200.h) plus the host mirrors the reference hands out by const
// reference (_vertices/_lines/_triangles, Solver.h:65-69).  All simulation state lives
// in HBM behind the handle; there is no CPU solver behind this class — if no CUDA
// device is usable every call is a no-op with `failed()` set and `lastError()` telling
// why (the reference has no error channel either: its only failure mode is the silent
// _simFailed latch, Solver.cpp:26-28).
//
// Requirements on the host side: glm (the reference vendors Extern/glm; its types are
// in these signatures) and, only for addTriMeshVolume, TetGen's <tetgen.h> built with
// -DTETLIBRARY (the reference vendors Extern/tetgen; meshing is setup-time host work,
// PrimitiveUtilities.cpp:183-241, and stays the host's).
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <glm/glm.hpp>

#include "../../include/pies_b200.h"

#if defined(__has_include)
#if __has_include(<tetgen.h>) && defined(TETLIBRARY)
#include <tetgen.h>
#define PIES_B200_HAVE_TETGEN 1
#endif
#endif

namespace Pies {

struct Triangle {  // reference Include/Pies/Triangle.h
  uint32_t nodeIds[3];
};
struct Tetrahedron {  // reference Include/Pies/Tetrahedron.h
  uint32_t nodeIds[4];
};

enum class SolverName { PBD, PD };  // Solver.h:21

struct SolverOptions {  // Solver.h:23-38, same defaults
  float fixedTimestepSize = 0.012f;
  uint32_t timeSubsteps = 1;
  uint32_t iterations = 4;
  uint32_t collisionStabilizationIterations = 4;
  float collisionThresholdDistance = 0.1f;
  float collisionThickness = 0.05f;
  float gravity = 10.0f;
  float damping = 0.006f;
  float friction = 0.01f;
  float staticFrictionThreshold = 0.f;
  float floorHeight = 0.0f;
  float gridSpacing = 2.0f;
  uint32_t threadCount = 8;
  SolverName solver = SolverName::PD;
};

class Solver {
public:
  struct Vertex {  // Solver.h:42-49, 36 bytes, bit-compatible with PiesB200Vertex
    glm::vec3 position{};
    float radius{};
    glm::vec3 baseColor{};
    float roughness{};
    float metallic{};
  };
  static_assert(sizeof(Vertex) == sizeof(PiesB200Vertex), "Vertex layout");
  static_assert(sizeof(Triangle) == 3 * sizeof(uint32_t), "Triangle layout");

  bool renderStateDirty = true;  // Solver.h:51
  bool releaseHinge = false;     // Solver.h:52 (PBD only, Solver.cpp:59)

  // The reference's default constructor leaves its per-thread scratch empty and crashes on the first PD
  // tick; here Solver() == Solver(SolverOptions{}).
  Solver() : Solver(SolverOptions{}) {}
  Solver(const SolverOptions& options, int device = -1) : _options(options) {  // not explicit, like the reference (Solver.h:55)
    PiesB200Options o;
    pies_b200_default_options(&o);
    o.fixedTimestepSize = options.fixedTimestepSize;
    o.timeSubsteps = options.timeSubsteps;
    o.iterations = options.iterations;
    o.collisionStabilizationIterations = options.collisionStabilizationIterations;
    o.collisionThresholdDistance = options.collisionThresholdDistance;
    o.collisionThickness = options.collisionThickness;
    o.gravity = options.gravity;
    o.damping = options.damping;
    o.friction = options.friction;
    o.staticFrictionThreshold = options.staticFrictionThreshold;
    o.floorHeight = options.floorHeight;
    o.gridSpacing = options.gridSpacing;
    o.threadCount = options.threadCount;
    o.solver = options.solver == SolverName::PD ? 1u : 0u;
    if (pies_b200_create(&o, device, &_h) != PIES_B200_OK) {
      _h = nullptr;
      _error = pies_b200_last_error(nullptr);
    }
  }
  Solver(const Solver&) = delete;
  Solver& operator=(const Solver&) = delete;
  Solver(Solver&& rhs) noexcept { *this = std::move(rhs); }
  Solver& operator=(Solver&& rhs) noexcept {
    if (this != &rhs) {
      if (_h) pies_b200_destroy(_h);
      _h = rhs._h; rhs._h = nullptr;
      _options = rhs._options;
      _vertices = std::move(rhs._vertices); _lines = std::move(rhs._lines); _triangles = std::move(rhs._triangles);
      _error = std::move(rhs._error);
      renderStateDirty = rhs.renderStateDirty; releaseHinge = rhs.releaseHinge;
    }
    return *this;
  }
  ~Solver() { if (_h) pies_b200_destroy(_h); }

  // ---- stepping (Solver.cpp:25-38; deltaTime is ignored exactly like the reference, Solver.cpp:162) ----
  void tick(float deltaTime) { _step(&pies_b200_tick, deltaTime); }
  void tickPBD(float deltaTime) { _step(&pies_b200_tick_pbd, deltaTime); }
  void tickPD(float deltaTime) { _step(&pies_b200_tick_pd, deltaTime); }

  const std::vector<Vertex>& getVertices() const { return _vertices; }
  const std::vector<uint32_t>& getLines() const { return _lines; }
  const std::vector<Triangle>& getTriangles() const { return _triangles; }
  const SolverOptions& getOptions() const { return _options; }

  void clear() {  // Solver.cpp:488-507
    if (_ok(pies_b200_clear(_h))) _syncTopology();
    renderStateDirty = true;
  }

  // ---- mesh import (PrimitiveUtilities.cpp:42-328) ----
  void addNodes(const std::vector<glm::vec3>& vertices) {
    if (_ok(pies_b200_add_nodes(_h, (uint32_t)vertices.size(), _f(vertices)))) _syncTopology();
  }
  void addTriMeshVolume(const std::vector<glm::vec3>& vertices, const std::vector<uint32_t>& triIndices,
                        const glm::vec3& initialVelocity, float density, float strainStiffness, float minStrain,
                        float maxStrain, float volumeStiffness, float compression, float stretching) {
#ifdef PIES_B200_HAVE_TETGEN
    // Host-side meshing with the reference's TetGen switches (PrimitiveUtilities.cpp:183-241).
    tetgenio in{}, out{};
    in.numberofpoints = (int)vertices.size();
    in.pointlist = new double[vertices.size() * 3];
    for (size_t i = 0; i < vertices.size(); ++i)
      for (int k = 0; k < 3; ++k) in.pointlist[3 * i + k] = vertices[i][k];
    in.numberoffacets = (int)(triIndices.size() / 3);
    in.facetlist = new tetgenio::facet[in.numberoffacets];
    for (int i = 0; i < in.numberoffacets; ++i) {
      tetgenio::facet& f = in.facetlist[i];
      f.numberofpolygons = 1;
      f.polygonlist = new tetgenio::polygon[1];
      f.polygonlist[0].numberofvertices = 3;
      f.polygonlist[0].vertexlist = new int[3];
      for (int k = 0; k < 3; ++k) f.polygonlist[0].vertexlist[k] = (int)triIndices[3 * i + k];
      f.numberofholes = 0;
      f.holelist = nullptr;
    }
    tetgenbehavior b{};
    b.plc = 1; b.facesout = 1; b.neighout = 2; b.zeroindex = 1; b.quality = 1; b.minratio = 1.5; b.regionattrib = 1;
    tetrahedralize(&b, &in, &out);
    // boundary faces = faces with a missing neighbour tet, winding flipped (:249-267)
    std::vector<uint32_t> tris;
    for (int i = 0; i < out.numberoftrifaces; ++i) {
      if (out.face2tetlist[2 * i] >= 0 && out.face2tetlist[2 * i + 1] >= 0) continue;
      tris.push_back((uint32_t)out.trifacelist[3 * i]);
      tris.push_back((uint32_t)out.trifacelist[3 * i + 2]);
      tris.push_back((uint32_t)out.trifacelist[3 * i + 1]);
    }
    std::vector<float> pts(3 * (size_t)out.numberofpoints);
    for (size_t i = 0; i < pts.size(); ++i) pts[i] = (float)out.pointlist[i];
    std::vector<uint32_t> tets(4 * (size_t)out.numberoftetrahedra);
    for (size_t i = 0; i < tets.size(); ++i) tets[i] = (uint32_t)out.tetrahedronlist[i];
    if (_ok(pies_b200_add_tet_mesh_volume(_h, (uint32_t)out.numberofpoints, pts.data(), (uint32_t)out.numberoftetrahedra,
                                          tets.data(), (uint32_t)(tris.size() / 3), tris.data(), &initialVelocity[0],
                                          density, strainStiffness, minStrain, maxStrain, volumeStiffness, compression,
                                          stretching)))
      _syncTopology();
#else
    (void)vertices; (void)triIndices; (void)initialVelocity; (void)density; (void)strainStiffness; (void)minStrain;
    (void)maxStrain; (void)volumeStiffness; (void)compression; (void)stretching;
    _error = "addTriMeshVolume needs <tetgen.h> and -DTETLIBRARY on the host side (or call addTetMeshVolume)";
#endif
  }
  // [additive] the post-TetGen half of addTriMeshVolume for hosts that mesh elsewhere.
  void addTetMeshVolume(const std::vector<glm::vec3>& points, const std::vector<Tetrahedron>& tets,
                        const std::vector<Triangle>& boundary, const glm::vec3& initialVelocity, float density,
                        float strainStiffness, float minStrain, float maxStrain, float volumeStiffness,
                        float compression, float stretching) {
    if (_ok(pies_b200_add_tet_mesh_volume(_h, (uint32_t)points.size(), _f(points), (uint32_t)tets.size(),
                                          tets.empty() ? nullptr : tets[0].nodeIds, (uint32_t)boundary.size(),
                                          boundary.empty() ? nullptr : boundary[0].nodeIds, &initialVelocity[0], density,
                                          strainStiffness, minStrain, maxStrain, volumeStiffness, compression, stretching)))
      _syncTopology();
  }
  void addFixedRegions(const std::vector<glm::mat4>& regionMatrices, float w) {
    if (_ok(pies_b200_add_fixed_regions(_h, (uint32_t)regionMatrices.size(), _m(regionMatrices), w))) _syncTopology();
  }
  void updateFixedRegions(const std::vector<glm::mat4>& regionMatrices) {
    _ok(pies_b200_update_fixed_regions(_h, (uint32_t)regionMatrices.size(), _m(regionMatrices)));
  }
  void addLinkedRegions(const std::vector<glm::mat4>& regionMatrices, float w) {
    if (_ok(pies_b200_add_linked_regions(_h, (uint32_t)regionMatrices.size(), _m(regionMatrices), w))) _syncTopology();
  }

  // ---- primitives (PrimitiveUtilities.cpp:330-1289) ----
  void createBox(const glm::vec3& translation, float scale, float w) {
    if (_ok(pies_b200_create_box(_h, &translation[0], scale, w))) _syncTopology();
  }
  void createTetBox(const glm::vec3& translation, float scale, const glm::vec3& initialVelocity, float w, float mass,
                    bool hinged) {
    if (_ok(pies_b200_create_tet_box(_h, &translation[0], scale, &initialVelocity[0], w, mass, hinged ? 1 : 0)))
      _syncTopology();
  }
  void createSheet(const glm::vec3& translation, float scale, float mass, float k) {
    if (_ok(pies_b200_create_sheet(_h, &translation[0], scale, mass, k))) _syncTopology();
  }
  void createShapeMatchingBox(const glm::vec3& translation, uint32_t countX, uint32_t countY, uint32_t countZ,
                              float scale, const glm::vec3& initialVelocity, float w) {
    if (_ok(pies_b200_create_shape_matching_box(_h, &translation[0], countX, countY, countZ, scale, &initialVelocity[0], w)))
      _syncTopology();
  }
  void createShapeMatchingSheet(const glm::vec3& translation, float scale, const glm::vec3& initialVelocity, float w) {
    if (_ok(pies_b200_create_shape_matching_sheet(_h, &translation[0], scale, &initialVelocity[0], w))) _syncTopology();
  }
  void createBendSheet(const glm::vec3& translation, float scale, float w) {
    if (_ok(pies_b200_create_bend_sheet(_h, &translation[0], scale, w))) _syncTopology();
  }

  // ---- [additive] what the reference has no public API for ----
  bool failed() const { return !_h || pies_b200_sim_failed(_h) != 0; }  // the _simFailed latch, made visible
  const std::string& lastError() const { return _error; }
  PiesB200Solver* handle() const { return _h; }  // for the bulk builders / state access of pies_b200.h
  std::vector<glm::vec3> getVelocities() const {
    std::vector<glm::vec3> v(_vertices.size());
    if (_h && !v.empty()) pies_b200_get_velocities(_h, &v[0][0]);
    return v;
  }
  // n ticks without refreshing the vertex mirror in between (state stays in HBM).
  void tickN(uint32_t n) {
    if (!_h) return;
    pies_b200_set_release_hinge(_h, releaseHinge ? 1 : 0);
    if (_ok(pies_b200_tick_n(_h, n))) _syncPositions();
  }

private:
  static_assert(sizeof(glm::vec3) == 12 && sizeof(glm::mat4) == 64, "glm packing");
  static const float* _f(const std::vector<glm::vec3>& v) { return v.empty() ? nullptr : &v[0][0]; }
  static const float* _m(const std::vector<glm::mat4>& v) { return v.empty() ? nullptr : &v[0][0][0]; }
  bool _ok(int rc) {
    if (!_h) return false;
    if (rc == PIES_B200_OK) return true;
    _error = pies_b200_last_error(_h);
    return false;
  }
  void _step(int (*fn)(PiesB200Solver*, float), float dt) {
    if (!_h) return;
    pies_b200_set_release_hinge(_h, releaseHinge ? 1 : 0);
    if (_ok(fn(_h, dt))) _syncPositions();
  }
  // Positions change every tick (the reference refreshes _vertices[i].position per substep, Solver.cpp:157,393).
  void _syncPositions() {
    uint32_t n = pies_b200_vertex_count(_h);
    const PiesB200Vertex* src = pies_b200_get_vertices(_h);
    if (_vertices.size() != n) { _syncTopology(); return; }
    if (n && src) std::memcpy(_vertices.data(), src, (size_t)n * sizeof(Vertex));
    renderStateDirty = pies_b200_get_render_state_dirty(_h) != 0 || renderStateDirty;
  }
  void _syncTopology() {
    uint32_t n = pies_b200_vertex_count(_h);
    _vertices.resize(n);
    if (n) std::memcpy(_vertices.data(), pies_b200_get_vertices(_h), (size_t)n * sizeof(Vertex));
    uint32_t nl = pies_b200_line_index_count(_h);
    _lines.resize(nl);
    if (nl) std::memcpy(_lines.data(), pies_b200_get_lines(_h), (size_t)nl * sizeof(uint32_t));
    uint32_t nt = pies_b200_triangle_count(_h);
    _triangles.resize(nt);
    if (nt) std::memcpy(_triangles.data(), pies_b200_get_triangles(_h), (size_t)nt * sizeof(Triangle));
    renderStateDirty = true;
  }

  PiesB200Solver* _h = nullptr;
  SolverOptions _options{};
  std::vector<Vertex> _vertices;
  std::vector<uint32_t> _lines;
  std::vector<Triangle> _triangles;
  std::string _error;
};

}  // namespace Pies
