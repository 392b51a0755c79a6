// scene.h — host-side scene tables of the B200 solver (plain C++, no glm).
//
// Scene construction is setup-time work that stays on the host, like in the
// reference (Src/PrimitiveUtilities.cpp); it must however reproduce the
// reference's node ids, constraint order and rest data exactly, because the
// per-timestep kernels consume those tables.  Each builder cites the reference
// lines it follows.  The tables are SoA so they upload to HBM without repacking.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pies_b200.h"

namespace pies {

struct Mat4 { float m[16]; };  // column-major, glm::mat4 memory layout

struct HostScene {
  // nodes (reference Include/Pies/Node.h:8-20), SoA
  std::vector<float> pos, prev, vel;  // 3 per node
  std::vector<float> radius, invMass;

  // Constraint<1,PositionConstraintProjection> (Constraints.h:159-169)
  std::vector<uint32_t> posId; std::vector<float> posTarget; std::vector<float> posW;
  // Constraint<2,DistanceConstraintProjection> (Constraints.h:147-157)
  std::vector<uint32_t> distId; std::vector<float> distRest; std::vector<float> distW;
  // TetrahedralConstraint (Constraints.h:171-192): ids, Qinv (glm column-major), w, strain limits
  std::vector<uint32_t> tetId; std::vector<float> tetQinv, tetW, tetMin, tetMax;
  // VolumeConstraint (Constraints.h:194-213)
  std::vector<uint32_t> volId; std::vector<float> volQinv, volW, volMin, volMax;
  // BendConstraint (Constraints.h:215-230)
  std::vector<uint32_t> bendId; std::vector<float> bendAngle, bendW;
  // ShapeMatchingConstraint (ShapeMatchingConstraint.h:15-37): CSR over clusters
  std::vector<uint32_t> shapeOff{0}, shapeId; std::vector<double> shapeMat /*3 per member, centred*/;
  std::vector<double> shapeQinv /*9 per cluster, column-major*/; std::vector<float> shapeW;
  std::vector<double> shapeQuat;  // warm-started rotation per cluster (x,y,z,w), SURVEY F11
  // GoalMatchingConstraint (ShapeMatchingConstraint.h:39-59)
  std::vector<uint32_t> goalOff{0}, goalId; std::vector<float> goalMat /*3 per member*/;
  std::vector<Mat4> goalXform; std::vector<float> goalW;
  // FixedRegion (Solver.h:141-145)
  struct FixedRegion { Mat4 initial, invInitial; uint32_t goal; };
  std::vector<FixedRegion> fixedRegions;

  // render mirrors (Solver.h:192-194)
  std::vector<uint32_t> triangles;  // 3 per triangle
  std::vector<uint32_t> lines;
  std::vector<PiesB200Vertex> vertices;

  uint32_t constraintId = 0;  // Solver::_constraintId (Solver.h:138)
  uint64_t topologyVersion = 0;  // bumped by every mutation; the device side rebuilds when it changes
  bool goalXformDirty = false;

  uint32_t nodeCount() const { return (uint32_t)radius.size(); }
  uint32_t triCount() const { return (uint32_t)(triangles.size() / 3); }

  // ---- reference factories ----
  void addNodes(uint32_t n, const float* xyz);
  void createBox(const float t[3], float scale, float w);
  void createTetBox(const float t[3], float scale, const float v0[3], float w, float mass, bool hinged);
  void createSheet(const float t[3], float scale, float mass, float k);
  void createShapeMatchingBox(const float t[3], uint32_t cx, uint32_t cy, uint32_t cz, float scale,
                              const float v0[3], float w);
  void createShapeMatchingSheet(const float t[3], float scale, const float v0[3], float w);
  void createBendSheet(const float t[3], float scale, float w);
  void addTetMeshVolume(uint32_t nPoints, const float* xyz, uint32_t nTets, const uint32_t* tetIdx,
                        uint32_t nTris, const uint32_t* triIdx, const float v0[3], float density,
                        float strainStiffness, float minStrain, float maxStrain, float volumeStiffness,
                        float compression, float stretching);
  void addFixedRegions(uint32_t n, const float* mats, float w);
  bool updateFixedRegions(uint32_t n, const float* mats);
  void addLinkedRegions(uint32_t n, const float* mats, float w);
  void clear();

  // ---- additive builders (Constraints.h:155-230 factories, bulk) ----
  uint32_t appendNode(const float p[3], const float v[3], float radius, float invMass);
  void appendDistance(uint32_t a, uint32_t b, float w);
  void appendPosition(uint32_t a, float w);
  void appendTet(const uint32_t ids[4], float w, float minStrain, float maxStrain);
  void appendVolume(const uint32_t ids[4], float w, float compression, float stretching);
  void appendBend(const uint32_t ids[4], float w);
  void appendShape(uint32_t n, const uint32_t* ids, const float* materialXYZ, float w);
  void appendGoal(uint32_t n, const uint32_t* ids, float w);

 private:
  struct Cosmetics { float color[3]; float roughness; float metallic; };
  Cosmetics rollCosmetics();
  void syncVertices(size_t first, const Cosmetics& c);
  void tetQinvOf(const uint32_t ids[4], float out[9]) const;
};

}  // namespace pies
