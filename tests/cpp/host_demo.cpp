// host_demo.cpp — a minimal host application written ONLY against the reference's public API
// (Pies::Solver, reference Include/Pies/Solver.h:40-116).  tests/cpp/Makefile compiles this one
// source twice: against the reference's header + objects (host_demo_ref, CPU) and against this
// repo's Include/Pies/Solver.h + libpies_b200.so (host_demo_b200, B200).  The drop-in claim is
// that nothing in this file changes between the two builds; tests/test_cpp_dropin.py runs both
// and compares the vertex positions they write.
//
// usage: host_demo <scene> <ticks> <out.bin>     scene in {twobox, sheet, boxes, shapes, tetcube}
#include <Pies/Solver.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

// closed triangulated surface of an axis-aligned cube, n x n quads per face, shared vertices: the input of
// Solver::addTriMeshVolume (reference Solver.h:77-87), which tetrahedralises it with TetGen on both builds
static void cubeSurface(float side, int n, glm::vec3 origin, std::vector<glm::vec3>& verts, std::vector<uint32_t>& tris) {
  std::map<std::tuple<int, int, int>, uint32_t> index;
  auto vid = [&](int i, int j, int k) {
    auto key = std::make_tuple(i, j, k);
    auto it = index.find(key);
    if (it != index.end()) return it->second;
    uint32_t id = (uint32_t)verts.size();
    verts.push_back(origin + glm::vec3(side * i / n, side * j / n, side * k / n));
    index.emplace(key, id);
    return id;
  };
  auto quad = [&](uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    tris.insert(tris.end(), {a, b, c, a, c, d});
  };
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b)
      for (int f = 0; f <= n; f += n) {
        quad(vid(f, a, b), vid(f, a + 1, b), vid(f, a + 1, b + 1), vid(f, a, b + 1));
        quad(vid(a, f, b), vid(a + 1, f, b), vid(a + 1, f, b + 1), vid(a, f, b + 1));
        quad(vid(a, b, f), vid(a + 1, b, f), vid(a + 1, b + 1, f), vid(a, b + 1, f));
      }
}

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s <scene> <ticks> <out.bin>\n", argv[0]); return 2; }
  const char* scene = argv[1];
  int ticks = std::atoi(argv[2]);
  std::srand(1);
  Pies::SolverOptions opt;
  opt.iterations = 10;
  if (!std::strcmp(scene, "boxes")) opt.solver = Pies::SolverName::PBD;
  Pies::Solver solver(opt);
  if (!std::strcmp(scene, "twobox")) {
    solver.createTetBox(glm::vec3(0.1f, 0.3f, 0.1f), 1.0f, glm::vec3(0.0f), 1000.0f, 1.0f, false);
    solver.createTetBox(glm::vec3(0.4f, 2.6f, 0.3f), 1.0f, glm::vec3(0.0f, -5.0f, 0.0f), 1000.0f, 1.0f, false);
  } else if (!std::strcmp(scene, "sheet")) {
    solver.createSheet(glm::vec3(0.0f, 4.0f, 0.0f), 1.0f, 1.0f, 100.0f);
    solver.createBendSheet(glm::vec3(30.0f, 4.0f, 0.0f), 1.0f, 100.0f);  // clear of the 20-wide sheet: coplanar overlap is degenerate
  } else if (!std::strcmp(scene, "boxes")) {
    for (int i = 0; i < 8; ++i)
      solver.createBox(glm::vec3(6.0f * (i % 2), 3.0f + 6.0f * (i / 4), 6.0f * ((i / 2) % 2)), 1.0f, 0.5f);
  } else if (!std::strcmp(scene, "shapes")) {
    solver.createShapeMatchingBox(glm::vec3(0.0f, 3.0f, 0.0f), 3, 3, 3, 1.0f, glm::vec3(0.0f), 1000.0f);
    solver.createShapeMatchingBox(glm::vec3(4.0f, 2.0f, 1.0f), 4, 3, 5, 1.0f, glm::vec3(0.0f), 800.0f);
  } else if (!std::strcmp(scene, "tetcube")) {
    std::vector<glm::vec3> verts;
    std::vector<uint32_t> tris;
    cubeSurface(4.0f, 4, glm::vec3(0.0f, 1.07f, 0.0f), verts, tris);
    solver.addTriMeshVolume(verts, tris, glm::vec3(0.0f), 1.0f, 1000.0f, 0.8f, 1.0f, 1000.0f, 1.0f, 1.0f);
  } else {
    std::fprintf(stderr, "unknown scene %s\n", scene);
    return 2;
  }
  for (int t = 0; t < ticks; ++t) solver.tick(0.012f);
#ifdef PIES_B200_H  // only the B200 build has an error channel (additive API)
  if (solver.failed() || !solver.lastError().empty()) std::fprintf(stderr, "pies_b200: %s\n", solver.lastError().c_str());
#endif
  const std::vector<Pies::Solver::Vertex>& v = solver.getVertices();
  std::FILE* f = std::fopen(argv[3], "wb");
  if (!f) return 3;
  uint32_t n = (uint32_t)v.size(), nt = (uint32_t)solver.getTriangles().size(), nl = (uint32_t)solver.getLines().size();
  std::fwrite(&n, 4, 1, f); std::fwrite(&nt, 4, 1, f); std::fwrite(&nl, 4, 1, f);
  for (const auto& x : v) { std::fwrite(&x.position, 12, 1, f); std::fwrite(&x.radius, 4, 1, f); }
  if (nt) std::fwrite(solver.getTriangles().data(), 12, nt, f);
  if (nl) std::fwrite(solver.getLines().data(), 4, nl, f);
  std::fclose(f);
  std::printf("%s: %u vertices, %u triangles, %u line indices after %d ticks\n", scene, n, nt, nl, ticks);
  return 0;
}
