// system.h — once-per-topology host precomputation for the PD global/local steps.
#pragma once

#include <cstdint>
#include <vector>

#include "scene.h"

namespace pies {

struct HostSystem {
  uint32_t n = 0;

  // fused tet elements (strain + volume on the same tet share one record), plane layout of kernels.h
  uint32_t nElems = 0;
  std::vector<uint32_t> elemIds;              // 4 per element
  std::vector<float> elemQa, elemQb, elemPc, elemPd;  // 4 floats per element each

  // contribution slots (one float4 each) and the per-node gather lists
  uint64_t baseTet = 0, baseDist = 0, baseBend = 0, baseShape = 0, baseGoal = 0, basePos = 0, nContrib = 0;
  std::vector<float> posContrib;              // constant w * fixedPosition, 4 floats per position constraint
  std::vector<int> incPtr;                    // n + 1
  std::vector<uint32_t> inc;

  // S = M/h^2 + sum w A^T A  (reference Solver.cpp:174-210), CSR with both triangles
  std::vector<int> rowPtr, col;
  std::vector<float> val;

  // Row batches for the CG mat-vec (CSR-stream): consecutive rows whose non-zeros fit one shared-memory
  // tile (<= kBatchNnz entries, <= kBatchRows rows).  A CTA streams a batch's (col, val) pairs fully coalesced,
  // parks the products in shared memory and then sums them per row in CSR order.
  std::vector<uint32_t> rowBatch;             // nBatches + 1 first rows
  static constexpr uint32_t kBatchNnz = 2048, kBatchRows = 256;

  // block-Jacobi preconditioner
  uint32_t nBlocks = 0;
  std::vector<int> blockNodes;                // 32 per block, -1 padded
  std::vector<float> blockInv;                // 1024 per block, symmetric

  uint64_t staticProjections = 0;             // per PD iteration, shape/goal count one per member
};

// h = fixedTimestepSize / timeSubsteps.  threads: worker threads for the row-wise assembly.
void buildSystem(const HostScene& sc, float h, HostSystem& out, unsigned threads);

}  // namespace pies
