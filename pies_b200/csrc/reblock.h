// reblock.h — per-substep contact-aware preconditioner blocks (reblock.cu).
#pragma once

#include <algorithm>

#include "engine.h"

namespace pies {

constexpr uint32_t kMidClusterMax = 1024;  // nodes of a contact cluster one warp can stage in shared memory (2 x 16 B each)

struct BlockWork {
  DevBuf<int> blockNodes;        // (static + dynamic blocks) * 32, members packed to the front, -1 padded
  DevBuf<float> blockInv;        // packed lower triangle per block (m = members), at blockOff[b]
  DevBuf<uint32_t> blockCount, blockOff;
  DevBuf<uint2> blockMeta;       // (blockOff, m) per block, what the CG kernels read
  DevBuf<uint32_t> slotOf;       // node -> block * 32 + lane
  DevBuf<uint32_t> flag, parent, vals, tmpVals, heads, start, blkOff, sortHist, scanScratch, nBlocksDev;
  DevBuf<uint64_t> keys, tmpKeys;
  DevBuf<uint8_t> dirty, gsClass;  // gsClass[node]: 0 not in a contact, 1 small cluster (<= 32 nodes), 3 mid cluster (<= kMidClusterMax), 2 large cluster
  DevBuf<uint32_t> clusterOf;      // node -> contact cluster (touched nodes only)
  DevBuf<uint32_t> rankOf;         // node -> position inside its cluster (touched nodes only)
  // clusters of 33 .. kMidClusterMax nodes (swept by one warp from shared memory, contact.cu): list + counters
  // [0] = mid clusters, [1] = clusters larger than that (dataflow sweeps); host copy in host[1..2] once countsReady fired
  DevBuf<uint32_t> midList, midCount;
  cudaEvent_t countsReady = nullptr;
  // the dense inversions of the touched blocks (k_block_factor) run on a side stream beside the cluster and island
  // bookkeeping that follows (dozens of tiny launches); whoever applies the inverses first waits for factorDone
  cudaStream_t side = nullptr;
  cudaEvent_t factorFork = nullptr, factorDone = nullptr;
  bool factorPending = false;
  ~BlockWork() {
    if (countsReady) cudaEventDestroy(countsReady);
    if (factorFork) cudaEventDestroy(factorFork);
    if (factorDone) cudaEventDestroy(factorDone);
    if (side) cudaStreamDestroy(side);
  }
  uint32_t midClusterMax = kMidClusterMax;  // tuning: 0 sends every cluster above 32 nodes to the dataflow sweeps
  uint32_t* host = nullptr;      // pinned, 4 words
  uint32_t nBlocksBound = 0, nTouched = 0;
  // contact clusters of this substep: nodes sorted by cluster in vals[start[c] .. start[c+1]), count at heads[nTouched]
  uint64_t scanCap = 0;
  cudaError_t lastError = cudaSuccess;
  // what rebuildBlocks() last pointed the CG at (kept so the per-iteration phases of a tick can rebuild their views)
  struct Current { int* blockNodes = nullptr; float* blockInv = nullptr; const uint2* blockMeta = nullptr;
                   uint32_t nBlocks = 0; const uint32_t* nBlocksDev = nullptr; } cur;
};

// Regroups the nodes joined by this substep's contacts into dense blocks and re-inverts every block the
// collision terms touch; points pw.blockNodes / blockInv / nBlocksDev at the result.  Returns 0 or -1.
int rebuildBlocks(BlockWork& w, cudaStream_t s, uint32_t n, const CsrMatrix& S, const int* staticNodes,
                  const float* staticInv, uint32_t nStatic, const ContactLists& c, const float4* q, PcgWork& pw,
                  int* launches);

}  // namespace pies
