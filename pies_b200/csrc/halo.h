// halo.h — slab-partitioned multi-GPU runs: ghost rows of the node state exchanged with NCCL from inside the tick (halo.cu).
#pragma once

#include <vector>

#include "engine.h"

namespace pies {

struct HaloWork {
  void* comm = nullptr;          // ncclComm_t
  int rank = 0, world = 1;
  // exchange lists: for peer k, rows sendIdx[sendOff[k] .. sendOff[k+1]) of this rank's state are ghosts there, and rows
  // recvIdx[recvOff[k] .. recvOff[k+1]) here are ghosts owned by it; both sides enumerate shared bodies in global order
  std::vector<int> peers;
  std::vector<uint32_t> sendOff, recvOff;
  DevBuf<uint32_t> sendIdx, recvIdx;
  DevBuf<float4> sendBuf, recvBuf;   // three planes per row, peer segments back to back
  DevBuf<int> status;                // [0] this rank's failure flag for the all-reduce, [1] the reduced flag,
                                     // [2] / [3] the same for "an island mixes owned and ghost rows"
  DevBuf<uint8_t> ghost;             // per local row: 1 = ghost (owned by a neighbour)
  DevBuf<uint32_t> bodyMix;          // per body (island root): bit 0 an owned row, bit 1 a ghost row in its island
  int* hostStatus = nullptr;         // pinned
  uint64_t bytesLastTick = 0, exchangesLastTick = 0;
  bool listsSet = false;
  ~HaloWork();
};

// 128 bytes: ncclGetUniqueId (rank 0; the host hands it to the other ranks, e.g. over its own rendezvous)
int haloUniqueId(void* out128, std::string* err);
int haloInit(PiesB200Solver* s, int rank, int world, const void* id128);
int haloSetLists(PiesB200Solver* s, int nPeers, const int* peers, const uint32_t* sendCounts, const uint32_t* sendIdx,
                 const uint32_t* recvCounts, const uint32_t* recvIdx);
// Overwrites the ghost rows with their owners' rows: planes = 1 (positions) or 3 (positions, previous positions, velocities).
int haloExchange(PiesB200Solver* s, int planes);
// Collective over all ranks: true if any rank's simulation has failed (so they all stop together instead of hanging).
int haloAnyFailed(PiesB200Solver* s, bool mine, bool* any);
// Collective, once per substep after the islands are built: true if on any rank an island (a connected component of
// S + C_t) contains both owned and ghost rows.  If none does, the owned rows of every rank are independent of the ghost
// rows for the rest of the substep (ghost bodies are complete islands of their own, simulated redundantly) and the
// per-iteration exchanges can be skipped: the next substep's three-plane exchange refreshes the ghosts.  localOk = false
// (this rank's substep failed) counts as mixed.
int haloMixedIslands(PiesB200Solver* s, bool localOk, bool* anyMixed);
void haloDestroy(PiesB200Solver* s);

}  // namespace pies
