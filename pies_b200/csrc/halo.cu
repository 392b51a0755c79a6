// halo.cu — NCCL halo exchange of slab-partitioned scenes, inside libpies_b200.so.
//
// The reference is a single-process CPU library with no communication at all; this is the multi-GPU extension the
// north_star asks for: bodies are partitioned by x slab (host side: pies_b200/multigpu.py or any C++ host), every rank
// simulates its bodies plus ghost copies of the neighbouring slabs' boundary bodies, and the ghost rows of the node
// state are overwritten with their owners' values
//   * at every substep start: position, previous position, velocity (reference Solver.cpp:229-240: before detection),
//   * after every local/global iteration: position (Solver.cpp:264-365),
// by one pack kernel, one ncclGroup of ncclSend / ncclRecv with the slab neighbours, and one unpack kernel, all on the
// solver's stream.  NVSwitch gives every pair full bandwidth, payloads are a few hundred KB, so the cost is launch and
// NCCL latency, not bytes.
//
// NCCL is loaded with dlopen at the first use: a process that already holds torch's libnccl.so.2 gets that same
// instance (no second copy of the library in the process), a plain C++ host gets the system one.  No NCCL, no multi-GPU:
// haloInit fails loudly; nothing falls back to host staging.
#include "halo.h"

#include <cstdlib>

#include "islands.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace pies {

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD);  // the copy the process already holds (torch's), if any
      if (a.lib) break;
    }
    for (const char* n : names) {
      if (a.lib) break;
      a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!a.lib) { a.error = std::string("libnccl.so.2 not found: ") + dlerror(); return a; }
    auto sym = [&](const char* name) { void* p = dlsym(a.lib, name); if (!p && a.error.empty()) a.error = std::string("missing NCCL symbol ") + name; return p; };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    return a;
  }();
  return api;
}

int failNccl(PiesB200Solver* s, ncclResult_t r, const char* what) {
  char buf[384];
  std::snprintf(buf, sizeof(buf), "NCCL error %d (%s) in %s", (int)r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?", what);
  if (s) { s->err = buf; s->simFailed = true; }
  return PIES_B200_ECUDA;
}
#define NCHECK(s, expr)                                            \
  do {                                                             \
    ncclResult_t _r = (expr);                                      \
    if (_r != ncclSuccess) return failNccl((s), _r, #expr);        \
  } while (0)

// rows idx[i] of up to three planes -> buf[plane * n + i]  (peer segments are contiguous ranges of i, so a peer's slice of
// every plane is contiguous once the planes of ONE peer are laid out back to back: see segment())
__global__ void __launch_bounds__(kThreads) k_halo_pack(uint32_t n, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ segOf,
                                                        const uint32_t* __restrict__ segStart, const uint32_t* __restrict__ segCount,
                                                        int planes, const float4* __restrict__ q, const float4* __restrict__ prev,
                                                        const float4* __restrict__ vel, float4* __restrict__ buf) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t row = idx[i], sg = segOf[i];
  const uint32_t base = (uint32_t)planes * segStart[sg], cnt = segCount[sg], j = i - segStart[sg];
  buf[base + j] = q[row];
  if (planes == 3) { buf[base + cnt + j] = prev[row]; buf[base + 2u * cnt + j] = vel[row]; }
}

__global__ void __launch_bounds__(kThreads) k_halo_unpack(uint32_t n, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ segOf,
                                                          const uint32_t* __restrict__ segStart, const uint32_t* __restrict__ segCount,
                                                          int planes, float4* __restrict__ q, float4* __restrict__ prev,
                                                          float4* __restrict__ vel, const float4* __restrict__ buf) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t row = idx[i], sg = segOf[i];
  const uint32_t base = (uint32_t)planes * segStart[sg], cnt = segCount[sg], j = i - segStart[sg];
  // positions / velocities only: .w carries invMass and radius, which belong to the local copy
  float4 a = buf[base + j], o = q[row];
  q[row] = make_float4(a.x, a.y, a.z, o.w);
  if (planes == 3) {
    a = buf[base + cnt + j]; o = prev[row];
    prev[row] = make_float4(a.x, a.y, a.z, o.w);
    a = buf[base + 2u * cnt + j];
    vel[row] = make_float4(a.x, a.y, a.z, 0.0f);
  }
}
// island root of every row's body: which kinds of rows (owned / ghost) its island holds
__global__ void __launch_bounds__(kThreads) k_halo_mix_mark(uint32_t n, const uint32_t* __restrict__ bodyOf,
                                                            const uint32_t* __restrict__ parent, const uint8_t* __restrict__ ghost,
                                                            uint32_t* __restrict__ bodyMix) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = bodyOf[i], p = parent[x];
  while (p != x) { x = p; p = parent[x]; }
  atomicOr(bodyMix + x, ghost[i] ? 2u : 1u);
}
__global__ void __launch_bounds__(kThreads) k_halo_mix_any(uint32_t nB, const uint32_t* __restrict__ bodyMix, int* __restrict__ flag) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nB && bodyMix[b] == 3u) *flag = 1;
}
}  // namespace

HaloWork::~HaloWork() {
  if (comm && nccl().CommDestroy) nccl().CommDestroy(static_cast<ncclComm_t>(comm));
  if (hostStatus) cudaFreeHost(hostStatus);
}

int haloUniqueId(void* out128, std::string* err) {
  NcclApi& a = nccl();
  if (!a.error.empty() || !a.GetUniqueId) { if (err) *err = a.error; return PIES_B200_ENODEV; }
  ncclUniqueId id;
  ncclResult_t r = a.GetUniqueId(&id);
  if (r != ncclSuccess) { if (err) *err = a.GetErrorString(r); return PIES_B200_ECUDA; }
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(out128, &id, sizeof(id));
  return PIES_B200_OK;
}

int haloInit(PiesB200Solver* s, int rank, int world, const void* id128) {
  NcclApi& a = nccl();
  if (!a.error.empty()) return fail(s, PIES_B200_ENODEV, ("multi-GPU runs need NCCL: " + a.error).c_str());
  if (world < 1 || rank < 0 || rank >= world || !id128) return fail(s, PIES_B200_EINVAL, "halo_init: bad rank / world / id");
  haloDestroy(s);
  HaloWork* h = new HaloWork();
  h->rank = rank; h->world = world;
  s->halo = h;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  NCHECK(s, a.CommInitRank(&comm, world, id, rank));
  h->comm = comm;
  PIES_CHECK(s, cudaMallocHost(&h->hostStatus, 4 * sizeof(int)));
  PIES_CHECK(s, h->status.reserve(4));
  return PIES_B200_OK;
}

void haloDestroy(PiesB200Solver* s) {
  delete s->halo;
  s->halo = nullptr;
}

int haloSetLists(PiesB200Solver* s, int nPeers, const int* peers, const uint32_t* sendCounts, const uint32_t* sendIdx,
                 const uint32_t* recvCounts, const uint32_t* recvIdx) {
  HaloWork* h = s->halo;
  if (!h) return fail(s, PIES_B200_EINVAL, "halo_set_lists before halo_init");
  if (nPeers < 0 || (nPeers && (!peers || !sendCounts || !recvCounts))) return fail(s, PIES_B200_EINVAL, "halo_set_lists: null argument");
  int rc = ensureBuilt(s);
  if (rc) return rc;
  h->peers.assign(peers, peers + nPeers);
  h->sendOff.assign(nPeers + 1, 0); h->recvOff.assign(nPeers + 1, 0);
  for (int k = 0; k < nPeers; ++k) {
    if (peers[k] < 0 || peers[k] >= h->world || peers[k] == h->rank) return fail(s, PIES_B200_EINVAL, "halo_set_lists: bad peer rank");
    h->sendOff[k + 1] = h->sendOff[k] + sendCounts[k];
    h->recvOff[k + 1] = h->recvOff[k] + recvCounts[k];
  }
  const uint32_t nS = h->sendOff[nPeers], nR = h->recvOff[nPeers];
  if ((nS && !sendIdx) || (nR && !recvIdx)) return fail(s, PIES_B200_EINVAL, "halo_set_lists: null index list");
  for (uint32_t i = 0; i < nS; ++i) if (sendIdx[i] >= s->n) return fail(s, PIES_B200_EINVAL, "halo_set_lists: send row out of range");
  for (uint32_t i = 0; i < nR; ++i) if (recvIdx[i] >= s->n) return fail(s, PIES_B200_EINVAL, "halo_set_lists: receive row out of range");
  // device tables: [idx | segOf | segStart | segCount] for each direction
  auto upload = [&](DevBuf<uint32_t>& d, const uint32_t* idx, const std::vector<uint32_t>& off) -> cudaError_t {
    const uint32_t n = off[nPeers];
    std::vector<uint32_t> t(2ull * n + 2ull * nPeers + 2, 0u);
    for (int k = 0; k < nPeers; ++k) {
      for (uint32_t i = off[k]; i < off[k + 1]; ++i) { t[i] = idx[i]; t[n + i] = (uint32_t)k; }
      t[2ull * n + k] = off[k];
      t[2ull * n + nPeers + k] = off[k + 1] - off[k];
    }
    return d.upload(t.data(), t.size(), s->stream);
  };
  PIES_CHECK(s, upload(h->sendIdx, sendIdx, h->sendOff));
  PIES_CHECK(s, upload(h->recvIdx, recvIdx, h->recvOff));
  {
    std::vector<uint8_t> g(s->n, 0);
    for (uint32_t i = 0; i < nR; ++i) g[recvIdx[i]] = 1;
    PIES_CHECK(s, h->ghost.upload(g.data(), g.size(), s->stream));
  }
  PIES_CHECK(s, h->sendBuf.reserve(3ull * nS + 1));
  PIES_CHECK(s, h->recvBuf.reserve(3ull * nR + 1));
  PIES_CHECK(s, cudaStreamSynchronize(s->stream));
  h->listsSet = true;
  return PIES_B200_OK;
}

int haloExchange(PiesB200Solver* s, int planes) {
  HaloWork* h = s->halo;
  if (!h || !h->listsSet || h->peers.empty()) return PIES_B200_OK;
  if (planes != 1 && planes != 3) return fail(s, PIES_B200_EINVAL, "halo_exchange: planes must be 1 or 3");
  NcclApi& a = nccl();
  const int nPeers = (int)h->peers.size();
  const uint32_t nS = h->sendOff[nPeers], nR = h->recvOff[nPeers];
  cudaStream_t st = s->stream;
  ncclComm_t comm = static_cast<ncclComm_t>(h->comm);
  if (nS) {
    const uint32_t* t = h->sendIdx.p;
    k_halo_pack<<<(nS + kThreads - 1) / kThreads, kThreads, 0, st>>>(nS, t, t + nS, t + 2ull * nS, t + 2ull * nS + nPeers, planes,
                                                                    s->q.p, s->prev.p, s->vel.p, h->sendBuf.p);
    ++s->launches;
  }
  NCHECK(s, a.GroupStart());
  for (int k = 0; k < nPeers; ++k) {
    const size_t cs = (size_t)(h->sendOff[k + 1] - h->sendOff[k]) * planes * 4, cr = (size_t)(h->recvOff[k + 1] - h->recvOff[k]) * planes * 4;
    if (cs) NCHECK(s, a.Send(h->sendBuf.p + (size_t)planes * h->sendOff[k], cs, ncclFloat, h->peers[k], comm, st));
    if (cr) NCHECK(s, a.Recv(h->recvBuf.p + (size_t)planes * h->recvOff[k], cr, ncclFloat, h->peers[k], comm, st));
  }
  NCHECK(s, a.GroupEnd());
  if (nR) {
    const uint32_t* t = h->recvIdx.p;
    k_halo_unpack<<<(nR + kThreads - 1) / kThreads, kThreads, 0, st>>>(nR, t, t + nR, t + 2ull * nR, t + 2ull * nR + nPeers, planes,
                                                                      s->q.p, s->prev.p, s->vel.p, h->recvBuf.p);
    ++s->launches;
  }
  h->bytesLastTick += 16ull * planes * ((uint64_t)nS + nR);
  ++h->exchangesLastTick;
  return PIES_B200_OK;
}

int haloMixedIslands(PiesB200Solver* s, bool localOk, bool* anyMixed) {
  HaloWork* h = s->halo;
  *anyMixed = true;
  if (!h || h->world == 1) { *anyMixed = false; return PIES_B200_OK; }
  static const bool always = std::getenv("PIES_B200_HALO_EVERY_ITERATION") != nullptr;   // A/B switch: never skip
  NcclApi& a = nccl();
  cudaStream_t st = s->stream;
  IslandWork* w = s->islands;
  const bool canTell = localOk && !always && h->listsSet && w && w->nBodies && h->ghost.cap >= s->n;
  h->hostStatus[2] = canTell ? 0 : 1;
  PIES_CHECK(s, cudaMemcpyAsync(h->status.p + 2, h->hostStatus + 2, sizeof(int), cudaMemcpyHostToDevice, st));
  if (canTell && w->united) {   // without contacts every body is its own island, and a body is owned or ghost as a whole
    PIES_CHECK(s, h->bodyMix.reserve(w->nBodies));
    PIES_CHECK(s, cudaMemsetAsync(h->bodyMix.p, 0, (size_t)w->nBodies * sizeof(uint32_t), st));
    k_halo_mix_mark<<<(s->n + kThreads - 1) / kThreads, kThreads, 0, st>>>(s->n, w->bodyOf.p, w->parent.p, h->ghost.p, h->bodyMix.p);
    k_halo_mix_any<<<(w->nBodies + kThreads - 1) / kThreads, kThreads, 0, st>>>(w->nBodies, h->bodyMix.p, h->status.p + 2);
    s->launches += 2;
  }
  NCHECK(s, a.AllReduce(h->status.p + 2, h->status.p + 3, 1, ncclInt, ncclMax, static_cast<ncclComm_t>(h->comm), st));
  PIES_CHECK(s, cudaMemcpyAsync(h->hostStatus + 3, h->status.p + 3, sizeof(int), cudaMemcpyDeviceToHost, st));
  PIES_CHECK(s, cudaStreamSynchronize(st));
  *anyMixed = h->hostStatus[3] != 0;
  return PIES_B200_OK;
}

int haloAnyFailed(PiesB200Solver* s, bool mine, bool* any) {
  HaloWork* h = s->halo;
  *any = mine;
  if (!h || h->world == 1) return PIES_B200_OK;
  NcclApi& a = nccl();
  h->hostStatus[0] = mine ? 1 : 0;
  PIES_CHECK(s, cudaMemcpyAsync(h->status.p, h->hostStatus, sizeof(int), cudaMemcpyHostToDevice, s->stream));
  NCHECK(s, a.AllReduce(h->status.p, h->status.p + 1, 1, ncclInt, ncclMax, static_cast<ncclComm_t>(h->comm), s->stream));
  PIES_CHECK(s, cudaMemcpyAsync(h->hostStatus + 1, h->status.p + 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  PIES_CHECK(s, cudaStreamSynchronize(s->stream));
  *any = h->hostStatus[1] != 0;
  return PIES_B200_OK;
}

}  // namespace pies
