// probes.cu — [additive] per-kernel probes: they run the very device functions the hot
// kernels use (svd3.cuh, ccd.cuh, sort.cu) on caller-provided data, so the unit parity tests
// exercise the product code, not a copy of it.
#include <vector>

#include "../../include/pies_b200.h"
#include "ccd.cuh"
#include "kernels.h"
#include "svd3.cuh"

namespace pies {

__global__ void k_probe_tet(uint32_t n, const float* __restrict__ pos, const float* __restrict__ qinv, float lo, float hi,
                            int volume, float* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pos + 12ull * i;
  float qi[9];
  for (int k = 0; k < 9; ++k) qi[k] = qinv[9ull * i + k];
  M3 F = deformationGradient(v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), v3(p[6], p[7], p[8]), v3(p[9], p[10], p[11]), qi);
  M3 U, V;
  float sg[3], d[3];
  svd3(F, U, sg, V);
  if (volume) volumeSigma(sg, lo, hi, d); else strainSigma(sg, det3(F), lo, hi, d);
  M3 R = recompose(U, d, V);
  float* o = out + 12ull * i;
  o[0] = o[1] = o[2] = 0.0f;                      // projected[0] = 0 (Constraints.cpp:124)
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) o[3 + 3 * c + r] = R.m[r][c];  // projected[c+1] = column c
}

__global__ void k_probe_ccd(uint32_t n, const float* __restrict__ in, float threshold, int32_t* __restrict__ hit,
                            float* __restrict__ t) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = in + 18ull * i;
  float tt = -1.0f;
  bool h = ex::pointTriangleCCD(v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), v3(p[6], p[7], p[8]), v3(p[9], p[10], p[11]),
                                v3(p[12], p[13], p[14]), v3(p[15], p[16], p[17]), threshold, tt);
  hit[i] = h ? 1 : 0;
  t[i] = h ? tt : -1.0f;
}

__global__ void k_probe_edge_ccd(uint32_t n, const float* __restrict__ in, int32_t* __restrict__ hit, float* __restrict__ t) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = in + 18ull * i;
  float tt = -1.0f;
  bool h = ex::edgeEdgeCCD(v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), v3(p[6], p[7], p[8]), v3(p[9], p[10], p[11]),
                           v3(p[12], p[13], p[14]), v3(p[15], p[16], p[17]), tt);
  hit[i] = h ? 1 : 0;
  t[i] = h ? tt : -1.0f;
}

__global__ void k_probe_tri_range(uint32_t n, const float* __restrict__ pos, const float* __restrict__ prev,
                                  long long* __restrict__ mins, uint32_t* __restrict__ lens) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pos + 9ull * i;
  const float* o = prev + 9ull * i;
  int mx, my, mz; unsigned lx, ly, lz; bool bad;
  ex::triCellRange(v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), v3(p[6], p[7], p[8]), v3(o[0], o[1], o[2]), v3(o[3], o[4], o[5]),
                   v3(o[6], o[7], o[8]), mx, my, mz, lx, ly, lz, bad);
  if (lx > 50u || ly > 50u || lz > 50u) { lx = ly = lz = 0; mx = my = mz = 0; }
  mins[3 * i] = mx; mins[3 * i + 1] = my; mins[3 * i + 2] = mz;
  lens[3 * i] = lx; lens[3 * i + 1] = ly; lens[3 * i + 2] = lz;
}

__global__ void k_probe_node_range(uint32_t n, const float* __restrict__ pos, const float* __restrict__ radius, float scale,
                                   long long* __restrict__ mins, uint32_t* __restrict__ lens) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int mx, my, mz; unsigned lx, ly, lz; bool bad;
  ex::nodeCellRange(v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), radius[i], scale, mx, my, mz, lx, ly, lz, bad);
  mins[3 * i] = mx; mins[3 * i + 1] = my; mins[3 * i + 2] = mz;
  lens[3 * i] = lx; lens[3 * i + 1] = ly; lens[3 * i + 2] = lz;
}

struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() { for (void* p : ptrs) cudaFree(p); }
  template <typename T>
  T* dev(size_t n, const T* host = nullptr) {
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    if (host && n) cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice);
    return static_cast<T*>(p);
  }
};

static int finish() { return cudaDeviceSynchronize() == cudaSuccess && cudaGetLastError() == cudaSuccess ? PIES_B200_OK : PIES_B200_ECUDA; }
static int haveDevice() { int c = 0; return cudaGetDeviceCount(&c) == cudaSuccess && c > 0; }

static int probeTet(uint32_t n, const float* pos, const float* qinv, float lo, float hi, int volume, float* out) {
  if (!haveDevice()) return PIES_B200_ENODEV;
  if (!n) return PIES_B200_OK;
  Scratch sc;
  float* dp = sc.dev<float>(12ull * n, pos); float* dq = sc.dev<float>(9ull * n, qinv); float* dout = sc.dev<float>(12ull * n);
  if (!dp || !dq || !dout) return PIES_B200_ECUDA;
  k_probe_tet<<<(n + 127) / 128, 128>>>(n, dp, dq, lo, hi, volume, dout);
  if (finish()) return PIES_B200_ECUDA;
  cudaMemcpy(out, dout, 48ull * n, cudaMemcpyDeviceToHost);
  return PIES_B200_OK;
}

}  // namespace pies

using namespace pies;

extern "C" {

int pies_b200_probe_tet_projection(uint32_t n, const float* pos, const float* qinv, float lo, float hi, float* out) {
  return probeTet(n, pos, qinv, lo, hi, 0, out);
}
int pies_b200_probe_volume_projection(uint32_t n, const float* pos, const float* qinv, float lo, float hi, float* out) {
  return probeTet(n, pos, qinv, lo, hi, 1, out);
}
int pies_b200_probe_ccd(uint32_t n, const float* in, float threshold, int32_t* hit, float* t) {
  if (!haveDevice()) return PIES_B200_ENODEV;
  if (!n) return PIES_B200_OK;
  Scratch sc;
  float* din = sc.dev<float>(18ull * n, in); int32_t* dh = sc.dev<int32_t>(n); float* dt = sc.dev<float>(n);
  if (!din || !dh || !dt) return PIES_B200_ECUDA;
  k_probe_ccd<<<(n + 127) / 128, 128>>>(n, din, threshold, dh, dt);
  if (finish()) return PIES_B200_ECUDA;
  cudaMemcpy(hit, dh, 4ull * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(t, dt, 4ull * n, cudaMemcpyDeviceToHost);
  return PIES_B200_OK;
}
int pies_b200_probe_edge_ccd(uint32_t n, const float* in, int32_t* hit, float* t) {
  if (!haveDevice()) return PIES_B200_ENODEV;
  if (!n) return PIES_B200_OK;
  Scratch sc;
  float* din = sc.dev<float>(18ull * n, in); int32_t* dh = sc.dev<int32_t>(n); float* dt = sc.dev<float>(n);
  if (!din || !dh || !dt) return PIES_B200_ECUDA;
  k_probe_edge_ccd<<<(n + 127) / 128, 128>>>(n, din, dh, dt);
  if (finish()) return PIES_B200_ECUDA;
  cudaMemcpy(hit, dh, 4ull * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(t, dt, 4ull * n, cudaMemcpyDeviceToHost);
  return PIES_B200_OK;
}
int pies_b200_probe_tri_range(uint32_t n, const float* pos, const float* prev, int64_t* mins, uint32_t* lens) {
  if (!haveDevice()) return PIES_B200_ENODEV;
  if (!n) return PIES_B200_OK;
  Scratch sc;
  float* dp = sc.dev<float>(9ull * n, pos); float* dq = sc.dev<float>(9ull * n, prev);
  long long* dm = sc.dev<long long>(3ull * n); uint32_t* dl = sc.dev<uint32_t>(3ull * n);
  if (!dp || !dq || !dm || !dl) return PIES_B200_ECUDA;
  k_probe_tri_range<<<(n + 127) / 128, 128>>>(n, dp, dq, dm, dl);
  if (finish()) return PIES_B200_ECUDA;
  cudaMemcpy(mins, dm, 24ull * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(lens, dl, 12ull * n, cudaMemcpyDeviceToHost);
  return PIES_B200_OK;
}
int pies_b200_probe_node_range(uint32_t n, const float* pos, const float* radius, float gridScale, int64_t* mins, uint32_t* lens) {
  if (!haveDevice()) return PIES_B200_ENODEV;
  if (!n) return PIES_B200_OK;
  Scratch sc;
  float* dp = sc.dev<float>(3ull * n, pos); float* dr = sc.dev<float>(n, radius);
  long long* dm = sc.dev<long long>(3ull * n); uint32_t* dl = sc.dev<uint32_t>(3ull * n);
  if (!dp || !dr || !dm || !dl) return PIES_B200_ECUDA;
  k_probe_node_range<<<(n + 127) / 128, 128>>>(n, dp, dr, gridScale, dm, dl);
  if (finish()) return PIES_B200_ECUDA;
  cudaMemcpy(mins, dm, 24ull * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(lens, dl, 12ull * n, cudaMemcpyDeviceToHost);
  return PIES_B200_OK;
}
int pies_b200_probe_sort_pairs(uint64_t n, uint64_t* keys, uint32_t* vals, int keyBits) {
  if (!haveDevice()) return PIES_B200_ENODEV;
  if (!n) return PIES_B200_OK;
  Scratch sc;
  uint64_t* dk = sc.dev<uint64_t>(n, keys); uint64_t* tk = sc.dev<uint64_t>(n);
  uint32_t* dv = sc.dev<uint32_t>(n, vals); uint32_t* tv = sc.dev<uint32_t>(n);
  uint32_t* hist = sc.dev<uint32_t>(sortHistBytes(n) / 4 + 4);
  if (!dk || !tk || !dv || !tv || !hist) return PIES_B200_ECUDA;
  launchSortPairs(nullptr, n, dk, dv, tk, tv, hist, keyBits);
  if (finish()) return PIES_B200_ECUDA;
  cudaMemcpy(keys, dk, 8ull * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(vals, dv, 4ull * n, cudaMemcpyDeviceToHost);
  return PIES_B200_OK;
}

}  // extern "C"
