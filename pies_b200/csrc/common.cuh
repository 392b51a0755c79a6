// common.cuh — shared device helpers for the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pies {

constexpr int kNumSMs = 148;          // B200
constexpr int kThreads = 256;         // default CTA size for streaming kernels
constexpr int kReduceBlocks = 592;    // 148 SMs x 4 resident CTAs: fixed grid => fixed-order reductions
constexpr int kMaxReduceBlocks = 2368; // 148 x 16: largest grid a kernel with a last-CTA reduction may use (partials capacity)

struct V3 { float x, y, z; };

__host__ __device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 v3(float4 a) { return V3{a.x, a.y, a.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
__device__ __forceinline__ V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 x, V3 y) {
  return V3{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
__device__ __forceinline__ float4 f4(V3 a, float w) { return make_float4(a.x, a.y, a.z, w); }

// Streaming (touch-once) accesses: keep them out of L1 so gathered node data stays resident.
__device__ __forceinline__ float4 ldStream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ uint4 ldStream(const uint4* p) { return __ldcs(p); }
__device__ __forceinline__ void stStream(float4* p, float4 v) { __stcs(p, v); }

// Acquire / release accesses of the per-node progress counters of the ordered (dataflow) sweeps.
__device__ __forceinline__ uint32_t ldAcquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stRelease(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float warpSum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Fixed-order CTA reduction of up to K values per thread; result valid in thread 0.
template <int K>
__device__ __forceinline__ void blockSum(float (&v)[K], float* smem /* K * 32 floats */) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warpSum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) smem[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float t = lane < nwarps ? smem[k * 32 + lane] : 0.0f;
      v[k] = warpSum(t);
    }
  }
  __syncthreads();
}

}  // namespace pies
