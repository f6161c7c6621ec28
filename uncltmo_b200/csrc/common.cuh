// Shared helpers for the uncltmo_b200 kernels (sm_100a only).
//
// Activation layout used by every kernel ("C8-blocked"): [N][C/8][H][W][8], i.e. channels are split in
// blocks of 8 that sit innermost.  One 8-channel group of a pixel is 16 B in bf16 / 32 B in fp32, so
//   * consecutive pixels of a channel block are contiguous -> coalesced epilogue stores and TMA boxes whose
//     shared-memory image is directly the no-swizzle K-major UMMA core-matrix layout (8 rows x 16 B), and
//   * a tensor may be a view into a wider "concat" buffer: only the per-image stride differs.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "uncltmo_b200.h"

enum { UNCL_F32 = 0, UNCL_BF16 = 1 };
enum { UNCL_ACT_NONE = 0, UNCL_ACT_RELU = 1, UNCL_ACT_LRELU = 2, UNCL_ACT_GELU = 3, UNCL_ACT_SIGMOID = 4 };

int uncl_set_error(int code, const char* fmt, ...);
int uncl_check_launch(const char* what);

#define UNCL_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return uncl_set_error(UNCL_EINVAL, __VA_ARGS__); \
  } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void load8(const float* p, float v[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float v[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8(float* p, const float v[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void store8(bf16* p, const float v[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ void from_f(float& d, float x) { d = x; }
__device__ __forceinline__ void from_f(bf16& d, float x) { d = __float2bfloat16_rn(x); }

// MUFU.SQRT: max relative error 2^-23 (PTX ISA, sqrt.approx.f32) - one instruction instead of the ~10 of the IEEE sqrtf;
// used where the result is an activation that is rounded to bf16 or feeds further fp32 arithmetic with looser tolerances
// (x * rsqrt(x) measured the same: the skip operators are bound by shared-memory bandwidth, not by the MUFU pipe)
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case UNCL_ACT_RELU: return fmaxf(x, 0.f);
    case UNCL_ACT_LRELU: return x > 0.f ? x : 0.2f * x;
    case UNCL_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    case UNCL_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum (all threads get the result).  `red` must hold >= 33 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

#define UNCL_DISPATCH_DTYPE(dtype, T, ...)                                      \
  do {                                                                          \
    if ((dtype) == UNCL_F32) { typedef float T; __VA_ARGS__; }                  \
    else if ((dtype) == UNCL_BF16) { typedef bf16 T; __VA_ARGS__; }             \
    else return uncl_set_error(UNCL_EINVAL, "unknown dtype %d", (int)(dtype));  \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
