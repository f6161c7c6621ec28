// inc.conv (Conv2d(1, 32, 3) valid + bias + ReLU, unet_parts.py:57-87 via inconv :196-203) on the tensor cores, bf16 path.
//
// The layer has ONE input channel: K = 9 taps.  As CUDA-core FMAs it is 288 FMA per pixel and FMA-bound at 62 % of the
// write bandwidth (profiles/r1_hbm_kernels.txt).  Here every CTA builds the im2col rows of 128 consecutive output pixels
// in shared memory itself - one thread per pixel writes its 9 taps as a K-major operand row - and ONE pair of
// tcgen05.mma (M = 128, N = 32, K = 2 x 16) produces all 32 channels.  To keep fp32-grade accuracy on the fp32 input
// image the row holds a three-term bf16 split: K slots [x_hi(9) | x_hi(9) | x_lo(9) | 0(5)] against weight rows
// [w_hi | w_lo | w_hi | 0] give x_hi.w_hi + x_hi.w_lo + x_lo.w_hi in the fp32 accumulator (~2^-16 relative per product).
// Persistent CTAs of 128 threads (64 TMEM columns, 18 KB of shared memory each, 8 resident per SM) walk 128-pixel blocks,
// double-buffered so that the MMAs of block i overlap the stores of block i-1.
// MEASURED (tools/convt_bench.py, 60 tiles): 81 us against 74 us for the CUDA-core kernel - the tensor core removes the
// 288 FMAs per pixel, but building the operand row (split, pack, 4 shared stores) and the epilogue cost ~300 instructions
// per pixel, i.e. the same ~9 warp instructions per pixel as the FMAs did.  Kept as an opt-in (UNCL_CONV_FIRST_TC) and as
// the worked example of an im2col-in-shared-memory tcgen05 operand; a one-allocation-per-block first version took 167 us
// (tcgen05.alloc / dealloc serialise per SM).
#include <cstdlib>
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

__device__ __forceinline__ void load_taps(const float* __restrict__ x, int H, int W, int Wo, int npix, int blocks_per_img,
                                          int blk, int t, float (&v)[9]) {
  const int n = blk / blocks_per_img;
  const int q = (blk - n * blocks_per_img) * 128 + t;
  const bool valid = q < npix;
  const int oy = valid ? q / Wo : 0, ox = valid ? q - oy * Wo : 0;
  const float* xi = x + ((long)n * H + oy) * W + ox;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) v[ky * 3 + kx] = valid ? __ldg(xi + ky * W + kx) : 0.f;
}

__device__ __forceinline__ void first_epilogue(uint32_t tacc, int warp, const float* s_bias, bf16* __restrict__ out,
                                               long out_img_stride, long cb_stride, int npix, int blocks_per_img, int blk,
                                               int t, int act) {
  uint32_t r[32];
  tc_ld32(tacc + ((uint32_t)(warp * 32) << 16), r);
  const int n = blk / blocks_per_img;
  const int q = (blk - n * blocks_per_img) * 128 + t;
  if (q < npix) {
    bf16* o = out + (long)n * out_img_stride + (long)q * 8;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float vv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) vv[j] = apply_act(__uint_as_float(r[g * 8 + j]) + s_bias[g * 8 + j], act);
      store8(o + g * cb_stride, vv);
    }
  }
}

// Persistent: a CTA allocates its 64 TMEM columns once (tcgen05.alloc / dealloc serialise per SM: one allocation per
// 128-pixel block cost 167 us for the 1080p frame) and walks blocks with two A tiles / two accumulators: the MMAs of
// block i run while the CTA stores block i-1, and the taps of block i+1 are already in flight.
__global__ void __launch_bounds__(128, 8)
conv_first_tc_kernel(const float* __restrict__ x, const bf16* __restrict__ wb, const float* __restrict__ bias,
                     bf16* __restrict__ out, long out_img_stride, int H, int W, int act, int blocks_per_img,
                     int total_blocks) {
  __shared__ __align__(128) uint8_t s_a[2][4 * 128 * 16];   // [buffer][k group][pixel][8 bf16]
  __shared__ __align__(128) uint8_t s_b[4 * 32 * 16];       // [k group][channel][8 bf16]
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  __shared__ float s_bias[32];
  const int t = threadIdx.x, warp = t >> 5;
  const int Ho = H - 2, Wo = W - 2, npix = Ho * Wo;

  if (t == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  reinterpret_cast<uint4*>(s_b)[t] = __ldg(reinterpret_cast<const uint4*>(wb) + t);   // weights (2 KB, L2-resident)
  if (t < 32) s_bias[t] = bias[t];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t dhi = (128u >> 4) | (1u << 14);                            // SBO 128 B, descriptor version 1
  const uint32_t lbo_a = (128u * 16u >> 4) << 16;                           // LBO_A = 128 rows x 16 B
  const uint32_t b0 = (smem_u32(s_b) >> 4) | ((32u * 16u >> 4) << 16);      // LBO_B = 32 rows x 16 B
  const long cb_stride = (long)npix * 8;
  uint32_t phase[2] = {0, 0};
  float v[9];
  int blk = blockIdx.x, prev = -1, buf = 0;
  if (blk < total_blocks) load_taps(x, H, W, Wo, npix, blocks_per_img, blk, t, v);
  for (; blk < total_blocks; blk += gridDim.x, buf ^= 1) {
    // this pixel's im2col row: 9 taps, split into bf16 hi / lo -> K slots [hi | hi | lo | 0]
    {
      unsigned short k16[32];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const bf16 hi = __float2bfloat16_rn(v[j]);
        const bf16 lo = __float2bfloat16_rn(v[j] - __bfloat162float(hi));
        k16[j] = __bfloat16_as_ushort(hi);
        k16[9 + j] = __bfloat16_as_ushort(hi);
        k16[18 + j] = __bfloat16_as_ushort(lo);
      }
#pragma unroll
      for (int j = 27; j < 32; ++j) k16[j] = 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = (uint32_t)k16[g * 8 + 0] | ((uint32_t)k16[g * 8 + 1] << 16);
        u.y = (uint32_t)k16[g * 8 + 2] | ((uint32_t)k16[g * 8 + 3] << 16);
        u.z = (uint32_t)k16[g * 8 + 4] | ((uint32_t)k16[g * 8 + 5] << 16);
        u.w = (uint32_t)k16[g * 8 + 6] | ((uint32_t)k16[g * 8 + 7] << 16);
        *reinterpret_cast<uint4*>(s_a[buf] + (g * 128 + t) * 16) = u;
      }
    }
    const int next = blk + gridDim.x;
    if (next < total_blocks) load_taps(x, H, W, Wo, npix, blocks_per_img, next, t, v);   // in flight during the MMA + epilogue
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();   // A[buf] complete; every warp has finished reading accumulator `buf` (block i-2)
    tc_fence_after();
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t a0 = (smem_u32(s_a[buf]) >> 4) | lbo_a, d = tmem + (uint32_t)(buf * 32);
        tc_mma_bf16(d, a0, dhi, b0, dhi, idesc, 0u);
        tc_mma_bf16(d, a0 + (2u * 128u * 16u >> 4), dhi, b0 + (2u * 32u * 16u >> 4), dhi, idesc, 1u);
        tc_commit(&s_bar[buf]);
      }
      __syncwarp();
    }
    if (prev >= 0) {   // store block i-1 while the MMAs of block i run
      mbar_wait(&s_bar[buf ^ 1], phase[buf ^ 1]);
      phase[buf ^ 1] ^= 1;
      tc_fence_after();
      first_epilogue(tmem + (uint32_t)((buf ^ 1) * 32), warp, s_bias, out, out_img_stride, cb_stride, npix, blocks_per_img, prev, t, act);
    }
    prev = blk;
  }
  if (prev >= 0) {
    buf ^= 1;   // the buffer of the last issued block
    mbar_wait(&s_bar[buf], phase[buf]);
    tc_fence_after();
    first_epilogue(tmem + (uint32_t)(buf * 32), warp, s_bias, out, out_img_stride, cb_stride, npix, blocks_per_img, prev, t, act);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
  }
}

}  // namespace

// x fp32 [N][H][W]; w_split bf16 [4][32][8] (uncltmo_b200/packing.py:conv_first_tc_split); out bf16 blocked [N][4][H-2][W-2][8]
extern "C" int uncl_conv_first_tc(const float* x, const void* w_split, const float* bias, void* out, long out_img_stride,
                                  int N, int H, int W, int C_out, int act, cudaStream_t stream) {
  UNCL_REQUIRE(C_out == 32 && H > 2 && W > 2 && N > 0, "conv_first_tc: bad shape N=%d H=%d W=%d C_out=%d", N, H, W, C_out);
  UNCL_REQUIRE(act == UNCL_ACT_RELU || act == UNCL_ACT_NONE, "conv_first_tc: only ReLU / identity");
  UNCL_REQUIRE((reinterpret_cast<uintptr_t>(w_split) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && out_img_stride % 8 == 0,
               "conv_first_tc: operands must be 16-byte aligned");
  const int blocks_per_img = (int)(((long)(H - 2) * (W - 2) + 127) / 128);
  const long total = (long)blocks_per_img * N;
  UNCL_REQUIRE(total < (1L << 31), "conv_first_tc: too many blocks");
  int dev = 0, sms = 148, per_sm = 8;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  per_sm = 8;   // __launch_bounds__(128, 8): 64 registers x 128 threads x 8 = the register file; 8 x 64 TMEM columns
  if (const char* e = getenv("UNCL_CONV_FIRST_CTAS")) { const int v = atoi(e); if (v >= 1 && v <= 8) per_sm = v; }
  const long grid = total < (long)sms * per_sm ? total : (long)sms * per_sm;
  conv_first_tc_kernel<<<(unsigned)grid, 128, 0, stream>>>(x, reinterpret_cast<const bf16*>(w_split), bias,
                                                         reinterpret_cast<bf16*>(out), out_img_stride, H, W, act,
                                                         blocks_per_img, (int)total);
  return uncl_check_launch("conv_first_tc");
}
