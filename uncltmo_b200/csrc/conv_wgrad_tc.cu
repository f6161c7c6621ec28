// 3x3 weight gradient on the tensor cores:  dW9[t][ci][co] += sum_pixels X[pix + off_t, ci] * dZ[pix, co].
//
// GEMM with K = pixels.  Both operands are C8-blocked bf16 tensors whose TMA image in shared memory, [block][pixel][8ch],
// is read as an MN-MAJOR UMMA operand: a core matrix is 8 channels (M or N, contiguous 16 B) x 8 pixels (K, 16 B apart),
// LBO = 128 B between 8-pixel groups, SBO = the channel-block stride.  One MMA consumes 16 consecutive pixels of a tile
// row; a filter tap (ky,kx) is again only a shifted A start address, so one X tile serves 9 (or 3) taps.
// D (TMEM): lane = input channel (M = 128), column = tap * Nc + output channel; a CTA accumulates over all its pixel
// tiles in TMEM and adds its partial sum to dW once, with fp32 atomics.
//   co <= 48 : all 9 taps per CTA (9*Nc <= 512 TMEM columns), X tile with a 2-row halo
//   co >= 64 : one filter row (3 taps) per CTA, Nc = min(co, 128) output channels
// Zero padding (ConvTranspose) and ragged tile edges are TMA out-of-bounds zero fill on X and dZ.
//
// Reference: autograd of nn.Conv2d / nn.ConvTranspose2d in models/unet_multi_filters/unet_parts.py.
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int kWgThreads = 320;   // warp 0 producer, warp 1 MMA issuer, warps 2-9 epilogue
constexpr int kWgMaxStages = 4;

struct WgParams {
  float* dW;
  int N, C_in, C_out, Ho, Wo, pad;
  int taps9, nky, nkx, Nc, n_co_chunks, n_ci_chunks, Mc_blocks;   // nkx = 1: pointwise (1 tap) mode
  int RB, BW, RBx, PWx;
  int bands, row_tiles, tiles_per_img, total_tiles;
  int groups, ctas_per_group;
  int stages, x_bytes, z_bytes, x_stage_bytes, stage_bytes;
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_z,
                        const __grid_constant__ WgParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  // the MMA reads 16 channel blocks of X even when fewer exist (garbage rows of D, never stored): keep that overrun
  // inside the allocation by placing the barriers in front of the stages
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kWgMaxStages;
  uint64_t* done = empty + kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  uint8_t* stage_base = smem + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = p.stages, stage_bytes = p.stage_bytes;
  const int group = blockIdx.x / p.ctas_per_group, slot = blockIdx.x % p.ctas_per_group;
  const int kyi = group % p.nky, cic = (group / p.nky) % p.n_ci_chunks, coc = group / (p.nky * p.n_ci_chunks);
  const int ntaps = p.taps9 ? 9 : p.nkx;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_z) : "memory");
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(p.x_bytes + p.z_bytes);
      for (int t = slot; t < p.total_tiles; t += p.ctas_per_group) {
        const int n = t / p.tiles_per_img, r = t - n * p.tiles_per_img;
        const int band = r % p.bands, rt = r / p.bands;
        const int x0 = band * p.BW, y0 = rt * p.RB;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sx = stage_base + (size_t)stage * stage_bytes;
        mbar_expect_tx(&full[stage], tx);
        tma_load_4d(sx, &tmap_x, &full[stage], (x0 - p.pad) * 2, y0 - p.pad + (p.taps9 ? 0 : kyi), cic * 16, n);
        tma_load_4d(sx + p.x_stage_bytes, &tmap_z, &full[stage], x0 * 2, y0, coc * (p.Nc / 8), n);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // M = 128, N = Nc, bf16 x bf16 -> fp32, A and B both MN-major (bits 15 / 16)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Nc >> 3) << 17) |
                           ((128u >> 4) << 24);
    const uint32_t lo_const = (128u >> 4) << 16;                                        // LBO = 128 B (next 8 pixels)
    const uint32_t a_hi = ((uint32_t)(p.RBx * p.PWx) & 0x3fffu) | (1u << 14);           // SBO_A = RBx*PWx*16 B
    const uint32_t b_hi = ((uint32_t)(p.RB * p.BW) & 0x3fffu) | (1u << 14);             // SBO_B = RB*BW*16 B
    const uint32_t stage0_16 = smem_u32(stage_base) >> 4, stage_16 = (uint32_t)stage_bytes >> 4;
    const uint32_t xs_16 = (uint32_t)p.x_stage_bytes >> 4;
    const uint32_t pwx = (uint32_t)p.PWx, bw = (uint32_t)p.BW, nc = (uint32_t)p.Nc;
    const int ksteps = p.BW / 16, nky_local = p.taps9 ? 3 : 1;
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int t = slot; t < p.total_tiles; t += p.ctas_per_group) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t sx16 = stage0_16 + (uint32_t)stage * stage_16;
      const uint32_t sz16 = sx16 + xs_16;
      if (elect_one()) {
        for (int r = 0; r < p.RB; ++r) {
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t b_lo = lo_const | (sz16 + (uint32_t)r * bw + 16u * (uint32_t)k);
            const uint32_t accum = (first && r == 0 && k == 0) ? 0u : 1u;
            uint32_t d = tmem_base;
            for (int ky = 0; ky < nky_local; ++ky) {
              const uint32_t a_row = lo_const | (sx16 + (uint32_t)(r + ky) * pwx + 16u * (uint32_t)k);
              if (p.nkx == 3) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  tc_mma_bf16(d, a_row + (uint32_t)kx, a_hi, b_lo, b_hi, idesc, accum);
                  d += nc;
                }
              } else {
                tc_mma_bf16(d, a_row, a_hi, b_lo, b_hi, idesc, accum);
              }
            }
          }
        }
      }
      __syncwarp();
      first = false;
      if (elect_one()) tc_commit(&empty[stage]);
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) tc_commit(done);
    __syncwarp();
  } else {
    // epilogue: TMEM -> fp32 atomics on dW9[tap][ci][co]
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int ci = cic * 128 + quarter * 32 + lane;
    const bool valid = ci < p.C_in && slot < p.total_tiles;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    mbar_wait(done, 0);
    tc_fence_after();
    const int cols = ntaps * p.Nc;
    for (int c0 = half * 32; c0 < cols; c0 += 64) {
      uint32_t r[32];
      tc_ld32(tmem_base + lane_base + (uint32_t)c0, r);
      if (valid) {
        const int tl = c0 / p.Nc, co = coc * p.Nc + (c0 - tl * p.Nc);
        const int tap = p.taps9 ? tl : kyi * 3 + tl;
        float* dst = p.dW + ((long)tap * p.C_in + ci) * p.C_out + co;
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(r[j]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// C_out = 32 (52 % of the generator's FLOPs: inc.conv1, up3.conv / conv1, up2.conv / conv1): the TRANSPOSED weight gradient.
//
// With only 32 output channels the kernel above issues nine N = 32 instructions per K step - each reads its 4 KB A tile
// for 40 cycles to do 16 cycles of math (tools/mma_probe.cu) - and for C_in = 32 three quarters of the M = 128 rows are
// padding.  Here the roles are swapped and the filter rows are stacked in M for free:
//     D[(j, co)][kx * Nci + ci] += sum_pixels dZ[y + j][x - kx][co] * X[Y][c][ci]
//   * A = dZ.  Its TMA map is declared with the dimensions ordered (x, channel block, y, image), so the box lands in shared
//     memory as [row][channel block][pixel][8]: the M-block index (j, channel block) = 4 j + cb then has ONE stride (a row
//     of PWz pixels), i.e. three consecutive dZ rows ARE a 96-row MN-major operand - no copies.  Row j pairs with filter
//     row ky = 2 - j.
//   * B = X, N = min(C_in, 128) input channels: at N = 128 the instruction is math-bound (64 cycles), M is 75 % used.
//   * the filter column kx is a start-address shift of A by kx pixels (three instructions per K step, three accumulators
//     side by side in TMEM: 3 N <= 384 columns).
// Work is partitioned over X pixels (row tiles x column bands x images), each pixel meets all nine taps exactly once;
// everything outside X or dZ is TMA zero fill (which is also the ConvTranspose padding).  Per 16 pixels: 3 x 64 cycles at
// N = 128 (was 9 x 40), 3 x 40 at N = 32 (was 9 x 40 with a quarter of the rows used).
// The epilogue's atomics are coalesced: lanes are output channels, contiguous in dW9[tap][ci][co].
struct WtParams {
  float* dW;
  int C_in, C_out, pad;
  int Nci, n_ci_chunks;          // input channels per CTA group (MMA N) and number of groups
  int RB, BW, PWz, RBz;          // X tile rows / columns; dZ tile columns (BW + 2) / rows (RB + 2)
  int bands, row_tiles, tiles_per_img, total_tiles, ctas_per_group;
  int stages, a_bytes, b_bytes, a_stage_bytes, stage_bytes;
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv3x3_wgrad_t32_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_z,
                         const __grid_constant__ WtParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kWgMaxStages;
  uint64_t* done = empty + kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  uint8_t* stage_base = smem + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = p.stages, stage_bytes = p.stage_bytes;
  const int cic = blockIdx.x / p.ctas_per_group, slot = blockIdx.x % p.ctas_per_group;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_z) : "memory");
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(p.a_bytes + p.b_bytes);
      for (int t = slot; t < p.total_tiles; t += p.ctas_per_group) {
        const int n = t / p.tiles_per_img, r = t - n * p.tiles_per_img;
        const int band = r % p.bands, rt = r / p.bands;
        const int c0 = band * p.BW, Y0 = rt * p.RB;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = stage_base + (size_t)stage * stage_bytes;
        mbar_expect_tx(&full[stage], tx);
        // dZ: columns c0 + pad - 2 .., rows Y0 + pad - 2 .., all four channel blocks (map dims: x, cb, y, n)
        tma_load_4d(sa, &tmap_z, &full[stage], (c0 + p.pad - 2) * 2, 0, Y0 + p.pad - 2, n);
        tma_load_4d(sa + p.a_stage_bytes, &tmap_x, &full[stage], c0 * 2, Y0, cic * (p.Nci / 8), n);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // M = 128 (96 used), N = Nci, bf16 x bf16 -> fp32, A and B both MN-major (bits 15 / 16)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Nci >> 3) << 17) |
                           ((128u >> 4) << 24);
    const uint32_t lo_const = (128u >> 4) << 16;                                  // LBO = 128 B (next 8 pixels)
    const uint32_t a_hi = ((uint32_t)p.PWz & 0x3fffu) | (1u << 14);               // SBO_A = PWz*16 B: next (row, channel block)
    const uint32_t b_hi = ((uint32_t)(p.RB * p.BW) & 0x3fffu) | (1u << 14);       // SBO_B = RB*BW*16 B: next channel block
    const uint32_t stage0_16 = smem_u32(stage_base) >> 4, stage_16 = (uint32_t)stage_bytes >> 4;
    const uint32_t as_16 = (uint32_t)p.a_stage_bytes >> 4;
    const uint32_t pwz4 = 4u * (uint32_t)p.PWz, bw = (uint32_t)p.BW, nci = (uint32_t)p.Nci;
    const int ksteps = p.BW / 16, RB = p.RB;
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int t = slot; t < p.total_tiles; t += p.ctas_per_group) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t sz16 = stage0_16 + (uint32_t)stage * stage_16;
      const uint32_t sx16 = sz16 + as_16;
      if (elect_one()) {
        for (int r = 0; r < RB; ++r) {
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t b_lo = lo_const | (sx16 + (uint32_t)r * bw + 16u * (uint32_t)k);
            const uint32_t a_row = lo_const | (sz16 + (uint32_t)r * pwz4 + 16u * (uint32_t)k + 2u);
            const uint32_t accum = (first && r == 0 && k == 0) ? 0u : 1u;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              tc_mma_bf16(tmem_base + (uint32_t)kx * nci, a_row - (uint32_t)kx, a_hi, b_lo, b_hi, idesc, accum);
          }
        }
      }
      __syncwarp();
      first = false;
      if (elect_one()) tc_commit(&empty[stage]);
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) tc_commit(done);
    __syncwarp();
  } else {
    // epilogue: lane = (j, co); tap = (2 - j) * 3 + kx; column = kx * Nci + ci
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const bool valid = quarter < 3 && slot < p.total_tiles;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    mbar_wait(done, 0);
    tc_fence_after();
    const int cols = 3 * p.Nci;
    for (int c0 = half * 32; c0 < cols; c0 += 64) {
      uint32_t r[32];
      tc_ld32(tmem_base + lane_base + (uint32_t)c0, r);
      if (valid) {
        const int kx = c0 / p.Nci, ci = cic * p.Nci + (c0 - kx * p.Nci);
        const int tap = (2 - quarter) * 3 + kx;
        float* dst = p.dW + ((long)tap * p.C_in + ci) * p.C_out + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + (long)j * p.C_out, __uint_as_float(r[j]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

int launch_wgrad_t32(const void* X, long x_img_stride, const void* dZ, long dz_img_stride, float* dW9, int N, int C_in, int H,
                     int W, int pad, cudaStream_t stream) {
  WtParams p{};
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  p.dW = dW9; p.C_in = C_in; p.C_out = 32; p.pad = pad;
  p.Nci = C_in < 128 ? C_in : 128;
  p.n_ci_chunks = C_in / p.Nci;
  p.BW = W >= 64 ? 64 : ((W + 15) / 16) * 16;
  p.PWz = p.BW + 2;
  int rb = 4;
  for (;; rb >>= 1) {
    const long ab = (long)(rb + 2) * 4 * p.PWz * 16, bb = (long)(p.Nci / 8) * rb * p.BW * 16;
    if (rb == 1 || (3 * (((ab + 127) & ~127L) + ((bb + 127) & ~127L)) <= 200 * 1024 && rb <= H)) break;
  }
  p.RB = rb; p.RBz = rb + 2;
  p.a_bytes = p.RBz * 4 * p.PWz * 16;
  p.b_bytes = (p.Nci / 8) * p.RB * p.BW * 16;
  p.a_stage_bytes = (p.a_bytes + 127) & ~127;
  p.stage_bytes = p.a_stage_bytes + ((p.b_bytes + 127) & ~127);
  // the A operand always spans 16 (row, channel block) slots = 4 dZ rows from row r (the fourth is garbage that lands in
  // the unused lanes 96..127): the read past the last row of the last stage must stay inside the allocation
  const int overrun = 4 * p.PWz * 16 + 4096;
  const int budget = 227 * 1024 - 512 - overrun;
  p.stages = budget / p.stage_bytes;
  if (p.stages > kWgMaxStages) p.stages = kWgMaxStages;
  UNCL_REQUIRE(p.stages >= 2, "conv3x3_wgrad_tc(t32): tile does not fit shared memory (%d B per stage)", p.stage_bytes);
  int smem_bytes = 256 + p.stages * p.stage_bytes + overrun;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
  p.bands = ceil_div(W, p.BW);
  p.row_tiles = ceil_div(H, p.RB);
  p.tiles_per_img = p.bands * p.row_tiles;
  p.total_tiles = N * p.tiles_per_img;
  const int sms = sm_count();
  p.ctas_per_group = sms / p.n_ci_chunks;
  if (p.ctas_per_group < 1) p.ctas_per_group = 1;
  if (p.ctas_per_group > p.total_tiles) p.ctas_per_group = p.total_tiles;

  CUtensorMap tmx, tmz;
  CUresult r = encode_blocked_bf16(&tmx, X, W, H, C_in / 8, N, x_img_stride, p.BW, p.RB, p.Nci / 8);
  if (r != CUDA_SUCCESS) return uncl_set_error(UNCL_ECUDA, "conv3x3_wgrad_tc(t32): tensor map (X) failed (%d)", (int)r);
  {
    // dZ with the dimensions ordered (x, channel block, y, image): the box lands as [row][channel block][pixel][8]
    EncodeTiledFn encode = get_encode();
    if (!encode) return uncl_set_error(UNCL_ECUDA, "conv3x3_wgrad_tc(t32): cuTensorMapEncodeTiled unavailable");
    const cuuint64_t gdim[4] = {(cuuint64_t)Wo * 2, 4, (cuuint64_t)Ho, (cuuint64_t)N};
    const cuuint64_t gstr[3] = {(cuuint64_t)Ho * Wo * 16, (cuuint64_t)Wo * 16, (cuuint64_t)dz_img_stride * 2};
    const cuuint32_t box[4] = {(cuuint32_t)p.PWz * 2, 4, (cuuint32_t)p.RBz, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    r = encode(&tmz, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(dZ), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return uncl_set_error(UNCL_ECUDA, "conv3x3_wgrad_tc(t32): tensor map (dZ) failed (%d)", (int)r);
  }
  static thread_local int smem_ok = 0, smem_dev = -1;
  cudaError_t e = ensure_smem(conv3x3_wgrad_t32_kernel, smem_bytes, smem_ok, smem_dev);
  if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "conv3x3_wgrad_tc(t32): smem attr: %s", cudaGetErrorString(e));
  conv3x3_wgrad_t32_kernel<<<p.n_ci_chunks * p.ctas_per_group, kWgThreads, smem_bytes, stream>>>(tmx, tmz, p);
  return uncl_check_launch("conv3x3_wgrad_tc(t32)");
}

}  // namespace

// X: bf16 blocked [N][C_in/8][H][W][8] (image stride x_img_stride elements); dZ: bf16 blocked dense
// [N][C_out/8][Ho][Wo][8]; dW9: fp32 [9][C_in][C_out], accumulated (zeroed by the caller).
static int wgrad_tc_impl(const void* X, long x_img_stride, const void* dZ, long dz_img_stride, float* dW9, int N, int C_in,
                         int H, int W, int C_out, int pad, int pointwise, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C_in % 32 == 0 && C_out % 32 == 0 && (pad == 0 || pad == 2) && (C_in <= 128 || C_in % 128 == 0),
               "conv3x3_wgrad_tc: unsupported C_in=%d C_out=%d pad=%d", C_in, C_out, pad);
  UNCL_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(dZ) & 15) == 0 && x_img_stride % 8 == 0,
               "conv3x3_wgrad_tc: operands must be 16-byte aligned");
  if (!pointwise && C_out == 32) return launch_wgrad_t32(X, x_img_stride, dZ, dz_img_stride, dW9, N, C_in, H, W, pad, stream);
  WgParams p{};
  p.dW = dW9;
  p.N = N; p.C_in = C_in; p.C_out = C_out; p.pad = pad;
  p.Ho = pointwise ? H : H + 2 * pad - 2; p.Wo = pointwise ? W : W + 2 * pad - 2;
  UNCL_REQUIRE(p.Ho > 0 && p.Wo > 0, "conv3x3_wgrad_tc: empty output");
  p.taps9 = !pointwise && C_out <= 48;
  p.nky = (p.taps9 || pointwise) ? 1 : 3;
  p.nkx = pointwise ? 1 : 3;
  p.Nc = p.taps9 ? C_out : (C_out < 128 ? C_out : 128);
  UNCL_REQUIRE(C_out % p.Nc == 0 && p.Nc % 32 == 0, "conv3x3_wgrad_tc: unsupported C_out=%d", C_out);
  p.n_co_chunks = C_out / p.Nc;
  p.n_ci_chunks = (C_in + 127) / 128;
  p.Mc_blocks = C_in < 128 ? C_in / 8 : 16;
  p.BW = p.Wo >= 64 ? 64 : ((p.Wo + 15) / 16) * 16;
  p.PWx = pointwise ? p.BW : p.BW + 2;
  const int halo = p.taps9 ? 2 : 0;
  int rb = 4;
  for (;; rb >>= 1) {
    const long xb = (long)p.Mc_blocks * (rb + halo) * p.PWx * 16, zb = (long)(p.Nc / 8) * rb * p.BW * 16;
    if (rb == 1 || (2 * (((xb + 127) & ~127L) + ((zb + 127) & ~127L)) <= 150 * 1024 && rb <= p.Ho)) break;
  }
  p.RB = rb;
  p.RBx = rb + halo;
  p.x_bytes = p.Mc_blocks * p.RBx * p.PWx * 16;
  p.z_bytes = (p.Nc / 8) * p.RB * p.BW * 16;
  p.x_stage_bytes = (p.x_bytes + 127) & ~127;
  p.stage_bytes = p.x_stage_bytes + ((p.z_bytes + 127) & ~127);
  // 16 channel blocks are always read: the overrun past the last stage must stay inside the allocation
  const int overrun = 16 * p.RBx * p.PWx * 16 + 4096;
  const int budget = 227 * 1024 - 512 - overrun;
  p.stages = budget / p.stage_bytes;
  if (p.stages > kWgMaxStages) p.stages = kWgMaxStages;
  UNCL_REQUIRE(p.stages >= 2, "conv3x3_wgrad_tc: tile does not fit shared memory (%d B per stage)", p.stage_bytes);
  int smem_bytes = 256 + p.stages * p.stage_bytes + overrun;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;   // one CTA per SM: each owns all 512 TMEM columns
  p.bands = ceil_div(p.Wo, p.BW);
  p.row_tiles = ceil_div(p.Ho, p.RB);
  p.tiles_per_img = p.bands * p.row_tiles;
  p.total_tiles = N * p.tiles_per_img;
  p.groups = p.nky * p.n_ci_chunks * p.n_co_chunks;
  const int sms = sm_count();
  p.ctas_per_group = sms / p.groups;
  if (p.ctas_per_group < 1) p.ctas_per_group = 1;
  if (p.ctas_per_group > p.total_tiles) p.ctas_per_group = p.total_tiles;

  CUtensorMap tmx, tmz;
  CUresult r = encode_blocked_bf16(&tmx, X, W, H, C_in / 8, N, x_img_stride, p.PWx, p.RBx, p.Mc_blocks);
  if (r != CUDA_SUCCESS) return uncl_set_error(UNCL_ECUDA, "conv3x3_wgrad_tc: tensor map (X) failed (%d)", (int)r);
  r = encode_blocked_bf16(&tmz, dZ, p.Wo, p.Ho, C_out / 8, N, dz_img_stride, p.BW, p.RB, p.Nc / 8);
  if (r != CUDA_SUCCESS) return uncl_set_error(UNCL_ECUDA, "conv3x3_wgrad_tc: tensor map (dZ) failed (%d)", (int)r);
  static thread_local int smem_ok = 0, smem_dev = -1;
  cudaError_t e = ensure_smem(conv3x3_wgrad_tc_kernel, smem_bytes, smem_ok, smem_dev);
  if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "conv3x3_wgrad_tc: smem attr: %s", cudaGetErrorString(e));
  conv3x3_wgrad_tc_kernel<<<p.groups * p.ctas_per_group, kWgThreads, smem_bytes, stream>>>(tmx, tmz, p);
  return uncl_check_launch("conv3x3_wgrad_tc");
}

extern "C" int uncl_conv3x3_wgrad_tc(const void* X, long x_img_stride, const void* dZ, float* dW9, int N, int C_in, int H,
                                     int W, int C_out, int pad, cudaStream_t stream) {
  return wgrad_tc_impl(X, x_img_stride, dZ, (long)C_out * (H + 2 * pad - 2) * (W + 2 * pad - 2), dW9, N, C_in, H, W, C_out, pad,
                       0, stream);
}

// The same with an explicit dZ image stride (elements): dZ may be a channel slice of a wider tensor - the exact path's
// three-term split passes the hi / lo thirds of [dz_hi | dz_hi | dz_lo] (uncl_split_bf16) without copying them out.
extern "C" int uncl_conv3x3_wgrad_tc_strided(const void* X, long x_img_stride, const void* dZ, long dz_img_stride, float* dW9,
                                             int N, int C_in, int H, int W, int C_out, int pad, cudaStream_t stream) {
  UNCL_REQUIRE(dz_img_stride % 8 == 0, "conv3x3_wgrad_tc_strided: dZ stride must be a multiple of 8");
  return wgrad_tc_impl(X, x_img_stride, dZ, dz_img_stride, dW9, N, C_in, H, W, C_out, pad, 0, stream);
}

// Pointwise (1x1 / GEMM) weight gradient on the same kernel with one tap:  dW[ci][co] += sum_pix X[pix, ci] * dZ[pix, co].
// The k2 s2 up-convolution's weight gradient is this GEMM over the space-to-depth output gradient (C_out = 4C).
// X: bf16 blocked [N][C_in/8][H][W][8]; dZ: bf16 blocked [N][C_out/8][H][W][8]; both with their own image strides.
extern "C" int uncl_pw_wgrad_tc(const void* X, long x_img_stride, const void* dZ, long dz_img_stride, float* dW, int N,
                                int C_in, int C_out, int H, int W, cudaStream_t stream) {
  UNCL_REQUIRE(dz_img_stride % 8 == 0, "pw_wgrad_tc: dZ stride must be a multiple of 8");
  return wgrad_tc_impl(X, x_img_stride, dZ, dz_img_stride, dW, N, C_in, H, W, C_out, 0, 1, stream);
}
