// 3x3 stride-1 convolution (and ConvTranspose 3x3 s1 as its padded/flipped twin) as an implicit GEMM on the
// Blackwell tensor cores: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM), operands staged by
// TMA / bulk async copies, persistent warp-specialised CTAs (one per SM: 1 TMA producer warp, 1 MMA-issuing warp,
// 16 epilogue warps) with a double-buffered TMEM accumulator.  Narrow layers with a long K loop (C_out <= 64,
// C_in >= 64) are routed to the kx-merged variant in conv_tc_merged.cu.
//
// Reference operator: models/unet_multi_filters/unet_parts.py:57-87 (double_conv), :126-141, :183-193
// (ConvTranspose2d 3x3 s1 p0 == conv over a 2-px zero-padded input with flipped kernels; the zero padding is TMA
// out-of-bounds fill), :319-322 (skip operators x^2 / sqrt(x+eps), emitted by the producing layer's epilogue),
// :338-345 + Unet_singleFrame.py:207-209 (1x1 out conv + sigmoid, fused into the last conv's epilogue).
//
// How the conv becomes a GEMM without im2col and without re-loading the input once per tap:
//   * activations are C8-blocked [N][C/8][H][W][8] bf16, so a TMA box {8ch, PW px, PH rows, 2 blocks} lands in
//     shared memory as [2][PH*PW][8] - which IS the no-swizzle K-major UMMA operand layout (core matrix = 8
//     consecutive pixels x 16 B, SBO = 128 B between 8-pixel groups, LBO = PH*PW*16 B between the two 8-channel
//     halves of a K=16 step).
//   * rows of the A operand are 128 CONSECUTIVE pixels of that halo tile.  A filter tap (ky,kx) is then just a
//     different descriptor start address (+ (ky*PW+kx)*16 B) into the same tile: 9 MMAs per K-step reuse one
//     halo tile, so shared memory is filled ~1.3x the input instead of 9x.
//   * where 128 consecutive pixels wrap around the end of a tile row the accumulator row is garbage; every
//     A row only feeds its own D row, so those rows are simply masked in the epilogue.
// Weights are pre-packed per 16-channel K-chunk as [9 taps][2][NT][8] bf16 (K-major B operand, LBO = NT*16 B),
// fetched with one cp.async.bulk per stage.
#include <cstdlib>
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int kThreads = 640;          // warp 0: TMA producer, warp 1: MMA issuer, warps 2-17: epilogue, warps 18-19: further MMA issuers
constexpr int kMma2Warp = 18;          // first of the extra issuing warps (TcParams::mma_warps == 2 or 3: the M blocks of a tile are dealt
                                       // out; 20 warps = 5 per scheduler keep 96 registers per thread, a 21st would cap them at 80)
constexpr int kEpiWarps = 16;          // four warps per TMEM lane quarter; (M block, 32-column chunk) units dealt round-robin
constexpr int kEpiPerQuarter = kEpiWarps / 4;
constexpr int kAccCols = 256;          // TMEM columns per accumulator stage when double-buffered (2 stages = 512 = all of TMEM)
constexpr int kMaxStages = 6;

struct TcParams {
  const bf16* w;            // packed weights [NS][nchunk][ntaps][2][NT][8]
  const float* bias;        // [C_out]
  void* out;                // blocked output, bf16 or fp32 (may be null when fuse_outc)
  int out_f32;              // 1: `out` is fp32 (training path: fp32 tensors between the tensor-core convs)
  long out_img_stride;
  const float* outc_w;      // [C_out] (fuse_outc)
  const float* outc_b;
  float* out_img;           // [N][Ho][Wo] fp32 (fuse_outc)
  float* out_logit;         // optional
  int N, C_in, C_out, Ho, Wo, pad;
  int x0;                   // first output column of this launch (the row kernel's trailing columns, conv_tc_rows.cu); bands cover [x0, Wo)
  int NT, NS;               // N tile (<=128) and number of N splits
  int MB;                   // M blocks (128 pixels each) per tile
  int PW, PH;               // TMA box (halo tile) extent in pixels; PW is also the flattening pitch of a band
  int BW, nbands;           // output columns per band (PW = BW + 2), bands per image row
  int band_total;           // Ho * PW: flattened (pitch PW) output positions of one band
  int tiles_per_band, tiles_per_img, num_items;
  int nchunk, ntaps, stages;
  int ksteps;               // K = 16 steps per pipeline stage (pointwise GEMMs: up to 4 = 64 input channels per barrier round trip)
  int nacc, acc_cols;       // accumulator stages (2 x 256 columns, or 1 x 512 for K-heavy layers: see launch_tc)
  int mma_warps;            // 1, 2 or 3 issuing warps: warp slot w issues the M blocks [w * per, (w + 1) * per), per = ceil(MB / warps)
  int a_box_bytes, a_stage_bytes, b_stage_bytes, stage_bytes;
  int act, emit_skip, fuse_outc;
  int sh_C, sh_H2, sh_W2;   // pixel-shuffle epilogue (ConvTranspose k2 s2): channels, target extent (replicate pad)
  // pointwise epilogue (EPI 2): out = scale[n] * act(acc + bias) + res
  const void* res;          // blocked, C_out channels, fp32 or bf16 (null: none)
  long res_img_stride;
  int res_f32;
  const float* scale;       // per-image factor (DropPath) or null
  const bf16* mask;         // data-gradient mode: out = (mask > 0) ? acc : 0 - the ReLU of the layer that PRODUCED this
  long mask_img_stride;     // conv's input, applied where its gradient is formed (blocked bf16, C_out channels, Ho x Wo)
  int ns_per_group;         // grouped 1x1 conv: N splits per group (K range of a split = its group's input channels)
  unsigned long long* dbg;  // diagnostics (uncl_conv_tc_set_debug): cycle counters of the three roles, or null
  unsigned long long m_PW;  // 2^40 / PW + 1 (fastdiv_pw)
  int probe_noload;         // timing probe: the producer only loads the first `stages` chunks, then re-signals stale stages
};

// ---------------------------------------------------------------- tile geometry shared by the three roles
// A tile is MB*128 CONSECUTIVE positions of one band, flattened with pitch PW (= BW + 2): position q -> output
// (q / PW, band*BW + q % PW); the last two positions of every pitch row are the wrap-around garbage columns.
struct Item {
  int n, ns;          // image, N split
  int band;           // column band
  int bx, by;         // TMA box origin (input pixel coordinates, may be negative: zero fill)
  int moff0;          // start pixel (inside the box) of M block 0
  int mb_act;         // active M blocks
  int q0;             // first flattened position
};

struct Geo {  // hot scalars of TcParams, hoisted into registers
  int NS, tiles_per_img, tiles_per_band, MB, PW, BW, pad, band_total, x0;
};

// q / PW with the host-computed magic m = 2^40 / PW + 1 (exact for q < 2^24, PW <= 128)
__device__ __forceinline__ int fastdiv_pw(int q, unsigned long long m) {
  return (int)(((unsigned long long)(uint32_t)q * m) >> 40);
}

__device__ __forceinline__ Item decode_item(const Geo& g, int item) {
  Item it;
  const int tile = item / g.NS;
  it.ns = item - tile * g.NS;
  it.n = tile / g.tiles_per_img;
  const int t = tile - it.n * g.tiles_per_img;
  it.band = t / g.tiles_per_band;
  const int tb = t - it.band * g.tiles_per_band;
  it.q0 = tb * 128 * g.MB;
  const int y0 = it.q0 / g.PW;
  it.bx = g.x0 + it.band * g.BW - g.pad;
  it.by = y0 - g.pad;
  it.moff0 = it.q0 - y0 * g.PW;
  it.mb_act = min(g.MB, (g.band_total - it.q0 + 127) / 128);
  return it;
}

// 9 taps x MB blocks of one K chunk, fully unrolled: descriptor low words are base + compile-time-shaped offsets
// B0..B1: the blocks of the tile this warp issues (each block owns its accumulator columns, so two issuing warps
// never touch the same TMEM columns and need no ordering between their instruction streams)
// NB: the number of blocks this warp issues; d0 / a_row already point at its first block
template <int NB>
__device__ __forceinline__ void issue_taps9(uint32_t d0, uint32_t a_row, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                            uint32_t pw, uint32_t nt, uint32_t b_tap_16, uint32_t first) {
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        tc_mma_bf16(d0 + (uint32_t)b * nt, a_row + (uint32_t)ky * pw + (uint32_t)(kx + b * 128), desc_hi,
                    b_lo + (uint32_t)(ky * 3 + kx) * b_tap_16, desc_hi, idesc, (ky > 0 || kx > 0) ? 1u : first);
      }
    }
  }
}

// EPI 0: conv epilogue (bias, ReLU, optional skip emission / fused 1x1 out conv + sigmoid)
// EPI 1: pixel-shuffle epilogue of ConvTranspose k2 s2 (columns = (dy, dx, co)), replicate pad into (H2 x W2)
// EPI 2: pointwise (1x1, optionally grouped) conv epilogue: bias, ReLU / GELU / identity, per-image scale, residual
template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // carve: [stages x (A | B)] | barriers | tmem ptr | bias
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);  // keeps the shared address space
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes + 128);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);  // [C_out] (+ [C_out] outc weights)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Geo geo = {p.NS, p.tiles_per_img, p.tiles_per_band, p.MB, p.PW, p.BW, p.pad, p.band_total, p.x0};
  const int num_items = p.num_items, nchunk = p.nchunk, stages = p.stages, stage_bytes = p.stage_bytes;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (uint32_t)p.mma_warps); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], (uint32_t)p.mma_warps); mbar_init(&tempty[s], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < p.C_out; i += kThreads) {
    s_bias[i] = p.bias ? p.bias[i] : 0.f;
    if (p.fuse_outc) s_bias[p.C_out + i] = p.outc_w[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (uint32_t)(p.a_box_bytes + p.b_stage_bytes), b_bytes = (uint32_t)p.b_stage_bytes;
      const int a_stage_bytes = p.a_stage_bytes;
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w);
      const int ns_per_group = p.ns_per_group;
      unsigned long long* const dbg = UNCL_PROBE(1, 1) ? p.dbg : nullptr;
      long long w_empty = 0;
      const long long t_begin = dbg ? clock64() : 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const Item it = decode_item(geo, item);
        const uint8_t* wsrc = wbase + (size_t)it.ns * nchunk * b_bytes;
        const int kblk0 = (it.ns / ns_per_group) * nchunk * 2 * p.ksteps;   // first input channel block of this split's group
        for (int ch = 0; ch < nchunk; ++ch) {
          if (UNCL_PROBE(p.probe_noload, 4)) continue;
          const long long tw0 = dbg ? clock64() : 0;
          mbar_wait(&empty[stage], phase ^ 1);
          if (dbg) w_empty += clock64() - tw0;
          uint8_t* sa = stage_base + (size_t)stage * stage_bytes;
          if ((UNCL_PROBE(p.probe_noload, 1)) && (item != (int)blockIdx.x || ch >= stages)) { mbar_arrive(&full[stage]); }
          else {
          mbar_expect_tx(&full[stage], tx_bytes);
          tma_load_4d(sa, &tmap, &full[stage], it.bx * 2, it.by, kblk0 + ch * 2 * p.ksteps, it.n);
          bulk_load(sa + a_stage_bytes, wsrc + (size_t)ch * b_bytes, b_bytes, &full[stage]);
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg) {
        atomicAdd(dbg + 0, (unsigned long long)(clock64() - t_begin));
        atomicAdd(dbg + 1, (unsigned long long)w_empty);
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp >= kMma2Warp) {
    // =============================== MMA issuer(s) ===============================
    // the whole warp walks the (uniform) loop so that descriptors live in uniform registers; one elected lane issues.
    // The issuing warp is a serial, latency-bound instruction stream (~85 cycles per MMA measured against 64 of math at
    // N = 128, ~57 against 40 at N = 32): with mma_warps == 2 / 3 further warps on the other schedulers issue their share
    // of every tile's M blocks (each block has its own accumulator columns: the streams need no ordering between them).
    const int slot = warp == 1 ? 0 : warp - kMma2Warp + 1;
    const bool second = slot != 0;
    if (slot < p.mma_warps) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((128u >> 4) << 24);
      // K-major no-swizzle descriptors: lo = (addr >> 4) | (LBO >> 4) << 16 ; hi = (SBO >> 4) | version 1 << 14
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo_const = ((uint32_t)(p.PH * p.PW) & 0x3fffu) << 16;   // LBO_A = PH*PW*16 B
      const uint32_t b_lo_const = ((uint32_t)p.NT & 0x3fffu) << 16;            // LBO_B = NT*16 B
      const uint32_t stage0_16 = smem_u32(stage_base) >> 4, stage_16 = (uint32_t)stage_bytes >> 4;
      const uint32_t a_bytes_16 = (uint32_t)p.a_stage_bytes >> 4, b_tap_16 = (uint32_t)(2 * p.NT);
      const uint32_t mstep = 128u;
      const uint32_t pw = (uint32_t)p.PW, nt = (uint32_t)p.NT;
      const bool taps9 = p.ntaps == 9;
      const int nacc = p.nacc, acc_cols = p.acc_cols;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      unsigned long long* const dbg = (UNCL_PROBE(1, 1) && !second) ? p.dbg : nullptr;
      long long w_full = 0, w_tempty = 0;
      const long long t_begin = dbg ? clock64() : 0;
      const unsigned long long g_begin = dbg ? globaltimer_ns() : 0ull;
      const uint32_t per = (uint32_t)((p.MB + p.mma_warps - 1) / p.mma_warps);
      const uint32_t blk_lo = min((uint32_t)slot * per, (uint32_t)p.MB), blk_hi = min(blk_lo + per, (uint32_t)p.MB);
      const uint32_t nblk = blk_hi - blk_lo;   // blocks this warp issues (launch-uniform)
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const Item it = decode_item(geo, item);
        const long long tw0 = dbg ? clock64() : 0;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        if (dbg) w_tempty += clock64() - tw0;
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * acc_cols);
        // runtime block range of this warp (pointwise GEMMs and unusual tile heights)
        const uint32_t mb_lo = blk_lo;
        const uint32_t mb = min((uint32_t)it.mb_act, blk_hi);
        for (int ch = 0; ch < nchunk; ++ch) {
          const long long tw1 = dbg ? clock64() : 0;
          if (!(UNCL_PROBE(p.probe_noload, 4))) mbar_wait(&full[stage], phase);
          if (dbg) w_full += clock64() - tw1;
          tc_fence_after();
          const uint32_t sa16 = stage0_16 + (uint32_t)stage * stage_16;
          uint32_t a_row = a_lo_const | (sa16 + (uint32_t)it.moff0);
          uint32_t b_lo = b_lo_const | (sa16 + a_bytes_16);
          if (taps9) {
            // one elected region per K chunk: 9 taps x MB MMAs issued back to back by the same thread.  For the tile
            // heights the generator uses the region is straight-line code (issue_taps9<MB>): the issuing warp is a
            // serial, latency-bound instruction stream and the remainder path of a runtime `mb` loop (R2UR / LDCU per
            // tap) cost ~240 cycles per tap at MB = 2 - twice the 128 cycles its two N = 128 MMAs execute in.
            // Every block of the tile is issued; blocks past the end of a band read zero-filled rows of the halo box
            // and are skipped by the epilogue.
            if (elect_one()) {
              const uint32_t first = ch > 0 ? 1u : 0u;
              const uint32_t dw = d0 + blk_lo * nt, aw = a_row + blk_lo * mstep;   // this warp's first block
              switch (nblk) {
                case 0: break;
                case 1: issue_taps9<1>(dw, aw, b_lo, desc_hi, idesc, pw, nt, b_tap_16, first); break;
                case 2: issue_taps9<2>(dw, aw, b_lo, desc_hi, idesc, pw, nt, b_tap_16, first); break;
                case 3: issue_taps9<3>(dw, aw, b_lo, desc_hi, idesc, pw, nt, b_tap_16, first); break;
                case 4: issue_taps9<4>(dw, aw, b_lo, desc_hi, idesc, pw, nt, b_tap_16, first); break;
                case 8: issue_taps9<8>(dw, aw, b_lo, desc_hi, idesc, pw, nt, b_tap_16, first); break;
                default:
#pragma unroll
                  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                      const uint32_t accum = (ky > 0 || kx > 0) ? 1u : first;
                      uint32_t a_lo = a_row + (uint32_t)kx + mb_lo * mstep, d = d0 + mb_lo * nt;
                      for (uint32_t b = mb_lo; b < mb; ++b) {
                        tc_mma_bf16(d, a_lo, desc_hi, b_lo, desc_hi, idesc, accum);
                        a_lo += mstep;
                        d += nt;
                      }
                      b_lo += b_tap_16;
                    }
                    a_row += pw;
                  }
              }
            }
            __syncwarp();
          } else {
            // pointwise GEMM: a stage holds `ksteps` K = 16 steps (their A halves / B tiles follow each other in the stage)
            if (elect_one()) {
              const uint32_t a_kstep_16 = (uint32_t)(2 * p.PH * p.PW), ksteps = (uint32_t)p.ksteps;
              for (uint32_t kc = 0; kc < ksteps; ++kc) {
                uint32_t a_lo = a_row + kc * a_kstep_16 + mb_lo * mstep, d = d0 + mb_lo * nt;
                const uint32_t accum = (ch > 0 || kc > 0) ? 1u : 0u;
                for (uint32_t b = mb_lo; b < mb; ++b) {
                  tc_mma_bf16(d, a_lo, desc_hi, b_lo + kc * b_tap_16, desc_hi, idesc, accum);
                  a_lo += mstep;
                  d += nt;
                }
              }
            }
            __syncwarp();
          }
          if (!(UNCL_PROBE(p.probe_noload, 4)) && elect_one()) tc_commit(&empty[stage]);
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit(&tfull[acc]);
        __syncwarp();
        if (++acc == nacc) { acc = 0; acc_phase ^= 1; }
      }
      if (dbg && lane == 0) {
        atomicAdd(dbg + 2, (unsigned long long)(clock64() - t_begin));
        atomicAdd(dbg + 3, (unsigned long long)w_full);
        atomicAdd(dbg + 4, (unsigned long long)w_tempty);
        atomicAdd(dbg + 8, globaltimer_ns() - g_begin);   // with [2]: the SM clock this launch actually ran at
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue (16 warps: 4 TMEM lane quarters x 4 round-robin unit slots) ========
    const int quarter = warp & 3;
    const int k4 = (warp - 2) >> 2;     // this warp takes the units u = (block, 32-column chunk) with u % 4 == k4
    const int cpb = p.NT >> 5;          // 32-column chunks per M block
    const int row = quarter * 32 + lane;
    const int Ho = p.Ho, Wo = p.Wo, NT = p.NT, C_out = p.C_out;
    const long cb_stride = (long)Ho * Wo * 8;
    const long skip2 = (long)(2 * (C_out / 8)) * cb_stride, skip3 = (long)(3 * (C_out / 8)) * cb_stride;
    const float act_floor = (p.act == UNCL_ACT_RELU) ? 0.f : -INFINITY;   // ReLU or identity, branch-free
    const bool emit_skip = p.emit_skip != 0, fuse_outc = p.fuse_outc != 0;
    bf16* const out = reinterpret_cast<bf16*>(p.out);
    float* const outf = reinterpret_cast<float*>(p.out);
    const bool out_f32 = p.out_f32 != 0;
    const long out_img_stride = p.out_img_stride;
    float* const out_img = p.out_img;
    float* const out_logit = p.out_logit;
    const float outc_b = fuse_outc ? __ldg(p.outc_b) : 0.f;
    const bf16* const mask = p.mask;
    const long mask_img_stride = p.mask_img_stride;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int nacc = p.nacc, acc_cols = p.acc_cols;
    int acc = 0;
    uint32_t acc_phase = 0;
    unsigned long long* const dbg = (UNCL_PROBE(1, 1) && warp == 2 && lane == 0) ? p.dbg : nullptr;
    long long w_tfull = 0;
    const long long t_begin = dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item it = decode_item(geo, item);
      const long long tw0 = dbg ? clock64() : 0;
      mbar_wait(&tfull[acc], acc_phase);
      if (dbg) w_tfull += clock64() - tw0;
      tc_fence_after();
      if constexpr (EPI == 0) {
        const int cbase0 = it.ns * NT;
        const int units = ((UNCL_PROBE(p.probe_noload, 2)) ? 0 : it.mb_act) * cpb;
        for (int u = k4, b = 0, cc = k4; u < units; u += kEpiPerQuarter, cc += kEpiPerQuarter) {
          while (cc >= cpb) { cc -= cpb; ++b; }
          const int c0 = cc * 32;
          const int q = it.q0 + b * 128 + row;
          const int oy = fastdiv_pw(q, p.m_PW), xl = q - oy * geo.PW;
          const int ox = geo.x0 + it.band * geo.BW + xl;
          const bool valid = (oy < Ho) && (xl < geo.BW) && (ox < Wo);
          const long pix = (long)oy * Wo + ox;
          float logit = 0.f;
          {
            uint32_t r[32];
            tc_ld32(tmem_base + lane_base + (uint32_t)(acc * acc_cols + b * NT + c0), r);
            if (valid) {
              const float* bias = s_bias + cbase0 + c0;
  #pragma unroll
              for (int g = 0; g < 4; ++g) {
                float v[8];
  #pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(__uint_as_float(r[g * 8 + j]) + bias[g * 8 + j], act_floor);
                if (mask != nullptr) {
                  float m[8];
                  load8(mask + (long)it.n * mask_img_stride + (long)(cbase0 / 8 + c0 / 8 + g) * cb_stride + pix * 8, m);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = m[j] > 0.f ? v[j] : 0.f;
                }
                if (out != nullptr) {
                  const long off = (long)it.n * out_img_stride + (long)(cbase0 / 8 + c0 / 8 + g) * cb_stride + pix * 8;
                  float s2[8], s3[8];
                  if (emit_skip) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s2[j] = v[j] * v[j]; s3[j] = fast_sqrt(v[j] + 1e-8f); }
                  }
                  if (out_f32) {
                    store8(outf + off, v);
                    if (emit_skip) { store8(outf + off + skip2, s2); store8(outf + off + skip3, s3); }
                  } else {
                    store8(out + off, v);
                    if (emit_skip) { store8(out + off + skip2, s2); store8(out + off + skip3, s3); }
                  }
                }
                if (fuse_outc) {
                  const float* ow = s_bias + C_out + cbase0 + c0 + g * 8;
  #pragma unroll
                  for (int j = 0; j < 8; ++j) logit = fmaf(v[j], ow[j], logit);
                }
              }
            }
          }
          if (fuse_outc && valid) {   // C_out == 32: the unit holds all channels of its pixels
            logit += outc_b;
            const long o = (long)it.n * Ho * Wo + pix;
            if (out_logit) out_logit[o] = logit;
            out_img[o] = 1.f / (1.f + __expf(-logit));
          }
        }
      } else if constexpr (EPI == 2) {
        const int cbase0 = it.ns * NT;
        const int act = p.act;
        const float sc = p.scale ? __ldg(p.scale + it.n) : 1.f;
        const bf16* const resb = reinterpret_cast<const bf16*>(p.res);
        const float* const resf = reinterpret_cast<const float*>(p.res);
        const bool res_f32 = p.res_f32 != 0;
        const long res_img_stride = p.res_img_stride;
        const int units = it.mb_act * cpb;
        for (int u = k4, b = 0, cc = k4; u < units; u += kEpiPerQuarter, cc += kEpiPerQuarter) {
          while (cc >= cpb) { cc -= cpb; ++b; }
          const int c0 = cc * 32;
          const int q = it.q0 + b * 128 + row;
          const int oy = fastdiv_pw(q, p.m_PW), xl = q - oy * geo.PW;
          const int ox = geo.x0 + it.band * geo.BW + xl;
          const bool valid = (oy < Ho) && (xl < geo.BW) && (ox < Wo);
          const long pix = (long)oy * Wo + ox;
          {
            uint32_t r[32];
            tc_ld32(tmem_base + lane_base + (uint32_t)(acc * acc_cols + b * NT + c0), r);
            if (valid) {
              const float* bias = s_bias + cbase0 + c0;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = sc * apply_act(__uint_as_float(r[g * 8 + j]) + bias[g * 8 + j], act);
                const long cpix = (long)(cbase0 / 8 + c0 / 8 + g) * cb_stride + pix * 8;
                if (resf != nullptr) {
                  float rr[8];
                  if (res_f32) load8(resf + (long)it.n * res_img_stride + cpix, rr);
                  else load8(resb + (long)it.n * res_img_stride + cpix, rr);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] += rr[j];
                }
                if (mask != nullptr) {
                  float m[8];
                  load8(mask + (long)it.n * mask_img_stride + cpix, m);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = m[j] > 0.f ? v[j] : 0.f;
                }
                const long off = (long)it.n * out_img_stride + cpix;
                if (out_f32) store8(outf + off, v);
                else store8(out + off, v);
              }
            }
          }
        }
      } else {
        // ConvTranspose k2 s2: column j = pos * C + co, pos = dy * 2 + dx; out pixel (2y+dy, 2x+dx) (+ replicate pad)
        // C % 32 == 0, so a 32-column chunk lies inside one pos: all index arithmetic is per chunk, not per store
        const int C = p.sh_C, H2 = p.sh_H2, W2 = p.sh_W2, Hi = Ho, Wi = Wo;
        const int padT = (H2 - 2 * Hi) / 2, padL = (W2 - 2 * Wi) / 2;
        const bool padded = (H2 != 2 * Hi) || (W2 != 2 * Wi);
        const long cb2 = (long)H2 * W2 * 8;
        bf16* const out_img_n = out + (long)it.n * out_img_stride;
        float* const outf_img_n = outf + (long)it.n * out_img_stride;
        // Paired mode (bf16, no replicate pad, both dx of a dy inside one N split, even W2): a thread takes the dx = 0 and
        // dx = 1 columns of the same 32 channels, whose outputs are neighbouring pixels = 32 contiguous bytes per channel
        // block -> one 256-bit store per lane, a warp writes 1 KB contiguous (the 16-byte stores at a 32-byte stride of
        // the unpaired path half-fill every sector they touch).
        const bool paired = !out_f32 && !padded && 2 * C <= NT && NT == 128 && (W2 & 1) == 0;
        const int units = paired ? it.mb_act * 2 : it.mb_act * cpb;
        if (paired) {
          for (int u = k4; u < units; u += kEpiPerQuarter) {
            const int b = u >> 1, pu = u & 1;
            const int colA = C == 32 ? pu * 64 : pu * 32;   // dx = 0 columns; dx = 1 is C columns further
            const int q = it.q0 + b * 128 + row;
            const int y = fastdiv_pw(q, p.m_PW), x = q - y * geo.PW;
            uint32_t ra[32], rb[32];
            tc_ld32_nowait(tmem_base + lane_base + (uint32_t)(acc * acc_cols + b * NT + colA), ra);
            tc_ld32(tmem_base + lane_base + (uint32_t)(acc * acc_cols + b * NT + colA + C), rb);
            if (y < Hi) {
              const int j0 = it.ns * NT + colA;
              const int pos = j0 / C, co0 = j0 - pos * C;   // pos = 2 * dy
              const int Y = 2 * y + (pos >> 1), X = 2 * x;
              const float* bias = s_bias + co0;
              bf16* o = out_img_n + (long)(co0 / 8) * cb2 + ((long)Y * W2 + X) * 8;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float b0 = bias[g * 8 + 2 * j], b1 = bias[g * 8 + 2 * j + 1];
                  w[j] = pack_bf16x2(__uint_as_float(ra[g * 8 + 2 * j]) + b0, __uint_as_float(ra[g * 8 + 2 * j + 1]) + b1);
                  w[4 + j] = pack_bf16x2(__uint_as_float(rb[g * 8 + 2 * j]) + b0, __uint_as_float(rb[g * 8 + 2 * j + 1]) + b1);
                }
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + g * cb2), "r"(w[0]), "r"(w[1]),
                             "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
              }
            }
          }
        } else
        for (int u = k4, b = 0, cc = k4; u < units; u += kEpiPerQuarter, cc += kEpiPerQuarter) {
          while (cc >= cpb) { cc -= cpb; ++b; }
          const int c0 = cc * 32;
          const int q = it.q0 + b * 128 + row;
          const int y = fastdiv_pw(q, p.m_PW), x = q - y * geo.PW;
          const bool valid = y < Hi;
          {
            uint32_t r[32];
            tc_ld32(tmem_base + lane_base + (uint32_t)(acc * acc_cols + b * NT + c0), r);
            if (valid) {
              const int j0 = it.ns * NT + c0;
              const int pos = j0 / C, co0 = j0 - pos * C;
              const int Y = 2 * y + (pos >> 1), X = 2 * x + (pos & 1);
              const float* bias = s_bias + co0;
              const long o0 = (long)(co0 / 8) * cb2;
              if (!padded) {
                const long o = o0 + ((long)Y * W2 + X) * 8;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  float v[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) + bias[g * 8 + j];
                  if (out_f32) store8(outf_img_n + o + g * cb2, v);
                  else store8(out_img_n + o + g * cb2, v);
                }
              } else {
                const int y_lo = (Y == 0) ? 0 : Y + padT, y_hi = (Y == 2 * Hi - 1) ? H2 - 1 : Y + padT;
                const int x_lo = (X == 0) ? 0 : X + padL, x_hi = (X == 2 * Wi - 1) ? W2 - 1 : X + padL;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  float v[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) + bias[g * 8 + j];
                  for (int yy = y_lo; yy <= y_hi; ++yy)
                    for (int xx = x_lo; xx <= x_hi; ++xx) {
                      const long o = o0 + g * cb2 + ((long)yy * W2 + xx) * 8;
                      if (out_f32) store8(outf_img_n + o, v);
                      else store8(out_img_n + o, v);
                    }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == nacc) { acc = 0; acc_phase ^= 1; }
    }
    if (dbg) {
      atomicAdd(dbg + 5, (unsigned long long)(clock64() - t_begin));
      atomicAdd(dbg + 6, (unsigned long long)w_tfull);
      atomicAdd(dbg + 7, 1ull);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ---------------------------------------------------------------- host side
}  // namespace

namespace {

#ifdef UNCL_PROBES
unsigned long long* g_dbg = nullptr;   // probe build only (tools/), see uncl_conv_tc_set_debug
static const char* probe_env(const char* name) { return getenv(name); }
#else
constexpr unsigned long long* g_dbg = nullptr;
static const char* probe_env(const char*) { return nullptr; }   // the product library reads no environment variable
#endif

// fills the tile geometry for an (ntaps = 9: 3x3 with halo | ntaps = 1: pointwise GEMM) problem and launches
// pure host arithmetic (no CUDA calls): also behind uncl_conv3x3_tc_plan
int plan_tc(TcParams& p, int N, int C_in, int bias_floats, const char* what, int* smem_bytes_out) {
  const int halo = p.ntaps == 9 ? 2 : 0;
  if (p.nchunk <= 0) p.nchunk = C_in / 16;
  p.ksteps = 1;
  // Accumulator staging.  An SS-mode MMA costs max(N/2, (4096 + 32 N)/128) cycles whatever the previous one used
  // (tools/mma_probe.cu); what the tile height buys is issue-side: the MMA-issuing warp is one serial, latency-bound
  // instruction stream and every K chunk costs it a barrier round trip (wait full, fence, elect, commit: ~300+ cycles).
  // More M blocks per tile put more MMAs behind each round trip, so K-heavy layers take all 512 TMEM columns for ONE
  // accumulator stage (twice the M blocks; the exposed epilogue is < 3 % of such a tile), short-K layers keep two
  // stages so that the epilogue of tile i overlaps the MMAs of tile i+1.  Measured per layer (profiles/README.md): the
  // single stage pays off for N = 128 with K >= 512*9 (up0.conv: 167 -> 129 us); for narrower N the exposed epilogue
  // and the coarser tiles cost more than the amortisation gains.
  p.nacc = (p.NT == 128 && C_in >= 512 && p.ntaps == 9 && probe_env("UNCL_PROBE_DOUBLE_ACC") == nullptr) ? 1 : 2;
  p.acc_cols = p.nacc == 1 ? 512 : kAccCols;
  const int mb_max = p.acc_cols / p.NT;
  // column bands: the TMA box row is PW pixels = 2*PW 8-byte elements and a box dimension holds <= 256 elements
  p.nbands = ceil_div(p.Wo - p.x0, 128 - halo);
  p.BW = ceil_div(p.Wo - p.x0, p.nbands);
  p.PW = p.BW + halo;
  p.band_total = p.Ho * p.PW;
  {
    // equal-sized tiles: the band's M blocks are spread evenly over the smallest number of tiles that fit TMEM
    const int blocks = ceil_div(p.band_total, 128);
    const int tiles = ceil_div(blocks, mb_max);
    p.MB = ceil_div(blocks, tiles);
  }
  if (const char* e = probe_env("UNCL_PROBE_MB")) {   // timing probe: force the M blocks per tile
    const int want = atoi(e);
    if (want >= 1 && want < p.MB) p.MB = want;
  }
  p.mma_warps = p.MB >= 6 ? 3 : (p.MB >= 2 ? 2 : 1);
  if (const char* e = probe_env("UNCL_MMA_WARPS")) { const int w = atoi(e); if (w >= 1 && w <= 3 && w <= p.MB) p.mma_warps = w; }
  p.PH = (p.PW - 1 + 128 * p.MB - 1) / p.PW + 1 + halo;
  p.tiles_per_band = ceil_div(p.band_total, 128 * p.MB);
  p.tiles_per_img = p.nbands * p.tiles_per_band;
  UNCL_REQUIRE(p.PW <= 128 && p.PH <= 256, "%s: halo tile too large (%d x %d)", what, p.PW, p.PH);
  p.num_items = N * p.tiles_per_img * p.NS;
  p.m_PW = (1ull << 40) / (unsigned)p.PW + 1;
  const int tail = 128 + (2 * kMaxStages + 4) * 8 + 16 + bias_floats * 4 + 256;
  const int budget = 227 * 1024 - tail;
  // pointwise GEMMs are short, latency-bound launches: fewer, fatter pipeline stages (up to 64 input channels) cut the
  // issuing warp's barrier round trips - as long as four stages still fit
  if (p.ntaps == 1 && probe_env("UNCL_PROBE_PW_KSTEPS1") == nullptr) {
    for (int ks = 4; ks > 1; ks >>= 1) {
      const int sb = ((2 * ks * p.PH * p.PW * 16 + 127) & ~127) + ks * 2 * p.NT * 16;
      if (p.nchunk % ks == 0 && 4 * sb <= budget) { p.ksteps = ks; break; }
    }
    p.nchunk /= p.ksteps;
  }
  p.a_box_bytes = 2 * p.ksteps * p.PH * p.PW * 16;
  p.a_stage_bytes = (p.a_box_bytes + 127) & ~127;
  p.b_stage_bytes = p.ksteps * p.ntaps * 2 * p.NT * 16;
  p.stage_bytes = p.a_stage_bytes + p.b_stage_bytes;  // both multiples of 128
  p.stages = budget / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  UNCL_REQUIRE(p.stages >= 2, "%s: tile does not fit shared memory (%d B per stage)", what, p.stage_bytes);
  int smem_bytes = p.stages * p.stage_bytes + tail;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;  // force one CTA per SM (each CTA owns all 512 TMEM columns)
  *smem_bytes_out = smem_bytes;
  return UNCL_OK;
}

int launch_tc(TcParams& p, const void* in, long in_img_stride, int N, int C_in, int H, int W, int epi, int bias_floats,
              const char* what, cudaStream_t stream) {
  UNCL_REQUIRE(in_img_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0, "%s: input must be 16-byte aligned", what);
  int smem_bytes = 0;
  if (int rc = plan_tc(p, N, C_in, bias_floats, what, &smem_bytes)) return rc;

  // 4-D map over 8-byte elements: (x*2 + half, y, channel block, image); one pixel's 8 bf16 channels = 2 elements,
  // so a box row is PW*16 contiguous bytes in global memory (full 32-byte sectors) and lands pixel-major in smem.
  CUtensorMap tmap;
  CUresult r = encode_blocked_bf16(&tmap, in, W, H, C_in / 8, N, in_img_stride, p.PW, p.PH, 2 * p.ksteps);
  if (r != CUDA_SUCCESS) return uncl_set_error(UNCL_ECUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);

  if (p.ns_per_group <= 0) p.ns_per_group = p.NS;
  p.dbg = g_dbg;
  p.probe_noload = probe_env("UNCL_PROBE_NOLOAD") != nullptr ? 1 : 0;
  if (const char* e = probe_env("UNCL_PROBE_FLAGS")) p.probe_noload = atoi(e);
  auto kern = epi == 0 ? conv3x3_tc_kernel<0> : (epi == 1 ? conv3x3_tc_kernel<1> : conv3x3_tc_kernel<2>);
  static thread_local int smem_ok[3] = {0, 0, 0}, smem_dev[3] = {-1, -1, -1};
  cudaError_t e = ensure_smem(kern, smem_bytes, smem_ok[epi], smem_dev[epi]);
  if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "%s: smem attr: %s", what, cudaGetErrorString(e));
  const int sms = sm_count();
  const int grid = p.num_items < sms ? p.num_items : sms;
  kern<<<grid, kThreads, smem_bytes, stream>>>(tmap, p);
  return uncl_check_launch(what);
}

}  // namespace

int uncl_launch_conv3x3_tc_merged(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                                  long out_img_stride, int out_f32, int N, int C_in, int H, int W, int C_out, int pad,
                                  int act, int emit_skip, int fuse_outc, const float* outc_w, const float* outc_b,
                                  float* out_img, float* out_logit, const void* mask, long mask_img_stride,
                                  unsigned long long* dbg, int derive, int x0, cudaStream_t stream);

// Which 3x3 layers run in the kx-merged kernel (conv_tc_merged.cu); packing.conv3x3_tc packs the weights to match.
// Measured per layer of the 1080p frame (profiles/README.md): with C_out <= 64 the merged formulation wins when the
// K loop is long enough to hide its heavier epilogue (C_in >= 64: 500 -> 327 us, 253 -> 155 us, 184 -> 113 us) and loses
// on the C_in = 32 layers, which are epilogue-bound either way.
static bool use_merged(int C_in, int C_out) {
  const char* e = probe_env("UNCL_MERGED_MIN_CI");
  return C_out <= 64 && C_in % 32 == 0 && C_in >= (e ? atoi(e) : 64);
}

static int conv3x3_tc_impl(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                           long out_img_stride, int out_dtype, int N, int C_in, int H, int W, int C_out, int pad,
                           int act, int emit_skip, int fuse_outc, const float* outc_w, const float* outc_b,
                           float* out_img, float* out_logit, const void* mask, long mask_img_stride, int x0,
                           cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C_in % 16 == 0 && C_out % 32 == 0 && (pad == 0 || pad == 2),
               "conv3x3_tc: unsupported C_in=%d C_out=%d pad=%d", C_in, C_out, pad);
  UNCL_REQUIRE(x0 >= 0 && x0 < W + 2 * pad - 2, "conv3x3_tc: first column %d outside the output", x0);
  if (use_merged(C_in, C_out)) {
    UNCL_REQUIRE(out_dtype == UNCL_F32 || out_dtype == UNCL_BF16, "conv3x3_tc: bad out_dtype");
    UNCL_REQUIRE(!fuse_outc || (outc_w && outc_b && out_img), "conv3x3_tc: fuse_outc needs outc params");
    UNCL_REQUIRE(out != nullptr || fuse_outc, "conv3x3_tc: no output requested");
    UNCL_REQUIRE(act == UNCL_ACT_RELU || act == UNCL_ACT_NONE, "conv3x3_tc: only ReLU / identity epilogues are built");
    UNCL_REQUIRE(H + 2 * pad - 2 > 0 && W + 2 * pad - 2 > 0, "conv3x3_tc: empty output");
    return uncl_launch_conv3x3_tc_merged(in, in_img_stride, w_packed, bias, out, out_img_stride, out_dtype == UNCL_F32, N, C_in,
                                         H, W, C_out, pad, act, emit_skip, fuse_outc, outc_w, outc_b, out_img, out_logit,
                                         mask, mask_img_stride, g_dbg, 0, x0, stream);
  }
  TcParams p{};
  p.x0 = x0;
  p.NT = C_out < 128 ? C_out : 128;
  if (C_out == 256 && probe_env("UNCL_PROBE_NT256") != nullptr) p.NT = 256;   // timing probe (weights packed by the caller)
  UNCL_REQUIRE(C_out % p.NT == 0 && C_out <= 1024, "conv3x3_tc: unsupported C_out=%d", C_out);
  UNCL_REQUIRE(out_dtype == UNCL_F32 || out_dtype == UNCL_BF16, "conv3x3_tc: bad out_dtype");
  UNCL_REQUIRE(!fuse_outc || (C_out == 32 && outc_w && outc_b && out_img), "conv3x3_tc: fuse_outc needs C_out == 32 and outc params");
  UNCL_REQUIRE(out != nullptr || fuse_outc, "conv3x3_tc: no output requested");
  UNCL_REQUIRE(act == UNCL_ACT_RELU || act == UNCL_ACT_NONE, "conv3x3_tc: only ReLU / identity epilogues are built");
  p.NS = C_out / p.NT;
  p.w = reinterpret_cast<const bf16*>(w_packed);
  p.bias = bias;
  p.out = out;
  p.out_f32 = out_dtype == UNCL_F32;
  p.out_img_stride = out_img_stride;
  p.outc_w = outc_w; p.outc_b = outc_b; p.out_img = out_img; p.out_logit = out_logit;
  p.N = N; p.C_in = C_in; p.C_out = C_out; p.pad = pad;
  p.Ho = H + 2 * pad - 2; p.Wo = W + 2 * pad - 2;
  UNCL_REQUIRE(p.Ho > 0 && p.Wo > 0, "conv3x3_tc: empty output");
  p.act = act; p.emit_skip = emit_skip; p.fuse_outc = fuse_outc;
  p.mask = reinterpret_cast<const bf16*>(mask); p.mask_img_stride = mask_img_stride;
  p.ntaps = 9;
  return launch_tc(p, in, in_img_stride, N, C_in, H, W, 0, 2 * C_out, "conv3x3_tc", stream);
}

extern "C" int uncl_conv3x3_tc(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                               long out_img_stride, int out_dtype, int N, int C_in, int H, int W, int C_out, int pad,
                               int act, int emit_skip, int fuse_outc, const float* outc_w, const float* outc_b,
                               float* out_img, float* out_logit, cudaStream_t stream) {
  return conv3x3_tc_impl(in, in_img_stride, w_packed, bias, out, out_img_stride, out_dtype, N, C_in, H, W, C_out, pad, act,
                         emit_skip, fuse_outc, outc_w, outc_b, out_img, out_logit, nullptr, 0, 0, stream);
}

// uncl_conv3x3_tc restricted to the output columns [x0, Wo): the trailing columns of a layer whose whole 126-column bands
// run in the row kernel (conv_tc_rows.cu).  Not part of the C ABI.
int uncl_conv3x3_tc_cols(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                         long out_img_stride, int out_dtype, int N, int C_in, int H, int W, int C_out, int pad, int act,
                         int emit_skip, int fuse_outc, const float* outc_w, const float* outc_b, float* out_img,
                         float* out_logit, int x0, cudaStream_t stream) {
  return conv3x3_tc_impl(in, in_img_stride, w_packed, bias, out, out_img_stride, out_dtype, N, C_in, H, W, C_out, pad, act,
                         emit_skip, fuse_outc, outc_w, outc_b, out_img, out_logit, nullptr, 0, x0, stream);
}

// uncl_conv3x3_tc_skipcat restricted to the output columns [x0, Wo) (arguments validated by the callers).
int uncl_conv3x3_tc_skipcat_cols(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                                 long out_img_stride, int out_dtype, int N, int C_skip, int H, int W, int C_out, int pad,
                                 int act, int x0, cudaStream_t stream) {
  UNCL_REQUIRE(x0 >= 0 && x0 < W + 2 * pad - 2, "conv3x3_tc_skipcat: first column %d outside the output", x0);
  return uncl_launch_conv3x3_tc_merged(in, in_img_stride, w_packed, bias, out, out_img_stride, out_dtype == UNCL_F32, N,
                                       4 * C_skip, H, W, C_out, pad, act, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                                       g_dbg, 1, x0, stream);
}

// First conv of an `up` block with the skip operators FUSED (unet_parts.py:311-332: x2 -> [x2, x2^2, sqrt(x2 + 1e-8)],
// cat with the up-sampled x1, ConvTranspose 3x3).  `in` holds only [skip (C_skip) | up-sampled (C_skip)] channels; the
// squared and square-root planes are never written to memory: the kernel builds them in shared memory from the skip
// chunk it has just loaded (conv_tc_merged.cu, derive mode), which halves the layer's DRAM reads and lets the producing
// layer write one plane instead of three.  w_packed: packing.conv3x3_tc of the full [9][4*C_skip][C_out] filter bank
// (concat order skip | up | skip^2 | sqrt).  Built for C_out == 32 (the 252^2 and 124^2 levels of the generator).
extern "C" int uncl_conv3x3_tc_skipcat(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                                       long out_img_stride, int out_dtype, int N, int C_skip, int H, int W, int C_out, int pad,
                                       int act, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C_skip > 0 && C_skip % 32 == 0 && C_out == 32 && (pad == 0 || pad == 2) && out != nullptr,
               "conv3x3_tc_skipcat: unsupported C_skip=%d C_out=%d pad=%d", C_skip, C_out, pad);
  UNCL_REQUIRE(out_dtype == UNCL_F32 || out_dtype == UNCL_BF16, "conv3x3_tc_skipcat: bad out_dtype");
  UNCL_REQUIRE(act == UNCL_ACT_RELU || act == UNCL_ACT_NONE, "conv3x3_tc_skipcat: only ReLU / identity epilogues are built");
  UNCL_REQUIRE(H + 2 * pad - 2 > 0 && W + 2 * pad - 2 > 0, "conv3x3_tc_skipcat: empty output");
  return uncl_conv3x3_tc_skipcat_cols(in, in_img_stride, w_packed, bias, out, out_img_stride, out_dtype, N, C_skip, H, W, C_out,
                                      pad, act, 0, stream);
}

// Data gradient of a 3x3 conv / ConvTranspose 3x3 with the ReLU backward of the PRODUCING layer fused into the epilogue:
// dX = corr(dZ, flipped / transposed taps; pad 2 - p), then dX *= (mask > 0) where `mask` is that layer's (post-ReLU)
// output = this conv's forward input.  The result is the pre-activation gradient dZ of the previous layer, written as
// bf16 (the operand of its own data / weight gradient GEMMs) or fp32.  mask == NULL: plain data gradient.
// in: dZ bf16 blocked [N][C_in/8][H][W][8]; w_packed: packing.conv3x3_tc of the transposed taps; mask: bf16 blocked with
// C_out channels and the output extent, image stride mask_img_stride elements (it may live in a concat buffer).
extern "C" int uncl_conv3x3_tc_dgrad(const void* in, long in_img_stride, const void* w_packed, const void* mask,
                                     long mask_img_stride, void* out, long out_img_stride, int out_dtype, int N, int C_in,
                                     int H, int W, int C_out, int pad, cudaStream_t stream) {
  UNCL_REQUIRE(out != nullptr && (mask == nullptr || ((reinterpret_cast<uintptr_t>(mask) & 15) == 0 && mask_img_stride % 8 == 0)),
               "conv3x3_tc_dgrad: bad output / mask");
  return conv3x3_tc_impl(in, in_img_stride, w_packed, nullptr, out, out_img_stride, out_dtype, N, C_in, H, W, C_out, pad,
                         UNCL_ACT_NONE, 0, 0, nullptr, nullptr, nullptr, nullptr, mask, mask_img_stride, 0, stream);
}

// ConvTranspose2d(C, C, 2, stride=2) as a GEMM [pixels x C] . [C x 4C] with a pixel-shuffle epilogue.
// w_packed: bf16 [NS][C/16][1][2][NT][8], column j = (dy*2+dx)*C + co, NT = min(4C, 128).
static int convT2x2_tc_impl(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                            long out_img_stride, int out_dtype, int N, int C_in, int C, int H, int W, int H2, int W2,
                            cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 32 == 0 && C_in % 16 == 0 && H2 >= 2 * H && W2 >= 2 * W && W <= 128,
               "convT2x2_tc: unsupported C_in=%d C=%d W=%d", C_in, C, W);
  UNCL_REQUIRE(out_dtype == UNCL_F32 || out_dtype == UNCL_BF16, "convT2x2_tc: bad out_dtype");
  TcParams p{};
  const int n_total = 4 * C;
  p.NT = n_total < 128 ? n_total : 128;
  p.NS = n_total / p.NT;
  p.w = reinterpret_cast<const bf16*>(w_packed);
  p.bias = bias;
  p.out = out;
  p.out_f32 = out_dtype == UNCL_F32;
  p.out_img_stride = out_img_stride;
  p.N = N; p.C_in = C_in; p.C_out = C; p.pad = 0;
  p.Ho = H; p.Wo = W;
  p.act = UNCL_ACT_NONE;
  p.ntaps = 1;
  p.sh_C = C; p.sh_H2 = H2; p.sh_W2 = W2;
  return launch_tc(p, in, in_img_stride, N, C_in, H, W, 1, 2 * C, "convT2x2_tc", stream);
}

extern "C" int uncl_convT2x2_tc(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                                long out_img_stride, int out_dtype, int N, int C, int H, int W, int H2, int W2,
                                cudaStream_t stream) {
  return convT2x2_tc_impl(in, in_img_stride, w_packed, bias, out, out_img_stride, out_dtype, N, C, C, H, W, H2, W2, stream);
}

// The same GEMM with C_in input channels != C output channels: the exact path feeds the three-term bf16 split
// [x_hi | x_hi | x_lo] (C_in = 3C) against [w_hi ; w_lo ; w_hi] (packing.convT2x2_tc_split).
extern "C" int uncl_convT2x2_tc_cin(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                                    long out_img_stride, int out_dtype, int N, int C_in, int C, int H, int W, int H2, int W2,
                                    cudaStream_t stream) {
  return convT2x2_tc_impl(in, in_img_stride, w_packed, bias, out, out_img_stride, out_dtype, N, C_in, C, H, W, H2, W2, stream);
}

// 1x1 (optionally grouped) convolution as a tensor-core GEMM [pixels x C_in/g] . [C_in/g x C_out/g] per group:
// out = scale[n] * act(W x + b) + res.  gcn_lib/torch_nn.py:54-78 (BasicConv), torch_vertex.py:219-227, Unet_singleFrame.py:36-42;
// also the data gradient of the k2 s2 up-convolution (a 4C -> C pointwise GEMM over the space-to-depth gradient).
// in: bf16 blocked [N][C_in/8][H][W][8], W <= 128.  w_packed: bf16 [NS][C_in/g/16][2][NT][8], NT = min(C_out/g, 128).
static int pw_conv_tc_impl(const void* in, long in_img_stride, const void* w_packed, const float* bias, const void* res,
                           long res_img_stride, int res_dtype, const float* scale, void* out, long out_img_stride,
                           int out_dtype, int N, int C_in, int C_out, int groups, int H, int W, int act, const void* mask,
                           long mask_img_stride, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && groups > 0 && C_in % (16 * groups) == 0 && C_out % (32 * groups) == 0 && W <= 128 && H > 0,
               "pw_conv_tc: unsupported C_in=%d C_out=%d groups=%d W=%d", C_in, C_out, groups, W);
  UNCL_REQUIRE((out_dtype == UNCL_F32 || out_dtype == UNCL_BF16) && out != nullptr, "pw_conv_tc: bad output");
  UNCL_REQUIRE(res == nullptr || res_dtype == UNCL_F32 || res_dtype == UNCL_BF16, "pw_conv_tc: bad res_dtype");
  UNCL_REQUIRE(act == UNCL_ACT_NONE || act == UNCL_ACT_RELU || act == UNCL_ACT_GELU, "pw_conv_tc: unsupported activation");
  TcParams p{};
  const int cout_g = C_out / groups;
  p.NT = cout_g < 128 ? cout_g : 128;
  UNCL_REQUIRE(cout_g % p.NT == 0 && C_out <= 2048, "pw_conv_tc: unsupported C_out/groups=%d", cout_g);
  p.NS = C_out / p.NT;
  p.ns_per_group = cout_g / p.NT;
  p.nchunk = C_in / groups / 16;
  p.w = reinterpret_cast<const bf16*>(w_packed);
  p.bias = bias;
  p.out = out;
  p.out_f32 = out_dtype == UNCL_F32;
  p.out_img_stride = out_img_stride;
  p.res = res; p.res_img_stride = res_img_stride; p.res_f32 = res_dtype == UNCL_F32; p.scale = scale;
  p.N = N; p.C_in = C_in; p.C_out = C_out; p.pad = 0;
  p.Ho = H; p.Wo = W;
  p.act = act;
  p.mask = reinterpret_cast<const bf16*>(mask); p.mask_img_stride = mask_img_stride;
  p.ntaps = 1;
  return launch_tc(p, in, in_img_stride, N, C_in, H, W, 2, 2 * C_out, "pw_conv_tc", stream);
}

extern "C" int uncl_pw_conv_tc(const void* in, long in_img_stride, const void* w_packed, const float* bias, const void* res,
                               long res_img_stride, int res_dtype, const float* scale, void* out, long out_img_stride,
                               int out_dtype, int N, int C_in, int C_out, int groups, int H, int W, int act,
                               cudaStream_t stream) {
  return pw_conv_tc_impl(in, in_img_stride, w_packed, bias, res, res_img_stride, res_dtype, scale, out, out_img_stride,
                         out_dtype, N, C_in, C_out, groups, H, W, act, nullptr, 0, stream);
}

// Pointwise data gradient with the ReLU backward of the producing layer fused (see uncl_conv3x3_tc_dgrad): the k2 s2
// up-convolution's dX = [space-to-depth dY (4C)] . W^T, masked by the decoder activation it up-sampled.
extern "C" int uncl_pw_conv_tc_dgrad(const void* in, long in_img_stride, const void* w_packed, const void* mask,
                                     long mask_img_stride, void* out, long out_img_stride, int out_dtype, int N, int C_in,
                                     int C_out, int groups, int H, int W, cudaStream_t stream) {
  return pw_conv_tc_impl(in, in_img_stride, w_packed, nullptr, nullptr, 0, UNCL_BF16, nullptr, out, out_img_stride, out_dtype,
                         N, C_in, C_out, groups, H, W, UNCL_ACT_NONE, mask, mask_img_stride, stream);
}

int uncl_plan_conv3x3_tc_merged(int N, int C_in, int H, int W, int C_out, int pad, int* plan);

// Tile plan of uncl_conv3x3_tc for a problem - pure host arithmetic, no GPU needed.  plan[16]: kind (0 one-tap kernel,
// bit 0 kx-merged kernel, bit 1 its tiles are row-aligned, bit 2 its weights stay resident in shared memory), NT, NS, MMA N, M blocks per tile, tile advance in positions, PW, PH, BW, bands, tiles per band,
// work items, pipeline stages, accumulator stages, K=16 steps per stage, dynamic shared memory bytes.
extern "C" int uncl_conv3x3_tc_plan(int N, int C_in, int H, int W, int C_out, int pad, int* plan) {
  UNCL_REQUIRE(plan != nullptr && N > 0 && C_in % 16 == 0 && C_out % 32 == 0 && (pad == 0 || pad == 2) &&
                   H + 2 * pad - 2 > 0 && W + 2 * pad - 2 > 0,
               "conv3x3_tc_plan: unsupported C_in=%d C_out=%d pad=%d H=%d W=%d", C_in, C_out, pad, H, W);
  if (use_merged(C_in, C_out)) return uncl_plan_conv3x3_tc_merged(N, C_in, H, W, C_out, pad, plan);
  TcParams p{};
  p.NT = C_out < 128 ? C_out : 128;
  UNCL_REQUIRE(C_out % p.NT == 0 && C_out <= 1024, "conv3x3_tc_plan: unsupported C_out=%d", C_out);
  p.NS = C_out / p.NT;
  p.Ho = H + 2 * pad - 2; p.Wo = W + 2 * pad - 2;
  p.ntaps = 9;
  int smem = 0;
  if (int rc = plan_tc(p, N, C_in, 2 * C_out, "conv3x3_tc_plan", &smem)) return rc;
  const int v[16] = {0, p.NT, p.NS, p.NT, p.MB, 128 * p.MB, p.PW, p.PH, p.BW, p.nbands, p.tiles_per_band, p.num_items,
                     p.stages, p.nacc, p.ksteps, smem};
  for (int i = 0; i < 16; ++i) plan[i] = v[i];
  return UNCL_OK;
}

// Diagnostics: when set (device pointer to 8 zeroed uint64), every tensor-core conv launch adds, summed over its CTAs,
// [0] producer cycles, [1] producer wait-for-empty-stage, [2] MMA-issuer cycles, [3] its wait-for-full-stage,
// [4] its wait-for-free-accumulator, [5] epilogue-warp cycles, [6] its wait-for-accumulator, [7] CTA count.
// Exists only in the -DUNCL_PROBES build (tools/): the product library has no mutable global and returns
// UNCL_EUNSUPPORTED here.  Not thread-safe; pass NULL to switch off.
extern "C" int uncl_conv_tc_set_debug(void* counters) {
#ifdef UNCL_PROBES
  g_dbg = reinterpret_cast<unsigned long long*>(counters);
  return UNCL_OK;
#else
  (void)counters;
  return uncl_set_error(UNCL_EUNSUPPORTED, "conv_tc_set_debug: this library was built without -DUNCL_PROBES "
                                           "(python -m uncltmo_b200.build --probes)");
#endif
}
