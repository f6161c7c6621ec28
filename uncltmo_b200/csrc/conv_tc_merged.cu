// 3x3 stride-1 convolution / ConvTranspose 3x3 for NARROW outputs (C_out <= 64) on the tensor cores with the three
// kx taps of a kernel row merged into the N dimension of one tcgen05.mma.
//
// Why (profiles/r1_mma_probe.csv): a kind::f16 M=128 K=16 instruction with both operands in shared memory costs
// max(N/2, (4096 + 32 N) / 128) cycles - it reads its 4 KB A tile (128 pixels x 16 channels) and its B tile at
// 128 B/cycle every time, whatever the previous instruction used.  With N = C_out = 32 (52 % of the generator's FLOPs)
// the A read alone is twice the math.  Merging the three kx taps makes N' = 3 C_out columns share one A read:
//     D[p][kx*C + co] = sum_{ky, ci} X[p + ky*PW][ci] * W[ky][kx][ci][co]        (3 MMAs per K chunk instead of 9)
//     out[q][co]      = D[q][co] + D[q+1][C + co] + D[q+2][2C + co]              (epilogue)
// N' = 96 costs 56 cycles for what took 3 x 40, N' = 192 runs at the math floor (96 cycles).
// Used for C_out <= 64 with C_in >= 64 (uncl_conv3x3_tc decides): with only 32 input channels the K loop is too short
// to hide this kernel's heavier epilogue (~7500 warp instructions per 256-position tile, ncu) and conv_tc.cu stays ahead.
// The price is the epilogue's cross-row sum: accumulator rows are TMEM lanes, a warp only reaches its own 32 lanes, so
// the kx = 1, 2 column groups come from the neighbouring lanes by warp shuffle and, for the last two lanes of a warp,
// from the next warp through a small shared-memory exchange.  Tiles overlap by two positions so that no exchange
// crosses a tile.
//
// Everything else follows conv_tc.cu: C8-blocked bf16 activations, one TMA halo box per K chunk whose shared-memory
// image is the no-swizzle K-major operand, a filter ROW (ky) is a descriptor start-address shift of ky*PW pixels,
// persistent warp-specialised CTAs (1 TMA producer warp, 1 MMA-issuing warp, 16 epilogue warps), double-buffered TMEM
// accumulator (2 x 256 columns: two M blocks of N' = 96 or one of N' = 192 per stage).
// Reference operator: models/unet_multi_filters/unet_parts.py:57-87, :126-141, :183-193, :319-322, :338-345.
// Weights: bf16 [NS][C_in/16][3 ky][2][3*NT (kx, n)][8], NT = min(C_out, 64) (packing.conv3x3_tc).
#include <cstdlib>
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int kThreads = 608;     // warp 0: TMA producer, 1: MMA issuer, 2-17: epilogue, 18: second MMA issuer
constexpr int kThreadsDerive = 736;   // ... + warps 19-22: skip-operator warps (the kDerive instantiation only)
constexpr int kMma2Warp = 18;     // issues the second M block of a tile when MgParams::mma_warps == 2
constexpr int kDeriveWarp0 = 19;  // fused skip operators (MgParams::derive): x^2 and sqrt(x + eps) chunks built in shared memory
constexpr int kDeriveWarps = 4;
constexpr int kMaxProg = 16;      // ring positions (K chunks) per tile in derive mode
constexpr int kEpiWarps = 16;     // 4 TMEM lane quarters x 2 work units x 2 halves of 16 channels
constexpr int kMaxStages = 8;
constexpr int kUnits = 2;         // (M block, 32-channel chunk) pairs per tile: 2 blocks x 32 ch or 1 block x 64 ch
constexpr int kXSlot = 96;        // floats per exchange slot: lane 0's kx=1 and kx=2 groups, lane 1's kx=2 group

struct MgParams {
  const bf16* w;
  const float* bias;
  void* out;
  int out_f32;
  long out_img_stride;
  const float* outc_w;
  const float* outc_b;
  float* out_img;
  float* out_logit;
  const bf16* mask;         // data-gradient mode (uncl_conv3x3_tc_dgrad): out = (mask > 0) ? acc : 0
  long mask_img_stride;
  int C_out, Ho, Wo, pad;
  int x0;                   // first output column of this launch: bands cover [x0, Wo) (trailing columns of conv_tc_rows.cu)
  int NT, NS, NP;           // channels per N split, N splits, MMA N = 3 * NT
  int MB, ADV;              // M blocks per tile, tile advance in positions (128 * MB - 2)
  int PW, PH, BW;
  int band_total, tiles_per_band, tiles_per_img, num_items;
  int nchunk, ksteps, stages, nacc, acc_cols;   // ksteps: K = 16 steps (16 input channels each) per pipeline stage
  int mma_warps;            // 1, or 2: the M blocks of a tile are issued by two warps (each commits its own arrivals)
  int w_res, w_total;       // weights resident in shared memory (loaded once per CTA) and their size in bytes
  int aligned;              // row-aligned tiles (ADV = R * PW) instead of flat tiles (ADV = 128 * MB - 2)
  // Fused skip operators (unet_parts.py:319-322).  The input tensor holds [skip (Cs) | up-sampled (Cs)] only; the K loop
  // also visits skip^2 and sqrt(skip + 1e-8), which four extra warps build from the skip chunk's shared-memory image.
  // A tile walks `nchunk` ring positions; position k is a TMA chunk (kind 0, channel block prog_cb[k]) or a derived chunk
  // (kind 1: square, 2: square root) of the TMA chunk prog_back[k] positions earlier; prog_wch[k] = its weight chunk.
  int derive, H_in, W_in;
  unsigned char prog_kind[kMaxProg], prog_cb[kMaxProg], prog_wch[kMaxProg], prog_back[kMaxProg];
  int a_box_bytes, a_stage_bytes, b_stage_bytes, stage_bytes;
  int act, emit_skip, fuse_outc;
  int probe_noload;         // timing probe: the producer only loads the first `stages` chunks, then re-signals stale stages
  unsigned long long m_NS, m_tpi, m_tpb, m_PW;   // 2^40 / d + 1: exact x / d for x < 2^20, d < 2^12 ... (see fastdiv)
  unsigned long long* dbg;
};

// x / d with the host-computed magic m = 2^40 / d + 1 (exact for x < 2^24 and d <= 2^16: x * (m - 2^40/d) < 2^40 / d)
__device__ __forceinline__ int fastdiv(int x, unsigned long long m) {
  return (int)(((unsigned long long)(uint32_t)x * m) >> 40);
}

struct MgItem {
  int n, ns, band, bx, by, moff0, q0;
};

__device__ __forceinline__ MgItem mg_decode(const MgParams& p, int item) {
  MgItem it;
  const int tile = fastdiv(item, p.m_NS);
  it.ns = item - tile * p.NS;
  it.n = fastdiv(tile, p.m_tpi);
  const int t = tile - it.n * p.tiles_per_img;
  it.band = fastdiv(t, p.m_tpb);
  const int tb = t - it.band * p.tiles_per_band;
  it.q0 = tb * p.ADV;
  const int y0 = fastdiv(it.q0, p.m_PW);
  it.bx = p.x0 + it.band * p.BW - p.pad;
  it.by = y0 - p.pad;
  it.moff0 = it.q0 - y0 * p.PW;
  return it;
}

__device__ __forceinline__ void tc_ld16_nw(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// two bf16 in one register: elementwise product (one rounding of the exact product, like bf16(x * x) of a float x)
__device__ __forceinline__ uint32_t bf16x2_mul(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// two bf16 -> bf16(sqrt(x + 1e-8)) each (unet_parts.py:321), MUFU square root as in the producing epilogue
__device__ __forceinline__ uint32_t bf16x2_sqrt_eps(uint32_t v) {
  const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
  return pack_bf16x2(fast_sqrt(lo + 1e-8f), fast_sqrt(hi + 1e-8f));
}

// MMA issuer: 3 filter rows x MB blocks per K chunk, straight-line (MB is a compile-time constant), every block of
// the tile is always issued (blocks past the end of a band read zero-filled / stale rows and are masked later).
// B0..B1: the M blocks this warp issues (with two issuing warps each block has its own accumulator columns, so the two
// instruction streams never touch the same TMEM columns and need no ordering between them).
template <int MB, int kKSteps, int B0, int B1>
__device__ __forceinline__ void mg_mma_role(const MgParams& p, uint8_t* stage_base, uint64_t* full, uint64_t* empty,
                                            uint64_t* tfull, uint64_t* tempty, uint64_t* wfull, uint32_t tmem_base, int lane) {
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NP >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t desc_hi = (128u >> 4) | (1u << 14);
  const uint32_t a_lo_const = ((uint32_t)(p.PH * p.PW) & 0x3fffu) << 16;   // LBO_A = PH*PW*16 B
  const uint32_t b_lo_const = ((uint32_t)p.NP & 0x3fffu) << 16;            // LBO_B = NP*16 B
  const uint32_t stage0_16 = smem_u32(stage_base) >> 4, stage_16 = (uint32_t)p.stage_bytes >> 4;
  const uint32_t a_bytes_16 = (uint32_t)p.a_stage_bytes >> 4, b_row_16 = (uint32_t)(2 * p.NP);
  // resident weights live behind the stage ring: chunk ch at wres + ch * b_stage_bytes
  const bool w_res = p.w_res != 0;
  const uint32_t wres_16 = stage0_16 + (uint32_t)p.stages * stage_16, b_stage_16 = (uint32_t)p.b_stage_bytes >> 4;
  const uint32_t pw = (uint32_t)p.PW, np = (uint32_t)p.NP;
  const uint32_t a_kstep_16 = (uint32_t)(2 * p.PH * p.PW);   // the next two channel blocks of the halo box
  const int nacc = p.nacc, acc_cols = p.acc_cols, nchunk = p.nchunk, stages = p.stages, num_items = p.num_items;
  int stage = 0, acc = 0;
  uint32_t phase = 0, acc_phase = 0;
  unsigned long long* const dbg = (UNCL_PROBE(1, 1) && B0 == 0) ? p.dbg : nullptr;
  long long w_full = 0, w_tempty = 0;
  const long long t_begin = dbg ? clock64() : 0;
  const unsigned long long g_begin = dbg ? globaltimer_ns() : 0ull;
  if (w_res) mbar_wait(wfull, 0);
  for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
    const MgItem it = mg_decode(p, item);
    const long long tw0 = dbg ? clock64() : 0;
    mbar_wait(&tempty[acc], acc_phase ^ 1);
    if (dbg) w_tempty += clock64() - tw0;
    tc_fence_after();
    const uint32_t d0 = tmem_base + (uint32_t)(acc * acc_cols);
    for (int ch = 0; ch < nchunk; ++ch) {
      const long long tw1 = dbg ? clock64() : 0;
      if (!(UNCL_PROBE(p.probe_noload, 4))) mbar_wait(&full[stage], phase);
      if (dbg) w_full += clock64() - tw1;
      tc_fence_after();
      const uint32_t sa16 = stage0_16 + (uint32_t)stage * stage_16;
      const uint32_t a_row = a_lo_const | (sa16 + (uint32_t)it.moff0);
      const uint32_t wch = p.derive ? (uint32_t)p.prog_wch[ch] : (uint32_t)ch;
      const uint32_t b_lo = b_lo_const | (w_res ? wres_16 + wch * b_stage_16 : sa16 + a_bytes_16);
      if (elect_one()) {
        // a stage holds kKSteps K=16 steps (32 input channels): kKSteps x 3 filter rows x MB blocks, straight-line
#pragma unroll
        for (int kc = 0; kc < kKSteps; ++kc) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int b = B0; b < B1; ++b) {
              tc_mma_bf16(d0 + (uint32_t)b * np, a_row + (uint32_t)kc * a_kstep_16 + (uint32_t)ky * pw + (uint32_t)b * 128u,
                          desc_hi, b_lo + (uint32_t)(kc * 3 + ky) * b_row_16, desc_hi, idesc,
                          (kc > 0 || ky > 0) ? 1u : (ch > 0 ? 1u : 0u));
            }
          }
        }
        if (!(UNCL_PROBE(p.probe_noload, 4))) tc_commit(&empty[stage]);
      }
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

template <bool kDerive>
__global__ void __launch_bounds__(kDerive ? kThreadsDerive : kThreads, 1)
conv3x3_tc_merged_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ MgParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* stage_base = smem;
  float* xbuf = reinterpret_cast<float*>(smem + (size_t)p.stages * p.stage_bytes + (p.w_res ? p.w_total : 0));   // [2][kUnits][4][kXSlot]
  float* lbuf = xbuf + 2 * kUnits * 4 * kXSlot;                                       // [2][kUnits][128] partial logits
  uint64_t* bars = reinterpret_cast<uint64_t*>(lbuf + 2 * kUnits * 128);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [C_out] (+ [C_out] outc weights)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.num_items, nchunk = p.nchunk, stages = p.stages, stage_bytes = p.stage_bytes;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    // derive mode: the skip-operator warps arrive on every full (they fill the derived stages) and on every empty (they
    // read the skip stages)
    const uint32_t extra = p.derive ? (uint32_t)kDeriveWarps : 0u;
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1 + extra); mbar_init(&empty[s], (uint32_t)p.mma_warps + extra); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], (uint32_t)p.mma_warps); mbar_init(&tempty[s], kEpiWarps); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < p.C_out; i += (int)blockDim.x) {
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
      const bool w_res = p.w_res != 0;
      const uint32_t b_bytes = (uint32_t)p.b_stage_bytes, tx_bytes = (uint32_t)p.a_box_bytes + (w_res ? 0u : b_bytes);
      const int a_stage_bytes = p.a_stage_bytes;
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w);
      if (w_res) {   // the whole filter bank once per CTA (NS == 1), chunk by chunk behind the stage ring
        uint8_t* wres = stage_base + (size_t)stages * stage_bytes;
        mbar_expect_tx(wfull, (uint32_t)p.w_total);
        for (int ch = 0; ch < nchunk; ++ch) bulk_load(wres + (size_t)ch * b_bytes, wbase + (size_t)ch * b_bytes, b_bytes, wfull);
      }
      unsigned long long* const dbg = UNCL_PROBE(1, 1) ? p.dbg : nullptr;
      long long w_empty = 0;
      const long long t_begin = dbg ? clock64() : 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const MgItem it = mg_decode(p, item);
        const uint8_t* wsrc = wbase + (size_t)it.ns * nchunk * b_bytes;
        for (int ch = 0; ch < nchunk; ++ch) {
          if (UNCL_PROBE(p.probe_noload, 4)) continue;
          const long long tw0 = dbg ? clock64() : 0;
          mbar_wait(&empty[stage], phase ^ 1);
          if (dbg) w_empty += clock64() - tw0;
          uint8_t* sa = stage_base + (size_t)stage * stage_bytes;
          if ((UNCL_PROBE(p.probe_noload, 1)) && (item != (int)blockIdx.x || ch >= stages)) { mbar_arrive(&full[stage]); }
          else if (p.derive) {
            const int wch = p.prog_wch[ch];
            if (p.prog_kind[ch] == 0) {
              mbar_expect_tx(&full[stage], tx_bytes);
              tma_load_4d(sa, &tmap, &full[stage], it.bx * 2, it.by, (int)p.prog_cb[ch], it.n);
              if (!w_res) bulk_load(sa + a_stage_bytes, wsrc + (size_t)wch * b_bytes, b_bytes, &full[stage]);
            } else if (!w_res) {   // derived chunk: only its weights come from memory
              mbar_expect_tx(&full[stage], b_bytes);
              bulk_load(sa + a_stage_bytes, wsrc + (size_t)wch * b_bytes, b_bytes, &full[stage]);
            } else {
              mbar_arrive(&full[stage]);
            }
          }
          else {
          mbar_expect_tx(&full[stage], tx_bytes);
          tma_load_4d(sa, &tmap, &full[stage], it.bx * 2, it.by, ch * 2 * p.ksteps, it.n);
          if (!w_res) bulk_load(sa + a_stage_bytes, wsrc + (size_t)ch * b_bytes, b_bytes, &full[stage]);
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
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // N' = 96: two M blocks, 32 input channels per stage (halves the barrier round trips of the latency-bound issuing
    // warp: 329 -> 302 us on up3.conv); N' = 192: one M block, 16 channels per stage (the 18 KB weight stage would
    // leave only 3 pipeline stages otherwise: 101 -> 108 us on up1.conv)
    // The issuing warp is a serial, latency-bound instruction stream (uniform-datapath adds, R2UR, barrier polls): with
    // two M blocks per tile a second warp on another scheduler takes block 1, halving the instructions behind each MMA.
    if (p.MB == 1 && p.ksteps == 1) mg_mma_role<1, 1, 0, 1>(p, stage_base, full, empty, tfull, tempty, wfull, tmem_base, lane);
    else if (p.MB == 1) mg_mma_role<1, 2, 0, 1>(p, stage_base, full, empty, tfull, tempty, wfull, tmem_base, lane);
    else if (p.mma_warps == 2) mg_mma_role<2, 2, 0, 1>(p, stage_base, full, empty, tfull, tempty, wfull, tmem_base, lane);
    else mg_mma_role<2, 2, 0, 2>(p, stage_base, full, empty, tfull, tempty, wfull, tmem_base, lane);
    __syncwarp();
  } else if (warp == kMma2Warp) {
    if (p.mma_warps == 2) mg_mma_role<2, 2, 1, 2>(p, stage_base, full, empty, tfull, tempty, wfull, tmem_base, lane);
    __syncwarp();
  } else if (warp >= kDeriveWarp0) {
    // =============================== fused skip operators ===============================
    // These warps walk the stage ring like a producer.  At a TMA position they only add their arrival; at a derived
    // position they wait for the source chunk (the skip activations, landed by TMA `back` positions earlier), read its
    // shared-memory image [2*ksteps channel blocks][PH*PW pixels][8] and write x*x or sqrt(x + 1e-8) at the same offsets
    // of their own stage - the layout IS the MMA operand layout, so the transform is elementwise.  The zero padding of the
    // transposed conv applies to the concatenated tensor: sqrt is 0 outside the image, not sqrt(eps).
    if constexpr (kDerive) {
      const int dtid = (warp - kDeriveWarp0) * 32 + lane;
      const int PHPW = p.PH * p.PW, PW = p.PW, H_in = p.H_in, W_in = p.W_in;   // derive mode: ksteps == 2, four channel blocks per stage
      int stage = 0;
      uint32_t phase = 0;
      unsigned long long* const dbg = (UNCL_PROBE(1, 1) && warp == kDeriveWarp0 && lane == 0) ? p.dbg : nullptr;
      long long w_e = 0, w_s = 0;
      const long long t_begin = dbg ? clock64() : 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const MgItem it = mg_decode(p, item);
        for (int ch = 0; ch < nchunk; ++ch) {
          const int kind = p.prog_kind[ch];
          if (kind == 2) {   // built together with the square (previous position)
            if (++stage == stages) { stage = 0; phase ^= 1; }
            continue;
          }
          const long long tw0 = dbg ? clock64() : 0;
          mbar_wait(&empty[stage], phase ^ 1);
          if (kind != 0) {
            // ONE pass builds both derived chunks: the square into this stage, the square root into the next one
            int src = stage - 1, nst = stage + 1;
            uint32_t src_phase = phase, n_phase = phase;
            if (src < 0) { src += stages; src_phase ^= 1; }
            if (nst == stages) { nst = 0; n_phase ^= 1; }
            mbar_wait(&empty[nst], n_phase ^ 1);
            if (dbg) w_e += clock64() - tw0;
            const long long tw1 = dbg ? clock64() : 0;
            mbar_wait(&full[src], src_phase);
            if (dbg) w_s += clock64() - tw1;
            const uint8_t* sp = stage_base + (size_t)src * stage_bytes;
            uint8_t* dp = stage_base + (size_t)stage * stage_bytes;
            uint8_t* dq = stage_base + (size_t)nst * stage_bytes;
            constexpr int kStep = kDeriveWarps * 32;
            for (int pos = dtid; pos < PHPW; pos += kStep) {
              if (UNCL_PROBE(p.probe_noload, 16)) continue;   // timing probe: no transform at all (wrong results)
              uint4 v[4], q[4];
#pragma unroll
              for (int cb = 0; cb < 4; ++cb) v[cb] = *reinterpret_cast<const uint4*>(sp + (cb * PHPW + pos) * 16);
              const int r0 = fastdiv(pos, p.m_PW), c0 = pos - r0 * PW;
              const uint32_t in0 = ((unsigned)(it.by + r0) < (unsigned)H_in && (unsigned)(it.bx + c0) < (unsigned)W_in) ? 0xffffffffu : 0u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                q[j].x = bf16x2_sqrt_eps(v[j].x) & in0; q[j].y = bf16x2_sqrt_eps(v[j].y) & in0;
                q[j].z = bf16x2_sqrt_eps(v[j].z) & in0; q[j].w = bf16x2_sqrt_eps(v[j].w) & in0;
                v[j].x = bf16x2_mul(v[j].x, v[j].x); v[j].y = bf16x2_mul(v[j].y, v[j].y);
                v[j].z = bf16x2_mul(v[j].z, v[j].z); v[j].w = bf16x2_mul(v[j].w, v[j].w);
              }
#pragma unroll
              for (int cb = 0; cb < 4; ++cb) {
                *reinterpret_cast<uint4*>(dp + (cb * PHPW + pos) * 16) = v[cb];
                *reinterpret_cast<uint4*>(dq + (cb * PHPW + pos) * 16) = q[cb];
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> the MMA's async-proxy reads
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&full[stage]); mbar_arrive(&empty[stage]);   // the derived stages are never read by these warps
              mbar_arrive(&full[nst]); mbar_arrive(&empty[nst]);
              mbar_arrive(&empty[src]);                                 // done reading the skip chunk
            }
          } else {
            if (dbg) w_e += clock64() - tw0;
            if (lane == 0) {
              mbar_arrive(&full[stage]);
              if (p.prog_back[ch] == 0) mbar_arrive(&empty[stage]);   // a TMA chunk nothing is derived from (prog_back = 1 marks a source)
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg) {   // [9] skip-operator warp cycles, [10] its wait for a free stage, [11] its wait for the skip chunk
        atomicAdd(dbg + 9, (unsigned long long)(clock64() - t_begin));
        atomicAdd(dbg + 10, (unsigned long long)w_e);
        atomicAdd(dbg + 11, (unsigned long long)w_s);
      }
    }
  } else {
    // =============================== epilogue ===============================
    // A tile has two work units (M block b, 32-channel chunk c): 2 x 32 channels (NT = 32) or 1 x 64 channels (NT = 64).
    // Each unit is handled, per TMEM lane quarter, by two warps that take 16 channels each: 4 x 2 x 2 = 16 warps.
    const int quarter = warp & 3;                 // a warp reaches TMEM lanes 32 * (warp % 4) ..
    const int k = (warp - 2) >> 2;                // 0..3
    const int u = k & 1, h = k >> 1;
    const int row = quarter * 32 + lane;
    const int Ho = p.Ho, Wo = p.Wo, NT = p.NT, NP = p.NP, C_out = p.C_out, ADV = p.ADV, PW = p.PW, BW = p.BW;
    const int b = NT == 32 ? u : 0, c = NT == 32 ? 0 : u;
    const uint32_t col = (uint32_t)(b * NP + c * 32 + h * 16);
    const long cb_stride = (long)Ho * Wo * 8;
    const long skip2 = (long)(2 * (C_out / 8)) * cb_stride, skip3 = (long)(3 * (C_out / 8)) * cb_stride;
    const float act_floor = (p.act == UNCL_ACT_RELU) ? 0.f : -INFINITY;
    const bool emit_skip = p.emit_skip != 0, fuse_outc = p.fuse_outc != 0;
    bf16* const out = reinterpret_cast<bf16*>(p.out);
    float* const outf = reinterpret_cast<float*>(p.out);
    const bool out_f32 = p.out_f32 != 0;
    const long out_img_stride = p.out_img_stride;
    float* const out_img = p.out_img;
    float* const out_logit = p.out_logit;
    const float outc_b = fuse_outc ? __ldg(p.outc_b) : 0.f;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int nacc = p.nacc, acc_cols = p.acc_cols;
    // rows q+1, q+2 of lanes 30 / 31 live in the next warp: next quarter, or quarter 0 of the next M block
    const bool has_next = quarter < 3 || (NT == 32 && u == 0);
    const int nu = quarter < 3 ? u : 1, nq = quarter < 3 ? quarter + 1 : 0;
    const int my_slot = (u * 4 + quarter) * kXSlot + h * 16, next_slot = (nu * 4 + nq) * kXSlot + h * 16;
    const int src1 = (lane + 1) & 31, src2 = (lane + 2) & 31;
    const int l = b * 128 + row;
    const int cb0 = c * 32 + h * 16;   // first channel of this warp inside the N split
    int acc = 0, par = 0;
    uint32_t acc_phase = 0;
    unsigned long long* const dbg = (UNCL_PROBE(1, 1) && warp == 2 && lane == 0) ? p.dbg : nullptr;
    long long w_tfull = 0;
    const long long t_begin = dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const MgItem it = mg_decode(p, item);
      const long long tw0 = dbg ? clock64() : 0;
      mbar_wait(&tfull[acc], acc_phase);
      if (dbg) w_tfull += clock64() - tw0;
      tc_fence_after();
      const uint32_t tacc = tmem_base + lane_base + (uint32_t)(acc * acc_cols) + col;
      uint32_t r0[16], r1[16], r2[16];
      tc_ld16_nw(tacc, r0);
      tc_ld16_nw(tacc + (uint32_t)NT, r1);
      tc_ld16_nw(tacc + (uint32_t)(2 * NT), r2);
      tc_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);   // everything this warp needs is in registers: release the accumulator
      if (UNCL_PROBE(p.probe_noload, 2)) { par ^= 1; if (++acc == nacc) { acc = 0; acc_phase ^= 1; } continue; }
      float* const xpar = xbuf + par * (kUnits * 4 * kXSlot);
      // lanes 0 and 1 publish what the previous warp's lanes 30 / 31 need ...
      if (lane < 2) {
        float4* s = reinterpret_cast<float4*>(xpar + my_slot + (lane == 0 ? 32 : 64));
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            s[j - 8] = make_float4(__uint_as_float(r1[4 * j]), __uint_as_float(r1[4 * j + 1]), __uint_as_float(r1[4 * j + 2]),
                                   __uint_as_float(r1[4 * j + 3]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          s[j] = make_float4(__uint_as_float(r2[4 * j]), __uint_as_float(r2[4 * j + 1]), __uint_as_float(r2[4 * j + 2]),
                             __uint_as_float(r2[4 * j + 3]));
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      // ... and take over the next warp's first rows: nobody in this warp reads lane 0's kx=1 group or lanes 0/1's kx=2
      // groups, so they become the sources the rotating shuffles deliver to lanes 31 / 30, 31
      if (lane < 2 && has_next) {
        const float4* s = reinterpret_cast<const float4*>(xpar + next_slot + (lane == 0 ? 32 : 64));
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 f = s[j - 8];
            r1[4 * j] = __float_as_uint(f.x); r1[4 * j + 1] = __float_as_uint(f.y);
            r1[4 * j + 2] = __float_as_uint(f.z); r1[4 * j + 3] = __float_as_uint(f.w);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = s[j];
          r2[4 * j] = __float_as_uint(f.x); r2[4 * j + 1] = __float_as_uint(f.y);
          r2[4 * j + 2] = __float_as_uint(f.z); r2[4 * j + 3] = __float_as_uint(f.w);
        }
      }
      __syncwarp();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        v[j] = __uint_as_float(r0[j]) + __shfl_sync(0xffffffffu, __uint_as_float(r1[j]), src1) +
               __shfl_sync(0xffffffffu, __uint_as_float(r2[j]), src2);
      const int q = it.q0 + l;
      const int oy = fastdiv(q, p.m_PW), xl = q - oy * PW;
      const int ox = p.x0 + it.band * BW + xl;
      const bool valid = (l < ADV) && (oy < Ho) && (xl < BW) && (ox < Wo);
      const long pix = (long)oy * Wo + ox;
      const int cbase = it.ns * NT + cb0;
      float logit = 0.f;
      if (valid) {
        const float* bias = s_bias + cbase;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaxf(v[g * 8 + j] + bias[g * 8 + j], act_floor);
          if (p.mask != nullptr) {
            float m[8];
            load8(p.mask + (long)it.n * p.mask_img_stride + (long)(cbase / 8 + g) * cb_stride + pix * 8, m);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = m[j] > 0.f ? o[j] : 0.f;
          }
          if (out != nullptr) {
            const long off = (long)it.n * out_img_stride + (long)(cbase / 8 + g) * cb_stride + pix * 8;
            float s2[8], s3[8];
            if (emit_skip) {
#pragma unroll
              for (int j = 0; j < 8; ++j) { s2[j] = o[j] * o[j]; s3[j] = fast_sqrt(o[j] + 1e-8f); }
            }
            if (out_f32) {
              store8(outf + off, o);
              if (emit_skip) { store8(outf + off + skip2, s2); store8(outf + off + skip3, s3); }
            } else {
              store8(out + off, o);
              if (emit_skip) { store8(out + off + skip2, s2); store8(out + off + skip3, s3); }
            }
          }
          if (fuse_outc) {
            const float* ow = s_bias + C_out + cbase + g * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) logit = fmaf(o[j], ow[j], logit);
          }
        }
      }
      if (fuse_outc) {   // the two 16-channel halves of a pixel sit in different warps: combine through shared memory
        float* lp = lbuf + (par * kUnits + u) * 128 + row;
        if (h == 1) *lp = logit;
        asm volatile("bar.sync 2, 512;" ::: "memory");
        if (h == 0 && valid) {
          logit += *lp + outc_b;
          const long o1 = (long)it.n * Ho * Wo + pix;
          if (out_logit) out_logit[o1] = logit;
          out_img[o1] = 1.f / (1.f + __expf(-logit));
        }
      }
      par ^= 1;
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

}  // namespace

namespace {
#ifdef UNCL_PROBES
static const char* probe_env(const char* name) { return getenv(name); }
#else
static const char* probe_env(const char*) { return nullptr; }   // the product library reads no environment variable
#endif
// tile geometry and pipeline sizing: pure host arithmetic (no CUDA calls), shared with uncl_plan_conv3x3_tc_merged
// derive: C_in is the LOGICAL channel count 4*Cs of [skip | up | skip^2 | sqrt(skip)]; the tensor holds the first 2*Cs
int mg_plan(MgParams& p, int N, int C_in, int H, int W, int C_out, int pad, int fuse_outc, int derive, const char* what,
            int* smem_bytes_out, int x0 = 0) {
  p.x0 = x0;
  p.NT = C_out < 64 ? C_out : 64;
  UNCL_REQUIRE(p.NT % 32 == 0 && C_out % p.NT == 0, "%s: unsupported C_out=%d", what, C_out);
  UNCL_REQUIRE(!fuse_outc || C_out == 32, "%s: the fused out conv needs C_out == 32", what);
  p.NS = C_out / p.NT;
  p.NP = 3 * p.NT;
  p.C_out = C_out; p.pad = pad;
  p.Ho = H + 2 * pad - 2; p.Wo = W + 2 * pad - 2;
  // accumulator staging: two stages of 256 TMEM columns (epilogue of tile i overlaps the MMAs of tile i+1)
  p.nacc = 2;
  p.acc_cols = 256;
  p.MB = p.acc_cols / p.NP;   // 2 blocks of 96 columns or 1 block of 192: always two 32-channel work units per tile
  int bw_max = 126;
  if (const char* e = probe_env("UNCL_MG_BW")) { const int want = atoi(e); if (want >= 8 && want < bw_max) bw_max = want; }
  const int tail = 128 + 2 * kUnits * 4 * kXSlot * 4 + 2 * kUnits * 128 * 4 + (2 * kMaxStages + 5) * 8 + 16 + 2 * C_out * 4 + 256;
  const int budget = 227 * 1024 - tail;
  const int wcols = p.Wo - x0;   // output columns of this launch
  const int nbands0 = ceil_div(wcols, bw_max);
  // Fused skip operators hold three ring slots per skip chunk (x, x^2, sqrt): with a ring of exactly one tile (4 slots)
  // the next tile's skip chunk could only load after this tile's derived chunks were built - a serial TMA + transform
  // chain per tile.  Narrower column bands shrink the halo box until FIVE stages fit, which rotates the slots from tile
  // to tile and lets the next skip chunk land early.
  const int want_stages = derive ? 5 : 0;
  int nbands = nbands0;
  for (;; ++nbands) {
  p.BW = ceil_div(wcols, nbands);
  p.PW = p.BW + 2;
  p.band_total = p.Ho * p.PW;
  // Tile = 128*MB consecutive positions of the band flattened with pitch PW.  Flat tiling advances by 128*MB - 2 (tiles
  // overlap by two positions so that the kx sum never crosses a tile) and a tile may start anywhere in a row: the halo
  // box needs ceil rows + 1 + 2.  When one or two whole rows nearly fill the tile (R*PW >= 97 % of it: PW = 124..128) tiles
  // are row-aligned instead: R rows per tile, box = R + 2 rows (3 instead of 5 rows for one 124-wide row, 4 instead of 6
  // for two), no overlap needed because a row ends with its two wrap-around columns.  Measured (1080p frame): 256 -> 32
  // at 124^2 141 -> 130 us; narrower pitches (61, 63) lose more to the unused tile tail than the smaller box saves.
  const int rows_al = (128 * p.MB) / p.PW;
  const bool aligned = rows_al >= 1 && (derive || (rows_al <= 2 && p.PW >= 100)) &&
                       rows_al * p.PW * 100 >= 97 * (128 * p.MB - 2) && probe_env("UNCL_MG_FLAT") == nullptr;
  p.aligned = aligned ? 1 : 0;
  if (aligned) {
    p.ADV = rows_al * p.PW;
    p.tiles_per_band = ceil_div(p.Ho, rows_al);
    p.PH = rows_al + 2;
  } else {
    p.ADV = 128 * p.MB - 2;
    p.tiles_per_band = ceil_div(p.band_total - 2, p.ADV);
    if (p.tiles_per_band < 1) p.tiles_per_band = 1;
    p.PH = (p.PW - 1 + 128 * p.MB - 1) / p.PW + 1 + 2;
  }
  p.tiles_per_img = nbands * p.tiles_per_band;
  UNCL_REQUIRE(p.PW <= 128 && p.PH <= 256, "%s: halo tile too large (%d x %d)", what, p.PW, p.PH);
  p.num_items = N * p.tiles_per_img * p.NS;
  UNCL_REQUIRE(p.num_items < (1 << 24), "%s: too many tiles (%d)", what, p.num_items);
  p.m_NS = (1ull << 40) / (unsigned)p.NS + 1; p.m_tpi = (1ull << 40) / (unsigned)p.tiles_per_img + 1;
  p.m_tpb = (1ull << 40) / (unsigned)p.tiles_per_band + 1; p.m_PW = (1ull << 40) / (unsigned)p.PW + 1;
  p.mma_warps = p.MB == 2 ? 2 : 1;
  if (const char* e = probe_env("UNCL_MMA_WARPS")) { if (atoi(e) == 1) p.mma_warps = 1; }
  // Pipeline stage = the halo box of `ksteps` K = 16 steps (+ their weights).  When the whole filter bank fits next to
  // at least three A stages it is loaded ONCE per CTA and stays resident (the weight stage is otherwise re-fetched from
  // L2 for every tile: 35 % of the L2 -> shared-memory traffic of the 128 -> 32 layer); N' = 96 issues 32 channels per
  // stage (halves the barrier round trips of the issuing warps), N' = 192 too when its weights are resident.
  p.w_total = (C_in / 16) * 3 * 2 * p.NP * 16;
  auto a_bytes = [&](int ks) { return (2 * ks * p.PH * p.PW * 16 + 127) & ~127; };
  int ks = (p.NT == 32 && C_in % 32 == 0) ? 2 : 1;
  p.w_res = (p.NS == 1 && p.w_total + (derive ? 4 : 3) * a_bytes(ks) <= budget && probe_env("UNCL_MG_NO_WRES") == nullptr) ? 1 : 0;
  UNCL_REQUIRE(p.NT == 64 || ks == 2, "%s: C_in must be a multiple of 32 for C_out = 32", what);
  p.ksteps = ks;
  p.nchunk = C_in / (16 * ks);
  p.a_box_bytes = 2 * ks * p.PH * p.PW * 16;
  p.a_stage_bytes = a_bytes(ks);
  p.b_stage_bytes = ks * 3 * 2 * p.NP * 16;
  p.stage_bytes = p.a_stage_bytes + (p.w_res ? 0 : p.b_stage_bytes);
  p.stages = (budget - (p.w_res ? p.w_total : 0)) / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (p.stages >= want_stages || nbands >= nbands0 + 8 || p.BW <= 16) break;
  }
  if (const char* e = probe_env("UNCL_MG_STAGES")) { const int want = atoi(e); if (want >= 2 && want < p.stages) p.stages = want; }
  UNCL_REQUIRE(p.stages >= 2, "%s: tile does not fit shared memory (%d B per stage)", what, p.stage_bytes);
  int smem_bytes = p.stages * p.stage_bytes + (p.w_res ? p.w_total : 0) + tail;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;   // one CTA per SM (each owns all 512 TMEM columns)
  p.derive = derive ? 1 : 0;
  p.H_in = H; p.W_in = W;
  if (derive) {
    // ring program of a tile: per 32-channel skip chunk j  [skip_j (TMA), skip_j^2, sqrt(skip_j)], then the up-sampled chunks
    const int cs32 = C_in / 4 / 32;   // 32-channel chunks of the skip tensor
    UNCL_REQUIRE(p.NT == 32 && p.ksteps == 2 && C_in % 128 == 0 && p.nchunk == 4 * cs32 && p.nchunk <= kMaxProg,
                 "%s: fused skip operators need C_out == 32 and C_skip a multiple of 32 (<= 128)", what);
    UNCL_REQUIRE(p.stages >= 4, "%s: fused skip operators need four pipeline stages (%d fit)", what, p.stages);
    int k = 0;
    for (int j = 0; j < cs32; ++j) {
      p.prog_kind[k] = 0; p.prog_cb[k] = (unsigned char)(4 * j); p.prog_wch[k] = (unsigned char)j; p.prog_back[k] = 1; ++k;
      p.prog_kind[k] = 1; p.prog_cb[k] = 0; p.prog_wch[k] = (unsigned char)(2 * cs32 + j); p.prog_back[k] = 1; ++k;
      p.prog_kind[k] = 2; p.prog_cb[k] = 0; p.prog_wch[k] = (unsigned char)(3 * cs32 + j); p.prog_back[k] = 2; ++k;
    }
    for (int j = 0; j < cs32; ++j) {
      p.prog_kind[k] = 0; p.prog_cb[k] = (unsigned char)(4 * (cs32 + j)); p.prog_wch[k] = (unsigned char)(cs32 + j); p.prog_back[k] = 0; ++k;
    }
  }
  *smem_bytes_out = smem_bytes;
  return UNCL_OK;
}
}  // namespace

static int plan_merged_impl(int N, int C_in, int H, int W, int C_out, int pad, int derive, int* plan) {
  MgParams p{};
  int smem = 0;
  if (int rc = mg_plan(p, N, C_in, H, W, C_out, pad, 0, derive, "conv3x3_tc_plan(merged)", &smem)) return rc;
  const int v[16] = {1 | (p.aligned << 1) | (p.w_res << 2), p.NT, p.NS, p.NP, p.MB, p.ADV, p.PW, p.PH, p.BW, p.tiles_per_img / p.tiles_per_band, p.tiles_per_band,
                     p.num_items, p.stages, p.nacc, p.ksteps, smem};
  for (int i = 0; i < 16; ++i) plan[i] = v[i];
  return UNCL_OK;
}

// Called by uncl_conv3x3_tc (conv_tc.cu) for the layers use_merged() selects; arguments already validated there.
int uncl_plan_conv3x3_tc_merged(int N, int C_in, int H, int W, int C_out, int pad, int* plan) {
  return plan_merged_impl(N, C_in, H, W, C_out, pad, 0, plan);
}

// Tile plan of uncl_conv3x3_tc_skipcat (same 16 fields as uncl_conv3x3_tc_plan): pure host arithmetic.
extern "C" int uncl_conv3x3_tc_skipcat_plan(int N, int C_skip, int H, int W, int C_out, int pad, int* plan) {
  UNCL_REQUIRE(plan != nullptr && N > 0 && C_skip > 0 && C_skip % 32 == 0 && C_out == 32 && (pad == 0 || pad == 2) &&
                   H + 2 * pad - 2 > 0 && W + 2 * pad - 2 > 0,
               "conv3x3_tc_skipcat_plan: unsupported C_skip=%d C_out=%d pad=%d H=%d W=%d", C_skip, C_out, pad, H, W);
  return plan_merged_impl(N, 4 * C_skip, H, W, C_out, pad, 1, plan);
}

int uncl_launch_conv3x3_tc_merged(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                                  long out_img_stride, int out_f32, int N, int C_in, int H, int W, int C_out, int pad,
                                  int act, int emit_skip, int fuse_outc, const float* outc_w, const float* outc_b,
                                  float* out_img, float* out_logit, const void* mask, long mask_img_stride,
                                  unsigned long long* dbg, int derive, int x0, cudaStream_t stream) {
  const char* what = derive ? "conv3x3_tc_skipcat" : "conv3x3_tc(merged)";
  UNCL_REQUIRE(in_img_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0, "%s: input must be 16-byte aligned", what);
  MgParams p{};
  int smem_bytes = 0;
  if (int rc = mg_plan(p, N, C_in, H, W, C_out, pad, fuse_outc, derive, what, &smem_bytes, x0)) return rc;
  p.w = reinterpret_cast<const bf16*>(w_packed);
  p.bias = bias; p.out = out; p.out_f32 = out_f32; p.out_img_stride = out_img_stride;
  p.outc_w = outc_w; p.outc_b = outc_b; p.out_img = out_img; p.out_logit = out_logit;
  p.act = act; p.emit_skip = emit_skip; p.fuse_outc = fuse_outc;
  p.mask = reinterpret_cast<const bf16*>(mask); p.mask_img_stride = mask_img_stride;

  CUtensorMap tmap;
  CUresult r = encode_blocked_bf16(&tmap, in, W, H, (derive ? C_in / 2 : C_in) / 8, N, in_img_stride, p.PW, p.PH, 2 * p.ksteps);
  if (r != CUDA_SUCCESS) return uncl_set_error(UNCL_ECUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
  p.dbg = dbg;
  p.probe_noload = probe_env("UNCL_PROBE_NOLOAD") != nullptr ? 1 : 0;   // bit 0: no loads, 1: no epilogue work, 2: MMA free-runs
  if (const char* e = probe_env("UNCL_PROBE_FLAGS")) p.probe_noload = atoi(e);
  static thread_local int smem_ok = 0, smem_dev = -1;
  static thread_local int smem_ok_d = 0, smem_dev_d = -1;
  cudaError_t e = derive ? ensure_smem(conv3x3_tc_merged_kernel<true>, smem_bytes, smem_ok_d, smem_dev_d)
                         : ensure_smem(conv3x3_tc_merged_kernel<false>, smem_bytes, smem_ok, smem_dev);
  if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "%s: smem attr: %s", what, cudaGetErrorString(e));
  const int sms = sm_count();
  const int grid = p.num_items < sms ? p.num_items : sms;
  if (derive) conv3x3_tc_merged_kernel<true><<<grid, kThreadsDerive, smem_bytes, stream>>>(tmap, p);
  else conv3x3_tc_merged_kernel<false><<<grid, kThreads, smem_bytes, stream>>>(tmap, p);
  return uncl_check_launch(what);
}
