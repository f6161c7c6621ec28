// 3x3 stride-1 convolution / ConvTranspose 3x3 for the NARROW, LARGE layers of the generator (C_out = 32 / 64 at 122..256
// pixels) on the tensor cores: the three ky taps (filter ROWS) merged into the N dimension of one tcgen05.mma, the input
// streamed through shared memory ONE IMAGE ROW AT A TIME, accumulators rotating through TMEM.
//
// Why (profiles/r1_mma_probe.csv): a kind::f16 M=128 K=16 instruction reads its 4 KB A tile at
// 128 B/cycle whatever N is, so N = C_out = 32 runs at 16/40 of the tensor peak.  conv_tc_merged.cu merges the three kx
// taps (N' = 96: 56 cycles for three taps) but pays for it twice: its accumulator rows of one output pixel sit in
// neighbouring TMEM LANES (a shuffle / shared-memory exchange epilogue of ~7500 warp instructions per tile) and every
// 2-row tile reloads a 4-row halo box (2x the input through the shared-memory port the MMAs already saturate).  Here
//     M block  = one input row of a 126-column band (128 positions, pitch 128)
//     D_r[x][ky*C + co]  = sum_{kx, ci} X[r][x + kx][ci] * W[ky][kx][ci][co]      (3 kx x C_in/16 MMAs of N' = 3 C per row;
//                                                                                  kx = descriptor start + kx*16 B)
//     out[y][x][co]      = D_y[x][co] + D_{y+1}[x][C + co] + D_{y+2}[x][2C + co]
// The three terms of an output pixel are in the SAME TMEM lane, and every input row is loaded into shared memory exactly
// once per strip (a CTA walks a strip of R output rows = R + 2 input rows).  Three variants behind one template:
//   slot  (C_out = 32, several K chunks per row, incl. the fused skip operators): five accumulator slots of 96 columns, a
//         row's MMAs are an ordinary K loop into its slot, the epilogue adds the three column groups from three slots;
//   ring  (one-chunk layers, C_out = 64): output row u owns ONE group of C columns at position (-u) mod 16 (mod 8), the
//         N' columns of input row r land on the adjacent groups of outputs r, r-1, r-2, every MMA accumulates, the
//         epilogue reads one group and hands it back cleared;
//   front (inc.conv1): ring of twelve groups; the stage ring is not loaded by TMA but COMPUTED - four extra warps build
//         inc.conv (1 -> 32, unet_parts.py:57-87 with in_ch = 1) from the fp32 image with im2col rows and three-term
//         bf16 split MMAs, so that layer's output never reaches memory.
// The fused skip operators (unet_parts.py:319-322: x^2, sqrt(x + 1e-8) of the skip tensor built in shared memory) use
// the same ring program as conv_tc_merged.cu, one row group at a time - each skip row is transformed once, not twice.
//
// Bands are 126 output columns (TMA box rows hold <= 256 8-byte elements = 128 pixels); the host entry points hand the
// columns past the last whole band (2 of 254, 4 of 256) to the older kernels (x0 argument), see conv_tc.cu.
// Reference operator: models/unet_multi_filters/unet_parts.py:57-87 (double_conv), :126-141 / :183-193 (ConvTranspose
// 3x3 pair of `up`), :311-332 (skip operators + concat), :338-345 (outconv).
// Weights: bf16 [C_in/32][2 ksteps][3 kx][2][3 C_out (ky, co)][8] (packing.conv3x3_tc_rows).
#include <cstdio>
#include <cstdlib>

#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

// warp 0: TMA producer, 1: MMA issuer, then the epilogue warps (sets of four, one per TMEM lane quarter).  An output row
// costs an epilogue warp ~3000 cycles of mostly exposed latency (three barrier waits, three tcgen05.ld round trips, the
// bias / activation / store tail: tools/gpu_rows_probe2.sh), so the one-chunk layers (336 MMA cycles per row) are bound by
// the epilogue warps' instruction stream (six sets were slower than four: 344 -> 550 us on inc.conv1 - more warps only
// dilute the issue slots).  The one-chunk layers therefore use the ring variant below (one tcgen05.ld per output row).
constexpr int kRwSets = 4, kRwSetsDerive = 4;
constexpr int kRwThreads = (2 + 4 * kRwSets) * 32;                 // 576
constexpr int kRwDeriveWarp0 = 2 + 4 * kRwSetsDerive;              // warps 18-21: skip-operator warps, or the first-conv warps of the front mode
constexpr int kRwThreadsDerive = (kRwDeriveWarp0 + 4) * 32;        // 704
constexpr int kRwDeriveWarps = 4;
constexpr int kRwMaxStages = 8;
constexpr int kRwSlots = 5;            // accumulator slots of 96 TMEM columns
constexpr int kRwRing = 16;            // ring variant: accumulator groups of C_out TMEM columns, one per output row (16 x 32 or 8 x 64)
constexpr int kRwMaxProg = 16;
constexpr int kRwFrontScratch = 4 * 2 * 4096;          // front mode: x_hi / x_lo im2col tiles of the four rows of a stage
constexpr int kRwFrontBytes = 2 * kRwFrontScratch + 2048;  // two such buffers (the next stage is built while this one's MMAs run) + the first conv's w_hi / w_lo tiles
constexpr int kRwFrontCols = 384;                      // front mode: TMEM columns 384..511 hold the first conv's accumulators
constexpr int kRwRowBytes = 128 * 16;  // one row of one channel block in shared memory
constexpr int kRwWChunk = 2 * 3 * 2 * 96 * 16;   // packed weights of 32 input channels, C_out = 32: 18432 B (x2 for C_out = 64)

struct RwParams {
  const bf16* w;
  const float* bias;
  bf16* out;
  long out_img_stride;
  const float* outc_w;
  const float* outc_b;
  float* out_img;
  float* out_logit;
  int NT;                              // C_out: 32, or 64 (ring variant only)
  int Ho, Wo, Wc, pad, H_in, W_in;     // Wc: output columns [0, Wc) are computed here
  int BW, nbands, R, nstrips, items_per_img, num_items;
  int G, nchunk, stages, stage_bytes, w_total;
  int derive;
  unsigned char prog_kind[kRwMaxProg], prog_cb[kRwMaxProg], prog_wch[kRwMaxProg], prog_back[kRwMaxProg];
  int act, emit_skip, fuse_outc;
  // front mode (inc.conv fused in front of inc.conv1): the stage ring is not loaded by TMA but COMPUTED by four extra warps from
  // the fp32 image - im2col rows in shared memory, three-term bf16 split MMAs (x_hi w_hi + x_lo w_hi + x_hi w_lo), bias + ReLU
  const float* fx;      // [N][H0][W0] fp32 image (1 channel)
  const bf16* fw;       // packing.conv_first_rows: [2 (hi, lo)][2 halves][32][8] bf16, tap 9 = the bias
  long fx_img_stride;
  int front, H0, W0, f_bytes;
  int ring;    // one-chunk layers: ring variant of the accumulators (host switch: 1 by default)
  int probe;   // -DUNCL_PROBES build only (tools/): 1 = no TMA traffic after the first ring fill, 2 = epilogue does the barrier protocol only, 4 = one MMA per row instead of six (results are wrong while set), 8 = print the per-role cycle accounting of CTA 0
};

struct RwItem {
  int n, bx, by, y0, x0, rows_out, rows_in;
};

__device__ __forceinline__ RwItem rw_decode(const RwParams& p, int item) {
  RwItem it;
  it.n = item / p.items_per_img;
  const int t = item - it.n * p.items_per_img;
  const int band = t / p.nstrips;
  const int strip = t - band * p.nstrips;
  it.y0 = strip * p.R;
  it.x0 = band * p.BW;
  it.rows_out = min(p.R, p.Ho - it.y0);
  it.rows_in = it.rows_out + 2;
  it.bx = it.x0 - p.pad;
  it.by = it.y0 - p.pad;
  return it;
}

__device__ __forceinline__ uint32_t rw_bf16x2_mul(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t rw_bf16x2_sqrt_eps(uint32_t v) {
  const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
  return pack_bf16x2(fast_sqrt(lo + 1e-8f), fast_sqrt(hi + 1e-8f));
}

// two floats -> bf16x2 with ReLU (low half = first argument, like pack_bf16x2)
__device__ __forceinline__ uint32_t rw_pack_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// 32 lanes x 32 columns of zeros into TMEM (ring variant: an accumulator group is handed back cleared)
__device__ __forceinline__ void rw_tmem_zero32(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z)
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <bool kDerive, int G, bool kRing, int NT, bool kFront>
__global__ void __launch_bounds__((kDerive || kFront) ? kRwThreadsDerive : kRwThreads, 1)
conv3x3_tc_rows_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ RwParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* stage_base = smem;
  uint8_t* wres = smem + (size_t)p.stages * p.stage_bytes;            // resident filter bank
  // 128 B of slack (the kx = 2 reads of a stage's last row), then [32] bias + [32] out conv weights, 16-byte aligned: the
  // epilogue reads them as broadcast LDS.128 - every shared-memory wavefront competes with the MMAs' operand reads
  uint8_t* fscr = wres + p.w_total + 128;                             // front mode: im2col tiles + first-conv filter tiles
  float* s_bias = reinterpret_cast<float*>(fscr + (kFront ? kRwFrontBytes : 0));
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 128);   // [0, NT) bias, [64, 96) out conv weights, [96, 128) first-conv bias
  uint64_t* full = bars;
  uint64_t* empty = bars + kRwMaxStages;
  uint64_t* rowdone = bars + 2 * kRwMaxStages;
  uint64_t* sfree = rowdone + kRwRing;
  uint64_t* wfull = sfree + kRwRing;
  uint64_t* fbar = wfull + 1;                                         // front mode: the first conv's MMAs of a stage are done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.num_items, nchunk = p.nchunk, stages = p.stages;
  constexpr int stage_bytes = G * 4 * kRwRowBytes;

  if (warp == 0 && lane == 0) {
    if (!kFront) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    const uint32_t extra = kDerive ? (uint32_t)kRwDeriveWarps : 0u;
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], kFront ? 4u : 1 + extra); mbar_init(&empty[s], 1 + extra); }
    mbar_init(fbar, 1);
    // a slot is read by the epilogues of three output rows (its ky = 0, 1, 2 column groups), four warps each
    for (int s = 0; s < kRwRing; ++s) { mbar_init(&rowdone[s], 1); mbar_init(&sfree[s], kRing ? 4 : 12); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  static_assert(NT == 32 || (NT == 64 && kRing && !kDerive), "C_out = 64 exists as the ring variant only");
  static_assert(!kFront || (kRing && !kDerive && NT == 32 && G == 4), "front mode: one-chunk ring variant, C_out = 32");
  constexpr int kGroups = kFront ? kRwFrontCols / 32 : 512 / NT;   // ring groups: 16 x 32 or 8 x 64 columns; 12 x 32 next to the front accumulators
  constexpr int kWChunk = kRwWChunk * (NT / 32);
  for (int i = threadIdx.x; i < 64; i += (int)blockDim.x) {
    s_bias[i] = (p.bias && i < NT) ? p.bias[i] : 0.f;
    s_bias[64 + i] = (p.fuse_outc && i < 32) ? p.outc_w[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if constexpr (kRing) {   // the MMAs only ever accumulate: all sixteen groups start cleared
    if (warp >= 2 && warp < 2 + 4 * kRwSets) {
      const uint32_t lb = (uint32_t)((warp & 3) * 32) << 16;
      for (int g = (warp - 2) >> 2; g < 16; g += kRwSets) rw_tmem_zero32(tmem_base + lb + (uint32_t)(g * 32));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      mbar_expect_tx(wfull, (uint32_t)p.w_total + (kFront ? 2048u : 0u));
      {
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w);
        for (int off = 0; off < p.w_total; off += kWChunk) bulk_load(wres + off, wsrc + off, kWChunk, wfull);
        if (kFront) bulk_load(fscr + 2 * kRwFrontScratch, p.fw, 2048, wfull);
      }
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t box_bytes = (uint32_t)stage_bytes;
      for (int item = blockIdx.x; !kFront && item < num_items; item += gridDim.x) {
        const RwItem it = rw_decode(p, item);
        for (int r0 = 0; r0 < it.rows_in; r0 += G) {
          for (int ch = 0; ch < nchunk; ++ch) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = stage_base + (size_t)stage * stage_bytes;
            if (UNCL_PROBE(p.probe, 1) && (item != (int)blockIdx.x || r0 >= G * stages)) {
              mbar_arrive(&full[stage]);
            } else if (!kDerive || p.prog_kind[ch] == 0) {
              mbar_expect_tx(&full[stage], box_bytes);
              tma_load_4d(sa, &tmap, &full[stage], it.bx * 2, it.by + r0, kDerive ? (int)p.prog_cb[ch] : ch * 4, it.n);
            } else {
              mbar_arrive(&full[stage]);   // derived chunk: filled by the skip-operator warps
            }
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // One warp: the rows of a strip accumulate into rotating TMEM slots and a row's MMAs are ordinary K-loop accumulation
    // into ONE slot (first MMA overwrites).  The issuing warp is a serial, latency-bound instruction stream (R2UR, uniform
    // adds, barrier polls: profiles/README.md), so one elected lane issues ALL rows of a (row group, K chunk) stage as
    // straight-line code - descriptor offsets are immediates (G is a template parameter) - and polls the slot barriers
    // itself between rows.
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(3 * NT >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
    constexpr uint32_t a_lo_const = ((uint32_t)(G * 128) & 0x3fffu) << 16;   // LBO_A: channel-block stride = G rows x 2048 B
    constexpr uint32_t b_lo_const = (uint32_t)(3 * NT) << 16;                // LBO_B = 3 C_out rows x 16 B
    constexpr uint32_t b_tile_16 = (uint32_t)(2 * 3 * NT);                   // one (kstep, kx) tile of the filter bank
    constexpr uint32_t stage_16 = (uint32_t)stage_bytes >> 4;
    constexpr uint32_t a_kstep_16 = (uint32_t)(2 * G * 128);
    const uint32_t stage0_16 = smem_u32(stage_base) >> 4;
    const uint32_t wres_16 = smem_u32(wres) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int slot0 = 0;          // slot of the next row group's first row
    uint32_t par0 = 1;      // parity its sfree wait uses: (use count & 1) ^ 1
    mbar_wait(wfull, 0);
#ifdef UNCL_PROBES
    long long mw_full = 0, mw_issue = 0, mrows = 0;
    const long long mt0 = clock64();
#endif
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const RwItem it = rw_decode(p, item);
      for (int r0 = 0; r0 < it.rows_in; r0 += G) {
        const int rows = min(G, it.rows_in - r0);
        for (int ch = 0; ch < nchunk; ++ch) {
#ifdef UNCL_PROBES
          const long long ma = clock64();
#endif
          mbar_wait(&full[stage], phase);
#ifdef UNCL_PROBES
          const long long mb = clock64();
          mw_full += mb - ma;
#endif
          tc_fence_after();
          const uint32_t sa16 = stage0_16 + (uint32_t)stage * stage_16;
          const uint32_t wch = kDerive ? (uint32_t)p.prog_wch[ch] : (uint32_t)ch;
          const uint32_t b_base = b_lo_const | (wres_16 + wch * (uint32_t)(kWChunk >> 4));
          const bool first = ch == 0, last = ch == nchunk - 1;
          if constexpr (kRing) {
            // Ring variant (one K chunk per row): output row u owns ONE group of 32 TMEM columns at position (-u) mod 16,
            // so the N' = 96 columns (ky = 0, 1, 2) of input row tt land on the groups of outputs tt, tt-1, tt-2, which are
            // adjacent - every MMA accumulates, the epilogue reads one group and hands it back cleared.  Rows whose three
            // groups wrap around the ring (2 of 16) issue an N = 64 and an N = 32 instruction instead of one N = 96.
            constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(2 * NT >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t kg = (uint32_t)kGroups;
            if (elect_one()) {
              uint32_t tt = (uint32_t)slot0;   // ring variant: slot0 counts rows (never wrapped at five)
#pragma unroll
              for (int r = 0; r < G; ++r) {
                if (r < rows) {
                  const uint32_t gi = tt % kg;
                  if (first) {   // group of output tt: cleared by its previous owner
                    mbar_wait(&sfree[gi], (((tt + 2u) / kg) & 1u) ^ 1u);
                    tc_fence_after();
                  }
                  const uint32_t pos = (kg - gi) % kg;
                  const uint32_t d = tmem_base + pos * (uint32_t)NT;
                  const uint32_t a_row = a_lo_const | (sa16 + (uint32_t)(r * 128));
                  if (pos <= kg - 3u) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                      for (int kx = 0; kx < 3; ++kx) {
                        if (UNCL_PROBE(p.probe, 4) && (ks > 0 || kx > 0)) continue;
                        tc_mma_bf16(d, a_row + (uint32_t)(ks * a_kstep_16 + kx), desc_hi, b_base + (uint32_t)(ks * 3 + kx) * b_tile_16,
                                    desc_hi, idesc, 1u);
                      }
                    }
                  } else {
                    const bool two_first = pos == kg - 2u;   // groups (last - 1, last | 0)   or   (last | 0, 1)
                    const uint32_t id_a = two_first ? idesc2 : idesc1, id_b = two_first ? idesc1 : idesc2;
                    const uint32_t nb = two_first ? (uint32_t)(2 * NT) : (uint32_t)NT;   // B rows (16 B each) of the first instruction
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                      for (int kx = 0; kx < 3; ++kx) {
                        const uint32_t a = a_row + (uint32_t)(ks * a_kstep_16 + kx), b = b_base + (uint32_t)(ks * 3 + kx) * b_tile_16;
                        tc_mma_bf16(d, a, desc_hi, b, desc_hi, id_a, 1u);
                        tc_mma_bf16(tmem_base, a, desc_hi, b + nb, desc_hi, id_b, 1u);
                      }
                    }
                  }
                  if (last) tc_commit(&rowdone[gi]);
                  ++tt;
                }
              }
              tc_commit(&empty[stage]);
            }
          } else
          if (elect_one()) {
            int slot = slot0;
            uint32_t par = par0;
#pragma unroll
            for (int r = 0; r < G; ++r) {
              if (r < rows) {
                if (first) {   // the slot's previous contents have been read by the three epilogues that needed them
                  mbar_wait(&sfree[slot], par);
                  tc_fence_after();
                }
                const uint32_t d = tmem_base + (uint32_t)(slot * 96);
                const uint32_t a_row = a_lo_const | (sa16 + (uint32_t)(r * 128));
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                  for (int kx = 0; kx < 3; ++kx) {
                    if (UNCL_PROBE(p.probe, 4) && (ks > 0 || kx > 0)) continue;
                    tc_mma_bf16(d, a_row + (uint32_t)(ks * a_kstep_16 + kx), desc_hi, b_base + (uint32_t)(ks * 3 + kx) * b_tile_16,
                                desc_hi, idesc, (ks > 0 || kx > 0) ? 1u : (first ? 0u : 1u));
                  }
                }
                if (last) tc_commit(&rowdone[slot]);
                if (++slot == kRwSlots) { slot = 0; par ^= 1; }
              }
            }
            tc_commit(&empty[stage]);
          }
          __syncwarp();
#ifdef UNCL_PROBES
          mw_issue += clock64() - mb;
#endif
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
#ifdef UNCL_PROBES
        mrows += rows;
#endif
        slot0 += rows;
        if (!kRing && slot0 >= kRwSlots) { slot0 -= kRwSlots; par0 ^= 1; }
      }
    }
#ifdef UNCL_PROBES
    if ((p.probe & 8) && blockIdx.x == 0 && lane == 0)
      printf("rows probe: MMA warp total %lld cycles, %lld rows, per row: wait full %lld, issue block (incl. slot waits) %lld\n",
             clock64() - mt0, mrows, mw_full / max(mrows, 1ll), mw_issue / max(mrows, 1ll));
#endif
    __syncwarp();
  } else if (kDerive && warp >= kRwDeriveWarp0) {
    // =============================== fused skip operators ===============================
    // Same ring program as conv_tc_merged.cu: at a TMA position these warps only add their arrival; at the `square`
    // position they wait for the skip chunk (previous ring position), and write x*x into this stage and sqrt(x + 1e-8)
    // into the next one at identical offsets (the shared-memory image IS the MMA operand layout).  The zero padding of
    // the transposed conv applies to the concatenated tensor: the square root is 0 outside the image, not sqrt(eps).
    if constexpr (kDerive) {
      const int dtid = (warp - kRwDeriveWarp0) * 32 + lane;
      const int npos = G * 128, H_in = p.H_in, W_in = p.W_in;
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const RwItem it = rw_decode(p, item);
        for (int r0 = 0; r0 < it.rows_in; r0 += G) {
          for (int ch = 0; ch < nchunk; ++ch) {
            const int kind = p.prog_kind[ch];
            if (kind == 2) {   // built together with the square (previous position)
              if (++stage == stages) { stage = 0; phase ^= 1; }
              continue;
            }
            mbar_wait(&empty[stage], phase ^ 1);
            if (kind != 0) {
              int src = stage - 1, nst = stage + 1;
              uint32_t src_phase = phase, n_phase = phase;
              if (src < 0) { src += stages; src_phase ^= 1; }
              if (nst == stages) { nst = 0; n_phase ^= 1; }
              mbar_wait(&empty[nst], n_phase ^ 1);
              mbar_wait(&full[src], src_phase);
              const uint8_t* sp = stage_base + (size_t)src * stage_bytes;
              uint8_t* dp = stage_base + (size_t)stage * stage_bytes;
              uint8_t* dq = stage_base + (size_t)nst * stage_bytes;
              for (int pos = dtid; pos < npos; pos += kRwDeriveWarps * 32) {
                uint4 v[4], q[4];
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) v[cb] = *reinterpret_cast<const uint4*>(sp + (cb * npos + pos) * 16);
                const int rr = pos >> 7, cc = pos & 127;
                const uint32_t in0 = ((unsigned)(it.by + r0 + rr) < (unsigned)H_in && (unsigned)(it.bx + cc) < (unsigned)W_in) ? 0xffffffffu : 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  q[j].x = rw_bf16x2_sqrt_eps(v[j].x) & in0; q[j].y = rw_bf16x2_sqrt_eps(v[j].y) & in0;
                  q[j].z = rw_bf16x2_sqrt_eps(v[j].z) & in0; q[j].w = rw_bf16x2_sqrt_eps(v[j].w) & in0;
                  v[j].x = rw_bf16x2_mul(v[j].x, v[j].x); v[j].y = rw_bf16x2_mul(v[j].y, v[j].y);
                  v[j].z = rw_bf16x2_mul(v[j].z, v[j].z); v[j].w = rw_bf16x2_mul(v[j].w, v[j].w);
                }
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                  *reinterpret_cast<uint4*>(dp + (cb * npos + pos) * 16) = v[cb];
                  *reinterpret_cast<uint4*>(dq + (cb * npos + pos) * 16) = q[cb];
                }
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> the MMA's async-proxy reads
              __syncwarp();
              if (lane == 0) {
                mbar_arrive(&full[stage]); mbar_arrive(&empty[stage]);
                mbar_arrive(&full[nst]); mbar_arrive(&empty[nst]);
                mbar_arrive(&empty[src]);
              }
            } else {
              if (lane == 0) {
                mbar_arrive(&full[stage]);
                if (p.prog_back[ch] == 0) mbar_arrive(&empty[stage]);   // a TMA chunk nothing is derived from
              }
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (kFront && warp >= kRwDeriveWarp0) {
    // =============================== first conv (front mode) ===============================
    // inc.conv (Conv2d 1 -> 32, 3x3, ReLU: unet_parts.py:57-87) computed straight into the stage ring: these four warps are
    // the 128 positions of a row.  Per stage (four rows of inc.conv's output): every lane gathers the 3x3 fp32 neighbourhood
    // of its pixel and writes it as two bf16 im2col rows (x_hi, x_lo: K = 9 taps padded to 16) in the K-major operand
    // layout; one elected lane issues x_hi w_hi + x_lo w_hi + x_hi w_lo (three N = 32 MMAs per row, ~2^-16 relative like the
    // other split GEMMs; the bias is a tenth tap whose input is the constant 1) into the TMEM columns behind the ring; then
    // each warp reads its lane quarter back, applies ReLU inside the bf16 conversion and stores the 32 bf16 channels of its pixel where the TMA box of the unfused kernel would have put them.
    if constexpr (kFront) {
      const int fw_ = warp - kRwDeriveWarp0;        // 0..3; TMEM lane quarter = warp % 4
      const int quarter = warp & 3, pos = quarter * 32 + lane;
      const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
      constexpr uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t fscr_16 = smem_u32(fscr) >> 4;
      const uint32_t a_lo_c = 128u << 16;            // LBO_A: 128 positions x 16 B between the two K halves
      const uint32_t b_lo_c = 32u << 16;             // LBO_B: 32 filter rows x 16 B
      const uint32_t bw_16 = fscr_16 + (uint32_t)(2 * kRwFrontScratch >> 4);
      const int H0 = p.H0, W0 = p.W0, Ha = H0 - 2, Wa = W0 - 2;
      // im2col rows (x_hi, x_lo) of the four output rows of the stage that starts at row `ya0` of image `n`, into buffer `buf`
      // the 6 x 3 input pixels behind this lane's four output pixels (issued early: their latency hides behind the epilogue)
      auto fetch = [&](int n, int ya0, int xa, float (&v)[G + 2][3]) {
        const float* xn = p.fx + (long)n * p.fx_img_stride;
        const bool xin = xa >= 0 && xa + 2 < W0;
#pragma unroll
        for (int j = 0; j < G + 2; ++j) {
          const int yi = ya0 + j;
          const bool in = xin && yi >= 0 && yi < H0;
#pragma unroll
          for (int c = 0; c < 3; ++c) v[j][c] = in ? __ldg(xn + (long)yi * W0 + xa + c) : 0.f;
        }
      };
      // ... split x = x_hi + x_lo (x_hi kept as an fp32 whose low 16 bits are zero: two of them pack into a bf16x2 with one PRMT)
      // and written as the im2col rows of the four output rows
      auto build = [&](const float (&v)[G + 2][3], int buf) {
        uint32_t hb[G + 2][3];
        float lo[G + 2][3];
#pragma unroll
        for (int j = 0; j < G + 2; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float h = __bfloat162float(__float2bfloat16_rn(v[j][c]));
            hb[j][c] = __float_as_uint(h);
            lo[j][c] = v[j][c] - h;
          }
        }
#pragma unroll
        for (int r = 0; r < G; ++r) {
          // taps k = 3 ky + kx -> pixel (r + ky, kx); tap 9 of x_hi is the constant 1 that multiplies the bias row of the filters
          uint8_t* th = fscr + buf * kRwFrontScratch + r * 8192 + pos * 16;   // x_hi tile of row r: [2 halves][128 positions][8 taps]
          uint4 h0, l0;
          h0.x = __byte_perm(hb[r][0], hb[r][1], 0x7632);     h0.y = __byte_perm(hb[r][2], hb[r + 1][0], 0x7632);
          h0.z = __byte_perm(hb[r + 1][1], hb[r + 1][2], 0x7632); h0.w = __byte_perm(hb[r + 2][0], hb[r + 2][1], 0x7632);
          l0.x = pack_bf16x2(lo[r][0], lo[r][1]);             l0.y = pack_bf16x2(lo[r][2], lo[r + 1][0]);
          l0.z = pack_bf16x2(lo[r + 1][1], lo[r + 1][2]);     l0.w = pack_bf16x2(lo[r + 2][0], lo[r + 2][1]);
          *reinterpret_cast<uint4*>(th) = h0;
          *reinterpret_cast<uint4*>(th + 2048) = make_uint4((hb[r + 2][2] >> 16) | 0x3f800000u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(th + 4096) = l0;
          *reinterpret_cast<uint4*>(th + 4096 + 2048) = make_uint4(pack_bf16x2(lo[r + 2][2], 0.f), 0u, 0u, 0u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      };
      // x_hi w_hi + x_lo w_hi + x_hi w_lo of the four rows in buffer `buf` into the front accumulators (one elected lane)
      auto issue = [&](int buf) {
        if (fw_ == 0) {
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int r = 0; r < G; ++r) {
              const uint32_t d = tmem_base + (uint32_t)(kRwFrontCols + 32 * r);
              const uint32_t xh = a_lo_c | (fscr_16 + (uint32_t)(buf * (kRwFrontScratch >> 4) + r * 512)), xl = xh + 256u;
              tc_mma_bf16(d, xh, desc_hi, b_lo_c | bw_16, desc_hi, idesc1, 0u);
              tc_mma_bf16(d, xl, desc_hi, b_lo_c | bw_16, desc_hi, idesc1, 1u);
              tc_mma_bf16(d, xh, desc_hi, b_lo_c | (bw_16 + 64u), desc_hi, idesc1, 1u);
            }
            tc_commit(fbar);
          }
          __syncwarp();
        }
      };
      int stage = 0, buf = 0;
      uint32_t phase = 0, fpar = 0;
      if (fw_ == 0) mbar_wait(wfull, 0);             // the filter tiles have landed (only the issuing warp needs to know)
      int item = blockIdx.x, r0 = 0;
      RwItem it = rw_decode(p, item);
      float xv[G + 2][3];
      fetch(it.n, it.by, it.bx + pos, xv);
      build(xv, 0);
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      issue(0);
      while (item < num_items) {
        // the stage after this one (same strip, or the first of the CTA's next strip): built while this one's MMAs run
        int nitem = item, nr0 = r0 + G;
        RwItem nit = it;
        if (nr0 >= it.rows_in) {
          nitem = item + (int)gridDim.x; nr0 = 0;
          if (nitem < num_items) nit = rw_decode(p, nitem);
        }
        const bool has_next = nitem < num_items;
        if (has_next) fetch(nit.n, nit.by + nr0, nit.bx + pos, xv);
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_wait(fbar, fpar);
        fpar ^= 1;
        tc_fence_after();
        // ---- bias, ReLU, bf16: the stage image [4 channel blocks][G rows][128 positions][8 channels]
        uint8_t* sa = stage_base + (size_t)stage * stage_bytes + pos * 16;
        const int xa = it.bx + pos;
#pragma unroll
        for (int r = 0; r < G; ++r) {
          uint32_t rb[32];
          tc_ld32(tmem_base + lane_base + (uint32_t)(kRwFrontCols + 32 * r), rb);
          const int ya = it.by + r0 + r;
          const bool in = ya >= 0 && ya < Ha && xa >= 0 && xa < Wa;
#pragma unroll
          for (int g = 0; g < 4; ++g) {   // the bias arrived through tap 9; ReLU inside the bf16 conversion
            uint4 o;
            o.x = rw_pack_relu(__uint_as_float(rb[8 * g + 0]), __uint_as_float(rb[8 * g + 1]));
            o.y = rw_pack_relu(__uint_as_float(rb[8 * g + 2]), __uint_as_float(rb[8 * g + 3]));
            o.z = rw_pack_relu(__uint_as_float(rb[8 * g + 4]), __uint_as_float(rb[8 * g + 5]));
            o.w = rw_pack_relu(__uint_as_float(rb[8 * g + 6]), __uint_as_float(rb[8 * g + 7]));
            if (!in) o = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sa + g * (G * kRwRowBytes) + r * kRwRowBytes) = o;
          }
        }
        if (has_next) build(xv, buf ^ 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");   // every warp has read its accumulators and built its share of the next tiles
        if (has_next) issue(buf ^ 1);
        if (lane == 0) mbar_arrive(&full[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
        buf ^= 1;
        item = nitem; r0 = nr0; it = nit;
      }
    }
  } else {
    // =============================== epilogue ===============================
    // kSets sets x 4 TMEM lane quarters; set s finishes the output rows u with u % kSets == s.  Output row u (CTA-wide
    // row counter; rows u, u+1, u+2 are its three input rows) takes the ky = 0 / 1 / 2 column group of slot u / u+1 / u+2
    // as each of those rows completes and releases its share of that slot at once.  Rows whose three inputs are not in
    // one strip (the last two of a strip, and the virtual rows -2, -1 before the first) only do the barrier protocol.
    constexpr int kSets = kDerive ? kRwSetsDerive : kRwSets;
    const int quarter = warp & 3, set = (warp - 2) >> 2;
    const int xl = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int Ho = p.Ho, Wo = p.Wo, Wc = p.Wc, BW = p.BW;
    const long cb_stride = (long)Ho * Wo * 8;
    const float act_floor = (p.act == UNCL_ACT_RELU) ? 0.f : -INFINITY;
    const bool emit_skip = p.emit_skip != 0, fuse_outc = p.fuse_outc != 0;
    bf16* const out = p.out;
    const float outc_b = fuse_outc ? __ldg(p.outc_b) : 0.f;
    int t = 0;
#ifdef UNCL_PROBES
    long long pw[3] = {0, 0, 0}, pwork = 0, prows = 0;
    const long long pt0 = clock64();
#endif
    for (int item = (int)blockIdx.x - (int)gridDim.x; item < num_items; item += (int)gridDim.x) {
      // item < 0: the two virtual rows before the CTA's first strip
      const bool virt = item < 0;
      const RwItem it = rw_decode(p, virt ? 0 : item);
      const int rows_in = virt ? 2 : it.rows_in, rows_out = virt ? 0 : it.rows_out;
      const bool last_item = !virt && item + (int)gridDim.x >= num_items;
      const int tb = virt ? -2 : t;
      for (int j = 0; j < rows_in; ++j) {
        const int u = tb + j;
        if ((u + 2 * kSets) % kSets != set) continue;
        const bool valid = j < rows_out;
        if (!valid && last_item) break;   // no later row exists: nothing waits for these slots any more
        // Slot r is read right after row r completes by all three output rows that need it (r, r-1, r-2), so it is free
        // again one epilogue turn after its own MMAs - the issuing warp can run up to four rows ahead of the epilogue.
        const int oy = it.y0 + j, ox = it.x0 + xl;
        const bool store = valid && !UNCL_PROBE(p.probe, 2) && xl < BW && ox < Wc;
        const long pix = (long)oy * Wo + ox;
        float logit = outc_b;
        // bias, activation, bf16 store (+ skip planes / out conv) of the 32 channels [32 h, 32 h + 32) of this lane's pixel
        auto finish = [&](const float (&v)[32], int h) {
          const float4* sb4 = reinterpret_cast<const float4*>(s_bias) + 8 * h;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 b0 = sb4[2 * g], b1 = sb4[2 * g + 1];
            float o[8];
            o[0] = fmaxf(v[g * 8 + 0] + b0.x, act_floor); o[1] = fmaxf(v[g * 8 + 1] + b0.y, act_floor);
            o[2] = fmaxf(v[g * 8 + 2] + b0.z, act_floor); o[3] = fmaxf(v[g * 8 + 3] + b0.w, act_floor);
            o[4] = fmaxf(v[g * 8 + 4] + b1.x, act_floor); o[5] = fmaxf(v[g * 8 + 5] + b1.y, act_floor);
            o[6] = fmaxf(v[g * 8 + 6] + b1.z, act_floor); o[7] = fmaxf(v[g * 8 + 7] + b1.w, act_floor);
            if (out != nullptr) {
              bf16* op = out + (long)it.n * p.out_img_stride + (long)(4 * h + g) * cb_stride + pix * 8;
              store8(op, o);
              if (emit_skip) {   // concat buffer [skip | up-sampled | skip^2 | sqrt(skip + eps)], C_out channels each
                float s2[8], s3[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) { s2[c] = o[c] * o[c]; s3[c] = fast_sqrt(o[c] + 1e-8f); }
                store8(op + (2 * NT / 8) * cb_stride, s2);
                store8(op + (3 * NT / 8) * cb_stride, s3);
              }
            }
            if (NT == 32 && fuse_outc) {
              const float4 w0 = sb4[16 + 2 * g], w1 = sb4[16 + 2 * g + 1];
              logit = fmaf(o[0], w0.x, logit); logit = fmaf(o[1], w0.y, logit); logit = fmaf(o[2], w0.z, logit);
              logit = fmaf(o[3], w0.w, logit); logit = fmaf(o[4], w1.x, logit); logit = fmaf(o[5], w1.y, logit);
              logit = fmaf(o[6], w1.z, logit); logit = fmaf(o[7], w1.w, logit);
            }
          }
        };
#ifdef UNCL_PROBES
        long long pb = clock64();
#endif
        if constexpr (kRing) {
          const int r2 = u + 2, gu = (u + kGroups) % kGroups;
#ifdef UNCL_PROBES
          const long long pa = clock64();
#endif
          mbar_wait(&rowdone[r2 % kGroups], (uint32_t)((r2 / kGroups) & 1));   // all three contributions have landed
#ifdef UNCL_PROBES
          pw[2] += clock64() - pa;
          pb = clock64();
#endif
          tc_fence_after();
          const uint32_t taddr = tmem_base + lane_base + (uint32_t)(((kGroups - gu) % kGroups) * NT);
#pragma unroll
          for (int h = 0; h < NT / 32; ++h) {
            float v[32];
            if (valid && !UNCL_PROBE(p.probe, 2)) {
              uint32_t rb[32];
              tc_ld32(taddr + (uint32_t)(32 * h), rb);
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(rb[c]);
            }
            rw_tmem_zero32(taddr + (uint32_t)(32 * h));
            if (h == NT / 32 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&sfree[gu]);
            }
            if (store) finish(v, h);
          }
        } else {
          float v[32];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int r = u + k;
            if (r < 0) continue;
            const int slot = r % kRwSlots;
#ifdef UNCL_PROBES
            const long long pa = clock64();
#endif
            mbar_wait(&rowdone[slot], (uint32_t)((r / kRwSlots) & 1));
#ifdef UNCL_PROBES
            pw[k] += clock64() - pa;
#endif
            tc_fence_after();
            if (valid && !UNCL_PROBE(p.probe, 2)) {
              uint32_t rb[32];
              tc_ld32(tmem_base + lane_base + (uint32_t)(slot * 96 + k * 32), rb);
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = k == 0 ? __uint_as_float(rb[c]) : v[c] + __uint_as_float(rb[c]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sfree[slot]);
          }
#ifdef UNCL_PROBES
          pb = clock64();
#endif
          if (store) finish(v, 0);
        }
        if (NT == 32 && fuse_outc && store) {
          const long o1 = (long)it.n * Ho * Wo + pix;
          if (p.out_logit) p.out_logit[o1] = logit;
          p.out_img[o1] = 1.f / (1.f + __expf(-logit));
        }
        if (!valid) continue;
#ifdef UNCL_PROBES
        pwork += clock64() - pb; ++prows;
#endif
      }
      if (!virt) t += rows_in;
    }
#ifdef UNCL_PROBES
    if ((p.probe & 8) && blockIdx.x == 0 && warp == 2 && lane == 0)
      printf("rows probe: epilogue warp total %lld cycles, %lld rows, per row: wait A %lld B %lld C %lld, finish %lld\n", clock64() - pt0,
             prows, pw[0] / max(prows, 1ll), pw[1] / max(prows, 1ll), pw[2] / max(prows, 1ll), pwork / max(prows, 1ll));
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// Strip height: the fewest waves of (image, band, strip) items over the SMs, weighted by the rows a strip really
// computes (R + 2 input rows for R output rows).
int rw_pick_R(int N, int nbands, int Ho, int sms) {
  int best_R = Ho < 8 ? Ho : 8;
  double best = 1e30;
  for (int ns = 1; ns <= Ho; ++ns) {
    const int R = ceil_div(Ho, ns);
    if (R < 6 && ns > 1) break;
    const int strips = ceil_div(Ho, R);
    const long items = (long)N * nbands * strips;
    const long waves = (items + sms - 1) / sms;
    const double cost = (double)waves * (R + 2) + 0.5 * waves;   // + per-strip pipeline fill
    if (cost < best) { best = cost; best_R = R; }
  }
  return best_R;
}

int rw_plan(RwParams& p, int N, int C_in, int H, int W, int C_out, int pad, int Wc, int derive, int sms, const char* what,
            int* smem_bytes_out, int front = 0) {
  p.front = front; p.f_bytes = front ? kRwFrontBytes : 0;
  UNCL_REQUIRE(C_out == 32 || (C_out == 64 && !derive), "%s: C_out must be 32 (or 64 without fused skip operators), got %d", what, C_out);
  p.NT = C_out;
  p.pad = pad; p.H_in = H; p.W_in = W;
  p.Ho = H + 2 * pad - 2; p.Wo = W + 2 * pad - 2;
  UNCL_REQUIRE(p.Ho > 0 && p.Wo > 0 && Wc > 0 && Wc <= p.Wo, "%s: bad extent (Ho=%d Wo=%d cols=%d)", what, p.Ho, p.Wo, Wc);
  UNCL_REQUIRE(C_in % 32 == 0 && C_in >= 32, "%s: C_in must be a multiple of 32", what);
  p.Wc = Wc;
  p.nbands = ceil_div(Wc, 126);
  p.BW = ceil_div(Wc, p.nbands);
  p.derive = derive ? 1 : 0;
  p.nchunk = C_in / 32;
  p.w_total = p.nchunk * kRwWChunk * (C_out / 32);
  UNCL_REQUIRE(p.nchunk <= kRwMaxProg, "%s: C_in=%d too deep", what, C_in);
  const int tail = 128 + 128 + (2 * kRwMaxStages + 2 * kRwRing + 2) * 8 + 16 + 128 * 4 + 256;
  const int budget = 227 * 1024 - tail - p.w_total - p.f_bytes;
  // Rows per pipeline stage.  Slot variant (C_out = 32 with several K chunks per row group): a group's rows complete together,
  // and row a + G + k of the next group needs the slot of row a + G + k - 5, whose epilogue waits for row a + G + k - 3: that
  // row must belong to an earlier group, so G <= 3.  The ring variant (one-chunk layers, C_out = 64) has 16 / 8 groups of
  // slack and takes four rows per barrier round trip.
  const bool ring = !derive && (C_out == 64 || p.nchunk == 1);
  const int want_stages = derive ? 5 : 3;
  int G = ring ? 4 : 3;
  for (; G >= 2; --G) {
    p.stage_bytes = G * 4 * kRwRowBytes;
    p.stages = budget / p.stage_bytes;
    if (p.stages >= want_stages) break;
  }
  UNCL_REQUIRE(G >= 2 && p.stages >= want_stages, "%s: filter bank of C_in=%d does not fit shared memory next to the stage ring", what, C_in);
  if (p.stages > kRwMaxStages) p.stages = kRwMaxStages;
  p.G = G;
  p.R = rw_pick_R(N, p.nbands, p.Ho, sms);
  // Ring variant with several K chunks: a group accumulates (row, chunk) contributions in the order the stages deliver them,
  // chunk-major inside a row group.  Strips that start on a multiple of G keep the row groups at the same absolute rows
  // whatever the strip height, so the summation order - and with it every output bit - does not depend on the batch size.
  if (ring && p.nchunk > 1) p.R = ceil_div(p.R, G) * G;
  p.nstrips = ceil_div(p.Ho, p.R);
  p.items_per_img = p.nbands * p.nstrips;
  UNCL_REQUIRE((long)N * p.items_per_img < (1 << 24), "%s: too many strips", what);
  p.num_items = N * p.items_per_img;
  if (derive) {
    const int cs32 = C_in / 4 / 32;
    UNCL_REQUIRE(C_in % 128 == 0 && p.nchunk == 4 * cs32, "%s: fused skip operators need C_skip a multiple of 32", what);
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
  int smem_bytes = p.stages * p.stage_bytes + p.w_total + p.f_bytes + tail;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;   // one CTA per SM (each owns all 512 TMEM columns)
  *smem_bytes_out = smem_bytes;
  return UNCL_OK;
}

int rw_launch(const void* in, long in_img_stride, const void* w_rows, const float* bias, void* out, long out_img_stride, int N,
              int C_in, int H, int W, int C_out, int pad, int Wc, int act, int emit_skip, int fuse_outc, const float* outc_w,
              const float* outc_b, float* out_img, float* out_logit, int derive, const char* what, cudaStream_t stream,
              const float* fx = nullptr, long fx_img_stride = 0, const void* fw = nullptr) {
  const int front = fx != nullptr ? 1 : 0;   // `in` is unused then: H, W are the extent of the first conv's OUTPUT
  UNCL_REQUIRE(front || (in_img_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0), "%s: input must be 16-byte aligned", what);
  UNCL_REQUIRE((reinterpret_cast<uintptr_t>(w_rows) & 15) == 0 && (reinterpret_cast<uintptr_t>(fw) & 15) == 0,
               "%s: weights must be 16-byte aligned", what);
  RwParams p{};
  int smem_bytes = 0;
  const int sms = sm_count();
  if (int rc = rw_plan(p, N, C_in, H, W, C_out, pad, Wc, derive, sms, what, &smem_bytes, front)) return rc;
  p.fx = fx; p.fx_img_stride = fx_img_stride; p.fw = reinterpret_cast<const bf16*>(fw); p.H0 = H + 2; p.W0 = W + 2;
  p.w = reinterpret_cast<const bf16*>(w_rows);
  p.bias = bias; p.out = reinterpret_cast<bf16*>(out); p.out_img_stride = out_img_stride;
  p.outc_w = outc_w; p.outc_b = outc_b; p.out_img = out_img; p.out_logit = out_logit;
  p.act = act; p.emit_skip = emit_skip; p.fuse_outc = fuse_outc;
  p.ring = 1;
#ifdef UNCL_PROBES
  if (getenv("UNCL_RW_NORING")) p.ring = 0;
  if (const char* e = getenv("UNCL_RW_PROBE")) p.probe = atoi(e);
#endif
  CUtensorMap tmap{};   // front mode: no activation tensor exists, the map is never touched
  if (!front) {
    CUresult r = encode_blocked_bf16(&tmap, in, W, H, (derive ? C_in / 2 : C_in) / 8, N, in_img_stride, 128, p.G, 4);
    if (r != CUDA_SUCCESS) return uncl_set_error(UNCL_ECUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
  }
  const int grid = p.num_items < sms ? p.num_items : sms;
  // one instantiation per (fused skip operators, rows per stage): the issuing warp's descriptor offsets are immediates
  static thread_local int smem_ok[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, smem_dev[10] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
#define RW_LAUNCH(D, GG, RING, NTT, slot_) RW_LAUNCH5(D, GG, RING, NTT, false, slot_)
#define RW_LAUNCH5(D, GG, RING, NTT, FR, slot_)                                                                                   \
  do {                                                                                                                      \
    cudaError_t e = ensure_smem(conv3x3_tc_rows_kernel<D, GG, RING, NTT, FR>, smem_bytes, smem_ok[slot_], smem_dev[slot_]);   \
    if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "%s: smem attr: %s", what, cudaGetErrorString(e));               \
    conv3x3_tc_rows_kernel<D, GG, RING, NTT, FR><<<grid, (D || FR) ? kRwThreadsDerive : kRwThreads, smem_bytes, stream>>>(tmap, p); \
  } while (0)
  const bool ring = !derive && (C_out == 64 || (p.nchunk == 1 && p.ring));
  if (front) {
    UNCL_REQUIRE(p.G == 4 && p.nchunk == 1 && C_out == 32 && !derive, "%s: front mode needs the one-chunk C_out = 32 layer", what);
    RW_LAUNCH5(false, 4, true, 32, true, 9);
  } else if (derive) {
    if (p.G == 3) RW_LAUNCH(true, 3, false, 32, 0);
    else if (p.G == 2) RW_LAUNCH(true, 2, false, 32, 1);
    else return uncl_set_error(UNCL_EINVAL, "%s: no kernel for %d rows per stage", what, p.G);
  } else if (C_out == 64) {
    if (p.G == 4) RW_LAUNCH(false, 4, true, 64, 6);
    else if (p.G == 3) RW_LAUNCH(false, 3, true, 64, 7);
    else if (p.G == 2) RW_LAUNCH(false, 2, true, 64, 8);
    else return uncl_set_error(UNCL_EINVAL, "%s: no kernel for %d rows per stage", what, p.G);
  } else {
    if (p.G == 4 && ring) RW_LAUNCH(false, 4, true, 32, 5);
    else if (p.G == 4) RW_LAUNCH(false, 4, false, 32, 2);
    else if (p.G == 3) RW_LAUNCH(false, 3, false, 32, 3);
    else if (p.G == 2) RW_LAUNCH(false, 2, false, 32, 4);
    else return uncl_set_error(UNCL_EINVAL, "%s: no kernel for %d rows per stage", what, p.G);
  }
#undef RW_LAUNCH
#undef RW_LAUNCH5
  return uncl_check_launch(what);
}

// columns the row kernel computes: whole 126-column bands, unless the remainder is wide enough to fill a band usefully
int rw_cols(int Wo) {
  const int rem = Wo % 126;
  return (rem == 0 || rem >= 64 || Wo < 126) ? Wo : Wo - rem;
}

}  // namespace

// the older kernels on the output columns [x0, Wo) (conv_tc.cu)
int uncl_conv3x3_tc_cols(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                         long out_img_stride, int out_dtype, int N, int C_in, int H, int W, int C_out, int pad, int act,
                         int emit_skip, int fuse_outc, const float* outc_w, const float* outc_b, float* out_img,
                         float* out_logit, int x0, cudaStream_t stream);
int uncl_conv3x3_tc_skipcat_cols(const void* in, long in_img_stride, const void* w_packed, const float* bias, void* out,
                                 long out_img_stride, int out_dtype, int N, int C_skip, int H, int W, int C_out, int pad,
                                 int act, int x0, cudaStream_t stream);

// 3x3 conv / ConvTranspose 3x3 (pad 0 / 2) with C_out = 32 or 64, bf16 blocked in and out, bias + ReLU / identity, optional
// skip-plane emission and (C_out = 32) fused 1x1 out conv + sigmoid: the arguments of uncl_conv3x3_tc with two packed
// filter banks.  w_rows: packing.conv3x3_tc_rows; w_tail: packing.conv3x3_tc (used for the columns past the last whole
// 126-column band, may be NULL when uncl_conv3x3_tc_rows_plan reports none).
extern "C" int uncl_conv3x3_tc_rows(const void* in, long in_img_stride, const void* w_rows, const void* w_tail,
                                    const float* bias, void* out, long out_img_stride, int N, int C_in, int H, int W, int C_out,
                                    int pad, int act, int emit_skip, int fuse_outc, const float* outc_w, const float* outc_b,
                                    float* out_img, float* out_logit, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C_in > 0 && C_in % 32 == 0 && (C_out == 32 || C_out == 64) && (pad == 0 || pad == 2) && w_rows != nullptr,
               "conv3x3_tc_rows: unsupported C_in=%d C_out=%d pad=%d", C_in, C_out, pad);
  UNCL_REQUIRE(act == UNCL_ACT_RELU || act == UNCL_ACT_NONE, "conv3x3_tc_rows: only ReLU / identity epilogues are built");
  UNCL_REQUIRE(!fuse_outc || (C_out == 32 && outc_w && outc_b && out_img), "conv3x3_tc_rows: fuse_outc needs C_out == 32 and outc params");
  UNCL_REQUIRE(out != nullptr || fuse_outc, "conv3x3_tc_rows: no output requested");
  const int Wo = W + 2 * pad - 2;
  UNCL_REQUIRE(Wo > 0 && H + 2 * pad - 2 > 0, "conv3x3_tc_rows: empty output");
  const int Wc = rw_cols(Wo);
  if (int rc = rw_launch(in, in_img_stride, w_rows, bias, out, out_img_stride, N, C_in, H, W, C_out, pad, Wc, act, emit_skip,
                         fuse_outc, outc_w, outc_b, out_img, out_logit, 0, "conv3x3_tc_rows", stream))
    return rc;
  if (Wc < Wo) {
    UNCL_REQUIRE(w_tail != nullptr, "conv3x3_tc_rows: %d trailing columns need the older kernels' filter bank (w_tail)", Wo - Wc);
    return uncl_conv3x3_tc_cols(in, in_img_stride, w_tail, bias, out, out_img_stride, UNCL_BF16, N, C_in, H, W, C_out, pad, act,
                                emit_skip, fuse_outc, outc_w, outc_b, out_img, out_logit, Wc, stream);
  }
  return UNCL_OK;
}

// inc.conv + inc.conv1 of the generator in ONE launch (Conv2d 1 -> 32 3x3 + ReLU, Conv2d 32 -> 32 3x3 + ReLU:
// unet_parts.py:57-87 with in_ch = 1): the first conv's output is never written to memory - four extra warps compute it into
// the stage ring of the row kernel (front mode, see the kernel).  x: [N][H0][W0] fp32; fw = packing.conv_first_rows(w1, b1)
// (the bias rides on a constant-one tap); w_rows = packing.conv3x3_tc_rows(w9 of the second conv); out: bf16 blocked [N][4 (+ skip planes)][H0-4][W0-4][8].
// W0 - 4 must be a whole number of 126-column bands (252 for the 256-pixel tiles of the generator).
extern "C" int uncl_conv_first_conv3x3_tc_rows(const float* x, long x_img_stride, const void* fw, const void* w_rows,
                                               const float* bias, void* out, long out_img_stride, int N, int H0, int W0, int act,
                                               int emit_skip, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && x != nullptr && fw != nullptr && w_rows != nullptr && out != nullptr && H0 > 4 && W0 > 4,
               "conv_first_conv3x3_tc_rows: bad arguments (N=%d H0=%d W0=%d)", N, H0, W0);
  UNCL_REQUIRE(act == UNCL_ACT_RELU || act == UNCL_ACT_NONE, "conv_first_conv3x3_tc_rows: only ReLU / identity epilogues are built");
  const int Wo = W0 - 4;
  UNCL_REQUIRE(rw_cols(Wo) == Wo, "conv_first_conv3x3_tc_rows: %d output columns are not whole 126-column bands", Wo);
  return rw_launch(nullptr, 0, w_rows, bias, out, out_img_stride, N, 32, H0 - 2, W0 - 2, 32, 0, Wo, act, emit_skip, 0, nullptr, nullptr,
                   nullptr, nullptr, 0, "conv_first_conv3x3_tc_rows", stream, x, x_img_stride, fw);
}

// uncl_conv3x3_tc_skipcat (fused skip operators, `in` = [skip (C_skip) | up-sampled (C_skip)], C_out = 32) through the row
// kernel.  w_rows / w_tail: packing.conv3x3_tc_rows / packing.conv3x3_tc of the full [9][4*C_skip][32] filter bank.
extern "C" int uncl_conv3x3_tc_rows_skipcat(const void* in, long in_img_stride, const void* w_rows, const void* w_tail,
                                            const float* bias, void* out, long out_img_stride, int N, int C_skip, int H, int W,
                                            int pad, int act, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C_skip > 0 && C_skip % 32 == 0 && (pad == 0 || pad == 2) && out != nullptr && w_rows != nullptr,
               "conv3x3_tc_rows_skipcat: unsupported C_skip=%d pad=%d", C_skip, pad);
  UNCL_REQUIRE(act == UNCL_ACT_RELU || act == UNCL_ACT_NONE, "conv3x3_tc_rows_skipcat: only ReLU / identity epilogues are built");
  const int Wo = W + 2 * pad - 2;
  UNCL_REQUIRE(Wo > 0 && H + 2 * pad - 2 > 0, "conv3x3_tc_rows_skipcat: empty output");
  const int Wc = rw_cols(Wo);
  if (int rc = rw_launch(in, in_img_stride, w_rows, bias, out, out_img_stride, N, 4 * C_skip, H, W, 32, pad, Wc, act, 0, 0, nullptr,
                         nullptr, nullptr, nullptr, 1, "conv3x3_tc_rows_skipcat", stream))
    return rc;
  if (Wc < Wo) {
    UNCL_REQUIRE(w_tail != nullptr, "conv3x3_tc_rows_skipcat: %d trailing columns need the kx-merged filter bank (w_tail)", Wo - Wc);
    return uncl_conv3x3_tc_skipcat_cols(in, in_img_stride, w_tail, bias, out, out_img_stride, UNCL_BF16, N, C_skip, H, W, 32, pad,
                                        act, Wc, stream);
  }
  return UNCL_OK;
}

// Plan of the row kernel - pure host arithmetic, no GPU needed.  C_in is the logical channel count (4 * C_skip with
// derive).  plan[16]: 1 when the problem fits (0: filter bank too large for shared memory - use uncl_conv3x3_tc),
// columns computed by the row kernel, trailing columns left to the older kernels, bands, band width, strip height,
// strips per band, work items, rows per stage, K chunks per row group, pipeline stages, stage bytes, resident weight
// bytes, dynamic shared memory, SMs assumed, 1 for the ring variant of the accumulators.
extern "C" int uncl_conv3x3_tc_rows_plan(int N, int C_in, int H, int W, int C_out, int pad, int derive, int sms, int* plan) {
  UNCL_REQUIRE(plan != nullptr && N > 0 && C_in > 0 && C_in % 32 == 0 && (C_out == 32 || (C_out == 64 && !derive)) &&
                   (pad == 0 || pad == 2) && sms > 0 && H + 2 * pad - 2 > 0 && W + 2 * pad - 2 > 0,
               "conv3x3_tc_rows_plan: unsupported C_in=%d C_out=%d pad=%d H=%d W=%d", C_in, C_out, pad, H, W);
  for (int i = 0; i < 16; ++i) plan[i] = 0;
  const int Wo = W + 2 * pad - 2;
  const int nchunk = C_in / 32;
  const int tail = 128 + 128 + (2 * kRwMaxStages + 2 * kRwRing + 2) * 8 + 16 + 128 * 4 + 256;
  if (nchunk > kRwMaxProg || (derive && C_in % 128 != 0) ||
      227 * 1024 - tail - nchunk * kRwWChunk * (C_out / 32) < (derive ? 5 : 3) * 2 * 4 * kRwRowBytes)
    return UNCL_OK;   // plan[0] = 0: not eligible
  RwParams p{};
  int smem = 0;
  const int Wc = rw_cols(Wo);
  if (int rc = rw_plan(p, N, C_in, H, W, C_out, pad, Wc, derive, sms, "conv3x3_tc_rows_plan", &smem)) return rc;
  const int v[16] = {1, Wc, Wo - Wc, p.nbands, p.BW, p.R, p.nstrips, p.num_items, p.G, p.nchunk, p.stages, p.stage_bytes,
                     p.w_total, smem, sms, (!derive && (C_out == 64 || p.nchunk == 1)) ? 1 : 0};
  for (int i = 0; i < 16; ++i) plan[i] = v[i];
  return UNCL_OK;
}
