// Frame path around the generator: log-lambda HDR normalisation fused with the replicate pad, tile gather,
// closed-form cross-fade blend of the 256x256 tiles, on-device percentiles (radix select), post-process
// (clamp / stretch / back-to-colour / crop) and the final 8-bit stretch.  All HBM-bound, vectorised where the
// layout allows, no host synchronisation anywhere.
//
// Reference: utils/model_save_util.py:219-240, 242-263 (log-lambda normalise), :409-486, 488-565 (tiling + blend),
// :389-402, 589-606 (post-process); utils/hdr_image_util.py:76-82 (to_gray), :122-132 (back_to_color_tensor),
// :93-102, 237-245 (8-bit stretch); utils/data_loader_util.py:135-157, 175-179 (replicate pad).
#include <cstdlib>
#include "common.cuh"

namespace {

__device__ __forceinline__ float gray_of(float r, float g, float b) { return 0.299f * r + 0.587f * g + 0.114f * b; }

// ---- pass 1: per-block partial (min rgb, min Y, max Y) --------------------------------------------------------
__global__ void __launch_bounds__(256) frame_stats_kernel(const float* __restrict__ rgb, long HW, const float* shift_src,
                                                         float* __restrict__ partials) {
  __shared__ float red[3][8];
  // shift_src == nullptr: first pass (no shift).  Otherwise stats[0] holds min(rgb); if it is >= 0 nothing changes.
  float shift = 0.f;
  if (shift_src != nullptr) {
    shift = fminf(shift_src[0], 0.f);
    if (shift == 0.f) return;  // first-pass statistics stand
  }
  float mn = INFINITY, ymn = INFINITY, ymx = -INFINITY;
  const long n4 = HW / 4;
  const float4* r4 = reinterpret_cast<const float4*>(rgb);
  const float4* g4 = reinterpret_cast<const float4*>(rgb + HW);
  const float4* b4 = reinterpret_cast<const float4*>(rgb + 2 * HW);
  const bool vec = (HW % 4 == 0);
  if (vec) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
      const float4 r = __ldg(r4 + i), g = __ldg(g4 + i), b = __ldg(b4 + i);
      const float rr[4] = {r.x, r.y, r.z, r.w}, gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mn = fminf(mn, fminf(rr[k], fminf(gg[k], bb[k])));
        const float y = gray_of(rr[k] - shift, gg[k] - shift, bb[k] - shift);
        ymn = fminf(ymn, y);
        ymx = fmaxf(ymx, y);
      }
    }
  } else {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long)gridDim.x * blockDim.x) {
      const float r = rgb[i], g = rgb[HW + i], b = rgb[2 * HW + i];
      mn = fminf(mn, fminf(r, fminf(g, b)));
      const float y = gray_of(r - shift, g - shift, b - shift);
      ymn = fminf(ymn, y);
      ymx = fmaxf(ymx, y);
    }
  }
  mn = warp_min(mn); ymn = warp_min(ymn); ymx = warp_max(ymx);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = mn; red[1][wid] = ymn; red[2][wid] = ymx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, red[0][w]); ymn = fminf(ymn, red[1][w]); ymx = fmaxf(ymx, red[2][w]); }
    partials[blockIdx.x * 3 + 0] = mn;
    partials[blockIdx.x * 3 + 1] = ymn;
    partials[blockIdx.x * 3 + 2] = ymx;
  }
}

// reduce the per-block partials into stats[0..2] = (min rgb, min Y, max Y); on the second pass keeps min rgb
__global__ void frame_stats_final_kernel(const float* __restrict__ partials, int nparts, float* __restrict__ stats,
                                         int second_pass) {
  __shared__ float red[3][32];
  if (second_pass && fminf(stats[0], 0.f) == 0.f) return;
  float mn = INFINITY, ymn = INFINITY, ymx = -INFINITY;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    mn = fminf(mn, partials[3 * i]);
    ymn = fminf(ymn, partials[3 * i + 1]);
    ymx = fmaxf(ymx, partials[3 * i + 2]);
  }
  mn = warp_min(mn); ymn = warp_min(ymn); ymx = warp_max(ymx);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = mn; red[1][wid] = ymn; red[2][wid] = ymx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fminf(mn, red[0][w]); ymn = fminf(ymn, red[1][w]); ymx = fmaxf(ymx, red[2][w]); }
    if (!second_pass) stats[0] = mn;
    stats[1] = ymn;
    stats[2] = ymx;
  }
}

// ---- pass 2: gray_log = log10((Y - Ymin) / max(Y - Ymin) * f + 1) / max(...), written replicate-padded --------
__global__ void __launch_bounds__(256) frame_normalise_pad_kernel(const float* __restrict__ rgb, int H, int W,
                                                                 const float* __restrict__ stats, float f,
                                                                 float* __restrict__ out, int H1, int W1, int padT,
                                                                 int padL) {
  const float shift = fminf(stats[0], 0.f), ymin = stats[1];
  const float gmax = stats[2] - ymin;          // max of (gray - gray.min())
  const float lmax = log10f((gmax / gmax) * f + 1.f);  // max of the log image is attained at gray == gmax
  const long HW = (long)H * W;
  const long total = (long)H1 * W1;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x1 = i % W1, y1 = i / W1;
    const int x = min(max(x1 - padL, 0), W - 1), y = min(max(y1 - padT, 0), H - 1);
    const long s = (long)y * W + x;
    const float g = gray_of(__ldg(rgb + s) - shift, __ldg(rgb + HW + s) - shift, __ldg(rgb + 2 * HW + s) - shift) - ymin;
    out[i] = log10f((g / gmax) * f + 1.f) / lmax;
  }
}

// ---- tiles: gather [T][256][256] from the padded frame, blend back ---------------------------------------------
__global__ void __launch_bounds__(256) tiles_gather_kernel(const float* __restrict__ frame, int W1,
                                                          const int* __restrict__ origins, float* __restrict__ tiles) {
  const int t = blockIdx.y;
  const int oy = origins[2 * t], ox = origins[2 * t + 1];
  const float* src = frame + (long)oy * W1 + ox;
  float* dst = tiles + (long)t * 65536;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 65536; i += gridDim.x * blockDim.x)
    dst[i] = __ldg(src + (long)(i >> 8) * W1 + (i & 255));
}

// out(y,x) = sum_{a<K} sum_{b<K} wy[y][a] * wx[x][b] * tile[ty[y][a]][tx[x][b]](y - y0, x - x0)
// (weights/indices are the closed form of the reference's sequential cross-fade; unused slots have weight 0)
template <int KMAX>
__global__ void __launch_bounds__(256) tiles_blend_kernel(const float* __restrict__ tiles, const int* __restrict__ yidx,
                                                         const float* __restrict__ yw, const int* __restrict__ ystart,
                                                         const int* __restrict__ xidx, const float* __restrict__ xw,
                                                         const int* __restrict__ xstart, int TX, int K,
                                                         float* __restrict__ out, int H1, int W1) {
  // one CTA row per output row: the row's tile indices / weights are CTA-uniform, no per-pixel division
  const int y = blockIdx.y;
  float wa_[KMAX];
  int ty_[KMAX], ly_[KMAX];
#pragma unroll
  for (int a = 0; a < KMAX; ++a) {
    wa_[a] = a < K ? yw[K * y + a] : 0.f;
    ty_[a] = a < K ? yidx[K * y + a] : 0;
    ly_[a] = y - ystart[ty_[a]];
  }
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W1; x += gridDim.x * blockDim.x) {
    float wb_[KMAX];
    long off_[KMAX];
#pragma unroll
    for (int b = 0; b < KMAX; ++b) {
      wb_[b] = b < K ? xw[K * x + b] : 0.f;
      const int tx = b < K ? xidx[K * x + b] : 0;
      off_[b] = ((long)tx << 16) + (x - xstart[tx]);
    }
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < KMAX; ++a) {
      if (wa_[a] == 0.f) continue;
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < KMAX; ++b) {
        if (wb_[b] == 0.f) continue;
        row = fmaf(wb_[b], __ldg(tiles + ((long)(ty_[a] * TX) << 16) + off_[b] + (ly_[a] << 8)), row);
      }
      acc = fmaf(wa_[a], row, acc);
    }
    out[(long)y * W1 + x] = acc;
  }
}

// ---- radix select: up to 4 order statistics of a float array in 3 histogram passes (11 + 11 + 10 bits) --------
constexpr int kSelQ = 4;
struct SelState {
  unsigned prefix[kSelQ];
  unsigned rank[kSelQ];
};
struct SelRanks {
  unsigned r[kSelQ];
};
__device__ __forceinline__ unsigned order_key(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_value(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// one warp per query: lanes own contiguous slices of the histogram, warp prefix sum locates the bin holding `rank`.
// Runs in the first 4 warps of the LAST histogram CTA to finish (see select_pass_kernel).
template <int PASS>
__device__ __forceinline__ void select_scan(SelState* st, const unsigned* hist, const SelRanks& ranks_in) {
  constexpr int BITS = PASS == 2 ? 10 : 11;
  constexpr int PER = (1 << BITS) / 32;
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned rank = PASS == 0 ? ranks_in.r[q] : st->rank[q];
  const unsigned pre = PASS == 0 ? 0u : st->prefix[q];
  int slot = 0;
  if (PASS != 0) {
    slot = q;
    for (int q2 = q - 1; q2 >= 0; --q2)
      if (st->prefix[q2] == pre) slot = q2;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");  // every warp has read the old prefixes / ranks before anyone overwrites them
  const unsigned* h = hist + ((size_t)slot << BITS) + lane * PER;
  unsigned mine = 0;
  for (int i = 0; i < PER; ++i) mine += h[i];
  unsigned incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned excl = incl - mine;
  // the owner lane is the first one whose inclusive count exceeds rank (the last lane if counts fall short)
  const unsigned ballot = __ballot_sync(0xffffffffu, rank < incl);
  const int owner = ballot ? __ffs(ballot) - 1 : 31;
  if (lane == owner) {
    unsigned cum = excl;
    int b = 0;
    for (; b < PER - 1; ++b) {
      const unsigned c = h[b];
      if (rank < cum + c) break;
      cum += c;
    }
    st->rank[q] = rank - cum;
    st->prefix[q] = (pre << BITS) | (unsigned)(lane * PER + b);
  }
}

// One radix pass: per-CTA shared histograms of the digit (restricted to the prefixes found so far), merged with global
// atomics; the last CTA to arrive scans the merged histogram for the 4 ranks, clears it for the next pass and, after the
// third pass, interpolates the two percentiles (numpy 'linear').  `done` and `hist` must be zero on entry of pass 0.
template <int PASS>
__global__ void __launch_bounds__(1024) select_pass_kernel(const float* __restrict__ data, long n, float clamp_lo,
                                                         float clamp_hi, SelState* st, unsigned* __restrict__ hist,
                                                         unsigned* done, const SelRanks ranks, double t0, double t1,
                                                         float* out) {
  constexpr int BITS = PASS == 2 ? 10 : 11;
  constexpr int SHIFT = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0);
  constexpr int NQ = PASS == 0 ? 1 : kSelQ;
  __shared__ unsigned s_hist[NQ << BITS];
  __shared__ bool s_last;
  for (int i = threadIdx.x; i < (NQ << BITS); i += blockDim.x) s_hist[i] = 0;
  unsigned pre[kSelQ];
#pragma unroll
  for (int q = 0; q < kSelQ; ++q) pre[q] = PASS == 0 ? 0u : st->prefix[q];
  __syncthreads();
  auto visit = [&](float raw) {
    const float v = fminf(fmaxf(raw, clamp_lo), clamp_hi);
    const unsigned k = order_key(v);
    if constexpr (PASS == 0) {
      // tone-mapped values cluster in a few exponent bins: aggregate equal digits inside the warp, one atomic per group
      const unsigned bin = k >> SHIFT;
      const unsigned active = __activemask();
      const unsigned peers = __match_any_sync(active, bin);
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[bin], (unsigned)__popc(peers));
    } else {
      const unsigned hi = k >> (SHIFT + BITS), bin = (k >> SHIFT) & ((1u << BITS) - 1);
#pragma unroll
      for (int q = 0; q < kSelQ; ++q) {
        // identical prefixes share a histogram slot (the lowest q) - the scan reads it from there
        bool first = true;
#pragma unroll
        for (int q2 = 0; q2 < q; ++q2) first = first && (pre[q2] != pre[q]);
        if (first && hi == pre[q]) atomicAdd(&s_hist[(q << BITS) + bin], 1u);
      }
    }
  };
  // 128-bit loads, two in flight per thread: the pass is a pure stream and would otherwise be latency-bound
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(data) & 15) == 0) {
    const float4* d4 = reinterpret_cast<const float4*>(data);
    const long n4 = n >> 2;
    long i = tid;
    for (; i + nthreads < n4; i += 2 * nthreads) {
      const float4 a = __ldg(d4 + i), b = __ldg(d4 + i + nthreads);
      visit(a.x); visit(a.y); visit(a.z); visit(a.w);
      visit(b.x); visit(b.y); visit(b.z); visit(b.w);
    }
    if (i < n4) {
      const float4 a = __ldg(d4 + i);
      visit(a.x); visit(a.y); visit(a.z); visit(a.w);
    }
    for (long j = (n4 << 2) + tid; j < n; j += nthreads) visit(__ldg(data + j));
  } else {
    for (long i = tid; i < n; i += nthreads) visit(__ldg(data + i));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (NQ << BITS); i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // the merged histogram comes back into shared memory in one parallel sweep (L2 reads, bypassing L1); the scan's
  // dependent look-ups then run at shared-memory latency
  for (int i = threadIdx.x; i < (NQ << BITS); i += blockDim.x) s_hist[i] = __ldcg(hist + i);
  __syncthreads();
  if (threadIdx.x < 32 * kSelQ) select_scan<PASS>(st, s_hist, ranks);
  __syncthreads();
  for (int i = threadIdx.x; i < (kSelQ << 11); i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) {
    *done = 0;
    if (PASS == 2) {
      const double a0 = key_value(st->prefix[0]), b0 = key_value(st->prefix[1]);
      const double a1 = key_value(st->prefix[2]), b1 = key_value(st->prefix[3]);
      out[0] = (float)(t0 < 0.5 ? a0 + (b0 - a0) * t0 : b0 - (b0 - a0) * (1.0 - t0));
      out[1] = (float)(t1 < 0.5 ? a1 + (b1 - a1) * t1 : b1 - (b1 - a1) * (1.0 - t1));
    }
  }
}


// ---- fused select: the three radix passes in ONE cooperative kernel, keys resident in shared memory ----------
// One CTA per SM owns a contiguous slice of the data and keeps its order keys in shared memory (<= 46 K keys per CTA:
// planes up to ~6.8 M values, i.e. both percentile calls of a 1080p frame); passes 1 and 2 then never touch global
// memory again, and the passes are separated by a grid barrier instead of a kernel boundary plus a last-CTA tail.
// Every CTA scans the merged histogram itself (identical result everywhere), so one barrier per pass is enough; the
// three passes use three separate, pre-zeroed global histograms.  Same digits, ranks and interpolation as
// select_pass_kernel: the results are bit-identical.
constexpr int kFusedThreads = 1024;
constexpr int kFusedMaxKeys = 46 * 1024;

__device__ __forceinline__ void select_grid_barrier(unsigned* ctr, unsigned target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(ctr, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// scan of one pass on shared-memory state (s_pre / s_rank are this CTA's copies of SelState), by the whole CTA:
// 256 threads per query, 8 (pass 2: 4) consecutive bins per thread, warp scan + 8 warp totals.  Same owner rule as
// select_scan: the first position whose inclusive count exceeds the rank, the last bin if the counts fall short.
template <int PASS>
__device__ __forceinline__ void select_scan_block(unsigned* s_pre, unsigned* s_rank, const unsigned* hist,
                                                  const SelRanks& ranks_in, unsigned* s_wtot /* [4][8] */) {
  constexpr int BITS = PASS == 2 ? 10 : 11;
  constexpr int PER = (1 << BITS) / 256;
  const int q = threadIdx.x >> 8, t = threadIdx.x & 255, lane = threadIdx.x & 31, w = t >> 5;
  const unsigned rank = PASS == 0 ? ranks_in.r[q] : s_rank[q];
  const unsigned pre = PASS == 0 ? 0u : s_pre[q];
  int slot = 0;
  if (PASS != 0) {
    slot = q;
    for (int q2 = q - 1; q2 >= 0; --q2)
      if (s_pre[q2] == pre) slot = q2;
  }
  const unsigned* h = hist + ((size_t)slot << BITS) + t * PER;
  unsigned c[PER], mine = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) { c[i] = h[i]; mine += c[i]; }
  unsigned incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_wtot[q * 8 + w] = incl;
  __syncthreads();   // also: every thread has read the old prefixes / ranks before anyone overwrites them
  unsigned base = 0;
  for (int i = 0; i < w; ++i) base += s_wtot[q * 8 + i];
  incl += base;
  const unsigned excl = incl - mine;
  // owner: rank in [excl, incl), or the very last thread of the query when the counts fall short
  const bool owner = (rank >= excl && rank < incl) || (t == 255 && rank >= incl);
  if (owner) {
    unsigned cum = excl;
    int b = 0;
#pragma unroll
    for (; b < PER - 1; ++b) {
      if (rank < cum + c[b]) break;
      cum += c[b];
    }
    s_rank[q] = rank - cum;
    s_pre[q] = (pre << BITS) | (unsigned)(t * PER + b);
  }
}

// ---- pieces of the fused select, shared by the plain kernel and by the frame-stage kernels that produce their values on
// the fly (blend -> select, post-process -> select -> 8-bit image) ----------------------------------------------------
struct SelSmem {
  unsigned* hist;    // [4][2048]
  unsigned* keys;    // this CTA's order keys
  unsigned* pre;     // [4]
  unsigned* rank;    // [4]
  unsigned* wtot;    // [4 * 8]
};

// pass 0 visit of one (already clamped) value: remember its order key, count its top 11 bits.  Warp-aggregated atomics;
// may be called from a partially active warp.
__device__ __forceinline__ void sel_visit0(const SelSmem& sm, float clamped, int idx) {
  const unsigned k = order_key(clamped);
  sm.keys[idx] = k;
  const unsigned bin = k >> 21;
  const unsigned peers = __match_any_sync(__activemask(), bin);
  if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sm.hist[bin], (unsigned)__popc(peers));
}

__device__ __forceinline__ void sel_clear_hist(const SelSmem& sm) {
  for (int i = threadIdx.x; i < (kSelQ << 11); i += blockDim.x) sm.hist[i] = 0;
  __syncthreads();
}

// after every thread's sel_visit0 calls: merge pass 0, then passes 1 and 2 from the shared-memory keys.  `bar_base` is the
// value of the grid-barrier counter when the kernel's select started (a kernel may run other grid barriers before).
// On return sm.pre[0..3] hold the keys of the four order statistics - in EVERY CTA.
__device__ __forceinline__ void sel_finish(const SelSmem& sm, int cnt, unsigned* __restrict__ hist3, unsigned* barrier,
                                           unsigned bar_base, const SelRanks& ranks) {
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x)
    if (sm.hist[i]) atomicAdd(&hist3[i], sm.hist[i]);
  select_grid_barrier(barrier, bar_base + gridDim.x);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm.hist[i] = __ldcg(hist3 + i);
  __syncthreads();
  select_scan_block<0>(sm.pre, sm.rank, sm.hist, ranks, sm.wtot);
  __syncthreads();
#pragma unroll
  for (int pass = 1; pass <= 2; ++pass) {
    const int BITS = pass == 2 ? 10 : 11, SHIFT = pass == 1 ? 10 : 0;
    unsigned pre[kSelQ];
    bool first[kSelQ];
#pragma unroll
    for (int q = 0; q < kSelQ; ++q) {
      pre[q] = sm.pre[q];
      first[q] = true;
#pragma unroll
      for (int q2 = 0; q2 < q; ++q2) first[q] = first[q] && (sm.pre[q2] != pre[q]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (kSelQ << 11); i += blockDim.x) sm.hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const unsigned k = sm.keys[i];
      const unsigned hi = k >> (SHIFT + BITS), bin = (k >> SHIFT) & ((1u << BITS) - 1);
#pragma unroll
      for (int q = 0; q < kSelQ; ++q)
        if (first[q] && hi == pre[q]) atomicAdd(&sm.hist[(q << BITS) + bin], 1u);
    }
    __syncthreads();
    unsigned* gh = hist3 + pass * (kSelQ << 11);
    for (int i = threadIdx.x; i < (kSelQ << BITS); i += blockDim.x)
      if (sm.hist[i]) atomicAdd(&gh[i], sm.hist[i]);
    select_grid_barrier(barrier, bar_base + gridDim.x * (unsigned)(pass + 1));
    for (int i = threadIdx.x; i < (kSelQ << BITS); i += blockDim.x) sm.hist[i] = __ldcg(gh + i);
    __syncthreads();
    if (pass == 1) select_scan_block<1>(sm.pre, sm.rank, sm.hist, ranks, sm.wtot);
    else select_scan_block<2>(sm.pre, sm.rank, sm.hist, ranks, sm.wtot);
    __syncthreads();
  }
}

// numpy 'linear' interpolation between the two neighbouring order statistics of each percentile
__device__ __forceinline__ void sel_result(const SelSmem& sm, double t0, double t1, float& lo, float& hi) {
  const double a0 = key_value(sm.pre[0]), b0 = key_value(sm.pre[1]);
  const double a1 = key_value(sm.pre[2]), b1 = key_value(sm.pre[3]);
  lo = (float)(t0 < 0.5 ? a0 + (b0 - a0) * t0 : b0 - (b0 - a0) * (1.0 - t0));
  hi = (float)(t1 < 0.5 ? a1 + (b1 - a1) * t1 : b1 - (b1 - a1) * (1.0 - t1));
}

#define UNCL_SEL_SMEM(sm)                                                          \
  extern __shared__ __align__(16) unsigned s_dyn[];                                \
  __shared__ unsigned s_pre_[kSelQ], s_rank_[kSelQ], s_wtot_[kSelQ * 8];           \
  const SelSmem sm = {s_dyn, s_dyn + (kSelQ << 11), s_pre_, s_rank_, s_wtot_}

__global__ void __launch_bounds__(kFusedThreads, 1)
select_fused_kernel(const float* __restrict__ data, long n, long per_cta, float clamp_lo, float clamp_hi,
                    unsigned* __restrict__ hist3, unsigned* barrier, const SelRanks ranks, double t0, double t1,
                    float* out) {
  UNCL_SEL_SMEM(sm);
  const long begin = (long)blockIdx.x * per_cta;
  const int cnt = (int)max(0L, min(per_cta, n - begin));
  sel_clear_hist(sm);
  // ---- load + pass 0 (top 11 bits)
  {
    auto visit0 = [&](float raw, int idx) { sel_visit0(sm, fminf(fmaxf(raw, clamp_lo), clamp_hi), idx); };
    const float* d = data + begin;   // begin is a multiple of 4 and data is 16-byte aligned (checked by the host)
    const int c4 = cnt >> 2;
    const float4* d4 = reinterpret_cast<const float4*>(d);
    int i = threadIdx.x;
    for (; i + kFusedThreads < c4; i += 2 * kFusedThreads) {
      const float4 a = __ldg(d4 + i), b = __ldg(d4 + i + kFusedThreads);
      visit0(a.x, 4 * i); visit0(a.y, 4 * i + 1); visit0(a.z, 4 * i + 2); visit0(a.w, 4 * i + 3);
      const int j = i + kFusedThreads;
      visit0(b.x, 4 * j); visit0(b.y, 4 * j + 1); visit0(b.z, 4 * j + 2); visit0(b.w, 4 * j + 3);
    }
    if (i < c4) {
      const float4 a = __ldg(d4 + i);
      visit0(a.x, 4 * i); visit0(a.y, 4 * i + 1); visit0(a.z, 4 * i + 2); visit0(a.w, 4 * i + 3);
    }
    for (int t = (c4 << 2) + threadIdx.x; t < cnt; t += kFusedThreads) visit0(__ldg(d + t), t);
  }
  sel_finish(sm, cnt, hist3, barrier, 0u, ranks);
  if (blockIdx.x == 0 && threadIdx.x == 0) sel_result(sm, t0, t1, out[0], out[1]);
}

// ---- frame stage 1 in ONE cooperative launch: statistics -> (shifted statistics) -> log-lambda normalisation written
// straight into the generator's 256 x 256 tiles (the replicate-padded frame is never materialised).  Same arithmetic as
// frame_stats_kernel / frame_normalise_pad_kernel / tiles_gather_kernel: the tiles are bit-identical to the staged path.
__device__ __forceinline__ void cta_minmax3(float& mn, float& ymn, float& ymx, float (*red)[32]) {
  mn = warp_min(mn); ymn = warp_min(ymn); ymx = warp_max(ymx);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) { red[0][wid] = mn; red[1][wid] = ymn; red[2][wid] = ymx; }
  __syncthreads();
  mn = red[0][0]; ymn = red[1][0]; ymx = red[2][0];
  for (int w = 1; w < nw; ++w) { mn = fminf(mn, red[0][w]); ymn = fminf(ymn, red[1][w]); ymx = fmaxf(ymx, red[2][w]); }
}

__device__ __forceinline__ void frame_stats_slice(const float* __restrict__ rgb, long HW, float shift, float& mn, float& ymn,
                                                  float& ymx) {
  mn = INFINITY; ymn = INFINITY; ymx = -INFINITY;
  const long n4 = HW / 4;
  if (HW % 4 == 0) {
    const float4* r4 = reinterpret_cast<const float4*>(rgb);
    const float4* g4 = reinterpret_cast<const float4*>(rgb + HW);
    const float4* b4 = reinterpret_cast<const float4*>(rgb + 2 * HW);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
      const float4 r = __ldg(r4 + i), g = __ldg(g4 + i), b = __ldg(b4 + i);
      const float rr[4] = {r.x, r.y, r.z, r.w}, gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mn = fminf(mn, fminf(rr[k], fminf(gg[k], bb[k])));
        const float y = gray_of(rr[k] - shift, gg[k] - shift, bb[k] - shift);
        ymn = fminf(ymn, y);
        ymx = fmaxf(ymx, y);
      }
    }
  } else {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long)gridDim.x * blockDim.x) {
      const float r = rgb[i], g = rgb[HW + i], b = rgb[2 * HW + i];
      mn = fminf(mn, fminf(r, fminf(g, b)));
      const float y = gray_of(r - shift, g - shift, b - shift);
      ymn = fminf(ymn, y);
      ymx = fmaxf(ymx, y);
    }
  }
}

__global__ void __launch_bounds__(kFusedThreads, 1)
frame_normalise_tiles_kernel(const float* __restrict__ rgb, int H, int W, float f, int H1, int W1,
                             const int* __restrict__ origins, int T, float* __restrict__ tiles, float* __restrict__ stats_out,
                             float* __restrict__ partials, unsigned* barrier) {
  __shared__ float red[3][32];
  const long HW = (long)H * W;
  // ---- statistics: per-CTA partials, grid barrier, every CTA reduces the (<= 148) partials itself
  float mn, ymn, ymx;
  frame_stats_slice(rgb, HW, 0.f, mn, ymn, ymx);
  cta_minmax3(mn, ymn, ymx, red);
  if (threadIdx.x == 0) { partials[3 * blockIdx.x] = mn; partials[3 * blockIdx.x + 1] = ymn; partials[3 * blockIdx.x + 2] = ymx; }
  select_grid_barrier(barrier, gridDim.x);
  mn = INFINITY; ymn = INFINITY; ymx = -INFINITY;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    mn = fminf(mn, __ldcg(partials + 3 * i)); ymn = fminf(ymn, __ldcg(partials + 3 * i + 1)); ymx = fmaxf(ymx, __ldcg(partials + 3 * i + 2));
  }
  cta_minmax3(mn, ymn, ymx, red);
  const float shift = fminf(mn, 0.f);
  if (shift != 0.f) {   // negative inputs (exr): luminance range of the shifted image (grid-uniform branch)
    float m2, y0, y1;
    frame_stats_slice(rgb, HW, shift, m2, y0, y1);
    cta_minmax3(m2, y0, y1, red);
    float* p2 = partials + 3 * gridDim.x;
    if (threadIdx.x == 0) { p2[3 * blockIdx.x + 1] = y0; p2[3 * blockIdx.x + 2] = y1; }
    select_grid_barrier(barrier, 2 * gridDim.x);
    ymn = INFINITY; ymx = -INFINITY;
    float dummy = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      ymn = fminf(ymn, __ldcg(p2 + 3 * i + 1)); ymx = fmaxf(ymx, __ldcg(p2 + 3 * i + 2));
    }
    cta_minmax3(dummy, ymn, ymx, red);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { stats_out[0] = mn; stats_out[1] = ymn; stats_out[2] = ymx; }
  // ---- normalise, written as tiles: tile pixel (t, ly, lx) = padded pixel (oy + ly, ox + lx) = source pixel clamped
  const float gmax = ymx - ymn;
  const float lmax = log10f((gmax / gmax) * f + 1.f);
  const int padT = (H1 - H) / 2, padL = (W1 - W) / 2;
  const long total4 = (long)T * (65536 / 4);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const int t = (int)(i >> 14), r = (int)(i & 16383);
    const int ly = r >> 6, lx = (r & 63) << 2;
    const int y = min(max(__ldg(origins + 2 * t) + ly - padT, 0), H - 1);
    const int x1 = __ldg(origins + 2 * t + 1) + lx - padL;
    const float* row = rgb + (long)y * W;
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = min(max(x1 + k, 0), W - 1);
      const float g = gray_of(__ldg(row + x) - shift, __ldg(row + HW + x) - shift, __ldg(row + 2 * HW + x) - shift) - ymn;
      o[k] = log10f((g / gmax) * f + 1.f) / lmax;
    }
    *reinterpret_cast<float4*>(tiles + (i << 2)) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---- frame stage 2 in ONE cooperative launch: closed-form blend of the tiles -> padded plane (written out for the
// post-process) with its order keys kept in shared memory -> percentiles of the plane.  K <= 3 (overlap <= 3 tiles).
__global__ void __launch_bounds__(kFusedThreads, 1)
blend_select_kernel(const float* __restrict__ tiles, const int* __restrict__ yidx, const float* __restrict__ yw,
                    const int* __restrict__ ystart, const int* __restrict__ xidx, const float* __restrict__ xw,
                    const int* __restrict__ xstart, int TX, int K, float* __restrict__ plane, int H1, int W1, long per_cta,
                    unsigned* __restrict__ hist3, unsigned* barrier, const SelRanks ranks, double t0, double t1,
                    float* pct_out) {
  UNCL_SEL_SMEM(sm);
  const long n = (long)H1 * W1;
  const long begin = (long)blockIdx.x * per_cta;
  const int cnt = (int)max(0L, min(per_cta, n - begin));
  sel_clear_hist(sm);
  // four consecutive pixels of a row per thread and step (W1, per_cta and hence `begin` are multiples of 4): the row's
  // tables are read once, the 4 x up-to-9 tile loads are independent of each other, the plane is written as float4
  const int W4 = W1 >> 2, cnt4 = cnt >> 2;
  const int g0 = (int)(begin >> 2);
  for (int j = threadIdx.x; j < cnt4; j += blockDim.x) {
    const int g = g0 + j;
    const int y = g / W4, x0 = (g - y * W4) << 2;
    float wa[3];
    long rowoff[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      wa[a] = a < K ? __ldg(yw + K * y + a) : 0.f;
      const int ty = a < K ? __ldg(yidx + K * y + a) : 0;
      rowoff[a] = ((long)(ty * TX) << 16) + ((y - __ldg(ystart + ty)) << 8);
    }
    float acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k;
      float wb[3];
      long off[3];
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        wb[b] = b < K ? __ldg(xw + K * x + b) : 0.f;
        const int tx = b < K ? __ldg(xidx + K * x + b) : 0;
        off[b] = ((long)tx << 16) + (x - __ldg(xstart + tx));
      }
      float t = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (wa[a] == 0.f) continue;
        float row = 0.f;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          if (wb[b] == 0.f) continue;
          row = fmaf(wb[b], __ldg(tiles + rowoff[a] + off[b]), row);
        }
        t = fmaf(wa[a], row, t);
      }
      acc[k] = t;
    }
    *reinterpret_cast<float4*>(plane + ((long)g << 2)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) sel_visit0(sm, acc[k], 4 * j + k);
  }
  sel_finish(sm, cnt, hist3, barrier, 0u, ranks);
  if (blockIdx.x == 0 && threadIdx.x == 0) sel_result(sm, t0, t1, pct_out[0], pct_out[1]);
}

// ---- frame stage 3 in ONE cooperative launch: post-process (clamp to the plane's percentiles, stretch, back to colour,
// crop) -> colour values as order keys in shared memory (and, if asked, in `col_out`) -> their percentiles after the
// clamp to [0, 1] -> the 8-bit image straight from the shared-memory keys.  A CTA owns a contiguous range of PIXELS and
// keeps their three channels side by side, so its slice of the interleaved 8-bit image is contiguous too.
__global__ void __launch_bounds__(kFusedThreads, 1)
post_select_u8_kernel(const float* __restrict__ fake, int W1, int padT, int padL, const float* __restrict__ rgb, int H, int W,
                      const float* __restrict__ stats, const float* __restrict__ pct_plane, float* __restrict__ col_out,
                      long px_per_cta, unsigned* __restrict__ hist3, unsigned* barrier, const SelRanks ranks, double t0,
                      double t1, float* pct_out, unsigned char* __restrict__ u8_out) {
  UNCL_SEL_SMEM(sm);
  const long HW = (long)H * W;
  const long pbegin = (long)blockIdx.x * px_per_cta;
  const int pcnt = (int)max(0L, min(px_per_cta, HW - pbegin));
  sel_clear_hist(sm);
  {
    const float shift = fminf(stats[0], 0.f);
    const float lo = pct_plane[0], hi = pct_plane[1];
    const float inv = 1.f / (hi - lo);
    // two pixels per thread and step: their eight loads are issued before any arithmetic
    for (int j = threadIdx.x; j < pcnt; j += 2 * blockDim.x) {
      const int j1 = j + blockDim.x;
      const bool has1 = j1 < pcnt;
      float fk[2], r[2], g[2], b[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long i = pbegin + (u == 0 || has1 ? (u == 0 ? j : j1) : j);
        const int y = (int)(i / W), x = (int)(i - (long)y * W);
        fk[u] = __ldg(fake + (long)(y + padT) * W1 + (x + padL));
        r[u] = __ldg(rgb + i); g[u] = __ldg(rgb + HW + i); b[u] = __ldg(rgb + 2 * HW + i);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !has1) break;
        const int jj = u == 0 ? j : j1;
        const long i = pbegin + jj;
        const float s = (fminf(fmaxf(fk[u], lo), hi) - lo) * inv;
        const float rr = r[u] - shift, gg = g[u] - shift, bb = b[u] - shift;
        const float d = gray_of(rr, gg, bb) + 1e-8f;
        const float c0 = sqrtf(rr / d) * s, c1 = sqrtf(gg / d) * s, c2 = sqrtf(bb / d) * s;
        if (col_out != nullptr) { col_out[i] = c0; col_out[HW + i] = c1; col_out[2 * HW + i] = c2; }
        sel_visit0(sm, fminf(fmaxf(c0, 0.f), 1.f), 3 * jj);
        sel_visit0(sm, fminf(fmaxf(c1, 0.f), 1.f), 3 * jj + 1);
        sel_visit0(sm, fminf(fmaxf(c2, 0.f), 1.f), 3 * jj + 2);
      }
    }
  }
  sel_finish(sm, 3 * pcnt, hist3, barrier, 0u, ranks);
  float lo, hi;
  sel_result(sm, t0, t1, lo, hi);
  if (blockIdx.x == 0 && threadIdx.x == 0) { pct_out[0] = lo; pct_out[1] = hi; }
  // 8-bit image: clip((clamp(c,0,1) - lo) / (hi - lo), 0, 1) * 255 truncated - the keys ARE the clamped values
  unsigned char* o = u8_out + 3 * pbegin;
  for (int j = threadIdx.x; j < 3 * pcnt; j += blockDim.x) {
    float v = key_value(sm.keys[j]);
    v = fminf(fmaxf((v - lo) / (hi - lo), 0.f), 1.f);
    o[j] = (unsigned char)(v * 255.f);
  }
}

// ---- post-process: clamp to percentiles, stretch, back to colour, crop --------------------------------------
__global__ void __launch_bounds__(256) frame_postprocess_kernel(const float* __restrict__ fake, int W1, int padT,
                                                               int padL, const float* __restrict__ rgb, int H, int W,
                                                               const float* __restrict__ stats,
                                                               const float* __restrict__ pct, float* __restrict__ out) {
  const float shift = fminf(stats[0], 0.f);
  const float lo = pct[0], hi = pct[1];
  const float inv = 1.f / (hi - lo);
  const long HW = (long)H * W;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long)gridDim.x * blockDim.x) {
    const int x = i % W, y = i / W;
    const float f = __ldg(fake + (long)(y + padT) * W1 + (x + padL));
    const float s = (fminf(fmaxf(f, lo), hi) - lo) * inv;
    const float r = __ldg(rgb + i) - shift, g = __ldg(rgb + HW + i) - shift, b = __ldg(rgb + 2 * HW + i) - shift;
    const float d = gray_of(r, g, b) + 1e-8f;
    out[i] = sqrtf(r / d) * s;
    out[HW + i] = sqrtf(g / d) * s;
    out[2 * HW + i] = sqrtf(b / d) * s;
  }
}

// final 8-bit image, HWC: clip((clamp(c,0,1) - lo) / (hi - lo), 0, 1) * 255 truncated
__global__ void __launch_bounds__(256) frame_to_u8_kernel(const float* __restrict__ col, long HW,
                                                         const float* __restrict__ pct, unsigned char* __restrict__ out) {
  const float lo = pct[0], hi = pct[1];
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = fminf(fmaxf(__ldg(col + c * HW + i), 0.f), 1.f);
      v = fminf(fmaxf((v - lo) / (hi - lo), 0.f), 1.f);
      out[3 * i + c] = (unsigned char)(v * 255.f);
    }
  }
}

inline int grid_for(long total, int block, int per_sm) {
  long g = (total + block - 1) / block;
  const long cap = 148L * per_sm;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
constexpr int kStatBlocks = 148 * 8;

extern "C" long uncl_frame_workspace_bytes() {
  // partials [kStatBlocks][3] floats | stats[4] | pct[4] | SelState | ranks[4] | hist [4][2048] | fused: barrier[4] + hist [3][4][2048]
  return (long)kStatBlocks * 3 * 4 + 16 + 16 + (long)sizeof(SelState) + 16 + 4L * 2048 * 4 + 16 + 3L * 4 * 2048 * 4 + 256;
}

struct FrameWs {
  float* partials; float* stats; float* pct; SelState* sel; unsigned* ranks; unsigned* hist;
  unsigned* fused_barrier; unsigned* fused_hist;
};
static FrameWs carve(void* ws) {
  FrameWs w;
  char* p = reinterpret_cast<char*>(ws);
  w.partials = reinterpret_cast<float*>(p); p += (size_t)kStatBlocks * 3 * 4;
  w.stats = reinterpret_cast<float*>(p); p += 16;
  w.pct = reinterpret_cast<float*>(p); p += 16;
  w.sel = reinterpret_cast<SelState*>(p); p += sizeof(SelState);
  w.ranks = reinterpret_cast<unsigned*>(p); p += 16;
  w.hist = reinterpret_cast<unsigned*>(p); p += 4 * 2048 * 4;
  w.fused_barrier = reinterpret_cast<unsigned*>(p); p += 16;
  w.fused_hist = reinterpret_cast<unsigned*>(p);
  return w;
}

extern "C" int uncl_frame_normalise_pad(const float* rgb, int H, int W, float f_factor, float* gray_out, int H1, int W1,
                                        float* stats_out, void* workspace, cudaStream_t stream) {
  UNCL_REQUIRE(H > 0 && W > 0 && H1 >= H && W1 >= W && f_factor > 0.f, "frame_normalise_pad: bad arguments");
  UNCL_REQUIRE((reinterpret_cast<uintptr_t>(rgb) & 15) == 0, "frame_normalise_pad: rgb must be 16-byte aligned");
  FrameWs w = carve(workspace);
  const long HW = (long)H * W;
  const int nb = grid_for(HW / 4 + 1, 256, 8);
  frame_stats_kernel<<<nb, 256, 0, stream>>>(rgb, HW, nullptr, w.partials);
  frame_stats_final_kernel<<<1, 1024, 0, stream>>>(w.partials, nb, stats_out, 0);
  // negative inputs (exr): shift by min(rgb) and recompute the luminance range; both kernels exit at once otherwise
  frame_stats_kernel<<<nb, 256, 0, stream>>>(rgb, HW, stats_out, w.partials);
  frame_stats_final_kernel<<<1, 1024, 0, stream>>>(w.partials, nb, stats_out, 1);
  const int padT = (H1 - H) / 2, padL = (W1 - W) / 2;
  frame_normalise_pad_kernel<<<grid_for((long)H1 * W1, 256, 8), 256, 0, stream>>>(rgb, H, W, stats_out, f_factor, gray_out, H1, W1, padT, padL);
  return uncl_check_launch("frame_normalise_pad");
}

extern "C" int uncl_tiles_gather(const float* frame, int H1, int W1, const int* origins, int T, float* tiles,
                                 cudaStream_t stream) {
  UNCL_REQUIRE(T > 0 && H1 >= 256 && W1 >= 256, "tiles_gather: bad arguments");
  tiles_gather_kernel<<<dim3(16, T), 256, 0, stream>>>(frame, W1, origins, tiles);
  return uncl_check_launch("tiles_gather");
}

extern "C" int uncl_tiles_blend(const float* tiles, const int* yidx, const float* yw, const int* ystart,
                                const int* xidx, const float* xw, const int* xstart, int TX, int K, float* out,
                                int H1, int W1, cudaStream_t stream) {
  UNCL_REQUIRE(TX > 0 && K > 0 && K <= 8 && H1 >= 256 && W1 >= 256, "tiles_blend: bad arguments (K=%d)", K);
  const dim3 grid((W1 + 255) / 256, H1);
  if (K <= 3) tiles_blend_kernel<3><<<grid, 256, 0, stream>>>(tiles, yidx, yw, ystart, xidx, xw, xstart, TX, K, out, H1, W1);
  else tiles_blend_kernel<8><<<grid, 256, 0, stream>>>(tiles, yidx, yw, ystart, xidx, xw, xstart, TX, K, out, H1, W1);
  return uncl_check_launch("tiles_blend");
}

// numpy.percentile(clamp(data, clamp_lo, clamp_hi), [p_lo, p_hi]) with the default 'linear' method -> pct_out[0..1]
extern "C" int uncl_percentile_pair(const float* data, long n, float clamp_lo, float clamp_hi, double p_lo,
                                    double p_hi, float* pct_out, void* workspace, cudaStream_t stream) {
  UNCL_REQUIRE(n >= 2 && p_lo >= 0 && p_hi <= 100 && p_lo <= p_hi, "percentile_pair: bad arguments");
  FrameWs w = carve(workspace);
  const double v0 = p_lo / 100.0 * (double)(n - 1), v1 = p_hi / 100.0 * (double)(n - 1);
  long k0 = (long)v0, k1 = (long)v1;
  if (k0 > n - 2) k0 = n - 2;
  if (k1 > n - 2) k1 = n - 2;
  UNCL_REQUIRE(n < (1L << 32), "percentile_pair: n too large");
  const SelRanks ranks = {{(unsigned)k0, (unsigned)(k0 + 1), (unsigned)k1, (unsigned)(k1 + 1)}};
  const double t0 = v0 - (double)k0, t1 = v1 - (double)k1;
  // fused path: one cooperative launch when every CTA's slice of keys fits in its shared memory
#ifdef UNCL_PROBES
  const bool force_3pass = getenv("UNCL_SELECT_3PASS") != nullptr;   // probe build only
#else
  const bool force_3pass = false;
#endif
  {
    int dev = 0, sms = 148, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    long per_cta = ((n + sms - 1) / sms + 3) & ~3L;
    if (coop && per_cta <= kFusedMaxKeys && (reinterpret_cast<uintptr_t>(data) & 15) == 0 && !force_3pass) {
      const size_t smem = (size_t)((kSelQ << 11) + per_cta) * sizeof(unsigned);
      cudaError_t e = cudaFuncSetAttribute(select_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "percentile_pair: smem attr: %s", cudaGetErrorString(e));
      cudaMemsetAsync(w.fused_barrier, 0, 16 + 3 * 4 * 2048 * 4, stream);
      unsigned* hist3 = w.fused_hist;
      unsigned* bar = w.fused_barrier;
      void* args[] = {(void*)&data, (void*)&n, (void*)&per_cta, (void*)&clamp_lo, (void*)&clamp_hi, (void*)&hist3, (void*)&bar,
                      (void*)&ranks, (void*)&t0, (void*)&t1, (void*)&pct_out};
      e = cudaLaunchCooperativeKernel((const void*)select_fused_kernel, dim3(sms), dim3(kFusedThreads), args, smem, stream);
      if (e == cudaSuccess) return uncl_check_launch("percentile_pair");
      // a device partition that cannot hold one CTA per SM at once (MPS / green-context limits) refuses the cooperative
      // launch synchronously: clear the error and take the three-launch path, which needs no co-residency
      (void)cudaGetLastError();
    }
  }
  // three-pass path (large planes).  One 1024-thread CTA per SM: every CTA merges up to 4 x 2048 shared-memory bins
  // into the global histogram with atomics
  const int nb = grid_for(n, 1024 * 8, 1);
  cudaMemsetAsync(w.ranks, 0, 16 + 4 * 2048 * 4, stream);   // arrival counter (w.ranks[0]) + histograms, contiguous
  select_pass_kernel<0><<<nb, 1024, 0, stream>>>(data, n, clamp_lo, clamp_hi, w.sel, w.hist, w.ranks, ranks, t0, t1, pct_out);
  select_pass_kernel<1><<<nb, 1024, 0, stream>>>(data, n, clamp_lo, clamp_hi, w.sel, w.hist, w.ranks, ranks, t0, t1, pct_out);
  select_pass_kernel<2><<<nb, 1024, 0, stream>>>(data, n, clamp_lo, clamp_hi, w.sel, w.hist, w.ranks, ranks, t0, t1, pct_out);
  return uncl_check_launch("percentile_pair");
}

extern "C" int uncl_frame_postprocess(const float* fake, int H1, int W1, const float* rgb, int H, int W,
                                      const float* stats, const float* pct, float* out, cudaStream_t stream) {
  UNCL_REQUIRE(H > 0 && W > 0 && H1 >= H && W1 >= W, "frame_postprocess: bad arguments");
  frame_postprocess_kernel<<<grid_for((long)H * W, 256, 8), 256, 0, stream>>>(fake, W1, (H1 - H) / 2, (W1 - W) / 2, rgb, H, W, stats, pct, out);
  return uncl_check_launch("frame_postprocess");
}

extern "C" int uncl_frame_to_u8(const float* col, int H, int W, const float* pct, unsigned char* out,
                                cudaStream_t stream) {
  UNCL_REQUIRE(H > 0 && W > 0, "frame_to_u8: bad arguments");
  frame_to_u8_kernel<<<grid_for((long)H * W, 256, 8), 256, 0, stream>>>(col, (long)H * W, pct, out);
  return uncl_check_launch("frame_to_u8");
}


// ================================================================================================
// Fused frame stages: one cooperative launch each (one 1024-thread CTA per SM, grid barriers inside)
// ================================================================================================
namespace {

struct SelPlan { SelRanks ranks; double t0, t1; };
// numpy.percentile 'linear': the two neighbouring order statistics of each percentile and the interpolation weights
inline SelPlan sel_plan(long n, double p_lo, double p_hi) {
  const double v0 = p_lo / 100.0 * (double)(n - 1), v1 = p_hi / 100.0 * (double)(n - 1);
  long k0 = (long)v0, k1 = (long)v1;
  if (k0 > n - 2) k0 = n - 2;
  if (k1 > n - 2) k1 = n - 2;
  SelPlan sp;
  sp.ranks = {{(unsigned)k0, (unsigned)(k0 + 1), (unsigned)k1, (unsigned)(k1 + 1)}};
  sp.t0 = v0 - (double)k0; sp.t1 = v1 - (double)k1;
  return sp;
}

struct CoopDev { int sms, coop; };
inline CoopDev coop_device() {
  static thread_local int cached_dev = -1;
  static thread_local CoopDev cd = {148, 0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cd.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&cd.coop, cudaDevAttrCooperativeLaunch, dev);
    cached_dev = dev;
  }
  return cd;
}

template <typename K>
inline int coop_launch(K kern, int sms, void** args, size_t smem, cudaStream_t stream, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024));
  if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "%s: smem attr: %s", what, cudaGetErrorString(e));
  e = cudaLaunchCooperativeKernel((const void*)kern, dim3(sms), dim3(kFusedThreads), args, smem, stream);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return uncl_set_error(UNCL_EUNSUPPORTED, "%s: cooperative launch refused (%s) - use the staged calls", what, cudaGetErrorString(e));
  }
  return uncl_check_launch(what);
}

}  // namespace

// 1 when the three fused frame-stage calls below can run for this geometry on the current device (cooperative launch
// available, every CTA's slice of order keys fits its shared memory, blend overlap of at most 3 tiles per axis), else 0:
// callers then use the staged calls (uncl_frame_normalise_pad ... uncl_frame_to_u8), which compute the same values.
extern "C" int uncl_frame_fused_supported(int H, int W, int H1, int W1, int K) {
  if (H <= 0 || W <= 0 || H1 < H || W1 < W || K <= 0 || K > 3) return 0;
  const CoopDev cd = coop_device();
  if (!cd.coop) return 0;
  const long per_plane = (((long)H1 * W1 + cd.sms - 1) / cd.sms + 3) & ~3L;
  const long px = ((long)H * W + cd.sms - 1) / cd.sms;
  return (per_plane <= kFusedMaxKeys && 3 * px <= kFusedMaxKeys && (long)H1 * W1 < (1L << 31)) ? 1 : 0;
}

// uncl_frame_normalise_pad + uncl_tiles_gather in one launch: statistics, (shifted statistics), normalisation written
// directly as the T 256x256 generator tiles at origins[t] = (y, x) of the padded (H1 x W1) frame, which is never stored.
extern "C" int uncl_frame_normalise_tiles(const float* rgb, int H, int W, float f_factor, int H1, int W1, const int* origins,
                                          int T, float* tiles, float* stats_out, void* workspace, cudaStream_t stream) {
  UNCL_REQUIRE(H > 0 && W > 0 && H1 >= H && W1 >= W && H1 >= 256 && W1 >= 256 && T > 0 && f_factor > 0.f,
               "frame_normalise_tiles: bad arguments");
  UNCL_REQUIRE((reinterpret_cast<uintptr_t>(rgb) & 15) == 0 && (reinterpret_cast<uintptr_t>(tiles) & 15) == 0,
               "frame_normalise_tiles: rgb / tiles must be 16-byte aligned");
  const CoopDev cd = coop_device();
  if (!cd.coop || 2 * cd.sms > kStatBlocks) return uncl_set_error(UNCL_EUNSUPPORTED, "frame_normalise_tiles: no cooperative launch");
  FrameWs w = carve(workspace);
  cudaMemsetAsync(w.fused_barrier, 0, 16, stream);
  float* partials = w.partials;
  unsigned* bar = w.fused_barrier;
  void* args[] = {(void*)&rgb, (void*)&H, (void*)&W, (void*)&f_factor, (void*)&H1, (void*)&W1, (void*)&origins, (void*)&T,
                  (void*)&tiles, (void*)&stats_out, (void*)&partials, (void*)&bar};
  return coop_launch(frame_normalise_tiles_kernel, cd.sms, args, 0, stream, "frame_normalise_tiles");
}

// uncl_tiles_blend + uncl_percentile_pair(plane, p_lo, p_hi) in one launch: the blended plane is written to `plane_out`
// and its percentiles to pct_out[0..1].  K <= 3.
extern "C" int uncl_frame_blend_percentiles(const float* tiles, const int* yidx, const float* yw, const int* ystart,
                                            const int* xidx, const float* xw, const int* xstart, int TX, int K,
                                            float* plane_out, int H1, int W1, double p_lo, double p_hi, float* pct_out,
                                            void* workspace, cudaStream_t stream) {
  UNCL_REQUIRE(TX > 0 && K > 0 && K <= 3 && H1 >= 256 && W1 >= 256 && p_lo >= 0 && p_hi <= 100 && p_lo <= p_hi,
               "frame_blend_percentiles: bad arguments (K=%d)", K);
  const CoopDev cd = coop_device();
  const long n = (long)H1 * W1;
  long per_cta = ((n + cd.sms - 1) / cd.sms + 3) & ~3L;
  if (!cd.coop || per_cta > kFusedMaxKeys || n >= (1L << 31))
    return uncl_set_error(UNCL_EUNSUPPORTED, "frame_blend_percentiles: plane too large for the fused kernel");
  FrameWs w = carve(workspace);
  const SelPlan sp = sel_plan(n, p_lo, p_hi);
  cudaMemsetAsync(w.fused_barrier, 0, 16 + 3 * 4 * 2048 * 4, stream);
  unsigned* hist3 = w.fused_hist;
  unsigned* bar = w.fused_barrier;
  const size_t smem = (size_t)((kSelQ << 11) + per_cta) * sizeof(unsigned);
  void* args[] = {(void*)&tiles, (void*)&yidx, (void*)&yw, (void*)&ystart, (void*)&xidx, (void*)&xw, (void*)&xstart, (void*)&TX,
                  (void*)&K, (void*)&plane_out, (void*)&H1, (void*)&W1, (void*)&per_cta, (void*)&hist3, (void*)&bar,
                  (void*)&sp.ranks, (void*)&sp.t0, (void*)&sp.t1, (void*)&pct_out};
  return coop_launch(blend_select_kernel, cd.sms, args, smem, stream, "frame_blend_percentiles");
}

// uncl_frame_postprocess + uncl_percentile_pair(clip(col, 0, 1), p_lo, p_hi) + uncl_frame_to_u8 in one launch.
// col_out ([3][H][W] fp32) may be NULL when only the 8-bit HWC image is wanted; pct_out[0..1] receives the colour percentiles.
extern "C" int uncl_frame_post_u8(const float* fake, int H1, int W1, const float* rgb, int H, int W, const float* stats,
                                  const float* pct_plane, float* col_out, double p_lo, double p_hi, float* pct_out,
                                  unsigned char* u8_out, void* workspace, cudaStream_t stream) {
  UNCL_REQUIRE(H > 0 && W > 0 && H1 >= H && W1 >= W && u8_out != nullptr && p_lo >= 0 && p_hi <= 100 && p_lo <= p_hi,
               "frame_post_u8: bad arguments");
  const CoopDev cd = coop_device();
  const long HW = (long)H * W;
  long px_per_cta = (HW + cd.sms - 1) / cd.sms;
  if (!cd.coop || 3 * px_per_cta > kFusedMaxKeys || 3 * HW >= (1L << 32))
    return uncl_set_error(UNCL_EUNSUPPORTED, "frame_post_u8: frame too large for the fused kernel");
  FrameWs w = carve(workspace);
  const SelPlan sp = sel_plan(3 * HW, p_lo, p_hi);
  cudaMemsetAsync(w.fused_barrier, 0, 16 + 3 * 4 * 2048 * 4, stream);
  unsigned* hist3 = w.fused_hist;
  unsigned* bar = w.fused_barrier;
  const int padT = (H1 - H) / 2, padL = (W1 - W) / 2;
  const size_t smem = (size_t)((kSelQ << 11) + 3 * px_per_cta) * sizeof(unsigned);
  void* args[] = {(void*)&fake, (void*)&W1, (void*)&padT, (void*)&padL, (void*)&rgb, (void*)&H, (void*)&W, (void*)&stats,
                  (void*)&pct_plane, (void*)&col_out, (void*)&px_per_cta, (void*)&hist3, (void*)&bar, (void*)&sp.ranks,
                  (void*)&sp.t0, (void*)&sp.t1, (void*)&pct_out, (void*)&u8_out};
  return coop_launch(post_select_u8_kernel, cd.sms, args, smem, stream, "frame_post_u8");
}
