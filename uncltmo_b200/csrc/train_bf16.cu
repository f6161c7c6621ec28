// Elementwise / reduction kernels of the bf16-activation training path (uncltmo_b200/train_graph.py).
//
// The generator backward keeps every activation and every pre-activation gradient as a C8-blocked bf16 tensor, exactly
// like the inference path.  The tensor-core kernels produce most of those gradients themselves (data gradient with the
// ReLU mask of the producing layer fused, uncl_conv3x3_tc_dgrad); what is left are the places where several gradient
// paths meet or a layout changes:
//   * skip_pool_bwd : a skip tensor x2 receives gradient from the decoder's concat [x2 | up | x2^2 | sqrt(x2+eps)]
//                     (unet_parts.py:319-322) and from the next encoder stage through MaxPool2d(2) (:210-213); both are
//                     combined, masked by the ReLU of the layer that produced x2, written once as bf16, and the bias
//                     gradient of that layer is reduced in the same pass.
//   * convT2x2_s2d  : space-to-depth of the up-convolution's output gradient (a channel slice of the concat gradient),
//                     replicate-pad fold included, + its bias gradient.
//   * outc_feat_bwd : 1x1 out conv + sigmoid backward, merged with the gradient arriving through the feature output
//                     (infoNCE2 on up_x, GanTrainerImg.py:384-408) and the ReLU mask of up3.conv1.
//   * bias_grad     : column sums of a bf16 gradient tensor.
//   * pack / unpack : all weights of a step re-laid out by ONE gather launch each (index maps built once on the host
//                     from uncltmo_b200/packing.py), instead of ~60 torch permute/contiguous kernels per step.
//   * nce (self)    : GanTrainer.nce for infoNCE2, positive / negative = rows of the anchor tensor chosen on the device.
//   * adam_flat     : torch.optim.Adam's update on one flat parameter buffer.
#include "common.cuh"

namespace {

__device__ __forceinline__ void block_reduce8_atomic(float (&s)[8], float* dst, float* sm /* [8][32] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = warp_sum(s[j]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[j * 32 + wid] = s[j];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = lane < nw ? sm[j * 32 + lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) atomicAdd(dst + j, t);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// grid (chunks, N * C/8); every block reduces its share of db for one channel block
// ---------------------------------------------------------------------------------------------------------
// Video recurrence (Unet.py:244): the max-pool read cat(prev[:, :r], x2[:, r:]).  With `prev` given (first channel block
// only, r <= 8), the window values of channels < r come from prev, their pool gradient goes to d_prev (dense [N][1][H][W][8],
// channels >= r zero) instead of dz, and d_state (dense [N][1][H][W][8], may be null) - the gradient the NEXT frame sent to
// this frame's own first r channels - is added before the ReLU mask.
__global__ void __launch_bounds__(256) skip_pool_bwd_kernel(const bf16* __restrict__ x2, long x2_img_stride,
                                                           const bf16* __restrict__ dcat, const bf16* __restrict__ dpool,
                                                           bf16* __restrict__ dz, float* __restrict__ db, int C, int H,
                                                           int W, const bf16* __restrict__ prev, long prev_img_stride,
                                                           int r, bf16* __restrict__ d_prev,
                                                           const bf16* __restrict__ d_state) {
  __shared__ float sm[8 * 32];
  const int Cb = C / 8, cb = blockIdx.y % Cb, n = blockIdx.y / Cb;
  const int HW = H * W, Hp = H / 2, Wp = W / 2;
  const bf16* xb = x2 + (long)n * x2_img_stride + (long)cb * HW * 8;
  const bool rec0 = cb == 0;       // the recurrent channels live in channel block 0
  const bf16* pb = (prev != nullptr && rec0) ? prev + (long)n * prev_img_stride : nullptr;
  bf16* dpo = (d_prev != nullptr && rec0) ? d_prev + (long)n * HW * 8 : nullptr;
  const bf16* dst_in = (d_state != nullptr && rec0) ? d_state + (long)n * HW * 8 : nullptr;
  const bf16* g0 = dcat ? dcat + ((long)n * 4 * Cb + cb) * HW * 8 : nullptr;
  const bf16* dp = dpool ? dpool + ((long)n * Cb + cb) * Hp * Wp * 8 : nullptr;
  bf16* o = dz + ((long)n * Cb + cb) * HW * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
    const int y = i / W, x = i - y * W;
    float a[8], g[8];
    load8(xb + (long)i * 8, a);
    if (g0 != nullptr) {
      float d0[8], d2[8], d3[8];
      load8(g0 + (long)i * 8, d0);
      load8(g0 + ((long)2 * Cb * HW + i) * 8, d2);
      load8(g0 + ((long)3 * Cb * HW + i) * 8, d3);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = d0[j] + 2.f * a[j] * d2[j] + 0.5f * d3[j] * rsqrtf(a[j] + 1e-8f);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = 0.f;
    }
    const int py = y >> 1, px = x >> 1;
    float gp[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // pool gradient owed to the previous frame (channels < r)
    if (dp != nullptr && py < Hp && px < Wp) {
      // MaxPool2d(2) backward: the FIRST maximum of the window in row-major order takes the gradient (as PyTorch)
      float v[4][8], d[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long o = ((long)(2 * py + (k >> 1)) * W + 2 * px + (k & 1)) * 8;
        load8(xb + o, v[k]);
        if (pb != nullptr) {
          float pv[8];
          load8(pb + o, pv);
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < r) v[k][j] = pv[j];
        }
      }
      load8(dp + ((long)py * Wp + px) * 8, d);
      const int me = ((y & 1) << 1) | (x & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int best = 0;
        float bv = v[0][j];
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (v[k][j] > bv) { bv = v[k][j]; best = k; }
        if (best == me) {
          if (pb != nullptr && j < r) gp[j] = d[j];
          else g[j] += d[j];
        }
      }
    }
    if (dpo != nullptr) store8(dpo + (long)i * 8, gp);
    if (dst_in != nullptr) {
      float ds[8];
      load8(dst_in + (long)i * 8, ds);
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < r) g[j] += ds[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      g[j] = a[j] > 0.f ? g[j] : 0.f;      // ReLU of the layer that produced x2
      g[j] = __bfloat162float(__float2bfloat16_rn(g[j]));   // db sums what the gradient GEMMs will read
      acc[j] += g[j];
    }
    store8(o + (long)i * 8, g);
  }
  if (db != nullptr) block_reduce8_atomic(acc, db + cb * 8, sm);
}

// Video recurrence on a decoder tensor (Unet.py:270: the up-convolution read cat(prev[:, :r], up[:, r:])), channel block 0
// of the gradient dz [N][C/8][HW][8] that the up-conv's data gradient produced for its (spliced) input:
//   spliced (d_prev != null): d_prev[ch < r] = dz[ch], dz[ch < r] = 0     (those channels belonged to the previous frame)
//   d_state != null          : dz[ch < r] += (own[ch] > 0) * d_state[ch]   (what the NEXT frame sent to this frame's slice)
__global__ void __launch_bounds__(256) splice_grad_kernel(bf16* __restrict__ dz, long dz_img_stride,
                                                         const bf16* __restrict__ own, long own_img_stride, int r,
                                                         bf16* __restrict__ d_prev, const bf16* __restrict__ d_state,
                                                         long HW, long total) {
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const long n = i / HW, p = i - n * HW;
    float g[8], o[8];
    bf16* gz = dz + n * dz_img_stride + p * 8;
    load8(gz, g);
    if (d_prev != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j] = j < r ? g[j] : 0.f; if (j < r) g[j] = 0.f; }
      store8(d_prev + (n * HW + p) * 8, o);
    }
    if (d_state != nullptr) {
      float ds[8], m[8];
      load8(d_state + (n * HW + p) * 8, ds);
      if (own != nullptr) load8(own + n * own_img_stride + p * 8, m);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < r && (own == nullptr || m[j] > 0.f)) g[j] += ds[j];
    }
    store8(gz, g);
  }
}

__global__ void __launch_bounds__(256) bias_grad_kernel(const bf16* __restrict__ dz, long img_stride, float* __restrict__ db,
                                                       int C, int HW) {
  __shared__ float sm[8 * 32];
  const int Cb = C / 8, cb = blockIdx.y % Cb, n = blockIdx.y / Cb;
  const bf16* d = dz + (long)n * img_stride + (long)cb * HW * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
    float v[8];
    load8(d + (long)i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
  block_reduce8_atomic(acc, db + cb * 8, sm);
}

// out[n, pos*C + co, y, x] = sum over the replicate-padded copies of dY[n, co, 2y+dy, 2x+dx]; db[co] += everything
__global__ void __launch_bounds__(256) convT2x2_s2d_bf16_kernel(const bf16* __restrict__ dY, long dy_img_stride,
                                                               bf16* __restrict__ out, float* __restrict__ db, int C,
                                                               int H, int W, int H2, int W2) {
  __shared__ float sm[8 * 32];
  const int Cb = C / 8, cb4 = blockIdx.y % (4 * Cb), n = blockIdx.y / (4 * Cb);
  const int pos = cb4 / Cb, cb = cb4 - pos * Cb;
  const int padT = (H2 - 2 * H) / 2, padL = (W2 - 2 * W) / 2;
  const bf16* d = dY + (long)n * dy_img_stride + (long)cb * H2 * W2 * 8;
  bf16* o = out + ((long)n * 4 * Cb + cb4) * H * W * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = blockIdx.x * 256 + threadIdx.x; i < H * W; i += gridDim.x * 256) {
    const int y = i / W, x = i - y * W;
    const int Y = 2 * y + (pos >> 1), X = 2 * x + (pos & 1);
    const int y_lo = (Y == 0) ? 0 : Y + padT, y_hi = (Y == 2 * H - 1) ? H2 - 1 : Y + padT;
    const int x_lo = (X == 0) ? 0 : X + padL, x_hi = (X == 2 * W - 1) ? W2 - 1 : X + padL;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int yy = y_lo; yy <= y_hi; ++yy)
      for (int xx = x_lo; xx <= x_hi; ++xx) {
        float v[8];
        load8(d + ((long)yy * W2 + xx) * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += v[j];
      }
    store8(o + (long)i * 8, s);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += s[j];
  }
  if (db != nullptr) block_reduce8_atomic(acc, db + cb * 8, sm);
}

// ---------------------------------------------------------------------------------------------------------
// first conv (C_in = 1) weight + bias gradient in one pass over dZ:  dW[t][c] += sum x[n, y+ky, x+kx] * dZ[n, c, y, x],
// db[c] += sum dZ.  One thread takes the 8 channels of a channel block at one pixel (one 16 / 32-byte load) and keeps
// 9 x 8 + 8 partial sums in registers; grid (chunks, N * C/8).
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv_first_wgrad_bias_kernel(const float* __restrict__ x, const T* __restrict__ dZ,
                                                                   float* __restrict__ dW, float* __restrict__ db, int H,
                                                                   int W, int C) {
  __shared__ float sm[8 * 32];
  const int Cb = C / 8, cb = blockIdx.y % Cb, n = blockIdx.y / Cb;
  const int Ho = H - 2, Wo = W - 2, HWo = Ho * Wo;
  const T* g0 = dZ + ((long)n * Cb + cb) * HWo * 8;
  const float* xn = x + (long)n * H * W;
  float acc[10][8];
#pragma unroll
  for (int t = 0; t < 10; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HWo; p += gridDim.x * 256) {
    const int oy = p / Wo, ox = p - oy * Wo;
    float g[8];
    load8(g0 + (long)p * 8, g);
    const float* xi = xn + (long)oy * W + ox;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float xv = __ldg(xi + (t / 3) * W + t % 3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(xv, g[j], acc[t][j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[9][j] += g[j];
  }
#pragma unroll
  for (int t = 0; t < 10; ++t) {
    float* dst = t < 9 ? dW + t * C + cb * 8 : (db != nullptr ? db + cb * 8 : nullptr);
    if (dst != nullptr) block_reduce8_atomic(acc[t], dst, sm);   // (uniform branch: every thread takes it)
  }
}

// ---------------------------------------------------------------------------------------------------------
// out conv (1x1, C = 32 -> 1) + sigmoid backward, merged with the feature-path gradient and the ReLU of `up`:
//   dl = d_out * o * (1 - o);  dz[c] = up[c] > 0 ? dl * w[c] + d_feat[c] : 0
//   dw[c] += sum dl * up[c];  db_out += sum dl;  db_up[c] += sum dz[c]
// one thread per pixel; per-thread partial sums over a grid-stride loop, reduced once per block
// ---------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) outc_feat_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ out,
                                                           const bf16* __restrict__ up, long up_img_stride,
                                                           const bf16* __restrict__ d_feat, const float* __restrict__ w,
                                                           bf16* __restrict__ dz, float* __restrict__ dw,
                                                           float* __restrict__ db_out, float* __restrict__ db_up, int HW,
                                                           int N) {
  __shared__ float s_acc[2 * C + 1];
  for (int i = threadIdx.x; i <= 2 * C; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  float wl[C];
#pragma unroll
  for (int c = 0; c < C; ++c) wl[c] = __ldg(w + c);
  float a_dw[C], a_db[C], a_lb = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) { a_dw[c] = 0.f; a_db[c] = 0.f; }
  const long total = (long)N * HW;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int p = (int)(i % HW), n = (int)(i / HW);
    const float o = out[i];
    const float dl = d_out ? d_out[i] * o * (1.f - o) : 0.f;
    a_lb += dl;
#pragma unroll
    for (int cb = 0; cb < C / 8; ++cb) {
      float u[8], g[8], f[8];
      load8(up + (long)n * up_img_stride + ((long)cb * HW + p) * 8, u);
      if (d_feat != nullptr) load8(d_feat + (((long)n * (C / 8) + cb) * HW + p) * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = dl * wl[cb * 8 + j] + (d_feat != nullptr ? f[j] : 0.f);
        v = u[j] > 0.f ? v : 0.f;
        v = __bfloat162float(__float2bfloat16_rn(v));
        g[j] = v;
        a_dw[cb * 8 + j] = fmaf(dl, u[j], a_dw[cb * 8 + j]);
        a_db[cb * 8 + j] += v;
      }
      store8(dz + (((long)n * (C / 8) + cb) * HW + p) * 8, g);
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float s1 = warp_sum(a_dw[c]), s2 = warp_sum(a_db[c]);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_acc[c], s1); atomicAdd(&s_acc[C + c], s2); }
  }
  a_lb = warp_sum(a_lb);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[2 * C], a_lb);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    atomicAdd(dw + i, s_acc[i]);
    if (db_up) atomicAdd(db_up + i, s_acc[C + i]);
  }
  if (threadIdx.x == 0) atomicAdd(db_out, s_acc[2 * C]);
}

// ---------------------------------------------------------------------------------------------------------
// weight packing / gradient unpacking by index map
// ---------------------------------------------------------------------------------------------------------
// idx bit 30: store the bf16 RESIDUAL w - bf16(w) (the `lo` term of a split-bf16 operand); bits 0-29: source element
__global__ void __launch_bounds__(256) pack_gather_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                         bf16* __restrict__ dst, long n) {
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const int k = idx[i];
    if (k < 0) { dst[i] = __float2bfloat16_rn(0.f); continue; }
    const float v = __ldg(src + (k & 0x3fffffff));
    const bf16 hi = __float2bfloat16_rn(v);
    dst[i] = (k & 0x40000000) ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
  }
}
__global__ void __launch_bounds__(256) unpack_gather_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                           float* __restrict__ dst, long n) {
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const int k = idx[i];
    if (k >= 0) dst[i] = __ldg(src + k);
  }
}

__global__ void __launch_bounds__(256) unpack_add_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                        float* __restrict__ dst, long n) {
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const int k = idx[i];
    if (k >= 0) dst[i] += __ldg(src + k);
  }
}

// ---------------------------------------------------------------------------------------------------------
// nce on one feature tensor whose positive / negative are two of its own rows (infoNCE2), any element layout
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float nce_term(float a, float q, float k, float c0) {
  return __fdividef(a * q, c0 + k * fabsf(a - q));
}
// d/da and d/dq of a*q / (c0 + k|a-q|)
__device__ __forceinline__ void nce_grad(float a, float q, float k, float c0, float& da, float& dq) {
  const float d = a - q, inv = __fdividef(1.f, c0 + k * fabsf(d));
  const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  const float t = a * q * k * sg * inv * inv;
  da = q * inv - t;
  dq = a * inv + t;
}

__global__ void __launch_bounds__(256) nce_self_fwd_kernel(const bf16* __restrict__ fea, const long* __restrict__ sel,
                                                          const bf16* __restrict__ ext, int B, long CHW, float k,
                                                          float c0, float inv_hw, float* __restrict__ logits) {
  __shared__ float red[33];
  const int b = blockIdx.y;
  const bf16* ab = fea + (long)b * CHW;
  const long ip = sel[0], in = sel[1];   // rows >= B live in `ext` (selected on another rank, data parallel)
  const bf16* pb = ip < B ? fea + ip * CHW : ext + (ip - B) * CHW;
  const bf16* nb = in < B ? fea + in * CHW : ext + (in - B) * CHW;
  float sp = 0.f, sn = 0.f;
  for (long i = ((long)blockIdx.x * 256 + threadIdx.x) * 8; i < CHW; i += (long)gridDim.x * 256 * 8) {
    float a[8], p[8], q[8];
    load8(ab + i, a); load8(pb + i, p); load8(nb + i, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) { sp += nce_term(a[j], p[j], k, c0); sn += nce_term(a[j], q[j], k, c0); }
  }
  sp = block_sum(sp * inv_hw, red);
  sn = block_sum(sn * inv_hw, red);
  if (threadIdx.x == 0) { atomicAdd(logits + 2 * b, sp); atomicAdd(logits + 2 * b + 1, sn); }
}

__global__ void nce_ce_kernel(const float* __restrict__ logits, int B, float* __restrict__ loss) {
  // cross entropy of [pos, neg] rows against class 0, mean over the batch (one warp)
  float v = 0.f;
  for (int b = threadIdx.x; b < B; b += 32) {
    const float l0 = logits[2 * b], l1 = logits[2 * b + 1], mx = fmaxf(l0, l1);
    v += mx + logf(expf(l0 - mx) + expf(l1 - mx)) - l0;
  }
  v = warp_sum(v);
  if (threadIdx.x == 0) loss[0] = v / (float)B;
}

// One thread owns 8 consecutive elements of EVERY row: it walks the batch once, writes each row's anchor gradient and
// accumulates the gradient of the broadcast positive / negative in registers; the two selected rows are stored last,
// with that sum added.  (A row-parallel grid leaves the two selected rows with B times the work of the others.)
template <bool OUT_BF16>
__global__ void __launch_bounds__(256) nce_self_bwd_kernel(const bf16* __restrict__ fea, const long* __restrict__ sel,
                                                          const bf16* __restrict__ ext, long CHW, float k, float c0,
                                                          float inv_hw, const float* __restrict__ logits, int B,
                                                          const float* __restrict__ g_up, void* __restrict__ d_fea_v,
                                                          float* __restrict__ d_ext) {
  extern __shared__ float s_dl[];   // [2][B]
  for (int b = threadIdx.x; b < B; b += 256) {
    const float l0 = logits[2 * b], l1 = logits[2 * b + 1], mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), p0 = e0 / (e0 + e1);
    const float g = __ldg(g_up) / (float)B * inv_hw;
    s_dl[b] = g * (p0 - 1.f);
    s_dl[B + b] = g * (1.f - p0);
  }
  __syncthreads();
  const long ip = sel[0], in = sel[1];
  bf16* const d16 = reinterpret_cast<bf16*>(d_fea_v);
  float* const d32 = reinterpret_cast<float*>(d_fea_v);
  for (long i = ((long)blockIdx.x * 256 + threadIdx.x) * 8; i < CHW; i += (long)gridDim.x * 256 * 8) {
    float p[8], q[8], dp[8], dq[8], op[8], oq[8];
    load8((ip < B ? fea + ip * CHW : ext + (ip - B) * CHW) + i, p);
    load8((in < B ? fea + in * CHW : ext + (in - B) * CHW) + i, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) { dp[j] = 0.f; dq[j] = 0.f; op[j] = 0.f; oq[j] = 0.f; }
    for (int b = 0; b < B; ++b) {
      float a[8], o[8];
      load8(fea + (long)b * CHW + i, a);
      const float dl0 = s_dl[b], dl1 = s_dl[B + b];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float da_p, dp_, da_n, dn_;
        nce_grad(a[j], p[j], k, c0, da_p, dp_);
        nce_grad(a[j], q[j], k, c0, da_n, dn_);
        o[j] = dl0 * da_p + dl1 * da_n;
        dp[j] = fmaf(dl0, dp_, dp[j]);
        dq[j] = fmaf(dl1, dn_, dq[j]);
      }
      if (b == ip || b == in) {
        // deferred: this row also receives the batch-summed gradient of the broadcast operand (both, if ip == in)
#pragma unroll
        for (int j = 0; j < 8; ++j) { if (b == ip) op[j] = o[j]; if (b == in) oq[j] = o[j]; }
      } else if (OUT_BF16) {
        store8(d16 + (long)b * CHW + i, o);
      } else {
        store8(d32 + (long)b * CHW + i, o);
      }
    }
    if (ip >= B || in >= B) {
      // external rows: their (partial, this rank's anchors only) gradient goes to d_ext [2][CHW] fp32
      if (ip >= B) store8(d_ext + (ip - B) * CHW + i, dp);
      if (in >= B) store8(d_ext + (in - B) * CHW + i, dq);
      if (ip < B) {
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] += dp[j];
        if (OUT_BF16) store8(d16 + ip * CHW + i, op); else store8(d32 + ip * CHW + i, op);
      }
      if (in < B) {
#pragma unroll
        for (int j = 0; j < 8; ++j) oq[j] += dq[j];
        if (OUT_BF16) store8(d16 + in * CHW + i, oq); else store8(d32 + in * CHW + i, oq);
      }
    } else if (ip == in) {
#pragma unroll
      for (int j = 0; j < 8; ++j) op[j] += dp[j] + dq[j];
      if (OUT_BF16) store8(d16 + ip * CHW + i, op); else store8(d32 + ip * CHW + i, op);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { op[j] += dp[j]; oq[j] += dq[j]; }
      if (OUT_BF16) { store8(d16 + ip * CHW + i, op); store8(d16 + in * CHW + i, oq); }
      else { store8(d32 + ip * CHW + i, op); store8(d32 + in * CHW + i, oq); }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam semantics, no amsgrad / weight decay / maximize) on a flat buffer; `step` lives on the device
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, long n, float lr,
                                                       float b1, float b2, float eps, const float* __restrict__ step) {
  const float t = step[0];
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}
__global__ void bump_step_kernel(float* step) { step[0] += 1.f; }

int grid_for(long n, int per_sm = 8) {
  long g = (n + 255) / 256;
  const long cap = 148L * per_sm;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

extern "C" int uncl_skip_pool_bwd(const void* x2, long x2_img_stride, const void* dcat, const void* dpool, void* dz, float* db,
                                  int N, int C, int H, int W, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && H > 1 && W > 1 && x2 && dz && x2_img_stride % 8 == 0, "skip_pool_bwd: bad arguments");
  int chunks = ceil_div(H * W, 256 * 4);
  if (chunks > 64) chunks = 64;
  skip_pool_bwd_kernel<<<dim3(chunks, N * (C / 8)), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(x2), x2_img_stride, reinterpret_cast<const bf16*>(dcat),
      reinterpret_cast<const bf16*>(dpool), reinterpret_cast<bf16*>(dz), db, C, H, W, nullptr, 0, 0, nullptr, nullptr);
  return uncl_check_launch("skip_pool_bwd");
}

extern "C" int uncl_skip_pool_bwd_rec(const void* x2, long x2_img_stride, const void* dcat, const void* dpool, void* dz,
                                      float* db, int N, int C, int H, int W, const void* prev, long prev_img_stride, int r,
                                      void* d_prev, const void* d_state, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && H > 1 && W > 1 && x2 && dz && x2_img_stride % 8 == 0 && r >= 0 && r <= 8 &&
                   prev_img_stride % 8 == 0 && (prev == nullptr) == (d_prev == nullptr),
               "skip_pool_bwd_rec: bad arguments");
  int chunks = ceil_div(H * W, 256 * 4);
  if (chunks > 64) chunks = 64;
  skip_pool_bwd_kernel<<<dim3(chunks, N * (C / 8)), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(x2), x2_img_stride, reinterpret_cast<const bf16*>(dcat),
      reinterpret_cast<const bf16*>(dpool), reinterpret_cast<bf16*>(dz), db, C, H, W, reinterpret_cast<const bf16*>(prev),
      prev_img_stride, r, reinterpret_cast<bf16*>(d_prev), reinterpret_cast<const bf16*>(d_state));
  return uncl_check_launch("skip_pool_bwd_rec");
}

extern "C" int uncl_splice_grad(void* dz, long dz_img_stride, const void* own, long own_img_stride, int r, void* d_prev,
                                const void* d_state, int N, long HW, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && HW > 0 && dz && r >= 0 && r <= 8 && dz_img_stride % 8 == 0 && own_img_stride % 8 == 0,
               "splice_grad: bad arguments");
  const long total = (long)N * HW;
  splice_grad_kernel<<<grid_for(total), 256, 0, stream>>>(reinterpret_cast<bf16*>(dz), dz_img_stride,
                                                         reinterpret_cast<const bf16*>(own), own_img_stride, r,
                                                         reinterpret_cast<bf16*>(d_prev), reinterpret_cast<const bf16*>(d_state),
                                                         HW, total);
  return uncl_check_launch("splice_grad");
}

extern "C" int uncl_bias_grad_bf16(const void* dz, long img_stride, float* db, int N, int C, int HW, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && HW > 0 && dz && db && img_stride % 8 == 0, "bias_grad_bf16: bad arguments");
  int chunks = ceil_div(HW, 256 * 8);
  if (chunks > 32) chunks = 32;
  bias_grad_kernel<<<dim3(chunks, N * (C / 8)), 256, 0, stream>>>(reinterpret_cast<const bf16*>(dz), img_stride, db, C, HW);
  return uncl_check_launch("bias_grad_bf16");
}

extern "C" int uncl_convT2x2_s2d_bf16(const void* dY, long dy_img_stride, void* out, float* db, int N, int C, int H, int W,
                                      int H2, int W2, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && H2 >= 2 * H && W2 >= 2 * W && dY && out && dy_img_stride % 8 == 0,
               "convT2x2_s2d_bf16: bad arguments");
  int chunks = ceil_div(H * W, 256 * 4);
  if (chunks > 32) chunks = 32;
  convT2x2_s2d_bf16_kernel<<<dim3(chunks, N * 4 * (C / 8)), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(dY), dy_img_stride, reinterpret_cast<bf16*>(out), db, C, H, W, H2, W2);
  return uncl_check_launch("convT2x2_s2d_bf16");
}

extern "C" int uncl_conv_first_wgrad_bias(const float* x, const void* dZ, int dz_dtype, float* dW, float* db, int N, int H,
                                          int W, int C, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && H > 2 && W > 2 && x && dZ && dW, "conv_first_wgrad_bias: bad arguments");
  int chunks = ceil_div((H - 2) * (W - 2), 256 * 8);
  if (chunks > 32) chunks = 32;
  const dim3 grid(chunks, N * (C / 8));
  if (dz_dtype == UNCL_BF16)
    conv_first_wgrad_bias_kernel<bf16><<<grid, 256, 0, stream>>>(x, reinterpret_cast<const bf16*>(dZ), dW, db, H, W, C);
  else if (dz_dtype == UNCL_F32)
    conv_first_wgrad_bias_kernel<float><<<grid, 256, 0, stream>>>(x, reinterpret_cast<const float*>(dZ), dW, db, H, W, C);
  else
    return uncl_set_error(UNCL_EINVAL, "conv_first_wgrad_bias: bad dz_dtype");
  return uncl_check_launch("conv_first_wgrad_bias");
}

extern "C" int uncl_outc_feat_bwd(const float* d_out, const float* out, const void* up, long up_img_stride,
                                  const void* d_feat, const float* w, void* dz, float* dw, float* db_out, float* db_up, int N,
                                  int C, int HW, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C == 32 && HW > 0 && out && up && w && dz && dw && db_out, "outc_feat_bwd: only C = 32 is built");
  outc_feat_bwd_kernel<32><<<148 * 2, 256, 0, stream>>>(d_out, out, reinterpret_cast<const bf16*>(up), up_img_stride,
                                                      reinterpret_cast<const bf16*>(d_feat), w, reinterpret_cast<bf16*>(dz),
                                                      dw, db_out, db_up, HW, N);
  return uncl_check_launch("outc_feat_bwd");
}

extern "C" int uncl_pack_gather(const float* src, const int* idx, void* dst_bf16, long n, cudaStream_t stream) {
  UNCL_REQUIRE(n > 0 && src && idx && dst_bf16, "pack_gather: bad arguments");
  pack_gather_kernel<<<grid_for(n), 256, 0, stream>>>(src, idx, reinterpret_cast<bf16*>(dst_bf16), n);
  return uncl_check_launch("pack_gather");
}

extern "C" int uncl_unpack_gather(const float* src, const int* idx, float* dst, long n, cudaStream_t stream) {
  UNCL_REQUIRE(n > 0 && src && idx && dst, "unpack_gather: bad arguments");
  unpack_gather_kernel<<<grid_for(n), 256, 0, stream>>>(src, idx, dst, n);
  return uncl_check_launch("unpack_gather");
}

extern "C" int uncl_unpack_add(const float* src, const int* idx, float* dst, long n, cudaStream_t stream) {
  UNCL_REQUIRE(n > 0 && src && idx && dst, "unpack_add: bad arguments");
  unpack_add_kernel<<<grid_for(n), 256, 0, stream>>>(src, idx, dst, n);
  return uncl_check_launch("unpack_add");
}

extern "C" int uncl_nce_self_fwd(const void* fea, const long* sel, const void* ext, int B, long CHW, int HW, float k,
                                 float constant, float* logits_scratch, float* loss_out, cudaStream_t stream) {
  UNCL_REQUIRE(B > 0 && CHW % 8 == 0 && HW > 0 && fea && sel && logits_scratch && loss_out, "nce_self_fwd: bad arguments");
  cudaMemsetAsync(logits_scratch, 0, 2 * B * sizeof(float), stream);
  int gx = grid_for(CHW / 8, 4) / B;
  if (gx < 1) gx = 1;
  nce_self_fwd_kernel<<<dim3(gx, B), 256, 0, stream>>>(reinterpret_cast<const bf16*>(fea), sel,
                                                      reinterpret_cast<const bf16*>(ext), B, CHW, k, constant,
                                                      1.f / (float)HW, logits_scratch);
  nce_ce_kernel<<<1, 32, 0, stream>>>(logits_scratch, B, loss_out);
  return uncl_check_launch("nce_self_fwd");
}

extern "C" int uncl_nce_self_bwd(const void* fea, const long* sel, const void* ext, int B, long CHW, int HW, float k,
                                 float constant, const float* logits, const float* g_up, void* d_fea, int d_dtype,
                                 float* d_ext, cudaStream_t stream) {
  UNCL_REQUIRE(B > 0 && B <= 1024 && CHW % 8 == 0 && fea && sel && logits && g_up && d_fea, "nce_self_bwd: bad arguments");
  UNCL_REQUIRE(d_dtype == UNCL_F32 || d_dtype == UNCL_BF16, "nce_self_bwd: bad d_dtype");
  const int gx = grid_for(CHW / 8, 8);
  if (d_dtype == UNCL_BF16)
    nce_self_bwd_kernel<true><<<gx, 256, 2 * B * sizeof(float), stream>>>(
        reinterpret_cast<const bf16*>(fea), sel, reinterpret_cast<const bf16*>(ext), CHW, k, constant, 1.f / (float)HW, logits,
        B, g_up, d_fea, d_ext);
  else
    nce_self_bwd_kernel<false><<<gx, 256, 2 * B * sizeof(float), stream>>>(
        reinterpret_cast<const bf16*>(fea), sel, reinterpret_cast<const bf16*>(ext), CHW, k, constant, 1.f / (float)HW, logits,
        B, g_up, d_fea, d_ext);
  return uncl_check_launch("nce_self_bwd");
}

extern "C" int uncl_adam_flat(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
                              float eps, float* step, cudaStream_t stream) {
  UNCL_REQUIRE(n > 0 && p && g && m && v && step, "adam_flat: bad arguments");
  bump_step_kernel<<<1, 1, 0, stream>>>(step);
  adam_flat_kernel<<<grid_for(n), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, step);
  return uncl_check_launch("adam_flat");
}
