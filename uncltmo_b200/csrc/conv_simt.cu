// CUDA-core (fp32-accumulate) kernels of the generator: first 1->C conv, generic 3x3 conv / ConvTranspose-as-conv,
// ConvTranspose k2 s2 (+ replicate pad into the skip concat buffer), 2x2 max-pool, 1x1 out conv + sigmoid,
// layout transforms.  These are the exact-fp32 path (generator rel-L2 <= 1e-4 gate) and the building blocks the
// tcgen05 path (conv_tc.cu) reuses for the non-GEMM layers.
//
// Reference semantics: models/unet_multi_filters/unet_parts.py:57-87 (double_conv), :126-141 (double_last_conv),
// :183-193 (double_conv_traspose), :283-335 (up: ConvTranspose k2 s2, replicate pad, concat operators), :338-345 (outconv).
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// 3x3 conv, C_in = 1 (inc.conv): x [N][H][W] fp32 -> blocked [N][C_out/8][H-2][W-2][8], bias + ReLU
// ------------------------------------------------------------------------------------------------
// thread = one output column, two output rows (they share 6 of their 12 input values); the weights of 4 output
// channels travel as one 128-bit shared-memory broadcast, so the kernel is FMA-bound (288 per pixel), not LDS-bound
template <typename T>
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, T* __restrict__ out,
                                                        long out_img_stride, int H, int W, int C_out, int act) {
  // w: [9][C_out]
  extern __shared__ __align__(16) float s_w[];  // 9*C_out + C_out
  float* s_b = s_w + 9 * C_out;
  for (int i = threadIdx.x; i < 9 * C_out; i += blockDim.x) s_w[i] = w[i];
  for (int i = threadIdx.x; i < C_out; i += blockDim.x) s_b[i] = bias[i];
  __syncthreads();
  const int Ho = H - 2, Wo = W - 2;
  const int n = blockIdx.z;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy = blockIdx.y * 16 + 2 * (threadIdx.x >> 5);
  if (ox >= Wo || oy >= Ho) return;
  const bool two = oy + 1 < Ho;
  const float* xi = x + (long)n * H * W + (long)oy * W + ox;
  float a[12];
#pragma unroll
  for (int ky = 0; ky < 4; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) a[ky * 3 + kx] = (ky < 3 || two) ? __ldg(xi + ky * W + kx) : 0.f;
  T* o = out + (long)n * out_img_stride + ((long)oy * Wo + ox) * 8;
  const long cb_stride = (long)Ho * Wo * 8;
  const float4* w4 = reinterpret_cast<const float4*>(s_w);
  const float4* b4 = reinterpret_cast<const float4*>(s_b);
  const int c4 = C_out / 4;
  for (int cb = 0; cb < C_out / 8; ++cb) {
    float v0[8], v1[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 bb = b4[cb * 2 + h];
      float p0[4] = {bb.x, bb.y, bb.z, bb.w}, p1[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 ww = w4[t * c4 + cb * 2 + h];
        const float wv[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          p0[j] = fmaf(a[t], wv[j], p0[j]);
          p1[j] = fmaf(a[t + 3], wv[j], p1[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v0[h * 4 + j] = apply_act(p0[j], act);
        v1[h * 4 + j] = apply_act(p1[j], act);
      }
    }
    store8(o + cb * cb_stride, v0);
    if (two) store8(o + cb * cb_stride + (long)Wo * 8, v1);
  }
}

// ------------------------------------------------------------------------------------------------
// generic 3x3 stride-1 conv on blocked tensors.  pad = 0 (valid) or 2 ("full": ConvTranspose 3x3 s1 p0 with
// flipped weights).  Weights packed [9][C_in][C_out] fp32.  Optional skip emission: besides y, write y^2 and
// sqrt(y + 1e-8) at channel-block offsets 2*Cb and 3*Cb of the same (concat) buffer - the concat operator
// 'square_and_square_root' of unet_parts.py:319-322 produced by the layer that makes the skip tensor.
// ------------------------------------------------------------------------------------------------
constexpr int CS_TX = 32, CS_TY = 8, CS_CO = 32;

template <typename T>
__global__ void __launch_bounds__(256) conv3x3_simt_kernel(const T* __restrict__ in, long in_img_stride,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          T* __restrict__ out, long out_img_stride, int C_in, int H,
                                                          int W, int C_out, int pad, int act, int emit_skip) {
  __shared__ float s_in[8][CS_TY + 2][CS_TX + 2];
  __shared__ __align__(16) float s_w[9][8][CS_CO];
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  const int co_tiles = C_out / CS_CO;
  const int n = blockIdx.z / co_tiles, co0 = (blockIdx.z % co_tiles) * CS_CO;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ox0 = blockIdx.x * CS_TX, oy0 = blockIdx.y * CS_TY;
  const T* inn = in + (long)n * in_img_stride;
  float acc[CS_CO];
#pragma unroll
  for (int j = 0; j < CS_CO; ++j) acc[j] = 0.f;

  for (int cb = 0; cb < C_in / 8; ++cb) {
    __syncthreads();
    for (int p = threadIdx.x; p < (CS_TY + 2) * (CS_TX + 2); p += 256) {
      const int ly = p / (CS_TX + 2), lx = p % (CS_TX + 2);
      const int iy = oy0 + ly - pad, ix = ox0 + lx - pad;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) load8(inn + (((long)cb * H + iy) * W + ix) * 8, v);
#pragma unroll
      for (int c = 0; c < 8; ++c) s_in[c][ly][lx] = v[c];
    }
    for (int i = threadIdx.x; i < 9 * 8 * CS_CO; i += 256) {
      const int co = i % CS_CO, c = (i / CS_CO) % 8, t = i / (CS_CO * 8);
      s_w[t][c][co] = w[((long)t * C_in + cb * 8 + c) * C_out + co0 + co];
    }
    __syncthreads();
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float a = s_in[c][ty + t / 3][tx + t % 3];
        const float4* wv = reinterpret_cast<const float4*>(&s_w[t][c][0]);
#pragma unroll
        for (int j = 0; j < CS_CO / 4; ++j) {
          const float4 q = wv[j];
          acc[4 * j + 0] = fmaf(a, q.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(a, q.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(a, q.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(a, q.w, acc[4 * j + 3]);
        }
      }
    }
  }
  const int ox = ox0 + tx, oy = oy0 + ty;
  if (ox >= Wo || oy >= Ho) return;
  const long cb_stride = (long)Ho * Wo * 8;
  const int Cb = C_out / 8;
  T* o = out + (long)n * out_img_stride + (long)(co0 / 8) * cb_stride + ((long)oy * Wo + ox) * 8;
#pragma unroll
  for (int b = 0; b < CS_CO / 8; ++b) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(acc[b * 8 + j] + __ldg(bias + co0 + b * 8 + j), act);
    store8(o + b * cb_stride, v);
    if (emit_skip) {
      float s[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = v[j] * v[j];
      store8(o + (b + 2 * Cb) * cb_stride, s);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = sqrtf(v[j] + 1e-8f);
      store8(o + (b + 3 * Cb) * cb_stride, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// ConvTranspose k2 s2 (+bias) on blocked tensors, written into a (H2 x W2) target with replicate padding
// (F.pad(..., mode='replicate') of unet_parts.py:292-299; only up_path.1 pads: 56 -> 57).
// Weights packed [C_in][4][C_out] fp32 (pos = dy*2+dx).  Optional recurrent splice: the first r input channels
// are read from `prev` (video generator, Unet.py:270).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) convT2x2_kernel(const T* __restrict__ in, long in_img_stride,
                                                      const T* __restrict__ prev, long prev_img_stride, int r,
                                                      const float* __restrict__ w, const float* __restrict__ bias,
                                                      T* __restrict__ out, long out_img_stride, int C, int H, int W,
                                                      int H2, int W2) {
  // CTA: 64 input pixels x 32 output channels x 4 positions. thread: 1 pixel x 1 position-pair... see below
  __shared__ float s_in[8][64];
  __shared__ __align__(16) float s_w[8][4][32];
  const int co_tiles = C / 32;
  const int n = blockIdx.y / co_tiles, co0 = (blockIdx.y % co_tiles) * 32;
  const int px = threadIdx.x & 63, pos = threadIdx.x >> 6;  // 4 positions
  const int p = blockIdx.x * 64 + px;
  const int HW = H * W;
  const T* inn = in + (long)n * in_img_stride;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (int cb = 0; cb < C / 8; ++cb) {
    __syncthreads();
    if (threadIdx.x < 64) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p < HW) {
        load8(inn + ((long)cb * HW + p) * 8, v);
        if (prev != nullptr && cb * 8 < r) {
          float pv[8];
          load8(prev + (long)n * prev_img_stride + ((long)cb * HW + p) * 8, pv);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (cb * 8 + c < r) v[c] = pv[c];
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) s_in[c][px] = v[c];
    }
    for (int i = threadIdx.x; i < 8 * 4 * 32; i += 256) {
      const int co = i & 31, ps = (i >> 5) & 3, c = i >> 7;
      s_w[c][ps][co] = w[((long)(cb * 8 + c) * 4 + ps) * C + co0 + co];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float a = s_in[c][px];
      const float4* wv = reinterpret_cast<const float4*>(&s_w[c][pos][0]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q = wv[j];
        acc[4 * j + 0] = fmaf(a, q.x, acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(a, q.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(a, q.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(a, q.w, acc[4 * j + 3]);
      }
    }
  }
  if (p >= HW) return;
  const int y = p / W, x = p % W;
  const int Y = 2 * y + (pos >> 1), X = 2 * x + (pos & 1);
  const int padT = (H2 - 2 * H) / 2, padL = (W2 - 2 * W) / 2;
  const int y_lo = (Y == 0) ? 0 : Y + padT, y_hi = (Y == 2 * H - 1) ? H2 - 1 : Y + padT;
  const int x_lo = (X == 0) ? 0 : X + padL, x_hi = (X == 2 * W - 1) ? W2 - 1 : X + padL;
  const long cb_stride = (long)H2 * W2 * 8;
  T* on = out + (long)n * out_img_stride + (long)(co0 / 8) * cb_stride;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[b * 8 + j] + __ldg(bias + co0 + b * 8 + j);
    for (int yy = y_lo; yy <= y_hi; ++yy)
      for (int xx = x_lo; xx <= x_hi; ++xx) store8(on + b * cb_stride + ((long)yy * W2 + xx) * 8, v);
  }
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(2) (floor) on blocked tensors; optional recurrent splice of the first r channels from `prev`
// (video generator, Unet.py:244: the pooled stage input is cat(prev[:, :r], cur[:, r:])).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_kernel(const T* __restrict__ in, long in_img_stride,
                                                      const T* __restrict__ prev, long prev_img_stride, int r,
                                                      T* __restrict__ out, long out_img_stride, int C, int H, int W,
                                                      int total_rows, int tpr, int rows_per_cta) {
  // a CTA covers `rows_per_cta` output rows (n, channel block, y) with `tpr` threads each (tpr = 16..128, a power of two
  // >= Wo where possible): narrow deep layers get full 128-thread CTAs instead of 12-of-32-lane ones
  const int Ho = H / 2, Wo = W / 2, Cb = C / 8;
  const int row = blockIdx.x * rows_per_cta + (int)(threadIdx.x / tpr);
  if (row >= total_rows) return;
  const int y = row % Ho, cb = (row / Ho) % Cb, n = row / (Ho * Cb);
  const T* in_row = in + (long)n * in_img_stride + ((long)cb * H + 2 * y) * W * 8;
  const T* pv_row = (prev != nullptr && cb * 8 < r) ? prev + (long)n * prev_img_stride + ((long)cb * H + 2 * y) * W * 8 : nullptr;
  T* out_row = out + (long)n * out_img_stride + ((long)cb * Ho + y) * Wo * 8;
  for (int x = threadIdx.x % tpr; x < Wo; x += tpr) {
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const long off = ((long)dy * W + 2 * x + dx) * 8;
        float v[8];
        load8(in_row + off, v);
        if (pv_row != nullptr) {
          float pv[8];
          load8(pv_row + off, pv);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (cb * 8 + c < r) v[c] = pv[c];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    store8(out_row + (long)x * 8, m);
  }
}

// ------------------------------------------------------------------------------------------------
// outc: 1x1 conv C -> 1 (+bias) and sigmoid.  blocked in -> plain [N][H][W] fp32 logits / sigmoid.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) outc_sigmoid_kernel(const T* __restrict__ in, long in_img_stride,
                                                          const float* __restrict__ w, const float* __restrict__ b,
                                                          float* __restrict__ out, float* __restrict__ logit, int C,
                                                          int HW, int N) {
  const long total = (long)N * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = i % HW, n = i / HW;
    float acc = __ldg(b);
    for (int cb = 0; cb < C / 8; ++cb) {
      float v[8];
      load8(in + (long)n * in_img_stride + ((long)cb * HW + p) * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(v[j], __ldg(w + cb * 8 + j), acc);
    }
    if (logit) logit[i] = acc;
    out[i] = 1.f / (1.f + expf(-acc));
  }
}

// ------------------------------------------------------------------------------------------------
// layout transforms: blocked <-> NCHW fp32 (used at the module boundary: up_x features, test hooks)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void blocked_to_nchw_kernel(const T* __restrict__ in, long in_img_stride, float* __restrict__ out, int C,
                                       int HW, int N) {
  const long total = (long)N * C * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = i % HW, c = (i / HW) % C, n = i / ((long)HW * C);
    out[i] = to_f(in[(long)n * in_img_stride + ((long)(c >> 3) * HW + p) * 8 + (c & 7)]);
  }
}
template <typename T>
__global__ void nchw_to_blocked_kernel(const float* __restrict__ in, T* __restrict__ out, long out_img_stride, int C,
                                       int HW, int N) {
  const int Cp = (C + 7) & ~7;
  const long total = (long)N * Cp * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = i & 7;
    const long q = i >> 3;
    const int p = q % HW, cb = (q / HW) % (Cp / 8), n = q / ((long)HW * (Cp / 8));
    const int c = cb * 8 + j;
    from_f(out[(long)n * out_img_stride + ((long)cb * HW + p) * 8 + j], c < C ? in[((long)n * C + c) * HW + p] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// video recurrence for kernels that have no `prev` input: overwrite the first r (<= 8) channels of `dst`
// with those of `src` (both blocked; only channel block 0 is touched).  Unet.py:244, 270.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void splice_channels_kernel(T* __restrict__ dst, long dst_img_stride, const T* __restrict__ src,
                                       long src_img_stride, int r, int HW, int N) {
  const long total = (long)N * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = i % HW, n = i / HW;
    float d[8], s[8];
    load8(dst + (long)n * dst_img_stride + (long)p * 8, d);
    load8(src + (long)n * src_img_stride + (long)p * 8, s);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < r) d[c] = s[c];
    store8(dst + (long)n * dst_img_stride + (long)p * 8, d);
  }
}

// dense element-wise dtype conversion (fp32 <-> bf16); n must be a multiple of 8 (blocked tensors always are)
template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, long n8) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float v[8];
    load8(in + i * 8, v);
    store8(out + i * 8, v);
  }
}

// ================================================================================================
// C ABI
// ================================================================================================
static inline int grid1d(long total, int block = 256) {
  long g = (total + block - 1) / block;
  return (int)(g > 148L * 16 ? 148L * 16 : (g < 1 ? 1 : g));
}

extern "C" int uncl_conv_first(const float* x, const float* w, const float* bias, void* out, long out_img_stride,
                               int N, int H, int W, int C_out, int act, int dtype, cudaStream_t stream) {
  UNCL_REQUIRE(C_out % 8 == 0 && H > 2 && W > 2 && N > 0, "conv_first: bad shape N=%d H=%d W=%d C_out=%d", N, H, W, C_out);
  dim3 grid(ceil_div(W - 2, 32), ceil_div(H - 2, 16), N);
  size_t smem = (size_t)10 * C_out * sizeof(float);
  UNCL_DISPATCH_DTYPE(dtype, T, (conv_first_kernel<T><<<grid, 256, smem, stream>>>(x, w, bias, (T*)out, out_img_stride, H, W, C_out, act)));
  return uncl_check_launch("conv_first");
}

extern "C" int uncl_conv3x3_simt(const void* in, long in_img_stride, const float* w, const float* bias, void* out,
                                 long out_img_stride, int N, int C_in, int H, int W, int C_out, int pad, int act,
                                 int emit_skip, int dtype, cudaStream_t stream) {
  UNCL_REQUIRE(C_in % 8 == 0 && C_out % CS_CO == 0 && (pad == 0 || pad == 2) && N > 0,
               "conv3x3_simt: unsupported C_in=%d C_out=%d pad=%d", C_in, C_out, pad);
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  UNCL_REQUIRE(Ho > 0 && Wo > 0, "conv3x3_simt: empty output");
  dim3 grid(ceil_div(Wo, CS_TX), ceil_div(Ho, CS_TY), N * (C_out / CS_CO));
  UNCL_DISPATCH_DTYPE(dtype, T, (conv3x3_simt_kernel<T><<<grid, 256, 0, stream>>>((const T*)in, in_img_stride, w, bias, (T*)out, out_img_stride, C_in, H, W, C_out, pad, act, emit_skip)));
  return uncl_check_launch("conv3x3_simt");
}

extern "C" int uncl_convT2x2(const void* in, long in_img_stride, const void* prev, long prev_img_stride, int r,
                             const float* w, const float* bias, void* out, long out_img_stride, int N, int C, int H,
                             int W, int H2, int W2, int dtype, cudaStream_t stream) {
  UNCL_REQUIRE(C % 32 == 0 && H2 >= 2 * H && W2 >= 2 * W && N > 0, "convT2x2: unsupported C=%d H2=%d W2=%d", C, H2, W2);
  dim3 grid(ceil_div(H * W, 64), N * (C / 32));
  UNCL_DISPATCH_DTYPE(dtype, T, (convT2x2_kernel<T><<<grid, 256, 0, stream>>>((const T*)in, in_img_stride, (const T*)prev, prev_img_stride, r, w, bias, (T*)out, out_img_stride, C, H, W, H2, W2)));
  return uncl_check_launch("convT2x2");
}

extern "C" int uncl_maxpool2(const void* in, long in_img_stride, const void* prev, long prev_img_stride, int r,
                             void* out, long out_img_stride, int N, int C, int H, int W, int dtype,
                             cudaStream_t stream) {
  UNCL_REQUIRE(C % 8 == 0 && H >= 2 && W >= 2 && N > 0, "maxpool2: bad shape");
  const long rows = (long)N * (C / 8) * (H / 2);
  UNCL_REQUIRE(rows < (1L << 31), "maxpool2: too many rows");
  const int wo = W / 2;
  const int tpr = wo <= 16 ? 16 : (wo <= 32 ? 32 : (wo <= 64 ? 64 : 128));
  const int rows_per_cta = 128 / tpr;
  const unsigned grid = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
  UNCL_DISPATCH_DTYPE(dtype, T, (maxpool2_kernel<T><<<grid, 128, 0, stream>>>((const T*)in, in_img_stride, (const T*)prev, prev_img_stride, r, (T*)out, out_img_stride, C, H, W, (int)rows, tpr, rows_per_cta)));
  return uncl_check_launch("maxpool2");
}

extern "C" int uncl_outc_sigmoid(const void* in, long in_img_stride, const float* w, const float* b, float* out,
                                 float* logit, int N, int C, int HW, int dtype, cudaStream_t stream) {
  UNCL_REQUIRE(C % 8 == 0 && N > 0 && HW > 0, "outc_sigmoid: bad shape");
  UNCL_DISPATCH_DTYPE(dtype, T, (outc_sigmoid_kernel<T><<<grid1d((long)N * HW), 256, 0, stream>>>((const T*)in, in_img_stride, w, b, out, logit, C, HW, N)));
  return uncl_check_launch("outc_sigmoid");
}

extern "C" int uncl_blocked_to_nchw(const void* in, long in_img_stride, float* out, int N, int C, int HW, int dtype,
                                    cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C > 0 && HW > 0, "blocked_to_nchw: bad shape");
  UNCL_DISPATCH_DTYPE(dtype, T, (blocked_to_nchw_kernel<T><<<grid1d((long)N * C * HW), 256, 0, stream>>>((const T*)in, in_img_stride, out, C, HW, N)));
  return uncl_check_launch("blocked_to_nchw");
}

extern "C" int uncl_nchw_to_blocked(const float* in, void* out, long out_img_stride, int N, int C, int HW, int dtype,
                                    cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C > 0 && HW > 0, "nchw_to_blocked: bad shape");
  const long total = (long)N * ((C + 7) & ~7) * HW;
  UNCL_DISPATCH_DTYPE(dtype, T, (nchw_to_blocked_kernel<T><<<grid1d(total), 256, 0, stream>>>(in, (T*)out, out_img_stride, C, HW, N)));
  return uncl_check_launch("nchw_to_blocked");
}

extern "C" int uncl_splice_channels(void* dst, long dst_img_stride, const void* src, long src_img_stride, int r, int N,
                                    int HW, int dtype, cudaStream_t stream) {
  UNCL_REQUIRE(r >= 1 && r <= 8 && N > 0 && HW > 0, "splice_channels: r must be in 1..8");
  UNCL_DISPATCH_DTYPE(dtype, T, (splice_channels_kernel<T><<<grid1d((long)N * HW), 256, 0, stream>>>((T*)dst, dst_img_stride, (const T*)src, src_img_stride, r, HW, N)));
  return uncl_check_launch("splice_channels");
}

// fp32 blocked [N][C/8][HW][8] (image stride in_img_stride) -> bf16 blocked [N][3C/8][HW][8] = [hi | hi | lo],
// hi = bf16(x), lo = bf16(x - hi): the A operand of a three-term bf16 product x_hi.w_hi + x_hi.w_lo + x_lo.w_hi
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ in, long in_img_stride,
                                                        bf16* __restrict__ out, long out_img_stride, int Cb, long HW,
                                                        long total) {
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const long p = i % HW;
    const int cb = (int)((i / HW) % Cb);
    const long n = i / (HW * Cb);
    float v[8], hi[8], lo[8];
    load8(in + n * in_img_stride + ((long)cb * HW + p) * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
      lo[j] = v[j] - hi[j];
    }
    bf16* o = out + n * out_img_stride + ((long)cb * HW + p) * 8;
    store8(o, hi);
    store8(o + (long)Cb * HW * 8, hi);
    store8(o + (long)2 * Cb * HW * 8, lo);
  }
}

extern "C" int uncl_split_bf16(const float* in, long in_img_stride, void* out, long out_img_stride, int N, int C, long HW,
                               cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && HW > 0 && in && out && in_img_stride % 8 == 0 && out_img_stride % 8 == 0,
               "split_bf16: bad arguments");
  const long total = (long)N * (C / 8) * HW;
  split_bf16_kernel<<<grid1d(total), 256, 0, stream>>>(in, in_img_stride, reinterpret_cast<bf16*>(out), out_img_stride, C / 8, HW,
                                                       total);
  return uncl_check_launch("split_bf16");
}

extern "C" int uncl_convert(const void* in, int in_dtype, void* out, int out_dtype, long n, cudaStream_t stream) {
  UNCL_REQUIRE(n > 0 && n % 8 == 0, "convert: element count must be a positive multiple of 8");
  const long n8 = n / 8;
  if (in_dtype == UNCL_F32 && out_dtype == UNCL_BF16)
    convert_kernel<float, bf16><<<grid1d(n8), 256, 0, stream>>>((const float*)in, (bf16*)out, n8);
  else if (in_dtype == UNCL_BF16 && out_dtype == UNCL_F32)
    convert_kernel<bf16, float><<<grid1d(n8), 256, 0, stream>>>((const bf16*)in, (float*)out, n8);
  else
    return uncl_set_error(UNCL_EINVAL, "convert: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
  return uncl_check_launch("convert");
}
