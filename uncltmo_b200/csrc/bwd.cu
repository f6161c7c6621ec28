// Backward kernels of the generator (fp32, C8-blocked tensors): ReLU mask + bias gradient, 3x3 weight gradient,
// max-pool / skip-concat / ConvTranspose-k2 helpers, pointwise weight gradient, GELU, graph aggregation, out conv.
// Data gradients of the 3x3 convs reuse the forward conv kernels with flipped / transposed weights (dgrad of a
// correlation with pad p is a correlation with pad 2-p); data gradients of the pointwise convs reuse pw_conv.
//
// Reference: autograd of models/unet_multi_filters/unet_parts.py (conv / ConvTranspose / MaxPool / concat operators),
// Unet_singleFrame.py:20-99 (FFN, GCNBlock), gcn_lib/torch_vertex.py:13-30 (MRConv2d).
#include "common.cuh"

namespace {

inline int cap_grid(long total, int block, int per_sm) {
  long g = (total + block - 1) / block;
  const long cap = 148L * per_sm;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

// ---------------------------------------------------------------------------------------------------------
// dZ = dY * (Y > 0) in place (optional) and db[c] += sum over images and pixels of dZ
// grid: (pixel chunks, C/8, N)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relu_bwd_bias_kernel(float* __restrict__ dY, const float* __restrict__ Y,
                                                           long y_img_stride, float* __restrict__ db, int C, int HW,
                                                           int apply_relu) {
  __shared__ float red[8][8];
  const int cb = blockIdx.y, n = blockIdx.z;
  float* d = dY + ((long)n * (C / 8) + cb) * HW * 8;
  const float* y = Y ? Y + (long)n * y_img_stride + (long)cb * HW * 8 : nullptr;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
    float g[8];
    load8(d + (long)p * 8, g);
    if (apply_relu) {
      float v[8];
      load8(y + (long)p * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = v[j] > 0.f ? g[j] : 0.f;
      store8(d + (long)p * 8, g);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += g[j];
  }
  if (db == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = warp_sum(s[j]);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[wid][j] = s[j];
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(db + cb * 8 + threadIdx.x, t);
  }
}

// out-of-place variant: dZ = dY * (Y > 0) written as fp32 or bf16 (the tensor-core dgrad / wgrad operand), db += sum dZ
template <typename TO>
__global__ void __launch_bounds__(256) relu_bwd_bias_out_kernel(const float* __restrict__ dY, const float* __restrict__ Y,
                                                               long y_img_stride, TO* __restrict__ dZ,
                                                               float* __restrict__ db, int C, int HW, int apply_relu) {
  __shared__ float red[8][8];
  const int cb = blockIdx.y, n = blockIdx.z;
  const long base = ((long)n * (C / 8) + cb) * HW * 8;
  const float* y = Y ? Y + (long)n * y_img_stride + (long)cb * HW * 8 : nullptr;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
    float g[8];
    load8(dY + base + (long)p * 8, g);
    if (apply_relu) {
      float v[8];
      load8(y + (long)p * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = v[j] > 0.f ? g[j] : 0.f;
    }
    store8(dZ + base + (long)p * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += g[j];
  }
  if (db == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = warp_sum(s[j]);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[wid][j] = s[j];
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(db + cb * 8 + threadIdx.x, t);
  }
}

// ---------------------------------------------------------------------------------------------------------
// 3x3 weight gradient: dW9[t][ci][co] += sum_{n,y,x} X[n, ci, y+ky-pad, x+kx-pad] * dZ[n, co, y, x]
// CTA: one input channel block (8 ci) x CO_T output channels x 9 taps, over a strided subset of 8x32 pixel tiles.
// thread: 1 ci x CPT co x 9 taps.
// ---------------------------------------------------------------------------------------------------------
template <int CPT>
__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(const float* __restrict__ X, long x_img_stride,
                                                           const float* __restrict__ dZ, float* __restrict__ dW,
                                                           int N, int C_in, int H, int W, int C_out, int pad) {
  constexpr int CO_T = 32 * CPT;
  constexpr int TY = CPT == 4 ? 4 : 8, TX = 32, NPX = TY * TX, LDZ = CO_T + 4;
  extern __shared__ __align__(16) float wg_smem[];
  float (*s_x)[TY + 2][TX + 2] = reinterpret_cast<float (*)[TY + 2][TX + 2]>(wg_smem);       // [8][TY+2][TX+2]
  float* s_dz = wg_smem + ((8 * (TY + 2) * (TX + 2) + 3) & ~3);                               // [NPX][LDZ]
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  const int cib = blockIdx.x, cot = blockIdx.y;
  const int tiles_x = (Wo + TX - 1) / TX, tiles_y = (Ho + TY - 1) / TY;
  const int tiles = N * tiles_y * tiles_x;
  const int ci = threadIdx.x >> 5, cl = threadIdx.x & 31;  // warp = one ci, lanes = co groups
  float acc[9][CPT];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[t][j] = 0.f;

  for (int tile = blockIdx.z; tile < tiles; tile += gridDim.z) {
    const int n = tile / (tiles_y * tiles_x), r = tile % (tiles_y * tiles_x);
    const int oy0 = (r / tiles_x) * TY, ox0 = (r % tiles_x) * TX;
    __syncthreads();
    const float* xn = X + (long)n * x_img_stride + (long)cib * H * W * 8;
    for (int p = threadIdx.x; p < (TY + 2) * (TX + 2); p += 256) {
      const int ly = p / (TX + 2), lx = p % (TX + 2);
      const int iy = oy0 + ly - pad, ix = ox0 + lx - pad;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) load8(xn + ((long)iy * W + ix) * 8, v);
#pragma unroll
      for (int c = 0; c < 8; ++c) s_x[c][ly][lx] = v[c];
    }
    const float* dn0 = dZ + ((long)n * (C_out / 8) + (long)cot * (CO_T / 8)) * Ho * Wo * 8;
    for (int i = threadIdx.x; i < NPX * (CO_T / 8); i += 256) {
      const int p = i % NPX, b = i / NPX;
      const int oy = oy0 + p / TX, ox = ox0 + p % TX;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (oy < Ho && ox < Wo) load8(dn0 + (long)b * Ho * Wo * 8 + ((long)oy * Wo + ox) * 8, v);
      *reinterpret_cast<float4*>(&s_dz[p * LDZ + b * 8]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&s_dz[p * LDZ + b * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
#pragma unroll 4
    for (int p = 0; p < NPX; ++p) {
      const int ly = p / TX, lx = p % TX;
      float dz[CPT];
      if constexpr (CPT == 4) {
        const float4 q = *reinterpret_cast<const float4*>(&s_dz[p * LDZ + cl * 4]);
        dz[0] = q.x; dz[1] = q.y; dz[2] = q.z; dz[3] = q.w;
      } else if constexpr (CPT == 2) {
        const float2 q = *reinterpret_cast<const float2*>(&s_dz[p * LDZ + cl * 2]);
        dz[0] = q.x; dz[1] = q.y;
      } else {
        dz[0] = s_dz[p * LDZ + cl];
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float xv = s_x[ci][ly + t / 3][lx + t % 3];
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[t][j] = fmaf(xv, dz[j], acc[t][j]);
      }
    }
  }
  const int cig = cib * 8 + ci;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < CPT; ++j)
      atomicAdd(dW + ((long)t * C_in + cig) * C_out + cot * CO_T + cl * CPT + j, acc[t][j]);
}

template <int CPT>
int launch_wgrad(dim3 grid, const float* X, long x_img_stride, const float* dZ, float* dW9, int N, int C_in, int H,
                 int W, int C_out, int pad, cudaStream_t stream) {
  constexpr int TY = CPT == 4 ? 4 : 8;
  const size_t smem = (size_t)(((8 * (TY + 2) * 34 + 3) & ~3) + TY * 32 * (32 * CPT + 4)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_kernel<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "conv3x3_wgrad: smem attr: %s", cudaGetErrorString(e));
  conv3x3_wgrad_kernel<CPT><<<grid, 256, smem, stream>>>(X, x_img_stride, dZ, dW9, N, C_in, H, W, C_out, pad);
  return uncl_check_launch("conv3x3_wgrad");
}

// first conv (C_in = 1): dW[t][c] += sum x[n, y+ky, x+kx] * dZ[n, c, y, x]
__global__ void __launch_bounds__(256) conv_first_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dZ,
                                                              float* __restrict__ dW, int N, int H, int W, int C) {
  __shared__ float red[8][9];
  const int Ho = H - 2, Wo = W - 2, HWo = Ho * Wo;
  const int c = blockIdx.y;
  float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long total = (long)N * HWo;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int p = i % HWo, n = i / HWo;
    const int oy = p / Wo, ox = p % Wo;
    const float g = dZ[((long)n * (C / 8) + (c >> 3)) * HWo * 8 + (long)p * 8 + (c & 7)];
    const float* xi = x + (long)n * H * W + (long)oy * W + ox;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = fmaf(__ldg(xi + (t / 3) * W + t % 3), g, acc[t]);
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = warp_sum(acc[t]);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int t = 0; t < 9; ++t) red[wid][t] = acc[t];
  __syncthreads();
  if (threadIdx.x < 9) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(dW + threadIdx.x * C + c, s);
  }
}

// ---------------------------------------------------------------------------------------------------------
// MaxPool2d(2) backward: the first maximum of each 2x2 window (row-major scan, as PyTorch) receives the gradient
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ X, long x_img_stride,
                                                          const float* __restrict__ dP, float* __restrict__ dX, int C,
                                                          int H, int W, int N) {
  const int Ho = H / 2, Wo = W / 2, Cb = C / 8;
  const long total = (long)N * Cb * H * W;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int x = i % W, y = (i / W) % H, cb = (i / ((long)W * H)) % Cb, n = i / ((long)W * H * Cb);
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int py = y >> 1, px = x >> 1;
    if (py < Ho && px < Wo) {
      const float* xb = X + (long)n * x_img_stride + (long)cb * H * W * 8;
      float v[4][8], dp[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) load8(xb + ((long)(2 * py + (k >> 1)) * W + 2 * px + (k & 1)) * 8, v[k]);
      load8(dP + (((long)n * Cb + cb) * Ho * Wo + (long)py * Wo + px) * 8, dp);
      const int me = ((y & 1) << 1) | (x & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int best = 0;
        float bv = v[0][j];
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (v[k][j] > bv) { bv = v[k][j]; best = k; }
        g[j] = best == me ? dp[j] : 0.f;
      }
    }
    store8(dX + (((long)n * Cb + cb) * H * W + (long)y * W + x) * 8, g);
  }
}

// ---------------------------------------------------------------------------------------------------------
// skip concat: cat = [x2 | x1 | x2^2 | sqrt(x2 + 1e-8)] and its backward  (unet_parts.py:319-322)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) skip_concat_fwd_kernel(const float* __restrict__ x2, long x2_img_stride,
                                                             const float* __restrict__ x1, float* __restrict__ cat,
                                                             int C, int HW, int N) {
  const int Cb = C / 8;
  const long total = (long)N * Cb * HW;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int p = i % HW, cb = (i / HW) % Cb, n = i / ((long)HW * Cb);
    float a[8], b[8], s[8];
    load8(x2 + (long)n * x2_img_stride + ((long)cb * HW + p) * 8, a);
    load8(x1 + (((long)n * Cb + cb) * HW + p) * 8, b);
    float* o = cat + (long)n * 4 * C * HW + ((long)cb * HW + p) * 8;
    store8(o, a);
    store8(o + (long)Cb * HW * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = a[j] * a[j];
    store8(o + (long)2 * Cb * HW * 8, s);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = sqrtf(a[j] + 1e-8f);
    store8(o + (long)3 * Cb * HW * 8, s);
  }
}
__global__ void __launch_bounds__(256) skip_concat_bwd_kernel(const float* __restrict__ dcat,
                                                             const float* __restrict__ x2, long x2_img_stride,
                                                             float* __restrict__ dx2, float* __restrict__ dx1, int C,
                                                             int HW, int N) {
  const int Cb = C / 8;
  const long total = (long)N * Cb * HW;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int p = i % HW, cb = (i / HW) % Cb, n = i / ((long)HW * Cb);
    const float* g = dcat + (long)n * 4 * C * HW + ((long)cb * HW + p) * 8;
    float a[8], g0[8], g1[8], g2[8], g3[8], o[8];
    load8(x2 + (long)n * x2_img_stride + ((long)cb * HW + p) * 8, a);
    load8(g, g0);
    load8(g + (long)Cb * HW * 8, g1);
    load8(g + (long)2 * Cb * HW * 8, g2);
    load8(g + (long)3 * Cb * HW * 8, g3);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = g0[j] + 2.f * a[j] * g2[j] + 0.5f * g3[j] / sqrtf(a[j] + 1e-8f);
    store8(dx2 + (((long)n * Cb + cb) * HW + p) * 8, o);
    store8(dx1 + (((long)n * Cb + cb) * HW + p) * 8, g1);
  }
}

// ---------------------------------------------------------------------------------------------------------
// ConvTranspose k2 s2 backward helper: space-to-depth of the (replicate-padded) output gradient,
//   out[n, pos*C + co, y, x] = sum over the padded copies of dY[n, co, 2y+dy, 2x+dx]   (pos = dy*2+dx)
// after which dX = pointwise conv with W^T, dW = pointwise weight gradient, db = column sums.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) convT2x2_s2d_kernel(const float* __restrict__ dY, float* __restrict__ out,
                                                          bf16* __restrict__ out_b, int C, int H, int W, int H2, int W2,
                                                          int N) {
  const int Cb = C / 8;
  const long total = (long)N * 4 * Cb * H * W;
  const int padT = (H2 - 2 * H) / 2, padL = (W2 - 2 * W) / 2;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int x = i % W, y = (i / W) % H;
    const int cb4 = (i / ((long)W * H)) % (4 * Cb), n = i / ((long)W * H * 4 * Cb);
    const int pos = cb4 / Cb, cb = cb4 % Cb;
    const int Y = 2 * y + (pos >> 1), X = 2 * x + (pos & 1);
    const int y_lo = (Y == 0) ? 0 : Y + padT, y_hi = (Y == 2 * H - 1) ? H2 - 1 : Y + padT;
    const int x_lo = (X == 0) ? 0 : X + padL, x_hi = (X == 2 * W - 1) ? W2 - 1 : X + padL;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* d = dY + ((long)n * Cb + cb) * H2 * W2 * 8;
    for (int yy = y_lo; yy <= y_hi; ++yy)
      for (int xx = x_lo; xx <= x_hi; ++xx) {
        float v[8];
        load8(d + ((long)yy * W2 + xx) * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += v[j];
      }
    const long o = (((long)n * 4 * Cb + cb4) * H * W + (long)y * W + x) * 8;
    store8(out + o, s);
    if (out_b != nullptr) store8(out_b + o, s);
  }
}

// ---------------------------------------------------------------------------------------------------------
// pointwise weight gradient: dW[g][ci][co] += sum_pix X[pix, ci] * dZ[pix, co]
// CTA: 64 ci x 64 co, K chunks of 32 pixels, split over pixels with atomics.  thread: 4 ci x 4 co.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pw_wgrad_kernel(const float* __restrict__ X, const float* __restrict__ dZ,
                                                      float* __restrict__ dW, int C_in, int C_out, int groups, int HW,
                                                      int N) {
  __shared__ __align__(16) float s_x[32][64 + 4];
  __shared__ __align__(16) float s_d[32][64 + 4];
  const int cin_g = C_in / groups, cout_g = C_out / groups;
  const int tiles_i = (cin_g + 63) / 64, tiles_j = (cout_g + 63) / 64;
  const int g = blockIdx.x / tiles_i;
  const int ci0 = (blockIdx.x % tiles_i) * 64, co0 = blockIdx.y * 64;   // within the group
  (void)tiles_j;
  const int P = N * HW;
  const int ti = (threadIdx.x & 15) * 4, tj = (threadIdx.x >> 4) * 4;
  float acc[4][4] = {};
  for (int p0 = blockIdx.z * 32; p0 < P; p0 += gridDim.z * 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 8; i += 256) {   // 32 pixels x 8 channel blocks of 8
      const int px = i & 31, b = i >> 5;
      const int p = p0 + px;
      float vx[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vd[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p < P) {
        const int n = p / HW, q = p - n * HW;
        if (ci0 + b * 8 < cin_g) load8(X + ((long)n * (C_in / 8) + ((g * cin_g + ci0) >> 3) + b) * HW * 8 + (long)q * 8, vx);
        if (co0 + b * 8 < cout_g) load8(dZ + ((long)n * (C_out / 8) + ((g * cout_g + co0) >> 3) + b) * HW * 8 + (long)q * 8, vd);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { s_x[px][b * 8 + j] = vx[j]; s_d[px][b * 8 + j] = vd[j]; }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&s_x[k][ti]);
      const float4 b = *reinterpret_cast<const float4*>(&s_d[k][tj]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (ci0 + ti + i < cin_g && co0 + tj + j < cout_g)
        atomicAdd(dW + ((long)g * cin_g + ci0 + ti + i) * cout_g + co0 + tj + j, acc[i][j]);
}

// ---------------------------------------------------------------------------------------------------------
// elementwise helpers of the graph block
// ---------------------------------------------------------------------------------------------------------
__global__ void gelu_fwd_kernel(const float* __restrict__ u, float* __restrict__ g, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    g[i] = apply_act(u[i], UNCL_ACT_GELU);
}
__global__ void gelu_bwd_kernel(const float* __restrict__ u, const float* __restrict__ dg, float* __restrict__ du, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float x = u[i];
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    du[i] = dg[i] * (cdf + x * pdf);
  }
}
__global__ void scale_rows_kernel(float* __restrict__ x, const float* __restrict__ scale, long per_image, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    x[i] *= scale[i / per_image];
}
__global__ void batch_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int N, long M) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += x[(long)n * M + i];
    out[i] = s;
  }
}

// MRConv aggregation backward.  z = interleave(y, agg), agg[c,i] = max_k (y[c, idx[i,k]] - y[c,i]).
//   dy[c,i] += dz_y[c,i] - dagg[c,i];  dy[c, idx[i,k*]] += dagg[c,i]   (k* = first arg-max)
// dy must be zero-initialised.  One thread per (n, i, channel block).
__global__ void __launch_bounds__(256) gcn_agg_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                                         const int* __restrict__ idx, float* __restrict__ dy, int C,
                                                         int N) {
  const int Cb = C / 8;
  const long total = (long)N * 144 * Cb;
  for (long t = (long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long)gridDim.x * 256) {
    const int cb = t % Cb, i = (t / Cb) % 144, n = t / ((long)Cb * 144);
    const float* yn = y + (long)n * C * 144;
    const float* dzn = dz + (long)n * 2 * C * 144;
    float* dyn = dy + (long)n * C * 144;
    float yi[8], lo[8], hi[8];
    load8(yn + ((long)cb * 144 + i) * 8, yi);
    load8(dzn + ((long)(2 * cb) * 144 + i) * 8, lo);
    load8(dzn + ((long)(2 * cb + 1) * 144 + i) * 8, hi);
    const float dyv[8] = {lo[0], lo[2], lo[4], lo[6], hi[0], hi[2], hi[4], hi[6]};
    const float dag[8] = {lo[1], lo[3], lo[5], lo[7], hi[1], hi[3], hi[5], hi[7]};
    float best[8];
    int bk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bk[j] = 0; }
    const int* id = idx + ((long)n * 144 + i) * 9;
    for (int k = 0; k < 9; ++k) {
      float yj[8];
      load8(yn + ((long)cb * 144 + id[k]) * 8, yj);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (yj[j] - yi[j] > best[j]) { best[j] = yj[j] - yi[j]; bk[j] = k; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(dyn + ((long)cb * 144 + i) * 8 + j, dyv[j] - dag[j]);
      atomicAdd(dyn + ((long)cb * 144 + id[bk[j]]) * 8 + j, dag[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// out conv (1x1, C -> 1) + sigmoid backward
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) outc_sigmoid_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ out,
                                                              const float* __restrict__ up, long up_img_stride,
                                                              const float* __restrict__ w, float* __restrict__ d_up,
                                                              float* __restrict__ dw, float* __restrict__ db, int C,
                                                              int HW, int N) {
  extern __shared__ float s_acc[];  // [C + 1]
  for (int i = threadIdx.x; i <= C; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const long total = (long)N * HW;
  float lb = 0.f;
  for (long base = (long)blockIdx.x * 256; base < total; base += (long)gridDim.x * 256) {
    const long i = base + threadIdx.x;
    const bool valid = i < total;
    const int p = valid ? (int)(i % HW) : 0, n = valid ? (int)(i / HW) : 0;
    const float o = valid ? out[i] : 0.f;
    const float dl = valid ? d_out[i] * o * (1.f - o) : 0.f;
    lb += dl;
    for (int cb = 0; cb < C / 8; ++cb) {
      float u[8], g[8];
      load8(up + (long)n * up_img_stride + ((long)cb * HW + p) * 8, u);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[j] = dl * __ldg(w + cb * 8 + j);
        const float c = warp_sum(dl * u[j]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[cb * 8 + j], c);
      }
      if (valid) store8(d_up + (((long)n * (C / 8) + cb) * HW + p) * 8, g);
    }
  }
  lb = warp_sum(lb);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[C], lb);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) atomicAdd(dw + i, s_acc[i]);
  if (threadIdx.x == 0) atomicAdd(db, s_acc[C]);
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int uncl_relu_bwd_bias(float* dY, const float* Y, long y_img_stride, float* db, int N, int C, int HW,
                                  int apply_relu, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && HW > 0 && (!apply_relu || Y), "relu_bwd_bias: bad arguments");
  int gx = (HW + 255) / 256;
  if (gx > 64) gx = 64;
  relu_bwd_bias_kernel<<<dim3(gx, C / 8, N), 256, 0, stream>>>(dY, Y, y_img_stride, db, C, HW, apply_relu);
  return uncl_check_launch("relu_bwd_bias");
}

extern "C" int uncl_relu_bwd_bias_out(const float* dY, const float* Y, long y_img_stride, void* dZ, int dz_dtype, float* db,
                                      int N, int C, int HW, int apply_relu, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && HW > 0 && (!apply_relu || Y) && dZ, "relu_bwd_bias_out: bad arguments");
  int gx = (HW + 255) / 256;
  if (gx > 64) gx = 64;
  UNCL_DISPATCH_DTYPE(dz_dtype, T, (relu_bwd_bias_out_kernel<T><<<dim3(gx, C / 8, N), 256, 0, stream>>>(dY, Y, y_img_stride, (T*)dZ, db, C, HW, apply_relu)));
  return uncl_check_launch("relu_bwd_bias_out");
}

extern "C" int uncl_conv3x3_wgrad(const float* X, long x_img_stride, const float* dZ, float* dW9, int N, int C_in,
                                  int H, int W, int C_out, int pad, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C_in % 8 == 0 && C_out % 32 == 0 && (pad == 0 || pad == 2), "conv3x3_wgrad: unsupported shape");
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  const int cpt = C_out % 128 == 0 ? 4 : (C_out % 64 == 0 ? 2 : 1);
  const int ty = cpt == 4 ? 4 : 8;
  const int tiles = N * ((Ho + ty - 1) / ty) * ((Wo + 31) / 32);
  const int base = (C_in / 8) * (C_out / (32 * cpt));
  int split = (148 * 4 + base - 1) / base;
  if (split > tiles) split = tiles;
  if (split < 1) split = 1;
  dim3 grid(C_in / 8, C_out / (32 * cpt), split);
  if (cpt == 4) return launch_wgrad<4>(grid, X, x_img_stride, dZ, dW9, N, C_in, H, W, C_out, pad, stream);
  if (cpt == 2) return launch_wgrad<2>(grid, X, x_img_stride, dZ, dW9, N, C_in, H, W, C_out, pad, stream);
  return launch_wgrad<1>(grid, X, x_img_stride, dZ, dW9, N, C_in, H, W, C_out, pad, stream);
}

extern "C" int uncl_conv_first_wgrad(const float* x, const float* dZ, float* dW, int N, int H, int W, int C,
                                     cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && H > 2 && W > 2, "conv_first_wgrad: bad shape");
  conv_first_wgrad_kernel<<<dim3(32, C), 256, 0, stream>>>(x, dZ, dW, N, H, W, C);
  return uncl_check_launch("conv_first_wgrad");
}

extern "C" int uncl_maxpool2_bwd(const float* X, long x_img_stride, const float* dP, float* dX, int N, int C, int H,
                                 int W, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && H >= 2 && W >= 2, "maxpool2_bwd: bad shape");
  maxpool2_bwd_kernel<<<cap_grid((long)N * (C / 8) * H * W, 256, 8), 256, 0, stream>>>(X, x_img_stride, dP, dX, C, H, W, N);
  return uncl_check_launch("maxpool2_bwd");
}

extern "C" int uncl_skip_concat_fwd(const float* x2, long x2_img_stride, const float* x1, float* cat, int N, int C,
                                    int HW, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && HW > 0, "skip_concat_fwd: bad shape");
  skip_concat_fwd_kernel<<<cap_grid((long)N * (C / 8) * HW, 256, 8), 256, 0, stream>>>(x2, x2_img_stride, x1, cat, C, HW, N);
  return uncl_check_launch("skip_concat_fwd");
}

extern "C" int uncl_skip_concat_bwd(const float* dcat, const float* x2, long x2_img_stride, float* dx2, float* dx1,
                                    int N, int C, int HW, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && HW > 0, "skip_concat_bwd: bad shape");
  skip_concat_bwd_kernel<<<cap_grid((long)N * (C / 8) * HW, 256, 8), 256, 0, stream>>>(dcat, x2, x2_img_stride, dx2, dx1, C, HW, N);
  return uncl_check_launch("skip_concat_bwd");
}

extern "C" int uncl_convT2x2_s2d(const float* dY, float* out, void* out_bf16, int N, int C, int H, int W, int H2, int W2,
                                 cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && H2 >= 2 * H && W2 >= 2 * W, "convT2x2_s2d: bad shape");
  convT2x2_s2d_kernel<<<cap_grid((long)N * 4 * (C / 8) * H * W, 256, 8), 256, 0, stream>>>(dY, out, (bf16*)out_bf16, C, H, W, H2, W2, N);
  return uncl_check_launch("convT2x2_s2d");
}

extern "C" int uncl_pw_wgrad(const float* X, const float* dZ, float* dW, int N, int C_in, int C_out, int groups, int HW,
                             cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && groups > 0 && C_in % (8 * groups) == 0 && C_out % (8 * groups) == 0, "pw_wgrad: unsupported shape");
  const int P = N * HW;
  const int cin_g = C_in / groups, cout_g = C_out / groups;
  const int tiles_i = (cin_g + 63) / 64, tiles_j = (cout_g + 63) / 64;
  const int base = groups * tiles_i * tiles_j;
  int split = (148 * 2 + base - 1) / base;
  if (split > (P + 31) / 32) split = (P + 31) / 32;
  pw_wgrad_kernel<<<dim3(groups * tiles_i, tiles_j, split), 256, 0, stream>>>(X, dZ, dW, C_in, C_out, groups, HW, N);
  return uncl_check_launch("pw_wgrad");
}

extern "C" int uncl_gelu_fwd(const float* u, float* g, long n, cudaStream_t stream) {
  UNCL_REQUIRE(n > 0, "gelu_fwd: empty");
  gelu_fwd_kernel<<<cap_grid(n, 256, 8), 256, 0, stream>>>(u, g, n);
  return uncl_check_launch("gelu_fwd");
}
extern "C" int uncl_gelu_bwd(const float* u, const float* dg, float* du, long n, cudaStream_t stream) {
  UNCL_REQUIRE(n > 0, "gelu_bwd: empty");
  gelu_bwd_kernel<<<cap_grid(n, 256, 8), 256, 0, stream>>>(u, dg, du, n);
  return uncl_check_launch("gelu_bwd");
}
extern "C" int uncl_scale_rows(float* x, const float* scale, int N, long per_image, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && per_image > 0, "scale_rows: empty");
  scale_rows_kernel<<<cap_grid((long)N * per_image, 256, 8), 256, 0, stream>>>(x, scale, per_image, (long)N * per_image);
  return uncl_check_launch("scale_rows");
}
extern "C" int uncl_batch_sum(const float* x, float* out, int N, long M, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && M > 0, "batch_sum: empty");
  batch_sum_kernel<<<cap_grid(M, 256, 8), 256, 0, stream>>>(x, out, N, M);
  return uncl_check_launch("batch_sum");
}

extern "C" int uncl_gcn_agg_bwd(const float* dz, const float* y, const int* idx, float* dy, int N, int C,
                                cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0, "gcn_agg_bwd: bad shape");
  gcn_agg_bwd_kernel<<<cap_grid((long)N * 144 * (C / 8), 256, 8), 256, 0, stream>>>(dz, y, idx, dy, C, N);
  return uncl_check_launch("gcn_agg_bwd");
}

extern "C" int uncl_outc_sigmoid_bwd(const float* d_out, const float* out, const float* up, long up_img_stride,
                                     const float* w, float* d_up, float* dw, float* db, int N, int C, int HW,
                                     cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && C % 8 == 0 && HW > 0, "outc_sigmoid_bwd: bad shape");
  outc_sigmoid_bwd_kernel<<<cap_grid((long)N * HW, 256, 4), 256, (C + 1) * sizeof(float), stream>>>(d_out, out, up, up_img_stride, w, d_up, dw, db, C, HW, N);
  return uncl_check_launch("outc_sigmoid_bwd");
}
