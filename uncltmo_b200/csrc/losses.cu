// Training-loss kernels (forward) and the SimpleDiscriminator forward.  All operate on dense fp32 planes
// [M][H][W] (single-channel images / feature planes) - these are HBM-bound stencils and reductions.
//
// Reference: models/struct_loss.py:46-104 (pyramid structural loss), models/Discriminator.py:49-126
// (SimpleDiscriminator, ContrastExtracter), GanTrainerImg.py:219-229 (contrastive_D_loss), :410-439 (nce),
// :308-313 (mean / contrast L1), GanTrainer.py:669-682 (L_TV).
#include "common.cuh"

namespace {

__device__ __forceinline__ void atomic_add_block(float v, float* dst, float* red) {
  v = block_sum(v, red);
  if (threadIdx.x == 0) atomicAdd(dst, v);
}

// ---------------------------------------------------------------------------------------------------------
// per-plane mean and mean local variance under the 11x11 gaussian (sigma 1.5, valid):  E_g[x^2] - E_g[x]^2
// tile 32x32 outputs, halo 10, separable filter in shared memory.  sums[2*m] += sum x, sums[2*m+1] += sum var
// ---------------------------------------------------------------------------------------------------------
// fspecial_gauss(11, 1.5) is separable: g2[i][j] = g[i] g[j], g = exp(-t^2 / 4.5) / sum (statically initialised: the library
// keeps no mutable global and uploads nothing at run time)
__constant__ float c_gauss11[11] = {1.028380084e-03f, 7.598758135e-03f, 3.600077213e-02f, 1.093606895e-01f, 2.130055377e-01f, 2.660117249e-01f,
                                   2.130055377e-01f, 1.093606895e-01f, 3.600077213e-02f, 7.598758135e-03f, 1.028380084e-03f};

__global__ void __launch_bounds__(256) plane_contrast_kernel(const float* __restrict__ x, long plane_stride, int H,
                                                            int W, float* __restrict__ sums) {
  __shared__ float s_x[42][43];
  __shared__ float s_h1[42][33];
  __shared__ float s_h2[42][33];
  __shared__ float red[33];
  const int m = blockIdx.z;
  const float* xp = x + (long)m * plane_stride;
  const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * 32;
  float msum = 0.f;
  for (int i = threadIdx.x; i < 42 * 42; i += 256) {
    const int ly = i / 42, lx = i % 42;
    const int gy = oy0 + ly, gx = ox0 + lx;
    float v = 0.f;
    if (gy < H && gx < W) {
      v = __ldg(xp + (long)gy * W + gx);
      if (ly < 32 && lx < 32) msum += v;  // every pixel belongs to exactly one tile's top-left 32x32
    }
    s_x[ly][lx] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 42 * 32; i += 256) {
    const int ly = i / 32, lx = i % 32;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float v = s_x[ly][lx + k];
      a = fmaf(c_gauss11[k], v, a);
      b = fmaf(c_gauss11[k], v * v, b);
    }
    s_h1[ly][lx] = a;
    s_h2[ly][lx] = b;
  }
  __syncthreads();
  float vsum = 0.f;
  const int Ho = H - 10, Wo = W - 10;
  for (int i = threadIdx.x; i < 32 * 32; i += 256) {
    const int ly = i / 32, lx = i % 32;
    if (oy0 + ly < Ho && ox0 + lx < Wo) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        a = fmaf(c_gauss11[k], s_h1[ly + k][lx], a);
        b = fmaf(c_gauss11[k], s_h2[ly + k][lx], b);
      }
      vsum += b - a * a;
    }
  }
  atomic_add_block(msum, sums + 2 * m, red);
  atomic_add_block(vsum, sums + 2 * m + 1, red);
}

__global__ void plane_contrast_final_kernel(const float* __restrict__ sums, int M, int H, int W,
                                            float* __restrict__ mean_out, float* __restrict__ cmean_out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) {
    if (mean_out) mean_out[m] = sums[2 * m] / (float)((long)H * W);
    if (cmean_out) cmean_out[m] = sums[2 * m + 1] / (float)((long)(H - 10) * (W - 10));
  }
}

// ---------------------------------------------------------------------------------------------------------
// bicubic x0.5 (align_corners=False, A=-0.75): taps [-3/32, 19/32, 19/32, -3/32] on in[2i-1..2i+2], clamped
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bicubic_half_kernel(const float* __restrict__ in, float* __restrict__ out, int H,
                                                          int W, long total) {
  const int Ho = H / 2, Wo = W / 2;
  const float t[4] = {-0.09375f, 0.59375f, 0.59375f, -0.09375f};
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = i % Wo, y = (i / Wo) % Ho;
    const long m = i / ((long)Wo * Ho);
    const float* p = in + m * H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(2 * y - 1 + a, 0), H - 1);
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(2 * x - 1 + b, 0), W - 1);
        row = fmaf(t[b], __ldg(p + (long)yy * W + xx), row);
      }
      acc = fmaf(t[a], row, acc);
    }
    out[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// one level of the structural loss: mean over all 5x5 windows and their 25 elements of
//   ((a - mu_a)/(std_a + e) - (b - mu_b)/(std_b + e))^2,  std = sqrt(max(E[x^2]-mu^2, 0) + e),  e = 1e-5.
// Per window this is  [S_aa/d_a^2 + S_bb/d_b^2 - 2 S_ab/(d_a d_b)] with centred sums S (computed about the window
// means: no E[x^2]-mu^2 cancellation).  *acc += weight * sum / (25 * windows).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) struct_level_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          int H, int W, float scale, float* __restrict__ acc) {
  __shared__ float s_a[20][37];
  __shared__ float s_b[20][37];
  __shared__ float red[33];
  const long m = blockIdx.z;
  const float* ap = a + m * H * W;
  const float* bp = b + m * H * W;
  const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * 16;
  for (int i = threadIdx.x; i < 20 * 36; i += 256) {
    const int ly = i / 36, lx = i % 36;
    const int gy = oy0 + ly, gx = ox0 + lx;
    const bool in = gy < H && gx < W;
    s_a[ly][lx] = in ? __ldg(ap + (long)gy * W + gx) : 0.f;
    s_b[ly][lx] = in ? __ldg(bp + (long)gy * W + gx) : 0.f;
  }
  __syncthreads();
  float v = 0.f;
  for (int i = threadIdx.x; i < 16 * 32; i += 256) {
    const int ly = i / 32, lx = i % 32;
    if (oy0 + ly < H - 4 && ox0 + lx < W - 4) {
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) { sa += s_a[ly + dy][lx + dx]; sb += s_b[ly + dy][lx + dx]; }
      const float ma = sa * 0.04f, mb = sb * 0.04f;
      float saa = 0.f, sbb = 0.f, sab = 0.f;
#pragma unroll
      for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
          const float da = s_a[ly + dy][lx + dx] - ma, db = s_b[ly + dy][lx + dx] - mb;
          saa = fmaf(da, da, saa);
          sbb = fmaf(db, db, sbb);
          sab = fmaf(da, db, sab);
        }
      const float e = 1e-5f;
      const float d_a = sqrtf(fmaxf(saa * 0.04f, 0.f) + e) + e, d_b = sqrtf(fmaxf(sbb * 0.04f, 0.f) + e) + e;
      v += saa / (d_a * d_a) + sbb / (d_b * d_b) - 2.f * sab / (d_a * d_b);
    }
  }
  atomic_add_block(v * scale, acc, red);
}

// ---------------------------------------------------------------------------------------------------------
// SimpleDiscriminator: Conv4x4 s2 (1->16) + LReLU ; Conv4x4 s2 (16->32) + LReLU + Conv1x1 (32->1) ; tail dot
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) disc_conv1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        int H, int W, int Ho, int Wo, long total) {
  __shared__ float s_w[16 * 16 + 16];
  for (int i = threadIdx.x; i < 16 * 16; i += 256) s_w[i] = w[i];
  if (threadIdx.x < 16) s_w[256 + threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ox = i % Wo, oy = (i / Wo) % Ho;
    const long n = i / ((long)Wo * Ho);
    const float* p = x + n * H * W + (long)(2 * oy) * W + 2 * ox;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __ldg(p + (k >> 2) * W + (k & 3));
#pragma unroll 4
    for (int c = 0; c < 16; ++c) {
      float acc = s_w[256 + c];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc = fmaf(v[k], s_w[c * 16 + k], acc);
      out[((n * 16 + c) * Ho + oy) * Wo + ox] = acc > 0.f ? acc : 0.2f * acc;
    }
  }
}

// h [N][16][Hi][Wi] -> fea [N][Ho][Wo];  w2 [32][16][4][4], w3 [32]
__global__ void __launch_bounds__(128) disc_conv2_kernel(const float* __restrict__ h, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, const float* __restrict__ w3,
                                                        const float* __restrict__ b3, float* __restrict__ fea,
                                                        float* __restrict__ a2_out, int Hi, int Wi, int Ho, int Wo,
                                                        long total) {
  extern __shared__ float s_w2[];  // 32*256 + 32 + 32
  for (int i = threadIdx.x; i < 32 * 256; i += blockDim.x) s_w2[i] = w2[i];
  if (threadIdx.x < 32) { s_w2[8192 + threadIdx.x] = b2[threadIdx.x]; s_w2[8224 + threadIdx.x] = w3[threadIdx.x]; }
  __syncthreads();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ox = i % Wo, oy = (i / Wo) % Ho;
    const long n = i / ((long)Wo * Ho);
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = s_w2[8192 + c];
    for (int ci = 0; ci < 16; ++ci) {
      const float* p = h + ((n * 16 + ci) * Hi + 2 * oy) * Wi + 2 * ox;
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = __ldg(p + (k >> 2) * Wi + (k & 3));
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float* wr = s_w2 + c * 256 + ci * 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[c] = fmaf(v[k], wr[k], acc[c]);
      }
    }
    float o = __ldg(b3);
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float av = acc[c] > 0.f ? acc[c] : 0.2f * acc[c];
      if (a2_out) a2_out[((n * 32 + c) * Ho + oy) * Wo + ox] = av;
      o = fmaf(av, s_w2[8224 + c], o);
    }
    fea[i] = o;
  }
}

// logits[n] = sum_p fea[n][p] * w[p]
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ fea, const float* __restrict__ w, int P,
                                                    float* __restrict__ out) {
  __shared__ float red[33];
  const float* f = fea + (long)blockIdx.x * P;
  float v = 0.f;
  for (int i = threadIdx.x; i < P; i += 256) v = fmaf(f[i], __ldg(w + i), v);
  v = block_sum(v, red);
  if (threadIdx.x == 0) out[blockIdx.x] = v;
}

// ---------------------------------------------------------------------------------------------------------
// contrastive D loss:  mean_i CE([r_i, f_*], 0) + mean_i CE([-f_i, -r_*], 0)   (one block)
// ---------------------------------------------------------------------------------------------------------
__global__ void contrastive_d_kernel(const float* __restrict__ r, const float* __restrict__ f, int B,
                                     float* __restrict__ out) {
  __shared__ float red[33];
  float v = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    {  // row [r_i, f_0..]
      float mx = r[i];
      for (int j = 0; j < B; ++j) mx = fmaxf(mx, f[j]);
      float s = expf(r[i] - mx);
      for (int j = 0; j < B; ++j) s += expf(f[j] - mx);
      v += (mx + logf(s)) - r[i];
    }
    {  // row [-f_i, -r_0..]
      float mx = -f[i];
      for (int j = 0; j < B; ++j) mx = fmaxf(mx, -r[j]);
      float s = expf(-f[i] - mx);
      for (int j = 0; j < B; ++j) s += expf(-r[j] - mx);
      v += (mx + logf(s)) + f[i];
    }
  }
  v = block_sum(v, red);
  if (threadIdx.x == 0) out[0] = v / (float)B;
}

// ---------------------------------------------------------------------------------------------------------
// nce similarity: logits[b][0] = mean_hw sum_c a*p / (c0 + k|a-p|), logits[b][1] same with the negative.
// a [B][C*HW]; p, n with image stride ps / ns (0 = one sample broadcast over the batch, infoNCE2).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nce_sim_kernel(const float* __restrict__ a, const float* __restrict__ p,
                                                     long ps, const float* __restrict__ n, long ns, long CHW,
                                                     float k, float c0, float inv_hw, float* __restrict__ logits) {
  __shared__ float red[33];
  const int b = blockIdx.y;
  const float* ab = a + (long)b * CHW;
  const float* pb = p + (long)b * ps;
  const float* nb = n + (long)b * ns;
  float sp = 0.f, sn = 0.f;
  if ((CHW & 3) == 0 && ((ps | ns) & 3) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(ab);
    const float4* p4 = reinterpret_cast<const float4*>(pb);
    const float4* n4 = reinterpret_cast<const float4*>(nb);
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < CHW / 4; i += (long)gridDim.x * 256) {
      const float4 av = __ldg(a4 + i), pv = __ldg(p4 + i), nv = __ldg(n4 + i);
      const float aa[4] = {av.x, av.y, av.z, av.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w}, nn[4] = {nv.x, nv.y, nv.z, nv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sp += (aa[j] * pp[j]) * (1.f / (c0 + k * fabsf(aa[j] - pp[j])));
        sn += (aa[j] * nn[j]) * (1.f / (c0 + k * fabsf(aa[j] - nn[j])));
      }
    }
  } else {
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < CHW; i += (long)gridDim.x * 256) {
      const float av = ab[i], pv = pb[i], nv = nb[i];
      sp += (av * pv) * (1.f / (c0 + k * fabsf(av - pv)));
      sn += (av * nv) * (1.f / (c0 + k * fabsf(av - nv)));
    }
  }
  atomic_add_block(sp * inv_hw, logits + 2 * b, red);
  atomic_add_block(sn * inv_hw, logits + 2 * b + 1, red);
}

// loss = mean_b [logsumexp(l0, l1) - l0]
__global__ void ce2_kernel(const float* __restrict__ logits, int B, float* __restrict__ out) {
  __shared__ float red[33];
  float v = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float l0 = logits[2 * b], l1 = logits[2 * b + 1];
    const float mx = fmaxf(l0, l1);
    v += mx + logf(expf(l0 - mx) + expf(l1 - mx)) - l0;
  }
  v = block_sum(v, red);
  if (threadIdx.x == 0) out[0] = v / (float)B;
}

// out = mean_i |a_i - b_i|
__global__ void l1_mean_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ out) {
  __shared__ float red[33];
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += fabsf(a[i] - b[i]);
  v = block_sum(v, red);
  if (threadIdx.x == 0) out[0] = v / (float)n;
}

// TV: acc[0] += sum (x[y+1]-x[y])^2, acc[1] += sum (x[x+1]-x[x])^2 over planes
__global__ void __launch_bounds__(256) tv_kernel(const float* __restrict__ x, int H, int W, long total,
                                                float* __restrict__ acc) {
  __shared__ float red[33];
  float h = 0.f, w = 0.f;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int xx = i % W, yy = (i / W) % H;
    const float v = x[i];
    if (yy + 1 < H) { const float d = x[i + W] - v; h = fmaf(d, d, h); }
    if (xx + 1 < W) { const float d = x[i + 1] - v; w = fmaf(d, d, w); }
  }
  atomic_add_block(h, acc, red);
  atomic_add_block(w, acc + 1, red);
}
__global__ void tv_final_kernel(const float* __restrict__ acc, int B, int C, int H, int W, float* __restrict__ out) {
  if (threadIdx.x == 0) {
    const float count_h = (float)((long)(H - 1) * W), count_w = (float)((long)H * (W - 1));
    out[0] = 2.f * (acc[0] / count_h + acc[1] / count_w) / (float)B;
    (void)C;
  }
}

// ---------------------------------------------------------------------------------------------------------
// TMQI statistical naturalness (TMQI.py:210-242, original=True): u = mean(L), sig = mean over zero-padded 11x11
// blocks of the block standard deviation, N = normpdf(u)/normpdf(mode) * betapdf(sig/64.29)/betapdf(mode).
// L = 255 * x.  One thread per block; sums[2m] += sum L, sums[2m+1] += sum block std.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tmqi_blocks_kernel(const float* __restrict__ x, int H, int W, int nby, int nbx,
                                                         float* __restrict__ sums) {
  __shared__ float red[33];
  const int m = blockIdx.y;
  const float* xp = x + (long)m * H * W;
  float spix = 0.f, sstd = 0.f;
  for (int b = blockIdx.x * 128 + threadIdx.x; b < nby * nbx; b += gridDim.x * 128) {
    const int y0 = (b / nbx) * 11, x0 = (b % nbx) * 11;
    float s = 0.f;
    for (int dy = 0; dy < 11; ++dy)
      for (int dx = 0; dx < 11; ++dx) {
        const int yy = y0 + dy, xx = x0 + dx;
        if (yy < H && xx < W) s += 255.f * __ldg(xp + (long)yy * W + xx);
      }
    const float mean = s / 121.f;
    float v = 0.f;
    for (int dy = 0; dy < 11; ++dy)
      for (int dx = 0; dx < 11; ++dx) {
        const int yy = y0 + dy, xx = x0 + dx;
        const float val = (yy < H && xx < W) ? 255.f * __ldg(xp + (long)yy * W + xx) : 0.f;
        v = fmaf(val - mean, val - mean, v);
      }
    spix += s;
    sstd += sqrtf(v / 121.f);
  }
  atomic_add_block(spix, sums + 2 * m, red);
  atomic_add_block(sstd, sums + 2 * m + 1, red);
}
__global__ void tmqi_final_kernel(const float* __restrict__ sums, int M, int H, int W, int nblocks, float* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float phat1 = 4.4f, phat2 = 10.1f, muhat = 115.94f, sigmahat = 27.99f;
  const float u = sums[2 * m] / (float)((long)H * W);
  const float sig = sums[2 * m + 1] / (float)nblocks;
  const float mode = (phat1 - 1.f) / (phat1 + phat2 - 2.f);
  const float xq = sig / 64.29f;
  float pc = 0.f;
  if (xq > 0.f && xq < 1.f)
    pc = expf((phat1 - 1.f) * (logf(xq) - logf(mode)) + (phat2 - 1.f) * (logf(1.f - xq) - logf(1.f - mode)));
  const float z = (u - muhat) / sigmahat;
  out[m] = expf(-0.5f * z * z) * pc;
}

inline int cap_grid(long total, int block, int per_sm) {
  long g = (total + block - 1) / block;
  const long cap = 148L * per_sm;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

int ensure_gauss() { return 0; }

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int uncl_plane_mean_contrast(const float* x, long plane_stride, int M, int H, int W, float* mean_out,
                                        float* cmean_out, float* scratch, cudaStream_t stream) {
  UNCL_REQUIRE(M > 0 && H > 10 && W > 10, "plane_mean_contrast: planes must be larger than the 11x11 window");
  if (ensure_gauss() != 0) return uncl_set_error(UNCL_ECUDA, "plane_mean_contrast: constant upload failed");
  cudaMemsetAsync(scratch, 0, (size_t)M * 2 * sizeof(float), stream);
  plane_contrast_kernel<<<dim3(ceil_div(W, 32), ceil_div(H, 32), M), 256, 0, stream>>>(x, plane_stride, H, W, scratch);
  plane_contrast_final_kernel<<<ceil_div(M, 128), 128, 0, stream>>>(scratch, M, H, W, mean_out, cmean_out);
  return uncl_check_launch("plane_mean_contrast");
}

extern "C" int uncl_bicubic_half(const float* in, float* out, int M, int H, int W, cudaStream_t stream) {
  UNCL_REQUIRE(M > 0 && H >= 2 && W >= 2, "bicubic_half: bad shape");
  const long total = (long)M * (H / 2) * (W / 2);
  bicubic_half_kernel<<<cap_grid(total, 256, 8), 256, 0, stream>>>(in, out, H, W, total);
  return uncl_check_launch("bicubic_half");
}

// loss_out[0] = sum_l weights[l] * struct_level(fake_l, hdr_l), levels by bicubic x0.5.
// scratch: floats, >= 2 * M * (H/2) * (W/2) * (1 + 1/4) + 4
extern "C" int uncl_struct_loss_fwd(const float* fake, const float* hdr, int M, int H, int W, int levels,
                                    const float* weights_host, float* loss_out, float* scratch, cudaStream_t stream) {
  UNCL_REQUIRE(M > 0 && levels >= 1 && levels <= 4 && (H >> (levels - 1)) >= 5 && (W >> (levels - 1)) >= 5,
               "struct_loss_fwd: bad arguments");
  cudaMemsetAsync(loss_out, 0, sizeof(float), stream);
  const float* a = fake;
  const float* b = hdr;
  float* next = scratch;
  int h = H, w = W;
  for (int l = 0; l < levels; ++l) {
    const float scale = weights_host[l] / (25.f * (float)((long)M * (h - 4) * (w - 4)));
    struct_level_kernel<<<dim3(ceil_div(w - 4, 32), ceil_div(h - 4, 16), M), 256, 0, stream>>>(a, b, h, w, scale, loss_out);
    if (l + 1 < levels) {
      const long n2 = (long)M * (h / 2) * (w / 2);
      float* a2 = next;
      float* b2 = next + n2;
      next += 2 * n2;
      bicubic_half_kernel<<<cap_grid(n2, 256, 8), 256, 0, stream>>>(a, a2, h, w, n2);
      bicubic_half_kernel<<<cap_grid(n2, 256, 8), 256, 0, stream>>>(b, b2, h, w, n2);
      a = a2; b = b2; h /= 2; w /= 2;
    }
  }
  return uncl_check_launch("struct_loss_fwd");
}

// SimpleDiscriminator forward (input 256x256, dim 16, no padding).  h_scratch: N*16*127*127 floats.
// fea [N][62][62], logits [N].
extern "C" int uncl_disc_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                 const float* w3, const float* b3, const float* w_tail, float* h_scratch, float* a2_out,
                                 float* fea, float* logits, int N, int H, int W, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && H >= 10 && W >= 10, "disc_forward: bad shape");
  const int H1 = (H - 4) / 2 + 1, W1 = (W - 4) / 2 + 1, H2 = (H1 - 4) / 2 + 1, W2 = (W1 - 4) / 2 + 1;
  const long t1 = (long)N * H1 * W1, t2 = (long)N * H2 * W2;
  disc_conv1_kernel<<<cap_grid(t1, 256, 8), 256, 0, stream>>>(x, w1, b1, h_scratch, H, W, H1, W1, t1);
  const size_t smem = (32 * 256 + 64) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(disc_conv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "disc_forward: %s", cudaGetErrorString(e));
  disc_conv2_kernel<<<cap_grid(t2, 128, 4), 128, smem, stream>>>(h_scratch, w2, b2, w3, b3, fea, a2_out, H1, W1, H2, W2, t2);
  rowdot_kernel<<<N, 256, 0, stream>>>(fea, w_tail, H2 * W2, logits);
  return uncl_check_launch("disc_forward");
}

extern "C" int uncl_contrastive_d_loss(const float* real_logits, const float* fake_logits, int B, float* out,
                                       cudaStream_t stream) {
  UNCL_REQUIRE(B > 0, "contrastive_d_loss: empty batch");
  contrastive_d_kernel<<<1, 128, 0, stream>>>(real_logits, fake_logits, B, out);
  return uncl_check_launch("contrastive_d_loss");
}

// logits_scratch: 2*B floats
extern "C" int uncl_nce_fwd(const float* anchor, const float* pos, long pos_stride, const float* neg, long neg_stride,
                            int B, int C, int HW, float k, float constant, float* logits_scratch, float* loss_out,
                            cudaStream_t stream) {
  UNCL_REQUIRE(B > 0 && C > 0 && HW > 0, "nce_fwd: bad shape");
  const long CHW = (long)C * HW;
  cudaMemsetAsync(logits_scratch, 0, (size_t)2 * B * sizeof(float), stream);
  int gx = cap_grid(CHW / 4 + 1, 256, 8) / B;
  if (gx < 1) gx = 1;
  nce_sim_kernel<<<dim3(gx, B), 256, 0, stream>>>(anchor, pos, pos_stride, neg, neg_stride, CHW, k, constant, 1.f / (float)HW, logits_scratch);
  ce2_kernel<<<1, 128, 0, stream>>>(logits_scratch, B, loss_out);
  return uncl_check_launch("nce_fwd");
}

extern "C" int uncl_l1_mean(const float* a, const float* b, int n, float* out, cudaStream_t stream) {
  UNCL_REQUIRE(n > 0, "l1_mean: empty");
  l1_mean_kernel<<<1, 256, 0, stream>>>(a, b, n, out);
  return uncl_check_launch("l1_mean");
}

// scratch: 2 floats
extern "C" int uncl_tv_loss(const float* x, int B, int C, int H, int W, float* scratch, float* out,
                            cudaStream_t stream) {
  UNCL_REQUIRE(B > 0 && C > 0 && H > 1 && W > 1, "tv_loss: bad shape");
  cudaMemsetAsync(scratch, 0, 2 * sizeof(float), stream);
  const long total = (long)B * C * H * W;
  tv_kernel<<<cap_grid(total, 256, 8), 256, 0, stream>>>(x, H, W, total, scratch);
  tv_final_kernel<<<1, 32, 0, stream>>>(scratch, B, C, H, W, out);
  return uncl_check_launch("tv_loss");
}

// TMQI naturalness score of M images in [0,1] (scaled by 255 inside).  scratch: 2*M floats.
extern "C" int uncl_tmqi_naturalness(const float* x, int M, int H, int W, float* scratch, float* out,
                                     cudaStream_t stream) {
  UNCL_REQUIRE(M > 0 && H > 10 && W > 10, "tmqi_naturalness: bad shape");
  const int nby = (H + (11 - H % 11)) / 11, nbx = (W + (11 - W % 11)) / 11;
  cudaMemsetAsync(scratch, 0, (size_t)2 * M * sizeof(float), stream);
  int gx = ceil_div(nby * nbx, 128);
  if (gx > 16) gx = 16;
  tmqi_blocks_kernel<<<dim3(gx, M), 128, 0, stream>>>(x, H, W, nby, nbx, scratch);
  tmqi_final_kernel<<<ceil_div(M, 128), 128, 0, stream>>>(scratch, M, H, W, nby * nbx, out);
  return uncl_check_launch("tmqi_naturalness");
}
