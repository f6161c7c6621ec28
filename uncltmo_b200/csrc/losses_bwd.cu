// Backward kernels of the training losses and of the SimpleDiscriminator (dense fp32 planes [M][H][W]).
// Every kernel takes the upstream gradient as a DEVICE scalar pointer so that no host synchronisation is needed.
//
// Reference: autograd of models/struct_loss.py:46-104, GanTrainerImg.py:219-229 (contrastive_D_loss), :410-439 (nce),
// :24-56 / models/Discriminator.py:61-83 (ContrastExtracter), GanTrainer.py:669-682 (L_TV), Discriminator.py:87-126.
#include "common.cuh"

namespace {

inline int cap_grid(long total, int block, int per_sm) {
  long g = (total + block - 1) / block;
  const long cap = 148L * per_sm;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

__constant__ float c_g11[11] = {1.028380084e-03f, 7.598758135e-03f, 3.600077213e-02f, 1.093606895e-01f, 2.130055377e-01f, 2.660117249e-01f,
                                   2.130055377e-01f, 1.093606895e-01f, 3.600077213e-02f, 7.598758135e-03f, 1.028380084e-03f};   // fspecial_gauss(11, 1.5), separable factor
int ensure_gauss() { return 0; }

// ---------------------------------------------------------------------------------------------------------
// structural loss, one level.  Per 5x5 window w (u = a - mu_a, v = b - mu_b, S = centred sums, d = std + e):
//   L_w = S_aa/d_a^2 + S_bb/d_b^2 - 2 S_ab/(d_a d_b)
//   dL_w/da_k = P_w u_k - Q_w v_k,  P_w = 2/d_a^2 - 2 S_aa/(25 std_a d_a^3) + 2 S_ab/(25 std_a d_a^2 d_b),
//                                   Q_w = 2/(d_a d_b)
// coef[0..3] = scale * (P, P mu_a, Q, Q mu_b); the pixel gradient sums them over the <=25 windows that cover it.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) struct_coef_kernel(const float* __restrict__ a, const float* __restrict__ b, int H,
                                                         int W, float scale, const float* __restrict__ g_up,
                                                         float* __restrict__ coef, long plane) {
  __shared__ float s_a[20][37];
  __shared__ float s_b[20][37];
  const long m = blockIdx.z;
  const float* ap = a + m * H * W;
  const float* bp = b + m * H * W;
  const int Ho = H - 4, Wo = W - 4;
  const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * 16;
  for (int i = threadIdx.x; i < 20 * 36; i += 256) {
    const int ly = i / 36, lx = i % 36;
    const int gy = oy0 + ly, gx = ox0 + lx;
    const bool in = gy < H && gx < W;
    s_a[ly][lx] = in ? __ldg(ap + (long)gy * W + gx) : 0.f;
    s_b[ly][lx] = in ? __ldg(bp + (long)gy * W + gx) : 0.f;
  }
  __syncthreads();
  const float sc = scale * __ldg(g_up);
  for (int i = threadIdx.x; i < 16 * 32; i += 256) {
    const int ly = i / 32, lx = i % 32;
    if (oy0 + ly < Ho && ox0 + lx < Wo) {
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
          saa = fmaf(da, da, saa); sbb = fmaf(db, db, sbb); sab = fmaf(da, db, sab);
        }
      const float e = 1e-5f;
      const float std_a = sqrtf(saa * 0.04f + e), std_b = sqrtf(sbb * 0.04f + e);
      const float d_a = std_a + e, d_b = std_b + e;
      // d std_a / d a_k = u_k / (25 std_a) while the variance is positive (max(var, 0) in the reference)
      const float dstd = saa > 0.f ? 1.f / (25.f * std_a) : 0.f;
      const float P = sc * (2.f / (d_a * d_a) - 2.f * saa / (d_a * d_a * d_a) * dstd + 2.f * sab / (d_a * d_a * d_b) * dstd);
      const float Q = sc * 2.f / (d_a * d_b);
      const long o = m * Ho * Wo + (long)(oy0 + ly) * Wo + ox0 + lx;
      coef[o] = P;
      coef[plane + o] = P * ma;
      coef[2 * plane + o] = Q;
      coef[3 * plane + o] = Q * mb;
    }
  }
}

__global__ void __launch_bounds__(256) struct_grad_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         const float* __restrict__ coef, long plane, int H, int W,
                                                         long total, float* __restrict__ grad) {
  const int Ho = H - 4, Wo = W - 4;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int x = i % W, y = (i / W) % H;
    const long m = i / ((long)W * H);
    const int y0 = max(0, y - 4), y1 = min(y, Ho - 1), x0 = max(0, x - 4), x1 = min(x, Wo - 1);
    float sp = 0.f, spm = 0.f, sq = 0.f, sqm = 0.f;
    for (int wy = y0; wy <= y1; ++wy)
      for (int wx = x0; wx <= x1; ++wx) {
        const long o = m * Ho * Wo + (long)wy * Wo + wx;
        sp += __ldg(coef + o);
        spm += __ldg(coef + plane + o);
        sq += __ldg(coef + 2 * plane + o);
        sqm += __ldg(coef + 3 * plane + o);
      }
    grad[i] = a[i] * sp - spm - b[i] * sq + sqm;
  }
}

// transpose of the bicubic x0.5 down-sampler: d_in += taps * d_out (d_in already holds the level's own gradient)
__global__ void __launch_bounds__(256) bicubic_half_bwd_kernel(const float* __restrict__ d_out, float* __restrict__ d_in,
                                                              int H, int W, long total) {
  const int Ho = H / 2, Wo = W / 2;
  const float t[4] = {-0.09375f, 0.59375f, 0.59375f, -0.09375f};
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int x = i % Wo, y = (i / Wo) % Ho;
    const long m = i / ((long)Wo * Ho);
    const float g = d_out[i];
    float* p = d_in + m * H * W;
#pragma unroll
    for (int aa = 0; aa < 4; ++aa) {
      const int yy = min(max(2 * y - 1 + aa, 0), H - 1);
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int xx = min(max(2 * x - 1 + bb, 0), W - 1);
        atomicAdd(p + (long)yy * W + xx, t[aa] * t[bb] * g);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// nce backward.  loss = mean_b [lse(l0,l1) - l0];  dl0 = g (p0 - 1)/B, dl1 = g (1 - p0)/B
//   sim(a,p) = (1/HW) sum a p / (c0 + k|a-p|):  d/da = (p den - a p k sgn(a-p)) / den^2,  d/dp = (a den + a p k sgn(a-p)) / den^2
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void nce_pair_grad(float a, float p, float k, float c0, float& da, float& dp) {
  const float diff = a - p;
  const float den = c0 + k * fabsf(diff);
  const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
  const float inv2 = 1.f / (den * den);
  da = (p * den - a * p * k * sg) * inv2;
  dp = (a * den + a * p * k * sg) * inv2;
}

// d_anchor for every sample (+ d_pos / d_neg when they are per-sample tensors)
__global__ void __launch_bounds__(256) nce_bwd_kernel(const float* __restrict__ a, const float* __restrict__ p, long ps,
                                                     const float* __restrict__ n, long ns, long CHW, float k, float c0,
                                                     float inv_hw, const float* __restrict__ logits, int B,
                                                     const float* __restrict__ g_up, float* __restrict__ da,
                                                     float* __restrict__ dp, float* __restrict__ dn) {
  const int b = blockIdx.y;
  const float l0 = logits[2 * b], l1 = logits[2 * b + 1];
  const float mx = fmaxf(l0, l1);
  const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
  const float p0 = e0 / (e0 + e1);
  const float g = __ldg(g_up) / (float)B * inv_hw;
  const float dl0 = g * (p0 - 1.f), dl1 = g * (1.f - p0);
  const float* ab = a + (long)b * CHW;
  const float* pb = p + (long)b * ps;
  const float* nb = n + (long)b * ns;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < CHW; i += (long)gridDim.x * 256) {
    const float av = ab[i];
    float da_p, dp_p, da_n, dn_n;
    nce_pair_grad(av, pb[i], k, c0, da_p, dp_p);
    nce_pair_grad(av, nb[i], k, c0, da_n, dn_n);
    da[(long)b * CHW + i] = dl0 * da_p + dl1 * da_n;
    if (dp && ps) dp[(long)b * CHW + i] = dl0 * dp_p;
    if (dn && ns) dn[(long)b * CHW + i] = dl1 * dn_n;
  }
}
// broadcast positive / negative (stride 0): their gradient sums over the batch
__global__ void __launch_bounds__(256) nce_bwd_bcast_kernel(const float* __restrict__ a, const float* __restrict__ q,
                                                           long CHW, float k, float c0, float inv_hw,
                                                           const float* __restrict__ logits, int B, int which,
                                                           const float* __restrict__ g_up, float* __restrict__ dq) {
  extern __shared__ float s_dl[];  // [B]
  for (int b = threadIdx.x; b < B; b += 256) {
    const float l0 = logits[2 * b], l1 = logits[2 * b + 1];
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    const float p0 = e0 / (e0 + e1);
    const float g = __ldg(g_up) / (float)B * inv_hw;
    s_dl[b] = which == 0 ? g * (p0 - 1.f) : g * (1.f - p0);
  }
  __syncthreads();
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < CHW; i += (long)gridDim.x * 256) {
    const float qv = q[i];
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      float da_, dq_;
      nce_pair_grad(a[(long)b * CHW + i], qv, k, c0, da_, dq_);
      acc = fmaf(s_dl[b], dq_, acc);
    }
    dq[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// contrastive D loss backward (one block, O(B^2))
// ---------------------------------------------------------------------------------------------------------
__global__ void contrastive_d_bwd_kernel(const float* __restrict__ r, const float* __restrict__ f, int B,
                                         const float* __restrict__ g_up, float* __restrict__ dr, float* __restrict__ df) {
  extern __shared__ float s[];  // lse1[B], lse2[B]
  float* lse1 = s;
  float* lse2 = s + B;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    float mx = r[i];
    for (int j = 0; j < B; ++j) mx = fmaxf(mx, f[j]);
    float sum = expf(r[i] - mx);
    for (int j = 0; j < B; ++j) sum += expf(f[j] - mx);
    lse1[i] = mx + logf(sum);
    mx = -f[i];
    for (int j = 0; j < B; ++j) mx = fmaxf(mx, -r[j]);
    sum = expf(-f[i] - mx);
    for (int j = 0; j < B; ++j) sum += expf(-r[j] - mx);
    lse2[i] = mx + logf(sum);
  }
  __syncthreads();
  const float g = __ldg(g_up) / (float)B;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    // half 1 rows: [r_i, f_0..f_{B-1}] ; half 2 rows: [-f_i, -r_0..-r_{B-1}]
    float gr = expf(r[i] - lse1[i]) - 1.f;          // own row of half 1
    float gf = 1.f - expf(-f[i] - lse2[i]);         // own row of half 2 (d(-f_i)/df_i = -1)
    for (int j = 0; j < B; ++j) {
      gf += expf(f[i] - lse1[j]);                   // f_i appears in every row j of half 1
      gr -= expf(-r[i] - lse2[j]);                  // -r_i appears in every row j of half 2
    }
    dr[i] = g * gr;
    df[i] = g * gf;
  }
}

// ---------------------------------------------------------------------------------------------------------
// per-plane mean / mean-local-variance backward
//   mean = sum x / HW ; cmean = (1/No) sum_o [ (g*x^2)_o - (g*x)_o^2 ]
//   dx_p = d_mean/HW + d_cmean/No * ( 2 x_p A_p - 2 B_p ),  A_p = sum_{o covers p} g2(p-o),  B_p = sum g2(p-o) mu_o
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) plane_mu_kernel(const float* __restrict__ x, int H, int W, long total,
                                                      float* __restrict__ mu) {
  const int Ho = H - 10, Wo = W - 10;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int ox = i % Wo, oy = (i / Wo) % Ho;
    const long m = i / ((long)Wo * Ho);
    const float* p = x + m * H * W + (long)oy * W + ox;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 11; ++ky) {
      float row = 0.f;
#pragma unroll
      for (int kx = 0; kx < 11; ++kx) row = fmaf(c_g11[kx], __ldg(p + ky * W + kx), row);
      acc = fmaf(c_g11[ky], row, acc);
    }
    mu[i] = acc;
  }
}
__global__ void __launch_bounds__(256) plane_mean_contrast_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mu,
                                                                     const float* __restrict__ d_mean,
                                                                     const float* __restrict__ d_cmean, int H, int W,
                                                                     long total, float* __restrict__ dx) {
  const int Ho = H - 10, Wo = W - 10;
  const float inv_hw = 1.f / (float)((long)H * W), inv_no = 1.f / (float)((long)Ho * Wo);
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int px = i % W, py = (i / W) % H;
    const long m = i / ((long)W * H);
    float g = d_mean ? d_mean[m] * inv_hw : 0.f;
    if (d_cmean) {
      const float dc = d_cmean[m] * inv_no;
      float A = 0.f, Bv = 0.f;
      const float* mp = mu + m * Ho * Wo;
      for (int ky = 0; ky < 11; ++ky) {
        const int oy = py - ky;
        if (oy < 0 || oy >= Ho) continue;
        float ra = 0.f, rb = 0.f;
        for (int kx = 0; kx < 11; ++kx) {
          const int ox = px - kx;
          if (ox < 0 || ox >= Wo) continue;
          ra += c_g11[kx];
          rb = fmaf(c_g11[kx], __ldg(mp + (long)oy * Wo + ox), rb);
        }
        A = fmaf(c_g11[ky], ra, A);
        Bv = fmaf(c_g11[ky], rb, Bv);
      }
      g += dc * (2.f * x[i] * A - 2.f * Bv);
    }
    dx[i] = g;
  }
}

// The same gradient as ONE tiled, separable kernel (the two kernels above do 2 x 121 taps per pixel).  A tile of 32 x 32 dx
// pixels needs mu on the 42 x 42 outputs that cover it, i.e. x on 52 x 52: horizontal then vertical 11-tap pass give mu
// (zero outside the valid output range), the transposed passes give B; A_p is the product of two 1-D partial sums of g.
__global__ void __launch_bounds__(256) plane_mean_contrast_bwd_tiled_kernel(const float* __restrict__ x,
                                                                           const float* __restrict__ d_mean,
                                                                           const float* __restrict__ d_cmean, int H, int W,
                                                                           float* __restrict__ dx) {
  __shared__ float s_x[52][53];
  __shared__ float s_h[52][43];
  __shared__ float s_mu[42][43];
  __shared__ float s_t[42][33];
  const int m = blockIdx.z;
  const int Ho = H - 10, Wo = W - 10;
  const int px0 = blockIdx.x * 32, py0 = blockIdx.y * 32;
  const float* xp = x + (long)m * H * W;
  const float gm = d_mean ? d_mean[m] / (float)((long)H * W) : 0.f;
  if (d_cmean == nullptr) {
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
      const int py = py0 + i / 32, px = px0 + i % 32;
      if (py < H && px < W) dx[(long)m * H * W + (long)py * W + px] = gm;
    }
    return;
  }
  const float dc = d_cmean[m] / (float)((long)Ho * Wo);
  // x tile: rows py0 - 10 .. py0 + 41, cols px0 - 10 .. px0 + 41
  for (int i = threadIdx.x; i < 52 * 52; i += 256) {
    const int ly = i / 52, lx = i % 52;
    const int gy = py0 - 10 + ly, gx = px0 - 10 + lx;
    s_x[ly][lx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(xp + (long)gy * W + gx) : 0.f;
  }
  __syncthreads();
  // horizontal pass: h[ly][ox] for output columns ox = px0 - 10 + lx, lx < 42
  for (int i = threadIdx.x; i < 52 * 42; i += 256) {
    const int ly = i / 42, lx = i % 42;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) a = fmaf(c_g11[k], s_x[ly][lx + k], a);
    s_h[ly][lx] = a;
  }
  __syncthreads();
  // vertical pass: mu at outputs (oy, ox) = (py0 - 10 + ly, px0 - 10 + lx); zero outside [0, Ho) x [0, Wo)
  for (int i = threadIdx.x; i < 42 * 42; i += 256) {
    const int ly = i / 42, lx = i % 42;
    const int oy = py0 - 10 + ly, ox = px0 - 10 + lx;
    float a = 0.f;
    if (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) {
#pragma unroll
      for (int k = 0; k < 11; ++k) a = fmaf(c_g11[k], s_h[ly + k][lx], a);
    }
    s_mu[ly][lx] = a;
  }
  __syncthreads();
  // transposed horizontal pass: t[ly][lxp] = sum_kx g[kx] mu(oy, px - kx), px = px0 + lxp  ->  mu column lxp + 10 - kx
  for (int i = threadIdx.x; i < 42 * 32; i += 256) {
    const int ly = i / 32, lx = i % 32;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) a = fmaf(c_g11[k], s_mu[ly][lx + 10 - k], a);
    s_t[ly][lx] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 32; i += 256) {
    const int ly = i / 32, lx = i % 32;
    const int py = py0 + ly, px = px0 + lx;
    if (py >= H || px >= W) continue;
    float Bv = 0.f, ay = 0.f, ax = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      Bv = fmaf(c_g11[k], s_t[ly + 10 - k][lx], Bv);
      const int oy = py - k, ox = px - k;
      if (oy >= 0 && oy < Ho) ay += c_g11[k];
      if (ox >= 0 && ox < Wo) ax += c_g11[k];
    }
    const float xv = s_x[ly + 10][lx + 10];
    dx[(long)m * H * W + (long)py * W + px] = gm + dc * (2.f * xv * ay * ax - 2.f * Bv);
  }
}

// L1 of two vectors: out = mean |a - b| ; da = g sign(a-b)/n, db = -da
__global__ void l1_mean_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int n,
                                   const float* __restrict__ g_up, float* __restrict__ da, float* __restrict__ db) {
  const float g = __ldg(g_up) / (float)n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    const float s = d > 0.f ? g : (d < 0.f ? -g : 0.f);
    if (da) da[i] = s;
    if (db) db[i] = -s;
  }
}

// TV backward: loss = 2 (sum dh^2 / count_h + sum dw^2 / count_w) / B
__global__ void __launch_bounds__(256) tv_bwd_kernel(const float* __restrict__ x, int H, int W, long total, float ch,
                                                    float cw, const float* __restrict__ g_up, float* __restrict__ dx) {
  const float g = __ldg(g_up);
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int xx = i % W, yy = (i / W) % H;
    const float v = x[i];
    float acc = 0.f;
    if (yy > 0) acc += ch * 2.f * (v - x[i - W]);
    if (yy + 1 < H) acc -= ch * 2.f * (x[i + W] - v);
    if (xx > 0) acc += cw * 2.f * (v - x[i - 1]);
    if (xx + 1 < W) acc -= cw * 2.f * (x[i + 1] - v);
    dx[i] = g * acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// SimpleDiscriminator backward
// ---------------------------------------------------------------------------------------------------------
// fea = conv3(a2) + b3, logits = fea . w_tail.  d_fea (in/out): adds d_logits[n] * w_tail[p]; d_w_tail[p] += sum_n d_logits fea
__global__ void __launch_bounds__(256) disc_tail_bwd_kernel(const float* __restrict__ d_logits, const float* __restrict__ w_tail,
                                                           const float* __restrict__ fea, float* __restrict__ d_fea,
                                                           float* __restrict__ d_w_tail, int P, int N) {
  for (int p = blockIdx.x * 256 + threadIdx.x; p < P; p += gridDim.x * 256) {
    float acc = 0.f;
    const float wt = w_tail[p];
    for (int n = 0; n < N; ++n) {
      const float dl = d_logits ? d_logits[n] : 0.f;
      d_fea[(long)n * P + p] += dl * wt;
      acc = fmaf(dl, fea[(long)n * P + p], acc);
    }
    if (d_w_tail) d_w_tail[p] = acc;
  }
}
// d_z2[n][c][p] = d_fea[n][p] * w3[c] * lrelu'(a2) ; dw3[c] += sum d_fea a2[c] ; db3 += sum d_fea
__global__ void __launch_bounds__(256) disc_top_bwd_kernel(const float* __restrict__ d_fea, const float* __restrict__ a2,
                                                          const float* __restrict__ w3, float* __restrict__ d_z2,
                                                          float* __restrict__ dw3, float* __restrict__ db3, int P, int N) {
  __shared__ float s_acc[33];
  if (threadIdx.x < 33) s_acc[threadIdx.x] = 0.f;
  __syncthreads();
  const long total = (long)N * P;
  for (long base = (long)blockIdx.x * 256; base < total; base += (long)gridDim.x * 256) {
    const long i = base + threadIdx.x;
    const bool valid = i < total;
    const int p = valid ? (int)(i % P) : 0, n = valid ? (int)(i / P) : 0;
    const float df = valid ? d_fea[i] : 0.f;
    float sb = warp_sum(df);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[32], sb);
    for (int c = 0; c < 32; ++c) {
      const float av = valid ? a2[((long)n * 32 + c) * P + p] : 0.f;
      if (valid) d_z2[((long)n * 32 + c) * P + p] = df * __ldg(w3 + c) * (av > 0.f ? 1.f : 0.2f);
      const float sw = warp_sum(df * av);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[c], sw);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) atomicAdd(dw3 + threadIdx.x, s_acc[threadIdx.x]);
  if (threadIdx.x == 32) atomicAdd(db3, s_acc[32]);
}
// 4x4 stride-2 conv weight gradient: dW[co][ci][k] += sum in[n][ci][2oy+ky][2ox+kx] * dz[n][co][oy][ox]; db[co] += sum dz
// A CTA takes a contiguous slice of the N*Ho*Wo output pixels.  Thread = (patch element pi = ci*16 + k, pixel group pg),
// COUT accumulators in registers; the dz row of a pixel (COUT values) is broadcast from shared memory, the input value
// is read once per (pixel, patch element) - the previous version re-read the patch once per output channel.
// Partial sums are combined in shared memory, then one global atomic per (co, ci, k) per CTA (dW / db zeroed by caller).
template <int CIN, int COUT>
__global__ void __launch_bounds__(256) conv4s2_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dz,
                                                           float* __restrict__ dW, float* __restrict__ db, int Hi, int Wi,
                                                           int Ho, int Wo, long total, long per_cta) {
  constexpr int PL = CIN * 16;        // patch lanes
  constexpr int G = 256 / PL;         // pixel groups
  constexpr int CH = 32;              // pixels staged per chunk
  __shared__ __align__(16) float s_dz[CH][COUT];
  __shared__ long s_off[CH];
  __shared__ float s_acc[COUT][PL + 1];
  __shared__ float s_db[COUT];
  const int pi = threadIdx.x % PL, pg = threadIdx.x / PL;
  const int ci = pi >> 4, k = pi & 15;
  const long poff = (long)ci * Hi * Wi + (k >> 2) * Wi + (k & 3);
  const long HWo = (long)Ho * Wo;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
  float accb = 0.f;
  for (int i = threadIdx.x; i < COUT * (PL + 1); i += 256) (&s_acc[0][0])[i] = 0.f;
  if (threadIdx.x < COUT) s_db[threadIdx.x] = 0.f;
  const long p_begin = (long)blockIdx.x * per_cta;
  const long p_end = p_begin + per_cta < total ? p_begin + per_cta : total;
  for (long p0 = p_begin; p0 < p_end; p0 += CH) {
    __syncthreads();
    for (int i = threadIdx.x; i < CH * COUT; i += 256) {
      const int px = i % CH, co = i / CH;
      const long p = p0 + px;
      float v = 0.f;
      if (p < p_end) {
        const long n = p / HWo, r = p - n * HWo;
        v = dz[(n * COUT + co) * HWo + r];
        if (co == 0) {
          const int oy = (int)(r / Wo), ox = (int)(r - (long)oy * Wo);
          s_off[px] = (n * CIN * Hi + 2 * oy) * (long)Wi + 2 * ox;
        }
      } else if (co == 0) {
        s_off[px] = -1;
      }
      s_dz[px][co] = v;
    }
    __syncthreads();
    if (threadIdx.x < COUT) {
      float t = 0.f;
#pragma unroll 8
      for (int px = 0; px < CH; ++px) t += s_dz[px][threadIdx.x];
      accb += t;
    }
    for (int px = pg; px < CH; px += G) {
      const long off = s_off[px];
      if (off < 0) break;
      const float v = __ldg(in + off + poff);
      const float4* d4 = reinterpret_cast<const float4*>(&s_dz[px][0]);
#pragma unroll
      for (int c4 = 0; c4 < COUT / 4; ++c4) {
        const float4 d = d4[c4];
        acc[c4 * 4 + 0] = fmaf(v, d.x, acc[c4 * 4 + 0]);
        acc[c4 * 4 + 1] = fmaf(v, d.y, acc[c4 * 4 + 1]);
        acc[c4 * 4 + 2] = fmaf(v, d.z, acc[c4 * 4 + 2]);
        acc[c4 * 4 + 3] = fmaf(v, d.w, acc[c4 * 4 + 3]);
      }
    }
  }
  __syncthreads();
  if (G == 1) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) s_acc[c][pi] = acc[c];
  } else {
#pragma unroll
    for (int c = 0; c < COUT; ++c) atomicAdd(&s_acc[c][pi], acc[c]);
  }
  if (threadIdx.x < COUT) s_db[threadIdx.x] = accb;
  __syncthreads();
  for (int i = threadIdx.x; i < COUT * PL; i += 256) {
    const int co = i / PL, q = i - co * PL;   // q = ci*16 + k: dW is [co][ci][16]
    atomicAdd(dW + (long)co * PL + q, s_acc[co][q]);
  }
  if (db && threadIdx.x < COUT) atomicAdd(db + threadIdx.x, s_db[threadIdx.x]);
}

// First conv of the discriminator (1 -> 16 channels, 4x4 stride 2): dW[co][k] = sum_pix x[2oy+ky][2ox+kx] * dz[co][oy][ox].
// The generic kernel above stages 32 pixels per barrier pair and is bound by those round trips (108 us for 132 MFLOP at 32
// images).  Here a thread owns (output pixel, group of 4 output channels): 16 patch values x 4 dz values = 64 accumulators,
// no shared memory in the loop; a warp's lanes are 32 consecutive pixels of the same channel group, so the final reduction is
// a warp shuffle tree + one shared-memory stage + 256 atomics per CTA.  db[co] = sum dz rides along.
__global__ void __launch_bounds__(256) conv4s2_wgrad_c1_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                              float* __restrict__ dW, float* __restrict__ db, int Hi, int Wi,
                                                              int Ho, int Wo, long total) {
  __shared__ float s_red[8][64 + 4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int cg = wid & 3;                       // channel group of this warp: co = 4 cg .. 4 cg + 3
  const long HWo = (long)Ho * Wo, HWi = (long)Hi * Wi;
  float acc[4][16];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[c][k] = 0.f;
  // warps (wid >> 2) in {0, 1} of every CTA interleave over pixel chunks of 32
  const long nchunks = (total + 31) / 32;
  for (long ch = (long)blockIdx.x * 2 + (wid >> 2); ch < nchunks; ch += (long)gridDim.x * 2) {
    const long p = ch * 32 + lane;
    if (p >= total) continue;
    const long n = p / HWo, r = p - n * HWo;
    const int oy = (int)(r / Wo), ox = (int)(r - (long)oy * Wo);
    const float* xp = x + n * HWi + (long)(2 * oy) * Wi + 2 * ox;
    float v[16];
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const float2 a = __ldg(reinterpret_cast<const float2*>(xp + (long)ky * Wi));       // 2 ox is even and Wi is even: 8-byte aligned
      const float2 b = __ldg(reinterpret_cast<const float2*>(xp + (long)ky * Wi + 2));
      v[ky * 4 + 0] = a.x; v[ky * 4 + 1] = a.y; v[ky * 4 + 2] = b.x; v[ky * 4 + 3] = b.y;
    }
    const float* dp = dz + (n * 16 + 4 * cg) * HWo + r;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float g = __ldg(dp + c * HWo);
      accb[c] += g;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[c][k] = fmaf(v[k], g, acc[c][k]);
    }
  }
  // warp reduction of the 64 + 4 partial sums, then across the two warps of a channel group, then atomics
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float t = acc[c][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) s_red[wid][c * 16 + k] = t;
    }
    float t = accb[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) s_red[wid][64 + c] = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * 68; i += 256) {
    const int g = i / 68, j = i - g * 68;
    const float t = s_red[g][j] + s_red[g + 4][j];
    if (j < 64) atomicAdd(dW + (4 * g + (j >> 4)) * 16 + (j & 15), t);       // dW is [co][1][16]
    else if (db) atomicAdd(db + 4 * g + (j - 64), t);
  }
}

// 4x4 stride-2 conv data gradient: d_in[n][ci][y][x] = sum_{co,k} dz[n][co][(y-ky)/2][(x-kx)/2] * w[co][ci][k],
// optionally times lrelu'(act_in) (act_in = the post-LeakyReLU tensor that fed the conv).
// blockIdx.y = parity class (y&1, x&1): all threads of a CTA use the same (ky, kx) taps, so the weights are shared-memory
// broadcasts; a thread owns one pixel and all CIN input channels, each dz value is loaded once for CIN FMAs.
template <int CIN>
__global__ void __launch_bounds__(256) conv4s2_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ w,
                                                           const float* __restrict__ act_in, float* __restrict__ d_in,
                                                           int Cout, int Hi, int Wi, int Ho, int Wo, int N) {
  extern __shared__ __align__(16) float s_w[];  // [16 k][Cout][CIN]
  for (int i = threadIdx.x; i < Cout * CIN * 16; i += 256) {
    const int k = i & 15, ci = (i >> 4) % CIN, co = i / (16 * CIN);
    s_w[(k * Cout + co) * CIN + ci] = w[i];   // w is [co][ci][k]
  }
  __syncthreads();
  const int py = blockIdx.y >> 1, px = blockIdx.y & 1;
  const int Hh = (Hi - py + 1) / 2, Wh = (Wi - px + 1) / 2;
  const long total = (long)N * Hh * Wh;
  const long HWo = (long)Ho * Wo, HWi = (long)Hi * Wi;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int x2 = (int)(i % Wh), y2 = (int)((i / Wh) % Hh);
    const long n = i / ((long)Wh * Hh);
    const int y = 2 * y2 + py, x = 2 * x2 + px;
    float acc[CIN];
#pragma unroll
    for (int c = 0; c < CIN; ++c) acc[c] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int oy = y2 - a, ky = py + 2 * a;
      if (oy < 0 || oy >= Ho) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ox = x2 - b, kx = px + 2 * b;
        if (ox < 0 || ox >= Wo) continue;
        const float* dp = dz + n * Cout * HWo + (long)oy * Wo + ox;
        const float* wk = s_w + (size_t)(ky * 4 + kx) * Cout * CIN;
        for (int co = 0; co < Cout; ++co) {
          const float g = __ldg(dp + co * HWo);
#pragma unroll
          for (int c = 0; c < CIN; ++c) acc[c] = fmaf(g, wk[co * CIN + c], acc[c]);
        }
      }
    }
    const long o = n * CIN * HWi + (long)y * Wi + x;
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      float v = acc[c];
      if (act_in) v *= act_in[o + c * HWi] > 0.f ? 1.f : 0.2f;
      d_in[o + c * HWi] = v;
    }
  }
}

// Data gradient of the FIRST discriminator conv (16 -> 1 channel, 4x4 stride 2; the gradient that flows back into the
// generator).  A thread owns a 2x2 block of input pixels (all four parity classes): they share the same four dz positions
// (oy in {y2, y2-1}, ox in {x2, x2-1}), so each dz value is loaded once for four FMAs instead of once per FMA as in the
// per-pixel kernel above (49 -> ~12 us at 16 images).  Weights: 16 channels x 16 taps, shared-memory broadcasts.
__global__ void __launch_bounds__(256) conv4s2_dgrad_c1_kernel(const float* __restrict__ dz, const float* __restrict__ w,
                                                              float* __restrict__ d_in, int Hi, int Wi, int Ho, int Wo,
                                                              int N) {
  __shared__ __align__(16) float s_w[16 * 16];   // [co][ky*4+kx]
  if (threadIdx.x < 256) s_w[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const int Hb = (Hi + 1) / 2, Wb = (Wi + 1) / 2;
  const long total = (long)N * Hb * Wb;
  const long HWo = (long)Ho * Wo, HWi = (long)Hi * Wi;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int x2 = (int)(i % Wb), y2 = (int)((i / Wb) % Hb);
    const long n = i / ((long)Wb * Hb);
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    const float* dn = dz + n * 16 * HWo;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int oy = y2 - a;
      if (oy < 0 || oy >= Ho) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ox = x2 - b;
        if (ox < 0 || ox >= Wo) continue;
        const float* dp = dn + (long)oy * Wo + ox;
#pragma unroll
        for (int co = 0; co < 16; ++co) {
          const float g = __ldg(dp + co * HWo);
          // pixel (py, px) of the block takes tap (ky, kx) = (py + 2a, px + 2b)
          const float* wk = s_w + co * 16 + (2 * a) * 4 + 2 * b;
          acc[0][0] = fmaf(g, wk[0], acc[0][0]);
          acc[0][1] = fmaf(g, wk[1], acc[0][1]);
          acc[1][0] = fmaf(g, wk[4], acc[1][0]);
          acc[1][1] = fmaf(g, wk[5], acc[1][1]);
        }
      }
    }
    const int y = 2 * y2, x = 2 * x2;
    float* o = d_in + n * HWi + (long)y * Wi + x;
    if (x + 1 < Wi && (Wi & 1) == 0) {
      *reinterpret_cast<float2*>(o) = make_float2(acc[0][0], acc[0][1]);
      if (y + 1 < Hi) *reinterpret_cast<float2*>(o + Wi) = make_float2(acc[1][0], acc[1][1]);
    } else {
      o[0] = acc[0][0];
      if (x + 1 < Wi) o[1] = acc[0][1];
      if (y + 1 < Hi) { o[Wi] = acc[1][0]; if (x + 1 < Wi) o[Wi + 1] = acc[1][1]; }
    }
  }
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
// d_fake [M][H][W] (written).  scratch floats: >= 6.5 * M*H*W + 64
extern "C" int uncl_struct_loss_bwd(const float* fake, const float* hdr, int M, int H, int W, int levels,
                                    const float* weights_host, const float* g_up, float* d_fake, float* scratch,
                                    cudaStream_t stream) {
  UNCL_REQUIRE(M > 0 && levels >= 1 && levels <= 4 && (H >> (levels - 1)) >= 5 && (W >> (levels - 1)) >= 5,
               "struct_loss_bwd: bad arguments");
  const float* a[4];
  const float* b[4];
  float* g[4];
  int hs[4], ws[4];
  a[0] = fake; b[0] = hdr; g[0] = d_fake; hs[0] = H; ws[0] = W;
  float* next = scratch;
  float* coef = next; next += 4L * M * H * W;
  for (int l = 1; l < levels; ++l) {
    hs[l] = hs[l - 1] / 2; ws[l] = ws[l - 1] / 2;
    const long n = (long)M * hs[l] * ws[l];
    float* al = next; float* bl = next + n; g[l] = next + 2 * n; next += 3 * n;
    a[l] = al; b[l] = bl;
    int rc = uncl_bicubic_half(a[l - 1], al, M, hs[l - 1], ws[l - 1], stream);
    if (rc) return rc;
    rc = uncl_bicubic_half(b[l - 1], bl, M, hs[l - 1], ws[l - 1], stream);
    if (rc) return rc;
  }
  for (int l = 0; l < levels; ++l) {
    const int h = hs[l], w = ws[l];
    const long plane = (long)M * (h - 4) * (w - 4);
    const float scale = weights_host[l] / (25.f * (float)plane);
    struct_coef_kernel<<<dim3(ceil_div(w - 4, 32), ceil_div(h - 4, 16), M), 256, 0, stream>>>(a[l], b[l], h, w, scale, g_up, coef, plane);
    const long total = (long)M * h * w;
    struct_grad_kernel<<<cap_grid(total, 256, 8), 256, 0, stream>>>(a[l], b[l], coef, plane, h, w, total, g[l]);
  }
  for (int l = levels - 1; l >= 1; --l) {
    const long total = (long)M * hs[l] * ws[l];
    bicubic_half_bwd_kernel<<<cap_grid(total, 256, 8), 256, 0, stream>>>(g[l], g[l - 1], hs[l - 1], ws[l - 1], total);
  }
  return uncl_check_launch("struct_loss_bwd");
}

extern "C" int uncl_nce_bwd(const float* anchor, const float* pos, long pos_stride, const float* neg, long neg_stride,
                            int B, int C, int HW, float k, float constant, const float* logits, const float* g_up,
                            float* d_anchor, float* d_pos, float* d_neg, cudaStream_t stream) {
  UNCL_REQUIRE(B > 0 && C > 0 && HW > 0 && d_anchor, "nce_bwd: bad arguments");
  const long CHW = (long)C * HW;
  int gx = cap_grid(CHW, 256, 8) / B;
  if (gx < 1) gx = 1;
  nce_bwd_kernel<<<dim3(gx, B), 256, 0, stream>>>(anchor, pos, pos_stride, neg, neg_stride, CHW, k, constant, 1.f / (float)HW, logits, B, g_up, d_anchor, d_pos, d_neg);
  if (d_pos && pos_stride == 0)
    nce_bwd_bcast_kernel<<<cap_grid(CHW, 256, 8), 256, B * sizeof(float), stream>>>(anchor, pos, CHW, k, constant, 1.f / (float)HW, logits, B, 0, g_up, d_pos);
  if (d_neg && neg_stride == 0)
    nce_bwd_bcast_kernel<<<cap_grid(CHW, 256, 8), 256, B * sizeof(float), stream>>>(anchor, neg, CHW, k, constant, 1.f / (float)HW, logits, B, 1, g_up, d_neg);
  return uncl_check_launch("nce_bwd");
}

extern "C" int uncl_contrastive_d_bwd(const float* real_logits, const float* fake_logits, int B, const float* g_up,
                                      float* d_real, float* d_fake, cudaStream_t stream) {
  UNCL_REQUIRE(B > 0 && B <= 4096, "contrastive_d_bwd: bad batch");
  contrastive_d_bwd_kernel<<<1, 128, 2 * B * sizeof(float), stream>>>(real_logits, fake_logits, B, g_up, d_real, d_fake);
  return uncl_check_launch("contrastive_d_bwd");
}

// mu_scratch: M*(H-10)*(W-10) floats.  d_mean / d_cmean may be NULL.
extern "C" int uncl_plane_mean_contrast_bwd(const float* x, int M, int H, int W, const float* d_mean, const float* d_cmean,
                                            float* dx, float* mu_scratch, cudaStream_t stream) {
  UNCL_REQUIRE(M > 0 && H > 10 && W > 10, "plane_mean_contrast_bwd: bad shape");
  if (ensure_gauss() != 0) return uncl_set_error(UNCL_ECUDA, "plane_mean_contrast_bwd: constant upload failed");
  (void)mu_scratch;   // the tiled kernel keeps mu in shared memory (the argument stays for ABI stability)
  if ((long)M <= 65535) {
    plane_mean_contrast_bwd_tiled_kernel<<<dim3(ceil_div(W, 32), ceil_div(H, 32), M), 256, 0, stream>>>(x, d_mean, d_cmean, H, W, dx);
    return uncl_check_launch("plane_mean_contrast_bwd");
  }
  if (d_cmean) {
    const long to = (long)M * (H - 10) * (W - 10);
    plane_mu_kernel<<<cap_grid(to, 256, 8), 256, 0, stream>>>(x, H, W, to, mu_scratch);
  }
  const long total = (long)M * H * W;
  plane_mean_contrast_bwd_kernel<<<cap_grid(total, 256, 8), 256, 0, stream>>>(x, mu_scratch, d_mean, d_cmean, H, W, total, dx);
  return uncl_check_launch("plane_mean_contrast_bwd");
}

extern "C" int uncl_l1_mean_bwd(const float* a, const float* b, int n, const float* g_up, float* da, float* db,
                                cudaStream_t stream) {
  UNCL_REQUIRE(n > 0, "l1_mean_bwd: empty");
  l1_mean_bwd_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(a, b, n, g_up, da, db);
  return uncl_check_launch("l1_mean_bwd");
}

extern "C" int uncl_tv_bwd(const float* x, int B, int C, int H, int W, const float* g_up, float* dx, cudaStream_t stream) {
  UNCL_REQUIRE(B > 0 && C > 0 && H > 1 && W > 1, "tv_bwd: bad shape");
  const long total = (long)B * C * H * W;
  const float ch = 2.f / ((float)((long)(H - 1) * W) * (float)B), cw = 2.f / ((float)((long)H * (W - 1)) * (float)B);
  tv_bwd_kernel<<<cap_grid(total, 256, 8), 256, 0, stream>>>(x, H, W, total, ch, cw, g_up, dx);
  return uncl_check_launch("tv_bwd");
}

// SimpleDiscriminator backward.  Inputs of the forward: x [N][256][256], h1 [N][16][127][127] (post-LReLU),
// a2 [N][32][62][62] (post-LReLU), fea [N][62][62].  d_fea (in/out): gradient from the feature branch, the tail
// contribution is added here.  Outputs: dx (NULL to skip) and parameter gradients (dw*, db* zeroed by the caller
// where accumulated: dw3, db3).  scratch: N*32*62*62 + N*16*127*127 floats.
extern "C" int uncl_disc_backward(const float* x, const float* h1, const float* a2, const float* fea, const float* w1,
                                  const float* w2, const float* w3, const float* w_tail, const float* d_logits,
                                  float* d_fea, float* dx, float* dw1, float* db1, float* dw2, float* db2, float* dw3,
                                  float* db3, float* dw_tail, float* scratch, int N, cudaStream_t stream) {
  UNCL_REQUIRE(N > 0, "disc_backward: empty batch");
  const int H = 256, H1 = 127, H2 = 62, P = H2 * H2;
  float* d_z2 = scratch;
  float* d_z1 = scratch + (long)N * 32 * P;
  disc_tail_bwd_kernel<<<ceil_div(P, 256), 256, 0, stream>>>(d_logits, w_tail, fea, d_fea, dw_tail, P, N);
  disc_top_bwd_kernel<<<cap_grid((long)N * P, 256, 2), 256, 0, stream>>>(d_fea, a2, w3, d_z2, dw3, db3, P, N);
  const int ctas = 148 * 2;
  const long t2 = (long)N * H2 * H2;
  // dw2 / dw1 == NULL: the caller does not want the convolutions' weight gradients (the generator step back-propagates
  // THROUGH the discriminator only; its parameter gradients would be thrown away by the next zero_grad).
  // The second conv's weight gradient is a leaf: it runs on a forked stream next to the data-gradient -> first-conv
  // weight-gradient chain (both need only d_z2) and is joined before this call returns; inside a CUDA-graph capture the
  // fork becomes a parallel branch.  (Each of the three kernels is ~100 us at 32 images and none fills the GPU.)
  static thread_local cudaStream_t fork_stream = nullptr;
  static thread_local cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  static thread_local int fork_dev = -1;
  bool forked = false;
  if (dw2 != nullptr) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (fork_dev != dev) {
      fork_stream = nullptr;
      if (cudaStreamCreateWithFlags(&fork_stream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess)
        fork_stream = nullptr;
      fork_dev = dev;
    }
    forked = fork_stream != nullptr && cudaEventRecord(ev_fork, stream) == cudaSuccess &&
             cudaStreamWaitEvent(fork_stream, ev_fork, 0) == cudaSuccess;
    conv4s2_wgrad_kernel<16, 32><<<ctas, 256, 0, forked ? fork_stream : stream>>>(h1, d_z2, dw2, db2, H1, H1, H2, H2, t2,
                                                                                (t2 + ctas - 1) / ctas);
    if (forked) cudaEventRecord(ev_join, fork_stream);
  }
  cudaError_t e = cudaFuncSetAttribute(conv4s2_dgrad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 16 * 16 * 4);
  if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "disc_backward: %s", cudaGetErrorString(e));
  const long q1 = (long)N * 64 * 64;   // pixels per parity class (upper bound)
  conv4s2_dgrad_kernel<16><<<dim3(cap_grid(q1, 256, 2), 4), 256, 32 * 16 * 16 * 4, stream>>>(d_z2, w2, h1, d_z1, 32, H1, H1, H2, H2, N);
  const long t1 = (long)N * H1 * H1;
  if (dw1 != nullptr)
    conv4s2_wgrad_c1_kernel<<<148 * 4, 256, 0, stream>>>(x, d_z1, dw1, db1, H, H, H1, H1, t1);
  if (dx) {
    const long q0 = (long)N * 128 * 128;
    conv4s2_dgrad_c1_kernel<<<cap_grid(q0, 256, 4), 256, 0, stream>>>(d_z1, w1, dx, H, H, H1, H1, N);
  }
  if (forked) cudaStreamWaitEvent(stream, ev_join, 0);
  return uncl_check_launch("disc_backward");
}
