// Off-default-path operators of SURVEY.md §8(a): the lambda-search objective (a19) and the optional output
// activations / normalisers of models/Blocks.py (a20).
//
// Reference: utils/adaptive_lambda.py:7-21 (cross_entropy: log10(g*lambda+1)/max -> 20-bin density histogram ->
// cross entropy against the mean LDR histogram), models/Blocks.py:77-138 (Exp, MySig, Clip, Max / MinMax normalisers).
#include "common.cuh"

namespace {

inline int cap_grid(long total, int block, int per_sm) {
  long g = (total + block - 1) / block;
  const long cap = 148L * per_sm;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

// hist[l][bin] += count of pixels whose log10(g*lambda_l+1)/log10(gmax*lambda_l+1) falls in bin (range [0,1], last bin closed)
// grid: (pixel chunks, L)
__global__ void __launch_bounds__(256) lambda_hist_kernel(const float* __restrict__ gray, long n, const float* __restrict__ gmax_ptr,
                                                         const double* __restrict__ lambdas, int bins,
                                                         unsigned* __restrict__ hist) {
  extern __shared__ unsigned s_h[];
  for (int i = threadIdx.x; i < bins; i += 256) s_h[i] = 0;
  __syncthreads();
  const float lam = (float)lambdas[blockIdx.y];
  const float inv = 1.f / log10f(__ldg(gmax_ptr) * lam + 1.f);
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const float v = log10f(__ldg(gray + i) * lam + 1.f) * inv;
    if (v >= 0.f && v <= 1.f) {
      int b = (int)(v * bins);
      if (b >= bins) b = bins - 1;
      atomicAdd(&s_h[b], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += 256)
    if (s_h[i]) atomicAdd(hist + (long)blockIdx.y * bins + i, s_h[i]);
}
// ce[l] = -sum_b targets[b] * log(density_b + 1e-9) / bins, density = count * bins / n
__global__ void lambda_ce_kernel(const unsigned* __restrict__ hist, const float* __restrict__ targets, int bins, long n,
                                 int L, float* __restrict__ ce) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  double acc = 0.0;
  for (int b = 0; b < bins; ++b) {
    const double dens = (double)hist[(long)l * bins + b] * bins / (double)n;
    acc += (double)targets[b] * log(dens + 1e-9);
  }
  ce[l] = (float)(-acc / bins);
}

__global__ void plane_minmax_kernel(const float* __restrict__ x, long per, float* __restrict__ mn, float* __restrict__ mx) {
  __shared__ float r0[32], r1[32];
  const float* p = x + (long)blockIdx.x * per;
  float a = INFINITY, b = -INFINITY;
  for (long i = threadIdx.x; i < per; i += blockDim.x) { const float v = p[i]; a = fminf(a, v); b = fmaxf(b, v); }
  a = warp_min(a); b = warp_max(b);
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = a; r1[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a = fminf(a, r0[w]); b = fmaxf(b, r1[w]); }
    mn[blockIdx.x] = a; mx[blockIdx.x] = b;
  }
}

// mode: 0 Exp (e^x - 1), 1 MySig(factor), 2 Clip, 3 MaxNormalization (per sample), 4 MaxNormalizationEpsilon,
//       5 BatchMaxNormalization, 6 MinMaxNormalization (per sample)
__global__ void blocks_apply_kernel(const float* __restrict__ x, float* __restrict__ out, long per, long n, int mode,
                                    float param, const float* __restrict__ mn, const float* __restrict__ mx, int N) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const int s = (int)(i / per);
    float r;
    switch (mode) {
      case 0: r = expf(v) - 1.f; break;
      case 1: r = 1.f / (1.f + expf(-param * v)); break;
      case 2: r = fminf(fmaxf(v * 1.1f - 0.05f, 0.f), 1.f); break;
      case 3: r = v / mx[s]; break;
      case 4: r = v / mx[s] - 1e-8f; break;
      case 5: { float m = mx[0]; for (int k = 1; k < N; ++k) m = fmaxf(m, mx[k]); r = v / m; break; }
      default: r = (v - mn[s]) / (mx[s] - mn[s] + 1e-8f); break;
    }
    out[i] = r;
  }
}

}  // namespace

// ce_out[l] = adaptive_lambda.cross_entropy(lambdas[l], gray, targets, bins) for a whole population at once.
// gray: n fp32 pixels (the caller's gray / gray.max(), or any non-negative image); lambdas: L doubles (device);
// targets: `bins` floats; workspace: (L*bins + 4) * 4 bytes.
extern "C" int uncl_lambda_cross_entropy(const float* gray, long n, const double* lambdas, int L, const float* targets,
                                         int bins, float* ce_out, void* workspace, cudaStream_t stream) {
  UNCL_REQUIRE(n > 0 && L > 0 && bins > 0 && bins <= 1024, "lambda_cross_entropy: bad arguments");
  unsigned* hist = reinterpret_cast<unsigned*>(workspace);
  float* mm = reinterpret_cast<float*>(hist + (long)L * bins);
  cudaMemsetAsync(hist, 0, (size_t)L * bins * sizeof(unsigned), stream);
  plane_minmax_kernel<<<1, 1024, 0, stream>>>(gray, n, mm, mm + 1);
  int gx = cap_grid(n, 256, 8) / L;
  if (gx < 1) gx = 1;
  lambda_hist_kernel<<<dim3(gx, L), 256, bins * sizeof(unsigned), stream>>>(gray, n, mm + 1, lambdas, bins, hist);
  lambda_ce_kernel<<<ceil_div(L, 128), 128, 0, stream>>>(hist, targets, bins, n, L, ce_out);
  return uncl_check_launch("lambda_cross_entropy");
}

// models/Blocks.py activations / normalisers on a dense [N][per] fp32 tensor.  scratch: 2*N floats.
extern "C" int uncl_blocks_apply(const float* x, float* out, int N, long per, int mode, float param, float* scratch,
                                 cudaStream_t stream) {
  UNCL_REQUIRE(N > 0 && per > 0 && mode >= 0 && mode <= 6, "blocks_apply: bad arguments");
  if (mode >= 3) plane_minmax_kernel<<<N, 512, 0, stream>>>(x, per, scratch, scratch + N);
  blocks_apply_kernel<<<cap_grid((long)N * per, 256, 8), 256, 0, stream>>>(x, out, per, (long)N * per, mode, param, scratch, scratch + N, N);
  return uncl_check_launch("blocks_apply");
}
