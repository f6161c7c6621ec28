// Training-sample preparation on the device: the per-crop work of the reference's dataset
// loaders - cv2.resize (bilinear) to a random square size, random 256x256 crop, HWC -> CHW, RGB -> Y, and the
// LDR / log-lambda HDR normalisations - for a whole batch of crops per launch.
//
// Reference: utils/ProcessedDatasetFolderImg.py:44-168 (npy_loader), :13-22 (get_ldr_im), utils/ProcessedDatasetFolder.py:43-215
// (video variant: x-crop only), utils/hdr_image_util.py:76-82 (to_gray_tensor).
// The random draws (resize size, crop origin) stay on the host and follow the reference's np.random call order
// (uncltmo_b200/data.py); the kernels are deterministic functions of (source image, draw).
#include "common.cuh"

namespace {

struct SampleMeta {  // one crop; mirrors the int32[8] rows the host writes
  int H, W;          // source extent
  int RH, RW;        // extent after cv2.resize (== H, W when the image is not resized)
  int xx, yy;        // crop origin in the resized image
  int pad0, pad1;
};

__device__ __forceinline__ float gray_of(float r, float g, float b) { return 0.299f * r + 0.587f * g + 0.114f * b; }

// cv2.resize(..., interpolation=INTER_LINEAR) on float32: source coordinate (d + 0.5) * scale - 0.5, floor, clamp
__device__ __forceinline__ void lin_coord(int d, int src, int dst, int& s0, int& s1, float& f) {
  const double scale = (double)src / (double)dst;
  float fx = (float)(((double)d + 0.5) * scale - 0.5);
  int sx = (int)floorf(fx);
  fx -= (float)sx;
  if (sx < 0) { sx = 0; fx = 0.f; }
  if (sx >= src - 1) { sx = src - 1; fx = 0.f; }
  s0 = sx;
  s1 = min(sx + 1, src - 1);
  f = fx;
}

// grid (P*P/256, S).  src: HWC fp32 images packed in one buffer at src_off[s] (floats).  color: [S][3][P][P].
__global__ void __launch_bounds__(256) sample_crop_resize_kernel(const float* __restrict__ src_base,
                                                                const long* __restrict__ src_off,
                                                                const SampleMeta* __restrict__ meta,
                                                                float* __restrict__ color, int P) {
  const int s = blockIdx.y;
  const SampleMeta m = meta[s];
  const float* src = src_base + src_off[s];
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= P * P) return;
  const int y = i / P, x = i - y * P;
  float v[3];
  if (m.RH == m.H && m.RW == m.W) {
    const float* p = src + ((long)(y + m.yy) * m.W + (x + m.xx)) * 3;
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
  } else {
    int x0, x1, y0, y1;
    float fx, fy;
    lin_coord(x + m.xx, m.W, m.RW, x0, x1, fx);
    lin_coord(y + m.yy, m.H, m.RH, y0, y1, fy);
    const float* p00 = src + ((long)y0 * m.W + x0) * 3;
    const float* p01 = src + ((long)y0 * m.W + x1) * 3;
    const float* p10 = src + ((long)y1 * m.W + x0) * 3;
    const float* p11 = src + ((long)y1 * m.W + x1) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // horizontal pass, then vertical (the order of cv2's separable float path)
      const float top = p00[c] * (1.f - fx) + p01[c] * fx;
      const float bot = p10[c] * (1.f - fx) + p11[c] * fx;
      v[c] = top * (1.f - fy) + bot * fy;
    }
  }
  float* o = color + (long)s * 3 * P * P + i;
  o[0] = v[0];
  o[(long)P * P] = v[1];
  o[2L * P * P] = v[2];
}

// one CTA per crop: (min Y, max Y)
__global__ void __launch_bounds__(1024) sample_stats_kernel(const float* __restrict__ color, float* __restrict__ stats, int PP) {
  __shared__ float red[2][32];
  const float* c = color + (long)blockIdx.x * 3 * PP;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < PP; i += 1024) {
    const float y = gray_of(c[i], c[PP + i], c[2 * PP + i]);
    mn = fminf(mn, y);
    mx = fmaxf(mx, y);
  }
  mn = warp_min(mn); mx = warp_max(mx);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = mn; red[1][wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w) { mn = fminf(mn, red[0][w]); mx = fmaxf(mx, red[1][w]); }
    stats[2 * blockIdx.x] = mn;
    stats[2 * blockIdx.x + 1] = mx;
  }
}

// mode 0: HDR   input = log10((Y - min) / max(Y - min) * f + 1) / log10(f + 1);  gray_norm = Y / max;  gray_shift = Y - min
// mode 1: LDR   input = Y / max            (max_normalization)
// mode 2: LDR   input = Y / 255            (bugy_max_normalization)
// mode 3: LDR   input = clip(((Y - min) / max) * max_stretch - min_stretch, 0, 1)   (stretch)
__global__ void __launch_bounds__(256) sample_normalise_kernel(const float* __restrict__ color, const float* __restrict__ stats,
                                                              const float* __restrict__ f_per_sample, int mode,
                                                              float max_stretch, float min_stretch, float* __restrict__ input,
                                                              float* __restrict__ gray_norm, float* __restrict__ gray_shift,
                                                              int PP) {
  const int s = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= PP) return;
  const float* c = color + (long)s * 3 * PP;
  const float y = gray_of(c[i], c[PP + i], c[2 * PP + i]);
  const float mn = stats[2 * s], mx = stats[2 * s + 1];
  const long o = (long)s * PP + i;
  if (mode == 0) {
    const float f = f_per_sample[s];
    const float g = y - mn, gmax = mx - mn;
    const float lmax = log10f((gmax / gmax) * f + 1.f);
    input[o] = log10f((g / gmax) * f + 1.f) / lmax;
    if (gray_norm) gray_norm[o] = y / mx;
    if (gray_shift) gray_shift[o] = g;
  } else if (mode == 1) {
    input[o] = y / mx;
  } else if (mode == 2) {
    input[o] = y / 255.f;
  } else {
    input[o] = fminf(fmaxf(((y - mn) / mx) * max_stretch - min_stretch, 0.f), 1.f);
  }
}

}  // namespace

// src_base / src_off: the batch's source images (HWC fp32) packed in one device buffer, offsets in floats.
// meta: int32 [S][8] = {H, W, RH, RW, xx, yy, 0, 0}.  color: [S][3][P][P] fp32.
extern "C" int uncl_sample_crop_resize(const float* src_base, const long* src_off, const int* meta, float* color, int S,
                                       int P, cudaStream_t stream) {
  UNCL_REQUIRE(S > 0 && P > 0 && src_base && src_off && meta && color, "sample_crop_resize: bad arguments");
  sample_crop_resize_kernel<<<dim3(ceil_div(P * P, 256), S), 256, 0, stream>>>(src_base, src_off,
                                                                              reinterpret_cast<const SampleMeta*>(meta), color, P);
  return uncl_check_launch("sample_crop_resize");
}

// color [S][3][P][P] -> input [S][P][P] (+ gray_norm, gray_shift in HDR mode; may be NULL).  stats_scratch: 2*S floats.
extern "C" int uncl_sample_normalise(const float* color, int S, int P, int mode, const float* f_per_sample,
                                     float max_stretch, float min_stretch, float* input, float* gray_norm, float* gray_shift,
                                     float* stats_scratch, cudaStream_t stream) {
  UNCL_REQUIRE(S > 0 && P > 0 && mode >= 0 && mode <= 3 && color && input && stats_scratch, "sample_normalise: bad arguments");
  UNCL_REQUIRE(mode != 0 || f_per_sample != nullptr, "sample_normalise: HDR mode needs the per-sample brightness factors");
  sample_stats_kernel<<<S, 1024, 0, stream>>>(color, stats_scratch, P * P);
  sample_normalise_kernel<<<dim3(ceil_div(P * P, 256), S), 256, 0, stream>>>(color, stats_scratch, f_per_sample, mode, max_stretch,
                                                                            min_stretch, input, gray_norm, gray_shift, P * P);
  return uncl_check_launch("sample_normalise");
}
