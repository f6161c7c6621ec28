// Radiance RGBE (.hdr) decode for the entry path: the reference reads HDR frames on the host through
// imageio / FreeImage (utils/hdr_image_util.py:35-53); here the file's bytes are uploaded as they are and expanded on the GPU.
//
// The pixel stream is run-length coded per scanline and per component ("new RLE": 2 2 hi lo, then for each of the four
// components a sequence of <count><bytes> packets, count > 128 = a run).  Where a scanline starts is only known after walking
// the packets before it, so the HOST does that one cheap sequential pass over the packet headers (uncl_hdr_scan_host, no pixel is
// touched: ~1 byte in 3 is read) and the DEVICE does everything that scales with the pixel count: one CTA per scanline stages
// the coded bytes in shared memory, four threads expand the four component planes, and all threads convert RGBE to float
// (mantissa * 2^(e - 136), exactly OpenCV's / Radiance's rgbe2float) and write the three fp32 planes coalesced.
// Flat (uncoded) files take the same kernel with fixed offsets.
#include <cstdint>
#include "common.cuh"

namespace {

constexpr int kMaxW = 8192;   // scanline width bound of one CTA's shared-memory staging (4 planes + coded bytes)

__global__ void __launch_bounds__(256) hdr_decode_kernel(const unsigned char* __restrict__ data,
                                                        const long* __restrict__ offsets, int W, int H, int rle,
                                                        float* __restrict__ out) {
  extern __shared__ unsigned char sm[];
  unsigned char* planes = sm;                 // [4][W]
  unsigned char* coded = sm + 4 * (size_t)W;  // up to 4 + 4 * (W + W/127 + 2) bytes
  const int y = blockIdx.x;
  const long beg = offsets[y], end = offsets[y + 1];
  const int nbytes = (int)(end - beg);
  if (rle) {
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) coded[i] = data[beg + i];
    __syncthreads();
    if (threadIdx.x < 4) {
      // component c starts where component c-1 ended: each of the four threads walks from the scanline header
      int pos = 4;
      for (int c = 0; c <= (int)threadIdx.x; ++c) {
        unsigned char* dst = planes + (size_t)c * W;
        const bool mine = c == (int)threadIdx.x;
        int x = 0;
        while (x < W && pos < nbytes) {
          int cnt = coded[pos++];
          if (cnt > 128) {
            cnt -= 128;
            const unsigned char v = coded[pos++];
            if (mine) for (int k = 0; k < cnt && x + k < W; ++k) dst[x + k] = v;
          } else {
            if (mine) for (int k = 0; k < cnt && x + k < W; ++k) dst[x + k] = coded[pos + k];
            pos += cnt;
          }
          x += cnt;
        }
      }
    }
    __syncthreads();
  }
  const long HW = (long)H * W;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    unsigned r, g, b, e;
    if (rle) {
      r = planes[x]; g = planes[W + x]; b = planes[2 * W + x]; e = planes[3 * W + x];
    } else {
      const unsigned char* p = data + beg + 4L * x;
      r = p[0]; g = p[1]; b = p[2]; e = p[3];
    }
    const float f = e ? ldexpf(1.f, (int)e - 136) : 0.f;
    const long o = (long)y * W + x;
    out[o] = (float)r * f;
    out[HW + o] = (float)g * f;
    out[2 * HW + o] = (float)b * f;
  }
}

}  // namespace

// HOST function: offsets_host[y] = byte offset of scanline y inside `data_host` (offsets_host[H] = end of the pixel data).
// Returns 0 and *rle_out = 1 for run-length coded files, *rle_out = 0 for flat RGBE; UNCL_EUNSUPPORTED for the old RLE scheme.
extern "C" int uncl_hdr_scan_host(const unsigned char* data_host, long size, long pixel_offset, int W, int H,
                                  long* offsets_host, int* rle_out) {
  UNCL_REQUIRE(data_host && offsets_host && rle_out && W > 0 && H > 0 && pixel_offset >= 0 && pixel_offset <= size,
               "hdr_scan_host: bad arguments");
  UNCL_REQUIRE(W <= kMaxW, "hdr_scan_host: scanlines wider than %d pixels are not built", kMaxW);
  long pos = pixel_offset;
  const bool maybe_rle = W >= 8 && W < 32768 && pos + 4 <= size && data_host[pos] == 2 && data_host[pos + 1] == 2 &&
                         ((data_host[pos + 2] << 8) | data_host[pos + 3]) == W;
  if (!maybe_rle) {
    UNCL_REQUIRE(pixel_offset + 4L * W * H <= size, "hdr_scan_host: truncated flat RGBE data");
    if (W >= 8 && W < 32768) {
      // a flat file whose first pixel happens to look like an old-RLE marker (1 1 1 n) cannot be told apart cheaply: refuse
      if (data_host[pos] == 1 && data_host[pos + 1] == 1 && data_host[pos + 2] == 1)
        return uncl_set_error(UNCL_EUNSUPPORTED, "hdr_scan_host: old-style run-length coding is not built");
    }
    for (int y = 0; y <= H; ++y) offsets_host[y] = pixel_offset + 4L * W * y;
    *rle_out = 0;
    return UNCL_OK;
  }
  for (int y = 0; y < H; ++y) {
    offsets_host[y] = pos;
    if (pos + 4 > size || data_host[pos] != 2 || data_host[pos + 1] != 2 ||
        ((data_host[pos + 2] << 8) | data_host[pos + 3]) != W)
      return uncl_set_error(UNCL_EINVAL, "hdr_scan_host: bad scanline header at row %d", y);
    pos += 4;
    for (int c = 0; c < 4; ++c) {
      int x = 0;
      while (x < W) {
        if (pos >= size) return uncl_set_error(UNCL_EINVAL, "hdr_scan_host: truncated data at row %d", y);
        int cnt = data_host[pos++];
        if (cnt > 128) { cnt -= 128; pos += 1; }
        else { if (cnt == 0) return uncl_set_error(UNCL_EINVAL, "hdr_scan_host: zero-length packet at row %d", y); pos += cnt; }
        x += cnt;
      }
      if (x != W || pos > size) return uncl_set_error(UNCL_EINVAL, "hdr_scan_host: packet overruns the scanline at row %d", y);
    }
  }
  offsets_host[H] = pos;
  *rle_out = 1;
  return UNCL_OK;
}

// data / offsets: DEVICE copies of the file bytes and of uncl_hdr_scan_host's offsets; out: fp32 [3][H][W] (R, G, B planes).
extern "C" int uncl_hdr_decode(const unsigned char* data, const long* offsets, int W, int H, int rle, float* out,
                               cudaStream_t stream) {
  UNCL_REQUIRE(data && offsets && out && W > 0 && H > 0 && W <= kMaxW, "hdr_decode: bad arguments");
  const size_t smem = rle ? 4 * (size_t)W + 4 + 4 * ((size_t)W + W / 127 + 4) : 0;
  static thread_local int smem_ok = 0, smem_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if ((int)smem > 48 * 1024 && !(dev == smem_dev && (int)smem <= smem_ok)) {
    cudaError_t e = cudaFuncSetAttribute(hdr_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "hdr_decode: smem attr: %s", cudaGetErrorString(e));
    smem_ok = 96 * 1024; smem_dev = dev;
  }
  hdr_decode_kernel<<<H, 256, smem, stream>>>(data, offsets, W, H, rle, out);
  return uncl_check_launch("hdr_decode");
}
