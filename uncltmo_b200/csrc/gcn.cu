// Bottleneck graph block of the generator (ViG Grapher + FFN on the 12x12 = 144-node grid).
//
// Reference: Unet_singleFrame.py:44-99 (GCNBlock), :20-42 (FFN); gcn_lib/torch_vertex.py:181-227 (Grapher_noBN),
// :109-130 (DyGraphConv2d), :13-30 (MRConv2d); gcn_lib/torch_edge.py:135-159, 54-86, 9-20 (normalise, pairwise
// distance + relative_pos, top-9); gcn_lib/torch_nn.py:81-102 (batched_index_select), :54-78 (BasicConv, groups=4).
//
// All internals are fp32 regardless of the conv precision: the KNN is a discrete choice (SURVEY.md "hard parts").
// Tensors here are C8-blocked fp32 with HW = 144: [N][C/8][144][8].
#include "common.cuh"

constexpr int GN = 144;   // nodes
constexpr int GK = 9;     // neighbours

// x0 = in + pos_embed (pos_embed pre-blocked [C/8][144][8] fp32)
// `split` (optional, bf16 blocked [N][3C/8][144][8]): x0 as [hi | hi | lo] with hi = bf16(x0), lo = bf16(x0 - hi), the A
// operand of the three-term bf16 GEMM x_hi.w_hi + x_hi.w_lo + x_lo.w_hi that stands in for the fp32 fc1 on the tensor
// cores (relative error ~2^-16 per product instead of bf16's 2^-9; packing.pointwise_tc_split builds the matching weights)
template <typename T>
__global__ void gcn_add_pos_kernel(const T* __restrict__ in, long in_img_stride, const float* __restrict__ pos,
                                   float* __restrict__ out, bf16* __restrict__ split, int C, int N) {
  const long per = (long)C * GN;
  const long total = (long)N * per;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / per, o = i % per;
    const float v = to_f(in[n * in_img_stride + o]) + pos[o];
    out[i] = v;
    if (split != nullptr) {
      const bf16 hi = __float2bfloat16_rn(v);
      const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      bf16* s = split + n * 3 * per + o;
      s[0] = hi;
      s[per] = hi;
      s[2 * per] = lo;
    }
  }
}

// Pointwise (1x1) conv on blocked fp32 tensors with groups, bias, activation, optional residual and per-sample
// scale of the branch (DropPath: out = scale[n] * f(x) + res).  Weights packed [groups][Cin_g][Cout_g].
// CTA: 128 pixels x PW_CO output channels, K step 16; thread: 8 pixels x (PW_CO/16) channels.
constexpr int PW_PX = 128, PW_K = 16;

template <typename TOut, int PW_CO>
__global__ void __launch_bounds__(256) pw_conv_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                     const float* __restrict__ bias, const float* __restrict__ res,
                                                     const float* __restrict__ scale, TOut* __restrict__ out,
                                                     long out_img_stride, int C_in, int C_out, int groups, int HW,
                                                     int N, int act) {
  constexpr int TC = PW_CO / 16;  // channels per thread (8 or 4)
  __shared__ __align__(16) float s_a[PW_K][PW_PX];  // [ci][pixel]
  __shared__ __align__(16) float s_w[PW_K][PW_CO];  // [ci][co]
  const int cin_g = C_in / groups, cout_g = C_out / groups;
  const int co0 = blockIdx.y * PW_CO;
  const int g = co0 / cout_g;
  const int P = N * HW;
  const int p0 = blockIdx.x * PW_PX;
  const int tp = (threadIdx.x & 15) * 8, tc = (threadIdx.x >> 4) * TC;
  float acc[8][TC];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < cin_g; k0 += PW_K) {
    __syncthreads();
    // A tile: 16 ci x 128 px = two 8-channel blocks; each thread moves one pixel of one block (32 B)
    {
      const int px = threadIdx.x & 127, blk = threadIdx.x >> 7;
      const int p = p0 + px, c = g * cin_g + k0 + blk * 8;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p < P) {
        const int n = p / HW, q = p - n * HW;
        load8(in + ((long)n * (C_in / 8) + (c >> 3)) * HW * 8 + (long)q * 8, v);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s_a[blk * 8 + j][px] = v[j];
    }
    for (int i = threadIdx.x; i < PW_K * PW_CO; i += 256) {
      const int co = i % PW_CO, ci = i / PW_CO;
      s_w[ci][co] = w[((long)g * cin_g + k0 + ci) * cout_g + (co0 - g * cout_g) + co];
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < PW_K; ++ci) {
      float av[8], bv[TC];
      *reinterpret_cast<float4*>(av) = *reinterpret_cast<const float4*>(&s_a[ci][tp]);
      *reinterpret_cast<float4*>(av + 4) = *reinterpret_cast<const float4*>(&s_a[ci][tp + 4]);
#pragma unroll
      for (int j = 0; j < TC; ++j) bv[j] = s_w[ci][tc + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = p0 + tp + i;
    if (p >= P) continue;
    const int n = p / HW, q = p - n * HW;
    const float sc = scale ? scale[n] : 1.f;
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      const int c = co0 + tc + j;
      const long o = ((long)(c >> 3) * HW + q) * 8 + (c & 7);
      float v = apply_act(acc[i][j] + (bias ? bias[c] : 0.f), act) * sc;
      if (res) v += res[(long)n * C_out * HW + o];
      from_f(out[(long)n * out_img_stride + o], v);
    }
  }
}

// KNN graph + max-relative aggregation (C = 256).  KNN_SPLIT CTAs per image, each owning a contiguous range of nodes.
//   yn = y / max(|y|_2, 1e-12) over channels; dist[i][j] = (|yn_i|^2 - 2 yn_i.yn_j) + |yn_j|^2 + relpos[i][j];
//   idx[i] = 9 smallest; z = interleave(y, max_k(y[idx_k] - y_i)) -> [N][2C/8][144][8]
// Distances: a warp takes six rows i at a time; each lane keeps 5 columns j (lane + 32 t) x 6 rows in registers and
// streams the channel axis with 128-bit shared loads (row stride C+4 floats keeps them conflict-free).
constexpr int KNN_SPLIT = 2;

template <int C, typename TZ>
__global__ void __launch_bounds__(512) gcn_knn_agg_kernel(const float* __restrict__ y, const float* __restrict__ relpos,
                                                         TZ* __restrict__ z, int* __restrict__ idx_out) {
  extern __shared__ __align__(16) float smem[];
  constexpr int LD = C + 4;
  constexpr int ROWS = GN / KNN_SPLIT;
  float* s_yn = smem;                 // [144][C+4]
  float* s_sq = smem + GN * LD;       // [144]
  float* s_nrm = s_sq + GN;           // [144] max(|y|, 1e-12): de-normalises the shared-memory copy for the aggregation
  int* s_idx = reinterpret_cast<int*>(s_nrm + GN);  // [ROWS][9]
  const int n = blockIdx.x / KNN_SPLIT, part = blockIdx.x % KNN_SPLIT;
  const int i_begin = part * ROWS;
  const float* yn_g = y + (long)n * C * GN;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;

  {
    const float4* y4 = reinterpret_cast<const float4*>(yn_g);   // 128-bit loads: half a pixel block each
    for (int i = threadIdx.x; i < C * GN / 4; i += blockDim.x) {
      const int half = i & 1, node = (i >> 1) % GN, cb = i / (2 * GN);
      *reinterpret_cast<float4*>(&s_yn[node * LD + cb * 8 + half * 4]) = __ldg(y4 + i);
    }
  }
  __syncthreads();
  for (int node = wid; node < GN; node += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = s_yn[node * LD + c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    const float nrm = fmaxf(sqrtf(s), 1e-12f);
    const float inv = 1.f / nrm;
    if (lane == 0) s_nrm[node] = nrm;
    float s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = s_yn[node * LD + c] * inv;
      s_yn[node * LD + c] = v;
      s2 = fmaf(v, v, s2);
    }
    s2 = warp_sum(s2);
    if (lane == 0) s_sq[node] = s2;
  }
  __syncthreads();
  // KR rows per warp: the B rows (5 x 128-bit loads per lane and K step) are what saturates shared memory, so every
  // extra row held in registers divides that traffic; 72 rows / 6 = 12 warps take one group each
  constexpr int KR = 6;
  static_assert(ROWS % KR == 0, "row groups");
  for (int ip = wid; ip < ROWS / KR; ip += nw) {
    const int i0 = i_begin + KR * ip;
    float dot[KR][5];
#pragma unroll
    for (int r = 0; r < KR; ++r)
#pragma unroll
      for (int t = 0; t < 5; ++t) dot[r][t] = 0.f;
    const float* yi0 = s_yn + i0 * LD;
    const float* yj[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) yj[t] = s_yn + min(lane + 32 * t, GN - 1) * LD;
#pragma unroll 2
    for (int c = 0; c < C; c += 4) {
      float4 a[KR];
#pragma unroll
      for (int r = 0; r < KR; ++r) a[r] = *reinterpret_cast<const float4*>(yi0 + r * LD + c);
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const float4 b = *reinterpret_cast<const float4*>(yj[t] + c);
#pragma unroll
        for (int r = 0; r < KR; ++r) {
          dot[r][t] = fmaf(a[r].x, b.x, dot[r][t]); dot[r][t] = fmaf(a[r].y, b.y, dot[r][t]);
          dot[r][t] = fmaf(a[r].z, b.z, dot[r][t]); dot[r][t] = fmaf(a[r].w, b.w, dot[r][t]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < KR; ++r) {
      const int i = i0 + r;
      float d[5];
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const int j = lane + 32 * t;
        d[t] = j < GN ? ((s_sq[i] + (-2.f * dot[r][t])) + s_sq[j]) + relpos[i * GN + j] : INFINITY;
      }
      for (int k = 0; k < GK; ++k) {
        float best = d[0];
        int bj = lane;
#pragma unroll
        for (int t = 1; t < 5; ++t)
          if (d[t] < best) { best = d[t]; bj = lane + 32 * t; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
          if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
#pragma unroll
        for (int t = 0; t < 5; ++t)
          if (bj == lane + 32 * t) d[t] = INFINITY;
        if (lane == 0) s_idx[(i - i_begin) * GK + k] = bj;
      }
    }
  }
  __syncthreads();
  if (idx_out)
    for (int i = threadIdx.x; i < ROWS * GK; i += blockDim.x) idx_out[((long)n * GN + i_begin) * GK + i] = s_idx[i];
  // aggregation on the raw (un-normalised) features: the node's own row from global memory (exact), its nine neighbours
  // from the shared-memory copy times their norm (within 2 ulp of the raw value; the 9 dependent 32-byte gathers per lane
  // from L2 were 58 % of this kernel's time, ncu source page)
  TZ* zn = z + (long)n * 2 * C * GN;
  for (int il = wid; il < ROWS; il += nw) {
    const int i = i_begin + il;
    for (int cb = lane; cb < C / 8; cb += 32) {
      float yi[8], m[8];
      load8(yn_g + ((long)cb * GN + i) * 8, yi);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
      for (int k = 0; k < GK; ++k) {
        const int jn = s_idx[il * GK + k];
        const float nj = s_nrm[jn];
        const float4 p0 = *reinterpret_cast<const float4*>(s_yn + jn * LD + cb * 8);
        const float4 p1 = *reinterpret_cast<const float4*>(s_yn + jn * LD + cb * 8 + 4);
        const float yj8[8] = {p0.x * nj, p0.y * nj, p0.z * nj, p0.w * nj, p1.x * nj, p1.y * nj, p1.z * nj, p1.w * nj};
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], yj8[j] - yi[j]);
      }
      const float lo[8] = {yi[0], m[0], yi[1], m[1], yi[2], m[2], yi[3], m[3]};
      const float hi[8] = {yi[4], m[4], yi[5], m[5], yi[6], m[6], yi[7], m[7]};
      store8(zn + ((long)(2 * cb) * GN + i) * 8, lo);
      store8(zn + ((long)(2 * cb + 1) * GN + i) * 8, hi);
    }
  }
}

// ================================================================================================
// C ABI
// ================================================================================================
static int launch_add_pos(const void* in, long in_img_stride, const float* pos, float* out, void* split, int N, int C,
                          int dtype, cudaStream_t stream) {
  UNCL_REQUIRE(C % 8 == 0 && N > 0, "gcn_add_pos: bad shape");
  const long total = (long)N * C * GN;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  UNCL_DISPATCH_DTYPE(dtype, T, (gcn_add_pos_kernel<T><<<grid, 256, 0, stream>>>((const T*)in, in_img_stride, pos, out, (bf16*)split, C, N)));
  return uncl_check_launch("gcn_add_pos");
}

extern "C" int uncl_gcn_add_pos(const void* in, long in_img_stride, const float* pos, float* out, int N, int C,
                                int dtype, cudaStream_t stream) {
  return launch_add_pos(in, in_img_stride, pos, out, nullptr, N, C, dtype, stream);
}

extern "C" int uncl_gcn_add_pos_split(const void* in, long in_img_stride, const float* pos, float* out, void* split, int N,
                                      int C, int dtype, cudaStream_t stream) {
  UNCL_REQUIRE(split != nullptr, "gcn_add_pos_split: no split output");
  return launch_add_pos(in, in_img_stride, pos, out, split, N, C, dtype, stream);
}

// out dtype: fp32 or bf16 (the last FFN conv writes the generator's working dtype)
extern "C" int uncl_pw_conv(const float* in, const float* w, const float* bias, const float* res, const float* scale,
                            void* out, long out_img_stride, int N, int C_in, int C_out, int groups, int HW, int act,
                            int out_dtype, cudaStream_t stream) {
  UNCL_REQUIRE(groups > 0 && C_in % groups == 0 && C_out % groups == 0 && (C_in / groups) % 16 == 0 &&
                   (C_out / groups) % 32 == 0 && N > 0,
               "pw_conv: unsupported C_in=%d C_out=%d groups=%d", C_in, C_out, groups);
  // 128-wide channel tiles when that still fills the machine, else 64-wide (twice the CTAs)
  const int px_tiles = ceil_div(N * HW, PW_PX);
  const bool wide = (C_out / groups) % 128 == 0 && px_tiles * (C_out / 128) >= 120;   // (64-wide tiles measured slower: 72 -> 77 us on fc1)
  if (wide) {
    dim3 grid(px_tiles, C_out / 128);
    UNCL_DISPATCH_DTYPE(out_dtype, T, (pw_conv_kernel<T, 128><<<grid, 256, 0, stream>>>(in, w, bias, res, scale, (T*)out, out_img_stride, C_in, C_out, groups, HW, N, act)));
  } else if ((C_out / groups) % 64 == 0) {
    dim3 grid(px_tiles, C_out / 64);
    UNCL_DISPATCH_DTYPE(out_dtype, T, (pw_conv_kernel<T, 64><<<grid, 256, 0, stream>>>(in, w, bias, res, scale, (T*)out, out_img_stride, C_in, C_out, groups, HW, N, act)));
  } else {
    dim3 grid(px_tiles, C_out / 32);
    UNCL_DISPATCH_DTYPE(out_dtype, T, (pw_conv_kernel<T, 32><<<grid, 256, 0, stream>>>(in, w, bias, res, scale, (T*)out, out_img_stride, C_in, C_out, groups, HW, N, act)));
  }
  return uncl_check_launch("pw_conv");
}

extern "C" int uncl_gcn_knn_aggregate(const float* y, const float* relpos, void* z, int z_dtype, int* idx_out, int N, int C,
                                      cudaStream_t stream) {
  UNCL_REQUIRE(C == 256 && N > 0, "gcn_knn_aggregate: only C=256 (shipped config) is built, got %d", C);
  const size_t smem = (size_t)(GN * (C + 4) + 2 * GN) * sizeof(float) + (size_t)GN * GK * sizeof(int);
  cudaError_t e = cudaFuncSetAttribute(gcn_knn_agg_kernel<256, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gcn_knn_agg_kernel<256, bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "gcn_knn_aggregate: smem attr: %s", cudaGetErrorString(e));
  UNCL_DISPATCH_DTYPE(z_dtype, T, (gcn_knn_agg_kernel<256, T><<<N * KNN_SPLIT, 512, smem, stream>>>(y, relpos, (T*)z, idx_out)));
  return uncl_check_launch("gcn_knn_aggregate");
}
