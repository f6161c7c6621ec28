// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (conv_tc.cu, conv_wgrad_tc.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

// Timing probes and per-role cycle counters exist only in the `-DUNCL_PROBES` build that tools/ uses
// (`python -m uncltmo_b200.build --probes` -> libuncltmo_b200_probes.so); the product library carries none of them,
// reads no environment variable and keeps no mutable global.
#ifdef UNCL_PROBES
#define UNCL_PROBE(flags, bit) ((flags) & (bit))
#else
#define UNCL_PROBE(flags, bit) 0
#endif

namespace tcptx {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// descriptors travel as (lo, hi) 32-bit halves: only `lo` (start address | LBO) changes between MMAs
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// K-major, no-swizzle shared-memory matrix descriptor (sm_100 "version 1").
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 4-D tensor map over 8-byte elements of a C8-blocked bf16 tensor: (2*x + half, y, channel block, image).
// One pixel's 8 channels are 2 elements, so a box row of `box_w` pixels is box_w*16 contiguous bytes.
// Encoded maps are kept in a small thread-local direct-mapped cache keyed on (pointer, shape, box): a network replays
// the same ~60 (tensor, tile) pairs every step and the driver call is the most expensive host work of a launch.
struct TmapKey {
  const void* base; long stride; int W, H, Cb, N, bw, bh, bc;
  bool operator==(const TmapKey& o) const {
    return base == o.base && stride == o.stride && W == o.W && H == o.H && Cb == o.Cb && N == o.N && bw == o.bw &&
           bh == o.bh && bc == o.bc;
  }
};
inline CUresult encode_blocked_bf16(CUtensorMap* tmap, const void* base, int W, int H, int Cb, int N, long img_stride_elems,
                                    int box_w, int box_h, int box_cb) {
  constexpr int kSlots = 256;
  struct Slot { TmapKey key; CUtensorMap map; bool used; };
  static thread_local Slot cache[kSlots];
  const TmapKey key{base, img_stride_elems, W, H, Cb, N, box_w, box_h, box_cb};
  size_t hsh = reinterpret_cast<uintptr_t>(base) >> 8;
  hsh = hsh * 1000003u ^ (size_t)(W * 131 + H * 31 + Cb * 7 + box_w * 3 + box_h + box_cb * 17 + N * 1009) ^ (size_t)img_stride_elems;
  Slot& slot = cache[hsh % kSlots];
  if (slot.used && slot.key == key) { *tmap = slot.map; return CUDA_SUCCESS; }
  EncodeTiledFn encode = get_encode();
  if (!encode) return CUDA_ERROR_NOT_SUPPORTED;
  const cuuint64_t gdim[4] = {(cuuint64_t)W * 2, (cuuint64_t)H, (cuuint64_t)Cb, (cuuint64_t)N};
  const cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)img_stride_elems * 2};
  const cuuint32_t box[4] = {(cuuint32_t)box_w * 2, (cuuint32_t)box_h, (cuuint32_t)box_cb, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) { slot.key = key; slot.map = *tmap; slot.used = true; }
  return r;
}

// SM count of the current device (thread-local cache: one attribute query per thread and device)
inline int sm_count() {
  static thread_local int cached_dev = -1, cached_sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached_sms;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (thread, device, kernel): `state` is a thread_local int of the
// caller holding the largest size already granted on `state_dev`
template <typename K>
inline cudaError_t ensure_smem(K kern, int smem_bytes, int& state, int& state_dev) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev == state_dev && smem_bytes <= state) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) { state = 227 * 1024; state_dev = dev; }
  return e;
}

}  // namespace tcptx
