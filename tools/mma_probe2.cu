// Micro-benchmark: cycles per tcgen05.mma.cta_group::2 (kind::f16, M = 256 over a CTA pair, K = 16, bf16) as a function
// of N, next to the one-CTA figures of tools/mma_probe.cu.  Each CTA of the pair holds its own 128-row A tile and HALF of
// the B tile (N/2 rows), so per instruction an SM reads 4 KB of A + 16 N bytes of B instead of 4 KB + 32 N.
// Operands are resident (noise), one thread of the leader CTA issues, every pair runs the same loop.  Timing only.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I uncltmo_b200/csrc -I include -o tools/mma_probe2 tools/mma_probe2.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_ptx.cuh"

using namespace tcptx;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_mma2_bf16(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

struct Cfg {
  int N, G, MB;   // MMA N, outer iterations (taps x chunks), M-block pairs per outer iteration (accumulators rotate)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2(Cfg c, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)(i + 7 * rank) * 2654435761u;
    w[i] = 0x3c003c00u | (h & 0x007f007fu) | ((h >> 3) & 0x80008000u);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    long long t0 = 0, t1 = 0;
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t hi = (128u >> 4) | (1u << 14);                       // SBO 128 B
      const uint32_t a_lbo = ((40u * 1024u) >> 4) << 16;                  // the two 8-channel halves of A are 40 KB apart
      const uint32_t b_lbo = ((uint32_t)(c.N / 2)) << 16;                 // per-CTA half of B: LBO = (N/2) * 16 B
      const uint32_t a_base = smem_u32(smem) >> 4, b_base = (smem_u32(smem) + 96u * 1024u) >> 4;
      const uint32_t n = (uint32_t)c.N;
      const int G = c.G, MB = c.MB;
      t0 = clock64();
      if (elect_one()) {
        for (int g = 0; g < G; ++g) {
          uint32_t a = a_lbo | (a_base + (uint32_t)(g & 3));                               // tap shift: +16 B
          const uint32_t b = b_lbo | (b_base + (uint32_t)(g & 7) * (uint32_t)(c.N / 2) * 2u);   // next tap's B tile
          uint32_t d = tmem;
          for (int i = 0; i < MB; ++i) {
            tc_mma2_bf16(d, a, hi, b, hi, idesc, 1u);
            a += 128u; d += n;                                                              // next M block, next accumulator
          }
        }
      }
      __syncwarp();
      if (elect_one()) tc_commit2(&bar, 3);
      __syncwarp();
    }
    mbar_wait(&bar, 0);      // the leader's commit arrives on both CTAs' barriers
    t1 = clock64();
    if (lane == 0 && rank == 0) out[blockIdx.x / 2] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

int main(int argc, char** argv) {
  int grid = argc > 1 ? atoi(argv[1]) : 148;
  grid &= ~1;
  cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
  unsigned long long* d_out;
  cudaMalloc(&d_out, 1024 * sizeof(unsigned long long));
  printf("grid %d (pairs %d)\nN,MB,cycles_per_2cta_mma,cycles_per_128_rows,one_cta_model\n", grid, grid / 2);
  const int Ns[] = {32, 64, 96, 128, 192, 256};
  for (int N : Ns) {
    for (int MB = 1; MB <= 512 / N && MB <= 4; MB *= 2) {
      Cfg c{N, 2048 / MB, MB};
      std::vector<unsigned long long> h(grid / 2);
      for (int rep = 0; rep < 2; ++rep) {
        probe2<<<grid, 128, 202 * 1024>>>(c, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h.data(), d_out, (grid / 2) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < grid / 2; ++i) s += (double)h[i];
      const double cyc = s / (grid / 2) / 2048.0;
      const double one = N / 2.0 > (4096 + 32.0 * N) / 128 ? N / 2.0 : (4096 + 32.0 * N) / 128;
      printf("%d,%d,%.1f,%.1f,%.1f\n", N, MB, cyc, cyc / 2, one);
    }
  }
  return 0;
}
