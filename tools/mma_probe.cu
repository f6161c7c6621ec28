// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, bf16) as a function of
//   * N,
//   * the shared-memory operand layout (no-swizzle "interleave" as conv_tc.cu uses it, or SWIZZLE_128B K-major),
//   * which operand changes between consecutive instructions (A / B / both / neither) and how many accumulators rotate,
//   * the A operand's home (shared memory or TMEM).
// Operands are resident (no TMA in flight), one thread issues, every SM runs the same loop.  Timing only: the data is noise.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I uncltmo_b200/csrc -I include -o tools/mma_probe tools/mma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_ptx.cuh"

using namespace tcptx;

struct Cfg {
  int N;            // MMA N
  int swz;          // 0: no-swizzle (LBO/SBO core matrices), 1: SWIZZLE_128B K-major
  int G;            // outer iterations ("taps x chunks")
  int MB;           // inner iterations (M blocks); accumulator of inner m = (m % nD) * N
  int nD;
  int a_g, a_m;     // A start-address step (bytes) per outer / inner iteration (wrapped by a_wrap groups)
  int b_g, b_m;     // same for B
  int a_tmem;       // 1: A operand from TMEM
  int a_shift;      // extra byte offset of A (tap-shift emulation)
  int a_gmask, b_gmask;   // outer index is masked with these before scaling
};

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 db;\n"
      "setp.ne.b32 p, 1, 0;\n"
      "mov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(Cfg c, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // noise operands (small bf16 values)
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u;
    w[i] = 0x3c003c00u | (h & 0x007f007fu) | ((h >> 3) & 0x80008000u);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    // the whole warp walks the uniform loop (descriptors stay in uniform registers), one elected lane issues
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t a_hi, b_hi, a_lbo, b_lbo;
    if (c.swz) {
      a_hi = b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      a_lbo = b_lbo = 1u << 16;
    } else {
      a_hi = b_hi = (128u >> 4) | (1u << 14);                 // SBO 128 B
      a_lbo = ((40u * 1024u) >> 4) << 16;                     // the two 8-channel halves of A are 40 KB apart
      b_lbo = ((uint32_t)c.N) << 16;                          // LBO_B = N * 16 B
    }
    const uint32_t a_base = (smem_u32(smem) + (uint32_t)c.a_shift) >> 4;
    const uint32_t b_base = (smem_u32(smem) + 96u * 1024u) >> 4;
    const uint32_t a_tm = tmem + 448u;   // 8 columns of TMEM hold a 128 x 16 bf16 A tile (TS mode)
    const uint32_t a_g = (uint32_t)c.a_g >> 4, b_g = (uint32_t)c.b_g >> 4, a_m = (uint32_t)c.a_m >> 4, b_m = (uint32_t)c.b_m >> 4;
    const uint32_t n = (uint32_t)c.N, nD = (uint32_t)c.nD;
    const int G = c.G, MB = c.MB, a_gmask = c.a_gmask, b_gmask = c.b_gmask;
    const bool ts = c.a_tmem != 0, ts_walk = c.a_m != 0;
    const long long t0 = clock64();
    const uint32_t d_step = nD > 1 ? n : 0u;      // nD is either 1 or MB in every configuration below
    const uint32_t at_step = ts_walk ? 8u : 0u;
    if (elect_one()) {
      if (ts) {
        for (int g = 0; g < G; ++g) {
          uint32_t b = b_lbo | (b_base + (uint32_t)(g & b_gmask) * b_g);
          uint32_t d = tmem, at = a_tm;
          for (int i = 0; i < MB; ++i) {
            mma_ts(d, at, b, b_hi, idesc);
            b += b_m; d += d_step; at += at_step;
          }
        }
      } else {
        for (int g = 0; g < G; ++g) {
          uint32_t a = a_lbo | (a_base + (uint32_t)(g & a_gmask) * a_g);
          uint32_t b = b_lbo | (b_base + (uint32_t)(g & b_gmask) * b_g);
          uint32_t d = tmem;
          for (int i = 0; i < MB; ++i) {
            tc_mma_bf16(d, a, a_hi, b, b_hi, idesc, 1u);
            a += a_m; b += b_m; d += d_step;
          }
        }
      }
    }
    __syncwarp();
    if (elect_one()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

static unsigned long long* d_out;
static int g_grid = 148;

static double run(const Cfg& c) {
  std::vector<unsigned long long> h(g_grid);
  for (int rep = 0; rep < 2; ++rep) {
    probe<<<g_grid, 128, 202 * 1024>>>(c, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  cudaMemcpy(h.data(), d_out, g_grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < g_grid; ++i) s += (double)h[i];
  return s / g_grid / ((double)c.G * c.MB);
}

int main(int argc, char** argv) {
  if (argc > 1) g_grid = atoi(argv[1]);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
  cudaMalloc(&d_out, 1024 * sizeof(unsigned long long));
  printf("grid %d\n", g_grid);
  printf("layout,N,pattern,MB,nD,cycles_per_mma,floor\n");
  const int Ns[] = {32, 64, 96, 128, 192, 256};
  for (int swz = 0; swz < 2; ++swz) {
    for (int N : Ns) {
      const int mbmax = 512 / N;
      // A block step (next 128 rows): no-swizzle 128 rows x 16 B = 2 KB, SW128 128 rows x 128 B = 16 KB
      const int a_blk = swz ? 16384 : 2048;
      // A tap / K step: no-swizzle +16 B (one pixel) ; SW128 +32 B (next K=16 inside the 128 B row)
      const int a_tap = swz ? 32 : 16;
      // B step per tap: no-swizzle 2 * N * 16 B ; SW128 +32 B inside the atom
      const int b_tap = swz ? 32 : 2 * N * 16;
      const int b_gmask = swz ? 3 : (N <= 128 ? 7 : 3);
      const int a_mmask_blocks = swz ? 4 : 16;   // A region holds this many 128-row blocks
      for (int MB = 1; MB <= mbmax; MB *= 2) {
        if (MB > a_mmask_blocks) break;
        const int G = 2048 / MB;
        // 1. conv_tc pattern: B stationary over the inner loop, A walks the M blocks, MB accumulators
        Cfg c1 = {N, swz, G, MB, MB, a_tap, a_blk, b_tap, 0, 0, 0, 3, b_gmask};
        printf("%s,%d,B-stationary(A walks; acc=MB),%d,%d,%.1f,%d\n", swz ? "sw128" : "none", N, MB, MB, run(c1), N / 2);
      }
      // 2. everything fixed: same A, same B, one accumulator / rotating accumulators
      for (int nD : {1, mbmax}) {
        Cfg c2 = {N, swz, 2048 / (nD > 1 ? nD : 8), nD > 1 ? nD : 8, nD, 0, 0, 0, 0, 0, 0, 0, 0};
        printf("%s,%d,same A same B,%d,%d,%.1f,%d\n", swz ? "sw128" : "none", N, c2.MB, nD, run(c2), N / 2);
      }
      // 3. GEMM pattern: A and B both change every instruction, one accumulator (K loop) / two
      for (int nD : {1, 2}) {
        if (nD * N > 512) continue;
        Cfg c3 = {N, swz, 2048, 1, 1, a_tap, 0, b_tap, 0, 0, 0, 3, b_gmask};
        if (nD == 2) { c3.G = 1024; c3.MB = 2; c3.nD = 2; c3.a_m = a_tap; c3.b_m = b_tap; c3.a_g = 2 * a_tap; c3.b_g = 2 * b_tap; c3.a_gmask = 1; c3.b_gmask = 1; }
        printf("%s,%d,A and B change (K loop),%d,%d,%.1f,%d\n", swz ? "sw128" : "none", N, c3.MB, nD, run(c3), N / 2);
      }
      // 4. A stationary, B changes every instruction (tap-merged / transposed formulations), acc rotates
      for (int nD : {1, 2, 4}) {
        if (nD * N > 512) continue;
        Cfg c4 = {N, swz, 2048 / nD, nD, nD, a_tap, 0, 0, b_tap, 0, 0, 3, 0};
        printf("%s,%d,A-stationary(B walks; acc=MB),%d,%d,%.1f,%d\n", swz ? "sw128" : "none", N, nD, nD, run(c4), N / 2);
      }
      // 5. B stationary but ONE accumulator (A walks): separates "B switch" from "accumulator switch"
      {
        Cfg c5 = {N, swz, 512, 4, 1, a_tap, a_blk, b_tap, 0, 0, 0, 3, b_gmask};
        printf("%s,%d,B-stationary(A walks; acc=1),%d,%d,%.1f,%d\n", swz ? "sw128" : "none", N, 4, 1, run(c5), N / 2);
      }
      // 6. A from TMEM, B changes every instruction
      {
        Cfg c6 = {N, swz, 2048, 1, 1, 0, 0, b_tap, 0, 1, 0, 0, b_gmask};
        printf("%s,%d,A in TMEM; B changes,%d,%d,%.1f,%d\n", swz ? "sw128" : "none", N, 1, 1, run(c6), N / 2);
        Cfg c7 = {N, swz, 512, (384 / N < 4 ? (384 / N < 1 ? 1 : 384 / N) : 4), (384 / N < 4 ? (384 / N < 1 ? 1 : 384 / N) : 4), 0, 1, b_tap, 0, 1, 0, 0, b_gmask};
        printf("%s,%d,A in TMEM walks; B stationary,%d,%d,%.1f,%d\n", swz ? "sw128" : "none", N, c7.MB, c7.nD, run(c7), N / 2);
      }
      // 7. tap shift alignment (no-swizzle: +16 B is what conv_tc does; SW128: +128 B = one pixel row of the atom)
      if (swz) {
        Cfg c8 = {N, swz, 2048 / mbmax, mbmax > 4 ? 4 : mbmax, mbmax > 4 ? 4 : mbmax, a_tap, a_blk, b_tap, 0, 0, 128, 3, b_gmask};
        printf("%s,%d,B-stationary; A start +128 B (row shift),%d,%d,%.1f,%d\n", "sw128", N, c8.MB, c8.nD, run(c8), N / 2);
      }
    }
  }
  return 0;
}
