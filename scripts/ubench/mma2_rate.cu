// Micro-benchmark + semantics check for tcgen05.mma.cta_group::2 (kind::f16, M=256 over a CTA pair, K=16).
//  (1) numerics: A_c[r][k] = (k == 0), B_c[n][0] = 48*c + n  ->  which B rows land in which D columns of each CTA
//  (2) rate: cycles per MMA for N in {48..256}, A tiles fresh, accumulators in rotation
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I selfc_b200/csrc scripts/ubench/mma2_rate.cu -o /tmp/mma2_rate -lcuda
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace selfc::tc;
namespace selfc { namespace tc { bool pdl_enabled() { return false; } } }

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_elect(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(bar), "h"(mask)
      : "memory");
}

struct Cfg { int n, iters, d_rot, check; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(Cfg c, long long* out, float* dout) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int nh = c.n / 2;
  // A: 16 tiles of 4 KB (128 rows x 16 k), no-swizzle core matrices [kcore][rowgroup 16][8][8]; B at +64 KB: [kcore][ngroup nh/8][8][8]
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(gbase)[i] = 0;
  __syncthreads();
  if (c.check) {
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(gbase);
    __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(gbase + 64 * 1024);
    for (int r = threadIdx.x; r < 128; r += 128) A[(r / 8) * 64 + (r % 8) * 8 + 0] = __float2bfloat16(1.0f);          // k = 0
    for (int n = threadIdx.x; n < nh; n += 128) B[(n / 8) * 64 + (n % 8) * 8 + 0] = __float2bfloat16((float)(rank * nh + n));
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc2(smem_u32(&slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, c.n);
    const uint32_t a0 = base, b0 = base + 64 * 1024;
    const uint32_t hi = desc_hi(128, 0);
    const long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      const uint32_t a_off = c.check ? 0u : (uint32_t)(i & 15) * 4096u;
      const uint32_t d = tmem + (uint32_t)((i & (c.d_rot - 1)) * c.n);
      const uint64_t ad = desc_join(desc_lo(a0 + a_off, 2048), hi);
      const uint64_t bd = desc_join(desc_lo(b0, (uint32_t)(nh / 8) * 128), hi);
      umma2_bf16_elect(d, ad, bd, idesc, i >= c.d_rot ? 1u : 0u);
    }
    umma2_commit_elect(smem_u32(&bar), 3);
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else {
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);       // the multicast commit arrives on the peer's barrier too
  }
  tc_fence_after();
  if (c.check && blockIdx.x < 2) {
    // lane quarter `warp`: rows 32*warp + lane of this CTA's half of D
    for (int n0 = 0; n0 < c.n; n0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) dout[((size_t)rank * 128 + warp * 32 + lane) * 256 + n0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(tmem, 512); }
}

static cudaError_t launch(int grid, Cfg c, long long* d, float* dout) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 100 * 1024;
  return cudaLaunchKernelEx(&cfg, k, c, d, dout);
}

int main() {
  long long* d;
  float* dout;
  cudaMalloc(&d, 8);
  cudaMalloc(&dout, 256 * 256 * 4);
  cudaMemset(dout, 0, 256 * 256 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  {
    Cfg c{96, 1, 1, 1};
    cudaError_t e = launch(2, c, d, dout);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("check launch error %s\n", cudaGetErrorString(e)); return 1; }
    static float h[256 * 256];
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 256; ++r)
      for (int n = 0; n < 96; ++n)
        if (h[r * 256 + n] != (float)n) ++bad;
    printf("numerics: D[r][n] == n (B rows of CTA0 -> columns 0..47, CTA1 -> 48..95, every row of both CTAs): %s (%d mismatches)\n",
           bad ? "NO" : "yes", bad);
    printf("  row 0 cols 0,1,47,48,95: %g %g %g %g %g   row 128 (CTA1 row 0): %g %g %g %g %g\n", h[0], h[1], h[47], h[48], h[95],
           h[128 * 256], h[128 * 256 + 1], h[128 * 256 + 47], h[128 * 256 + 48], h[128 * 256 + 95]);
  }
  printf("cycles per M=256 (2 x 128),K=16 MMA pair-issue (2000 back-to-back, 74 CTA pairs)\n%5s %6s %10s\n", "N", "D rot", "cyc/MMA");
  for (int n : {48, 64, 96, 128, 192, 256})
    for (int rot : {1, 2, 4}) {
      if (rot * n > 512) continue;
      Cfg c{n, 2000, rot, 0};
      cudaError_t e = launch(148, c, d, dout);
      long long h = 0;
      if (e == cudaSuccess) e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("%5d %6d %10.1f\n", n, rot, (double)h / c.iters);
    }
  return 0;
}
