// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of N, the A/B shared-memory layouts and
// whether consecutive MMAs read a fresh A tile / write a fresh accumulator.  Operands are zeros; only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I selfc_b200/csrc scripts/ubench/mma_rate.cu -o gpurun_out/mma_rate -lcuda
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace selfc::tc;
namespace selfc { namespace tc { bool pdl_enabled() { return false; } } }

struct Cfg { int n, a_layout, a_fresh, d_rot, iters, b_lbo; };

__global__ void __launch_bounds__(64, 1) k(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, c.n);
    const uint32_t a0 = base, b0 = base + 64 * 1024;
    // A: 16 tiles of 4 KB; layout 6 = SW32 (SBO 256), 2 = SW128 (SBO 1024), 0 = no-swizzle core matrices (LBO 2048, SBO 128)
    const uint32_t hi_a = c.a_layout == 6 ? desc_hi(256, 6) : (c.a_layout == 2 ? desc_hi(1024, 2) : desc_hi(128, 0));
    const uint32_t lbo_a = c.a_layout == 0 ? 2048 : 16;
    const uint32_t hi_b = desc_hi(128, 0);
    const long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      const uint32_t a_off = c.a_fresh ? (uint32_t)(i & 15) * 4096u : 0u;
      const uint32_t d = tmem + (uint32_t)((i & (c.d_rot - 1)) * c.n);      // d_rot = 1, 2 or 4 accumulators in rotation
      const uint64_t ad = desc_join(desc_lo(a0 + a_off, lbo_a), hi_a);
      const uint64_t bd = desc_join(desc_lo(b0, (uint32_t)c.b_lbo), hi_b);
      umma_bf16_elect(d, ad, bd, idesc, i >= c.d_rot ? 1u : 0u);
    }
    umma_commit_elect(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int ns[] = {16, 32, 48, 64, 96, 128, 256};
  printf("cycles per M=128,K=16 MMA (2000 back-to-back, 1 CTA per SM on 148 SMs)\n");
  printf("%5s %8s %7s %6s %10s\n", "N", "A layout", "A fresh", "D rot", "cyc/MMA");
  for (int layout : {6})
    for (int n : ns)
      for (int fresh : {1, 0})
        for (int rot : {1, 2, 4}) {
          if (rot * n > 512) continue;
          Cfg c{n, layout, fresh, rot, 2000, n * 16};
          k<<<148, 64, 100 * 1024>>>(c, d);
          long long h = 0;
          cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("%5d %8s %7d %6d %10.1f\n", n, layout == 6 ? "SW32" : (layout == 2 ? "SW128" : "none"), fresh, rot, (double)h / c.iters);
        }
  return 0;
}
