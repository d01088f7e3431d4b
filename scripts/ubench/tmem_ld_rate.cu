// Micro-benchmark: tcgen05.ld throughput (4 warps, 32x32b.x16) alone and while another warp issues tcgen05.mma
// back to back, and the MMA rate while the loads run.   Zeros as operands; only timing matters.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace selfc::tc;
namespace selfc { namespace tc { bool pdl_enabled() { return false; } } }

struct Cfg { int n, do_mma, do_ld, iters; };

__global__ void __launch_bounds__(192, 1) k(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 192) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && c.do_mma) {
    const uint32_t idesc = umma_idesc_bf16(128, c.n);
    const uint32_t a0 = base, b0 = base + 64 * 1024;
    const uint32_t hi_a = desc_hi(256, 6), hi_b = desc_hi(128, 0);
    const long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      const uint64_t ad = desc_join(desc_lo(a0 + (uint32_t)(i & 15) * 4096u, 16), hi_a);
      const uint64_t bd = desc_join(desc_lo(b0, (uint32_t)c.n * 16), hi_b);
      umma_bf16_elect(tmem + 256u + (uint32_t)((i & 1) * 128), ad, bd, idesc, i >= 2 ? 1u : 0u);   // accumulators in columns 256..511
    }
    umma_commit_elect(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  if (warp >= 2 && c.do_ld) {
    const int q = warp & 3;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      // 96 columns = 6 x16 loads per "tile row", columns 0..95 (not touched by the MMAs)
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        uint32_t r[16];
        tmem_ld16(lane_addr + (uint32_t)(16 * j), r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) acc ^= r[e];
      }
    }
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0 && warp == 2) out[1] = t1 - t0;
    if (acc == 0x12345678u) out[2] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("4 warps x tcgen05.ld 32x32b.x16 of 96 columns (48 KB per iteration over the SM) vs tcgen05.mma M=128 K=16\n");
  printf("%5s %6s %6s %14s %18s\n", "N", "mma", "ld", "cyc/MMA", "cyc per 48 KB ld");
  for (int n : {48, 96})
    for (int mode = 0; mode < 3; ++mode) {
      Cfg c{n, mode != 1, mode != 0, 2000};
      long long h[2] = {0, 0};
      cudaMemset(d, 0, 64);
      k<<<148, 192, 100 * 1024>>>(c, d);
      cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("%5d %6d %6d %14.1f %18.1f\n", n, c.do_mma, c.do_ld, (double)h[0] / c.iters, (double)h[1] / c.iters);
    }
  return 0;
}
