// Semantics check + rate of tcgen05.mma with the A operand in TENSOR MEMORY (kind::f16, M = 128 per CTA, K = 16), for
// cta_group::1 and cta_group::2.  The fused dense-block kernel keeps the growth channels of a dense block as A operands in
// TMEM (written by the epilogue with tcgen05.st), so it needs to know (1) which bits of which column hold A[m][k] and (2) the
// issue rate against the shared-memory A form (scripts/ubench/mma_rate.cu, mma2_rate.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I selfc_b200/csrc scripts/ubench/mma_tmem_a.cu -o /tmp/mma_tmem_a -lcuda
// Layout probed: lane = row m, 8 consecutive 32-bit columns per K = 16 step, column j = {A[m][2j] (low 16 bits), A[m][2j+1]}.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace selfc::tc;
namespace selfc { namespace tc { bool pdl_enabled() { return false; } } }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t r[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
template <int CG>
__device__ __forceinline__ void umma_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct Cfg { int n, iters, mode, a_rot; };   // mode 0: rate; 1: full numerics; 2 + p: one-hot at packed position p (0..15)

__host__ __device__ inline float a_val(int m, int k) { return (float)(((m * 5 + k) % 13) - 6); }
__host__ __device__ inline float b_val(int n, int k) { return (float)(((n * 7 + k * 3) % 11) - 5); }

template <int CG>
__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* out, float* dout) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int nh = c.n / CG;                 // B rows held by this CTA
  for (int i = threadIdx.x; i < 16 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(gbase)[i] = 0;
  __syncthreads();
  // B: no-swizzle core matrices [kcore(2)][ngroup nh/8][8 rows][8 k]
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(gbase);
  if (c.mode != 0)
    for (int idx = threadIdx.x; idx < nh * 16; idx += 128) {
      const int n = idx / 16, kk = idx % 16;
      const float v = c.mode == 1 ? b_val((int)rank * nh + n, kk) : (float)(kk + 1);
      B[(size_t)((kk / 8) * (nh / 8) + n / 8) * 64 + (n % 8) * 8 + (kk % 8)] = __float2bfloat16(v);
    }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { if (CG == 2) tmem_alloc2(smem_u32(&slot), 512); else tmem_alloc(smem_u32(&slot), 512); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t a_col = 256;              // A tiles: 8 columns each, 16 of them in rotation for the rate test
  // every thread writes its row (lane quarter `warp`) of every A tile
  {
    const int m = warp * 32 + lane;
    for (int t = 0; t < 16; ++t) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) {
        float lo = 0.f, hi = 0.f;
        if (c.mode == 1) { lo = a_val((int)rank * 128 + m, 2 * j); hi = a_val((int)rank * 128 + m, 2 * j + 1); }
        else if (c.mode >= 2) { const int p = c.mode - 2; lo = (p == 2 * j) ? 1.f : 0.f; hi = (p == 2 * j + 1) ? 1.f : 0.f; }
        __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);      // .x = low 16 bits
        r[j] = *reinterpret_cast<uint32_t*>(&v);
      }
      tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + a_col + (uint32_t)t * 8, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(128 * CG, c.n);
    const uint32_t hi = desc_hi(128, 0);
    const uint64_t bd = desc_join(desc_lo(base, (uint32_t)(nh / 8) * 128), hi);
    const long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      const uint32_t a = tmem + a_col + (uint32_t)(c.a_rot ? (i & 15) * 8 : 0);
      umma_ts_elect<CG>(tmem, a, bd, idesc, i > 0 ? 1u : 0u);
    }
    if (CG == 2) umma2_commit_elect(smem_u32(&bar)); else umma_commit_elect(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else {
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
  }
  tc_fence_after();
  if (c.mode != 0 && blockIdx.x < CG) {
    for (int n0 = 0; n0 < c.n; n0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) dout[((size_t)rank * 128 + warp * 32 + lane) * 256 + n0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) { tc_fence_after(); if (CG == 2) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int CG>
static cudaError_t launch(int grid, Cfg c, long long* d, float* dout) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 20 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k<CG>, c, d, dout);
}

template <int CG>
static int run_all(long long* d, float* dout) {
  static float h[256 * 256];
  const int rows = 128 * CG, N = 96;
  {
    Cfg c{N, 1, 1, 0};
    cudaMemset(dout, 0, sizeof(h));
    cudaError_t e = launch<CG>(CG, c, d, dout);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cta_group::%d numerics launch error %s\n", CG, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < rows; ++r)
      for (int n = 0; n < N; ++n) {
        float ref = 0.f;
        for (int kk = 0; kk < 16; ++kk) ref += a_val(r, kk) * b_val(n, kk);
        if (h[r * 256 + n] != ref) { if (bad < 4) printf("  mismatch row %d col %d: got %g want %g\n", r, n, h[r * 256 + n], ref); ++bad; }
      }
    printf("cta_group::%d numerics (A in TMEM: lane = row, column j = {k=2j low half, k=2j+1 high half}): %s (%d mismatches of %d)\n", CG,
           bad ? "NO" : "yes", bad, rows * N);
  }
  printf("cta_group::%d one-hot map, packed position p (column p/2, half p%%2) -> k: ", CG);
  for (int p = 0; p < 16; ++p) {
    Cfg c{N, 1, 2 + p, 0};
    cudaMemset(dout, 0, sizeof(h));
    cudaError_t e = launch<CG>(CG, c, d, dout);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("one-hot launch error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%d->%g(row %d: %g) ", p, h[0] - 1.f, rows - 1, h[(rows - 1) * 256 + N - 1] - 1.f);
  }
  printf("\n");
  printf("cta_group::%d cycles per MMA, A in TMEM (2000 back-to-back, all SMs)\n%5s %6s %10s\n", CG, "N", "A rot", "cyc/MMA");
  for (int n : {32, 48, 64, 96, 128})
    for (int rot : {0, 1}) {
      if (n % (8 * CG * (CG == 2 ? 2 : 1)) != 0 && CG == 2 && n % 16 != 0) continue;
      Cfg c{n, 2000, 0, rot};
      cudaError_t e = launch<CG>(148, c, d, dout);
      long long hh = 0;
      if (e == cudaSuccess) e = cudaMemcpy(&hh, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("%5d %6d %10.1f\n", n, rot, (double)hh / c.iters);
    }
  return 0;
}

int main() {
  long long* d;
  float* dout;
  cudaMalloc(&d, 8);
  cudaMalloc(&dout, 256 * 256 * 4);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
  if (run_all<1>(d, dout)) return 1;
  if (run_all<2>(d, dout)) return 1;
  return 0;
}
