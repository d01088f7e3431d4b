// Micro-benchmark: TMA tiled-load throughput per SM as a function of the box's inner row size (SWIZZLE_32B / 64B / 128B), for the
// halo-tile access pattern of conv3x3_tc3 (10 rows x 32 positions per tile, tiles walked band-major over a 7 x 270 x 480 clip).
// One producer thread per CTA keeps NST boxes of ~20 KB in flight; no consumer.  Reports bytes / cycle / SM and TB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I selfc_b200/csrc scripts/ubench/tma_rate.cu -o /tmp/tma_rate -lcuda
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace selfc::tc;
namespace selfc { namespace tc { bool pdl_enabled() { return false; } } }

struct Cfg { int cpr; int nslab_box; int nst; int iters; int hot; int tiles_x, tiles_y, N; int box_bytes; int box_y; };

__device__ __forceinline__ void tma5(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__global__ void __launch_bounds__(32, 1) k(const __grid_constant__ CUtensorMap tmap, Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.nst; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ntiles = c.tiles_x * c.tiles_y * c.N;
    const long long t0 = clock64();
    int tile = blockIdx.x;
    for (int i = 0; i < c.iters; ++i) {
      const int s = i % c.nst;
      if (i >= c.nst) mbar_wait(smem_u32(&bars[s]), ((i / c.nst) - 1) & 1, nullptr, 0);
      const int t = c.hot ? (int)blockIdx.x : tile;
      const int tx = t % c.tiles_x, n = (t / c.tiles_x) % c.N, ty = t / (c.tiles_x * c.N);
      mbar_expect_tx(smem_u32(&bars[s]), (uint32_t)c.box_bytes);
      tma5(base + s * c.box_bytes, &tmap, smem_u32(&bars[s]), 0, tx * 30 - 1, ty * (c.box_y - 2) - 1, n, (i % 4) * c.nslab_box);
      tile += gridDim.x;
      if (tile >= ntiles) tile -= ntiles;
    }
    for (int i = c.iters; i < c.iters + c.nst; ++i) {
      const int s = i % c.nst;
      mbar_wait(smem_u32(&bars[s]), ((i / c.nst) - 1) & 1, nullptr, 0);
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

int main() {
  const int N = 7, H = 270, W = 480, CH = 256;          // 256 channels of bf16 per pixel in total, as slabs of cpr channels
  const size_t bytes = (size_t)N * H * W * CH * 2;
  void* buf;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  EncodeTiledFn encode = nullptr;
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    encode = (EncodeTiledFn)p;
  }
  printf("%8s %6s %6s %5s %5s %12s %10s\n", "row B", "box y", "box KB", "nst", "hot", "B/clk/SM", "TB/s@1.9");
  for (int cpr : {16, 32, 64}) {                         // channels per slab row: 32 / 64 / 128 bytes
    for (int box_y : {10, 5, 20}) {
      for (int swz : {1, 0}) {
        if (swz == 0 && !(cpr == 16 && box_y == 10)) continue;       // one un-swizzled reference point
        const int nslab = CH / cpr;
        const int nslab_box = 32 / cpr > 0 ? 32 / cpr : 1;
        CUtensorMap tmap;
        const cuuint64_t gdim[5] = {(cuuint64_t)cpr, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)nslab};
        const cuuint64_t gstr[4] = {(cuuint64_t)cpr * 2, (cuuint64_t)W * cpr * 2, (cuuint64_t)H * W * cpr * 2, (cuuint64_t)N * H * W * cpr * 2};
        const cuuint32_t box[5] = {(cuuint32_t)cpr, 32, (cuuint32_t)box_y, 1, (cuuint32_t)nslab_box};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        const CUtensorMapSwizzle sw = !swz ? CU_TENSOR_MAP_SWIZZLE_NONE
                                           : (cpr == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (cpr == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B));
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int box_bytes = cpr * 2 * 32 * box_y * nslab_box;
        for (int nst : {4}) {
          if (nst * box_bytes > 190 * 1024) continue;
          for (int hot : {1, 0}) {
            Cfg c{cpr, nslab_box, nst, 4000, hot, 16, 270 / (box_y - 2) + 1, N, box_bytes, box_y};
            k<<<148, 32, 200 * 1024>>>(tmap, c, d);
            long long h = 0;
            cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            const double bpc = (double)c.iters * box_bytes / (double)h;
            printf("%7d%s %6d %6.1f %5d %5d %12.2f %10.2f\n", cpr * 2, swz ? " " : "n", box_y, box_bytes / 1024.0, nst, hot, bpc, bpc * 148 * 1.9e9 / 1e12);
          }
        }
      }
    }
  }
  return 0;
}
