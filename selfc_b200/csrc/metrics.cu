// Quality metrics of the validation loop on the GPU: BT.601 luma, per-frame squared error (PSNR) and SSIM.
//
// Reference behaviour restated (not copied): data/util.py:239-245 (rgb_to_ycbcr: Y = (65.481 R + 128.553 G + 24.966 B + 16) / 255),
// utils/util.py:198-221 (calculate_psnr: per frame, mean over C,H,W, 20 log10(1 / sqrt(mse))), utils/util.py:361-488 and
// :596-603 (calculate_ssim: per frame, 11-tap sigma 1.5 separable Gaussian, VALID convolution per channel, K1 = 0.01,
// K2 = 0.03, data_range 1, mean of the SSIM map over C x (H-10) x (W-10)).
//
// One fused kernel per pair of clips: the (optional) luma conversion happens on load, the five blurred moments
// (x, y, xx, yy, xy) of a 16x16 output tile are built in shared memory from one 26x26 input tile, the SSIM map is never
// written, and the per-frame sums are accumulated in fp64.  The host divides and takes the logarithm.
#include "common.cuh"
#include "kernels.h"

namespace selfc {

__device__ __forceinline__ float luma601(float r, float g, float b) {
  // products and sums rounded one by one, as the eager CPU expression does (no FMA contraction)
  float y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, 65.481f), __fmul_rn(g, 128.553f)), __fmul_rn(b, 24.966f)), 16.0f);
  return __fdiv_rn(y, 255.0f);
}

__global__ void __launch_bounds__(256) rgb_to_y_kernel(const float* __restrict__ x, float* __restrict__ y, long long N, long long HW) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N * HW) return;
  const long long n = p / HW, q = p - n * HW;
  const float* s = x + n * 3 * HW + q;
  y[p] = luma601(s[0], s[HW], s[2 * HW]);
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  __syncthreads();
  return t;
}

// sse[n] += sum over the frame of (a - b)^2; C channels per frame (to_y: the inputs have 3 channels, the metric 1)
__global__ void __launch_bounds__(256) frame_sse_kernel(const float* __restrict__ a, const float* __restrict__ b, int C, long long HW,
                                                        int to_y, double* __restrict__ sse) {
  __shared__ double red[8];
  const int n = blockIdx.y;
  const long long total = to_y ? HW : (long long)C * HW;
  const float* pa = a + (long long)n * C * HW;
  const float* pb = b + (long long)n * C * HW;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float va, vb;
    if (to_y) {
      va = luma601(pa[i], pa[HW + i], pa[2 * HW + i]);
      vb = luma601(pb[i], pb[HW + i], pb[2 * HW + i]);
    } else {
      va = pa[i];
      vb = pb[i];
    }
    const float d = va - vb;
    acc += (double)(d * d);
  }
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(sse + n, t);
}

constexpr int SS_T = 16;             // output tile
constexpr int SS_W = 11;             // window
constexpr int SS_IN = SS_T + SS_W - 1;

// grid: (tiles_x, tiles_y, N * Ceff); ssim_sum[n] += sum of the SSIM map of this tile
__global__ void __launch_bounds__(SS_T * SS_T) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int C, int H, int W,
                                                            int to_y, const float* __restrict__ win, double* __restrict__ ssim_sum) {
  __shared__ float sx[SS_IN][SS_IN + 1], sy[SS_IN][SS_IN + 1];
  __shared__ float hz[5][SS_IN][SS_T + 1];
  __shared__ float w[SS_W];
  __shared__ double red[8];
  const int ceff = to_y ? 1 : C;
  const int n = blockIdx.z / ceff, c = blockIdx.z % ceff;
  const long long HW = (long long)H * W;
  const float* pa = a + ((long long)n * C + (to_y ? 0 : c)) * HW;
  const float* pb = b + ((long long)n * C + (to_y ? 0 : c)) * HW;
  const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
  const int tid = threadIdx.x;
  if (tid < SS_W) w[tid] = win[tid];
  for (int i = tid; i < SS_IN * SS_IN; i += SS_T * SS_T) {
    const int ly = i / SS_IN, lx = i % SS_IN;
    const int gy = y0 + ly, gx = x0 + lx;
    float va = 0.f, vb = 0.f;
    if (gy < H && gx < W) {
      const long long o = (long long)gy * W + gx;
      if (to_y) {
        va = luma601(pa[o], pa[HW + o], pa[2 * HW + o]);
        vb = luma601(pb[o], pb[HW + o], pb[2 * HW + o]);
      } else {
        va = pa[o];
        vb = pb[o];
      }
    }
    sx[ly][lx] = va;
    sy[ly][lx] = vb;
  }
  __syncthreads();
  // horizontal pass: 26 rows x 16 columns x 5 moments
  for (int i = tid; i < SS_IN * SS_T; i += SS_T * SS_T) {
    const int ly = i / SS_T, lx = i % SS_T;
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < SS_W; ++k) {
      const float u = sx[ly][lx + k], v = sy[ly][lx + k], g = w[k];
      m[0] += g * u;
      m[1] += g * v;
      m[2] += g * (u * u);
      m[3] += g * (v * v);
      m[4] += g * (u * v);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) hz[q][ly][lx] = m[q];
  }
  __syncthreads();
  const int ty = tid / SS_T, tx = tid % SS_T;
  double val = 0.0;
  if (y0 + ty < H - (SS_W - 1) && x0 + tx < W - (SS_W - 1)) {
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < SS_W; ++k) {
      const float g = w[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] += g * hz[q][ty + k][tx];
    }
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float mu1 = m[0], mu2 = m[1];
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float s1 = m[2] - mu1_sq, s2 = m[3] - mu2_sq, s12 = m[4] - mu12;
    const float cs = (2.f * s12 + C2) / (s1 + s2 + C2);
    val = (double)(((2.f * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs);
  }
  const double t = block_sum(val, red);
  if (tid == 0) atomicAdd(ssim_sum + n, t);
}

int launch_rgb_to_y(const float* x, float* y, long long N, long long HW, cudaStream_t st) {
  if (N * HW == 0) return 0;
  rgb_to_y_kernel<<<cdiv(N * HW, 256), 256, 0, st>>>(x, y, N, HW);
  SELFC_LAUNCH_CHECK("rgb_to_y_kernel");
  return 0;
}

int launch_frame_metrics(const float* a, const float* b, int N, int C, int H, int W, int to_y, const float* win11, double* sse,
                         double* ssim_sum, cudaStream_t st) {
  if (N == 0) return 0;
  const long long HW = (long long)H * W;
  SELFC_CUDA(cudaMemsetAsync(sse, 0, (size_t)N * sizeof(double), st));
  int gx = (int)cdiv((to_y ? HW : (long long)C * HW), 256 * 8);
  if (gx > 592) gx = 592;
  if (gx < 1) gx = 1;
  frame_sse_kernel<<<dim3(gx, N), 256, 0, st>>>(a, b, C, HW, to_y, sse);
  SELFC_LAUNCH_CHECK("frame_sse_kernel");
  if (ssim_sum != nullptr) {
    SELFC_CUDA(cudaMemsetAsync(ssim_sum, 0, (size_t)N * sizeof(double), st));
    const int oh = H - (SS_W - 1), ow = W - (SS_W - 1);
    const int ceff = to_y ? 1 : C;
    ssim_kernel<<<dim3(cdiv(ow, SS_T), cdiv(oh, SS_T), N * ceff), SS_T * SS_T, 0, st>>>(a, b, C, H, W, to_y, win11, ssim_sum);
    SELFC_LAUNCH_CHECK("ssim_kernel");
  }
  return 0;
}

}  // namespace selfc
