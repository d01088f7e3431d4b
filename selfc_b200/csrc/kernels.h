// Host-side launcher declarations shared by the translation units of libselfc_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace selfc {

// ---- layout.cu ------------------------------------------------------------------------------------------
int launch_fa_fwd_nchw(const float* x, float* out51, int N, int h, int w, cudaStream_t st);
template <typename T>
int launch_fa_fwd_z(const float* x, float* z, T* fbuf, int fpitch, long long fslabM, int N, int h, int w, cudaStream_t st);
int launch_fa_rev(const float* z, bool z_is_nchw, float* y, int N, int h, int w, cudaStream_t st);
// 2x variants (f3): FrequencyAnalyzer(k=2) [N,3,2h,2w] <-> [N,15,h,w]; HaarDownsampling [N,C,2h,2w] <-> [N,4C,h,w]
int launch_fa2(const float* in, float* out, bool rev, int N, int h, int w, cudaStream_t st);
int launch_haar(const float* in, float* out, bool rev, int N, int C, int h, int w, cudaStream_t st);
// validation metrics (metrics.cu): BT.601 luma; per-frame sum of squared errors and sum of the SSIM map (fp64), to_y = convert
// the 3-channel inputs to luma on load
int launch_rgb_to_y(const float* x, float* y, long long N, long long HW, cudaStream_t st);
int launch_frame_metrics(const float* a, const float* b, int N, int C, int H, int W, int to_y, const float* win11, double* sse,
                         double* ssim_sum, cudaStream_t st);
// 8-bit frames in cv2 layout [N][H][W][3] (B,G,R): fused into the FrequencyAnalyzer kernels (HR side) and stand-alone
template <typename T>
int launch_fa_fwd_z_u8(const uint8_t* x, float* z, T* fbuf, int fpitch, long long fslabM, int N, int h, int w, cudaStream_t st);
int launch_fa_rev_u8(const float* z, uint8_t* y, int N, int h, int w, cudaStream_t st);
int launch_frames_from_u8(const uint8_t* img, float* x, long long N, long long HW, cudaStream_t st);
int launch_frames_to_u8(const float* x, uint8_t* img, long long N, long long HW, cudaStream_t st);
int launch_quantize(const float* x, uint8_t* q8, float* qf, size_t n, cudaStream_t st);
int launch_export_down(const float* z, float* out51, uint8_t* lr_u8, float* lr_q, long long M, long long hw, cudaStream_t st);
int launch_export_hf(const float* z, float* hf, long long M, long long hw, cudaStream_t st);
int launch_gaussian_down(const float* x, const float* k13, float* y, int NC, int H, int W, cudaStream_t st);
// BF16 mode: lr [N,3,h,w] -> quad 0 of z + the X slab (16 channels) of up to three slab-planar dense buffers, one pass
// x2: the buffers hold (hi, lo) bf16 pairs (BF16X3 mode, common.cuh): 64-byte rows [16 x hi | 16 x lo]
int launch_lr_ingest_slab(const float* lr, float* z, __nv_bfloat16* d0, __nv_bfloat16* d1, __nv_bfloat16* d2, long long M, long long hw,
                          cudaStream_t st, bool x2 = false);
template <typename T>
int launch_nchw_to_dense(const float* x, T* dst, int pitch, long long slabM, int off, int C, int cpad, long long M, long long hw,
                         cudaStream_t st);
template <typename T>
int launch_nchw_slice_to_dense(const float* x, int ctot, int c0, T* dst, int pitch, long long slabM, int off, int C, int cpad,
                               long long M, long long hw, cudaStream_t st);
template <typename T>
int launch_dense_to_nchw(const T* src, int pitch, long long slabM, int off, float* y, int C, long long M, long long hw, cudaStream_t st);

// ---- conv_simt.cu: fp32-FMA implicit GEMM (strict-fp32 mode and the small GEMMs of both modes) --------------
enum TapMode { TAP_POINT = 0, TAP_SPATIAL = 1, TAP_TEMPORAL = 2, TAP_TMIX = 3 };
enum Epilogue { EPI_STORE = 0, EPI_COUPLE_Y1 = 1, EPI_COUPLE_S = 2, EPI_COUPLE_Y2 = 3, EPI_GA = 4, EPI_GMM = 5, EPI_COUPLE_HG = 6, EPI_ACCUM = 7 };

template <typename T>
struct ConvArgs {
  // input: dense buffer (common.cuh: pixel-major, or slab-planar when in_slabM != 0), channels [0,cin) consumed (cin % 4 == 0)
  const T* in = nullptr;
  int in_pitch = 0, cin = 0;
  long long in_slabM = 0;
  // packed weights [taps*cin][np] fp32 (np % 32 == 0) + bias [np]
  const float* w = nullptr;
  const float* bias = nullptr;
  int np = 0, cout = 0;
  int taps = 1, tap_mode = TAP_POINT, in_lrelu = 0;
  int BT = 0, Tn = 1, h = 0, w_ = 0;
  // TAP_TMIX: temporal mixing matrix [B][T][T] (GlobalAgg), column sums [B][T]
  const float* wmat = nullptr;
  const float* wsum = nullptr;
  // epilogue
  int epi = EPI_STORE, act = 0, rev = 0;
  T* outT = nullptr;
  int outT_pitch = 0, outT_off = 0;
  long long outT_slabM = 0;
  float* outF = nullptr;
  int outF_pitch = 0, outF_off = 0;
  float* z = nullptr;       // latent state [M][52]
  float* sbuf = nullptr;    // coupling log-scale s [M][48]
  T* copyA = nullptr;       // typed copies of the coupling result into conv-input slots
  int copyA_pitch = 0;
  T* copyB = nullptr;
  int copyB_pitch = 0;
  int copy_pad = 0;         // Y1: zero-fill channels [3, copy_pad)
  long long copy_slabM = 0; // layout of the copyA / copyB dense buffers
  const T* resid = nullptr; // EPI_GA residual
  int resid_pitch = 0;
  T* outAct = nullptr;      // EPI_GA: optional LeakyReLU'd copy of the result (input of the GMM head)
  int outAct_pitch = 0;
};

template <typename T>
int launch_conv_simt(const ConvArgs<T>& a, cudaStream_t st);

// reference conv weight [cout][cin_ref][taps...] -> [taps*cin_buf][np]; buffer channel c maps to reference
// channel c (c < xreal), nothing (xreal <= c < xpad), or c - xpad + xreal (growth channels)
int launch_pack_conv_simt(const float* wref, const float* bref, float* wpk, float* bpk, int cout, int cin_ref, int taps,
                          int cin_buf, int xreal, int xpad, int np, cudaStream_t st);

// ---- stp.cu: GlobalAgg statistics, GMM sampler -----------------------------------------------------------------
int launch_ga_wmap(const float* fcw, float* wmap, int h, int w, cudaStream_t st);
template <typename T>
int launch_ga_stat(const T* x, int pitch, const float* wmap, float* partial, int nsplit, int BT, int hw, cudaStream_t st);
// BF16 mode GlobalAgg apply: out = x + sum_t W[b,t,t'] * P[t], P = proj1(x) + bias (pixel-major [M][64] bf16)
int launch_ga_mix(const __nv_bfloat16* P, const __nv_bfloat16* x, const float* wmat, __nv_bfloat16* outT, int outT_pitch,
                  long long outT_slabM, float* outF, int outF_pitch, __nv_bfloat16* outAct, int B, int T, long long hw, cudaStream_t st,
                  bool x2 = false /* BF16X3 mode: P, x, outT, outAct hold (hi, lo) pairs */);
int launch_ga_weights(const float* partial, int nsplit, const float* fcb, const float* p2w, const float* p2b, const float* p3w,
                      const float* p3b, float* wmat, float* wsum, int B, int T, cudaStream_t st);
// params: pixel-major [M][720] fp32 (ppitch floats per pixel) or NCHW [BT,720,h,w] (params_nchw)
int launch_gmm_sample(const float* params, bool params_nchw, const float* eps, uint64_t seed, uint64_t offset, float* v,
                      bool v_nchw, int vpitch, int voff, int B, int T, int h, int w, cudaStream_t st);
// params as planar quads [180][M][4] in the permuted channel order n' = j*240 + k*48 + hf (tcgen05 head); v -> planar z
// params_half: the quads are fp16 (what the tcgen05 head writes by default), read by the warp-split form only
int launch_gmm_sample_planar(const float* params, const float* eps, uint64_t seed, uint64_t offset, float* z, int B, int T, int h,
                             int w, cudaStream_t st, int form = -1 /* 0 thread-per-pixel, 1 warp-split, -1 default / SELFC_GMM_SPLIT */,
                             bool params_half = false);
// by_component: rows k*144 + j*48 + hf (fused head + sampler, one GEMM per mixture component); else j*240 + k*48 + hf
int launch_permute_gmm_rows(const float* w, const float* b, float* wp, float* bp, bool by_component, cudaStream_t st);
int launch_export_eps(float* eps, uint64_t seed, uint64_t offset, int B, int T, long long hw, cudaStream_t st);

}  // namespace selfc
