// tcgen05 / TMEM / TMA implicit-GEMM convolution (BF16 mode) -- host-side interface.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace selfc {

// One (1,3,3) convolution's weights as the exact shared-memory image the UMMA B-operand descriptors read
// (see conv_tc3.cu), plus its fp32 bias.
struct TcConvW {
  void* img = nullptr;      // device, bf16
  void* img_pair = nullptr; // device, bf16: the same rows split in two halves for the CTA-pair kernel
  void* img_x2 = nullptr;   // device, bf16 (BF16X3 mode): the pair image with a hi and a lo tile per (ky, K-step), 2 x img_bytes
  float* bias = nullptr;    // device, [32]
  size_t img_bytes = 0;
  int cin_buf = 0;          // input channels consumed from the dense buffer (multiple of 16)
};

int pack_tc_weights(TcConvW& w, const float* wref, const float* bref, int cin_ref, int cin_buf, int xreal, int xpad, cudaStream_t st,
                    bool x2 = false);
void free_tc_weights(TcConvW& w);
// conv_k of a dense block, in place on the slab-planar buffer (slabM = N*h*wd pixels per 16-channel slab): reads
// channels [0,cin), writes lrelu(conv+bias) to [out_off,out_off+32)   (conv_tc3.cu)
// (w2, buf2): optional second problem of the same shape run by the other half of the grid in the same launch (G and H).
// The buffers must have 32 readable bytes before their first and after their last slab (the position-pair tensor map reads one
// position either side; the values never reach an output).  Workspace regions satisfy this: none is first, all carry a guard.
// x2 (BF16X3 mode, common.cuh): the buffers hold (hi, lo) bf16 pairs -- 64-byte rows per position and slab -- and every tap is three MMAs
// acc (x2 only; training): instead of bias + LeakyReLU + a (hi, lo) store, ADD the 32 results to channels [off, off + n) of an fp32
// pixel-major buffer -- an input-gradient launch: buf holds the concatenated output gradients, w.img_x2 pack_tc3_dgrad_slot_images' images
struct TcAccum {
  float* out = nullptr;
  int pitch = 0, off = 0, n = 0;      // n channels in ngroups groups of 32 (w.img_x2 = the first of ngroups consecutive images)
  int ngroups = 1;
  long long slabM = 0;                // layout of `out`: pixel-major [M][pitch], or 16-channel fp32 slabs [pitch / 16][slabM][16]
};
int launch_conv3x3_tc(const TcConvW& w, __nv_bfloat16* buf, long long slabM, int cin, int out_off, int N, int h, int wd, cudaStream_t st,
                      const TcConvW* w2 = nullptr, __nv_bfloat16* buf2 = nullptr, bool x2 = false, const TcAccum* acc = nullptr);
// input-gradient images of one 32-channel-group SLOT [c0, c0 + ncover) of a dense buffer: the later convs' transposed, tap-flipped
// weights stacked in K in the order [conv4 | conv3 | ...] (nconv of them; wf / cin = the forward packs [9][cin_k][32] of conv1..4);
// ceil(ncover / 32) images of tc3_dgrad_slot_image_bytes(nconv) each
size_t tc3_dgrad_slot_image_bytes(int nconv);
int pack_tc3_dgrad_slot_images(const float* const wf[4], const int cin[4], void* img, int c0, int ncover, int nconv, cudaStream_t st);

// ---- dense_fused.cu: conv1..convL of a dense block in ONE launch, growth channels kept in tensor memory ----------------
// w[0..L-1] = the block's conv_k weights (TcConvW::img_pair is what the kernel reads), L = dense_fused_layers(cin): 4, or 3 when the fourth layer's
// weights do not fit shared memory next to the X ring (cin = 64).  Reads channels [0,cin) of the slab-planar buffer, writes
// x1..xL to [cin, cin + 32 L); bit-identical to L launches of launch_conv3x3_tc.  (w2, buf2): second problem of the same shape.
int dense_fused_schedule(int sch, int* out18);       // the compiled schedule tables (C-ABI selfc_dense_fused_schedule)
int dense_fused_layers(int cin);      // how many layers (4, 3 or 0 = unsupported) one launch fuses for this X width
// (f5img, f5part): for a block with 3 outputs (F of a coupling; cin 48) the launch also applies conv5's three temporal taps to
// every row while [X | x1..x4] is on chip and writes 9 fp32 partial products per pixel to f5part [3 taps][M][4]; x1..x4 are then
// NOT written to HBM and launch_f5_combine finishes conv5 + the additive coupling (replaces the temporal kernel's EPI_COUPLE_Y1).
int launch_dense_fused(const TcConvW* w, int L, __nv_bfloat16* buf, long long slabM, int cin, int N, int h, int wd, cudaStream_t st,
                       const TcConvW* w2 = nullptr, __nv_bfloat16* buf2 = nullptr, const void* f5img = nullptr, float* f5part = nullptr);
int pack_f5_weights(void** img, const float* wref, int cin, cudaStream_t st);      // conv5.weight [3][cin][3] -> *img (allocated on first use)
int launch_f5_combine(const float* part, const float* bias, float* z, __nv_bfloat16* copyA, __nv_bfloat16* copyB, int T, long long hw,
                      long long M, int rev, cudaStream_t st);

// ---- temporal_tc.cu: (3,1,1) conv5 + coupling epilogues, GlobalAgg apply ---------------------------------------
struct TcTempW {
  void* img = nullptr;      // device, bf16 B-operand image [tap][kstep][...]
  float* bias = nullptr;    // device, [64]
  size_t img_bytes = 0;
  int cin_buf = 0, npad = 0, taps = 0, cout = 0;
  bool x2 = false;          // BF16X3 mode: a hi and a lo tile per (tap, K-step); the launch then reads (hi, lo) activations
};

struct TcTempArgs {
  const __nv_bfloat16* in = nullptr;
  int in_pitch = 0, B = 0, T = 0, hw = 0;
  long long in_slabM = 0;            // != 0: `in` is a slab-planar dense buffer (common.cuh) of in_slabM pixels per slab
  // EPI_COUPLE_HG: H(y1) and G(y1) of a coupling in ONE launch -- `in` is H's dense buffer, `in2` G's (same shape); the
  // accumulator holds [H | G] (48 + 48 columns), s = 2*sigmoid(H)-1 never leaves the SM
  const __nv_bfloat16* in2 = nullptr;
  int epi = 0, rev = 0, act = 0;
  __nv_bfloat16* outT = nullptr;
  int outT_pitch = 0, outT_off = 0;
  long long outT_slabM = 0, copy_slabM = 0;
  float* outF = nullptr;
  int outF_pitch = 0, outF_off = 0, outF_planar = 0;
  int outF_blk = 0, outF_blk_stride = 0;   // planar outF: column n -> channel outF_off + (n / blk) * stride + n % blk
  long long outF_slabM = 0;                // non-planar outF: 16-channel fp32 slabs instead of pixel-major rows
  long long m_limit = 0;             // > 0: rows >= m_limit do not exist (pointwise mode over pseudo-frames)
  float* z = nullptr;
  float* sbuf = nullptr;
  __nv_bfloat16* copyA = nullptr;
  int copyA_pitch = 0;
  __nv_bfloat16* copyB = nullptr;
  int copyB_pitch = 0;
  int copy_pad = 0;
  const float* wmat = nullptr;
  const float* wsum = nullptr;
  const __nv_bfloat16* resid = nullptr;
  int resid_pitch = 0;
  void* dbg = nullptr;               // optional device buffer of 16 x int64: per-role barrier wait cycles of CTA 0
  __nv_bfloat16* outAct = nullptr;   // EPI_GA: optional LeakyReLU'd copy (input of the GMM head)
  int outAct_pitch = 0;
  // EPI_GMM (fused GMM head + sampler): this launch's accumulator columns are [logit | log-scale | mean] x 48 hf of mixture
  // component gmm_k; v[hf] (+)= softmax_hf(logit) * (eps * exp(clamp(ls)) + mean) goes to quads 1..12 of z (stored for
  // gmm_k == 0, accumulated otherwise).  eps: injected [B,48,5,T,h,w] or null -> Philox stream (seed, offset).
  int gmm_k = 0, gmm_T = 1;
  long long gmm_hw = 0;
  const float* eps = nullptr;
  uint64_t seed = 0, offset = 0;
};

int pack_temporal_weights(TcTempW& w, const float* wref, const float* bref, int cout, int cin_ref, int taps, int cin_buf, int xreal,
                          int xpad, cudaStream_t st, bool x2 = false);
void free_temporal_weights(TcTempW& w);
bool temporal_tc_supported(const TcTempW& w, int T);
// w2: weights of the second conv of an EPI_COUPLE_HG launch (G), same shape as w (H)
int launch_temporal_tc(const TcTempW& w, const TcTempArgs& a, cudaStream_t st, const TcTempW* w2 = nullptr);

}  // namespace selfc
