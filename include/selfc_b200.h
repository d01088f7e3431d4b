/* selfc_b200 C-ABI -- the drop-in boundary of the SelfC-large 4x rescaling hot path on B200 (sm_100a).
 *
 * Plain C symbols, device pointers and sizes only: no torch / C++ types cross this boundary.
 * The reference (tianyuan168326/SelfC) has no native code; each entry point below replaces a
 * stretch of stock-PyTorch ops in the reference's Python.  Citations are relative to
 * /root/reference/codes.  The host-side mirror of the reference's operator interface
 * (selfc_b200/arch.py: SelfCInvNet.forward(x, rev)) binds these with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative SELFC_E_* code; selfc_last_error() gives the
 *     text for the calling thread.  Nothing throws, nothing synchronises the device.
 *   - all pointers are DEVICE pointers (16-byte aligned) unless the name says host; `stream` is a
 *     cudaStream_t passed as void* (the caller's current torch stream).
 *   - frames are the reference's layouts: [B*T, C, H, W] fp32 NCHW; H, W multiples of 4; h=H/4, w=W/4.
 *   - the library allocates nothing persistent except the packed weight cache inside a selfc_ctx; all
 *     activations live in a caller-provided workspace (selfc_workspace_bytes).
 */
#ifndef SELFC_B200_H
#define SELFC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SELFC_OK 0
#define SELFC_E_ARG (-1)       /* bad shape / null pointer / misaligned pointer */
#define SELFC_E_STATE (-2)     /* weights not loaded, wrong device, workspace too small */
#define SELFC_E_CUDA (-3)      /* a CUDA runtime / driver call or a launch failed */
#define SELFC_E_UNSUPPORTED (-4)

/* precision modes (new knob, supplied out of band; SURVEY 8b "Config") */
#define SELFC_MODE_FP32 0      /* fp32 storage, fp32 FMA convolutions: the <=1e-3 numerics gate */
#define SELFC_MODE_BF16 1      /* bf16 activations, tcgen05 kind::f16 implicit-GEMM, fp32 accumulate/state */
#define SELFC_MODE_BF16X3 2    /* the numerics-gate mode ON TENSOR CORES (BASELINE configs[1]): activations and weights as (hi, lo) bf16
                                  pairs (16 mantissa bits), every product as three tcgen05 MMAs (hi.hi + hi.lo + lo.hi) with fp32
                                  accumulation, fp32 latent state -- HR within 1e-3 of the fp32 reference at ~1/3 of the BF16 rate */

#define SELFC_NUM_PARAMS 354   /* state_dict tensors of SelfCInvNet (vid4 YAML) */

typedef struct selfc_ctx selfc_ctx;

int selfc_version(void);
const char* selfc_last_error(void);

/* ---- context: packed-weight cache for one device --------------------------------------------------- */
int selfc_ctx_create(selfc_ctx** out, int device, int mode);
int selfc_ctx_destroy(selfc_ctx* ctx);
int selfc_ctx_mode(const selfc_ctx* ctx);
/* Re-pack the reference-layout parameters (models/modules/SelfC_GMM_arch_inv.py:433-448 registration order,
 * SURVEY A.8; `params[i]` = device pointer to the i-th fp32 tensor) into kernel layout.  Called after
 * load_state_dict / .to(); replaces base_model.py:87-107's implicit "weights are just nn.Parameters". */
int selfc_ctx_load_weights(selfc_ctx* ctx, const float* const* params_host_array_of_dev_ptrs, int n_params, void* stream);

size_t selfc_workspace_bytes(const selfc_ctx* ctx, int B, int T, int h, int w);

/* ---- the path: SelfCInvNet.forward (SelfC_GMM_arch_inv.py:450-490) ---------------------------------- */
/* rev=False (:454-469): FrequencyAnalyzer + 8 InvBlockExp.  out51 [B*T,51,h,w] fp32 (may be NULL);
 * lr_u8 [B*T,3,h,w] = Quantization of out51[:, :3] (Quantization.py:4-17, models/SelfC_model.py:217-222)
 * as 8-bit codes, lr_q the same on the 1/255 fp32 grid (either may be NULL). */
int selfc_down(selfc_ctx* ctx, const float* hr, float* out51, uint8_t* lr_u8, float* lr_q,
               int B, int T, int H, int W, void* workspace, size_t workspace_bytes, void* stream);
/* rev=True (:470-490): STPNet prior + soft-GMM sample + 8 inverse couplings + FrequencyAnalyzer reverse.
 * lr [B*T,3,h,w] fp32.  eps: NULL -> counter-based Philox4x32-10 noise keyed (seed, offset) and indexed by
 * the reference's eps linear index in [B,48,5,T,h,w] (:412-415); non-NULL -> injected noise in that layout.
 * hr [B*T,3,H,W]; hf [B*T,48,h,w] (recon_hf, may be NULL). */
int selfc_up(selfc_ctx* ctx, const float* lr, const float* eps, uint64_t seed, uint64_t offset,
             float* hr, float* hf, int B, int T, int H, int W,
             void* workspace, size_t workspace_bytes, void* stream);

/* ---- 8-bit frames at the boundary --------------------------------------------------------
 * Images are what cv2.imread / cv2.imwrite hold: [n][H][W][3] bytes, channel order B,G,R.  Ingest replaces
 * read_img1's astype(float32)/255 (data/util.py:103-115) + the dataset's BGR->RGB, HWC->CHW
 * (data/LQGTVID_dataset.py:150-154); egress replaces tensor2img's clamp(0,1), *255, round-half-even, RGB->BGR,
 * CHW->HWC (utils/util.py:104-133) before save_img (:181-182).  Both are bit-exact with those CPU conversions and,
 * on the HR side, fused into the FrequencyAnalyzer kernels (no fp32 frame ever exists in HBM). */
/* hr_img [B*T][H][W][3] -> lr_img [B*T][h][w][3] (may be NULL), lr_q [B*T,3,h,w] fp32 on the 1/255 grid (may be NULL) */
int selfc_down_u8(selfc_ctx* ctx, const uint8_t* hr_img, uint8_t* lr_img, float* lr_q, int B, int T, int H, int W,
                  void* workspace, size_t workspace_bytes, void* stream);
/* lr_img [B*T][h][w][3] (the stored LR video) -> hr_img [B*T][H][W][3]; eps/seed/offset as selfc_up */
int selfc_up_u8(selfc_ctx* ctx, const uint8_t* lr_img, const float* eps, uint64_t seed, uint64_t offset, uint8_t* hr_img,
                int B, int T, int H, int W, void* workspace, size_t workspace_bytes, void* stream);
/* both halves in one call (models/SelfC_model.py:213-233 on 8-bit frames); lr_img may be NULL */
int selfc_rescale_u8(selfc_ctx* ctx, const uint8_t* hr_img, const float* eps, uint64_t seed, uint64_t offset,
                     uint8_t* lr_img, uint8_t* hr_out_img, int B, int T, int H, int W,
                     void* workspace, size_t workspace_bytes, void* stream);
/* stand-alone conversions, any H and W: img [N][H][W][3] BGR bytes <-> x [N,3,H,W] RGB fp32 */
int selfc_frames_from_u8(const uint8_t* img, float* x, int N, int H, int W, void* stream);
int selfc_frames_to_u8(const float* x, uint8_t* img, int N, int H, int W, void* stream);

/* ---- components (each is one row of SURVEY 8a; used by the parity tests) ---------------------------- */
/* a1 FrequencyAnalyzer.forward(rev=False) :62-78 -> [N,51,h,w]; a10 rev=True :79-82 -> [N,3,H,W] */
int selfc_fa_fwd(const float* x, float* out51, int N, int H, int W, void* stream);
int selfc_fa_rev(const float* z51, float* y, int N, int h, int w, void* stream);
/* f1: validation metrics (train.py:28-86 cal_metric).  rgb_to_ycbcr (data/util.py:239-245): x [N,3,H,W] -> y [N,1,H,W].
 * frame_metrics: per frame n, sse[n] = sum (a-b)^2 and ssim_sum[n] = sum of the SSIM map (utils/util.py:361-488: 11-tap
 * window win11, valid convolution, K1 .01, K2 .03, data_range 1) over C channels, or over the BT.601 luma when to_y != 0
 * (C must be 3).  Outputs are fp64 device arrays of N; ssim_sum may be NULL.  PSNR = 20 log10(1/sqrt(sse / count)),
 * SSIM = ssim_sum / (Ceff (H-10) (W-10)) are left to the caller (utils/util.py:198-221, :596-603). */
int selfc_rgb_to_y(const float* x, float* y, int N, int H, int W, void* stream);
int selfc_frame_metrics(const float* a, const float* b, int N, int C, int H, int W, int to_y, const float* win11,
                        double* sse, double* ssim_sum, void* stream);
/* f3: the 2x operators of the sibling configurations.  FrequencyAnalyzer(k=2) of the compression model's rescaler half
 * (SelfC_Codec_arch_inv.py:78-98): x [N,3,H,W] <-> [N,15,H/2,W/2].  HaarDownsampling of `model: SelfC` / IRN
 * (SelfC_arch_inv.py:44-84, Inv_arch.py:44-84): x [N,C,H,W] <-> [N,4C,H/2,W/2], output channel k*C+c. */
int selfc_fa2_fwd(const float* x, float* out15, int N, int H, int W, void* stream);
int selfc_fa2_rev(const float* z15, float* y, int N, int h, int w, void* stream);
int selfc_haar_fwd(const float* x, float* out, int N, int C, int H, int W, void* stream);
int selfc_haar_rev(const float* z, float* y, int N, int C, int h, int w, void* stream);
/* a4 Quantization.py:4-17 on [n] floats */
int selfc_quantize(const float* x, uint8_t* q_u8, float* q_f32, size_t n, void* stream);
/* caller-side neighbour of the path: LR_ref of `distortion: sr_bd`, models/Guassian.py:7-52 as called at
 * models/SelfC_model.py:128-129.  x [N,C,H,W] -> y [N,C,H/4,W/4]; k13 = the 13x13 taps (device, 169 floats). */
int selfc_gaussian_down(const float* x, const float* k13, float* y, int N, int C, int H, int W, void* stream);
/* a3 D2DTInput.forward (Subnet_constructor.py:115-133) for the dense block stored at parameter index
 * `first_param` (its conv1.weight); x [B*T,Cin,h,w] -> y [B*T,Cout,h,w], both fp32 NCHW. */
int selfc_d2dt(selfc_ctx* ctx, int first_param, const float* x, float* y, int B, int T, int h, int w,
               void* workspace, size_t workspace_bytes, void* stream);
/* one (1,3,3) conv of that dense block (Subnet_constructor.py:102-105,126-129), k = 0..3 for conv1..conv4:
 * x [B*T, Cin+32k, h, w] (the concatenated input) -> y [B*T,32,h,w] = LeakyReLU_0.2(conv(x) + bias) */
int selfc_conv3x3(selfc_ctx* ctx, int first_param, int k, const float* x, float* y, int B, int T, int h, int w,
                  void* workspace, size_t workspace_bytes, void* stream);
/* a13 (training step, models/SelfC_model.py:148-183) building block: backward of D2DTInput (Subnet_constructor.py:115-133)
 * for the dense block at `first_param`.  x [B*T,Cin,h,w], gy [B*T,Cout,h,w] -> gx [B*T,Cin,h,w]; gparams[10] (conv1.weight,
 * conv1.bias, ..., conv5.bias in the reference layouts, device fp32) are ACCUMULATED into; a NULL weight entry skips that
 * conv's weight and bias gradient.  FP32 or BF16X3 mode (BF16X3: forward, input and weight gradients on the tcgen05 kernels).  The forward activations are recomputed here. */
int selfc_d2dt_backward(selfc_ctx* ctx, int first_param, const float* x, const float* gy, float* gx, float* const* gparams,
                        int B, int T, int h, int w, void* workspace, size_t workspace_bytes, void* stream);
/* a13 building block: backward of InvBlockExp (SelfC_GMM_arch_inv.py:21-33) number blk (0..7) in the forward (rev = 0) or
 * reverse (rev = 1) direction.  z_in [B*T,51,h,w]: the block's input (its forward is recomputed); gz [B*T,51,h,w]: gradient
 * w.r.t. the block's output on entry, overwritten with the gradient w.r.t. its input; gparams[30] (F, G, H x conv1..5 weight,
 * bias; reference layouts) are accumulated into.  FP32 or BF16X3 mode. */
int selfc_invblock_backward(selfc_ctx* ctx, int blk, int rev, const float* z_in, float* gz, float* const* gparams,
                            int B, int T, int h, int w, void* workspace, size_t workspace_bytes, void* stream);
/* extra device memory ("tape": gradient state, saved block inputs, scratch) the training-step entry points need */
size_t selfc_train_tape_bytes(int B, int T, int h, int w);
/* a13 building block: backward of the GMM head (tail_gmm, SelfC_GMM_arch_inv.py:336-344) + soft-GMM sampler (:383-394).
 * feat [B*T,64,h,w]: the STP feature; gv [B*T,48,h,w]: gradient w.r.t. the sampled HF latents; eps as in selfc_up (NULL: the
 * Philox stream seed/offset); gfeat [B*T,64,h,w] out; gparams[6]: tail_gmm.{1,3,5}.{weight,bias}, accumulated into. */
int selfc_head_sampler_backward(selfc_ctx* ctx, const float* feat, const float* gv, const float* eps, uint64_t seed,
                                uint64_t offset, float* gfeat, float* const* gparams, int B, int T, int h, int w,
                                void* workspace, size_t workspace_bytes, void* tape, size_t tape_bytes, void* stream);
/* a13 building block: backward of GlobalAgg (SelfC_GMM_arch_inv.py:265-285) whose fc.weight is parameter `first_param`.
 * x, gout [B*T,64,h,w] -> gx; gparams[8] (fc.weight, fc.bias, proj1.weight, proj1.bias, proj2.weight, proj2.bias,
 * proj3.weight, proj3.bias; reference layouts) are accumulated into.  FP32 or BF16X3 mode, T <= 16. */
int selfc_global_agg_backward(selfc_ctx* ctx, int first_param, const float* x, const float* gout, float* gx,
                              float* const* gparams, int B, int T, int h, int w, void* workspace, size_t workspace_bytes,
                              void* tape, size_t tape_bytes, void* stream);
/* a13 SelfCModel.optimize_parameters' forward + backward (models/SelfC_model.py:148-170, models/modules/loss.py:5-21,
 * Quantization.py:4-17 with its straight-through gradient; losses of train_rescaling_selfc_large.yml: l2 forward fit against
 * ref_l, Charbonnier reconstruction, total x 144*144*3).  hr [B*T,3,H,W], ref_l [B*T,3,H/4,W/4], eps as in selfc_up.
 * grads[354]: device fp32 buffers in the reference parameter layouts (state_dict order), ACCUMULATED into -- zero them for a
 * plain step; losses: 3 device floats (total, l_forw_fit, l_back_rec).  Block inputs are kept on the tape and every block's
 * coupling block keeps its activations inside the tape for its backward (the STP blocks are recomputed).  FP32 mode (fp32-FMA kernels) or
 * BF16X3 mode (every convolution pass on the tcgen05 kernels with (hi, lo) bf16 operands; fp32 gradients), T <= 16. */
int selfc_train_grads(selfc_ctx* ctx, const float* hr, const float* ref_l, const float* eps, uint64_t seed, uint64_t offset,
                      float* const* grads, int n_grads, float* losses, int B, int T, int H, int W, void* workspace,
                      size_t workspace_bytes, void* tape, size_t tape_bytes, void* stream);
/* a13 optimiser step (models/SelfC_model.py:66-68 Adam, :172-176 clip_grad_norm_ + step) on a flat gradient buffer.
 * params[n_tensors]: DEVICE array of device pointers to the parameter tensors; offsets[n_tensors+1]: DEVICE array of the
 * tensors' element offsets inside grad / m / v (flat fp32, `total` elements); gscale multiplies the gradient first (1/world
 * after a SUM all-reduce); max_norm <= 0 disables clipping; step >= 1 (bias correction); sqnorm_scratch: one device float. */
int selfc_adam_step(float* const* params, const long long* offsets, int n_tensors, long long total, const float* grad,
                    float* m, float* v, float* sqnorm_scratch, float gscale, float max_norm, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, void* stream);
/* a6 GlobalAgg.forward (:265-285) for the module whose fc.weight is parameter `first_param`;
 * x,y [B*T,64,h,w]; wmat_out (may be NULL) receives the [B,T,T] mixing matrix. */
int selfc_global_agg(selfc_ctx* ctx, int first_param, const float* x, float* y, float* wmat_out,
                     int B, int T, int h, int w, void* workspace, size_t workspace_bytes, void* stream);
/* a7 sampler (:383-394): params [B*T,720,h,w] (channel hf*15+k*3+j), eps as in selfc_up -> v [B*T,48,h,w] */
int selfc_gmm_sample(const float* params, const float* eps, uint64_t seed, uint64_t offset, float* v,
                     int B, int T, int h, int w, void* stream);
/* the same draw on the bf16 mode's internal layout (the sampler selfc_up launches after the tcgen05 head): params_planar
 * [180][M][4] fp32, quad j*60 + k*12 + i = channels hf 4i..4i+3 of kind j (0 logit, 1 log-sigma, 2 mu), component k, with
 * M = B*T*h*w pixels m = (b*T + t)*h*w + pix; z_planar [13][M][4]: quads 1..12 receive the 48 HF channels (quad 0, the LR
 * frame, is not touched).  form: 0 thread-per-pixel kernel, 1 warp-split kernel, -1 the default (SELFC_GMM_SPLIT). */
int selfc_gmm_sample_planar(const float* params_planar, const float* eps, uint64_t seed, uint64_t offset, float* z_planar,
                            int B, int T, int h, int w, int form, void* stream);
/* the noise selfc_up would draw for (seed, offset), in the reference layout [B,48,5,T,h,w] */
int selfc_export_eps(float* eps, uint64_t seed, uint64_t offset, int B, int T, int h, int w, void* stream);

/* Optional per-launch timing for the roofline report: while enabled, every launch of selfc_down / selfc_up is
 * bracketed by CUDA events on the caller's stream.  Classes: 0 (1,3,3) dense-block convs [work = FLOPs],
 * 1 (3,1,1) conv5 + coupling [FLOPs], 2 GlobalAgg [bytes], 3 GMM head [FLOPs], 4 sampler [bytes], 5 layout [bytes].
 * selfc_prof_read synchronises on the recorded events, sums ms / algorithmic work / launches per class, and clears. */
#define SELFC_PROF_CLASSES 6
int selfc_prof_enable(selfc_ctx* ctx, int on);
int selfc_prof_read(selfc_ctx* ctx, int ncls, double* ms, double* work, uint64_t* launches);

/* The row schedule of dense_fused_kernel (csrc/dense_fused.cu) for schedule id `sch` (0: one-slab X, 1: F, 2: 64->64 STP, 3: F with
 * conv5's taps), for tests that verify the ring sizes by simulating the issue order (host only, no device needed):
 * out[0] = conv layers L, out[1] = groups per step, out[2] = X-ring slots, out[3..7] = lag of group j (rows behind conv1),
 * out[8..12] = issue order, out[13..16] = tensor-memory ring rows of x1..x4 (0 = not kept on chip), out[17] = tensor-memory
 * columns in use.  Returns 0, or SELFC_E_ARG for an unknown id. */
int selfc_dense_fused_schedule(int sch, int* out18);
/* Host-side view of the zero-padded pixel planes the tensor-core weight-gradient kernel (BF16X3 training, csrc/wgrad_tc.cu)
 * contracts over: out4 = {Wp (row pitch, a multiple of 8 >= w + 2), Fp = (h + 2) * Wp, P (plane length), Pa (allocated, multiple of 32)}.
 * Pixel (b, t, y, x) of a [B][T][h][w] clip batch lives at P = ((b * (T + 1) + t + 1) * (h + 2) + y + 1) * Wp + x + 1; a spatial tap
 * (ky, kx) is the offset (ky - 1) * Wp + (kx - 1), a temporal tap dt the offset (dt - 1) * Fp, and every such neighbour of a real pixel
 * is either the real neighbour or padding (tests/test_boundary_cpu.py checks exactly that).  No device work. */
int selfc_wgrad_geometry(int B, int T, int h, int w, long long* out4);

/* number of kernels this library has launched on the calling thread since load (bench.py's gpu_launches) */
uint64_t selfc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SELFC_B200_H */
