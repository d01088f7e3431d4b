// Context, workspace layout and a few host helpers shared by net.cu (inference orchestration + C-ABI) and train.cu
// (backward kernels of the training step).
#pragma once
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "conv_tc.h"
#include "pack_batch.h"

namespace selfc {

// ---- parameter index map (module registration order of the reference) ------------------------
constexpr int P_INV0 = 0;           // operations.{1..8}.{F,G,H}.conv{1..5}.{weight,bias}: 8 * 3 * 10
constexpr int P_LOCAL1 = 240;       // stp_net.local_m1 (10)
constexpr int P_LOCAL2 = 250;       // stp_net.local_m2 (10)
constexpr int P_GLOBAL1 = 260;      // stp_net.global_m1 (8: fc, proj1, proj2, proj3)
constexpr int P_GLOBAL2 = 268;
constexpr int P_OTHER = 276;        // 4 x (D2DT 10, GlobalAgg 8)
constexpr int P_TAIL = 348;         // tail_gmm.{1,3,5}.{weight,bias}
static_assert(P_TAIL + 6 == SELFC_NUM_PARAMS, "parameter map");

struct DenseW {            // one D2DTInput in kernel layout
  int cin = 0, cout = 0, xpad = 0;
  float* w[5] = {};        // SIMT fp32 [taps*cin_buf][np]
  float* b[5] = {};
  int np[5] = {};
  TcConvW tc[5];           // tcgen05 bf16 images (BF16 mode)
  TcTempW t5;              // conv5 image (BF16 mode)
  void* f5img = nullptr;   // BF16 mode, cout == 3 (F blocks): conv5's taps as a pointwise GEMM inside the fused dense-block kernel
  // BF16X3 training: input-gradient images, packed on the first backward after every weight load (dg_valid)
  void* dg_img[4] = {};    // slots X, x1, x2, x3 of the buffer: pack_tc3_dgrad_slot_images (4, 3, 2, 1 contributing convs)
  TcTempW dg5[2];          // conv5: the buffer channels in two column groups (<= 96 each)
  int dg5_c0[2] = {}, dg5_n[2] = {};
  float* dg5_tmp = nullptr; // conv5's flipped / transposed weights in reference layout, the source of dg5's images
  bool dg_valid = false;
};
struct GaW {
  float *fcw = nullptr, *fcb = nullptr, *p2w = nullptr, *p2b = nullptr, *p3w = nullptr, *p3b = nullptr;
  float *p1w = nullptr, *p1b = nullptr;   // packed [64][64] (k-major rows), bias [64]
  TcTempW tp;                             // proj1 image (BF16 mode)
  // fc(adaptive_avg_pool2d(.)) as a per-pixel weight map [h*w]: depends only on fc.weight and (h, w), so it is built once per
  // (weights, frame size, stream) and kept by the context (invalidated by selfc_ctx_load_weights, freed by selfc_ctx_destroy)
  float* wmap_cache = nullptr;
  size_t wmap_cap = 0;
  int wmap_h = 0, wmap_w = 0;
  cudaStream_t wmap_stream = nullptr;
};
struct ProfRec {
  cudaEvent_t a = nullptr, b = nullptr;
  int cls = 0;
  double work = 0.0;
};
struct HeadW {
  float* w[3] = {};
  float* b[3] = {};
  int cin[3] = {64, 128, 256}, cout[3] = {128, 256, 720}, np[3] = {128, 256, 736};
  TcTempW t[5];            // BF16 mode: 64->128, 128->256, 256->240 x3 as tcgen05 pointwise GEMMs
  TcTempW g[5];            // BF16 mode: 256->144 per mixture component (fused head + sampler)
  // BF16X3 training: 256 -> 720 in five 144-row groups in REFERENCE channel order (the sampler's backward reads pixel-major [M][720]),
  // and the input-gradient images of the three layers (transposed weights; 256 -> 720 in four 64-column groups)
  TcTempW r[5];
  TcTempW dg3[4], dg2, dg1;
};

}  // namespace selfc

using selfc::DenseW;
using selfc::GaW;
using selfc::HeadW;
using selfc::ProfRec;

struct selfc_ctx {
  int device = 0, mode = SELFC_MODE_FP32;
  bool loaded = false;
  int xpad3 = 4;            // X-slot width of the 3-channel dense blocks (G, H, local_m1)
  DenseW inv[8][3];         // [block][F,G,H]
  DenseW stp[6];            // local_m1, local_m2, other 0,2,4,6
  GaW ga[6];
  HeadW head;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  std::mutex mu;
  // training-step scratch owned by the context (dgrad weights + weight-gradient scratch, a zero bias vector): allocated on first
  // use under `mu`, freed by selfc_ctx_destroy -- two contexts (or two trainers) on one device do not share it
  float* train_scratch = nullptr;
  float* train_zero_bias = nullptr;
  // BF16X3 training: zero-padded pixel planes of the tensor-core weight-gradient kernel (wgrad_tc.cu), zeroed whenever the clip
  // geometry (B, T, h, w) changes; same ownership as the scratch above
  void* wg_planes = nullptr;
  size_t wg_bytes = 0;
  int wg_key[4] = {0, 0, 0, 0};
  void* dg_gslab = nullptr;      // the output gradient of one conv as (hi, lo) slabs (input of a tensor-core input-gradient launch)
  size_t dg_gslab_bytes = 0;
  selfc::PackBatch pack;         // job tables of selfc_ctx_load_weights' batched weight packing (pack_batch.h)
  selfc::PackBatch pack_dg;      // ... and of the training step's input-gradient images (train.cu)
  // optional per-launch timing (bench.py's roofline leg): CUDA events around every launch, by kernel class
  bool prof_on = false;
  std::vector<ProfRec> prof;
};

namespace selfc {

// ---- workspace layout -------------------------------------------------------------------------------------
struct Workspace {
  size_t total = 0;
  size_t z, sbuf, fbuf, gbuf, hbuf, stpbuf, feat, fact, h1, h2, params, wmap, partial, wmat, wsum, lrq;
  int nsplit = 1;
  int fpitch = 176, gpitch = 0, spitch = 192;
};

Workspace make_workspace(const selfc_ctx* ctx, int B, int T, int h, int w);

struct Dims {
  int B, T, h, w;
  long long M() const { return (long long)B * T * h * w; }
  long long hw() const { return (long long)h * w; }
};

// layout of the dense-block buffers (common.cuh): slab-planar in the BF16 / BF16X3 modes, pixel-major in FP32 mode
inline long long dense_slab(const selfc_ctx* ctx, const Dims& d) { return ctx->mode != SELFC_MODE_FP32 ? d.M() : 0; }


// what the training step asks the reverse pass to keep (fp32, device): ga_save = 7 x [M][64] (slot i+1 <- output of
// GlobalAgg i), z_save = 8 x planar state (slot blk <- the state reverse block blk starts from)
// Dense buffers of ONE coupling block kept for its backward (training): its own F / G / H buffers and log-scale, and where the block's
// last epilogue puts the NEXT block's input (forward direction: y2 -> the next block's F; reverse: y1 -> the next block's G and H).
// Inference runs every block in the same three workspace buffers (all of these equal the workspace's).
struct BlockBufsV {
  void *f = nullptr, *g = nullptr, *h = nullptr, *f_next = nullptr, *g_next = nullptr, *h_next = nullptr;
  float* s = nullptr;
};
struct TrainHooks {
  float* ga_save = nullptr;
  float* z_save = nullptr;              // slot blk <- the state reverse block blk starts from; slot 8 <- the state after the last one
  const BlockBufsV* up_bufs = nullptr;  // [8]: per-block buffers of the reverse pass (the activations stay for the backward)
};
// forward pieces re-used by the training step (net.cu); E = float (FP32 mode) or bfx2 (BF16X3 mode: the same launches on the tcgen05
// kernels, dense buffers slab-planar (hi, lo) pairs)
template <typename E>
int up_hooked(selfc_ctx* ctx, const float* lr, const float* eps, uint64_t seed, uint64_t offset, float* hr, const Dims& d, char* wsp,
              const Workspace& ws, cudaStream_t st, const TrainHooks* hooks);
// one STP stage's dense block: conv1..4 + conv5 on the buffer whose X slot is already filled -> feat [M][64] fp32
template <typename E>
int stp_dense(const selfc_ctx* ctx, int i, E* stpbuf, int pitch, float* feat, const Dims& d, cudaStream_t st);
template <typename E>
int dense_convs(const selfc_ctx* ctx, const DenseW& W, E* buf, int pitch, const Dims& d, cudaStream_t st);
// InvBlockExp forward / reverse on the latent state in the workspace (ws.z), leaving the F / G / H dense buffers and the
// log-scale (ws.sbuf) behind; the X slot of the first dense block must already hold its input (x2 for F when !rev, x1 for
// G and H when rev)
template <typename E>
int invblock_fwd(const selfc_ctx* ctx, int blk, bool rev, char* wsp, const Workspace& ws, const Dims& d, cudaStream_t st,
                 const BlockBufsV* bufs = nullptr);
int check_run(selfc_ctx* ctx, int B, int T, int H, int W, void* workspace, size_t workspace_bytes, Workspace* ws);
const DenseW* find_dense(selfc_ctx* ctx, int first_param);
const GaW* find_ga(selfc_ctx* ctx, int first_param);

}  // namespace selfc
