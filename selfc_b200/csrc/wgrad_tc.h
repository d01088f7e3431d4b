// Tensor-core weight gradients of the dense-block convolutions (wgrad_tc.cu) -- host-side interface (BF16X3 training mode).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "net_ctx.h"

namespace selfc {

constexpr int kWgRows = 257;       // AT rows: the ones row + up to 256 channels (dense buffers: 192; the GMM head's 256-channel layer)
constexpr int kWgWindowDefault = 0;  // default of SELFC_WGRAD_WINDOW (1 once measured faster on the box)
constexpr int kWgWindowMaxR = 4;    // row pitches up to 4 x 32 pixels take the sliding-window form of the spatial kernel
constexpr int kWgSpatialRows = 96;  // GT rows [0, 96): three one-pixel-shifted copies of a spatial conv's 32 gradient channels
constexpr int kWgGradRows = 160;    // ... rows [96, 160): conv5's (unshifted) gradient, up to 64 channels

// zero-padded pixel planes of one clip batch: P = ((f' * (h + 2) + y + 1) * Wp + x + 1), f' = b * (T + 1) + t + 1
struct WgGeom {
  int B, T, h, w, Wp;
  long long Fp, P, Pa;             // padded frame size, plane length, allocated plane length (multiple of 32)
};
WgGeom wg_geometry(const Dims& d);
size_t wg_plane_bytes(const WgGeom& g);      // [2][257][Pa] + [2][160][Pa] bf16; must be ZEROED once per geometry (the padding)

// The backward of a dense block is a chain of ~25 short dependent kernels (mask, gradient planes, weight gradient, unpack, input
// gradient, five times).  Kernels of the chain start with chain_entry() (common.cuh) and are launched through launch_chain(): the next
// kernel is scheduled while this one drains (programmatic dependent launch) and waits for its results before touching global memory.
// SELFC_TRAIN_PDL=0: plain stream order (A/B).
bool train_pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = train_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// activations of a whole dense buffer (slab-planar (hi, lo) pairs, `pitch` channels) -> AT planes
int launch_wg_planes_act(const bfx2* buf, int pitch, const Dims& d, const WgGeom& g, void* planes, cudaStream_t st);
// output gradient channels [off, off + ncols) of an fp32 buffer (pixel-major, or 16-channel slabs when sslabM != 0) -> GT rows, rows >= ncols zero
// (zero / zero_n: an fp32 range the kernel clears on the way -- the weight-gradient scratch of the launch_wgrad_tc that follows)
int launch_wg_planes_grad(const float* gsrc, int pitch, int off, long long sslabM, int ncols, int nb, bool temporal, const Dims& d,
                          const WgGeom& g, void* planes, cudaStream_t st, float* zero = nullptr, long long zero_n = 0);
// the same from an fp32 pixel-major tensor [M][pitch], channels [0, C) (the pointwise convs of the GMM head and of GlobalAgg)
int launch_wg_planes_act_f32(const float* src, int pitch, int C, const Dims& d, const WgGeom& g, void* planes, cudaStream_t st);
// dw[(tap * cin + c) * np + n_off + n] += sum_p in[p + shift(tap)][c] * g[p][n]; dw[taps * cin * np + n_off + n] += sum_p g[p][n]   (dw zeroed
// by the caller, or by the launch_wg_planes_grad before it).  kind: WG_SPATIAL (9 taps, 32 outputs), WG_TEMPORAL (3 taps, <= 64 outputs), WG_POINT (1 tap, <= 64 outputs per launch:
// a wider layer is a loop over n_off; the gradient planes are built with temporal = true)
enum WgKind { WG_SPATIAL = 0, WG_TEMPORAL = 1, WG_POINT = 2 };
int launch_wgrad_tc(void* planes, const WgGeom& g, int cin, int ncols, int nb, int kind, float* dw, int np, cudaStream_t st, int n_off = 0);

}  // namespace selfc
