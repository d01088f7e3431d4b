// (1,3,3) dense-block convolution as a tcgen05 / TMEM / TMA implicit GEMM (BF16 mode).
//
// What it computes (Subnet_constructor.py:102-105,126-129): for conv_k of a D2DTInput block stored in a pixel-major
// bf16 dense buffer [M][pitch],  out[:, cin:cin+32] = lrelu_0.2( conv3x3(buf[:, 0:cin]) + bias ),  zero padding.
//
// Mapping to the hardware
//   * GEMM view: D[pixels, 32] = sum over 9 taps, cin channels of A_tap[pixels, ch] * W_tap[ch, 32].
//   * Tile = 8 output rows x 30 output columns.  The TMA loads the (8+2) x (30+2) halo ONCE per 16-channel slice
//     as two 8-channel planes [10][32][8ch] (one cp.async.bulk.tensor.4d each; out-of-image coordinates are
//     zero-filled by the TMA = the conv's zero padding).  A plane is exactly the no-swizzle K-major UMMA core-
//     matrix layout with rows = halo pixels in flattened order (16 B per pixel per plane), so the A operand of tap
//     (ky,kx) is the SAME shared-memory tile with the descriptor start address advanced by (ky*32+kx)*16 bytes:
//     nine taps re-use one load.  M = 128 consecutive flattened positions; two M-blocks cover the 8x32 positions
//     (the 2 halo columns per row are computed and discarded: 240/256 useful).
//   * The whole conv's weights (<= 92 KB bf16) stay resident in shared memory for the life of the persistent CTA,
//     pre-packed on the host side of the ABI as the exact B-operand core-matrix image.
//   * Accumulators: 2 M-blocks x 32 fp32 columns in TMEM, double-buffered across tiles (128 columns) so the
//     epilogue of tile i (tcgen05.ld -> +bias -> LeakyReLU -> bf16 -> global, written in place into the dense
//     buffer: the concat is address arithmetic) overlaps the MMAs of tile i+1.
//   * Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread tcgen05.mma issuer,
//     warps 2..5 = epilogue (one TMEM lane quarter each).  8-stage mbarrier ring between TMA and MMA.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.h"
#include "tc_ptx.cuh"

namespace selfc {
namespace tc {

constexpr int WT = 32;                    // tile pitch (30 valid columns + 2 halo)
constexpr int VALID_W = WT - 2;
constexpr int ROWS = 8;                   // output rows per tile
constexpr int HT = ROWS + 2;
constexpr int PLANE_BYTES = HT * WT * 16; // one 8-channel plane of the halo tile
constexpr int KSTEP_BYTES = 2 * PLANE_BYTES;   // 16 channels = one UMMA K step = two planes
constexpr int MAX_KPS = 4;                // K steps per pipeline stage (fewer, larger stages amortise the per-stage
                                          // barrier wait / fence / commit of the single issuing warp: measured ~400-570 cycles)
constexpr int NSTAGE = 8;                 // barrier slots (upper bound on stages)
constexpr int MBLK = 2;                   // 2 x 128 flattened positions per tile
constexpr int NOUT = 32;
constexpr int ACC_COLS = MBLK * NOUT;     // TMEM columns per accumulator buffer
constexpr int TMEM_COLS = 2 * ACC_COLS;   // double buffered
constexpr int MAX_CIN = 160;
constexpr int WTILE_BYTES = NOUT * 16 * 2;   // one (tap, k-step) B tile: 32 x 16 bf16
constexpr int THREADS = 192;

struct SmemLayout {
  // offsets from the 1024-aligned dynamic smem base: [A stages][pad][barriers][weights]
  static constexpr int A_PAD = 1024;      // garbage rows of the last M-block read a little past the tile
  static constexpr int BAR_BYTES = 256;
  __host__ __device__ static int bar_off(int nst, int stage_bytes) { return nst * stage_bytes + A_PAD; }
  static int total(int cin, int nst, int stage_bytes) { return bar_off(nst, stage_bytes) + BAR_BYTES + 9 * (cin / 16) * WTILE_BYTES + 1024; }
};

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// A,B K-major, N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct Params {
  const void* wimg;
  const float* bias;
  __nv_bfloat16* buf;
  int pitch, out_off, N, h, w, nchunk;   // nchunk = cin / 16 (K steps)
  int nst, kps;                          // pipeline stages; K steps (16 channels) per stage
  int tiles_x, tiles_y, ntiles;
  int* err;
};

__global__ void __launch_bounds__(THREADS, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16-byte aligned: round up to 1024 by hand
  const int STAGE = p.kps * KSTEP_BYTES;
  const int NST = p.nst;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const int bar_off = SmemLayout::bar_off(NST, STAGE);
  const uint32_t bar_base = base + bar_off;
  const uint32_t w_base = bar_base + SmemLayout::BAR_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * NSTAGE);
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + 1 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE + 3 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE + 5);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + bar_off + 8 * (2 * NSTAGE + 5));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(w_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();

  const int nchunk = p.nchunk;
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t wbytes = 9u * nchunk * WTILE_BYTES;
      mbar_expect_tx(w_bar, wbytes);
      for (int tap = 0; tap < 9; ++tap)
        bulk_g2s(w_base + tap * nchunk * WTILE_BYTES, (const uint8_t*)p.wimg + (size_t)tap * nchunk * WTILE_BYTES,
                 (uint32_t)nchunk * WTILE_BYTES, w_bar);
      pdl_wait();      // weights are static; activations come from the previous kernel in the stream
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int tx = tile % p.tiles_x;
        const int ty = (tile / p.tiles_x) % p.tiles_y;
        const int n = tile / (p.tiles_x * p.tiles_y);
        const int x0 = tx * VALID_W - 1, y0 = ty * ROWS - 1;
        for (int c0 = 0; c0 < nchunk; c0 += p.kps) {
          const int nk = nchunk - c0 < p.kps ? nchunk - c0 : p.kps;
          mbar_wait(empty_bar(s), ph ^ 1u, p.err, 1);
          mbar_expect_tx(full_bar(s), (uint32_t)nk * KSTEP_BYTES);
          const uint32_t dst = a_base + s * STAGE;
          for (int pl = 0; pl < 2 * nk; ++pl) tma_load_4d(dst + pl * PLANE_BYTES, &tmap, full_bar(s), c0 * 16 + pl * 8, x0, y0, n);
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: whole warp runs the loop, one elected lane issues =====================
    {
      mbar_wait(w_bar, 0, p.err, 2);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(tempty_bar(acc), (use & 1u) ^ 1u, p.err, 3);
        tc_fence_after();
        const uint32_t hi = desc_hi(128, 0);
        const uint32_t b_tap = (uint32_t)nchunk * (WTILE_BYTES >> 4);
        for (int c0 = 0; c0 < nchunk; c0 += p.kps) {
          const int nk = nchunk - c0 < p.kps ? nchunk - c0 : p.kps;
          mbar_wait(full_bar(s), ph, p.err, 4);
          tc_fence_after();
          const uint32_t a_stage = a_base + s * STAGE;
          for (int ks = 0; ks < nk; ++ks) {
            // descriptor low words advance by constants (16-byte units): one halo pixel = 1, one weight tile = 64
            const uint32_t a_lo = desc_lo(a_stage + (uint32_t)ks * KSTEP_BYTES, PLANE_BYTES);
            const uint32_t b_lo = desc_lo(w_base + (uint32_t)(c0 + ks) * WTILE_BYTES, 512);
            const bool first = (c0 + ks) == 0;
#pragma unroll
            for (int mb = 0; mb < MBLK; ++mb) {
              const uint32_t d = tmem_base + (uint32_t)(acc * ACC_COLS + mb * NOUT);
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap % 3;
                const uint64_t ad = desc_join(a_lo + (uint32_t)(mb * 128 + ky * WT + kx), hi);
                const uint64_t bd = desc_join(b_lo + (uint32_t)tap * b_tap, hi);
                umma_bf16_elect(d, ad, bd, kIdesc, (!first || tap > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit_elect(empty_bar(s));      // frees the stage when these MMAs have read it
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
        umma_commit_elect(tfull_bar(acc));      // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    pdl_wait();
    float bias[NOUT];
#pragma unroll
    for (int j = 0; j < NOUT; ++j) bias[j] = __ldg(p.bias + j);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int n = tile / (p.tiles_x * p.tiles_y);
      mbar_wait(tfull_bar(acc), use & 1u, p.err, 5);
      tc_fence_after();
#pragma unroll
      for (int mb = 0; mb < MBLK; ++mb) {
        uint32_t r[NOUT];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + mb * NOUT);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // flattened halo-tile position of this accumulator row
        const int f = WT + 1 + mb * 128 + q * 32 + lane;
        const int fy = f / WT, fx = f % WT;
        const int y = ty * ROWS + fy - 1, x = tx * VALID_W + fx - 1;
        if (fx >= 1 && fx <= VALID_W && fy <= ROWS && y < p.h && x < p.w) {
          __nv_bfloat16* o = p.buf + ((size_t)((size_t)n * p.h + y) * p.w + x) * p.pitch + p.out_off;
#pragma unroll
          for (int j = 0; j < NOUT; j += 8) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = lrelu02(__uint_as_float(r[j + e]) + bias[j + e]);
            uint4 pk;
            __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]);
            __nv_bfloat162 b1 = __floats2bfloat162_rn(v[2], v[3]);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(v[4], v[5]);
            __nv_bfloat162 b3 = __floats2bfloat162_rn(v[6], v[7]);
            pk.x = *reinterpret_cast<uint32_t*>(&b0);
            pk.y = *reinterpret_cast<uint32_t*>(&b1);
            pk.z = *reinterpret_cast<uint32_t*>(&b2);
            pk.w = *reinterpret_cast<uint32_t*>(&b3);
            *reinterpret_cast<uint4*>(o + j) = pk;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---- weight image ------------------------------------------------------------------------------------------
// wref [32][cin_ref][3][3] fp32 -> bf16 image [tap][kstep][kcore(2)][ngroup(4)][n%8][k%8]
__global__ void pack_tc_kernel(const float* __restrict__ wref, __nv_bfloat16* __restrict__ img, int cin_ref, int cin_buf, int xreal,
                               int xpad) {
  const int nchunk = cin_buf / 16;
  const int total = 9 * cin_buf * NOUT;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = idx % NOUT;
  const int c = (idx / NOUT) % cin_buf;
  const int tap = idx / (NOUT * cin_buf);
  const int cref = c < xreal ? c : (c < xpad ? -1 : c - xpad + xreal);
  float v = 0.f;
  if (cref >= 0 && cref < cin_ref) v = wref[((size_t)n * cin_ref + cref) * 9 + tap];
  const int ks = c / 16, kk = c % 16;
  const size_t off = (size_t)(tap * nchunk + ks) * (WTILE_BYTES / 2) + (size_t)((kk / 8) * 4 + n / 8) * 64 + (n % 8) * 8 + (kk % 8);
  img[off] = __float2bfloat16_rn(v);
}

}  // namespace tc

// ---- host side ---------------------------------------------------------------------------------------------
namespace tc {
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int* g_err_flag[64] = {};

int* err_flag_for_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!g_err_flag[dev]) {
    if (cudaMalloc(&g_err_flag[dev], sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(g_err_flag[dev], 0, sizeof(int));
  }
  return g_err_flag[dev];
}

static long long* g_dbg_dev = nullptr;
static long long g_dbg_tags[4096];
static int g_dbg_n = 0;
bool debug_slots() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SELFC_TC_DBG");
    on = (e && atoi(e) != 0) ? 1 : 0;
  }
  return on == 1;
}
long long* debug_next_slot(long long tag) {
  if (!g_dbg_dev) {
    if (cudaMalloc(&g_dbg_dev, 4096 * 16 * sizeof(long long)) != cudaSuccess) return nullptr;
    cudaMemset(g_dbg_dev, 0, 4096 * 16 * sizeof(long long));
  }
  if (g_dbg_n >= 4096) return nullptr;
  g_dbg_tags[g_dbg_n] = tag;
  return g_dbg_dev + 16 * (g_dbg_n++);
}
int debug_read(long long* out, int cap) {
  cudaDeviceSynchronize();
  int n = g_dbg_n < cap ? g_dbg_n : cap;
  for (int i = 0; i < n; ++i) {
    out[17 * i] = g_dbg_tags[i];
    cudaMemcpy(out + 17 * i + 1, g_dbg_dev + 16 * i, 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  }
  g_dbg_n = 0;
  return n;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SELFC_NO_PDL");
    on = (e && atoi(e) != 0) ? 0 : 1;
  }
  return on == 1;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}
}  // namespace tc

int pack_tc_weights(TcConvW& w, const float* wref, const float* bref, int cin_ref, int cin_buf, int xreal, int xpad, cudaStream_t st) {
  SELFC_CHECK_ARG(cin_buf % 16 == 0 && cin_buf <= tc::MAX_CIN, "conv3x3_tc: cin %d must be a multiple of 16 and <= %d", cin_buf,
                  tc::MAX_CIN);
  const size_t bytes = (size_t)9 * (cin_buf / 16) * tc::WTILE_BYTES;
  if (w.img == nullptr || w.img_bytes != bytes) {
    free_tc_weights(w);
    SELFC_CUDA(cudaMalloc(&w.img, bytes));
    SELFC_CUDA(cudaMalloc(&w.bias, tc::NOUT * sizeof(float)));
    w.img_bytes = bytes;
  }
  w.cin_buf = cin_buf;
  const int total = 9 * cin_buf * tc::NOUT;
  tc::pack_tc_kernel<<<cdiv(total, 256), 256, 0, st>>>(wref, reinterpret_cast<__nv_bfloat16*>(w.img), cin_ref, cin_buf, xreal, xpad);
  SELFC_LAUNCH_CHECK("pack_tc_kernel");
  SELFC_CUDA(cudaMemcpyAsync(w.bias, bref, tc::NOUT * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

void free_tc_weights(TcConvW& w) {
  if (w.img2) cudaFree(w.img2);
  w.img2 = nullptr;
  w.img2_bytes = 0;
  if (w.img) cudaFree(w.img);
  if (w.bias) cudaFree(w.bias);
  w.img = nullptr;
  w.bias = nullptr;
  w.img_bytes = 0;
}

int launch_conv3x3_tc(const TcConvW& w, __nv_bfloat16* buf, int pitch, int cin, int out_off, int N, int h, int wd, cudaStream_t st) {
  SELFC_CHECK_ARG(w.img != nullptr && cin == w.cin_buf, "conv3x3_tc: weights not packed for cin=%d", cin);
  SELFC_CHECK_ARG(pitch % 8 == 0 && out_off % 8 == 0 && aligned16(buf), "conv3x3_tc: pitch/out_off/buffer alignment");
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SELFC_E_CUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {(cuuint64_t)pitch, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)N};
  const cuuint64_t gstr[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)wd * pitch * 2, (cuuint64_t)h * wd * pitch * 2};
  const cuuint32_t box[4] = {8u, (cuuint32_t)tc::WT, (cuuint32_t)tc::HT, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (pitch %d, %dx%dx%d)", (int)r, pitch, N, h, wd);
    return SELFC_E_CUDA;
  }
  tc::Params p;
  p.wimg = w.img;
  p.bias = w.bias;
  p.buf = buf;
  p.pitch = pitch;
  p.out_off = out_off;
  p.N = N;
  p.h = h;
  p.w = wd;
  p.nchunk = cin / 16;
  p.tiles_x = cdiv(wd, tc::VALID_W);
  p.tiles_y = cdiv(h, tc::ROWS);
  p.ntiles = p.tiles_x * p.tiles_y * N;
  p.err = tc::err_flag_for_device();
  if (p.ntiles == 0) return 0;
  const int num_sms = tc::num_sms();
  // stage size: as many K steps per stage as leave >= 3 stages next to the resident weights
  static int kps_env = -1;
  if (kps_env < 0) {
    const char* e = getenv("SELFC_TC_KPS");
    kps_env = e ? atoi(e) : 0;
  }
  const int nchunk = cin / 16;
  int kps = kps_env > 0 ? kps_env : tc::MAX_KPS;
  if (kps > tc::MAX_KPS) kps = tc::MAX_KPS;
  if (kps > nchunk) kps = nchunk;
  int nst = 0;
  for (;; --kps) {
    const int stage = kps * tc::KSTEP_BYTES;
    nst = (227 * 1024 - tc::SmemLayout::total(cin, 0, stage)) / stage;
    if (nst >= 3 || kps == 1) break;
  }
  if (nst > tc::NSTAGE) nst = tc::NSTAGE;
  SELFC_CHECK_ARG(nst >= 2, "conv3x3_tc: no room for the A pipeline (cin %d)", cin);
  p.nst = nst;
  p.kps = kps;
  const int smem = tc::SmemLayout::total(cin, nst, kps * tc::KSTEP_BYTES);
  static bool attr_set = false;
  if (!attr_set) {
    SELFC_CUDA(cudaFuncSetAttribute(tc::conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = p.ntiles < num_sms ? p.ntiles : num_sms;
  SELFC_CUDA(tc::launch_pdl(tc::conv3x3_tc_kernel, grid, tc::THREADS, smem, st, tmap, p));
  SELFC_LAUNCH_CHECK("conv3x3_tc_kernel");
  return 0;
}

}  // namespace selfc
