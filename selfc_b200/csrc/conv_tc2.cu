// (1,3,3) dense-block convolution, second tcgen05 formulation: WEIGHTS are the A operand with the three kx taps
// stacked in M, ACTIVATIONS are the B operand with N = pixels (BF16 mode).
//
// Why: with pixels in M and N = 32 output channels (conv_tc.cu) every MMA re-reads a 4 KB activation tile from shared
// memory for 16 cycles of math, and the tile is read once per tap (9x): the UMMA is bound by the shared-memory feed
// (measured ~40-46 cycles per M128xN32xK16 MMA, tensor pipe <= 33 % active).  Here one MMA computes, for one ky,
//     D[kx*32 + n][j] += sum_c W[ky,kx][n][c] * X[c][pos0 + ky*32 + j],      M = 96 (->128), N = 144 positions, K = 16
// so the activation tile is read once per ky (3x instead of 9x) and feeds 4.5x more MACs per byte; the kx shift is
// applied when the accumulator is read back:  out[n][i] = sum_kx D[kx*32 + n][i + kx]  (a TMEM column offset).
// Same halo-tile trick as conv_tc.cu: the TMA loads the (4+2) x 32 halo of a 4x30 output tile once per 16-channel slice
// as two 8-channel planes; a plane is the no-swizzle K-major core-matrix layout with rows = flattened halo positions,
// so the ky shift is a descriptor start-address offset of ky*32 rows.
//
// Per CTA (persistent, 192 threads): warp 0 TMA producer, warp 1 MMA issuer (whole warp, elected lane), warps 2..5
// epilogue.  Accumulator = 144 fp32 columns, double-buffered across tiles (288 TMEM columns).  Epilogue phase 1: the
// three warps owning lane quarters kx = 0,1,2 read their rows with the column offset kx and park the partial sums in
// shared memory [pixel][kx][n]; phase 2: one thread per pixel adds the three partials + bias, LeakyReLU, packs bf16 and
// stores its 64 bytes in place into the dense buffer.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.h"
#include "tc_ptx.cuh"

namespace selfc {
namespace tc2 {

using namespace tc;

constexpr int WT = 32;                        // tile pitch (30 valid columns + 2 halo)
constexpr int VALID_W = WT - 2;
constexpr int ROWS = 8;                       // output rows per tile = 2 M-blocks of 128 flattened positions
constexpr int HT = ROWS + 2;
constexpr int HALO_POS = HT * WT;             // 320 halo positions = rows of the activation tile
constexpr int MBLK = 2;
constexpr int NB = 128;                       // MMA N: valid outputs i <= 125 only need columns i + kx <= 127
constexpr int NOUT = 32;
constexpr int MROWS = 128;                    // MMA M: 3 kx x 32 channels = 96 real rows, padded to 128
constexpr int WTILE_BYTES = MROWS * 16 * 2;   // one (ky, k-step) A tile: 128 x 16 bf16 = 4 KB
constexpr int TMEM_COLS = 512;                // 2 tiles x 2 M-blocks x 128 columns
constexpr int MAX_CIN = 160;
constexpr int THREADS = 192;
constexpr int NSTAGE_MAX = 8;
constexpr int RED_PITCH = 97;                 // floats per pixel in the reduction buffer (3 x 32 + 1: conflict-free reads)
constexpr int RED_BYTES = 128 * RED_PITCH * 4;
constexpr int A_PAD = 1024;
constexpr int BAR_BYTES = 256;

struct Params {
  const void* wimg;
  const float* bias;
  __nv_bfloat16* buf;
  int pitch, out_off, N, h, w, nks;
  int tiles_x, tiles_y, ntiles;
  int nst;
  int dbg;       // experiment knob (SELFC_TC2_DBG): bit 0 = skip the epilogue work, bit 1 = skip the MMAs
  int* err;
};

// KC channels per pipeline stage = one TMA box of HALO_POS rows x KC*2 bytes, swizzled with the matching width
// (KC = 64 -> SWIZZLE_128B, 32 -> 64B, 16 -> 32B): wider rows = fewer, fuller L2 requests.  The swizzle XOR is applied
// by the hardware on absolute shared-memory address bits (verified on B200), so a tile-row shift is just a start offset.
template <int KC>
__global__ void __launch_bounds__(THREADS, 1) conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int ROW_BYTES = KC * 2;
  constexpr int STAGE = HALO_POS * ROW_BYTES;
  constexpr int KSTEPS = KC / 16;
  constexpr uint32_t LAYOUT = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);
  constexpr uint32_t SBO_B = 8 * ROW_BYTES;
  const int NST = p.nst;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  // [A stages (activations)][pad][barriers][reduction buffer][weights]
  const uint32_t a_base = base;
  const int bar_off = NST * STAGE + A_PAD;
  const uint32_t bar_base = base + bar_off;
  const int red_off = bar_off + BAR_BYTES;
  const uint32_t w_base = base + red_off + RED_BYTES;
  float* red = reinterpret_cast<float*>(gen_base + red_off);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE_MAX + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * NSTAGE_MAX);
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE_MAX + 1 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE_MAX + 3 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE_MAX + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + bar_off + 8 * (2 * NSTAGE_MAX + 5));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSTAGE_MAX; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(w_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();

  const int nks = p.nks;
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t wbytes = 3u * nks * WTILE_BYTES;
      mbar_expect_tx(w_bar, wbytes);
      for (int ky = 0; ky < 3; ++ky)
        bulk_g2s(w_base + ky * nks * WTILE_BYTES, (const uint8_t*)p.wimg + (size_t)ky * nks * WTILE_BYTES, (uint32_t)nks * WTILE_BYTES,
                 w_bar);
      pdl_wait();
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int tx = tile % p.tiles_x;
        const int ty = (tile / p.tiles_x) % p.tiles_y;
        const int n = tile / (p.tiles_x * p.tiles_y);
        const int x0 = tx * VALID_W - 1, y0 = ty * ROWS - 1;
        for (int c0 = 0; c0 < nks; c0 += KSTEPS) {
          mbar_wait(empty_bar(s), ph ^ 1u, p.err, 21);
          mbar_expect_tx(full_bar(s), (uint32_t)STAGE);
          tma_load_4d(a_base + s * STAGE, &tmap, full_bar(s), c0 * 16, x0, y0, n);
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: whole warp runs the loop, one elected lane issues =====================
    constexpr uint32_t idesc = umma_idesc_bf16(MROWS, NB);
    mbar_wait(w_bar, 0, p.err, 22);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    const uint32_t hi_a = desc_hi(128, 0);            // weights: no-swizzle core matrices
    const uint32_t hi_b = desc_hi(SBO_B, LAYOUT);      // activations: swizzled rows
    const uint32_t a_ky = (uint32_t)nks * (WTILE_BYTES >> 4);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(tempty_bar(acc), (use & 1u) ^ 1u, p.err, 23);
      tc_fence_after();
      for (int c0 = 0; c0 < nks; c0 += KSTEPS) {
        const int nk = nks - c0 < KSTEPS ? nks - c0 : KSTEPS;
        mbar_wait(full_bar(s), ph, p.err, 24);
        tc_fence_after();
        const uint32_t a_stage = a_base + s * STAGE;
        for (int ks = 0; ks < nk; ++ks) {
          // A (weights): (ky, k-step) tiles of 4 KB, K-core stride 2 KB
          const uint32_t a_lo = desc_lo(w_base + (uint32_t)(c0 + ks) * WTILE_BYTES, 2048);
          // B (activations): rows = flattened halo positions; K step = +32 bytes inside the swizzled row
          const uint32_t b_lo = desc_lo(a_stage + (uint32_t)ks * 32u, 16);
#pragma unroll
          for (int mb = 0; mb < MBLK; ++mb) {
            const uint32_t d = tmem_base + (uint32_t)((acc * MBLK + mb) * NB);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint64_t ad = desc_join(a_lo + (uint32_t)ky * a_ky, hi_a);
              const uint64_t bd = desc_join(b_lo + (uint32_t)((mb * 128 + ky * WT) * (ROW_BYTES >> 4)), hi_b);
              if (!(p.dbg & 2)) umma_bf16_elect(d, ad, bd, idesc, ((c0 + ks) > 0 || ky > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit_elect(empty_bar(s));
        if (++s == NST) { s = 0; ph ^= 1u; }
      }
      umma_commit_elect(tfull_bar(acc));
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp & 3;               // TMEM lane quarter = kx (quarter 3 holds the zero rows 96..127)
    const int et = q * 32 + lane;         // epilogue thread index 0..127 = output position of phase 2
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
      mbar_wait(tfull_bar(acc), use & 1u, p.err, 25);
      tc_fence_after();
     for (int mb = 0; mb < MBLK; ++mb) {
      // ---- phase 1: rows kx*32 + lane, columns i + kx -> partial[i][kx][lane] in shared memory ----
      if (q < 3 && !(p.dbg & 1)) {
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MBLK + mb) * NB);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          uint32_t r[16];
          tmem_ld16(trow + (uint32_t)(c * 16 + q), r);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) red[(c * 16 + e) * RED_PITCH + q * 32 + lane] = __uint_as_float(r[e]);
        }
        {
          // last chunk: columns 112+kx .. 127 (the missing kx columns only feed the invalid outputs i >= 126)
          uint32_t r[16];
          tmem_ld16(trow + (uint32_t)(NB - 16), r);
          tmem_ld_wait();
          if (q == 0) {
#pragma unroll
            for (int e = 0; e < 16; ++e) red[(112 + e) * RED_PITCH + lane] = __uint_as_float(r[e]);
          } else if (q == 1) {
#pragma unroll
            for (int e = 0; e < 15; ++e) red[(112 + e) * RED_PITCH + 32 + lane] = __uint_as_float(r[e + 1]);
          } else {
#pragma unroll
            for (int e = 0; e < 14; ++e) red[(112 + e) * RED_PITCH + 64 + lane] = __uint_as_float(r[e + 2]);
          }
        }
      }
      if (mb == MBLK - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));      // both accumulators drained: the next-but-one tile may start
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");       // partials visible to the 4 epilogue warps
      // ---- phase 2: one thread per output position ----
      {
        const int f = WT + 1 + mb * 128 + et;              // flattened halo-tile position
        const int fy = f / WT, fx = f % WT;
        const int y = ty * ROWS + fy - 1, x = tx * VALID_W + fx - 1;
        if (fx >= 1 && fx <= VALID_W && fy <= ROWS && y < p.h && x < p.w && !(p.dbg & 1)) {
          const float* pr = red + et * RED_PITCH;
          __nv_bfloat16* o = p.buf + ((size_t)((size_t)n * p.h + y) * p.w + x) * p.pitch + p.out_off;
#pragma unroll
          for (int j = 0; j < NOUT; j += 8) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = lrelu02(pr[j + e] + pr[32 + j + e] + pr[64 + j + e] + bias[j + e]);
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
      asm volatile("bar.sync 1, 128;" ::: "memory");       // reduction buffer free for the next block
     }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)TMEM_COLS);
  }
}

// wref [32][cin_ref][3][3] fp32 -> bf16 A-operand image [ky][kstep][kcore(2)][mgroup(16)][r%8][k%8], row r = kx*32 + n
__global__ void pack_tc2_kernel(const float* __restrict__ wref, __nv_bfloat16* __restrict__ img, int cin_ref, int cin_buf, int xreal,
                                int xpad) {
  const int nks = cin_buf / 16;
  const int total = 3 * cin_buf * MROWS;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int r = idx % MROWS;
  const int c = (idx / MROWS) % cin_buf;
  const int ky = idx / (MROWS * cin_buf);
  const int kx = r / 32, n = r % 32;
  const int cref = c < xreal ? c : (c < xpad ? -1 : c - xpad + xreal);
  float v = 0.f;
  if (kx < 3 && cref >= 0 && cref < cin_ref) v = wref[((size_t)n * cin_ref + cref) * 9 + ky * 3 + kx];
  const int ks = c / 16, kk = c % 16;
  const size_t off = (size_t)(ky * nks + ks) * (WTILE_BYTES / 2) + (size_t)((kk / 8) * 16 + r / 8) * 64 + (r % 8) * 8 + (kk % 8);
  img[off] = __float2bfloat16_rn(v);
}

}  // namespace tc2

bool conv3x3_tc2_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SELFC_TC_CONV2");
    on = e ? (atoi(e) != 0 ? 1 : 0) : 1;
  }
  return on == 1;
}

int pack_tc2_weights(TcConvW& w, const float* wref, int cin_ref, int cin_buf, int xreal, int xpad, cudaStream_t st) {
  SELFC_CHECK_ARG(cin_buf % 16 == 0 && cin_buf <= tc2::MAX_CIN, "conv3x3_tc2: cin %d must be a multiple of 16 and <= %d", cin_buf,
                  tc2::MAX_CIN);
  const size_t bytes = (size_t)3 * (cin_buf / 16) * tc2::WTILE_BYTES;
  if (w.img2 == nullptr || w.img2_bytes != bytes) {
    if (w.img2) cudaFree(w.img2);
    w.img2 = nullptr;
    SELFC_CUDA(cudaMalloc(&w.img2, bytes));
    w.img2_bytes = bytes;
  }
  const int total = 3 * cin_buf * tc2::MROWS;
  tc2::pack_tc2_kernel<<<cdiv(total, 256), 256, 0, st>>>(wref, reinterpret_cast<__nv_bfloat16*>(w.img2), cin_ref, cin_buf, xreal, xpad);
  SELFC_LAUNCH_CHECK("pack_tc2_kernel");
  return 0;
}

int launch_conv3x3_tc2(const TcConvW& w, __nv_bfloat16* buf, int pitch, int cin, int out_off, int N, int h, int wd, cudaStream_t st) {
  SELFC_CHECK_ARG(w.img2 != nullptr && cin == w.cin_buf, "conv3x3_tc2: weights not packed for cin=%d", cin);
  SELFC_CHECK_ARG(pitch % 8 == 0 && out_off % 8 == 0 && aligned16(buf), "conv3x3_tc2: pitch/out_off/buffer alignment");
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SELFC_E_CUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {(cuuint64_t)pitch, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)N};
  const cuuint64_t gstr[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)wd * pitch * 2, (cuuint64_t)h * wd * pitch * 2};
  static int kc_pref = 0;     // SELFC_TC2_KC = 64 | 32 | 16: channels per TMA box / swizzle width of the activation tile
  if (!kc_pref) {
    const char* e = getenv("SELFC_TC2_KC");
    kc_pref = e ? atoi(e) : 32;
    if (kc_pref != 64 && kc_pref != 32 && kc_pref != 16) kc_pref = 32;
  }
  // widest box that still leaves two pipeline stages next to the resident weights and the reduction buffer
  const int fixed = tc2::A_PAD + tc2::BAR_BYTES + tc2::RED_BYTES + (int)w.img2_bytes + 1024;
  int kc = kc_pref;
  while (kc > 16 && (227 * 1024 - fixed) / (tc2::HALO_POS * kc * 2) < 2) kc >>= 1;
  const cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)tc2::WT, (cuuint32_t)tc2::HT, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (pitch %d, %dx%dx%d)", (int)r, pitch, N, h, wd);
    return SELFC_E_CUDA;
  }
  tc2::Params p;
  p.wimg = w.img2;
  p.bias = w.bias;
  p.buf = buf;
  p.pitch = pitch;
  p.out_off = out_off;
  p.N = N;
  p.h = h;
  p.w = wd;
  p.nks = cin / 16;
  p.tiles_x = cdiv(wd, tc2::VALID_W);
  p.tiles_y = cdiv(h, tc2::ROWS);
  p.ntiles = p.tiles_x * p.tiles_y * N;
  p.err = tc::err_flag_for_device();
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("SELFC_TC2_DBG");
    dbg = e ? atoi(e) : 0;
  }
  p.dbg = dbg;
  if (p.ntiles == 0) return 0;
  const int stage = tc2::HALO_POS * kc * 2;
  int nst = (227 * 1024 - fixed) / stage;
  if (nst > tc2::NSTAGE_MAX) nst = tc2::NSTAGE_MAX;
  SELFC_CHECK_ARG(nst >= 2, "conv3x3_tc2: no room for the activation pipeline (cin %d)", cin);
  p.nst = nst;
  const int smem = fixed + nst * stage;
  static bool attr_set = false;
  if (!attr_set) {
    SELFC_CUDA(cudaFuncSetAttribute(tc2::conv3x3_tc2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(tc2::conv3x3_tc2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(tc2::conv3x3_tc2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int nsm = tc::num_sms();
  const int grid = p.ntiles < nsm ? p.ntiles : nsm;
  if (kc == 64) SELFC_CUDA(tc::launch_pdl(tc2::conv3x3_tc2_kernel<64>, grid, tc2::THREADS, smem, st, tmap, p));
  else if (kc == 32) SELFC_CUDA(tc::launch_pdl(tc2::conv3x3_tc2_kernel<32>, grid, tc2::THREADS, smem, st, tmap, p));
  else SELFC_CUDA(tc::launch_pdl(tc2::conv3x3_tc2_kernel<16>, grid, tc2::THREADS, smem, st, tmap, p));
  SELFC_LAUNCH_CHECK("conv3x3_tc2_kernel");
  return 0;
}

}  // namespace selfc
