// tcgen05 / TMEM / TMA implicit-GEMM convolution (BF16 mode) -- host-side interface.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace selfc {

// One (1,3,3) convolution's weights as the exact shared-memory image the UMMA B-operand descriptors read
// (see conv_tc.cu), plus its fp32 bias.
struct TcConvW {
  void* img = nullptr;      // device, bf16
  float* bias = nullptr;    // device, [32]
  size_t img_bytes = 0;
  int cin_buf = 0;          // input channels consumed from the dense buffer (multiple of 16)
};

int pack_tc_weights(TcConvW& w, const float* wref, const float* bref, int cin_ref, int cin_buf, int xreal, int xpad, cudaStream_t st);
void free_tc_weights(TcConvW& w);
// conv_k of a dense block, in place: reads channels [0,cin) of buf, writes lrelu(conv+bias) to [out_off,out_off+32)
int launch_conv3x3_tc(const TcConvW& w, __nv_bfloat16* buf, int pitch, int cin, int out_off, int N, int h, int wd, cudaStream_t st);

}  // namespace selfc
