#pragma once
#include <cuda_runtime.h>

namespace b200 {
constexpr int kMaxTaps = 20;

// Host-side description of one convolution launch (all device pointers unless noted).
struct ConvDesc {
  const void* x;         // fp16 NHWC [batch, h, w, cin_pad]
  const void* x2;        // optional second input of the same shape: the conv sees cat[x, x2] along channels
  int batch, h, w, cin_pad;
  const void* w_packed;  // fp16 [cout_pad, kh*kw*cin_pad]
  int kh, kw, stride;    // 1x1 / 3x3, stride 1 / 2, padding kh/2
  int cout, cout_pad, nt;
  int th, tw;            // spatial tile, th*tw == 128
  const float* bias;     // [cout_pad] fp32
  const void* res;       // optional fp16 NHWC residual, same shape as the output (or one image if res_broadcast)
  int res_broadcast;
  const float* rank1_w;  // optional [cout_pad]
  const float* rank1_x;  // optional [batch, ho*wo]
  void* out_raw;         // optional fp16 NHWC
  void* out_relu;        // optional fp16 NHWC, max(.,0)
  float* out_f32;        // optional fp32 NHWC
};

int launch_conv(const ConvDesc& d, cudaStream_t stream);
}  // namespace b200
