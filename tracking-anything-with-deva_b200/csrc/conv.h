#pragma once
#include <cuda_runtime.h>

namespace b200 {
constexpr int kMaxTaps = 32;
constexpr int kMaxHead = 9;

// Host-side description of one convolution launch (all device pointers unless noted).
struct ConvDesc {
  const void* x;         // fp16 NHWC [batch, h, w, cin_pad]
  const void* x2;        // optional second input of the same shape: the conv sees cat[x, x2] along channels
  const void* x_lo;      // optional fp16 low-order part of x: split-precision mode, D = Xh.Wh + Xl.Wh + Xh.Wl
                         // (weights packed [cout_pad, 2 (hi, lo), taps, cin_pad]); excludes x2
  int split_mode;        // 0: x_lo => three passes (Xh.Wh + Xl.Wh + Xh.Wl); 1: x_lo with single weights (Xh.W + Xl.W);
                         // 2: no x_lo, weights packed (hi, lo) (X.Wh + X.Wl)
  int batch, h, w, cin_pad;
  const void* w_packed;  // fp16 [cout_pad, kh*kw*cin_pad]
  int kh, kw, stride;    // 1x1 / 3x3, stride 1 / 2, padding kh/2
  int cout, cout_pad, nt;
  int th, tw;            // spatial tile, th*tw == 128
  const float* bias;     // [cout_pad] fp32
  const void* res;       // optional fp16 NHWC residual, same shape as the output (or one image if res_broadcast)
  const void* res_lo;    // optional low-order part of the residual
  int res_broadcast;
  const float* rank1_w;  // optional [cout_pad]
  const float* rank1_x;  // optional [batch, ho*wo]
  void* out_raw;         // optional fp16 NHWC
  void* out_relu;        // optional fp16 NHWC, max(.,0)
  float* out_f32;        // optional fp32 NHWC
  void* out_raw_lo;      // optional fp16 low-order parts: value - fp16(value) of out_raw / out_relu
  void* out_relu_lo;
  // optional fused 1x1 'head' on the ReLU'd fp32 result (needs cout_pad == nt): head_out[pixel, t] =
  // sum_c relu(out[pixel, c]) * head_w[t, c], t < head_n <= kMaxHead.  Used to fold the 3x3 logit conv
  // (MaskDecoder.pred) into the last decoder conv: the 9 taps are gathered afterwards by ew_head_gather3x3.
  const float* head_w;   // [head_n, cout] fp32
  float* head_out;       // [batch*ho*wo, head_n] fp32
  int head_n;
  // optional fused gate epilogue (see include/deva_b200.h): h' = f*h*(1-u) + u*tanh(n) on the fp32 accumulators
  const void* gate_h;
  void* gate_out;
  int ksplit;            // > 1: split the K loop; out_f32 is [ksplit_effective, batch, ho, wo, cout] partial sums
  // split_mode 3: D = (X . W16 + Xlo8 . W8) * acc_scale with W16 = fp16(W * 2^S), W8 = e4m3(W * 2^(S-12)),
  // Xlo8 = e4m3((x - fp16(x)) * 4096), acc_scale = 2^-S: the low-order activation pass on the fp8 tensor-core path
  const void* x_lo8;     // u8 NHWC [batch, h, w, cin_pad]
  const void* w8_packed; // u8 [cout_pad, kh*kw*cin_pad]
  float acc_scale;       // 0 = 1
  void* out_relu_lo8;    // optional u8 NHWC: e4m3 low-order part of the ReLU'd output (x 4096)
};

int launch_conv(const ConvDesc& d, cudaStream_t stream);
}  // namespace b200
