// Host-side TMA tensor-map construction.  cuTensorMapEncodeTiled is a driver-API symbol;
// it is resolved at run time through the CUDA runtime so the library links (and loads on a
// GPU-less build box) without libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum class TmapType { F16, F32, BF16, U8 };

// Returns 0 on success; on failure returns non-zero and `err` points at a static message.
int make_tmap_2d(CUtensorMap* out, TmapType type, const void* base, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, const char** err);

// fp16 activation map for implicit-GEMM convolution: rank-5 dims (C, W, H, B, 1) with explicit element
// strides for W / H / B (so strided "phase" views of an NHWC buffer can be described), 128-byte swizzle,
// box (64, box_w, box_h, 1, 1).  Out-of-range coordinates (conv padding) are zero-filled.
int make_tmap_act5(CUtensorMap* out, const void* base, uint64_t c, uint64_t w, uint64_t h, uint64_t b,
                   long long stride_w, long long stride_h, long long stride_b, uint32_t box_w, uint32_t box_h,
                   const char** err);

// Same for an 8-bit (e4m3) activation tensor: dims (C bytes, W, H, B, 1), box (128, box_w, box_h, 1, 1), strides in bytes.
int make_tmap_act5_u8(CUtensorMap* out, const void* base, uint64_t c, uint64_t w, uint64_t h, uint64_t b,
                      long long stride_w, long long stride_h, long long stride_b, uint32_t box_w, uint32_t box_h,
                      const char** err);

// General rank-5 fp16 tiled map, 128-byte swizzle: dims / box in elements, strides (dims 1..4) in elements.
int make_tmap_f16_5d(CUtensorMap* out, const void* base, const uint64_t (&dims)[5], const long long (&strides)[4],
                     const uint32_t (&box)[5], const char** err);

}  // namespace b200
