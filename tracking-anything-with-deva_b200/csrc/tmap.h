// Host-side TMA tensor-map construction.  cuTensorMapEncodeTiled is a driver-API symbol;
// it is resolved at run time through the CUDA runtime so the library links (and loads on a
// GPU-less build box) without libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum class TmapType { F16, F32, BF16 };

// Returns 0 on success; on failure returns non-zero and `err` points at a static message.
int make_tmap_2d(CUtensorMap* out, TmapType type, const void* base, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, const char** err);

// NHWC activation map for implicit-GEMM convolution: dims (C, W, H, N), 128-byte swizzle,
// box (box_c, box_w, box_h, 1).  Out-of-range coordinates (conv padding) are zero-filled.
int make_tmap_nhwc(CUtensorMap* out, TmapType type, const void* base, uint64_t c, uint64_t w, uint64_t h,
                   uint64_t n, uint32_t box_c, uint32_t box_w, uint32_t box_h, const char** err);

}  // namespace b200
