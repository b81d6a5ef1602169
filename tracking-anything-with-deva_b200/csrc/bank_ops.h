#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace b200 {
int launch_pack_query(const float* qk, const float* qe, long long sc, long long sq, int ck, int q, __half* hi,
                      __half* lo, float* bsq, cudaStream_t stream);
int launch_pack_keys(const float* key, const float* sel, long long sc, long long sn, const float* shr, int ck, int n,
                     __half* hi, __half* lo, float* neg_s, float* raw_key, float* raw_sel, float* raw_shr,
                     cudaStream_t stream);
int launch_append_values(const float* src, long long ld_src, __half* dst, long long ld_dst, int rows, int n,
                         cudaStream_t stream);
int launch_gather_rows(void* dst, const void* src, const int* idx, int n, int row_bytes, cudaStream_t stream);
int launch_gather_f32(float* dst, const float* src, const int* idx, int n, cudaStream_t stream);
int launch_gather_cols_f16(__half* dst, long long ld_dst, const __half* src, long long ld_src, const int* idx,
                           int rows, int n, cudaStream_t stream);
int launch_usage(float* out, const float* use, const float* life, int n, cudaStream_t stream);
}  // namespace b200
