#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace b200 {
constexpr int kMaxGroups = 256;  // objects per readout launch
constexpr int kMaxPeers = 8;     // ranks of one NVLink box (scatter-reduce mode)

// out[out_row[g] + r, q] = sum_n values[val_row[g] + r, n] * P[q, n]
// for g < n_groups, r < rows_per_group, q < q, n < n_window.  val_row/out_row are HOST arrays.
int launch_readout(const __half* values, long long values_ld, long long values_rows, const int* val_row,
                   const int* out_row, int n_groups, int rows_per_group, const __half* P, long long ldP,
                   int n_window, int q, float* out, long long ldo, __half* out_tok, cudaStream_t stream);
// out_tok (optional, replaces `out`): fp16 token-major [object, q, rows_per_group] with object = out_row / rows_per_group

// Fused sparse-affinity variant: the B operand is generated in shared memory from the top-k lists
// (idx/w: [q, 32] as written by launch_sim_topk), no dense affinity.  workspace: readout_sparse_workspace_bytes().
size_t readout_sparse_workspace_bytes(int q, int n_window);
int launch_readout_sparse(const __half* values, long long values_ld, long long values_rows, const int* val_row,
                          const int* out_row, int n_groups, int rows_per_group, const int* idx, const float* w,
                          int top_k, int n_window, int q, void* workspace, float* out, long long ldo, __half* out_tok,
                          cudaStream_t stream, const int* owner = nullptr, float* const* rank_dst = nullptr,
                          int n_ranks = 0);
// owner / rank_dst (scatter-reduce, bank-sharded read): tile of group g is red.add'ed into
// rank_dst[owner[g]] + out_row[g] * ldo - the buffer of the rank that owns the object (peer memory over NVLink).
}  // namespace b200
