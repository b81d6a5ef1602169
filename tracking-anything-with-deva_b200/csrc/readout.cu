// Memory readout  out[K*CV, Q] = V[K*CV, N] . P[N, Q]   (reference: MemoryManager._readout,
// deva/inference/memory_manager.py:64-75; training twin memory_utils.py:87-94).
//
// tcgen05 GEMM, fp16 operands / fp32 accumulation in TMEM.
//   A operand = value bank rows (M = value channels of the active objects), K-major: the bank
//               keeps every value row contiguous along the memory-slot axis N.
//   B operand = dense affinity P stored [Q, N] (K-major), produced by the top-k/softmax stage.
//   D         = 128 x 256 fp32 tile per CTA, double-buffered in TMEM (2 x 256 columns).
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-5 epilogue.
// Persistent CTAs (one per SM) walk a grouped raster of the tile grid so that CTAs running
// together share value panels and affinity panels in L2.
#include <cuda_fp16.h>

#include "common.h"
#include "ptx.cuh"
#include "readout.h"
#include "tmap.h"

namespace b200 {
namespace readout {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int GROUP_M = 16;

struct Params {
  int m_tiles, n_tiles, k_blocks, tiles_per_group;
  int q;
  long long ldo;
  float* out;
  __half* out_tok;  // optional fp16 token-major output [object, q, rows_per_group]
  int rows_per_group;
  int val_row[kMaxGroups];
  int out_row[kMaxGroups];
};

__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int& m, int& n) {
  const int per_band = GROUP_M * n_tiles;
  const int band = tile / per_band;
  const int first_m = band * GROUP_M;
  const int rows = min(GROUP_M, m_tiles - first_m);
  const int r = tile - band * per_band;
  m = first_m + r % rows;
  n = r / rows;
}

__global__ void __launch_bounds__(THREADS, 1)
readout_kernel(const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_p,
               const __grid_constant__ Params p) {
  // 1024-byte alignment (128B swizzle atoms) comes from the declaration: deriving an aligned pointer through an
  // integer cast would make the compiler lose the shared address space (generic LD/ST instead of LDS/STS).
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_v);
    tma_prefetch_desc(&map_p);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m, n;
        tile_coords(tile, p.m_tiles, p.n_tiles, m, n);
        const int g = m / p.tiles_per_group;
        const int a_row = p.val_row[g] + (m - g * p.tiles_per_group) * BM;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          tma_load_2d(sA + stage * A_BYTES, &map_v, &full[stage], kb * BK, a_row);
          tma_load_2d(sB + stage * B_BYTES, &map_p, &full[stage], kb * BK, n * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(0, BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_f16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                     (kb | k) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);
      }
    }
  } else {
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    const bool vec_ok = (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int m, n;
      tile_coords(tile, p.m_tiles, p.n_tiles, m, n);
      const int g = m / p.tiles_per_group;
      const long long o_row = p.out_row[g] + (long long)(m - g * p.tiles_per_group) * BM + row;
      float* dst = p.out + o_row * p.ldo;
      const int acc = it & 1;
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + c * 32, r);
        tmem_ld_wait();
        const int q0 = n * BN + c * 32;
        if (p.out_tok) {
          const long long R = p.out_row[g] + (long long)(m - g * p.tiles_per_group) * BM + row;
          const long long obj = R / p.rows_per_group;
          const int ch = (int)(R - obj * p.rows_per_group);
          __half* dt = p.out_tok + (obj * p.q) * p.rows_per_group + ch;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (q0 + j < p.q) dt[(long long)(q0 + j) * p.rows_per_group] = __float2half_rn(__uint_as_float(r[j]));
        } else if (q0 + 32 <= p.q && vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                   __uint_as_float(r[j + 3]));
            *reinterpret_cast<float4*>(dst + q0 + j) = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (q0 + j < p.q) dst[q0 + j] = __uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace readout

int launch_readout(const __half* values, long long values_ld, long long values_rows, const int* val_row,
                   const int* out_row, int n_groups, int rows_per_group, const __half* P, long long ldP,
                   int n_window, int q, float* out, long long ldo, __half* out_tok, cudaStream_t stream) {
  using namespace readout;
  B200_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups, "readout: n_groups %d out of range [1,%d]", n_groups,
               kMaxGroups);
  B200_REQUIRE(rows_per_group % BM == 0, "readout: rows_per_group %d must be a multiple of %d", rows_per_group, BM);
  B200_REQUIRE(n_window >= 1 && q >= 1, "readout: empty problem");
  B200_REQUIRE(values_ld % 8 == 0 && ldP % 8 == 0, "readout: leading dimensions must be multiples of 8 halves");
  const char* err = nullptr;
  CUtensorMap map_v, map_p;
  if (make_tmap_2d(&map_v, TmapType::F16, values, n_window, values_rows, values_ld * 2, BK, BM, &err) ||
      make_tmap_2d(&map_p, TmapType::F16, P, n_window, q, ldP * 2, BK, BN, &err)) {
    set_error("readout: %s", err ? err : "tensor map");
    return 3;
  }
  Params p;
  p.tiles_per_group = rows_per_group / BM;
  p.m_tiles = n_groups * p.tiles_per_group;
  p.n_tiles = ceil_div(q, BN);
  p.k_blocks = ceil_div(n_window, BK);
  p.q = q;
  p.ldo = ldo;
  p.out = out;
  p.out_tok = out_tok;
  p.rows_per_group = rows_per_group;
  for (int i = 0; i < n_groups; ++i) {
    p.val_row[i] = val_row[i];
    p.out_row[i] = out_row[i];
  }
  static bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(readout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < sm_count() ? total : sm_count();
  readout_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(map_v, map_p, p);
  B200_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200
