// Anisotropic-L2 similarity + per-query top-k + softmax, fused
// (reference: get_similarity + do_softmax, deva/model/memory_utils.py:6-76).
//
// sim[n,q] = s_n * ( sum_c mk_cn^2 * (-qe_cq) + sum_c mk_cn * (2 qk_cq qe_cq) - bsq_q ),  s_n = shrinkage_n / sqrt(CK)
// runs as ONE tcgen05 GEMM with K = 2*CK:   key row n = [ s*mk^2 | s*mk ],  query row q = [ -qe | 2*qk*qe ],
// and the  - s_n * bsq_q  term is applied in the epilogue.  fp32-equivalent accuracy comes from a
// three-term split of both operands into fp16 (hi, lo) pairs: D = Qh.Kh + Ql.Kh + Qh.Kl, fp32 accumulate in TMEM
// (11+11 significand bits; the dropped Ql.Kl term is ~2^-22 relative).
//
// CTA = 128 queries (TMEM lanes) x a contiguous range of 128-slot memory tiles; the query operand is TMA-loaded
// once and stays resident, key tiles stream through a 2-stage ring.  Each epilogue thread owns one query: it scans
// the accumulator row with a running threshold (current k-th best) and keeps its k best (value, slot) pairs in a
// bank-conflict-free shared-memory column.  The memory axis can be split over several CTAs (grid.y); a small merge
// kernel picks the global top-k, applies the softmax, accumulates per-slot usage and scatters the dense fp16
// affinity rows consumed by the readout GEMM.
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.h"
#include "ptx.cuh"
#include "simtopk.h"
#include "tmap.h"

namespace b200 {
namespace simtopk {

constexpr int BQ = 128, BNK = 64, BK = 64;
constexpr int QTILE_BYTES = BQ * 128;   // 128 query rows x 64 halves
constexpr int KTILE_BYTES = BNK * 128;  // 64 key rows x 64 halves
constexpr int MAX_KB = 2;               // 2*CK/64 with CK <= 64
constexpr int KSTAGES = 3;
constexpr int ACC_STAGES = 4;           // 4 x 64 TMEM columns; epilogue group g owns stages g, g+2
constexpr int GROUPS = 2;               // epilogue warp groups working on alternating tiles
constexpr int THREADS = 64 + GROUPS * 128;
constexpr int Q_BYTES = 2 * MAX_KB * QTILE_BYTES;  // hi/lo x k-blocks
constexpr int KSTAGE_BYTES = 2 * MAX_KB * KTILE_BYTES;
constexpr int LIST_BYTES = GROUPS * 2 * kListCap * BQ * 4;
constexpr int NS_BYTES = ACC_STAGES * BNK * 4;
static_assert(GROUPS == 2, "the threshold exchange pairs group g with group g ^ 1");
constexpr int THR_BYTES = GROUPS * BQ * 4;  // each epilogue group publishes its running k-th best per query
constexpr int SMEM_BYTES = Q_BYTES + KSTAGES * KSTAGE_BYTES + LIST_BYTES + NS_BYTES + 256 + THR_BYTES;
static_assert(SMEM_BYTES <= 232448, "simtopk: shared-memory budget");

struct Params {
  int q, n_window, n_lead, kblocks, top_k;
  int tiles_total, tiles_per_split;
  int qpad;
  const float* neg_s;   // [n_window]  -shrinkage/sqrt(CK)
  const float* bsq;     // [q]
  float* part_val;      // [nsplit*GROUPS][kListCap][qpad]
  int* part_idx;
  float* dense_out;     // DENSE mode: [q][ld_dense]
  long long ld_dense;
  const float* thr_floor;  // optional [q]: a lower bound of each query's k-th best similarity (temporal warm start)
};

// v[j] for a run-time j without spilling v to local memory: 5-level select tree (31 selects).
__device__ __forceinline__ float pick32(const float (&v)[32], int j) {
  float a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 16) ? v[i + 16] : v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 8) ? a[i + 8] : a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[i + 4] : b[i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 2) ? c[i + 2] : c[i];
  return (j & 1) ? d[1] : d[0];
}

// Per-thread running top-k: kListCap (32) candidate slots in a shared-memory column (entries >= top_k hold
// +inf), tracked as 4 groups of 8 with the group minima (value, position) in registers.  Replacing the global
// minimum re-scans only its group: 8 shared loads + a handful of selects per insertion.
struct TopK {
  float gmin[4];
  int gpos[4];
  float thr;
  int minpos;
  __device__ __forceinline__ void refresh() {
    float m01 = gmin[0]; int p01 = gpos[0];
    if (gmin[1] < m01) { m01 = gmin[1]; p01 = gpos[1]; }
    float m23 = gmin[2]; int p23 = gpos[2];
    if (gmin[3] < m23) { m23 = gmin[3]; p23 = gpos[3]; }
    const bool c = m23 < m01;
    thr = c ? m23 : m01;
    minpos = c ? p23 : p01;
  }
  __device__ __forceinline__ void insert(float* lv, int* li, int row, float v, int n) {
    lv[minpos * BQ + row] = v;
    li[minpos * BQ + row] = n;
    const int g = minpos >> 3;
    const float* base = lv + (g * 8) * BQ + row;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = base[e * BQ];
    float m = x[0]; int mp = 0;
#pragma unroll
    for (int e = 1; e < 8; ++e)
      if (x[e] < m) { m = x[e]; mp = e; }
    mp += g * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i == g) { gmin[i] = m; gpos[i] = mp; }
    refresh();
  }
};

template <bool DENSE>
__global__ void __launch_bounds__(THREADS, 1)
simtopk_kernel(const __grid_constant__ CUtensorMap map_qh, const __grid_constant__ CUtensorMap map_ql,
               const __grid_constant__ CUtensorMap map_kh, const __grid_constant__ CUtensorMap map_kl,
               const __grid_constant__ Params p) {
  // 1024-byte alignment (128B swizzle atoms) comes from the declaration: deriving an aligned pointer through an
  // integer cast would make the compiler lose the shared address space (generic LD/ST instead of LDS/STS).
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                                   // [hi/lo][kb] query tiles, resident
  uint8_t* sK = smem + Q_BYTES;                         // [stage][hi/lo][kb] key tiles
  float* list_val = reinterpret_cast<float*>(sK + KSTAGES * KSTAGE_BYTES);   // [group][kListCap][BQ]
  int* list_idx = reinterpret_cast<int*>(list_val + GROUPS * kListCap * BQ);
  float* ns = reinterpret_cast<float*>(list_idx + GROUPS * kListCap * BQ);   // [ACC_STAGES][BNK]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ns + ACC_STAGES * BNK);
  uint64_t* q_full = bars;
  uint64_t* full = bars + 1;                    // [KSTAGES]
  uint64_t* empty = full + KSTAGES;             // [KSTAGES]
  uint64_t* acc_full = empty + KSTAGES;         // [ACC_STAGES]
  uint64_t* acc_empty = acc_full + ACC_STAGES;  // [ACC_STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
  volatile float* thr_pub = reinterpret_cast<volatile float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [GROUPS][BQ]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_qh); tma_prefetch_desc(&map_ql);
    tma_prefetch_desc(&map_kh); tma_prefetch_desc(&map_kl);
    mbar_init(q_full, 1);
    for (int i = 0; i < KSTAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2u * p.kblocks * QTILE_BYTES);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        tma_load_2d(sQ + (0 * MAX_KB + kb) * QTILE_BYTES, &map_qh, q_full, kb * BK, q0);
        tma_load_2d(sQ + (1 * MAX_KB + kb) * QTILE_BYTES, &map_ql, q_full, kb * BK, q0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], 2u * p.kblocks * KTILE_BYTES);
        uint8_t* dst = sK + stage * KSTAGE_BYTES;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tma_load_2d(dst + (0 * MAX_KB + kb) * KTILE_BYTES, &map_kh, &full[stage], kb * BK, t * BNK);
          tma_load_2d(dst + (1 * MAX_KB + kb) * KTILE_BYTES, &map_kl, &full[stage], kb * BK, t * BNK);
        }
        if (++stage == KSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(0, BQ, BNK);
      mbar_wait(q_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int acc = it & (ACC_STAGES - 1);
        mbar_wait(&acc_empty[acc], ((it / ACC_STAGES) & 1) ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BNK;
        const uint32_t k_base = smem_u32(sK + stage * KSTAGE_BYTES);
        const uint32_t q_base = smem_u32(sQ);
        for (int kb = 0; kb < p.kblocks; ++kb) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t qh = umma_desc_sw128(q_base + (0 * MAX_KB + kb) * QTILE_BYTES + k * 32);
            const uint64_t ql = umma_desc_sw128(q_base + (1 * MAX_KB + kb) * QTILE_BYTES + k * 32);
            const uint64_t kh = umma_desc_sw128(k_base + (0 * MAX_KB + kb) * KTILE_BYTES + k * 32);
            const uint64_t kl = umma_desc_sw128(k_base + (1 * MAX_KB + kb) * KTILE_BYTES + k * 32);
            umma_f16(d_tmem, ql, kh, idesc, (kb | k) != 0);
            umma_f16(d_tmem, qh, kl, idesc, 1);
            umma_f16(d_tmem, qh, kh, idesc, 1);
          }
        }
        umma_commit(&empty[stage]);
        umma_commit(&acc_full[acc]);
        if (++stage == KSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int group = (warp - 2) >> 2;  // 0: even local tiles, 1: odd local tiles
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;   // query within the tile == TMEM lane
    const int qg = q0 + row;
    const float bsq = (qg < p.q) ? p.bsq[qg] : 0.f;
    const int top_k = p.top_k;
    // Nothing at or below a valid lower bound of the k-th best similarity can be in the top-k: with the bound taken from
    // the slots the previous frame selected, only a handful of candidates per query ever reach the insertion path.
    const float floor_thr = (!DENSE && p.thr_floor && qg < p.q) ? p.thr_floor[qg] : -CUDART_INF_F;
    float* lv = list_val + group * kListCap * BQ;
    int* li = list_idx + group * kListCap * BQ;
    TopK tk;
    if (!DENSE) {
      for (int j = 0; j < kListCap; ++j) {
        lv[j * BQ + row] = (j < top_k) ? -CUDART_INF_F : CUDART_INF_F;
        li[j * BQ + row] = -1;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        tk.gmin[i] = (i * 8 < top_k) ? -CUDART_INF_F : CUDART_INF_F;
        tk.gpos[i] = i * 8;
      }
      tk.refresh();
      // The two groups scan alternating tiles of the SAME queries into separate lists.  Each publishes its running k-th
      // best; the other group's value is a valid lower bound of the query's global k-th best (its list holds k real
      // candidates at least that large), so gating on it keeps every global top-k member while cutting the insertions
      // of two half-length lists to about those of one full-length list.  Stale reads are fine: the bound only grows.
      thr_pub[group * BQ + row] = -CUDART_INF_F;
      named_bar_sync(3, GROUPS * 128);
    }
    volatile const float* thr_other = thr_pub + (group ^ 1) * BQ + row;
    for (int it = group; t_begin + it < t_end; it += GROUPS) {
      const int acc = it & (ACC_STAGES - 1);
      const int n0 = (t_begin + it) * BNK;
      if (row < BNK) {
        const int n = n0 + row;
        ns[acc * BNK + row] = (n >= p.n_lead && n < p.n_window) ? p.neg_s[n] : -CUDART_INF_F;
      }
      named_bar_sync(1 + group, 128);
      mbar_wait(&acc_full[acc], (it / ACC_STAGES) & 1);
      tc_fence_after();
      const float4* ns4 = reinterpret_cast<const float4*>(ns + acc * BNK);
#pragma unroll 1
      for (int c = 0; c < BNK / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc * BNK + c * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 s4 = ns4[c * 8 + j];
          v[4 * j + 0] = fmaf(s4.x, bsq, __uint_as_float(r[4 * j + 0]));
          v[4 * j + 1] = fmaf(s4.y, bsq, __uint_as_float(r[4 * j + 1]));
          v[4 * j + 2] = fmaf(s4.z, bsq, __uint_as_float(r[4 * j + 2]));
          v[4 * j + 3] = fmaf(s4.w, bsq, __uint_as_float(r[4 * j + 3]));
        }
        if (DENSE) {
          if (qg < p.q) {
            float* dst = p.dense_out + (long long)qg * p.ld_dense + n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c * 32 + j < p.n_window) dst[j] = (v[j] == v[j]) ? v[j] : -CUDART_INF_F;
          }
        } else {
          // Bit j of `pending` = column j beats this query's current k-th best.  Lanes then retire their
          // pending candidates in lock-step (1st of every lane, 2nd of every lane, ...): the warp pays for
          // max-per-lane insertions per chunk, not for every distinct column position.
          uint32_t pending = 0;
          const float gate = fmaxf(fmaxf(tk.thr, floor_thr), *thr_other);
#pragma unroll
          for (int j = 0; j < 32; ++j) pending |= (v[j] > gate) ? (1u << j) : 0u;
          while (__any_sync(0xffffffffu, pending != 0)) {
            if (pending) {
              const int j = __ffs(pending) - 1;
              pending &= pending - 1;
              const float x = pick32(v, j);
              if (x > tk.thr) tk.insert(lv, li, row, x, n0 + c * 32 + j);
            }
          }
          thr_pub[group * BQ + row] = tk.thr;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
    if (!DENSE) {
      const long long base = ((long long)blockIdx.y * GROUPS + group) * kListCap;
      for (int j = 0; j < top_k; ++j) {
        p.part_val[(base + j) * p.qpad + qg] = lv[j * BQ + row];
        p.part_idx[(base + j) * p.qpad + qg] = li[j * BQ + row];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// One warp per query: global top-k over the per-split candidate lists, softmax, usage, dense scatter.
constexpr int MERGE_WARPS = 8;
__global__ void __launch_bounds__(MERGE_WARPS * 32)
merge_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int nsplit, int top_k, int q,
             int qpad, int* __restrict__ out_idx, float* __restrict__ out_w, __half* __restrict__ P, long long ldP,
             float* __restrict__ use_cnt, int n_long, int add_long, int add_work, float* __restrict__ out_sim) {
  // row pitch = 4 mod 32: the staging writes of a warp (4 candidates x 8 queries) and the per-warp scans are conflict-free
  __shared__ float cv[MERGE_WARPS][kMaxSplit * GROUPS * kListCap + 4];
  __shared__ int ci[MERGE_WARPS][kMaxSplit * GROUPS * kListCap + 4];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x * MERGE_WARPS;
  const int C = nsplit * top_k;
  // cooperative, sector-coalesced staging of the [C candidates] x [8 queries] tile (rows are qpad apart)
  for (int i = threadIdx.x; i < C * MERGE_WARPS; i += MERGE_WARPS * 32) {
    const int c = i / MERGE_WARPS, qo = i - c * MERGE_WARPS;
    const int s_ = c / top_k, j = c - s_ * top_k;
    const long long off = ((long long)s_ * kListCap + j) * qpad + qb + qo;
    const bool ok = qb + qo < q;
    cv[qo][c] = ok ? part_val[off] : -CUDART_INF_F;
    ci[qo][c] = ok ? part_idx[off] : -1;
  }
  __syncthreads();
  const int qi = qb + w;
  if (qi >= q) return;
  float sel_v = -CUDART_INF_F;
  int sel_i = 0;
  for (int round = 0; round < top_k; ++round) {
    float bv = -CUDART_INF_F;
    int bi = 0x7fffffff, bc = -1;
    for (int c = lane; c < C; c += 32) {
      const float x = cv[w][c];
      const int xi = ci[w][c];
      if (xi >= 0 && (x > bv || (x == bv && xi < bi))) { bv = x; bi = xi; bc = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bc = oc; }
    }
    if (bc >= 0 && (bc & 31) == lane) ci[w][bc] = -1;  // consumed
    if (lane == round) { sel_v = bv; sel_i = (bc >= 0) ? bi : 0; }
    __syncwarp();
  }
  const float vmax = __shfl_sync(0xffffffffu, sel_v, 0);
  float e = (lane < top_k && sel_v > -CUDART_INF_F) ? expf(sel_v - vmax) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float wgt = e / sum;
  out_idx[(long long)qi * kListCap + lane] = (lane < top_k) ? sel_i : 0;
  out_w[(long long)qi * kListCap + lane] = (lane < top_k) ? wgt : 0.f;
  if (out_sim) out_sim[(long long)qi * kListCap + lane] = (lane < top_k) ? sel_v : -CUDART_INF_F;
  if (lane < top_k) {
    if (P) P[(long long)qi * ldP + sel_i] = __float2half_rn(wgt);
    if (use_cnt && ((sel_i < n_long) ? add_long : add_work)) atomicAdd(use_cnt + sel_i, wgt);
  }
}

// Full softmax over the window for every query row (consolidation: memory_utils.py:66-71), plus the
// shrinkage read-out  sum_n p[n] * shr[n]  (memory_manager.py:273).  One block per query.
__global__ void __launch_bounds__(256)
row_softmax_kernel(const float* __restrict__ sim, long long ld_sim, int n_window, const float* __restrict__ shr,
                   __half* __restrict__ P, long long ldP, float* __restrict__ shr_out) {
  __shared__ float red[8];
  const int qi = blockIdx.x;
  const float* row = sim + (long long)qi * ld_sim;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float m = -CUDART_INF_F;
  for (int n = threadIdx.x; n < n_window; n += 256) m = fmaxf(m, row[n]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[w] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int n = threadIdx.x; n < n_window; n += 256) s += expf(row[n] - m);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[w] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < 8; ++i) s += red[i];
  __syncthreads();
  const float inv = 1.f / s;
  float acc = 0.f;
  for (int n = threadIdx.x; n < n_window; n += 256) {
    const float pn = expf(row[n] - m) * inv;
    P[(long long)qi * ldP + n] = __float2half_rn(pn);
    if (shr) acc += pn * shr[n];
  }
  if (shr_out) {
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += red[i];
      shr_out[qi] = t;
    }
  }
}

__global__ void tick_kernel(float* __restrict__ life, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) life[i] += 1.f;
}

}  // namespace simtopk

namespace simtopk {
// Temporal warm start: thr[q] = min over the slots the PREVIOUS read selected for query q of their similarity to the
// CURRENT query, minus a margin.  The k-th best of the whole window is >= the k-th best of any k of its slots, so this is
// a valid lower bound; consecutive frames share most of their top-k, so it is tight.  One warp per query, one lane per
// slot; operands are the hi parts of the packed rows the GEMM uses, evaluated in fp32 with a rigorous allowance for the
// dropped low-order parts.  Any invalid slot -> -inf (no bound).
__global__ void __launch_bounds__(256)
thr_floor_kernel(const int* __restrict__ prev_idx, int top_k, int q, int n_lead, int n_window, int kdim,
                 const __half* __restrict__ k_hi, const float* __restrict__ neg_s, const __half* __restrict__ q_hi,
                 const float* __restrict__ bsq, float* __restrict__ thr) {
  __shared__ float qrow[8][128];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * 8 + w;
  if (qi >= q) return;
  // hi parts only (half the L2 traffic): sum (qh+ql)(kh+kl) = sum qh.kh + R with |R| <= sum |qh.kh| * 2^-9.9 since
  // |ql| <= 2^-11 |qh| and |kl| <= 2^-11 |kh|; the bound is lowered by that much, which keeps it valid.
  for (int c = lane; c < kdim; c += 32) qrow[w][c] = __half2float(q_hi[(long long)qi * kdim + c]);
  __syncwarp();
  float sim = CUDART_INF_F;
  bool bad = false;
  if (lane < top_k) {
    const int n = prev_idx[(long long)qi * kListCap + lane];
    if (n < n_lead || n >= n_window) {
      bad = true;
    } else {
      const uint4* ph = reinterpret_cast<const uint4*>(k_hi + (long long)n * kdim);
      float acc = 0.f, mag = 0.f;
      for (int c8 = 0; c8 < kdim / 8; ++c8) {
        const uint4 vh = ph[c8];
        const __half2* h2 = reinterpret_cast<const __half2*>(&vh);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __half22float2(h2[e]);
          const float t0 = a.x * qrow[w][c8 * 8 + 2 * e], t1 = a.y * qrow[w][c8 * 8 + 2 * e + 1];
          acc += t0 + t1;
          mag += fabsf(t0) + fabsf(t1);
        }
      }
      sim = fmaf(neg_s[n], bsq[qi], acc) - mag * (1.f / 512.f);
    }
  }
  bad = __any_sync(0xffffffffu, bad);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sim = fminf(sim, __shfl_xor_sync(0xffffffffu, sim, o));
  if (lane == 0) thr[qi] = (bad || !(sim == sim)) ? -CUDART_INF_F : sim - 1e-3f * (1.f + fabsf(sim));
}
}  // namespace simtopk

static int make_maps(CUtensorMap* m, const __half* q_hi, const __half* q_lo, int q, const __half* k_hi,
                     const __half* k_lo, int n_window, int ck) {
  const char* err = nullptr;
  const uint64_t kdim = 2ull * ck;
  if (make_tmap_2d(&m[0], TmapType::F16, q_hi, kdim, q, kdim * 2, 64, 128, &err) ||
      make_tmap_2d(&m[1], TmapType::F16, q_lo, kdim, q, kdim * 2, 64, 128, &err) ||
      make_tmap_2d(&m[2], TmapType::F16, k_hi, kdim, n_window, kdim * 2, 64, 64, &err) ||
      make_tmap_2d(&m[3], TmapType::F16, k_lo, kdim, n_window, kdim * 2, 64, 64, &err)) {
    set_error("simtopk: %s", err ? err : "tensor map");
    return 3;
  }
  return 0;
}


// Split of the memory axis over CTAs: minimise (waves) x (tiles per CTA), with a small per-split
// charge for the list warm-up every split repeats.
static int choose_split(int q_tiles, int tiles, int max_split) {
  const int sms = sm_count();
  int best = 1;
  long long best_cost = -1;
  for (int ns = 1; ns <= max_split && ns <= tiles; ++ns) {
    const long long waves = ceil_div((long long)q_tiles * ns, sms);
    const long long cost = waves * (ceil_div(tiles, ns) * 4 + 3);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ns; }
  }
  return best;
}

size_t simtopk_workspace_bytes(int q) {
  const size_t qpad = (size_t)ceil_div(q, 128) * 128;
  return (size_t)kMaxSplit * simtopk::GROUPS * kListCap * qpad * 8;
}

int launch_sim_topk(const __half* k_hi, const __half* k_lo, const float* neg_s, int n_window, int n_lead,
                    const __half* q_hi, const __half* q_lo, const float* bsq, int q, int ck, int top_k,
                    void* workspace, int* out_idx, float* out_w, __half* P, long long ldP, float* use_cnt,
                    float* life_cnt, int n_long, int count_long, int count_work, float* out_sim, const int* prev_idx,
                    float* thr_ws, cudaStream_t stream) {
  using namespace simtopk;
  B200_REQUIRE(ck == 32 || ck == 64, "simtopk: key_dim %d unsupported (32 or 64)", ck);
  B200_REQUIRE(top_k >= 1 && top_k <= kListCap, "simtopk: top_k %d out of range [1,%d]", top_k, kListCap);
  B200_REQUIRE(n_window - n_lead >= top_k, "simtopk: top_k %d exceeds the %d memory slots (the reference raises too)",
               top_k, n_window - n_lead);
  B200_REQUIRE(q >= 1 && n_lead >= 0 && n_lead < 8, "simtopk: bad q / n_lead");
  CUtensorMap maps[4];
  if (int rc = make_maps(maps, q_hi, q_lo, q, k_hi, k_lo, n_window, ck)) return rc;
  Params p{};
  p.q = q; p.n_window = n_window; p.n_lead = n_lead; p.kblocks = 2 * ck / 64; p.top_k = top_k;
  p.tiles_total = ceil_div(n_window, BNK);
  const int q_tiles = ceil_div(q, BQ);
  int nsplit = choose_split(q_tiles, p.tiles_total, kMaxSplit);
  p.tiles_per_split = ceil_div(p.tiles_total, nsplit);
  nsplit = ceil_div(p.tiles_total, p.tiles_per_split);
  p.qpad = q_tiles * BQ;
  p.neg_s = neg_s; p.bsq = bsq;
  p.part_val = reinterpret_cast<float*>(workspace);
  p.part_idx = reinterpret_cast<int*>(p.part_val + (size_t)kMaxSplit * GROUPS * kListCap * p.qpad);
  static bool configured_dev[kMaxDevices] = {false};  // function attributes are per device
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(simtopk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    B200_CUDA(cudaFuncSetAttribute(simtopk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  if (P) B200_CUDA(cudaMemsetAsync(P, 0, (size_t)q * ldP * sizeof(__half), stream));
  if (prev_idx) {
    B200_REQUIRE(thr_ws != nullptr, "simtopk: prev_idx needs the thr_ws scratch ([q] floats)");
    thr_floor_kernel<<<ceil_div(q, 8), 256, 0, stream>>>(prev_idx, top_k, q, n_lead, n_window, 2 * ck, k_hi, neg_s, q_hi, bsq,
                                                         thr_ws);
    B200_LAUNCH_CHECK();
    p.thr_floor = thr_ws;
  }
  simtopk_kernel<false><<<dim3(q_tiles, nsplit), THREADS, SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  B200_LAUNCH_CHECK();
  merge_kernel<<<ceil_div(q, MERGE_WARPS), MERGE_WARPS * 32, 0, stream>>>(
      p.part_val, p.part_idx, nsplit * GROUPS, top_k, q, p.qpad, out_idx, out_w, P, ldP, use_cnt, n_long, count_long,
      count_work, out_sim);
  B200_LAUNCH_CHECK();
  if (life_cnt) {
    // life += 1 for every slot whose usage is being counted (kv_memory_store.py:118-125)
    const int lo = count_long ? n_lead : n_long;
    const int hi = count_work ? n_window : n_long;
    if (hi > lo) {
      tick_kernel<<<ceil_div(hi - lo, 256), 256, 0, stream>>>(life_cnt + lo, hi - lo);
      B200_LAUNCH_CHECK();
    }
  }
  return 0;
}

// Global top-k + softmax over `n_lists` candidate lists per query (layout [n_lists][kListCap][qpad], entries with
// idx < 0 are empty): the cross-rank merge of a bank-sharded read, and the second stage of launch_sim_topk.
int launch_merge_lists(const float* part_val, const int* part_idx, int n_lists, int top_k, int q, int qpad, int* out_idx,
                       float* out_w, float* out_sim, cudaStream_t stream) {
  using namespace simtopk;
  B200_REQUIRE(n_lists >= 1 && n_lists <= kMaxSplit * GROUPS, "merge: n_lists %d out of range [1,%d]", n_lists, kMaxSplit * GROUPS);
  B200_REQUIRE(top_k >= 1 && top_k <= kListCap && q >= 1 && qpad >= q, "merge: bad shape");
  merge_kernel<<<ceil_div(q, MERGE_WARPS), MERGE_WARPS * 32, 0, stream>>>(part_val, part_idx, n_lists, top_k, q, qpad, out_idx,
                                                                         out_w, nullptr, 0, nullptr, 0, 0, 0, out_sim);
  B200_LAUNCH_CHECK();
  return 0;
}

int launch_sim_dense_softmax(const __half* k_hi, const __half* k_lo, const float* neg_s, const float* shr,
                             int n_window, int n_lead, const __half* q_hi, const __half* q_lo, const float* bsq,
                             int q, int ck, float* sim_ws, long long ld_sim, __half* P, long long ldP, float* shr_out,
                             cudaStream_t stream) {
  using namespace simtopk;
  B200_REQUIRE(ck == 32 || ck == 64, "simdense: key_dim %d unsupported (32 or 64)", ck);
  B200_REQUIRE(q >= 1 && n_window > n_lead && n_lead >= 0 && n_lead < 8, "simdense: bad shape");
  CUtensorMap maps[4];
  if (int rc = make_maps(maps, q_hi, q_lo, q, k_hi, k_lo, n_window, ck)) return rc;
  Params p{};
  p.q = q; p.n_window = n_window; p.n_lead = n_lead; p.kblocks = 2 * ck / 64; p.top_k = 1;
  p.tiles_total = ceil_div(n_window, BNK);
  const int q_tiles = ceil_div(q, BQ);
  int nsplit = ceil_div(sm_count(), q_tiles);
  if (nsplit > p.tiles_total) nsplit = p.tiles_total;
  p.tiles_per_split = ceil_div(p.tiles_total, nsplit);
  nsplit = ceil_div(p.tiles_total, p.tiles_per_split);
  p.qpad = q_tiles * BQ;
  p.neg_s = neg_s; p.bsq = bsq;
  p.dense_out = sim_ws; p.ld_dense = ld_sim;
  static bool configured_dev[kMaxDevices] = {false};  // function attributes are per device
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(simtopk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  simtopk_kernel<true><<<dim3(q_tiles, nsplit), THREADS, SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  B200_LAUNCH_CHECK();
  row_softmax_kernel<<<q, 256, 0, stream>>>(sim_ws, ld_sim, n_window, shr, P, ldP, shr_out);
  B200_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200
