// Implicit-GEMM convolution on tcgen05, NHWC fp16 activations, fp32 accumulation in TMEM.
// Replaces the nn.Conv2d + BatchNorm(eval) + ReLU (+ residual add) call chains of the reference network
// (deva/model/resnet.py:46-152, big_modules.py:23-212, modules.py:22-169, group_modules.py:41-67).
//
//   D[pixel, cout] = sum_{tap, cin} X[b, y*s + dy(tap), x*s + dx(tap), cin] * W[cout, tap, cin]
//
//   M = 128 output pixels = a TH x TW rectangle of one image (TMEM lanes)
//   N = NT output channels (<= 256, TMEM columns), K = 64 input channels of one filter tap per k-iteration.
//
// A operand: the activation tensor is described to TMA as a rank-5 tiled tensor (C, W, H, B, 1); the box
//   (64, TW, TH, 1, 1) lands in shared memory as 128 rows of 128 bytes in exactly the 128B-swizzled K-major
//   layout tcgen05 wants.  A filter tap is just a shifted box origin; convolution padding is TMA's
//   out-of-bounds zero fill (also for negative coordinates).  Stride-2 convolutions use one tensor map per
//   input phase (y%2, x%2) over the same memory, so every tap is again a unit-stride box.
// B operand: weights packed [Cout_pad, taps * Cin_pad] fp16 (K-major), 2-D TMA.
// Epilogue (TMEM -> registers): + bias (folded BatchNorm), + optional residual (own or batch-broadcast),
//   + optional rank-1 term w1[cout] * x1[b, pixel] (the "+1" mask / logit input channel of
//   sensory_compress and g4_conv), then writes any of: raw fp16, ReLU'd fp16, raw fp32.
// Warp roles: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2..5 = epilogue.  Persistent CTAs, 4-stage
//   smem ring, double-buffered accumulators.
#include <cuda_fp16.h>

#include "common.h"
#include "conv.h"
#include "ptx.cuh"
#include "tmap.h"

namespace b200 {
namespace conv {

constexpr int BM = 128, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES_MAX = 256 * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int THREADS = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

struct Params {
  int batch, ho, wo, cout;
  int tiles_x, tiles_y, tw, th;
  int n_tiles, nt;
  int taps, cblocks;
  int pos_x, pos_y, pos_b;  // which TMA coordinate carries x / y / image index
  signed char tap_map[kMaxTaps], tap_dx[kMaxTaps], tap_dy[kMaxTaps];
  const float* bias;
  const __half* res;
  long long res_batch_stride;  // elements; 0 = broadcast one image over the batch
  const float* rank1_w;
  const float* rank1_x;        // [batch, ho*wo]
  __half* out_raw;
  __half* out_relu;
  float* out_f32;
};

struct Maps {
  CUtensorMap act[4];
  CUtensorMap wgt;
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, const int (&c)[5]) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
      : "memory");
}

__device__ __forceinline__ void tile_decode(int tile, const Params& p, int& b, int& y0, int& x0, int& n0) {
  const int nidx = tile % p.n_tiles;
  int m = tile / p.n_tiles;
  const int tx = m % p.tiles_x;
  m /= p.tiles_x;
  const int ty = m % p.tiles_y;
  b = m / p.tiles_y;
  y0 = ty * p.th;
  x0 = tx * p.tw;
  n0 = nidx * p.nt;
}

__global__ void __launch_bounds__(THREADS, 1)
conv_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.batch * p.tiles_y * p.tiles_x * p.n_tiles;
  const int k_iters = p.taps * p.cblocks;
  const uint32_t stage_tx = A_BYTES + p.nt * BK * 2;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.act[i]);
    tma_prefetch_desc(&maps.wgt);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
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
        int b, y0, x0, n0;
        tile_decode(tile, p, b, y0, x0, n0);
        for (int t = 0; t < p.taps; ++t) {
          int c[5] = {0, 0, 0, 0, 0};
          c[p.pos_x] = x0 + p.tap_dx[t];
          c[p.pos_y] = y0 + p.tap_dy[t];
          c[p.pos_b] = b;
          const CUtensorMap* am = &maps.act[p.tap_map[t]];
          for (int cb = 0; cb < p.cblocks; ++cb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], stage_tx);
            c[0] = cb * BK;
            tma_load_5d(sA + stage * A_BYTES, am, &full[stage], c);
            tma_load_2d(sB + stage * B_BYTES_MAX, &maps.wgt, &full[stage], (t * p.cblocks + cb) * BK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(0, BM, p.nt);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * B_BYTES_MAX);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                     (ki | k) != 0);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // pixel within the tile
    const int ty = row / p.tw, tx = row - ty * p.tw;
    const bool vec_ok = (p.cout % 8) == 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int b, y0, x0, n0;
      tile_decode(tile, p, b, y0, x0, n0);
      const int y = y0 + ty, x = x0 + tx;
      const bool live = (y < p.ho) && (x < p.wo);
      const long long pix = ((long long)b * p.ho + y) * p.wo + x;
      const long long off = pix * p.cout + n0;
      const __half* res = p.res ? p.res + (p.res_batch_stride ? off : ((long long)y * p.wo + x) * p.cout + n0)
                                : nullptr;
      const float r1x = (p.rank1_x && live) ? p.rank1_x[pix] : 0.f;
      const int acc = it & 1;
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < p.nt / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc * 256 + c * 32, r);
        tmem_ld_wait();
        const int ch0 = n0 + c * 32;
        if (live && ch0 < p.cout) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (ch0 + 32 <= p.cout && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(p.bias + ch0 + j);
              v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
            }
            if (p.rank1_w) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 wv = *reinterpret_cast<const float4*>(p.rank1_w + ch0 + j);
                v[j] = fmaf(wv.x, r1x, v[j]); v[j + 1] = fmaf(wv.y, r1x, v[j + 1]);
                v[j + 2] = fmaf(wv.z, r1x, v[j + 2]); v[j + 3] = fmaf(wv.w, r1x, v[j + 3]);
              }
            }
            if (res) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 rv = *reinterpret_cast<const uint4*>(res + c * 32 + j);
                const __half2* h2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __half22float2(h2[e]);
                  v[j + 2 * e] += f.x;
                  v[j + 2 * e + 1] += f.y;
                }
              }
            }
            if (p.out_raw) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 o;
                __half2* h2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(v[j + 2 * e], v[j + 2 * e + 1]);
                *reinterpret_cast<uint4*>(p.out_raw + off + c * 32 + j) = o;
              }
            }
            if (p.out_relu) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 o;
                __half2* h2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  h2[e] = __floats2half2_rn(fmaxf(v[j + 2 * e], 0.f), fmaxf(v[j + 2 * e + 1], 0.f));
                *reinterpret_cast<uint4*>(p.out_relu + off + c * 32 + j) = o;
              }
            }
            if (p.out_f32) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(p.out_f32 + off + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {  // ragged channel tail (e.g. Cout = 1, 129): scalar path
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int ch = ch0 + j;
              if (ch < p.cout) {
                float o = v[j] + p.bias[ch];
                if (p.rank1_w) o = fmaf(p.rank1_w[ch], r1x, o);
                if (res) o += __half2float(res[c * 32 + j]);
                if (p.out_raw) p.out_raw[off + c * 32 + j] = __float2half_rn(o);
                if (p.out_relu) p.out_relu[off + c * 32 + j] = __float2half_rn(fmaxf(o, 0.f));
                if (p.out_f32) p.out_f32[off + c * 32 + j] = o;
              }
            }
          }
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

}  // namespace conv

int launch_conv(const ConvDesc& d, cudaStream_t stream) {
  using namespace conv;
  B200_REQUIRE(d.cin_pad % 64 == 0 && d.cin_pad >= 64, "conv: cin_pad %d must be a multiple of 64", d.cin_pad);
  B200_REQUIRE(d.stride == 1 || d.stride == 2, "conv: stride %d unsupported", d.stride);
  B200_REQUIRE(d.kh == d.kw && (d.kh == 1 || d.kh == 3), "conv: %dx%d filter unsupported here", d.kh, d.kw);
  B200_REQUIRE(d.nt % 32 == 0 && d.nt >= 32 && d.nt <= 256 && d.cout_pad % d.nt == 0, "conv: bad channel tile %d", d.nt);
  B200_REQUIRE(d.tw * d.th == 128 && d.tw <= 256 && d.th <= 256, "conv: spatial tile %dx%d must cover 128 pixels", d.th, d.tw);
  const int pad = d.kh / 2;
  const int ho = (d.h + 2 * pad - d.kh) / d.stride + 1;
  const int wo = (d.w + 2 * pad - d.kw) / d.stride + 1;
  Maps maps;
  Params p{};
  const char* err = nullptr;
  const __half* x = reinterpret_cast<const __half*>(d.x);
  const int ktaps = d.kh * d.kw;
  const int sources = d.x2 ? 2 : 1;
  p.taps = ktaps * sources;
  B200_REQUIRE(p.taps <= kMaxTaps, "conv: too many taps");
  B200_REQUIRE(sources == 1 || d.stride == 1, "conv: the two-input form is stride-1 only");
  if (d.stride == 1) {
    if (make_tmap_act5(&maps.act[0], x, d.cin_pad, d.w, d.h, d.batch, (long long)d.cin_pad, (long long)d.w * d.cin_pad,
                       (long long)d.h * d.w * d.cin_pad, d.tw, d.th, &err)) {
      set_error("conv: %s", err ? err : "tensor map");
      return 3;
    }
    maps.act[1] = maps.act[2] = maps.act[3] = maps.act[0];
    if (d.x2 && make_tmap_act5(&maps.act[1], d.x2, d.cin_pad, d.w, d.h, d.batch, (long long)d.cin_pad,
                               (long long)d.w * d.cin_pad, (long long)d.h * d.w * d.cin_pad, d.tw, d.th, &err)) {
      set_error("conv: %s", err ? err : "tensor map (x2)");
      return 3;
    }
    for (int t = 0; t < p.taps; ++t) {
      const int kt = t % ktaps;
      p.tap_map[t] = (signed char)(t / ktaps);
      p.tap_dy[t] = (signed char)(kt / d.kw - pad);
      p.tap_dx[t] = (signed char)(kt % d.kw - pad);
    }
  } else {
    // phase (py, px): rows y = 2*i + py, cols x = 2*j + px of the same NHWC buffer
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const int hp = (d.h - py + 1) / 2, wp = (d.w - px + 1) / 2;
        const __half* base = x + ((long long)py * d.w + px) * d.cin_pad;
        if (hp <= 0 || wp <= 0) { maps.act[py * 2 + px] = maps.act[0]; continue; }
        if (make_tmap_act5(&maps.act[py * 2 + px], base, d.cin_pad, wp, hp, d.batch, 2ll * d.cin_pad,
                           2ll * d.w * d.cin_pad, (long long)d.h * d.w * d.cin_pad, d.tw, d.th, &err)) {
          set_error("conv: %s", err ? err : "tensor map");
          return 3;
        }
      }
    for (int t = 0; t < p.taps; ++t) {
      const int oy = t / d.kw - pad, ox = t % d.kw - pad;  // input offset relative to 2*yo, 2*xo
      const int py = ((oy % 2) + 2) % 2, px = ((ox % 2) + 2) % 2;
      p.tap_map[t] = (signed char)(py * 2 + px);
      p.tap_dy[t] = (signed char)((oy - py) / 2);
      p.tap_dx[t] = (signed char)((ox - px) / 2);
    }
  }
  if (make_tmap_2d(&maps.wgt, TmapType::F16, d.w_packed, (uint64_t)p.taps * d.cin_pad, d.cout_pad,
                   (uint64_t)p.taps * d.cin_pad * 2, 64, d.nt, &err)) {
    set_error("conv: %s", err ? err : "weight tensor map");
    return 3;
  }
  p.batch = d.batch; p.ho = ho; p.wo = wo; p.cout = d.cout;
  p.tw = d.tw; p.th = d.th;
  p.tiles_x = ceil_div(wo, d.tw); p.tiles_y = ceil_div(ho, d.th);
  p.nt = d.nt; p.n_tiles = d.cout_pad / d.nt;
  p.cblocks = d.cin_pad / 64;
  p.pos_x = 1; p.pos_y = 2; p.pos_b = 3;
  p.bias = d.bias;
  p.res = reinterpret_cast<const __half*>(d.res);
  p.res_batch_stride = d.res_broadcast ? 0 : (long long)ho * wo * d.cout;
  p.rank1_w = d.rank1_w; p.rank1_x = d.rank1_x;
  p.out_raw = reinterpret_cast<__half*>(d.out_raw);
  p.out_relu = reinterpret_cast<__half*>(d.out_relu);
  p.out_f32 = d.out_f32;
  static bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const long long total = (long long)p.batch * p.tiles_y * p.tiles_x * p.n_tiles;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  conv_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(maps, p);
  B200_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200
