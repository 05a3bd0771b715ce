// Bandwidth-bound kernels of the denoiser step (NHWC activations, channel-contiguous rows).
// Reference anchors are given per kernel; paths are relative to the reference tree.
#pragma once

#include "common.cuh"

namespace hd {

// Per-step state living in device memory so that a captured CUDA graph can be replayed for every
// timestep: kernels read the current step from here instead of from launch parameters.
struct StepState {
  int step;  // index into the time-modulation table and the coefficient array
};

struct StepCoef {  // mirrors hd_step_coef
  float timestep, sqrt_beta_prod, sqrt_alpha_prod, clip, k_x0, k_eps, k_x, k_noise;
};

// Where a kernel finds the AdaLN vectors of the face it is working on.
struct ModRef {
  const float* table;    // [rows][stride]
  const int* row_idx;    // per-face table row (hd_denoise_step: face -> its t; sampler: all = current step)
  int stride;            // floats per table row (124928 for the 32 blocks)
  __device__ __forceinline__ const float* row(int face) const {
    return table + static_cast<size_t>(row_idx[face]) * stride;
  }
};

// ------------------------------------------------------------------------------------------------
// intro: 3x3 conv 4 -> 128, NCHW fp32 latents -> NHWC fp32 residual stream (model.py:159-167,235)
// one block per face; thread = (pixel column, 8 output channels); weights [36][128] in smem
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) intro_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         int S) {
  extern __shared__ float s_intro[];
  float* s_w = s_intro;                 // [36][128]  (k = c*9 + ky*3 + kx major, channel minor)
  float* s_in = s_intro + 36 * 128;     // [4][S+2][S+2] zero-padded face
  const int b = blockIdx.x;
  const int W2 = S + 2;
  pdl_trigger();
  // w is already [36][128] (host-side transpose): 4608 floats = 1152 float4, all loads of a thread in flight
  {
    float4 t[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int idx = i * 256 + threadIdx.x;
      if (idx < 36 * 32) t[i] = __ldg(reinterpret_cast<const float4*>(w) + idx);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int idx = i * 256 + threadIdx.x;
      if (idx < 36 * 32) reinterpret_cast<float4*>(s_w)[idx] = t[i];
    }
  }
  pdl_wait();
  for (int i = threadIdx.x; i < 4 * W2 * W2; i += blockDim.x) {
    const int c = i / (W2 * W2), r = (i / W2) % W2, col = i % W2;
    const int hh = r - 1, ww = col - 1;
    float v = 0.f;
    if (hh >= 0 && hh < S && ww >= 0 && ww < S) v = x[((static_cast<size_t>(b) * 4 + c) * S + hh) * S + ww];
    s_in[i] = v;
  }
  __syncthreads();
  const int cg = threadIdx.x & 15;        // 8 output channels each
  const int pl = threadIdx.x >> 4;        // 16 pixel lanes
  float bo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bo[i] = bias[cg * 8 + i];
  // register tile: 4 pixels x 8 channels per pass; the channel loop is kept rolled so the weight
  // loads stay inside it (fully unrolled, the compiler hoists all 288 weights and spills)
  for (int p0 = pl; p0 < S * S; p0 += 64) {
    float acc[4][8];
    int py[4], px[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = min(p0 + 16 * j, S * S - 1);
      py[j] = p / S;
      px[j] = p - py[j] * S;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[j][i] = bo[i];
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float* wp = s_w + (c * 9 + t) * 128 + cg * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wp);
        const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v = s_in[(c * W2 + py[j] + t / 3) * W2 + px[j] + t % 3];
          acc[j][0] = fmaf(v, w0.x, acc[j][0]); acc[j][1] = fmaf(v, w0.y, acc[j][1]);
          acc[j][2] = fmaf(v, w0.z, acc[j][2]); acc[j][3] = fmaf(v, w0.w, acc[j][3]);
          acc[j][4] = fmaf(v, w1.x, acc[j][4]); acc[j][5] = fmaf(v, w1.y, acc[j][5]);
          acc[j][6] = fmaf(v, w1.z, acc[j][6]); acc[j][7] = fmaf(v, w1.w, acc[j][7]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = p0 + 16 * j;
      if (p < S * S) store8(out + (static_cast<size_t>(b) * S * S + p) * 128 + cg * 8, acc[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm2d + AdaLN modulation (utils.py:16-24, conditional_naf.py:114-115,126-127)
//   y = (x - mu) / sqrt(var + eps) ; y = w*y + b ; out = y*(scale+1) + shift
// LPR = min(32, C/16) lanes per pixel row (>= 4 independent float4 loads per lane), 32/LPR rows per
// warp, 4 warps per block; fp32 residual in, T out.  C in {128,...,2048}
// ------------------------------------------------------------------------------------------------
// SPLIT3 (TOut = bf16): the row goes out as [hi | lo | hi] bf16 column blocks of width C (row stride 3C), the A
// operand of the split-precision tensor-core GEMM (cr_split3_kernel's mode 0 without the fp32 round trip).
template <int C, typename TOut, bool SPLIT3 = false>
__global__ void __launch_bounds__(128) ln_mod_kernel(const float* __restrict__ x, const float* __restrict__ lw,
                                                     const float* __restrict__ lb, TOut* __restrict__ out, int rows,
                                                     int rows_per_face, ModRef mod, int shift_off, int scale_off,
                                                     int has_mod) {
  constexpr int LPR = (C / 16) < 32 ? (C / 16) : 32;  // lanes per row
  constexpr int RPW = 32 / LPR;                        // rows per warp
  constexpr int NV = C / (4 * LPR);                    // float4 per lane
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
  const bool ok = row < rows;
  const int rowc = ok ? row : 0;
  pdl_trigger();
  // every load is issued before the reductions: one memory latency deep
  float4 w4[NV], b4[NV], sc[NV], sh[NV], xv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * LPR + sl) * 4;
    w4[i] = __ldg(reinterpret_cast<const float4*>(lw + c0));
    b4[i] = __ldg(reinterpret_cast<const float4*>(lb + c0));
  }
  pdl_wait();
  const float* mrow = has_mod ? mod.row(rowc / rows_per_face) : nullptr;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * LPR + sl) * 4;
    xv[i] = *reinterpret_cast<const float4*>(x + static_cast<size_t>(rowc) * C + c0);
    sc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    sh[i] = sc[i];
    if (has_mod) {
      sc[i] = __ldg(reinterpret_cast<const float4*>(mrow + scale_off + c0));
      sh[i] = __ldg(reinterpret_cast<const float4*>(mrow + shift_off + c0));
    }
    s += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s * (1.f / C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float d0 = xv[i].x - mu, d1 = xv[i].y - mu, d2 = xv[i].z - mu, d3 = xv[i].w - mu;
    ss = fmaf(d0, d0, ss); ss = fmaf(d1, d1, ss); ss = fmaf(d2, d2, ss); ss = fmaf(d3, d3, ss);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (!ok) return;
  const float denom = sqrtf(ss * (1.f / C) + 1e-6f);
  TOut* orow = out + static_cast<size_t>(row) * (SPLIT3 ? 3 * C : C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * LPR + sl) * 4;
    float y[4];
    y[0] = w4[i].x * ((xv[i].x - mu) / denom) + b4[i].x;
    y[1] = w4[i].y * ((xv[i].y - mu) / denom) + b4[i].y;
    y[2] = w4[i].z * ((xv[i].z - mu) / denom) + b4[i].z;
    y[3] = w4[i].w * ((xv[i].w - mu) / denom) + b4[i].w;
    if (has_mod) {
      y[0] = y[0] * (sc[i].x + 1.f) + sh[i].x;
      y[1] = y[1] * (sc[i].y + 1.f) + sh[i].y;
      y[2] = y[2] * (sc[i].z + 1.f) + sh[i].z;
      y[3] = y[3] * (sc[i].w + 1.f) + sh[i].w;
    }
    if (SPLIT3) {
      float lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) lo[e] = y[e] - __bfloat162float(__float2bfloat16_rn(y[e]));
      const uint2 ph = make_uint2(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]));
      const uint2 pl = make_uint2(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]));
      bf16* o = reinterpret_cast<bf16*>(orow) + c0;
      *reinterpret_cast<uint2*>(o) = ph;
      *reinterpret_cast<uint2*>(o + C) = pl;
      *reinterpret_cast<uint2*>(o + 2 * C) = ph;
    } else if (sizeof(TOut) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(orow) + c0) = make_float4(y[0], y[1], y[2], y[3]);
    } else {
      uint2 p;
      p.x = pack_bf16x2(y[0], y[1]);
      p.y = pack_bf16x2(y[2], y[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(orow) + c0) = p;
    }
  }
}

// Wide rows, few of them (1024 / 2048 channels at the 2x2 / 1x1 levels): one 256-thread block per pixel
// row, every load (x, LayerNorm affine, modulation) issued before the block reduction so the kernel
// is one memory latency deep instead of three.
template <int C, typename TOut>
__global__ void __launch_bounds__(256) ln_mod_wide_kernel(const float* __restrict__ x, const float* __restrict__ lw,
                                                          const float* __restrict__ lb, TOut* __restrict__ out, int rows,
                                                          int rows_per_face, ModRef mod, int shift_off, int scale_off,
                                                          int has_mod) {
  constexpr int NV = C / 1024;  // float4 per thread
  __shared__ float s_red[2][8];
  const int row = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, wp = t >> 5;
  pdl_trigger();
  float4 w4[NV], b4[NV], sc[NV], sh[NV], xv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {  // constants first (independent of the predecessor)
    const int c0 = (i * 256 + t) * 4;
    w4[i] = __ldg(reinterpret_cast<const float4*>(lw + c0));
    b4[i] = __ldg(reinterpret_cast<const float4*>(lb + c0));
  }
  pdl_wait();
  const float* mrow = has_mod ? mod.row(row / rows_per_face) : nullptr;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * 256 + t) * 4;
    xv[i] = *reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * C + c0);
    sc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    sh[i] = sc[i];
    if (has_mod) {
      sc[i] = __ldg(reinterpret_cast<const float4*>(mrow + scale_off + c0));
      sh[i] = __ldg(reinterpret_cast<const float4*>(mrow + shift_off + c0));
    }
    s += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
  }
  s = warp_sum(s);
  if (lane == 0) s_red[0][wp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += s_red[0][i];
  const float mu = tot * (1.f / C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float d0 = xv[i].x - mu, d1 = xv[i].y - mu, d2 = xv[i].z - mu, d3 = xv[i].w - mu;
    ss = fmaf(d0, d0, ss); ss = fmaf(d1, d1, ss); ss = fmaf(d2, d2, ss); ss = fmaf(d3, d3, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) s_red[1][wp] = ss;
  __syncthreads();
  float vt = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) vt += s_red[1][i];
  const float denom = sqrtf(vt * (1.f / C) + 1e-6f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * 256 + t) * 4;
    float y[4];
    y[0] = w4[i].x * ((xv[i].x - mu) / denom) + b4[i].x;
    y[1] = w4[i].y * ((xv[i].y - mu) / denom) + b4[i].y;
    y[2] = w4[i].z * ((xv[i].z - mu) / denom) + b4[i].z;
    y[3] = w4[i].w * ((xv[i].w - mu) / denom) + b4[i].w;
    if (has_mod) {
      y[0] = y[0] * (sc[i].x + 1.f) + sh[i].x;
      y[1] = y[1] * (sc[i].y + 1.f) + sh[i].y;
      y[2] = y[2] * (sc[i].z + 1.f) + sh[i].z;
      y[3] = y[3] * (sc[i].w + 1.f) + sh[i].w;
    }
    TOut* orow = out + static_cast<size_t>(row) * C;
    if (sizeof(TOut) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(orow) + c0) = make_float4(y[0], y[1], y[2], y[3]);
    } else {
      uint2 p;
      p.x = pack_bf16x2(y[0], y[1]);
      p.y = pack_bf16x2(y[2], y[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(orow) + c0) = p;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 (pad 1, bias) + SimpleGate + global average pool (conditional_naf.py:117-119, 54-65)
//   in  h [B, sp, sp, 2c]   (x1 = channels [0,c), x2 = channels [c,2c))
//   out g [B, sp, sp, c] = dw(x1) * dw(x2) ;  pooled[B, c] = mean_hw(g)
// One block = 256 pixels (256/sp^2 whole faces) x 64 gate channels, staged in shared memory with
// independent 16-byte loads.  A thread owns 2 gate channels (2 x1 + 2 x2 inputs) with all 36 filter
// taps in registers and walks image columns top to bottom with a 3x3 register window, so each new
// output pixel costs 3 shared-memory loads per half instead of 9 (the kernel is instruction-bound,
// not bandwidth-bound).  Column sums go through shared memory and are added per face in fixed
// order (deterministic).  tile_px = 256 at 16x16 (one face), 64 below (more blocks).
// grid (c/64, ceil(B*sp^2/tile_px)), block 256, dynamic smem tile_px*128*sizeof(T)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ld_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld_pair(const bf16* p) { return unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ void st_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st_pair(bf16* p, float a, float b) { *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(a, b); }

template <typename T>
__global__ void __launch_bounds__(256, 2) dwconv_gate_pool_kernel(const T* __restrict__ h, const float* __restrict__ w9,
                                                                  const float* __restrict__ bias, T* __restrict__ g,
                                                                  T* __restrict__ pooled, int sp, int c, int total_px,
                                                                  int tile_px) {
  extern __shared__ __align__(16) uint8_t s_dw_raw[];
  T* tile = reinterpret_cast<T*>(s_dw_raw);             // [tile_px][128 ch]: 64 x1 | 64 x2
  float* colsum = reinterpret_cast<float*>(s_dw_raw);   // aliases the tile after the stencil: [tile_px/sp columns][64]
  const int j0 = blockIdx.x * 64;
  const int px0 = blockIdx.y * tile_px;
  const int C2 = 2 * c;
  const int npix = sp * sp;
  const int cl = threadIdx.x & 31;   // channel lane: gate channels j0 + 2*cl, +1
  const int pl = threadIdx.x >> 5;   // column lane (8)
  pdl_trigger();
  // filter taps and biases of this thread's 2+2 channels (constants: before the dependency wait)
  float w1[9][2], w2[9][2], b1[2], b2[2];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float2 a = *reinterpret_cast<const float2*>(w9 + t * C2 + j0 + 2 * cl);
    const float2 b = *reinterpret_cast<const float2*>(w9 + t * C2 + c + j0 + 2 * cl);
    w1[t][0] = a.x; w1[t][1] = a.y; w2[t][0] = b.x; w2[t][1] = b.y;
  }
  {
    const float2 a = *reinterpret_cast<const float2*>(bias + j0 + 2 * cl);
    const float2 b = *reinterpret_cast<const float2*>(bias + c + j0 + 2 * cl);
    b1[0] = a.x; b1[1] = a.y; b2[0] = b.x; b2[1] = b.y;
  }
  pdl_wait();
  // stage: 256 px x 16 chunks of 16 bytes (bf16: 8 ch; fp32: two 16-byte halves of 8 ch)
  {
    const int chunk = threadIdx.x & 15;
    const int half = chunk >> 3, cc = (chunk & 7) * 8;
#pragma unroll 4
    for (int r = threadIdx.x >> 4; r < tile_px; r += 16) {
      const int p = px0 + r;
      uint4 v = make_uint4(0, 0, 0, 0);
      uint4 v2 = make_uint4(0, 0, 0, 0);
      if (p < total_px) {
        const T* src = h + static_cast<size_t>(p) * C2 + half * c + j0 + cc;
        v = *reinterpret_cast<const uint4*>(src);
        if (sizeof(T) == 4) v2 = *reinterpret_cast<const uint4*>(src + 4);
      }
      T* dst = tile + r * 128 + chunk * 8;
      *reinterpret_cast<uint4*>(dst) = v;
      if (sizeof(T) == 4) *reinterpret_cast<uint4*>(dst + 4) = v2;
    }
  }
  __syncthreads();

  const int ncols = tile_px / sp;  // image columns in the tile (all faces), each sp pixels tall
  const T* t1 = tile + 2 * cl;
  const T* t2 = tile + 64 + 2 * cl;
  float csum[32][2];           // per-column sums of this thread (at most 256/sp/8 = 32 columns for sp = 1)
  int ncol_mine = 0;
  for (int col = pl; col < ncols; col += 8, ++ncol_mine) {
    const int f = col / sp, x = col - f * sp;        // face within the tile, image column
    const int base = f * npix + x;                   // pixel (y = 0, x) of that face, tile-relative
    const bool has_l = x > 0, has_r = x + 1 < sp;
    // window rows: top (y-1), mid (y), bot (y+1); each 3 pixels x (2 x1 + 2 x2)
    float top1[3][2], top2[3][2], mid1[3][2], mid2[3][2], bot1[3][2], bot2[3][2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      top1[k][0] = top1[k][1] = top2[k][0] = top2[k][1] = 0.f;
      mid1[k][0] = mid1[k][1] = mid2[k][0] = mid2[k][1] = 0.f;
    }
    auto load_row = [&](int y, float (&r1)[3][2], float (&r2)[3][2]) {
      const int p = base + y * sp;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const bool ok = (k == 1) || (k == 0 ? has_l : has_r);
        float2 a = make_float2(0.f, 0.f), b = a;
        if (ok) {
          a = ld_pair(t1 + (p + k - 1) * 128);
          b = ld_pair(t2 + (p + k - 1) * 128);
        }
        r1[k][0] = a.x; r1[k][1] = a.y; r2[k][0] = b.x; r2[k][1] = b.y;
      }
    };
    load_row(0, mid1, mid2);
    float s0 = 0.f, s1 = 0.f;
    for (int y = 0; y < sp; ++y) {
      if (y + 1 < sp) {
        load_row(y + 1, bot1, bot2);
      } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) bot1[k][0] = bot1[k][1] = bot2[k][0] = bot2[k][1] = 0.f;
      }
      float a1[2] = {b1[0], b1[1]}, a2[2] = {b2[0], b2[1]};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          a1[e] = fmaf(top1[k][e], w1[k][e], a1[e]);
          a1[e] = fmaf(mid1[k][e], w1[3 + k][e], a1[e]);
          a1[e] = fmaf(bot1[k][e], w1[6 + k][e], a1[e]);
          a2[e] = fmaf(top2[k][e], w2[k][e], a2[e]);
          a2[e] = fmaf(mid2[k][e], w2[3 + k][e], a2[e]);
          a2[e] = fmaf(bot2[k][e], w2[6 + k][e], a2[e]);
        }
      }
      const float o0 = a1[0] * a2[0], o1 = a1[1] * a2[1];
      s0 += o0; s1 += o1;
      const int p = px0 + base + y * sp;
      if (p < total_px) st_pair(g + static_cast<size_t>(p) * c + j0 + 2 * cl, o0, o1);
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          top1[k][e] = mid1[k][e]; top2[k][e] = mid2[k][e];
          mid1[k][e] = bot1[k][e]; mid2[k][e] = bot2[k][e];
        }
    }
    csum[ncol_mine][0] = s0;
    csum[ncol_mine][1] = s1;
  }
  __syncthreads();  // everyone is done reading the input tile
  {
    int i = 0;
    for (int col = pl; col < ncols; col += 8, ++i) st_pair(colsum + col * 64 + 2 * cl, csum[i][0], csum[i][1]);
  }
  __syncthreads();
  // per-face means: (256 / npix) faces x 64 channels, columns added in fixed order
  const int faces_in_tile = tile_px / npix;
  for (int i = threadIdx.x; i < faces_in_tile * 64; i += blockDim.x) {
    const int f = i >> 6, ch = i & 63;
    const int face = px0 / npix + f;
    if (face * npix >= total_px) continue;
    float sum = 0.f;
    for (int q = 0; q < sp; ++q) sum += colsum[(f * sp + q) * 64 + ch];
    pooled[static_cast<size_t>(face) * c + j0 + ch] = from_f32<T>(sum / static_cast<float>(npix));
  }
}

// Faces too large to stage whole (latent 32: 32x32 at the first level): depthwise 3x3 + SimpleGate with the taps read
// from global memory (L1/L2-resident neighbours), thread = (pixel, 2 gate channels); the pool is its own kernel.
// Correctness path for `image_res` 256 (train_refiner.py:27): not tuned.
template <typename T>
__global__ void __launch_bounds__(256) dwconv_gate_any_kernel(const T* __restrict__ h, const float* __restrict__ w9,
                                                              const float* __restrict__ bias, T* __restrict__ g, int sp,
                                                              int c, size_t total) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int half_c = c >> 1;
  const size_t pix = i / half_c;
  const int j = static_cast<int>(i - pix * half_c) * 2;
  const int npix = sp * sp;
  const int pf = static_cast<int>(pix % npix);
  const int py = pf / sp, px = pf - py * sp;
  const int C2 = 2 * c;
  float a1[2] = {bias[j], bias[j + 1]}, a2[2] = {bias[c + j], bias[c + j + 1]};
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
    if (yy < 0 || yy >= sp || xx < 0 || xx >= sp) continue;
    const T* src = h + (pix + static_cast<size_t>((t / 3 - 1) * sp + (t % 3 - 1))) * C2;
    const float2 x1 = ld_pair(src + j), x2 = ld_pair(src + c + j);
    a1[0] = fmaf(x1.x, w9[t * C2 + j], a1[0]);
    a1[1] = fmaf(x1.y, w9[t * C2 + j + 1], a1[1]);
    a2[0] = fmaf(x2.x, w9[t * C2 + c + j], a2[0]);
    a2[1] = fmaf(x2.y, w9[t * C2 + c + j + 1], a2[1]);
  }
  st_pair(g + pix * c + j, a1[0] * a2[0], a1[1] * a2[1]);
}

// pooled[face, ch] = mean over the face's pixels of g[face, p, ch]; thread = (face, channel), pixels in order
template <typename T>
__global__ void __launch_bounds__(256) pool_faces_kernel(const T* __restrict__ g, T* __restrict__ pooled, int npix, int c,
                                                         int faces) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(faces) * c) return;
  const size_t face = i / c;
  const int ch = static_cast<int>(i - face * c);
  const T* src = g + face * npix * c + ch;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int p = 0; p < npix; p += 4) {
    s0 += to_f32(src[static_cast<size_t>(p) * c]);
    s1 += to_f32(src[static_cast<size_t>(p + 1) * c]);
    s2 += to_f32(src[static_cast<size_t>(p + 2) * c]);
    s3 += to_f32(src[static_cast<size_t>(p + 3) * c]);
  }
  pooled[i] = from_f32<T>(((s0 + s1) + (s2 + s3)) / static_cast<float>(npix));
}

// nchw_rows[b, p, ch] = src[b, ch * hw + p]: idc_conv(identity).reshape(B, 2048, n, n) as NHWC rows (model.py:245-246)
__global__ void chw_to_hwc_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int hw) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * C * hw) return;
  const int ch = static_cast<int>(i % C);
  const size_t bp = i / C;
  const int p = static_cast<int>(bp % hw);
  const size_t b = bp / hw;
  dst[i] = src[(b * C + ch) * hw + p];
}

// The same op at the 2x2 and 4x4 levels, where a face is 4 / 16 pixels: one thread owns 4 gate channels of one
// face, holds all of the face's pixels in registers (x1 pass, then x2 pass multiplied in), and needs no shared
// memory, no halo and no cross-thread reduction for the pool.  A warp reads 128 consecutive channels per pixel
// (256 B of bf16).  Same FMA and summation order as dwconv_gate_pool_kernel (zero-padding taps are skipped:
// they add exact zeros), so both kernels produce the same bits.
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
  const uint2 a = *reinterpret_cast<const uint2*>(p);
  float2 f = unpack_bf16x2(a.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(a.y); v[2] = f.x; v[3] = f.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
  uint2 a;
  a.x = pack_bf16x2(v[0], v[1]);
  a.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = a;
}

template <typename T, int SP>
__global__ void __launch_bounds__(256) dwconv_small_kernel(const T* __restrict__ h, const float* __restrict__ w9,
                                                           const float* __restrict__ bias, T* __restrict__ g,
                                                           T* __restrict__ pooled, int c, int faces) {
  constexpr int NP = SP * SP, CH = 4;
  pdl_trigger();
  const int per_face = c / CH;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int face = static_cast<int>(i / per_face);
  const int j = static_cast<int>(i - static_cast<size_t>(face) * per_face) * CH;
  const int C2 = 2 * c;
  pdl_wait();
  if (face >= faces) return;
  float res[NP][CH];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    float in[NP][CH], wt[9][CH], b[CH];
    const T* src = h + static_cast<size_t>(face) * NP * C2 + hf * c + j;
#pragma unroll
    for (int p = 0; p < NP; ++p) ld4(src + static_cast<size_t>(p) * C2, in[p]);
#pragma unroll
    for (int t = 0; t < 9; ++t) ld4(w9 + t * C2 + hf * c + j, wt[t]);
    ld4(bias + hf * c + j, b);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int y = p / SP, x = p % SP;
      float a[CH];
#pragma unroll
      for (int e = 0; e < CH; ++e) a[e] = b[e];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int yy = y + r - 1, xx = x + k - 1;
          if (yy < 0 || yy >= SP || xx < 0 || xx >= SP) continue;
#pragma unroll
          for (int e = 0; e < CH; ++e) a[e] = fmaf(in[yy * SP + xx][e], wt[r * 3 + k][e], a[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < CH; ++e) res[p][e] = hf == 0 ? a[e] : res[p][e] * a[e];
    }
  }
  float tot[CH] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int x = 0; x < SP; ++x) {
    float s[CH] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int y = 0; y < SP; ++y)
#pragma unroll
      for (int e = 0; e < CH; ++e) s[e] += res[y * SP + x][e];
#pragma unroll
    for (int e = 0; e < CH; ++e) tot[e] += s[e];
  }
  T* dst = g + static_cast<size_t>(face) * NP * c + j;
#pragma unroll
  for (int p = 0; p < NP; ++p) st4(dst + static_cast<size_t>(p) * c, res[p]);
#pragma unroll
  for (int e = 0; e < CH; ++e) tot[e] = tot[e] / static_cast<float>(NP);
  st4(pooled + static_cast<size_t>(face) * c + j, tot);
}

// g[m, k] *= s[face(m), k]   (the SCA channel scale, conditional_naf.py:119)
template <typename T>
__global__ void __launch_bounds__(256) scale_rows_kernel(T* __restrict__ g, const float* __restrict__ s, size_t total8,
                                                         int c, int rows_per_face) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const size_t e = i * 8;
  const size_t row = e / c;
  const int k = static_cast<int>(e - row * c);
  const int face = static_cast<int>(row / rows_per_face);
  float v[8], sc[8];
  load8(g + e, v);
  load8(s + static_cast<size_t>(face) * c + k, sc);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= sc[j];
  store8(g + e, v);
}

// SimpleGate on an unpacked [R, 2H] fp32 matrix (time path: model.py:49, conditional_naf.py:19)
__global__ void gate_split_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int H) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(R) * H) return;
  const size_t r = i / H;
  const int j = static_cast<int>(i - r * H);
  out[i] = in[r * 2 * H + j] * in[r * 2 * H + H + j];
}

// SimpleGate on the gate-packed conv4 output (fp32 mode): 128-column groups [x1(64) | x2(64)]
template <typename TOut>
__global__ void gate_packed_kernel(const float* __restrict__ in, TOut* __restrict__ out, size_t rows, int c) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * c) return;
  const size_t r = i / c;
  const int j = static_cast<int>(i - r * c);
  const int grp = j >> 6, k = j & 63;
  const float* src = in + r * 2 * c + grp * 128;
  out[i] = from_f32<TOut>(src[k] * src[64 + k]);
}

// space-to-depth for the 2x2 stride-2 down conv (model.py:86): [B,n,n,c] fp32 -> [B,(n/2)^2, (i,j,c)] T
template <typename T>
__global__ void __launch_bounds__(256) s2d_kernel(const float* __restrict__ x, T* __restrict__ out, int B, int n, int c) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n2 = n >> 1;
  const size_t total8 = static_cast<size_t>(B) * n2 * n2 * 4 * c / 8;
  if (i >= total8) return;
  const size_t e = i * 8;
  const int K = 4 * c;
  const size_t orow = e / K;
  const int k = static_cast<int>(e - orow * K);
  const int q = k / c, ch = k - q * c;
  const int face = static_cast<int>(orow / (n2 * n2));
  const int rem = static_cast<int>(orow - static_cast<size_t>(face) * n2 * n2);
  const int h2 = rem / n2, w2 = rem - h2 * n2;
  const size_t irow = (static_cast<size_t>(face) * n + (2 * h2 + (q >> 1))) * n + (2 * w2 + (q & 1));
  float v[8];
  load8(x + irow * c + ch, v);
  store8(out + e, v);
}

template <typename T>
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ x, T* __restrict__ out, size_t total8) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  float v[8];
  load8(x + i * 8, v);
  store8(out + i * 8, v);
}

// HCA gate application (hca.py:28): f_o = f_d + w_c*f_d + w_s*f_d, optionally after adding the
// hoisted idc_conv(identity) vector to f_d (model.py:245-246, level 0 only).
template <typename T>
__global__ void __launch_bounds__(256) hca_apply_kernel(const float* __restrict__ fd, const float* __restrict__ wc,
                                                        const float* __restrict__ ws, const float* __restrict__ idc,
                                                        T* __restrict__ out, size_t total8, int c, int rows_per_face) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const size_t e = i * 8;
  const size_t row = e / c;
  const int k = static_cast<int>(e - row * c);
  const int face = static_cast<int>(row / rows_per_face);
  float v[8], g[8];
  load8(fd + e, v);
  load8(wc + static_cast<size_t>(face) * c + k, g);
  if (idc != nullptr) {
    float a[8];
    load8(idc + row * c + k, a);   // idc_conv(identity) in NHWC rows (model.py:245-246 reshape(B, 2048, n, n))
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += a[j];
  }
  const float s = ws[row];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = v[j] + g[j] * v[j] + s * v[j];
  store8(out + e, v);
}

// ------------------------------------------------------------------------------------------------
// ending: 3x3 conv 128 -> 4 (model.py:168-176,261-262), NHWC in, NCHW fp32 epsilon out.
// one block per face: the S*S x 128 input tile is staged in shared memory (16-byte chunks XOR-
// swizzled by row so that pixel-per-thread reads are conflict-free), thread = pixel.
// dynamic smem: S*S*128*sizeof(TIn) + 4*9*128*4
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) ending_conv_kernel(const TIn* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ eps,
                                                          int B, int S) {
  extern __shared__ __align__(16) uint8_t s_end_raw[];
  constexpr int EPC = 16 / sizeof(TIn);          // elements per 16-byte chunk
  constexpr int CPR = 128 / EPC;                 // chunks per pixel row
  const int npix = S * S;
  TIn* tile = reinterpret_cast<TIn*>(s_end_raw);
  float* s_w = reinterpret_cast<float*>(s_end_raw + static_cast<size_t>(npix) * 128 * sizeof(TIn));
  const int face = blockIdx.x;
  pdl_trigger();
  {  // 4608 floats = 1152 float4, all loads of a thread in flight
    float4 t[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int idx = i * 256 + threadIdx.x;
      if (idx < 1152) t[i] = __ldg(reinterpret_cast<const float4*>(w) + idx);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int idx = i * 256 + threadIdx.x;
      if (idx < 1152) reinterpret_cast<float4*>(s_w)[idx] = t[i];
    }
  }
  pdl_wait();
  const TIn* xf = x + static_cast<size_t>(face) * npix * 128;
  for (int i0 = threadIdx.x; i0 < npix * CPR; i0 += 8 * blockDim.x) {  // 8 loads in flight per thread
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < npix * CPR) v[u] = *reinterpret_cast<const uint4*>(xf + static_cast<size_t>(i / CPR) * 128 + (i % CPR) * EPC);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      const int r = i / CPR, ck = i % CPR;
      if (i < npix * CPR) *reinterpret_cast<uint4*>(tile + r * 128 + ((ck ^ (r % CPR)) * EPC)) = v[u];
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < npix; p += blockDim.x) {
    const int py = p / S, px = p - py * S;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
      if (yy < 0 || yy >= S || xx < 0 || xx >= S) continue;
      const int r = yy * S + xx;
      const TIn* row = tile + r * 128;
      const float* wt = s_w + tap * 128;
#pragma unroll 4
      for (int ck = 0; ck < CPR; ++ck) {
        float v[EPC];
        if (sizeof(TIn) == 2) {
          float t8[8];
          load8(reinterpret_cast<const bf16*>(row) + ((ck ^ (r % CPR)) * EPC), t8);
#pragma unroll
          for (int k = 0; k < EPC; ++k) v[k] = t8[k];
        } else {
          const float4 q = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + ((ck ^ (r % CPR)) * EPC));
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const float* wr = wt + o * 9 * 128 + ck * EPC;
#pragma unroll
          for (int k = 0; k < EPC; ++k) acc[o] = fmaf(v[k], wr[k], acc[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) eps[((static_cast<size_t>(face) * 4 + o) * S + py) * S + px] = acc[o] + bias[o];
  }
}

// ------------------------------------------------------------------------------------------------
// x_{t-1} update (diffusers DDIMScheduler.step / DDPMScheduler.step; reference call site
// train_refiner.py:120) with Philox4x32-10 + Box-Muller noise keyed by (seed, face, step, group).
// One thread = 4 consecutive elements = one Philox group.  Pure streaming: 12-16 B/element.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r > 0) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  }
}
// ending for faces too large for one tile (latent 32): one block = `band` image rows of one face plus a one-row halo,
// thread = pixel.  grid (B, S / band); dynamic smem (band + 2) * S * 128 * sizeof(TIn) + 4*9*128*4.  Correctness path
// for `image_res` 256: same arithmetic as ending_conv_kernel, not tuned.
template <typename TIn>
__global__ void __launch_bounds__(256) ending_conv_band_kernel(const TIn* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ eps,
                                                               int S, int band) {
  extern __shared__ __align__(16) uint8_t s_endb_raw[];
  constexpr int EPC = 16 / sizeof(TIn);
  constexpr int CPR = 128 / EPC;
  const int rows_t = band + 2;
  TIn* tile = reinterpret_cast<TIn*>(s_endb_raw);
  float* s_w = reinterpret_cast<float*>(s_endb_raw + static_cast<size_t>(rows_t) * S * 128 * sizeof(TIn));
  const int face = blockIdx.x, y0 = blockIdx.y * band;
  pdl_trigger();
  for (int i = threadIdx.x; i < 1152; i += blockDim.x) reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  pdl_wait();
  const TIn* xf = x + static_cast<size_t>(face) * S * S * 128;
  for (int i = threadIdx.x; i < rows_t * S * CPR; i += blockDim.x) {
    const int r = i / CPR, ck = i % CPR;           // r = tile pixel: (ty, xx)
    const int ty = r / S, xx = r - ty * S;
    const int yy = y0 - 1 + ty;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (yy >= 0 && yy < S) v = *reinterpret_cast<const uint4*>(xf + static_cast<size_t>(yy * S + xx) * 128 + ck * EPC);
    *reinterpret_cast<uint4*>(tile + r * 128 + ((ck ^ (r % CPR)) * EPC)) = v;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < band * S; p += blockDim.x) {
    const int ly = p / S, px = p - ly * S;
    const int py = y0 + ly;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int xx = px + tap % 3 - 1;
      if (xx < 0 || xx >= S) continue;              // rows outside the face are zero rows of the tile
      const int r = (ly + tap / 3) * S + xx;
      const TIn* row = tile + r * 128;
      const float* wt = s_w + tap * 128;
#pragma unroll 4
      for (int ck = 0; ck < CPR; ++ck) {
        float v[EPC];
        if (sizeof(TIn) == 2) {
          float t8[8];
          load8(reinterpret_cast<const bf16*>(row) + ((ck ^ (r % CPR)) * EPC), t8);
#pragma unroll
          for (int k = 0; k < EPC; ++k) v[k] = t8[k];
        } else {
          const float4 q = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + ((ck ^ (r % CPR)) * EPC));
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const float* wr = wt + o * 9 * 128 + ck * EPC;
#pragma unroll
          for (int k = 0; k < EPC; ++k) acc[o] = fmaf(v[k], wr[k], acc[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) eps[((static_cast<size_t>(face) * 4 + o) * S + py) * S + px] = acc[o] + bias[o];
  }
}

__device__ __forceinline__ float philox_unit(uint32_t r) {
  return (static_cast<float>(r >> 9) + 0.5f) * 1.1920928955078125e-07f;  // 2^-23
}

// The scheduler step (DDIM / DDPM, coefficients from schedulers.py::step_coefficients) on one group of four
// consecutive latent values: i = face * groups + grp indexes float4s of x; Philox counter = (grp, global face, step).
// Shared by sampler_update_kernel and by the fused ending kernel (edge_convs.cuh) so both produce the same bits.
__device__ __forceinline__ void sampler_update_group(float* __restrict__ x, float4 ev, const StepCoef cf, int step, int face,
                                                     int grp, size_t i, const float* __restrict__ noise,
                                                     unsigned long long seed, long long first_face, int batch, int groups) {
  const float4 xv = *reinterpret_cast<const float4*>(x + i * 4);
  float xs[4] = {xv.x, xv.y, xv.z, xv.w};
  const float es[4] = {ev.x, ev.y, ev.z, ev.w};
  float z[4] = {0.f, 0.f, 0.f, 0.f};
  if (cf.k_noise != 0.f) {
    if (noise != nullptr) {
      const float4 nv = *reinterpret_cast<const float4*>(noise + (static_cast<size_t>(step) * batch * groups + i) * 4);
      z[0] = nv.x; z[1] = nv.y; z[2] = nv.z; z[3] = nv.w;
    } else {
      uint32_t c[4] = {static_cast<uint32_t>(grp), static_cast<uint32_t>(first_face + face),
                       static_cast<uint32_t>(step), 0x48494644u};
      philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
      const float r0 = sqrtf(-2.f * logf(philox_unit(c[0])));
      const float a0 = 6.283185307179586f * philox_unit(c[1]);
      const float r1 = sqrtf(-2.f * logf(philox_unit(c[2])));
      const float a1 = 6.283185307179586f * philox_unit(c[3]);
      z[0] = r0 * cosf(a0); z[1] = r0 * sinf(a0);
      z[2] = r1 * cosf(a1); z[3] = r1 * sinf(a1);
    }
  }
#pragma unroll
  // every product and sum rounded on its own (no FMA contraction): that is how the torch ops of diffusers' step()
  // evaluate it, and it keeps the two kernels that inline this function bit-identical
  for (int j = 0; j < 4; ++j) {
    float x0 = __fdiv_rn(__fsub_rn(xs[j], __fmul_rn(cf.sqrt_beta_prod, es[j])), cf.sqrt_alpha_prod);
    if (cf.clip > 0.f) x0 = fminf(fmaxf(x0, -cf.clip), cf.clip);
    float r = __fmul_rn(cf.k_x0, x0);
    if (cf.k_eps != 0.f) r = __fadd_rn(r, __fmul_rn(cf.k_eps, es[j]));
    if (cf.k_x != 0.f) r = __fadd_rn(r, __fmul_rn(cf.k_x, xs[j]));
    if (cf.k_noise != 0.f) r = __fadd_rn(r, __fmul_rn(cf.k_noise, z[j]));
    xs[j] = r;
  }
  *reinterpret_cast<float4*>(x + i * 4) = make_float4(xs[0], xs[1], xs[2], xs[3]);
}

__global__ void __launch_bounds__(256) sampler_update_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                                             const StepCoef* __restrict__ coefs,
                                                             const StepState* __restrict__ state, int fixed_step,
                                                             const float* __restrict__ noise, unsigned long long seed,
                                                             long long first_face, int batch, int elems_per_face) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int groups = elems_per_face >> 2;
  pdl_trigger();
  pdl_wait();
  if (i >= static_cast<size_t>(batch) * groups) return;
  const int step = state != nullptr ? state->step : fixed_step;
  const int face = static_cast<int>(i / groups);
  const int grp = static_cast<int>(i - static_cast<size_t>(face) * groups);
  const float4 ev = *reinterpret_cast<const float4*>(eps + i * 4);
  sampler_update_group(x, ev, coefs[step], step, face, grp, i, noise, seed, first_face, batch, groups);
}

// end of a sampler step: step += 1 and point every face at the next table row (single block)
__global__ void advance_rows_kernel(StepState* st, int* __restrict__ row_idx, int n) {
  pdl_trigger();
  pdl_wait();
  const int next = st->step + 1;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) row_idx[i] = next;
  if (threadIdx.x == 0) st->step = next;
}
__global__ void set_step_kernel(StepState* st, int v) { st->step = v; }

// sinusoidal embedding (model.py:22-29): emb[r, i] = sin(t_r f_i), emb[r, 64+i] = cos(t_r f_i)
__global__ void time_embed_kernel(const float* __restrict__ t, const float* __restrict__ freqs, float* __restrict__ emb,
                                  int R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 64) return;
  const int r = i >> 6, j = i & 63;
  const float a = t[r] * freqs[j];
  emb[r * 128 + j] = sinf(a);
  emb[r * 128 + 64 + j] = cosf(a);
}

// layout conversion at the boundary and for taps
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, int HW) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * C * HW) return;
  const int c = static_cast<int>(i % C);
  const size_t r = i / C;
  const int p = static_cast<int>(r % HW);
  const int b = static_cast<int>(r / HW);
  out[i] = in[(static_cast<size_t>(b) * C + c) * HW + p];
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int B, int C, int HW, int ld) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * C * HW) return;
  const int p = static_cast<int>(i % HW);
  const size_t r = i / HW;
  const int c = static_cast<int>(r % C);
  const int b = static_cast<int>(r / C);
  out[i] = to_f32(in[(static_cast<size_t>(b) * HW + p) * ld + c]);
}

// avg-pool + max-pool over pixels of an NHWC fp32 tensor (hca.py:34-36): out[b,c] = mean + max
__global__ void pool_avgmax_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int HW, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  float s = 0.f, mx = -INFINITY;
  for (int p = 0; p < HW; ++p) {
    const float v = in[(static_cast<size_t>(b) * HW + p) * C + c];
    s += v;
    mx = fmaxf(mx, v);
  }
  out[i] = s / static_cast<float>(HW) + mx;
}

// weight repack: dst[n, kd] = rs[n] * src[perm[n]][kmap(kd)]
//   taps == 1: src is [N_src, K];  taps > 1: src is OIHW [N_src, C, taps], kd = tap*C + c
template <typename TDst>
__global__ void pack_rows_kernel(const float* __restrict__ src, TDst* __restrict__ dst, const int* __restrict__ perm,
                                 const float* __restrict__ rs, int N, int Kd, int taps) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(N) * Kd) return;
  const int n = static_cast<int>(i / Kd), kd = static_cast<int>(i - static_cast<size_t>(n) * Kd);
  const int sn = perm != nullptr ? perm[n] : n;
  int sk = kd;
  if (taps > 1) {
    const int C = Kd / taps;
    const int tap = kd / C, c = kd - tap * C;
    sk = c * taps + tap;
  }
  float v = src[static_cast<size_t>(sn) * Kd + sk];
  if (rs != nullptr) v *= rs[n];
  dst[i] = from_f32<TDst>(v);
}

}  // namespace hd
