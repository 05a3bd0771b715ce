// CoarseRestoration (reference models/cr/model.py:8-88, models/cr/stn.py:9-52): the kernels that are not GEMMs.
// CR runs once per face before the sampling loop, in fp32 throughout (the spatial transformers resample the
// residual stream with data-dependent coordinates, so this stage keeps the fp32 arithmetic of the reference; its
// 1x1 / 2x2 convolutions run on the FFMA GEMM, gemm_simt.cuh).  Activations are NHWC fp32.
#pragma once
#include "common.cuh"

namespace hd {

// intro: 3x3 conv 3 -> 32, pad 1 (cr/model.py:41-49,78).  x NCHW [B][3][H][H] -> out NHWC [B][H][H][32].
// w [27][32] (k = (c*3 + ky)*3 + kx), thread = four pixels of a row x 8 output channels: every weight vector read
// from shared memory serves four pixels (the one-pixel version issued 81 loads for 216 FMAs and was LSU-bound at
// 3x the time of writing its output).  Taps are added in (c, ky, kx) order after the bias.
__global__ void __launch_bounds__(256) cr_intro_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, float* __restrict__ out, int B, int H) {
  __shared__ __align__(16) float sw[27 * 32];
  pdl_trigger();
  for (int i = threadIdx.x; i < 27 * 32; i += 256) sw[i] = w[i];
  __syncthreads();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * H * (H / 4) * 4) return;
  const int cg = static_cast<int>(i & 3);
  size_t r = i >> 2;
  const int px0 = static_cast<int>(r % (H / 4)) * 4; r /= (H / 4);
  const int py = static_cast<int>(r % H);
  const int face = static_cast<int>(r / H);
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = acc[2][j] = acc[3][j] = b[cg * 8 + j];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = py + ky - 1;
      if (yy < 0 || yy >= H) continue;
      const float* row = x + ((static_cast<size_t>(face) * 3 + c) * H + yy) * H + px0;
      float v[6];
      v[0] = px0 > 0 ? row[-1] : 0.f;
      const float4 mid = *reinterpret_cast<const float4*>(row);
      v[1] = mid.x; v[2] = mid.y; v[3] = mid.z; v[4] = mid.w;
      v[5] = px0 + 4 < H ? row[4] : 0.f;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* wr = sw + ((c * 3 + ky) * 3 + kx) * 32 + cg * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int xx = px0 + p + kx - 1;
          if (xx < 0 || xx >= H) continue;     // the reference's zero padding: the tap is skipped, not added as 0
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(v[p + kx], ww[j], acc[p][j]);
        }
      }
    }
#pragma unroll
  for (int p = 0; p < 4; ++p) store8(out + ((static_cast<size_t>(face) * H + py) * H + px0 + p) * 32 + cg * 8, acc[p]);
}

// outro: 3x3 conv 32 -> 3, pad 1 (cr/model.py:50-58,86).  x NHWC [B][H][H][32] -> out NCHW [B][3][H][H].
// w [3][9][32] in shared memory.  Eight lanes per pixel, four input channels each: a tap of a pixel is one contiguous
// 128-byte read (one thread per pixel made every load instruction touch 32 cache lines), then a three-step shuffle
// reduction over the eight lanes.
__global__ void __launch_bounds__(256) cr_outro_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, float* __restrict__ out, int B, int H) {
  __shared__ __align__(16) float sw[3 * 9 * 32];
  pdl_trigger();
  for (int i = threadIdx.x; i < 3 * 9 * 32; i += 256) sw[i] = w[i];
  __syncthreads();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // grid covers B*H*H*8 exactly (H % 4 == 0)
  const int l = static_cast<int>(i & 7);
  size_t r = i >> 3;
  const bool valid = r < static_cast<size_t>(B) * H * H;
  if (!valid) r = 0;
  const int px = static_cast<int>(r % H); r /= H;
  const int py = static_cast<int>(r % H);
  const int face = static_cast<int>(r / H);
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = px + kx - 1;
      if (xx < 0 || xx >= H) continue;
      const float4 v = *reinterpret_cast<const float4*>(x + ((static_cast<size_t>(face) * H + yy) * H + xx) * 32 + l * 4);
      const int tap = ky * 3 + kx;
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const float4 ww = *reinterpret_cast<const float4*>(sw + (o * 9 + tap) * 32 + l * 4);
        acc[o] = fmaf(v.x, ww.x, acc[o]); acc[o] = fmaf(v.y, ww.y, acc[o]);
        acc[o] = fmaf(v.z, ww.z, acc[o]); acc[o] = fmaf(v.w, ww.w, acc[o]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 3; ++o) {
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], d);
  }
  if (valid && l < 3) {
    const float v = l == 0 ? acc[0] : (l == 1 ? acc[1] : acc[2]);
    out[((static_cast<size_t>(face) * 3 + l) * H + py) * H + px] = v + b[l];
  }
}

// depthwise 3x3 (pad 1, bias) + SimpleGate at any spatial size (cr/naf.py:34-42,113-114): h [B][n][n][2c] ->
// g [B][n][n][c] = dw(x1) * dw(x2).  Neighbours come straight from global memory (L1/L2): thread = one pixel x 4
// gate channels.  w9 [9][2c], bias [2c].
__global__ void __launch_bounds__(256) cr_dwconv_gate_kernel(const float* __restrict__ h, const float* __restrict__ w9,
                                                             const float* __restrict__ bias, float* __restrict__ g, int B,
                                                             int n, int c) {
  pdl_trigger();
  pdl_wait();
  const int c4 = c / 4, C2 = 2 * c;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * n * n * c4) return;
  const int j = static_cast<int>(i % c4) * 4;
  size_t r = i / c4;
  const int px = static_cast<int>(r % n); r /= n;
  const int py = static_cast<int>(r % n);
  const int face = static_cast<int>(r / n);
  float4 a1 = *reinterpret_cast<const float4*>(bias + j), a2 = *reinterpret_cast<const float4*>(bias + c + j);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
    if (yy < 0 || yy >= n || xx < 0 || xx >= n) continue;
    const float* src = h + ((static_cast<size_t>(face) * n + yy) * n + xx) * C2;
    const float4 x1 = *reinterpret_cast<const float4*>(src + j), x2 = *reinterpret_cast<const float4*>(src + c + j);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w9 + t * C2 + j));
    const float4 w2 = __ldg(reinterpret_cast<const float4*>(w9 + t * C2 + c + j));
    a1.x = fmaf(x1.x, w1.x, a1.x); a1.y = fmaf(x1.y, w1.y, a1.y); a1.z = fmaf(x1.z, w1.z, a1.z); a1.w = fmaf(x1.w, w1.w, a1.w);
    a2.x = fmaf(x2.x, w2.x, a2.x); a2.y = fmaf(x2.y, w2.y, a2.y); a2.z = fmaf(x2.z, w2.z, a2.z); a2.w = fmaf(x2.w, w2.w, a2.w);
  }
  *reinterpret_cast<float4*>(g + ((static_cast<size_t>(face) * n + py) * n + px) * c + j) =
      make_float4(a1.x * a2.x, a1.y * a2.y, a1.z * a2.z, a1.w * a2.w);
}

// The same, one thread per (column, 4 gate channels) walking down a strip of R rows: every input row is loaded once
// (3 columns x 2 halves) and scattered into the accumulators of the three output rows it touches, the 18 weight
// vectors stay in registers — 6 loads per output instead of 36 (the per-pixel kernel above was L1-bound at a quarter
// of the HBM rate on the 128x128 and 64x64 stages).  The taps reach each accumulator in the same order as above
// (bias, then ky-major, kx-minor), so the results are bit-identical.
template <int R>
__global__ void __launch_bounds__(128, 3) cr_dwconv_gate_strip_kernel(const float* __restrict__ h, const float* __restrict__ w9,
                                                                   const float* __restrict__ bias, float* __restrict__ g,
                                                                   int B, int n, int c) {
  pdl_trigger();
  const int c4 = c / 4, C2 = 2 * c, strips = n / R;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * strips * n * c4) return;
  const int j = static_cast<int>(i % c4) * 4;
  size_t r = i / c4;
  const int px = static_cast<int>(r % n); r /= n;
  const int y0 = static_cast<int>(r % strips) * R;
  const int face = static_cast<int>(r / strips);
  float4 w[9][2];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    w[t][0] = __ldg(reinterpret_cast<const float4*>(w9 + t * C2 + j));
    w[t][1] = __ldg(reinterpret_cast<const float4*>(w9 + t * C2 + c + j));
  }
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + j)), b2 = __ldg(reinterpret_cast<const float4*>(bias + c + j));
  pdl_wait();
  const float* base = h + static_cast<size_t>(face) * n * n * C2 + j;
  float* out = g + static_cast<size_t>(face) * n * n * c + j;
  const bool has_l = px > 0, has_r = px + 1 < n;
  float4 acc[3][2];   // output rows y - 1, y, y + 1 of the input row being scattered, at slots (y - 1) % 3 ...
#pragma unroll
  for (int k = 0; k < 3; ++k) { acc[k][0] = b1; acc[k][1] = b2; }
  auto fma4 = [](float4& a, const float4& x, const float4& ww) {
    a.x = fmaf(x.x, ww.x, a.x); a.y = fmaf(x.y, ww.y, a.y); a.z = fmaf(x.z, ww.z, a.z); a.w = fmaf(x.w, ww.w, a.w);
  };
#pragma unroll
  for (int s = 0; s < R + 2; ++s) {          // input row y = y0 - 1 + s
    const int y = y0 - 1 + s;
    if (y >= 0 && y < n) {
      const float* row = base + (static_cast<size_t>(y) * n + px) * C2;
      float4 x[3][2];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const bool ok = kx == 0 ? has_l : (kx == 2 ? has_r : true);
        if (ok) {
          x[kx][0] = *reinterpret_cast<const float4*>(row + (kx - 1) * C2);
          x[kx][1] = *reinterpret_cast<const float4*>(row + (kx - 1) * C2 + c);
        }
      }
      // input row y is tap row ky of output row Y = y0 + s - ky, whose accumulators live in slot (s - ky) % 3
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        if (s - ky < 0 || s - ky >= R) continue;   // static after unrolling
        const int slot = (s - ky + 3) % 3;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const bool ok = kx == 0 ? has_l : (kx == 2 ? has_r : true);
          if (ok) { fma4(acc[slot][0], x[kx][0], w[ky * 3 + kx][0]); fma4(acc[slot][1], x[kx][1], w[ky * 3 + kx][1]); }
        }
      }
    }
    if (s >= 2) {                              // output row y0 + s - 2 has received its three tap rows
      const int slot = (s + 1) % 3;
      const float4 a1 = acc[slot][0], a2 = acc[slot][1];
      *reinterpret_cast<float4*>(out + (static_cast<size_t>(y0 + s - 2) * n + px) * c) = make_float4(a1.x * a2.x, a1.y * a2.y, a1.z * a2.z, a1.w * a2.w);
      acc[slot][0] = b1; acc[slot][1] = b2;
    }
  }
}

// per-face channel mean (AdaptiveAvgPool2d(1), cr/naf.py:57): g [B][HW][c] -> pooled [B][c].
// grid (c/32, B), block 1024: lane = channel, 32 row lanes, fixed-order reduction.
constexpr int kCrPoolLanes = 32;
__global__ void __launch_bounds__(1024) cr_pool_kernel(const float* __restrict__ g, float* __restrict__ pooled, int HW, int c) {
  __shared__ float red[kCrPoolLanes][32];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane, face = blockIdx.y;
  const float* src = g + static_cast<size_t>(face) * HW * c + ch;
  float s = 0.f;
  for (int p = rl; p < HW; p += kCrPoolLanes) s += src[static_cast<size_t>(p) * c];
  red[rl][lane] = s;
  __syncthreads();
  if (rl == 0) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < kCrPoolLanes; ++k) tot += red[k][lane];
    pooled[static_cast<size_t>(face) * c + ch] = tot / static_cast<float>(HW);
  }
}

// SimpleGate on a plain [rows][2c] tensor (cr/naf.py:120-121; utils.py:57-60): out[r][k] = in[r][k] * in[r][c + k].
__global__ void __launch_bounds__(256) cr_gate_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, int c) {
  pdl_trigger();
  pdl_wait();
  const int c4 = c / 4;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * c4) return;
  const size_t r = i / c4;
  const int j = static_cast<int>(i - r * c4) * 4;
  const float4 a = *reinterpret_cast<const float4*>(in + r * 2 * c + j);
  const float4 b = *reinterpret_cast<const float4*>(in + r * 2 * c + c + j);
  *reinterpret_cast<float4*>(out + r * c + j) = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}

// a += b  (the skip add in front of the first decoder, cr/model.py:83)
__global__ void __launch_bounds__(256) cr_add_kernel(float* __restrict__ a, const float* __restrict__ b, size_t total4) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  float4 x = reinterpret_cast<float4*>(a)[i];
  const float4 y = reinterpret_cast<const float4*>(b)[i];
  x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
  reinterpret_cast<float4*>(a)[i] = x;
}

__global__ void __launch_bounds__(256) cr_zero_kernel(float* __restrict__ a, size_t total4) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total4) reinterpret_cast<float4*>(a)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// fp32 -> three bf16 column blocks for the split-precision tensor-core GEMM: x = hi + lo with hi = bf16(x),
// lo = bf16(x - hi).  Activations (mode 0) are laid out [hi | lo | hi] and weights (mode 1) [hi | hi | lo], so one
// bf16 GEMM over K' = 3K accumulates a_hi w_hi + a_lo w_hi + a_hi w_lo in fp32 — everything but the lo*lo term
// (2^-16 relative), i.e. fp32-grade products on the tcgen05 pipe.  in [rows][K] -> out [rows][3K].
__global__ void __launch_bounds__(256) cr_split3_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t rows, int K,
                                                        int mode) {
  pdl_trigger();
  pdl_wait();
  const int k8 = K / 8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * k8) return;
  const size_t r = i / k8;
  const int k0 = static_cast<int>(i - r * k8) * 8;
  float v[8], hi[8], lo[8];
  load8(in + r * K + k0, v);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    hi[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
    lo[e] = v[e] - hi[e];
  }
  bf16* o = out + r * 3 * K + k0;
  store8(o, hi);
  store8(o + K, mode == 0 ? lo : hi);
  store8(o + 2 * K, mode == 0 ? hi : lo);
}

// The two element-wise producers of a split GEMM operand, writing [hi | lo | hi] directly (no fp32 round trip):
//   SCA channel scale in front of conv3 (cr/naf.py:116-117): out3 = split(g[m, k] * s[face(m), k])
__global__ void __launch_bounds__(256) cr_scale_split3_kernel(const float* __restrict__ g, const float* __restrict__ s,
                                                              bf16* __restrict__ out3, size_t total8, int c, int rows_per_face) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const size_t e = i * 8;
  const size_t row = e / c;
  const int k = static_cast<int>(e - row * c);
  const int face = static_cast<int>(row / rows_per_face);
  float v[8], sc[8], hi[8], lo[8];
  load8(g + e, v);
  load8(s + static_cast<size_t>(face) * c + k, sc);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] *= sc[j];
    hi[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
    lo[j] = v[j] - hi[j];
  }
  bf16* o = out3 + row * 3 * c + k;
  store8(o, hi);
  store8(o + c, lo);
  store8(o + 2 * c, hi);
}
//   SimpleGate in front of conv5 (cr/naf.py:122-123): out3 = split(in[m, k] * in[m, c + k])
__global__ void __launch_bounds__(256) cr_gate_split3_kernel(const float* __restrict__ in, bf16* __restrict__ out3, size_t rows, int c) {
  pdl_trigger();
  pdl_wait();
  const int c8 = c / 8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * c8) return;
  const size_t r = i / c8;
  const int k = static_cast<int>(i - r * c8) * 8;
  float a[8], b[8], hi[8], lo[8];
  load8(in + r * 2 * c + k, a);
  load8(in + r * 2 * c + c + k, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float v = a[j] * b[j];
    hi[j] = __bfloat162float(__float2bfloat16_rn(v));
    lo[j] = v - hi[j];
  }
  bf16* o = out3 + r * 3 * c + k;
  store8(o, hi);
  store8(o + c, lo);
  store8(o + 2 * c, hi);
}

// STN localisation stage (stn.py:20-27): valid k x k conv (Cin -> Cout <= 10) + MaxPool2d(2,2) + ReLU, fused.
//   in NHWC [B][n][n][Cin], w [Cout][k][k][Cin], out NHWC [B][no][no][Cout], no = (n - k + 1) / 2.
// Thread = one pooled pixel x OG output channels (COUT / OG groups): the 2x2 conv outputs under the pool window
// share their inputs.
template <int COUT, int OG>
__global__ void __launch_bounds__(128) cr_stn_conv_pool_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                               const float* __restrict__ b, float* __restrict__ out, int B,
                                                               int n, int Cin, int k, int no) {
  pdl_trigger();
  pdl_wait();
  constexpr int NG = COUT / OG;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * no * no * NG) return;
  const int og = static_cast<int>(i % NG) * OG;
  size_t r = i / NG;
  const size_t pix = r;
  const int px = static_cast<int>(r % no); r /= no;
  const int py = static_cast<int>(r % no);
  const int face = static_cast<int>(r / no);
  float acc[4][OG];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int o = 0; o < OG; ++o) acc[q][o] = 0.f;
  const float* base = in + (static_cast<size_t>(face) * n + 2 * py) * n * Cin + static_cast<size_t>(2 * px) * Cin;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      const float* wk = w + (static_cast<size_t>(og) * k * k + ky * k + kx) * Cin;
      const float* s00 = base + (static_cast<size_t>(ky) * n + kx) * Cin;
      for (int c = 0; c < Cin; c += 4) {
        const float4 v00 = *reinterpret_cast<const float4*>(s00 + c);
        const float4 v01 = *reinterpret_cast<const float4*>(s00 + Cin + c);
        const float4 v10 = *reinterpret_cast<const float4*>(s00 + static_cast<size_t>(n) * Cin + c);
        const float4 v11 = *reinterpret_cast<const float4*>(s00 + static_cast<size_t>(n) * Cin + Cin + c);
#pragma unroll
        for (int o = 0; o < OG; ++o) {
          const float4 ww = __ldg(reinterpret_cast<const float4*>(wk + static_cast<size_t>(o) * k * k * Cin + c));
          acc[0][o] = fmaf(v00.x, ww.x, acc[0][o]); acc[0][o] = fmaf(v00.y, ww.y, acc[0][o]);
          acc[0][o] = fmaf(v00.z, ww.z, acc[0][o]); acc[0][o] = fmaf(v00.w, ww.w, acc[0][o]);
          acc[1][o] = fmaf(v01.x, ww.x, acc[1][o]); acc[1][o] = fmaf(v01.y, ww.y, acc[1][o]);
          acc[1][o] = fmaf(v01.z, ww.z, acc[1][o]); acc[1][o] = fmaf(v01.w, ww.w, acc[1][o]);
          acc[2][o] = fmaf(v10.x, ww.x, acc[2][o]); acc[2][o] = fmaf(v10.y, ww.y, acc[2][o]);
          acc[2][o] = fmaf(v10.z, ww.z, acc[2][o]); acc[2][o] = fmaf(v10.w, ww.w, acc[2][o]);
          acc[3][o] = fmaf(v11.x, ww.x, acc[3][o]); acc[3][o] = fmaf(v11.y, ww.y, acc[3][o]);
          acc[3][o] = fmaf(v11.z, ww.z, acc[3][o]); acc[3][o] = fmaf(v11.w, ww.w, acc[3][o]);
        }
      }
    }
  float* o_ptr = out + pix * COUT + og;
#pragma unroll
  for (int o = 0; o < OG; ++o) {
    const float m = fmaxf(fmaxf(acc[0][o], acc[1][o]), fmaxf(acc[2][o], acc[3][o])) + b[og + o];  // bias commutes with max
    o_ptr[o] = fmaxf(m, 0.f);
  }
}

// The same stage for wide inputs (Cin % 16 == 0: the first localisation conv): the channel reduction is split over
// CS = 4 adjacent lanes, each doing all COUT output channels for a quarter of the channels (4 input + COUT weight
// loads per 16 * COUT FMAs, and a pixel's 4 lanes read 64 contiguous bytes), combined with two shuffles at the end.
// The one-thread-per-pixel form above is bound by L1 bandwidth at 6 loads per 32 FMAs.
template <int COUT>
__global__ void __launch_bounds__(128) cr_stn_conv_pool_cs_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                  const float* __restrict__ b, float* __restrict__ out, int B,
                                                                  int n, int Cin, int k, int no) {
  constexpr int CS = 4;
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int slice = static_cast<int>(i % CS);
  const size_t npix = static_cast<size_t>(B) * no * no;
  const bool valid = i / CS < npix;
  const size_t pix = valid ? i / CS : npix - 1;   // out-of-range lanes redo the last pixel (all lanes stay in the shuffles)
  size_t r = pix;
  const int px = static_cast<int>(r % no); r /= no;
  const int py = static_cast<int>(r % no);
  const int face = static_cast<int>(r / no);
  const int cper = Cin / CS, cbeg = slice * cper;
  float acc[4][COUT];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[q][o] = 0.f;
  const float* base = in + (static_cast<size_t>(face) * n + 2 * py) * n * Cin + static_cast<size_t>(2 * px) * Cin + cbeg;
  const size_t wstride = static_cast<size_t>(k) * k * Cin;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      const float* wk = w + static_cast<size_t>(ky * k + kx) * Cin + cbeg;
      const float* s00 = base + (static_cast<size_t>(ky) * n + kx) * Cin;
      for (int c = 0; c < cper; c += 4) {
        const float4 v00 = *reinterpret_cast<const float4*>(s00 + c);
        const float4 v01 = *reinterpret_cast<const float4*>(s00 + Cin + c);
        const float4 v10 = *reinterpret_cast<const float4*>(s00 + static_cast<size_t>(n) * Cin + c);
        const float4 v11 = *reinterpret_cast<const float4*>(s00 + static_cast<size_t>(n) * Cin + Cin + c);
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
          const float4 ww = __ldg(reinterpret_cast<const float4*>(wk + o * wstride + c));
          acc[0][o] = fmaf(v00.x, ww.x, acc[0][o]); acc[0][o] = fmaf(v00.y, ww.y, acc[0][o]);
          acc[0][o] = fmaf(v00.z, ww.z, acc[0][o]); acc[0][o] = fmaf(v00.w, ww.w, acc[0][o]);
          acc[1][o] = fmaf(v01.x, ww.x, acc[1][o]); acc[1][o] = fmaf(v01.y, ww.y, acc[1][o]);
          acc[1][o] = fmaf(v01.z, ww.z, acc[1][o]); acc[1][o] = fmaf(v01.w, ww.w, acc[1][o]);
          acc[2][o] = fmaf(v10.x, ww.x, acc[2][o]); acc[2][o] = fmaf(v10.y, ww.y, acc[2][o]);
          acc[2][o] = fmaf(v10.z, ww.z, acc[2][o]); acc[2][o] = fmaf(v10.w, ww.w, acc[2][o]);
          acc[3][o] = fmaf(v11.x, ww.x, acc[3][o]); acc[3][o] = fmaf(v11.y, ww.y, acc[3][o]);
          acc[3][o] = fmaf(v11.z, ww.z, acc[3][o]); acc[3][o] = fmaf(v11.w, ww.w, acc[3][o]);
        }
      }
    }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      float v = acc[q][o];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      acc[q][o] = v;
    }
  if (slice == 0 && valid) {
    float* o_ptr = out + pix * COUT;
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      const float m = fmaxf(fmaxf(acc[0][o], acc[1][o]), fmaxf(acc[2][o], acc[3][o])) + b[o];
      o_ptr[o] = fmaxf(m, 0.f);
    }
  }
}

// STN regressor (stn.py:29-33,45-47): theta = W2 relu(W1 xs + b1) + b2.
//   xs [fc] (NHWC order of the localisation output; W1's columns are permuted to match at load), W1 [hid][fc],
//   W2 [6][hid] -> theta [B][6].  hid <= 96.
// grid (ceil(hid / 8), B): one warp per hidden unit (fc is up to 7290 at the 128x128 stage: one block per face left
// 85 serial dot products to 8 warps, 440 us); the hidden vector goes through global memory and the last block of a
// face to finish (ticket, left at zero) applies W2.  Every sum keeps a fixed order.
__global__ void __launch_bounds__(256) cr_stn_fc_kernel(const float* __restrict__ xs, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, float* __restrict__ theta,
                                                        float* __restrict__ hidden, unsigned int* __restrict__ ticket,
                                                        int fc, int hid) {
  __shared__ float hbuf[96];
  __shared__ bool last;
  pdl_trigger();
  pdl_wait();
  const int face = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 8 + warp;
  if (j < hid) {
    const float* x = xs + static_cast<size_t>(face) * fc;
    const float* wr = w1 + static_cast<size_t>(j) * fc;
    float s = 0.f;
#pragma unroll 8
    for (int k = lane; k < fc; k += 32) s = fmaf(x[k], __ldg(wr + k), s);
    s = warp_sum(s);
    if (lane == 0) hidden[face * 96 + j] = fmaxf(s + b1[j], 0.f);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket + face, 1u) + 1u == gridDim.x;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < hid) hbuf[threadIdx.x] = __ldcg(hidden + face * 96 + threadIdx.x);
  __syncthreads();
  if (threadIdx.x < 6) {
    float s = b2[threadIdx.x];
    for (int k = 0; k < hid; ++k) s = fmaf(hbuf[k], w2[threadIdx.x * hid + k], s);
    theta[face * 6 + threadIdx.x] = s;
  }
  if (threadIdx.x == 0) ticket[face] = 0u;
}

// affine_grid + grid_sample (bilinear, zeros padding, align_corners=False; stn.py:49-50) on the NHWC stream.
// Base grid as torch builds it: linspace(-1, 1, W) * (W - 1) / W, the linspace evaluated from both ends.
__device__ __forceinline__ float cr_base_coord(int i, int W) {
  const float step = 2.f / static_cast<float>(W - 1);
  const float v = i < W / 2 ? -1.f + step * static_cast<float>(i) : 1.f - step * static_cast<float>(W - 1 - i);
  return v * static_cast<float>(W - 1) / static_cast<float>(W);
}
__global__ void __launch_bounds__(256) cr_stn_sample_kernel(const float* __restrict__ in, const float* __restrict__ theta,
                                                            float* __restrict__ out, int B, int n, int c) {
  pdl_trigger();
  pdl_wait();
  const int c4 = c / 4;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * n * n * c4) return;
  const int j = static_cast<int>(i % c4) * 4;
  size_t r = i / c4;
  const int px = static_cast<int>(r % n); r /= n;
  const int py = static_cast<int>(r % n);
  const int face = static_cast<int>(r / n);
  const float* th = theta + face * 6;
  const float xb = cr_base_coord(px, n), yb = cr_base_coord(py, n);
  const float gx = xb * th[0] + yb * th[1] + th[2];
  const float gy = xb * th[3] + yb * th[4] + th[5];
  const float ix = ((gx + 1.f) * static_cast<float>(n) - 1.f) / 2.f;
  const float iy = ((gy + 1.f) * static_cast<float>(n) - 1.f) / 2.f;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  const float wgt[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};  // nw, ne, sw, se
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* fbase = in + static_cast<size_t>(face) * n * n * c + j;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int xx = x0 + (q & 1), yy = y0 + (q >> 1);
    if (xx < 0 || xx >= n || yy < 0 || yy >= n) continue;
    const float4 v = *reinterpret_cast<const float4*>(fbase + (static_cast<size_t>(yy) * n + xx) * c);
    acc.x += v.x * wgt[q]; acc.y += v.y * wgt[q]; acc.z += v.z * wgt[q]; acc.w += v.w * wgt[q];
  }
  *reinterpret_cast<float4*>(out + ((static_cast<size_t>(face) * n + py) * n + px) * c + j) = acc;
}

}  // namespace hd
