// IDC identity network (ResNet-50 trunk, reference models/idc/model.py:10-55,102-166): the kernels that are
// not GEMM-shaped.  The 1x1 / 3x3 convolutions of the 16 bottlenecks run on the tcgen05 GEMM (gemm_tc.cuh)
// with eval-mode BatchNorm folded into the packed weights; this file holds the 7x7 stem, the max-pool, the
// strided patch gather that feeds the three stride-2 3x3 convs and the stride-2 projections, the
// ReLU + operand cast after each residual add, and the final average pool.  Activations are NHWC.
#pragma once
#include "common.cuh"

namespace hd {

// Weight repack with zero padding: src OIHW [N][C][taps] fp32 -> dst [Npad][taps*Cpad], k = tap*Cpad + c,
// row n scaled by rs[n] (BatchNorm scale).  Rows >= N and channels >= C are zero.
template <typename TDst>
__global__ void idc_pack_conv_kernel(const float* __restrict__ src, TDst* __restrict__ dst, const float* __restrict__ rs,
                                     int N, int C, int taps, int Npad, int Cpad) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int Kd = taps * Cpad;
  if (i >= static_cast<size_t>(Npad) * Kd) return;
  const int n = static_cast<int>(i / Kd), kd = static_cast<int>(i - static_cast<size_t>(n) * Kd);
  const int tap = kd / Cpad, c = kd - tap * Cpad;
  float v = 0.f;
  if (n < N && c < C) v = src[(static_cast<size_t>(n) * C + c) * taps + tap] * rs[n];
  dst[i] = from_f32<TDst>(v);
}

// Stem: conv 7x7 stride 2 pad 3, 3 -> 64 channels, no conv bias, BatchNorm folded, ReLU (idc/model.py:107-110,124).
//   x   NCHW fp32 [B][3][H][H]   (the caller's cr_face, read in place)
//   w   [147][64] fp32, k = (c*7 + ky)*7 + kx, BN scale folded;  b [64] = BN shift
//   out NHWC [B][H/2][H/2][64]
// One block = an 8x8 tile of output pixels x 64 channels; thread = 1 pixel x 16 channels.
constexpr int kStemPatch = 21;  // 8 outputs * stride 2 + 7 - 2
template <typename TOut>
__global__ void __launch_bounds__(256) idc_stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, TOut* __restrict__ out, int H) {
  extern __shared__ float smem[];
  float* sw = smem;                                 // [147][64]
  float* sx = smem + 147 * 64;                      // [3][21][22]
  const int tid = threadIdx.x;
  const int face = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x, Ho = H / 2;
  for (int i = tid; i < 147 * 64 / 4; i += 256) reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(w)[i];
  pdl_trigger();
  pdl_wait();
  const int iy0 = ty * 16 - 3, ix0 = tx * 16 - 3;
  for (int i = tid; i < 3 * kStemPatch * kStemPatch; i += 256) {
    const int c = i / (kStemPatch * kStemPatch), r = i - c * kStemPatch * kStemPatch;
    const int py = r / kStemPatch, px = r - py * kStemPatch;
    const int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < H) v = x[((static_cast<size_t>(face) * 3 + c) * H + iy) * H + ix];
    sx[(c * kStemPatch + py) * 22 + px] = v;
  }
  __syncthreads();
  const int p = tid & 63, cg = tid >> 6;
  const int py = p >> 3, px = p & 7;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = b[cg * 16 + j];
  for (int c = 0; c < 3; ++c)
    for (int ky = 0; ky < 7; ++ky) {
      const float* xr = sx + (c * kStemPatch + 2 * py + ky) * 22 + 2 * px;
      const float* wr = sw + ((c * 7 + ky) * 7) * 64 + cg * 16;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float v = xr[kx];
        const float4* w4 = reinterpret_cast<const float4*>(wr + kx * 64);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ww = w4[q];
          acc[q * 4 + 0] = fmaf(v, ww.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(v, ww.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(v, ww.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(v, ww.w, acc[q * 4 + 3]);
        }
      }
    }
  const int oy = ty * 8 + py, ox = tx * 8 + px;
  TOut* o = out + ((static_cast<size_t>(face) * Ho + oy) * Ho + ox) * 64 + cg * 16;
  float v8[8];
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v8[j] = fmaxf(acc[hlf * 8 + j], 0.f);
    store8(o + hlf * 8, v8);
  }
}

// MaxPool2d(3, stride 2, pad 1) on NHWC [B][n][n][C] -> fp32 identity copy and operand copy [B][n/2][n/2][C]
// (idc/model.py:111,125).  One thread = 8 channels of one output pixel.
template <typename T>
__global__ void __launch_bounds__(256) idc_maxpool_kernel(const T* __restrict__ in, float* __restrict__ out_f,
                                                          T* __restrict__ out_t, int B, int n, int C) {
  pdl_trigger();
  pdl_wait();
  const int no = n / 2, c8 = C / 8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * no * no * c8) return;
  const int cc = static_cast<int>(i % c8);
  size_t r = i / c8;
  const int ox = static_cast<int>(r % no); r /= no;
  const int oy = static_cast<int>(r % no);
  const int face = static_cast<int>(r / no);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
  for (int dy = -1; dy <= 1; ++dy) {
    const int iy = 2 * oy + dy;
    if (iy < 0 || iy >= n) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int ix = 2 * ox + dx;
      if (ix < 0 || ix >= n) continue;
      float v[8];
      load8(in + ((static_cast<size_t>(face) * n + iy) * n + ix) * C + cc * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
    }
  }
  store8(out_f + i * 8, m);
  store8(out_t + i * 8, m);
}

// Patch gather (im2col) for the strided convs: in NHWC [B][n][n][C] -> out [B*no*no][k*k*C], column = tap*C + c,
// no = (n + 2*pad - k)/stride + 1, zero padding.  k=3,stride=2,pad=1: the first conv2 of layer2..4
// (idc/model.py:21-23); k=1,stride=2,pad=0: the stride-2 projection (idc/model.py:141-149).
template <typename T>
__global__ void __launch_bounds__(256) idc_gather_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int n, int C,
                                                         int k, int stride, int pad, int no) {
  pdl_trigger();
  pdl_wait();
  const int c8 = C / 8, per_row = k * k * c8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * no * no * per_row) return;
  const int col = static_cast<int>(i % per_row);
  size_t r = i / per_row;
  const int tap = col / c8, cc = col - tap * c8;
  const int ox = static_cast<int>(r % no); r /= no;
  const int oy = static_cast<int>(r % no);
  const int face = static_cast<int>(r / no);
  const int iy = oy * stride - pad + tap / k, ix = ox * stride - pad + tap % k;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (iy >= 0 && iy < n && ix >= 0 && ix < n) load8(in + ((static_cast<size_t>(face) * n + iy) * n + ix) * C + cc * 8, v);
  store8(out + i * 8, v);
}

// ReLU after the residual add (idc/model.py:52-53): x <- max(x, 0) in place (the next block's identity) and the
// operand copy for the next conv1.
template <typename T>
__global__ void __launch_bounds__(256) idc_relu_cast_kernel(float* __restrict__ x, T* __restrict__ out, size_t total8) {
  pdl_trigger();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  float v[8];
  load8(x + i * 8, v);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
  store8(x + i * 8, v);
  store8(out + i * 8, v);
}

// ReLU + AdaptiveAvgPool2d(1) of the last block (idc/model.py:53,132-133): x fp32 [B][HW][C] -> out [B][C]
// (= NCHW (B, C, 1, 1)).
__global__ void __launch_bounds__(256) idc_relu_avgpool_kernel(const float* __restrict__ x, float* __restrict__ out, int B,
                                                               int HW, int C) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int face = i / C, c = i - face * C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += fmaxf(x[(static_cast<size_t>(face) * HW + p) * C + c], 0.f);
  out[i] = s / static_cast<float>(HW);
}

}  // namespace hd
