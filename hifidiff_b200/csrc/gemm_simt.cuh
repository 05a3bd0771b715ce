// fp32 FFMA GEMM (CUDA cores):  out = epilogue(A[M,K] * W[N,K]^T), fully bounds-checked in M and N.
// This is the fp32 correctness mode (rel-L2 <= 1e-5 against the reference, which tcgen05 cannot give:
// kind::tf32 keeps 10 mantissa bits) and the engine for the once-per-schedule / once-per-face work
// (time-modulation tables, HCA gates, idc_conv).  K must be a multiple of 4.
#pragma once

#include "common.cuh"

namespace hd {
namespace simt {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtArgs {
  int M, N, K;
  const void* A; int lda; int a_mode; int sp; int C;
  const void* W; int ldw;
  const float* bias;
  void* out; int ldo;
  const float* resid; int ldr;
};

template <typename TA>
__device__ __forceinline__ void load_a4(const SimtArgs& g, int m, int k, float (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if (m >= g.M || k >= g.K) return;
  const TA* src;
  if (g.a_mode == A_CONV3) {
    const int tap = k / g.C, c = k - tap * g.C;
    const int sp = g.sp;
    const int face = m / (sp * sp);
    const int rem = m - face * sp * sp;
    const int h = rem / sp + tap / 3 - 1, w = rem % sp + tap % 3 - 1;
    if (h < 0 || h >= sp || w < 0 || w >= sp) return;
    src = reinterpret_cast<const TA*>(g.A) + (static_cast<size_t>(face) * sp * sp + h * sp + w) * g.C + c;
  } else {
    src = reinterpret_cast<const TA*>(g.A) + static_cast<size_t>(m) * g.lda + k;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = to_f32(src[i]);
}

template <typename TA, typename TW, typename TOut, int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtArgs g) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  pdl_trigger();
  pdl_wait();
  const int lr = t >> 2, lk = (t & 3) * 4;  // loader: row/col within tile, k quad
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += TK) {
    float a[4], w[4];
    load_a4<TA>(g, m0 + lr, k0 + lk, a);
    {
      const int n = n0 + lr, k = k0 + lk;
      w[0] = w[1] = w[2] = w[3] = 0.f;
      if (n < g.N && k < g.K) {
        const TW* src = reinterpret_cast<const TW*>(g.W) + static_cast<size_t>(n) * g.ldw + k;
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = to_f32(src[i]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[lk + i][lr] = a[i]; Ws[lk + i][lr] = w[i]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (EPI == EPI_PIXSHUF) {
        const int quarter = g.N >> 2;
        const int q = n / quarter, kch = n - q * quarter;
        const int sp = g.sp;
        const int face = m / (sp * sp);
        const int rem = m - face * sp * sp;
        const int h = rem / sp, w = rem - h * sp;
        const size_t orow = (static_cast<size_t>(face) * (2 * sp) + (2 * h + (q >> 1))) * (2 * sp) + (2 * w + (q & 1));
        float* o = reinterpret_cast<float*>(g.out) + orow * g.ldo + kch;
        *o += v;
        continue;
      }
      if (g.bias != nullptr) v += g.bias[n];
      if (EPI == EPI_RELU) v = fmaxf(v, 0.f);
      if (EPI == EPI_SIGMOID) v = 1.f / (1.f + expf(-v));
      if (EPI == EPI_RESID) v += g.resid[static_cast<size_t>(m) * g.ldr + n];
      reinterpret_cast<TOut*>(g.out)[static_cast<size_t>(m) * g.ldo + n] = from_f32<TOut>(v);
    }
  }
}

}  // namespace simt
}  // namespace hd
