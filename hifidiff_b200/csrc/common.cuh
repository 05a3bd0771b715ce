// Shared declarations for the hifidiff_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace hd {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// GEMM description shared by the FFMA (fp32) and tcgen05 (bf16) back ends.
//   out = epilogue( A[M,K] * W[N,K]^T )      A and W are both K-contiguous ("TN" GEMM)
// ---------------------------------------------------------------------------------------------
enum Epi : int {
  EPI_BIAS = 0,     // out[m,n] = acc + bias[n]
  EPI_RELU = 1,     // out[m,n] = max(acc + bias[n], 0)
  EPI_SIGMOID = 2,  // out[m,n] = 1 / (1 + exp(-(acc + bias[n])))
  EPI_RESID = 3,    // out[m,n] = resid[m,n] + acc + bias[n]         (fp32 residual stream)
  EPI_GATE = 4,     // packed 128-column groups: out[m, g*64+i] = (acc[m,g*128+i]+b) * (acc[m,g*128+64+i]+b)
  EPI_PIXSHUF = 5,  // 1x1 up-conv + PixelShuffle(2) + skip add, in place on the fp32 skip buffer
  EPI_MUL = 6,      // out[m,n] = (acc + bias[n]) * mul[m,n]: SCA at 1x1 spatial, where the pooled mean is the gated tensor
                    // itself and a face is one row (conditional_naf.py:119 `x * self.sca(x)`); mul is bf16, passed as `resid`
};

enum AMode : int {
  A_PLAIN = 0,  // A is a dense [M, lda] matrix
  A_CONV3 = 1,  // implicit 3x3/pad-1 im2col over an NHWC tensor [B, n, n, C]; K = 9*C, k = tap*C + c
};

enum DType : int { DT_F32 = 0, DT_BF16 = 1 };

struct GemmDesc {
  int M = 0, N = 0, K = 0;
  const void* A = nullptr;
  int lda = 0;
  int a_dtype = DT_F32;
  int a_mode = A_PLAIN;
  int sp = 0;  // A_CONV3: spatial size n;  EPI_PIXSHUF: spatial size of the GEMM rows (input level)
  int C = 0;   // A_CONV3: channels of the NHWC tensor
  const void* W = nullptr;
  int ldw = 0;
  int w_dtype = DT_F32;
  const float* bias = nullptr;
  int epi = EPI_BIAS;
  void* out = nullptr;
  int ldo = 0;
  int out_dtype = DT_F32;
  const float* resid = nullptr;  // EPI_RESID / EPI_PIXSHUF: fp32 [M, ldr]; EPI_MUL: bf16 [M, ldr] multiplicand
  int ldr = 0;
  // split-K until the grid has at least this many CTAs (one chain of a split region gets its share of the SMs)
  int cta_target = 120;
};

// Device-side error word shared by all kernels of a handle (pipeline watchdog).
struct DeviceStatus {
  unsigned int error;  // 0 = ok
  unsigned int where;  // (kernel id << 16) | site
};

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 p = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(p);
}

// Loads 8 consecutive elements as fp32 (pointer must be 16B-aligned for bf16, 32B for fp32).
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 a = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = unpack_bf16x2(a.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(a.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(a.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(a.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 a;
  a.x = pack_bf16x2(v[0], v[1]);
  a.y = pack_bf16x2(v[2], v[3]);
  a.z = pack_bf16x2(v[4], v[5]);
  a.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = a;
}

// Programmatic dependent launch: every per-step kernel lets its successor start launching right away
// (pdl_trigger) and touches activation memory only after its predecessor has fully completed
// (pdl_wait).  Constants (weights, biases) may be fetched before the wait.  Both are no-ops when the
// kernel was launched without the programmatic-stream-serialization attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace hd
