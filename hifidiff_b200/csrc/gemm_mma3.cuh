// Small-K fp32-grade GEMM for the wide, shallow stages of CoarseRestoration (models/cr/model.py:8-88: the 1x1 convs
// of the NAF blocks at c = 32 / 64, the 2x2 stride-2 down convs and the 1x1 up convs + PixelShuffle):
//   out[M, N] = epilogue(A[M, K] W[N, K]^T),   M = faces x pixels (10^5 .. 10^6 rows), K = 32 .. 512, N = 32 .. 1024
// These are HBM-bound by shape (a few hundred flops per row against 128 .. 768 bytes of fp32 traffic), but on CUDA
// cores the FFMA issue rate was the limit (18 TFLOP/s, 4-8x off the memory floor).  Here the arithmetic is
// mma.sync.m16n8k8 on TF32 with split operands ("3xTF32") — A and W as tf32 hi + lo (11 + 11 mantissa bits),
// a_hi w_hi + a_lo w_hi + a_hi w_lo accumulated in fp32: everything but the lo x lo term, ~2^-21 relative, and the
// fp32 exponent range — so that the kernel waits on memory, not on the FMA pipe, and the nine data-dependent
// resamplings downstream see fp32-grade features (a bf16 split, 2^-17, cost 8e-4 of the network's output error).
// K is too small here for a TMA / tcgen05 pipeline to amortise its set-up, and A needs the fp32 -> hi / lo
// conversion on the way into shared memory anyway.
//
// CTA = 256 threads = 8 warps, tile 128 rows x BN columns, warp w owns rows 16w .. 16w+15 and all BN columns.
// K goes through in chunks of 32: global -> registers (the next chunk's loads are in flight during the MMAs) ->
// hi / lo split -> shared memory rows of 128 + 16 bytes (conflict-free ldmatrix; a 32-bit element is a b16 pair).
#pragma once

#include "common.cuh"
#include "edge_convs.cuh"

namespace hd {
namespace mma3 {

struct Args {
  const float* A;      // [M, lda] fp32
  const float* w_hi;   // [N, K] tf32
  const float* w_lo;   // [N, K] tf32
  const float* bias;   // [N] or nullptr
  float* out;          // [M, ldo] fp32 (EPI_PIXSHUF: the up-sampled skip buffer, accumulated in place)
  const float* resid;  // EPI_RESID: [M, ldr]
  int lda, ldo, ldr;
  int M, N, K;
  int sp;              // EPI_PIXSHUF: spatial size of the GEMM rows
};

constexpr int BM = 128, KC = 32, ROWB = KC * 4 + 16;

__device__ __forceinline__ float to_tf32(float f) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
  return __uint_as_float(u);
}
__device__ __forceinline__ void mma_1688_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// W fp32 [N, K] -> hi, lo tf32 [N, K] (once, at load)
__global__ void split_hl_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float f = w[i], h = to_tf32(f);
  hi[i] = h;
  lo[i] = to_tf32(f - h);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(256, 2) gemm_mma3_kernel(const Args g) {
  constexpr int NT = BN / 8;
  extern __shared__ __align__(128) uint8_t s_raw[];
  uint8_t* s_a[2] = {s_raw, s_raw + BM * ROWB};                                  // hi, lo
  uint8_t* s_w[2] = {s_raw + 2 * BM * ROWB, s_raw + 2 * BM * ROWB + BN * ROWB};
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  constexpr int WJ = BN * 8 / 256;             // 16-byte weight chunks per thread and matrix
  float4 areg[4];
  uint4 whreg[WJ], wlreg[WJ];
  auto load_w = [&](int k0) {
#pragma unroll
    for (int j = 0; j < WJ; ++j) {
      const int i = tid + 256 * j;
      const size_t off = static_cast<size_t>(n0 + (i >> 3)) * g.K + k0 + (i & 7) * 4;
      whreg[j] = __ldg(reinterpret_cast<const uint4*>(g.w_hi + off));
      wlreg[j] = __ldg(reinterpret_cast<const uint4*>(g.w_lo + off));
    }
  };
  auto load_a = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = tid + 256 * j, row = m0 + (i >> 3);
      areg[j] = row < g.M ? *reinterpret_cast<const float4*>(g.A + static_cast<size_t>(row) * g.lda + k0 + (i & 7) * 4)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = tid + 256 * j;
      const float4 f = areg[j];
      const float4 hi = make_float4(to_tf32(f.x), to_tf32(f.y), to_tf32(f.z), to_tf32(f.w));
      const uint32_t off = (i >> 3) * ROWB + (i & 7) * 16;
      *reinterpret_cast<float4*>(s_a[0] + off) = hi;
      *reinterpret_cast<float4*>(s_a[1] + off) = make_float4(to_tf32(f.x - hi.x), to_tf32(f.y - hi.y), to_tf32(f.z - hi.z), to_tf32(f.w - hi.w));
    }
#pragma unroll
    for (int j = 0; j < WJ; ++j) {
      const int i = tid + 256 * j;
      const uint32_t off = (i >> 3) * ROWB + (i & 7) * 16;
      *reinterpret_cast<uint4*>(s_w[0] + off) = whreg[j];
      *reinterpret_cast<uint4*>(s_w[1] + off) = wlreg[j];
    }
  };
  pdl_trigger();
  load_w(0);                                   // constants: before the dependency wait
  pdl_wait();
  load_a(0);
  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const uint32_t a_lane = (warp * 16 + (lane & 15)) * ROWB + (lane >> 4) * 16;
  const uint32_t w_lane = ((lane >> 4) * 8 + (lane & 7)) * ROWB + ((lane >> 3) & 1) * 16;
  const uint32_t ah_u32 = edge::smem_addr(s_a[0]) + a_lane, al_u32 = edge::smem_addr(s_a[1]) + a_lane;
  const uint32_t wh_u32 = edge::smem_addr(s_w[0]) + w_lane, wl_u32 = edge::smem_addr(s_w[1]) + w_lane;
#pragma unroll 1
  for (int k0 = 0; k0 < g.K; k0 += KC) {
    if (k0 > 0) __syncthreads();               // the previous chunk is no longer read
    stage();
    __syncthreads();
    if (k0 + KC < g.K) { load_w(k0 + KC); load_a(k0 + KC); }
#pragma unroll
    for (int s = 0; s < KC / 8; ++s) {
      uint32_t ah[4], al[4];
      edge::ldmatrix_x4(ah_u32 + s * 32, ah);
      edge::ldmatrix_x4(al_u32 + s * 32, al);
#pragma unroll
      for (int jp = 0; jp < NT / 2; ++jp) {
        uint32_t bh[4], bl[4];
        edge::ldmatrix_x4(wh_u32 + jp * 16 * ROWB + s * 32, bh);   // n-tiles 2jp (bh[0..1]) and 2jp+1 (bh[2..3])
        edge::ldmatrix_x4(wl_u32 + jp * 16 * ROWB + s * 32, bl);
        const uint32_t bh0[2] = {bh[0], bh[1]}, bh1[2] = {bh[2], bh[3]}, bl0[2] = {bl[0], bl[1]}, bl1[2] = {bl[2], bl[3]};
        mma_1688_tf32(acc[2 * jp], ah, bh0);     mma_1688_tf32(acc[2 * jp + 1], ah, bh1);
        mma_1688_tf32(acc[2 * jp], al, bh0);     mma_1688_tf32(acc[2 * jp + 1], al, bh1);
        mma_1688_tf32(acc[2 * jp], ah, bl0);     mma_1688_tf32(acc[2 * jp + 1], ah, bl1);
      }
    }
  }
  // C fragment: rows g and g + 8 of the warp's 16, columns 8j + 2q + {0, 1}
  const int gq = lane >> 2, q = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int m = m0 + warp * 16 + gq + half * 8;
    if (m >= g.M) continue;
    if (EPI == EPI_PIXSHUF) {
      // column n = quarter * (N/4) + channel: quarter (dy, dx) of the 2x2 up-sampled pixel (PixelShuffle(2))
      const int quarter = g.N >> 2, sp = g.sp;
      const int face = m / (sp * sp), rem = m - face * sp * sp;
      const int py = rem / sp, px = rem - py * sp;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int n = n0 + j * 8 + 2 * q;
        const int qd = n / quarter, kch = n - qd * quarter;
        const size_t orow = (static_cast<size_t>(face) * (2 * sp) + (2 * py + (qd >> 1))) * (2 * sp) + (2 * px + (qd & 1));
        float2* o = reinterpret_cast<float2*>(g.out + orow * g.ldo + kch);
        float2 v = *o;
        v.x += acc[j][2 * half];
        v.y += acc[j][2 * half + 1];
        *o = v;
      }
    } else {
      float* orow = g.out + static_cast<size_t>(m) * g.ldo + n0 + 2 * q;
      const float* rrow = EPI == EPI_RESID ? g.resid + static_cast<size_t>(m) * g.ldr + n0 + 2 * q : nullptr;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        float2 v = make_float2(acc[j][2 * half], acc[j][2 * half + 1]);
        if (g.bias != nullptr) {
          const float2 b = __ldg(reinterpret_cast<const float2*>(g.bias + n0 + j * 8 + 2 * q));
          v.x += b.x;
          v.y += b.y;
        }
        if (EPI == EPI_RESID) {
          const float2 r = *reinterpret_cast<const float2*>(rrow + j * 8);
          v.x += r.x;
          v.y += r.y;
        }
        *reinterpret_cast<float2*>(orow + j * 8) = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K = 32 / 64 (the 1x1 convs of the NAF blocks at c = 32 / 64: most of the pixels of the network): the 3xTF32 kernel
// above is bound by the MMA issue rate there (130 TFLOP/s of k8 instructions).  With the whole K extent of a row in
// one shared-memory stage the split can be fp16 instead — x * s = hi + lo, 11 + 11 mantissa bits, s the power of two
// that brings the ROW's maximum into [2^14, 2^15) so that fp16's exponent range never clips (anything lost to
// underflow is below 2^-39 of the row maximum); weights likewise with one scale per layer — and the MMAs are k16:
// half the instructions and half the ldmatrix traffic for the same 22 bits.  The accumulator is unscaled per row in
// the epilogue (powers of two: exact).
// ------------------------------------------------------------------------------------------------------------------
struct ArgsH {
  const float* A;       // [M, lda] fp32
  const __half* w_hi;   // [N, K] fp16 of w * 2^e
  const __half* w_lo;
  const float* bias;
  float* out;
  const float* resid;
  int lda, ldo, ldr;
  int M, N;
  float w_unscale;      // 2^-e
};

template <int BN, int K> constexpr int smem_bytes_h() { return 2 * (BM + BN) * (K * 2 + 16) + BM * 4; }

template <int BN, int K, int EPI>
__global__ void __launch_bounds__(256, 2) gemm_mma3h_kernel(const ArgsH g) {
  constexpr int NT = BN / 8, RB = K * 2 + 16;
  constexpr int LPR = K / 4;                    // lanes (float4) per row
  constexpr int AJ = BM * LPR / 256;            // float4 per thread
  constexpr int WCH = BN * K / 8;               // 16-byte chunks per weight matrix
  constexpr int WJ = (WCH + 255) / 256;
  extern __shared__ __align__(128) uint8_t s_raw[];
  uint8_t* s_a[2] = {s_raw, s_raw + BM * RB};
  uint8_t* s_w[2] = {s_raw + 2 * BM * RB, s_raw + 2 * BM * RB + BN * RB};
  float* s_inv = reinterpret_cast<float*>(s_raw + 2 * (BM + BN) * RB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  pdl_trigger();
#pragma unroll
  for (int j = 0; j < WJ; ++j) {                // weights: constants, before the dependency wait
    const int i = tid + 256 * j;
    if (i < WCH) {
      const size_t off = static_cast<size_t>(n0 + i / (K / 8)) * K + (i % (K / 8)) * 8;
      const uint32_t dst = (i / (K / 8)) * RB + (i % (K / 8)) * 16;
      *reinterpret_cast<uint4*>(s_w[0] + dst) = __ldg(reinterpret_cast<const uint4*>(g.w_hi + off));
      *reinterpret_cast<uint4*>(s_w[1] + dst) = __ldg(reinterpret_cast<const uint4*>(g.w_lo + off));
    }
  }
  pdl_wait();
  float4 areg[AJ];
#pragma unroll
  for (int j = 0; j < AJ; ++j) {
    const int i = tid + 256 * j, row = m0 + i / LPR;
    areg[j] = row < g.M ? *reinterpret_cast<const float4*>(g.A + static_cast<size_t>(row) * g.lda + (i % LPR) * 4)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int j = 0; j < AJ; ++j) {
    const int i = tid + 256 * j;
    float4 f = areg[j];
    float mx = fmaxf(fmaxf(fabsf(f.x), fabsf(f.y)), fmaxf(fabsf(f.z), fabsf(f.w)));
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));   // the row's LPR lanes are adjacent
    const uint32_t e = (__float_as_uint(mx) >> 23) & 0xFFu;
    const float sc = e >= 15u ? __uint_as_float((268u - e) << 23) : 1.f;
    if (i % LPR == 0) s_inv[i / LPR] = e >= 15u ? __uint_as_float((e - 14u) << 23) : 1.f;
    f.x *= sc; f.y *= sc; f.z *= sc; f.w *= sc;
    const float r0 = __half2float(__float2half_rn(f.x)), r1 = __half2float(__float2half_rn(f.y));
    const float r2 = __half2float(__float2half_rn(f.z)), r3 = __half2float(__float2half_rn(f.w));
    const uint32_t off = (i / LPR) * RB + (i % LPR) * 8;
    *reinterpret_cast<uint2*>(s_a[0] + off) = make_uint2(edge::pack_half2(f.x, f.y), edge::pack_half2(f.z, f.w));
    *reinterpret_cast<uint2*>(s_a[1] + off) = make_uint2(edge::pack_half2(f.x - r0, f.y - r1), edge::pack_half2(f.z - r2, f.w - r3));
  }
  __syncthreads();
  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const uint32_t a_lane = (warp * 16 + (lane & 15)) * RB + (lane >> 4) * 16;
  const uint32_t w_lane = ((lane >> 4) * 8 + (lane & 7)) * RB + ((lane >> 3) & 1) * 16;
  const uint32_t ah_u32 = edge::smem_addr(s_a[0]) + a_lane, al_u32 = edge::smem_addr(s_a[1]) + a_lane;
  const uint32_t wh_u32 = edge::smem_addr(s_w[0]) + w_lane, wl_u32 = edge::smem_addr(s_w[1]) + w_lane;
#pragma unroll
  for (int s = 0; s < K / 16; ++s) {
    uint32_t ah[4], al[4];
    edge::ldmatrix_x4(ah_u32 + s * 32, ah);
    edge::ldmatrix_x4(al_u32 + s * 32, al);
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) {
      uint32_t bh[4], bl[4];
      edge::ldmatrix_x4(wh_u32 + jp * 16 * RB + s * 32, bh);
      edge::ldmatrix_x4(wl_u32 + jp * 16 * RB + s * 32, bl);
      const uint32_t bh0[2] = {bh[0], bh[1]}, bh1[2] = {bh[2], bh[3]}, bl0[2] = {bl[0], bl[1]}, bl1[2] = {bl[2], bl[3]};
      edge::mma_16816_f16(acc[2 * jp], ah, bh0);     edge::mma_16816_f16(acc[2 * jp + 1], ah, bh1);
      edge::mma_16816_f16(acc[2 * jp], al, bh0);     edge::mma_16816_f16(acc[2 * jp + 1], al, bh1);
      edge::mma_16816_f16(acc[2 * jp], ah, bl0);     edge::mma_16816_f16(acc[2 * jp + 1], ah, bl1);
    }
  }
  const int gq = lane >> 2, q = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int rl = warp * 16 + gq + half * 8, m = m0 + rl;
    if (m >= g.M) continue;
    const float us = s_inv[rl] * g.w_unscale;
    float* orow = g.out + static_cast<size_t>(m) * g.ldo + n0 + 2 * q;
    const float* rrow = EPI == EPI_RESID ? g.resid + static_cast<size_t>(m) * g.ldr + n0 + 2 * q : nullptr;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      float2 v = make_float2(acc[j][2 * half] * us, acc[j][2 * half + 1] * us);
      if (g.bias != nullptr) {
        const float2 b = __ldg(reinterpret_cast<const float2*>(g.bias + n0 + j * 8 + 2 * q));
        v.x += b.x;
        v.y += b.y;
      }
      if (EPI == EPI_RESID) {
        const float2 r = *reinterpret_cast<const float2*>(rrow + j * 8);
        v.x += r.x;
        v.y += r.y;
      }
      *reinterpret_cast<float2*>(orow + j * 8) = v;
    }
  }
}

inline bool eligible_h(int M, int N, int K, int epi) {
  return M >= 1024 && (K == 32 || K == 64) && N % 32 == 0 && (epi == EPI_BIAS || epi == EPI_RESID);
}

template <int BN> constexpr int smem_bytes() { return 2 * (BM + BN) * ROWB; }

inline bool eligible(int M, int N, int K, int epi) {
  return M >= 1024 && K % KC == 0 && N % 32 == 0 && (epi == EPI_BIAS || epi == EPI_RESID || (epi == EPI_PIXSHUF && (N / 4) % 2 == 0));
}

}  // namespace mma3
}  // namespace hd
