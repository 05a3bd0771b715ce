// Simplified channel attention, fused: s = W_sca . pooled + b per face, then g <- g * s on every pixel row of the
// face (conditional_naf.py:54-65,119), in ONE launch.
//
// Why not the tcgen05 GEMM: this is a skinny GEMM (M = faces = 256, N = K = c) in the middle of a dependent chain.
// On the tcgen05 kernel it costs 8 us — TMEM allocation, tensor-map fetch, an 8-way split-K cluster and its DSMEM
// reduction, for 0.5 GFLOP — plus a separate 2 us launch for the rescale.  Latency is what matters here, not
// tensor throughput, so this kernel uses warp-level mma.sync (m16n8k16, bf16 in, fp32 accumulate) fed by a
// 3-stage cp.async ring: no TMEM, no cluster, 64x32 output tiles so the grid covers the chip at every level
// (c/32 x faces/64 CTAs), and the CTA that owns an s tile rescales the matching columns of its faces' rows.
//
//   A      [faces][c]  bf16   pooled means (or, at the 1x1 level, the gated tensor itself)
//   W      [c][c]      bf16   K-major, as packed for the tcgen05 GEMM
//   g      [faces*rpf][c] bf16 gated tensor;  out = g * s  (out of place: at the 1x1 level A aliases g)
//   s_out  [faces][c]  fp32   optional copy of the scale
#pragma once
#include "common.cuh"

namespace hd {
namespace sca {

constexpr int BM = 64, BN = 32, BK = 128, STAGES = 3, THREADS = 256;
constexpr int LDS = BK + 8;                         // padded smem row (elements): ldmatrix conflict-free
constexpr int STAGE_ELEMS = (BM + BN) * LDS;
constexpr int SMEM_BYTES = STAGES * STAGE_ELEMS * 2;  // 78336

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(THREADS) sca_scale_kernel(const bf16* __restrict__ A, const bf16* __restrict__ W,
                                                            const float* __restrict__ bias, const bf16* __restrict__ g,
                                                            bf16* __restrict__ out, float* __restrict__ s_out, int faces,
                                                            int c, int rpf) {
  extern __shared__ __align__(16) uint8_t sca_smem[];
  bf16* smem = reinterpret_cast<bf16*>(sca_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int nchunks = c / BK;
  const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  pdl_trigger();

  // 16-byte copies of one k-chunk: W tile BN x BK (512 copies), A tile BM x BK (1024 copies)
  auto load_w = [&](int kc, int stage) {
#pragma unroll
    for (int i = 0; i < BN * (BK / 8) / THREADS; ++i) {
      const int idx = tid + i * THREADS, r = idx >> 4, ck = idx & 15;
      cp_async16(smem_base + ((stage * STAGE_ELEMS + (BM + r) * LDS + ck * 8) << 1),
                 W + static_cast<size_t>(n0 + r) * c + kc * BK + ck * 8);
    }
  };
  auto load_a = [&](int kc, int stage) {
#pragma unroll
    for (int i = 0; i < BM * (BK / 8) / THREADS; ++i) {
      const int idx = tid + i * THREADS, r = idx >> 4, ck = idx & 15;
      cp_async16(smem_base + ((stage * STAGE_ELEMS + r * LDS + ck * 8) << 1),
                 A + static_cast<size_t>(m0 + r) * c + kc * BK + ck * 8);
    }
  };
  // weights are constants: their first chunks are in flight before the dependency wait
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s)
    if (s < nchunks) load_w(s, s);
  pdl_wait();
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nchunks) load_a(s, s);
    cp_commit();
  }

  const int mw = warp & 3;   // 16-row slice of the tile
  const int kg = warp >> 2;  // k-group: k16 steps [kg*4, kg*4 + 4) of each chunk
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  // ldmatrix source rows / k offsets of this lane (A: 16x16 tile; B: two n8 tiles x k16 per x4)
  const int a_row = mw * 16 + (lane & 15), a_k = (lane >> 4) * 8;
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_k = ((lane >> 3) & 1) * 8;

  for (int kc = 0; kc < nchunks; ++kc) {
    cp_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kc + STAGES - 1;
      if (nk < nchunks) { load_w(nk, nk % STAGES); load_a(nk, nk % STAGES); }
      cp_commit();
    }
    const int stage = kc % STAGES;
    const uint32_t sa = smem_base + ((stage * STAGE_ELEMS) << 1);
    const uint32_t sb = sa + ((BM * LDS) << 1);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int k0 = (kg * 4 + ks) * 16;
      uint32_t af[4], b01[4], b23[4];
      ldmatrix_x4(sa + ((a_row * LDS + k0 + a_k) << 1), af);
      ldmatrix_x4(sb + ((b_row * LDS + k0 + b_k) << 1), b01);          // n-tiles 0, 1
      ldmatrix_x4(sb + (((16 + b_row) * LDS + k0 + b_k) << 1), b23);   // n-tiles 2, 3
      mma_bf16(acc[0], af, b01[0], b01[1]);
      mma_bf16(acc[1], af, b01[2], b01[3]);
      mma_bf16(acc[2], af, b23[0], b23[1]);
      mma_bf16(acc[3], af, b23[2], b23[3]);
    }
  }
  cp_wait<0>();
  __syncthreads();  // the ring is dead: reuse it for the k-group reduction and the s tile

  float* red = reinterpret_cast<float*>(sca_smem);   // [BM][BN + 1]
  float* stile = red + BM * (BN + 1);                // [BM][BN + 1]
  const int r0 = mw * 16 + (lane >> 2), cq = (lane & 3) * 2;
  if (kg == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[r0 * (BN + 1) + j * 8 + cq] = acc[j][0];
      red[r0 * (BN + 1) + j * 8 + cq + 1] = acc[j][1];
      red[(r0 + 8) * (BN + 1) + j * 8 + cq] = acc[j][2];
      red[(r0 + 8) * (BN + 1) + j * 8 + cq + 1] = acc[j][3];
    }
  }
  __syncthreads();
  if (kg == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = r0 + (e >> 1) * 8, col = j * 8 + cq + (e & 1);
        const float v = acc[j][e] + red[r * (BN + 1) + col] + bias[n0 + col];
        stile[r * (BN + 1) + col] = v;
        if (s_out != nullptr && m0 + r < faces) s_out[static_cast<size_t>(m0 + r) * c + n0 + col] = v;
      }
    }
  }
  __syncthreads();
  // rescale: 64 faces x rpf rows x 32 channels (4 chunks of 8 bf16)
  const int total = BM * rpf * (BN / 8);
  for (int idx = tid; idx < total; idx += THREADS) {
    const int ck = idx & 3, r = idx >> 2;
    const int fl = r / rpf, face = m0 + fl;
    if (face >= faces) continue;
    const size_t off = (static_cast<size_t>(face) * rpf + (r - fl * rpf)) * c + n0 + ck * 8;
    float v[8];
    load8(g + off, v);
    const float* sv = stile + fl * (BN + 1) + ck * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= sv[e];
    store8(out + off, v);
  }
}

}  // namespace sca
}  // namespace hd
