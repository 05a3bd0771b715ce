// Persistent "level chain" kernel: a whole run of NAF blocks at a small-spatial level (M = faces * sp^2 <= ~1024
// rows) executed by ONE cooperative launch.  Every GEMM of such a level is a weight-streaming GEMM that one
// kernel launch cannot make efficient (fixed per-launch latency ~5 us against a 2-4 us mainloop), so the chain
// replaces kernel boundaries by grid barriers:
//
//   GEMM phase   each CTA computes (m-tile, n-tile, k-split) units with the same TMA -> smem ring ->
//                tcgen05.mma -> TMEM pipeline as gemm_tc_kernel and drains the fp32 accumulator tile straight to
//                an L2-resident partial-tile workspace (no epilogue math, no DSMEM)
//   grid barrier
//   fix-up phase row-parallel over all CTAs: sums the K-split partial tiles in fixed order (deterministic) and
//                applies what the reference does between two GEMMs — bias + SimpleGate, SCA scale,
//                bias + residual add + LayerNorm2d + AdaLN modulation — writing the next GEMM's bf16 operand
//   grid barrier
//
// The phase list, the TMA descriptors (global memory) and all buffers are prepared by the host once per batch
// size.  Reference arithmetic: models/denoiser/conditional_naf.py:108-136 at 1x1 spatial (depthwise 3x3 folded
// into conv1, pooled mean == the gated tensor itself).
#pragma once

#include "common.cuh"
#include "gemm_tc.cuh"

namespace hd {
namespace chain {

enum PhaseKind : int {
  PH_GEMM = 0,       // partial[unit] = A[m-tile, k-slice] * W[n-tile, k-slice]^T
  PH_FIX_GATE = 1,   // out_bf16[m, j] = (P[m, x1(j)] + b[x1]) * (P[m, x2(j)] + b[x2])      (gate-packed 128-col groups)
  PH_FIX_SCALE = 2,  // out_bf16[m, n] = in_bf16[m, n] * (P[m, n] + b[n])                   (SCA, 1x1 level)
  PH_FIX_RESID = 3,  // x[m, n] += P[m, n] + b[n];  optionally out_bf16[m, :] = LNmod(x[m, :])
};

struct Phase {
  int kind;
  // GEMM
  int map_a, map_w;      // indices into the tensor-map array
  int n_tiles, num_kb, split;
  // fix-up (describes the GEMM whose partials it consumes)
  int N;                 // columns of the partial matrix (packed width for PH_FIX_GATE)
  int src_n_tiles, src_split;
  const float* bias;
  float* x;              // fp32 residual stream [rows, N]
  bf16* out;             // bf16 output [rows, N] (N/2 for gate)
  const bf16* in;        // PH_FIX_SCALE multiplicand
  const float* ln_w;     // PH_FIX_RESID: nullptr = no LayerNorm
  const float* ln_b;
  int shift_off, scale_off;
};

struct ChainArgs {
  const Phase* phases;
  int n_phases;
  const CUtensorMap* maps;
  float* partial;        // [units][128][128] fp32
  unsigned int* barrier; // grid barrier counter, zeroed before the launch
  int rows;              // valid rows (faces at the 1x1 level)
  int m_tiles;
  const float* mod_table;
  const int* mod_row_idx;
  int mod_stride, rows_per_face;
  DeviceStatus* status;
};

constexpr int STAGES = 6;
constexpr int EW = 8;
constexpr int THREADS = 64 + 32 * EW;
constexpr int STAGE_BYTES = 2 * tc::BM * tc::BK * 2;  // A 128x64 + W 128x64 bf16
constexpr int RING_BYTES = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = RING_BYTES + 512 + 1024;

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// All CTAs of the (cooperative, co-resident) grid.  `target` is advanced identically by every thread.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, DeviceStatus* st) {
  fence_proxy_async_all();  // this thread's generic-proxy global writes vs. later TMA (async proxy) reads
  __syncwarp();             // the single-lane producer / MMA loops rejoin their warps
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    unsigned int spins = 0;
    while (ld_acquire_gpu(counter) < target) {
      if ((++spins & 0x3FFu) == 0u) {
        if (*reinterpret_cast<volatile unsigned int*>(&st->error) != 0u) break;
        if (clock64() - t0 > 4000000000ll) {
          if (atomicCAS(&st->error, 0u, 2u) == 0u) st->where = 0x900u;
          break;
        }
      }
    }
    __threadfence();
  }
  __syncthreads();
  fence_proxy_async_all();
}

__global__ void __launch_bounds__(THREADS, 1) chain_kernel(const ChainArgs args) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);
  float* s_red = reinterpret_cast<float*>(bar_base + 256);  // [2][16] block-reduction scratch

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    mbar_init(smem_u32(tmem_empty_bar), EW);  // one arrival per epilogue warp after draining
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 128);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  unsigned int bar_target = 0;
  uint32_t ring_it = 0;  // k-blocks pushed through the ring so far (same sequence in producer and MMA thread)
  uint32_t acc_it = 0;   // accumulator tiles produced so far

  for (int ph = 0; ph < args.n_phases; ++ph) {
    const Phase P = args.phases[ph];
    if (P.kind == PH_GEMM) {
      const int units = args.m_tiles * P.n_tiles * P.split;
      const int kb_count = P.num_kb / P.split;
      const CUtensorMap* mapA = args.maps + P.map_a;
      const CUtensorMap* mapW = args.maps + P.map_w;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        // unit -> (split, n-tile, m-tile); m fastest so CTAs sharing a W tile run together
        const int mt = u % args.m_tiles;
        const int nt = (u / args.m_tiles) % P.n_tiles;
        const int ks = u / (args.m_tiles * P.n_tiles);
        const int kb_begin = ks * kb_count;
        if (warp == 0) {
          if (lane == 0) {
            for (int i = 0; i < kb_count; ++i) {
              const uint32_t it = ring_it + i;
              const int s = it % STAGES;
              const uint32_t phs = (it / STAGES) & 1;
              mbar_wait(smem_u32(&empty_bar[s]), phs ^ 1u, args.status, 0x910u);
              const uint32_t fb = smem_u32(&full_bar[s]);
              const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
              mbar_expect_tx(fb, STAGE_BYTES);
              tma_load_2d(sa, mapA, (kb_begin + i) * BK, mt * BM, fb);
              tma_load_2d(sa + BM * BK * 2, mapW, (kb_begin + i) * BK, nt * 128, fb);
            }
          }
        } else if (warp == 1) {
          if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, 128);
            // the previous accumulator must be drained before it is overwritten
            mbar_wait(smem_u32(tmem_empty_bar), (acc_it & 1u) ^ 1u, args.status, 0x920u);
            tc_fence_after_sync();
            for (int i = 0; i < kb_count; ++i) {
              const uint32_t it = ring_it + i;
              const int s = it % STAGES;
              const uint32_t phs = (it / STAGES) & 1;
              mbar_wait(smem_u32(&full_bar[s]), phs, args.status, 0x930u);
              tc_fence_after_sync();
              const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
              const uint64_t da = make_smem_desc(sa);
              const uint64_t db = make_smem_desc(sa + BM * BK * 2);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(da + 2 * k, db + 2 * k, tmem_base, (i | k) != 0 ? 1u : 0u, idesc);
              umma_commit(smem_u32(&empty_bar[s]));
            }
            umma_commit(smem_u32(tmem_full_bar));
          }
        } else {
          // drain: TMEM lane = row; each thread writes 128-byte row segments of the partial tile
          const int quad = warp & 3;
          mbar_wait(smem_u32(tmem_full_bar), acc_it & 1u, args.status, 0x940u);
          tc_fence_after_sync();
          const int r = quad * 32 + lane;
          const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
          float* prow = args.partial + (static_cast<size_t>(u) * 128 + r) * 128;
          const int cbeg = ((warp - 2) >> 2) * 64;
#pragma unroll 1
          for (int c0 = cbeg; c0 < cbeg + 64; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr_row + c0, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(prow + c0 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tmem_empty_bar)) : "memory");
          }
        }
        ring_it += kb_count;
        acc_it += 1;
      }
    } else {
      // ---------------- fix-up phases: one row at a time per CTA, 256 worker threads, float4 chunks ----------------
      const int t = threadIdx.x;
      const int chunks_out = (P.kind == PH_FIX_GATE ? P.N / 2 : P.N) / 4;   // float4 chunks of the output row
      for (int row = blockIdx.x; row < args.rows; row += gridDim.x) {
        const int mt = row >> 7, rin = row & 127;
        float4 keep[2];   // PH_FIX_RESID keeps the finished row in registers for the LayerNorm (N <= 2048)
        float lsum = 0.f;
        for (int rep = 0; rep < 2; ++rep) {
          const int j = rep * 256 + t;
          keep[rep] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t >= 256 || j >= chunks_out) continue;
          // partial chunk(s) feeding output chunk j
          int pc = j;
          if (P.kind == PH_FIX_GATE) pc = (j >> 4) * 32 + (j & 15);
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          {
            const int nt = pc >> 5, cin = (pc & 31) * 4;
            for (int s = 0; s < P.src_split; ++s) {
              const size_t u = (static_cast<size_t>(s) * P.src_n_tiles + nt) * args.m_tiles + mt;
              const float4 v = __ldcg(reinterpret_cast<const float4*>(args.partial + (u * 128 + rin) * 128 + cin));
              a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
              if (P.kind == PH_FIX_GATE) {
                const float4 w = __ldcg(reinterpret_cast<const float4*>(args.partial + (u * 128 + rin) * 128 + cin + 64));
                b.x += w.x; b.y += w.y; b.z += w.z; b.w += w.w;
              }
            }
          }
          const float4 bi = __ldg(reinterpret_cast<const float4*>(P.bias + pc * 4));
          a.x += bi.x; a.y += bi.y; a.z += bi.z; a.w += bi.w;
          if (P.kind == PH_FIX_GATE) {
            const float4 b2 = __ldg(reinterpret_cast<const float4*>(P.bias + pc * 4 + 64));
            a.x *= b.x + b2.x; a.y *= b.y + b2.y; a.z *= b.z + b2.z; a.w *= b.w + b2.w;
            tc::store4<bf16>(P.out + static_cast<size_t>(row) * (P.N / 2) + j * 4, a);
          } else if (P.kind == PH_FIX_SCALE) {
            const uint2 g = *reinterpret_cast<const uint2*>(P.in + static_cast<size_t>(row) * P.N + j * 4);
            const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y);
            a.x *= g0.x; a.y *= g0.y; a.z *= g1.x; a.w *= g1.y;
            tc::store4<bf16>(P.out + static_cast<size_t>(row) * P.N + j * 4, a);
          } else {  // PH_FIX_RESID
            float* xp = P.x + static_cast<size_t>(row) * P.N + j * 4;
            const float4 xv = *reinterpret_cast<const float4*>(xp);
            a.x += xv.x; a.y += xv.y; a.z += xv.z; a.w += xv.w;
            *reinterpret_cast<float4*>(xp) = a;
            keep[rep] = a;
            lsum += a.x + a.y + a.z + a.w;
          }
        }
        if (P.kind == PH_FIX_RESID && P.ln_w != nullptr) {
          // LayerNorm2d over the finished row (two-pass, utils.py:16-24) + AdaLN modulation -> bf16
          lsum = warp_sum(lsum);
          if (lane == 0) s_red[warp] = lsum;
          __syncthreads();
          float tot = 0.f;
#pragma unroll
          for (int i = 0; i < THREADS / 32; ++i) tot += s_red[i];
          const float mu = tot / static_cast<float>(P.N);
          float ss = 0.f;
          for (int rep = 0; rep < 2; ++rep) {
            const int j = rep * 256 + t;
            if (t >= 256 || j >= chunks_out) continue;
            const float d0 = keep[rep].x - mu, d1 = keep[rep].y - mu, d2 = keep[rep].z - mu, d3 = keep[rep].w - mu;
            ss += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          }
          ss = warp_sum(ss);
          if (lane == 0) s_red[16 + warp] = ss;
          __syncthreads();
          float vt = 0.f;
#pragma unroll
          for (int i = 0; i < THREADS / 32; ++i) vt += s_red[16 + i];
          const float denom = sqrtf(vt / static_cast<float>(P.N) + 1e-6f);
          const float* mrow = args.mod_table + static_cast<size_t>(args.mod_row_idx[row / args.rows_per_face]) * args.mod_stride;
          for (int rep = 0; rep < 2; ++rep) {
            const int j = rep * 256 + t;
            if (t >= 256 || j >= chunks_out) continue;
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(P.ln_w + j * 4));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.ln_b + j * 4));
            const float4 sc = __ldg(reinterpret_cast<const float4*>(mrow + P.scale_off + j * 4));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(mrow + P.shift_off + j * 4));
            float4 y;
            y.x = (w4.x * ((keep[rep].x - mu) / denom) + b4.x) * (sc.x + 1.f) + sh.x;
            y.y = (w4.y * ((keep[rep].y - mu) / denom) + b4.y) * (sc.y + 1.f) + sh.y;
            y.z = (w4.z * ((keep[rep].z - mu) / denom) + b4.z) * (sc.z + 1.f) + sh.z;
            y.w = (w4.w * ((keep[rep].w - mu) / denom) + b4.w) * (sc.w + 1.f) + sh.w;
            tc::store4<bf16>(P.out + static_cast<size_t>(row) * P.N + j * 4, y);
          }
          __syncthreads();  // s_red is reused by the next row
        }
      }
    }
    grid_barrier(args.barrier, bar_target, args.status);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace chain
}  // namespace hd
