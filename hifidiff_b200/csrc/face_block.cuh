// Fused ConditionalNAFBlock kernel for the 16x16 level (c = 128): ONE CTA runs one face (256 pixels) through a
// whole run of consecutive blocks, and nothing but the weights is read from memory in between.
//
//   residual stream x   fp32, lives in TENSOR MEMORY (2 m-tiles x 128 columns) for the whole kernel: conv3 and
//                       conv5 are issued with accumulate=1 straight onto it, so the residual add is free
//   GEMM operands       A (LayerNorm / gate output, bf16) is written by the CTA's own threads into shared memory
//                       in the K-major SWIZZLE_128B layout tcgen05.mma expects; W tiles arrive by TMA
//   conv1 -> dw3x3      the conv1 accumulator is drained to a bf16 [plane][pixel][128] tile in shared memory; the
//                       depthwise 3x3, SimpleGate and the SCA pool run out of it (a face is a whole image, so the
//                       zero padding is the face border and no halo leaves the CTA)
//   SCA                 per-face mean -> 128x128 GEMV on CUDA cores -> rescale of the gated operand
//
// Reference arithmetic: models/denoiser/conditional_naf.py:108-136 (block), utils.py:16-24 (LayerNorm2d),
// utils.py:57-60 (SimpleGate); beta/gamma are folded into conv3/conv5, conv4 is gate-packed (hd_weights.inl).
//
// Shared memory (232000 B):  A 64 KB | T 128 KB (two planes; also parking space for W1 / W4) | W3/W5 32 KB | misc
// Tensor memory (512 cols):  x m-tile 0 | x m-tile 1 | accumulator m-tile 0 | accumulator m-tile 1
// Threads: 8 worker warps (thread <-> pixel row <-> TMEM lane); thread 0 also issues the TMA loads and MMAs.
#pragma once

#include "common.cuh"
#include "gemm_tc.cuh"

namespace hd {
namespace fb {

constexpr int C = 128;
constexpr int SP = 16;
constexpr int PX = SP * SP;
constexpr int WORKERS = 256;
constexpr int THREADS = WORKERS;          // thread 0 doubles as the controller (TMA + MMA issue): 8 warps keep 255 regs/thread
constexpr int TILE = 16384;               // one 128-row x 64-col bf16 operand tile
constexpr int A_OFF = 0;                  // [m-tile 2][k-block 2] tiles
constexpr int T_OFF = 4 * TILE;           // plane 0 | plane 1, each [256 px][128 ch] bf16
constexpr int PLANE = 4 * TILE;
constexpr int W3_OFF = T_OFF + 2 * PLANE; // W3 / W5: [k-block 2] tiles
constexpr int BAR_OFF = W3_OFF + 2 * TILE;
constexpr int SCR_OFF = BAR_OFF + 64;     // 2 KB scratch: pool partials [4][128] | mean [128] + LN params [2][128]
constexpr int S_OFF = SCR_OFF + 2048;     // SCA scale [128]
constexpr int SMEM_BYTES = S_OFF + 512;
constexpr int MAX_BLOCKS = 4;
constexpr uint32_t X_COL = 0, ACC_COL = 256;

struct BlockParams {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const float *b1, *dw_w, *dw_b;   // conv1 bias [256]; depthwise taps [9][256] and bias [256]
  const float *wsca_t, *bsca;      // SCA weight transposed [k][n] fp32, bias
  const float *b4;                 // gate-packed conv4 bias
  const float *cb3, *cb5;          // cumulative residual bias after this block's conv3 / conv5 (the stream in TMEM is bias-free)
  int mod_off, pad;
};

struct Args {
  const CUtensorMap* maps;         // [n_blocks][4]: w1 [256,128], w3 [128,128], w4 (gate-packed) [256,128], w5 [128,128]
  const BlockParams* blocks;
  int n_blocks;
  float* x;                        // residual stream [faces * 256, 128] fp32, updated in place
  const float* mod_table;
  const int* mod_row_idx;
  int mod_stride;
  const float* zero_bias;          // 128 zeros (no bias before the first conv3)
  DeviceStatus* status;
  long long* trace;                // optional: clock64 stamps of one CTA's phase boundaries (diagnostics)
  int trace_cta;
  int first_wave;                  // CTAs [0, first_wave) are the first on their SM: they warm L2 and the instruction cache
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
// packed pairs of fp32 for the 2-wide FMA pipe (fma.rn.f32x2)
__device__ __forceinline__ uint64_t pack_f2(float2 v) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
  return r;
}
__device__ __forceinline__ float2 unpack_f2(uint64_t r) {
  float2 v;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
  return v;
}
__device__ __forceinline__ uint64_t bf2_to_f2(uint32_t u) {  // (bf16 lo, bf16 hi) -> (f32, f32)
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(u << 16), "r"(u & 0xffff0000u));
  return r;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
               : "memory");
}

// ---- code-size discipline --------------------------------------------------------------------------------------
// These fused kernels are long straight-line programs; measured on B200, the first execution of every code region
// after the rest of the step has flushed the instruction caches costs about as much as the work itself (the SM
// fetches ~2-3 B of instructions per clock from L2).  So everything that is called from several places and does not
// carry register arrays across the call lives in ONE non-inlined copy.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, DeviceStatus* st, uint32_t site) {
  tc::mbar_wait(bar, parity, st, site);
}
__device__ __forceinline__ void mbar_wait_c(uint32_t bar, uint32_t parity, DeviceStatus* st, uint32_t site) {
  if (!tc::mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, st, site);
}
// one 64-wide k-block: D[tmem] (+)= A tile (128 x 64) * B tile (N x 64)^T, four UMMA_K steps
__device__ __noinline__ void issue_kblock(uint32_t a_tile, uint32_t b_tile, uint32_t tmem_d, uint32_t accumulate_first,
                                          uint32_t idesc) {
  const uint64_t da = tc::make_smem_desc(a_tile), db = tc::make_smem_desc(b_tile);
#pragma unroll
  for (int k = 0; k < tc::BK / tc::UMMA_K; ++k)
    tc::umma_bf16(da + 2 * k, db + 2 * k, tmem_d, k == 0 ? accumulate_first : 1u, idesc);
}

// the controller lane / sub-warp branches rejoin their warps before the CTA barrier
__device__ __forceinline__ void block_sync() {
  __syncwarp();
  __syncthreads();
}

// byte offset inside the A operand of (pixel row R, 16-byte chunk q of the 256-byte channel row)
__device__ __forceinline__ uint32_t a_chunk_off(int R, int q) {
  const int mt = R >> 7, r = R & 127;
  return static_cast<uint32_t>(((mt * 2 + (q >> 3)) * TILE) + r * 128 + (((q & 7) ^ (r & 7)) << 4));
}
// byte offset inside one T plane of (pixel px, 16-byte chunk q of the 256-byte channel row)
__device__ __forceinline__ uint32_t t_chunk_off(int px, int q) {
  return static_cast<uint32_t>(px * 256 + ((q ^ (px & 7)) << 4));
}

// LayerNorm2d over the 128 channels of one pixel (two-pass, fp32) with the AdaLN modulation folded into
// (eff_w, eff_b), written as the bf16 A operand row R.
__device__ __forceinline__ void ln_row_to_a(const float (&v)[C], const float* eff_w, const float* eff_b, uint32_t a_base,
                                            int R) {
  float sa[8];  // 8 independent partial sums: a 16-deep dependency chain instead of 128
#pragma unroll
  for (int i = 0; i < 8; ++i) sa[i] = v[i];
#pragma unroll
  for (int i = 8; i < C; ++i) sa[i & 7] += v[i];
  const float mu = (((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]))) * (1.f / C);
#pragma unroll
  for (int i = 0; i < 8; ++i) sa[i] = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const float d = v[i] - mu;
    sa[i & 7] = fmaf(d, d, sa[i & 7]);
  }
  const float ss = ((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]));
  const float rstd = 1.f / sqrtf(ss * (1.f / C) + 1e-6f);
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float4 w0 = *reinterpret_cast<const float4*>(eff_w + q * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(eff_w + q * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(eff_b + q * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(eff_b + q * 8 + 4);
    const float y0 = (v[q * 8 + 0] - mu) * rstd * w0.x + b0.x, y1 = (v[q * 8 + 1] - mu) * rstd * w0.y + b0.y;
    const float y2 = (v[q * 8 + 2] - mu) * rstd * w0.z + b0.z, y3 = (v[q * 8 + 3] - mu) * rstd * w0.w + b0.w;
    const float y4 = (v[q * 8 + 4] - mu) * rstd * w1.x + b1.x, y5 = (v[q * 8 + 5] - mu) * rstd * w1.y + b1.y;
    const float y6 = (v[q * 8 + 6] - mu) * rstd * w1.z + b1.z, y7 = (v[q * 8 + 7] - mu) * rstd * w1.w + b1.w;
    sts128(a_base + a_chunk_off(R, q), pack_bf16x2(y0, y1), pack_bf16x2(y2, y3), pack_bf16x2(y4, y5), pack_bf16x2(y6, y7));
  }
}

// eff_w = w (1 + scale), eff_b = b (1 + scale) + shift   (conditional_naf.py:24-25 applied to LayerNorm2d's affine)
__device__ __forceinline__ void make_ln_params(float* eff, const float* lw, const float* lb, const float* mrow, int shift_off,
                                               int scale_off, int c) {
  const float sc = 1.f + __ldg(mrow + scale_off + c);
  eff[c] = __ldg(lw + c) * sc;
  eff[C + c] = __ldg(lb + c) * sc + __ldg(mrow + shift_off + c);
}

// residual row R (= x_tmem + cbias) -> LayerNorm2d + modulation -> bf16 A operand row; one copy for every call site
__device__ __noinline__ void residual_ln(uint32_t t_x, int R, const float* __restrict__ cbias, const float* eff, uint32_t sA) {
  using namespace tc;
  float v[C];
  {
    uint32_t r[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld32(t_x + c * 32, r[c]);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 bb = cbias != nullptr ? __ldg(reinterpret_cast<const float4*>(cbias + c * 32 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[c * 32 + i] = __uint_as_float(r[c][i]) + bb.x; v[c * 32 + i + 1] = __uint_as_float(r[c][i + 1]) + bb.y;
        v[c * 32 + i + 2] = __uint_as_float(r[c][i + 2]) + bb.z; v[c * 32 + i + 3] = __uint_as_float(r[c][i + 3]) + bb.w;
      }
  }
  ln_row_to_a(v, eff, eff + C, sA, R);
}

__global__ void __launch_bounds__(THREADS, 1) face_block_kernel(const Args args) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sA = sbase + A_OFF, sT = sbase + T_OFF, sW3 = sbase + W3_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [0..3] weight tiles w1,w3,w4,w5; [4] MMA done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 48);
  float* scr = reinterpret_cast<float*>(smem + SCR_OFF);
  float* eff = scr + 256;                                         // LN params: eff_w [128], eff_b [128]
  float* s_sca = reinterpret_cast<float*>(smem + S_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool ctrl = tid == 0;
  const int face = blockIdx.x;
  const uint32_t wbar0 = smem_u32(&bars[0]), mma_bar = smem_u32(&bars[4]);
  const int nb = args.n_blocks;
  int n_stamp = 0;
  auto stamp = [&]() {
    if (args.trace != nullptr && blockIdx.x == args.trace_cta && tid == 0) args.trace[n_stamp] = clock64();
    ++n_stamp;
  };

  pdl_trigger();
  stamp();
  if (tid == 0) {
    if ((sbase & 1023u) != 0u) {
      if (atomicCAS(&args.status->error, 0u, 3u) == 0u) args.status->where = 0xA00u;
    }
    for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  block_sync();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  auto load_w = [&](int map_idx, uint32_t dst, int halves, int bar) {  // controller only
    const CUtensorMap* m = args.maps + map_idx;
    const uint32_t b = wbar0 + bar * 8;
    mbar_expect_tx(b, halves * 2 * TILE);
    for (int h = 0; h < halves; ++h)
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(dst + (h * 2 + kb) * TILE, m, kb * BK, h * 128, b);
  };
  constexpr uint32_t idesc = make_idesc(128, 128);
  auto issue = [&](uint32_t w_base, uint32_t d_col, uint32_t accumulate) {  // controller only: D[2 m-tiles] (+)= A W^T
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      issue_kblock(sA + (mt * 2) * TILE, w_base, tmem_base + d_col + mt * 128, accumulate, idesc);
      issue_kblock(sA + (mt * 2 + 1) * TILE, w_base + TILE, tmem_base + d_col + mt * 128, 1u, idesc);
    }
    umma_commit(mma_bar);
  };

  // weights of the first block are constants: fetch them before waiting for the predecessor kernel
  if (ctrl) {
    load_w(0, sT + PLANE, 2, 0);  // W1 parks in plane 1 (plane 0 is written while its second half is still needed)
    load_w(1, sW3, 1, 1);
    // everything else this kernel will need is pulled towards L2 now (cold after the rest of the step)
    for (int i = 2; i < 4 * nb && blockIdx.x < args.first_wave; ++i) {
      const int halves = (i & 1) ? 1 : 2;
      for (int h = 0; h < halves; ++h)
        for (int kb = 0; kb < 2; ++kb) tma_prefetch_2d(args.maps + i, kb * BK, h * 128);
    }
  }
  for (int b = 0; b < nb && blockIdx.x < args.first_wave; ++b) {
    const BlockParams bp = args.blocks[b];
    prefetch_l2(bp.wsca_t + tid * 32);
    prefetch_l2(bp.wsca_t + (256 + tid) * 32);
    if (tid < 72) prefetch_l2(bp.dw_w + tid * 32);
    if (tid >= 96 && tid < 104) prefetch_l2(bp.b1 + (tid - 96) * 32);
    if (tid >= 104 && tid < 112) prefetch_l2(bp.b4 + (tid - 104) * 32);
    if (tid >= 112 && tid < 120) prefetch_l2(bp.dw_b + (tid - 112) * 32);
    if (tid >= 128 && tid < 132) prefetch_l2(bp.cb3 + (tid - 128) * 32);
    if (tid >= 132 && tid < 136) prefetch_l2(bp.cb5 + (tid - 132) * 32);
    if (tid >= 136 && tid < 140) prefetch_l2(bp.bsca + (tid - 136) * 32);
    if (tid >= 140 && tid < 144) prefetch_l2(bp.ln1_w + (tid - 140) * 32);
    if (tid >= 144 && tid < 148) prefetch_l2(bp.ln1_b + (tid - 144) * 32);
    if (tid >= 148 && tid < 152) prefetch_l2(bp.ln2_w + (tid - 148) * 32);
    if (tid >= 152 && tid < 156) prefetch_l2(bp.ln2_b + (tid - 152) * 32);
  }
  // worker geometry
  const int R = tid;                                  // pixel row (workers)
  const int quad = warp & 3, mt_own = (warp >> 2) & 1;
  const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
  const uint32_t t_x = tmem_base + lane_addr + X_COL + mt_own * 128;
  const uint32_t t_acc = tmem_base + lane_addr + ACC_COL + mt_own * 128;
  if (static_cast<int>(blockIdx.x) < args.first_wave) {
    // The first CTA on an SM finds the instruction cache cold (the other 200-odd kernels of the step have been
    // through it): the first LayerNorm cost 16 k clocks instead of 4 k.  Run it once on whatever the A region and
    // tensor memory hold while the predecessor kernel is still finishing; its output is overwritten below.
    residual_ln(t_x, R, nullptr, eff, sA);
    block_sync();
  }
  pdl_wait();
  stamp();

  const float* mrow = args.mod_table + static_cast<size_t>(__ldg(args.mod_row_idx + face)) * args.mod_stride;
  // depthwise geometry: warp -> (channel block, 4-row strip), lane -> channel pair
  const int cb = warp & 1, strip = (warp >> 1) & 3, j = cb * 64 + lane * 2;
  // fp32 row staging for the coalesced load / store of x: rows 0..127 in the A region, 128..255 in T plane 0;
  // 16-byte chunks XOR-swizzled by row so that both the row-per-warp and the row-per-thread side are conflict-free
  auto stage_row = [&](int r) { return (r < 128 ? sA : sT - 128 * 512) + static_cast<uint32_t>(r) * 512u; };

  uint32_t mph = 0;  // parity of the MMA-done barrier

  {
    const BlockParams bp = args.blocks[0];
    if (tid < C) make_ln_params(eff, bp.ln1_w, bp.ln1_b, mrow, bp.mod_off, bp.mod_off + C, tid);
    {
      const float* xw = args.x + (static_cast<size_t>(face) * PX + warp * 32) * C + lane * 4;
      float4 t[32];  // all 32 rows of this warp in flight at once
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = *reinterpret_cast<const float4*>(xw + static_cast<size_t>(i) * C);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int r = warp * 32 + i;
        sts128(stage_row(r) + ((lane ^ (r & 7)) << 4), __float_as_uint(t[i].x), __float_as_uint(t[i].y), __float_as_uint(t[i].z),
               __float_as_uint(t[i].w));
      }
    }
    stamp();
    block_sync();
    {
      const uint32_t srow = stage_row(R);
#pragma unroll 1
      for (int c0 = 0; c0 < 4; ++c0) {
        uint32_t r[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 f = lds128(srow + (((c0 * 8 + q) ^ (R & 7)) << 4));
          r[4 * q] = __float_as_uint(f.x); r[4 * q + 1] = __float_as_uint(f.y); r[4 * q + 2] = __float_as_uint(f.z); r[4 * q + 3] = __float_as_uint(f.w);
        }
        tmem_st32(t_x + c0 * 32, r);
      }
      tmem_wait_st();
    }
    block_sync();  // every row has left the staging area before the A operand is written over it
    stamp();
    residual_ln(t_x, R, nullptr, eff, sA);
  }

  for (int b = 0; b < nb; ++b) {
    stamp();  // A ready (norm1)
    const BlockParams bp = args.blocks[b];
    const bool last = b + 1 == nb;
    const uint32_t wpar = static_cast<uint32_t>(b & 1);

    // ---------------- conv1: two 128-column halves through the accumulator, drained to the T planes ----------------
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      block_sync();
      if (ctrl) {
        tc_fence_after_sync();
        if (h == 0) mbar_wait_c(wbar0, wpar, args.status, 0xA10u);
        issue(sT + PLANE + h * 2 * TILE, ACC_COL, 0u);
      }
      mbar_wait_c(mma_bar, mph, args.status, 0xA11u);
      tc_fence_after_sync();
      {
        const uint32_t prow = sT + h * PLANE + static_cast<uint32_t>(R) * 256u;
        uint32_t r[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(t_acc + c * 32, r[c]);
        tmem_wait_ld();
        const float4* bias = reinterpret_cast<const float4*>(bp.b1 + h * 128);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4 b0 = __ldg(bias + 2 * q), b1 = __ldg(bias + 2 * q + 1);
          const uint32_t* rr = &r[q >> 2][(q & 3) * 8];
          sts128(prow + ((q ^ (R & 7)) << 4), pack_bf16x2(__uint_as_float(rr[0]) + b0.x, __uint_as_float(rr[1]) + b0.y),
                 pack_bf16x2(__uint_as_float(rr[2]) + b0.z, __uint_as_float(rr[3]) + b0.w),
                 pack_bf16x2(__uint_as_float(rr[4]) + b1.x, __uint_as_float(rr[5]) + b1.y),
                 pack_bf16x2(__uint_as_float(rr[6]) + b1.z, __uint_as_float(rr[7]) + b1.w));
        }
      }
      mph ^= 1u;
    }
    block_sync();  // T complete
    stamp();

    // ---------------- depthwise 3x3 + bias + SimpleGate -> A operand, pool partials ----------------
    {
      uint64_t wk1[9], wk2[9];  // taps of this lane's (x1a, x1b) and (x2a, x2b) channel pairs, packed for FFMA2
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        wk1[t] = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_w + t * 256 + j)));
        wk2[t] = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_w + t * 256 + 128 + j)));
      }
      const uint64_t bz1 = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_b + j)));
      const uint64_t bz2 = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_b + 128 + j)));
      uint32_t lx[8], ax[8];    // swizzle terms by (x & 7): load side (T planes) and store side (A operand)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        lx[k] = static_cast<uint32_t>((((j >> 3) ^ k) << 4) + (j & 7) * 2);
        ax[k] = static_cast<uint32_t>((((lane >> 2) ^ k) << 4) + (lane & 3) * 4);
      }
      float ps0 = 0.f, ps1 = 0.f;
      // two image rows at a time: 4 input rows per column feed 2 output pixels (fewer loads, 4 independent FMA chains)
#pragma unroll 1
      for (int y0 = strip * 4; y0 < strip * 4 + 4; y0 += 2) {
        const bool up = y0 > 0, dn = y0 + 2 < SP;
        const uint32_t trow = sT + static_cast<uint32_t>(y0 * SP) * 256u;                       // pixel (y0, 0) in plane 0
        const uint32_t arow = sA + static_cast<uint32_t>(((y0 >> 3) * 2 + cb) * TILE + (y0 & 7) * SP * 128);
        uint64_t w1v[3][4], w2v[3][4];  // [column slot][input row y0-1 .. y0+2]
        auto load_col = [&](int x, uint64_t (&c1)[4], uint64_t (&c2)[4]) {
          const uint32_t a = trow + lx[x & 7] + x * 256;
          c1[0] = up ? bf2_to_f2(lds32(a - SP * 256)) : 0ull;
          c2[0] = up ? bf2_to_f2(lds32(a - SP * 256 + PLANE)) : 0ull;
          c1[1] = bf2_to_f2(lds32(a));
          c2[1] = bf2_to_f2(lds32(a + PLANE));
          c1[2] = bf2_to_f2(lds32(a + SP * 256));
          c2[2] = bf2_to_f2(lds32(a + SP * 256 + PLANE));
          c1[3] = dn ? bf2_to_f2(lds32(a + 2 * SP * 256)) : 0ull;
          c2[3] = dn ? bf2_to_f2(lds32(a + 2 * SP * 256 + PLANE)) : 0ull;
        };
#pragma unroll
        for (int dy = 0; dy < 4; ++dy) w1v[0][dy] = w2v[0][dy] = 0ull;
        load_col(0, w1v[1], w2v[1]);
#pragma unroll
        for (int x = 0; x < SP; ++x) {
          uint64_t (&l1)[4] = w1v[x % 3], (&l2)[4] = w2v[x % 3];
          uint64_t (&m1)[4] = w1v[(x + 1) % 3], (&m2)[4] = w2v[(x + 1) % 3];
          uint64_t (&r1)[4] = w1v[(x + 2) % 3], (&r2)[4] = w2v[(x + 2) % 3];
          if (x + 1 < SP) {
            load_col(x + 1, r1, r2);
          } else {
#pragma unroll
            for (int dy = 0; dy < 4; ++dy) r1[dy] = r2[dy] = 0ull;
          }
          uint64_t a1 = bz1, a2 = bz2, b1 = bz1, b2 = bz2;  // a: output row y0, b: output row y0 + 1
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            a1 = ffma2(wk1[dy * 3 + 0], l1[dy], a1); a2 = ffma2(wk2[dy * 3 + 0], l2[dy], a2);
            b1 = ffma2(wk1[dy * 3 + 0], l1[dy + 1], b1); b2 = ffma2(wk2[dy * 3 + 0], l2[dy + 1], b2);
            a1 = ffma2(wk1[dy * 3 + 1], m1[dy], a1); a2 = ffma2(wk2[dy * 3 + 1], m2[dy], a2);
            b1 = ffma2(wk1[dy * 3 + 1], m1[dy + 1], b1); b2 = ffma2(wk2[dy * 3 + 1], m2[dy + 1], b2);
            a1 = ffma2(wk1[dy * 3 + 2], r1[dy], a1); a2 = ffma2(wk2[dy * 3 + 2], r2[dy], a2);
            b1 = ffma2(wk1[dy * 3 + 2], r1[dy + 1], b1); b2 = ffma2(wk2[dy * 3 + 2], r2[dy + 1], b2);
          }
          const float2 fa1 = unpack_f2(a1), fa2 = unpack_f2(a2), fb1 = unpack_f2(b1), fb2 = unpack_f2(b2);
          const float ga0 = fa1.x * fa2.x, ga1 = fa1.y * fa2.y, gb0 = fb1.x * fb2.x, gb1 = fb1.y * fb2.y;
          ps0 += ga0 + gb0; ps1 += ga1 + gb1;
          sts32(arow + ax[x & 7] + x * 128, pack_bf16x2(ga0, ga1));
          sts32(arow + ax[x & 7] + x * 128 + SP * 128, pack_bf16x2(gb0, gb1));
        }
      }
      *reinterpret_cast<float2*>(scr + strip * C + j) = make_float2(ps0, ps1);
    }
    // SCA weights for this thread's half of the GEMV: fetched before the barrier, consumed after it
    const int sn = tid & (C - 1), skh = tid >> 7;
    float wv[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) wv[i] = __ldg(bp.wsca_t + (skh * 64 + i) * C + sn);
    block_sync();  // T free; partials visible
    stamp();
    if (ctrl) {
      load_w(b * 4 + 2, sT, 2, 2);                           // W4 -> plane 0
      if (!last) load_w((b + 1) * 4 + 0, sT + PLANE, 2, 0);  // next block's W1 -> plane 1
    }
    // ---------------- SCA: s = Wsca mean + b ----------------
    if (tid < C) scr[tid] = (scr[tid] + scr[C + tid] + scr[2 * C + tid] + scr[3 * C + tid]) * (1.f / PX);
    block_sync();
    {
      float a0 = skh == 0 ? __ldg(bp.bsca + sn) : 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float4 m = *reinterpret_cast<const float4*>(scr + skh * 64 + i);
        a0 = fmaf(wv[i], m.x, a0); a1 = fmaf(wv[i + 1], m.y, a1); a2 = fmaf(wv[i + 2], m.z, a2); a3 = fmaf(wv[i + 3], m.w, a3);
      }
      // halves: k < 64 -> s_sca, k >= 64 -> scr[128..255] (strip-1 partials, dead since the mean was taken)
      (skh == 0 ? s_sca : scr + C)[sn] = (a0 + a1) + (a2 + a3);
      if (tid >= C) make_ln_params(eff, bp.ln2_w, bp.ln2_b, mrow, bp.mod_off + 2 * C, bp.mod_off + 3 * C, tid - C);
    }
    block_sync();
    {  // rescale this thread's own gated values
      const float s0 = s_sca[j] + scr[C + j], s1 = s_sca[j + 1] + scr[C + j + 1];
      const uint32_t abase = sA + static_cast<uint32_t>(((strip >> 1) * 2 + cb) * TILE + (strip & 1) * 64 * 128) +
                             static_cast<uint32_t>((lane & 3) * 4);
#pragma unroll 8
      for (int i = 0; i < 64; ++i) {
        const uint32_t a = abase + i * 128 + (((lane >> 2) ^ (i & 7)) << 4);
        const float2 g = unpack_bf16x2(lds32(a));
        sts32(a, pack_bf16x2(g.x * s0, g.y * s1));
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
    }
    block_sync();
    stamp();

    // ---------------- conv3 (+beta) accumulated onto x; then x += b3, norm2 + modulation -> A ----------------
    if (ctrl) {
      tc_fence_after_sync();
      mbar_wait_c(wbar0 + 8, wpar, args.status, 0xA20u);
      issue(sW3, X_COL, 1u);
      mbar_wait_c(mma_bar, mph, args.status, 0xA21u);
      load_w(b * 4 + 3, sW3, 1, 3);  // W5 takes W3's place
    }
    mbar_wait_c(mma_bar, mph, args.status, 0xA22u);
    tc_fence_after_sync();
    residual_ln(t_x, R, bp.cb3, eff, sA);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    mph ^= 1u;
    block_sync();
    stamp();

    // ---------------- conv4 + SimpleGate: first half kept in registers until the second half's MMAs are done ----------------
    uint32_t hold[32];
    if (ctrl) {
      tc_fence_after_sync();
      mbar_wait_c(wbar0 + 16, wpar, args.status, 0xA30u);
      issue(sT, ACC_COL, 0u);
    }
    if (!last && tid < C) {
      const BlockParams nx = args.blocks[b + 1];
      make_ln_params(eff, nx.ln1_w, nx.ln1_b, mrow, nx.mod_off, nx.mod_off + C, tid);
    }
    auto gate_half = [&](int h, uint32_t (&out)[32]) {  // 64 gated values of this row as 32 bf16x2
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(t_acc + c * 32, r[c]);
      tmem_wait_ld();
      const float4* bias = reinterpret_cast<const float4*>(bp.b4 + h * 128);
#pragma unroll
      for (int i = 0; i < 16; ++i) {  // gated values 4i .. 4i+3
        const float4 p = __ldg(bias + i), q = __ldg(bias + 16 + i);
        const uint32_t* x1 = &r[i >> 3][(i & 7) * 4];
        const uint32_t* x2 = &r[2 + (i >> 3)][(i & 7) * 4];
        out[2 * i] = pack_bf16x2((__uint_as_float(x1[0]) + p.x) * (__uint_as_float(x2[0]) + q.x),
                                 (__uint_as_float(x1[1]) + p.y) * (__uint_as_float(x2[1]) + q.y));
        out[2 * i + 1] = pack_bf16x2((__uint_as_float(x1[2]) + p.z) * (__uint_as_float(x2[2]) + q.z),
                                     (__uint_as_float(x1[3]) + p.w) * (__uint_as_float(x2[3]) + q.w));
      }
    };
    auto store_half = [&](int h, const uint32_t (&in)[32]) {  // gated channels h*64 .. h*64+63 = k-block h of the A row
#pragma unroll
      for (int q = 0; q < 8; ++q) sts128(sA + a_chunk_off(R, h * 8 + q), in[4 * q], in[4 * q + 1], in[4 * q + 2], in[4 * q + 3]);
    };
    mbar_wait_c(mma_bar, mph, args.status, 0xA31u);
    tc_fence_after_sync();
    gate_half(0, hold);
    tc_fence_before_sync();
    mph ^= 1u;
    block_sync();
    if (ctrl) {
      tc_fence_after_sync();
      issue(sT + 2 * TILE, ACC_COL, 0u);
    }
    mbar_wait_c(mma_bar, mph, args.status, 0xA32u);
    tc_fence_after_sync();
    {
      uint32_t g2[32];
      gate_half(1, g2);
      store_half(0, hold);
      store_half(1, g2);
      fence_proxy_async_smem();
      tc_fence_before_sync();
    }
    mph ^= 1u;
    block_sync();
    stamp();

    // ---------------- conv5 (+gamma) accumulated onto x; x += b5; next block's norm1 or the final store ----------------
    if (ctrl) {
      tc_fence_after_sync();
      mbar_wait_c(wbar0 + 24, wpar, args.status, 0xA40u);
      issue(sW3, X_COL, 1u);
      if (!last) {
        mbar_wait_c(mma_bar, mph, args.status, 0xA41u);
        load_w((b + 1) * 4 + 1, sW3, 1, 1);  // next block's W3
      }
    }
    mbar_wait_c(mma_bar, mph, args.status, 0xA42u);
    tc_fence_after_sync();
    if (!last) {
      residual_ln(t_x, R, bp.cb5, eff, sA);
    } else {
      // rows -> swizzled staging (A region / plane 0 are dead: every MMA has completed) -> coalesced row stores
      const uint32_t srow = stage_row(R);
#pragma unroll 1
      for (int c0 = 0; c0 < 4; ++c0) {
        uint32_t r[32];
        tmem_ld32(t_x + c0 * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bp.cb5 + c0 * 32 + q * 4));
          sts128(srow + (((c0 * 8 + q) ^ (R & 7)) << 4), __float_as_uint(__uint_as_float(r[4 * q]) + bb.x),
                 __float_as_uint(__uint_as_float(r[4 * q + 1]) + bb.y), __float_as_uint(__uint_as_float(r[4 * q + 2]) + bb.z),
                 __float_as_uint(__uint_as_float(r[4 * q + 3]) + bb.w));
        }
      }
    }
    mph ^= 1u;
  }
  block_sync();
  {
    float* xw = args.x + (static_cast<size_t>(face) * PX + warp * 32) * C + lane * 4;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const int r = warp * 32 + i;
      *reinterpret_cast<float4*>(xw + static_cast<size_t>(i) * C) = lds128(stage_row(r) + ((lane ^ (r & 7)) << 4));
    }
  }

  tc_fence_before_sync();
  block_sync();
  stamp();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fb
}  // namespace hd
