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
// utils.py:57-60 (SimpleGate); beta/gamma are folded into conv3/conv5, conv4 is gate-packed (hd_lib.cu).
//
// Shared memory (232000 B):  A 64 KB | T 128 KB (two planes; also parking space for W1 / W4) | W3/W5 32 KB | misc
// Tensor memory (512 cols):  x m-tile 0 | x m-tile 1 | accumulator m-tile 0 | accumulator m-tile 1
// Threads: 8 worker warps (thread <-> pixel row <-> TMEM lane) + 1 controller warp (TMA + MMA issue).
#pragma once

#include "common.cuh"
#include "gemm_tc.cuh"

namespace hd {
namespace fb {

constexpr int C = 128;
constexpr int SP = 16;
constexpr int PX = SP * SP;
constexpr int WORKERS = 256;
constexpr int THREADS = WORKERS + 32;
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
  const float *b3, *b4, *b5;       // beta*b3, gate-packed b4, gamma*b5
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
  DeviceStatus* status;
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
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) s += v[i];
  const float mu = s * (1.f / C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const float d = v[i] - mu;
    ss += d * d;
  }
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
  const bool worker = warp < 8;
  const bool ctrl = warp == 8 && lane == 0;
  const int face = blockIdx.x;
  const uint32_t wbar0 = smem_u32(&bars[0]), mma_bar = smem_u32(&bars[4]);
  const int nb = args.n_blocks;

  pdl_trigger();
  if (tid == 0) {
    if ((sbase & 1023u) != 0u) {
      if (atomicCAS(&args.status->error, 0u, 3u) == 0u) args.status->where = 0xA00u;
    }
    for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 8) {
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
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t da = make_smem_desc(sA + (mt * 2 + kb) * TILE);
        const uint64_t db = make_smem_desc(w_base + kb * TILE);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(da + 2 * k, db + 2 * k, tmem_base + d_col + mt * 128, accumulate | static_cast<uint32_t>((kb | k) != 0), idesc);
      }
    }
    umma_commit(mma_bar);
  };

  // weights of the first block are constants: fetch them before waiting for the predecessor kernel
  if (ctrl) {
    load_w(0, sT + PLANE, 2, 0);  // W1 parks in plane 1 (plane 0 is written while its second half is still needed)
    load_w(1, sW3, 1, 1);
  }
  pdl_wait();

  // worker geometry
  const int R = tid;                                  // pixel row (workers)
  const int quad = warp & 3, mt_own = (warp >> 2) & 1;
  const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
  const uint32_t t_x = tmem_base + lane_addr + X_COL + mt_own * 128;
  const uint32_t t_acc = tmem_base + lane_addr + ACC_COL + mt_own * 128;
  const float* mrow = args.mod_table + static_cast<size_t>(__ldg(args.mod_row_idx + face)) * args.mod_stride;
  // depthwise geometry: warp -> (channel block, 4-row strip), lane -> channel pair
  const int cb = warp & 1, strip = (warp >> 1) & 3, j = cb * 64 + lane * 2;

  uint32_t mph = 0;  // parity of the MMA-done barrier
  float v[C];        // this thread's pixel row of the residual stream (workers)

  {
    const BlockParams bp = args.blocks[0];
    if (tid < C) make_ln_params(eff, bp.ln1_w, bp.ln1_b, mrow, bp.mod_off, bp.mod_off + C, tid);
    if (worker) {
      const float* xr = args.x + (static_cast<size_t>(face) * PX + R) * C;
#pragma unroll
      for (int i = 0; i < C / 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(xr + 4 * i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[c0 + i]);
        tmem_st32(t_x + c0, r);
      }
    }
    block_sync();
    if (worker) {
      ln_row_to_a(v, eff, eff + C, sA, R);
      tmem_wait_st();
    }
  }

  for (int b = 0; b < nb; ++b) {
    const BlockParams bp = args.blocks[b];
    const bool last = b + 1 == nb;
    const uint32_t wpar = static_cast<uint32_t>(b & 1);

    // ---------------- conv1: two 128-column halves through the accumulator, drained to the T planes ----------------
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      if (worker) {
        fence_proxy_async_smem();
        tc_fence_before_sync();
      }
      block_sync();
      if (ctrl) {
        tc_fence_after_sync();
        if (h == 0) mbar_wait(wbar0, wpar, args.status, 0xA10u);
        issue(sT + PLANE + h * 2 * TILE, ACC_COL, 0u);
      }
      if (worker) {
        mbar_wait(mma_bar, mph, args.status, 0xA11u);
        tc_fence_after_sync();
        const uint32_t plane = sT + h * PLANE;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_acc + c0, r);
          tmem_wait_ld();
          const float* bias = bp.b1 + h * 128 + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t p[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              p[e] = pack_bf16x2(__uint_as_float(r[q * 8 + 2 * e]) + __ldg(bias + q * 8 + 2 * e),
                                 __uint_as_float(r[q * 8 + 2 * e + 1]) + __ldg(bias + q * 8 + 2 * e + 1));
            sts128(plane + t_chunk_off(R, (c0 >> 3) + q), p[0], p[1], p[2], p[3]);
          }
        }
      }
      mph ^= 1u;
    }
    block_sync();  // T complete

    // ---------------- depthwise 3x3 + bias + SimpleGate -> A operand, pool partials ----------------
    if (worker) {
      float wk[9][4], bz[4];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(bp.dw_w + t * 256 + j));
        const float2 c2 = __ldg(reinterpret_cast<const float2*>(bp.dw_w + t * 256 + 128 + j));
        wk[t][0] = a.x; wk[t][1] = a.y; wk[t][2] = c2.x; wk[t][3] = c2.y;
      }
      {
        const float2 a = __ldg(reinterpret_cast<const float2*>(bp.dw_b + j));
        const float2 c2 = __ldg(reinterpret_cast<const float2*>(bp.dw_b + 128 + j));
        bz[0] = a.x; bz[1] = a.y; bz[2] = c2.x; bz[3] = c2.y;
      }
      float ps0 = 0.f, ps1 = 0.f;
      const int jq = j >> 3, jin = (j & 7) * 2;
#pragma unroll 1
      for (int y = strip * 4; y < strip * 4 + 4; ++y) {
        float win[3][3][4];  // [column slot][dy][x1a, x1b, x2a, x2b]
        auto load_col = [&](int x, float (&col)[3][4]) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int yy = y + dy - 1;
            if (yy >= 0 && yy < SP) {
              const int px = yy * SP + x;
              const uint32_t off = t_chunk_off(px, jq) + jin;
              const float2 a = unpack_bf16x2(lds32(sT + off));
              const float2 c2 = unpack_bf16x2(lds32(sT + PLANE + off));
              col[dy][0] = a.x; col[dy][1] = a.y; col[dy][2] = c2.x; col[dy][3] = c2.y;
            } else {
              col[dy][0] = col[dy][1] = col[dy][2] = col[dy][3] = 0.f;
            }
          }
        };
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int e = 0; e < 4; ++e) win[0][dy][e] = 0.f;
        load_col(0, win[1]);
#pragma unroll
        for (int x = 0; x < SP; ++x) {
          float (&cl)[3][4] = win[x % 3];
          float (&cm)[3][4] = win[(x + 1) % 3];
          float (&cr)[3][4] = win[(x + 2) % 3];
          if (x + 1 < SP) {
            load_col(x + 1, cr);
          } else {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
              for (int e = 0; e < 4; ++e) cr[dy][e] = 0.f;
          }
          float acc[4] = {bz[0], bz[1], bz[2], bz[3]};
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[e] = fmaf(wk[dy * 3 + 0][e], cl[dy][e], acc[e]);
              acc[e] = fmaf(wk[dy * 3 + 1][e], cm[dy][e], acc[e]);
              acc[e] = fmaf(wk[dy * 3 + 2][e], cr[dy][e], acc[e]);
            }
          const float g0 = acc[0] * acc[2], g1 = acc[1] * acc[3];
          ps0 += g0; ps1 += g1;
          const int Rp = y * SP + x;
          sts32(sA + a_chunk_off(Rp, cb * 8 + (lane >> 2)) + (lane & 3) * 4, pack_bf16x2(g0, g1));
        }
      }
      *reinterpret_cast<float2*>(scr + strip * C + j) = make_float2(ps0, ps1);
    }
    block_sync();  // T free; partials visible
    if (ctrl) {
      load_w(b * 4 + 2, sT, 2, 2);                           // W4 -> plane 0
      if (!last) load_w((b + 1) * 4 + 0, sT + PLANE, 2, 0);  // next block's W1 -> plane 1
    }
    // ---------------- SCA: s = Wsca mean + b ----------------
    if (tid < C) scr[tid] = (scr[tid] + scr[C + tid] + scr[2 * C + tid] + scr[3 * C + tid]) * (1.f / PX);
    block_sync();
    if (tid < C) {
      float a0 = __ldg(bp.bsca + tid), a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
      for (int k = 0; k < C; k += 4) {
        a0 = fmaf(__ldg(bp.wsca_t + (k + 0) * C + tid), scr[k + 0], a0);
        a1 = fmaf(__ldg(bp.wsca_t + (k + 1) * C + tid), scr[k + 1], a1);
        a2 = fmaf(__ldg(bp.wsca_t + (k + 2) * C + tid), scr[k + 2], a2);
        a3 = fmaf(__ldg(bp.wsca_t + (k + 3) * C + tid), scr[k + 3], a3);
      }
      s_sca[tid] = (a0 + a1) + (a2 + a3);
    } else if (tid < 2 * C) {
      make_ln_params(eff, bp.ln2_w, bp.ln2_b, mrow, bp.mod_off + 2 * C, bp.mod_off + 3 * C, tid - C);
    }
    block_sync();
    if (worker) {  // rescale this thread's own gated values
      const float s0 = s_sca[j], s1 = s_sca[j + 1];
#pragma unroll 4
      for (int i = 0; i < 64; ++i) {
        const int Rp = strip * 64 + i;
        const uint32_t a = sA + a_chunk_off(Rp, cb * 8 + (lane >> 2)) + (lane & 3) * 4;
        const float2 g = unpack_bf16x2(lds32(a));
        sts32(a, pack_bf16x2(g.x * s0, g.y * s1));
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
    }
    block_sync();

    // ---------------- conv3 (+beta) accumulated onto x; then x += b3, norm2 + modulation -> A ----------------
    if (ctrl) {
      tc_fence_after_sync();
      mbar_wait(wbar0 + 8, wpar, args.status, 0xA20u);
      issue(sW3, X_COL, 1u);
      mbar_wait(mma_bar, mph, args.status, 0xA21u);
      load_w(b * 4 + 3, sW3, 1, 3);  // W5 takes W3's place
    }
    if (worker) {
      mbar_wait(mma_bar, mph, args.status, 0xA22u);
      tc_fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_x + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[c0 + i] = __uint_as_float(r[i]) + __ldg(bp.b3 + c0 + i);
          r[i] = __float_as_uint(v[c0 + i]);
        }
        tmem_st32(t_x + c0, r);
      }
      ln_row_to_a(v, eff, eff + C, sA, R);
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before_sync();
    }
    mph ^= 1u;
    block_sync();

    // ---------------- conv4 + SimpleGate: first half kept in registers until the second half's MMAs are done ----------------
    uint32_t hold[32];
    if (ctrl) {
      tc_fence_after_sync();
      mbar_wait(wbar0 + 16, wpar, args.status, 0xA30u);
      issue(sT, ACC_COL, 0u);
    }
    if (!last && tid < C) {
      const BlockParams nx = args.blocks[b + 1];
      make_ln_params(eff, nx.ln1_w, nx.ln1_b, mrow, nx.mod_off, nx.mod_off + C, tid);
    }
    auto gate_half = [&](int h, uint32_t (&out)[32]) {  // 64 gated values of this row as 32 bf16x2
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        uint32_t r1[32], r2[32];
        tmem_ld32(t_acc + part * 32, r1);
        tmem_ld32(t_acc + 64 + part * 32, r2);
        tmem_wait_ld();
        const float* bias = bp.b4 + h * 128 + part * 32;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float g0 = (__uint_as_float(r1[2 * i]) + __ldg(bias + 2 * i)) * (__uint_as_float(r2[2 * i]) + __ldg(bias + 64 + 2 * i));
          const float g1 = (__uint_as_float(r1[2 * i + 1]) + __ldg(bias + 2 * i + 1)) *
                           (__uint_as_float(r2[2 * i + 1]) + __ldg(bias + 64 + 2 * i + 1));
          out[part * 16 + i] = pack_bf16x2(g0, g1);
        }
      }
    };
    auto store_half = [&](int h, const uint32_t (&in)[32]) {  // gated channels h*64 .. h*64+63 = k-block h of the A row
#pragma unroll
      for (int q = 0; q < 8; ++q) sts128(sA + a_chunk_off(R, h * 8 + q), in[4 * q], in[4 * q + 1], in[4 * q + 2], in[4 * q + 3]);
    };
    if (worker) {
      mbar_wait(mma_bar, mph, args.status, 0xA31u);
      tc_fence_after_sync();
      gate_half(0, hold);
      tc_fence_before_sync();
    }
    mph ^= 1u;
    block_sync();
    if (ctrl) {
      tc_fence_after_sync();
      issue(sT + 2 * TILE, ACC_COL, 0u);
    }
    if (worker) {
      mbar_wait(mma_bar, mph, args.status, 0xA32u);
      tc_fence_after_sync();
      uint32_t g2[32];
      gate_half(1, g2);
      store_half(0, hold);
      store_half(1, g2);
      fence_proxy_async_smem();
      tc_fence_before_sync();
    }
    mph ^= 1u;
    block_sync();

    // ---------------- conv5 (+gamma) accumulated onto x; x += b5; next block's norm1 or the final store ----------------
    if (ctrl) {
      tc_fence_after_sync();
      mbar_wait(wbar0 + 24, wpar, args.status, 0xA40u);
      issue(sW3, X_COL, 1u);
      if (!last) {
        mbar_wait(mma_bar, mph, args.status, 0xA41u);
        load_w((b + 1) * 4 + 1, sW3, 1, 1);  // next block's W3
      }
    }
    if (worker) {
      mbar_wait(mma_bar, mph, args.status, 0xA42u);
      tc_fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_x + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[c0 + i] = __uint_as_float(r[i]) + __ldg(bp.b5 + c0 + i);
          r[i] = __float_as_uint(v[c0 + i]);
        }
        if (!last) tmem_st32(t_x + c0, r);
      }
      if (!last) {
        ln_row_to_a(v, eff, eff + C, sA, R);
        tmem_wait_st();
      } else {
        float* xr = args.x + (static_cast<size_t>(face) * PX + R) * C;
#pragma unroll
        for (int i = 0; i < C / 4; ++i)
          *reinterpret_cast<float4*>(xr + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
    mph ^= 1u;
  }

  tc_fence_before_sync();
  block_sync();
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fb
}  // namespace hd
