// Fused ConditionalNAFBlock kernel for the 8x8 level (c = 256): ONE CTA runs TWO faces (2 x 64 pixels = one
// 128-row MMA tile) through a run of consecutive blocks.  Same idea as face_block.cuh (residual stream in tensor
// memory, operands built in shared memory, depthwise 3x3 / SimpleGate / SCA on the CTA's own data), but the block's
// weights (768 KB bf16) no longer fit next to the activations, so they stream through five 32 KB shared-memory slots
// whose lifetimes are scheduled by hand around the places where the conv1 tile is not live:
//
//   smem   A operand 64 KB | T = P0..P3 (conv1 output planes [128 px][128 ch] bf16, 32 KB each) | R 32 KB | misc
//   slots  P0..P3 and R hold weight tiles whenever the plane / scratch they alias is dead:
//            conv1  q0 (P2,P3)  q1 (P1,R)  q2 (P2,P3)  q3 (R,P3)       drains: quarter q -> plane q
//            conv3  k-block kb -> P(kb)                                 (T is dead after the depthwise conv)
//            conv4  q0 (P0,P1)  q1 (P2,P3)  q2 (R,P0)   q3 (P2,P3)
//            conv5  kb0 P1, kb1 P0, kb2 P2, kb3 P3
//          R also carries the per-face LayerNorm parameters and the SCA scratch while no weights are parked in it.
//   TMEM   x (residual stream, 256 fp32 columns, bias-free: the conv3 / conv5 biases are constants per channel and
//          are carried as a cumulative vector added on read) | accumulator 0 | accumulator 1 (128 columns each)
//   threads 256: thread <-> (pixel row r = TMEM lane, column half hf); thread 0 also issues the TMA loads and MMAs.
//
// Reference arithmetic: models/denoiser/conditional_naf.py:108-136, utils.py:16-24,57-60.
#pragma once

#include "common.cuh"
#include "face_block.cuh"
#include "gemm_tc.cuh"

namespace hd {
namespace pb {

using fb::bf2_to_f2;
using fb::block_sync;
using fb::ffma2;
using fb::lds128;
using fb::lds32;
using fb::pack_f2;
using fb::prefetch_l2;
using fb::sts128;
using fb::sts32;
using fb::unpack_f2;

constexpr int C = 256;
constexpr int SP = 8;
constexpr int FPX = SP * SP;              // pixels per face
constexpr int ROWS = 2 * FPX;             // rows per CTA
constexpr int THREADS = 256;
constexpr int TILE = 16384;               // 128 rows x 64 bf16
constexpr int SLOT = 2 * TILE;
constexpr int A_OFF = 0;                  // 4 k-block tiles
constexpr int P_OFF = 4 * TILE;           // planes / slots P0..P3
constexpr int R_OFF = P_OFF + 4 * SLOT;   // slot R
constexpr int BAR_OFF = R_OFF + SLOT;
constexpr int SMEM_BYTES = BAR_OFF + 128;
constexpr int MAX_BLOCKS = 4;
constexpr uint32_t X_COL = 0, ACC_COL = 256;
// scratch inside R (valid only while no weights are parked there)
constexpr int R_EFF = 0;                  // [face 2][w|b][256] fp32 = 4 KB
constexpr int R_MEAN = 4096;              // [face 2][256] fp32 = 2 KB
constexpr int R_PART = 6144;              // [k-half 2][face 2][256] fp32 = 4 KB
constexpr int R_XCHG = 12288;             // [half 2][row 128] (mean, M2) = 2 KB

struct BlockParams {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const float *b1, *dw_w, *dw_b;   // conv1 bias [512]; depthwise taps [9][512] and bias [512]
  const bf16* wsca_t;              // SCA weight transposed [k][n] bf16
  const float *bsca, *b4;          // SCA bias; gate-packed conv4 bias [512]
  const float *cb3, *cb5;          // cumulative residual bias after this block's conv3 / conv5 (see header)
  int mod_off, pad;
};

struct Args {
  const CUtensorMap* maps;         // [n_blocks][4]: w1 [512,256], w3 [256,256], w4 (gate-packed) [512,256], w5 [256,256]
  const BlockParams* blocks;
  int n_blocks, n_faces;
  float* x;                        // residual stream [faces * 64, 256] fp32, updated in place
  const float* mod_table;
  const int* mod_row_idx;
  int mod_stride;
  const float* zero_bias;          // 256 zeros (the stream carries no bias before the first conv3)
  DeviceStatus* status;
  long long* trace;                // optional: clock64 stamps of one CTA's phase boundaries (diagnostics)
  int trace_cta;
  int warm;                        // run the LayerNorm code once before the dependency wait (cold instruction cache)
};

// LayerNorm2d + AdaLN modulation of residual row r (= x_tmem + cbias).  The two threads of a row each hold one
// 128-channel half in registers, take its mean / centred second moment (two-pass), swap them through shared memory
// (one CTA barrier) and combine them exactly (Chan et al.); each then writes its half as k-blocks 2hf, 2hf+1 of the
// bf16 A operand.  Must be called by all threads of the CTA.
__device__ __noinline__ void residual_ln(uint32_t t_own, int hf, int r, const float* __restrict__ cbias_own, const float* eff_w,
                                         const float* eff_b, uint32_t sA, float2* xchg) {
  using namespace tc;
  float v[128];
  {
    uint32_t t[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld32(t_own + c * 32, t[c]);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 bb = cbias_own != nullptr ? __ldg(reinterpret_cast<const float4*>(cbias_own + c * 32 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[c * 32 + i] = __uint_as_float(t[c][i]) + bb.x; v[c * 32 + i + 1] = __uint_as_float(t[c][i + 1]) + bb.y;
        v[c * 32 + i + 2] = __uint_as_float(t[c][i + 2]) + bb.z; v[c * 32 + i + 3] = __uint_as_float(t[c][i + 3]) + bb.w;
      }
  }
  float sa[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sa[i] = v[i];
#pragma unroll
  for (int i = 8; i < 128; ++i) sa[i & 7] += v[i];
  const float mean_h = (((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]))) * (1.f / 128);
#pragma unroll
  for (int i = 0; i < 8; ++i) sa[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 128; ++i) {
    const float d = v[i] - mean_h;
    sa[i & 7] = fmaf(d, d, sa[i & 7]);
  }
  const float m2_h = ((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]));
  xchg[hf * 128 + r] = make_float2(mean_h, m2_h);
  block_sync();
  const float2 p = xchg[(1 - hf) * 128 + r];
  const float mu = 0.5f * (mean_h + p.x), dm = mean_h - p.x;
  const float rstd = 1.f / sqrtf((m2_h + p.y + dm * dm * 64.f) * (1.f / C) + 1e-6f);
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float4 w0 = *reinterpret_cast<const float4*>(eff_w + q * 8), w1 = *reinterpret_cast<const float4*>(eff_w + q * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(eff_b + q * 8), b1 = *reinterpret_cast<const float4*>(eff_b + q * 8 + 4);
    const float y0 = (v[q * 8 + 0] - mu) * rstd * w0.x + b0.x, y1 = (v[q * 8 + 1] - mu) * rstd * w0.y + b0.y;
    const float y2 = (v[q * 8 + 2] - mu) * rstd * w0.z + b0.z, y3 = (v[q * 8 + 3] - mu) * rstd * w0.w + b0.w;
    const float y4 = (v[q * 8 + 4] - mu) * rstd * w1.x + b1.x, y5 = (v[q * 8 + 5] - mu) * rstd * w1.y + b1.y;
    const float y6 = (v[q * 8 + 6] - mu) * rstd * w1.z + b1.z, y7 = (v[q * 8 + 7] - mu) * rstd * w1.w + b1.w;
    const uint32_t a = sA + static_cast<uint32_t>((hf * 2 + (q >> 3)) * TILE + r * 128 + (((q & 7) ^ (r & 7)) << 4));
    sts128(a, pack_bf16x2(y0, y1), pack_bf16x2(y2, y3), pack_bf16x2(y4, y5), pack_bf16x2(y6, y7));
  }
}

// conv1 accumulator quarter -> bf16 plane: this thread's 64 columns of row r, + bias
__device__ __noinline__ void drain_quarter(uint32_t acc, const float* __restrict__ bias_f, uint32_t prow, int hf, int r) {
  using namespace tc;
  uint32_t t[2][32];
  tmem_ld32(acc + hf * 64, t[0]);
  tmem_ld32(acc + hf * 64 + 32, t[1]);
  tmem_wait_ld();
  const float4* bias = reinterpret_cast<const float4*>(bias_f + hf * 64);
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    const float4 b0 = __ldg(bias + 2 * ch), b1 = __ldg(bias + 2 * ch + 1);
    const uint32_t* rr = &t[ch >> 2][(ch & 3) * 8];
    sts128(prow + (((hf * 8 + ch) ^ (r & 7)) << 4), pack_bf16x2(__uint_as_float(rr[0]) + b0.x, __uint_as_float(rr[1]) + b0.y),
           pack_bf16x2(__uint_as_float(rr[2]) + b0.z, __uint_as_float(rr[3]) + b0.w),
           pack_bf16x2(__uint_as_float(rr[4]) + b1.x, __uint_as_float(rr[5]) + b1.y),
           pack_bf16x2(__uint_as_float(rr[6]) + b1.z, __uint_as_float(rr[7]) + b1.w));
  }
}

__global__ void __launch_bounds__(THREADS, 1) pair_block_kernel(const Args args) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sA = sbase + A_OFF, sP = sbase + P_OFF, sR = sbase + R_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [0..4] slots P0..P3, R; [5], [6] accumulators 0 / 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 64);
  float* r_eff = reinterpret_cast<float*>(smem + R_OFF + R_EFF);
  float* r_mean = reinterpret_cast<float*>(smem + R_OFF + R_MEAN);
  float* r_part = reinterpret_cast<float*>(smem + R_OFF + R_PART);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool ctrl = tid == 0;
  const int nb = args.n_blocks;
  const int face0 = blockIdx.x * 2;
  const uint32_t wbar = smem_u32(&bars[0]);
  const uint32_t mbar[2] = {smem_u32(&bars[5]), smem_u32(&bars[6])};

  int n_stamp = 0;
  auto stamp = [&]() {
    if (args.trace != nullptr && blockIdx.x == args.trace_cta && tid == 0) args.trace[n_stamp] = clock64();
    ++n_stamp;
  };
  pdl_trigger();
  stamp();
  if (tid == 0) {
    if ((sbase & 1023u) != 0u) {
      if (atomicCAS(&args.status->error, 0u, 3u) == 0u) args.status->where = 0xB00u;
    }
    for (int i = 0; i < 7; ++i) mbar_init(smem_u32(&bars[i]), 1);
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

  // ---- controller helpers (thread 0 only) ----
  enum { P0 = 0, P1 = 1, P2 = 2, P3 = 3, RS = 4 };
  auto slot_addr = [&](int s) { return s == RS ? sR : sP + static_cast<uint32_t>(s) * SLOT; };
  uint32_t wph = 0;  // per-slot parity bits of the weight barriers (controller)
  // two k-block tiles (k-blocks 2*pair, 2*pair+1) of the 128-row quarter q of W1 / W4
  auto load_pair = [&](int map_idx, int q, int pair, int s) {
    const uint32_t b = wbar + s * 8;
    mbar_expect_tx(b, SLOT);
    tma_load_2d(slot_addr(s), args.maps + map_idx, (pair * 2) * BK, q * 128, b);
    tma_load_2d(slot_addr(s) + TILE, args.maps + map_idx, (pair * 2 + 1) * BK, q * 128, b);
  };
  // one k-block of the 256-row W3 / W5 (N = 256 operand: two stacked 128-row tiles)
  auto load_kb = [&](int map_idx, int kb, int s) {
    const uint32_t b = wbar + s * 8;
    mbar_expect_tx(b, SLOT);
    tma_load_2d(slot_addr(s), args.maps + map_idx, kb * BK, 0, b);
    tma_load_2d(slot_addr(s) + TILE, args.maps + map_idx, kb * BK, 128, b);
  };
  auto wait_slot = [&](int s) {
    fb::mbar_wait_c(wbar + s * 8, (wph >> s) & 1u, args.status, 0xB10u + s);
    wph ^= 1u << s;
  };
  constexpr uint32_t idesc128 = make_idesc(128, 128), idesc256 = make_idesc(128, 256);
  // acc (+)= A[:, k-blocks 2*pair..2*pair+1] * Wq[:, same]^T    (N = 128)
  auto mma_pair = [&](int s, int pair, uint32_t acc_col, bool first) {
    fb::issue_kblock(sA + (pair * 2) * TILE, slot_addr(s), tmem_base + acc_col, first ? 0u : 1u, idesc128);
    fb::issue_kblock(sA + (pair * 2 + 1) * TILE, slot_addr(s) + TILE, tmem_base + acc_col, 1u, idesc128);
  };
  // x += A[:, k-block kb] * W[:, kb]^T    (N = 256, accumulated onto the residual stream)
  auto mma_kb = [&](int s, int kb) { fb::issue_kblock(sA + kb * TILE, slot_addr(s), tmem_base + X_COL, 1u, idesc256); };

  if (ctrl) load_pair(0, 0, 0, P2), load_pair(0, 0, 1, P3);  // conv1 q0 of the first block (constants: before the wait)
  for (int b = 0; b < nb; ++b) {
    const BlockParams bp = args.blocks[b];
    prefetch_l2(reinterpret_cast<const char*>(bp.wsca_t) + tid * 512);
    prefetch_l2(reinterpret_cast<const char*>(bp.wsca_t) + tid * 512 + 128);
    prefetch_l2(reinterpret_cast<const char*>(bp.wsca_t) + tid * 512 + 256);
    prefetch_l2(reinterpret_cast<const char*>(bp.wsca_t) + tid * 512 + 384);
    if (tid < 144) prefetch_l2(bp.dw_w + tid * 32);
  }
  // ---- thread geometry ----
  const int r = (warp & 3) * 32 + lane;               // pixel row = TMEM lane
  const int hf = warp >> 2;                           // column half owned by this thread
  const int fl = r >> 6;                              // local face of this row
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t t_x = tmem_base + lane_addr + X_COL + hf * 128;        // own half of the residual row
  const uint32_t t_acc0 = tmem_base + lane_addr + ACC_COL;
  if (args.warm) {
    // every CTA is the first on its SM (one wave) and finds the instruction cache cold: run the LayerNorm once on
    // whatever tensor memory and the A region hold while the predecessor kernel finishes (face_block.cuh has the numbers)
    residual_ln(t_x, hf, r, args.zero_bias + hf * 128, r_eff + hf * 128, r_eff + C + hf * 128, sA,
                reinterpret_cast<float2*>(smem + R_OFF + R_XCHG));
    block_sync();
  }
  pdl_wait();
  stamp();

  const bool face_ok[2] = {face0 < args.n_faces, face0 + 1 < args.n_faces};
  const float* mrow[2];
#pragma unroll
  for (int f = 0; f < 2; ++f)
    mrow[f] = args.mod_table + static_cast<size_t>(__ldg(args.mod_row_idx + min(face0 + f, args.n_faces - 1))) * args.mod_stride;
  // depthwise geometry: warp -> (64-channel block, face), lane -> channel pair
  const int cb = warp & 3, df = warp >> 2, j = cb * 64 + lane * 2;

  uint32_t mph0 = 0, mph1 = 0;  // parities of the two MMA-done barriers (tracked by every thread)

  // LayerNorm parameters of both faces -> R: eff_w = w (1 + scale), eff_b = b (1 + scale) + shift
  auto make_eff = [&](const float* lw, const float* lb, int shift_off, int scale_off) {
    const float w = __ldg(lw + tid), bb = __ldg(lb + tid);
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const float sc = 1.f + __ldg(mrow[f] + scale_off + tid);
      r_eff[(f * 2 + 0) * C + tid] = w * sc;
      r_eff[(f * 2 + 1) * C + tid] = bb * sc + __ldg(mrow[f] + shift_off + tid);
    }
  };
  auto ln_row = [&](const float* cbias) {
    residual_ln(t_x, hf, r, cbias + hf * 128, r_eff + (fl * 2 + 0) * C + hf * 128, r_eff + (fl * 2 + 1) * C + hf * 128, sA,
                reinterpret_cast<float2*>(smem + R_OFF + R_XCHG));
  };

  // ---------------- prologue: x -> staging (coalesced) -> registers / TMEM; norm1 of the first block ----------------
  {
    const BlockParams bp = args.blocks[0];
    // staging: row rr at rr * 1024 over [A | P0 | P1], 16-byte chunks XOR-swizzled by row
    {
      float4 t[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int rr = warp * 16 + (i >> 1);
        const bool ok = face_ok[rr >> 6];
        const float* src = args.x + (static_cast<size_t>(face0) * FPX + rr) * C + (i & 1) * 128 + lane * 4;
        t[i] = ok ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int rr = warp * 16 + (i >> 1);
        const int q = (i & 1) * 32 + lane;
        sts128(sbase + rr * 1024 + ((q ^ (rr & 7)) << 4), __float_as_uint(t[i].x), __float_as_uint(t[i].y), __float_as_uint(t[i].z),
               __float_as_uint(t[i].w));
      }
    }
    make_eff(bp.ln1_w, bp.ln1_b, bp.mod_off, bp.mod_off + C);
    block_sync();
    {
      const uint32_t srow = sbase + r * 1024;
#pragma unroll 1
      for (int c0 = 0; c0 < 4; ++c0) {
        uint32_t t[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 f = lds128(srow + (((hf * 32 + c0 * 8 + q) ^ (r & 7)) << 4));
          t[4 * q] = __float_as_uint(f.x); t[4 * q + 1] = __float_as_uint(f.y); t[4 * q + 2] = __float_as_uint(f.z); t[4 * q + 3] = __float_as_uint(f.w);
        }
        fb::tmem_st32(t_x + c0 * 32, t);
      }
      fb::tmem_wait_st();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    block_sync();  // every row has left the staging area and sits in tensor memory
    tc_fence_after_sync();
    if (ctrl) load_pair(0, 1, 0, P1);  // conv1 q1, first pair (P1 was staging space)
    ln_row(args.zero_bias);
  }

  for (int b = 0; b < nb; ++b) {
    const BlockParams bp = args.blocks[b];
    const bool last = b + 1 == nb;
    const int m1 = b * 4, m3 = b * 4 + 1, m4 = b * 4 + 2, m5 = b * 4 + 3;
    stamp();  // A ready (norm1)

    // ---------------- conv1: four 128-column quarters, two accumulators, quarter q drained to plane q ----------------
    auto drain = [&](int q, uint32_t acc) {  // this thread's 64 columns of the quarter -> plane q, + bias, bf16
      drain_quarter(acc, bp.b1 + q * 128, sP + static_cast<uint32_t>(q) * SLOT + static_cast<uint32_t>(r) * 256u, hf, r);
    };
    fence_proxy_async_smem();
    tc_fence_before_sync();
    block_sync();                                   // A (norm1) complete; R no longer holds LayerNorm parameters
    if (ctrl) {
      tc_fence_after_sync();
      load_pair(m1, 1, 1, RS);
      wait_slot(P2); wait_slot(P3);
      mma_pair(P2, 0, ACC_COL, true); mma_pair(P3, 1, ACC_COL, false);
      umma_commit(mbar[0]);
      wait_slot(P1); wait_slot(RS);
      mma_pair(P1, 0, ACC_COL + 128, true); mma_pair(RS, 1, ACC_COL + 128, false);
      umma_commit(mbar[1]);
    }
    fb::mbar_wait_c(mbar[0], mph0, args.status, 0xB20u); mph0 ^= 1u;  // q0 done
    tc_fence_after_sync();
    if (ctrl) load_pair(m1, 2, 0, P2), load_pair(m1, 2, 1, P3);
    drain(0, t_acc0);
    tc_fence_before_sync();
    block_sync();                                   // accumulator 0 free
    if (ctrl) {
      tc_fence_after_sync();
      wait_slot(P2); wait_slot(P3);
      mma_pair(P2, 0, ACC_COL, true); mma_pair(P3, 1, ACC_COL, false);
      umma_commit(mbar[0]);
    }
    fb::mbar_wait_c(mbar[1], mph1, args.status, 0xB21u); mph1 ^= 1u;  // q1 done
    tc_fence_after_sync();
    if (ctrl) load_pair(m1, 3, 0, RS);
    drain(1, t_acc0 + 128);
    tc_fence_before_sync();
    block_sync();                                   // accumulator 1 free
    if (ctrl) {
      tc_fence_after_sync();
      wait_slot(RS);
      mma_pair(RS, 0, ACC_COL + 128, true);
    }
    fb::mbar_wait_c(mbar[0], mph0, args.status, 0xB22u); mph0 ^= 1u;  // q2 done: P2, P3 are dead
    tc_fence_after_sync();
    if (ctrl) {
      load_pair(m1, 3, 1, P3);
      wait_slot(P3);
      mma_pair(P3, 1, ACC_COL + 128, false);
      umma_commit(mbar[1]);
    }
    drain(2, t_acc0);
    fb::mbar_wait_c(mbar[1], mph1, args.status, 0xB23u); mph1 ^= 1u;  // q3 done
    tc_fence_after_sync();
    drain(3, t_acc0 + 128);
    tc_fence_before_sync();
    block_sync();                                   // T complete, R free
    stamp();

    // ---------------- depthwise 3x3 + bias + SimpleGate -> A operand; per-face means -> R ----------------
    {
      uint64_t wk1[9], wk2[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        wk1[t] = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_w + t * 512 + j)));
        wk2[t] = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_w + t * 512 + 256 + j)));
      }
      const uint64_t bz1 = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_b + j)));
      const uint64_t bz2 = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_b + 256 + j)));
      const int jj = j & 127;                       // channel inside its plane
      uint32_t lx[8], ax[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        lx[k] = static_cast<uint32_t>((((jj >> 3) ^ k) << 4) + (jj & 7) * 2);
        ax[k] = static_cast<uint32_t>((((lane >> 2) ^ k) << 4) + (lane & 3) * 4);
      }
      const uint32_t plane1 = sP + static_cast<uint32_t>(cb >> 1) * SLOT;  // x1 half lives in planes 0/1, x2 half in 2/3
      constexpr uint32_t X2 = 2 * SLOT;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll 1
      for (int y0 = 0; y0 < SP; y0 += 2) {
        const bool up = y0 > 0, dn = y0 + 2 < SP;
        const int px0 = df * FPX + y0 * SP;         // pixel (y0, 0) of this thread's face
        const uint32_t trow = plane1 + static_cast<uint32_t>(px0) * 256u;
        const uint32_t arow = sA + static_cast<uint32_t>(cb * TILE + px0 * 128);
        uint64_t w1v[3][4], w2v[3][4];
        auto load_col = [&](int x, uint64_t (&c1)[4], uint64_t (&c2)[4]) {
          const uint32_t a = trow + lx[x & 7] + x * 256;
          c1[0] = up ? bf2_to_f2(lds32(a - SP * 256)) : 0ull;
          c2[0] = up ? bf2_to_f2(lds32(a - SP * 256 + X2)) : 0ull;
          c1[1] = bf2_to_f2(lds32(a));
          c2[1] = bf2_to_f2(lds32(a + X2));
          c1[2] = bf2_to_f2(lds32(a + SP * 256));
          c2[2] = bf2_to_f2(lds32(a + SP * 256 + X2));
          c1[3] = dn ? bf2_to_f2(lds32(a + 2 * SP * 256)) : 0ull;
          c2[3] = dn ? bf2_to_f2(lds32(a + 2 * SP * 256 + X2)) : 0ull;
        };
#pragma unroll
        for (int dy = 0; dy < 4; ++dy) w1v[0][dy] = w2v[0][dy] = 0ull;
        load_col(0, w1v[1], w2v[1]);
#pragma unroll
        for (int x = 0; x < SP; ++x) {
          uint64_t (&l1)[4] = w1v[x % 3], (&l2)[4] = w2v[x % 3];
          uint64_t (&c1)[4] = w1v[(x + 1) % 3], (&c2)[4] = w2v[(x + 1) % 3];
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
            a1 = ffma2(wk1[dy * 3 + 1], c1[dy], a1); a2 = ffma2(wk2[dy * 3 + 1], c2[dy], a2);
            b1 = ffma2(wk1[dy * 3 + 1], c1[dy + 1], b1); b2 = ffma2(wk2[dy * 3 + 1], c2[dy + 1], b2);
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
      *reinterpret_cast<float2*>(r_mean + df * C + j) = make_float2(ps0 * (1.f / FPX), ps1 * (1.f / FPX));
    }
    // SCA GEMV: thread -> (output pair n2, n2 + 1; k half); first batch of weights fetched before the barrier
    const int n2 = (tid & 127) * 2, kh = tid >> 7;
    const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(bp.wsca_t + static_cast<size_t>(kh) * 128 * C + n2);
    uint32_t wv[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) wv[i] = __ldg(wsrc + i * (C / 2));
    fence_proxy_async_smem();
    block_sync();                                   // T free; means visible
    stamp();
    if (ctrl) {
      load_kb(m3, 0, P0); load_kb(m3, 1, P1); load_kb(m3, 2, P2); load_kb(m3, 3, P3);
    }
    {
      float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float2 w = unpack_bf16x2(wv[i]);
          const int k = kh * 128 + half * 64 + i;
          const float m0 = r_mean[k], m1v = r_mean[C + k];
          acc[0][0] = fmaf(w.x, m0, acc[0][0]); acc[0][1] = fmaf(w.y, m0, acc[0][1]);
          acc[1][0] = fmaf(w.x, m1v, acc[1][0]); acc[1][1] = fmaf(w.y, m1v, acc[1][1]);
        }
        if (half == 0) {
#pragma unroll
          for (int i = 0; i < 64; ++i) wv[i] = __ldg(wsrc + (64 + i) * (C / 2));
        }
      }
#pragma unroll
      for (int f = 0; f < 2; ++f) *reinterpret_cast<float2*>(r_part + (kh * 2 + f) * C + n2) = make_float2(acc[f][0], acc[f][1]);
    }
    block_sync();
    {  // rescale this thread's own gated values (face df, channels j, j + 1)
      const float2 p0 = *reinterpret_cast<const float2*>(r_part + (0 * 2 + df) * C + j);
      const float2 p1 = *reinterpret_cast<const float2*>(r_part + (1 * 2 + df) * C + j);
      const float2 bs = __ldg(reinterpret_cast<const float2*>(bp.bsca + j));
      const float s0 = p0.x + p1.x + bs.x, s1 = p0.y + p1.y + bs.y;
      const uint32_t abase = sA + static_cast<uint32_t>(cb * TILE + df * FPX * 128) + static_cast<uint32_t>((lane & 3) * 4);
#pragma unroll 8
      for (int i = 0; i < FPX; ++i) {
        const uint32_t a = abase + i * 128 + (((lane >> 2) ^ (i & 7)) << 4);
        const float2 g = unpack_bf16x2(lds32(a));
        sts32(a, pack_bf16x2(g.x * s0, g.y * s1));
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
    }
    block_sync();                                   // A (gated, scaled) complete; R scratch dead
    stamp();

    // ---------------- conv3 (+beta) accumulated onto x; norm2 + modulation -> A ----------------
    if (ctrl) {
      tc_fence_after_sync();
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        wait_slot(kb);
        mma_kb(kb, kb);
      }
      umma_commit(mbar[0]);
    }
    make_eff(bp.ln2_w, bp.ln2_b, bp.mod_off + 2 * C, bp.mod_off + 3 * C);
    block_sync();                                   // LayerNorm parameters visible
    fb::mbar_wait_c(mbar[0], mph0, args.status, 0xB30u); mph0 ^= 1u;
    tc_fence_after_sync();
    if (ctrl) {
      load_pair(m4, 0, 0, P0); load_pair(m4, 0, 1, P1);
      load_pair(m4, 1, 0, P2); load_pair(m4, 1, 1, P3);
    }
    ln_row(bp.cb3);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    block_sync();                                   // A (norm2) complete; R free
    stamp();

    // ---------------- conv4 + SimpleGate: gated values wait in registers until the last quarter's MMAs are done ----------------
    uint32_t hold[4][16];
    auto gate = [&](int q, uint32_t acc, uint32_t (&out)[16]) {  // gated channels q*64 + hf*32 .. +31 of this row
      uint32_t x1[32], x2[32];
      tmem_ld32(acc + hf * 32, x1);
      tmem_ld32(acc + 64 + hf * 32, x2);
      tmem_wait_ld();
      const float4* bias = reinterpret_cast<const float4*>(bp.b4 + q * 128 + hf * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 p = __ldg(bias + i), s = __ldg(bias + 16 + i);
        out[2 * i] = pack_bf16x2((__uint_as_float(x1[4 * i]) + p.x) * (__uint_as_float(x2[4 * i]) + s.x),
                                 (__uint_as_float(x1[4 * i + 1]) + p.y) * (__uint_as_float(x2[4 * i + 1]) + s.y));
        out[2 * i + 1] = pack_bf16x2((__uint_as_float(x1[4 * i + 2]) + p.z) * (__uint_as_float(x2[4 * i + 2]) + s.z),
                                     (__uint_as_float(x1[4 * i + 3]) + p.w) * (__uint_as_float(x2[4 * i + 3]) + s.w));
      }
    };
    if (ctrl) {
      tc_fence_after_sync();
      load_pair(m4, 2, 0, RS);
      wait_slot(P0); wait_slot(P1);
      mma_pair(P0, 0, ACC_COL, true); mma_pair(P1, 1, ACC_COL, false);
      umma_commit(mbar[0]);
      wait_slot(P2); wait_slot(P3);
      mma_pair(P2, 0, ACC_COL + 128, true); mma_pair(P3, 1, ACC_COL + 128, false);
      umma_commit(mbar[1]);
    }
    fb::mbar_wait_c(mbar[0], mph0, args.status, 0xB40u); mph0 ^= 1u;  // q0 done: P0, P1 dead
    tc_fence_after_sync();
    if (ctrl) {
      load_pair(m4, 2, 1, P0);
      load_kb(m5, 0, P1);
    }
    gate(0, t_acc0, hold[0]);
    tc_fence_before_sync();
    block_sync();                                   // accumulator 0 free
    if (ctrl) {
      tc_fence_after_sync();
      wait_slot(RS); wait_slot(P0);
      mma_pair(RS, 0, ACC_COL, true); mma_pair(P0, 1, ACC_COL, false);
      umma_commit(mbar[0]);
    }
    fb::mbar_wait_c(mbar[1], mph1, args.status, 0xB41u); mph1 ^= 1u;  // q1 done: P2, P3 dead
    tc_fence_after_sync();
    if (ctrl) load_pair(m4, 3, 0, P2), load_pair(m4, 3, 1, P3);
    gate(1, t_acc0 + 128, hold[1]);
    tc_fence_before_sync();
    block_sync();                                   // accumulator 1 free
    if (ctrl) {
      tc_fence_after_sync();
      wait_slot(P2); wait_slot(P3);
      mma_pair(P2, 0, ACC_COL + 128, true); mma_pair(P3, 1, ACC_COL + 128, false);
      umma_commit(mbar[1]);
    }
    fb::mbar_wait_c(mbar[0], mph0, args.status, 0xB42u); mph0 ^= 1u;  // q2 done: R, P0 dead
    tc_fence_after_sync();
    if (ctrl) load_kb(m5, 1, P0);
    gate(2, t_acc0, hold[2]);
    fb::mbar_wait_c(mbar[1], mph1, args.status, 0xB43u); mph1 ^= 1u;  // q3 done: every conv4 MMA has read A
    tc_fence_after_sync();
    if (ctrl) load_kb(m5, 2, P2), load_kb(m5, 3, P3);
    gate(3, t_acc0 + 128, hold[3]);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        sts128(sA + static_cast<uint32_t>(q * TILE + r * 128 + (((hf * 4 + ch) ^ (r & 7)) << 4)), hold[q][4 * ch], hold[q][4 * ch + 1],
               hold[q][4 * ch + 2], hold[q][4 * ch + 3]);
    if (!last) {
      const BlockParams nx = args.blocks[b + 1];
      make_eff(nx.ln1_w, nx.ln1_b, nx.mod_off, nx.mod_off + C);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    block_sync();                                   // A (gated) complete; LayerNorm parameters visible
    stamp();

    // ---------------- conv5 (+gamma) accumulated onto x; next block's norm1 or the final store ----------------
    if (ctrl) {
      tc_fence_after_sync();
      wait_slot(P1); mma_kb(P1, 0);
      wait_slot(P0); mma_kb(P0, 1);
      wait_slot(P2); mma_kb(P2, 2);
      wait_slot(P3); mma_kb(P3, 3);
      umma_commit(mbar[0]);
    }
    fb::mbar_wait_c(mbar[0], mph0, args.status, 0xB50u); mph0 ^= 1u;
    tc_fence_after_sync();
    if (ctrl && !last) {
      load_pair(m1 + 4, 0, 0, P2); load_pair(m1 + 4, 0, 1, P3);
      load_pair(m1 + 4, 1, 0, P1);
    }
    if (!last) ln_row(bp.cb5);
  }

  // ---------------- final store: rows -> swizzled staging over [A | P0 | P1] -> coalesced global rows ----------------
  {
    const float* cb = args.blocks[nb - 1].cb5;
    const uint32_t srow = sbase + r * 1024;
#pragma unroll 1
    for (int c0 = 0; c0 < 4; ++c0) {
      uint32_t t[32];
      tmem_ld32(t_x + c0 * 32, t);
      tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(cb + hf * 128 + c0 * 32 + q * 4));
        sts128(srow + (((hf * 32 + c0 * 8 + q) ^ (r & 7)) << 4), __float_as_uint(__uint_as_float(t[4 * q]) + bb.x),
               __float_as_uint(__uint_as_float(t[4 * q + 1]) + bb.y), __float_as_uint(__uint_as_float(t[4 * q + 2]) + bb.z),
               __float_as_uint(__uint_as_float(t[4 * q + 3]) + bb.w));
      }
    }
  }
  block_sync();
#pragma unroll 8
  for (int i = 0; i < 32; ++i) {
    const int rr = warp * 16 + (i >> 1);
    const int q = (i & 1) * 32 + lane;
    if (face_ok[rr >> 6])
      *reinterpret_cast<float4*>(args.x + (static_cast<size_t>(face0) * FPX + rr) * C + (i & 1) * 128 + lane * 4) =
          lds128(sbase + rr * 1024 + ((q ^ (rr & 7)) << 4));
  }
  tc_fence_before_sync();
  block_sync();
  stamp();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pb
}  // namespace hd
