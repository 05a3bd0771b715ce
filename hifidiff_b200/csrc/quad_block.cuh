// Fused ConditionalNAFBlock kernel for the 4x4 level (c = 512): a CLUSTER of four CTAs runs eight faces (8 x 16
// pixels = one 128-row MMA tile) through a run of consecutive blocks.  The block's weights (3.5 MB bf16) are far
// too large for one SM to stream per m-tile, so the channel dimension is split over the cluster: CTA `rank` owns
// channels [128 rank, 128 rank + 128) of the residual stream (in its tensor memory), of the gated tensors and of
// every GEMM's output, and streams only the matching quarter of each weight matrix.  What a GEMM needs from the
// other three CTAs - the full-width bf16 A operand (LayerNorm output / gated tensor), LayerNorm partial statistics
// and the per-face SCA means - is exchanged through small L2-resident buffers between hardware cluster barriers
// (barrier.cluster release / acquire); no grid-wide synchronisation and no kernel boundary inside the run.
//
//   smem   A operand 128 KB (8 k-block tiles, all 512 channels) | P0, P1: conv1 output planes (x1 / x2 slice,
//          [128 px][128 ch] bf16) | R 32 KB | barriers.  P0, P1 and R double as the 3-slot weight ring while the
//          planes are dead and as scratch (LayerNorm parameters, SCA means / partial sums) between GEMMs.
//   TMEM   x slice (128 fp32 columns, bias-free, see pair_block.cuh) | accumulator 0 | accumulator 1
//   every GEMM of the block is the same shape for one CTA: [128 rows x 512] x [128 x 512]^T, streamed as 4 pairs of
//   64-wide k-blocks through the ring by thread 0 (TMA producer and MMA issuer).
//
// Reference arithmetic: models/denoiser/conditional_naf.py:108-136, utils.py:16-24,57-60.
#pragma once

#include "common.cuh"
#include "face_block.cuh"
#include "gemm_tc.cuh"

namespace hd {
namespace qb {

using fb::bf2_to_f2;
using fb::block_sync;
using fb::ffma2;
using fb::lds128;
using fb::lds32;
using fb::pack_f2;
using fb::sts128;
using fb::sts32;
using fb::unpack_f2;

constexpr int C = 512;
constexpr int SP = 4;
constexpr int FPX = SP * SP;              // 16 pixels per face
constexpr int FACES = 8;                  // faces per cluster (one 128-row tile)
constexpr int CL = 4;                     // CTAs per cluster
constexpr int CS = C / CL;                // channels owned by one CTA
constexpr int THREADS = 256;
constexpr int TILE = 16384;
constexpr int SLOT = 2 * TILE;
constexpr int A_OFF = 0;                  // 8 k-block tiles
constexpr int P_OFF = 8 * TILE;           // planes P0, P1
constexpr int R_OFF = P_OFF + 2 * SLOT;
constexpr int BAR_OFF = R_OFF + SLOT;
constexpr int SMEM_BYTES = BAR_OFF + 128;
constexpr int MAX_BLOCKS = 4;
constexpr uint32_t X_COL = 0, ACC_COL = 128;
// scratch (bytes) inside R / the planes while they hold neither weights nor the conv1 tile
constexpr int R_EFF = 0;                  // R: [face 8][w|b][128] fp32 = 8 KB
constexpr int P_MEAN = 0;                 // P0: [face 8][512] fp32 = 16 KB
constexpr int P_PART = SLOT;              // P1: [k quarter 4][face 8][128] fp32 = 16 KB

struct BlockParams {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const float *b1, *dw_w, *dw_b;   // conv1 bias [1024]; depthwise taps [9][1024] and bias [1024]
  const bf16* wsca_t;              // SCA weight transposed [k][n] bf16
  const float *bsca, *b4;          // SCA bias; gate-packed conv4 bias [1024]
  const float *cb3, *cb5;          // cumulative residual bias after this block's conv3 / conv5
  int mod_off, pad;
};

struct Args {
  const CUtensorMap* maps;         // [n_blocks][4]: w1 [1024,512], w3 [512,512], w4 (gate-packed) [1024,512], w5 [512,512]
  const BlockParams* blocks;
  int n_blocks, n_faces, n_mtiles;
  float* x;                        // residual stream [faces * 16, 512] fp32, updated in place
  bf16* xa;                        // exchange: [parity 2][m-tile][128 rows][512] bf16, the full-width A operand.  Two
                                   // buffers alternate: a CTA may publish exchange n+1 while a peer still gathers n
  float2* stats;                   // exchange: [m-tile][rank 4][half 2][128 rows] (mean, M2) of 64 channels
  float* means;                    // exchange: [m-tile][face 8][512] SCA means
  const float* zero_bias;          // 512 zeros
  const float* mod_table;
  const int* mod_row_idx;
  int mod_stride;
  DeviceStatus* status;
  long long* trace;                // optional: clock64 stamps of one CTA's phase boundaries (diagnostics)
  int trace_cta;
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t v;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(v));
  return v;
}
__device__ __forceinline__ uint4 ldcg128(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

// LayerNorm2d + AdaLN modulation of this CTA's 128-channel slice of the residual rows.  Each thread holds 64 channels
// of row r; the (mean, M2) of all eight 64-channel groups of the row meet in an L2 buffer across a cluster barrier
// and are combined exactly.  The bf16 result goes to k-block 2*rank + hf of the local A operand and to the exchange
// buffer for the other three CTAs.  Must be called by every thread of every CTA of the cluster.
__device__ __noinline__ void residual_ln(uint32_t t_own, int hf, int r, int rank, const float* __restrict__ cbias_own,
                                         const float* eff_w, const float* eff_b, uint32_t sA, float2* stats_tile,
                                         bf16* xa_row) {
  using namespace tc;
  float v[64];
  {
    uint32_t t[2][32];
    tmem_ld32(t_own, t[0]);
    tmem_ld32(t_own + 32, t[1]);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(cbias_own + c * 32 + i));
        v[c * 32 + i] = __uint_as_float(t[c][i]) + bb.x; v[c * 32 + i + 1] = __uint_as_float(t[c][i + 1]) + bb.y;
        v[c * 32 + i + 2] = __uint_as_float(t[c][i + 2]) + bb.z; v[c * 32 + i + 3] = __uint_as_float(t[c][i + 3]) + bb.w;
      }
  }
  float sa[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sa[i] = v[i];
#pragma unroll
  for (int i = 8; i < 64; ++i) sa[i & 7] += v[i];
  const float mean_g = (((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]))) * (1.f / 64);
#pragma unroll
  for (int i = 0; i < 8; ++i) sa[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    const float d = v[i] - mean_g;
    sa[i & 7] = fmaf(d, d, sa[i & 7]);
  }
  const float m2_g = ((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]));
  stats_tile[(rank * 2 + hf) * 128 + r] = make_float2(mean_g, m2_g);
  cluster_sync_all();
  float2 g[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] = __ldcg(stats_tile + i * 128 + r);
  float mu = 0.f, m2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { mu += g[i].x; m2 += g[i].y; }
  mu *= 0.125f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = g[i].x - mu; m2 = fmaf(64.f * d, d, m2); }
  const float rstd = 1.f / sqrtf(m2 * (1.f / C) + 1e-6f);
  const uint32_t arow = sA + static_cast<uint32_t>((rank * 2 + hf) * TILE + r * 128);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 w0 = *reinterpret_cast<const float4*>(eff_w + q * 8), w1 = *reinterpret_cast<const float4*>(eff_w + q * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(eff_b + q * 8), b1 = *reinterpret_cast<const float4*>(eff_b + q * 8 + 4);
    const uint32_t p0 = pack_bf16x2((v[q * 8 + 0] - mu) * rstd * w0.x + b0.x, (v[q * 8 + 1] - mu) * rstd * w0.y + b0.y);
    const uint32_t p1 = pack_bf16x2((v[q * 8 + 2] - mu) * rstd * w0.z + b0.z, (v[q * 8 + 3] - mu) * rstd * w0.w + b0.w);
    const uint32_t p2 = pack_bf16x2((v[q * 8 + 4] - mu) * rstd * w1.x + b1.x, (v[q * 8 + 5] - mu) * rstd * w1.y + b1.y);
    const uint32_t p3 = pack_bf16x2((v[q * 8 + 6] - mu) * rstd * w1.z + b1.z, (v[q * 8 + 7] - mu) * rstd * w1.w + b1.w);
    sts128(arow + ((q ^ (r & 7)) << 4), p0, p1, p2, p3);
    *reinterpret_cast<uint4*>(xa_row + q * 8) = make_uint4(p0, p1, p2, p3);
  }
}

// the other three CTAs' 128-channel slices of the A operand: exchange buffer (L2) -> local k-block tiles
__device__ __noinline__ void gather_a(const bf16* xa_tile, int rank, uint32_t sA, int tid) {
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    uint4 u[12];  // twelve 16-byte loads in flight per thread: the L2 round trip is paid twice, not 24 times
#pragma unroll
    for (int it = 0; it < 12; ++it) {
      const int idx = (half * 12 + it) * THREADS + tid;
      const int cc = idx & 7, row = (idx >> 3) & 127, sel = idx >> 10;  // sel 0..5: the six foreign k-blocks
      const int kb = sel + (sel >= rank * 2 ? 2 : 0);
      u[it] = ldcg128(xa_tile + static_cast<size_t>(row) * C + kb * 64 + cc * 8);
    }
#pragma unroll
    for (int it = 0; it < 12; ++it) {
      const int idx = (half * 12 + it) * THREADS + tid;
      const int cc = idx & 7, row = (idx >> 3) & 127, sel = idx >> 10;
      const int kb = sel + (sel >= rank * 2 ? 2 : 0);
      sts128(sA + static_cast<uint32_t>(kb * TILE + row * 128 + ((cc ^ (row & 7)) << 4)), u[it].x, u[it].y, u[it].z, u[it].w);
    }
  }
}

// own slice of the A operand (k-blocks 2 rank, 2 rank + 1) -> exchange buffer
__device__ __noinline__ void publish_a(bf16* xa_tile, int rank, uint32_t sA, int tid) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * THREADS + tid;
    const int cc = idx & 7, row = (idx >> 3) & 127, kb = rank * 2 + (idx >> 10);
    const float4 f = lds128(sA + static_cast<uint32_t>(kb * TILE + row * 128 + ((cc ^ (row & 7)) << 4)));
    *reinterpret_cast<float4*>(xa_tile + static_cast<size_t>(row) * C + kb * 64 + cc * 8) = f;
  }
}

__global__ void __launch_bounds__(THREADS, 1) quad_block_kernel(const Args args) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sA = sbase + A_OFF, sP = sbase + P_OFF, sR = sbase + R_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [0..2] ring full, [3..5] ring empty, [6], [7] GEMM done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 64);
  float* r_eff = reinterpret_cast<float*>(smem + R_OFF + R_EFF);
  float* p_mean = reinterpret_cast<float*>(smem + P_OFF + P_MEAN);
  float* p_part = reinterpret_cast<float*>(smem + P_OFF + P_PART);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool ctrl = tid == 0;
  const int nb = args.n_blocks;
  const int rank = static_cast<int>(cluster_rank());
  const int mtile = blockIdx.x / CL;
  const int face0 = mtile * FACES;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[3]);
  const uint32_t dbar[2] = {smem_u32(&bars[6]), smem_u32(&bars[7])};

  int n_stamp = 0;
  auto stamp = [&]() {
    if (args.trace != nullptr && blockIdx.x == args.trace_cta && tid == 0) args.trace[n_stamp] = clock64();
    ++n_stamp;
  };
  pdl_trigger();
  stamp();
  if (tid == 0) {
    if ((sbase & 1023u) != 0u) {
      if (atomicCAS(&args.status->error, 0u, 3u) == 0u) args.status->where = 0xC00u;
    }
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
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

  // ---- controller: one [128 x 512] x [128 x 512]^T GEMM streamed through the 3-slot ring (R, P0, P1) ----
  uint32_t n_pairs = 0;  // k-block pairs streamed so far (ring position, carried across GEMMs)
  auto slot_addr = [&](uint32_t n) { const uint32_t s = n % 3u; return s == 0 ? sR : sP + (s - 1) * SLOT; };
  constexpr uint32_t idesc = make_idesc(128, 128);
  auto gemm = [&](int map_idx, int row0, uint32_t tmem_d, uint32_t accumulate_first, uint32_t done_bar) {
    const CUtensorMap* m = args.maps + map_idx;
    uint32_t loaded = n_pairs;
#pragma unroll 1
    for (uint32_t p = 0; p < 4; ++p) {
      const uint32_t n = n_pairs + p;
#pragma unroll 1
      while (loaded <= n + 2 && loaded < n_pairs + 4) {  // keep up to three pairs in flight
        const uint32_t s = loaded % 3u;
        if (loaded >= 3) fb::mbar_wait_c(empty0 + s * 8, ((loaded / 3u) - 1u) & 1u, args.status, 0xC10u);
        const uint32_t fb_ = full0 + s * 8, dst = slot_addr(loaded);
        const int kb = static_cast<int>(loaded - n_pairs) * 2;
        mbar_expect_tx(fb_, SLOT);
        tma_load_2d(dst, m, kb * BK, row0, fb_);
        tma_load_2d(dst + TILE, m, (kb + 1) * BK, row0, fb_);
        ++loaded;
      }
      const uint32_t s = n % 3u;
      fb::mbar_wait_c(full0 + s * 8, (n / 3u) & 1u, args.status, 0xC11u);
      tc_fence_after_sync();
      fb::issue_kblock(sA + (2 * p) * TILE, slot_addr(n), tmem_d, p == 0 ? accumulate_first : 1u, idesc);
      fb::issue_kblock(sA + (2 * p + 1) * TILE, slot_addr(n) + TILE, tmem_d, 1u, idesc);
      umma_commit(empty0 + s * 8);
    }
    umma_commit(done_bar);
    n_pairs += 4;
  };

  pdl_wait();
  stamp();

  // ---- thread geometry ----
  const int r = (warp & 3) * 32 + lane;               // pixel row = TMEM lane
  const int hf = warp >> 2;                           // 64-column half of every 128-column entity
  const int fl = r >> 4;                              // local face of this row
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t t_x = tmem_base + lane_addr + X_COL + hf * 64;
  const uint32_t t_acc0 = tmem_base + lane_addr + ACC_COL;
  const bool row_ok = face0 + fl < args.n_faces;
  float* x_row = args.x + (static_cast<size_t>(face0) * FPX + r) * C + rank * CS + hf * 64;
  uint32_t xph = 0;  // exchange-buffer parity, flipped after every gather (identically in every thread of the cluster)
  auto xa_tile = [&]() { return args.xa + (static_cast<size_t>(xph) * args.n_mtiles + mtile) * 128 * C; };
  float2* stats_tile = args.stats + static_cast<size_t>(mtile) * (CL * 2 * 128);
  float* means_tile = args.means + static_cast<size_t>(mtile) * FACES * C;
  // depthwise geometry: warp -> (64-channel block of the slice, face pair), lane -> channel pair
  const int cb = warp & 1, fq = warp >> 1, jj = cb * 64 + lane * 2;

  uint32_t dph0 = 0, dph1 = 0;  // parities of the two GEMM-done barriers

  // LayerNorm parameters of the own slice for all eight faces: eff_w = w (1 + scale), eff_b = b (1 + scale) + shift.
  // Loaded into registers while a GEMM still streams through R, stored to R once the ring is idle.
  float effv[4][2];
  auto eff_load = [&](const float* lw, const float* lb, int shift_off, int scale_off) {
    const int c = tid & 127, part = tid >> 7;
    const float w = __ldg(lw + rank * CS + c), bb = __ldg(lb + rank * CS + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = part * 4 + i;
      const float* mrow =
          args.mod_table + static_cast<size_t>(__ldg(args.mod_row_idx + min(face0 + f, args.n_faces - 1))) * args.mod_stride;
      const float sc = 1.f + __ldg(mrow + scale_off + rank * CS + c);
      effv[i][0] = w * sc;
      effv[i][1] = bb * sc + __ldg(mrow + shift_off + rank * CS + c);
    }
  };
  auto eff_store = [&]() {
    const int c = tid & 127, part = tid >> 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r_eff[((part * 4 + i) * 2 + 0) * CS + c] = effv[i][0];
      r_eff[((part * 4 + i) * 2 + 1) * CS + c] = effv[i][1];
    }
  };
  auto exchange = [&]() {  // every CTA has published its slice of the A operand: pull in the other three
    cluster_sync_all();
    gather_a(xa_tile(), rank, sA, tid);
    xph ^= 1u;
  };
  auto ln_and_gather = [&](const float* cbias) {
    residual_ln(t_x, hf, r, rank, cbias + rank * CS + hf * 64, r_eff + (fl * 2 + 0) * CS + hf * 64,
                r_eff + (fl * 2 + 1) * CS + hf * 64, sA, stats_tile, xa_tile() + static_cast<size_t>(r) * C + rank * CS + hf * 64);
    exchange();
  };

  // ---------------- prologue: own slice of x -> tensor memory; norm1 of the first block ----------------
  {
    const BlockParams bp = args.blocks[0];
    eff_load(bp.ln1_w, bp.ln1_b, bp.mod_off, bp.mod_off + C);
    eff_store();
#pragma unroll 1
    for (int c0 = 0; c0 < 2; ++c0) {
      uint32_t t[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 f = row_ok ? *reinterpret_cast<const float4*>(x_row + c0 * 32 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        t[4 * q] = __float_as_uint(f.x); t[4 * q + 1] = __float_as_uint(f.y); t[4 * q + 2] = __float_as_uint(f.z); t[4 * q + 3] = __float_as_uint(f.w);
      }
      fb::tmem_st32(t_x + c0 * 32, t);
    }
    fb::tmem_wait_st();
    block_sync();  // LayerNorm parameters visible
    ln_and_gather(args.zero_bias);
  }

  for (int b = 0; b < nb; ++b) {
    const BlockParams bp = args.blocks[b];
    const bool last = b + 1 == nb;
    const int m1 = b * 4, m3 = b * 4 + 1, m4 = b * 4 + 2, m5 = b * 4 + 3;
    stamp();  // 0: A ready (norm1 gathered)

    // ---------------- conv1: x1 quarter -> accumulator 0, x2 quarter -> accumulator 1 ----------------
    fence_proxy_async_smem();
    tc_fence_before_sync();
    block_sync();                                   // A (norm1, all 512 channels) complete; R / planes free
    if (ctrl) {
      tc_fence_after_sync();
      gemm(m1, rank * CS, tmem_base + ACC_COL, 0u, dbar[0]);
      gemm(m1, C + rank * CS, tmem_base + ACC_COL + 128, 0u, dbar[1]);
    }
    uint32_t hold[32];
    // accumulator quarter (+ bias) -> 64 bf16 of this row
    auto take = [&](uint32_t acc, const float* bias_f, uint32_t (&out)[32]) {
      uint32_t t[2][32];
      tmem_ld32(acc + hf * 64, t[0]);
      tmem_ld32(acc + hf * 64 + 32, t[1]);
      tmem_wait_ld();
      const float4* bias = reinterpret_cast<const float4*>(bias_f + hf * 64);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 b4 = __ldg(bias + i);
        const uint32_t* rr = &t[i >> 3][(i & 7) * 4];
        out[2 * i] = pack_bf16x2(__uint_as_float(rr[0]) + b4.x, __uint_as_float(rr[1]) + b4.y);
        out[2 * i + 1] = pack_bf16x2(__uint_as_float(rr[2]) + b4.z, __uint_as_float(rr[3]) + b4.w);
      }
    };
    auto put_plane = [&](int plane, const uint32_t (&in)[32]) {
      const uint32_t prow = sP + static_cast<uint32_t>(plane) * SLOT + static_cast<uint32_t>(r) * 256u;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        sts128(prow + (((hf * 8 + ch) ^ (r & 7)) << 4), in[4 * ch], in[4 * ch + 1], in[4 * ch + 2], in[4 * ch + 3]);
    };
    fb::mbar_wait_c(dbar[0], dph0, args.status, 0xC20u); dph0 ^= 1u;
    tc_fence_after_sync();
    take(t_acc0, bp.b1 + rank * CS, hold);          // the planes are still ring slots of the second quarter
    fb::mbar_wait_c(dbar[1], dph1, args.status, 0xC21u); dph1 ^= 1u;
    tc_fence_after_sync();
    put_plane(0, hold);
    take(t_acc0 + 128, bp.b1 + C + rank * CS, hold);
    put_plane(1, hold);
    tc_fence_before_sync();
    block_sync();                                   // conv1 tile complete
    stamp();  // 1: conv1 done

    // ---------------- depthwise 3x3 + bias + SimpleGate -> own A k-blocks; per-face means -> exchange ----------------
    {
      uint64_t wk1[9], wk2[9];
      const int gch = rank * CS + jj;               // global gated channel of this lane's pair
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        wk1[t] = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_w + t * 2 * C + gch)));
        wk2[t] = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_w + t * 2 * C + C + gch)));
      }
      const uint64_t bz1 = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_b + gch)));
      const uint64_t bz2 = pack_f2(__ldg(reinterpret_cast<const float2*>(bp.dw_b + C + gch)));
      uint32_t lx[8], ax[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        lx[k] = static_cast<uint32_t>((((jj >> 3) ^ k) << 4) + (jj & 7) * 2);
        ax[k] = static_cast<uint32_t>((((lane >> 2) ^ k) << 4) + (lane & 3) * 4);
      }
#pragma unroll 1
      for (int fi = 0; fi < 2; ++fi) {
        const int f = fq * 2 + fi;
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll 1
        for (int y0 = 0; y0 < SP; y0 += 2) {
          const bool up = y0 > 0, dn = y0 + 2 < SP;
          const int px0 = f * FPX + y0 * SP;          // pixel (y0, 0); px0 & 7 == 0
          const uint32_t trow = sP + static_cast<uint32_t>(px0) * 256u;
          const uint32_t arow = sA + static_cast<uint32_t>((rank * 2 + cb) * TILE + px0 * 128);
          uint64_t w1v[3][4], w2v[3][4];
          // input rows y0-1, y0, y0+1, y0+2 of column x: their (pixel & 7) is x+4, x, x+4, x
          auto load_col = [&](int x, uint64_t (&c1)[4], uint64_t (&c2)[4]) {
            const uint32_t ae = trow + lx[x] + x * 256, ao = trow + lx[x + 4] + x * 256;
            c1[0] = up ? bf2_to_f2(lds32(ao - SP * 256)) : 0ull;
            c2[0] = up ? bf2_to_f2(lds32(ao - SP * 256 + SLOT)) : 0ull;
            c1[1] = bf2_to_f2(lds32(ae));
            c2[1] = bf2_to_f2(lds32(ae + SLOT));
            c1[2] = bf2_to_f2(lds32(ao + SP * 256));
            c2[2] = bf2_to_f2(lds32(ao + SP * 256 + SLOT));
            c1[3] = dn ? bf2_to_f2(lds32(ae + 2 * SP * 256)) : 0ull;
            c2[3] = dn ? bf2_to_f2(lds32(ae + 2 * SP * 256 + SLOT)) : 0ull;
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
            sts32(arow + ax[x] + x * 128, pack_bf16x2(ga0, ga1));               // row px0 + x:      (px & 7) = x
            sts32(arow + ax[x + 4] + (SP + x) * 128, pack_bf16x2(gb0, gb1));    // row px0 + 4 + x:  (px & 7) = x + 4
          }
        }
        *reinterpret_cast<float2*>(means_tile + f * C + rank * CS + jj) = make_float2(ps0 * (1.f / FPX), ps1 * (1.f / FPX));
      }
    }
    // SCA weights of this thread (output pair, k quarter): first half fetched before the barrier
    const int np = (tid & 63) * 2, kq = tid >> 6;
    const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(bp.wsca_t + static_cast<size_t>(kq) * 128 * C + rank * CS + np);
    uint32_t wv[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) wv[i] = __ldg(wsrc + i * (C / 2));
    stamp();  // 2: depthwise done
    cluster_sync_all();                             // every CTA's means are in the exchange buffer; conv1 tile dead
    stamp();  // 3: means barrier
    // ---------------- SCA: s = Wsca mean + b for the own 128 outputs, all 8 faces ----------------
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * THREADS + tid;           // 1024 float4 = [8 faces][512] -> staged as [k][8 faces]
      const float4 m = __ldcg(reinterpret_cast<const float4*>(means_tile) + idx);
      const int f = idx >> 7, k = (idx & 127) * 4;
      p_mean[(k + 0) * FACES + f] = m.x; p_mean[(k + 1) * FACES + f] = m.y;
      p_mean[(k + 2) * FACES + f] = m.z; p_mean[(k + 3) * FACES + f] = m.w;
    }
    block_sync();
    {
      float acc[FACES][2];
#pragma unroll
      for (int f = 0; f < FACES; ++f) acc[f][0] = acc[f][1] = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float2 w = unpack_bf16x2(wv[i]);
          const int k = kq * 128 + half * 64 + i;
          const float4 ma = *reinterpret_cast<const float4*>(p_mean + k * FACES), mb = *reinterpret_cast<const float4*>(p_mean + k * FACES + 4);
          const float mm[FACES] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
#pragma unroll
          for (int f = 0; f < FACES; ++f) {
            acc[f][0] = fmaf(w.x, mm[f], acc[f][0]);
            acc[f][1] = fmaf(w.y, mm[f], acc[f][1]);
          }
        }
        if (half == 0) {
#pragma unroll
          for (int i = 0; i < 64; ++i) wv[i] = __ldg(wsrc + (64 + i) * (C / 2));
        }
      }
#pragma unroll
      for (int f = 0; f < FACES; ++f) *reinterpret_cast<float2*>(p_part + (kq * FACES + f) * CS + np) = make_float2(acc[f][0], acc[f][1]);
    }
    block_sync();
    {  // rescale this thread's own gated values (faces 2fq, 2fq+1; channels jj, jj+1 of the slice)
      const float2 bs = __ldg(reinterpret_cast<const float2*>(bp.bsca + rank * CS + jj));
#pragma unroll
      for (int fi = 0; fi < 2; ++fi) {
        const int f = fq * 2 + fi;
        float s0 = bs.x, s1 = bs.y;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float2 p = *reinterpret_cast<const float2*>(p_part + (k4 * FACES + f) * CS + jj);
          s0 += p.x; s1 += p.y;
        }
        const uint32_t abase = sA + static_cast<uint32_t>((rank * 2 + cb) * TILE + f * FPX * 128) + static_cast<uint32_t>((lane & 3) * 4);
#pragma unroll
        for (int i = 0; i < FPX; ++i) {
          const uint32_t a = abase + i * 128 + (((lane >> 2) ^ (i & 7)) << 4);
          const float2 g = unpack_bf16x2(lds32(a));
          sts32(a, pack_bf16x2(g.x * s0, g.y * s1));
        }
      }
    }
    block_sync();
    stamp();  // 4: SCA + rescale done
    publish_a(xa_tile(), rank, sA, tid);
    exchange();
    stamp();  // 5: gated operand gathered
    fence_proxy_async_smem();
    tc_fence_before_sync();
    block_sync();                                   // A (gated, scaled, all 512 channels) complete

    // ---------------- conv3 (+beta) accumulated onto the x slice; norm2 + modulation -> A ----------------
    if (ctrl) {
      tc_fence_after_sync();
      gemm(m3, rank * CS, tmem_base + X_COL, 1u, dbar[0]);
    }
    eff_load(bp.ln2_w, bp.ln2_b, bp.mod_off + 2 * C, bp.mod_off + 3 * C);
    fb::mbar_wait_c(dbar[0], dph0, args.status, 0xC30u); dph0 ^= 1u;
    tc_fence_after_sync();
    block_sync();                                   // ring idle in every thread's view: R is scratch again
    stamp();  // 6: conv3 done
    eff_store();
    block_sync();
    ln_and_gather(bp.cb3);
    stamp();  // 7: norm2 gathered

    // ---------------- conv4 + SimpleGate ----------------
    fence_proxy_async_smem();
    tc_fence_before_sync();
    block_sync();                                   // A (norm2) complete
    if (ctrl) {
      tc_fence_after_sync();
      gemm(m4, rank * 256, tmem_base + ACC_COL, 0u, dbar[0]);
      gemm(m4, rank * 256 + 128, tmem_base + ACC_COL + 128, 0u, dbar[1]);
    }
    auto gate = [&](int q, uint32_t acc, uint32_t (&out)[16]) {  // gated channels q*64 + hf*32 .. +31 of the slice
      uint32_t x1[32], x2[32];
      tmem_ld32(acc + hf * 32, x1);
      tmem_ld32(acc + 64 + hf * 32, x2);
      tmem_wait_ld();
      const float4* bias = reinterpret_cast<const float4*>(bp.b4 + rank * 256 + q * 128 + hf * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 p = __ldg(bias + i), s = __ldg(bias + 16 + i);
        out[2 * i] = pack_bf16x2((__uint_as_float(x1[4 * i]) + p.x) * (__uint_as_float(x2[4 * i]) + s.x),
                                 (__uint_as_float(x1[4 * i + 1]) + p.y) * (__uint_as_float(x2[4 * i + 1]) + s.y));
        out[2 * i + 1] = pack_bf16x2((__uint_as_float(x1[4 * i + 2]) + p.z) * (__uint_as_float(x2[4 * i + 2]) + s.z),
                                     (__uint_as_float(x1[4 * i + 3]) + p.w) * (__uint_as_float(x2[4 * i + 3]) + s.w));
      }
    };
    uint32_t g0[16], g1[16];
    fb::mbar_wait_c(dbar[0], dph0, args.status, 0xC40u); dph0 ^= 1u;
    tc_fence_after_sync();
    gate(0, t_acc0, g0);
    fb::mbar_wait_c(dbar[1], dph1, args.status, 0xC41u); dph1 ^= 1u;  // every conv4 MMA has read A
    tc_fence_after_sync();
    gate(1, t_acc0 + 128, g1);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      sts128(sA + static_cast<uint32_t>((rank * 2 + 0) * TILE + r * 128 + (((hf * 4 + ch) ^ (r & 7)) << 4)), g0[4 * ch], g0[4 * ch + 1],
             g0[4 * ch + 2], g0[4 * ch + 3]);
      sts128(sA + static_cast<uint32_t>((rank * 2 + 1) * TILE + r * 128 + (((hf * 4 + ch) ^ (r & 7)) << 4)), g1[4 * ch], g1[4 * ch + 1],
             g1[4 * ch + 2], g1[4 * ch + 3]);
      bf16* xr = xa_tile() + static_cast<size_t>(r) * C + rank * CS + hf * 32 + ch * 8;
      *reinterpret_cast<uint4*>(xr) = make_uint4(g0[4 * ch], g0[4 * ch + 1], g0[4 * ch + 2], g0[4 * ch + 3]);
      *reinterpret_cast<uint4*>(xr + 64) = make_uint4(g1[4 * ch], g1[4 * ch + 1], g1[4 * ch + 2], g1[4 * ch + 3]);
    }
    stamp();  // 8: conv4 + gate done
    exchange();
    fence_proxy_async_smem();
    tc_fence_before_sync();
    block_sync();                                   // A (gated, all 512 channels) complete
    stamp();  // 9: gated operand gathered

    // ---------------- conv5 (+gamma) accumulated onto the x slice; next block's norm1 ----------------
    if (ctrl) {
      tc_fence_after_sync();
      gemm(m5, rank * CS, tmem_base + X_COL, 1u, dbar[0]);
    }
    if (!last) {
      const BlockParams nx = args.blocks[b + 1];
      eff_load(nx.ln1_w, nx.ln1_b, nx.mod_off, nx.mod_off + C);
    }
    fb::mbar_wait_c(dbar[0], dph0, args.status, 0xC50u); dph0 ^= 1u;
    tc_fence_after_sync();
    stamp();  // 10: conv5 done
    if (!last) {
      block_sync();                                 // ring idle: R is scratch again
      eff_store();
      block_sync();
      ln_and_gather(bp.cb5);
    }
  }

  // ---------------- final store of the own slice ----------------
  {
    const float* cb = args.blocks[nb - 1].cb5 + rank * CS + hf * 64;
#pragma unroll 1
    for (int c0 = 0; c0 < 2; ++c0) {
      uint32_t t[32];
      tmem_ld32(t_x + c0 * 32, t);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(cb + c0 * 32 + q * 4));
          *reinterpret_cast<float4*>(x_row + c0 * 32 + q * 4) =
              make_float4(__uint_as_float(t[4 * q]) + bb.x, __uint_as_float(t[4 * q + 1]) + bb.y, __uint_as_float(t[4 * q + 2]) + bb.z,
                          __uint_as_float(t[4 * q + 3]) + bb.w);
        }
      }
    }
  }
  tc_fence_before_sync();
  block_sync();
  stamp();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
  cluster_sync_all();  // no CTA of the cluster exits while a peer may still be reading the exchange buffers
}

}  // namespace qb
}  // namespace hd
