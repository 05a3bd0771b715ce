// Persistent "level chain" kernel (v2): a whole run of ConditionalNAFBlocks at one of the small-spatial levels
// (4x4 / 2x2 / 1x1, c = 512 / 1024 / 2048) executed by ONE launch of 128 co-resident CTAs (32 clusters of 4).
//
// Why.  At B = 256 the GEMMs of these levels are 3-9 us of mainloop wrapped in ~4 us of per-launch fixed cost
// (launch, TMEM allocation, barrier setup, first-operand latency, drain), nine launches per block.  The chain keeps
// the CTAs, their TMEM and their barriers alive across all GEMMs of the run and replaces kernel boundaries by grid
// barriers that only the warps that need them wait on:
//
//   warp 0      TMA producer.  Walks the whole (phase, unit, k-block) sequence with TWO cursors: the W cursor runs
//               ahead — weights depend on nothing, so the next phase's weight tiles stream into their own ring while
//               this phase is still in its epilogue / barrier — and the A cursor waits at every phase boundary for the
//               grid barrier (the A operand is the previous phase's output).
//   warp 1      MMA issuer: tcgen05.mma M=128 N=128 K=16 from separate A / W rings into one of two TMEM accumulators.
//   warps 2-9   epilogue: TMEM -> fp32 staging tile in shared memory; the CTAs that split K for the same output tile
//               are cluster-mates and exchange partial tiles over distributed shared memory (fixed summation order:
//               deterministic); bias + SimpleGate / SCA multiply / residual add; LayerNorm2d + AdaLN modulation of the
//               finished residual rows (per-tile (mean, M2) statistics merged across the N tiles after one extra grid
//               barrier); the next GEMM's bf16 A operand is written to global memory (L2-resident).
//
// Reference arithmetic: models/denoiser/conditional_naf.py:108-136 (block), utils.py:16-24 (LayerNorm2d).
#pragma once

#include "common.cuh"
#include "gemm_tc.cuh"

namespace hd {
namespace lv {

enum EpiKind : int {
  LV_GATE = 0,   // out_bf16[m, nt*64 + j] = (acc[m, j] + b[j]) * (acc[m, 64 + j] + b[64 + j])   gate-packed 128-column groups
  LV_MUL = 1,    // out_bf16[m, n] = (acc + b[n]) * mul_bf16[m, n]          SCA at 1x1 spatial (pooled mean == the tensor itself)
  LV_RESID = 2,  // x[m, n] += acc + b[n];  ln != 0: out_bf16[m, :] = modulate(LayerNorm2d(x[m, :]))
};

struct Phase {
  int kind;
  int map_a, map_w;              // indices into the tensor-map array
  int m_tiles, n_tiles, num_kb;  // 128-row tiles, 128-column tiles of the (packed) N, K / 64
  int split;                     // K split over `split` cluster-mates: 1, 2 or 4
  int N;                         // packed GEMM width (n_tiles * 128)
  const float* bias;             // [N] in packed column order
  bf16* out;                     // GATE: [rows, N/2]; MUL / RESID+ln: [rows, N]
  const bf16* mul;               // MUL: [rows, N]
  float* x;                      // RESID: fp32 residual stream [rows, N], updated in place
  int ln;                        // RESID: LayerNorm2d + modulation of the finished rows -> out
  const float* ln_w;
  const float* ln_b;
  int shift_off, scale_off;      // offsets of the block's modulation vectors in a table row
};

struct Args {
  const Phase* phases;
  int n_phases;
  const CUtensorMap* maps;
  int rows;                      // valid rows (faces * pixels per face)
  int rows_per_face;
  float2* stats;                 // [rows][32] per-tile (mean, M2) of the finished residual rows
  unsigned int* sync;            // [0]: phase barrier counter, [32]: LayerNorm barrier counter (zeroed before the launch)
  const float* mod_table;
  const int* mod_row_idx;
  int mod_stride;
  DeviceStatus* status;
  long long* trace;              // optional [CTA][phase][8] clock64 timeline, nullptr in production
};

constexpr int CL = 4;                         // cluster size = largest K split
constexpr int GRID = 128;                     // co-resident CTAs (132 is the most clusters of 4 the chip takes)
constexpr int EW = 8;                         // epilogue warps
constexpr int THREADS = 64 + 32 * EW;
constexpr int NA = 4, NW = 6;                 // A ring / W ring depth (16 KB tiles)
constexpr int TILE_BYTES = tc::BM * tc::BK * 2;
constexpr int STG_BYTES = 128 * 128 * 4;      // fp32 staging tile
constexpr int OFF_A = 0;
constexpr int OFF_W = NA * TILE_BYTES;
constexpr int OFF_STG = (NA + NW) * TILE_BYTES;
constexpr int OFF_BAR = OFF_STG + STG_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;  // + barriers + 1 KB alignment slack
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// ---- small PTX helpers on top of gemm_tc.cuh's -----------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity, DeviceStatus* st, uint32_t site) {
  if (mbar_try_wait_cluster(bar, parity)) return true;
  const long long t0 = clock64();
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait_cluster(bar, parity)) return true;
    if ((it & 0xFFFu) == 0u) {
      if (*reinterpret_cast<volatile unsigned int*>(&st->error) != 0u) return false;
      if (clock64() - t0 > 4000000000ll) {
        if (atomicCAS(&st->error, 0u, 3u) == 0u) st->where = site;
        return false;
      }
    }
  }
}
// spin until *counter >= target (grid barrier wait side); bounded like mbar_wait
__device__ __forceinline__ bool wait_counter(const unsigned int* counter, unsigned int target, DeviceStatus* st, uint32_t site) {
  if (ld_acquire_gpu(counter) >= target) return true;
  const long long t0 = clock64();
  for (uint32_t it = 1;; ++it) {
    if (ld_acquire_gpu(counter) >= target) return true;
    if ((it & 0x3FFu) == 0u) {
      if (*reinterpret_cast<volatile unsigned int*>(&st->error) != 0u) return false;
      if (clock64() - t0 > 4000000000ll) {
        if (atomicCAS(&st->error, 0u, 2u) == 0u) st->where = site;
        return false;
      }
    }
  }
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory"); }

// One unit of GEMM work of this CTA: (phase, iteration) -> tile and K range.  Both producer cursors, the MMA issuer
// and the epilogue walk the same sequence.
struct Unit {
  int mt, nt, z, kb_begin, kb_count;
  bool valid;
};
__device__ __forceinline__ int phase_iters(const Phase& P) {
  const int units = P.m_tiles * P.n_tiles * P.split;
  return (units + GRID - 1) / GRID;
}
__device__ __forceinline__ Unit make_unit(const Phase& P, int iter) {
  Unit u;
  const int units = P.m_tiles * P.n_tiles * P.split;
  const int id = static_cast<int>(blockIdx.x) + iter * GRID;
  u.valid = id < units;
  u.z = id % P.split;  // == cluster rank % split (GRID % CL == 0)
  const int t = id / P.split;
  u.mt = t % P.m_tiles;
  u.nt = t / P.m_tiles;
  u.kb_count = P.num_kb / P.split;
  u.kb_begin = u.z * u.kb_count;
  return u;
}

// cursor over the (phase, iteration, k-block) sequence of this CTA
struct Cursor {
  int ph, it, kb;
  Unit u;
  bool done;
};
__device__ __forceinline__ void cursor_settle(Cursor& c, const Args& a) {
  // moves (ph, it) forward to the next valid unit; kb is reset by the caller
  while (c.ph < a.n_phases) {
    const Phase& P = a.phases[c.ph];
    const int iters = phase_iters(P);
    while (c.it < iters) {
      c.u = make_unit(P, c.it);
      if (c.u.valid) return;
      ++c.it;
    }
    ++c.ph;
    c.it = 0;
  }
  c.done = true;
}
__device__ __forceinline__ void cursor_init(Cursor& c, const Args& a) {
  c.ph = 0; c.it = 0; c.kb = 0; c.done = false;
  cursor_settle(c, a);
}
__device__ __forceinline__ void cursor_next(Cursor& c, const Args& a) {
  if (++c.kb < c.u.kb_count) return;
  c.kb = 0;
  ++c.it;
  cursor_settle(c, a);
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1) level_chain_kernel(const Args args) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full_a = bars;                  // [NA]
  uint64_t* empty_a = full_a + NA;          // [NA]
  uint64_t* full_w = empty_a + NA;          // [NW]
  uint64_t* empty_w = full_w + NW;          // [NW]
  uint64_t* tmem_full = empty_w + NW;       // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint64_t* stage_ready = tmem_empty + 2;   // all CL cluster-mates have staged their partial tile
  uint64_t* reads_done = stage_ready + 1;   // all CL cluster-mates have finished reading the staged tiles
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(reads_done + 1);
  float* stage = reinterpret_cast<float*>(smem + OFF_STG);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  DeviceStatus* status = args.status;
  long long* trace = args.trace != nullptr ? args.trace + static_cast<size_t>(blockIdx.x) * args.n_phases * 8 : nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NA; ++s) { mbar_init(smem_u32(&full_a[s]), 1); mbar_init(smem_u32(&empty_a[s]), 1); }
    for (int s = 0; s < NW; ++s) { mbar_init(smem_u32(&full_w[s]), 1); mbar_init(smem_u32(&empty_w[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tmem_full[s]), 1); mbar_init(smem_u32(&tmem_empty[s]), EW); }
    mbar_init(smem_u32(stage_ready), CL);
    mbar_init(smem_u32(reads_done), CL);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  cluster_sync_all();  // every CTA's barriers exist before a cluster-mate can arrive on them
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  unsigned int* bar_phase = args.sync;
  unsigned int* bar_ln = args.sync + 32;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer: W cursor ahead of A cursor =================
      Cursor ca, cw;
      cursor_init(ca, args);
      cursor_init(cw, args);
      uint32_t a_it = 0, w_it = 0;   // k-blocks pushed into each ring
      uint32_t seq_a = 0, seq_w = 0; // items issued by each cursor
      int synced = 0;                // phases [0, synced) are known complete grid-wide
      auto issue_w = [&](bool blocking) -> bool {
        if (cw.done) return false;
        const int s = w_it % NW;
        const uint32_t ph = (w_it / NW) & 1;
        const uint32_t eb = smem_u32(&empty_w[s]);
        if (w_it >= static_cast<uint32_t>(NW)) {
          if (blocking) { if (!mbar_wait(eb, ph ^ 1u, status, 0xA10u)) { cw.done = true; return false; } }
          else if (!mbar_try_wait(eb, ph ^ 1u)) return false;
        }
        const int map_w = args.phases[cw.ph].map_w;
        const uint32_t fb = smem_u32(&full_w[s]);
        mbar_expect_tx(fb, TILE_BYTES);
        tma_load_2d(smem_u32(smem + OFF_W + s * TILE_BYTES), args.maps + map_w, (cw.u.kb_begin + cw.kb) * BK, cw.u.nt * 128, fb);
        ++w_it; ++seq_w;
        cursor_next(cw, args);
        return true;
      };
      while (!ca.done) {
        if (ca.ph > synced) {
          // the A operand of phase ca.ph is written by the epilogues of phase ca.ph - 1: wait for the whole grid,
          // streaming as many weight tiles as the W ring takes first
          while (issue_w(false)) {}
          if (!wait_counter(bar_phase, static_cast<unsigned int>(GRID) * ca.ph, status, 0xA20u)) break;
          __threadfence();
          fence_proxy_async_all();
          synced = ca.ph;
          if (trace != nullptr) trace[ca.ph * 8 + 0] = clock64();
        }
        while (seq_w <= seq_a) { if (!issue_w(true)) break; }  // this item's W tile is in flight
        const int s = a_it % NA;
        const uint32_t ph = (a_it / NA) & 1;
        const uint32_t eb = smem_u32(&empty_a[s]);
        if (a_it >= static_cast<uint32_t>(NA) && !mbar_try_wait(eb, ph ^ 1u)) {
          while (issue_w(false)) {}
          if (!mbar_wait(eb, ph ^ 1u, status, 0xA30u)) break;
        }
        const int map_a = args.phases[ca.ph].map_a;
        const uint32_t fb = smem_u32(&full_a[s]);
        mbar_expect_tx(fb, TILE_BYTES);
        tma_load_2d(smem_u32(smem + OFF_A + s * TILE_BYTES), args.maps + map_a, (ca.u.kb_begin + ca.kb) * BK, ca.u.mt * 128, fb);
        ++a_it; ++seq_a;
        cursor_next(ca, args);
        issue_w(false);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc(BM, 128);
      uint32_t a_it = 0, w_it = 0, acc_it = 0;
      bool ok = true;
      for (int ph = 0; ph < args.n_phases && ok; ++ph) {
        const Phase P = args.phases[ph];
        const int iters = phase_iters(P);
        for (int it = 0; it < iters && ok; ++it) {
          const Unit u = make_unit(P, it);
          if (!u.valid) continue;
          const uint32_t buf = acc_it & 1u;
          ok = mbar_wait(smem_u32(&tmem_empty[buf]), ((acc_it >> 1) & 1u) ^ 1u, status, 0xB10u);
          tc_fence_after_sync();
          for (int kb = 0; kb < u.kb_count && ok; ++kb) {
            const int sa = a_it % NA, sw = w_it % NW;
            ok = mbar_wait(smem_u32(&full_w[sw]), (w_it / NW) & 1u, status, 0xB20u) &&
                 mbar_wait(smem_u32(&full_a[sa]), (a_it / NA) & 1u, status, 0xB30u);
            tc_fence_after_sync();
            const uint64_t da = make_smem_desc(smem_u32(smem + OFF_A + sa * TILE_BYTES));
            const uint64_t db = make_smem_desc(smem_u32(smem + OFF_W + sw * TILE_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(da + 2 * k, db + 2 * k, tmem_base + buf * 128u, (kb | k) != 0 ? 1u : 0u, idesc);
            umma_commit(smem_u32(&empty_a[sa]));
            umma_commit(smem_u32(&empty_w[sw]));
            ++a_it; ++w_it;
          }
          umma_commit(smem_u32(&tmem_full[buf]));
          ++acc_it;
        }
      }
    }
  } else {
    // ================= epilogue warps =================
    // Control flow here never depends on whether a wait succeeded: a tripped watchdog sets the status word, every
    // later wait then gives up at once, and all warps still meet at every barrier (garbage out, HD_ERR_KERNEL on
    // the host) instead of hanging the GPU.
    const int ew = warp - 2;                 // 0..7
    const int quad = warp & 3;               // TMEM lane quadrant this warp may read
    const int chalf = ew >> 2;               // which 64 accumulator columns this warp drains
    uint32_t acc_it = 0, hs_it = 0;          // accumulators consumed; cluster handshakes done
    bool hs_pending = false;                 // peers may still be reading this CTA's staging tile
    unsigned int ln_target = 0;
    const uint32_t stage_u32 = smem_u32(stage);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ph = 0; ph < args.n_phases; ++ph) {
      const Phase P = args.phases[ph];
      const int iters = phase_iters(P);
      const int S = P.split;
      const int rows_mine = 128 / S;
      const uint32_t gbase = (crank / S) * S;   // first cluster rank of this CTA's K-split group
      const bool gate = P.kind == LV_GATE;
      // cross-CTA data of the previous phase (mul source, residual rows) must be visible to this warp's loads
      if (ph > 0) {
        if (lane == 0) wait_counter(bar_phase, static_cast<unsigned int>(GRID) * ph, status, 0xC10u);
        __syncwarp();
      }
      for (int it = 0; it < iters; ++it) {
        const Unit u = make_unit(P, it);
        const int ncol = u.nt * 128;
        const int row_base = u.z * rows_mine;
        // rows of the non-gate epilogues: one row per warp pass, lanes along the 32 four-column chunks
        constexpr int MAXP = 16;
        const int passes = rows_mine / EW;       // non-gate: 4 / 8 / 16
        float4 e_pre[MAXP];
        if (u.valid && !gate) {
          // residual rows / SCA multiplicand of this CTA's rows: fetched while the MMAs are still running
#pragma unroll
          for (int p = 0; p < MAXP; ++p) {
            e_pre[p] = zero4;
            const int m = u.mt * 128 + row_base + p * EW + ew;
            if (p < passes && m < args.rows) {
              if (P.kind == LV_RESID) {
                e_pre[p] = *reinterpret_cast<const float4*>(P.x + static_cast<size_t>(m) * P.N + ncol + lane * 4);
              } else {
                const uint2 g = *reinterpret_cast<const uint2*>(P.mul + static_cast<size_t>(m) * P.N + ncol + lane * 4);
                const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y);
                e_pre[p] = make_float4(g0.x, g0.y, g1.x, g1.y);
              }
            }
          }
        }
        if (u.valid) {
          const uint32_t buf = acc_it & 1u;
          if (hs_pending) {  // peers finished reading the previous partial tile staged here
            mbar_wait_cluster(smem_u32(reads_done), (hs_it - 1u) & 1u, status, 0xC20u);
            hs_pending = false;
          }
          mbar_wait(smem_u32(&tmem_full[buf]), (acc_it >> 1) & 1u, status, 0xC30u);
          tc_fence_after_sync();
          if (trace != nullptr && threadIdx.x == 64) trace[ph * 8 + 1] = clock64();
          // ---- TMEM -> staging (16-byte chunks XOR-swizzled by row) ----
          {
            const int r = quad * 32 + lane;
            const uint32_t taddr = tmem_base + buf * 128u + (static_cast<uint32_t>(quad * 32) << 16);
            float* srow = stage + r * 128;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(taddr + chalf * 64 + c0, v);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int ck = ((chalf * 64 + c0) >> 2) + j;
                *reinterpret_cast<uint4*>(srow + ((ck ^ (r & 7)) << 2)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
            }
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_local(smem_u32(&tmem_empty[buf]));
          ++acc_it;
        }
        // ---- all staged: CTA-wide, then cluster-wide when K is split ----
        epi_bar_sync();
        if (S > 1) {
          if (threadIdx.x == 64) {
#pragma unroll
            for (uint32_t r = 0; r < static_cast<uint32_t>(CL); ++r) mbar_arrive_remote(smem_u32(stage_ready), r);
          }
          mbar_wait_cluster(smem_u32(stage_ready), hs_it & 1u, status, 0xC40u);
        }
        if (trace != nullptr && threadIdx.x == 64) trace[ph * 8 + 2] = clock64();
        if (u.valid && gate) {
          // ---- SimpleGate: two rows per warp pass, 16 lanes per row; x1 chunk j pairs with x2 chunk j + 16 ----
          const int sub = lane >> 4, sl = lane & 15;
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(P.bias + ncol + sl * 4));
          const float4 b2 = __ldg(reinterpret_cast<const float4*>(P.bias + ncol + 64 + sl * 4));
          const int gpasses = rows_mine / (2 * EW);   // 2 / 4 / 8
#pragma unroll 1
          for (int p0 = 0; p0 < gpasses; p0 += 2) {
            float4 part[2][CL], part2[2][CL];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int r = row_base + ((p0 + q) * EW + ew) * 2 + sub;
              const uint32_t a1 = stage_u32 + static_cast<uint32_t>((r * 128 + ((sl ^ (r & 7)) << 2)) * 4);
              const uint32_t a2 = stage_u32 + static_cast<uint32_t>((r * 128 + (((sl + 16) ^ (r & 7)) << 2)) * 4);
#pragma unroll
              for (int s = 0; s < CL; ++s) {
                if (s < S) {
                  part[q][s] = ld_dsmem_f4(a1, gbase + s);
                  part2[q][s] = ld_dsmem_f4(a2, gbase + s);
                }
              }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int r = row_base + ((p0 + q) * EW + ew) * 2 + sub;
              const int m = u.mt * 128 + r;
              float4 v = zero4, w = zero4;
#pragma unroll
              for (int s = 0; s < CL; ++s) {
                if (s < S) {
                  v.x += part[q][s].x; v.y += part[q][s].y; v.z += part[q][s].z; v.w += part[q][s].w;
                  w.x += part2[q][s].x; w.y += part2[q][s].y; w.z += part2[q][s].z; w.w += part2[q][s].w;
                }
              }
              v.x = (v.x + b1.x) * (w.x + b2.x); v.y = (v.y + b1.y) * (w.y + b2.y);
              v.z = (v.z + b1.z) * (w.z + b2.z); v.w = (v.w + b1.w) * (w.w + b2.w);
              if (m < args.rows) store4<bf16>(P.out + static_cast<size_t>(m) * (P.N >> 1) + (ncol >> 1) + sl * 4, v);
            }
          }
        } else if (u.valid) {
          // ---- SCA multiply / residual add (+ per-tile LayerNorm statistics): one row per warp pass ----
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(P.bias + ncol + lane * 4));
#pragma unroll
          for (int p0 = 0; p0 < MAXP; p0 += 4) {
            if (p0 < passes) {
              float4 part[4][CL];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int r = row_base + (p0 + q) * EW + ew;
                const uint32_t a1 = stage_u32 + static_cast<uint32_t>((r * 128 + ((lane ^ (r & 7)) << 2)) * 4);
#pragma unroll
                for (int s = 0; s < CL; ++s)
                  if (s < S) part[q][s] = ld_dsmem_f4(a1, gbase + s);
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int r = row_base + (p0 + q) * EW + ew;
                const int m = u.mt * 128 + r;
                const bool live = m < args.rows;
                float4 v = zero4;
#pragma unroll
                for (int s = 0; s < CL; ++s)
                  if (s < S) { v.x += part[q][s].x; v.y += part[q][s].y; v.z += part[q][s].z; v.w += part[q][s].w; }
                v.x += b1.x; v.y += b1.y; v.z += b1.z; v.w += b1.w;
                const float4 e = e_pre[p0 + q];
                if (P.kind == LV_MUL) {
                  v.x *= e.x; v.y *= e.y; v.z *= e.z; v.w *= e.w;
                  if (live) store4<bf16>(P.out + static_cast<size_t>(m) * P.N + ncol + lane * 4, v);
                } else {
                  v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
                  if (live) *reinterpret_cast<float4*>(P.x + static_cast<size_t>(m) * P.N + ncol + lane * 4) = v;
                  if (P.ln) {
                    // statistics of this row's 128 columns (two-pass inside the tile); merged after the barrier
                    const float mean_t = warp_sum(v.x + v.y + v.z + v.w) * (1.f / 128.f);
                    const float d0 = v.x - mean_t, d1 = v.y - mean_t, d2 = v.z - mean_t, d3 = v.w - mean_t;
                    const float m2_t = warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
                    if (lane == 0 && live) args.stats[static_cast<size_t>(m) * 32 + u.nt] = make_float2(mean_t, m2_t);
                  }
                }
              }
            }
          }
        }
        if (trace != nullptr && threadIdx.x == 64) trace[ph * 8 + 3] = clock64();
        if (S > 1) {
          epi_bar_sync();  // every warp of this CTA is done with its remote reads
          if (threadIdx.x == 64) {
#pragma unroll
            for (uint32_t r = 0; r < static_cast<uint32_t>(CL); ++r) mbar_arrive_remote(smem_u32(reads_done), r);
          }
          hs_pending = true;
          ++hs_it;
        }
      }
      if (P.kind == LV_RESID && P.ln) {
        // ---- grid barrier among the epilogue warps, then LayerNorm2d + modulation of the finished rows ----
        __threadfence();
        epi_bar_sync();
        ln_target += GRID;
        if (threadIdx.x == 64) {
          atomicAdd(bar_ln, 1u);
          wait_counter(bar_ln, ln_target, status, 0xC50u);
          __threadfence();
        }
        epi_bar_sync();
        if (trace != nullptr && threadIdx.x == 64) trace[ph * 8 + 4] = clock64();
        const int T = P.n_tiles;
        const float inv_n = 1.f / static_cast<float>(P.N);
        for (int it = 0; it < iters; ++it) {
          const Unit u = make_unit(P, it);
          if (!u.valid) continue;
          const int ncol = u.nt * 128 + lane * 4;
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(P.ln_w + ncol));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.ln_b + ncol));
          for (int rr = ew; rr < rows_mine; rr += EW) {
            const int m = u.mt * 128 + u.z * rows_mine + rr;
            if (m >= args.rows) continue;
            const float4 xv = *reinterpret_cast<const float4*>(P.x + static_cast<size_t>(m) * P.N + ncol);
            float2 st = make_float2(0.f, 0.f);
            if (lane < T) st = __ldcg(args.stats + static_cast<size_t>(m) * 32 + lane);
            const float mu = warp_sum(st.x) / static_cast<float>(T);
            const float dm = lane < T ? st.x - mu : 0.f;
            const float m2 = warp_sum(st.y + 128.f * dm * dm);
            const float denom = sqrtf(m2 * inv_n + 1e-6f);
            const float* mrow = args.mod_table + static_cast<size_t>(args.mod_row_idx[m / args.rows_per_face]) * args.mod_stride;
            const float4 sc = __ldg(reinterpret_cast<const float4*>(mrow + P.scale_off + ncol));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(mrow + P.shift_off + ncol));
            float4 y;
            y.x = (w4.x * ((xv.x - mu) / denom) + b4.x) * (sc.x + 1.f) + sh.x;
            y.y = (w4.y * ((xv.y - mu) / denom) + b4.y) * (sc.y + 1.f) + sh.y;
            y.z = (w4.z * ((xv.z - mu) / denom) + b4.z) * (sc.z + 1.f) + sh.z;
            y.w = (w4.w * ((xv.w - mu) / denom) + b4.w) * (sc.w + 1.f) + sh.w;
            store4<bf16>(P.out + static_cast<size_t>(m) * P.N + ncol, y);
          }
        }
      }
      // ---- end of phase: this CTA's global writes are published, one arrival on the grid counter ----
      fence_proxy_async_all();
      __threadfence();
      epi_bar_sync();
      if (threadIdx.x == 64) {
        atomicAdd(bar_phase, 1u);
        if (trace != nullptr) trace[ph * 8 + 5] = clock64();
      }
    }
  }

  // no CTA may exit (or free TMEM) while a cluster-mate can still read its staging tile or arrive on its barriers
  __syncwarp();
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace lv
}  // namespace hd
