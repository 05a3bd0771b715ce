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
//   warps 2-9   epilogue: the CTAs that split K for the same output tile are cluster-mates; each pushes the rows another
//               CTA finishes straight from TMEM registers into that CTA's shared memory (st.async over distributed
//               shared memory, completion counted on the receiver's mbarrier) and adds the partial rows it received
//               in fixed order (deterministic); bias + SimpleGate / SCA multiply / residual add; LayerNorm2d + AdaLN modulation of the
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
// 16-byte store into the shared memory of a cluster-mate; completes 16 bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ uint32_t map_shared_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  return remote;
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
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
  const CUtensorMap* map;  // this cursor's operand of the current phase
  int row0;                // first row of the operand tile (A: mt * 128, W: nt * 128)
  bool is_w;
  bool done;
};
__device__ __forceinline__ void cursor_settle(Cursor& c, const Args& a) {
  // moves (ph, it) forward to the next valid unit; kb is reset by the caller
  while (c.ph < a.n_phases) {
    const Phase& P = a.phases[c.ph];
    const int iters = phase_iters(P);
    while (c.it < iters) {
      c.u = make_unit(P, c.it);
      if (c.u.valid) {
        c.map = a.maps + (c.is_w ? P.map_w : P.map_a);
        c.row0 = (c.is_w ? c.u.nt : c.u.mt) * 128;
        return;
      }
      ++c.it;
    }
    ++c.ph;
    c.it = 0;
  }
  c.done = true;
}
__device__ __forceinline__ void cursor_init(Cursor& c, const Args& a, bool is_w) {
  c.ph = 0; c.it = 0; c.kb = 0; c.done = false; c.is_w = is_w; c.map = nullptr; c.row0 = 0;
  cursor_settle(c, a);
}
__device__ __forceinline__ void cursor_next(Cursor& c, const Args& a) {
  if (++c.kb < c.u.kb_count) return;
  c.kb = 0;
  ++c.it;
  cursor_settle(c, a);
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(320, 1) level_chain_kernel(const Args args) {
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
  uint64_t* phase_go = reads_done + 1;      // the producer has seen the grid barrier that opens the next phase
  uint64_t* recv_full = phase_go + 1;       // the cluster-mates' partial rows have landed in this CTA's receive slots
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(recv_full + 1);
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
    mbar_init(smem_u32(phase_go), 1);
    mbar_init(smem_u32(recv_full), 1);
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
      // The only thread of the CTA that polls the grid-barrier counter (polling from every warp slows the arrivals
      // down); the epilogue warps wait on the local phase_go barrier it signals.
      for (int i = 0; i < 3; ++i) prefetch_tensormap(args.maps + i);
      Cursor ca, cw;
      cursor_init(ca, args, false);
      cursor_init(cw, args, true);
      uint32_t a_slot = 0, a_par = 0, a_fills = 0;   // A ring: next slot, its parity, fills so far
      uint32_t w_slot = 0, w_par = 0, w_fills = 0;
      uint32_t seq_a = 0, seq_w = 0;                 // items issued by each cursor
      const CUtensorMap* w_map_seen = nullptr;
      auto issue_w = [&](bool blocking) -> bool {
        if (cw.done) return false;
        const uint32_t eb = smem_u32(&empty_w[w_slot]);
        if (w_fills >= static_cast<uint32_t>(NW)) {
          if (blocking) mbar_wait(eb, w_par ^ 1u, status, 0xA10u);
          else if (!mbar_try_wait(eb, w_par ^ 1u)) return false;
        }
        if (cw.map != w_map_seen) {  // the descriptor of the phase after this one: fetched a whole phase ahead
          w_map_seen = cw.map;
          if (cw.ph + 1 < args.n_phases) prefetch_tensormap(args.maps + args.phases[cw.ph + 1].map_w);
        }
        const uint32_t fb = smem_u32(&full_w[w_slot]);
        mbar_expect_tx(fb, TILE_BYTES);
        tma_load_2d(smem_u32(smem + OFF_W + w_slot * TILE_BYTES), cw.map, (cw.u.kb_begin + cw.kb) * BK, cw.row0, fb);
        ++w_fills; ++seq_w;
        if (++w_slot == static_cast<uint32_t>(NW)) { w_slot = 0; w_par ^= 1u; }
        cursor_next(cw, args);
        return true;
      };
      for (int ph = 0; ph < args.n_phases; ++ph) {
        if (ph > 0) {
          // the A operand of phase ph is written by the epilogues of phase ph - 1: wait for the whole grid,
          // streaming as many weight tiles as the W ring takes first
          while (issue_w(false)) {}
          wait_counter(bar_phase, static_cast<unsigned int>(GRID) * ph, status, 0xA20u);
          fence_proxy_async_all();
          mbar_arrive_local(smem_u32(phase_go));
          if (trace != nullptr) trace[ph * 8 + 0] = clock64();
        }
        while (!ca.done && ca.ph == ph) {
          while (seq_w <= seq_a) { if (!issue_w(true)) break; }  // this item's W tile is in flight
          const uint32_t eb = smem_u32(&empty_a[a_slot]);
          if (a_fills >= static_cast<uint32_t>(NA) && !mbar_try_wait(eb, a_par ^ 1u)) {
            while (issue_w(false)) {}
            mbar_wait(eb, a_par ^ 1u, status, 0xA30u);
          }
          const uint32_t fb = smem_u32(&full_a[a_slot]);
          mbar_expect_tx(fb, TILE_BYTES);
          tma_load_2d(smem_u32(smem + OFF_A + a_slot * TILE_BYTES), ca.map, (ca.u.kb_begin + ca.kb) * BK, ca.row0, fb);
          ++a_fills; ++seq_a;
          if (++a_slot == static_cast<uint32_t>(NA)) { a_slot = 0; a_par ^= 1u; }
          cursor_next(ca, args);
          issue_w(false);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc(BM, 128);
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0, acc_it = 0;
      for (int ph = 0; ph < args.n_phases; ++ph) {
        const Phase P = args.phases[ph];
        const int iters = phase_iters(P);
        for (int it = 0; it < iters; ++it) {
          const Unit u = make_unit(P, it);
          if (!u.valid) continue;
          const uint32_t buf = acc_it & 1u;
          mbar_wait(smem_u32(&tmem_empty[buf]), ((acc_it >> 1) & 1u) ^ 1u, status, 0xB10u);
          tc_fence_after_sync();
          for (int kb = 0; kb < u.kb_count; ++kb) {
            mbar_wait(smem_u32(&full_w[sw]), pw, status, 0xB20u);
            mbar_wait(smem_u32(&full_a[sa]), pa, status, 0xB30u);
            tc_fence_after_sync();
            const uint64_t da = make_smem_desc(smem_u32(smem + OFF_A + sa * TILE_BYTES));
            const uint64_t db = make_smem_desc(smem_u32(smem + OFF_W + sw * TILE_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(da + 2 * k, db + 2 * k, tmem_base + buf * 128u, (kb | k) != 0 ? 1u : 0u, idesc);
            umma_commit(smem_u32(&empty_a[sa]));
            umma_commit(smem_u32(&empty_w[sw]));
            if (++sa == static_cast<uint32_t>(NA)) { sa = 0; pa ^= 1u; }
            if (++sw == static_cast<uint32_t>(NW)) { sw = 0; pw ^= 1u; }
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
    //
    // K-split reduction by PUSH: the staging buffer is S receive slots of 128/S rows.  The CTA with split rank z
    // finishes rows [z * 128/S, (z+1) * 128/S) of the tile; every CTA of the group sends those rows of ITS partial
    // accumulator straight from registers into slot (its own z) of that CTA — st.async over distributed shared
    // memory, counted in bytes on the receiver's mbarrier — and keeps its own rows with plain stores.  The receiver
    // then adds the S slots in slot order (fixed: deterministic) out of its own shared memory.  (Pulling the partial
    // tiles with ld.shared::cluster instead measured 8-10 k clocks per 128x128 tile: remote loads are latency-bound.)
    const int ew = warp - 2;                 // 0..7
    const int quad = warp & 3;               // TMEM lane quadrant this warp may read
    const int chalf = ew >> 2;               // which 64 accumulator columns this warp drains
    uint32_t acc_it = 0, rx_it = 0, hs_it = 0;   // accumulators consumed; receive phases; read-done handshakes
    bool hs_pending = false;                 // cluster-mates may still be reading... writing: see reads_done below
    unsigned int ln_target = 0;
    const uint32_t stage_u32 = smem_u32(stage);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ph = 0; ph < args.n_phases; ++ph) {
      const Phase P = args.phases[ph];
      const int iters = phase_iters(P);
      const int S = P.split;
      const int rows_mine = 128 / S;
      const uint32_t slot_bytes = static_cast<uint32_t>(STG_BYTES / S);
      const uint32_t gbase = (crank / S) * S;   // first cluster rank of this CTA's K-split group
      const bool gate = P.kind == LV_GATE;
      // cross-CTA data of the previous phase (mul source, residual rows) is visible once the producer saw the barrier
      if (ph > 0) mbar_wait(smem_u32(phase_go), (ph - 1) & 1u, status, 0xC10u);
      for (int it = 0; it < iters; ++it) {
        const Unit u = make_unit(P, it);
        const int ncol = u.nt * 128;
        const int row_base = u.z * rows_mine;
        // rows of the non-gate epilogues: one row per warp pass, lanes along the 32 four-column chunks
        constexpr int MAXP = 8;
        const int passes = rows_mine / EW;       // non-gate: 4 / 8 (split 4 / 2)
        float4 e_pre[MAXP];
        if (u.valid && S > 1 && threadIdx.x == 64) mbar_expect_tx(smem_u32(recv_full), (S - 1) * slot_bytes);
        if (u.valid && !gate) {
          // residual rows / SCA multiplicand of this CTA's rows: fetched while the MMAs are still running
#pragma unroll
          for (int p = 0; p < MAXP; ++p) {
            e_pre[p] = zero4;
            const int m = u.mt * 128 + row_base + p * EW + ew;
            if (p < passes && m < args.rows) {
              if (P.kind == LV_RESID) {
                e_pre[p] = __ldcg(reinterpret_cast<const float4*>(P.x + static_cast<size_t>(m) * P.N + ncol + lane * 4));
              } else {
                const uint2 g = __ldcg(reinterpret_cast<const uint2*>(P.mul + static_cast<size_t>(m) * P.N + ncol + lane * 4));
                const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y);
                e_pre[p] = make_float4(g0.x, g0.y, g1.x, g1.y);
              }
            }
          }
        }
        if (u.valid) {
          const uint32_t buf = acc_it & 1u;
          if (hs_pending) {  // (several units per phase only) the group has consumed the previous unit's slots
            mbar_wait_cluster(smem_u32(reads_done), (hs_it - 1u) & 1u, status, 0xC20u);
            hs_pending = false;
          }
          mbar_wait(smem_u32(&tmem_full[buf]), (acc_it >> 1) & 1u, status, 0xC30u);
          tc_fence_after_sync();
          if (trace != nullptr && threadIdx.x == 64) trace[ph * 8 + 1] = clock64();
          // ---- TMEM -> receive slots: own rows locally, the other rows into their owners' shared memory ----
          {
            const int r = quad * 32 + lane;                 // accumulator row of this thread
            const int zdst = r / rows_mine;                 // split rank that finishes this row
            const int rl = r - zdst * rows_mine;            // row inside its owner's slot
            const uint32_t off = static_cast<uint32_t>(u.z) * slot_bytes + static_cast<uint32_t>(rl) * 512u;
            const bool local = zdst == u.z;
            const uint32_t dst_base = local ? stage_u32 + off : map_shared_rank(stage_u32 + off, gbase + zdst);
            const uint32_t dst_bar = local ? 0u : map_shared_rank(smem_u32(recv_full), gbase + zdst);
            const uint32_t taddr = tmem_base + buf * 128u + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(taddr + chalf * 64 + c0, v);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t ck = static_cast<uint32_t>(((chalf * 64 + c0) >> 2) + j);
                const uint32_t a = dst_base + ((ck ^ static_cast<uint32_t>(rl & 7)) << 4);
                if (local) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
                else st_async_v4(a, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], dst_bar);
              }
            }
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_local(smem_u32(&tmem_empty[buf]));
          ++acc_it;
          if (trace != nullptr && threadIdx.x == 64) trace[ph * 8 + 6] = clock64();
        }
        // ---- own rows staged by every warp of this CTA; the cluster-mates' rows have landed ----
        epi_bar_sync();
        if (u.valid && S > 1) {
          mbar_wait_cluster(smem_u32(recv_full), rx_it & 1u, status, 0xC40u);
          ++rx_it;
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
              const int rl = ((p0 + q) * EW + ew) * 2 + sub;
              const uint32_t a1 = stage_u32 + static_cast<uint32_t>(rl * 512 + ((sl ^ (rl & 7)) << 4));
              const uint32_t a2 = stage_u32 + static_cast<uint32_t>(rl * 512 + (((sl + 16) ^ (rl & 7)) << 4));
#pragma unroll
              for (int s = 0; s < CL; ++s) {
                if (s < S) {
                  part[q][s] = ld_shared_f4(a1 + s * slot_bytes);
                  part2[q][s] = ld_shared_f4(a2 + s * slot_bytes);
                }
              }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int rl = ((p0 + q) * EW + ew) * 2 + sub;
              const int m = u.mt * 128 + row_base + rl;
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
                const int rl = (p0 + q) * EW + ew;
                const uint32_t a1 = stage_u32 + static_cast<uint32_t>(rl * 512 + ((lane ^ (rl & 7)) << 4));
#pragma unroll
                for (int s = 0; s < CL; ++s)
                  if (s < S) part[q][s] = ld_shared_f4(a1 + s * slot_bytes);
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int rl = (p0 + q) * EW + ew;
                const int m = u.mt * 128 + row_base + rl;
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
        if (S > 1 && iters > 1) {
          // another unit of this phase follows without a grid barrier in between: the group agrees that everybody
          // has read its slots before anybody pushes the next partial rows
          epi_bar_sync();
          if (threadIdx.x == 64) {
#pragma unroll
            for (uint32_t r = 0; r < static_cast<uint32_t>(CL); ++r) mbar_arrive_remote(smem_u32(reads_done), r);
          }
          hs_pending = true;
          ++hs_it;
        }
      }
      if (hs_pending) {  // leave the phase with the handshake barrier consumed
        mbar_wait_cluster(smem_u32(reads_done), (hs_it - 1u) & 1u, status, 0xC25u);
        hs_pending = false;
      }
      if (P.kind == LV_RESID && P.ln) {
        // ---- grid barrier among the epilogue warps, then LayerNorm2d + modulation of the finished rows ----
        // (bar.sync orders every warp's stores before thread 64's gpu-scope fence: one fence releases them all)
        epi_bar_sync();
        ln_target += GRID;
        if (threadIdx.x == 64) {
          __threadfence();
          atomicAdd(bar_ln, 1u);
          wait_counter(bar_ln, ln_target, status, 0xC50u);
        }
        epi_bar_sync();
        if (trace != nullptr && threadIdx.x == 64) trace[ph * 8 + 4] = clock64();
        const int T = P.n_tiles;
        const float inv_t = 1.f / static_cast<float>(T);
        const float inv_n = 1.f / static_cast<float>(P.N);
        for (int it = 0; it < iters; ++it) {
          const Unit u = make_unit(P, it);
          if (!u.valid) continue;
          const int ncol = u.nt * 128 + lane * 4;
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(P.ln_w + ncol));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.ln_b + ncol));
          // four rows per batch: all loads of the batch are issued before the first reduction
#pragma unroll 1
          for (int rr0 = ew; rr0 < rows_mine; rr0 += 4 * EW) {
            float4 xv[4], sc[4], sh[4];
            float2 st[4];
            int mrow[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int rr = rr0 + q * EW;
              const int m = u.mt * 128 + u.z * rows_mine + rr;
              mrow[q] = (rr < rows_mine && m < args.rows) ? m : -1;
              xv[q] = zero4; sc[q] = zero4; sh[q] = zero4; st[q] = make_float2(0.f, 0.f);
              if (mrow[q] >= 0) {
                xv[q] = __ldcg(reinterpret_cast<const float4*>(P.x + static_cast<size_t>(m) * P.N + ncol));
                if (lane < T) st[q] = __ldcg(args.stats + static_cast<size_t>(m) * 32 + lane);
                const float* mr = args.mod_table + static_cast<size_t>(args.mod_row_idx[m / args.rows_per_face]) * args.mod_stride;
                sc[q] = __ldg(reinterpret_cast<const float4*>(mr + P.scale_off + ncol));
                sh[q] = __ldg(reinterpret_cast<const float4*>(mr + P.shift_off + ncol));
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float mu = warp_sum(st[q].x) * inv_t;
              const float dm = lane < T ? st[q].x - mu : 0.f;
              const float m2 = warp_sum(st[q].y + 128.f * dm * dm);
              const float denom = sqrtf(m2 * inv_n + 1e-6f);
              float4 y;
              y.x = (w4.x * ((xv[q].x - mu) / denom) + b4.x) * (sc[q].x + 1.f) + sh[q].x;
              y.y = (w4.y * ((xv[q].y - mu) / denom) + b4.y) * (sc[q].y + 1.f) + sh[q].y;
              y.z = (w4.z * ((xv[q].z - mu) / denom) + b4.z) * (sc[q].z + 1.f) + sh[q].z;
              y.w = (w4.w * ((xv[q].w - mu) / denom) + b4.w) * (sc[q].w + 1.f) + sh[q].w;
              if (mrow[q] >= 0) store4<bf16>(P.out + static_cast<size_t>(mrow[q]) * P.N + ncol, y);
            }
          }
        }
      }
      // ---- end of phase: this CTA's global writes are published, one arrival on the grid counter ----
      fence_proxy_async_all();  // this thread's generic-proxy stores vs. the TMA (async-proxy) reads of the next phase
      epi_bar_sync();
      if (threadIdx.x == 64) {
        __threadfence();
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
