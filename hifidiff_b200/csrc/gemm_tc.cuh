// tcgen05 / TMEM / TMA GEMM for sm_100a:   out = epilogue(A[M,K] * W[N,K]^T), bf16 in, fp32 accumulate.
//
// One CTA computes a 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      TMA producer: A tile (128 x 64 bf16) and W tile (BN x 64 bf16) per k-block into a
//               multi-stage shared-memory ring, 128B-swizzled, completion on an mbarrier.
//   warp 1      allocates TMEM, issues tcgen05.mma (M=128, N=BN, K=16, kind::f16) from one thread,
//               frees ring slots with tcgen05.commit.
//   warps 2-9   epilogue: tcgen05.ld the fp32 accumulator (one TMEM lane = one output row; two warps per
//               lane quadrant, half the columns each) into a staging tile, then apply bias / ReLU /
//               residual add / SimpleGate / PixelShuffle scatter row-coalesced.  Eight warps because the
//               epilogue is issue-latency-bound: rows are independent, more warps = more rows in flight.
// The A operand is either a dense [M,K] matrix (1x1 convs, linears, packed 2x2-s2 convs) or an
// implicit 3x3/pad-1 im2col over an NHWC tensor fetched with a 4-D tensor map whose out-of-bounds
// zero fill supplies the padding (the HCA fused 3x3, reference models/fpg/hca.py:21-23).
#pragma once

#include "common.cuh"

namespace hd {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;       // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
// warp 0 TMA, warp 1 MMA, warps 2.. epilogue: EW = 8 (two per TMEM lane quadrant) for multi-wave grids,
// EW = 16 for grids that fit one wave (one CTA per SM anyway: spend the idle issue slots on the epilogue)
constexpr int num_threads(int ew) { return 64 + 32 * ew; }

struct TcArgs {
  int M, N, num_kb;           // rows, packed weight rows, K / 64
  const float* bias;
  void* out;
  int ldo;
  const float* resid;
  int ldr;
  int sp;                     // A_CONV3: spatial n; EPI_PIXSHUF: spatial n of the GEMM rows
  int kb_per_tap;             // A_CONV3: C / 64
  int conv_bh, conv_bb;       // A_CONV3: box rows in h and in batch (conv_bh * sp * conv_bb == 128)
  DeviceStatus* status;
  const void* pf_ptr;         // weights of the NEXT tensor-core GEMM of the plan: each CTA prefetches its share into L2
  unsigned int pf_bytes;      // (0 = none) so that GEMM's weight stream starts from L2 instead of HBM
  unsigned long long w_policy; // L2 eviction priority of the weight tiles (kL2EvictNormal / kL2EvictFirst)
  long long* trace;           // optional per-CTA timeline (16 slots per CTA), nullptr in production
};

template <int BN, int STAGES, int EW = 8> struct TileCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STAGING_BYTES = BM * BN * 4;  // fp32 accumulator tile, aliases the ring after the mainloop
  static_assert(RING_BYTES >= STAGING_BYTES, "epilogue staging must fit in the operand ring");
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;               // power of two >= 32
  // co-resident CTAs per SM (228 KB shared memory, 512 TMEM columns): short-K GEMMs are dominated by
  // prologue/epilogue, so several small CTAs per SM overlap one tile's epilogue with another's mainloop
  // (320 threads per CTA: two CTAs keep the register budget at ~100 per thread; 576 threads: one CTA)
  static constexpr int MIN_BLOCKS = (EW > 8) ? 1 : (SMEM_BYTES <= 113 * 1024 ? 2 : 1);
  static_assert(MIN_BLOCKS * TMEM_COLS <= 512, "TMEM over-subscribed");
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait.  The fast path is a tight spin on mbarrier.try_wait (which itself suspends the thread in
// hardware for a bounded time); only every 4096 failed polls does the thread look at the SM clock and
// at the handle's error word.  A broken pipeline therefore trips a watchdog (~2 s) instead of hanging
// the GPU, and once any CTA has tripped every other wait gives up quickly.  The host reports
// HD_ERR_KERNEL.  (Polling %globaltimer / global memory on every spin costs ~1 us per poll and
// serialises the TMA->MMA pipeline: measured 3x slower GEMMs.)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, DeviceStatus* st, uint32_t site) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((it & 0xFFFu) == 0u) {
      if (*reinterpret_cast<volatile unsigned int*>(&st->error) != 0u) return false;
      if (clock64() - t0 > 4000000000ll) {
        if (atomicCAS(&st->error, 0u, 1u) == 0u) st->where = site;
        return false;
      }
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// the same with an L2 eviction-priority hint (policy = one of the kL2* encodings below)
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull, kL2EvictNormal = 0x1000000000000000ull;
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// L2 prefetch of `bytes` (multiple of 16) at a 16-byte-aligned global address
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(p)), "r"(bytes) : "memory");
}
// This CTA's share of the next GEMM's weight matrix, 2 KB per instruction, issued by the lanes of one warp.  The
// ~0.8 GB of weights are streamed from HBM once per denoise step whatever happens; fetching GEMM i+1's matrix while
// GEMM i runs turns the DRAM latency at the head of every weight stream into an L2 hit.
__device__ __forceinline__ void prefetch_next_weights(const void* ptr, uint32_t bytes, int lane) {
  if (bytes == 0u) return;
  const uint32_t ncta = gridDim.x * gridDim.y * gridDim.z;
  const uint32_t cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const uint32_t per = ((bytes / ncta) + 15u) & ~15u;
  const uint32_t begin = cta * per;
  if (begin >= bytes) return;
  const uint32_t end = begin + per < bytes ? begin + per : bytes;
  for (uint32_t off = begin + static_cast<uint32_t>(lane) * 2048u; off < end; off += 32u * 2048u) {
    const uint32_t n = end - off < 2048u ? end - off : 2048u;
    prefetch_l2_bulk(static_cast<const char*>(ptr) + off, n & ~15u);
  }
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], single-CTA, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint64_t desc_a, uint64_t desc_b, uint32_t tmem_d, uint32_t accumulate,
                                          uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 = 1024B (8 rows x 128B)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor: c=f32 (bit 4), a=b=bf16 (bits 7,10), K-major both, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// Epilogue.  Phase A: each epilogue warp drains its 32 TMEM lanes (= 32 output rows) into an fp32
// staging tile in shared memory (16-byte chunks XOR-swizzled by row: conflict-free both ways).
// Phase B: the same warp walks its rows with lanes along the columns, so every global access
// (bias, residual read, output write) is a contiguous, fully coalesced row segment.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_nctaid_z() {
  uint32_t v;
  asm volatile("mov.u32 %0, %%cluster_nctaid.z;" : "=r"(v));
  return v;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16-byte load from the shared memory of CTA `rank` of this cluster (DSMEM); rank == own rank is local
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(remote)
               : "memory");
  return v;
}
template <typename TOut> __device__ __forceinline__ void store4(TOut* dst, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* dst, float4 v) {
  *reinterpret_cast<float4*>(dst) = v;
}
template <> __device__ __forceinline__ void store4<bf16>(bf16* dst, float4 v) {
  uint2 p;
  p.x = pack_bf16x2(v.x, v.y);
  p.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(dst) = p;
}

// Split-K: gridDim.z CTAs forming one thread-block cluster (1,1,S) share an output tile; CTA z
// accumulates k-blocks [z*num_kb/S, (z+1)*num_kb/S) in its own TMEM, stages the partial tile in its
// own shared memory, and after a cluster barrier reduces rows [z*128/S, (z+1)*128/S) of all S
// partial tiles over distributed shared memory in fixed order (deterministic), applying the real
// epilogue once.  This turns the small-M, weight-streaming GEMMs of the 2x2 / 1x1 levels into
// >= 120 CTAs with deep TMA rings without atomics or a global workspace.
template <int BN, int STAGES, int EPI, int AMODE, typename TOut, int EW>
__global__ void __launch_bounds__(num_threads(EW), (TileCfg<BN, STAGES, EW>::MIN_BLOCKS))
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcArgs args) {
  using Cfg = TileCfg<BN, STAGES, EW>;
  constexpr int NUM_EPI_WARPS = EW;
  static_assert(EW == 8 || EW == 16, "epilogue warps: 2 or 4 per TMEM lane quadrant");
  static_assert(BN / (EW / 4) >= 32, "each epilogue warp drains at least one 32-column TMEM chunk");
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert(EPI != EPI_GATE || BN == 128, "gate epilogue needs 128-column packed groups");
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment (same offset in every CTA of the cluster)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + Cfg::RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x;
  const int n0 = blockIdx.y * BN;
  const int m0 = m_tile * BM;
  long long* trace = args.trace;
  if (trace != nullptr) {
    trace += 16 * ((static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
    if (threadIdx.x == 0) {
      unsigned long long gt;
      unsigned int smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      trace[0] = clock64(); trace[9] = static_cast<long long>(gt); trace[10] = smid;
    }
  }
  pdl_trigger();
  const int nsplit = static_cast<int>(cluster_nctaid_z());
  const int zrank = blockIdx.z;  // == rank in the (1,1,S) cluster
  // Epilogue operands fetched BEFORE the accumulator is complete (their L2 / DRAM latency then hides under the
  // mainloop instead of sitting in the epilogue of these latency-bound one-wave GEMMs): the bias chunks of every
  // variant, and — for the 16-epilogue-warp variants with a per-element second operand — the residual rows
  // (EPI_RESID) or the bf16 multiplicand (EPI_MUL) of the rows this warp will finish.
  constexpr int E_CH = (EPI == EPI_GATE) ? BN / 8 : BN / 4;
  constexpr int E_LPR = E_CH < 32 ? E_CH : 32;
  constexpr int E_CPL = E_CH / E_LPR;
  constexpr bool PRE = (EPI == EPI_RESID || EPI == EPI_MUL) && EW == 16 && BN == 128;
  float4 bias_pre[E_CPL], bias2_pre[E_CPL];
  float4 ext_pre[PRE ? 8 : 1];
  const int kb_count = args.num_kb / nsplit;
  const int kb_begin = zrank * kb_count;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&mapA);
    prefetch_tensormap(&mapB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    fence_barrier_init();
    fence_proxy_async_smem();
    // Weights do not depend on the preceding kernel: fill the ring's W halves right away — before the
    // CTA-wide setup barrier (TMEM allocation) and before the programmatic-dependency wait — so the
    // weight stream overlaps both this CTA's setup and the predecessor's tail.
    const int pre0 = kb_count < STAGES ? kb_count : STAGES;
    for (int i = 0; i < pre0; ++i) {
      const uint32_t fb = smem_u32(&full_bar[i]);
      mbar_expect_tx(fb, Cfg::STAGE_BYTES);
      tma_load_2d_hint(smem_u32(smem + i * Cfg::STAGE_BYTES) + Cfg::A_BYTES, &mapB, (kb_begin + i) * BK, n0, fb, args.w_policy);
    }
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  float* stage = reinterpret_cast<float*>(smem);
  if (trace != nullptr && threadIdx.x == 0) trace[1] = clock64();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      int conv_b0 = 0, conv_h0 = 0;
      if (AMODE == A_CONV3) {
        if (args.conv_bb > 1) {
          conv_b0 = m_tile * args.conv_bb;
        } else {
          const int tiles_per_face = args.sp / args.conv_bh;
          conv_b0 = m_tile / tiles_per_face;
          conv_h0 = (m_tile % tiles_per_face) * args.conv_bh;
        }
      }
      const int pre = kb_count < STAGES ? kb_count : STAGES;  // W halves already in flight (setup)
      pdl_wait();
      for (int i = 0; i < kb_count; ++i) {
        const int kb = kb_begin + i;
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
        if (i >= pre) {
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u, args.status, 0x100u);
          mbar_expect_tx(fb, Cfg::STAGE_BYTES);
        }
        if (AMODE == A_CONV3) {
          const int tap = kb / args.kb_per_tap;
          const int c0 = (kb - tap * args.kb_per_tap) * BK;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma_load_4d(sa, &mapA, c0, dx, conv_h0 + dy, conv_b0, fb);
        } else {
          tma_load_2d(sa, &mapA, kb * BK, m0, fb);
        }
        if (i >= pre) tma_load_2d_hint(sb, &mapB, kb * BK, n0, fb, args.w_policy);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = make_idesc(BM, BN);
      for (int i = 0; i < kb_count; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph, args.status, 0x200u);
        tc_fence_after_sync();
        if (trace != nullptr && i == 0) trace[2] = clock64();
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t da = make_smem_desc(sa);
        const uint64_t db = make_smem_desc(sa + Cfg::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128B swizzle row: +2 in the (addr >> 4) field
          umma_bf16(da + 2 * k, db + 2 * k, tmem_base, (i | k) != 0 ? 1u : 0u, idesc);
        }
        umma_commit(smem_u32(&empty_bar[s]));
      }
      umma_commit(smem_u32(tmem_full_bar));
      if (trace != nullptr) trace[3] = clock64();
    }
  } else {
    // ---------------- epilogue phase A (warps 2..5): TMEM -> swizzled fp32 staging tile ----------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    if (warp == 2) prefetch_next_weights(args.pf_ptr, args.pf_bytes, lane);
    {
      const int sl = lane % E_LPR;
#pragma unroll
      for (int i = 0; i < E_CPL; ++i) {
        const int ck = i * E_LPR + sl;
        bias_pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        bias2_pre[i] = bias_pre[i];
        if (EPI != EPI_PIXSHUF) bias_pre[i] = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + ck * 4));
        if (EPI == EPI_GATE) bias2_pre[i] = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + 64 + ck * 4));
      }
    }
    pdl_wait();                 // phase B reads the residual / overwrites buffers the predecessor may still use
    if (PRE) {
      // pass p of this warp finishes row (zrank * BM / nsplit) + p * 16 + (warp - 2) of the tile, lanes along the columns
      const int rows_here = BM / nsplit, row_base = zrank * rows_here, ew = warp - 2;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int rl = p * NUM_EPI_WARPS + ew;
        const int m = m0 + row_base + rl;
        ext_pre[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rl < rows_here && m < args.M) {
          if (EPI == EPI_RESID) {
            ext_pre[p] = *reinterpret_cast<const float4*>(args.resid + static_cast<size_t>(m) * args.ldr + n0 + lane * 4);
          } else {
            const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(args.resid) + static_cast<size_t>(m) * args.ldr + n0 + lane * 4);
            const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
            ext_pre[p] = make_float4(a.x, a.y, b.x, b.y);
          }
        }
      }
    }
    mbar_wait(smem_u32(tmem_full_bar), 0u, args.status, 0x300u);
    tc_fence_after_sync();
    if (trace != nullptr && threadIdx.x == 64) trace[4] = clock64();
    // All MMAs have completed: every TMA load has landed and been consumed, the ring is free.
    const int r = quad * 32 + lane;
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    float* srow = stage + r * BN;
    constexpr int COLS_PER_WARP = BN / (NUM_EPI_WARPS / 4);
    const int cbeg = ((warp - 2) >> 2) * COLS_PER_WARP;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + COLS_PER_WARP; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr_row + c0, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ck = (c0 >> 2) + j;
        *reinterpret_cast<uint4*>(srow + ((ck ^ (r & 7)) << 2)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }

  __syncwarp();
  if (trace != nullptr && threadIdx.x == 64) trace[5] = clock64();
  tc_fence_before_sync();
  if (nsplit > 1) cluster_sync_all(); else __syncthreads();
  if (trace != nullptr && threadIdx.x == 64) trace[6] = clock64();

  if (warp >= 2) {
    // ---------------- epilogue phase B: lanes along columns, coalesced global traffic ----------------
    constexpr int CH = (EPI == EPI_GATE) ? BN / 8 : BN / 4;  // output 4-column chunks per row
    constexpr int LPR = CH < 32 ? CH : 32;                  // lanes per row
    constexpr int RPI = 32 / LPR;                           // rows per warp pass
    constexpr int CPL = CH / LPR;                           // chunks per lane
    constexpr int U = 4;                                    // passes batched for memory-level parallelism
    const int ew = warp - 2;
    const int sub = lane / LPR, sl = lane % LPR;
    const int rows_here = BM / nsplit;                      // rows of the tile this CTA finishes
    const int row_base = zrank * rows_here;
    const uint32_t stage_u32 = smem_u32(stage);

    int out_col0 = n0;
    int q = 0;
    if (EPI == EPI_PIXSHUF) {
      const int quarter = args.N >> 2;
      q = n0 / quarter;  // q = 2*i + j of PixelShuffle(2)
      out_col0 = n0 - q * quarter;
    }
    if (EPI == EPI_GATE) out_col0 = n0 >> 1;

    static_assert(CPL == E_CPL && LPR == E_LPR, "bias preload uses the phase B lane mapping");
    float4 (&bias_r)[CPL] = bias_pre;
    float4 (&bias2_r)[CPL] = bias2_pre;

    const int passes = (rows_here + NUM_EPI_WARPS * RPI - 1) / (NUM_EPI_WARPS * RPI);

    // one row-chunk: bias / activation / residual, then the store
    auto finish = [&](float4 v, float4 g, float4 e, int i, TOut* d, bool okay) -> float4 {
      v.x += bias_r[i].x; v.y += bias_r[i].y; v.z += bias_r[i].z; v.w += bias_r[i].w;
      if (EPI == EPI_GATE) {
        v.x *= g.x + bias2_r[i].x; v.y *= g.y + bias2_r[i].y; v.z *= g.z + bias2_r[i].z; v.w *= g.w + bias2_r[i].w;
      }
      if (EPI == EPI_RELU) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      if (EPI == EPI_SIGMOID) {
        v.x = 1.f / (1.f + __expf(-v.x)); v.y = 1.f / (1.f + __expf(-v.y));
        v.z = 1.f / (1.f + __expf(-v.z)); v.w = 1.f / (1.f + __expf(-v.w));
      }
      if (EPI == EPI_RESID || EPI == EPI_PIXSHUF) { v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w; }
      if (EPI == EPI_MUL) { v.x *= e.x; v.y *= e.y; v.z *= e.z; v.w *= e.w; }
      if (okay) store4<TOut>(d + i * LPR * 4, v);
      return v;
    };
    // EPI_MUL: the bf16 multiplicand of one 4-column chunk
    auto load_mul = [&](const bf16* p) -> float4 {
      const uint2 u = *reinterpret_cast<const uint2*>(p);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
      return make_float4(a.x, a.y, b.x, b.y);
    };
    const bf16* mul_src = reinterpret_cast<const bf16*>(args.resid);
    // row bookkeeping shared by both paths
    auto locate = [&](int pass, bool& okay, int& rc, int& mc, TOut*& d) {
      const int rl = (pass * NUM_EPI_WARPS + ew) * RPI + sub;  // row within this CTA's slice
      const int r = row_base + rl;
      const int m = m0 + r;
      okay = pass < passes && rl < rows_here && m < args.M;
      rc = okay ? r : row_base;                            // clamp: loads stay in bounds
      mc = okay ? m : m0 + row_base;
      size_t out_row = static_cast<size_t>(mc);
      if (EPI == EPI_PIXSHUF) {
        const int sp = args.sp;
        const int face = mc / (sp * sp);
        const int rem = mc - face * sp * sp;
        const int hh = rem / sp, ww = rem - hh * sp;
        out_row = (static_cast<size_t>(face) * (2 * sp) + (2 * hh + (q >> 1))) * (2 * sp) + (2 * ww + (q & 1));
      }
      d = reinterpret_cast<TOut*>(args.out) + out_row * args.ldo + out_col0 + sl * 4;
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    if (nsplit == 1 && EPI != EPI_PIXSHUF) {
      // Fast path: the CTA finishes all 128 rows; a lane group's rows advance by STEP (a multiple of 8, so
      // the staging swizzle phase is pass-invariant) and every address is base + compile-time offset.
      constexpr int STEP = NUM_EPI_WARPS * RPI;
      constexpr int PASSES = BM / STEP;
      constexpr int U = PASSES < 4 ? PASSES : 4;            // passes batched for memory-level parallelism
      static_assert(STEP % 8 == 0 && PASSES % U == 0, "row walk");
      const int r0 = ew * RPI + sub;
      const int swz = r0 & 7;
      const float* sbase = stage + r0 * BN;
      TOut* dbase = reinterpret_cast<TOut*>(args.out) + static_cast<size_t>(m0 + r0) * args.ldo + out_col0 + sl * 4;
      const float* rbase = EPI == EPI_RESID ? args.resid + static_cast<size_t>(m0 + r0) * args.ldr + n0 + sl * 4 : nullptr;
      const bf16* mbase = EPI == EPI_MUL ? mul_src + static_cast<size_t>(m0 + r0) * args.ldr + n0 + sl * 4 : nullptr;
      const size_t dstep = static_cast<size_t>(STEP) * args.ldo, rstep = static_cast<size_t>(STEP) * args.ldr;
      int soff[CPL], soff2[CPL];
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        soff[i] = ((i * LPR + sl) ^ swz) << 2;
        soff2[i] = ((i * LPR + sl + 16) ^ swz) << 2;
      }
      if (PRE) {
        // second operand already in registers (ext_pre, fetched under the mainloop): PASSES == 8, CPL == 1
#pragma unroll
        for (int it0 = 0; it0 < PASSES; it0 += U) {
          float4 acc[U];
#pragma unroll
          for (int u = 0; u < U; ++u) acc[u] = *reinterpret_cast<const float4*>(sbase + (it0 + u) * (STEP * BN) + soff[0]);
#pragma unroll
          for (int u = 0; u < U; ++u)
            finish(acc[u], zero4, ext_pre[(it0 + u) & 7], 0, dbase + (it0 + u) * dstep, m0 + r0 + (it0 + u) * STEP < args.M);
        }
      } else
#pragma unroll 1
      for (int it0 = 0; it0 < PASSES; it0 += U) {
        float4 acc[U][CPL], acc2[U][CPL], ext[U][CPL];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int p = it0 + u;
          ok[u] = m0 + r0 + p * STEP < args.M;
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            acc[u][i] = *reinterpret_cast<const float4*>(sbase + p * (STEP * BN) + soff[i]);
            acc2[u][i] = zero4;
            ext[u][i] = zero4;
            if (EPI == EPI_GATE) acc2[u][i] = *reinterpret_cast<const float4*>(sbase + p * (STEP * BN) + soff2[i]);
            if (EPI == EPI_RESID && ok[u]) ext[u][i] = *reinterpret_cast<const float4*>(rbase + p * rstep + i * LPR * 4);
            if (EPI == EPI_MUL && ok[u]) ext[u][i] = load_mul(mbase + p * rstep + i * LPR * 4);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            finish(acc[u][i], acc2[u][i], ext[u][i], i, dbase + (it0 + u) * dstep, ok[u]);
          }
      }
    } else if (nsplit == 1) {
      constexpr int U = 4;                                  // passes batched for memory-level parallelism
#pragma unroll 1
      for (int it0 = 0; it0 < passes; it0 += U) {
        float4 acc[U][CPL], acc2[U][CPL], ext[U][CPL];
        bool ok[U];
        TOut* dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          int rc, mc;
          locate(it0 + u, ok[u], rc, mc, dst[u]);
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            const int ck = i * LPR + sl;
            acc[u][i] = *reinterpret_cast<const float4*>(stage + rc * BN + ((ck ^ (rc & 7)) << 2));
            acc2[u][i] = zero4;
            ext[u][i] = zero4;
            if (EPI == EPI_GATE) acc2[u][i] = *reinterpret_cast<const float4*>(stage + rc * BN + (((ck + 16) ^ (rc & 7)) << 2));
            if (EPI == EPI_RESID)
              ext[u][i] = *reinterpret_cast<const float4*>(args.resid + static_cast<size_t>(mc) * args.ldr + n0 + ck * 4);
            if (EPI == EPI_MUL) ext[u][i] = load_mul(mul_src + static_cast<size_t>(mc) * args.ldr + n0 + ck * 4);
            if (EPI == EPI_PIXSHUF) ext[u][i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dst[u]) + i * LPR * 4);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            finish(acc[u][i], acc2[u][i], ext[u][i], i, dst[u], ok[u]);
          }
      }
    } else {
      // split-K: all DSMEM loads of a pass are issued before the first add (they are ~200+ cycles each
      // and would otherwise serialise), then summed in fixed split order -> deterministic
      constexpr int MAXS = 8;
      constexpr int NP = (EPI == EPI_GATE) ? 2 : 1;
      auto do_pass = [&](int pass, float4 e_pre) {
        bool okay;
        int rc, mc;
        TOut* d;
        locate(pass, okay, rc, mc, d);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const int ck = i * LPR + sl;
          float4 part[NP][MAXS];
          float4 e = e_pre;
#pragma unroll
          for (int pi = 0; pi < NP; ++pi) {
            const uint32_t a = stage_u32 + static_cast<uint32_t>((rc * BN + (((ck + 16 * pi) ^ (rc & 7)) << 2)) * 4);
#pragma unroll
            for (int sidx = 0; sidx < MAXS; ++sidx)
              if (sidx < nsplit) part[pi][sidx] = ld_dsmem_f4(a, static_cast<uint32_t>(sidx));
          }
          if (!PRE) {
            if (EPI == EPI_RESID)
              e = *reinterpret_cast<const float4*>(args.resid + static_cast<size_t>(mc) * args.ldr + n0 + ck * 4);
            if (EPI == EPI_MUL) e = load_mul(mul_src + static_cast<size_t>(mc) * args.ldr + n0 + ck * 4);
          }
          if (EPI == EPI_PIXSHUF) e = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d) + i * LPR * 4);
          float4 v = zero4, g = zero4;
#pragma unroll
          for (int sidx = 0; sidx < MAXS; ++sidx) {
            if (sidx < nsplit) {
              v.x += part[0][sidx].x; v.y += part[0][sidx].y; v.z += part[0][sidx].z; v.w += part[0][sidx].w;
              if (EPI == EPI_GATE) {
                g.x += part[NP - 1][sidx].x; g.y += part[NP - 1][sidx].y; g.z += part[NP - 1][sidx].z; g.w += part[NP - 1][sidx].w;
              }
            }
          }
          finish(v, g, e, i, d, okay);
        }
      };
      if (PRE) {  // at most 128 / 2 / 16 = 4 passes; the second operand comes from ext_pre
#pragma unroll
        for (int pass = 0; pass < 4; ++pass)
          if (pass < passes) do_pass(pass, ext_pre[pass & 7]);
      } else {
#pragma unroll 1
        for (int pass = 0; pass < passes; ++pass) do_pass(pass, zero4);
      }
    }
  }

  if (trace != nullptr && threadIdx.x == 64) trace[7] = clock64();
  // no CTA may exit (or free TMEM) while a peer can still read its staging tile
  if (nsplit > 1) cluster_sync_all(); else __syncthreads();
  if (trace != nullptr && threadIdx.x == 0) trace[8] = clock64();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// ================================================================================================
// 2-CTA variant (tcgen05 cta_group::2): a thread-block cluster (2,1,1) of two SMs computes one 256 x 256
// output tile.  Each CTA TMA-loads its own 128 A rows and its half (128 rows) of the W tile; the leader
// CTA issues M=256, N=256, K=16 MMAs that read A and W from both CTAs' shared memory and write each CTA's
// 128 accumulator rows into that CTA's TMEM.  Operand fill per CTA and k-block stays 32 KB for twice the
// flops of the 128x128 tile — the fill (~60 B/clk/SM) is what bounds the 1-CTA mainloop.
//   - TMA loads of both CTAs complete on the LEADER's full barrier (peer bit of the barrier address
//     cleared); the leader's producer arms it with the bytes of both CTAs.
//   - tcgen05.commit.cta_group::2 multicasts slot-free / accumulator-ready arrivals to both CTAs.
//   - TMEM is allocated with cta_group::2 by the same warp in both CTAs.
// No split-K; epilogues: bias, ReLU, SimpleGate, +residual (rows walked coalesced, as in the 1-CTA kernel).
// ================================================================================================
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t v;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(v));
  return v;
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d_hint(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                             uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint64_t desc_a, uint64_t desc_b, uint32_t tmem_d, uint32_t accumulate,
                                           uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {  // arrive on the same barrier in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3))
               : "memory");
}

struct Tile2Cfg {
  static constexpr int BN = 256;                 // pair tile: 256 rows x 256 columns
  static constexpr int STAGES = 4;
  static constexpr int A_BYTES = BM * BK * 2;    // this CTA's 128 A rows
  static constexpr int B_BYTES = (BN / 2) * BK * 2;  // this CTA's half of the W tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;   // 128 KB == fp32 staging of 128 x 256
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES + 1024;
  static constexpr int EW = 8;
};

template <int EPI, int AMODE, typename TOut>
__global__ void __launch_bounds__(num_threads(Tile2Cfg::EW), 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcArgs args) {
  using Cfg = Tile2Cfg;
  constexpr int BN = Cfg::BN, STAGES = Cfg::STAGES, NUM_EPI_WARPS = Cfg::EW;
  static_assert(EPI == EPI_BIAS || EPI == EPI_RELU || EPI == EPI_GATE || EPI == EPI_RESID, "2-CTA epilogues");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + Cfg::RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  const int m_tile = blockIdx.x;                 // 128-row tile of this CTA (pair = blockIdx.x / 2)
  const int m0 = m_tile * BM;
  const int n0 = blockIdx.y * BN;
  const int num_kb = args.num_kb;
  pdl_trigger();

  if (threadIdx.x == 0) {
    prefetch_tensormap(&mapA);
    prefetch_tensormap(&mapB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);   // leader: its producer's arrive.expect_tx (bytes of both CTAs)
      mbar_init(smem_u32(&empty_bar[s]), 1);  // one multicast tcgen05.commit per use
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc2(smem_u32(tmem_slot), 256);
    tmem_relinquish2();
  }
  tc_fence_before_sync();
  cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  float* stage = reinterpret_cast<float*>(smem);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs) ----------------
      int conv_b0 = 0, conv_h0 = 0;
      if (AMODE == A_CONV3) {
        if (args.conv_bb > 1) {
          conv_b0 = m_tile * args.conv_bb;
        } else {
          const int tiles_per_face = args.sp / args.conv_bh;
          conv_b0 = m_tile / tiles_per_face;
          conv_h0 = (m_tile % tiles_per_face) * args.conv_bh;
        }
      }
      pdl_wait();
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u, args.status, 0x500u);
        const uint32_t fb_local = smem_u32(&full_bar[s]);
        const uint32_t fb_leader = fb_local & kPeerBitMask;
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
        if (rank == 0) mbar_expect_tx(fb_local, 2 * Cfg::STAGE_BYTES);
        if (AMODE == A_CONV3) {
          const int tap = i / args.kb_per_tap;
          const int c0 = (i - tap * args.kb_per_tap) * BK;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma2_load_4d(sa, &mapA, c0, dx, conv_h0 + dy, conv_b0, fb_leader);
        } else {
          tma2_load_2d(sa, &mapA, i * BK, m0, fb_leader);
        }
        tma2_load_2d_hint(sb, &mapB, i * BK, n0 + static_cast<int>(rank) * (BN / 2), fb_leader, args.w_policy);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ---------------- MMA issuer (leader CTA only) ----------------
      constexpr uint32_t idesc = make_idesc(2 * BM, BN);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph, args.status, 0x600u);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t da = make_smem_desc(sa);
        const uint64_t db = make_smem_desc(sa + Cfg::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) umma2_bf16(da + 2 * k, db + 2 * k, tmem_base, (i | k) != 0 ? 1u : 0u, idesc);
        umma2_commit_mc(smem_u32(&empty_bar[s]));
      }
      umma2_commit_mc(smem_u32(tmem_full_bar));
    }
  } else {
    // ---------------- epilogue phase A: this CTA's 128 accumulator rows -> staging ----------------
    const int quad = warp & 3;
    if (warp == 2) prefetch_next_weights(args.pf_ptr, args.pf_bytes, lane);
    pdl_wait();
    mbar_wait(smem_u32(tmem_full_bar), 0u, args.status, 0x700u);
    tc_fence_after_sync();
    const int r = quad * 32 + lane;
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    float* srow = stage + r * BN;
    constexpr int COLS_PER_WARP = BN / (NUM_EPI_WARPS / 4);
    const int cbeg = ((warp - 2) >> 2) * COLS_PER_WARP;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + COLS_PER_WARP; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr_row + c0, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ck = (c0 >> 2) + j;
        *reinterpret_cast<uint4*>(srow + ((ck ^ (r & 7)) << 2)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }
  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();

  if (warp >= 2) {
    // ---------------- epilogue phase B ----------------
    // gate: the 256 packed columns are two 128-column groups [64 x1 | 64 x2]; output chunk j (of 32) takes
    // staged chunks (j/16)*32 + j%16 and +16
    constexpr int CH = (EPI == EPI_GATE) ? BN / 8 : BN / 4;
    constexpr int LPR = 32, RPI = 1, CPL = CH / LPR;
    constexpr int STEP = NUM_EPI_WARPS * RPI, PASSES = BM / STEP, U = 4;
    const int ew = warp - 2, sl = lane;
    const int out_col0 = (EPI == EPI_GATE) ? (n0 >> 1) : n0;
    const int r0 = ew;
    const int swz = r0 & 7;
    const float* sbase = stage + r0 * BN;
    TOut* dbase = reinterpret_cast<TOut*>(args.out) + static_cast<size_t>(m0 + r0) * args.ldo + out_col0 + sl * 4;
    const float* rbase = (EPI == EPI_RESID) ? args.resid + static_cast<size_t>(m0 + r0) * args.ldr + n0 + sl * 4 : nullptr;
    const size_t dstep = static_cast<size_t>(STEP) * args.ldo, rstep = static_cast<size_t>(STEP) * args.ldr;
    float4 bias_r[CPL], bias2_r[CPL];
    int soff[CPL], soff2[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int j = i * LPR + sl;                                   // output chunk
      const int ck = (EPI == EPI_GATE) ? ((j >> 4) * 32 + (j & 15)) : j;  // staged chunk
      soff[i] = (ck ^ swz) << 2;
      soff2[i] = ((ck + 16) ^ swz) << 2;
      bias_r[i] = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + ck * 4));
      bias2_r[i] = bias_r[i];
      if (EPI == EPI_GATE) bias2_r[i] = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + (ck + 16) * 4));
    }
#pragma unroll 1
    for (int it0 = 0; it0 < PASSES; it0 += U) {
      float4 acc[U][CPL], acc2[U][CPL], ext[U][CPL];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = it0 + u;
        ok[u] = m0 + r0 + p * STEP < args.M;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          acc[u][i] = *reinterpret_cast<const float4*>(sbase + p * (STEP * BN) + soff[i]);
          acc2[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          ext[u][i] = acc2[u][i];
          if (EPI == EPI_GATE) acc2[u][i] = *reinterpret_cast<const float4*>(sbase + p * (STEP * BN) + soff2[i]);
          if (EPI == EPI_RESID && ok[u]) ext[u][i] = *reinterpret_cast<const float4*>(rbase + p * rstep + i * LPR * 4);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          float4 v = acc[u][i];
          v.x += bias_r[i].x; v.y += bias_r[i].y; v.z += bias_r[i].z; v.w += bias_r[i].w;
          if (EPI == EPI_GATE) {
            const float4 g = acc2[u][i];
            v.x *= g.x + bias2_r[i].x; v.y *= g.y + bias2_r[i].y; v.z *= g.z + bias2_r[i].z; v.w *= g.w + bias2_r[i].w;
          }
          if (EPI == EPI_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          if (EPI == EPI_RESID) { v.x += ext[u][i].x; v.y += ext[u][i].y; v.z += ext[u][i].z; v.w += ext[u][i].w; }
          if (ok[u]) store4<TOut>(dbase + (it0 + u) * dstep + i * LPR * 4, v);
        }
    }
  }
  // neither CTA may exit (its shared memory is an MMA operand of the pair) or free TMEM before both are done
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc2(tmem_base, 256);
  }
}

}  // namespace tc
}  // namespace hd
