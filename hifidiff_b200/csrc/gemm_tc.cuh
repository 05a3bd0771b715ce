// tcgen05 / TMEM / TMA GEMM for sm_100a:   out = epilogue(A[M,K] * W[N,K]^T), bf16 in, fp32 accumulate.
//
// One CTA computes a 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      TMA producer: A tile (128 x 64 bf16) and W tile (BN x 64 bf16) per k-block into a
//               multi-stage shared-memory ring, 128B-swizzled, completion on an mbarrier.
//   warp 1      allocates TMEM, issues tcgen05.mma (M=128, N=BN, K=16, kind::f16) from one thread,
//               frees ring slots with tcgen05.commit.
//   warps 2-5   epilogue: tcgen05.ld the fp32 accumulator (one TMEM lane = one output row) and
//               apply bias / ReLU / residual add / SimpleGate / PixelShuffle scatter.
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
constexpr int NUM_THREADS = 192;

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
};

template <int BN> struct TileCfg {
  static constexpr int STAGES = (BN >= 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;                         // power of two >= 32
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
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a broken pipeline trips a watchdog (2 s) instead of hanging the GPU; once any CTA
// has tripped, every other wait gives up immediately.  The host reports HD_ERR_KERNEL.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, DeviceStatus* st, uint32_t site) {
  if (mbar_try_wait(bar, parity)) return true;
  const uint64_t t0 = global_timer_ns();
  while (true) {
    if (mbar_try_wait(bar, parity)) return true;
    if (*reinterpret_cast<volatile unsigned int*>(&st->error) != 0u) return false;
    if (global_timer_ns() - t0 > 2000000000ull) {
      if (atomicCAS(&st->error, 0u, 1u) == 0u) st->where = site;
      return false;
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
// epilogue helpers: `v` holds 32 consecutive accumulator columns of one output row
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void add_bias32(float (&v)[32], const float* __restrict__ bias) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float4 b = __ldg(reinterpret_cast<const float4*>(bias + j));
    v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
  }
}
__device__ __forceinline__ void store_row32(float* dst, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}
__device__ __forceinline__ void store_row32(bf16* dst, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 p;
    p.x = pack_bf16x2(v[j], v[j + 1]);
    p.y = pack_bf16x2(v[j + 2], v[j + 3]);
    p.z = pack_bf16x2(v[j + 4], v[j + 5]);
    p.w = pack_bf16x2(v[j + 6], v[j + 7]);
    *reinterpret_cast<uint4*>(dst + j) = p;
  }
}

template <int BN, int EPI, int AMODE, typename TOut>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcArgs args) {
  using Cfg = TileCfg<BN>;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert(EPI != EPI_GATE || BN == 128, "gate epilogue needs 128-column packed groups");
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x;
  const int n0 = blockIdx.y * BN;
  const int m0 = m_tile * BM;
  const int num_kb = args.num_kb;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&mapA);
    prefetch_tensormap(&mapB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

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
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u, args.status, 0x100u);
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
        mbar_expect_tx(fb, Cfg::STAGE_BYTES);
        if (AMODE == A_CONV3) {
          const int tap = kb / args.kb_per_tap;
          const int c0 = (kb - tap * args.kb_per_tap) * BK;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma_load_4d(sa, &mapA, c0, dx, conv_h0 + dy, conv_b0, fb);
        } else {
          tma_load_2d(sa, &mapA, kb * BK, m0, fb);
        }
        tma_load_2d(sb, &mapB, kb * BK, n0, fb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = make_idesc(BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph, args.status, 0x200u);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t da = make_smem_desc(sa);
        const uint64_t db = make_smem_desc(sa + Cfg::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128B swizzle row: +2 in the (addr >> 4) field
          umma_bf16(da + 2 * k, db + 2 * k, tmem_base, (kb | k) != 0 ? 1u : 0u, idesc);
        }
        umma_commit(smem_u32(&empty_bar[s]));
      }
      umma_commit(smem_u32(tmem_full_bar));
    }
  } else {
    // ---------------- epilogue (warps 2..5) ----------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int row_in_tile = quad * 32 + lane;
    const int m = m0 + row_in_tile;
    mbar_wait(smem_u32(tmem_full_bar), 0u, args.status, 0x300u);
    tc_fence_after_sync();
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const bool row_ok = m < args.M;

    if (EPI == EPI_GATE) {
      // packed 128-column groups: columns [0,64) are x1 channels, [64,128) the matching x2 channels
      TOut* orow = reinterpret_cast<TOut*>(args.out) + static_cast<size_t>(m) * args.ldo + (n0 >> 1);
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r1[32], r2[32];
        tmem_ld32(taddr_row + c0, r1);
        tmem_ld32(taddr_row + 64 + c0, r2);
        tmem_wait_ld();
        float v1[32], v2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { v1[j] = __uint_as_float(r1[j]); v2[j] = __uint_as_float(r2[j]); }
        add_bias32(v1, args.bias + n0 + c0);
        add_bias32(v2, args.bias + n0 + 64 + c0);
#pragma unroll
        for (int j = 0; j < 32; ++j) v1[j] *= v2[j];
        if (row_ok) store_row32(orow + c0, v1);
      }
    } else {
      size_t out_row = static_cast<size_t>(m);
      int out_col0 = n0;
      if (EPI == EPI_PIXSHUF) {
        const int quarter = args.N >> 2;
        const int q = n0 / quarter;  // q = 2*i + j of PixelShuffle(2)
        out_col0 = n0 - q * quarter;
        const int sp = args.sp;
        const int face = m / (sp * sp);
        const int rem = m - face * sp * sp;
        const int h = rem / sp, w = rem - h * sp;
        out_row = (static_cast<size_t>(face) * (2 * sp) + (2 * h + (q >> 1))) * (2 * sp) + (2 * w + (q & 1));
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr_row + c0, r);
        tmem_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (EPI != EPI_PIXSHUF) add_bias32(v, args.bias + n0 + c0);
        if (EPI == EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (EPI == EPI_SIGMOID) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
        }
        if (row_ok) {
          if (EPI == EPI_RESID) {
            const float* rrow = args.resid + static_cast<size_t>(m) * args.ldr + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 x = *reinterpret_cast<const float4*>(rrow + j);
              v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
            }
          }
          if (EPI == EPI_PIXSHUF) {
            float* orow = reinterpret_cast<float*>(args.out) + out_row * args.ldo + out_col0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 x = *reinterpret_cast<const float4*>(orow + j);
              v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
            }
            store_row32(orow, v);
          } else {
            TOut* orow = reinterpret_cast<TOut*>(args.out) + out_row * args.ldo + out_col0 + c0;
            store_row32(orow, v);
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace hd
