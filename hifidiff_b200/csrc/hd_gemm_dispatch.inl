// hd_gemm_dispatch.inl — part of hd_lib.cu (one translation unit; included there in order, not compiled on its own).
// GEMM dispatch: tensor-map construction and the choice of tile / stage / split-K / cta_group::2 variant per shape.
// ------------------------------------------------------------------------------------------------
// GEMM dispatch
// ------------------------------------------------------------------------------------------------
template <typename TA, typename TW, typename TOut>
void launch_simt_typed(const GemmDesc& d, cudaStream_t st) {
  simt::SimtArgs g;
  g.M = d.M; g.N = d.N; g.K = d.K;
  g.A = d.A; g.lda = d.lda; g.a_mode = d.a_mode; g.sp = d.sp; g.C = d.C;
  g.W = d.W; g.ldw = d.ldw; g.bias = d.bias;
  g.out = d.out; g.ldo = d.ldo; g.resid = d.resid; g.ldr = d.ldr;
  dim3 grid(cdiv(d.M, simt::TM), cdiv(d.N, simt::TN));
  switch (d.epi) {
    case EPI_BIAS: launch_k(simt::gemm_simt_kernel<TA, TW, TOut, EPI_BIAS>, grid, 256, 0, st, g); break;
    case EPI_RELU: launch_k(simt::gemm_simt_kernel<TA, TW, TOut, EPI_RELU>, grid, 256, 0, st, g); break;
    case EPI_SIGMOID: launch_k(simt::gemm_simt_kernel<TA, TW, TOut, EPI_SIGMOID>, grid, 256, 0, st, g); break;
    case EPI_RESID: launch_k(simt::gemm_simt_kernel<TA, TW, TOut, EPI_RESID>, grid, 256, 0, st, g); break;
    case EPI_PIXSHUF: launch_k(simt::gemm_simt_kernel<TA, TW, TOut, EPI_PIXSHUF>, grid, 256, 0, st, g); break;
    default: break;
  }
}

void launch_simt(const GemmDesc& d, cudaStream_t st) {
  const bool abf = d.a_dtype == DT_BF16, wbf = d.w_dtype == DT_BF16, obf = d.out_dtype == DT_BF16;
  if (!abf && !wbf && !obf) launch_simt_typed<float, float, float>(d, st);
  else if (abf && wbf && obf) launch_simt_typed<bf16, bf16, bf16>(d, st);
  else if (abf && wbf && !obf) launch_simt_typed<bf16, bf16, float>(d, st);
  else if (!abf && !wbf && obf) launch_simt_typed<float, float, bf16>(d, st);
  else launch_simt_typed<float, float, float>(d, st);  // unreachable by construction (checked in add_gemm)
}

// split-precision mma.sync GEMM for the shallow CoarseRestoration stages (gemm_mma3.cuh)
template <int BN>
void launch_mma3_bn(const mma3::Args& a, int epi, cudaStream_t st) {
  const dim3 grid(static_cast<unsigned>((a.M + mma3::BM - 1) / mma3::BM), static_cast<unsigned>(a.N / BN));
  constexpr int smem = mma3::smem_bytes<BN>();
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(mma3::gemm_mma3_kernel<BN, EPI_BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_CHECK(cudaFuncSetAttribute(mma3::gemm_mma3_kernel<BN, EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_CHECK(cudaFuncSetAttribute(mma3::gemm_mma3_kernel<BN, EPI_PIXSHUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (epi == EPI_BIAS) launch_k(mma3::gemm_mma3_kernel<BN, EPI_BIAS>, grid, dim3(256), smem, st, a);
  else if (epi == EPI_RESID) launch_k(mma3::gemm_mma3_kernel<BN, EPI_RESID>, grid, dim3(256), smem, st, a);
  else launch_k(mma3::gemm_mma3_kernel<BN, EPI_PIXSHUF>, grid, dim3(256), smem, st, a);
}

void launch_mma3(const mma3::Args& a, int epi, cudaStream_t st) {
  if (a.N % 128 == 0) launch_mma3_bn<128>(a, epi, st);
  else if (a.N % 64 == 0) launch_mma3_bn<64>(a, epi, st);
  else launch_mma3_bn<32>(a, epi, st);
}

template <int BN, int K>
void launch_mma3h_bnk(const mma3::ArgsH& a, int epi, cudaStream_t st) {
  const dim3 grid(static_cast<unsigned>((a.M + mma3::BM - 1) / mma3::BM), static_cast<unsigned>(a.N / BN));
  constexpr int smem = mma3::smem_bytes_h<BN, K>();
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(mma3::gemm_mma3h_kernel<BN, K, EPI_BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_CHECK(cudaFuncSetAttribute(mma3::gemm_mma3h_kernel<BN, K, EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (epi == EPI_BIAS) launch_k(mma3::gemm_mma3h_kernel<BN, K, EPI_BIAS>, grid, dim3(256), smem, st, a);
  else launch_k(mma3::gemm_mma3h_kernel<BN, K, EPI_RESID>, grid, dim3(256), smem, st, a);
}
template <int K>
void launch_mma3h_k(const mma3::ArgsH& a, int epi, cudaStream_t st) {
  if (a.N % 128 == 0) launch_mma3h_bnk<128, K>(a, epi, st);
  else if (a.N % 64 == 0) launch_mma3h_bnk<64, K>(a, epi, st);
  else launch_mma3h_bnk<32, K>(a, epi, st);
}
void launch_mma3h(const mma3::ArgsH& a, int K, int epi, cudaStream_t st) {
  if (K == 32) launch_mma3h_k<32>(a, epi, st);
  else launch_mma3h_k<64>(a, epi, st);
}


struct TcLaunch {
  CUtensorMap mapA, mapB;
  tc::TcArgs args;
  dim3 grid;
  int epi, a_mode, out_dtype, stages, bn;
  bool two_cta;  // cta_group::2: CTA pairs on 256x256 tiles
};

template <int STAGES, int EW, int EPI, int AMODE, typename TOut, int BN = 128>
void launch_tc_inst2(const TcLaunch& L, cudaStream_t st) {
  auto kern = tc::gemm_tc_kernel<BN, STAGES, EPI, AMODE, TOut, EW>;
  using Cfg = tc::TileCfg<BN, STAGES, EW>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = L.grid;
  cfg.blockDim = dim3(tc::num_threads(EW));
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = L.grid.z;  // split-K CTAs of one tile form a cluster
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = t_use_pdl ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kern, L.mapA, L.mapB, L.args);
}

// ring depth: <= 2 k-blocks per CTA needs two stages; a grid that fits in one wave gets the deep 6-stage
// ring (weight streaming, one CTA per SM) and 16 epilogue warps; everything else 3 stages (2 CTAs/SM)
template <int EPI, int AMODE, typename TOut>
void launch_tc_inst(const TcLaunch& L, cudaStream_t st) {
  if (L.stages == 2) launch_tc_inst2<2, 8, EPI, AMODE, TOut>(L, st);
  else if (L.stages == 6) launch_tc_inst2<6, 16, EPI, AMODE, TOut>(L, st);
  else launch_tc_inst2<3, 8, EPI, AMODE, TOut>(L, st);
}

template <int EPI, int AMODE, typename TOut>
void launch_tc2_inst(const TcLaunch& L, cudaStream_t st) {
  auto kern = tc::gemm_tc2_kernel<EPI, AMODE, TOut>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Tile2Cfg::SMEM_BYTES);
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = L.grid;
  cfg.blockDim = dim3(tc::num_threads(tc::Tile2Cfg::EW));
  cfg.dynamicSmemBytes = tc::Tile2Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;  // the CTA pair
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = t_use_pdl ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kern, L.mapA, L.mapB, L.args);
}

void launch_tc2(const TcLaunch& L, cudaStream_t st) {
  const bool obf = L.out_dtype == DT_BF16;
  if (L.a_mode == A_CONV3) {
    launch_tc2_inst<EPI_RELU, A_CONV3, bf16>(L, st);
    return;
  }
  switch (L.epi) {
    case EPI_BIAS:
      if (obf) launch_tc2_inst<EPI_BIAS, A_PLAIN, bf16>(L, st);
      else launch_tc2_inst<EPI_BIAS, A_PLAIN, float>(L, st);
      break;
    case EPI_RELU: launch_tc2_inst<EPI_RELU, A_PLAIN, bf16>(L, st); break;
    case EPI_GATE: launch_tc2_inst<EPI_GATE, A_PLAIN, bf16>(L, st); break;
    case EPI_RESID: launch_tc2_inst<EPI_RESID, A_PLAIN, float>(L, st); break;
    default: break;
  }
}

void launch_tc(const TcLaunch& L, cudaStream_t st) {
  if (L.two_cta) { launch_tc2(L, st); return; }
  const bool obf = L.out_dtype == DT_BF16;
  if (L.a_mode == A_CONV3) {
    if (L.epi == EPI_RELU && obf) {
      // dense 3x3 (K = 9C, operand-fill-bound): 128x256 tiles halve the A fill per flop
      if (L.bn == 256) launch_tc_inst2<3, 8, EPI_RELU, A_CONV3, bf16, 256>(L, st);
      else launch_tc_inst<EPI_RELU, A_CONV3, bf16>(L, st);
    }
    return;
  }
  switch (L.epi) {
    case EPI_BIAS:
      if (obf) launch_tc_inst<EPI_BIAS, A_PLAIN, bf16>(L, st);
      else launch_tc_inst<EPI_BIAS, A_PLAIN, float>(L, st);
      break;
    case EPI_RELU: launch_tc_inst<EPI_RELU, A_PLAIN, bf16>(L, st); break;
    case EPI_RESID: launch_tc_inst<EPI_RESID, A_PLAIN, float>(L, st); break;
    case EPI_GATE: launch_tc_inst<EPI_GATE, A_PLAIN, bf16>(L, st); break;
    case EPI_PIXSHUF: launch_tc_inst<EPI_PIXSHUF, A_PLAIN, float>(L, st); break;
    case EPI_MUL: launch_tc_inst<EPI_MUL, A_PLAIN, bf16>(L, st); break;
    default: break;
  }
}

void encode_map(hd_handle* h, CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = h->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) HD_THROW(HD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
}

// rows_alloc: number of valid rows in the A allocation (>= M, multiple of 128)
TcLaunch build_tc(hd_handle* h, const GemmDesc& d, long long a_rows_alloc) {
  TcLaunch L;
  memset(&L, 0, sizeof(L));
  L.epi = d.epi; L.a_mode = d.a_mode; L.out_dtype = d.out_dtype;
  tc::TcArgs& a = L.args;
  a.M = d.M; a.N = d.N; a.num_kb = d.K / tc::BK;
  a.bias = d.bias; a.out = d.out; a.ldo = d.ldo; a.resid = d.resid; a.ldr = d.ldr;
  a.sp = d.sp; a.kb_per_tap = 1; a.conv_bh = 1; a.conv_bb = 1;
  a.status = h->d_status;
  a.trace = nullptr;
  a.w_policy = h->tun.w_evict_first ? tc::kL2EvictFirst : tc::kL2EvictNormal;
  if (d.a_mode == A_CONV3) {
    const int n = d.sp, C = d.C;
    if (128 % n != 0 || (n * n < 128 && 128 % (n * n) != 0)) HD_THROW(HD_ERR_UNSUPPORTED, "conv tile: spatial %d", n);
    int bh, bb;
    if (n * n >= 128) { bh = 128 / n; bb = 1; } else { bh = n; bb = 128 / (n * n); }
    a.kb_per_tap = C / tc::BK; a.conv_bh = bh; a.conv_bb = bb;
    const long long faces_alloc = a_rows_alloc / (n * n);
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)faces_alloc};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)n * C * 2, (cuuint64_t)n * n * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)n, (cuuint32_t)bh, (cuuint32_t)bb};
    encode_map(h, &L.mapA, d.A, 4, dims, strides, box);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)d.K, (cuuint64_t)a_rows_alloc};
    cuuint64_t strides[1] = {(cuuint64_t)d.lda * 2};
    cuuint32_t box[2] = {64, 128};
    encode_map(h, &L.mapA, d.A, 2, dims, strides, box);
  }
  // cta_group::2 pairs on 256x256 tiles: only where the mainloop dominates (K >= 1152: +5..15 % measured, -8..-25 %
  // on short-K shapes where one CTA per SM loses the inter-CTA overlap), the pair grid still covers the chip, and
  // the epilogue kind is supported
  const int m_tiles = cdiv(d.M, 128);
  const bool epi2 = d.epi == EPI_BIAS || d.epi == EPI_GATE || d.epi == EPI_RESID ||
                    (d.epi == EPI_RELU && d.out_dtype == DT_BF16);
  const bool conv_ok = d.a_mode != A_CONV3 || (d.epi == EPI_RELU && d.out_dtype == DT_BF16);
  const long long pair_ctas = static_cast<long long>((m_tiles + 1) / 2) * 2 * (d.N / 256);
  L.two_cta = h->tun.two_cta != 0 && epi2 && conv_ok && d.N % 256 == 0 && a_rows_alloc % 256 == 0 &&
              (h->tun.two_cta == 2 || (a.num_kb >= 18 && pair_ctas >= 128));  // measured: wins from K >= 1152 (tools/gemm_bench.py)
  if (L.two_cta) {
    L.bn = 256;
    cuuint64_t dims[2] = {(cuuint64_t)d.K, (cuuint64_t)d.N};
    cuuint64_t strides[1] = {(cuuint64_t)d.ldw * 2};
    cuuint32_t box[2] = {64, 128};  // each CTA of the pair loads its half of the 256-row W tile
    encode_map(h, &L.mapB, d.W, 2, dims, strides, box);
    L.grid = dim3(((m_tiles + 1) / 2) * 2, d.N / 256, 1);
    L.stages = 4;
    return L;
  }
  const int bn = (h->tun.bn256 && d.a_mode == A_CONV3 && d.epi == EPI_RELU && d.out_dtype == DT_BF16 && d.N % 256 == 0) ? 256 : 128;
  L.bn = bn;
  {
    cuuint64_t dims[2] = {(cuuint64_t)d.K, (cuuint64_t)d.N};
    cuuint64_t strides[1] = {(cuuint64_t)d.ldw * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    encode_map(h, &L.mapB, d.W, 2, dims, strides, box);
  }
  // split-K over a (1,1,S) cluster until the grid can cover the chip (>= 120 CTAs)
  const int tiles = cdiv(d.M, 128) * (d.N / bn);
  int split = 1;
  while (tiles * split < d.cta_target && split < h->tun.max_split && a.num_kb % (2 * split) == 0 && a.num_kb / (2 * split) >= 2) split *= 2;
  L.grid = dim3(cdiv(d.M, 128), d.N / bn, split);
  const int local_kb = a.num_kb / split;
  L.stages = local_kb <= 2 ? 2 : (tiles * split <= 160 ? 6 : 3);
  return L;
}

bool tc_eligible(const hd_handle* h, const GemmDesc& d) {
  if (!h->bf16) return false;
  if (d.a_dtype != DT_BF16 || d.w_dtype != DT_BF16) return false;
  if (d.K % 64 != 0 || d.N % 128 != 0) return false;
  if (d.a_mode == A_CONV3 && (d.C % 64 != 0)) return false;
  if (d.epi == EPI_SIGMOID) return false;
  return true;
}

