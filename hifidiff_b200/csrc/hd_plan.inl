// hd_plan.inl — part of hd_lib.cu (one translation unit; included there in order, not compiled on its own).
// Plan construction for one denoise step (the list of launches replayed per timestep).
// ------------------------------------------------------------------------------------------------
// plan construction
// ------------------------------------------------------------------------------------------------
void add_op(Plan& P, std::function<void(cudaStream_t)> fn, const std::string& tap = std::string(),
            TapInfo info = TapInfo()) {
  Op op;
  op.fn = std::move(fn);
  op.tap = tap;
  op.info = info;
  op.label = g_label;
  P.ops.push_back(std::move(op));
}

void add_gemm(hd_handle* h, Plan& P, GemmDesc d, long long a_rows_alloc, const std::string& tap = std::string(),
              TapInfo info = TapInfo()) {
  const long long taps_exec = d.a_mode == A_CONV3 ? 1 : 1;
  (void)taps_exec;
  P.flops_per_face += 2.0 * d.M * static_cast<double>(d.N) * d.K / P.batch;
  if (d.cta_target == 120) d.cta_target = h->tun.cta_target;
  static const char* epi_names[] = {"bias", "relu", "sigmoid", "resid", "gate", "pixshuf", "bias*mul"};
  const std::string what = g_label;
  if (tc_eligible(h, d)) {
    std::shared_ptr<TcLaunch> Lp = std::make_shared<TcLaunch>(build_tc(h, d, a_rows_alloc));
    TcLaunch& L = *Lp;
    if (h->tun.w_prefetch && d.ldw == d.K && (d.a_mode != A_CONV3)) {
      // this GEMM's weights are what the previous GEMM of the plan prefetches into L2 (the first one of the step is
      // prefetched by the last: the same plan runs again for the next timestep)
      const unsigned int wbytes = static_cast<unsigned int>(static_cast<size_t>(d.N) * d.K * 2);
      if (P.last_tc) { P.last_tc->args.pf_ptr = d.W; P.last_tc->args.pf_bytes = wbytes; }
      if (P.first_w == nullptr) { P.first_w = d.W; P.first_w_bytes = wbytes; }
    }
    if (h->tun.w_prefetch) P.last_tc = Lp;
    g_label = fmt("%s gemm_tc %s%s M=%d N=%d K=%d grid=(%d,%d,%d) stages=%d", what.c_str(), epi_names[d.epi],
                  L.two_cta ? (d.a_mode == A_CONV3 ? "+conv3 2CTA" : " 2CTA") : d.a_mode == A_CONV3 ? (L.bn == 256 ? "+conv3 BN=256" : "+conv3") : "",
                  d.M, d.N, d.K, L.grid.x, L.grid.y, L.grid.z, L.stages);
    add_op(P, [Lp](cudaStream_t st) { launch_tc(*Lp, st); }, tap, info);
    return;
  }
  if (d.epi == EPI_MUL) HD_THROW(HD_ERR_INVALID, "EPI_MUL exists on the tcgen05 path only (M=%d N=%d K=%d)", d.M, d.N, d.K);
  if (d.epi == EPI_GATE) {
    // fp32 mode: bias epilogue into a packed fp32 buffer, then the SimpleGate kernel
    GemmDesc g = d;
    g.epi = EPI_BIAS;
    g.out = h->gate_tmp;
    g.ldo = d.N;
    g.out_dtype = DT_F32;
    const int c = d.N / 2;
    void* out = d.out;
    const int odt = d.out_dtype;
    const size_t rows = d.M;
    float* tmp = h->gate_tmp;
    g_label = fmt("%s gemm_ffma bias M=%d N=%d K=%d", what.c_str(), d.M, d.N, d.K);
    add_op(P, [g](cudaStream_t st) { launch_simt(g, st); });
    g_label = what + " gate_packed";
    add_op(P, [=](cudaStream_t st) {
      const int blocks = cdiv(rows * c, 256);
      if (odt == DT_BF16) launch_k(gate_packed_kernel<bf16>, dim3(blocks), dim3(256), 0, st, tmp, static_cast<bf16*>(out), rows, c);
      else launch_k(gate_packed_kernel<float>, dim3(blocks), dim3(256), 0, st, tmp, static_cast<float*>(out), rows, c);
    }, tap, info);
    return;
  }
  const bool abf = d.a_dtype == DT_BF16, wbf = d.w_dtype == DT_BF16;
  if (abf != wbf) HD_THROW(HD_ERR_INVALID, "mixed-precision operands reached the FFMA GEMM");
  g_label = fmt("%s gemm_ffma %s M=%d N=%d K=%d", what.c_str(), epi_names[d.epi], d.M, d.N, d.K);
  add_op(P, [d](cudaStream_t st) { launch_simt(d, st); }, tap, info);
}

template <typename T>
void launch_ln(int c, const float* x, const float* lw, const float* lb, T* out, int rows, int rpf, ModRef mod,
               int shift_off, int scale_off, int has_mod, cudaStream_t st) {
  const int lpr = std::min(32, c / 16);
  const int grid = cdiv(rows, 4 * (32 / lpr));  // 4 warps per block, 32/lpr rows per warp
  switch (c) {
    case 32: launch_k(ln_mod_kernel<32, T>, dim3(grid), dim3(128), 0, st, x, lw, lb, out, rows, rpf, mod, shift_off, scale_off, has_mod); break;
    case 64: launch_k(ln_mod_kernel<64, T>, dim3(grid), dim3(128), 0, st, x, lw, lb, out, rows, rpf, mod, shift_off, scale_off, has_mod); break;
    case 128: launch_k(ln_mod_kernel<128, T>, dim3(grid), dim3(128), 0, st, x, lw, lb, out, rows, rpf, mod, shift_off, scale_off, has_mod); break;
    case 256: launch_k(ln_mod_kernel<256, T>, dim3(grid), dim3(128), 0, st, x, lw, lb, out, rows, rpf, mod, shift_off, scale_off, has_mod); break;
    case 512: launch_k(ln_mod_kernel<512, T>, dim3(grid), dim3(128), 0, st, x, lw, lb, out, rows, rpf, mod, shift_off, scale_off, has_mod); break;
    case 1024: launch_k(ln_mod_wide_kernel<1024, T>, dim3(rows), dim3(256), 0, st, x, lw, lb, out, rows, rpf, mod, shift_off, scale_off, has_mod); break;
    case 2048: launch_k(ln_mod_wide_kernel<2048, T>, dim3(rows), dim3(256), 0, st, x, lw, lb, out, rows, rpf, mod, shift_off, scale_off, has_mod); break;
    default: break;
  }
}

// LayerNorm straight into the [hi | lo | hi] bf16 operand of the split-precision GEMM (CoarseRestoration, c >= 128)
void launch_ln_split3(int c, const float* x, const float* lw, const float* lb, bf16* out3, int rows, int rpf, cudaStream_t st) {
  const int lpr = std::min(32, c / 16);
  const int grid = cdiv(rows, 4 * (32 / lpr));
  ModRef nomod{nullptr, nullptr, 0};
  switch (c) {
    case 128: launch_k(ln_mod_kernel<128, bf16, true>, dim3(grid), dim3(128), 0, st, x, lw, lb, out3, rows, rpf, nomod, 0, 0, 0); break;
    case 256: launch_k(ln_mod_kernel<256, bf16, true>, dim3(grid), dim3(128), 0, st, x, lw, lb, out3, rows, rpf, nomod, 0, 0, 0); break;
    case 512: launch_k(ln_mod_kernel<512, bf16, true>, dim3(grid), dim3(128), 0, st, x, lw, lb, out3, rows, rpf, nomod, 0, 0, 0); break;
    default: HD_THROW(HD_ERR_UNSUPPORTED, "split LayerNorm for c = %d", c);
  }
}

// rows a GEMM A operand in the shared workspace may claim (a multiple of 128 >= the rows actually used)
long long rows_cap(const hd_handle* h, int rpf) { return static_cast<long long>(h->Bcap) * rpf; }

// One ConditionalNAFBlock / NAFBlock as one kernel per op (conditional_naf.py:108-136): the plan of the 4x4, 2x2 and
// 1x1 levels, of the FPG encoder, and of every level in the debug (per-layer tap) plan.
void add_block(hd_handle* h, Plan& P, const BlockW& bw, const std::string& tapname) {
  const int B = P.batch, l = bw.level, c = bw.c, sp = h->sp[l];
  const int rows = B * sp * sp, rpf = sp * sp;
  const long long rows_alloc = rows_cap(h, rpf);
  const int adt = h->bf16 ? DT_BF16 : DT_F32;
  float* resid = h->resid[l];
  ModRef mod{h->mod_table, h->row_idx, h->mod_stride};
  const bool bf = h->bf16;
  void *act_a = h->act_a, *act_h = h->act_h, *act_g = h->act_g, *pooled = h->pooled;
  float* sca_s = h->sca_s;
  const int has_mod = bw.has_mod ? 1 : 0;
  auto ln = [=](const float* lw, const float* lb, int shift_off, int scale_off) {
    return [=](cudaStream_t st) {
      if (bf) launch_ln<bf16>(c, resid, lw, lb, static_cast<bf16*>(act_a), rows, rpf, mod, shift_off, scale_off, has_mod, st);
      else launch_ln<float>(c, resid, lw, lb, static_cast<float*>(act_a), rows, rpf, mod, shift_off, scale_off, has_mod, st);
    };
  };
  const std::string L0 = fmt("L%d c=%d ", l, c);
  // norm1 + modulation (shift_att = chunk 0, scale_att = chunk 1)
  g_label = L0 + "ln1";
  add_op(P, ln(bw.ln1_w, bw.ln1_b, bw.mod_off, bw.mod_off + c));
  g_label = L0 + "conv1";
  if (bw.dw_folded) {  // conv1 + (folded) depthwise + SimpleGate; the pooled mean over 1 pixel is g itself
    GemmDesc d;
    d.M = rows; d.N = 2 * c; d.K = c; d.A = act_a; d.lda = c; d.a_dtype = adt;
    d.W = bw.w1; d.ldw = c; d.w_dtype = adt; d.bias = bw.b1; d.epi = EPI_GATE;
    d.out = act_g; d.ldo = c; d.out_dtype = adt;
    add_gemm(h, P, d, rows_alloc);
  } else {  // conv1
    GemmDesc d;
    d.M = rows; d.N = 2 * c; d.K = c; d.A = act_a; d.lda = c; d.a_dtype = adt;
    d.W = bw.w1; d.ldw = c; d.w_dtype = adt; d.bias = bw.b1; d.epi = EPI_BIAS;
    d.out = act_h; d.ldo = 2 * c; d.out_dtype = adt;
    add_gemm(h, P, d, rows_alloc);
  }
  g_label = L0 + "dwconv_gate_pool";
  if (!bw.dw_folded) {  // depthwise 3x3 + SimpleGate + pool
    const float *dw_w = bw.dw_w, *dw_b = bw.dw_b;
    if (h->tun.dw_small && (sp == 2 || sp == 4) && c % 4 == 0) {  // register-resident faces (dwconv_small_kernel)
      add_op(P, [=](cudaStream_t st) {
        const dim3 grid(cdiv(static_cast<size_t>(B) * (c / 4), 256));
        if (bf && sp == 2) launch_k(dwconv_small_kernel<bf16, 2>, grid, dim3(256), 0, st, static_cast<const bf16*>(act_h), dw_w, dw_b, static_cast<bf16*>(act_g), static_cast<bf16*>(pooled), c, B);
        else if (bf) launch_k(dwconv_small_kernel<bf16, 4>, grid, dim3(256), 0, st, static_cast<const bf16*>(act_h), dw_w, dw_b, static_cast<bf16*>(act_g), static_cast<bf16*>(pooled), c, B);
        else if (sp == 2) launch_k(dwconv_small_kernel<float, 2>, grid, dim3(256), 0, st, static_cast<const float*>(act_h), dw_w, dw_b, static_cast<float*>(act_g), static_cast<float*>(pooled), c, B);
        else launch_k(dwconv_small_kernel<float, 4>, grid, dim3(256), 0, st, static_cast<const float*>(act_h), dw_w, dw_b, static_cast<float*>(act_g), static_cast<float*>(pooled), c, B);
      });
    } else if (sp > 16) {  // faces too large to stage whole (latent 32): taps from global memory, pool as its own kernel
      add_op(P, [=](cudaStream_t st) {
        const size_t total = static_cast<size_t>(rows) * c / 2;
        if (bf) launch_k(dwconv_gate_any_kernel<bf16>, dim3(cdiv(total, 256)), dim3(256), 0, st, static_cast<const bf16*>(act_h), dw_w, dw_b, static_cast<bf16*>(act_g), sp, c, total);
        else launch_k(dwconv_gate_any_kernel<float>, dim3(cdiv(total, 256)), dim3(256), 0, st, static_cast<const float*>(act_h), dw_w, dw_b, static_cast<float*>(act_g), sp, c, total);
      });
      g_label = L0 + "pool_faces";
      add_op(P, [=](cudaStream_t st) {
        const size_t total = static_cast<size_t>(B) * c;
        if (bf) launch_k(pool_faces_kernel<bf16>, dim3(cdiv(total, 256)), dim3(256), 0, st, static_cast<const bf16*>(act_g), static_cast<bf16*>(pooled), rpf, c, B);
        else launch_k(pool_faces_kernel<float>, dim3(cdiv(total, 256)), dim3(256), 0, st, static_cast<const float*>(act_g), static_cast<float*>(pooled), rpf, c, B);
      });
    } else
    add_op(P, [=](cudaStream_t st) {
      const int tile_px = sp >= 16 ? sp * sp : 64;   // whole faces per tile; small tiles below 16x16 for parallelism
      dim3 grid(c / 64, cdiv(rows, tile_px));
      if (bf) launch_k(dwconv_gate_pool_kernel<bf16>, dim3(grid), dim3(256), tile_px * 128 * 2, st, static_cast<const bf16*>(act_h), dw_w, dw_b,
                                                                  static_cast<bf16*>(act_g), static_cast<bf16*>(pooled), sp, c, rows, tile_px);
      else launch_k(dwconv_gate_pool_kernel<float>, dim3(grid), dim3(256), tile_px * 128 * 4, st, static_cast<const float*>(act_h), dw_w, dw_b,
                                                                static_cast<float*>(act_g), static_cast<float*>(pooled), sp, c, rows, tile_px);
    });
    P.flops_per_face += 2.0 * 9 * 2 * c * rpf;
  }
  // SCA (conditional_naf.py:119  x * sca(x)).  At 1x1 spatial on the tensor-core path the pooled mean is the gated
  // tensor itself and one face is one row, so the rescale rides in the SCA GEMM's epilogue (EPI_MUL, out of place
  // into act_h); elsewhere: SCA GEMM on the pooled vectors, then the per-face rescale of the gated rows.
  const bool sca_mul = bw.dw_folded && bf && h->tun.sca_mul;
  g_label = L0 + "sca";
  {
    GemmDesc d;
    d.M = B; d.N = c; d.K = c; d.A = bw.dw_folded ? act_g : pooled; d.lda = c; d.a_dtype = adt;
    d.W = bw.wsca; d.ldw = c; d.w_dtype = adt; d.bias = bw.bsca; d.epi = EPI_BIAS;
    d.out = sca_s; d.ldo = c; d.out_dtype = DT_F32;
    d.cta_target = h->tun.sca_target;
    if (sca_mul) { d.epi = EPI_MUL; d.resid = static_cast<const float*>(act_g); d.ldr = c; d.out = act_h; d.out_dtype = adt; }
    add_gemm(h, P, d, rows_cap(h, 1));
  }
  if (!sca_mul) {
    g_label = L0 + "scale_rows";
    add_op(P, [=](cudaStream_t st) {
      const size_t total8 = static_cast<size_t>(rows) * c / 8;
      if (bf) launch_k(scale_rows_kernel<bf16>, dim3(cdiv(total8, 256)), dim3(256), 0, st, static_cast<bf16*>(act_g), sca_s, total8, c, rpf);
      else launch_k(scale_rows_kernel<float>, dim3(cdiv(total8, 256)), dim3(256), 0, st, static_cast<float*>(act_g), sca_s, total8, c, rpf);
    });
  }
  g_label = L0 + "conv3";
  {  // conv3 (+beta) + residual
    GemmDesc d;
    d.M = rows; d.N = c; d.K = c; d.A = sca_mul ? act_h : act_g; d.lda = c; d.a_dtype = adt;
    d.W = bw.w3; d.ldw = c; d.w_dtype = adt; d.bias = bw.b3; d.epi = EPI_RESID;
    d.out = resid; d.ldo = c; d.out_dtype = DT_F32; d.resid = resid; d.ldr = c;
    add_gemm(h, P, d, rows_alloc);
  }
  // norm2 + modulation (shift_ffn = chunk 2, scale_ffn = chunk 3)
  g_label = L0 + "ln2";
  add_op(P, ln(bw.ln2_w, bw.ln2_b, bw.mod_off + 2 * c, bw.mod_off + 3 * c));
  g_label = L0 + "conv4";
  {  // conv4 + SimpleGate
    GemmDesc d;
    d.M = rows; d.N = 2 * c; d.K = c; d.A = act_a; d.lda = c; d.a_dtype = adt;
    d.W = bw.w4; d.ldw = c; d.w_dtype = adt; d.bias = bw.b4; d.epi = EPI_GATE;
    d.out = act_g; d.ldo = c; d.out_dtype = adt;
    add_gemm(h, P, d, rows_alloc);
  }
  g_label = L0 + "conv5";
  {  // conv5 (+gamma) + residual
    GemmDesc d;
    d.M = rows; d.N = c; d.K = c; d.A = act_g; d.lda = c; d.a_dtype = adt;
    d.W = bw.w5; d.ldw = c; d.w_dtype = adt; d.bias = bw.b5; d.epi = EPI_RESID;
    d.out = resid; d.ldo = c; d.out_dtype = DT_F32; d.resid = resid; d.ldr = c;
    TapInfo ti;
    ti.ptr = resid; ti.dtype = DT_F32; ti.C = c; ti.HW = rpf; ti.ld = c;
    add_gemm(h, P, d, rows_alloc, tapname, ti);
  }
}

void add_hca(hd_handle* h, Plan& P, int j, int level) {
  const HcaW& w = h->hca[j];
  const int B = P.batch, d = w.d, sp = w.sp, rpf = sp * sp, rows = B * rpf;
  const long long rows_alloc = rows_cap(h, rpf);
  const int adt = h->bf16 ? DT_BF16 : DT_F32;
  const bool bf = h->bf16;
  const float* fd = h->resid[level];
  const float *wc = w.wc, *ws = w.ws;
  const float* idc = j == 0 ? h->idc_add : nullptr;
  void* act_a = h->act_a;
  void* hca_out = h->hca_out;
  g_label = fmt("hca%d apply", j);
  add_op(P, [=](cudaStream_t st) {
    const size_t total8 = static_cast<size_t>(rows) * d / 8;
    if (bf) launch_k(hca_apply_kernel<bf16>, dim3(cdiv(total8, 256)), dim3(256), 0, st, fd, wc, ws, idc, static_cast<bf16*>(act_a), total8, d, rpf);
    else launch_k(hca_apply_kernel<float>, dim3(cdiv(total8, 256)), dim3(256), 0, st, fd, wc, ws, idc, static_cast<float*>(act_a), total8, d, rpf);
  });
  GemmDesc g;
  g.M = rows; g.N = d; g.A = act_a; g.a_dtype = adt; g.w_dtype = adt; g.bias = w.bf; g.epi = EPI_RELU;
  g.out = hca_out; g.ldo = d; g.out_dtype = adt; g.ldw = 9 * d;
  if (sp == 1) {  // only the centre tap sees a pixel
    g.a_mode = A_PLAIN; g.K = d; g.lda = d;
    g.W = static_cast<const char*>(w.wf) + static_cast<size_t>(4) * d * esize(adt);
  } else {
    g.a_mode = A_CONV3; g.K = 9 * d; g.sp = sp; g.C = d; g.lda = d; g.W = w.wf;
  }
  TapInfo ti;
  ti.ptr = hca_out; ti.dtype = adt; ti.C = d; ti.HW = rpf; ti.ld = d;
  g_label = fmt("hca%d conv3x3", j);
  add_gemm(h, P, g, rows_alloc, "hcas." + std::to_string(j), ti);
}

// Fused per-face kernel (face_block.cuh) over the blocks [first, first + count) of the 16x16 level:
// reads and writes the level's fp32 residual stream in place.
bool face_blocks_ok(hd_handle* h, size_t first, int count, bool debug) {
  if (!h->tun.face || debug || !h->bf16 || count > fb::MAX_BLOCKS) return false;
  for (int i = 0; i < count; ++i) {
    const BlockW& bw = h->blocks[first + i];
    if (bw.c != fb::C || h->sp[bw.level] != fb::SP || !bw.has_mod || bw.dw_folded || bw.wsca_t == nullptr) return false;
  }
  return true;
}

void add_face_blocks(hd_handle* h, Plan& P, size_t first, int count) {
  const int B = P.batch;
  const int c = fb::C, rpf = fb::PX;
  std::vector<CUtensorMap> maps;
  std::vector<fb::BlockParams> bps;
  std::vector<float> cum(fb::C, 0.f);
  auto add_map = [&](const void* base, int N) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)c * 2};
    cuuint32_t box[2] = {64, 128};
    encode_map(h, &m, base, 2, dims, strides, box);
    maps.push_back(m);
  };
  for (int i = 0; i < count; ++i) {
    const BlockW& bw = h->blocks[first + i];
    add_map(bw.w1, 2 * c);
    add_map(bw.w3, c);
    add_map(bw.w4, 2 * c);
    add_map(bw.w5, c);
    fb::BlockParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.ln1_w = bw.ln1_w; bp.ln1_b = bw.ln1_b; bp.ln2_w = bw.ln2_w; bp.ln2_b = bw.ln2_b;
    bp.b1 = bw.b1; bp.dw_w = bw.dw_w; bp.dw_b = bw.dw_b; bp.wsca_t = bw.wsca_t; bp.bsca = bw.bsca;
    bp.b4 = bw.b4; bp.mod_off = bw.mod_off;
    for (int k = 0; k < c; ++k) cum[k] += bw.b3_h[k];
    bp.cb3 = upload_f32(h, cum);
    for (int k = 0; k < c; ++k) cum[k] += bw.b5_h[k];
    bp.cb5 = upload_f32(h, cum);
    bps.push_back(bp);
    P.flops_per_face += 2.0 * rpf * 6.0 * c * c + 2.0 * c * c + 2.0 * 9 * 2 * c * rpf;
  }
  fb::Args a;
  memset(&a, 0, sizeof(a));
  CUtensorMap* d_maps = static_cast<CUtensorMap*>(h->arena.alloc(maps.size() * sizeof(CUtensorMap)));
  fb::BlockParams* d_bps = static_cast<fb::BlockParams*>(h->arena.alloc(bps.size() * sizeof(fb::BlockParams)));
  CUDA_CHECK(cudaMemcpy(d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(d_bps, bps.data(), bps.size() * sizeof(fb::BlockParams), cudaMemcpyHostToDevice));
  a.maps = d_maps;
  a.blocks = d_bps;
  a.n_blocks = count;
  a.zero_bias = upload_f32(h, std::vector<float>(c, 0.f));
  a.x = h->resid[h->blocks[first].level];
  a.mod_table = h->mod_table;
  a.mod_row_idx = h->row_idx;
  a.mod_stride = h->mod_stride;
  a.status = h->d_status;
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(fb::face_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fb::SMEM_BYTES));
    configured = true;
  }
  g_label = fmt("L%d c=%d face_block x%d (%s)", h->blocks[first].level, c, count, h->blocks[first].prefix.c_str());
  // the run's last block output is still observable (per-layer parity of the fused kernel itself)
  TapInfo ti;
  ti.ptr = a.x; ti.dtype = DT_F32; ti.C = c; ti.HW = rpf; ti.ld = c;
  std::string tap = h->blocks[first + count - 1].prefix;
  if (!tap.empty() && tap.back() == '.') tap.pop_back();
  a.first_wave = h->tun.face_warm ? h->sm_count : 0;
  if (getenv("HD_FACE_TRACE") != nullptr) {  // diagnostics: phase timeline of CTA 0, printed after every eager launch
    long long* tr = h->arena.get<long long>(64);
    a.trace = tr;
    a.trace_cta = atoi(getenv("HD_FACE_TRACE"));
    const int n_st = 5 + 6 * count;
    add_op(P, [=](cudaStream_t st) {
      launch_k(fb::face_block_kernel, dim3(B), dim3(fb::THREADS), fb::SMEM_BYTES, st, a);
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(st, &cs);
      if (cs != cudaStreamCaptureStatusNone) return;
      long long hst[64];
      cudaStreamSynchronize(st);
      cudaMemcpy(hst, tr, sizeof(hst), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[face_block trace, clocks since start]");
      for (int i = 1; i < n_st; ++i) fprintf(stderr, " %lld", hst[i] - hst[0]);
      fprintf(stderr, "\n");
    }, tap, ti);
    return;
  }
  add_op(P, [=](cudaStream_t st) { launch_k(fb::face_block_kernel, dim3(B), dim3(fb::THREADS), fb::SMEM_BYTES, st, a); }, tap, ti);
}

// Fused face-pair kernel (pair_block.cuh) over the blocks [first, first + count) of the 8x8 level.
bool pair_blocks_ok(hd_handle* h, size_t first, int count, bool debug) {
  if (!h->tun.pair || debug || !h->bf16 || count > pb::MAX_BLOCKS) return false;
  for (int i = 0; i < count; ++i) {
    const BlockW& bw = h->blocks[first + i];
    if (bw.c != pb::C || h->sp[bw.level] != pb::SP || !bw.has_mod || bw.dw_folded || bw.wsca_tb == nullptr) return false;
  }
  return true;
}

void add_pair_blocks(hd_handle* h, Plan& P, size_t first, int count) {
  const int B = P.batch;
  const int c = pb::C, rpf = pb::FPX;
  std::vector<CUtensorMap> maps;
  std::vector<pb::BlockParams> bps;
  auto add_map = [&](const void* base, int N) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)c * 2};
    cuuint32_t box[2] = {64, 128};
    encode_map(h, &m, base, 2, dims, strides, box);
    maps.push_back(m);
  };
  std::vector<float> cum(c, 0.f);
  for (int i = 0; i < count; ++i) {
    const BlockW& bw = h->blocks[first + i];
    add_map(bw.w1, 2 * c);
    add_map(bw.w3, c);
    add_map(bw.w4, 2 * c);
    add_map(bw.w5, c);
    pb::BlockParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.ln1_w = bw.ln1_w; bp.ln1_b = bw.ln1_b; bp.ln2_w = bw.ln2_w; bp.ln2_b = bw.ln2_b;
    bp.b1 = bw.b1; bp.dw_w = bw.dw_w; bp.dw_b = bw.dw_b; bp.wsca_t = static_cast<const bf16*>(bw.wsca_tb); bp.bsca = bw.bsca;
    bp.b4 = bw.b4; bp.mod_off = bw.mod_off;
    // the residual stream stays bias-free in tensor memory: x_true = x_tmem + (sum of the conv3 / conv5 biases so far)
    for (int k = 0; k < c; ++k) cum[k] += bw.b3_h[k];
    bp.cb3 = upload_f32(h, cum);
    for (int k = 0; k < c; ++k) cum[k] += bw.b5_h[k];
    bp.cb5 = upload_f32(h, cum);
    bps.push_back(bp);
    P.flops_per_face += 2.0 * rpf * 6.0 * c * c + 2.0 * c * c + 2.0 * 9 * 2 * c * rpf;
  }
  pb::Args a;
  memset(&a, 0, sizeof(a));
  CUtensorMap* d_maps = static_cast<CUtensorMap*>(h->arena.alloc(maps.size() * sizeof(CUtensorMap)));
  pb::BlockParams* d_bps = static_cast<pb::BlockParams*>(h->arena.alloc(bps.size() * sizeof(pb::BlockParams)));
  CUDA_CHECK(cudaMemcpy(d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(d_bps, bps.data(), bps.size() * sizeof(pb::BlockParams), cudaMemcpyHostToDevice));
  a.maps = d_maps;
  a.blocks = d_bps;
  a.n_blocks = count;
  a.n_faces = B;
  a.zero_bias = upload_f32(h, std::vector<float>(c, 0.f));
  a.x = h->resid[h->blocks[first].level];
  a.mod_table = h->mod_table;
  a.mod_row_idx = h->row_idx;
  a.mod_stride = h->mod_stride;
  a.status = h->d_status;
  a.warm = h->tun.face_warm ? 1 : 0;
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(pb::pair_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pb::SMEM_BYTES));
    configured = true;
  }
  g_label = fmt("L%d c=%d pair_block x%d (%s)", h->blocks[first].level, c, count, h->blocks[first].prefix.c_str());
  TapInfo ti;
  ti.ptr = a.x; ti.dtype = DT_F32; ti.C = c; ti.HW = rpf; ti.ld = c;
  std::string tap = h->blocks[first + count - 1].prefix;
  if (!tap.empty() && tap.back() == '.') tap.pop_back();
  if (getenv("HD_PAIR_TRACE") != nullptr) {  // diagnostics: phase timeline of one CTA, printed after every eager launch
    long long* tr = h->arena.get<long long>(64);
    a.trace = tr;
    a.trace_cta = atoi(getenv("HD_PAIR_TRACE"));
    const int n_st = 3 + 6 * count;
    add_op(P, [=](cudaStream_t st) {
      launch_k(pb::pair_block_kernel, dim3((B + 1) / 2), dim3(pb::THREADS), pb::SMEM_BYTES, st, a);
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(st, &cs);
      if (cs != cudaStreamCaptureStatusNone) return;
      long long hst[64];
      cudaStreamSynchronize(st);
      cudaMemcpy(hst, tr, sizeof(hst), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[pair_block trace, clocks since start]");
      for (int i = 1; i < n_st; ++i) fprintf(stderr, " %lld", hst[i] - hst[0]);
      fprintf(stderr, "\n");
    }, tap, ti);
    return;
  }
  add_op(P, [=](cudaStream_t st) { launch_k(pb::pair_block_kernel, dim3((B + 1) / 2), dim3(pb::THREADS), pb::SMEM_BYTES, st, a); }, tap, ti);
}

Plan* get_plan(hd_handle* h, int B, bool debug = false) {
  auto& cache = debug ? h->plans_dbg : h->plans;
  auto it = cache.find(B);
  if (it != cache.end()) return it->second.get();
  std::unique_ptr<Plan> up(new Plan());
  Plan& P = *up;
  P.batch = B;
  const int S = h->S;
  const bool bf = h->bf16;
  const int adt = bf ? DT_BF16 : DT_F32;

  {  // intro
    float* out = h->resid[0];
    const float *w = h->intro_w, *b = h->intro_b;
    TapInfo ti;
    ti.ptr = out; ti.dtype = DT_F32; ti.C = kWidth; ti.HW = S * S; ti.ld = kWidth;
    if (h->tun.edge_mma && h->intro_mma_hi != nullptr) {
      const bf16 *whi = h->intro_mma_hi, *wlo = h->intro_mma_lo;
      g_label = "intro conv3x3 mma.sync (3 x bf16 split)";
      add_op(P, [=](cudaStream_t st) { launch_k(edge::intro_mma_kernel, dim3(B), dim3(256), edge::IN_SMEM, st, h->cur_x, whi, wlo, b, out); },
             "intro", ti);
    } else {
      g_label = "intro conv3x3";
      add_op(P, [=](cudaStream_t st) {
        launch_k(intro_conv_kernel, dim3(B), dim3(256), (36 * 128 + 4 * (S + 2) * (S + 2)) * sizeof(float), st, h->cur_x, w, b, out, S);
      }, "intro", ti);
    }
    P.flops_per_face += 2.0 * 36 * 128 * S * S;
  }
  // ---- builders for one UNet stage ----
  auto emit_blocks = [&](size_t first, int count, const std::string& prefix) {
    if (face_blocks_ok(h, first, count, debug)) { add_face_blocks(h, P, first, count); return; }
    if (pair_blocks_ok(h, first, count, debug)) { add_pair_blocks(h, P, first, count); return; }
    for (int i = 0; i < count; ++i) add_block(h, P, h->blocks[first + i], prefix + std::to_string(i));
  };
  auto emit_down = [&](int l) {  // 2x2 stride-2 conv as space-to-depth + GEMM: resid[l] -> resid[l + 1]
    const int Bq = P.batch;
    const int c = h->c[l], n = h->sp[l], rows_out = Bq * (n / 2) * (n / 2);
    const float* src = h->resid[l];
    void* act_a = h->act_a;
    g_label = fmt("down%d s2d", l);
    add_op(P, [=](cudaStream_t st) {
      const size_t total8 = static_cast<size_t>(rows_out) * 4 * c / 8;
      if (bf) launch_k(s2d_kernel<bf16>, dim3(cdiv(total8, 256)), dim3(256), 0, st, src, static_cast<bf16*>(act_a), Bq, n, c);
      else launch_k(s2d_kernel<float>, dim3(cdiv(total8, 256)), dim3(256), 0, st, src, static_cast<float*>(act_a), Bq, n, c);
    });
    float* out = h->resid[l + 1];
    GemmDesc d;
    d.M = rows_out; d.N = 2 * c; d.K = 4 * c; d.A = act_a; d.lda = 4 * c; d.a_dtype = adt;
    d.W = h->down_w[l]; d.ldw = 4 * c; d.w_dtype = adt; d.bias = h->down_b[l]; d.epi = EPI_BIAS;
    d.out = out; d.ldo = 2 * c; d.out_dtype = DT_F32;
    TapInfo ti;
    ti.ptr = out; ti.dtype = DT_F32; ti.C = 2 * c; ti.HW = (n / 2) * (n / 2); ti.ld = 2 * c;
    g_label = fmt("down%d", l);
    add_gemm(h, P, d, rows_cap(h, (n / 2) * (n / 2)), "downs." + std::to_string(l), ti);
  };
  auto emit_up = [&](int L) {  // 1x1 conv + PixelShuffle(2) + skip add: level 4 - L -> resid[3 - L] (in place on the skip)
    const int Bq = P.batch;
    const int lin = 4 - L, lout = 3 - L;
    const int cin = h->c[lin], n = h->sp[lin], rows_in = Bq * n * n;
    const void* a_ptr;
    if (h->fused) {
      a_ptr = h->hca_out;
    } else {
      const float* src = h->resid[lin];
      void* act_a = h->act_a;
      a_ptr = act_a;
      g_label = fmt("up%d cast", L);
      add_op(P, [=](cudaStream_t st) {
        const size_t total8 = static_cast<size_t>(rows_in) * cin / 8;
        if (bf) launch_k(cast_kernel<bf16>, dim3(cdiv(total8, 256)), dim3(256), 0, st, src, static_cast<bf16*>(act_a), total8);
        else launch_k(cast_kernel<float>, dim3(cdiv(total8, 256)), dim3(256), 0, st, src, static_cast<float*>(act_a), total8);
      });
    }
    float* out = h->resid[lout];
    GemmDesc d;
    d.M = rows_in; d.N = 2 * cin; d.K = cin; d.A = a_ptr; d.lda = cin; d.a_dtype = adt;
    d.W = h->up_w[L]; d.ldw = cin; d.w_dtype = adt; d.bias = nullptr; d.epi = EPI_PIXSHUF; d.sp = n;
    d.out = out; d.ldo = cin / 2; d.out_dtype = DT_F32;
    TapInfo ti;
    ti.ptr = out; ti.dtype = DT_F32; ti.C = cin / 2; ti.HW = 4 * n * n; ti.ld = cin / 2;
    g_label = fmt("up%d", L);
    add_gemm(h, P, d, rows_cap(h, n * n), "ups." + std::to_string(L), ti);
  };
  // block index of the first block of each stage, in execution order
  size_t enc_first[4], dec_first[4], mid_first;
  {
    size_t bi = 0;
    for (int l = 0; l < 4; ++l) { enc_first[l] = bi; bi += kEncBlocks[l]; }
    mid_first = bi; bi += kMidBlocks;
    for (int L = 0; L < 4; ++L) { dec_first[L] = bi; bi += kDecBlocks[L]; }
  }
  emit_blocks(enc_first[0], kEncBlocks[0], "encoders.0.");
  emit_down(0);
  emit_blocks(enc_first[1], kEncBlocks[1], "encoders.1.");
  emit_down(1);
  emit_blocks(enc_first[2], kEncBlocks[2], "encoders.2.");
  emit_down(2);
  emit_blocks(enc_first[3], kEncBlocks[3], "encoders.3.");
  emit_down(3);
  emit_blocks(mid_first, kMidBlocks, "middle_blks.");
  if (h->fused) add_hca(h, P, 0, 4);
  emit_up(0);
  emit_blocks(dec_first[0], kDecBlocks[0], "decoders.0.");
  if (h->fused) add_hca(h, P, 1, 3);
  emit_up(1);
  emit_blocks(dec_first[1], kDecBlocks[1], "decoders.1.");
  if (h->fused) add_hca(h, P, 2, 2);
  emit_up(2);
  emit_blocks(dec_first[2], kDecBlocks[2], "decoders.2.");
  if (h->fused) add_hca(h, P, 3, 1);
  emit_up(3);
  emit_blocks(dec_first[3], kDecBlocks[3], "decoders.3.");
  if (h->fused) add_hca(h, P, 4, 0);
  {  // ending
    const float *w = h->end_w, *b = h->end_b;
    const void* in = h->fused ? h->hca_out : static_cast<const void*>(h->resid[0]);
    const bool in_bf = h->fused && bf;
    if (in_bf && h->tun.edge_mma && h->end_mma_hi != nullptr) {
      edge::EndArgs ea;
      memset(&ea, 0, sizeof(ea));
      ea.x = static_cast<const bf16*>(in); ea.w_hi = h->end_mma_hi; ea.w_lo = h->end_mma_lo; ea.bias = b;
      g_label = "ending conv3x3 mma.sync";
      add_op(P, [=](cudaStream_t st) {
        edge::EndArgs e2 = ea;
        e2.eps = h->cur_eps;
        launch_k(edge::ending_mma_kernel<false>, dim3(B), dim3(256), edge::END_SMEM, st, e2);
      });
      P.ending_idx = static_cast<int>(P.ops.size()) - 1;
    } else if (S > 16) {
      const int band = in_bf ? 16 : 8;   // image rows per block: (band + 2) * S * 128 elements of shared memory
      g_label = "ending conv3x3 (row bands)";
      add_op(P, [=](cudaStream_t st) {
        const size_t wbytes = 4 * 9 * 128 * sizeof(float);
        if (in_bf) launch_k(ending_conv_band_kernel<bf16>, dim3(B, S / band), dim3(256), static_cast<size_t>(band + 2) * S * 128 * 2 + wbytes, st, static_cast<const bf16*>(in), w, b, h->cur_eps, S, band);
        else launch_k(ending_conv_band_kernel<float>, dim3(B, S / band), dim3(256), static_cast<size_t>(band + 2) * S * 128 * 4 + wbytes, st, static_cast<const float*>(in), w, b, h->cur_eps, S, band);
      });
    } else {
    g_label = "ending conv3x3";
    add_op(P, [=](cudaStream_t st) {
      const size_t wbytes = 4 * 9 * 128 * sizeof(float);
      if (in_bf) launch_k(ending_conv_kernel<bf16>, dim3(B), dim3(256), S * S * 128 * 2 + wbytes, st, static_cast<const bf16*>(in), w, b, h->cur_eps, B, S);
      else launch_k(ending_conv_kernel<float>, dim3(B), dim3(256), S * S * 128 * 4 + wbytes, st, static_cast<const float*>(in), w, b, h->cur_eps, B, S);
    });
    }
    P.flops_per_face += 2.0 * 9 * 128 * 4 * S * S;
  }
  if (P.last_tc && P.first_w != nullptr) { P.last_tc->args.pf_ptr = P.first_w; P.last_tc->args.pf_bytes = P.first_w_bytes; }
  Plan* raw = up.get();
  cache[B] = std::move(up);
  return raw;
}

