// hd_idc.inl — part of hd_lib.cu (one translation unit; included there in order, not compiled on its own).
// IDC ResNet-50: weights (BatchNorm folded) and forward.
// ------------------------------------------------------------------------------------------------
// IDC identity network: ResNet-50 trunk (models/idc/model.py:102-166), once per face.
// Every conv is followed by an eval-mode BatchNorm, folded into the packed weights / bias.  The 64-plane
// tensors of layer1 are zero-padded to 128 channels so that every GEMM has N % 128 == 0.
// ------------------------------------------------------------------------------------------------
constexpr int kIdcChunk = 64;  // faces per pass: bounds the activation workspace (~5 MB per face in bf16)
constexpr int kIdcLayers[4] = {3, 4, 6, 3};
constexpr int kIdcOut = 2048;

IdcConvW idc_pack(hd_handle* h, const std::string& conv, const std::string& bn, int N, int C, int taps, int Npad,
                  int Cpad, int wdt) {
  std::vector<float> cb = host_vec(h, need(h, conv + "bias", {N}));
  std::vector<float> shift;
  std::vector<float> rs = bn_scale(h, bn, N, &shift, cb);
  const float* src = dev_src(h, need(h, conv + "weight", {N, C, taps}));
  IdcConvW cw;
  cw.N = Npad;
  cw.K = taps * Cpad;
  const size_t total = static_cast<size_t>(Npad) * cw.K;
  cw.w = h->arena.alloc(total * esize(wdt));
  const float* drs = upload_f32(h, rs);
  if (wdt == DT_BF16)
    idc_pack_conv_kernel<bf16><<<cdiv(total, 256), 256, 0, h->stream>>>(src, static_cast<bf16*>(cw.w), drs, N, C, taps, Npad, Cpad);
  else
    idc_pack_conv_kernel<float><<<cdiv(total, 256), 256, 0, h->stream>>>(src, static_cast<float*>(cw.w), drs, N, C, taps, Npad, Cpad);
  CUDA_CHECK(cudaGetLastError());
  shift.resize(Npad, 0.f);
  cw.b = upload_f32(h, shift);
  return cw;
}

void load_idc_impl(hd_handle* h) {
  const int wdt = h->bf16 ? DT_BF16 : DT_F32;
  const size_t es = esize(wdt);
  IdcW& I = h->idc;
  I.H = 8 * h->S;  // cr_face is the pixel-space face: 8x the latent size (SD-VAE factor), 128 for 16x16 latents
  if (I.H % 64 != 0) HD_THROW(HD_ERR_UNSUPPORTED, "IDC needs an image size that is a multiple of 64 (latent size %d)", h->S);
  I.cap = kIdcChunk;
  {  // stem: conv1 (64,3,7,7) no bias + batch_norm1 -> [147][64] fp32, BN scale folded (idc/model.py:107-110)
    std::vector<float> zero(64, 0.f), shift;
    std::vector<float> rs = bn_scale(h, "batch_norm1.", 64, &shift, zero);
    std::vector<float> w = host_vec(h, need(h, "conv1.weight", {64, 147}));
    std::vector<float> t(147 * 64);
    for (int o = 0; o < 64; ++o)
      for (int k = 0; k < 147; ++k) t[static_cast<size_t>(k) * 64 + o] = w[static_cast<size_t>(o) * 147 + k] * rs[o];
    I.stem_w = upload_f32(h, t);
    I.stem_b = upload_f32(h, shift);
  }
  I.blocks.clear();
  int cin = 64, n = I.H / 4;
  for (int li = 0; li < 4; ++li) {
    const int planes = 64 << li;
    for (int bi = 0; bi < kIdcLayers[li]; ++bi) {
      IdcBlockW b;
      const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(bi) + ".";
      b.stride = (bi == 0 && li > 0) ? 2 : 1;
      b.cin = cin; b.planes = planes; b.pp = std::max(planes, 128); b.n_in = n;
      b.has_proj = bi == 0;
      b.c1 = idc_pack(h, p + "conv1.", p + "batch_norm1.", planes, cin, 1, b.pp, cin, wdt);
      b.c2 = idc_pack(h, p + "conv2.", p + "batch_norm2.", planes, planes, 9, b.pp, b.pp, wdt);
      b.c3 = idc_pack(h, p + "conv3.", p + "batch_norm3.", 4 * planes, planes, 1, 4 * planes, b.pp, wdt);
      if (b.has_proj) b.proj = idc_pack(h, p + "i_downsample.0.", p + "i_downsample.1.", 4 * planes, cin, 1, 4 * planes, cin, wdt);
      I.blocks.push_back(b);
      cin = 4 * planes;
      n /= b.stride;
    }
  }
  // workspace for one chunk of faces (elements per face; see get_idc_plan for who writes what)
  const size_t H = I.H, cap = I.cap, n1 = H / 4;
  const size_t x_elems = n1 * n1 * 256;                       // widest residual tensor: layer1 output
  const size_t t_elems = n1 * n1 * 128;                       // conv1 / conv2 outputs (layer1 padded, layer2.0.conv1)
  const size_t col_elems = (n1 / 2) * (n1 / 2) * 9 * 128;     // patches of layer2.0.conv2 (the largest gather)
  I.stem_out = h->arena.alloc(cap * (H / 2) * (H / 2) * 64 * es);
  for (int k = 0; k < 2; ++k) {
    I.xf[k] = h->arena.get<float>(cap * x_elems);
    I.xb[k] = h->arena.alloc(cap * x_elems * es);
  }
  I.t1 = h->arena.alloc(cap * t_elems * es);
  I.t2 = h->arena.alloc(cap * t_elems * es);
  I.col = h->arena.alloc(cap * col_elems * es);
  I.stage = h->arena.get<float>(cap * 3 * H * H);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (void* p : h->temp_dev) cudaFree(p);
  h->temp_dev.clear();
  h->src.clear();
  I.loaded = true;
  h->idc_plans.clear();
}

Plan* get_idc_plan(hd_handle* h, int B) {
  auto it = h->idc_plans.find(B);
  if (it != h->idc_plans.end()) return it->second.get();
  std::unique_ptr<Plan> up(new Plan());
  Plan& P = *up;
  P.batch = B;
  const bool bf = h->bf16;
  const int adt = bf ? DT_BF16 : DT_F32;
  const IdcW& I = h->idc;
  const int H = I.H, cap = I.cap;
  {  // stem 7x7 s2 + BN + ReLU -> stem_out NHWC [B][H/2][H/2][64]
    const float *w = I.stem_w, *b = I.stem_b;
    void* out = I.stem_out;
    const size_t smem = (147 * 64 + 3 * kStemPatch * 22) * sizeof(float);
    g_label = "idc stem conv7x7 s2";
    add_op(P, [=](cudaStream_t st) {
      if (bf) launch_k(idc_stem_kernel<bf16>, dim3(H / 16, H / 16, B), dim3(256), smem, st, h->idc_in, w, b, static_cast<bf16*>(out), H);
      else launch_k(idc_stem_kernel<float>, dim3(H / 16, H / 16, B), dim3(256), smem, st, h->idc_in, w, b, static_cast<float*>(out), H);
    });
    P.flops_per_face += 2.0 * 147 * 64 * (H / 2) * (H / 2);
  }
  {  // max-pool 3x3 s2 -> xf[0] (identity) + xb[0] (operand), NHWC [B][H/4][H/4][64]
    const void* in = I.stem_out;
    float* of = I.xf[0];
    void* ot = I.xb[0];
    const int n = H / 2;
    const size_t total = static_cast<size_t>(B) * (n / 2) * (n / 2) * 8;
    g_label = "idc maxpool";
    add_op(P, [=](cudaStream_t st) {
      if (bf) launch_k(idc_maxpool_kernel<bf16>, dim3(cdiv(total, 256)), dim3(256), 0, st, static_cast<const bf16*>(in), of, static_cast<bf16*>(ot), B, n, 64);
      else launch_k(idc_maxpool_kernel<float>, dim3(cdiv(total, 256)), dim3(256), 0, st, static_cast<const float*>(in), of, static_cast<float*>(ot), B, n, 64);
    });
  }
  auto gather = [&](const void* in, void* out, int n, int C, int k, int stride, int pad, const std::string& label) {
    const int no = (n + 2 * pad - k) / stride + 1;
    const size_t total = static_cast<size_t>(B) * no * no * k * k * (C / 8);
    g_label = label;
    add_op(P, [=](cudaStream_t st) {
      if (bf) launch_k(idc_gather_kernel<bf16>, dim3(cdiv(total, 256)), dim3(256), 0, st, static_cast<const bf16*>(in), static_cast<bf16*>(out), B, n, C, k, stride, pad, no);
      else launch_k(idc_gather_kernel<float>, dim3(cdiv(total, 256)), dim3(256), 0, st, static_cast<const float*>(in), static_cast<float*>(out), B, n, C, k, stride, pad, no);
    });
  };
  int cur = 0;
  for (size_t bi = 0; bi < I.blocks.size(); ++bi) {
    const IdcBlockW& b = I.blocks[bi];
    const int n = b.n_in, no = n / b.stride;
    const int rows_in = B * n * n, rows_out = B * no * no;
    const long long alloc_in = static_cast<long long>(cap) * n * n, alloc_out = static_cast<long long>(cap) * no * no;
    const std::string L0 = fmt("idc b%d ", static_cast<int>(bi));
    {  // conv1 1x1 + BN + ReLU (idc/model.py:41)
      GemmDesc d;
      d.M = rows_in; d.N = b.pp; d.K = b.cin; d.A = I.xb[cur]; d.lda = b.cin; d.a_dtype = adt;
      d.W = b.c1.w; d.ldw = b.cin; d.w_dtype = adt; d.bias = b.c1.b; d.epi = EPI_RELU;
      d.out = I.t1; d.ldo = b.pp; d.out_dtype = adt;
      g_label = L0 + "conv1";
      add_gemm(h, P, d, alloc_in);
    }
    {  // conv2 3x3 (stride 1: implicit GEMM; stride 2: patch gather + GEMM) + BN + ReLU (idc/model.py:43)
      GemmDesc d;
      d.M = rows_out; d.N = b.pp; d.K = 9 * b.pp; d.a_dtype = adt;
      d.W = b.c2.w; d.ldw = 9 * b.pp; d.w_dtype = adt; d.bias = b.c2.b; d.epi = EPI_RELU;
      d.out = I.t2; d.ldo = b.pp; d.out_dtype = adt;
      if (b.stride == 1) {
        d.A = I.t1; d.lda = b.pp; d.a_mode = A_CONV3; d.sp = n; d.C = b.pp;
      } else {
        gather(I.t1, I.col, n, b.pp, 3, 2, 1, L0 + "conv2 patches s2");
        d.A = I.col; d.lda = 9 * b.pp;
      }
      g_label = L0 + "conv2";
      add_gemm(h, P, d, alloc_out);
    }
    int nxt = cur;
    if (b.has_proj) {  // projection shortcut: 1x1 (stride s) + BN (idc/model.py:141-149)
      nxt = cur ^ 1;
      GemmDesc d;
      d.M = rows_out; d.N = 4 * b.planes; d.K = b.cin; d.a_dtype = adt; d.lda = b.cin;
      d.W = b.proj.w; d.ldw = b.cin; d.w_dtype = adt; d.bias = b.proj.b; d.epi = EPI_BIAS;
      d.out = I.xf[nxt]; d.ldo = 4 * b.planes; d.out_dtype = DT_F32;
      if (b.stride == 1) {
        d.A = I.xb[cur];
      } else {
        gather(I.xb[cur], I.col, n, b.cin, 1, 2, 0, L0 + "proj rows s2");
        d.A = I.col;
      }
      g_label = L0 + "proj";
      add_gemm(h, P, d, b.stride == 1 ? alloc_in : alloc_out);
    }
    {  // conv3 1x1 + BN + identity (idc/model.py:45-51)
      GemmDesc d;
      d.M = rows_out; d.N = 4 * b.planes; d.K = b.pp; d.A = I.t2; d.lda = b.pp; d.a_dtype = adt;
      d.W = b.c3.w; d.ldw = b.pp; d.w_dtype = adt; d.bias = b.c3.b; d.epi = EPI_RESID;
      d.out = I.xf[nxt]; d.ldo = 4 * b.planes; d.out_dtype = DT_F32; d.resid = I.xf[nxt]; d.ldr = 4 * b.planes;
      g_label = L0 + "conv3";
      add_gemm(h, P, d, alloc_out);
    }
    float* xf = I.xf[nxt];
    if (bi + 1 < I.blocks.size()) {  // ReLU (idc/model.py:52) -> identity + operand of the next block
      void* xb = I.xb[nxt];
      const size_t total8 = static_cast<size_t>(rows_out) * 4 * b.planes / 8;
      g_label = L0 + "relu+cast";
      add_op(P, [=](cudaStream_t st) {
        if (bf) launch_k(idc_relu_cast_kernel<bf16>, dim3(cdiv(total8, 256)), dim3(256), 0, st, xf, static_cast<bf16*>(xb), total8);
        else launch_k(idc_relu_cast_kernel<float>, dim3(cdiv(total8, 256)), dim3(256), 0, st, xf, static_cast<float*>(xb), total8);
      });
    } else {  // ReLU + global average pool -> (B, 2048, 1, 1) (idc/model.py:132-133)
      const int hw = no * no, C = 4 * b.planes;
      g_label = "idc relu+avgpool";
      add_op(P, [=](cudaStream_t st) {
        launch_k(idc_relu_avgpool_kernel, dim3(cdiv(B * C, 256)), dim3(256), 0, st, static_cast<const float*>(xf), h->idc_out, B, hw, C);
      });
    }
    cur = nxt;
  }
  Plan* raw = up.get();
  h->idc_plans[B] = std::move(up);
  return raw;
}

