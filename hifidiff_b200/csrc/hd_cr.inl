// hd_cr.inl — part of hd_lib.cu (one translation unit; included there in order, not compiled on its own).
// CoarseRestoration: weights and forward plan.
// ------------------------------------------------------------------------------------------------
// CoarseRestoration: NAFNet U-Net with a spatial transformer after every stage
// (models/cr/model.py:8-88, models/cr/stn.py:9-52), once per face before the sampling loop, fp32 throughout.
// ------------------------------------------------------------------------------------------------
constexpr int kCrC[5] = {32, 64, 128, 256, 512};
constexpr int kCrRes[5] = {128, 64, 32, 16, 8};

float* cr_mat(hd_handle* h, const std::string& name, int N, int Kd, int taps, const std::vector<int>* perm,
              const std::vector<float>* rs) {
  return static_cast<float*>(pack_matrix(h, need(h, name, {N, Kd}), N, Kd, taps, perm, rs, DT_F32));
}
float* cr_vec(hd_handle* h, const std::string& name, int n) { return upload_f32(h, host_vec(h, need(h, name, {n}))); }

void load_cr_block(hd_handle* h, CrBlockW& b, const std::string& p, int c) {
  b.c = c;
  b.ln1_w = cr_vec(h, p + "norm1.weight", c); b.ln1_b = cr_vec(h, p + "norm1.bias", c);
  b.ln2_w = cr_vec(h, p + "norm2.weight", c); b.ln2_b = cr_vec(h, p + "norm2.bias", c);
  auto beta = host_vec(h, need(h, p + "beta", {c}));
  auto gamma = host_vec(h, need(h, p + "gamma", {c}));
  b.w1 = cr_mat(h, p + "conv1.weight", 2 * c, c, 1, nullptr, nullptr);
  b.b1 = cr_vec(h, p + "conv1.bias", 2 * c);
  {  // depthwise 3x3: [2c,1,3,3] -> [9][2c]
    auto w = host_vec(h, need(h, p + "conv2.weight", {2 * c, 9}));
    std::vector<float> t(static_cast<size_t>(18) * c);
    for (int ch = 0; ch < 2 * c; ++ch)
      for (int tap = 0; tap < 9; ++tap) t[static_cast<size_t>(tap) * 2 * c + ch] = w[static_cast<size_t>(ch) * 9 + tap];
    b.dw_w = upload_f32(h, t);
    b.dw_b = cr_vec(h, p + "conv2.bias", 2 * c);
  }
  b.wsca = cr_mat(h, p + "sca.1.weight", c, c, 1, nullptr, nullptr);
  b.bsca = cr_vec(h, p + "sca.1.bias", c);
  auto b3 = host_vec(h, need(h, p + "conv3.bias", {c}));
  auto b5 = host_vec(h, need(h, p + "conv5.bias", {c}));
  for (int i = 0; i < c; ++i) { b3[i] *= beta[i]; b5[i] *= gamma[i]; }
  b.w3 = cr_mat(h, p + "conv3.weight", c, c, 1, nullptr, &beta);   // y = inp + beta * (W3 x + b3)
  b.b3 = upload_f32(h, b3);
  b.w4 = cr_mat(h, p + "conv4.weight", 2 * c, c, 1, nullptr, nullptr);
  b.b4 = cr_vec(h, p + "conv4.bias", 2 * c);
  b.w5 = cr_mat(h, p + "conv5.weight", c, c, 1, nullptr, &gamma);
  b.b5 = upload_f32(h, b5);
  if (c >= 128 && h->bf16 && h->tun.cr_tc) {
    auto split = [&](const float* w, int N, int K) {
      bf16* out = h->arena.get<bf16>(static_cast<size_t>(N) * 3 * K);
      const size_t total = static_cast<size_t>(N) * (K / 8);
      cr_split3_kernel<<<cdiv(total, 256), 256, 0, h->stream>>>(w, out, static_cast<size_t>(N), K, 1);
      CUDA_CHECK(cudaGetLastError());
      return out;
    };
    b.w1s = split(b.w1, 2 * c, c); b.w3s = split(b.w3, c, c); b.w4s = split(b.w4, 2 * c, c); b.w5s = split(b.w5, c, c);
  }
}

// conv weight OIHW [O][I][k][k] -> [O][k][k][I] (channels innermost, as the NHWC kernels read them)
float* cr_conv_ohwi(hd_handle* h, const std::string& name, int O, int I, int k) {
  auto w = host_vec(h, need(h, name, {O, I, k * k}));
  std::vector<float> t(w.size());
  for (int o = 0; o < O; ++o)
    for (int i = 0; i < I; ++i)
      for (int q = 0; q < k * k; ++q) t[(static_cast<size_t>(o) * k * k + q) * I + i] = w[(static_cast<size_t>(o) * I + i) * k * k + q];
  return upload_f32(h, t);
}

void load_cr_stn(hd_handle* h, CrStnW& s, const std::string& p, int c, int res) {
  // kernel sizes and regressor width as STNBlock.__init__ derives them (stn.py:13-22,29-33)
  if (res <= 8) { s.k1 = 3; s.k2 = 1; } else if (res <= 16) { s.k1 = 5; s.k2 = 3; } else if (res <= 32) { s.k1 = 7; s.k2 = 5; } else { s.k1 = 9; s.k2 = 7; }
  s.n1 = (res - s.k1 + 1) / 2;
  s.n2 = (s.n1 - s.k2 + 1) / 2;
  s.fc = 10 * s.n2 * s.n2;
  s.hid = static_cast<int>(std::sqrt(static_cast<double>(s.fc)));
  if (s.hid > 96) HD_THROW(HD_ERR_UNSUPPORTED, "STN regressor width %d", s.hid);
  s.w1 = cr_conv_ohwi(h, p + "localization.0.weight", 8, c, s.k1);
  if ((c == 32 || c % 64 == 0) && h->bf16 && h->tun.cr_tc && h->tun.cr_stn_mma) {
    // edge::stn_conv_mma_kernel's order: [pass][tap][chunk] k-steps of 512 bytes, each {hi, lo} x {k 0-7, k 8-15} 8x8
    // matrices [n = 8][8 k]; fp16 hi + lo of w * 2^e with max |w| * 2^e in [2^14, 2^15)
    auto w = host_vec(h, need(h, p + "localization.0.weight", {8, c, s.k1 * s.k1}));  // [o][i][tap]
    const int ch = c == 32 ? 32 : 64, cch = ch / 16, taps = s.k1 * s.k1;
    float wmax = 0.f;
    for (float f : w) wmax = std::max(wmax, std::fabs(f));
    const float wscale = wmax > 0.f ? std::ldexp(1.f, 14 - std::ilogb(wmax)) : 1.f;
    s.w1_unscale = 1.f / wscale;
    std::vector<uint16_t> v(static_cast<size_t>(c / 16) * taps * 256);
    auto to_half = [](float f) {
      const __half_raw r = static_cast<__half_raw>(__float2half_rn(f));
      return r.x;
    };
    auto from_half = [](uint16_t x) {
      __half_raw r;
      r.x = x;
      return __half2float(__half(r));
    };
    for (int pass = 0; pass < c / ch; ++pass)
      for (int tap = 0; tap < taps; ++tap)
        for (int cc = 0; cc < cch; ++cc) {
          uint16_t* blk = v.data() + ((static_cast<size_t>(pass) * taps + tap) * cch + cc) * 256;
          for (int o = 0; o < 8; ++o)
            for (int kk = 0; kk < 16; ++kk) {
              const float f = w[(static_cast<size_t>(o) * c + pass * ch + cc * 16 + kk) * taps + tap] * wscale;
              const uint16_t hi = to_half(f);
              const int at = (kk >> 3) * 64 + o * 8 + (kk & 7);
              blk[at] = hi;
              blk[128 + at] = to_half(f - from_half(hi));
            }
        }
    s.w1_mma = static_cast<__half*>(h->arena.alloc(v.size() * 2));
    CUDA_CHECK(cudaMemcpy(s.w1_mma, v.data(), v.size() * 2, cudaMemcpyHostToDevice));
  }
  s.b1 = cr_vec(h, p + "localization.0.bias", 8);
  s.w2 = cr_conv_ohwi(h, p + "localization.3.weight", 10, 8, s.k2);
  s.b2 = cr_vec(h, p + "localization.3.bias", 10);
  {  // fc_loc.0: columns from the reference's (C,H,W) flattening to the kernels' (H,W,C)
    auto w = host_vec(h, need(h, p + "fc_loc.0.weight", {s.hid, s.fc}));
    std::vector<float> t(w.size());
    const int hw = s.n2 * s.n2;
    for (int j = 0; j < s.hid; ++j)
      for (int q = 0; q < hw; ++q)
        for (int o = 0; o < 10; ++o) t[static_cast<size_t>(j) * s.fc + q * 10 + o] = w[static_cast<size_t>(j) * s.fc + o * hw + q];
    s.f1 = upload_f32(h, t);
  }
  s.fb1 = cr_vec(h, p + "fc_loc.0.bias", s.hid);
  s.f2 = upload_f32(h, host_vec(h, need(h, p + "fc_loc.2.weight", {6, s.hid})));
  s.fb2 = cr_vec(h, p + "fc_loc.2.bias", 6);
}

// Split copies of an fp32 weight matrix [N, K] for the mma.sync GEMMs of CoarseRestoration's shallow stages, made
// once per matrix (at load for every matrix that can take the path; the look-up at plan time only builds one if a
// shape was not foreseen): scaled fp16 hi + lo for gemm_mma3h (K = 32 / 64), tf32 hi + lo for gemm_mma3.
const CrW::SplitH& cr_split_h(hd_handle* h, const float* W, int N, int K) {
  auto& cache = h->cr.split_h;
  auto it = cache.find(W);
  if (it != cache.end()) return it->second;
  const size_t nw = static_cast<size_t>(N) * K;
  std::vector<float> wf(nw);
  CUDA_CHECK(cudaMemcpyAsync(wf.data(), W, nw * 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  float wmax = 0.f;
  for (float f : wf) wmax = std::max(wmax, std::fabs(f));
  const float wscale = wmax > 0.f && std::isfinite(wmax) ? std::ldexp(1.f, 14 - std::ilogb(wmax)) : 1.f;
  std::vector<uint16_t> vh(nw), vl(nw);
  for (size_t i = 0; i < nw; ++i) {
    const float f = wf[i] * wscale;
    const __half hh = __float2half_rn(f);
    vh[i] = static_cast<__half_raw>(hh).x;
    vl[i] = static_cast<__half_raw>(__float2half_rn(f - __half2float(hh))).x;
  }
  CrW::SplitH sp;
  sp.hi = static_cast<__half*>(h->arena.alloc(nw * 2));
  sp.lo = static_cast<__half*>(h->arena.alloc(nw * 2));
  sp.unscale = 1.f / wscale;
  CUDA_CHECK(cudaMemcpy(sp.hi, vh.data(), nw * 2, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(sp.lo, vl.data(), nw * 2, cudaMemcpyHostToDevice));
  return cache.emplace(W, sp).first->second;
}
const std::pair<float*, float*>& cr_split_tf32(hd_handle* h, const float* W, int N, int K) {
  auto& cache = h->cr.split_hl;
  auto it = cache.find(W);
  if (it != cache.end()) return it->second;
  const size_t nw = static_cast<size_t>(N) * K;
  float *hi = h->arena.get<float>(nw), *lo = h->arena.get<float>(nw);
  mma3::split_hl_kernel<<<cdiv(nw, static_cast<size_t>(256)), 256, 0, h->stream>>>(W, hi, lo, nw);
  CUDA_CHECK(cudaGetLastError());
  return cache.emplace(W, std::make_pair(hi, lo)).first->second;
}

void load_cr_impl(hd_handle* h) {
  CrW& R = h->cr;
  R.H = 8 * h->S;
  if (R.H != 128) HD_THROW(HD_ERR_UNSUPPORTED, "CoarseRestoration is built for 128x128 faces (latent size 16)");
  R.cap = h->tun.cr_chunk;
  {  // intro (32,3,3,3) -> [27][32]; outro (3,32,3,3) -> [3][9][32]
    auto w = host_vec(h, need(h, "intro.weight", {32, 27}));
    std::vector<float> t(27 * 32);
    for (int o = 0; o < 32; ++o)
      for (int k = 0; k < 27; ++k) t[static_cast<size_t>(k) * 32 + o] = w[static_cast<size_t>(o) * 27 + k];
    R.intro_w = upload_f32(h, t);
    R.intro_b = cr_vec(h, "intro.bias", 32);
    R.outro_w = cr_conv_ohwi(h, "outro.weight", 3, 32, 3);
    R.outro_b = cr_vec(h, "outro.bias", 3);
  }
  R.stages.clear();
  const int enc_naf[4] = {2, 2, 4, 8};
  auto add_stage = [&](const std::string& p, int level, int num_naf, int sampling) {
    CrStageW st;
    st.c = kCrC[level]; st.res = kCrRes[level]; st.sampling = sampling;
    st.blocks.resize(num_naf);
    for (int i = 0; i < num_naf; ++i) load_cr_block(h, st.blocks[i], p + "nfbs." + std::to_string(i) + ".", st.c);
    load_cr_stn(h, st.stn, p + "stn.", st.c, st.res);
    const int c = st.c;
    if (sampling == 1) {  // Conv2d(c, 2c, 2, 2): K order (dy, dx, c) of the space-to-depth rows
      st.samp_w = cr_mat(h, p + "sampling.weight", 2 * c, 4 * c, 4, nullptr, nullptr);
      st.samp_b = cr_vec(h, p + "sampling.bias", 2 * c);
    } else if (sampling == 2) {  // Conv2d(c, 2c, 1, bias=False) + PixelShuffle(2): rows grouped by quadrant
      const int N = 2 * c, quarter = N / 4;
      std::vector<int> perm(N);
      for (int n = 0; n < N; ++n) perm[n] = 4 * (n % quarter) + n / quarter;
      st.samp_w = cr_mat(h, p + "sampling.0.weight", N, c, 1, &perm, nullptr);
    }
    R.stages.push_back(std::move(st));
  };
  for (int i = 0; i < 4; ++i) add_stage("encoders." + std::to_string(i) + ".", i, enc_naf[i], 1);
  add_stage("middle_blocks.", 4, 8, 0);
  for (int j = 0; j < 4; ++j) add_stage("decoders." + std::to_string(j) + ".", 4 - j, 2, 2);
  // workspace
  const size_t cap = R.cap;
  size_t e[5];
  for (int l = 0; l < 5; ++l) e[l] = static_cast<size_t>(kCrRes[l]) * kCrRes[l] * kCrC[l];
  for (int l = 0; l < 5; ++l) R.r[l] = h->arena.get<float>(cap * e[l]);
  for (int l = 1; l < 5; ++l) R.sk[l] = h->arena.get<float>(cap * e[l]);
  R.ln_out = h->arena.get<float>(cap * e[0]);
  R.act_h = h->arena.get<float>(cap * 2 * e[0]);
  R.act_g = h->arena.get<float>(cap * e[0]);
  R.tmp = h->arena.get<float>(cap * e[0]);
  R.pooled = h->arena.get<float>(cap * 512);
  R.sca_s = h->arena.get<float>(cap * 512);
  R.loc1 = h->arena.get<float>(cap * 60 * 60 * 8);
  R.loc2 = h->arena.get<float>(cap * 27 * 27 * 10);
  R.theta = h->arena.get<float>(cap * 6);
  R.stn_hidden = h->arena.get<float>(cap * 96);
  R.stn_ticket = h->arena.get<unsigned int>(cap);   // zero (arena memory is cleared), and every kernel leaves it so
  R.stage = h->arena.get<float>(cap * 3 * R.H * R.H);
  R.use_tc = h->bf16 && h->tun.cr_tc;
  if (R.use_tc) R.a3 = h->arena.get<bf16>(cap * 3 * e[2]);  // levels with c >= 128: rows x c <= e[2]
  R.split_h.clear();
  R.split_hl.clear();
  if (R.use_tc && h->tun.cr_mma3) {  // split weights of the mma.sync GEMMs: here, not at the first forward
    for (const CrStageW& st : R.stages) {
      const int c = st.c;
      if (h->tun.cr_mma3h && (c == 32 || c == 64))
        for (const CrBlockW& b : st.blocks) {
          cr_split_h(h, b.w1, 2 * c, c); cr_split_h(h, b.w3, c, c); cr_split_h(h, b.w4, 2 * c, c); cr_split_h(h, b.w5, c, c);
        }
      if (st.sampling == 1) cr_split_tf32(h, st.samp_w, 2 * c, 4 * c);
      if (st.sampling == 2) cr_split_tf32(h, st.samp_w, 2 * c, c);
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (void* p : h->temp_dev) cudaFree(p);
  h->temp_dev.clear();
  h->src.clear();
  R.loaded = true;
  h->cr_plans.clear();
}

Plan* get_cr_plan(hd_handle* h, int B) {
  auto it = h->cr_plans.find(B);
  if (it != h->cr_plans.end()) return it->second.get();
  std::unique_ptr<Plan> up(new Plan());
  Plan& P = *up;
  P.batch = B;
  const CrW& R = h->cr;
  const int H = R.H;
  auto gemm = [&](int M, int N, int K, const float* A, int lda, const float* W, const float* bias, int epi, float* out, int ldo,
                  const float* resid, int sp, const std::string& label) {
    GemmDesc d;
    d.M = M; d.N = N; d.K = K; d.A = A; d.lda = lda; d.a_dtype = DT_F32; d.W = W; d.ldw = K; d.w_dtype = DT_F32;
    d.bias = bias; d.epi = epi; d.out = out; d.ldo = ldo; d.out_dtype = DT_F32; d.resid = resid; d.ldr = ldo; d.sp = sp;
    if (R.use_tc && h->tun.cr_mma3 && h->tun.cr_mma3h && lda % 4 == 0 && ldo % 2 == 0 && mma3::eligible_h(M, N, K, epi)) {
      // K = 32 / 64: the whole K extent in one stage, row-scaled fp16 split (k16 MMAs)
      const CrW::SplitH& sp = cr_split_h(h, W, N, K);
      mma3::ArgsH a;
      a.A = A; a.w_hi = sp.hi; a.w_lo = sp.lo; a.bias = bias; a.out = out; a.resid = resid;
      a.lda = lda; a.ldo = ldo; a.ldr = ldo; a.M = M; a.N = N; a.w_unscale = sp.unscale;
      g_label = label + fmt(" gemm_mma3h M=%d N=%d K=%d (3 x fp16 split, row-scaled)", M, N, K);
      add_op(P, [a, K, epi](cudaStream_t st) { launch_mma3h(a, K, epi, st); });
      P.flops_per_face += 2.0 * M * static_cast<double>(N) * K / P.batch;
      return;
    }
    if (R.use_tc && h->tun.cr_mma3 && lda % 4 == 0 && ldo % 2 == 0 && mma3::eligible(M, N, K, epi)) {
      // shallow stages: FFMA-bound on CUDA cores, memory-bound on mma.sync with split operands
      const std::pair<float*, float*>& wsp = cr_split_tf32(h, W, N, K);
      mma3::Args a;
      a.A = A; a.w_hi = wsp.first; a.w_lo = wsp.second; a.bias = bias; a.out = out; a.resid = resid;
      a.lda = lda; a.ldo = ldo; a.ldr = ldo; a.M = M; a.N = N; a.K = K; a.sp = sp;
      g_label = label + fmt(" gemm_mma3 M=%d N=%d K=%d (3 x tf32 split)", M, N, K);
      add_op(P, [a, epi](cudaStream_t st) { launch_mma3(a, epi, st); });
      P.flops_per_face += 2.0 * M * static_cast<double>(N) * K / P.batch;
      return;
    }
    g_label = label + fmt(" gemm_ffma M=%d N=%d K=%d", M, N, K);
    add_op(P, [d](cudaStream_t st) { launch_simt(d, st); });
    P.flops_per_face += 2.0 * M * static_cast<double>(N) * K / P.batch;
  };
  auto ew = [&](size_t total, int per_block = 256) { return dim3(static_cast<unsigned>(cdiv(total, static_cast<size_t>(per_block)))); };
  // split-precision tensor-core GEMM (c >= 128): A fp32 -> [hi | lo | hi] bf16, W pre-split [hi | hi | lo], K' = 3K,
  // fp32 accumulate in TMEM: a_hi w_hi + a_lo w_hi + a_hi w_lo
  // A == nullptr: the producer (LayerNorm, SCA scale, SimpleGate) has already written the split operand into R.a3
  auto gemm_tc3 = [&](int M, int N, int K, const float* A, const bf16* Ws, const float* bias, int epi, float* out, int ldo,
                      const float* resid, long long rows_alloc, const std::string& label) {
    bf16* a3 = R.a3;
    const size_t total = static_cast<size_t>(M) * (K / 8);
    if (A != nullptr) {
      g_label = label + " split3";
      add_op(P, [=](cudaStream_t st) { launch_k(cr_split3_kernel, ew(total), dim3(256), 0, st, A, a3, static_cast<size_t>(M), K, 0); });
    }
    GemmDesc d;
    d.M = M; d.N = N; d.K = 3 * K; d.A = a3; d.lda = 3 * K; d.a_dtype = DT_BF16; d.W = Ws; d.ldw = 3 * K; d.w_dtype = DT_BF16;
    d.bias = bias; d.epi = epi; d.out = out; d.ldo = ldo; d.out_dtype = DT_F32; d.resid = resid; d.ldr = ldo;
    g_label = label + " (3xbf16)";
    add_gemm(h, P, d, rows_alloc);
  };
  auto naf_block = [&](const CrBlockW& b, float* x, int n, const std::string& L0) {
    const int c = b.c, rpf = n * n, rows = B * rpf;
    const bool tc = R.use_tc && b.w1s != nullptr;
    const long long rows_alloc = static_cast<long long>(R.cap) * rpf;
    float *ln_out = R.ln_out, *act_h = R.act_h, *act_g = R.act_g, *pooled = R.pooled, *sca_s = R.sca_s;
    ModRef nomod{nullptr, nullptr, 0};
    const float *l1w = b.ln1_w, *l1b = b.ln1_b, *l2w = b.ln2_w, *l2b = b.ln2_b;
    g_label = L0 + "ln1";
    const bool fs = tc && h->tun.cr_fuse_split;   // producers write the split GEMM operand themselves
    bf16* a3 = R.a3;
    if (fs) add_op(P, [=](cudaStream_t st) { launch_ln_split3(c, x, l1w, l1b, a3, rows, rpf, st); });
    else add_op(P, [=](cudaStream_t st) { launch_ln<float>(c, x, l1w, l1b, ln_out, rows, rpf, nomod, 0, 0, 0, st); });
    if (tc) gemm_tc3(rows, 2 * c, c, fs ? nullptr : ln_out, b.w1s, b.b1, EPI_BIAS, act_h, 2 * c, nullptr, rows_alloc, L0 + "conv1");
    else gemm(rows, 2 * c, c, ln_out, c, b.w1, b.b1, EPI_BIAS, act_h, 2 * c, nullptr, 0, L0 + "conv1");
    const float *dw_w = b.dw_w, *dw_b = b.dw_b;
    g_label = L0 + "dwconv+gate";
    add_op(P, [=](cudaStream_t st) {
      if (!h->tun.cr_dw_strip) launch_k(cr_dwconv_gate_kernel, ew(static_cast<size_t>(rows) * (c / 4)), dim3(256), 0, st, static_cast<const float*>(act_h), dw_w, dw_b, act_g, B, n, c);
      else if (n >= 16) launch_k(cr_dwconv_gate_strip_kernel<16>, ew(static_cast<size_t>(rows / 16) * (c / 4), 128), dim3(128), 0, st, static_cast<const float*>(act_h), dw_w, dw_b, act_g, B, n, c);
      else launch_k(cr_dwconv_gate_strip_kernel<8>, ew(static_cast<size_t>(rows / 8) * (c / 4), 128), dim3(128), 0, st, static_cast<const float*>(act_h), dw_w, dw_b, act_g, B, n, c);
    });
    g_label = L0 + "pool";
    add_op(P, [=](cudaStream_t st) {
      launch_k(cr_pool_kernel, dim3(c / 32, B), dim3(1024), 0, st, static_cast<const float*>(act_g), pooled, rpf, c);
    });
    gemm(B, c, c, pooled, c, b.wsca, b.bsca, EPI_BIAS, sca_s, c, nullptr, 0, L0 + "sca");
    g_label = L0 + (fs ? "scale_rows -> split3" : "scale_rows");
    add_op(P, [=](cudaStream_t st) {
      const size_t total8 = static_cast<size_t>(rows) * c / 8;
      if (fs) launch_k(cr_scale_split3_kernel, ew(total8), dim3(256), 0, st, static_cast<const float*>(act_g), static_cast<const float*>(sca_s), a3, total8, c, rpf);
      else launch_k(scale_rows_kernel<float>, ew(total8), dim3(256), 0, st, act_g, static_cast<const float*>(sca_s), total8, c, rpf);
    });
    if (tc) gemm_tc3(rows, c, c, fs ? nullptr : act_g, b.w3s, b.b3, EPI_RESID, x, c, x, rows_alloc, L0 + "conv3");
    else gemm(rows, c, c, act_g, c, b.w3, b.b3, EPI_RESID, x, c, x, 0, L0 + "conv3");
    g_label = L0 + "ln2";
    if (fs) add_op(P, [=](cudaStream_t st) { launch_ln_split3(c, x, l2w, l2b, a3, rows, rpf, st); });
    else add_op(P, [=](cudaStream_t st) { launch_ln<float>(c, x, l2w, l2b, ln_out, rows, rpf, nomod, 0, 0, 0, st); });
    if (tc) gemm_tc3(rows, 2 * c, c, fs ? nullptr : ln_out, b.w4s, b.b4, EPI_BIAS, act_h, 2 * c, nullptr, rows_alloc, L0 + "conv4");
    else gemm(rows, 2 * c, c, ln_out, c, b.w4, b.b4, EPI_BIAS, act_h, 2 * c, nullptr, 0, L0 + "conv4");
    g_label = L0 + "gate";
    add_op(P, [=](cudaStream_t st) {
      if (fs) launch_k(cr_gate_split3_kernel, ew(static_cast<size_t>(rows) * (c / 8)), dim3(256), 0, st, static_cast<const float*>(act_h), a3, static_cast<size_t>(rows), c);
      else launch_k(cr_gate_kernel, ew(static_cast<size_t>(rows) * (c / 4)), dim3(256), 0, st, static_cast<const float*>(act_h), act_g, static_cast<size_t>(rows), c);
    });
    if (tc) gemm_tc3(rows, c, c, fs ? nullptr : act_g, b.w5s, b.b5, EPI_RESID, x, c, x, rows_alloc, L0 + "conv5");
    else gemm(rows, c, c, act_g, c, b.w5, b.b5, EPI_RESID, x, c, x, 0, L0 + "conv5");
    P.flops_per_face += 2.0 * 9 * 2 * c * rpf;
  };
  auto stn = [&](const CrStnW& s, const float* x, float* out, int n, int c, const std::string& L0) {
    float *loc1 = R.loc1, *loc2 = R.loc2, *theta = R.theta, *stn_hidden = R.stn_hidden;
    unsigned int* stn_ticket = R.stn_ticket;
    const float *w1 = s.w1, *b1 = s.b1, *w2 = s.w2, *b2 = s.b2, *f1 = s.f1, *fb1 = s.fb1, *f2 = s.f2, *fb2 = s.fb2;
    const int k1 = s.k1, k2 = s.k2, n1 = s.n1, n2 = s.n2, fc = s.fc, hid = s.hid;
    const __half* w1m = s.w1_mma;
    const float w1u = s.w1_unscale;
    if (R.use_tc && w1m != nullptr) {
      // implicit GEMM on mma.sync with split-precision operands (edge_convs.cuh)
      const int conv_n = n - k1 + 1, tiles = cdiv(conv_n, 16), ch = c == 32 ? 32 : 64;
      const size_t smem = edge::stn_conv_smem(ch, k1);
      static bool configured = false;
      if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(edge::stn_conv_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(edge::stn_conv_smem(32, 9))));
        CUDA_CHECK(cudaFuncSetAttribute(edge::stn_conv_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(edge::stn_conv_smem(64, 9))));
        configured = true;
      }
      g_label = L0 + fmt("stn conv%dx%d+pool+relu mma.sync (3 x fp16 split, scaled)", k1, k1);
      add_op(P, [=](cudaStream_t st) {
        if (ch == 32) launch_k(edge::stn_conv_mma_kernel<32>, dim3(tiles, tiles, B), dim3(256), smem, st, x, w1m, b1, loc1, n, c, k1, n1, w1u);
        else launch_k(edge::stn_conv_mma_kernel<64>, dim3(tiles, tiles, B), dim3(256), smem, st, x, w1m, b1, loc1, n, c, k1, n1, w1u);
      });
    } else {
    g_label = L0 + fmt("stn conv%dx%d+pool+relu", k1, k1);
    add_op(P, [=](cudaStream_t st) {
      if (h->tun.cr_stn_cs) launch_k(cr_stn_conv_pool_cs_kernel<8>, ew(static_cast<size_t>(B) * n1 * n1 * 4, 128), dim3(128), 0, st, x, w1, b1, loc1, B, n, c, k1, n1);
      else launch_k(cr_stn_conv_pool_kernel<8, 2>, ew(static_cast<size_t>(B) * n1 * n1 * 4, 128), dim3(128), 0, st, x, w1, b1, loc1, B, n, c, k1, n1);
    });
    }
    g_label = L0 + fmt("stn conv%dx%d+pool+relu", k2, k2);
    add_op(P, [=](cudaStream_t st) {
      launch_k(cr_stn_conv_pool_kernel<10, 5>, ew(static_cast<size_t>(B) * n2 * n2 * 2, 128), dim3(128), 0, st, static_cast<const float*>(loc1), w2, b2, loc2, B, n1, 8, k2, n2);
    });
    g_label = L0 + "stn fc";
    add_op(P, [=](cudaStream_t st) {
      launch_k(cr_stn_fc_kernel, dim3(cdiv(hid, 8), B), dim3(256), 0, st, static_cast<const float*>(loc2), f1, fb1, f2, fb2, theta, stn_hidden, stn_ticket, fc, hid);
    });
    g_label = L0 + "stn affine_grid+grid_sample";
    add_op(P, [=](cudaStream_t st) {
      launch_k(cr_stn_sample_kernel, ew(static_cast<size_t>(B) * n * n * (c / 4)), dim3(256), 0, st, x, static_cast<const float*>(theta), out, B, n, c);
    });
    P.flops_per_face += 2.0 * (static_cast<double>(k1) * k1 * c * 8 * (2 * n1) * (2 * n1) + static_cast<double>(k2) * k2 * 80 * (2 * n2) * (2 * n2));
  };
  auto copy = [&](const float* src, float* dst, size_t elems, const std::string& label) {
    g_label = label;
    add_op(P, [=](cudaStream_t st) { launch_k(cast_kernel<float>, ew(elems / 8), dim3(256), 0, st, src, dst, elems / 8); });
  };
  {  // intro
    const float *w = R.intro_w, *b = R.intro_b;
    float* out = R.r[0];
    g_label = "cr intro conv3x3";
    add_op(P, [=](cudaStream_t st) { launch_k(cr_intro_kernel, ew(static_cast<size_t>(B) * H * H), dim3(256), 0, st, h->cr_in, w, b, out, B, H); });
  }
  float* tmp = R.tmp;
  for (int i = 0; i < 4; ++i) {  // encoders: NAF blocks, STN, 2x2 stride-2 conv; the result is also the skip (model.py:79-81)
    const CrStageW& S = R.stages[i];
    const int n = S.res, c = S.c;
    const std::string L0 = fmt("cr enc%d c=%d ", i, c);
    for (size_t k = 0; k < S.blocks.size(); ++k) naf_block(S.blocks[k], R.r[i], n, L0 + fmt("b%d ", static_cast<int>(k)));
    stn(S.stn, R.r[i], tmp, n, c, L0);
    float* s2d = R.act_h;
    const int rows_out = B * (n / 2) * (n / 2);
    g_label = L0 + "down s2d";
    add_op(P, [=](cudaStream_t st) {
      const size_t total8 = static_cast<size_t>(rows_out) * 4 * c / 8;
      launch_k(s2d_kernel<float>, ew(total8), dim3(256), 0, st, static_cast<const float*>(tmp), s2d, B, n, c);
    });
    gemm(rows_out, 2 * c, 4 * c, s2d, 4 * c, S.samp_w, S.samp_b, EPI_BIAS, R.r[i + 1], 2 * c, nullptr, 0, L0 + "down");
    copy(R.r[i + 1], R.sk[i + 1], static_cast<size_t>(rows_out) * 2 * c, L0 + "skip copy");
  }
  {  // middle: NAF blocks + STN, no sampling; then x + enc_skips[3] for the first decoder (model.py:82-84)
    const CrStageW& S = R.stages[4];
    const int n = S.res, c = S.c;
    const std::string L0 = fmt("cr mid c=%d ", c);
    for (size_t k = 0; k < S.blocks.size(); ++k) naf_block(S.blocks[k], R.r[4], n, L0 + fmt("b%d ", static_cast<int>(k)));
    stn(S.stn, R.r[4], tmp, n, c, L0);
    float *a = R.sk[4];
    const size_t total4 = static_cast<size_t>(B) * n * n * c / 4;
    g_label = L0 + "add skip";   // sk[4] <- stn(middle) + enc_skips[3]: the first decoder's stream
    add_op(P, [=](cudaStream_t st) { launch_k(cr_add_kernel, ew(total4), dim3(256), 0, st, a, static_cast<const float*>(tmp), total4); });
  }
  for (int j = 0; j < 4; ++j) {  // decoders: stream = (up-sampled previous stage + skip), accumulated in the skip buffer
    const CrStageW& S = R.stages[5 + j];
    const int l = 4 - j, n = S.res, c = S.c;
    const std::string L0 = fmt("cr dec%d c=%d ", j, c);
    float* x = R.sk[l];
    for (size_t k = 0; k < S.blocks.size(); ++k) naf_block(S.blocks[k], x, n, L0 + fmt("b%d ", static_cast<int>(k)));
    stn(S.stn, x, tmp, n, c, L0);
    float* target = l - 1 >= 1 ? R.sk[l - 1] : R.r[0];
    if (l - 1 == 0) {  // the last up-sampling has no skip to land on: start from zeros
      const size_t total4 = static_cast<size_t>(B) * H * H * 32 / 4;
      g_label = L0 + "zero";
      add_op(P, [=](cudaStream_t st) { launch_k(cr_zero_kernel, ew(total4), dim3(256), 0, st, target, total4); });
    }
    gemm(B * n * n, 2 * c, c, tmp, c, S.samp_w, nullptr, EPI_PIXSHUF, target, c / 2, nullptr, n, L0 + "up");
  }
  {  // outro
    const float *w = R.outro_w, *b = R.outro_b;
    const float* in = R.r[0];
    g_label = "cr outro conv3x3";
    add_op(P, [=](cudaStream_t st) { launch_k(cr_outro_kernel, ew(static_cast<size_t>(B) * H * H * 8), dim3(256), 0, st, in, w, b, h->cr_out, B, H); });
  }
  Plan* raw = up.get();
  h->cr_plans[B] = std::move(up);
  return raw;
}

