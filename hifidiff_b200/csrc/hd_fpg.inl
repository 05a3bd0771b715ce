// hd_fpg.inl — part of hd_lib.cu (one translation unit; included there in order, not compiled on its own).
// FacialPriorGuidance: weights and forward plan.
// ------------------------------------------------------------------------------------------------
// FacialPriorGuidance: the same NAF-block kernels without modulation
// ------------------------------------------------------------------------------------------------
void load_fpg_impl(hd_handle* h) {
  const int wdt = h->bf16 ? DT_BF16 : DT_F32;
  FpgW& F = h->fpg;
  F.blocks.clear();
  for (int l = 0; l < 4; ++l)
    for (int i = 0; i < kEncBlocks[l]; ++i) {
      BlockW b;
      b.prefix = "encoders." + std::to_string(l) + "." + std::to_string(i) + ".";
      b.level = l; b.c = h->c[l]; b.mod_off = 0; b.has_mod = false;
      F.blocks.push_back(b);
    }
  const int64_t keep = h->weight_elems_step;
  for (auto& b : F.blocks) load_block(h, b, wdt);
  F.intro_w = upload_f32(h, intro_taps_major(host_vec(h, need(h, "intro.weight", {kWidth, 36}))));
  F.intro_b = upload_f32(h, host_vec(h, need(h, "intro.bias", {kWidth})));
  for (int l = 0; l < 4; ++l) {
    const int c = h->c[l];
    const std::string p = "downs." + std::to_string(l) + ".";
    F.down_w[l] = pack_matrix(h, need(h, p + "weight", {2 * c, c, 4}), 2 * c, 4 * c, 4, nullptr, nullptr, wdt);
    F.down_b[l] = upload_f32(h, host_vec(h, need(h, p + "bias", {2 * c})));
  }
  F.convs_w[0] = pack_matrix(h, need(h, "convs.0.0.weight", {2048, 2048}), 2048, 2048, 1, nullptr, nullptr, wdt);
  for (int j = 1; j < 5; ++j) {
    const int cin = h->c[5 - j];  // 2048, 1024, 512, 256
    const int N = 2 * cin, quarter = N / 4;
    std::vector<int> perm(N);
    for (int n = 0; n < N; ++n) perm[n] = 4 * (n % quarter) + n / quarter;
    F.convs_w[j] = pack_matrix(h, need(h, "convs." + std::to_string(j) + ".0.weight", {N, cin}), N, cin, 1, &perm, nullptr, wdt);
  }
  F.zero_bias = h->arena.get<float>(4096);  // arena memory is zero-initialised
  F.p0 = h->arena.get<float>(static_cast<size_t>(h->Bcap) * 2048);
  h->weight_elems_step = keep;
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (void* p : h->temp_dev) cudaFree(p);
  h->temp_dev.clear();
  h->src.clear();
  F.loaded = true;
  h->fpg_plans.clear();
}

Plan* get_fpg_plan(hd_handle* h, int B) {
  auto it = h->fpg_plans.find(B);
  if (it != h->fpg_plans.end()) return it->second.get();
  std::unique_ptr<Plan> up(new Plan());
  Plan& P = *up;
  P.batch = B;
  const int S = h->S;
  const bool bf = h->bf16;
  const int adt = bf ? DT_BF16 : DT_F32;
  const FpgW& F = h->fpg;
  {
    float* out = h->resid[0];
    const float *w = F.intro_w, *b = F.intro_b;
    g_label = "fpg intro conv3x3";
    add_op(P, [=](cudaStream_t st) {
      launch_k(intro_conv_kernel, dim3(B), dim3(256), (36 * 128 + 4 * (S + 2) * (S + 2)) * sizeof(float), st, h->fpg_in, w, b, out, S);
    });
  }
  size_t bi = 0;
  for (int l = 0; l < 4; ++l) {
    for (int i = 0; i < kEncBlocks[l]; ++i, ++bi) add_block(h, P, F.blocks[bi], std::string());
    const int c = h->c[l], n = h->sp[l], rows_out = B * (n / 2) * (n / 2);
    const float* src = h->resid[l];
    void* act_a = h->act_a;
    g_label = fmt("fpg down%d s2d", l);
    add_op(P, [=](cudaStream_t st) {
      const size_t total8 = static_cast<size_t>(rows_out) * 4 * c / 8;
      if (bf) launch_k(s2d_kernel<bf16>, dim3(cdiv(total8, 256)), dim3(256), 0, st, src, static_cast<bf16*>(act_a), B, n, c);
      else launch_k(s2d_kernel<float>, dim3(cdiv(total8, 256)), dim3(256), 0, st, src, static_cast<float*>(act_a), B, n, c);
    });
    GemmDesc d;
    d.M = rows_out; d.N = 2 * c; d.K = 4 * c; d.A = act_a; d.lda = 4 * c; d.a_dtype = adt;
    d.W = F.down_w[l]; d.ldw = 4 * c; d.w_dtype = adt; d.bias = F.down_b[l]; d.epi = EPI_BIAS;
    d.out = h->resid[l + 1]; d.ldo = 2 * c; d.out_dtype = DT_F32;
    g_label = fmt("fpg down%d", l);
    add_gemm(h, P, d, static_cast<long long>(h->Bcap) * (n / 2) * (n / 2));
  }
  auto cast_to_act = [&](const float* src, size_t elems) {
    void* act_a = h->act_a;
    g_label = "fpg cast";
    add_op(P, [=](cudaStream_t st) {
      const size_t total8 = elems / 8;
      if (bf) launch_k(cast_kernel<bf16>, dim3(cdiv(total8, 256)), dim3(256), 0, st, src, static_cast<bf16*>(act_a), total8);
      else launch_k(cast_kernel<float>, dim3(cdiv(total8, 256)), dim3(256), 0, st, src, static_cast<float*>(act_a), total8);
    });
  };
  {  // convs[0]: 1x1 2048 -> 2048, no bias, PixelShuffle(1) == identity  (fpg/model.py:34-36,56-57)
    const int n = h->sp[4], rows = B * n * n;
    cast_to_act(h->resid[4], static_cast<size_t>(rows) * 2048);
    GemmDesc d;
    d.M = rows; d.N = 2048; d.K = 2048; d.A = h->act_a; d.lda = 2048; d.a_dtype = adt;
    d.W = F.convs_w[0]; d.ldw = 2048; d.w_dtype = adt; d.bias = F.zero_bias; d.epi = EPI_BIAS;
    d.out = F.p0; d.ldo = 2048; d.out_dtype = DT_F32;
    g_label = "fpg convs0";
    add_gemm(h, P, d, static_cast<long long>(h->Bcap) * n * n);
  }
  for (int j = 1; j < 5; ++j) {  // x = PixelShuffle(conv(x)) + skip, accumulated in place on the skip buffer
    const int lin = 5 - j, lout = 4 - j;
    const int cin = h->c[lin], n = h->sp[lin], rows_in = B * n * n;
    cast_to_act(j == 1 ? F.p0 : h->resid[lin], static_cast<size_t>(rows_in) * cin);
    GemmDesc d;
    d.M = rows_in; d.N = 2 * cin; d.K = cin; d.A = h->act_a; d.lda = cin; d.a_dtype = adt;
    d.W = F.convs_w[j]; d.ldw = cin; d.w_dtype = adt; d.bias = nullptr; d.epi = EPI_PIXSHUF; d.sp = n;
    d.out = h->resid[lout]; d.ldo = cin / 2; d.out_dtype = DT_F32;
    g_label = fmt("fpg convs%d", j);
    add_gemm(h, P, d, static_cast<long long>(h->Bcap) * n * n);
  }
  Plan* raw = up.get();
  h->fpg_plans[B] = std::move(up);
  return raw;
}

