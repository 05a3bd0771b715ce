// hd_weights.inl — part of hd_lib.cu (one translation unit; included there in order, not compiled on its own).
// Denoiser weight loading: name lookup in the caller's state_dict and repacking into the library-owned arena.
// ------------------------------------------------------------------------------------------------
// weight loading
// ------------------------------------------------------------------------------------------------
const SrcTensor& need(hd_handle* h, const std::string& name, std::initializer_list<int64_t> shape) {
  auto it = h->src.find(name);
  if (it == h->src.end()) HD_THROW(HD_ERR_INVALID, "missing tensor '%s'", name.c_str());
  const SrcTensor& t = it->second;
  size_t n = 1;
  for (int64_t s : shape) n *= static_cast<size_t>(s);
  if (t.dtype != 0) HD_THROW(HD_ERR_INVALID, "tensor '%s' must be fp32", name.c_str());
  if (t.numel != n) HD_THROW(HD_ERR_INVALID, "tensor '%s' has %zu elements, expected %zu", name.c_str(), t.numel, n);
  return t;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// device view of a source tensor (uploads host memory into a temporary)
const float* dev_src(hd_handle* h, const SrcTensor& t) {
  if (is_device_ptr(t.data)) return static_cast<const float*>(t.data);
  void* p = nullptr;
  CUDA_CHECK(cudaMalloc(&p, t.numel * 4));
  h->temp_dev.push_back(p);
  CUDA_CHECK(cudaMemcpy(p, t.data, t.numel * 4, cudaMemcpyHostToDevice));
  return static_cast<const float*>(p);
}

std::vector<float> host_vec(hd_handle* h, const SrcTensor& t) {
  std::vector<float> v(t.numel);
  CUDA_CHECK(cudaMemcpy(v.data(), t.data, t.numel * 4, cudaMemcpyDefault));
  return v;
}

float* upload_f32(hd_handle* h, const std::vector<float>& v) {
  float* d = h->arena.get<float>(v.size());
  CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return d;
}
int* upload_i32(hd_handle* h, const std::vector<int>& v) {
  int* d = h->arena.get<int>(v.size());
  CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return d;
}

// dst[n, kd] = rs[n] * src[perm[n]][kmap(kd)]  into a freshly allocated arena matrix of dtype dt
void* pack_matrix(hd_handle* h, const SrcTensor& t, int N, int Kd, int taps, const std::vector<int>* perm,
                  const std::vector<float>* rs, int dt, void* dst_override = nullptr) {
  const float* src = dev_src(h, t);
  void* dst = dst_override ? dst_override : h->arena.alloc(static_cast<size_t>(N) * Kd * esize(dt));
  const int* dperm = perm ? upload_i32(h, *perm) : nullptr;
  const float* drs = rs ? upload_f32(h, *rs) : nullptr;
  const size_t total = static_cast<size_t>(N) * Kd;
  const int blocks = cdiv(total, 256);
  if (dt == DT_BF16)
    pack_rows_kernel<bf16><<<blocks, 256, 0, h->stream>>>(src, static_cast<bf16*>(dst), dperm, drs, N, Kd, taps);
  else
    pack_rows_kernel<float><<<blocks, 256, 0, h->stream>>>(src, static_cast<float*>(dst), dperm, drs, N, Kd, taps);
  CUDA_CHECK(cudaGetLastError());
  return dst;
}

std::vector<float> bn_scale(hd_handle* h, const std::string& p, int n, std::vector<float>* shift_out,
                            const std::vector<float>& conv_bias) {
  // eval-mode BatchNorm folded into the preceding conv: y = rs * (conv + b - mean) + beta
  auto w = host_vec(h, need(h, p + "weight", {n}));
  auto b = host_vec(h, need(h, p + "bias", {n}));
  auto mu = host_vec(h, need(h, p + "running_mean", {n}));
  auto var = host_vec(h, need(h, p + "running_var", {n}));
  std::vector<float> rs(n);
  shift_out->resize(n);
  for (int i = 0; i < n; ++i) {
    rs[i] = w[i] / std::sqrt(var[i] + kBnEps);
    (*shift_out)[i] = (conv_bias[i] - mu[i]) * rs[i] + b[i];
  }
  return rs;
}

// intro.weight [128, 4*9] -> [36][128] (tap-major, output channel minor): the kernel's shared-memory layout, so the
// per-block fill is a straight coalesced copy
std::vector<float> intro_taps_major(const std::vector<float>& w) {
  std::vector<float> t(w.size());
  for (int o = 0; o < kWidth; ++o)
    for (int k = 0; k < 36; ++k) t[static_cast<size_t>(k) * kWidth + o] = w[static_cast<size_t>(o) * 36 + k];
  return t;
}

void load_block(hd_handle* h, BlockW& bw, int wdt) {
  const std::string& p = bw.prefix;
  const int c = bw.c;
  bw.ln1_w = upload_f32(h, host_vec(h, need(h, p + "norm1.weight", {c})));
  bw.ln1_b = upload_f32(h, host_vec(h, need(h, p + "norm1.bias", {c})));
  bw.ln2_w = upload_f32(h, host_vec(h, need(h, p + "norm2.weight", {c})));
  bw.ln2_b = upload_f32(h, host_vec(h, need(h, p + "norm2.bias", {c})));
  auto beta = host_vec(h, need(h, p + "beta", {c}));
  auto gamma = host_vec(h, need(h, p + "gamma", {c}));

  if (h->sp[bw.level] == 1) {
    // At 1x1 spatial only the centre tap of the depthwise 3x3 sees a pixel (zero padding), so
    // conv2(conv1(x)) = dwc * (W1 x + b1) + bdw per channel: fold it into conv1 and let the
    // SimpleGate run in conv1's epilogue (same 128-row [64 x1 | 64 x2] packing as conv4).
    auto dw = host_vec(h, need(h, p + "conv2.weight", {2 * c, 9}));
    auto dwb = host_vec(h, need(h, p + "conv2.bias", {2 * c}));
    auto b1 = host_vec(h, need(h, p + "conv1.bias", {2 * c}));
    std::vector<int> perm(2 * c);
    std::vector<float> rs(2 * c), bp(2 * c);
    for (int n = 0; n < 2 * c; ++n) {
      const int g = n / 128, r = n % 128;
      const int ch = r < 64 ? g * 64 + r : c + g * 64 + (r - 64);
      perm[n] = ch;
      rs[n] = dw[static_cast<size_t>(ch) * 9 + 4];
      bp[n] = rs[n] * b1[ch] + dwb[ch];
    }
    bw.w1 = pack_matrix(h, need(h, p + "conv1.weight", {2 * c, c}), 2 * c, c, 1, &perm, &rs, wdt);
    bw.b1 = upload_f32(h, bp);
    bw.dw_folded = true;
  } else {
    bw.w1 = pack_matrix(h, need(h, p + "conv1.weight", {2 * c, c}), 2 * c, c, 1, nullptr, nullptr, wdt);
    bw.b1 = upload_f32(h, host_vec(h, need(h, p + "conv1.bias", {2 * c})));
  }

  {  // depthwise 3x3: [2c,1,3,3] -> [9][2c]
    auto w = host_vec(h, need(h, p + "conv2.weight", {2 * c, 9}));
    std::vector<float> t(static_cast<size_t>(18) * c);
    for (int ch = 0; ch < 2 * c; ++ch)
      for (int tap = 0; tap < 9; ++tap) t[static_cast<size_t>(tap) * 2 * c + ch] = w[static_cast<size_t>(ch) * 9 + tap];
    bw.dw_w = upload_f32(h, t);
    bw.dw_b = upload_f32(h, host_vec(h, need(h, p + "conv2.bias", {2 * c})));
  }
  bw.wsca = pack_matrix(h, need(h, p + "sca.1.weight", {c, c}), c, c, 1, nullptr, nullptr, wdt);
  bw.bsca = upload_f32(h, host_vec(h, need(h, p + "sca.1.bias", {c})));
  if (c == fb::C && h->sp[bw.level] == fb::SP) {
    auto w = host_vec(h, need(h, p + "sca.1.weight", {c, c}));
    std::vector<float> t(w.size());
    for (int n = 0; n < c; ++n)
      for (int k = 0; k < c; ++k) t[static_cast<size_t>(k) * c + n] = w[static_cast<size_t>(n) * c + k];
    bw.wsca_t = upload_f32(h, t);
  }
  if (c == pb::C && h->sp[bw.level] == pb::SP) {
    auto w = host_vec(h, need(h, p + "sca.1.weight", {c, c}));
    std::vector<uint16_t> t(w.size());
    for (int n = 0; n < c; ++n)
      for (int k = 0; k < c; ++k) {
        uint32_t u;
        memcpy(&u, &w[static_cast<size_t>(n) * c + k], 4);
        u += 0x7FFFu + ((u >> 16) & 1u);  // round to nearest even (weights are finite)
        t[static_cast<size_t>(k) * c + n] = static_cast<uint16_t>(u >> 16);
      }
    bw.wsca_tb = h->arena.alloc(t.size() * 2);
    CUDA_CHECK(cudaMemcpy(bw.wsca_tb, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
  }

  {  // conv3 with beta folded: y = inp + beta * (W3 x + b3)
    auto b3 = host_vec(h, need(h, p + "conv3.bias", {c}));
    for (int i = 0; i < c; ++i) b3[i] *= beta[i];
    bw.w3 = pack_matrix(h, need(h, p + "conv3.weight", {c, c}), c, c, 1, nullptr, &beta, wdt);
    bw.b3 = upload_f32(h, b3);
    bw.b3_h = b3;
  }
  {  // conv4, gate-packed: 128-row groups [64 x1 rows | 64 matching x2 rows]
    std::vector<int> perm(2 * c);
    for (int n = 0; n < 2 * c; ++n) {
      const int g = n / 128, r = n % 128;
      perm[n] = r < 64 ? g * 64 + r : c + g * 64 + (r - 64);
    }
    auto b4 = host_vec(h, need(h, p + "conv4.bias", {2 * c}));
    std::vector<float> b4p(2 * c);
    for (int n = 0; n < 2 * c; ++n) b4p[n] = b4[perm[n]];
    bw.w4 = pack_matrix(h, need(h, p + "conv4.weight", {2 * c, c}), 2 * c, c, 1, &perm, nullptr, wdt);
    bw.b4 = upload_f32(h, b4p);
  }
  {  // conv5 with gamma folded
    auto b5 = host_vec(h, need(h, p + "conv5.bias", {c}));
    for (int i = 0; i < c; ++i) b5[i] *= gamma[i];
    bw.w5 = pack_matrix(h, need(h, p + "conv5.weight", {c, c}), c, c, 1, nullptr, &gamma, wdt);
    bw.b5 = upload_f32(h, b5);
    bw.b5_h = b5;
  }
  if (bw.has_mod) {
    // per-block time MLP rows go into the concatenated [mod_stride, 256] matrix
    pack_matrix(h, need(h, p + "mlp.1.weight", {4 * c, 256}), 4 * c, 256, 1, nullptr, nullptr, DT_F32,
                h->mlp_w + static_cast<size_t>(bw.mod_off) * 256);
    const SrcTensor& t = need(h, p + "mlp.1.bias", {4 * c});
    CUDA_CHECK(cudaMemcpy(h->mlp_b + bw.mod_off, t.data, static_cast<size_t>(4) * c * 4, cudaMemcpyDefault));
    h->weight_elems_step += static_cast<int64_t>(c) * c * 7 + 18 * c;
  }
}

void load_weights_impl(hd_handle* h) {
  const int wdt = h->bf16 ? DT_BF16 : DT_F32;
  h->weight_elems_step = 0;
  h->mlp_w = h->arena.get<float>(static_cast<size_t>(h->mod_stride) * 256);
  h->mlp_b = h->arena.get<float>(h->mod_stride);
  for (auto& b : h->blocks) load_block(h, b, wdt);

  h->tm1_w = upload_f32(h, host_vec(h, need(h, "time_mlp.1.weight", {2 * kTimeDim, kWidth})));
  h->tm1_b = upload_f32(h, host_vec(h, need(h, "time_mlp.1.bias", {2 * kTimeDim})));
  h->tm3_w = upload_f32(h, host_vec(h, need(h, "time_mlp.3.weight", {kTimeDim, kTimeDim})));
  h->tm3_b = upload_f32(h, host_vec(h, need(h, "time_mlp.3.bias", {kTimeDim})));
  h->intro_w = upload_f32(h, intro_taps_major(host_vec(h, need(h, "intro.weight", {kWidth, 36}))));
  h->intro_b = upload_f32(h, host_vec(h, need(h, "intro.bias", {kWidth})));
  h->end_w = static_cast<float*>(pack_matrix(h, need(h, "ending.weight", {4, kWidth, 9}), 4, 9 * kWidth, 9, nullptr,
                                             nullptr, DT_F32));
  h->end_b = upload_f32(h, host_vec(h, need(h, "ending.bias", {4})));
  h->weight_elems_step += 128 * 36 + 4 * 9 * 128;
  if (h->bf16 && h->S == edge::S) {
    // intro / ending weights as bf16 hi + lo in mma.sync B-fragment order [k-step][n][16 k] (edge_convs.cuh)
    auto split = [&](const std::vector<float>& v, hd::bf16** hi, hd::bf16** lo) {
      std::vector<uint16_t> vh(v.size()), vl(v.size());
      auto to_bf16 = [](float f) {
        uint32_t u;
        memcpy(&u, &f, 4);
        u += 0x7FFFu + ((u >> 16) & 1u);
        return static_cast<uint16_t>(u >> 16);
      };
      for (size_t i = 0; i < v.size(); ++i) {
        vh[i] = to_bf16(v[i]);
        const uint32_t hb = static_cast<uint32_t>(vh[i]) << 16;
        float hf;
        memcpy(&hf, &hb, 4);
        vl[i] = to_bf16(v[i] - hf);
      }
      *hi = static_cast<hd::bf16*>(h->arena.alloc(v.size() * 2));
      *lo = static_cast<hd::bf16*>(h->arena.alloc(v.size() * 2));
      CUDA_CHECK(cudaMemcpy(*hi, vh.data(), v.size() * 2, cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaMemcpy(*lo, vl.data(), v.size() * 2, cudaMemcpyHostToDevice));
    };
    {
      auto w = host_vec(h, need(h, "ending.weight", {4, kWidth, 9}));  // [o][c][tap]
      std::vector<float> f(static_cast<size_t>(edge::END_KSTEPS) * 8 * 16, 0.f);
      for (int tap = 0; tap < 9; ++tap)
        for (int cc = 0; cc < 8; ++cc)
          for (int n = 0; n < 4; ++n)
            for (int k = 0; k < 16; ++k)
              f[((static_cast<size_t>(tap) * 8 + cc) * 8 + n) * 16 + k] = w[(static_cast<size_t>(n) * kWidth + cc * 16 + k) * 9 + tap];
      split(f, &h->end_mma_hi, &h->end_mma_lo);
    }
    {
      auto w = host_vec(h, need(h, "intro.weight", {kWidth, 36}));  // [o][ci * 9 + tap]
      std::vector<float> f(static_cast<size_t>(3) * 16 * 8 * 16, 0.f);
      for (int ks = 0; ks < 3; ++ks)
        for (int nt = 0; nt < 16; ++nt)
          for (int n = 0; n < 8; ++n)
            for (int k = 0; k < 16; ++k) {
              const int kk = ks * 16 + k;
              if (kk < 36) f[((static_cast<size_t>(ks) * 16 + nt) * 8 + n) * 16 + k] = w[static_cast<size_t>(nt * 8 + n) * 36 + kk];
            }
      split(f, &h->intro_mma_hi, &h->intro_mma_lo);
    }
    h->end_ticket = h->arena.get<unsigned int>(64);
    CUDA_CHECK(cudaFuncSetAttribute(edge::ending_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, edge::END_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(edge::ending_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, edge::END_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(edge::intro_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, edge::IN_SMEM));
  }

  for (int l = 0; l < 4; ++l) {
    const int c = h->c[l];
    const std::string p = "downs." + std::to_string(l) + ".";
    h->down_w[l] = pack_matrix(h, need(h, p + "weight", {2 * c, c, 4}), 2 * c, 4 * c, 4, nullptr, nullptr, wdt);
    h->down_b[l] = upload_f32(h, host_vec(h, need(h, p + "bias", {2 * c})));
    h->weight_elems_step += static_cast<int64_t>(8) * c * c;
  }
  for (int L = 0; L < 4; ++L) {
    const int cin = h->c[4 - L];  // 2048, 1024, 512, 256
    const int N = 2 * cin, quarter = N / 4;
    std::vector<int> perm(N);
    for (int n = 0; n < N; ++n) perm[n] = 4 * (n % quarter) + n / quarter;  // packed row q*quarter+k <- 4k+q
    h->up_w[L] = pack_matrix(h, need(h, "ups." + std::to_string(L) + ".0.weight", {N, cin}), N, cin, 1, &perm,
                             nullptr, wdt);
    h->weight_elems_step += static_cast<int64_t>(N) * cin;
  }
  if (h->fused) {
    const int idc_out = 2048 * (h->S / 16) * (h->S / 16);
    h->idc_w = upload_f32(h, host_vec(h, need(h, "idc_conv.weight", {idc_out, 2048})));
    h->idc_b = upload_f32(h, host_vec(h, need(h, "idc_conv.bias", {idc_out})));
    for (int j = 0; j < kNumLevels; ++j) {
      HcaW& w = h->hca[j];
      const int d = w.d;
      const std::string p = "hcas." + std::to_string(j) + ".";
      w.c0w = upload_f32(h, host_vec(h, need(h, p + "channel_mlp.0.weight", {d, d})));
      w.c0b = upload_f32(h, host_vec(h, need(h, p + "channel_mlp.0.bias", {d})));
      w.c2w = upload_f32(h, host_vec(h, need(h, p + "channel_mlp.2.weight", {d, d})));
      w.c2b = upload_f32(h, host_vec(h, need(h, p + "channel_mlp.2.bias", {d})));
      {
        std::vector<float> shift;
        auto cb = host_vec(h, need(h, p + "spatial_mlp.0.bias", {d / 2}));
        auto rs = bn_scale(h, p + "spatial_mlp.1.", d / 2, &shift, cb);
        w.s0w = static_cast<float*>(pack_matrix(h, need(h, p + "spatial_mlp.0.weight", {d / 2, d}), d / 2, d, 1,
                                                nullptr, &rs, DT_F32));
        w.s0b = upload_f32(h, shift);
      }
      {
        std::vector<float> shift;
        auto cb = host_vec(h, need(h, p + "spatial_mlp.3.bias", {1}));
        auto rs = bn_scale(h, p + "spatial_mlp.4.", 1, &shift, cb);
        w.s3w = static_cast<float*>(pack_matrix(h, need(h, p + "spatial_mlp.3.weight", {1, d / 2}), 1, d / 2, 1,
                                                nullptr, &rs, DT_F32));
        w.s3b = upload_f32(h, shift);
      }
      {
        std::vector<float> shift;
        auto cb = host_vec(h, need(h, p + "fused_mlp.0.bias", {d}));
        auto rs = bn_scale(h, p + "fused_mlp.1.", d, &shift, cb);
        w.wf = pack_matrix(h, need(h, p + "fused_mlp.0.weight", {d, d, 9}), d, 9 * d, 9, nullptr, &rs, wdt);
        w.bf = upload_f32(h, shift);
      }
      // taps that can touch a real pixel: all 9 unless the level is 1x1 (centre tap only)
      h->weight_elems_step += static_cast<int64_t>(d) * d * (w.sp == 1 ? 1 : 9);
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (void* p : h->temp_dev) cudaFree(p);
  h->temp_dev.clear();
  h->src.clear();
  h->weights_loaded = true;
  h->table_key.clear();
}

