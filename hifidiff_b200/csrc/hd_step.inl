// hd_step.inl — part of hd_lib.cu (one translation unit; included there in order, not compiled on its own).
// Time-modulation table, per-call step execution (eager / CUDA-graph), stream joins and the device status check.
// ------------------------------------------------------------------------------------------------
// time-modulation table: rows r -> all 32 blocks' [shift_att, scale_att, shift_ffn, scale_ffn]
// (model.py:22-29,46-51 ; conditional_naf.py:18-22,103-106) — fp32 FFMA, depends on t only
// ------------------------------------------------------------------------------------------------
void simt_f32(int M, int N, int K, const float* A, const float* W, const float* bias, float* out, int epi,
              cudaStream_t st) {
  GemmDesc d;
  d.M = M; d.N = N; d.K = K; d.A = A; d.lda = K; d.W = W; d.ldw = K; d.bias = bias; d.epi = epi; d.out = out; d.ldo = N;
  launch_simt(d, st);
}

void compute_time_rows(hd_handle* h, int R) {  // t_vals[0..R) already on device
  cudaStream_t st = h->stream;
  time_embed_kernel<<<cdiv(R * 64, 256), 256, 0, st>>>(h->t_vals, h->freqs, h->t_emb, R);
  simt_f32(R, 2 * kTimeDim, kWidth, h->t_emb, h->tm1_w, h->tm1_b, h->t_h1, EPI_BIAS, st);
  gate_split_kernel<<<cdiv(static_cast<long long>(R) * kTimeDim, 256), 256, 0, st>>>(h->t_h1, h->t_g1, R, kTimeDim);
  simt_f32(R, kTimeDim, kTimeDim, h->t_g1, h->tm3_w, h->tm3_b, h->t_temb, EPI_BIAS, st);
  gate_split_kernel<<<cdiv(static_cast<long long>(R) * 256, 256), 256, 0, st>>>(h->t_temb, h->t_g2, R, 256);
  simt_f32(R, h->mod_stride, 256, h->t_g2, h->mlp_w, h->mlp_b, h->mod_table, EPI_BIAS, st);
  CUDA_CHECK(cudaGetLastError());
}

void ensure_time_table(hd_handle* h, const std::vector<float>& ts) {
  if (h->table_key == ts) return;
  const int R = static_cast<int>(ts.size());
  if (R > h->max_steps) HD_THROW(HD_ERR_INVALID, "%d timesteps exceed the table capacity %d", R, h->max_steps);
  CUDA_CHECK(cudaMemcpyAsync(h->t_vals, ts.data(), R * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));  // ts may be a temporary
  compute_time_rows(h, R);
  h->table_key = ts;
}

void check_device_status(hd_handle* h) {
  DeviceStatus s;
  CUDA_CHECK(cudaMemcpy(&s, h->d_status, sizeof(s), cudaMemcpyDeviceToHost));
  if (s.error != 0) {
    DeviceStatus z{0, 0};
    cudaMemcpy(h->d_status, &z, sizeof(z), cudaMemcpyHostToDevice);
    HD_THROW(HD_ERR_KERNEL, "tcgen05 pipeline watchdog tripped (site 0x%x)", s.where);
  }
}

// Deferred watchdog report for the asynchronous entry points: every call that enqueues kernels ends with an async
// copy of the 8-byte status word into pinned memory (post_status); the NEXT call on the handle, and hd_synchronize,
// look at it once the copy has completed (poll_status).  A tripped tcgen05 pipeline therefore surfaces as
// HD_ERR_KERNEL on the following call instead of staying silent until somebody calls hd_synchronize.
void post_status(hd_handle* h) {
  cudaMemcpyAsync(h->status_host, h->d_status, sizeof(DeviceStatus), cudaMemcpyDeviceToHost, h->stream);
  cudaEventRecord(h->ev_status, h->stream);
  h->status_posted = true;
}
void poll_status(hd_handle* h) {
  if (!h->status_posted || cudaEventQuery(h->ev_status) != cudaSuccess) { cudaGetLastError(); return; }
  h->status_posted = false;
  if (h->status_host->error != 0) {
    const unsigned int where = h->status_host->where;
    DeviceStatus z{0, 0};
    *h->status_host = z;
    cudaMemcpy(h->d_status, &z, sizeof(z), cudaMemcpyHostToDevice);
    HD_THROW(HD_ERR_KERNEL, "tcgen05 pipeline watchdog tripped in an earlier call (site 0x%x)", where);
  }
}

void join_in(hd_handle* h, void* user_stream) {
  poll_status(h);
  t_use_pdl = h->tun.pdl;
  cudaStream_t us = static_cast<cudaStream_t>(user_stream);
  CUDA_CHECK(cudaEventRecord(h->ev_in, us));
  CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->ev_in, 0));
}
void join_out(hd_handle* h, void* user_stream) {
  cudaStream_t us = static_cast<cudaStream_t>(user_stream);
  post_status(h);
  CUDA_CHECK(cudaEventRecord(h->ev_out, h->stream));
  CUDA_CHECK(cudaStreamWaitEvent(us, h->ev_out, 0));
}

void run_plan(hd_handle* h, Plan* P, cudaStream_t st, const char* const* tap_names, float* const* tap_out, int n_taps,
              int B) {
  for (auto& op : P->ops) {
    op.fn(st);
    if (n_taps > 0 && !op.tap.empty()) {
      for (int i = 0; i < n_taps; ++i) {
        if (op.tap != tap_names[i]) continue;
        const TapInfo& ti = op.info;
        const size_t total = static_cast<size_t>(B) * ti.C * ti.HW;
        float* dst = tap_out[i];
        float* dev_dst = dst;
        const bool host_dst = !is_device_ptr(dst);
        if (host_dst) CUDA_CHECK(cudaMalloc(&dev_dst, total * 4));
        if (ti.dtype == DT_BF16)
          nhwc_to_nchw_kernel<bf16><<<cdiv(total, 256), 256, 0, st>>>(static_cast<const bf16*>(ti.ptr), dev_dst, B, ti.C, ti.HW, ti.ld);
        else
          nhwc_to_nchw_kernel<float><<<cdiv(total, 256), 256, 0, st>>>(static_cast<const float*>(ti.ptr), dev_dst, B, ti.C, ti.HW, ti.ld);
        if (host_dst) {
          CUDA_CHECK(cudaStreamSynchronize(st));
          CUDA_CHECK(cudaMemcpy(dst, dev_dst, total * 4, cudaMemcpyDeviceToHost));
          cudaFree(dev_dst);
        }
      }
    }
  }
  CUDA_CHECK(cudaGetLastError());
}

void denoise_impl(hd_handle* h, const float* x, const float* t, int t_len, float* eps_out, int B,
                  const char* const* tap_names, float* const* tap_out, int n_taps, void* user_stream) {
  if (!h->weights_loaded) HD_THROW(HD_ERR_STATE, "hd_load_weights has not been called");
  if (h->fused && !h->condition_set) HD_THROW(HD_ERR_STATE, "hd_set_condition has not been called");
  if (B < 1 || B > h->cfg.max_batch) HD_THROW(HD_ERR_INVALID, "batch %d outside [1, %d]", B, h->cfg.max_batch);
  if (t_len != 1 && t_len != B) HD_THROW(HD_ERR_INVALID, "t_len must be 1 or batch");
  if (t_len > h->max_steps) HD_THROW(HD_ERR_INVALID, "per-face timesteps (%d) exceed table rows (%d)", t_len, h->max_steps);
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  join_in(h, user_stream);
  cudaStream_t st = h->stream;
  const size_t xe = static_cast<size_t>(B) * 4 * h->S * h->S;
  const bool x_dev = is_device_ptr(x), e_dev = is_device_ptr(eps_out);
  if (!x_dev) CUDA_CHECK(cudaMemcpyAsync(h->x_stage, x, xe * 4, cudaMemcpyHostToDevice, st));
  h->cur_x = x_dev ? x : h->x_stage;
  h->cur_eps = e_dev ? eps_out : h->eps_buf;
  // time rows: row r of the table <- t[r]
  std::vector<float> ts(t_len);
  CUDA_CHECK(cudaMemcpyAsync(ts.data(), t, t_len * sizeof(float), cudaMemcpyDefault, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  ensure_time_table(h, ts);
  std::vector<int> rows(B);
  for (int b = 0; b < B; ++b) rows[b] = t_len == 1 ? 0 : b;
  CUDA_CHECK(cudaMemcpyAsync(h->row_idx, rows.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  // taps come from the production plan when it exposes every requested one, else from the per-op plan
  Plan* P = get_plan(h, B, false);
  for (int i = 0; i < n_taps; ++i) {
    bool found = false;
    for (auto& op : P->ops) found = found || op.tap == tap_names[i];
    if (!found) { P = get_plan(h, B, true); break; }
  }
  run_plan(h, P, st, tap_names, tap_out, n_taps, B);
  // "time_mlp" tap: (B,512) embedding
  for (int i = 0; i < n_taps; ++i) {
    if (std::string(tap_names[i]) != "time_mlp") continue;
    for (int b = 0; b < B; ++b)
      CUDA_CHECK(cudaMemcpyAsync(tap_out[i] + static_cast<size_t>(b) * kTimeDim,
                                 h->t_temb + static_cast<size_t>(t_len == 1 ? 0 : b) * kTimeDim, kTimeDim * 4,
                                 cudaMemcpyDefault, st));
  }
  if (!e_dev) {
    CUDA_CHECK(cudaMemcpyAsync(eps_out, h->eps_buf, xe * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  }
  join_out(h, user_stream);
  if (!e_dev || n_taps > 0) check_device_status(h);
}

}  // namespace
