// hifidiff_b200 host side: handle, weight repacking, per-batch launch plan, C ABI (include/hifidiff_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <unordered_map>

#include "../../include/hifidiff_b200.h"
#include "common.cuh"
#include "elem_kernels.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "edge_convs.cuh"
#include "gemm_mma3.cuh"
#include "face_block.cuh"
#include "pair_block.cuh"
#include "idc_kernels.cuh"
#include "cr_kernels.cuh"

using namespace hd;

#include "hd_state.inl"
#include "hd_gemm_dispatch.inl"
#include "hd_weights.inl"
#include "hd_plan.inl"
#include "hd_fpg.inl"
#include "hd_idc.inl"
#include "hd_cr.inl"
#include "hd_step.inl"

// ================================================================================================
// C ABI
// ================================================================================================
#define HD_API_BEGIN try {
#define HD_API_END(h)                                      \
  }                                                        \
  catch (const HdError& e) {                               \
    if (h) (h)->err = e.msg; else g_create_error = e.msg;  \
    return e.code;                                         \
  }                                                        \
  catch (const std::exception& e) {                        \
    if (h) (h)->err = e.what(); else g_create_error = e.what(); \
    return HD_ERR_INVALID;                                 \
  }                                                        \
  return HD_OK;

extern "C" {

int32_t hd_abi_version(void) { return HD_ABI_VERSION; }

const char* hd_last_error(const hd_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void hd_destroy(hd_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (auto& kv : h->plans)
    if (kv.second->graph) cudaGraphExecDestroy(kv.second->graph);
  h->plans.clear();
  h->plans_dbg.clear();
  h->arena.release();
  if (h->ev_in) cudaEventDestroy(h->ev_in);
  if (h->ev_out) cudaEventDestroy(h->ev_out);
  if (h->ev_status) cudaEventDestroy(h->ev_status);
  if (h->status_host) cudaFreeHost(h->status_host);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int32_t hd_create(hd_handle** out, const hd_config* cfg) {
  hd_handle* h = nullptr;
  if (out) *out = nullptr;
  hd_handle* null_handle = nullptr;
  HD_API_BEGIN
  if (!out || !cfg) HD_THROW(HD_ERR_INVALID, "null argument");
  if (cfg->struct_size != (int32_t)sizeof(hd_config)) HD_THROW(HD_ERR_INVALID, "hd_config size mismatch");
  // model.py:198-200: the latent size must be a multiple of 16 (four stride-2 levels); 16 (image_res 128, the default,
  // train_refiner.py:27) runs the tuned kernels, 32 (image_res 256) the general ones
  if (cfg->latent_size != 16 && cfg->latent_size != 32)
    HD_THROW(HD_ERR_UNSUPPORTED, "latent_size %d: 16 (image_res 128) and 32 (image_res 256) are supported", cfg->latent_size);
  if (cfg->max_batch < 1) HD_THROW(HD_ERR_INVALID, "max_batch must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    HD_THROW(HD_ERR_UNSUPPORTED, "no CUDA device: hifidiff_b200 has no CPU path");
  }
  if (cfg->device < 0 || cfg->device >= ndev) HD_THROW(HD_ERR_INVALID, "device %d out of range", cfg->device);
  CUDA_CHECK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) HD_THROW(HD_ERR_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
  h = new hd_handle();
  struct Guard { hd_handle*& p; bool armed = true; ~Guard() { if (armed && p) { hd_destroy(p); p = nullptr; } } } guard{h};
  h->cfg = *cfg;
  h->tun.read_env();
  t_use_pdl = h->tun.pdl;
  h->fused = cfg->model == HD_MODEL_FUSED;
  h->bf16 = cfg->precision == HD_PRECISION_BF16;
  h->S = cfg->latent_size;
  h->sm_count = prop.multiProcessorCount; h->sm_major = prop.major; h->sm_minor = prop.minor;
  h->Bcap = ((cfg->max_batch + 127) / 128) * 128;
  h->max_steps = std::max(std::max(cfg->max_steps, 1), cfg->max_batch);
  for (int l = 0; l < kNumLevels; ++l) { h->c[l] = kWidth << l; h->sp[l] = h->S >> l; }
  CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) HD_THROW(HD_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  CUDA_CHECK(cudaFuncSetAttribute(dwconv_gate_pool_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 * 2));
  CUDA_CHECK(cudaFuncSetAttribute(dwconv_gate_pool_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 * 4));
  CUDA_CHECK(cudaFuncSetAttribute(ending_conv_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 * 2 + 4 * 9 * 128 * 4));
  CUDA_CHECK(cudaFuncSetAttribute(ending_conv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 * 4 + 4 * 9 * 128 * 4));
  CUDA_CHECK(cudaFuncSetAttribute(ending_conv_band_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 18 * 32 * 128 * 2 + 4 * 9 * 128 * 4));
  CUDA_CHECK(cudaFuncSetAttribute(ending_conv_band_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 10 * 32 * 128 * 4 + 4 * 9 * 128 * 4));
  // block table in execution order with table offsets
  int off = 0;
  auto push = [&](const std::string& prefix, int level) {
    BlockW b;
    b.prefix = prefix; b.level = level; b.c = h->c[level]; b.mod_off = off;
    off += 4 * b.c;
    h->blocks.push_back(b);
  };
  for (int l = 0; l < 4; ++l)
    for (int i = 0; i < kEncBlocks[l]; ++i) push("encoders." + std::to_string(l) + "." + std::to_string(i) + ".", l);
  for (int i = 0; i < kMidBlocks; ++i) push("middle_blks." + std::to_string(i) + ".", 4);
  for (int L = 0; L < 4; ++L)
    for (int i = 0; i < kDecBlocks[L]; ++i) push("decoders." + std::to_string(L) + "." + std::to_string(i) + ".", 3 - L);
  h->mod_stride = off;
  for (int j = 0; j < kNumLevels; ++j) { h->hca[j].d = h->c[4 - j]; h->hca[j].sp = h->sp[4 - j]; }

  // workspace
  Arena& A = h->arena;
  const size_t before = A.total;
  const size_t as = h->bf16 ? 2 : 4;
  const size_t Bc = h->Bcap;
  h->d_status = A.get<DeviceStatus>(1);
  CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&h->status_host), sizeof(DeviceStatus), cudaHostAllocDefault));
  memset(h->status_host, 0, sizeof(DeviceStatus));
  CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_status, cudaEventDisableTiming));
  size_t max_pc = 0;
  for (int l = 0; l < kNumLevels; ++l) {
    const size_t pc = Bc * h->sp[l] * h->sp[l] * h->c[l];
    h->resid[l] = A.get<float>(pc);
    max_pc = std::max(max_pc, pc);
  }
  h->act_bytes = max_pc * as;
  h->act_a = A.alloc(max_pc * as);
  h->act_h = A.alloc(2 * max_pc * as);
  h->act_g = A.alloc(max_pc * as);
  h->hca_out = A.alloc(max_pc * as);
  h->pooled_rows = Bc;
  h->pooled = A.alloc(h->pooled_rows * 2048 * as);
  h->sca_s = A.get<float>(h->pooled_rows * 2048);
  if (!h->bf16) h->gate_tmp = A.get<float>(2 * max_pc);
  const size_t xe = Bc * 4 * h->S * h->S;
  h->x_state = A.get<float>(xe);
  h->eps_buf = A.get<float>(xe);
  h->x_stage = A.get<float>(xe);
  const size_t R = h->max_steps;
  h->t_vals = A.get<float>(R);
  h->t_emb = A.get<float>(R * kWidth);
  h->t_h1 = A.get<float>(R * 2 * kTimeDim);
  h->t_g1 = A.get<float>(R * kTimeDim);
  h->t_temb = A.get<float>(R * kTimeDim);
  h->t_g2 = A.get<float>(R * 256);
  h->mod_table = A.get<float>(R * h->mod_stride);
  h->row_idx = A.get<int>(Bc);
  h->coefs = A.get<StepCoef>(R);
  h->state = A.get<StepState>(1);
  h->freqs = A.get<float>(64);
  {
    // model.py:24-26: exp(arange(64) * -(ln(1e4) / 63)) evaluated in fp32
    float f[64];
    const float step = static_cast<float>(-(std::log(10000.0) / 63.0));
    for (int i = 0; i < 64; ++i) f[i] = std::exp(static_cast<float>(i) * step);
    CUDA_CHECK(cudaMemcpy(h->freqs, f, sizeof(f), cudaMemcpyHostToDevice));
  }
  if (h->fused) {
    h->idc_add = A.get<float>(Bc * 2048 * (h->S / 16) * (h->S / 16));
    h->cond_nhwc = A.get<float>(max_pc);
    h->cond_stage = A.get<float>(max_pc);
    h->cond_pool = A.get<float>(Bc * 2048);
    h->cond_h = A.get<float>(Bc * 2048);
    h->cond_hs = A.get<float>(max_pc / 2);
    for (int j = 0; j < kNumLevels; ++j) {
      h->hca[j].wc = A.get<float>(Bc * h->hca[j].d);
      h->hca[j].ws = A.get<float>(Bc * h->hca[j].sp * h->hca[j].sp);
    }
  }
  h->workspace_bytes = A.total - before;
  guard.armed = false;
  *out = h;
  HD_API_END(null_handle)
}

int32_t hd_get_info(hd_handle* h, hd_info* info) {
  HD_API_BEGIN
  if (!h || !info) HD_THROW(HD_ERR_INVALID, "null argument");
  memset(info, 0, sizeof(*info));
  info->struct_size = sizeof(hd_info);
  info->abi_version = HD_ABI_VERSION;
  info->sm_major = h->sm_major; info->sm_minor = h->sm_minor; info->sm_count = h->sm_count;
  info->workspace_bytes = static_cast<int64_t>(h->workspace_bytes);
  info->weight_bytes = static_cast<int64_t>(h->arena.total - h->workspace_bytes);
  info->weight_elems_per_step = h->weight_elems_step;
  if (!h->plans.empty()) {
    const Plan& P = *h->plans.rbegin()->second;
    // launches of one SAMPLER step: the plan, plus the scheduler-step and advance kernels unless they are fused
    info->launches_per_step = static_cast<int32_t>(P.ops.size()) + (P.ending_idx >= 0 ? 0 : 2);
    info->flops_per_face_step = P.flops_per_face;
  }
  HD_API_END(h)
}

int32_t hd_load_weights(hd_handle* h, const hd_tensor_desc* tensors, int32_t n, void* stream) {
  HD_API_BEGIN
  if (!h || !tensors || n <= 0) HD_THROW(HD_ERR_INVALID, "null argument");
  if (h->weights_loaded) HD_THROW(HD_ERR_STATE, "weights already loaded; create a new handle to reload");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  (void)stream;
  CUDA_CHECK(cudaDeviceSynchronize());
  h->src.clear();
  for (int i = 0; i < n; ++i) {
    const hd_tensor_desc& t = tensors[i];
    if (!t.name || !t.data) continue;
    SrcTensor s;
    s.data = t.data; s.dtype = t.dtype;
    s.numel = 1;
    for (int k = 0; k < t.ndim && k < 4; ++k) { s.shape.push_back(t.shape[k]); s.numel *= static_cast<size_t>(t.shape[k]); }
    h->src[t.name] = s;
  }
  load_weights_impl(h);
  HD_API_END(h)
}

int32_t hd_set_time_frequencies(hd_handle* h, const float* freqs64) {
  HD_API_BEGIN
  if (!h || !freqs64) HD_THROW(HD_ERR_INVALID, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  CUDA_CHECK(cudaMemcpy(h->freqs, freqs64, 64 * sizeof(float), cudaMemcpyDefault));
  h->table_key.clear();
  HD_API_END(h)
}

int32_t hd_set_condition(hd_handle* h, const float* const priors[5], const float* identity, int32_t B, void* stream) {
  HD_API_BEGIN
  if (!h || !priors || !identity) HD_THROW(HD_ERR_INVALID, "null argument");
  if (!h->fused) HD_THROW(HD_ERR_STATE, "hd_set_condition needs HD_MODEL_FUSED");
  if (!h->weights_loaded) HD_THROW(HD_ERR_STATE, "hd_load_weights has not been called");
  if (B < 1 || B > h->cfg.max_batch) HD_THROW(HD_ERR_INVALID, "batch %d outside [1, %d]", B, h->cfg.max_batch);
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  join_in(h, stream);
  cudaStream_t st = h->stream;
  bool staged = false;  // a host-pointer input went through cond_stage
  for (int j = 0; j < kNumLevels; ++j) {
    HcaW& w = h->hca[j];
    const int d = w.d, hw = w.sp * w.sp, rows = B * hw;
    const size_t total = static_cast<size_t>(rows) * d;
    const float* src = priors[j];
    if (!src) HD_THROW(HD_ERR_INVALID, "priors[%d] is null", j);
    if (!is_device_ptr(src)) {
      if (staged) CUDA_CHECK(cudaStreamSynchronize(st));  // the previous prior still owns the staging buffer
      CUDA_CHECK(cudaMemcpyAsync(h->cond_stage, src, total * 4, cudaMemcpyHostToDevice, st));
      src = h->cond_stage;
      staged = true;
    }
    nchw_to_nhwc_kernel<<<cdiv(total, 256), 256, 0, st>>>(src, h->cond_nhwc, B, d, hw);
    // channel gate (hca.py:33-43)
    pool_avgmax_kernel<<<cdiv(static_cast<long long>(B) * d, 256), 256, 0, st>>>(h->cond_nhwc, h->cond_pool, B, hw, d);
    simt_f32(B, d, d, h->cond_pool, w.c0w, w.c0b, h->cond_h, EPI_RELU, st);
    simt_f32(B, d, d, h->cond_h, w.c2w, w.c2b, w.wc, EPI_SIGMOID, st);
    // spatial gate (hca.py:12-19,45-48), BN(eval) folded
    simt_f32(rows, d / 2, d, h->cond_nhwc, w.s0w, w.s0b, h->cond_hs, EPI_RELU, st);
    simt_f32(rows, 1, d / 2, h->cond_hs, w.s3w, w.s3b, w.ws, EPI_SIGMOID, st);
  }
  {  // idc_conv(identity) (model.py:245), identity is (B,2048,1,1) == (B,2048)
    const float* src = identity;
    if (!is_device_ptr(src)) {
      if (staged) CUDA_CHECK(cudaStreamSynchronize(st));
      CUDA_CHECK(cudaMemcpyAsync(h->cond_stage, src, static_cast<size_t>(B) * 2048 * 4, cudaMemcpyHostToDevice, st));
      src = h->cond_stage;
      staged = true;
    }
    const int hw = (h->S / 16) * (h->S / 16);   // bottleneck pixels per face; idc_conv has 2048 * hw output channels (model.py:198-200)
    if (hw == 1) {
      simt_f32(B, 2048, 2048, src, h->idc_w, h->idc_b, h->idc_add, EPI_BIAS, st);
    } else {  // (B, 2048 * hw) -> reshape(B, 2048, n, n) -> NHWC rows [B * hw][2048]
      simt_f32(B, 2048 * hw, 2048, src, h->idc_w, h->idc_b, h->cond_nhwc, EPI_BIAS, st);
      chw_to_hwc_rows_kernel<<<cdiv(static_cast<size_t>(B) * 2048 * hw, 256), 256, 0, st>>>(h->cond_nhwc, h->idc_add, B, 2048, hw);
    }
  }
  CUDA_CHECK(cudaGetLastError());
  // Device-pointer inputs: fully asynchronous (every buffer written above is consumed in stream order).  Host-pointer
  // inputs were copied from pageable memory the caller may free on return: wait for those copies.
  if (staged) CUDA_CHECK(cudaStreamSynchronize(st));
  h->condition_set = true;
  join_out(h, stream);
  HD_API_END(h)
}

int32_t hd_denoise_step(hd_handle* h, const float* x, const float* t, int32_t t_len, float* eps_out, int32_t batch,
                        void* stream) {
  HD_API_BEGIN
  if (!h || !x || !t || !eps_out) HD_THROW(HD_ERR_INVALID, "null argument");
  denoise_impl(h, x, t, t_len, eps_out, batch, nullptr, nullptr, 0, stream);
  HD_API_END(h)
}

int32_t hd_denoise_step_taps(hd_handle* h, const float* x, const float* t, int32_t t_len, float* eps_out,
                             int32_t batch, const char* const* tap_names, float* const* tap_out, int32_t n_taps,
                             void* stream) {
  HD_API_BEGIN
  if (!h || !x || !t || !eps_out) HD_THROW(HD_ERR_INVALID, "null argument");
  if (n_taps > 0 && (!tap_names || !tap_out)) HD_THROW(HD_ERR_INVALID, "null tap arrays");
  denoise_impl(h, x, t, t_len, eps_out, batch, tap_names, tap_out, n_taps, stream);
  HD_API_END(h)
}

int32_t hd_sampler_update(hd_handle* h, float* x_inout, const float* eps, const hd_step_coef* coef, int32_t step_index,
                          uint64_t seed, int64_t first_face, int32_t B, const float* noise, void* stream) {
  HD_API_BEGIN
  if (!h || !x_inout || !eps || !coef) HD_THROW(HD_ERR_INVALID, "null argument");
  if (B < 1 || B > h->cfg.max_batch) HD_THROW(HD_ERR_INVALID, "batch %d outside [1, %d]", B, h->cfg.max_batch);
  if (step_index < 0 || step_index >= h->max_steps) HD_THROW(HD_ERR_INVALID, "step_index out of range");
  if (!is_device_ptr(x_inout) || !is_device_ptr(eps) || (noise && !is_device_ptr(noise)))
    HD_THROW(HD_ERR_INVALID, "hd_sampler_update takes device pointers");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  join_in(h, stream);
  cudaStream_t st = h->stream;
  static_assert(sizeof(hd_step_coef) == sizeof(StepCoef), "coef layout");
  CUDA_CHECK(cudaMemcpyAsync(h->coefs + step_index, coef, sizeof(StepCoef), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  const int epf = 4 * h->S * h->S;
  const size_t threads = static_cast<size_t>(B) * epf / 4;
  // explicit noise here is (B, epf) for this one step: offset the pointer so the kernel's
  // step-major indexing lands on it
  const float* nz = noise ? noise - static_cast<size_t>(step_index) * B * epf : nullptr;
  sampler_update_kernel<<<cdiv(threads, 256), 256, 0, st>>>(x_inout, eps, h->coefs, nullptr, step_index, nz, seed,
                                                           first_face, B, epf);
  CUDA_CHECK(cudaGetLastError());
  join_out(h, stream);
  HD_API_END(h)
}

int32_t hd_sample(hd_handle* h, float* x_inout, const hd_step_coef* coef, int32_t n_steps, uint64_t seed,
                  int64_t first_face, int32_t B, const float* noise, void* stream) {
  HD_API_BEGIN
  if (!h || !x_inout || !coef) HD_THROW(HD_ERR_INVALID, "null argument");
  if (!h->weights_loaded) HD_THROW(HD_ERR_STATE, "hd_load_weights has not been called");
  if (h->fused && !h->condition_set) HD_THROW(HD_ERR_STATE, "hd_set_condition has not been called");
  if (B < 1 || B > h->cfg.max_batch) HD_THROW(HD_ERR_INVALID, "batch %d outside [1, %d]", B, h->cfg.max_batch);
  if (n_steps < 1 || n_steps > h->max_steps) HD_THROW(HD_ERR_INVALID, "n_steps %d outside [1, %d]", n_steps, h->max_steps);
  if (noise && !is_device_ptr(noise)) HD_THROW(HD_ERR_INVALID, "explicit noise must be a device pointer");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  join_in(h, stream);
  cudaStream_t st = h->stream;
  const int epf = 4 * h->S * h->S;
  const size_t xe = static_cast<size_t>(B) * epf;
  std::vector<float> ts(n_steps);
  for (int i = 0; i < n_steps; ++i) ts[i] = coef[i].timestep;
  ensure_time_table(h, ts);
  CUDA_CHECK(cudaMemcpyAsync(h->coefs, coef, n_steps * sizeof(StepCoef), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(h->x_state, x_inout, xe * 4, cudaMemcpyDefault, st));
  CUDA_CHECK(cudaMemsetAsync(h->row_idx, 0, h->Bcap * sizeof(int), st));
  set_step_kernel<<<1, 1, 0, st>>>(h->state, 0);
  h->cur_x = h->x_state;
  h->cur_eps = h->eps_buf;
  Plan* P = get_plan(h, B);
  const size_t threads = xe / 4;
  float* xs = h->x_state;
  float* eb = h->eps_buf;
  StepCoef* cf = h->coefs;
  StepState* ss = h->state;
  int* ridx = h->row_idx;
  const int Bcap = h->Bcap;
  const bool fused_end = P->ending_idx >= 0;
  edge::EndArgs ea;
  memset(&ea, 0, sizeof(ea));
  if (fused_end) {
    ea.x = static_cast<const bf16*>(h->hca_out); ea.w_hi = h->end_mma_hi; ea.w_lo = h->end_mma_lo; ea.bias = h->end_b;
    ea.x_state = xs; ea.coefs = cf; ea.state = ss; ea.noise = noise; ea.seed = seed; ea.first_face = first_face; ea.batch = B;
    ea.row_idx = ridx; ea.n_rows = Bcap; ea.ticket = h->end_ticket;
    CUDA_CHECK(cudaMemsetAsync(h->end_ticket, 0, sizeof(unsigned int), st));
  }
  auto one_step = [&](cudaStream_t s) {
    if (fused_end) {
      // ending conv + scheduler step + step advance in one launch (edge_convs.cuh)
      for (int i = 0; i < static_cast<int>(P->ops.size()); ++i)
        if (i != P->ending_idx) P->ops[i].fn(s);
      launch_k(edge::ending_mma_kernel<true>, dim3(B), dim3(256), edge::END_SMEM, s, ea);
      return;
    }
    for (auto& op : P->ops) op.fn(s);
    launch_k(sampler_update_kernel, dim3(cdiv(threads, 256)), dim3(256), 0, s, xs, eb, cf, ss, 0, noise, seed, first_face, B, epf);
    launch_k(advance_rows_kernel, dim3(1), dim3(256), 0, s, ss, ridx, Bcap);
  };
  // The graph bakes in seed / first_face / noise: re-capture when they change.
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "seed width");
  if (h->cfg.use_graph) {
    const bool stale = P->graph_seed != seed || P->graph_first != first_face || P->graph_noise != noise;
    if (P->graph == nullptr || stale) {
      if (P->graph) { cudaGraphExecDestroy(P->graph); P->graph = nullptr; }
      cudaGraph_t g = nullptr;
      CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      one_step(st);
      CUDA_CHECK(cudaStreamEndCapture(st, &g));
      CUDA_CHECK(cudaGraphInstantiate(&P->graph, g, 0));
      cudaGraphDestroy(g);
      P->graph_seed = seed; P->graph_first = first_face; P->graph_noise = noise;
    }
    for (int i = 0; i < n_steps; ++i) CUDA_CHECK(cudaGraphLaunch(P->graph, st));
  } else {
    for (int i = 0; i < n_steps; ++i) one_step(st);
  }
  CUDA_CHECK(cudaGetLastError());
  const bool x_dev = is_device_ptr(x_inout);
  CUDA_CHECK(cudaMemcpyAsync(x_inout, h->x_state, xe * 4, cudaMemcpyDefault, st));
  if (!x_dev) {
    CUDA_CHECK(cudaStreamSynchronize(st));
    check_device_status(h);
  }
  join_out(h, stream);
  HD_API_END(h)
}

int32_t hd_load_fpg_weights(hd_handle* h, const hd_tensor_desc* tensors, int32_t n, void* stream) {
  HD_API_BEGIN
  if (!h || !tensors || n <= 0) HD_THROW(HD_ERR_INVALID, "null argument");
  if (h->fpg.loaded) HD_THROW(HD_ERR_STATE, "FPG weights already loaded; create a new handle to reload");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  (void)stream;
  CUDA_CHECK(cudaDeviceSynchronize());
  h->src.clear();
  for (int i = 0; i < n; ++i) {
    const hd_tensor_desc& t = tensors[i];
    if (!t.name || !t.data) continue;
    SrcTensor s;
    s.data = t.data; s.dtype = t.dtype;
    s.numel = 1;
    for (int k = 0; k < t.ndim && k < 4; ++k) { s.shape.push_back(t.shape[k]); s.numel *= static_cast<size_t>(t.shape[k]); }
    h->src[t.name] = s;
  }
  load_fpg_impl(h);
  HD_API_END(h)
}

int32_t hd_fpg_forward(hd_handle* h, const float* cr_latent, float* const priors_out[5], int32_t B, void* stream) {
  HD_API_BEGIN
  if (!h || !cr_latent || !priors_out) HD_THROW(HD_ERR_INVALID, "null argument");
  if (!h->fpg.loaded) HD_THROW(HD_ERR_STATE, "hd_load_fpg_weights has not been called");
  if (B < 1 || B > h->cfg.max_batch) HD_THROW(HD_ERR_INVALID, "batch %d outside [1, %d]", B, h->cfg.max_batch);
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  join_in(h, stream);
  cudaStream_t st = h->stream;
  const size_t xe = static_cast<size_t>(B) * 4 * h->S * h->S;
  if (!is_device_ptr(cr_latent)) {
    CUDA_CHECK(cudaMemcpyAsync(h->x_stage, cr_latent, xe * 4, cudaMemcpyHostToDevice, st));
    h->fpg_in = h->x_stage;
  } else {
    h->fpg_in = cr_latent;
  }
  Plan* P = get_fpg_plan(h, B);
  for (auto& op : P->ops) op.fn(st);
  // priors: p0 and the four skip buffers (now skip + upsampled), NHWC fp32 -> NCHW fp32 (model.py:46-64 order)
  for (int j = 0; j < 5; ++j) {
    const int lvl = 4 - j, C = h->c[lvl], hw = h->sp[lvl] * h->sp[lvl];
    const float* src = j == 0 ? h->fpg.p0 : h->resid[lvl];
    float* dst = priors_out[j];
    if (!dst) HD_THROW(HD_ERR_INVALID, "priors_out[%d] is null", j);
    if (!is_device_ptr(dst)) HD_THROW(HD_ERR_INVALID, "hd_fpg_forward writes device buffers");
    const size_t total = static_cast<size_t>(B) * C * hw;
    nhwc_to_nchw_kernel<float><<<cdiv(total, 256), 256, 0, st>>>(src, dst, B, C, hw, C);
  }
  CUDA_CHECK(cudaGetLastError());
  join_out(h, stream);
  HD_API_END(h)
}

int32_t hd_load_idc_weights(hd_handle* h, const hd_tensor_desc* tensors, int32_t n, void* stream) {
  HD_API_BEGIN
  if (!h || !tensors || n <= 0) HD_THROW(HD_ERR_INVALID, "null argument");
  if (h->idc.loaded) HD_THROW(HD_ERR_STATE, "IDC weights already loaded; create a new handle to reload");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  (void)stream;
  CUDA_CHECK(cudaDeviceSynchronize());
  h->src.clear();
  for (int i = 0; i < n; ++i) {
    const hd_tensor_desc& t = tensors[i];
    if (!t.name || !t.data) continue;
    SrcTensor s;
    s.data = t.data; s.dtype = t.dtype;
    s.numel = 1;
    for (int k = 0; k < t.ndim && k < 4; ++k) { s.shape.push_back(t.shape[k]); s.numel *= static_cast<size_t>(t.shape[k]); }
    h->src[t.name] = s;
  }
  load_idc_impl(h);
  HD_API_END(h)
}

int32_t hd_idc_forward(hd_handle* h, const float* cr_face, int32_t image_size, float* identity_out, int32_t B,
                       void* stream) {
  HD_API_BEGIN
  if (!h || !cr_face || !identity_out) HD_THROW(HD_ERR_INVALID, "null argument");
  if (!h->idc.loaded) HD_THROW(HD_ERR_STATE, "hd_load_idc_weights has not been called");
  if (B < 1) HD_THROW(HD_ERR_INVALID, "batch %d < 1", B);
  const IdcW& I = h->idc;
  if (image_size != I.H) HD_THROW(HD_ERR_INVALID, "cr_face must be (B,3,%d,%d) for latent size %d, got %d", I.H, I.H, h->S, image_size);
  if (!is_device_ptr(identity_out)) HD_THROW(HD_ERR_INVALID, "hd_idc_forward writes a device buffer");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  join_in(h, stream);
  cudaStream_t st = h->stream;
  const bool in_dev = is_device_ptr(cr_face);
  const size_t per_face = static_cast<size_t>(3) * I.H * I.H;
  for (int off = 0; off < B; off += I.cap) {
    const int bc = std::min(I.cap, B - off);
    if (in_dev) {
      h->idc_in = cr_face + off * per_face;
    } else {
      CUDA_CHECK(cudaMemcpyAsync(I.stage, cr_face + off * per_face, bc * per_face * 4, cudaMemcpyHostToDevice, st));
      h->idc_in = I.stage;
    }
    h->idc_out = identity_out + static_cast<size_t>(off) * kIdcOut;
    Plan* P = get_idc_plan(h, bc);
    for (auto& op : P->ops) op.fn(st);
  }
  CUDA_CHECK(cudaGetLastError());
  if (!in_dev) CUDA_CHECK(cudaStreamSynchronize(st));
  if (!in_dev) check_device_status(h);
  join_out(h, stream);
  HD_API_END(h)
}

int32_t hd_load_cr_weights(hd_handle* h, const hd_tensor_desc* tensors, int32_t n, void* stream) {
  HD_API_BEGIN
  if (!h || !tensors || n <= 0) HD_THROW(HD_ERR_INVALID, "null argument");
  if (h->cr.loaded) HD_THROW(HD_ERR_STATE, "CR weights already loaded; create a new handle to reload");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  (void)stream;
  CUDA_CHECK(cudaDeviceSynchronize());
  h->src.clear();
  for (int i = 0; i < n; ++i) {
    const hd_tensor_desc& t = tensors[i];
    if (!t.name || !t.data) continue;
    SrcTensor s;
    s.data = t.data; s.dtype = t.dtype;
    s.numel = 1;
    for (int k = 0; k < t.ndim && k < 4; ++k) { s.shape.push_back(t.shape[k]); s.numel *= static_cast<size_t>(t.shape[k]); }
    h->src[t.name] = s;
  }
  load_cr_impl(h);
  HD_API_END(h)
}

int32_t hd_cr_forward(hd_handle* h, const float* ln_face, int32_t image_size, float* cr_face_out, int32_t B, void* stream) {
  HD_API_BEGIN
  if (!h || !ln_face || !cr_face_out) HD_THROW(HD_ERR_INVALID, "null argument");
  if (!h->cr.loaded) HD_THROW(HD_ERR_STATE, "hd_load_cr_weights has not been called");
  if (B < 1) HD_THROW(HD_ERR_INVALID, "batch %d < 1", B);
  const CrW& R = h->cr;
  if (image_size != R.H) HD_THROW(HD_ERR_INVALID, "ln_face must be (B,3,%d,%d), got %d", R.H, R.H, image_size);
  if (!is_device_ptr(cr_face_out)) HD_THROW(HD_ERR_INVALID, "hd_cr_forward writes a device buffer");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  join_in(h, stream);
  cudaStream_t st = h->stream;
  const bool in_dev = is_device_ptr(ln_face);
  const size_t per_face = static_cast<size_t>(3) * R.H * R.H;
  for (int off = 0; off < B; off += R.cap) {
    const int bc = std::min(R.cap, B - off);
    if (in_dev) {
      h->cr_in = ln_face + off * per_face;
    } else {
      CUDA_CHECK(cudaMemcpyAsync(R.stage, ln_face + off * per_face, bc * per_face * 4, cudaMemcpyHostToDevice, st));
      h->cr_in = R.stage;
    }
    h->cr_out = cr_face_out + off * per_face;
    Plan* P = get_cr_plan(h, bc);
    for (auto& op : P->ops) op.fn(st);
  }
  CUDA_CHECK(cudaGetLastError());
  if (!in_dev) CUDA_CHECK(cudaStreamSynchronize(st));
  join_out(h, stream);
  HD_API_END(h)
}

int32_t hd_profile_step(hd_handle* h, int32_t batch, int32_t reps, float* ms_out, char* labels_out, int32_t label_stride,
                        int32_t cap, int32_t* n_ops) {
  HD_API_BEGIN
  if (!h || !ms_out || !n_ops) HD_THROW(HD_ERR_INVALID, "null argument");
  if (!h->weights_loaded) HD_THROW(HD_ERR_STATE, "hd_load_weights has not been called");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  Plan* P = get_plan(h, batch);
  const int n = static_cast<int>(P->ops.size());
  *n_ops = n;
  if (n > cap) HD_THROW(HD_ERR_INVALID, "need room for %d ops", n);
  if (h->cur_x == nullptr) { h->cur_x = h->x_state; h->cur_eps = h->eps_buf; }
  cudaStream_t st = h->stream;
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
  std::vector<double> acc(n, 0.0);
  const int outer = reps < 0 ? -reps : reps;
  if (reps < 0) {
    // device-side cost: each launch captured 16x back to back into its own CUDA graph (no host launch
    // latency in the measurement; the step's numerics are meaningless in this mode)
    const int inner = 16;
    for (int i = 0; i < n; ++i) {
      cudaGraph_t g = nullptr;
      cudaGraphExec_t ge = nullptr;
      CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      for (int k = 0; k < inner; ++k) P->ops[i].fn(st);
      CUDA_CHECK(cudaStreamEndCapture(st, &g));
      CUDA_CHECK(cudaGraphInstantiate(&ge, g, 0));
      CUDA_CHECK(cudaGraphLaunch(ge, st));  // warm-up
      double best = 1e30;
      for (int rep = 0; rep < outer; ++rep) {
        CUDA_CHECK(cudaEventRecord(ev[0], st));
        CUDA_CHECK(cudaGraphLaunch(ge, st));
        CUDA_CHECK(cudaEventRecord(ev[1], st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, ev[0], ev[1]));
        best = std::min(best, static_cast<double>(ms));
      }
      acc[i] = best / inner * outer;
      cudaGraphExecDestroy(ge);
      cudaGraphDestroy(g);
    }
  } else {
    for (int rep = 0; rep < outer + 1; ++rep) {
      CUDA_CHECK(cudaEventRecord(ev[0], st));
      for (int i = 0; i < n; ++i) {
        P->ops[i].fn(st);
        CUDA_CHECK(cudaEventRecord(ev[i + 1], st));
      }
      CUDA_CHECK(cudaStreamSynchronize(st));
      if (rep == 0) continue;  // warm-up
      for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
        acc[i] += ms;
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    ms_out[i] = static_cast<float>(acc[i] / std::max(outer, 1));
    if (labels_out && label_stride > 0) {
      strncpy(labels_out + static_cast<size_t>(i) * label_stride, P->ops[i].label.c_str(), label_stride - 1);
      labels_out[static_cast<size_t>(i) * label_stride + label_stride - 1] = 0;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  check_device_status(h);
  HD_API_END(h)
}

int32_t hd_synchronize(hd_handle* h) {
  HD_API_BEGIN
  if (!h) HD_THROW(HD_ERR_INVALID, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  CUDA_CHECK(cudaGetLastError());
  check_device_status(h);
  HD_API_END(h)
}

static int g_time_reps = 0;       // set by hd_debug_gemm_time around one hd_debug_gemm call
static float g_time_ms = 0.f;
static long long* g_trace_dev = nullptr;  // set by hd_debug_gemm_trace around one hd_debug_gemm call
static int g_trace_ctas = 0;

int32_t hd_debug_gemm_time(hd_handle* h, const float* a, const float* w, float* out, int32_t m, int32_t n, int32_t k,
                           int32_t mode, int32_t reps, float* ms_per_launch) {
  if (!ms_per_launch || reps < 1) return HD_ERR_INVALID;
  g_time_reps = reps;
  g_time_ms = 0.f;
  const int32_t rc = hd_debug_gemm(h, a, w, nullptr, out, m, n, k, mode, nullptr);
  g_time_reps = 0;
  *ms_per_launch = g_time_ms;
  return rc;
}

int32_t hd_debug_gemm_trace(hd_handle* h, const float* a, const float* w, float* out, int32_t m, int32_t n, int32_t k,
                            long long* trace_host, int32_t cap_ctas, int32_t* n_ctas, int32_t* grid_xyz) {
  int32_t rc;
  {
    HD_API_BEGIN
    if (!h || !trace_host || !n_ctas) HD_THROW(HD_ERR_INVALID, "null argument");
    CUDA_CHECK(cudaSetDevice(h->cfg.device));
    CUDA_CHECK(cudaMalloc(&g_trace_dev, static_cast<size_t>(cap_ctas) * 16 * sizeof(long long)));
    CUDA_CHECK(cudaMemset(g_trace_dev, 0, static_cast<size_t>(cap_ctas) * 16 * sizeof(long long)));
    g_trace_ctas = cap_ctas;
    } catch (const HdError& e) { h->err = e.msg; return e.code; }
  }
  rc = hd_debug_gemm(h, a, w, nullptr, out, m, n, k, 1, nullptr);
  if (rc == HD_OK) {
    cudaMemcpy(trace_host, g_trace_dev, static_cast<size_t>(cap_ctas) * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
    *n_ctas = g_trace_ctas;
    if (grid_xyz) { grid_xyz[0] = g_trace_ctas; }
  }
  cudaFree(g_trace_dev);
  g_trace_dev = nullptr;
  return rc;
}

int32_t hd_debug_gemm(hd_handle* h, const float* a, const float* w, const float* bias, float* out, int32_t m,
                      int32_t n, int32_t k, int32_t use_tc, void* stream) {
  HD_API_BEGIN
  if (!h || !a || !w || !out) HD_THROW(HD_ERR_INVALID, "null argument");
  if (m < 1 || n < 1 || k < 1) HD_THROW(HD_ERR_INVALID, "bad shape");
  if (use_tc && (k % 64 != 0 || n % 128 != 0)) HD_THROW(HD_ERR_INVALID, "tensor-core GEMM needs K %% 64 == 0 and N %% 128 == 0");
  if (!is_device_ptr(a) || !is_device_ptr(w) || !is_device_ptr(out) || (bias && !is_device_ptr(bias)))
    HD_THROW(HD_ERR_INVALID, "hd_debug_gemm takes device pointers");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  join_in(h, stream);
  cudaStream_t st = h->stream;
  const long long m_alloc = ((m + 255) / 256) * 256;
  void *da = nullptr, *dw = nullptr;
  float* zb = nullptr;
  CUDA_CHECK(cudaMalloc(&zb, n * 4));
  CUDA_CHECK(cudaMemsetAsync(zb, 0, n * 4, st));
  GemmDesc d;
  d.M = m; d.N = n; d.K = k; d.lda = k; d.ldw = k; d.bias = bias ? bias : zb; d.epi = EPI_BIAS;
  d.out = out; d.ldo = n; d.out_dtype = DT_F32;
  if (use_tc) {
    CUDA_CHECK(cudaMalloc(&da, m_alloc * k * 2));
    CUDA_CHECK(cudaMalloc(&dw, static_cast<size_t>(n) * k * 2));
    CUDA_CHECK(cudaMemsetAsync(da, 0, m_alloc * k * 2, st));
    const size_t ta = static_cast<size_t>(m) * k / 8, tw = static_cast<size_t>(n) * k / 8;
    launch_k(cast_kernel<bf16>, dim3(cdiv(ta, 256)), dim3(256), 0, st, a, static_cast<bf16*>(da), ta);
    launch_k(cast_kernel<bf16>, dim3(cdiv(tw, 256)), dim3(256), 0, st, w, static_cast<bf16*>(dw), tw);
    d.A = da; d.W = dw; d.a_dtype = DT_BF16; d.w_dtype = DT_BF16;
    const int saved = h->tun.two_cta;
    h->tun.two_cta = use_tc == 2 ? 2 : (use_tc == 3 ? 0 : saved);  // 2: force cta_group::2 pairs, 3: force single-CTA tiles
    TcLaunch L = build_tc(h, d, m_alloc);
    h->tun.two_cta = saved;
    if (use_tc == 2 && !L.two_cta) HD_THROW(HD_ERR_INVALID, "shape not eligible for the 2-CTA kernel (N %% 256)");
    launch_tc(L, st);  // warm-up (also warms L2 with the operands)
    launch_tc(L, st);
    if (g_time_reps > 0) {
      cudaEvent_t e0, e1;
      CUDA_CHECK(cudaEventCreate(&e0));
      CUDA_CHECK(cudaEventCreate(&e1));
      CUDA_CHECK(cudaEventRecord(e0, st));
      for (int i = 0; i < g_time_reps; ++i) launch_tc(L, st);
      CUDA_CHECK(cudaEventRecord(e1, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      CUDA_CHECK(cudaEventElapsedTime(&g_time_ms, e0, e1));
      g_time_ms /= g_time_reps;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    if (g_trace_dev != nullptr) {
      const int ctas = static_cast<int>(L.grid.x * L.grid.y * L.grid.z);
      if (ctas <= g_trace_ctas) {
        L.args.trace = g_trace_dev;
        g_trace_ctas = ctas;
        launch_tc(L, st);
        L.args.trace = nullptr;
      }
    }
  } else {
    d.A = a; d.W = w;
    launch_simt(d, st);
  }
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(st));
  cudaFree(da); cudaFree(dw); cudaFree(zb);
  check_device_status(h);
  join_out(h, stream);
  HD_API_END(h)
}

}  // extern "C"
