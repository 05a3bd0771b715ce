/*
 * hifidiff_b200 — C ABI of the B200-native HifiDiff reverse-sampling hot path.
 *
 * The reference (js43o/HifiDiff) is pure PyTorch and has no FFI of its own; these entry points
 * are what a binding for its hot path would call.  Each one names the reference interface it
 * replaces (paths relative to the reference tree).  All functions return an hd_status
 * (0 = ok); the message for the last failure is available from hd_last_error().  Nothing
 * throws across this boundary, nothing here falls back to a CPU path: a device that is not
 * sm_100 is a hard error at hd_create().
 *
 * Conventions
 *   - Tensors crossing the boundary are dense fp32, NCHW, exactly as the reference's modules
 *     take and return them (latents (B,4,S,S), priors (B,C,n,n), identity (B,2048,1,1)).
 *   - Data pointers may be device pointers (on the handle's device) or host pointers; host
 *     buffers are staged through the library's own pinned/device buffers on `stream`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued on it; calls do not synchronise unless a host pointer has to be read or written
 *     (host inputs come from pageable memory the caller may free on return).
 *   - Device-side faults (the tcgen05 pipeline watchdog) are reported late, never dropped: every
 *     call that enqueues kernels ends with an asynchronous copy of the 8-byte status word into
 *     pinned memory; the NEXT call on the handle, and hd_synchronize(), return HD_ERR_KERNEL if
 *     that word is set.
 *   - One handle per (process, device); a handle is not thread-safe.
 *   - The caller owns every buffer it passes; the library owns only its packed-weight arena,
 *     activation workspace and time-modulation tables.
 */
#ifndef HIFIDIFF_B200_H_
#define HIFIDIFF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HD_ABI_VERSION 1

typedef struct hd_handle hd_handle;

typedef enum hd_status {
  HD_OK = 0,
  HD_ERR_INVALID = 1,      /* bad argument / shape / missing tensor */
  HD_ERR_CUDA = 2,         /* a CUDA runtime / driver call failed */
  HD_ERR_UNSUPPORTED = 3,  /* not an sm_100 device, unsupported size */
  HD_ERR_STATE = 4,        /* call order (weights / condition / schedule not set) */
  HD_ERR_KERNEL = 5        /* a device-side watchdog tripped (pipeline barrier timeout) */
} hd_status;

typedef enum hd_model_kind {
  HD_MODEL_DENOISER = 0, /* models/denoiser/model.py:32-134  Denoiser(latent_size)      */
  HD_MODEL_FUSED = 1     /* models/denoiser/model.py:137-266 FusedDenoiser(latent_size) */
} hd_model_kind;

typedef enum hd_precision {
  HD_PRECISION_BF16 = 0, /* bf16 operands on tcgen05, fp32 accumulate, fp32 residual stream */
  HD_PRECISION_FP32 = 1  /* fp32 FFMA everywhere (correctness mode, rel-L2 <= 1e-5)          */
} hd_precision;

typedef struct hd_config {
  int32_t struct_size;  /* = sizeof(hd_config) */
  int32_t model;        /* hd_model_kind */
  int32_t precision;    /* hd_precision */
  int32_t latent_size;  /* S: latents are (B,4,S,S); 16 or 32 (model.py:198-200; --image_res 128 / 256) */
  int32_t device;       /* CUDA device ordinal */
  int32_t max_batch;    /* workspace is sized for this many faces per call */
  int32_t max_steps;    /* rows of the time-modulation table (>= sampler steps) */
  int32_t use_graph;    /* 1: replay the per-step launch sequence as a CUDA graph in hd_sample */
} hd_config;

/* One named parameter/buffer of the module's state_dict().  dtype: 0 = fp32,
 * 1 = int64 (only BatchNorm num_batches_tracked, ignored). */
typedef struct hd_tensor_desc {
  const char* name;
  const void* data;
  int32_t dtype;
  int32_t ndim;
  int64_t shape[4];
} hd_tensor_desc;

/* Per-step coefficients of the x_{t-1} update, computed by the host scheduler exactly as
 * diffusers 0.32.2 does (DDIMScheduler.step / DDPMScheduler.step; reference call sites
 * train_refiner.py:120, pretrain_denoiser.py:110, test_refiner.py:91):
 *     x0   = (x - sqrt_beta_prod * eps) / sqrt_alpha_prod ;  x0 = clamp(x0, -clip, clip) if clip > 0
 *     x'   = k_x0 * x0 + k_eps * eps + k_x * x + k_noise * z
 * DDIM(eta): k_x0 = sqrt(a_prev), k_eps = sqrt(1 - a_prev - std^2), k_x = 0, k_noise = std
 * DDPM     : k_x0 = sqrt(a_prev) * beta_t / (1 - a_t), k_eps = 0,
 *            k_x = sqrt(alpha_t) * (1 - a_prev) / (1 - a_t), k_noise = sqrt(var) (0 at t = 0) */
typedef struct hd_step_coef {
  float timestep;        /* value fed to the time embedding (model.py:22-29) */
  float sqrt_beta_prod;  /* sqrt(1 - alphas_cumprod[t]) */
  float sqrt_alpha_prod; /* sqrt(alphas_cumprod[t]) */
  float clip;            /* <= 0: no clipping */
  float k_x0, k_eps, k_x, k_noise;
} hd_step_coef;

typedef struct hd_info {
  int32_t struct_size;
  int32_t abi_version;
  int32_t sm_major, sm_minor, sm_count;
  int32_t launches_per_step; /* kernels launched by one denoise step at the current batch */
  int64_t weight_bytes;      /* packed weight arena */
  int64_t workspace_bytes;   /* activations + tables */
  int64_t weight_elems_per_step; /* weight elements one denoise step streams */
  double flops_per_face_step;    /* executed MAC*2 per face per step (incl. padding taps) */
} hd_info;

int32_t hd_abi_version(void);

/* Create / destroy a handle.  Replaces module construction: Denoiser(latent_size) model.py:33,
 * FusedDenoiser(latent_size) model.py:138. */
int32_t hd_create(hd_handle** out, const hd_config* cfg);
void hd_destroy(hd_handle* h);
/* Message of the last error on this handle (h == NULL: last hd_create failure). */
const char* hd_last_error(const hd_handle* h);
int32_t hd_get_info(hd_handle* h, hd_info* info);

/* Hand the module's state_dict to the library, which repacks it into its own arena (bf16,
 * K-major, BatchNorm(eval) folded, beta/gamma folded into conv3/conv5, gate-interleaved conv4).
 * Replaces nn.Module.load_state_dict as used at refiner.py:22-25 / test_refiner.py:162-164.
 * Unknown names are ignored; a missing required tensor is HD_ERR_INVALID. */
int32_t hd_load_weights(hd_handle* h, const hd_tensor_desc* tensors, int32_t n, void* stream);

/* Optional: override the 64 sinusoidal frequencies (model.py:24-26) with host values. */
int32_t hd_set_time_frequencies(hd_handle* h, const float* freqs64);

/* FacialPriorGuidance on the same kernels (SURVEY.md §8f "next" row 1).  hd_load_fpg_weights takes the
 * FPG module's state_dict (intro.*, encoders.L.i.*, downs.L.*, convs.j.0.weight; replaces
 * fpg.load_state_dict, refiner.py:25); hd_fpg_forward replaces FacialPriorGuidance.forward
 * (models/fpg/model.py:46-64): cr_latent (B,4,S,S) -> priors_out[j] (B, C_j, n_j, n_j) fp32 device buffers. */
int32_t hd_load_fpg_weights(hd_handle* h, const hd_tensor_desc* tensors, int32_t n, void* stream);
int32_t hd_fpg_forward(hd_handle* h, const float* cr_latent, float* const priors_out[5], int32_t batch,
                       void* stream);

/* IDC identity network on the same kernels (SURVEY.md §8f "next" row 2).  hd_load_idc_weights takes the
 * ResNet-50 module's state_dict (conv1.weight, batch_norm1.*, layerL.i.{conv1,conv2,conv3}.{weight,bias},
 * layerL.i.batch_normK.*, layerL.0.i_downsample.{0,1}.*; replaces idc.load_state_dict, refiner.py:18-20) and
 * folds every eval-mode BatchNorm into its conv.  hd_idc_forward replaces ResNet.forward
 * (models/idc/model.py:123-136) as called at refiner.py:34: cr_face (B,3,image_size,image_size) fp32, device or
 * host, image_size == 8 * latent_size -> identity_out (B,2048,1,1) fp32 device buffer.  Any batch >= 1 (faces are
 * processed in chunks of 64). */
int32_t hd_load_idc_weights(hd_handle* h, const hd_tensor_desc* tensors, int32_t n, void* stream);
int32_t hd_idc_forward(hd_handle* h, const float* cr_face, int32_t image_size, float* identity_out,
                       int32_t batch, void* stream);

/* CoarseRestoration on CUDA kernels (SURVEY.md §8f "next" row 3), the stage before the sampling loop.
 * hd_load_cr_weights takes the CR module's state_dict (intro.*, outro.*, {encoders.i,middle_blocks,decoders.j}.
 * {nfbs.k.*, stn.localization.{0,3}.*, stn.fc_loc.{0,2}.*, sampling.*}; replaces cr_module.load_state_dict,
 * train_refiner.py:377-379).  hd_cr_forward replaces CoarseRestoration.forward (models/cr/model.py:75-88) as
 * called at train_refiner.py:106: ln_face (B,3,128,128) fp32, device or host -> cr_face_out (B,3,128,128) fp32
 * device buffer.  The residual stream and the spatial transformers are fp32 in both precision modes; with
 * HD_PRECISION_BF16 the 1x1 convs at c >= 128 run as split-precision (3 x bf16, fp32 accumulate) tcgen05 GEMMs,
 * with HD_PRECISION_FP32 every GEMM is FFMA.  Any batch >= 1 (faces are processed in chunks of 32). */
int32_t hd_load_cr_weights(hd_handle* h, const hd_tensor_desc* tensors, int32_t n, void* stream);
int32_t hd_cr_forward(hd_handle* h, const float* ln_face, int32_t image_size, float* cr_face_out, int32_t batch,
                      void* stream);

/* Condition-only work, hoisted out of the timestep loop (it depends on neither x_t nor t):
 * idc_conv(identity) (model.py:245-246) and the five HCA channel/spatial gates computed from
 * the priors (hca.py:33-48).  priors[j]: (B, C_j, n_j, n_j) with C = 2048,1024,512,256,128 and
 * n = S/16 * (1,2,4,8,16); identity: (B,2048,1,1).  HD_MODEL_FUSED only. */
int32_t hd_set_condition(hd_handle* h, const float* const priors[5], const float* identity,
                         int32_t batch, void* stream);

/* One epsilon prediction.  Replaces Denoiser.forward (model.py:106-134) / FusedDenoiser.forward
 * (model.py:217-266).  t has t_len == 1 (shared) or t_len == batch entries.
 * The time-modulation table is keyed by the VALUES of t, so this entry point reads t on the host and
 * waits for `stream` (the one exception to "calls do not synchronise" besides host buffers); the
 * sampling loop proper is hd_sample, which runs every step from device-resident state. */
int32_t hd_denoise_step(hd_handle* h, const float* x, const float* t, int32_t t_len,
                        float* eps_out, int32_t batch, void* stream);

/* Same, and additionally copies named intermediate activations (fp32 NCHW) for per-layer
 * parity.  tap_names use the reference module paths ("intro", "encoders.0.1", "downs.2",
 * "middle_blks.7", "ups.0", "decoders.3.1", "hcas.4", "time_mlp").  tap_out[i] must hold the
 * layer's full output for `batch` faces. */
int32_t hd_denoise_step_taps(hd_handle* h, const float* x, const float* t, int32_t t_len,
                             float* eps_out, int32_t batch, const char* const* tap_names,
                             float* const* tap_out, int32_t n_taps, void* stream);

/* Whole reverse-sampling loop on a batch of faces: for each step i,
 *   eps = model(x, coef[i].timestep) ; x = update(x, eps, coef[i], z_i)
 * Replaces the loop body of ddim_sample (train_refiner.py:111-120, pretrain_denoiser.py:101-110,
 * test_refiner.py:87-91).  z_i is Philox4x32-10 noise keyed by (seed, first_face + b, i, elem)
 * unless `noise` (n_steps, batch, 4*S*S) is given.  x_inout: (batch,4,S,S), updated in place. */
int32_t hd_sample(hd_handle* h, float* x_inout, const hd_step_coef* coef, int32_t n_steps,
                  uint64_t seed, int64_t first_face, int32_t batch, const float* noise,
                  void* stream);

/* The x_{t-1} update alone (one vectorised elementwise kernel); eps and x are (batch,4,S,S).
 * Replaces scheduler.step(noise_pred, t, latents, eta).prev_sample (train_refiner.py:120). */
int32_t hd_sampler_update(hd_handle* h, float* x_inout, const float* eps, const hd_step_coef* coef,
                          int32_t step_index, uint64_t seed, int64_t first_face, int32_t batch,
                          const float* noise, void* stream);

/* Waits for all work the handle has enqueued and reports a tripped device-side watchdog
 * (HD_ERR_KERNEL) or a sticky CUDA error.  The asynchronous entry points report a fault of an earlier
 * call when they are entered (see Conventions); this is the call that reports it for the last one. */
int32_t hd_synchronize(hd_handle* h);

/* Times every kernel of one denoise step (plain launches, CUDA events between launches, warm L2),
 * averaged over `reps` repetitions.  ms_out[i] / labels_out[i*label_stride] describe launch i. */
int32_t hd_profile_step(hd_handle* h, int32_t batch, int32_t reps, float* ms_out, char* labels_out,
                        int32_t label_stride, int32_t cap, int32_t* n_ops);

/* Standalone C = A[M,K] * W[N,K]^T (+bias) on the tcgen05 path (bf16 operands given as fp32,
 * converted internally); used by the parity tests to pin the tensor-core kernel alone.
 * use_tensor_cores: 0 = fp32 FFMA, 1 = library heuristics, 2 = force cta_group::2 pairs (N % 256 == 0),
 * 3 = force single-CTA tiles. */
int32_t hd_debug_gemm(hd_handle* h, const float* a, const float* w, const float* bias, float* out,
                      int32_t m, int32_t n, int32_t k, int32_t use_tensor_cores, void* stream);

/* hd_debug_gemm (mode = its use_tensor_cores) followed by `reps` back-to-back launches of the same GEMM timed
 * with CUDA events on the library stream: milliseconds per launch, operands L2-warm. */
int32_t hd_debug_gemm_time(hd_handle* h, const float* a, const float* w, float* out, int32_t m, int32_t n,
                           int32_t k, int32_t mode, int32_t reps, float* ms_per_launch);

/* hd_debug_gemm on the tensor-core path with a per-CTA clock64 timeline (16 slots per CTA:
 * 0 entry, 1 setup done, 2 first operands landed, 3 last MMA issued, 4 accumulator ready,
 * 5 staged, 6 after tile barrier, 7 stored, 8 exit, 9 globaltimer at entry, 10 SM id). */
int32_t hd_debug_gemm_trace(hd_handle* h, const float* a, const float* w, float* out, int32_t m, int32_t n,
                            int32_t k, long long* trace_host, int32_t cap_ctas, int32_t* n_ctas,
                            int32_t* grid_xyz);

#ifdef __cplusplus
}
#endif
#endif /* HIFIDIFF_B200_H_ */
