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

namespace {

thread_local std::string g_create_error;

std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return std::string(buf);
}

struct HdError {
  int code;
  std::string msg;
};

#define HD_THROW(code, ...) throw HdError{code, fmt(__VA_ARGS__)}
#define CUDA_CHECK(expr)                                                                              \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      HD_THROW(HD_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

constexpr int kNumLevels = 5;
constexpr int kEncBlocks[4] = {2, 2, 4, 8};  // models/denoiser/model.py:80
constexpr int kMidBlocks = 8;                // model.py:89-91
constexpr int kDecBlocks[4] = {2, 2, 2, 2};  // model.py:93
constexpr int kWidth = 128;                  // model.py:36
constexpr int kTimeDim = 512;                // model.py:44
constexpr float kBnEps = 1e-5f;

inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// Tuning switches, read from the environment once per handle at hd_create (a handle keeps its own copy: two handles
// created under different environments do not see each other's settings).
struct Tunables {
  bool pdl = true;        // HD_PDL=0 disables programmatic dependent launch
  bool bn256 = true;      // HD_BN256=0: 128x128 tiles for the dense 3x3 convs too
  int two_cta = 1;        // HD_TWO_CTA=0: never use cta_group::2 pairs; 2 (set by hd_debug_gemm only): wherever the shape allows
  bool face = true;       // HD_FACE=0: per-op kernels at the 16x16 level instead of the fused per-face block kernel
  bool pair = true;       // HD_PAIR=0: per-op kernels at the 8x8 level instead of the fused face-pair block kernel
  bool sca_mul = true;    // HD_SCA_MUL=0: separate scale_rows kernel at the 1x1 level too
  bool w_prefetch = true; // HD_W_PREFETCH=0: GEMMs do not prefetch the next GEMM's weights into L2
  bool edge_mma = true;   // HD_EDGE_MMA=0: CUDA-core intro / ending convs and separate sampler-update / advance launches
  bool cr_stn_cs = true;  // HD_CR_STN_CS=0: one thread per (pixel, 2 output channels) in the first STN localisation conv
  bool cr_tc = true;      // HD_CR_TC=0: every CoarseRestoration GEMM on the FFMA kernel (no split-precision tcgen05 path)
  bool cr_mma3 = true;    // HD_CR_MMA3=0: the shallow CoarseRestoration stages (c = 32 / 64, down / up convs) stay on the FFMA GEMM
  bool cr_mma3h = true;   // HD_CR_MMA3H=0: K = 32 / 64 GEMMs of CoarseRestoration on the 3xTF32 kernel instead of the row-scaled fp16 split
  bool cr_fuse_split = true; // HD_CR_FUSE_SPLIT=0: separate fp32 -> [hi|lo|hi] kernels in front of the split tcgen05 GEMMs
  bool cr_stn_mma = true; // HD_CR_STN_MMA=0: the first STN localisation conv stays on CUDA cores
  bool cr_dw_strip = true; // HD_CR_DW_STRIP=0: CoarseRestoration depthwise conv one thread per pixel instead of per column strip
  bool w_evict_first = false; // HD_W_EVICT_FIRST=1: weight tiles enter L2 with evict-first priority (activations and code stay)
  bool face_warm = true;  // HD_FACE_WARM=0: no instruction-cache warm-up / first-wave-only prefetch in the fused face kernel
  bool dw_small = true;   // HD_DW_SMALL=0: the generic tiled depthwise kernel at the 2x2 / 4x4 levels too
  int cta_target = 120;   // HD_CTA_TARGET: split-K until a GEMM's grid has at least this many CTAs
  int max_split = 4;      // HD_MAX_SPLIT: deepest split-K (cluster size along z); 8-way DSMEM reductions measured slower at every batch (1 .. 256 faces: -1 .. -12 % per step with 4)
  int sca_target = 120;   // HD_SCA_TARGET: the same for the SCA GEMMs (M = faces)
  int cr_chunk = 128;     // HD_CR_CHUNK: faces per CoarseRestoration pass (~14 MB of fp32 workspace per face; 32: 66 ms, 64: 53 ms, 128: 48 ms, 256: 46 ms per 256 faces)
  void read_env() {
    auto flag = [](const char* name, bool& v) { if (const char* e = getenv(name)) v = atoi(e) != 0; };
    flag("HD_PDL", pdl); flag("HD_BN256", bn256); flag("HD_FACE", face); flag("HD_PAIR", pair); flag("HD_SCA_MUL", sca_mul); flag("HD_EDGE_MMA", edge_mma); flag("HD_W_PREFETCH", w_prefetch);
    flag("HD_CR_STN_CS", cr_stn_cs); flag("HD_CR_TC", cr_tc); flag("HD_CR_MMA3", cr_mma3); flag("HD_CR_STN_MMA", cr_stn_mma); flag("HD_CR_FUSE_SPLIT", cr_fuse_split); flag("HD_CR_MMA3H", cr_mma3h); flag("HD_CR_DW_STRIP", cr_dw_strip); flag("HD_DW_SMALL", dw_small); flag("HD_FACE_WARM", face_warm); flag("HD_W_EVICT_FIRST", w_evict_first);
    if (const char* e = getenv("HD_TWO_CTA")) two_cta = atoi(e) != 0 ? 1 : 0;  // 2 (pairs wherever the shape allows) only through hd_debug_gemm: a whole plan forced onto pairs hung in round 2
    if (const char* e = getenv("HD_CTA_TARGET")) cta_target = std::max(atoi(e), 1);
    if (const char* e = getenv("HD_MAX_SPLIT")) max_split = std::min(std::max(atoi(e), 1), 8);
    if (const char* e = getenv("HD_SCA_TARGET")) sca_target = std::max(atoi(e), 1);
    if (const char* e = getenv("HD_CR_CHUNK")) cr_chunk = std::min(std::max(atoi(e), 1), 256);
  }
};
// PDL attribute of the launches issued by the calling thread: set from the handle's Tunables by every entry point
// that launches kernels (set_launch_tunables), so the launch helpers need no handle argument.
thread_local bool t_use_pdl = true;

// Per-step kernel launch: programmatic stream serialization lets kernel N+1 be scheduled (and run its
// prologue / weight prefetch) while kernel N drains; every such kernel executes pdl_wait() first.
template <typename... KArgs, typename... Args>
void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = t_use_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------
// chunked bump allocator for everything the library owns on the device
// ------------------------------------------------------------------------------------------------
struct Arena {
  std::vector<void*> chunks;
  char* cur = nullptr;
  size_t left = 0;
  size_t total = 0;
  size_t chunk_bytes = size_t(256) << 20;
  void* alloc(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes > left) {
      size_t sz = std::max(bytes, chunk_bytes);
      void* p = nullptr;
      CUDA_CHECK(cudaMalloc(&p, sz));
      CUDA_CHECK(cudaMemset(p, 0, sz));
      chunks.push_back(p);
      cur = static_cast<char*>(p);
      left = sz;
      total += sz;
    }
    void* r = cur;
    cur += bytes;
    left -= bytes;
    return r;
  }
  template <typename T> T* get(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }
  void release() {
    for (void* p : chunks) cudaFree(p);
    chunks.clear();
    cur = nullptr;
    left = total = 0;
  }
};

struct BlockW {
  std::string prefix;
  int level = 0, c = 0, mod_off = 0;
  float *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr;
  void *w1 = nullptr, *wsca = nullptr, *w3 = nullptr, *w4 = nullptr, *w5 = nullptr;
  float *b1 = nullptr, *bsca = nullptr, *b3 = nullptr, *b4 = nullptr, *b5 = nullptr;
  float *dw_w = nullptr, *dw_b = nullptr;
  float* wsca_t = nullptr;  // 16x16 level: SCA weight transposed [k][n] fp32 for the fused face kernel's GEMV
  void* wsca_tb = nullptr;  // 8x8 level: the same, bf16, for the fused face-pair kernel
  std::vector<float> b3_h, b5_h;  // host copies of the folded conv3 / conv5 biases (cumulative residual bias)
  bool dw_folded = false;  // 1x1 level: depthwise 3x3 == per-channel scale, folded into conv1 (gate-packed)
  bool has_mod = true;     // false: unconditional NAFBlock of the FPG encoder (models/fpg/naf.py:105-126)
};

// FacialPriorGuidance (models/fpg/model.py:7-64): NAFNet encoder over the CR latent, run once per face batch
struct FpgW {
  bool loaded = false;
  std::vector<BlockW> blocks;
  void* down_w[4] = {};
  float* down_b[4] = {};
  float *intro_w = nullptr, *intro_b = nullptr;
  void* convs_w[5] = {};     // convs[0]: plain 1x1; convs[1..4]: 1x1 + PixelShuffle(2), rows grouped by quadrant
  float* zero_bias = nullptr;
  float* p0 = nullptr;       // prior 0 (B, 2048) fp32
};

// IDC identity network (models/idc/model.py:10-55,102-166): ResNet-50 trunk, BatchNorm(eval) folded
struct IdcConvW {
  void* w = nullptr;   // [N][K] operand dtype, BN scale folded, zero-padded to N % 128 == 0 / C % 64 == 0
  float* b = nullptr;  // [N] folded shift
  int N = 0, K = 0;    // padded
};
struct IdcBlockW {
  IdcConvW c1, c2, c3, proj;
  bool has_proj = false;
  int stride = 1, cin = 0, planes = 0, pp = 0;  // pp = planes padded to a multiple of 128
  int n_in = 0;                                 // spatial size of the block input
};
struct IdcW {
  bool loaded = false;
  int H = 0, cap = 0;  // image size; faces per chunk the workspace holds
  float *stem_w = nullptr, *stem_b = nullptr;
  std::vector<IdcBlockW> blocks;
  void *stem_out = nullptr, *pool_t = nullptr, *xb[2] = {}, *t1 = nullptr, *t2 = nullptr, *col = nullptr;
  float *xf[2] = {}, *stage = nullptr, *out_stage = nullptr;
  double flops_per_face = 0;
};

// CoarseRestoration (models/cr/model.py:8-88): fp32 throughout, plain (unpacked) weights for the FFMA GEMM
struct CrBlockW {
  int c = 0;
  float *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr;
  float *w1 = nullptr, *b1 = nullptr, *dw_w = nullptr, *dw_b = nullptr, *wsca = nullptr, *bsca = nullptr;
  float *w3 = nullptr, *b3 = nullptr, *w4 = nullptr, *b4 = nullptr, *w5 = nullptr, *b5 = nullptr;  // beta / gamma folded
  bf16 *w1s = nullptr, *w3s = nullptr, *w4s = nullptr, *w5s = nullptr;  // c >= 128: [N][3K] bf16 hi|hi|lo for the split-precision tcgen05 GEMM
};
struct CrStnW {
  int k1 = 0, k2 = 0, n1 = 0, n2 = 0, fc = 0, hid = 0;
  float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;      // localisation convs, [Cout][k][k][Cin]
  float *f1 = nullptr, *fb1 = nullptr, *f2 = nullptr, *fb2 = nullptr;    // regressor, f1 columns in NHWC order
  __half* w1_mma = nullptr;                                              // first conv as scaled fp16 hi + lo 8x8 B matrices for edge::stn_conv_mma_kernel
  float w1_unscale = 1.f;                                                // 2^-e of that scale
};
struct CrStageW {
  int c = 0, res = 0, sampling = 0;  // 0 none, 1 down (2x2 s2 conv), 2 up (1x1 conv + PixelShuffle)
  std::vector<CrBlockW> blocks;
  CrStnW stn;
  float *samp_w = nullptr, *samp_b = nullptr;
};
struct CrW {
  bool loaded = false;
  int H = 128, cap = 32;
  float *intro_w = nullptr, *intro_b = nullptr, *outro_w = nullptr, *outro_b = nullptr;
  std::vector<CrStageW> stages;  // 4 encoders, middle, 4 decoders
  // workspace for one chunk of faces (all fp32)
  float *r[5] = {}, *sk[5] = {}, *ln_out = nullptr, *act_h = nullptr, *act_g = nullptr, *tmp = nullptr;
  float *pooled = nullptr, *sca_s = nullptr, *loc1 = nullptr, *loc2 = nullptr, *theta = nullptr, *stage = nullptr;
  float* stn_hidden = nullptr;          // [cap][96] hidden layer of the STN regressor
  unsigned int* stn_ticket = nullptr;   // [cap] block-completion counters of cr_stn_fc_kernel
  bf16* a3 = nullptr;  // split-precision A operand [rows][3K]
  bool use_tc = true;
  std::unordered_map<const float*, std::pair<float*, float*>> split_hl;  // fp32 weight -> tf32 hi / lo for gemm_mma3
  struct SplitH { __half *hi, *lo; float unscale; };
  std::unordered_map<const float*, SplitH> split_h;                      // fp32 weight -> scaled fp16 hi / lo for gemm_mma3h
};

struct HcaW {
  int d = 0, sp = 0;
  void* wf = nullptr;  // [d, 9d] fused 3x3, BN folded
  float* bf = nullptr;
  float *c0w = nullptr, *c0b = nullptr, *c2w = nullptr, *c2b = nullptr;  // channel_mlp
  float *s0w = nullptr, *s0b = nullptr, *s3w = nullptr, *s3b = nullptr;  // spatial_mlp, BN folded
  float *wc = nullptr, *ws = nullptr;                                    // per-face gates (set_condition)
};

struct TapInfo {
  const void* ptr = nullptr;
  int dtype = DT_F32, C = 0, HW = 0, ld = 0;
};

struct Op {
  std::function<void(cudaStream_t)> fn;
  std::string tap;
  TapInfo info;
  std::string label;  // kernel kind + shape, for hd_profile_step
};

thread_local std::string g_label;  // label picked up by the next add_op

struct TcLaunch;

struct Plan {
  int batch = 0;
  std::vector<Op> ops;
  cudaGraphExec_t graph = nullptr;  // one sampler step (plan + x_{t-1} update + advance), see hd_sample
  uint64_t graph_seed = 0;
  int64_t graph_first = 0;
  const float* graph_noise = nullptr;
  double flops_per_face = 0;
  std::shared_ptr<TcLaunch> last_tc;   // the previous tensor-core GEMM of the plan: it prefetches the next one's weights
  const void* first_w = nullptr;       // weights of the first such GEMM (prefetched by the last one: the plan repeats every step)
  unsigned int first_w_bytes = 0;
  int ending_idx = -1;  // index of the ending-conv op when hd_sample may replace it by the fused ending + scheduler-step kernel
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct SrcTensor {
  const void* data;
  int dtype;
  std::vector<int64_t> shape;
  size_t numel;
};

}  // namespace

struct hd_handle {
  hd_config cfg{};
  Tunables tun;
  std::string err;
  bool fused = false, bf16 = true, weights_loaded = false, condition_set = false;
  int S = 16, Bcap = 0, max_steps = 0, sm_count = 0, sm_major = 0, sm_minor = 0;
  int c[kNumLevels], sp[kNumLevels];
  int mod_stride = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  size_t act_bytes = 0, pooled_rows = 0;
  EncodeTiledFn encode = nullptr;
  Arena arena;
  DeviceStatus* d_status = nullptr;
  DeviceStatus* status_host = nullptr;  // pinned mirror of d_status, refreshed asynchronously after every enqueued call
  cudaEvent_t ev_status = nullptr;
  bool status_posted = false;

  // weights
  std::vector<BlockW> blocks;  // execution order
  HcaW hca[kNumLevels];
  void *down_w[4] = {}, *up_w[4] = {};
  float* down_b[4] = {};
  float *intro_w = nullptr, *intro_b = nullptr, *end_w = nullptr, *end_b = nullptr;
  hd::bf16 *intro_mma_hi = nullptr, *intro_mma_lo = nullptr, *end_mma_hi = nullptr, *end_mma_lo = nullptr;  // edge_convs.cuh fragment order
  unsigned int* end_ticket = nullptr;
  float *tm1_w = nullptr, *tm1_b = nullptr, *tm3_w = nullptr, *tm3_b = nullptr, *mlp_w = nullptr, *mlp_b = nullptr;
  float *idc_w = nullptr, *idc_b = nullptr, *freqs = nullptr;
  int64_t weight_elems_step = 0;

  // workspace
  float* resid[kNumLevels] = {};
  void *act_a = nullptr, *act_h = nullptr, *act_g = nullptr, *hca_out = nullptr, *pooled = nullptr;
  float *sca_s = nullptr, *gate_tmp = nullptr;
  float *x_state = nullptr, *eps_buf = nullptr, *x_stage = nullptr;
  float *t_vals = nullptr, *t_emb = nullptr, *t_h1 = nullptr, *t_g1 = nullptr, *t_temb = nullptr, *t_g2 = nullptr,
        *mod_table = nullptr;
  int* row_idx = nullptr;
  StepCoef* coefs = nullptr;
  StepState* state = nullptr;
  float *idc_add = nullptr, *cond_nhwc = nullptr, *cond_pool = nullptr, *cond_h = nullptr, *cond_hs = nullptr,
        *cond_stage = nullptr;
  std::vector<float> table_key;  // timesteps currently held by mod_table rows
  const float* cur_x = nullptr;
  float* cur_eps = nullptr;
  std::map<int, std::unique_ptr<Plan>> plans;      // fast plans (persistent chain kernel where enabled)
  std::map<int, std::unique_ptr<Plan>> plans_dbg;  // one kernel per op: per-layer taps
  FpgW fpg;
  std::map<int, std::unique_ptr<Plan>> fpg_plans;
  const float* fpg_in = nullptr;
  IdcW idc;
  std::map<int, std::unique_ptr<Plan>> idc_plans;
  const float* idc_in = nullptr;
  float* idc_out = nullptr;
  CrW cr;
  std::map<int, std::unique_ptr<Plan>> cr_plans;
  const float* cr_in = nullptr;
  float* cr_out = nullptr;
  size_t workspace_bytes = 0;

  // transient during load
  std::map<std::string, SrcTensor> src;
  std::vector<void*> temp_dev;
};

namespace {

size_t esize(int dt) { return dt == DT_BF16 ? 2 : 4; }

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
