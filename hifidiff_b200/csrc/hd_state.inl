// hd_state.inl — part of hd_lib.cu (one translation unit; included there in order, not compiled on its own).
// Library state: error helpers, per-handle tuning switches, the PDL launch helper, the device arena, packed-weight records of every network, the per-batch launch plan and hd_handle itself.
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

