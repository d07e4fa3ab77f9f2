// libsamaudio_b200.so — engine + C ABI (include/samaudio_b200.h).
// Weights are repacked to bf16 GEMM operands at load; every (Bc, T, L) shape gets a static plan
// (workspace + TMA descriptors + launch records) that each ODE evaluation replays.
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <functional>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <array>

#include "../../include/samaudio_b200.h"
#include "host_util.h"
#include "attention.cuh"
#include "attention_tc.cuh"
#include "attention_tc2.cuh"
#include "elementwise.cuh"
#include "codec_kernels.cuh"
#include "vision_kernels.cuh"

using namespace sab;
typedef __nv_bfloat16 bf16;

static thread_local std::string g_last_error;

// ---- per-device state (several engines on different GPUs may live in one process) ----
constexpr int kMaxDevices = 64;
static int cur_device() {
  int d = 0;
  SAB_CUDA(cudaGetDevice(&d));
  SAB_CHECK(d >= 0 && d < kMaxDevices, "device index %d out of range", d);
  return d;
}
static int sm_count() {   // of the current device
  static int cached[kMaxDevices] = {0};
  const int d = cur_device();
  if (!cached[d]) {
    cudaDeviceProp prop;
    SAB_CUDA(cudaGetDeviceProperties(&prop, d));
    cached[d] = prop.multiProcessorCount;
  }
  return cached[d];
}
// cudaFuncSetAttribute is per device: once per (kernel, device).  Keyed by the function ADDRESS: all instantiations
// of one kernel template share a function-pointer type, so a per-type static would configure only the first one.
static void ensure_dynamic_smem(const void* kern, int smem) {
  static std::unordered_map<const void*, std::array<bool, kMaxDevices>> configured;
  const int d = cur_device();
  auto& flags = configured[kern];   // value-initialised to all false on first use
  if (!flags[d]) {
    SAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    flags[d] = true;
  }
}
// every entry point runs on its engine's device whatever the caller's current device is, and restores it
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    int cur = 0;
    SAB_CUDA(cudaGetDevice(&cur));
    if (cur != dev) { SAB_CUDA(cudaSetDevice(dev)); prev = cur; }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// =====================================================================================================
// GEMM dispatch
// =====================================================================================================
template <int BN, int BK, int MODE, int CG, bool B2B = false>
static void launch_gemm_inst(const GemmOp& op, cudaStream_t st) {
  auto kern = gemm_tc_kernel<BN, BK, MODE, CG, B2B>;
  constexpr int smem = GemmSmem<BN, BK, CG, B2B>::kTotal;
  ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
  if (CG == 1) {
    kern<<<op.grid, GEMM_THREADS, smem, st>>>(op.tmA, op.tmB, op.b2b ? op.tmW : op.tmB, op.P);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(op.grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SAB_CUDA(cudaLaunchKernelEx(&cfg, kern, op.tmA, op.tmB, op.tmB, op.P));
  }
  SAB_CUDA(cudaGetLastError());
}

static void launch_gemm(const GemmOp& op, cudaStream_t st) {
  if (op.b2b) {
    if (op.BN == 64 && op.BK == 64) return launch_gemm_inst<64, 64, EPI_AFFINE, 1, true>(op, st);
    if (op.BN == 96 && op.BK == 32) return launch_gemm_inst<96, 32, EPI_AFFINE, 1, true>(op, st);
    if (op.BN == 128 && op.BK == 64) return launch_gemm_inst<128, 64, EPI_AFFINE, 1, true>(op, st);
    throw Error(fmt("no back-to-back GEMM instantiation for BN=%d BK=%d (%s)", op.BN, op.BK, op.tag));
  }
#define SAB_CASE(bn_, bk_, md_, cg_) \
  if (op.BN == bn_ && op.BK == bk_ && op.mode == md_ && op.cg == cg_) return launch_gemm_inst<bn_, bk_, md_, cg_>(op, st);
  SAB_CASE(256, 64, EPI_AFFINE, 2)
  SAB_CASE(256, 64, EPI_AFFINE_NORM, 2)
  SAB_CASE(256, 64, EPI_AFFINE_NORM, 1)
  SAB_CASE(256, 64, EPI_SWIGLU, 2)
  SAB_CASE(256, 64, EPI_QKV, 2)
  SAB_CASE(256, 64, EPI_AFFINE, 1)
  SAB_CASE(192, 64, EPI_AFFINE, 1)
  SAB_CASE(128, 64, EPI_AFFINE, 1)
  SAB_CASE(96, 64, EPI_AFFINE, 1)
  SAB_CASE(64, 64, EPI_AFFINE, 1)
  SAB_CASE(96, 32, EPI_AFFINE, 1)
  SAB_CASE(128, 32, EPI_AFFINE, 1)
  SAB_CASE(256, 64, EPI_SWIGLU, 1)
  SAB_CASE(256, 64, EPI_QKV, 1)
  SAB_CASE(128, 64, EPI_QKV, 1)
#undef SAB_CASE
  throw Error(fmt("no GEMM instantiation for BN=%d BK=%d mode=%d cg=%d (%s)", op.BN, op.BK, op.mode, op.cg, op.tag));
}

// A operand view
struct AView {
  const bf16* ptr;
  int64_t cols;        // row width visible to TMA (elements)
  int64_t rows;        // rows per item
  int64_t items;
  int64_t row_pitch;   // elements
  int64_t item_pitch;  // elements
};
static AView flat_view(const bf16* p, int64_t M, int64_t K, int64_t ld = -1) {
  return AView{p, K, M, 1, ld < 0 ? K : ld, 0};
}
static AView seq_view(const bf16* p, int64_t items, int64_t T, int64_t C) { return AView{p, C, T, items, C, T * C}; }

struct RunList {
  int n[2] = {0, 0};
  KRun r[2][GEMM_MAX_RUNS];
  int period = 0, sw = 0;
  void add(int list, int shift, int col, int nkb) {
    SAB_CHECK(n[list] < GEMM_MAX_RUNS, "too many K runs");
    r[list][n[list]++] = KRun{shift, col, nkb};
  }
  int total_kb(int list) const {
    int t = 0;
    for (int i = 0; i < n[list]; ++i) t += r[list][i].nkb;
    return t;
  }
};

static int g_force_cg = 0;   // SAB_FORCE_CG=1|2 overrides the CTA-group choice (A/B measurements)

static GemmOp make_gemm(const char* tag, const AView& A, const bf16* B, int N, int BN, int BK, int mode,
                        const RunList& runs, int cg = 0) {
  GemmOp op;
  memset(&op.P, 0, sizeof(op.P));
  op.tag = tag;
  op.BN = BN; op.BK = BK; op.mode = mode;
  // 2-CTA pairs (cta_group::2, 256 x BN tiles) for the big tensor-bound GEMMs
  if (cg == 0) cg = (BN == 256 && BK == 64 && A.rows * A.items >= 1024 && A.rows >= 200) ? 2 : 1;
  if (g_force_cg && BN == 256 && BK == 64) cg = g_force_cg;
  op.cg = cg;
  const int ktot = runs.total_kb(0) * BK;
  if (runs.period > 0) SAB_CHECK(runs.total_kb(1) * BK == ktot, "%s: run lists differ in K", tag);
  op.tmA = make_tmap_3d(A.ptr, A.cols, A.rows, A.items, A.row_pitch, A.item_pitch == 0 ? A.rows * A.row_pitch : A.item_pitch,
                        BK, GEMM_BM);
  op.tmB = make_tmap_2d(B, ktot, N, ktot, BK, BN / cg);
  GemmParams& P = op.P;
  P.rows_per_item = (int)A.rows;
  P.n_items = (int)A.items;
  P.tiles_per_item = (int)((A.rows + GEMM_BM * cg - 1) / (GEMM_BM * cg));
  P.N = N;
  P.n_tiles_n = (N + BN - 1) / BN;
  // rasterisation: keep one group's A panel (group_m x 128*cg x K bf16) around 24 MB so it stays L2-resident
  {
    const double panel_bytes = (double)GEMM_BM * cg * ktot * 2.0;
    int gm = (int)(24.0e6 / panel_bytes);
    gm = gm < 4 ? 4 : (gm > 32 ? 32 : gm);
    P.group_m = gm;
  }
  for (int l = 0; l < 2; ++l) {
    P.n_runs[l] = runs.n[l];
    for (int i = 0; i < runs.n[l]; ++i) P.runs[l][i] = runs.r[l][i];
  }
  P.n_period = runs.period;
  P.n_switch = runs.sw;
  if (runs.period > 0) SAB_CHECK(runs.sw % BN == 0 && runs.period % BN == 0, "%s: BN=%d does not divide the tap switch %d/%d", tag, BN, runs.sw, runs.period);
  P.alpha = 1.f;
  P.gate_div = 1;
  P.eps = 1e-5f;
  const long long tiles = (long long)P.n_items * P.tiles_per_item * P.n_tiles_n;
  op.grid = (int)std::min<long long>(tiles, sm_count() / cg) * cg;
  op.flops = 2.0 * (double)A.rows * (double)A.items * (double)N * (double)ktot;
  op.rows = (double)A.rows * (double)A.items;
  op.in_bytes = op.rows * (double)A.cols * 2.0 + (double)N * (double)ktot * 2.0;
  SAB_CHECK(N % 32 == 0, "%s: N=%d must be a multiple of 32", tag, N);
  return op;
}
static GemmOp make_linear(const char* tag, const bf16* A, int64_t M, int K, const bf16* B, int N, int BN, int mode,
                          int64_t lda = -1) {
  SAB_CHECK(K % 64 == 0, "%s: K=%d must be a multiple of 64", tag, K);
  RunList rl;
  rl.add(0, 0, 0, K / 64);
  return make_gemm(tag, flat_view(A, M, K, lda), B, N, BN, 64, mode, rl);
}

// =====================================================================================================
// weight packing kernels
// =====================================================================================================
enum RowMap { ROW_IDENT = 0, ROW_HEADS = 1, ROW_SWIGLU_A = 2, ROW_SWIGLU_B = 3 };

// dst[(row0 + map(r)) * ld + col0 + c] = bf16(src[r*rs + c*cs])
__global__ void pack_bf16_kernel(bf16* __restrict__ dst, long long ld, long long row0, long long col0,
                                 const float* __restrict__ src, long long R, long long C, long long rs, long long cs,
                                 int rowmap, int H) {
  const long long n = R * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C, c = i % C;
    long long pr = r;
    if (rowmap == ROW_HEADS) { const long long dd = r / H, h = r % H; pr = h * (R / H) + dd; }
    else if (rowmap == ROW_SWIGLU_A) pr = (r >> 5) * 64 + (r & 31);
    else if (rowmap == ROW_SWIGLU_B) pr = (r >> 5) * 64 + 32 + (r & 31);
    dst[(row0 + pr) * ld + col0 + c] = __float2bfloat16(src[r * rs + c * cs]);
  }
}
static void pack(bf16* dst, long long ld, long long row0, long long col0, const float* src, long long R, long long C,
                 long long rs, long long cs, int rowmap, int H, cudaStream_t st) {
  const long long n = R * C;
  const int grid = (int)std::min<long long>((n + 255) / 256, 4096);
  pack_bf16_kernel<<<grid, 256, 0, st>>>(dst, ld, row0, col0, src, R, C, rs, cs, rowmap, H);
  SAB_CUDA(cudaGetLastError());
}

// =====================================================================================================
// engine state
// =====================================================================================================
struct Slot {
  std::vector<int64_t> shape;
  std::function<void(const float*, cudaStream_t)> load;
  bool loaded = false;
};

struct LayerW {
  bf16 *wqkv, *wo, *wq_c, *wkv_c, *wo_c, *w13, *w2;
  float *qn, *kn, *qn_c, *kn_c, *attn_norm, *ffn_norm;
  float* qn_scaled = nullptr;   // q_norm.weight * log2(e)/sqrt(hd): the QKV epilogue emits q in the softmax's log2 domain
  // self-attention logit bound (log2 domain) from the QK-norm weights: |q.k| / sqrt(hd) <= sqrt(hd) max|w_q| max|w_k|
  // (RMSNorm'd vectors have norm <= sqrt(hd) max|w|, RoPE preserves norms); < 0: unknown -> exact two-pass softmax
  float att_shift_log2 = -1.f;
};

struct ConvLayer {   // a multi-channel codec conv lowered to the GEMM
  bf16* w = nullptr;
  float* bias = nullptr;
  int cin = 0, cout = 0, k = 0, stride = 1, dil = 1;
  bool transposed = false;
};
struct ResUnitW {
  float* a0; ConvLayer c7; float* a1; ConvLayer c1;
};
struct EncBlockW { ResUnitW ru[3]; float* a_down; ConvLayer down; int cin, cout, stride; };
struct DecBlockW { float* a_up; ConvLayer up; ResUnitW ru[3]; int cin, cout, stride; };

struct DitPlan;
struct CodecPlan;
struct ProfRec { const char* tag; double flops; double bytes; };

struct sab_engine {
  sab_config cfg;
  int device = 0;
  DevicePool wpool;
  std::unordered_map<std::string, Slot> slots;
  float* staging = nullptr;
  int64_t staging_elems = 0;
  bool finalized = false;
  int64_t launches = 0;
  cudaStream_t capture_stream = nullptr;   // private stream for graph capture (the caller's may be the NULL stream)
  bool prof = false;
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_events;

  // --- DiT weights ---
  std::vector<LayerW> layers;
  float* tables = nullptr;          // [L, 6, d]
  bf16* wkv_c_all = nullptr;        // [L, 2d, d]
  float* kn_c_all = nullptr;        // [L, 128]
  bf16 *wp_y, *wp_f, *wmem, *wvid;  // proj (noisy third / feature third), memory_proj, align conv
  float *proj_b, *mem_b, *vid_b, *vid_ln_w, *vid_ln_b, *vid_gate, *vid_const;
  float *anchor_embed, *anchor_proj, *anchor_gate, *anchor_table;
  bf16 *xe_w[2]; float *xe_b[2], *xe_gn_w[2], *xe_gn_b[2];
  bf16 *t_w13, *t_w2, *tb_w, *y_w13, *y_w2, *w_out;
  float *tb_b, *final_norm, *final_table;
  float* norm_w_all = nullptr;      // [2L, d]: attention_norm / ffn_norm weights of every layer, contiguous (finalize)
  float2* rope = nullptr;
  int rope_len = 0;

  // --- codec weights ---
  float *enc0_w, *enc0_b;                 // Conv1d(1, C0, 7)
  std::vector<EncBlockW> enc_blocks;
  float* enc_final_alpha; ConvLayer enc_final;   // Snake + Conv k3
  ConvLayer in_proj, out_proj;            // in_proj packed as [2*cz, latent] (mean rows duplicated)
  ConvLayer dec0;                         // Conv k7 latent -> decoder_dim
  std::vector<DecBlockW> dec_blocks;
  float* dec_final_alpha; float *dec_last_w, *dec_last_b;  // Snake + Conv1d(C,1,7) + tanh

  std::unique_ptr<DitPlan> dit;
  std::map<std::pair<int, long long>, std::unique_ptr<CodecPlan>> enc_plans, dec_plans;
  int64_t plan_clock = 0;

  ~sab_engine();
};

// Every kernel launch goes through mark(): it counts the launch and, when profiling is on, drops one event
// in front of it; the time between consecutive events is attributed to the launch that follows the first.
static void mark(sab_engine* e, cudaStream_t st, const char* tag, double flops = 0, double bytes = 0) {
  e->launches++;
  if (!e->prof) return;
  if (e->prof_recs.size() >= e->prof_events.size()) {
    cudaEvent_t ev;
    SAB_CUDA(cudaEventCreate(&ev));
    e->prof_events.push_back(ev);
  }
  SAB_CUDA(cudaEventRecord(e->prof_events[e->prof_recs.size()], st));
  e->prof_recs.push_back(ProfRec{tag, flops, bytes});
}

// =====================================================================================================
// weight registry
// =====================================================================================================
static void reg(sab_engine* e, const std::string& name, std::vector<int64_t> shape,
                std::function<void(const float*, cudaStream_t)> fn) {
  Slot s;
  s.shape = std::move(shape);
  s.load = std::move(fn);
  e->slots[name] = std::move(s);
}
static void reg_f32(sab_engine* e, const std::string& name, std::vector<int64_t> shape, float** dst) {
  int64_t n = 1;
  for (auto v : shape) n *= v;
  *dst = e->wpool.alloc<float>(n, true);
  float* d = *dst;
  reg(e, name, shape, [d, n](const float* src, cudaStream_t st) {
    SAB_CUDA(cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  });
}
// plain [out, in] linear -> bf16 [out, in] (optionally a column window of the source)
static void reg_linear(sab_engine* e, const std::string& name, int out, int in, bf16** dst, int rowmap = ROW_IDENT,
                       int H = 1, bf16* into = nullptr, long long row0 = 0, long long ld = -1, int src_col0 = 0,
                       int src_cols = -1) {
  const int cols = src_cols < 0 ? in : src_cols;
  if (!into) { *dst = e->wpool.alloc<bf16>((int64_t)out * cols, true); into = *dst; }
  if (ld < 0) ld = cols;
  reg(e, name, {out, in}, [=](const float* src, cudaStream_t st) {
    pack(into, ld, row0, 0, src + src_col0, out, cols, in, 1, rowmap, H, st);
  });
}
// Conv1d weight [co, ci, k] -> bf16 [co, k*ci] (tap-major K)
static void reg_conv(sab_engine* e, const std::string& name, ConvLayer* L, int co, int ci, int k, int stride, int dil) {
  L->cin = ci; L->cout = co; L->k = k; L->stride = stride; L->dil = dil; L->transposed = false;
  L->w = e->wpool.alloc<bf16>((int64_t)co * ci * k, true);
  bf16* w = L->w;
  reg(e, name + ".weight", {co, ci, k}, [=](const float* src, cudaStream_t st) {
    for (int kk = 0; kk < k; ++kk) pack(w, (long long)k * ci, 0, (long long)kk * ci, src + kk, co, ci, (long long)ci * k, k, ROW_IDENT, 1, st);
  });
  reg_f32(e, name + ".bias", {co}, &L->bias);
}
// ConvTranspose1d weight [ci, co, 2s] (stride s, padding s/2) -> bf16 [s*co, 2*ci]:
// output sample r*s+u is x[r] * W[:, :, u+p] + (u < s/2 ? x[r-1] * W[:, :, u+p+s] : x[r+1] * W[:, :, u+p-s])
static void reg_convT(sab_engine* e, const std::string& name, ConvLayer* L, int ci, int co, int s) {
  SAB_CHECK(s % 2 == 0, "codec rates must be even (got %d)", s);
  L->cin = ci; L->cout = co; L->k = 2 * s; L->stride = s; L->dil = 1; L->transposed = true;
  L->w = e->wpool.alloc<bf16>((int64_t)s * co * 2 * ci, true);
  bf16* w = L->w;
  const int K = 2 * s, p = s / 2;
  reg(e, name + ".weight", {ci, co, K}, [=](const float* src, cudaStream_t st) {
    for (int u = 0; u < s; ++u) {
      const int k0 = u + p;
      const int k1 = (u < s / 2) ? u + p + s : u + p - s;
      pack(w, 2LL * ci, (long long)u * co, 0, src + k0, co, ci, K, (long long)co * K, ROW_IDENT, 1, st);
      pack(w, 2LL * ci, (long long)u * co, ci, src + k1, co, ci, K, (long long)co * K, ROW_IDENT, 1, st);
    }
  });
  reg_f32(e, name + ".bias", {co}, &L->bias);
}
static void reg_resunit(sab_engine* e, const std::string& p, ResUnitW* R, int c, int dil) {
  reg_f32(e, p + ".block.0.alpha", {1, c, 1}, &R->a0);
  reg_conv(e, p + ".block.1", &R->c7, c, c, 7, 1, dil);
  reg_f32(e, p + ".block.2.alpha", {1, c, 1}, &R->a1);
  reg_conv(e, p + ".block.3", &R->c1, c, c, 1, 1, 1);
}

static void register_weights(sab_engine* e) {
  const sab_config& c = e->cfg;
  const int d = c.dim, H = c.n_heads, hid = c.ffn_hidden, L = c.n_layers;
  auto S = [](const char* f, int i) { return fmt(f, i); };
  // ---- conditioning ----
  e->wp_y = e->wpool.alloc<bf16>((int64_t)d * 256, true);
  e->wp_f = e->wpool.alloc<bf16>((int64_t)d * 256, true);
  SAB_CHECK(c.in_channels == 768 && c.out_channels == 256, "in_channels/out_channels must be 768/256");
  {
    bf16 *wy = e->wp_y, *wf = e->wp_f;
    const int in = c.in_channels;
    reg(e, "proj.weight", {d, in}, [=](const float* src, cudaStream_t st) {
      pack(wy, 256, 0, 0, src, d, 256, in, 1, ROW_IDENT, 1, st);         // multiplies the noisy latent
      pack(wf, 256, 0, 0, src + 512, d, 256, in, 1, ROW_IDENT, 1, st);   // multiplies the mixture features
    });
  }
  reg_f32(e, "proj.bias", {d}, &e->proj_b);
  reg_linear(e, "memory_proj.weight", d, c.text_dim, &e->wmem);
  reg_f32(e, "memory_proj.bias", {d}, &e->mem_b);
  e->wvid = e->wpool.alloc<bf16>((int64_t)d * c.vision_dim, true);
  {
    bf16* wv = e->wvid;
    const int vd = c.vision_dim;
    reg(e, "align_masked_video.conv.weight", {d, vd, 1}, [=](const float* src, cudaStream_t st) {
      pack(wv, vd, 0, 0, src, d, vd, vd, 1, ROW_IDENT, 1, st);
    });
  }
  reg_f32(e, "align_masked_video.conv.bias", {d}, &e->vid_b);
  reg_f32(e, "align_masked_video.layer_norm.weight", {d}, &e->vid_ln_w);
  reg_f32(e, "align_masked_video.layer_norm.bias", {d}, &e->vid_ln_b);
  reg_f32(e, "align_masked_video.gate", {1}, &e->vid_gate);
  reg_f32(e, "embed_anchors.embed.weight", {c.n_anchor_tokens, c.anchor_dim}, &e->anchor_embed);
  reg_f32(e, "embed_anchors.proj.weight", {d, c.anchor_dim}, &e->anchor_proj);
  reg_f32(e, "embed_anchors.gate", {1}, &e->anchor_gate);
  e->vid_const = e->wpool.alloc<float>(d, true);
  e->anchor_table = e->wpool.alloc<float>((int64_t)c.n_anchor_tokens * d, true);
  // ---- DiT prologue / epilogue ----
  for (int b = 0; b < 2; ++b) {
    const std::string p = fmt("transformer.x_embedder.block.block%d", b + 1);
    reg_f32(e, p + ".groupnorm.weight", {d}, &e->xe_gn_w[b]);
    reg_f32(e, p + ".groupnorm.bias", {d}, &e->xe_gn_b[b]);
    e->xe_w[b] = e->wpool.alloc<bf16>((int64_t)d * 3 * d, true);
    bf16* w = e->xe_w[b];
    reg(e, p + ".project.weight", {d, d, 3}, [=](const float* src, cudaStream_t st) {
      for (int k = 0; k < 3; ++k) pack(w, 3LL * d, 0, (long long)k * d, src + k, d, d, 3LL * d, 3, ROW_IDENT, 1, st);
    });
    reg_f32(e, p + ".project.bias", {d}, &e->xe_b[b]);
  }
  e->t_w13 = e->wpool.alloc<bf16>((int64_t)2 * d * 256, true);
  reg_linear(e, "transformer.t_embedder.projection.w1.weight", d, 256, nullptr, ROW_SWIGLU_A, 1, e->t_w13, 0, 256);
  reg_linear(e, "transformer.t_embedder.projection.w3.weight", d, 256, nullptr, ROW_SWIGLU_B, 1, e->t_w13, 0, 256);
  reg_linear(e, "transformer.t_embedder.projection.w2.weight", d, d, &e->t_w2);
  reg_linear(e, "transformer.t_block.weight", 6 * d, d, &e->tb_w);
  reg_f32(e, "transformer.t_block.bias", {6 * d}, &e->tb_b);
  e->y_w13 = e->wpool.alloc<bf16>((int64_t)2 * d * d, true);
  reg_linear(e, "transformer.y_embedder.projection.w1.weight", d, d, nullptr, ROW_SWIGLU_A, 1, e->y_w13, 0, d);
  reg_linear(e, "transformer.y_embedder.projection.w3.weight", d, d, nullptr, ROW_SWIGLU_B, 1, e->y_w13, 0, d);
  reg_linear(e, "transformer.y_embedder.projection.w2.weight", d, d, &e->y_w2);
  reg_f32(e, "transformer.norm.weight", {d}, &e->final_norm);
  reg_linear(e, "transformer.output.weight", c.out_channels, d, &e->w_out);
  reg_f32(e, "transformer.final_layer_scale_shift_table", {2, d}, &e->final_table);
  // ---- DiT layers ----
  e->norm_w_all = e->wpool.alloc<float>((int64_t)2 * L * d, true);
  e->tables = e->wpool.alloc<float>((int64_t)L * 6 * d, true);
  e->wkv_c_all = e->wpool.alloc<bf16>((int64_t)L * 2 * d * d, true);   // all layers' cross wk|wv: one GEMM per evaluation
  e->kn_c_all = e->wpool.alloc<float>((int64_t)L * 128, true);
  e->layers.resize(L);
  for (int l = 0; l < L; ++l) {
    LayerW& W = e->layers[l];
    const std::string p = S("transformer.layers.%d", l);
    W.wqkv = e->wpool.alloc<bf16>((int64_t)3 * d * d, true);
    reg_linear(e, p + ".attention.wq.weight", d, d, nullptr, ROW_HEADS, H, W.wqkv, 0, d);
    reg_linear(e, p + ".attention.wk.weight", d, d, nullptr, ROW_HEADS, H, W.wqkv, d, d);
    reg_linear(e, p + ".attention.wv.weight", d, d, nullptr, ROW_HEADS, H, W.wqkv, 2 * d, d);
    reg_linear(e, p + ".attention.wo.weight", d, d, &W.wo);
    reg_f32(e, p + ".attention.q_norm.weight", {128}, &W.qn);
    W.qn_scaled = e->wpool.alloc<float>(128, true);
    reg_f32(e, p + ".attention.k_norm.weight", {128}, &W.kn);
    reg_linear(e, p + ".cross_attention.wq.weight", d, d, &W.wq_c, ROW_HEADS, H);
    W.wkv_c = e->wkv_c_all + (int64_t)l * 2 * d * d;
    reg_linear(e, p + ".cross_attention.wk.weight", d, d, nullptr, ROW_HEADS, H, W.wkv_c, 0, d);
    reg_linear(e, p + ".cross_attention.wv.weight", d, d, nullptr, ROW_HEADS, H, W.wkv_c, d, d);
    reg_linear(e, p + ".cross_attention.wo.weight", d, d, &W.wo_c);
    reg_f32(e, p + ".cross_attention.q_norm.weight", {128}, &W.qn_c);
    W.kn_c = e->kn_c_all + (int64_t)l * 128;
    {
      float* dst = W.kn_c;
      reg(e, p + ".cross_attention.k_norm.weight", {128}, [dst](const float* src, cudaStream_t st) {
        SAB_CUDA(cudaMemcpyAsync(dst, src, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
      });
    }
    W.w13 = e->wpool.alloc<bf16>((int64_t)2 * hid * d, true);
    reg_linear(e, p + ".feed_forward.w1.weight", hid, d, nullptr, ROW_SWIGLU_A, 1, W.w13, 0, d);
    reg_linear(e, p + ".feed_forward.w3.weight", hid, d, nullptr, ROW_SWIGLU_B, 1, W.w13, 0, d);
    reg_linear(e, p + ".feed_forward.w2.weight", d, hid, &W.w2);
    reg_f32(e, p + ".attention_norm.weight", {d}, &W.attn_norm);
    reg_f32(e, p + ".ffn_norm.weight", {d}, &W.ffn_norm);
    float* tab = e->tables + (int64_t)l * 6 * d;
    reg(e, p + ".scale_shift_table", {6, d}, [=](const float* src, cudaStream_t st) {
      SAB_CUDA(cudaMemcpyAsync(tab, src, 6LL * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    });
  }
  // ---- codec ----
  const std::string E = "audio_codec.encoder", D = "audio_codec.decoder", Q = "audio_codec.quantizer";
  const int nr = c.codec_n_rates;
  int ch = c.codec_encoder_dim;
  SAB_CHECK(ch % 64 == 0, "codec encoder_dim must be a multiple of 64");
  reg_f32(e, E + ".block.0.weight", {ch, 1, 7}, &e->enc0_w);
  reg_f32(e, E + ".block.0.bias", {ch}, &e->enc0_b);
  e->enc_blocks.resize(nr);
  for (int i = 0; i < nr; ++i) {
    EncBlockW& B = e->enc_blocks[i];
    const int s = c.codec_encoder_rates[i];
    SAB_CHECK(s % 2 == 0, "codec rates must be even");
    B.cin = ch; B.cout = 2 * ch; B.stride = s;
    const std::string p = fmt("%s.block.%d", E.c_str(), i + 1);
    const int dil[3] = {1, 3, 9};
    for (int j = 0; j < 3; ++j) reg_resunit(e, fmt("%s.block.%d", p.c_str(), j), &B.ru[j], ch, dil[j]);
    reg_f32(e, p + ".block.3.alpha", {1, ch, 1}, &B.a_down);
    reg_conv(e, p + ".block.4", &B.down, 2 * ch, ch, 2 * s, s, 1);
    ch *= 2;
  }
  reg_f32(e, fmt("%s.block.%d.alpha", E.c_str(), nr + 1), {1, ch, 1}, &e->enc_final_alpha);
  reg_conv(e, fmt("%s.block.%d", E.c_str(), nr + 2), &e->enc_final, c.codec_latent_dim, ch, 3, 1, 1);
  {
    // in_proj: only the mean half is used (codec.py:68); packed twice so one GEMM writes [.., 2*cz]
    const int cz = c.codec_codebook_dim, ld = c.codec_latent_dim;
    ConvLayer& P = e->in_proj;
    P.cin = ld; P.cout = 2 * cz; P.k = 1;
    P.w = e->wpool.alloc<bf16>((int64_t)2 * cz * ld, true);
    P.bias = e->wpool.alloc<float>(2 * cz, true);
    bf16* w = P.w; float* bptr = P.bias;
    reg(e, Q + ".in_proj.weight", {2 * cz, ld, 1}, [=](const float* src, cudaStream_t st) {
      pack(w, ld, 0, 0, src, cz, ld, ld, 1, ROW_IDENT, 1, st);
      pack(w, ld, cz, 0, src, cz, ld, ld, 1, ROW_IDENT, 1, st);
    });
    reg(e, Q + ".in_proj.bias", {2 * cz}, [=](const float* src, cudaStream_t st) {
      SAB_CUDA(cudaMemcpyAsync(bptr, src, cz * sizeof(float), cudaMemcpyDeviceToDevice, st));
      SAB_CUDA(cudaMemcpyAsync(bptr + cz, src, cz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    });
    reg_conv(e, Q + ".out_proj", &e->out_proj, ld, cz, 1, 1, 1);
  }
  int dch = c.codec_decoder_dim;
  reg_conv(e, D + ".model.0", &e->dec0, dch, c.codec_latent_dim, 7, 1, 1);
  e->dec_blocks.resize(nr);
  for (int i = 0; i < nr; ++i) {
    DecBlockW& B = e->dec_blocks[i];
    const int s = c.codec_decoder_rates[i];
    B.cin = dch; B.cout = dch / 2; B.stride = s;
    const std::string p = fmt("%s.model.%d", D.c_str(), i + 1);
    reg_f32(e, p + ".block.0.alpha", {1, dch, 1}, &B.a_up);
    reg_convT(e, p + ".block.1", &B.up, dch, dch / 2, s);
    const int dil[3] = {1, 3, 9};
    for (int j = 0; j < 3; ++j) reg_resunit(e, fmt("%s.block.%d", p.c_str(), j + 2), &B.ru[j], dch / 2, dil[j]);
    dch /= 2;
  }
  reg_f32(e, fmt("%s.model.%d.alpha", D.c_str(), nr + 1), {1, dch, 1}, &e->dec_final_alpha);
  reg_f32(e, fmt("%s.model.%d.weight", D.c_str(), nr + 2), {1, dch, 7}, &e->dec_last_w);
  reg_f32(e, fmt("%s.model.%d.bias", D.c_str(), nr + 2), {1}, &e->dec_last_b);
}

// =====================================================================================================
// DiT plan: workspace + launch records for one (Bc, T, L)
// =====================================================================================================
struct LayerOps {
  GemmOp qkv, wo, q_c, kv_c, wo_c, w13, w2;
};
// Everything of a DiT evaluation that depends on the evaluation TIME only (transformer.py:482-493, 363-371, 507-509):
// the timestep embedding, t_block, the adaLN tables of every layer, and — for the fused RMSNorm — the column scales
// w*(1+scale) and the biases shift @ W^T.  R rows = one per sequence (sab_dit_forward: every sequence has its own time)
// or one per evaluation of a solve (sab_solve: all sequences share the time, so the 2*n_steps rows are computed ONCE
// per plan instead of inside the ODE loop, and every consumer reads row k with an item stride of 0).
struct TimeState {
  int R = 0;
  float *time = nullptr, *t = nullptr, *t0 = nullptr, *mod = nullptr, *fin = nullptr, *cs = nullptr;
  float *bias_qkv = nullptr, *bias_w13 = nullptr;
  bf16 *tfreq = nullptr, *t_h = nullptr, *t_silu = nullptr, *shift = nullptr;
  GemmOp g_t13, g_t2, g_tb;
  std::vector<GemmOp> g_bias_qkv, g_bias_w13;
  bool valid = false;      // hoisted state: tables match the current weights and step count
  int n_steps = 0;
};

struct DitPlan {
  int B = 0, cand = 1;                  // clips and candidates per clip: Bc = B * cand sequences (candidate-minor)
  int Bc = 0, T = 0, L = 0;
  int64_t M = 0, MB = 0, ML = 0;        // sequence rows Bc*T, clip rows B*T, clip text rows B*L
  float* condB = nullptr;               // clip-level conditioning GEMM output (aliases cond when cand == 1)
  // fused RMSNorm + modulate (gemm_tc.cuh): column scales w*(1+scale) [2NL, Bc, d], bf16 shifts [2NL, Bc, d],
  // per-(row, n-tile, warp half) partial sums of squares [M, ssq_n], per-item biases shift @ W^T
  bool fused_norm = false;
  float* nrm_ssq = nullptr;
  int ssq_n = 0;
  TimeState ts_item;     // per-sequence times (sab_dit_forward), recomputed every evaluation
  TimeState ts_solve;    // one row per evaluation of the ODE grid (sab_solve), computed once
  DevicePool pool;
  // activations
  float *y, *ymid, *cond, *x0, *c1, *h, *mem_base, *time_dev, *vproj;
  bf16 *y_bf, *gn_a, *hb, *xn, *qkv, *att, *qc, *kvc, *u, *mem_in, *y_h, *ymem, *feat_bf, *text_bf, *vid_bf;
  double* gn_partial;
  uint8_t *pad_mask, *text_mask;
  long long *anchor_ids, *anchor_align;
  int n_ids_cap = 0;
  // ops
  GemmOp g_y13, g_y2, g_in, g_xe[2], g_out, g_cond, g_mem, g_vid, g_kvc_all;
  std::vector<LayerOps> lay;
  CUtensorMap tm_att_q, tm_att_kv;   // fused-QKV buffer viewed as (3d cols, T rows, Bc items): box 64x128 / 64x256
  CUtensorMap tm_att_o;              // attention output viewed as (d cols, T rows, Bc items): box 64x128 (TMA store)
  bool att_tc = false;               // tcgen05 self-attention usable (T <= 256)
  bool xa_fused = false;             // cross-attention folded into the cross.wq GEMM epilogue (L <= XA_MAX_TK)
  double flops_per_eval = 0;
  // the whole ODE solve (2*n_steps evaluations for midpoint, ~5.9k launches at 24 layers) as one CUDA graph, captured on the second solve
  // of a plan (the first one runs eagerly and configures every kernel's attributes)
  cudaGraphExec_t solve_graph = nullptr;
  int solve_graph_steps = 0;
  int64_t solve_graph_launches = 0;
  int solves = 0;
  int time_steps_uploaded = 0;
  int solve_method = 0;              // SAB_ODE_* of the cached time tables / captured graph
  int64_t time_cap = 0;              // capacity of time_dev (floats)
  float* rk[4] = {nullptr, nullptr, nullptr, nullptr};   // rk4 stage velocities
  ~DitPlan() { if (solve_graph) cudaGraphExecDestroy(solve_graph); }
};

static void build_time_state(sab_engine* e, DitPlan& p, TimeState& ts, int R) {
  const sab_config& c = e->cfg;
  const int d = c.dim, hid = c.ffn_hidden, NL = c.n_layers;
  DevicePool& w = p.pool;
  ts.R = R;
  ts.time = w.alloc<float>(R);
  ts.t = w.alloc<float>((int64_t)R * d); ts.t0 = w.alloc<float>((int64_t)R * 6 * d);
  ts.mod = w.alloc<float>((int64_t)NL * R * 6 * d); ts.fin = w.alloc<float>((int64_t)R * 2 * d);
  ts.tfreq = w.alloc<bf16>((int64_t)R * 256); ts.t_h = w.alloc<bf16>((int64_t)R * d);
  ts.t_silu = w.alloc<bf16>((int64_t)R * d);
  ts.g_t13 = make_linear("t_embedder.w13", ts.tfreq, R, 256, e->t_w13, 2 * d, 256, EPI_SWIGLU);
  ts.g_t13.P.out_bf16 = ts.t_h; ts.g_t13.P.out_bf16_ld = d;
  ts.g_t2 = make_linear("t_embedder.w2", ts.t_h, R, d, e->t_w2, d, 256, EPI_AFFINE);
  ts.g_t2.P.out_f32 = ts.t; ts.g_t2.P.out_f32_ld = d;
  ts.g_tb = make_linear("t_block", ts.t_silu, R, d, e->tb_w, 6 * d, 256, EPI_AFFINE);
  ts.g_tb.P.bias = e->tb_b; ts.g_tb.P.out_f32 = ts.t0; ts.g_tb.P.out_f32_ld = 6 * d;
  if (p.fused_norm) {
    ts.cs = w.alloc<float>((int64_t)2 * NL * R * d);
    ts.shift = w.alloc<bf16>((int64_t)2 * NL * R * d);
    ts.bias_qkv = w.alloc<float>((int64_t)NL * R * 3 * d);
    ts.bias_w13 = w.alloc<float>((int64_t)NL * R * 2 * hid);
    ts.g_bias_qkv.resize(NL);
    ts.g_bias_w13.resize(NL);
    for (int l = 0; l < NL; ++l) {
      const LayerW& W = e->layers[l];
      // shift @ W^T: [R, d] x [N, d]^T with the SAME packed weights as the main GEMM (so the bias columns line up
      // with its accumulator columns, head permutation and gate/up interleave included).  Weight-streaming (HBM) bound
      // at a handful of rows: 64-column tiles so that every SM pulls a share of the weight
      ts.g_bias_qkv[l] = make_linear("norm.bias.qkv", ts.shift + (int64_t)(2 * l) * R * d, R, d, W.wqkv, 3 * d, 64, EPI_AFFINE);
      ts.g_bias_qkv[l].P.out_f32 = ts.bias_qkv + (int64_t)l * R * 3 * d; ts.g_bias_qkv[l].P.out_f32_ld = 3 * d;
      ts.g_bias_w13[l] = make_linear("norm.bias.w13", ts.shift + (int64_t)(2 * l + 1) * R * d, R, d, W.w13, 2 * hid, 64, EPI_AFFINE);
      ts.g_bias_w13[l].P.out_f32 = ts.bias_w13 + (int64_t)l * R * 2 * hid; ts.g_bias_w13[l].P.out_f32_ld = 2 * hid;
    }
  }
}

// Conditioning (features, text, video, anchors, masks) is per CLIP; the candidates of a clip (model.py:193-203) share it:
// the once-per-call GEMMs and the text path (memory, y_embedder, cross K/V) run on B clips and the per-sequence
// consumers index clip = sequence / cand.  Inside sab_solve every sequence sees the same time, so the text path of
// a clip is the same for all of its candidates; sab_dit_forward is only used with cand == 1.
static void build_dit_plan(sab_engine* e, int B, int cand, int T, int L, int n_ids_cap) {
  const int Bc = B * cand;
  const sab_config& c = e->cfg;
  const int d = c.dim, hid = c.ffn_hidden, NL = c.n_layers;
  auto P = std::make_unique<DitPlan>();
  DitPlan& p = *P;
  p.B = B; p.cand = cand; p.Bc = Bc; p.T = T; p.L = L;
  const int64_t M = (int64_t)Bc * T, MB = (int64_t)B * T, ML = (int64_t)B * L;
  p.M = M; p.MB = MB; p.ML = ML;
  SAB_CHECK(T <= e->rope_len, "T=%d exceeds the RoPE table (%d)", T, e->rope_len);
  DevicePool& w = p.pool;
  p.y = w.alloc<float>(M * 256); p.ymid = w.alloc<float>(M * 256);
  p.cond = w.alloc<float>(M * d); p.x0 = w.alloc<float>(M * d); p.c1 = w.alloc<float>(M * d);
  p.h = w.alloc<float>(M * d); p.vproj = w.alloc<float>(MB * d);
  p.condB = cand == 1 ? p.cond : w.alloc<float>(MB * d);
  p.mem_base = w.alloc<float>(ML * d);
  p.time_dev = w.alloc<float>((int64_t)Bc * 64); p.time_cap = (int64_t)Bc * 64;
  p.y_bf = w.alloc<bf16>(M * 256); p.gn_a = w.alloc<bf16>(M * d); p.hb = w.alloc<bf16>(M * d);
  p.xn = w.alloc<bf16>(M * d); p.qkv = w.alloc<bf16>(M * 3 * d); p.att = w.alloc<bf16>(M * d);
  p.qc = w.alloc<bf16>(M * d); p.kvc = w.alloc<bf16>(ML * 2 * d * NL); p.u = w.alloc<bf16>(M * hid);
  p.mem_in = w.alloc<bf16>(ML * d); p.y_h = w.alloc<bf16>(ML * d); p.ymem = w.alloc<bf16>(ML * d);
  p.feat_bf = w.alloc<bf16>(MB * 256); p.text_bf = w.alloc<bf16>(ML * c.text_dim);
  p.vid_bf = w.alloc<bf16>(MB * c.vision_dim);
  p.gn_partial = w.alloc<double>((int64_t)Bc * GN_CHUNKS * 2);
  p.pad_mask = w.alloc<uint8_t>(MB); p.text_mask = w.alloc<uint8_t>(ML);
  p.fused_norm = !getenv("SAB_NO_FUSED_NORM");
  if (p.fused_norm) {
    p.ssq_n = 2 * ((d + 255) / 256);
    p.nrm_ssq = w.alloc<float>(M * p.ssq_n, true);
  }
  p.n_ids_cap = n_ids_cap;
  p.anchor_ids = w.alloc<long long>((int64_t)B * p.n_ids_cap); p.anchor_align = w.alloc<long long>(MB);

  // ---- once-per-call conditioning GEMMs ----
  p.g_cond = make_linear("cond.proj_feat", p.feat_bf, MB, 256, e->wp_f, d, 256, EPI_AFFINE);
  p.g_cond.P.bias = e->proj_b; p.g_cond.P.out_f32 = p.condB; p.g_cond.P.out_f32_ld = d;
  p.g_mem = make_linear("cond.memory_proj", p.text_bf, ML, c.text_dim, e->wmem, d, 256, EPI_AFFINE);
  p.g_mem.P.bias = e->mem_b; p.g_mem.P.out_f32 = p.mem_base; p.g_mem.P.out_f32_ld = d;
  p.g_vid = make_linear("cond.align_video", p.vid_bf, MB, c.vision_dim, e->wvid, d, 256, EPI_AFFINE);
  p.g_vid.P.bias = e->vid_b; p.g_vid.P.out_f32 = p.vproj; p.g_vid.P.out_f32_ld = d;

  // ---- per-evaluation prologue (the time-only part lives in a TimeState, built on first use) ----
  p.g_y13 = make_linear("y_embedder.w13", p.mem_in, ML, d, e->y_w13, 2 * d, 256, EPI_SWIGLU);
  p.g_y13.P.out_bf16 = p.y_h; p.g_y13.P.out_bf16_ld = d;
  p.g_y2 = make_linear("y_embedder.w2", p.y_h, ML, d, e->y_w2, d, 256, EPI_AFFINE);
  p.g_y2.P.out_bf16 = p.ymem; p.g_y2.P.out_bf16_ld = d;
  p.g_in = make_linear("proj.noisy", p.y_bf, M, 256, e->wp_y, d, 256, EPI_AFFINE);
  p.g_in.P.res = p.cond; p.g_in.P.res_ld = d; p.g_in.P.out_f32 = p.x0; p.g_in.P.out_f32_ld = d;
  for (int b = 0; b < 2; ++b) {
    RunList rl;   // Conv1d k=3, zero padding (1,1): taps t-1, t, t+1  (patcher.py:52-67)
    for (int k = 0; k < 3; ++k) rl.add(0, k - 1, 0, d / 64);
    p.g_xe[b] = make_gemm(b == 0 ? "x_embedder.conv1" : "x_embedder.conv2", seq_view(p.gn_a, Bc, T, d), e->xe_w[b], d,
                          256, 64, EPI_AFFINE, rl);
    p.g_xe[b].P.bias = e->xe_b[b];
    p.g_xe[b].P.out_f32 = b == 0 ? p.c1 : p.h; p.g_xe[b].P.out_f32_ld = d;
    if (b == 1) { p.g_xe[b].P.res = p.x0; p.g_xe[b].P.res_ld = d; }
  }
  // fused RMSNorm: the kernel that last writes h before a norm also emits that norm's GEMM operand h*w*(1+scale) (into
  // xn) and the partial sums of squares; `which` = 2l (attention_norm of layer l) or 2l+1 (ffn_norm)
  auto norm_producer = [&](GemmOp& op, int which) {
    if (!p.fused_norm) return;
    (void)which;                  // colscale (TimeState row) is bound per evaluation in dit_eval
    op.mode = EPI_AFFINE_NORM;
    op.P.gate_div = T;
    op.P.out_scaled = p.xn; op.P.out_scaled_ld = d;
    op.P.ssq_out = p.nrm_ssq; op.P.ssq_ld = p.ssq_n;
  };
  auto norm_consumer = [&](GemmOp& op) {   // ibias (TimeState row) is bound per evaluation in dit_eval
    if (!p.fused_norm) return;
    op.P.ssq_in = p.nrm_ssq; op.P.ssq_n = p.ssq_n; op.P.ssq_inv_dim = 1.0f / (float)d;
    op.P.ibias_div = T;
  };
  norm_producer(p.g_xe[1], 0);
  // ---- layers ----
  p.lay.resize(NL);
  p.xa_fused = (L <= XA_MAX_TK) && !getenv("SAB_NO_FUSED_XATTN");
  for (int l = 0; l < NL; ++l) {
    const LayerW& W = e->layers[l];
    LayerOps& o = p.lay[l];
    const long long kv_ld = 2LL * d * NL;            // layer l's K|V live at columns [l*2d, (l+1)*2d) of p.kvc
    bf16* kvc_l = p.kvc + (long long)l * 2 * d;
    o.qkv = make_linear("attention.qkv", p.xn, M, d, W.wqkv, 3 * d, 256, EPI_QKV);
    o.qkv.P.out_bf16 = p.qkv; o.qkv.P.out_bf16_ld = 3 * d;
    o.qkv.P.qnorm_w = W.qn_scaled; o.qkv.P.knorm_w = W.kn; o.qkv.P.n_q_end = d; o.qkv.P.n_k_end = 2 * d;
    o.qkv.P.rope = e->rope; o.qkv.P.rope_T = T; o.qkv.P.use_rope = 1; o.qkv.P.eps = c.norm_eps;
    norm_consumer(o.qkv);
    o.wo = make_linear("attention.wo", p.att, M, d, W.wo, d, 256, EPI_AFFINE);
    o.wo.P.gate_div = T;     // gate rows (adaLN gate_msa / gate_mlp of the TimeState) are bound per evaluation
    o.wo.P.res = p.h; o.wo.P.res_ld = d; o.wo.P.out_f32 = p.h; o.wo.P.out_f32_ld = d;
    o.wo.P.out_bf16 = p.hb; o.wo.P.out_bf16_ld = d;
    o.q_c = make_linear("cross.wq", p.hb, M, d, W.wq_c, d, 256, EPI_QKV);
    o.q_c.P.out_bf16 = p.qc; o.q_c.P.out_bf16_ld = d;
    o.q_c.P.qnorm_w = W.qn_c; o.q_c.P.knorm_w = W.kn_c; o.q_c.P.n_q_end = d; o.q_c.P.n_k_end = d;
    o.q_c.P.rope_T = T; o.q_c.P.use_rope = 0; o.q_c.P.eps = c.norm_eps;
    if (p.xa_fused) {   // epilogue attends to the layer's text K/V and writes the attention output directly
      SAB_CHECK(o.q_c.BN == 256, "fused cross-attention needs BN=256 tiles");
      o.q_c.tag = "cross.wq+attn";
      o.q_c.P.out_bf16 = p.att; o.q_c.P.out_bf16_ld = d;
      o.q_c.P.xa_kv = kvc_l; o.q_c.P.xa_kv_ld = kv_ld; o.q_c.P.xa_v_col0 = d; o.q_c.P.xa_Tk = L;
      o.q_c.P.xa_T = T * cand;   // query rows per K/V item: the cand sequences of a clip attend to the clip's text
      o.q_c.P.xa_mask = p.text_mask;
      o.q_c.P.xa_scale_log2 = (1.0f / sqrtf(128.f)) * 1.4426950408889634f;
      o.q_c.flops += 4.0 * Bc * c.n_heads * (double)T * L * 128;
    }

    o.wo_c = make_linear("cross.wo", p.att, M, d, W.wo_c, d, 256, EPI_AFFINE);
    o.wo_c.P.res = p.h; o.wo_c.P.res_ld = d; o.wo_c.P.out_f32 = p.h; o.wo_c.P.out_f32_ld = d;
    norm_producer(o.wo_c, 2 * l + 1);
    o.w13 = make_linear("ffn.w13", p.xn, M, d, W.w13, 2 * hid, 256, EPI_SWIGLU);
    o.w13.P.out_bf16 = p.u; o.w13.P.out_bf16_ld = hid; o.w13.P.eps = c.norm_eps;
    norm_consumer(o.w13);
    o.w2 = make_linear("ffn.w2", p.u, M, hid, W.w2, d, 256, EPI_AFFINE);
    o.w2.P.gate_div = T;
    o.w2.P.res = p.h; o.w2.P.res_ld = d; o.w2.P.out_f32 = p.h; o.w2.P.out_f32_ld = d;
    if (l + 1 < NL) norm_producer(o.w2, 2 * (l + 1));      // the last layer feeds the final norm (standalone kernel)
  }
  p.g_out = make_linear("output", p.xn, M, d, e->w_out, c.out_channels, 256, EPI_AFFINE);
  // every layer's text K|V in one GEMM per evaluation: [Bc*L, d] x [NL*2d, d]^T, k-norm per layer
  p.g_kvc_all = make_linear("cross.wkv(all layers)", p.ymem, ML, d, e->wkv_c_all, 2 * d * NL, 256, EPI_QKV);
  p.g_kvc_all.P.out_bf16 = p.kvc; p.g_kvc_all.P.out_bf16_ld = 2LL * d * NL;
  p.g_kvc_all.P.qnorm_w = e->kn_c_all; p.g_kvc_all.P.knorm_w = e->kn_c_all;
  p.g_kvc_all.P.n_q_end = 0; p.g_kvc_all.P.n_k_end = d; p.g_kvc_all.P.qkv_period = 2 * d; p.g_kvc_all.P.norm_w_stride = 128;
  p.g_kvc_all.P.rope_T = L; p.g_kvc_all.P.use_rope = 0; p.g_kvc_all.P.eps = c.norm_eps;
  p.att_tc = (T <= 256) && !getenv("SAB_NO_TC_ATTENTION");
  if (p.att_tc) {
    p.tm_att_q = make_tmap_3d(p.qkv, 3LL * d, T, Bc, 3LL * d, (int64_t)T * 3 * d, 64, 128);
    p.tm_att_kv = make_tmap_3d(p.qkv, 3LL * d, T, Bc, 3LL * d, (int64_t)T * 3 * d, 64, 256);
    p.tm_att_o = make_tmap_3d(p.att, d, T, Bc, d, (int64_t)T * d, 64, 128);
  }

  // algorithmic FLOPs of one evaluation (GEMMs + attention), for roofline reporting
  double f = p.g_kvc_all.flops + p.g_y13.flops + p.g_y2.flops + p.g_in.flops +
             p.g_xe[0].flops + p.g_xe[1].flops + p.g_out.flops;
  for (auto& o : p.lay)
    f += o.qkv.flops + o.wo.flops + o.q_c.flops + o.wo_c.flops + o.w13.flops + o.w2.flops +
         4.0 * Bc * (double)T * T * d + 4.0 * Bc * (double)T * L * d;
  p.flops_per_eval = f;
  e->dit = std::move(P);
}

// =====================================================================================================
// DiT evaluation
// =====================================================================================================
template <int V>
static void launch_rmsnorm(const float* x, const float* w, const float* shift, const float* scale, long long mod_ld,
                           int rows_per_item, bf16* out, int M, float eps, int reverse, cudaStream_t st) {
  rmsnorm_mod_kernel<V><<<(M + 7) / 8, 256, 0, st>>>(x, w, shift, scale, mod_ld, rows_per_item, out, M, eps, reverse);
}
static void rmsnorm_mod(sab_engine* e, const float* x, const float* w, const float* shift, const float* scale,
                        long long mod_ld, int rows_per_item, bf16* out, int M, cudaStream_t st, int reverse = 0) {
  const int d = e->cfg.dim;
  const float eps = e->cfg.norm_eps;
  mark(e, st, "rmsnorm_mod", 0, (double)M * d * 6.0);
  switch (d / 128) {
#define SAB_RN(v) case v: launch_rmsnorm<v>(x, w, shift, scale, mod_ld, rows_per_item, out, M, eps, reverse, st); break;
    SAB_RN(2) SAB_RN(4) SAB_RN(8) SAB_RN(12) SAB_RN(16) SAB_RN(20) SAB_RN(22) SAB_RN(24) SAB_RN(32)
#undef SAB_RN
    default: throw Error(fmt("rmsnorm: unsupported dim %d (add an instantiation)", d));
  }
  SAB_CUDA(cudaGetLastError());
}

static void attention(sab_engine* e, const AttnParams& ap, int items, int heads, cudaStream_t st) {
  ensure_dynamic_smem(reinterpret_cast<const void*>(attention_kernel), ATT_SMEM);
  const double fl = 4.0 * items * heads * (double)ap.Tq * ap.Tk * 128;
  if (ap.Tk <= XATT_MAX_TK && ap.k != ap.q) {   // a handful of text tokens: HBM-bound warp-per-row kernel
    mark(e, st, "sdpa.cross", fl, (double)items * ap.Tq * heads * 128 * 4.0);
    launch_xattn_small(ap, items, heads, st);
    SAB_CUDA(cudaGetLastError());
    return;
  }
  dim3 grid((ap.Tq + ATT_BQ - 1) / ATT_BQ, heads, items);
  mark(e, st, ap.k == ap.q ? "sdpa.self" : "sdpa.cross", fl, 0);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, st>>>(ap);
  SAB_CUDA(cudaGetLastError());
}

static void launch_attention_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& ap,
                                int items, cudaStream_t st) {
  ensure_dynamic_smem(reinterpret_cast<const void*>(attention_tc_kernel), ATC_SMEM);
  AttnTcParams p2 = ap;
  p2.items = items;
  const int n_work = ap.heads * items;
  attention_tc_kernel<<<n_work < sm_count() ? n_work : sm_count(), ATC_THREADS, ATC_SMEM, st>>>(tq, tk, tv, p2);
  SAB_CUDA(cudaGetLastError());
}

// algorithmic HBM bytes of one GEMM launch: operands once + every epilogue stream once (what a perfect kernel moves)
static double gemm_bytes(const GemmOp& op) {
  const GemmParams& P = op.P;
  double per_elem = 0;
  if (P.res) per_elem += 4;
  if (P.out_f32) per_elem += 4;
  if (P.out_bf16) per_elem += 2;
  if (P.out_act) per_elem += 2;
  const double n_out = (op.mode == EPI_SWIGLU) ? P.N / 2.0 : (double)P.N;
  return op.in_bytes + op.rows * n_out * per_elem;
}
// second-generation tcgen05 self-attention (attention_tc2.cuh).  shift_log2 < 0 selects the exact two-pass softmax.
static int g_attn_poly = 2;      // SAB_ATTN_POLY=0..4: pairs of every 8 whose exponential runs on the FMA pipe
template <bool kExact, int kPoly, bool kFolded>
static void launch_attention_tc2_inst(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                      const CUtensorMap& to, const AttnTc2Params& ap, cudaStream_t st) {
  auto kern = attention_tc2_kernel<kExact, kPoly, kFolded>;
  ensure_dynamic_smem(reinterpret_cast<const void*>(kern), AT2_SMEM);
  const int n_work = ap.heads * ap.items;
  kern<<<n_work < sm_count() ? n_work : sm_count(), AT2_THREADS, AT2_SMEM, st>>>(tq, tk, tv, to, ap);
  SAB_CUDA(cudaGetLastError());
}
// variants: exact (two-pass row maximum; shift_log2 < 0), single pass with scale and shift, and `folded`
// (scale already in q, |logit| <= 50: p = 2^s) — the engine's production path
static void launch_attention_tc2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                                 const AttnTc2Params& ap, bool folded, cudaStream_t st) {
  const bool exact = ap.shift_log2 < 0.f;
  SAB_CHECK(!(exact && folded), "folded softmax needs a logit bound");
#define SAB_AT2(p_) case p_: return exact ? launch_attention_tc2_inst<true, p_, false>(tq, tk, tv, to, ap, st)   \
                                  : folded ? launch_attention_tc2_inst<false, p_, true>(tq, tk, tv, to, ap, st)  \
                                           : launch_attention_tc2_inst<false, p_, false>(tq, tk, tv, to, ap, st);
  switch (g_attn_poly) {
    SAB_AT2(0) SAB_AT2(2) SAB_AT2(3) SAB_AT2(4)
    default: throw Error(fmt("SAB_ATTN_POLY=%d not built (0, 2, 3, 4)", g_attn_poly));
  }
#undef SAB_AT2
}

static void gemm(sab_engine* e, const GemmOp& op, cudaStream_t st) {
  mark(e, st, op.tag, op.flops, gemm_bytes(op));
  launch_gemm(op, st);
}
// Serpentine row order.  Every kernel of a DiT layer is row-parallel and streams activations that exceed the 126 MB
// L2 (fp32 residual 180 MB, fused QKV 270 MB, FFN hidden 241 MB at B = 64): a consumer that walks the rows in the
// same direction as its producer finds the first rows evicted.  Alternating the direction from one launch to the
// next makes each kernel start on the rows its producer wrote last, so roughly an L2's worth of its input (and of the
// in-place residual it updates) never travels to HBM.
static bool g_serpentine = true;   // SAB_NO_SERPENTINE=1 disables (A/B)
static void gemm_dir(sab_engine* e, const GemmOp& op, int& dir, cudaStream_t st) {
  GemmOp o2 = op;
  o2.P.reverse_m = g_serpentine ? dir : 0;
  dir ^= 1;
  gemm(e, o2, st);
}

// final-layer variants: out = base + coef * velocity (ODE axpy fused in the output GEMM's epilogue)
struct FinalSpec {
  const float* base;   // nullptr: plain velocity
  float coef;
  float* out_f32;
  bf16* out_bf16;      // next evaluation's GEMM operand (or nullptr)
};

// the time-only part of an evaluation for the R rows of a TimeState (ts.time holds the R times)
static void run_time_state(sab_engine* e, DitPlan& p, TimeState& ts, cudaStream_t st) {
  const sab_config& c = e->cfg;
  const int d = c.dim, NL = c.n_layers, R = ts.R;
  mark(e, st, "time_features_kernel");
  tfreq_kernel<<<(R * 256 + 255) / 256, 256, 0, st>>>(ts.time, R, ts.tfreq);
  gemm(e, ts.g_t13, st);
  gemm(e, ts.g_t2, st);
  mark(e, st, "silu_cast_kernel");
  silu_cast_kernel<<<64, 256, 0, st>>>(ts.t, ts.t_silu, (long long)R * d);
  gemm(e, ts.g_tb, st);
  mark(e, st, "build_mod_kernel");
  build_mod_kernel<<<512, 256, 0, st>>>(e->tables, ts.t0, ts.mod, NL, R, d, e->final_table, ts.t, ts.fin);
  if (p.fused_norm) {
    mark(e, st, "norm_tables_kernel");
    norm_tables_kernel<<<512, 256, 0, st>>>(ts.mod, e->norm_w_all, NL, R, d, ts.cs, ts.shift);
    for (int l = 0; l < NL; ++l) {
      gemm(e, ts.g_bias_qkv[l], st);
      gemm(e, ts.g_bias_w13[l], st);
    }
  }
  SAB_CUDA(cudaGetLastError());
}

// One evaluation.  `row` < 0: ts has one row per sequence (per-item adaLN state); `row` >= 0: ts row `row` holds the
// state of this evaluation's time, shared by every sequence (item stride 0).  time_dev: the Bc times (text memory).
static void dit_eval(sab_engine* e, const float* time_dev, const FinalSpec& fs, const TimeState& ts, int row,
                     cudaStream_t st) {
  DitPlan& p = *e->dit;
  const sab_config& c = e->cfg;
  const int d = c.dim, Bc = p.Bc, T = p.T, L = p.L, NL = c.n_layers, H = c.n_heads, cand = p.cand, hid = c.ffn_hidden;
  const int M = (int)p.M;
  const float sl2 = (1.0f / sqrtf(128.f)) * 1.4426950408889634f;
  const int R = ts.R, r0 = row < 0 ? 0 : row;
  const int istride = row < 0 ? 1 : 0;        // per-item tables advance by one row per sequence, shared ones do not
  auto mod_row = [&](int l, int r) { return ts.mod + (((int64_t)l * R + r0) * 6 + r) * d; };
  auto bind_gate = [&](GemmOp& op, int l, int r) { op.P.gate = mod_row(l, r); op.P.gate_ld = istride * 6 * d; };
  auto bind_cs = [&](GemmOp& op, int which) {
    if (p.fused_norm) { op.P.colscale = ts.cs + ((int64_t)which * R + r0) * d; op.P.colscale_ld = (long long)istride * d; }
  };

  mark(e, st, "mem_time_kernel");
  mem_time_kernel<<<256, 256, 0, st>>>(time_dev, Bc, cand, d, L, p.mem_base, p.mem_in);
  gemm(e, p.g_y13, st);
  gemm(e, p.g_y2, st);
  gemm(e, p.g_in, st);
  // x_embedder (patcher.py:138-164): GN(1) -> SiLU -> conv3, twice, + x
  for (int b = 0; b < 2; ++b) {
    const float* src = b == 0 ? p.x0 : p.c1;
    mark(e, st, "gn_stats_kernel");
    gn_stats_kernel<<<dim3(GN_CHUNKS, Bc), 256, 0, st>>>(src, (long long)T * d, p.gn_partial);
    mark(e, st, "gn_silu_kernel");
    gn_silu_kernel<<<dim3(64, Bc), 256, 0, st>>>(src, p.gn_partial, e->xe_gn_w[b], e->xe_gn_b[b], d, (long long)T * d,
                                               1e-5f, p.gn_a);
    GemmOp xe = p.g_xe[b];
    if (b == 1) bind_cs(xe, 0);
    gemm(e, xe, st);
  }
  SAB_CUDA(cudaGetLastError());
  gemm(e, p.g_kvc_all, st);
  int dir = 1;   // x_embedder.conv2 (the producer of h) walked the rows forwards
  for (int l = 0; l < NL; ++l) {
    const LayerW& W = e->layers[l];
    LayerOps o = p.lay[l];      // by value: the TimeState rows of this evaluation are bound below
    const long long mod_ld = (long long)istride * 6 * d;
    if (!p.fused_norm) {
      rmsnorm_mod(e, p.h, W.attn_norm, mod_row(l, 0), mod_row(l, 1), mod_ld, T, p.xn, M, st, g_serpentine ? dir : 0);
      dir ^= 1;
    } else {
      o.qkv.P.ibias = ts.bias_qkv + ((int64_t)l * R + r0) * 3 * d; o.qkv.P.ibias_ld = (long long)istride * 3 * d;
      o.w13.P.ibias = ts.bias_w13 + ((int64_t)l * R + r0) * 2 * hid; o.w13.P.ibias_ld = (long long)istride * 2 * hid;
      bind_cs(o.wo_c, 2 * l + 1);
      if (l + 1 < NL) bind_cs(o.w2, 2 * (l + 1));
    }
    bind_gate(o.wo, l, 2);
    bind_gate(o.w2, l, 5);
    gemm_dir(e, o.qkv, dir, st);
    AttnParams a{};
    a.q = p.qkv; a.q_ld = 3 * d; a.q_col0 = 0;
    a.k = p.qkv; a.k_ld = 3 * d; a.k_col0 = d;
    a.v = p.qkv; a.v_ld = 3 * d; a.v_col0 = 2 * d;
    a.o = p.att; a.o_ld = d; a.key_mask = p.pad_mask; a.Tq = T; a.Tk = T; a.mask_div = cand;
    a.scale_log2 = 1.f;          // self-attention: log2(e)/sqrt(hd) is folded into q by the QKV epilogue (W.qn_scaled)
    static const bool attn_v1 = getenv("SAB_ATTN_V1") != nullptr;   // A/B against the first-generation kernel
    if (p.att_tc && !attn_v1) {
      AttnTc2Params tp{};
      tp.key_mask = p.pad_mask; tp.T = T; tp.heads = H; tp.items = Bc; tp.mask_div = cand;
      // the softmax scale is already in q (QKV epilogue, W.qn_scaled): scale 1 for the exact fallback, and with a
      // logit bound the folded single-pass kernel needs neither scale nor shift
      tp.q_col0 = 0; tp.k_col0 = d; tp.v_col0 = 2 * d; tp.scale_log2 = 1.f; tp.shift_log2 = W.att_shift_log2;
      tp.reverse = g_serpentine ? dir : 0;
      dir ^= 1;
      mark(e, st, "sdpa.self", 4.0 * Bc * H * (double)T * T * 128, (double)M * 4.0 * d * 2.0);   // reads q|k|v, writes o (bf16)
      launch_attention_tc2(p.tm_att_q, p.tm_att_kv, p.tm_att_kv, p.tm_att_o, tp, W.att_shift_log2 >= 0.f, st);
    } else if (p.att_tc) {
      AttnTcParams tp{};
      tp.o = p.att; tp.o_ld = d; tp.key_mask = p.pad_mask; tp.Tq = T; tp.Tk = T; tp.heads = H; tp.mask_div = cand;
      tp.q_col0 = 0; tp.k_col0 = d; tp.v_col0 = 2 * d; tp.scale_log2 = 1.f; tp.v_lbo = ATC_KV_BYTES / 2; tp.v_sbo = 1024;
      mark(e, st, "sdpa.self", 4.0 * Bc * H * (double)T * T * 128, 0);
      launch_attention_tc(p.tm_att_q, p.tm_att_kv, p.tm_att_kv, tp, Bc, st);
    } else {
      attention(e, a, Bc, H, st);
    }
    gemm_dir(e, o.wo, dir, st);
    gemm_dir(e, o.q_c, dir, st);
    if (!p.xa_fused) {
      AttnParams x{};
      x.q = p.qc; x.q_ld = d; x.q_col0 = 0;
      x.k = p.kvc; x.k_ld = 2LL * d * NL; x.k_col0 = l * 2 * d;
      x.v = p.kvc; x.v_ld = 2LL * d * NL; x.v_col0 = l * 2 * d + d;
      x.o = p.att; x.o_ld = d; x.key_mask = p.text_mask; x.Tq = T; x.Tk = L; x.scale_log2 = sl2;
      x.kv_div = cand; x.mask_div = cand;
      attention(e, x, Bc, H, st);
    }
    gemm_dir(e, o.wo_c, dir, st);
    if (!p.fused_norm) {
      rmsnorm_mod(e, p.h, W.ffn_norm, mod_row(l, 3), mod_row(l, 4), mod_ld, T, p.xn, M, st, g_serpentine ? dir : 0);
      dir ^= 1;
    }
    gemm_dir(e, o.w13, dir, st);
    gemm_dir(e, o.w2, dir, st);
  }
  rmsnorm_mod(e, p.h, e->final_norm, ts.fin + (int64_t)r0 * 2 * d, ts.fin + (int64_t)r0 * 2 * d + d, (long long)istride * 2 * d, T,
              p.xn, M, st);
  GemmOp out = p.g_out;
  out.P.res = fs.base; out.P.res_ld = 256; out.P.alpha = fs.coef;
  out.P.out_f32 = fs.out_f32; out.P.out_f32_ld = 256;
  out.P.out_bf16 = fs.out_bf16; out.P.out_bf16_ld = 256;
  gemm(e, out, st);
}

// =====================================================================================================
// codec plans
// =====================================================================================================
struct CodecStep {
  enum Kind { GEMM, ENC0, DEC_LAST, LATENT_SPLIT } kind = GEMM;
  GemmOp op;
};
struct CodecPlan {
  std::vector<std::unique_ptr<std::string>> tags;   // owned tag strings (profiling labels)
  const char* tag(const std::string& t) { tags.push_back(std::make_unique<std::string>(t)); return tags.back()->c_str(); }
  int items = 0;       // waveforms per chunk
  long long S = 0;     // samples per waveform
  DevicePool pool;
  std::vector<CodecStep> steps;
  // end-point buffers
  float* x_first = nullptr; bf16* a_first = nullptr;     // encoder stage 0
  bf16* z_in = nullptr;                                  // decoder input [items, T, cz] bf16
  bf16* a_last = nullptr;                                // decoder last-stage activated input
  GemmOp* enc_out = nullptr;                             // in_proj (output pointer patched per chunk)
  double flops = 0;
  int64_t last_use = 0;
};
// A few resident plans per direction (ragged traffic alternates between clip lengths / tail chunks): the least
// recently used one goes when a new shape arrives and kMaxCodecPlans are resident.
constexpr size_t kMaxCodecPlans = 3;
static void evict_codec_plans(sab_engine* e, std::map<std::pair<int, long long>, std::unique_ptr<CodecPlan>>& plans) {
  if (plans.size() < kMaxCodecPlans) return;
  SAB_CUDA(cudaDeviceSynchronize());          // the victim's workspace may still be in use by enqueued work
  auto victim = plans.begin();
  for (auto it = plans.begin(); it != plans.end(); ++it)
    if (it->second->last_use < victim->second->last_use) victim = it;
  plans.erase(victim);
}

static int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N % 96 == 0) return 96;
  if (N % 64 == 0) return 64;
  throw Error(fmt("codec: no GEMM tile for N=%d", N));
}
static int pick_bk(int cin) {
  if (cin % 64 == 0) return 64;
  if (cin % 32 == 0) return 32;
  throw Error(fmt("codec: channel count %d not a multiple of 32", cin));
}

struct StageBufs { float* x; bf16* a; bf16* mid; };

// Conv1d(c, c, k=7, dilation) 'same' + Snake  ->  Conv1d(c, c, 1) + residual  (+ Snake for the consumer)
// C <= 128: ONE launch — the 1x1 conv runs back to back on the tensor core from the activated tile in shared
// memory (the intermediate never touches HBM); its output goes to the other operand buffer (sb.a / sb.mid swap)
// because neighbouring tiles still read this unit's input through their dilated taps.
static void plan_resunit(CodecPlan& cp, const ResUnitW& R, int items, long long Tn, int C, StageBufs& sb,
                         const float* next_alpha, bool keep_x) {
  const int bk = pick_bk(C), bn = pick_bn(C);
  static const bool no_b2b = getenv("SAB_NO_B2B") != nullptr;
  if (C <= 128 && bn == C && !no_b2b) {
    RunList rl;
    for (int k = 0; k < 7; ++k) rl.add(0, (k - 3) * R.c7.dil, 0, C / bk);
    CodecStep s;
    s.op = make_gemm(cp.tag(fmt("codec.res.b2b.c%d", C)), seq_view(sb.a, items, Tn, C), R.c7.w, C, bn, bk, EPI_AFFINE, rl, 1);
    s.op.b2b = true;
    s.op.tmW = make_tmap_2d(R.c1.w, C, C, C, bk, bn);
    s.op.P.b2b_bias = R.c7.bias; s.op.P.b2b_alpha = R.a1;
    s.op.P.bias = R.c1.bias;
    s.op.P.res = sb.x; s.op.P.res_ld = C;
    if (keep_x) { s.op.P.out_f32 = sb.x; s.op.P.out_f32_ld = C; }
    s.op.P.out_act = sb.mid; s.op.P.out_act_ld = C; s.op.P.snake_alpha = next_alpha;
    s.op.flops += 2.0 * items * (double)Tn * C * C;
    cp.flops += s.op.flops;
    cp.steps.push_back(s);
    std::swap(sb.a, sb.mid);
    return;
  }
  {
    RunList rl;
    for (int k = 0; k < 7; ++k) rl.add(0, (k - 3) * R.c7.dil, 0, C / bk);
    CodecStep s;
    s.op = make_gemm(cp.tag(fmt("codec.res.conv7.c%d", C)), seq_view(sb.a, items, Tn, C), R.c7.w, C, bn, bk, EPI_AFFINE, rl);
    s.op.P.bias = R.c7.bias;
    s.op.P.out_act = sb.mid; s.op.P.out_act_ld = C; s.op.P.snake_alpha = R.a1;
    cp.flops += s.op.flops;
    cp.steps.push_back(s);
  }
  {
    RunList rl;
    rl.add(0, 0, 0, C / bk);
    CodecStep s;
    s.op = make_gemm(cp.tag(fmt("codec.res.conv1.c%d", C)), seq_view(sb.mid, items, Tn, C), R.c1.w, C, bn, bk, EPI_AFFINE, rl);
    s.op.P.bias = R.c1.bias;
    s.op.P.res = sb.x; s.op.P.res_ld = C;
    if (keep_x) { s.op.P.out_f32 = sb.x; s.op.P.out_f32_ld = C; }
    s.op.P.out_act = sb.a; s.op.P.out_act_ld = C; s.op.P.snake_alpha = next_alpha;
    cp.flops += s.op.flops;
    cp.steps.push_back(s);
  }
}

static CodecPlan* get_enc_plan(sab_engine* e, int items, long long S) {
  auto key = std::make_pair(items, S);
  auto it = e->enc_plans.find(key);
  if (it != e->enc_plans.end()) { it->second->last_use = ++e->plan_clock; return it->second.get(); }
  evict_codec_plans(e, e->enc_plans);
  const sab_config& c = e->cfg;
  auto CP = std::make_unique<CodecPlan>();
  CodecPlan& cp = *CP;
  cp.items = items; cp.S = S;
  int C = c.codec_encoder_dim;
  long long Tn = S;
  StageBufs sb{cp.pool.alloc<float>(items * Tn * C), cp.pool.alloc<bf16>(items * Tn * C), cp.pool.alloc<bf16>(items * Tn * C)};
  cp.x_first = sb.x; cp.a_first = sb.a;
  { CodecStep s; s.kind = CodecStep::ENC0; cp.steps.push_back(s); }
  for (int i = 0; i < c.codec_n_rates; ++i) {
    const EncBlockW& B = e->enc_blocks[i];
    for (int j = 0; j < 3; ++j)
      plan_resunit(cp, B.ru[j], items, Tn, C, sb, j < 2 ? B.ru[j + 1].a0 : B.a_down, j < 2);
    // strided Conv1d(C, 2C, k=2s, stride s, pad s/2) over the [Tn/s, s*C] view of the activated stream
    const int s = B.stride, pd = s / 2;
    SAB_CHECK(Tn % s == 0, "encoder length %lld not divisible by rate %d", Tn, s);
    const long long To = Tn / s;
    const int C2 = 2 * C;
    StageBufs nb{cp.pool.alloc<float>(items * To * C2), cp.pool.alloc<bf16>(items * To * C2),
                 cp.pool.alloc<bf16>(items * To * C2)};
    const int bk = 64;
    SAB_CHECK((pd * C) % bk == 0, "encoder stride tap split not BK aligned");
    RunList rl;
    rl.add(0, -1, (s - pd) * C, pd * C / bk);
    rl.add(0, 0, 0, s * C / bk);
    rl.add(0, +1, 0, (s - pd) * C / bk);
    CodecStep st;
    AView av{sb.a, (int64_t)s * C, To, items, (int64_t)s * C, To * s * C};
    st.op = make_gemm(cp.tag(fmt("codec.enc.down.c%d", C)), av, B.down.w, C2, pick_bn(C2), bk, EPI_AFFINE, rl);
    st.op.P.bias = B.down.bias;
    const bool last = (i == c.codec_n_rates - 1);
    st.op.P.out_f32 = nb.x; st.op.P.out_f32_ld = C2;
    st.op.P.out_act = nb.a; st.op.P.out_act_ld = C2;
    st.op.P.snake_alpha = last ? e->enc_final_alpha : e->enc_blocks[i + 1].ru[0].a0;
    cp.flops += st.op.flops;
    cp.steps.push_back(st);
    sb = nb; C = C2; Tn = To;
  }
  {  // Snake (already applied) -> Conv1d(C, latent, k=3, pad 1) -> bf16 for in_proj
    RunList rl;
    for (int k = 0; k < 3; ++k) rl.add(0, k - 1, 0, C / 64);
    bf16* zl = cp.pool.alloc<bf16>(items * Tn * c.codec_latent_dim);
    CodecStep st;
    st.op = make_gemm("codec.enc.final", seq_view(sb.a, items, Tn, C), e->enc_final.w, c.codec_latent_dim,
                      pick_bn(c.codec_latent_dim), 64, EPI_AFFINE, rl);
    st.op.P.bias = e->enc_final.bias;
    st.op.P.out_bf16 = zl; st.op.P.out_bf16_ld = c.codec_latent_dim;
    cp.flops += st.op.flops;
    cp.steps.push_back(st);
    RunList r1;
    r1.add(0, 0, 0, c.codec_latent_dim / 64);
    CodecStep s2;
    s2.op = make_gemm("codec.in_proj", seq_view(zl, items, Tn, c.codec_latent_dim), e->in_proj.w,
                      2 * c.codec_codebook_dim, pick_bn(2 * c.codec_codebook_dim), 64, EPI_AFFINE, r1);
    s2.op.P.bias = e->in_proj.bias;
    s2.op.P.out_f32_ld = 2 * c.codec_codebook_dim;  // out_f32 patched per call
    cp.flops += s2.op.flops;
    cp.steps.push_back(s2);
  }
  cp.enc_out = &cp.steps.back().op;
  cp.last_use = ++e->plan_clock;
  CodecPlan* r = CP.get();
  e->enc_plans[key] = std::move(CP);
  return r;
}

static CodecPlan* get_dec_plan(sab_engine* e, int items, long long T) {
  auto key = std::make_pair(items, T);
  auto it = e->dec_plans.find(key);
  if (it != e->dec_plans.end()) { it->second->last_use = ++e->plan_clock; return it->second.get(); }
  evict_codec_plans(e, e->dec_plans);
  const sab_config& c = e->cfg;
  auto CP = std::make_unique<CodecPlan>();
  CodecPlan& cp = *CP;
  cp.items = items;
  const int cz = c.codec_codebook_dim, ld = c.codec_latent_dim;
  long long Tn = T;
  cp.z_in = cp.pool.alloc<bf16>(items * Tn * cz);
  { CodecStep s; s.kind = CodecStep::LATENT_SPLIT; cp.steps.push_back(s); }
  bf16* e0 = cp.pool.alloc<bf16>(items * Tn * ld);
  {
    RunList rl;
    rl.add(0, 0, 0, cz / 64);
    CodecStep s;
    s.op = make_gemm("codec.out_proj", seq_view(cp.z_in, items, Tn, cz), e->out_proj.w, ld, pick_bn(ld), 64, EPI_AFFINE, rl);
    s.op.P.bias = e->out_proj.bias;
    s.op.P.out_bf16 = e0; s.op.P.out_bf16_ld = ld;
    cp.flops += s.op.flops;
    cp.steps.push_back(s);
  }
  int C = c.codec_decoder_dim;
  bf16* a_cur = cp.pool.alloc<bf16>(items * Tn * C);
  {
    RunList rl;
    for (int k = 0; k < 7; ++k) rl.add(0, k - 3, 0, ld / 64);
    CodecStep s;
    s.op = make_gemm("codec.dec.conv_in", seq_view(e0, items, Tn, ld), e->dec0.w, C, pick_bn(C), 64, EPI_AFFINE, rl);
    s.op.P.bias = e->dec0.bias;
    s.op.P.out_act = a_cur; s.op.P.out_act_ld = C; s.op.P.snake_alpha = e->dec_blocks[0].a_up;
    cp.flops += s.op.flops;
    cp.steps.push_back(s);
  }
  for (int i = 0; i < c.codec_n_rates; ++i) {
    const DecBlockW& B = e->dec_blocks[i];
    const int s = B.stride, Co = B.cout;
    const long long To = Tn * s;
    StageBufs sb{cp.pool.alloc<float>(items * To * Co), cp.pool.alloc<bf16>(items * To * Co),
                 cp.pool.alloc<bf16>(items * To * Co)};
    {
      // ConvTranspose1d(C, Co, k=2s, stride s, pad s/2): output rows viewed as [Tn, s*Co]
      const int bk = pick_bk(C);
      int bn = pick_bn(Co);                       // must divide the tap switch (s/2)*Co and the period s*Co
      while (bn > 64 && (((s / 2) * Co) % bn != 0)) bn = (bn == 256) ? 192 : (bn == 192) ? 128 : (bn == 128) ? 96 : 64;
      RunList rl;
      rl.period = s * Co; rl.sw = (s / 2) * Co;
      rl.add(0, 0, 0, C / bk); rl.add(0, -1, 0, C / bk);
      rl.add(1, 0, 0, C / bk); rl.add(1, +1, 0, C / bk);
      CodecStep st;
      st.op = make_gemm(cp.tag(fmt("codec.dec.up.c%d", C)), seq_view(a_cur, items, Tn, C), B.up.w, s * Co, bn, bk, EPI_AFFINE, rl);
      st.op.P.bias = B.up.bias; st.op.P.bias_mod = Co;
      st.op.P.out_f32 = sb.x; st.op.P.out_f32_ld = (long long)s * Co;
      st.op.P.out_act = sb.a; st.op.P.out_act_ld = (long long)s * Co; st.op.P.snake_alpha = B.ru[0].a0;
      cp.flops += st.op.flops;
      cp.steps.push_back(st);
    }
    const bool last = (i == c.codec_n_rates - 1);
    const float* after = last ? e->dec_final_alpha : e->dec_blocks[i + 1].a_up;
    for (int j = 0; j < 3; ++j) plan_resunit(cp, B.ru[j], items, To, Co, sb, j < 2 ? B.ru[j + 1].a0 : after, j < 2);
    a_cur = sb.a; C = Co; Tn = To;
  }
  cp.a_last = a_cur;
  cp.S = Tn;
  { CodecStep s; s.kind = CodecStep::DEC_LAST; cp.steps.push_back(s); }
  cp.flops += 2.0 * items * (double)Tn * C * 7;
  cp.last_use = ++e->plan_clock;
  CodecPlan* r = CP.get();
  e->dec_plans[key] = std::move(CP);
  return r;
}

sab_engine::~sab_engine() {
  dit.reset();
  enc_plans.clear();
  dec_plans.clear();
  if (staging) cudaFree(staging);
  if (capture_stream) cudaStreamDestroy(capture_stream);
  for (auto ev : prof_events) cudaEventDestroy(ev);
}

// =====================================================================================================
// C ABI
// =====================================================================================================
#define SAB_API_BEGIN try {
#define SAB_API_BEGIN_E(e) try { SAB_CHECK((e) != nullptr, "null engine"); DeviceGuard _dg((e)->device);
#define SAB_API_END                                   \
  }                                                   \
  catch (const std::exception& ex) {                  \
    g_last_error = ex.what();                         \
    return 1;                                         \
  }                                                   \
  return 0;

extern "C" {

const char* sab_last_error(void) { return g_last_error.c_str(); }
int sab_version(void) { return 1; }

int sab_create(const sab_config* cfg, int device, sab_engine** out) {
  SAB_API_BEGIN
  SAB_CHECK(cfg && out, "null argument");
  int n_dev = 0;
  cudaError_t ce = cudaGetDeviceCount(&n_dev);
  SAB_CHECK(ce == cudaSuccess && n_dev > 0, "no CUDA device: the SAM-Audio B200 path has no CPU fallback (%s)",
            cudaGetErrorString(ce));
  SAB_CHECK(device >= 0 && device < n_dev && device < kMaxDevices, "bad device index %d", device);
  DeviceGuard _dg(device);
  cudaDeviceProp prop;
  SAB_CUDA(cudaGetDeviceProperties(&prop, device));
  SAB_CHECK(prop.major == 10, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  if (const char* f = getenv("SAB_FORCE_CG")) g_force_cg = atoi(f);
  if (const char* f = getenv("SAB_ATTN_POLY")) g_attn_poly = atoi(f);
  g_serpentine = getenv("SAB_NO_SERPENTINE") == nullptr;
  SAB_CHECK(cfg->dim % 128 == 0 && cfg->dim / cfg->n_heads == 128, "dim must be n_heads*128");
  SAB_CHECK(cfg->ffn_hidden % 64 == 0, "ffn_hidden must be a multiple of 64");
  SAB_CHECK(cfg->codec_n_rates >= 1 && cfg->codec_n_rates <= 8, "bad codec_n_rates");
  auto* e = new sab_engine();
  e->cfg = *cfg;
  e->device = device;
  try {
    register_weights(e);
    e->rope_len = cfg->max_positions;
    e->rope = reinterpret_cast<float2*>(e->wpool.alloc<float>((int64_t)e->rope_len * 64 * 2));
    rope_table_kernel<<<(e->rope_len * 64 + 255) / 256, 256>>>(e->rope, e->rope_len, 128, cfg->rope_theta);
    SAB_CUDA(cudaGetLastError());
    SAB_CUDA(cudaDeviceSynchronize());
  } catch (...) {
    delete e;
    throw;
  }
  *out = e;
  SAB_API_END
}

int sab_destroy(sab_engine* e) {
  SAB_API_BEGIN
  if (e) {
    DeviceGuard _dg(e->device);
    cudaDeviceSynchronize();
    delete e;
  }
  SAB_API_END
}

int sab_load_weight(sab_engine* e, const char* name, const float* data, const int64_t* shape, int ndim, int is_device,
                    void* stream) {
  SAB_API_BEGIN_E(e)
  SAB_CHECK(e && name && data, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  auto it = e->slots.find(name);
  SAB_CHECK(it != e->slots.end(), "unexpected weight '%s'", name);
  Slot& s = it->second;
  int64_t n = 1, n_expected = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  for (auto v : s.shape) n_expected *= v;
  std::vector<int64_t> sq_a, sq_b;  // compare with singleton dims squeezed ([1,C,1] alpha == [C])
  for (auto v : s.shape) if (v != 1) sq_a.push_back(v);
  for (int i = 0; i < ndim; ++i) if (shape[i] != 1) sq_b.push_back(shape[i]);
  SAB_CHECK(sq_a == sq_b, "weight '%s': shape mismatch (expected %lld elements, got %lld)", name,
            (long long)n_expected, (long long)n);
  const float* src = data;
  if (!is_device) {
    if (e->staging_elems < n) {
      if (e->staging) { SAB_CUDA(cudaStreamSynchronize(st)); SAB_CUDA(cudaFree(e->staging)); }
      SAB_CUDA(cudaMalloc(&e->staging, n * sizeof(float)));
      e->staging_elems = n;
    }
    SAB_CUDA(cudaMemcpyAsync(e->staging, data, n * sizeof(float), cudaMemcpyHostToDevice, st));
    src = e->staging;
  }
  s.load(src, st);
  if (!is_device) SAB_CUDA(cudaStreamSynchronize(st));  // staging buffer is reused by the next call
  s.loaded = true;
  e->finalized = false;
  if (e->dit) e->dit->ts_solve.valid = false;           // hoisted adaLN tables depend on the weights
  SAB_API_END
}

int sab_finalize_weights(sab_engine* e, int allow_missing, char* missing_out, int64_t missing_cap, void* stream) {
  SAB_API_BEGIN_E(e)
  cudaStream_t st = (cudaStream_t)stream;
  std::string missing, all_missing;
  int n_missing = 0;
  for (auto& kv : e->slots)
    if (!kv.second.loaded) {
      if (n_missing++ < 8) missing += kv.first + " ";
      all_missing += kv.first + "\n";
    }
  SAB_CHECK(n_missing == 0 || allow_missing, "Missing keys (%d): %s", n_missing, missing.c_str());
  if (missing_out && missing_cap > 0) {
    const size_t n = std::min<size_t>(all_missing.size(), (size_t)missing_cap - 1);
    memcpy(missing_out, all_missing.data(), n);
    missing_out[n] = 0;
  }
  const int d = e->cfg.dim;
  // constant video term for text-only prompts: LayerNorm(conv bias)   
  ln_vector_kernel<<<1, 256, 0, st>>>(e->vid_b, e->vid_ln_w, e->vid_ln_b, d, e->vid_const);
  // anchor table: embed [n,128] @ proj^T [128,d]
  const int n = e->cfg.n_anchor_tokens;
  small_abt_kernel<<<(n * d + 255) / 256, 256, 0, st>>>(e->anchor_embed, e->anchor_proj, e->anchor_table, n, d,
                                                         e->cfg.anchor_dim);
  SAB_CUDA(cudaGetLastError());
  SAB_CUDA(cudaStreamSynchronize(st));
  for (size_t l = 0; l < e->layers.size(); ++l) {
    SAB_CUDA(cudaMemcpyAsync(e->norm_w_all + (2 * l) * d, e->layers[l].attn_norm, d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SAB_CUDA(cudaMemcpyAsync(e->norm_w_all + (2 * l + 1) * d, e->layers[l].ffn_norm, d * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  SAB_CUDA(cudaStreamSynchronize(st));
  // self-attention logit bound per layer (see LayerW::att_shift_log2).  The single-pass softmax subtracts this
  // constant instead of the row maximum; it is used only while even a row whose keys are all anti-aligned keeps
  // its largest probability above 2^-100 (2 x shift <= 100), otherwise the layer runs the exact two-pass variant.
  for (auto& W : e->layers) {
    float qn[128], kn[128];
    SAB_CUDA(cudaMemcpy(qn, W.qn, sizeof(qn), cudaMemcpyDeviceToHost));
    SAB_CUDA(cudaMemcpy(kn, W.kn, sizeof(kn), cudaMemcpyDeviceToHost));
    float mq = 0.f, mk = 0.f;
    for (int i = 0; i < 128; ++i) { mq = fmaxf(mq, fabsf(qn[i])); mk = fmaxf(mk, fabsf(kn[i])); }
    const float sl2 = (1.0f / sqrtf(128.f)) * 1.4426950408889634f;
    for (int i = 0; i < 128; ++i) qn[i] *= sl2;
    SAB_CUDA(cudaMemcpy(W.qn_scaled, qn, sizeof(qn), cudaMemcpyHostToDevice));
    const float bound = sqrtf(128.f) * mq * mk * 1.4426950408889634f * 1.02f + 0.25f;   // + bf16 rounding of q, k
    W.att_shift_log2 = (std::isfinite(bound) && bound <= 50.f && !getenv("SAB_ATTN_EXACT")) ? bound : -1.f;
  }
  e->finalized = true;
  SAB_API_END
}

int sab_prepare(sab_engine* e, int B, int candidates, int T, int L, const float* features, const float* text_features,
                const uint8_t* text_mask, const float* video_features, const int64_t* anchor_ids, int n_ids,
                const int64_t* anchor_alignment, const uint8_t* audio_pad_mask, int flags, void* stream) {
  SAB_API_BEGIN_E(e)
  SAB_CHECK(e->finalized, "weights not finalized");
  SAB_CHECK(B > 0 && candidates > 0 && T > 0 && L > 0, "bad shape");
  SAB_CHECK(features && text_mask && audio_pad_mask, "null argument");
  const bool no_text = (flags & SAB_PREP_NO_TEXT) != 0, no_anchor = (flags & SAB_PREP_NO_ANCHORS) != 0;
  SAB_CHECK(no_text ? L == 1 : text_features != nullptr, "text_features missing (or SAB_PREP_NO_TEXT with L != 1)");
  SAB_CHECK(no_anchor || (anchor_ids && anchor_alignment && n_ids > 0), "anchor tensors missing");
  cudaStream_t st = (cudaStream_t)stream;
  if (!e->dit || e->dit->B != B || e->dit->cand != candidates || e->dit->T != T || e->dit->L != L ||
      n_ids > e->dit->n_ids_cap) {
    SAB_CUDA(cudaStreamSynchronize(st));
    e->dit.reset();
    build_dit_plan(e, B, candidates, T, L, std::max(64, n_ids));   // the anchor-id table grows with the longest id list
  }
  DitPlan& p = *e->dit;
  const sab_config& c = e->cfg;
  const int d = c.dim;
  SAB_CHECK(n_ids <= p.n_ids_cap, "too many anchors per clip (%d > %d)", n_ids, p.n_ids_cap);
  const long long M = p.M, MB = p.MB, ML = p.ML;
  SAB_CUDA(cudaMemcpyAsync(p.pad_mask, audio_pad_mask, MB, cudaMemcpyDeviceToDevice, st));
  SAB_CUDA(cudaMemcpyAsync(p.text_mask, text_mask, ML, cudaMemcpyDeviceToDevice, st));
  if (!no_anchor) {
    SAB_CUDA(cudaMemcpyAsync(p.anchor_ids, anchor_ids, (size_t)B * n_ids * 8, cudaMemcpyDeviceToDevice, st));
    SAB_CUDA(cudaMemcpyAsync(p.anchor_align, anchor_alignment, (size_t)MB * 8, cudaMemcpyDeviceToDevice, st));
  }
  mark(e, st, "cast_bf16_kernel");
  cast_bf16_kernel<<<512, 256, 0, st>>>(features, p.feat_bf, MB * 256);
  gemm(e, p.g_cond, st);
  if (no_text) {   // model.py:170-172 with text_features=None: the memory is the time embedding alone
    SAB_CUDA(cudaMemsetAsync(p.mem_base, 0, (size_t)ML * d * sizeof(float), st));
  } else {
    mark(e, st, "cast_bf16_kernel");
    cast_bf16_kernel<<<64, 256, 0, st>>>(text_features, p.text_bf, ML * c.text_dim);
    gemm(e, p.g_mem, st);
  }
  const float* vproj = nullptr;
  if (video_features) {
    mark(e, st, "transpose_cast_kernel");
    transpose_cast_kernel<<<1024, 256, 0, st>>>(video_features, c.vision_dim, T, MB * c.vision_dim, p.vid_bf);
    gemm(e, p.g_vid, st);
    vproj = p.vproj;
  }
  mark(e, st, "cond_finish_kernel");
  cond_finish_kernel<<<(int)((M + 7) / 8), 256, 0, st>>>(p.cond, p.condB, (int)M, d, T, candidates, vproj, e->vid_ln_w,
                                                        e->vid_ln_b, e->vid_const, e->vid_gate, e->anchor_table,
                                                        reinterpret_cast<const long long*>(p.anchor_ids), n_ids,
                                                        reinterpret_cast<const long long*>(p.anchor_align), e->anchor_gate,
                                                        (flags & SAB_PREP_NO_VIDEO_TERM) ? 0 : 1, no_anchor ? 0 : 1);
  SAB_CUDA(cudaGetLastError());
  SAB_API_END
}

int sab_dit_forward(sab_engine* e, const float* noisy, const float* time, float* velocity, void* stream) {
  SAB_API_BEGIN_E(e)
  SAB_CHECK(e && e->dit, "sab_prepare must be called first");
  SAB_CHECK(e->dit->cand == 1, "sab_dit_forward takes per-sequence times: prepare with candidates == 1");
  cudaStream_t st = (cudaStream_t)stream;
  DitPlan& p = *e->dit;
  mark(e, st, "cast_bf16_kernel");
  cast_bf16_kernel<<<512, 256, 0, st>>>(noisy, p.y_bf, p.M * 256);
  FinalSpec fs{nullptr, 1.f, velocity, nullptr};
  if (p.ts_item.R == 0) build_time_state(e, p, p.ts_item, p.Bc);
  SAB_CUDA(cudaMemcpyAsync(p.ts_item.time, time, (size_t)p.Bc * sizeof(float), cudaMemcpyDeviceToDevice, st));
  run_time_state(e, p, p.ts_item, st);
  dit_eval(e, time, fs, p.ts_item, -1, st);
  SAB_API_END
}

// stage times (fractions of a step) of the fixed-grid solvers torchdiffeq offers and the reference forwards
// `ode_opt` to (model.py:285-290): euler, midpoint, rk4 (torchdiffeq's rk4 is the 3/8 rule)
static int solver_stages(int method, float (&frac)[4]) {
  switch (method) {
    case SAB_ODE_EULER: frac[0] = 0.f; return 1;
    case SAB_ODE_MIDPOINT: frac[0] = 0.f; frac[1] = 0.5f; return 2;
    case SAB_ODE_RK4: frac[0] = 0.f; frac[1] = 1.f / 3.f; frac[2] = 2.f / 3.f; frac[3] = 1.f; return 4;
    default: throw Error(fmt("unknown ODE method %d (0 midpoint, 1 euler, 2 rk4)", method));
  }
}

int sab_solve(sab_engine* e, const float* noise, int n_steps, int method, float* latent, void* stream) {
  SAB_API_BEGIN_E(e)
  SAB_CHECK(e && e->dit, "sab_prepare must be called first");
  float frac[4];
  const int S = solver_stages(method, frac);
  const int E = S * n_steps;                   // evaluations per solve
  SAB_CHECK(n_steps >= 1 && E <= 128, "n_steps out of range");
  cudaStream_t st = (cudaStream_t)stream;
  DitPlan& p = *e->dit;
  const long long n = p.M * 256;
  // evaluation times (k + frac_s)/n, broadcast over the batch (model.py:280 t.expand); uploaded once per plan
  if (p.time_steps_uploaded != n_steps || p.solve_method != method) {
    std::vector<float> times((size_t)E * p.Bc);
    for (int k = 0; k < n_steps; ++k)
      for (int sg = 0; sg < S; ++sg)
        for (int b = 0; b < p.Bc; ++b) times[(size_t)(S * k + sg) * p.Bc + b] = ((float)k + frac[sg]) / (float)n_steps;
    if (p.time_cap < (int64_t)times.size()) {
      SAB_CUDA(cudaStreamSynchronize(st));
      p.time_dev = p.pool.alloc<float>((int64_t)times.size());
      p.time_cap = (int64_t)times.size();
    }
    SAB_CUDA(cudaMemcpyAsync(p.time_dev, times.data(), times.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    SAB_CUDA(cudaStreamSynchronize(st));  // `times` is a stack-owned host buffer
    p.time_steps_uploaded = n_steps;
  }
  // The adaLN state depends on the evaluation time only and every sequence of a solve shares it: one row per
  // evaluation, computed once per (plan, solver, step count, weights) instead of inside the loop.
  if (p.ts_solve.R != E || p.solve_method != method) {
    if (p.ts_solve.R != 0) SAB_CUDA(cudaStreamSynchronize(st));
    p.ts_solve = TimeState();     // a new grid: new tables (the old buffers stay with the plan's pool)
    build_time_state(e, p, p.ts_solve, E);
  }
  if (!p.ts_solve.valid) {
    std::vector<float> tk((size_t)E);
    for (int k = 0; k < n_steps; ++k)
      for (int sg = 0; sg < S; ++sg) tk[S * k + sg] = ((float)k + frac[sg]) / (float)n_steps;
    SAB_CUDA(cudaMemcpyAsync(p.ts_solve.time, tk.data(), tk.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    SAB_CUDA(cudaStreamSynchronize(st));
    run_time_state(e, p, p.ts_solve, st);
    p.ts_solve.valid = true;
  }
  if (method == SAB_ODE_RK4 && !p.rk[0])
    for (int i = 0; i < 4; ++i) p.rk[i] = p.pool.alloc<float>(n);
  const bool new_method = p.solve_method != method;
  p.solve_method = method;
  SAB_CUDA(cudaMemcpyAsync(p.y, noise, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  auto enqueue = [&](cudaStream_t st) {
    mark(e, st, "cast_bf16_kernel");
    cast_bf16_kernel<<<512, 256, 0, st>>>(p.y, p.y_bf, n);
    const float dt = 1.0f / (float)n_steps;
    auto tm = [&](int ev) { return p.time_dev + (size_t)ev * p.Bc; };
    for (int k = 0; k < n_steps; ++k) {
      if (method == SAB_ODE_MIDPOINT) {
        // f0 = f(t_k, y); y_mid = y + f0*dt/2     |     y += dt * f(t_k + dt/2, y_mid)   (axpy fused in the output GEMM)
        FinalSpec a{p.y, 0.5f * dt, p.ymid, p.y_bf};
        dit_eval(e, tm(2 * k), a, p.ts_solve, 2 * k, st);
        FinalSpec b{p.y, dt, p.y, p.y_bf};
        dit_eval(e, tm(2 * k + 1), b, p.ts_solve, 2 * k + 1, st);
      } else if (method == SAB_ODE_EULER) {
        FinalSpec a{p.y, dt, p.y, p.y_bf};         // y += dt * f(t_k, y)
        dit_eval(e, tm(k), a, p.ts_solve, k, st);
      } else {
        // torchdiffeq rk4 (3/8 rule): k1 = f(t, y); k2 = f(t + dt/3, y + dt k1/3); k3 = f(t + 2dt/3, y + dt (k2 - k1/3));
        // k4 = f(t + dt, y + dt (k1 - k2 + k3)); y += dt (k1 + 3 k2 + 3 k3 + k4) / 8.  The stage inputs need the raw
        // velocities, so they are combined by a small elementwise kernel instead of the output GEMM's epilogue.
        const float c3 = dt / 3.f;
        auto comb = [&](float* out, float a1, float a2, float a3, float a4) {
          mark(e, st, "ode_combine_kernel");
          ode_combine_kernel<<<512, 256, 0, st>>>(p.y, p.rk[0], a1, p.rk[1], a2, p.rk[2], a3, p.rk[3], a4, out, p.y_bf, n);
        };
        for (int sg = 0; sg < 4; ++sg) {
          FinalSpec v{nullptr, 1.f, p.rk[sg], nullptr};
          dit_eval(e, tm(4 * k + sg), v, p.ts_solve, 4 * k + sg, st);
          if (sg == 0) comb(p.ymid, c3, 0.f, 0.f, 0.f);
          else if (sg == 1) comb(p.ymid, -c3, dt, 0.f, 0.f);
          else if (sg == 2) comb(p.ymid, dt, -dt, dt, 0.f);
          else comb(p.y, dt * 0.125f, dt * 0.375f, dt * 0.375f, dt * 0.125f);
        }
      }
    }
  };
  static const bool use_graph = !getenv("SAB_NO_GRAPH");
  const bool graphable = use_graph && !e->prof && p.solves >= 1;
  if (graphable && (!p.solve_graph || p.solve_graph_steps != n_steps || new_method)) {
    if (p.solve_graph) { SAB_CUDA(cudaGraphExecDestroy(p.solve_graph)); p.solve_graph = nullptr; }
    const int64_t launches_before = e->launches;
    cudaGraph_t g = nullptr;
    if (!e->capture_stream) SAB_CUDA(cudaStreamCreateWithFlags(&e->capture_stream, cudaStreamNonBlocking));
    SAB_CUDA(cudaStreamBeginCapture(e->capture_stream, cudaStreamCaptureModeThreadLocal));
    try {
      enqueue(e->capture_stream);
    } catch (...) {
      cudaStreamEndCapture(e->capture_stream, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    SAB_CUDA(cudaStreamEndCapture(e->capture_stream, &g));
    SAB_CUDA(cudaGraphInstantiate(&p.solve_graph, g, 0));
    SAB_CUDA(cudaGraphDestroy(g));
    p.solve_graph_steps = n_steps;
    p.solve_graph_launches = e->launches - launches_before;
    e->launches = launches_before;   // counted when the graph is launched
  }
  if (graphable) {
    SAB_CUDA(cudaGraphLaunch(p.solve_graph, st));
    e->launches += p.solve_graph_launches;
  } else {
    enqueue(st);
  }
  p.solves++;
  SAB_CUDA(cudaMemcpyAsync(latent, p.y, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  SAB_API_END
}

static void run_codec(sab_engine* e, CodecPlan& cp, int items, const float* wav_in, const float* latent_in, int T,
                      float* out, cudaStream_t st) {
  const sab_config& c = e->cfg;
  for (auto& s : cp.steps) {
    switch (s.kind) {
      case CodecStep::ENC0: {
        const int C0 = c.codec_encoder_dim;
        const long long n = cp.S * (C0 / 4);
        mark(e, st, "codec.enc.conv0", 2.0 * items * (double)cp.S * C0 * 7, (double)items * cp.S * (4.0 + C0 * 6.0));
        if (C0 % 4 == 0 && 256 % (C0 / 4) == 0) {
          const long long per_block = (256 / (C0 / 4)) * ENC0_ITER;
          enc_conv0_reg_kernel<<<dim3((unsigned)((cp.S + per_block - 1) / per_block), items), 256, 0, st>>>(
              wav_in, cp.S, C0, e->enc0_w, e->enc0_b, e->enc_blocks[0].ru[0].a0, cp.x_first, cp.a_first);
        } else {
          enc_conv0_kernel<<<dim3((unsigned)((n + 255) / 256), items), 256, 0, st>>>(
              wav_in, cp.S, C0, e->enc0_w, e->enc0_b, e->enc_blocks[0].ru[0].a0, cp.x_first, cp.a_first);
        }
        break;
      }
      case CodecStep::LATENT_SPLIT: {
        const int cz = c.codec_codebook_dim;
        mark(e, st, "latent_split_kernel");
        latent_split_kernel<<<512, 256, 0, st>>>(latent_in, T, cz, (long long)items * T * cz, cp.z_in);
        break;
      }
      case CodecStep::DEC_LAST: {
        const int C = c.codec_decoder_dim >> c.codec_n_rates;
        const int smem = ((DEC_LAST_TB + 6) * (C + 2) + 2) * 2 + 7 * C * 4;
        mark(e, st, "codec.dec.last", 2.0 * items * (double)cp.S * C * 7, (double)items * cp.S * (4.0 + C * 2.0));
        const dim3 rgrid((unsigned)((cp.S + 32 * DEC_SEG - 1) / (32 * DEC_SEG)), items);
        if (C == 96) dec_last_reg_kernel<3><<<rgrid, 256, 0, st>>>(cp.a_last, cp.S, e->dec_last_w, e->dec_last_b, out);
        else if (C == 64) dec_last_reg_kernel<2><<<rgrid, 256, 0, st>>>(cp.a_last, cp.S, e->dec_last_w, e->dec_last_b, out);
        else if (C == 128) dec_last_reg_kernel<4><<<rgrid, 256, 0, st>>>(cp.a_last, cp.S, e->dec_last_w, e->dec_last_b, out);
        else
          dec_last_kernel<<<dim3((unsigned)((cp.S + DEC_LAST_TB - 1) / DEC_LAST_TB), items), DEC_LAST_TB, smem, st>>>(
              cp.a_last, cp.S, C, e->dec_last_w, e->dec_last_b, out);
        break;
      }
      default: {
        GemmOp op = s.op;
        op.P.n_items = items;
        if (&s.op == cp.enc_out) op.P.out_f32 = out;
        const long long tiles = (long long)items * op.P.tiles_per_item * op.P.n_tiles_n;
        op.grid = (int)std::min<long long>(tiles, sm_count() / op.cg) * op.cg;
        op.flops = s.op.flops * (double)items / (double)cp.items;
        op.rows = s.op.rows * (double)items / (double)cp.items;
        op.in_bytes = s.op.in_bytes * (double)items / (double)cp.items;   // (the weight share is negligible here)
        gemm(e, op, st);
      }
    }
  }
  SAB_CUDA(cudaGetLastError());
}

static int codec_chunk(const sab_engine* e, long long S_samples, bool decoder) {
  // fp32 stream + two bf16 operand buffers per stage (8 B per activation element): ~0.56 GB per 10 s clip in the
  // encoder, ~0.85 GB per 10 s waveform in the decoder; keep a chunk's workspace near 12 GB
  (void)e;
  const double per_item = (double)S_samples / 480000.0 * (decoder ? 0.85e9 : 0.56e9);
  double budget = 12e9;
  if (const char* b = getenv("SAB_CODEC_CHUNK_BYTES")) budget = atof(b);   // tests force small chunks through this
  int n = (int)(budget / per_item);
  return n < 1 ? 1 : (n > 64 ? 64 : n);
}

int sab_encode(sab_engine* e, const float* wav, int B, int64_t S, float* features, void* stream) {
  SAB_API_BEGIN_E(e)
  SAB_CHECK(e && e->finalized, "weights not finalized");
  cudaStream_t st = (cudaStream_t)stream;
  long long hop = 1;
  for (int i = 0; i < e->cfg.codec_n_rates; ++i) hop *= e->cfg.codec_encoder_rates[i];
  SAB_CHECK(S % hop == 0, "S=%lld must be padded to a multiple of hop=%lld (codec.py:72-78)", (long long)S, hop);
  const long long T = S / hop;
  const int chunk = std::min(B, codec_chunk(e, S, false));
  CodecPlan* cp = get_enc_plan(e, chunk, S);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = std::min(chunk, B - b0);
    run_codec(e, *cp, n, wav + (long long)b0 * S, nullptr, (int)T,
              features + (long long)b0 * T * 2 * e->cfg.codec_codebook_dim, st);
  }
  SAB_API_END
}

int sab_decode(sab_engine* e, const float* latent, int Bc, int T, float* wav, void* stream) {
  SAB_API_BEGIN_E(e)
  SAB_CHECK(e && e->finalized, "weights not finalized");
  cudaStream_t st = (cudaStream_t)stream;
  long long hop = 1;
  for (int i = 0; i < e->cfg.codec_n_rates; ++i) hop *= e->cfg.codec_decoder_rates[i];
  const long long S = (long long)T * hop;
  const int items = 2 * Bc;
  // whole clips per chunk: items (2b, 2b+1) = (target, residual) come from the two halves of latent row b, so a chunk
  // is always an even number of items, at least one clip (a single-item chunk would decode the target half twice)
  int chunk = std::min(items, std::max(2, codec_chunk(e, S, true) & ~1));
  CodecPlan* cp = get_dec_plan(e, chunk, T);
  const int cz = e->cfg.codec_codebook_dim;
  for (int i0 = 0; i0 < items; i0 += chunk) {
    const int n = std::min(chunk, items - i0);
    // items (2b, 2b+1) live in latent row b: chunk boundaries are clip boundaries
    run_codec(e, *cp, n, nullptr, latent + (long long)(i0 / 2) * T * 2 * cz, T, wav + (long long)i0 * S, st);
  }
  SAB_API_END
}

int64_t sab_launch_count(sab_engine* e, int reset) {
  if (!e) return -1;
  const int64_t v = e->launches;
  if (reset) e->launches = 0;
  return v;
}

int64_t sab_workspace_bytes(sab_engine* e) {
  if (!e) return -1;
  int64_t b = e->wpool.bytes;
  if (e->dit) b += e->dit->pool.bytes;
  for (auto& kv : e->enc_plans) b += kv.second->pool.bytes;
  for (auto& kv : e->dec_plans) b += kv.second->pool.bytes;
  return b;
}

int sab_profile(sab_engine* e, int enable, void* stream) {
  SAB_API_BEGIN_E(e)
  SAB_CHECK(e, "null engine");
  SAB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  e->prof = enable != 0;
  e->prof_recs.clear();
  SAB_API_END
}

// JSON: {"tag": {"launches": n, "ms": t, "flops": f, "bytes": b}, ...} aggregated since sab_profile(e, 1).
int sab_profile_report(sab_engine* e, char* buf, int64_t cap, void* stream) {
  SAB_API_BEGIN_E(e)
  SAB_CHECK(e && buf && cap > 2, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = e->prof_recs.size();
  std::map<std::string, std::array<double, 4>> agg;
  if (n > 0) {
    if (e->prof_events.size() <= n) {
      cudaEvent_t ev;
      SAB_CUDA(cudaEventCreate(&ev));
      e->prof_events.push_back(ev);
    }
    SAB_CUDA(cudaEventRecord(e->prof_events[n], st));
    SAB_CUDA(cudaStreamSynchronize(st));
    for (size_t i = 0; i < n; ++i) {
      float ms = 0.f;
      SAB_CUDA(cudaEventElapsedTime(&ms, e->prof_events[i], e->prof_events[i + 1]));
      auto& a = agg[e->prof_recs[i].tag];
      a[0] += 1; a[1] += ms; a[2] += e->prof_recs[i].flops; a[3] += e->prof_recs[i].bytes;
    }
  }
  std::string js = "{";
  bool first = true;
  for (auto& kv : agg) {
    js += fmt("%s\"%s\": {\"launches\": %.0f, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
              kv.first.c_str(), kv.second[0], kv.second[1], kv.second[2], kv.second[3]);
    first = false;
  }
  js += "}";
  SAB_CHECK((int64_t)js.size() + 1 <= cap, "profile buffer too small (%zu needed)", js.size() + 1);
  memcpy(buf, js.c_str(), js.size() + 1);
  e->prof_recs.clear();
  SAB_API_END
}

// ---- test seams ----
int sab_test_gemm(int M, int N, int K, const void* a_bf16, const void* b_bf16, float* c, int bn, int bk, int cg,
                  void* stream) {
  SAB_API_BEGIN
  RunList rl;
  SAB_CHECK(K % bk == 0, "K must be a multiple of bk");
  rl.add(0, 0, 0, K / bk);
  GemmOp op = make_gemm("test", flat_view((const bf16*)a_bf16, M, K), (const bf16*)b_bf16, N, bn, bk, EPI_AFFINE, rl, cg);
  op.P.out_f32 = c; op.P.out_f32_ld = N;
  launch_gemm(op, (cudaStream_t)stream);
  SAB_API_END
}

int sab_test_attention(int items, int heads, int Tq, int Tk, const void* q, const void* k, const void* v,
                       const uint8_t* key_mask, void* o, void* stream) {
  SAB_API_BEGIN
  ensure_dynamic_smem(reinterpret_cast<const void*>(attention_kernel), ATT_SMEM);
  AttnParams a{};
  const long long ld = (long long)heads * 128;
  a.q = (const bf16*)q; a.q_ld = ld; a.k = (const bf16*)k; a.k_ld = ld; a.v = (const bf16*)v; a.v_ld = ld;
  a.o = (bf16*)o; a.o_ld = ld; a.key_mask = key_mask; a.Tq = Tq; a.Tk = Tk;
  a.scale_log2 = (1.0f / sqrtf(128.f)) * 1.4426950408889634f;
  if (Tk <= XATT_MAX_TK && k != q) {
    launch_xattn_small(a, items, heads, (cudaStream_t)stream);
  } else {
    dim3 grid((Tq + ATT_BQ - 1) / ATT_BQ, heads, items);
    attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, (cudaStream_t)stream>>>(a);
  }
  SAB_CUDA(cudaGetLastError());
  SAB_API_END
}

int sab_test_attention_tc(int items, int heads, int T, const void* q, const void* k, const void* v,
                          const uint8_t* key_mask, void* o, int v_lbo, int v_sbo, void* stream) {
  SAB_API_BEGIN
  SAB_CHECK(T <= 256, "tcgen05 attention handles T <= 256");
  const long long ld = (long long)heads * 128;
  CUtensorMap tq = make_tmap_3d(q, ld, T, items, ld, (int64_t)T * ld, 64, 128);
  CUtensorMap tk = make_tmap_3d(k, ld, T, items, ld, (int64_t)T * ld, 64, 256);
  CUtensorMap tv = make_tmap_3d(v, ld, T, items, ld, (int64_t)T * ld, 64, 256);
  AttnTcParams tp{};
  tp.o = (bf16*)o; tp.o_ld = ld; tp.key_mask = key_mask; tp.Tq = T; tp.Tk = T; tp.heads = heads;
  tp.q_col0 = tp.k_col0 = tp.v_col0 = 0;
  tp.scale_log2 = (1.0f / sqrtf(128.f)) * 1.4426950408889634f;
  tp.v_lbo = v_lbo > 0 ? v_lbo : ATC_KV_BYTES / 2;
  tp.v_sbo = v_sbo > 0 ? v_sbo : 1024;
  launch_attention_tc(tq, tk, tv, tp, items, (cudaStream_t)stream);
  SAB_API_END
}

int sab_test_attention_tc2(int items, int heads, int T, const void* q, const void* k, const void* v,
                           const uint8_t* key_mask, void* o, float shift_log2, int poly, long long* trace, void* stream) {
  SAB_API_BEGIN
  SAB_CHECK(T <= 256 && T >= 1, "tcgen05 attention handles 1 <= T <= 256");
  const long long ld = (long long)heads * 128;
  CUtensorMap tq = make_tmap_3d(q, ld, T, items, ld, (int64_t)T * ld, 64, 128);
  CUtensorMap tk = make_tmap_3d(k, ld, T, items, ld, (int64_t)T * ld, 64, 256);
  CUtensorMap tv = make_tmap_3d(v, ld, T, items, ld, (int64_t)T * ld, 64, 256);
  CUtensorMap to = make_tmap_3d(o, ld, T, items, ld, (int64_t)T * ld, 64, 128);
  AttnTc2Params tp{};
  tp.key_mask = key_mask; tp.T = T; tp.heads = heads; tp.items = items; tp.mask_div = 1;
  tp.q_col0 = tp.k_col0 = tp.v_col0 = 0;
  tp.scale_log2 = (1.0f / sqrtf(128.f)) * 1.4426950408889634f;
  tp.shift_log2 = shift_log2;
  tp.trace = trace;
  const int saved = g_attn_poly;
  if (poly >= 0) g_attn_poly = poly;
  try {
    // shift_log2 == 0: folded path (the caller passes q already multiplied by log2(e)/sqrt(hd), |logit| <= 50)
    if (shift_log2 == 0.f) tp.shift_log2 = 50.f;
    launch_attention_tc2(tq, tk, tv, to, tp, shift_log2 == 0.f, (cudaStream_t)stream);
  } catch (...) {
    g_attn_poly = saved;
    throw;
  }
  g_attn_poly = saved;
  SAB_API_END
}

// ---- visual prompting: frame pre-processing (vision_kernels.cuh) ----
// tap windows + normalised weights of one axis, fp32 arithmetic of ATen's _compute_indices_min_size_weights_aa
// (cubic a = -1/2, antialias: support and filter stretch by scale = in / out when down-sampling)
struct AaTaps { int taps = 0; std::vector<int> lo, cnt; std::vector<float> w; };
static AaTaps aa_taps(int in_size, int out_size) {
  AaTaps t;
  const float scale = (float)in_size / (float)out_size;
  const float support = scale >= 1.f ? 2.0f * scale : 2.0f;
  const float invscale = scale >= 1.f ? 1.0f / scale : 1.0f;
  t.taps = (int)ceilf(support) * 2 + 1;
  t.lo.resize(out_size); t.cnt.resize(out_size); t.w.assign((size_t)out_size * t.taps, 0.f);
  const float a = -0.5f;
  for (int i = 0; i < out_size; ++i) {
    const float center = scale * ((float)i + 0.5f);
    int lo = (int)(center - support + 0.5f);
    lo = lo < 0 ? 0 : lo;
    int hi = (int)(center + support + 0.5f);
    hi = hi > in_size ? in_size : hi;
    int n = hi - lo;
    n = n < 0 ? 0 : (n > t.taps ? t.taps : n);
    float tot = 0.f;
    float* w = &t.w[(size_t)i * t.taps];
    for (int j = 0; j < n; ++j) {
      float x = fabsf(((float)(j + lo) - center + 0.5f)) * invscale;
      float v;
      if (x < 1.f) v = ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
      else if (x < 2.f) v = (((x - 5.f) * x + 8.f) * x - 4.f) * a;
      else v = 0.f;
      w[j] = v;
      tot = tot + v;
    }
    for (int j = 0; j < n; ++j) w[j] = w[j] / tot;
    t.lo[i] = lo; t.cnt[i] = n;
  }
  return t;
}
struct AaTapsDev { int taps; int *lo, *cnt; float* w; };
static AaTapsDev aa_taps_device(int in_size, int out_size) {   // cached per (device, in, out): a few KB each
  static std::map<std::array<int, 3>, AaTapsDev> cache;
  const std::array<int, 3> key = {cur_device(), in_size, out_size};
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const AaTaps h = aa_taps(in_size, out_size);
  AaTapsDev d{h.taps, nullptr, nullptr, nullptr};
  SAB_CUDA(cudaMalloc(&d.lo, out_size * sizeof(int)));
  SAB_CUDA(cudaMalloc(&d.cnt, out_size * sizeof(int)));
  SAB_CUDA(cudaMalloc(&d.w, h.w.size() * sizeof(float)));
  SAB_CUDA(cudaMemcpy(d.lo, h.lo.data(), out_size * sizeof(int), cudaMemcpyHostToDevice));
  SAB_CUDA(cudaMemcpy(d.cnt, h.cnt.data(), out_size * sizeof(int), cudaMemcpyHostToDevice));
  SAB_CUDA(cudaMemcpy(d.w, h.w.data(), h.w.size() * sizeof(float), cudaMemcpyHostToDevice));
  cache[key] = d;
  return d;
}

/* host-only test seam: the evaluation times sab_solve uploads for a solver and step count (no GPU needed) */
int sab_test_solver_grid(int method, int n_steps, int cap, int* n_evals, float* times) {
  SAB_API_BEGIN
  SAB_CHECK(n_evals && times && n_steps >= 1, "bad argument");
  float frac[4];
  const int S = solver_stages(method, frac);
  SAB_CHECK(S * n_steps <= cap, "capacity %d < %d", cap, S * n_steps);
  *n_evals = S * n_steps;
  for (int k = 0; k < n_steps; ++k)
    for (int sg = 0; sg < S; ++sg) times[S * k + sg] = ((float)k + frac[sg]) / (float)n_steps;
  SAB_API_END
}

/* host-only test seam: the tap windows / weights the resize kernels use (no GPU needed) */
int sab_test_aa_taps(int in_size, int out_size, int cap, int* taps, int* lo, int* cnt, float* w) {
  SAB_API_BEGIN
  SAB_CHECK(in_size >= 1 && out_size >= 1 && taps && lo && cnt && w, "bad argument");
  const AaTaps t = aa_taps(in_size, out_size);
  SAB_CHECK(t.taps <= cap, "tap capacity %d < %d", cap, t.taps);
  *taps = t.taps;
  for (int i = 0; i < out_size; ++i) {
    lo[i] = t.lo[i];
    cnt[i] = t.cnt[i];
    for (int j = 0; j < cap; ++j) w[(size_t)i * cap + j] = j < t.taps ? t.w[(size_t)i * t.taps + j] : 0.f;
  }
  SAB_API_END
}

int sab_preprocess_frames(const uint8_t* frames, int n_frames, int H, int W, int out_size, float* workspace, float* out,
                          void* stream) {
  SAB_API_BEGIN
  SAB_CHECK(frames && out && workspace, "null argument");
  SAB_CHECK(n_frames >= 1 && H >= 1 && W >= 1 && out_size >= 1 && W <= 48 * 1024, "bad frame shape");
  cudaStream_t st = (cudaStream_t)stream;
  const AaTapsDev tx = aa_taps_device(W, out_size), ty = aa_taps_device(H, out_size);
  const int planes = n_frames * 3;
  resize_rows_kernel<<<dim3(H, planes), 256, W, st>>>(frames, H, W, out_size, tx.lo, tx.cnt, tx.w, tx.taps, workspace);
  resize_cols_finish_kernel<<<dim3(out_size, planes), 256, 0, st>>>(workspace, H, out_size, ty.lo, ty.cnt, ty.w, ty.taps, out);
  SAB_CUDA(cudaGetLastError());
  SAB_API_END
}

}  // extern "C"

#include "t5_engine.inc"
