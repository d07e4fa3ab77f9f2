/* samaudio_b200.h — C ABI of libsamaudio_b200.so: the B200 (sm_100a) implementation of the
 * SAMAudio.separate() inference hot path of facebookresearch/sam-audio.
 *
 * The reference has no FFI: its seam is the Python class surface.  Each entry point
 * below names the reference method it replaces (paths relative to the reference tree); the Python
 * host mirror (sam_audio_b200/model.py) and INTEGRATION.md show the binding.
 *
 * Conventions
 *  - plain C types only; every device pointer is BORROWED (the caller — PyTorch in the Python mirror —
 *    owns activations and I/O buffers); the library owns its packed weights and workspace.
 *  - every function returns 0 on success, non-zero on failure; sab_last_error() gives the message
 *    (reference convention is Python exceptions; the mirror raises RuntimeError from it).
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *    unless stated.  A handle is not thread-safe.
 *  - layouts: "rows x cols" row-major, fp32 unless stated; T = latent frames (25 Hz), S = samples.
 */
#ifndef SAMAUDIO_B200_H
#define SAMAUDIO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sab_engine sab_engine;

/* Shapes of the model (reference: sam_audio/model/config.py:86-130 TransformerConfig,
 * :10-41 DACVAEConfig, :204-231 SAMAudioConfig). */
typedef struct sab_config {
  int32_t dim;              /* transformer.dim (multiple of 128)            */
  int32_t n_heads;          /* transformer.n_heads (head_dim must be 128)   */
  int32_t n_layers;         /* transformer.n_layers                         */
  int32_t ffn_hidden;       /* SwiGLU hidden width (transformer.py:179-185) */
  int32_t out_channels;     /* 256                                          */
  int32_t in_channels;      /* SAMAudioConfig.in_channels = 768             */
  int32_t text_dim;         /* 768                                          */
  int32_t vision_dim;       /* 1024                                         */
  int32_t n_anchor_tokens;  /* num_anchors + 1 = 4                          */
  int32_t anchor_dim;       /* 128                                          */
  int32_t max_positions;    /* RoPE table length (10000)                    */
  float   rope_theta;       /* max(10000, 2*max_positions)                  */
  float   norm_eps;         /* 1e-5                                         */
  /* DAC-VAE codec */
  int32_t codec_encoder_dim;      /* 64   */
  int32_t codec_latent_dim;       /* 1024 */
  int32_t codec_decoder_dim;      /* 1536 */
  int32_t codec_codebook_dim;     /* 128  */
  int32_t codec_n_rates;          /* 4    */
  int32_t codec_encoder_rates[8]; /* 2,8,10,12 */
  int32_t codec_decoder_rates[8]; /* 12,10,8,2 */
} sab_config;

const char* sab_last_error(void);
int sab_version(void);

/* Lifecycle.  replaces: SAMAudio.__init__ (model.py:79-102). */
int sab_create(const sab_config* cfg, int device, sab_engine** out);
int sab_destroy(sab_engine* e);

/* Weights.  replaces: BaseModel._from_pretrained / load_state_dict (base.py:47-61, model.py:346-359).
 * `name` is the reference state-dict key (e.g. "transformer.layers.0.attention.wq.weight"; codec keys
 * "audio_codec.encoder.block.0.weight", weight-norm already folded).  `data` is fp32, contiguous, on host
 * (is_device=0) or device (is_device=1); it is consumed (repacked to bf16 / permuted) before returning.
 * Unknown names fail.  sab_finalize_weights() checks completeness and builds derived tables; with
 * allow_missing != 0 (load_state_dict(strict=False)) missing tensors keep their zero initialisation and their
 * names are written, newline-separated, into missing_out (may be NULL). */
int sab_load_weight(sab_engine* e, const char* name, const float* data, const int64_t* shape, int ndim,
                    int is_device, void* stream);
int sab_finalize_weights(sab_engine* e, int allow_missing, char* missing_out, int64_t missing_cap, void* stream);

/* Codec analysis.  replaces: DACVAE.forward (codec.py:65-78) + SAMAudio._get_audio_features (model.py:182-184).
 * wav [B, S] mono fp32 (S already padded to a multiple of hop by the caller = codec.py:72-78);
 * features [B, T, 2*codebook_dim] fp32: the 128-d mean latent duplicated along channels. */
int sab_encode(sab_engine* e, const float* wav, int B, int64_t S, float* features, void* stream);

/* Conditioning, once per separate() call.  replaces the time-independent part of SAMAudio.forward:
 * align_inputs (model.py:108-128), AlignModalities (align.py:30-50), EmbedAnchors (model.py:54-65),
 * memory_proj (model.py:92,172), and _repeat_for_reranking (model.py:193-203): every input is per CLIP
 * and the `candidates` sequences of a clip (candidate-minor, sequence = clip * candidates + k) share it
 * without being materialised; the ODE state has B * candidates sequences.
 *  features       [B, T, 256]      (sab_encode output)
 *  text_features  [B, L, text_dim] ; text_mask [B, L] uint8 (1 = token)
 *  video_features [B, vision_dim, T] or NULL (= zeros, model.py:188-189)
 *  anchor_ids [B, n_ids] int64 ; anchor_alignment [B, T] int64 ; audio_pad_mask [B, T] uint8 (1 = frame)
 *  flags: SAB_PREP_* — the three `None` cases of SAMAudio.forward (separate() never passes None). */
#define SAB_PREP_NO_VIDEO_TERM 1 /* masked_video_features=None: AlignModalities returns its input (align.py:41-42) */
#define SAB_PREP_NO_ANCHORS 2    /* anchor_ids=None: EmbedAnchors returns its input (model.py:57-58); ids may be NULL */
#define SAB_PREP_NO_TEXT 4       /* text_features=None: memory = time embedding only (model.py:170-172); L must be 1 */
int sab_prepare(sab_engine* e, int B, int candidates, int T, int L, const float* features, const float* text_features,
                const uint8_t* text_mask, const float* video_features, const int64_t* anchor_ids, int n_ids,
                const int64_t* anchor_alignment, const uint8_t* audio_pad_mask, int flags, void* stream);

/* One ODE function evaluation.  replaces: SAMAudio.forward / DiT.forward (model.py:130-180,
 * transformer.py:473-524) on the conditioning installed by sab_prepare.
 * noisy [Bc, T, 256], time [Bc] (device), velocity out [Bc, T, 256]. */
int sab_dit_forward(sab_engine* e, const float* noisy, const float* time, float* velocity, void* stream);

/* The ODE solve.  replaces: torchdiffeq.odeint(method=..., options={"step_size": 1/n_steps}) as called at
 * model.py:285-290 with the reference's **ode_opt: fixed-grid midpoint (the default, 2*n_steps evaluations), euler
 * (n_steps) or rk4 (torchdiffeq's 3/8 rule, 4*n_steps).  noise [Bc, T, 256] in, latent [Bc, T, 256] out (may alias). */
#define SAB_ODE_MIDPOINT 0
#define SAB_ODE_EULER 1
#define SAB_ODE_RK4 2
int sab_solve(sab_engine* e, const float* noise, int n_steps, int method, float* latent, void* stream);

/* Codec synthesis.  replaces: DACVAE.decode (codec.py:86-89) as called at model.py:291-295.
 * latent [Bc, T, 256] (target half = channels [0,128), residual half = [128,256));
 * wav [Bc, 2, T*hop] fp32 (row 0 target, row 1 residual). */
int sab_decode(sab_engine* e, const float* latent, int Bc, int T, float* wav, void* stream);

/* Introspection for benchmarks/tests: kernels launched by this engine since creation / last reset. */
int64_t sab_launch_count(sab_engine* e, int reset);
/* Workspace bytes currently held. */
int64_t sab_workspace_bytes(sab_engine* e);
/* Live per-kernel timing for roofline reporting: with profiling on, one CUDA event is recorded in front of
 * every launch on `stream`; sab_profile_report aggregates the inter-event times per kernel tag as JSON
 * ({"tag": {"launches","ms","flops","bytes"}}) and clears the log. */
int sab_profile(sab_engine* e, int enable, void* stream);
int sab_profile_report(sab_engine* e, char* json_out, int64_t capacity, void* stream);

/* ---- T5 text encoder (SURVEY §8f-3 "next" row) ----
 * replaces: the T5EncoderModel call inside T5TextEncoder.forward (reference sam_audio/model/text_encoder.py:29-35;
 * HF transformers T5Stack arithmetic).  Tokenisation stays on the host (HF tokenizer, text_encoder.py:21-27).
 * Weight names are the HF T5EncoderModel state-dict keys ("shared.weight",
 * "encoder.block.N.layer.0.SelfAttention.q.weight", ...). */
typedef struct sab_t5 sab_t5;
typedef struct sab_t5_config {
  int32_t vocab_size;   /* 32128 */
  int32_t d_model;      /* 768   */
  int32_t d_kv;         /* 64    */
  int32_t d_ff;         /* 3072  */
  int32_t n_layers;     /* 12    */
  int32_t n_heads;      /* 12    */
  int32_t n_buckets;    /* relative_attention_num_buckets = 32 */
  float   eps;          /* layer_norm_epsilon = 1e-6 */
} sab_t5_config;
int sab_t5_create(const sab_t5_config* cfg, int device, sab_t5** out);
int sab_t5_destroy(sab_t5* e);
int sab_t5_load_weight(sab_t5* e, const char* name, const float* data, const int64_t* shape, int ndim, int is_device,
                       void* stream);
int sab_t5_finalize(sab_t5* e, void* stream);
/* ids [B, L] int64, mask [B, L] uint8 (1 = token), rel_bucket [2L-1] int32 = T5's relative-position bucket of
 * (key - query) + L - 1 (computed by the host exactly as transformers does); out [B, L, d_model] fp32. */
int sab_t5_forward(sab_t5* e, const int64_t* ids, const uint8_t* mask, const int32_t* rel_bucket, int B, int L, float* out,
                   void* stream);
int64_t sab_t5_launch_count(sab_t5* e, int reset);

/* ---- unit-test seams (used by tests/ only; stable but not part of the drop-in surface) ---- */
/* C[M,N] (fp32) = A[M,K] (bf16) * B[N,K]^T (bf16) through the tcgen05 GEMM (tile BN x BK; cg = 1: one CTA per
 * 128-row tile, cg = 2: cta_group::2 pairs on 256-row tiles, cg = 0: the engine's default choice). */
int sab_test_gemm(int M, int N, int K, const void* a_bf16, const void* b_bf16, float* c, int bn, int bk, int cg,
                  void* stream);
/* O = softmax(Q K^T / sqrt(128) + mask) V through the attention kernel; all [items*T, heads*128] bf16. */
int sab_test_attention(int items, int heads, int Tq, int Tk, const void* q, const void* k, const void* v,
                       const uint8_t* key_mask, void* o, void* stream);

/* Same, through the tcgen05 self-attention kernel (Tq == Tk == T <= 256); v_lbo / v_sbo <= 0 select the
 * defaults of the MN-major V descriptor. */
int sab_test_attention_tc(int items, int heads, int T, const void* q, const void* k, const void* v,
                          const uint8_t* key_mask, void* o, int v_lbo, int v_sbo, void* stream);
/* second-generation kernel (attention_tc2.cuh): shift_log2 >= 0 = single-pass softmax with that logit bound,
 * < 0 = exact two-pass; poly = pairs of every 8 exponentials evaluated on the FMA pipe (0..4, < 0: default);
 * trace: null, or a device buffer of 8 x 32 x 8 int64 that CTA 0 fills with clock64() stamps (tools/attn_trace.py). */
int sab_test_attention_tc2(int items, int heads, int T, const void* q_bf16, const void* k_bf16, const void* v_bf16,
                           const uint8_t* key_mask, void* o_bf16, float shift_log2, int poly, long long* trace, void* stream);

/* Visual prompting, frame pre-processing.  replaces: PerceptionEncoder.get_transform (vision_encoder.py:91-113) =
 * torchvision Resize((S, S), BICUBIC, antialias) on uint8 frames, x / 255, Normalize(0.5, 0.5).
 * frames [n_frames, 3, H, W] uint8 (device), workspace [n_frames, 3, H, S] fp32 (device), out [n_frames, 3, S, S] fp32.
 * Engine-independent (no weights). */
int sab_preprocess_frames(const uint8_t* frames, int n_frames, int H, int W, int out_size, float* workspace, float* out,
                          void* stream);
/* host-only test seam: evaluation times of sab_solve's grid (method = SAB_ODE_*), in evaluation order */
int sab_test_solver_grid(int method, int n_steps, int cap, int* n_evals, float* times);
/* host-only test seam: tap window start / count / normalised weights ([out_size, cap], zero padded) of one axis */
int sab_test_aa_taps(int in_size, int out_size, int cap, int* taps, int* lo, int* cnt, float* w);

#ifdef __cplusplus
}
#endif
#endif /* SAMAUDIO_B200_H */
