// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   D[m, n] = sum_k A[m', k'] * B[n, k]
//
//  * A (activations, bf16, K-major rows) arrives through a 3-D TMA map (cols, rows-per-item, items);
//    the K loop is a short list of "runs" (row_shift, a_col, n k-blocks), so the same kernel serves
//    plain linears (1 run), dilated / strided Conv1d and ConvTranspose1d over channels-last audio
//    (one run per tap; TMA zero-fills rows that fall off either end of an item = the conv's padding).
//  * B (weights, bf16, [N, Ktot] K-major, packed at load time) arrives through a 2-D TMA map.
//  * Accumulators live in TMEM (2 x BN fp32 columns, double buffered): the epilogue of tile i
//    overlaps the MMAs of tile i+1.  One elected thread issues tcgen05.mma (M=128, N=BN, K=16).
//  * Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue
//    (warp w reads TMEM lanes 32*(w%4).. : thread <-> output row).
//  * Epilogues (fused, no extra pass over HBM):
//      EPI_AFFINE : v = (acc + bias[n]) * gate[row/gate_div, n] * alpha + res[row, n]
//                   -> fp32 and/or bf16 and/or Snake(v) bf16   (DiT residual/gate, ODE axpy, codec convs)
//      EPI_SWIGLU : silu(gate_cols) * up_cols -> bf16          (w1|w3 interleaved in 32-col chunks)
//      EPI_QKV    : per-head RMSNorm (q,k) + RoPE on adjacent pairs -> bf16   (heads are 128 contiguous cols)
#pragma once
#include "common.cuh"

namespace sab {

constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_MAX_RUNS = 8;

enum EpiMode { EPI_AFFINE = 0, EPI_SWIGLU = 1, EPI_QKV = 2 };

struct KRun {
  int row_shift;  // added to the tile's first row (rows outside [0, rows_per_item) read as zero)
  int a_col;      // first A column (elements)
  int nkb;        // number of BK-wide k-blocks
};

struct GemmParams {
  // tiling
  int rows_per_item;    // T: rows per batch item (tiles never straddle items)
  int n_items;
  int tiles_per_item;   // ceil(T / 128)
  int N;                // valid output columns
  int n_tiles_n;
  // K loop
  int n_runs[2];
  KRun runs[2][GEMM_MAX_RUNS];
  int n_period, n_switch;  // run list 1 iff n_period > 0 and (n0 % n_period) >= n_switch
  // affine epilogue
  const float* bias; int bias_mod;           // bias[n % bias_mod] (bias_mod == 0: bias[n])
  const float* gate; int gate_ld; int gate_div;
  float alpha;
  const float* res; long long res_ld;
  float* out_f32; long long out_f32_ld;
  __nv_bfloat16* out_bf16; long long out_bf16_ld;
  __nv_bfloat16* out_act; long long out_act_ld; const float* snake_alpha;  // snake_alpha[n % bias_mod]
  // qkv epilogue
  const float* qnorm_w; const float* knorm_w;  // [128]
  int n_q_end, n_k_end;                        // cols [0,n_q_end): q-norm, [n_q_end,n_k_end): k-norm, rest plain
  const float2* rope; int rope_T; int use_rope; float eps;
};

template <int BN, int BK>
struct GemmSmem {
  static constexpr int kABytes = GEMM_BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static constexpr int kBarBytes = 1024;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual 1 KB alignment
};

template <int BN, int BK, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmParams P) {
  using S = GemmSmem<BN, BK>;
  constexpr int kStages = S::kStages;
  constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr int kAccStride = (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;  // column offset of accumulator 1
  static_assert(2 * kAccStride <= 512, "accumulators exceed TMEM");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");
  static_assert(BK == 64 || BK == 32, "BK must be one swizzle row (128B or 64B)");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + kStages * S::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_m = P.n_items * P.tiles_per_item;
  const int n_tiles = n_tiles_m * P.n_tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = tile % n_tiles_m, nt = tile / n_tiles_m;
        const int item = mt / P.tiles_per_item;
        const int t0 = (mt % P.tiles_per_item) * GEMM_BM;
        const int n0 = nt * BN;
        const int list = (P.n_period > 0 && (n0 % P.n_period) >= P.n_switch) ? 1 : 0;
        int kb_global = 0;
        for (int r = 0; r < P.n_runs[list]; ++r) {
          const KRun run = P.runs[list][r];
          for (int kb = 0; kb < run.nkb; ++kb, ++kb_global) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * S::kStageBytes;
            uint8_t* sb = sa + S::kABytes;
            mbar_expect_tx(&full_bar[stage], S::kStageBytes);
            tma_load_3d(sa, &tmA, &full_bar[stage], run.a_col + kb * BK, t0 + run.row_shift, item);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb_global * BK, n0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int nt = tile / n_tiles_m;
        const int n0 = nt * BN;
        const int list = (P.n_period > 0 && (n0 % P.n_period) >= P.n_switch) ? 1 : 0;
        int total_kb = 0;
        for (int r = 0; r < P.n_runs[list]; ++r) total_kb += P.runs[list][r].nkb;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
          const uint64_t da = make_kmajor_desc<BK * 2>(sa);
          const uint64_t db = make_kmajor_desc<BK * 2>(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in the (addr >> 4) field
            umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int mt = tile % n_tiles_m, nt = tile / n_tiles_m;
      const int item = mt / P.tiles_per_item;
      const int t_in_item = (mt % P.tiles_per_item) * GEMM_BM + q * 32 + lane;
      const bool row_ok = t_in_item < P.rows_per_item;
      const long long row = (long long)item * P.rows_per_item + t_in_item;
      const int n0 = nt * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * kAccStride + ((uint32_t)(q * 32) << 16);

      if constexpr (MODE == EPI_AFFINE) {
        const float* gate_row = P.gate ? P.gate + (long long)(row / P.gate_div) * P.gate_ld : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          float v[32];
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          const int n = n0 + c;
          if (row_ok && n < P.N) {
            if (P.bias) {
              const int nb = P.bias_mod ? (n % P.bias_mod) : n;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = *reinterpret_cast<const float4*>(P.bias + nb + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
            if (gate_row) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 g = *reinterpret_cast<const float4*>(gate_row + n + j);
                v[j] *= g.x; v[j + 1] *= g.y; v[j + 2] *= g.z; v[j + 3] *= g.w;
              }
            }
            if (P.alpha != 1.f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= P.alpha;
            }
            if (P.res) {
              const float* rp = P.res + row * P.res_ld + n;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 r = *reinterpret_cast<const float4*>(rp + j);
                v[j] += r.x; v[j + 1] += r.y; v[j + 2] += r.z; v[j + 3] += r.w;
              }
            }
            if (P.out_f32) {
              float* op = P.out_f32 + row * P.out_f32_ld + n;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            if (P.out_bf16) {
              uint4* op = reinterpret_cast<uint4*>(P.out_bf16 + row * P.out_bf16_ld + n);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                op[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                   pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
            }
            if (P.out_act) {
              const int na = P.bias_mod ? (n % P.bias_mod) : n;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float a = P.snake_alpha[na + j];
                const float s = __sinf(a * v[j]);
                v[j] += s * s * __frcp_rn(a + 1e-9f);
              }
              uint4* op = reinterpret_cast<uint4*>(P.out_act + row * P.out_act_ld + n);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                op[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                   pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
            }
          }
        }
      } else if constexpr (MODE == EPI_SWIGLU) {
        // tile columns: [32 gate | 32 up] pairs -> BN/2 outputs at column n0/2
#pragma unroll 1
        for (int c = 0; c < BN; c += 64) {
          float g[32], u[32];
          tmem_ld32(t_addr + c, g);
          tmem_ld32(t_addr + c + 32, u);
          tmem_ld_wait();
          const int n_out = (n0 + c) >> 1;
          if (row_ok && (n0 + c) < P.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) g[j] = silu_f(g[j]) * u[j];
            uint4* op = reinterpret_cast<uint4*>(P.out_bf16 + row * P.out_bf16_ld + n_out);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              op[j] = make_uint4(pack_bf16(g[8 * j], g[8 * j + 1]), pack_bf16(g[8 * j + 2], g[8 * j + 3]),
                                 pack_bf16(g[8 * j + 4], g[8 * j + 5]), pack_bf16(g[8 * j + 6], g[8 * j + 7]));
          }
        }
      } else {  // EPI_QKV: BN is a multiple of 128; each 128-col group is one head
        const int pos = (int)(row % P.rope_T);
#pragma unroll 1
        for (int hc = 0; hc < BN; hc += 128) {
          const int n = n0 + hc;
          const float* nw = (n < P.n_q_end) ? P.qnorm_w : (n < P.n_k_end ? P.knorm_w : nullptr);
          float rstd = 1.f;
          if (nw) {
            float ss = 0.f;
#pragma unroll 1
            for (int c = 0; c < 128; c += 32) {
              float v[32];
              tmem_ld32(t_addr + hc + c, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) ss = fmaf(v[j], v[j], ss);
            }
            rstd = rsqrtf(ss * (1.f / 128.f) + P.eps);
          }
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            float v[32];
            tmem_ld32(t_addr + hc + c, v);
            tmem_ld_wait();
            if (row_ok && n < P.N) {
              if (nw) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = v[j] * rstd * nw[c + j];
                if (P.use_rope) {
                  const float2* rp = P.rope + (long long)pos * 64 + (c >> 1);
#pragma unroll
                  for (int j = 0; j < 32; j += 2) {
                    const float2 cs = rp[j >> 1];
                    const float x0 = v[j], x1 = v[j + 1];
                    v[j] = x0 * cs.x - x1 * cs.y;
                    v[j + 1] = x0 * cs.y + x1 * cs.x;
                  }
                }
              }
              uint4* op = reinterpret_cast<uint4*>(P.out_bf16 + row * P.out_bf16_ld + n + c);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                op[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                   pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
            }
          }
        }
      }
      // release the accumulator back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace sab
