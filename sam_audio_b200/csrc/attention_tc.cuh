// tcgen05 self-attention for sequences of up to 256 keys (10 s clips: T = 250), head_dim = 128.
// reference: sam_audio/model/transformer.py:153-160 (SDPA, scale 1/sqrt(hd), bool key mask, True = attend).
//
// Persistent CTAs loop over (item, head) work items; the next item's Q/K (then V) tiles are prefetched as soon as
// the tensor core has retired the MMAs that read them.  Per work item both 128-query tiles share one K/V load and run as two independent softmax
// groups (warps 1-4 and 5-8, two warps per SM sub-partition) against one TMA/MMA issuing thread, so the second
// tile's QK^T runs under the first tile's softmax and the first tile's PV under the second tile's softmax.
// Everything between the two HBM touches stays on chip:
//   TMA  : Q [128 x 128], K [256 x 128], V [256 x 128] tiles (SWIZZLE_128B, zero-filled past the sequence end)
//   UMMA : S = Q K^T            (M=128, N=256, K=128: 8 x tcgen05.mma, A/B from shared memory) -> TMEM [0,256)
//   CUDA : row softmax, thread == query row == TMEM lane: max / exp2 / sum straight out of TMEM,
//          P written back to TMEM as packed bf16 IN PLACE over S (cols [0,128))
//   UMMA : O = P V              (M=128, N=128, K=256: 16 x tcgen05.mma, A = P from TMEM, B = V MN-major smem)
//          -> TMEM [128,256) (dead half of S)
//   CUDA : O / rowsum -> bf16 -> global, head-major "(h d)"
// Longer sequences use the streaming kernel in attention.cuh.
#pragma once
#include "common.cuh"
#include "attention.cuh"

namespace sab {

constexpr int ATC_THREADS = 288;                  // warp 0: TMA + MMA issue + TMEM alloc; warps 1..8: two softmax groups
constexpr int ATC_Q_BYTES = 128 * 128 * 2;        // 32 KB per 128-query tile (2 sub-tiles of [128 x 64])
constexpr int ATC_KV_BYTES = 256 * 128 * 2;       // 64 KB  (2 sub-tiles of [256 x 64])
constexpr int ATC_SMEM = 2 * ATC_Q_BYTES + 2 * ATC_KV_BYTES + 1024 /*barriers*/ + 1024 /*alignment*/;

struct AttnTcParams {
  __nv_bfloat16* o; long long o_ld;
  const uint8_t* key_mask;          // [items, Tk] or null
  int Tq, Tk, heads, items;
  int mask_div;                     // mask row = item / mask_div (candidates share their clip's pad mask); 0/1: item
  int q_col0, k_col0, v_col0;       // column of head 0 inside the fused QKV row
  float scale_log2;
  // debug knobs for bring-up of the MN-major V descriptor (bytes)
  int v_lbo, v_sbo;
};

SAB_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SAB_DEVICE void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
SAB_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]
SAB_DEVICE void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major SWIZZLE_128B operand descriptor: 64-element (128 B) chunks along MN are `lbo` bytes apart,
// 8-row groups along K are `sbo` bytes apart (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).
SAB_DEVICE uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_bmn(int M, int N) {   // as make_idesc_bf16, B operand MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q /*box 64 x 128*/,
                    const __grid_constant__ CUtensorMap tm_k /*box 64 x 256*/,
                    const __grid_constant__ CUtensorMap tm_v /*box 64 x 256*/, const __grid_constant__ AttnTcParams P) {
  extern __shared__ uint8_t atc_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // two query tiles
  uint8_t* sK = sQ + 2 * ATC_Q_BYTES;
  uint8_t* sV = sK + ATC_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATC_KV_BYTES);
  uint64_t *bar_q0k = bars, *bar_q1 = bars + 1, *bar_v = bars + 2;
  uint64_t *bar_s = bars + 3 /*[2]*/, *bar_p = bars + 5 /*[2]*/, *bar_o = bars + 7 /*[2]*/, *bar_e = bars + 9 /*[2]*/;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mt = (P.Tq + 127) / 128;             // 1 or 2 query tiles
  const int n_work = P.heads * P.items;            // persistent CTAs loop over (item, head)

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
      mbar_init(bar_q0k, 1);
      mbar_init(bar_q1, 1);
      mbar_init(bar_v, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_s + i, 1);
        mbar_init(bar_p + i, 128);
        mbar_init(bar_o + i, 1);
        mbar_init(bar_e + i, 128);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // (Q0, K) first so the first QK^T can start, then Q1; V separately (it is only needed for P V)
      auto load_qk = [&](int w) {
        const int head = w % P.heads, item = w / P.heads;
        const int qc = P.q_col0 + head * 128, kc = P.k_col0 + head * 128;
        mbar_expect_tx(bar_q0k, ATC_Q_BYTES + ATC_KV_BYTES);
        tma_load_3d(sQ, &tm_q, bar_q0k, qc, 0, item);
        tma_load_3d(sQ + ATC_Q_BYTES / 2, &tm_q, bar_q0k, qc + 64, 0, item);
        tma_load_3d(sK, &tm_k, bar_q0k, kc, 0, item);
        tma_load_3d(sK + ATC_KV_BYTES / 2, &tm_k, bar_q0k, kc + 64, 0, item);
        if (n_mt > 1) {
          mbar_expect_tx(bar_q1, ATC_Q_BYTES);
          tma_load_3d(sQ + ATC_Q_BYTES, &tm_q, bar_q1, qc, 128, item);
          tma_load_3d(sQ + ATC_Q_BYTES + ATC_Q_BYTES / 2, &tm_q, bar_q1, qc + 64, 128, item);
        }
      };
      auto load_v = [&](int w) {
        const int head = w % P.heads, item = w / P.heads;
        const int vc = P.v_col0 + head * 128;
        mbar_expect_tx(bar_v, ATC_KV_BYTES);
        tma_load_3d(sV, &tm_v, bar_v, vc, 0, item);
        tma_load_3d(sV + ATC_KV_BYTES / 2, &tm_v, bar_v, vc + 64, 0, item);
      };
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 256);
      constexpr uint32_t idesc_o = make_idesc_bf16_bmn(128, 128);
      if ((int)blockIdx.x < n_work) { load_qk(blockIdx.x); load_v(blockIdx.x); }
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int w_next = w + gridDim.x;
        // ---- S_m = Q_m K^T  -> TMEM cols [256m, 256m + 256); the buffer is free once the previous item's
        //      epilogue of tile m has read its O (which aliases the upper half of S_m) ----
        for (int m = 0; m < n_mt; ++m) {
          if (it > 0) mbar_wait(bar_e + m, ph ^ 1);
          mbar_wait(m == 0 ? bar_q0k : bar_q1, ph);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t qa = smem_u32(sQ) + m * ATC_Q_BYTES + (ks >> 2) * (ATC_Q_BYTES / 2);
            const uint32_t ka = smem_u32(sK) + (ks >> 2) * (ATC_KV_BYTES / 2);
            umma_f16(tmem + m * 256, make_kmajor_desc<128>(qa) + (uint64_t)((ks & 3) * 2),
                     make_kmajor_desc<128>(ka) + (uint64_t)((ks & 3) * 2), idesc_s, ks ? 1u : 0u);
          }
          umma_commit(bar_s + m);
        }
        // Q and K are dead once the last QK^T has retired: prefetch the next item's under this item's softmax
        mbar_wait(bar_s + (n_mt - 1), ph);
        if (w_next < n_work) load_qk(w_next);
        // ---- O_m = P_m V  (A = P from TMEM cols [256m, +128), D = cols [256m + 128, +128)) ----
        mbar_wait(bar_v, ph);
        for (int m = 0; m < n_mt; ++m) {
          mbar_wait(bar_p + m, ph);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {      // 16 keys per instruction
            const uint64_t vb = make_mnmajor_desc(smem_u32(sV) + ks * 16 * 128, (uint32_t)P.v_lbo, (uint32_t)P.v_sbo);
            umma_f16_ts(tmem + m * 256 + 128, tmem + m * 256 + ks * 8, vb, idesc_o, ks ? 1u : 0u);
          }
          umma_commit(bar_o + m);
        }
        mbar_wait(bar_o + (n_mt - 1), ph);     // V is dead: prefetch the next item's
        if (w_next < n_work) load_v(w_next);
      }
    }
  } else if (((warp - 1) >> 2) < n_mt) {
    // ===================== softmax / epilogue group m: thread == query row == TMEM lane =====================
    const int m = (warp - 1) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int t = m * 128 + row;
    const uint32_t tS = tmem + m * 256 + ((uint32_t)(q * 32) << 16);
    const uint32_t tO = tS + 128;
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int head = w % P.heads, item = w / P.heads;
      // key validity bits of this item (sequence end + padding mask), one ballot per 32 keys
      uint32_t kbits[8];
      {
        const uint8_t* mk = P.key_mask ? P.key_mask + (long long)(P.mask_div > 1 ? item / P.mask_div : item) * P.Tk : nullptr;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int key = c * 32 + lane;
          kbits[c] = __ballot_sync(0xffffffffu, key < P.Tk && (!mk || mk[key]));
        }
      }
      mbar_wait(bar_s + m, ph);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v[32];
        tmem_ld32(tS + c * 32, v);
        tmem_ld_wait();
        const uint32_t bits = kbits[c];
        if (bits == 0xffffffffu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, ((bits >> j) & 1u) ? v[j] : -INFINITY);
        }
      }
      const float mscaled = (mx == -INFINITY) ? 0.f : mx * P.scale_log2;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v[32];
        tmem_ld32(tS + c * 32, v);
        tmem_ld_wait();
        const uint32_t bits = kbits[c];
        uint32_t pk[16];
        if (bits == 0xffffffffu) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2_approx(fmaf(v[j], P.scale_log2, -mscaled));
            const float p1 = ex2_approx(fmaf(v[j + 1], P.scale_log2, -mscaled));
            sum += p0 + p1;
            pk[j >> 1] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = ((bits >> j) & 1u) ? ex2_approx(fmaf(v[j], P.scale_log2, -mscaled)) : 0.f;
            const float p1 = ((bits >> (j + 1)) & 1u) ? ex2_approx(fmaf(v[j + 1], P.scale_log2, -mscaled)) : 0.f;
            sum += p0 + p1;
            pk[j >> 1] = pack_bf16(p0, p1);
          }
        }
        tmem_st16(tS + c * 16, pk);   // in place: cols [16c, 16c+16) were consumed by chunk c/2 <= c
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p + m);
      // ---- O / sum -> global ----
      mbar_wait(bar_o + m, ph);
      tc_fence_after();
      const float inv = 1.f / sum;
      __nv_bfloat16* op = P.o + ((long long)item * P.Tq + t) * P.o_ld + head * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld32(tO + c * 32, v);
        tmem_ld_wait();
        if (t < P.Tq) {
          uint4* dst = reinterpret_cast<uint4*>(op + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_uint4(pack_bf16(v[8 * j] * inv, v[8 * j + 1] * inv), pack_bf16(v[8 * j + 2] * inv, v[8 * j + 3] * inv),
                                pack_bf16(v[8 * j + 4] * inv, v[8 * j + 5] * inv), pack_bf16(v[8 * j + 6] * inv, v[8 * j + 7] * inv));
        }
      }
      tc_fence_before();
      mbar_arrive(bar_e + m);   // S_m / O_m may be overwritten by the next item's QK^T
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace sab
