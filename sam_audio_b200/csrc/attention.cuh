// Masked softmax attention for head_dim = 128 (self- and cross-attention of the DiT block;
// reference: sam_audio/model/transformer.py:153-160 — SDPA, scale 1/sqrt(hd), bool key mask, True = attend).
//
// v0 kernel: flash-style, one CTA = 64 query rows of one (item, head), 4 warps x 16 rows,
// K/V streamed in 64-key tiles through a 2-stage cp.async pipeline, bf16 mma.sync m16n8k16 with fp32
// accumulation and an online (running max / sum) softmax in the exp2 domain.
// Q/K/V are read in place from the fused-QKV GEMM output ([rows, ld] with the head's 128 channels
// contiguous); O is written head-major "(h d)" so that `wo` consumes it directly.
#pragma once
#include "common.cuh"

namespace sab {

struct AttnParams {
  const __nv_bfloat16* q; long long q_ld; int q_col0;   // element (item*Tq + t, q_col0 + h*128 + d)
  const __nv_bfloat16* k; long long k_ld; int k_col0;   // element (item*Tk + s, k_col0 + h*128 + d)
  const __nv_bfloat16* v; long long v_ld; int v_col0;
  __nv_bfloat16* o; long long o_ld;                     // element (item*Tq + t, h*128 + d)
  const uint8_t* key_mask;                              // [items, Tk] (1 = attend) or null
  int Tq, Tk;
  float scale_log2;                                     // (1/sqrt(128)) * log2(e)
  int kv_div, mask_div;                                 // K/V item = query item / kv_div, mask row = query item / mask_div
                                                        // (the candidates of one clip share the clip's text K/V and its
                                                        // masks); 0 or 1: same item
};
SAB_DEVICE int attn_kv_item(const AttnParams& P, int item) { return P.kv_div > 1 ? item / P.kv_div : item; }
SAB_DEVICE int attn_mask_item(const AttnParams& P, int item) { return P.mask_div > 1 ? item / P.mask_div : item; }

constexpr int ATT_BQ = 64, ATT_BK = 64, ATT_D = 128, ATT_THREADS = 128;
constexpr int ATT_TILE_BYTES = 64 * ATT_D * 2;  // 16 KB
constexpr int ATT_SMEM = ATT_TILE_BYTES * 5;    // Q + 2 x (K, V)

SAB_DEVICE void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
SAB_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SAB_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

SAB_DEVICE void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
SAB_DEVICE void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
SAB_DEVICE void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tile: 64 rows x 256 B; 16-byte chunk index XOR (row & 7) -> conflict-free ldmatrix
SAB_DEVICE uint32_t tile_off(int row, int chunk) { return (uint32_t)(row * 256 + ((chunk ^ (row & 7)) << 4)); }

SAB_DEVICE void load_tile(uint32_t smem_tile, const __nv_bfloat16* base, long long ld, int row0, int n_rows_valid,
                          int tid) {
  // 64 rows x 16 chunks = 1024 chunks / 128 threads
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = tid + i * ATT_THREADS;
    const int r = idx >> 4, c = idx & 15;
    const bool ok = (row0 + r) < n_rows_valid;
    const __nv_bfloat16* src = base + (long long)(ok ? (row0 + r) : 0) * ld + c * 8;
    cp_async16(smem_tile + tile_off(r, c), src, ok);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const AttnParams P) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int qt = blockIdx.x, head = blockIdx.y, item = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const uint32_t sQ = smem_u32(att_smem);
  const uint32_t sK0 = sQ + ATT_TILE_BYTES;           // stage s: K at sK0 + s*2*TILE, V right after
  const int q0 = qt * ATT_BQ;

  const __nv_bfloat16* qb = P.q + (long long)item * P.Tq * P.q_ld + P.q_col0 + head * ATT_D;
  const int kv_item = attn_kv_item(P, item);
  const __nv_bfloat16* kb = P.k + (long long)kv_item * P.Tk * P.k_ld + P.k_col0 + head * ATT_D;
  const __nv_bfloat16* vb = P.v + (long long)kv_item * P.Tk * P.v_ld + P.v_col0 + head * ATT_D;
  const uint8_t* mask = P.key_mask ? P.key_mask + (long long)attn_mask_item(P, item) * P.Tk : nullptr;
  const int n_kt = (P.Tk + ATT_BK - 1) / ATT_BK;

  load_tile(sQ, qb, P.q_ld, q0, P.Tq, tid);
  load_tile(sK0, kb, P.k_ld, 0, P.Tk, tid);
  load_tile(sK0 + ATT_TILE_BYTES, vb, P.v_ld, 0, P.Tk, tid);
  cp_async_commit();

  uint32_t qf[8][4];
  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int kt = 0; kt < n_kt; ++kt) {
    const int st = kt & 1;
    if (kt + 1 < n_kt) {
      const uint32_t nk = sK0 + (st ^ 1) * 2 * ATT_TILE_BYTES;
      load_tile(nk, kb, P.k_ld, (kt + 1) * ATT_BK, P.Tk, tid);
      load_tile(nk + ATT_TILE_BYTES, vb, P.v_ld, (kt + 1) * ATT_BK, P.Tk, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
      // Q fragments: warp's 16 rows x 128 -> 8 k-steps
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int r = warp * 16 + (lane & 15);
        const int c = ks * 2 + (lane >> 4);
        ldsm_x4(sQ + tile_off(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint32_t sK = sK0 + st * 2 * ATT_TILE_BYTES, sV = sK + ATT_TILE_BYTES;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {        // 8 keys per n-block
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {   // two k-steps (32 of the 128 dims) per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        const int r = j * 8 + (lane & 7);
        const int c = kp * 4 + (lane >> 3);
        ldsm_x4(sK + tile_off(r, c), b0, b1, b2, b3);
        mma_bf16_16816(s[j], qf[kp * 2], b0, b1);
        mma_bf16_16816(s[j], qf[kp * 2 + 1], b2, b3);
      }
    }
    // ---- mask + online softmax ----
    const int key0 = kt * ATT_BK;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = key0 + j * 8 + t4 * 2 + e;
        const bool ok = key < P.Tk && (!mask || mask[key]);
        if (!ok) { s[j][e] = -INFINITY; s[j][2 + e] = -INFINITY; }
      }
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      mnew[h] = fmaxf(m_run[h], mx[h]);
      const float base = (mnew[h] == -INFINITY) ? 0.f : mnew[h];
      corr[h] = exp2f((m_run[h] - base) * P.scale_log2);   // m_run = -inf -> 0
      m_run[h] = mnew[h];
      mnew[h] = base * P.scale_log2;
      l_run[h] *= corr[h];
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] * P.scale_log2 - mnew[0]);
      const float p1 = exp2f(s[j][1] * P.scale_log2 - mnew[0]);
      const float p2 = exp2f(s[j][2] * P.scale_log2 - mnew[1]);
      const float p3 = exp2f(s[j][3] * P.scale_log2 - mnew[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      pf[j >> 1][(j & 1) * 2] = pack_bf16(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {       // 16 keys per k-step
#pragma unroll
      for (int np = 0; np < 8; ++np) {     // 16 output dims per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        const int r = kk * 16 + (lane & 15);
        const int c = np * 2 + (lane >> 4);
        ldsm_x4_t(sV + tile_off(r, c), b0, b1, b2, b3);
        mma_bf16_16816(o[np * 2], pf[kk], b0, b1);
        mma_bf16_16816(o[np * 2 + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();  // all warps done with this stage before it is refilled
  }

  // ---- normalise and store ----
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  __nv_bfloat16* ob = P.o + (long long)item * P.Tq * P.o_ld + head * ATT_D;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int col = i * 8 + t4 * 2;
    if (r0 < P.Tq) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * P.o_ld + col) = pack_bf16(o[i][0] * inv0, o[i][1] * inv0);
    if (r1 < P.Tq) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * P.o_ld + col) = pack_bf16(o[i][2] * inv1, o[i][3] * inv1);
  }
}

// ---------------------------------------------------------------------------------------------
// Cross-attention to a handful of text tokens (Tk <= 16: "man speaking" is 3 tokens).  HBM-bound:
// 8 lanes per (query row, head), 16 head dims per lane (32 B loads, 256 B contiguous per row); K/V of the
// (item, head) live in shared memory; scores by a 3-step shuffle reduction, softmax in registers.
// Reads Q once, writes O once.
// ---------------------------------------------------------------------------------------------
constexpr int XATT_MAX_TK = 16, XATT_ROWS = 128, XATT_THREADS = 256;
SAB_DEVICE void bf16x8_to_f32(const uint4& r, float* f) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
  const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.z));
  const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.w));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// TK = compile-time key count bucket (keys >= P.Tk are zero rows with a -inf bias), so every loop is straight-line
template <int TK>
__global__ void __launch_bounds__(XATT_THREADS)
xattn_small_kernel(const AttnParams P) {
  __shared__ __align__(16) float sK[TK][ATT_D];
  __shared__ __align__(16) float sV[TK][ATT_D];
  __shared__ float sBias[TK];
  const int qt = blockIdx.x, head = blockIdx.y, item = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kv_item = attn_kv_item(P, item);
  const __nv_bfloat16* kb = P.k + (long long)kv_item * P.Tk * P.k_ld + P.k_col0 + head * ATT_D;
  const __nv_bfloat16* vb = P.v + (long long)kv_item * P.Tk * P.v_ld + P.v_col0 + head * ATT_D;
  for (int i = tid; i < TK * (ATT_D / 8); i += XATT_THREADS) {   // 8 dims per thread, converted to fp32 once
    const int j = i / (ATT_D / 8), c = (i % (ATT_D / 8)) * 8;
    uint4 kr = make_uint4(0u, 0u, 0u, 0u), vr = kr;
    if (j < P.Tk) {
      kr = *reinterpret_cast<const uint4*>(kb + (long long)j * P.k_ld + c);
      vr = *reinterpret_cast<const uint4*>(vb + (long long)j * P.v_ld + c);
    }
    // shared layout: lane group l8 = dim/16 owns dims [16 l8, +16) as four float4 "i"; float4 (i, l8) sits at
    // index i*8 + l8 so that the 8 lanes of a row read consecutive 16 B words (no bank conflicts)
    float kf[8], vf[8];
    bf16x8_to_f32(kr, kf);
    bf16x8_to_f32(vr, vf);
    const int l8 = c >> 4, i0 = (c & 15) >> 2;
    *reinterpret_cast<float4*>(&sK[j][((i0) * 8 + l8) * 4]) = make_float4(kf[0], kf[1], kf[2], kf[3]);
    *reinterpret_cast<float4*>(&sK[j][((i0 + 1) * 8 + l8) * 4]) = make_float4(kf[4], kf[5], kf[6], kf[7]);
    *reinterpret_cast<float4*>(&sV[j][((i0) * 8 + l8) * 4]) = make_float4(vf[0], vf[1], vf[2], vf[3]);
    *reinterpret_cast<float4*>(&sV[j][((i0 + 1) * 8 + l8) * 4]) = make_float4(vf[4], vf[5], vf[6], vf[7]);
  }
  if (tid < TK)
    sBias[tid] = (tid < P.Tk && (!P.key_mask || P.key_mask[(long long)attn_mask_item(P, item) * P.Tk + tid])) ? 0.f : -INFINITY;
  // this lane: row (pass*32 + warp*4 + lane/8), dims [16*(lane%8), +16)
  const int sub = lane >> 3, d0 = (lane & 7) * 16;
  const __nv_bfloat16* qb = P.q + (long long)item * P.Tq * P.q_ld + P.q_col0 + head * ATT_D + d0;
  __nv_bfloat16* ob = P.o + (long long)item * P.Tq * P.o_ld + head * ATT_D + d0;
  constexpr int kPasses = XATT_ROWS / 32;
  uint4 qraw[kPasses][2];
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {       // all query loads in flight before the K/V barrier
    const int r = qt * XATT_ROWS + p * 32 + warp * 4 + sub;
    qraw[p][0] = qraw[p][1] = make_uint4(0u, 0u, 0u, 0u);
    if (r < P.Tq) {
      const uint4* src = reinterpret_cast<const uint4*>(qb + (long long)r * P.q_ld);
      qraw[p][0] = src[0];
      qraw[p][1] = src[1];
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int p = 0; p < kPasses; ++p) {
    const int r = qt * XATT_ROWS + p * 32 + warp * 4 + sub;
    float qf[16], o[16];
    bf16x8_to_f32(qraw[p][0], qf);
    bf16x8_to_f32(qraw[p][1], qf + 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = 0.f;
    float m_run = -INFINITY, den = 0.f;     // online softmax over the (few) keys keeps the live state small
#pragma unroll 4
    for (int j = 0; j < TK; ++j) {
      const float4* kr = reinterpret_cast<const float4*>(&sK[j][0]) + (lane & 7);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 k4 = kr[i * 8];
        d = fmaf(qf[4 * i], k4.x, d); d = fmaf(qf[4 * i + 1], k4.y, d);
        d = fmaf(qf[4 * i + 2], k4.z, d); d = fmaf(qf[4 * i + 3], k4.w, d);
      }
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      const float s = fmaf(d, P.scale_log2, sBias[j]);
      const float m_new = fmaxf(m_run, s);
      const float base = (m_new == -INFINITY) ? 0.f : m_new;
      float corr, pj;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(corr) : "f"(m_run - base));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pj) : "f"(s - base));
      m_run = m_new;
      den = fmaf(den, corr, pj);
      const float4* vr = reinterpret_cast<const float4*>(&sV[j][0]) + (lane & 7);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v4 = vr[i * 8];
        o[4 * i] = fmaf(o[4 * i], corr, pj * v4.x); o[4 * i + 1] = fmaf(o[4 * i + 1], corr, pj * v4.y);
        o[4 * i + 2] = fmaf(o[4 * i + 2], corr, pj * v4.z); o[4 * i + 3] = fmaf(o[4 * i + 3], corr, pj * v4.w);
      }
    }
    const float inv = 1.f / den;
    if (r < P.Tq) {
      uint4* dst = reinterpret_cast<uint4*>(ob + (long long)r * P.o_ld);
      dst[0] = make_uint4(pack_bf16(o[0] * inv, o[1] * inv), pack_bf16(o[2] * inv, o[3] * inv),
                          pack_bf16(o[4] * inv, o[5] * inv), pack_bf16(o[6] * inv, o[7] * inv));
      dst[1] = make_uint4(pack_bf16(o[8] * inv, o[9] * inv), pack_bf16(o[10] * inv, o[11] * inv),
                          pack_bf16(o[12] * inv, o[13] * inv), pack_bf16(o[14] * inv, o[15] * inv));
    }
  }
}
inline void launch_xattn_small(const AttnParams& ap, int items, int heads, cudaStream_t st) {
  const dim3 grid((ap.Tq + XATT_ROWS - 1) / XATT_ROWS, heads, items);
  if (ap.Tk <= 4) xattn_small_kernel<4><<<grid, XATT_THREADS, 0, st>>>(ap);
  else if (ap.Tk <= 8) xattn_small_kernel<8><<<grid, XATT_THREADS, 0, st>>>(ap);
  else xattn_small_kernel<16><<<grid, XATT_THREADS, 0, st>>>(ap);
}

}  // namespace sab
