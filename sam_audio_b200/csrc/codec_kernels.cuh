// SIMT kernels at the two 1-channel ends of the DAC-VAE codec, plus layout helpers.
// The multi-channel convolutions run on the tcgen05 GEMM (gemm_tc.cuh) over channels-last audio.
// (DAC-VAE arithmetic is restated from the Descript-DAC layout; reference call sites codec.py:65-89.)
#pragma once
#include "common.cuh"

namespace sab {

SAB_DEVICE float snake_f(float v, float a) {
  const float s = __sinf(a * v);
  return v + s * s * __frcp_rn(a + 1e-9f);
}

// encoder.block.0: Conv1d(1 -> C0, k=7, pad=3) on mono PCM, output channels-last.
//   x_out[b,t,c] = bias[c] + sum_k w[c,k] * wav[b,t+k-3]        (fp32 residual stream)
//   a_out[b,t,c] = bf16(Snake_alpha(x_out))                      (operand of the first residual unit)
// 16 threads per sample (4 channels each, C0 = 64): 256 B / 128 B coalesced stores per sample.
__global__ void __launch_bounds__(256)
enc_conv0_kernel(const float* __restrict__ wav, long long S, int C0, const float* __restrict__ w /*[C0,7]*/,
                 const float* __restrict__ bias, const float* __restrict__ alpha, float* __restrict__ x_out,
                 __nv_bfloat16* __restrict__ a_out) {
  const int tpc = C0 / 4;  // threads per sample
  const int b = blockIdx.y;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long t = idx / tpc;
  const int c = (int)(idx % tpc) * 4;
  if (t >= S) return;
  const float* wp = wav + b * S;
  float xin[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const long long tt = t + k - 3;
    xin[k] = (tt >= 0 && tt < S) ? wp[tt] : 0.f;
  }
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float acc = bias[c + j];
#pragma unroll
    for (int k = 0; k < 7; ++k) acc = fmaf(w[(c + j) * 7 + k], xin[k], acc);
    v[j] = acc;
  }
  const long long o = (b * S + t) * C0 + c;
  *reinterpret_cast<float4*>(x_out + o) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<uint2*>(a_out + o) =
      make_uint2(pack_bf16(snake_f(v[0], alpha[c]), snake_f(v[1], alpha[c + 1])),
                 pack_bf16(snake_f(v[2], alpha[c + 2]), snake_f(v[3], alpha[c + 3])));
}

// decoder last layer: Conv1d(C -> 1, k=7, pad=3) + tanh on the Snake-activated channels-last input.
//   wav[i,t] = tanh(bias + sum_{k,c} w[c,k] * a[i,t+k-3,c])
// CTA = 128 samples (+6 halo) staged in shared memory (row padded by 2 bf16 -> conflict-free column reads).
constexpr int DEC_LAST_TB = 128;
__global__ void __launch_bounds__(DEC_LAST_TB)
dec_last_kernel(const __nv_bfloat16* __restrict__ a, long long S, int C, const float* __restrict__ w /*[C,7]*/,
                const float* __restrict__ bias, float* __restrict__ wav) {
  extern __shared__ __align__(16) uint8_t dl_smem[];
  const int ldc = C + 2;
  __nv_bfloat16* sa = reinterpret_cast<__nv_bfloat16*>(dl_smem);           // [(TB+6), ldc]
  float* sw = reinterpret_cast<float*>(sa + (DEC_LAST_TB + 6) * ldc + 2);  // [7, C] (k-major)
  const int item = blockIdx.y;
  const long long t0 = blockIdx.x * (long long)DEC_LAST_TB;
  for (int i = threadIdx.x; i < C * 7; i += DEC_LAST_TB) {
    const int c = i / 7, k = i % 7;
    sw[k * C + c] = w[i];
  }
  const int c2 = C / 2;
  for (int i = threadIdx.x; i < (DEC_LAST_TB + 6) * c2; i += DEC_LAST_TB) {
    const int r = i / c2, cc = (i % c2) * 2;
    const long long t = t0 + r - 3;
    uint32_t v = 0;
    if (t >= 0 && t < S) v = *reinterpret_cast<const uint32_t*>(a + ((long long)item * S + t) * C + cc);
    *reinterpret_cast<uint32_t*>(sa + r * ldc + cc) = v;
  }
  __syncthreads();
  const long long t = t0 + threadIdx.x;
  if (t >= S) return;
  float acc = bias[0];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const __nv_bfloat16* row = sa + (threadIdx.x + k) * ldc;
    const float* wk = sw + k * C;
    for (int c = 0; c < C; c += 2) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + c));
      acc = fmaf(f.x, wk[c], acc);
      acc = fmaf(f.y, wk[c + 1], acc);
    }
  }
  wav[(long long)item * S + t] = tanhf(acc);
}

// latent [Bc, T, 2*Cz] fp32 -> bf16 [2*Bc, T, Cz]  (model.py:291-295: row 2b = target half, 2b+1 = residual half)
__global__ void latent_split_kernel(const float* __restrict__ lat, int T, int Cz, long long n_total,
                                    __nv_bfloat16* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cz);
    const long long r = i / Cz;
    const int t = (int)(r % T);
    const long long item = r / T;  // 2b + half
    const long long b = item >> 1;
    const int half = (int)(item & 1);
    out[i] = __float2bfloat16(lat[(b * T + t) * (2 * Cz) + half * Cz + c]);
  }
}

// [B, C, T] fp32 -> [B, T, C] bf16 (video features arrive channels-first: model.py:191)
__global__ void transpose_cast_kernel(const float* __restrict__ x, int C, int T, long long n_total,
                                      __nv_bfloat16* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int t = (int)(r % T);
    const long long b = r / T;
    y[i] = __float2bfloat16(x[(b * C + c) * T + t]);
  }
}

}  // namespace sab
