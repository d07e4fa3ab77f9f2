// Frame pre-processing of the visual-prompting path (reference: sam_audio/model/vision_encoder.py:91-113 — torchvision
// Resize((S, S), BICUBIC, antialias=True) on uint8 frames, x / 255, Normalize(0.5, 0.5)):
//   uint8 [N, 3, H, W]  ->  antialiased bicubic resample (PIL-style cubic, a = -1/2, support 2 * max(scale, 1), weights
//   normalised per output pixel; width first, then height, fp32)  ->  clamp [0, 255]  ->  round half to even (the uint8
//   cast torchvision performs)  ->  (v / 255 - 0.5) / 0.5  ->  fp32 [N, 3, S, S]
// The per-output-pixel tap windows and weights are computed on the host in fp32 with the arithmetic of ATen's
// _compute_indices_min_size_weights_aa and uploaded (a few KB); the kernels are HBM-bound streaming passes.
#pragma once
#include "common.cuh"

namespace sab {

// horizontal pass: one CTA per (frame*channel, row); the uint8 row is staged in shared memory
__global__ void __launch_bounds__(256)
resize_rows_kernel(const uint8_t* __restrict__ src, int H, int W, int S, const int* __restrict__ xmin,
                   const int* __restrict__ xcnt, const float* __restrict__ wts, int taps, float* __restrict__ tmp) {
  extern __shared__ uint8_t s_row[];
  const long long plane = blockIdx.y;          // frame * 3 + channel
  const int y = blockIdx.x;
  const uint8_t* row = src + (plane * H + y) * (long long)W;
  for (int i = threadIdx.x; i < W; i += blockDim.x) s_row[i] = row[i];
  __syncthreads();
  float* out = tmp + (plane * H + y) * (long long)S;
  for (int ox = threadIdx.x; ox < S; ox += blockDim.x) {
    const int x0 = xmin[ox], n = xcnt[ox];
    const float* w = wts + (long long)ox * taps;
    float t = __fmul_rn((float)s_row[x0], w[0]);     // separate multiply and add, in tap order, as the scalar reference
    for (int j = 1; j < n; ++j) t = __fadd_rn(t, __fmul_rn((float)s_row[x0 + j], w[j]));
    out[ox] = t;
  }
}

// vertical pass + uint8 rounding + normalisation: thread = output column (coalesced reads of the intermediate rows)
__global__ void __launch_bounds__(256)
resize_cols_finish_kernel(const float* __restrict__ tmp, int H, int S, const int* __restrict__ ymin,
                          const int* __restrict__ ycnt, const float* __restrict__ wts, int taps, float* __restrict__ out) {
  const long long plane = blockIdx.y;
  const int oy = blockIdx.x;
  const int y0 = ymin[oy], n = ycnt[oy];
  const float* w = wts + (long long)oy * taps;
  const float* base = tmp + (plane * H + y0) * (long long)S;
  float* orow = out + (plane * S + oy) * (long long)S;
  for (int ox = threadIdx.x; ox < S; ox += blockDim.x) {
    float t = __fmul_rn(base[ox], w[0]);
    for (int j = 1; j < n; ++j) t = __fadd_rn(t, __fmul_rn(base[(long long)j * S + ox], w[j]));
    t = fminf(fmaxf(t, 0.f), 255.f);
    const float r = rintf(t);                         // torch.round: half to even, then the uint8 cast
    orow[ox] = __fdiv_rn(__fsub_rn(__fdiv_rn(r, 255.f), 0.5f), 0.5f);   // x.float() / 255 ; (x - 0.5) / 0.5, as torch
  }
}

}  // namespace sab
