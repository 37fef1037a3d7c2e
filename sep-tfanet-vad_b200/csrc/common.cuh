// Shared device helpers for the septfa kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace septfa {

constexpr int kNfft = 512;
constexpr int kHop = 256;
constexpr int kBins = 257;       // n_fft/2 + 1
constexpr int kC = 256;          // BN_dim: residual-stream channels (= kBins - 1, DC dropped)
constexpr int kH = 512;          // H_dim: depthwise hidden channels
constexpr int kLogitStride = 576;  // row pitch of the frame-major logits buffer (514 padded to 3 x 192)
constexpr int kMaxSegs = 72;     // max utterance segments touched by one 128-row tile (T >= 2)

// Per-utterance GroupNorm statistics accumulators: {sum, sum of squares} in double.
struct Stat2 { double s, ss; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float prelu(float x, float a) { return x >= 0.f ? x : a * x; }
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// mean / rstd of one utterance from its double accumulators (biased variance, GroupNorm(1,C)).
__device__ __forceinline__ float2 stat_mean_rstd(const Stat2* st, double inv_n, float eps) {
  double m = st->s * inv_n;
  double var = st->ss * inv_n - m * m;
  if (var < 0.0) var = 0.0;
  return make_float2((float)m, (float)(1.0 / sqrt(var + (double)eps)));
}

// Block-wide reduction of (s, ss) for blocks whose rows all belong to ONE utterance, followed by
// one double atomicAdd pair. `red` is >= 64 floats of shared memory. Must be called by all threads.
__device__ __forceinline__ void block_stat_atomic(float s, float ss, Stat2* dst, float* red) {
  s = warp_sum(s);
  ss = warp_sum(ss);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (l == 0) { red[w] = s; red[32 + w] = ss; }
  __syncthreads();
  if (w == 0) {
    double a = l < nw ? (double)red[l] : 0.0;
    double b = l < nw ? (double)red[32 + l] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
    if (l == 0) { atomicAdd(&dst->s, a); atomicAdd(&dst->ss, b); }
  }
  __syncthreads();
}

}  // namespace septfa
