// Pre-processing on the device: the min-max normalisation of only_inference.py:81,
//   normalized = 1.8 * (audio - audio.min()) / (audio.max() - audio.min()) - 0.9,
// batched over utterances (optionally ragged: `lengths`), in the reference's float32 operation order - subtract,
// multiply by float32(1.8), IEEE divide, subtract float32(0.9) - with explicit round-to-nearest intrinsics so that the
// compiler cannot contract them into FMAs: the result is bit-identical to numpy's. A constant signal gives 0/0 = NaN,
// like the reference. Two kernels: per-utterance extrema (ordered-integer atomics), then the element-wise map with
// 16-byte accesses.
#include "kernels.h"

namespace septfa {

namespace {

constexpr int kNormChunk = 8192;   // samples per CTA

// monotone float <-> unsigned map, so that unsigned atomicMin / atomicMax order floats (including negatives)
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// ordered(+inf) for the running minimum, ordered(-inf) for the running maximum
__global__ void k_ext_init(unsigned* ext, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ext[i] = (i & 1) ? f2ord(-INFINITY) : f2ord(INFINITY);
}

__global__ void __launch_bounds__(256) k_minmax(const float* __restrict__ x, int64_t L, const int64_t* __restrict__ lengths,
                                                unsigned* __restrict__ ext /*[B][2]: ordered min, ordered max*/) {
  __shared__ float red[2][8];
  const int b = blockIdx.y;
  const int64_t n = lengths != nullptr ? min(lengths[b], L) : L;
  const int64_t i0 = (int64_t)blockIdx.x * kNormChunk;
  if (i0 >= n) return;
  const float* xb = x + (int64_t)b * L;
  float lo = INFINITY, hi = -INFINITY;
  const int64_t i1 = min(i0 + kNormChunk, n);
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
    const float v = __ldg(xb + i);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = lo; red[1][w] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { lo = fminf(lo, red[0][i]); hi = fmaxf(hi, red[1][i]); }
    atomicMin(ext + 2 * b, f2ord(lo));
    atomicMax(ext + 2 * b + 1, f2ord(hi));
  }
}

__global__ void __launch_bounds__(256) k_norm_apply(const float* __restrict__ x, int64_t L, const int64_t* __restrict__ lengths,
                                                    const unsigned* __restrict__ ext, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int64_t n = lengths != nullptr ? min(lengths[b], L) : L;
  const float mn = ord2f(ext[2 * b]), mx = ord2f(ext[2 * b + 1]);
  const float range = __fsub_rn(mx, mn);
  const float* xb = x + (int64_t)b * L;
  float* ob = out + (int64_t)b * L;
  const int64_t i0 = (int64_t)blockIdx.x * kNormChunk, i1 = min(i0 + kNormChunk, L);
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
    float v = 0.f;   // samples past a ragged utterance's length are written as zeros
    if (i < n) v = __fsub_rn(__fdiv_rn(__fmul_rn(1.8f, __fsub_rn(__ldg(xb + i), mn)), range), 0.9f);
    ob[i] = v;
  }
}

}  // namespace

void launch_minmax_normalize(const float* x, int B, int64_t L, const int64_t* lengths, unsigned* ext, float* out, cudaStream_t st) {
  k_ext_init<<<(2 * B + 255) / 256, 256, 0, st>>>(ext, 2 * B);
  dim3 grid((unsigned)((L + kNormChunk - 1) / kNormChunk), B);
  k_minmax<<<grid, 256, 0, st>>>(x, L, lengths, ext);
  k_norm_apply<<<grid, 256, 0, st>>>(x, L, lengths, ext, out);
  g_launch_count += 3;
}

}  // namespace septfa
