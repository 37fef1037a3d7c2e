// Pre-processing on the device: the min-max normalisation of only_inference.py:81,
//   normalized = 1.8 * (audio - audio.min()) / (audio.max() - audio.min()) - 0.9,
// batched over utterances (optionally ragged: `lengths`), in the reference's float32 operation order - subtract,
// multiply by float32(1.8), IEEE divide, subtract float32(0.9) - with explicit round-to-nearest intrinsics so that the
// compiler cannot contract them into FMAs: the result is bit-identical to numpy's. A constant signal gives 0/0 = NaN,
// like the reference. Two kernels: per-utterance extrema (ordered-integer atomics), then the element-wise map with
// 16-byte accesses.
#include <algorithm>
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

// XT = float, or int16_t: PCM samples as scipy.io.wavfile.read returns them, converted like only_inference.py:69
// (`np.array(audio, dtype=np.float32)`, exact for 16-bit integers).
template <typename XT>
__global__ void __launch_bounds__(256) k_minmax(const XT* __restrict__ x, int64_t L, const int64_t* __restrict__ lengths,
                                                unsigned* __restrict__ ext /*[B][2]: ordered min, ordered max*/) {
  __shared__ float red[2][8];
  const int b = blockIdx.y;
  const int64_t n = lengths != nullptr ? min(lengths[b], L) : L;
  const int64_t i0 = (int64_t)blockIdx.x * kNormChunk;
  if (i0 >= n) return;
  const XT* xb = x + (int64_t)b * L;
  float lo = INFINITY, hi = -INFINITY;
  const int64_t i1 = min(i0 + kNormChunk, n);
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
    const float v = (float)__ldg(xb + i);
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

template <typename XT>
__global__ void __launch_bounds__(256) k_norm_apply(const XT* __restrict__ x, int64_t L, const int64_t* __restrict__ lengths,
                                                    const unsigned* __restrict__ ext, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int64_t n = lengths != nullptr ? min(lengths[b], L) : L;
  const float mn = ord2f(ext[2 * b]), mx = ord2f(ext[2 * b + 1]);
  const float range = __fsub_rn(mx, mn);
  const XT* xb = x + (int64_t)b * L;
  float* ob = out + (int64_t)b * L;
  const int64_t i0 = (int64_t)blockIdx.x * kNormChunk, i1 = min(i0 + kNormChunk, L);
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
    float v = 0.f;   // samples past a ragged utterance's length are written as zeros
    if (i < n) v = __fsub_rn(__fdiv_rn(__fmul_rn(1.8f, __fsub_rn((float)__ldg(xb + i), mn)), range), 0.9f);
    ob[i] = v;
  }
}

// float32 -> IEEE half, round to nearest even: numpy's `.astype(np.float16)` of save_audio (Our_utils/utlis_inference.py:30-32)
__global__ void __launch_bounds__(256) k_to_half(const float4* __restrict__ in, uint2* __restrict__ out, int64_t n4,
                                                 const float* __restrict__ in_tail, __half* __restrict__ out_tail, int ntail) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(in + i);
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    out[i] = make_uint2(*reinterpret_cast<const unsigned*>(&a), *reinterpret_cast<const unsigned*>(&b));
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) out_tail[threadIdx.x] = __float2half_rn(in_tail[threadIdx.x]);
}

template <typename XT>
void minmax_normalize_t(const XT* x, int B, int64_t L, const int64_t* lengths, unsigned* ext, float* out, cudaStream_t st) {
  k_ext_init<<<(2 * B + 255) / 256, 256, 0, st>>>(ext, 2 * B);
  dim3 grid((unsigned)((L + kNormChunk - 1) / kNormChunk), B);
  k_minmax<XT><<<grid, 256, 0, st>>>(x, L, lengths, ext);
  k_norm_apply<XT><<<grid, 256, 0, st>>>(x, L, lengths, ext, out);
  ctx().launches += 3;
}

}  // namespace

void launch_minmax_normalize(const float* x, int B, int64_t L, const int64_t* lengths, unsigned* ext, float* out, cudaStream_t st) {
  minmax_normalize_t<float>(x, B, L, lengths, ext, out, st);
}
void launch_minmax_normalize_pcm16(const int16_t* x, int B, int64_t L, const int64_t* lengths, unsigned* ext, float* out,
                                   cudaStream_t st) {
  minmax_normalize_t<int16_t>(x, B, L, lengths, ext, out, st);
}
void launch_to_half(const float* in, __half* out, int64_t n, cudaStream_t st) {
  const int64_t n4 = n / 4;
  const int blocks = (int)std::min<int64_t>(148 * 8, (n4 + 255) / 256 + 1);
  k_to_half<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(out), n4, in + n4 * 4,
                                    out + n4 * 4, (int)(n - n4 * 4));
  ctx().launches += 1;
}

}  // namespace septfa

// ---------------------------------------------------------------------------------------------------------------
// SI-SDR (model/combined_loss.py:16-56, `calc_sisdr`) of `rows` signal pairs in one pass over the data: the five
// moments sum(p), sum(t), sum(p t), sum(t t), sum(p p) per row (accumulated in double),
// then, in double,
//   alpha = (<p', t'> + eps) / (<t', t'> + eps),  val = 10 log10((alpha^2 <t', t'> + eps) / (|alpha t' - p'|^2 + eps))
// with p' = p - mean(p), t' = t - mean(t) when zero_mean, eps = float32 machine epsilon as in the reference. The
// reference evaluates the same expression element-wise in float32; the two agree to ~1e-3 dB (tests).
namespace septfa {

namespace {

constexpr int kSdrChunk = 8192;

__global__ void __launch_bounds__(256) k_sisdr_moments(const float* __restrict__ p, const float* __restrict__ t, int64_t n,
                                                       double* __restrict__ acc /*[rows][5]*/) {
  __shared__ double red[5][8];
  const int64_t row = blockIdx.y;
  const float* pr = p + row * n;
  const float* tr = t + row * n;
  const int64_t i0 = (int64_t)blockIdx.x * kSdrChunk, i1 = min(i0 + kSdrChunk, n);
  // products and sums in double: at 60 dB the noise energy is 1e-6 of the terms it is the difference of
  double sp = 0.0, st = 0.0, spt = 0.0, stt = 0.0, spp = 0.0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
    const double a = (double)__ldg(pr + i), b = (double)__ldg(tr + i);
    sp += a; st += b;
    spt = fma(a, b, spt); stt = fma(b, b, stt); spp = fma(a, a, spp);
  }
  sp = warp_sum(sp); st = warp_sum(st); spt = warp_sum(spt); stt = warp_sum(stt); spp = warp_sum(spp);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = sp; red[1][w] = st; red[2][w] = spt; red[3][w] = stt; red[4][w] = spp; }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    atomicAdd(acc + row * 5 + threadIdx.x, s);
  }
}

__global__ void k_sisdr_final(const double* __restrict__ acc, int64_t rows, int64_t n, int zero_mean, float* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const double* a = acc + r * 5;
  const double eps = 1.1920928955078125e-07;   // torch.finfo(torch.float32).eps
  const double mp = zero_mean ? a[0] / (double)n : 0.0, mt = zero_mean ? a[1] / (double)n : 0.0;
  const double spt = a[2] - (double)n * mp * mt, stt = a[3] - (double)n * mt * mt, spp = a[4] - (double)n * mp * mp;
  const double alpha = (spt + eps) / (stt + eps);
  const double ts2 = alpha * alpha * stt;
  double noise2 = ts2 - 2.0 * alpha * spt + spp;
  if (noise2 < 0.0) noise2 = 0.0;
  out[r] = (float)(10.0 * log10((ts2 + eps) / (noise2 + eps)));
}

}  // namespace

void launch_sisdr(const float* p, const float* t, int64_t rows, int64_t n, int zero_mean, double* scratch, float* out,
                  cudaStream_t st) {
  cudaMemsetAsync(scratch, 0, sizeof(double) * 5 * rows, st);
  dim3 grid((unsigned)((n + kSdrChunk - 1) / kSdrChunk), (unsigned)rows);
  k_sisdr_moments<<<grid, 256, 0, st>>>(p, t, n, scratch);
  k_sisdr_final<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(scratch, rows, n, zero_mean, out);
  ctx().launches += 2;
}

}  // namespace septfa
