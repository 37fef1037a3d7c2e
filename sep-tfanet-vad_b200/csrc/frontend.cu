// Front-end kernels: STFT (frame, window, 512-point shared-memory radix-2 FFT), power-dB,
// activity gate and the statistics of the TCN input norm.
// Reference: model/model.py:408-419 (Spectrogram/InputSpec -> torch.stft, AmplitudeToDB,
// activity_input Conv2d 3x3 + PReLU) and :333 (TCN.LN statistics).
#include "kernels.h"

namespace septfa {

int g_launch_count = 0;

// Twiddle table exp(-2*pi*i*j/512), j < 256, computed in double on the host.
void make_twiddles(float2* h) {
  for (int j = 0; j < 256; ++j) {
    double a = -2.0 * 3.14159265358979323846 * (double)j / 512.0;
    h[j] = make_float2((float)cos(a), (float)sin(a));
  }
}

// In-place 512-point complex FFT over shared memory, decimation in time; the caller has stored
// the input in bit-reversed order. 256 threads, one butterfly per thread per stage.
// INVERSE uses conjugated twiddles (unnormalised inverse).
template <bool INVERSE>
__device__ __forceinline__ void fft512_smem(float2* buf, const float2* tw) {
  const int k = threadIdx.x;
#pragma unroll
  for (int s = 0; s < 9; ++s) {
    const int half = 1 << s;
    const int pos = k & (half - 1);
    const int i0 = ((k >> s) << (s + 1)) + pos;
    const int i1 = i0 + half;
    __syncthreads();  // also orders the caller's writes of buf / tw before the first stage
    float2 w = tw[pos << (8 - s)];
    if (INVERSE) w.y = -w.y;
    const float2 a = buf[i0], b = buf[i1];
    const float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
    buf[i0] = make_float2(a.x + t.x, a.y + t.y);
    buf[i1] = make_float2(a.x - t.x, a.y - t.y);
  }
  __syncthreads();
}

// One CTA per STFT frame. torch.stft(center=True, pad_mode='reflect', onesided), no normalisation,
// DC bin zeroed (model.py:24,410). S[row, f] complex, P[row, f] = 10 log10(max(|S|^2, 1e-10)).
__global__ void __launch_bounds__(256) k_stft(const float* __restrict__ x, int64_t L, int T,
                                              const float* __restrict__ window,
                                              const float2* __restrict__ twiddle,
                                              float2* __restrict__ S, float* __restrict__ P) {
  __shared__ float2 buf[kNfft];
  __shared__ float2 tw[256];
  const int row = blockIdx.x;
  const int b = row / T, t = row - b * T;
  const float* xb = x + (int64_t)b * L;
  tw[threadIdx.x] = __ldg(twiddle + threadIdx.x);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = threadIdx.x + h * 256;
    int64_t i = (int64_t)t * kHop + n - kNfft / 2;
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    const float v = __ldg(xb + i) * __ldg(window + n);
    buf[__brev((unsigned)n) >> 23] = make_float2(v, 0.f);
  }
  fft512_smem<false>(buf, tw);
  float2* Sr = S + (int64_t)row * kBins;
  float* Pr = P + (int64_t)row * kBins;
  for (int f = threadIdx.x; f < kBins; f += 256) {
    float2 v = buf[f];
    if (f == 0) v = make_float2(0.f, 0.f);
    Sr[f] = v;
    Pr[f] = 10.f * log10f(fmaxf(v.x * v.x + v.y * v.y, 1e-10f));
  }
}

// One CTA per frame: spectrum *= PReLU(Conv2d 3x3 (zero pad 1) over the (257, T) plane), rows
// 1..256 go to the TCN (model.py:414-421); accumulates the TCN.LN statistics.
__global__ void __launch_bounds__(256) k_activity_gate(const float* __restrict__ P, int T, int enabled,
                                                       float k00, float k01, float k02, float k10, float k11,
                                                       float k12, float k20, float k21, float k22, float bias,
                                                       float slope, float* __restrict__ z0,
                                                       float* __restrict__ dc_gated, Stat2* __restrict__ st0) {
  __shared__ float rows[3][kBins + 2];
  __shared__ float red[64];
  const int row = blockIdx.x;
  const int b = row / T, t = row - b * T;
  for (int i = threadIdx.x; i < 3 * (kBins + 2); i += 256) {
    const int j = i / (kBins + 2), f = i - j * (kBins + 2) - 1;  // f in [-1, 257]
    const int tt = t + j - 1;
    float v = 0.f;
    if (tt >= 0 && tt < T && f >= 0 && f < kBins) v = __ldg(P + (int64_t)(row + j - 1) * kBins + f);
    rows[j][f + 1] = v;
  }
  __syncthreads();
  // kernel index [i][j]: i over frequency, j over time (input plane is [257, T])
  auto gate = [&](int f) {
    const float c = rows[1][f + 1];
    if (!enabled) return c;
    float acc = k00 * rows[0][f] + k01 * rows[1][f] + k02 * rows[2][f];
    acc += k10 * rows[0][f + 1] + k11 * rows[1][f + 1] + k12 * rows[2][f + 1];
    acc += k20 * rows[0][f + 2] + k21 * rows[1][f + 2] + k22 * rows[2][f + 2];
    acc += bias;
    return c * prelu(acc, slope);
  };
  const float z = gate(threadIdx.x + 1);
  z0[(int64_t)row * kC + threadIdx.x] = z;
  if (threadIdx.x == 0) dc_gated[row] = gate(0);
  block_stat_atomic(z, z * z, st0 + b, red);
}

void launch_stft(const float* x, int B, int64_t L, int T, const float* window, const float2* twiddle, float2* S, float* P,
                 cudaStream_t st) {
  k_stft<<<B * T, 256, 0, st>>>(x, L, T, window, twiddle, S, P);
  ++g_launch_count;
}

void launch_activity_gate(const float* P, int B, int T, int enabled, const float* k, float bias, float slope,
                          float* z0, float* dc_gated, Stat2* st0, cudaStream_t st) {
  // Conv2d weight [1,1,3,3]: k[i*3+j], i = frequency tap, j = time tap; the kernel's kIJ multiplies
  // rows[J][f+I] (rows[J] holds time tap J), i.e. P[f+I-1][t+J-1].
  k_activity_gate<<<B * T, 256, 0, st>>>(P, T, enabled, k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], k[8], bias,
                                         slope, z0, dc_gated, st0);
  ++g_launch_count;
}

}  // namespace septfa
