// Back-end kernels: VAD head (+ threshold / [1,0,1] smoothing), mask application fused with the
// inverse STFT overlap-add, and the optional torch-layout exports.
// Reference: model/model.py:173-179 (VAD), :423-457 (masks, noisy phase, inference gating),
// :460 (InverseSpectrogram -> torch.istft).
#include "kernels.h"

namespace septfa {

template <bool INVERSE>
__device__ __forceinline__ void fft512_smem_be(float2* buf, const float2* tw) {
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

// ------------------------------------------------------------------------------------------
// VAD.common.conv1_1 (257 -> 4, k5, pad 2) + PReLU on the mask logits of each speaker, plus the
// statistics of GroupNorm(1,4) over the [4,T] plane. One CTA per frame, warp = (speaker, channel).
__global__ void __launch_bounds__(256) k_vad_conv(VadParams p) {
  __shared__ float vals[8];
  const int row = blockIdx.x, b = row / p.T, t = row - b * p.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = warp >> 2, j = warp & 3;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int tt = t + k - 2;
    if (tt < 0 || tt >= p.T) continue;
    const float* lr = p.logits + (int64_t)(row + k - 2) * kLogitStride + s * kBins;
    const float* wr = p.w1t + (k * 4 + j) * kBins;
    for (int f = lane; f < kBins; f += 32) acc = fmaf(__ldg(wr + f), __ldg(lr + f), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const float v = prelu(acc + p.b1[j], p.slope);
    p.c4[(((int64_t)(b * 2 + s)) * p.T + t) * 4 + j] = v;
    vals[warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double sm = 0.0, ssm = 0.0;
    for (int i = 0; i < 4; ++i) {
      const double v = vals[threadIdx.x * 4 + i];
      sm += v;
      ssm += v * v;
    }
    atomicAdd(&p.st_v[b * 2 + threadIdx.x].s, sm);
    atomicAdd(&p.st_v[b * 2 + threadIdx.x].ss, ssm);
  }
}

// GroupNorm(1,4) -> output_layer_vad (4 -> 1, k3, pad 1) -> sigmoid; then the inference-only
// threshold (>=) and [1,0,1] neighbour-OR smoothing with edge copy (model.py:449-451).
// One CTA per (utterance, speaker).
__global__ void __launch_bounds__(256) k_vad_final(VadParams p) {
  const int bs = blockIdx.x;
  const float2 mr = stat_mean_rstd(p.st_v + bs, 1.0 / (4.0 * p.T), 1e-8f);
  const float* c4 = p.c4 + (int64_t)bs * p.T * 4;
  float* prob = p.prob + (int64_t)bs * p.T;
  for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
    float acc = p.b2;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int tt = t + k - 1;
      if (tt < 0 || tt >= p.T) continue;
      const float4 c = __ldg(reinterpret_cast<const float4*>(c4 + (int64_t)tt * 4));
      acc += p.w2[0 * 3 + k] * (((c.x - mr.x) * mr.y) * p.g[0] + p.be[0]);
      acc += p.w2[1 * 3 + k] * (((c.y - mr.x) * mr.y) * p.g[1] + p.be[1]);
      acc += p.w2[2 * 3 + k] * (((c.z - mr.x) * mr.y) * p.g[2] + p.be[2]);
      acc += p.w2[3 * 3 + k] * (((c.w - mr.x) * mr.y) * p.g[3] + p.be[3]);
    }
    prob[t] = sigmoidf_acc(acc);
  }
  if (!p.do_smooth) return;
  __syncthreads();  // prob[] written by this CTA is visible to it after the barrier
  float* sm = p.smooth + (int64_t)bs * p.T;
  for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
    const float d = prob[t] >= p.thr ? 1.f : 0.f;
    float v = d;
    if (t > 0 && t < p.T - 1) {
      const float dl = prob[t - 1] >= p.thr ? 1.f : 0.f;
      const float dr = prob[t + 1] >= p.thr ? 1.f : 0.f;
      v = fminf(dl + dr, 1.f);
    }
    sm[t] = v;
  }
}

void launch_vad(const VadParams& p, cudaStream_t st) {
  k_vad_conv<<<p.M, 256, 0, st>>>(p);
  k_vad_final<<<p.B * 2, 256, 0, st>>>(p);
  g_launch_count += 2;
}

// ------------------------------------------------------------------------------------------
// Mask application + inverse STFT + overlap-add, one CTA per 256-sample output block of one
// (utterance, speaker): out[256 j + n] = (w[256+n] fr_j[256+n] + w[n] fr_{j+1}[n]) / env.
// fr_t = irfft_512(S[t] * sigmoid(logit[s,:,t]) * gate[s,t]) (model.py:429-437,452-455,460).
__global__ void __launch_bounds__(256) k_mask_istft(const float2* __restrict__ S, const float* __restrict__ logits,
                                                    const float* __restrict__ gate, const float* __restrict__ window,
                                                    const float2* __restrict__ twiddle, int64_t L, int T,
                                                    float* __restrict__ out) {
  __shared__ float2 buf[kNfft];
  __shared__ float2 tw[256];
  const int j = blockIdx.x, s = blockIdx.y, b = blockIdx.z;
  const int n = threadIdx.x;
  if ((int64_t)j * kHop >= L) return;
  tw[n] = __ldg(twiddle + n);
  float acc = 0.f, env = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = j + h;
    if (t >= T) break;
    const int64_t row = (int64_t)b * T + t;
    const float g = gate != nullptr ? __ldg(gate + ((int64_t)b * 2 + s) * T + t) : 1.f;
    __syncthreads();  // previous iteration's reads of buf are done
    // Hermitian-extended spectrum in bit-reversed order; thread n fills bins n and 512-n (n>=1), 0 and 256.
    {
      const int f = n;  // 0..255
      float2 e = make_float2(0.f, 0.f);
      if (f >= 1) {
        const float2 sv = __ldg(S + row * kBins + f);
        const float m = sigmoidf_acc(__ldg(logits + row * kLogitStride + s * kBins + f)) * g;
        e = make_float2(sv.x * m, sv.y * m);
      }
      buf[__brev((unsigned)f) >> 23] = e;
      if (f >= 1) buf[__brev((unsigned)(kNfft - f)) >> 23] = make_float2(e.x, -e.y);
      if (f == 0) {
        const float2 sv = __ldg(S + row * kBins + 256);
        const float m = sigmoidf_acc(__ldg(logits + row * kLogitStride + s * kBins + 256)) * g;
        buf[__brev(256u) >> 23] = make_float2(sv.x * m, 0.f);  // imaginary part of Nyquist is ignored by irfft
      }
    }
    fft512_smem_be<true>(buf, tw);
    const int idx = (h == 0) ? (kHop + n) : n;
    const float wv = __ldg(window + idx);
    acc += buf[idx].x * (1.f / (float)kNfft) * wv;
    env += wv * wv;
  }
  const int64_t o = (int64_t)j * kHop + n;
  if (o < L) out[((int64_t)b * 2 + s) * L + o] = acc / env;
}

void launch_mask_istft(const float2* S, const float* logits, const float* gate, const float* window,
                       const float2* twiddle, int B, int64_t L, int T, float* out, cudaStream_t st) {
  dim3 grid(T, 2, B);
  k_mask_istft<<<grid, 256, 0, st>>>(S, logits, gate, window, twiddle, L, T, out);
  ++g_launch_count;
}

// ------------------------------------------------------------------------------------------
// Optional exports in the reference's torch layouts ([.., 257, T], T contiguous): tile transposes
// from the frame-major internal buffers.
__global__ void __launch_bounds__(256) k_export(const float2* __restrict__ S, const float* __restrict__ logits,
                                                const float* __restrict__ gate, const float* __restrict__ z0,
                                                const float* __restrict__ dc_gated, int T, float2* __restrict__ est,
                                                float* __restrict__ mask, float* __restrict__ spectrum,
                                                float* __restrict__ logits_out) {
  __shared__ float tl[32][33];   // logits tile [t][f]
  __shared__ float2 ts[32][33];  // S tile
  __shared__ float tz[32][33];   // gated spectrum tile
  const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32, b = blockIdx.z >> 1, s = blockIdx.z & 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, f = f0 + tx;
    if (t < T && f < kBins) {
      const int64_t row = (int64_t)b * T + t;
      tl[i][tx] = __ldg(logits + row * kLogitStride + s * kBins + f);
      ts[i][tx] = __ldg(S + row * kBins + f);
      if (spectrum != nullptr && s == 0) tz[i][tx] = (f == 0) ? __ldg(dc_gated + row) : __ldg(z0 + row * kC + f - 1);
    }
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int f = f0 + i, t = t0 + tx;
    if (t < T && f < kBins) {
      const float lg = tl[tx][i];
      const float m = sigmoidf_acc(lg);
      const int64_t o = (((int64_t)b * 2 + s) * kBins + f) * T + t;
      if (mask != nullptr) mask[o] = m;
      if (logits_out != nullptr) logits_out[o] = lg;  // [B, 514, T] with n = s*257 + f
      if (est != nullptr) {
        const float g = gate != nullptr ? __ldg(gate + ((int64_t)b * 2 + s) * T + t) : 1.f;
        const float2 sv = ts[tx][i];
        est[o] = make_float2(sv.x * m * g, sv.y * m * g);
      }
      if (spectrum != nullptr && s == 0) spectrum[((int64_t)b * kBins + f) * T + t] = tz[tx][i];
    }
  }
}

void launch_export(const float2* S, const float* logits, const float* gate, const float* z0, const float* dc_gated,
                   int B, int T, float2* est, float* mask, float* spectrum, float* logits_out, cudaStream_t st) {
  if (est == nullptr && mask == nullptr && spectrum == nullptr && logits_out == nullptr) return;
  dim3 grid((T + 31) / 32, (kBins + 31) / 32, B * 2);
  k_export<<<grid, 256, 0, st>>>(S, logits, gate, z0, dc_gated, T, est, mask, spectrum, logits_out);
  ++g_launch_count;
}

}  // namespace septfa
