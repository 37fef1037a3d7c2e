// Online-mode kernels: per-stream L1 permutation-invariant matching of a new window against the
// already-emitted signal, and the reorder + emit + tail update.
// Reference: model/online_class_unknown_targets.py:84-94; PITLossWrapper('pw_pt') with nn.L1Loss,
// n_src = 2 (model/pit_wrapper.py:149-177, 261-312); reorder_source_mse (model/combined_loss.py:63-78).
// The reference is only ever run with batch 1, where nn.L1Loss's batch-mean is the per-stream
// mean; here every stream gets its own permutation.
#include "kernels.h"

namespace septfa {

constexpr int kPitChunk = 4096;

// acc[s][i*2+j] += sum_n |a[s,i,n] - b[s,j,n]| over this CTA's chunk.
__global__ void __launch_bounds__(256) k_pit_l1(const float* __restrict__ a, int64_t a_bs, int64_t a_ss,
                                                const float* __restrict__ b, int64_t b_bs, int64_t b_ss, int64_t n,
                                                double* __restrict__ acc) {
  __shared__ float red[4][8];
  const int s = blockIdx.y;
  const int64_t n0 = (int64_t)blockIdx.x * kPitChunk;
  const float* a0 = a + s * a_bs;
  const float* a1 = a0 + a_ss;
  const float* b0 = b + s * b_bs;
  const float* b1 = b0 + b_ss;
  float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
  for (int64_t i = n0 + threadIdx.x; i < min(n0 + kPitChunk, n); i += 256) {
    const float x0 = a0[i], x1 = a1[i], y0 = b0[i], y1 = b1[i];
    v00 += fabsf(x0 - y0);
    v01 += fabsf(x0 - y1);
    v10 += fabsf(x1 - y0);
    v11 += fabsf(x1 - y1);
  }
  v00 = warp_sum(v00); v01 = warp_sum(v01); v10 = warp_sum(v10); v11 = warp_sum(v11);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = v00; red[1][w] = v01; red[2][w] = v10; red[3][w] = v11; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += (double)red[threadIdx.x][i];
    atomicAdd(acc + s * 4 + threadIdx.x, t);
  }
}

// identity if (pw[0][0] + pw[1][1]) <= (pw[1][0] + pw[0][1]) else swap (torch.min: first index wins ties).
__global__ void k_pit_perm(const double* __restrict__ acc, int S, int32_t* __restrict__ perm) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const double ident = acc[s * 4 + 0] + acc[s * 4 + 3];
  const double swap = acc[s * 4 + 2] + acc[s * 4 + 1];
  const int sw = ident <= swap ? 0 : 1;
  perm[s * 2 + 0] = sw;
  perm[s * 2 + 1] = 1 - sw;
}

void launch_pit(const float* a, int64_t a_bs, int64_t a_ss, const float* b, int64_t b_bs, int64_t b_ss, int S, int64_t n,
                double* acc, int32_t* perm, cudaStream_t st) {
  cudaMemsetAsync(acc, 0, sizeof(double) * 4 * S, st);
  dim3 grid((unsigned)((n + kPitChunk - 1) / kPitChunk), S);
  k_pit_l1<<<grid, 256, 0, st>>>(a, a_bs, a_ss, b, b_bs, b_ss, n, acc);
  k_pit_perm<<<(S + 127) / 128, 128, 0, st>>>(acc, S, perm);
  ctx().launches += 2;
}

// emitted[s,i,:] = pred[s, perm[s][i], Lw-hop:]; tail_out = concat(tail_in[:, :, :tail_len], emitted)[-tail_cap:].
__global__ void __launch_bounds__(256) k_online_emit(const float* __restrict__ pred, int64_t Lw,
                                                     const int32_t* __restrict__ perm, int hop, int tail_cap,
                                                     const float* __restrict__ tail_in, int tail_len,
                                                     float* __restrict__ tail_out, float* __restrict__ emitted) {
  const int s = blockIdx.z, i = blockIdx.y;
  const int src = perm[s * 2 + i];
  const float* ps = pred + ((int64_t)s * 2 + src) * Lw + (Lw - hop);
  const int new_len = min(tail_len + hop, tail_cap);
  const int keep = new_len - hop;            // samples carried over from the old tail
  const int drop = tail_len - keep;          // old samples falling off the front
  const float* ti = tail_in + ((int64_t)s * 2 + i) * tail_cap;
  float* to = tail_out + ((int64_t)s * 2 + i) * tail_cap;
  float* em = emitted + ((int64_t)s * 2 + i) * hop;
  for (int n = blockIdx.x * 256 + threadIdx.x; n < new_len; n += gridDim.x * 256) {
    float v;
    if (n < keep) {
      v = ti[drop + n];
    } else {
      v = ps[n - keep];
      em[n - keep] = v;
    }
    to[n] = v;
  }
}

void launch_online_emit(const float* pred, int64_t Lw, const int32_t* perm, int S, int hop, int tail_cap,
                        const float* tail_in, int tail_len, float* tail_out, float* emitted, cudaStream_t st) {
  dim3 grid(16, 2, S);
  k_online_emit<<<grid, 256, 0, st>>>(pred, Lw, perm, hop, tail_cap, tail_in, tail_len, tail_out, emitted);
  ++ctx().launches;
}

}  // namespace septfa
