// Shared device helpers for the septfa kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace septfa {

// Programmatic dependent launch (PDL): every kernel of the forward is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, lets its successor be scheduled early (launch_dependents)
// and waits for its predecessor to complete and flush (griddepcontrol.wait) before it touches any global
// memory another kernel of the chain writes or reads. Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int kNfft = 512;
constexpr int kHop = 256;
constexpr int kBins = 257;       // n_fft/2 + 1
constexpr int kC = 256;          // BN_dim: residual-stream channels (= kBins - 1, DC dropped)
constexpr int kH = 512;          // H_dim: depthwise hidden channels
constexpr int kLogitStride = 576;  // row pitch of the frame-major logits buffer (514 padded to 3 x 192)
constexpr int kVadCol0 = 516;      // logits columns 516..555: the 2 x 20 partial products of VAD.common.conv1_1 (16 B aligned)
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
// ex2.approx + rcp.approx: ~1e-6 absolute on a mask in (0, 1); for the masks only - the VAD probabilities that are
// compared with a threshold use sigmoidf_acc.
// (the approximate reciprocal, one MUFU, instead of the IEEE-rounded one: the masks move by <= 1 ulp, and in the iSTFT kernel the
// rounded reciprocal's instruction sequence held a third of the stall samples)
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// mean / rstd of one utterance from its double accumulators (biased variance, GroupNorm(1,C)).
__device__ __forceinline__ float2 stat_mean_rstd(const Stat2* st, double inv_n, float eps) {
  // mean and the E[x^2] - mean^2 cancellation in double (three cheap DFMA-class ops); the reciprocal square
  // root in float (IEEE sqrt + divide) - a double sqrt/divide is a ~200-instruction software routine that every
  // CTA would wait on before touching its tile.
  const double m = st->s * inv_n;
  double var = st->ss * inv_n - m * m;
  if (var < 0.0) var = 0.0;
  return make_float2((float)m, 1.0f / sqrtf((float)var + eps));
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

// Running GroupNorm statistics of one thread over the utterance segments of its tile. Slots 0 and 1
// (the tile's first two utterances) live in registers; further segments (only when T is smaller
// than the tile) fall back to shared float atomics. Shared float atomics are CAS loops (ATOMS.CAST.SPIN)
// and collapse under same-address contention, so they stay off the common path: flush_warp() reduces
// over the warp with shuffles and issues one atomic per warp and value.
struct SegStat2 {
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  template <int N>
  __device__ __forceinline__ void add(int sg, const float (&v)[N], float* sm) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) { s += v[i]; q = fmaf(v[i], v[i], q); }
    if (sg == 0) { s0 += s; q0 += q; }
    else if (sg == 1) { s1 += s; q1 += q; }
    else { atomicAdd(sm + 2 * sg, s); atomicAdd(sm + 2 * sg + 1, q); }
  }
  // Must be called by all 32 lanes of the warp: shuffle-reduce, then lane 0 stores the warp's partials
  // into its own slot (slots[warp_slot*4 ..]); seg_stats_commit() adds the slots in a fixed order.
  __device__ __forceinline__ void flush_warp(float* slots, int warp_slot) {
    s0 = warp_sum(s0); q0 = warp_sum(q0);
    s1 = warp_sum(s1); q1 = warp_sum(q1);
    if ((threadIdx.x & 31) == 0) {
      slots[warp_slot * 4 + 0] = s0; slots[warp_slot * 4 + 1] = q0;
      slots[warp_slot * 4 + 2] = s1; slots[warp_slot * 4 + 3] = q1;
    }
  }
};

// After a block barrier: thread i < nseg adds the tile's statistics of segment i to the utterance's
// double accumulators. Slots are summed in warp order, so a tile's contribution is bit-reproducible;
// only the order of the (double) global atomics varies between runs.
__device__ __forceinline__ void seg_stats_commit(const float* slots, int nslots, const float* seg_acc, int nseg,
                                                 Stat2* dst) {
  for (int i = threadIdx.x; i < nseg; i += blockDim.x) {
    float s = seg_acc[2 * i], q = seg_acc[2 * i + 1];
    if (i < 2)
      for (int w = 0; w < nslots; ++w) { s += slots[w * 4 + 2 * i]; q += slots[w * 4 + 2 * i + 1]; }
    atomicAdd(&dst[i].s, (double)s);
    atomicAdd(&dst[i].ss, (double)q);
  }
}

// Row -> utterance segment of a tile without an integer division on the common path.
struct SegMap {
  int e1, e2, T, b_first;  // e1 / e2: first global row of the tile's 2nd / 3rd utterance
  __device__ __forceinline__ SegMap(int r0, int T_) : T(T_) {
    b_first = r0 / T_;
    e1 = (b_first + 1) * T_;
    e2 = e1 + T_;
  }
  __device__ __forceinline__ int seg(int row) const { return row < e1 ? 0 : (row < e2 ? 1 : row / T - b_first); }
  __device__ __forceinline__ int frame(int row, int sg) const { return row - (b_first + sg) * T; }
};

}  // namespace septfa
