// 512-point complex FFT over shared memory, shared by the STFT front-end and the iSTFT back-end.
#pragma once
#include <cuda_runtime.h>

namespace septfa {

// In-place, decimation in time; the caller has stored the input in bit-reversed order
// (index __brev(n) >> 23). 256 threads, one radix-2 butterfly per thread per stage, twiddles
// tw[j] = exp(-2*pi*i*j/512), j < 256. INVERSE conjugates the twiddles (unnormalised inverse).
// Starts and ends with a block barrier.
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

}  // namespace septfa
