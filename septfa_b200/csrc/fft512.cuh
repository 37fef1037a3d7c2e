// 512-point complex FFT in shared memory, radix-8 Stockham (3 passes), shared by the STFT front-end and the
// iSTFT back-end. One FFT is computed by a group of 64 threads (8 points per thread per pass, held in
// registers), so a 256-thread CTA runs 4 independent FFTs side by side and needs 6 block barriers per FFT
// (the first version was radix-2: one butterfly per thread, 10 barriers, 2x the instructions).
#pragma once
#include <cuda_runtime.h>

namespace septfa {

constexpr int kFftPad = 512 + 64;  // float2 elements per padded FFT buffer

// Padded index: one float2 of padding every 8 keeps the stride-8 stores of the passes at the 2-wavefront minimum.
__device__ __forceinline__ int fft_idx(int i) { return i + (i >> 3); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// complex add / subtract on the packed fp32 pipe (FADD2, sm_100): half the instructions of two scalar adds
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

// 8-point DFT in registers (decimation in frequency, natural-order output). INVERSE conjugates the roots.
template <bool INVERSE>
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
  constexpr float h = 0.70710678118654752440f;
  const float sgn = INVERSE ? 1.f : -1.f;  // imaginary sign of exp(-/+ i theta)
  auto mul_i = [&](float2 v) { return INVERSE ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x); };  // * (-/+ i)
  float2 b0 = cadd(a[0], a[4]), b4 = csub(a[0], a[4]);
  float2 b1 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]);
  float2 b2 = cadd(a[2], a[6]), b6 = csub(a[2], a[6]);
  float2 b3 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
  b5 = cmul(b5, make_float2(h, sgn * h));    // W8^1
  b6 = mul_i(b6);                            // W8^2
  b7 = cmul(b7, make_float2(-h, sgn * h));   // W8^3
  const float2 c0 = cadd(b0, b2), c2 = csub(b0, b2), c1 = cadd(b1, b3), c3 = mul_i(csub(b1, b3));
  const float2 c4 = cadd(b4, b6), c6 = csub(b4, b6), c5 = cadd(b5, b7), c7 = mul_i(csub(b5, b7));
  a[0] = cadd(c0, c1); a[4] = csub(c0, c1);
  a[2] = cadd(c2, c3); a[6] = csub(c2, c3);
  a[1] = cadd(c4, c5); a[5] = csub(c4, c5);
  a[3] = cadd(c6, c7); a[7] = csub(c6, c7);
}

// Barrier among the 64 threads (two warps) of FFT group `grp` only: the four groups of a CTA work on independent
// buffers, so they need not wait for each other between passes (named barriers 1..4; 0 is __syncthreads).
__device__ __forceinline__ void group_barrier(int grp) { asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory"); }

// In-place (natural order in, natural order out) FFT of the padded buffer `buf` (kFftPad float2) by the 64 threads
// j = 0..63 of one group. `tw` is the per-pass twiddle table of make_twiddles (512 float2):
//   tw[(r-1)*8 + k]       = exp(-2 pi i r k / 64),  k < 8   (pass 1)
//   tw[64 + (r-1)*64 + k] = exp(-2 pi i r k / 512), k < 64  (pass 2)
// so that the lanes of a warp (consecutive k) read consecutive words: the natural exp(-2 pi i n / 512) table read at
// r * step was an up-to-8-way bank conflict (25 M conflicts per iSTFT launch).
// ALL threads of the CTA must call it together: the passes synchronise per group, the final barrier is block-wide
// (callers read other groups' buffers afterwards). The caller's writes of buf - by the group's own threads - are
// ordered by the barrier of pass 0.
template <bool INVERSE>
__device__ __forceinline__ void fft512_r8(float2* buf, const float2* tw, int j, int grp) {
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const int ns = pass == 0 ? 1 : (pass == 1 ? 8 : 64);
    const int k = j & (ns - 1);
    const int j0 = ((j - k) << 3) + k;
    float2 u[8];
    group_barrier(grp);
#pragma unroll
    for (int r = 0; r < 8; ++r) u[r] = buf[fft_idx(j + 64 * r)];
    if (pass > 0) {
      const float2* twp = pass == 1 ? tw + k : tw + 64 + k;
      const int pitch = pass == 1 ? 8 : 64;
#pragma unroll
      for (int r = 1; r < 8; ++r) {
        float2 w = twp[(r - 1) * pitch];
        if (INVERSE) w.y = -w.y;
        u[r] = cmul(u[r], w);
      }
    }
    dft8<INVERSE>(u);
    group_barrier(grp);
#pragma unroll
    for (int r = 0; r < 8; ++r) buf[fft_idx(j0 + r * ns)] = u[r];
  }
  __syncthreads();
}

}  // namespace septfa
