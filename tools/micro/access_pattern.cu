// Micro-benchmark: HBM read efficiency of "K-chunked" tile reads (128-row tiles read as 4 x 256 B or 8 x 128 B
// pieces per row, one piece pass at a time) versus full-row reads. nvcc -arch=sm_100a -O3 access_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

// Each CTA (256 threads) owns a 128-row tile of a [M,256] fp32 matrix. PIECES passes; in pass j every warp step reads
// 4 rows x (1024/PIECES) bytes.
template <int PIECES>
__global__ void __launch_bounds__(256, 2) k_chunked(const float* __restrict__ in, float* __restrict__ out, int M) {
  const int r0 = blockIdx.x * 128;
  constexpr int BYTES = 1024 / PIECES;        // bytes per row piece
  constexpr int E = BYTES / 16;               // 16-byte elements per row piece
  constexpr int TOTAL = 128 * E;              // elements per pass
  float acc = 0.f;
  for (int j = 0; j < PIECES; ++j) {
    for (int base = 0; base < TOTAL; base += 256 * 8) {   // 8 x 16 B loads in flight per thread
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = base + k * 256 + threadIdx.x;
        const int rl = e / E, c = e % E;
        v[k] = __ldg(reinterpret_cast<const float4*>(in + (size_t)(r0 + rl) * 256 + j * (BYTES / 4) + c * 4));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    __syncthreads();
  }
  if (acc == 123.456f) out[blockIdx.x] = acc;
}

int main() {
  const int M = 64256 / 128 * 128;
  float *in, *out;
  cudaMalloc(&in, (size_t)M * 1024 * 3);
  cudaMalloc(&out, 4096 * 4);
  cudaMemset(in, 0, (size_t)M * 1024 * 3);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](auto kern, const char* name) {
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
      const float* src = in + (size_t)(rep % 3) * M * 256;  // rotate buffers (198 MB total > L2)
      cudaEventRecord(e0);
      kern<<<M / 128, 256>>>(src, out, M);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2 && ms < best) best = ms;
    }
    printf("%-28s %.1f us  %.2f TB/s\n", name, best * 1e3, (double)M * 1024 / (best * 1e-3) / 1e12);
  };
  run(k_chunked<1>, "full rows (1 x 1024 B)");
  run(k_chunked<2>, "2 x 512 B pieces");
  run(k_chunked<4>, "4 x 256 B pieces (conv1)");
  run(k_chunked<8>, "8 x 128 B pieces (dconv)");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
