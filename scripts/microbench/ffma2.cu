// Microbenchmark (GPU): FP32 FMA issue throughput, scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2.cu -o ffma2 && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int PACKED>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
  const float2 av = make_float2(a, a), bv = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) {
        unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&acc[i]);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(*reinterpret_cast<const unsigned long long*>(&av)),
                     "l"(*reinterpret_cast<const unsigned long long*>(&bv)));
        acc[i] = *reinterpret_cast<float2*>(&r);
      } else {
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(acc[i].x) : "f"(acc[i].x), "f"(a), "f"(b));
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(acc[i].y) : "f"(acc[i].y), "f"(a), "f"(b));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, iters = 20000;
  float* out;
  cudaMalloc(&out, sizeof(float) * blocks * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int packed = 0; packed < 2; ++packed) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (packed) k<1><<<blocks, 256>>>(out, iters, 0.999f, 1e-3f); else k<0><<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flops = 2.0 * 16 * (double)iters * blocks * 256;
      if (rep) printf("%s: %.3f ms, %.1f TFLOP/s fp32\n", packed ? "FFMA2 (f32x2)" : "FFMA scalar", ms, flops / ms / 1e9);
    }
  }
  return 0;
}
