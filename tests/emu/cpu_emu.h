// TEST INFRASTRUCTURE ONLY -- a tiny CPU emulator for the SIMT subset of CUDA used by
// csrc/*.cu.  It lets the `-m "not gpu"` suite execute the very same kernel source (index
// math, shared-memory staging, barriers, warp shuffles, atomics) on host threads in the build
// container, where no GPU exists.  It is never linked into the product library
// (libeqv2_b200.so is built by nvcc only and the Python package refuses to run without it).
//
// One block at a time; every CUDA thread of the block is a host thread; __syncthreads is a
// pthread barrier; __shared__ is a plain static (blocks run sequentially).
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <pthread.h>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyToSymbolAsyncEmu(void* d, const void* s, size_t n) { memcpy(d, s, n); return 0; }
#define cudaFuncSetAttribute(...) 0
#define cudaFuncAttributeMaxDynamicSharedMemorySize 0

extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern thread_local dim3 blockDim;
extern thread_local dim3 gridDim;
extern pthread_barrier_t emu_block_barrier;
extern pthread_barrier_t emu_warp_barrier[64];
extern volatile uint32_t emu_shfl_buf[64][32];
extern unsigned char* emu_dyn_smem;

static inline void __syncthreads() { pthread_barrier_wait(&emu_block_barrier); }
static inline void __syncwarp(unsigned = 0xffffffffu);
static inline void __threadfence() {}

static inline int emu_linear_tid() {
  return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
}
static inline void __syncwarp(unsigned) { pthread_barrier_wait(&emu_warp_barrier[emu_linear_tid() >> 5]); }
template <typename T>
static inline T emu_shfl(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "4-byte shuffles only");
  int tid = emu_linear_tid();
  int w = tid >> 5, l = tid & 31;
  uint32_t bits;
  memcpy(&bits, &v, 4);
  emu_shfl_buf[w][l] = bits;
  pthread_barrier_wait(&emu_warp_barrier[w]);
  uint32_t got = emu_shfl_buf[w][src_lane & 31];
  pthread_barrier_wait(&emu_warp_barrier[w]);
  T out;
  memcpy(&out, &got, 4);
  return out;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) { (void)width; return emu_shfl(v, (emu_linear_tid() & 31) ^ m); }
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int d, int width = 32) {
  int l = emu_linear_tid() & 31;
  int s = l + d;
  if ((s / width) != (l / width)) s = l;
  return emu_shfl(v, s);
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int s, int width = 32) {
  int l = emu_linear_tid() & 31;
  return emu_shfl(v, (l / width) * width + (s % width));
}

template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int d, int width = 32) {
  int l = emu_linear_tid() & 31;
  int s = l - d;
  if (s < (l / width) * width) s = l;
  return emu_shfl(v, s);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  int tid = emu_linear_tid();
  int w = tid >> 5, l = tid & 31;
  emu_shfl_buf[w][l] = pred ? 1u : 0u;
  pthread_barrier_wait(&emu_warp_barrier[w]);
  unsigned out = 0;
  for (int i = 0; i < 32; ++i) out |= (emu_shfl_buf[w][i] & 1u) << i;
  pthread_barrier_wait(&emu_warp_barrier[w]);
  return out;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline int atomicExch(int* p, int v) { return reinterpret_cast<std::atomic<int>*>(p)->exchange(v); }

static inline float atomicAdd(float* p, float v) {
  std::atomic<uint32_t>* a = reinterpret_cast<std::atomic<uint32_t>*>(p);
  uint32_t old = a->load();
  for (;;) {
    float f;
    memcpy(&f, &old, 4);
    float n = f + v;
    uint32_t nb;
    memcpy(&nb, &n, 4);
    if (a->compare_exchange_weak(old, nb)) return f;
  }
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
  std::atomic<unsigned>* a = reinterpret_cast<std::atomic<unsigned>*>(p);
  unsigned old = a->load();
  while (old < v && !a->compare_exchange_weak(old, v)) {}
  return old;
}
static inline int atomicAdd(int* p, int v) { return reinterpret_cast<std::atomic<int>*>(p)->fetch_add(v); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return reinterpret_cast<std::atomic<unsigned>*>(p)->fetch_add(v); }

template <typename T> static inline T __ldg(const T* p) { return *p; }
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline void sincosf_emu(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }
static inline long long min(long long a, int b) { return a < b ? a : b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }

void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

#define EQV2_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu_launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define EQV2_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu_dyn_smem)
