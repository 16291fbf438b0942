// Shared helpers for the sm_100a kernels behind the C-ABI of include/eqv2_b200.h.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifdef EQV2_CPU_EMU
#include "cpu_emu.h"
#else
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define EQV2_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define EQV2_DYN_SMEM(type, name)                              \
  extern __shared__ __align__(16) unsigned char eqv2_dyn_smem_[]; \
  type* name = reinterpret_cast<type*>(eqv2_dyn_smem_)
#endif

#include "../../include/eqv2_b200.h"

#define EQV2_MAX_LMAX 8
#define EQV2_MAX_K 81

void eqv2_set_error(const char* fmt, ...);

#define EQV2_CHECK_LAUNCH(name)                                                  \
  do {                                                                           \
    cudaError_t e_ = cudaGetLastError();                                         \
    if (e_ != cudaSuccess) {                                                     \
      eqv2_set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));     \
      return 2;                                                                  \
    }                                                                            \
  } while (0)

#define EQV2_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      eqv2_set_error(__VA_ARGS__);     \
      return 1;                        \
    }                                  \
  } while (0)

// MUFU.EX2 + MUFU.RCP (2 ulp): the S2 grid evaluates 324 sigmoids per (edge, channel)
__device__ __forceinline__ float eqv2_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float eqv2_silu(float x) { return x * eqv2_sigmoid(x); }
// d/dx silu(x) = s (1 + x (1 - s))
__device__ __forceinline__ float eqv2_dsilu(float x) {
  float s = eqv2_sigmoid(x);
  return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float eqv2_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float eqv2_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// offset of the (2l+1)x(2l+1) block of degree l inside a packed block-diagonal Wigner row
__host__ __device__ __forceinline__ int eqv2_wig_off(int l) { return l * (4 * l * l - 1) / 3; }

// Running max |v| of everything a launch writes (the f16x3 GEMM engine's operand scale, csrc/gemm_f16.cu).  The slot is
// EQV2_ABSMAX_SLOTS floats (zero-initialised by the caller); the tensor's maximum is the maximum over the slots.
// Block reduction, then ONE fire-and-forget atomic per block, spread over the slots by block index: same-address atomics
// serialise in L2, and waiting for a read of the running value at the tail of every (short) CTA cost 20-28 % in the
// edge-parallel rotate kernels.  Every thread of the block must call (contains __syncthreads).
#define EQV2_ABSMAX_SLOTS 64
__device__ __forceinline__ void eqv2_commit_absmax(float m, float* slot) {
  __shared__ float eqv2_absmax_red[32];
  m = eqv2_warp_max(m);
  if ((threadIdx.x & 31) == 0) eqv2_absmax_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 1; i < nw; ++i) m = fmaxf(m, eqv2_absmax_red[i]);
    // non-negative floats order like their bit patterns
    if (m > 0.f) atomicMax(reinterpret_cast<unsigned*>(slot) + (blockIdx.x & (EQV2_ABSMAX_SLOTS - 1)), __float_as_uint(m));
  }
}
// max over the slots, computed by a full warp (every lane returns it)
__device__ __forceinline__ float eqv2_read_absmax(const float* slot) {
  const int lane = threadIdx.x & 31;
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < EQV2_ABSMAX_SLOTS / 32; ++i) m = fmaxf(m, slot[lane + 32 * i]);
  return eqv2_warp_max(m);
}

// ---- producer-side operand planes (f16x3 GEMM engine, csrc/gemm_f16.cu) ----------------------------------------------
// The engine multiplies fp32 matrices as scaled fp16 hi/lo planes: T ~ (hi + lo) / s_T with s_T a power of two derived
// from a value B >= max |T| (scale_of: s_T B in [2^14, 2^15)).  B need not be the exact maximum: hi + lo keeps 22
// significant bits for |v| >= 2^-18 B and an absolute error <= 2^-40 B below, so a bound that overshoots the true maximum
// by up to ~2^8 still leaves every dot product fp32-class (error floor 2^-32 of the tensor maximum).  A kernel that
// PRODUCES a GEMM operand therefore writes the planes itself -- the fp32 tensor and the separate split pass (one read +
// one write of the tensor) disappear -- with B computed on the device from the maxima of its own inputs:
//   B = max(bound_a) * max(bound_b) * bound_c     (bound_b may be null = 1)
// and stores B where the GEMM reads "max |T|" (bound_out, a zero-initialised absmax slot).
struct Eqv2PlaneArgs {
  void* hi;               // __half [2][rows][ld]; plane 1 (lo) starts `plane` elements after plane 0
  long long plane, ld;
  const float* bound_a;
  const float* bound_b;
  float bound_c;
  float* bound_out;
};
#ifndef EQV2_CPU_EMU
__device__ __forceinline__ void eqv2_scale_of(float amax, float& s, float& inv) {
  if (!(amax > 0.f) || !(amax < 3.0e38f)) { s = 1.f; inv = 1.f; return; }
  int e;
  frexpf(amax, &e);                       // amax = f * 2^e, f in [0.5, 1)
  e = max(e, -100);
  s = ldexpf(1.f, 15 - e);                // s * amax in [2^14, 2^15)
  inv = ldexpf(1.f, e - 15);
}
// scale of this launch's plane output (full warps must call: warp-collective slot reads); one thread publishes B
__device__ __forceinline__ float eqv2_plane_scale(const Eqv2PlaneArgs& P, bool publisher) {
  float B = eqv2_read_absmax(P.bound_a) * P.bound_c;
  if (P.bound_b != nullptr) B *= eqv2_read_absmax(P.bound_b);
  if (publisher) P.bound_out[0] = B;
  float s, inv;
  eqv2_scale_of(B, s, inv);
  return s;
}
__device__ __forceinline__ void eqv2_plane_store(const Eqv2PlaneArgs& P, long long idx, float v, float s) {
  const float x = v * s;
  const __half h = __float2half_rn(x);
  __half* hp = reinterpret_cast<__half*>(P.hi);
  hp[idx] = h;
  hp[idx + P.plane] = __float2half_rn(x - __half2float(h));
}
// Same, for a FULL warp whose lanes hold consecutive columns (idx = base + lane, base even): lane pairs exchange their
// halves so that every lane issues ONE 32-bit store -- even lanes the hi pair, odd lanes the lo pair -- instead of two
// 16-bit stores (the producer kernels issue one store per coefficient row and thread; 16-bit stores doubled that).
__device__ __forceinline__ void eqv2_plane_store_warp(const Eqv2PlaneArgs& P, long long idx, float v, float s) {
  const float x = v * s;
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn(x - __half2float(h));
  const unsigned mine = (unsigned)__half_as_ushort(h) | ((unsigned)__half_as_ushort(l) << 16);
  const unsigned other = __shfl_xor_sync(0xffffffffu, mine, 1);
  __half* hp = reinterpret_cast<__half*>(P.hi);
  if ((threadIdx.x & 1) == 0) *reinterpret_cast<unsigned*>(hp + idx) = (mine & 0xffffu) | (other << 16);
  else *reinterpret_cast<unsigned*>(hp + idx - 1 + P.plane) = (other >> 16) | (mine & 0xffff0000u);
}
#else   // the CPU emulator (tests/emu) builds the fp32 instances only
inline void eqv2_plane_store_warp(const Eqv2PlaneArgs&, long long, float, float) {}
inline float eqv2_plane_scale(const Eqv2PlaneArgs&, bool) { return 1.f; }
inline void eqv2_plane_store(const Eqv2PlaneArgs&, long long, float, float) {}
#endif
