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

// MUFU.EX2 + MUFU.RCP (2 ulp): the S2 grid evaluates 324 sigmoids per (edge, channel).  `__expf` wraps the EX2 in range
// scaling for denormal results (FSETP + two predicated FMULs per call: 972 of the 8 400 instructions an S2-forward thread
// executes, and that kernel is issue-bound); the flush-to-zero form needs none -- below 2^-126 the sigmoid is 1 to 38 digits
// and an overflowing exponent still gives 1 / inf = 0.  Same bits as before everywhere else.
#ifndef EQV2_CPU_EMU
__device__ __forceinline__ float eqv2_sigmoid(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
#else
__device__ __forceinline__ float eqv2_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
#endif
__device__ __forceinline__ float eqv2_silu(float x) { return x * eqv2_sigmoid(x); }
// d/dx silu(x) = s (1 + x (1 - s))
__device__ __forceinline__ float eqv2_dsilu(float x) {
  float s = eqv2_sigmoid(x);
  return s * (1.0f + x * (1.0f - s));
}

// d2/dx2 silu(x) = s (1 - s) (2 + x (1 - 2 s))
__device__ __forceinline__ float eqv2_silu_d2(float x) {
  const float s = eqv2_sigmoid(x);
  return s * (1.0f - s) * (2.0f + x * (1.0f - 2.0f * s));
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
// Asynchronous global -> shared copies (cp.async, SASS LDGSTS) for the software-pipelined node-centric kernels: the data of
// the NEXT edge of a node's serial edge walk is in flight while the current edge is rotated.  (The emulator copies at once.)
#ifndef EQV2_CPU_EMU
__device__ __forceinline__ void eqv2_async_copy4(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void eqv2_async_copy16(float* dst_smem, const float* src) {     // both 16-byte aligned
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void eqv2_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void eqv2_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
#else
inline void eqv2_async_copy4(float* d, const float* s) { *d = *s; }
inline void eqv2_async_copy16(float* d, const float* s) { memcpy(d, s, 16); }
inline void eqv2_async_commit() {}
template <int PENDING>
inline void eqv2_async_wait() {}
#endif

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
// Staged variant used by the producer kernels (measured: per-thread 16-bit / paired 32-bit global stores, one per coefficient
// row and thread, made the plane-writing kernels 40-75 % slower than their fp32 twins -- twice the store requests at half
// the width).  Every thread instead deposits its values in a shared-memory image of the CTA's output tile,
// hi[rows][T] | lo[rows][T] (T = threads = columns of the tile), and ONE thread hands the tile to the bulk-copy engine
// (cp.async.bulk shared -> global, SASS UBLKCP): when the tile's rows are adjacent in the plane it is one contiguous chunk
// per plane (two instructions per CTA), else one chunk per row.  Sizes / addresses are multiples of 16 bytes (T % 8 == 0,
// ld % 8 == 0, plane % 8 == 0).
__device__ __forceinline__ void eqv2_plane_stage(__half* st, int rows, int T, int row, int t, float v, float s) {
  const float x = v * s;
  const __half h = __float2half_rn(x);
  st[row * T + t] = h;
  st[(rows + row) * T + t] = __float2half_rn(x - __half2float(h));
}
__device__ __forceinline__ void eqv2_bulk_s2g(void* dst, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(bytes)
               : "memory");
}
// every thread of the CTA must call; idx0 = element offset of tile row 0, row_stride = elements between tile rows
__device__ __forceinline__ void eqv2_plane_flush(const Eqv2PlaneArgs& P, const __half* st, int rows, int T, long long idx0,
                                                 long long row_stride) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the copy engine
  __syncthreads();
  if (threadIdx.x == 0) {
    __half* hp = reinterpret_cast<__half*>(P.hi);
    if (row_stride == (long long)T) {
      eqv2_bulk_s2g(hp + idx0, st, (unsigned)(rows * T * 2));
      eqv2_bulk_s2g(hp + idx0 + P.plane, st + rows * T, (unsigned)(rows * T * 2));
    } else {
      for (int r = 0; r < rows; ++r) {
        eqv2_bulk_s2g(hp + idx0 + r * row_stride, st + r * T, (unsigned)(T * 2));
        eqv2_bulk_s2g(hp + idx0 + r * row_stride + P.plane, st + (rows + r) * T, (unsigned)(T * 2));
      }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads
  }
}
#else   // the CPU emulator (tests/emu) builds the fp32 instances only
struct __half;
inline void eqv2_plane_stage(__half*, int, int, int, int, float, float) {}
inline void eqv2_plane_flush(const Eqv2PlaneArgs&, const __half*, int, int, long long, long long) {}
inline float eqv2_plane_scale(const Eqv2PlaneArgs&, bool) { return 1.f; }
inline void eqv2_plane_store(const Eqv2PlaneArgs&, long long, float, float) {}
#endif
