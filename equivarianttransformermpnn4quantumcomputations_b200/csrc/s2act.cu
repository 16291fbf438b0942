// Separable S2 activation (activation.py:153-192, grids from so3.py:552-646), fused:
//   grid = to_grid[G,Kr] @ x[Kr,C]  ->  SiLU  ->  out = from_grid[G,Kr]^T @ grid,
//   out row 0 (l = 0) replaced by SiLU(gate).
// The reference materialises the [E,18,18,C] grid tensor (62-166 KB per edge); here the 324 grid
// samples of one (row, channel) never leave registers.  Persistent CTAs keep both grid matrices
// in shared memory and stride over (row, 64-channel chunk) work items.
// The coefficient order is whatever order the matrix columns are given in (m-primary for the
// edge path, l-primary for the node FFN path); columns are zero-padded to KP.
#include "common.cuh"

namespace {

constexpr int CH = 64;       // channels per work item
constexpr int NG = 4;        // grid-point groups per channel
constexpr int S2_THREADS = CH * NG;

template <int KP>
__global__ void __launch_bounds__(S2_THREADS)
s2act_fwd_kernel(const float* __restrict__ X, long long x_rs, const float* __restrict__ gate, long long g_rs,
                 float* __restrict__ O, long long o_rs, const float* __restrict__ Tm, const float* __restrict__ Fm,
                 long long R, int C, int Kr, int G) {
  EQV2_DYN_SMEM(float, smem);
  float* sT = smem;                 // [G][KP]
  float* sF = sT + (size_t)G * KP;  // [G][KP]
  float* red = sF + (size_t)G * KP; // [NG][KP][CH]
  for (int i = threadIdx.x; i < G * KP; i += blockDim.x) {
    sT[i] = Tm[i];
    sF[i] = Fm[i];
  }
  __syncthreads();
  const int cl = threadIdx.x % CH, q = threadIdx.x / CH;
  const int nchunk = (C + CH - 1) / CH;
  const long long nwork = R * nchunk;
  const int gper = (G + NG - 1) / NG;
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
    const long long r = w / nchunk;
    const int c = (int)(w % nchunk) * CH + cl;
    const bool live = c < C;
    float x[KP], o[KP];
#pragma unroll
    for (int p = 0; p < KP; ++p) {
      x[p] = (live && p < Kr) ? __ldg(X + r * x_rs + (long long)p * C + c) : 0.f;
      o[p] = 0.f;
    }
    const int g0 = q * gper, g1 = min(G, g0 + gper);
    for (int g = g0; g < g1; ++g) {
      const float4* t4 = reinterpret_cast<const float4*>(sT + (size_t)g * KP);
      const float4* f4 = reinterpret_cast<const float4*>(sF + (size_t)g * KP);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int p = 0; p < KP / 4; ++p) {
        const float4 t = t4[p];
        a0 = fmaf(t.x, x[4 * p], a0);
        a1 = fmaf(t.y, x[4 * p + 1], a1);
        a2 = fmaf(t.z, x[4 * p + 2], a2);
        a3 = fmaf(t.w, x[4 * p + 3], a3);
      }
      const float s = eqv2_silu((a0 + a1) + (a2 + a3));
#pragma unroll
      for (int p = 0; p < KP / 4; ++p) {
        const float4 f = f4[p];
        o[4 * p] = fmaf(f.x, s, o[4 * p]);
        o[4 * p + 1] = fmaf(f.y, s, o[4 * p + 1]);
        o[4 * p + 2] = fmaf(f.z, s, o[4 * p + 2]);
        o[4 * p + 3] = fmaf(f.w, s, o[4 * p + 3]);
      }
    }
#pragma unroll
    for (int p = 0; p < KP; ++p) red[((size_t)q * KP + p) * CH + cl] = o[p];
    __syncthreads();
    if (live) {
      for (int p = q; p < Kr; p += NG) {
        float v;
        if (p == 0 && gate != nullptr) {
          v = eqv2_silu(__ldg(gate + r * g_rs + c));
        } else {
          v = 0.f;
#pragma unroll
          for (int qq = 0; qq < NG; ++qq) v += red[((size_t)qq * KP + p) * CH + cl];
        }
        O[r * o_rs + (long long)p * C + c] = v;
      }
    }
    __syncthreads();
  }
}

template <int KP>
__global__ void __launch_bounds__(S2_THREADS)
s2act_bwd_kernel(const float* __restrict__ X, long long x_rs, const float* __restrict__ gate, long long g_rs,
                 const float* __restrict__ dO, long long o_rs, float* __restrict__ dX, long long dx_rs,
                 float* __restrict__ dgate, long long dg_rs, const float* __restrict__ Tm,
                 const float* __restrict__ Fm, long long R, int C, int Kr, int G) {
  EQV2_DYN_SMEM(float, smem);
  float* sT = smem;
  float* sF = sT + (size_t)G * KP;
  float* red = sF + (size_t)G * KP;
  for (int i = threadIdx.x; i < G * KP; i += blockDim.x) {
    sT[i] = Tm[i];
    sF[i] = Fm[i];
  }
  __syncthreads();
  const int cl = threadIdx.x % CH, q = threadIdx.x / CH;
  const int nchunk = (C + CH - 1) / CH;
  const long long nwork = R * nchunk;
  const int gper = (G + NG - 1) / NG;
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
    const long long r = w / nchunk;
    const int c = (int)(w % nchunk) * CH + cl;
    const bool live = c < C;
    float x[KP], go[KP], dx[KP];
#pragma unroll
    for (int p = 0; p < KP; ++p) {
      x[p] = (live && p < Kr) ? __ldg(X + r * x_rs + (long long)p * C + c) : 0.f;
      // the l=0 output row is overwritten by the gate path -> no gradient through the grid
      go[p] = (live && p < Kr && (p > 0 || gate == nullptr)) ? __ldg(dO + r * o_rs + (long long)p * C + c) : 0.f;
      dx[p] = 0.f;
    }
    const int g0 = q * gper, g1 = min(G, g0 + gper);
    for (int g = g0; g < g1; ++g) {
      const float4* t4 = reinterpret_cast<const float4*>(sT + (size_t)g * KP);
      const float4* f4 = reinterpret_cast<const float4*>(sF + (size_t)g * KP);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
      for (int p = 0; p < KP / 4; ++p) {
        const float4 t = t4[p];
        const float4 f = f4[p];
        a0 = fmaf(t.x, x[4 * p], a0);
        a1 = fmaf(t.y, x[4 * p + 1], a1);
        a2 = fmaf(t.z, x[4 * p + 2], a2);
        a3 = fmaf(t.w, x[4 * p + 3], a3);
        d0 = fmaf(f.x, go[4 * p], d0);
        d1 = fmaf(f.y, go[4 * p + 1], d1);
        d2 = fmaf(f.z, go[4 * p + 2], d2);
        d3 = fmaf(f.w, go[4 * p + 3], d3);
      }
      const float dg = ((d0 + d1) + (d2 + d3)) * eqv2_dsilu((a0 + a1) + (a2 + a3));
#pragma unroll
      for (int p = 0; p < KP / 4; ++p) {
        const float4 t = t4[p];
        dx[4 * p] = fmaf(t.x, dg, dx[4 * p]);
        dx[4 * p + 1] = fmaf(t.y, dg, dx[4 * p + 1]);
        dx[4 * p + 2] = fmaf(t.z, dg, dx[4 * p + 2]);
        dx[4 * p + 3] = fmaf(t.w, dg, dx[4 * p + 3]);
      }
    }
#pragma unroll
    for (int p = 0; p < KP; ++p) red[((size_t)q * KP + p) * CH + cl] = dx[p];
    __syncthreads();
    if (live) {
      for (int p = q; p < Kr; p += NG) {
        float v = 0.f;
#pragma unroll
        for (int qq = 0; qq < NG; ++qq) v += red[((size_t)qq * KP + p) * CH + cl];
        dX[r * dx_rs + (long long)p * C + c] = v;
      }
      if (q == 0 && gate != nullptr) {
        const float gv = __ldg(gate + r * g_rs + c);
        dgate[r * dg_rs + c] = __ldg(dO + r * o_rs + c) * eqv2_dsilu(gv);
      }
    }
    __syncthreads();
  }
}

template <int KP>
int launch_fwd(const float* X, long long x_rs, const float* gate, long long g_rs, float* O, long long o_rs,
               const float* Tm, const float* Fm, long long R, int C, int Kr, int G, int nblocks, void* stream) {
  const size_t smem = ((size_t)2 * G * KP + (size_t)NG * KP * CH) * sizeof(float);
#ifndef EQV2_CPU_EMU
  cudaFuncSetAttribute(s2act_fwd_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  EQV2_LAUNCH(s2act_fwd_kernel<KP>, dim3(nblocks), dim3(S2_THREADS), smem, stream, X, x_rs, gate, g_rs, O, o_rs, Tm, Fm, R, C, Kr, G);
  return 0;
}
template <int KP>
int launch_bwd(const float* X, long long x_rs, const float* gate, long long g_rs, const float* dO, long long o_rs,
               float* dX, long long dx_rs, float* dgate, long long dg_rs, const float* Tm, const float* Fm,
               long long R, int C, int Kr, int G, int nblocks, void* stream) {
  const size_t smem = ((size_t)2 * G * KP + (size_t)NG * KP * CH) * sizeof(float);
#ifndef EQV2_CPU_EMU
  cudaFuncSetAttribute(s2act_bwd_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  EQV2_LAUNCH(s2act_bwd_kernel<KP>, dim3(nblocks), dim3(S2_THREADS), smem, stream, X, x_rs, gate, g_rs, dO, o_rs, dX, dx_rs, dgate, dg_rs, Tm, Fm, R, C, Kr, G);
  return 0;
}

}  // namespace

#define EQV2_DISPATCH_KP(KPv, CALL)                                               \
  switch (KPv) {                                                                  \
    case 12: CALL(12); break;                                                     \
    case 16: CALL(16); break;                                                     \
    case 20: CALL(20); break;                                                     \
    case 28: CALL(28); break;                                                     \
    case 32: CALL(32); break;                                                     \
    case 52: CALL(52); break;                                                     \
    default: eqv2_set_error("s2act: padded coefficient count %d unsupported", KPv); return 1; \
  }

extern "C" int eqv2_s2act_padded_rows(int Kr) {
  const int opts[6] = {12, 16, 20, 28, 32, 52};
  for (int i = 0; i < 6; ++i)
    if (Kr <= opts[i]) return opts[i];
  return -1;
}

extern "C" int eqv2_s2act_fwd(const float* X, long long x_rs, const float* gate, long long g_rs, float* O,
                              long long o_rs, const float* Tm, const float* Fm, long long R, int C, int Kr, int KP,
                              int G, int nblocks, void* stream) {
  if (R == 0) return 0;
  EQV2_REQUIRE(Kr <= KP && C > 0 && G > 0 && nblocks > 0, "s2act_fwd: bad sizes");
#define CALL(K_) launch_fwd<K_>(X, x_rs, gate, g_rs, O, o_rs, Tm, Fm, R, C, Kr, G, nblocks, stream)
  EQV2_DISPATCH_KP(KP, CALL)
#undef CALL
  EQV2_CHECK_LAUNCH("eqv2_s2act_fwd");
  return 0;
}

extern "C" int eqv2_s2act_bwd(const float* X, long long x_rs, const float* gate, long long g_rs, const float* dO,
                              long long o_rs, float* dX, long long dx_rs, float* dgate, long long dg_rs,
                              const float* Tm, const float* Fm, long long R, int C, int Kr, int KP, int G,
                              int nblocks, void* stream) {
  if (R == 0) return 0;
  EQV2_REQUIRE(Kr <= KP && C > 0 && G > 0 && nblocks > 0, "s2act_bwd: bad sizes");
#define CALL(K_) launch_bwd<K_>(X, x_rs, gate, g_rs, dO, o_rs, dX, dx_rs, dgate, dg_rs, Tm, Fm, R, C, Kr, G, nblocks, stream)
  EQV2_DISPATCH_KP(KP, CALL)
#undef CALL
  EQV2_CHECK_LAUNCH("eqv2_s2act_bwd");
  return 0;
}
