// Wigner-D edge-frame rotation kernels (HBM-bound gather / scatter side of the block).
//
//  gather_rotate_fwd : x[src] | x[dst] -> rotate into the edge frame, keep |m| <= mmax, emit rows in
//                      m-primary order and (optionally) apply the radial modulation.
//                      Replaces transformer_block.py:250-275 + so3.py:343-360,509-512 + so3.py:322-334
//                      + the `x * x_edge` products of so2_ops.py:150-175 in one pass.
//  gather_rotate_bwd : node-centric, deterministic (no atomics): for every node, walk its outgoing
//                      (src half) and incoming (dst half) edges in CSR order.
//  rotinv_reduce_fwd : value * alpha -> rotate back (Wigner^T with the l > mmax rescale of
//                      so3.py:175-195) -> dst-sorted segmented sum.  Replaces
//                      transformer_block.py:321-331 + so3.py:367-387,516-521 + so3.py:304-318
//                      (index_add_) deterministically; also serves input_block.py:113-129.
//  rotinv_reduce_bwd : edge-parallel gather of the node gradient, rotate, split into d(value), d(alpha).
//
// Wigner matrices are stored block-diagonal: [E, sum_l (2l+1)^2] (so3.py:537 stores dense [E,K,K]).
// One CTA per edge (or node), one thread per channel; the per-thread coefficient column lives in
// registers (LMAX is a template parameter), the edge's Wigner blocks are broadcast from shared memory.
#include "common.cuh"

namespace {

template <int LMAX>
struct KDim { static constexpr int K = (LMAX + 1) * (LMAX + 1); static constexpr int WS = (LMAX + 1) * (4 * (LMAX + 1) * (LMAX + 1) - 1) / 3; };

__device__ __forceinline__ float rescale_l(int l, int mmax) {
  return (l > mmax) ? sqrtf((float)(2 * l + 1) / (float)(2 * mmax + 1)) : 1.0f;
}

// ------------------------------------------------------------------------------------------
template <int LMAX>
__global__ void gather_rotate_fwd_kernel(const float* __restrict__ x, const long long* __restrict__ src,
                                         const long long* __restrict__ dst, const float* __restrict__ wig,
                                         const float* __restrict__ rad, float* __restrict__ out,
                                         const int* __restrict__ pos_of_full, const int* __restrict__ rad_slot,
                                         int C, int Kr, int mmax, int nrad) {
  constexpr int K = KDim<LMAX>::K, WS = KDim<LMAX>::WS;
  __shared__ float sw[WS];
  __shared__ int spos[K];
  __shared__ int sslot[K];
  const long long e = blockIdx.x;
  for (int i = threadIdx.x; i < WS; i += blockDim.x) sw[i] = wig[e * WS + i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    spos[i] = pos_of_full[i];
    sslot[i] = (i < Kr) ? rad_slot[i] : 0;
  }
  __syncthreads();
  const long long ns = src[e], nd = dst[e];
  const int C2 = 2 * C;
  for (int ch = threadIdx.x; ch < C2; ch += blockDim.x) {
    const long long node = (ch < C) ? ns : nd;
    const int c = (ch < C) ? ch : ch - C;
    const float* xp = x + node * (long long)K * C + c;
    float xc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) xc[k] = __ldg(xp + (long long)k * C);
    const float* rp = rad ? rad + e * (long long)nrad + ch : nullptr;
    float* op = out + e * (long long)Kr * C2 + ch;
#pragma unroll
    for (int l = 0; l <= LMAX; ++l) {
      const int n = 2 * l + 1;
      const int mm = (l < mmax) ? l : mmax;
      const float* wl = sw + eqv2_wig_off(l);
      for (int m = -mm; m <= mm; ++m) {
        const int row = l + m;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < n; ++j) acc = fmaf(wl[row * n + j], xc[l * l + j], acc);
        const int p = spos[l * l + row];
        if (rp) acc *= rp[(long long)sslot[p] * C2];
        op[(long long)p * C2] = acc;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
template <int LMAX>
__global__ void gather_rotate_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wig,
                                         const float* __restrict__ rad, const float* __restrict__ dA,
                                         const int* __restrict__ rowptr_src, const int* __restrict__ perm_src,
                                         const int* __restrict__ rowptr_dst, const int* __restrict__ perm_dst,
                                         float* __restrict__ dx, float* __restrict__ drad,
                                         const int* __restrict__ pos_of_full, const int* __restrict__ rad_slot,
                                         int C, int Kr, int mmax, int nrad) {
  constexpr int K = KDim<LMAX>::K, WS = KDim<LMAX>::WS;
  __shared__ float sw[WS];
  __shared__ int spos[K];
  __shared__ int sslot[K];
  const long long node = blockIdx.x;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    spos[i] = pos_of_full[i];
    sslot[i] = (i < Kr) ? rad_slot[i] : 0;
  }
  const int c = threadIdx.x;
  const bool live = c < C;
  const int C2 = 2 * C;
  float xc[K], acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    xc[k] = live ? __ldg(x + (node * K + k) * (long long)C + c) : 0.f;
    acc[k] = 0.f;
  }
  for (int half = 0; half < 2; ++half) {
    const int* rowptr = half ? rowptr_dst : rowptr_src;
    const int* perm = half ? perm_dst : perm_src;
    const int beg = rowptr[node], end = rowptr[node + 1];
    const int ch = half * C + c;
    for (int idx = beg; idx < end; ++idx) {
      const long long e = perm[idx];
      __syncthreads();
      for (int i = threadIdx.x; i < WS; i += blockDim.x) sw[i] = wig[e * WS + i];
      __syncthreads();
      if (!live) continue;
      const float* gp = dA + e * (long long)Kr * C2 + ch;
      const float* rp = rad ? rad + e * (long long)nrad + ch : nullptr;
      float* drp = drad ? drad + e * (long long)nrad + ch : nullptr;
#pragma unroll
      for (int l = 0; l <= LMAX; ++l) {
        const int n = 2 * l + 1;
        const int mm = (l < mmax) ? l : mmax;
        const float* wl = sw + eqv2_wig_off(l);
        for (int m = 0; m <= mm; ++m) {
          // rows (l, +m) and (l, -m) share one radial weight (so2_ops.py:170-175)
          const int rowp = l + m, rown = l - m;
          const int pp = spos[l * l + rowp], pn = spos[l * l + rown];
          const float r = rp ? rp[(long long)sslot[pp] * C2] : 1.0f;
          const float gpv = gp[(long long)pp * C2];
          const float gnv = (m > 0) ? gp[(long long)pn * C2] : 0.f;
          if (drp) {
            float xrp = 0.f, xrn = 0.f;
#pragma unroll
            for (int j = 0; j < n; ++j) {
              xrp = fmaf(wl[rowp * n + j], xc[l * l + j], xrp);
              xrn = fmaf(wl[rown * n + j], xc[l * l + j], xrn);
            }
            drp[(long long)sslot[pp] * C2] = (m > 0) ? (gpv * xrp + gnv * xrn) : gpv * xrp;
          }
          const float gmp = gpv * r, gmn = gnv * r;
#pragma unroll
          for (int j = 0; j < n; ++j) {
            float t = fmaf(wl[rowp * n + j], gmp, acc[l * l + j]);
            acc[l * l + j] = (m > 0) ? fmaf(wl[rown * n + j], gmn, t) : t;
          }
        }
      }
    }
  }
  if (live) {
#pragma unroll
    for (int k = 0; k < K; ++k) dx[(node * K + k) * (long long)C + c] = acc[k];
  }
}

// ------------------------------------------------------------------------------------------
template <int LMAX>
__global__ void rotinv_reduce_fwd_kernel(const float* __restrict__ val, const float* __restrict__ alpha,
                                         const float* __restrict__ wig, const int* __restrict__ rowptr_dst,
                                         const int* __restrict__ perm_dst, float* __restrict__ out,
                                         const int* __restrict__ pos_of_full, int Cv, int rows_used,
                                         long long val_estride, int heads, int mmax, float scale) {
  constexpr int K = KDim<LMAX>::K, WS = KDim<LMAX>::WS;
  __shared__ float sw[WS];
  __shared__ int spos[K];
  const long long node = blockIdx.x;
  for (int i = threadIdx.x; i < K; i += blockDim.x) spos[i] = pos_of_full[i];
  const int c = threadIdx.x;
  const bool live = c < Cv;
  const int vch = heads > 0 ? Cv / heads : Cv;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  const int beg = rowptr_dst[node], end = rowptr_dst[node + 1];
  for (int idx = beg; idx < end; ++idx) {
    const long long e = perm_dst[idx];
    __syncthreads();
    for (int i = threadIdx.x; i < WS; i += blockDim.x) sw[i] = wig[e * WS + i];
    __syncthreads();
    if (!live) continue;
    const float a = alpha ? alpha[e * heads + c / vch] : 1.0f;
    const float* vp = val + e * val_estride + c;
#pragma unroll
    for (int l = 0; l <= LMAX; ++l) {
      const int n = 2 * l + 1;
      const int mm = (l < mmax) ? l : mmax;
      const float* wl = sw + eqv2_wig_off(l);
      for (int m = -mm; m <= mm; ++m) {
        const int row = l + m;
        const int p = spos[l * l + row];
        if (p >= rows_used) continue;
        const float v = vp[(long long)p * Cv] * a;
#pragma unroll
        for (int j = 0; j < n; ++j) acc[l * l + j] = fmaf(wl[row * n + j], v, acc[l * l + j]);
      }
    }
  }
  if (live) {
#pragma unroll
    for (int l = 0; l <= LMAX; ++l) {
      const float f = rescale_l(l, mmax) * scale;
#pragma unroll
      for (int j = 0; j < 2 * l + 1; ++j) out[(node * K + l * l + j) * (long long)Cv + c] = acc[l * l + j] * f;
    }
  }
}

// ------------------------------------------------------------------------------------------
template <int LMAX>
__global__ void rotinv_reduce_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ val,
                                         const float* __restrict__ alpha, const float* __restrict__ wig,
                                         const long long* __restrict__ dst, float* __restrict__ dval,
                                         float* __restrict__ dalpha, const int* __restrict__ pos_of_full,
                                         int Cv, int rows_used, long long val_estride, int heads, int mmax,
                                         float scale) {
  constexpr int K = KDim<LMAX>::K, WS = KDim<LMAX>::WS;
  __shared__ float sw[WS];
  __shared__ int spos[K];
  EQV2_DYN_SMEM(float, spart);   // [blockDim.x] partial d(alpha)
  const long long e = blockIdx.x;
  for (int i = threadIdx.x; i < WS; i += blockDim.x) sw[i] = wig[e * WS + i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) spos[i] = pos_of_full[i];
  __syncthreads();
  const int c = threadIdx.x;
  const bool live = c < Cv;
  const int vch = heads > 0 ? Cv / heads : Cv;
  const long long node = dst[e];
  float da = 0.f;
  if (live) {
    float g[K];
#pragma unroll
    for (int l = 0; l <= LMAX; ++l) {
      const float f = rescale_l(l, mmax) * scale;
#pragma unroll
      for (int j = 0; j < 2 * l + 1; ++j) g[l * l + j] = __ldg(dout + (node * K + l * l + j) * (long long)Cv + c) * f;
    }
    const float a = alpha ? alpha[e * heads + c / vch] : 1.0f;
    const float* vp = val + e * val_estride + c;
    float* dvp = dval + e * val_estride + c;
#pragma unroll
    for (int l = 0; l <= LMAX; ++l) {
      const int n = 2 * l + 1;
      const int mm = (l < mmax) ? l : mmax;
      const float* wl = sw + eqv2_wig_off(l);
      for (int m = -mm; m <= mm; ++m) {
        const int row = l + m;
        const int p = spos[l * l + row];
        if (p >= rows_used) continue;
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < n; ++j) t = fmaf(wl[row * n + j], g[l * l + j], t);
        if (alpha) da = fmaf(t, vp[(long long)p * Cv], da);
        dvp[(long long)p * Cv] = t * a;
      }
    }
  }
  if (dalpha) {
    spart[threadIdx.x] = da;
    __syncthreads();
    if (threadIdx.x < heads) {
      float s = 0.f;
      for (int i = 0; i < vch; ++i) s += spart[threadIdx.x * vch + i];
      dalpha[e * heads + threadIdx.x] = s;
    }
  }
}

inline int round32(int v) { return (v + 31) / 32 * 32; }

}  // namespace

#define EQV2_DISPATCH_LMAX(lmax, CALL)                                          \
  switch (lmax) {                                                               \
    case 1: { CALL(1); } break;                                                 \
    case 2: { CALL(2); } break;                                                 \
    case 3: { CALL(3); } break;                                                 \
    case 4: { CALL(4); } break;                                                 \
    case 5: { CALL(5); } break;                                                 \
    case 6: { CALL(6); } break;                                                 \
    default: eqv2_set_error("lmax=%d unsupported (1..6)", lmax); return 1;      \
  }

extern "C" int eqv2_gather_rotate_fwd(const float* x, const long long* src, const long long* dst, const float* wig,
                                      const float* rad, float* out, const int* pos_of_full, const int* rad_slot,
                                      long long E, int C, int lmax, int mmax, int Kr, int nrad, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(C > 0 && Kr > 0, "gather_rotate_fwd: bad sizes");
  const int threads = min(256, round32(2 * C));
#define CALL(L) EQV2_LAUNCH(gather_rotate_fwd_kernel<L>, dim3((unsigned)E), dim3(threads), 0, stream, x, src, dst, wig, rad, out, pos_of_full, rad_slot, C, Kr, mmax, nrad)
  EQV2_DISPATCH_LMAX(lmax, CALL)
#undef CALL
  EQV2_CHECK_LAUNCH("eqv2_gather_rotate_fwd");
  return 0;
}

extern "C" int eqv2_gather_rotate_bwd(const float* x, const float* wig, const float* rad, const float* dA,
                                      const int* rowptr_src, const int* perm_src, const int* rowptr_dst,
                                      const int* perm_dst, float* dx, float* drad, const int* pos_of_full,
                                      const int* rad_slot, long long N, int C, int lmax, int mmax, int Kr, int nrad,
                                      void* stream) {
  if (N == 0) return 0;
  EQV2_REQUIRE(C > 0 && C <= 1024, "gather_rotate_bwd: C=%d out of range", C);
  const int threads = round32(C);
#define CALL(L) EQV2_LAUNCH(gather_rotate_bwd_kernel<L>, dim3((unsigned)N), dim3(threads), 0, stream, x, wig, rad, dA, rowptr_src, perm_src, rowptr_dst, perm_dst, dx, drad, pos_of_full, rad_slot, C, Kr, mmax, nrad)
  EQV2_DISPATCH_LMAX(lmax, CALL)
#undef CALL
  EQV2_CHECK_LAUNCH("eqv2_gather_rotate_bwd");
  return 0;
}

extern "C" int eqv2_rotinv_reduce_fwd(const float* val, const float* alpha, const float* wig, const int* rowptr_dst,
                                      const int* perm_dst, float* out, const int* pos_of_full, long long N, int Cv,
                                      int rows_used, long long val_estride, int heads, int lmax, int mmax, float scale,
                                      void* stream) {
  if (N == 0) return 0;
  EQV2_REQUIRE(Cv > 0 && Cv <= 1024, "rotinv_reduce_fwd: Cv=%d out of range", Cv);
  EQV2_REQUIRE(alpha == nullptr || (heads > 0 && Cv % heads == 0), "rotinv_reduce_fwd: heads must divide Cv");
  const int threads = round32(Cv);
#define CALL(L) EQV2_LAUNCH(rotinv_reduce_fwd_kernel<L>, dim3((unsigned)N), dim3(threads), 0, stream, val, alpha, wig, rowptr_dst, perm_dst, out, pos_of_full, Cv, rows_used, val_estride, heads, mmax, scale)
  EQV2_DISPATCH_LMAX(lmax, CALL)
#undef CALL
  EQV2_CHECK_LAUNCH("eqv2_rotinv_reduce_fwd");
  return 0;
}

extern "C" int eqv2_rotinv_reduce_bwd(const float* dout, const float* val, const float* alpha, const float* wig,
                                      const long long* dst, float* dval, float* dalpha, const int* pos_of_full,
                                      long long E, int Cv, int rows_used, long long val_estride, int heads, int lmax,
                                      int mmax, float scale, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(Cv > 0 && Cv <= 1024, "rotinv_reduce_bwd: Cv=%d out of range", Cv);
  EQV2_REQUIRE(alpha == nullptr || (heads > 0 && Cv % heads == 0), "rotinv_reduce_bwd: heads must divide Cv");
  const int threads = round32(max(Cv, heads));
#define CALL(L) EQV2_LAUNCH(rotinv_reduce_bwd_kernel<L>, dim3((unsigned)E), dim3(threads), threads * sizeof(float), stream, dout, val, alpha, wig, dst, dval, dalpha, pos_of_full, Cv, rows_used, val_estride, heads, mmax, scale)
  EQV2_DISPATCH_LMAX(lmax, CALL)
#undef CALL
  EQV2_CHECK_LAUNCH("eqv2_rotinv_reduce_bwd");
  return 0;
}
