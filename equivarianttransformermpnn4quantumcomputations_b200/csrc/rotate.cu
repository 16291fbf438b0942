// Wigner-D edge-frame rotation kernels (HBM / L2-bound gather / scatter side of the block).
//
//  gather_rotate_fwd : x[src] | x[dst] -> rotate into the edge frame, keep |m| <= mmax, emit rows in
//                      m-primary order and (optionally) apply the radial modulation.
//                      Replaces transformer_block.py:250-275 + so3.py:343-360,509-512 + so3.py:322-334
//                      + the `x * x_edge` products of so2_ops.py:150-175 in one pass.
//  gather_rotate_bwd : node-centric, deterministic (no atomics): for every node, walk its outgoing
//                      (src half) and incoming (dst half) edges in CSR order -- the two halves run
//                      concurrently in two thread groups and are summed through shared memory.
//  rotinv_reduce_fwd : value * alpha -> rotate back (Wigner^T with the l > mmax rescale of
//                      so3.py:175-195) -> dst-sorted segmented sum.  Replaces
//                      transformer_block.py:321-331 + so3.py:367-387,516-521 + so3.py:304-318
//                      (index_add_) deterministically; also serves input_block.py:113-129.
//                      Two thread groups take alternate edges of the node's segment.
//  rotinv_reduce_bwd : edge-parallel gather of the node gradient, rotate, split into d(value), d(alpha).
//
// Wigner matrices are stored block-diagonal: [E, sum_l (2l+1)^2] (so3.py:537 stores dense [E,K,K]).
// (LMAX, MMAX) are template parameters: every coefficient index (l-major position, m-primary position,
// radial slot) is a compile-time constant, the per-thread coefficient column lives in registers, and the
// edge's Wigner blocks sit in shared memory with rows padded to a multiple of 4 floats so that a row is
// read with broadcast 128-bit loads (one LDS.128 per 4 FMAs instead of one LDS.32 per FMA).
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"

namespace {

template <int L>
struct Dim {
  static constexpr int K = (L + 1) * (L + 1);
  static constexpr int WS = (L + 1) * (4 * (L + 1) * (L + 1) - 1) / 3;
};
__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }
// offset of degree l inside the padded shared-memory copy (rows of 2l+1 floats padded to a multiple of 4)
__host__ __device__ constexpr int wpad_off(int l) {
  int s = 0;
  for (int j = 0; j < l; ++j) s += (2 * j + 1) * pad4(2 * j + 1);
  return s;
}
// m-primary position of coefficient (l, m), |m| <= min(l, M)      (so3.py:97-109)
template <int L, int M>
__host__ __device__ constexpr int mpos(int l, int m) {
  if (m == 0) return l;
  const int am = m < 0 ? -m : m;
  int base = L + 1;
  for (int j = 1; j < am; ++j) base += 2 * (L - j + 1);
  return base + (m < 0 ? (L - am + 1) : 0) + (l - am);
}
// radial-weight slot of rows (l, +-am): slot * C_in + channel indexes the rad vector (so2_ops.py:100-133)
template <int L, int M>
__host__ __device__ constexpr int rslot(int l, int am) {
  if (am == 0) return l;
  int base = L + 1;
  for (int j = 1; j < am; ++j) base += (L - j + 1);
  return base + (l - am);
}
// number of radial-weight slots = number of (l, |m|) pairs with |m| <= min(l, M)
template <int L, int M>
__host__ __device__ constexpr int nslots() {
  int n = L + 1;
  for (int j = 1; j <= M; ++j) n += (L - j + 1);
  return n;
}
__host__ __device__ constexpr float rescale_l(int l, int mmax) {
  // sqrt((2l+1)/(2 mmax+1)) for l > mmax, evaluated at run time where needed (sqrtf is not constexpr)
  return (l > mmax) ? (float)(2 * l + 1) / (float)(2 * mmax + 1) : 1.0f;
}

// cooperative copy of one edge's packed Wigner blocks into the padded shared layout
template <int L>
__device__ __forceinline__ void stage_wigner(float* __restrict__ sw, const float* __restrict__ wig_e, int tid, int nthreads) {
#pragma unroll
  for (int l = 0; l <= L; ++l) {
    const int n = 2 * l + 1, ns = pad4(n);
    const float* src = wig_e + eqv2_wig_off(l);
    float* dst = sw + wpad_off(l);
    for (int i = tid; i < n * n; i += nthreads) dst[(i / n) * ns + (i % n)] = src[i];
  }
}

// acc = sum_j w[row][j] * v[j]   (row of degree l, padded smem, 128-bit broadcast loads)
template <int LDEG>
__device__ __forceinline__ float row_dot(const float* __restrict__ sw, int row, const float* v /* n values */) {
  constexpr int n = 2 * LDEG + 1, ns = pad4(n);
  const float4* wr = reinterpret_cast<const float4*>(sw + wpad_off(LDEG) + row * ns);
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < ns / 4; ++q) {
    const float4 w = wr[q];
    acc = fmaf(w.x, v[4 * q], acc);
    if (4 * q + 1 < n) acc = fmaf(w.y, v[4 * q + 1], acc);
    if (4 * q + 2 < n) acc = fmaf(w.z, v[4 * q + 2], acc);
    if (4 * q + 3 < n) acc = fmaf(w.w, v[4 * q + 3], acc);
  }
  return acc;
}
// v[j] += w[row][j] * s
template <int LDEG>
__device__ __forceinline__ void row_axpy(const float* __restrict__ sw, int row, float s, float* v) {
  constexpr int n = 2 * LDEG + 1, ns = pad4(n);
  const float4* wr = reinterpret_cast<const float4*>(sw + wpad_off(LDEG) + row * ns);
#pragma unroll
  for (int q = 0; q < ns / 4; ++q) {
    const float4 w = wr[q];
    v[4 * q] = fmaf(w.x, s, v[4 * q]);
    if (4 * q + 1 < n) v[4 * q + 1] = fmaf(w.y, s, v[4 * q + 1]);
    if (4 * q + 2 < n) v[4 * q + 2] = fmaf(w.z, s, v[4 * q + 2]);
    if (4 * q + 3 < n) v[4 * q + 3] = fmaf(w.w, s, v[4 * q + 3]);
  }
}

// compile-time loop over degrees
template <int LDEG, int L, int M, typename F>
__device__ __forceinline__ void for_each_degree(F&& f) {
  f(std::integral_constant<int, LDEG>{});
  if constexpr (LDEG < L) for_each_degree<LDEG + 1, L, M>(static_cast<F&&>(f));
}

// compile-time loop i = I .. N-1 (register arrays indexed by i stay in registers)
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(static_cast<F&&>(f));
  }
}

// column `n` values, `stride` apart, into a register array
template <int N>
__device__ __forceinline__ void load_column(float (&v)[N], const float* __restrict__ p, long long stride) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = __ldg(p + (long long)i * stride);
}

// ------------------------------------------------------------------------------------------
// PL = true: the result goes out as scaled fp16 hi/lo planes (the A operand of the first SO(2) convolution's GEMM) instead
// of fp32 (Eqv2PlaneArgs, common.cuh); bound = max |rad| * max |x| * sqrt(2 lmax + 1)  (|(W x)_row| <= ||x_l||_2).
// CT > 0: the channel count is the compile-time constant CT (the launcher checks C == CT): every row stride, the CTA width and
// the staged tile's shape become immediates -- the 67 strided loads and 58 staged stores per thread lose their 64-bit address
// arithmetic (SASS r02: 471 of 1 704 instructions of the CT = 0 instance were IADD3 / IMAD / LEA).
template <int L, int M, bool PL, int CT = 0>
__global__ void __launch_bounds__(256)
gather_rotate_fwd_kernel(const float* __restrict__ x, const long long* __restrict__ src,
                         const long long* __restrict__ dst, const float* __restrict__ wig,
                         const float* __restrict__ rad, float* __restrict__ out, int C_rt, int Kr_rt, int nrad_rt,
                         float* __restrict__ absmax, const Eqv2PlaneArgs PA) {
  constexpr int K = Dim<L>::K, WS = Dim<L>::WS;
  constexpr int NS = nslots<L, M>();
  const int C = CT ? CT : C_rt, Kr = CT ? mpos<L, M>(L, -M) + 1 : Kr_rt, nrad = CT ? NS * 2 * CT : nrad_rt;
  const int T = CT ? (2 * CT < 256 ? 2 * CT : 256) : (int)blockDim.x;
  __shared__ __align__(16) float sw[wpad_off(L + 1)];
  const long long e = blockIdx.x;
  float pscale = 1.f;
  if constexpr (PL) pscale = eqv2_plane_scale(PA, blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0);
  EQV2_DYN_SMEM(__half, stage);          // PL: hi[Kr][T] | lo[Kr][T] image of this CTA's output tile
  const long long ns_ = src[e], nd_ = dst[e];
  const int C2 = 2 * C;
  float amax = 0.f;
  // The x column and all radial weights of this (edge, channel) are requested BEFORE the Wigner blocks are staged:
  // the barrier then waits on one memory latency, not on wigner-then-x-then-rad in sequence (ncu r01: long_scoreboard
  // 12.5 and barrier 2.8 stall cycles per issue with one short-lived CTA per edge).
  float xc[K], rv[NS];
  const int ch = blockIdx.y * T + threadIdx.x;       // channel chunks of T are a grid dimension
  if (ch < C2) {
    load_column<K>(xc, x + ((ch < C) ? ns_ : nd_) * (long long)K * C + ((ch < C) ? ch : ch - C), C);
    if (rad) load_column<NS>(rv, rad + e * (long long)nrad + ch, C2);
  }
  if (!rad) {
#pragma unroll
    for (int i = 0; i < NS; ++i) rv[i] = 1.0f;
  }
  stage_wigner<L>(sw, wig + e * WS, threadIdx.x, T);
  __syncthreads();
  if (ch < C2) {
    float* op = out + e * (long long)Kr * C2 + ch;
    for_each_degree<0, L, M>([&](auto ld) {
      constexpr int l = decltype(ld)::value;
      constexpr int mm = l < M ? l : M;
      static_for<0, 2 * mm + 1>([&](auto mc) {
        constexpr int m = decltype(mc)::value - mm;
        constexpr int p = mpos<L, M>(l, m), sl = rslot<L, M>(l, m < 0 ? -m : m);
        const float acc = row_dot<l>(sw, l + m, xc + l * l) * rv[sl];
        if constexpr (PL) eqv2_plane_stage(stage, Kr, T, p, threadIdx.x, acc, pscale);
        else op[(long long)p * C2] = acc;
        amax = fmaxf(amax, fabsf(acc));
      });
    });
  }
  if constexpr (PL) eqv2_plane_flush(PA, stage, Kr, T, e * PA.ld + (long long)blockIdx.y * T, C2);
  if (absmax != nullptr) eqv2_commit_absmax(amax, absmax);
}

// ------------------------------------------------------------------------------------------
// d(rad): edge-parallel (same shape as the forward).  drad[e, slot, ch] = sum over the rows (l, +-m) sharing the
// slot of  dA[e, row, ch] * (W_e x)[row, ch]   (so2_ops.py:170-175: rows +m and -m share one radial weight).
// PL = true: drad goes out as fp16 hi/lo planes (operand of the radial MLP's backward GEMMs);
// bound = 2 * max |dA| * max |x| * sqrt(2 lmax + 1)  (two rows share a slot).
template <int L, int M, bool PL, int CT = 0>
__global__ void __launch_bounds__(256)
gather_rotate_drad_kernel(const float* __restrict__ x, const long long* __restrict__ src,
                          const long long* __restrict__ dst, const float* __restrict__ wig,
                          const float* __restrict__ dA, float* __restrict__ drad, int C_rt, int Kr_rt, int nrad_rt,
                          float* __restrict__ absmax, const Eqv2PlaneArgs PA) {
  constexpr int K = Dim<L>::K, WS = Dim<L>::WS;
  const int C = CT ? CT : C_rt, Kr = CT ? mpos<L, M>(L, -M) + 1 : Kr_rt, nrad = CT ? nslots<L, M>() * 2 * CT : nrad_rt;
  const int T = CT ? (2 * CT < 256 ? 2 * CT : 256) : (int)blockDim.x;
  __shared__ __align__(16) float sw[wpad_off(L + 1)];
  const long long e = blockIdx.x;
  float pscale = 1.f;
  if constexpr (PL) pscale = eqv2_plane_scale(PA, blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0);
  EQV2_DYN_SMEM(__half, stage);          // PL: hi[NS][T] | lo[NS][T]
  constexpr int NSL = nslots<L, M>();
  const long long ns_ = src[e], nd_ = dst[e];
  const int C2 = 2 * C;
  float amax = 0.f;
  // x and dA columns first, Wigner staging + barrier second (see gather_rotate_fwd_kernel)
  constexpr int KR = mpos<L, M>(L, -M) + 1;      // Kr consecutive m-primary rows
  float xc[K], gv[KR];
  const int ch = blockIdx.y * T + threadIdx.x;
  if (ch < C2) {
    load_column<K>(xc, x + ((ch < C) ? ns_ : nd_) * (long long)K * C + ((ch < C) ? ch : ch - C), C);
    load_column<KR>(gv, dA + e * (long long)Kr * C2 + ch, C2);
  }
  stage_wigner<L>(sw, wig + e * WS, threadIdx.x, T);
  __syncthreads();
  if (ch < C2) {
    float* drp = drad + e * (long long)nrad + ch;
    for_each_degree<0, L, M>([&](auto ld) {
      constexpr int l = decltype(ld)::value;
      constexpr int mm = l < M ? l : M;
      static_for<0, mm + 1>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        constexpr int pp = mpos<L, M>(l, m), pm = mpos<L, M>(l, -m), sl = rslot<L, M>(l, m);
        float d = gv[pp] * row_dot<l>(sw, l + m, xc + l * l);
        if constexpr (m > 0) d = fmaf(gv[pm], row_dot<l>(sw, l - m, xc + l * l), d);
        if constexpr (PL) eqv2_plane_stage(stage, NSL, T, sl, threadIdx.x, d, pscale);
        else drp[(long long)sl * C2] = d;
        amax = fmaxf(amax, fabsf(d));
      });
    });
  }
  if constexpr (PL) eqv2_plane_flush(PA, stage, NSL, T, e * PA.ld + (long long)blockIdx.y * T, C2);
  if (absmax != nullptr) eqv2_commit_absmax(amax, absmax);
}

// Node-centric kernels (deterministic, no atomics).  CTA = (node, slice of CW channels), 256 threads = G groups of CW
// threads (measured: CW = 128 / G = 2 beats 64 / 4 beats 32 / 8 -- every extra group costs a Wigner staging and a
// reduction round); every group walks its share of the node's edges, one edge per iteration:
//   * the edge's data columns (dA and radial weights, or values and attention weight) are requested first and the
//     Wigner blocks staged second, so an iteration exposes one memory latency instead of one per degree
//     (ncu r01: long_scoreboard 6.7 / 25 stall cycles per issue, 640 CTAs of serial edge walks at 12-32 % of HBM peak);
//   * the G partial sums meet in shared memory in group order.
constexpr int NODE_THREADS = 256;
constexpr int NODE_MAX_GROUPS = 8;

// d(x): groups with even index walk the edges where the node is the source (first C columns of dA), odd groups the
// edges where it is the destination (last C columns); sub-group g >> 1 of G / 2 takes every (G/2)-th edge.
template <int L, int M>
__global__ void __launch_bounds__(NODE_THREADS, 2)
gather_rotate_dx_kernel(const float* __restrict__ wig, const float* __restrict__ rad, const float* __restrict__ dA,
                        const int* __restrict__ rowptr_src, const int* __restrict__ perm_src,
                        const int* __restrict__ rowptr_dst, const int* __restrict__ perm_dst, float* __restrict__ dx,
                        int C, int CW /* channels per CTA: 32 or 64 */, int Kr, int nrad) {
  constexpr int K = Dim<L>::K, WS = Dim<L>::WS, WP = wpad_off(L + 1);
  constexpr int KR = mpos<L, M>(L, -M) + 1, NS = nslots<L, M>();
  __shared__ __align__(16) float sw[NODE_MAX_GROUPS][WP];
  EQV2_DYN_SMEM(float, sred);     // [K][CW]
  const long long node = blockIdx.x;
  const int G = NODE_THREADS / CW;
  const int grp = threadIdx.x / CW, ct = threadIdx.x % CW;
  const int half = grp & 1, sub = grp >> 1, nsub = G >> 1;
  const int c = blockIdx.y * CW + ct;
  const bool live = c < C;
  const int C2 = 2 * C;
  const int* rowptr = half ? rowptr_dst : rowptr_src;
  const int* perm = half ? perm_dst : perm_src;
  const int beg = rowptr[node], len = rowptr[node + 1] - beg;
  const int len_other = (half ? rowptr_src : rowptr_dst)[node + 1] - (half ? rowptr_src : rowptr_dst)[node];
  const int maxlen = len > len_other ? len : len_other;
  const int ch = half * C + c;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  for (int it = sub; it < maxlen + sub; it += nsub) {      // same trip count in every group (barriers inside)
    const bool has = it < len;
    long long e = 0;
    float gv[KR], rv[NS];
    if (has) {
      e = perm[beg + it];
      if (live) {
        load_column<KR>(gv, dA + e * (long long)Kr * C2 + ch, C2);
        if (rad) load_column<NS>(rv, rad + e * (long long)nrad + ch, C2);
      }
    }
    __syncthreads();                                       // the previous iteration has finished reading sw
    if (has) stage_wigner<L>(sw[grp], wig + e * WS, ct, CW);
    __syncthreads();
    if (!has || !live) continue;
    const float* w = sw[grp];
    for_each_degree<0, L, M>([&](auto ld) {
      constexpr int l = decltype(ld)::value;
      constexpr int mm = l < M ? l : M;
      static_for<0, mm + 1>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        constexpr int pp = mpos<L, M>(l, m), pm = mpos<L, M>(l, -m), sl = rslot<L, M>(l, m);
        const float r = rad ? rv[sl] : 1.0f;
        row_axpy<l>(w, l + m, gv[pp] * r, acc + l * l);
        if constexpr (m > 0) row_axpy<l>(w, l - m, gv[pm] * r, acc + l * l);
      });
    });
  }
  for (int g = 1; g < G; ++g) {                            // acc(group 0) += acc(group g), in group order
    __syncthreads();
    if (grp == g && live) {
#pragma unroll
      for (int k = 0; k < K; ++k) sred[k * CW + ct] = acc[k];
    }
    __syncthreads();
    if (grp == 0 && live) {
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += sred[k * CW + ct];
    }
  }
  if (grp == 0 && live) {
#pragma unroll
    for (int k = 0; k < K; ++k) dx[(node * K + k) * (long long)C + c] = acc[k];
  }
}

// ---- software-pipelined versions of the two node-centric kernels -------------------------------------------------
// ncu (profiles/r02_ncu_kernel_summary.txt): the serial edge walk exposes one full memory latency per edge (4.4
// long-scoreboard stall cycles per issue, 36 % of HBM peak at 16 resident warps per SM).  Here every group keeps S stages
// in shared memory -- stage = the edge's data columns [ROWS][256] (16-byte cp.async, coalesced 512-byte rows) and its padded
// Wigner blocks -- and requests edge i + S - 1 before it rotates edge i: one barrier per edge, no exposed load latency.
// Same arithmetic and summation order as the kernels above (bit-identical results).
template <int L>
__device__ __forceinline__ void stage_wigner_async(float* __restrict__ sw, const float* __restrict__ wig_e, int tid, int nthreads) {
#pragma unroll
  for (int l = 0; l <= L; ++l) {
    const int n = 2 * l + 1, ns = pad4(n);
    const float* src = wig_e + eqv2_wig_off(l);
    float* dst = sw + wpad_off(l);
    for (int i = tid; i < n * n; i += nthreads) eqv2_async_copy4(dst + (i / n) * ns + (i % n), src + i);
  }
}
// rows x CW floats of one group: global row stride `gstride`, shared row stride NODE_THREADS; c_lim = valid channels
__device__ __forceinline__ void stage_rows_async(float* __restrict__ sdst, const float* __restrict__ gsrc, int rows,
                                                 long long gstride, int CW, int ct, int c_lim) {
  const int mask = (CW >> 2) - 1;                                  // CW / 4 is a power of two (8, 16 or 32)
  const int sh = CW >= 128 ? 5 : (CW >= 64 ? 4 : 3);
  for (int q = ct; q < (rows << sh); q += CW) {
    const int row = q >> sh, c4 = (q & mask) << 2;
    if (c4 < c_lim) eqv2_async_copy16(sdst + row * NODE_THREADS + c4, gsrc + (long long)row * gstride + c4);
  }
}

// CT > 0: C = CW = CT at compile time (the launcher checks), as in the edge-centric kernels: strides and group shape immediate.
template <int L, int M, int S, int CT = 0>
__global__ void __launch_bounds__(NODE_THREADS, 2)
gather_rotate_dx_pipe_kernel(const float* __restrict__ wig, const float* __restrict__ rad, const float* __restrict__ dA,
                             const int* __restrict__ rowptr_src, const int* __restrict__ perm_src,
                             const int* __restrict__ rowptr_dst, const int* __restrict__ perm_dst, float* __restrict__ dx,
                             int C_rt, int CW_rt, int Kr_rt, int nrad_rt) {
  constexpr int K = Dim<L>::K, WS = Dim<L>::WS, WP = wpad_off(L + 1);
  constexpr int KR = mpos<L, M>(L, -M) + 1, NS = nslots<L, M>(), ROWS = KR + NS;
  const int C = CT ? CT : C_rt, CW = CT ? CT : CW_rt, Kr = CT ? KR : Kr_rt, nrad = CT ? NS * 2 * CT : nrad_rt;
  EQV2_DYN_SMEM(float, smem);     // S stages of { columns [ROWS][NODE_THREADS] | Wigner [G][WP] }; reused as sred [K][CW]
  const long long node = blockIdx.x;
  const int G = NODE_THREADS / CW;
  const int stage_floats = ROWS * NODE_THREADS + G * WP;
  const int grp = threadIdx.x / CW, ct = threadIdx.x % CW;
  const int half = grp & 1, sub = grp >> 1, nsub = G >> 1;
  const int c0 = blockIdx.y * CW, c = c0 + ct;
  const bool live = c < C;
  const int C2 = 2 * C;
  const int* rowptr = half ? rowptr_dst : rowptr_src;
  const int* perm = half ? perm_dst : perm_src;
  const int beg = rowptr[node], len = rowptr[node + 1] - beg;
  const int len_other = (half ? rowptr_src : rowptr_dst)[node + 1] - (half ? rowptr_src : rowptr_dst)[node];
  const int maxlen = len > len_other ? len : len_other;
  const int n_iter = (maxlen + nsub - 1) / nsub;            // same trip count in every group (barriers inside)
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  auto issue = [&](int i) {
    const int it = sub + i * nsub;
    if (i < n_iter && it < len) {
      float* st = smem + (i % S) * stage_floats;
      const long long e = perm[beg + it];
      stage_rows_async(st + grp * CW, dA + e * (long long)Kr * C2 + half * C + c0, KR, C2, CW, ct, C - c0);
      if (rad) stage_rows_async(st + KR * NODE_THREADS + grp * CW, rad + e * (long long)nrad + half * C + c0, NS, C2, CW, ct, C - c0);
      stage_wigner_async<L>(st + ROWS * NODE_THREADS + grp * WP, wig + e * WS, ct, CW);
    }
    eqv2_async_commit();
  };
#pragma unroll
  for (int i = 0; i < S - 1; ++i) issue(i);
  for (int i = 0; i < n_iter; ++i) {
    eqv2_async_wait<S - 2>();                               // this thread's copies of stage i have landed ...
    __syncthreads();                                        // ... and everybody's; stage i - 1 is free again
    issue(i + S - 1);
    if (sub + i * nsub >= len || !live) continue;
    const float* st = smem + (i % S) * stage_floats;
    const float* gv = st + threadIdx.x;
    const float* rv = st + KR * NODE_THREADS + threadIdx.x;
    const float* w = st + ROWS * NODE_THREADS + grp * WP;
    for_each_degree<0, L, M>([&](auto ld) {
      constexpr int l = decltype(ld)::value;
      constexpr int mm = l < M ? l : M;
      static_for<0, mm + 1>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        constexpr int pp = mpos<L, M>(l, m), pm = mpos<L, M>(l, -m), sl = rslot<L, M>(l, m);
        const float r = rad ? rv[sl * NODE_THREADS] : 1.0f;
        row_axpy<l>(w, l + m, gv[pp * NODE_THREADS] * r, acc + l * l);
        if constexpr (m > 0) row_axpy<l>(w, l - m, gv[pm * NODE_THREADS] * r, acc + l * l);
      });
    });
  }
  float* sred = smem;
  for (int g = 1; g < G; ++g) {                            // acc(group 0) += acc(group g), in group order
    __syncthreads();
    if (grp == g && live) {
#pragma unroll
      for (int k = 0; k < K; ++k) sred[k * CW + ct] = acc[k];
    }
    __syncthreads();
    if (grp == 0 && live) {
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += sred[k * CW + ct];
    }
  }
  if (grp == 0 && live) {
#pragma unroll
    for (int k = 0; k < K; ++k) dx[(node * K + k) * (long long)C + c] = acc[k];
  }
}

template <int L, int M, int S, int CT = 0>
__global__ void __launch_bounds__(NODE_THREADS, 2)
rotinv_reduce_fwd_pipe_kernel(const float* __restrict__ val, const float* __restrict__ alpha, const float* __restrict__ wig,
                              const int* __restrict__ rowptr_dst, const int* __restrict__ perm_dst, float* __restrict__ out,
                              int Cv_rt, int CW_rt, int rows_used, long long val_estride, int heads, float scale) {
  constexpr int K = Dim<L>::K, WS = Dim<L>::WS, WP = wpad_off(L + 1);
  const int Cv = CT ? CT : Cv_rt, CW = CT ? CT : CW_rt;
  constexpr int KR = mpos<L, M>(L, -M) + 1, ROWS = KR + 1;         // value rows + the attention weight
  EQV2_DYN_SMEM(float, smem);
  const long long node = blockIdx.x;
  const int G = NODE_THREADS / CW;
  const int stage_floats = ROWS * NODE_THREADS + G * WP;
  const int grp = threadIdx.x / CW, ct = threadIdx.x % CW;
  const int c0 = blockIdx.y * CW, c = c0 + ct;
  const bool live = c < Cv;
  const int vch = heads > 0 ? Cv / heads : Cv;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  const int beg = rowptr_dst[node], end = rowptr_dst[node + 1];
  const int n_iter = (end - beg + G - 1) / G;
  auto issue = [&](int i) {
    const int idx = beg + i * G + grp;
    if (i < n_iter && idx < end) {
      float* st = smem + (i % S) * stage_floats;
      const long long e = perm_dst[idx];
      stage_rows_async(st + grp * CW, val + e * val_estride + c0, rows_used, Cv, CW, ct, Cv - c0);
      if (alpha && live) eqv2_async_copy4(st + KR * NODE_THREADS + threadIdx.x, alpha + e * heads + c / vch);
      stage_wigner_async<L>(st + ROWS * NODE_THREADS + grp * WP, wig + e * WS, ct, CW);
    }
    eqv2_async_commit();
  };
#pragma unroll
  for (int i = 0; i < S - 1; ++i) issue(i);
  for (int i = 0; i < n_iter; ++i) {
    eqv2_async_wait<S - 2>();
    __syncthreads();
    issue(i + S - 1);
    if (beg + i * G + grp >= end || !live) continue;
    const float* st = smem + (i % S) * stage_floats;
    const float* vv = st + threadIdx.x;
    const float a = alpha ? st[KR * NODE_THREADS + threadIdx.x] : 1.0f;
    const float* w = st + ROWS * NODE_THREADS + grp * WP;
    for_each_degree<0, L, M>([&](auto ld) {
      constexpr int l = decltype(ld)::value;
      constexpr int mm = l < M ? l : M;
      static_for<0, 2 * mm + 1>([&](auto mc) {
        constexpr int m = decltype(mc)::value - mm;
        constexpr int p = mpos<L, M>(l, m);
        if (p < rows_used) row_axpy<l>(w, l + m, vv[p * NODE_THREADS] * a, acc + l * l);
      });
    });
  }
  float* sred = smem;
  for (int g = 1; g < G; ++g) {
    __syncthreads();
    if (grp == g && live) {
#pragma unroll
      for (int k = 0; k < K; ++k) sred[k * CW + ct] = acc[k];
    }
    __syncthreads();
    if (grp == 0 && live) {
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += sred[k * CW + ct];
    }
  }
  if (grp == 0 && live) {
    for_each_degree<0, L, M>([&](auto ld) {
      constexpr int l = decltype(ld)::value;
      const float f = (l > M ? sqrtf(rescale_l(l, M)) : 1.0f) * scale;
#pragma unroll
      for (int j = 0; j < 2 * l + 1; ++j) out[(node * K + l * l + j) * (long long)Cv + c] = acc[l * l + j] * f;
    });
  }
}

// ------------------------------------------------------------------------------------------
// CTA = (destination node, channel slice); group g takes every G-th edge of the node's segment
template <int L, int M>
__global__ void __launch_bounds__(NODE_THREADS, 2)
rotinv_reduce_fwd_kernel(const float* __restrict__ val, const float* __restrict__ alpha, const float* __restrict__ wig,
                         const int* __restrict__ rowptr_dst, const int* __restrict__ perm_dst, float* __restrict__ out,
                         int Cv, int CW, int rows_used, long long val_estride, int heads, float scale) {
  constexpr int K = Dim<L>::K, WS = Dim<L>::WS, WP = wpad_off(L + 1);
  constexpr int KR = mpos<L, M>(L, -M) + 1;
  __shared__ __align__(16) float sw[NODE_MAX_GROUPS][WP];
  EQV2_DYN_SMEM(float, sred);     // [K][CW]
  const long long node = blockIdx.x;
  const int G = NODE_THREADS / CW;
  const int grp = threadIdx.x / CW, ct = threadIdx.x % CW;
  const int c = blockIdx.y * CW + ct;
  const bool live = c < Cv;
  const int vch = heads > 0 ? Cv / heads : Cv;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  const int beg = rowptr_dst[node], end = rowptr_dst[node + 1];
  for (int idx0 = beg; idx0 < end; idx0 += G) {
    const int idx = idx0 + grp;
    const bool has = idx < end;
    long long e = 0;
    float vv[KR];
    float a = 1.0f;
    if (has) {
      e = perm_dst[idx];
      if (live) {
        if (alpha) a = __ldg(alpha + e * heads + c / vch);
        const float* vp = val + e * val_estride + c;
#pragma unroll
        for (int p = 0; p < KR; ++p) vv[p] = (p < rows_used) ? __ldg(vp + (long long)p * Cv) : 0.f;
      }
    }
    __syncthreads();
    if (has) stage_wigner<L>(sw[grp], wig + e * WS, ct, CW);
    __syncthreads();
    if (!has || !live) continue;
    const float* w = sw[grp];
    for_each_degree<0, L, M>([&](auto ld) {
      constexpr int l = decltype(ld)::value;
      constexpr int mm = l < M ? l : M;
      static_for<0, 2 * mm + 1>([&](auto mc) {
        constexpr int m = decltype(mc)::value - mm;
        constexpr int p = mpos<L, M>(l, m);
        if (p < rows_used) row_axpy<l>(w, l + m, vv[p] * a, acc + l * l);
      });
    });
  }
  for (int g = 1; g < G; ++g) {
    __syncthreads();
    if (grp == g && live) {
#pragma unroll
      for (int k = 0; k < K; ++k) sred[k * CW + ct] = acc[k];
    }
    __syncthreads();
    if (grp == 0 && live) {
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += sred[k * CW + ct];
    }
  }
  if (grp == 0 && live) {
    for_each_degree<0, L, M>([&](auto ld) {
      constexpr int l = decltype(ld)::value;
      const float f = (l > M ? sqrtf(rescale_l(l, M)) : 1.0f) * scale;
#pragma unroll
      for (int j = 0; j < 2 * l + 1; ++j) out[(node * K + l * l + j) * (long long)Cv + c] = acc[l * l + j] * f;
    });
  }
}

// ------------------------------------------------------------------------------------------
// PL = true: d(value) goes out as fp16 hi/lo planes (operand of the second SO(2) convolution's backward GEMMs);
// bound = max |dout| * alpha_bound * scale * sqrt((2 lmax + 1) * max(1, (2 lmax + 1) / (2 mmax + 1))).
template <int L, int M, bool PL, int CT = 0>
__global__ void __launch_bounds__(256)
rotinv_reduce_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ val, const float* __restrict__ alpha,
                         const float* __restrict__ wig, const long long* __restrict__ dst, float* __restrict__ dval,
                         float* __restrict__ dalpha, int Cv_rt, int rows_used, long long val_estride, int heads, float scale,
                         float* __restrict__ absmax, const Eqv2PlaneArgs PA) {
  constexpr int K = Dim<L>::K, WS = Dim<L>::WS;
  const int Cv = CT ? CT : Cv_rt;                       // CT > 0: the launcher checks Cv == CT == CTA width
  const int T = CT ? CT : (int)blockDim.x;
  __shared__ __align__(16) float sw[wpad_off(L + 1)];
  float pscale = 1.f;
  if constexpr (PL) pscale = eqv2_plane_scale(PA, blockIdx.x == 0 && threadIdx.x == 0);
  EQV2_DYN_SMEM(float, spart);   // [T] partial d(alpha); PL: then hi[rows_used][T] | lo[rows_used][T]
  __half* stage = reinterpret_cast<__half*>(spart + T);
  constexpr int KR = mpos<L, M>(L, -M) + 1;
  const long long e = blockIdx.x;
  const int c = threadIdx.x;
  const bool live = c < Cv;
  const int vch = heads > 0 ? Cv / heads : Cv;
  const long long node = dst[e];
  float da = 0.f, amax = 0.f;
  // inputs first (node gradient column, attention weight, value column), Wigner staging + barrier second: one exposed
  // memory latency per CTA instead of three in sequence
  float g[K], vv[KR];
  float a = 1.0f;
  if (live) {
    for_each_degree<0, L, M>([&](auto ld) {
      constexpr int l = decltype(ld)::value;
      const float f = (l > M ? sqrtf(rescale_l(l, M)) : 1.0f) * scale;
#pragma unroll
      for (int j = 0; j < 2 * l + 1; ++j) g[l * l + j] = __ldg(dout + (node * K + l * l + j) * (long long)Cv + c) * f;
    });
    if (alpha) {
      a = __ldg(alpha + e * heads + c / vch);
      const float* vp = val + e * val_estride + c;
#pragma unroll
      for (int p = 0; p < KR; ++p) vv[p] = (p < rows_used) ? __ldg(vp + (long long)p * Cv) : 0.f;
    }
  }
  stage_wigner<L>(sw, wig + e * WS, threadIdx.x, T);
  __syncthreads();
  if (live) {
    float* dvp = dval + e * val_estride + c;
    for_each_degree<0, L, M>([&](auto ld) {
      constexpr int l = decltype(ld)::value;
      constexpr int mm = l < M ? l : M;
      static_for<0, 2 * mm + 1>([&](auto mc) {
        constexpr int m = decltype(mc)::value - mm;
        constexpr int p = mpos<L, M>(l, m);
        if (p < rows_used) {
          const float t = row_dot<l>(sw, l + m, g + l * l);
          if (alpha) da = fmaf(t, vv[p], da);
          if constexpr (PL) eqv2_plane_stage(stage, rows_used, T, p, c, t * a, pscale);
          else dvp[(long long)p * Cv] = t * a;
          amax = fmaxf(amax, fabsf(t * a));
        }
      });
    });
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
  if constexpr (PL) eqv2_plane_flush(PA, stage, rows_used, T, e * PA.ld, Cv);
  if (absmax != nullptr) eqv2_commit_absmax(amax, absmax);
}

inline int round32(int v) { return (v + 31) / 32 * 32; }

}  // namespace

// (lmax, mmax) pairs with kernels: every reference config and the test fixtures
// kernels whose dynamic shared memory can exceed the 48 KB default (staged plane tiles, pipelined node kernels)
static int plane_smem_attr(const void* kfn, size_t smem) {
  if (smem <= 48 * 1024) return 0;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  EQV2_REQUIRE(e == cudaSuccess, "rotate kernels: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
  return 0;
}
constexpr size_t NODE_PIPE_SMEM_MAX = 112 * 1024;        // two CTAs per SM
// EQV2_NODE_PIPE=0 selects the unpipelined node-centric kernels (A/B measurements)
static int node_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v != nullptr && v[0] != 0) ? atoi(v) : dflt;
}
static bool node_ct() {                                 // EQV2_NODE_CT=0: run-time channel count in the pipelined kernels
  const char* v = getenv("EQV2_NODE_CT");
  return v == nullptr || v[0] != '0';
}
static bool node_pipe_enabled() {
  const char* v = getenv("EQV2_NODE_PIPE");      // read per call: tests toggle it
  return v == nullptr || v[0] != '0';
}

#define EQV2_ROT_CONFIGS(X) X(1, 1) X(2, 1) X(2, 2) X(3, 2) X(3, 3) X(4, 2) X(4, 4) X(5, 2) X(6, 2) X(6, 4) X(6, 6)

#define EQV2_ROT_DISPATCH(NAME, BODY)                                                    \
  EQV2_ROT_CONFIGS(BODY)                                                                 \
  eqv2_set_error(NAME ": (lmax, mmax) = (%d, %d) has no kernel instance", lmax, mmax);   \
  return 1;

extern "C" int eqv2_gather_rotate_fwd(const float* x, const long long* src, const long long* dst, const float* wig,
                                      const float* rad, float* out, const int* pos_of_full, const int* rad_slot,
                                      long long E, int C, int lmax, int mmax, int Kr, int nrad, float* absmax,
                                      void* stream) {
  (void)pos_of_full; (void)rad_slot;   // index maps are compile-time now; kept in the ABI for the host tables
  if (E == 0) return 0;
  EQV2_REQUIRE(C > 0 && Kr > 0, "gather_rotate_fwd: bad sizes");
  const int threads = min(256, round32(2 * C));
#define X(L_, M_)                                                                                              \
  if (lmax == L_ && mmax == M_) {                                                                              \
    auto kfn = gather_rotate_fwd_kernel<L_, M_, false>;                                                                             \
    EQV2_LAUNCH(kfn, dim3((unsigned)E, (2 * C + threads - 1) / threads), dim3(threads), 0, stream, x, src, dst, wig, rad, out, C, Kr, nrad, absmax, Eqv2PlaneArgs{}); \
    EQV2_CHECK_LAUNCH("eqv2_gather_rotate_fwd");                                                               \
    return 0;                                                                                                  \
  }
  EQV2_ROT_DISPATCH("gather_rotate_fwd", X)
#undef X
}

// dx[N,K,C] = sum over the node's edges of W_e^T (dA * rad)     (deterministic, node-centric)
extern "C" int eqv2_gather_rotate_dx(const float* wig, const float* rad, const float* dA, const int* rowptr_src,
                                     const int* perm_src, const int* rowptr_dst, const int* perm_dst, float* dx,
                                     long long N, int C, int lmax, int mmax, int Kr, int nrad, void* stream) {
  if (N == 0) return 0;
  EQV2_REQUIRE(C > 0, "gather_rotate_dx: C=%d out of range", C);
  const int CW = C > 64 ? 128 : (C > 32 ? 64 : 32);      // 2 edge groups (source / destination role) x 128 channels
  const size_t smem = (size_t)(lmax + 1) * (lmax + 1) * CW * sizeof(float);
  // pipelined version: 16-byte aligned rows, two stages within half an SM's shared memory (2 CTAs per SM)
  const bool aligned = (C % 4 == 0) && (nrad % 4 == 0) && ((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(rad)) & 15) == 0;
#define X(L_, M_)                                                                                              \
  if (lmax == L_ && mmax == M_) {                                                                              \
    constexpr int ROWS_ = (mpos<L_, M_>(L_, -M_) + 1) + nslots<L_, M_>();                                      \
    const int PCW = node_env_int("EQV2_NODE_CW", CW), PS = node_env_int("EQV2_NODE_STAGES", 2);                 \
    const size_t psmem = PS * ((size_t)ROWS_ * NODE_THREADS + (NODE_THREADS / PCW) * wpad_off(L_ + 1)) * sizeof(float); \
    if (aligned && node_pipe_enabled() && psmem <= (PS == 2 ? NODE_PIPE_SMEM_MAX : 224 * 1024)) {               \
      auto pfn = PS == 3 ? gather_rotate_dx_pipe_kernel<L_, M_, 3>                                             \
                 : (C == 128 && PCW == 128 && Kr == mpos<L_, M_>(L_, -M_) + 1 && nrad == nslots<L_, M_>() * 256 && node_ct()) \
                       ? gather_rotate_dx_pipe_kernel<L_, M_, 2, 128> : gather_rotate_dx_pipe_kernel<L_, M_, 2>; \
      if (plane_smem_attr((const void*)pfn, psmem)) return 1;                                                  \
      EQV2_LAUNCH(pfn, dim3((unsigned)N, (C + PCW - 1) / PCW), dim3(NODE_THREADS), psmem, stream, wig, rad, dA, rowptr_src, perm_src, rowptr_dst, perm_dst, dx, C, PCW, Kr, nrad); \
      EQV2_CHECK_LAUNCH("eqv2_gather_rotate_dx");                                                              \
      return 0;                                                                                                \
    }                                                                                                          \
    auto kfn = gather_rotate_dx_kernel<L_, M_>;                                                                \
    EQV2_LAUNCH(kfn, dim3((unsigned)N, (C + CW - 1) / CW), dim3(NODE_THREADS), smem, stream, wig, rad, dA, rowptr_src, perm_src, rowptr_dst, perm_dst, dx, C, CW, Kr, nrad); \
    EQV2_CHECK_LAUNCH("eqv2_gather_rotate_dx");                                                                \
    return 0;                                                                                                  \
  }
  EQV2_ROT_DISPATCH("gather_rotate_dx", X)
#undef X
}

// drad[E,nrad] = per-edge gradient of the radial weights          (edge-parallel)
extern "C" int eqv2_gather_rotate_drad(const float* x, const long long* src, const long long* dst, const float* wig,
                                       const float* dA, float* drad, long long E, int C, int lmax, int mmax, int Kr,
                                       int nrad, float* absmax, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(C > 0 && Kr > 0, "gather_rotate_drad: bad sizes");
  const int threads = min(256, round32(2 * C));
#define X(L_, M_)                                                                                              \
  if (lmax == L_ && mmax == M_) {                                                                              \
    auto kfn = gather_rotate_drad_kernel<L_, M_, false>;                                                       \
    EQV2_LAUNCH(kfn, dim3((unsigned)E, (2 * C + threads - 1) / threads), dim3(threads), 0, stream, x, src, dst, wig, dA, drad, C, Kr, nrad, absmax, Eqv2PlaneArgs{});     \
    EQV2_CHECK_LAUNCH("eqv2_gather_rotate_drad");                                                              \
    return 0;                                                                                                  \
  }
  EQV2_ROT_DISPATCH("gather_rotate_drad", X)
#undef X
}

extern "C" int eqv2_rotinv_reduce_fwd(const float* val, const float* alpha, const float* wig, const int* rowptr_dst,
                                      const int* perm_dst, float* out, const int* pos_of_full, long long N, int Cv,
                                      int rows_used, long long val_estride, int heads, int lmax, int mmax, float scale,
                                      void* stream) {
  (void)pos_of_full;
  if (N == 0) return 0;
  EQV2_REQUIRE(Cv > 0, "rotinv_reduce_fwd: Cv=%d out of range", Cv);
  EQV2_REQUIRE(alpha == nullptr || (heads > 0 && Cv % heads == 0), "rotinv_reduce_fwd: heads must divide Cv");
  const int CW = Cv > 64 ? 128 : (Cv > 32 ? 64 : 32);
  const size_t smem = (size_t)(lmax + 1) * (lmax + 1) * CW * sizeof(float);
  const bool aligned = (Cv % 4 == 0) && (val_estride % 4 == 0) && (reinterpret_cast<uintptr_t>(val) & 15) == 0;
#define X(L_, M_)                                                                                              \
  if (lmax == L_ && mmax == M_) {                                                                              \
    constexpr int ROWS_ = (mpos<L_, M_>(L_, -M_) + 1) + 1;                                                     \
    const int PCW = node_env_int("EQV2_NODE_CW", CW), PS = node_env_int("EQV2_NODE_STAGES", 2);                 \
    const size_t psmem = PS * ((size_t)ROWS_ * NODE_THREADS + (NODE_THREADS / PCW) * wpad_off(L_ + 1)) * sizeof(float); \
    if (aligned && node_pipe_enabled() && psmem <= (PS == 2 ? NODE_PIPE_SMEM_MAX : 224 * 1024) && rows_used <= ROWS_ - 1) { \
      auto pfn = PS == 3 ? rotinv_reduce_fwd_pipe_kernel<L_, M_, 3>                                            \
                 : (Cv == 128 && PCW == 128 && node_ct()) ? rotinv_reduce_fwd_pipe_kernel<L_, M_, 2, 128>       \
                                                          : rotinv_reduce_fwd_pipe_kernel<L_, M_, 2>;           \
      if (plane_smem_attr((const void*)pfn, psmem)) return 1;                                                  \
      EQV2_LAUNCH(pfn, dim3((unsigned)N, (Cv + PCW - 1) / PCW), dim3(NODE_THREADS), psmem, stream, val, alpha, wig, rowptr_dst, perm_dst, out, Cv, PCW, rows_used, val_estride, heads, scale); \
      EQV2_CHECK_LAUNCH("eqv2_rotinv_reduce_fwd");                                                             \
      return 0;                                                                                                \
    }                                                                                                          \
    auto kfn = rotinv_reduce_fwd_kernel<L_, M_>;                                                                                    \
    EQV2_LAUNCH(kfn, dim3((unsigned)N, (Cv + CW - 1) / CW), dim3(NODE_THREADS), smem, stream, val, alpha, wig, rowptr_dst, perm_dst, out, Cv, CW, rows_used, val_estride, heads, scale); \
    EQV2_CHECK_LAUNCH("eqv2_rotinv_reduce_fwd");                                                               \
    return 0;                                                                                                  \
  }
  EQV2_ROT_DISPATCH("rotinv_reduce_fwd", X)
#undef X
}

extern "C" int eqv2_rotinv_reduce_bwd(const float* dout, const float* val, const float* alpha, const float* wig,
                                      const long long* dst, float* dval, float* dalpha, const int* pos_of_full,
                                      long long E, int Cv, int rows_used, long long val_estride, int heads, int lmax,
                                      int mmax, float scale, float* absmax, void* stream) {
  (void)pos_of_full;
  if (E == 0) return 0;
  EQV2_REQUIRE(Cv > 0 && Cv <= 256, "rotinv_reduce_bwd: Cv=%d out of range", Cv);
  EQV2_REQUIRE(alpha == nullptr || (heads > 0 && Cv % heads == 0), "rotinv_reduce_bwd: heads must divide Cv");
  const int threads = round32(max(Cv, heads));
#define X(L_, M_)                                                                                              \
  if (lmax == L_ && mmax == M_) {                                                                              \
    auto kfn = rotinv_reduce_bwd_kernel<L_, M_, false>;                                                                             \
    EQV2_LAUNCH(kfn, dim3((unsigned)E), dim3(threads), threads * sizeof(float), stream, dout, val, alpha, wig, dst, dval, dalpha, Cv, rows_used, val_estride, heads, scale, absmax, Eqv2PlaneArgs{}); \
    EQV2_CHECK_LAUNCH("eqv2_rotinv_reduce_bwd");                                                               \
    return 0;                                                                                                  \
  }
  EQV2_ROT_DISPATCH("rotinv_reduce_bwd", X)
#undef X
}

#ifndef EQV2_CPU_EMU
// ---- producer-side operand planes: the same three kernels writing scaled fp16 hi/lo planes instead of fp32 ----------
// (the staged tile can exceed the 48 KB default of dynamic shared memory: lmax 6 / mmax 6 is 49 rows x 256 columns x 4 B;
// plane_smem_attr above)

// EQV2_ROT_CT=0 selects the run-time-channel-count instances (A/B measurements)
static bool ct_enabled() {
  const char* v = getenv("EQV2_ROT_CT");
  return v == nullptr || v[0] != '0';
}

static int check_planes(const char* who, const void* planes, long long plane, long long ld, long long cols,
                        const float* bound_a, const float* bound_out) {
  EQV2_REQUIRE(planes != nullptr && bound_a != nullptr && bound_out != nullptr, "%s: null plane / bound pointer", who);
  EQV2_REQUIRE(ld >= cols && (ld % 8) == 0 && (plane % 8) == 0 && (((uintptr_t)planes) & 15) == 0,
               "%s: planes must be 16-byte aligned, ld %% 8 == 0 and ld >= cols", who);
  return 0;
}

extern "C" int eqv2_gather_rotate_fwd_planes(const float* x, const long long* src, const long long* dst,
                                             const float* wig, const float* rad, void* planes, long long plane,
                                             long long ld, const float* bound_x, const float* bound_rad, float* bound_out,
                                             long long E, int C, int lmax, int mmax, int Kr, int nrad, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(C > 0 && Kr > 0 && rad != nullptr && bound_rad != nullptr, "gather_rotate_fwd_planes: bad arguments");
  if (check_planes("gather_rotate_fwd_planes", planes, plane, ld, (long long)Kr * 2 * C, bound_x, bound_out)) return 1;
  EQV2_REQUIRE((2 * C) % 32 == 0 && (2 * C) % min(256, 2 * C) == 0,
               "gather_rotate_fwd_planes: 2 C must be a multiple of 32 and of the CTA width (whole staged tiles)");
  const Eqv2PlaneArgs PA{planes, plane, ld, bound_x, bound_rad, 1.01f * sqrtf((float)(2 * lmax + 1)), bound_out};
  const int threads = min(256, round32(2 * C));
#define X(L_, M_)                                                                                              \
  if (lmax == L_ && mmax == M_) {                                                                              \
    auto kfn = (C == 128 && Kr == mpos<L_, M_>(L_, -M_) + 1 && nrad == nslots<L_, M_>() * 256 && ct_enabled())       \
                   ? gather_rotate_fwd_kernel<L_, M_, true, 128> : gather_rotate_fwd_kernel<L_, M_, true>;     \
    const size_t smem = (size_t)Kr * threads * 4;                                                              \
    if (plane_smem_attr((const void*)kfn, smem)) return 1;                                                     \
    EQV2_LAUNCH(kfn, dim3((unsigned)E, (2 * C + threads - 1) / threads), dim3(threads), smem, stream, x, src, dst, wig, rad, (float*)nullptr, C, Kr, nrad, (float*)nullptr, PA); \
    EQV2_CHECK_LAUNCH("eqv2_gather_rotate_fwd_planes");                                                        \
    return 0;                                                                                                  \
  }
  EQV2_ROT_DISPATCH("gather_rotate_fwd_planes", X)
#undef X
}

extern "C" int eqv2_gather_rotate_drad_planes(const float* x, const long long* src, const long long* dst,
                                              const float* wig, const float* dA, void* planes, long long plane,
                                              long long ld, const float* bound_x, const float* bound_dA, float* bound_out,
                                              long long E, int C, int lmax, int mmax, int Kr, int nrad, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(C > 0 && Kr > 0 && bound_dA != nullptr, "gather_rotate_drad_planes: bad arguments");
  if (check_planes("gather_rotate_drad_planes", planes, plane, ld, nrad, bound_x, bound_out)) return 1;
  EQV2_REQUIRE((2 * C) % 32 == 0 && (2 * C) % min(256, 2 * C) == 0,
               "gather_rotate_drad_planes: 2 C must be a multiple of 32 and of the CTA width (whole staged tiles)");
  const Eqv2PlaneArgs PA{planes, plane, ld, bound_x, bound_dA, 2.02f * sqrtf((float)(2 * lmax + 1)), bound_out};
  const int threads = min(256, round32(2 * C));
#define X(L_, M_)                                                                                              \
  if (lmax == L_ && mmax == M_) {                                                                              \
    auto kfn = (C == 128 && Kr == mpos<L_, M_>(L_, -M_) + 1 && nrad == nslots<L_, M_>() * 256 && ct_enabled())       \
                   ? gather_rotate_drad_kernel<L_, M_, true, 128> : gather_rotate_drad_kernel<L_, M_, true>;   \
    const size_t smem = (size_t)(nrad / (2 * C)) * threads * 4;                                                \
    if (plane_smem_attr((const void*)kfn, smem)) return 1;                                                     \
    EQV2_LAUNCH(kfn, dim3((unsigned)E, (2 * C + threads - 1) / threads), dim3(threads), smem, stream, x, src, dst, wig, dA, (float*)nullptr, C, Kr, nrad, (float*)nullptr, PA); \
    EQV2_CHECK_LAUNCH("eqv2_gather_rotate_drad_planes");                                                       \
    return 0;                                                                                                  \
  }
  EQV2_ROT_DISPATCH("gather_rotate_drad_planes", X)
#undef X
}

extern "C" int eqv2_rotinv_reduce_bwd_planes(const float* dout, const float* val, const float* alpha, const float* wig,
                                             const long long* dst, void* planes, long long plane, long long ld,
                                             const float* bound_dout, float alpha_bound, float* bound_out, float* dalpha,
                                             long long E, int Cv, int rows_used, long long val_estride, int heads,
                                             int lmax, int mmax, float scale, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(Cv > 0 && Cv <= 256, "rotinv_reduce_bwd_planes: Cv=%d out of range", Cv);
  EQV2_REQUIRE(alpha == nullptr || (heads > 0 && Cv % heads == 0), "rotinv_reduce_bwd_planes: heads must divide Cv");
  EQV2_REQUIRE(alpha_bound > 0.f, "rotinv_reduce_bwd_planes: alpha_bound must be positive");
  if (check_planes("rotinv_reduce_bwd_planes", planes, plane, ld, (long long)rows_used * Cv, bound_dout, bound_out)) return 1;
  EQV2_REQUIRE(Cv % 32 == 0 && heads <= Cv, "rotinv_reduce_bwd_planes: Cv must be a multiple of 32 (whole staged tiles)");
  const float resc = (lmax > mmax) ? (float)(2 * lmax + 1) / (float)(2 * mmax + 1) : 1.0f;
  const Eqv2PlaneArgs PA{planes, plane, ld, bound_dout, nullptr,
                         1.01f * alpha_bound * fabsf(scale) * sqrtf((float)(2 * lmax + 1) * resc), bound_out};
  const int threads = round32(max(Cv, heads));
#define X(L_, M_)                                                                                              \
  if (lmax == L_ && mmax == M_) {                                                                              \
    auto kfn = (Cv == 128 && threads == 128 && ct_enabled()) ? rotinv_reduce_bwd_kernel<L_, M_, true, 128>      \
                                                              : rotinv_reduce_bwd_kernel<L_, M_, true>;         \
    const size_t smem = threads * sizeof(float) + (size_t)rows_used * threads * 4;                             \
    if (plane_smem_attr((const void*)kfn, smem)) return 1;                                                     \
    EQV2_LAUNCH(kfn, dim3((unsigned)E), dim3(threads), smem, stream, dout, val, alpha, wig, dst, (float*)nullptr, dalpha, Cv, rows_used, val_estride, heads, scale, (float*)nullptr, PA); \
    EQV2_CHECK_LAUNCH("eqv2_rotinv_reduce_bwd_planes");                                                        \
    return 0;                                                                                                  \
  }
  EQV2_ROT_DISPATCH("rotinv_reduce_bwd_planes", X)
#undef X
}
#endif  // EQV2_CPU_EMU
