// Separable S2 activation (activation.py:153-192 on the grids of so3.py:552-646), latitude/longitude
// factorised.  The reference multiplies by dense [18,18,Kr] matrices; those matrices are products
//     to_grid[b,a,(l,m)]   = Pt[|m|][b][l] * trig[a][m]
//     from_grid[b,a,(l,m)] = Pf[|m|][b][l] * trig[a][m]          trig = 1 | sqrt2 cos(m alpha_a) | sqrt2 sin(|m| alpha_a)
// (e3nn ToS2Grid / FromS2Grid `shb`/`sha`, so3.py:584-608), so per latitude ring b
//     u[m]   = sum_l Pt[|m|][b][l] x[l,m]                       (Kr FMAs)
//     g[a]   = sum_m trig[a][m] u[m]  ->  s = SiLU(g[a])  ->  v[m] += trig[a][m] s      (2 (2M+1) FMAs per grid point)
//     o[l,m] += Pf[|m|][b][l] v[m]                              (Kr FMAs)
// = 18 (2 Kr + 36 (2M+1)) FMAs instead of 2*324*Kr: 4.8x fewer at (L,M) = (6,2), 3.1x at (6,6).
// The factor tables live in __constant__ memory (uniform operands of the FMAs, no shared-memory traffic);
// the host checks them against the module's to_grid/from_grid buffers before use.
// One thread per (row, channel); x / o stay in registers; coefficient order (l- or m-primary) is a
// compile-time index map.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int S2_MAXL = 6;
constexpr int S2_RES = 18;
constexpr int S2_THREADS = 128;

struct S2Tables {
  float Pt[S2_MAXL + 1][S2_RES][S2_MAXL + 1];   // [|m|][b][l]
  float Pf[S2_MAXL + 1][S2_RES][S2_MAXL + 1];
  float ct[S2_RES][S2_MAXL + 1];                // sqrt2 cos(m alpha_a)  (m = 0: 1)
  float st[S2_RES][S2_MAXL + 1];                // sqrt2 sin(m alpha_a)  (m = 0: unused)
};

constexpr int S2_SLOTS = 6;       // table sets resident at once; ops.py owns the slot map (LRU, slots used by a captured graph are pinned)
#ifdef EQV2_CPU_EMU
static S2Tables g_tab[S2_SLOTS];
#else
__constant__ S2Tables g_tab[S2_SLOTS];
#endif

// position of coefficient (l, +-mi) in the reduced tensor
template <int L, int M, bool MPRIMARY>
__host__ __device__ constexpr int coef_pos(int l, int mi, bool neg) {
  if (MPRIMARY) {
    if (mi == 0) return l;
    int base = L + 1;
    for (int j = 1; j < mi; ++j) base += 2 * (L - j + 1);
    return base + (neg ? (L - mi + 1) : 0) + (l - mi);
  }
  int base = 0;
  for (int j = 0; j < l; ++j) base += (2 * j + 1 < 2 * M + 1) ? 2 * j + 1 : 2 * M + 1;
  const int mm = l < M ? l : M;
  return base + mm + (neg ? -mi : mi);
}

template <int L, int M>
struct KrOf {
  static constexpr int value() {
    int s = 0;
    for (int l = 0; l <= L; ++l) s += (2 * l + 1 < 2 * M + 1) ? 2 * l + 1 : 2 * M + 1;
    return s;
  }
};

// latitude transform of ring b: u_p[mi] / u_n[mi] from coefficient registers
template <int L, int M, bool MP>
__device__ __forceinline__ void lat_fwd(const float (&x)[KrOf<L, M>::value()], const float (*P)[S2_RES][S2_MAXL + 1], int b,
                                        float (&up)[M + 1], float (&un)[M + 1]) {
#pragma unroll
  for (int mi = 0; mi <= M; ++mi) {
    float sp = 0.f, sn = 0.f;
#pragma unroll
    for (int l = mi; l <= L; ++l) {
      const float p = P[mi][b][l];
      sp = fmaf(p, x[coef_pos<L, M, MP>(l, mi, false)], sp);
      if (mi > 0) sn = fmaf(p, x[coef_pos<L, M, MP>(l, mi, true)], sn);
    }
    up[mi] = sp;
    un[mi] = sn;
  }
}
template <int L, int M, bool MP>
__device__ __forceinline__ void lat_bwd(float (&o)[KrOf<L, M>::value()], const float (*P)[S2_RES][S2_MAXL + 1], int b,
                                        const float (&vp)[M + 1], const float (&vn)[M + 1]) {
#pragma unroll
  for (int mi = 0; mi <= M; ++mi) {
#pragma unroll
    for (int l = mi; l <= L; ++l) {
      const float p = P[mi][b][l];
      o[coef_pos<L, M, MP>(l, mi, false)] = fmaf(p, vp[mi], o[coef_pos<L, M, MP>(l, mi, false)]);
      if (mi > 0) o[coef_pos<L, M, MP>(l, mi, true)] = fmaf(p, vn[mi], o[coef_pos<L, M, MP>(l, mi, true)]);
    }
  }
}

// MB = minimum resident CTAs per SM the register allocation must allow.  Unconstrained (MB = 1) ptxas takes 197-255 registers
// for these kernels -> 2 CTAs = 8 warps per SM; MB = 3 caps them at 168 (no spills at lmax 6 / mmax 2, 160 bytes at mmax 6)
// -> 12 warps per SM: measured on one box (profiles/r02an_*) the OC20 step takes 65.33 (MB = 1) / 63.37 (3) / 64.15 ms (4:
// 128 registers), the S2 backward 250 / 209 / 216 us per launch.  3 is the default.
template <int L, int M, bool MP, int MB>
__global__ void __launch_bounds__(S2_THREADS, MB)
s2sep_fwd_kernel(const float* __restrict__ X, long long x_rs, const float* __restrict__ gate, long long g_rs,
                 float* __restrict__ O, long long o_rs, long long R, int C, int slot, float* __restrict__ absmax) {
  constexpr int Kr = KrOf<L, M>::value();
  const S2Tables& T = g_tab[slot];
  const long long total = R * C;
  float amax = 0.f;          // max |O| of what this thread writes (operand scale of the GEMM that consumes O)
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const long long r = w / C;
    const int c = (int)(w % C);
    float x[Kr], o[Kr];
#pragma unroll
    for (int p = 0; p < Kr; ++p) {
      x[p] = __ldg(X + r * x_rs + (long long)p * C + c);
      o[p] = 0.f;
    }
#pragma unroll 1
    for (int b = 0; b < S2_RES; ++b) {
      float up[M + 1], un[M + 1], vp[M + 1], vn[M + 1];
      lat_fwd<L, M, MP>(x, T.Pt, b, up, un);
#pragma unroll
      for (int mi = 0; mi <= M; ++mi) vp[mi] = vn[mi] = 0.f;
#pragma unroll
      for (int a = 0; a < S2_RES; ++a) {
        float g = up[0];
#pragma unroll
        for (int mi = 1; mi <= M; ++mi) g = fmaf(T.ct[a][mi], up[mi], fmaf(T.st[a][mi], un[mi], g));
        const float s = eqv2_silu(g);
        vp[0] += s;
#pragma unroll
        for (int mi = 1; mi <= M; ++mi) {
          vp[mi] = fmaf(T.ct[a][mi], s, vp[mi]);
          vn[mi] = fmaf(T.st[a][mi], s, vn[mi]);
        }
      }
      lat_bwd<L, M, MP>(o, T.Pf, b, vp, vn);
    }
    if (gate != nullptr) o[0] = eqv2_silu(__ldg(gate + r * g_rs + c));
#pragma unroll
    for (int p = 0; p < Kr; ++p) {
      O[r * o_rs + (long long)p * C + c] = o[p];
      amax = fmaxf(amax, fabsf(o[p]));
    }
  }
  if (absmax != nullptr) eqv2_commit_absmax(amax, absmax);
}

#ifndef EQV2_CPU_EMU
// The same forward with the result written as scaled fp16 hi/lo operand planes (the A operand of the second SO(2)
// convolution's GEMM, common.cuh) instead of fp32: the [E, Kr*C] activation tensor and its operand-split pass disappear.
// bound = max |Y| * max(1, ||F||_1 ||T||_inf)  (|SiLU(g)| <= |g| <= ||T||_inf max |x|; the gate row is |SiLU(gate)| <= |gate|),
// a factor 25-80 above typical maxima: inside the 2^8 margin of the plane format.  A CTA-iteration covers 128 consecutive
// (row, channel) pairs = 128 / C whole rows (C divides 128); their [Kr][C] blocks are contiguous in the plane rows and leave
// through the bulk-copy engine, one copy per row and plane.
template <int L, int M, bool MP, int MB>
__global__ void __launch_bounds__(S2_THREADS, MB)
s2sep_fwd_planes_kernel(const float* __restrict__ X, long long x_rs, const float* __restrict__ gate, long long g_rs,
                        long long R, int C, int slot, const Eqv2PlaneArgs PA) {
  constexpr int Kr = KrOf<L, M>::value();
  const S2Tables& T = g_tab[slot];
  EQV2_DYN_SMEM(__half, stage);            // hi [128 / C rows][Kr][C] | lo [...]
  const float pscale = eqv2_plane_scale(PA, blockIdx.x == 0 && threadIdx.x == 0);
  const long long total = R * C;
  const int r_local = threadIdx.x / C, c = threadIdx.x - r_local * C;
  for (long long base = (long long)blockIdx.x * S2_THREADS; base < total; base += (long long)gridDim.x * S2_THREADS) {
    const long long r0 = base / C;
    const long long r = r0 + r_local;
    if (r < R) {
      float x[Kr], o[Kr];
#pragma unroll
      for (int p = 0; p < Kr; ++p) {
        x[p] = __ldg(X + r * x_rs + (long long)p * C + c);
        o[p] = 0.f;
      }
#pragma unroll 1
      for (int b = 0; b < S2_RES; ++b) {
        float up[M + 1], un[M + 1], vp[M + 1], vn[M + 1];
        lat_fwd<L, M, MP>(x, T.Pt, b, up, un);
#pragma unroll
        for (int mi = 0; mi <= M; ++mi) vp[mi] = vn[mi] = 0.f;
#pragma unroll
        for (int a = 0; a < S2_RES; ++a) {
          float g = up[0];
#pragma unroll
          for (int mi = 1; mi <= M; ++mi) g = fmaf(T.ct[a][mi], up[mi], fmaf(T.st[a][mi], un[mi], g));
          const float sv = eqv2_silu(g);
          vp[0] += sv;
#pragma unroll
          for (int mi = 1; mi <= M; ++mi) {
            vp[mi] = fmaf(T.ct[a][mi], sv, vp[mi]);
            vn[mi] = fmaf(T.st[a][mi], sv, vn[mi]);
          }
        }
        lat_bwd<L, M, MP>(o, T.Pf, b, vp, vn);
      }
      if (gate != nullptr) o[0] = eqv2_silu(__ldg(gate + r * g_rs + c));
      __half* sh = stage + r_local * (Kr * C) + c;
#pragma unroll
      for (int p = 0; p < Kr; ++p) {
        const float v = o[p] * pscale;
        const __half h = __float2half_rn(v);
        sh[p * C] = h;
        sh[Kr * S2_THREADS + p * C] = __float2half_rn(v - __half2float(h));
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long rows_here = (R - r0) < (long long)(S2_THREADS / C) ? (R - r0) : (long long)(S2_THREADS / C);
      __half* hp = reinterpret_cast<__half*>(PA.hi);
      for (long long rl = 0; rl < rows_here; ++rl) {
        eqv2_bulk_s2g(hp + (r0 + rl) * PA.ld, stage + rl * (Kr * C), (unsigned)(Kr * C * 2));
        eqv2_bulk_s2g(hp + (r0 + rl) * PA.ld + PA.plane, stage + Kr * S2_THREADS + rl * (Kr * C), (unsigned)(Kr * C * 2));
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the staged tile is rewritten next iteration
    }
    __syncthreads();
  }
}
#endif

template <int L, int M, bool MP, int MB>
__global__ void __launch_bounds__(S2_THREADS, MB)
s2sep_bwd_kernel(const float* __restrict__ X, long long x_rs, const float* __restrict__ gate, long long g_rs,
                 const float* __restrict__ dO, long long o_rs, float* __restrict__ dX, long long dx_rs,
                 float* __restrict__ dgate, long long dg_rs, long long R, int C, int slot, float* __restrict__ absmax) {
  constexpr int Kr = KrOf<L, M>::value();
  const S2Tables& T = g_tab[slot];
  const long long total = R * C;
  float amax = 0.f;          // max |dX|, |dgate| of what this thread writes
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const long long r = w / C;
    const int c = (int)(w % C);
    float x[Kr], go[Kr], dx[Kr];
#pragma unroll
    for (int p = 0; p < Kr; ++p) {
      x[p] = __ldg(X + r * x_rs + (long long)p * C + c);
      // the l = 0 output row is overwritten by the gate path -> no gradient through the grid
      go[p] = (p > 0 || gate == nullptr) ? __ldg(dO + r * o_rs + (long long)p * C + c) : 0.f;
      dx[p] = 0.f;
    }
#pragma unroll 1
    for (int b = 0; b < S2_RES; ++b) {
      float up[M + 1], un[M + 1], hp[M + 1], hn[M + 1], wp[M + 1], wn[M + 1];
      lat_fwd<L, M, MP>(x, T.Pt, b, up, un);     // grid values of x on ring b
      lat_fwd<L, M, MP>(go, T.Pf, b, hp, hn);    // from_grid^T applied to dO on ring b
#pragma unroll
      for (int mi = 0; mi <= M; ++mi) wp[mi] = wn[mi] = 0.f;
#pragma unroll
      for (int a = 0; a < S2_RES; ++a) {
        float g = up[0], h = hp[0];
#pragma unroll
        for (int mi = 1; mi <= M; ++mi) {
          g = fmaf(T.ct[a][mi], up[mi], fmaf(T.st[a][mi], un[mi], g));
          h = fmaf(T.ct[a][mi], hp[mi], fmaf(T.st[a][mi], hn[mi], h));
        }
        const float t = h * eqv2_dsilu(g);
        wp[0] += t;
#pragma unroll
        for (int mi = 1; mi <= M; ++mi) {
          wp[mi] = fmaf(T.ct[a][mi], t, wp[mi]);
          wn[mi] = fmaf(T.st[a][mi], t, wn[mi]);
        }
      }
      lat_bwd<L, M, MP>(dx, T.Pt, b, wp, wn);
    }
#pragma unroll
    for (int p = 0; p < Kr; ++p) {
      dX[r * dx_rs + (long long)p * C + c] = dx[p];
      amax = fmaxf(amax, fabsf(dx[p]));
    }
    if (gate != nullptr) {
      const float gv = __ldg(gate + r * g_rs + c);
      const float dg = __ldg(dO + r * o_rs + c) * eqv2_dsilu(gv);
      dgate[r * dg_rs + c] = dg;
      amax = fmaxf(amax, fabsf(dg));
    }
  }
  if (absmax != nullptr) eqv2_commit_absmax(amax, absmax);
}

// d/dz of SiLU'(z):  s (1 - s) (2 + z (1 - 2 s))
__device__ __forceinline__ float eqv2_d2silu(float z) {
  const float sg = eqv2_sigmoid(z);
  return sg * (1.f - sg) * (2.f + z * (1.f - 2.f * sg));
}

// Derivative of the BACKWARD pass (needed when forces = -dE/dpos are themselves trained on).
// Backward computed  dx = T^t[ SiLU'(T x) . (F go') ],  dgate = go_0 SiLU'(gate)   (go' = go without row 0 if gated).
// With cotangents u (of dx) and w (of dgate):   S = <u, dx> + <w, dgate>
//   dS/dx    = T^t[ (T u) . SiLU''(T x) . (F go') ]
//   dS/dgo'  = F^t[ (T u) . SiLU'(T x) ]             dS/dgo_0 = w SiLU'(gate)      (gated)
//   dS/dgate = w go_0 SiLU''(gate)
template <int L, int M, bool MP>
__global__ void __launch_bounds__(S2_THREADS, 3)
s2sep_bwd2_kernel(const float* __restrict__ X, long long x_rs, const float* __restrict__ gate, long long g_rs,
                  const float* __restrict__ dO, long long o_rs, const float* __restrict__ U, long long u_rs,
                  const float* __restrict__ Wg, long long w_rs, float* __restrict__ d2X, long long d2x_rs,
                  float* __restrict__ d2gate, long long d2g_rs, float* __restrict__ d2O, long long d2o_rs, long long R,
                  int C, int slot) {
  constexpr int Kr = KrOf<L, M>::value();
  const S2Tables& T = g_tab[slot];
  const long long total = R * C;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const long long r = w / C;
    const int c = (int)(w % C);
    float x[Kr], go[Kr], u[Kr], ax[Kr], ao[Kr];
#pragma unroll
    for (int p = 0; p < Kr; ++p) {
      x[p] = __ldg(X + r * x_rs + (long long)p * C + c);
      go[p] = (p > 0 || gate == nullptr) ? __ldg(dO + r * o_rs + (long long)p * C + c) : 0.f;
      u[p] = (U != nullptr) ? __ldg(U + r * u_rs + (long long)p * C + c) : 0.f;
      ax[p] = 0.f;
      ao[p] = 0.f;
    }
#pragma unroll 1
    for (int b = 0; b < S2_RES; ++b) {
      float xp[M + 1], xn[M + 1], hp[M + 1], hn[M + 1], qp[M + 1], qn[M + 1], ap[M + 1], an[M + 1], bp[M + 1], bn[M + 1];
      lat_fwd<L, M, MP>(x, T.Pt, b, xp, xn);
      lat_fwd<L, M, MP>(go, T.Pf, b, hp, hn);
      lat_fwd<L, M, MP>(u, T.Pt, b, qp, qn);
#pragma unroll
      for (int mi = 0; mi <= M; ++mi) ap[mi] = an[mi] = bp[mi] = bn[mi] = 0.f;
#pragma unroll
      for (int a = 0; a < S2_RES; ++a) {
        float g = xp[0], h = hp[0], q = qp[0];
#pragma unroll
        for (int mi = 1; mi <= M; ++mi) {
          g = fmaf(T.ct[a][mi], xp[mi], fmaf(T.st[a][mi], xn[mi], g));
          h = fmaf(T.ct[a][mi], hp[mi], fmaf(T.st[a][mi], hn[mi], h));
          q = fmaf(T.ct[a][mi], qp[mi], fmaf(T.st[a][mi], qn[mi], q));
        }
        const float t1 = q * h * eqv2_d2silu(g);
        const float t2 = q * eqv2_dsilu(g);
        ap[0] += t1;
        bp[0] += t2;
#pragma unroll
        for (int mi = 1; mi <= M; ++mi) {
          ap[mi] = fmaf(T.ct[a][mi], t1, ap[mi]);
          an[mi] = fmaf(T.st[a][mi], t1, an[mi]);
          bp[mi] = fmaf(T.ct[a][mi], t2, bp[mi]);
          bn[mi] = fmaf(T.st[a][mi], t2, bn[mi]);
        }
      }
      lat_bwd<L, M, MP>(ax, T.Pt, b, ap, an);
      lat_bwd<L, M, MP>(ao, T.Pf, b, bp, bn);
    }
    if (gate != nullptr) {
      const float gv = __ldg(gate + r * g_rs + c);
      const float wv = (Wg != nullptr) ? __ldg(Wg + r * w_rs + c) : 0.f;
      ao[0] = wv * eqv2_dsilu(gv);
      d2gate[r * d2g_rs + c] = wv * __ldg(dO + r * o_rs + c) * eqv2_d2silu(gv);
    }
#pragma unroll
    for (int p = 0; p < Kr; ++p) {
      d2X[r * d2x_rs + (long long)p * C + c] = ax[p];
      d2O[r * d2o_rs + (long long)p * C + c] = ao[p];
    }
  }
}

inline unsigned s2_grid_blocks(long long total) {
  long long b = (total + S2_THREADS - 1) / S2_THREADS;
  const long long cap = 148LL * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

// EQV2_S2_MB=1 selects the unconstrained-register instances of the forward / backward kernels (A/B measurements)
static int s2_min_blocks() {
  const char* v = getenv("EQV2_S2_MB");
  return (v != nullptr && v[0] == '1') ? 1 : 3;
}
#define S2_PICK(K, L_, M_)                                                                                               \
  (s2_min_blocks() == 3 ? (m_primary ? K<L_, M_, true, 3> : K<L_, M_, false, 3>)                                         \
                        : (m_primary ? K<L_, M_, true, 1> : K<L_, M_, false, 1>))

#define EQV2_S2_CONFIGS(X) X(2, 2) X(3, 2) X(3, 3) X(4, 2) X(4, 4) X(6, 2) X(6, 6) X(1, 1) X(2, 1) X(5, 2) X(5, 5)

extern "C" int eqv2_s2sep_supported(int lmax, int mmax) {
#define X(L_, M_) if (lmax == L_ && mmax == M_) return 1;
  EQV2_S2_CONFIGS(X)
#undef X
  return 0;
}

// tables: host pointer to an S2Tables-shaped float block (see ops.py::S2Factors), copied into constant slot 0/1
extern "C" int eqv2_s2sep_set_tables(const float* host_tables, int nfloats, int slot, void* stream) {
  EQV2_REQUIRE(slot >= 0 && slot < S2_SLOTS, "s2sep_set_tables: slot must be in [0, %d)", S2_SLOTS);
  EQV2_REQUIRE(nfloats == (int)(sizeof(S2Tables) / sizeof(float)), "s2sep_set_tables: expected %d floats, got %d",
               (int)(sizeof(S2Tables) / sizeof(float)), nfloats);
#ifdef EQV2_CPU_EMU
  memcpy(&g_tab[slot], host_tables, sizeof(S2Tables));
#else
  cudaError_t e = cudaMemcpyToSymbolAsync(g_tab, host_tables, sizeof(S2Tables), (size_t)slot * sizeof(S2Tables),
                                          cudaMemcpyHostToDevice, (cudaStream_t)stream);
  EQV2_REQUIRE(e == cudaSuccess, "s2sep_set_tables: %s", cudaGetErrorString(e));
#endif
  return 0;
}

extern "C" int eqv2_s2sep_fwd(const float* Xp, long long x_rs, const float* gate, long long g_rs, float* O, long long o_rs,
                              long long R, int C, int lmax, int mmax, int m_primary, int slot, float* absmax,
                              void* stream) {
  if (R == 0) return 0;
  const unsigned blocks = s2_grid_blocks(R * C);
#define X(L_, M_)                                                                                                        \
  if (lmax == L_ && mmax == M_) {                                                                                        \
    auto kfn = S2_PICK(s2sep_fwd_kernel, L_, M_);                                                                        \
    EQV2_LAUNCH(kfn, dim3(blocks), dim3(S2_THREADS), 0, stream, Xp, x_rs, gate, g_rs, O, o_rs, R, C, slot, absmax);       \
    EQV2_CHECK_LAUNCH("eqv2_s2sep_fwd");                                                                                 \
    return 0;                                                                                                            \
  }
  EQV2_S2_CONFIGS(X)
#undef X
  eqv2_set_error("s2sep_fwd: (lmax, mmax) = (%d, %d) not instantiated", lmax, mmax);
  return 1;
}

#ifndef EQV2_CPU_EMU
extern "C" int eqv2_s2sep_fwd_planes(const float* Xp, long long x_rs, const float* gate, long long g_rs, void* planes,
                                     long long plane, long long ld, const float* bound_in, float bound_c, float* bound_out,
                                     long long R, int C, int lmax, int mmax, int m_primary, int slot, void* stream) {
  if (R == 0) return 0;
  EQV2_REQUIRE(planes != nullptr && bound_in != nullptr && bound_out != nullptr && bound_c >= 1.0f,
               "s2sep_fwd_planes: null plane / bound pointer or bound factor < 1");
  EQV2_REQUIRE(C > 0 && C <= S2_THREADS && S2_THREADS % C == 0 && C % 8 == 0,
               "s2sep_fwd_planes: C = %d must divide %d and be a multiple of 8", C, S2_THREADS);
  EQV2_REQUIRE((ld % 8) == 0 && (plane % 8) == 0 && (((uintptr_t)planes) & 15) == 0,
               "s2sep_fwd_planes: planes must be 16-byte aligned with ld %% 8 == 0");
  const Eqv2PlaneArgs PA{planes, plane, ld, bound_in, nullptr, bound_c, bound_out};
  const unsigned blocks = s2_grid_blocks(R * C);
#define X(L_, M_)                                                                                                        \
  if (lmax == L_ && mmax == M_) {                                                                                        \
    auto kfn = S2_PICK(s2sep_fwd_planes_kernel, L_, M_);                                                                 \
    constexpr int kr_ = KrOf<L_, M_>::value();                                                                           \
    const size_t smem = (size_t)2 * kr_ * S2_THREADS * sizeof(__half);                                                   \
    EQV2_REQUIRE(ld >= (long long)kr_ * C, "s2sep_fwd_planes: ld smaller than Kr * C");                                  \
    EQV2_LAUNCH(kfn, dim3(blocks), dim3(S2_THREADS), smem, stream, Xp, x_rs, gate, g_rs, R, C, slot, PA);                 \
    EQV2_CHECK_LAUNCH("eqv2_s2sep_fwd_planes");                                                                          \
    return 0;                                                                                                            \
  }
  EQV2_S2_CONFIGS(X)
#undef X
  eqv2_set_error("s2sep_fwd_planes: (lmax, mmax) = (%d, %d) not instantiated", lmax, mmax);
  return 1;
}
#endif

extern "C" int eqv2_s2sep_bwd(const float* Xp, long long x_rs, const float* gate, long long g_rs, const float* dO,
                              long long o_rs, float* dX, long long dx_rs, float* dgate, long long dg_rs, long long R, int C,
                              int lmax, int mmax, int m_primary, int slot, float* absmax, void* stream) {
  if (R == 0) return 0;
  const unsigned blocks = s2_grid_blocks(R * C);
#define X(L_, M_)                                                                                                        \
  if (lmax == L_ && mmax == M_) {                                                                                        \
    auto kfn = S2_PICK(s2sep_bwd_kernel, L_, M_);                                                                        \
    EQV2_LAUNCH(kfn, dim3(blocks), dim3(S2_THREADS), 0, stream, Xp, x_rs, gate, g_rs, dO, o_rs, dX, dx_rs, dgate, dg_rs, R, C, slot, absmax); \
    EQV2_CHECK_LAUNCH("eqv2_s2sep_bwd");                                                                                 \
    return 0;                                                                                                            \
  }
  EQV2_S2_CONFIGS(X)
#undef X
  eqv2_set_error("s2sep_bwd: (lmax, mmax) = (%d, %d) not instantiated", lmax, mmax);
  return 1;
}

extern "C" int eqv2_s2sep_bwd2(const float* Xp, long long x_rs, const float* gate, long long g_rs, const float* dO,
                               long long o_rs, const float* U, long long u_rs, const float* Wg, long long w_rs,
                               float* d2X, long long d2x_rs, float* d2gate, long long d2g_rs, float* d2O, long long d2o_rs,
                               long long R, int C, int lmax, int mmax, int m_primary, int slot, void* stream) {
  if (R == 0) return 0;
  const unsigned blocks = s2_grid_blocks(R * C);
#define X(L_, M_)                                                                                                        \
  if (lmax == L_ && mmax == M_) {                                                                                        \
    auto kfn = m_primary ? s2sep_bwd2_kernel<L_, M_, true> : s2sep_bwd2_kernel<L_, M_, false>;                           \
    EQV2_LAUNCH(kfn, dim3(blocks), dim3(S2_THREADS), 0, stream, Xp, x_rs, gate, g_rs, dO, o_rs, U, u_rs, Wg, w_rs, d2X,  \
                d2x_rs, d2gate, d2g_rs, d2O, d2o_rs, R, C, slot);                                                        \
    EQV2_CHECK_LAUNCH("eqv2_s2sep_bwd2");                                                                                \
    return 0;                                                                                                            \
  }
  EQV2_S2_CONFIGS(X)
#undef X
  eqv2_set_error("s2sep_bwd2: (lmax, mmax) = (%d, %d) not instantiated", lmax, mmax);
  return 1;
}
