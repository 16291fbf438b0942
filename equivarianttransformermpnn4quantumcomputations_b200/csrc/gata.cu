// Per-edge operators of the GATA / HTR forks (BASELINE configs 4-5), one thread per (edge, channel):
//
//  HTR (NewFunctions/Gotennet_morethaninspired/activation.py:166-264): the scalar the edge stream is updated with is
//      w[e,c] = sum_{l>=1} <reject(q^l, r^l), reject(k^l, -r^l)> / (2l+1)
//             = sum_l [ q^l.k^l - (2 - |r^l|^2) (q^l.r^l)(k^l.r^l) ] / (2l+1)              =: T(q, k; r)
//    a BILINEAR form of the two projected feature tensors q, k [E, M, H] (M = (lmax+1)^2 - 1 rows, l = 1..lmax) with
//    coefficients from the detached edge harmonics r [E, M].  Its gradient map
//      G(g, b; r)[e,m,c] = g[e,c] / (2l+1) * ( b[e,m,c] - (2 - |r^l|^2) (b^l.r^l)[e,c] r[e,m] )
//    satisfies  <u, G(g, b)> = g . T(u, b),  so {T, G} is closed under differentiation: dT/dq = G(g, k), dT/dk = G(g, q),
//    dG/dg = T(u, b), dG/db = G(g, u).  Forces by autograd (double backward) therefore need exactly these two kernels.
//    They replace ~10 element-wise / reduction launches per degree and pass over [E, 2l+1, H] slices.
//
//  GATAValueActivation (:270-414): combined [E, (1 + 2 lmax) H] = (o_s | o_d^l | o_t^l), Xp = xj_proj X_j [E, M, H]:
//      out[e, 0, c]        = SiLU(o_s[e,c])
//      out[e, row(l,m), c] = o_d^l[e,c] r[e, off_l + m] + o_t^l[e,c] Xp[e, off_l + m, c],   m < w_l = min(2l+1, 2 mmax+1)
//    (the FIRST w_l rows of each degree, reference :389-393), rows in l-primary reduced order.  Forward, backward and the
//    derivative of the backward (SiLU'' on the scalar row, products of cotangents elsewhere) are one kernel each.
#include "common.cuh"

namespace {

constexpr int GATA_MAXL = EQV2_MAX_LMAX;

__global__ void htr_inner_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ rl,
                                 float* __restrict__ out, long long E, int H, int lmax) {
  const long long e = blockIdx.x;
  const int M = (lmax + 1) * (lmax + 1) - 1;
  const float* re = rl + e * M;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float* qp = q + (e * M) * (long long)H + c;
    const float* kp = k + (e * M) * (long long)H + c;
    float w = 0.f;
    for (int l = 1; l <= lmax; ++l) {
      const int n = 2 * l + 1, off = l * l - 1;
      float qk = 0.f, qr = 0.f, kr = 0.f, rr = 0.f;
      for (int m = 0; m < n; ++m) {
        const float r = re[off + m];
        const float qv = qp[(long long)(off + m) * H], kv = kp[(long long)(off + m) * H];
        qk = fmaf(qv, kv, qk);
        qr = fmaf(qv, r, qr);
        kr = fmaf(kv, r, kr);
        rr = fmaf(r, r, rr);
      }
      w += (qk - (2.0f - rr) * qr * kr) / (float)n;
    }
    out[e * H + c] = w;
  }
}

__global__ void htr_grad_kernel(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ rl,
                                float* __restrict__ out, long long E, int H, int lmax) {
  const long long e = blockIdx.x;
  const int M = (lmax + 1) * (lmax + 1) - 1;
  const float* re = rl + e * M;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float* bp = b + (e * M) * (long long)H + c;
    float* op = out + (e * M) * (long long)H + c;
    const float gv = g[e * H + c];
    for (int l = 1; l <= lmax; ++l) {
      const int n = 2 * l + 1, off = l * l - 1;
      float bv[2 * GATA_MAXL + 1];
      float br = 0.f, rr = 0.f;
#pragma unroll
      for (int m = 0; m < 2 * GATA_MAXL + 1; ++m) {
        if (m < n) {
          const float r = re[off + m];
          bv[m] = bp[(long long)(off + m) * H];
          br = fmaf(bv[m], r, br);
          rr = fmaf(r, r, rr);
        }
      }
      const float s = gv / (float)n, t = (2.0f - rr) * br;
#pragma unroll
      for (int m = 0; m < 2 * GATA_MAXL + 1; ++m)
        if (m < n) op[(long long)(off + m) * H] = s * (bv[m] - t * re[off + m]);
    }
  }
}

// d2/dx2 silu(x) = s (1 - s) (2 + x (1 - 2 s))
__device__ __forceinline__ float gata_d2silu(float x) {
  const float s = eqv2_sigmoid(x);
  return s * (1.0f - s) * (2.0f + x * (1.0f - 2.0f * s));
}

// row offset of degree l >= 1 inside the l-primary reduced output ([0] = scalar row)
__device__ __forceinline__ int gata_row0(int l, int mmax) {
  int r = 1;
  for (int j = 1; j < l; ++j) r += (2 * j + 1 < 2 * mmax + 1) ? 2 * j + 1 : 2 * mmax + 1;
  return r;
}

__global__ void gata_value_fwd_kernel(const float* __restrict__ comb, const float* __restrict__ Xp,
                                      const float* __restrict__ rl, float* __restrict__ out, long long E, int H, int lmax,
                                      int mmax, int Kr) {
  const long long e = blockIdx.x;
  const int M = (lmax + 1) * (lmax + 1) - 1, S = 1 + 2 * lmax;
  const float* re = rl + e * M;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float* cp = comb + e * (long long)S * H + c;
    const float* xp = Xp + (e * M) * (long long)H + c;
    float* op = out + (e * Kr) * (long long)H + c;
    op[0] = eqv2_silu(cp[0]);
    for (int l = 1; l <= lmax; ++l) {
      const int n = 2 * l + 1, w = n < 2 * mmax + 1 ? n : 2 * mmax + 1, off = l * l - 1, r0 = gata_row0(l, mmax);
      const float od = cp[(long long)l * H], ot = cp[(long long)(lmax + l) * H];
      for (int m = 0; m < w; ++m)
        op[(long long)(r0 + m) * H] = fmaf(od, re[off + m], ot * xp[(long long)(off + m) * H]);
    }
  }
}

// d_comb [E, S H], d_Xp [E, M, H] from g [E, Kr, H]
__global__ void gata_value_bwd_kernel(const float* __restrict__ comb, const float* __restrict__ Xp,
                                      const float* __restrict__ rl, const float* __restrict__ g, float* __restrict__ dcomb,
                                      float* __restrict__ dXp, long long E, int H, int lmax, int mmax, int Kr) {
  const long long e = blockIdx.x;
  const int M = (lmax + 1) * (lmax + 1) - 1, S = 1 + 2 * lmax;
  const float* re = rl + e * M;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float* cp = comb + e * (long long)S * H + c;
    const float* xp = Xp + (e * M) * (long long)H + c;
    const float* gp = g + (e * Kr) * (long long)H + c;
    float* dc = dcomb + e * (long long)S * H + c;
    float* dx = dXp + (e * M) * (long long)H + c;
    dc[0] = gp[0] * eqv2_dsilu(cp[0]);
    for (int l = 1; l <= lmax; ++l) {
      const int n = 2 * l + 1, w = n < 2 * mmax + 1 ? n : 2 * mmax + 1, off = l * l - 1, r0 = gata_row0(l, mmax);
      const float ot = cp[(long long)(lmax + l) * H];
      float sd = 0.f, st = 0.f;
      for (int m = 0; m < n; ++m) {
        float d = 0.f;
        if (m < w) {
          const float gv = gp[(long long)(r0 + m) * H];
          sd = fmaf(gv, re[off + m], sd);
          st = fmaf(gv, xp[(long long)(off + m) * H], st);
          d = gv * ot;
        }
        dx[(long long)(off + m) * H] = d;
      }
      dc[(long long)l * H] = sd;
      dc[(long long)(lmax + l) * H] = st;
    }
  }
}

// derivative of the backward: cotangents u [E, S H] (of d_comb), v [E, M, H] (of d_Xp) -> d_g [E,Kr,H], d2comb, d2Xp
__global__ void gata_value_bwd2_kernel(const float* __restrict__ comb, const float* __restrict__ Xp,
                                       const float* __restrict__ rl, const float* __restrict__ g,
                                       const float* __restrict__ u, const float* __restrict__ v, float* __restrict__ dg,
                                       float* __restrict__ d2comb, float* __restrict__ d2Xp, long long E, int H, int lmax,
                                       int mmax, int Kr) {
  const long long e = blockIdx.x;
  const int M = (lmax + 1) * (lmax + 1) - 1, S = 1 + 2 * lmax;
  const float* re = rl + e * M;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float* cp = comb + e * (long long)S * H + c;
    const float* xp = Xp + (e * M) * (long long)H + c;
    const float* gp = g + (e * Kr) * (long long)H + c;
    const float* up = u != nullptr ? u + e * (long long)S * H + c : nullptr;
    const float* vp = v != nullptr ? v + (e * M) * (long long)H + c : nullptr;
    float* dgp = dg + (e * Kr) * (long long)H + c;
    float* d2c = d2comb + e * (long long)S * H + c;
    float* d2x = d2Xp + (e * M) * (long long)H + c;
    const float u0 = up ? up[0] : 0.f;
    dgp[0] = u0 * eqv2_dsilu(cp[0]);
    d2c[0] = u0 * gp[0] * gata_d2silu(cp[0]);
    for (int l = 1; l <= lmax; ++l) {
      const int n = 2 * l + 1, w = n < 2 * mmax + 1 ? n : 2 * mmax + 1, off = l * l - 1, r0 = gata_row0(l, mmax);
      const float ot = cp[(long long)(lmax + l) * H];
      const float ud = up ? up[(long long)l * H] : 0.f, ut = up ? up[(long long)(lmax + l) * H] : 0.f;
      float s_ot = 0.f;
      for (int m = 0; m < n; ++m) {
        float dx = 0.f;
        if (m < w) {
          const float gv = gp[(long long)(r0 + m) * H];
          const float vv = vp ? vp[(long long)(off + m) * H] : 0.f;
          dgp[(long long)(r0 + m) * H] = fmaf(ud, re[off + m], fmaf(ut, xp[(long long)(off + m) * H], vv * ot));
          s_ot = fmaf(vv, gv, s_ot);
          dx = ut * gv;
        }
        d2x[(long long)(off + m) * H] = dx;
      }
      d2c[(long long)l * H] = 0.f;
      d2c[(long long)(lmax + l) * H] = s_ot;
    }
  }
}

inline int gata_threads(int H) { return H >= 128 ? 128 : (H >= 64 ? 64 : 32); }

}  // namespace

extern "C" int eqv2_htr_inner(const float* q, const float* k, const float* rl, float* out, long long E, int H, int lmax,
                              void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(H > 0 && lmax >= 1 && lmax <= GATA_MAXL, "htr_inner: bad sizes");
  EQV2_LAUNCH(htr_inner_kernel, dim3((unsigned)E), dim3(gata_threads(H)), 0, stream, q, k, rl, out, E, H, lmax);
  EQV2_CHECK_LAUNCH("eqv2_htr_inner");
  return 0;
}

extern "C" int eqv2_htr_grad(const float* g, const float* b, const float* rl, float* out, long long E, int H, int lmax,
                             void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(H > 0 && lmax >= 1 && lmax <= GATA_MAXL, "htr_grad: bad sizes");
  EQV2_LAUNCH(htr_grad_kernel, dim3((unsigned)E), dim3(gata_threads(H)), 0, stream, g, b, rl, out, E, H, lmax);
  EQV2_CHECK_LAUNCH("eqv2_htr_grad");
  return 0;
}

extern "C" int eqv2_gata_value_fwd(const float* comb, const float* Xp, const float* rl, float* out, long long E, int H,
                                   int lmax, int mmax, int Kr, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(H > 0 && lmax >= 1 && lmax <= GATA_MAXL && mmax >= 0, "gata_value_fwd: bad sizes");
  EQV2_LAUNCH(gata_value_fwd_kernel, dim3((unsigned)E), dim3(gata_threads(H)), 0, stream, comb, Xp, rl, out, E, H, lmax, mmax, Kr);
  EQV2_CHECK_LAUNCH("eqv2_gata_value_fwd");
  return 0;
}

extern "C" int eqv2_gata_value_bwd(const float* comb, const float* Xp, const float* rl, const float* g, float* dcomb,
                                   float* dXp, long long E, int H, int lmax, int mmax, int Kr, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(H > 0 && lmax >= 1 && lmax <= GATA_MAXL && mmax >= 0, "gata_value_bwd: bad sizes");
  EQV2_LAUNCH(gata_value_bwd_kernel, dim3((unsigned)E), dim3(gata_threads(H)), 0, stream, comb, Xp, rl, g, dcomb, dXp, E, H, lmax, mmax, Kr);
  EQV2_CHECK_LAUNCH("eqv2_gata_value_bwd");
  return 0;
}

extern "C" int eqv2_gata_value_bwd2(const float* comb, const float* Xp, const float* rl, const float* g, const float* u,
                                    const float* v, float* dg, float* d2comb, float* d2Xp, long long E, int H, int lmax,
                                    int mmax, int Kr, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(H > 0 && lmax >= 1 && lmax <= GATA_MAXL && mmax >= 0, "gata_value_bwd2: bad sizes");
  EQV2_LAUNCH(gata_value_bwd2_kernel, dim3((unsigned)E), dim3(gata_threads(H)), 0, stream, comb, Xp, rl, g, u, v, dg, d2comb, d2Xp, E, H, lmax, mmax, Kr);
  EQV2_CHECK_LAUNCH("eqv2_gata_value_bwd2");
  return 0;
}
