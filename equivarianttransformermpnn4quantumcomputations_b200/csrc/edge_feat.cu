// Per-edge scalar-feature kernels:
//   * Gaussian smearing of edge distances (equiformerv2_oc20.py:43-60) fwd / bwd-wrt-distance,
//   * fused LayerNorm + SiLU of the radial MLP (radial_function.py:21-22) fwd / bwd,
//   * Euler angles + Wigner-D blocks of the edge frames (so3.py:525-545, wigner.py:17-39),
//     written block-diagonal [E, sum_l (2l+1)^2] instead of dense [E,K,K].
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------- RBF
__global__ void rbf_fwd_kernel(const float* __restrict__ d, float* __restrict__ out, long long E, int R,
                               const float* __restrict__ offset, float coeff) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E * R) return;
  const long long e = i / R;
  const int k = (int)(i % R);
  const float t = d[e] - offset[k];
  out[i] = expf(coeff * t * t);
}

// dd[e] = sum_k go[e,k] * 2 coeff (d - mu_k) exp(coeff (d - mu_k)^2)
__global__ void rbf_bwd_kernel(const float* __restrict__ d, const float* __restrict__ go, float* __restrict__ dd,
                               long long E, int R, const float* __restrict__ offset, float coeff) {
  const int lane = threadIdx.x & 31;
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= E) return;   // whole warp exits together
  const float de = d[e];
  float s = 0.f;
  for (int k = lane; k < R; k += 32) {
    const float t = de - offset[k];
    s = fmaf(go[e * R + k], 2.f * coeff * t * expf(coeff * t * t), s);
  }
  s = eqv2_warp_sum(s);
  if (lane == 0) dd[e] = s;
}

// ---------------------------------------------------------------------------------------- LN + SiLU
constexpr int LN_PL = 8;  // features per lane (width <= 256)

__global__ void ln_silu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ b, float* __restrict__ y, long long rows, int width,
                                   float eps) {
  const int lane = threadIdx.x & 31;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const float* xp = x + r * width;
  float v[LN_PL];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LN_PL; ++k) {
    const int i = lane + 32 * k;
    v[k] = (i < width) ? xp[i] : 0.f;
    s += v[k];
  }
  const float mean = eqv2_warp_sum(s) / width;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < LN_PL; ++k) {
    const int i = lane + 32 * k;
    const float t = (i < width) ? v[k] - mean : 0.f;
    q = fmaf(t, t, q);
  }
  const float rstd = rsqrtf(eqv2_warp_sum(q) / width + eps);
#pragma unroll
  for (int k = 0; k < LN_PL; ++k) {
    const int i = lane + 32 * k;
    if (i < width) y[r * width + i] = eqv2_silu((v[k] - mean) * rstd * w[i] + b[i]);
  }
}

__global__ void ln_silu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ b, const float* __restrict__ gy, float* __restrict__ gx,
                                   float* __restrict__ gw, float* __restrict__ gb, long long rows, int width,
                                   float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float aw[LN_PL], ab[LN_PL];
#pragma unroll
  for (int k = 0; k < LN_PL; ++k) aw[k] = ab[k] = 0.f;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* xp = x + r * width;
    float v[LN_PL];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_PL; ++k) {
      const int i = lane + 32 * k;
      v[k] = (i < width) ? xp[i] : 0.f;
      s += v[k];
    }
    const float mean = eqv2_warp_sum(s) / width;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_PL; ++k) {
      const int i = lane + 32 * k;
      const float t = (i < width) ? v[k] - mean : 0.f;
      q = fmaf(t, t, q);
    }
    const float rstd = rsqrtf(eqv2_warp_sum(q) / width + eps);
    float dxh[LN_PL], xh[LN_PL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_PL; ++k) {
      const int i = lane + 32 * k;
      dxh[k] = xh[k] = 0.f;
      if (i < width) {
        xh[k] = (v[k] - mean) * rstd;
        const float u = xh[k] * w[i] + b[i];
        const float du = gy[r * width + i] * eqv2_dsilu(u);
        aw[k] = fmaf(du, xh[k], aw[k]);
        ab[k] += du;
        dxh[k] = du * w[i];
        s1 += dxh[k];
        s2 = fmaf(dxh[k], xh[k], s2);
      }
    }
    s1 = eqv2_warp_sum(s1) / width;
    s2 = eqv2_warp_sum(s2) / width;
#pragma unroll
    for (int k = 0; k < LN_PL; ++k) {
      const int i = lane + 32 * k;
      if (i < width) gx[r * width + i] = rstd * (dxh[k] - s1 - xh[k] * s2);
    }
  }
  // block-level sum of the 8 warps' partial weight / bias gradients first: one atomic per (block, column) instead of
  // one per (warp, column) -- same-address atomics serialise in L2
  __shared__ float sred[2][8][32 * LN_PL];
  const int wib = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < LN_PL; ++k) {
    sred[0][wib][lane + 32 * k] = aw[k];
    sred[1][wib][lane + 32 * k] = ab[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    float sw_ = 0.f, sb_ = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      sw_ += sred[0][q][i];
      sb_ += sred[1][q][i];
    }
    atomicAdd(&gw[i], sw_);
    atomicAdd(&gb[i], sb_);
  }
}

// ---------------------------------------------------------------------------------------- Wigner-D
constexpr int WIG_WARPS = 4;
constexpr int MAXN = 2 * EQV2_MAX_LMAX + 1;

__global__ void wigner_kernel(const float* __restrict__ rot, const float* __restrict__ Jd, float* __restrict__ wig,
                              long long E, int lmax, int WS) {
  EQV2_DYN_SMEM(float, smem);
  float* sJ = smem;                                           // [WS]
  float* sA = sJ + WS;                                        // [WIG_WARPS][MAXN*MAXN]
  float* sT = sA + WIG_WARPS * MAXN * MAXN;                   // [WIG_WARPS][3][2][MAXN]  (angle, sin/cos, f + lmax)
  for (int i = threadIdx.x; i < WS; i += blockDim.x) sJ[i] = Jd[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long e = (long long)blockIdx.x * WIG_WARPS + wib;
  if (e >= E) return;   // warp-uniform; no block barriers below
  float* A = sA + wib * MAXN * MAXN;
  float* T = sT + wib * 6 * MAXN;
  const float* R = rot + e * 9;
  // x = R @ (0,1,0) -> column 1, normalised and clamped (e3nn xyz_to_angles)
  // (a matmul accumulates from +0: a degenerate frame's -0 entries become +0, which decides atan2(0, +-0);
  //  zero-length self-image edges of the MatPES v2 builder hit exactly this case)
  float x0 = __fadd_rn(R[1], 0.f), x1 = __fadd_rn(R[4], 0.f), x2 = __fadd_rn(R[7], 0.f);
  const float nrm = fmaxf(sqrtf(x0 * x0 + x1 * x1 + x2 * x2), 1e-12f);
  x0 = fminf(fmaxf(x0 / nrm, -1.f), 1.f);
  x1 = fminf(fmaxf(x1 / nrm, -1.f), 1.f);
  x2 = fminf(fmaxf(x2 / nrm, -1.f), 1.f);
  const float beta = acosf(x1);
  const float alpha = atan2f(x0, x2);
  const float ca = cosf(alpha), sa = sinf(alpha);
  // first row of (R_y(alpha) R_x(beta))^T @ R
  const float r00 = __fadd_rn(ca * R[0] - sa * R[6], 0.f);
  const float r02 = __fadd_rn(ca * R[2] - sa * R[8], 0.f);
  const float gamma = atan2f(r02, r00);
  // trig tables for frequencies f = -lmax..lmax
  if (lane <= 2 * lmax) {
    const float f = (float)(lane - lmax);
    T[0 * MAXN + lane] = sinf(f * alpha);
    T[1 * MAXN + lane] = cosf(f * alpha);
    T[2 * MAXN + lane] = sinf(f * beta);
    T[3 * MAXN + lane] = cosf(f * beta);
    T[4 * MAXN + lane] = sinf(f * gamma);
    T[5 * MAXN + lane] = cosf(f * gamma);
  }
  __syncwarp();
  float* out = wig + e * (long long)WS;
  for (int l = 0; l <= lmax; ++l) {
    const int n = 2 * l + 1;
    const float* J = sJ + eqv2_wig_off(l);
    // A = J @ (Z(beta) @ J);   row k of Z has frequency f_k = l - k
    for (int idx = lane; idx < n * n; idx += 32) {
      const int i = idx / n, j = idx % n;
      float s = 0.f;
      for (int k = 0; k < n; ++k) {
        const int fi = (l - k) + lmax;
        const float zj = T[3 * MAXN + fi] * J[k * n + j] + ((2 * k + 1 == n) ? 0.f : T[2 * MAXN + fi] * J[(n - 1 - k) * n + j]);
        s = fmaf(J[i * n + k], zj, s);
      }
      A[idx] = s;
    }
    __syncwarp();
    for (int idx = lane; idx < n * n; idx += 32) {
      const int i = idx / n, j = idx % n;
      const int fi = (l - i) + lmax, fj = (l - j) + lmax;
      const float cai = T[1 * MAXN + fi], sai = (2 * i + 1 == n) ? 0.f : T[0 * MAXN + fi];
      const float ccj = T[5 * MAXN + fj], scj = (2 * j + 1 == n) ? 0.f : T[4 * MAXN + fj];
      const float y_i = A[i * n + j] * ccj - A[i * n + (n - 1 - j)] * scj;
      const float y_r = A[(n - 1 - i) * n + j] * ccj - A[(n - 1 - i) * n + (n - 1 - j)] * scj;
      out[eqv2_wig_off(l) + idx] = cai * y_i + sai * y_r;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------- edge SH
// Real spherical harmonics of the edge direction for l = 1..lmax ("norm" normalisation: |Y_l| = 1), as
// e3nn o3.SphericalHarmonics(normalize=False, normalization='norm') is used at
// equiformerv2_MatPES_GATAV2.py:137-140,232-241 on unit = vec / max(|vec|, 1e-8).  Basis of SURVEY App. B.1
// (polar axis y, azimuth atan2(x, z), no Condon-Shortley phase), evaluated pole-free as Cartesian polynomials:
//   A_m + i B_m = (z + i x)^m ,  Q_l^m(y) = P_l^m(y) / sin^m(beta)  (upward recurrences in y).
// One thread per edge; output [E, (lmax+1)^2 - 1].
constexpr int SH_MAXL = 6;

__global__ void edge_sh_kernel(const float* __restrict__ vec, float* __restrict__ out, long long E, int lmax) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float x = vec[3 * e], y = vec[3 * e + 1], z = vec[3 * e + 2];
  const float inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-8f);
  x *= inv; y *= inv; z *= inv;
  float A[SH_MAXL + 1], B[SH_MAXL + 1];
  A[0] = 1.f; B[0] = 0.f;
  for (int m = 1; m <= lmax; ++m) {
    A[m] = A[m - 1] * z - B[m - 1] * x;
    B[m] = B[m - 1] * z + A[m - 1] * x;
  }
  const int K1 = (lmax + 1) * (lmax + 1) - 1;
  float* o = out + e * K1;
  for (int m = 0; m <= lmax; ++m) {
    float dfact = 1.f;
    for (int k = 1; k < 2 * m; k += 2) dfact *= (float)k;
    float qm2 = 0.f, qm1 = 0.f;          // Q_{l-2}^m, Q_{l-1}^m
    for (int l = m; l <= lmax; ++l) {
      float q;
      if (l == m) q = dfact;
      else if (l == m + 1) q = (float)(2 * m + 1) * y * qm1;
      else q = ((float)(2 * l - 1) * y * qm1 - (float)(l + m - 1) * qm2) / (float)(l - m);
      qm2 = qm1; qm1 = q;
      if (l == 0) continue;
      // N_l^m * sqrt(4 pi / (2l+1)) = sqrt((l-m)! / (l+m)!)
      float ratio = 1.f;
      for (int k = l - m + 1; k <= l + m; ++k) ratio /= (float)k;
      const float n = sqrtf(ratio);
      const int base = l * l - 1 + l;    // index of (l, m=0) in the output (l = 0 dropped)
      if (m == 0) o[base] = n * q;
      else {
        o[base + m] = n * 1.41421356237309515f * q * A[m];
        o[base - m] = n * 1.41421356237309515f * q * B[m];
      }
    }
  }
}

}  // namespace

extern "C" int eqv2_rbf_fwd(const float* d, float* out, long long E, int R, const float* offset, float coeff,
                            void* stream) {
  if (E == 0) return 0;
  const long long n = E * R;
  EQV2_LAUNCH(rbf_fwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, d, out, E, R, offset, coeff);
  EQV2_CHECK_LAUNCH("eqv2_rbf_fwd");
  return 0;
}

// derivative of rbf_bwd for a cotangent u [E] of dd:  dd[e] = sum_k go[e,k] g_k(d),  g_k = 2 c (d - mu_k) rbf_k(d)
//   d(go)[e,k] = u[e] g_k(d[e]);   d(d)[e] = u[e] sum_k go[e,k] (2 c + (2 c (d - mu_k))^2) rbf_k(d[e])
__global__ void rbf_bwd2_kernel(const float* __restrict__ d, const float* __restrict__ go, const float* __restrict__ u,
                                float* __restrict__ dgo, float* __restrict__ d2d, long long E, int R,
                                const float* __restrict__ offset, float coeff) {
  const int lane = threadIdx.x & 31;
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= E) return;   // whole warp exits together
  const float de = d[e], ue = u[e];
  float acc = 0.f;
  for (int k = lane; k < R; k += 32) {
    const float t = de - offset[k];
    const float r = expf(coeff * t * t);
    const float g1 = 2.f * coeff * t;
    dgo[e * R + k] = ue * g1 * r;
    acc = fmaf(go[e * R + k], (2.f * coeff + g1 * g1) * r, acc);
  }
  acc = eqv2_warp_sum(acc);
  if (lane == 0) d2d[e] = ue * acc;
}

extern "C" int eqv2_rbf_bwd2(const float* d, const float* go, const float* u, float* dgo, float* d2d, long long E, int R,
                             const float* offset, float coeff, void* stream) {
  if (E == 0) return 0;
  EQV2_LAUNCH(rbf_bwd2_kernel, dim3((unsigned)((E * 32 + 255) / 256)), dim3(256), 0, stream, d, go, u, dgo, d2d, E, R, offset, coeff);
  EQV2_CHECK_LAUNCH("eqv2_rbf_bwd2");
  return 0;
}

extern "C" int eqv2_rbf_bwd(const float* d, const float* go, float* dd, long long E, int R, const float* offset,
                            float coeff, void* stream) {
  if (E == 0) return 0;
  EQV2_LAUNCH(rbf_bwd_kernel, dim3((unsigned)((E * 32 + 255) / 256)), dim3(256), 0, stream, d, go, dd, E, R, offset, coeff);
  EQV2_CHECK_LAUNCH("eqv2_rbf_bwd");
  return 0;
}

// Derivative of the BACKWARD pass (forces by autograd: the loss on the forces is back-propagated through the first
// backward).  First backward, per row:  x^ = (x - mu) / sigma,  z = x^ w + b,  a = w gy SiLU'(z),
//   gx = J a,   J = (I - 1 1^T / n - x^ x^^T / n) / sigma  (the symmetric Jacobian of x -> x^).
// For a cotangent u of gx:  S = <u, J a> = <J u, a> = sum_i w_i gy_i SiLU'(z_i) p_i  with  p = J u.  Then
//   dS/dgy_i = w_i SiLU'(z_i) p_i
//   dS/db_i  = c_i := w_i gy_i SiLU''(z_i) p_i                         (summed over rows)
//   dS/dw_i  = gy_i SiLU'(z_i) p_i + c_i x^_i                         (summed over rows)
//   dS/dx    = J (e + q) - S x^ / (n sigma),   e = c w  (through z),   q = -(a mean(u x^) + u mean(a x^)) / sigma  (through J)
// One warp per row, two rounds of warp reductions; parameter gradients via block partial sums + one atomic per column.
__global__ void ln_silu_bwd2_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                    const float* __restrict__ gy, const float* __restrict__ u, float* __restrict__ dx,
                                    float* __restrict__ dgy, float* __restrict__ dw, float* __restrict__ db, long long rows,
                                    int width, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float aw[LN_PL], ab[LN_PL];
#pragma unroll
  for (int k = 0; k < LN_PL; ++k) aw[k] = ab[k] = 0.f;
  const float inv_n = 1.0f / (float)width;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* xp = x + r * width;
    float v[LN_PL];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_PL; ++k) {
      const int i = lane + 32 * k;
      v[k] = (i < width) ? xp[i] : 0.f;
      s += v[k];
    }
    const float mean = eqv2_warp_sum(s) * inv_n;
    float q2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_PL; ++k) {
      const int i = lane + 32 * k;
      const float t = (i < width) ? v[k] - mean : 0.f;
      q2 = fmaf(t, t, q2);
    }
    const float rstd = rsqrtf(eqv2_warp_sum(q2) * inv_n + eps);
    float xh[LN_PL], a[LN_PL], uu[LN_PL], d1[LN_PL], cc[LN_PL];      // x^, a, u, gy SiLU'(z), w gy SiLU''(z)
    float su = 0.f, sux = 0.f, sax = 0.f;
#pragma unroll
    for (int k = 0; k < LN_PL; ++k) {
      const int i = lane + 32 * k;
      xh[k] = a[k] = uu[k] = d1[k] = cc[k] = 0.f;
      if (i < width) {
        xh[k] = (v[k] - mean) * rstd;
        const float wi = w[i];
        const float z = fmaf(xh[k], wi, b[i]);
        const float gyi = gy[r * width + i];
        d1[k] = gyi * eqv2_dsilu(z);
        cc[k] = wi * gyi * eqv2_silu_d2(z);
        a[k] = wi * d1[k];
        uu[k] = u[r * width + i];
        su += uu[k];
        sux = fmaf(uu[k], xh[k], sux);
        sax = fmaf(a[k], xh[k], sax);
      }
    }
    const float ubar = eqv2_warp_sum(su) * inv_n, ux = eqv2_warp_sum(sux) * inv_n, m2 = eqv2_warp_sum(sax) * inv_n;
    float f[LN_PL];
    float sf = 0.f, sfx = 0.f, sT = 0.f;
#pragma unroll
    for (int k = 0; k < LN_PL; ++k) {
      const int i = lane + 32 * k;
      f[k] = 0.f;
      if (i < width) {
        const float p = (uu[k] - ubar - xh[k] * ux) * rstd;
        const float c = cc[k] * p;                      // dS/dz_i
        const float wi = w[i];
        dgy[r * width + i] = wi * eqv2_dsilu(fmaf(xh[k], wi, b[i])) * p;
        ab[k] += c;
        aw[k] += fmaf(d1[k], p, c * xh[k]);
        f[k] = c * wi - rstd * (a[k] * ux + uu[k] * m2);
        sf += f[k];
        sfx = fmaf(f[k], xh[k], sfx);
        sT = fmaf(a[k], p, sT);
      }
    }
    const float mf = eqv2_warp_sum(sf) * inv_n, mfx = eqv2_warp_sum(sfx) * inv_n, T = eqv2_warp_sum(sT);
#pragma unroll
    for (int k = 0; k < LN_PL; ++k) {
      const int i = lane + 32 * k;
      if (i < width) dx[r * width + i] = rstd * (f[k] - mf - xh[k] * mfx) - T * rstd * inv_n * xh[k];
    }
  }
  __shared__ float sred[2][8][32 * LN_PL];
  const int wib = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < LN_PL; ++k) {
    sred[0][wib][lane + 32 * k] = aw[k];
    sred[1][wib][lane + 32 * k] = ab[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    float tw = 0.f, tb = 0.f;
    for (int q = 0; q < 8; ++q) {
      tw += sred[0][q][i];
      tb += sred[1][q][i];
    }
    atomicAdd(dw + i, tw);
    atomicAdd(db + i, tb);
  }
}

extern "C" int eqv2_ln_silu_bwd2(const float* x, const float* w, const float* b, const float* gy, const float* u, float* dx,
                                 float* dgy, float* dw, float* db, long long rows, int width, float eps, void* stream) {
  if (rows == 0) return 0;
  EQV2_REQUIRE(width > 0 && width <= 32 * LN_PL, "ln_silu_bwd2: width %d > %d", width, 32 * LN_PL);
  long long blocks = (rows * 32 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  EQV2_LAUNCH(ln_silu_bwd2_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, x, w, b, gy, u, dx, dgy, dw, db, rows, width, eps);
  EQV2_CHECK_LAUNCH("eqv2_ln_silu_bwd2");
  return 0;
}

extern "C" int eqv2_ln_silu_fwd(const float* x, const float* w, const float* b, float* y, long long rows, int width,
                                float eps, void* stream) {
  if (rows == 0) return 0;
  EQV2_REQUIRE(width > 0 && width <= 32 * LN_PL, "ln_silu_fwd: width %d > %d", width, 32 * LN_PL);
  EQV2_LAUNCH(ln_silu_fwd_kernel, dim3((unsigned)((rows * 32 + 255) / 256)), dim3(256), 0, stream, x, w, b, y, rows, width, eps);
  EQV2_CHECK_LAUNCH("eqv2_ln_silu_fwd");
  return 0;
}

extern "C" int eqv2_ln_silu_bwd(const float* x, const float* w, const float* b, const float* gy, float* gx, float* gw,
                                float* gb, long long rows, int width, float eps, void* stream) {
  if (rows == 0) return 0;
  EQV2_REQUIRE(width > 0 && width <= 32 * LN_PL, "ln_silu_bwd: width %d > %d", width, 32 * LN_PL);
  long long blocks = (rows * 32 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  EQV2_LAUNCH(ln_silu_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, x, w, b, gy, gx, gw, gb, rows, width, eps);
  EQV2_CHECK_LAUNCH("eqv2_ln_silu_bwd");
  return 0;
}

extern "C" int eqv2_wigner_from_rot(const float* rot, const float* Jd, float* wig, long long E, int lmax,
                                    void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(lmax >= 0 && lmax <= EQV2_MAX_LMAX, "wigner_from_rot: lmax=%d out of range", lmax);
  const int WS = eqv2_wig_off(lmax + 1);
  const size_t smem = (size_t)(WS + WIG_WARPS * MAXN * MAXN + WIG_WARPS * 6 * MAXN) * sizeof(float);
  EQV2_LAUNCH(wigner_kernel, dim3((unsigned)((E + WIG_WARPS - 1) / WIG_WARPS)), dim3(WIG_WARPS * 32), smem, stream, rot, Jd, wig, E, lmax, WS);
  EQV2_CHECK_LAUNCH("eqv2_wigner_from_rot");
  return 0;
}

extern "C" int eqv2_edge_sh(const float* vec, float* out, long long E, int lmax, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(lmax >= 1 && lmax <= SH_MAXL, "edge_sh: lmax=%d out of range (1..%d)", lmax, SH_MAXL);
  EQV2_LAUNCH(edge_sh_kernel, dim3((unsigned)((E + 127) / 128)), dim3(128), 0, stream, vec, out, E, lmax);
  EQV2_CHECK_LAUNCH("eqv2_edge_sh");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Edge frames (edge_rot_mat.py:13-80 with a caller-supplied helper draw; equiformerv2_MatPESv2.py:41-66 deterministic).
// One thread per edge; fp32 in the reference's operation order (no FMA contraction).  Rows of the result: z, x_edge, -y.
//   mode 0: helper = draw[e] (already `rand - 0.5`), normalised; swapped for one of its two 90-degree alternates when
//           (anti)parallel to the edge (edge_rot_mat.py:32-55).  stats[0] = ~bits(min |vec|), stats[1] = bits(max |<helper, x>|)
//           (atomicMax on unsigned; the host checks the reference's two conditions from them in ONE read-back).
//   mode 1: helper = cardinal axis of the smallest |component| of the direction; every norm clamped at 1e-8.
namespace {
__device__ __forceinline__ float ef_norm3(float a, float b, float c) {
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
}
__device__ __forceinline__ float ef_dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}
__global__ void edge_frames_kernel(const float* __restrict__ vec, const float* __restrict__ draw, float* __restrict__ out,
                                   long long E, int mode, unsigned* __restrict__ stats) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float v0 = vec[3 * e], v1 = vec[3 * e + 1], v2 = vec[3 * e + 2];
  float len = ef_norm3(v0, v1, v2);
  float h0, h1, h2;
  if (mode == 0) {
    atomicMax(stats, ~__float_as_uint(len));
  } else {
    len = fmaxf(len, 1e-8f);
  }
  const float x0 = v0 / len, x1 = v1 / len, x2 = v2 / len;
  if (mode == 0) {
    const float d0 = draw[3 * e], d1 = draw[3 * e + 1], d2 = draw[3 * e + 2];
    const float dn = ef_norm3(d0, d1, d2);
    h0 = d0 / dn; h1 = d1 / dn; h2 = d2 / dn;
    const float b0 = -h1, b1 = h0, b2 = h2;            // alternates (edge_rot_mat.py:36-41)
    const float c0 = h0, c1 = -h2, c2 = h1;
    const float dot_b = fabsf(ef_dot3(b0, b1, b2, x0, x1, x2)), dot_c = fabsf(ef_dot3(c0, c1, c2, x0, x1, x2));
    if (fabsf(ef_dot3(h0, h1, h2, x0, x1, x2)) > dot_b) { h0 = b0; h1 = b1; h2 = b2; }
    if (fabsf(ef_dot3(h0, h1, h2, x0, x1, x2)) > dot_c) { h0 = c0; h1 = c1; h2 = c2; }
    atomicMax(stats + 1, __float_as_uint(fabsf(ef_dot3(h0, h1, h2, x0, x1, x2))));
  } else {
    const float a0 = fabsf(x0), a1 = fabsf(x1), a2 = fabsf(x2);
    const int best = (a0 <= a1 && a0 <= a2) ? 0 : ((a1 <= a2) ? 1 : 2);      // torch.argmin: first minimum
    h0 = best == 0 ? 1.f : 0.f; h1 = best == 1 ? 1.f : 0.f; h2 = best == 2 ? 1.f : 0.f;
  }
  // z = x cross helper
  float z0 = __fadd_rn(__fmul_rn(x1, h2), -__fmul_rn(x2, h1));
  float z1 = __fadd_rn(__fmul_rn(x2, h0), -__fmul_rn(x0, h2));
  float z2 = __fadd_rn(__fmul_rn(x0, h1), -__fmul_rn(x1, h0));
  float zn = ef_norm3(z0, z1, z2);
  if (mode == 1) zn = fmaxf(zn, 1e-8f);
  z0 /= zn; z1 /= zn; z2 /= zn;
  if (mode == 0) {                                     // the reference normalises twice (edge_rot_mat.py:60-63)
    zn = ef_norm3(z0, z1, z2);
    z0 /= zn; z1 /= zn; z2 /= zn;
  }
  float y0 = __fadd_rn(__fmul_rn(x1, z2), -__fmul_rn(x2, z1));
  float y1 = __fadd_rn(__fmul_rn(x2, z0), -__fmul_rn(x0, z2));
  float y2 = __fadd_rn(__fmul_rn(x0, z1), -__fmul_rn(x1, z0));
  float yn = ef_norm3(y0, y1, y2);
  if (mode == 1) yn = fmaxf(yn, 1e-8f);
  y0 /= yn; y1 /= yn; y2 /= yn;
  float* o = out + 9 * e;
  o[0] = z0; o[1] = z1; o[2] = z2;
  o[3] = x0; o[4] = x1; o[5] = x2;
  o[6] = -y0; o[7] = -y1; o[8] = -y2;
}
}  // namespace

extern "C" int eqv2_edge_frames(const float* vec, const float* draw, float* out, long long E, int mode, unsigned* stats,
                                void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(mode == 1 || (draw != nullptr && stats != nullptr), "edge_frames: mode 0 needs the helper draw and stats");
  EQV2_LAUNCH(edge_frames_kernel, dim3((unsigned)((E + 127) / 128)), dim3(128), 0, stream, vec, draw, out, E, mode, stats);
  EQV2_CHECK_LAUNCH("eqv2_edge_frames");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// GaussianSmearing + first layer of the radial MLP, fused (equiformerv2_oc20.py:43-60 + radial_function.py:5-30 with the
// edge-scalar input of transformer_block.py:241-248):
//   h[e, :] = W1[:, :R] rbf(d_e) + W1[:, R:R+Ce] E_src[Z_src(e)] + W1[:, R+Ce:] E_dst[Z_dst(e)] + b1
// * exp(-j^2 / (2 w^2)) (w = basis_width_scalar) is below 1e-12 beyond |j| = `band` basis functions from the nearest one,
//   so the R = 600 term sum is a `2 band + 1`-term banded product (<= 1e-6 of the dense result, tests/test_rbf_linear.py);
// * the two embedding terms are row lookups into tables pre-multiplied by their weight slices (T = E W^T, [V, H]).
// The [E, R + 2 Ce] feature matrix (46 MB per block at the OC20 shape), its torch.cat, its operand split and the dense
// K = 856 GEMM disappear; per edge the kernel reads (2 band + 1) rows of W1^T from L2 and writes H floats.
// First-order only: d carries no gradient here (configs 1-2; the MatPES family keeps the differentiable dense path).
namespace {
__global__ void rbf_linear_fwd_kernel(const float* __restrict__ d, const float* __restrict__ offset,
                                      const float* __restrict__ Wt /*[R,H]*/, const float* __restrict__ Ts,
                                      const float* __restrict__ Td, const long long* __restrict__ zs,
                                      const long long* __restrict__ zd, const float* __restrict__ bias,
                                      float* __restrict__ out, long long E, int R, int H, float start, float inv_delta,
                                      float coeff, int band) {
  const long long e = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  if (e >= E) return;
  const float de = d[e];
  int k0 = (int)rintf((de - start) * inv_delta);
  k0 = k0 < 0 ? 0 : (k0 > R - 1 ? R - 1 : k0);
  const int lo = k0 - band < 0 ? 0 : k0 - band, hi = k0 + band > R - 1 ? R - 1 : k0 + band;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = bias != nullptr ? bias[c] : 0.f;
    if (Ts != nullptr) acc += Ts[zs[e] * H + c] + Td[zd[e] * H + c];
    // four independent partial sums: the row loads of W1^T (L2) are what the loop waits for
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = lo;
    for (; k + 3 <= hi; k += 4) {
      const float t0 = de - offset[k], t1 = de - offset[k + 1], t2 = de - offset[k + 2], t3 = de - offset[k + 3];
      const float w0 = __ldg(Wt + (long long)k * H + c), w1 = __ldg(Wt + (long long)(k + 1) * H + c);
      const float w2 = __ldg(Wt + (long long)(k + 2) * H + c), w3 = __ldg(Wt + (long long)(k + 3) * H + c);
      a0 = fmaf(expf(coeff * t0 * t0), w0, a0);
      a1 = fmaf(expf(coeff * t1 * t1), w1, a1);
      a2 = fmaf(expf(coeff * t2 * t2), w2, a2);
      a3 = fmaf(expf(coeff * t3 * t3), w3, a3);
    }
    for (; k <= hi; ++k) {
      const float t = de - offset[k];
      a0 = fmaf(expf(coeff * t * t), __ldg(Wt + (long long)k * H + c), a0);
    }
    out[e * H + c] = acc + ((a0 + a1) + (a2 + a3));
  }
}

// Weight gradient gWt[k, c] = sum_e rbf_k(d_e) gh[e, c], deterministic, in two stages over the edges SORTED by nearest
// basis index (perm; built once per graph).  A first version (one CTA per basis function walking its band of bins) was
// badly balanced: with max_neighbors = 20 the kept distances populate ~40 % of the 600 bins, so ~240 CTAs did all the
// work (0.14-0.49 ms per launch).  Now the EDGES are the parallel dimension:
//   stage 1: CTA = chunk of RL_CHUNK consecutive sorted edges.  The basis functions its edges can touch are the contiguous
//            range [kmin, kmax] (chunk_k[2 chunk], chunk_k[2 chunk + 1]; bins +- band); the CTA stages the chunk's gradient
//            rows in shared memory, evaluates the weights w[e][k] tile by tile (32 basis functions at a time) and writes one
//            partial row per touched k to partial[chunk_base[chunk] + k - kmin].
//   stage 2: CTA = basis function k: adds the partial rows of the chunks that touch k -- the contiguous chunk range
//            [k_chunks[2k], k_chunks[2k+1]] -- in chunk order.
constexpr int RL_CHUNK = 64;
constexpr int RL_KT = 32;
__global__ void __launch_bounds__(128)
rbf_linear_wgrad_partial_kernel(const float* __restrict__ d, const float* __restrict__ offset,
                                const int* __restrict__ perm, const int* __restrict__ chunk_k,
                                const int* __restrict__ chunk_base, const float* __restrict__ gh,
                                float* __restrict__ partial, long long E, int H, float coeff) {
  EQV2_DYN_SMEM(float, sm);                  // gh rows [RL_CHUNK][H] | w [RL_CHUNK][RL_KT] | d [RL_CHUNK]
  float* sg = sm;
  float* sw = sm + RL_CHUNK * H;
  float* sd = sw + RL_CHUNK * RL_KT;
  const int chunk = blockIdx.x;
  const long long e0 = (long long)chunk * RL_CHUNK;
  const int ne = (int)(E - e0 < RL_CHUNK ? E - e0 : RL_CHUNK);
  const int kmin = chunk_k[2 * chunk], kmax = chunk_k[2 * chunk + 1];
  for (int i = threadIdx.x; i < ne; i += blockDim.x) sd[i] = d[perm[e0 + i]];
  for (int i = threadIdx.x; i < ne * (H / 4); i += blockDim.x) {
    const int r = i / (H / 4), q = i - r * (H / 4);
    reinterpret_cast<float4*>(sg + r * H)[q] = __ldg(reinterpret_cast<const float4*>(gh + (long long)perm[e0 + r] * H) + q);
  }
  __syncthreads();
  float* prow = partial + (long long)chunk_base[chunk] * H;
  for (int kt = kmin; kt <= kmax; kt += RL_KT) {
    const int nk = kmax - kt + 1 < RL_KT ? kmax - kt + 1 : RL_KT;
    for (int i = threadIdx.x; i < ne * RL_KT; i += blockDim.x) {
      const int r = i / RL_KT, kk = i - r * RL_KT;
      float w = 0.f;
      if (kk < nk) {
        const float t = sd[r] - offset[kt + kk];
        w = expf(coeff * t * t);
      }
      sw[i] = w;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
      float acc[RL_KT];
#pragma unroll
      for (int kk = 0; kk < RL_KT; ++kk) acc[kk] = 0.f;
      for (int r = 0; r < ne; ++r) {
        const float g = sg[r * H + c];
        const float4* wr = reinterpret_cast<const float4*>(sw + r * RL_KT);       // broadcast reads
#pragma unroll
        for (int q = 0; q < RL_KT / 4; ++q) {
          const float4 w = wr[q];
          acc[4 * q] = fmaf(w.x, g, acc[4 * q]);
          acc[4 * q + 1] = fmaf(w.y, g, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(w.z, g, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(w.w, g, acc[4 * q + 3]);
        }
      }
#pragma unroll
      for (int kk = 0; kk < RL_KT; ++kk)
        if (kk < nk) prow[(long long)(kt - kmin + kk) * H + c] = acc[kk];
    }
    __syncthreads();
  }
}

__global__ void rbf_linear_wgrad_final_kernel(const float* __restrict__ partial, const int* __restrict__ chunk_k,
                                              const int* __restrict__ chunk_base, const int* __restrict__ k_chunks,
                                              float* __restrict__ gWt, int H) {
  const int k = blockIdx.x;
  const int clo = k_chunks[2 * k], chi = k_chunks[2 * k + 1];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float a = 0.f;
    for (int ch = clo; ch <= chi; ++ch)
      a += partial[(long long)(chunk_base[ch] + k - chunk_k[2 * ch]) * H + c];
    gWt[(long long)k * H + c] = a;
  }
}
}  // namespace

extern "C" int eqv2_rbf_linear_fwd(const float* d, const float* offset, const float* Wt, const float* Ts, const float* Td,
                                   const long long* zs, const long long* zd, const float* bias, float* out, long long E,
                                   int R, int H, float start, float delta, float coeff, int band, void* stream) {
  if (E == 0) return 0;
  EQV2_REQUIRE(R >= 2 && H > 0 && band >= 0 && delta > 0.f, "rbf_linear_fwd: bad sizes");
  EQV2_REQUIRE((Ts == nullptr) == (Td == nullptr) && (Ts == nullptr || (zs != nullptr && zd != nullptr)),
               "rbf_linear_fwd: embedding tables and element types come together");
  const int tx = H >= 128 ? 128 : (H >= 64 ? 64 : 32), ty = 256 / tx;
  EQV2_LAUNCH(rbf_linear_fwd_kernel, dim3((unsigned)((E + ty - 1) / ty)), dim3(tx, ty), 0, stream, d, offset, Wt, Ts, Td,
              zs, zd, bias, out, E, R, H, start, 1.0f / delta, coeff, band);
  EQV2_CHECK_LAUNCH("eqv2_rbf_linear_fwd");
  return 0;
}

extern "C" int eqv2_rbf_linear_chunk(void) { return RL_CHUNK; }

extern "C" int eqv2_rbf_linear_wgrad(const float* d, const float* offset, const int* perm, const int* chunk_k,
                                     const int* chunk_base, const int* k_chunks, const float* gh, float* partial,
                                     float* gWt, long long E, int R, int H, float coeff, void* stream) {
  EQV2_REQUIRE(R >= 2 && H > 0 && (H % 4) == 0, "rbf_linear_wgrad: H must be a positive multiple of 4");
  EQV2_REQUIRE((((uintptr_t)gh) & 15) == 0, "rbf_linear_wgrad: gh must be 16-byte aligned");
  const size_t smem = (size_t)(RL_CHUNK * H + RL_CHUNK * RL_KT + RL_CHUNK) * sizeof(float);
  EQV2_REQUIRE(smem <= 48 * 1024, "rbf_linear_wgrad: H = %d too wide for the staged chunk", H);
  if (E > 0) {
    const unsigned nchunks = (unsigned)((E + RL_CHUNK - 1) / RL_CHUNK);
    EQV2_LAUNCH(rbf_linear_wgrad_partial_kernel, dim3(nchunks), dim3(128), smem, stream, d, offset, perm, chunk_k, chunk_base,
                gh, partial, E, H, coeff);
    EQV2_CHECK_LAUNCH("eqv2_rbf_linear_wgrad (partial)");
  }
  EQV2_LAUNCH(rbf_linear_wgrad_final_kernel, dim3((unsigned)R), dim3(H >= 128 ? 128 : (H >= 64 ? 64 : 32)), 0, stream, partial,
              chunk_k, chunk_base, k_chunks, gWt, H);
  EQV2_CHECK_LAUNCH("eqv2_rbf_linear_wgrad (final)");
  return 0;
}
