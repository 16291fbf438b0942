// Equivariant normalisation layers (layer_norm.py): 'layer_norm' (:38-108), 'layer_norm_sh'
// (:112-201) and 'rms_norm_sh' V2 (:265-351) expressed as ONE kernel over "degree groups":
//   * the l = 0 row is centred over channels,
//   * every group g has s_g = mean_c sum_{k in g} bw_k f[k,c]^2 and inv_g = (s_g + eps)^-1/2,
//   * out[k,c] = f[k,c] * inv_g(k) * w[l(k),c]  (+ b[c] on k = 0).
//   rms_norm_sh  : one group, bw_k = 1/((2l+1)(L+1))
//   layer_norm_sh: {l=0} with bw = 1 (== nn.LayerNorm), {l>0} with bw_k = 1/((2l+1) L)
//   layer_norm   : one group per degree, bw_k = 1/(2l+1)
// One CTA per node, one thread per channel; inv_g and the l=0 mean are saved for backward.
#include "common.cuh"

namespace {

constexpr int MAXG = EQV2_MAX_LMAX + 1;

struct NormMeta {
  int group_of_l[MAXG];
  float bw_l[MAXG];
  int ngroups;
  int lmax;
};

// sums `nv` per-thread values over the CTA; result broadcast in smem `res`
__device__ __forceinline__ void block_sum(float* vals, int nv, float* scratch, float* res) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = 0; i < nv; ++i) {
    const float s = eqv2_warp_sum(vals[i]);
    if (lane == 0) scratch[i * 32 + warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < nv) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += scratch[threadIdx.x * 32 + w];
    res[threadIdx.x] = s;
  }
  __syncthreads();
}

__global__ void equiv_norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ b, float* __restrict__ out, float* __restrict__ inv_out,
                                      float* __restrict__ mean_out, const NormMeta M, int C, float eps,
                                      float* __restrict__ absmax) {
  __shared__ float scratch[MAXG * 32];
  __shared__ float res[MAXG];
  const int K = (M.lmax + 1) * (M.lmax + 1);
  const long long n = blockIdx.x;
  const int c = threadIdx.x;
  const bool live = c < C;
  const float* xp = x + n * (long long)K * C + c;
  float v[MAXG];
  v[0] = live ? xp[0] : 0.f;
  block_sum(v, 1, scratch, res);
  const float mean0 = res[0] / C;
  __syncthreads();
  for (int g = 0; g < MAXG; ++g) v[g] = 0.f;
  if (live) {
    for (int l = 0; l <= M.lmax; ++l) {
      float s = 0.f;
      for (int k = l * l; k < (l + 1) * (l + 1); ++k) {
        float f = xp[(long long)k * C];
        if (k == 0) f -= mean0;
        s = fmaf(f, f, s);
      }
      // group ids are < MAXG; static indexing keeps v[] in registers
#pragma unroll
      for (int g = 0; g < MAXG; ++g)
        if (g == M.group_of_l[l]) v[g] = fmaf(s, M.bw_l[l], v[g]);
    }
  }
  block_sum(v, M.ngroups, scratch, res);
  float inv[MAXG];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) inv[g] = (g < M.ngroups) ? rsqrtf(res[g] / C + eps) : 0.f;
  if (threadIdx.x < M.ngroups) inv_out[n * M.ngroups + threadIdx.x] = inv[threadIdx.x];
  if (threadIdx.x == 0) mean_out[n] = mean0;
  float amax = 0.f;
  if (live) {
    float* op = out + n * (long long)K * C + c;
    for (int l = 0; l <= M.lmax; ++l) {
      float ig = 0.f;
#pragma unroll
      for (int g = 0; g < MAXG; ++g)
        if (g == M.group_of_l[l]) ig = inv[g];
      const float sc = ig * w[l * C + c];
      for (int k = l * l; k < (l + 1) * (l + 1); ++k) {
        float f = xp[(long long)k * C];
        if (k == 0) f -= mean0;
        float o = f * sc;
        if (k == 0) o += b[c];
        op[(long long)k * C] = o;
        amax = fmaxf(amax, fabsf(o));
      }
    }
  }
  // max |out|: the operand bound of the gather/rotate kernel that consumes the normed embedding (common.cuh)
  if (absmax != nullptr) eqv2_commit_absmax(amax, absmax);
}

__global__ void equiv_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ go, const float* __restrict__ inv_in,
                                      const float* __restrict__ mean_in, float* __restrict__ dx,
                                      float* __restrict__ dw, float* __restrict__ db, const NormMeta M, int C) {
  __shared__ float scratch[MAXG * 32];
  __shared__ float res[MAXG];
  const int K = (M.lmax + 1) * (M.lmax + 1);
  const long long n = blockIdx.x;
  const int c = threadIdx.x;
  const bool live = c < C;
  const float* xp = x + n * (long long)K * C + c;
  const float* gp = go + n * (long long)K * C + c;
  const float mean0 = mean_in[n];
  float inv[MAXG], v[MAXG];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) {
    inv[g] = (g < M.ngroups) ? inv_in[n * M.ngroups + g] : 0.f;
    v[g] = 0.f;
  }
  // T_g = sum_{k in g, c} go * w * f
  if (live) {
    for (int l = 0; l <= M.lmax; ++l) {
      float s = 0.f;
      for (int k = l * l; k < (l + 1) * (l + 1); ++k) {
        float f = xp[(long long)k * C];
        if (k == 0) f -= mean0;
        s = fmaf(gp[(long long)k * C], f, s);
      }
      s *= w[l * C + c];
#pragma unroll
      for (int g = 0; g < MAXG; ++g)
        if (g == M.group_of_l[l]) v[g] += s;
    }
  }
  block_sum(v, M.ngroups, scratch, res);
  float ds[MAXG];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) ds[g] = (g < M.ngroups) ? -0.5f * inv[g] * inv[g] * inv[g] * res[g] : 0.f;
  __syncthreads();
  // row 0 needs the channel mean of its own gradient (centring)
  float g0 = 0.f;
  if (live) {
    const float f = xp[0] - mean0;
    float ig = 0.f, dg = 0.f;
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
      if (g == M.group_of_l[0]) { ig = inv[g]; dg = ds[g]; }
    g0 = gp[0] * w[c] * ig + dg * 2.f * M.bw_l[0] * f / C;
  }
  v[0] = g0;
  block_sum(v, 1, scratch, res);
  const float g0mean = res[0] / C;
  if (live) {
    float* dp = dx + n * (long long)K * C + c;
    for (int l = 0; l <= M.lmax; ++l) {
      float ig = 0.f, dg = 0.f;
#pragma unroll
      for (int g = 0; g < MAXG; ++g)
        if (g == M.group_of_l[l]) { ig = inv[g]; dg = ds[g]; }
      const float wl = w[l * C + c];
      float dwl = 0.f;
      for (int k = l * l; k < (l + 1) * (l + 1); ++k) {
        float f = xp[(long long)k * C];
        if (k == 0) f -= mean0;
        const float gk = gp[(long long)k * C];
        dwl = fmaf(gk, f, dwl);
        float gf = gk * wl * ig + dg * 2.f * M.bw_l[l] * f / C;
        if (k == 0) gf -= g0mean;
        dp[(long long)k * C] = gf;
      }
      atomicAdd(&dw[l * C + c], dwl * ig);
    }
    atomicAdd(&db[c], gp[0]);
  }
}

// Derivative of the BACKWARD pass for a cotangent u of dx (forces by autograd; see edge_feat.cu ln_silu_bwd2 for the
// pattern).  With a = go w, f = centred x, inv_g = (s_g + eps)^-1/2, t = bw f / C, u~ = u with its l = 0 row centred:
//   first backward:  gf = inv_g a - inv_g^3 bw f D_g,  D_g = (1/C) sum_{g} a f;  dx = gf with its l = 0 row centred.
//   S = <u, dx> = sum_g [ inv_g A_g - inv_g^3 D_g B_g ],   A_g = sum_g u~ a,   B_g = sum_g bw u~ f.
//   dS/dgo = w (inv_g u~ - inv_g^3 B_g f / C)                  dS/dw = sum over nodes and rows of go (same bracket)
//   dS/df  = -inv_g^3 t A_g + 3 inv_g^5 t D_g B_g - inv_g^3 B_g a / C - inv_g^3 D_g bw u~;   dS/dx = dS/df, l = 0 row centred.
__global__ void equiv_norm_bwd2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                       const float* __restrict__ go, const float* __restrict__ inv_in,
                                       const float* __restrict__ mean_in, const float* __restrict__ u,
                                       float* __restrict__ d2x, float* __restrict__ dgo, float* __restrict__ dw,
                                       const NormMeta M, int C) {
  __shared__ float scratch[3 * MAXG * 32];
  __shared__ float res[3 * MAXG];
  const int K = (M.lmax + 1) * (M.lmax + 1);
  const long long n = blockIdx.x;
  const int c = threadIdx.x;
  const bool live = c < C;
  const long long base = n * (long long)K * C + c;
  const float* xp = x + base;
  const float* gp = go + base;
  const float* up = u + base;
  const float mean0 = mean_in[n];
  float inv[MAXG];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) inv[g] = (g < M.ngroups) ? inv_in[n * M.ngroups + g] : 0.f;
  float v[3 * MAXG];
  v[0] = live ? up[0] : 0.f;
  block_sum(v, 1, scratch, res);
  const float u0mean = res[0] / C;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 3 * MAXG; ++i) v[i] = 0.f;
  if (live) {
    for (int l = 0; l <= M.lmax; ++l) {
      const float wl = w[l * C + c], bw = M.bw_l[l];
      float sA = 0.f, sB = 0.f, sD = 0.f;
      for (int k = l * l; k < (l + 1) * (l + 1); ++k) {
        float f = xp[(long long)k * C], ut = up[(long long)k * C];
        if (k == 0) { f -= mean0; ut -= u0mean; }
        const float a = gp[(long long)k * C] * wl;
        sA = fmaf(ut, a, sA);
        sB = fmaf(bw * ut, f, sB);
        sD = fmaf(a, f, sD);
      }
#pragma unroll
      for (int g = 0; g < MAXG; ++g)
        if (g == M.group_of_l[l]) { v[3 * g] += sA; v[3 * g + 1] += sB; v[3 * g + 2] += sD; }
    }
  }
  block_sum(v, 3 * M.ngroups, scratch, res);
  float A[MAXG], B[MAXG], D[MAXG];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) {
    A[g] = (g < M.ngroups) ? res[3 * g] : 0.f;
    B[g] = (g < M.ngroups) ? res[3 * g + 1] : 0.f;
    D[g] = (g < M.ngroups) ? res[3 * g + 2] / C : 0.f;
  }
  __syncthreads();
  // the l = 0 row of dS/df needs its channel mean (centring)
  float h0 = 0.f;
  if (live) {
    float ig = 0.f, Ag = 0.f, Bg = 0.f, Dg = 0.f;
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
      if (g == M.group_of_l[0]) { ig = inv[g]; Ag = A[g]; Bg = B[g]; Dg = D[g]; }
    const float f = xp[0] - mean0, ut = up[0] - u0mean, a = gp[0] * w[c], bw = M.bw_l[0];
    const float i3 = ig * ig * ig, t = bw * f / C;
    h0 = -i3 * t * Ag + 3.f * i3 * ig * ig * t * Dg * Bg - i3 * Bg * a / C - i3 * Dg * bw * ut;
  }
  v[0] = h0;
  block_sum(v, 1, scratch, res);
  const float h0mean = res[0] / C;
  if (live) {
    float* dxp = d2x + base;
    float* dgp = dgo + base;
    for (int l = 0; l <= M.lmax; ++l) {
      float ig = 0.f, Ag = 0.f, Bg = 0.f, Dg = 0.f;
#pragma unroll
      for (int g = 0; g < MAXG; ++g)
        if (g == M.group_of_l[l]) { ig = inv[g]; Ag = A[g]; Bg = B[g]; Dg = D[g]; }
      const float wl = w[l * C + c], bw = M.bw_l[l];
      const float i3 = ig * ig * ig, i5 = i3 * ig * ig;
      float dwl = 0.f;
      for (int k = l * l; k < (l + 1) * (l + 1); ++k) {
        float f = xp[(long long)k * C], ut = up[(long long)k * C];
        if (k == 0) { f -= mean0; ut -= u0mean; }
        const float gk = gp[(long long)k * C];
        const float a = gk * wl, t = bw * f / C;
        const float br = ig * ut - i3 * Bg * f / C;          // dS/da
        dgp[(long long)k * C] = wl * br;
        dwl = fmaf(gk, br, dwl);
        float h = -i3 * t * Ag + 3.f * i5 * t * Dg * Bg - i3 * Bg * a / C - i3 * Dg * bw * ut;
        if (k == 0) h -= h0mean;
        dxp[(long long)k * C] = h;
      }
      atomicAdd(&dw[l * C + c], dwl);
    }
  }
}

}  // namespace

static int fill_meta(NormMeta& M, int lmax, int ngroups, const int* group_of_l, const float* bw_l) {
  if (lmax < 0 || lmax > EQV2_MAX_LMAX) return 1;
  memset(&M, 0, sizeof(M));
  M.lmax = lmax;
  M.ngroups = ngroups;
  for (int l = 0; l <= lmax; ++l) {
    if (group_of_l[l] < 0 || group_of_l[l] >= ngroups) return 1;
    M.group_of_l[l] = group_of_l[l];
    M.bw_l[l] = bw_l[l];
  }
  return 0;
}

extern "C" int eqv2_equiv_norm_fwd(const float* x, const float* w, const float* b, float* out, float* inv_out,
                                   float* mean_out, long long N, int C, int lmax, int ngroups, const int* group_of_l,
                                   const float* bw_l, float eps, float* absmax, void* stream) {
  if (N == 0) return 0;
  EQV2_REQUIRE(C > 0 && C <= 1024, "equiv_norm_fwd: C=%d out of range", C);
  NormMeta M;
  EQV2_REQUIRE(fill_meta(M, lmax, ngroups, group_of_l, bw_l) == 0, "equiv_norm_fwd: bad group table");
  const int threads = (C + 31) / 32 * 32;
  EQV2_LAUNCH(equiv_norm_fwd_kernel, dim3((unsigned)N), dim3(threads), 0, stream, x, w, b, out, inv_out, mean_out, M, C, eps, absmax);
  EQV2_CHECK_LAUNCH("eqv2_equiv_norm_fwd");
  return 0;
}

extern "C" int eqv2_equiv_norm_bwd(const float* x, const float* w, const float* go, const float* inv_in,
                                   const float* mean_in, float* dx, float* dw, float* db, long long N, int C, int lmax,
                                   int ngroups, const int* group_of_l, const float* bw_l, void* stream) {
  if (N == 0) return 0;
  EQV2_REQUIRE(C > 0 && C <= 1024, "equiv_norm_bwd: C=%d out of range", C);
  NormMeta M;
  EQV2_REQUIRE(fill_meta(M, lmax, ngroups, group_of_l, bw_l) == 0, "equiv_norm_bwd: bad group table");
  const int threads = (C + 31) / 32 * 32;
  EQV2_LAUNCH(equiv_norm_bwd_kernel, dim3((unsigned)N), dim3(threads), 0, stream, x, w, go, inv_in, mean_in, dx, dw, db, M, C);
  EQV2_CHECK_LAUNCH("eqv2_equiv_norm_bwd");
  return 0;
}

extern "C" int eqv2_equiv_norm_bwd2(const float* x, const float* w, const float* go, const float* inv_in,
                                    const float* mean_in, const float* u, float* d2x, float* dgo, float* dw, long long N,
                                    int C, int lmax, int ngroups, const int* group_of_l, const float* bw_l,
                                    void* stream) {
  if (N == 0) return 0;
  EQV2_REQUIRE(C > 0 && C <= 1024, "equiv_norm_bwd2: C=%d out of range", C);
  NormMeta M;
  EQV2_REQUIRE(fill_meta(M, lmax, ngroups, group_of_l, bw_l) == 0, "equiv_norm_bwd2: bad group table");
  const int threads = (C + 31) / 32 * 32;
  EQV2_LAUNCH(equiv_norm_bwd2_kernel, dim3((unsigned)N), dim3(threads), 0, stream, x, w, go, inv_in, mean_in, u, d2x, dgo, dw, M, C);
  EQV2_CHECK_LAUNCH("eqv2_equiv_norm_bwd2");
  return 0;
}
