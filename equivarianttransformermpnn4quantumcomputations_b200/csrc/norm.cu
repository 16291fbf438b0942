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
