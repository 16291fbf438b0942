// Grouped fp32 GEMM on the CUDA cores (FFMA) -- the exact-fp32 engine behind every dense
// contraction of the path: SO(2) convolution blocks (so2_ops.py:150-185), radial MLP linears
// (radial_function.py:29), SO3_LinearV2 (so3.py:722-727) and their dgrad / wgrad.
//
//   C[i,j] (+)= sum_k opA(i,k) * opB(k,j) (+ bias[j])
//
// Every operand row is addressed with a two-level stride
//   row_off(r) = (r / rpb) * bstride + (r % rpb) * ld
// so that a degree-l slab x[:, l*l:(l+1)^2, :] of an [N,K,C] node tensor is a GEMM operand
// without a copy.  Up to EQV2_GEMM_MAX_GROUPS independent problems per launch (the m = 0..mmax
// blocks of one SO(2) convolution, or the l = 0..lmax blocks of one SO3 linear), optional
// split-K with fp32 atomics for the reduction-heavy weight gradients.
//
// Tile: 128x128x16, 256 threads, 8x8 outputs per thread, register-staged double buffering.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int PAD = 4;

struct GemmParams {
  eqv2_gemm_desc g[EQV2_GEMM_MAX_GROUPS];
  int ngroups;
  int split_k;
};

// rows-per-block >= 2^31 marks a plain row-major operand: no 64-bit division on the hot path
__device__ __forceinline__ long long row_off(long long r, long long rpb, long long bs, long long ld) {
  if (rpb >= (1ll << 31)) return r * ld;
  const int q = (int)r / (int)rpb;
  return (long long)q * bs + (long long)((int)r - q * (int)rpb) * ld;
}

__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmParams P) {
  const int zi = blockIdx.z;
  const int gi = zi / P.split_k;
  const int ks = zi % P.split_k;
  const eqv2_gemm_desc& d = P.g[gi];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const bool active = (m0 < d.M) && (n0 < d.N);
  if (!active) return;   // CTA-uniform: the grid is sized for the largest group

  __shared__ float As[2][BK][BM + PAD];
  __shared__ float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // K range of this split (multiples of BK)
  const int ktiles = (d.K + BK - 1) / BK;
  const int per = (ktiles + P.split_k - 1) / P.split_k;
  const int kt0 = ks * per;
  const int kt1 = min(ktiles, kt0 + per);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];

  // loader geometry: 128x16 = 2048 elements / 256 threads = 8 each.
  // contiguous-in-k operands: thread -> (row = e*16 + tid/16, k = tid%16)
  // contiguous-in-row operands: thread -> (k = e*2 + tid/128, row = tid%128)
  auto load_tiles = [&](int kt) {
    const int kbase = kt * BK;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int r, k;
      if (!d.transA) { r = e * 16 + (tid >> 4); k = tid & 15; }
      else           { k = e * 2 + (tid >> 7);  r = tid & 127; }
      const int gm = m0 + r, gk = kbase + k;
      float v = 0.f;
      if (active && gm < d.M && gk < d.K) {
        v = d.transA ? d.A[row_off(gk, d.a_rpb, d.a_bs, d.a_ld) + gm]
                     : d.A[row_off(gm, d.a_rpb, d.a_bs, d.a_ld) + gk];
      }
      ra[e] = v;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int c, k;
      if (d.transB) { c = e * 16 + (tid >> 4); k = tid & 15; }
      else          { k = e * 2 + (tid >> 7);  c = tid & 127; }
      const int gn = n0 + c, gk = kbase + k;
      float v = 0.f;
      if (active && gn < d.N && gk < d.K) {
        v = d.transB ? d.B[row_off(gn, d.b_rpb, d.b_bs, d.b_ld) + gk]
                     : d.B[row_off(gk, d.b_rpb, d.b_bs, d.b_ld) + gn];
      }
      rb[e] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int r, k;
      if (!d.transA) { r = e * 16 + (tid >> 4); k = tid & 15; }
      else           { k = e * 2 + (tid >> 7);  r = tid & 127; }
      As[buf][k][r] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int c, k;
      if (d.transB) { c = e * 16 + (tid >> 4); k = tid & 15; }
      else          { k = e * 2 + (tid >> 7);  c = tid & 127; }
      Bs[buf][k][c] = rb[e];
    }
  };

  if (kt0 < kt1) {
    load_tiles(kt0);
    store_tiles(0);
  }
  __syncthreads();
  int buf = 0;
  for (int kt = kt0; kt < kt1; ++kt) {
    if (kt + 1 < kt1) load_tiles(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < kt1) store_tiles(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  if (!active) return;
  const bool use_atomic = (P.split_k > 1);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= d.M) continue;
    float* crow = d.C + row_off(gm, d.c_rpb, d.c_bs, d.c_ld);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= d.N) continue;
      float v = acc[i][j];
      if (d.bias != nullptr && ks == 0) v += d.bias[gn];
      if (use_atomic) atomicAdd(&crow[gn], v);
      else if (d.accumulate) crow[gn] += v;
      else crow[gn] = v;
    }
  }
}

}  // namespace

extern "C" int eqv2_gemm_f32(const eqv2_gemm_desc* descs, int ngroups, int split_k, void* stream) {
  EQV2_REQUIRE(ngroups >= 1 && ngroups <= EQV2_GEMM_MAX_GROUPS, "eqv2_gemm_f32: ngroups=%d out of range", ngroups);
  EQV2_REQUIRE(split_k >= 1, "eqv2_gemm_f32: split_k must be >= 1");
  GemmParams P;
  memset(&P, 0, sizeof(P));
  int maxM = 0, maxN = 0;
  for (int i = 0; i < ngroups; ++i) {
    P.g[i] = descs[i];
    EQV2_REQUIRE(descs[i].A && descs[i].B && descs[i].C, "eqv2_gemm_f32: null operand in group %d", i);
    EQV2_REQUIRE(descs[i].a_rpb > 0 && descs[i].b_rpb > 0 && descs[i].c_rpb > 0, "eqv2_gemm_f32: rpb must be > 0");
    if (descs[i].M > maxM) maxM = descs[i].M;
    if (descs[i].N > maxN) maxN = descs[i].N;
  }
  if (maxM == 0 || maxN == 0) return 0;
  P.ngroups = ngroups;
  P.split_k = split_k;
  dim3 grid((maxN + BN - 1) / BN, (maxM + BM - 1) / BM, ngroups * split_k);
  EQV2_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "eqv2_gemm_f32: grid too large (M=%d)", maxM);
  EQV2_LAUNCH(gemm_simt_kernel, grid, dim3(NT), 0, stream, P);
  EQV2_CHECK_LAUNCH("eqv2_gemm_f32");
  return 0;
}
