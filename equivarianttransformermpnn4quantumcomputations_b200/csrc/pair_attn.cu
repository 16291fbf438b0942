// All-pairs attention core of the global-attention classes (BASELINE config 5; reference
// NewFunctions/GATA_and_all2all/activation.py:419-1567).  The reference builds an [N_tot, N_tot] attention map over ALL
// atoms of the batch and masks cross-structure pairs (or pads to [B, N_max]); only the pairs inside a structure carry
// weight, so the map is stored here RAGGED: row i (a query atom of structure g, n_g atoms, first atom s_g) owns the n_g
// contiguous entries  pair(i, j) = rowptr[i] + (j - s_g),  each with H heads: tensors [P, H], P = sum_g n_g^2.
//
// Three operators, closed under differentiation (forces by autograd need the derivative of the backward pass):
//   scores : S[pair(i,j), h] = scale * sum_{m,d} a[i,m,h,d] b[j,m,h,d]                       a, b [N, M, H, D]
//   mix    : out[i,m,h,d]    = sum_j W[pair(i,j), h] b[j,m,h,d]            (transpose: out[j] = sum_i W[pair(i,j)] b[i])
//   softmax over each row (+ its backward, + the derivative of its backward)
//     d scores / d a = scale * mix(gS, b),   d scores / d b = scale * mix^T(gS, a)
//     d mix / d W    = scores(gout, b)  (transpose: scores(b, gout)),      d mix / d b = mix^T(W, gout)
// Every output element is produced by one thread / one warp in a fixed order: deterministic, no atomics.
// One launch per operator for the whole batch replaces the reference's (and round 1's) per-structure einsum / softmax /
// einsum launches (5 + 2 lmax library calls per structure and layer).
#include "common.cuh"

namespace {

constexpr int PA_THREADS = 256;

// scores: CTA = query row i; a[i] (M x C floats) staged in shared memory; warp w takes keys j = w, w + 8, ...;
// lane handles the float4 chunks c4 = lane, lane + 32, ... of every m-row; D / 4 adjacent lanes form a head.
__global__ void __launch_bounds__(PA_THREADS)
pair_scores_kernel(const float* __restrict__ a, const float* __restrict__ b, const int* __restrict__ gstart,
                   const int* __restrict__ gcount, const long long* __restrict__ rowptr, float* __restrict__ S,
                   int M, int H, int D, float scale) {
  EQV2_DYN_SMEM(float, sa);            // [M][C]
  const int i = blockIdx.x;
  const int C = H * D, C4 = C >> 2, lanes_per_head = D >> 2;
  const int s = gstart[i], n = gcount[i];
  const long long MC = (long long)M * C;
  for (int t = threadIdx.x; t < M * C4; t += blockDim.x)
    reinterpret_cast<float4*>(sa)[t] = __ldg(reinterpret_cast<const float4*>(a + i * MC) + t);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* Srow = S + rowptr[i] * H;
  for (int j = warp; j < n; j += nwarps) {
    const float4* bj = reinterpret_cast<const float4*>(b + (long long)(s + j) * MC);
    for (int c0 = 0; c0 < C4; c0 += 32) {              // chunks of 32 float4 = 128 channels
      const int c4 = c0 + lane;
      float acc = 0.f;
      if (c4 < C4) {
        for (int m = 0; m < M; ++m) {
          const float4 x = reinterpret_cast<const float4*>(sa)[m * C4 + c4];
          const float4 y = __ldg(bj + m * C4 + c4);
          acc = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, acc))));
        }
      }
      for (int o = 1; o < lanes_per_head; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (c4 < C4 && (lane & (lanes_per_head - 1)) == 0) Srow[(long long)j * H + c4 / lanes_per_head] = acc * scale;
    }
  }
}

// mix: CTA = output row; the row's weights [n][H] staged in shared memory (transpose: gathered from the n rows of the
// structure); thread t owns float4 column chunks (m, c4) and walks the structure's atoms in order.
__global__ void __launch_bounds__(PA_THREADS)
pair_mix_kernel(const float* __restrict__ W, const float* __restrict__ b, const int* __restrict__ gstart,
                const int* __restrict__ gcount, const long long* __restrict__ rowptr, float* __restrict__ out,
                int M, int H, int D, int transpose) {
  EQV2_DYN_SMEM(float, sw);            // [n][H]
  const int i = blockIdx.x;
  const int C = H * D, C4 = C >> 2, lanes_per_head = D >> 2;
  const int s = gstart[i], n = gcount[i];
  const long long MC = (long long)M * C;
  for (int t = threadIdx.x; t < n * H; t += blockDim.x) {
    const int j = t / H, h = t - j * H;
    sw[t] = transpose ? __ldg(W + (rowptr[s + j] + (i - s)) * H + h) : __ldg(W + (rowptr[i] + j) * H + h);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < M * C4; t += blockDim.x) {
    const int c4 = t % C4, h = c4 / lanes_per_head;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* bp = reinterpret_cast<const float4*>(b + (long long)s * MC) + t;
    for (int j = 0; j < n; ++j) {
      const float w = sw[j * H + h];
      const float4 y = __ldg(bp + (long long)j * (MC >> 2));
      acc.x = fmaf(w, y.x, acc.x);
      acc.y = fmaf(w, y.y, acc.y);
      acc.z = fmaf(w, y.z, acc.z);
      acc.w = fmaf(w, y.w, acc.w);
    }
    reinterpret_cast<float4*>(out + i * MC)[t] = acc;
  }
}

// ---- row softmax: one warp per (row, head); lanes stride over the row's n entries ----
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void pair_softmax_fwd_kernel(const float* __restrict__ S, const int* __restrict__ gcount,
                                        const long long* __restrict__ rowptr, float* __restrict__ Pw, long long N, int H) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= N * H) return;
  const long long i = w / H;
  const int h = (int)(w - i * H), n = gcount[i];
  const float* sp = S + rowptr[i] * H + h;
  float* pp = Pw + rowptr[i] * H + h;
  float mx = -3.0e38f;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sp[(long long)j * H]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) sum += __expf(sp[(long long)j * H] - mx);
  sum = eqv2_warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < n; j += 32) pp[(long long)j * H] = __expf(sp[(long long)j * H] - mx) * inv;
}

// gS = P (gP - sum_j P gP)
__global__ void pair_softmax_bwd_kernel(const float* __restrict__ Pw, const float* __restrict__ gP,
                                        const int* __restrict__ gcount, const long long* __restrict__ rowptr,
                                        float* __restrict__ gS, long long N, int H) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= N * H) return;
  const long long i = w / H;
  const int h = (int)(w - i * H), n = gcount[i];
  const long long base = rowptr[i] * H + h;
  float r = 0.f;
  for (int j = lane; j < n; j += 32) r = fmaf(Pw[base + (long long)j * H], gP[base + (long long)j * H], r);
  r = eqv2_warp_sum(r);
  for (int j = lane; j < n; j += 32) gS[base + (long long)j * H] = Pw[base + (long long)j * H] * (gP[base + (long long)j * H] - r);
}

// derivative of the backward: cotangent u of gS = P (gP - r), r = sum P gP:
//   d/dgP = P (u - q)                 q = sum_j u P
//   d/dP  = u (gP - r) - gP q
__global__ void pair_softmax_bwd2_kernel(const float* __restrict__ Pw, const float* __restrict__ gP, const float* __restrict__ u,
                                         const int* __restrict__ gcount, const long long* __restrict__ rowptr,
                                         float* __restrict__ dP, float* __restrict__ dgP, long long N, int H) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= N * H) return;
  const long long i = w / H;
  const int h = (int)(w - i * H), n = gcount[i];
  const long long base = rowptr[i] * H + h;
  float r = 0.f, q = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float p = Pw[base + (long long)j * H];
    r = fmaf(p, gP[base + (long long)j * H], r);
    q = fmaf(p, u[base + (long long)j * H], q);
  }
  r = eqv2_warp_sum(r);
  q = eqv2_warp_sum(q);
  for (int j = lane; j < n; j += 32) {
    const long long k = base + (long long)j * H;
    const float p = Pw[k], g = gP[k], uu = u[k];
    dgP[k] = p * (uu - q);
    dP[k] = uu * (g - r) - g * q;
  }
}

int check_shape(const char* who, int M, int H, int D) {
  EQV2_REQUIRE(M > 0 && H > 0 && D > 0 && D % 4 == 0 && ((D / 4) & (D / 4 - 1)) == 0 && D / 4 <= 32,
               "%s: head width %d must be 4 x a power of two <= 128 (M = %d, H = %d)", who, D, M, H);
  return 0;
}

}  // namespace

extern "C" int eqv2_pair_scores(const float* a, const float* b, const int* gstart, const int* gcount,
                                const long long* rowptr, float* S, long long N, int M, int H, int D, float scale,
                                void* stream) {
  if (N == 0) return 0;
  if (check_shape("eqv2_pair_scores", M, H, D)) return 1;
  const size_t smem = (size_t)M * H * D * sizeof(float);
  EQV2_REQUIRE(smem <= 200 * 1024, "eqv2_pair_scores: a row of %zu B does not fit shared memory", smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute((const void*)pair_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    EQV2_REQUIRE(e == cudaSuccess, "eqv2_pair_scores: cannot reserve %zu B of shared memory", smem);
  }
  EQV2_LAUNCH(pair_scores_kernel, dim3((unsigned)N), dim3(PA_THREADS), smem, stream, a, b, gstart, gcount, rowptr, S, M, H, D, scale);
  EQV2_CHECK_LAUNCH("eqv2_pair_scores");
  return 0;
}

extern "C" int eqv2_pair_mix(const float* W, const float* b, const int* gstart, const int* gcount, const long long* rowptr,
                             float* out, long long N, int M, int H, int D, int max_count, int transpose, void* stream) {
  if (N == 0) return 0;
  if (check_shape("eqv2_pair_mix", M, H, D)) return 1;
  const size_t smem = (size_t)max_count * H * sizeof(float);
  EQV2_REQUIRE(smem <= 200 * 1024, "eqv2_pair_mix: %d atoms x %d heads do not fit shared memory", max_count, H);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute((const void*)pair_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    EQV2_REQUIRE(e == cudaSuccess, "eqv2_pair_mix: cannot reserve %zu B of shared memory", smem);
  }
  EQV2_LAUNCH(pair_mix_kernel, dim3((unsigned)N), dim3(PA_THREADS), smem, stream, W, b, gstart, gcount, rowptr, out, M, H, D, transpose);
  EQV2_CHECK_LAUNCH("eqv2_pair_mix");
  return 0;
}

extern "C" int eqv2_pair_softmax_fwd(const float* S, const int* gcount, const long long* rowptr, float* Pw, long long N,
                                     int H, void* stream) {
  if (N == 0) return 0;
  const long long warps = N * H;
  EQV2_LAUNCH(pair_softmax_fwd_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, stream, S, gcount, rowptr, Pw, N, H);
  EQV2_CHECK_LAUNCH("eqv2_pair_softmax_fwd");
  return 0;
}

extern "C" int eqv2_pair_softmax_bwd(const float* Pw, const float* gP, const int* gcount, const long long* rowptr, float* gS,
                                     long long N, int H, void* stream) {
  if (N == 0) return 0;
  const long long warps = N * H;
  EQV2_LAUNCH(pair_softmax_bwd_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, stream, Pw, gP, gcount, rowptr, gS, N, H);
  EQV2_CHECK_LAUNCH("eqv2_pair_softmax_bwd");
  return 0;
}

extern "C" int eqv2_pair_softmax_bwd2(const float* Pw, const float* gP, const float* u, const int* gcount,
                                      const long long* rowptr, float* dP, float* dgP, long long N, int H, void* stream) {
  if (N == 0) return 0;
  const long long warps = N * H;
  EQV2_LAUNCH(pair_softmax_bwd2_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, stream, Pw, gP, u, gcount, rowptr, dP, dgP, N, H);
  EQV2_CHECK_LAUNCH("eqv2_pair_softmax_bwd2");
  return 0;
}
