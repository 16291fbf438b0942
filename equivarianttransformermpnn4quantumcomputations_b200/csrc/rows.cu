// Row gathers and deterministic segmented column sums: the edge-level embedding lookups
// (transformer_block.py:241-248, input_block.py:93-100: source_embedding(Z[src]), target_embedding(Z[dst])), their
// weight gradients, and the bias gradients of the dense layers (column sums over all edges).  They replace
// F.embedding's sort-based backward (8 radix-sort launches per call) and torch's strided column reduction.
//
//   embed_rows : out[e, :] = table[idx[e], :]
//   seg_colsum : out[v, c] = sum_{i in [rowptr[v], rowptr[v+1])} src[perm[i] * ld + c]     (perm / rowptr optional)
//                two stages, fixed summation order -> bit-reproducible: stage 1 gives every (segment, split s) the
//                rows i = beg + s, beg + s + S, ...; stage 2 adds the S partial rows in order.
#include "common.cuh"

namespace {

__global__ void embed_rows_kernel(const float* __restrict__ table, const long long* __restrict__ idx,
                                  float* __restrict__ out, long long E, int C) {
  const long long e = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  if (e >= E) return;
  const float* row = table + idx[e] * (long long)C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) out[e * C + c] = __ldg(row + c);
}

__global__ void seg_colsum_partial_kernel(const float* __restrict__ src, long long ld, const int* __restrict__ rowptr,
                                          const int* __restrict__ perm, long long rows, int C, int S,
                                          float* __restrict__ partial) {
  const int v = blockIdx.z, s = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long beg = rowptr != nullptr ? rowptr[v] : 0, end = rowptr != nullptr ? rowptr[v + 1] : rows;
  // 8 independent chains keep 8 loads in flight per thread (the kernel is latency-bound otherwise); fixed order
  float a[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) a[u] = 0.f;
  long long i = beg + s;
  for (; i + 7ll * S < end; i += 8ll * S) {
    long long r[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) r[u] = perm != nullptr ? (long long)perm[i + (long long)u * S] : i + (long long)u * S;
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] += __ldg(src + r[u] * ld + c);
  }
  for (; i < end; i += S) a[0] += __ldg(src + (perm != nullptr ? (long long)perm[i] : i) * ld + c);
  const float a0 = a[0] + a[4], a1 = a[1] + a[5], a2 = a[2] + a[6], a3 = a[3] + a[7];
  partial[((long long)v * S + s) * C + c] = (a0 + a1) + (a2 + a3);
}

// blockDim = (32 columns, 8 split lanes): lane y adds splits y, y + 8, ... (fixed order), the 8 partial sums meet in
// shared memory and are added in order
__global__ void seg_colsum_final_kernel(const float* __restrict__ partial, int C, int S, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int v = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (c < C)
    for (int s = threadIdx.y; s < S; s += 8) a += partial[((long long)v * S + s) * C + c];
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x];
    out[(long long)v * C + c] = t;
  }
}

// SO(2) convolution weights of order m > 0 in real 2x2 block form (so2_ops.py:53-61: out+ = x_r W_r - x_i W_i,
// out- = x_r W_i + x_i W_r):  B[2h, 2k] = [[Wr, -Wi], [Wi, Wr]] from fc.weight W[2h, k] = [Wr; Wi], and its adjoint
// gW = [gB00 + gB11; gB10 - gB01].  One launch each instead of 2 slices + neg + 3 cats (and, in backward, 3 narrow
// views + neg + 2 x (zero-fill + copy) + add): ~12 graph nodes per m-block and pass.
__global__ void so2_block_weight_kernel(const float* __restrict__ W, float* __restrict__ B, int h, int k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = 4ll * h * k;
  if (i >= total) return;
  const int r = (int)(i / (2 * k)), c = (int)(i % (2 * k));
  float v;
  if (r < h) v = (c < k) ? W[(long long)r * k + c] : -W[(long long)(h + r) * k + (c - k)];
  else v = (c < k) ? W[(long long)r * k + c] : W[(long long)(r - h) * k + (c - k)];
  B[i] = v;
}

__global__ void so2_block_weight_adj_kernel(const float* __restrict__ gB, float* __restrict__ gW, int h, int k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = 2ll * h * k;
  if (i >= total) return;
  const int r = (int)(i / k), c = (int)(i % k);
  const long long ld = 2ll * k;
  float v;
  if (r < h) v = gB[(long long)r * ld + c] + gB[(long long)(h + r) * ld + k + c];
  else v = gB[(long long)r * ld + c] - gB[(long long)(r - h) * ld + k + c];
  gW[i] = v;
}

}  // namespace

extern "C" int eqv2_so2_block_weight(const float* W, float* B, int h, int k, void* stream) {
  const long long total = 4ll * h * k;
  if (total == 0) return 0;
  EQV2_LAUNCH(so2_block_weight_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, W, B, h, k);
  EQV2_CHECK_LAUNCH("eqv2_so2_block_weight");
  return 0;
}

extern "C" int eqv2_so2_block_weight_adj(const float* gB, float* gW, int h, int k, void* stream) {
  const long long total = 2ll * h * k;
  if (total == 0) return 0;
  EQV2_LAUNCH(so2_block_weight_adj_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, gB, gW, h, k);
  EQV2_CHECK_LAUNCH("eqv2_so2_block_weight_adj");
  return 0;
}

extern "C" int eqv2_embed_rows(const float* table, const long long* idx, float* out, long long E, int C, void* stream) {
  if (E == 0 || C == 0) return 0;
  const int tx = C >= 128 ? 128 : 32, ty = 256 / tx;
  EQV2_LAUNCH(embed_rows_kernel, dim3((unsigned)((E + ty - 1) / ty)), dim3(tx, ty), 0, stream, table, idx, out, E, C);
  EQV2_CHECK_LAUNCH("eqv2_embed_rows");
  return 0;
}

// Per-graph stochastic depth (reference drop.py:16-27,49-68): out[n, :] = x[n, :] * scale[batch[n]],
// scale[g] = floor(keep + u[g]) * (1 / keep) -- the reference's `ones.div(keep) * (keep + rand).floor_()` gathered by `batch`
// and multiplied into the node tensor (seven element-wise launches per call), bit for bit; u is the reference's draw.
namespace {
__global__ void drop_path_scale_kernel(const float* __restrict__ x, const float* __restrict__ u,
                                       const long long* __restrict__ batch, float keep, float* __restrict__ out,
                                       long long N, long long row) {
  const float inv = 1.0f / keep;
  const long long total = N * row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / row;
    const float scale = inv * floorf(keep + __ldg(u + __ldg(batch + n)));
    out[i] = x[i] * scale;
  }
}
}  // namespace

extern "C" int eqv2_drop_path_scale(const float* x, const float* u, const long long* batch, float keep, float* out,
                                    long long N, long long row, void* stream) {
  if (N == 0 || row == 0) return 0;
  EQV2_REQUIRE(keep > 0.f && keep <= 1.f, "eqv2_drop_path_scale: keep probability %f out of (0, 1]", (double)keep);
  const long long total = N * row;
  const long long blocks = (total + 255) / 256;
  EQV2_LAUNCH(drop_path_scale_kernel, dim3((unsigned)(blocks < 148 * 16 ? blocks : 148 * 16)), dim3(256), 0, stream, x, u, batch,
              keep, out, N, row);
  EQV2_CHECK_LAUNCH("eqv2_drop_path_scale");
  return 0;
}

extern "C" int eqv2_seg_colsum(const float* src, long long ld, const int* rowptr, const int* perm, long long rows, int V,
                               int C, int S, float* partial, float* out, void* stream) {
  EQV2_REQUIRE(V >= 1 && S >= 1 && C >= 0, "eqv2_seg_colsum: bad V/S/C");
  EQV2_REQUIRE(V == 1 || rowptr != nullptr, "eqv2_seg_colsum: several segments need rowptr");
  if (C == 0) return 0;
  const int tx = 128, bx = (C + tx - 1) / tx;
  EQV2_LAUNCH(seg_colsum_partial_kernel, dim3(bx, S, V), dim3(tx), 0, stream, src, ld, rowptr, perm, rows, C, S, partial);
  EQV2_CHECK_LAUNCH("eqv2_seg_colsum (partial)");
  EQV2_LAUNCH(seg_colsum_final_kernel, dim3((C + 31) / 32, V), dim3(32, 8), 0, stream, partial, C, S, out);
  EQV2_CHECK_LAUNCH("eqv2_seg_colsum (final)");
  return 0;
}

#ifndef EQV2_CPU_EMU
// Column sums of a matrix that exists only as scaled fp16 hi/lo operand planes (bias gradient of a layer whose output
// gradient was written as planes by its producer kernel): out[c] = (1/s) sum_r (hi + lo)[r, col_off + c], s from the
// plane's bound slot.  Same two stages and fixed order as seg_colsum.
namespace {
__global__ void planes_colsum_partial_kernel(const __half* __restrict__ hi, long long plane, long long ld, long long rows,
                                             int C, int S, float* __restrict__ partial) {
  const int s = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  long long i = s;
  for (; i + 3ll * S < rows; i += 4ll * S) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long o = (i + (long long)u * S) * ld + c;
      a[u] += __half2float(hi[o]) + __half2float(hi[o + plane]);
    }
  }
  for (; i < rows; i += S) a[0] += __half2float(hi[i * ld + c]) + __half2float(hi[i * ld + c + plane]);
  partial[(long long)s * C + c] = (a[0] + a[1]) + (a[2] + a[3]);
}
// 16-byte loads: thread (cx, ry) = (tid % 16, tid / 16) sums 8 adjacent columns over the rows s + (ry + 8 k) S of slice s;
// the 8 row lanes meet in shared memory in fixed order (deterministic).  (The scalar version above read 2 bytes per thread
// and load: 2.9 TB/s on the [E, 4608] radial-weight gradient, 1.1 ms per OC20 step.)
__global__ void __launch_bounds__(128)
planes_colsum_partial_vec_kernel(const __half* __restrict__ hi, long long plane, long long ld, long long rows, int C, int S,
                                 float* __restrict__ partial) {
  __shared__ float red[8][16][9];
  const int s = blockIdx.y, cx = threadIdx.x & 15, ry = threadIdx.x >> 4;
  const int c = blockIdx.x * 128 + cx * 8;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    for (long long i = s + (long long)ry * S; i < rows; i += 8ll * S) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + i * ld + c));
      const uint4 l = __ldg(reinterpret_cast<const uint4*>(hi + i * ld + c + plane));
      const __half2* hp = reinterpret_cast<const __half2*>(&h);
      const __half2* lp = reinterpret_cast<const __half2*>(&l);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 x = __half22float2(hp[u]), y = __half22float2(lp[u]);
        a[2 * u] += x.x + y.x;
        a[2 * u + 1] += x.y + y.y;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) red[ry][cx][u] = a[u];
  __syncthreads();
  const int col = threadIdx.x;              // 128 columns of this block
  if (blockIdx.x * 128 + col < C) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][col >> 3][col & 7];
    partial[(long long)s * C + blockIdx.x * 128 + col] = t;
  }
}
__global__ void planes_colsum_final_kernel(const float* __restrict__ partial, int C, int S, const float* __restrict__ bound,
                                           float* __restrict__ out) {
  __shared__ float red[8][33];
  float sc, inv;
  eqv2_scale_of(eqv2_read_absmax(bound), sc, inv);
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (c < C)
    for (int s = threadIdx.y; s < S; s += 8) a += partial[(long long)s * C + c];
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x];
    out[c] = t * inv;
  }
}
}  // namespace

extern "C" int eqv2_planes_colsum(const void* planes, long long plane, long long ld, long long col_off, long long rows,
                                  int C, int S, const float* bound, float* partial, float* out, void* stream) {
  EQV2_REQUIRE(S >= 1 && C >= 0 && planes && bound && partial && out, "eqv2_planes_colsum: bad arguments");
  if (C == 0) return 0;
  const __half* hi = reinterpret_cast<const __half*>(planes) + col_off;
  if (C % 8 == 0 && col_off % 8 == 0 && ld % 8 == 0 && plane % 8 == 0 && (reinterpret_cast<uintptr_t>(planes) & 15) == 0) {
    EQV2_LAUNCH(planes_colsum_partial_vec_kernel, dim3((C + 127) / 128, S), dim3(128), 0, stream, hi, plane, ld, rows, C, S, partial);
  } else {
    EQV2_LAUNCH(planes_colsum_partial_kernel, dim3((C + 127) / 128, S), dim3(128), 0, stream, hi, plane, ld, rows, C, S, partial);
  }
  EQV2_CHECK_LAUNCH("eqv2_planes_colsum (partial)");
  EQV2_LAUNCH(planes_colsum_final_kernel, dim3((C + 31) / 32), dim3(32, 8), 0, stream, partial, C, S, bound, out);
  EQV2_CHECK_LAUNCH("eqv2_planes_colsum (final)");
  return 0;
}
#endif
