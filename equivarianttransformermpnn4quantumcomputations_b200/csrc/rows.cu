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
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;          // 4 independent chains hide the load latency; fixed order
  long long i = beg + s;
  for (; i + 3ll * S < end; i += 4ll * S) {
    const long long r0 = perm != nullptr ? perm[i] : i, r1 = perm != nullptr ? perm[i + S] : i + S;
    const long long r2 = perm != nullptr ? perm[i + 2ll * S] : i + 2ll * S, r3 = perm != nullptr ? perm[i + 3ll * S] : i + 3ll * S;
    a0 += __ldg(src + r0 * ld + c);
    a1 += __ldg(src + r1 * ld + c);
    a2 += __ldg(src + r2 * ld + c);
    a3 += __ldg(src + r3 * ld + c);
  }
  for (; i < end; i += S) a0 += __ldg(src + (perm != nullptr ? (long long)perm[i] : i) * ld + c);
  partial[((long long)v * S + s) * C + c] = (a0 + a1) + (a2 + a3);
}

__global__ void seg_colsum_final_kernel(const float* __restrict__ partial, int C, int S, float* __restrict__ out) {
  const int v = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += partial[((long long)v * S + s) * C + c];
  out[(long long)v * C + c] = a;
}

}  // namespace

extern "C" int eqv2_embed_rows(const float* table, const long long* idx, float* out, long long E, int C, void* stream) {
  if (E == 0 || C == 0) return 0;
  const int tx = C >= 128 ? 128 : 32, ty = 256 / tx;
  EQV2_LAUNCH(embed_rows_kernel, dim3((unsigned)((E + ty - 1) / ty)), dim3(tx, ty), 0, stream, table, idx, out, E, C);
  EQV2_CHECK_LAUNCH("eqv2_embed_rows");
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
  EQV2_LAUNCH(seg_colsum_final_kernel, dim3(bx, V), dim3(tx), 0, stream, partial, C, S, out);
  EQV2_CHECK_LAUNCH("eqv2_seg_colsum (final)");
  return 0;
}
