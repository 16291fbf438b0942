// Grouped GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM)
// for the dense contraction of the path: the SO(2) convolution blocks (so2_ops.py:150-185), their
// dgrad / wgrad, and the radial-MLP linears (radial_function.py:29).
//
//   C[M,N] (+)= opA[M,K] * opB[K,N] (+ bias[N])         fp32 in, fp32 out
//
// Precision modes
//   mode 0 "3xTF32": every operand value v is split into hi = v rounded to tf32 and lo = v - hi (exact in fp32; the
//           tensor core reads lo's upper 19 bits); the tensor core computes hi*hi, lo*hi and hi*lo (the dropped
//           lo*lo term and the truncation of lo are O(2^-21) relative).
//           MEASURED on B200 (scripts/gemm_accuracy.py): every tcgen05.mma accumulation into TMEM truncates
//           (round-toward-zero at 24 bits, ~5e-8 relative per instruction, systematic), so a plain K-long
//           chain is 1e-5 off at K ~ 2000.  The kernel therefore keeps the hi*hi chain short: it accumulates
//           CHUNK_KB k-blocks (16 instructions) into one of two ping-pong TMEM accumulators and the worker warps
//           add each finished chunk into fp32 registers (round-to-nearest) while the tensor core works on the
//           other buffer.  The small lo terms use a third TMEM accumulator for the whole K (their truncation is
//           2^-11 times smaller).  Result: 4e-7 max-normalised error at K = 2048..8192 -- fp32-class accuracy, as
//           the 1e-5 parity bound of the fp32 mode needs (the FFMA engine and cuBLAS fp32 give 5e-7..2.7e-6).
//   mode 1 "1xTF32": raw fp32 tiles straight to the tensor core, one TMEM chain (separately stated tolerance).
//
// CTA = one 128 x 128 output tile, 4-stage mbarrier ring.  Warp roles (416 threads, 13 warps -> 128 regs/thread):
//   warps 0-3   producers: cp.async (LDGSTS, zero-filling out-of-range chunks) of the raw fp32 A and B tiles
//               (128 rows x 32 k each) straight into their final swizzled position, up to 4 k-blocks ahead.
//               Every 16-byte chunk is addressed individually, so K-contiguous ([rows,K]) sources land in the
//               K-major SWIZZLE_128B layout and row-contiguous ([K,rows]: dgrad weights, both wgrad operands)
//               sources in the MN-major SWIZZLE_128B_BASE32B layout (the only legal one for 32-bit MN-major
//               operands) -- no transposed copies in HBM -- with ragged M/N/K tails zero-filled; two-level row
//               strides address the degree slabs of [N,K,C] node tensors.
//   warps 4-11  workers, two groups of four on ALTERNATE k-blocks:
//               B: split the landed tile in shared memory (hi in place, lo tile);
//               A: shared memory -> registers -> hi/lo -> TENSOR MEMORY (tcgen05.st): the A operand of all three
//                  MMAs is read from TMEM, so no A_hi/A_lo tiles are written to or re-read from shared memory
//                  (shared-memory bandwidth was the first bound: 224 KB of traffic per k-block in r01 v3);
//               promotion of finished hi*hi chunks (tcgen05.ld -> register accumulators) and the epilogue
//               (bias / accumulate / split-K atomics -> global).
//   warp  12    TMEM allocation (512 columns: D_hi[2] | D_lo | A ring 2 x (hi 32 | lo 32)) + single-thread
//               tcgen05.mma issue (A from TMEM, B from smem descriptors) + tcgen05.commit.
// Barriers: raw_full[s] (cp.async completion -> workers), full[s] (workers -> MMA), empty[s] (commit ->
// producers), a_empty[g] (commit -> workers: TMEM A slot free), acc_full[b] / acc_empty[b] (promotion).
#include "common.cuh"

#ifndef EQV2_CPU_EMU

namespace {

constexpr int BM = 128, BN = 128, BK = 32;         // BK fp32 = 128 bytes = one swizzle row
constexpr int STAGES = 4;
constexpr int TILE_BYTES = BM * BK * 4;            // 16 KB per operand tile
constexpr int STAGE_BYTES = 3 * TILE_BYTES;        // A raw | B hi (raw lands here) | B lo
constexpr int A_SLOTS = 2;                         // TMEM ring of split A tiles: per slot hi (32 columns) + lo (32)
// 13 warps = (4,3,3,3) per scheduler partition -> 128 registers per thread without spills
constexpr int NUM_PRODUCER_WARPS = 4;              // warps 0-3
constexpr int NUM_EPI_WARPS = 8;                   // warps 4-11 (lane quarter = warp & 3)
constexpr int MMA_WARP = NUM_PRODUCER_WARPS + NUM_EPI_WARPS;
constexpr int NUM_THREADS = (NUM_PRODUCER_WARPS + NUM_EPI_WARPS + 1) * 32;
constexpr int TMEM_COLS = 512;                     // D_hi[0] | D_hi[1] | D_lo (128 columns each) | A ring (2 x 64)
constexpr int PROMOTE_LAG = 3;                     // blocks between the split of a chunk's last stage and its promotion
constexpr int CHUNK_KB = 4;                        // k-blocks per promoted hi*hi chain (16 tcgen05.mma)
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct TcGroup {
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  int M, N, K;
  long long lda, ldb, ldc;
  // two-level addressing of the NON-contiguous index r of an operand (degree slabs of [N,K,C] node tensors):
  // offset(r) = (r / rpb) * bs + (r % rpb) * ld ; rpb == 0 means plain r * ld
  int a_rpb, b_rpb, c_rpb;
  long long a_bs, b_bs, c_bs;
  int a_mn;        // 1: A stored [K,M] (row index contiguous)   0: [M,K]
  int b_mn;        // 1: B stored [K,N]                          0: [N,K]
  int accumulate;  // C += result
  int tiles_m, tiles_n;
  int tile_start;  // first linear tile of this group
};

struct TcParams {
  TcGroup g[EQV2_GEMM_MAX_GROUPS];
  int ngroups;
  int split_k;
  int mode;
};

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n\t}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// A operand from tensor memory (128 lanes = rows, 8 columns = the K=8 slice), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
// 32 lanes x 16 columns: lane i of the warp writes its 16 registers to columns [col, col+16) of TMEM lane (base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t o;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o) : "f"(v));
  return __uint_as_float(o);
}

// ---- descriptors ---------------------------------------------------------------------------------
// SM100 shared-memory matrix descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
// | layout type [61,64).
// K-major tile [rows][32 fp32], SWIZZLE_128B (type 2): 16-byte chunk index ^= (row & 7); 8-row groups are
//   SBO = 1024 B apart; LBO unused (1).
// MN-major tile (the row index is the contiguous one).  For 32-bit operands the only legal swizzle is
//   SWIZZLE_128B_BASE32B (type 1): atom = 32 rows (128 B) x 4 k, 32-byte chunk index ^= (k & 3).
//   Tile = [k/4 (8)][rows/32 (4)][k%4][128 B]  ->  LBO (next 32 rows) = 512 B, SBO (next 4 k) = 2048 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, bool mn_major) {
  const uint64_t lbo = mn_major ? (512u >> 4) : 1u;
  const uint64_t sbo = mn_major ? (2048u >> 4) : (1024u >> 4);
  const uint64_t type = mn_major ? 1ull : 2ull;
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (type << 61);
}
// instruction descriptor: D fp32 (1<<4), A/B tf32 (2<<7, 2<<10), majors (bits 15/16), N>>3 (17..22), M>>4 (24..28)
__device__ __forceinline__ uint32_t make_idesc(bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ---- producer: one operand tile (128 rows x 32 k) global -> smem with cp.async (LDGSTS) ------------------
// tid in [0,128).  src element (row, k) = mn ? src[k*ld + row] : src[row*ld + k].  Out-of-range chunks are
// zero-filled through the src-size operand.  The raw fp32 tile lands in the `hi` buffer in its final layout.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ long long strided_off(int r, int rpb, long long bs, long long ld) {
  if (rpb == 0) return (long long)r * ld;
  const int q = r / rpb;
  return (long long)q * bs + (long long)(r - q * rpb) * ld;
}

// Per-thread view of its 8 chunks of one operand tile: everything that does not change along K.
struct ChunkPlan {
  long long rowoff[8];   // element offset of the chunk's row part (-1: row out of range)
  uint32_t off[8];       // byte offset inside the tile (final swizzled position)
  uint32_t bytes[8];     // 16, or fewer for a ragged row tail of a row-contiguous source
  int kk[8];             // k of the chunk inside the k-block (first of 4 for K-contiguous sources)
};

__device__ __forceinline__ void plan_tile(ChunkPlan& P, long long ld, int rpb, long long bs, int mn, int row0, int rows,
                                          int tid) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * 128 + tid;
    if (!mn) {
      const int r = idx >> 3, c = idx & 7;                  // 8 x 16 B chunks per row
      const int gr = row0 + r;
      P.rowoff[it] = (gr < rows) ? strided_off(gr, rpb, bs, ld) : -1;
      P.bytes[it] = 16;
      P.kk[it] = c * 4;
      P.off[it] = (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
    } else {
      const int kk = idx >> 5, c32 = idx & 31;              // 32 x 16 B chunks (128 rows) per k
      const int gr = row0 + c32 * 4;
      P.rowoff[it] = (gr < rows) ? (long long)gr : -1;
      P.bytes[it] = (uint32_t)max(0, min(16, (rows - gr) * 4));
      P.kk[it] = kk;
      const int rb = c32 >> 3, c = c32 & 7, k4 = kk & 3, kg = kk >> 2;
      P.off[it] = (uint32_t)(kg * 2048 + rb * 512 + k4 * 128 + (((((c >> 1) ^ k4) << 1) | (c & 1)) << 4));
    }
  }
}

__device__ __forceinline__ void issue_tile(const ChunkPlan& P, const float* __restrict__ base, long long ld, int rpb,
                                           long long bs, int mn, int k0, int K, uint32_t hi) {
  if (!mn || rpb == 0) {          // common case: the k offset is a plain multiple -> no division in the loop
    const long long kmul = mn ? ld : 1;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int gk = k0 + P.kk[it];
      const bool ok = (P.rowoff[it] >= 0) && (gk < K);
      cp_async16(hi + P.off[it], ok ? base + P.rowoff[it] + (long long)gk * kmul : base, ok ? P.bytes[it] : 0u);
    }
  } else {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int gk = k0 + P.kk[it];
      const bool ok = (P.rowoff[it] >= 0) && (gk < K);
      cp_async16(hi + P.off[it], ok ? base + P.rowoff[it] + strided_off(gk, rpb, bs, ld) : base, ok ? P.bytes[it] : 0u);
    }
  }
}

// hi/lo split of a landed tile, elementwise and therefore layout-agnostic: chunk i of `hi` <-> chunk i of `lo`.
// hi = v rounded to tf32 (ties away: add half an ulp to the magnitude bits, clear the low 13 bits),
// lo = v - hi exactly (fp32); the tensor core reads lo's upper 19 bits, i.e. drops O(2^-21 |v|).
// 128 threads of one worker group, 4 chunks each over an 8 KB half tile (called twice per tile).
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ void split_tile(unsigned char* hi, unsigned char* lo, int wt) {
  float4 v[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) v[it] = *reinterpret_cast<const float4*>(hi + ((uint32_t)(it * 128 + wt) << 4));
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const uint32_t off = (uint32_t)(it * 128 + wt) << 4;
    float4 h, l;
    h.x = tf32_hi(v[it].x); h.y = tf32_hi(v[it].y); h.z = tf32_hi(v[it].z); h.w = tf32_hi(v[it].w);
    l.x = v[it].x - h.x; l.y = v[it].y - h.y; l.z = v[it].z - h.z; l.w = v[it].w - h.w;
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

// A tile (raw fp32 in shared memory, K-major SWIZZLE_128B or MN-major SWIZZLE_128B_BASE32B) -> hi / lo halves in
// TENSOR MEMORY: the worker thread of TMEM lane r reads 16 of row r's 32 k-values, splits them and stores
// hi to columns [k, k+16) and lo to columns [32 + k, ...) of the slot.  The tensor core then reads A from TMEM:
// no A_hi / A_lo tiles are written to or re-read from shared memory (the kernel is shared-memory-bandwidth bound).
__device__ __forceinline__ void split_a_to_tmem(const unsigned char* a_raw, int mn, int q, int half, int lane,
                                                uint32_t tmem_slot_addr) {
  const int r = q * 32 + lane;
  float v[16];
  if (!mn) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * half + j;
      const float4 f = *reinterpret_cast<const float4*>(a_raw + r * 128 + ((c ^ (r & 7)) << 4));
      v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
    }
  } else {
    const int c16 = lane >> 2;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int kk = 16 * half + t, k4 = kk & 3, kg = kk >> 2;
      const int cs = ((((c16 >> 1) ^ k4) << 1) | (c16 & 1));
      v[t] = *reinterpret_cast<const float*>(a_raw + kg * 2048 + q * 512 + k4 * 128 + cs * 16 + (lane & 3) * 4);
    }
  }
  uint32_t h[16], l[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const float hv = tf32_hi(v[t]);
    h[t] = __float_as_uint(hv);
    l[t] = __float_as_uint(v[t] - hv);
  }
  const uint32_t ta = tmem_slot_addr + ((uint32_t)(q * 32) << 16) + (uint32_t)(16 * half);
  tmem_st16(ta, h);
  tmem_st16(ta + 32u, l);     // the caller issues tcgen05.wait::st once for the whole tile
}

__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  unsigned char* tiles = smem_raw + pad;                                    // STAGES * STAGE_BYTES, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * STAGE_BYTES);
  // bars: [0,S) full | [S,2S) empty | [2S,3S) raw_full | 3S+b acc_full | 3S+2+b acc_empty | 3S+4+a a_empty ; TMEM base
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* raw_full = bars + 2 * STAGES;
  uint64_t* acc_full = bars + 3 * STAGES;
  uint64_t* acc_empty = bars + 3 * STAGES + 2;
  uint64_t* a_empty = bars + 3 * STAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4 + A_SLOTS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- which tile ----
  const int tile_lin = blockIdx.x / P.split_k;
  const int ks = blockIdx.x % P.split_k;
  int gi = 0;
#pragma unroll 1
  for (int i = 1; i < P.ngroups; ++i)
    if (tile_lin >= P.g[i].tile_start) gi = i;
  const TcGroup& G = P.g[gi];
  const int t = tile_lin - G.tile_start;
  // n fastest: CTAs that run together share the (large, streamed) A row panel through L2; the B panels are
  // weights and stay L2-resident anyway (ncu r01c: m-fastest re-read A 5x from DRAM)
  const int tn = t % G.tiles_n, tm = t / G.tiles_n;
  const int m0 = tm * BM, n0 = tn * BN;
  const int nkb_all = (G.K + BK - 1) / BK;
  const int per = (nkb_all + P.split_k - 1) / P.split_k;
  const int kb0 = ks * per, kb1 = min(nkb_all, kb0 + per);
  const int nkb = max(0, kb1 - kb0);
  const bool want_lo = ((P.mode & 1) == 0);
  // bit 1 of mode (diagnostics): one long TMEM chain even in 3xTF32 mode -> isolates the cost of promotion
  const int chunk = (want_lo && !(P.mode & 2)) ? CHUNK_KB : max(nkb, 1);
  const int nchunks = (nkb + chunk - 1) / chunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), NUM_EPI_WARPS / 2);             // one arrive per warp of the owning worker group
      mbar_init(smem_u32(&empty[s]), 1);
      mbar_init(smem_u32(&raw_full[s]), NUM_PRODUCER_WARPS * 32);   // one cp.async completion arrive per thread
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&acc_full[b]), 1);
      mbar_init(smem_u32(&acc_empty[b]), NUM_EPI_WARPS);
      mbar_init(smem_u32(&a_empty[b]), 1);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_lo = tmem_base + 2 * BN;
  const uint32_t tmem_a = tmem_base + 3 * BN;            // + slot * 64: hi [0,32), lo [32,64)

  if (warp < NUM_PRODUCER_WARPS) {
    // ================= producers: cp.async only, up to STAGES blocks ahead of the tensor core =================
    const int tid = threadIdx.x;          // 0..127
    ChunkPlan pa, pb;
    plan_tile(pa, G.lda, G.a_rpb, G.a_bs, G.a_mn, m0, G.M, tid);
    plan_tile(pb, G.ldb, G.b_rpb, G.b_bs, G.b_mn, n0, G.N, tid);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES, round = i / STAGES;
      mbar_wait(smem_u32(&empty[s]), (uint32_t)((round & 1) ^ 1));
      const uint32_t st = smem_u32(tiles + (size_t)s * STAGE_BYTES);
      const int k0 = (kb0 + i) * BK;
      issue_tile(pa, G.A, G.lda, G.a_rpb, G.a_bs, G.a_mn, k0, G.K, st);
      issue_tile(pb, G.B, G.ldb, G.b_rpb, G.b_bs, G.b_mn, k0, G.K, st + TILE_BYTES);
      cp_async_arrive_noinc(smem_u32(&raw_full[s]));
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(G.a_mn != 0, G.b_mn != 0);
      const uint32_t idesc_ts = make_idesc(false, G.b_mn != 0);   // A from TMEM is always K-major (row per lane)
      const uint32_t a_step = G.a_mn ? 4096u : 32u;       // bytes per K=8 slice (MN-major: two 4-k groups)
      const uint32_t b_step = G.b_mn ? 4096u : 32u;
      int i = 0;
      for (int j = 0; j < nchunks; ++j) {
        const int b = j & 1;
        mbar_wait(smem_u32(&acc_empty[b]), (uint32_t)(((j >> 1) & 1) ^ 1));   // promoters drained this buffer
        tc_fence_after();
        const uint32_t d_hi = tmem_base + (uint32_t)(b * BN);
        const int iend = min(nkb, i + chunk);
        for (int i0 = i; i < iend; ++i) {
          const int s = i % STAGES, round = i / STAGES;
          if (want_lo) {
            mbar_wait(smem_u32(&full[s]), (uint32_t)(round & 1));
          } else {                                         // 1xTF32: the raw fp32 tile feeds the tensor core as is
            mbar_wait(smem_u32(&raw_full[s]), (uint32_t)(round & 1));
            fence_proxy_async_smem();
          }
          tc_fence_after();
          const uint32_t base = smem_u32(tiles + (size_t)s * STAGE_BYTES);
          const uint32_t a_raw = base, b_hi = base + TILE_BYTES, b_lo = base + 2 * TILE_BYTES;
          if (want_lo) {
            // A (hi | lo) from tensor memory slot i % 2, B (hi, lo) from shared memory
            const uint32_t ah = tmem_a + (uint32_t)((i % A_SLOTS) * 64), al = ah + 32u;
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint64_t dbh = make_desc(b_hi + k * b_step, G.b_mn != 0);
              const uint64_t dbl = make_desc(b_lo + k * b_step, G.b_mn != 0);
              umma_tf32_ts(d_hi, ah + 8u * k, dbh, idesc_ts, (i > i0 || k > 0) ? 1u : 0u);
              umma_tf32_ts(tmem_lo, al + 8u * k, dbh, idesc_ts, (i > 0 || k > 0) ? 1u : 0u);
              umma_tf32_ts(tmem_lo, ah + 8u * k, dbl, idesc_ts, 1u);
            }
            umma_commit(smem_u32(&a_empty[i % A_SLOTS]));  // the A slot may be overwritten when these retire
          } else {
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint64_t da = make_desc(a_raw + k * a_step, G.a_mn != 0);
              const uint64_t db = make_desc(b_hi + k * b_step, G.b_mn != 0);
              umma_tf32(d_hi, da, db, idesc, (i > i0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&empty[s]));                // frees the stage when these MMAs retire
        }
        umma_commit(smem_u32(&acc_full[b]));               // chunk (and everything before it) complete
      }
    }
    __syncwarp();
  } else if (nkb > 0) {
    // ================= workers: hi/lo split of landed stages + promotion of finished chunks + epilogue ==========
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    const int half = (warp - NUM_PRODUCER_WARPS) >> 2;     // column half
    const int wt = threadIdx.x - NUM_PRODUCER_WARPS * 32;  // 0..255
    const uint32_t lane_col = ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
    float acc[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = 0.f;
    auto promote = [&](int j) {
      const int b = j & 1;
      mbar_wait(smem_u32(&acc_full[b]), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      uint32_t r[32];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        tmem_ld32(tmem_base + (uint32_t)(b * BN) + lane_col + (uint32_t)(cc * 32), r);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[cc * 32 + c] += __uint_as_float(r[c]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[b]));
    };
    int promoted = 0;
    if (want_lo) {
      // The split of one k-block is a chain of long-latency steps (LDS -> STTM -> wait::st -> LDS -> STS -> proxy
      // fence -> arrive; ncu r01 v4: the MMA thread waited on `full` 140 polls per block).  Two groups of four warps
      // therefore work on ALTERNATE k-blocks (group g owns blocks i = g mod 2 and TMEM A slot g), so two chains are
      // in flight; both groups promote every finished chunk (their own 64 accumulator columns).
      const int grp = half, gt = wt & 127;
      for (int i = grp; i < nkb; i += 2) {
        const int s = i % STAGES, round = i / STAGES;
        mbar_wait(smem_u32(&raw_full[s]), (uint32_t)(round & 1));
        unsigned char* st = tiles + (size_t)s * STAGE_BYTES;
        // B first (it does not depend on the tensor core): split in place (hi) + lo tile, 128 threads x 8 chunks
        split_tile(st + TILE_BYTES, st + 2 * TILE_BYTES, gt);
        split_tile(st + TILE_BYTES + 8192, st + 2 * TILE_BYTES + 8192, gt);
        // A: shared memory (raw) -> registers -> hi/lo -> tensor memory slot g (free once block i-2 retired)
        mbar_wait(smem_u32(&a_empty[grp]), (uint32_t)((((i >> 1) & 1) ^ 1)));
        tc_fence_after();
        split_a_to_tmem(st, G.a_mn, q, 0, lane, tmem_a + (uint32_t)(grp * 64));
        split_a_to_tmem(st, G.a_mn, q, 1, lane, tmem_a + (uint32_t)(grp * 64));
        fence_proxy_async_smem();    // generic-proxy stores of B -> visible to the tensor core (async proxy)
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&full[s]));
        // promote chunks whose MMAs were issued >= PROMOTE_LAG blocks ago (they have very likely retired)
        while (promoted < nchunks && (promoted + 1) * chunk + PROMOTE_LAG <= i + 1) promote(promoted++);
      }
    }
    while (promoted < nchunks) promote(promoted++);
    if (want_lo) {
      uint32_t r[32];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        tmem_ld32(tmem_lo + lane_col + (uint32_t)(cc * 32), r);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[cc * 32 + c] += __uint_as_float(r[c]);
      }
    }
    const int row = m0 + q * 32 + lane;
    const bool atomic = P.split_k > 1;
    if (row < G.M) {
      const int col0 = half * 64;
      float* crow = G.C + strided_off(row, G.c_rpb, G.c_bs, G.ldc) + n0 + col0;
      const int ncol = min(64, G.N - n0 - col0);
      const bool vec = (ncol == 64) && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0) && !atomic;
      if (vec) {
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          float4 o = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
          if (G.bias != nullptr) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(G.bias + n0 + col0 + c));
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          if (G.accumulate) {
            const float4 p = *reinterpret_cast<const float4*>(crow + c);
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          *reinterpret_cast<float4*>(crow + c) = o;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          if (c < ncol) {
            float o = acc[c];
            if (G.bias != nullptr && ks == 0) o += __ldg(G.bias + n0 + col0 + c);
            if (atomic) atomicAdd(crow + c, o);
            else if (G.accumulate) crow[c] += o;
            else crow[c] = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

extern "C" int eqv2_gemm_tc(const eqv2_gemm_desc* descs, int ngroups, int split_k, int mode, void* stream) {
  EQV2_REQUIRE(ngroups >= 1 && ngroups <= EQV2_GEMM_MAX_GROUPS, "eqv2_gemm_tc: ngroups=%d out of range", ngroups);
  EQV2_REQUIRE(split_k >= 1 && mode >= 0 && mode <= 3, "eqv2_gemm_tc: bad split_k/mode");
  TcParams P;
  memset(&P, 0, sizeof(P));
  int tiles = 0;
  for (int i = 0; i < ngroups; ++i) {
    const eqv2_gemm_desc& d = descs[i];
    EQV2_REQUIRE(d.A && d.B && d.C, "eqv2_gemm_tc: null operand in group %d", i);
    EQV2_REQUIRE((d.a_ld % 4) == 0 && (d.b_ld % 4) == 0 && (((uintptr_t)d.A | (uintptr_t)d.B) & 15) == 0,
                 "eqv2_gemm_tc: operands must be 16-byte aligned with leading dimensions multiple of 4");
    EQV2_REQUIRE((d.a_rpb >= (1ll << 31) || d.a_bs % 4 == 0) && (d.b_rpb >= (1ll << 31) || d.b_bs % 4 == 0),
                 "eqv2_gemm_tc: block strides of two-level operands must be multiples of 4");
    // the contiguous dimension is read in float4 units
    EQV2_REQUIRE(d.transA ? true : (d.K % 4) == 0, "eqv2_gemm_tc: K must be a multiple of 4 for K-contiguous A");
    EQV2_REQUIRE(d.transB ? (d.K % 4) == 0 : true, "eqv2_gemm_tc: K must be a multiple of 4 for K-contiguous B");
    TcGroup& g = P.g[i];
    g.A = d.A; g.B = d.B; g.C = d.C; g.bias = d.bias;
    g.M = d.M; g.N = d.N; g.K = d.K;
    g.lda = d.a_ld; g.ldb = d.b_ld; g.ldc = d.c_ld;
    g.a_rpb = d.a_rpb >= (1ll << 31) ? 0 : (int)d.a_rpb; g.a_bs = d.a_bs;
    g.b_rpb = d.b_rpb >= (1ll << 31) ? 0 : (int)d.b_rpb; g.b_bs = d.b_bs;
    g.c_rpb = d.c_rpb >= (1ll << 31) ? 0 : (int)d.c_rpb; g.c_bs = d.c_bs;
    g.a_mn = d.transA ? 1 : 0;
    g.b_mn = d.transB ? 0 : 1;
    g.accumulate = d.accumulate;
    g.tiles_m = (d.M + BM - 1) / BM;
    g.tiles_n = (d.N + BN - 1) / BN;
    g.tile_start = tiles;
    tiles += g.tiles_m * g.tiles_n;
  }
  if (tiles == 0) return 0;
  P.ngroups = ngroups;
  P.split_k = split_k;
  P.mode = mode;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    EQV2_REQUIRE(e == cudaSuccess, "eqv2_gemm_tc: cannot reserve %zu B of shared memory: %s", SMEM_BYTES,
                 cudaGetErrorString(e));
    attr_set = true;
  }
  gemm_tc_kernel<<<dim3((unsigned)(tiles * split_k)), dim3(NUM_THREADS), SMEM_BYTES, (cudaStream_t)stream>>>(P);
  EQV2_CHECK_LAUNCH("eqv2_gemm_tc");
  return 0;
}

#endif  // EQV2_CPU_EMU
