// Grouped GEMM on the 5th-generation tensor cores, second engine: fp32-class accuracy from THREE fp16 passes
// (tcgen05.mma kind::f16 runs at twice the kind::tf32 rate), operands pre-split by a streaming kernel and moved
// by TMA.  Same contraction as gemm_tc.cu (SO(2) convolution blocks so2_ops.py:150-185, their dgrad / wgrad,
// radial_function.py:29):
//
//   C[M,N] (+)= opA[M,K] * opB[K,N] (+ bias[N])         fp32 in, fp32 out
//
// 1. eqv2_split_f16: every fp32 operand tensor T is scaled by a power of two s_T (from its absolute maximum, so
//    that max |s_T v| lies in [2^14, 2^15)) and written as two fp16 planes  hi = fp16(s v),  lo = fp16(s v - hi)
//    into a zero-padded buffer [2][rows_pad][cols_pad].  hi + lo carries 22 significant bits of every value down
//    to 2^-18 of the tensor maximum and an absolute error <= 2^-40 max|T| below that (fp16 subnormals) -- the
//    max-normalised error of a dot product stays fp32-class.  One split serves every use of the tensor
//    (forward A operand and weight-gradient A^T operand read the same planes through different tensor maps).
// 2. eqv2_gemm_f16: persistent kernel, one CTA per SM looping over 128 x 128 output tiles (x split-K slices).
//      warp 0     TMA producer (one thread): per 64-wide k-block four tiles A_hi | A_lo | B_hi | B_lo (16 KB each,
//                 SWIZZLE_128B; K-contiguous sources as K-major tiles, row-contiguous sources as MN-major tiles,
//                 ragged M / N / K edges zero-filled by the tensor map) into a 3-stage mbarrier ring.
//      warp 1     TMEM allocation (512 columns = 2 x [D_hh 128 | D_lo 128]) + single-thread tcgen05.mma issue:
//                 per K=16 slice  A_hi x [B_hi | B_lo] (one N=256 instruction -> D_hh | D_lo) and
//                 A_lo x B_hi (N=128 -> D_lo): 2 instructions, 20 KB of shared-memory operand reads instead of the
//                 24 KB of three separate products.
//      warps 2-17 promotion + epilogue.  MEASURED on B200 (scripts/gemm_accuracy.py): every tcgen05.mma
//                 accumulation into TMEM truncates (~5e-8 relative per instruction, systematic), so the hi*hi
//                 chain is kept to CHUNK_KB k-blocks (16 instructions) in one of two ping-pong accumulator
//                 buffers; the workers add every finished chunk into fp32 registers (round to nearest) while the
//                 tensor core fills the other buffer.  The small lo terms stay in TMEM for the whole K (their
//                 truncation is 2^-11 times smaller) and are read with the last chunk of each buffer.  The
//                 accumulators live in registers, so the epilogue (scale 1/(s_A s_B), bias, accumulate or split-K
//                 atomics, store) of tile t overlaps the main loop of tile t+1.
#include "common.cuh"

#ifndef EQV2_CPU_EMU
#include <cuda.h>
#include <cuda_fp16.h>

namespace {

constexpr int BM = 128, BN = 128, BK = 64;         // BK fp16 = 128 bytes = one swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 2;            // 16 KB per operand plane tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;        // A_hi | A_lo | B_hi | B_lo
#ifndef EQV2_GEMM_WORKERS
#define EQV2_GEMM_WORKERS 16
#endif
// worker warps 2.. (TMEM lane quarter = warp & 3); each owns 32 rows x WCOLS columns of the tile.  16 workers (4 column
// quarters): the per-tile promotion + epilogue chain of a warp is half as long and twice as many warps hide its latencies
// (ncu r02: the workers ran at 6.2 cycles per instruction with 2 warps per scheduler, and their per-tile chain, not the
// MMAs or the stores, bounded the K = 128 launches).
constexpr int NUM_WORKER_WARPS = EQV2_GEMM_WORKERS;
constexpr int WCOLS = BN / (NUM_WORKER_WARPS / 4);          // 64 or 32 accumulator columns per worker thread
constexpr int PATCH_ROWS = 128 / NUM_WORKER_WARPS;          // 16 or 8 rows per transposition round
constexpr int NUM_THREADS = (2 + NUM_WORKER_WARPS) * 32;
constexpr int TMEM_COLS = 512;                     // buffer b: D_hh at b*256, D_lo at b*256 + 128
constexpr int CHUNK_KB = 4;                        // k-blocks per promoted hi*hi chain (16 tcgen05.mma)
constexpr int EPI_PATCH_BYTES = NUM_WORKER_WARPS * PATCH_ROWS * 36 * 4;   // per worker warp: a padded PATCH_ROWS x 32 fp32 transposition patch
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_PATCH_BYTES;

struct alignas(64) HGroup {
  CUtensorMap mapA, mapB;
  float* C;
  const float* bias;
  const float* a_absmax;
  const float* b_absmax;
  float* c_absmax;     // optional: receives max |C| of what this launch stores (64 slots, zero-initialised by the caller)
  long long ldc, c_bs;
  int c_rpb;           // C row r lives at (r / c_rpb) * c_bs + (r % c_rpb) * ldc (degree slabs of [N,K,C]); 0 = plain
  int M, N, K;
  int a_mn, b_mn;      // 1: the row index (M resp. N) is the contiguous one
  int accumulate;
  int tiles_m, tiles_n;
  int tile_start;
};

struct HParams {
  HGroup g[EQV2_GEMM_MAX_GROUPS];
  int ngroups;
  int split_k;
  int total_work;
  int passes;          // 3: hi*hi + hi*lo + lo*hi (fp32-class); 1: hi*hi only (single fp16 pass, reduced precision)
};

// ---- scale shared by the split kernel, the producer kernels and the GEMM epilogue: eqv2_scale_of (common.cuh) ----
__device__ __forceinline__ void scale_of(float amax, float& s, float& inv) { eqv2_scale_of(amax, s, inv); }

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one box of a 3-D tensor map (plane index = coordinate 2) -> shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------------------
// SM100 shared-memory matrix descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
// | layout type [61,64) (2 = SWIZZLE_128B: 16-byte chunk index ^= 128-byte row index & 7, as the TMA writes it).
// K-major tile [rows][64 fp16]: 8-row groups SBO = 1024 B apart, LBO unused; a K=16 slice is 32 B further.
// MN-major tile = TMA boxes of [64 k][64 rows] (8 KB each): 64-row blocks LBO = 8192 B apart, 8-k groups
//   SBO = 1024 B apart; a K=16 slice is 2048 B further.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, bool mn_major) {
  const uint64_t lbo = mn_major ? (8192u >> 4) : 1u;
  const uint64_t sbo = 1024u >> 4;
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32 (1<<4), A/B fp16 (0), majors (bits 15/16), N>>3 (17..22), M>>4 (24..28)
__device__ __forceinline__ uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

struct Work {
  int gi, m0, n0, ks, nkb, kb0;
};

__device__ __forceinline__ Work decode_work(const HParams& P, int w) {
  Work W;
  const int tile_lin = w / P.split_k;
  W.ks = w - tile_lin * P.split_k;
  int gi = 0;
#pragma unroll 1
  for (int i = 1; i < P.ngroups; ++i)
    if (tile_lin >= P.g[i].tile_start) gi = i;
  W.gi = gi;
  const HGroup& G = P.g[gi];
  const int t = tile_lin - G.tile_start;
  // n fastest: CTAs that run together share the (large, streamed) A row panel through L2
  const int tn = t % G.tiles_n, tm = t / G.tiles_n;
  W.m0 = tm * BM;
  W.n0 = tn * BN;
  const int nkb_all = (G.K + BK - 1) / BK;
  const int per = (nkb_all + P.split_k - 1) / P.split_k;
  W.kb0 = W.ks * per;
  W.nkb = max(0, min(nkb_all, W.kb0 + per) - W.kb0);
  return W;
}

// PASSES (3: fp32-class, 1: hi planes only) and CMAX (reduce max |C| in the epilogue) are compile-time: the single-thread
// TMA-issue and MMA-issue loops are latency-critical, and a run-time `passes` test inside them cost the three-pass kernel
// ~10 % (measured r02: conv1 forward 439 -> 383 TFLOP/s).
template <int PASSES, bool CMAX>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_f16_kernel(const __grid_constant__ HParams P) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  unsigned char* tiles = smem_raw + pad;                                    // STAGES * STAGE_BYTES, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * STAGE_BYTES);
  uint64_t* full = bars;                      // [STAGES]  TMA bytes landed
  uint64_t* empty = bars + STAGES;            // [STAGES]  MMAs that read the stage retired
  uint64_t* acc_full = bars + 2 * STAGES;     // [2]       chunk complete in TMEM buffer b
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]     workers drained buffer b
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* epi_patch = reinterpret_cast<float*>(tiles + (size_t)STAGES * STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&acc_full[b]), 1);
      mbar_init(smem_u32(&acc_empty[b]), NUM_WORKER_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (one thread) =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < P.total_work; w += gridDim.x) {
        const Work W = decode_work(P, w);
        const HGroup& G = P.g[W.gi];
        for (int i = 0; i < W.nkb; ++i, ++it) {
          const uint32_t s = it % STAGES, round = it / STAGES;
          mbar_wait(smem_u32(&empty[s]), (round & 1u) ^ 1u);
          const uint32_t bar = smem_u32(&full[s]);
          constexpr int nplanes = PASSES == 1 ? 1 : 2;
          mbar_expect_tx(bar, (uint32_t)(nplanes * 2 * TILE_BYTES));
          const uint32_t st = smem_u32(tiles + (size_t)s * STAGE_BYTES);
          const int k0 = (W.kb0 + i) * BK;
#pragma unroll
          for (int plane = 0; plane < nplanes; ++plane) {
            const uint32_t da = st + plane * TILE_BYTES, db = st + (2 + plane) * TILE_BYTES;
            if (!G.a_mn) {
              tma_load_3d(da, &G.mapA, bar, k0, W.m0, plane);
            } else {
              tma_load_3d(da, &G.mapA, bar, W.m0, k0, plane);
              tma_load_3d(da + 8192, &G.mapA, bar, W.m0 + 64, k0, plane);
            }
            if (!G.b_mn) {
              tma_load_3d(db, &G.mapB, bar, k0, W.n0, plane);
            } else {
              tma_load_3d(db, &G.mapB, bar, W.n0, k0, plane);
              tma_load_3d(db + 8192, &G.mapB, bar, W.n0 + 64, k0, plane);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      uint32_t it = 0, c = 0;
      for (int w = blockIdx.x; w < P.total_work; w += gridDim.x) {
        const Work W = decode_work(P, w);
        const HGroup& G = P.g[W.gi];
        const bool amn = G.a_mn != 0, bmn = G.b_mn != 0;
        const uint32_t idesc128 = make_idesc(128, amn, bmn), idesc256 = make_idesc(256, amn, bmn);
        const uint32_t a_step = amn ? 2048u : 32u;          // bytes per K=16 slice
        const uint32_t b_step = bmn ? 2048u : 32u;
        const int nchunks = (W.nkb + CHUNK_KB - 1) / CHUNK_KB;
        int i = 0;
        for (int j = 0; j < nchunks; ++j, ++c) {
          const uint32_t b = c & 1u;
          mbar_wait(smem_u32(&acc_empty[b]), ((c >> 1) & 1u) ^ 1u);      // workers drained this buffer
          tc_fence_after();
          const uint32_t d_hh = tmem_base + b * 256u, d_lo = d_hh + 128u;
          const int iend = min(W.nkb, i + CHUNK_KB);
          for (int i0 = i; i < iend; ++i, ++it) {
            const uint32_t s = it % STAGES, round = it / STAGES;
            mbar_wait(smem_u32(&full[s]), round & 1u);
            tc_fence_after();
            const uint32_t base = smem_u32(tiles + (size_t)s * STAGE_BYTES);
            const uint32_t a_hi = base, a_lo = base + TILE_BYTES, b_hi = base + 2 * TILE_BYTES, b_lo = base + 3 * TILE_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t dah = make_desc(a_hi + k * a_step, amn);
              const uint64_t dal = make_desc(a_lo + k * a_step, amn);
              const uint64_t dbh = make_desc(b_hi + k * b_step, bmn);     // as an N=256 operand: B_hi | B_lo
              if constexpr (PASSES == 1) {                                 // single pass: D_hh (+)= A_hi B_hi
                umma_f16(d_hh, dah, dbh, idesc128, (i == i0 && k == 0) ? 0u : 1u);
                continue;
              }
              if (i == i0 && k == 0) {
                // first slice of a chunk: D_hh restarts; D_lo restarts only on the first chunk of each buffer
                const uint64_t dbl = make_desc(b_lo + k * b_step, bmn);
                umma_f16(d_hh, dah, dbh, idesc128, 0u);
                umma_f16(d_lo, dah, dbl, idesc128, j < 2 ? 0u : 1u);
              } else {
                umma_f16(d_hh, dah, dbh, idesc256, 1u);                   // D_hh += A_hi B_hi, D_lo += A_hi B_lo
              }
              umma_f16(d_lo, dal, dbh, idesc128, 1u);                     // D_lo += A_lo B_hi
            }
            umma_commit(smem_u32(&empty[s]));                // frees the stage when these MMAs retire
          }
          umma_commit(smem_u32(&acc_full[b]));               // chunk (and everything before it) complete
        }
      }
    }
    __syncwarp();
  } else {
    // ================= workers: promotion of finished chunks + epilogue =================
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                      // column block (WCOLS wide) of this warp
    const uint32_t lane_col = ((uint32_t)(q * 32) << 16) + (uint32_t)(half * WCOLS);
    uint32_t c = 0;
    int scale_gi = -1;                                      // group whose 1/(s_A s_B) is cached in ia / ib
    float ia = 1.f, ib = 1.f;
    for (int w = blockIdx.x; w < P.total_work; w += gridDim.x) {
      const Work W = decode_work(P, w);
      const HGroup& G = P.g[W.gi];
      if (W.nkb == 0) continue;
      if (W.gi != scale_gi) {       // the operand maxima are read once per group, BEFORE the tile's promotions: four
        float sa, sb;               // dependent global loads per tile on the workers' path delayed the next promotion
        scale_of(eqv2_read_absmax(G.a_absmax), sa, ia);
        scale_of(eqv2_read_absmax(G.b_absmax), sb, ib);
        scale_gi = W.gi;
      }
      const int nchunks = (W.nkb + CHUNK_KB - 1) / CHUNK_KB;
      float acc[WCOLS];
#pragma unroll
      for (int x = 0; x < WCOLS; ++x) acc[x] = 0.f;
      for (int j = 0; j < nchunks; ++j, ++c) {
        const uint32_t b = c & 1u;
        mbar_wait(smem_u32(&acc_full[b]), (c >> 1) & 1u);
        tc_fence_after();
        const uint32_t d_hh = tmem_base + b * 256u + lane_col;
        uint32_t r0[32], r1[WCOLS > 32 ? 32 : 1];
        tmem_ld32_nowait(d_hh, r0);
        if constexpr (WCOLS > 32) tmem_ld32_nowait(d_hh + 32u, r1);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          acc[x] += __uint_as_float(r0[x]);
          if constexpr (WCOLS > 32) acc[32 + x] += __uint_as_float(r1[x]);
        }
        if (PASSES != 1 && j >= nchunks - 2) {               // last chunk on this buffer: its D_lo is final
          tmem_ld32_nowait(d_hh + 128u, r0);
          if constexpr (WCOLS > 32) tmem_ld32_nowait(d_hh + 160u, r1);
          tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 32; ++x) {
            acc[x] += __uint_as_float(r0[x]);
            if constexpr (WCOLS > 32) acc[32 + x] += __uint_as_float(r1[x]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_empty[b]));
      }
      // ---- epilogue (overlaps the next tile's main loop: TMEM is already released) ----
      const int row = W.m0 + q * 32 + lane;
      const bool atomic = P.split_k > 1;
      const float sc = ia * ib;
      float cmax = 0.f;                                     // max |C| of what this thread stores (c_absmax)
      // Fast path (plain row-major C, 16-byte aligned rows, N % 4 == 0): the accumulators are row-per-lane,
      // so a direct float4 store touches 32 different 128-byte lines per warp instruction (512 LSU tag cycles per warp and
      // tile, 4 096 per tile over the 8 worker warps: longer than the MMAs of a K = 128 tile).  Each warp instead passes
      // 16 x 32 blocks through a padded shared-memory patch and stores 4 rows x 128 bytes per instruction: same number
      // of store instructions, 8x fewer lines per instruction.  Split-K partial tiles take the same route with one
      // 16-byte vector reduction (red.global.add.v4.f32) per float4 instead of four scalar atomics on 32 lines each.
      const bool fast = (G.ldc & 3) == 0 && (G.c_bs & 3) == 0 && (G.N & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(G.C) & 15) == 0 && ((W.n0 & 3) == 0);
      if (fast) {
        float* patch = epi_patch + (warp - 2) * (PATCH_ROWS * 36);
        const int r16 = lane & (PATCH_ROWS - 1), c4 = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
        for (int cc = 0; cc < WCOLS / 32; ++cc) {
          const int col = W.n0 + half * WCOLS + cc * 32 + c4;
          const bool col_ok = col < G.N;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (G.bias != nullptr && col_ok && W.ks == 0) bv = __ldg(reinterpret_cast<const float4*>(G.bias + col));
#pragma unroll
          for (int rh = 0; rh < 32 / PATCH_ROWS; ++rh) {
            if ((lane / PATCH_ROWS) == rh) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(patch + r16 * 36 + 4 * j) =
                    make_float4(acc[cc * 32 + 4 * j] * sc, acc[cc * 32 + 4 * j + 1] * sc, acc[cc * 32 + 4 * j + 2] * sc,
                                acc[cc * 32 + 4 * j + 3] * sc);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < PATCH_ROWS / 4; ++i) {
              const int rl = 4 * i + rsub;
              const int grow = W.m0 + q * 32 + rh * PATCH_ROWS + rl;
              float4 o = *reinterpret_cast<const float4*>(patch + rl * 36 + c4);
              if (grow < G.M && col_ok) {
                long long roff = (long long)grow * G.ldc;
                if (G.c_rpb != 0) {                      // two-level row map (degree slabs of [N, K, C])
                  const int qd = grow / G.c_rpb;
                  roff = (long long)qd * G.c_bs + (long long)(grow - qd * G.c_rpb) * G.ldc;
                }
                float* cp = G.C + roff + col;
                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                if (atomic) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                               : "memory");
                  continue;
                }
                if (G.accumulate) {
                  const float4 p = *reinterpret_cast<const float4*>(cp);
                  o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                }
                *reinterpret_cast<float4*>(cp) = o;
                if constexpr (CMAX) cmax = fmaxf(cmax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
              }
            }
            __syncwarp();
          }
        }
      } else if (row < G.M) {
        const int col0 = half * WCOLS;
        long long roff = (long long)row * G.ldc;
        if (G.c_rpb != 0) {
          const int qd = row / G.c_rpb;
          roff = (long long)qd * G.c_bs + (long long)(row - qd * G.c_rpb) * G.ldc;
        }
        float* crow = G.C + roff + W.n0 + col0;
        const int ncol = min(WCOLS, G.N - W.n0 - col0);
#pragma unroll
        for (int x = 0; x < WCOLS; ++x) {
          if (x < ncol) {
            float o = acc[x] * sc;
            if (G.bias != nullptr && W.ks == 0) o += __ldg(G.bias + W.n0 + col0 + x);
            if (atomic) atomicAdd(crow + x, o);
            else if (G.accumulate) { o += crow[x]; crow[x] = o; }
            else crow[x] = o;
            if constexpr (CMAX) cmax = fmaxf(cmax, fabsf(o));
          }
        }
      }
      if (CMAX && G.c_absmax != nullptr && !atomic) {       // one fire-and-forget atomic per warp and tile
        cmax = eqv2_warp_max(cmax);
        if (lane == 0 && cmax > 0.f)
          atomicMax(reinterpret_cast<unsigned*>(G.c_absmax) + ((blockIdx.x * NUM_WORKER_WARPS + warp) & (EQV2_ABSMAX_SLOTS - 1)),
                    __float_as_uint(cmax));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- operand split: fp32 tensor -> scaled fp16 hi / lo planes ---------------------------------------------
struct SplitItem {
  const float* src;
  __half* dst;
  float* absmax;
  long long rows, cols, rows_pad, cols_pad;
  int absmax_given;    // the producer already wrote max |src| into *absmax
  int slab_k;          // > 0: src is [rows / slab_k, slab_k, cols] (l-major coefficients); dst holds the degree slabs
                       // one after the other: slab l = rows [n_nodes l^2, n_nodes (l+1)^2), row (node, j) at node (2l+1) + j
  int block_start, nblocks;
};
struct SplitParams {
  SplitItem it[EQV2_SPLIT_MAX_ITEMS];
  int n;
};

__device__ __forceinline__ const SplitItem& find_item(const SplitParams& P, int blk) {
  int i = 0;
#pragma unroll 1
  for (int k = 1; k < P.n; ++k)
    if (blk >= P.it[k].block_start) i = k;
  return P.it[i];
}

__global__ void __launch_bounds__(256) absmax_kernel(const __grid_constant__ SplitParams P) {
  const SplitItem& T = find_item(P, blockIdx.x);
  if (T.absmax_given) return;
  const int lb = blockIdx.x - T.block_start;
  const long long n = T.rows * T.cols;
  float m = 0.f;
  if ((reinterpret_cast<uintptr_t>(T.src) & 15) == 0) {
    const long long n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(T.src);
    const long long stride = (long long)T.nblocks * 256;
    for (long long i0 = (long long)lb * 256 + threadIdx.x; i0 < n4; i0 += 4 * stride) {      // 4 loads in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i0 + u * stride < n4) ? __ldg(s4 + i0 + u * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
    }
    if (lb == 0 && threadIdx.x < (int)(n & 3)) m = fmaxf(m, fabsf(__ldg(T.src + (n4 << 2) + threadIdx.x)));
  } else {
    for (long long i = (long long)lb * 256 + threadIdx.x; i < n; i += (long long)T.nblocks * 256)
      m = fmaxf(m, fabsf(__ldg(T.src + i)));
  }
  m = eqv2_warp_max(m);
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = wm[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    // non-negative floats order like their bit patterns; +inf is kept (the split then uses scale 1)
    if (threadIdx.x == 0 && m > 0.f)
      atomicMax(reinterpret_cast<unsigned int*>(T.absmax) + (blockIdx.x & (EQV2_ABSMAX_SLOTS - 1)), __float_as_uint(m));
  }
}

__global__ void __launch_bounds__(256) split_kernel(const __grid_constant__ SplitParams P) {
  const SplitItem& T = find_item(P, blockIdx.x);
  const int lb = blockIdx.x - T.block_start;
  float s, inv;
  scale_of(eqv2_read_absmax(T.absmax), s, inv);
  const unsigned gpr = (unsigned)(T.cols_pad >> 3);                      // 8-element groups per padded row
  const long long ngroups = T.rows_pad * (long long)gpr;
  const long long plane = T.rows_pad * T.cols_pad;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(T.src) & 15) == 0) && ((T.cols & 3) == 0);
  if (vec_ok && T.cols == T.cols_pad && T.rows == T.rows_pad && T.slab_k == 0) {
    // Fast path (every large activation / gradient tensor): source and planes are the same linear sequence of
    // 8-element groups; four groups per thread and iteration (8 x 128-bit loads in flight per thread).
    const long long stride = (long long)T.nblocks * 256;
    const float4* s4 = reinterpret_cast<const float4*>(T.src);
    uint4* dh = reinterpret_cast<uint4*>(T.dst);
    uint4* dl = reinterpret_cast<uint4*>(T.dst + plane);
    for (long long g0 = (long long)lb * 256 + threadIdx.x; g0 < ngroups; g0 += 4 * stride) {
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long g = g0 + u * stride;
        if (g < ngroups) {
          a[u] = __ldg(s4 + 2 * g);
          b[u] = __ldg(s4 + 2 * g + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long g = g0 + u * stride;
        if (g < ngroups) {
          const float v[8] = {a[u].x, a[u].y, a[u].z, a[u].w, b[u].x, b[u].y, b[u].z, b[u].w};
          __align__(16) __half h[8];
          __align__(16) __half l[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float x = v[t] * s;
            h[t] = __float2half_rn(x);
            l[t] = __float2half_rn(x - __half2float(h[t]));
          }
          dh[g] = *reinterpret_cast<const uint4*>(h);
          dl[g] = *reinterpret_cast<const uint4*>(l);
        }
      }
    }
    return;
  }
  for (long long g = (long long)lb * 256 + threadIdx.x; g < ngroups; g += (long long)T.nblocks * 256) {
    const long long r = (ngroups <= 0xFFFFFFFFll) ? (long long)((unsigned)g / gpr) : g / gpr;   // 32-bit division
    const long long c8 = (g - r * gpr) << 3;
    long long sr = r;                                    // source row
    if (T.slab_k > 0 && r < T.rows) {
      const int nn = (int)(T.rows / T.slab_k);
      int l = (int)sqrtf((float)((int)r / nn));
      while ((long long)(l + 1) * (l + 1) * nn <= r) ++l;
      while ((long long)l * l * nn > r) --l;
      const int rr = (int)(r - (long long)l * l * nn), w = 2 * l + 1;
      const int node = rr / w;
      sr = (long long)node * T.slab_k + l * l + (rr - node * w);
    }
    float v[8];
    if (r < T.rows && c8 + 8 <= T.cols && vec_ok) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(T.src + sr * T.cols + c8));
      const float4 b = __ldg(reinterpret_cast<const float4*>(T.src + sr * T.cols + c8 + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = (r < T.rows && c8 + t < T.cols) ? __ldg(T.src + sr * T.cols + c8 + t) : 0.f;
    }
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float u = v[t] * s;
      h[t] = __float2half_rn(u);
      l[t] = __float2half_rn(u - __half2float(h[t]));
    }
    __half* d = T.dst + r * T.cols_pad + c8;
    *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(d + plane) = *reinterpret_cast<const uint4*>(l);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// planes [2][outer][inner] of fp16 with leading dimension ld and plane stride `plane` (elements)
bool make_map(CUtensorMap* m, const void* base, long long inner, long long outer, long long ld, long long plane,
              int box_outer) {
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, 2};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)box_outer, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int sm_count() {
  static int cache[64] = {};              // per device ordinal
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

}  // namespace

extern "C" int eqv2_split_f16(const eqv2_split_desc* descs, int n, void* stream) {
  EQV2_REQUIRE(n >= 1 && n <= EQV2_SPLIT_MAX_ITEMS, "eqv2_split_f16: n=%d out of range", n);
  SplitParams P;
  memset(&P, 0, sizeof(P));
  int blocks = 0;
  bool need_pass = false;
  for (int i = 0; i < n; ++i) {
    const eqv2_split_desc& d = descs[i];
    EQV2_REQUIRE(d.src && d.dst && d.absmax, "eqv2_split_f16: null pointer in item %d", i);
    EQV2_REQUIRE(d.rows >= 0 && d.cols >= 0 && d.rows_pad >= d.rows && d.cols_pad >= d.cols && (d.cols_pad % 64) == 0 &&
                     (((uintptr_t)d.dst) & 15) == 0,
                 "eqv2_split_f16: item %d: padded extents must cover the tensor, cols_pad %% 64 == 0, dst 16-byte aligned", i);
    SplitItem& t = P.it[i];
    t.src = d.src; t.dst = reinterpret_cast<__half*>(d.dst); t.absmax = d.absmax;
    t.rows = d.rows; t.cols = d.cols; t.rows_pad = d.rows_pad; t.cols_pad = d.cols_pad;
    EQV2_REQUIRE(d.slab_k >= 0 && (d.slab_k == 0 || (d.rows % d.slab_k == 0 && d.rows < (1ll << 31))),
                 "eqv2_split_f16: item %d: rows must be a multiple of slab_k", i);
    t.slab_k = d.slab_k;
    t.absmax_given = (d.absmax_given == 1) ? 1 : 0;
    const long long groups = d.rows_pad * (d.cols_pad / 8);
    long long nb = (groups + 2047) / 2048;                 // >= 8 groups (128 B of output per plane) per thread
    if (nb < 1) nb = 1;
    if (nb > 4 * sm_count()) nb = 4 * sm_count();
    t.block_start = blocks;
    t.nblocks = (int)nb;
    blocks += (int)nb;
    if (d.absmax_given != 1) need_pass = true;
    if (d.absmax_given == 0) {                      // 2: the caller hands in a slot that is already zero
      cudaError_t e = cudaMemsetAsync(d.absmax, 0, EQV2_ABSMAX_SLOTS * sizeof(float), (cudaStream_t)stream);
      EQV2_REQUIRE(e == cudaSuccess, "eqv2_split_f16: memset failed: %s", cudaGetErrorString(e));
    }
  }
  P.n = n;
  if (need_pass) {
    absmax_kernel<<<dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream>>>(P);
    EQV2_CHECK_LAUNCH("eqv2_split_f16 (absmax)");
  }
  split_kernel<<<dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream>>>(P);
  EQV2_CHECK_LAUNCH("eqv2_split_f16 (split)");
  return 0;
}

extern "C" int eqv2_gemm_f16(const eqv2_gemm16_desc* descs, int ngroups, int split_k, void* stream) {
  return eqv2_gemm_f16_ex(descs, ngroups, split_k, 3, stream);
}

extern "C" int eqv2_gemm_f16_ex(const eqv2_gemm16_desc* descs, int ngroups, int split_k, int passes, void* stream) {
  EQV2_REQUIRE(ngroups >= 1 && ngroups <= EQV2_GEMM_MAX_GROUPS, "eqv2_gemm_f16: ngroups=%d out of range", ngroups);
  EQV2_REQUIRE(split_k >= 1, "eqv2_gemm_f16: bad split_k");
  EQV2_REQUIRE(passes == 1 || passes == 3, "eqv2_gemm_f16: passes must be 1 or 3");
  HParams P;
  memset(&P, 0, sizeof(P));
  int tiles = 0;
  for (int i = 0; i < ngroups; ++i) {
    const eqv2_gemm16_desc& d = descs[i];
    EQV2_REQUIRE(d.A && d.B && d.C && d.a_absmax && d.b_absmax, "eqv2_gemm_f16: null operand in group %d", i);
    EQV2_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "eqv2_gemm_f16: empty problem in group %d", i);
    EQV2_REQUIRE((d.a_ld % 8) == 0 && (d.b_ld % 8) == 0 && (d.a_plane % 8) == 0 && (d.b_plane % 8) == 0 &&
                     (((uintptr_t)d.A | (uintptr_t)d.B) & 15) == 0,
                 "eqv2_gemm_f16: split operands must be 16-byte aligned with leading dimensions multiple of 8");
    HGroup& g = P.g[i];
    g.C = d.C; g.bias = d.bias; g.a_absmax = d.a_absmax; g.b_absmax = d.b_absmax;
    g.c_absmax = d.c_absmax;
    g.ldc = d.c_ld;
    g.c_rpb = (d.c_rpb <= 0 || d.c_rpb >= (1ll << 31)) ? 0 : (int)d.c_rpb;
    g.c_bs = d.c_bs;
    g.M = d.M; g.N = d.N; g.K = d.K;
    g.a_mn = d.transA ? 1 : 0;
    g.b_mn = d.transB ? 0 : 1;
    g.accumulate = d.accumulate;
    const bool oka = g.a_mn ? make_map(&g.mapA, d.A, d.M, d.K, d.a_ld, d.a_plane, 64)
                            : make_map(&g.mapA, d.A, d.K, d.M, d.a_ld, d.a_plane, 128);
    const bool okb = g.b_mn ? make_map(&g.mapB, d.B, d.N, d.K, d.b_ld, d.b_plane, 64)
                            : make_map(&g.mapB, d.B, d.K, d.N, d.b_ld, d.b_plane, 128);
    EQV2_REQUIRE(oka && okb, "eqv2_gemm_f16: cuTensorMapEncodeTiled failed for group %d", i);
    g.tiles_m = (d.M + BM - 1) / BM;
    g.tiles_n = (d.N + BN - 1) / BN;
    g.tile_start = tiles;
    tiles += g.tiles_m * g.tiles_n;
  }
  P.ngroups = ngroups;
  P.split_k = split_k;
  P.total_work = tiles * split_k;
  P.passes = passes;
  bool want_cmax = false;
  for (int i = 0; i < ngroups; ++i) want_cmax = want_cmax || (descs[i].c_absmax != nullptr && split_k == 1);
  typedef void (*KernelFn)(const HParams);
  const KernelFn variants[4] = {gemm_f16_kernel<3, false>, gemm_f16_kernel<3, true>, gemm_f16_kernel<1, false>,
                                gemm_f16_kernel<1, true>};
  const int vi = (passes == 1 ? 2 : 0) + (want_cmax ? 1 : 0);
  static bool attr_set[64][4] = {};       // the attribute is per device and per kernel instance
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev][vi]) {
    cudaError_t e = cudaFuncSetAttribute(variants[vi], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    EQV2_REQUIRE(e == cudaSuccess, "eqv2_gemm_f16: cannot reserve %zu B of shared memory: %s", SMEM_BYTES,
                 cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev][vi] = true;
  }
  const int grid = P.total_work < sm_count() ? P.total_work : sm_count();
  variants[vi]<<<dim3((unsigned)grid), dim3(NUM_THREADS), SMEM_BYTES, (cudaStream_t)stream>>>(P);
  EQV2_CHECK_LAUNCH("eqv2_gemm_f16");
  return 0;
}

#endif  // EQV2_CPU_EMU
