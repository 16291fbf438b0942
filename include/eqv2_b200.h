/*
 * eqv2_b200.h -- C ABI of libeqv2_b200.so: hand-written sm_100a kernels for the
 * EquiformerV2 SO(2)-equivariant graph-attention block.
 *
 * The reference (johannebirkchristensen/EquivariantTransformerMPNN4QuantumComputations) has
 * no FFI: its hot path is the Python package models/EquiformerV2Functions.  Each
 * entry point below replaces the eager-PyTorch op sequence of the cited reference lines and
 * is what a binding for that path would call.  Conventions:
 *   - plain pointers to DEVICE memory, fp32 values, int64 node indices (as the reference),
 *     int32 CSR tables; sizes as int / long long; `stream` is a cudaStream_t;
 *   - no allocation inside: the caller passes outputs and workspaces;
 *   - return 0 on success, non-zero otherwise with a thread-local message from
 *     eqv2_last_error(); launches are asynchronous on `stream`.
 */
#ifndef EQV2_B200_H
#define EQV2_B200_H

#ifdef __cplusplus
extern "C" {
#endif

const char* eqv2_last_error(void);
int eqv2_abi_version(void);

/* ---- grouped GEMM engine ---------------------------------------------------------------
 * C[i,j] (+)= sum_k opA(i,k) opB(k,j) (+ bias[j]);  row r of an operand lives at
 * (r / rpb) * bs + (r % rpb) * ld  (two-level stride: degree-l slabs of [N,K,C] tensors).
 * Replaces: F.linear in so2_ops.py:150-185 (SO2 convolution blocks), radial_function.py:29,
 * the einsum of so3.py:722-727 (SO3_LinearV2) and their autograd backward. */
#define EQV2_GEMM_MAX_GROUPS 10
typedef struct {
  const float* A;
  const float* B;
  float* C;
  const float* bias; /* may be NULL */
  int M, N, K;
  int transA, transB;
  long long a_rpb, a_bs, a_ld;
  long long b_rpb, b_bs, b_ld;
  long long c_rpb, c_bs, c_ld;
  int accumulate; /* C += (ignored when split_k > 1: C must then be pre-initialised) */
} eqv2_gemm_desc;

/* exact fp32 (FFMA) engine: any operand addressing */
int eqv2_gemm_f32(const eqv2_gemm_desc* descs, int ngroups, int split_k, void* stream);
/* tensor-core engine (tcgen05.mma kind::tf32, TMEM accumulators): operands 16-byte aligned, leading
 * dimensions and block strides multiples of 4 (the two-level stride applies to the non-contiguous index).
 * mode 0 = 3xTF32 split (fp32-class accuracy), mode 1 = 1xTF32. */
int eqv2_gemm_tc(const eqv2_gemm_desc* descs, int ngroups, int split_k, int mode, void* stream);

/* second tensor-core engine (tcgen05.mma kind::f16, TMA operands, persistent CTAs): fp32-class accuracy from three
 * fp16 passes over PRE-SPLIT operands.
 *   eqv2_split_f16: for each contiguous fp32 tensor [rows, cols] write the planes hi = fp16(s v), lo = fp16(s v - hi)
 *     into dst[2][rows_pad][cols_pad] (fp16, zero padding; cols_pad % 64 == 0) with s the power of two that puts
 *     max |s v| into [2^14, 2^15); absmax[64] receives max |v| (spread over 64 slots; the GEMM derives 1/s from their
 *     maximum on the device).
 *   eqv2_gemm_f16: A / B point at the hi plane of the operand's sub-block (lo plane `*_plane` elements further,
 *     leading dimension `*_ld`; both multiples of 8 elements, pointers 16-byte aligned).
 *     transA = 0: A(m,k) at A[m*a_ld + k]; 1: A[k*a_ld + m].  transB = 1: B(k,n) at B[n*b_ld + k]; 0: B[k*b_ld + n].
 *     C is fp32 [M, c_ld]; accumulate / split_k / bias as for eqv2_gemm_f32. */
#define EQV2_SPLIT_MAX_ITEMS 16
typedef struct {
  const float* src;
  void* dst;       /* fp16 [2][rows_pad][cols_pad] */
  float* absmax;   /* EQV2_ABSMAX_SLOTS (64) floats whose maximum is max |v| of src; written by the call */
  long long rows, cols, rows_pad, cols_pad;
  int absmax_given; /* 0: compute max |v| (the call zeroes the slot first); 1: absmax already holds it (written by the
                       producing kernel): skip the reduction pass; 2: compute it, the slot is already zero */
  int slab_k;      /* 0: plain.  > 0: src is a node tensor [rows / slab_k, slab_k = (lmax+1)^2, cols] (so3.py:76-88) and dst
                      receives its degree slabs one after the other -- slab l = rows [n l^2, n (l+1)^2), row (node, j) at
                      node (2l+1) + j -- so that every SO3_LinearV2 block (so3.py:722-727) is a plain matrix */
} eqv2_split_desc;
typedef struct {
  const void* A;
  const void* B;
  float* C;
  const float* bias; /* may be NULL */
  const float* a_absmax;
  const float* b_absmax;
  long long a_ld, a_plane, b_ld, b_plane, c_ld;
  long long c_rpb, c_bs; /* C row r at (r / c_rpb) * c_bs + (r % c_rpb) * c_ld; c_rpb <= 0 or >= 2^31: plain r * c_ld */
  int M, N, K;
  int transA, transB;
  int accumulate;
  float* c_absmax;   /* may be NULL.  Else (split_k == 1 only): 64 zero-initialised floats whose maximum becomes max |C| of
                        everything the launch stores -- the operand bound of whichever kernel consumes C next */
} eqv2_gemm16_desc;
int eqv2_split_f16(const eqv2_split_desc* descs, int n, void* stream);
int eqv2_gemm_f16(const eqv2_gemm16_desc* descs, int ngroups, int split_k, void* stream);
/* passes = 3: the fp32-class product above; passes = 1: hi planes only -- ONE fp16 tensor-core pass, ~3x the rate,
 * relative error ~5e-4 per product (the reduced-precision "bf16/TF32 GEMM mode" of BASELINE configs[3]; tolerance stated
 * in tests/test_model_parity.py).  The lo planes are not read. */
int eqv2_gemm_f16_ex(const eqv2_gemm16_desc* descs, int ngroups, int split_k, int passes, void* stream);

/* ---- Wigner-D rotation (so3.py:343-387,499-545; transformer_block.py:250-275,321-331) ---- */
/* Edge frames [E,3,3] (rows z, x_edge, -y) from edge vectors.  mode 0: edge_rot_mat.py:13-80 with the helper draw
 * (`torch.rand_like(vec) - 0.5`, line 28) supplied by the caller so that the RNG stays the reference's; stats[2] (zeroed
 * unsigned) receive ~bits(min |vec|) and bits(max |<helper, x>|) -- the two conditions the reference checks on the host
 * (lines 19-24, 58).  mode 1: the deterministic variant of equiformerv2_MatPESv2.py:41-66 (draw / stats may be NULL). */
int eqv2_edge_frames(const float* vec /*[E,3]*/, const float* draw /*[E,3]*/, float* out /*[E,3,3]*/, long long E, int mode,
                     unsigned* stats, void* stream);
int eqv2_wigner_from_rot(const float* rot /*[E,3,3]*/, const float* Jd /*packed blocks*/,
                         float* wig /*[E, sum (2l+1)^2]*/, long long E, int lmax, void* stream);

int eqv2_gather_rotate_fwd(const float* x /*[N,K,C]*/, const long long* src, const long long* dst,
                           const float* wig, const float* rad /*[E,nrad] or NULL*/,
                           float* out /*[E,Kr,2C] m-primary*/, const int* pos_of_full /*[K]*/,
                           const int* rad_slot /*[Kr]*/, long long E, int C, int lmax, int mmax, int Kr,
                           int nrad, float* absmax /*or NULL*/, void* stream);

/* `absmax` (gather_rotate_fwd / _drad, rotinv_reduce_bwd): optional device array of 64 floats, zero-initialised by the
 * caller, whose maximum becomes max |v| over everything the launch writes -- the operand scale the f16x3 GEMM engine needs for this tensor,
 * produced for free instead of by a separate pass (eqv2_split_desc.absmax_given). */
/* backward of the gather/rotate, split by its two outputs:
 *   dx   [N,K,C]   = sum over the node's outgoing (src half) and incoming (dst half) edges of W_e^T (dA * rad)
 *                    -- node-centric over the two CSR views, deterministic (no atomics);
 *   drad [E,nrad]  = dA * (W_e x) summed over the rows that share a radial weight -- edge-parallel. */
int eqv2_gather_rotate_dx(const float* wig, const float* rad /*or NULL*/, const float* dA /*[E,Kr,2C]*/,
                          const int* rowptr_src, const int* perm_src, const int* rowptr_dst, const int* perm_dst,
                          float* dx, long long N, int C, int lmax, int mmax, int Kr, int nrad, void* stream);
int eqv2_gather_rotate_drad(const float* x, const long long* src, const long long* dst, const float* wig,
                            const float* dA, float* drad, long long E, int C, int lmax, int mmax, int Kr, int nrad,
                            float* absmax /*or NULL*/, void* stream);

int eqv2_rotinv_reduce_fwd(const float* val /*[E,rows,Cv]*/, const float* alpha /*[E,heads] or NULL*/,
                           const float* wig, const int* rowptr_dst, const int* perm_dst,
                           float* out /*[N,K,Cv]*/, const int* pos_of_full, long long N, int Cv, int rows_used,
                           long long val_estride, int heads, int lmax, int mmax, float scale, void* stream);

int eqv2_rotinv_reduce_bwd(const float* dout /*[N,K,Cv]*/, const float* val, const float* alpha,
                           const float* wig, const long long* dst, float* dval, float* dalpha /*or NULL*/,
                           const int* pos_of_full, long long E, int Cv, int rows_used, long long val_estride,
                           int heads, int lmax, int mmax, float scale, float* absmax /*or NULL*/, void* stream);

/* Producer-side operand planes (f16x3 engine).  The three edge-parallel rotate kernels can write their result directly as
 * the scaled fp16 hi/lo planes the GEMM consumes ([2][E][ld] fp16, plane 1 `plane` elements after plane 0) instead of an
 * fp32 tensor that a separate eqv2_split_f16 pass would re-read: the scale comes from a BOUND on the result computed on
 * the device from the maxima of the kernel's own inputs (64-float slots as written by the `absmax` outputs of the kernels
 * above / `c_absmax` of eqv2_gemm_f16), and the bound is stored to bound_out[0] -- the slot the consuming GEMM is handed
 * as a_absmax / b_absmax (the other 63 floats must be zero).  A bound that overshoots the true maximum by up to ~2^8 keeps
 * the product fp32-class (csrc/common.cuh).  Same argument meaning as the fp32 entry points otherwise.
 *   gather_rotate_fwd_planes : bound = max|x| max|rad| sqrt(2 lmax + 1)                (transformer_block.py:250-275)
 *   gather_rotate_drad_planes: bound = 2 max|x| max|dA| sqrt(2 lmax + 1)
 *   rotinv_reduce_bwd_planes : bound = max|dout| alpha_bound |scale| sqrt((2 lmax+1) max(1, (2 lmax+1)/(2 mmax+1)));
 *                              alpha_bound >= max |alpha| is the caller's (softmax output: 1, with dropout p: 1/(1-p)) */
/* out[c] = sum_r T[r, col_off + c] of a matrix T held only as operand planes (bias gradients); `bound` = the planes'
 * bound slot; partial: S * C floats of workspace; deterministic (fixed summation order). */
int eqv2_planes_colsum(const void* planes, long long plane, long long ld, long long col_off, long long rows, int C, int S,
                       const float* bound, float* partial, float* out, void* stream);
int eqv2_gather_rotate_fwd_planes(const float* x, const long long* src, const long long* dst, const float* wig,
                                  const float* rad, void* planes, long long plane, long long ld, const float* bound_x,
                                  const float* bound_rad, float* bound_out, long long E, int C, int lmax, int mmax,
                                  int Kr, int nrad, void* stream);
int eqv2_gather_rotate_drad_planes(const float* x, const long long* src, const long long* dst, const float* wig,
                                   const float* dA, void* planes, long long plane, long long ld, const float* bound_x,
                                   const float* bound_dA, float* bound_out, long long E, int C, int lmax, int mmax,
                                   int Kr, int nrad, void* stream);
int eqv2_rotinv_reduce_bwd_planes(const float* dout, const float* val, const float* alpha, const float* wig,
                                  const long long* dst, void* planes, long long plane, long long ld,
                                  const float* bound_dout, float alpha_bound, float* bound_out, float* dalpha,
                                  long long E, int Cv, int rows_used, long long val_estride, int heads, int lmax,
                                  int mmax, float scale, void* stream);

/* ---- separable S2 activation (activation.py:153-192, so3.py:552-646) --------------------- */
int eqv2_s2act_padded_rows(int Kr);
int eqv2_s2act_fwd(const float* X, long long x_rs, const float* gate /*or NULL*/, long long g_rs, float* O,
                   long long o_rs, const float* to_grid /*[G,KP]*/, const float* from_grid /*[G,KP]*/,
                   long long R, int C, int Kr, int KP, int G, int nblocks, void* stream);
int eqv2_s2act_bwd(const float* X, long long x_rs, const float* gate, long long g_rs, const float* dO,
                   long long o_rs, float* dX, long long dx_rs, float* dgate, long long dg_rs,
                   const float* to_grid, const float* from_grid, long long R, int C, int Kr, int KP, int G,
                   int nblocks, void* stream);

/* latitude/longitude-factorised version of the same operator (csrc/s2act_sep.cu): resolution-18 grids,
 * factor tables (float block laid out as [7][18][7] Pt | [7][18][7] Pf | [18][7] cos | [18][7] sin) in one
 * of six __constant__ slots (the host side owns the slot -> table-set map; slots used inside a captured CUDA graph are never rebound); m_primary selects the coefficient order of X / O. */
int eqv2_s2sep_supported(int lmax, int mmax);
int eqv2_s2sep_set_tables(const float* host_tables, int nfloats, int slot, void* stream);
int eqv2_s2sep_fwd(const float* X, long long x_rs, const float* gate, long long g_rs, float* O, long long o_rs,
                   long long R, int C, int lmax, int mmax, int m_primary, int slot,
                   float* absmax /*may be NULL; else 64 zeroed floats receiving max |O|*/, void* stream);
/* eqv2_s2sep_fwd with the result written as scaled fp16 hi/lo operand planes [2][R][ld] (the A operand of the second SO(2)
 * convolution, so2_ops.py:150-185) instead of fp32; bound_out[0] <- max(bound_in slots) * bound_c, bound_c >= the operator
 * norm ||from_grid||_1 ||to_grid||_inf of the activation (and >= 1 for the gate row).  C must divide 128, C % 8 == 0. */
int eqv2_s2sep_fwd_planes(const float* Xp, long long x_rs, const float* gate, long long g_rs, void* planes, long long plane,
                          long long ld, const float* bound_in, float bound_c, float* bound_out, long long R, int C, int lmax,
                          int mmax, int m_primary, int slot, void* stream);
int eqv2_s2sep_bwd(const float* X, long long x_rs, const float* gate, long long g_rs, const float* dO, long long o_rs,
                   float* dX, long long dx_rs, float* dgate, long long dg_rs, long long R, int C, int lmax, int mmax,
                   int m_primary, int slot, float* absmax /*may be NULL; max |dX|, |dgate|*/, void* stream);

/* derivative of eqv2_s2sep_bwd w.r.t. (X, gate, dO) for cotangents U (of dX) and Wg (of dgate): the second-order
 * term needed when forces = -dE/dpos are trained on (train_MatPES_GATAWandB.py:72-91). */
int eqv2_s2sep_bwd2(const float* X, long long x_rs, const float* gate, long long g_rs, const float* dO, long long o_rs,
                    const float* U, long long u_rs, const float* Wg, long long w_rs, float* d2X, long long d2x_rs,
                    float* d2gate, long long d2g_rs, float* d2O, long long d2o_rs, long long R, int C, int lmax, int mmax,
                    int m_primary, int slot, void* stream);

/* ---- attention logits + segment softmax (transformer_block.py:311-315) ------------------- */
int eqv2_attn_alpha_fwd(const float* Y, long long y_rs, const float* ln_w /*or NULL*/, const float* ln_b,
                        const float* alpha_dot /*[heads,ach]*/, const int* rowptr_dst, const int* perm_dst,
                        float* logits /*[E,heads]*/, float* alpha /*[E,heads]*/, long long E, long long N,
                        int heads, int ach, float eps, void* stream);
int eqv2_attn_alpha_bwd(const float* Y, long long y_rs, const float* ln_w, const float* ln_b,
                        const float* alpha_dot, const int* rowptr_dst, const int* perm_dst, const float* alpha,
                        const float* dalpha, float* dlogits, float* dY, long long dy_rs,
                        float* d_ln_w /*zeroed*/, float* d_ln_b /*zeroed*/, float* d_alpha_dot /*zeroed*/,
                        long long E, long long N, int heads, int ach, float eps,
                        float* absmax /*may be NULL: max |dY| into the same kind of slot*/, void* stream);
/* derivative of eqv2_attn_alpha_bwd's dY w.r.t. (Y, ln_w, ln_b, alpha_dot, alpha, dalpha) for a cotangent U of dY (double
 * backward of the force loss).  dlogits = the first backward's softmax gradient; R [E, heads] workspace; parameter
 * gradients accumulate atomically into zeroed buffers; d_alpha is the cotangent of `alpha` AS AN INPUT of the backward
 * (the caller's autograd carries it on through the forward operator's backward). */
int eqv2_attn_alpha_bwd2(const float* Y, long long y_rs, const float* ln_w, const float* ln_b, const float* alpha_dot,
                         const int* rowptr_dst, const int* perm_dst, const float* alpha, const float* dalpha,
                         const float* dlogits, const float* U, long long u_rs, float* R, float* d2Y, long long d_rs,
                         float* d_ln_w /*zeroed*/, float* d_ln_b /*zeroed*/, float* d_alpha_dot /*zeroed*/,
                         float* d_dalpha, float* d_alpha, long long E, long long N, int heads, int ach, float eps,
                         void* stream);

/* ---- equivariant norms (layer_norm.py:38-108,112-201,265-351) ---------------------------- */
int eqv2_equiv_norm_fwd(const float* x /*[N,K,C]*/, const float* w /*[lmax+1,C]*/, const float* b /*[C]*/,
                        float* out, float* inv_out /*[N,ngroups]*/, float* mean_out /*[N]*/, long long N, int C,
                        int lmax, int ngroups, const int* group_of_l /*host*/, const float* bw_l /*host*/,
                        float eps, float* absmax /*may be NULL; else 64 zeroed floats receiving max |out|*/, void* stream);
int eqv2_equiv_norm_bwd(const float* x, const float* w, const float* go, const float* inv_in,
                        const float* mean_in, float* dx, float* dw /*zeroed*/, float* db /*zeroed*/,
                        long long N, int C, int lmax, int ngroups, const int* group_of_l /*host*/,
                        const float* bw_l /*host*/, void* stream);
/* derivative of eqv2_equiv_norm_bwd's dx w.r.t. (x, go, w) for a cotangent u of dx (double backward of the force loss);
 * dw accumulated atomically into a zeroed buffer; the bias does not enter dx */
int eqv2_equiv_norm_bwd2(const float* x, const float* w, const float* go, const float* inv_in, const float* mean_in,
                         const float* u, float* d2x, float* dgo, float* dw /*zeroed*/, long long N, int C, int lmax,
                         int ngroups, const int* group_of_l, const float* bw_l, void* stream);

/* ---- edge scalar features (equiformerv2_oc20.py:43-60, radial_function.py:21-22) --------- */
int eqv2_rbf_fwd(const float* d, float* out /*[E,R]*/, long long E, int R, const float* offset /*[R]*/, float coeff,
                 void* stream);
int eqv2_rbf_bwd(const float* d, const float* go, float* dd, long long E, int R, const float* offset, float coeff,
                 void* stream);
/* derivative of eqv2_rbf_bwd's dd w.r.t. (go, d) for a cotangent u [E] of dd (double backward of the force loss) */
int eqv2_rbf_bwd2(const float* d, const float* go, const float* u, float* dgo /*[E,R]*/, float* d2d /*[E]*/, long long E,
                  int R, const float* offset, float coeff, void* stream);
/* real SH of the edge direction, l = 1..lmax, |Y_l| = 1 (e3nn SphericalHarmonics(normalize=False,
 * normalization='norm') at equiformerv2_MatPES_GATAV2.py:137-140,232-241); out [E, (lmax+1)^2 - 1] */
int eqv2_edge_sh(const float* vec /*[E,3]*/, float* out, long long E, int lmax, void* stream);
int eqv2_ln_silu_fwd(const float* x, const float* w, const float* b, float* y, long long rows, int width,
                     float eps, void* stream);
int eqv2_ln_silu_bwd(const float* x, const float* w, const float* b, const float* gy, float* gx,
                     float* gw /*zeroed*/, float* gb /*zeroed*/, long long rows, int width, float eps,
                     void* stream);
/* derivative of eqv2_ln_silu_bwd's gx w.r.t. (x, gy, w, b) for a cotangent u of gx (forces by autograd: double backward
 * of the radial MLP); dw / db are accumulated atomically into zeroed buffers */
int eqv2_ln_silu_bwd2(const float* x, const float* w, const float* b, const float* gy, const float* u, float* dx,
                      float* dgy, float* dw /*zeroed*/, float* db /*zeroed*/, long long rows, int width, float eps,
                      void* stream);

/* ---- neighbour lists and per-graph reductions -------------------------------------------
 * Builders run count (mode 0: deg[N]) -> eqv2_exclusive_scan -> fill (mode 1) and emit edges sorted
 * by destination; `err` is a device int set to 1 if a destination had more in-cutoff candidates
 * than the kernel's shared-memory capacity.
 * Replaces: equiformerv2_qm9.py:423-525 (generate_graph), the fairchem generate_graph call at
 * equiformerv2_oc20.py:223-234, and the index_add_ readouts equiformerv2_oc20.py:278-281. */
int eqv2_graph_ptr(const long long* natoms /*[B]*/, int* graph_ptr /*[B+1]*/, int B, void* stream);
int eqv2_exclusive_scan(const int* in /*[n]*/, int* out /*[n+1]*/, int n, void* stream);
int eqv2_radius_graph(const float* pos /*[N,3]*/, const int* graph_ptr, const long long* batch, long long N,
                      float cutoff, int max_nb /* <0: unlimited */, int mode, int* deg, const int* rowptr,
                      long long* src, long long* dst, float* dist, float* vec /*[E,3]*/, int* err, void* stream);
int eqv2_pbc_reps(const float* cell /*[B,3,3]*/, double cutoff, int* reps /*[B,3]*/, int B, void* stream);
int eqv2_radius_graph_pbc(const float* pos, const float* cell, const int* graph_ptr, const long long* batch,
                          const int* reps, long long N, double cutoff, int max_nb, int strict, int mode, int* deg,
                          const int* rowptr, long long* nbr, long long* ctr, float* dist, float* vec, int* err,
                          void* stream);
/* MatPES builders: 27 images, fp32; version 1 = equiformerv2_MatPES.py:258-340 (true image vectors),
 * version 2 = equiformerv2_MatPESv2.py:177-240 / equiformerv2_MatPES_GATAV2.py:285-349 (ranking and vectors
 * without the image offset).  row src = i, row dst = j of diff[i,j] = pos[j] + offset - pos[i]. */
int eqv2_radius_graph_pbc27(const float* pos, const float* cell, const int* graph_ptr, const long long* batch,
                            long long N, float cutoff, int max_nb, int version, int mode, int* deg, const int* rowptr,
                            long long* src, long long* dst, int* img, float* dist, float* vec, int* err, void* stream);
int eqv2_csr_from_index(const long long* idx /*[E]*/, long long E, long long N, int* counts /*zeroed [N]*/,
                        int* rowptr /*[N+1]*/, int* cursor /*zeroed [N]*/, int* perm /*[E]*/, void* stream);
int eqv2_segment_sum_fwd(const float* v, long long v_stride, const long long* batch /*non-decreasing [N]*/,
                         float* out /*[B]*/, long long N, int B, void* stream);
int eqv2_segment_sum_bwd(const float* gout /*[B]*/, const long long* batch, float* gv /*[N]*/, long long N,
                         void* stream);

/* ---- row gathers and deterministic segmented column sums (csrc/rows.cu) --------------------------
 * eqv2_embed_rows: out[e,:] = table[idx[e],:] -- nn.Embedding lookups source_embedding(Z[src]) / target_embedding(Z[dst])
 *   of transformer_block.py:241-248 and input_block.py:93-100.
 * eqv2_seg_colsum: out[v,c] = sum over rows i in [rowptr[v], rowptr[v+1]) of src[perm[i]*ld + c] (rowptr NULL: one
 *   segment [0, rows); perm NULL: identity) in a fixed order; `partial` is a [V, S, C] workspace.  Embedding weight
 *   gradients (segments = element types) and bias gradients of the dense layers. */
/* eqv2_so2_block_weight: B[2h,2k] = [[Wr,-Wi],[Wi,Wr]] from fc.weight W[2h,k] = [Wr;Wi] of an order-m > 0 SO(2)
 *   convolution (so2_ops.py:53-61), so that the +- recombination is part of the GEMM; _adj: the adjoint map
 *   gW = [gB00 + gB11; gB10 - gB01] (weight gradient). */
int eqv2_so2_block_weight(const float* W, float* B, int h, int k, void* stream);
int eqv2_so2_block_weight_adj(const float* gB, float* gW, int h, int k, void* stream);
int eqv2_embed_rows(const float* table, const long long* idx, float* out, long long E, int C, void* stream);
int eqv2_seg_colsum(const float* src, long long ld, const int* rowptr, const int* perm, long long rows, int V, int C,
                    int S, float* partial, float* out, void* stream);

/* ---- GaussianSmearing + first radial-MLP layer, fused (equiformerv2_oc20.py:43-60, radial_function.py:5-30,
 * transformer_block.py:241-248; north_star piece 2) -------------------------------------------------
 *   out[e, :] = Wt^T rbf(d_e) + Ts[zs[e]] + Td[zd[e]] + bias,   rbf_k(d) = exp(coeff (d - offset[k])^2),
 * evaluated over the 2 band + 1 basis functions nearest to d_e (offset = linspace(start, start + (R-1) delta, R); the
 * neglected terms are < exp(-band^2 / (2 w^2))).  Wt = W1[:, :R]^T [R, H]; Ts / Td [V, H] = embedding tables already
 * multiplied by their weight slices (NULL: no embedding terms).  First order: no gradient w.r.t. d.
 *   eqv2_rbf_linear_wgrad: gWt[k, :] = sum_e rbf_k(d_e) gh[e, :], deterministic, two stages over the edges sorted by nearest
 *   basis index (perm): chunks of eqv2_rbf_linear_chunk() sorted edges write one partial row per basis function they can
 *   touch -- chunk_k[2c], chunk_k[2c+1] = that contiguous range, chunk_base[c] = its first row in `partial`
 *   (sum of the range lengths rows x H floats) -- and every basis function k adds the rows of the chunks
 *   k_chunks[2k] .. k_chunks[2k+1] that touch it, in chunk order (an empty range: first > last). */
int eqv2_rbf_linear_fwd(const float* d, const float* offset, const float* Wt, const float* Ts, const float* Td,
                        const long long* zs, const long long* zd, const float* bias, float* out, long long E, int R, int H,
                        float start, float delta, float coeff, int band, void* stream);
int eqv2_rbf_linear_chunk(void);
int eqv2_rbf_linear_wgrad(const float* d, const float* offset, const int* perm, const int* chunk_k, const int* chunk_base,
                          const int* k_chunks, const float* gh, float* partial, float* gWt, long long E, int R, int H,
                          float coeff, void* stream);

/* ---- GATA / HTR per-edge operators (BASELINE configs 4-5; NewFunctions/Gotennet_morethaninspired/activation.py) ----
 * HTR (:166-264): T(q, k; r)[e,c] = sum_{l>=1} [ q^l.k^l - (2 - |r^l|^2)(q^l.r^l)(k^l.r^l) ] / (2l+1)  -- the inner product of the
 *   two vector rejections -- and its gradient map G(g, b; r)[e,m,c] = g[e,c]/(2l+1) (b - (2 - |r^l|^2)(b^l.r^l) r)[e,m,c];
 *   q, k, b, G: [E, M, H] with M = (lmax+1)^2 - 1 rows (l = 1..lmax), r: [E, M] (detached), T, g: [E, H].
 *   {T, G} is closed under differentiation (dT/dq = G(g,k), dT/dk = G(g,q), dG/dg = T(u,b), dG/db = G(g,u)).
 * GATAValueActivation (:270-414): comb [E, (1 + 2 lmax) H] = (o_s | o_d^l | o_t^l), Xp [E, M, H] ->
 *   out [E, Kr, H]: row 0 = SiLU(o_s); degree l rows m < min(2l+1, 2 mmax+1): o_d^l r[off_l+m] + o_t^l Xp[off_l+m]
 *   (Kr = 1 + sum_l min(2l+1, 2 mmax+1), l-primary).  bwd: g -> (d_comb, d_Xp); bwd2: cotangents u (of d_comb, may be
 *   NULL), v (of d_Xp, may be NULL) -> (d_g, d2_comb, d2_Xp). */
int eqv2_htr_inner(const float* q, const float* k, const float* rl, float* out, long long E, int H, int lmax, void* stream);
int eqv2_htr_grad(const float* g, const float* b, const float* rl, float* out, long long E, int H, int lmax, void* stream);
int eqv2_gata_value_fwd(const float* comb, const float* Xp, const float* rl, float* out, long long E, int H, int lmax,
                        int mmax, int Kr, void* stream);
int eqv2_gata_value_bwd(const float* comb, const float* Xp, const float* rl, const float* g, float* dcomb, float* dXp,
                        long long E, int H, int lmax, int mmax, int Kr, void* stream);
int eqv2_gata_value_bwd2(const float* comb, const float* Xp, const float* rl, const float* g, const float* u,
                         const float* v, float* dg, float* d2comb, float* d2Xp, long long E, int H, int lmax, int mmax,
                         int Kr, void* stream);

/* ---- per-graph stochastic depth (EquiformerV2Functions/drop.py:16-27,49-68 `GraphDropPath`) --------------------------
 * out[n, 0:row] = x[n, 0:row] * (1 / keep) * floor(keep + u[batch[n]]);  u [num_graphs] is the reference's uniform draw. */
int eqv2_drop_path_scale(const float* x, const float* u, const long long* batch, float keep, float* out, long long N,
                         long long row, void* stream);

/* ---- all-pairs attention core of the global-attention classes (NewFunctions/GATA_and_all2all/activation.py:419-1567:
 * `attn = softmax(q k^T * scale + bias)` over the atoms of the query's own structure, `out_l = attn @ v_l`; the reference
 * forms an [N_tot, N_tot] map and masks cross-structure pairs) -------------------------------------------------------
 * Ragged pair tensors [P, H], P = sum_g n_g^2: row i (atom of structure g: first atom gstart[i], gcount[i] atoms) owns the
 * contiguous entries rowptr[i] + (j - gstart[i]).  Node tensors a, b, out are [N, M, H, D] fp32, D = 4 x a power of two.
 *   eqv2_pair_scores      : S[pair(i,j), h] = scale * sum_{m,d} a[i,m,h,d] b[j,m,h,d]
 *   eqv2_pair_mix         : out[i,m,h,d] = sum_j W[pair(i,j), h] b[j,m,h,d]   (transpose != 0: out[j] = sum_i W[pair(i,j)] b[i]);
 *                           max_count = largest structure (shared-memory sizing)
 *   eqv2_pair_softmax_fwd : Pw = softmax over each row and head
 *   eqv2_pair_softmax_bwd : gS = Pw (gP - sum_j Pw gP)
 *   eqv2_pair_softmax_bwd2: for a cotangent u of gS: dgP = Pw (u - q), dP = u (gP - r) - gP q  (r = sum Pw gP, q = sum Pw u)
 * scores / mix are each other's derivatives (csrc/pair_attn.cu), so the double backward of a force loss needs no more. */
int eqv2_pair_scores(const float* a, const float* b, const int* gstart, const int* gcount, const long long* rowptr,
                     float* S, long long N, int M, int H, int D, float scale, void* stream);
int eqv2_pair_mix(const float* W, const float* b, const int* gstart, const int* gcount, const long long* rowptr, float* out,
                  long long N, int M, int H, int D, int max_count, int transpose, void* stream);
int eqv2_pair_softmax_fwd(const float* S, const int* gcount, const long long* rowptr, float* Pw, long long N, int H,
                          void* stream);
int eqv2_pair_softmax_bwd(const float* Pw, const float* gP, const int* gcount, const long long* rowptr, float* gS,
                          long long N, int H, void* stream);
int eqv2_pair_softmax_bwd2(const float* Pw, const float* gP, const float* u, const int* gcount, const long long* rowptr,
                           float* dP, float* dgP, long long N, int H, void* stream);

/* ---- optimizer-side step (train_oc20v2_parallel.py:95-126,177-186) -----------------------------------
 * Multi-tensor kernels over ONE device table of the model's parameter tensors, processed in chunks of
 * eqv2_opt_chunk_elems() elements: chunk c covers elements [chunk_index[c] * chunk, ...) of tensors[chunk_tensor[c]].
 *   eqv2_grad_sqnorm   : out[0] = global L2 norm of all gradients (deterministic summation order), out[1] = the
 *                        clip coefficient min(1, max_norm / (norm + 1e-6)) of torch.nn.utils.clip_grad_norm_
 *                        (1 when max_norm <= 0); partial: nchunks floats of workspace.
 *   eqv2_adamw_ema_step: torch.optim.AdamW update (decoupled weight decay, bias corrections 1 - beta^t with the
 *                        tensor's own t = step - lag, evaluated in double) on the clipped gradient
 *                        (clip = out of eqv2_grad_sqnorm, or NULL), followed by the reference's
 *                        ExponentialMovingAverage.update on `ema` (NULL: none).  Tensors with g == NULL are skipped. */
typedef struct {
  float* p;        /* parameter */
  const float* g;  /* gradient, may be NULL */
  float* m;        /* exp_avg */
  float* v;        /* exp_avg_sq */
  float* ema;      /* EMA shadow, may be NULL */
  long long n;
  float lr, wd;
  int lag;         /* optimizer steps this tensor has NOT taken (no gradient): its own step count = step - lag, as torch */
  int pad_;
} eqv2_opt_tensor;
int eqv2_opt_chunk_elems(void);
int eqv2_grad_sqnorm(const eqv2_opt_tensor* tensors /*device*/, const int* chunk_tensor /*device*/,
                     const int* chunk_index /*device*/, int nchunks, float max_norm, float* partial, float* out /*[2]*/,
                     void* stream);
int eqv2_adamw_ema_step(const eqv2_opt_tensor* tensors, const int* chunk_tensor, const int* chunk_index, int nchunks,
                        const float* clip /*[2] or NULL*/, float beta1, float beta2, float eps, int step /*>= 1*/,
                        float ema_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EQV2_B200_H */
