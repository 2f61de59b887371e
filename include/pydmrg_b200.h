/*
 * pydmrg_b200 -- C ABI of the B200-native (sm_100a) DMRG sweep hot path.
 *
 * The reference (philliphelms/PyDMRG) is pure Python and has no FFI of its own; its seam
 * for this path is the closure / step-function contract
 *     Hfun(x) -> -H_eff x                          pydmrg/tools/diag_tools.py:108-125, 146-162
 *     update_envR / update_envL                    pydmrg/tools/env_tools.py:40-50, 28-38
 *     davidson(aop, x0, precond, ...)              pydmrg/tools/la_tools.py:315-557
 *     renorm_inf (two-site SVD truncation)         pydmrg/tools/mps_tools.py:98-120
 *     renormalizeR / renormalizeL (RDM route)      pydmrg/dmrg.py:49-136
 * Each entry point below names the reference lines it replaces.  The reference-side binding
 * is a ctypes stub; pydmrg_b200/_lib.py is that stub.
 *
 * Conventions
 *   - every pointer argument is a DEVICE pointer unless its name ends in _host;
 *   - tensors are C-order (row-major) exactly as the reference's NumPy arrays;
 *   - dtype: PDMRG_F64 = float64, PDMRG_C128 = interleaved complex128 (NumPy layout);
 *   - all work is enqueued on the cudaStream_t passed as `void* stream`; no call
 *     synchronises the stream unless stated;
 *   - return value 0 = success, non-zero = error, message from pdmrg_last_error();
 *   - no hidden device allocation: scratch space is passed in by the caller (sizes from the
 *     *_workspace_bytes queries); plan objects own only a few KB of MPO coefficients.
 */
#ifndef PYDMRG_B200_H
#define PYDMRG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDMRG_F64 0
#define PDMRG_C128 1

/* ---- library ------------------------------------------------------------------------- */
const char* pdmrg_last_error(void);
int pdmrg_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long pdmrg_launch_count(void);
/* which GEMM path the last pdmrg_gemm_nt call took: 1 = DMMA+TMA tiles, 2 = generic */
int pdmrg_last_gemm_path(void);
/* force the generic (non-TMA) GEMM kernel; testing only */
void pdmrg_set_force_generic_gemm(int on);

/* ---- dense building block -------------------------------------------------------------
 * C[b] (M x N, ldc) = alpha * A[b] (M x K, lda) * op(B[b] (N x K, ldb))^T + beta * C[b]
 * op = conj when conjB != 0 (complex only).  Leading dimensions and batch strides are in
 * ELEMENTS of the dtype.  strideA / strideB = 0 broadcasts that operand over the batch.
 * beta must be 0 or 1.  This is the tile engine under every contraction of
 * tools/diag_tools.py:114-119,153-156 and tools/env_tools.py:35-37,47-49. */
int pdmrg_gemm_nt(int dtype, int M, int N, int K, int batch,
                  const void* A, long long lda, long long strideA,
                  const void* B, long long ldb, long long strideB, int conjB,
                  void* C, long long ldc, long long strideC,
                  double alpha_re, double alpha_im, int beta, void* stream);

/* Same product with TWO batch levels, b = (b1, b2): the operand / result offsets are
 * b1*stride?1 + idx2[b2]*stride?2 (idx2_host = NULL means idx2[b2] = b2; nb2 <= 32).  The index
 * list lets one launch visit only the non-empty MPO-bond slabs. */
int pdmrg_gemm_nt_batched2(int dtype, int M, int N, int K, int nb1, int nb2, const int* idx2_host,
                           const void* A, long long lda, long long strideA1, long long strideA2,
                           const void* B, long long ldb, long long strideB1, long long strideB2, int conjB,
                           void* C, long long ldc, long long strideC1, long long strideC2,
                           double alpha_re, double alpha_im, int beta, void* stream);
/* force the 128- or 64-row CTA tile (0 = automatic wave-quantisation model); testing only */
void pdmrg_set_force_tile_rows(int bm);
/* Stream-K scheduling of the tile kernel (csrc/gemm_nt.cu): 0 = never, 1 = by the cost model (default: problems whose
 * whole tiles fill the SMs badly, i.e. bond dimensions <= 512), 2 = whenever possible (tests).  The result does not
 * depend on the mode beyond rounding (fixed summation order: deterministic).  pdmrg_last_gemm_streamk(): whether the
 * last pdmrg_gemm_nt* launch used it. */
void pdmrg_set_streamk_mode(int mode);
int pdmrg_last_gemm_streamk(void);

/* ---- strided 4-index copy:  out[i0,i1,i2,i3] (C-order, dims n0..n3) = f(in[sum i_k*s_k])
 * f = conj if conj != 0; if scale != NULL the element is multiplied by the REAL vector
 * scale[i_{scale_axis}] (device pointer).  Strides in elements. */
int pdmrg_permute4(int dtype, const void* in, void* out,
                   int n0, int n1, int n2, int n3,
                   long long s0, long long s1, long long s2, long long s3,
                   int conj, const double* scale, int scale_axis, void* stream);

/* ---- effective-Hamiltonian mat-vec  (H1: tools/diag_tools.py:108-125,
 *                                       H2: tools/diag_tools.py:146-162) -------------------
 * One plan per local problem.  P = physical dimension of the local block (d for one-site,
 * d1*d2 for two-site).  For each MPO term m the caller passes the COMBINED local operator
 *   Wc[m] : dense (wL*P) x (wR*P) C-order, Wc[(j,pout),(o,pin)]
 *           one-site: Wc[(n,o),(j,l)] = W[n,j,o,l];
 *           two-site: Wc[(j,(m,p)),(o,(n,q))] = sum_l W1[j,l,m,n] W2[l,o,p,q]
 * (host pointer, dtype of the plan).  Zero entries are skipped (MPO block sparsity).
 * y[pout,i,r] = sign * sum_m sum L_m[i,j,k] Wc_m[(j,pout),(o,pin)] R_m[r,o,s] x[pin,k,s]
 * with L_m (Dl,wL,Dl), R_m (Dr,wR,Dr), x and y (P,Dl,Dr).  The reference's Hfun returns
 * -H x, i.e. sign = -1. */
typedef struct pdmrg_heff_plan pdmrg_heff_plan;
int pdmrg_heff_plan_create(pdmrg_heff_plan** plan, int dtype, int P, int Dl, int Dr, int n_mpo,
                           const int* wL_host, const int* wR_host, const void* const* Wc_host);
void pdmrg_heff_plan_destroy(pdmrg_heff_plan* plan);
size_t pdmrg_heff_workspace_bytes(const pdmrg_heff_plan* plan);
int pdmrg_heff_apply(pdmrg_heff_plan* plan, const void* const* L_host_ptrs, const void* const* R_host_ptrs,
                     const void* x, void* y, double sign, void* workspace, size_t workspace_bytes,
                     void* stream);
/* Same as pdmrg_heff_apply, but records CUDA events on `stream` around the three stages and
 * SYNCHRONISES: ms3_host[0..2] = milliseconds spent in stage A (GEMM), the W mix, stage C (GEMM),
 * summed over MPO terms.  Measurement aid for bench.py's roofline object. */
int pdmrg_heff_apply_timed(pdmrg_heff_plan* plan, const void* const* L_host_ptrs, const void* const* R_host_ptrs,
                           const void* x, void* y, double sign, void* workspace, size_t workspace_bytes,
                           void* stream, float* ms3_host);
/* MPO-bond-sharded variant: this rank applies only the (j,o) blocks whose
 * flag in block_mask_host[m][j*wR+o] is non-zero; y holds the PARTIAL sum (caller
 * all-reduces).  The mask is fixed at plan creation. */
int pdmrg_heff_plan_create_sharded(pdmrg_heff_plan** plan, int dtype, int P, int Dl, int Dr, int n_mpo,
                                   const int* wL_host, const int* wR_host, const void* const* Wc_host,
                                   const unsigned char* const* block_mask_host);
/* ---- bond-sharded mat-vec with the exchange fused into the kernels -----------------
 * One process per GPU.  Each rank allocates three buffers with pdmrg_ipc_alloc (slab buffer Z,
 * y staging buffer, flag array of 2*world uint64), ships the 64-byte handles to its peers through
 * any host channel (torch.distributed / MPI), opens the peers' handles and builds a communicator.
 * pdmrg_heff_apply_fused then runs stage A and the W mix locally, stage C with a SCATTER epilogue
 * that stores every output tile straight into the memory of the rank owning those rows of y
 * (peer st.global over NVLink), and a reduce + all-gather-by-peer-stores kernel guarded by
 * system-scope flags; every rank ends with a bit-identical y.  No NCCL call on this path.
 * jmask_all_ranks[r] = bit mask of the MPO-bond slabs j rank r produces (pdmrg_heff_needed_j). */
typedef struct pdmrg_comm pdmrg_comm;
int pdmrg_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64);
int pdmrg_ipc_open(const unsigned char* handle64, void** dev_ptr);
int pdmrg_ipc_close(void* dev_ptr);
int pdmrg_ipc_free(void* dev_ptr);
int pdmrg_comm_create(pdmrg_comm** comm, int rank, int world, void* const* Zpeer, void* const* Ypeer,
                      void* const* Fpeer);
void pdmrg_comm_destroy(pdmrg_comm* comm);
/* Sticky error word of this rank's flag buffer (the flag buffer needs (2*world + 1) * 8 bytes): non-zero when a flag wait
 * gave up after several seconds (a peer died or diverged).  Synchronises the stream. */
int pdmrg_comm_error(pdmrg_comm* comm, unsigned long long* err_out, void* stream);
size_t pdmrg_comm_z_bytes(const pdmrg_heff_plan* plan, int world);
size_t pdmrg_comm_y_bytes(const pdmrg_heff_plan* plan);
int pdmrg_heff_needed_j(const pdmrg_heff_plan* plan, int mpo, unsigned int* mask_out);
int pdmrg_heff_apply_fused(pdmrg_heff_plan* plan, pdmrg_comm* comm, const void* const* L_host_ptrs,
                           const void* const* R_host_ptrs, const void* x, void* y, double sign,
                           const unsigned int* jmask_all_ranks, void* workspace, size_t workspace_bytes,
                           void* stream);
/* the GEMM with the scatter epilogue on its own: rows [q*rows_per_owner, ...) of C go to Cpeer_host[q].
 * accumulate_b2 != 0: the nb2 second-level batch entries are summed into ONE output per b1 (their
 * k-ranges are concatenated inside the kernel's k-loop; strideC2 is ignored). */
int pdmrg_gemm_nt_scatter(int dtype, int M, int N, int K, int nb1, int nb2, const int* idx2_host,
                          const void* A, long long lda, long long strideA1, long long strideA2,
                          const void* B, long long ldb, long long strideB1, long long strideB2, int conjB,
                          const void* const* Cpeer_host, int n_owner, int rows_per_owner,
                          long long ldc, long long strideC1, long long strideC2,
                          double alpha_re, double alpha_im, int accumulate_b2, void* stream);
/* Ket-split mat-vec with the exchange fused into the kernels: `plan` from pdmrg_heff_plan_create_ketsplit
 * (this rank's ket range), stage C runs as ONE long-K GEMM (MPO bond folded into the k-loop) whose scatter
 * epilogue stores every tile of this rank's PARTIAL y straight into the owner rank's buffer
 * Z[src][pout][il][r]; the flag-guarded reduce + all-gather kernel then sums the `world` partials and
 * pushes the rows to every rank.  Z needs world * P * chunk * Dr elements (pdmrg_comm_z_bytes suffices). */
int pdmrg_heff_apply_fused_ket(pdmrg_heff_plan* plan, pdmrg_comm* comm, const void* const* L_host_ptrs,
                               const void* const* R_host_ptrs, const void* x, void* y, double sign,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Ket-split shard: this rank contracts only the ket indices k in [k0, k0+klen) of the LEFT bond
 * (x[pin, k, s], L[i, j, k]); x and y keep their full (P, Dl, Dr) shape, y holds the PARTIAL sum.
 * Perfectly balanced for any rank count (every rank does 1/n of both GEMM stages); combine with a
 * block mask (or pass NULL) as needed.  klen = 0 makes the rank a no-op (y = 0). */
int pdmrg_heff_plan_create_ketsplit(pdmrg_heff_plan** plan, int dtype, int P, int Dl, int Dr, int n_mpo,
                                    const int* wL_host, const int* wR_host, const void* const* Wc_host,
                                    const unsigned char* const* block_mask_host, int k0, int klen);
/* Plan creation WITHOUT any device allocation, legacy-stream copy or synchronisation (what a sweep uses: one plan per
 * bond step, possibly several sweeps running concurrently on different streams of one GPU).  The host-side operator
 * tables are built here; the caller then provides a device buffer of pdmrg_heff_plan_blob_bytes() bytes (256-byte
 * aligned, alive until the last apply has run) and the stream on which the small H2D copy is ordered.
 * block_mask_host may be NULL; k0 = 0, klen = -1 selects the whole left bond. */
int pdmrg_heff_plan_create_host(pdmrg_heff_plan** plan, int dtype, int P, int Dl, int Dr, int n_mpo,
                                const int* wL, const int* wR, const void* const* Wc_host,
                                const unsigned char* const* block_mask_host, int k0, int klen);
size_t pdmrg_heff_plan_blob_bytes(const pdmrg_heff_plan* plan);
int pdmrg_heff_plan_upload(pdmrg_heff_plan* plan, void* blob_dev, size_t blob_bytes, void* stream);
/* algorithmic FLOPs of one apply (dense-W count of SURVEY 8d) and with block sparsity */
double pdmrg_heff_flops_dense(const pdmrg_heff_plan* plan);
double pdmrg_heff_flops_sparse(const pdmrg_heff_plan* plan);
/* FLOPs the two GEMM stages of one apply actually execute (chi^3 terms of the non-empty slabs) */
double pdmrg_heff_flops_gemm(const pdmrg_heff_plan* plan);

/* ---- environment (block) updates  (E1: tools/env_tools.py:40-50, E2: :28-38) --------------
 * right-growing:  out[k,m,q] = sum L[j,l,p] Abra[i,j,k] W[l,m,i,n] A[n,p,q]
 *   A, Abra (d,Dl,Dr); L (Dl,wl,Dl); W host (wl,wr,d,d) dtype-typed or NULL (= identity,
 *   wr = wl); out (Dr,wr,Dr).  Abra = NULL means conj(A) (env_tools.py:41).
 * left-growing:   out[x,y,a] = sum B[e,a,f] R[c,d,f] W[y,d,b,e] Bbra[b,x,c]
 *   B, Bbra (d,Dl,Dr); R (Dr,wr,Dr); out (Dl,wl,Dl). */
size_t pdmrg_env_workspace_bytes(int dtype, int d, int Dl, int Dr, int wl, int wr);
int pdmrg_env_update_right(int dtype, int d, int Dl, int Dr, int wl, int wr, const void* W_host,
                           const void* L, const void* A, const void* Abra, void* out,
                           void* workspace, size_t workspace_bytes, void* stream);
int pdmrg_env_update_left(int dtype, int d, int Dl, int Dr, int wl, int wr, const void* W_host,
                          const void* R, const void* B, const void* Bbra, void* out,
                          void* workspace, size_t workspace_bytes, void* stream);
/* Column-restricted variants for sharded sweeps (same contractions, tools/env_tools.py:40-50 /
 * 28-38): only the entries of the new block whose KET index (last index of `out`) lies in [k0, k0+kn) are computed
 * and stored, in place, into the full-size `out` (D_bra, w, D_ket); everything else in `out` is left untouched.
 * Ranks owning disjoint ket ranges build the whole block with one exchange.  Workspace: pdmrg_env_workspace_bytes.
 * k0 should be a multiple of 2 for f64 (16-byte aligned stores). */
int pdmrg_env_update_right_cols(int dtype, int d, int Dl, int Dr, int wl, int wr, const void* W_host,
                                const void* L, const void* A, const void* Abra, int k0, int kn, void* out,
                                void* workspace, size_t workspace_bytes, void* stream);
int pdmrg_env_update_left_cols(int dtype, int d, int Dl, int Dr, int wl, int wr, const void* W_host,
                               const void* R, const void* B, const void* Bbra, int k0, int kn, void* out,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---- Krylov vector kernels for the Davidson loop (tools/la_tools.py:449-457, 469-470,
 *      479-482, 524-536).  Vectors have n elements; a "basis" is m vectors at X + i*stride. */
/* out_dev[i] = <X_i | y> = sum conj(X_i) y, i < m   (out_dev: m elements of dtype) */
int pdmrg_vec_dots(int dtype, long long n, int m, const void* X, long long stride,
                   const void* y, void* out_dev, void* scratch, size_t scratch_bytes, void* stream);
/* out = sum_i coef[i] X_i  (+ out if accumulate);  coef_host: m elements of dtype */
int pdmrg_vec_combine(int dtype, long long n, int m, const void* coef_host, const void* X,
                      long long stride, void* out, int accumulate, void* stream);
/* r = sum_i v[i] (AX_i - e XS_i);  norm2_dev[0] = ||r||^2 (double).  v_host, e_host of dtype */
int pdmrg_davidson_residual(int dtype, long long n, int m, const void* v_host, const void* e_host,
                            const void* XS, const void* AX, long long stride, void* r,
                            double* norm2_dev, void* scratch, size_t scratch_bytes, void* stream);
/* r = (r - sum_i c_dev[i] X_i) * (scale_dev ? scale_dev[0]^-1/2 : 1) ; norm2_dev[0] = ||r||^2 of the result.
 * Coefficients stay on the device so that dots -> project needs no host round trip. */
int pdmrg_vec_project(int dtype, long long n, int m, const void* X, long long stride,
                      const void* c_dev, const double* scale_dev, void* r,
                      double* norm2_dev, void* scratch, size_t scratch_bytes, void* stream);
/* out = in * (norm2_dev[0])^-1/2  (out may alias in) */
int pdmrg_vec_normalize(int dtype, long long n, const void* in, const double* norm2_dev,
                        void* out, void* stream);
/* Scratch for the reductions above.  Its last 256 bytes hold the ticket of the fused second reduction
 * stage: the buffer must be ZERO-filled once before first use (the kernels leave the ticket at zero), and
 * one scratch buffer must not be shared by kernels running concurrently on different streams. */
size_t pdmrg_vec_scratch_bytes(int m);

/* ---- the whole local eigensolve behind one call (K + D: tools/la_tools.py:402-557 as driven by
 * tools/diag_tools.py:243-290) -------------------------------------------------------------------
 * Restarted non-symmetric Davidson on the operator  sign * H_eff  of `plan` (the reference minimises
 * Re of -H, i.e. sign = -1): same subspace growth, restart rule (space + nroots > 12 + 3*nroots),
 * lindep pruning and stopping rule as the reference; res_tol < 0 keeps the reference's effective
 * criterion ||r||^2 <= lindep.  x0: nx0 <= nroots guesses (device, x0_stride apart).  XS / AX:
 * basis and image slabs with >= max_space + 4*nroots rows of `stride` elements; Rv: 2*nroots + 2
 * rows.  e_out_host: nroots (re, im) pairs; vecs_out: nroots Ritz vectors.  SYNCHRONISES the stream
 * twice per cycle (O(m) scalars).  The projected eigenproblem is solved on the host by a complex
 * Hessenberg-QR (pdmrg_host_eig exposes it for testing). */
size_t pdmrg_davidson_scratch_bytes(int nroots, int max_space);
int pdmrg_davidson(pdmrg_heff_plan* plan, const void* const* L_host_ptrs, const void* const* R_host_ptrs, double sign,
                   int dtype, long long n, const void* x0, long long x0_stride, int nx0, int nroots, int max_space,
                   double lindep, double res_tol, int max_cycle, int fixed_cycles,
                   void* XS, void* AX, void* Rv, long long stride,
                   void* heff_workspace, size_t heff_workspace_bytes, void* scratch, size_t scratch_bytes,
                   double* e_out_host, void* vecs_out, long long vecs_stride,
                   int* n_matvec_out, int* cycles_out, double* last_residual_out, void* stream);
/* The same loop with a diagonal preconditioner in the correction step (opt-in: iteration counts then differ from
 * the reference, whose preconditioner is the identity, tools/diag_tools.py:137-139; it acts at la_tools.py:510):
 *   t = r / (diag - theta),  |diag_i - theta| kept >= pc_floor (absolute, in units of the operator: the spacing of the
 *   diagonal is set by the local couplings, not by the total eigenvalue, which grows with the chain length),
 * `diag` = diagonal of the operator sign * H_eff (n elements of dtype, e.g. from pdmrg_heff_diag with the same sign).
 * Safeguard: if the residual is not halved within two restart periods the loop continues with the identity. */
int pdmrg_davidson_precond(pdmrg_heff_plan* plan, const void* const* L_host_ptrs, const void* const* R_host_ptrs, double sign,
                           int dtype, long long n, const void* x0, long long x0_stride, int nx0, int nroots, int max_space,
                           double lindep, double res_tol, int max_cycle, int fixed_cycles,
                           void* XS, void* AX, void* Rv, long long stride,
                           void* heff_workspace, size_t heff_workspace_bytes, void* scratch, size_t scratch_bytes,
                           double* e_out_host, void* vecs_out, long long vecs_stride,
                           int* n_matvec_out, int* cycles_out, double* last_residual_out, void* stream,
                           const void* diag, double pc_floor);
/* The same loop with the solver options that go beyond the reference's algorithm (all opt-in; csrc/davidson.cu):
 *   restart_keep > 0 : thick restart -- at a restart the basis is compressed to an orthonormal basis of the
 *                      restart_keep lowest Ritz vectors (XS' = XS Q, AX' = AX Q, no new mat-vec) instead of being thrown
 *                      away as la_tools.py:544-546 does; XS / AX then need max_space + 4*nroots + restart_keep rows;
 *   tol_relative     : res_tol is relative to |theta| (ARPACK's rule; scipy eigs(tol=1e-5), tools/diag_tools.py:225);
 *   diag             : optional diagonal preconditioner as in pdmrg_davidson_precond (NULL = the reference's identity);
 *   stagnation > 0   : stop a solve whose residual has not been halved within that many cycles (the reference goes on to
 *                      max_cycle = 1000; nearly defective local problems next to s = 0 stall at ||r|| ~ 1e-4).
 * With the identity preconditioner the subspace is the Krylov space of the start vector, so restart_keep > 0 is
 * thick-restarted Arnoldi (Krylov-Schur): the device-resident counterpart of calc_eigs_arnoldi (diag_tools.py:214-241). */
int pdmrg_davidson_ex(pdmrg_heff_plan* plan, const void* const* L_host_ptrs, const void* const* R_host_ptrs, double sign,
                      int dtype, long long n, const void* x0, long long x0_stride, int nx0, int nroots, int max_space,
                      double lindep, double res_tol, int max_cycle, int fixed_cycles,
                      void* XS, void* AX, void* Rv, long long stride,
                      void* heff_workspace, size_t heff_workspace_bytes, void* scratch, size_t scratch_bytes,
                      double* e_out_host, void* vecs_out, long long vecs_stride,
                      int* n_matvec_out, int* cycles_out, double* last_residual_out, void* stream,
                      const void* diag, double pc_floor, int restart_keep, int tol_relative, int* n_restarts_out,
                      int stagnation);
/* pdmrg_davidson_ex over a SHARDED operator: `plan` = this rank's ket-split shard, every mat-vec =
 * pdmrg_heff_apply_fused_ket through `comm`.  Identical start vectors and options on every rank; the ranks then run in
 * lock step without exchanging control scalars (bit-identical images, fixed-order reductions). */
int pdmrg_davidson_sharded(pdmrg_heff_plan* plan, pdmrg_comm* comm, const void* const* L_host_ptrs, const void* const* R_host_ptrs,
                           double sign, int dtype, long long n, const void* x0, long long x0_stride, int nx0, int nroots,
                           int max_space, double lindep, double res_tol, int max_cycle, int fixed_cycles,
                           void* XS, void* AX, void* Rv, long long stride,
                           void* heff_workspace, size_t heff_workspace_bytes, void* scratch, size_t scratch_bytes,
                           double* e_out_host, void* vecs_out, long long vecs_stride,
                           int* n_matvec_out, int* cycles_out, double* last_residual_out, void* stream,
                           const void* diag, double pc_floor, int restart_keep, int tol_relative, int* n_restarts_out,
                           int stagnation);
/* r <- r / (diag - theta) element-wise with the floor above; norm2_dev[0] = ||r||^2 of the result.  scratch as for the
 * other vector kernels (pdmrg_vec_scratch_bytes). */
int pdmrg_vec_precond(int dtype, long long n, void* r, const void* diag, double theta_re, double theta_im,
                      double floor_abs, double* norm2_dev, void* scratch, size_t scratch_bytes, void* stream);
/* diag[p,k,s] (+)= sign * sum_{j,o} L[k,j,k] Wpp[p,j,o] R[s,o,s]: diagonal of one MPO term of H_eff (H1/H2 with
 * P = d or d1*d2); Wpp (P, wL, wR) on the DEVICE holds the physical-diagonal blocks Wc[(j,p),(o,p)] of the combined
 * local operator.  accumulate != 0 adds to `out` (several MPO terms). */
int pdmrg_heff_diag(int dtype, int P, int Dl, int Dr, int wL, int wR, const void* L, const void* R,
                    const void* Wpp_dev, double sign, int accumulate, void* out, void* stream);
/* eigenvalues w (n complex) and unit-norm right eigenvectors (columns of v) of a small general
 * complex matrix, column-major interleaved; host only (no GPU needed) */
int pdmrg_host_eig(int n, const double* a_colmajor_interleaved, double* w_out, double* v_out);

/* ---- truncation ---------------------------------------------------------------------------
 * T2 (tools/mps_tools.py:98-120): thin SVD of a row-major (rows x cols) matrix a:
 *   a = U diag(S) Vh,  U (rows x k) row-major, S (k) descending, Vh (k x cols) row-major,
 *   k = min(rows, cols).  `a` is destroyed (methods 0, 1).  method: 0 = QR-iteration (gesvd),
 *   1 = Jacobi (gesvdj), 2 = Gram matrix (DMMA GEMM) + heevd + Householder QR with a gated tail
 *   refinement level: both factors are exact isometries, reconstruction error <= 1e-9*S_0
 *   (typically 1e-12, see csrc/solver.cu); 5-10x faster at 2048^2; SYNCHRONISES the stream.  info_dev: one int
 *   (device).  cuSOLVER is loaded on first use. */
size_t pdmrg_svd_workspace_bytes(int dtype, int rows, int cols, int method);
int pdmrg_svd(int dtype, int rows, int cols, void* a, void* U, double* S, void* Vh,
              int method, void* workspace, size_t workspace_bytes, int* info_dev, void* stream);
/* Same, when the caller will only use the `keep` leading singular triplets (chi truncation):
 * U (rows x k), S (k), Vh (k x cols) keep their full-size layout but only columns / entries /
 * rows [0, keep) are defined for method 2 (S[keep:] = 0); methods 0 and 1 compute everything. */
int pdmrg_svd_topk(int dtype, int rows, int cols, int keep, void* a, void* U, double* S, void* Vh,
                   int method, void* workspace, size_t workspace_bytes, int* info_dev, void* stream);
/* Left singular basis of a row-major X (n x c) for projective truncation of a moving centre:
 * Ut (n x n, row a = eigenvector u_a of X X^H, eigenvalues descending) and lam[a] = S_a^2, with the
 * small-singular-value tail re-diagonalised from the projected Gram matrix (directions and
 * weights accurate to ~1e-12 S_0, see csrc/solver.cu) whenever one of the `keep` leading vectors
 * lies in it.  X is read-only.  SYNCHRONISES the stream.  Returns 77 if cuSOLVER's heevd did not
 * converge (caller falls back to pdmrg_svd method 0). */
size_t pdmrg_left_basis_workspace_bytes(int dtype, int n, int c);
int pdmrg_left_basis(int dtype, int n, int c, int keep, const void* X, void* Ut, double* lam,
                     void* workspace, size_t workspace_bytes, int* info_dev, void* stream);
/* Warm-started replacement for pdmrg_left_basis in a converging sweep (same role as the SVD of
 * tools/mps_tools.py:106 for a moving centre; see csrc/solver.cu "Recycled ... step"):
 *   T = Wt X1^H (K x a)  ->  Householder QR keeping the column order  ->  Pt (K x a, orthonormal rows),
 *   Ct = Pt X2^T (K x c),  then rows [k-b, K) of Pt and Ct are rotated by the exact Rayleigh-Ritz
 *   vectors of that window (sorted by decreasing weight);  S[i] = ||Ct[i, :]||.
 * X1 (a x c) and X2 = X1^T (c x a) are row-major and read-only, Wt (K x c) holds the previous basis
 * (k kept rows, then K-k spare rows, ordered by weight).  Needs K <= min(a, c).  The caller keeps rows
 * [0, k).  method 0: Householder QR (cuSOLVER geqrf/ungqr); method 1: Cholesky-QR2 on the row-normalised T
 * (DMMA GEMMs + potrf + trsm; same triangular structure, ~2x faster, valid while the normalised rows are
 * not numerically dependent).  S has K + 2 entries: S[K] = ||P1 P1^H - I||_F^2 after the first Cholesky
 * pass (the caller must fall back to method 0 when it is not << 1), S[K+1] = number of non-zero cuSOLVER
 * infos.  info_dev: int[4] = {geqrf/ungqr, potrf pass 1, potrf pass 2, window eigensolver}. */
size_t pdmrg_recycle_basis_workspace_bytes(int dtype, int a, int c, int K, int k, int b);
int pdmrg_recycle_basis(int dtype, int a, int c, int K, int k, int b, int method, const void* X1,
                        const void* X2, const void* Wt, void* Pt, void* Ct, double* S, void* workspace,
                        size_t workspace_bytes, int* info_dev, void* stream);
/* T1 (dmrg.py:64-76): eigen-decomposition of a Hermitian (n x n) row-major matrix, in place:
 * on exit row i of `a` is the i-th eigenvector (ascending eigenvalues in w). */
size_t pdmrg_heevd_workspace_bytes(int dtype, int n);
int pdmrg_heevd(int dtype, int n, void* a, double* w, void* workspace, size_t workspace_bytes,
                int* info_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
