// Truncation back-ends (T2: pydmrg/tools/mps_tools.py:106 np.linalg.svd;
// T1: pydmrg/dmrg.py:68 eig of the reduced density matrix) through cuSOLVER, which is
// dlopen'ed on first use so that the rest of the library has no link-time dependency on it.
// cuSOLVER is column-major; the reference's matrices are row-major.  A row-major (r x c)
// array is the column-major (c x r) transpose, so no data is moved when c >= r:
//   a^T = Uc S VTc   =>   a = VTc^T S Uc^T,   row-major U = VTc buffer, row-major Vh = Uc buffer.
#include "common.cuh"

#include <cublas_v2.h>
#include <cusolverDn.h>
#include <cstdlib>
#include <dlfcn.h>
#include <mutex>
#include <vector>

namespace pdmrg {
namespace {

struct Solver {
  void* lib = nullptr;
  cusolverDnHandle_t handle = nullptr;
  gesvdjInfo_t jinfo = nullptr;
#define PDMRG_SOLVER_FN(name) decltype(&name) p_##name = nullptr;
  PDMRG_SOLVER_FN(cusolverDnCreate)
  PDMRG_SOLVER_FN(cusolverDnSetStream)
  PDMRG_SOLVER_FN(cusolverDnZgesvd_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnDgesvd_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnZgesvd)
  PDMRG_SOLVER_FN(cusolverDnDgesvd)
  PDMRG_SOLVER_FN(cusolverDnCreateGesvdjInfo)
  PDMRG_SOLVER_FN(cusolverDnXgesvdjSetTolerance)
  PDMRG_SOLVER_FN(cusolverDnXgesvdjSetMaxSweeps)
  PDMRG_SOLVER_FN(cusolverDnXgesvdjSetSortEig)
  PDMRG_SOLVER_FN(cusolverDnZgesvdj_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnDgesvdj_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnZgesvdj)
  PDMRG_SOLVER_FN(cusolverDnDgesvdj)
  PDMRG_SOLVER_FN(cusolverDnZheevd_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnDsyevd_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnZheevd)
  PDMRG_SOLVER_FN(cusolverDnDsyevd)
  PDMRG_SOLVER_FN(cusolverDnZgeqrf_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnDgeqrf_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnZgeqrf)
  PDMRG_SOLVER_FN(cusolverDnDgeqrf)
  PDMRG_SOLVER_FN(cusolverDnZungqr_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnDorgqr_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnZungqr)
  PDMRG_SOLVER_FN(cusolverDnDorgqr)
  PDMRG_SOLVER_FN(cusolverDnZpotrf_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnDpotrf_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnZpotrf)
  PDMRG_SOLVER_FN(cusolverDnDpotrf)
  PDMRG_SOLVER_FN(cusolverDnCreateSyevjInfo)
  PDMRG_SOLVER_FN(cusolverDnXsyevjSetTolerance)
  PDMRG_SOLVER_FN(cusolverDnXsyevjSetMaxSweeps)
  PDMRG_SOLVER_FN(cusolverDnXsyevjSetSortEig)
  PDMRG_SOLVER_FN(cusolverDnZheevj_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnDsyevj_bufferSize)
  PDMRG_SOLVER_FN(cusolverDnZheevj)
  PDMRG_SOLVER_FN(cusolverDnDsyevj)
#undef PDMRG_SOLVER_FN
  syevjInfo_t evj = nullptr;
  // cuBLAS (triangular solves of the Cholesky-QR route), dlopen'ed like cuSOLVER
  void* blas_lib = nullptr;
  cublasHandle_t blas = nullptr;
  decltype(&cublasCreate_v2) p_cublasCreate = nullptr;
  decltype(&cublasSetStream_v2) p_cublasSetStream = nullptr;
  decltype(&cublasZtrsm_v2) p_cublasZtrsm = nullptr;
  decltype(&cublasDtrsm_v2) p_cublasDtrsm = nullptr;
};

Solver* get_solver() {
  static Solver s;
  static std::once_flag once;
  static bool ok = false;
  std::call_once(once, [] {
    const char* names[] = {"libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so"};
    for (const char* n : names) {
      s.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (s.lib) break;
    }
    if (!s.lib) { set_error("cannot dlopen libcusolver: %s", dlerror()); return; }
    bool all = true;
#define PDMRG_LOAD(name)                                                     \
  s.p_##name = reinterpret_cast<decltype(&name)>(dlsym(s.lib, #name));       \
  if (!s.p_##name) { set_error("cusolver symbol %s missing", #name); all = false; }
    PDMRG_LOAD(cusolverDnCreate)
    PDMRG_LOAD(cusolverDnSetStream)
    PDMRG_LOAD(cusolverDnZgesvd_bufferSize)
    PDMRG_LOAD(cusolverDnDgesvd_bufferSize)
    PDMRG_LOAD(cusolverDnZgesvd)
    PDMRG_LOAD(cusolverDnDgesvd)
    PDMRG_LOAD(cusolverDnCreateGesvdjInfo)
    PDMRG_LOAD(cusolverDnXgesvdjSetTolerance)
    PDMRG_LOAD(cusolverDnXgesvdjSetMaxSweeps)
    PDMRG_LOAD(cusolverDnXgesvdjSetSortEig)
    PDMRG_LOAD(cusolverDnZgesvdj_bufferSize)
    PDMRG_LOAD(cusolverDnDgesvdj_bufferSize)
    PDMRG_LOAD(cusolverDnZgesvdj)
    PDMRG_LOAD(cusolverDnDgesvdj)
    PDMRG_LOAD(cusolverDnZheevd_bufferSize)
    PDMRG_LOAD(cusolverDnDsyevd_bufferSize)
    PDMRG_LOAD(cusolverDnZheevd)
    PDMRG_LOAD(cusolverDnDsyevd)
    PDMRG_LOAD(cusolverDnZgeqrf_bufferSize)
    PDMRG_LOAD(cusolverDnDgeqrf_bufferSize)
    PDMRG_LOAD(cusolverDnZgeqrf)
    PDMRG_LOAD(cusolverDnDgeqrf)
    PDMRG_LOAD(cusolverDnZungqr_bufferSize)
    PDMRG_LOAD(cusolverDnDorgqr_bufferSize)
    PDMRG_LOAD(cusolverDnZungqr)
    PDMRG_LOAD(cusolverDnDorgqr)
    PDMRG_LOAD(cusolverDnZpotrf_bufferSize)
    PDMRG_LOAD(cusolverDnDpotrf_bufferSize)
    PDMRG_LOAD(cusolverDnZpotrf)
    PDMRG_LOAD(cusolverDnDpotrf)
    PDMRG_LOAD(cusolverDnCreateSyevjInfo)
    PDMRG_LOAD(cusolverDnXsyevjSetTolerance)
    PDMRG_LOAD(cusolverDnXsyevjSetMaxSweeps)
    PDMRG_LOAD(cusolverDnXsyevjSetSortEig)
    PDMRG_LOAD(cusolverDnZheevj_bufferSize)
    PDMRG_LOAD(cusolverDnDsyevj_bufferSize)
    PDMRG_LOAD(cusolverDnZheevj)
    PDMRG_LOAD(cusolverDnDsyevj)
#undef PDMRG_LOAD
    if (!all) return;
    {
      const char* bn[] = {"libcublas.so.12", "/usr/local/cuda/lib64/libcublas.so.12", "libcublas.so"};
      for (const char* n : bn) {
        s.blas_lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (s.blas_lib) break;
      }
      if (!s.blas_lib) { set_error("cannot dlopen libcublas: %s", dlerror()); return; }
      s.p_cublasCreate = reinterpret_cast<decltype(&cublasCreate_v2)>(dlsym(s.blas_lib, "cublasCreate_v2"));
      s.p_cublasSetStream = reinterpret_cast<decltype(&cublasSetStream_v2)>(dlsym(s.blas_lib, "cublasSetStream_v2"));
      s.p_cublasZtrsm = reinterpret_cast<decltype(&cublasZtrsm_v2)>(dlsym(s.blas_lib, "cublasZtrsm_v2"));
      s.p_cublasDtrsm = reinterpret_cast<decltype(&cublasDtrsm_v2)>(dlsym(s.blas_lib, "cublasDtrsm_v2"));
      if (!s.p_cublasCreate || !s.p_cublasSetStream || !s.p_cublasZtrsm || !s.p_cublasDtrsm) { set_error("cublas symbols missing"); return; }
    }
    ok = true;
  });
  if (!ok) return nullptr;
  // Handles (and their stream binding) are per host thread: several scan points may be driven concurrently from
  // different threads on different streams of one GPU (pydmrg_b200/scan.py workers), and a cuSOLVER / cuBLAS handle
  // carries its stream.
  thread_local Solver tl;
  thread_local int tl_state = 0;          // 0 = not created, 1 = ready, -1 = failed
  if (tl_state == 0) {
    tl = s;
    tl_state = -1;
    if (tl.p_cublasCreate(&tl.blas) != CUBLAS_STATUS_SUCCESS) { set_error("cublasCreate failed"); return nullptr; }
    if (tl.p_cusolverDnCreate(&tl.handle) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnCreate failed"); return nullptr; }
    if (tl.p_cusolverDnCreateGesvdjInfo(&tl.jinfo) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnCreateGesvdjInfo failed"); return nullptr; }
    tl.p_cusolverDnXgesvdjSetTolerance(tl.jinfo, 1e-15);
    tl.p_cusolverDnXgesvdjSetMaxSweeps(tl.jinfo, 100);
    tl.p_cusolverDnXgesvdjSetSortEig(tl.jinfo, 1);
    if (tl.p_cusolverDnCreateSyevjInfo(&tl.evj) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnCreateSyevjInfo failed"); return nullptr; }
    tl.p_cusolverDnXsyevjSetTolerance(tl.evj, 1e-15);
    tl.p_cusolverDnXsyevjSetMaxSweeps(tl.evj, 100);
    tl.p_cusolverDnXsyevjSetSortEig(tl.evj, 1);
    tl_state = 1;
  }
  return tl_state == 1 ? &tl : nullptr;
}

struct SvdLayout {
  bool transpose;          // rows > cols: cuSOLVER works on an explicit transpose
  int p, q, k;             // column-major problem is p x q, p >= q
  size_t at, uc, vc, lwork_bytes;
  size_t total() const { return at + uc + vc + lwork_bytes; }
};

int svd_layout(Solver* s, int dtype, int rows, int cols, int method, SvdLayout* l) {
  const size_t eb = elem_bytes(dtype);
  l->transpose = rows > cols;
  l->p = l->transpose ? rows : cols;
  l->q = l->transpose ? cols : rows;
  l->k = l->q;
  l->at = l->transpose ? align_up((size_t)rows * cols * eb, 256) : 0;
  // scratch copies of the factors whenever they need a permutation afterwards
  l->uc = align_up((size_t)l->p * l->k * eb, 256);
  l->vc = align_up((size_t)l->q * l->k * eb, 256);
  int lwork = 0;
  cusolverStatus_t st;
  if (method == 0) {
    st = dtype == PDMRG_C128 ? s->p_cusolverDnZgesvd_bufferSize(s->handle, l->p, l->q, &lwork)
                             : s->p_cusolverDnDgesvd_bufferSize(s->handle, l->p, l->q, &lwork);
  } else {
    if (dtype == PDMRG_C128)
      st = s->p_cusolverDnZgesvdj_bufferSize(s->handle, CUSOLVER_EIG_MODE_VECTOR, 1, l->p, l->q, nullptr, l->p, nullptr,
                                             nullptr, l->p, nullptr, l->q, &lwork, s->jinfo);
    else
      st = s->p_cusolverDnDgesvdj_bufferSize(s->handle, CUSOLVER_EIG_MODE_VECTOR, 1, l->p, l->q, nullptr, l->p, nullptr,
                                             nullptr, l->p, nullptr, l->q, &lwork, s->jinfo);
  }
  if (st != CUSOLVER_STATUS_SUCCESS) { set_error("cusolver bufferSize failed (%d)", (int)st); return 1; }
  // gesvd (QR iteration) additionally needs rwork of min(m,n)-1 reals for complex input
  l->lwork_bytes = align_up((size_t)lwork * eb, 256) + align_up((size_t)l->q * 8 + 8, 256);
  return 0;
}

// ---- method 2: Gram matrix + Hermitian eigensolver + Householder QR ----------------------
// For a (rows x cols), rows <= cols:  G = a a^H (DMMA GEMM) -> heevd -> U (exactly unitary);
// Y = a^H U = V S has orthogonal columns of norm S_i; its Householder QR  Y = Q R  gives an
// exactly isometric V = Q phase with S_i = |R_ii|.  Both factors are isometries to machine
// precision.  A single Gram level resolves direction i only to eps S_0^2 / S_i, so the tail
// (S_i < 1e-4 of the block's largest) is re-diagonalised from the Gram matrix of the PROJECTED
// rows U_tail^H a, whose norm is 1e-8 S_0^2: one refinement level brings the reconstruction error
// to ~1e-12 S_0 for every singular value above 1e-12 S_0 (below that is rounding noise of the
// input).  The level is skipped when the accurate row norms |u_i^H a| show that the kept tail
// vectors cannot be off by more than 1e-9 S_0 in total.  5-10x faster than gesvd at 2048 x 2048
// complex128 (tools/svd_bench.py).
template <bool CPLX>
__global__ void qr_diag_kernel(const double* __restrict__ Y, long long ld, int k, double* __restrict__ S,
                               double* __restrict__ phase) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= k) return;
  if (CPLX) {
    const double re = Y[2 * (a * ld + a)], im = Y[2 * (a * ld + a) + 1];
    const double r = hypot(re, im);
    S[a] = r;
    phase[2 * a] = r > 0 ? re / r : 1.0;
    phase[2 * a + 1] = r > 0 ? im / r : 0.0;
  } else {
    const double re = Y[a * ld + a];
    S[a] = fabs(re);
    phase[a] = re < 0 ? -1.0 : 1.0;
  }
}

// Vh[a,c] = conj(Qt[a,c] * phase_a)
template <bool CPLX>
__global__ void vh_from_q_kernel(const double* __restrict__ Qt, const double* __restrict__ phase, int k, long long cols,
                                 double* __restrict__ Vh) {
  const long long total = (long long)k * cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(e / cols);
    if (CPLX) {
      const double2 q = reinterpret_cast<const double2*>(Qt)[e];
      const double2 ph = reinterpret_cast<const double2*>(phase)[a];
      const double2 v = cmul(q, ph);
      reinterpret_cast<double2*>(Vh)[e] = make_double2(v.x, -v.y);
    } else {
      Vh[e] = Qt[e] * phase[a];
    }
  }
}

// G_ii += rel * mean(diag G): lifts entries that underflowed towards the denormal range (rows of
// norm ~1e-85 appear after zero-padding in a chi ramp and make cuSOLVER's heevd fail to converge);
// rel = 1e-28 is far below every accuracy the factorisation delivers.
template <bool CPLX>
__global__ void __launch_bounds__(256) shift_diag_kernel(double* G, int n, double rel) {
  __shared__ double sm[256];
  constexpr int E = CPLX ? 2 : 1;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += G[(size_t)E * ((size_t)i * n + i)];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  const double delta = rel * fabs(sm[0]) / n;
  for (int i = threadIdx.x; i < n; i += 256) G[(size_t)E * ((size_t)i * n + i)] += delta;
}

struct GramLayout {
  size_t G, Ut, mT, Y, tau, phase, tail, lwork_bytes;   // tail: 6 staging blocks for the refinement levels
  int lwork_elems;
  size_t total() const { return G + Ut + mT + Y + tau + phase + 6 * tail + lwork_bytes; }
};

int gram_layout(Solver* s, int dtype, int rows, int cols, GramLayout* l) {
  const size_t eb = elem_bytes(dtype);
  const int k = rows;
  l->G = align_up((size_t)rows * rows * eb, 256);
  l->Ut = align_up((size_t)k * rows * eb, 256);
  l->mT = align_up((size_t)cols * rows * eb, 256);
  l->Y = align_up((size_t)k * cols * eb, 256);
  l->tau = align_up((size_t)k * eb, 256);
  l->phase = align_up((size_t)k * 16, 256);
  l->tail = align_up((size_t)k * (size_t)(cols > rows ? cols : rows) * eb, 256);   // staging blocks: up to k x max(rows, cols)
  int lw_e = 0, lw_q = 0, lw_g = 0;
  cusolverStatus_t st;
  if (dtype == PDMRG_C128) {
    st = s->p_cusolverDnZheevd_bufferSize(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, rows, nullptr, rows, nullptr, &lw_e);
    if (st == CUSOLVER_STATUS_SUCCESS) st = s->p_cusolverDnZgeqrf_bufferSize(s->handle, cols, k, nullptr, cols, &lw_q);
    if (st == CUSOLVER_STATUS_SUCCESS) st = s->p_cusolverDnZungqr_bufferSize(s->handle, cols, k, k, nullptr, cols, nullptr, &lw_g);
  } else {
    st = s->p_cusolverDnDsyevd_bufferSize(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, rows, nullptr, rows, nullptr, &lw_e);
    if (st == CUSOLVER_STATUS_SUCCESS) st = s->p_cusolverDnDgeqrf_bufferSize(s->handle, cols, k, nullptr, cols, &lw_q);
    if (st == CUSOLVER_STATUS_SUCCESS) st = s->p_cusolverDnDorgqr_bufferSize(s->handle, cols, k, k, nullptr, cols, nullptr, &lw_g);
  }
  if (st != CUSOLVER_STATUS_SUCCESS) { set_error("cusolver bufferSize (gram SVD) failed (%d)", (int)st); return 1; }
  int lw = lw_e > lw_q ? lw_e : lw_q;
  if (lw_g > lw) lw = lw_g;
  l->lwork_elems = lw;
  l->lwork_bytes = align_up((size_t)lw * eb, 256) + 256;
  return 0;
}

int heevd_inplace(Solver* s, int dtype, int n, double* A, double* w, void* work, int lwork, int* info_dev,
                  cudaStream_t stream) {
  if (dtype == PDMRG_C128) shift_diag_kernel<true><<<1, 256, 0, stream>>>(A, n, 1e-28);
  else shift_diag_kernel<false><<<1, 256, 0, stream>>>(A, n, 1e-28);
  PDMRG_CHECK_LAUNCH();
  cusolverStatus_t st;
  if (dtype == PDMRG_C128)
    st = s->p_cusolverDnZheevd(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, (cuDoubleComplex*)A, n, w,
                               (cuDoubleComplex*)work, lwork, info_dev);
  else
    st = s->p_cusolverDnDsyevd(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, w, (double*)work, lwork, info_dev);
  if (st != CUSOLVER_STATUS_SUCCESS) { set_error("gram SVD: heevd failed (%d)", (int)st); return 1; }
  count_launch();
  return 0;
}

constexpr int GRAM_FALLBACK = 77;
constexpr double GRAM_TAU2 = 1e-8;     // tail = eigenvalues below tau^2 = (1e-4)^2 of the block's largest
constexpr int GRAM_MAX_REFINE = 1;   // one refinement level already reaches ~1e-12 S_0 (tau * sqrt(eps) = 1e-12)

// rows <= cols.  a is read only.  Synchronises `stream` (reads eigenvalues to place the tail cuts).
// basis_only: stop after the (refined) left basis -- U receives Ut (k x rows, row a = eigenvector u_a of
// a a^H, descending) and S the eigenvalues lambda_a = S_a^2 (refined in the tail); Vh is not touched.
int svd_gram_wide(Solver* s, int dtype, int rows, int cols, int keep, const void* a, void* U, double* S, void* Vh,
                  unsigned char* ws, int* info_dev, cudaStream_t stream, bool basis_only = false) {
  GramLayout l;
  if (gram_layout(s, dtype, rows, cols, &l)) return 1;
  const bool cplx = dtype == PDMRG_C128;
  const int k = rows;
  const int nkeep = (keep <= 0 || keep > k) ? k : keep;      // leading singular triplets the caller uses
  const int E = cplx ? 2 : 1;
  const size_t eb = elem_bytes(dtype);
  unsigned char* q = ws;
  auto take = [&](size_t bytes) { double* r = reinterpret_cast<double*>(q); q += bytes; return r; };
  double* G = take(l.G);
  double* Ut = take(l.Ut);
  double* mT = take(l.mT);
  double* Y = take(l.Y);
  double* tau = take(l.tau);
  double* phase = take(l.phase);
  double* Uc = take(l.tail);      // conj of the tail rows of Ut
  double* X2 = take(l.tail);      // tail rows of U^H a
  double* G2 = take(l.tail);
  double* W2 = take(l.tail);
  double* Bt = take(l.tail);      // transposed tail of Ut
  double* Ut2 = take(l.tail);
  void* work = q;

  // ---- level 0: G = a a^H, eigenvectors -> Ut (row a = u_a, descending eigenvalue)
  if (pdmrg_gemm_nt(dtype, rows, rows, cols, 1, a, cols, 0, a, cols, 0, 1, G, rows, 0, 1.0, 0.0, 0, stream)) return 1;
  double* wdev = Y;               // eigenvalues (k doubles) parked in Y's storage
  if (heevd_inplace(s, dtype, rows, G, wdev, work, l.lwork_elems, info_dev, stream)) return 1;
  // memory row i of G = conj(eigenvector i) (column-major solver on the row-major array), ascending
  if (pdmrg_permute4(dtype, G + (size_t)E * (rows - 1) * rows, Ut, 1, 1, k, rows, 0, 0, -(long long)rows, 1, 1, nullptr, 0, stream)) return 1;
  if (pdmrg_permute4(dtype, a, mT, 1, 1, cols, rows, 0, 0, 1, cols, 0, nullptr, 0, stream)) return 1;   // mT[c,r] = a[r,c]
  std::vector<double> lam(k);
  int info_host = 0;
  PDMRG_CHECK_CUDA(cudaMemcpyAsync(lam.data(), wdev, (size_t)k * 8, cudaMemcpyDeviceToHost, stream));
  PDMRG_CHECK_CUDA(cudaMemcpyAsync(&info_host, info_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PDMRG_CHECK_CUDA(cudaStreamSynchronize(stream));
  // cuSOLVER's divide-and-conquer heevd can fail (info > 0) on exactly rank-deficient Gram matrices
  // (zero-padded tensors right after a chi increase): hand the call to gesvd
  if (info_host != 0) return GRAM_FALLBACK;
  const double lam_top = lam[k - 1];
  std::vector<double> lam_all(k);
  for (int i = 0; i < k; ++i) lam_all[i] = lam[k - 1 - i];
  // ---- refinement of the small-singular-value tail: the Gram matrix of the projected rows
  //      U_tail^H a has norm (tau S_0)^2, so its eigen-decomposition resolves what level 0 cannot
  int t_prev = 0, nblk = k;       // current block = rows [t_prev, k), eigenvalues lam[0..nblk) ascending
  for (int level = 0; level < GRAM_MAX_REFINE; ++level) {
    const double top = lam[nblk - 1];
    if (!(top > 0.0) || top < 1e-28 * lam_top) break;
    int keep = 0;
    while (keep < nblk && lam[nblk - 1 - keep] >= GRAM_TAU2 * top) ++keep;
    const int t = t_prev + keep, nt = k - t;
    if (nt < 2 || t >= nkeep) break;      // every singular vector the caller keeps is already resolved
    // Uc[i,r] = conj(Ut[t+i,r]);  X2 = Uc mT^T = rows t.. of U^H a
    if (pdmrg_permute4(dtype, Ut + (size_t)E * t * rows, Uc, 1, 1, nt, rows, 0, 0, rows, 1, 1, nullptr, 0, stream)) return 1;
    if (pdmrg_gemm_nt(dtype, nt, cols, rows, 1, Uc, rows, 0, mT, rows, 0, 0, X2, cols, 0, 1.0, 0.0, 0, stream)) return 1;
    if (pdmrg_gemm_nt(dtype, nt, nt, cols, 1, X2, cols, 0, X2, cols, 0, 1, G2, nt, 0, 1.0, 0.0, 0, stream)) return 1;
    // diag(G2)_i = |u_{t+i}^H a|^2 is an ACCURATE estimate of lambda_{t+i} (the level-0 eigenvalue carries noise
    // eps*lambda_0).  Error model of the level-0 basis: direction i is off by min(S_i, eps S_0^2 / S_i); the second
    // eigen-decomposition is skipped when the kept tail vectors [t, nkeep) cannot be wrong by more than 1e-9 S_0
    // in total (product-like states with an empty tail, slowly decaying spectra).
    {
      std::vector<double> dg(nt);
      PDMRG_CHECK_CUDA(cudaMemcpy2DAsync(dg.data(), 8, G2, (size_t)(nt + 1) * eb, 8, nt, cudaMemcpyDeviceToHost, stream));
      PDMRG_CHECK_CUDA(cudaStreamSynchronize(stream));
      const double eps_l0 = 2.2e-16 * lam_top;
      double err2 = 0.0;
      for (int i = 0; i < nkeep - t; ++i) {
        const double li = dg[i] > 0.0 ? dg[i] : 0.0;
        const double alt = li > 0.0 ? eps_l0 * eps_l0 / li : 0.0;
        err2 += li < alt ? li : alt;
      }
      if (err2 <= 1e-18 * lam_top) break;
    }
    if (heevd_inplace(s, dtype, nt, G2, wdev, work, l.lwork_elems, info_dev, stream)) return 1;
    PDMRG_CHECK_CUDA(cudaMemcpyAsync(&info_host, info_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PDMRG_CHECK_CUDA(cudaMemcpyAsync(lam.data(), wdev, (size_t)nt * 8, cudaMemcpyDeviceToHost, stream));
    PDMRG_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (info_host != 0) return GRAM_FALLBACK;
    // W2[j,i] = w_j[i] (descending j);  Bt[r,i] = Ut[t+i,r];  Ut2[j,r] = sum_i W2[j,i] Bt[r,i]
    if (pdmrg_permute4(dtype, G2 + (size_t)E * (nt - 1) * nt, W2, 1, 1, nt, nt, 0, 0, -(long long)nt, 1, 1, nullptr, 0, stream)) return 1;
    if (pdmrg_permute4(dtype, Ut + (size_t)E * t * rows, Bt, 1, 1, rows, nt, 0, 0, 1, rows, 0, nullptr, 0, stream)) return 1;
    if (pdmrg_gemm_nt(dtype, nt, rows, nt, 1, W2, nt, 0, Bt, nt, 0, 0, Ut2, rows, 0, 1.0, 0.0, 0, stream)) return 1;
    PDMRG_CHECK_CUDA(cudaMemcpyAsync(Ut + (size_t)E * t * rows, Ut2, (size_t)nt * rows * eb, cudaMemcpyDeviceToDevice, stream));
    for (int j = 0; j < nt; ++j) lam_all[t + j] = lam[nt - 1 - j];
    t_prev = t; nblk = nt;
  }
  if (basis_only) {
    PDMRG_CHECK_CUDA(cudaMemcpyAsync(U, Ut, (size_t)k * rows * eb, cudaMemcpyDeviceToDevice, stream));
    PDMRG_CHECK_CUDA(cudaMemcpyAsync(S, lam_all.data(), (size_t)k * 8, cudaMemcpyHostToDevice, stream));
    PDMRG_CHECK_CUDA(cudaStreamSynchronize(stream));      // lam_all is a stack-owned host buffer
    return 0;
  }
  // ---- Y = a^H U (column-major cols x k == row-major Yrow[a,c] = sum_r Ut[a,r] conj(mT[c,r])), Householder QR
  //      (only the nkeep leading columns: the discarded triplets are never formed)
  const int nq = nkeep;
  if (pdmrg_gemm_nt(dtype, nq, cols, rows, 1, Ut, rows, 0, mT, rows, 0, 1, Y, cols, 0, 1.0, 0.0, 0, stream)) return 1;
  cusolverStatus_t st;
  if (cplx) st = s->p_cusolverDnZgeqrf(s->handle, cols, nq, (cuDoubleComplex*)Y, cols, (cuDoubleComplex*)tau, (cuDoubleComplex*)work, l.lwork_elems, info_dev);
  else st = s->p_cusolverDnDgeqrf(s->handle, cols, nq, Y, cols, tau, (double*)work, l.lwork_elems, info_dev);
  if (st != CUSOLVER_STATUS_SUCCESS) { set_error("gram SVD: geqrf failed (%d)", (int)st); return 1; }
  if (nq < k) PDMRG_CHECK_CUDA(cudaMemsetAsync(S + nq, 0, (size_t)(k - nq) * 8, stream));
  if (cplx) qr_diag_kernel<true><<<(nq + 127) / 128, 128, 0, stream>>>(Y, cols, nq, S, phase);
  else qr_diag_kernel<false><<<(nq + 127) / 128, 128, 0, stream>>>(Y, cols, nq, S, phase);
  PDMRG_CHECK_LAUNCH();
  if (cplx) st = s->p_cusolverDnZungqr(s->handle, cols, nq, nq, (cuDoubleComplex*)Y, cols, (cuDoubleComplex*)tau, (cuDoubleComplex*)work, l.lwork_elems, info_dev);
  else st = s->p_cusolverDnDorgqr(s->handle, cols, nq, nq, Y, cols, tau, (double*)work, l.lwork_elems, info_dev);
  if (st != CUSOLVER_STATUS_SUCCESS) { set_error("gram SVD: ungqr failed (%d)", (int)st); return 1; }
  count_launch(3);
  int grid = (int)(((long long)nq * cols + 255) / 256);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  if (cplx) vh_from_q_kernel<true><<<grid, 256, 0, stream>>>(Y, phase, nq, cols, (double*)Vh);
  else vh_from_q_kernel<false><<<grid, 256, 0, stream>>>(Y, phase, nq, cols, (double*)Vh);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  // U[r, a] = Ut[a, r]
  if (pdmrg_permute4(dtype, Ut, U, 1, 1, rows, k, 0, 0, 1, rows, 0, nullptr, 0, stream)) return 1;
  return 0;
}

template <bool CPLX>
__global__ void conj_inplace_kernel(double* a, long long n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    a[2 * e + 1] = -a[2 * e + 1];
}

}  // namespace
}  // namespace pdmrg

using namespace pdmrg;

extern "C" size_t pdmrg_svd_workspace_bytes(int dtype, int rows, int cols, int method) {
  Solver* s = get_solver();
  if (!s) return 0;
  if (method == 2) {
    const int r = rows <= cols ? rows : cols, c = rows <= cols ? cols : rows;
    GramLayout g;
    if (gram_layout(s, dtype, r, c, &g)) return 0;
    const size_t eb = elem_bytes(dtype);
    // transposed problem (rows > cols): a^T, U', Vh' staging
    const size_t extra = rows > cols ? 3 * align_up((size_t)rows * cols * eb, 256) : 0;
    SvdLayout fb;                                   // gesvd fallback shares the workspace
    if (svd_layout(s, dtype, rows, cols, 0, &fb)) return 0;
    const size_t need = g.total() + extra + 256;
    return need > fb.total() + 256 ? need : fb.total() + 256;
  }
  SvdLayout l;
  if (svd_layout(s, dtype, rows, cols, method, &l)) return 0;
  return l.total() + 256;
}

extern "C" int pdmrg_svd_topk(int dtype, int rows, int cols, int keep, void* a, void* U, double* S, void* Vh, int method,
                              void* workspace, size_t workspace_bytes, int* info_dev, void* stream_);

extern "C" int pdmrg_svd(int dtype, int rows, int cols, void* a, void* U, double* S, void* Vh, int method,
                         void* workspace, size_t workspace_bytes, int* info_dev, void* stream_) {
  return pdmrg_svd_topk(dtype, rows, cols, 0, a, U, S, Vh, method, workspace, workspace_bytes, info_dev, stream_);
}

extern "C" int pdmrg_svd_topk(int dtype, int rows, int cols, int keep, void* a, void* U, double* S, void* Vh, int method,
                              void* workspace, size_t workspace_bytes, int* info_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Solver* s = get_solver();
  if (!s) return 1;
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "pdmrg_svd: bad dtype");
  PDMRG_REQUIRE(rows > 0 && cols > 0, "pdmrg_svd: bad extents");
  PDMRG_REQUIRE(method >= 0 && method <= 2, "pdmrg_svd: method must be 0 (gesvd), 1 (gesvdj) or 2 (gram)");
  if (method == 2) {
    PDMRG_REQUIRE(workspace_bytes >= pdmrg_svd_workspace_bytes(dtype, rows, cols, 2), "pdmrg_svd: workspace too small");
    if (s->p_cusolverDnSetStream(s->handle, stream) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnSetStream failed"); return 1; }
    unsigned char* w8 = static_cast<unsigned char*>(workspace);
    int rc;
    if (rows <= cols) {
      rc = svd_gram_wide(s, dtype, rows, cols, keep, a, U, S, Vh, w8, info_dev, stream);
      if (rc != GRAM_FALLBACK) return rc;
      return pdmrg_svd(dtype, rows, cols, a, U, S, Vh, 0, workspace, workspace_bytes, info_dev, stream_);
    }
    // a^T = U' S Vh'  =>  a = Vh'^T S U'^T
    const size_t blk = align_up((size_t)rows * cols * elem_bytes(dtype), 256);
    void* aT = w8; void* Up = w8 + blk; void* Vhp = w8 + 2 * blk;
    const int k = cols;
    if (pdmrg_permute4(dtype, a, aT, 1, 1, cols, rows, 0, 0, 1, cols, 0, nullptr, 0, stream)) return 1;
    rc = svd_gram_wide(s, dtype, cols, rows, keep, aT, Up, S, Vhp, w8 + 3 * blk, info_dev, stream);
    if (rc == GRAM_FALLBACK) return pdmrg_svd(dtype, rows, cols, a, U, S, Vh, 0, workspace, workspace_bytes, info_dev, stream_);
    if (rc) return 1;
    if (pdmrg_permute4(dtype, Vhp, U, 1, 1, rows, k, 0, 0, 1, rows, 0, nullptr, 0, stream)) return 1;    // U[r,j] = Vh'[j,r]
    if (pdmrg_permute4(dtype, Up, Vh, 1, 1, k, cols, 0, 0, 1, k, 0, nullptr, 0, stream)) return 1;       // Vh[j,c] = U'[c,j]
    return 0;
  }
  SvdLayout l;
  if (svd_layout(s, dtype, rows, cols, method, &l)) return 1;
  PDMRG_REQUIRE(workspace_bytes >= l.total(), "pdmrg_svd: workspace too small");
  if (s->p_cusolverDnSetStream(s->handle, stream) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnSetStream failed"); return 1; }
  const bool cplx = dtype == PDMRG_C128;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  void* at = ws;
  void* uc = ws + l.at;
  void* vc = ws + l.at + l.uc;
  void* work = ws + l.at + l.uc + l.vc;
  const size_t eb = elem_bytes(dtype);
  const int lwork = (int)((l.lwork_bytes - align_up((size_t)l.q * 8 + 8, 256)) / eb);
  double* rwork = reinterpret_cast<double*>(ws + l.total() - align_up((size_t)l.q * 8 + 8, 256));
  const int p = l.p, q = l.q, k = l.k;
  // column-major (p x q) input G: the row-major a itself (cols >= rows) or its explicit transpose
  void* G = a;
  if (l.transpose) {
    // at[c][r] = a[r][c]  (row-major cols x rows)  == column-major rows x cols
    if (pdmrg_permute4(dtype, a, at, 1, 1, cols, rows, 0, 0, 1, cols, 0, nullptr, 0, stream)) return 1;
    G = at;
  }
  cusolverStatus_t st;
  if (method == 0) {
    // G = Uc S VTc ; Uc (p x k, ld p), VTc (k x q, ld k)
    void* Uc = l.transpose ? uc : Vh;   // no transpose: row-major Vh (k x cols) IS Uc
    void* VTc = l.transpose ? vc : U;   // no transpose: row-major U (rows x k) IS VTc
    if (cplx)
      st = s->p_cusolverDnZgesvd(s->handle, 'S', 'S', p, q, (cuDoubleComplex*)G, p, S, (cuDoubleComplex*)Uc, p,
                                 (cuDoubleComplex*)VTc, k, (cuDoubleComplex*)work, lwork, rwork, info_dev);
    else
      st = s->p_cusolverDnDgesvd(s->handle, 'S', 'S', p, q, (double*)G, p, S, (double*)Uc, p, (double*)VTc, k,
                                 (double*)work, lwork, rwork, info_dev);
    if (st != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDn?gesvd failed (%d)", (int)st); return 1; }
    count_launch();
    if (l.transpose) {
      // a = Uc S VTc directly: U[r][j] = Uc(r,j) at r + j*p ;  Vh[j][c] = VTc(j,c) at j + c*k
      if (pdmrg_permute4(dtype, Uc, U, 1, 1, rows, k, 0, 0, 1, p, 0, nullptr, 0, stream)) return 1;
      if (pdmrg_permute4(dtype, VTc, Vh, 1, 1, k, cols, 0, 0, 1, k, 0, nullptr, 0, stream)) return 1;
    }
  } else {
    // G = Uc S Vc^H ; Uc (p x k, ld p), Vc (q x k, ld q)
    void* Uc = l.transpose ? uc : Vh;
    void* Vc = vc;
    if (cplx)
      st = s->p_cusolverDnZgesvdj(s->handle, CUSOLVER_EIG_MODE_VECTOR, 1, p, q, (cuDoubleComplex*)G, p, S,
                                  (cuDoubleComplex*)Uc, p, (cuDoubleComplex*)Vc, q, (cuDoubleComplex*)work, lwork,
                                  info_dev, s->jinfo);
    else
      st = s->p_cusolverDnDgesvdj(s->handle, CUSOLVER_EIG_MODE_VECTOR, 1, p, q, (double*)G, p, S, (double*)Uc, p,
                                  (double*)Vc, q, (double*)work, lwork, info_dev, s->jinfo);
    if (st != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDn?gesvdj failed (%d)", (int)st); return 1; }
    count_launch();
    if (l.transpose) {
      // a = Uc S Vc^H : U[r][j] = Uc(r,j) at r + j*p ; Vh[j][c] = conj(Vc(c,j)) at c + j*q
      if (pdmrg_permute4(dtype, Uc, U, 1, 1, rows, k, 0, 0, 1, p, 0, nullptr, 0, stream)) return 1;
      if (pdmrg_permute4(dtype, Vc, Vh, 1, 1, k, cols, 0, 0, q, 1, 1, nullptr, 0, stream)) return 1;
    } else {
      // a^T = Uc S Vc^H  =>  a = conj(Vc) S Uc^T : U[r][j] = conj(Vc(r,j)) at r + j*q ; Vh = Uc buffer
      if (pdmrg_permute4(dtype, Vc, U, 1, 1, rows, k, 0, 0, 1, q, 1, nullptr, 0, stream)) return 1;
    }
  }
  return 0;
}

extern "C" size_t pdmrg_left_basis_workspace_bytes(int dtype, int n, int c) {
  Solver* s = get_solver();
  if (!s) return 0;
  GramLayout g;
  if (gram_layout(s, dtype, n, c, &g)) return 0;
  return g.total() + 256;
}

extern "C" int pdmrg_left_basis(int dtype, int n, int c, int keep, const void* X, void* Ut, double* lam, void* workspace,
                                size_t workspace_bytes, int* info_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Solver* s = get_solver();
  if (!s) return 1;
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "pdmrg_left_basis: bad dtype");
  PDMRG_REQUIRE(n > 0 && c > 0, "pdmrg_left_basis: bad extents");
  PDMRG_REQUIRE(workspace_bytes >= pdmrg_left_basis_workspace_bytes(dtype, n, c), "pdmrg_left_basis: workspace too small");
  if (s->p_cusolverDnSetStream(s->handle, stream) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnSetStream failed"); return 1; }
  const int rc = svd_gram_wide(s, dtype, n, c, keep, X, Ut, lam, nullptr, static_cast<unsigned char*>(workspace), info_dev,
                               stream, true);
  if (rc == GRAM_FALLBACK) { set_error("pdmrg_left_basis: cuSOLVER heevd did not converge"); return GRAM_FALLBACK; }
  return rc;
}

// ---------------------------------------------------------------------------------------
// Recycled (warm-started) dominant-subspace step for the moving-centre two-site truncation.
//
// In a converging sweep the cut's singular subspaces barely move between two visits of the same bond, so
// the O(n^3) eigen-decomposition of the reduced density matrix can be replaced by ONE step of ordered
// two-sided subspace iteration started from the previous bases (k kept + p spare vectors, ordered by
// weight):   T = W X1^H  ->  Householder QR (column order kept, so the nested dominant subspaces and the
// grading of the columns survive)  ->  P,   C = P X2^T   (the projection of the wave-function on P),
// followed by an exact Rayleigh-Ritz rotation restricted to the WINDOW [k-b, k+p) around the cut, which is
// the only place where the kept/discarded split is decided.  Everything is GEMMs on the DMMA kernel plus
// cuSOLVER geqrf/ungqr and a (b+p)-sized heevd.  The caller alternates X1/X2 = (psi, psi^T) moving right
// and (psi^T, psi) moving left; see truncate.recycled_truncate for the conjugation bookkeeping.
// ---------------------------------------------------------------------------------------
namespace pdmrg {
namespace {

// window sizes up to this use cuSOLVER's Jacobi eigensolver (PDMRG_JACOBI_MAX overrides, for experiments)
int recycle_jacobi_max() {
  static const int v = [] { const char* e = getenv("PDMRG_JACOBI_MAX"); return e ? atoi(e) : 160; }();
  return v;
}
#define RECYCLE_JACOBI_MAX (recycle_jacobi_max())

struct RecycleLayout {
  size_t tau, H, lam, Tw, G, lwork_bytes;
  int lwork_elems;
  size_t total() const { return tau + H + lam + Tw + G + lwork_bytes; }
};

int recycle_layout(Solver* s, int dtype, int a, int c, int K, int w, RecycleLayout* l) {
  const size_t eb = elem_bytes(dtype);
  if (w < 1) w = 1;                      // no Rayleigh-Ritz window (plain ordered QR): keep the queries valid
  l->tau = align_up((size_t)K * eb, 256);
  l->H = align_up((size_t)w * w * eb, 256);
  l->lam = align_up((size_t)w * 8, 256);
  l->Tw = align_up((size_t)w * (size_t)(a > c ? a : c) * eb, 256);
  l->G = align_up((size_t)K * K * eb, 256);
  int lw_e = 0, lw_q = 0, lw_g = 0, lw_p = 0, lw_j = 0;
  cusolverStatus_t st;
  if (dtype == PDMRG_C128) {
    st = s->p_cusolverDnZheevd_bufferSize(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, w, nullptr, w, nullptr, &lw_e);
    if (st == CUSOLVER_STATUS_SUCCESS) st = s->p_cusolverDnZgeqrf_bufferSize(s->handle, a, K, nullptr, a, &lw_q);
    if (st == CUSOLVER_STATUS_SUCCESS) st = s->p_cusolverDnZungqr_bufferSize(s->handle, a, K, K, nullptr, a, nullptr, &lw_g);
    if (st == CUSOLVER_STATUS_SUCCESS) st = s->p_cusolverDnZpotrf_bufferSize(s->handle, CUBLAS_FILL_MODE_LOWER, K, nullptr, K, &lw_p);
    if (st == CUSOLVER_STATUS_SUCCESS && w <= RECYCLE_JACOBI_MAX)
      st = s->p_cusolverDnZheevj_bufferSize(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, w, nullptr, w, nullptr, &lw_j, s->evj);
  } else {
    st = s->p_cusolverDnDsyevd_bufferSize(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, w, nullptr, w, nullptr, &lw_e);
    if (st == CUSOLVER_STATUS_SUCCESS) st = s->p_cusolverDnDgeqrf_bufferSize(s->handle, a, K, nullptr, a, &lw_q);
    if (st == CUSOLVER_STATUS_SUCCESS) st = s->p_cusolverDnDorgqr_bufferSize(s->handle, a, K, K, nullptr, a, nullptr, &lw_g);
    if (st == CUSOLVER_STATUS_SUCCESS) st = s->p_cusolverDnDpotrf_bufferSize(s->handle, CUBLAS_FILL_MODE_LOWER, K, nullptr, K, &lw_p);
    if (st == CUSOLVER_STATUS_SUCCESS && w <= RECYCLE_JACOBI_MAX)
      st = s->p_cusolverDnDsyevj_bufferSize(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, w, nullptr, w, nullptr, &lw_j, s->evj);
  }
  if (st != CUSOLVER_STATUS_SUCCESS) { set_error("cusolver bufferSize (recycle) failed (%d)", (int)st); return 1; }
  int lw = lw_e > lw_q ? lw_e : lw_q;
  if (lw_g > lw) lw = lw_g;
  if (lw_p > lw) lw = lw_p;
  if (lw_j > lw) lw = lw_j;
  l->lwork_elems = lw;
  l->lwork_bytes = align_up((size_t)lw * eb, 256) + 256;
  return 0;
}

// out[r] = || M[r, :] ||_2 for a row-major (rows x len) matrix of E-word elements; one block per row
template <int E>
__global__ void __launch_bounds__(256) row_norm_kernel(const double* __restrict__ M, long long len, double* __restrict__ out) {
  __shared__ double sm[256];
  const double* row = M + (size_t)blockIdx.x * len * E;
  double s = 0.0;
  for (long long i = threadIdx.x; i < len * E; i += 256) { const double v = row[i]; s = fma(v, v, s); }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sqrt(sm[0]);
}

// Row normalisation before the Cholesky-QR passes; one block per row, norms[] = row norms (row_norm_kernel).
// Rows whose norm is below 1e-12 of the largest are numerical noise (the wave-function has no weight there:
// rank-deficient cuts, e.g. product-like states at large chi).  Their computed direction is rounding error
// confined to the few rows of psi that carry weight, so hundreds of them are numerically DEPENDENT and would
// break the Cholesky factorisation; any orthonormal completion is as good as any other for such rows, so
// they are replaced by reproducible pseudo-random vectors (splitmix64 of (row, element)) before normalising.
template <int E>
__global__ void __launch_bounds__(256) row_normalize_kernel(double* __restrict__ M, long long len, const double* __restrict__ norms,
                                                            int nrows) {
  __shared__ double sm[256];
  double mx = 0.0;
  for (int i = threadIdx.x; i < nrows; i += 256) mx = fmax(mx, norms[i]);
  sm[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
    __syncthreads();
  }
  mx = sm[0];
  __syncthreads();
  double* row = M + (size_t)blockIdx.x * len * E;
  const double nrm = norms[blockIdx.x];
  if (nrm > 1e-12 * mx) {
    const double sc = 1.0 / nrm;
    for (long long i = threadIdx.x; i < len * E; i += 256) row[i] *= sc;
    return;
  }
  double s = 0.0;
  for (long long i = threadIdx.x; i < len * E; i += 256) {
    unsigned long long z = ((unsigned long long)blockIdx.x << 32) + (unsigned long long)i + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const double v = (double)(long long)(z >> 11) * (1.0 / 4503599627370496.0) - 1.0;     // uniform in [-1, 1)
    row[i] = v;
    s = fma(v, v, s);
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  const double sc = 1.0 / sqrt(sm[0]);
  for (long long i = threadIdx.x; i < len * E; i += 256) row[i] *= sc;
}

// out[0] += || G - I ||_F^2 over this block's share of the row-major (K x K) matrix G
template <int E>
__global__ void __launch_bounds__(256) gram_defect_kernel(const double* __restrict__ G, int K, double* __restrict__ out) {
  __shared__ double sm[256];
  double s = 0.0;
  const long long tot = (long long)K * K;
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < tot; t += (long long)gridDim.x * 256) {
    const int i = (int)(t / K), j = (int)(t - (long long)i * K);
    double re = G[(size_t)E * t] - (i == j ? 1.0 : 0.0);
    double im = E == 2 ? G[(size_t)E * t + 1] : 0.0;
    s = fma(re, re, s); s = fma(im, im, s);
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, sm[0]);
}

__global__ void recycle_status_kernel(const int* __restrict__ infos, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += infos[i] != 0 ? 1.0 : 0.0;
    out[0] = s;
  }
}

// One Cholesky-QR pass on the rows of T (K x a):  G = T T^H,  G = L L^H,  T <- L^-1 T.
// Row-major G is the column-major conj(G) = Lc Lc^H with L = conj(Lc); the row-major T is the column-major
// T^T, and (L^-1 T)^T = T^T (Lc^H)^-1: one right-sided triangular solve with op = conjugate-transpose.
int cholqr_pass(Solver* s, int dtype, int K, int a, void* T, void* G, void* work, int lwork, int* info_dev, void* stream_) {
  if (pdmrg_gemm_nt(dtype, K, K, a, 1, T, a, 0, T, a, 0, 1, G, K, 0, 1.0, 0.0, 0, stream_)) return 1;
  cusolverStatus_t st;
  if (dtype == PDMRG_C128) st = s->p_cusolverDnZpotrf(s->handle, CUBLAS_FILL_MODE_LOWER, K, (cuDoubleComplex*)G, K, (cuDoubleComplex*)work, lwork, info_dev);
  else st = s->p_cusolverDnDpotrf(s->handle, CUBLAS_FILL_MODE_LOWER, K, (double*)G, K, (double*)work, lwork, info_dev);
  if (st != CUSOLVER_STATUS_SUCCESS) { set_error("recycle: potrf failed (%d)", (int)st); return 1; }
  cublasStatus_t bs;
  if (dtype == PDMRG_C128) {
    const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0);
    bs = s->p_cublasZtrsm(s->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_C, CUBLAS_DIAG_NON_UNIT, a, K, &one,
                          (const cuDoubleComplex*)G, K, (cuDoubleComplex*)T, a);
  } else {
    const double one = 1.0;
    bs = s->p_cublasDtrsm(s->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, a, K, &one,
                          (const double*)G, K, (double*)T, a);
  }
  if (bs != CUBLAS_STATUS_SUCCESS) { set_error("recycle: trsm failed (%d)", (int)bs); return 1; }
  count_launch(2);
  return 0;
}

}  // namespace
}  // namespace pdmrg

extern "C" size_t pdmrg_recycle_basis_workspace_bytes(int dtype, int a, int c, int K, int k, int b) {
  Solver* s = get_solver();
  if (!s) return 0;
  if (K < 1 || k < 1 || k > K || b < 0 || b > k) { set_error("pdmrg_recycle_basis_workspace_bytes: bad extents"); return 0; }
  RecycleLayout l;
  if (recycle_layout(s, dtype, a, c, K, K - (k - b), &l)) return 0;
  return l.total() + 256;
}

extern "C" int pdmrg_recycle_basis(int dtype, int a, int c, int K, int k, int b, int method, const void* X1, const void* X2,
                                   const void* Wt, void* Pt, void* Ct, double* S, void* workspace, size_t workspace_bytes,
                                   int* info_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Solver* s = get_solver();
  if (!s) return 1;
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "pdmrg_recycle_basis: bad dtype");
  PDMRG_REQUIRE(a > 0 && c > 0 && K > 0 && K <= a && K <= c, "pdmrg_recycle_basis: need 0 < K <= min(a, c)");
  PDMRG_REQUIRE(k > 0 && k <= K && b >= 0 && b <= k, "pdmrg_recycle_basis: need 0 < k <= K and 0 <= b <= k");
  const int w0 = k - b, w = K - w0;
  RecycleLayout l;
  if (recycle_layout(s, dtype, a, c, K, w, &l)) return 1;
  PDMRG_REQUIRE(workspace_bytes >= l.total(), "pdmrg_recycle_basis: workspace too small");
  if (s->p_cusolverDnSetStream(s->handle, stream) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnSetStream failed"); return 1; }
  const bool cplx = dtype == PDMRG_C128;
  const size_t eb = elem_bytes(dtype);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  void* tau = ws;
  void* H = ws + l.tau;
  double* lam = reinterpret_cast<double*>(ws + l.tau + l.H);
  void* Tw = ws + l.tau + l.H + l.lam;
  void* G = ws + l.tau + l.H + l.lam + l.Tw;
  void* work = ws + l.tau + l.H + l.lam + l.Tw + l.G;
  PDMRG_REQUIRE(method == 0 || method == 1, "pdmrg_recycle_basis: method must be 0 (Householder) or 1 (Cholesky-QR2)");
  PDMRG_CHECK_CUDA(cudaMemsetAsync(info_dev, 0, 4 * sizeof(int), stream));
  PDMRG_CHECK_CUDA(cudaMemsetAsync(S + K, 0, 2 * sizeof(double), stream));

  // T[i, r] = sum_x W[i, x] conj(X1[r, x])     (K x a), written where P will live
  if (pdmrg_gemm_nt(dtype, K, a, c, 1, Wt, c, 0, X1, c, 0, 1, Pt, a, 0, 1.0, 0.0, 0, stream_)) return 1;
  cusolverStatus_t st;
  if (method == 1) {
    // Cholesky-QR2 on the row-normalised T: triangular like Householder (row i only mixes with rows < i, so the
    // ordering of the basis survives), all GEMM / potrf / trsm.  Valid while the normalised rows are not
    // numerically dependent; the defect of the first pass (S[K]) and the potrf infos (S[K+1]) tell the caller.
    if (s->p_cublasSetStream(s->blas, stream) != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetStream failed"); return 1; }
    if (cplx) row_norm_kernel<2><<<K, 256, 0, stream>>>((const double*)Pt, a, S);
    else row_norm_kernel<1><<<K, 256, 0, stream>>>((const double*)Pt, a, S);
    PDMRG_CHECK_LAUNCH();
    if (cplx) row_normalize_kernel<2><<<K, 256, 0, stream>>>((double*)Pt, a, S, K);
    else row_normalize_kernel<1><<<K, 256, 0, stream>>>((double*)Pt, a, S, K);
    PDMRG_CHECK_LAUNCH();
    if (cholqr_pass(s, dtype, K, a, Pt, G, work, l.lwork_elems, info_dev + 1, stream_)) return 1;
    // second pass; its Gram matrix measures what the first one achieved
    if (pdmrg_gemm_nt(dtype, K, K, a, 1, Pt, a, 0, Pt, a, 0, 1, G, K, 0, 1.0, 0.0, 0, stream_)) return 1;
    if (cplx) gram_defect_kernel<2><<<64, 256, 0, stream>>>((const double*)G, K, S + K);
    else gram_defect_kernel<1><<<64, 256, 0, stream>>>((const double*)G, K, S + K);
    PDMRG_CHECK_LAUNCH();
    if (cplx) st = s->p_cusolverDnZpotrf(s->handle, CUBLAS_FILL_MODE_LOWER, K, (cuDoubleComplex*)G, K, (cuDoubleComplex*)work, l.lwork_elems, info_dev + 2);
    else st = s->p_cusolverDnDpotrf(s->handle, CUBLAS_FILL_MODE_LOWER, K, (double*)G, K, (double*)work, l.lwork_elems, info_dev + 2);
    if (st != CUSOLVER_STATUS_SUCCESS) { set_error("recycle: potrf failed (%d)", (int)st); return 1; }
    cublasStatus_t bs;
    if (cplx) {
      const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0);
      bs = s->p_cublasZtrsm(s->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_C, CUBLAS_DIAG_NON_UNIT, a, K, &one,
                            (const cuDoubleComplex*)G, K, (cuDoubleComplex*)Pt, a);
    } else {
      const double one = 1.0;
      bs = s->p_cublasDtrsm(s->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, a, K, &one,
                            (const double*)G, K, (double*)Pt, a);
    }
    if (bs != CUBLAS_STATUS_SUCCESS) { set_error("recycle: trsm failed (%d)", (int)bs); return 1; }
    count_launch(5);
  } else {
    // Householder QR of the column-major (a x K) view: column i of it is row i of T
    if (cplx) st = s->p_cusolverDnZgeqrf(s->handle, a, K, (cuDoubleComplex*)Pt, a, (cuDoubleComplex*)tau, (cuDoubleComplex*)work, l.lwork_elems, info_dev);
    else st = s->p_cusolverDnDgeqrf(s->handle, a, K, (double*)Pt, a, (double*)tau, (double*)work, l.lwork_elems, info_dev);
    if (st != CUSOLVER_STATUS_SUCCESS) { set_error("recycle: geqrf failed (%d)", (int)st); return 1; }
    if (cplx) st = s->p_cusolverDnZungqr(s->handle, a, K, K, (cuDoubleComplex*)Pt, a, (cuDoubleComplex*)tau, (cuDoubleComplex*)work, l.lwork_elems, info_dev);
    else st = s->p_cusolverDnDorgqr(s->handle, a, K, K, (double*)Pt, a, (double*)tau, (double*)work, l.lwork_elems, info_dev);
    if (st != CUSOLVER_STATUS_SUCCESS) { set_error("recycle: ungqr failed (%d)", (int)st); return 1; }
    count_launch(2);
  }
  // C[i, x] = sum_r P[i, r] X2[x, r]            (K x c)
  if (pdmrg_gemm_nt(dtype, K, c, a, 1, Pt, a, 0, X2, a, 0, 0, Ct, c, 0, 1.0, 0.0, 0, stream_)) return 1;

  if (w > 1) {
    // Rayleigh-Ritz on the window rows [w0, K):  H = -(Cw Cw^H); ascending eigenvalues of -H = descending weights.
    // Row-major H is column-major conj(H), so after heevd memory row q holds conj(v_q) = the coefficients
    // Vh[q, i] with which window rows are recombined:  row'_q = sum_i Vh[q, i] row_i  (same for C and P).
    unsigned char* Cw = static_cast<unsigned char*>(Ct) + (size_t)w0 * c * eb;
    unsigned char* Pw = static_cast<unsigned char*>(Pt) + (size_t)w0 * a * eb;
    if (pdmrg_gemm_nt(dtype, w, w, c, 1, Cw, c, 0, Cw, c, 0, 1, H, w, 0, -1.0, 0.0, 0, stream_)) return 1;
    if (w <= RECYCLE_JACOBI_MAX) {
      if (cplx) st = s->p_cusolverDnZheevj(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, w, (cuDoubleComplex*)H, w, lam,
                                           (cuDoubleComplex*)work, l.lwork_elems, info_dev + 3, s->evj);
      else st = s->p_cusolverDnDsyevj(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, w, (double*)H, w, lam, (double*)work,
                                      l.lwork_elems, info_dev + 3, s->evj);
      if (st != CUSOLVER_STATUS_SUCCESS) { set_error("recycle: heevj failed (%d)", (int)st); return 1; }
      count_launch();
    } else {
      if (heevd_inplace(s, dtype, w, (double*)H, lam, work, l.lwork_elems, info_dev + 3, stream)) return 1;
    }
    if (pdmrg_permute4(dtype, Cw, Tw, 1, 1, c, w, 0, 0, 1, c, 0, nullptr, 0, stream_)) return 1;       // Tw[x, i] = Cw[i, x]
    if (pdmrg_gemm_nt(dtype, w, c, w, 1, H, w, 0, Tw, w, 0, 0, Cw, c, 0, 1.0, 0.0, 0, stream_)) return 1;
    if (pdmrg_permute4(dtype, Pw, Tw, 1, 1, a, w, 0, 0, 1, a, 0, nullptr, 0, stream_)) return 1;
    if (pdmrg_gemm_nt(dtype, w, a, w, 1, H, w, 0, Tw, w, 0, 0, Pw, a, 0, 1.0, 0.0, 0, stream_)) return 1;
  }
  if (cplx) row_norm_kernel<2><<<K, 256, 0, stream>>>((const double*)Ct, c, S);
  else row_norm_kernel<1><<<K, 256, 0, stream>>>((const double*)Ct, c, S);
  PDMRG_CHECK_LAUNCH();
  recycle_status_kernel<<<1, 32, 0, stream>>>(info_dev, 4, S + K + 1);
  PDMRG_CHECK_LAUNCH();
  count_launch(2);
  return 0;
}

extern "C" size_t pdmrg_heevd_workspace_bytes(int dtype, int n) {
  Solver* s = get_solver();
  if (!s) return 0;
  int lwork = 0;
  cusolverStatus_t st = dtype == PDMRG_C128
      ? s->p_cusolverDnZheevd_bufferSize(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, nullptr, n, nullptr, &lwork)
      : s->p_cusolverDnDsyevd_bufferSize(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, nullptr, n, nullptr, &lwork);
  if (st != CUSOLVER_STATUS_SUCCESS) { set_error("cusolver heevd bufferSize failed (%d)", (int)st); return 0; }
  return align_up((size_t)lwork * elem_bytes(dtype), 256) + 256;
}

extern "C" int pdmrg_heevd(int dtype, int n, void* a, double* w, void* workspace, size_t workspace_bytes,
                           int* info_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Solver* s = get_solver();
  if (!s) return 1;
  PDMRG_REQUIRE(n > 0, "pdmrg_heevd: bad extent");
  if (s->p_cusolverDnSetStream(s->handle, stream) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnSetStream failed"); return 1; }
  const int lwork = (int)(workspace_bytes / elem_bytes(dtype));
  cusolverStatus_t st;
  // guard against underflowed diagonal entries (see shift_diag_kernel)
  if (dtype == PDMRG_C128) shift_diag_kernel<true><<<1, 256, 0, stream>>>((double*)a, n, 1e-28);
  else shift_diag_kernel<false><<<1, 256, 0, stream>>>((double*)a, n, 1e-28);
  PDMRG_CHECK_LAUNCH();
  // row-major Hermitian a == column-major conj(a); its eigenvectors are the conjugates.
  if (dtype == PDMRG_C128) {
    st = s->p_cusolverDnZheevd(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, (cuDoubleComplex*)a, n, w,
                               (cuDoubleComplex*)workspace, lwork, info_dev);
    if (st != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnZheevd failed (%d)", (int)st); return 1; }
    const long long tot = (long long)n * n;
    int grid = (int)((tot + 255) / 256); if (grid > num_sms() * 8) grid = num_sms() * 8;
    conj_inplace_kernel<true><<<grid, 256, 0, stream>>>((double*)a, tot);
    PDMRG_CHECK_LAUNCH();
    count_launch(2);
  } else {
    st = s->p_cusolverDnDsyevd(s->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, (double*)a, n, w,
                               (double*)workspace, lwork, info_dev);
    if (st != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnDsyevd failed (%d)", (int)st); return 1; }
    count_launch();
  }
  return 0;
}
