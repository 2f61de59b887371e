// Truncation back-ends (T2: pydmrg/tools/mps_tools.py:106 np.linalg.svd;
// T1: pydmrg/dmrg.py:68 eig of the reduced density matrix) through cuSOLVER, which is
// dlopen'ed on first use so that the rest of the library has no link-time dependency on it.
// cuSOLVER is column-major; the reference's matrices are row-major.  A row-major (r x c)
// array is the column-major (c x r) transpose, so no data is moved when c >= r:
//   a^T = Uc S VTc   =>   a = VTc^T S Uc^T,   row-major U = VTc buffer, row-major Vh = Uc buffer.
#include "common.cuh"

#include <cusolverDn.h>
#include <dlfcn.h>
#include <mutex>

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
#undef PDMRG_SOLVER_FN
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
#undef PDMRG_LOAD
    if (!all) return;
    if (s.p_cusolverDnCreate(&s.handle) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnCreate failed"); return; }
    if (s.p_cusolverDnCreateGesvdjInfo(&s.jinfo) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnCreateGesvdjInfo failed"); return; }
    s.p_cusolverDnXgesvdjSetTolerance(s.jinfo, 1e-15);
    s.p_cusolverDnXgesvdjSetMaxSweeps(s.jinfo, 100);
    s.p_cusolverDnXgesvdjSetSortEig(s.jinfo, 1);
    ok = true;
  });
  return ok ? &s : nullptr;
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
  SvdLayout l;
  if (svd_layout(s, dtype, rows, cols, method, &l)) return 0;
  return l.total() + 256;
}

extern "C" int pdmrg_svd(int dtype, int rows, int cols, void* a, void* U, double* S, void* Vh, int method,
                         void* workspace, size_t workspace_bytes, int* info_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Solver* s = get_solver();
  if (!s) return 1;
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "pdmrg_svd: bad dtype");
  PDMRG_REQUIRE(rows > 0 && cols > 0, "pdmrg_svd: bad extents");
  PDMRG_REQUIRE(method == 0 || method == 1, "pdmrg_svd: method must be 0 (gesvd) or 1 (gesvdj)");
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
