// TEST INFRASTRUCTURE ONLY.  Host stand-ins for everything pydmrg_b200/csrc/davidson.cu calls, so that the C++
// Davidson loop itself (pdmrg_davidson / pdmrg_davidson_precond: cycle sequencing, projected-matrix bookkeeping,
// restarts, lindep pruning, the host Hessenberg-QR, the preconditioner hook and its safeguard) can be compiled with
// g++ and executed on plain host memory in the CPU test-suite (tests/test_native_davidson_host_cpu.py).
// The vector "kernels" follow the semantics documented in include/pydmrg_b200.h; the mat-vec is a callback.
#include <complex>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

typedef std::complex<double> cd;
struct pdmrg_heff_plan;
struct pdmrg_comm;
typedef int (*matvec_cb)(const double* x, double* y, double sign);

namespace pdmrg {
static char g_err[512];
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
}  // namespace pdmrg

static matvec_cb g_cb = nullptr;
static long long g_n = 0;
static long long g_launches = 0;

extern "C" {

const char* shim_last_error() { return pdmrg::g_err; }
void shim_set_matvec(matvec_cb cb, long long n) { g_cb = cb; g_n = n; }
long long shim_launches() { return g_launches; }

// ---- CUDA runtime subset --------------------------------------------------------------------------------------
int cudaMemcpyAsync(void* dst, const void* src, size_t bytes, int, void*) { memmove(dst, src, bytes); return 0; }
int cudaStreamSynchronize(void*) { return 0; }
int cudaHostAlloc(void** p, size_t bytes, unsigned) { *p = malloc(bytes); return *p ? 0 : 2; }
int cudaFreeHost(void* p) { free(p); return 0; }
const char* cudaGetErrorString(int) { return "host shim"; }

// ---- mat-vec ----------------------------------------------------------------------------------------------------
int pdmrg_heff_apply(pdmrg_heff_plan*, const void* const*, const void* const*, const void* x, void* y, double sign, void*,
                     size_t, void*) {
  ++g_launches;
  return g_cb((const double*)x, (double*)y, sign);
}

// the sharded loop's mat-vec (pdmrg_davidson_sharded): one "rank" on the host = the same callback
int pdmrg_heff_apply_fused_ket(pdmrg_heff_plan*, pdmrg_comm*, const void* const*, const void* const*, const void* x, void* y,
                               double sign, void*, size_t, void*) {
  ++g_launches;
  return g_cb((const double*)x, (double*)y, sign);
}

// ---- vector kernels (include/pydmrg_b200.h) ---------------------------------------------------------------------
size_t pdmrg_vec_scratch_bytes(int) { return 1024; }

int pdmrg_vec_dots(int dtype, long long n, int m, const void* X_, long long stride, const void* y_, void* out_, void*, size_t, void*) {
  ++g_launches;
  if (dtype == 1) {
    const cd* X = (const cd*)X_; const cd* y = (const cd*)y_; cd* out = (cd*)out_;
    for (int i = 0; i < m; ++i) { cd s = 0; for (long long e = 0; e < n; ++e) s += std::conj(X[i * stride + e]) * y[e]; out[i] = s; }
  } else {
    const double* X = (const double*)X_; const double* y = (const double*)y_; double* out = (double*)out_;
    for (int i = 0; i < m; ++i) { double s = 0; for (long long e = 0; e < n; ++e) s += X[i * stride + e] * y[e]; out[i] = s; }
  }
  return 0;
}

int pdmrg_vec_combine(int dtype, long long n, int m, const void* c_, const void* X_, long long stride, void* out_, int acc, void*) {
  ++g_launches;
  if (dtype == 1) {
    const cd* c = (const cd*)c_; const cd* X = (const cd*)X_; cd* out = (cd*)out_;
    for (long long e = 0; e < n; ++e) { cd s = acc ? out[e] : cd(0); for (int i = 0; i < m; ++i) s += c[i] * X[i * stride + e]; out[e] = s; }
  } else {
    const double* c = (const double*)c_; const double* X = (const double*)X_; double* out = (double*)out_;
    for (long long e = 0; e < n; ++e) { double s = acc ? out[e] : 0.0; for (int i = 0; i < m; ++i) s += c[i] * X[i * stride + e]; out[e] = s; }
  }
  return 0;
}

int pdmrg_davidson_residual(int dtype, long long n, int m, const void* v_, const void* e_, const void* XS_, const void* AX_,
                            long long stride, void* r_, double* norm2, void*, size_t, void*) {
  ++g_launches;
  double nrm = 0;
  if (dtype == 1) {
    const cd* v = (const cd*)v_; const cd ev = *(const cd*)e_; const cd* XS = (const cd*)XS_; const cd* AX = (const cd*)AX_; cd* r = (cd*)r_;
    for (long long e = 0; e < n; ++e) {
      cd s = 0; for (int i = 0; i < m; ++i) s += v[i] * (AX[i * stride + e] - ev * XS[i * stride + e]);
      r[e] = s; nrm += std::norm(s);
    }
  } else {
    const double* v = (const double*)v_; const double ev = *(const double*)e_; const double* XS = (const double*)XS_;
    const double* AX = (const double*)AX_; double* r = (double*)r_;
    for (long long e = 0; e < n; ++e) {
      double s = 0; for (int i = 0; i < m; ++i) s += v[i] * (AX[i * stride + e] - ev * XS[i * stride + e]);
      r[e] = s; nrm += s * s;
    }
  }
  norm2[0] = nrm;
  return 0;
}

int pdmrg_vec_project(int dtype, long long n, int m, const void* X_, long long stride, const void* c_, const double* scale,
                      void* r_, double* norm2, void*, size_t, void*) {
  ++g_launches;
  const double s = scale ? 1.0 / std::sqrt(scale[0]) : 1.0;
  double nrm = 0;
  if (dtype == 1) {
    const cd* X = (const cd*)X_; const cd* c = (const cd*)c_; cd* r = (cd*)r_;
    for (long long e = 0; e < n; ++e) { cd t = r[e]; for (int i = 0; i < m; ++i) t -= c[i] * X[i * stride + e]; t *= s; r[e] = t; nrm += std::norm(t); }
  } else {
    const double* X = (const double*)X_; const double* c = (const double*)c_; double* r = (double*)r_;
    for (long long e = 0; e < n; ++e) { double t = r[e]; for (int i = 0; i < m; ++i) t -= c[i] * X[i * stride + e]; t *= s; r[e] = t; nrm += t * t; }
  }
  norm2[0] = nrm;
  return 0;
}

int pdmrg_vec_normalize(int dtype, long long n, const void* in_, const double* norm2, void* out_, void*) {
  ++g_launches;
  const double s = 1.0 / std::sqrt(norm2[0]);
  const long long nd = dtype == 1 ? 2 * n : n;
  const double* in = (const double*)in_; double* out = (double*)out_;
  for (long long e = 0; e < nd; ++e) out[e] = in[e] * s;
  return 0;
}

int pdmrg_vec_precond(int dtype, long long n, void* r_, const void* d_, double th_re, double th_im, double floor_, double* norm2,
                      void*, size_t, void*) {
  ++g_launches;
  double nrm = 0;
  if (dtype == 1) {
    cd* r = (cd*)r_; const cd* d = (const cd*)d_; const cd th(th_re, th_im);
    for (long long e = 0; e < n; ++e) {
      cd den = d[e] - th; const double a = std::abs(den);
      if (a < floor_) den = a > 0 ? den * (floor_ / a) : cd(floor_, 0);
      r[e] = r[e] / den; nrm += std::norm(r[e]);
    }
  } else {
    double* r = (double*)r_; const double* d = (const double*)d_;
    for (long long e = 0; e < n; ++e) {
      double den = d[e] - th_re; if (std::fabs(den) < floor_) den = den < 0 ? -floor_ : floor_;
      r[e] = r[e] / den; nrm += r[e] * r[e];
    }
  }
  norm2[0] = nrm;
  return 0;
}

}  // extern "C"
