// TEST INFRASTRUCTURE ONLY.  Host stand-ins for what pydmrg_b200/csrc/env_cols.cu calls (pdmrg_permute4,
// pdmrg_gemm_nt, the sparse W mix and its operator upload), so that the file itself -- the orchestration of the
// column-restricted block updates: offsets, extents, leading dimensions, batch strides -- is compiled with g++ and
// executed on host memory in the CPU test-suite (tests/test_env_cols_model_cpu.py).  Semantics as documented in
// include/pydmrg_b200.h and csrc/ops.cuh; the CUDA kernels themselves are covered by the ``-m gpu`` tests.
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../pydmrg_b200/csrc/ops.cuh"

typedef std::complex<double> cd;

namespace pdmrg {
static char g_err[512];
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}

size_t MixOpHost::device_bytes() const {
  return align_up(rowptr.size() * sizeof(int), 16) + align_up(in_off.size() * sizeof(long long), 16) +
         align_up(val.size() * sizeof(double), 16) + align_up(out_off.size() * sizeof(long long), 16) + 64;
}

int upload_mix_op(const MixOpHost& h, void* dst, MixOpDev* dev, cudaStream_t) {
  unsigned char* d = static_cast<unsigned char*>(dst);
  size_t o0 = 0;
  size_t o1 = o0 + align_up(h.rowptr.size() * sizeof(int), 16);
  size_t o2 = o1 + align_up(h.in_off.size() * sizeof(long long), 16);
  size_t o3 = o2 + align_up(h.val.size() * sizeof(double), 16);
  memcpy(d + o0, h.rowptr.data(), h.rowptr.size() * sizeof(int));
  if (!h.in_off.empty()) memcpy(d + o1, h.in_off.data(), h.in_off.size() * sizeof(long long));
  if (!h.val.empty()) memcpy(d + o2, h.val.data(), h.val.size() * sizeof(double));
  memcpy(d + o3, h.out_off.data(), h.out_off.size() * sizeof(long long));
  dev->rowptr = reinterpret_cast<const int*>(d + o0);
  dev->in_off = reinterpret_cast<const long long*>(d + o1);
  dev->val = reinterpret_cast<const double*>(d + o2);
  dev->out_off = reinterpret_cast<const long long*>(d + o3);
  dev->rows = h.rows();
  return 0;
}

// out[out_off[row] + x*sox + y*soy] = sum_nz val[nz] * in[in_off[nz] + x*six + y*siy]   (element offsets)
int launch_mix(int dtype, const MixOpDev& op, const void* in_, void* out_, int nx, int ny, long long six, long long siy,
               long long sox, long long soy, cudaStream_t) {
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) {
      const long long ib = x * six + y * siy, ob = x * sox + y * soy;
      for (int r = 0; r < op.rows; ++r) {
        if (dtype == PDMRG_C128) {
          cd acc = 0;
          for (int k = op.rowptr[r]; k < op.rowptr[r + 1]; ++k)
            acc += cd(op.val[2 * k], op.val[2 * k + 1]) * ((const cd*)in_)[op.in_off[k] + ib];
          ((cd*)out_)[op.out_off[r] + ob] = acc;
        } else {
          double acc = 0;
          for (int k = op.rowptr[r]; k < op.rowptr[r + 1]; ++k) acc += op.val[2 * k] * ((const double*)in_)[op.in_off[k] + ib];
          ((double*)out_)[op.out_off[r] + ob] = acc;
        }
      }
    }
  return 0;
}
}  // namespace pdmrg

extern "C" {
const char* shim_last_error() { return pdmrg::g_err; }

// out[contiguous (n0,n1,n2,n3)] = in[i0*s0 + i1*s1 + i2*s2 + i3*s3]  (conjugated; scaled along scale_axis)
int pdmrg_permute4(int dtype, const void* in_, void* out_, int n0, int n1, int n2, int n3, long long s0, long long s1,
                   long long s2, long long s3, int conj, const double* scale, int scale_axis, void*) {
  long long o = 0;
  for (int i0 = 0; i0 < n0; ++i0) for (int i1 = 0; i1 < n1; ++i1) for (int i2 = 0; i2 < n2; ++i2) for (int i3 = 0; i3 < n3; ++i3, ++o) {
    const long long src = i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3;
    const int idx[4] = {i0, i1, i2, i3};
    const double sc = scale ? scale[idx[scale_axis]] : 1.0;
    if (dtype == PDMRG_C128) { cd v = ((const cd*)in_)[src]; ((cd*)out_)[o] = (conj ? std::conj(v) : v) * sc; }
    else ((double*)out_)[o] = ((const double*)in_)[src] * sc;
  }
  return 0;
}

// C[b][m,n] = alpha * sum_k A[b][m,k] * op(B[b][n,k]) + beta * C[b][m,n]   (element strides)
int pdmrg_gemm_nt(int dtype, int M, int N, int K, int batch, const void* A_, long long lda, long long sA, const void* B_,
                  long long ldb, long long sB, int conjB, void* C_, long long ldc, long long sC, double are, double aim, int beta,
                  void*) {
  for (int b = 0; b < batch; ++b) for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    if (dtype == PDMRG_C128) {
      const cd* A = (const cd*)A_ + b * sA; const cd* B = (const cd*)B_ + b * sB; cd* C = (cd*)C_ + b * sC;
      cd s = 0;
      for (int k = 0; k < K; ++k) s += A[m * lda + k] * (conjB ? std::conj(B[n * ldb + k]) : B[n * ldb + k]);
      C[m * ldc + n] = cd(are, aim) * s + (beta ? C[m * ldc + n] : cd(0));
    } else {
      const double* A = (const double*)A_ + b * sA; const double* B = (const double*)B_ + b * sB; double* C = (double*)C_ + b * sC;
      double s = 0;
      for (int k = 0; k < K; ++k) s += A[m * lda + k] * B[n * ldb + k];
      C[m * ldc + n] = are * s + (beta ? C[m * ldc + n] : 0.0);
    }
  }
  return 0;
}
}  // extern "C"
