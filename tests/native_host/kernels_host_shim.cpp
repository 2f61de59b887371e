// TEST INFRASTRUCTURE ONLY.  Host stand-ins for what the ORCHESTRATION files of the hot path call -- pdmrg_permute4,
// pdmrg_gemm_nt / _batched2, the sparse W mix and its operator upload, the ordered slab reduction, a handful of CUDA
// runtime calls -- so that pydmrg_b200/csrc/heff.cu (mat-vec plans, block updates) and csrc/env_cols.cu (column-
// restricted block updates) themselves, i.e. their offsets, extents, leading dimensions, batch strides and sparsity
// lists, are compiled with g++ and executed on host memory in the CPU test-suite (tests/test_env_cols_model_cpu.py,
// tests/test_heff_host_cpu.py).  Semantics as documented in include/pydmrg_b200.h and csrc/ops.cuh; the CUDA kernels
// themselves are covered by the ``-m gpu`` tests.  The peer-memory (fused) paths are not served here.
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
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
// y[b, :] = sign * sum_q Z[b, idx[q], :] (+ y[b, :])
int launch_reduce_slabs(int dtype, const void* Z_, void* y_, long long n_elems, int nb, int nq, const int* idx, long long zsb,
                        long long zsq, double sign, int accumulate, cudaStream_t) {
  const int E = dtype == PDMRG_C128 ? 2 : 1;
  const double* Z = (const double*)Z_; double* y = (double*)y_;
  for (int b = 0; b < nb; ++b)
    for (long long e = 0; e < n_elems * E; ++e) {
      double s = 0;
      for (int q = 0; q < nq; ++q) s += Z[(b * zsb + idx[q] * zsq) * E + e];
      y[b * n_elems * E + e] = sign * s + (accumulate ? y[b * n_elems * E + e] : 0.0);
    }
  return 0;
}

// peer-memory communicator: not available on the host
int comm_signal(pdmrg_comm*, int, cudaStream_t) { set_error("host shim: no peer memory"); return 1; }
int comm_reduce_gather(pdmrg_comm*, const unsigned*, int, int, int, int, int, int, double, cudaStream_t) { set_error("host shim: no peer memory"); return 1; }
int comm_wait_copy(pdmrg_comm*, void*, long long, cudaStream_t) { set_error("host shim: no peer memory"); return 1; }
void comm_next_epoch(pdmrg_comm*) {}
int comm_rank(const pdmrg_comm*) { return 0; }
int comm_world(const pdmrg_comm*) { return 1; }
double* comm_zpeer(const pdmrg_comm*, int) { return nullptr; }
}  // namespace pdmrg

static long long g_shim_launches = 0;

extern "C" {
const char* shim_last_error() { return pdmrg::g_err; }

// ---- CUDA runtime subset ("device" memory is host memory) ------------------------------------------------------------
cudaError_t cudaMalloc(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "host shim"; }

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

static void gemm_one(int dtype, int M, int N, int K, const void* A_, long long lda, const void* B_, long long ldb, int conjB,
                     void* C_, long long ldc, double are, double aim, int beta) {
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    if (dtype == PDMRG_C128) {
      const cd* A = (const cd*)A_; const cd* B = (const cd*)B_; cd* C = (cd*)C_;
      cd s = 0;
      for (int k = 0; k < K; ++k) s += A[m * lda + k] * (conjB ? std::conj(B[n * ldb + k]) : B[n * ldb + k]);
      C[m * ldc + n] = cd(are, aim) * s + (beta ? C[m * ldc + n] : cd(0));
    } else {
      const double* A = (const double*)A_; const double* B = (const double*)B_; double* C = (double*)C_;
      double s = 0;
      for (int k = 0; k < K; ++k) s += A[m * lda + k] * B[n * ldb + k];
      C[m * ldc + n] = are * s + (beta ? C[m * ldc + n] : 0.0);
    }
  }
}

// two batch levels: offsets b1*stride?1 + idx2[b2]*stride?2 (idx2 = NULL: identity)
int pdmrg_gemm_nt_batched2(int dtype, int M, int N, int K, int nb1, int nb2, const int* idx2, const void* A_, long long lda,
                           long long sA1, long long sA2, const void* B_, long long ldb, long long sB1, long long sB2, int conjB,
                           void* C_, long long ldc, long long sC1, long long sC2, double are, double aim, int beta, void*) {
  ++g_shim_launches;
  const size_t eb = dtype == PDMRG_C128 ? 16 : 8;
  for (int b1 = 0; b1 < nb1; ++b1) for (int q = 0; q < nb2; ++q) {
    const long long b2 = idx2 ? idx2[q] : q;
    gemm_one(dtype, M, N, K, (const char*)A_ + (b1 * sA1 + b2 * sA2) * eb, lda, (const char*)B_ + (b1 * sB1 + b2 * sB2) * eb, ldb,
             conjB, (char*)C_ + (b1 * sC1 + b2 * sC2) * eb, ldc, are, aim, beta);
  }
  return 0;
}

int pdmrg_gemm_nt_scatter(int, int, int, int, int, int, const int*, const void*, long long, long long, long long, const void*,
                          long long, long long, long long, int, const void* const*, int, int, long long, long long, long long,
                          double, double, int, void*) {
  pdmrg::set_error("host shim: no peer memory");
  return 1;
}

long long shim_launches() { return g_shim_launches; }

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
