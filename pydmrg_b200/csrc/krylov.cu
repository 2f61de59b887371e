// HBM-bound vector kernels of the device-resident Davidson solver
// (reference: pydmrg/tools/la_tools.py:449-457 projected-matrix dots, :469-470 Ritz
// combination, :479-482 residual and norm, :524-536 projection against the basis).
//
// The Krylov basis and its images stay resident in HBM as two (max_space x n) slabs; every
// kernel below makes ONE pass over the vectors it touches (all m dot products of a vector
// against the basis share a single read of that vector; the residual, its norm and the
// Ritz combination are fused).  Reductions are two-stage and ordered (per-block partials,
// then a single-block tree), so results are bit-reproducible run to run.
#include "common.cuh"

namespace pdmrg {
namespace {

constexpr int VT = 256;        // threads per block
constexpr int MAXM = 16;       // basis vectors handled per pass
constexpr int MAX_BLOCKS = 148 * 4;
constexpr int MAX_DOT_PARTIALS = 4736;   // slice streams of the dots kernel (148 SMs x 8 blocks x up to 4 slices)

struct Coefs {
  double re[2 * MAXM + 2];
  double im[2 * MAXM + 2];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-reduce `nv` doubles per thread (in vals[]) into partial[blockIdx.x * nv + i]
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&vals)[NV], int nv, double* partial) {
  __shared__ double sm[VT / 32][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nv) {
      const double s = warp_sum(vals[i]);
      if (lane == 0) sm[warp][i] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < nv) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < VT / 32; ++w) s += sm[w][threadIdx.x];
    partial[(size_t)blockIdx.x * nv + threadIdx.x] = s;
  }
}

// Fused second stage of the ordered two-stage reductions: every block publishes its partials, takes a
// ticket, and the block that draws the last ticket sums ALL partials in a fixed order (warp w reduces values
// w, w+8, ...; every lane adds its rows in order, four independent chains, then a fixed shuffle tree) -- the
// result does not depend on which block happens to be last, so it is bit-reproducible, and the separate
// one-block kernel launch is gone (the Davidson loop at small bond dimension is launch-latency bound).
// `ticket` lives in the last 256 bytes of the caller's scratch buffer: zero before first use, left at zero.
__device__ __forceinline__ void last_block_reduce(const double* partial, int nrows, int nv, double* __restrict__ out,
                                                  unsigned int* ticket) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < nv; i += VT / 32) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int b = lane;
    for (; b + 96 < nrows; b += 128) {
      s0 += __ldcg(partial + (size_t)b * nv + i);
      s1 += __ldcg(partial + (size_t)(b + 32) * nv + i);
      s2 += __ldcg(partial + (size_t)(b + 64) * nv + i);
      s3 += __ldcg(partial + (size_t)(b + 96) * nv + i);
    }
    for (; b < nrows; b += 32) s0 += __ldcg(partial + (size_t)b * nv + i);
    const double s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) out[i] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// out-of-register formulation: the 8 warps of a block split into P = 2^k <= min(m, 8) vector groups and
// S = 8 / P element slices.  Warp (g, s) walks the chunks of slice stream (block, s) and accumulates
// <X_i|y> for i = g, g + P (< m <= 16): two complex accumulators per lane instead of 2*m, so eight blocks
// stay resident per SM and every lane keeps four independent 16-byte loads of X in flight; the P warps of a
// slice read the same y chunk, which is served by L1 after the first.  partial[(block*S + s)*nv + comp].
template <bool CPLX>
__global__ void __launch_bounds__(VT)
dots_kernel(long long n, int m, int P, const double* __restrict__ X, long long stride, const double* __restrict__ y,
            double* partial, double* __restrict__ out, unsigned int* ticket) {
  constexpr int U = 4;                                   // elements per lane per chunk
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = (VT / 32) / P;
  const int g = warp % P, sl = warp / P;
  const int i0 = g, i1 = g + P;                          // i1 may be >= m (then unused)
  const bool two = i1 < m;
  double a0r = 0.0, a0i = 0.0, a1r = 0.0, a1i = 0.0;
  const long long chunk = 32 * U;
  const long long nchunks = (n + chunk - 1) / chunk;
  for (long long q = (long long)blockIdx.x * S + sl; q < nchunks; q += (long long)gridDim.x * S) {
    const long long base = q * chunk + lane;
    if (CPLX) {
      double2 yv[U], x0[U], x1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = base + 32 * u;
        const bool ok = e < n;
        yv[u] = ok ? reinterpret_cast<const double2*>(y)[e] : make_double2(0.0, 0.0);
        x0[u] = ok ? reinterpret_cast<const double2*>(X + 2 * i0 * stride)[e] : make_double2(0.0, 0.0);
        x1[u] = (ok && two) ? reinterpret_cast<const double2*>(X + 2 * i1 * stride)[e] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a0r = fma(x0[u].x, yv[u].x, a0r); a0r = fma(x0[u].y, yv[u].y, a0r);
        a0i = fma(x0[u].x, yv[u].y, a0i); a0i = fma(-x0[u].y, yv[u].x, a0i);
        a1r = fma(x1[u].x, yv[u].x, a1r); a1r = fma(x1[u].y, yv[u].y, a1r);
        a1i = fma(x1[u].x, yv[u].y, a1i); a1i = fma(-x1[u].y, yv[u].x, a1i);
      }
    } else {
      double yv[U], x0[U], x1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = base + 32 * u;
        const bool ok = e < n;
        yv[u] = ok ? y[e] : 0.0;
        x0[u] = ok ? X[i0 * stride + e] : 0.0;
        x1[u] = (ok && two) ? X[i1 * stride + e] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { a0r = fma(x0[u], yv[u], a0r); a1r = fma(x1[u], yv[u], a1r); }
    }
  }
  a0r = warp_sum(a0r); a0i = warp_sum(a0i); a1r = warp_sum(a1r); a1i = warp_sum(a1i);
  if (lane == 0) {
    constexpr int E = CPLX ? 2 : 1;
    double* dst = partial + ((size_t)blockIdx.x * S + sl) * (E * m);
    dst[E * i0] = a0r;
    if (CPLX) dst[E * i0 + 1] = a0i;
    if (two) { dst[E * i1] = a1r; if (CPLX) dst[E * i1 + 1] = a1i; }
  }
  last_block_reduce(partial, gridDim.x * S, (CPLX ? 2 : 1) * m, out, ticket);
}

template <bool CPLX>
__global__ void __launch_bounds__(VT)
combine_kernel(long long n, int m, Coefs c, const double* __restrict__ X, long long stride,
               double* __restrict__ out, int accumulate) {
  for (long long e = blockIdx.x * (long long)VT + threadIdx.x; e < n; e += (long long)gridDim.x * VT) {
    if (CPLX) {
      double2 acc = accumulate ? reinterpret_cast<double2*>(out)[e] : make_double2(0.0, 0.0);
      for (int i = 0; i < m; ++i)
        cfma(acc, make_double2(c.re[i], c.im[i]), reinterpret_cast<const double2*>(X + 2 * i * stride)[e]);
      reinterpret_cast<double2*>(out)[e] = acc;
    } else {
      double acc = accumulate ? out[e] : 0.0;
      for (int i = 0; i < m; ++i) acc = fma(c.re[i], X[i * stride + e], acc);
      out[e] = acc;
    }
  }
}

// r = sum_i v_i (AX_i - e XS_i); coefficient slot m holds e
template <bool CPLX>
__global__ void __launch_bounds__(VT)
residual_kernel(long long n, int m, Coefs c, const double* __restrict__ XS, const double* __restrict__ AX,
                long long stride, double* __restrict__ r, int accumulate, double* partial, double* __restrict__ norm2_out,
                unsigned int* ticket) {
  double nrm[1] = {0.0};
  for (long long e = blockIdx.x * (long long)VT + threadIdx.x; e < n; e += (long long)gridDim.x * VT) {
    if (CPLX) {
      double2 ax = make_double2(0.0, 0.0), xs = make_double2(0.0, 0.0);
      for (int i = 0; i < m; ++i) {
        const double2 v = make_double2(c.re[i], c.im[i]);
        cfma(ax, v, reinterpret_cast<const double2*>(AX + 2 * i * stride)[e]);
        cfma(xs, v, reinterpret_cast<const double2*>(XS + 2 * i * stride)[e]);
      }
      const double2 ev = make_double2(c.re[2 * MAXM], c.im[2 * MAXM]);
      double2 res = accumulate ? reinterpret_cast<double2*>(r)[e] : make_double2(0.0, 0.0);
      res.x += ax.x - (ev.x * xs.x - ev.y * xs.y);
      res.y += ax.y - (ev.x * xs.y + ev.y * xs.x);
      reinterpret_cast<double2*>(r)[e] = res;
      nrm[0] = fma(res.x, res.x, nrm[0]); nrm[0] = fma(res.y, res.y, nrm[0]);
    } else {
      double ax = 0.0, xs = 0.0;
      for (int i = 0; i < m; ++i) { ax = fma(c.re[i], AX[i * stride + e], ax); xs = fma(c.re[i], XS[i * stride + e], xs); }
      double res = (accumulate ? r[e] : 0.0) + ax - c.re[2 * MAXM] * xs;
      r[e] = res;
      nrm[0] = fma(res, res, nrm[0]);
    }
  }
  block_reduce_store<1>(nrm, 1, partial);
  last_block_reduce(partial, gridDim.x, 1, norm2_out, ticket);
}

// r = (r - sum_i c_i X_i) * s,  s = (1.0 / sqrt(scale[0])) if scale given on the LAST pass
template <bool CPLX>
__global__ void __launch_bounds__(VT)
project_kernel(long long n, int m, const double* __restrict__ X, long long stride, const double* __restrict__ cdev,
               const double* __restrict__ scale, double* __restrict__ r, double* partial, double* __restrict__ norm2_out,
               unsigned int* ticket) {
  const double s = scale ? (1.0 / sqrt(scale[0])) : 1.0;
  __shared__ double csh[2 * MAXM];                 // the coefficients come from a device-side dots result
  if ((int)threadIdx.x < (CPLX ? 2 : 1) * m) csh[threadIdx.x] = -cdev[threadIdx.x];
  __syncthreads();
  double nrm[1] = {0.0};
  for (long long e = blockIdx.x * (long long)VT + threadIdx.x; e < n; e += (long long)gridDim.x * VT) {
    if (CPLX) {
      double2 res = reinterpret_cast<double2*>(r)[e];
#pragma unroll 8
      for (int i = 0; i < m; ++i) {
        const double2 ci = make_double2(csh[2 * i], csh[2 * i + 1]);
        cfma(res, ci, reinterpret_cast<const double2*>(X + 2 * i * stride)[e]);
      }
      res.x *= s; res.y *= s;
      reinterpret_cast<double2*>(r)[e] = res;
      nrm[0] = fma(res.x, res.x, nrm[0]); nrm[0] = fma(res.y, res.y, nrm[0]);
    } else {
      double res = r[e];
#pragma unroll 8
      for (int i = 0; i < m; ++i) res = fma(csh[i], X[i * stride + e], res);
      res *= s;
      r[e] = res;
      nrm[0] = fma(res, res, nrm[0]);
    }
  }
  block_reduce_store<1>(nrm, 1, partial);
  last_block_reduce(partial, gridDim.x, 1, norm2_out, ticket);
}

// t = r / (d - theta), |d - theta| kept >= floor (direction of d - theta preserved); norm2_out = ||t||^2.
// The diagonal (Jacobi-Davidson style) preconditioner of the Davidson correction step (the reference passes the
// identity, tools/diag_tools.py:137-139; la_tools.py:510 is where a preconditioner acts).
template <bool CPLX>
__global__ void __launch_bounds__(VT)
precond_kernel(long long n, double* __restrict__ r, const double* __restrict__ diag, double th_re, double th_im, double floor_,
               double* partial, double* __restrict__ norm2_out, unsigned int* ticket) {
  double nrm[1] = {0.0};
  for (long long e = blockIdx.x * (long long)VT + threadIdx.x; e < n; e += (long long)gridDim.x * VT) {
    if (CPLX) {
      const double2 d = reinterpret_cast<const double2*>(diag)[e];
      double dr = d.x - th_re, di = d.y - th_im;
      const double a = sqrt(dr * dr + di * di);
      if (a < floor_) {
        if (a > 0.0) { dr *= floor_ / a; di *= floor_ / a; } else { dr = floor_; di = 0.0; }
      }
      const double inv = 1.0 / (dr * dr + di * di);
      const double2 v = reinterpret_cast<double2*>(r)[e];
      double2 t;                                                    // v * conj(den) / |den|^2
      t.x = (v.x * dr + v.y * di) * inv;
      t.y = (v.y * dr - v.x * di) * inv;
      reinterpret_cast<double2*>(r)[e] = t;
      nrm[0] = fma(t.x, t.x, nrm[0]); nrm[0] = fma(t.y, t.y, nrm[0]);
    } else {
      double den = diag[e] - th_re;
      if (fabs(den) < floor_) den = den < 0.0 ? -floor_ : floor_;
      const double t = r[e] / den;
      r[e] = t;
      nrm[0] = fma(t, t, nrm[0]);
    }
  }
  block_reduce_store<1>(nrm, 1, partial);
  last_block_reduce(partial, gridDim.x, 1, norm2_out, ticket);
}

template <bool CPLX>
__global__ void __launch_bounds__(VT)
normalize_kernel(long long n, const double* __restrict__ in, const double* __restrict__ norm2, double* __restrict__ out) {
  const double s = (1.0 / sqrt(norm2[0]));
  const long long nd = CPLX ? 2 * n : n;
  for (long long e = blockIdx.x * (long long)VT + threadIdx.x; e < nd; e += (long long)gridDim.x * VT) out[e] = in[e] * s;
}

int vec_grid(long long n) {
  long long b = (n + VT - 1) / VT;
  const long long cap = (long long)num_sms() * 4 < MAX_BLOCKS ? (long long)num_sms() * 4 : MAX_BLOCKS;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace
}  // namespace pdmrg

using namespace pdmrg;

extern "C" size_t pdmrg_vec_scratch_bytes(int m) {
  (void)m;
  return (size_t)MAX_DOT_PARTIALS * 2 * MAXM * sizeof(double) + 256;
}

// the reduction ticket lives in the last 256 bytes of the vector scratch (ZERO before first use)
static unsigned int* ticket_of(void* scratch) {
  return reinterpret_cast<unsigned int*>(static_cast<unsigned char*>(scratch) + (size_t)MAX_DOT_PARTIALS * 2 * MAXM * sizeof(double));
}

extern "C" int pdmrg_vec_dots(int dtype, long long n, int m, const void* X, long long stride, const void* y,
                              void* out_dev, void* scratch, size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(scratch_bytes >= pdmrg_vec_scratch_bytes(m), "pdmrg_vec_dots: scratch too small");
  PDMRG_REQUIRE(n > 0 && m >= 0, "pdmrg_vec_dots: bad extents");
  const bool cplx = dtype == PDMRG_C128;
  const int E = cplx ? 2 : 1;
  double* partial = static_cast<double*>(scratch);
  for (int i0 = 0; i0 < m; i0 += MAXM) {
    const int mm = (m - i0) < MAXM ? (m - i0) : MAXM;
    const double* Xp = static_cast<const double*>(X) + (size_t)E * i0 * stride;
    int P = 1;
    while (2 * P <= mm && 2 * P <= VT / 32) P *= 2;
    const int S = (VT / 32) / P;
    const long long nchunks = (n + 127) / 128;
    long long want = (nchunks + S - 1) / S;
    long long cap = (long long)num_sms() * 8;
    if (cap * S > MAX_DOT_PARTIALS) cap = MAX_DOT_PARTIALS / S;
    if (want > cap) want = cap;
    const int grid = (int)(want < 1 ? 1 : want);
    // complex: partial holds (re,im) pairs -> out is interleaved complex
    double* outp = static_cast<double*>(out_dev) + E * i0;
    if (cplx) dots_kernel<true><<<grid, VT, 0, stream>>>(n, mm, P, Xp, stride, (const double*)y, partial, outp, ticket_of(scratch));
    else dots_kernel<false><<<grid, VT, 0, stream>>>(n, mm, P, Xp, stride, (const double*)y, partial, outp, ticket_of(scratch));
    PDMRG_CHECK_LAUNCH();
    count_launch();
  }
  return 0;
}

static int fill_coefs(Coefs& c, int dtype, const void* host, int m0, int m) {
  const double* h = static_cast<const double*>(host);
  for (int i = 0; i < m; ++i) {
    if (dtype == PDMRG_C128) { c.re[i] = h[2 * (m0 + i)]; c.im[i] = h[2 * (m0 + i) + 1]; }
    else { c.re[i] = h[m0 + i]; c.im[i] = 0.0; }
  }
  return 0;
}

extern "C" int pdmrg_vec_combine(int dtype, long long n, int m, const void* coef_host, const void* X,
                                 long long stride, void* out, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(n > 0 && m > 0, "pdmrg_vec_combine: bad extents");
  const bool cplx = dtype == PDMRG_C128;
  const int E = cplx ? 2 : 1;
  const int grid = vec_grid(n);
  for (int i0 = 0; i0 < m; i0 += 2 * MAXM) {
    const int mm = (m - i0) < 2 * MAXM ? (m - i0) : 2 * MAXM;
    Coefs c;
    fill_coefs(c, dtype, coef_host, i0, mm);
    const double* Xp = static_cast<const double*>(X) + (size_t)E * i0 * stride;
    const int acc = (accumulate || i0 > 0) ? 1 : 0;
    if (cplx) combine_kernel<true><<<grid, VT, 0, stream>>>(n, mm, c, Xp, stride, (double*)out, acc);
    else combine_kernel<false><<<grid, VT, 0, stream>>>(n, mm, c, Xp, stride, (double*)out, acc);
    PDMRG_CHECK_LAUNCH();
    count_launch();
  }
  return 0;
}

extern "C" int pdmrg_davidson_residual(int dtype, long long n, int m, const void* v_host, const void* e_host,
                                       const void* XS, const void* AX, long long stride, void* r,
                                       double* norm2_dev, void* scratch, size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(scratch_bytes >= pdmrg_vec_scratch_bytes(m), "pdmrg_davidson_residual: scratch too small");
  PDMRG_REQUIRE(n > 0 && m > 0, "pdmrg_davidson_residual: bad extents");
  const bool cplx = dtype == PDMRG_C128;
  const int E = cplx ? 2 : 1;
  const int grid = vec_grid(n);
  double* partial = static_cast<double*>(scratch);
  const double* eh = static_cast<const double*>(e_host);
  for (int i0 = 0; i0 < m; i0 += 2 * MAXM) {
    const int mm = (m - i0) < 2 * MAXM ? (m - i0) : 2 * MAXM;
    Coefs c;
    fill_coefs(c, dtype, v_host, i0, mm);
    c.re[2 * MAXM] = eh[0];
    c.im[2 * MAXM] = cplx ? eh[1] : 0.0;
    const double* XSp = static_cast<const double*>(XS) + (size_t)E * i0 * stride;
    const double* AXp = static_cast<const double*>(AX) + (size_t)E * i0 * stride;
    if (cplx) residual_kernel<true><<<grid, VT, 0, stream>>>(n, mm, c, XSp, AXp, stride, (double*)r, i0 > 0, partial, norm2_dev, ticket_of(scratch));
    else residual_kernel<false><<<grid, VT, 0, stream>>>(n, mm, c, XSp, AXp, stride, (double*)r, i0 > 0, partial, norm2_dev, ticket_of(scratch));
    PDMRG_CHECK_LAUNCH();
    count_launch();
  }
  return 0;
}

extern "C" int pdmrg_vec_project(int dtype, long long n, int m, const void* X, long long stride, const void* c_dev,
                                 const double* scale_dev, void* r, double* norm2_dev, void* scratch,
                                 size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(scratch_bytes >= pdmrg_vec_scratch_bytes(m), "pdmrg_vec_project: scratch too small");
  PDMRG_REQUIRE(n > 0 && m >= 0, "pdmrg_vec_project: bad extents");
  const bool cplx = dtype == PDMRG_C128;
  const int E = cplx ? 2 : 1;
  const int grid = vec_grid(n);
  double* partial = static_cast<double*>(scratch);
  // the kernel stages at most MAXM coefficients in shared memory: larger subspaces (nroots >= 2: up to 12 + 4*nroots
  // vectors) go in passes of MAXM; the scaling and the norm that is kept belong to the last pass
  int i0 = 0;
  do {
    const int mm = (m - i0) < MAXM ? (m - i0) : MAXM;
    const bool last = i0 + mm >= m;
    const double* Xp = static_cast<const double*>(X) + (size_t)E * i0 * stride;
    const double* cp = static_cast<const double*>(c_dev) + (size_t)E * i0;
    if (cplx) project_kernel<true><<<grid, VT, 0, stream>>>(n, mm, Xp, stride, cp, last ? scale_dev : nullptr, (double*)r, partial, norm2_dev, ticket_of(scratch));
    else project_kernel<false><<<grid, VT, 0, stream>>>(n, mm, Xp, stride, cp, last ? scale_dev : nullptr, (double*)r, partial, norm2_dev, ticket_of(scratch));
    PDMRG_CHECK_LAUNCH();
    count_launch();
    i0 += mm;
  } while (i0 < m);
  return 0;
}

extern "C" int pdmrg_vec_normalize(int dtype, long long n, const void* in, const double* norm2_dev, void* out,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(n > 0, "pdmrg_vec_normalize: bad extent");
  const int grid = vec_grid(dtype == PDMRG_C128 ? 2 * n : n);
  if (dtype == PDMRG_C128) normalize_kernel<true><<<grid, VT, 0, stream>>>(n, (const double*)in, norm2_dev, (double*)out);
  else normalize_kernel<false><<<grid, VT, 0, stream>>>(n, (const double*)in, norm2_dev, (double*)out);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  return 0;
}

extern "C" int pdmrg_vec_precond(int dtype, long long n, void* r, const void* diag, double theta_re, double theta_im,
                                 double floor_abs, double* norm2_dev, void* scratch, size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(scratch_bytes >= pdmrg_vec_scratch_bytes(1), "pdmrg_vec_precond: scratch too small");
  PDMRG_REQUIRE(n > 0 && floor_abs > 0.0, "pdmrg_vec_precond: bad extent / floor");
  const int grid = vec_grid(n);
  double* partial = static_cast<double*>(scratch);
  if (dtype == PDMRG_C128) precond_kernel<true><<<grid, VT, 0, stream>>>(n, (double*)r, (const double*)diag, theta_re, theta_im, floor_abs, partial, norm2_dev, ticket_of(scratch));
  else precond_kernel<false><<<grid, VT, 0, stream>>>(n, (double*)r, (const double*)diag, theta_re, 0.0, floor_abs, partial, norm2_dev, ticket_of(scratch));
  PDMRG_CHECK_LAUNCH();
  count_launch();
  return 0;
}
