// HBM-bound tensor kernels of the hot path: strided 4-index copy (transposes, conjugation,
// gauge scaling) and the sparse MPO "mix" that applies the small local operator W between
// the two chi^3 GEMM stages.  Both are pure streaming kernels: coalesced along the fastest
// output index, 16-byte accesses for complex128, grid sized in multiples of the SM count.
#include "ops.cuh"

namespace pdmrg {
namespace {

template <bool CPLX>
__global__ void __launch_bounds__(256)
permute4_kernel(const double* __restrict__ in, double* __restrict__ out, long long total,
                int n1, int n2, int n3, long long s0, long long s1, long long s2, long long s3,
                int conj, const double* __restrict__ scale, int scale_axis) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx;
    const int i3 = (int)(r % n3); r /= n3;
    const int i2 = (int)(r % n2); r /= n2;
    const int i1 = (int)(r % n1); r /= n1;
    const int i0 = (int)r;
    const long long src = i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3;
    double sc = 1.0;
    if (scale) sc = scale[scale_axis == 0 ? i0 : scale_axis == 1 ? i1 : scale_axis == 2 ? i2 : i3];
    if (CPLX) {
      double2 v = reinterpret_cast<const double2*>(in)[src];
      if (conj) v.y = -v.y;
      v.x *= sc; v.y *= sc;
      reinterpret_cast<double2*>(out)[idx] = v;
    } else {
      out[idx] = in[src] * sc;
    }
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(256)
mix_kernel(MixOpDev op, const double* __restrict__ in, double* __restrict__ out, int nx, int ny,
           long long six, long long siy, long long sox, long long soy) {
  const long long total = (long long)nx * ny;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx / ny), y = (int)(idx % ny);
    const long long ib = x * six + y * siy, ob = x * sox + y * soy;
    for (int r = 0; r < op.rows; ++r) {
      const int k0 = __ldg(op.rowptr + r), k1 = __ldg(op.rowptr + r + 1);
      if (CPLX) {
        double2 acc = make_double2(0.0, 0.0);
        for (int k = k0; k < k1; ++k) {
          const double2 w = __ldg(reinterpret_cast<const double2*>(op.val) + k);
          const double2 v = __ldg(reinterpret_cast<const double2*>(in) + __ldg(op.in_off + k) + ib);
          cfma(acc, w, v);
        }
        reinterpret_cast<double2*>(out)[__ldg(op.out_off + r) + ob] = acc;
      } else {
        double acc = 0.0;
        for (int k = k0; k < k1; ++k)
          acc = fma(__ldg(op.val + 2 * k), __ldg(in + __ldg(op.in_off + k) + ib), acc);
        out[__ldg(op.out_off + r) + ob] = acc;
      }
    }
  }
}

// The same operator with one output row per blockIdx.y: for small point spaces (chi <= 256) the kernel above has only
// nx*ny threads, each walking all rows through dependent metadata loads (latency bound: 15 us for 3 us of traffic at
// chi = 256); here rows x points threads run, the row's metadata is uniform per block, and the re-reads of `in` by the
// rows that share an input slab hit L2 (the whole intermediate is L2-resident at these sizes).
template <bool CPLX>
__global__ void __launch_bounds__(256)
mix_rows_kernel(MixOpDev op, const double* __restrict__ in, double* __restrict__ out, int nx, int ny,
                long long six, long long siy, long long sox, long long soy) {
  const int r = blockIdx.y;
  const int k0 = __ldg(op.rowptr + r), k1 = __ldg(op.rowptr + r + 1);
  const long long obase = __ldg(op.out_off + r);
  const long long total = (long long)nx * ny;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx / ny), y = (int)(idx % ny);
    const long long ib = x * six + y * siy, ob = x * sox + y * soy;
    if (CPLX) {
      double2 acc = make_double2(0.0, 0.0);
      for (int k = k0; k < k1; ++k) {
        const double2 w = __ldg(reinterpret_cast<const double2*>(op.val) + k);
        const double2 v = __ldg(reinterpret_cast<const double2*>(in) + __ldg(op.in_off + k) + ib);
        cfma(acc, w, v);
      }
      reinterpret_cast<double2*>(out)[obase + ob] = acc;
    } else {
      double acc = 0.0;
      for (int k = k0; k < k1; ++k)
        acc = fma(__ldg(op.val + 2 * k), __ldg(in + __ldg(op.in_off + k) + ib), acc);
      out[obase + ob] = acc;
    }
  }
}

// y[b, e] = sign * sum_{q < nq} Z[b, idx[q], e]  (+ y[b, e] if accumulate); doubles (complex = 2 per element)
struct SlabIdx { int v[32]; };
__global__ void __launch_bounds__(256)
reduce_slabs_kernel(const double* __restrict__ Z, double* __restrict__ y, long long nd, int nb, int nq, SlabIdx idx,
                    long long zstride_b, long long zstride_q, double sign, int accumulate) {
  const long long total = nd * nb;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / nd, e = t - b * nd;
    const double* z = Z + b * zstride_b + e;
    double acc = 0.0;
    for (int q = 0; q < nq; ++q) acc += z[idx.v[q] * zstride_q];
    acc *= sign;
    if (accumulate) acc += y[t];
    y[t] = acc;
  }
}

int grid_for(long long total, int threads) {
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

size_t MixOpHost::device_bytes() const {
  return align_up(rowptr.size() * sizeof(int), 16) + align_up(in_off.size() * sizeof(long long), 16) +
         align_up(val.size() * sizeof(double), 16) + align_up(out_off.size() * sizeof(long long), 16) + 64;
}

int upload_mix_op(const MixOpHost& h, void* dst, MixOpDev* dev, cudaStream_t stream) {
  // pack into one host blob so that a single H2D copy moves the operator
  std::vector<unsigned char> blob(h.device_bytes(), 0);
  size_t o0 = 0;
  size_t o1 = o0 + align_up(h.rowptr.size() * sizeof(int), 16);
  size_t o2 = o1 + align_up(h.in_off.size() * sizeof(long long), 16);
  size_t o3 = o2 + align_up(h.val.size() * sizeof(double), 16);
  memcpy(blob.data() + o0, h.rowptr.data(), h.rowptr.size() * sizeof(int));
  if (!h.in_off.empty()) memcpy(blob.data() + o1, h.in_off.data(), h.in_off.size() * sizeof(long long));
  if (!h.val.empty()) memcpy(blob.data() + o2, h.val.data(), h.val.size() * sizeof(double));
  memcpy(blob.data() + o3, h.out_off.data(), h.out_off.size() * sizeof(long long));
  PDMRG_CHECK_CUDA(cudaMemcpyAsync(dst, blob.data(), blob.size(), cudaMemcpyHostToDevice, stream));
  unsigned char* d = static_cast<unsigned char*>(dst);
  dev->rowptr = reinterpret_cast<const int*>(d + o0);
  dev->in_off = reinterpret_cast<const long long*>(d + o1);
  dev->val = reinterpret_cast<const double*>(d + o2);
  dev->out_off = reinterpret_cast<const long long*>(d + o3);
  dev->rows = h.rows();
  return 0;
}

int launch_reduce_slabs(int dtype, const void* Z, void* y, long long n_elems, int nb, int nq, const int* idx_host,
                        long long zstride_b, long long zstride_q, double sign, int accumulate, cudaStream_t stream) {
  if (n_elems == 0 || nb == 0) return 0;
  PDMRG_REQUIRE(nq <= 32, "reduce_slabs: more than 32 slabs");
  const int E = dtype == PDMRG_C128 ? 2 : 1;
  SlabIdx idx;
  for (int q = 0; q < 32; ++q) idx.v[q] = q < nq ? idx_host[q] : 0;
  const long long nd = n_elems * E;
  const int grid = grid_for(nd * nb, 256);
  reduce_slabs_kernel<<<grid, 256, 0, stream>>>((const double*)Z, (double*)y, nd, nb, nq, idx, zstride_b * E, zstride_q * E, sign, accumulate);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  return 0;
}

int launch_mix(int dtype, const MixOpDev& op, const void* in, void* out, int nx, int ny,
               long long six, long long siy, long long sox, long long soy, cudaStream_t stream) {
  const long long total = (long long)nx * ny;
  if (total == 0 || op.rows == 0) return 0;
  const int grid = grid_for(total, 256);
  if (total <= 256ll * 256ll && op.rows > 1 && op.rows <= 65535) {     // measured: 15.5 -> 10.6 us at chi=256, but 19 -> 30 us at chi=512 (f64)
    const dim3 g2(grid, op.rows);
    if (dtype == PDMRG_C128)
      mix_rows_kernel<true><<<g2, 256, 0, stream>>>(op, (const double*)in, (double*)out, nx, ny, six, siy, sox, soy);
    else
      mix_rows_kernel<false><<<g2, 256, 0, stream>>>(op, (const double*)in, (double*)out, nx, ny, six, siy, sox, soy);
  } else if (dtype == PDMRG_C128)
    mix_kernel<true><<<grid, 256, 0, stream>>>(op, (const double*)in, (double*)out, nx, ny, six, siy, sox, soy);
  else
    mix_kernel<false><<<grid, 256, 0, stream>>>(op, (const double*)in, (double*)out, nx, ny, six, siy, sox, soy);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  return 0;
}

}  // namespace pdmrg

using namespace pdmrg;

extern "C" int pdmrg_permute4(int dtype, const void* in, void* out, int n0, int n1, int n2, int n3,
                              long long s0, long long s1, long long s2, long long s3, int conj,
                              const double* scale, int scale_axis, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "pdmrg_permute4: bad dtype");
  PDMRG_REQUIRE(n0 >= 0 && n1 >= 0 && n2 >= 0 && n3 >= 0, "pdmrg_permute4: negative extent");
  PDMRG_REQUIRE(scale_axis >= 0 && scale_axis < 4, "pdmrg_permute4: bad scale axis");
  const long long total = (long long)n0 * n1 * n2 * n3;
  if (total == 0) return 0;
  const int grid = grid_for(total, 256);
  if (dtype == PDMRG_C128)
    permute4_kernel<true><<<grid, 256, 0, stream>>>((const double*)in, (double*)out, total, n1, n2, n3, s0, s1, s2, s3, conj, scale, scale_axis);
  else
    permute4_kernel<false><<<grid, 256, 0, stream>>>((const double*)in, (double*)out, total, n1, n2, n3, s0, s1, s2, s3, 0, scale, scale_axis);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  return 0;
}

// ---- diagonal of the effective Hamiltonian (for the diagonal Davidson preconditioner) ----------------------------
//   diag[p,k,s] (+)= sign * sum_{j,o} L[k,j,k] * Wpp[p,j,o] * R[s,o,s],   Wpp[p,j,o] = Wc[(j,p),(o,p)]
namespace pdmrg {
template <bool CPLX>
__global__ void __launch_bounds__(256)
heff_diag_kernel(int P, int Dl, int Dr, int wL, int wR, const double* __restrict__ L, const double* __restrict__ R,
                 const double* __restrict__ Wpp, double sign, int accumulate, double* __restrict__ out) {
  constexpr int E = CPLX ? 2 : 1;
  const long long total = (long long)Dl * Dr;
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
    const int k = (int)(t / Dr), s = (int)(t % Dr);
    for (int p = 0; p < P; ++p) {
      double ar = 0.0, ai = 0.0;
      for (int j = 0; j < wL; ++j) {
        const double* lp = L + (size_t)E * (((size_t)k * wL + j) * Dl + k);          // L[k,j,k]
        const double lr = lp[0], li = CPLX ? lp[1] : 0.0;
        double br = 0.0, bi = 0.0;                                                    // sum_o Wpp[p,j,o] R[s,o,s]
        for (int o = 0; o < wR; ++o) {
          const double* wp = Wpp + (size_t)E * (((size_t)p * wL + j) * wR + o);
          const double wr_ = wp[0], wi = CPLX ? wp[1] : 0.0;
          if (wr_ == 0.0 && wi == 0.0) continue;
          const double* rp = R + (size_t)E * (((size_t)s * wR + o) * Dr + s);        // R[s,o,s]
          const double rr = rp[0], ri = CPLX ? rp[1] : 0.0;
          br += wr_ * rr - wi * ri;
          bi += wr_ * ri + wi * rr;
        }
        ar += lr * br - li * bi;
        ai += lr * bi + li * br;
      }
      double* op = out + (size_t)E * (((size_t)p * Dl + k) * Dr + s);
      if (accumulate) { op[0] += sign * ar; if (CPLX) op[1] += sign * ai; }
      else { op[0] = sign * ar; if (CPLX) op[1] = sign * ai; }
    }
  }
}
}  // namespace pdmrg

extern "C" int pdmrg_heff_diag(int dtype, int P, int Dl, int Dr, int wL, int wR, const void* L, const void* R,
                               const void* Wpp_dev, double sign, int accumulate, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "pdmrg_heff_diag: bad dtype");
  PDMRG_REQUIRE(P > 0 && Dl > 0 && Dr > 0 && wL > 0 && wR > 0, "pdmrg_heff_diag: bad extents");
  const int grid = grid_for((long long)Dl * Dr, 256);
  if (dtype == PDMRG_C128)
    heff_diag_kernel<true><<<grid, 256, 0, stream>>>(P, Dl, Dr, wL, wR, (const double*)L, (const double*)R, (const double*)Wpp_dev, sign, accumulate, (double*)out);
  else
    heff_diag_kernel<false><<<grid, 256, 0, stream>>>(P, Dl, Dr, wL, wR, (const double*)L, (const double*)R, (const double*)Wpp_dev, sign, accumulate, (double*)out);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  return 0;
}
