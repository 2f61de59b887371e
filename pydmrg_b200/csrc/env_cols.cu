// Column-restricted environment (block) updates for sharded sweeps.
//
// A rank that owns the ket-index range [k0, k0+kn) of the NEW block computes only
//   right-growing:  out[k, m, q]  for q in the range   (tools/env_tools.py:40-50)
//   left-growing :  out[x, y, a]  for a in the range   (tools/env_tools.py:28-38)
// i.e. 1/n of both chi^3 GEMM stages and of the W mix, and writes its columns IN PLACE into the full-size,
// reference-layout output (every other entry is left untouched).  Ranks owning disjoint ranges therefore
// assemble the whole block with ONE exchange (all-reduce of the zero-initialised block, or an all-gather of
// the column ranges).  Same three stages and the same kernels as pdmrg_env_update_right / _left in heff.cu
// (DMMA NT GEMM, sparse W mix, permute), with the ket extent of the new block replaced by the owned range;
// the final GEMM is batched over the MPO bond of the new block so that it can store straight into the
// (D_bra, w, D_ket) layout at column offset k0.
#include "ops.cuh"

using namespace pdmrg;

namespace {

struct ColsLayout {
  size_t op_bytes, pack_a, pack_b, t1, v;   // byte sizes (256-aligned); identical bound to heff.cu's EnvLayout
  size_t total() const { return op_bytes + pack_a + pack_b + t1 + v; }
};

ColsLayout cols_layout(int dtype, int d, int Dl, int Dr, int wl, int wr) {
  const size_t eb = elem_bytes(dtype);
  ColsLayout l;
  const size_t nnz = (size_t)wl * wr * d * d;
  l.op_bytes = align_up(nnz * (8 + 16) + ((size_t)(wl > wr ? wl : wr) * d + 2) * 16 + 256, 256);
  l.pack_a = align_up((size_t)d * Dl * Dr * eb, 256);
  l.pack_b = l.pack_a;
  const int wmax = wl > wr ? wl : wr;
  l.t1 = align_up((size_t)d * Dl * Dr * wmax * eb, 256);
  l.v = l.t1;
  return l;
}

inline const double* w_at(const double* W, int E, int wb, int d, int a, int b, int i, int j) {
  return W + (size_t)E * ((((size_t)a * wb + b) * d + i) * d + j);
}

}  // namespace

// out[k, m, k0+q] = sum L[j,l,p] Abra[i,j,k] W[l,m,i,n] A[n,p,k0+q],  q < kn   (Abra == NULL: conj(A))
extern "C" int pdmrg_env_update_right_cols(int dtype, int d, int Dl, int Dr, int wl, int wr, const void* W_host,
                                           const void* L, const void* A, const void* Abra, int k0, int kn,
                                           void* out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "env_update_right_cols: bad dtype");
  PDMRG_REQUIRE(k0 >= 0 && kn >= 0 && k0 + kn <= Dr, "env_update_right_cols: ket range outside [0, Dr)");
  if (!W_host) PDMRG_REQUIRE(wl == wr, "env_update_right_cols: identity MPO entry needs wl == wr");
  const ColsLayout lay = cols_layout(dtype, d, Dl, Dr, wl, wr);
  PDMRG_REQUIRE(workspace_bytes >= lay.total(), "env_update_right_cols: workspace too small (pdmrg_env_workspace_bytes)");
  if (kn == 0) return 0;
  const int E = dtype == PDMRG_C128 ? 2 : 1;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  void* op_dev = ws;
  double* At = reinterpret_cast<double*>(ws + lay.op_bytes);
  double* AbT = reinterpret_cast<double*>(ws + lay.op_bytes + lay.pack_a);
  double* T1 = reinterpret_cast<double*>(ws + lay.op_bytes + lay.pack_a + lay.pack_b);
  double* V = reinterpret_cast<double*>(ws + lay.op_bytes + lay.pack_a + lay.pack_b + lay.t1);
  const double* Acols = static_cast<const double*>(A) + (size_t)E * k0;
  // At[n,q,p] = A[n,p,k0+q]
  if (pdmrg_permute4(dtype, Acols, At, 1, d, kn, Dl, 0, (long long)Dl * Dr, 1, Dr, 0, nullptr, 0, stream)) return 1;
  // AbT[k,i,j] = Abra[i,j,k]  for EVERY bra index k   (default bra = conj(A), tools/env_tools.py:41)
  if (pdmrg_permute4(dtype, Abra ? Abra : A, AbT, 1, Dr, d, Dl, 0, 1, (long long)Dl * Dr, Dr, Abra ? 0 : 1, nullptr, 0, stream)) return 1;
  // T1[(n,q),(j,l)] = sum_p At[(n,q),p] L[(j,l),p]
  if (pdmrg_gemm_nt(dtype, d * kn, Dl * wl, Dl, 1, At, Dl, 0, L, Dl, 0, 0, T1, (long long)Dl * wl, 0, 1.0, 0.0, 0, stream)) return 1;
  // V[m,q,i,j] = sum_{l,n} W[l,m,i,n] T1[n,q,j,l]
  MixOpHost h;
  h.rowptr.push_back(0);
  const double* W = static_cast<const double*>(W_host);
  for (int m = 0; m < wr; ++m)
    for (int i = 0; i < d; ++i) {
      for (int l = 0; l < wl; ++l)
        for (int n = 0; n < d; ++n) {
          double re, im = 0.0;
          if (W) { const double* w = w_at(W, E, wr, d, l, m, i, n); re = w[0]; if (E == 2) im = w[1]; }
          else { re = (l == m && i == n) ? 1.0 : 0.0; }
          if (re == 0.0 && im == 0.0) continue;
          h.in_off.push_back((long long)l + (long long)n * kn * Dl * wl);
          h.val.push_back(re); h.val.push_back(im);
        }
      h.rowptr.push_back((int)h.in_off.size());
      h.out_off.push_back((long long)m * kn * d * Dl + (long long)i * Dl);
    }
  PDMRG_REQUIRE(h.device_bytes() <= lay.op_bytes, "env_update_right_cols: operator blob overflow");
  MixOpDev dev;
  if (upload_mix_op(h, op_dev, &dev, stream)) return 1;
  if (launch_mix(dtype, dev, T1, V, kn, Dl, (long long)Dl * wl, wl, (long long)d * Dl, 1, stream)) return 1;
  // out[k, m, k0+q] = sum_{(i,j)} AbT[k,(i,j)] V[m][q,(i,j)]    -- batched over m, stored at column offset k0
  double* Ocols = static_cast<double*>(out) + (size_t)E * k0;
  if (pdmrg_gemm_nt(dtype, Dr, kn, d * Dl, wr, AbT, (long long)d * Dl, 0, V, (long long)d * Dl, (long long)kn * d * Dl, 0,
                    Ocols, (long long)wr * Dr, (long long)Dr, 1.0, 0.0, 0, stream))
    return 1;
  return 0;
}

// out[x, y, k0+a] = sum B[e,k0+a,f] R[c,dd,f] W[y,dd,b,e] Bbra[b,x,c],  a < kn   (Bbra == NULL: conj(B))
extern "C" int pdmrg_env_update_left_cols(int dtype, int d, int Dl, int Dr, int wl, int wr, const void* W_host,
                                          const void* R, const void* B, const void* Bbra, int k0, int kn,
                                          void* out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "env_update_left_cols: bad dtype");
  PDMRG_REQUIRE(k0 >= 0 && kn >= 0 && k0 + kn <= Dl, "env_update_left_cols: ket range outside [0, Dl)");
  if (!W_host) PDMRG_REQUIRE(wl == wr, "env_update_left_cols: identity MPO entry needs wl == wr");
  const ColsLayout lay = cols_layout(dtype, d, Dl, Dr, wl, wr);
  PDMRG_REQUIRE(workspace_bytes >= lay.total(), "env_update_left_cols: workspace too small (pdmrg_env_workspace_bytes)");
  if (kn == 0) return 0;
  const int E = dtype == PDMRG_C128 ? 2 : 1;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  void* op_dev = ws;
  double* BbT = reinterpret_cast<double*>(ws + lay.op_bytes);
  double* T1 = reinterpret_cast<double*>(ws + lay.op_bytes + lay.pack_a + lay.pack_b);
  double* V = reinterpret_cast<double*>(ws + lay.op_bytes + lay.pack_a + lay.pack_b + lay.t1);
  // BbT[x,b,c] = Bbra[b,x,c]  for EVERY bra index x
  if (pdmrg_permute4(dtype, Bbra ? Bbra : B, BbT, 1, Dl, d, Dr, 0, Dr, (long long)Dl * Dr, 1, Bbra ? 0 : 1, nullptr, 0, stream)) return 1;
  // T1[e][(a),(c,dd)] = sum_f B[e,k0+a,f] R[(c,dd),f]    -- batched over the physical index e (rows k0.. of each B[e])
  const double* Brows = static_cast<const double*>(B) + (size_t)E * k0 * Dr;
  if (pdmrg_gemm_nt(dtype, kn, Dr * wr, Dr, d, Brows, Dr, (long long)Dl * Dr, R, Dr, 0, 0, T1, (long long)Dr * wr,
                    (long long)kn * Dr * wr, 1.0, 0.0, 0, stream))
    return 1;
  // V[y,a,b,c] = sum_{dd,e} W[y,dd,b,e] T1[e,a,c,dd]
  MixOpHost h;
  h.rowptr.push_back(0);
  const double* W = static_cast<const double*>(W_host);
  for (int y = 0; y < wl; ++y)
    for (int b = 0; b < d; ++b) {
      for (int dd = 0; dd < wr; ++dd)
        for (int e = 0; e < d; ++e) {
          double re, im = 0.0;
          if (W) { const double* w = w_at(W, E, wr, d, y, dd, b, e); re = w[0]; if (E == 2) im = w[1]; }
          else { re = (y == dd && b == e) ? 1.0 : 0.0; }
          if (re == 0.0 && im == 0.0) continue;
          h.in_off.push_back((long long)dd + (long long)e * kn * Dr * wr);
          h.val.push_back(re); h.val.push_back(im);
        }
      h.rowptr.push_back((int)h.in_off.size());
      h.out_off.push_back((long long)y * kn * d * Dr + (long long)b * Dr);
    }
  PDMRG_REQUIRE(h.device_bytes() <= lay.op_bytes, "env_update_left_cols: operator blob overflow");
  MixOpDev dev;
  if (upload_mix_op(h, op_dev, &dev, stream)) return 1;
  if (launch_mix(dtype, dev, T1, V, kn, Dr, (long long)Dr * wr, wr, (long long)d * Dr, 1, stream)) return 1;
  // out[x, y, k0+a] = sum_{(b,c)} BbT[x,(b,c)] V[y][a,(b,c)]    -- batched over y, stored at column offset k0
  double* Ocols = static_cast<double*>(out) + (size_t)E * k0;
  if (pdmrg_gemm_nt(dtype, Dl, kn, d * Dr, wl, BbT, (long long)d * Dr, 0, V, (long long)d * Dr, (long long)kn * d * Dr, 0,
                    Ocols, (long long)wl * Dl, (long long)Dl, 1.0, 0.0, 0, stream))
    return 1;
  return 0;
}
