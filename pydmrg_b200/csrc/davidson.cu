// Restarted non-symmetric Davidson driven from C++ (reference: pydmrg/tools/la_tools.py:402-557
// with the arguments of pydmrg/tools/diag_tools.py:260).  Same algorithm as
// pydmrg_b200/davidson.py -- which remains the readable specification and the fallback for
// nroots > 4 -- but the whole cycle (mat-vec, projected-matrix update, small eigenproblem,
// residual, projection, restart) is sequenced here, so a cycle costs ~15 kernel launches and two
// pinned D2H reads instead of ~40 Python/ctypes round trips.  This matters at small bond
// dimension (chi <= 256), where a mat-vec takes 0.02-0.3 ms.
//
// The projected (<= 15+3k) x (<= 15+3k) general eigenproblem is solved on the host by a complex
// Hessenberg-QR (Householder reduction, Wilkinson-shifted QR with deflation, back substitution
// for the Schur vectors) -- the role scipy.linalg.eig plays in the reference (la_tools.py:461).
#include "ops.cuh"

#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace pdmrg {
namespace {

typedef std::complex<double> cd;

// ---- dense complex eigen-decomposition of a small matrix (column-major n x n in `a`) ---------
// returns eigenvalues w[n] and right eigenvectors in the columns of v (unit 2-norm)
int small_eig(int n, std::vector<cd> a, std::vector<cd>& w, std::vector<cd>& v) {
  auto A = [&](int i, int j) -> cd& { return a[(size_t)j * n + i]; };
  std::vector<cd> z((size_t)n * n, cd(0));
  auto Z = [&](int i, int j) -> cd& { return z[(size_t)j * n + i]; };
  for (int i = 0; i < n; ++i) Z(i, i) = 1.0;
  // Householder reduction to upper Hessenberg form, Q accumulated in Z
  for (int k = 0; k + 2 < n; ++k) {
    double alpha2 = 0;
    for (int i = k + 1; i < n; ++i) alpha2 += std::norm(A(i, k));
    const double alpha = std::sqrt(alpha2);
    if (alpha == 0.0) continue;
    std::vector<cd> u(n, cd(0));
    const cd x0 = A(k + 1, k);
    const cd ph = std::abs(x0) > 0 ? x0 / std::abs(x0) : cd(1.0);
    for (int i = k + 1; i < n; ++i) u[i] = A(i, k);
    u[k + 1] += ph * alpha;
    double un2 = 0;
    for (int i = k + 1; i < n; ++i) un2 += std::norm(u[i]);
    if (un2 == 0.0) continue;
    // A <- (I - 2 u u^H / |u|^2) A (I - 2 u u^H / |u|^2)
    for (int j = 0; j < n; ++j) {
      cd s = 0;
      for (int i = k + 1; i < n; ++i) s += std::conj(u[i]) * A(i, j);
      s *= 2.0 / un2;
      for (int i = k + 1; i < n; ++i) A(i, j) -= u[i] * s;
    }
    for (int i = 0; i < n; ++i) {
      cd s = 0;
      for (int j = k + 1; j < n; ++j) s += A(i, j) * u[j];
      s *= 2.0 / un2;
      for (int j = k + 1; j < n; ++j) A(i, j) -= s * std::conj(u[j]);
    }
    for (int i = 0; i < n; ++i) {
      cd s = 0;
      for (int j = k + 1; j < n; ++j) s += Z(i, j) * u[j];
      s *= 2.0 / un2;
      for (int j = k + 1; j < n; ++j) Z(i, j) -= s * std::conj(u[j]);
    }
    for (int i = k + 2; i < n; ++i) A(i, k) = 0.0;
  }
  // shifted QR on the Hessenberg matrix -> upper triangular T (in A), Schur vectors in Z
  double anorm = 0;
  for (int j = 0; j < n; ++j) for (int i = 0; i <= std::min(j + 1, n - 1); ++i) anorm += std::abs(A(i, j));
  if (anorm == 0) anorm = 1;
  const double eps = 2.220446049250313e-16;
  int hi = n - 1, iter = 0;
  while (hi >= 0) {
    int l = hi;
    for (; l > 0; --l) {
      double sc = std::abs(A(l - 1, l - 1)) + std::abs(A(l, l));
      if (sc == 0) sc = anorm;
      if (std::abs(A(l, l - 1)) <= eps * sc) { A(l, l - 1) = 0.0; break; }
    }
    if (l == hi) { --hi; iter = 0; continue; }
    if (++iter > 60 * n) return 1;
    // Wilkinson shift from the trailing 2x2 block; exceptional shifts now and then
    cd shift;
    if (iter % 11 == 10) {
      shift = std::abs(A(hi, hi - 1)) + std::abs(A(hi - 1, hi - 2 >= 0 ? hi - 2 : 0));
    } else {
      const cd a11 = A(hi - 1, hi - 1), a12 = A(hi - 1, hi), a21 = A(hi, hi - 1), a22 = A(hi, hi);
      const cd tr = a11 + a22, det = a11 * a22 - a12 * a21;
      const cd disc = std::sqrt(tr * tr * 0.25 - det);
      const cd e1 = tr * 0.5 + disc, e2 = tr * 0.5 - disc;
      shift = std::abs(e1 - a22) < std::abs(e2 - a22) ? e1 : e2;
    }
    // one QR sweep with Givens rotations on rows/cols l..hi
    for (int i = l; i <= hi; ++i) A(i, i) -= shift;
    std::vector<cd> cs(hi - l), sn(hi - l);
    for (int k = l; k < hi; ++k) {
      const cd f = A(k, k), g = A(k + 1, k);
      const double r = std::sqrt(std::norm(f) + std::norm(g));
      cd c = 1.0, s = 0.0;
      if (r > 0) { c = f / r; s = g / r; }
      cs[k - l] = c; sn[k - l] = s;
      // rows k, k+1:  [ conj(c)  conj(s) ; -s  c ]
      for (int j = k; j < n; ++j) {
        const cd t1 = A(k, j), t2 = A(k + 1, j);
        A(k, j) = std::conj(c) * t1 + std::conj(s) * t2;
        A(k + 1, j) = -s * t1 + c * t2;
      }
    }
    for (int k = l; k < hi; ++k) {
      const cd c = cs[k - l], s = sn[k - l];
      // columns k, k+1 times the conjugate transpose of the rotation
      for (int i = 0; i <= std::min(k + 2, hi); ++i) {
        const cd t1 = A(i, k), t2 = A(i, k + 1);
        A(i, k) = t1 * c + t2 * s;
        A(i, k + 1) = -t1 * std::conj(s) + t2 * std::conj(c);
      }
      for (int i = 0; i < n; ++i) {
        const cd t1 = Z(i, k), t2 = Z(i, k + 1);
        Z(i, k) = t1 * c + t2 * s;
        Z(i, k + 1) = -t1 * std::conj(s) + t2 * std::conj(c);
      }
    }
    for (int i = l; i <= hi; ++i) A(i, i) += shift;
  }
  // eigenvectors of the triangular T by back substitution, then x = Z y
  w.resize(n);
  v.assign((size_t)n * n, cd(0));
  for (int i = 0; i < n; ++i) w[i] = A(i, i);
  const double small = eps * anorm;
  std::vector<cd> y(n);
  for (int k = 0; k < n; ++k) {
    std::fill(y.begin(), y.end(), cd(0));
    y[k] = 1.0;
    for (int i = k - 1; i >= 0; --i) {
      cd s = 0;
      for (int j = i + 1; j <= k; ++j) s += A(i, j) * y[j];
      cd den = A(i, i) - w[k];
      if (std::abs(den) < small) den = small;
      y[i] = -s / den;
    }
    double nrm = 0;
    for (int i = 0; i < n; ++i) {
      cd s = 0;
      for (int j = 0; j <= k; ++j) s += Z(i, j) * y[j];
      v[(size_t)k * n + i] = s;
      nrm += std::norm(s);
    }
    nrm = std::sqrt(nrm);
    if (nrm > 0) for (int i = 0; i < n; ++i) v[(size_t)k * n + i] /= nrm;
  }
  return 0;
}

struct Pinned {
  double* p = nullptr;
  size_t n = 0;
  double* get(size_t doubles) {
    if (doubles > n) {
      if (p) cudaFreeHost(p);
      p = nullptr; n = 0;
      if (cudaHostAlloc((void**)&p, doubles * 8, cudaHostAllocDefault) != cudaSuccess) return nullptr;
      n = doubles;
    }
    return p;
  }
};
thread_local Pinned g_pin;      // one staging buffer per host thread (concurrent solves on different streams)

}  // namespace
}  // namespace pdmrg

using namespace pdmrg;

extern "C" int pdmrg_host_eig(int n, const double* a_colmajor_interleaved, double* w_out, double* v_out) {
  PDMRG_REQUIRE(n > 0 && n <= 256, "pdmrg_host_eig: n must be in 1..256");
  std::vector<cd> a((size_t)n * n), w, v;
  for (size_t i = 0; i < (size_t)n * n; ++i) a[i] = cd(a_colmajor_interleaved[2 * i], a_colmajor_interleaved[2 * i + 1]);
  if (small_eig(n, a, w, v)) { set_error("pdmrg_host_eig: QR iteration did not converge"); return 1; }
  for (int i = 0; i < n; ++i) { w_out[2 * i] = w[i].real(); w_out[2 * i + 1] = w[i].imag(); }
  for (size_t i = 0; i < (size_t)n * n; ++i) { v_out[2 * i] = v[i].real(); v_out[2 * i + 1] = v[i].imag(); }
  return 0;
}

extern "C" size_t pdmrg_davidson_scratch_bytes(int nroots, int max_space) {
  const int ms = max_space + 3 * nroots, rows = ms + nroots;
  return pdmrg_vec_scratch_bytes(rows) + (size_t)(2 * rows * (nroots + 3) + 64) * 16 + (size_t)(2 * nroots + 64) * 8 + 1024;
}

// XS, AX: (rows >= max_space + 4*nroots) x stride slabs;  Rv: (2*nroots + 2) x stride;  x0: nx0 vectors at x0 + i*x0_stride
// diag (optional): diagonal of the operator (sign included) -> correction t = r / (diag - theta) instead of t = r
static int davidson_impl(pdmrg_heff_plan* plan, const void* const* Ls, const void* const* Rs, double sign,
                         int dtype, long long n, const void* x0, long long x0_stride, int nx0, int nroots, int max_space,
                         double lindep, double res_tol, int max_cycle, int fixed_cycles,
                         void* XS_, void* AX_, void* Rv_, long long stride,
                         void* heff_ws, size_t heff_ws_bytes, void* scratch, size_t scratch_bytes,
                         double* e_out_host, void* vecs_out, long long vecs_stride,
                         int* n_matvec_out, int* cycles_out, double* last_res_out, void* stream_,
                         const void* diag, double pc_floor, int restart_keep = 0, int tol_relative = 0,
                         int* n_restarts_out = nullptr, int stagnation = 0, pdmrg_comm* comm = nullptr) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(plan != nullptr, "pdmrg_davidson: null plan");
  PDMRG_REQUIRE(restart_keep >= 0 && restart_keep <= 16, "pdmrg_davidson: restart_keep must be in 0..16");
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "pdmrg_davidson: bad dtype");
  PDMRG_REQUIRE(nroots >= 1 && nroots <= 4 && nx0 >= 1 && nx0 <= nroots, "pdmrg_davidson: 1..4 roots, at most one guess per root");
  PDMRG_REQUIRE(scratch_bytes >= pdmrg_davidson_scratch_bytes(nroots, max_space), "pdmrg_davidson: scratch too small");
  const bool cplx = dtype == PDMRG_C128;
  const int E = cplx ? 2 : 1;
  const size_t eb = elem_bytes(dtype);
  const int ms = max_space + 3 * nroots;             // la_tools.py:412
  const int rows = ms + nroots;
  unsigned char* sp = static_cast<unsigned char*>(scratch);
  void* vscratch = sp;
  const size_t vsb = pdmrg_vec_scratch_bytes(rows);
  double* small = reinterpret_cast<double*>(sp + vsb);                                   // 2*rows*(nroots+3)+64 elements
  double* norms = reinterpret_cast<double*>(sp + vsb + (size_t)(2 * rows * (nroots + 3) + 64) * 16);
  double* XS = static_cast<double*>(XS_);
  double* AX = static_cast<double*>(AX_);
  double* Rv = static_cast<double*>(Rv_);
  auto xs = [&](int i) { return XS + (size_t)E * i * stride; };
  auto ax = [&](int i) { return AX + (size_t)E * i * stride; };
  auto rv = [&](int i) { return Rv + (size_t)E * i * stride; };
  double* pin = g_pin.get((size_t)2 * rows * (nroots + 2) * 2 + 64);
  PDMRG_REQUIRE(pin != nullptr, "pdmrg_davidson: pinned staging allocation failed");
  double thresh2 = res_tol >= 0 ? res_tol * res_tol : lindep;
  const bool thick = restart_keep > 0;
  int n_restarts = 0;

  auto normalise_from = [&](double* vec, double* n2slot) -> int {   // vec <- vec / sqrt(n2slot[0])
    return pdmrg_vec_normalize(dtype, n, vec, n2slot, vec, stream);
  };
  // Gram-Schmidt of a short list of vectors in place (la_tools.py:586-598); uses norms[62], small tail
  auto orthonormalise = [&](std::vector<double*>& vs) -> int {
    double* gs = norms + (2 * nroots + 62);
    double* ctmp = small + (size_t)E * (2 * rows * (nroots + 3) + 32);
    for (size_t k = 0; k < vs.size(); ++k) {
      if (k == 0) {
        if (pdmrg_vec_project(dtype, n, 0, vs[0], stride, ctmp, nullptr, vs[0], gs, vscratch, vsb, stream)) return 1;
      }
      for (size_t j = 0; j < k; ++j) {
        if (pdmrg_vec_dots(dtype, n, 1, vs[j], stride, vs[k], ctmp, vscratch, vsb, stream)) return 1;
        if (pdmrg_vec_project(dtype, n, 1, vs[j], stride, ctmp, nullptr, vs[k], gs, vscratch, vsb, stream)) return 1;
      }
      if (normalise_from(vs[k], gs)) return 1;
    }
    return 0;
  };

  std::vector<cd> heff((size_t)rows * rows, cd(0));            // column-major
  auto H = [&](int i, int j) -> cd& { return heff[(size_t)j * rows + i]; };
  std::vector<cd> e(nroots), vsel((size_t)rows * nroots);
  int n_mv = 0, space = 0, cyc = 0, nr = nroots;
  bool fresh = true;
  std::vector<double*> trial;
  std::vector<const double*> x0v;
  for (int i = 0; i < nx0; ++i) x0v.push_back(static_cast<const double*>(x0) + (size_t)E * i * x0_stride);
  double last_res = 0.0;
  std::vector<double> coef;
  bool use_pc = diag != nullptr;
  double pc_best = 1e300;
  int pc_stall = 0;
  double stag_best = 1e300;
  int stag_count = 0;

  for (cyc = 0; cyc < max_cycle; ++cyc) {
    if (fresh) {
      space = 0;
      trial.clear();
      for (size_t k = 0; k < x0v.size(); ++k) {
        if (x0v[k] != xs((int)k)) PDMRG_CHECK_CUDA(cudaMemcpyAsync(xs((int)k), x0v[k], (size_t)n * eb, cudaMemcpyDeviceToDevice, stream));
        trial.push_back(xs((int)k));
      }
      if (orthonormalise(trial)) return 1;
    } else if (trial.size() > 1) {
      if (orthonormalise(trial)) return 1;
    }
    const int rnow = (int)trial.size(), head = space;
    for (int k = 0; k < rnow; ++k) {
      if (trial[k] != xs(head + k)) PDMRG_CHECK_CUDA(cudaMemcpyAsync(xs(head + k), trial[k], (size_t)n * eb, cudaMemcpyDeviceToDevice, stream));
      // comm != NULL: the plan is this rank's ket-split shard and the exchange is fused into the kernels (heff.cu
      // pdmrg_heff_apply_fused_ket): every rank gets a bit-identical image, so all ranks run this loop in lock step
      if (comm ? pdmrg_heff_apply_fused_ket(plan, comm, Ls, Rs, xs(head + k), ax(head + k), sign, heff_ws, heff_ws_bytes, stream)
               : pdmrg_heff_apply(plan, Ls, Rs, xs(head + k), ax(head + k), sign, heff_ws, heff_ws_bytes, stream))
        return 1;
      ++n_mv;
    }
    space = head + rnow;
    // projected matrix (la_tools.py:449-457): new columns <xs_i|A xt_k> (i < space), new rows conj <A xs_i|xt_k> (i < head)
    for (int k = 0; k < rnow; ++k) {
      if (pdmrg_vec_dots(dtype, n, space, XS, stride, ax(head + k), small + (size_t)E * k * 2 * rows, vscratch, vsb, stream)) return 1;
      if (head > 0 && pdmrg_vec_dots(dtype, n, head, AX, stride, xs(head + k), small + (size_t)E * (k * 2 * rows + rows), vscratch, vsb, stream)) return 1;
    }
    PDMRG_CHECK_CUDA(cudaMemcpyAsync(pin, small, (size_t)rnow * 2 * rows * eb, cudaMemcpyDeviceToHost, stream));
    PDMRG_CHECK_CUDA(cudaStreamSynchronize(stream));                       // D2H #1 of the cycle
    auto at = [&](int idx) { return cplx ? cd(pin[2 * idx], pin[2 * idx + 1]) : cd(pin[idx], 0.0); };
    for (int k = 0; k < rnow; ++k) {
      for (int i = 0; i < space; ++i) H(i, head + k) = at(k * 2 * rows + i);
      for (int i = 0; i < head; ++i) H(head + k, i) = std::conj(at(k * 2 * rows + rows + i));
    }
    std::vector<cd> sub((size_t)space * space), w, vv;
    for (int j = 0; j < space; ++j) for (int i = 0; i < space; ++i) sub[(size_t)j * space + i] = H(i, j);
    if (small_eig(space, sub, w, vv)) { set_error("pdmrg_davidson: projected eigenproblem did not converge"); return 1; }
    std::vector<int> order(space);
    for (int i = 0; i < space; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w[a].real() < w[b].real(); });   // pick_eigs, diag_tools.py:175-179
    // Tie-break (the reference sorts by the real part alone): a complex-conjugate pair has EQUAL real parts up to
    // rounding, so the target would flip between its two members from cycle to cycle and the solve would burn all
    // max_cycle iterations without converging (seen next to s = 0, where the two leading real eigenvalues of the
    // truncated generator collide).  Among (numerically) equal real parts the larger imaginary part comes first.
    for (int pass = 0; pass < 2; ++pass)
      for (int i = 0; i + 1 < space; ++i) {
        const cd a = w[order[i]], b = w[order[i + 1]];
        const double sc = std::max(1.0, std::max(std::abs(a), std::abs(b)));
        if (std::abs(a.real() - b.real()) <= 1e-9 * sc && a.imag() < b.imag()) std::swap(order[i], order[i + 1]);
      }
    nr = std::min(nroots, space);
    for (int k = 0; k < nr; ++k) {
      e[k] = w[order[k]];
      for (int i = 0; i < space; ++i) vsel[(size_t)k * rows + i] = vv[(size_t)order[k] * space + i];
      if (!cplx) {
        const double scale = std::max(1.0, std::abs(e[k].real()));
        if (std::abs(e[k].imag()) > 1e-9 * scale) { set_error("pdmrg_davidson: complex Ritz value in f64 mode; use c128"); return 3; }
        int piv = 0;
        for (int i = 1; i < space; ++i) if (std::abs(vsel[(size_t)k * rows + i]) > std::abs(vsel[(size_t)k * rows + piv])) piv = i;
        const cd ph = std::conj(vsel[(size_t)k * rows + piv]) / std::abs(vsel[(size_t)k * rows + piv]);
        for (int i = 0; i < space; ++i) vsel[(size_t)k * rows + i] = cd((vsel[(size_t)k * rows + i] * ph).real(), 0.0);
        e[k] = cd(e[k].real(), 0.0);
      }
    }
    if (tol_relative && res_tol >= 0) {                                     // ARPACK's criterion: ||r|| <= tol * |theta|
      const double sc = std::max(std::abs(e[0]), 2.220446049250313e-16);
      thresh2 = res_tol * res_tol * sc * sc;
    }
    // residuals, their projection against the basis and both norms, without a host round trip in between
    coef.resize((size_t)2 * std::max(space, rows));
    for (int k = 0; k < nr; ++k) {
      for (int i = 0; i < space; ++i) {
        if (cplx) { coef[2 * i] = vsel[(size_t)k * rows + i].real(); coef[2 * i + 1] = vsel[(size_t)k * rows + i].imag(); }
        else coef[i] = vsel[(size_t)k * rows + i].real();
      }
      double eh[2] = {e[k].real(), e[k].imag()};
      if (pdmrg_davidson_residual(dtype, n, space, coef.data(), eh, XS, AX, stride, rv(k), norms + 2 * k, vscratch, vsb, stream)) return 1;
      const double* scale_slot = norms + 2 * k;                       // ||r||^2: the projected direction starts from r / ||r||
      if (use_pc) {                                                       // ... or from t / ||t||, t = r / (diag - theta)
        double* tn = norms + 2 * nroots + 8 + k;
        if (pdmrg_vec_precond(dtype, n, rv(k), diag, e[k].real(), e[k].imag(), pc_floor, tn, vscratch, vsb, stream)) return 1;
        scale_slot = tn;
      }
      double* cdev = small + (size_t)E * (2 * rows * (nroots + 2));
      if (pdmrg_vec_dots(dtype, n, space, XS, stride, rv(k), cdev, vscratch, vsb, stream)) return 1;
      if (pdmrg_vec_project(dtype, n, space, XS, stride, cdev, scale_slot, rv(k), norms + 2 * k + 1, vscratch, vsb, stream)) return 1;
    }
    PDMRG_CHECK_CUDA(cudaMemcpyAsync(pin, norms, (size_t)2 * nr * 8, cudaMemcpyDeviceToHost, stream));
    PDMRG_CHECK_CUDA(cudaStreamSynchronize(stream));                       // D2H #2 of the cycle
    last_res = 0.0;
    trial.clear();
    std::vector<int> keep;
    for (int k = 0; k < nr; ++k) {
      const double dx2 = pin[2 * k], px2 = pin[2 * k + 1];
      if (std::sqrt(std::max(dx2, 0.0)) > last_res) last_res = std::sqrt(std::max(dx2, 0.0));
      if (dx2 > thresh2 && px2 > lindep) keep.push_back(k);                 // la_tools.py:505-537
    }
    if (use_pc) {
      // safeguard: a diagonal that is a poor model of the operator (no diagonal dominance) can stall the restarted
      // iteration; without a halving of the residual within two restart periods the identity takes over for good
      if (last_res < 0.5 * pc_best) { pc_best = last_res; pc_stall = 0; }
      else if (++pc_stall > 2 * ms) use_pc = false;
    }
    if (keep.empty()) { ++cyc; break; }                                     // la_tools.py:539-541
    if (stagnation > 0) {
      // opt-in: give up on a local problem whose residual has not been halved within `stagnation` cycles (the reference
      // would go on to max_cycle = 1000).  Next to s = 0 the truncated generator has a nearly defective leading pair and
      // single-state solves stall at ||r|| ~ 1e-4 whatever the subspace; the sweep moves on with the best Ritz pair.
      if (last_res < 0.5 * stag_best) { stag_best = last_res; stag_count = 0; }
      else if (++stag_count >= stagnation) { ++cyc; break; }
    }
    if (fixed_cycles > 0 && cyc + 1 >= fixed_cycles) { ++cyc; break; }
    fresh = space + nroots > ms;                                            // la_tools.py:544
    if (fresh && thick) {
      // Thick restart (Krylov-Schur style; opt-in, the reference restarts from the nroots Ritz vectors alone): the
      // basis is compressed to an orthonormal basis Q of the span of the kk lowest Ritz vectors,
      //   XS' = XS Q,  AX' = AX Q (no new mat-vec),  H' = Q^H H Q,
      // and the iteration goes on with the residual direction of THIS cycle, which is orthogonal to span(XS)
      // and therefore to XS'.  Nearly degenerate partners of the wanted state survive the restart, which is what
      // makes the restarted iteration converge when the gap to the second state closes (s -> 0 in the s-ensemble).
      int kk = std::min(std::max(restart_keep, nr), space - 1);
      std::vector<cd> Q((size_t)space * kk);
      int got = 0;
      for (int c = 0; c < space && got < kk; ++c) {
        const int src = order[c];
        if (!cplx && std::abs(w[src].imag()) > 1e-9 * std::max(1.0, std::abs(w[src].real()))) continue;   // f64: real pairs only
        cd* q = &Q[(size_t)got * space];
        if (c < nr) for (int i = 0; i < space; ++i) q[i] = vsel[(size_t)c * rows + i];                    // phase-fixed copies
        else for (int i = 0; i < space; ++i) q[i] = vv[(size_t)src * space + i];
        if (!cplx && c >= nr) {
          int piv = 0;
          for (int i = 1; i < space; ++i) if (std::abs(q[i]) > std::abs(q[piv])) piv = i;
          const cd ph = std::conj(q[piv]) / std::abs(q[piv]);
          for (int i = 0; i < space; ++i) q[i] = cd((q[i] * ph).real(), 0.0);
        }
        bool ok = true;
        for (int pass = 0; pass < 2 && ok; ++pass) {                        // modified Gram-Schmidt, twice
          for (int j = 0; j < got; ++j) {
            const cd* qj = &Q[(size_t)j * space];
            cd d = 0;
            for (int i = 0; i < space; ++i) d += std::conj(qj[i]) * q[i];
            for (int i = 0; i < space; ++i) q[i] -= d * qj[i];
          }
          double nn = 0;
          for (int i = 0; i < space; ++i) nn += std::norm(q[i]);
          if (nn < 1e-20) { ok = false; break; }
          nn = 1.0 / std::sqrt(nn);
          for (int i = 0; i < space; ++i) q[i] *= nn;
        }
        if (ok) ++got;
      }
      kk = got;
      if (kk >= nr) {
        for (int j = 0; j < kk; ++j) {
          for (int i = 0; i < space; ++i) {
            if (cplx) { coef[2 * i] = Q[(size_t)j * space + i].real(); coef[2 * i + 1] = Q[(size_t)j * space + i].imag(); }
            else coef[i] = Q[(size_t)j * space + i].real();
          }
          if (pdmrg_vec_combine(dtype, n, space, coef.data(), XS, stride, xs(space + j), 0, stream)) return 1;
          if (pdmrg_vec_combine(dtype, n, space, coef.data(), AX, stride, ax(space + j), 0, stream)) return 1;
        }
        for (int j = 0; j < kk; ++j) {
          PDMRG_CHECK_CUDA(cudaMemcpyAsync(xs(j), xs(space + j), (size_t)n * eb, cudaMemcpyDeviceToDevice, stream));
          PDMRG_CHECK_CUDA(cudaMemcpyAsync(ax(j), ax(space + j), (size_t)n * eb, cudaMemcpyDeviceToDevice, stream));
        }
        std::vector<cd> HQ((size_t)space * kk, cd(0)), Hn((size_t)kk * kk, cd(0));
        for (int j = 0; j < kk; ++j)
          for (int i = 0; i < space; ++i) {
            cd sacc = 0;
            for (int l = 0; l < space; ++l) sacc += H(i, l) * Q[(size_t)j * space + l];
            HQ[(size_t)j * space + i] = sacc;
          }
        for (int j = 0; j < kk; ++j)
          for (int i = 0; i < kk; ++i) {
            cd sacc = 0;
            for (int l = 0; l < space; ++l) sacc += std::conj(Q[(size_t)i * space + l]) * HQ[(size_t)j * space + l];
            Hn[(size_t)j * kk + i] = sacc;
          }
        for (int j = 0; j < rows; ++j) for (int i = 0; i < rows; ++i) H(i, j) = 0.0;
        for (int j = 0; j < kk; ++j) for (int i = 0; i < kk; ++i) H(i, j) = Hn[(size_t)j * kk + i];
        // the Ritz coefficients of the wanted states in the NEW basis (needed if the loop ends before the next solve)
        for (int k = 0; k < nr; ++k) {
          std::vector<cd> vn(kk, cd(0));
          for (int j = 0; j < kk; ++j)
            for (int l = 0; l < space; ++l) vn[j] += std::conj(Q[(size_t)j * space + l]) * vsel[(size_t)k * rows + l];
          for (int i = 0; i < rows; ++i) vsel[(size_t)k * rows + i] = i < kk ? vn[i] : cd(0);
        }
        space = kk;
        fresh = false;
        ++n_restarts;
      }
    }
    if (!fresh) {
      for (int k : keep) {
        // a single new direction is normalised straight into its basis slot (no copy next cycle)
        double* dst = keep.size() == 1 ? xs(space) : rv(k);
        if (pdmrg_vec_normalize(dtype, n, rv(k), norms + 2 * k + 1, dst, stream)) return 1;
        trial.push_back(dst);
      }
    }
    if (fresh) {
      x0v.clear();
      for (int k = 0; k < nr; ++k) {
        for (int i = 0; i < space; ++i) {
          if (cplx) { coef[2 * i] = vsel[(size_t)k * rows + i].real(); coef[2 * i + 1] = vsel[(size_t)k * rows + i].imag(); }
          else coef[i] = vsel[(size_t)k * rows + i].real();
        }
        if (pdmrg_vec_combine(dtype, n, space, coef.data(), XS, stride, rv(nr + k), 0, stream)) return 1;
        x0v.push_back(rv(nr + k));
      }
    }
  }
  // Ritz vectors (la_tools.py:469)
  for (int k = 0; k < nr; ++k) {
    for (int i = 0; i < space; ++i) {
      if (cplx) { coef[2 * i] = vsel[(size_t)k * rows + i].real(); coef[2 * i + 1] = vsel[(size_t)k * rows + i].imag(); }
      else coef[i] = vsel[(size_t)k * rows + i].real();
    }
    if (pdmrg_vec_combine(dtype, n, space, coef.data(), XS, stride, static_cast<double*>(vecs_out) + (size_t)E * k * vecs_stride, 0, stream)) return 1;
    e_out_host[2 * k] = e[k].real();
    e_out_host[2 * k + 1] = e[k].imag();
  }
  if (n_matvec_out) *n_matvec_out = n_mv;
  if (n_restarts_out) *n_restarts_out = n_restarts;
  if (cycles_out) *cycles_out = cyc;
  if (last_res_out) *last_res_out = last_res;
  return 0;
}

extern "C" int pdmrg_davidson(pdmrg_heff_plan* plan, const void* const* Ls, const void* const* Rs, double sign,
                              int dtype, long long n, const void* x0, long long x0_stride, int nx0, int nroots, int max_space,
                              double lindep, double res_tol, int max_cycle, int fixed_cycles,
                              void* XS_, void* AX_, void* Rv_, long long stride,
                              void* heff_ws, size_t heff_ws_bytes, void* scratch, size_t scratch_bytes,
                              double* e_out_host, void* vecs_out, long long vecs_stride,
                              int* n_matvec_out, int* cycles_out, double* last_res_out, void* stream_) {
  return davidson_impl(plan, Ls, Rs, sign, dtype, n, x0, x0_stride, nx0, nroots, max_space, lindep, res_tol, max_cycle,
                       fixed_cycles, XS_, AX_, Rv_, stride, heff_ws, heff_ws_bytes, scratch, scratch_bytes, e_out_host,
                       vecs_out, vecs_stride, n_matvec_out, cycles_out, last_res_out, stream_, nullptr, 0.0);
}

extern "C" int pdmrg_davidson_precond(pdmrg_heff_plan* plan, const void* const* Ls, const void* const* Rs, double sign,
                                      int dtype, long long n, const void* x0, long long x0_stride, int nx0, int nroots,
                                      int max_space, double lindep, double res_tol, int max_cycle, int fixed_cycles,
                                      void* XS_, void* AX_, void* Rv_, long long stride,
                                      void* heff_ws, size_t heff_ws_bytes, void* scratch, size_t scratch_bytes,
                                      double* e_out_host, void* vecs_out, long long vecs_stride,
                                      int* n_matvec_out, int* cycles_out, double* last_res_out, void* stream_,
                                      const void* diag, double pc_floor) {
  PDMRG_REQUIRE(diag != nullptr && pc_floor > 0.0, "pdmrg_davidson_precond: needs the operator diagonal and a positive floor");
  return davidson_impl(plan, Ls, Rs, sign, dtype, n, x0, x0_stride, nx0, nroots, max_space, lindep, res_tol, max_cycle,
                       fixed_cycles, XS_, AX_, Rv_, stride, heff_ws, heff_ws_bytes, scratch, scratch_bytes, e_out_host,
                       vecs_out, vecs_stride, n_matvec_out, cycles_out, last_res_out, stream_, diag, pc_floor);
}

// The same loop with the solver options that go beyond the reference's algorithm (all opt-in):
//   restart_keep > 0 : thick restart -- the basis is compressed to the restart_keep lowest Ritz vectors instead of
//                      being thrown away (XS / AX need max_space + 4*nroots + restart_keep rows);
//   tol_relative     : res_tol is relative to |theta| (ARPACK's stopping rule, scipy eigs(tol=...), diag_tools.py:225);
//   diag (optional)  : diagonal preconditioner as in pdmrg_davidson_precond;
//   stagnation > 0   : stop when the residual has not been halved within that many cycles.
// With the identity preconditioner the expansion vectors are residuals, i.e. the subspace is the Krylov space of the
// start vector: restart_keep > 0 then IS thick-restarted Arnoldi (Krylov-Schur), the device-resident counterpart of the
// reference's ARPACK path (calc_eigs_arnoldi, diag_tools.py:214-241).
extern "C" int pdmrg_davidson_ex(pdmrg_heff_plan* plan, const void* const* Ls, const void* const* Rs, double sign,
                                 int dtype, long long n, const void* x0, long long x0_stride, int nx0, int nroots,
                                 int max_space, double lindep, double res_tol, int max_cycle, int fixed_cycles,
                                 void* XS_, void* AX_, void* Rv_, long long stride,
                                 void* heff_ws, size_t heff_ws_bytes, void* scratch, size_t scratch_bytes,
                                 double* e_out_host, void* vecs_out, long long vecs_stride,
                                 int* n_matvec_out, int* cycles_out, double* last_res_out, void* stream_,
                                 const void* diag, double pc_floor, int restart_keep, int tol_relative, int* n_restarts_out,
                                 int stagnation) {
  PDMRG_REQUIRE(diag == nullptr || pc_floor > 0.0, "pdmrg_davidson_ex: the preconditioner needs a positive floor");
  return davidson_impl(plan, Ls, Rs, sign, dtype, n, x0, x0_stride, nx0, nroots, max_space, lindep, res_tol, max_cycle,
                       fixed_cycles, XS_, AX_, Rv_, stride, heff_ws, heff_ws_bytes, scratch, scratch_bytes, e_out_host,
                       vecs_out, vecs_stride, n_matvec_out, cycles_out, last_res_out, stream_, diag, pc_floor, restart_keep,
                       tol_relative, n_restarts_out, stagnation);
}

// The loop of pdmrg_davidson_ex over a SHARDED operator: `plan` is this rank's ket-split shard
// (pdmrg_heff_plan_create_ketsplit / _create_host with a ket range), `comm` the peer-memory communicator; each mat-vec is
// pdmrg_heff_apply_fused_ket.  Every rank must call it with identical start vectors and options: the images are
// bit-identical on all ranks and every vector kernel reduces in a fixed order, so the ranks take identical control
// decisions without exchanging a single control scalar (a rank that diverged would trip the bounded flag waits:
// pdmrg_comm_error).
extern "C" int pdmrg_davidson_sharded(pdmrg_heff_plan* plan, pdmrg_comm* comm, const void* const* Ls, const void* const* Rs,
                                      double sign, int dtype, long long n, const void* x0, long long x0_stride, int nx0,
                                      int nroots, int max_space, double lindep, double res_tol, int max_cycle, int fixed_cycles,
                                      void* XS_, void* AX_, void* Rv_, long long stride,
                                      void* heff_ws, size_t heff_ws_bytes, void* scratch, size_t scratch_bytes,
                                      double* e_out_host, void* vecs_out, long long vecs_stride,
                                      int* n_matvec_out, int* cycles_out, double* last_res_out, void* stream_,
                                      const void* diag, double pc_floor, int restart_keep, int tol_relative,
                                      int* n_restarts_out, int stagnation) {
  PDMRG_REQUIRE(comm != nullptr, "pdmrg_davidson_sharded: null communicator");
  PDMRG_REQUIRE(diag == nullptr || pc_floor > 0.0, "pdmrg_davidson_sharded: the preconditioner needs a positive floor");
  return davidson_impl(plan, Ls, Rs, sign, dtype, n, x0, x0_stride, nx0, nroots, max_space, lindep, res_tol, max_cycle,
                       fixed_cycles, XS_, AX_, Rv_, stride, heff_ws, heff_ws_bytes, scratch, scratch_bytes, e_out_host,
                       vecs_out, vecs_stride, n_matvec_out, cycles_out, last_res_out, stream_, diag, pc_floor, restart_keep,
                       tol_relative, n_restarts_out, stagnation, comm);
}
