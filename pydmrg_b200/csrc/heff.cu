// Effective-Hamiltonian mat-vec (H1/H2) and environment updates (E1/E2) composed from the
// DMMA GEMM tiles (gemm_nt.cu) and the sparse W "mix" (tensor_ops.cu).
//
// Mat-vec, per MPO term  (reference: pydmrg/tools/diag_tools.py:114-119 one-site,
// :153-156 two-site; P = d or d1*d2):
//   stage A   T[(r,o),(pin,k)] = sum_s R[r,o,s] x[pin,k,s]        GEMM  (Dr*wR) x (P*Dl) x Dr
//   mix       U[pout,r,j,k]    = sum Wc[(j,pout),(o,pin)] T[r,o,pin,k]   point-wise, sparse
//   stage C   y[pout,i,r]     += sign sum_{j,k} L[i,j,k] U[pout,r,j,k]   GEMM  Dl x Dr x (wL*Dl), batch P
// Stage A is issued as a GEMM batched over the MPO-bond index o of R (and stage C over
// contiguous runs of j), so all-zero columns / rows of the local operator -- MPO block
// sparsity, or the blocks owned by other ranks in the bond-sharded multi-GPU mode --
// cost no FLOPs.  Every index order is chosen so that both GEMMs are plain "NT" products
// on the tensors exactly as the reference stores them (no packing of x, L, R or y).
#include "ops.cuh"

#include <cmath>

using namespace pdmrg;

struct pdmrg_heff_plan {
  int dtype, P, Dl, Dr, n_mpo;
  std::vector<int> wL, wR;
  std::vector<MixOpDev> mix;                         // per MPO term
  std::vector<std::vector<int>> need_o, need_j;      // per MPO term: flags
  void* dev_blob = nullptr;
  bool own_blob = true;                              // false: the caller owns the device blob (pdmrg_heff_plan_upload)
  std::vector<MixOpHost> pending;                    // host operators waiting for pdmrg_heff_plan_upload
  size_t blob_bytes = 0;
  double flops_dense = 0, flops_sparse = 0, flops_gemm = 0;
  size_t ws_bytes = 0;
  int k0 = 0, kl = 0;      // ket-index range [k0, k0+kl) of the LEFT bond this plan contracts (kl = Dl: all)
};

namespace {

int build_plan(pdmrg_heff_plan** out, int dtype, int P, int Dl, int Dr, int n_mpo, const int* wL,
               const int* wR, const void* const* Wc_host, const unsigned char* const* mask, int k0 = 0, int klen = -1,
               bool defer_upload = false) {
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "heff plan: bad dtype");
  PDMRG_REQUIRE(P > 0 && Dl > 0 && Dr > 0 && n_mpo > 0, "heff plan: bad extents");
  if (klen < 0) klen = Dl;
  PDMRG_REQUIRE(k0 >= 0 && klen >= 0 && k0 + klen <= Dl, "heff plan: ket range outside the left bond");
  auto* pl = new pdmrg_heff_plan();
  pl->dtype = dtype; pl->P = P; pl->Dl = Dl; pl->Dr = Dr; pl->n_mpo = n_mpo;
  pl->k0 = k0; pl->kl = klen;
  const int kl = klen;
  const int E = dtype == PDMRG_C128 ? 2 : 1;
  const double cmac = dtype == PDMRG_C128 ? 8.0 : 2.0;
  std::vector<MixOpHost> hosts(n_mpo);
  size_t blob_bytes = 0;
  for (int m = 0; m < n_mpo; ++m) {
    const int wl = wL[m], wr = wR[m];
    pl->wL.push_back(wl); pl->wR.push_back(wr);
    const double* W = static_cast<const double*>(Wc_host[m]);
    const int rows = wl * P, cols = wr * P;
    std::vector<int> need_o(wr, 0), need_j(wl, 0);
    MixOpHost& h = hosts[m];
    h.rowptr.push_back(0);
    long long nnz = 0;
    // rows are ordered (j, pout); input T[r,o,pin,k]; output U[pout,r,j,k]
    for (int j = 0; j < wl; ++j)
      for (int po = 0; po < P; ++po) {
        const int row = j * P + po;
        for (int o = 0; o < wr; ++o) {
          if (mask && mask[m] && !mask[m][j * wr + o]) continue;
          for (int pi = 0; pi < P; ++pi) {
            const double* w = W + (size_t)E * ((size_t)row * cols + o * P + pi);
            const bool nz = (w[0] != 0.0) || (E == 2 && w[1] != 0.0);
            if (!nz) continue;
            h.in_off.push_back(((long long)o * P + pi) * kl);
            h.val.push_back(w[0]);
            h.val.push_back(E == 2 ? w[1] : 0.0);
            need_o[o] = 1; need_j[j] = 1;
            ++nnz;
          }
        }
        h.rowptr.push_back((int)h.in_off.size());
        h.out_off.push_back((long long)po * Dr * wl * kl + (long long)j * kl);
      }
    // drop rows of slabs j that are never read by stage C
    MixOpHost hc;
    hc.rowptr.push_back(0);
    for (int j = 0; j < wl; ++j) {
      if (!need_j[j]) continue;
      for (int po = 0; po < P; ++po) {
        const int row = j * P + po;
        for (int k = h.rowptr[row]; k < h.rowptr[row + 1]; ++k) {
          hc.in_off.push_back(h.in_off[k]);
          hc.val.push_back(h.val[2 * k]); hc.val.push_back(h.val[2 * k + 1]);
        }
        hc.rowptr.push_back((int)hc.in_off.size());
        hc.out_off.push_back(h.out_off[row]);
      }
    }
    h = hc;
    pl->need_o.push_back(need_o); pl->need_j.push_back(need_j);
    blob_bytes += align_up(h.device_bytes(), 256);
    int no = 0, nj = 0;
    for (int v : need_o) no += v;
    for (int v : need_j) nj += v;
    const double cube_a = (double)Dr * Dr * kl, cube_c = (double)Dl * kl * Dr;
    pl->flops_dense += cmac * (P * (wr * cube_a + wl * cube_c) + (double)rows * cols * kl * Dr);
    pl->flops_sparse += cmac * (P * (no * cube_a + nj * cube_c) + (double)nnz * kl * Dr);
    pl->flops_gemm += cmac * P * (no * cube_a + nj * cube_c);
    const size_t t_elems = (size_t)Dr * (wr > wl ? wr : wl) * P * Dl, u_elems = (size_t)P * Dr * wl * kl;   // T also hosts Z
    const size_t need = align_up(t_elems * 8 * E, 256) + align_up(u_elems * 8 * E, 256);
    if (need > pl->ws_bytes) pl->ws_bytes = need;
  }
  pl->blob_bytes = blob_bytes ? blob_bytes : 256;
  if (defer_upload) {
    // no device allocation, no copy on the legacy stream, no synchronisation: the caller provides the device blob and
    // the stream (pdmrg_heff_plan_upload) -- plans are created and dropped once per bond step, and cudaMalloc / cudaFree
    // would synchronise the whole device under the other scan points running concurrently on their own streams
    pl->own_blob = false;
    pl->pending = std::move(hosts);
    *out = pl;
    return 0;
  }
  cudaError_t e = cudaMalloc(&pl->dev_blob, blob_bytes ? blob_bytes : 256);
  if (e != cudaSuccess) { set_error("heff plan: cudaMalloc(%zu) failed: %s", blob_bytes, cudaGetErrorString(e)); delete pl; return 1; }
  size_t off = 0;
  for (int m = 0; m < n_mpo; ++m) {
    MixOpDev dev;
    if (upload_mix_op(hosts[m], static_cast<unsigned char*>(pl->dev_blob) + off, &dev, 0)) { cudaFree(pl->dev_blob); delete pl; return 1; }
    pl->mix.push_back(dev);
    off += align_up(hosts[m].device_bytes(), 256);
  }
  e = cudaStreamSynchronize(0);
  if (e != cudaSuccess) { set_error("heff plan: upload failed: %s", cudaGetErrorString(e)); cudaFree(pl->dev_blob); delete pl; return 1; }
  *out = pl;
  return 0;
}

}  // namespace

extern "C" int pdmrg_heff_plan_create(pdmrg_heff_plan** plan, int dtype, int P, int Dl, int Dr, int n_mpo,
                                      const int* wL, const int* wR, const void* const* Wc_host) {
  return build_plan(plan, dtype, P, Dl, Dr, n_mpo, wL, wR, Wc_host, nullptr);
}

extern "C" int pdmrg_heff_plan_create_sharded(pdmrg_heff_plan** plan, int dtype, int P, int Dl, int Dr, int n_mpo,
                                              const int* wL, const int* wR, const void* const* Wc_host,
                                              const unsigned char* const* block_mask_host) {
  return build_plan(plan, dtype, P, Dl, Dr, n_mpo, wL, wR, Wc_host, block_mask_host);
}

extern "C" int pdmrg_heff_plan_create_ketsplit(pdmrg_heff_plan** plan, int dtype, int P, int Dl, int Dr, int n_mpo,
                                               const int* wL, const int* wR, const void* const* Wc_host,
                                               const unsigned char* const* block_mask_host, int k0, int klen) {
  return build_plan(plan, dtype, P, Dl, Dr, n_mpo, wL, wR, Wc_host, block_mask_host, k0, klen);
}

extern "C" int pdmrg_heff_plan_create_host(pdmrg_heff_plan** plan, int dtype, int P, int Dl, int Dr, int n_mpo,
                                           const int* wL, const int* wR, const void* const* Wc_host,
                                           const unsigned char* const* block_mask_host, int k0, int klen) {
  return build_plan(plan, dtype, P, Dl, Dr, n_mpo, wL, wR, Wc_host, block_mask_host, k0, klen, true);
}

extern "C" size_t pdmrg_heff_plan_blob_bytes(const pdmrg_heff_plan* plan) { return plan ? plan->blob_bytes : 0; }

extern "C" int pdmrg_heff_plan_upload(pdmrg_heff_plan* pl, void* blob_dev, size_t blob_bytes, void* stream_) {
  PDMRG_REQUIRE(pl && !pl->own_blob && (int)pl->pending.size() == pl->n_mpo, "pdmrg_heff_plan_upload: plan was not created by pdmrg_heff_plan_create_host (or is already uploaded)");
  PDMRG_REQUIRE(blob_dev != nullptr && blob_bytes >= pl->blob_bytes, "pdmrg_heff_plan_upload: device blob too small");
  PDMRG_REQUIRE((reinterpret_cast<uintptr_t>(blob_dev) & 255) == 0, "pdmrg_heff_plan_upload: device blob must be 256-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  size_t off = 0;
  pl->mix.clear();
  for (int m = 0; m < pl->n_mpo; ++m) {
    MixOpDev dev;
    if (upload_mix_op(pl->pending[m], static_cast<unsigned char*>(blob_dev) + off, &dev, stream)) return 1;
    pl->mix.push_back(dev);
    off += align_up(pl->pending[m].device_bytes(), 256);
  }
  pl->dev_blob = blob_dev;
  pl->pending.clear();
  return 0;
}

extern "C" void pdmrg_heff_plan_destroy(pdmrg_heff_plan* plan) {
  if (!plan) return;
  if (plan->dev_blob && plan->own_blob) cudaFree(plan->dev_blob);
  delete plan;
}

extern "C" size_t pdmrg_heff_workspace_bytes(const pdmrg_heff_plan* plan) { return plan ? plan->ws_bytes : 0; }
extern "C" double pdmrg_heff_flops_dense(const pdmrg_heff_plan* plan) { return plan ? plan->flops_dense : 0; }
extern "C" double pdmrg_heff_flops_sparse(const pdmrg_heff_plan* plan) { return plan ? plan->flops_sparse : 0; }
extern "C" double pdmrg_heff_flops_gemm(const pdmrg_heff_plan* plan) { return plan ? plan->flops_gemm : 0; }

namespace {
struct StageTimer {
  cudaEvent_t ev[4];
  float ms[3] = {0, 0, 0};
  bool on = false;
  cudaStream_t stream = nullptr;
  int mark(int i) {
    if (!on) return 0;
    PDMRG_CHECK_CUDA(cudaEventRecord(ev[i], stream));
    return 0;
  }
  int collect() {
    if (!on) return 0;
    PDMRG_CHECK_CUDA(cudaEventSynchronize(ev[3]));
    for (int i = 0; i < 3; ++i) {
      float t = 0;
      PDMRG_CHECK_CUDA(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
      ms[i] += t;
    }
    return 0;
  }
};
int heff_apply_impl(pdmrg_heff_plan* pl, const void* const* Ls, const void* const* Rs, const void* x, void* y,
                    double sign, void* workspace, size_t workspace_bytes, cudaStream_t stream, StageTimer* tm);
}  // namespace

extern "C" int pdmrg_heff_apply(pdmrg_heff_plan* pl, const void* const* Ls, const void* const* Rs, const void* x,
                                void* y, double sign, void* workspace, size_t workspace_bytes, void* stream_) {
  return heff_apply_impl(pl, Ls, Rs, x, y, sign, workspace, workspace_bytes, static_cast<cudaStream_t>(stream_), nullptr);
}

extern "C" int pdmrg_heff_apply_timed(pdmrg_heff_plan* pl, const void* const* Ls, const void* const* Rs, const void* x,
                                      void* y, double sign, void* workspace, size_t workspace_bytes, void* stream_,
                                      float* ms3_host) {
  StageTimer tm;
  tm.on = true;
  tm.stream = static_cast<cudaStream_t>(stream_);
  for (int i = 0; i < 4; ++i) PDMRG_CHECK_CUDA(cudaEventCreate(&tm.ev[i]));
  int rc = heff_apply_impl(pl, Ls, Rs, x, y, sign, workspace, workspace_bytes, tm.stream, &tm);
  for (int i = 0; i < 3; ++i) ms3_host[i] = tm.ms[i];
  for (int i = 0; i < 4; ++i) cudaEventDestroy(tm.ev[i]);
  return rc;
}

namespace {
int heff_apply_impl(pdmrg_heff_plan* pl, const void* const* Ls, const void* const* Rs, const void* x, void* y,
                    double sign, void* workspace, size_t workspace_bytes, cudaStream_t stream, StageTimer* tm) {
  PDMRG_REQUIRE(pl != nullptr, "pdmrg_heff_apply: null plan");
  PDMRG_REQUIRE((int)pl->mix.size() == pl->n_mpo, "pdmrg_heff_apply: plan not uploaded (pdmrg_heff_plan_upload)");
  PDMRG_REQUIRE(workspace_bytes >= pl->ws_bytes, "pdmrg_heff_apply: workspace too small");
  const int E = pl->dtype == PDMRG_C128 ? 2 : 1;
  const int P = pl->P, Dl = pl->Dl, Dr = pl->Dr;
  bool wrote_y = false;
  for (int m = 0; m < pl->n_mpo; ++m) {
    const int wl = pl->wL[m], wr = pl->wR[m];
    const size_t t_elems = (size_t)Dr * (wr > wl ? wr : wl) * P * Dl;
    double* T = static_cast<double*>(workspace);
    double* U = reinterpret_cast<double*>(static_cast<unsigned char*>(workspace) + align_up(t_elems * 8 * E, 256));
    const double* R = static_cast<const double*>(Rs[m]);
    const double* L = static_cast<const double*>(Ls[m]);
    // ---- stage A: ONE launch, batched over the non-empty o slabs (index list)
    if (tm && tm->mark(0)) return 1;
    std::vector<int> olist, jlist;
    for (int o = 0; o < wr; ++o) if (pl->need_o[m][o]) olist.push_back(o);
    for (int j = 0; j < wl; ++j) if (pl->need_j[m][j]) jlist.push_back(j);
    if (olist.empty() || jlist.empty()) continue;
    PDMRG_REQUIRE(olist.size() <= 32 && jlist.size() <= 32, "pdmrg_heff_apply: MPO bond dimension above 32 not supported yet");
    const int kl = pl->kl, k0 = pl->k0;
    if (kl == 0) continue;
    if (kl == Dl) {
      if (pdmrg_gemm_nt_batched2(pl->dtype, Dr, P * Dl, Dr, 1, (int)olist.size(), olist.data(),
                                 R, (long long)wr * Dr, 0, Dr,
                                 x, Dr, 0, 0, 0,
                                 T, (long long)wr * P * Dl, 0, (long long)P * Dl,
                                 1.0, 0.0, 0, stream))
        return 1;
    } else {
      // ket-split shard: only the rows k in [k0, k0+kl) of every x[pin] block enter (batch level 1 = pin)
      if (pdmrg_gemm_nt_batched2(pl->dtype, Dr, kl, Dr, P, (int)olist.size(), olist.data(),
                                 R, (long long)wr * Dr, 0, Dr,
                                 static_cast<const double*>(x) + (size_t)E * k0 * Dr, Dr, (long long)Dl * Dr, 0, 0,
                                 T, (long long)wr * P * kl, kl, (long long)P * kl,
                                 1.0, 0.0, 0, stream))
        return 1;
    }
    // ---- mix
    if (tm && tm->mark(1)) return 1;
    if (launch_mix(pl->dtype, pl->mix[m], T, U, Dr, kl, (long long)wr * P * kl, 1, (long long)wl * kl, 1, stream)) return 1;
    if (tm && tm->mark(2)) return 1;
    // ---- stage C: split over the MPO bond j (ONE launch, P x |j| batches, K = Dl each) into the slabs
    //      Z[pout][j][i][r] (aliases T, which is dead after the mix), then an ordered reduction over j:
    //      removes the partial-wave tail of the long-K formulation and fills the GPU at small chi.
    double* Z = T;
    if (pdmrg_gemm_nt_batched2(pl->dtype, Dl, Dr, kl, P, (int)jlist.size(), jlist.data(),
                               L + (size_t)E * k0, (long long)wl * Dl, 0, Dl,
                               U, (long long)wl * kl, (long long)Dr * wl * kl, kl, 0,
                               Z, Dr, (long long)wl * Dl * Dr, (long long)Dl * Dr,
                               1.0, 0.0, 0, stream))
      return 1;
    if (launch_reduce_slabs(pl->dtype, Z, y, (long long)Dl * Dr, P, (int)jlist.size(), jlist.data(),
                            (long long)wl * Dl * Dr, (long long)Dl * Dr, sign, wrote_y ? 1 : 0, stream))
      return 1;
    wrote_y = true;
    if (tm && (tm->mark(3) || tm->collect())) return 1;
  }
  if (!wrote_y) PDMRG_CHECK_CUDA(cudaMemsetAsync(y, 0, (size_t)P * Dl * Dr * 8 * E, stream));
  return 0;
}
}  // namespace

// ---------------------------------------------------------------------------------------
// Bond-sharded mat-vec with the exchange fused into the kernels (peer memory, no NCCL)
// ---------------------------------------------------------------------------------------
extern "C" size_t pdmrg_comm_z_bytes(const pdmrg_heff_plan* pl, int world) {
  if (!pl) return 0;
  const int E = pl->dtype == PDMRG_C128 ? 2 : 1;
  const int chunk = (pl->Dl + world - 1) / world;
  int wl = 1;
  for (int v : pl->wL) if (v > wl) wl = v;
  return (size_t)world * pl->P * wl * chunk * pl->Dr * 8 * E + 256;
}

extern "C" size_t pdmrg_comm_y_bytes(const pdmrg_heff_plan* pl) {
  if (!pl) return 0;
  return (size_t)pl->P * pl->Dl * pl->Dr * elem_bytes(pl->dtype) + 256;
}

extern "C" int pdmrg_heff_needed_j(const pdmrg_heff_plan* pl, int mpo, unsigned int* mask_out) {
  PDMRG_REQUIRE(pl && mpo >= 0 && mpo < pl->n_mpo, "pdmrg_heff_needed_j: bad arguments");
  unsigned int m = 0;
  for (int j = 0; j < pl->wL[mpo] && j < 32; ++j) if (pl->need_j[mpo][j]) m |= 1u << j;
  *mask_out = m;
  return 0;
}

extern "C" int pdmrg_heff_apply_fused(pdmrg_heff_plan* pl, pdmrg_comm* comm, const void* const* Ls, const void* const* Rs,
                                      const void* x, void* y, double sign, const unsigned int* jmask_all_ranks,
                                      void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(pl && comm, "pdmrg_heff_apply_fused: null plan / communicator");
  PDMRG_REQUIRE(pl->n_mpo == 1, "pdmrg_heff_apply_fused: one MPO term per plan");
  PDMRG_REQUIRE(pl->kl == pl->Dl, "pdmrg_heff_apply_fused: ket-split plans use the all-reduce path");
  PDMRG_REQUIRE(workspace_bytes >= pl->ws_bytes, "pdmrg_heff_apply_fused: workspace too small");
  const int E = pl->dtype == PDMRG_C128 ? 2 : 1;
  const int P = pl->P, Dl = pl->Dl, Dr = pl->Dr, wl = pl->wL[0], wr = pl->wR[0];
  const int world = comm_world(comm), rank = comm_rank(comm);
  const int chunk = (Dl + world - 1) / world;
  const size_t t_elems = (size_t)Dr * (wr > wl ? wr : wl) * P * Dl;
  double* T = static_cast<double*>(workspace);
  double* U = reinterpret_cast<double*>(static_cast<unsigned char*>(workspace) + align_up(t_elems * 8 * E, 256));
  std::vector<int> olist, jlist;
  for (int o = 0; o < wr; ++o) if (pl->need_o[0][o]) olist.push_back(o);
  for (int j = 0; j < wl; ++j) if (pl->need_j[0][j]) jlist.push_back(j);
  PDMRG_REQUIRE(olist.size() <= 32 && jlist.size() <= 32 && wl <= 32, "pdmrg_heff_apply_fused: MPO bond above 32");
  comm_next_epoch(comm);
  if (!olist.empty() && !jlist.empty()) {
    const double* R = static_cast<const double*>(Rs[0]);
    const double* L = static_cast<const double*>(Ls[0]);
    if (pdmrg_gemm_nt_batched2(pl->dtype, Dr, P * Dl, Dr, 1, (int)olist.size(), olist.data(), R, (long long)wr * Dr, 0, Dr,
                               x, Dr, 0, 0, 0, T, (long long)wr * P * Dl, 0, (long long)P * Dl, 1.0, 0.0, 0, stream))
      return 1;
    if (launch_mix(pl->dtype, pl->mix[0], T, U, Dr, Dl, (long long)wr * P * Dl, 1, (long long)wl * Dl, 1, stream)) return 1;
    // stage C with the SCATTER epilogue: rows [q*chunk, (q+1)*chunk) of every (pout, j) slab are stored
    // into rank q's Z buffer at [src = this rank][pout][j][il][r]
    const void* Cpeer[8];
    for (int q = 0; q < world; ++q) Cpeer[q] = comm_zpeer(comm, q) + (size_t)E * rank * P * wl * chunk * Dr;
    if (pdmrg_gemm_nt_scatter(pl->dtype, Dl, Dr, Dl, P, (int)jlist.size(), jlist.data(), L, (long long)wl * Dl, 0, Dl,
                              U, (long long)wl * Dl, (long long)Dr * wl * Dl, Dl, 0, Cpeer, world, chunk, Dr,
                              (long long)wl * chunk * Dr, (long long)chunk * Dr, 1.0, 0.0, 0, stream))
      return 1;
  }
  if (comm_signal(comm, 0, stream)) return 1;
  if (comm_reduce_gather(comm, jmask_all_ranks, P, wl, chunk, Dl, Dr, E, sign, stream)) return 1;
  if (comm_signal(comm, 1, stream)) return 1;
  if (comm_wait_copy(comm, y, (long long)P * Dl * Dr * E, stream)) return 1;
  return 0;
}

extern "C" int pdmrg_heff_apply_fused_ket(pdmrg_heff_plan* pl, pdmrg_comm* comm, const void* const* Ls, const void* const* Rs,
                                          const void* x, void* y, double sign, void* workspace, size_t workspace_bytes,
                                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(pl && comm, "pdmrg_heff_apply_fused_ket: null plan / communicator");
  PDMRG_REQUIRE(pl->n_mpo == 1, "pdmrg_heff_apply_fused_ket: one MPO term per plan");
  PDMRG_REQUIRE(workspace_bytes >= pl->ws_bytes, "pdmrg_heff_apply_fused_ket: workspace too small");
  const int E = pl->dtype == PDMRG_C128 ? 2 : 1;
  const int P = pl->P, Dl = pl->Dl, Dr = pl->Dr, wl = pl->wL[0], wr = pl->wR[0];
  const int kl = pl->kl, k0 = pl->k0;
  const int world = comm_world(comm), rank = comm_rank(comm);
  const int chunk = (Dl + world - 1) / world;
  const size_t t_elems = (size_t)Dr * (wr > wl ? wr : wl) * P * Dl;
  double* T = static_cast<double*>(workspace);
  double* U = reinterpret_cast<double*>(static_cast<unsigned char*>(workspace) + align_up(t_elems * 8 * E, 256));
  std::vector<int> olist, jlist;
  for (int o = 0; o < wr; ++o) if (pl->need_o[0][o]) olist.push_back(o);
  for (int j = 0; j < wl; ++j) if (pl->need_j[0][j]) jlist.push_back(j);
  PDMRG_REQUIRE(olist.size() <= 32 && jlist.size() <= 32, "pdmrg_heff_apply_fused_ket: MPO bond above 32");
  PDMRG_REQUIRE(kl > 0 && !olist.empty() && !jlist.empty(), "pdmrg_heff_apply_fused_ket: every rank needs a non-empty ket range");
  comm_next_epoch(comm);
  {
    const double* R = static_cast<const double*>(Rs[0]);
    const double* L = static_cast<const double*>(Ls[0]);
    if (pdmrg_gemm_nt_batched2(pl->dtype, Dr, kl, Dr, P, (int)olist.size(), olist.data(), R, (long long)wr * Dr, 0, Dr,
                               static_cast<const double*>(x) + (size_t)E * k0 * Dr, Dr, (long long)Dl * Dr, 0, 0,
                               T, (long long)wr * P * kl, kl, (long long)P * kl, 1.0, 0.0, 0, stream))
      return 1;
    if (launch_mix(pl->dtype, pl->mix[0], T, U, Dr, kl, (long long)wr * P * kl, 1, (long long)wl * kl, 1, stream)) return 1;
    // stage C as ONE long-K product: y_partial[pout][i, r] = sum_j sum_{k in range} L[i, j, k] U[pout][r, j, k]
    // (the |j| slabs are folded into the k-loop), tiles scattered to Z_q[src = rank][pout][il][r]
    const void* Cpeer[8];
    for (int q = 0; q < world; ++q) Cpeer[q] = comm_zpeer(comm, q) + (size_t)E * rank * P * chunk * Dr;
    if (pdmrg_gemm_nt_scatter(pl->dtype, Dl, Dr, kl, P, (int)jlist.size(), jlist.data(),
                              L + (size_t)E * k0, (long long)wl * Dl, 0, Dl,
                              U, (long long)wl * kl, (long long)Dr * wl * kl, kl, 0, Cpeer, world, chunk, Dr,
                              (long long)chunk * Dr, 0, 1.0, 0.0, 1, stream))
      return 1;
  }
  // every rank contributes exactly one slab per pout (ranks with an empty range contribute none): the masks
  // are the same on all ranks because the ket ranges are (empty only for world > Dl, which create() rejects)
  unsigned int masks[8];
  for (int q = 0; q < 8; ++q) masks[q] = q < world ? 1u : 0u;
  if (comm_signal(comm, 0, stream)) return 1;
  if (comm_reduce_gather(comm, masks, P, 1, chunk, Dl, Dr, E, sign, stream)) return 1;
  if (comm_signal(comm, 1, stream)) return 1;
  if (comm_wait_copy(comm, y, (long long)P * Dl * Dr * E, stream)) return 1;
  return 0;
}

// ---------------------------------------------------------------------------------------
// Environment updates
// ---------------------------------------------------------------------------------------
namespace {

struct EnvLayout {
  size_t op_bytes, pack_a, pack_b, t1, v;   // byte sizes (256-aligned)
  size_t total() const { return op_bytes + pack_a + pack_b + t1 + v; }
};

EnvLayout env_layout(int dtype, int d, int Dl, int Dr, int wl, int wr) {
  const size_t eb = elem_bytes(dtype);
  EnvLayout l;
  // generous bound for the CSR blob: dense wl*wr*d*d entries
  const size_t nnz = (size_t)wl * wr * d * d;
  l.op_bytes = align_up(nnz * (8 + 16) + ((size_t)(wl > wr ? wl : wr) * d + 2) * 16 + 256, 256);
  l.pack_a = align_up((size_t)d * Dl * Dr * eb, 256);
  l.pack_b = l.pack_a;
  const int wmax = wl > wr ? wl : wr;
  l.t1 = align_up((size_t)d * Dl * Dr * wmax * eb, 256);
  l.v = l.t1;
  return l;
}

// W host tensor (w_a, w_b, d, d) element accessor
inline const double* w_at(const double* W, int E, int wb, int d, int a, int b, int i, int j) {
  return W + (size_t)E * ((((size_t)a * wb + b) * d + i) * d + j);
}

}  // namespace

extern "C" size_t pdmrg_env_workspace_bytes(int dtype, int d, int Dl, int Dr, int wl, int wr) {
  return env_layout(dtype, d, Dl, Dr, wl, wr).total();
}

extern "C" int pdmrg_env_update_right(int dtype, int d, int Dl, int Dr, int wl, int wr, const void* W_host,
                                      const void* L, const void* A, const void* Abra, void* out,
                                      void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "env_update_right: bad dtype");
  if (!W_host) PDMRG_REQUIRE(wl == wr, "env_update_right: identity MPO entry needs wl == wr");
  const EnvLayout lay = env_layout(dtype, d, Dl, Dr, wl, wr);
  PDMRG_REQUIRE(workspace_bytes >= lay.total(), "env_update_right: workspace too small");
  const int E = dtype == PDMRG_C128 ? 2 : 1;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  void* op_dev = ws;
  double* At = reinterpret_cast<double*>(ws + lay.op_bytes);
  double* AbT = reinterpret_cast<double*>(ws + lay.op_bytes + lay.pack_a);
  double* T1 = reinterpret_cast<double*>(ws + lay.op_bytes + lay.pack_a + lay.pack_b);
  double* V = reinterpret_cast<double*>(ws + lay.op_bytes + lay.pack_a + lay.pack_b + lay.t1);
  // At[n,q,p] = A[n,p,q]
  if (pdmrg_permute4(dtype, A, At, 1, d, Dr, Dl, 0, (long long)Dl * Dr, 1, Dr, 0, nullptr, 0, stream)) return 1;
  // AbT[k,i,j] = Abra[i,j,k]   (default bra = conj(A), tools/env_tools.py:41)
  if (pdmrg_permute4(dtype, Abra ? Abra : A, AbT, 1, Dr, d, Dl, 0, 1, (long long)Dl * Dr, Dr, Abra ? 0 : 1, nullptr, 0, stream)) return 1;
  // T1[(n,q),(j,l)] = sum_p At[(n,q),p] L[(j,l),p]
  if (pdmrg_gemm_nt(dtype, d * Dr, Dl * wl, Dl, 1, At, Dl, 0, L, Dl, 0, 0, T1, (long long)Dl * wl, 0, 1.0, 0.0, 0, stream)) return 1;
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
          h.in_off.push_back((long long)l + (long long)n * Dr * Dl * wl);
          h.val.push_back(re); h.val.push_back(im);
        }
      h.rowptr.push_back((int)h.in_off.size());
      h.out_off.push_back((long long)m * Dr * d * Dl + (long long)i * Dl);
    }
  PDMRG_REQUIRE(h.device_bytes() <= lay.op_bytes, "env_update_right: operator blob overflow");
  MixOpDev dev;
  if (upload_mix_op(h, op_dev, &dev, stream)) return 1;
  if (launch_mix(dtype, dev, T1, V, Dr, Dl, (long long)Dl * wl, wl, (long long)d * Dl, 1, stream)) return 1;
  // out[k,(m,q)] = sum_{(i,j)} AbT[k,(i,j)] V[(m,q),(i,j)]
  if (pdmrg_gemm_nt(dtype, Dr, wr * Dr, d * Dl, 1, AbT, (long long)d * Dl, 0, V, (long long)d * Dl, 0, 0, out,
                    (long long)wr * Dr, 0, 1.0, 0.0, 0, stream))
    return 1;
  return 0;
}

extern "C" int pdmrg_env_update_left(int dtype, int d, int Dl, int Dr, int wl, int wr, const void* W_host,
                                     const void* R, const void* B, const void* Bbra, void* out,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "env_update_left: bad dtype");
  if (!W_host) PDMRG_REQUIRE(wl == wr, "env_update_left: identity MPO entry needs wl == wr");
  const EnvLayout lay = env_layout(dtype, d, Dl, Dr, wl, wr);
  PDMRG_REQUIRE(workspace_bytes >= lay.total(), "env_update_left: workspace too small");
  const int E = dtype == PDMRG_C128 ? 2 : 1;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  void* op_dev = ws;
  double* BbT = reinterpret_cast<double*>(ws + lay.op_bytes);
  double* T1 = reinterpret_cast<double*>(ws + lay.op_bytes + lay.pack_a + lay.pack_b);
  double* V = reinterpret_cast<double*>(ws + lay.op_bytes + lay.pack_a + lay.pack_b + lay.t1);
  // BbT[x,b,c] = Bbra[b,x,c]
  if (pdmrg_permute4(dtype, Bbra ? Bbra : B, BbT, 1, Dl, d, Dr, 0, Dr, (long long)Dl * Dr, 1, Bbra ? 0 : 1, nullptr, 0, stream)) return 1;
  // T1[(e,a),(c,dd)] = sum_f B[(e,a),f] R[(c,dd),f]
  if (pdmrg_gemm_nt(dtype, d * Dl, Dr * wr, Dr, 1, B, Dr, 0, R, Dr, 0, 0, T1, (long long)Dr * wr, 0, 1.0, 0.0, 0, stream)) return 1;
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
          h.in_off.push_back((long long)dd + (long long)e * Dl * Dr * wr);
          h.val.push_back(re); h.val.push_back(im);
        }
      h.rowptr.push_back((int)h.in_off.size());
      h.out_off.push_back((long long)y * Dl * d * Dr + (long long)b * Dr);
    }
  PDMRG_REQUIRE(h.device_bytes() <= lay.op_bytes, "env_update_left: operator blob overflow");
  MixOpDev dev;
  if (upload_mix_op(h, op_dev, &dev, stream)) return 1;
  if (launch_mix(dtype, dev, T1, V, Dl, Dr, (long long)Dr * wr, wr, (long long)d * Dr, 1, stream)) return 1;
  // out[x,(y,a)] = sum_{(b,c)} BbT[x,(b,c)] V[(y,a),(b,c)]
  if (pdmrg_gemm_nt(dtype, Dl, wl * Dl, d * Dr, 1, BbT, (long long)d * Dr, 0, V, (long long)d * Dr, 0, 0, out,
                    (long long)wl * Dl, 0, 1.0, 0.0, 0, stream))
    return 1;
  return 0;
}
