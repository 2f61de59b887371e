// Internal (non-ABI) interfaces between the pydmrg_b200 CUDA translation units.
#pragma once
#include "common.cuh"
#include <vector>

namespace pdmrg {

// Sparse "mixing" operator applied point-wise over a 2-D index space (x, y):
//   out[row, x, y] = sum_{nz in row} val[nz] * in[col[nz], x, y]
// Rows / columns are composite (MPO-bond, physical) indices whose memory offsets are
// pre-resolved on the host, so one kernel serves every W-application of the hot path
// (tools/diag_tools.py:118,154-155; tools/env_tools.py:36,48).
struct MixOpHost {
  std::vector<int> rowptr;            // rows+1
  std::vector<long long> in_off;      // per nz: element offset of the input slab
  std::vector<double> val;            // per nz: (re, im) pairs (im ignored for f64)
  std::vector<long long> out_off;     // per row: element offset of the output slab
  int rows() const { return (int)rowptr.size() - 1; }
  int nnz() const { return (int)in_off.size(); }
  size_t device_bytes() const;
};

struct MixOpDev {
  const int* rowptr;
  const long long* in_off;
  const double* val;
  const long long* out_off;
  int rows;
};

// copies the operator into `dst` (device, >= device_bytes()) on `stream`
int upload_mix_op(const MixOpHost& h, void* dst, MixOpDev* dev, cudaStream_t stream);

// in/out strides (elements) of the point-wise index space; y is the fastest thread index
int launch_mix(int dtype, const MixOpDev& op, const void* in, void* out, int nx, int ny,
               long long six, long long siy, long long sox, long long soy, cudaStream_t stream);

// y[b,:] = sign * sum_q Z[b, idx[q], :] (+ y): ordered (deterministic) reduction of split-K slabs
int launch_reduce_slabs(int dtype, const void* Z, void* y, long long n_elems, int nb, int nq, const int* idx_host,
                        long long zstride_b, long long zstride_q, double sign, int accumulate, cudaStream_t stream);

}  // namespace pdmrg

// peer-memory communicator of the bond-sharded mat-vec (comm.cu)
struct pdmrg_comm;
namespace pdmrg {
int comm_signal(pdmrg_comm* c, int phase, cudaStream_t stream);
int comm_reduce_gather(pdmrg_comm* c, const unsigned int* jmask, int P, int wl, int chunk, int Dl, int Dr, int E, double sign,
                       cudaStream_t stream);
int comm_wait_copy(pdmrg_comm* c, void* y, long long n_doubles, cudaStream_t stream);
int comm_rank(const pdmrg_comm* c);
int comm_world(const pdmrg_comm* c);
void comm_next_epoch(pdmrg_comm* c);
double* comm_zpeer(const pdmrg_comm* c, int q);
}  // namespace pdmrg
