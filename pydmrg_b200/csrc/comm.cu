// Peer-memory plumbing for the MPO-bond-sharded mat-vec: one process per GPU,
// buffers exchanged through CUDA IPC, data moved by the kernels themselves with plain
// ld/st.global on peer-mapped pointers over NVLink (no NCCL on this path):
//
//   stage-C GEMM (gemm_nt.cu, SCATTER epilogue)  -- every output tile is stored straight into the
//        memory of the rank that owns those rows of y (reduce-scatter riding on the epilogue);
//   signal                                         -- release: flag[rank] = epoch on every peer;
//   reduce_gather_kernel                           -- acquire the flags, sum the slabs all ranks
//        delivered for this rank's rows, and PUSH the reduced rows to every rank's y staging
//        buffer (all-gather by peer stores);
//   signal, wait_copy_kernel                       -- acquire, copy the staged y to the caller.
//
// Every rank ends up with a bit-identical y (each row chunk is reduced by exactly one owner).
#include "ops.cuh"

struct pdmrg_comm {
  int rank, world;
  double* Zpeer[8];
  double* Ypeer[8];
  unsigned long long* Fpeer[8];
  unsigned long long epoch;
};

namespace pdmrg {
namespace {

struct PeerPtrs { double* p[8]; };
struct FlagPtrs { unsigned long long* p[8]; };
struct JMasks { unsigned int m[8]; };

__global__ void signal_kernel(FlagPtrs f, int world, int slot, unsigned long long epoch) {
  if ((int)threadIdx.x < world) {
    __threadfence_system();                       // order this GPU's earlier stores (incl. peer stores)
    volatile unsigned long long* dst = f.p[threadIdx.x] + slot;
    *dst = epoch;
  }
}

// Bounded: a peer that never signals (a rank that died, or -- it must not happen -- took a different control decision)
// would otherwise hang this GPU for good.  After ~2^23 polls (several seconds) the wait gives up and raises the sticky
// error word behind the flags (slot 2*world), which the host checks (pdmrg_comm_error).
__device__ __forceinline__ void wait_flags(unsigned long long* flags, int world, int phase, unsigned long long epoch) {
  if ((int)threadIdx.x < world) {
    const volatile unsigned long long* src = flags + phase * world + threadIdx.x;
    unsigned int polls = 0;
    while (*src < epoch) {
      __nanosleep(400);
      if (++polls > (1u << 23)) { flags[2 * world] = epoch; break; }
    }
    __threadfence_system();
  }
  __syncthreads();
}

// Zloc[src][pout][j][il][r]  ->  Ypeer[q][pout][rank*chunk + il][r]  for every q
template <int V>   // V = doubles per access: 2 (16-byte loads and peer stores; rowlen even) or 1
__global__ void __launch_bounds__(256)
reduce_gather_kernel(const double* __restrict__ Zloc, PeerPtrs Y, unsigned long long* __restrict__ flags,
                     int world, int rank, unsigned long long epoch, JMasks jm, int P, int wl, int chunk, int rows_here,
                     long long Dl, long long rowlen /* doubles per row */, double sign) {
  wait_flags(flags, world, 0, epoch);
  const long long per_p = (long long)rows_here * rowlen;
  const long long total = (long long)P * per_p / V;
  const long long slab = (long long)chunk * rowlen;             // one (src, pout, j) slab
  for (long long tv = blockIdx.x * (long long)blockDim.x + threadIdx.x; tv < total; tv += (long long)gridDim.x * blockDim.x) {
    const long long t = tv * V;
    const long long po = t / per_p, e = t - po * per_p;         // e = il * rowlen + r
    double acc0 = 0.0, acc1 = 0.0;
    for (int src = 0; src < world; ++src) {
      const double* z = Zloc + ((long long)src * P + po) * wl * slab + e;
      unsigned int m = jm.m[src];
      for (int j = 0; m; ++j, m >>= 1)
        if (m & 1u) {
          if (V == 2) { const double2 v = *reinterpret_cast<const double2*>(z + (long long)j * slab); acc0 += v.x; acc1 += v.y; }
          else acc0 += z[(long long)j * slab];
        }
    }
    acc0 *= sign; acc1 *= sign;
    const long long dst = po * Dl * rowlen + (long long)rank * chunk * rowlen + e;
    for (int q = 0; q < world; ++q) {                           // all-gather by peer stores (16 bytes each when V == 2)
      if (V == 2) *reinterpret_cast<double2*>(Y.p[q] + dst) = make_double2(acc0, acc1);
      else Y.p[q][dst] = acc0;
    }
  }
  __threadfence_system();      // the pushes are released to the peers by the flag kernel that follows
}

__global__ void __launch_bounds__(256)
wait_copy_kernel(const double* __restrict__ Yloc, double* __restrict__ y, long long n, unsigned long long* __restrict__ flags,
                 int world, unsigned long long epoch) {
  wait_flags(flags, world, 1, epoch);
  if ((n & 1) == 0 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(Yloc)) & 15) == 0) {
    const double2* s2 = reinterpret_cast<const double2*>(Yloc);
    double2* d2 = reinterpret_cast<double2*>(y);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n / 2; t += (long long)gridDim.x * blockDim.x) d2[t] = s2[t];
    return;
  }
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) y[t] = Yloc[t];
}

}  // namespace

int comm_signal(pdmrg_comm* c, int phase, cudaStream_t stream) {
  FlagPtrs f;
  for (int i = 0; i < 8; ++i) f.p[i] = c->Fpeer[i];
  signal_kernel<<<1, 32, 0, stream>>>(f, c->world, phase * c->world + c->rank, c->epoch);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  return 0;
}

int comm_reduce_gather(pdmrg_comm* c, const unsigned int* jmask, int P, int wl, int chunk, int Dl, int Dr, int E, double sign,
                       cudaStream_t stream) {
  PeerPtrs Y; JMasks jm;
  for (int i = 0; i < 8; ++i) { Y.p[i] = c->Ypeer[i]; jm.m[i] = i < c->world ? jmask[i] : 0u; }
  int rows_here = Dl - c->rank * chunk;
  if (rows_here > chunk) rows_here = chunk;
  if (rows_here < 0) rows_here = 0;
  const long long rowlen = (long long)Dr * E;
  const bool vec2 = (rowlen % 2) == 0;
  long long total = (long long)P * rows_here * rowlen / (vec2 ? 2 : 1);
  int grid = (int)((total + 255) / 256);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  if (grid < 1) grid = 1;
  if (vec2)
    reduce_gather_kernel<2><<<grid, 256, 0, stream>>>(c->Zpeer[c->rank], Y, c->Fpeer[c->rank], c->world, c->rank, c->epoch, jm, P, wl,
                                                      chunk, rows_here, Dl, rowlen, sign);
  else
    reduce_gather_kernel<1><<<grid, 256, 0, stream>>>(c->Zpeer[c->rank], Y, c->Fpeer[c->rank], c->world, c->rank, c->epoch, jm, P, wl,
                                                      chunk, rows_here, Dl, rowlen, sign);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  return 0;
}

int comm_wait_copy(pdmrg_comm* c, void* y, long long n_doubles, cudaStream_t stream) {
  int grid = (int)((n_doubles + 255) / 256);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  wait_copy_kernel<<<grid, 256, 0, stream>>>(c->Ypeer[c->rank], (double*)y, n_doubles, c->Fpeer[c->rank], c->world, c->epoch);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  return 0;
}

int comm_rank(const pdmrg_comm* c) { return c->rank; }
int comm_world(const pdmrg_comm* c) { return c->world; }
void comm_next_epoch(pdmrg_comm* c) { ++c->epoch; }
double* comm_zpeer(const pdmrg_comm* c, int q) { return c->Zpeer[q]; }

}  // namespace pdmrg

using namespace pdmrg;

extern "C" int pdmrg_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
  PDMRG_REQUIRE(sizeof(cudaIpcMemHandle_t) == 64, "unexpected cudaIpcMemHandle_t size");
  void* p = nullptr;
  PDMRG_CHECK_CUDA(cudaMalloc(&p, bytes));
  PDMRG_CHECK_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  PDMRG_CHECK_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}

extern "C" int pdmrg_ipc_open(const unsigned char* handle64, void** dev_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  PDMRG_CHECK_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int pdmrg_ipc_close(void* dev_ptr) {
  PDMRG_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

extern "C" int pdmrg_ipc_free(void* dev_ptr) {
  PDMRG_CHECK_CUDA(cudaFree(dev_ptr));
  return 0;
}

extern "C" int pdmrg_comm_create(pdmrg_comm** out, int rank, int world, void* const* Zpeer, void* const* Ypeer,
                                 void* const* Fpeer) {
  PDMRG_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "pdmrg_comm_create: bad rank / world (<= 8 GPUs)");
  auto* c = new pdmrg_comm();
  c->rank = rank; c->world = world; c->epoch = 0;
  for (int i = 0; i < 8; ++i) {
    c->Zpeer[i] = i < world ? static_cast<double*>(Zpeer[i]) : nullptr;
    c->Ypeer[i] = i < world ? static_cast<double*>(Ypeer[i]) : nullptr;
    c->Fpeer[i] = i < world ? static_cast<unsigned long long*>(Fpeer[i]) : nullptr;
  }
  *out = c;
  return 0;
}

extern "C" void pdmrg_comm_destroy(pdmrg_comm* c) { delete c; }

// Sticky error word of this rank's flag buffer: non-zero = a flag wait timed out (value = the epoch that was waited for).
// Synchronises `stream`.
extern "C" int pdmrg_comm_error(pdmrg_comm* c, unsigned long long* err_out, void* stream_) {
  PDMRG_REQUIRE(c && err_out, "pdmrg_comm_error: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_CHECK_CUDA(cudaMemcpyAsync(err_out, c->Fpeer[c->rank] + 2 * c->world, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  PDMRG_CHECK_CUDA(cudaStreamSynchronize(stream));
  return 0;
}
