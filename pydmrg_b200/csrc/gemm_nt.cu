// FP64 / complex128 "NT" GEMM for sm_100a:  C = alpha * A * op(B)^T + beta * C
//
// Blackwell's tcgen05 tensor path has no f64 kind, so the FP64 tensor pipe is reached with
// warp-level mma.sync.m8n8k4.f64 (SASS: DMMA.8x8x4; measured 36.9 TFLOP/s on B200,
// profiles/dmma_peak_r01.txt).  Operand tiles are staged by TMA (cp.async.bulk.tensor,
// 128-byte swizzle) into a 6-stage shared-memory ring guarded by mbarriers; one producer
// warp issues the copies, eight consumer warps (2 x 4, 64x32 warp tiles) issue DMMA.
//
// Complex128 runs on the same real tiles through the real embedding of B only:
//   C_re = sum_k  Ar*Br - Ai*Bi ,  C_im = sum_k Ar*Bi + Ai*Br
// i.e. C_real (M x 2N) = A_real (M x 2K, the interleaved array as it lies in memory)
//                        * B'(2N x 2K)^T,  B'[2n] = [Br,-Bi,...],  B'[2n+1] = [Bi,Br,...].
// B' is never materialised: the B fragment loader reads the swapped element of the plain
// interleaved tile and flips a sign bit (one LOP3), so a complex GEMM costs exactly
// 4 real DMMA products and writes interleaved complex C directly.
//
// This is the tile engine under every contraction of the reference's hot path
// (pydmrg/tools/diag_tools.py:114-119,153-156; pydmrg/tools/env_tools.py:35-37,47-49).
#include "common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace pdmrg {
namespace {

constexpr int BM = 128;          // rows of A per CTA tile
constexpr int BNR = 128;         // REAL columns of C per CTA tile (64 complex columns)
constexpr int BK = 16;           // doubles per k-iteration (= one 128-byte swizzle row)
constexpr int STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int A_BYTES = BM * BK * 8;           // 16 KiB
constexpr int STAGE_BYTES = 2 * A_BYTES;       // B tile lives at +A_BYTES (8 or 16 KiB used)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8;

struct GemmParams {
  int M, N;                 // N = columns of C in ELEMENTS (complex columns when CPLX)
  int kiters, batch, tiles_m, tiles_n;
  long long ldc, strideC;   // elements
  double* C;
  double alpha_re, alpha_im;
  int beta, conjB, bcastA, bcastB;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <bool CPLX>
__global__ void __launch_bounds__(THREADS, 1)
gemm_nt_dmma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  // 128-byte swizzle needs 1024-byte aligned tiles (no static shared memory in this kernel)
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars);
  const uint32_t empty0 = smem_u32(bars + STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int total_tiles = p.tiles_m * p.tiles_n * p.batch;
  constexpr int B_ROWS = CPLX ? BNR / 2 : BNR;
  constexpr uint32_t TX_BYTES = A_BYTES + B_ROWS * BK * 8;

  if (warp == CONSUMER_WARPS) {
    // ------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tm = tile % p.tiles_m;
        const int rest = tile / p.tiles_m;
        const int tn = rest % p.tiles_n;
        const int b = rest / p.tiles_n;
        const int rowA = tm * BM, rowB = tn * B_ROWS;
        const int bA = p.bcastA ? 0 : b, bB = p.bcastB ? 0 : b;
        for (int kit = 0; kit < p.kiters; ++kit) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t full = full0 + 8 * stage;
          mbar_expect_tx(full, TX_BYTES);
          const uint32_t sA = smem_base + stage * STAGE_BYTES;
          tma_load_3d(sA, &tmA, kit * BK, rowA, bA, full);
          tma_load_3d(sA + A_BYTES, &tmB, kit * BK, rowB, bB, full);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // -------------------------------- DMMA consumers ----------------------------------
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;
  // fragment row g of an 8-row MMA tile maps to tile row perm(g): the two half-warps of a
  // 64-bit shared load then touch rows {0,2,4,6} / {1,3,5,7}, which the 128B swizzle
  // spreads over all banks (conflict-free).
  const int permg = ((g & 3) << 1) | (g >> 2);
  const uint32_t offA = (wm * 64 + permg) * 128 + ((t & 1) << 3) + ((((t >> 1) ^ permg) & 7) << 4);
  uint32_t offB;
  uint32_t signmask = 0;
  if (CPLX) {
    const int part = g & 1, jc = g >> 1;
    offB = (wn * 16 + jc * 2) * 128 + (((t & 1) ^ part) << 3) + ((((t >> 1) ^ (jc << 1)) & 7) << 4);
    const bool neg = p.conjB ? (part == 1 && (t & 1) == 0) : (part == 0 && (t & 1) == 1);
    signmask = neg ? 0x80000000u : 0u;
  } else {
    offB = (wn * 32 + permg) * 128 + ((t & 1) << 3) + ((((t >> 1) ^ permg) & 7) << 4);
  }

  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tm = tile % p.tiles_m;
    const int rest = tile / p.tiles_m;
    const int tn = rest % p.tiles_n;
    const int b = rest / p.tiles_n;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    for (int kit = 0; kit < p.kiters; ++kit) {
      mbar_wait(full0 + 8 * stage, phase);
      const unsigned char* sA = smem + stage * STAGE_BYTES;
      const unsigned char* sB = sA + A_BYTES;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        double a[8], bf[4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          a[i] = *reinterpret_cast<const double*>(sA + (offA ^ (kk << 5)) + i * 1024);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (CPLX) {
            const uint32_t o = (offB ^ ((kk << 5) | ((j & 1) << 4))) + ((j & 1) + (j >> 1) * 8) * 128;
            double v = *reinterpret_cast<const double*>(sB + o);
            bf[j] = __hiloint2double(__double2hiint(v) ^ signmask, __double2loint(v));
          } else {
            bf[j] = *reinterpret_cast<const double*>(sB + (offB ^ (kk << 5)) + j * 1024);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * stage);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }

    // ------------------------------ epilogue ----------------------------------------
    double* Cb = p.C + (CPLX ? 2 : 1) * (long long)b * p.strideC;
    const int row0 = tm * BM + wm * 64 + permg;
    if (CPLX) {
      const int col0 = tn * (BNR / 2) + wn * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = row0 + i * 8;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = col0 + t * 2 + (j & 1) + (j >> 1) * 8;
          if (col >= p.N) continue;
          double2* dst = reinterpret_cast<double2*>(Cb + 2 * ((long long)row * p.ldc + col));
          const double re = acc[i][j][0], im = acc[i][j][1];
          double2 out = make_double2(p.alpha_re * re - p.alpha_im * im, p.alpha_re * im + p.alpha_im * re);
          if (p.beta) { const double2 old = *dst; out.x += old.x; out.y += old.y; }
          *dst = out;
        }
      }
    } else {
      const int col0 = tn * BNR + wn * 32;
      const int g0 = 2 * t, g1 = 2 * t + 1;
      const int pc0 = ((g0 & 3) << 1) | (g0 >> 2), pc1 = ((g1 & 3) << 1) | (g1 >> 2);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = row0 + i * 8;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double* dst = Cb + (long long)row * p.ldc + col0 + j * 8;
          if (col0 + j * 8 + pc0 < p.N) {
            double v = p.alpha_re * acc[i][j][0];
            if (p.beta) v += dst[pc0];
            dst[pc0] = v;
          }
          if (col0 + j * 8 + pc1 < p.N) {
            double v = p.alpha_re * acc[i][j][1];
            if (p.beta) v += dst[pc1];
            dst[pc1] = v;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Generic kernel: any alignment / leading dimension, small problems (bond dimensions 1,2,4..
// near the chain ends).  Plain DFMA, 32x32 tile per 16x16 thread block.
// ---------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256)
gemm_nt_generic_kernel(int M, int N, int K, const double* __restrict__ A, long long lda, long long strideA,
                       const double* __restrict__ B, long long ldb, long long strideB, int conjB,
                       double* __restrict__ C, long long ldc, long long strideC,
                       double alpha_re, double alpha_im, int beta) {
  constexpr int E = CPLX ? 2 : 1;
  __shared__ double sA[32][16 * E + 1];
  __shared__ double sB[32][16 * E + 1];
  const int b = blockIdx.z;
  const double* Ab = A + E * (long long)b * strideA;
  const double* Bb = B + E * (long long)b * strideB;
  double* Cb = C + E * (long long)b * strideC;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  double accr[2][2] = {{0, 0}, {0, 0}}, acci[2][2] = {{0, 0}, {0, 0}};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int idx = threadIdx.x; idx < 32 * 16; idx += 256) {
      const int r = idx >> 4, c = idx & 15;
      const int k = k0 + c;
      for (int e = 0; e < E; ++e) {
        sA[r][c * E + e] = (m0 + r < M && k < K) ? Ab[E * ((long long)(m0 + r) * lda + k) + e] : 0.0;
        sB[r][c * E + e] = (n0 + r < N && k < K) ? Bb[E * ((long long)(n0 + r) * ldb + k) + e] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 16; ++c) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int r = ty + 16 * i, q = tx + 16 * j;
          if (CPLX) {
            const double ar = sA[r][2 * c], ai = sA[r][2 * c + 1];
            const double br = sB[q][2 * c], bi = conjB ? -sB[q][2 * c + 1] : sB[q][2 * c + 1];
            accr[i][j] = fma(ar, br, accr[i][j]); accr[i][j] = fma(-ai, bi, accr[i][j]);
            acci[i][j] = fma(ar, bi, acci[i][j]); acci[i][j] = fma(ai, br, acci[i][j]);
          } else {
            accr[i][j] = fma(sA[r][c], sB[q][c], accr[i][j]);
          }
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int row = m0 + ty + 16 * i, col = n0 + tx + 16 * j;
      if (row >= M || col >= N) continue;
      double* dst = Cb + E * ((long long)row * ldc + col);
      if (CPLX) {
        double re = alpha_re * accr[i][j] - alpha_im * acci[i][j];
        double im = alpha_re * acci[i][j] + alpha_im * accr[i][j];
        if (beta) { re += dst[0]; im += dst[1]; }
        dst[0] = re; dst[1] = im;
      } else {
        double v = alpha_re * accr[i][j];
        if (beta) v += dst[0];
        dst[0] = v;
      }
    }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// 3-D map over a row-major matrix viewed as REAL doubles: dims {kreal, rows, batch}
int make_map(CUtensorMap* tm, const void* base, long long kreal, long long rows, long long ld_real,
             long long batch, long long stride_real, int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
  cuuint64_t dims[3] = {(cuuint64_t)kreal, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld_real * 8,
                           (cuuint64_t)((batch > 1 && stride_real > 0) ? stride_real : rows * ld_real) * 8};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return 1; }
  return 0;
}

std::atomic<int> g_last_path{0};
std::atomic<int> g_force_generic{0};

}  // namespace
}  // namespace pdmrg

using namespace pdmrg;

extern "C" int pdmrg_last_gemm_path(void) { return g_last_path.load(); }
extern "C" void pdmrg_set_force_generic_gemm(int on) { g_force_generic.store(on); }

extern "C" int pdmrg_gemm_nt(int dtype, int M, int N, int K, int batch, const void* A, long long lda,
                             long long strideA, const void* B, long long ldb, long long strideB, int conjB,
                             void* C, long long ldc, long long strideC, double alpha_re, double alpha_im,
                             int beta, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "pdmrg_gemm_nt: bad dtype");
  PDMRG_REQUIRE(beta == 0 || beta == 1, "pdmrg_gemm_nt: beta must be 0 or 1");
  PDMRG_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "pdmrg_gemm_nt: negative extent");
  if (M == 0 || N == 0 || batch == 0) return 0;
  const bool cplx = dtype == PDMRG_C128;
  if (!cplx) { PDMRG_REQUIRE(alpha_im == 0.0, "pdmrg_gemm_nt: complex alpha with real dtype"); }
  const int E = cplx ? 2 : 1;

  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool strides_ok = ((lda * E) % 2 == 0) && ((ldb * E) % 2 == 0) && ((strideA * E) % 2 == 0) &&
                          ((strideB * E) % 2 == 0);
  const double work = (double)M * N * (double)(K > 0 ? K : 1) * batch * (cplx ? 4 : 1);
  const bool use_tma = !g_force_generic.load() && K > 0 && aligned16(A) && aligned16(B) && strides_ok &&
                       (!cplx || aligned16(C)) && work >= 64.0 * 64.0 * 64.0;

  if (use_tma) {
    static bool attr_set = false;
    if (!attr_set) {
      PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_set = true;
    }
    CUtensorMap tmA, tmB;
    const long long kreal = (long long)K * E;
    if (make_map(&tmA, A, kreal, M, lda * E, strideA == 0 ? 1 : batch, strideA * E, BM)) return 1;
    if (make_map(&tmB, B, kreal, N, ldb * E, strideB == 0 ? 1 : batch, strideB * E, cplx ? BNR / 2 : BNR)) return 1;
    GemmParams p;
    p.M = M; p.N = N;
    p.kiters = (int)((kreal + BK - 1) / BK);
    p.batch = batch;
    p.tiles_m = (M + BM - 1) / BM;
    const int bn = cplx ? BNR / 2 : BNR;
    p.tiles_n = (N + bn - 1) / bn;
    p.ldc = ldc; p.strideC = strideC;
    p.C = static_cast<double*>(C);
    p.alpha_re = alpha_re; p.alpha_im = alpha_im;
    p.beta = beta; p.conjB = conjB;
    p.bcastA = (strideA == 0 || batch == 1); p.bcastB = (strideB == 0 || batch == 1);
    const long long total = (long long)p.tiles_m * p.tiles_n * batch;
    const int grid = (int)(total < num_sms() ? total : num_sms());
    if (cplx)
      gemm_nt_dmma_kernel<true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
    else
      gemm_nt_dmma_kernel<false><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
    PDMRG_CHECK_LAUNCH();
    count_launch();
    g_last_path.store(1);
    return 0;
  }

  dim3 grid((N + 31) / 32, (M + 31) / 32, batch);
  PDMRG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "pdmrg_gemm_nt: generic path extent too large");
  if (cplx)
    gemm_nt_generic_kernel<true><<<grid, 256, 0, stream>>>(M, N, K, (const double*)A, lda, strideA, (const double*)B, ldb,
                                                           strideB, conjB, (double*)C, ldc, strideC, alpha_re, alpha_im, beta);
  else
    gemm_nt_generic_kernel<false><<<grid, 256, 0, stream>>>(M, N, K, (const double*)A, lda, strideA, (const double*)B, ldb,
                                                            strideB, 0, (double*)C, ldc, strideC, alpha_re, 0.0, beta);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  g_last_path.store(2);
  return 0;
}
