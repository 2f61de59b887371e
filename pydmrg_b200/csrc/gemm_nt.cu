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

#include <map>
#include <mutex>

namespace pdmrg {
namespace {

constexpr int BM_MAX = 128;      // rows of A per CTA tile: 128 (8 m-tiles per warp) or 64 (4)
constexpr int BNR = 128;         // REAL columns of C per CTA tile (64 complex columns)
constexpr int BK = 16;           // doubles per k-iteration (= one 128-byte swizzle row)
constexpr int STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int A_BYTES = BM_MAX * BK * 8;       // 16 KiB (the 64-row variant uses the first half)
constexpr int STAGE_BYTES = 2 * A_BYTES;       // B tile lives at +A_BYTES (8 or 16 KiB used)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8;

constexpr int MAX_IDX2 = 32;

// Two batch levels: b = b1 * nb2 + b2; the second level may be an index list (idx2[b2]) so that
// e.g. only the non-empty MPO-bond slabs are visited, in ONE launch (no per-slab wave tails).
struct GemmParams {
  int M, N;                 // N = columns of C in ELEMENTS (complex columns when CPLX)
  int kiters, nb1, nb2, tiles_m, tiles_n;
  long long ldc, strideC1, strideC2;   // elements
  double* C;
  double alpha_re, alpha_im;
  int beta, conjB;
  int bcastA1, bcastA2, bcastB1, bcastB2;
  int idx2[MAX_IDX2];
  // SCATTER epilogue: rows [q*rows_per_owner, (q+1)*rows_per_owner) of C go to Cpeer[q] (a peer GPU's
  // memory mapped through CUDA IPC) -- the reduce-scatter of a bond-sharded mat-vec rides on the
  // GEMM epilogue, tile by tile, as plain st.global over NVLink
  double* Cpeer[8];
  int rows_per_owner;
  // accum2: the nb2 entries of the second batch level are SUMMED into one output tile (the k-loop runs over
  // all of them) instead of producing nb2 separate outputs -- a long-K product whose K is not contiguous
  int accum2;
  // stream-K (small problems: fewer tiles than ~4 waves): the global iteration space (tiles x k-iterations) is cut into
  // gridDim.x equal contiguous ranges; a tile split between CTAs is finished by the CTA that holds its FIRST k-segment
  // (its last work item), which adds the fragment-layout partials the other holders left in `sk_partial` (fixed order:
  // deterministic).  sk_flags[c] == sk_epoch announces the partial of CTA c.
  int streamk;
  long long sk_per;          // k-iterations per CTA
  double* sk_partial;        // gridDim.x x (MI*8*256) doubles
  unsigned int* sk_flags;
  unsigned int sk_epoch;
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Work items of one CTA: (tile, [k0, k1)) in processing order.  Data-parallel mode: whole tiles, strided over the grid.
// Stream-K mode: CTA c owns the contiguous range r = gridDim.x-1-c of the global iteration space (reversed, so that the
// finisher of a split tile only ever waits for CTAs with a LOWER index, which the hardware dispatches first).
template <bool SK>
struct WorkIter {
  static constexpr bool mode = SK;
  int total_tiles, ktot, tile, step;
  long long it, it_end;
  __device__ WorkIter(const GemmParams& p, int total_tiles_, int ktot_) : total_tiles(total_tiles_), ktot(ktot_) {
    if (mode) {
      const long long total = (long long)total_tiles * ktot;
      const long long r = (long long)gridDim.x - 1 - blockIdx.x;
      it = r * p.sk_per;
      it_end = it + p.sk_per < total ? it + p.sk_per : total;
    } else {
      tile = blockIdx.x;
      step = gridDim.x;
    }
  }
  __device__ bool next(int& t, int& k0, int& k1) {
    if (mode) {
      if (it >= it_end) return false;
      t = (int)(it / ktot);
      k0 = (int)(it - (long long)t * ktot);
      const long long room = it_end - it;
      k1 = (long long)(ktot - k0) <= room ? ktot : (int)(k0 + room);
      it += k1 - k0;
      return true;
    }
    if (tile >= total_tiles) return false;
    t = tile; k0 = 0; k1 = ktot;
    tile += step;
    return true;
  }
};

template <bool CPLX, int MI, bool SCATTER = false, bool SK = false>
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

  constexpr int BM = 16 * MI;                       // 128 or 64 rows per CTA tile
  const int nb2_out = p.accum2 ? 1 : p.nb2;            // outputs along the second batch level
  const int nb2_k = p.accum2 ? p.nb2 : 1;              // second-level entries folded into the k-loop
  const int total_tiles = p.tiles_m * p.tiles_n * p.nb1 * nb2_out;
  constexpr int B_ROWS = CPLX ? BNR / 2 : BNR;
  constexpr uint32_t TX_BYTES = BM * BK * 8 + B_ROWS * BK * 8;

  if (warp == CONSUMER_WARPS) {
    // ------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int ktot_p = p.kiters * nb2_k;
      WorkIter<SK> wi(p, total_tiles, ktot_p);
      int tile, k0, k1;
      while (wi.next(tile, k0, k1)) {
        const int tm = tile % p.tiles_m;
        const int rest = tile / p.tiles_m;
        const int tn = rest % p.tiles_n;
        const int b = rest / p.tiles_n;
        const int rowA = tm * BM, rowB = tn * B_ROWS;
        const int b1 = b / nb2_out;
        const int a1 = p.bcastA1 ? 0 : b1, c1 = p.bcastB1 ? 0 : b1;
        int q = SK ? k0 / p.kiters : 0, kit = SK ? k0 - q * p.kiters : 0;       // data-parallel segments start at k = 0
        for (int kglob = k0; kglob < k1; ++kglob, ++kit) {
          if (kit == p.kiters) { kit = 0; ++q; }
          const int b2 = p.idx2[p.accum2 ? q : b % p.nb2];
          const int a2 = p.bcastA2 ? 0 : b2, c2 = p.bcastB2 ? 0 : b2;
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t full = full0 + 8 * stage;
          mbar_expect_tx(full, TX_BYTES);
          const uint32_t sA = smem_base + stage * STAGE_BYTES;
          tma_load_4d(sA, &tmA, kit * BK, rowA, a2, a1, full);
          tma_load_4d(sA + A_BYTES, &tmB, kit * BK, rowB, c2, c1, full);
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
  const uint32_t offA = (wm * (8 * MI) + permg) * 128 + ((t & 1) << 3) + ((((t >> 1) ^ permg) & 7) << 4);
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
  const int ktot = p.kiters * nb2_k;
  WorkIter<SK> wi(p, total_tiles, ktot);
  int tile, seg_k0, seg_k1;
  while (wi.next(tile, seg_k0, seg_k1)) {
    const int tm = tile % p.tiles_m;
    const int rest = tile / p.tiles_m;
    const int tn = rest % p.tiles_n;
    const int b = rest / p.tiles_n;

    double acc[MI][4][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    for (int kit = seg_k0; kit < seg_k1; ++kit) {
      mbar_wait(full0 + 8 * stage, phase);
      const unsigned char* sA = smem + stage * STAGE_BYTES;
      const unsigned char* sB = sA + A_BYTES;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        double a[MI], bf[4];
#pragma unroll
        for (int i = 0; i < MI; ++i)
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
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
      }
      // The stage may only be handed back once every lane's LDS of it has been PERFORMED, not merely
      // issued: ptxas turns the converged __syncwarp into a NOP and schedules the arrive right behind the
      // last LDS, and under load/store-unit back-pressure (peer stores of sibling warps in the scatter
      // epilogue) the arrive was observed to overtake them -- TMA then refilled rows still being read.
      __threadfence_block();
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * stage);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }

    if (SK && !(seg_k0 == 0 && seg_k1 == ktot)) {
      constexpr int FRAG = MI * 8;                               // accumulator doubles per thread
      const int ctid = threadIdx.x;                              // consumer thread id, 0..255
      if (seg_k0 != 0) {
        // END / middle segment: leave the partial in fragment layout and announce it
        double* dst = p.sk_partial + (size_t)blockIdx.x * (FRAG * 256) + ctid;
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __stcg(dst + ((i * 4 + j) * 2 + 0) * 256, acc[i][j][0]);
            __stcg(dst + ((i * 4 + j) * 2 + 1) * 256, acc[i][j][1]);
          }
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ctid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sk_flags + blockIdx.x), "r"(p.sk_epoch) : "memory");
        continue;
      }
      // START segment: this CTA finishes the tile.  The remaining segments [seg_k1, ktot) belong to the ranges
      // r+1, r+2, ... = CTAs blockIdx.x-1, blockIdx.x-2, ... (all dispatched before this one); add them in order.
      int have = seg_k1;
      for (int c = (int)blockIdx.x - 1; have < ktot; --c) {
        if (ctid == 0) {
          unsigned int v;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.sk_flags + c) : "memory");
          } while (v != p.sk_epoch);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const double* src = p.sk_partial + (size_t)c * (FRAG * 256) + ctid;
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[i][j][0] += __ldcg(src + ((i * 4 + j) * 2 + 0) * 256);
            acc[i][j][1] += __ldcg(src + ((i * 4 + j) * 2 + 1) * 256);
          }
        have += (int)((long long)(ktot - have) < p.sk_per ? (ktot - have) : p.sk_per);
      }
    }

    // ------------------------------ epilogue ----------------------------------------
    const long long boff = (CPLX ? 2 : 1) * ((long long)(b / nb2_out) * p.strideC1 +
                                             (p.accum2 ? 0ll : (long long)p.idx2[b % p.nb2] * p.strideC2));
    double* Cb = p.C + boff;
    const int row0 = tm * BM + wm * (8 * MI) + permg;
    if (CPLX) {
      const int col0 = tn * (BNR / 2) + wn * 16;
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int row = row0 + i * 8;
        if (row >= p.M) continue;
        double* Crow;
        if (SCATTER) {
          const int owner = row / p.rows_per_owner;
          Crow = p.Cpeer[owner] + boff + 2 * (long long)(row - owner * p.rows_per_owner) * p.ldc;
        } else {
          Crow = Cb + 2 * (long long)row * p.ldc;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = col0 + t * 2 + (j & 1) + (j >> 1) * 8;
          if (col >= p.N) continue;
          double2* dst = reinterpret_cast<double2*>(Crow + 2 * col);
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
      for (int i = 0; i < MI; ++i) {
        const int row = row0 + i * 8;
        if (row >= p.M) continue;
        double* Crow;
        if (SCATTER) {
          const int owner = row / p.rows_per_owner;
          Crow = p.Cpeer[owner] + boff + (long long)(row - owner * p.rows_per_owner) * p.ldc;
        } else {
          Crow = Cb + (long long)row * p.ldc;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double* dst = Crow + col0 + j * 8;
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
    // peer stores of this tile must be visible system-wide before the kernel retires: the flag that
    // releases them to the owner rank is written by a later, stream-ordered kernel
    if (SCATTER) __threadfence_system();
  }
}

// ---------------------------------------------------------------------------------------
// Generic kernel: any alignment / leading dimension, small problems (bond dimensions 1,2,4..
// near the chain ends).  Plain DFMA, 32x32 tile per 16x16 thread block.
// ---------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256)
gemm_nt_generic_kernel(int M, int N, int K, const double* __restrict__ A, long long lda, long long strideA1, long long strideA2,
                       const double* __restrict__ B, long long ldb, long long strideB1, long long strideB2, int conjB,
                       double* __restrict__ C, long long ldc, long long strideC1, long long strideC2,
                       double alpha_re, double alpha_im, int beta, int nb2, GemmParams idx) {
  constexpr int E = CPLX ? 2 : 1;
  __shared__ double sA[32][16 * E + 1];
  __shared__ double sB[32][16 * E + 1];
  const long long b1 = blockIdx.z / nb2, b2 = idx.idx2[blockIdx.z % nb2];
  const double* Ab = A + E * (b1 * strideA1 + b2 * strideA2);
  const double* Bb = B + E * (b1 * strideB1 + b2 * strideB2);
  double* Cb = C + E * (b1 * strideC1 + b2 * strideC2);
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

// 4-D map over a row-major matrix viewed as REAL doubles: dims {kreal, rows, n2, n1}
int make_map(CUtensorMap* tm, const void* base, long long kreal, long long rows, long long ld_real,
             long long n2, long long stride2_real, long long n1, long long stride1_real, int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
  if (n2 < 1) n2 = 1;
  if (n1 < 1) n1 = 1;
  const long long dflt = rows * ld_real;          // any valid (16-byte multiple) stride for extent-1 dims
  cuuint64_t dims[4] = {(cuuint64_t)kreal, (cuuint64_t)rows, (cuuint64_t)n2, (cuuint64_t)n1};
  cuuint64_t strides[3] = {(cuuint64_t)ld_real * 8, (cuuint64_t)((n2 > 1 && stride2_real > 0) ? stride2_real : dflt) * 8,
                           (cuuint64_t)((n1 > 1 && stride1_real > 0) ? stride1_real : dflt) * 8};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return 1; }
  return 0;
}

std::atomic<int> g_last_path{0};
std::atomic<int> g_force_generic{0};
std::atomic<int> g_force_bm{0};

// wave-quantisation cost model: time ~ waves * tile height; the 64-row tile pays a small
// efficiency penalty (more B traffic per FLOP) but halves the tail of a partial wave
int choose_bm(long long M, long long tiles_n, long long nbatch) {
  if (g_force_bm.load() == 64 || g_force_bm.load() == 128) return g_force_bm.load();
  const long long sms = num_sms();
  const long long t128 = ((M + 127) / 128) * tiles_n * nbatch, t64 = ((M + 63) / 64) * tiles_n * nbatch;
  const double c128 = (double)((t128 + sms - 1) / sms) * 1.0;
  const double c64 = (double)((t64 + sms - 1) / sms) * 0.5 * 1.04;
  return c64 < c128 ? 64 : 128;
}

// ---- stream-K scratch: one set of partial-tile buffers + flags per CUDA stream (launches on one stream are ordered;
// concurrent streams must not share partials).  Allocated on first use, kept for the life of the process.
struct StreamKScratch {
  double* partial = nullptr;
  unsigned int* flags = nullptr;
  unsigned int epoch = 0;
};
std::mutex g_sk_mutex;
std::map<std::pair<int, cudaStream_t>, StreamKScratch> g_sk_scratch;
std::atomic<int> g_streamk_mode{1};     // 0 = never, 1 = cost model, 2 = always (tests)
std::atomic<int> g_last_streamk{0};

int streamk_scratch(cudaStream_t stream, StreamKScratch* out) {
  int dev = 0;
  PDMRG_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_sk_mutex);
  StreamKScratch& sc = g_sk_scratch[std::make_pair(dev, stream)];
  if (!sc.partial) {
    const size_t ctas = (size_t)num_sms();
    PDMRG_CHECK_CUDA(cudaMalloc(&sc.partial, ctas * 8 * 8 * 256 * sizeof(double)));
    PDMRG_CHECK_CUDA(cudaMalloc(&sc.flags, ctas * sizeof(unsigned int)));
    PDMRG_CHECK_CUDA(cudaMemset(sc.flags, 0, ctas * sizeof(unsigned int)));      // legacy-stream memset: ordered before any launch
    PDMRG_CHECK_CUDA(cudaDeviceSynchronize());
  }
  sc.epoch += 1;
  if (sc.epoch == 0) sc.epoch = 1;
  *out = sc;
  return 0;
}

}  // namespace
}  // namespace pdmrg

using namespace pdmrg;

extern "C" int pdmrg_last_gemm_path(void) { return g_last_path.load(); }
extern "C" void pdmrg_set_streamk_mode(int mode) { g_streamk_mode.store(mode); }
extern "C" int pdmrg_last_gemm_streamk(void) { return g_last_streamk.load(); }
extern "C" void pdmrg_set_force_generic_gemm(int on) { g_force_generic.store(on); }
extern "C" void pdmrg_set_force_tile_rows(int bm) { g_force_bm.store(bm); }

static int gemm_nt_impl(int dtype, int M, int N, int K, int nb1, int nb2, const int* idx2_host,
                        const void* A, long long lda, long long strideA1, long long strideA2,
                        const void* B, long long ldb, long long strideB1, long long strideB2, int conjB,
                        void* C, long long ldc, long long strideC1, long long strideC2,
                        double alpha_re, double alpha_im, int beta, void* stream_,
                        const void* const* Cpeer, int n_owner, int rows_per_owner, int accum2);

extern "C" int pdmrg_gemm_nt_batched2(int dtype, int M, int N, int K, int nb1, int nb2, const int* idx2_host,
                                      const void* A, long long lda, long long strideA1, long long strideA2,
                                      const void* B, long long ldb, long long strideB1, long long strideB2, int conjB,
                                      void* C, long long ldc, long long strideC1, long long strideC2,
                                      double alpha_re, double alpha_im, int beta, void* stream_) {
  return gemm_nt_impl(dtype, M, N, K, nb1, nb2, idx2_host, A, lda, strideA1, strideA2, B, ldb, strideB1, strideB2, conjB, C, ldc,
                      strideC1, strideC2, alpha_re, alpha_im, beta, stream_, nullptr, 0, 0, 0);
}

extern "C" int pdmrg_gemm_nt_scatter(int dtype, int M, int N, int K, int nb1, int nb2, const int* idx2_host,
                                     const void* A, long long lda, long long strideA1, long long strideA2,
                                     const void* B, long long ldb, long long strideB1, long long strideB2, int conjB,
                                     const void* const* Cpeer_host, int n_owner, int rows_per_owner,
                                     long long ldc, long long strideC1, long long strideC2,
                                     double alpha_re, double alpha_im, int accumulate_b2, void* stream_) {
  PDMRG_REQUIRE(n_owner >= 1 && n_owner <= 8 && rows_per_owner > 0, "pdmrg_gemm_nt_scatter: bad owner layout");
  PDMRG_REQUIRE((long long)n_owner * rows_per_owner >= M, "pdmrg_gemm_nt_scatter: owners do not cover the rows");
  return gemm_nt_impl(dtype, M, N, K, nb1, nb2, idx2_host, A, lda, strideA1, strideA2, B, ldb, strideB1, strideB2, conjB,
                      const_cast<void*>(Cpeer_host[0]), ldc, strideC1, strideC2, alpha_re, alpha_im, 0, stream_, Cpeer_host, n_owner,
                      rows_per_owner, accumulate_b2 ? 1 : 0);
}

static int gemm_nt_impl(int dtype, int M, int N, int K, int nb1, int nb2, const int* idx2_host,
                        const void* A, long long lda, long long strideA1, long long strideA2,
                        const void* B, long long ldb, long long strideB1, long long strideB2, int conjB,
                        void* C, long long ldc, long long strideC1, long long strideC2,
                        double alpha_re, double alpha_im, int beta, void* stream_,
                        const void* const* Cpeer, int n_owner, int rows_per_owner, int accum2) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PDMRG_REQUIRE(dtype == PDMRG_F64 || dtype == PDMRG_C128, "pdmrg_gemm_nt: bad dtype");
  PDMRG_REQUIRE(beta == 0 || beta == 1, "pdmrg_gemm_nt: beta must be 0 or 1");
  PDMRG_REQUIRE(M >= 0 && N >= 0 && K >= 0 && nb1 >= 0 && nb2 >= 0, "pdmrg_gemm_nt: negative extent");
  PDMRG_REQUIRE(nb2 <= MAX_IDX2, "pdmrg_gemm_nt: second batch level limited to 32 entries");
  if (M == 0 || N == 0 || nb1 == 0 || nb2 == 0) return 0;
  const bool cplx = dtype == PDMRG_C128;
  if (!cplx) { PDMRG_REQUIRE(alpha_im == 0.0, "pdmrg_gemm_nt: complex alpha with real dtype"); }
  const int E = cplx ? 2 : 1;

  GemmParams p;
  int max2 = 0;
  for (int i = 0; i < MAX_IDX2; ++i) p.idx2[i] = 0;
  for (int i = 0; i < nb2; ++i) {
    p.idx2[i] = idx2_host ? idx2_host[i] : i;
    PDMRG_REQUIRE(p.idx2[i] >= 0, "pdmrg_gemm_nt: negative batch index");
    if (p.idx2[i] > max2) max2 = p.idx2[i];
  }
  const long long nbatch = (long long)nb1 * (accum2 ? 1 : nb2);       // output batches

  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  auto even = [&](long long v) { return (v * E) % 2 == 0; };
  const bool strides_ok = even(lda) && even(ldb) && even(strideA1) && even(strideA2) && even(strideB1) && even(strideB2);
  const double work = (double)M * N * (double)(K > 0 ? K : 1) * (double)nb1 * nb2 * (cplx ? 4 : 1);
  const bool use_tma = !g_force_generic.load() && K > 0 && aligned16(A) && aligned16(B) && strides_ok &&
                       (!cplx || aligned16(C)) && work >= 64.0 * 64.0 * 64.0;

  p.M = M; p.N = N;
  p.nb1 = nb1; p.nb2 = nb2;
  p.ldc = ldc; p.strideC1 = strideC1; p.strideC2 = strideC2;
  p.C = static_cast<double*>(C);
  p.alpha_re = alpha_re; p.alpha_im = alpha_im;
  p.beta = beta; p.conjB = conjB;
  p.bcastA1 = (strideA1 == 0 || nb1 == 1); p.bcastA2 = (strideA2 == 0 || max2 == 0);
  p.bcastB1 = (strideB1 == 0 || nb1 == 1); p.bcastB2 = (strideB2 == 0 || max2 == 0);
  for (int i = 0; i < 8; ++i) p.Cpeer[i] = (Cpeer && i < n_owner) ? static_cast<double*>(const_cast<void*>(Cpeer[i])) : nullptr;
  p.rows_per_owner = rows_per_owner > 0 ? rows_per_owner : 1;
  p.accum2 = accum2;
  const bool scatter = Cpeer != nullptr;
  if (scatter) PDMRG_REQUIRE(beta == 0, "pdmrg_gemm_nt_scatter: beta must be 0");

  if (use_tma) {
    static bool attr_set = false;
    if (!attr_set) {
      PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_set = true;
    }
    const long long kreal = (long long)K * E;
    const int bn = cplx ? BNR / 2 : BNR;
    p.kiters = (int)((kreal + BK - 1) / BK);
    p.tiles_n = (N + bn - 1) / bn;
    int bm = choose_bm(M, p.tiles_n, nbatch);
    // stream-K when whole tiles quantise badly over the SMs (bond dimensions <= 512: a fraction of a wave to a few
    // waves): cost in k-iterations of the busiest SM, data-parallel = ceil(tiles / SMs) * ktot (x 0.52 for the 64-row
    // tile), stream-K = ceil(tiles * ktot / SMs) + ~4 for the partial store / fix-up (measured: ~10 us)
    p.streamk = 0; p.sk_per = 0; p.sk_partial = nullptr; p.sk_flags = nullptr; p.sk_epoch = 0;
    int grid_sk = 0;
    {
      const long long sms = num_sms();
      const long long ktot = (long long)p.kiters * (accum2 ? nb2 : 1);
      const long long t128 = ((M + 127) / 128) * (long long)p.tiles_n * nbatch, t64 = ((M + 63) / 64) * (long long)p.tiles_n * nbatch;
      const double dp = bm == 128 ? (double)((t128 + sms - 1) / sms) * ktot : (double)((t64 + sms - 1) / sms) * ktot * 0.52;
      const long long iters = t128 * ktot;
      const double sk = (double)((iters + sms - 1) / sms) + 4.0;
      const int mode = g_streamk_mode.load();
      if (!scatter && mode != 0 && iters >= 8 && ktot >= 2 && (mode == 2 || (sk < 0.95 * dp && t128 < 6 * sms))) {
        long long g = sms < iters / 4 ? sms : iters / 4;
        if (g < 1) g = 1;
        const long long per = (iters + g - 1) / g;
        g = (iters + per - 1) / per;
        StreamKScratch sc;
        if (streamk_scratch(stream, &sc)) return 1;
        bm = 128;
        p.streamk = 1; p.sk_per = per; p.sk_partial = sc.partial; p.sk_flags = sc.flags; p.sk_epoch = sc.epoch;
        grid_sk = (int)g;
      }
    }
    g_last_streamk.store(p.streamk);
    p.tiles_m = (M + bm - 1) / bm;
    CUtensorMap tmA, tmB;
    if (make_map(&tmA, A, kreal, M, lda * E, p.bcastA2 ? 1 : max2 + 1, strideA2 * E, p.bcastA1 ? 1 : nb1, strideA1 * E, bm)) return 1;
    if (make_map(&tmB, B, kreal, N, ldb * E, p.bcastB2 ? 1 : max2 + 1, strideB2 * E, p.bcastB1 ? 1 : nb1, strideB1 * E, bn)) return 1;
    const long long total = (long long)p.tiles_m * p.tiles_n * nbatch;
    PDMRG_REQUIRE(total < (1ll << 31), "pdmrg_gemm_nt: too many tiles");
    const int grid = p.streamk ? grid_sk : (int)(total < num_sms() ? total : num_sms());
    if (scatter) {
      static bool attr2 = false;
      if (!attr2) {
        PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<true, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<false, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<true, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<false, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr2 = true;
      }
      if (bm == 128) {
        if (cplx) gemm_nt_dmma_kernel<true, 8, true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
        else gemm_nt_dmma_kernel<false, 8, true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
      } else {
        if (cplx) gemm_nt_dmma_kernel<true, 4, true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
        else gemm_nt_dmma_kernel<false, 4, true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
      }
    } else if (p.streamk) {
      static bool attr3 = false;
      if (!attr3) {
        PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<true, 8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        PDMRG_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_dmma_kernel<false, 8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr3 = true;
      }
      if (cplx) gemm_nt_dmma_kernel<true, 8, false, true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
      else gemm_nt_dmma_kernel<false, 8, false, true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
    } else if (bm == 128) {
      if (cplx) gemm_nt_dmma_kernel<true, 8><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
      else gemm_nt_dmma_kernel<false, 8><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
    } else {
      if (cplx) gemm_nt_dmma_kernel<true, 4><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
      else gemm_nt_dmma_kernel<false, 4><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
    }
    PDMRG_CHECK_LAUNCH();
    count_launch();
    g_last_path.store(1);
    return 0;
  }

  PDMRG_REQUIRE(!scatter && !accum2, "pdmrg_gemm_nt_scatter: operands must be TMA-compatible (16-byte aligned, even strides)");
  dim3 grid((N + 31) / 32, (M + 31) / 32, (unsigned)nbatch);
  PDMRG_REQUIRE(grid.y <= 65535 && nbatch <= 65535, "pdmrg_gemm_nt: generic path extent too large");
  p.kiters = 0; p.tiles_m = p.tiles_n = 0;
  p.streamk = 0; p.sk_per = 0; p.sk_partial = nullptr; p.sk_flags = nullptr; p.sk_epoch = 0;
  g_last_streamk.store(0);
  if (cplx)
    gemm_nt_generic_kernel<true><<<grid, 256, 0, stream>>>(M, N, K, (const double*)A, lda, strideA1, strideA2, (const double*)B, ldb,
                                                           strideB1, strideB2, conjB, (double*)C, ldc, strideC1, strideC2,
                                                           alpha_re, alpha_im, beta, nb2, p);
  else
    gemm_nt_generic_kernel<false><<<grid, 256, 0, stream>>>(M, N, K, (const double*)A, lda, strideA1, strideA2, (const double*)B, ldb,
                                                            strideB1, strideB2, 0, (double*)C, ldc, strideC1, strideC2,
                                                            alpha_re, 0.0, beta, nb2, p);
  PDMRG_CHECK_LAUNCH();
  count_launch();
  g_last_path.store(2);
  return 0;
}

extern "C" int pdmrg_gemm_nt(int dtype, int M, int N, int K, int batch, const void* A, long long lda,
                             long long strideA, const void* B, long long ldb, long long strideB, int conjB,
                             void* C, long long ldc, long long strideC, double alpha_re, double alpha_im,
                             int beta, void* stream_) {
  return pdmrg_gemm_nt_batched2(dtype, M, N, K, batch, 1, nullptr, A, lda, strideA, 0, B, ldb, strideB, 0, conjB, C, ldc,
                                strideC, 0, alpha_re, alpha_im, beta, stream_);
}
