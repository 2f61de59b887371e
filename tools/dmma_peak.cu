// FP64 pipe micro-benchmark for B200 (sm_100a): DMMA.8x8x4 issue rate vs independent
// accumulators / resident warps, DFMA peak, and DMMA fed by LDS.64 fragment loads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_peak tools/dmma_peak.cu
// The numbers it prints are the measured FP64 denominators quoted in DESIGN.md.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = 0; c1[i] = 0; }
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}

// 64x32 warp tile: per k4 step 8 A frags + 4 B frags from shared memory, 32 DMMA
__global__ void __launch_bounds__(256) k_dmma_lds(double* out, int iters) {
  __shared__ double sA[128 * 20];
  __shared__ double sB[128 * 20];
  for (int i = threadIdx.x; i < 128 * 20; i += blockDim.x) { sA[i] = i * 1e-6; sB[i] = i * 2e-6; }
  __syncthreads();
  double c0[8][4], c1[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) { c0[i][j] = 0; c1[i][j] = 0; }
  int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  int w = threadIdx.x >> 5;
  const double* pa = sA + ((w & 1) * 64 + g) * 20 + t;
  const double* pb = sB + ((w >> 1) * 32 + g) * 20 + t;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = pa[i * 8 * 20 + kk * 4];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = pb[j * 8 * 20 + kk * 4];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma(c0[i][j], c1[i][j], a[i], b[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) s += c0[i][j] + c1[i][j];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps = 3) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

template <int NACC>
void run_dmma(double* out, int nsm, int warps, int ctas_per_sm) {
  int iters = 20000 / NACC;
  int threads = warps * 32;
  int grid = nsm * ctas_per_sm;
  float ms = time_ms([&] { k_dmma<NACC><<<grid, threads>>>(out, iters); });
  double flop = 2.0 * 256 * (double)NACC * iters * warps * grid;
  printf("dmma  nacc=%2d warps/cta=%2d ctas/sm=%d  %8.3f ms  %7.2f TFLOP/s\n", NACC, warps, ctas_per_sm, ms, flop / ms * 1e-9);
}

template <int NACC>
void run_dfma(double* out, int nsm, int warps) {
  int iters = 40000 / NACC;
  int grid = nsm;
  float ms = time_ms([&] { k_dfma<NACC><<<grid, warps * 32>>>(out, iters); });
  double flop = 2.0 * 32 * (double)NACC * iters * warps * grid;
  printf("dfma  nacc=%2d warps/cta=%2d            %8.3f ms  %7.2f TFLOP/s\n", NACC, warps, ms, flop / ms * 1e-9);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("device %s  SMs %d  clock %d kHz\n", p.name, nsm, p.clockRate);
  double* out; CK(cudaMalloc(&out, 1024));
  for (int warps : {4, 8, 16, 32}) {
    run_dmma<1>(out, nsm, warps, 1);
    run_dmma<2>(out, nsm, warps, 1);
    run_dmma<4>(out, nsm, warps, 1);
    run_dmma<8>(out, nsm, warps, 1);
    run_dmma<16>(out, nsm, warps, 1);
    run_dmma<32>(out, nsm, warps, 1);
  }
  for (int warps : {8, 16, 32}) { run_dfma<8>(out, nsm, warps); run_dfma<16>(out, nsm, warps); }
  for (int cps : {1, 2}) {
    int iters = 2000;
    float ms = time_ms([&] { k_dmma_lds<<<nsm * cps, 256>>>(out, iters); });
    double flop = 2.0 * 256 * 32 * 4.0 * iters * 8 * nsm * cps;
    printf("dmma+lds 64x32 warp tile, 8 warps, ctas/sm=%d  %8.3f ms  %7.2f TFLOP/s\n", cps, ms, flop / ms * 1e-9);
  }
  // sustained: 3 s of back-to-back DMMA
  {
    int iters = 20000 / 16;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 400; n++) k_dmma<16><<<nsm, 512>>>(out, iters * 10);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double flop = 2.0 * 256 * 16.0 * iters * 10 * 16 * nsm * n;
    printf("dmma sustained nacc=16 warps=16: %8.1f ms  %7.2f TFLOP/s\n", ms, flop / ms * 1e-9);
  }
  return 0;
}
