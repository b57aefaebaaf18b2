// Dependent-issue latencies on the GPU at hand (cycles per op in a dependent chain, one warp).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP>
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
  __shared__ double sm[64];
  double x = a + threadIdx.x, y = b;
  sm[threadIdx.x] = x;
  sm[threadIdx.x + 32] = y;
  __syncwarp();
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) {
    if (OP == 0) x = fma(x, y, a);
    if (OP == 1) x = x + y;
    if (OP == 2) x = x * y;
    if (OP == 3) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    if (OP == 4) { idx = (int)sm[idx & 31] & 31; x += idx; }            // LDS -> F2I -> addr (upper bound)
    if (OP == 5) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x)); }
    if (OP == 6) x = x / y;
    if (OP == 7) x = sqrt(x + 2.0);
    if (OP == 8) { float f = (float)x; f = fmaf(f, 1.0001f, 0.5f); x = f; }
    if (OP == 9) { x = (x > y) ? x - y : x + a; }                         // DSETP + select + DADD
    if (OP == 10) { int4 v = *reinterpret_cast<int4*>(&sm[(idx & 15) * 2]); idx = v.x & 15; }  // LDS.128 pointer chase
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + idx;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP>
void run(const char* name) {
  double* out; long long* cyc;
  cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
  k<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, N);
  k<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, N);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %7.1f cycles/op\n", name, (double)h / N);
}
// throughput: independent chains in one warp
__global__ void thr(double* out, long long* cyc, double a, int n) {
  double x[8];
  for (int j = 0; j < 8; ++j) x[j] = a + j + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fma(x[j], 0.9999999, a);
  }
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  run<0>("DFMA dependent"); run<1>("DADD dependent"); run<2>("DMUL dependent"); run<3>("SHFL.64 dependent");
  run<4>("LDS.64 + F2I chase"); run<5>("MUFU.RCP64H dependent"); run<6>("fp64 divide dependent"); run<7>("fp64 sqrt dependent");
  run<8>("F2F+FFMA+F2F roundtrip"); run<9>("DSETP+SEL+DADD"); run<10>("LDS.128 pointer chase");
  double* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
  thr<<<1, 32>>>(out, cyc, 1.0, 1024); thr<<<1, 32>>>(out, cyc, 1.0, 1024);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %7.2f cycles per warp-DFMA (8 independent chains, 1 warp)\n", "DFMA issue rate", (double)h / (1024 * 8));
  return 0;
}
