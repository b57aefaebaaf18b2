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
    if (OP == 11) { float f = (float)x; f = fmaf(f, 0.9999f, 0.25f); f = fmaf(f, 0.9999f, 0.25f); f = fmaf(f, 0.9999f, 0.25f);
                    f = fmaf(f, 0.9999f, 0.25f); x = f; }                    // 4 dependent FFMA (+ 2 conversions)
    if (OP == 12) {  // one level of the tree elimination in fp64: det -> 1/det -> T = adj(D) [L | f] -> C = U T / det
      const double d00 = x, d01 = y, d10 = a, d11 = x + 2.0, l0 = y, l1 = a, u0 = a, u1 = y;
      const double det = d00 * d11 - d01 * d10;
      double r;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(det));
      r = fma(fma(-det, r, 1.0), r, r);
      r = fma(fma(-det, r, 1.0), r, r);
      const double t0_ = d11 * l0 - d01 * l1, t1_ = d00 * l1 - d10 * l0;
      x = x - (u0 * t0_ + u1 * t1_) * r;  // folded into the parent's block
    }
    if (OP == 13 || OP == 14) {  // the same level in fp32 (MUFU.RCP + one Newton step); OP 14: twice (solve + refinement
                                 // solve) with an fp64 residual in between -- what an fp64-accurate mixed step needs
      float xf = (float)x;
      const float yf = (float)y, af = (float)a;
#pragma unroll
      for (int rep = 0; rep < (OP == 14 ? 2 : 1); ++rep) {
        const float d00 = xf, d01 = yf, d10 = af, d11 = xf + 2.0f, l0 = yf, l1 = af, u0 = af, u1 = yf;
        const float det = d00 * d11 - d01 * d10;
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(det));
        r = fmaf(fmaf(-det, r, 1.0f), r, r);
        const float t0_ = d11 * l0 - d01 * l1, t1_ = d00 * l1 - d10 * l0;
        xf = xf - (u0 * t0_ + u1 * t1_) * r;
        if (OP == 14 && rep == 0) {  // fp64 residual of the 2x2 system: r = f - D x (4 dependent DFMA-class operations)
          const double xd = xf;
          double res = fma(-x, xd, y);
          res = fma(-a, xd, res);
          xf = (float)res + xf;
        }
      }
      x = xf;
    }
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
  run<11>("4 x FFMA dependent (+2 F2F)");
  run<12>("tree level fp64 (det,rcp,fold)"); run<13>("tree level fp32"); run<14>("tree level fp32 x2 + fp64 residual");
  double* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
  thr<<<1, 32>>>(out, cyc, 1.0, 1024); thr<<<1, 32>>>(out, cyc, 1.0, 1024);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %7.2f cycles per warp-DFMA (8 independent chains, 1 warp)\n", "DFMA issue rate", (double)h / (1024 * 8));
  return 0;
}
