// Latency of anm::project_polygon in isolation (one warp = 4 lane groups), cycles per call, ANM6-like tables.
#include <cstdio>
#include <vector>
#include <array>
#include <cmath>
#include "../gym_anm_b200/csrc/anm_kernels.cuh"

struct Tab { std::vector<double> a, b, h, coef; std::vector<int> info; };
static Tab make(const std::vector<std::array<double, 3>>& rows) {
  Tab t; int R = (int)rows.size();
  t.a.assign(10, 0); t.b.assign(10, 0); t.h.assign(10, 1e300);
  for (int k = 0; k < R; ++k) { t.a[k] = rows[k][0]; t.b[k] = rows[k][1]; t.h[k] = rows[k][2]; }
  auto push = [&](int s1, int s2, unsigned need, std::array<double,4> kx, std::array<double,4> ky) {
    t.info.push_back(s1 | (s2 << 8) | (int)(need << 16));
    t.coef.insert(t.coef.end(), kx.begin(), kx.end()); t.coef.insert(t.coef.end(), ky.begin(), ky.end()); };
  push(0, 0, 0, {1,0,0,0}, {0,1,0,0});
  for (int k = 0; k < R; ++k) {
    double a = t.a[k], b = t.b[k];
    if (b == 0) push(k, k, 1u << k, {0,0,1/a,0}, {0,1,0,0});
    else if (a == 0) push(k, k, 1u << k, {1,0,0,0}, {0,0,1/b,0});
    else { double w = 1/(a*a+b*b); push(k, k, 1u << k, {1-a*a*w, -a*b*w, a*w, 0}, {-a*b*w, 1-b*b*w, b*w, 0}); }
  }
  for (int j = 1; j < R; ++j) for (int i = 0; i < j; ++i) {
    double det = t.a[i]*t.b[j] - t.a[j]*t.b[i]; if (det == 0) continue;
    push(i, j, (1u<<i)|(1u<<j), {0,0,t.b[j]/det,-t.b[i]/det}, {0,0,-t.a[j]/det,t.a[i]/det}); }
  return t;
}
__global__ void k(const double* ab, const double* h, const double* coef, const int* info, int ncand, double* out, long long* cyc, int reps) {
  __shared__ double s_ab[20], s_h[4][10], s_nh[4][10];
  __shared__ __align__(16) double s_coef[56 * 8];
  __shared__ int s_info[56];
  for (int i = threadIdx.x; i < 20; i += 32) s_ab[i] = ab[i];
  for (int i = threadIdx.x; i < 40; i += 32) s_h[i / 10][i % 10] = h[i % 10], s_nh[i / 10][i % 10] = -h[i % 10];
  for (int i = threadIdx.x; i < ncand * 8; i += 32) s_coef[i] = coef[i];
  for (int i = threadIdx.x; i < ncand; i += 32) s_info[i] = info[i];
  __syncwarp();
  const int lane = threadIdx.x % 8, grp = threadIdx.x / 8;
  const unsigned gm = 0xffu << (grp * 8);
  double p = 0.45 + 0.01 * grp, q = -0.6, po = 0, qo = 0, acc = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    anm::project_polygon<8, true>(s_ab, s_ab + 10, s_nh[grp], s_h[grp], 0x3ffu, reinterpret_cast<const double4*>(s_coef), s_info, ncand, ncand > 30 ? 10 : 7, p, q, lane, gm, po, qo);
    p = po + 0.3; q = qo - 0.2; acc += po + qo;   // dependent calls
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / reps;
}
int main() {
  Tab g = make({{-1,0,0},{1,0,0.3},{1,0,0.25},{0,-1,0.3},{0,1,0.3},{1.5,1,0.6},{1.5,-1,0.6}});
  Tab s = make({{-1,0,0.5},{1,0,0.5},{0,-1,0.5},{0,1,0.5},{1.25,1,0.875},{1.25,-1,0.875},{-1.25,-1,0.875},{-1.25,1,0.875},{-1,0,0.1},{1,0,0.2}});
  for (Tab* t : {&g, &s}) {
    double *ab, *h, *coef, *out; int* info; long long* cyc;
    std::vector<double> abv(t->a); abv.insert(abv.end(), t->b.begin(), t->b.end());
    cudaMalloc(&ab, 160); cudaMalloc(&h, 80); cudaMalloc(&coef, t->coef.size() * 8); cudaMalloc(&info, t->info.size() * 4);
    cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
    cudaMemcpy(ab, abv.data(), 160, cudaMemcpyHostToDevice); cudaMemcpy(h, t->h.data(), 80, cudaMemcpyHostToDevice);
    cudaMemcpy(coef, t->coef.data(), t->coef.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(info, t->info.data(), t->info.size() * 4, cudaMemcpyHostToDevice);
    for (int it = 0; it < 2; ++it) k<<<1, 32>>>(ab, h, coef, info, (int)t->info.size(), out, cyc, 200);
    long long c; double o[32];
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(o, out, 256, cudaMemcpyDeviceToHost);
    printf("ncand=%d: %lld cycles per projection (%.1f per round), checksum %.6f  err=%s\n", (int)t->info.size(), c,
           (double)c / ((t->info.size() + 7) / 8), o[0], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
