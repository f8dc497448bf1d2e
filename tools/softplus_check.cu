// Accuracy of bnn_softplus_fast / bnn_log_ge1 (sigma head of the Gaussian likelihood, k_fwd3) against libm on the device.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/softplus_check tools/softplus_check.cu && tools/softplus_check
#include <cmath>
#include <cstdio>
#include <vector>
#include "../npbnn_b200/csrc/bnn_common.cuh"

__global__ void k_check(const double* tab_g, int n, double* err_sp, double* err_log, double* err_pos_log) {
  __shared__ double tab[BNN_EXP_TAB_SIZE];
  for (int i = threadIdx.x; i < BNN_EXP_TAB_SIZE; i += blockDim.x) tab[i] = tab_g[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double z = -699.0 + 1398.0 * (double)i / (double)(n - 1);
  const double ref = fmax(z, 0.0) + log1p(exp(-fabs(z)));
  const double sp = bnn_softplus_fast(z, tab);
  err_sp[i] = fabs(sp - ref) / ref;
  const double lr = log(ref), lf = bnn_log_ge1(sp);
  err_log[i] = fabs(lf - lr) / fmax(fabs(lr), 1e-300);
  // log of arbitrary positive normals: 2^-1000 .. 2^1000
  const double x = ldexp(1.0 + (double)(i % 997) / 997.0, -1000 + (int)(2000.0 * i / n));
  err_pos_log[i] = fabs(bnn_log_ge1(x) - log(x)) / fmax(fabs(log(x)), 1e-300);
}

int main() {
  const int n = 1 << 20;
  std::vector<double> tab(BNN_EXP_TAB_SIZE);
  for (int j = 0; j < BNN_EXP_TAB_SIZE; ++j) tab[j] = exp2((double)j / BNN_EXP_TAB_SIZE);
  double *dt, *e1, *e2, *e3;
  cudaMalloc(&dt, sizeof(double) * BNN_EXP_TAB_SIZE);
  cudaMalloc(&e1, sizeof(double) * n); cudaMalloc(&e2, sizeof(double) * n); cudaMalloc(&e3, sizeof(double) * n);
  cudaMemcpy(dt, tab.data(), sizeof(double) * BNN_EXP_TAB_SIZE, cudaMemcpyHostToDevice);
  k_check<<<n / 256, 256>>>(dt, n, e1, e2, e3);
  std::vector<double> h1(n), h2(n), h3(n);
  cudaMemcpy(h1.data(), e1, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(h2.data(), e2, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(h3.data(), e3, sizeof(double) * n, cudaMemcpyDeviceToHost);
  if (cudaGetLastError() != cudaSuccess) { printf("cuda error\n"); return 1; }
  double m1 = 0, m2 = 0, m3 = 0;
  for (int i = 0; i < n; ++i) { m1 = fmax(m1, h1[i]); m2 = fmax(m2, h2[i]); m3 = fmax(m3, h3[i]); }
  printf("z in [-699, 699], %d points: max rel err softplus %.3e | log(softplus) %.3e (|log| < 1e-3 excluded: %s) | log(x), x in 2^[-1000,1000] %.3e\n",
         n, m1, m2, "no", m3);
  return (m1 < 1e-14 && m3 < 1e-14) ? 0 : 2;
}
