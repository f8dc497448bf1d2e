// Accuracy + SASS check of the fixed-point exponential (bnn_exp_split) and the activations built on it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp_fix_check tools/exp_fix_check.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "exp_fixed_point.cuh"

__global__ void k_eval(const double* __restrict__ x, int n, const double* __restrict__ gtab, double* o_exp, double* o_sw,
                       double* o_th, double* o_swf, double* o_thf) {
  __shared__ double tab[BNN_EXP_TAB_SIZE];
  for (int i = threadIdx.x; i < BNN_EXP_TAB_SIZE; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = x[i];
    o_exp[i] = bnn_exp_neg_fast(-fabs(v), tab);
    // reference columns: the library's table + polynomial activations need the UNSCALED table -> evaluated in double
    // precision library calls here (the host compares everything with long double anyway)
    o_sw[i] = v / (1.0 + exp(-v));
    o_th[i] = tanh(v);
    const bool ok1 = !expfix_needs_care<BNN_ACT_SWISH>(v), ok2 = !expfix_needs_care<BNN_ACT_TANH>(v);
    o_swf[i] = ok1 ? expfix_act_fast<BNN_ACT_SWISH>(v, tab) : o_sw[i];
    o_thf[i] = ok2 ? expfix_act_fast<BNN_ACT_TANH>(v, tab) : o_th[i];
  }
}

int main() {
  const int n = 1 << 22;
  std::vector<double> x(n), tab(BNN_EXP_TAB_SIZE);
  for (int j = 0; j < BNN_EXP_TAB_SIZE; ++j) tab[j] = ldexp(exp2((double)j / BNN_EXP_TAB_SIZE), -EXPFIX_TAB_BIAS);
  unsigned long long s = 88172645463325252ULL;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
  for (int i = 0; i < n; ++i) {
    const double u = rnd(), sc = (i % 4 == 0) ? 1.0 : (i % 4 == 1) ? 10.0 : (i % 4 == 2) ? 43.9 : 800.0;
    x[i] = (2.0 * u - 1.0) * sc;
  }
  x[0] = 0.0; x[1] = -0.0; x[2] = INFINITY; x[3] = -INFINITY; x[4] = NAN; x[5] = 44.0; x[6] = -44.0; x[7] = 43.999999;
  x[8] = 22.0; x[9] = -22.0; x[10] = 1e-300; x[11] = -1e-300; x[12] = 708.0; x[13] = -708.0; x[14] = 1e300; x[15] = -1e300;
  double *dx, *dt, *o[5];
  cudaMalloc(&dx, n * 8); cudaMalloc(&dt, BNN_EXP_TAB_SIZE * 8);
  for (auto& p : o) cudaMalloc(&p, n * 8);
  cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dt, tab.data(), BNN_EXP_TAB_SIZE * 8, cudaMemcpyHostToDevice);
  k_eval<<<296, 256>>>(dx, n, dt, o[0], o[1], o[2], o[3], o[4]);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
  std::vector<double> r[5];
  for (int k = 0; k < 5; ++k) { r[k].resize(n); cudaMemcpy(r[k].data(), o[k], n * 8, cudaMemcpyDeviceToHost); }
  const char* names[5] = {"exp(-|x|) fix", "swish libm", "tanh libm", "swish fix", "tanh fix"};
  for (int k = 0; k < 5; ++k) {
    double worst_rel = 0, worst_abs = 0, wx = 0, sum2 = 0; long cnt = 0;
    for (int i = 16; i < n; ++i) {
      const long double v = x[i];
      long double ref;
      if (k == 0) ref = expl(-fabsl(v));
      else if (k == 1 || k == 3) ref = v / (1.0L + expl(-v));
      else ref = tanhl(v);
      const double got = r[k][i];
      const double ab = (double)fabsl(got - ref);
      // exp: relative error where the result is not flushed; activations: absolute error relative to max(1, |z|)
      double rel;
      if (k == 0) { if (fabs(x[i]) >= 44.0) { if (got != 0.0) { printf("exp not flushed at %g\n", x[i]); } continue; } rel = ab / (double)ref; }
      else rel = ab / fmax(1.0, fabs(x[i]));
      sum2 += rel * rel; ++cnt;
      if (rel > worst_rel) { worst_rel = rel; wx = x[i]; }
      if (ab > worst_abs) worst_abs = ab;
    }
    printf("%-12s max err %.3e (at x = %.6g)  rms %.3e  max abs %.3e\n", names[k], worst_rel, wx, sqrt(sum2 / cnt), worst_abs);
  }
  printf("special values (x, exp(-|x|), swish, tanh):\n");
  for (int i = 0; i < 16; ++i) printf("  %-12g %-24.17g %-24.17g %-24.17g\n", x[i], r[0][i], r[1][i], r[2][i]);
  return 0;
}
