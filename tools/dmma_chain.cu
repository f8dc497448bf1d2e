// How many independent accumulator chains does a warp need to keep the FP64 tensor pipe of B200 busy, and what does
// a dependent FP64 chain (an activation) cost when it shares the pipe?  12 warps per SM as in k_fwd3.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_chain tools/dmma_chain.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// NA independent accumulators; per outer iteration every accumulator gets one m16n8k8 (4 dependent-pair DMMA.884);
// NCH independent FP64 chains of CL dependent FMAs are interleaved per iteration (0 = none)
template <int NA, int NCH, int CL>
__global__ void __launch_bounds__(384, 1) k_chain(double* out, int iters, double x, double y) {
  double c[NA][4];
#pragma unroll
  for (int i = 0; i < NA; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  double a[4], b[2], f[NCH > 0 ? NCH : 1];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = x + threadIdx.x * 1e-6 + i * 1e-3;
  b[0] = y + threadIdx.x * 1e-6; b[1] = y * 0.5;
#pragma unroll
  for (int i = 0; i < (NCH > 0 ? NCH : 1); ++i) f[i] = i + threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NA; ++i) mma1688(c[i], a, b);
    if (NCH > 0) {
#pragma unroll
      for (int s = 0; s < CL; ++s)
#pragma unroll
        for (int q = 0; q < NCH; ++q) f[q] = fma(f[q], x, y);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
#pragma unroll
  for (int i = 0; i < (NCH > 0 ? NCH : 1); ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NA, int NCH, int CL>
void run(int sms, double* out) {
  const int iters = 8192 / NA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_chain<NA, NCH, CL><<<sms, 384>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k_chain<NA, NCH, CL><<<sms, 384>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  const double warps = sms * 12.0, mma = (double)iters * NA * warps, fp = (double)iters * NCH * CL * warps;
  const double clk = ms * 1e-3 * 1.965e9;                       // per SMSP: 3 warps
  const double per_smsp_mma = (double)iters * NA * 3, per_smsp_fp = (double)iters * NCH * CL * 3;
  printf("acc=%d chains=%dx%-2d  %.3f ms  dmma %.2f TFLOP/s  | clk per m16n8k8 if FP64 instr were free: %.1f"
         "  | clk per FP64 instr beyond 64/mma: %.2f\n",
         NA, NCH, CL, ms, 2.0 * 1024 * mma / (ms * 1e-3) / 1e12, clk / per_smsp_mma,
         per_smsp_fp > 0 ? (clk - 64.0 * per_smsp_mma) / per_smsp_fp : 0.0);
  (void)fp;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 384);
  printf("%s, %d SMs, 12 warps/SM\n", prop.name, sms);
  run<1, 0, 0>(sms, out); run<2, 0, 0>(sms, out); run<4, 0, 0>(sms, out); run<8, 0, 0>(sms, out);
  // activation-like chains between MMA groups: 4 mma (layer 2 k-group) + 4 chains x 12 dependent FMAs
  run<4, 4, 12>(sms, out); run<4, 8, 12>(sms, out); run<4, 4, 24>(sms, out);
  run<8, 4, 12>(sms, out); run<8, 8, 12>(sms, out); run<2, 4, 12>(sms, out);
  run<8, 4, 6>(sms, out); run<8, 16, 3>(sms, out); run<8, 48, 1>(sms, out);
  return 0;
}
