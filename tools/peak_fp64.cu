// FP64 peak micro-benchmarks for B200 (sm_100a).
//
// MEASURED_PEAKS.json carries HBM and bf16 peaks only; the npBNN hot path is bound by the
// FP64 pipe, so the roofline denominator has to be measured here:
//   * DFMA         : vector fused multiply-add, 8 independent chains per thread
//   * DMMA m8n8k4  : mma.sync.aligned.m8n8k4.f64   (sm_80+)
//   * DMMA m16n8k4 / k8 / k16 : mma.sync.aligned.m16n8k{4,8,16}.f64 (sm_90+)
//   * DMMA + DFMA interleaved : do the two share one pipe?
//   * layout probe : checks the fragment <-> matrix coordinate mapping the kernels assume
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peak_fp64 tools/peak_fp64.cu
// Run  :  tools/peak_fp64 [out.json]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
               "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

constexpr int NACC = 8;

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double x, double y) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>   // 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double x, double y) {
  double c2[NACC][2];
  double c4[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    c2[i][0] = c2[i][1] = 0.0;
    c4[i][0] = c4[i][1] = c4[i][2] = c4[i][3] = 0.0;
  }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = x + threadIdx.x * 1e-6 + i * 1e-3;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = y + threadIdx.x * 1e-6 + i * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (KIND == 0) mma884(c2[i], a[0], b[0]);
        if (KIND == 1) { double aa[2] = {a[0], a[1]}; mma1684(c4[i], aa, b[0]); }
        if (KIND == 2) { double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; mma1688(c4[i], aa, bb); }
        if (KIND == 3) mma16816(c4[i], a, b);
      }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c2[i][0] + c2[i][1] + c4[i][0] + c4[i][1] + c4[i][2] + c4[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA (m16n8k8) with NF independent DFMA per thread per MMA interleaved: shared pipe or not?
template <int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double x, double y) {
  double c4[NACC][4];
  double f[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c4[i][0] = c4[i][1] = c4[i][2] = c4[i][3] = 0.0; f[i] = i + threadIdx.x * 1e-9; }
  double a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = x + threadIdx.x * 1e-6 + i * 1e-3;
  b[0] = y; b[1] = y + 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        mma1688(c4[i], a, b);
#pragma unroll
        for (int q = 0; q < NF; ++q) f[(i + q) % NACC] = fma(f[(i + q) % NACC], x, y);
      }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c4[i][0] + c4[i][1] + c4[i][2] + c4[i][3] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- layout probe: C[16x8] = A[16x8] * B[8x8] with the K-permuted fragment mapping the BNN kernels use:
//   a0=A[g][2t] a1=A[g+8][2t] a2=A[g][2t+1] a3=A[g+8][2t+1];  b0=Bt[g][2t] b1=Bt[g][2t+1]  (Bt[n][k], "W[out][in]")
//   c0=C[g][2t] c1=C[g][2t+1] c2=C[g+8][2t] c3=C[g+8][2t+1]
__global__ void k_probe(const double* A, const double* Bt, double* C88, double* C84a, double* C884) {
  int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  {
    double a[4] = {A[g * 8 + 2 * t], A[(g + 8) * 8 + 2 * t], A[g * 8 + 2 * t + 1], A[(g + 8) * 8 + 2 * t + 1]};
    double b[2] = {Bt[g * 8 + 2 * t], Bt[g * 8 + 2 * t + 1]};
    double c[4] = {0, 0, 0, 0};
    mma1688(c, a, b);
    C88[g * 8 + 2 * t] = c[0]; C88[g * 8 + 2 * t + 1] = c[1];
    C88[(g + 8) * 8 + 2 * t] = c[2]; C88[(g + 8) * 8 + 2 * t + 1] = c[3];
  }
  {  // same product as two m16n8k4 steps
    double c[4] = {0, 0, 0, 0};
    double a0[2] = {A[g * 8 + 2 * t], A[(g + 8) * 8 + 2 * t]};
    mma1684(c, a0, Bt[g * 8 + 2 * t]);
    double a1[2] = {A[g * 8 + 2 * t + 1], A[(g + 8) * 8 + 2 * t + 1]};
    mma1684(c, a1, Bt[g * 8 + 2 * t + 1]);
    C84a[g * 8 + 2 * t] = c[0]; C84a[g * 8 + 2 * t + 1] = c[1];
    C84a[(g + 8) * 8 + 2 * t] = c[2]; C84a[(g + 8) * 8 + 2 * t + 1] = c[3];
  }
  {  // m8n8k4 on the first 8 rows
    double c[2] = {0, 0};
    mma884(c, A[g * 8 + 2 * t], Bt[g * 8 + 2 * t]);
    mma884(c, A[g * 8 + 2 * t + 1], Bt[g * 8 + 2 * t + 1]);
    C884[g * 8 + 2 * t] = c[0]; C884[g * 8 + 2 * t + 1] = c[1];
  }
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device: %s  SMs=%d  cc=%d.%d\n", prop.name, sms, prop.major, prop.minor);
  const int threads = 256;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * threads));

  // ---- layout probe
  {
    std::vector<double> A(16 * 8), Bt(8 * 8), ref(16 * 8, 0.0);
    for (int i = 0; i < 16 * 8; ++i) A[i] = sin(0.37 * i + 0.1);
    for (int i = 0; i < 8 * 8; ++i) Bt[i] = cos(0.53 * i + 0.2);
    for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) for (int k = 0; k < 8; ++k) ref[m * 8 + n] += A[m * 8 + k] * Bt[n * 8 + k];
    double *dA, *dB, *dC; CK(cudaMalloc(&dA, 128 * 8)); CK(cudaMalloc(&dB, 64 * 8)); CK(cudaMalloc(&dC, 3 * 128 * 8));
    CK(cudaMemcpy(dA, A.data(), 128 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bt.data(), 64 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, 3 * 128 * 8));
    k_probe<<<1, 32>>>(dA, dB, dC, dC + 128, dC + 256);
    CK(cudaDeviceSynchronize());
    std::vector<double> C(3 * 128);
    CK(cudaMemcpy(C.data(), dC, 3 * 128 * 8, cudaMemcpyDeviceToHost));
    double e88 = 0, e84 = 0, e884 = 0;
    for (int i = 0; i < 128; ++i) { e88 = fmax(e88, fabs(C[i] - ref[i])); e84 = fmax(e84, fabs(C[128 + i] - ref[i])); }
    for (int i = 0; i < 64; ++i) e884 = fmax(e884, fabs(C[256 + i] - ref[i]));
    printf("layout probe max|err|: m16n8k8=%.3e  2x m16n8k4=%.3e  2x m8n8k4=%.3e\n", e88, e84, e884);
  }

  FILE* js = fopen(argc > 1 ? argv[1] : "fp64_peaks.json", "w");
  fprintf(js, "{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);

  const int iters = 4096;
  for (int bps = 1; bps <= 4; bps *= 2) {   // blocks per SM: 8, 16, 32 warps/SM
    int grid = sms * bps;
    double fl_thread = (double)iters * 4 * NACC;
    {
      double ms = time_ms([&] { k_dfma<<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double tf = 2.0 * fl_thread * grid * threads / (ms * 1e-3) / 1e12;
      printf("DFMA            blocks/SM=%d  %.3f ms  %.2f TFLOP/s\n", bps, ms, tf);
      fprintf(js, ", \"dfma_bps%d\": %.3f", bps, tf);
    }
    double mma_per_warp = (double)iters * 4 * NACC;
    int warps = grid * threads / 32;
    struct { const char* name; double flop; } kinds[4] = {
      {"dmma_m8n8k4", 2.0 * 8 * 8 * 4}, {"dmma_m16n8k4", 2.0 * 16 * 8 * 4}, {"dmma_m16n8k8", 2.0 * 16 * 8 * 8}, {"dmma_m16n8k16", 2.0 * 16 * 8 * 16}};
    for (int k = 0; k < 4; ++k) {
      int it2 = iters / (k == 3 ? 4 : (k == 2 ? 2 : 1));
      double ms = time_ms([&] {
        if (k == 0) k_dmma<0><<<grid, threads>>>(out, it2, 1.0, 1e-9);
        if (k == 1) k_dmma<1><<<grid, threads>>>(out, it2, 1.0, 1e-9);
        if (k == 2) k_dmma<2><<<grid, threads>>>(out, it2, 1.0, 1e-9);
        if (k == 3) k_dmma<3><<<grid, threads>>>(out, it2, 1.0, 1e-9);
      }, 5);
      double tf = kinds[k].flop * (mma_per_warp * it2 / iters) * warps / (ms * 1e-3) / 1e12;
      printf("%-15s blocks/SM=%d  %.3f ms  %.2f TFLOP/s\n", kinds[k].name, bps, ms, tf);
      fprintf(js, ", \"%s_bps%d\": %.3f", kinds[k].name, bps, tf);
    }
  }
  // mixed: m16n8k8 (1024 FMA/warp = 32/thread) + NF DFMA/thread per MMA
  {
    int grid = sms * 2, it2 = iters / 2;
    int warps = grid * threads / 32;
    double mmas = (double)it2 * 4 * NACC;
    for (int nf = 0; nf <= 2; ++nf) {
      int NFv[3] = {4, 8, 16};
      double ms = time_ms([&] {
        if (nf == 0) k_mix<4><<<grid, threads>>>(out, it2, 1.0000001, 1e-9);
        if (nf == 1) k_mix<8><<<grid, threads>>>(out, it2, 1.0000001, 1e-9);
        if (nf == 2) k_mix<16><<<grid, threads>>>(out, it2, 1.0000001, 1e-9);
      }, 5);
      double tf_mma = 2048.0 * mmas * warps / (ms * 1e-3) / 1e12;
      double tf_fma = 2.0 * NFv[nf] * 32 * mmas * warps / (ms * 1e-3) / 1e12;
      printf("mix m16n8k8 + %2d DFMA/thread/MMA: %.3f ms  dmma %.2f + dfma %.2f = %.2f TFLOP/s\n", NFv[nf], ms, tf_mma, tf_fma, tf_mma + tf_fma);
      fprintf(js, ", \"mix_nf%d_dmma\": %.3f, \"mix_nf%d_dfma\": %.3f", NFv[nf], tf_mma, NFv[nf], tf_fma);
    }
  }
  // sustained: DMMA m16n8k8 back to back for ~3 s (power-capped clocks)
  {
    int grid = sms * 2, it2 = iters * 4;
    int warps = grid * threads / 32;
    double mmas = (double)it2 * 4 * NACC;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_dmma<2><<<grid, threads>>>(out, it2, 1.0, 1e-9);
    CK(cudaDeviceSynchronize());
    int n = 0; float ms = 0;
    CK(cudaEventRecord(e0));
    do {
      for (int r = 0; r < 10; ++r) k_dmma<2><<<grid, threads>>>(out, it2, 1.0, 1e-9);
      n += 10;
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
    } while (ms < 3000.f);
    double tf = 2048.0 * mmas * warps * n / (ms * 1e-3) / 1e12;
    printf("sustained dmma_m16n8k8 over %.1f s: %.2f TFLOP/s\n", ms * 1e-3, tf);
    fprintf(js, ", \"dmma_m16n8k8_sustained\": %.3f", tf);
    n = 0;
    CK(cudaEventRecord(e0));
    do {
      for (int r = 0; r < 10; ++r) k_dfma<<<grid, threads>>>(out, it2, 1.0000001, 1e-9);
      n += 10;
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
    } while (ms < 3000.f);
    tf = 2.0 * (double)it2 * 4 * NACC * grid * threads * n / (ms * 1e-3) / 1e12;
    printf("sustained dfma over %.1f s: %.2f TFLOP/s\n", ms * 1e-3, tf);
    fprintf(js, ", \"dfma_sustained\": %.3f", tf);
  }
  fprintf(js, "}\n");
  fclose(js);
  return 0;
}
