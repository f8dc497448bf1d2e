// Incremental cost (clocks per warp instruction, per sub-partition) of one instruction of a given kind when it is
// mixed into a stream of FP64 MMAs (m16n8k8 = 4 x DMMA.8x8x4, 64 clk of the FP64 pipe each) -- the cost model
// behind k_fwd3's instruction budget.  Payload instructions are independent (8 registers round robin), so this is
// throughput, not latency.  12 warps per SM (3 per sub-partition), like k_fwd3.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_cost tools/pipe_cost.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

enum { P_NONE, P_DFMA_RRR, P_DFMA_RIR, P_DFMA_RRI, P_DADD_RR, P_DADD_RI, P_DMUL_RR, P_DMUL_RI, P_LOP3, P_IADD, P_IMAD,
       P_FSEL, P_FFMA, P_MUFU_RCP64H, P_MUFU_EX2, P_LDS64, P_LDS128, P_SHFL, P_F2F_64_32, P_F2F_32_64, P_I2F64, P_ISETP_SEL,
       P_F2I64, P_DSETP, P_IMAD_HI, P_IMAD_HI_C, P_IMAD_WIDE, P_SHF, P_LEA, P_IADD3 };

template <int KIND>
__device__ __forceinline__ void payload(double (&f)[8], int (&x)[8], float (&s)[8], int i, double u, double v, int k1, int k2,
                                        const double* sm) {
  const int j = i & 7;
  if (KIND == P_DFMA_RRR) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[j]) : "d"(u), "d"(v));
  if (KIND == P_DFMA_RIR) asm volatile("fma.rn.f64 %0, %0, 0d3FEFFFFFFFFFF000, %1;" : "+d"(f[j]) : "d"(v));
  if (KIND == P_DFMA_RRI) asm volatile("fma.rn.f64 %0, %0, %1, 0d3FE0000000000000;" : "+d"(f[j]) : "d"(u));
  if (KIND == P_DADD_RR) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[j]) : "d"(v));
  if (KIND == P_DADD_RI) asm volatile("add.rn.f64 %0, %0, 0d3FF0000000000000;" : "+d"(f[j]));
  if (KIND == P_DMUL_RR) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(f[j]) : "d"(u));
  if (KIND == P_DMUL_RI) asm volatile("mul.rn.f64 %0, %0, 0d3FEFFFFFFFFFF000;" : "+d"(f[j]));
  if (KIND == P_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(k1), "r"(k2));
  if (KIND == P_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(k1));
  if (KIND == P_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(k1), "r"(k2));
  if (KIND == P_IMAD_HI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(k1));
  if (KIND == P_IMAD_HI_C) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(k1), "r"(k2));
  if (KIND == P_IMAD_WIDE) {
    unsigned long long w;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[j]), "r"(k1));
    asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %1; xor.b32 %0, lo, hi;}" : "=r"(x[j]) : "l"(w));
  }
  if (KIND == P_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 9;" : "+r"(x[j]) : "r"(k1));
  if (KIND == P_LEA) asm volatile("{.reg .u32 t; shl.b32 t, %0, 9; add.u32 %0, t, %1;}" : "+r"(x[j]) : "r"(k1));
  if (KIND == P_IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[j]) : "r"(k1), "r"(k2));
  if (KIND == P_FSEL) asm volatile("{.reg .pred p; setp.gt.s32 p, %1, 0; selp.b32 %0, %0, %2, p;}" : "+r"(x[j]) : "r"(k1), "r"(k2));
  if (KIND == P_ISETP_SEL) asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; selp.b32 %0, %0, %2, p;}" : "+r"(x[j]) : "r"(k1), "r"(k2));
  if (KIND == P_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[j]) : "f"(0.999f), "f"(s[(j + 1) & 7]));
  if (KIND == P_MUFU_RCP64H) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(f[j]));
  if (KIND == P_MUFU_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[j]));
  if (KIND == P_LDS64) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(f[j]) : "r"((unsigned)__cvta_generic_to_shared(sm + ((x[0] + 33 * j + threadIdx.x) & 2047))));
  if (KIND == P_LDS128) {
    double t2;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(f[j]), "=d"(t2) : "r"((unsigned)__cvta_generic_to_shared(sm + 2 * ((j * 32 + threadIdx.x) & 1023))));
    asm volatile("" :: "d"(t2));
  }
  if (KIND == P_SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(x[j]));
  // conversions: ptxas folds what it can prove, so each is a data-dependent round trip with one LOP3 in between
  // (2 payload "instructions" = cvt + lop3 [+ cvt back]; subtract LOP3's own cost when reading the numbers)
  if (KIND == P_F2F_64_32) {          // F2F.F64.F32 + LOP3 + F2F.F32.F64
    double d; int lo, hi;
    asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(s[j]));
    asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(d));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo) : "r"(k1), "r"(x[j]));
    asm volatile("mov.b64 %0, {%1, %2};" : "=d"(d) : "r"(lo), "r"(hi));
    asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(s[j]) : "d"(d));
  }
  if (KIND == P_F2F_32_64) {          // F2F.F32.F64 + LOP3 (on the float) feeding the low word of the double
    float t; int lo, hi;
    asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t) : "d"(f[j]));
    asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(f[j]));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo) : "r"(__float_as_int(t)), "r"(k2));
    asm volatile("mov.b64 %0, {%1, %2};" : "=d"(f[j]) : "r"(lo), "r"(hi));
  }
  if (KIND == P_I2F64) {              // I2F.F64.S32 + LOP3 of its low word into the integer
    double d; int lo, hi;
    asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d) : "r"(x[j]));
    asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(d));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(hi), "r"(k2));
  }
  if (KIND == P_F2I64) {              // F2I.S32.F64 + LOP3 into the double's low word
    int q, lo, hi;
    asm volatile("cvt.rni.s32.f64 %0, %1;" : "=r"(q) : "d"(f[j]));
    asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(f[j]));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo) : "r"(q), "r"(k2));
    asm volatile("mov.b64 %0, {%1, %2};" : "=d"(f[j]) : "r"(lo), "r"(hi));
  }
  if (KIND == P_DSETP) asm volatile("{.reg .pred p; setp.gt.f64 p, %1, %2; selp.b32 %0, %0, %3, p;}" : "+r"(x[j]) : "d"(f[j]), "d"(u), "r"(k2));
}

// NM MMAs + NP payload instructions per iteration
template <int KIND, int NM, int NP>
__global__ void __launch_bounds__(384, 1) k_cost(double* out, int iters, double u, double v, int k1, int k2) {
  __shared__ double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
  __syncthreads();
  double c[NM > 0 ? NM : 1][4];
#pragma unroll
  for (int i = 0; i < (NM > 0 ? NM : 1); ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  double a[4], b[2], f[8];
  int x[8];
  float s[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = u + threadIdx.x * 1e-6 + i * 1e-3;
  b[0] = v + threadIdx.x * 1e-6; b[1] = v * 0.5;
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = 1.0 + 0.1 * i + threadIdx.x * 1e-3; x[i] = threadIdx.x * 7 + i; s[i] = 0.5f + i; }
  for (int it = 0; it < iters; ++it) {
    // payload interleaved with the MMAs in program order (asm volatile keeps the order)
    if (NM > 0) {
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        mma1688(c[m], a, b);
#pragma unroll
        for (int i = 0; i < NP / NM; ++i) payload<KIND>(f, x, s, m * (NP / NM) + i, u, v, k1, k2, sm);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NP; ++i) payload<KIND>(f, x, s, i, u, v, k1, k2, sm);
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < (NM > 0 ? NM : 1); ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
#pragma unroll
  for (int i = 0; i < 8; ++i) r += f[i] + x[i] + s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// latency: 1 warp per sub-partition, NP dependent instructions of one kind on ONE register per iteration
template <int KIND, int NP>
__global__ void __launch_bounds__(128, 1) k_lat(double* out, int iters, double u, double v, int k1, int k2) {
  __shared__ double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
  __syncthreads();
  double f[8]; int x[8]; float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = 1.0 + 0.1 * i + threadIdx.x * 1e-3; x[i] = threadIdx.x * 7 + i; s[i] = 0.5f + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NP; ++i) payload<KIND>(f, x, s, 0, u, v, k1, k2, sm);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += f[i] + x[i] + s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int KIND, int NP>
void lat(int sms, double* out, const char* what) {
  const int iters = 2048;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_lat<KIND, NP><<<sms, 128>>>(out, iters, 0.9999999, 1e-9, 0x5bd1e995, 12345);
  cudaEventRecord(e0);
  k_lat<KIND, NP><<<sms, 128>>>(out, iters, 0.9999999, 1e-9, 0x5bd1e995, 12345);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("latency %-24s %.2f clk per dependent instr\n", what, ms * 1e-3 * 1.965e9 / iters / NP);
}

template <int KIND, int NM, int NP>
double run(int sms, double* out, const char* what, double base_per_it) {
  const int iters = 2048;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_cost<KIND, NM, NP><<<sms, 384>>>(out, iters, 0.9999999, 1e-9, 0x5bd1e995, 12345);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) k_cost<KIND, NM, NP><<<sms, 384>>>(out, iters, 0.9999999, 1e-9, 0x5bd1e995, 12345);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 3;
  const double clk_per_it = ms * 1e-3 * 1.965e9 / iters / 3.0;      // per warp-iteration on a sub-partition (3 warps share it)
  printf("%-28s mma/it=%d payload/it=%2d  %.3f ms  clk per warp-iteration %.1f", what, NM, NP, ms, clk_per_it);
  if (NP) printf("  => %.2f clk per payload instr", (clk_per_it - base_per_it) / NP);
  printf("  [%s]\n", cudaGetErrorString(cudaGetLastError()));
  return clk_per_it;
}

#define ROW(K, name)                                                   \
  run<K, 4, 8>(sms, out, name " (4 mma + 8)", base4);                  \
  run<K, 4, 32>(sms, out, name " (4 mma + 32)", base4);                \
  run<K, 0, 32>(sms, out, name " (no mma, 32)", 0.0);

int main(int argc, char**) {   // no args: throughput table; 1 arg: conversions; 2 args: latencies; 3 args: integer multiplies / shifts
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 384);
  printf("%s, %d SMs, 12 warps/SM\n", prop.name, sms);
  const double base4 = run<P_NONE, 4, 0>(sms, out, "4 mma", 0.0);
  if (argc == 3) {
    lat<P_DFMA_RRR, 32>(sms, out, "DFMA");
    lat<P_DADD_RR, 32>(sms, out, "DADD");
    lat<P_DMUL_RR, 32>(sms, out, "DMUL");
    lat<P_LOP3, 32>(sms, out, "LOP3");
    lat<P_IMAD, 32>(sms, out, "IMAD");
    lat<P_FFMA, 32>(sms, out, "FFMA (2 regs)");
    lat<P_MUFU_RCP64H, 32>(sms, out, "MUFU.RCP64H");
    lat<P_MUFU_EX2, 32>(sms, out, "MUFU.EX2");
    lat<P_SHFL, 32>(sms, out, "SHFL");
    lat<P_I2F64, 32>(sms, out, "I2F.F64+LOP3");
    lat<P_F2I64, 32>(sms, out, "F2I.F64+LOP3");
    return 0;
  }
  if (argc > 3) {
    ROW(P_IMAD, "IMAD")
    ROW(P_IMAD_HI, "IMAD.HI")
    ROW(P_IMAD_HI_C, "IMAD.HI + c")
    ROW(P_IMAD_WIDE, "IMAD.WIDE + LOP3")
    ROW(P_SHF, "SHF")
    ROW(P_LEA, "SHL+ADD")
    ROW(P_IADD3, "ADD+ADD")
    ROW(P_LOP3, "LOP3")
    return 0;
  }
  if (argc > 1) {
    ROW(P_F2F_64_32, "F2F.64.32+LOP3+F2F.32.64")
    ROW(P_F2F_32_64, "F2F.F32.F64+LOP3")
    ROW(P_I2F64, "I2F.F64.S32+LOP3")
    ROW(P_F2I64, "F2I.S32.F64+LOP3")
    return 0;
  }
  ROW(P_DFMA_RRR, "DFMA r,r,r")
  ROW(P_DFMA_RIR, "DFMA r,imm,r")
  ROW(P_DFMA_RRI, "DFMA r,r,imm")
  ROW(P_DADD_RR, "DADD r,r")
  ROW(P_DADD_RI, "DADD r,imm")
  ROW(P_DMUL_RR, "DMUL r,r")
  ROW(P_DMUL_RI, "DMUL r,imm")
  ROW(P_DSETP, "DSETP+SEL")
  ROW(P_LOP3, "LOP3")
  ROW(P_IADD, "IADD")
  ROW(P_IMAD, "IMAD")
  ROW(P_FSEL, "ISETP+SEL (const pred)")
  ROW(P_ISETP_SEL, "ISETP+SEL")
  ROW(P_FFMA, "FFMA")
  ROW(P_MUFU_RCP64H, "MUFU.RCP64H")
  ROW(P_MUFU_EX2, "MUFU.EX2")
  ROW(P_LDS64, "LDS.64 scattered")
  ROW(P_LDS128, "LDS.128 linear")
  ROW(P_SHFL, "SHFL")
  ROW(P_F2F_64_32, "F2F.F64.F32")
  ROW(P_F2F_32_64, "F2F.F32.F64")
  ROW(P_I2F64, "I2F.F64.S32")
  ROW(P_F2I64, "F2I.S32.F64")
  return 0;
}
