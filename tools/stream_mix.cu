// Does spreading the activation instructions evenly between the FP64 MMAs (one warp's stream = MMA, a few activation
// instructions, MMA, ...) hide the activation's integer / FP32 / load instructions in the issue slots the MMAs leave
// free?  Cost model measured on the existing kernels (tools/fp64_mix.cu, profiles/r02_fwd3_v*_ncu.txt):
//   T = 64.4 clk per m16n8k8 + 2.9 clk per non-MMA FP64 instruction + 0.75 clk per other instruction
// while an instruction that lands in the shadow of an MMA costs 0.2 clk (tools/pipe_cost.cu).
//   MODE 0  per iteration NM MMAs, then NE fast swish evaluations (plain C++: ptxas schedules)
//   MODE 1  the same work with every activation instruction an asm volatile statement, issued in slices after each MMA
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/stream_mix tools/stream_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "exp_fixed_point.cuh"

// ---- the fast swish as single volatile instructions (one element), cut into 10 slices
struct SwishV {
  double z, t, T, d, y, e;
  unsigned hi, F, a, b, c;
  float y32, ff, h, gm;
  unsigned live;        // a 32-bit value produced by the last slice (token source)
  __device__ __forceinline__ unsigned tok() const { return live; }
  // make the next slice depend on `v` (a freshly loaded B fragment): OR (v & zero) into the values it reads first
  __device__ __forceinline__ void tie(unsigned v, unsigned zero) {
    unsigned k;
    asm volatile("and.b32 %0, %1, %2;" : "=r"(k) : "r"(v), "r"(zero));
    unsigned lo, hi2;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi2) : "d"(z));
    asm volatile("or.b32 %0, %0, %1;" : "+r"(lo) : "r"(k));
    asm volatile("mov.b64 %0, {%1, %2};" : "=d"(z) : "r"(lo), "r"(hi2));
    asm volatile("or.b32 %0, %0, %1;" : "+r"(hi) : "r"(k));
    asm volatile("or.b32 %0, %0, %1;" : "+r"(a) : "r"(k));
  }
  template <int S>
  __device__ __forceinline__ void slice(const double* tab, unsigned c433, unsigned c3f8) {
    if (S == 0) asm volatile("fma.rn.f64 %0, %1, 0dC0A71547652B82FE, 0d4118000000000000;" : "=d"(t) : "d"(z));   // -INV, 393216
    if (S == 1) {
      asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(F), "=r"(hi) : "d"(t));
      asm volatile("shl.b32 %0, %1, 1;" : "=r"(a) : "r"(hi));
      asm volatile("and.b32 %0, %0, 0x3FF8;" : "+r"(a));
      asm volatile("shf.r.wrap.b32 %0, %1, %2, 11;" : "=r"(b) : "r"(F), "r"(hi));
    }
    if (S == 2) {
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"((unsigned)__cvta_generic_to_shared(tab) + a));
      asm volatile("lop3.b32 %0, %0, 0x007FFFFF, %1, 0xEA;" : "+r"(b) : "r"(c3f8));          // (b & m) | c
      asm volatile("mov.b32 %0, %1;" : "=f"(y32) : "r"(b));
      asm volatile("add.f32 %0, %1, 0fBF7FFFFF;" : "=f"(ff) : "f"(y32));
    }
    if (S == 3) {
      asm volatile("fma.rn.f32 %0, %1, 0f43A3FEA0, 0f4A317198;" : "=f"(h) : "f"(ff));        // 327.98926, 2907270
      asm volatile("mul.f32 %0, %1, %1;" : "=f"(gm) : "f"(ff));
      asm volatile("and.b32 %0, %1, 0x000FE000;" : "=r"(c) : "r"(hi));
      asm volatile("shl.b32 %0, %0, 7;" : "+r"(c));
    }
    if (S == 4) {
      asm volatile("fma.rn.f32 %0, %0, %1, 0f4B000000;" : "+f"(gm) : "f"(h));
      asm volatile("lop3.b32 %0, %0, 3, %1, 0xEA;" : "+r"(hi) : "r"(c433));                   // (hi & 3) | 0x43300000
      unsigned tl, th;
      asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(tl), "=r"(th) : "d"(T));
      asm volatile("add.s32 %0, %0, %1;" : "+r"(th) : "r"(c));
      asm volatile("mov.b64 %0, {%1, %2};" : "=d"(T) : "r"(tl), "r"(th));
    }
    if (S == 5) {
      unsigned g;
      asm volatile("mov.b32 %0, %1;" : "=r"(g) : "f"(gm));
      asm volatile("add.s32 %0, %0, 0xB5000000;" : "+r"(g));
      asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(F) : "r"(g));
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(hi));
      asm volatile("mov.b64 %0, {%1, %2};" : "=d"(d) : "r"(F), "r"(hi));
    }
    if (S == 6) asm volatile("fma.rn.f64 %0, %0, 0d3D162E42FEFA39EF, 0dC055EE42FEFA39EF;" : "+d"(d));             // D1
    if (S == 7) {
      asm volatile("fma.rn.f64 %0, %1, %0, 0d3FF0000000000000;" : "+d"(d) : "d"(T));                             // 1 + exp
      asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    }
    if (S == 8) {
      asm volatile("{.reg .f64 nd; neg.f64 nd, %1; fma.rn.f64 %0, nd, %2, 0d3FF0000000000000;}" : "=d"(e) : "d"(d), "d"(y));
    }
    if (S == 9) asm volatile("fma.rn.f64 %0, %0, %0, %0;" : "+d"(e));
    if (S == 10) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(y) : "d"(e));
    if (S == 11) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(z) : "d"(y));
    // token: some 32-bit piece of what this slice produced
    if (S == 0) live = (unsigned)__double2loint(t);
    else if (S <= 5) live = hi ^ a ^ b ^ c ^ F ^ __float_as_uint(gm) ^ __float_as_uint(ff);
    else if (S <= 7) live = (unsigned)__double2loint(d);
    else if (S <= 9) live = (unsigned)__double2loint(e);
    else if (S == 10) live = (unsigned)__double2loint(y);
    else live = (unsigned)__double2loint(z);
  }
};

template <int MODE, int NM, int NE>
__global__ void __launch_bounds__(384, 1) k_stream(double* out, int iters, double x, double y, const double* gtab, unsigned c433,
                                                    unsigned c3f8) {
  __shared__ double tab[BNN_EXP_TAB_SIZE];
  for (int i = threadIdx.x; i < BNN_EXP_TAB_SIZE; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  double c[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  double a[4], b[2], f[NE];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = x + threadIdx.x * 1e-6 + i * 1e-3;
  b[0] = y + threadIdx.x * 1e-6; b[1] = y * 0.5;
#pragma unroll
  for (int i = 0; i < NE; ++i) f[i] = 0.3 * i + threadIdx.x * 1e-3;
  const double off = 0.25 + (threadIdx.x & 31) * 0.37;
  __shared__ double wsm[32 * 64];                                   // B fragments: one 16-byte load per MMA and lane
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) wsm[i] = y * (1 + (i & 7));
  __syncthreads();
  const unsigned wbase = (unsigned)__cvta_generic_to_shared(wsm) + (threadIdx.x & 31) * 16;
  const unsigned zero = c433 & 1u;                                  // a zero ptxas cannot see through
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2) {
      // the same slices, pinned: the B-fragment load of MMA m takes its address through a slice-s value, and the
      // first instruction of the next slice takes an operand through the loaded value => s -> LDS_m -> s+1
      SwishV sv[NE];
#pragma unroll
      for (int q = 0; q < NE; ++q) sv[q].z = f[q];
      unsigned tok = 0;
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        unsigned addr;
        asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(addr) : "r"(tok), "r"(zero), "r"(wbase + (m & 3) * 512));
        double b0, b1;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(b0), "=d"(b1) : "r"(addr));
        dmma16x8x8(c[m & 1], a[0], a[1], a[2], a[3], b0, b1);
        unsigned blo = (unsigned)__double2loint(b0);
#define SL(S) if (S * NM / 12 == m) { _Pragma("unroll") for (int q = 0; q < NE; ++q) { sv[q].tie(blo, zero); sv[q].template slice<S>(tab, c433, c3f8); } tok = sv[0].tok(); }
        SL(0) SL(1) SL(2) SL(3) SL(4) SL(5) SL(6) SL(7) SL(8) SL(9) SL(10) SL(11)
#undef SL
      }
#pragma unroll
      for (int q = 0; q < NE; ++q) f[q] = sv[q].z + off;
    } else if (MODE == 0) {
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const double2 bb = *reinterpret_cast<const double2*>(wsm + (threadIdx.x & 31) * 2 + (m & 3) * 64);
        dmma16x8x8(c[m & 1], a[0], a[1], a[2], a[3], bb.x, bb.y);
      }
#pragma unroll
      for (int q = 0; q < NE; ++q) f[q] = expfix_act_fast<BNN_ACT_SWISH>(f[q], tab) + off;
    } else {
      // 12 slices x NE elements spread over NM MMAs: slice s of all elements goes after MMA floor(s * NM / 12)
      SwishV sv[NE];
#pragma unroll
      for (int q = 0; q < NE; ++q) sv[q].z = f[q];
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const double2 bb = *reinterpret_cast<const double2*>(wsm + (threadIdx.x & 31) * 2 + (m & 3) * 64);
        dmma16x8x8(c[m & 1], a[0], a[1], a[2], a[3], bb.x, bb.y);
#define SL(S) if (S * NM / 12 == m) { _Pragma("unroll") for (int q = 0; q < NE; ++q) sv[q].template slice<S>(tab, c433, c3f8); }
        SL(0) SL(1) SL(2) SL(3) SL(4) SL(5) SL(6) SL(7) SL(8) SL(9) SL(10) SL(11)
#undef SL
      }
#pragma unroll
      for (int q = 0; q < NE; ++q) f[q] = sv[q].z + off;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
#pragma unroll
  for (int i = 0; i < NE; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NM, int NE>
void run(int sms, double* out, const double* tab, const char* what) {
  const int iters = 4096 / NM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_stream<MODE, NM, NE><<<sms, 384>>>(out, iters, 1.0000001, 1e-9, tab, 0x43300000u, 0x3F800000u);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k_stream<MODE, NM, NE><<<sms, 384>>>(out, iters, 1.0000001, 1e-9, tab, 0x43300000u, 0x3F800000u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  const double clk_it = ms * 1e-3 * 1.965e9 / iters / 3.0;      // per warp-iteration on a sub-partition (3 warps share it)
  printf("%-28s mma/it=%2d swish/it=%d  %.3f ms  clk per warp-iteration %.1f = %d x 64.4 + %d x %.1f   [%s]\n", what, NM, NE, ms,
         clk_it, NM, NE, (clk_it - 64.4 * NM) / NE, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double *out, *dt;
  cudaMalloc(&out, sizeof(double) * sms * 384);
  cudaMalloc(&dt, sizeof(double) * BNN_EXP_TAB_SIZE);
  double tab[BNN_EXP_TAB_SIZE];
  for (int j = 0; j < BNN_EXP_TAB_SIZE; ++j) tab[j] = ldexp(exp2((double)j / BNN_EXP_TAB_SIZE), -EXPFIX_TAB_BIAS);
  cudaMemcpy(dt, tab, sizeof(tab), cudaMemcpyHostToDevice);
  printf("%s, %d SMs, 12 warps/SM\n", prop.name, sms);
  run<0, 12, 4>(sms, out, dt, "ptxas order");
  run<1, 12, 4>(sms, out, dt, "volatile slices");
  run<2, 12, 4>(sms, out, dt, "pinned slices");
  run<2, 24, 8>(sms, out, dt, "pinned slices");
  run<2, 12, 8>(sms, out, dt, "pinned slices");
  run<2, 24, 4>(sms, out, dt, "pinned slices");
  run<0, 24, 4>(sms, out, dt, "ptxas order");
  run<0, 24, 8>(sms, out, dt, "ptxas order");
  run<1, 24, 8>(sms, out, dt, "volatile slices");
  run<0, 12, 8>(sms, out, dt, "ptxas order");
  run<1, 12, 8>(sms, out, dt, "volatile slices");
  run<0, 8, 4>(sms, out, dt, "ptxas order");
  run<1, 8, 4>(sms, out, dt, "volatile slices");
  return 0;
}
