// Shared device/host definitions for the npbnn_b200 kernels (sm_100a).
//
// Data layout in HBM:
//   X     "swizzled rows": [n_pad16, F_pad] f64, element (r, c) at r*F_pad + (c ^ ((r&1)*x_swz)).
//         F_pad = F rounded up to 8, rows rounded up to 16, zero filled.  The XOR swaps the two
//         64-byte halves of each 128-byte group on odd rows so that the 16-byte fragment loads of a
//         quarter-warp (2 rows x 4 chunks) hit 8 different 16-byte bank groups.
//   W     "packed weight set": per layer a [out_pad, stride] matrix in the same swizzled-row form
//         (bias column removed) followed by bias[out_pad]; padding is zero and never written.
//   both are moved to shared memory verbatim (cp.async.bulk), so global layout == smem layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/npbnn_b200.h"

#ifndef BNN_EXP_TAB_BITS
#define BNN_EXP_TAB_BITS 11      // 2048-entry table of 2^(j/2048) (16 KB of shared memory) + degree-3 polynomial
#endif
#define BNN_EXP_TAB_SIZE (1 << BNN_EXP_TAB_BITS)

struct LayerGeom {
  int in, out, bias;        // canonical: W is [out, in + bias]
  int in_pad, out_pad;      // rounded up to 8
  int stride, swz;          // packed row stride (= in_pad) and XOR constant (8 or 0)
  int w_off, b_off;         // offsets (doubles) of matrix / bias inside a packed weight set
  int c_off;                // offset (doubles) of the layer inside a canonical weight set
};

struct NetGeom {
  int L;
  int F, F_pad, x_swz;
  int act, lik;
  int O;                    // width of the output layer
  int K;                    // classes (CATEGORICAL) or number of modelled outputs (Gaussian)
  int P;                    // canonical parameters per weight set
  int PB;                   // packed doubles per weight set
  int max_w;                // widest padded activation (for smem staging)
  LayerGeom l[BNN_MAX_LAYERS];
};

__host__ __device__ inline int bnn_round_up(int v, int m) { return (v + m - 1) / m * m; }
__host__ __device__ inline int bnn_swz_for(int stride) { return (stride % 16 == 0) ? 8 : 0; }

// canonical (layer, r, cc) -> offset inside a packed weight set
__host__ __device__ inline int bnn_packed_index(const LayerGeom& g, int r, int cc) {
  if (g.bias) {
    if (cc == 0) return g.b_off + r;
    cc -= 1;
  }
  return g.w_off + r * g.stride + (cc ^ ((r & 1) * g.swz));
}

// ---------------------------------------------------------------------------------------------
// FP64 tensor-core MMA (DMMA).  Fragment <-> matrix mapping (g = lane>>2, t = lane&3), with the K
// index permuted inside each group of 8 columns (k-slot t <-> column 2t, k-slot t+4 <-> column 2t+1;
// a sum over k is order-free) so that one 16-byte load feeds both k-slots and the accumulator of
// one layer IS the A operand of the next:
//   a0=A[g][8j+2t] a1=A[g+8][8j+2t] a2=A[g][8j+2t+1] a3=A[g+8][8j+2t+1]
//   b0=W[n0+g][8j+2t] b1=W[n0+g][8j+2t+1]
//   c0=C[g][n0+2t] c1=C[g][n0+2t+1] c2=C[g+8][n0+2t] c3=C[g+8][n0+2t+1]
// Verified on a B200 by tools/peak_fp64.cu (layout probe, profiles/r01_fp64_peaks.log).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma16x8x8(double (&c)[4], double a0, double a1, double a2, double a3,
                                           double b0, double b1) {
#ifdef BNN_DBG_NODMMA         // tuning experiment only: is the FP64 MMA what something else is waiting for?
  c[0] += a0 * b0; c[1] += a1 * b1; c[2] += a2 * b0; c[3] += a3 * b1;
  return;
#endif
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}

// ---------------------------------------------------------------------------------------------
// FP64 transcendentals tuned for the shared FP64 pipe (DMMA and DFMA issue to the same pipe on B200,
// profiles/r01_fp64_peaks.log), i.e. as few FP64 instructions as possible:
//   exp : 2048-entry table of 2^(j/2048) + degree-3 polynomial, 8 FP64 instructions, |rel err| < 3e-16
//         (BNN_EXP_TAB_BITS=8: 256 entries + degree 4, 9 instructions)
//   rcp : MUFU.RCP64H seed + one third-order step, 3 FP64 instructions
// ---------------------------------------------------------------------------------------------
// TB = log2(table entries): 11 (2048 entries + degree 3) or 8 (256 entries + degree 4, one more instruction;
// used where shared memory is short)
//
// ONE_STEP: argument reduction with the single rounded constant ln2/2^TB (one FMA instead of two).  The
// reduced argument is then off by at most |x| * 3.4e-17, i.e. exp(x) by that relative amount -- harmless where
// the exponential feeds a sigmoid: d swish / d r = z s (1-s) and d tanh / d r are bounded, so the activation
// moves by < 1e-16 ABSOLUTE for every z.  The softmax terms (arguments <= 0, see bnn_exp_neg) use it as well.
template <int TB = BNN_EXP_TAB_BITS, bool ONE_STEP = false>
__device__ __forceinline__ double bnn_exp_core(double x, const double* __restrict__ tab) {
  static_assert(TB == 8 || TB == 11, "exp table: 256 or 2048 entries");
  const double MAGIC = 6755399441055744.0;           // 1.5 * 2^52: round-to-nearest-integer trick
  // INV = 2^TB / ln 2 ; C_HI + C_LO = ln2 / 2^TB with the low 24 (21) mantissa bits of C_HI zero
  const double INV = (TB == 11) ? 2954.639443740597 : 369.3299304675746271;
  const double C_HI = (TB == 11) ? 0.0003384507708688034 : 0.00270760617331689;
  const double C_LO = (TB == 11) ? 8.889824211446026e-13 : 7.453964567463233e-13;
  double t = fma(x, INV, MAGIC);
  int k = __double2loint(t);
  double kd = t - MAGIC;
  double r;
  if (ONE_STEP) {
    r = fma(kd, (TB == 11) ? -0.0003384507717577858 : -0.0027076061740622863, x);
  } else {
    r = fma(kd, -C_HI, x);
    r = fma(kd, -C_LO, r);
  }
  double q;
  if (TB == 11) {
    q = fma(r, 1.66666666666666657e-01, 0.5);        // |r| <= ln2/4096: r^4/24 < 4e-17
  } else {
    q = fma(r, 4.16666666666666644e-02, 1.66666666666666657e-01);
    q = fma(r, q, 0.5);
  }
  double r2 = r * r;
  double p = fma(r2, q, r);
  double T = tab[k & ((1 << TB) - 1)];
  double res = fma(T, p, T);
  int n = k >> TB;
  return __hiloint2double(__double2hiint(res) + (n << 20), __double2loint(res));
}

// exp(SCALE * z) for SCALE = -1 (swish: exp(-z)) or 2 (tanh: exp(2z)) with the scale folded into the constants, so
// that neither -z nor 2z is formed in an FP64 instruction (ptxas emits DADD for a negation that feeds an FMA chain).
// One-step argument reduction as above (activations only); 7 FP64 instructions (8 with the 256-entry table).
//   rs = z - k * (C / SCALE)  (= r / SCALE, exact scaling);  exp(r) = 1 + p
//   SCALE -1:  p = -rs + rs^2 (1/2 - rs/6)
//   SCALE  2:  p = 2 [rs + rs^2 (1 + 2 rs / 3)], the factor 2 goes into the table value (exponent + 1, integer add)
template <int TB, int SCALE>
__device__ __forceinline__ double bnn_exp_scaled(double z, const double* __restrict__ tab) {
  static_assert(TB == 8 || TB == 11, "exp table: 256 or 2048 entries");
  static_assert(SCALE == -1 || SCALE == 2, "exp(-z) or exp(2z)");
  const double MAGIC = 6755399441055744.0;
  const double INV = (TB == 11) ? 2954.639443740597 : 369.3299304675746271;
  const double C1 = (TB == 11) ? 0.0003384507717577858 : 0.0027076061740622863;      // ln2 / 2^TB
  double t = fma(z, (double)SCALE * INV, MAGIC);
  int k = __double2loint(t);
  double kd = t - MAGIC;
#ifdef BNN_DBG_NOTAB          // tuning experiment only: what do the table lookups (random shared-memory reads) cost?
  double T = 1.0;
#else
  double T = tab[k & ((1 << TB) - 1)];
#endif
  double res;
  if (SCALE == -1) {
    double rs = fma(kd, C1, z);                       // = -r
    double q;
    if (TB == 11) {
      q = fma(rs, -1.66666666666666657e-01, 0.5);
    } else {                                          // 256 entries: |r| <= ln2/512 needs the fourth-order term
      q = fma(rs, 4.16666666666666644e-02, -1.66666666666666657e-01);
      q = fma(rs, q, 0.5);
    }
    double r2 = rs * rs;
    double p = fma(r2, q, -rs);
    res = fma(T, p, T);
  } else {
    double rs = fma(kd, -0.5 * C1, z);                // = r / 2
    double q;
    if (TB == 11) {
      q = fma(rs, 6.66666666666666630e-01, 1.0);
    } else {
      q = fma(rs, 3.33333333333333315e-01, 6.66666666666666630e-01);
      q = fma(rs, q, 1.0);
    }
    double r2 = rs * rs;
    double p = fma(r2, q, rs);
    const double T2 = __hiloint2double(__double2hiint(T) + (1 << 20), __double2loint(T));    // 2 T (T in [1, 2))
    res = fma(T2, p, T);
  }
  int n = k >> TB;
  return __hiloint2double(__double2hiint(res) + (n << 20), __double2loint(res));
}

// the same with z clamped to +-708 / |SCALE| (selects on the integer view of z; NaN handling is the caller's)
template <int TB, int SCALE>
__device__ __forceinline__ double bnn_exp_scaled_clamped(double z, const double* __restrict__ tab) {
  constexpr int THR = (SCALE == 2) ? 0x40762000 : 0x40862000;          // 354.0 / 708.0
  const int hx = __double2hiint(z);
  const bool big = (hx & 0x7fffffff) >= THR;                            // also inf / NaN
  const double zc = big ? __hiloint2double((hx & 0x80000000) | THR, 0) : z;
  return bnn_exp_scaled<TB, SCALE>(zc, tab);
}

// x is NaN, as integer instructions only (the 64-bit AND goes through inline PTX: written in C++, ptxas recognises
// fabs() and materialises it with a DADD on the FP64 pipe)
__device__ __forceinline__ bool bnn_is_nan_int(double x) {
  unsigned long long a;
  asm("and.b64 %0, %1, 0x7fffffffffffffff;" : "=l"(a) : "l"(__double_as_longlong(x)));
  return a > 0x7ff0000000000000ULL;
}

// Branch-free range handling (selects on the integer view of x; no BSSY/BSYNC, so ptxas can interleave the
// FP64 chains of several independent evaluations with the surrounding DMMAs).
//
// exp for arguments <= 0 (softmax): results below 2^-1022 flush to 0, NaN propagates.
template <int TB = BNN_EXP_TAB_BITS>
__device__ __forceinline__ double bnn_exp_neg(double x, const double* __restrict__ tab) {
  const int hx = __double2hiint(x);
  const int ax = hx & 0x7fffffff;
  const bool big = ax >= 0x4086232c;                  // |x| >= 708.3965 (also inf / NaN): exp(x) < 2^-1022
  // one-step reduction: exp(x) is off by at most |x| * 3.4e-17 relative, i.e. a softmax term e^x <= 1 by at most
  // 0.37 * 3.4e-17 absolute -- invisible in the row sum (>= 1) and 2e-14 relative in a probability of 1e-300
  double res = bnn_exp_core<TB, true>(big ? -708.0 : x, tab);   // clamped so the exponent arithmetic stays in range
  res = big ? 0.0 : res;
  // NaN in => NaN out, by OR-ing quiet-NaN bits into the result (an integer op: a select here makes ptxas
  // branch around the whole evaluation, which breaks the interleaving with the surrounding MMAs)
  const bool is_nan = bnn_is_nan_int(x);
  return __hiloint2double(__double2hiint(res) | (is_nan ? 0x7ff80000 : 0), __double2loint(res));
}

// exp for activations: argument clamped to [-708, 708] (1/(1+e) is then NaN-free and the clamped tails
// differ from the exact value by < 1e-300 in the activation); NaN handling is done by the caller.
template <int TB = BNN_EXP_TAB_BITS>
__device__ __forceinline__ double bnn_exp_clamped(double x, const double* __restrict__ tab) {
  const int hx = __double2hiint(x);
  const bool big = (hx & 0x7fffffff) >= 0x40862000;           // |x| >= 708 (also inf / NaN)
  const double xc = big ? __hiloint2double((hx & 0x80000000) | 0x40862000, 0) : x;
  return bnn_exp_core<TB, true>(xc, tab);
}

// 1/d for finite d >= 1: MUFU.RCP64H seed (rcp.approx.ftz.f64) + one third-order step
//   e = 1 - d*y0 ; y = y0 + y0*(e + e*e)      (error e^3; parity tests pass at 1e-9 with margin ~1e-13)
// BNN_RCP_NEWTON2 selects two Newton steps (4 instructions) instead.
__device__ __forceinline__ double bnn_rcp(double d) {
  double y;
#ifdef BNN_DBG_NOMUFU         // tuning experiment only: what does the MUFU seed cost?
  y = __hiloint2double(0x7fe00000 - __double2hiint(d), 0);
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#endif
  double e = fma(-d, y, 1.0);
#ifndef BNN_RCP_NEWTON2
  e = fma(e, e, e);
  y = fma(y, e, y);
#else
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  y = fma(y, e, y);
#endif
  return y;
}

// log(x) for finite x >= 1 (sum of softmax terms, or a product of up to 16 of them), branch-free, ~22 FP64
// instructions: x = 2^e m, m in [0.71, 1.42); log m = 2 atanh(f), f = (m-1)/(m+1), series to f^21; |err| < 3e-16.
__device__ __forceinline__ double bnn_log_ge1(double x) {
  int hi = __double2hiint(x);
  int e = (hi >> 20) - 1023;
  const int up = ((hi & 0xfffff) > 0x6a09e) ? 1 : 0;           // mantissa above sqrt(2): halve it
  e += up;
  const double m = __hiloint2double((hi & 0xfffff) | ((0x3ff - up) << 20), __double2loint(x));
  const double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;   // (double)e
  const double f = (m - 1.0) * bnn_rcp(m + 1.0);
  const double f2 = f * f;
  double p = 4.76190476190476164e-02;                           // 1/21
  p = fma(p, f2, 5.26315789473684181e-02);                      // 1/19
  p = fma(p, f2, 5.88235294117647051e-02);                      // 1/17
  p = fma(p, f2, 6.66666666666666657e-02);                      // 1/15
  p = fma(p, f2, 7.69230769230769273e-02);                      // 1/13
  p = fma(p, f2, 9.09090909090909116e-02);                      // 1/11
  p = fma(p, f2, 1.11111111111111105e-01);                      // 1/9
  p = fma(p, f2, 1.42857142857142849e-01);                      // 1/7
  p = fma(p, f2, 2.00000000000000011e-01);                      // 1/5
  p = fma(p, f2, 3.33333333333333315e-01);                      // 1/3
  const double f2x = f + f;
  const double lm = fma(f2x * f2, p, f2x);
  return fma(ed, 0.6931471675634384, fma(ed, 1.2996506893889889e-08, lm));   // ln2 split hi (26 bits) + lo
}

// softplus(z) = np.logaddexp(0, z) = max(z, 0) + log1p(exp(-|z|)) (BNN_lib.py:170-172) and its logarithm, for the
// sigma head of the Gaussian likelihood, without libm calls (each is a 400-500 clk dependent chain that the warps of
// k_fwd3 would all sit behind): exp from the table routine, log1p(e) = log(u) + (e - (u - 1)) / u with u = fl(1 + e)
// (the correction restores what the rounding of 1 + e lost: full relative accuracy down to e ~ 1e-304), log by
// bnn_log_ge1's algorithm, which holds for every positive normal argument.  |z| >= 700 / inf / NaN: caller's libm path.
__device__ __forceinline__ double bnn_softplus_fast(double z, const double* __restrict__ tab) {
  const double az = fabs(z);
  const double e = bnn_exp_neg<BNN_EXP_TAB_BITS>(-az, tab);          // (0, 1]
  const double u = 1.0 + e;
  const double c = e - (u - 1.0);
  const double l1p = fma(c, bnn_rcp(u), bnn_log_ge1(u));
  return fmax(z, 0.0) + l1p;
}

// Hidden-layer activation, matching the reference formulas (BNN_lib.py:50-66):
//   swish z*(1+exp(-z))^-1 ; tanh 1 - 2/(exp(2z)+1) ; ReLU ; leaky (alpha*z for z<0)
template <int ACT, int TB = BNN_EXP_TAB_BITS>
__device__ __forceinline__ double bnn_act(double z, double alpha, const double* __restrict__ tab) {
  if (ACT == BNN_ACT_RELU) return z < 0.0 ? 0.0 : z;
  if (ACT == BNN_ACT_LEAKY) return z < 0.0 ? alpha * z : z;
  if (ACT == BNN_ACT_SWISH) {
    // NaN: z * finite = NaN.  +-inf: inf * 1 = inf, -inf * ~0 -> the reference gives NaN (-inf * 0); here
    // -inf * 3e-308 = -inf.  Both poison the likelihood (NaN or -inf log-posterior => proposal rejected).
    double e = bnn_exp_scaled_clamped<TB, -1>(z, tab);
    return z * bnn_rcp(1.0 + e);
  }
  double e = bnn_exp_scaled_clamped<TB, 2>(z, tab);
  double r = fma(-2.0, bnn_rcp(e + 1.0), 1.0);
  const bool is_nan = bnn_is_nan_int(z);
  return __hiloint2double(__double2hiint(r) | (is_nan ? 0x7ff80000 : 0), __double2loint(r));
}

// Fast-path activation for callers that have established (e.g. by a warp vote on bnn_act_needs_care) that the
// argument of the exponential is finite and below 708 in magnitude: no clamp, no NaN fix-up -- the integer
// instructions of those two are a third of the slow path's issue slots.
template <int ACT>
__device__ __forceinline__ bool bnn_act_needs_care(double z) {
  if (ACT == BNN_ACT_RELU || ACT == BNN_ACT_LEAKY) return false;
  // swish: exp(-z), |z| < 708 ; tanh: exp(2z), |z| < 354 (inf / NaN compare as large)
  return (__double2hiint(z) & 0x7fffffff) >= (ACT == BNN_ACT_SWISH ? 0x40862000 : 0x40762000);
}
template <int ACT>
__device__ __forceinline__ double bnn_act_fast(double z, double alpha, const double* __restrict__ tab) {
  if (ACT == BNN_ACT_RELU) return z < 0.0 ? 0.0 : z;
  if (ACT == BNN_ACT_LEAKY) return z < 0.0 ? alpha * z : z;
  if (ACT == BNN_ACT_SWISH) return z * bnn_rcp(1.0 + bnn_exp_scaled<BNN_EXP_TAB_BITS, -1>(z, tab));
  return fma(-2.0, bnn_rcp(bnn_exp_scaled<BNN_EXP_TAB_BITS, 2>(z, tab) + 1.0), 1.0);
}

// ------------------------------------------------------------------------------------------------
// The same activations for N elements at once, cut into dependency LEVELS ("stages"): stage S of all N elements is
// N independent instructions, and consecutive stages are what the FP64 latency (8 clk) separates.  The caller places
// the stages BETWEEN its MMAs in program order; nvcc emits PTX in that order and ptxas largely keeps it, which is
// the only way to get the activation chains interleaved with the DMMAs instead of scheduled as one long FP64-only
// stretch (measured: what ptxas does with sequentially written activations).  Fast path only (bnn_act_fast: the
// caller has voted that no element needs the clamp / NaN fix-up); ReLU / leaky do everything in stage 0.
// ------------------------------------------------------------------------------------------------
template <int ACT, int N>
struct ActPipe {
  static constexpr int STAGES = (ACT == BNN_ACT_SWISH || ACT == BNN_ACT_TANH) ? 11 : 1;
  static constexpr int TB = BNN_EXP_TAB_BITS;
  double z[N], a[N], b[N], T[N];     // a, b: the two live temporaries of the chain
  int k[N];
  template <int S>
  __device__ __forceinline__ void stage(double alpha, const double* __restrict__ tab) { stage_range<S>(0, N, alpha, tab); }
  // elements [i0, i1) only (compile-time bounds after unrolling): per-chain slopes of the leaky ReLU
  template <int S>
  __device__ __forceinline__ void stage_range(int i0, int i1, double alpha, const double* __restrict__ tab) {
    static_assert(TB == 11, "staged activations use the 2048-entry table");
    constexpr double MAGIC = 6755399441055744.0, INV = 2954.639443740597, C1 = 0.0003384507717577858;
    constexpr bool SW = (ACT == BNN_ACT_SWISH);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i < i0 || i >= i1) continue;
      if constexpr (ACT == BNN_ACT_RELU) {
        if (S == 0) z[i] = z[i] < 0.0 ? 0.0 : z[i];
      } else if constexpr (ACT == BNN_ACT_LEAKY) {
        if (S == 0) z[i] = z[i] < 0.0 ? alpha * z[i] : z[i];
      } else {
        if (S == 0) a[i] = fma(z[i], SW ? -INV : 2.0 * INV, MAGIC);                  // t
        if (S == 1) {
          k[i] = __double2loint(a[i]);
          T[i] = tab[k[i] & ((1 << TB) - 1)];
          a[i] = a[i] - MAGIC;                                                        // kd
        }
        if (S == 2) a[i] = fma(a[i], SW ? C1 : -0.5 * C1, z[i]);                      // rs = -r (swish) or r/2 (tanh)
        if (S == 3) {
          b[i] = SW ? fma(a[i], -1.66666666666666657e-01, 0.5) : fma(a[i], 6.66666666666666630e-01, 1.0);   // q
          T[i] = T[i];
        }
        if (S == 4) {
          const double r2 = a[i] * a[i];
          a[i] = SW ? -a[i] : a[i];                                                   // folded into the FMA below by ptxas
          b[i] = fma(r2, b[i], a[i]);                                                 // p
        }
        if (S == 5) {
          const double T2 = SW ? T[i] : __hiloint2double(__double2hiint(T[i]) + (1 << 20), __double2loint(T[i]));
          const double res = fma(T2, b[i], T[i]);
          a[i] = __hiloint2double(__double2hiint(res) + ((k[i] >> TB) << 20), __double2loint(res));   // exp(..)
        }
        if (S == 6) {
          a[i] = a[i] + 1.0;                                                          // d
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(b[i]) : "d"(a[i]));                 // y0
        }
        if (S == 7) a[i] = fma(-a[i], b[i], 1.0);                                     // e
        if (S == 8) a[i] = fma(a[i], a[i], a[i]);
        if (S == 9) b[i] = fma(b[i], a[i], b[i]);                                     // y
        if (S == 10) z[i] = SW ? z[i] * b[i] : fma(-2.0, b[i], 1.0);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (free-running proposals, posterior-predictive resampling)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {   // [0,1)
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// uniform of (row, weight set) for the posterior-predictive resampling (sample_from_categorical, BNN_lib.py:682-713)
__device__ __forceinline__ double bnn_samp_uniform(unsigned long long seed, long long row, int set) {
  const uint4 r = philox4x32(make_uint4((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)set, 9u),
                             make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return u53(r.x, r.y);
}

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA 1-D) helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// wait with back-off: for a control thread that would otherwise spin on the issue slots of the warps it feeds
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
    __nanosleep(32);
  }
}
// global -> shared bulk copy, completion signalled on `bar` (bytes multiple of 16, 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// forward / likelihood kernel parameters
// ---------------------------------------------------------------------------------------------
#define BNN_MAX_SETS_PER_PASS 64

struct FwdParams {
  NetGeom g;
  const double* x;          // swizzled rows
  long long n_train, n_total, n_tiles16;
  const int* labels;        // [n_total] (categorical)
  const double* targets;    // [n_total, K] (Gaussian)
  const double* inst_w;     // [n_train] or null
  const double* class_w;    // [K] or null
  const double* wp;         // packed weight sets [C, PB]
  const double* alpha;      // [C, L] or null
  int C;
  // likelihood mode outputs
  double* part;             // [n_part, C, NF] per-warp partial sums (fixed order -> deterministic)
  int NF;                   // 1 (categorical / head: loglik) + Gaussian: 3*K (sum r, sum r^2, sum r^2 test)
  int* counts;              // [C, 2 + 2K] (atomics on integers: order independent)
  // prediction mode outputs
  double* mean_out;         // [n, K] or null
  double* votes_out;        // [n, K] or null
  double* dense_out;        // [C, n, K] or null
  double inv_sets;         // number of weight sets as a double: summaries are divided by it
  // posterior-predictive resampling (sample_from_categorical, BNN_lib.py:682-713); samp_u null = off
  const double* samp_u;     // [n, C] injected uniforms (parity with the reference's stream), or null
  int samp_philox;          // 1: uniforms are generated in the kernel, counter (row, set), key samp_seed -- O(1) memory
  unsigned long long samp_seed;
  int* samp_counts;         // [C, K] instances per class and set, or null
  double* samp_dense;       // [n, C] drawn class, or null   (the per-row shares go to votes_out)
  const double* exp_tab;    // [BNN_EXP_TAB_SIZE] 2^(j / BNN_EXP_TAB_SIZE)
  const double* exp_tab_small;   // [256] 2^(j / 256)
  // tensor-core first layer (k_fwd3t): int8 slice tiles of X, per-row scales, per-set slices of W1; null = off
  const uint8_t* xsl;
  const double* x_rowscale;
  const uint8_t* wt;
  long long n_tiles128;
  // block-masked networks (create_mask, BNN_lib.py:16-47): dataflow program over the dense blocks that cover
  // the mask of the hidden layers (format: bnn_forward.cu, k_fwd_sparse); null = dense evaluation
  const int* sp_prog;
  const int* sp_widx;       // [sp_wlen] offset inside a packed weight set of each weight-stream entry (-1: 0.0)
  int sp_prog_len, sp_n_items, sp_wlen;
  int sp_slots;             // scratch slots (hidden units alive at the same time) per row and chain
  int sp_group;             // chains whose weight streams are resident in shared memory (set by the launcher)
  int sp_share;             // k_fwd_pairs: warps that share one X tile (set by the launcher)
  int sp_pair_nc1, sp_pair_nr1, sp_pair_nr2;   // uniform block pairs (sparse_pair): features / units / units per pair; nc1 == 0: not uniform
};
