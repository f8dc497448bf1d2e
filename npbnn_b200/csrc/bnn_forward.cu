// K1 / K3: chain-batched forward pass fused with likelihood, accuracy counters or posterior summaries.
//
// Replaces (reference file:line): RunHiddenLayer / MatrixMultiplicationD / ActFun.eval
// (BNN_lib.py:184-193, 154-162, 83-87), SoftMax / RegressTransform(Error) (BNN_lib.py:166-182),
// calc_likelihood* (BNN_lib.py:100-143), CalcAccuracy / CalcLabelAccuracy / CalcLabelFreq
// (BNN_lib.py:195-233) as called from MCMC.mh_step (BNN_env.py:449-518), and the RunPredict loops of
// get_posterior_cat_prob / get_posterior_est / get_pdp (BNN_lib.py:376-392, 731-737, BNN_pdp.py:63-73).
//
// Work decomposition: a warp owns 16 rows of X ("warp tile") and runs every weight set of the pass
// over them, so X is read from HBM exactly once per pass.  Each layer is a chain of FP64 tensor-core
// MMAs (m16n8k8, bnn_common.cuh) whose accumulators are the next layer's A operand.
//   k_fwd3        3-layer nets with compile-time padded widths: X warp tile in shared memory (bulk
//                 async copy), weight sets streamed through a 2-deep shared-memory ring by bulk async
//                 copies + mbarriers, hidden activations never leave registers.
//   k_fwd_generic any depth/width: hidden activations staged in per-warp shared memory, weights read
//                 through L1/L2.
#include <cstdio>
#include <type_traits>
#include "bnn_common.cuh"
#include "bnn_generic_body.cuh"

// =============================================================================================
// generic kernel
// =============================================================================================
constexpr int GEN_WARPS = 4;

template <int ACT, bool PREDICT>
__global__ void __launch_bounds__(GEN_WARPS * 32) k_fwd_generic(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const NetGeom& g = p.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2;
  const int ZS = g.l[g.L - 1].out_pad + 1;        // the last layer may be padded beyond round_up(O, 8) (k_fwd3 width families)
  const int PW = bnn_pred_width(g);

  double* tab = reinterpret_cast<double*>(smem_raw);
  double* hbase = tab + GEN_TAB_SIZE;
  const int per_warp = 2 * 16 * g.max_w + 16 * ZS + (PREDICT ? 16 * PW : 0);
  double* h0 = hbase + warp * per_warp;
  double* h1 = h0 + 16 * g.max_w;
  double* zs = h1 + 16 * g.max_w;
  double* pacc = zs + 16 * ZS;
  int* ibase = reinterpret_cast<int*>(hbase + GEN_WARPS * per_warp);
  int* pvote = ibase + warp * (PREDICT ? 16 * PW : 0);
  int* cnt = ibase;   // likelihood mode: [C][2+2K]
  const int n_cnt = (!PREDICT && g.lik == BNN_LIK_CATEGORICAL) ? p.C * (2 + 2 * g.K) : 0;

  for (int i = threadIdx.x; i < GEN_TAB_SIZE; i += blockDim.x) tab[i] = p.exp_tab_small[i];
  for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) cnt[i] = 0;
  __syncthreads();

  // likelihood mode: when there are fewer row tiles than the GPU has room for, the weight sets are split over
  // gridDim.y as well (every (tile, set) still has exactly one owner); prediction accumulates over sets per tile
  const int c_beg = PREDICT ? 0 : (int)((long long)p.C * blockIdx.y / gridDim.y);
  const int c_end = PREDICT ? p.C : (int)((long long)p.C * (blockIdx.y + 1) / gridDim.y);
  const long long total_warps = (long long)gridDim.x * GEN_WARPS;
  for (long long wt = (long long)blockIdx.x * GEN_WARPS + warp; wt < p.n_tiles16; wt += total_warps) {
    if (PREDICT) {
      for (int i = lane; i < 16 * PW; i += 32) { pacc[i] = 0.0; pvote[i] = 0; }
      __syncwarp();
    }
    const double* xrow0 = p.x + (wt * 16 + gq) * (long long)g.F_pad;
    const double* xrow1 = xrow0 + 8LL * g.F_pad;
    for (int c = c_beg; c < c_end; ++c) {
      fwd_generic_layers<ACT, true, true>(g, xrow0, xrow1, p.wp + (long long)c * g.PB,
                                    (ACT == BNN_ACT_LEAKY && p.alpha) ? p.alpha + (long long)c * g.L : nullptr, h0, h1, zs, ZS,
                                    tab, lane);
      bnn_epilogue<PREDICT, false, GEN_TB>(p, c, wt, lane, zs, ZS, tab, cnt, pacc, pvote);
      __syncwarp();
    }
    bnn_pred_flush<PREDICT>(p, wt, lane, pacc, pvote);
  }
  if (n_cnt) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_cnt; i += blockDim.x)
      if (cnt[i]) atomicAdd(&p.counts[i], cnt[i]);
  }
}

// =============================================================================================
// block-masked networks (create_mask, BNN_lib.py:16-47; apply_mask, BNN_env.py:259-267)
// =============================================================================================
// A masked hidden layer is a list of small dense blocks ("items": rows r0..r0+nr-1, nr <= 4, columns
// c0..c0+nc-1) that cover every entry the mask keeps; entries outside the blocks are exactly zero in every
// proposal (w' *= mask, BNN_env.py:461-462), so skipping them changes nothing but the order of the sum.
// At the block sizes create_mask produces (a feature feeds a handful of nodes) the contraction is far too
// thin for tensor-core tiles and the FP64 work is dominated by the activations, so this kernel is plain DFMA
// with one THREAD per row: lane = row of a 32-row tile, every lane runs the same item, weights are warp-
// uniform shared-memory loads.  The host (build_sparse_program, bnn_capi.cu) orders the items of all hidden
// layers as a dataflow program -- an item runs as soon as the units it reads exist -- and assigns each hidden
// unit a scratch slot that is recycled after its last reader, so a thread needs F + (live units) doubles of
// shared memory instead of every layer's width; the units of the LAST hidden layer are never stored: they are
// multiplied into the (dense, O <= 8) output layer's accumulators, which stay in registers.
//   * the weights the program touches are gathered once per CTA from the packed weight sets into shared
//     memory IN PROGRAM ORDER (sp_widx), so the item loop reads them with a running pointer and 16-byte
//     loads -- no address arithmetic per multiply-add;
//   * four (or two) chains are evaluated together (same program, different weights) for instruction-level
//     parallelism;
//   * a work unit is (32-row tile, chain subset); a CTA owns a contiguous range of tiles and its warps draw
//     units from a shared counter, so the load is balanced to one unit.
//
// program (ints), per item: [0] layer | nr << 8 | to_out << 16  [1] nc  [2] 1 if the inputs are scratch slots
//     (0: X columns, shared by both chains)  [3] 0  [4..7] element offset of each produced unit's slot
//     [8..8+nc) element offset of each input, padded to a multiple of 4 ints (offsets are for the first chain of the unit; chain q's
//     scratch slots follow q * sp_slots * SP_US elements later)
// weight stream (doubles), per chain: [0..8) output-layer bias; then per item: bias[nrp], nc x w[nrp] (column
//     c0+k, rows r0..r0+nr-1) and, for to_out items, nr x w_out[op] (the output-layer column of each produced
//     unit); nrp / op = nr / O rounded up to even, padding is 0.
constexpr int SP_US = 33;        // per-unit stride in doubles (32 rows + 1: conflict-free transposed stores)
constexpr int SP_MAX_O = 8;
constexpr int SP_HDR = 8;

template <int ACT, int NCH, int NR>
__device__ __forceinline__ void sparse_item(const int* it, int nc, int qs, int sq, bool to_out, int O,
                                            const double* (&w)[NCH], double* bufl, const double (&alpha)[NCH],
                                            const double* tab, double (&out)[NCH][SP_MAX_O]) {
  constexpr int NRP = (NR + 1) & ~1;
  double acc[NCH][NRP];
#pragma unroll
  for (int q = 0; q < NCH; ++q)
#pragma unroll
    for (int i = 0; i < NRP; i += 2) {
      const double2 b = *reinterpret_cast<const double2*>(w[q] + i);
      acc[q][i] = b.x; acc[q][i + 1] = b.y;
    }
#pragma unroll 2
  for (int k = 0; k < nc; ++k) {
    const int src = it[SP_HDR + k];
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
      const double a = bufl[src + q * qs];
#pragma unroll
      for (int i = 0; i < NRP; i += 2) {
        const double2 ww = *reinterpret_cast<const double2*>(w[q] + (k + 1) * NRP + i);
        acc[q][i] = fma(ww.x, a, acc[q][i]);
        if (i + 1 < NR) acc[q][i + 1] = fma(ww.y, a, acc[q][i + 1]);
      }
    }
  }
  bool care = false;
#pragma unroll
  for (int q = 0; q < NCH; ++q)
#pragma unroll
    for (int i = 0; i < NR; ++i) care |= bnn_act_needs_care<ACT>(acc[q][i]);
  if (!__any_sync(FULL_MASK, care)) {
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int i = 0; i < NR; ++i) acc[q][i] = bnn_act_fast<ACT>(acc[q][i], alpha[q], tab);
  } else {
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int i = 0; i < NR; ++i) acc[q][i] = bnn_act<ACT>(acc[q][i], alpha[q], tab);
  }
  int adv = (nc + 1) * NRP;
  if (to_out) {
    const int op = (O + 1) & ~1;
#pragma unroll
    for (int i = 0; i < NR; ++i) {
#pragma unroll
      for (int j = 0; j < SP_MAX_O / 2; ++j) {
        if (2 * j < O) {
#pragma unroll
          for (int q = 0; q < NCH; ++q) {
            const double2 ww = *reinterpret_cast<const double2*>(w[q] + adv + i * op + 2 * j);
            out[q][2 * j] = fma(ww.x, acc[q][i], out[q][2 * j]);
            out[q][2 * j + 1] = fma(ww.y, acc[q][i], out[q][2 * j + 1]);
          }
        }
      }
    }
    adv += NR * op;
  } else {
    const int4 sl = *reinterpret_cast<const int4*>(it + 4);
    const int slot[4] = {sl.x, sl.y, sl.z, sl.w};
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int i = 0; i < NR; ++i) bufl[slot[i] + q * sq] = acc[q][i];
  }
#pragma unroll
  for (int q = 0; q < NCH; ++q) w[q] += adv;
}

// Fused evaluation of one "block pair" for networks whose program is a uniform sequence of pairs (what create_mask
// produces when every feature group has the same node counts, BASELINE config 3: 40 x [1 feature -> 3 nodes -> 2
// nodes]): a first-layer item (NR1 units reading nc1 features) directly followed by the second-layer item that reads
// exactly those units (NR2 units, feeding the output layer).  Same program, same weight stream as the interpreter
// (sparse_item), but the hidden units stay in registers (no scratch slots), the item headers are not decoded and
// there is no dispatch on the item height -- the interpreter spends two thirds of its instructions on those.
template <int ACT, int NCH, int NR1, int NR2, int NC1 = 0>
__device__ __forceinline__ void sparse_pair(const int* it, int nc1_rt, int O, const double* (&w)[NCH], const double* bufl,
                                            const double (&alpha1)[NCH], const double (&alpha2)[NCH], const double* tab,
                                            double (&out)[NCH][SP_MAX_O]) {
  constexpr int P1 = (NR1 + 1) & ~1, P2 = (NR2 + 1) & ~1;
  double h1[NCH][P1], h2[NCH][P2];
#pragma unroll
  for (int q = 0; q < NCH; ++q)
#pragma unroll
    for (int i = 0; i < P1; i += 2) {
      const double2 b = *reinterpret_cast<const double2*>(w[q] + i);
      h1[q][i] = b.x; h1[q][i + 1] = b.y;
    }
  const int nc1 = NC1 > 0 ? NC1 : nc1_rt;                    // NC1 > 0: features per pair known at compile time
#pragma unroll
  for (int k = 0; k < nc1; ++k) {
    const double a = bufl[it[SP_HDR + k]];                   // the feature is shared by the chains
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int i = 0; i < P1; i += 2) {
        const double2 ww = *reinterpret_cast<const double2*>(w[q] + (k + 1) * P1 + i);
        h1[q][i] = fma(ww.x, a, h1[q][i]);
        if (i + 1 < NR1) h1[q][i + 1] = fma(ww.y, a, h1[q][i + 1]);
      }
  }
  const int adv1 = (nc1 + 1) * P1;
  bool care = false;
#pragma unroll
  for (int q = 0; q < NCH; ++q)
#pragma unroll
    for (int i = 0; i < NR1; ++i) care |= bnn_act_needs_care<ACT>(h1[q][i]);
  if (!__any_sync(FULL_MASK, care)) {
#ifndef SPARSE_STAGED_ACT
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int i = 0; i < NR1; ++i) h1[q][i] = bnn_act_fast<ACT>(h1[q][i], alpha1[q], tab);
#else
    // all NCH * NR1 evaluations level by level (ActPipe): the program order ptxas gets is already interleaved
    ActPipe<ACT, NCH * NR1> ap;
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int i = 0; i < NR1; ++i) ap.z[q * NR1 + i] = h1[q][i];
    static_for<0, ActPipe<ACT, NCH * NR1>::STAGES>([&](auto st) {
      if constexpr (ACT == BNN_ACT_LEAKY) {
#pragma unroll
        for (int q = 0; q < NCH; ++q) ap.template stage_range<decltype(st)::value>(q * NR1, (q + 1) * NR1, alpha1[q], tab);
      } else {
        ap.template stage<decltype(st)::value>(0.0, tab);
      }
    });
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int i = 0; i < NR1; ++i) h1[q][i] = ap.z[q * NR1 + i];
#endif
  } else {
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int i = 0; i < NR1; ++i) h1[q][i] = bnn_act<ACT>(h1[q][i], alpha1[q], tab);
  }
  // second-layer item: bias[P2], then one column of P2 weights per first-layer unit
#pragma unroll
  for (int q = 0; q < NCH; ++q)
#pragma unroll
    for (int j = 0; j < P2; j += 2) {
      const double2 b = *reinterpret_cast<const double2*>(w[q] + adv1 + j);
      h2[q][j] = b.x; h2[q][j + 1] = b.y;
    }
#pragma unroll
  for (int i = 0; i < NR1; ++i)
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int j = 0; j < P2; j += 2) {
        const double2 ww = *reinterpret_cast<const double2*>(w[q] + adv1 + (i + 1) * P2 + j);
        h2[q][j] = fma(ww.x, h1[q][i], h2[q][j]);
        if (j + 1 < NR2) h2[q][j + 1] = fma(ww.y, h1[q][i], h2[q][j + 1]);
      }
  care = false;
#pragma unroll
  for (int q = 0; q < NCH; ++q)
#pragma unroll
    for (int j = 0; j < NR2; ++j) care |= bnn_act_needs_care<ACT>(h2[q][j]);
  if (!__any_sync(FULL_MASK, care)) {
#ifndef SPARSE_STAGED_ACT
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int j = 0; j < NR2; ++j) h2[q][j] = bnn_act_fast<ACT>(h2[q][j], alpha2[q], tab);
#else
    ActPipe<ACT, NCH * NR2> ap;
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int j = 0; j < NR2; ++j) ap.z[q * NR2 + j] = h2[q][j];
    static_for<0, ActPipe<ACT, NCH * NR2>::STAGES>([&](auto st) {
      if constexpr (ACT == BNN_ACT_LEAKY) {
#pragma unroll
        for (int q = 0; q < NCH; ++q) ap.template stage_range<decltype(st)::value>(q * NR2, (q + 1) * NR2, alpha2[q], tab);
      } else {
        ap.template stage<decltype(st)::value>(0.0, tab);
      }
    });
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int j = 0; j < NR2; ++j) h2[q][j] = ap.z[q * NR2 + j];
#endif
  } else {
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
      for (int j = 0; j < NR2; ++j) h2[q][j] = bnn_act<ACT>(h2[q][j], alpha2[q], tab);
  }
  const int adv2 = adv1 + (NR1 + 1) * P2;
  const int op = (O + 1) & ~1;
#pragma unroll
  for (int j = 0; j < NR2; ++j)
#pragma unroll
    for (int o = 0; o < SP_MAX_O / 2; ++o)
      if (2 * o < O) {
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
          const double2 ww = *reinterpret_cast<const double2*>(w[q] + adv2 + j * op + 2 * o);
          out[q][2 * o] = fma(ww.x, h2[q][j], out[q][2 * o]);
          out[q][2 * o + 1] = fma(ww.y, h2[q][j], out[q][2 * o + 1]);
        }
      }
#pragma unroll
  for (int q = 0; q < NCH; ++q) w[q] += adv2 + NR2 * op;
}

// NCH chains per work unit: 4 (at most 10 warps, <= 204 registers) or 2 (at most 16 warps, <= 128 registers).
template <int ACT, int NCH>
__global__ void __launch_bounds__(NCH == 4 ? 320 : 512, 1) k_fwd_sparse(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const NetGeom& g = p.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int ZS = SP_MAX_O + 1;
  const int O = g.O;
  const int G = p.sp_group;                                   // chains whose weight streams are resident
  const int sq = p.sp_slots * SP_US;

  double* tab = reinterpret_cast<double*>(smem_raw);
  double* wsm = tab + BNN_EXP_TAB_SIZE;                       // [G][sp_wlen]
  const int per_warp = (g.F + NCH * p.sp_slots) * SP_US + 32 * ZS;
  double* buf = wsm + (size_t)G * p.sp_wlen + warp * per_warp;
  double* zs = buf + (g.F + NCH * p.sp_slots) * SP_US;
  int* cnt = reinterpret_cast<int*>(wsm + (size_t)G * p.sp_wlen + nwarps * per_warp);
  const int n_cnt = (g.lik == BNN_LIK_CATEGORICAL) ? p.C * (2 + 2 * g.K) : 0;
  int* next_unit = cnt + n_cnt;
  int* prog = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(next_unit + 1) + 15) & ~(uintptr_t)15);

  for (int i = threadIdx.x; i < BNN_EXP_TAB_SIZE; i += blockDim.x) tab[i] = p.exp_tab[i];
  for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) cnt[i] = 0;
  for (int i = threadIdx.x; i < p.sp_prog_len; i += blockDim.x) prog[i] = p.sp_prog[i];

  // this CTA's contiguous range of 32-row tiles
  const long long n_tiles32 = (p.n_tiles16 + 1) >> 1;
  const long long n_pad = p.n_tiles16 * 16;
  const long long t0 = n_tiles32 * blockIdx.x / gridDim.x, t1 = n_tiles32 * (blockIdx.x + 1) / gridDim.x;
  double* bufl = buf + lane;

  for (int c0 = 0; c0 < p.C; c0 += G) {
    const int gc = min(G, p.C - c0);                          // chains of this group
    const int n_sub = (gc + NCH - 1) / NCH;                   // chain subsets of NCH
    __syncthreads();                                           // previous group's readers are done
    for (int i = threadIdx.x; i < gc * p.sp_wlen; i += blockDim.x) {
      const int c = i / p.sp_wlen, j = i - c * p.sp_wlen;
      const int idx = __ldg(p.sp_widx + j);
      wsm[i] = idx < 0 ? 0.0 : __ldg(p.wp + (long long)(c0 + c) * g.PB + idx);
    }
    if (threadIdx.x == 0) *next_unit = 0;
    __syncthreads();
    const int n_units = (int)(t1 - t0) * n_sub;
    long long cur_tile = -1;
    for (;;) {
      int u = 0;
      if (lane == 0) u = atomicAdd(next_unit, 1);
      u = __shfl_sync(FULL_MASK, u, 0);
      if (u >= n_units) break;
      const long long w32 = t0 + u / n_sub;
      const int sub = u - (int)(w32 - t0) * n_sub;
      if (w32 != cur_tile) {
        // X tile -> [feature][row]; global reads are contiguous (32 * F_pad consecutive doubles)
        const double* xt = p.x + w32 * 32 * (long long)g.F_pad;
        __syncwarp();
        // 8 independent loads in flight per lane (F_pad is a multiple of 8)
        for (int j0 = 0; j0 < g.F_pad; j0 += 8) {
          double v[8];
#pragma unroll
          for (int uu = 0; uu < 8; ++uu) {
            const int idx = (j0 + uu) * 32 + lane;
            v[uu] = (w32 * 32 + idx / g.F_pad < n_pad) ? __ldg(xt + idx) : 0.0;
          }
#pragma unroll
          for (int uu = 0; uu < 8; ++uu) {
            const int idx = (j0 + uu) * 32 + lane;
            const int r = idx / g.F_pad, cp = idx - r * g.F_pad;
            const int cidx = cp ^ ((r & 1) * g.x_swz);
            if (cidx < g.F) buf[cidx * SP_US + r] = v[uu];
          }
        }
        __syncwarp();
        cur_tile = w32;
      }
      // chains of this unit (a short last subset repeats its last chain; the copies are not scored)
      int ch[NCH];
      const double* w[NCH];
      double out[NCH][SP_MAX_O];
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        ch[q] = min(NCH * sub + q, gc - 1);
        w[q] = wsm + (size_t)ch[q] * p.sp_wlen;
#pragma unroll
        for (int o = 0; o < SP_MAX_O; ++o) out[q][o] = w[q][o];
        w[q] += SP_MAX_O;
      }
      const int* it = prog;
      for (int n = 0; n < p.sp_n_items; ++n) {
        const int4 h = *reinterpret_cast<const int4*>(it);
        const int layer = h.x & 0xff, nr = (h.x >> 8) & 0xff, nc = h.y;
        const bool to_out = (h.x >> 16) & 1;
        const int qs = h.z ? sq : 0;
        double alpha[NCH];
#pragma unroll
        for (int q = 0; q < NCH; ++q)
          alpha[q] = (ACT == BNN_ACT_LEAKY && p.alpha) ? p.alpha[(c0 + ch[q]) * g.L + layer] : 0.0;
        switch (nr) {
          case 1: sparse_item<ACT, NCH, 1>(it, nc, qs, sq, to_out, O, w, bufl, alpha, tab, out); break;
          case 2: sparse_item<ACT, NCH, 2>(it, nc, qs, sq, to_out, O, w, bufl, alpha, tab, out); break;
          case 3: sparse_item<ACT, NCH, 3>(it, nc, qs, sq, to_out, O, w, bufl, alpha, tab, out); break;
          default: sparse_item<ACT, NCH, 4>(it, nc, qs, sq, to_out, O, w, bufl, alpha, tab, out); break;
        }
        it += SP_HDR + ((nc + 3) & ~3);         // items are padded to 16 bytes
      }
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        if (NCH * sub + q < gc) {
          __syncwarp();
#pragma unroll
          for (int o = 0; o < SP_MAX_O; ++o) zs[lane * ZS + o] = out[q][o];
          __syncwarp();
          bnn_epilogue<false, true>(p, c0 + ch[q], w32 * 2, lane, zs, ZS, tab, cnt, nullptr, nullptr);
        }
      }
    }
  }
  if (n_cnt) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_cnt; i += blockDim.x)
      if (cnt[i]) atomicAdd(&p.counts[i], cnt[i]);
  }
}

// Uniform block pairs (sparse_pair) for all chains of the pass.  Thread = row as in k_fwd_sparse, but the in-flight
// work comes from MANY warps with few chains each instead of few warps with four chains each: groups of GW warps
// share one transposed X tile (the tile's rows are the same for every chain subset), so a warp's private shared
// memory is only the 32 x 9 staging area of the likelihood epilogue and 16-24 warps fit next to the weight streams.
//   NCH 1: up to 24 warps (<= 80 registers), 2: up to 16 warps (<= 128 registers)
// Measured at BASELINE config 3 (8 chains, 200k rows): interpreter 0.90 ms per step, pairs with 4 chains x 10 warps
// 0.69, 1 chain x 24 warps 0.67, 2 chains x 16 warps 0.63 (issue slots 65 % busy, FP64 pipe 52 %).
template <int ACT, int NCH, int NR1, int NR2, int NC1>
__global__ void __launch_bounds__(NCH == 1 ? 768 : 512, 1) k_fwd_pairs(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const NetGeom& g = p.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int ZS = SP_MAX_O + 1;
  const int O = g.O;
  const int G = p.sp_group;                                   // chains whose weight streams are resident
  const int GW = p.sp_share;                                   // warps per group (share one X tile)
  const int n_groups = nwarps / GW, grp = warp / GW, wi = warp - grp * GW;

  double* tab = reinterpret_cast<double*>(smem_raw);
  double* wsm = tab + BNN_EXP_TAB_SIZE;                       // [G][sp_wlen]
  double* tile = wsm + (size_t)G * p.sp_wlen + (size_t)grp * g.F * SP_US;       // [n_groups][F][33]
  double* zs = wsm + (size_t)G * p.sp_wlen + (size_t)n_groups * g.F * SP_US + (size_t)warp * 32 * ZS;
  int* cnt = reinterpret_cast<int*>(wsm + (size_t)G * p.sp_wlen + (size_t)n_groups * g.F * SP_US + (size_t)nwarps * 32 * ZS);
  const int n_cnt = (g.lik == BNN_LIK_CATEGORICAL) ? p.C * (2 + 2 * g.K) : 0;
  int* prog = cnt + ((n_cnt + 3) & ~3);                        // 16-byte aligned: the areas before it are

  for (int i = threadIdx.x; i < BNN_EXP_TAB_SIZE; i += blockDim.x) tab[i] = p.exp_tab[i];
  for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) cnt[i] = 0;
  for (int i = threadIdx.x; i < p.sp_prog_len; i += blockDim.x) prog[i] = p.sp_prog[i];

  const long long n_tiles32 = (p.n_tiles16 + 1) >> 1;
  const long long n_pad = p.n_tiles16 * 16;
  const long long t0 = n_tiles32 * blockIdx.x / gridDim.x, t1 = n_tiles32 * (blockIdx.x + 1) / gridDim.x;
  const double* bufl = tile + lane;
  const int nc1 = p.sp_pair_nc1;
  const int stride = 2 * SP_HDR + ((nc1 + 3) & ~3) + ((NR1 + 3) & ~3);          // two padded items per pair
  const int gthreads = GW * 32, gtid = wi * 32 + lane;
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(gthreads) : "memory"); };

  for (int c0 = 0; c0 < p.C; c0 += G) {
    const int gc = min(G, p.C - c0);                          // chains of this group
    const int n_sub = (gc + NCH - 1) / NCH;                   // chain subsets of NCH
    __syncthreads();                                           // previous group's readers are done
    for (int i = threadIdx.x; i < gc * p.sp_wlen; i += blockDim.x) {
      const int c = i / p.sp_wlen, j = i - c * p.sp_wlen;
      const int idx = __ldg(p.sp_widx + j);
      wsm[i] = idx < 0 ? 0.0 : __ldg(p.wp + (long long)(c0 + c) * g.PB + idx);
    }
    __syncthreads();
    if (grp >= n_groups) continue;                             // (left-over warps when nwarps is not a multiple of GW)
    for (long long w32 = t0 + grp; w32 < t1; w32 += n_groups) {
      // X tile -> [feature][row], loaded by the whole group; global reads are contiguous (32 * F_pad doubles)
      const double* xt = p.x + w32 * 32 * (long long)g.F_pad;
      group_sync();                                            // the previous tile's readers are done
      for (int e0 = gtid; e0 < 32 * g.F_pad; e0 += 4 * gthreads) {
        double v[4];
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int idx = e0 + uu * gthreads;
          v[uu] = (idx < 32 * g.F_pad && w32 * 32 + idx / g.F_pad < n_pad) ? __ldg(xt + idx) : 0.0;
        }
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int idx = e0 + uu * gthreads;
          const int r = idx / g.F_pad, cp = idx - r * g.F_pad;
          const int cidx = cp ^ ((r & 1) * g.x_swz);
          if (idx < 32 * g.F_pad && cidx < g.F) tile[cidx * SP_US + r] = v[uu];
        }
      }
      group_sync();
      for (int sub = wi; sub < n_sub; sub += GW) {
        int ch[NCH];
        const double* w[NCH];
        double out[NCH][SP_MAX_O];
        double alpha1[NCH], alpha2[NCH];
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
          ch[q] = min(NCH * sub + q, gc - 1);
          w[q] = wsm + (size_t)ch[q] * p.sp_wlen;
#pragma unroll
          for (int o = 0; o < SP_MAX_O; ++o) out[q][o] = w[q][o];
          w[q] += SP_MAX_O;
          alpha1[q] = (ACT == BNN_ACT_LEAKY && p.alpha) ? p.alpha[(c0 + ch[q]) * g.L + 0] : 0.0;
          alpha2[q] = (ACT == BNN_ACT_LEAKY && p.alpha) ? p.alpha[(c0 + ch[q]) * g.L + 1] : 0.0;
        }
        const int* it = prog;
        for (int n = 0; n < p.sp_n_items; n += 2, it += stride)
          sparse_pair<ACT, NCH, NR1, NR2, NC1>(it, nc1, O, w, bufl, alpha1, alpha2, tab, out);
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
          if (NCH * sub + q < gc) {
            __syncwarp();
#pragma unroll
            for (int o = 0; o < SP_MAX_O; ++o) zs[lane * ZS + o] = out[q][o];
            __syncwarp();
            bnn_epilogue<false, true>(p, c0 + ch[q], w32 * 2, lane, zs, ZS, tab, cnt, nullptr, nullptr);
          }
        }
      }
    }
  }
  if (n_cnt) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_cnt; i += blockDim.x)
      if (cnt[i]) atomicAdd(&p.counts[i], cnt[i]);
  }
}

// =============================================================================================
// specialised 3-layer kernel (compile-time padded widths KP0 -> N1 -> N2 -> N3), categorical likelihood
// =============================================================================================
template <int KP0, int N1, int N2, int N3>
struct Fwd3Geom {
  static constexpr int SW0 = (KP0 % 16 == 0) ? 8 : 0;
  static constexpr int SW1 = (N1 % 16 == 0) ? 8 : 0;
  static constexpr int SW2 = (N2 % 16 == 0) ? 8 : 0;
  static constexpr int W1_OFF = 0, B1_OFF = N1 * KP0;
  static constexpr int W2_OFF = B1_OFF + N1, B2_OFF = W2_OFF + N2 * N1;
  static constexpr int W3_OFF = B2_OFF + N2, B3_OFF = W3_OFF + N3 * N2;
  static constexpr int PB = B3_OFF + N3;
};

#ifndef FWD3_WARPS
#define FWD3_WARPS 12
#endif
#ifndef FWD3_ACT_MODE
#define FWD3_ACT_MODE 0         // how layer-1 activations interleave with the layer-2 MMAs (tuning variants 1, 2)
#endif

template <int ACT, bool FAST = false, int TB = BNN_EXP_TAB_BITS>
__device__ __forceinline__ void act_tile(double (&a)[4], double alpha, const double* tab) {
#ifdef BNN_DBG_NOACT      // tuning experiment only: how fast is the kernel without activations?
  return;
#endif
#ifdef BNN_DBG_ACTCHAIN   // tuning experiment only: N dependent register DFMAs per element instead of the activation
#pragma unroll
  for (int s = 0; s < BNN_DBG_ACTCHAIN; ++s)
#pragma unroll
    for (int e = 0; e < 4; ++e) a[e] = fma(a[e], 0.999, alpha);
  return;
#endif
#pragma unroll
  for (int e = 0; e < 4; ++e) a[e] = FAST ? bnn_act_fast<ACT>(a[e], alpha, tab) : bnn_act<ACT, TB>(a[e], alpha, tab);
}

// Does any pre-activation of this warp's fragments need the guarded evaluation (|z| beyond the exponential's range,
// inf, NaN)?  One vote per layer: the common case then runs a straight-line copy of the layer without the clamp /
// NaN fix-up instructions (6 integer instructions per element on the 16-lane ALU pipe, which the activation chains
// of three warps otherwise keep as busy as the FP64 pipe).
template <int ACT, int NT>
__device__ __forceinline__ bool frag_needs_care(const double (&acc)[NT][4]) {
  if (ACT == BNN_ACT_RELU || ACT == BNN_ACT_LEAKY) return false;
  int worst = 0;
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) worst = max(worst, __double2hiint(acc[j][e]) & 0x7fffffff);
  return __any_sync(FULL_MASK, worst >= (ACT == BNN_ACT_SWISH ? 0x40862000 : 0x40762000));
}

// ---------------------------------------------------------------------------------------------
// Epilogues on the accumulator fragments of the last layer.  The 4 lanes of a quad hold one row pair
// (rows g and g+8), columns 8j+2t+{0,1}; row statistics are combined with two xor-shuffles and all exps
// of a thread are independent (ILP).
// ---------------------------------------------------------------------------------------------
template <int N3>
struct RowStats {
  double m[2], S[2], zy[2];
  int arg[2];
  double ev[2][N3 / 4];
};

// stage 1: row maximum and arg-max (first maximum wins, as np.argmax)
template <int N3>
__device__ __forceinline__ void qs_max(const double (&acc)[N3 / 8][4], int K, int t, RowStats<N3>& r) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double m = -INFINITY;
    int arg = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < N3 / 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 8 * j + 2 * t + e;
        const double v = acc[j][2 * h + e];
        const bool take = col < K && v > m;      // columns ascend within a thread
        m = take ? v : m;
        arg = take ? col : arg;
      }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const double om = __shfl_xor_sync(FULL_MASK, m, o);
      const int oa = __shfl_xor_sync(FULL_MASK, arg, o);
      const bool take = om > m || (om == m && oa < arg);
      m = take ? om : m;
      arg = take ? oa : arg;
    }
    r.m[h] = m;
    r.arg[h] = (arg == 0x7fffffff) ? 0 : arg;
  }
}

// stage 2: exp(z - max) of this thread's columns of row h, local sums
template <int N3, int H, bool NEED_ZY, int TB = BNN_EXP_TAB_BITS>
__device__ __forceinline__ void qs_exp(const double (&acc)[N3 / 8][4], int K, int t, const int (&y)[2],
                                       const double* tab, RowStats<N3>& r) {
  double S = 0.0, zy = 0.0;
#pragma unroll
  for (int j = 0; j < N3 / 8; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 8 * j + 2 * t + e;
      const double v = acc[j][2 * H + e];
      // padded columns: argument -1000 => exp flushes to exactly 0 (no branch around the evaluation)
      const double ex = bnn_exp_neg<TB>((col < K) ? v - r.m[H] : -1000.0, tab);
      r.ev[H][2 * j + e] = ex;
      S += ex;
      if (NEED_ZY) zy = (col == y[H]) ? v : zy;
    }
  r.S[H] = S;
  r.zy[H] = zy;
}

// stage 3: quad reductions of the sums
template <int N3, bool NEED_ZY>
__device__ __forceinline__ void qs_reduce(RowStats<N3>& r) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    r.S[h] += __shfl_xor_sync(FULL_MASK, r.S[h], 1);
    r.S[h] += __shfl_xor_sync(FULL_MASK, r.S[h], 2);
    if (NEED_ZY) {
      r.zy[h] += __shfl_xor_sync(FULL_MASK, r.zy[h], 1);
      r.zy[h] += __shfl_xor_sync(FULL_MASK, r.zy[h], 2);
    }
  }
}

template <int N3, bool NEED_ZY, int TB = BNN_EXP_TAB_BITS>
__device__ __forceinline__ void quad_softmax_stats(const double (&acc)[N3 / 8][4], int K, int t, const int (&y)[2],
                                                   const double* tab, RowStats<N3>& r) {
  qs_max<N3>(acc, K, t, r);
  qs_exp<N3, 0, NEED_ZY, TB>(acc, K, t, y, tab, r);
  qs_exp<N3, 1, NEED_ZY, TB>(acc, K, t, y, tab, r);
  qs_reduce<N3, NEED_ZY>(r);
}

// Likelihood + accuracy counters of one (warp tile, weight set), in two parts so that the kernel can place
// the arithmetic (part 1: softmax statistics and log-likelihood, branch-free, latency-bound FP64 chains and
// shuffles) in the same basic block as the first MMAs of the NEXT weight set -- ptxas then interleaves the
// two -- and the side effects (part 2: counters and the partial-sum store) at the end of that block.
//   * without class / instance weights  sum_r log softmax = sum_r (z_y - max) - log(prod_r S_r): ONE log per
//     warp tile (the product of 16 sums of at most 32 terms each stays below 2^80)
//   * counters: ballots for n_correct, full-mask match.any groups for the per-class counters: at most one
//     shared-memory atomic per distinct class and warp tile instead of three per row
struct LikRow {
  double ll;          // warp-tile log-likelihood (all lanes)
  int my_y, my_arg;   // lane t == 0 speaks for row g, lane t == 1 for row g + 8
  bool ok, train, role;
};

// stage 4: per-row roles, warp-tile reduction and the logarithm
template <int N3, bool WEIGHTED>
__device__ __forceinline__ LikRow quad_lik_finish(const FwdParams& p, long long wt, int lane, const RowStats<N3>& r,
                                                  const int (&y)[2], const double (&wgt)[2], bool valid) {
  const int gq = lane >> 2, t = lane & 3;
  LikRow o;
  const int h = t & 1;
  const long long row = wt * 16 + gq + 8 * h;
  o.role = valid && t < 2 && row < p.n_total;
  o.train = o.role && row < p.n_train;
  o.my_y = h ? y[1] : y[0];
  o.my_arg = h ? r.arg[1] : r.arg[0];
  const double my_S = h ? r.S[1] : r.S[0];
  const double d = (h ? r.zy[1] : r.zy[0]) - (h ? r.m[1] : r.m[0]);
  o.ok = o.role && o.my_arg == o.my_y;
  // log(softmax) of the reference is -inf once exp(d) underflows to 0 (BNN_lib.py:121,168)
  const bool neg_inf = o.train && d < -745.1332191019412;
  double ll;
  if (!WEIGHTED) {
    double sd = o.train ? d : 0.0, ps = o.train ? my_S : 1.0;
#pragma unroll
    for (int s = 1; s <= 16; s <<= 1) {        // fixed xor tree => deterministic
      sd += __shfl_xor_sync(FULL_MASK, sd, s);
      ps *= __shfl_xor_sync(FULL_MASK, ps, s);
    }
    ll = sd - bnn_log_ge1(ps);
  } else {
    ll = o.train ? (d - bnn_log_ge1(my_S)) * (h ? wgt[1] : wgt[0]) : 0.0;
#pragma unroll
    for (int s = 1; s <= 16; s <<= 1) ll += __shfl_xor_sync(FULL_MASK, ll, s);
  }
  o.ll = __any_sync(FULL_MASK, neg_inf) ? -INFINITY : ll;
  return o;
}

template <int N3>
__device__ __forceinline__ void quad_lik_commit(const FwdParams& p, int c, long long wt, int lane, int* cnt,
                                                const LikRow& o, bool valid) {
  static_assert(N3 < 32, "one lane per class plus one for the totals");
  const int K = p.g.K;
  int* cc = cnt + c * (2 + 2 * K);
  // Counters without data-dependent control flow: one ballot per class and table (independent VOTEs that pipeline),
  // lane k keeps the count of class k, then ONE predicated shared-memory atomic per table with distinct addresses per
  // lane.  (Measured: the counters cost nothing -- a build without them runs in the same 17.3 ms.)
  const unsigned ok_train = __ballot_sync(FULL_MASK, o.ok && o.train);
  const unsigned ok_test = __ballot_sync(FULL_MASK, o.ok && !o.train);
  int n_y = 0, n_arg = 0;
#pragma unroll
  for (int k = 0; k < N3; ++k) {
    const unsigned by = __ballot_sync(FULL_MASK, o.ok && o.train && o.my_y == k);
    const unsigned ba = __ballot_sync(FULL_MASK, o.train && o.my_arg == k);
    if (lane == k) { n_y = __popc(by); n_arg = __popc(ba); }
  }
  if (lane == N3) { n_y = __popc(ok_train); n_arg = __popc(ok_test); }
  // lanes 0..K-1: class tables; lane N3: the two totals (cc[0], cc[1])
  const bool tab = lane < K;
  const int i_y = tab ? 2 + lane : 0, i_arg = tab ? 2 + K + lane : 1;
  if ((tab || lane == N3) && n_y) atomicAdd(&cc[i_y], n_y);
  if ((tab || lane == N3) && n_arg) atomicAdd(&cc[i_arg], n_arg);
  if (valid && lane == 0) p.part[((long long)c * p.NF) * p.n_tiles16 + wt] = o.ll;
}

// Posterior summaries: softmax probabilities accumulated per (row, class) in registers over the weight sets.
template <int N3>
__device__ __forceinline__ void quad_epilogue_pred(const FwdParams& p, int c, long long wt, int lane,
                                                   const double (&acc)[N3 / 8][4], const double* tab,
                                                   double (&pacc)[2][N3 / 4], int (&pvote)[2][N3 / 4]) {
  const int K = p.g.K;
  const int gq = lane >> 2, t = lane & 3;
  const int y[2] = {-1, -1};
  RowStats<N3> r;
  quad_softmax_stats<N3, false>(acc, K, t, y, tab, r);
  if (p.samp_philox) {
    // Posterior-predictive resampling (sample_from_categorical, BNN_lib.py:682-713) with in-kernel uniforms: the class of
    // (row, set) is the first one whose cumulative probability reaches u (none: class 0).  The classes of a row are
    // spread over the 4 lanes of a quad (columns 8j + 2t + e): each lane forms the cumulative sums of its own columns on
    // top of the quad-exclusive prefix of the 8-column block and the totals of the blocks before it.
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long row = wt * 16 + gq + 8 * h;
      const bool live = row < p.n_total;
      const double inv = 1.0 / r.S[h];
      const double u = live ? bnn_samp_uniform(p.samp_seed, row, c) : 2.0;
      double base = 0.0;
      int drawn = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < N3 / 8; ++j) {
        const int c0 = 8 * j + 2 * t;
        const double p0 = (c0 < K) ? r.ev[h][2 * j] * inv : 0.0, p1 = (c0 + 1 < K) ? r.ev[h][2 * j + 1] * inv : 0.0;
        double incl = p0 + p1;                                    // inclusive scan over t (lanes of the quad)
        const double up1 = __shfl_up_sync(FULL_MASK, incl, 1, 4);
        if (t >= 1) incl += up1;
        const double up2 = __shfl_up_sync(FULL_MASK, incl, 2, 4);
        if (t >= 2) incl += up2;
        const double excl = base + incl - (p0 + p1);
        const double cum0 = excl + p0, cum1 = cum0 + p1;
        if (c0 + 1 < K && cum1 - u >= 0.0 && c0 + 1 < drawn) drawn = c0 + 1;
        if (c0 < K && cum0 - u >= 0.0 && c0 < drawn) drawn = c0;
        base += __shfl_sync(FULL_MASK, incl, 3, 4);               // total of this 8-column block
      }
      drawn = min(drawn, __shfl_xor_sync(FULL_MASK, drawn, 1));
      drawn = min(drawn, __shfl_xor_sync(FULL_MASK, drawn, 2));
      if (drawn == 0x7fffffff) drawn = 0;
      if (live) {
#pragma unroll
        for (int j = 0; j < N3 / 8; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (8 * j + 2 * t + e == drawn) pvote[h][2 * j + e] += 1;
        if (t == 0 && p.samp_dense) p.samp_dense[row * p.C + c] = (double)drawn;
      }
      if (p.samp_counts) {
        // one atomic per distinct class among the rows of this half tile (lanes t == 0 speak for their row)
        const unsigned grp = __match_any_sync(FULL_MASK, (live && t == 0) ? drawn : 64 + lane);
        if (live && t == 0 && lane == __ffs(grp) - 1) atomicAdd(&p.samp_counts[c * K + drawn], __popc(grp));
      }
    }
    return;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long row = wt * 16 + gq + 8 * h;
    if (row < p.n_total) {
      const double inv = 1.0 / r.S[h];
#pragma unroll
      for (int j = 0; j < N3 / 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * j + 2 * t + e;
          if (col < K) {
            const double pk = r.ev[h][2 * j + e] * inv;
            pacc[h][2 * j + e] += pk;
            if (p.dense_out) p.dense_out[((long long)c * p.n_total + row) * K + col] = pk;
            if (col == r.arg[h]) pvote[h][2 * j + e] += 1;
          }
        }
    }
  }
}

// Gaussian likelihoods (calc_likelihood_regression / _error, BNN_lib.py:123-143) on the last layer's accumulator
// fragment (N3 = 8: thread (g, t) holds rows g and g + 8, columns 2t and 2t + 1).  Per (warp tile, weight set) and
// modelled output j the kernel leaves sum r, sum r^2 over the training rows and sum r^2 over the test rows
// (r = prediction - target), from which finalize_loglik forms the log-likelihood (fixed or empirical sigma) and the
// MSE statistics; with the sigma head (columns K..2K-1 through softplus) the per-row log-density is summed here.
template <int N3>
__device__ __forceinline__ void quad_gauss_lik(const FwdParams& p, int c, long long wt, int lane,
                                               const double (&acc)[N3 / 8][4], const double* tab) {
  static_assert(N3 == 8, "Gaussian outputs fit one 8-column tile");
  const int K = p.g.K;
  const bool head = (p.g.lik == BNN_LIK_GAUSSIAN_HEAD);
  const int gq = lane >> 2, t = lane & 3;
  const long long nt = p.n_tiles16;
  double ll = 0.0, sr[2] = {0.0, 0.0}, sr2[2] = {0.0, 0.0}, st2[2] = {0.0, 0.0};
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int col = 2 * t + e;
    const int scol = (col + K) & 7;                                // column of this output's sigma parameter (head)
    const int src = (gq << 2) | (scol >> 1);
    const double s00 = __shfl_sync(FULL_MASK, acc[0][0], src), s01 = __shfl_sync(FULL_MASK, acc[0][1], src);
    const double s10 = __shfl_sync(FULL_MASK, acc[0][2], src), s11 = __shfl_sync(FULL_MASK, acc[0][3], src);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long row = wt * 16 + gq + 8 * h;
      const bool active = row < p.n_total && col < K;
      const bool train = active && row < p.n_train;
      double r = 0.0;
      if (active) {
        const double tv = p.targets[row * K + col];
        const double mu = acc[0][2 * h + e];
        r = mu - tv;
        if (head && train) {
          const double zs = (scol & 1) ? (h ? s11 : s01) : (h ? s10 : s00);
          if (fabs(zs) < 700.0) {
            const double sd = bnn_softplus_fast(zs, tab);
            const double u = (tv - mu) * bnn_rcp(sd);
            ll += -0.5 * u * u - kLogSqrt2Pi - bnn_log_ge1(sd);
          } else {                                                   // out of the table routine's range, inf, NaN: libm
            const double sd = softplus_ref(zs);
            const double u = (tv - mu) / sd;
            ll += -0.5 * u * u - kLogSqrt2Pi - log(sd);
          }
        }
      }
      sr[e] += train ? r : 0.0;
      sr2[e] += train ? r * r : 0.0;
      st2[e] += (active && !train) ? r * r : 0.0;
    }
  }
#pragma unroll
  for (int o = 4; o <= 16; o <<= 1)                                 // rows: fixed xor tree over g => deterministic
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      sr[e] += __shfl_xor_sync(FULL_MASK, sr[e], o);
      sr2[e] += __shfl_xor_sync(FULL_MASK, sr2[e], o);
      st2[e] += __shfl_xor_sync(FULL_MASK, st2[e], o);
    }
#pragma unroll
  for (int o = 1; o <= 16; o <<= 1) ll += __shfl_xor_sync(FULL_MASK, ll, o);
  if (gq == 0) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 2 * t + e;
      if (col < K) {
        p.part[((long long)c * p.NF + 1 + col) * nt + wt] = sr[e];
        p.part[((long long)c * p.NF + 1 + K + col) * nt + wt] = sr2[e];
        p.part[((long long)c * p.NF + 1 + 2 * K + col) * nt + wt] = st2[e];
      }
    }
  }
  if (lane == 0) p.part[((long long)c * p.NF) * nt + wt] = ll;
}

// prediction mode: transformed outputs (identity, softplus on the sigma head) accumulated over the weight sets
template <int N3>
__device__ __forceinline__ void quad_gauss_pred(const FwdParams& p, int c, long long wt, int lane,
                                                const double (&acc)[N3 / 8][4], double (&pacc)[2][N3 / 4], const double* tab) {
  static_assert(N3 == 8, "Gaussian outputs fit one 8-column tile");
  const int K = p.g.K, O = p.g.O;
  const bool head = (p.g.lik == BNN_LIK_GAUSSIAN_HEAD);
  const int gq = lane >> 2, t = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long row = wt * 16 + gq + 8 * h;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 2 * t + e;
      if (row < p.n_total && col < O) {
        double v = acc[0][2 * h + e];
        if (head && col >= K) v = (fabs(v) < 700.0) ? bnn_softplus_fast(v, tab) : softplus_ref(v);
        pacc[h][e] += v;
        if (p.dense_out) p.dense_out[((long long)c * p.n_total + row) * O + col] = v;
      }
    }
  }
}

// MODE: 0 likelihood, 1 likelihood with class / instance weights, 2 posterior prediction
enum { FWD3_LIK = 0, FWD3_LIK_W = 1, FWD3_PRED = 2 };
// LIKK: 0 categorical (softmax epilogue, software-pipelined into the next weight set), 1 Gaussian (plain or sigma head)
enum { FWD3_CAT = 0, FWD3_GAUSS = 1 };

template <int ACT, int KP0, int N1, int N2, int N3, int NWARPS, int MODE, int LIKK = FWD3_CAT>
__global__ void __launch_bounds__(NWARPS * 32, 1) k_fwd3(const __grid_constant__ FwdParams p) {
  using G3 = Fwd3Geom<KP0, N1, N2, N3>;
  constexpr bool PREDICT = (MODE == FWD3_PRED);
  constexpr bool CAT = (LIKK == FWD3_CAT);
  constexpr bool DEFER = !PREDICT && CAT;        // categorical likelihood: epilogue pipelined into the next set's layer 1
#ifdef BNN_DBG_NOEPI      // tuning experiment only: no likelihood epilogue at all (results are garbage)
  constexpr bool EPI = false;
#else
  constexpr bool EPI = DEFER;
#endif
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const NetGeom& g = p.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;

  // ---- shared memory carve-up
  // Weight ring, two slots per part: part 0 = first-layer matrix + bias (W1_D doubles), part 1 = the rest of the
  // packed weight set (W23_D doubles).  The parts are released separately: the first layer is 61 % of a use, so its
  // slot is free (and its refill under way) long before the warp is done with the use.
  constexpr int W1_D = G3::W2_OFF, W23_D = G3::PB - G3::W2_OFF;
  double* w1buf = reinterpret_cast<double*>(smem_raw);                  // [2][W1_D]
  double* w23buf = w1buf + 2 * W1_D;                                    // [2][W23_D]
  double* xs = w23buf + 2 * W23_D + warp * 16 * KP0;                    // [NWARPS][16*KP0] X warp tiles
  double* tab = w23buf + 2 * W23_D + NWARPS * 16 * KP0;                 // exp table
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab + BNN_EXP_TAB_SIZE);
  uint64_t* full1 = bars;           // [2]
  uint64_t* full23 = bars + 2;      // [2]
  uint64_t* xbar = bars + 4 + warp; // [NWARPS]
  int* rel = reinterpret_cast<int*>(bars + 4 + NWARPS);   // [4] warps that released slot b of part 0 / part 1
  int* cnt = reinterpret_cast<int*>(bars + 6 + NWARPS);
  const int n_cnt = DEFER ? p.C * (2 + 2 * g.K) : 0;
  const int PW = CAT ? g.K : g.O;                 // columns of a prediction row

  for (int i = threadIdx.x; i < BNN_EXP_TAB_SIZE; i += blockDim.x) tab[i] = p.exp_tab[i];
  for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) cnt[i] = 0;
  if (threadIdx.x == 0) {
    rel[0] = rel[1] = rel[2] = rel[3] = 0;
    mbar_init(&full1[0], 1); mbar_init(&full1[1], 1);
    mbar_init(&full23[0], 1); mbar_init(&full23[1], 1);
    for (int w = 0; w < NWARPS; ++w) mbar_init(&bars[4 + w], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // Work decomposition.  n_full rounds in which every warp of the grid owns one 16-row tile and runs all C
  // weight sets over it; the remaining tiles (fewer than one per warp) form `groups` CTA-sized tile groups
  // whose (group, weight set) pairs are dealt out evenly over the CTAs, so that the last round costs
  // ceil(groups*C/grid) weight sets per CTA instead of C (c4: 7 instead of 32, 36 rounds -> 35.2).  A CTA's
  // share is a contiguous range of pairs, i.e. at most two segments (tile group, [lo, hi)).  Every (tile,
  // weight set) is still computed by exactly one warp with the same instruction sequence, so results do not
  // depend on the decomposition.  Prediction accumulates over weight sets in registers: no split there.
  const long long total_warps = (long long)gridDim.x * NWARPS;
  const long long n_full = p.n_tiles16 / total_warps;
  const long long tail_base = n_full * total_warps;
  const long long rem = p.n_tiles16 - tail_base;
  int seg_tg[2] = {0, 0}, seg_lo[2] = {0, 0}, seg_hi[2] = {0, 0};
  if (rem > 0) {
    const long long groups = (rem + NWARPS - 1) / NWARPS;
    if (PREDICT) {
      if ((long long)blockIdx.x < groups) { seg_tg[0] = blockIdx.x; seg_hi[0] = p.C; }
    } else {
      const long long pairs = groups * p.C;
      const long long per_cta = (pairs + gridDim.x - 1) / gridDim.x;
      const long long s0 = (long long)blockIdx.x * per_cta;
      const long long s1 = (s0 + per_cta < pairs) ? s0 + per_cta : pairs;
      if (s0 < s1) {
        seg_tg[0] = (int)(s0 / p.C);
        seg_lo[0] = (int)(s0 % p.C);
        const long long n0 = ((s1 - s0) < (long long)(p.C - seg_lo[0])) ? (s1 - s0) : (long long)(p.C - seg_lo[0]);
        seg_hi[0] = seg_lo[0] + (int)n0;
        if (s1 - s0 > n0) { seg_tg[1] = seg_tg[0] + 1; seg_hi[1] = (int)(s1 - s0 - n0); }
      }
    }
  }
  const long long q_seg0 = n_full * p.C;                         // first weight-set use of segment 0
  const long long q_seg1 = q_seg0 + (seg_hi[0] - seg_lo[0]);     // ... of segment 1
  const long long total_q = q_seg1 + (seg_hi[1] - seg_lo[1]);    // weight-set uses, identical for every warp of the CTA
  const long long n_iter = n_full + (seg_hi[0] > seg_lo[0] ? 1 : 0) + (seg_hi[1] > seg_lo[1] ? 1 : 0);
  auto set_of = [&](long long qq) -> long long {                 // weight set consumed by use qq of this CTA
    return qq < q_seg0 ? qq % p.C : (qq < q_seg1 ? seg_lo[0] + (qq - q_seg0) : seg_lo[1] + (qq - q_seg1));
  };
  constexpr uint32_t W1_BYTES = W1_D * sizeof(double), W23_BYTES = W23_D * sizeof(double);
  constexpr uint32_t X_BYTES = 16 * KP0 * sizeof(double);
  static_assert(W1_BYTES % 16 == 0 && W23_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");

  auto load_part = [&](int part, long long qq) {
    const int b = (int)(qq & 1);
    const double* src = p.wp + set_of(qq) * G3::PB;
    if (part == 0) {
      mbar_arrive_expect_tx(&full1[b], W1_BYTES);
      bulk_g2s(w1buf + b * W1_D, src, W1_BYTES, &full1[b]);
    } else {
      mbar_arrive_expect_tx(&full23[b], W23_BYTES);
      bulk_g2s(w23buf + b * W23_D, src + W1_D, W23_BYTES, &full23[b]);
    }
  };
  if (threadIdx.x == 0) {
    for (int q = 0; q < 2 && q < total_q; ++q) { load_part(0, q); load_part(1, q); }
  }

  // A slot is released by every warp when it is done with use qq of that part; the warp whose release is the last
  // one issues the bulk copy of use qq+2 into the slot itself, so a refill starts the moment the slot is free (a fixed
  // producer thread would start it only when its own warp next reaches the top of the weight-set loop, and would make
  // that warp wait for the slowest one before every use: 17.01 -> 16.68 ms at c4).  The warp schedulers prefer the
  // younger warps (measured ring waits per warp: 0.3 / 2 / 4.5 % of the run for warps 0-3 / 4-7 / 8-11), so the
  // slowest warp is rarely warp 0.
  auto release_part = [&](int part, long long qq) {
    if (lane == 0) {
      int* r = &rel[2 * part + (int)(qq & 1)];
      __threadfence_block();                       // this warp's reads of the slot happen before the release
      const int old = atomicAdd(r, 1);
      if (old == NWARPS - 1) {                     // last warp out refills the slot
        *r = 0;
        if (qq + 2 < total_q) {
          __threadfence_block();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          load_part(part, qq + 2);
        }
      }
    }
  };
#ifdef BNN_DBG_RINGCLK
  long long dbg_ring = 0;
  const long long dbg_start = clock64();
#endif
  long long q = 0;
  uint32_t x_uses = 0;               // X tiles this warp has loaded (parity of its mbarrier)
  bool x_prefetched = false;         // the X tile of this iteration was requested during the previous one
  auto tile_of = [&](long long i) -> long long {             // 16-row tile of this warp in iteration i
    const int s2 = (int)(i - n_full);
    return s2 < 0 ? i * total_warps + (long long)blockIdx.x * NWARPS + warp
                  : tail_base + (long long)(s2 == 0 ? seg_tg[0] : seg_tg[1]) * NWARPS + warp;
  };
  // Likelihood modes: the epilogue of a (tile, weight set) is software-pipelined into layer 1 of the NEXT weight set
  // this warp runs -- also across a tile boundary, so that a tile's last set is not finished in a phase of its own
  // (that phase is 1/32 of a tile at 32 sets per GPU, but a quarter at 4 and everything for a single chain).
  // acc3 / ep_* describe the pending epilogue: last-layer accumulators, tile, weight set, labels (and weights).
  double acc3[N3 / 8][4];
  bool prev_valid = false;
  long long ep_wt = 0;
  int ep_c = 0;
  int ep_y[2] = {0, 0};
  double ep_wgt[2] = {1.0, 1.0};
#pragma unroll
  for (int j = 0; j < N3 / 8; ++j) acc3[j][0] = acc3[j][1] = acc3[j][2] = acc3[j][3] = 0.0;
  for (long long it = 0; it < n_iter; ++it) {
    const int sg = (int)(it - n_full);                       // tail segment (>= 0) or full round (< 0)
    const long long wt = tile_of(it);
    const int c_lo = sg < 0 ? 0 : (sg == 0 ? seg_lo[0] : seg_lo[1]);
    const int c_hi = sg < 0 ? p.C : (sg == 0 ? seg_hi[0] : seg_hi[1]);
    const bool have_tile = wt < p.n_tiles16;
    int y[2] = {0, 0};
    double wgt[2] = {1.0, 1.0};
    double pacc[2][N3 / 4];
    int pvote[2][N3 / 4];
    if (have_tile) {
      if (lane == 0 && !x_prefetched) {
        // order this warp's earlier generic-proxy reads of xs before the async-proxy overwrite
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(xbar, X_BYTES);
        bulk_g2s(xs, p.x + wt * 16 * KP0, X_BYTES, xbar);
      }
      x_prefetched = false;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long row = wt * 16 + gq + 8 * h;
        if (DEFER && row < p.n_total) {
          y[h] = p.labels[row];
          if (MODE == FWD3_LIK_W) {
            if (p.class_w) wgt[h] *= p.class_w[y[h]];
            if (p.inst_w && row < p.n_train) wgt[h] *= p.inst_w[row];
          }
        }
#pragma unroll
        for (int i = 0; i < N3 / 4; ++i) { pacc[h][i] = 0.0; pvote[h][i] = 0; }
      }
      mbar_wait(xbar, x_uses & 1);
      ++x_uses;
    }
    for (int c = c_lo; c < c_hi; ++c, ++q) {
      const int b = (int)(q & 1);
      __syncwarp();
#ifdef BNN_DBG_RINGCLK      // tuning experiment only: clocks this warp spends waiting for the weight ring
      const long long dbg_t0 = clock64();
#endif
      mbar_wait(&full1[b], (uint32_t)((q >> 1) & 1));
#ifdef BNN_DBG_RINGCLK
      dbg_ring += clock64() - dbg_t0;
#endif
      if (have_tile) {
        const double* W = w1buf + b * W1_D;                          // part 0; layers 2 / 3 index W23 with the same offsets
        const double* W23 = w23buf + b * W23_D - G3::W2_OFF;
        const double a1 = (ACT == BNN_ACT_LEAKY && p.alpha) ? p.alpha[c * 3 + 0] : 0.0;
        const double a2 = (ACT == BNN_ACT_LEAKY && p.alpha) ? p.alpha[c * 3 + 1] : 0.0;
        // ---------------- layer 1: [16 x KP0] x [KP0 x N1]
        double acc1[N1 / 8][4];
#pragma unroll
        for (int j = 0; j < N1 / 8; ++j) {
          const double2 bb = *reinterpret_cast<const double2*>(W + G3::B1_OFF + 8 * j + 2 * t);
          acc1[j][0] = bb.x; acc1[j][1] = bb.y; acc1[j][2] = bb.x; acc1[j][3] = bb.y;
        }
        {
          const double* xr0 = xs + gq * KP0;
          const double* xr1 = xs + (gq + 8) * KP0;
          const double* wr = W + G3::W1_OFF + gq * KP0;
          const int sw = (gq & 1) * G3::SW0;
          // Software pipeline over weight sets: the epilogue of the PREVIOUS set (softmax, log-likelihood,
          // counters -- latency-bound FP64 chains and shuffles) sits in the same basic block as the first two
          // k-groups of this set's layer 1, whose 64 DMMAs need no FP64 issue slots of their own.
          // Software pipeline over weight sets: the epilogue of the PREVIOUS set (softmax statistics, log-
          // likelihood: latency-bound FP64 chains and shuffles) is cut into four stages that are placed
          // between the MMA groups of the first two k-groups of this set's layer 1, whose DMMAs leave the
          // FP64 issue slots free.
          RowStats<N3> rs;
          LikRow lr;
          const int K = g.K;
#pragma unroll
          for (int kg = 0; kg < 2; ++kg) {
            const int col = (8 * kg + 2 * t) ^ sw;
            const double2 alo = *reinterpret_cast<const double2*>(xr0 + col);
            const double2 ahi = *reinterpret_cast<const double2*>(xr1 + col);
            if (EPI) {
              if (kg == 0) qs_max<N3>(acc3, K, t, rs);
              else qs_exp<N3, 1, true>(acc3, K, t, ep_y, tab, rs);
            }
#pragma unroll
            for (int j = 0; j < N1 / 16; ++j) {
              const double2 bb = *reinterpret_cast<const double2*>(wr + j * 8 * KP0 + col);
              dmma16x8x8(acc1[j], alo.x, ahi.x, alo.y, ahi.y, bb.x, bb.y);
            }
            if (EPI) {
              if (kg == 0) qs_exp<N3, 0, true>(acc3, K, t, ep_y, tab, rs);
              else {
                qs_reduce<N3, true>(rs);
                lr = quad_lik_finish<N3, MODE == FWD3_LIK_W>(p, ep_wt, lane, rs, ep_y, ep_wgt, prev_valid);
              }
            }
#pragma unroll
            for (int j = N1 / 16; j < N1 / 8; ++j) {
              const double2 bb = *reinterpret_cast<const double2*>(wr + j * 8 * KP0 + col);
              dmma16x8x8(acc1[j], alo.x, ahi.x, alo.y, ahi.y, bb.x, bb.y);
            }
          }
#ifndef BNN_DBG_NOCOMMIT      // tuning experiment only: counters / partial store of the previous set
          if (EPI) quad_lik_commit<N3>(p, ep_c, ep_wt, lane, cnt, lr, prev_valid);
#endif
#pragma unroll 2
          for (int kg = 2; kg < KP0 / 8 - 2; ++kg) {
            const int col = (8 * kg + 2 * t) ^ sw;
            const double2 alo = *reinterpret_cast<const double2*>(xr0 + col);
            const double2 ahi = *reinterpret_cast<const double2*>(xr1 + col);
#pragma unroll
            for (int j = 0; j < N1 / 8; ++j) {
              const double2 bb = *reinterpret_cast<const double2*>(wr + j * 8 * KP0 + col);
              dmma16x8x8(acc1[j], alo.x, ahi.x, alo.y, ahi.y, bb.x, bb.y);
            }
          }
          // last two k-groups tile by tile: acc1[0] is complete first, so its activation (below, same basic
          // block) overlaps with the MMAs that finish the other tiles
          {
            const int c0 = (8 * (KP0 / 8 - 2) + 2 * t) ^ sw, c1 = (8 * (KP0 / 8 - 1) + 2 * t) ^ sw;
            const double2 alo0 = *reinterpret_cast<const double2*>(xr0 + c0), ahi0 = *reinterpret_cast<const double2*>(xr1 + c0);
            const double2 alo1 = *reinterpret_cast<const double2*>(xr0 + c1), ahi1 = *reinterpret_cast<const double2*>(xr1 + c1);
#pragma unroll
            for (int j = 0; j < N1 / 8; ++j) {
              const double2 b0 = *reinterpret_cast<const double2*>(wr + j * 8 * KP0 + c0);
              const double2 b1 = *reinterpret_cast<const double2*>(wr + j * 8 * KP0 + c1);
              dmma16x8x8(acc1[j], alo0.x, ahi0.x, alo0.y, ahi0.y, b0.x, b0.y);
              dmma16x8x8(acc1[j], alo1.x, ahi1.x, alo1.y, ahi1.y, b1.x, b1.y);
            }
          }
        }
        // the first-layer weights of this use are no longer needed by this warp; layers 2 / 3 need part 1
        __syncwarp();
        release_part(0, q);
        mbar_wait(&full23[b], (uint32_t)((q >> 1) & 1));
        // X is read by layer 1 only: after layer 1 of the tile's LAST weight set the buffer is free, and the next
        // tile's X (one bulk copy, ~2 us from HBM) arrives under layers 2 / 3 and the epilogue instead of stalling
        // the start of the next tile -- 0.4 % of a 32-set tile, but 3 % at 4 sets per GPU and 12 % for a single chain.
        if (c == c_hi - 1 && it + 1 < n_iter) {
          const long long wt_next = tile_of(it + 1);
          if (wt_next < p.n_tiles16) {
            __syncwarp();                                      // every lane's reads of xs are done
            if (lane == 0) {
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              mbar_arrive_expect_tx(xbar, X_BYTES);
              bulk_g2s(xs, p.x + wt_next * 16 * KP0, X_BYTES, xbar);
            }
            x_prefetched = true;
          }
        }
        // ---------------- layer 2: [16 x N1] x [N1 x N2]   (A operand = activated acc1, no data movement).
        // The activation of k-group kg+1 is independent of the MMAs of k-group kg, so its FP64 chains fill
        // the issue slots between the DMMAs instead of running as a separate latency-bound phase.
        // Two straight-line copies per layer: one warp vote decides whether any pre-activation needs the guarded
        // evaluation (frag_needs_care); the common copy carries no clamp / NaN fix-up instructions.
        double acc2[N2 / 8][4];
#pragma unroll
        for (int j = 0; j < N2 / 8; ++j) {
          const double2 bb = *reinterpret_cast<const double2*>(W23 + G3::B2_OFF + 8 * j + 2 * t);
          acc2[j][0] = bb.x; acc2[j][1] = bb.y; acc2[j][2] = bb.x; acc2[j][3] = bb.y;
        }
        auto layer2 = [&](auto fast_tag) {
          constexpr bool FAST = decltype(fast_tag)::value;
          const double* wr = W23 + G3::W2_OFF + gq * N1;
          const int sw = (gq & 1) * G3::SW1;
          act_tile<ACT, FAST>(acc1[0], a1, tab);
#pragma unroll
          for (int kg = 0; kg < N1 / 8 - 2; ++kg) {
            act_tile<ACT, FAST>(acc1[kg + 1], a1, tab);
            const int col = (8 * kg + 2 * t) ^ sw;
#pragma unroll
            for (int j = 0; j < N2 / 8; ++j) {
              const double2 bb = *reinterpret_cast<const double2*>(wr + j * 8 * N1 + col);
              dmma16x8x8(acc2[j], acc1[kg][0], acc1[kg][2], acc1[kg][1], acc1[kg][3], bb.x, bb.y);
            }
          }
          {
            constexpr int k0 = N1 / 8 - 2, k1 = N1 / 8 - 1;
            act_tile<ACT, FAST>(acc1[k1], a1, tab);
            const int c0 = (8 * k0 + 2 * t) ^ sw, c1 = (8 * k1 + 2 * t) ^ sw;
#pragma unroll
            for (int j = 0; j < N2 / 8; ++j) {
              const double2 b0 = *reinterpret_cast<const double2*>(wr + j * 8 * N1 + c0);
              const double2 b1 = *reinterpret_cast<const double2*>(wr + j * 8 * N1 + c1);
              dmma16x8x8(acc2[j], acc1[k0][0], acc1[k0][2], acc1[k0][1], acc1[k0][3], b0.x, b0.y);
              dmma16x8x8(acc2[j], acc1[k1][0], acc1[k1][2], acc1[k1][1], acc1[k1][3], b1.x, b1.y);
            }
          }
        };
#ifdef FWD3_NO_FAST_ACT
        layer2(std::false_type{});
#else
        if (frag_needs_care<ACT, N1 / 8>(acc1)) layer2(std::false_type{});
        else layer2(std::true_type{});
#endif
        // ---------------- layer 3: [16 x N2] x [N2 x N3]
#pragma unroll
        for (int j = 0; j < N3 / 8; ++j) {
          const double2 bb = *reinterpret_cast<const double2*>(W23 + G3::B3_OFF + 8 * j + 2 * t);
          acc3[j][0] = bb.x; acc3[j][1] = bb.y; acc3[j][2] = bb.x; acc3[j][3] = bb.y;
        }
        auto layer3 = [&](auto fast_tag) {
          constexpr bool FAST = decltype(fast_tag)::value;
          const double* wr = W23 + G3::W3_OFF + gq * N2;
          const int sw = (gq & 1) * G3::SW2;
          act_tile<ACT, FAST>(acc2[0], a2, tab);
#pragma unroll
          for (int kg = 0; kg < N2 / 8; ++kg) {
            if (kg + 1 < N2 / 8) act_tile<ACT, FAST>(acc2[kg + 1], a2, tab);
            const int col = (8 * kg + 2 * t) ^ sw;
#pragma unroll
            for (int j = 0; j < N3 / 8; ++j) {
              const double2 bb = *reinterpret_cast<const double2*>(wr + j * 8 * N2 + col);
              dmma16x8x8(acc3[j], acc2[kg][0], acc2[kg][2], acc2[kg][1], acc2[kg][3], bb.x, bb.y);
            }
          }
        };
#ifdef FWD3_NO_FAST_ACT
        layer3(std::false_type{});
#else
        if (frag_needs_care<ACT, N2 / 8>(acc2)) layer3(std::false_type{});
        else layer3(std::true_type{});
#endif
        // weights of this use are no longer needed by this warp
        __syncwarp();
        release_part(1, q);
        if (DEFER) {
          prev_valid = true;
          ep_wt = wt; ep_c = c;
          ep_y[0] = y[0]; ep_y[1] = y[1];
          if (MODE == FWD3_LIK_W) { ep_wgt[0] = wgt[0]; ep_wgt[1] = wgt[1]; }
        }
        if constexpr (CAT) {
          if (PREDICT) quad_epilogue_pred<N3>(p, c, wt, lane, acc3, tab, pacc, pvote);
        } else {
          if (PREDICT) quad_gauss_pred<N3>(p, c, wt, lane, acc3, pacc, tab);
          else quad_gauss_lik<N3>(p, c, wt, lane, acc3, tab);
        }
      } else {
        // (a warp without a tile waits for both parts before releasing them, so that its releases cannot run ahead)
        mbar_wait(&full23[b], (uint32_t)((q >> 1) & 1));
        release_part(0, q);
        release_part(1, q);
      }
    }
    if (have_tile) {
      if (PREDICT) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const long long row = wt * 16 + gq + 8 * h;
#pragma unroll
          for (int j = 0; j < N3 / 8; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = 8 * j + 2 * t + e;
              if (row < p.n_total && col < PW) {
                if (p.mean_out) p.mean_out[row * PW + col] = pacc[h][2 * j + e] / p.inv_sets;
                if (CAT && p.votes_out) p.votes_out[row * PW + col] = (double)pvote[h][2 * j + e] / p.inv_sets;
              }
            }
        }
      }
    }
  }
#ifdef BNN_DBG_RINGCLK
  if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77) && p.C >= 32)
    printf("cta %d warp %2d (smsp %d): ring wait %lld of %lld clk (%.1f%%)\n", blockIdx.x, warp, warp & 3, dbg_ring,
           clock64() - dbg_start, 100.0 * dbg_ring / (double)(clock64() - dbg_start));
#endif
  // drain the software pipeline: the epilogue of the last (tile, weight set) this warp ran
  if (DEFER && prev_valid) {
    RowStats<N3> rs;
    quad_softmax_stats<N3, true>(acc3, g.K, t, ep_y, tab, rs);
    const LikRow lr = quad_lik_finish<N3, MODE == FWD3_LIK_W>(p, ep_wt, lane, rs, ep_y, ep_wgt, true);
    quad_lik_commit<N3>(p, ep_c, ep_wt, lane, cnt, lr, true);
  }
  if (n_cnt) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_cnt; i += blockDim.x)
      if (cnt[i]) atomicAdd(&p.counts[i], cnt[i]);
  }
}

// The int8 tensor-core first layer (k_fwd3t, below) is EXPERIMENTAL and compiled only with -DBNN_EXPERIMENTAL_TENSOR_L1:
// with its helper warps allowed to run two weight sets ahead an intermittent corruption was observed whose cause is not
// understood, so the shipped library does not contain the kernel and "tensor_l1" cannot be enabled.
#ifndef BNN_EXPERIMENTAL_TENSOR_L1
cudaError_t bnn_debug_set_trace_ptr(unsigned long long*) { return cudaErrorNotSupported; }
cudaError_t bnn_debug_counters_read(unsigned long long* out48) {
  for (int i = 0; i < 48; ++i) out48[i] = 0;
  return cudaSuccess;
}
cudaError_t bnn_launch_slice_x(const double*, long long, uint8_t*, double*, long long, int*, cudaStream_t) {
  return cudaErrorNotSupported;
}
cudaError_t bnn_launch_slice_w1(const double*, int, uint8_t*, int, cudaStream_t) { return cudaErrorNotSupported; }
size_t bnn_slice_x_tile_bytes() { return 0; }
size_t bnn_slice_w1_bytes() { return 0; }
#else
// =============================================================================================
// k_fwd3t: the 3-layer kernel with layer 1 on the 5th-generation tensor cores (tcgen05 + TMEM)
// =============================================================================================
// tcgen05 has no FP64 kind, but the first-layer contraction X W1^T can be evaluated EXACTLY in integer
// arithmetic (Ozaki splitting): every row of X and every row of W1 is scaled by a power of two and rounded to a
// 47-bit signed integer, which is cut into 6 balanced base-256 digits (int8 "slices"):
//     x_ik = 2^(e_i - 46) * sum_s a_s[i,k] 256^s          w_nk = 2^(f_n - 46) * sum_t b_t[n,k] 256^t
//     sum_k x_ik w_nk = 2^(e_i + f_n - 92) * sum_T 256^T * ( sum_{s+t=T} sum_k a_s[i,k] b_t[n,k] )
// The inner sums are int8 x int8 -> int32 tensor-core products (tcgen05.mma kind::i8, exact, no overflow:
// 6 pairs * 64 * 128^2 < 2^31); the 6 most significant diagonals T = 10..5 (21 slice pairs) are kept, the
// dropped ones are zero-mean and below 2^-41 of max|x_i| * max|w_n| (measured: log-likelihood within 1e-12
// relative of the FP64 kernels).  X is sliced once at bnn_set_data (k_slice_x), W1 once per proposal
// (k_slice_w1); both are stored in HBM in the tensor core's K-major core-matrix layout, so a plain bulk copy
// brings them to shared memory.  Layers 2 and 3 stay on the FP64 DMMA path of k_fwd3 -- their A operand is a
// per-chain activation, so the chains cannot share an MMA and a 32- or 16-column tcgen05.mma costs as much as
// a 128-column one (70 clk floor, tools/umma_probe.cu).
//
// CTA = one 128-row tile at a time: 8 compute warps (16 rows each, the TMEM lane quarter of warp w is w & 3)
// + 1 control warp whose elected lane streams the operands (bulk copies) and issues the MMAs.  Per chain:
//   control : wait weights(q) and "TMEM drained(q-1)" -> 42 MMAs into 6 x 64 TMEM columns -> commit
//   compute : wait commit(q) -> tcgen05.ld.16x256b (= the m16n8 accumulator fragment layout) -> recombine the
//             6 diagonals to FP64 in registers (two int64 groups, 5 FP64 instructions per element) ->
//             signal "drained" -> activation, layers 2/3, likelihood epilogue exactly as k_fwd3
// so the integer MMAs of chain q+1 run under the FP64 work of chain q.
constexpr int OZ_S = 6;                               // slices per operand
constexpr int OZ_P = 46;                              // operands are rounded to |v| <= 2^46
constexpr int OZ_XROW = OZ_S * 64;                    // bytes of one row of X slices
constexpr int OZ_XTILE = 128 * OZ_XROW;
constexpr int OZ_WPLANE = 64 * 64;
constexpr int OZ_W1_BYTES = OZ_S * OZ_WPLANE + 64 * 8;   // 6 planes + colscale[64]
constexpr int OZ_LBO = 128, OZ_SBO = 512;             // K-major, no swizzle: 8-row x 16-byte core matrices

__host__ __device__ inline int oz_kmajor_offset(int r, int k) {
  return (r >> 3) * OZ_SBO + (k >> 4) * OZ_LBO + (r & 7) * 16 + (k & 15);
}

// 47-bit integer image of v * 2^(46 - e) as 6 balanced base-256 digits
__device__ __forceinline__ void oz_digits(double v, int e, int8_t (&d)[OZ_S]) {
  long long q = __double2ll_rn(scalbn(v, OZ_P - e));
#pragma unroll
  for (int s = 0; s < OZ_S - 1; ++s) {
    const int b = (int)((q + 128) & 255) - 128;
    d[s] = (int8_t)b;
    q = (q - b) >> 8;
  }
  d[OZ_S - 1] = (int8_t)q;
}
// exponent e with max < 2^e (0 for an all-zero row)
__device__ __forceinline__ int oz_exponent(double mx) { return mx > 0.0 ? ilogb(mx) + 1 : 0; }

// X (swizzled rows, F_pad = 64) -> slices [n_rows128][6][64 B] (a row's 6 x 64 bytes are contiguous: the helper
// thread that owns the row copies them to TMEM, where they are the A operand) + rowscale[n_rows128] = 2^e_i.
// One thread per row.  flag is set when a row holds a non-finite value (the tensor path is then not used).
__global__ void k_slice_x(const double* __restrict__ x, long long n_pad16, uint8_t* __restrict__ xsl,
                          double* __restrict__ rowscale, long long n_rows128, int* __restrict__ flag) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows128) return;
  const int sw = (int)(r & 1) * 8;
  double mx = 0.0;
  bool bad = false;
  if (r < n_pad16)
    for (int k = 0; k < 64; ++k) {
      const double v = x[r * 64 + (k ^ sw)];
      bad |= !isfinite(v);
      mx = fmax(mx, fabs(v));
    }
  if (bad) { atomicExch(flag, 1); mx = 0.0; }
  const int e = oz_exponent(mx);
  rowscale[r] = scalbn(1.0, e);
  uint32_t* out = reinterpret_cast<uint32_t*>(xsl + r * (long long)(OZ_S * 64));
  for (int k4 = 0; k4 < 16; ++k4) {
    uint32_t w[OZ_S] = {0, 0, 0, 0, 0, 0};
    for (int kk = 0; kk < 4; ++kk) {
      const int k = 4 * k4 + kk;
      int8_t d[OZ_S];
      const double v = (r < n_pad16 && !bad) ? x[r * 64 + (k ^ sw)] : 0.0;
      oz_digits(v, e, d);
#pragma unroll
      for (int sl = 0; sl < OZ_S; ++sl) w[sl] |= (uint32_t)(uint8_t)d[sl] << (8 * kk);
    }
#pragma unroll
    for (int sl = 0; sl < OZ_S; ++sl) out[sl * 16 + k4] = w[sl];
  }
}

// packed weight sets (W1 = [64][64] swizzled rows at offset 0) -> per set: 6 slice planes + colscale[n] = 2^(f_n - 48)
__global__ void k_slice_w1(const double* __restrict__ wp, int PB, uint8_t* __restrict__ wt) {
  const int c = blockIdx.x, n = threadIdx.x;       // 64 threads: one per output unit
  const double* w = wp + (long long)c * PB + n * 64;
  const int sw = (n & 1) * 8;
  double mx = 0.0;
  for (int k = 0; k < 64; ++k) mx = fmax(mx, fabs(w[k ^ sw]));
  const int f = oz_exponent(mx);
  uint8_t* out = wt + (long long)c * OZ_W1_BYTES;
  reinterpret_cast<double*>(out + OZ_S * OZ_WPLANE)[n] = scalbn(1.0, f - 48);   // 2^(f - 92) * 2^44: the helpers park the recombined sum scaled by 2^-44
  for (int k = 0; k < 64; ++k) {
    int8_t d[OZ_S];
    oz_digits(w[k ^ sw], f, d);
#pragma unroll
    for (int s = 0; s < OZ_S; ++s) out[s * OZ_WPLANE + oz_kmajor_offset(n, k)] = (uint8_t)d[s];
  }
}

__device__ __forceinline__ uint64_t oz_smem_desc(uint32_t saddr) {
  // cute::UMMA::SmemDescriptor: start address [0,14), LBO [16,30), SBO [32,46) (all >> 4), version 1 at [46,48),
  // layout type SWIZZLE_NONE (0) at [61,64)
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(OZ_LBO >> 4) << 16) | ((uint64_t)(OZ_SBO >> 4) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 16 TMEM lanes x 16 columns in the m16n8 accumulator-fragment layout (verified by tools/umma_probe.cu):
// v[4h + {0,1,2,3}] = (row g, col 8h+2t), (g, 8h+2t+1), (g+8, 8h+2t), (g+8, 8h+2t+1)
__device__ __forceinline__ void oz_ld_frag(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
// (hi * 65536 + mid * 256 + lo) as a double: |value| < 2^51, so adding it to the integer image of 1.5 * 2^52 and
// subtracting 1.5 * 2^52 converts exactly with ONE FP64 instruction (I2F.F64 runs at a fraction of the DFMA rate)
__device__ __forceinline__ double oz_group(int hi, int mid, int lo) {
  const long long v = (long long)hi * 65536 + (long long)mid * 256 + (long long)lo;
  return __longlong_as_double(v + 0x4338000000000000LL) - 6755399441055744.0;
}

// 32 TMEM lanes (thread = lane = row) x 8 consecutive 32-bit columns
__device__ __forceinline__ void oz_ld_row8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

constexpr int FWD3T_COMPUTE_WARPS = 8;     // 16 rows each: layers 2 / 3 (DMMA) + likelihood epilogue
constexpr int FWD3T_HELPER_WARPS = 4;      // one per TMEM lane quarter: drain, recombine, activate layer 1
constexpr int FWD3T_CONTROL_WARP = 12;     // lane 0: weight streaming + MMA issue (warps 13-15 only donate registers)
constexpr int FWD3T_THREADS = 16 * 32;
// registers per thread after setmaxnreg (the kernel is launched with 512 x 128): 8 x 168 + 4 x 128 + 4 x 40 <= 2048
constexpr int FWD3T_REGS_COMPUTE = 168, FWD3T_REGS_CONTROL = 40;

// activated layer-1 outputs of the 128-row tile in shared memory: row r, 16-byte chunk c (columns 2c, 2c+1) at
// r * 512 + ((c ^ s(r)) * 16), s(r) = ((r & 1) << 2) | ((r >> 1) & 3): conflict-free both for the helpers'
// stores (lane = row, same chunk) and for the compute warps' A-fragment loads (rows g, chunks 4kg + t)
__device__ __forceinline__ int a1_swz(int r) { return ((r & 1) << 2) | ((r >> 1) & 3); }

__device__ __forceinline__ void oz_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void oz_ld_row16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void oz_st_row16(uint32_t taddr, const uint4 (&w)[4]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(w[0].x), "r"(w[0].y), "r"(w[0].z), "r"(w[0].w), "r"(w[1].x), "r"(w[1].y), "r"(w[1].z), "r"(w[1].w), "r"(w[2].x),
      "r"(w[2].y), "r"(w[2].z), "r"(w[2].w), "r"(w[3].x), "r"(w[3].y), "r"(w[3].z), "r"(w[3].w)
      : "memory");
}

// tuning instrumentation (build with -DBNN_DBG_WAITCLK): clocks spent per wait site / phase, summed over lane 0 of
// every warp; slots 0-15 helper 0 (control), 16-31 helpers 1-3, 32-47 compute warps.  Read and reset with bnn_debug_counters().
__device__ unsigned long long g_dbg_clk[48];
__device__ unsigned long long* g_dbg_trace = nullptr;   // BNN_DBG_CSUM: [CTA][use][warp][2] checksums (a1 read, logits)
cudaError_t bnn_debug_set_trace_ptr(unsigned long long* dev_ptr) { return cudaMemcpyToSymbol(g_dbg_trace, &dev_ptr, sizeof(dev_ptr)); }
#ifdef BNN_DBG_WAITCLK
#define DBG_T0() const long long dbg_t0 = clock64()
#define DBG_ADD(slot) dbg[slot] += clock64() - dbg_t0
#define DBG_WAIT(slot, bar, par) do { const long long t0__ = clock64(); mbar_wait_sleep(bar, par); dbg[slot] += clock64() - t0__; } while (0)
#else
#define DBG_T0()
#define DBG_ADD(slot)
#define DBG_WAIT(slot, bar, par) mbar_wait_sleep(bar, par)
#endif
cudaError_t bnn_debug_counters_read(unsigned long long* out48) {
  cudaError_t e = cudaMemcpyFromSymbol(out48, g_dbg_clk, sizeof(unsigned long long) * 48);
  if (e != cudaSuccess) return e;
  unsigned long long z[48] = {0};
  return cudaMemcpyToSymbol(g_dbg_clk, z, sizeof(z));
}

constexpr uint32_t OZ_TMEM_X = OZ_S * 64;     // TMEM columns [384, 480): X slices (A operand), 16 columns per slice

template <int ACT, int MODE>
__global__ void __launch_bounds__(FWD3T_THREADS, 1) k_fwd3t(const __grid_constant__ FwdParams p) {
  constexpr int KP0 = 64, N1 = 64, N2 = 32, N3 = 16;
  using G3 = Fwd3Geom<KP0, N1, N2, N3>;
  constexpr bool PREDICT = (MODE == FWD3_PRED);
  constexpr int TB = 8;                                            // 256-entry exp table (shared memory is short)
  constexpr int REST = G3::PB - G3::W2_OFF;                       // FP64 part for the compute warps: W2, b2, W3, b3
  constexpr uint32_t W1_BYTES = OZ_S * OZ_WPLANE;                  // slice planes of W1
  constexpr uint32_t RSLOT_BYTES = REST * 8;
  constexpr uint32_t HSLOT_BYTES = 2 * 64 * 8;                     // colscale[64] + b1[64] for the helpers
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const NetGeom& g = p.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;
  long long dbg[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  (void)dbg;

  // ---- shared memory carve-up
  uint8_t* w1ring = smem_raw;                                      // [2][W1_BYTES]
  uint8_t* rring = w1ring + 2 * W1_BYTES;                          // [2][RSLOT_BYTES]
  double* a1s = reinterpret_cast<double*>(rring + 2 * RSLOT_BYTES);  // [2][128][64] layer-1 outputs (double buffer)
  double* tab = a1s + 2 * 128 * 64;
  uint8_t* hring = reinterpret_cast<uint8_t*>(tab + (1 << TB));     // [4][HSLOT_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(hring + 4 * HSLOT_BYTES);
  uint64_t* w1full = bars;          // [2] W1 slices of a use have landed
  uint64_t* rfull = bars + 2;       // [2] colscale + FP64 part of a use have landed
  uint64_t* rempty = bars + 4;      // [2] all compute warps are done with the slot
  uint64_t* tfull = bars + 6;       // layer-1 accumulators of a use are complete in TMEM
  uint64_t* tfree = bars + 7;       // all helpers have drained the accumulators (and refreshed X at a tile change)
  uint64_t* a1full = bars + 8;      // [4][2] activated layer-1 rows of a lane quarter are in a1s buffer b
  uint64_t* a1free = bars + 16;     // [4][2] both compute warps of the quarter have consumed them
  uint64_t* hfull = bars + 24;      // [4] colscale + b1 of a use have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  int* cnt = reinterpret_cast<int*>(bars + 30);
  const int n_cnt = PREDICT ? 0 : p.C * (2 + 2 * g.K);
#ifdef BNN_DBG_CSUM            // debugging aid: checksums of the activation hand-off
  volatile unsigned long long* dbg_cs = reinterpret_cast<volatile unsigned long long*>(cnt + ((n_cnt + 3) & ~3));   // [2][8]
#endif

  for (int i = threadIdx.x; i < (1 << TB); i += blockDim.x) tab[i] = p.exp_tab_small[i];
  for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) cnt[i] = 0;
  if (threadIdx.x == 0) {
    rel[0] = rel[1] = 0;
    for (int i = 0; i < 2; ++i) { mbar_init(&w1full[i], 1); mbar_init(&rfull[i], 1); mbar_init(&rempty[i], FWD3T_COMPUTE_WARPS); }
    mbar_init(tfull, 1); mbar_init(tfree, FWD3T_HELPER_WARPS);
    for (int i = 0; i < 8; ++i) { mbar_init(&a1full[i], 1); mbar_init(&a1free[i], 2); }
    for (int i = 0; i < 4; ++i) mbar_init(&hfull[i], 1);
    mbar_fence_init();
  }
  if (warp == FWD3T_COMPUTE_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const long long n_iter = (p.n_tiles128 + gridDim.x - 1) / gridDim.x;
  const long long total_q = n_iter * p.C;            // weight-set uses, identical for every warp of the CTA

  if (warp >= FWD3T_CONTROL_WARP) {
    // ======================= control warp: lane 0 streams the weights and issues the MMAs; it owns no data, so
    // none of its waits delays a warp that computes.  Its warpgroup gives its registers to the compute warps.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FWD3T_REGS_CONTROL));
    if (warp == FWD3T_CONTROL_WARP && lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, S8 x S8, K-major, N = 64, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      auto load_w1 = [&](long long use) {
        const int b = (int)(use & 1), c = (int)(use % p.C);
        mbar_arrive_expect_tx(&w1full[b], W1_BYTES);
        bulk_g2s(w1ring + (size_t)b * W1_BYTES, p.wt + (size_t)c * OZ_W1_BYTES, W1_BYTES, &w1full[b]);
      };
      auto load_rest = [&](long long use) {
        const int b = (int)(use & 1), c = (int)(use % p.C);
        mbar_arrive_expect_tx(&rfull[b], RSLOT_BYTES);
        bulk_g2s(rring + (size_t)b * RSLOT_BYTES, p.wp + (size_t)c * G3::PB + G3::W2_OFF, RSLOT_BYTES, &rfull[b]);
      };
      auto load_hslot = [&](long long use) {
        const int b = (int)(use & 3), c = (int)(use % p.C);
        uint8_t* slot = hring + (size_t)b * HSLOT_BYTES;
        mbar_arrive_expect_tx(&hfull[b], HSLOT_BYTES);
        bulk_g2s(slot, p.wt + (size_t)c * OZ_W1_BYTES + W1_BYTES, 64 * 8, &hfull[b]);
        bulk_g2s(slot + 64 * 8, p.wp + (size_t)c * G3::PB + G3::B1_OFF, 64 * 8, &hfull[b]);
      };
      auto issue_mma = [&](long long use) {
#ifdef BNN_DBG_NOMMA          // tuning experiment only
        return;
#endif
        // all 21 slice pairs: A = X slices in TMEM, B = W1 slices of `use` in shared memory
        const uint64_t wdesc = oz_smem_desc(smem_u32(w1ring + (size_t)(use & 1) * W1_BYTES));
#pragma unroll 1
        for (int d = 0; d < OZ_S; ++d) {               // diagonal T = 10 - d -> TMEM columns [64 d, 64 d + 64)
          const int T = 2 * (OZ_S - 1) - d;
          uint32_t acc = 0;
          for (int sx = T - (OZ_S - 1); sx <= OZ_S - 1; ++sx) {
            const int sw = T - sx;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              oz_mma_ts(tmem + 64 * d, tmem + OZ_TMEM_X + 16 * sx + 8 * ks,
                        wdesc + (uint64_t)((sw * OZ_WPLANE + ks * 2 * OZ_LBO) >> 4), idesc, acc);
              acc = 1;
            }
          }
        }
      };
      for (long long u = 0; u < 2 && u < total_q; ++u) { load_w1(u); load_rest(u); load_hslot(u); }
      // use 0: the helpers have staged the first tile's X slices (phase 0 of tfree)
      mbar_wait_sleep(tfree, 0);
      mbar_wait_sleep(&w1full[0], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if ((long long)blockIdx.x < p.n_tiles128) issue_mma(0);
      oz_commit(tfull);
      for (long long q = 0; q + 1 < total_q; ++q) {
        const long long it = q / p.C;
        const int c = (int)(q - it * p.C);
        const long long tile = it * gridDim.x + blockIdx.x;
        // the MMAs of use q are complete: their W1 slot takes the slices of use q + 2 (every helper has finished
        // pass 1 of use q - 1, so the helper slot of use q - 2 is free as well)
        DBG_WAIT(1, tfull, (uint32_t)(q & 1));
        if (q + 2 < total_q) { load_w1(q + 2); load_hslot(q + 2); }
        // the helpers have drained use q (and staged the next tile's X slices at a tile change)
        DBG_WAIT(3, tfree, (uint32_t)((q + 1) & 1));
        DBG_WAIT(4, &w1full[(q + 1) & 1], (uint32_t)(((q + 1) >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long ntile = (c + 1 == p.C) ? tile + gridDim.x : tile;
        {
          DBG_T0();
          if (ntile < p.n_tiles128) issue_mma(q + 1);
          oz_commit(tfull);          // arrives when every MMA issued so far is complete (also with none issued)
          DBG_ADD(8);
        }
        // FP64 part of use q + 1 goes into the slot of use q - 1 once the compute warps have released it
        if (q >= 1) {
          DBG_WAIT(5, &rempty[(q + 1) & 1], (uint32_t)(((q - 1) >> 1) & 1));
          load_rest(q + 1);
        }
      }
    }
  } else if (warp >= FWD3T_COMPUTE_WARPS) {
    // ======================= helper warps: X slices -> TMEM, accumulators -> activated layer-1 outputs in a1s
    const int h = warp - FWD3T_COMPUTE_WARPS;                     // TMEM lane quarter (== warp & 3)
    const int r = 32 * h + lane;                                   // row inside the 128-row tile
    const uint32_t tlane = (uint32_t)(32 * h) << 16;
    // this thread's row of X slices (6 x 64 bytes, contiguous) -> TMEM columns [384, 480) of its lane
    auto stage_x = [&](long long tile) {
      const uint4* src = reinterpret_cast<const uint4*>(p.xsl + ((size_t)tile * 128 + r) * OZ_XROW);
#pragma unroll
      for (int sl = 0; sl < OZ_S; ++sl) {
        uint4 w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = __ldg(src + 4 * sl + i);
        oz_st_row16(tmem + tlane + OZ_TMEM_X + 16 * sl, w);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    };
    if ((long long)blockIdx.x < p.n_tiles128) stage_x(blockIdx.x);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(tfree);                       // phase 0 of tfree: the first tile's X is in TMEM
    long long q = 0;
    for (long long it = 0; it < n_iter; ++it) {
      const long long tile = it * gridDim.x + blockIdx.x;
      const bool tile_ok = tile < p.n_tiles128;
      const double rsc = tile_ok ? p.x_rowscale[tile * 128 + r] : 0.0;
      for (int c = 0; c < p.C; ++c, ++q) {
        const int b = (int)(q & 1);
        DBG_WAIT(0, &hfull[q & 3], (uint32_t)((q >> 2) & 1));
        DBG_WAIT(1, tfull, (uint32_t)(q & 1));
        if (q > 1) DBG_WAIT(2, &a1free[2 * h + b], (uint32_t)(((q >> 1) - 1) & 1));
        // ... and stay at most ONE weight set ahead of this quarter's compute warps (wait until they have consumed
        // use q - 1 from the other buffer).  Without this bound -- helpers two sets ahead at the moment the next
        // tile's X slices are stored to TMEM -- a few rows of ONE weight set per launch came out wrong (always the
        // third-from-last set of an early tile, rows of TMEM lane quarter 0; tools/race_probe.py).  In failing runs
        // the compute warps read exactly the activations the helper wrote (-DBNN_DBG_CSUM), so the fault is not in
        // this hand-off; any added delay (trace stores, weights read from global memory) hides it.  The cause is
        // not understood; with the bound 64 launches over 8 chain counts were clean.  The kernel is opt-in.
#ifndef BNN_DBG_NOLOCKSTEP
        if (q > 0) mbar_wait_sleep(&a1free[2 * h + (b ^ 1)], (uint32_t)(((q - 1) >> 1) & 1));
#endif
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint8_t* slot = hring + (size_t)(q & 3) * HSLOT_BYTES;
        const double* csc = reinterpret_cast<const double*>(slot);
        const double* b1 = reinterpret_cast<const double*>(slot + 64 * 8);
        const double al = (ACT == BNN_ACT_LEAKY && p.alpha) ? p.alpha[c * 3 + 0] : 0.0;
        double* arow = a1s + (size_t)b * 128 * 64 + r * 64;
        const int sw = a1_swz(r);
        const long long dbg_p1 = clock64(); (void)dbg_p1;
        // ---- pass 1: drain the accumulators (the tensor core is idle until this is done).  Integer instructions
        // only -- the FP64 pipe is busy with the compute warps' DMMAs, and every FP64 instruction here would
        // queue behind them: the 6 diagonals are folded into one 64-bit integer per element,
        // (a0 2^16 + a1 2^8 + a2) 2^20 + ((a3 2^16 + a4 2^8 + a5) >> 4)   (|.| < 2^60; the 4 dropped bits are
        // 2^-64 of full scale), which is parked in a1s
#ifdef BNN_DBG_NOHELPER       // tuning experiment only
        if (false)
#endif
#pragma unroll 1
        for (int cb = 0; cb < N1 / 16; ++cb) {
          uint32_t v[OZ_S][16];
#pragma unroll
          for (int d = 0; d < OZ_S; ++d) oz_ld_row16(tmem + tlane + 64 * d + 16 * cb, v[d]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            long long w[2];
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int e = 2 * i + e2;
              const long long hi = (long long)(int)v[0][e] * 65536 + (long long)(int)v[1][e] * 256 + (long long)(int)v[2][e];
              const long long lo = (long long)(int)v[3][e] * 65536 + (long long)(int)v[4][e] * 256 + (long long)(int)v[5][e];
              w[e2] = hi * 1048576 + (lo >> 4);
            }
            *reinterpret_cast<longlong2*>(arow + 2 * ((8 * cb + i) ^ sw)) = make_longlong2(w[0], w[1]);
          }
        }
        dbg[6] += clock64() - dbg_p1;
        // a new tile follows: its X slices replace the current ones (no MMA is in flight: tfull(q) was waited for)
        if (c + 1 == p.C && tile + gridDim.x < p.n_tiles128) {
          DBG_T0();
          stage_x(tile + gridDim.x);
          DBG_ADD(9);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tfree);
        const long long dbg_p2 = clock64(); (void)dbg_p2;
        // ---- pass 2 (under the MMAs of the next weight set): integer -> FP64, scale, bias, activation, in place
        // on this thread's own row.  v = vh 2^32 + vl with the two halves converted by the 2^52 trick.
#ifdef BNN_DBG_NOHELPER
        if (false)
#endif
#ifdef BNN_DBG_CSUM
        unsigned long long csum = 0;
#endif
#pragma unroll 1
        for (int cb = 0; cb < N1 / 8; ++cb) {
          double z[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const longlong2 ww = *reinterpret_cast<const longlong2*>(arow + 2 * ((4 * cb + i) ^ sw));
            const long long w2[2] = {ww.x, ww.y};
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int e = 2 * i + e2;
              const int vh = (int)(w2[e2] >> 32);
              const unsigned vl = (unsigned)w2[e2];
              const double dh = __hiloint2double(0x43300000, vh ^ 0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
              const double dl = __hiloint2double(0x43300000, (int)vl) - 4503599627370496.0;           // 2^52
              const double hsum = fma(dh, 4294967296.0, dl);
              z[e] = fma(hsum * rsc, csc[8 * cb + e], b1[8 * cb + e]);
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) z[e] = bnn_act<ACT, TB>(z[e], al, tab);
#ifdef BNN_DBG_CSUM
#pragma unroll
          for (int e = 0; e < 8; ++e) csum ^= (unsigned long long)__double_as_longlong(z[e]);
#endif
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<double2*>(arow + 2 * ((4 * cb + i) ^ sw)) = make_double2(z[2 * i], z[2 * i + 1]);
        }
#ifdef BNN_DBG_CSUM
        for (int o = 8; o > 0; o >>= 1) csum ^= __shfl_xor_sync(FULL_MASK, csum, o);
        if ((lane & 15) == 0) dbg_cs[b * 8 + 2 * h + (lane >> 4)] = csum;
#endif
        __syncwarp();
        dbg[7] += clock64() - dbg_p2;
        if (lane == 0) mbar_arrive(&a1full[2 * h + b]);
      }
    }
  } else {
    // ======================= compute warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FWD3T_REGS_COMPUTE));
    const int h = warp & 3;
    const int rbase = 32 * h + 16 * (warp >> 2);                // rows of this warp inside the 128-row tile
    long long q = 0;
    for (long long it = 0; it < n_iter; ++it) {
      const long long tile = it * gridDim.x + blockIdx.x;
      const long long wt = tile * 8 + (rbase >> 4);
      const bool have_tile = tile < p.n_tiles128 && wt < p.n_tiles16;
      int y[2] = {0, 0};
      double wgt[2] = {1.0, 1.0};
      double acc3[N3 / 8][4];
      bool prev_valid = false;
#pragma unroll
      for (int j = 0; j < N3 / 8; ++j) acc3[j][0] = acc3[j][1] = acc3[j][2] = acc3[j][3] = 0.0;
      if (have_tile) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const long long row = wt * 16 + gq + 8 * hh;
          if (row < p.n_total) {
            y[hh] = p.labels[row];
            if (MODE == FWD3_LIK_W) {
              if (p.class_w) wgt[hh] *= p.class_w[y[hh]];
              if (p.inst_w && row < p.n_train) wgt[hh] *= p.inst_w[row];
            }
          }
        }
      }
      for (int c = 0; c < p.C; ++c, ++q) {
        const int b = (int)(q & 1);
        DBG_WAIT(0, &rfull[b], (uint32_t)((q >> 1) & 1));
        DBG_WAIT(1, &a1full[2 * h + b], (uint32_t)((q >> 1) & 1));
        const long long dbg_c0 = clock64(); (void)dbg_c0;
        const double* W = reinterpret_cast<const double*>(rring + (size_t)b * RSLOT_BYTES) - G3::W2_OFF;
        const double a2 = (ACT == BNN_ACT_LEAKY && p.alpha) ? p.alpha[c * 3 + 1] : 0.0;
        // ---------------- layer 2: [16 x N1] x [N1 x N2], A operand = activated layer-1 rows from a1s.  The
        // likelihood epilogue of the previous weight set (latency-bound FP64 chains and shuffles) is cut into
        // stages between the MMA groups of the first k-groups.
        double acc2[N2 / 8][4];
#pragma unroll
        for (int j = 0; j < N2 / 8; ++j) {
          const double2 bb = *reinterpret_cast<const double2*>(W + G3::B2_OFF + 8 * j + 2 * t);
          acc2[j][0] = bb.x; acc2[j][1] = bb.y; acc2[j][2] = bb.x; acc2[j][3] = bb.y;
        }
        RowStats<N3> rs;
        LikRow lr;
        const int K = g.K;
#ifdef BNN_DBG_CSUM
        unsigned long long ccs = 0;
        const unsigned long long want_cs = dbg_cs[b * 8 + 2 * h + (warp >> 2)];
#endif
#ifdef BNN_DBG_NOCOMPUTE      // tuning experiment only
        if (p.C < 0)
#endif
        {
          const double* wr = W + G3::W2_OFF + gq * N1;
          const int sw = (gq & 1) * G3::SW1;
          const double* ar0 = a1s + (size_t)b * 128 * 64 + (rbase + gq) * 64;
          const double* ar1 = ar0 + 8 * 64;
          const int asw = a1_swz(gq);                            // rows g and g + 8 share the swizzle
#pragma unroll
          for (int kg = 0; kg < N1 / 8; ++kg) {
            const double2 alo = *reinterpret_cast<const double2*>(ar0 + 2 * ((4 * kg + t) ^ asw));
            const double2 ahi = *reinterpret_cast<const double2*>(ar1 + 2 * ((4 * kg + t) ^ asw));
#ifdef BNN_DBG_CSUM
            ccs ^= (unsigned long long)__double_as_longlong(alo.x) ^ (unsigned long long)__double_as_longlong(alo.y) ^
                   (unsigned long long)__double_as_longlong(ahi.x) ^ (unsigned long long)__double_as_longlong(ahi.y);
#endif
            if (!PREDICT) {
              if (kg == 0) qs_max<N3>(acc3, K, t, rs);
              else if (kg == 1) qs_exp<N3, 0, true, TB>(acc3, K, t, y, tab, rs);
              else if (kg == 2) qs_exp<N3, 1, true, TB>(acc3, K, t, y, tab, rs);
              else if (kg == 3) {
                qs_reduce<N3, true>(rs);
                lr = quad_lik_finish<N3, MODE == FWD3_LIK_W>(p, wt, lane, rs, y, wgt, prev_valid && have_tile);
              }
            }
            const int col = (8 * kg + 2 * t) ^ sw;
#pragma unroll
            for (int j = 0; j < N2 / 8; ++j) {
              const double2 bb = *reinterpret_cast<const double2*>(wr + j * 8 * N1 + col);
              dmma16x8x8(acc2[j], alo.x, ahi.x, alo.y, ahi.y, bb.x, bb.y);
            }
          }
        }
#ifdef BNN_DBG_CSUM
        for (int o = 16; o > 0; o >>= 1) ccs ^= __shfl_xor_sync(FULL_MASK, ccs, o);
        if (lane == 0 && ccs != want_cs && have_tile) {
          atomicAdd(&g_dbg_clk[40], 1ULL);
          g_dbg_clk[41] = (unsigned long long)q; g_dbg_clk[42] = (unsigned long long)warp; g_dbg_clk[43] = (unsigned long long)it;
        }
#endif
        // the activated layer-1 rows have been consumed: the helpers may overwrite this buffer (two weight sets on)
        __syncwarp();
        if (lane == 0) mbar_arrive(&a1free[2 * h + b]);
        dbg[2] += clock64() - dbg_c0;
        if (!PREDICT) quad_lik_commit<N3>(p, c - 1, wt, lane, cnt, lr, prev_valid && have_tile);
        // ---------------- layer 3: [16 x N2] x [N2 x N3]
#pragma unroll
        for (int j = 0; j < N3 / 8; ++j) {
          const double2 bb = *reinterpret_cast<const double2*>(W + G3::B3_OFF + 8 * j + 2 * t);
          acc3[j][0] = bb.x; acc3[j][1] = bb.y; acc3[j][2] = bb.x; acc3[j][3] = bb.y;
        }
#ifdef BNN_DBG_NOCOMPUTE
        if (p.C < 0)
#endif
        {
          const double* wr = W + G3::W3_OFF + gq * N2;
          const int sw = (gq & 1) * G3::SW2;
          act_tile<ACT, false, TB>(acc2[0], a2, tab);
#pragma unroll
          for (int kg = 0; kg < N2 / 8; ++kg) {
            if (kg + 1 < N2 / 8) act_tile<ACT, false, TB>(acc2[kg + 1], a2, tab);
            const int col = (8 * kg + 2 * t) ^ sw;
#pragma unroll
            for (int j = 0; j < N3 / 8; ++j) {
              const double2 bb = *reinterpret_cast<const double2*>(wr + j * 8 * N2 + col);
              dmma16x8x8(acc3[j], acc2[kg][0], acc2[kg][2], acc2[kg][1], acc2[kg][3], bb.x, bb.y);
            }
          }
        }
#ifdef BNN_DBG_CSUM
        {
          unsigned long long lcs = 0;
#pragma unroll
          for (int j = 0; j < N3 / 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) lcs ^= (unsigned long long)__double_as_longlong(acc3[j][e]) * (unsigned long long)(2 * (4 * j + e) + 1);
          for (int o = 16; o > 0; o >>= 1) lcs ^= __shfl_xor_sync(FULL_MASK, lcs, o);
          if (lane == 0 && g_dbg_trace) {
            unsigned long long* tr = g_dbg_trace + (((size_t)blockIdx.x * total_q + q) * 8 + warp) * 2;
            tr[0] = ccs; tr[1] = lcs;
          }
        }
#endif
        // weights of this use are no longer needed by this warp
        __syncwarp();
        if (lane == 0) mbar_arrive(&rempty[b]);
        dbg[3] += clock64() - dbg_c0;
        prev_valid = true;
      }
      if (have_tile && !PREDICT) {
        // drain the software pipeline: epilogue of the last weight set of this tile
        RowStats<N3> rs;
        quad_softmax_stats<N3, true, TB>(acc3, g.K, t, y, tab, rs);
        const LikRow lr = quad_lik_finish<N3, MODE == FWD3_LIK_W>(p, wt, lane, rs, y, wgt, true);
        quad_lik_commit<N3>(p, p.C - 1, wt, lane, cnt, lr, true);
      }
    }
  }
#ifdef BNN_DBG_WAITCLK
  if (lane == 0)
    for (int i = 0; i < 12; ++i) atomicAdd(&g_dbg_clk[(warp >= FWD3T_CONTROL_WARP ? 0 : warp >= FWD3T_COMPUTE_WARPS ? 16 : 32) + i], (unsigned long long)dbg[i]);
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == FWD3T_COMPUTE_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  if (n_cnt) {
    for (int i = threadIdx.x; i < n_cnt; i += blockDim.x)
      if (cnt[i]) atomicAdd(&p.counts[i], cnt[i]);
  }
}

cudaError_t bnn_launch_slice_x(const double* x, long long n_pad16, uint8_t* xsl, double* rowscale, long long n_tiles128,
                               int* flag, cudaStream_t st) {
  const long long rows = n_tiles128 * 128;
  k_slice_x<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(x, n_pad16, xsl, rowscale, rows, flag);
  return cudaGetLastError();
}
cudaError_t bnn_launch_slice_w1(const double* wp, int PB, uint8_t* wt, int n_sets, cudaStream_t st) {
  k_slice_w1<<<n_sets, 64, 0, st>>>(wp, PB, wt);
  return cudaGetLastError();
}
size_t bnn_slice_x_tile_bytes() { return OZ_XTILE; }
size_t bnn_slice_w1_bytes() { return OZ_W1_BYTES; }

static size_t fwd3t_smem_bytes(int C, int K) {
  using G3 = Fwd3Geom<64, 64, 32, 16>;
  return 2 * (size_t)(OZ_S * OZ_WPLANE) + 2 * (size_t)(G3::PB - G3::W2_OFF) * 8 + 2 * 128 * 64 * sizeof(double) +
         256 * sizeof(double) + 4 * 1024 + 30 * sizeof(uint64_t) + (size_t)C * (2 + 2 * K) * sizeof(int)
#ifdef BNN_DBG_CSUM
         + 256
#endif
      ;
}

template <int ACT, int MODE>
static cudaError_t launch_fwd3t(const FwdParams& p, int n_sms, cudaStream_t st) {
  auto kern = k_fwd3t<ACT, MODE>;
  const size_t smem = fwd3t_smem_bytes(p.C, p.g.K);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  if (smem > 232448) return cudaErrorInvalidConfiguration;
  int grid = (int)(p.n_tiles128 < n_sms ? p.n_tiles128 : n_sms);
  kern<<<grid, FWD3T_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

#endif  // BNN_EXPERIMENTAL_TENSOR_L1

// =============================================================================================
// host-side launchers
// =============================================================================================
template <int KP0, int N1, int N2, int N3, int NWARPS, int MODE, int LIKK>
static size_t fwd3_smem_bytes(const FwdParams& p) {
  constexpr bool DEFER = (MODE != FWD3_PRED) && (LIKK == FWD3_CAT);
  using G3 = Fwd3Geom<KP0, N1, N2, N3>;
  size_t d = 2 * (size_t)G3::PB + (size_t)NWARPS * 16 * KP0 + BNN_EXP_TAB_SIZE;
  size_t bytes = d * sizeof(double) + (6 + NWARPS) * sizeof(uint64_t);     // barriers + the ring's release counters
  size_t ints = DEFER ? (size_t)p.C * (2 + 2 * p.g.K) : 0;
  return bytes + ints * sizeof(int);
}

template <int ACT, int KP0, int N1, int N2, int N3, int NWARPS, int MODE, int LIKK>
static cudaError_t launch_fwd3(const FwdParams& p, int n_sms, cudaStream_t st) {
  auto kern = k_fwd3<ACT, KP0, N1, N2, N3, NWARPS, MODE, LIKK>;
  size_t smem = fwd3_smem_bytes<KP0, N1, N2, N3, NWARPS, MODE, LIKK>(p);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  if (smem > 232448) return cudaErrorInvalidConfiguration;
  long long ctas = (p.n_tiles16 + NWARPS - 1) / NWARPS;
  if (MODE != FWD3_PRED) ctas *= p.C;          // likelihood modes deal (tile group, weight set) pairs out over the CTAs
  int grid = (int)(ctas < n_sms ? ctas : n_sms);
  kern<<<grid, NWARPS * 32, smem, st>>>(p);
  return cudaGetLastError();
}

// k_fwd3 instantiations: two padded width families x four activations x {categorical (N3 = 16: likelihood, weighted
// likelihood, prediction), Gaussian / sigma head (N3 = 8: likelihood, prediction)}
//   family A   64 -> 64 -> 32 -> N3   12 warps per SM (BASELINE config 4 / 5 is its swish / categorical member)
//   family B   32 -> 32 -> 16 -> N3   16 warps per SM
template <int ACT, int KP0, int N1, int N2, int NWARPS>
static cudaError_t launch_fwd3_family(const FwdParams& p, bool predict, int n_sms, cudaStream_t st) {
  if (p.g.lik == BNN_LIK_CATEGORICAL) {
    if (predict) return launch_fwd3<ACT, KP0, N1, N2, 16, NWARPS, FWD3_PRED, FWD3_CAT>(p, n_sms, st);
    if (p.class_w || p.inst_w) return launch_fwd3<ACT, KP0, N1, N2, 16, NWARPS, FWD3_LIK_W, FWD3_CAT>(p, n_sms, st);
    return launch_fwd3<ACT, KP0, N1, N2, 16, NWARPS, FWD3_LIK, FWD3_CAT>(p, n_sms, st);
  }
  if (predict) return launch_fwd3<ACT, KP0, N1, N2, 8, NWARPS, FWD3_PRED, FWD3_GAUSS>(p, n_sms, st);
  return launch_fwd3<ACT, KP0, N1, N2, 8, NWARPS, FWD3_LIK, FWD3_GAUSS>(p, n_sms, st);
}

template <int KP0, int N1, int N2, int NWARPS>
static cudaError_t launch_fwd3_act(const FwdParams& p, bool predict, int n_sms, cudaStream_t st) {
  switch (p.g.act) {
    case BNN_ACT_RELU: return launch_fwd3_family<BNN_ACT_RELU, KP0, N1, N2, NWARPS>(p, predict, n_sms, st);
    case BNN_ACT_LEAKY: return launch_fwd3_family<BNN_ACT_LEAKY, KP0, N1, N2, NWARPS>(p, predict, n_sms, st);
    case BNN_ACT_SWISH: return launch_fwd3_family<BNN_ACT_SWISH, KP0, N1, N2, NWARPS>(p, predict, n_sms, st);
    default: return launch_fwd3_family<BNN_ACT_TANH, KP0, N1, N2, NWARPS>(p, predict, n_sms, st);
  }
}

// which k_fwd3 family the padded geometry matches exactly: 1 = A, 2 = B, 0 = none
int bnn_fwd3_family(const NetGeom& g) {
  if (g.L != 3) return 0;
  const int n3 = (g.lik == BNN_LIK_CATEGORICAL) ? 16 : 8;
  if (g.l[2].out_pad != n3) return 0;
  if (g.F_pad == 64 && g.l[0].out_pad == 64 && g.l[1].out_pad == 32) return 1;
  if (g.F_pad == 32 && g.l[0].out_pad == 32 && g.l[1].out_pad == 16) return 2;
  return 0;
}

template <int ACT, bool PREDICT>
static cudaError_t launch_generic_t(const FwdParams& p, int n_sms, cudaStream_t st) {
  auto kern = k_fwd_generic<ACT, PREDICT>;
  const int ZS = p.g.l[p.g.L - 1].out_pad + 1;
  const int PW = (p.g.lik == BNN_LIK_CATEGORICAL) ? p.g.K : p.g.O;
  size_t per_warp = 2 * 16 * (size_t)p.g.max_w + 16 * ZS + (PREDICT ? 16 * PW : 0);
  size_t bytes = (GEN_TAB_SIZE + GEN_WARPS * per_warp) * sizeof(double);
  size_t ints = PREDICT ? (size_t)GEN_WARPS * 16 * PW
                        : (p.g.lik == BNN_LIK_CATEGORICAL ? (size_t)p.C * (2 + 2 * p.g.K) : 0);
  bytes += ints * sizeof(int);
  if (bytes > 232448) return cudaErrorInvalidConfiguration;
  static size_t attr_bytes = 0;
  if (bytes > 48 * 1024 && bytes > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_bytes = 232448;
  }
  long long ctas = (p.n_tiles16 + GEN_WARPS - 1) / GEN_WARPS;
  // resident CTAs per SM limited by shared memory; cap the grid at a few waves of resident CTAs
  long long max_grid = (long long)n_sms * 8;
  int grid = (int)(ctas < max_grid ? ctas : max_grid);
  int ny = 1;
  if (!PREDICT && p.C > 1 && grid < n_sms * 4) {
    const int want = (n_sms * 4 + grid - 1) / grid;
    ny = want < p.C ? want : p.C;
  }
  kern<<<dim3(grid, ny), GEN_WARPS * 32, bytes, st>>>(p);
  return cudaGetLastError();
}

template <bool PREDICT>
static cudaError_t launch_generic(const FwdParams& p, int n_sms, cudaStream_t st) {
  switch (p.g.act) {
    case BNN_ACT_RELU: return launch_generic_t<BNN_ACT_RELU, PREDICT>(p, n_sms, st);
    case BNN_ACT_LEAKY: return launch_generic_t<BNN_ACT_LEAKY, PREDICT>(p, n_sms, st);
    case BNN_ACT_SWISH: return launch_generic_t<BNN_ACT_SWISH, PREDICT>(p, n_sms, st);
    default: return launch_generic_t<BNN_ACT_TANH, PREDICT>(p, n_sms, st);
  }
}

// shared-memory plan of k_fwd_sparse: chains per unit, chains per resident weight group and warps per CTA
static size_t sparse_fixed_bytes(const FwdParams& p, int C) {
  return BNN_EXP_TAB_SIZE * sizeof(double) +
         (p.g.lik == BNN_LIK_CATEGORICAL ? (size_t)C * (2 + 2 * p.g.K) * sizeof(int) : 0) +
         (size_t)(p.sp_prog_len + 8) * sizeof(int) + 16;
}
static size_t sparse_warp_bytes(const FwdParams& p, int nch) {
  return ((size_t)(p.g.F + nch * p.sp_slots) * SP_US + 32 * (SP_MAX_O + 1)) * sizeof(double);
}
static bool sparse_plan(const FwdParams& p, int C, int nch, int* group, int* nwarps) {
  const size_t cap = 232448;
  const size_t fixed = sparse_fixed_bytes(p, C), per_warp = sparse_warp_bytes(p, nch);
  const size_t per_chain = (size_t)p.sp_wlen * sizeof(double);
  const int c_up = (C + nch - 1) / nch * nch;
  const int max_warps = (nch == 4) ? 10 : 16;
  // Warps first: the kernel is latency-bound (two to three warps per scheduler), and a smaller resident chain group
  // only costs another sweep over the tiles (X comes from L2 / HBM at a few percent of the step).  The rest of the
  // shared memory then takes as many chains' weight streams as fit.
  for (int w = max_warps; w >= 1; --w) {
    if (fixed + (size_t)w * per_warp + nch * per_chain > cap) continue;
    size_t g = (cap - fixed - (size_t)w * per_warp) / per_chain;
    g = g / nch * nch;
    if (g > (size_t)c_up) g = c_up;
    *group = (int)g;
    *nwarps = w;
    return true;
  }
  return false;
}

template <int ACT, int NCH>
static cudaError_t launch_sparse_n(const FwdParams& p0, int n_sms, cudaStream_t st) {
  auto kern = k_fwd_sparse<ACT, NCH>;
  FwdParams p = p0;
  int group = 0, nwarps = 0;
  if (!sparse_plan(p, p.C, NCH, &group, &nwarps)) return cudaErrorInvalidConfiguration;
  p.sp_group = group;
  const long long n_tiles32 = (p.n_tiles16 + 1) >> 1;
  int grid = (int)(n_tiles32 < n_sms ? n_tiles32 : n_sms);
  // small problems: no more warps than units per CTA
  const long long units = ((n_tiles32 + grid - 1) / grid) * ((group + NCH - 1) / NCH);
  if (units < nwarps) nwarps = (int)units;
  const size_t bytes = sparse_fixed_bytes(p, p.C) + (size_t)group * p.sp_wlen * sizeof(double) +
                       (size_t)nwarps * sparse_warp_bytes(p, NCH);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  kern<<<grid, nwarps * 32, bytes, st>>>(p);
  return cudaGetLastError();
}

// the (units, units) shapes of uniform block pairs with a compiled instantiation of k_fwd_pairs
static int sparse_pair_shape(const FwdParams& p) {
  const int pr = p.sp_pair_nc1 > 0 ? p.sp_pair_nr1 * 16 + p.sp_pair_nr2 : 0;
  return (pr == 3 * 16 + 2 || pr == 2 * 16 + 1 || pr == 2 * 16 + 2 || pr == 4 * 16 + 2) ? pr : 0;
}
#ifndef SPARSE_PAIR_NCH
#define SPARSE_PAIR_NCH 2
#endif
// shared-memory plan of k_fwd_pairs: resident chains, warps per group, warps
static bool pairs_plan(const FwdParams& p, int nch, int* group, int* share, int* nwarps, size_t* bytes) {
  const size_t cap = 232448;
  const size_t n_cnt = (p.g.lik == BNN_LIK_CATEGORICAL) ? (size_t)p.C * (2 + 2 * p.g.K) : 0;
  const size_t fixed = BNN_EXP_TAB_SIZE * sizeof(double) + (((n_cnt + 3) & ~(size_t)3) + p.sp_prog_len + 4) * sizeof(int);
  const size_t per_chain = (size_t)p.sp_wlen * sizeof(double), tile_b = (size_t)p.g.F * SP_US * sizeof(double);
  const size_t zs_b = 32 * (SP_MAX_O + 1) * sizeof(double);
  const int max_warps = (nch == 1) ? 24 : 16;
  const int c_up = (p.C + nch - 1) / nch * nch;
  // all chains resident if they fit next to a useful number of warps, else as many as fit with the full set of warps
  for (int gch = c_up; gch >= nch; gch -= nch) {
    const int n_sub = gch / nch;
    int gw = n_sub < 8 ? n_sub : 8;
    while (max_warps % gw) --gw;                       // groups tile the CTA exactly
    for (int w = max_warps; w >= gw; w -= gw) {
      const size_t need = fixed + (size_t)gch * per_chain + (size_t)(w / gw) * tile_b + (size_t)w * zs_b;
      if (need <= cap && (w >= max_warps / 2 || gch == nch)) {
        *group = gch; *share = gw; *nwarps = w; *bytes = need;
        return true;
      }
    }
  }
  return false;
}
static bool sparse_uses_pairs(const FwdParams& p) {
  int group, share, nwarps;
  size_t bytes;
  return p.C >= 2 && sparse_pair_shape(p) && pairs_plan(p, SPARSE_PAIR_NCH, &group, &share, &nwarps, &bytes);
}
template <int ACT, int NR1, int NR2, int NC1 = 0>
static cudaError_t launch_pairs(const FwdParams& p0, int n_sms, cudaStream_t st) {
  if (NC1 == 0 && p0.sp_pair_nc1 == 1) return launch_pairs<ACT, NR1, NR2, 1>(p0, n_sms, st);
  auto kern = k_fwd_pairs<ACT, SPARSE_PAIR_NCH, NR1, NR2, NC1>;
  FwdParams p = p0;
  int group = 0, share = 1, nwarps = 0;
  size_t bytes = 0;
  if (!pairs_plan(p, SPARSE_PAIR_NCH, &group, &share, &nwarps, &bytes)) return cudaErrorInvalidConfiguration;
  p.sp_group = group;
  p.sp_share = share;
  const long long n_tiles32 = (p.n_tiles16 + 1) >> 1;
  int grid = (int)(n_tiles32 < n_sms ? n_tiles32 : n_sms);
  // small problems: no more groups than tiles per CTA
  const long long tiles_per_cta = (n_tiles32 + grid - 1) / grid;
  if (tiles_per_cta * share < nwarps) {
    nwarps = (int)tiles_per_cta * share;
    const size_t n_cnt = (p.g.lik == BNN_LIK_CATEGORICAL) ? (size_t)p.C * (2 + 2 * p.g.K) : 0;
    bytes = BNN_EXP_TAB_SIZE * sizeof(double) + (((n_cnt + 3) & ~(size_t)3) + p.sp_prog_len + 4) * sizeof(int) +
            (size_t)group * p.sp_wlen * sizeof(double) + (size_t)(nwarps / share) * p.g.F * SP_US * sizeof(double) +
            (size_t)nwarps * 32 * (SP_MAX_O + 1) * sizeof(double);
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  kern<<<grid, nwarps * 32, bytes, st>>>(p);
  return cudaGetLastError();
}

template <int ACT>
static cudaError_t launch_sparse_t(const FwdParams& p, int n_sms, cudaStream_t st) {
  if (sparse_uses_pairs(p)) {
    // uniform block pairs (create_mask with equal node counts per feature group): fused, register-resident path
    switch (sparse_pair_shape(p)) {
      case 3 * 16 + 2: return launch_pairs<ACT, 3, 2>(p, n_sms, st);
      case 2 * 16 + 1: return launch_pairs<ACT, 2, 1>(p, n_sms, st);
      case 2 * 16 + 2: return launch_pairs<ACT, 2, 2>(p, n_sms, st);
      default: return launch_pairs<ACT, 4, 2>(p, n_sms, st);
    }
  }
  int group, nwarps;
  if (p.C >= 3 && sparse_plan(p, p.C, 4, &group, &nwarps)) return launch_sparse_n<ACT, 4>(p, n_sms, st);
  return launch_sparse_n<ACT, 2>(p, n_sms, st);
}

// true when the block-sparse kernel can run this problem (output width, shared-memory footprint)
bool bnn_sparse_fits(const FwdParams& p) {
  if (p.g.L < 2 || p.g.O > SP_MAX_O) return false;
  int group, nwarps;
  return sparse_plan(p, BNN_MAX_SETS_PER_PASS, 2, &group, &nwarps);
}

static cudaError_t launch_sparse(const FwdParams& p, int n_sms, cudaStream_t st) {
  switch (p.g.act) {
    case BNN_ACT_RELU: return launch_sparse_t<BNN_ACT_RELU>(p, n_sms, st);
    case BNN_ACT_LEAKY: return launch_sparse_t<BNN_ACT_LEAKY>(p, n_sms, st);
    case BNN_ACT_SWISH: return launch_sparse_t<BNN_ACT_SWISH>(p, n_sms, st);
    default: return launch_sparse_t<BNN_ACT_TANH>(p, n_sms, st);
  }
}

// Dispatch: specialised kernel when the padded shape is one of the compiled instantiations.
// force_generic != 0 disables the specialised path (used by the tests to cross-check both kernels).
cudaError_t bnn_launch_forward(const FwdParams& p, bool predict, int n_sms, int force_generic, cudaStream_t st,
                               const char** which) {
  const NetGeom& g = p.g;
  if (p.sp_prog && !predict) {
    if (which) *which = sparse_uses_pairs(p) ? "k_fwd_sparse<pairs>" : "k_fwd_sparse";
    return launch_sparse(p, n_sms, st);
  }
  if (!force_generic && !p.samp_u) {
    static const char* const names[2][4] = {
        {"k_fwd3<relu,64,64,32>", "k_fwd3<leaky,64,64,32>", "k_fwd3<swish,64,64,32,16>", "k_fwd3<tanh,64,64,32>"},
        {"k_fwd3<relu,32,32,16>", "k_fwd3<leaky,32,32,16>", "k_fwd3<swish,32,32,16>", "k_fwd3<tanh,32,32,16>"}};
    const int fam = bnn_fwd3_family(g);
    if (fam) {
      if (which) *which = names[fam - 1][g.act];
      return fam == 1 ? launch_fwd3_act<64, 64, 32, FWD3_WARPS>(p, predict, n_sms, st)
                      : launch_fwd3_act<32, 32, 16, 16>(p, predict, n_sms, st);
    }
  }
  if (which) *which = "k_fwd_generic";
  return predict ? launch_generic<true>(p, n_sms, st) : launch_generic<false>(p, n_sms, st);
}
