// Persistent Metropolis-Hastings loop for small data sets (BASELINE configs 1 / 2: bnn_classify.py, bnn_regress.py --
// a few thousand rows, a few hundred weights): ONE launch runs n_steps iterations of MCMC.mh_step (BNN_env.py:381-532).
//
// With two launches per iteration (k_mh_update, k_fwd_generic) a step of such a chain costs ~30 us even when replayed
// from a CUDA graph: two kernel boundaries, the proposal travelling through L2 / HBM, and a 2.5 MB feature matrix
// re-read through L2 by 40 CTAs -- for 3.5 MFLOP of arithmetic.  Here a thread-block CLUSTER owns a chain:
//   * the leader CTA (cluster rank 0) runs the update body (bnn_mh_body.cuh: accept step s-1, adapt, propose step s,
//     prior, pack) -- the same code, thread count and reduction order as k_mh_update, so chains are bit-identical;
//   * barrier.cluster (release / acquire);
//   * every CTA of the cluster copies the packed proposal into shared memory and runs the generic forward body
//     (bnn_generic_body.cuh) over its share of the 16-row tiles, whose X rows stay RESIDENT in shared memory for the
//     whole launch when they fit (config 1: 2.5 MB over 16 CTAs); per-tile partials and the accuracy counters go
//     straight into the leader's shared memory (distributed shared memory stores / atomics);
//   * barrier.cluster; next step.
// Chains are independent, so the C clusters of a launch need not be co-resident.
#include <cooperative_groups.h>
#include "bnn_mh_body.cuh"
#include "bnn_generic_body.cuh"

namespace cg = cooperative_groups;

constexpr int CL_WARPS = 8;                 // 256 threads = the update body's thread count for networks of <= 2048 weights
constexpr int CL_THREADS = CL_WARPS * 32;

struct ChainLoopPlan {
  int cluster;            // CTAs per chain
  int tiles_per_warp;     // ceil(n_tiles16 / (cluster * CL_WARPS))
  int x_resident;         // X tiles kept in shared memory
  size_t smem;            // dynamic shared memory per CTA
};

// shared-memory carve-up (doubles unless noted), identical in every CTA of the cluster:
//   tab [BNN_EXP_TAB_SIZE] | w [PB] | alpha [BNN_MAX_LAYERS] | part [NF * n_tiles16] | stage [CL_WARPS * per_warp] |
//   x [CL_WARPS * tiles_per_warp * 16 * F_pad] (resident only) | cnt_leader [NC] (ints) | cnt_local [NC] (ints)
__host__ __device__ inline size_t chain_loop_per_warp(const NetGeom& g) {
  const int ZS = g.l[g.L - 1].out_pad + 1;
  return 2 * 16 * (size_t)g.max_w + 16 * (size_t)ZS;
}

template <int ACT, bool XRES>
__global__ void __launch_bounds__(CL_THREADS, 1) k_chain_loop(const __grid_constant__ ChainDev d,
                                                               const __grid_constant__ FwdParams p0, int n_steps,
                                                               int tiles_per_warp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ FwdParams sp;                  // this chain's forward parameters (partials -> the leader's shared memory)
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.dim_blocks().x;
  const int rank = (int)cluster.block_rank();
  const int c = (int)(blockIdx.x / CL);
  const NetGeom& g = d.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2;
  const int NC = 2 + 2 * g.K;
  const int ZS = g.l[g.L - 1].out_pad + 1;
  const long long nt = d.n_tiles16;

  double* tab = reinterpret_cast<double*>(smem_raw);
  double* wsm = tab + BNN_EXP_TAB_SIZE;
  double* alpha_sm = wsm + g.PB;
  double* part_sm = alpha_sm + ((BNN_MAX_LAYERS + 1) & ~1);
  double* stage = part_sm + (((size_t)d.NF * nt + 1) & ~(size_t)1);      // 16-byte aligned: the staging buffers take double2 accesses
  const size_t per_warp = chain_loop_per_warp(g);
  double* h0 = stage + warp * per_warp;
  double* h1 = h0 + 16 * g.max_w;
  double* zs = h1 + 16 * g.max_w;
  double* xsm = stage + CL_WARPS * per_warp;
  const size_t x_doubles = XRES ? (size_t)CL_WARPS * tiles_per_warp * 16 * g.F_pad : 0;
  int* cnt_leader = reinterpret_cast<int*>(xsm + x_doubles);
  int* cnt_local = cnt_leader + ((NC + 3) & ~3);

  double* part_leader = cluster.map_shared_rank(part_sm, 0);
  int* cnt_dst = cluster.map_shared_rank(cnt_leader, 0);

  for (int i = tid; i < BNN_EXP_TAB_SIZE; i += CL_THREADS) tab[i] = p0.exp_tab[i];
  if (tid == 0) {
    sp = p0;
    sp.C = 1;
    sp.part = part_leader;
    sp.counts = nullptr;
    sp.wp = wsm;
    sp.alpha = alpha_sm;
  }
  // this warp's tiles: (i * CL + rank) * CL_WARPS + warp, i = 0 .. tiles_per_warp - 1
  auto tile_of = [&](int i) -> long long { return ((long long)i * CL + rank) * CL_WARPS + warp; };
  if (XRES) {
    // X rows of the tiles this warp owns, once per launch (global layout == shared layout: swizzled rows)
    for (int i = 0; i < tiles_per_warp; ++i) {
      const long long wt = tile_of(i);
      if (wt >= nt) break;
      const double2* src = reinterpret_cast<const double2*>(p0.x + wt * 16 * (long long)g.F_pad);
      double2* dst = reinterpret_cast<double2*>(xsm + ((size_t)i * CL_WARPS + warp) * 16 * g.F_pad);
      for (int e = lane; e < 8 * g.F_pad; e += 32) dst[e] = __ldg(src + e);
    }
  }
  __syncthreads();
  const FwdParams& p = sp;
  const double2* wsrc = reinterpret_cast<const double2*>(d.wp_prop + (long long)c * g.PB);

  for (int s = 0; s <= n_steps; ++s) {
    if (rank == 0) mh_update_body<true>(d, c, s > 0 ? 1 : 0, s < n_steps ? 1 : 0, s, part_sm, cnt_leader);
    if (s == n_steps) break;
    cluster.sync();                          // the proposal of step s (packed weights, slopes, zeroed counters) is visible
    for (int i = tid; i < g.PB / 2; i += CL_THREADS) reinterpret_cast<double2*>(wsm)[i] = __ldcg(wsrc + i);
    if (tid < g.L) alpha_sm[tid] = __ldcg(d.alpha_fwd + (long long)c * g.L + tid);
    for (int i = tid; i < NC; i += CL_THREADS) cnt_local[i] = 0;
    __syncthreads();
    for (int i = 0; i < tiles_per_warp; ++i) {
      const long long wt = tile_of(i);
      if (wt >= nt) break;
      const double* xt = XRES ? xsm + ((size_t)i * CL_WARPS + warp) * 16 * g.F_pad : p.x + wt * 16 * (long long)g.F_pad;
      const double* xrow0 = xt + gq * g.F_pad;
      const double* xrow1 = xrow0 + 8 * g.F_pad;
      fwd_generic_layers<ACT, !XRES, false>(g, xrow0, xrow1, wsm, (ACT == BNN_ACT_LEAKY) ? alpha_sm : nullptr, h0, h1, zs,
                                            ZS, tab, lane);
      bnn_epilogue<false>(p, 0, wt, lane, zs, ZS, tab, cnt_local, nullptr, nullptr);
      __syncwarp();
    }
    __syncthreads();
    if (g.lik == BNN_LIK_CATEGORICAL)
      for (int i = tid; i < NC; i += CL_THREADS)
        if (cnt_local[i]) atomicAdd(&cnt_dst[i], cnt_local[i]);
    cluster.sync();                          // partials and counters of step s are in the leader's shared memory
  }
  // no CTA may exit while others can still address its shared memory
  cluster.sync();
}

static bool chain_loop_plan(const NetGeom& g, int NF, long long nt, int C, int n_sms, ChainLoopPlan* plan) {
  if (g.P > 2048 || nt < 1 || nt >= 4096 || g.PB % 2) return false;      // (k_mh_update would run 1,024 threads / slice partials)
  const size_t cap = 232448 - sizeof(FwdParams) - 4096;                   // static shared memory of the two bodies
  const int NC = 2 + 2 * g.K;
  const size_t fixed = (BNN_EXP_TAB_SIZE + (size_t)g.PB + ((BNN_MAX_LAYERS + 1) & ~1) + (((size_t)NF * nt + 1) & ~(size_t)1) +
                        CL_WARPS * chain_loop_per_warp(g)) * sizeof(double) + 2 * (size_t)((NC + 3) & ~3) * sizeof(int);
  if (fixed > cap) return false;
  // CTAs per chain: as many as there are tiles for (one tile per warp), within the GPU when all chains run at once
  int cl = 16;
  while (cl > 1 && ((long long)(cl / 2) * CL_WARPS >= nt || (long long)C * cl > n_sms)) cl >>= 1;
  const int tpw = (int)((nt + (long long)cl * CL_WARPS - 1) / ((long long)cl * CL_WARPS));
  if (tpw > 8) return false;                 // larger data: the grid-wide two-kernel path uses all SMs for every chain
  const size_t xb = (size_t)CL_WARPS * tpw * 16 * g.F_pad * sizeof(double);
  plan->cluster = cl;
  plan->tiles_per_warp = tpw;
  plan->x_resident = fixed + xb <= cap;
  plan->smem = fixed + (plan->x_resident ? xb : 0);
  return true;
}

bool bnn_chain_loop_fits(const NetGeom& g, int NF, long long nt, int C, int n_sms) {
  ChainLoopPlan plan;
  return chain_loop_plan(g, NF, nt, C, n_sms, &plan);
}

template <int ACT, bool XRES>
static cudaError_t launch_chain_loop_t(const ChainDev& d, const FwdParams& p, int n_steps, const ChainLoopPlan& plan,
                                       cudaStream_t st) {
  auto kern = k_chain_loop<ACT, XRES>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(232448 - sizeof(FwdParams) - 4096));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(d.C * plan.cluster));
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)plan.cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int tpw = plan.tiles_per_warp;
  return cudaLaunchKernelEx(&cfg, kern, d, p, n_steps, tpw);
}

template <int ACT>
static cudaError_t launch_chain_loop_a(const ChainDev& d, const FwdParams& p, int n_steps, const ChainLoopPlan& plan,
                                       cudaStream_t st) {
  return plan.x_resident ? launch_chain_loop_t<ACT, true>(d, p, n_steps, plan, st)
                         : launch_chain_loop_t<ACT, false>(d, p, n_steps, plan, st);
}

// n_steps MH iterations of every chain of d in one launch; cudaErrorNotSupported when the problem does not fit the
// persistent path (the caller then issues the per-step launch sequence)
cudaError_t bnn_launch_chain_loop(const ChainDev& d, const FwdParams& p, int n_steps, int n_sms, int max_cluster,
                                  cudaStream_t st, int* cluster_out) {
  ChainLoopPlan plan;
  if (!chain_loop_plan(d.g, d.NF, d.n_tiles16, d.C, n_sms, &plan)) return cudaErrorNotSupported;
  if (max_cluster >= 1 && plan.cluster > max_cluster) {
    // (option "chain_loop_cluster": smaller clusters, e.g. where 16-CTA clusters cannot be scheduled)
    while (plan.cluster > max_cluster) plan.cluster >>= 1;
    ChainLoopPlan q = plan;
    const long long per = (long long)q.cluster * CL_WARPS;
    q.tiles_per_warp = (int)((d.n_tiles16 + per - 1) / per);
    if (q.tiles_per_warp > 8) return cudaErrorNotSupported;
    const size_t xb_old = plan.x_resident ? (size_t)CL_WARPS * plan.tiles_per_warp * 16 * d.g.F_pad * sizeof(double) : 0;
    const size_t fixed = plan.smem - xb_old;
    const size_t xb = (size_t)CL_WARPS * q.tiles_per_warp * 16 * d.g.F_pad * sizeof(double);
    q.x_resident = fixed + xb <= 232448 - sizeof(FwdParams) - 4096;
    q.smem = fixed + (q.x_resident ? xb : 0);
    plan = q;
  }
  if (cluster_out) *cluster_out = plan.cluster;
  switch (d.g.act) {
    case BNN_ACT_RELU: return launch_chain_loop_a<BNN_ACT_RELU>(d, p, n_steps, plan, st);
    case BNN_ACT_LEAKY: return launch_chain_loop_a<BNN_ACT_LEAKY>(d, p, n_steps, plan, st);
    case BNN_ACT_SWISH: return launch_chain_loop_a<BNN_ACT_SWISH>(d, p, n_steps, plan, st);
    default: return launch_chain_loop_a<BNN_ACT_TANH>(d, p, n_steps, plan, st);
  }
}
