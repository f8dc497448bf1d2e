// K2: proposal generation, log-prior, Metropolis-Hastings accept/reject and state commit, plus the
// small layout kernels (X / weight packing) and the deterministic reduction of the forward partials.
//
// Replaces (reference file:line): UpdateNormal (BNN_mcmc.py:57-69), the proposal / mask / prior /
// accept / adaptation / acceptance-window parts of MCMC.mh_step (BNN_env.py:392-413, 446-466, 481,
// 492-530) and npBNN.calc_prior (BNN_env.py:180-194).
#include "bnn_common.cuh"
#include "bnn_kernels.h"

static constexpr double kLogSqrt2Pi = 0.91893853320467274178;
static constexpr double kLogPi = 1.14472988584940017414;
static constexpr double kLog2 = 0.69314718055994530942;
#define UPD_THREADS 1024      // upper bound (launch bounds); small networks launch 256 (upd_threads)

// ------------------------------------------------------------------------------------------------
// layout kernels
// ------------------------------------------------------------------------------------------------
// canonical X [n, F] -> swizzled, padded rows [n_pad16, F_pad]; optional column override (PDP)
__global__ void k_pack_x(const double* __restrict__ x, double* __restrict__ xs, long long n, long long n_pad, int F,
                         int F_pad, int swz, const int* __restrict__ ov_cols, const double* __restrict__ ov_vals,
                         int n_ov) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = n_pad * F_pad;
  if (i >= total) return;
  long long r = i / F_pad;
  int c = (int)(i % F_pad);
  double v = 0.0;
  if (r < n && c < F) {
    v = x[r * F + c];
    for (int k = 0; k < n_ov; ++k)
      if (ov_cols[k] == c) v = ov_vals[k];
  }
  xs[r * F_pad + (c ^ ((int)(r & 1) * swz))] = v;
}

__device__ __forceinline__ int layer_of(const NetGeom& g, int i) {
  int l = 0;
#pragma unroll 1
  for (int k = 1; k < g.L; ++k)
    if (i >= g.l[k].c_off) l = k;
  return l;
}

// canonical weight sets [n_sets, P] -> packed [n_sets, PB] (padding entries are pre-zeroed, never written)
__global__ void k_pack_w(const __grid_constant__ NetGeom g, const double* __restrict__ w, double* __restrict__ wp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.P) return;
  int l = layer_of(g, i);
  const LayerGeom& lg = g.l[l];
  int cols = lg.in + lg.bias;
  int r = (i - lg.c_off) / cols, cc = (i - lg.c_off) % cols;
  wp[(long long)blockIdx.y * g.PB + bnn_packed_index(lg, r, cc)] = w[(long long)blockIdx.y * g.P + i];
}

// ------------------------------------------------------------------------------------------------
// deterministic block reductions (fixed order: per-thread strided sum -> xor tree -> warps in order)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_fixed(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
  return s;
}

// Sum of src[tid], src[tid + B], src[tid + 2B], ... in exactly that order, with the loads of eight terms issued
// before the first add: the plain loop is one L2 round trip per term (62,500 partials per chain at 1M rows).
__device__ __forceinline__ double strided_sum_ordered(const double* __restrict__ src, long long nt) {
  const long long B = blockDim.x;
  long long i = threadIdx.x;
  double v = 0.0;
  for (; i + 7 * B < nt; i += 8 * B) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __ldg(src + i + k * B);
#pragma unroll
    for (int k = 0; k < 8; ++k) v += a[k];
  }
  for (; i < nt; i += B) v += __ldg(src + i);
  return v;
}

__device__ __forceinline__ double logpdf_prior(double w, int kind, double scale, double log_scale) {
  // closed forms of scipy.stats.{norm,cauchy,laplace}.logpdf(w, 0, scale) (BNN_env.py:139-150)
  double x = w / scale;
  if (kind == BNN_PRIOR_CAUCHY) return -kLogPi - log1p(x * x) - log_scale;
  if (kind == BNN_PRIOR_LAPLACE) return -kLog2 - fabs(x) - log_scale;
  return -0.5 * x * x - kLogSqrt2Pi - log_scale;
}

// Reduce the per-warp-tile partials of chain `c` and turn them into the log-likelihood.
//   red : shared [1 + 3*BNN_MAX_OUT] receives the reduced slots; sig_out [K] the sigma that was used
__device__ double finalize_loglik(const NetGeom& g, const double* __restrict__ part, int NF, long long nt, int c,
                                  long long n_train, double lik_temp, int sigma_mode,
                                  const double* __restrict__ sigma_in, double* red, double* sig_out, double* sh) {
  for (int slot = 0; slot < NF; ++slot) {
    const double* src = part + ((long long)c * NF + slot) * nt;
    const double v = strided_sum_ordered(src, nt);
    double s = block_sum_fixed(v, sh);
    if (threadIdx.x == 0) red[slot] = s;
  }
  __syncthreads();
  double ll = red[0];
  if (g.lik == BNN_LIK_GAUSSIAN) {
    // calc_likelihood_regression (BNN_lib.py:123-131) from sum r, sum r^2:
    //   sum_i logN(y_i; mu_i, s) = -SSR/(2 s^2) - N log s - N log sqrt(2 pi)
    ll = 0.0;
    const double N = (double)n_train;
    for (int j = 0; j < g.K; ++j) {
      double sr = red[1 + j], ssr = red[1 + g.K + j];
      double s;
      if (sigma_mode == BNN_SIGMA_EMPIRICAL) {
        double mu = sr / N;                       // np.std(y' - labels, axis=0) (BNN_env.py:475-476)
        s = sqrt(fmax(ssr / N - mu * mu, 0.0));
      } else {
        s = sigma_in ? sigma_in[j] : 1.0;
      }
      if (threadIdx.x == 0) sig_out[j] = s;
      ll += -ssr / (2.0 * s * s) - N * log(s) - N * kLogSqrt2Pi;
    }
  }
  __syncthreads();
  return lik_temp * ll;
}

// stateless scoring: loglik / sums for bnn_forward_lik
__global__ void __launch_bounds__(UPD_THREADS) k_finalize_lik(const __grid_constant__ NetGeom g, const double* part,
                                                               int NF, long long nt, long long n_train,
                                                               double lik_temp, int sigma_mode, const double* sigma,
                                                               double* loglik, double* sums, int set0) {
  __shared__ double red[1 + 3 * BNN_MAX_OUT];
  __shared__ double sig[BNN_MAX_OUT];
  __shared__ double sh[32];
  const int c = blockIdx.x;
  const double* sg = sigma ? sigma + (long long)(set0 + c) * g.K : nullptr;
  double ll = finalize_loglik(g, part, NF, nt, c, n_train, lik_temp, sigma_mode, sg, red, sig, sh);
  if (threadIdx.x == 0) loglik[set0 + c] = ll;
  if (sums && g.lik != BNN_LIK_CATEGORICAL)
    for (int i = threadIdx.x; i < 3 * g.K; i += blockDim.x) sums[(long long)(set0 + c) * 3 * g.K + i] = red[1 + i];
}

// ps_entry / pls_entry: per-entry scales and their logs (hyper-priors: npBNN._prior_scale holds a vector per input
// node or a matrix per weight after sample_prior_scale, BNN_env.py:196-219), set `entry_stride` doubles apart
// (0: the same scales for every set); null: one scale per layer.
__global__ void __launch_bounds__(UPD_THREADS) k_log_prior(const __grid_constant__ NetGeom g, const double* w, int prior,
                                                            PriorScales ps, const double* ps_entry,
                                                            const double* pls_entry, int entry_stride, double* out) {
  __shared__ double sh[32];
  const int c = blockIdx.x;
  double v = 0.0;
  if (prior != BNN_PRIOR_UNIFORM)
    for (int i = threadIdx.x; i < g.P; i += blockDim.x) {
      int l = layer_of(g, i);
      const long long e = (long long)c * entry_stride + i;
      v += logpdf_prior(w[(long long)c * g.P + i], prior, ps_entry ? ps_entry[e] : ps.s[l], ps_entry ? pls_entry[e] : ps.ls[l]);
    }
  double s = block_sum_fixed(v, sh);
  if (threadIdx.x == 0) out[c] = s;
}

// MCMC.gibbs_step (BNN_env.py:534-538) after new prior scales: logPrior = calc_prior(), logPost = logLik + logPrior
// of the CURRENT weights of every chain.
__global__ void __launch_bounds__(UPD_THREADS) k_prior_refresh(const __grid_constant__ ChainDev d) {
  __shared__ double sh[32];
  const int c = blockIdx.x;
  const NetGeom& g = d.g;
  double v = 0.0;
  if (d.cfg.prior != BNN_PRIOR_UNIFORM)
    for (int i = threadIdx.x; i < g.P; i += blockDim.x) {
      int l = layer_of(g, i);
      const long long e = (long long)c * g.P + i;
      v += logpdf_prior(d.w_cur[e], d.cfg.prior, d.ps_entry ? d.ps_entry[e] : d.ps.s[l], d.ps_entry ? d.pls_entry[e] : d.ps.ls[l]);
    }
  double s = block_sum_fixed(v, sh);
  if (d.ind_cur && d.cfg.use_indicators) {
    // calc_prior's indicator term (BNN_env.py:191-193), as in k_mh_update: gibbs_step recomputes the whole log-prior
    const int P0 = g.l[0].out * (g.l[0].in + g.l[0].bias);
    const double* ind = d.ind_cur + (long long)c * P0;
    double n1 = 0.0;
    for (int i = threadIdx.x; i < P0; i += (int)blockDim.x) n1 += ind[i];
    n1 = block_sum_fixed(n1, sh);
    s += n1 * log(d.cfg.prior_ind1) + ((double)P0 - n1) * log(1.0 - d.cfg.prior_ind1);
  }
  if (threadIdx.x == 0) {
    double* sf = d.sf + (long long)c * BNN_F_STRIDE;
    sf[BNN_F_LOGPRIOR] = s;
    sf[BNN_F_LOGPOST] = sf[BNN_F_LOGLIK] + s;
  }
}

struct Draw { int ix, iy; double dz; };
// k-th proposal of layer l of chain c at iteration it
__device__ __forceinline__ Draw philox_draw(uint64_t seed, int c, int it, int l, int k, int rows, int cols, double ws) {
  uint2 key = make_uint2((uint32_t)seed ^ (uint32_t)c, (uint32_t)(seed >> 32));
  uint4 a = philox4x32(make_uint4((uint32_t)it, (uint32_t)k, (uint32_t)l, 0u), key);
  uint4 b = philox4x32(make_uint4((uint32_t)it, (uint32_t)k, (uint32_t)l, 1u), key);
  Draw d;
  d.ix = (int)__umulhi(a.x, (uint32_t)rows);
  d.iy = (int)__umulhi(a.y, (uint32_t)cols);
  double u1 = 1.0 - u53(a.z, a.w);            // (0,1]
  double u2 = u53(b.x, b.y);
  double s, co;
  sincospi(2.0 * u2, &s, &co);
  d.dz = ws * sqrt(-2.0 * log(u1)) * co;
  return d;
}

// ------------------------------------------------------------------------------------------------
// one CTA per chain: [accept previous proposal] + [adapt, propose, prior, pack]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UPD_THREADS) k_mh_update(const __grid_constant__ ChainDev d, int accept_mode,
                                                            int propose_mode, int step) {
  __shared__ double red[1 + 3 * BNN_MAX_OUT];
  __shared__ double sig[BNN_MAX_OUT];
  __shared__ double sh[32];
  __shared__ int s_flag;
  __shared__ int s_prop[BNN_MAX_LAYERS], s_cnt[BNN_MAX_LAYERS], s_off[BNN_MAX_LAYERS];
  // on-device generator (free-running chains): indicator moves decided by thread 0, flip probabilities for the block
  __shared__ int s_ind_move, s_fi_move;
  __shared__ double s_ind_p, s_fi_p;
  const NetGeom& g = d.g;
  const int c = blockIdx.x, tid = threadIdx.x;
  // The chain's scalar state is staged in shared memory for the whole launch: the accept / adapt / propose logic is
  // a serial chain of ~100 reads and writes by one thread, each of which would otherwise be an L2 round trip.
  __shared__ double ssf[BNN_F_STRIDE];
  __shared__ int ssi[BNN_I_STRIDE];
  double* const gsf = d.sf + (long long)c * BNN_F_STRIDE;
  int* const gsi = d.si + (long long)c * BNN_I_STRIDE;
  for (int i = tid; i < BNN_F_STRIDE; i += (int)blockDim.x) ssf[i] = gsf[i];
  for (int i = tid; i < BNN_I_STRIDE; i += (int)blockDim.x) ssi[i] = gsi[i];
  __syncthreads();
  double* sf = ssf;
  int* si = ssi;
  auto write_back = [&]() {
    __syncthreads();
    for (int i = tid; i < BNN_F_STRIDE; i += (int)blockDim.x) gsf[i] = ssf[i];
    for (int i = tid; i < BNN_I_STRIDE; i += (int)blockDim.x) gsi[i] = ssi[i];
  };
  double* wc = d.w_cur + (long long)c * g.P;
  double* wn = d.w_prop + (long long)c * g.P;
  const int NC = 2 + 2 * g.K;
  const int P0 = g.l[0].out * (g.l[0].in + g.l[0].bias);      // size of the first weight matrix (indicator shape)

  // ------------------------------------------------------------------ accept / reject
  if (accept_mode) {
    const bool smode_is_empirical = (accept_mode != 2) && d.cfg.sigma_mode == BNN_SIGMA_EMPIRICAL;
    // the initial likelihood of MCMC.__init__ always uses the stored error_prm (ones), even with
    // empirical_error=True (BNN_env.py:313-319); the empirical std only enters in mh_step (:475-476)
    const int smode = (accept_mode == 2) ? BNN_SIGMA_FIXED : d.cfg.sigma_mode;
    const double* sg_in = (g.lik == BNN_LIK_GAUSSIAN && smode == BNN_SIGMA_FIXED) ? sf + BNN_F_SIGMA : nullptr;
    double ll = finalize_loglik(g, d.part, d.NF, d.n_tiles16, c, d.n_train, d.cfg.lik_temp, smode, sg_in, red, sig, sh);
    if (d.cfg.sample_from_prior) ll = 0.0;
    if (tid == 0) {
      double lp = sf[BNN_F_LOGPRIOR_PROP];
      double post = ll + lp;
      sf[BNN_F_LOGLIK_PROP] = ll;
      // accept iff (logPost' - logPost) * T + hastings >= log u   (BNN_env.py:493-494); NaN compares false
      int acc = (accept_mode == 2) ? 1 : (((post - sf[BNN_F_LOGPOST]) * sf[BNN_F_TEMPERATURE] + 0.0 >= sf[BNN_F_LOG_U]) ? 1 : 0);
      s_flag = acc;
      if (acc) {
        sf[BNN_F_LOGLIK] = ll; sf[BNN_F_LOGPRIOR] = lp; sf[BNN_F_LOGPOST] = post;
        // ActFun.reset_accepted_prm (BNN_env.py:502-503)
        for (int l = 0; l < g.L; ++l) sf[BNN_F_ALPHA + l] = sf[BNN_F_ALPHA_PROP + l];
      }
      if (accept_mode == 1) {
        si[BNN_I_LAST_ACCEPTED] = acc;
        si[BNN_I_N_ACCEPTED] += acc;
        // acceptance window (BNN_env.py:523-529): mean over the stored outcomes + the new one, then keep 100
        int len = si[BNN_I_RING_LEN], head = si[BNN_I_RING_HEAD], sum = si[BNN_I_RING_SUM];
        sf[BNN_F_ACC_RATE] = (double)(sum + acc) / (double)(len + 1);
        if (len < 100) {
          si[BNN_I_RING + (head + len) % 100] = acc;
          si[BNN_I_RING_LEN] = len + 1;
          si[BNN_I_RING_SUM] = sum + acc;
        } else {
          si[BNN_I_RING_SUM] = sum + acc - si[BNN_I_RING + head];
          si[BNN_I_RING + head] = acc;
          si[BNN_I_RING_HEAD] = (head + 1) % 100;
        }
        si[BNN_I_ITERATION] += 1;
      }
    }
    __syncthreads();
    if (s_flag) {
      for (int i = tid; i < g.P; i += (int)blockDim.x) wc[i] = wn[i];
      // reset_indicators / _feature_indicators on accept (BNN_env.py:497-499)
      if (d.ind_cur)
        for (int i = tid; i < P0; i += (int)blockDim.x) d.ind_cur[(long long)c * P0 + i] = d.ind_prop[(long long)c * P0 + i];
      if (d.fi_cur)
        for (int i = tid; i < g.F; i += (int)blockDim.x) d.fi_cur[(long long)c * g.F + i] = d.fi_prop[(long long)c * g.F + i];
      if (g.lik == BNN_LIK_CATEGORICAL) {
        const int* cp = d.counts_prop + (long long)c * NC;
        if (tid < 2) si[BNN_I_N_CORRECT + tid] = cp[tid];
        for (int i = tid; i < g.K; i += (int)blockDim.x) {
          si[BNN_I_CLASS_CORRECT + i] = cp[2 + i];
          si[BNN_I_PRED_HIST + i] = cp[2 + g.K + i];
        }
      } else {
        for (int i = tid; i < g.K; i += (int)blockDim.x) {
          sf[BNN_F_SUM_R + i] = red[1 + i];
          sf[BNN_F_SUM_R2 + i] = red[1 + g.K + i];
          sf[BNN_F_SUM_R2_TEST + i] = red[1 + 2 * g.K + i];
          // reset_error_prm on accept, regression mode only (BNN_env.py:500-501)
          if (g.lik == BNN_LIK_GAUSSIAN && smode_is_empirical) sf[BNN_F_SIGMA + i] = sig[i];
        }
      }
    }
    __syncthreads();
  }
  if (!propose_mode) { write_back(); return; }

  // ------------------------------------------------------------------ adaptation + which layers
  const int it = si[BNN_I_ITERATION];
  if (tid == 0) {
    if (propose_mode == 1) {
      // BNN_env.py:392-413
      if (it % d.cfg.adapt_freq == 0 && it < d.cfg.adapt_stop) {
        double ar = sf[BNN_F_ACC_RATE];
        if (ar < d.cfg.adapt_f) {
          for (int l = 0; l < g.L; ++l) {
            sf[BNN_F_FREQ_LAYER + l] *= 0.8;
            sf[BNN_F_UPDATE_F + l] *= 0.85;
            int n = (int)((double)si[BNN_I_MAX_N + l] * sf[BNN_F_UPDATE_F + l]);
            si[BNN_I_UPDATE_N + l] = n < 1 ? 1 : n;
            sf[BNN_F_UPDATE_WS + l] *= 0.9;
          }
        }
        int tot = 0;
        for (int l = 0; l < g.L; ++l) tot += si[BNN_I_UPDATE_N + l];
        if (ar > d.cfg.adapt_fM && tot < g.P) {
          for (int l = 0; l < g.L; ++l) {
            sf[BNN_F_UPDATE_F + l] = exp(log(sf[BNN_F_UPDATE_F + l]) * 0.85);
            int n = (int)((double)si[BNN_I_MAX_N + l] * sf[BNN_F_UPDATE_F + l]);
            si[BNN_I_UPDATE_N + l] = n < 1 ? 1 : n;
            sf[BNN_F_UPDATE_WS + l] *= 1.2;
          }
        }
      }
      int off = 0;
      if (d.inj_proposed) {
        const long long base = ((long long)step * d.C + c) * g.L;
        for (int l = 0; l < g.L; ++l) {
          s_prop[l] = d.inj_proposed[base + l];
          s_cnt[l] = s_prop[l] ? d.inj_count[base + l] : 0;
          s_off[l] = off;
          off += s_cnt[l];
        }
        sf[BNN_F_LOG_U] = d.inj_logu[(long long)step * d.C + c];
      } else {
        // rr = rs.random(L); rr[argmin] = 0; layer proposed iff rr < freq_layer_update (BNN_env.py:446-451)
        uint2 key = make_uint2((uint32_t)d.cfg.seed ^ (uint32_t)(d.cfg.chain_offset + c), (uint32_t)(d.cfg.seed >> 32));
        double rr[BNN_MAX_LAYERS];
        int amin = 0;
        for (int l = 0; l < g.L; ++l) {
          uint4 r = philox4x32(make_uint4((uint32_t)it, 0xFFFFFFFFu, (uint32_t)l, 2u), key);
          rr[l] = u53(r.x, r.y);
          if (rr[l] < rr[amin]) amin = l;
        }
        rr[amin] = 0.0;
        // weight indicators (BNN_env.py:449-460): the first layer is proposed only if rr[0] >= freq_indicator, otherwise
        // its indicators move: UpdateBinomial(ind, update_f[3], shape) flips each entry with probability u * update_f[3]
        s_ind_move = 0;
        if (d.cfg.use_indicators && rr[0] < d.cfg.freq_indicator) {
          uint4 r = philox4x32(make_uint4((uint32_t)it, 0xFFFFFFFCu, 0u, 5u), key);
          s_ind_move = 1;
          s_ind_p = u53(r.x, r.y) * sf[BNN_F_UPDATE_F + 3];
        }
        // feature indicators (BNN_env.py:423-431): past adapt_stop, with probability 0.2, flips with probability u * 0.5
        s_fi_move = 0;
        if (d.cfg.use_feature_indicators && it > d.cfg.adapt_stop) {
          uint4 r = philox4x32(make_uint4((uint32_t)it, 0xFFFFFFFBu, 0u, 6u), key);
          if (u53(r.x, r.y) < 0.2) { s_fi_move = 1; s_fi_p = u53(r.z, r.w) * 0.5; }
        }
        for (int l = 0; l < g.L; ++l) {
          s_prop[l] = rr[l] < sf[BNN_F_FREQ_LAYER + l] && !(l == 0 && s_ind_move);
          s_cnt[l] = s_prop[l] ? si[BNN_I_UPDATE_N + l] : 0;
          s_off[l] = off;
          off += s_cnt[l];
        }
        uint4 r = philox4x32(make_uint4((uint32_t)it, 0xFFFFFFFEu, 0u, 3u), key);
        sf[BNN_F_LOG_U] = log(u53(r.x, r.y));
      }
      for (int l = 0; l < g.L; ++l) si[BNN_I_PROPOSED + l] = s_prop[l];
      // trainable activation parameters (BNN_env.py:416-421): UpdateNormal1D(_acc_prm, d=0.05, n=1, Mb=1, mb=0) with
      // the injected draw, both reflections over every entry (BNN_mcmc.py:46-56), Exp(10) term into additional_prob
      double addp = d.inj_add_prob ? d.inj_add_prob[(long long)step * d.C + c] : 0.0;
      for (int l = 0; l < g.L; ++l) sf[BNN_F_ALPHA_PROP + l] = sf[BNN_F_ALPHA + l];
      if (d.cfg.n_act_prm > 0 && (d.inj_alpha_ix || !d.inj_proposed)) {
        int ix;
        double adz;
        if (d.inj_alpha_ix) {
          ix = d.inj_alpha_ix[(long long)step * d.C + c];
          adz = d.inj_alpha_dz[(long long)step * d.C + c];
        } else {                                          // rs.integers(0, n, 1), rs.normal(0, 0.05, 1) on the device
          uint2 key = make_uint2((uint32_t)d.cfg.seed ^ (uint32_t)(d.cfg.chain_offset + c), (uint32_t)(d.cfg.seed >> 32));
          uint4 r = philox4x32(make_uint4((uint32_t)it, 0xFFFFFFFDu, 0u, 4u), key);
          ix = (int)__umulhi(r.x, (uint32_t)d.cfg.n_act_prm);
          double sn, cs;
          sincospi(2.0 * u53(r.y, r.z), &sn, &cs);
          uint4 r2 = philox4x32(make_uint4((uint32_t)it, 0xFFFFFFFDu, 1u, 4u), key);
          adz = 0.05 * sqrt(-2.0 * log(1.0 - u53(r2.x, r2.y))) * cs;
        }
        sf[BNN_F_ALPHA_PROP + ix] = sf[BNN_F_ALPHA + ix] + adz;
        double sum = 0.0;
        for (int l = 0; l < d.cfg.n_act_prm; ++l) {
          double z = sf[BNN_F_ALPHA_PROP + l];
          if (z > 1.0) z = 1.0 - (z - 1.0);
          if (z < 0.0) z = 0.0 + (0.0 - z);
          sf[BNN_F_ALPHA_PROP + l] = z;
          sum += z;
        }
        addp += log(10.0) * (-sum) * 10.0;
      }
      sf[BNN_F_ADD_PROB] = addp;
      for (int l = 0; l < g.L; ++l) d.alpha_fwd[(long long)c * g.L + l] = sf[BNN_F_ALPHA_PROP + l];
    } else {
      s_ind_move = 0; s_fi_move = 0;
      for (int l = 0; l < g.L; ++l) { s_prop[l] = 0; s_cnt[l] = 0; s_off[l] = 0; }
      // initial state (MCMC.__init__, BNN_env.py:313-320): stored parameters, init_additional_prob
      for (int l = 0; l < g.L; ++l) sf[BNN_F_ALPHA_PROP + l] = sf[BNN_F_ALPHA + l];
      sf[BNN_F_ADD_PROB] = d.cfg.init_additional_prob;
    }
  }
  // zero the proposal's counters (the forward kernel accumulates into them)
  if (g.lik == BNN_LIK_CATEGORICAL)
    for (int i = tid; i < NC; i += (int)blockDim.x) d.counts_prop[(long long)c * NC + i] = 0;
  for (int i = tid; i < g.P; i += (int)blockDim.x) wn[i] = wc[i];
  // indicator proposals: UpdateBinomial = |ind - flip| with the injected flips (BNN_mcmc.py:98-99), else unchanged
  const double* ind_p = nullptr;
  const double* fi_p = nullptr;
  __syncthreads();                                   // s_ind_move / s_fi_move and their probabilities are visible
  const bool gen_moves = propose_mode == 1 && !d.inj_proposed;          // free-running chains draw the flips here
  const uint2 fkey = make_uint2((uint32_t)d.cfg.seed ^ (uint32_t)(d.cfg.chain_offset + c), (uint32_t)(d.cfg.seed >> 32));
  if (d.ind_cur) {
    const long long sc = (long long)step * d.C + c;
    const bool mv = propose_mode == 1 && d.inj_ind_move && d.inj_ind_move[sc];
    double* dst = d.ind_prop + (long long)c * P0;
    for (int i = tid; i < P0; i += (int)blockDim.x) {
      double v = d.ind_cur[(long long)c * P0 + i];
      bool flip = mv && d.inj_ind_flip[sc * P0 + i];
      if (gen_moves && s_ind_move) {
        const uint4 r = philox4x32(make_uint4((uint32_t)it, (uint32_t)i, 0u, 7u), fkey);
        flip = u53(r.x, r.y) < s_ind_p;
      }
      if (flip) v = fabs(v - 1.0);
      dst[i] = v;
    }
    ind_p = dst;
  }
  if (d.fi_cur) {
    const long long sc = (long long)step * d.C + c;
    const bool mv = propose_mode == 1 && d.inj_fi_move && d.inj_fi_move[sc];
    double* dst = d.fi_prop + (long long)c * g.F;
    for (int i = tid; i < g.F; i += (int)blockDim.x) {
      double v = d.fi_cur[(long long)c * g.F + i];
      bool flip = mv && d.inj_fi_flip[sc * g.F + i];
      if (gen_moves && s_fi_move) {
        const uint4 r = philox4x32(make_uint4((uint32_t)it, (uint32_t)i, 0u, 8u), fkey);
        flip = u53(r.x, r.y) < s_fi_p;
      }
      if (flip) v = fabs(v - 1.0);
      dst[i] = v;
    }
    fi_p = dst;
  }
  __syncthreads();

  // ------------------------------------------------------------------ UpdateNormal (BNN_mcmc.py:57-69)
  // z[Ix,Iy] = z[Ix,Iy] + N(0, d): fancy assignment => for duplicate (ix,iy) the LAST draw wins and
  // increments are not accumulated.  owner[idx] = largest draw index touching idx.
  int* owner = d.owner + (long long)c * g.P;
  const long long inj_base = ((long long)step * d.C + c) * d.inj_cap;
  for (int pass = 0; pass < 3; ++pass) {
    for (int l = 0; l < g.L; ++l) {
      if (!s_prop[l]) continue;
      const LayerGeom& lg = g.l[l];
      const int cols = lg.in + lg.bias;
      const double ws = sf[BNN_F_UPDATE_WS + l];
      for (int k = tid; k < s_cnt[l]; k += (int)blockDim.x) {
        Draw dr;
        if (d.inj_proposed) {
          dr.ix = d.inj_ix[inj_base + s_off[l] + k];
          dr.iy = d.inj_iy[inj_base + s_off[l] + k];
          dr.dz = d.inj_dz[inj_base + s_off[l] + k];
        } else {
          dr = philox_draw(d.cfg.seed, d.cfg.chain_offset + c, it, l, k, lg.out, cols, ws);
        }
        const int idx = lg.c_off + dr.ix * cols + dr.iy;
        if (pass == 0) atomicMax(&owner[idx], k);
        else if (pass == 1) { if (owner[idx] == k) wn[idx] = wc[idx] + dr.dz; }
        else owner[idx] = -1;
      }
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------ reflect, mask, prior, pack
  const double hi = d.cfg.w_bound, lo = -d.cfg.w_bound;
  double lp = 0.0;
  double* wpk = d.wp_prop + (long long)c * g.PB;
  for (int i = tid; i < g.P; i += (int)blockDim.x) {
    const int l = layer_of(g, i);
    const LayerGeom& lg = g.l[l];
    double z = wn[i];
    if (propose_mode == 1) {
      if (s_prop[l]) {                 // single reflection at the bounds (BNN_mcmc.py:66-67)
        if (z > hi) z = hi - (z - hi);
        if (z < lo) z = lo + (lo - z);
      }
      if (d.mask) z *= d.mask[i];      // w' *= mask for every layer (BNN_env.py:461-462)
      wn[i] = z;
    }
    if (d.cfg.prior != BNN_PRIOR_UNIFORM) {
      const long long e = (long long)c * g.P + i;
      lp += logpdf_prior(z, d.cfg.prior, d.ps_entry ? d.ps_entry[e] : d.ps.s[l], d.ps_entry ? d.pls_entry[e] : d.ps.ls[l]);
    }
    const int cols = lg.in + lg.bias;
    const int r = (i - lg.c_off) / cols, cc = (i - lg.c_off) % cols;
    // the forward pass sees w0' * indicators' (BNN_env.py:463-466; the prior above does not), and a feature whose
    // indicator is 0 is replaced by its mean: its weight column leaves the contraction and enters the bias below
    if (l == 0 && ind_p) z *= ind_p[i];
    if (l == 0 && fi_p && !(lg.bias && cc == 0) && fi_p[cc - lg.bias] == 0.0) z = 0.0;
    wpk[bnn_packed_index(lg, r, cc)] = z;
  }
  double s = block_sum_fixed(lp, sh);
  if (ind_p && d.cfg.use_indicators) {
    // + sum(ind) log(pi1) + (size - sum(ind)) log(1 - pi1)   (BNN_env.py:191-193)
    double n1 = 0.0;
    for (int i = tid; i < P0; i += (int)blockDim.x) n1 += ind_p[i];
    n1 = block_sum_fixed(n1, sh);
    s += n1 * log(d.cfg.prior_ind1) + ((double)P0 - n1) * log(1.0 - d.cfg.prior_ind1);
  }
  if (fi_p) {
    // data_transform (BNN_env.py:14-17): x'[:, j] = mean_j where the feature indicator is 0, i.e. every first-layer
    // node gets the constant  sum_j mean_j * w0'[r, j] * ind'[r, j]  on top of its bias
    __syncthreads();
    const LayerGeom& l0 = g.l[0];
    const int cols0 = l0.in + l0.bias;
    for (int r = tid; r < l0.out; r += (int)blockDim.x) {
      double adj = 0.0;
      for (int j = 0; j < l0.in; ++j)
        if (fi_p[j] == 0.0) {
          const int e = r * cols0 + l0.bias + j;
          adj += d.feat_mean[j] * wn[e] * (ind_p ? ind_p[e] : 1.0);
        }
      const double b = l0.bias ? wn[r * cols0] * (ind_p ? ind_p[r * cols0] : 1.0) : 0.0;
      wpk[l0.b_off + r] = b + adj;
    }
  }
  if (tid == 0) sf[BNN_F_LOGPRIOR_PROP] = s + sf[BNN_F_ADD_PROB];     // calc_prior(...) + additional_prob (BNN_env.py:481)
  write_back();
}

// ------------------------------------------------------------------------------------------------
// launch wrappers (called from bnn_capi.cu)
// ------------------------------------------------------------------------------------------------
// threads per CTA of the one-CTA-per-chain kernels: the per-entry loops want the full 1024 at c4 size (6,474 entries,
// 62,500 partials), while at a few hundred entries the block-wide synchronisations of 32 warps cost more than they save
// (a function of the network only, so that sharded and unsharded runs of the same chains reduce in the same order)
static inline int upd_threads(const NetGeom& g) { return g.P > 2048 ? 1024 : 256; }

cudaError_t bnn_launch_pack_x(const double* x, double* xs, long long n, long long n_pad, int F, int F_pad, int swz,
                              const int* ov_cols, const double* ov_vals, int n_ov, cudaStream_t st) {
  long long total = n_pad * F_pad;
  if (total == 0) return cudaSuccess;
  k_pack_x<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, xs, n, n_pad, F, F_pad, swz, ov_cols, ov_vals, n_ov);
  return cudaGetLastError();
}
cudaError_t bnn_launch_pack_w(const NetGeom& g, const double* w, double* wp, int n_sets, cudaStream_t st) {
  dim3 grid((g.P + 255) / 256, n_sets);
  k_pack_w<<<grid, 256, 0, st>>>(g, w, wp);
  return cudaGetLastError();
}
cudaError_t bnn_launch_finalize_lik(const NetGeom& g, const double* part, int NF, long long nt, long long n_train,
                                    double lik_temp, int sigma_mode, const double* sigma, double* loglik, double* sums,
                                    int set0, int n_sets, cudaStream_t st) {
  k_finalize_lik<<<n_sets, upd_threads(g), 0, st>>>(g, part, NF, nt, n_train, lik_temp, sigma_mode, sigma, loglik, sums, set0);
  return cudaGetLastError();
}
cudaError_t bnn_launch_log_prior(const NetGeom& g, const double* w, int n_sets, int prior, const PriorScales& ps,
                                 const double* ps_entry, const double* pls_entry, int entry_stride, double* out,
                                 cudaStream_t st) {
  k_log_prior<<<n_sets, upd_threads(g), 0, st>>>(g, w, prior, ps, ps_entry, pls_entry, entry_stride, out);
  return cudaGetLastError();
}
cudaError_t bnn_launch_prior_refresh(const ChainDev& d, cudaStream_t st) {
  k_prior_refresh<<<d.C, upd_threads(d.g), 0, st>>>(d);
  return cudaGetLastError();
}
// ------------------------------------------------------------------------------------------------
// row sharding: the rows of X are split over the ranks, every rank runs the same chains.
// After the forward pass each rank reduces its per-tile partials to one vector per chain (k_rowshard_local, same
// fixed order as finalize_loglik); the vectors are summed over the ranks (one all-reduce of C * (NF + 2 + 2K)
// doubles); k_rowshard_commit stores the sums where k_mh_update expects partials of a single tile, so the accept
// step runs unchanged and every rank takes the same decision.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UPD_THREADS) k_rowshard_local(const double* __restrict__ part, int NF, long long nt,
                                                                 const int* __restrict__ counts, int NC,
                                                                 double* __restrict__ out) {
  __shared__ double sh[32];
  const int c = blockIdx.x;
  for (int slot = 0; slot < NF; ++slot) {
    const double* src = part + ((long long)c * NF + slot) * nt;
    const double v = strided_sum_ordered(src, nt);
    const double s = block_sum_fixed(v, sh);
    if (threadIdx.x == 0) out[(long long)c * (NF + NC) + slot] = s;
  }
  for (int i = threadIdx.x; i < NC; i += blockDim.x)
    out[(long long)c * (NF + NC) + NF + i] = counts ? (double)counts[(long long)c * NC + i] : 0.0;
}
__global__ void k_rowshard_commit(const double* __restrict__ in, int NF, int NC, double* __restrict__ part_red,
                                  int* __restrict__ counts) {
  const int c = blockIdx.x;
  for (int i = threadIdx.x; i < NF; i += blockDim.x) part_red[(long long)c * NF + i] = in[(long long)c * (NF + NC) + i];
  if (counts)
    for (int i = threadIdx.x; i < NC; i += blockDim.x)
      counts[(long long)c * NC + i] = (int)llrint(in[(long long)c * (NF + NC) + NF + i]);    // exact: integer-valued sums
}
cudaError_t bnn_launch_rowshard_local(const NetGeom& g, const double* part, int NF, long long nt, const int* counts, int NC,
                                      double* out, int n_chains, cudaStream_t st) {
  k_rowshard_local<<<n_chains, upd_threads(g), 0, st>>>(part, NF, nt, counts, NC, out);
  return cudaGetLastError();
}
cudaError_t bnn_launch_rowshard_commit(const double* in, int NF, int NC, double* part_red, int* counts, int n_chains,
                                       cudaStream_t st) {
  k_rowshard_commit<<<n_chains, 64, 0, st>>>(in, NF, NC, part_red, counts);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// First level of the partial-sum reduction.  The forward kernels leave one partial per (chain, slot, 16-row tile):
// 62,500 per chain at 1M rows, which k_mh_update -- one CTA per chain -- used to fold alone (18.5 MB through 32 SMs,
// 40 us per step, 2-3 % of a step at 4 chains per GPU).  Here C * NF * n_slices CTAs fold fixed slices of the tile
// axis in a fixed order (thread-strided terms, then the fixed block tree), so k_mh_update reads n_slices values per
// chain.  The slice boundaries depend on the tile count only: the result does not depend on how the forward kernel
// dealt its tiles out, nor on the number of chains in the pass.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reduce_part(const double* __restrict__ part, long long nt, int n_slices,
                                                      double* __restrict__ out) {
  __shared__ double sh[32];
  const int slice = blockIdx.x;
  const long long cs = blockIdx.y;                       // chain * NF + slot
  const long long per = (nt + n_slices - 1) / n_slices;
  const long long a = slice * per;
  long long b = a + per;
  if (b > nt) b = nt;
  const double v = (a < b) ? strided_sum_ordered(part + cs * nt + a, b - a) : 0.0;
  const double s = block_sum_fixed(v, sh);
  if (threadIdx.x == 0) out[cs * n_slices + slice] = s;
}

int bnn_part_slices(long long nt) {
  if (nt < 4096) return 0;                               // small data sets: k_mh_update reads the tile partials itself
  long long s = nt / 1024;
  return (int)(s > 64 ? 64 : s);
}

cudaError_t bnn_launch_reduce_part(const double* part, long long nt, int n_slices, int n_rows, double* out, cudaStream_t st) {
  dim3 grid((unsigned)n_slices, (unsigned)n_rows);
  k_reduce_part<<<grid, 256, 0, st>>>(part, nt, n_slices, out);
  return cudaGetLastError();
}

cudaError_t bnn_launch_mh_update(const ChainDev& d, int accept_mode, int propose_mode, int step, cudaStream_t st) {
  k_mh_update<<<d.C, upd_threads(d.g), 0, st>>>(d, accept_mode, propose_mode, step);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// FP64 peak probe (same loop as tools/peak_fp64.cu, m16n8k8): 8 independent accumulators per warp
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dmma_peak(double* out, int iters) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  const double a0 = 1.0 + threadIdx.x * 1e-6, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  const double b0 = 1e-9 + threadIdx.x * 1e-12, b1 = b0 * 2;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma16x8x8(c[i], a0, a1, a2, a3, b0, b1);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Labels index the class-probability row, the class weights and the per-class counters inside the forward epilogues
// (prediction[sample_id, labels], BNN_lib.py:104): a label outside [0, K) is an IndexError in the reference and would be
// an out-of-bounds access here, so bnn_set_data rejects it.
__global__ void k_check_labels(const int* __restrict__ labels, long long n_train, long long n_total, int K, int* flag) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  int bad = 0;
  for (; i < n_total; i += stride) {
    const int y = labels[i];
    bad |= (y < 0) || (y >= K);      // test labels feed the same counters (argmax == label), same range
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

cudaError_t bnn_launch_check_labels(const int* labels, long long n_train, long long n_total, int K, int* flag, cudaStream_t st) {
  long long blocks = (n_total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_check_labels<<<(int)blocks, 256, 0, st>>>(labels, n_train, n_total, K, flag);
  return cudaGetLastError();
}

cudaError_t bnn_measure_dmma_peak(int n_sms, double* tflops) {
  const int grid = n_sms * 2, threads = 256, iters = 2048;
  double* out = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(double) * grid * threads);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) k_dmma_peak<<<grid, threads>>>(out, iters);
  double best = 1e30;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k_dmma_peak<<<grid, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  e = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  if (e != cudaSuccess) return e;
  const double flops = 2.0 * 16 * 8 * 8 * (double)iters * 4 * 8 * (grid * threads / 32);
  *tflops = flops / (best * 1e-3) / 1e12;
  return cudaSuccess;
}
