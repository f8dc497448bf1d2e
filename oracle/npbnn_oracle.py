"""CPU oracle for the npBNN Metropolis-Hastings hot path -- TEST INFRASTRUCTURE ONLY.

This module is a plain-numpy restatement of the reference algorithm (dsilvestro/npBNN,
`np_bnn/`), written in explicit formulas (no scipy objects).  It exists to check the CUDA
path; nothing under `npbnn_b200/` may import it.  Allowed importers: `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs.

Parity status: PINNED.  Every function here is checked in `tests/test_oracle_golden.py`
against vectors produced by running the unmodified reference in the build container
(`tests/golden/make_golden.py`, which imports `/root/reference/np_bnn`); the reference
itself ships no tests.

Reference citations are `file:line` into the reference tree.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)
LOG_PI = math.log(math.pi)
LOG_2 = math.log(2.0)


# --------------------------------------------------------------------------------------
# activations  (BNN_lib.py:50-66, dispatch BNN_lib.py:68-87)
# --------------------------------------------------------------------------------------
def activation(z: np.ndarray, fun: str, alpha: float = 0.0) -> np.ndarray:
    """ReLU / leaky ("genReLU") / swish / tanh exactly as the reference evaluates them.

    * ReLU  : negative entries -> 0                                  (BNN_lib.py:50-52)
    * leaky : negative entries scaled by `alpha`                     (BNN_lib.py:54-56)
    * swish : z * (1 + exp(-z))**-1                                  (BNN_lib.py:58-61)
    * tanh  : 1 - 2 / (exp(2 z) + 1)  (not np.tanh)                  (BNN_lib.py:63-65)
    `alpha` is only honoured for fun == "genReLU" (BNN_lib.py:83-87).
    """
    if fun == "ReLU":
        return np.where(z < 0, 0.0, z)
    if fun == "genReLU":
        return np.where(z < 0, alpha * z, z)
    if fun == "swish":
        with np.errstate(over="ignore"):
            return z * (1.0 + np.exp(-z)) ** (-1)
    if fun == "tanh":
        with np.errstate(over="ignore"):
            return 1.0 - (2.0 / (np.exp(2.0 * z) + 1.0))
    raise ValueError("unknown activation %r" % (fun,))


def effective_act(fun: str, trainable: bool) -> str:
    """Which function ActFun ends up calling (BNN_lib.py:74-81: later `if`s override).

    fun="ReLU", trainable=True  -> leaky_relu_f is bound, but eval() passes prm 0 unless
    fun=="genReLU" (BNN_lib.py:83-87)  => behaves as ReLU.
    """
    if fun in ("swish", "tanh"):
        return fun
    if fun == "genReLU":
        return "genReLU"
    return "ReLU"


# --------------------------------------------------------------------------------------
# layers  (BNN_lib.py:154-162, 184-193)
# --------------------------------------------------------------------------------------
def dense(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """x @ w.T, with column 0 of `w` acting as bias when w has one more column than x
    (MatrixMultiplicationD, BNN_lib.py:154-162: presence of a bias is inferred from shapes)."""
    if x.shape[1] == w.shape[1]:
        return np.dot(x, w.T)
    return np.dot(x, w[:, 1:].T) + w[:, 0]


def softmax_rows(z: np.ndarray) -> np.ndarray:
    """scipy.special.softmax(z, axis=1) == exp(z - max) / sum(exp(z - max))  (BNN_lib.py:166-168)."""
    e = np.exp(z - np.max(z, axis=1, keepdims=True))
    return e / np.sum(e, axis=1, keepdims=True)


def softplus(z: np.ndarray) -> np.ndarray:
    """np.logaddexp(0, z)  (BNN_lib.py:170-172)."""
    return np.logaddexp(0.0, z)


def output_transform(z: np.ndarray, kind: str) -> np.ndarray:
    """'softmax' (classification, BNN_env.py:55), 'identity' (RegressTransform, BNN_lib.py:174),
    'regress-error' (softplus on the second half of the columns, BNN_lib.py:177-182)."""
    if kind == "softmax":
        return softmax_rows(z)
    if kind == "identity":
        return z
    if kind == "regress-error":
        h = z.shape[1] // 2
        out = z.copy()
        out[:, h:] = softplus(z[:, h:])
        return out
    raise ValueError(kind)


def forward(x: np.ndarray, weights: Sequence[np.ndarray], act: str = "ReLU",
            alphas: Optional[Sequence[float]] = None, out_kind: str = "softmax",
            indicators: Optional[np.ndarray] = None) -> np.ndarray:
    """RunPredict / RunPredictInd (BNN_lib.py:245-272): hidden layers with activation,
    last layer linear, then the output transform.  `indicators` multiplies layer-0 weights."""
    h = x
    n = len(weights)
    for i, w in enumerate(weights):
        if i == 0 and indicators is not None:
            w = w * indicators
        h = dense(h, w)
        if i < n - 1:
            a = 0.0
            if act == "genReLU" and alphas is not None:
                a = float(alphas[i])
            h = activation(h, act, a)
    return output_transform(h, out_kind)


# --------------------------------------------------------------------------------------
# likelihoods  (BNN_lib.py:100-143)
# --------------------------------------------------------------------------------------
def loglik_categorical(y: np.ndarray, labels: np.ndarray, class_w=None, inst_w=None,
                       lik_temp: float = 1.0) -> float:
    """lik_temp * sum_i log p[i, y_i] * cw[y_i] * iw_i   (BNN_lib.py:100-121).
    class_w together with inst_w raises in the reference (np.sum(1-D, axis=1), :105)."""
    with np.errstate(divide="ignore"):
        lp = np.log(y[np.arange(y.shape[0]), labels])
    has_cw = class_w is not None and len(class_w) > 0
    if has_cw and inst_w is not None:
        raise ValueError("class_weight with instance_weight: AxisError in the reference (BNN_lib.py:105)")
    if has_cw:
        lp = lp * np.asarray(class_w)[labels]
    if inst_w is not None:
        lp = lp * inst_w
    return float(lik_temp * np.sum(lp))


def norm_logpdf(x, mu, s):
    """scipy.stats.norm.logpdf(x, mu, s) = -((x-mu)/s)^2/2 - log(sqrt(2 pi)) - log s."""
    t = (x - mu) / s
    return -0.5 * t * t - LOG_SQRT_2PI - np.log(s)


def loglik_regression(y: np.ndarray, labels: np.ndarray, sig, lik_temp: float = 1.0) -> float:
    """lik_temp * sum_ij N(labels_ij; y_ij, sig_j)   (BNN_lib.py:123-131); sig scalar or [O]."""
    return float(lik_temp * np.sum(norm_logpdf(labels, y, sig)))


def loglik_regression_error(y: np.ndarray, labels: np.ndarray, lik_temp: float = 1.0) -> float:
    """mu = first O columns, sigma = last O columns  (BNN_lib.py:134-143)."""
    o = labels.shape[1]
    return float(lik_temp * np.sum(norm_logpdf(labels, y[:, :o], y[:, o:])))


# --------------------------------------------------------------------------------------
# prior  (BNN_env.py:135-154, 180-194)
# --------------------------------------------------------------------------------------
def logpdf_prior(w: np.ndarray, kind: int, scale) -> np.ndarray:
    """kind 1 (or anything not 0/2/3): Normal(0,s); 2: Cauchy(0,s); 3: Laplace(0,s).
    Closed forms of scipy.stats.{norm,cauchy,laplace}.logpdf (BNN_env.py:139-150)."""
    x = w / scale
    if kind == 2:
        return -LOG_PI - np.log1p(x * x) - np.log(scale)
    if kind == 3:
        return -LOG_2 - np.abs(x) - np.log(scale)
    return -0.5 * x * x - LOG_SQRT_2PI - np.log(scale)


def log_prior(weights: Sequence[np.ndarray], kind: int, scales: Sequence, indicators=None,
              freq_indicator: float = 0.0, prior_ind1: float = 0.5) -> float:
    """npBNN.calc_prior (BNN_env.py:180-194): sums over ALL entries (masked zeros and bias
    columns included); uniform prior (kind 0) contributes 0."""
    lp = 0.0
    if kind != 0:
        for w, s in zip(weights, scales):
            lp += np.sum(logpdf_prior(w, kind, s))
    if freq_indicator:
        k = np.sum(indicators)
        lp += k * np.log(prior_ind1) + (indicators.size - k) * np.log(1.0 - prior_ind1)
    return float(lp)


def feature_transform(x: np.ndarray, feature_indicators, feature_means) -> np.ndarray:
    """data_transform_obj.transform (BNN_env.py:9-17): columns whose indicator is 0 are replaced by the training mean
    of that feature."""
    d = np.array(x, dtype=np.float64, copy=True)
    off = np.asarray(feature_indicators) == 0
    d[:, off] = np.asarray(feature_means, dtype=np.float64)[off]
    return d


def update_binomial(ind: np.ndarray, flips: np.ndarray) -> np.ndarray:
    """UpdateBinomial (BNN_mcmc.py:98-99) with the binomial draw injected: |ind - flips|."""
    return np.abs(np.asarray(ind) - np.asarray(flips))


def gibbs_prior_scales(weights: Sequence[np.ndarray], hyper_p: int, gamma=None) -> list:
    """npBNN.sample_prior_scale (BNN_env.py:196-219): one conjugate draw of the Normal prior's standard deviation per
    layer (hyper_p 1: GibbsSampleNormStdGammaVector, a=2), per input node (2: GibbsSampleNormStdGamma2D, a=1, sums
    over axis 0) or per weight (3: GibbsSampleNormStdGammaONE, a=1.5 + one observation); b = 0.1, mu = 0
    (BNN_mcmc.py:124-141).  tau ~ Gamma(a', scale = 1/b'), sd = 1/sqrt(tau).  `gamma(shape, scale=)` defaults to
    numpy's global generator, which is what the reference draws from; one call per layer."""
    gamma = np.random.gamma if gamma is None else gamma
    out = []
    for w in weights:
        w = np.asarray(w, dtype=np.float64)
        if hyper_p == 1:
            a, b = 2 + w.size / 2.0, 0.1 + np.sum(w.flatten() ** 2) / 2.0
        elif hyper_p == 2:
            a, b = 1 + w.shape[0] / 2.0, 0.1 + np.sum(w ** 2, axis=0) / 2.0
        elif hyper_p == 3:
            a, b = 1.5 + 0.5, 0.1 + (w ** 2) / 2.0
        else:
            raise ValueError("hyper_p must be 1, 2 or 3")
        out.append(1 / np.sqrt(gamma(a, scale=1.0 / b)))
    return out


# --------------------------------------------------------------------------------------
# proposal  (BNN_mcmc.py:57-69)
# --------------------------------------------------------------------------------------
def apply_update_normal(w: np.ndarray, ix: np.ndarray, iy: np.ndarray, dz: np.ndarray,
                        hi: float = np.inf, lo: float = -np.inf) -> np.ndarray:
    """UpdateNormal with the random draws injected: z[ix,iy] = w[ix,iy] + dz via fancy
    assignment (duplicates: the LAST draw wins, increments do not accumulate), then one
    reflection at the bounds (BNN_mcmc.py:64-67)."""
    z = np.array(w, dtype=np.float64, copy=True)
    z[ix, iy] = w[ix, iy] + dz
    over = z > hi
    z[over] = hi - (z[over] - hi)
    under = z < lo
    z[under] = lo + (lo - z[under])
    return z


# --------------------------------------------------------------------------------------
# accuracy counters  (BNN_lib.py:195-233)
# --------------------------------------------------------------------------------------
def class_counters(y: np.ndarray, labels: np.ndarray):
    """n_correct, per-class correct, per-class totals, predicted-label histogram.
    CalcAccuracy = n_correct/N (:203-209); CalcLabelAccuracy = correct_k/total_k over
    np.unique(labels) (:211-219); CalcLabelFreq = hist/N (:228-233). argmax = first maximum."""
    k = y.shape[1]
    pred = np.argmax(y, axis=1)
    ok = pred == labels
    return (int(np.sum(ok)),
            np.bincount(labels[ok], minlength=k).astype(np.int64),
            np.bincount(labels, minlength=k).astype(np.int64),
            np.bincount(pred, minlength=k).astype(np.int64))


def regression_sums(y: np.ndarray, labels: np.ndarray):
    """sum r_j and sum r_j^2 per output with r = y[:, :O] - labels.  MSE (CalcAccuracyRegression,
    BNN_lib.py:195-201) = sum r^2 / (N O); empirical sigma_j (BNN_env.py:476) = population std of r_j."""
    r = y[:, :labels.shape[1]] - labels
    return np.sum(r, axis=0), np.sum(r * r, axis=0)


# --------------------------------------------------------------------------------------
# chain state + one MH iteration  (BNN_env.py:273-532)
# --------------------------------------------------------------------------------------
@dataclass
class Model:
    """The fields of npBNN the hot path reads (BNN_env.py:36-173)."""
    x: np.ndarray
    labels: np.ndarray
    weights: List[np.ndarray]
    act: str = "ReLU"
    alphas: Optional[np.ndarray] = None
    mode: str = "classification"        # classification | regression | regression-error
    prior: int = 1
    prior_scale: Optional[Sequence] = None
    w_bound: float = np.inf
    mask: Optional[List[np.ndarray]] = None
    class_w: Optional[np.ndarray] = None
    inst_w: Optional[np.ndarray] = None
    empirical_error: bool = False
    error_prm: object = 1.0
    x_test: Optional[np.ndarray] = None
    labels_test: Optional[np.ndarray] = None
    act_trainable: bool = False         # ActFun(trainable=True): `alphas` are the accepted parameters (_acc_prm)

    def __post_init__(self):
        if self.prior_scale is None:
            self.prior_scale = np.ones(len(self.weights))
        if self.prior == 0:
            # uniform prior: the bound is p_scale (BNN_env.py:135-137); caller passes w_bound
            pass

    @property
    def out_kind(self):
        return {"classification": "softmax", "regression": "identity",
                "regression-error": "regress-error"}[self.mode]

    @property
    def n_params(self):
        return int(sum(w.size for w in self.weights))


@dataclass
class Sampler:
    """The fields of MCMC the hot path reads/writes (BNN_env.py:282-379)."""
    update_f: np.ndarray
    update_ws: np.ndarray               # one scalar per layer (the reference stores constant matrices)
    update_n: np.ndarray
    max_n: np.ndarray
    temperature: float = 1.0
    lik_temp: float = 1.0
    adapt_f: float = 0.0
    adapt_fM: float = 1.0
    adapt_freq: int = 1000
    adapt_stop: int = 0
    sample_from_prior: int = 0
    it: int = 0
    estimate_error: float = np.inf      # MCMC._estimate_error (BNN_env.py:374-377): error_prm is proposed after it
    freq_layer_update: Optional[np.ndarray] = None
    acc_mem: List[int] = field(default_factory=lambda: [1])
    acceptance_rate: float = 0.0
    last_accepted: int = 1
    logLik: float = 0.0
    logPrior: float = 0.0
    logPost: float = 0.0
    accuracy: float = 0.0
    label_acc: Optional[np.ndarray] = None
    label_freq: Optional[np.ndarray] = None
    test_accuracy: float = 0.0


def make_sampler(m: Model, update_f=None, update_ws=None, temperature=1.0, n_iteration=100000,
                 lik_temp=1.0, adapt_f=0.0, adapt_fM=1.0, adapt_freq=1000, adapt_stop=None,
                 sample_from_prior=0, estimate_error=False, init_additional_prob=0.0) -> Sampler:
    """MCMC.__init__ (BNN_env.py:274-379): update sizes, initial forward / lik / prior / accuracy."""
    nl = len(m.weights)
    if update_ws is None:
        update_ws = [0.075] * nl
    if update_f is None:
        update_f = [0.05] * nl
    update_f = np.array(update_f[:nl], dtype=np.float64)
    sizes = np.array([w.size for w in m.weights])
    update_n = np.array([max(1, int(np.round(sizes[i] * update_f[i]))) for i in range(nl)])
    s = Sampler(update_f=update_f, update_ws=np.array(update_ws[:nl], dtype=np.float64), update_n=update_n,
                max_n=sizes.astype(int), temperature=temperature, lik_temp=lik_temp, adapt_f=adapt_f,
                adapt_fM=adapt_fM, adapt_freq=adapt_freq,
                adapt_stop=int(n_iteration * 0.05) if adapt_stop is None else adapt_stop,
                sample_from_prior=sample_from_prior, freq_layer_update=np.ones(nl),
                # BNN_env.py:374-377: min(20000, 0.1 n_iteration) when estimate_error else n_iteration
                estimate_error=float(min(20000, 0.1 * n_iteration)) if estimate_error else float(n_iteration))
    y = forward(m.x, m.weights, m.act, m.alphas, m.out_kind)
    s.logLik = 0.0 if sample_from_prior else likelihood(m, y, m.error_prm, lik_temp)
    s.logPrior = log_prior(m.weights, m.prior, m.prior_scale) + init_additional_prob      # BNN_env.py:320
    s.logPost = s.logLik + s.logPrior
    _refresh_accuracy(m, s, y, m.weights)
    return s


def likelihood(m: Model, y: np.ndarray, sig, lik_temp: float) -> float:
    if m.mode == "classification":
        return loglik_categorical(y, m.labels, m.class_w, m.inst_w, lik_temp)
    if m.mode == "regression":
        return loglik_regression(y, m.labels, sig, lik_temp)
    return loglik_regression_error(y, m.labels, lik_temp)


def _refresh_accuracy(m: Model, s: Sampler, y: np.ndarray, weights) -> None:
    """What the reference recomputes on every accept (BNN_env.py:508-518)."""
    if m.mode == "classification":
        nc, ck, tk, hist = class_counters(y, m.labels)
        s.accuracy = nc / y.shape[0]
        present = tk > 0
        s.label_acc = ck[present] / tk[present]
        s.label_freq = hist / y.shape[0]
    else:
        sr, sr2 = regression_sums(y, m.labels)
        s.label_acc = sr2 / y.shape[0]
        s.accuracy = float(np.sum(sr2) / (y.shape[0] * m.labels.shape[1]))
        s.label_freq = None
    if m.x_test is not None and len(m.x_test) > 0:
        yt = forward(m.x_test, weights, m.act, m.alphas, m.out_kind)
        if m.mode == "classification":
            s.test_accuracy = class_counters(yt, m.labels_test)[0] / yt.shape[0]
        else:
            _, sr2 = regression_sums(yt, m.labels_test)
            s.test_accuracy = float(np.sum(sr2) / (yt.shape[0] * m.labels_test.shape[1]))
    else:
        s.test_accuracy = 0


def adapt(m: Model, s: Sampler) -> None:
    """Adaptation block of mh_step (BNN_env.py:392-413)."""
    if s.it % s.adapt_freq == 0 and s.it < s.adapt_stop:
        if s.acceptance_rate < s.adapt_f:
            s.freq_layer_update = s.freq_layer_update * 0.8
            s.update_f = np.array(s.update_f) * 0.85
            n = (s.max_n * s.update_f).astype(int)
            n[n < 1] = 1
            s.update_n = n
            s.update_ws = s.update_ws * 0.9
        if s.acceptance_rate > s.adapt_fM and np.sum(s.update_n) < m.n_params:
            s.update_f = np.exp(np.log(np.array(s.update_f)) * 0.85)
            n = (s.max_n * s.update_f).astype(int)
            n[n < 1] = 1
            s.update_n = n
            s.update_ws = s.update_ws * 1.2


@dataclass
class StepInjection:
    """The random draws of one MH iteration, in the reference's consumption order
    (BNN_env.py:446-453,493): rr = rs.random(L); per proposed layer ix, iy, normal(0, d, n);
    log_u = log(rs.random())."""
    rr: np.ndarray
    layers: List[Optional[tuple]]       # per layer: None (not proposed) or (ix, iy, dz)
    log_u: float
    # optional proposals drawn BEFORE rr (BNN_env.py:416-444), in this order:
    alpha_ix: Optional[int] = None      # UpdateNormal1D on ActFun._acc_prm: rs.integers(0, len, 1), rs.normal(0, 0.05, 1)
    alpha_dz: float = 0.0
    sigma_mult: Optional[np.ndarray] = None   # multiplier_proposal_vector: rs.binomial(1, .5, S), rs.random(S) -> m
    add_prob: float = 0.0               # the additional_prob argument of mh_step


def layer_is_proposed(rr: np.ndarray, freq_layer_update: np.ndarray, freq_indicator: float = 0.0) -> np.ndarray:
    """rr[argmin]=0 then layer i proposed iff (rr[i] >= freq_indicator or i > 0) and
    rr[i] < freq_layer_update[i]   (BNN_env.py:446-457)."""
    r = np.array(rr, dtype=np.float64, copy=True)
    r[np.argmin(r)] = 0.0
    ok = np.zeros(len(r), dtype=bool)
    for i in range(len(r)):
        ok[i] = (r[i] >= freq_indicator or i > 0) and r[i] < freq_layer_update[i]
    return ok


def error_prm_is_proposed(m: Model, s: Sampler) -> bool:
    """BNN_env.py:435-444: regression, past _estimate_error, no empirical error."""
    return m.mode == "regression" and s.it > s.estimate_error and not m.empirical_error


def multiplier_from_draws(ff: np.ndarray, u: np.ndarray, d: float = 1.1) -> np.ndarray:
    """multiplier_proposal_vector (BNN_mcmc.py:101-113): m = exp(2 log(d) (u - 0.5)), 1 where the mask is 0."""
    mult = np.exp(2 * np.log(d) * (u - .5))
    mult[ff == 0] = 1.
    return mult


def propose_act_prm(prm, ix: int, dz: float) -> np.ndarray:
    """UpdateNormal1D(prm, d=0.05, n=1, Mb=1, mb=0) (BNN_mcmc.py:46-56) with the draw injected; both reflections
    act on every entry."""
    z = np.zeros(np.shape(prm)) + np.asarray(prm, dtype=np.float64)
    z[ix] = z[ix] + dz
    z[z > 1] = 1 - (z[z > 1] - 1)
    z[z < 0] = 0 + (0 - z[z < 0])
    return z


def draw_injection(m: Model, s: Sampler, rs: np.random.Generator) -> StepInjection:
    """Consume `rs` exactly as mh_step + UpdateNormal do (BNN_env.py:416-453, BNN_mcmc.py:46-69,101-113)."""
    nl = len(m.weights)
    extra = {}
    if m.act_trainable:
        extra["alpha_ix"] = int(rs.integers(0, len(m.alphas), 1)[0])
        extra["alpha_dz"] = float(rs.normal(0, 0.05, 1)[0])
    if error_prm_is_proposed(m, s):
        shape = np.shape(m.error_prm)
        ff = rs.binomial(1, 0.5, shape)
        extra["sigma_mult"] = multiplier_from_draws(ff, rs.random(shape))
    rr = rs.random(nl)
    ok = layer_is_proposed(rr, s.freq_layer_update)
    layers = []
    for i in range(nl):
        if not ok[i]:
            layers.append(None)
            continue
        w = m.weights[i]
        n = int(s.update_n[i])
        ix = rs.integers(0, w.shape[0], n)
        iy = rs.integers(0, w.shape[1], n)
        dz = rs.normal(0, np.full(n, s.update_ws[i]), n)
        layers.append((ix, iy, dz))
    return StepInjection(rr=rr, layers=layers, log_u=float(np.log(rs.random())), **extra)


def mh_step(m: Model, s: Sampler, inj: Optional[StepInjection] = None,
            rs: Optional[np.random.Generator] = None) -> dict:
    """One MH iteration (BNN_env.py:381-532).  Randomness is either injected (`inj`) or drawn
    from `rs` AFTER the adaptation block, as the reference does.  Mutates `m.weights` (rebinds
    on accept), `m.error_prm` and `s`.  Returns the proposal's diagnostics."""
    adapt(m, s)
    if inj is None:
        inj = draw_injection(m, s, rs)
    hastings = 0.0
    additional_prob = inj.add_prob
    alphas = m.alphas
    if m.act_trainable:                                       # BNN_env.py:416-421
        alphas = propose_act_prm(m.alphas, inj.alpha_ix, inj.alpha_dz)
        r = 10
        additional_prob += np.log(r) * -np.sum(alphas) * r    # "aka exponential Exp(r)"
    # error parameter (BNN_env.py:435-444): multiplier proposal past _estimate_error, else the scalar 1;
    # regression + empirical_error replaces it by the residual std after the forward pass (:475-476)
    sig = 1
    if error_prm_is_proposed(m, s):
        sig = m.error_prm * inj.sigma_mult
        hastings += np.sum(np.log(inj.sigma_mult))
        r = 1
        additional_prob += np.log(r) * -np.sum(sig) * r
    w_prime = []
    for i, w in enumerate(m.weights):
        upd = inj.layers[i]
        if upd is None:
            z = w + 0
        else:
            z = apply_update_normal(w, upd[0], upd[1], upd[2], m.w_bound, -m.w_bound)
        if m.mask is not None:
            z = z * m.mask[i]
        w_prime.append(z)
    y = forward(m.x, w_prime, m.act, alphas, m.out_kind)
    if m.mode == "regression" and m.empirical_error:
        sig = np.std(y - m.labels, axis=0)
    lp = log_prior(w_prime, m.prior, m.prior_scale) + additional_prob
    ll = 0.0 if s.sample_from_prior else likelihood(m, y, sig, s.lik_temp)
    post = ll + lp
    accept = bool((post - s.logPost) * s.temperature + hastings >= inj.log_u)
    if accept:
        m.weights = w_prime
        if m.mode == "regression":
            m.error_prm = sig
        if m.act_trainable:
            m.alphas = alphas                                  # reset_accepted_prm
        s.logPost, s.logLik, s.logPrior = post, ll, lp
        _refresh_accuracy(m, s, y, w_prime)
        s.last_accepted = 1
    else:
        s.last_accepted = 0
    s.acc_mem.append(s.last_accepted)
    s.acceptance_rate = float(np.mean(s.acc_mem))
    if len(s.acc_mem) > 100:
        s.acc_mem = s.acc_mem[-100:]
    s.it += 1
    return {"logLik_prime": ll, "logPrior_prime": lp, "accepted": int(accept), "additional_prob": additional_prob,
            "hastings": hastings}


# --------------------------------------------------------------------------------------
# MC3 swap  (BNN_mc3.py:98-112)
# --------------------------------------------------------------------------------------
def mc3_swap(log_post: np.ndarray, temps: np.ndarray, j: int, k: int, log_u: float):
    """r = (lp_k - lp_j) T_j + (lp_j - lp_k) T_k ; swap the two temperatures iff r >= log u."""
    t = np.array(temps, dtype=np.float64, copy=True)
    r = (log_post[k] - log_post[j]) * t[j] + (log_post[j] - log_post[k]) * t[k]
    swapped = bool(r >= log_u)
    if swapped:
        t[j], t[k] = temps[k], temps[j]
    return t, swapped, float(r)


# --------------------------------------------------------------------------------------
# posterior prediction summaries  (BNN_lib.py:352-397, BNN_pdp.py:63-82)
# --------------------------------------------------------------------------------------
def posterior_predict(x: np.ndarray, post_weights: Sequence[Sequence[np.ndarray]], act: str,
                      post_alphas=None, out_kind: str = "softmax", mode: int = 1):
    """Dense [S,N,K] tensor plus its summary: mode 0 = argmax vote share, mode 1 = mean
    (BNN_lib.py:376-392)."""
    dense_out = np.array([forward(x, w, act, None if post_alphas is None else post_alphas[i], out_kind)
                          for i, w in enumerate(post_weights)])
    s, n, k = dense_out.shape
    if mode == 0:
        votes = np.zeros((n, k))
        am = np.argmax(dense_out, axis=2)
        for c in range(k):
            votes[:, c] = np.sum(am == c, axis=0)
        return dense_out, votes / s
    return dense_out, np.mean(dense_out, axis=0)


def resample_categorical(dense: np.ndarray, u: np.ndarray):
    """sample_from_categorical (BNN_lib.py:682-713) with the uniforms injected: dense [S,N,K] class probabilities,
    u [N,S] the np.random.random(S) draw of every instance.  For each (instance, sample) the class is the argmin over
    k of cumsum_k - u with negative entries replaced by 1 (so: the first class whose cumulative probability reaches u;
    class 0 when none does).  Returns (predictions [N,K], class_counts [S,K], post_predictions [N,S])."""
    s, n, k = dense.shape
    q = np.cumsum(dense, axis=2) - np.transpose(u)[:, :, None]
    q = np.where(q < 0, 1.0, q)
    drawn = np.argmin(q, axis=2)                               # [S, N]
    onehot = (drawn[:, :, None] == np.arange(k)[None, None, :])
    predictions = onehot.sum(axis=0) / float(s)                # share of the S draws per class
    class_counts = onehot.sum(axis=1).astype(np.float64)       # instances per class and sample
    return predictions, class_counts, drawn.T.astype(np.float64)


def pdp_step(x: np.ndarray, focal: Sequence[int], values: np.ndarray, post_weights, act: str,
             post_alphas=None, out_kind: str = "softmax", classification: bool = True):
    """One grid step of get_pdp (BNN_pdp.py:63-82): overwrite focal columns, predict with all
    samples, class-cumsum, mean over (S,N), 2.5/97.5 % row quantiles of the S-mean."""
    feat = np.copy(x)
    feat[:, focal] = values
    pred, _ = posterior_predict(feat, post_weights, act, post_alphas, out_kind, 1)
    if classification:
        pred = np.cumsum(pred, axis=2)
    mean = np.mean(pred, axis=(0, 1))
    q = np.quantile(np.mean(pred, axis=0), q=(0.025, 0.975), axis=0)
    return mean, q[0], q[1]


# --------------------------------------------------------------------------------------
# masks  (BNN_lib.py:16-47)
# --------------------------------------------------------------------------------------
def block_mask(shape, col_groups: Sequence[int], rows_per_group: Sequence[int]) -> np.ndarray:
    """One layer of create_mask: column i belongs to group col_groups[i]; consecutive columns
    with the same id share one block of rows_per_group[g] rows; blocks stack downwards."""
    if len(col_groups) == 0:
        return np.ones(shape)
    msk = np.zeros(shape)
    g, row0 = 0, 0
    for c in range(len(col_groups)):
        if c > 0 and col_groups[c] != col_groups[c - 1]:
            row0 += rows_per_group[g]
            g += 1
        msk[row0:row0 + rows_per_group[g], c] = 1
    return msk
