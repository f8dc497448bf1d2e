"""GPU: every instantiation family of the shape-specialised forward kernel k_fwd3 (two padded width families x four
activations x {categorical, Gaussian, Gaussian with sigma head}) against the oracle -- scoring, test-row counters,
weights, prediction summaries, and MH chains whose proposals are packed into the padded-up layout."""
import numpy as np
import pytest

from oracle import npbnn_oracle as orc
from tests.test_gpu_parity import LIK, make_engine, oracle_score, rel_close

pytestmark = pytest.mark.gpu

ACTS = ["ReLU", "genReLU", "swish", "tanh"]


def _net(rng, f, hidden, out, bias):
    """Weight shapes as init_weight_prm builds them (BNN_mcmc.py:9-25)."""
    b1, b2, b3 = int(bias >= 1), int(bias >= 2), int(bias in (3, -1))
    return [(hidden[0], f + b1), (hidden[1], hidden[0] + b2), (out, hidden[1] + b3)]


def _problem(mode, n, f, hidden, k, act, bias, seed, n_sets=4, n_test=0):
    rng = np.random.default_rng(seed)
    out = {"classification": k, "regression": k, "regression-error": 2 * k}[mode]
    shapes = _net(rng, f, hidden, out, bias)
    x = rng.standard_normal((n + n_test, f))
    alphas = np.array([0.05, 0.3]) if act == "genReLU" else None
    teacher = [rng.normal(0, 0.5, s) for s in shapes]
    okind = {"classification": "softmax", "regression": "identity", "regression-error": "regress-error"}[mode]
    yt = orc.forward(x, teacher, act, alphas, okind)
    if mode == "classification":
        lab = np.argmax(yt, axis=1)
        lab[:k] = np.arange(k)
        if n_test:
            lab[n:n + k] = np.arange(k)
    else:
        lab = yt[:, :k] + 0.2 * rng.standard_normal((n + n_test, k))
    sets = [[rng.normal(0, 0.3, s) for s in shapes] for _ in range(n_sets)]
    m = orc.Model(x=x[:n], labels=lab[:n], weights=sets[0], act=act, alphas=alphas, mode=mode,
                  x_test=x[n:] if n_test else None, labels_test=lab[n:] if n_test else None,
                  error_prm=np.ones(k) if mode == "regression" else 1.0)
    return m, sets, alphas


def _check_scores(eng, m, sets, alphas, kernel_prefix, sigma_mode=0):
    nl = 3
    al = None if alphas is None else np.tile(np.resize(alphas, nl), (len(sets), 1))
    res = eng.forward_lik(sets, alphas=al, sigma_mode=sigma_mode)
    assert eng.last_kernel.startswith(kernel_prefix), eng.last_kernel
    for i, w in enumerate(sets):
        ref = oracle_score(m, w, "empirical" if sigma_mode else None)
        assert rel_close(res["loglik"][i], ref["loglik"]), (i, res["loglik"][i], ref["loglik"])
        if m.mode == "classification":
            K = eng.K
            c = res["counts"][i]
            assert c[0] == ref["n_correct"]
            assert np.array_equal(c[2:2 + K], ref["class_correct"]) and np.array_equal(c[2 + K:2 + 2 * K], ref["pred_hist"])
            if m.x_test is not None:
                assert c[1] == ref["n_correct_test"]
        else:
            assert np.allclose(res["sums"][i][0], ref["sum_r"], rtol=1e-9, atol=1e-9)
            assert rel_close(res["sums"][i][1], ref["sum_r2"])
            if m.x_test is not None:
                assert rel_close(res["sums"][i][2], ref["sum_r2_test"])
    return res


@pytest.mark.parametrize("act", ACTS)
@pytest.mark.parametrize("mode", ["classification", "regression", "regression-error"])
@pytest.mark.parametrize("family", ["A", "B"])
def test_exact_width_families(family, mode, act):
    """Networks whose padded widths ARE a family's (64 -> 64 -> 32 / 32 -> 32 -> 16): k_fwd3 at any row count.
    4,099 + 200 test rows (ragged last tile, train / test boundary inside a tile)."""
    f, hidden = (64, (64, 32)) if family == "A" else (32, (32, 16))
    k = 10 if mode == "classification" else 2
    m, sets, alphas = _problem(mode, 4099, f, hidden, k, act, bias=-1 if family == "A" else 2, seed=11, n_test=200)
    eng = make_engine(m)
    prefix = "k_fwd3<%s,%d,%d,%d" % ({"ReLU": "relu", "genReLU": "leaky", "swish": "swish", "tanh": "tanh"}[act], f, hidden[0], hidden[1])
    res = _check_scores(eng, m, sets, alphas, prefix, sigma_mode=1 if mode == "regression" else 0)
    # the generic kernel agrees (same arithmetic, other summation order)
    eng.set_option("force_generic", 1)
    al = None if alphas is None else np.tile(np.resize(alphas, 3), (len(sets), 1))
    gen = eng.forward_lik(sets, alphas=al, sigma_mode=1 if mode == "regression" else 0)
    assert eng.last_kernel == "k_fwd_generic"
    assert np.allclose(gen["loglik"], res["loglik"], rtol=1e-10)
    eng.set_option("force_generic", 0)
    # prediction summaries through the same family
    out = eng.predict(m.x, sets, alphas=al, mean=True, votes=(mode == "classification"), dense=True)
    assert eng.last_kernel.startswith(prefix), eng.last_kernel
    dense_ref, mean_ref = orc.posterior_predict(m.x, sets, act, None if alphas is None else [alphas] * len(sets), m.out_kind, 1)
    # regression outputs cross zero (cancellation in the last layer): absolute tolerance at the scale of the outputs
    atol = 1e-300 if mode == "classification" else 1e-11
    assert np.allclose(out["dense"], dense_ref, rtol=1e-10, atol=atol)
    assert np.allclose(out["mean"], mean_ref, rtol=1e-10, atol=atol)
    if mode == "classification":
        _, votes_ref = orc.posterior_predict(m.x, sets, act, None if alphas is None else [alphas] * len(sets), "softmax", 0)
        assert np.array_equal(out["votes"], votes_ref)
    eng.close()


@pytest.mark.parametrize("case", [("classification", 40, (50, 20), 7, "tanh", 2, "k_fwd3<tanh,64,64,32"),
                                  ("classification", 20, (24, 12), 5, "ReLU", -1, "k_fwd3<relu,32,32,16"),
                                  ("regression", 20, (24, 12), 3, "swish", 2, "k_fwd3<swish,32,32,16"),
                                  ("regression-error", 48, (40, 24), 2, "genReLU", 3, "k_fwd3<leaky,64,64,32")])
def test_smaller_networks_are_padded_up_on_large_data(case):
    """A three-layer network that fits inside a family is padded up to it once the data set has at least one full
    round of warp tiles (28,416 rows); below that it keeps the minimal padding and the generic kernel.  Scores, MH chain
    (proposals packed into the padded layout by k_mh_update) and prediction against the oracle."""
    mode, f, hidden, k, act, bias, prefix = case
    n = 30_011
    m, sets, alphas = _problem(mode, n, f, hidden, k, act, bias, seed=5, n_sets=3, n_test=333)
    eng = make_engine(m)
    _check_scores(eng, m, sets, alphas, prefix)
    # the generic kernel on the padded-up layout (reached through force_generic and by injected-uniform resampling)
    eng.set_option("force_generic", 1)
    _check_scores(eng, m, sets, alphas, "k_fwd_generic")
    eng.set_option("force_generic", 0)
    # a chain on the padded layout: 6 MH iterations replayed against the oracle
    s = orc.make_sampler(m, n_iteration=1000, update_f=[0.1] * 3)
    eng.chains_init([sets[0]], update_f=[0.1] * 3, alphas=None if alphas is None else np.resize(alphas, 3),
                    adapt_stop=50)
    rs = np.random.default_rng(2)
    st = eng.read_state()
    assert rel_close(st.logLik[0], s.logLik) and rel_close(st.logPrior[0], s.logPrior)
    for it in range(6):
        inj = orc.draw_injection(m, s, rs)
        cap = max(1, sum(len(l[0]) for l in inj.layers if l is not None))
        arr = {"proposed": np.zeros((1, 1, 3), np.int32), "count": np.zeros((1, 1, 3), np.int32),
               "ix": np.zeros((1, 1, cap), np.int32), "iy": np.zeros((1, 1, cap), np.int32),
               "dz": np.zeros((1, 1, cap)), "log_u": np.full((1, 1), inj.log_u)}
        o = 0
        for l, lay in enumerate(inj.layers):
            if lay is None:
                continue
            cnt = len(lay[0])
            arr["proposed"][0, 0, l], arr["count"][0, 0, l] = 1, cnt
            arr["ix"][0, 0, o:o + cnt], arr["iy"][0, 0, o:o + cnt], arr["dz"][0, 0, o:o + cnt] = lay
            o += cnt
        d = orc.mh_step(m, s, inj)
        eng.mh_steps(1, arr)
        assert eng.last_kernel.startswith(prefix), eng.last_kernel
        st = eng.read_state()
        assert st.last_accepted[0] == d["accepted"], it
        assert rel_close(st.logLik_prop[0], d["logLik_prime"]) and rel_close(st.logPrior_prop[0], d["logPrior_prime"])
    for a, b in zip(st.weights(0), m.weights):
        assert np.array_equal(a, b)
    # the same network on a small data set keeps the minimal padding
    m2, sets2, alphas2 = _problem(mode, 999, f, hidden, k, act, bias, seed=6, n_sets=2)
    eng2 = make_engine(m2)
    _check_scores(eng2, m2, sets2, alphas2, "k_fwd_generic")
    # prediction picks its padding from its own row count
    al = None if alphas is None else np.tile(np.resize(alphas, 3), (len(sets), 1))
    big = eng2.predict(m.x, sets, alphas=al, mean=True)
    assert eng2.last_kernel.startswith(prefix), eng2.last_kernel
    small = eng2.predict(m.x[:999], sets, alphas=al, mean=True)
    assert eng2.last_kernel == "k_fwd_generic"
    atol = 1e-300 if mode == "classification" else 1e-11
    assert np.allclose(big["mean"][:999], small["mean"], rtol=1e-12, atol=atol)
    _, mean_ref = orc.posterior_predict(m.x[:2000], sets, act, None if alphas is None else [alphas] * len(sets), m.out_kind, 1)
    assert np.allclose(big["mean"][:2000], mean_ref, rtol=1e-10, atol=atol)
    eng.close()
    eng2.close()
