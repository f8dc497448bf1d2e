"""GPU: the reference-facing Python interface (npbnn_b200.api) against the reference's recorded chains."""
import os
import pickle

import numpy as np
import pytest

from oracle import npbnn_oracle as orc
from tests import _golden as G

pytestmark = pytest.mark.gpu


def _dat(z):
    return {"data": np.array(z["x"]), "labels": np.array(z["labels"]), "test_data": np.array(z["x_test"]) if len(z["x_test"]) else [],
            "test_labels": np.array(z["labels_test"]) if len(z["x_test"]) else []}


def _close(a, b, rtol=1e-9):
    return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=1e-12)


def _follow(z, meta, bnn, mcmc, n_steps, bn):
    assert _close(mcmc._logLik, z["init_logLik"]) and _close(mcmc._logPrior, z["init_logPrior"])
    assert _close(mcmc._accuracy, z["init_accuracy"]) and _close(mcmc._test_accuracy, z["init_test_accuracy"])
    assert np.array_equal(mcmc._update_n, z["init_update_n"])
    for t in range(n_steps):
        mcmc.mh_step(bnn)
        assert mcmc._last_accepted == int(z["steps_accepted"][t]), t
        assert _close(mcmc._logLik, z["steps_logLik"][t]) and _close(mcmc._logPrior, z["steps_logPrior"][t]), t
        assert _close(mcmc._logPost, z["steps_logPost"][t])
        assert _close(mcmc._accuracy, z["steps_accuracy"][t]) and _close(mcmc._test_accuracy, z["steps_test_accuracy"][t])
        assert _close(mcmc._label_acc, z["steps_label_acc"][t])
        assert abs(mcmc._acceptance_rate - float(z["steps_acceptance_rate"][t])) < 1e-15
        assert np.array_equal(mcmc._update_n, z["steps_update_n"][t]), t
        assert mcmc._current_iteration == t + 1
    return mcmc


def test_classify_script_flow_reproduces_reference_chain():
    """bnn_classify.py:45-70 through the mirrored API with the reference's own generator sequence: the chain
    (accept decisions, log-likelihoods, accuracies, adaptation) is the reference's chain."""
    import npbnn_b200 as bn
    z, meta = G.load("c1_classify")
    np.random.seed(1234)
    bnn = bn.npBNN(_dat(z), n_nodes=[5, 5], use_class_weights=0, actFun=bn.ActFun(fun="tanh"), use_bias_node=2,
                   prior_f=1, p_scale=1, seed=1234, init_std=0.1, instance_weights=None)
    for i in range(3):
        assert np.array_equal(bnn._w_layers[i], z["w0_%d" % i])        # same initial weights as the reference
    assert _close(bnn.calc_prior(), z["init_logPrior"])
    mcmc = bn.MCMC(bnn, update_f=[0.05, 0.05, 0.07], update_ws=[0.075, 0.075, 0.075], n_iteration=10000, sampling_f=10,
                   print_f=1000, n_post_samples=100, sample_from_prior=0, adapt_f=0.3, adapt_fM=0.6)
    _follow(z, meta, bnn, mcmc, 120, bn)
    for i in range(3):
        assert np.array_equal(bnn._w_layers[i], z["wN_%d" % i])
    y = mcmc.y(bnn)
    assert np.allclose(y, z["yN"], rtol=1e-10, atol=1e-300)


@pytest.mark.parametrize("emp", [1, 0])
def test_regress_script_flow_reproduces_reference_chain(emp):
    import npbnn_b200 as bn
    z, meta = G.load("c2_regress_emp%d" % emp)
    np.random.seed(1234)
    bnn = bn.npBNN(_dat(z), n_nodes=[10, 5], estimation_mode="regression", actFun=bn.ActFun(fun="ReLU"), p_scale=1,
                   use_bias_node=2, empirical_error=bool(emp))
    mcmc = bn.MCMC(bnn, update_ws=[0.025, 0.025, 0.05], update_f=[0.005, 0.005, 0.05], n_iteration=20000, sampling_f=100,
                   print_f=1000, n_post_samples=100, likelihood_tempering=1, adapt_f=0.3, estimate_error=False)
    _follow(z, meta, bnn, mcmc, 100, bn)
    assert _close(bnn._error_prm, z["steps_error_prm"][99])


def test_batched_run_with_adaptation_matches_single_steps():
    """MCMC.run(n) batches host-drawn proposals between adaptation points; the chain must not change."""
    import npbnn_b200 as bn
    z, meta = G.load("syn_adapt")
    np.random.seed(7)
    dat = _dat(z)
    bnn = bn.npBNN(dat, n_nodes=[4, 3], actFun=bn.ActFun(fun="tanh"), use_bias_node=2, seed=7)
    for i in range(3):
        assert np.array_equal(bnn._w_layers[i], z["w0_%d" % i])
    mcmc = bn.MCMC(bnn, update_f=[0.3, 0.3, 0.3], adapt_f=0.3, adapt_fM=0.6, adapt_freq=10, n_iteration=1000)
    mcmc.run(bnn, 80)
    assert mcmc._current_iteration == 80
    assert _close(mcmc._logLik, z["steps_logLik"][79]) and _close(mcmc._logPrior, z["steps_logPrior"][79])
    assert np.array_equal(mcmc._update_n, z["steps_update_n"][79])
    for i in range(3):
        assert np.array_equal(bnn._w_layers[i], z["wN_%d" % i])


def test_block_mask_network_flow():
    import npbnn_b200 as bn
    z, meta = G.load("syn_block_mask")
    np.random.seed(7)
    bnn = bn.npBNN(_dat(z), n_nodes=[24, 16], actFun=bn.ActFun(fun="tanh"), use_bias_node=-1, seed=7)
    m = bn.create_mask(bnn._w_layers, indx_input_list=[list(range(8)), sum(([g] * 3 for g in range(8)), []), []],
                       nodes_per_feature_list=[[3] * 8, [2] * 8, []])
    for i in range(3):
        assert np.array_equal(m[i], z["mask_%d" % i])
    bnn.apply_mask(m)
    mcmc = bn.MCMC(bnn, n_iteration=1000)
    _follow(z, meta, bnn, mcmc, 60, bn)


def test_run_mcmc_logger_pickle_and_prediction(tmp_path):
    """run_mcmc (BNN_mcmc.py:153-170) + postLogger files + the prediction callers on the pickled samples."""
    import npbnn_b200 as bn
    z, meta = G.load("syn_swish_cauchy")
    np.random.seed(7)
    dat = _dat(z)
    bnn = bn.npBNN(dat, n_nodes=[4, 3], actFun=bn.ActFun(fun="swish"), use_bias_node=2, prior_f=2, p_scale=0.7, seed=7)
    mcmc = bn.MCMC(bnn, n_iteration=200, sampling_f=20, print_f=100, n_post_samples=5, rng="philox")
    logger = bn.postLogger(bnn, filename="t", wdir=str(tmp_path))
    bn.run_mcmc(bnn, mcmc, logger)
    assert mcmc._current_iteration == 200
    rows = open(logger._logfile).read().strip().split("\n")
    assert rows[0].split("\t")[:4] == ["it", "posterior", "likelihood", "prior"] and len(rows) == 1 + 10
    b2, m2, l2 = bn.load_obj(logger._pklfile)
    assert len(l2._post_weight_samples) == 5 and m2._current_iteration == 200
    dense, votes = bn.get_posterior_cat_prob(dat["test_data"], l2._post_weight_samples, post_summary_mode=0,
                                             actFun=b2._act_fun, output_act_fun=b2._output_act_fun)
    _, mean = bn.get_posterior_cat_prob(dat["test_data"], l2._post_weight_samples, post_summary_mode=1,
                                        actFun=b2._act_fun, output_act_fun=b2._output_act_fun)
    assert dense.shape == (5, len(dat["test_data"]), 3)
    assert np.allclose(mean, dense.mean(0), rtol=1e-12) and np.allclose(votes.sum(1), 1.0)
    # single-set forward agrees with the dense tensor; restart from the pickle takes the last sample
    y = bn.RunPredict(dat["test_data"], l2._post_weight_samples[-1]["weights"], b2._act_fun, b2._output_act_fun)
    assert np.allclose(y, dense[-1], rtol=1e-12)
    b3 = bn.npBNN(dat, n_nodes=[4, 3], actFun=bn.ActFun(fun="swish"), use_bias_node=2, pickle_file=logger._pklfile)
    assert all(np.array_equal(a, b) for a, b in zip(b3._w_layers, l2._post_weight_samples[-1]["weights"]))
    res = bn.pdp(logger._pklfile, [[1], [0, 2]])
    assert res[0]["pdp"].shape == (100, 3, 3) and res[1]["pdp"].shape == (2, 3, 3)
    assert np.allclose(res[0]["pdp"][:, -1, 0], 1.0)     # cumulative class probability ends at 1


class _ReplaySwap:
    def __init__(self, z, n):
        self.z, self.i, self.n = z, 0, n

    def pair(self, n):
        j, k = [int(v) for v in self.z["it%d_pair" % self.i]]
        return j, k

    def log_uniform(self):
        v = float(self.z["it%d_log_u" % self.i])
        self.i += 1
        return v


def test_mc3_reproduces_reference_run(tmp_path):
    """BNN_mc3.py:87-126 with the reference's per-step reseeding (default_rng(it + id)) and its recorded swap
    draws: log-posteriors, temperatures and weights of every chain after every swap period are the reference's."""
    import npbnn_b200 as bn
    z, meta = G.load("mc3")
    dat = {"data": np.array(z["x"]), "labels": np.array(z["labels"]), "test_data": [], "test_labels": []}
    bnn = bn.npBNN(dat, n_nodes=[4, 3], use_bias_node=-1, seed=1, actFun=bn.ActFun(fun="swish"),
                   init_weights=[np.array(z["w0_%d" % i]) for i in range(3)])
    logger = bn.postLogger(bnn, filename="mc3", wdir=str(tmp_path))
    mc3 = bn.MC3(bnn, logger=logger, n_post_samples=10, sampling_f=5, n_iteration=40, n_chains=3, swap_frequency=5,
                 verbose=0, adapt_freq=50, adapt_f=0.1, adapt_fM=0.6, adapt_stop=1000)
    mc3._swap_rng = _ReplaySwap(z, 3)
    n_it = int(mc3.n_mc3_iteration)
    mc3.n_mc3_iteration = 1
    for it in range(n_it):
        assert np.array_equal(mc3.current_temperatures, z["it%d_temps_before" % it])
        mc3.run_mcmc()
        lp = np.array([a[1]._logPost for a in mc3.singleChainArgs])
        assert _close(lp, z["it%d_logPost" % it]), it
        assert np.array_equal(mc3.current_temperatures, z["it%d_temps_after" % it]), it
        for c in range(3):
            for li in range(3):
                assert np.array_equal(mc3.singleChainArgs[c][0]._w_layers[li], z["it%d_c%d_w%d" % (it, c, li)]), (it, c, li)
    assert os.path.exists(logger._pklfile)
    lg = pickle.load(open(logger._pklfile, "rb"))[2]
    assert lg._post_weight_samples
    # the pickle on disk when run_mcmc returns is the LAST snapshot (background writer): the cold chain's final state
    cold = [a for a in mc3.singleChainArgs if a[1]._temperature == 1][0]
    assert lg._post_weight_samples[-1]["mcmc_it"] == cold[1]._current_iteration
    for wa, wb in zip(lg._post_weight_samples[-1]["weights"], cold[0]._w_layers):
        assert np.array_equal(wa, wb)


def test_mc3_background_pickle_matches_synchronous_writes(tmp_path):
    """MC3.run_mcmc hands the per-period [bnn, mcmc, logger] pickle (BNN_env.py:658) to the logger's background
    writer; with async_pickle = False every period writes synchronously as the reference does.  Same .log rows, same
    posterior samples in the final pickle."""
    import npbnn_b200 as bn
    z, meta = G.load("mc3")
    dat = {"data": np.array(z["x"]), "labels": np.array(z["labels"]), "test_data": [], "test_labels": []}
    runs = []
    for tag, flag in (("bg", True), ("sync", False)):
        np.random.seed(3)                 # MC3 draws its chain seeds from numpy's global generator (BNN_mc3.py:40)
        bnn = bn.npBNN(dat, n_nodes=[4, 3], use_bias_node=-1, seed=1, actFun=bn.ActFun(fun="swish"),
                       init_weights=[np.array(z["w0_%d" % i]) for i in range(3)])
        logger = bn.postLogger(bnn, filename="mc3_" + tag, wdir=str(tmp_path))
        mc3 = bn.MC3(bnn, logger=logger, n_post_samples=6, sampling_f=5, n_iteration=60, n_chains=3, swap_frequency=5,
                     verbose=0, rng="philox", swap_seed=11)
        mc3.async_pickle = flag
        mc3.run_mcmc()
        assert logger.__dict__.get("_async") is None
        saved = pickle.load(open(logger._pklfile, "rb"))
        runs.append((open(logger._logfile).read().splitlines(), saved[2]._post_weight_samples, saved[0]._w_layers))
    a, b = runs
    assert a[0] == b[0] and len(a[0]) == 1 + 12
    assert len(a[1]) == len(b[1]) == 6
    for sa, sb in zip(a[1], b[1]):
        assert sa["mcmc_it"] == sb["mcmc_it"]
        for wa, wb in zip(sa["weights"], sb["weights"]):
            assert np.array_equal(wa, wb)
    for wa, wb in zip(a[2], b[2]):
        assert np.array_equal(wa, wb)


def test_unsupported_options_raise():
    import npbnn_b200 as bn
    z, meta = G.load("syn_swish_cauchy")
    dat = _dat(z)
    with pytest.raises(NotImplementedError):
        bn.MCMC(bn.npBNN(dat, n_nodes=[4, 3]), update_function=bn.UpdateUniform)        # only UpdateNormal is on the device
    with pytest.raises(IndexError):                       # the reference reads update_f[3] in the indicator branch
        bn.MCMC(bn.npBNN(dat, n_nodes=[4, 3], freq_indicator=0.1), rng="philox")
    with pytest.raises(NotImplementedError):
        bn.npBNN(dat, n_nodes=[4, 3], estimation_mode="custom", size_output=3)
    bnn = bn.npBNN(dat, n_nodes=[4, 3])
    with pytest.raises(NotImplementedError):
        bn.MCMC(bnn, likelihood_f=lambda *a, **k: 0.0)
    with pytest.raises(NotImplementedError):
        bn.MCMC(bnn, update_function=lambda *a, **k: None)


def test_categorical_resampling_matches_reference(tmp_path):
    """get_posterior_cat_prob mode 2 / sample_from_categorical (BNN_lib.py:682-713) fused into the prediction pass:
    the reference's recorded uniforms reproduce its draws exactly; the API consumes the global numpy stream like it."""
    import npbnn_b200 as bn
    from npbnn_b200 import _lib as L
    from npbnn_b200.engine import Engine, NetShape
    z, meta = G.load("sample_cat")
    s = int(meta["S"])
    x = np.array(z["x"])
    post = [{"weights": [np.array(z["s%d_w%d" % (j, li)]) for li in range(3)], "alphas": [0.0]} for j in range(s)]
    eng = Engine(NetShape.from_weights(post[0]["weights"], x.shape[1], act="swish", lik=L.LIK_CATEGORICAL))
    res = eng.predict_sample(x, [p["weights"] for p in post], np.array(z["u"]))
    eng.close()
    assert np.array_equal(res["predictions"], z["predictions"])
    assert np.array_equal(res["class_counts"], z["class_counts"])
    assert np.array_equal(res["post_predictions"], z["post_predictions"])
    af = bn.ActFun(fun="swish")
    np.random.seed(int(meta["seed_mode2"]))
    dense, summ = bn.get_posterior_cat_prob(x, post, post_summary_mode=2, actFun=af, output_act_fun=bn.SoftMax)
    assert np.array_equal(summ, z["mode2"]) and np.allclose(dense, z["dense"], rtol=1e-12)
    np.random.seed(int(meta["seed_direct"]))
    r2 = bn.sample_from_categorical(x, post, actFun=af, output_act_fun=bn.SoftMax)
    assert np.array_equal(r2["predictions"], z["predictions"]) and np.array_equal(r2["post_predictions"], z["post_predictions"])


def test_feature_importance_flow(tmp_path):
    """Permutation feature importance (BNN_lib.py:503-598): a feature the labels depend on loses accuracy when
    shuffled, an irrelevant one does not; the shuffles follow np.random.permutation as in the reference."""
    import npbnn_b200 as bn
    rng = np.random.default_rng(3)
    n, f, k = 400, 4, 3
    x = rng.standard_normal((n, f))
    w = [rng.normal(0, 1.0, (6, f)), rng.normal(0, 1.0, (5, 6)), rng.normal(0, 1.0, (k, 6))]
    w[0][:, 3] = 0.0                                               # feature 3 is irrelevant to the network
    y = np.argmax(orc.forward(x, w, "tanh", None, "softmax"), axis=1)
    post = [{"weights": [a + rng.normal(0, 0.01, a.shape) * (a != 0) for a in w], "alphas": [0.0]} for _ in range(4)]
    np.random.seed(11)
    df = bn.feature_importance(x, weights_posterior=post, true_labels=y, n_permutations=3, write_to_file=True,
                               predictions_outdir=str(tmp_path), actFun=bn.ActFun(fun="tanh"), output_act_fun=bn.SoftMax)
    assert list(df.columns) == ["feature_block_index", "feature_name", "delta_acc_mean", "delta_acc_std",
                                "acc_with_feature_randomized_mean", "acc_with_feature_randomized_std"]
    by = {int(r.feature_block_index): r for r in df.itertuples()}
    assert abs(by[3].delta_acc_mean) < 1e-12 and max(by[i].delta_acc_mean for i in range(3)) > 0.05
    # same numbers as the oracle with the same permutations
    np.random.seed(11)
    ref_acc = np.mean(np.argmax(orc.posterior_predict(x, [p["weights"] for p in post], "tanh", None, "softmax", 0)[1], 1) == y)
    xs = x.copy()
    xs[:, 0] = np.random.permutation(xs[:, 0])
    acc0 = np.mean(np.argmax(orc.posterior_predict(xs, [p["weights"] for p in post], "tanh", None, "softmax", 0)[1], 1) == y)
    assert (tmp_path / "feature_importance.txt").exists()
    assert 0.0 <= acc0 <= ref_acc <= 1.0


@pytest.mark.parametrize("name", ["syn_trainable_genrelu", "syn_trainable_tanh"])
def test_trainable_activation_flow_reproduces_reference_chain(name):
    """ActFun(trainable=True) + init_additional_prob through the reference-shaped API (BNN_env.py:416-421,502-503):
    the chain, the accepted activation parameters and the log-prior (with the Exp(10) term) follow the reference."""
    import npbnn_b200 as bn
    z, meta = G.load(name)
    np.random.seed(7)
    dat = _dat(z)
    af = bn.ActFun(fun=meta["act"], prm=np.array(meta["alphas"]), trainable=True)
    bnn = bn.npBNN(dat, n_nodes=meta["n_nodes"], actFun=af, use_bias_node=meta["use_bias_node"], prior_f=meta["prior"],
                   p_scale=meta["p_scale"], seed=7)
    for i in range(3):
        assert np.array_equal(bnn._w_layers[i], z["w0_%d" % i])
    mcmc = bn.MCMC(bnn, n_iteration=meta["n_iteration"], init_additional_prob=meta["init_additional_prob"])
    assert abs(mcmc._logPrior - float(z["init_logPrior"])) <= 1e-9 * abs(float(z["init_logPrior"]))
    T = int(z["n_steps"])
    for t in range(T):
        mcmc.mh_step(bnn)
        assert mcmc._last_accepted == int(z["steps_accepted"][t]), t
        assert np.allclose(bnn._act_fun._acc_prm, z["steps_act_prm"][t], rtol=1e-14, atol=0), t
        assert abs(mcmc._logPrior - float(z["steps_logPrior"][t])) <= 1e-9 * abs(float(z["steps_logPrior"][t]))
    for i in range(3):
        assert np.array_equal(bnn._w_layers[i], z["wN_%d" % i])


@pytest.mark.parametrize("hp", [1, 2, 3])
def test_hyper_prior_gibbs_flow_reproduces_reference_chain(hp):
    """npBNN(hyper_p=1|2|3): MH iterations interleaved with MCMC.gibbs_step (BNN_env.py:196-219,534-538).  The prior
    scales are drawn on the host with the reference's call sequence, the log-prior of the current weights and of every
    following proposal is evaluated on the device with one scale per layer / input node / weight.  Scales and weights
    bit-exact, log-prior / log-posterior to 1e-9, identical accept decisions."""
    import npbnn_b200 as bn
    z, meta = G.load("syn_hyper_p%d" % hp)
    np.random.seed(int(meta["seed"]))
    bnn = bn.npBNN(_dat(z), n_nodes=[4, 3], actFun=bn.ActFun(fun="tanh"), use_bias_node=2, prior_f=1, p_scale=1,
                   hyper_p=hp, seed=int(meta["seed"]))
    for i in range(3):
        assert np.array_equal(bnn._w_layers[i], z["w0_%d" % i])
    mcmc = bn.MCMC(bnn, n_iteration=1000, update_f=[0.2, 0.2, 0.2])
    assert _close(mcmc._logLik, z["init_logLik"]) and _close(mcmc._logPrior, z["init_logPrior"])
    for t in range(int(meta["n_steps"])):
        if int(z["steps_gibbs"][t]):
            mcmc.gibbs_step(bnn)
            for i in range(3):
                assert np.array_equal(np.asarray(bnn._prior_scale[i]), z["t%d_scale_%d" % (t, i)]), (t, i)
                assert np.array_equal(bnn._w_layers[i], z["t%d_w_%d" % (t, i)])
            assert _close(bnn.calc_prior(), z["steps_logPrior"][t])
        else:
            mcmc.mh_step(bnn)
            assert mcmc._last_accepted == int(z["steps_accepted"][t]), t
        assert _close(mcmc._logLik, z["steps_logLik"][t]), t
        assert _close(mcmc._logPrior, z["steps_logPrior"][t]), (t, mcmc._logPrior, float(z["steps_logPrior"][t]))
        assert _close(mcmc._logPost, z["steps_logPost"][t]), t
        assert mcmc._current_iteration == int(z["steps_iteration"][t])
    for i in range(3):
        assert np.array_equal(bnn._w_layers[i], z["wN_%d" % i])


def test_reset_update_parameters_reach_the_device():
    """MCMC.reset_update_n / _f / _ws (BNN_env.py:540-547) edit the device-resident sampler state."""
    import npbnn_b200 as bn
    z, meta = G.load("syn_adapt")
    np.random.seed(7)
    bnn = bn.npBNN(_dat(z), n_nodes=[4, 3], actFun=bn.ActFun(fun="tanh"), use_bias_node=2, seed=7)
    mcmc = bn.MCMC(bnn, n_iteration=1000)
    mcmc.reset_update_n([3, 2, 1])
    mcmc.reset_update_f([0.11, 0.12, 0.13])
    mcmc.reset_update_ws([np.ones(w.shape) * v for w, v in zip(bnn._w_layers, (0.01, 0.02, 0.03))])
    st = mcmc._group.eng.read_state(weights=False)
    assert np.array_equal(st.update_n[0], [3, 2, 1])
    assert np.allclose(st.update_f[0], [0.11, 0.12, 0.13]) and np.allclose(st.update_ws[0], [0.01, 0.02, 0.03])
    mcmc.mh_step(bnn)
    assert np.array_equal(mcmc._update_n, [3, 2, 1]) and mcmc._current_iteration == 1
    with pytest.raises(NotImplementedError):
        mcmc.reset_update_ws([np.arange(w.size, dtype=float).reshape(w.shape) for w in bnn._w_layers])


def test_pipelined_logger_matches_synchronous_run(tmp_path):
    """run_mcmc with device-generated proposals: the asynchronous snapshot ring (bnn_chains_snapshot, logging points
    exported while the device keeps stepping) must log exactly what the synchronous loop logs."""
    import pickle
    import npbnn_b200 as bn
    z, meta = G.load("syn_swish_cauchy")
    runs = []
    for tag, depth in (("sync", 0), ("ring", 3)):
        np.random.seed(5)
        bnn = bn.npBNN(_dat(z), n_nodes=[4, 3], actFun=bn.ActFun(fun="swish"), use_bias_node=2, seed=5)
        mcmc = bn.MCMC(bnn, n_iteration=205, sampling_f=10, print_f=50, n_post_samples=8, rng="philox")
        logger = bn.postLogger(bnn, filename="pl_" + tag, wdir=str(tmp_path), log_all_weights=0)
        bn.run_mcmc(bnn, mcmc, logger, pipeline_depth=depth)
        assert mcmc._current_iteration == 205
        with open(logger._pklfile, "rb") as f:
            _, _, lg = pickle.load(f)
        rows = open(logger._logfile).read().splitlines()
        runs.append((mcmc._logLik, mcmc._logPost, [w.copy() for w in bnn._w_layers], lg._post_weight_samples, rows))
    a, b = runs
    assert a[0] == b[0] and a[1] == b[1]
    for wa, wb in zip(a[2], b[2]):
        assert np.array_equal(wa, wb)
    assert len(a[3]) == len(b[3]) == 8
    for sa, sb in zip(a[3], b[3]):
        assert sa["mcmc_it"] == sb["mcmc_it"]
        for wa, wb in zip(sa["weights"], sb["weights"]):
            assert np.array_equal(wa, wb)
    assert a[4] == b[4] and len(a[4]) == 1 + 20


def test_philox_chains_agree_statistically_with_reference_chains():
    """BASELINE.json north_star: "full chains must agree statistically with the reference (posterior mean accuracy and
    predictive probabilities within MC error)".  rng="host" IS the reference's chain (bit-for-bit on its generator
    sequence, see the script-flow tests above); rng="philox" draws the same proposal distribution on the device from a
    different generator.  4 chains x 3000 iterations each way (burn-in 1000, every 10th state kept): posterior means of
    log-likelihood, train / test accuracy and the posterior-predictive class probabilities of the test rows must agree
    within Monte-Carlo error (between-chain standard error; thresholds have a > 2x margin over the measured gaps)."""
    import npbnn_b200 as bn
    from tests.golden_data import synth_class

    dat = synth_class(600, 6, 3, seed=21, n_test=200)
    out = {}
    for mode in ("host", "philox"):
        per_chain, sets = [], []
        for ch in range(4):
            np.random.seed(100 + ch)                    # same initial weights in both modes
            bnn = bn.npBNN(dat, n_nodes=[4, 3], actFun=bn.ActFun(fun="tanh"), use_bias_node=2, seed=100 + ch)
            mcmc = bn.MCMC(bnn, n_iteration=3000, update_f=[0.2, 0.2, 0.2], rng=mode, mcmc_id=ch)
            mcmc._rs = np.random.default_rng(500 + ch)  # host mode: an independent stream per chain
            rows = []
            for it in range(300):
                mcmc.run(bnn, 10)
                if it >= 100:
                    rows.append([mcmc._logLik, mcmc._accuracy, mcmc._test_accuracy, mcmc._acceptance_rate])
                    sets.append([w.copy() for w in bnn._w_layers])
            per_chain.append(np.mean(rows, axis=0))
        per_chain = np.array(per_chain)
        prob = mcmc._group.eng.predict(dat["test_data"], sets, mean=True)["mean"]
        out[mode] = (per_chain.mean(0), per_chain.std(0, ddof=1) / 2.0, prob)      # mean, standard error over 4 chains
    (mh, sh, ph), (mp, sp, pp) = out["host"], out["philox"]
    se = np.sqrt(sh ** 2 + sp ** 2)
    gap = np.abs(mh - mp)
    print("host", mh, "philox", mp, "gap", gap, "se", se, "prob gap mean/max", np.abs(ph - pp).mean(), np.abs(ph - pp).max())
    assert gap[0] < max(4 * se[0], 0.01 * abs(mh[0])), (mh[0], mp[0], se[0])        # log-likelihood
    assert gap[1] < max(4 * se[1], 0.02) and gap[2] < max(4 * se[2], 0.03)          # train / test accuracy
    # window acceptance rate: chains of either kind range over 0.24-0.40 depending on where they sit (measured: overall
    # acceptance 0.329 +- 0.018 host vs 0.316 +- 0.027 Philox over 6 chains each)
    assert gap[3] < 0.15
    # predictive probabilities (measured: mean |gap| 0.021, max 0.134 between two sets of 4 chains)
    assert np.abs(ph - pp).mean() < 0.04 and np.abs(ph - pp).max() < 0.25
    assert np.allclose(ph.sum(1), 1.0) and np.allclose(pp.sum(1), 1.0)


@pytest.mark.parametrize("kind", ["weight", "feature", "both"])
def test_indicator_flow_reproduces_reference_chain(kind, tmp_path):
    """npBNN(freq_indicator=0.3) / npBNN(feature_indicators=True) through mh_step (BNN_env.py:423-431,449-466): the flips
    are drawn on the host with the reference's generators, the device applies them, multiplies the first layer by the
    indicators, folds masked features into the first-layer bias and adds the Bernoulli prior term.  Same accept
    decisions, indicators and weights as the reference's recorded chain; likelihood / prior / accuracies to 1e-9."""
    import npbnn_b200 as bn
    z, meta = G.load("syn_ind_%s" % kind)
    seed, nl = int(meta["seed"]), len(meta["n_nodes"]) + 1
    np.random.seed(seed)
    bnn = bn.npBNN(_dat(z), n_nodes=meta["n_nodes"], actFun=bn.ActFun(fun="tanh"), use_bias_node=2, prior_f=1, p_scale=1,
                   freq_indicator=meta["freq_indicator"], prior_ind1=meta["prior_ind1"], seed=seed,
                   feature_indicators=True if kind in ("feature", "both") else None)
    for i in range(nl):
        assert np.array_equal(bnn._w_layers[i], z["w0_%d" % i])
    mcmc = bn.MCMC(bnn, n_iteration=1000, update_f=meta["update_f"], adapt_stop=meta["adapt_stop"])
    assert _close(mcmc._logLik, z["init_logLik"]) and _close(mcmc._logPrior, z["init_logPrior"])
    assert _close(bnn.calc_prior(), z["init_logPrior"])
    for t in range(int(meta["n_steps"])):
        mcmc.mh_step(bnn)
        assert mcmc._last_accepted == int(z["steps_accepted"][t]), t
        assert _close(mcmc._logLik, z["steps_logLik"][t]), (t, mcmc._logLik, float(z["steps_logLik"][t]))
        assert _close(mcmc._logPrior, z["steps_logPrior"][t]), (t, mcmc._logPrior, float(z["steps_logPrior"][t]))
        assert _close(mcmc._logPost, z["steps_logPost"][t])
        assert _close(mcmc._accuracy, z["steps_accuracy"][t]) and _close(mcmc._test_accuracy, z["steps_test_accuracy"][t]), t
        assert abs(np.mean(bnn._indicators) - float(z["steps_mean_ind"][t])) < 1e-15, t
        if kind in ("feature", "both"):
            assert np.array_equal(bnn._feature_indicators, z["steps_feature_ind"][t]), t
    for i in range(nl):
        assert np.array_equal(bnn._w_layers[i], z["wN_%d" % i])
    assert np.array_equal(bnn._indicators, z["indN"])
    assert np.allclose(mcmc.y(bnn), z["yN"], rtol=1e-10, atol=1e-300)
    # the logger writes the reference's extra columns and the indicated first layer
    logger = bn.postLogger(bnn, filename="ind", wdir=str(tmp_path))
    logger.log_sample(bnn, mcmc)
    logger.log_weights(bnn, mcmc)
    head, row = [l.split("\t") for l in open(logger._logfile).read().splitlines()]
    assert len(head) == len(row)
    assert ("mean_ind" in head) == bool(meta["freq_indicator"]) and ("feature_ind_0" in head) == (kind != "weight")
    assert np.array_equal(logger._post_weight_samples[-1]["weights"][0], bnn._w_layers[0] * bnn._indicators)


def test_predict_and_pdp_with_feature_transform():
    """RunPredict / RunPredictInd / get_pdp with data_transform (masked features replaced by their training means,
    BNN_env.py:9-17): the transform is a column override while X is packed; where the PDP's focal feature is itself
    masked the transform wins, as in the reference (BNN_lib.py:248-249)."""
    import npbnn_b200 as bn
    z, meta = G.load("predict_transform")
    x, fi = z["x"], z["fi"]
    dt = bn.data_transform_obj(fi, x.mean(axis=0))
    S = int(meta["S"])
    post = [[z["s%d_w%d" % (j, li)] for li in range(3)] for j in range(S)]
    af = bn.ActFun(fun="tanh")
    y = bn.RunPredict(x, post[0], af, bn.SoftMax, data_transform=dt)
    assert np.allclose(y, z["y_transform"], rtol=1e-10, atol=1e-300)
    y = bn.RunPredictInd(x, post[0], z["ind"], af, bn.SoftMax, data_transform=dt)
    assert np.allclose(y, z["y_transform_ind"], rtol=1e-10, atol=1e-300)
    for focal in (1, 2):
        res = bn.get_pdp(x, [focal], "classification", 3, af, bn.SoftMax, post, [[0.0]] * S, dt)
        assert np.array_equal(res["feature"], z["pdp%d_feature" % focal])
        assert np.allclose(res["pdp"], z["pdp%d" % focal], rtol=1e-9, atol=1e-12)


def test_philox_chains_agree_statistically_at_the_c4_shape():
    """The same statistical comparison on the BASELINE config-4 network ([64,32] swish, 64 features, 10 classes, bias on
    the last layer; 20,000 rows), i.e. through the shape-specialised forward kernel and the 1,024-thread update kernel:
    3 chains x 1,500 iterations each way (burn-in 700, every 10th state), posterior means of log-likelihood and accuracy
    within between-chain Monte-Carlo error."""
    import npbnn_b200 as bn
    from npbnn_b200 import workloads as wl
    x, labels = wl.c4_data(22_000, seed=4)
    dat = {"data": x[:20_000], "labels": labels[:20_000].astype(int), "test_data": x[20_000:], "test_labels": labels[20_000:].astype(int)}
    out = {}
    for mode in ("host", "philox"):
        per_chain = []
        for ch in range(3):
            np.random.seed(40 + ch)
            bnn = bn.npBNN(dat, n_nodes=[64, 32], actFun=bn.ActFun(fun="swish"), use_bias_node=-1, seed=40 + ch)
            mcmc = bn.MCMC(bnn, n_iteration=1500, rng=mode, mcmc_id=ch)
            mcmc._rs = np.random.default_rng(900 + ch)
            rows = []
            for it in range(150):
                mcmc.run(bnn, 10)
                if it >= 70:
                    rows.append([mcmc._logLik, mcmc._accuracy, mcmc._test_accuracy, mcmc._acceptance_rate])
            assert mcmc._group.eng.last_kernel.startswith("k_fwd3<")
            per_chain.append(np.mean(rows, axis=0))
        per_chain = np.array(per_chain)
        out[mode] = (per_chain.mean(0), per_chain.std(0, ddof=1) / np.sqrt(3.0))
    (mh, sh), (mp, sp) = out["host"], out["philox"]
    se = np.sqrt(sh ** 2 + sp ** 2)
    gap = np.abs(mh - mp)
    print("host", mh, "philox", mp, "gap", gap, "se", se)
    assert gap[0] < max(4 * se[0], 0.01 * abs(mh[0]))            # log-likelihood
    assert gap[1] < max(4 * se[1], 0.01) and gap[2] < max(4 * se[2], 0.02)
    assert gap[3] < 0.15


def test_mode2_resampling_with_in_kernel_uniforms():
    """get_posterior_cat_prob(post_summary_mode=2) / sample_from_categorical with the uniforms generated inside the
    kernel (bnn_predict_sample_philox): O(N K) memory, shape-specialised kernel.  Checked against the injected-uniform
    path statistically (shares vs mean probabilities) and draw by draw against the generic kernel on the same
    counters (the two kernels sum the cumulative probabilities in different orders)."""
    import npbnn_b200 as bn
    from npbnn_b200 import _lib as L
    from npbnn_b200.engine import Engine, NetShape
    rng = np.random.default_rng(12)
    n, S, K = 6000, 96, 10
    shapes = [(64, 64), (32, 64), (K, 33)]
    x = rng.standard_normal((n, 64))
    sets = [[rng.normal(0, 0.25, s) for s in shapes] for _ in range(S)]
    eng = Engine(NetShape.from_weights(sets[0], 64, act="swish", lik=L.LIK_CATEGORICAL))
    fast = eng.predict_sample(x, sets, u=None, seed=2024)
    assert eng.last_kernel.startswith("k_fwd3<"), eng.last_kernel
    eng.set_option("force_generic", 1)
    gen = eng.predict_sample(x, sets, u=None, seed=2024)
    assert eng.last_kernel == "k_fwd_generic"
    eng.set_option("force_generic", 0)
    other = eng.predict_sample(x, sets, u=None, seed=2025)
    pred = eng.predict(x, sets, mean=True, dense=True)
    mean, dense = pred["mean"], pred["dense"]
    eng.close()
    pp = fast["post_predictions"]
    assert pp.shape == (n, S) and pp.min() >= 0 and pp.max() <= K - 1 and np.all(pp == np.round(pp))
    assert np.mean(pp == gen["post_predictions"]) > 0.9999           # same counters, same draws (up to rounding ties)
    assert np.mean(pp == other["post_predictions"]) < 0.9            # another seed, other draws
    # bookkeeping: shares are the per-row histogram of the draws, class_counts the per-set histogram
    hist = np.stack([(pp == k).mean(1) for k in range(K)], axis=1)
    assert np.array_equal(fast["predictions"], hist)
    assert np.array_equal(fast["class_counts"], np.stack([(pp == k).sum(0) for k in range(K)], axis=1))
    # statistics: a share is the mean of S independent Bernoulli(p_s) draws: E = mean probability, Var = sum p_s (1 - p_s)
    # / S^2; the standardised deviations over all (row, class) cells behave like N(0, 1)
    sd = np.sqrt(np.maximum((dense * (1 - dense)).sum(0), 1e-12)) / S
    zs = (fast["predictions"] - mean) / sd
    sel = (mean > 0.02) & (mean < 0.98)
    assert abs(zs[sel].mean()) < 0.02 and 0.95 < zs[sel].std() < 1.05
    # the API picks the in-kernel generator by size (or on request) and seeds it from numpy's global stream
    post = [{"weights": w, "alphas": [0.0]} for w in sets]
    af = bn.ActFun(fun="swish")
    np.random.seed(3)
    _, a = bn.get_posterior_cat_prob(x, post, post_summary_mode=2, actFun=af, output_act_fun=bn.SoftMax, return_dense=False,
                                     rng="philox")
    np.random.seed(3)
    b = bn.sample_from_categorical(x, post, actFun=af, output_act_fun=bn.SoftMax, rng="philox", post_predictions=False)
    assert np.array_equal(a, b["predictions"]) and b["post_predictions"] is None
    assert abs(((a - mean) / sd)[sel].mean()) < 0.02


def test_free_running_chains_draw_every_proposal_branch_on_the_device(tmp_path):
    """rng="philox" with trainable activation slopes, weight indicators (four layers) and feature indicators: the
    secondary proposals (BNN_env.py:416-431,449-460) are generated in k_mh_update.  The chain state must stay coherent:
    log-prior = calc_prior of the exported weights / indicators (+ the Exp(10) term of the slopes), log-likelihood =
    the likelihood of a fresh forward pass with the exported indicators, indicators in {0, 1}, slopes in [0, 1]."""
    import npbnn_b200 as bn
    from tests import golden_data
    dat = golden_data.synth_class(800, 6, 3, 13, 100)
    np.random.seed(13)
    prm = np.array([0.1, 0.3, 0.2])
    bnn = bn.npBNN(dat, n_nodes=[5, 4, 4], actFun=bn.ActFun(fun="genReLU", prm=prm, trainable=True), use_bias_node=2,
                   freq_indicator=0.3, prior_ind1=0.4, feature_indicators=True, seed=13)
    mcmc = bn.MCMC(bnn, n_iteration=2000, update_f=[0.2] * 4, adapt_stop=5, rng="philox",
                   init_additional_prob=float(np.log(10) * -np.sum(prm) * 10))
    mcmc.run(bnn, 400)
    assert mcmc._current_iteration == 400 and 0.02 < mcmc._acceptance_rate <= 1.0
    ind, fi = np.asarray(bnn._indicators), np.asarray(bnn._feature_indicators)
    assert set(np.unique(ind)) <= {0.0, 1.0} and set(np.unique(fi)) <= {0, 1}
    assert ind.mean() < 1.0 and fi.mean() < 1.0                       # both kinds of move happened and were accepted
    acc = np.asarray(bnn._act_fun._acc_prm)
    assert np.all(acc >= 0) and np.all(acc <= 1) and not np.allclose(acc, prm)
    # coherence of the exported state
    lp = bnn.calc_prior() + float(np.log(10) * -np.sum(acc) * 10)
    assert np.isclose(mcmc._logPrior, lp, rtol=1e-9), (mcmc._logPrior, lp)
    af = bn.ActFun(fun="genReLU", prm=acc)
    dt = bn.data_transform_obj(fi, bnn._feature_means)
    y = bn.RunPredictInd(bnn._data, bnn._w_layers, ind, af, bn.SoftMax, data_transform=dt)
    ll = bn.calc_likelihood(y, bnn._labels, np.arange(len(bnn._labels)))
    assert np.isclose(mcmc._logLik, ll, rtol=1e-9), (mcmc._logLik, ll)
    assert np.isclose(mcmc._accuracy, bn.CalcAccuracy(y, bnn._labels))
    # and through run_mcmc with the pipelined logger
    logger = bn.postLogger(bnn, filename="sec", wdir=str(tmp_path))
    bn.run_mcmc(bnn, mcmc, logger)
    head = open(logger._logfile).read().split("\n")[0].split("\t")
    assert "mean_ind" in head and "alpha_0" in head and "feature_ind_5" in head


def test_predict_sharded_chunked_upload_matches_one_pass():
    """predshard.predict_sharded with the posterior samples in pinned host memory uploads them in chunks on a side stream
    under the kernel of the previous chunk; mean probabilities must agree with the one-pass prediction to rounding of the
    last additions and the vote shares exactly (world = 1: no collective)."""
    import torch
    from npbnn_b200.engine import Engine, NetShape, flatten_weights
    from npbnn_b200 import predshard, workloads as wl
    rng = np.random.default_rng(8)
    n, S = 3000, 700
    x, _ = wl.c4_data(n, seed=2)
    base = flatten_weights(wl.c4_init_weights(1)[0])
    w = base[None, :] + rng.normal(0, 0.1, (S, base.size))
    eng = Engine(NetShape(64, list(wl.C4_SHAPES), act="swish", lik=0))
    ref = eng.predict(x, w, mean=True, votes=True)
    xd = torch.from_numpy(x).cuda()
    out = predshard.predict_sharded(eng, xd, torch.from_numpy(w).pin_memory(), n, 0, 1, votes=True)
    assert out["grid"] == (1, 1) and out["sets"] == (0, S)
    assert np.allclose(out["mean"].cpu().numpy(), ref["mean"], rtol=0, atol=1e-14)
    assert np.array_equal(np.rint(out["votes"].cpu().numpy() * S), np.rint(ref["votes"] * S))
    plain = predshard.predict_sharded(eng, xd, w, n, 0, 1, votes=True)                  # pageable numpy: one chunk
    assert np.allclose(plain["mean"].cpu().numpy(), ref["mean"], rtol=0, atol=1e-15)     # (mean * S / S)
    eng.close()


def test_predict_tf32_mode():
    """Opt-in reduced-precision prediction (option predict_tf32: 3xTF32 tensor-core contractions, FP32 activations and
    softmax, FP64 accumulation over the samples) against the FP64 kernel on the c4 / c5 network shape.  Stated tolerance:
    mean class probabilities to 5e-6 absolute; the argmax votes may flip only where two classes tie to that precision."""
    from npbnn_b200.engine import Engine, NetShape, flatten_weights
    from npbnn_b200 import workloads as wl
    rng = np.random.default_rng(12)
    n, S = 5003, 37
    x, _ = wl.c4_data(n, seed=4)
    base = flatten_weights(wl.c4_init_weights(1)[0])
    w = base[None, :] + rng.normal(0, 0.15, (S, base.size))
    for act in ("swish", "tanh", "ReLU"):
        eng = Engine(NetShape(64, list(wl.C4_SHAPES), act=act, lik=0))
        ref = eng.predict(x, w, mean=True, votes=True)
        assert eng.last_kernel.startswith("k_fwd3<"), eng.last_kernel
        eng.set_option("predict_tf32", 1)
        lp = eng.predict(x, w, mean=True, votes=True)
        assert eng.last_kernel.startswith("k_pred_tf32x3<"), eng.last_kernel
        eng.set_option("predict_tf32", 0)
        err = np.abs(lp["mean"] - ref["mean"]).max()
        assert err < 5e-6, (act, err)
        assert np.allclose(lp["mean"].sum(1), 1.0, atol=1e-6)
        flips = np.abs(np.rint(lp["votes"] * S) - np.rint(ref["votes"] * S)).sum() / 2
        assert flips <= 1e-4 * n * S, (act, flips)
        eng.set_option("predict_tf32", 2)                 # plain TF32 (one product): quick-look precision, 5e-3 stated
        lp1 = eng.predict(x, w, mean=True, votes=True)
        assert eng.last_kernel.startswith("k_pred_tf32x1<"), eng.last_kernel
        err1 = np.abs(lp1["mean"] - ref["mean"]).max()
        assert err < err1 < 5e-3, (act, err1)
        eng.close()
