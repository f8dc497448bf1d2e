"""CPU: pin the oracle (oracle/npbnn_oracle.py) to the reference's own outputs (tests/golden/)."""
import numpy as np
import pytest

from oracle import npbnn_oracle as orc
from tests import _golden as G

RTOL = 1e-12   # same FP64 arithmetic, different summation order only


def close(a, b, rtol=RTOL, atol=0.0):
    return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", G.CHAIN_CASES + ["syn_c3_shape:0", "syn_c3_shape:5"])
def test_chain_replay_matches_reference(name):
    if ":" in name:                       # one chain of the 8-chain block-masked golden (BASELINE config 3's network)
        name, chain = name.split(":")
        z, meta = G.load(name)
        z = G.ChainView(z, int(chain))
    else:
        z, meta = G.load(name)
    m = G.build_model(z, meta)
    s = G.build_sampler(m, meta)
    assert close(s.logLik, z["init_logLik"]), (s.logLik, float(z["init_logLik"]))
    assert close(s.logPrior, z["init_logPrior"])
    assert close(s.accuracy, z["init_accuracy"])
    assert close(s.test_accuracy, z["init_test_accuracy"])
    assert close(s.label_acc, z["init_label_acc"])
    assert np.array_equal(s.update_n, z["init_update_n"])
    n_steps = int(z["n_steps"])
    near_ties = 0
    for t in range(n_steps):
        inj = G.injection(z, meta, t)
        # the proposed-layer rule must reproduce what the reference did
        d = orc.mh_step(m, s, inj)
        ok = orc.layer_is_proposed(inj.rr, z["steps_freq_layer_update"][t])
        assert np.array_equal(ok.astype(np.int32), z["steps_proposed"][t]), t
        assert close(d["logLik_prime"], z["steps_logLik_prime"][t]), (t, d["logLik_prime"], float(z["steps_logLik_prime"][t]))
        # the golden records calc_prior() alone; mh_step adds additional_prob to it (BNN_env.py:481)
        assert close(d["logPrior_prime"] - d["additional_prob"], z["steps_logPrior_prime"][t], rtol=1e-11), t
        assert d["accepted"] == int(z["steps_accepted"][t]), t
        assert close(s.logLik, z["steps_logLik"][t]) and close(s.logPrior, z["steps_logPrior"][t])
        assert close(s.logPost, z["steps_logPost"][t])
        assert close(s.accuracy, z["steps_accuracy"][t]) and close(s.test_accuracy, z["steps_test_accuracy"][t])
        assert close(s.label_acc, z["steps_label_acc"][t])
        if meta.get("trainable"):
            assert close(m.alphas, z["steps_act_prm"][t]), t
        if meta["mode"] == "classification":
            assert close(s.label_freq, z["steps_label_freq"][t])
        assert close(s.acceptance_rate, z["steps_acceptance_rate"][t], rtol=1e-15)
        assert np.array_equal(s.update_n, z["steps_update_n"][t]), t
        assert close(s.update_f, z["steps_update_f"][t], rtol=1e-15)
        assert close(s.update_ws, z["steps_update_ws"][t], rtol=1e-15)
        assert close(s.freq_layer_update, z["steps_freq_layer_update"][t], rtol=1e-15)
        if meta["mode"] == "regression":
            assert close(np.asarray(m.error_prm) * np.ones(m.labels.shape[1]), z["steps_error_prm"][t])
    for i, w in enumerate(m.weights):
        assert np.array_equal(w, z["wN_%d" % i]), "final weights layer %d" % i
    y = orc.forward(m.x, m.weights, m.act, m.alphas, m.out_kind)
    assert close(y, z["yN"], rtol=1e-11, atol=1e-300)
    assert near_ties == 0


def test_survey_anchor_values():
    """SURVEY.md section 8c anchors for BASELINE config 1 and 2 (measured by the survey on the reference)."""
    z, meta = G.load("c1_classify")
    assert abs(float(z["init_logLik"]) - (-3621.998513641989)) < 1e-9
    assert abs(float(z["init_logPrior"]) - (-646.541793988718)) < 1e-9
    assert abs(float(z["steps_logLik"][99]) - (-3623.157147588077)) < 1e-9
    assert abs(float(z["steps_logPrior"][99]) - (-650.529100740772)) < 1e-9
    assert int(np.sum(z["steps_accepted"][:100])) == 89
    assert list(z["steps_update_n"][99]) == [27, 1, 1]
    z, meta = G.load("c2_regress_emp1")
    assert abs(float(z["init_logLik"]) - (-16509.872057635934)) < 1e-8
    assert abs(float(z["init_logPrior"]) - (-96.989733054569)) < 1e-9


def test_masks_match_create_mask():
    z, meta = G.load("masks")
    for si, spec in enumerate(meta["specs"]):
        for li, shape in enumerate(spec["shapes"]):
            got = orc.block_mask(tuple(shape), spec["indx_input_list"][li], spec["nodes_per_feature_list"][li])
            assert np.array_equal(got, z["m%d_%d" % (si, li)]), (si, li)


def test_posterior_predict_and_pdp():
    z, meta = G.load("predict")
    x = z["x"]
    for ci, case in enumerate(meta["cases"]):
        post = [[z["p%d_s%d_w%d" % (ci, j, li)] for li in range(3)] for j in range(meta["S"])]
        alphas = [case["alphas"]] * meta["S"] if case["alphas"] else None
        dense0, votes = orc.posterior_predict(x, post, case["act"], alphas, "softmax", 0)
        _, mean = orc.posterior_predict(x, post, case["act"], alphas, "softmax", 1)
        assert close(dense0, z["p%d_dense" % ci], rtol=1e-12, atol=1e-300)
        assert np.array_equal(votes, z["p%d_mode0" % ci])
        assert close(mean, z["p%d_mode1" % ci], rtol=1e-13)
        for focal in (1, 3):
            feats = z["p%d_pdp%d_feature" % (ci, focal)]
            gold = z["p%d_pdp%d" % (ci, focal)]
            for n in range(feats.shape[0]):
                mean_, lo, hi = orc.pdp_step(x, [focal], feats[n, :], post, case["act"], alphas)
                assert close(mean_, gold[n, :, 0], rtol=1e-12)
                assert close(lo, gold[n, :, 1], rtol=1e-12) and close(hi, gold[n, :, 2], rtol=1e-12)


def test_mc3_swaps_and_chain_states():
    """Replay the reference MC3 run: per-chain steps reseed default_rng(it + id) every step
    (BNN_env.py:383-384); swap rule BNN_mc3.py:98-112."""
    z, meta = G.load("mc3")
    nc, sf = meta["n_chains"], meta["swap_frequency"]
    labels = z["labels"].astype(np.int64)
    temps = np.array(z["temps0"])
    models, samplers = [], []
    for c in range(nc):
        m = orc.Model(x=np.array(z["x"]), labels=labels, weights=[np.array(z["w0_%d" % i]) for i in range(3)],
                      act=meta["act"], mode="classification", prior=1, prior_scale=np.ones(3))
        s = orc.make_sampler(m, temperature=temps[c], n_iteration=sf, adapt_f=meta["adapt_f"],
                             adapt_fM=meta["adapt_fM"], adapt_freq=meta["adapt_freq"], adapt_stop=meta["adapt_stop"])
        models.append(m); samplers.append(s)
    for it in range(meta["n_mc3_iterations"]):
        assert close(temps, z["it%d_temps_before" % it], rtol=0)
        for c in range(nc):
            m, s = models[c], samplers[c]
            s.temperature = temps[c]
            for _ in range(sf):
                rs = np.random.default_rng(s.it + c)
                orc.mh_step(m, s, rs=rs)
            for li in range(3):
                assert np.array_equal(m.weights[li], z["it%d_c%d_w%d" % (it, c, li)]), (it, c, li)
        lp = np.array([s.logPost for s in samplers])
        assert close(lp, z["it%d_logPost" % it])
        j, k = [int(v) for v in z["it%d_pair" % it]]
        temps, swapped, r = orc.mc3_swap(lp, temps, j, k, float(z["it%d_log_u" % it]))
        assert close(temps, z["it%d_temps_after" % it], rtol=0), it


def test_resample_categorical_matches_reference():
    """sample_from_categorical (BNN_lib.py:682-713) on the reference's own probability tensor and uniforms."""
    z, meta = G.load("sample_cat")
    pred, counts, drawn = orc.resample_categorical(np.array(z["dense"]), np.array(z["u"]))
    assert np.array_equal(pred, z["predictions"])
    assert np.array_equal(counts, z["class_counts"])
    assert np.array_equal(drawn, z["post_predictions"])
    # the oracle's forward reproduces the tensor the reference sampled from
    s = int(meta["S"])
    for j in range(s):
        w = [np.array(z["s%d_w%d" % (j, li)]) for li in range(3)]
        assert close(orc.forward(np.array(z["x"]), w, "swish", None, "softmax"), z["dense"][j])


@pytest.mark.parametrize("hp", [1, 2, 3])
def test_hyper_prior_scales_and_prior(hp):
    """sample_prior_scale / gibbs_step (BNN_env.py:196-219,534-538): the oracle's conjugate draw reproduces the
    reference's scales when numpy's global generator is in the reference's state, and calc_prior under the sampled
    scales (scalar / per input node / per weight broadcasting) reproduces the recorded log-priors."""
    z, meta = G.load("syn_hyper_p%d" % hp)
    np.random.seed(int(meta["seed"]))
    w0 = [np.random.normal(0, 0.1, z["w0_%d" % i].shape) for i in range(3)]       # init_weight_prm draws (BNN_mcmc.py:19-24)
    for i in range(3):
        assert np.array_equal(w0[i], z["w0_%d" % i])
    # mh_step draws from mcmc._rs only, so the global generator is now where the first Gibbs step found it
    n_gibbs = 0
    for t in range(int(meta["n_steps"])):
        if not int(z["steps_gibbs"][t]):
            continue
        w = [z["t%d_w_%d" % (t, i)] for i in range(3)]
        scales = orc.gibbs_prior_scales(w, hp)
        for i in range(3):
            ref = z["t%d_scale_%d" % (t, i)]
            assert np.shape(scales[i]) == ref.shape and np.array_equal(np.asarray(scales[i]), ref), (t, i)
        lp = orc.log_prior(w, 1, scales)
        assert close(lp, z["steps_logPrior"][t]), (t, lp, float(z["steps_logPrior"][t]))
        assert close(float(z["steps_logLik"][t]) + lp, z["steps_logPost"][t])
        assert int(z["steps_iteration"][t]) == t + 1
        n_gibbs += 1
    assert n_gibbs == int(meta["n_steps"]) // 4


@pytest.mark.parametrize("kind", ["weight", "feature", "both"])
def test_indicator_state_matches_reference(kind):
    """Weight indicators (first-layer weights times a 0/1 matrix in the forward pass, Bernoulli term in the prior,
    BNN_env.py:191-193,458-466) and feature indicators (masked features replaced by their training mean,
    BNN_env.py:9-17,423-431): the oracle's forward pass and prior on the reference's final state reproduce the
    predictions and the log-prior the reference recorded for it."""
    z, meta = G.load("syn_ind_%s" % kind)
    nl = len(meta["n_nodes"]) + 1
    w = [z["wN_%d" % i] for i in range(nl)]
    x = z["x"]
    w_fwd = list(w)
    if meta["freq_indicator"]:
        w_fwd[0] = w[0] * z["indN"]
        assert set(np.unique(z["indN"])) <= {0.0, 1.0} and z["indN"].mean() < 1.0       # the indicators did move
    if kind in ("feature", "both"):
        fi = z["steps_feature_ind"][-1]
        assert fi.min() == 0                                                              # some feature is masked
        x = orc.feature_transform(x, fi, z["x"].mean(axis=0))
    y = orc.forward(x, w_fwd, "tanh", None, "softmax")
    assert close(y, z["yN"], rtol=1e-11)
    lp = orc.log_prior(w, 1, np.ones(nl), indicators=z["indN"], freq_indicator=meta["freq_indicator"],
                       prior_ind1=meta["prior_ind1"])
    assert close(lp, z["steps_logPrior"][-1])
    assert close(orc.loglik_categorical(y, z["labels"].astype(int), None, None, 1.0), z["steps_logLik"][-1])
    assert np.array_equal(orc.update_binomial(np.array([1, 0, 1, 0]), np.array([1, 1, 0, 0])), [0, 1, 1, 0])
