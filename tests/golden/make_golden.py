"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference has no tests of its own, so parity is pinned to what
the reference code itself computes here.  Randomness is recorded, not re-derived: the MCMC
object's generator (`mcmc._rs`, BNN_env.py:362) is wrapped by a proxy that logs every
integers()/normal()/random() call, so each MH iteration's proposal indices, increments and
accept uniform can be replayed into the oracle and the CUDA path ("identical injected
weights and proposals", BASELINE.json north_star).
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, "/root/reference")
import np_bnn as bn  # noqa: E402  (the reference)

OUT = os.path.dirname(os.path.abspath(__file__))
EX = "/root/reference/example_files"


class RecRNG:
    """Recording proxy around numpy Generator (only the methods mh_step/UpdateNormal call)."""

    def __init__(self, rng):
        self._rng = rng
        self.log = []

    def integers(self, *a, **k):
        v = self._rng.integers(*a, **k)
        self.log.append(("integers", np.array(v)))
        return v

    def normal(self, *a, **k):
        v = self._rng.normal(*a, **k)
        self.log.append(("normal", np.array(v)))
        return v

    def random(self, *a, **k):
        v = self._rng.random(*a, **k)
        self.log.append(("random", np.array(v)))
        return v

    def binomial(self, *a, **k):
        v = self._rng.binomial(*a, **k)
        self.log.append(("binomial", np.array(v)))
        return v


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def record_chain(bnn, mcmc, n_steps, out, prefix="", rs_seed=None):
    """Run n_steps of the reference's mh_step, recording injections and outcomes.
    rs_seed: give this chain its own proposal stream (every MCMC object starts from default_rng(1234), BNN_env.py:362)."""
    if rs_seed is not None:
        mcmc._rs = np.random.default_rng(rs_seed)
    rec = RecRNG(mcmc._rs)
    mcmc._rs = rec
    lik_log, prior_log = [], []
    lik_f = mcmc._likelihood_f

    def lik_wrap(*a, **k):
        v = lik_f(*a, **k)
        lik_log.append(v)
        return v

    mcmc._likelihood_f = lik_wrap
    prior_f = bnn.calc_prior

    def prior_wrap(*a, **k):
        v = prior_f(*a, **k)
        prior_log.append(v)
        return v

    bnn.calc_prior = prior_wrap
    nl = bnn._n_layers
    out[prefix + "n_steps"] = n_steps
    out[prefix + "init_logLik"] = mcmc._logLik
    out[prefix + "init_logPrior"] = mcmc._logPrior
    out[prefix + "init_accuracy"] = mcmc._accuracy
    out[prefix + "init_test_accuracy"] = mcmc._test_accuracy
    out[prefix + "init_label_acc"] = np.asarray(mcmc._label_acc, dtype=np.float64)
    out[prefix + "init_label_freq"] = np.asarray(mcmc._label_freq, dtype=np.float64)
    out[prefix + "init_update_n"] = np.asarray(mcmc._update_n)
    for i in range(nl):
        out[prefix + "w0_%d" % i] = np.array(bnn._w_layers[i])
    per_step = {}      # key -> list over steps (stacked at the end)
    ragged = {i: {"ix": [], "iy": [], "dz": [], "off": [0]} for i in range(nl)}

    def put(key, v):
        per_step.setdefault(key, []).append(np.asarray(v))

    for t in range(n_steps):
        rec.log.clear(); lik_log.clear(); prior_log.clear()
        quiet(mcmc.mh_step, bnn)
        log = list(rec.log)
        # optional proposals drawn before rr (BNN_env.py:416-444): activation parameter, error parameter
        if bnn._act_fun._trainable:
            assert log[0][0] == "integers" and log[1][0] == "normal", log[:2]
            put("alpha_ix", int(log[0][1][0])); put("alpha_dz", float(log[1][1][0]))
            log = log[2:]
        if bnn._estimation_mode == "regression":
            if log[0][0] == "binomial":
                put("sigma_on", 1); put("sigma_ff", log[0][1].astype(np.int32)); put("sigma_u", log[1][1])
                log = log[2:]
            else:
                put("sigma_on", 0); put("sigma_ff", np.zeros(bnn._size_output, np.int32)); put("sigma_u", np.zeros(bnn._size_output))
        assert log[0][0] == "random" and log[0][1].shape == (nl,), log[0]
        assert log[-1][0] == "random" and log[-1][1].shape == ()
        rr = log[0][1]
        body = log[1:-1]
        assert len(body) % 3 == 0
        # which layers were proposed: same rule as BNN_env.py:446-457 (freq_indicator == 0 here)
        r2 = rr.copy(); r2[np.argmin(r2)] = 0
        proposed = []
        bi = 0
        # the freq_layer_update used in THIS step is the post-adaptation one, which equals the
        # value stored on the object after the step
        flu = np.asarray(mcmc._freq_layer_update)
        for i in range(nl):
            if r2[i] < flu[i]:
                ix, iy, dz = body[bi][1], body[bi + 1][1], body[bi + 2][1]
                assert body[bi][0] == "integers" and body[bi + 2][0] == "normal"
                ragged[i]["ix"].append(ix.astype(np.int32))
                ragged[i]["iy"].append(iy.astype(np.int32))
                ragged[i]["dz"].append(dz)
                proposed.append(1)
                bi += 3
            else:
                proposed.append(0)
            ragged[i]["off"].append(ragged[i]["off"][-1] + (len(ragged[i]["ix"][-1]) if proposed[-1] else 0))
        assert bi == len(body), (bi, len(body))
        put("rr", rr)
        put("proposed", np.array(proposed, dtype=np.int32))
        put("log_u", np.log(log[-1][1]))
        put("logLik_prime", lik_log[-1] if lik_log else 0.0)
        put("logPrior_prime", prior_log[-1])
        put("accepted", mcmc._last_accepted)
        put("logLik", mcmc._logLik)
        put("logPrior", mcmc._logPrior)
        put("logPost", mcmc._logPost)
        put("accuracy", mcmc._accuracy)
        put("test_accuracy", mcmc._test_accuracy)
        put("label_acc", np.asarray(mcmc._label_acc, dtype=np.float64))
        put("label_freq", np.asarray(mcmc._label_freq, dtype=np.float64))
        put("acceptance_rate", mcmc._acceptance_rate)
        put("update_n", np.asarray(mcmc._update_n))
        put("update_f", np.asarray(mcmc._update_f, dtype=np.float64))
        put("update_ws", np.array([w.flat[0] for w in mcmc._update_ws]))
        put("freq_layer_update", flu)
        if bnn._estimation_mode == "regression":
            put("error_prm", np.asarray(bnn._error_prm, dtype=np.float64) * np.ones(bnn._size_output))
        if bnn._act_fun._trainable:
            put("act_prm", np.asarray(bnn._act_fun._acc_prm, dtype=np.float64))
    for k, v in per_step.items():
        out[prefix + "steps_" + k] = np.stack(v)
    for i in range(nl):
        r = ragged[i]
        out[prefix + "prop_l%d_off" % i] = np.array(r["off"], dtype=np.int64)
        out[prefix + "prop_l%d_ix" % i] = np.concatenate(r["ix"]) if r["ix"] else np.zeros(0, np.int32)
        out[prefix + "prop_l%d_iy" % i] = np.concatenate(r["iy"]) if r["iy"] else np.zeros(0, np.int32)
        out[prefix + "prop_l%d_dz" % i] = np.concatenate(r["dz"]) if r["dz"] else np.zeros(0)
        out[prefix + "wN_%d" % i] = np.array(bnn._w_layers[i])
    out[prefix + "yN"] = np.array(mcmc._y)


def save(name, out, meta):
    out = dict(out)
    out["meta"] = np.array(json.dumps(meta))
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f kB, %d arrays)" % (path, os.path.getsize(path) / 1e3, len(out)))


def store_data(out, dat):
    out["x"] = np.asarray(dat["data"], dtype=np.float64)
    out["labels"] = np.asarray(dat["labels"])
    out["x_test"] = np.asarray(dat["test_data"], dtype=np.float64) if len(dat["test_data"]) else np.zeros((0, out["x"].shape[1]))
    out["labels_test"] = np.asarray(dat["test_labels"]) if len(dat["test_labels"]) else np.zeros((0,))


# ---------------------------------------------------------------------------------------
def case_c1(n_steps=120):
    """BASELINE config 1: bnn_classify.py:15-70 on the shipped example files."""
    dat = quiet(bn.get_data, EX + "/data_features.txt", EX + "/data_labels.txt", seed=1234, testsize=0.1,
                all_class_in_testset=1, header=1, cv=0, instance_id=1)
    np.random.seed(1234)
    bnn = quiet(bn.npBNN, dat, n_nodes=[5, 5], use_class_weights=0, actFun=bn.ActFun(fun="tanh"),
                use_bias_node=2, prior_f=1, p_scale=1, seed=1234, init_std=0.1, instance_weights=None)
    mcmc = bn.MCMC(bnn, update_f=[0.05, 0.05, 0.07], update_ws=[0.075, 0.075, 0.075], n_iteration=10000,
                   sampling_f=10, print_f=1000, n_post_samples=100, sample_from_prior=0, adapt_f=0.3, adapt_fM=0.6)
    out = {}
    store_data(out, dat)
    out["x"] = out["x"].astype(np.float64)
    record_chain(bnn, mcmc, n_steps, out)
    meta = dict(act="tanh", mode="classification", prior=1, p_scale=1.0, use_bias_node=2, n_nodes=[5, 5],
                update_f=[0.05, 0.05, 0.07], update_ws=[0.075] * 3, n_iteration=10000, adapt_f=0.3, adapt_fM=0.6,
                adapt_freq=1000, temperature=1.0, lik_temp=1.0, w_bound=float("inf"))
    save("c1_classify", out, meta)


def case_c2(empirical, n_steps=120):
    """BASELINE config 2: bnn_regress.py:19-52 data handling, [10,5] ReLU, Gaussian likelihood."""
    dat = quiet(bn.get_data, EX + "/data_features_reg.txt", EX + "/data_lab_reg.txt", seed=1234, testsize=0.1,
                all_class_in_testset=0, cv=0, header=True, from_file=True, instance_id=0, randomize_order=True,
                label_mode="regression")
    np.random.seed(1234)
    bnn = quiet(bn.npBNN, dat, n_nodes=[10, 5], estimation_mode="regression", actFun=bn.ActFun(fun="ReLU"),
                p_scale=1, use_bias_node=2, empirical_error=empirical)
    mcmc = bn.MCMC(bnn, update_ws=[0.025, 0.025, 0.05], update_f=[0.005, 0.005, 0.05], n_iteration=20000,
                   sampling_f=100, print_f=1000, n_post_samples=100, likelihood_tempering=1, adapt_f=0.3,
                   estimate_error=False)
    out = {}
    store_data(out, dat)
    record_chain(bnn, mcmc, n_steps, out)
    meta = dict(act="ReLU", mode="regression", prior=1, p_scale=1.0, use_bias_node=2, n_nodes=[10, 5],
                update_f=[0.005, 0.005, 0.05], update_ws=[0.025, 0.025, 0.05], n_iteration=20000, adapt_f=0.3,
                adapt_fM=1.0, adapt_freq=1000, temperature=1.0, lik_temp=1.0, w_bound=float("inf"),
                empirical_error=bool(empirical))
    save("c2_regress_emp%d" % int(empirical), out, meta)


def synth_class(n, f, k, seed, n_test=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n + n_test, f))
    wt = rng.normal(0, 1.0, (k, f))
    lab = np.argmax(x @ wt.T + rng.gumbel(size=(n + n_test, k)), axis=1)
    lab[:k] = np.arange(k)  # every class present in train
    if n_test:
        lab[n:n + k] = np.arange(k)
    return {"data": x[:n], "labels": lab[:n], "test_data": x[n:] if n_test else [],
            "test_labels": lab[n:] if n_test else []}


def case_synth(name, n=500, f=6, k=3, n_nodes=(4, 3), act="swish", alphas=None, prior=1, p_scale=1.0,
               bias=2, class_w=0, inst_w=False, temperature=1.0, lik_temp=1.0, n_test=64, mask_spec=None,
               n_steps=60, seed=7, adapt_f=0.0, adapt_fM=1.0, adapt_freq=1000, update_f=None, update_ws=None,
               n_iteration=1000, w_bound=np.inf, trainable=False, init_additional_prob=0):
    dat = synth_class(n, f, k, seed, n_test)
    iw = None
    if inst_w:
        iw = np.random.default_rng(seed + 1).uniform(0.2, 1.5, n)
    np.random.seed(seed)
    af = bn.ActFun(fun=act, prm=np.array(alphas) if alphas is not None else np.zeros(1), trainable=trainable)
    bnn = quiet(bn.npBNN, dat, n_nodes=list(n_nodes), use_class_weights=class_w, actFun=af, use_bias_node=bias,
                prior_f=prior, p_scale=p_scale, seed=seed, instance_weights=iw, w_bound=w_bound)
    out = {}
    if mask_spec is not None:
        m = bn.create_mask(bnn._w_layers, indx_input_list=mask_spec[0], nodes_per_feature_list=mask_spec[1])
        quiet(bnn.apply_mask, m)
        for i, mm in enumerate(m):
            out["mask_%d" % i] = mm
    mcmc = bn.MCMC(bnn, update_f=update_f, update_ws=update_ws, temperature=temperature, n_iteration=n_iteration,
                   likelihood_tempering=lik_temp, adapt_f=adapt_f, adapt_fM=adapt_fM, adapt_freq=adapt_freq,
                   init_additional_prob=init_additional_prob)
    store_data(out, dat)
    if iw is not None:
        out["inst_w"] = iw
    if class_w:
        out["class_w"] = np.asarray(bnn._class_w)
    record_chain(bnn, mcmc, n_steps, out)
    nl = len(n_nodes) + 1
    meta = dict(act=act, alphas=list(alphas) if alphas is not None else None, mode="classification", prior=prior,
                p_scale=p_scale, use_bias_node=bias, n_nodes=list(n_nodes),
                update_f=list(update_f) if update_f else [0.05] * nl,
                update_ws=list(update_ws) if update_ws else [0.075] * nl, n_iteration=n_iteration, adapt_f=adapt_f,
                adapt_fM=adapt_fM, adapt_freq=adapt_freq, temperature=temperature, lik_temp=lik_temp,
                w_bound=float(bnn._w_bound), trainable=bool(trainable), init_additional_prob=float(init_additional_prob))
    save(name, out, meta)


def case_regerr(n_steps=60):
    rng = np.random.default_rng(11)
    n, f, o = 400, 5, 2
    x = rng.standard_normal((n + 50, f))
    yv = np.stack([x[:, 0] - 0.5 * x[:, 1], np.sin(x[:, 2])], axis=1) + 0.1 * rng.standard_normal((n + 50, o))
    dat = {"data": x[:n], "labels": yv[:n], "test_data": x[n:], "test_labels": yv[n:]}
    np.random.seed(11)
    bnn = quiet(bn.npBNN, dat, n_nodes=[6, 4], estimation_mode="regression-error", actFun=bn.ActFun(fun="tanh"),
                use_bias_node=3, output_act_fun=bn.RegressTransformError)
    mcmc = bn.MCMC(bnn, n_iteration=1000)
    out = {}
    store_data(out, dat)
    record_chain(bnn, mcmc, n_steps, out)
    meta = dict(act="tanh", mode="regression-error", prior=1, p_scale=1.0, use_bias_node=3, n_nodes=[6, 4],
                update_f=[0.05] * 3, update_ws=[0.075] * 3, n_iteration=1000, adapt_f=0.0, adapt_fM=1.0,
                adapt_freq=1000, temperature=1.0, lik_temp=1.0, w_bound=float("inf"))
    save("syn_regress_error", out, meta)


# NOTE (measured while writing this generator): the error-parameter multiplier proposal of regression mode
# (BNN_env.py:435-442, estimate_error=True, empirical_error=False) cannot be recorded: every accepted step before
# _estimate_error resets bnn._error_prm to the scalar 1 (BNN_env.py:443-444,500-501), and the first proposal after it
# calls multiplier_proposal_vector(1, ...) which raises AttributeError ('int' object has no attribute 'shape',
# BNN_mcmc.py:105).  The path only survives when no proposal at all is accepted in the first _estimate_error steps.


def case_masks():
    """The three block_bnns.py examples (block_bnns.py:39-41,57-59,79-81) + a 40-feature c3-style one."""
    out = {}
    specs = [
        ([(6, 3), (2, 6), (2, 3)], [[0, 1, 2], [], []], [[2, 2, 2], [], []]),
        ([(9, 3), (6, 9), (2, 7)], [[0, 1, 2], [0, 0, 0, 1, 1, 1, 2, 2, 2], []], [[3, 3, 3], [2, 2, 2], []]),
        ([(9, 3), (5, 9), (2, 6)], [[0, 1, 1], [0, 0, 0, 1, 1, 1, 1, 1, 1], []], [[3, 6], [2, 3], []]),
        ([(24, 8), (16, 24), (5, 17)], [list(range(8)), sum(([g] * 3 for g in range(8)), []), []],
         [[3] * 8, [2] * 8, []]),
    ]
    meta = {"specs": []}
    for si, (shapes, idx, npf) in enumerate(specs):
        w = [np.ones(s) for s in shapes]
        m = bn.create_mask(w, indx_input_list=idx, nodes_per_feature_list=npf)
        for li, mm in enumerate(m):
            out["m%d_%d" % (si, li)] = mm
        meta["specs"].append({"shapes": shapes, "indx_input_list": idx, "nodes_per_feature_list": npf})
    save("masks", out, meta)


def case_predict():
    """get_posterior_cat_prob modes 0 and 1 (BNN_lib.py:352-397) and one get_pdp call (BNN_pdp.py:48-84)."""
    rng = np.random.default_rng(5)
    n, f, k, s = 300, 7, 4, 9
    x = rng.standard_normal((n, f))
    x[:, 3] = rng.integers(0, 3, n)   # an ordinal feature for the PDP grid
    out = {"x": x}
    meta = {"S": s, "cases": []}
    for ci, (act, bias, alphas) in enumerate([("swish", -1, None), ("tanh", 2, None), ("genReLU", 3, [0.01, 0.2])]):
        post = []
        for j in range(s):
            np.random.seed(100 * ci + j)
            w = bn.init_weight_prm([6, 5], f, k, init_std=0.1, bias_node=bias)
            w = [wi + rng.normal(0, 0.6, wi.shape) for wi in w]
            post.append({"weights": w, "alphas": list(alphas) if alphas else [0.0]})
            for li, wi in enumerate(w):
                out["p%d_s%d_w%d" % (ci, j, li)] = wi
        af = bn.ActFun(fun=act, prm=np.array(alphas) if alphas else np.zeros(1))
        for mode in (0, 1):
            dense_out, summ = bn.get_posterior_cat_prob(x, post, post_summary_mode=mode, actFun=af,
                                                        output_act_fun=bn.SoftMax)
            out["p%d_mode%d" % (ci, mode)] = summ
        out["p%d_dense" % ci] = dense_out
        # PDP on one continuous and one ordinal feature
        for focal in ([1], [3]):
            res = bn.get_pdp(x, focal, "classification", k, af, bn.SoftMax, [p["weights"] for p in post],
                             [p["alphas"] for p in post], None)
            out["p%d_pdp%d_feature" % (ci, focal[0])] = res["feature"]
            out["p%d_pdp%d" % (ci, focal[0])] = res["pdp"]
        meta["cases"].append({"act": act, "use_bias_node": bias, "alphas": alphas})
    save("predict", out, meta)


def case_predict_transform():
    """RunPredict / RunPredictInd / get_pdp with a feature-indicator data_transform (BNN_env.py:9-17, BNN_lib.py:245-272,
    BNN_pdp.py:48-84): features 2 and 5 are replaced by their means; the PDP runs over a kept and over a masked feature."""
    from np_bnn.BNN_env import data_transform_obj
    rng = np.random.default_rng(8)
    n, f, k, s = 200, 7, 3, 5
    x = rng.standard_normal((n, f))
    fi = np.array([1, 1, 0, 1, 1, 0, 1])
    dt = data_transform_obj(fi, x.mean(axis=0))
    post = []
    out = {"x": x, "fi": fi}
    for j in range(s):
        np.random.seed(50 + j)
        w = [wi + rng.normal(0, 0.5, wi.shape) for wi in bn.init_weight_prm([5, 4], f, k, init_std=0.1, bias_node=2)]
        post.append(w)
        for li, wi in enumerate(w):
            out["s%d_w%d" % (j, li)] = wi
    af = bn.ActFun(fun="tanh")
    ind = (rng.random(post[0][0].shape) < 0.7).astype(float)
    out["ind"] = ind
    out["y_transform"] = bn.RunPredict(x, post[0], af, bn.SoftMax, data_transform=dt)
    out["y_transform_ind"] = bn.RunPredictInd(x, post[0], ind, af, bn.SoftMax, data_transform=dt)
    for focal in ([1], [2]):
        res = bn.get_pdp(x, focal, "classification", k, af, bn.SoftMax, post, [[0.0]] * s, dt)
        out["pdp%d_feature" % focal[0]] = res["feature"]
        out["pdp%d" % focal[0]] = res["pdp"]
    save("predict_transform", out, {"S": s, "act": "tanh", "use_bias_node": 2})


def case_sample_cat():
    """sample_from_categorical (BNN_lib.py:682-713) and get_posterior_cat_prob mode 2: the global numpy stream is
    seeded, so the uniforms the reference consumed are np.random.random((n, S)) after the same seed."""
    rng = np.random.default_rng(9)
    n, f, k, s = 120, 5, 4, 6
    x = rng.standard_normal((n, f))
    out = {"x": x}
    post = []
    for j in range(s):
        np.random.seed(500 + j)
        w = bn.init_weight_prm([6, 5], f, k, init_std=0.1, bias_node=-1)
        w = [wi + rng.normal(0, 0.8, wi.shape) for wi in w]
        post.append({"weights": w, "alphas": [0.0]})
        for li, wi in enumerate(w):
            out["s%d_w%d" % (j, li)] = wi
    af = bn.ActFun(fun="swish")
    dense_out, _ = bn.get_posterior_cat_prob(x, post, post_summary_mode=1, actFun=af, output_act_fun=bn.SoftMax)
    out["dense"] = dense_out
    np.random.seed(77)
    res = bn.sample_from_categorical(posterior_weights=dense_out)
    np.random.seed(77)
    out["u"] = np.random.random((n, s))
    out["predictions"], out["class_counts"], out["post_predictions"] = res["predictions"], res["class_counts"], res["post_predictions"]
    np.random.seed(78)
    _, out["mode2"] = bn.get_posterior_cat_prob(x, post, post_summary_mode=2, actFun=af, output_act_fun=bn.SoftMax)
    save("sample_cat", out, {"S": s, "act": "swish", "use_bias_node": -1, "seed_direct": 77, "seed_mode2": 78})


def case_mc3():
    """Reference MC3.run_mcmc (BNN_mc3.py:87-126) on a toy problem: 3 chains, swap every 5 steps.
    The swap RNG (global np.random, BNN_mc3.py:99,109) is recorded by monkeypatching."""
    dat = synth_class(200, 5, 3, 21)
    np.random.seed(21)
    bnn = quiet(bn.npBNN, dat, n_nodes=[4, 3], use_bias_node=-1, seed=1, actFun=bn.ActFun(fun="swish"))
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        logger = bn.postLogger(bnn, filename="mc3gold", log_all_weights=0)
        np.random.seed(77)
        mc3 = quiet(bn.MC3, bnn, logger=logger, n_post_samples=10, sampling_f=5, n_iteration=40, n_chains=3,
                    swap_frequency=5, verbose=0, adapt_freq=50, adapt_f=0.1, adapt_fM=0.6, adapt_stop=1000)
        out = {}
        store_data(out, dat)
        for i in range(3):
            out["w0_%d" % i] = np.array(bnn._w_layers[i])
        out["temps0"] = np.array(mc3.temperatures, dtype=np.float64)
        # drive the reference's outer loop one swap period at a time so the state can be recorded
        rec = {"choice": [], "random": []}
        real_choice, real_random = np.random.choice, np.random.random

        def choice(*a, **k):
            v = real_choice(*a, **k); rec["choice"].append(np.array(v)); return v

        def rnd(*a, **k):
            v = real_random(*a, **k); rec["random"].append(v); return v

        n_it = mc3.n_mc3_iteration
        mc3.n_mc3_iteration = 1
        np.random.choice, np.random.random = choice, rnd
        try:
            for it in range(n_it):
                rec["choice"].clear(); rec["random"].clear()
                temps_before = np.array([a[1]._temperature for a in mc3.singleChainArgs], dtype=np.float64)
                quiet(mc3.run_mcmc)
                out["it%d_logPost" % it] = np.array([a[1]._logPost for a in mc3.singleChainArgs])
                out["it%d_logLik" % it] = np.array([a[1]._logLik for a in mc3.singleChainArgs])
                out["it%d_temps_before" % it] = temps_before
                out["it%d_temps_after" % it] = np.array([a[1]._temperature for a in mc3.singleChainArgs], dtype=np.float64)
                out["it%d_pair" % it] = rec["choice"][0].astype(np.int32)
                out["it%d_log_u" % it] = np.log(rec["random"][0])
                out["it%d_iteration" % it] = np.array([a[1]._current_iteration for a in mc3.singleChainArgs])
                for c in range(3):
                    for li in range(3):
                        out["it%d_c%d_w%d" % (it, c, li)] = np.array(mc3.singleChainArgs[c][0]._w_layers[li])
        finally:
            np.random.choice, np.random.random = real_choice, real_random
    finally:
        os.chdir(cwd)
    meta = dict(act="swish", mode="classification", prior=1, p_scale=1.0, use_bias_node=-1, n_nodes=[4, 3],
                n_chains=3, swap_frequency=5, n_mc3_iterations=int(n_it), adapt_freq=50, adapt_f=0.1, adapt_fM=0.6,
                adapt_stop=1000, update_f=[0.05] * 3, update_ws=[0.075] * 3, min_temperature=0.8)
    save("mc3", out, meta)


def case_hyper(hyper_p, n_steps=40, seed=11):
    """Hyper-priors (BNN_env.py:196-219,534-538): MH iterations with a Gibbs step on the prior scales every 4th
    iteration.  sample_prior_scale draws from numpy's GLOBAL generator (BNN_mcmc.py:124-141), the MH proposals from
    mcmc._rs, so seeding the global generator before npBNN(...) reproduces both streams."""
    dat = synth_class(300, 5, 3, seed, 40)
    np.random.seed(seed)
    bnn = quiet(bn.npBNN, dat, n_nodes=[4, 3], actFun=bn.ActFun(fun="tanh"), use_bias_node=2, prior_f=1, p_scale=1,
                hyper_p=hyper_p, seed=seed)
    mcmc = bn.MCMC(bnn, n_iteration=1000, update_f=[0.2, 0.2, 0.2])
    out = {}
    store_data(out, dat)
    for i, w in enumerate(bnn._w_layers):
        out["w0_%d" % i] = np.array(w)
    out["init_logLik"], out["init_logPrior"] = np.float64(mcmc._logLik), np.float64(mcmc._logPrior)
    rows = {k: [] for k in ("logLik", "logPrior", "logPost", "accepted", "iteration", "gibbs")}
    for t in range(n_steps):
        gibbs = (t % 4 == 3)
        if gibbs:
            mcmc.gibbs_step(bnn)
            for i, sc in enumerate(bnn._prior_scale):
                out["t%d_scale_%d" % (t, i)] = np.array(sc, dtype=np.float64)
                out["t%d_w_%d" % (t, i)] = np.array(bnn._w_layers[i])
        else:
            quiet(mcmc.mh_step, bnn)
        rows["logLik"].append(mcmc._logLik); rows["logPrior"].append(mcmc._logPrior); rows["logPost"].append(mcmc._logPost)
        rows["accepted"].append(mcmc._last_accepted); rows["iteration"].append(mcmc._current_iteration)
        rows["gibbs"].append(int(gibbs))
    for k, v in rows.items():
        out["steps_" + k] = np.array(v)
    for i, w in enumerate(bnn._w_layers):
        out["wN_%d" % i] = np.array(w)
    save("syn_hyper_p%d" % hyper_p, out, dict(hyper_p=hyper_p, seed=seed, n_steps=n_steps, act="tanh", n_nodes=[4, 3],
                                              use_bias_node=2, update_f=[0.2] * 3, n_iteration=1000))


def case_indicators(kind, n_steps=80, seed=13):
    """Indicator moves of mh_step.  kind "weight": npBNN(freq_indicator=0.3) on a FOUR-layer network (the branch reads
    update_f[3], BNN_env.py:460, so it only exists from four layers on); kind "feature": npBNN(feature_indicators=True)
    with adapt_stop=5 so that the branch (BNN_env.py:423-431) is live from iteration 6; kind "both".  UpdateBinomial
    draws from numpy's GLOBAL generator, the rest of the step from mcmc._rs."""
    dat = synth_class(300, 6, 3, seed, 50)
    np.random.seed(seed)
    weight, feature = kind in ("weight", "both"), kind in ("feature", "both")
    n_nodes = [4, 3, 3] if weight else [4, 3]
    bnn = quiet(bn.npBNN, dat, n_nodes=n_nodes, actFun=bn.ActFun(fun="tanh"), use_bias_node=2, prior_f=1, p_scale=1,
                freq_indicator=0.3 if weight else 0, prior_ind1=0.4, seed=seed, feature_indicators=True if feature else None)
    nl = len(n_nodes) + 1
    mcmc = bn.MCMC(bnn, n_iteration=1000, update_f=[0.2] * nl, adapt_stop=5)
    out = {}
    store_data(out, dat)
    for i, w in enumerate(bnn._w_layers):
        out["w0_%d" % i] = np.array(w)
    out["init_logLik"], out["init_logPrior"] = np.float64(mcmc._logLik), np.float64(mcmc._logPrior)
    rows = {k: [] for k in ("logLik", "logPrior", "logPost", "accepted", "accuracy", "test_accuracy", "mean_ind", "feature_ind")}
    for t in range(n_steps):
        quiet(mcmc.mh_step, bnn)
        rows["logLik"].append(mcmc._logLik); rows["logPrior"].append(mcmc._logPrior); rows["logPost"].append(mcmc._logPost)
        rows["accepted"].append(mcmc._last_accepted); rows["accuracy"].append(mcmc._accuracy)
        rows["test_accuracy"].append(mcmc._test_accuracy)
        rows["mean_ind"].append(np.mean(bnn._indicators))
        rows["feature_ind"].append(np.array(bnn._feature_indicators) if feature else np.zeros(0))
    for k, v in rows.items():
        out["steps_" + k] = np.array(v)
    for i, w in enumerate(bnn._w_layers):
        out["wN_%d" % i] = np.array(w)
    out["indN"] = np.array(bnn._indicators, dtype=np.float64)
    out["yN"] = np.array(mcmc._y)
    save("syn_ind_%s" % kind, out, dict(kind=kind, seed=seed, n_steps=n_steps, act="tanh", n_nodes=n_nodes, use_bias_node=2,
                                        freq_indicator=0.3 if weight else 0, prior_ind1=0.4, update_f=[0.2] * nl,
                                        adapt_stop=5, n_iteration=1000))


C3_MASK_SPEC = ([list(range(40)), sum(([g] * 3 for g in range(40)), []), []], [[3] * 40, [2] * 40, []])


def case_c4_shape(n=4000, n_test=500, n_steps=60, seed=41):
    """BASELINE config 4's network at a size the reference runs in seconds: F=64, [64,32] swish, K=10, bias -1
    (bnn_runner_MC3.py:28-34) -- the shape k_fwd3 and the 1,024-thread k_mh_update are compiled for."""
    case_synth("syn_c4_shape", n=n, f=64, k=10, n_nodes=(64, 32), act="swish", bias=-1, n_test=n_test, n_steps=n_steps,
               seed=seed, n_iteration=10000)


def case_c3_shape(n=2000, n_test=200, n_steps=40, n_chains=8, seed=43):
    """BASELINE config 3's network (block_bnns.py:39-81 pattern at SURVEY.md 8d's size): F=40, [120,80] tanh, K=5,
    bias -1, each feature -> 3 nodes -> 2 nodes; 8 chains over the same data, chain c initialised after
    np.random.seed(1000+c) and proposing from its own generator default_rng(1234+c)."""
    dat = synth_class(n, 40, 5, seed, n_test)
    out = {}
    store_data(out, dat)
    for c in range(n_chains):
        np.random.seed(1000 + c)
        bnn = quiet(bn.npBNN, dat, n_nodes=[120, 80], actFun=bn.ActFun(fun="tanh"), use_bias_node=-1, prior_f=1,
                    p_scale=1, seed=1000 + c)
        m = bn.create_mask(bnn._w_layers, indx_input_list=C3_MASK_SPEC[0], nodes_per_feature_list=C3_MASK_SPEC[1])
        quiet(bnn.apply_mask, m)
        if c == 0:
            for i, mm in enumerate(m):
                out["mask_%d" % i] = mm.astype(np.int8)
        mcmc = bn.MCMC(bnn, n_iteration=10000, mcmc_id=c)
        record_chain(bnn, mcmc, n_steps, out, prefix="c%d_" % c, rs_seed=1234 + c)
    meta = dict(act="tanh", alphas=None, mode="classification", prior=1, p_scale=1.0, use_bias_node=-1, n_nodes=[120, 80],
                update_f=[0.05] * 3, update_ws=[0.075] * 3, n_iteration=10000, adapt_f=0.0, adapt_fM=1.0, adapt_freq=1000,
                temperature=1.0, lik_temp=1.0, w_bound=float("inf"), n_chains=n_chains)
    save("syn_c3_shape", out, meta)


# ---------------------------------------------------------------------------------------
# host-side callers around the path: file readers, summaries, the script flows
# ---------------------------------------------------------------------------------------
sys.path.insert(0, os.path.dirname(OUT))
import golden_data  # noqa: E402  (tests/golden_data.py: the synthetic example tables)

GET_DATA_CASES = {        # name -> (feature table, label table, kwargs): the calls the five scripts make
    "classify": ("features", "labels", dict(seed=1234, testsize=0.1, all_class_in_testset=1, header=1, cv=0, instance_id=1)),
    "mc3": ("features", "labels", dict(seed=1234, testsize=0.1, all_class_in_testset=1, header=1, instance_id=1)),
    "all": ("features", "labels", dict(testsize=0, header=1, instance_id=1)),
    "unlabeled": ("unlabeled", None, dict(header=1, instance_id=1)),
    "tail_split": ("features", "labels", dict(seed=5, testsize=0.2, all_class_in_testset=0, header=1, instance_id=1)),
    "cv2_subset": ("features", "labels", dict(seed=9, testsize=0.1, header=1, instance_id=1, cv=2, feature_indx=[0, 3, 5],
                                              batch_training=50)),
    "regress": ("features_reg", "labels_reg", dict(seed=1234, testsize=0.1, all_class_in_testset=0, cv=0, header=True,
                                                   from_file=True, instance_id=0, randomize_order=True, label_mode="regression")),
    "regress_ordered": ("features_reg", "labels_reg", dict(seed=3, testsize=0.1, header=0, instance_id=0,
                                                           randomize_order=False, label_mode="regression")),
}


def _store_dat(out, prefix, dat):
    for k, v in dat.items():
        a = np.asarray(v)
        if a.dtype.kind in "OU":
            a = a.astype(str)
        out[prefix + k] = a


def case_hostlib():
    """Outputs of the reference's host-side helpers (BNN_files.py, BNN_lib.py:195-348,627-679, BNN_mcmc.py:27-150,
    BNN_lik.py) on seeded inputs."""
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        paths = golden_data.write_example_tables(tmp)
        for name, (f, l, kw) in GET_DATA_CASES.items():
            dat = quiet(bn.get_data, paths[f], paths[l] if l else None, **kw)
            _store_dat(out, "gd_%s_" % name, dat)
    rng = np.random.default_rng(23)
    n, k = 400, 5
    z = rng.normal(0, 2.0, (n, k))
    probs = bn.SoftMax(z)
    prior = bn.SoftMax(rng.normal(0, 0.3, (n, k)))
    lab = rng.integers(0, k - 1, n)                 # class k-1 never appears in the labels
    lab[:50] = np.argmax(probs[:50], axis=1) % (k - 1)
    out["probs"], out["prior_probs"], out["lab"], out["z"] = probs, prior, lab, z
    out["CalcAccuracy"] = bn.CalcAccuracy(probs, lab)
    out["CalcAccuracy3"] = bn.CalcAccuracy(np.stack([probs, prior]), lab)
    out["CalcLabelAccuracy"] = bn.CalcLabelAccuracy(probs, lab)
    out["CalcLabelFreq"] = bn.CalcLabelFreq(probs)
    out["CalcConfusionMatrix"] = bn.CalcConfusionMatrix(probs, lab).values
    out["CalcTP"], out["CalcFP"] = bn.CalcTP(probs, lab, threshold=0.8), bn.CalcFP(probs, lab, threshold=0.8)
    out["CalcTP_BF"], out["CalcFP_BF"] = bn.CalcTP_BF(probs, prior, lab, threshold=20), bn.CalcFP_BF(probs, prior, lab, threshold=20)
    th = bn.get_accuracy_threshold(probs, lab, threshold=0.75)
    out["thr_predictions"], out["thr_accuracy"], out["thr_retained"] = th["predictions"], th["accuracy"], th["retained_samples"]
    out["thr_cm"] = th["confusion_matrix"].values
    keep = np.where(np.max(probs, axis=1) > 0.9)[0]
    out["low_pp"] = bn.turn_low_pp_instances_to_nan(probs, keep)
    out["hpd"] = np.array(bn.calcHPD(z[:, 0], 0.9))
    yreg, labreg = rng.normal(size=(50, 4)), rng.normal(size=(50, 2))
    out["yreg"], out["labreg"] = yreg, labreg
    out["CalcAccuracyRegression"] = bn.CalcAccuracyRegression(yreg, labreg)
    out["CalcLabelAccuracyRegression"] = bn.CalcLabelAccuracyRegression(yreg, labreg)
    out["relu"], out["leaky"] = bn.relu_f(z.copy(), 0), bn.leaky_relu_f(z.copy(), 0.2)
    out["swish"], out["tanh"] = bn.swish_f(z.copy(), 0), bn.tanh_f(z.copy(), 0)
    out["softplus"], out["regerr"] = bn.SoftPlus(z), bn.RegressTransformError(z.copy())
    out["lik_cat"] = bn.calc_likelihood(probs, lab, np.arange(n), class_weight=np.linspace(0.5, 1.5, k), lik_temp=0.7)
    out["lik_cat_iw"] = bn.calc_likelihood(probs, lab, np.arange(n), instance_weight=np.linspace(0.1, 2, n))
    out["lik_reg"] = bn.calc_likelihood_regression(yreg[:, :2], labreg, None, lik_temp=0.9, sig2=np.array([0.5, 2.0]))
    out["lik_regerr"] = bn.calc_likelihood_regression_error(bn.RegressTransformError(yreg.copy()), labreg, None)
    w = rng.normal(size=(7, 6))
    d = np.full(w.shape, 0.3)
    out["w"] = w
    for name in ("UpdateNormal", "UpdateFixedNormal", "UpdateNormalNormalized"):
        zz, (ix, iy), h = getattr(bn, name)(w, d=d, n=9, Mb=1.0, mb=-1.0, rs=np.random.default_rng(5))
        out[name + "_z"], out[name + "_ix"], out[name + "_iy"], out[name + "_h"] = zz, ix, iy, h
    zz, ix, h = bn.UpdateNormal1D(w[0], d=0.05, n=2, rs=np.random.default_rng(5))
    out["UpdateNormal1D_z"], out["UpdateNormal1D_ix"] = zz, ix
    np.random.seed(31)
    out["UpdateUniform_z"] = bn.UpdateUniform(w, d=d, n=4)[0]
    out["UpdateBinomial"] = bn.UpdateBinomial(np.ones((3, 4)), 0.5, (3, 4))
    q, _, u = bn.multiplier_proposal_vector(np.array([1.0, 2.0, 3.0]), d=1.2, f=0.6, rs=np.random.default_rng(5))
    out["mpv_q"], out["mpv_u"] = q, u
    out["mp"] = np.array(bn.multiplier_proposal(2.0, d=1.1)[::2])
    out["gibbs_vec"] = bn.GibbsSampleNormStdGammaVector(w.flatten())
    out["gibbs_2d"], out["gibbs_one"] = bn.GibbsSampleNormStdGamma2D(w), bn.GibbsSampleNormStdGammaONE(w)
    out["gibbs_rate"] = bn.GibbsSampleGammaRateExp(np.array([0.5, 1.0, 2.0]), 2.0)
    cnt = rng.integers(0, 20, (50, 2)).astype(float)
    out["cnt"] = cnt
    out["poi"], out["negbin"] = bn.poi_likelihood(yreg, cnt), bn.negbin_likelihood(yreg, cnt)
    out["negbin2d"], out["negbin10"] = bn.negbin_likelihood2d(yreg, cnt), bn.negbin_likelihood_base10(yreg * 0.3, cnt)
    out["gamma"] = bn.gamma_likelihood(yreg * 0.1, cnt[:, :1] + 3.0)
    out["negbin_acc"], out["negbin2d_acc"], out["poi_acc"] = bn.negbin_acc(yreg, cnt), bn.negbin2d_acc(yreg, cnt), bn.poi_acc(yreg, cnt)
    out["negbin_acc10"] = bn.negbin_acc_base10(yreg * 0.3, cnt)
    out["assign_indx"] = bn.assign_indx(["b", "a", "b", "c", "a"])
    out["unique_unsorted"] = bn.unique_unsorted(np.array([3, 1, 3, 2, 1]))
    xs = rng.normal(size=(30, 4))
    xs[:, 1] = rng.integers(0, 3, 30)
    out["xs"], out["feature_summary"] = xs, bn.get_feature_summary(xs, [0, 1])
    mu, var = bn.RecurMeanVar(4, [np.zeros((7, 6)), np.ones((7, 6))], w, (np.array([0, 2]), np.array([1, 3])))
    out["recur_mu"], out["recur_var"] = mu, var
    # one hidden layer through the reference's forward helpers (device-side in this repository)
    x = rng.normal(size=(40, 5))
    wb = rng.normal(size=(6, 6))
    out["rh_x"], out["rh_w"] = x, wb
    out["mmd_bias"], out["mmd_nobias"] = bn.MatrixMultiplicationD(x, wb), bn.MatrixMultiplicationD(x, wb[:, 1:])
    for fun, prm in (("ReLU", None), ("genReLU", [0.1, 0.3]), ("swish", None), ("tanh", None)):
        af = bn.ActFun(fun=fun, prm=np.array(prm) if prm else np.zeros(1))
        out["rh_" + fun] = bn.RunHiddenLayer(x, wb, af, 1 if prm else 0)
    out["rh_none"] = bn.RunHiddenLayer(x, wb, False, 2)
    save("hostlib", out, {"get_data_cases": {k: [v[0], v[1], v[2]] for k, v in GET_DATA_CASES.items()}})


def _read_rows(path):
    return [r.split("\t") for r in open(path).read().strip().split("\n")]


def case_flow_classify():
    """bnn_classify.py:15-154 with shortened chains on the synthetic tables: get_data -> npBNN -> MCMC -> postLogger ->
    run_mcmc -> predictBNN (test set, all data, unlabeled) -> restart from the pickle -> feature_importance."""
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        paths = golden_data.write_example_tables(tmp)
        os.chdir(tmp)
        try:
            dat = quiet(bn.get_data, paths["features"], paths["labels"], seed=1234, testsize=0.1, all_class_in_testset=1,
                        header=1, cv=0, instance_id=1)
            np.random.seed(1234)
            bnn = quiet(bn.npBNN, dat, n_nodes=[5, 5], use_class_weights=0, actFun=bn.ActFun(fun="tanh"), use_bias_node=2,
                        prior_f=1, p_scale=1, seed=1234, init_std=0.1, instance_weights=None)
            mcmc = bn.MCMC(bnn, update_f=[0.05, 0.05, 0.07], update_ws=[0.075, 0.075, 0.075], n_iteration=400,
                           sampling_f=10, print_f=100, n_post_samples=20, sample_from_prior=0, adapt_f=0.3, adapt_fM=0.6)
            logger = bn.postLogger(bnn, filename="BNN_cv0", log_all_weights=0)
            quiet(bn.run_mcmc, bnn, mcmc, logger)
            out["log_rows"] = np.array(_read_rows(logger._logfile)[1:], dtype=np.float64)
            out["log_head"] = np.array(_read_rows(logger._logfile)[0])
            pr = quiet(bn.predictBNN, dat["test_data"], pickle_file=logger._pklfile, test_labels=dat["test_labels"],
                       instance_id=dat["id_test_data"], fname=dat["file_name"], post_summary_mode=0)
            out["test_pp"], out["test_acc"], out["test_cm"] = pr["post_prob_predictions"], pr["mean_accuracy"], pr["confusion_matrix"]
            out["test_files"] = np.array(sorted(os.listdir(tmp)))
            out["test_mean_pr_txt"] = np.array(open(dat["file_name"] + "_BNN_cv0_l5_5_pred_mean_pr.txt").read())
            out["test_accuracy_txt"] = np.array(open(dat["file_name"] + "_BNN_cv0_l5_5_accuracy.txt").read())
            dat_all = quiet(bn.get_data, paths["features"], paths["labels"], testsize=0, header=1, instance_id=1)
            pa = quiet(bn.predictBNN, dat_all["data"], pickle_file=logger._pklfile, test_labels=dat_all["labels"],
                       instance_id=dat_all["id_data"], fname="all_data", post_summary_mode=1)
            out["all_pp"], out["all_acc"] = pa["post_prob_predictions"], pa["mean_accuracy"]
            new = quiet(bn.get_data, paths["unlabeled"], header=1, instance_id=1)
            pn = quiet(bn.predictBNN, new["data"], pickle_file=logger._pklfile, instance_id=new["id_data"], fname=new["file_name"])
            out["new_pp"] = pn["post_prob_predictions"]
            thr = quiet(bn.get_posterior_threshold, logger._pklfile, target_acc=0.5, post_summary_mode=1)
            out["threshold_row"] = np.asarray(thr)
            pc = quiet(bn.predictBNN, new["data"], pickle_file=logger._pklfile, post_cutoff=0.6, post_summary_mode=1, fname="cut")
            out["cut_pp"] = pc["post_prob_predictions"]
            # restart from the pickle (bnn_classify.py:114-134)
            bnn2 = quiet(bn.npBNN, dat, n_nodes=[5, 5], use_bias_node=1, prior_f=1, p_scale=1, pickle_file=logger._pklfile,
                         seed=1234, actFun=bn.ActFun(fun="tanh"))
            for i, w in enumerate(bnn2._w_layers):
                out["restart_w%d" % i] = np.array(w)
            # bn.feature_importance (bnn_classify.py:137-152) cannot be recorded in this container: the reference assigns
            # float columns into a string-typed frame (BNN_lib.py:580), which pandas 3.0 rejects with TypeError.
        finally:
            os.chdir(cwd)
    save("flow_classify", out, {"n_iteration": 400})


def case_flow_mc3():
    """bnn_runner_MC3.py:17-79 with a shortened run: np.random.seed -> get_data -> npBNN -> postLogger -> MC3.run_mcmc
    -> predictBNN on the test set."""
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        paths = golden_data.write_example_tables(tmp)
        os.chdir(tmp)
        try:
            np.random.seed(1234)
            dat = quiet(bn.get_data, paths["features"], paths["labels"], seed=1234, testsize=0.1, all_class_in_testset=1,
                        header=1, instance_id=1)
            bnn = quiet(bn.npBNN, dat, n_nodes=[5, 5], use_bias_node=-1, seed=1, init_std=0.1)
            logger = bn.postLogger(bnn, filename="BNNMC3", log_all_weights=0)
            mc3 = quiet(bn.MC3, bnn, logger=logger, n_post_samples=100, sampling_f=100, n_iteration=300, n_chains=4,
                        swap_frequency=20, verbose=1)
            quiet(mc3.run_mcmc)
            out["log_rows"] = np.array(_read_rows(logger._logfile)[1:], dtype=np.float64)
            out["final_logPost"] = np.array([a[1]._logPost for a in mc3.singleChainArgs])
            out["final_temps"] = np.array([a[1]._temperature for a in mc3.singleChainArgs], dtype=np.float64)
            pr = quiet(bn.predictBNN, dat["test_data"], pickle_file=logger._pklfile, test_labels=dat["test_labels"],
                       instance_id=dat["id_test_data"])
            out["test_pp"], out["test_acc"], out["test_cm"] = pr["post_prob_predictions"], pr["mean_accuracy"], pr["confusion_matrix"]
            b2, m2, l2 = bn.load_obj(logger._pklfile)
            out["n_samples"] = len(l2._post_weight_samples)
            for i, w in enumerate(l2._post_weight_samples[-1]["weights"]):
                out["last_sample_w%d" % i] = np.array(w)
        finally:
            os.chdir(cwd)
    save("flow_mc3", out, {"n_iteration": 300, "swap_frequency": 20, "n_chains": 4})


def case_flow_regress():
    """bnn_regress.py:19-92 (plots left out) with a shortened chain: the attributes the script reads (mcmc._update_n,
    mcmc._accuracy_lab_f(mcmc._y, labels), mcmc._y, mcmc._y_test) and its RunPredict loop over the pickled samples."""
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        paths = golden_data.write_example_tables(tmp)
        os.chdir(tmp)
        try:
            np.random.seed(1234)
            dat = quiet(bn.get_data, paths["features_reg"], paths["labels_reg"], seed=1234, testsize=0.1, all_class_in_testset=0,
                        cv=0, header=True, from_file=True, instance_id=0, randomize_order=True, label_mode="regression")
            bnn = quiet(bn.npBNN, dat, n_nodes=[6, 4], estimation_mode="regression", actFun=bn.ActFun(fun="tanh"), p_scale=1,
                        use_bias_node=2, empirical_error=True)
            mcmc = bn.MCMC(bnn, update_ws=[0.025, 0.025, 0.05], update_f=[0.005, 0.005, 0.05], n_iteration=300, sampling_f=20,
                           print_f=100, n_post_samples=10, likelihood_tempering=1, adapt_f=0.3, estimate_error=False)
            out["update_n"] = np.asarray(mcmc._update_n)
            out["label_acc0"] = mcmc._accuracy_lab_f(mcmc._y, bnn._labels)
            logger = bn.postLogger(bnn, filename="testM", log_all_weights=0)
            quiet(bn.run_mcmc, bnn, mcmc, logger)
            out["y"], out["y_test"] = np.array(mcmc._y), np.array(mcmc._y_test)
            out["log_rows"] = np.array(_read_rows(logger._logfile)[1:], dtype=np.float64)
            b2, m2, l2 = bn.load_obj(logger._pklfile)
            ps = l2._post_weight_samples
            preds = []
            for i in range(len(ps)):
                af = b2._act_fun
                af.reset_prm(ps[i]["alphas"])
                preds.append(bn.RunPredict(b2._data, ps[i]["weights"], actFun=af, output_act_fun=b2._output_act_fun))
            out["post_preds"] = np.array(preds)
            est = bn.get_posterior_est(logger._pklfile)
            out["prm_mean"], out["prm_mean_test"] = est["prm_mean"], est["prm_mean_test"]
            out["error_prm"] = np.array(est["error_prm"], dtype=np.float64)
            pd_ = bn.pdp(logger._pklfile, [[0], [1, 2]])
            out["pdp0_feature"], out["pdp0"], out["pdp1"] = pd_[0]["feature"], pd_[0]["pdp"], pd_[1]["pdp"]
        finally:
            os.chdir(cwd)
    save("flow_regress", out, {"n_iteration": 300})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "hyper":          # regenerate only the hyper-prior cases
        for hp in (1, 2, 3):
            case_hyper(hp)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "host":            # host-side helpers and the script flows
        if "flows" not in sys.argv:
            case_hostlib()
        case_flow_classify()
        case_flow_mc3()
        case_flow_regress()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "shapes":          # the two headline network shapes
        case_c4_shape()
        case_c3_shape()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c2":              # BASELINE config 2 only
        case_c2(True)
        case_c2(False)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "indicators":
        for kind in ("weight", "feature", "both"):
            case_indicators(kind)
        case_predict_transform()
        sys.exit(0)
    case_c1()
    case_c2(True)
    case_c2(False)
    case_synth("syn_swish_cauchy", act="swish", prior=2, p_scale=0.7, bias=2)
    case_synth("syn_genrelu_laplace", act="genReLU", alphas=[0.01, 0.3], prior=3, p_scale=1.3, bias=1)
    case_synth("syn_uniform_bound", act="tanh", prior=0, p_scale=0.25, bias=3, update_ws=[0.2, 0.2, 0.2])
    case_synth("syn_classw_temp", act="ReLU", class_w=1, temperature=0.8, lik_temp=0.9, bias=-1)
    case_synth("syn_instw", act="swish", inst_w=True, bias=0)
    case_synth("syn_adapt", act="tanh", adapt_f=0.3, adapt_fM=0.6, adapt_freq=10, n_iteration=1000, n_steps=80,
               update_f=[0.3, 0.3, 0.3])
    case_synth("syn_block_mask", f=8, k=5, n_nodes=(24, 16), act="tanh", bias=-1,
               mask_spec=([list(range(8)), sum(([g] * 3 for g in range(8)), []), []], [[3] * 8, [2] * 8, []]))
    case_regerr()
    case_masks()
    case_predict()
    case_mc3()
    case_sample_cat()
    # init_additional_prob = the Exp(10) term of the initial parameters, so that proposals are not all rejected
    case_synth("syn_trainable_genrelu", act="genReLU", alphas=[0.1, 0.3], trainable=True, bias=2,
               init_additional_prob=float(np.log(10) * -np.sum([0.1, 0.3]) * 10))
    case_synth("syn_trainable_tanh", act="tanh", alphas=[0.2, 0.9], trainable=True, bias=1,
               init_additional_prob=float(np.log(10) * -np.sum([0.2, 0.9]) * 10))
    for hp in (1, 2, 3):
        case_hyper(hp)
    for kind in ("weight", "feature", "both"):
        case_indicators(kind)
    case_predict_transform()
    case_c4_shape()
    case_c3_shape()
    case_hostlib()
    case_flow_classify()
    case_flow_mc3()
    case_flow_regress()
