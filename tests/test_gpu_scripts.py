"""GPU: the reference's example scripts, run against `import np_bnn as bn` (this repository's drop-in), reproduce
the files and numbers the unmodified reference produced for the same calls (tests/golden/flow_*.npz, written by
tests/golden/make_golden.py: case_flow_*).  The statements are the scripts' own (bnn_classify.py:15-134,
bnn_runner_MC3.py:17-79, bnn_regress.py:19-92 without the plots, block_bnns.py:13-81) with shorter chains and the
synthetic tables of tests/golden_data.write_example_tables in place of example_files/."""
import os

import numpy as np
import pytest

from tests import _golden as G
from tests import golden_data

pytestmark = pytest.mark.gpu


def _rows(path):
    return [r.split("\t") for r in open(path).read().strip().split("\n")]


def _close(a, b, rtol=1e-9, atol=1e-12):
    return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    return golden_data.write_example_tables(str(tmp_path))


def test_bnn_classify_script(workdir, tmp_path):
    import np_bnn as bn
    z, _ = G.load("flow_classify")
    rseed = 1234
    f, l = workdir["features"], workdir["labels"]
    cross_validation_batch = 0
    dat = bn.get_data(f, l, seed=rseed, testsize=0.1, all_class_in_testset=1, header=1, cv=cross_validation_batch,
                      instance_id=1)
    n_nodes_list = [5, 5]
    activation_function = bn.ActFun(fun="tanh")
    np.random.seed(1234)                 # the generator seeds numpy before npBNN (the script itself leaves it unseeded)
    bnn_model = bn.npBNN(dat, n_nodes=n_nodes_list, use_class_weights=0, actFun=activation_function, use_bias_node=2,
                         prior_f=1, p_scale=1, seed=rseed, init_std=0.1, instance_weights=None)
    mcmc = bn.MCMC(bnn_model, update_f=[0.05, 0.05, 0.07], update_ws=[0.075, 0.075, 0.075], n_iteration=400, sampling_f=10,
                   print_f=100, n_post_samples=20, sample_from_prior=0, adapt_f=0.3, adapt_fM=0.6)
    logger = bn.postLogger(bnn_model, filename="BNN_cv%s" % cross_validation_batch, log_all_weights=0)
    bn.run_mcmc(bnn_model, mcmc, logger)
    assert mcmc._group.eng.launch_count > 0
    rows = _rows(logger._logfile)
    assert rows[0] == list(z["log_head"])
    got = np.array(rows[1:], dtype=np.float64)
    assert got.shape == z["log_rows"].shape and _close(got, z["log_rows"])        # the reference's chain, row by row
    post_pr_test = bn.predictBNN(dat["test_data"], pickle_file=logger._pklfile, test_labels=dat["test_labels"],
                                 instance_id=dat["id_test_data"], fname=dat["file_name"], post_summary_mode=0)
    assert np.array_equal(post_pr_test["post_prob_predictions"], z["test_pp"])
    assert post_pr_test["mean_accuracy"] == float(z["test_acc"])
    assert np.array_equal(post_pr_test["confusion_matrix"], z["test_cm"])
    assert sorted(os.listdir(tmp_path)) == list(z["test_files"])
    assert open(dat["file_name"] + "_BNN_cv0_l5_5_pred_mean_pr.txt").read() == str(z["test_mean_pr_txt"])
    assert open(dat["file_name"] + "_BNN_cv0_l5_5_accuracy.txt").read() == str(z["test_accuracy_txt"])
    dat_all = bn.get_data(f, l, testsize=0, header=1, instance_id=1)
    post_pr_all = bn.predictBNN(dat_all["data"], pickle_file=logger._pklfile, test_labels=dat_all["labels"],
                                instance_id=dat_all["id_data"], fname="all_data", post_summary_mode=1)
    assert _close(post_pr_all["post_prob_predictions"], z["all_pp"]) and post_pr_all["mean_accuracy"] == float(z["all_acc"])
    new_dat = bn.get_data(f=workdir["unlabeled"], header=1, instance_id=1)
    post_pr_new = bn.predictBNN(new_dat["data"], pickle_file=logger._pklfile, instance_id=new_dat["id_data"],
                                fname=new_dat["file_name"])
    assert np.array_equal(post_pr_new["post_prob_predictions"], z["new_pp"])
    # threshold utilities (BNN_lib.py:627-679) on the same pickle
    thr = bn.get_posterior_threshold(logger._pklfile, target_acc=0.5, post_summary_mode=1)
    assert _close(thr, z["threshold_row"])
    cut = bn.predictBNN(new_dat["data"], pickle_file=logger._pklfile, post_cutoff=0.6, post_summary_mode=1, fname="cut")
    assert np.array_equal(np.isnan(cut["post_prob_predictions"]), np.isnan(z["cut_pp"]))
    assert _close(np.nan_to_num(cut["post_prob_predictions"]), np.nan_to_num(z["cut_pp"]))
    # restart from the pickle (bnn_classify.py:114-134)
    bnn_model = bn.npBNN(dat, n_nodes=n_nodes_list, use_bias_node=1, prior_f=1, p_scale=1, pickle_file=logger._pklfile,
                         seed=rseed, actFun=activation_function)
    for i, w in enumerate(bnn_model._w_layers):
        assert np.array_equal(w, z["restart_w%d" % i])
    # permutation feature importance (bnn_classify.py:137-152; the reference's own call fails under pandas 3, see the
    # generator): runs, writes its table, the irrelevant direction is consistent
    np.random.seed(99)
    fi = bn.feature_importance(dat["test_data"], weights_pkl=logger._pklfile, true_labels=dat["test_labels"],
                               fname_stem=dat["file_name"], feature_names=dat["feature_names"], n_permutations=3,
                               feature_blocks=[[0, 1, 2, 3], [4, 5], [6, 7, 8, 9, 10, 11]], unlink_features_within_block=True)
    assert len(fi) == 3 and os.path.exists(dat["file_name"] + "_feature_importance.txt")


def test_bnn_runner_mc3_script(workdir):
    import np_bnn as bn
    z, meta = G.load("flow_mc3")
    rseed = 1234
    np.random.seed(rseed)
    f, l = workdir["features"], workdir["labels"]
    dat = bn.get_data(f, l, seed=rseed, testsize=0.1, all_class_in_testset=1, header=1, instance_id=1)
    data_obj = bn.npBNN(dat, n_nodes=[5, 5], use_bias_node=-1, seed=1, init_std=0.1)
    logger = bn.postLogger(data_obj, filename="BNNMC3", log_all_weights=0)
    mc3 = bn.MC3(data_obj, logger=logger, n_post_samples=100, sampling_f=100, n_iteration=300, n_chains=4,
                 swap_frequency=20, verbose=1)
    mc3.run_mcmc()
    got = np.array(_rows(logger._logfile)[1:], dtype=np.float64)
    assert got.shape == z["log_rows"].shape and _close(got, z["log_rows"])
    assert _close([a[1]._logPost for a in mc3.singleChainArgs], z["final_logPost"])
    assert np.array_equal(np.array([a[1]._temperature for a in mc3.singleChainArgs]), z["final_temps"])
    post_pr_test = bn.predictBNN(dat["test_data"], pickle_file=logger._pklfile, test_labels=dat["test_labels"],
                                 instance_id=dat["id_test_data"])
    assert np.array_equal(post_pr_test["post_prob_predictions"], z["test_pp"])
    assert post_pr_test["mean_accuracy"] == float(z["test_acc"]) and np.array_equal(post_pr_test["confusion_matrix"], z["test_cm"])
    b2, m2, l2 = bn.load_obj(logger._pklfile)
    assert len(l2._post_weight_samples) == int(z["n_samples"])
    for i, w in enumerate(l2._post_weight_samples[-1]["weights"]):
        assert np.array_equal(w, z["last_sample_w%d" % i])


def test_bnn_regress_script(workdir):
    import np_bnn as bn
    z, _ = G.load("flow_regress")
    np.random.seed(1234)
    dat = bn.get_data(workdir["features_reg"], workdir["labels_reg"], seed=1234, testsize=0.1, all_class_in_testset=0, cv=0,
                      header=True, from_file=True, instance_id=0, randomize_order=True, label_mode="regression")
    bnn_model = bn.npBNN(dat, n_nodes=[6, 4], estimation_mode="regression", actFun=bn.ActFun(fun="tanh"), p_scale=1,
                         use_bias_node=2, empirical_error=True)
    mcmc = bn.MCMC(bnn_model, update_ws=[0.025, 0.025, 0.05], update_f=[0.005, 0.005, 0.05], n_iteration=300, sampling_f=20,
                   print_f=100, n_post_samples=10, likelihood_tempering=1, adapt_f=0.3, estimate_error=False)
    assert np.array_equal(mcmc._update_n, z["update_n"])
    assert _close(mcmc._accuracy_lab_f(mcmc._y, bnn_model._labels), z["label_acc0"])
    logger = bn.postLogger(bnn_model, filename="testM", log_all_weights=0)
    bn.run_mcmc(bnn_model, mcmc, logger)
    assert _close(mcmc._y, z["y"], rtol=1e-10) and _close(mcmc._y_test, z["y_test"], rtol=1e-10)
    got = np.array(_rows(logger._logfile)[1:], dtype=np.float64)
    assert got.shape == z["log_rows"].shape and _close(got, z["log_rows"])
    bnn_obj, mcmc_obj, logger_obj = bn.load_obj(logger._pklfile)
    post_samples = logger_obj._post_weight_samples
    post_weights = [post_samples[i]["weights"] for i in range(len(post_samples))]
    post_alphas = [post_samples[i]["alphas"] for i in range(len(post_samples))]
    actFun, output_act_fun = bnn_obj._act_fun, bnn_obj._output_act_fun
    post_cat_probs = []
    for i in range(len(post_weights)):
        actFun_i = actFun
        actFun_i.reset_prm(post_alphas[i])
        post_cat_probs.append(bn.RunPredict(bnn_obj._data, post_weights[i], actFun=actFun_i, output_act_fun=output_act_fun))
    assert _close(np.array(post_cat_probs), z["post_preds"], rtol=1e-10)
    est = bn.get_posterior_est(logger._pklfile)
    assert _close(est["prm_mean"], z["prm_mean"], rtol=1e-10) and _close(est["prm_mean_test"], z["prm_mean_test"], rtol=1e-10)
    assert _close(np.array(est["error_prm"], dtype=np.float64), z["error_prm"])
    pd_reg = bn.pdp(logger._pklfile, [[0], [1, 2]])
    assert _close(pd_reg[0]["feature"], z["pdp0_feature"]) and _close(pd_reg[0]["pdp"], z["pdp0"], rtol=1e-9)
    assert _close(pd_reg[1]["pdp"], z["pdp1"], rtol=1e-9)


def test_block_bnns_script(workdir):
    """block_bnns.py:13-81: the three masked networks are built and masked like the reference's (known answers of
    create_mask in tests/golden/masks.npz) and a short chain keeps every masked weight at zero."""
    import np_bnn as bn
    z, meta = G.load("masks")
    np.random.seed(1234)
    dat = bn.get_data(workdir["features_reg"], workdir["labels_reg"], seed=1234, testsize=0.1, all_class_in_testset=0, cv=0,
                      header=0, from_file=True, instance_id=0, randomize_order=True, label_mode="regression")
    specs = [([6, 2], [[0, 1, 2], [], []], [[2, 2, 2], [], []], "tanh"),
             ([9, 6], [[0, 1, 2], [0, 0, 0, 1, 1, 1, 2, 2, 2], []], [[3, 3, 3], [2, 2, 2], []], "ReLU"),
             ([9, 5], [[0, 1, 1], [0, 0, 0, 1, 1, 1, 1, 1, 1], []], [[3, 6], [2, 3], []], "ReLU")]
    for si, (nodes, idx, npf, act) in enumerate(specs):
        bnn_model = bn.npBNN(dat, n_nodes=nodes, estimation_mode="regression", actFun=bn.ActFun(fun=act), p_scale=1,
                             use_bias_node=-1, empirical_error=True)
        m = bn.create_mask(bnn_model._w_layers, indx_input_list=idx, nodes_per_feature_list=npf)
        for li in range(3):
            assert np.array_equal(m[li], z["m%d_%d" % (si, li)]), (si, li)
        bnn_model.apply_mask(m)
        mcmc = bn.MCMC(bnn_model, n_iteration=60, sampling_f=20, print_f=1000, estimate_error=False)
        for _ in range(40):
            mcmc.mh_step(bnn_model)
        assert mcmc._current_iteration == 40 and np.isfinite(mcmc._logLik)
        for li in range(3):
            assert np.all(bnn_model._w_layers[li][m[li] == 0] == 0)


def test_single_layer_helpers_run_on_the_device():
    """RunHiddenLayer / MatrixMultiplicationD (BNN_lib.py:154-193) through the prediction kernels against the
    reference's outputs."""
    import np_bnn as bn
    z, _ = G.load("hostlib")
    x, w = z["rh_x"], z["rh_w"]
    assert _close(bn.MatrixMultiplicationD(x, w), z["mmd_bias"], rtol=1e-12)
    assert _close(bn.MatrixMultiplicationD(x, w[:, 1:]), z["mmd_nobias"], rtol=1e-12)
    for fun, prm in (("ReLU", None), ("genReLU", [0.1, 0.3]), ("swish", None), ("tanh", None)):
        af = bn.ActFun(fun=fun, prm=np.array(prm) if prm else np.zeros(1))
        assert _close(bn.RunHiddenLayer(x, w, af, 1 if prm else 0), z["rh_" + fun], rtol=1e-12), fun
    assert _close(bn.RunHiddenLayer(x, w, False, 2), z["rh_none"], rtol=1e-12)


def test_labels_outside_the_class_range_are_rejected():
    """labels index shared-memory counters inside the epilogues; bnn_set_data refuses labels outside [0, K)
    (1-based labels, or init_weights with fewer outputs) instead of corrupting memory -- the reference raises IndexError."""
    from npbnn_b200 import _lib as L
    from npbnn_b200.engine import Engine, NetShape
    rng = np.random.default_rng(0)
    x = rng.standard_normal((100, 6))
    w = [rng.normal(size=(4, 6)), rng.normal(size=(3, 4)), rng.normal(size=(3, 3))]
    eng = Engine(NetShape.from_weights(w, 6, act="tanh", lik=L.LIK_CATEGORICAL))
    with pytest.raises(L.NpbnnError):
        eng.set_data(x, rng.integers(1, 4, 100))              # 1-based labels with K = 3
    with pytest.raises(L.NpbnnError):
        eng.set_data(x, rng.integers(0, 3, 100), x[:10], np.full(10, -1))
    eng.set_data(x, rng.integers(0, 3, 100))
    assert np.isfinite(eng.forward_lik([w])["loglik"][0])
    eng.close()


def test_philox_streams_follow_the_global_chain_index():
    """with chains sharded over ranks every rank used key seed ^ LOCAL index, i.e. identical proposal
    streams.  The key is now seed ^ (chain_offset + c): a shard holding global chains 2..3 reproduces chains 2..3 of
    the unsharded run bit for bit, and differs from the shard holding chains 0..1."""
    from npbnn_b200 import _lib as L
    from npbnn_b200.engine import Engine, NetShape
    d = golden_data.synth_class(500, 6, 3, 3)
    rs = np.random.RandomState(1)
    w0 = [rs.normal(0, 0.1, s) for s in ((5, 6), (4, 5), (3, 4))]
    net = NetShape.from_weights(w0, 6, act="tanh", lik=L.LIK_CATEGORICAL)

    def run(n, offset):
        eng = Engine(net)
        eng.set_data(d["data"], d["labels"])
        eng.chains_init([w0] * n, seed=77, chain_offset=offset)
        eng.mh_steps(25)
        st = eng.read_state()
        eng.close()
        return st
    full, lo, hi = run(4, 0), run(2, 0), run(2, 2)
    assert np.array_equal(hi.w, full.w[2:]) and np.array_equal(lo.w, full.w[:2])
    assert np.array_equal(hi.logLik, full.logLik[2:])
    assert not np.array_equal(hi.w[0], lo.w[0])
