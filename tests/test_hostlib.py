"""CPU: the host-side callers around the device path (npbnn_b200/hostlib.py) against outputs of the unmodified
reference (tests/golden/hostlib.npz, written by tests/golden/make_golden.py: case_hostlib), and the rank-0 logging /
seed-broadcast logic of the sharded MC3 driver on two gloo ranks."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import _golden as G
from tests import golden_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return G.load("hostlib")


@pytest.fixture(scope="module")
def tables(tmp_path_factory):
    d = tmp_path_factory.mktemp("tables")
    return golden_data.write_example_tables(str(d))


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind in "OUS" or b.dtype.kind in "OUS":
        return a.shape == b.shape and bool(np.all(a.astype(str) == b.astype(str)))
    return a.shape == b.shape and bool(np.array_equal(a, b, equal_nan=True))


def test_np_bnn_alias_exports_every_public_callable_of_the_reference():
    """The 88 public callables the reference package exposes through its star imports (np_bnn/__init__.py:6-25),
    listed here so that the test does not need /root/reference."""
    import np_bnn as bn
    names = """ActFun CalcAccAboveThreshold CalcAccuracy CalcAccuracyRegression CalcConfusionMatrix CalcFP CalcFP_BF
    CalcLabelAccuracy CalcLabelAccuracyRegression CalcLabelFreq CalcTP CalcTP_BF GibbsSampleGammaRateExp
    GibbsSampleNormStdGamma2D GibbsSampleNormStdGammaONE GibbsSampleNormStdGammaVector MC3 MCMC MatrixMultiplication
    MatrixMultiplicationD RecurMeanVar RegressTransform RegressTransformError RunHiddenLayer RunPredict RunPredictInd
    SaveObject SkipAccuracy SkipAccuracyVec SoftMax SoftPlus UpdateBinomial UpdateFixedNormal UpdateNormal UpdateNormal1D
    UpdateNormalNormalized UpdateUniform assign_indx calcHPD calc_likelihood calc_likelihood_regression
    calc_likelihood_regression_error combine_pkls create_mask data_transform_obj deepcopy feature_importance gamma_acc
    gamma_likelihood get_accuracy_threshold get_data get_feature_summary get_pdp get_posterior_cat_prob get_posterior_est
    get_posterior_threshold get_weights_from_tensorflow_model init_output_files init_weight_prm leaky_relu_f load_obj
    make_pdp_features merge_dict multiplier_proposal multiplier_proposal_vector negbin2d_acc negbin_acc negbin_acc_base10
    negbin_likelihood negbin_likelihood2d negbin_likelihood_base10 npBNN pdp poi_acc poi_likelihood postLogger predict
    predictBNN randomize_data relu_f run_mcmc sample_from_categorical save_data swish_f tanh_f turn_labels_to_numeric
    turn_low_pp_instances_to_nan unique_unsorted""".split()
    assert len(names) == 88
    missing = [n for n in names if not callable(getattr(bn, n, None))]
    assert missing == []


def test_get_data_reproduces_the_reference_splits(gold, tables):
    """get_data / randomize_data / turn_labels_to_numeric (BNN_files.py:10-99,190-260,302-313): same rows in the same
    order, same label coding, same instance names for the calls the five scripts make."""
    import np_bnn as bn
    z, meta = gold
    for name, (f, l, kw) in meta["get_data_cases"].items():
        dat = bn.get_data(tables[f], tables[l] if l else None, **kw)
        for key in ("data", "labels", "label_dict", "test_data", "test_labels", "id_data", "id_test_data", "file_name",
                    "feature_names"):
            assert _same(dat[key], z["gd_%s_%s" % (name, key)]), (name, key)


def test_prediction_table_summaries(gold):
    import np_bnn as bn
    z, _ = gold
    probs, prior, lab = z["probs"], z["prior_probs"], z["lab"]
    assert bn.CalcAccuracy(probs, lab) == float(z["CalcAccuracy"])
    assert _same(bn.CalcAccuracy(np.stack([probs, prior]), lab), z["CalcAccuracy3"])
    assert _same(bn.CalcLabelAccuracy(probs, lab), z["CalcLabelAccuracy"])
    assert _same(bn.CalcLabelFreq(probs), z["CalcLabelFreq"])
    cm = bn.CalcConfusionMatrix(probs, lab)
    assert _same(cm.values, z["CalcConfusionMatrix"]) and cm.index.name == "True" and cm.columns.name == "Predicted"
    assert str(cm.index[-1]) == "All"
    assert bn.CalcTP(probs, lab, threshold=0.8) == float(z["CalcTP"]) and bn.CalcFP(probs, lab, threshold=0.8) == float(z["CalcFP"])
    assert bn.CalcTP_BF(probs, prior, lab, threshold=20) == float(z["CalcTP_BF"])
    assert bn.CalcFP_BF(probs, prior, lab, threshold=20) == float(z["CalcFP_BF"])
    th = bn.get_accuracy_threshold(probs, lab, threshold=0.75)
    assert _same(th["predictions"], z["thr_predictions"]) and th["accuracy"] == float(z["thr_accuracy"])
    assert th["retained_samples"] == float(z["thr_retained"]) and _same(th["confusion_matrix"].values, z["thr_cm"])
    keep = np.where(np.max(probs, axis=1) > 0.9)[0]
    assert _same(bn.turn_low_pp_instances_to_nan(probs, keep), z["low_pp"])
    assert _same(np.array(bn.calcHPD(z["z"][:, 0], 0.9)), z["hpd"])
    assert np.isclose(bn.CalcAccuracyRegression(z["yreg"], z["labreg"]), float(z["CalcAccuracyRegression"]), rtol=1e-15)
    assert np.allclose(bn.CalcLabelAccuracyRegression(z["yreg"], z["labreg"]), z["CalcLabelAccuracyRegression"], rtol=1e-15)
    assert _same(bn.assign_indx(["b", "a", "b", "c", "a"]), z["assign_indx"])
    assert _same(bn.unique_unsorted(np.array([3, 1, 3, 2, 1])), z["unique_unsorted"])
    assert _same(bn.get_feature_summary(z["xs"], [0, 1]), z["feature_summary"])
    mu, var = bn.RecurMeanVar(4, [np.zeros((7, 6)), np.ones((7, 6))], z["w"], (np.array([0, 2]), np.array([1, 3])))
    assert _same(mu, z["recur_mu"]) and _same(var, z["recur_var"])


def test_elementwise_forms_and_host_likelihoods(gold):
    """The selector tokens evaluate the reference's formulas when called on a host table (post-processing use)."""
    import np_bnn as bn
    z, _ = gold
    zz = z["z"]
    assert _same(bn.relu_f(zz.copy(), 0), z["relu"]) and _same(bn.leaky_relu_f(zz.copy(), 0.2), z["leaky"])
    assert _same(bn.swish_f(zz.copy(), 0), z["swish"]) and _same(bn.tanh_f(zz.copy(), 0), z["tanh"])
    assert np.allclose(bn.SoftMax(zz), z["probs"], rtol=1e-14, atol=0)
    assert _same(bn.SoftPlus(zz), z["softplus"]) and _same(bn.RegressTransformError(zz.copy()), z["regerr"])
    assert _same(bn.RegressTransform(zz), zz)
    n, k = z["probs"].shape
    lik = bn.calc_likelihood(z["probs"], z["lab"], np.arange(n), class_weight=np.linspace(0.5, 1.5, k), lik_temp=0.7)
    assert np.isclose(lik, float(z["lik_cat"]), rtol=1e-13)
    lik = bn.calc_likelihood(z["probs"], z["lab"], np.arange(n), instance_weight=np.linspace(0.1, 2, n))
    assert np.isclose(lik, float(z["lik_cat_iw"]), rtol=1e-13)
    lik = bn.calc_likelihood_regression(z["yreg"][:, :2], z["labreg"], None, lik_temp=0.9, sig2=np.array([0.5, 2.0]))
    assert np.isclose(lik, float(z["lik_reg"]), rtol=1e-13)
    lik = bn.calc_likelihood_regression_error(bn.RegressTransformError(z["yreg"].copy()), z["labreg"], None)
    assert np.isclose(lik, float(z["lik_regerr"]), rtol=1e-13)
    with pytest.raises(SystemExit):
        bn.calc_likelihood_regression(z["yreg"][:, :2], z["labreg"], None, instance_weight=np.ones(50))
    cnt, y = z["cnt"], z["yreg"]
    for name, got in (("poi", bn.poi_likelihood(y, cnt)), ("negbin", bn.negbin_likelihood(y, cnt)),
                      ("negbin2d", bn.negbin_likelihood2d(y, cnt)), ("negbin10", bn.negbin_likelihood_base10(y * 0.3, cnt)),
                      ("gamma", bn.gamma_likelihood(y * 0.1, cnt[:, :1] + 3.0)), ("negbin_acc", bn.negbin_acc(y, cnt)),
                      ("negbin2d_acc", bn.negbin2d_acc(y, cnt)), ("poi_acc", bn.poi_acc(y, cnt)),
                      ("negbin_acc10", bn.negbin_acc_base10(y * 0.3, cnt))):
        assert np.isclose(got, float(z[name]), rtol=1e-12), name


def test_proposal_and_gibbs_helpers_consume_the_generators_like_the_reference(gold):
    import np_bnn as bn
    z, _ = gold
    w = z["w"]
    d = np.full(w.shape, 0.3)
    for name in ("UpdateNormal", "UpdateFixedNormal", "UpdateNormalNormalized"):
        zz, (ix, iy), h = getattr(bn, name)(w, d=d, n=9, Mb=1.0, mb=-1.0, rs=np.random.default_rng(5))
        assert _same(zz, z[name + "_z"]) and _same(ix, z[name + "_ix"]) and _same(iy, z[name + "_iy"]), name
        assert np.isclose(h, float(z[name + "_h"]), rtol=1e-13, atol=0), name
    zz, ix, h = bn.UpdateNormal1D(w[0], d=0.05, n=2, rs=np.random.default_rng(5))
    assert _same(zz, z["UpdateNormal1D_z"]) and _same(ix, z["UpdateNormal1D_ix"])
    np.random.seed(31)                      # one global stream for the next six calls, as in the generator
    assert _same(bn.UpdateUniform(w, d=d, n=4)[0], z["UpdateUniform_z"])
    assert _same(bn.UpdateBinomial(np.ones((3, 4)), 0.5, (3, 4)), z["UpdateBinomial"])
    q, _, u = bn.multiplier_proposal_vector(np.array([1.0, 2.0, 3.0]), d=1.2, f=0.6, rs=np.random.default_rng(5))
    assert _same(q, z["mpv_q"]) and np.isclose(u, float(z["mpv_u"]), rtol=1e-14)
    assert np.allclose(np.array(bn.multiplier_proposal(2.0, d=1.1)[::2]), z["mp"], rtol=1e-15)
    assert bn.GibbsSampleNormStdGammaVector(w.flatten()) == float(z["gibbs_vec"])
    assert _same(bn.GibbsSampleNormStdGamma2D(w), z["gibbs_2d"]) and _same(bn.GibbsSampleNormStdGammaONE(w), z["gibbs_one"])
    assert bn.GibbsSampleGammaRateExp(np.array([0.5, 1.0, 2.0]), 2.0) == float(z["gibbs_rate"])


def test_init_output_files_and_logger_header(tmp_path):
    import np_bnn as bn
    dat = golden_data.synth_class(60, 4, 3, 1, 10)
    bnn = bn.npBNN(dat, n_nodes=[4, 3], hyper_p=2, freq_indicator=0.1)
    log, wlog, pkl = bn.init_output_files(bnn, filename=str(tmp_path / "sub" / "run"), log_all_weights=1)
    head = open(log).read().strip().split("\t")
    assert head[:6] == ["it", "posterior", "likelihood", "prior", "accuracy", "test_accuracy"]
    assert head[6:9] == ["acc_C0", "acc_C1", "acc_C2"] and "mean_prior_std_w2" in head and "mean_ind" in head
    assert head[-2:] == ["acc_prob", "mcmc_id"]
    assert log.endswith("run_l4_3.log") and wlog.endswith("run_l4_3_W.log") and pkl.endswith("run_l4_3.pkl")
    assert len(open(wlog).read().strip().split("\t")) == 1 + bnn._n_params
    logger = bn.postLogger(bnn, filename=str(tmp_path / "lg"), log_all_weights=0)
    assert open(logger._logfile).read().strip().split("\t") == head        # one header builder for both entry points


def test_update_data_after_staging_is_refused_not_ignored():
    """npBNN.update_data / reset_weights after an MCMC staged the model would be silently ignored by the device state
    (ADVICE r1): the sampler raises instead.  Checked on the version bookkeeping (no device needed)."""
    import np_bnn as bn
    dat = golden_data.synth_class(60, 4, 3, 1, 10)
    bnn = bn.npBNN(dat, n_nodes=[4, 3])
    v0 = bnn.__dict__.get("_data_version", 0)
    bnn.update_data(dat)
    bnn.reset_weights(bnn._w_layers)
    assert bnn._data_version == v0 + 2


RANK0_LOGGING_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
rank, port, out = int(sys.argv[1]), sys.argv[2], sys.argv[3]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
dist.init_process_group("gloo", rank=rank, world_size=2)
np.random.seed(100 + rank)                       # deliberately different numpy states on the two ranks
import np_bnn as bn
from npbnn_b200 import api, mc3
from tests import golden_data
dat = golden_data.synth_class(60, 4, 3, 1, 10)
bnn = bn.npBNN(dat, n_nodes=[4, 3], init_weights=[np.full(s, 0.5) for s in ((4, 5), (3, 5), (3, 3))])
logger = bn.postLogger(bnn, filename="sharded", wdir=out)
seeds = mc3.broadcast_from_rank0(np.random.choice(range(1000, 9999), 4, replace=False))
m = object.__new__(api.MC3)                      # the logging step of MC3.run_mcmc without a device
m.n_chains, m.world, m.rank, m.logger, m._bnn = 4, 2, rank, logger, bnn
m.current_temperatures = np.array([0.8, 0.9, 1.0, 0.95])          # the cold chain (index 2) lives on rank 1
cold = None
if rank == 1:
    b = api.deepcopy(bnn)
    b._w_layers = [w + 1.0 for w in b._w_layers]
    mc = object.__new__(api.MCMC)
    mc.__dict__.update(_current_iteration=40, _logPost=-12.5, _logLik=-10.0, _logPrior=-2.5, _accuracy=0.75,
                       _test_accuracy=0.5, _label_acc=np.array([0.1, 0.2, 0.3]), _acceptance_rate=0.25, _mcmc_id=2,
                       _n_post_samples=5, _temperature=1.0)
    cold = (b, mc)
m._log_on_rank0(cold)
np.save(out + "/seeds_%%d.npy" %% rank, seeds)
np.save(out + "/nsamples_%%d.npy" %% rank, np.array([len(logger._post_weight_samples)]))
dist.destroy_process_group()
'''


def test_sharded_mc3_logs_on_rank0_only_two_ranks_gloo(tmp_path):
    """Two gloo ranks: the cold chain sits on rank 1; its exported state reaches rank 0, which alone writes the .log /
    .pkl and keeps the sample list; chain seeds are rank 0's on both ranks (logging and seeds per rank)."""
    script = tmp_path / "w.py"
    script.write_text(RANK0_LOGGING_WORKER % {"root": ROOT})
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), port, str(tmp_path)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    assert np.array_equal(np.load(tmp_path / "seeds_0.npy"), np.load(tmp_path / "seeds_1.npy"))
    assert int(np.load(tmp_path / "nsamples_0.npy")[0]) == 1 and int(np.load(tmp_path / "nsamples_1.npy")[0]) == 0
    rows = open(tmp_path / "sharded_l4_3.log").read().strip().split("\n")
    assert len(rows) == 2                                    # one header (rank 0 only), one sample
    vals = rows[1].split("\t")
    assert vals[0] == "40" and float(vals[1]) == -12.5 and vals[-1] == "2"
    import pickle
    b, mc, lg = pickle.load(open(tmp_path / "sharded_l4_3.pkl", "rb"))
    assert np.all(lg._post_weight_samples[0]["weights"][0] == 1.5) and mc._current_iteration == 40
