#!/usr/bin/env python
"""it/s of run_mcmc through the reference-facing Python API on the c1 shape (2250+250 rows, 128 features, [5,5] tanh):
rng="host" (the reference's generator sequence replayed on the device, default) vs rng="philox" (device-generated).
usage: tools/api_rate.py [iterations] [sampling_f] [out.json]"""
import cProfile, json, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import npbnn_b200 as bn

n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
sf = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rng = np.random.default_rng(0)
x = rng.standard_normal((2500, 128))
y = rng.integers(0, 5, 2500)
dat = {"data": x[:2250], "labels": y[:2250], "test_data": x[2250:], "test_labels": y[2250:]}
res = {}
with tempfile.TemporaryDirectory() as d:
    for mode in ("host", "philox"):
        np.random.seed(1)
        bnn = bn.npBNN(dat, n_nodes=[5, 5], actFun=bn.ActFun(fun="tanh"), use_bias_node=2, seed=1)
        mcmc = bn.MCMC(bnn, n_iteration=n_it, sampling_f=sf, print_f=10 ** 9, n_post_samples=100, rng=mode,
                       adapt_f=0.3, adapt_fM=0.6)
        logger = bn.postLogger(bnn, filename="ar_" + mode, wdir=d)
        pr = cProfile.Profile() if mode == "host" and os.environ.get("PROFILE") else None
        t0 = time.perf_counter()
        if pr:
            pr.enable()
        bn.run_mcmc(bnn, mcmc, logger)
        if pr:
            pr.disable()
            pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
        dt = time.perf_counter() - t0
        res[mode] = {"seconds": dt, "it_per_s": n_it / dt, "logLik": mcmc._logLik}
        print(mode, res[mode], flush=True)
if len(sys.argv) > 3:
    json.dump({"config": "c1 shape, run_mcmc, sampling_f=%d, %d iterations" % (sf, n_it), "results": res}, open(sys.argv[3], "w"), indent=1)
