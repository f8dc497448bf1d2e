#!/usr/bin/env python
"""run_mcmc on the c1 shape (2250+250 rows, 128 features, [5,5] tanh) with device-generated proposals and
sampling_f=10: synchronous logging points vs the asynchronous snapshot ring.
usage: tools/logger_bench.py [iterations] [out.json]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import npbnn_b200 as bn

n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(0)
x = rng.standard_normal((2500, 128))
y = rng.integers(0, 5, 2500)
dat = {"data": x[:2250], "labels": y[:2250], "test_data": x[2250:], "test_labels": y[2250:]}
res = {}
with tempfile.TemporaryDirectory() as d:
    for tag, depth in (("synchronous", 0), ("ring_depth4", 4), ("ring_depth16", 16)):
        np.random.seed(1)
        bnn = bn.npBNN(dat, n_nodes=[5, 5], actFun=bn.ActFun(fun="tanh"), use_bias_node=2, seed=1)
        mcmc = bn.MCMC(bnn, n_iteration=n_it, sampling_f=10, print_f=10 ** 9, n_post_samples=100, rng="philox")
        logger = bn.postLogger(bnn, filename="lb_" + tag, wdir=d)
        t0 = time.perf_counter()
        bn.run_mcmc(bnn, mcmc, logger, pipeline_depth=depth)
        dt = time.perf_counter() - t0
        res[tag] = {"seconds": dt, "it_per_s": n_it / dt, "logLik": mcmc._logLik}
        print(tag, res[tag], flush=True)
assert len({round(v["logLik"], 9) for v in res.values()}) == 1
if len(sys.argv) > 2:
    json.dump({"config": "c1 shape, rng=philox, sampling_f=10, %d iterations" % n_it, "results": res}, open(sys.argv[2], "w"), indent=1)
