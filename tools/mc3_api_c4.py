#!/usr/bin/env python
"""BASELINE config 4 through the reference-facing Python surface (`import np_bnn as bn`): the flow of
bnn_runner_MC3.py:17-48 (npBNN -> postLogger -> MC3 -> run_mcmc) on 1,000,000 x 64 rows, [64,32] swish, 10 classes,
bias on the last layer, 32 tempered chains, swap every 100 iterations -- what a user of the reference gets after changing
nothing but the package on the path.  Both proposal sources: rng="host" (the reference's numpy generator replayed, the
default: bit-compatible chains) and rng="philox" (device-generated).  bench.py measures the same workload through the
C ABI; this is the rate of the Python surface above it, logger and posterior samples included.

    python tools/mc3_api_c4.py OUT.json [n_iteration]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import np_bnn as bn
from npbnn_b200 import workloads as wl

out_path = sys.argv[1]
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 300
t0 = time.perf_counter()
x, y = wl.c4_data(1_000_000, seed=0)
t_data = time.perf_counter() - t0
dat = {"data": x, "labels": y.astype(np.int64), "label_dict": np.arange(10), "test_data": x[:1000],
       "test_labels": y[:1000].astype(np.int64)}
res = {"what": __doc__.split("\n\n")[0].replace("\n", " "), "rows": len(x), "chains": 32, "n_iteration": n_it,
       "host_data_seconds": t_data, "runs": []}
for rng in ("host", "philox"):
    with tempfile.TemporaryDirectory() as tmp:
        np.random.seed(1234)
        t0 = time.perf_counter()
        bnn = bn.npBNN(dat, n_nodes=[64, 32], use_bias_node=-1, seed=1, actFun=bn.ActFun(fun="swish"))
        logger = bn.postLogger(bnn, filename="c4", wdir=tmp)
        mc3 = bn.MC3(bnn, logger=logger, n_post_samples=10, sampling_f=100, n_iteration=n_it, n_chains=32,
                     swap_frequency=100, verbose=0, print_f=10 ** 9, rng=rng, swap_seed=77)
        torch.cuda.synchronize()
        t_setup = time.perf_counter() - t0
        t0 = time.perf_counter()
        mc3.run_mcmc()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rows = open(logger._logfile).read().splitlines()
        res["runs"].append({"rng": rng, "setup_seconds": t_setup, "run_seconds": dt,
                            "chain_steps_per_s": 32 * n_it / dt, "log_rows": len(rows), "last_log_row": rows[-1][:160]})
        print(res["runs"][-1], flush=True)
json.dump(res, open(out_path, "w"), indent=1)
