#!/usr/bin/env python
"""BASELINE config 5 shape through the reference-facing Python surface: bn.get_posterior_cat_prob (BNN_lib.py:352-397) on
1,000,000 host rows with S posterior samples given as the reference's list of {"weights", "alphas"} dicts -- posterior mean
(mode 1), votes (mode 0) and categorical resampling (mode 2) without the [S, N, K] tensor (return_dense=False).  The rate
includes the host copy of the features, their upload and packing, the packing of the samples and the read-back.

    python tools/predict_api_c5.py OUT.json [S]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import np_bnn as bn
from npbnn_b200 import workloads as wl

out_path = sys.argv[1]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
n = 1_000_000
x, y = wl.c4_data(n, seed=0)
rng = np.random.default_rng(5)
base = wl.c4_init_weights(1)[0]
post = [{"weights": [w + rng.normal(0, 0.05, w.shape) for w in base], "alphas": [0.0]} for _ in range(S)]
af = bn.ActFun(fun="swish")
res = {"what": __doc__.split("\n\n")[0].replace("\n", " "), "rows": n, "samples": S, "runs": []}
bn.get_posterior_cat_prob(x[:4096], post[:8], post_summary_mode=1, actFun=af, output_act_fun=bn.SoftMax, return_dense=False)
for mode in (1, 0, 2):
    np.random.seed(7)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, summ = bn.get_posterior_cat_prob(x, post, post_summary_mode=mode, actFun=af, output_act_fun=bn.SoftMax,
                                        return_dense=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res["runs"].append({"post_summary_mode": mode, "seconds": dt, "row_samples_per_s": n * S / dt,
                        "tflops": n * S * wl.C4_FLOP_PER_ROW / dt / 1e12, "summary_shape": list(np.shape(summ)),
                        "row_sums_ok": bool(np.allclose(np.sum(summ, axis=1), 1.0))})
    print(res["runs"][-1], flush=True)
json.dump(res, open(out_path, "w"), indent=1)
