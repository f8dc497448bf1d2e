#!/usr/bin/env python
"""Where does the persistent chain loop stop paying?  c1-shaped network ([5,5] tanh, 128 features) at growing row counts
and chain counts; option chain_loop = 2 (loop wherever it fits), 1 (automatic: the library's cost model), 0 (launch
sequence): us per MH step (device-generated proposals, 100-step calls).  Writes profiles/r02_chain_loop_threshold.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from npbnn_b200.engine import Engine, NetShape

shapes = [(5, 129), (5, 6), (5, 5)]
out = []
for n in (500, 1000, 2500, 4000, 5000, 10000, 30000):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, 128)); y = rng.integers(0, 5, n)
    for chains in (1, 2, 4, 8, 16, 32):
        sets = [[rng.normal(0, 0.1, s) for s in shapes] for _ in range(chains)]
        rec = {"rows": n, "chains": chains}
        for loop in (2, 1, 0):
            eng = Engine(NetShape(128, shapes, act="tanh", lik=0))
            eng.set_data(x, y)
            eng.chains_init(sets, seed=3)
            eng.set_option("chain_loop", loop)
            for _ in range(2):
                eng.mh_steps(100)
            eng.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                eng.mh_steps(100)
            eng.synchronize()
            key = {2: "loop", 1: "auto", 0: "sequence"}[loop]
            rec[key + "_us"] = (time.perf_counter() - t0) / 500 * 1e6
            rec["kernel_" + key] = eng.last_kernel
            eng.close()
        out.append(rec)
        print(json.dumps(rec), flush=True)
json.dump(out, open(os.path.join("gpurun_out" if os.path.isdir("gpurun_out") else "profiles", "r02_chain_loop_threshold.json"), "w"), indent=1)
