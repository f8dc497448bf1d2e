#!/usr/bin/env python
"""Opt-in 3xTF32 prediction (option predict_tf32) against the FP64 kernel on the config-5 shape: time and agreement."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from npbnn_b200 import _lib as L, workloads as wl
from npbnn_b200.engine import Engine, NetShape, flatten_weights
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = 1_000_000
x, _ = wl.c4_data(n, seed=0)
rng = np.random.default_rng(5)
base = flatten_weights(wl.c4_init_weights(1)[0])
w = base[None, :] + rng.normal(0, 0.05, (S, base.size))
eng = Engine(NetShape(64, list(wl.C4_SHAPES), act="swish", lik=L.LIK_CATEGORICAL), device=0)
xd, wd = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
out = {"rows": n, "samples": S}
res = {}
for mode in (0, 1, 2):
    eng.set_option("predict_tf32", mode)
    eng.predict(xd, wd[:32], mean=True, votes=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = eng.predict(xd, wd, mean=True, votes=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[mode] = r
    out[("fp64", "tf32x3", "tf32x1")[mode]] = {"seconds": dt, "row_samples_per_s": S * n / dt, "kernel": eng.last_kernel}
out["max_abs_diff_mean_prob"] = float(np.abs(res[0]["mean"] - res[1]["mean"]).max())
out["vote_flips_per_million"] = float(np.abs(np.rint(res[0]["votes"] * S) - np.rint(res[1]["votes"] * S)).sum() / 2 / (n * S) * 1e6)
out["speedup"] = out["fp64"]["seconds"] / out["tf32x3"]["seconds"]
out["max_abs_diff_mean_prob_tf32x1"] = float(np.abs(res[0]["mean"] - res[2]["mean"]).max())
out["speedup_tf32x1"] = out["fp64"]["seconds"] / out["tf32x1"]["seconds"]
print(json.dumps(out))
json.dump(out, open(os.path.join("gpurun_out" if os.path.isdir("gpurun_out") else "profiles", "r02_predict_tf32x3.json"), "w"), indent=1)
