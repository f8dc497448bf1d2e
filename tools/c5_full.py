#!/usr/bin/env python
"""BASELINE config 5 at full size on ONE B200: 10,000 posterior samples x 1,000,000 rows, [64,32] swish, 10 classes.
  (a) posterior mean + vote summaries (get_posterior_cat_prob modes 0 / 1: bnn_predict),
  (b) posterior-predictive resampling (mode 2, sample_from_categorical) with in-kernel Philox uniforms (bnn_predict_sample
      with u = None): O(N K) memory -- the reference's [N, S] uniform / prediction arrays would be 80 GB each.
Writes profiles/r02_c5_full_10k_samples.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from npbnn_b200 import _lib as L, workloads as wl
from npbnn_b200.engine import Engine, NetShape, flatten_weights

S = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
n = 1_000_000
x, _ = wl.c4_data(n, seed=0)
rng = np.random.default_rng(5)
base = flatten_weights(wl.c4_init_weights(1)[0])
w = base[None, :] + rng.normal(0, 0.05, (S, base.size))
eng = Engine(NetShape(64, list(wl.C4_SHAPES), act="swish", lik=L.LIK_CATEGORICAL), device=0)
xd = torch.from_numpy(x).cuda()
wd = torch.from_numpy(w).cuda()
out = {"rows": n, "samples": S, "gpu": torch.cuda.get_device_name(0)}
torch.cuda.reset_peak_memory_stats()
for name, fn in (("summaries_mean_votes", lambda: eng.predict(xd, wd, mean=True, votes=True)),
                 ("resampling_mode2_philox", lambda: eng.predict_sample(xd, wd, u=None, post_predictions=False, seed=11))):
    fn() if S <= 256 else None                       # (warm-up only for short runs; the long ones are 5 s each)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rec = {"seconds": dt, "row_samples_per_s": S * n / dt, "tflops": S * n * wl.C4_FLOP_PER_ROW / dt / 1e12, "kernel": eng.last_kernel}
    if name.startswith("summaries"):
        rec["mean_prob_sum_per_row"] = float(np.asarray(r["mean"]).sum() / n)
    else:
        est = np.asarray(r["predictions"])
        rec["share_sum_per_row"] = float(est.sum() / n)
        rec["class_counts_total"] = int(np.asarray(r["class_counts"]).sum())
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
out["peak_device_memory_GB"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(out))
dst = os.path.join("gpurun_out" if os.path.isdir("gpurun_out") else "profiles", "r02_c5_full_10k_samples.json")
json.dump(out, open(dst, "w"), indent=1)
