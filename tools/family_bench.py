#!/usr/bin/env python
"""Kernel timing of the k_fwd3 instantiation families against the generic kernel on 1M rows:
[64,32] swish / tanh / ReLU / genReLU categorical, ReLU regression (+ sigma head), a 32-32-16 member and a padded-up
network.  For every case: ms per launch of the specialised kernel, of the generic kernel on the same data (option
force_generic), algorithmic TFLOP/s and the ratio to the swish headline.  Writes profiles/r02_generic_vs_fwd3.json.
    python tools/family_bench.py [rows] [sets]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch  # noqa: E401,E402
from npbnn_b200 import _lib as L  # noqa: E402
from npbnn_b200.engine import Engine, NetShape  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sets = int(sys.argv[2]) if len(sys.argv) > 2 else 32

CASES = [  # name, F, hidden, K / outputs, act, lik, bias
    ("c4 [64,32] swish categorical K=10", 64, (64, 32), 10, "swish", L.LIK_CATEGORICAL, -1),
    ("[64,32] tanh categorical K=10", 64, (64, 32), 10, "tanh", L.LIK_CATEGORICAL, -1),
    ("[64,32] ReLU categorical K=10", 64, (64, 32), 10, "ReLU", L.LIK_CATEGORICAL, -1),
    ("[64,32] genReLU categorical K=10", 64, (64, 32), 10, "genReLU", L.LIK_CATEGORICAL, -1),
    ("[64,32] ReLU regression O=2", 64, (64, 32), 2, "ReLU", L.LIK_GAUSSIAN, 2),
    ("[64,32] tanh regression-error O=2x2", 64, (64, 32), 4, "tanh", L.LIK_GAUSSIAN_HEAD, 3),
    ("[32,16] swish categorical K=8, F=32", 32, (32, 16), 8, "swish", L.LIK_CATEGORICAL, 2),
    ("[50,20] tanh categorical K=7, F=40 (padded up to 64-64-32)", 40, (50, 20), 7, "tanh", L.LIK_CATEGORICAL, 2),
]


def shapes(f, hidden, out, bias):
    b1, b2, b3 = int(bias >= 1), int(bias >= 2), int(bias in (3, -1))
    return [(hidden[0], f + b1), (hidden[1], hidden[0] + b2), (out, hidden[1] + b3)]


def time_pass(eng, w, al, reps=4):
    for _ in range(2):
        eng.forward_lik(w, alphas=al)
    eng.set_option("time_forward", 1)
    eng.forward_time(True)
    for _ in range(reps):
        r = eng.forward_lik(w, alphas=al)
    ms, n = eng.forward_time(True)
    eng.set_option("time_forward", 0)
    return ms / n, eng.last_kernel, r["loglik"]


out = []
g = torch.Generator(device="cuda").manual_seed(0)
for name, f, hidden, k, act, lik, bias in CASES:
    shp = shapes(f, hidden, k, bias)
    x = torch.randn(rows, f, dtype=torch.float64, device="cuda", generator=g)
    if lik == L.LIK_CATEGORICAL:
        y = torch.randint(0, k, (rows,), dtype=torch.int32, device="cuda", generator=g)
    else:
        y = torch.randn(rows, k if lik == L.LIK_GAUSSIAN else k // 2, dtype=torch.float64, device="cuda", generator=g)
    eng = Engine(NetShape(f, shp, act=act, lik=lik))
    eng.set_data(x, y)
    rs = np.random.RandomState(1)
    w = torch.as_tensor(np.stack([np.concatenate([rs.normal(0, 0.1, s).ravel() for s in shp]) for _ in range(sets)])).cuda()
    al = np.tile([0.05, 0.3, 0.0], (sets, 1)) if act == "genReLU" else None
    fast_ms, fast_k, ll_f = time_pass(eng, w, al)
    eng.set_option("force_generic", 1)
    gen_ms, gen_k, ll_g = time_pass(eng, w, al, reps=2)
    eng.close()
    flop_row = sum(2 * (r * (c - b)) + b * r for (r, c), b in zip(shp, [int(bias >= 1), int(bias >= 2), int(bias in (3, -1))]))
    rec = {"case": name, "rows": rows, "sets": sets, "kernel": fast_k, "ms": fast_ms, "generic_ms": gen_ms,
           "speedup_vs_generic": gen_ms / fast_ms, "algorithmic_flop_per_row": flop_row,
           "tflops": sets * rows * flop_row / (fast_ms * 1e-3) / 1e12,
           "max_rel_diff_vs_generic": float(np.max(np.abs(ll_f - ll_g) / np.abs(ll_g)))}
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del x, y, w
    torch.cuda.empty_cache()
base = out[0]["ms"]
for r in out:
    r["ms_relative_to_swish_headline"] = r["ms"] / base
os.makedirs("profiles", exist_ok=True)
dst = os.path.join("gpurun_out" if os.path.isdir("gpurun_out") else "profiles", "r02_generic_vs_fwd3.json")
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst)
