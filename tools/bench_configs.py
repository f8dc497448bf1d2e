#!/usr/bin/env python
"""Device timings of the BASELINE.json configurations other than the bench.py headline (c4):

  c1  classification  N=2250+250, F=128, [5,5] tanh, K=5, bias mode 2      (bnn_classify.py shapes)
  c2  regression      N=900+99,   F=3,   [10,5] ReLU, O=2, empirical sigma (bnn_regress.py shapes)
  c3  block-masked    N=200k, F=40, [120,80] tanh, K=5, 8 chains           (block_bnns.py construction)
  c5  prediction      S posterior samples x N=1M rows, [64,32] swish, K=10 (RunPredict / get_posterior_cat_prob)

Synthetic data of the named shapes.  Each line: chain-steps/s (or row-samples/s), ms per
step, algorithmic TFLOP/s.  Timed with CUDA events on the launch stream after warm-up.

    python tools/bench_configs.py [--out profiles/rNN_configs.json] [--only c1,c3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from npbnn_b200 import _lib as L  # noqa: E402
from npbnn_b200 import api, workloads as wl  # noqa: E402
from npbnn_b200.engine import Engine, NetShape, flatten_weights  # noqa: E402


def timed(fn, reps=3):
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def shapes_for(n_features, n_nodes, n_out, bias_mode):
    b_first = 1 if bias_mode >= 1 else 0
    b_hidden = 1 if bias_mode >= 2 else 0
    b_last = 1 if (bias_mode >= 3 or bias_mode == -1) else 0
    s = [(n_nodes[0], n_features + b_first)]
    for i in range(1, len(n_nodes)):
        s.append((n_nodes[i], n_nodes[i - 1] + b_hidden))
    s.append((n_out, n_nodes[-1] + b_last))
    return s


def flop_per_row(shapes, n_features):
    f, width = 0, n_features
    for r, c in shapes:
        f += 2 * r * width + (r if c == width + 1 else 0)
        width = r
    return f


def mh_rate(name, x, y, xt, yt, shapes, act, lik, chains, steps, mask=None, sigma_mode=L.SIGMA_FIXED, flop_row=None,
            update_f=None, update_ws=None, graphs=True):
    F = x.shape[1]
    net = NetShape(F, shapes, act=act, lik=lik)
    eng = Engine(net, device=0)
    eng.set_data(x, y, xt, yt)
    rs = np.random.RandomState(1)
    w0 = []
    for c in range(chains):
        w = [rs.normal(0, 0.1, s) for s in shapes]
        if mask is not None:
            w = [a * m for a, m in zip(w, mask)]
        w0.append(w)
    eng.chains_init(w0, mask=mask, sigma_mode=sigma_mode, update_f=update_f, update_ws=update_ws, seed=7)
    # chunks of at most 100 iterations per call, as run_mcmc / MC3 issue them (sampling_f / swap_frequency); the
    # first call runs eagerly, the second captures the CUDA graph of the chunk, the timed ones replay it
    chunk = min(100, steps)
    if not graphs:
        eng.set_option("graphs", 0)
    for _ in range(3):
        eng.mh_steps(chunk)
    eng.synchronize()
    ms = timed(lambda: [eng.mh_steps(chunk) for _ in range(steps // chunk)])
    steps = steps // chunk * chunk
    st = eng.read_state(weights=False)
    n = x.shape[0] + (0 if xt is None else xt.shape[0])
    fr = flop_row if flop_row is not None else flop_per_row(shapes, F)
    out = {"config": name, "rows": int(n), "features": int(F), "shapes": [list(s) for s in shapes], "act": act, "chains": chains,
           "steps_timed": steps, "ms_per_step": ms / steps, "chain_steps_per_s": chains * steps / (ms * 1e-3),
           "algorithmic_flop_per_row": fr, "tflops": chains * steps * n * fr / (ms * 1e-3) / 1e12,
           "cuda_graphs": bool(graphs), "kernel": eng.last_kernel, "logLik_finite": bool(np.all(np.isfinite(st.logLik))),
           "acceptance": float(np.mean(st.n_accepted / np.maximum(st.iteration, 1)))}
    eng.close()
    return out


def c1(chains, steps, graphs=True):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2500, 128))
    y = rng.integers(0, 5, 2500).astype(np.int32)
    sh = shapes_for(128, [5, 5], 5, 2)
    return mh_rate("c1 classify [5,5] tanh (C=%d)" % chains, x[:2250], y[:2250], x[2250:], y[2250:], sh, "tanh",
                   L.LIK_CATEGORICAL, chains, steps, graphs=graphs)


def c2(chains, steps, graphs=True):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((999, 3))
    y = np.stack([x[:, 0] * 2 + x[:, 1], x[:, 2] - x[:, 0]], 1) + 0.1 * rng.standard_normal((999, 2))
    sh = shapes_for(3, [10, 5], 2, 2)
    return mh_rate("c2 regress [10,5] ReLU empirical sigma (C=%d)" % chains, x[:900], y[:900], x[900:], y[900:], sh, "ReLU",
                   L.LIK_GAUSSIAN, chains, steps, sigma_mode=L.SIGMA_EMPIRICAL, graphs=graphs)


def c3_mask():
    shapes = shapes_for(40, [120, 80], 5, -1)
    w = [np.zeros(s) for s in shapes]
    idx = [list(range(40)), sum(([g] * 3 for g in range(40)), []), []]
    npf = [[3] * 40, [2] * 40, []]
    return shapes, api.create_mask(w, idx, npf)


def c3(chains, steps, n=200_000):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, 40))
    shapes, mask = c3_mask()
    teacher = [rng.normal(0, 1, s) * m for s, m in zip(shapes, mask)]
    h = np.tanh(x @ teacher[0].T)
    h = np.tanh(h @ teacher[1].T)
    y = np.argmax(h @ teacher[2][:, 1:].T + teacher[2][:, 0], 1).astype(np.int32)
    sparse = 2 * (120 + 240 + 80 * 5) + 5                   # flop/row with the block sparsity exploited
    o = mh_rate("c3 block-masked [120,80] tanh (C=%d)" % chains, x, y, None, None, shapes, "tanh", L.LIK_CATEGORICAL,
                chains, steps, mask=mask, flop_row=sparse)
    o["dense_masked_flop_per_row"] = flop_per_row(shapes, 40)
    o["mask_nonzeros"] = [int(m.sum()) for m in mask]
    return o


def c5(S, n=1_000_000):
    x, _ = wl.c4_data(n, seed=0)
    rng = np.random.default_rng(5)
    base = flatten_weights(wl.c4_init_weights(1)[0])
    w = base[None, :] + rng.normal(0, 0.05, (S, base.size))
    net = NetShape(64, list(wl.C4_SHAPES), act="swish", lik=L.LIK_CATEGORICAL)
    eng = Engine(net, device=0)
    xd = torch.from_numpy(x).cuda()
    wd = torch.from_numpy(w).cuda()
    md = torch.empty((n, 10), dtype=torch.float64, device="cuda")
    vd = torch.empty((n, 10), dtype=torch.float64, device="cuda")
    import ctypes as C

    def run():
        L.check(eng.lib.bnn_predict(eng._h, C.c_void_p(xd.data_ptr()), n, C.c_void_p(wd.data_ptr()), S, None, None, None, 0,
                                    C.c_void_p(md.data_ptr()), C.c_void_p(vd.data_ptr()), None, eng._stream()))
    run()
    ms = timed(run, reps=2)
    ok = bool(torch.isfinite(md).all().item()) and abs(float(md.sum().item()) - n) < 1e-6 * n
    out = {"config": "c5 posterior prediction [64,32] swish, mean + votes summaries", "rows": n, "samples": S, "ms": ms,
           "row_samples_per_s": S * n / (ms * 1e-3), "rows_per_s_all_samples": n / (ms * 1e-3),
           "tflops": S * n * wl.C4_FLOP_PER_ROW / (ms * 1e-3) / 1e12, "kernel": eng.last_kernel, "probabilities_sum_to_1": ok}
    eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="c1,c2,c3,c5")
    ap.add_argument("--c5-samples", type=int, default=128)
    args = ap.parse_args()
    only = set(args.only.split(","))
    res = []
    if "c1" in only:
        res += [c1(1, 2000, graphs=False), c1(1, 2000), c1(32, 1000)]
    if "c2" in only:
        res += [c2(1, 2000, graphs=False), c2(1, 2000), c2(32, 1000)]
    if "c3" in only:
        res += [c3(8, 40)]
    if "c5" in only:
        res += [c5(args.c5_samples)]
    for r in res:
        print(json.dumps(r), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "gpu": torch.cuda.get_device_name(0),
                       "results": res}, f, indent=1)


if __name__ == "__main__":
    main()
