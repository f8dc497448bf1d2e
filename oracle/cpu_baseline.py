"""CPU baseline timing of the hot path with the numpy oracle -- TEST/BENCH INFRASTRUCTURE ONLY.

FALLBACK of bench.py's `cpu_baseline` leg and `--impl reference` arm: they time the unmodified reference
(oracle/_ref, oracle/ref_baseline.py) and only use this port (kind "port") when oracle/_ref is absent.  Same numpy
operations as the reference (np.dot, the `(1+exp(-z))**-1` swish, exp/sum softmax, fancy-index gather + log), one
process per chain on its own core like the reference's MC3 fork pool (BNN_mc3.py:89-96), on a row sample rescaled
linearly in rows; kept as a cross-check of the reference timing.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from oracle import npbnn_oracle as orc

_SHARED = {}


def _worker(args):
    chain, n_steps, seconds = args
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=1)
    except Exception:
        ctx = None
    x, labels, shapes, act = _SHARED["x"], _SHARED["labels"], _SHARED["shapes"], _SHARED["act"]
    rs = np.random.RandomState(1000 + chain)
    w = [rs.normal(0, 0.1, s) for s in shapes]
    m = orc.Model(x=x, labels=labels, weights=w, act=act, mode="classification", prior=1)
    s = orc.make_sampler(m, n_iteration=100000)
    rng = np.random.default_rng(chain)
    orc.mh_step(m, s, rs=rng)                    # warm-up
    t0 = time.perf_counter()
    done = 0
    while done < n_steps and (done == 0 or time.perf_counter() - t0 < seconds):
        orc.mh_step(m, s, rs=rng)
        done += 1
    el = time.perf_counter() - t0
    if ctx is not None:
        ctx.unregister() if hasattr(ctx, "unregister") else None
    return done, el


def mh_rate(x, labels, shapes, act, n_full_rows, n_procs=None, steps_per_proc=3, seconds=20.0):
    """Chain-steps per second of the oracle on `x` (a row sample), rescaled to n_full_rows rows.
    Returns dict(value, cores, sample)."""
    cores = os.cpu_count() or 1
    n_procs = max(1, min(n_procs or cores, cores, 32))
    _SHARED.update(x=x, labels=labels, shapes=shapes, act=act)
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(n_procs) as pool:
        res = pool.map(_worker, [(c, steps_per_proc, seconds) for c in range(n_procs)])
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    rate_sample = steps / slowest
    scale = x.shape[0] / float(n_full_rows)
    return {"value": rate_sample * scale, "cores": n_procs, "wall_s": wall,
            "sample": "%d chains x %d MH steps (numpy oracle port, 1 process per core, 1 BLAS thread each) on %d of %d rows; "
                      "rate rescaled linearly in rows" % (n_procs, steps // n_procs, x.shape[0], n_full_rows)}
