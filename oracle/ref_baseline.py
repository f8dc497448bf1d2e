"""Times the UNMODIFIED reference (oracle/_ref/np_bnn, see oracle/make_ref.py) on the host cores.

TEST / BENCH INFRASTRUCTURE ONLY -- run as a subprocess by bench.py (`--impl reference` and the `cpu_baseline` leg), so
that `import np_bnn` resolves to the reference copy and never to this repository's drop-in package:

    python oracle/ref_baseline.py --rows 1000000 --steps 20 --warmup 5 [--chains 32] [--mc3-period K]

Workload: BASELINE.json configs[3] (SURVEY.md 8d "c4") -- synthetic 1,000,000 x 64 float64 features, 10 teacher-generated
classes, n_nodes=[64,32], ActFun("swish"), use_bias_node=-1, Normal(0,1) prior, default update_f / update_ws, built with
the reference's own constructors (np_bnn.npBNN, np_bnn.MCMC; bnn_runner_MC3.py:28-48).

  chains leg   min(32, cores, memory) chains, one forked process per chain (the reference's MC3 runs its chains in a
               fork pool, BNN_mc3.py:89-96), each calling np_bnn.MCMC.mh_step (BNN_env.py:381-532) on ALL rows:
               W warm-up iterations, a barrier, K timed iterations, a barrier.  One BLAS thread per process (the
               processes already occupy every core).  No pickling of the data between swap periods, i.e. generous to
               the reference.  value = chains * K / wall.
  mc3 leg      (--mc3-period K > 0) np_bnn.MC3(n_chains=4, ...).run_mcmc() for ONE swap period of K iterations, the
               reference's stock driver, including its per-period pickling of every chain's [bnn, mcmc] (X included)
               through the pool (BNN_mc3.py:96).  Reported beside the chains leg, not as the headline value.

Prints one JSON object on stdout.
"""
import argparse
import contextlib
import io
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "_ref"))

import numpy as np  # noqa: E402

C4_SHAPES = [(64, 64), (32, 64), (10, 33)]


def c4_data(n_rows, seed=0):
    """Same generator as npbnn_b200/workloads.py:c4_data (kept separate: this process must not import the product)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n_rows, 64))
    teacher = [rng.normal(0, 0.5, s) for s in C4_SHAPES]
    sw = lambda z: z * (1.0 + np.exp(-z)) ** (-1)     # noqa: E731
    h = sw(sw(x @ teacher[0].T) @ teacher[1].T)
    return x, np.argmax(h @ teacher[2][:, 1:].T + teacher[2][:, 0], axis=1)


def _mem_limit_bytes():
    lim = None
    try:
        import psutil
        lim = psutil.virtual_memory().available
    except Exception:
        pass
    for p in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(p).read().strip()
            if v.isdigit():
                used = 0
                try:
                    used = int(open(os.path.join(os.path.dirname(p), "memory.current")).read())
                except Exception:
                    pass
                lim = min(lim, int(v) - used) if lim else int(v) - used
        except Exception:
            pass
    return lim


def _quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def _chain_proc(c, dat, warmup, steps, bar, out):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    import np_bnn as bn
    np.random.seed(1000 + c)
    bnn = _quiet(bn.npBNN, dat, n_nodes=[64, 32], use_bias_node=-1, actFun=bn.ActFun(fun="swish"), seed=1000 + c)
    mcmc = bn.MCMC(bnn, n_iteration=100000, mcmc_id=c, randomize_seed=True)
    ll0 = float(mcmc._logLik)
    for _ in range(warmup):
        mcmc.mh_step(bnn)
    bar.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        mcmc.mh_step(bnn)
    t1 = time.perf_counter()
    bar.wait()
    out.put((c, t1 - t0, ll0, float(mcmc._logLik), int(mcmc._current_iteration)))


def chains_leg(dat, n_chains, warmup, steps):
    ctx = mp.get_context("fork")
    bar = ctx.Barrier(n_chains + 1)
    out = ctx.Queue()
    procs = [ctx.Process(target=_chain_proc, args=(c, dat, warmup, steps, bar, out)) for c in range(n_chains)]
    for p in procs:
        p.start()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    wall = time.perf_counter() - t0
    res = [out.get() for _ in procs]
    for p in procs:
        p.join()
    assert all(p.exitcode == 0 for p in procs)
    assert all(r[4] == warmup + steps and np.isfinite(r[3]) for r in res)
    return {"wall_s": wall, "per_chain_s": sorted(r[1] for r in res), "value": n_chains * steps / wall}


def mc3_leg(dat, n_chains, period):
    import np_bnn as bn
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            np.random.seed(1234)
            t0 = time.perf_counter()
            bnn = _quiet(bn.npBNN, dat, n_nodes=[64, 32], use_bias_node=-1, actFun=bn.ActFun(fun="swish"), seed=1)
            logger = bn.postLogger(bnn, filename="ref_mc3", log_all_weights=0)
            mc3 = _quiet(bn.MC3, bnn, logger=logger, n_post_samples=10, sampling_f=period, n_iteration=period,
                         n_chains=n_chains, swap_frequency=period, verbose=0)
            t1 = time.perf_counter()
            _quiet(mc3.run_mcmc)
            t2 = time.perf_counter()
        finally:
            os.chdir(cwd)
    return {"setup_s": t1 - t0, "wall_s": t2 - t1, "value": n_chains * period / (t2 - t1), "swap_frequency": period,
            "what": "np_bnn.MC3(n_chains=%d, swap_frequency=%d, n_iteration=%d).run_mcmc(): one swap period incl. the "
                    "pool's pickling of every chain's [bnn, mcmc] and the logger's pickle" % (n_chains, period, period)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--chains", type=int, default=32)
    ap.add_argument("--mc3-period", type=int, default=0)
    ap.add_argument("--gb-per-chain", type=float, default=2.5)
    a = ap.parse_args()
    import np_bnn as bn
    assert os.path.realpath(os.path.dirname(bn.__file__)).startswith(os.path.realpath(os.path.join(HERE, "_ref"))), bn.__file__
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    n_chains = max(1, min(a.chains, cores))
    mem = _mem_limit_bytes()
    gb = a.gb_per_chain * a.rows / 1e6
    if mem:
        n_chains = max(1, min(n_chains, int((mem / 1e9 - 2.0) / max(gb, 1e-3))))
    x, lab = c4_data(a.rows, seed=0)
    dat = {"data": x, "labels": lab, "label_dict": np.unique(lab), "test_data": [], "test_labels": []}
    res = {"reference_version": getattr(bn, "__version__", "?"), "cores": cores, "chains": n_chains, "rows": a.rows,
           "steps": a.steps, "warmup": a.warmup, "blas_threads_per_process": 1,
           "numpy": np.__version__}
    ch = chains_leg(dat, n_chains, a.warmup, a.steps)
    res["chains_leg"] = ch
    res["value"] = ch["value"]
    res["sample"] = ("UNMODIFIED reference np_bnn %s (oracle/_ref): %d forked chains (1 process per core, 1 BLAS thread each) x "
                     "%d np_bnn.MCMC.mh_step iterations after %d warm-up, on %d of %d rows (no subsampling, no rescaling)"
                     % (res["reference_version"], n_chains, a.steps, a.warmup, a.rows, a.rows))
    if a.mc3_period > 0:
        try:
            # MC3's pool.map pickles the bound method's `self` -- the MC3 object with EVERY chain's copy of X -- once per
            # task (BNN_mc3.py:96), i.e. chains^2 x 512 MB per swap period at this size: 32 chains would move 0.5 TB
            # through pipes.  The leg therefore runs with the reference's default n_chains=4 (BNN_mc3.py:14).
            res["mc3_leg"] = mc3_leg(dat, min(n_chains, 4), a.mc3_period)
        except Exception as e:                                   # e.g. out of memory while pickling X per chain
            res["mc3_leg"] = {"error": repr(e)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
