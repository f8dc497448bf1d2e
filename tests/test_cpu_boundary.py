"""CPU: the C-ABI library loads and exports every symbol include/npbnn_b200.h declares; host-side
logic (MC3 partition / swap, net-shape inference, weight (un)flattening); product path has no CPU fallback."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "npbnn_b200.h")).read()
    return sorted(set(re.findall(r"\b(bnn_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as ge
    ge.build()
    from npbnn_b200 import _lib
    import ctypes
    h = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(h, n), "library does not export %s" % n
    # the ctypes binding covers exactly the header
    assert set(_lib.EXPORTED_SYMBOLS) == set(names)
    assert h.bnn_abi_version() == 1


def test_state_slot_constants_match_header():
    from npbnn_b200 import _lib
    src = open(os.path.join(ROOT, "include", "npbnn_b200.h")).read()
    consts = dict((k, int(v)) for k, v in re.findall(r"\b(BNN_[FI]_[A-Z0-9_]+)\s*=\s*(\d+)", src))
    for k, v in consts.items():
        py = k.replace("BNN_", "")
        assert getattr(_lib, py) == v, k
    assert _lib.MAX_LAYERS == 8 and _lib.MAX_OUT == 32


def test_no_cpu_fallback_engine_raises_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from npbnn_b200.engine import Engine, NetShape
    from npbnn_b200 import _lib
    with pytest.raises(_lib.NpbnnError):
        Engine(NetShape(4, [(3, 4), (2, 3)]))


def test_product_code_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "npbnn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(dirpath, f)


def test_net_shape_bias_inference_and_flatten_roundtrip():
    from npbnn_b200.engine import NetShape, flatten_weights, unflatten_weights
    rng = np.random.default_rng(0)
    w = [rng.normal(size=(5, 129)), rng.normal(size=(5, 6)), rng.normal(size=(5, 5))]
    net = NetShape.from_weights(w, 128)
    assert net.has_bias() == [1, 1, 0] and net.n_params == 700
    back = unflatten_weights(flatten_weights(w), net.shapes)
    assert all(np.array_equal(a, b) for a, b in zip(w, back))
    with pytest.raises(ValueError):
        NetShape(10, [(4, 13)]).has_bias()


def test_mc3_partition_and_swap_rule():
    from npbnn_b200 import mc3
    from oracle import npbnn_oracle as orc
    for n, world in [(32, 1), (32, 8), (5, 2), (3, 4)]:
        parts = [mc3.chain_partition(n, world, r) for r in range(world)]
        assert parts[0][0] == 0 and sum(p[1] for p in parts) == n
        for a, b in zip(parts, parts[1:]):
            assert a[0] + a[1] == b[0]
    assert np.array_equal(mc3.default_temperatures(4, 0.8), np.linspace(0.8, 1, 4))
    assert mc3.default_temperatures(1)[0] == 1.0
    rng = np.random.default_rng(0)
    for _ in range(50):
        lp = rng.normal(-100, 5, 6)
        t = np.linspace(0.8, 1, 6)
        j, k = rng.choice(6, 2, replace=False)
        lu = float(np.log(rng.random()))
        got, sw = mc3.swap_temperatures(lp, t, int(j), int(k), lu)
        exp, sw2, _ = orc.mc3_swap(lp, t, int(j), int(k), lu)
        assert sw == sw2 and np.array_equal(got, exp)


def test_mc3_swaps_match_reference_golden():
    from npbnn_b200 import mc3
    from tests import _golden as G
    z, meta = G.load("mc3")
    for it in range(meta["n_mc3_iterations"]):
        j, k = [int(v) for v in z["it%d_pair" % it]]
        t, _ = mc3.swap_temperatures(z["it%d_logPost" % it], z["it%d_temps_before" % it], j, k, float(z["it%d_log_u" % it]))
        assert np.array_equal(t, z["it%d_temps_after" % it])


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from npbnn_b200 import mc3
rank, world = int(sys.argv[1]), 2
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[2]
dist.init_process_group("gloo", rank=rank, world_size=world)
n = 5
lp_all = np.array([-10.0, -12.5, -9.25, -11.0, -8.75])
temps = mc3.default_temperatures(n, 0.8)
rng = mc3.SwapRNG(7)
start, cnt = mc3.chain_partition(n, world, rank)
log = []
for it in range(20):
    local = torch.tensor(lp_all[start:start + cnt] + 0.1 * it * (np.arange(cnt) + start))
    temps, swapped, pair, lp = mc3.exchange(local, temps, rng, None, world)
    log.append((pair, swapped, temps.copy(), lp.copy()))
np.save(sys.argv[3] + "/temps_%%d.npy" %% rank, np.array([l[2] for l in log]))
np.save(sys.argv[3] + "/lp_%%d.npy" %% rank, np.array([l[3] for l in log]))
dist.destroy_process_group()
'''


def test_mc3_exchange_two_ranks_gloo(tmp_path):
    """world_size-2 run of the swap step on CPU (gloo): both ranks gather the same log-posteriors and
    reach identical temperature vectors, equal to a single-process replay."""
    from npbnn_b200 import mc3
    script = tmp_path / "w.py"
    script.write_text(WORKER % {"root": ROOT})
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), port, str(tmp_path)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    t0, t1 = np.load(tmp_path / "temps_0.npy"), np.load(tmp_path / "temps_1.npy")
    l0, l1 = np.load(tmp_path / "lp_0.npy"), np.load(tmp_path / "lp_1.npy")
    assert np.array_equal(t0, t1) and np.array_equal(l0, l1)
    # single-process replay
    n = 5
    lp_all = np.array([-10.0, -12.5, -9.25, -11.0, -8.75])
    temps = mc3.default_temperatures(n, 0.8)
    rng = mc3.SwapRNG(7)
    for it in range(20):
        lp = lp_all + 0.1 * it * np.arange(n)
        assert np.allclose(l0[it], lp)
        j, k = rng.pair(n)
        temps, _ = mc3.swap_temperatures(lp, temps, j, k, rng.log_uniform())
        assert np.array_equal(temps, t0[it])
    assert sorted(t0[-1]) == sorted(mc3.default_temperatures(n, 0.8))


def test_row_partition_covers_all_rows_once():
    from npbnn_b200 import rowshard
    for n, w in ((10, 3), (1_000_003, 8), (5, 8), (64, 1)):
        spans = [rowshard.row_partition(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


RS_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from npbnn_b200 import rowshard
rank, world = int(sys.argv[1]), 2
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[2]
dist.init_process_group("gloo", rank=rank, world_size=world)
per_row = np.load(sys.argv[3] + "/per_row.npy")            # [C, n_rows, n_values]
a, b = rowshard.row_partition(per_row.shape[1], world, rank)
out = []
for it in range(3):
    red = torch.from_numpy(per_row[:, a:b].sum(axis=1) * (it + 1))
    rowshard.dist_all_reduce_sum(red)
    out.append(red.numpy().copy())
np.save(sys.argv[3] + "/red_%%d.npy" %% rank, np.array(out))
dist.destroy_process_group()
'''


def test_rowshard_exchange_two_ranks_gloo(tmp_path):
    """world_size-2 run of the row-sharding exchange on CPU (gloo): each rank sums its own rows' contributions, the
    all-reduce leaves the same bits on both ranks (so both take the same accept decision) and they equal the
    unsharded sums."""
    rs = np.random.default_rng(5)
    per_row = rs.normal(size=(3, 1001, 23))
    np.save(tmp_path / "per_row.npy", per_row)
    script = tmp_path / "w.py"
    script.write_text(RS_WORKER % {"root": ROOT})
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), port, str(tmp_path)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    r0, r1 = np.load(tmp_path / "red_0.npy"), np.load(tmp_path / "red_1.npy")
    assert np.array_equal(r0, r1)
    for it in range(3):
        np.testing.assert_allclose(r0[it], per_row.sum(axis=1) * (it + 1), rtol=1e-12, atol=1e-12)
