#!/usr/bin/env python
"""Stress probe for the opt-in tensor-core kernel (k_fwd3t): repeated scoring passes compared with the FP64 kernel,
down to the per-warp-tile partial sums, so that an intermittent corruption is located (weight set, 16-row tile,
128-row tile, iteration of the persistent loop).  Env: N rows, S weight sets, REPS passes.
    N=1000000 S=32 REPS=12 python tools/race_probe.py"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from npbnn_b200 import workloads as wl
from npbnn_b200.engine import Engine, NetShape, flatten_weights
n, S = int(os.environ.get("N", "1000000")), int(os.environ.get("S", "32"))
x, y = wl.c4_data(n, seed=0)
w = np.stack([flatten_weights(ws) for ws in wl.c4_init_weights(S)])
net = NetShape(64, list(wl.C4_SHAPES), act="swish", lik=0)
eng = Engine(net); eng.set_data(x, y)
wd = torch.from_numpy(w).cuda()
import ctypes as C
nt16 = (n + 15) // 16
def part():
    buf = np.empty((S, nt16))
    eng.lib.bnn_debug_read_part(eng._h, buf.ctypes.data_as(C.c_void_p), buf.size)
    return buf
eng.set_option("tensor_l1", 0)
ref = eng.forward_lik(wd)
pref = part()
eng.set_option("tensor_l1", 1)
bad = 0
dbuf = (C.c_ulonglong * 48)()
eng.lib.bnn_debug_counters(eng._h, C.cast(dbuf, C.c_void_p))
dbuf = (C.c_ulonglong * 48)()
eng.lib.bnn_debug_counters(eng._h, C.cast(dbuf, C.c_void_p))
for rep in range(int(os.environ.get("REPS", "6"))):
    r = eng.forward_lik(wd)
    eng.lib.bnn_debug_counters(eng._h, C.cast(dbuf, C.c_void_p))
    print("csum mismatches:", dbuf[40], "last q/warp/it:", dbuf[41], dbuf[42], dbuf[43], flush=True)
    d = r["loglik"] - ref["loglik"]
    cd = (r["counts"] != ref["counts"]).any(axis=1)
    idx = np.nonzero((np.abs(d) > 1e-6) | cd)[0]
    if len(idx):
        bad += 1
        pp = part()
        dd = np.abs(pp - pref) > 1e-9 * np.abs(pref) + 1e-12
        ch, tl = np.nonzero(dd)
        it0 = dd[:, : 148 * 8]
        ch0, tl0 = np.nonzero(it0)
        print("   iteration-0 bad (chain, tile16):", list(zip(ch0.tolist(), tl0.tolist()))[:16])
        for cc, tt in list(zip(ch0.tolist(), tl0.tolist()))[:3]:
            col = pref[:, tt]
            near = np.argsort(np.abs(col - pp[cc, tt]))[:3]
            print("      got", pp[cc, tt], "want", pref[cc, tt], " | same tile, other chains closest:", [(int(k), float(col[k])) for k in near],
                  " | next-iteration tile same chain:", float(pref[cc, tt + 148 * 8]) if tt + 148 * 8 < pref.shape[1] else None)
        print("   bad (chain, tile16):", list(zip(ch.tolist(), tl.tolist()))[:24], " tile128:", sorted(set((tl // 8).tolist()))[:12], " tile128 % 148:", sorted(set(((tl // 8) % 148).tolist()))[:12], " iter:", sorted(set(((tl // 8) // 148).tolist())))
        print("rep", rep, "bad chains", idx.tolist(), "dloglik", [float(d[i]) for i in idx], "dcorrect", [int(r["counts"][i,0]-ref["counts"][i,0]) for i in idx], flush=True)
print("bad reps", bad, "of 12")
