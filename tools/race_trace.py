#!/usr/bin/env python
"""Bisect an intermittent corruption of k_fwd3t with a -DBNN_DBG_CSUM build: every compute warp records, per weight-set
use, a checksum of the layer-1 activations it read and of the logits it produced.  Passes are compared with each
other: the first differing (CTA, use, warp, field) tells whether the fault enters before or after the hand-off."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from npbnn_b200 import workloads as wl
from npbnn_b200.engine import Engine, NetShape, flatten_weights

n, S = int(os.environ.get("N", "1000000")), int(os.environ.get("S", "32"))
reps = int(os.environ.get("REPS", "8"))
x, y = wl.c4_data(n, seed=0)
w = np.stack([flatten_weights(ws) for ws in wl.c4_init_weights(S)])
eng = Engine(NetShape(64, list(wl.C4_SHAPES), act="swish", lik=0))
eng.set_data(x, y)
wd = torch.from_numpy(w).cuda()
n_tiles = (n + 127) // 128
n_iter = -(-n_tiles // 148)
total_q = n_iter * S
trace = torch.zeros((148, total_q, 8, 2), dtype=torch.int64, device="cuda")
eng.lib.bnn_debug_set_trace(eng._h, C.c_void_p(trace.data_ptr()))
eng.set_option("tensor_l1", 0)
ref = eng.forward_lik(wd)["loglik"]
eng.set_option("tensor_l1", 1)
runs = []
for rep in range(reps):
    trace.zero_()
    ll = eng.forward_lik(wd)["loglik"]
    bad = np.nonzero(np.abs(ll - ref) > 1e-6)[0]
    runs.append((trace.cpu().numpy().copy(), bad.tolist()))
    print("rep", rep, "bad chains", bad.tolist(), flush=True)
good = [t for t, b in runs if not b]
if not good:
    print("no clean pass to compare with")
    sys.exit(0)
g = good[0]
for i, (t, b) in enumerate(runs):
    d = np.argwhere(t != g)
    if len(d) == 0:
        continue
    fields = sorted(set(d[:, 3].tolist()))
    uses = sorted(set((d[:, 1] % S).tolist()))
    its = sorted(set((d[:, 1] // S).tolist()))
    warps = sorted(set(d[:, 2].tolist()))
    print("rep", i, "bad" if b else "clean", ": %d differing entries; fields %s (0 = activations read, 1 = logits); chains %s; iterations %s; warps %s; CTAs %s"
          % (len(d), fields, uses, its, warps, sorted(set(d[:, 0].tolist()))[:10]))
eng.lib.bnn_debug_set_trace(eng._h, None)
