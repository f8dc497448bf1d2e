#!/usr/bin/env python
"""Kernel-only timing of the forward+likelihood pass on the c4 workload (tuning helper).
usage: tools/fwd_bench.py [rows] [chains] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from npbnn_b200 import _lib as L, workloads as wl
from npbnn_b200.engine import Engine, NetShape, flatten_weights

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(rows, 64, dtype=torch.float64, device="cuda", generator=g)
y = torch.randint(0, 10, (rows,), dtype=torch.int32, device="cuda", generator=g)
net = NetShape(64, list(wl.C4_SHAPES), act="swish", lik=L.LIK_CATEGORICAL)
eng = Engine(net)
eng.set_data(x, y)
w = torch.as_tensor(np.stack([flatten_weights(s) for s in wl.c4_init_weights(chains)])).cuda()
ll = torch.empty(chains, dtype=torch.float64, device="cuda")
cnt = torch.zeros(chains, 22, dtype=torch.int32, device="cuda")
import ctypes as C
def run():
    L.check(eng.lib.bnn_forward_lik(eng._h, C.c_void_p(w.data_ptr()), chains, None, None, 0, 1.0, C.c_void_p(ll.data_ptr()), None, C.c_void_p(cnt.data_ptr()), None))
for _ in range(2): run()
torch.cuda.synchronize()
eng.set_option("time_forward", 1); eng.forward_time(True)
for _ in range(reps): run()
torch.cuda.synchronize()
ms, n = eng.forward_time(True)
avg = ms / n
tf = chains * rows * wl.C4_FLOP_PER_ROW / (avg * 1e-3) / 1e12
print("%s lib=%s rows=%d chains=%d: %.3f ms/launch  %.2f TFLOP/s  (%.1f chain-steps/s)  loglik[0]=%.6f" % (
    eng.last_kernel, os.path.basename(L.LIB_PATH), rows, chains, avg, tf, chains / (avg * 1e-3), ll[0].item()))
