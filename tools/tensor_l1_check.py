#!/usr/bin/env python
"""Accuracy and speed of the tensor-core first layer (k_fwd3t) against the FP64 DMMA kernel (k_fwd3) on the c4 shape.
    python tools/tensor_l1_check.py [rows] [sets]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from npbnn_b200 import workloads as wl  # noqa: E402
from npbnn_b200.engine import Engine, NetShape, flatten_weights  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    x, y = wl.c4_data(n, seed=0)
    w = np.stack([flatten_weights(ws) for ws in wl.c4_init_weights(S)])
    net = NetShape(64, list(wl.C4_SHAPES), act="swish", lik=0)
    eng = Engine(net)
    eng.set_data(x, y)
    wd = torch.from_numpy(w).cuda()
    out = {}
    for name, opt in (("tensor", 1), ("f64", 0)):
        eng.set_option("tensor_l1", opt)
        eng.forward_lik(wd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = eng.forward_lik(wd)
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"kernel": eng.last_kernel, "ms_per_pass": e0.elapsed_time(e1) / 3, "loglik": r["loglik"], "counts": r["counts"]}
    if os.environ.get("NPBNN_DBG_COUNTERS"):
        import ctypes as C
        buf = (C.c_ulonglong * 48)()
        eng.set_option("tensor_l1", 1)
        eng.lib.bnn_debug_counters(eng._h, C.cast(buf, C.c_void_p))           # reset
        eng.forward_lik(wd)
        eng.lib.bnn_debug_counters(eng._h, C.cast(buf, C.c_void_p))
        v = list(buf)
        names_h = ["wait hfull", "wait tfull", "wait a1free", "ctl wait tfree", "ctl wait w1full", "ctl wait rempty", "pass1", "pass2",
                   "ctl issue x4", "stage_x", "issue->tfull seen"]
        names_c = ["wait rfull", "wait a1full", "L2 phase", "whole chain body"]
        uses = 148 * -(-(n // 128 + (n % 128 > 0)) // 148) * S
        print("control warp (clocks per use):", {k: round(v[i] / uses, 1) for i, k in enumerate(names_h)})
        print("helpers (4 warps, clocks per use per warp):", {k: round(v[16 + i] / (uses * 4), 1) for i, k in enumerate(names_h)})
        print("compute (8 warps/SM, clocks per use per warp):", {k: round(v[32 + i] / (uses * 8), 1) for i, k in enumerate(names_c)})
    a, b = out["tensor"], out["f64"]
    rel = np.abs(a["loglik"] - b["loglik"]) / np.abs(b["loglik"])
    print(json.dumps({"rows": n, "sets": S, "tensor_kernel": a["kernel"], "f64_kernel": b["kernel"],
                      "tensor_ms": a["ms_per_pass"], "f64_ms": b["ms_per_pass"], "speedup": b["ms_per_pass"] / a["ms_per_pass"],
                      "max_rel_diff_loglik": float(rel.max()), "counts_equal": bool(np.array_equal(a["counts"], b["counts"])),
                      "loglik0": [float(a["loglik"][0]), float(b["loglik"][0])]}))


if __name__ == "__main__":
    main()
