#!/usr/bin/env python
"""Row-sharded MH on the c4 shape with ONE chain (or a few): rows of X split over the ranks, one
all-reduce of C x 23 doubles per iteration.  Launch with torchrun (1 process per GPU); rank 0 prints a JSON line.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/rowshard_bench.py [--chains 1]"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from npbnn_b200 import rowshard, workloads as wl  # noqa: E402
from npbnn_b200.engine import Engine, NetShape, flatten_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=1)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=100)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    x, y = wl.c4_data(args.rows, seed=0)
    a, b = rowshard.row_partition(args.rows, world, rank)
    w0 = np.stack([flatten_weights(w) for w in wl.c4_init_weights(args.chains)])
    eng = Engine(NetShape(64, list(wl.C4_SHAPES), act="swish", lik=0), device=local)
    eng.set_data(x[a:b], y[a:b])
    if world > 1:
        eng.enable_rowshard(args.rows, rowshard.dist_all_reduce_sum)
    eng.chains_init(w0, seed=1234)
    eng.mh_steps(10)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.mh_steps(args.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    st = eng.read_state(weights=False)
    if rank == 0:
        print(json.dumps({"mode": "rows sharded over ranks" if world > 1 else "single GPU", "n_gpus": world, "chains": args.chains,
                          "rows": args.rows, "steps": args.steps, "ms_per_step": float(ms.item()) / args.steps,
                          "chain_steps_per_s": args.chains * args.steps / (float(ms.item()) * 1e-3),
                          "logLik": [float(v) for v in st.logLik], "n_accepted": [int(v) for v in st.n_accepted],
                          "exchange_bytes_per_step": args.chains * 23 * 8}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
