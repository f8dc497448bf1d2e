"""Row sharding of the MH step over ranks: every rank holds a contiguous block of the training rows
and of the test rows and runs the SAME chains; per MH iteration one all-reduce of C * (slots + 2 + 2K) doubles (the
per-chain log-likelihood / residual sums and the accuracy counters) is the only exchange.  Chain sharding (mc3.py) is
the mode for many chains; this is the mode for one or a few chains on many rows (run_mcmc on several GPUs)."""
import torch
import torch.distributed as dist


def row_partition(n_rows: int, world: int, rank: int):
    """Contiguous, balanced split: rank r owns rows [n r / world, n (r + 1) / world)."""
    return n_rows * rank // world, n_rows * (rank + 1) // world


def dist_all_reduce_sum(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM over the default process group (NCCL over NVLink on a B200 box, gloo in the CPU tests).  The result
    is bit-identical on every rank, so the accept decisions are."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_rows(x, y, world: int, rank: int):
    a, b = row_partition(len(x), world, rank)
    return x[a:b], y[a:b]
