"""MC3 (Metropolis-coupled MCMC) host logic: chain partitioning over ranks and the swap step.

Replaces the fork pool of the reference (BNN_mc3.py:87-126), which pickles every chain (with its copy
of the data) to a worker and back each swap period.  Here the chains never leave their GPU; the only
exchange is an all-gather of the per-chain log-posteriors (8 bytes per chain) over NCCL/NVLink, after
which every rank evaluates the same swap with the same seeded generator and updates the temperatures
of its own chains.  No tensor of the data path is communicated.
"""
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def default_temperatures(n_chains: int, min_temperature: float = 0.8) -> np.ndarray:
    """np.linspace(min_temperature, 1, n_chains), [1] for a single chain (BNN_mc3.py:46-51)."""
    if n_chains == 1:
        return np.ones(1)
    return np.linspace(min_temperature, 1, n_chains)


def chain_partition(n_chains: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of chains owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_chains, world_size)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


class SwapRNG:
    """The generator behind the swap step.  Every rank builds it from the same seed, so the pair and the
    uniform agree everywhere without communication (the reference uses the global np.random state,
    BNN_mc3.py:99,109)."""

    def __init__(self, seed: int = 4321):
        self._rs = np.random.RandomState(seed)

    def pair(self, n_chains: int):
        j, k = self._rs.choice(range(n_chains), 2, replace=False)
        return int(j), int(k)

    def log_uniform(self) -> float:
        return float(np.log(self._rs.random_sample()))


class GlobalSwapRNG:
    """Single-process default: the swap step consumes numpy's GLOBAL generator exactly as the reference does
    (np.random.choice(range(n), 2, replace=False) then np.log(np.random.random()), BNN_mc3.py:99,109), so a script that
    calls np.random.seed(...) first swaps the same pairs as the reference run."""

    def pair(self, n_chains: int):
        j, k = np.random.choice(range(n_chains), 2, replace=False)
        return int(j), int(k)

    def log_uniform(self) -> float:
        return float(np.log(np.random.random()))


def broadcast_from_rank0(obj, group: Optional[dist.ProcessGroup] = None):
    """Small host object of rank 0 on every rank (seeds); identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def owner_of(chain: int, n_chains: int, world_size: int) -> int:
    """Rank that holds global chain index `chain` under chain_partition."""
    for r in range(world_size):
        start, n = chain_partition(n_chains, world_size, r)
        if start <= chain < start + n:
            return r
    raise ValueError("chain %d of %d" % (chain, n_chains))


def swap_temperatures(log_post: np.ndarray, temps: np.ndarray, j: int, k: int, log_u: float):
    """r = (lp_k - lp_j) T_j + (lp_j - lp_k) T_k; the two chains exchange temperatures (states stay put)
    iff r >= log u  (BNN_mc3.py:101-112)."""
    t = np.array(temps, dtype=np.float64, copy=True)
    r = (log_post[k] - log_post[j]) * t[j] + (log_post[j] - log_post[k]) * t[k]
    swapped = bool(r >= log_u)
    if swapped:
        t[j], t[k] = temps[k], temps[j]
    return t, swapped


def gather_log_post(local: torch.Tensor, counts, group: Optional[dist.ProcessGroup] = None) -> np.ndarray:
    """All-gather of the local chains' log-posteriors -> host vector over all chains (rank order).
    `counts[r]` = chains owned by rank r.  Uses the process group's backend (NCCL on GPUs, gloo in tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local.detach().cpu().numpy().astype(np.float64)
    world = dist.get_world_size(group)
    cmax = max(counts)
    send = torch.zeros(cmax, dtype=torch.float64, device=local.device)
    send[:local.numel()] = local
    recv = torch.empty(world * cmax, dtype=torch.float64, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, cmax)
    return np.concatenate([recv[r, :counts[r]] for r in range(world)])


def exchange(local_log_post: torch.Tensor, temps_all: np.ndarray, rng: SwapRNG,
             group: Optional[dist.ProcessGroup] = None, world_size: int = 1):
    """One swap step.  Returns (new temperatures of ALL chains, swapped?, (j, k), gathered log-posteriors)."""
    n = len(temps_all)
    counts = [chain_partition(n, world_size, r)[1] for r in range(world_size)]
    lp = gather_log_post(local_log_post, counts, group)
    if n < 2:
        return np.array(temps_all, dtype=np.float64), False, (0, 0), lp
    j, k = rng.pair(n)
    t, swapped = swap_temperatures(lp, temps_all, j, k, rng.log_uniform())
    return t, swapped, (j, k), lp
