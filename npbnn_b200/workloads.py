"""Synthetic workloads of BASELINE.json: shapes, data and initial weights."""
import numpy as np

C4_SHAPES = [(64, 64), (32, 64), (10, 33)]      # [64,32] hidden, 10 classes, use_bias_node=-1 (bias on the last layer)
C4_FLOP_PER_ROW = 2 * (64 * 64 + 64 * 32 + 32 * 10) + 10     # contraction + bias = 12,938


def swish_np(z):
    return z * (1.0 + np.exp(-z)) ** (-1)


def c4_data(n_rows=1_000_000, seed=0):
    """X ~ N(0,1) [n,64]; labels = argmax of a [64,32] swish teacher with N(0,0.5) weights."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n_rows, 64))
    teacher = [rng.normal(0, 0.5, s) for s in C4_SHAPES]
    h = swish_np(x @ teacher[0].T)
    h = swish_np(h @ teacher[1].T)
    logits = h @ teacher[2][:, 1:].T + teacher[2][:, 0]
    return x, np.argmax(logits, axis=1).astype(np.int32)


def c4_init_weights(n_chains, first_chain=0):
    """Per-chain N(0, 0.1) initial weights, chain c seeded with 1000 + c (init_weight_prm, BNN_mcmc.py:9-25)."""
    out = []
    for c in range(first_chain, first_chain + n_chains):
        rs = np.random.RandomState(1000 + c)
        out.append([rs.normal(0, 0.1, s) for s in C4_SHAPES])
    return out
