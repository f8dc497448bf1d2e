"""Synthetic data shared by the golden generator (tests/golden/make_golden.py: synth_class) and the tests."""
import numpy as np


def synth_class(n, f, k, seed, n_test=0):
    """Noisy linear-teacher classification set; every class present in the first k rows of each split."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n + n_test, f))
    wt = rng.normal(0, 1.0, (k, f))
    lab = np.argmax(x @ wt.T + rng.gumbel(size=(n + n_test, k)), axis=1)
    lab[:k] = np.arange(k)
    if n_test:
        lab[n:n + k] = np.arange(k)
    return {"data": x[:n], "labels": lab[:n], "test_data": x[n:] if n_test else [],
            "test_labels": lab[n:] if n_test else []}
