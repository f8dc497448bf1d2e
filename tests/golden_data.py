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


def write_example_tables(outdir, seed=17, n=600, f=12, k=4, n_new=40, n_reg=300, f_reg=3):
    """Text tables in the formats of the reference's example_files/ (tab-separated; classification tables with a header
    row and instance names in column 0, string class names; regression tables without header or names), written from a
    seeded generator so that the golden generator and the tests see the same files.  Returns the paths."""
    import os
    rng = np.random.default_rng(seed)
    x = np.round(rng.standard_normal((n + n_new, f)), 6)
    x[:, 3] = rng.integers(0, 3, n + n_new)                       # an ordinal feature
    wt = rng.normal(0, 1.0, (k, f))
    lab = np.argmax(x @ wt.T + 0.5 * rng.gumbel(size=(n + n_new, k)), axis=1)
    names = np.array(["cls_%s" % "dacb"[i % 4] + str(i // 4) for i in range(k)])      # unsorted class names
    p = {key: os.path.join(outdir, key + ".txt") for key in ("features", "labels", "unlabeled", "features_reg", "labels_reg")}
    head = "\t".join(["id"] + ["feat%d" % j for j in range(f)])
    with open(p["features"], "w") as fh:
        fh.write(head + "\n")
        for i in range(n):
            fh.write("\t".join(["inst%d" % i] + [repr(float(v)) for v in x[i]]) + "\n")
    with open(p["labels"], "w") as fh:
        fh.write("id\tlabel\n")
        for i in range(n):
            fh.write("inst%d\t%s\n" % (i, names[lab[i]]))
    with open(p["unlabeled"], "w") as fh:
        fh.write(head + "\n")
        for i in range(n, n + n_new):
            fh.write("\t".join(["new%d" % i] + [repr(float(v)) for v in x[i]]) + "\n")
    xr = np.round(rng.standard_normal((n_reg, f_reg)), 6)
    yr = np.round(np.stack([xr[:, 0] - 0.5 * xr[:, 1], np.sin(xr[:, 2])], axis=1) + 0.1 * rng.standard_normal((n_reg, 2)), 6)
    np.savetxt(p["features_reg"], xr, delimiter="\t", fmt="%.6f")
    np.savetxt(p["labels_reg"], yr, delimiter="\t", fmt="%.6f")
    return p
