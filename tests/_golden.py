"""Load a golden case (tests/golden/*.npz, written by tests/golden/make_golden.py from the
unmodified reference) and turn it into oracle objects / injection sequences."""
import json
import os

import numpy as np

from oracle import npbnn_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CHAIN_CASES = ["c1_classify", "c2_regress_emp1", "c2_regress_emp0", "syn_swish_cauchy", "syn_genrelu_laplace",
               "syn_uniform_bound", "syn_classw_temp", "syn_instw", "syn_adapt", "syn_block_mask",
               "syn_regress_error", "syn_trainable_genrelu", "syn_trainable_tanh", "syn_c4_shape"]


class ChainView:
    """One chain of a multi-chain golden file (syn_c3_shape): keys are looked up as 'c<chain>_<key>' first, then as
    the shared '<key>' (data, masks), so the single-chain helpers below work on it unchanged."""

    def __init__(self, z, chain):
        self._z, self._p = z, "c%d_" % chain
        self.files = [k[len(self._p):] if k.startswith(self._p) else k for k in z.files]

    def __getitem__(self, key):
        return self._z[self._p + key] if self._p + key in self._z.files else self._z[key]


def stack_injections(per_chain):
    """Injection arrays of several single-chain replays -> one batch with the chains side by side."""
    cap = max(a["ix"].shape[2] for a in per_chain)
    out = {}
    for k in per_chain[0]:
        parts = []
        for a in per_chain:
            v = a[k]
            if k in ("ix", "iy", "dz") and v.shape[2] < cap:
                v = np.concatenate([v, np.zeros(v.shape[:2] + (cap - v.shape[2],), v.dtype)], axis=2)
            parts.append(v)
        out[k] = np.ascontiguousarray(np.concatenate(parts, axis=1))
    return out


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta


def n_layers(meta):
    return len(meta["n_nodes"]) + 1


def build_model(z, meta) -> orc.Model:
    nl = n_layers(meta)
    weights = [np.array(z["w0_%d" % i]) for i in range(nl)]
    mode = meta["mode"]
    labels = z["labels"].astype(np.int64) if mode == "classification" else z["labels"].astype(np.float64)
    lt = z["labels_test"]
    labels_test = lt.astype(np.int64) if mode == "classification" else lt.astype(np.float64)
    mask = [np.array(z["mask_%d" % i]) for i in range(nl)] if "mask_0" in z.files else None
    return orc.Model(
        x=np.array(z["x"]), labels=labels, weights=weights, act=meta["act"],
        alphas=np.array(meta["alphas"]) if meta.get("alphas") else None, mode=mode, prior=meta["prior"],
        prior_scale=np.ones(nl) * meta["p_scale"], w_bound=meta["w_bound"], mask=mask,
        class_w=np.array(z["class_w"]) if "class_w" in z.files else None,
        inst_w=np.array(z["inst_w"]) if "inst_w" in z.files else None,
        empirical_error=bool(meta.get("empirical_error", False)),
        error_prm=np.ones(labels.shape[1]) if mode == "regression" else 1.0,
        x_test=np.array(z["x_test"]) if len(z["x_test"]) else None,
        labels_test=labels_test if len(z["x_test"]) else None,
        act_trainable=bool(meta.get("trainable", False)))


def build_sampler(m, meta) -> orc.Sampler:
    return orc.make_sampler(m, update_f=meta["update_f"], update_ws=meta["update_ws"],
                            temperature=meta["temperature"], n_iteration=meta["n_iteration"],
                            lik_temp=meta["lik_temp"], adapt_f=meta["adapt_f"], adapt_fM=meta["adapt_fM"],
                            adapt_freq=meta["adapt_freq"], init_additional_prob=meta.get("init_additional_prob", 0.0))


def injection(z, meta, t) -> orc.StepInjection:
    nl = n_layers(meta)
    layers = []
    for i in range(nl):
        off = z["prop_l%d_off" % i]
        a, b = int(off[t]), int(off[t + 1])
        if z["steps_proposed"][t][i]:
            layers.append((z["prop_l%d_ix" % i][a:b].astype(np.int64), z["prop_l%d_iy" % i][a:b].astype(np.int64),
                           z["prop_l%d_dz" % i][a:b]))
        else:
            layers.append(None)
    extra = {}
    if meta.get("trainable"):
        extra = {"alpha_ix": int(z["steps_alpha_ix"][t]), "alpha_dz": float(z["steps_alpha_dz"][t])}
    return orc.StepInjection(rr=z["steps_rr"][t], layers=layers, log_u=float(z["steps_log_u"][t]), **extra)


def injection_arrays(z, meta, t0, t1, n_chains=1):
    """Pack steps [t0, t1) of a golden case into the arrays bnn_mh_steps expects (same draws for every chain)."""
    nl = n_layers(meta)
    T = t1 - t0
    cap = 1
    for t in range(t0, t1):
        tot = sum(int(z["prop_l%d_off" % i][t + 1] - z["prop_l%d_off" % i][t]) for i in range(nl))
        cap = max(cap, tot)
    proposed = np.zeros((T, n_chains, nl), np.int32)
    count = np.zeros((T, n_chains, nl), np.int32)
    ix = np.zeros((T, n_chains, cap), np.int32)
    iy = np.zeros((T, n_chains, cap), np.int32)
    dz = np.zeros((T, n_chains, cap))
    log_u = np.zeros((T, n_chains))
    for t in range(t0, t1):
        o = 0
        for i in range(nl):
            off = z["prop_l%d_off" % i]
            a, b = int(off[t]), int(off[t + 1])
            proposed[t - t0, :, i] = int(z["steps_proposed"][t][i])
            count[t - t0, :, i] = b - a
            ix[t - t0, :, o:o + b - a] = z["prop_l%d_ix" % i][a:b]
            iy[t - t0, :, o:o + b - a] = z["prop_l%d_iy" % i][a:b]
            dz[t - t0, :, o:o + b - a] = z["prop_l%d_dz" % i][a:b]
            o += b - a
        log_u[t - t0, :] = float(z["steps_log_u"][t])
    out = dict(proposed=proposed, count=count, ix=ix, iy=iy, dz=dz, log_u=log_u)
    if meta.get("trainable"):
        out["alpha_ix"] = np.repeat(z["steps_alpha_ix"][t0:t1].astype(np.int32)[:, None], n_chains, 1)
        out["alpha_dz"] = np.repeat(z["steps_alpha_dz"][t0:t1].astype(np.float64)[:, None], n_chains, 1)
    return out
