"""npbnn_b200 -- B200-native (sm_100a CUDA behind a C ABI) implementation of the npBNN MCMC hot path.

The names below mirror the reference package (np_bnn/__init__.py:6-25) for the path that is rebuilt here:
model / sampler state, the MH drivers, MC3, block masks and the posterior-prediction callers.
"""
__version__ = "0.1.0"

from .hostlib import *  # noqa: F401,F403,E402  (file readers, summaries, proposal / Gibbs helpers, BNN_lik functions)
from .api import (ActFun, MC3, MCMC, RunPredict, RunPredictInd, SaveObject,  # noqa: F401,E402
                  create_mask, data_transform_obj, feature_importance, get_pdp, get_posterior_cat_prob, get_posterior_est,
                  init_weight_prm, make_pdp_features, npBNN, pdp, postLogger, predict, run_mcmc, sample_from_categorical)
from .engine import Engine, NetShape  # noqa: F401
