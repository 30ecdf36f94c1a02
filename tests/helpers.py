"""Shared helpers for the parity tests: fixture loading and oracle construction."""
import json
import os

import numpy as np

from oracle import spec
from oracle.cem import CEMOracle
from oracle.mppi import MPPIOracle
from oracle.replay_rng import ReplayRNG
from oracle.rpgd import RPGDOracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix=""):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith(prefix))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["config"]))
    return z, meta


def make_predictor(meta):
    if meta["predictor"].startswith("ODE"):
        return spec.ODEPredictor(spec.CartPoleParams(dt=meta["cfg"]["mpc_timestep"]))
    return spec.MLPPredictor(spec.MLPWeights.random_init(meta["mlp_seed"]))


def make_oracle(meta, **over):
    pred = make_predictor(meta)
    cost = spec.CostParams(name=meta["cost"])
    cfg = dict(meta["cfg"])
    cfg.update(over)
    cls = {"mppi": MPPIOracle, "cem-tf": CEMOracle, "rpgd": RPGDOracle}[meta["optimizer"]]
    return cls(pred, cost, **cfg)


def replay(meta):
    return ReplayRNG(meta["noise_seed"])


def rel_err(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
