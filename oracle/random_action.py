"""oracle.random_action -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Line-by-line restatement (torch-CPU fp32) of the reference's random-shooting optimizer
``Optimizers/optimizer_random_action_tf.py:48-76`` with TensorFlow semantics restated (``tf.argsort`` ascending, ties to the
lower index).  Pinned by tests/test_oracle_golden.py against fixtures produced by the UNMODIFIED reference file."""
from __future__ import annotations

import numpy as np
import torch

from . import spec


class RandomActionOracle:
    def __init__(self, predictor, cost: spec.CostParams, *, mpc_horizon, num_rollouts, action_low=-1.0, action_high=1.0,
                 dtype=torch.float32, **_ignored):
        self.dtype = dtype
        self.predictor, self.cost = predictor, cost
        self.H, self.N = int(mpc_horizon), int(num_rollouts)
        self.low, self.high = np.float32(action_low), np.float32(action_high)
        self.u = 0.0  # Optimizers/__init__.py:35
        self.last = {}

    def reset(self, rng):
        rng.uniform(shape=(self.N, self.H, 1), dtype=torch.float32, minval=self.low, maxval=self.high)  # :80-87 (drawn, unused)

    def step(self, s: np.ndarray, rng) -> np.ndarray:
        s = torch.as_tensor(np.tile(np.asarray(s, np.float32), (self.N, 1))).to(self.dtype)  # :53-54
        Q = rng.uniform(shape=(self.N, self.H, 1), dtype=torch.float32, minval=self.low, maxval=self.high).to(self.dtype)  # :56-61
        rollout = self.predictor.predict_core(s, Q)  # :42
        traj_cost = spec.trajectory_cost(rollout, Q, self.u, self.cost)  # :43-45
        best_idx = int(torch.argsort(traj_cost, stable=True)[0])  # :65-66
        self.u = Q[best_idx, 0, :].squeeze().numpy().astype(np.float32)  # :68
        self.last = dict(J=traj_cost.numpy(), Q=Q.numpy(), rollouts=rollout.numpy(), best_idx=best_idx)
        return self.u
