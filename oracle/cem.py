"""oracle.cem -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Standalone torch-CPU fp32 restatement of the reference's CEM tick, following
``/root/reference/Optimizers/optimizer_cem_tf.py:54-117`` line by line with TensorFlow semantics
(tf.argsort ascending == top_k(-x): ties -> lower index; tf.math.reduce_std population std;
tf.clip_by_value == min(max(x,lo),hi)).  Pinned against the unmodified reference file (run through the
TF-API shim) via tests/golden/cem_*.npz.
"""
from __future__ import annotations

import numpy as np
import torch

from . import spec


class CEMOracle:
    def __init__(self, predictor, cost: spec.CostParams, *, mpc_horizon, num_rollouts, cem_outer_it,
                 cem_initial_action_stdev, cem_stdev_min, cem_best_k, warmup=False, warmup_iterations=0,
                 action_low=-1.0, action_high=1.0, dtype=torch.float32, **_ignored):
        self.dtype = dtype
        self.predictor = predictor
        self.cost = cost
        self.H, self.N = int(mpc_horizon), int(num_rollouts)
        self.outer_it = int(cem_outer_it)
        self.init_std = float(cem_initial_action_stdev)
        self.std_min = float(cem_stdev_min)
        self.k = int(cem_best_k)
        self.warmup, self.warmup_iterations = bool(warmup), int(warmup_iterations)
        self.low = torch.as_tensor(np.atleast_1d(np.asarray(action_low, np.float32))).to(dtype)
        self.high = torch.as_tensor(np.atleast_1d(np.asarray(action_high, np.float32))).to(dtype)
        self.nu = int(getattr(predictor, "num_control_inputs", 1))
        self.reset()

    def reset(self):  # optimizer_cem_tf.py:113-117
        self.dist_mue = (self.low + self.high) * 0.5 * torch.ones([1, self.H, self.nu], dtype=self.dtype)
        self.stdev = float(np.float32(self.init_std)) * torch.ones([1, self.H, self.nu], dtype=self.dtype)
        self.count = 0
        self.u = 0.0
        self.last = {}

    def step(self, s: np.ndarray, rng) -> np.ndarray:
        s = torch.as_tensor(np.tile(np.asarray(s, np.float32), (self.N, 1))).to(self.dtype)  # :86-87
        iterations = self.warmup_iterations if self.warmup and self.count == 0 else self.outer_it  # :92
        elite_log, cost_log = [], []
        for _ in range(iterations):  # :93-94 -> update_distribution :61-80
            Q = self.dist_mue.repeat(self.N, 1, 1) + torch.mul(
                rng.normal(shape=(self.N, self.H, self.nu), dtype=torch.float32).to(self.dtype), self.stdev)  # :64-65
            Q = torch.minimum(torch.maximum(Q, self.low), self.high)  # :66
            rollout = self.predictor.predict_core(s, Q)  # :57
            traj_cost = spec.trajectory_cost(rollout, Q, self.u, self.cost)  # :58
            sorted_cost = torch.argsort(traj_cost, stable=True)  # :73
            best_idx = sorted_cost[: self.k]  # :74
            elite_Q = torch.index_select(Q, 0, best_idx)  # :75
            self.dist_mue = torch.mean(elite_Q, dim=0, keepdim=True)  # :77
            mu = torch.mean(elite_Q, dim=0, keepdim=True)
            self.stdev = torch.sqrt(torch.mean((elite_Q - mu) * (elite_Q - mu), dim=0, keepdim=True))  # :78
            elite_log.append(best_idx.numpy().copy())
            cost_log.append(traj_cost.numpy().copy())
        # :99-102
        self.stdev = torch.minimum(torch.maximum(self.stdev, torch.tensor(float(np.float32(self.std_min)), dtype=self.dtype)), torch.tensor(1.0e8, dtype=self.dtype))
        self.stdev = torch.cat([self.stdev[:, 1:, :], float(np.float32(self.init_std)) * torch.ones((1, 1, self.nu), dtype=self.dtype)], dim=1)
        self.u = elite_Q[0, 0, :].squeeze().numpy().copy()
        self.dist_mue = torch.cat([self.dist_mue[:, 1:, :], (self.low + self.high) * 0.5 * torch.ones((1, 1, self.nu), dtype=self.dtype)], dim=1)
        self.last = dict(J=traj_cost.numpy(), Q=Q.numpy(), rollouts=rollout.numpy(), elite_idx=np.stack(elite_log), J_iters=np.stack(cost_log))
        self.count += 1  # :110
        return self.u
