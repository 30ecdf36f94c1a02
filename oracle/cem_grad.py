"""oracle.cem_grad -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatements (torch-CPU fp32) of the reference's two gradient-assisted CEM optimizers, TensorFlow semantics restated as in
oracle/cem.py and oracle/gradient.py (``tf.argsort`` ascending with ties to the lower index, population ``reduce_std``,
``tf.clip_by_norm(g, c, axes=[1,2])`` = g c / max(||g||_2, c) per trajectory, legacy Keras Adam):

* ``CEMNaiveGradOracle``   -- ``Optimizers/optimizer_cem_naive_grad_tf.py`` (``predict_and_cost`` :58-87, ``step`` :90-114,
  ``optimizer_reset`` :116-119): every sample takes ONE clipped plain gradient-descent step before it is ranked; u is the first
  element of the refit MEAN (:105), not of the best sample.
* ``CEMBharadhwajOracle``  -- ``Optimizers/optimizer_cem_grad_bharadhwaj_tf.py`` (``predict_and_cost`` :93-123, ``_sample_actions``
  :125-132, ``apply_time_delta`` :134-147, ``step`` :151-178, ``optimizer_reset`` :180-184): the k elites are carried between the
  outer iterations, N-k fresh samples are appended, the whole population takes ONE Keras-Adam step whose moments persist per ROW
  across iterations and ticks (never reset, never shifted), u is the first control of the best sample.
Pinned by tests/test_oracle_golden.py against fixtures produced by the UNMODIFIED reference files."""
from __future__ import annotations

import numpy as np
import torch

from . import spec


class _CEMGradBase:
    def __init__(self, predictor, cost: spec.CostParams, *, mpc_horizon, num_rollouts, cem_outer_it, cem_initial_action_stdev,
                 cem_stdev_min, cem_best_k, learning_rate, gradmax_clip, action_low=-1.0, action_high=1.0, dtype=torch.float32,
                 **_ignored):
        self.dtype = dtype
        self.predictor, self.cost = predictor, cost
        self.H, self.N = int(mpc_horizon), int(num_rollouts)
        self.outer_it = int(cem_outer_it)
        self.init_std = float(np.float32(cem_initial_action_stdev))
        self.std_min = float(np.float32(cem_stdev_min))
        self.k = int(cem_best_k)
        self.lr = float(np.float32(learning_rate)) if dtype == torch.float32 else float(learning_rate)
        self.gradmax_clip = float(np.float32(gradmax_clip))
        self.low, self.high = float(np.float32(action_low)), float(np.float32(action_high))
        self.u = 0.0
        self.last = {}
        self.reset()

    def reset(self, rng=None):
        self.dist_mue = (self.low + self.high) * 0.5 * torch.ones([1, self.H, 1], dtype=self.dtype)
        self.stdev = self.init_std * torch.ones([1, self.H, 1], dtype=self.dtype)
        self.count = 0

    def _cost(self, s, Q):
        rollout = self.predictor.predict_core(s, Q)
        return spec.trajectory_cost(rollout, Q, self.u, self.cost), rollout

    def _clipped_gradient(self, s, Q):
        Q = Q.detach().clone().requires_grad_(True)
        J, _ = self._cost(s, Q)
        (g,) = torch.autograd.grad(J.sum(), Q)
        n = torch.sqrt(torch.sum(g * g, dim=(1, 2), keepdim=True))
        c = torch.tensor(self.gradmax_clip, dtype=self.dtype)
        return g * c / torch.maximum(n, c)

    def _refit(self, Qn, traj_cost):
        best_idx = torch.argsort(traj_cost.detach(), stable=True)[: self.k]
        elite_Q = torch.index_select(Qn, 0, best_idx)
        self.dist_mue = torch.mean(elite_Q, dim=0, keepdim=True)
        mu = torch.mean(elite_Q, dim=0, keepdim=True)
        self.stdev = torch.sqrt(torch.mean((elite_Q - mu) * (elite_Q - mu), dim=0, keepdim=True))
        return best_idx, elite_Q

    def _time_shift(self):
        self.stdev = torch.clamp(self.stdev, self.std_min, 10.0)
        self.stdev = torch.cat([self.stdev[:, 1:, :], self.init_std * torch.ones((1, 1, 1), dtype=self.dtype)], dim=1)
        self.dist_mue = torch.cat([self.dist_mue[:, 1:, :], (self.low + self.high) * 0.5 * torch.ones((1, 1, 1), dtype=self.dtype)], dim=1)


class CEMNaiveGradOracle(_CEMGradBase):
    def step(self, s: np.ndarray, rng) -> np.ndarray:
        s = torch.as_tensor(np.tile(np.asarray(s, np.float32), (self.N, 1))).to(self.dtype)  # :94-95
        elite_log = []
        for _ in range(self.outer_it):  # :98-99 -> predict_and_cost :58-87
            Q = self.dist_mue.repeat(self.N, 1, 1) + rng.normal([self.N, self.H, 1], dtype=torch.float32).to(self.dtype) * self.stdev  # :61-62
            Q = torch.clamp(Q, self.low, self.high)  # :63
            g = self._clipped_gradient(s, Q)  # :65-72
            Qn = torch.clamp(Q - self.lr * g, self.low, self.high)  # :74-75
            traj_cost, rollout = self._cost(s, Qn)  # :77-78
            best_idx, _ = self._refit(Qn, traj_cost)  # :81-86
            elite_log.append(best_idx.numpy().copy())
        self.stdev = torch.clamp(self.stdev, self.std_min, 10.0)  # :103
        self.stdev = torch.cat([self.stdev[:, 1:, :], self.init_std * torch.ones((1, 1, 1), dtype=self.dtype)], dim=1)  # :104
        self.u = self.dist_mue[0, 0, :].squeeze().numpy().astype(np.float32).copy()  # :105 the refit MEAN
        self.dist_mue = torch.cat([self.dist_mue[:, 1:, :], (self.low + self.high) * 0.5 * torch.ones((1, 1, 1), dtype=self.dtype)], dim=1)  # :106
        self.last = dict(J=traj_cost.detach().numpy(), Q=Qn.detach().numpy(), rollouts=rollout.detach().numpy(), elite_idx=np.stack(elite_log))
        self.count += 1
        return self.u


class CEMBharadhwajOracle(_CEMGradBase):
    def __init__(self, predictor, cost, *, adam_beta_1, adam_beta_2, adam_epsilon, warmup=False, warmup_iterations=0, **kw):
        super().__init__(predictor, cost, **kw)
        self.b1, self.b2, self.eps = float(adam_beta_1), float(adam_beta_2), float(adam_epsilon)
        self.warmup, self.warmup_iterations = bool(warmup), int(warmup_iterations)
        self.adam_step, self.m, self.v = 0, None, None  # Keras slots: created by the first apply_gradients, never reset afterwards

    def _sample(self, rng, n):  # :125-132
        return self.dist_mue.repeat(n, 1, 1) + self.stdev * rng.normal([n, self.H, 1], dtype=torch.float32).to(self.dtype)

    def _adam(self, Q, g):  # tf.keras.optimizers.Adam (:64-69), one step on the whole [N,H,1] variable
        if self.m is None:
            self.m, self.v = torch.zeros_like(g), torch.zeros_like(g)
        t = self.adam_step + 1
        f32 = np.float32
        if self.dtype == torch.float32:
            lr_t = float(f32(f32(self.lr) * np.sqrt(f32(1) - np.power(f32(self.b2), f32(t))) / (f32(1) - np.power(f32(self.b1), f32(t)))))
            one_b1, one_b2, eps = float(f32(1 - self.b1)), float(f32(1 - self.b2)), float(f32(self.eps))
        else:
            lr_t = self.lr * np.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)
            one_b1, one_b2, eps = 1 - self.b1, 1 - self.b2, self.eps
        self.m = self.m + (g - self.m) * one_b1
        self.v = self.v + (g * g - self.v) * one_b2
        self.adam_step = t
        return Q - lr_t * self.m / (torch.sqrt(self.v) + eps)

    def step(self, s: np.ndarray, rng) -> np.ndarray:
        s = torch.as_tensor(np.tile(np.asarray(s, np.float32), (self.N, 1))).to(self.dtype)  # :155-156
        elite_Q = self._sample(rng, self.k)  # :159
        iterations = self.warmup_iterations if self.warmup and self.count == 0 else self.outer_it  # :162
        elite_log = []
        for _ in range(iterations):  # :163-164 -> predict_and_cost :93-123
            Q = torch.cat([elite_Q, self._sample(rng, self.N - self.k)], dim=0)  # :95-96
            Q = torch.clamp(Q, self.low, self.high)  # :97
            g = self._clipped_gradient(s, Q)  # :100-108
            Qn = torch.clamp(self._adam(Q, g), self.low, self.high)  # :110-111
            traj_cost, rollout = self._cost(s, Qn)  # :113-114
            best_idx, elite_Q = self._refit(Qn, traj_cost)  # :117-122
            elite_log.append(best_idx.numpy().copy())
        self.u = elite_Q[0, 0, :].squeeze().numpy().astype(np.float32).copy()  # :168 first control of the best sample
        self._time_shift()  # :169 -> :134-147
        self.last = dict(J=traj_cost.detach().numpy(), Q=Qn.detach().numpy(), rollouts=rollout.detach().numpy(), elite_idx=np.stack(elite_log))
        self.count += 1
        return self.u
