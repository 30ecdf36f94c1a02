"""oracle.gradient -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatement (torch-CPU fp32) of the reference's population gradient optimizer ``Optimizers/optimizer_gradient_tf.py``
(``gradient_optimization`` :82-98, ``step`` :101-167, ``optimizer_reset`` :169-185) with TensorFlow semantics restated:
``tf.clip_by_norm(g, c, axes=[1,2])`` = g c / max(||g||_2, c) per trajectory; legacy Keras Adam (lr_t = lr sqrt(1-b2^t)/(1-b1^t),
m += (g-m)(1-b1), v += (g^2-v)(1-b2), Q -= lr_t m / (sqrt(v)+eps)); ``tf.argsort`` ascending, ties to the lower index.
Pinned by tests/test_oracle_golden.py against fixtures produced by the UNMODIFIED reference file."""
from __future__ import annotations

import numpy as np
import torch

from . import spec


class GradientOracle:
    def __init__(self, predictor, cost: spec.CostParams, *, mpc_horizon, num_rollouts, gradient_steps, learning_rate, adam_beta_1,
                 adam_beta_2, adam_epsilon, gradmax_clip, warmup, warmup_iterations, action_low=-1.0, action_high=1.0,
                 dtype=torch.float32, **_ignored):
        self.dtype = dtype
        self.predictor, self.cost = predictor, cost
        self.H, self.N = int(mpc_horizon), int(num_rollouts)
        self.gradient_steps = int(gradient_steps)
        self.first_iter_count = int(warmup_iterations) if warmup else self.gradient_steps  # :66-68
        self.low, self.high = np.float32(action_low), np.float32(action_high)
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), float(adam_beta_1), float(adam_beta_2), float(adam_epsilon)
        self.gradmax_clip = float(np.float32(gradmax_clip))
        self.u = 0.0
        self.Q = None
        self.last = {}

    def reset(self, rng):  # :169-185
        Q = rng.uniform([self.N, self.H, 1], self.low, self.high, dtype=torch.float32).to(self.dtype)
        self.Q = torch.clamp(Q, float(self.low), float(self.high))
        self.count = 0
        self.adam_step, self.m, self.v = 0, None, None  # the Keras slots do not exist before the first apply_gradients

    def _cost(self, s, Q):  # :70-77
        rollout = self.predictor.predict_core(s, Q)
        return spec.trajectory_cost(rollout, Q, self.u, self.cost), rollout

    def _grad_step(self, s):  # :82-98
        Q = self.Q.detach().clone().requires_grad_(True)
        J, _ = self._cost(s, Q)
        (g,) = torch.autograd.grad(J.sum(), Q)
        n = torch.sqrt(torch.sum(g * g, dim=(1, 2), keepdim=True))
        c = torch.tensor(self.gradmax_clip, dtype=self.dtype)
        g = g * c / torch.maximum(n, c)
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
        Qn = Q.detach() - lr_t * self.m / (torch.sqrt(self.v) + eps)
        self.adam_step = t
        self.Q = torch.clamp(Qn, float(self.low), float(self.high))  # :94, assigned at :123

    def step(self, s: np.ndarray, rng) -> np.ndarray:
        s = torch.as_tensor(np.tile(np.asarray(s, np.float32), (self.N, 1))).to(self.dtype)  # :105-106
        iters = self.first_iter_count if self.count == 0 else self.gradient_steps  # :109-112
        for _ in range(iters):  # :116-118
            self._grad_step(s)
        Q = self.Q
        traj_cost, rollout = self._cost(s, Q)  # :126
        best_idx = int(torch.argsort(traj_cost.detach(), stable=True)[0])  # :129-130
        self.u = Q[best_idx, 0, :].squeeze().detach().numpy().astype(np.float32)  # :132
        self.last = dict(J=traj_cost.detach().numpy(), Q=Q.detach().numpy().copy(), rollouts=rollout.detach().numpy(), best_idx=best_idx)
        self.count += 1  # :141
        Q_s = rng.uniform(shape=[self.N, 1, 1], minval=self.low, maxval=self.high, dtype=torch.float32).to(self.dtype)  # :142-147
        self.Q = torch.cat([self.Q[:, 1:, :], Q_s], dim=1)  # :148-149
        if self.m is not None:  # :151-165  Adam moments shifted by one step, zero fill
            z = torch.zeros((self.N, 1, 1), dtype=self.dtype)
            self.m = torch.cat([self.m[:, 1:, :], z], dim=1)
            self.v = torch.cat([self.v[:, 1:, :], z], dim=1)
        return self.u
