"""oracle.rpgd -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Standalone torch-CPU fp32 restatement of the reference's RPGD tick, following
``/root/reference/Optimizers/optimizer_rpgd.py`` (torch branch: ``_grad_step_torch`` :329-338, manual Adam :56-82,
``_get_action`` :340-380, ``step`` :388-524, ``optimizer_reset`` :527-548) with TensorFlow *value* semantics for
``u_nom`` (copy, not a view of ``Q_tf``: see oracle/refharness/SI_Toolkit/computation_library.py).
``adam_form="keras"`` switches the update to the Keras/TF form.
Pinned against the unmodified reference file via tests/golden/rpgd_*.npz.
"""
from __future__ import annotations

import numpy as np
import torch

from . import spec
from .mppi import interpolation_matrix


class RPGDOracle:
    def __init__(self, predictor, cost: spec.CostParams, *, mpc_horizon, num_rollouts, outer_its, sample_stdev,
                 sample_mean, sample_whole_control_space, uniform_dist_min, uniform_dist_max, resamp_per,
                 period_interpolation_inducing_points, SAMPLING_DISTRIBUTION, shift_previous, warmup,
                 warmup_iterations, learning_rate, opt_keep_k_ratio, gradmax_clip, adam_beta_1, adam_beta_2,
                 adam_epsilon, action_low=-1.0, action_high=1.0, adam_form="torch", dtype=torch.float32, **_ignored):
        self.dtype = dtype
        self.predictor, self.cost = predictor, cost
        self.H, self.N = int(mpc_horizon), int(num_rollouts)
        self.outer_its = int(outer_its)
        self.low = torch.tensor([action_low], dtype=dtype)
        self.high = torch.tensor([action_high], dtype=dtype)
        self.sample_stdev = torch.tensor(float(np.float32(sample_stdev)), dtype=dtype)
        self.sample_mean = torch.tensor(float(np.float32(sample_mean)), dtype=dtype)
        if sample_whole_control_space:  # :200-205
            self.sample_min, self.sample_max = self.low.clone(), self.high.clone()
        else:
            self.sample_min = torch.tensor(float(np.float32(uniform_dist_min)), dtype=dtype)
            self.sample_max = torch.tensor(float(np.float32(uniform_dist_max)), dtype=dtype)
        self.resamp_per = int(resamp_per)
        self.shift_previous = int(shift_previous)
        self.first_iter_count = int(warmup_iterations) if warmup else self.outer_its  # :219-221
        self.k = int(max(int(num_rollouts * opt_keep_k_ratio), 1))  # :213
        self.gradmax_clip = torch.tensor(float(np.float32(gradmax_clip)), dtype=dtype)
        self.dist = SAMPLING_DISTRIBUTION
        self.lr, self.b1, self.b2, self.eps = learning_rate, adam_beta_1, adam_beta_2, adam_epsilon
        self.adam_form = adam_form
        self.W = torch.from_numpy(interpolation_matrix(self.H, int(period_interpolation_inducing_points))).to(dtype)
        self.n_ind = self.W.shape[0]
        self.Q = None
        self.u = np.float32(0.0)  # Optimizers/__init__.py:35 (constructor only; optimizer_reset :527-548 keeps the last applied control)

    # -- :275-296 -----------------------------------------------------------------------------------
    def sample_actions(self, rng, batch):
        if self.dist == "normal":
            Qn = rng.normal([batch, self.n_ind, 1], mean=self.sample_mean.float(), stddev=self.sample_stdev.float(), dtype=torch.float32).to(self.dtype)
        elif self.dist == "uniform":
            Qn = rng.uniform([batch, self.n_ind, 1], minval=self.sample_min.float(), maxval=self.sample_max.float(), dtype=torch.float32).to(self.dtype)
        else:
            raise ValueError(f"RPGD cannot interpret sampling type {self.dist}")
        Qn = torch.minimum(torch.maximum(Qn, self.low), self.high)
        return torch.matmul(Qn.permute(2, 0, 1), self.W[None]).permute(1, 2, 0)

    def reset(self, rng):  # :527-548
        self.Q = self.sample_actions(rng, self.N).clone()
        self.count = 0
        self.adam_step, self.m, self.v = 0, None, None
        self.ages = torch.zeros((self.N,), dtype=self.dtype)
        self.last = {}

    def _cost(self, s, Q):  # :298-304
        rollout = self.predictor.predict_core(s, Q)
        return spec.trajectory_cost(rollout, Q, self.u, self.cost), rollout

    def _adam(self, g, Q):  # :56-82 (torch) or Keras form
        self.adam_step += 1
        if self.m is None:
            self.m, self.v = torch.zeros_like(g), torch.zeros_like(g)
        if self.adam_form == "torch":
            m = self.m.mul(self.b1).add(g, alpha=1 - self.b1)
            v = self.v.mul(self.b2).add(g * g, alpha=1 - self.b2)
            self.m, self.v = m, v
            bc1 = 1 - self.b1 ** self.adam_step
            bc2 = 1 - self.b2 ** self.adam_step
            return Q - self.lr * m.div(bc1) / (v.div(bc2).sqrt() + self.eps)
        # Keras: m += (g-m)(1-b1); v += (g^2-v)(1-b2); var -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)
        self.m = self.m + (g - self.m) * (1 - self.b1)
        self.v = self.v + (g * g - self.v) * (1 - self.b2)
        alpha = self.lr * np.sqrt(1 - self.b2 ** self.adam_step) / (1 - self.b1 ** self.adam_step)
        return Q - alpha * self.m / (self.v.sqrt() + self.eps)

    def grad_step(self, s, Q):  # :329-338
        Q = Q.clone().detach().requires_grad_(True)
        traj_cost, _ = self._cost(s, Q)
        traj_cost.sum().backward()
        g = Q.grad
        l2 = torch.sqrt(torch.sum(g * g, dim=(1, 2), keepdim=True))  # clip_by_norm(axes=[1,2])
        g = g * self.gradmax_clip / torch.maximum(l2, self.gradmax_clip)
        Qn = self._adam(g, Q.detach())
        return torch.minimum(torch.maximum(Qn, self.low), self.high), traj_cost.detach()

    def step(self, s: np.ndarray, rng) -> np.ndarray:
        s = torch.as_tensor(np.tile(np.asarray(s, np.float32), (self.N, 1))).to(self.dtype)  # :393-394
        iters = self.first_iter_count if self.count == 0 else self.outer_its  # :397-400
        for _ in range(iters):  # :404-406
            self.Q, _ = self.grad_step(s, self.Q)
        with torch.no_grad():
            J, rollout = self._cost(s, self.Q)  # _get_action :342
            sorted_cost = torch.argsort(J, dim=0, stable=True)  # :345
            best_idx = sorted_cost[: self.k]  # :346
            sp = self.shift_previous
            Qn = torch.cat([self.Q[:, sp:, :], self.Q[:, -1:, :].repeat(1, sp, 1)], dim=1)  # :376-379
            u_nom = self.Q[best_idx[0]][None].clone()  # :426 (value semantics)
            Q_logged = self.Q.clone()
            ages_logged = self.ages.clone()
            m_t, v_t = self.m, self.v
            if self.count % self.resamp_per == 0:  # :449-495
                Qres = self.sample_actions(rng, self.N - self.k)
                Qn = torch.cat([Qres, torch.index_select(Qn, 0, best_idx)], 0)
                self.ages = torch.cat([torch.zeros((self.N - self.k,), dtype=self.dtype), torch.index_select(self.ages, 0, best_idx)], 0)
                wk1 = torch.cat([torch.index_select(m_t, 0, best_idx)[:, 1:, :], torch.zeros([self.k, 1, 1], dtype=self.dtype)], 1)
                wk2 = torch.cat([torch.index_select(v_t, 0, best_idx)[:, 1:, :], torch.zeros([self.k, 1, 1], dtype=self.dtype)], 1)
                z = torch.zeros([self.N - self.k, self.H, 1], dtype=self.dtype)
                self.m, self.v = torch.cat([z, wk1], 0), torch.cat([z, wk2], 0)
            else:  # :496-513
                self.m = torch.cat([m_t[:, 1:, :], torch.zeros([self.N, 1, 1], dtype=self.dtype)], 1)
                self.v = torch.cat([v_t[:, 1:, :], torch.zeros([self.N, 1, 1], dtype=self.dtype)], 1)
            self.ages = self.ages + 1  # :514
            self.Q = Qn.clone()  # :515
            self.count += 1  # :516
        self.u = u_nom[0, 0, :].numpy().copy()  # :523
        self.last = dict(J=J.numpy(), Q=Q_logged.numpy(), rollouts=rollout.numpy(), ages=ages_logged.numpy(),
                         u_nom=u_nom.numpy(), best_idx=best_idx.numpy())
        return self.u
