"""oracle.mppi -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Standalone torch-CPU fp32 restatement of the reference's MPPI tick, following
``/root/reference/Optimizers/optimizer_mppi.py`` line by line (cited per statement) with the interpolation of
``others/Interpolator.py:53-84,97-106`` and the trajectory cost of ``Cost_Functions/__init__.py:74-93``.
Pinned against the unmodified reference file through tests/golden/mppi_*.npz (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import spec


def interpolation_matrix(horizon: int, period: int, nu: int = 1) -> np.ndarray:
    """reference others/Interpolator.py:53-77 -> W [n_ind, H] (nu folded away: identical for every control)."""
    n_ind = int(math.ceil((horizon - 1) / period) + 1)  # Interpolator.py:79-84
    mat = np.zeros(((n_ind - 1) * period + 1, n_ind), dtype=np.float32)
    block = np.zeros((period, 2), dtype=np.float32)
    for j in range(period):
        block[j, 0] = period - j
        block[j, 1] = j
    for i in range(n_ind - 1):
        mat[i * period:(i + 1) * period, i:i + 2] = block
    mat[-1, -1] = 1
    mat = mat[:horizon, :] / np.float32(period)  # fp32 division, Interpolator.py:74
    return np.ascontiguousarray(mat.T)  # [n_ind, H]


class MPPIOracle:
    def __init__(self, predictor, cost: spec.CostParams, *, mpc_horizon, num_rollouts, cc_weight, R, LBD, NU,
                 SQRTRHOINV, period_interpolation_inducing_points, mpc_timestep=0.02,
                 action_low=-1.0, action_high=1.0, dtype=torch.float32, **_ignored):
        self.dtype = dtype  # torch.float64 gives the exact-arithmetic 'truth' used to measure the fp32 noise floor
        self.predictor = predictor
        self.cost = cost
        self.H = int(mpc_horizon)
        self.N = int(num_rollouts)
        self.cc_weight = torch.tensor(cc_weight, dtype=dtype)  # optimizer_mppi.py:155 (to_tensor)
        self.R = torch.tensor(R, dtype=dtype)  # :92
        self.LBD = float(LBD)  # :93 (python float)
        self.NU = torch.tensor(NU, dtype=dtype)  # :94
        self.period = int(period_interpolation_inducing_points)
        self.W = torch.from_numpy(interpolation_matrix(self.H, self.period)).to(dtype)  # [n_ind, H]
        self.n_ind = self.W.shape[0]
        # :130  SQRTRHODTINV = fp32( SQRTRHOINV * (1/sqrt(dt)) ) evaluated in float64
        self.SQRTRHODTINV = torch.tensor(np.float32(np.array(SQRTRHOINV) * (1 / np.sqrt(mpc_timestep))), dtype=dtype)
        self.low = torch.as_tensor(np.atleast_1d(np.asarray(action_low, np.float32))).to(dtype)  # [nu] (Optimizers/__init__.py:42-44)
        self.high = torch.as_tensor(np.atleast_1d(np.asarray(action_high, np.float32))).to(dtype)
        self.nu = int(getattr(predictor, "num_control_inputs", 1))
        self.u = 0.0  # Optimizers/__init__.py:35 (constructor only)
        self.reset()

    def reset(self):
        # optimizer_mppi.py:227-231: ONLY u_nom is reset; self.u (the previous_input of the cost) keeps the last applied control
        self.u_nom = 0.5 * (self.low + self.high) * torch.ones([1, self.H, self.nu], dtype=self.dtype)
        self.last = {}

    def _interpolate(self, y):  # y [N, n_ind, 1] -> [N, H, 1]; Interpolator.py:97-106 (batched matmul per control)
        return torch.matmul(y.permute(2, 0, 1), self.W[None]).permute(1, 2, 0)

    def step(self, s: np.ndarray, rng) -> np.ndarray:
        s = torch.as_tensor(np.asarray(s, np.float32)).to(self.dtype).reshape(1, -1)  # :208-209
        u_old = self.u
        s = s.repeat(self.N, 1)  # :182
        u_nom = torch.cat([self.u_nom[:, 1:, :], self.u_nom[:, -1:, :]], 1)  # :184
        delta_u = rng.normal([self.N, self.n_ind, self.nu], dtype=torch.float32).to(self.dtype) * self.SQRTRHODTINV  # :173-175
        delta_u = self._interpolate(delta_u)  # :177
        u_run = u_nom.repeat(self.N, 1, 1) + delta_u  # :186
        u_run = torch.minimum(torch.maximum(u_run, self.low), self.high)  # :187
        rollout = self.predictor.predict_core(s, u_run)  # :188
        total = spec.trajectory_cost(rollout, u_run, u_old, self.cost)  # :159
        # :154-155 mppi_correction_cost (evaluation order as written)
        corr = torch.sum(self.cc_weight * (0.5 * (1 - 1.0 / self.NU) * self.R * (delta_u ** 2)
                                           + self.R * u_run * delta_u + 0.5 * self.R * (u_run ** 2)), (1, 2))
        S = total + corr  # :160
        rho = torch.amin(S, 0)  # :164
        exp_s = torch.exp(-1.0 / self.LBD * (S - rho))  # :165
        a = torch.sum(exp_s, 0)  # :166
        b = torch.sum(exp_s[:, None, None] * delta_u, 0) / a  # :167
        u_nom = torch.minimum(torch.maximum(u_nom + b, self.low), self.high)  # :190
        self.u_nom = u_nom
        self.u = u_nom[0, 0, :].squeeze().numpy().copy()  # :191,212
        if hasattr(self.predictor, "update"):  # :192,195-197 RNN-state hook: the saved hidden state advances with (s, new u_nom[0])
            self.predictor.update(s=s, Q0=u_nom[:, :1, :].repeat(self.N, 1, 1))
        self.last = dict(J=S.numpy(), Q=u_run.numpy(), rollouts=rollout.numpy(), delta_u=delta_u.numpy())
        return self.u

    def step_chunked(self, s: np.ndarray, rng, chunk: int = 65536) -> np.ndarray:
        """The same tick evaluated ``chunk`` rollouts at a time (full-size configs: 10^6 x 100 would need > 5 GB of
        intermediates in one piece).  Identical arithmetic per rollout; the population sums ``a`` and ``b`` are accumulated per chunk
        (fp32 partial sums, then summed), i.e. a different -- equally valid -- fp32 summation tree than torch.sum's over N.
        tests/test_oracle_golden.py checks it against step()."""
        s1 = torch.as_tensor(np.asarray(s, np.float32)).to(self.dtype).reshape(1, -1)
        u_old = self.u
        u_nom = torch.cat([self.u_nom[:, 1:, :], self.u_nom[:, -1:, :]], 1)
        z_all = rng.normal([self.N, self.n_ind, self.nu], dtype=torch.float32).to(self.dtype)
        S_parts = []
        for c0 in range(0, self.N, chunk):
            z = z_all[c0:c0 + chunk]
            n = z.shape[0]
            delta_u = self._interpolate(z * self.SQRTRHODTINV)
            u_run = torch.minimum(torch.maximum(u_nom.repeat(n, 1, 1) + delta_u, self.low), self.high)
            rollout = self.predictor.predict_core(s1.repeat(n, 1), u_run)
            total = spec.trajectory_cost(rollout, u_run, u_old, self.cost)
            corr = torch.sum(self.cc_weight * (0.5 * (1 - 1.0 / self.NU) * self.R * (delta_u ** 2)
                                               + self.R * u_run * delta_u + 0.5 * self.R * (u_run ** 2)), (1, 2))
            S_parts.append(total + corr)
        S = torch.cat(S_parts)
        rho = torch.amin(S, 0)
        exp_s = torch.exp(-1.0 / self.LBD * (S - rho))
        a_parts, b_parts = [], []
        for c0 in range(0, self.N, chunk):
            e = exp_s[c0:c0 + chunk]
            delta_u = self._interpolate(z_all[c0:c0 + chunk] * self.SQRTRHODTINV)
            a_parts.append(torch.sum(e, 0))
            b_parts.append(torch.sum(e[:, None, None] * delta_u, 0))
        a = torch.sum(torch.stack(a_parts), 0)
        b = torch.sum(torch.stack(b_parts), 0) / a
        u_nom = torch.minimum(torch.maximum(u_nom + b, self.low), self.high)
        self.u_nom = u_nom
        self.u = u_nom[0, 0, :].squeeze().numpy().copy()
        self.last = dict(J=S.numpy())
        return self.u
