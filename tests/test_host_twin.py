"""CPU tests of the per-rollout arithmetic the kernels run (control_toolkit_b200/csrc/ctk_math.cuh), compiled for the
host as a TEST-ONLY twin (tests/host_twin): ODE step + cost vs the oracle spec, and the hand-derived RPGD adjoint vs
the oracle's torch autograd.  The twin is not a product path -- the product never loads it."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import spec

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def twin():
    sys.path.insert(0, os.path.join(HERE, "host_twin"))
    import build_twin
    lib = C.CDLL(build_twin.build())
    return lib


def _params(cost_name, dt=0.02):
    from control_toolkit_b200 import specs, _lib as L
    ode = specs.CartPoleODE().to_c(dt)
    cost = specs.resolve_cost("CartPole", cost_name).to_c(0.0, 1.0)
    return ode, cost


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.mark.parametrize("cost_name", ["default", "quadratic_boundary_grad"])
def test_rollout_and_cost_match_spec(twin, cost_name):
    N, H = 64, 50
    rng = np.random.default_rng(3)
    for s0 in spec.synthetic_states(4, seed=5):
        Q = rng.uniform(-1, 1, (N, H)).astype(np.float32)
        ode, cost = _params(cost_name)
        J = np.zeros(N, np.float32)
        traj = np.zeros((N, H + 1, 6), np.float32)
        twin.twin_rollout_cost(_fp(s0), _fp(Q), N, H, C.byref(ode), C.byref(cost), C.c_float(0.3), _fp(J), _fp(traj))
        pred = spec.ODEPredictor()
        ro = pred.predict_core(torch.from_numpy(np.tile(s0, (N, 1))), torch.from_numpy(Q[..., None]))
        Jo = spec.trajectory_cost(ro, torch.from_numpy(Q[..., None]), 0.3, spec.CostParams(name=cost_name)).numpy()
        for c in range(6):
            scale = max(np.abs(ro[..., c].numpy()).max(), 1e-6)
            assert np.abs(traj[..., c] - ro[..., c].numpy()).max() / scale < 2e-4
        assert np.max(np.abs(J - Jo) / (np.abs(Jo) + 1e-3)) < 1e-4


@pytest.mark.parametrize("cost_name", ["default", "quadratic_boundary_grad"])
def test_adjoint_matches_autograd(twin, cost_name):
    N, H = 32, 50
    rng = np.random.default_rng(4)
    for s0 in spec.synthetic_states(4, seed=6):
        Q = rng.uniform(-1, 1, (N, H)).astype(np.float32)
        ode, cost = _params(cost_name)
        g = np.zeros((N, H), np.float32)
        twin.twin_grad(_fp(s0), _fp(Q), N, H, C.byref(ode), C.byref(cost), C.c_float(-0.2), _fp(g))
        Qt = torch.from_numpy(Q[..., None]).clone().requires_grad_(True)
        ro = spec.ODEPredictor().predict_core(torch.from_numpy(np.tile(s0, (N, 1))), Qt)
        spec.trajectory_cost(ro, Qt, -0.2, spec.CostParams(name=cost_name)).sum().backward()
        go = Qt.grad[..., 0].numpy()
        for n in range(N):
            scale = np.abs(go[n]).max()
            assert np.abs(g[n] - go[n]).max() / scale < 2e-4, (n, np.abs(g[n] - go[n]).max(), scale)


@pytest.mark.parametrize("cost_name", ["default", "quadratic_boundary_grad"])
def test_coefficient_form_adjoint_matches_autograd_and_direct_form(twin, cost_name):
    """The reverse sweep the device runs (adjoint_coefficients folded into the forward pass + adjoint_apply) against torch
    autograd (same bound as the direct form) and against the direct form itself (pure re-association: 2e-5 of the row scale)."""
    N, H = 32, 50
    rng = np.random.default_rng(9)
    for s0 in spec.synthetic_states(4, seed=8):
        Q = rng.uniform(-1, 1, (N, H)).astype(np.float32)
        ode, cost = _params(cost_name)
        g, gd = np.zeros((N, H), np.float32), np.zeros((N, H), np.float32)
        twin.twin_grad_coef(_fp(s0), _fp(Q), N, H, C.byref(ode), C.byref(cost), C.c_float(-0.2), _fp(g))
        twin.twin_grad(_fp(s0), _fp(Q), N, H, C.byref(ode), C.byref(cost), C.c_float(-0.2), _fp(gd))
        Qt = torch.from_numpy(Q[..., None]).clone().requires_grad_(True)
        ro = spec.ODEPredictor().predict_core(torch.from_numpy(np.tile(s0, (N, 1))), Qt)
        spec.trajectory_cost(ro, Qt, -0.2, spec.CostParams(name=cost_name)).sum().backward()
        go = Qt.grad[..., 0].numpy()
        for n in range(N):
            scale = np.abs(go[n]).max()
            assert np.abs(g[n] - go[n]).max() / scale < 2e-4, (n, np.abs(g[n] - go[n]).max(), scale)
            assert np.abs(g[n] - gd[n]).max() / scale < 2e-5, (n, np.abs(g[n] - gd[n]).max(), scale)


@pytest.mark.parametrize("cost_name", ["default", "quadratic_boundary_grad"])
def test_scaled_variable_rollout_matches_spec(twin, cost_name):
    """The MPPI/ODE kernel's arithmetic (ctk_ode_scaled.cuh: scaled state variables, merged control terms, telescoped
    control-change cost) and derive_ode_hot against the oracle spec: trajectories and the total MPPI cost
    S = trajectory cost + sum_t cc (0.5 (1 - 1/NU) R du^2 + R u du + 0.5 R u^2)   (optimizer_mppi.py:154-161)."""
    N, H = 64, 100
    rng = np.random.default_rng(9)
    f = np.float32
    cc, R, NU = f(1.0), f(1.0), f(1000.0)
    coef_du2 = f(f(0.5) * (f(1) - f(1.0) / NU)) * R
    for s0 in spec.synthetic_states(4, seed=11):
        dU = (rng.standard_normal((N, H)) * 0.2).astype(np.float32)
        U = np.clip(rng.uniform(-0.5, 0.5, (1, H)).astype(np.float32) + dU, -1, 1).astype(np.float32)
        ode, cost = _params(cost_name)
        S = np.zeros(N, np.float32)
        traj = np.zeros((N, H + 1, 6), np.float32)
        twin.twin_rollout_scaled(_fp(s0), _fp(U), _fp(dU), N, H, C.byref(ode), C.byref(cost), C.c_float(0.3), C.c_float(cc),
                                 C.c_float(coef_du2), C.c_float(R), C.c_float(f(0.5) * R), _fp(S), _fp(traj))
        ro = spec.ODEPredictor().predict_core(torch.from_numpy(np.tile(s0, (N, 1))), torch.from_numpy(U[..., None]))
        Jo = spec.trajectory_cost(ro, torch.from_numpy(U[..., None]), 0.3, spec.CostParams(name=cost_name)).numpy().astype(np.float64)
        corr = (cc * (0.5 * (1 - 1.0 / 1000.0) * R * dU.astype(np.float64) ** 2 + R * U.astype(np.float64) * dU + 0.5 * R * U.astype(np.float64) ** 2)).sum(1)
        So = Jo + corr
        for c in range(6):
            scale = max(np.abs(ro[..., c].numpy()).max(), 1e-6)
            assert np.abs(traj[..., c] - ro[..., c].numpy()).max() / scale < 4e-4, c
        assert np.max(np.abs(S - So) / (np.abs(So) + 1e-3)) < 2e-4
