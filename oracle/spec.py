"""oracle.spec -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The build's pinned written specification of the arithmetic the reference delegates to code that is
NOT under /root/reference (**parity unpinned**):

* the CartPole next-state ODE behind ``PredictorWrapper.predict_core``
  (call sites: reference ``Optimizers/optimizer_mppi.py:188``, ``optimizer_cem_tf.py:57``,
  ``optimizer_rpgd.py:300``),
* the CartPole cost classes loaded by ``Cost_Functions/cost_function_wrapper.py:59-66``
  (``default`` for MPPI/CEM, ``quadratic_boundary_grad`` for RPGD), evaluated through
  ``Cost_Functions/__init__.py:49-93`` (stage cost shift, mean over H+1),
* the 6->128->128->5 tanh MLP autoregressive predictor of config C4.

Everything is fp32 torch-CPU, written one elementary operation at a time in the order the formulas
are stated, so the CUDA kernels can follow the same op order.  All *compound constants* are evaluated
in float64 on the host and rounded once to fp32 (``CartPoleParams.f32``); the CUDA side receives the
same rounded constants through the C-ABI.

State layout (alphabetical, upstream convention): [angle, angleD, angle_cos, angle_sin, position, positionD].
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

ANGLE, ANGLED, ANGLE_COS, ANGLE_SIN, POSITION, POSITIOND = range(6)
NUM_STATES = 6
NUM_CONTROLS = 1


def _f32(x: float) -> float:
    """Round a python double once to fp32 and return it as a python float."""
    return float(np.float32(x))


@dataclass
class CartPoleParams:
    # physical constants [UPSTREAM-RECALL, pinned here]
    k: float = 1.0 / 3.0
    M: float = 0.230
    m: float = 0.087
    L: float = 0.395 / 2.0
    g: float = 9.81
    J_fric: float = 2.5e-4
    M_fric: float = 4.77
    u_max: float = 2.62
    TrackHalfLength: float = 0.198
    # integration
    dt: float = 0.02
    intermediate_steps: int = 1

    def f32(self) -> dict:
        """Compound constants, float64 -> one rounding to fp32.  Same list as include/ctk_b200.h::ctk_ode_params."""
        h = self.dt / self.intermediate_steps
        return dict(
            u_max=_f32(self.u_max),
            kp1_Mm=_f32((self.k + 1.0) * (self.M + self.m)),
            m=_f32(self.m),
            neg_M_fric=_f32(-self.M_fric),
            neg_J_fric=_f32(-self.J_fric),
            mg=_f32(self.m * self.g),
            L=_f32(self.L),
            kp1=_f32(self.k + 1.0),
            mL=_f32(self.m * self.L),
            g=_f32(self.g),
            kp1L=_f32((self.k + 1.0) * self.L),
            h=_f32(h),
        )


@dataclass
class CostParams:
    """Weights of the CartPole cost classes.  quadratic_boundary_grad values are the ones visible at
    reference ``Control_Toolkit_ASF_Template/config_cost_function.yml:11-18``; ``default`` reuses them
    (its own are not visible anywhere in the reference)."""
    name: str = "default"  # "default" | "quadratic_boundary_grad"
    dd_weight: float = 600.0
    ep_weight: float = 20000.0
    ekp_weight: float = 80.0  # quadratic_boundary_grad only
    cc_weight: float = 1.0
    ccrc_weight: float = 1.0
    R: float = 1.0
    MAX_COST: float = 0.0  # reference Cost_Functions/__init__.py:13-15,63-64
    TrackHalfLength: float = 0.198
    target_position: float = 0.0
    target_equilibrium: float = 1.0

    def f32(self) -> dict:
        thl = self.TrackHalfLength
        return dict(
            dd_weight=_f32(self.dd_weight),
            ep_weight=_f32(self.ep_weight),
            ekp_weight=_f32(self.ekp_weight),
            cc_weight=_f32(self.cc_weight),
            ccrc_weight=_f32(self.ccrc_weight),
            R=_f32(self.R),
            MAX_COST=_f32(self.MAX_COST),
            two_thl=_f32(2.0 * thl),
            thl_095=_f32(0.95 * thl),
            thl_005=_f32(0.05 * thl),
            thl_09=_f32(0.9 * thl),
            thl_01=_f32(0.1 * thl),
            target_position=_f32(self.target_position),
            target_equilibrium=_f32(self.target_equilibrium),
        )


# ----------------------------------------------------------------------------------------------
# CartPole ODE (explicit Euler on (angle, angleD, position, positionD) with the OLD derivatives)
# ----------------------------------------------------------------------------------------------
def cartpole_step(s: torch.Tensor, Q: torch.Tensor, c: dict, intermediate_steps: int = 1) -> torch.Tensor:
    """One predictor step.  s [N,6] fp32, Q [N,1] fp32 in [-1,1]; returns the next state [N,6]."""
    angle, angleD, ca, sa, pos, posD = s.unbind(dim=1)
    u = c["u_max"] * Q[:, 0]
    h = c["h"]
    for _ in range(intermediate_steps):
        A = c["kp1_Mm"] - c["m"] * (ca * ca)
        F = c["neg_M_fric"] * posD
        T = c["neg_J_fric"] * angleD
        posDD = (
            c["mg"] * sa * ca
            + (T * ca) / c["L"]
            + c["kp1"] * (-(c["mL"] * (angleD * angleD) * sa) + F + u)
        ) / A
        angleDD = (c["g"] * sa + posDD * ca + T / c["mL"]) / c["kp1L"]
        angle = angle + angleD * h
        angleD = angleD + angleDD * h
        pos = pos + posD * h
        posD = posD + posDD * h
        ca = torch.cos(angle)
        sa = torch.sin(angle)
        angle = torch.atan2(sa, ca)  # wrap to (-pi, pi]
    return torch.stack([angle, angleD, ca, sa, pos, posD], dim=1)


class ODEPredictor:
    """``predict_core(s[N,6], Q[N,H,1]) -> [N,H+1,6]`` including s_0 (shape pinned by reference
    ``optimizer_cem_tf.py:70``)."""

    num_states = NUM_STATES
    num_control_inputs = NUM_CONTROLS

    def __init__(self, params: CartPoleParams | None = None):
        self.params = params or CartPoleParams()
        self.c = self.params.f32()

    def predict_core(self, s: torch.Tensor, Q: torch.Tensor) -> torch.Tensor:
        out = [s]
        for t in range(Q.shape[1]):
            s = cartpole_step(s, Q[:, t, :], self.c, self.params.intermediate_steps)
            out.append(s)
        return torch.stack(out, dim=1)


# ----------------------------------------------------------------------------------------------
# MLP autoregressive predictor (config C4)
# ----------------------------------------------------------------------------------------------
@dataclass
class MLPWeights:
    """Dense 6->H1 tanh -> H2 tanh -> 5 linear.  Row-major [in, out] matrices (y = x @ W + b)."""
    W1: np.ndarray
    b1: np.ndarray
    W2: np.ndarray
    b2: np.ndarray
    W3: np.ndarray
    b3: np.ndarray

    @staticmethod
    def random_init(seed: int = 2, hidden: int = 128, gain: float = 1.0) -> "MLPWeights":
        """weights N(0, 1/fan_in), biases 0, default_rng(2)."""
        rng = np.random.default_rng(seed)

        def w(i, o):
            return (rng.standard_normal((i, o)) * gain / math.sqrt(i)).astype(np.float32)

        return MLPWeights(
            W1=w(6, hidden), b1=np.zeros(hidden, np.float32),
            W2=w(hidden, hidden), b2=np.zeros(hidden, np.float32),
            W3=w(hidden, 5), b3=np.zeros(5, np.float32),
        )


class MLPPredictor:
    """net input  = [Q, angleD, angle_cos, angle_sin, position, positionD]   (normalisation = identity)
    net output = next [angleD, angle_cos, angle_sin, position, positionD]; angle = atan2(sin, cos)."""

    num_states = NUM_STATES
    num_control_inputs = NUM_CONTROLS

    def __init__(self, weights: MLPWeights, dtype=torch.float32, bf16_layer2: bool = False):
        """bf16_layer2: the input rounding of the opt-in 'tcgen05_bf16' / 'tcgen05_fast' engines -- the operands of the 128 x 128
        layer (h1 and W2) are rounded to bfloat16 (round to nearest even), products and sums stay fp32 (SURVEY section 7 hard part 4:
        'the MLP parity is defined against an oracle that applies the same input rounding')."""
        self.w = weights
        self.bf16_layer2 = bool(bf16_layer2)
        self.t = {k: torch.from_numpy(np.ascontiguousarray(getattr(weights, k))).to(dtype) for k in ("W1", "b1", "W2", "b2", "W3", "b3")}
        if self.bf16_layer2:
            self.t["W2"] = self.t["W2"].to(torch.bfloat16).to(dtype)

    def step(self, s: torch.Tensor, Q: torch.Tensor) -> torch.Tensor:
        x = torch.cat([Q, s[:, 1:]], dim=1)  # [N,6]
        h1 = torch.tanh(x @ self.t["W1"] + self.t["b1"])
        if self.bf16_layer2:
            h1 = h1.to(torch.bfloat16).to(h1.dtype)
        h2 = torch.tanh(h1 @ self.t["W2"] + self.t["b2"])
        y = h2 @ self.t["W3"] + self.t["b3"]  # [N,5]
        angle = torch.atan2(y[:, 2], y[:, 1])
        return torch.cat([angle[:, None], y], dim=1)

    def predict_core(self, s: torch.Tensor, Q: torch.Tensor) -> torch.Tensor:
        out = [s]
        for t in range(Q.shape[1]):
            s = self.step(s, Q[:, t, :])
            out.append(s)
        return torch.stack(out, dim=1)


# ----------------------------------------------------------------------------------------------
# GRU autoregressive predictor (recurrent predictors + the predictor.update state hook,
# reference Optimizers/optimizer_mppi.py:195-197)
# ----------------------------------------------------------------------------------------------
@dataclass
class GRUWeights:
    """Two stacked GRU layers of width ``hidden`` and a linear read-out: 6 -> GRU(hidden) -> GRU(hidden) -> Dense 5
    (naming pattern 'GRU-6IN-32H1-32H2-5OUT-0', cf. reference Control_Toolkit_ASF_Template/config_controllers.yml:8).
    Row-major [in, 3*hidden] matrices (g = x @ W + b), gate order [r, z, n]; per layer
        r = sigmoid(gi_r + gh_r)   z = sigmoid(gi_z + gh_z)   n = tanh(gi_n + r * gh_n)   h' = (1 - z) * n + z * h
    with gi = x @ Wi + bi and gh = h @ Wh + bh (the torch.nn.GRU / Keras reset_after=True cell)."""
    Wi1: np.ndarray
    Wh1: np.ndarray
    bi1: np.ndarray
    bh1: np.ndarray
    Wi2: np.ndarray
    Wh2: np.ndarray
    bi2: np.ndarray
    bh2: np.ndarray
    W3: np.ndarray
    b3: np.ndarray

    @property
    def hidden(self) -> int:
        return int(self.Wh1.shape[0])

    @staticmethod
    def random_init(seed: int = 3, hidden: int = 32) -> "GRUWeights":
        """weights N(0, 1/fan_in), biases N(0, 0.1^2) (non-zero so that every bias term is exercised), default_rng(seed); the
        read-out column / bias of the position output are scaled by 0.08 so that the predicted cart mostly stays on the track (0.198 m)
        and the softmin / elite selection see a spread of ordinary costs instead of one rollout that escapes the 1e9 barrier term."""
        rng = np.random.default_rng(seed)

        def w(i, o):
            return (rng.standard_normal((i, o)) / math.sqrt(i)).astype(np.float32)

        def b(o):
            return (0.1 * rng.standard_normal(o)).astype(np.float32)

        h = hidden
        g = GRUWeights(Wi1=w(6, 3 * h), Wh1=w(h, 3 * h), bi1=b(3 * h), bh1=b(3 * h),
                       Wi2=w(h, 3 * h), Wh2=w(h, 3 * h), bi2=b(3 * h), bh2=b(3 * h), W3=w(h, 5), b3=b(5))
        g.W3[:, 3] *= np.float32(0.08)
        g.b3[3] *= np.float32(0.08)
        return g


class GRUPredictor:
    """net input [Q, angleD, angle_cos, angle_sin, position, positionD] (normalisation = identity) -> next
    [angleD, angle_cos, angle_sin, position, positionD]; angle = atan2(sin, cos).

    Stateful like SI_Toolkit's autoregressive RNN predictor: ``predict_core`` starts every rollout from the SAVED hidden state and
    leaves it untouched; ``update(s, Q0)`` advances the saved state by one step with the measured state and the control that is
    about to be applied (the hook the reference calls at optimizer_mppi.py:192,195-197)."""

    num_states = NUM_STATES
    num_control_inputs = NUM_CONTROLS

    def __init__(self, weights: GRUWeights, dtype=torch.float32):
        self.w = weights
        self.dtype = dtype
        self.t = {k: torch.from_numpy(np.ascontiguousarray(getattr(weights, k))).to(dtype)
                  for k in ("Wi1", "Wh1", "bi1", "bh1", "Wi2", "Wh2", "bi2", "bh2", "W3", "b3")}
        self.hid = weights.hidden
        self.reset_state()

    def reset_state(self):
        self.h1 = torch.zeros(1, self.hid, dtype=self.dtype)
        self.h2 = torch.zeros(1, self.hid, dtype=self.dtype)

    def _cell(self, x, h, Wi, Wh, bi, bh):
        H = self.hid
        gi = x @ Wi + bi
        gh = h @ Wh + bh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        return (1.0 - z) * n + z * h

    def _net(self, s, Q, h1, h2):
        t = self.t
        x = torch.cat([Q, s[:, 1:]], dim=1)  # [N,6]
        h1 = self._cell(x, h1, t["Wi1"], t["Wh1"], t["bi1"], t["bh1"])
        h2 = self._cell(h1, h2, t["Wi2"], t["Wh2"], t["bi2"], t["bh2"])
        y = h2 @ t["W3"] + t["b3"]  # [N,5]
        angle = torch.atan2(y[:, 2], y[:, 1])
        return torch.cat([angle[:, None], y], dim=1), h1, h2

    def predict_core(self, s: torch.Tensor, Q: torch.Tensor) -> torch.Tensor:
        n = s.shape[0]
        h1, h2 = self.h1.repeat(n, 1), self.h2.repeat(n, 1)
        out = [s]
        for t in range(Q.shape[1]):
            s, h1, h2 = self._net(s, Q[:, t, :], h1, h2)
            out.append(s)
        return torch.stack(out, dim=1)

    def update(self, s=None, Q0=None):
        """s [N,6] (rows identical), Q0 [N,1,1] (rows identical): one step from the saved hidden state, saved back."""
        s = torch.as_tensor(s).to(self.dtype).reshape(-1, NUM_STATES)[:1]
        q = torch.as_tensor(Q0).to(self.dtype).reshape(-1, 1)[:1]
        _, self.h1, self.h2 = self._net(s, q, self.h1, self.h2)


# ----------------------------------------------------------------------------------------------
# Second environment: a Dubins car (SURVEY 8f.3 "a second environment's ODE + cost to prove the functor registry"; the reference's
# contract is any Control_Toolkit_ASF.Cost_Functions.<env>.<name>, Cost_Functions/cost_function_wrapper.py:59-66, and any
# num_states / num_control_inputs, Optimizers/optimizer_mppi.py:173-175).  3 states [x, y, yaw], 2 controls [throttle, steer] in
# [-1, 1]^2 -- so it also exercises num_control_inputs > 1.  Like the CartPole spec this arithmetic is the build's own pinned
# specification (no environment code lives in /root/reference): parity unpinned upstream, pinned between oracle and CUDA.
# ----------------------------------------------------------------------------------------------
DUBINS_NUM_STATES = 3
DUBINS_NUM_CONTROLS = 2


@dataclass
class DubinsParams:
    v_max: float = 1.0
    omega_max: float = 2.0
    dt: float = 0.02

    def f32(self) -> dict:
        return dict(v_max=_f32(self.v_max), omega_max=_f32(self.omega_max), h=_f32(self.dt))


def dubins_step(s: torch.Tensor, Q: torch.Tensor, c: dict) -> torch.Tensor:
    """s [N,3] = [x, y, yaw], Q [N,2] = [throttle, steer]; explicit Euler with the OLD yaw, then the yaw is wrapped to (-pi, pi]."""
    x, y, yaw = s.unbind(dim=1)
    v = c["v_max"] * Q[:, 0]
    w = c["omega_max"] * Q[:, 1]
    x = x + (v * torch.cos(yaw)) * c["h"]
    y = y + (v * torch.sin(yaw)) * c["h"]
    yaw = yaw + w * c["h"]
    yaw = torch.atan2(torch.sin(yaw), torch.cos(yaw))
    return torch.stack([x, y, yaw], dim=1)


class DubinsPredictor:
    num_states = DUBINS_NUM_STATES
    num_control_inputs = DUBINS_NUM_CONTROLS

    def __init__(self, params: DubinsParams | None = None):
        self.params = params or DubinsParams()
        self.c = self.params.f32()

    def predict_core(self, s: torch.Tensor, Q: torch.Tensor) -> torch.Tensor:
        out = [s]
        for t in range(Q.shape[1]):
            s = dubins_step(s, Q[:, t, :], self.c)
            out.append(s)
        return torch.stack(out, dim=1)


@dataclass
class DubinsCost:
    """``default`` cost of the Dubins car: squared distance to the target, a hard penalty inside a circular obstacle (an indicator,
    like the CartPole track barrier), control effort and control change rate; terminal cost = weighted squared distance."""
    name: str = "default"
    dd_weight: float = 10.0
    obstacle_weight: float = 1.0e4
    cc_weight: float = 1.0
    ccrc_weight: float = 1.0
    R: float = 0.1
    terminal_weight: float = 100.0
    MAX_COST: float = 0.0
    target_x: float = 1.0
    target_y: float = 0.5
    obstacle_x: float = 0.5
    obstacle_y: float = 0.2
    obstacle_r: float = 0.15

    def f32(self) -> dict:
        return dict(dd_weight=_f32(self.dd_weight), obstacle_weight=_f32(self.obstacle_weight), cc_weight=_f32(self.cc_weight),
                    ccrc_weight=_f32(self.ccrc_weight), R=_f32(self.R), terminal_weight=_f32(self.terminal_weight),
                    MAX_COST=_f32(self.MAX_COST), target_x=_f32(self.target_x), target_y=_f32(self.target_y),
                    obstacle_x=_f32(self.obstacle_x), obstacle_y=_f32(self.obstacle_y), obstacle_r2=_f32(self.obstacle_r * self.obstacle_r))

    def stage_cost(self, states: torch.Tensor, inputs: torch.Tensor, previous_input) -> torch.Tensor:
        """states [N,H,3], inputs [N,H,2], previous_input scalar / [2] -> [N,H]."""
        c = self.f32()
        previous_input = torch.as_tensor(previous_input).to(states.dtype)
        dx = states[:, :, 0] - c["target_x"]
        dy = states[:, :, 1] - c["target_y"]
        dd = c["dd_weight"] * (dx * dx + dy * dy)
        ox = states[:, :, 0] - c["obstacle_x"]
        oy = states[:, :, 1] - c["obstacle_y"]
        obs = c["obstacle_weight"] * ((ox * ox + oy * oy) < c["obstacle_r2"]).to(states.dtype)
        cc = c["cc_weight"] * _CC_cost(inputs, c)
        ccrc = c["ccrc_weight"] * _control_change_rate_cost(inputs, previous_input)
        return dd + obs + cc + ccrc

    def terminal_cost(self, terminal_states: torch.Tensor) -> torch.Tensor:
        c = self.f32()
        dx = terminal_states[:, 0] - c["target_x"]
        dy = terminal_states[:, 1] - c["target_y"]
        return c["terminal_weight"] * (dx * dx + dy * dy)


def dubins_synthetic_states(n: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.2, 0.2, n)
    y = rng.uniform(-0.2, 0.2, n)
    yaw = rng.uniform(-math.pi, math.pi, n)
    return np.stack([x, y, yaw], 1).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# Cost functions
# ----------------------------------------------------------------------------------------------
def _distance_difference_cost(position, c):
    return ((position - c["target_position"]) / c["two_thl"]) ** 2 + (
        (torch.abs(position) > c["thl_095"]).to(position.dtype)
        * 1.0e9
        * ((torch.abs(position) - c["thl_095"]) / c["thl_005"]) ** 2
    )


def _E_pot_cost(angle, c):
    return c["target_equilibrium"] * 0.25 * (1.0 - torch.cos(angle)) ** 2


def _CC_cost(u, c):
    return c["R"] * torch.sum(u ** 2, 2)


def _control_change_rate_cost(u, u_prev):
    u_prev_vec = torch.cat([torch.ones((u.shape[0], 1, u.shape[2]), dtype=u.dtype) * u_prev, u[:, :-1, :]], 1)
    return torch.sum((u - u_prev_vec) ** 2, 2)


def stage_cost(states: torch.Tensor, inputs: torch.Tensor, previous_input, cp: CostParams) -> torch.Tensor:
    """``_get_stage_cost``: states [N,H,6], inputs [N,H,1], previous_input scalar/0-d -> [N,H]."""
    c = cp.f32()
    previous_input = torch.as_tensor(previous_input).to(states.dtype)
    dd = c["dd_weight"] * _distance_difference_cost(states[:, :, POSITION], c)
    ep = c["ep_weight"] * _E_pot_cost(states[:, :, ANGLE], c)
    cc = c["cc_weight"] * _CC_cost(inputs, c)
    ccrc = c["ccrc_weight"] * _control_change_rate_cost(inputs, previous_input)
    if cp.name == "default":
        return dd + ep + cc + ccrc
    elif cp.name == "quadratic_boundary_grad":
        ekp = c["ekp_weight"] * states[:, :, ANGLED] ** 2
        border = (torch.abs(states[:, :, POSITION]) > c["thl_09"]).to(states.dtype) * 1.0e7
        return dd + ep + ekp + cc + ccrc + border
    raise ValueError(f"unknown cost function {cp.name}")


def terminal_cost(terminal_states: torch.Tensor, cp: CostParams) -> torch.Tensor:
    c = cp.f32()
    return 10000.0 * (
        (torch.abs(terminal_states[:, ANGLE]) > 0.2)
        | (torch.abs(terminal_states[:, POSITION] - c["target_position"]) > c["thl_01"])
    ).to(terminal_states.dtype)


def trajectory_cost(state_horizon: torch.Tensor, inputs: torch.Tensor, previous_input, cp: CostParams) -> torch.Tensor:
    """Restates reference ``Cost_Functions/__init__.py:74-93``:
    mean over H+1 of [stage costs (s_0..s_{H-1} with u_0..u_{H-1}) - MAX_COST, terminal cost(s_H)]."""
    if hasattr(cp, "stage_cost"):  # an environment that brings its own cost class (DubinsCost)
        sc = cp.stage_cost(state_horizon[:, :-1, :], inputs, previous_input) - cp.f32()["MAX_COST"]
        tc = cp.terminal_cost(state_horizon[:, -1, :]).reshape(-1, 1)
        return torch.mean(torch.cat([sc, tc], 1), 1)
    sc = stage_cost(state_horizon[:, :-1, :], inputs, previous_input, cp) - cp.f32()["MAX_COST"]
    tc = terminal_cost(state_horizon[:, -1, :], cp).reshape(-1, 1)
    return torch.mean(torch.cat([sc, tc], 1), 1)


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synthetic_states(n: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    angle = rng.uniform(-math.pi, math.pi, n)
    angleD = rng.uniform(-5.0, 5.0, n)
    position = rng.uniform(-0.15, 0.15, n)
    positionD = rng.uniform(-0.5, 0.5, n)
    s = np.stack([angle, angleD, np.cos(angle), np.sin(angle), position, positionD], 1)
    return s.astype(np.float32)
