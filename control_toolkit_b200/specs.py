"""Registry of the environments / predictors / cost functions that have DEVICE functors.

The reference's contract is "any Python subclass of cost_function_base" loaded by name
(reference Cost_Functions/cost_function_wrapper.py:59-66) and "any predictor_specification" resolved by SI_Toolkit
(reference Controllers/controller_mpc.py:67-73).  A fused CUDA rollout needs device code, so this backend keeps a
registry ``(environment_name, name) -> device functor id + parameter block`` and fails loudly (ValueError) for
anything unregistered -- there is no Python/CPU fallback.

Constants are the build's pinned spec (DESIGN.md "Spec").  Compound constants are evaluated in
float64 and rounded once to fp32, exactly as oracle/spec.py does (the two are written independently and compared in
tests/test_host_logic.py).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field, replace

import numpy as np

from . import _lib as L


def _f32(x) -> float:
    return float(np.float32(x))


@dataclass
class CartPoleODE:
    """CartPole Euler ODE parameters [UPSTREAM-RECALL of CartPoleSimulation, pinned]."""
    k: float = 1.0 / 3.0
    M: float = 0.230
    m: float = 0.087
    L: float = 0.395 / 2.0
    g: float = 9.81
    J_fric: float = 2.5e-4
    M_fric: float = 4.77
    u_max: float = 2.62
    intermediate_steps: int = 1

    def to_c(self, dt: float) -> L.ctk_ode_params:
        h = dt / self.intermediate_steps
        p = L.ctk_ode_params()
        p.u_max = _f32(self.u_max)
        p.kp1_Mm = _f32((self.k + 1.0) * (self.M + self.m))
        p.m = _f32(self.m)
        p.neg_M_fric = _f32(-self.M_fric)
        p.neg_J_fric = _f32(-self.J_fric)
        p.mg = _f32(self.m * self.g)
        p.L = _f32(self.L)
        p.kp1 = _f32(self.k + 1.0)
        p.mL = _f32(self.m * self.L)
        p.g = _f32(self.g)
        p.kp1L = _f32((self.k + 1.0) * self.L)
        p.h = _f32(h)
        p.intermediate_steps = int(self.intermediate_steps)
        return p


@dataclass
class CartPoleCost:
    """Weights of the CartPole cost classes (quadratic_boundary_grad values: reference
    Control_Toolkit_ASF_Template/config_cost_function.yml:11-18; `default` reuses them)."""
    kind: int = L.COST_DEFAULT
    dd_weight: float = 600.0
    ep_weight: float = 20000.0
    ekp_weight: float = 80.0
    cc_weight: float = 1.0
    ccrc_weight: float = 1.0
    R: float = 1.0
    MAX_COST: float = 0.0
    TrackHalfLength: float = 0.198

    def to_c(self, target_position: float = 0.0, target_equilibrium: float = 1.0) -> L.ctk_cost_params:
        thl = self.TrackHalfLength
        p = L.ctk_cost_params()
        p.kind = int(self.kind)
        p.dd_weight, p.ep_weight, p.ekp_weight = _f32(self.dd_weight), _f32(self.ep_weight), _f32(self.ekp_weight)
        p.cc_weight, p.ccrc_weight, p.R, p.MAX_COST = _f32(self.cc_weight), _f32(self.ccrc_weight), _f32(self.R), _f32(self.MAX_COST)
        p.two_thl, p.thl_095, p.thl_005 = _f32(2.0 * thl), _f32(0.95 * thl), _f32(0.05 * thl)
        p.thl_09, p.thl_01 = _f32(0.9 * thl), _f32(0.1 * thl)
        p.target_position, p.target_equilibrium = _f32(target_position), _f32(target_equilibrium)
        return p


@dataclass
class DubinsCarODE:
    """Second registered environment: Dubins car, state [x, y, yaw], controls [throttle, steer] in [-1, 1]^2;
    x += h v cos(yaw), y += h v sin(yaw), yaw = wrap(yaw + h w) with v = v_max Q0, w = omega_max Q1.  Pinned spec: oracle/spec.py
    dubins_step (written independently, compared in tests/test_host_logic.py)."""
    v_max: float = 1.0
    omega_max: float = 2.0
    intermediate_steps: int = 1


@dataclass
class DubinsCarCost:
    """`default` cost of the Dubins car (oracle/spec.py DubinsCost): squared distance to the target, an indicator penalty inside a
    circular obstacle, control effort, control change rate; terminal = weighted squared distance."""
    kind: int = L.COST_DEFAULT
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


def dubins_env_params(ode: DubinsCarODE, cost: DubinsCarCost, dt: float) -> np.ndarray:
    """The flat parameter block of DubinsEnv (ctk_kernels_env.cuh): compound constants in float64, rounded once."""
    return np.array([dt, ode.v_max, ode.omega_max, cost.dd_weight, cost.obstacle_weight, _f32(cost.cc_weight) * _f32(cost.R),
                     cost.ccrc_weight, cost.terminal_weight, cost.target_x, cost.target_y, cost.obstacle_x, cost.obstacle_y,
                     cost.obstacle_r * cost.obstacle_r, cost.MAX_COST], np.float32)


@dataclass(frozen=True)
class EnvironmentInfo:
    env_id: int
    num_states: int
    num_control_inputs: int
    env_params: object = None  # (ode_spec, cost_spec, dt) -> flat fp32 parameter block, for environments on the general kernels


# environment_name -> device functors.  CartPole: the fused kernels; anything else: the general-environment kernels (MPPI, CEM).
ENVIRONMENTS = {
    "CartPole": EnvironmentInfo(L.ENV_CARTPOLE, 6, 1),
    "DubinsCar": EnvironmentInfo(L.ENV_DUBINS_CAR, 3, 2, dubins_env_params),
}

# (environment_name, cost_function_name) -> prototype
COST_REGISTRY = {
    ("CartPole", "default"): CartPoleCost(kind=L.COST_DEFAULT),
    ("CartPole", "quadratic_boundary_grad"): CartPoleCost(kind=L.COST_QUADRATIC_BOUNDARY_GRAD),
    ("DubinsCar", "default"): DubinsCarCost(),
}
ODE_REGISTRY = {"CartPole": CartPoleODE(), "DubinsCar": DubinsCarODE()}


def environment_info(environment_name: str) -> EnvironmentInfo:
    if environment_name not in ENVIRONMENTS:
        raise ValueError(f"environment {environment_name!r} has no registered CUDA functors (registered: {sorted(ENVIRONMENTS)}); "
                         "this backend has no Python/CPU fallback")
    return ENVIRONMENTS[environment_name]


def resolve_cost(environment_name: str, cost_function_name: str, overrides: dict | None = None):
    key = (str(environment_name), str(cost_function_name).replace("-", "_"))
    if key not in COST_REGISTRY:
        raise ValueError(f"cost function {key[1]!r} of environment {key[0]!r} has no registered CUDA functor "
                         f"(registered: {sorted(COST_REGISTRY)}); this backend has no Python/CPU fallback")
    proto = COST_REGISTRY[key]
    names = {f for f in proto.__dataclass_fields__ if f != "kind"}
    ov = {k: float(v) for k, v in (overrides or {}).items() if k in names and k not in ("MAX_COST", "TrackHalfLength")}
    return replace(proto, **ov)


def cost_overrides_from_yaml(environment_name: str, cost_function_name: str, path: str | None = None) -> dict:
    """Weights from ``Control_Toolkit_ASF/config_cost_function.yml`` (CWD-relative, like the reference loads it at
    Cost_Functions/cost_function_wrapper.py:14), if that file and section exist."""
    path = path or os.path.join("Control_Toolkit_ASF", "config_cost_function.yml")
    if not os.path.isfile(path):
        return {}
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    return dict((cfg.get(environment_name) or {}).get(cost_function_name) or {})


@dataclass
class MLPSpec:
    """Dense 6 -> hidden tanh -> hidden tanh -> 5 autoregressive predictor (config C4).  Row-major [in,out]."""
    W1: np.ndarray
    b1: np.ndarray
    W2: np.ndarray
    b2: np.ndarray
    W3: np.ndarray
    b3: np.ndarray

    def __post_init__(self):
        for k in ("W1", "b1", "W2", "b2", "W3", "b3"):
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=np.float32))
        hid = self.W1.shape[1]
        if self.W1.shape != (6, hid) or self.W2.shape != (hid, hid) or self.W3.shape != (hid, 5) \
                or self.b1.shape != (hid,) or self.b2.shape != (hid,) or self.b3.shape != (5,):
            raise ValueError("MLP predictor must be 6 -> hidden -> hidden -> 5")
        if hid % 16 or not 16 <= hid <= 128:
            raise ValueError("MLP hidden width must be a multiple of 16 in [16, 128]")

    @property
    def hidden(self) -> int:
        return int(self.W1.shape[1])

    @staticmethod
    def random_init(seed: int = 2, hidden: int = 128) -> "MLPSpec":
        """weights N(0, 1/fan_in), biases 0, numpy default_rng(seed)."""
        rng = np.random.default_rng(seed)

        def w(i, o):
            return (rng.standard_normal((i, o)) / math.sqrt(i)).astype(np.float32)

        return MLPSpec(w(6, hidden), np.zeros(hidden, np.float32), w(hidden, hidden), np.zeros(hidden, np.float32),
                       w(hidden, 5), np.zeros(5, np.float32))

    def to_c(self) -> L.ctk_mlp_weights:
        w = L.ctk_mlp_weights()
        w.hidden = self.hidden
        for k in ("W1", "b1", "W2", "b2", "W3", "b3"):
            setattr(w, k, L.fptr(getattr(self, k)))
        return w


MLP_REGISTRY: dict[str, MLPSpec] = {}


def register_mlp(predictor_specification: str, spec: MLPSpec) -> None:
    """Make a network predictor available under a predictor_specification name (e.g. 'Dense-6IN-128H1-128H2-5OUT-0')."""
    MLP_REGISTRY[str(predictor_specification)] = spec


@dataclass
class GRUSpec:
    """Recurrent autoregressive predictor 6 -> GRU(hidden) -> GRU(hidden) -> Dense 5 ('GRU-6IN-32H1-32H2-5OUT-0').  Row-major
    [in, 3*hidden] matrices, gate order [r, z, n] (include/ctk_b200.h ctk_gru_weights).  The hidden state is SAVED on the device:
    every rollout starts from it and each MPPI tick advances it with (measured state, applied control) -- the reference's
    predictor.update hook (reference Optimizers/optimizer_mppi.py:192,195-197) -- inside the tick's kernel sequence."""
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
    _KEYS = ("Wi1", "Wh1", "bi1", "bh1", "Wi2", "Wh2", "bi2", "bh2", "W3", "b3")

    def __post_init__(self):
        for k in self._KEYS:
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=np.float32))
        h = self.Wh1.shape[0]
        shapes = dict(Wi1=(6, 3 * h), Wh1=(h, 3 * h), bi1=(3 * h,), bh1=(3 * h,), Wi2=(h, 3 * h), Wh2=(h, 3 * h), bi2=(3 * h,),
                      bh2=(3 * h,), W3=(h, 5), b3=(5,))
        for k, shp in shapes.items():
            if getattr(self, k).shape != shp:
                raise ValueError(f"GRU predictor: {k} must have shape {shp}, got {getattr(self, k).shape}")
        if h % 8 or not 8 <= h <= 32:
            raise ValueError("GRU hidden width must be a multiple of 8 in [8, 32]")

    @property
    def hidden(self) -> int:
        return int(self.Wh1.shape[0])

    @staticmethod
    def random_init(seed: int = 3, hidden: int = 32) -> "GRUSpec":
        """weights N(0, 1/fan_in), biases N(0, 0.1^2), numpy default_rng(seed); the position read-out is scaled by 0.08 (the same
        draws, in the same order, as oracle/spec.py GRUWeights.random_init -- written independently, compared in tests)."""
        rng = np.random.default_rng(seed)

        def w(i, o):
            return (rng.standard_normal((i, o)) / math.sqrt(i)).astype(np.float32)

        def b(o):
            return (0.1 * rng.standard_normal(o)).astype(np.float32)

        h = hidden
        Wi1, Wh1, bi1, bh1 = w(6, 3 * h), w(h, 3 * h), b(3 * h), b(3 * h)
        Wi2, Wh2, bi2, bh2 = w(h, 3 * h), w(h, 3 * h), b(3 * h), b(3 * h)
        W3, b3 = w(h, 5), b(5)
        W3[:, 3] *= np.float32(0.08)
        b3[3] *= np.float32(0.08)
        return GRUSpec(Wi1, Wh1, bi1, bh1, Wi2, Wh2, bi2, bh2, W3, b3)

    def to_c(self) -> L.ctk_gru_weights:
        w = L.ctk_gru_weights()
        w.hidden = self.hidden
        for k in self._KEYS:
            setattr(w, k, L.fptr(getattr(self, k)))
        return w


GRU_REGISTRY: dict[str, GRUSpec] = {}


def register_gru(predictor_specification: str, spec: GRUSpec) -> None:
    """Make a recurrent predictor available under a predictor_specification name (e.g. 'GRU-6IN-32H1-32H2-5OUT-0')."""
    GRU_REGISTRY[str(predictor_specification)] = spec


def resolve_predictor(environment_name: str, predictor_specification: str):
    """-> (PRED_ODE, CartPoleODE), (PRED_MLP, MLPSpec) or (PRED_GRU, GRUSpec).  ValueError if nothing is registered."""
    name = str(predictor_specification)
    if name.startswith("ODE"):
        if environment_name not in ODE_REGISTRY:
            raise ValueError(f"environment {environment_name!r} has no registered CUDA ODE (registered: {sorted(ODE_REGISTRY)})")
        return L.PRED_ODE, ODE_REGISTRY[environment_name]
    if name in MLP_REGISTRY:
        return L.PRED_MLP, MLP_REGISTRY[name]
    if name in GRU_REGISTRY:
        return L.PRED_GRU, GRU_REGISTRY[name]
    raise ValueError(f"predictor_specification {name!r} is neither 'ODE' nor a registered network "
                     f"(register_mlp / register_gru); registered networks: {sorted(MLP_REGISTRY) + sorted(GRU_REGISTRY)}")
