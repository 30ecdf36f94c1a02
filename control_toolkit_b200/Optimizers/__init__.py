"""template_optimizer for the B200 backend.

Mirrors the reference's plugin interface (reference Optimizers/__init__.py:10-79): same constructor contract, same
``configure(num_states, num_control_inputs, default_configure, **kwargs)``, abstract ``step`` / ``optimizer_reset``,
``optimizer_name`` property and ``logging_values`` dict.  What differs is underneath: instead of a computation library
evaluating ``predict_core`` + ``get_trajectory_cost`` op by op, a C handle (libctk_b200.so) owns the optimizer state on
the GPU and runs the whole tick as fused CUDA kernels.

Noise: ``self.rng`` is ``None`` by default -> counter-based Philox4x32-10 generated in-kernel from ``seed``.
Assigning an object with ``standard_draws(kind, shape) -> np.ndarray`` (e.g. oracle.replay_rng.ReplayRNG in the tests)
switches to injected-noise mode: the same standard draws the reference would consume through ``rng.normal`` /
``rng.uniform`` are uploaded and consumed by the kernels (SURVEY.md section 8c noise-injection contract).
"""
from __future__ import annotations

import ctypes as C
from datetime import datetime
from typing import Tuple

import numpy as np

from .. import _lib as L
from .. import specs


class template_optimizer:
    # kept for signature compatibility; this backend does not dispatch on a computation library
    supported_computation_libraries = (object,)
    _OPT = None  # L.OPT_*

    def __init__(
        self,
        predictor,
        cost_function,
        control_limits: "Tuple[np.ndarray, np.ndarray]",
        optimizer_logging: bool,
        seed: int,
        num_rollouts: int,
        mpc_horizon: int,
        computation_library=None,
        **kwargs,
    ) -> None:
        self.lib = computation_library
        self.num_rollouts = int(num_rollouts)
        self.mpc_horizon = int(mpc_horizon)
        self.cost_function = cost_function
        self.u = 0.0
        self.predictor = predictor
        self.num_states = None
        self.num_control_inputs = None
        self.action_low, self.action_high = (np.atleast_1d(np.asarray(a, dtype=np.float32)) for a in control_limits)
        if seed is None:  # reference others/globals_and_utils.py:87-91
            seed = int((datetime.now() - datetime(1970, 1, 1)).total_seconds() * 1000.0)
        self.seed = int(seed)
        self.rng = None  # None -> in-kernel Philox; object with standard_draws() -> injected noise
        self.logging_values = {}
        self.optimizer_logging = bool(optimizer_logging)
        # backend options (not in the reference; all optional)
        self.device = int(kwargs.pop("device_index", 0))
        self.freeze_previous_input = bool(kwargs.pop("freeze_previous_input", False))
        self.mlp_engine = str(kwargs.pop("mlp_engine", "simt"))
        self.shard = kwargs.pop("shard", None)  # control_toolkit_b200.distributed.ShardPlan or None
        self.environment_name = kwargs.pop("environment_name", None)
        # logs of at least this many bytes are handed out as read-only VIEWS of a pinned host buffer owned by the C handle (valid
        # until the next step(); controller_mpc.update_logs copies them like the reference, Controllers/__init__.py:159-178);
        # smaller logs are fresh copies.  None: always copy.
        self.log_view_min_bytes = kwargs.pop("log_view_min_bytes", 16 << 20)
        # Optional top-M-only logging: with optimizer_logging on and logging_top_m = M, Q_logged / J_logged /
        # rollout_trajectories_logged hold only the M lowest-cost rollouts of the tick, best first (ties to the lower index), plus
        # their global rollout ids in "top_m_indices_logged" -- selected and gathered on the device, so the host receives
        # M x ((H+1) ns + H nu + 2) floats instead of the whole population's logs.  None (default): the reference's full logs.
        m = kwargs.pop("logging_top_m", None)
        self.logging_top_m = None if m in (None, 0) else int(m)
        # > 1: the handle carries that many independent clients whose ticks run in ONE launch (optimizer_mppi.step_batch; the
        # serving edge of SURVEY 8f.4).  1: the reference's single controller.
        self.num_clients = int(kwargs.pop("num_clients", 1))
        self._h = None
        self._cost_spec = None
        self._cost_live = (None, None)
        self._ode_spec = None
        self._dt = None

    # -- reference API -----------------------------------------------------------------------------------------
    def configure(self, num_states: int, num_control_inputs: int, default_configure: bool = True, **kwargs) -> None:
        self.num_states = num_states
        self.num_control_inputs = num_control_inputs
        if default_configure:
            self.optimizer_reset()

    def step(self, s: np.ndarray, time=None):
        raise NotImplementedError("Implement this function in a subclass.")

    def optimizer_reset(self):
        raise NotImplementedError("Implement this function in a subclass.")

    @property
    def optimizer_name(self):
        name = self.__class__.__name__
        if name != "template_optimizer":
            return name.replace("optimizer_", "").replace("_", "-").lower()
        else:
            raise AttributeError()

    # -- backend plumbing --------------------------------------------------------------------------------------
    def _fill_config(self, cfg: L.ctk_config) -> None:
        """Subclasses add their optimizer-specific fields."""
        raise NotImplementedError

    def _environment(self) -> str:
        return (self.environment_name or getattr(self.cost_function, "environment_name", None) or "CartPole")

    def _live_targets(self):
        vp = getattr(self.cost_function, "variable_parameters", None)
        tp = float(getattr(vp, "target_position", 0.0)) if vp is not None else 0.0
        te = float(getattr(vp, "target_equilibrium", 1.0)) if vp is not None else 1.0
        return tp, te

    def _create_backend(self, dt: float, predictor_specification: str) -> None:
        if dt is None or predictor_specification is None:
            raise ValueError(f"{self.__class__.__name__} requires dt and predictor_specification to be passed.")
        lib = L.load()
        env = self._environment()
        cost_name = getattr(self.cost_function, "cost_function_name", None) or "default"
        overrides = dict(specs.cost_overrides_from_yaml(env, cost_name))
        overrides.update(getattr(self.cost_function, "weights", None) or {})
        self._cost_spec = specs.resolve_cost(env, cost_name, overrides)
        info = specs.environment_info(env)
        if (int(self.num_states), int(self.num_control_inputs)) != (info.num_states, info.num_control_inputs):
            raise ValueError(f"environment {env!r} has {info.num_states} states and {info.num_control_inputs} control inputs, the controller "
                             f"configured {self.num_states} / {self.num_control_inputs}")
        self._env_info = info
        pred_kind, pred_spec = specs.resolve_predictor(env, predictor_specification)
        if info.env_id != L.ENV_CARTPOLE and pred_kind != L.PRED_ODE:
            raise ValueError("network predictors are registered for the CartPole environment only")
        ode_spec = pred_spec if pred_kind == L.PRED_ODE else specs.ODE_REGISTRY.get(env, specs.CartPoleODE())
        self._dt = float(dt)
        self._ode_spec = ode_spec

        cfg = L.ctk_config()
        cfg.abi_version = L.CTK_ABI_VERSION
        cfg.optimizer = self._OPT
        cfg.predictor = pred_kind
        cfg.device = self.device
        n_glob = self.num_rollouts
        if self.shard is not None:
            cfg.num_rollouts, cfg.rollout_offset = self.shard.local_count(n_glob), self.shard.local_offset(n_glob)
        else:
            cfg.num_rollouts, cfg.rollout_offset = n_glob, 0
        cfg.num_rollouts_global = n_glob
        cfg.mpc_horizon = self.mpc_horizon
        cfg.num_states = int(self.num_states)
        cfg.num_control_inputs = int(self.num_control_inputs)
        nu = info.num_control_inputs
        if self.action_low.size not in (1, nu) or self.action_high.size not in (1, nu):
            raise ValueError(f"control_limits must have 1 or {nu} entries per side for environment {env!r}")
        low = np.ascontiguousarray(np.broadcast_to(self.action_low, (nu,)), np.float32)
        high = np.ascontiguousarray(np.broadcast_to(self.action_high, (nu,)), np.float32)
        cfg.action_low, cfg.action_high = float(low[0]), float(high[0])
        cfg.environment = info.env_id
        cfg.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        cfg.logging = int(self.optimizer_logging)
        cfg.freeze_previous_input = int(self.freeze_previous_input)
        engines = {"simt": L.MLP_SIMT, "tcgen05": L.MLP_TCGEN05, "tcgen05_bf16": L.MLP_TCGEN05_BF16, "tcgen05_fast": L.MLP_TCGEN05_FAST}
        if self.mlp_engine not in engines:
            raise ValueError(f"mlp_engine must be one of {sorted(engines)}, got {self.mlp_engine!r}")
        cfg.mlp_engine = engines[self.mlp_engine]
        cfg.num_clients = self.num_clients
        self._fill_config(cfg)

        tp, te = self._live_targets()
        self._cost_live = (tp, te)
        if info.env_id == L.ENV_CARTPOLE:
            ode_c, cost_c = ode_spec.to_c(self._dt), self._cost_spec.to_c(tp, te)
        else:  # the CartPole blocks of ctk_create are ignored for other environments (their constants go through ctk_set_env_params)
            ode_c, cost_c = specs.CartPoleODE().to_c(self._dt), specs.CartPoleCost().to_c()
        if self._h is not None:
            lib.ctk_destroy(self._h)
            self._h = None
        h = C.c_void_p()
        L.check(lib.ctk_create(C.byref(cfg), C.byref(ode_c), C.byref(cost_c), C.byref(h)))
        self._h = h
        self._cfg = cfg
        self._n_local = int(cfg.num_rollouts)
        if info.env_id != L.ENV_CARTPOLE:
            L.check(lib.ctk_set_control_limits(self._h, L.fptr(low), L.fptr(high), nu))
            ep = np.ascontiguousarray(info.env_params(ode_spec, self._cost_spec, self._dt), np.float32)
            L.check(lib.ctk_set_env_params(self._h, L.fptr(ep), ep.size))
        if pred_kind == L.PRED_MLP:
            self._mlp_keepalive = pred_spec
            w = pred_spec.to_c()
            L.check(lib.ctk_set_mlp_weights(self._h, C.byref(w)))
        elif pred_kind == L.PRED_GRU:
            self._mlp_keepalive = pred_spec
            w = pred_spec.to_c()
            L.check(lib.ctk_set_gru_weights(self._h, C.byref(w)))
        self._u_buf = np.zeros(nu, np.float32)
        self._state_buf = None

    def _require_backend(self):
        if self._h is None:
            raise RuntimeError(f"{self.__class__.__name__}.configure() has not been called")
        return L.load()

    def _refresh_live_cost(self, lib) -> None:
        """target_position / target_equilibrium -- and the pole length L of the ODE predictor (reference
        controller_server/controller_server.py:21-28 lists it among the live attributes) -- may change between ticks
        (controller update_attributes, reference Controllers/__init__.py:106-107)."""
        if self._env_info.env_id != L.ENV_CARTPOLE:
            return  # (target_position / target_equilibrium / L are CartPole attributes)
        live = self._live_targets()
        if live != self._cost_live:
            c = self._cost_spec.to_c(*live)
            L.check(lib.ctk_set_cost_params(self._h, C.byref(c)))
            self._cost_live = live
        vp = getattr(self.cost_function, "variable_parameters", None)
        L_live = getattr(vp, "L", None) if vp is not None else None
        # (the reference's server initialises L to 0.0 until a client sends the real value, controller_server.py:19-26: ignored)
        if L_live is not None and float(L_live) > 0.0 and self._ode_spec is not None and float(L_live) != float(self._ode_spec.L):
            from dataclasses import replace
            self._ode_spec = replace(self._ode_spec, L=float(L_live))
            o = self._ode_spec.to_c(self._dt)
            L.check(lib.ctk_set_ode_params(self._h, C.byref(o)))

    def _feed_noise(self, lib, blocks) -> None:
        """blocks: list of (kind, shape) standard-draw blocks this call will consume, in order."""
        if self.rng is None:
            return
        if not hasattr(self.rng, "standard_draws"):
            raise TypeError("optimizer.rng must be None (in-kernel Philox) or provide standard_draws(kind, shape)")
        for kind, shape in blocks:
            z = np.ascontiguousarray(self.rng.standard_draws(kind, shape), dtype=np.float32).ravel()
            L.check(lib.ctk_push_injected_noise(self._h, L.fptr(z), z.size))

    def _tick(self, lib, s: np.ndarray, state_which=None, state_n=0) -> np.ndarray:
        s32 = np.ascontiguousarray(np.asarray(s, dtype=np.float32).reshape(-1))
        if s32.size != int(self.num_states):
            raise ValueError(f"state must have {int(self.num_states)} entries, got {s32.size}")
        if self.shard is not None and self.shard.world_size > 1:
            return self.shard.run_tick(self, lib, s32, state_which, state_n)
        if state_which is not None:  # u and one [H] state array with a single device->host window and synchronisation
            if self._state_buf is None or self._state_buf.size != state_n:
                self._state_buf = np.empty(state_n, np.float32)
            L.check(lib.ctk_step_state(self._h, L.fptr(s32), L.fptr(self._u_buf), state_which, L.fptr(self._state_buf), state_n))
        else:
            L.check(lib.ctk_step(self._h, L.fptr(s32), L.fptr(self._u_buf)))
        return self._u_buf.copy()

    # -- state / logs ------------------------------------------------------------------------------------------
    def _get_state(self, which: int, shape) -> np.ndarray:
        lib = self._require_backend()
        out = np.empty(int(np.prod(shape)), np.float32)
        L.check(lib.ctk_get_state(self._h, which, L.fptr(out), out.size))
        return out.reshape(shape)

    def _set_state(self, which: int, value) -> None:
        lib = self._require_backend()
        a = np.ascontiguousarray(np.asarray(value, dtype=np.float32)).ravel()
        L.check(lib.ctk_set_state(self._h, which, L.fptr(a), a.size))

    def _get_counter(self, which: int) -> int:
        lib = self._require_backend()
        v = C.c_int64()
        L.check(lib.ctk_get_counter(self._h, which, C.byref(v)))
        return int(v.value)

    def _set_counter(self, which: int, value: int) -> None:
        lib = self._require_backend()
        L.check(lib.ctk_set_counter(self._h, which, int(value)))

    def _get_log(self, which: int, shape, dtype=np.float32) -> np.ndarray:
        lib = self._require_backend()
        count = int(np.prod(shape))
        if self.log_view_min_bytes is not None and count * np.dtype(dtype).itemsize >= self.log_view_min_bytes:
            ptr, n = C.c_void_p(), C.c_size_t()
            L.check(lib.ctk_get_log_view(self._h, which, C.byref(ptr), C.byref(n)))
            if n.value != count * np.dtype(dtype).itemsize:
                raise RuntimeError(f"log {which}: {n.value} bytes on the device, {count * np.dtype(dtype).itemsize} expected")
            view = np.frombuffer((C.c_char * n.value).from_address(ptr.value), dtype=dtype, count=count).reshape(shape)
            view.flags.writeable = False
            return view
        out = np.empty(count, dtype)
        L.check(lib.ctk_get_log(self._h, which, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out.reshape(shape)

    def _get_log_top(self, m: int, with_rows: bool = True) -> dict:
        """The logs of the m lowest-cost rollouts of the last tick (ctk_get_log_top): {"idx" [m] int32 global rollout ids, "J" [m],
        "Q" [m, H, nu], "traj" [m, H+1, ns]} -- Q / traj only when the handle logs (with_rows)."""
        lib = self._require_backend()
        H, nu, ns = self.mpc_horizon, int(self.num_control_inputs), int(self.num_states)
        idx, J = np.empty(m, np.int32), np.empty(m, np.float32)
        Q = np.empty((m, H, nu), np.float32) if with_rows else None
        traj = np.empty((m, H + 1, ns), np.float32) if with_rows else None
        L.check(lib.ctk_get_log_top(self._h, int(m), idx.ctypes.data_as(C.POINTER(C.c_int32)), L.fptr(J),
                                    L.fptr(Q) if with_rows else None, L.fptr(traj) if with_rows else None))
        return {"idx": idx, "J": J, "Q": Q, "traj": traj}

    def _collect_rollout_logs(self, N: int) -> np.ndarray:
        """Fills logging_values' Q_logged / J_logged / rollout_trajectories_logged (reference optimizer_mppi.py:214-218,
        optimizer_cem_tf.py:104-108) -- the whole population, or its logging_top_m best rollouts -- and returns the trajectories."""
        H, nu, ns = self.mpc_horizon, int(self.num_control_inputs), int(self.num_states)
        if self.logging_top_m is not None:
            top = self._get_log_top(min(self.logging_top_m, N))
            self.logging_values["Q_logged"] = top["Q"]
            self.logging_values["J_logged"] = top["J"]
            self.logging_values["rollout_trajectories_logged"] = top["traj"]
            self.logging_values["top_m_indices_logged"] = top["idx"]
            return top["traj"]
        traj = self._get_log(L.LOG_ROLLOUTS, (N, H + 1, ns))
        self.logging_values["Q_logged"] = self._get_log(L.LOG_Q, (N, H, nu))
        self.logging_values["J_logged"] = self._get_log(L.LOG_J, (N,))
        self.logging_values["rollout_trajectories_logged"] = traj
        return traj

    def export_philox(self, stream: int, tick: int, per_rollout: int, rows: int, uniform: bool = False, row0: int = 0,
                      sub: int = 0) -> np.ndarray:
        """The standard draws [rows, per_rollout] the kernels generate in-kernel for noise block (stream | sub << 8, tick) of this
        handle (verification of the production instantiations against the oracle: ctk_philox_export)."""
        lib = self._require_backend()
        out = np.empty((int(rows), int(per_rollout)), np.float32)
        L.check(lib.ctk_philox_export(self._h, int(stream) | (int(sub) << 8), int(tick), int(per_rollout), int(bool(uniform)),
                                      int(row0), int(rows), L.fptr(out)))
        return out

    @property
    def last_kernel(self) -> str:
        """Template instantiation of the last rollout-kernel launch (ctk_last_kernel)."""
        return (self._require_backend().ctk_last_kernel(self._h) or b"").decode()

    @property
    def tick_counter(self) -> int:
        return self._get_counter(L.COUNTER_TICK)

    @property
    def gpu_launches(self) -> int:
        lib = self._require_backend()
        v = C.c_int64()
        L.check(lib.ctk_get_launch_count(self._h, C.byref(v)))
        return int(v.value)

    def rollout_single(self, s: np.ndarray, Q: np.ndarray):
        """Nominal rollout of one control sequence -> (trajectory [1,H+1,ns], summed stage cost)."""
        lib = self._require_backend()
        s32 = np.ascontiguousarray(np.asarray(s, np.float32).reshape(-1))
        q32 = np.ascontiguousarray(np.asarray(Q, np.float32).reshape(-1))
        traj = np.empty((self.mpc_horizon + 1) * 6, np.float32)
        summed = np.zeros(1, np.float32)
        L.check(lib.ctk_rollout_single(self._h, L.fptr(s32), L.fptr(q32), L.fptr(traj), L.fptr(summed)))
        return traj.reshape(1, self.mpc_horizon + 1, 6), summed

    def close(self):
        if self._h is not None:
            try:
                L.load().ctk_destroy(self._h)
            finally:
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
