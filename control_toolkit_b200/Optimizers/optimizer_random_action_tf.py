"""optimizer_random_action_tf -- B200 backend behind the reference's random-shooting plugin
(reference Optimizers/optimizer_random_action_tf.py:12-87; the class keeps the reference's name so that the
``random-action-tf`` key of config_optimizers.yml resolves to it, although no TensorFlow is involved).

Every tick: Q ~ U[action_low, action_high) for the whole population, rollout + trajectory cost, u = first control of the
cheapest rollout (``tf.argsort`` semantics: ties to the lower index).  On the device this is the CEM machinery with one
outer iteration, one elite and a uniform sampling distribution (ctk_config.cem_uniform_actions): the same fused
sample -> rollout -> cost kernel, the same bitonic top-k, u read from the regenerated elite row.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .. import _lib as L
from . import template_optimizer


class optimizer_random_action_tf(template_optimizer):
    _OPT = L.OPT_CEM

    def __init__(
        self,
        predictor,
        cost_function,
        control_limits: "Tuple[np.ndarray, np.ndarray]",
        computation_library=None,
        seed: int = None,
        mpc_horizon: int = 40,
        num_rollouts: int = 200,
        optimizer_logging: bool = False,
        calculate_optimal_trajectory: bool = False,
        **kwargs,
    ):
        super().__init__(predictor=predictor, cost_function=cost_function, control_limits=control_limits,
                         optimizer_logging=optimizer_logging, seed=seed, num_rollouts=num_rollouts,
                         mpc_horizon=mpc_horizon, computation_library=computation_library, **kwargs)
        self.best_index = None

    def configure(self, num_states: int, num_control_inputs: int, default_configure: bool = True, **kwargs):
        # template_optimizer.configure (reference Optimizers/__init__.py:52-63); dt and predictor_specification arrive
        # through **kwargs (controller_mpc.py:84-89)
        self.num_states, self.num_control_inputs = num_states, num_control_inputs
        self._create_backend(kwargs.get("dt", None), kwargs.get("predictor_specification", None))
        if default_configure:
            self.optimizer_reset()

    def _fill_config(self, cfg: L.ctk_config) -> None:
        cfg.cem_outer_it = 1
        cfg.cem_best_k = 1
        cfg.cem_warmup = 0
        cfg.cem_warmup_iterations = 1
        cfg.cem_initial_action_stdev = 1.0
        cfg.cem_stdev_min = 0.0
        cfg.cem_uniform_actions = 1

    def _population_block(self):
        return ("uniform", (self.num_rollouts, self.mpc_horizon, self.num_control_inputs))

    def step(self, s: np.ndarray, time=None):
        lib = self._require_backend()
        if self.optimizer_logging:
            self.logging_values = {"s_logged": np.asarray(s).copy()}
        self._refresh_live_cost(lib)
        self._feed_noise(lib, [self._population_block()])  # :56-61
        u = self._tick(lib, s)
        self.u = np.squeeze(u)  # :68
        H, nu, N = self.mpc_horizon, self.num_control_inputs, self._n_local
        if self.optimizer_logging:
            self.logging_values["Q_logged"] = self._get_log(L.LOG_Q, (N, H, nu))
            self.logging_values["J_logged"] = self._get_log(L.LOG_J, (N,))
            self.logging_values["rollout_trajectories_logged"] = self._get_log(L.LOG_ROLLOUTS, (N, H + 1, 6))
            self.logging_values["u_logged"] = self.u
            self.best_index = int(self._get_log(L.LOG_ELITE_IDX, (1, 1), np.int32)[0, 0])
        return self.u

    def optimizer_reset(self):
        lib = self._require_backend()
        L.check(lib.ctk_reset(self._h))
        # self.u (the cost's previous_input) survives optimizer_reset() in the reference: only optimizer_cem_tf.py:117 resets it
        if self.rng is not None:  # the reference draws (and discards) one population here (:78-87): keep a replay rng in step
            self.rng.standard_draws(*self._population_block())

    def last_costs(self) -> np.ndarray:
        return self._get_log(L.LOG_J, (self._n_local,))
