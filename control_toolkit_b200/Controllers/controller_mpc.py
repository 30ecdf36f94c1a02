"""controller_mpc -- the boundary caller of the hot path, mirroring reference Controllers/controller_mpc.py:21-109
and the parts of template_controller it relies on (reference Controllers/__init__.py:27-178): same
``configure(optimizer_name, predictor_specification)``, ``step(s, time, updated_attributes) -> u``,
``controller_reset()``, ``update_logs`` / ``get_outputs`` with the reference's ``save_vars``.

Configuration: the reference reads three YAML files relative to the CWD at import time; here the same files are read
lazily at construction (``Control_Toolkit_ASF/config_{controllers,optimizers,cost_function}.yml``) unless dicts are
passed in.  The controller is NOT re-implemented on the GPU; it just owns the optimizer plugin.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from .. import import_optimizer_by_name
from ..wrappers import CostFunctionWrapper, PredictorWrapper, VariableParameters


def _load_yaml(name: str) -> dict:
    import yaml
    path = os.path.join("Control_Toolkit_ASF", name)
    with open(path) as f:
        return yaml.safe_load(f) or {}


class controller_mpc:
    _has_optimizer = True

    def __init__(self, environment_name: str, control_limits, initial_environment_attributes: dict,
                 config_controller: Optional[dict] = None, config_optimizers: Optional[dict] = None,
                 config_cost_function: Optional[dict] = None):
        self.config_controller = dict(config_controller if config_controller is not None
                                      else _load_yaml("config_controllers.yml")["mpc"])
        self._config_optimizers = config_optimizers
        self._config_cost_function = config_cost_function
        self.environment_name = environment_name
        self.control_limits = control_limits
        self.action_low, self.action_high = control_limits
        self.variable_parameters = VariableParameters()
        self.variable_parameters.set_attributes(dict(initial_environment_attributes))
        self.u = 0.0
        self.controller_logging = bool(self.config_controller.get("controller_logging", False))
        self.save_vars = ["Q_logged", "J_logged", "s_logged", "u_logged", "realized_cost_logged",
                          "trajectory_ages_logged", "rollout_trajectories_logged"]  # reference Controllers/__init__.py:89-97
        self.logs = {s: [] for s in self.save_vars}
        self.controller_data_for_csv = {}
        self.optimizer = None

    @property
    def controller_name(self):
        return "mpc"

    @property
    def has_optimizer(self):
        return self._has_optimizer

    def configure(self, optimizer_name: Optional[str] = None, predictor_specification: Optional[str] = None):
        if optimizer_name in {None, ""}:
            optimizer_name = str(self.config_controller["optimizer"])
        if predictor_specification in {None, ""}:
            predictor_specification = self.config_controller.get("predictor_specification", None)
        cfgs = self._config_optimizers if self._config_optimizers is not None else _load_yaml("config_optimizers.yml")
        config_optimizer = dict(cfgs[optimizer_name])
        cost_cfg = self._config_cost_function
        if cost_cfg is None and os.path.isfile(os.path.join("Control_Toolkit_ASF", "config_cost_function.yml")):
            cost_cfg = _load_yaml("config_cost_function.yml")
        cost_cfg = cost_cfg or {}
        cost_function_specification = self.config_controller.get("cost_function_specification", None)
        self.cost_function = CostFunctionWrapper(cost_cfg.get("cost_function_name_default", "default"))
        self.predictor = PredictorWrapper()
        # the reference's PredictorWrapper takes the dimensions from SI_Toolkit's configuration; here: from the environment registry
        from ..specs import environment_info
        info = environment_info(self.environment_name)
        self.predictor.num_states, self.predictor.num_control_inputs = info.num_states, info.num_control_inputs

        Optimizer = import_optimizer_by_name(optimizer_name)  # reference :56
        self.optimizer = Optimizer(
            predictor=self.predictor,
            cost_function=self.cost_function,
            control_limits=self.control_limits,
            optimizer_logging=self.controller_logging,
            computation_library=None,
            calculate_optimal_trajectory=bool(self.config_controller.get("calculate_optimal_trajectory")),
            **config_optimizer,
        )
        self.predictor.configure(batch_size=self.optimizer.num_rollouts, dt=config_optimizer["mpc_timestep"],
                                 variable_parameters=self.variable_parameters,
                                 predictor_specification=predictor_specification)
        name = (cost_function_specification or self.cost_function.cost_function_name_default).replace("-", "_")
        weights = dict((cost_cfg.get(self.environment_name) or {}).get(name) or {})
        self.cost_function.configure(batch_size=self.optimizer.num_rollouts, horizon=self.optimizer.mpc_horizon,
                                     variable_parameters=self.variable_parameters,
                                     environment_name=self.environment_name,
                                     cost_function_specification=cost_function_specification, weights=weights)
        self.optimizer.configure(dt=config_optimizer["mpc_timestep"], predictor_specification=predictor_specification,
                                 num_states=self.predictor.num_states, num_control_inputs=self.predictor.num_control_inputs)
        self.controller_data_for_csv = self.cost_function.cost_function.logged_attributes

    def update_attributes(self, updated_attributes: dict):
        self.variable_parameters.update_attributes(updated_attributes)

    def step(self, s: np.ndarray, time=None, updated_attributes: dict = {}):
        self.cost_function.update_cost_parameters_from_config()
        self.update_attributes(updated_attributes)
        u = self.optimizer.step(s, time)
        self.update_logs(self.optimizer.logging_values)
        return u

    def controller_reset(self):
        self.optimizer.optimizer_reset()

    def update_logs(self, logging_values: dict) -> None:
        if self.controller_logging:
            for name in self.save_vars:
                var = logging_values.get(name, None)
                if var is not None:
                    self.logs[name].append(np.array(var, copy=True))

    def get_outputs(self) -> dict:
        return {name: np.stack(v, axis=0) if len(v) > 0 else None for name, v in self.logs.items()}
