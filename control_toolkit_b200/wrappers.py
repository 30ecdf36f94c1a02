"""Spec-holding stand-ins for the two wrapper objects the reference's controller hands to an optimizer:

* ``PredictorWrapper``     (SI_Toolkit.Predictors.predictor_wrapper; interface as used at reference
  Controllers/controller_mpc.py:43,67-73, Optimizers/optimizer_mppi.py:87,133-137)
* ``CostFunctionWrapper``  (reference Cost_Functions/cost_function_wrapper.py:16-115)

In this backend the arithmetic runs inside the fused CUDA kernels, so these objects only carry the *specification*
(names, dt, live environment attributes).  The optimizers accept either these or the reference's own wrapper objects
(duck-typed on ``cost_function_name`` / ``environment_name`` / ``variable_parameters``).
"""
from __future__ import annotations

from types import SimpleNamespace


class VariableParameters(SimpleNamespace):
    """Live environment attributes (target_position, target_equilibrium, ...): reference Controllers/__init__.py:85-86."""

    def set_attributes(self, attributes: dict, device=None):
        for k, v in attributes.items():
            setattr(self, k, v)

    def update_attributes(self, attributes: dict):
        for k, v in attributes.items():
            setattr(self, k, v)


class PredictorWrapper:
    num_states = 6
    num_control_inputs = 1

    def __init__(self):
        self.predictor_specification = None
        self.batch_size = None
        self.dt = None
        self.horizon = None
        self.variable_parameters = None

    def configure(self, batch_size=None, dt=None, computation_library=None, variable_parameters=None,
                  predictor_specification=None, horizon=None, **kwargs):
        self.batch_size, self.dt, self.horizon = batch_size, dt, horizon
        self.variable_parameters = variable_parameters
        self.predictor_specification = predictor_specification

    def predict_core(self, s, Q):
        raise RuntimeError("control_toolkit_b200 predictors run inside the fused CUDA rollout kernels; "
                           "there is no host predict_core (and no CPU fallback)")

    def update(self, s=None, Q0=None):
        # RNN-state hook of the reference (optimizer_mppi.py:195-197).  Stateless predictors ignore it; for a recurrent predictor
        # (GRUSpec) the saved hidden state lives in the C handle and the hook runs on the device at the end of every MPPI tick
        # (gru_update_kernel), so there is nothing to do on the host either.
        pass

    def copy(self):
        return PredictorWrapper()


class CostFunctionWrapper:
    def __init__(self, cost_function_name_default: str = "default"):
        self.cost_function = None
        self.cost_function_name_default = cost_function_name_default
        self.cost_function_name = None
        self.environment_name = None
        self.variable_parameters = None
        self.weights = {}

    def configure(self, batch_size=None, horizon=None, variable_parameters=None, environment_name=None,
                  computation_library=None, cost_function_specification=None, weights: dict | None = None):
        self.batch_size, self.horizon = batch_size, horizon
        self.variable_parameters = variable_parameters
        self.environment_name = environment_name
        # reference cost_function_wrapper.py:76-86
        if cost_function_specification is None:
            self.cost_function_name = self.cost_function_name_default.replace("-", "_")
        elif isinstance(cost_function_specification, str):
            self.cost_function_name = cost_function_specification.replace("-", "_")
        else:
            raise ValueError(f"Cannot interpret cost function specification {cost_function_specification}.")
        self.weights = dict(weights or {})
        self.cost_function = SimpleNamespace(logged_attributes={}, reload_cost_parameters_from_config_flag=False)

    def update_cost_parameters_from_config(self):
        pass

    def copy(self):
        c = CostFunctionWrapper(self.cost_function_name_default)
        c.cost_function_name = self.cost_function_name
        return c
