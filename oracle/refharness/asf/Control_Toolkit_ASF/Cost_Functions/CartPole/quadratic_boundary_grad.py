"""TEST-ONLY application-side cost plugin (see default.py); RPGD's cost class."""
from Control_Toolkit.Cost_Functions import cost_function_base
from oracle import spec as _spec
from .default import _live

COST_PARAMS = _spec.CostParams(name="quadratic_boundary_grad")


class quadratic_boundary_grad(cost_function_base):
    MAX_COST = COST_PARAMS.MAX_COST

    def get_terminal_cost(self, terminal_states):
        return _spec.terminal_cost(terminal_states, _live(COST_PARAMS, self.variable_parameters))

    def _get_stage_cost(self, states, inputs, previous_input):
        return _spec.stage_cost(states, inputs, previous_input, _live(COST_PARAMS, self.variable_parameters))
