"""TEST-ONLY application-side cost plugin for the golden-vector harness: the class the reference loads by
name at ``Cost_Functions/cost_function_wrapper.py:59-66``.  It derives from the reference's own
``cost_function_base`` (so ``get_stage_cost`` / ``get_trajectory_cost`` are the reference's code,
``Cost_Functions/__init__.py:49-93``) and delegates the CartPole arithmetic to the build's pinned spec."""
from Control_Toolkit.Cost_Functions import cost_function_base
from oracle import spec as _spec

COST_PARAMS = _spec.CostParams(name="default")


def _live(params, vp):
    p = _spec.CostParams(**{**params.__dict__})
    if hasattr(vp, "target_position"):
        p.target_position = float(vp.target_position)
    if hasattr(vp, "target_equilibrium"):
        p.target_equilibrium = float(vp.target_equilibrium)
    return p


class default(cost_function_base):
    MAX_COST = COST_PARAMS.MAX_COST

    def get_terminal_cost(self, terminal_states):
        return _spec.terminal_cost(terminal_states, _live(COST_PARAMS, self.variable_parameters))

    def _get_stage_cost(self, states, inputs, previous_input):
        return _spec.stage_cost(states, inputs, previous_input, _live(COST_PARAMS, self.variable_parameters))
