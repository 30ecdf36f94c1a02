"""TEST-ONLY application-side cost plugin of the SECOND environment (Dubins car) for the golden-vector harness: loaded by name at
reference ``Cost_Functions/cost_function_wrapper.py:59-66`` as ``Control_Toolkit_ASF.Cost_Functions.DubinsCar.default``.  It derives
from the reference's own ``cost_function_base`` (``get_stage_cost`` / ``get_trajectory_cost`` are the reference's code,
``Cost_Functions/__init__.py:49-93``) and delegates the arithmetic to the build's pinned spec (oracle/spec.py DubinsCost)."""
from Control_Toolkit.Cost_Functions import cost_function_base
from oracle import spec as _spec

COST_PARAMS = _spec.DubinsCost()


class default(cost_function_base):
    MAX_COST = COST_PARAMS.MAX_COST

    def get_terminal_cost(self, terminal_states):
        return COST_PARAMS.terminal_cost(terminal_states)

    def _get_stage_cost(self, states, inputs, previous_input):
        return COST_PARAMS.stage_cost(states, inputs, previous_input)
