"""TEST-ONLY shim of ``SI_Toolkit.Predictors.predictor_wrapper.PredictorWrapper``.

Interface as used by the reference: ``configure(batch_size, dt,
computation_library, variable_parameters, predictor_specification[, horizon])``
(reference controller_mpc.py:67-73, optimizer_mppi.py:133-137, optimizer_rpgd.py:265-269),
``predict_core(s[N,ns], Q[N,H,nu]) -> [N,H+1,ns]``, ``update(s=, Q0=)`` (no-op for ODE / MLP),
``copy()``, ``.num_states`` / ``.num_control_inputs``.

The arithmetic is the build's pinned spec (oracle/spec.py).  ``predictor_specification``:
``"ODE"`` -> CartPole Euler ODE;  anything starting with ``"Dense"``/``"MLP"`` -> the MLP registered in
``MLP_REGISTRY[predictor_specification]`` (an ``oracle.spec.MLPWeights``);  ``"GRU..."`` -> the stateful recurrent predictor
registered in ``GRU_REGISTRY`` (an ``oracle.spec.GRUWeights``), whose ``update`` advances the saved hidden state.
"""
from oracle import spec as _spec

MLP_REGISTRY = {}
GRU_REGISTRY = {}
ENVIRONMENT = {"name": "CartPole"}  # SI_Toolkit takes the environment from its own configuration; the harness sets it per case
ODE_PARAMS = {"intermediate_steps": 1}


class PredictorWrapper:
    def __init__(self):
        dubins = ENVIRONMENT["name"] == "DubinsCar"
        self.num_states = _spec.DUBINS_NUM_STATES if dubins else _spec.NUM_STATES
        self.num_control_inputs = _spec.DUBINS_NUM_CONTROLS if dubins else _spec.NUM_CONTROLS
        self.predictor = None
        self.predictor_specification = None

    def configure(self, batch_size=None, dt=None, computation_library=None, variable_parameters=None,
                  predictor_specification=None, horizon=None, **kwargs):
        self.batch_size = batch_size
        self.horizon = horizon
        self.dt = dt
        self.predictor_specification = predictor_specification
        name = str(predictor_specification)
        if name.startswith("ODE") and ENVIRONMENT["name"] == "DubinsCar":
            self.predictor = _spec.DubinsPredictor(_spec.DubinsParams(dt=dt))
        elif name.startswith("ODE"):
            self.predictor = _spec.ODEPredictor(_spec.CartPoleParams(dt=dt, intermediate_steps=ODE_PARAMS["intermediate_steps"]))
        elif name.startswith(("Dense", "MLP")):
            self.predictor = _spec.MLPPredictor(MLP_REGISTRY[name])
        elif name.startswith("GRU"):
            self.predictor = _spec.GRUPredictor(GRU_REGISTRY[name])
        else:
            raise ValueError(f"unknown predictor_specification {predictor_specification}")

    def predict_core(self, s, Q):
        return self.predictor.predict_core(s, Q)

    def predict(self, s, Q):
        return self.predict_core(s, Q)

    def update(self, s=None, Q0=None):
        # RNN-state hook (reference optimizer_mppi.py:195-197); stateless predictors ignore it
        if hasattr(self.predictor, "update"):
            self.predictor.update(s=s, Q0=Q0)

    def copy(self):
        return PredictorWrapper()
