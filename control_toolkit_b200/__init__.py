"""control_toolkit_b200 -- B200 (sm_100a) backend for the batched MPC rollout hot path of SensorsINI/Control_Toolkit.

Drop-in optimizer plugins (``optimizer_mppi``, ``optimizer_cem_tf``, ``optimizer_rpgd``) behind the reference's
``template_optimizer`` / ``controller_mpc`` API; the arithmetic runs in hand-written CUDA kernels reached through the
C ABI of ``libctk_b200.so`` (include/ctk_b200.h).  No TensorFlow, no Triton, no CPU fallback.
"""
from ._lib import BackendUnavailable, LIB_PATH  # noqa: F401
from .specs import GRUSpec, MLPSpec, register_gru, register_mlp  # noqa: F401
from .wrappers import CostFunctionWrapper, PredictorWrapper, VariableParameters  # noqa: F401

__version__ = "0.1.0"


def import_optimizer_by_name(optimizer_name: str):
    """Name -> class, with the reference's naming rule (reference others/globals_and_utils.py:103-133:
    '-' <-> '_', file ``optimizer_<name>.py`` holding class ``optimizer_<name>``)."""
    from importlib import import_module
    name = optimizer_name.replace("-", "_")
    full = name if name.startswith("optimizer") else "optimizer_" + name
    try:
        mod = import_module(f"{__name__}.Optimizers.{full}")
    except ModuleNotFoundError as e:
        raise ValueError(f"Optimizer {full} not found.") from e
    return getattr(mod, full)
