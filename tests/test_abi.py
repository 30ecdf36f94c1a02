"""CPU tests of the drop-in boundary: the C-ABI shared library loads, exports every symbol include/ctk_b200.h declares,
the ctypes structs match the C structs byte for byte, and the product fails LOUDLY (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "ctk_b200.h")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from control_toolkit_b200 import _lib as L
    lib = L.load()  # raises if the .so is missing or a bound symbol is absent
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ctk_b200.h but not exported"
        assert name in L.SYMBOLS, f"{name} declared in include/ctk_b200.h but not bound in _lib.SYMBOLS"
    assert set(L.SYMBOLS) == set(declared)
    assert lib.ctk_abi_version() == L.CTK_ABI_VERSION


def test_struct_layout_matches_header(tmp_path):
    from control_toolkit_b200 import _lib as L
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ctk_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(ctk_config),sizeof(ctk_ode_params),sizeof(ctk_cost_params),sizeof(ctk_mlp_weights),'
                   'offsetof(ctk_config,seed),offsetof(ctk_config,rpgd_beta_1),offsetof(ctk_config,mlp_engine));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    exp = [C.sizeof(L.ctk_config), C.sizeof(L.ctk_ode_params), C.sizeof(L.ctk_cost_params), C.sizeof(L.ctk_mlp_weights),
           L.ctk_config.seed.offset, L.ctk_config.rpgd_beta_1.offset, L.ctk_config.mlp_engine.offset]
    assert got == exp


def test_error_conventions_without_compute():
    """Argument validation happens before any CUDA call: bad configs -> CTK_EINVAL -> ValueError (reference conventions:
    ValueError for unsupported configuration, Optimizers/__init__.py:27-28, optimizer_rpgd.py:271,291)."""
    from control_toolkit_b200 import _lib as L
    lib = L.load()
    cfg, ode, cost = L.ctk_config(), L.ctk_ode_params(), L.ctk_cost_params()
    h = C.c_void_p()
    cfg.abi_version = 999
    assert lib.ctk_create(C.byref(cfg), C.byref(ode), C.byref(cost), C.byref(h)) == L.CTK_EINVAL
    assert b"abi_version" in lib.ctk_last_error()
    cfg.abi_version = L.CTK_ABI_VERSION
    cfg.optimizer = 7
    with pytest.raises(ValueError):
        L.check(lib.ctk_create(C.byref(cfg), C.byref(ode), C.byref(cost), C.byref(h)))
    cfg.optimizer, cfg.num_states, cfg.num_control_inputs = L.OPT_MPPI, 4, 2
    with pytest.raises(ValueError, match="CartPole"):
        L.check(lib.ctk_create(C.byref(cfg), C.byref(ode), C.byref(cost), C.byref(h)))
    assert lib.ctk_destroy(None) == 0


@pytest.mark.skipif(_has_cuda(), reason="needs a machine WITHOUT a GPU")
def test_no_cpu_fallback():
    """Without a CUDA device the optimizer plugin must raise, never compute on the CPU."""
    import control_toolkit_b200 as ctk
    from control_toolkit_b200.Controllers.controller_mpc import controller_mpc
    ctrl = controller_mpc("CartPole", (np.array([-1.0], np.float32), np.array([1.0], np.float32)),
                          {"target_position": 0.0, "target_equilibrium": 1.0},
                          config_controller=dict(optimizer="mppi", predictor_specification="ODE", cost_function_specification="default",
                                                 controller_logging=False),
                          config_optimizers={"mppi": dict(seed=1, mpc_horizon=10, mpc_timestep=0.02, num_rollouts=32, cc_weight=1.0, R=1.0,
                                                          LBD=100.0, NU=1000.0, SQRTRHOINV=0.03, period_interpolation_inducing_points=5)},
                          config_cost_function={})
    with pytest.raises(RuntimeError):
        ctrl.configure(optimizer_name="mppi", predictor_specification="ODE")


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under control_toolkit_b200/ (or the C sources) may reference it."""
    bad = []
    for root, _, files in os.walk(os.path.join(REPO, "control_toolkit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "libctk_host_twin" in txt:
                    bad.append(os.path.join(root, f))
    assert not bad, bad
