"""Phase timeline of block 0 of the persistent CEM tick kernel (diagnostics, GPU box):  python tools/cem_trace.py
Per outer iteration, microseconds since the iteration's start: distribution in shared memory, rollouts done, block sort done,
grid candidates gathered + merged, refit done."""
import ctypes as C
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from control_toolkit_b200 import _lib as L  # noqa: E402


def main():
    lib = L.load()
    ctrl, N, H = bench.build_controller("cem_ode_c2")
    opt = ctrl.optimizer
    states = bench.synthetic_states(40, 0)
    grid = C.c_int()
    L.check(lib.ctk_debug_trace(opt._h, 1, None, 0, C.byref(grid)))
    buf = np.zeros(148 * 8, np.uint64)
    rows = []
    for i in range(30):
        ctrl.step(states[i])
        L.check(lib.ctk_debug_trace(opt._h, 1, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size, C.byref(grid)))
        if i >= 10:
            rows.append(buf[:3 * 8].reshape(3, 8).astype(np.int64).copy())
    r = np.stack(rows)  # [ticks, it, slot]
    names = ["dist ready", "rollouts done", "block sort done", "candidates merged", "refit done"]
    for it in range(3):
        d = (r[:, it, 1:6] - r[:, it, 0:1]) / 1e3
        print(f"iteration {it}: " + "  ".join(f"{n} {np.median(d[:, j]):6.2f}" for j, n in enumerate(names)))
    print("tick (first stamp -> last refit):", np.median((r[:, 2, 5] - r[:, 0, 0]) / 1e3), "us")


if __name__ == "__main__":
    main()
