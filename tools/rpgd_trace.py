"""Phase timeline of the one-launch RPGD tick (rpgd_grad_coef_kernel, C3), diagnostics, GPU box:  python tools/rpgd_trace.py
Microseconds between consecutive stamps of block 0: staging | per Adam iteration: forward, coefficients, reverse, update | write-back |
final rollout | select."""
import ctypes as C
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from control_toolkit_b200 import _lib as L  # noqa: E402


def main():
    lib = L.load()
    ctrl, N, H = bench.build_controller("rpgd_ode_c3")
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
            rows.append(buf[:16].astype(np.int64).copy())
    r = np.stack(rows)
    names = ["staging", "it0 forward", "it0 coefficients", "it0 reverse", "it0 update", "it1 forward", "it1 coefficients", "it1 reverse",
             "it1 update", "write-back", "final rollout", "select"]
    d = np.median(np.diff(r[:, :13], axis=1), axis=0) / 1e3
    for n, v in zip(names, d):
        print(f"  {n:18s} {v:6.2f} us")
    print(f"  total              {np.median(r[:, 12] - r[:, 0]) / 1e3:6.2f} us")


if __name__ == "__main__":
    main()
