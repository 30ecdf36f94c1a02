"""Timeline of a back-to-back CHAIN of MPPI/ODE ticks (diagnostics, GPU box):  python tools/k1_chain_trace.py [N] [ticks]
Launches `ticks` ticks with ctk_step_device_n (programmatic dependent launch chain, no events in between) and prints, for the
last three launches, when the blocks start (the overlap with the previous tick's finish), when the prologue ends (= the previous
launch has completed), when the rollouts end, when block 0 has finished the tick -- and the tick period."""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    lib = L.load()
    ctrl, N, H = bench.build_controller("mppi_ode_1m", n_override=n)
    opt = ctrl.optimizer
    st = torch.cuda.Stream()
    L.check(lib.ctk_set_stream(opt._h, C.c_void_p(st.cuda_stream)))
    states = torch.from_numpy(bench.synthetic_states(ticks, 0)).cuda()
    u = torch.zeros(ticks, 2, device="cuda")
    grid = C.c_int()
    L.check(lib.ctk_debug_trace(opt._h, 1, None, 0, C.byref(grid)))
    per = 320 * 8
    buf = np.zeros(4 * per, np.uint64)
    with torch.cuda.stream(st):
        L.check(lib.ctk_step_device_n(opt._h, C.c_void_p(states.data_ptr()), 6, C.c_void_p(u.data_ptr()), 2, 10))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.ctk_step_device_n(opt._h, C.c_void_p(states.data_ptr()), 6, C.c_void_p(u.data_ptr()), 2, ticks))
        e1.record()
        torch.cuda.synchronize()
    L.check(lib.ctk_debug_trace(opt._h, 1, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size, C.byref(grid)))
    g = grid.value
    t = buf.reshape(4, 320, 8)[:, :g, :].astype(np.int64)
    t0 = t[1, :, 0].min()
    t = (t - t0) / 1e3
    print(f"N={n} grid={g} chain of {ticks} ticks: {e0.elapsed_time(e1) * 1e3 / ticks:.2f} us per tick (one event pair around the chain)")
    print("  launch | blocks start: first / median / last (block 0) | prologue done: median (block 0) | rollouts done: median / last (block 0) | "
          "records stored (last) | tick finished | period")
    for i in range(1, 4):
        a = t[i]
        prev_fin = t[i - 1, 0, 5]
        print(f"  {i}: start {a[:, 0].min():8.2f} {np.median(a[:, 0]):8.2f} {a[:, 0].max():8.2f} ({a[0, 0]:8.2f}) | prologue {np.median(a[:, 1]):8.2f} ({a[0, 1]:8.2f}) | "
              f"rollouts {np.median(a[:, 2]):8.2f} {a[:, 2].max():8.2f} ({a[0, 2]:8.2f}) | records {a[:, 4].max():8.2f} | finished {a[0, 5]:8.2f} | "
              f"period {a[0, 5] - prev_fin:6.2f}  (previous finish -> median prologue done {np.median(a[:, 1]) - prev_fin:5.2f}; "
              f"finisher: last record stamp -> polled {a[0, 6] - a[:, 4].max():5.2f} -> combined {a[0, 7] - a[0, 6]:5.2f} -> finished {a[0, 5] - a[0, 7]:5.2f})")


if __name__ == "__main__":
    main()
