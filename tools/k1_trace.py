"""Per-block phase timeline of the MPPI/ODE rollout kernel (diagnostics, GPU box):  python tools/k1_trace.py [N] [flush]
Prints, relative to the earliest block start of one launch: when blocks start, finish the prologue, finish their rollouts,
store their record, and when the last block finishes the tick; plus the CUDA-event duration of the same tick."""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    do_flush = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
    lib = L.load()
    ctrl, N, H = bench.build_controller("mppi_ode_1m", n_override=n)
    opt = ctrl.optimizer
    st = torch.cuda.Stream()
    L.check(lib.ctk_set_stream(opt._h, C.c_void_p(st.cuda_stream)))
    states = torch.from_numpy(bench.synthetic_states(16, 0)).cuda()
    u = torch.zeros(4, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    grid = C.c_int()
    L.check(lib.ctk_debug_trace(opt._h, 1, None, 0, C.byref(grid)))
    buf = np.zeros(4 * 320 * 8, np.uint64)  # [last four launches][CTK_MBOX_BLOCKS][8 stamps]
    rows = []
    with torch.cuda.stream(st):
        for i in range(12):
            if do_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.ctk_step_device(opt._h, C.c_void_p(states[i].data_ptr()), C.c_void_p(u.data_ptr())))
            e1.record()
            torch.cuda.synchronize()
            L.check(lib.ctk_debug_trace(opt._h, 1, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size, C.byref(grid)))
            allt = buf.reshape(4, 320, 8)[:, :grid.value, :6].astype(np.int64)
            t = allt[int(np.argmax(allt[:, 0, 0]))]  # the most recent launch
            t0 = t[:, 0].min()
            t = (t - t0) / 1e3  # us
            if i >= 4:
                rows.append([e0.elapsed_time(e1) * 1e3, t[:, 0].max(), np.median(t[:, 1]), t[:, 1].max(), np.median(t[:, 2]), t[:, 2].min(),
                             t[:, 2].max(), t[:, 4].max(), t[:, 5].max()])
    r = np.median(np.array(rows), 0)
    print(f"N={n} grid={grid.value} flush={do_flush}  (microseconds, medians over 8 ticks, relative to the first block's start)")
    for name, v in zip(["event pair duration", "last block START", "prologue done (median block)", "prologue done (last block)",
                        "rollouts done (median block)", "rollouts done (first block)", "rollouts done (last block)",
                        "record stored (last block)", "tick finished (last block's finish)"], r):
        print(f"  {name:38s} {v:8.2f}")


if __name__ == "__main__":
    main()
