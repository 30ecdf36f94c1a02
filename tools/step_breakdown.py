"""Where the host side of one controller.step() goes (C1: MPPI N=2000 H=50; C5: N=1e6 H=100), logging off, in-kernel Philox.
GPU box:  python tools/step_breakdown.py [ticks]
Rows: device-resident tick (CUDA events) | raw ctypes ctk_step | ctk_step_state | optimizer.step | controller.step (p50, us)."""
import ctypes as C
import os
import statistics
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from gpu_helpers import make_controller  # noqa: E402
from helpers import load_golden  # noqa: E402
from oracle import spec  # noqa: E402  (synthetic states only)


def p50(f, n, warm=20):
    for i in range(warm):
        f(i)
    lat = []
    for i in range(n):
        t0 = time.perf_counter()
        f(warm + i)
        lat.append(time.perf_counter() - t0)
    return statistics.median(lat) * 1e6, np.quantile(lat, 0.9) * 1e6


def main():
    import torch
    from control_toolkit_b200 import _lib as L
    ticks = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    for label, fixture, over in (("C1 MPPI N=2000 H=50", "mppi_c1_n2000", {}), ("C5 MPPI N=1e6 H=100", "mppi_h100_n256", {"num_rollouts": 1_000_000}),
                                 ("C2 CEM N=4096 k=64", "cem_c2_n4096_k64", {}), ("C3 RPGD N=32", "rpgd_c3", {})):
        z, meta = load_golden(fixture)
        ctrl = make_controller(meta, rng=None, logging=False, **over)
        opt = ctrl.optimizer
        lib = L.load()
        states = spec.synthetic_states(ticks + 40, seed=7)
        H = opt.mpc_horizon
        sdev = torch.from_numpy(states).cuda()
        udev = torch.zeros(4, device="cuda")
        L.check(lib.ctk_set_stream(opt._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        for i in range(20):
            L.check(lib.ctk_step_device(opt._h, C.c_void_p(sdev[i].data_ptr()), C.c_void_p(udev.data_ptr())))
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ticks)]
        for i in range(ticks):
            evs[i][0].record()
            L.check(lib.ctk_step_device(opt._h, C.c_void_p(sdev[20 + i].data_ptr()), C.c_void_p(udev.data_ptr())))
            evs[i][1].record()
        torch.cuda.synchronize()
        dev_us = statistics.median(a.elapsed_time(b) for a, b in evs) * 1e3
        ubuf, sbuf = np.zeros(1, np.float32), np.zeros(H, np.float32)
        s32 = [np.ascontiguousarray(s, np.float32) for s in states]
        raw = p50(lambda i: lib.ctk_step(opt._h, L.fptr(s32[i]), L.fptr(ubuf)), ticks)
        row = [("device tick (events, no L2 flush)", dev_us, 0.0), ("ctypes ctk_step", *raw)]
        if meta["optimizer"] == "mppi":
            row.append(("ctypes ctk_step_state", *p50(lambda i: lib.ctk_step_state(opt._h, L.fptr(s32[i]), L.fptr(ubuf), L.STATE_U_NOM, L.fptr(sbuf), H), ticks)))
        row.append(("optimizer.step", *p50(lambda i: opt.step(states[i]), ticks)))
        row.append(("controller.step", *p50(lambda i: ctrl.step(states[i]), ticks)))
        print(label)
        for name, a, b in row:
            print(f"    {name:36s} p50 {a:8.1f} us   p90 {b:8.1f} us")
        opt.close()


if __name__ == "__main__":
    main()
