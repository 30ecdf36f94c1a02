"""p50 controller.step() latency (host state in -> host control out) and kernel launches per tick for the BASELINE configs,
in-kernel Philox noise, logging off.  GPU box:  python tools/latency_table.py [ticks]"""
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

CASES = [
    ("C1 MPPI  N=2000  H=50  ODE", "mppi_c1_n2000", {}, {}),
    ("C2 CEM   N=4096 k=64 H=50 ODE (3 outer iterations)", "cem_c2_n4096_k64", {}, {}),
    ("C3 RPGD  N=32   H=50  ODE (2 Adam steps + selection)", "rpgd_c3", {}, {}),
    ("C4 MPPI  N=65536 H=100 MLP 128x128, tcgen05", "mppi_mlp_c4_n256", {"num_rollouts": 65536, "mlp_engine": "tcgen05"}, {}),
    ("C4 MPPI  N=65536 H=100 MLP 128x128, FP32 pipe", "mppi_mlp_c4_n256", {"num_rollouts": 65536, "mlp_engine": "simt"}, {"ticks": 12}),
    ("C5 MPPI  N=1e6  H=100 ODE (1 GPU)", "mppi_h100_n256", {"num_rollouts": 1_000_000}, {}),
]


def main():
    ticks = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    print(f"{'config':58s} {'p50 ms':>9s} {'p90 ms':>9s} {'launches/tick':>14s} {'rollout-steps/s (p50)':>22s}")
    for label, fixture, over, opts in CASES:
        z, meta = load_golden(fixture)
        ctrl = make_controller(meta, rng=None, logging=False, **over)
        opt = ctrl.optimizer
        n = opts.get("ticks", ticks)
        states = spec.synthetic_states(n + 20, seed=7)
        for i in range(min(20, max(3, n // 4))):
            ctrl.step(states[i])
        l0 = opt.gpu_launches
        lat = []
        for i in range(n):
            t0 = time.perf_counter()
            ctrl.step(states[20 + i])
            lat.append(time.perf_counter() - t0)
        launches = (opt.gpu_launches - l0) / n
        p50 = statistics.median(lat)
        N, H = opt.num_rollouts, opt.mpc_horizon
        passes = {"mppi": 1, "cem-tf": meta["cfg"].get("cem_outer_it", 1), "rpgd": 2 * meta["cfg"].get("outer_its", 1) + 1}[meta["optimizer"]]
        print(f"{label:58s} {p50 * 1e3:9.4f} {np.quantile(lat, 0.9) * 1e3:9.4f} {launches:14.1f} {N * H * passes / p50:22.3e}")
        opt.close()


if __name__ == "__main__":
    main()
