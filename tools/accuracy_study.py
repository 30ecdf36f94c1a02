"""GPU accuracy study (not a test): distribution of |cuda - exact| vs |reference_fp32 - exact| for one MPPI tick over
many random states, where exact = the reference algorithm evaluated in float64 on the same injected noise.
Usage (on the GPU box):  python tools/accuracy_study.py [n_states]"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from gpu_helpers import make_controller  # noqa: E402
from helpers import load_golden, make_oracle  # noqa: E402
from oracle import spec  # noqa: E402
from oracle.replay_rng import ReplayRNG  # noqa: E402


def main():
    n_states = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rows = []
    for name in ("mppi_c1_n64", "mppi_c1_n2000", "mppi_h100_n256"):
        z, meta = load_golden(name)
        states = spec.synthetic_states(n_states, seed=123)
        e_c, e_r, e_cr = [], [], []
        ctrl = make_controller(meta, rng=None)
        for i, s in enumerate(states):
            o32, o64 = make_oracle(meta), make_oracle(meta, dtype=torch.float64)
            ctrl.optimizer.optimizer_reset()
            ctrl.optimizer.rng = ReplayRNG(1000 + i, as_torch=False)
            ctrl.step(s)
            o32.step(s, ReplayRNG(1000 + i))
            o64.step(s, ReplayRNG(1000 + i))
            truth = o64.u_nom.numpy()
            sc = max(np.abs(truth).max(), 1e-2)
            e_c.append(np.abs(ctrl.optimizer.u_nom - truth).max() / sc)
            e_r.append(np.abs(o32.u_nom.numpy() - truth).max() / sc)
            e_cr.append(np.abs(ctrl.optimizer.u_nom - o32.u_nom.numpy()).max() / sc)
        q = lambda a: "med %.1e p90 %.1e max %.1e" % (np.median(a), np.quantile(a, 0.9), np.max(a))
        print(f"{name}: cuda-vs-exact [{q(e_c)}]  ref32-vs-exact [{q(e_r)}]  cuda-vs-ref32 [{q(e_cr)}]  "
              f"frac(cuda-vs-ref32 < 1e-5) = {np.mean(np.array(e_cr) < 1e-5):.2f}")


if __name__ == "__main__":
    main()
