import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from gpu_helpers import make_controller
from helpers import load_golden
from oracle import spec
name = sys.argv[1]
z, meta = load_golden(name)
ctrl = make_controller(meta, rng=None, logging=False)
s = spec.synthetic_states(8, seed=7)
for i in range(6):
    ctrl.step(s[i])
