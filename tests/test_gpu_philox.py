"""-m gpu: the device generator K0 (ctk_device.cuh philox4x32_10 / noise4) against the CPU restatement oracle/philox.py, which is
pinned to Random123's known-answer vectors (tests/test_oracle_golden.py::test_philox4x32_10_known_answers).

Bar: the integer work and the word -> uniform mapping are BIT-EXACT (rounds, key schedule, counter layout (draw block, global rollout
id, tick, stream), 24-bit uniforms); the normals go through the MUFU approximations (lg2 / sqrt / sin / cos .approx), so they agree
with a float64 Box-Muller on the same words to approximation error: 2e-5 absolute where u1 <= 0.99, 2e-3 on the innermost 1 % of the
radius (r = sqrt(-2 ln u1) amplifies lg2's absolute error by 1 / (2 r) as u1 -> 1)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fill(seed, kind, n):
    from control_toolkit_b200 import _lib as L
    lib = L.load()
    a = np.empty(n, np.float32)
    L.check(lib.ctk_philox_fill(0, seed, kind, L.fptr(a), n))
    return a


@pytest.mark.parametrize("seed", [42, (0x9E3779B9 << 32) | 0x7F4A7C15])
def test_device_uniforms_are_philox4x32_10_bit_exact(seed):
    """ctk_philox_fill draws element i as draw i % 16 of rollout i // 16 of block (stream MPPI = 0, tick 1), key = (seed low, high)."""
    from oracle import philox as P
    n = 1 << 16
    u = _fill(seed, 1, n)
    ref = P.uniform24(P.draw_words(seed, stream=0, tick=1, rows=np.arange(n // 16), per_rollout=16)).ravel()
    np.testing.assert_array_equal(u, ref)


def test_device_normals_are_box_muller_of_the_same_words():
    from oracle import philox as P
    n = 1 << 18
    z = _fill(42, 0, n).astype(np.float64)
    ref, u1 = P.box_muller(P.draw_words(42, stream=0, tick=1, rows=np.arange(n // 16), per_rollout=16))
    ref, u1 = ref.ravel(), u1.ravel()
    err = np.abs(z - ref)
    outer = u1 <= 0.99
    assert np.isfinite(z).all()
    assert err[outer].max() < 2e-5, err[outer].max()
    assert err.max() < 2e-3, err.max()


def test_exported_block_of_a_handle_uses_global_rollout_ids():
    """ctk_philox_export (the hook the production-pinning tests replay through the oracle): rows [row0, row0 + rows) of block
    (stream, tick) == counter word 1 = row0 + local row -- the property that makes the sampled population independent of sharding."""
    from control_toolkit_b200 import _lib as L
    from gpu_helpers import make_controller
    from helpers import load_golden
    from oracle import philox as P
    _, meta = load_golden("cem_c2_n256_k16")
    ctrl = make_controller(meta, rng=None, logging=False)
    opt = ctrl.optimizer
    seed = int(opt.seed) & 0xFFFFFFFFFFFFFFFF
    lib = L.load()
    rows, per, row0, tick, stream = 40, 7, 123457, 9, 1 | (2 << 8)
    out = np.empty(rows * per, np.float32)
    L.check(lib.ctk_philox_export(opt._h, stream, tick, per, 1, row0, rows, L.fptr(out)))
    ref = P.uniform24(P.draw_words(seed, stream=stream, tick=tick, rows=row0 + np.arange(rows), per_rollout=per)).ravel()
    np.testing.assert_array_equal(out, ref)
