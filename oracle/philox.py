"""oracle.philox -- TEST INFRASTRUCTURE ONLY (never imported by control_toolkit_b200).

CPU restatement (numpy, integer arithmetic) of the generator behind the reference's ``self.rng``:
``tf.random.Generator.from_seed`` / ``create_rng`` (reference others/globals_and_utils.py:95-97) is Philox4x32-10 of
J. K. Salmon, M. A. Moraes, R. O. Dror, D. E. Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11.  The algorithm is a
third-party one (TensorFlow's / Random123's), absent from /root/reference, so it is restated here from the publication and PINNED
to the known-answer vectors Random123 1.09 ships for philox4x32-10 (``examples/kat_vectors``; tests/test_oracle_golden.py
``test_philox4x32_10_known_answers``).  TF's mapping of the words to a tensor's elements is internal to TF and not reproducible
offline (parity is defined under injected noise); what this module pins is the device generator K0
(control_toolkit_b200/csrc/ctk_device.cuh ``philox4x32_10`` / ``noise4``): its rounds, key schedule, counter layout
(draw block, global rollout id, tick, stream) and the word -> uniform / normal mapping -- bit-exact for the integer and uniform work
(tests/test_gpu_philox.py).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)   # round multipliers
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)   # Weyl key increments (golden ratio, sqrt(3) - 1)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr [..., 4] uint32, key [..., 2] uint32 (broadcast against ctr) -> [..., 4] uint32.  Ten rounds of
    (c0, c1, c2, c3) <- (hi(M1 c2) ^ c1 ^ k0, lo(M1 c2), hi(M0 c0) ^ c3 ^ k1, lo(M0 c0)), key bumped by (W0, W1) between rounds."""
    c = np.array(np.broadcast_to(np.asarray(ctr, np.uint32), np.broadcast_shapes(np.shape(ctr), np.shape(key)[:-1] + (4,))), np.uint32)
    k = np.array(np.broadcast_to(np.asarray(key, np.uint32), c.shape[:-1] + (2,)), np.uint32)
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = k0 + W0
            k1 = k1 + W1
    return np.stack([c0, c1, c2, c3], axis=-1)


def draw_words(seed: int, stream: int, tick: int, rows, per_rollout: int):
    """The 32-bit words behind draws [0, per_rollout) of the global rollout ids ``rows`` of noise block (stream, tick) of a handle
    seeded with ``seed``: counter = (draw // 4, rollout id, tick, stream), key = (seed low word, seed high word); draw i is word
    i % 4 of its counter's output (ctk_device.cuh noise4 / noise1).  -> uint32 [len(rows), per_rollout]."""
    rows = np.asarray(rows, np.uint32)
    nblk = (per_rollout + 3) // 4
    ctr = np.zeros((rows.size, nblk, 4), np.uint32)
    ctr[..., 0] = np.arange(nblk, dtype=np.uint32)[None, :]
    ctr[..., 1] = rows[:, None]
    ctr[..., 2] = np.uint32(tick & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32(stream)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32)
    w = philox4x32_10(ctr, key)
    return w.reshape(rows.size, nblk * 4)[:, :per_rollout]


def uniform24(words):
    """word -> U[0, 1) fp32 with 24 bits: (w >> 8) * 2^-24 (exact in fp32; ctk_device.cuh noise4, uniform branch)."""
    return ((np.asarray(words, np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def box_muller(words):
    """uint32 [..., 4k] -> float64 standard normals by the device's mapping (ctk_device.cuh noise4): per counter output
    (x, y, z, w): u1 = 2 - float(1.mantissa23(x >> 9)) in (0, 1], angle a2 = (float(1.mantissa23(y >> 9)) - 1.5) * 2 pi in [-pi, pi),
    out = (r1 cos a2, r1 sin a2, r3 cos a4, r3 sin a4) with r = sqrt(-2 ln u).  Evaluated in float64 on the fp32 uniforms: the device
    uses the MUFU approximations (lg2 / sqrt / sin / cos .approx), so device draws agree to approximation error, not bit for bit."""
    w = np.asarray(words, np.uint32)
    assert w.shape[-1] % 4 == 0
    m = ((w >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32)   # [1, 2)
    q = m.reshape(w.shape[:-1] + (w.shape[-1] // 4, 4))
    two_pi32 = np.float32(6.283185307)
    u1 = (np.float32(2.0) - q[..., 0]).astype(np.float64)
    u3 = (np.float32(2.0) - q[..., 2]).astype(np.float64)
    a2 = ((q[..., 1] - np.float32(1.5)) * two_pi32).astype(np.float64)   # fp32 product, as on the device
    a4 = ((q[..., 3] - np.float32(1.5)) * two_pi32).astype(np.float64)
    r1, r3 = np.sqrt(-2.0 * np.log(u1)), np.sqrt(-2.0 * np.log(u3))
    out = np.stack([r1 * np.cos(a2), r1 * np.sin(a2), r3 * np.cos(a4), r3 * np.sin(a4)], axis=-1)
    return out.reshape(w.shape), np.stack([u1, u1, u3, u3], axis=-1).reshape(w.shape)
