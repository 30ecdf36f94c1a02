"""oracle.replay_rng -- TEST INFRASTRUCTURE ONLY.

Injectable-noise rng (SURVEY.md section 8c 'Noise-injection contract').  The reference always draws through
``self.rng`` (``optimizer_mppi.py:173``, ``optimizer_cem_tf.py:64``, ``optimizer_rpgd.py:277,284``); replacing
that attribute by a ``ReplayRNG`` makes the unmodified reference files and the CUDA kernels consume the SAME
standard draws ``z``:

* ``normal(shape, dtype, mean, stddev)``  -> ``z * stddev + mean``   (fp32, z ~ N(0,1), C-order over shape)
* ``uniform(shape, dtype, minval, maxval)`` -> ``z * (maxval - minval) + minval`` (fp32, z ~ U[0,1))
* ``standard_draws(kind, shape)`` -> the raw ``z`` block as numpy fp32 (the product's injected-noise hook;
  control_toolkit_b200 optimizers call this when their ``rng`` attribute provides it).

Blocks are produced sequentially from ``numpy.random.default_rng(seed)`` so a fixture only needs the seed.
"""
import numpy as np


class ReplayRNG:
    def __init__(self, seed: int = 1, as_torch: bool = True):
        self.seed = seed
        self._g = np.random.default_rng(seed)
        self.as_torch = as_torch
        self.blocks = []  # (kind, shape) log of every draw, in order

    # -- raw blocks ---------------------------------------------------------------------------
    def standard_draws(self, kind: str, shape) -> np.ndarray:
        shape = tuple(int(s) for s in shape)
        n = int(np.prod(shape))
        if kind == "normal":
            z = self._g.standard_normal(n, dtype=np.float32)
        elif kind == "uniform":
            z = self._g.random(n, dtype=np.float32)
        else:
            raise ValueError(kind)
        self.blocks.append((kind, shape))
        return z.reshape(shape)

    # -- reference-facing duck type -----------------------------------------------------------
    def _wrap(self, a):
        if self.as_torch:
            import torch
            return torch.from_numpy(a)
        return a

    def _scalar(self, x):
        if hasattr(x, "detach"):
            return x.detach()
        return np.float32(x) if not self.as_torch else float(np.float32(x))

    # positional order as in tf.random.Generator (shape, mean, stddev, dtype) / (shape, minval, maxval, dtype): the reference's
    # optimizer_gradient_tf.py:173-178 passes minval / maxval positionally
    def normal(self, shape, mean=0.0, stddev=1.0, dtype=None):
        z = self._wrap(self.standard_draws("normal", shape))
        return z * self._scalar(stddev) + self._scalar(mean)

    def uniform(self, shape, minval=0.0, maxval=1.0, dtype=None):
        z = self._wrap(self.standard_draws("uniform", shape))
        lo, hi = self._scalar(minval), self._scalar(maxval)
        return z * (hi - lo) + lo


class QueueRNG(ReplayRNG):
    """Replays GIVEN blocks of standard draws, in order (e.g. the in-kernel Philox draws of a production tick, exported with
    ctk_philox_export): the oracle then consumes exactly the numbers the CUDA kernels generated."""

    def __init__(self, blocks=(), as_torch: bool = True):
        super().__init__(0, as_torch)
        self.queue = [np.ascontiguousarray(b, dtype=np.float32) for b in blocks]

    def push(self, block):
        self.queue.append(np.ascontiguousarray(block, dtype=np.float32))

    def standard_draws(self, kind: str, shape) -> np.ndarray:
        shape = tuple(int(s) for s in shape)
        if not self.queue:
            raise RuntimeError(f"QueueRNG: no block left for a {kind} draw of shape {shape}")
        z = self.queue.pop(0)
        if z.size != int(np.prod(shape)):
            raise RuntimeError(f"QueueRNG: next block has {z.size} draws, the consumer asks for {shape}")
        self.blocks.append((kind, shape))
        return z.reshape(shape)
