"""TEST-ONLY TensorFlow-*API* shim over torch-CPU so that the reference's UNMODIFIED
``Optimizers/optimizer_cem_tf.py`` (hard ``import tensorflow as tf`` at :5) can be executed when golden
vectors are generated.  TensorFlow itself is not installable here (no network).  Only the calls that file
makes are provided, with TF's documented semantics:

* ``tf.argsort(x)``        ascending; implemented by TF as top_k(-x) => ties keep the lower index first (stable);
* ``tf.math.reduce_std``   population std (ddof 0), sqrt(mean((x-mean)^2));
* ``tf.clip_by_value``     min(max(x, lo), hi);
* ``tf.gather(x, i, axis=0)`` row gather; ``tf.tile``; ``tf.concat``; ``tf.squeeze``; ``tf.ensure_shape`` (checked).
Tensors are torch tensors (they have ``.numpy()``).
"""
import numpy as _np
import torch as _torch

float32 = _torch.float32
int32 = _torch.int32
int64 = _torch.int64
Tensor = _torch.Tensor
Variable = _torch.Tensor


def convert_to_tensor(x, dtype=float32):
    return _torch.as_tensor(_np.asarray(x), dtype=dtype)


def constant(x, dtype=None):
    a = _np.asarray(x)
    if dtype is None and a.dtype.kind in "iu":
        return a  # python-side integer constants (np.tile reps at optimizer_cem_tf.py:86)
    return _torch.as_tensor(a, dtype=dtype or float32)


def zeros(shape, dtype=float32):
    return _torch.zeros(tuple(int(s) for s in (shape if hasattr(shape, "__iter__") else (shape,))), dtype=dtype)


def ones(shape, dtype=float32):
    return _torch.ones(tuple(int(s) for s in shape), dtype=dtype)


def tile(x, reps):
    return x.repeat(*[int(r) for r in reps])


def multiply(a, b):
    return _torch.mul(a, b)


def clip_by_value(x, lo, hi):
    lo = _torch.as_tensor(lo, dtype=x.dtype)
    hi = _torch.as_tensor(hi, dtype=x.dtype)
    return _torch.minimum(_torch.maximum(x, lo), hi)


def ensure_shape(x, shape):
    assert tuple(x.shape) == tuple(shape), (tuple(x.shape), tuple(shape))
    return x


def argsort(x, axis=-1):
    return _torch.argsort(x, dim=axis, stable=True)


def gather(x, idx, axis=0):
    return _torch.index_select(x, axis, idx)


def reduce_mean(x, axis=None, keepdims=False):
    return _torch.mean(x, dim=axis, keepdim=keepdims)


def concat(xs, axis):
    return _torch.cat(list(xs), dim=axis)


def squeeze(x):
    return _torch.squeeze(x)


class _Math:
    @staticmethod
    def reduce_std(x, axis=None, keepdims=False):
        mu = _torch.mean(x, dim=axis, keepdim=True)
        return _torch.sqrt(_torch.mean((x - mu) * (x - mu), dim=axis, keepdim=keepdims))


math = _Math()


class _Generator:
    """tf.random.Generator stand-in (Philox in real TF; its stream cannot be reproduced offline,
    parity is defined under injected noise only)."""

    def __init__(self, seed):
        self._g = _torch.Generator().manual_seed(int(seed) % (2 ** 63))

    @classmethod
    def from_seed(cls, seed):
        return cls(seed)

    def normal(self, shape, dtype=float32, mean=0.0, stddev=1.0):
        return _torch.randn(tuple(shape), generator=self._g, dtype=dtype) * stddev + mean

    def uniform(self, shape, dtype=float32, minval=0.0, maxval=1.0):
        return _torch.rand(tuple(shape), generator=self._g, dtype=dtype) * (maxval - minval) + minval


class _Random:
    Generator = _Generator


random = _Random()
