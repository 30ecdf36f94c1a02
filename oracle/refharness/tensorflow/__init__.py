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


def constant(x, dtype=None, shape=None):
    a = _np.asarray(x)
    if shape is not None:  # tf.constant(value, shape=...): a scalar fills, anything else is reshaped (optimizer_cem_naive_grad_tf.py:106)
        a = _np.full(tuple(shape), a.reshape(-1)[0], a.dtype) if a.size == 1 else a.reshape(tuple(shape))
    if dtype is None and a.dtype.kind in "iu":
        return a  # python-side integer constants (np.tile reps at optimizer_cem_tf.py:86)
    return _torch.as_tensor(a, dtype=dtype or float32)


def zeros(shape, dtype=float32):
    return _torch.zeros(tuple(int(s) for s in (shape if hasattr(shape, "__iter__") else (shape,))), dtype=dtype)


def reshape(x, shape):
    return _torch.as_tensor(_np.asarray(x) if not isinstance(x, _torch.Tensor) else x).reshape(tuple(int(v) for v in shape))


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
    return _torch.cat([x if isinstance(x, _torch.Tensor) else _torch.as_tensor(_np.asarray(x)) for x in xs], dim=axis)


def squeeze(x):
    return _torch.squeeze(x)


class _Math:
    @staticmethod
    def reduce_mean(x, axis=None, keepdims=False):
        return _torch.mean(x, dim=axis, keepdim=keepdims)

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

    def normal(self, shape, mean=0.0, stddev=1.0, dtype=float32):
        return _torch.randn(tuple(shape), generator=self._g, dtype=dtype) * stddev + mean

    def uniform(self, shape, minval=0.0, maxval=1.0, dtype=float32):
        return _torch.rand(tuple(shape), generator=self._g, dtype=dtype) * (maxval - minval) + minval


class _Random:
    Generator = _Generator


random = _Random()


# ----------------------------------------------------------------------------------------------------------------------
# Additions for the reference's UNMODIFIED ``Optimizers/optimizer_gradient_tf.py`` / ``optimizer_random_action_tf.py``:
# tf.Variable, tf.GradientTape, tf.clip_by_norm, tf.zeros_like and the LEGACY Keras Adam (``get_weights()`` returns
# ``[iterations, m, v]`` once slots exist, ``[]`` before the first apply_gradients), with TF's documented semantics:
# * ``tf.clip_by_norm(t, c, axes)`` = t * c / max(||t||_2, c), norm over ``axes`` with keepdims;
# * Keras Adam ``_resource_apply_dense``: t = iterations + 1; lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);
#   m += (g - m)(1 - b1); v += (g^2 - v)(1 - b2); var -= lr_t * m / (sqrt(v) + eps); iterations += 1.
# ----------------------------------------------------------------------------------------------------------------------
class _Var(_torch.Tensor):
    """tf.Variable stand-in: a leaf torch tensor with ``assign`` (in place) whose slices are new tensors."""

    __torch_function__ = _torch._C._disabled_torch_function_impl

    @staticmethod
    def __new__(cls, initial_value=None, trainable=True, dtype=float32, **_):
        return _torch.Tensor._make_subclass(cls, _torch.as_tensor(initial_value, dtype=dtype).detach().clone(), False)

    def __init__(self, initial_value=None, trainable=True, dtype=float32, **_):
        pass

    def assign(self, value):
        with _torch.no_grad():
            self.requires_grad_(False)
            self.copy_(_torch.as_tensor(value, dtype=self.dtype).detach())
        return self

    def __getitem__(self, idx):
        return _torch.Tensor.__getitem__(self.as_subclass(_torch.Tensor), idx).clone()

    def numpy(self):
        return self.as_subclass(_torch.Tensor).detach().numpy().copy()


def _unwrap(x):
    return x.as_subclass(_torch.Tensor) if isinstance(x, _Var) else x


Variable = _Var
_plain_clip_by_value = clip_by_value


def clip_by_value(x, lo, hi):  # noqa: F811  (accepts tf.Variable)
    return _plain_clip_by_value(_unwrap(x), lo, hi)


def zeros_like(x):
    return _torch.zeros_like(_unwrap(x) if not isinstance(x, (int, _np.integer)) else _torch.tensor(x))


def clip_by_norm(t, clip_norm, axes=None):
    t = _unwrap(t)
    n = _torch.sqrt(_torch.sum(t * t, dim=tuple(axes) if axes is not None else None, keepdim=True))
    c = _torch.as_tensor(clip_norm, dtype=t.dtype)
    return t * c / _torch.maximum(n, c)


class GradientTape:
    def __init__(self, watch_accessed_variables=True, persistent=False):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, var):
        var.requires_grad_(True)  # leaf: the ops inside the tape build the autograd graph on it

    def gradient(self, target, source):
        (g,) = _torch.autograd.grad(_torch.sum(target), source)
        source.requires_grad_(False)
        return g.as_subclass(_torch.Tensor) if isinstance(g, _Var) else g


class _KerasAdam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **_):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), float(beta_1), float(beta_2), float(epsilon)
        self.iterations = 0
        self._m = self._v = None

    def apply_gradients(self, grads_and_vars):
        for g, var in grads_and_vars:
            g = g.detach()
            if self._m is None:
                self._m, self._v = _torch.zeros_like(g), _torch.zeros_like(g)
            t = self.iterations + 1
            f32 = _np.float32
            lr_t = f32(f32(self.lr) * _np.sqrt(f32(1) - _np.power(f32(self.b2), f32(t))) / (f32(1) - _np.power(f32(self.b1), f32(t))))
            self._m = self._m + (g - self._m) * f32(1 - self.b1)
            self._v = self._v + (g * g - self._v) * f32(1 - self.b2)
            var.assign(_unwrap(var).detach() - float(lr_t) * self._m / (_torch.sqrt(self._v) + f32(self.eps)))
        self.iterations += 1

    def get_weights(self):
        if self._m is None:
            return []
        return [_np.int64(self.iterations), self._m.numpy().copy(), self._v.numpy().copy()]

    def set_weights(self, weights):
        if len(weights) == 0:
            return
        self.iterations = int(_np.asarray(weights[0]))
        self._m = _torch.as_tensor(_np.asarray(weights[1]), dtype=float32).clone()
        self._v = _torch.as_tensor(_np.asarray(weights[2]), dtype=float32).clone()


class _KerasOptimizers:
    Adam = _KerasAdam


class _Keras:
    optimizers = _KerasOptimizers()


keras = _Keras()
