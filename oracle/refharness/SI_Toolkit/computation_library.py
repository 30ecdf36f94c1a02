"""TEST-ONLY shim of ``SI_Toolkit.computation_library`` (the real package is an un-vendored, un-pinned
git submodule of the reference: reference README.md:19-26).  It provides just the ``self.lib.*`` surface
the reference's hot-path files use (SURVEY.md section 8c lists it), backed by torch-CPU fp32, so that the
reference's ``optimizer_mppi.py`` / ``optimizer_rpgd.py`` / ``optimizer_cem_tf.py`` / ``controller_mpc.py``
execute UNMODIFIED when golden vectors are generated (oracle/gen_golden.py).  Never imported by the product.

Semantics follow TensorFlow where the libraries differ, because the north star pins parity to the TF
implementation:
* ``argsort``  : ascending, stable (tf.argsort == top_k(-x): ties -> lower index first);
* ``reduce_std``: population std (ddof 0);
* ``clip_by_norm(t, c, axes)`` : t * c / max(||t||_2, c), norm over ``axes`` with keepdims;
* ``to_variable`` returns a tensor whose ``__getitem__`` COPIES (slicing a tf.Variable yields a new tensor;
  plain torch basic indexing would alias ``optimizer_rpgd.py:426`` ``u_nom`` to ``Q_tf`` which is overwritten at ``:515``).
"""
import numpy as np
import torch


class _TFLikeVariable(torch.Tensor):
    """A tensor that can be assigned in place but whose slices are copies (tf.Variable read semantics)."""

    __torch_function__ = torch._C._disabled_torch_function_impl

    @staticmethod
    def __new__(cls, data):
        return torch.Tensor._make_subclass(cls, data.detach().clone(), False)

    def __getitem__(self, idx):
        return torch.Tensor.__getitem__(self.as_subclass(torch.Tensor), idx).clone()

    def numpy(self):
        return self.as_subclass(torch.Tensor).numpy().copy()


class ComputationLibrary:
    lib = None
    float32 = torch.float32
    int32 = torch.int32
    int64 = torch.int64
    bool = torch.bool
    newaxis = None
    pi = float(np.pi)

    # ---- conversion -------------------------------------------------------------------------
    @staticmethod
    def to_tensor(x, dtype=torch.float32):
        if isinstance(x, torch.Tensor):
            return x.as_subclass(torch.Tensor).to(dtype)
        return torch.as_tensor(np.asarray(x), dtype=dtype)

    @staticmethod
    def constant(x, dtype=torch.float32):
        return ComputationLibrary.to_tensor(x, dtype)

    @staticmethod
    def to_numpy(x):
        if isinstance(x, torch.Tensor):
            return x.detach().as_subclass(torch.Tensor).cpu().numpy().copy()
        return np.asarray(x)

    @staticmethod
    def to_variable(x, dtype=torch.float32):
        return _TFLikeVariable(ComputationLibrary.to_tensor(x, dtype))

    @staticmethod
    def assign(v, x):
        with torch.no_grad():
            v.as_subclass(torch.Tensor).copy_(torch.as_tensor(x).detach())
        return v

    @staticmethod
    def cast(x, dtype):
        return x.to(dtype)

    # ---- shape ------------------------------------------------------------------------------
    @staticmethod
    def reshape(x, shape):
        return torch.reshape(x, tuple(shape))

    @staticmethod
    def permute(x, perm):
        return x.permute(*perm)

    @staticmethod
    def squeeze(x):
        return torch.squeeze(x)

    @staticmethod
    def ndim(x):
        return x.ndim

    @staticmethod
    def concat(xs, axis):
        return torch.cat(list(xs), dim=axis)

    @staticmethod
    def stack(xs, axis=0):
        return torch.stack(list(xs), dim=axis)

    @staticmethod
    def tile(x, reps):
        return x.repeat(*[int(r) for r in reps])

    @staticmethod
    def gather(x, idx, axis=0):
        return torch.index_select(x.as_subclass(torch.Tensor), axis, idx)

    # ---- creation ---------------------------------------------------------------------------
    @staticmethod
    def zeros(shape, dtype=torch.float32):
        return torch.zeros(tuple(shape), dtype=dtype)

    @staticmethod
    def ones(shape, dtype=torch.float32):
        return torch.ones(tuple(shape), dtype=dtype)

    @staticmethod
    def zeros_like(x):
        return torch.zeros_like(torch.as_tensor(x))

    @staticmethod
    def arange(*a):
        return torch.arange(*a)

    # ---- math -------------------------------------------------------------------------------
    @staticmethod
    def clip(x, lo, hi):
        return torch.minimum(torch.maximum(x, torch.as_tensor(lo, dtype=x.dtype)), torch.as_tensor(hi, dtype=x.dtype))

    @staticmethod
    def clip_by_norm(x, clip_norm, axes):
        l2 = torch.sqrt(torch.sum(x * x, dim=tuple(axes), keepdim=True))
        return x * clip_norm / torch.maximum(l2, torch.as_tensor(clip_norm, dtype=x.dtype))

    @staticmethod
    def sum(x, axis=None):
        return torch.sum(x, dim=axis)

    @staticmethod
    def mean(x, axis=None):
        return torch.mean(x, dim=axis)

    @staticmethod
    def reduce_mean(x, axis=None, keepdims=False):
        return torch.mean(x, dim=axis, keepdim=keepdims)

    @staticmethod
    def reduce_std(x, axis=None, keepdims=False):
        mu = torch.mean(x, dim=axis, keepdim=True)
        return torch.sqrt(torch.mean((x - mu) ** 2, dim=axis, keepdim=keepdims))

    @staticmethod
    def reduce_min(x, axis=None):
        return torch.amin(x, dim=axis)

    @staticmethod
    def argsort(x, axis=0):
        return torch.argsort(x, dim=axis, stable=True)

    @staticmethod
    def matmul(a, b):
        return torch.matmul(a, b)

    exp = staticmethod(torch.exp)
    abs = staticmethod(torch.abs)
    cos = staticmethod(torch.cos)
    sin = staticmethod(torch.sin)
    sqrt = staticmethod(torch.sqrt)

    @staticmethod
    def ceil(x):
        return np.ceil(x)

    @staticmethod
    def floor(x):
        return np.floor(x)

    @staticmethod
    def nan_to_num(x, nan=0.0):
        return torch.nan_to_num(x, nan=nan)

    # ---- device -----------------------------------------------------------------------------
    @staticmethod
    def set_device(device_name):
        def decorator(f):
            return f
        return decorator


class NumpyLibrary(ComputationLibrary):
    lib = "Numpy"


class TensorFlowLibrary(ComputationLibrary):
    """TF-API *semantics* over torch-CPU (TensorFlow is not installable offline)."""
    lib = "TF"


class PyTorchLibrary(ComputationLibrary):
    lib = "Pytorch"


ComputationClasses = (NumpyLibrary, TensorFlowLibrary, PyTorchLibrary)
TensorType = torch.Tensor
VariableType = torch.Tensor
