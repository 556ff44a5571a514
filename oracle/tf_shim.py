"""NumPy stand-ins for the TensorFlow / TFP / GPflow symbols the reference's hot-path files touch.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Why: TensorFlow 2.1, TensorFlow-Probability 0.9 and GPflow 2.0.0-rc1 (reference README.md:9-11)
cannot be installed offline in this image, so ``fourier_features.py``, ``acquisitions/pes.py`` and
``acquisitions/dts.py`` cannot be imported as shipped.  ``install()`` registers three tiny modules
named ``tensorflow``, ``tensorflow_probability`` and ``gpflow`` in ``sys.modules`` whose functions
are the float64 NumPy equivalents of the eager TF ops those files call (matmul/transpose/cos/
inv/argmax/gather_nd/scatter_nd/bincount/...).  With it the *unmodified reference sources* execute
here, which is how ``tests/golden/make_golden.py`` produces the golden vectors that pin the oracle.
Semantics mirrored: eager execution, float64, ``tf.linalg.inv`` raising
``tf.errors.InvalidArgumentError`` on a singular input, ``tf.scatter_nd`` summing duplicate indices,
``tf.math.argmax`` returning the lowest index on ties.  Random streams are NumPy's (seed with
``seed()``); parity is always defined on the arrays actually drawn, never on TF's RNG stream.
Autodiff for the one place the reference needs it (``optimizer.minimize(loss, var_list=[x_star])``, ff.py:150):
``_RMSprop.minimize`` re-runs the reference's own ``loss`` closure with the variable replaced by a torch float64
leaf -- every shim op and every operator on ``Tensor`` forwards to torch when it meets a torch operand -- takes
``torch.autograd`` gradients and applies the update of ``tf.keras.optimizers.RMSprop`` as TF 2.1 defines it for
momentum = 0, centered = False (optimizer_v2/rmsprop.py, python path): ``rms <- rho rms + (1 - rho) g^2``,
``var <- var - lr g / (sqrt(rms) + epsilon)`` (lr = 1e-3, epsilon = 1e-7 by default), then the variable's
``constraint``.  The gradient, the loop, the loss bookkeeping and the early stop are therefore the reference's own
code; only this update rule is restated from the TF sources (which are not available offline).
"""
from __future__ import annotations

import sys
import types

import numpy as np

try:                                                    # torch is only needed for the autodiff of ``minimize``
    import torch as _torch
except Exception:                                       # pragma: no cover
    _torch = None

_rng = np.random.default_rng(0)
_variables = []          # every tf.Variable created, in order (lets a test recover ff.py's x_star)


def seed(s):
    global _rng
    _rng = np.random.default_rng(s)
    _variables.clear()


def created_variables():
    return list(_variables)


def _is_torch(x):
    return _torch is not None and isinstance(x, _torch.Tensor)


def _as_torch(x):
    """torch float64 view of a shim tensor / array; a traced Variable maps to its autograd leaf."""
    if _is_torch(x):
        return x
    leaf = getattr(x, "_leaf", None)
    if leaf is not None:
        return leaf
    return _torch.from_numpy(np.array(np.asarray(x), dtype=np.float64))


def _binary(name, reflected=False):
    base = getattr(np.ndarray, name)

    def op(self, other):
        if _is_torch(other) or getattr(self, "_leaf", None) is not None or getattr(other, "_leaf", None) is not None:
            a, b = _as_torch(self), _as_torch(other)
            fn = getattr(_torch.Tensor, name.replace("__r", "__") if reflected else name)
            return fn(b, a) if reflected else fn(a, b)
        return base(self, other)
    return op


class Tensor(np.ndarray):
    """ndarray with the two eager-tensor members the reference uses: ``.numpy()`` and ``.assign``.  Binary operators
    forward to torch when the other operand is a torch tensor (the autodiff pass of ``_RMSprop.minimize``)."""
    __array_priority__ = 1000

    def numpy(self):
        return np.asarray(self)

    def assign(self, value):
        self[...] = np.asarray(value)
        return self

    def __neg__(self):
        leaf = getattr(self, "_leaf", None)
        return -leaf if leaf is not None else np.ndarray.__neg__(self)


for _n in ("__add__", "__sub__", "__mul__", "__truediv__", "__matmul__"):
    setattr(Tensor, _n, _binary(_n))
for _n in ("__radd__", "__rsub__", "__rmul__", "__rtruediv__", "__rmatmul__"):
    setattr(Tensor, _n, _binary(_n, reflected=True))


def _t(a, dtype=None):
    arr = np.asarray(a.numpy() if hasattr(a, "numpy") else a)
    if dtype is not None:
        arr = arr.astype(dtype)
    return arr.view(Tensor)


class _Variable(Tensor):
    def __new__(cls, initial_value=None, constraint=None, dtype=None, shape=None, **_):
        arr = np.array(np.asarray(initial_value), dtype=np.float64 if dtype is None else dtype, copy=True)
        obj = arr.view(cls)
        obj._constraint = constraint
        obj._leaf = None                               # torch autograd leaf while minimize() traces the loss
        obj.initial_value = arr.copy()                 # what the reference initialised it with (ff.py:141-144)
        _variables.append(obj)
        return obj

    def __array_finalize__(self, obj):
        self._constraint = getattr(obj, "_constraint", None)
        self._leaf = None

    def __getitem__(self, item):
        # slices stay writable views so `expected_log[:, i].assign(...)` (pes.py:119) lands in place
        return np.ndarray.__getitem__(self, item).view(Tensor)


class InvalidArgumentError(Exception):
    pass


def _inv(a):
    try:
        return _t(np.linalg.inv(np.asarray(a)))
    except np.linalg.LinAlgError as err:      # "Input is not invertible."
        raise InvalidArgumentError(str(err))


def _gather_nd(params, indices):
    idx = np.asarray(indices)
    return _t(np.asarray(params)[tuple(idx[..., i] for i in range(idx.shape[-1]))])


def _scatter_nd(indices, updates, shape):
    out = np.zeros(tuple(int(s) for s in np.asarray(shape).reshape(-1)), dtype=np.asarray(updates).dtype)
    idx = np.asarray(indices)
    np.add.at(out, tuple(idx[..., i] for i in range(idx.shape[-1])), np.asarray(updates))
    return _t(out)


def _logsumexp(a, axis=None):
    a = np.asarray(a)
    m = np.max(a, axis=axis, keepdims=True)
    out = m + np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True))
    return _t(np.squeeze(out) if axis is None else np.squeeze(out, axis=axis))


def _bincount(arr, minlength=0, dtype=np.int64, **_):
    return _t(np.bincount(np.asarray(arr).reshape(-1), minlength=minlength).astype(dtype))


def build_modules():
    tf = types.ModuleType("tensorflow")
    tf.float64, tf.float32, tf.int32, tf.int64 = np.float64, np.float32, np.int32, np.int64
    tf.dtypes = types.SimpleNamespace(float64=np.float64, int32=np.int32, int64=np.int64)
    tf.Variable = _Variable
    tf.function = lambda f=None, **kw: f if f is not None else (lambda g: g)
    tf.constant = lambda v, dtype=None, **kw: _t(v, dtype)
    tf.cast = lambda x, dtype: x if _is_torch(x) else _t(x, dtype)
    tf.sqrt = lambda x: _t(np.sqrt(x))
    tf.cos = lambda x: _torch.cos(x) if _is_torch(x) else _t(np.cos(np.asarray(x)))
    tf.exp = lambda x: _t(np.exp(np.asarray(x)))
    tf.zeros = lambda shape, dtype=np.float32: _t(np.zeros(shape, dtype=dtype))
    tf.ones = lambda shape, dtype=np.float32: _t(np.ones(shape, dtype=dtype))
    tf.expand_dims = lambda x, axis: _t(np.expand_dims(np.asarray(x), axis))
    tf.squeeze = lambda x, axis=None: _t(np.squeeze(np.asarray(x), axis=axis))
    tf.tile = lambda x, multiples: _t(np.tile(np.asarray(x), multiples))
    tf.transpose = lambda x: _t(np.transpose(np.asarray(x)))
    tf.stack = lambda vals, axis=0: _t(np.stack([np.asarray(v) for v in vals], axis=axis))
    tf.concat = lambda vals, axis: _t(np.concatenate([np.asarray(v) for v in vals], axis=axis))
    tf.range = lambda n, dtype=np.int32: _t(np.arange(n, dtype=dtype))
    tf.argmax = lambda x, axis=0, output_type=np.int64: _t(np.argmax(np.asarray(x), axis=axis).astype(output_type))
    tf.gather_nd = lambda params, indices: _gather_nd(params, indices)
    tf.scatter_nd = lambda indices, updates, shape: _scatter_nd(indices, updates, shape)
    tf.reduce_mean = lambda x: _torch.mean(x) if _is_torch(x) else _t(np.mean(np.asarray(x)))
    tf.reduce_sum = lambda x, axis=None: _t(np.sum(np.asarray(x), axis=axis))
    tf.reduce_logsumexp = lambda x, axis=None: _logsumexp(x, axis)
    tf.clip_by_value = lambda x, lo, hi: _t(np.clip(np.asarray(x), lo, hi))
    tf.cond = lambda pred, true_fn, false_fn: true_fn() if bool(pred) else false_fn()
    tf.linalg = types.SimpleNamespace(
        matrix_transpose=lambda x: (_torch.swapaxes(_as_torch(x), -1, -2)
                                    if (_is_torch(x) or getattr(x, "_leaf", None) is not None)
                                    else _t(np.swapaxes(np.asarray(x), -1, -2))),
        inv=_inv)
    tf.math = types.SimpleNamespace(
        argmax=lambda x, axis=0, output_type=np.int64: _t(np.argmax(np.asarray(x), axis=axis).astype(output_type)),
        bincount=_bincount,
        log=lambda x: _t(np.log(np.asarray(x))),
        exp=lambda x: _t(np.exp(np.asarray(x))),
        square=lambda x: _t(np.square(np.asarray(x))))
    tf.random = types.SimpleNamespace(
        normal=lambda shape, mean=0.0, stddev=1.0, dtype=np.float32:
            _t(_rng.standard_normal(tuple(shape)) * np.asarray(stddev) + mean, dtype),
        uniform=lambda shape, minval=0, maxval=None, dtype=np.float32:
            _t(_rng.uniform(minval, 1.0 if maxval is None else maxval, size=tuple(shape)), dtype))
    tf.errors = types.SimpleNamespace(InvalidArgumentError=InvalidArgumentError)

    class _RMSprop:                                   # constructed at ff.py:139 as RMSprop(rho=0.0)
        """TF 2.1 tf.keras.optimizers.RMSprop, momentum = 0, centered = False (see the module docstring)."""

        def __init__(self, learning_rate=0.001, rho=0.9, momentum=0.0, epsilon=1e-7, centered=False, **kw):
            assert momentum == 0.0 and not centered, "only the plain RMSprop path is restated"
            self.lr, self.rho, self.epsilon = float(learning_rate), float(rho), float(epsilon)
            self.rms = {}

        def minimize(self, loss, var_list):
            if _torch is None:
                raise NotImplementedError("tf_shim.minimize needs torch for autodiff")
            for v in var_list:
                v._leaf = _torch.from_numpy(np.array(np.asarray(v), dtype=np.float64)).requires_grad_(True)
            try:
                value = loss()                         # the reference's closure, executed on torch tensors
                grads = _torch.autograd.grad(value, [v._leaf for v in var_list])
            finally:
                for v in var_list:
                    v._leaf = None
            for v, g in zip(var_list, grads):
                g = g.numpy()
                rms = self.rho * self.rms.get(id(v), np.zeros_like(g)) + (1.0 - self.rho) * g * g
                self.rms[id(v)] = rms
                new = np.asarray(v) - self.lr * g / (np.sqrt(rms) + self.epsilon)
                if v._constraint is not None:
                    new = np.asarray(v._constraint(_t(new)))
                v[...] = new

    tf.keras = types.SimpleNamespace(
        backend=types.SimpleNamespace(epsilon=lambda: 1e-7),
        optimizers=types.SimpleNamespace(RMSprop=_RMSprop))

    tfp = types.ModuleType("tensorflow_probability")

    class _MVNDiag:                                   # ff.py:68-71
        def __init__(self, loc, scale_diag):
            self.loc, self.scale = np.asarray(loc), np.asarray(scale_diag)

        def sample(self, count):
            return _t(self.loc + self.scale * _rng.standard_normal((int(count), self.loc.shape[0])))

    tfp.distributions = types.SimpleNamespace(MultivariateNormalDiag=_MVNDiag)

    gpflow = types.ModuleType("gpflow")

    class Product:                                    # gpflow.kernels.base.Product (ff.py:33)
        def __init__(self, kernels):
            self.kernels = list(kernels)

    def _mvnquad(*a, **kw):
        raise NotImplementedError("gpflow.quadrature is outside the hot path")

    gpflow.kernels = types.SimpleNamespace(base=types.SimpleNamespace(Product=Product), Product=Product)
    gpflow.quadrature = types.SimpleNamespace(mvnquad=_mvnquad)
    return tf, tfp, gpflow


def install():
    """Register the shim modules unless real ones are importable. Returns (tf, tfp, gpflow)."""
    mods = build_modules()
    for m in mods:
        sys.modules.setdefault(m.__name__, m)
    return tuple(sys.modules[m.__name__] for m in mods)


# ---- duck-typed model pieces handed to the reference functions -------------------------------
class Param:
    """GPflow Parameter stand-in: ``.numpy()`` returns the value (ff.py:38,42)."""

    def __init__(self, value):
        self.value = np.asarray(value, dtype=np.float64)

    def numpy(self):
        return self.value


class RBFKernel:
    def __init__(self, lengthscale):
        self.lengthscale = Param(lengthscale)     # GPflow 2.0.0-rc1 spells it singular


class StubModel:
    """kernel / q_mu / q_sqrt as read by ff.py:96-97, and a ``predict_f_samples`` that returns
    pre-drawn samples (pes.py:106, rank_pes.py:28): shape (S, N, 1), object with ``.numpy()``."""

    def __init__(self, kernel=None, q_mu=None, q_sqrt=None, fixed_samples=None):
        self.kernel = kernel
        self.q_mu = None if q_mu is None else _t(q_mu)
        self.q_sqrt = None if q_sqrt is None else _t(q_sqrt)
        self.fixed_samples = fixed_samples
        self.calls = 0

    def predict_f_samples(self, x, num_samples):
        fs = self.fixed_samples
        if callable(fs):
            out = fs(np.asarray(x), int(num_samples), self.calls)
        else:
            out = np.asarray(fs)
        self.calls += 1
        out = np.asarray(out, dtype=np.float64)
        assert out.shape[0] == num_samples and out.shape[1] == np.asarray(x).shape[0]
        return _t(out.reshape(out.shape[0], out.shape[1], 1))
