"""Minimal eager stand-in for the TensorFlow-1.12 API surface the reference's stage-1 code touches
(TEST INFRASTRUCTURE ONLY).

Purpose: TensorFlow 1.12 cannot be installed here, but the reference's *own, unmodified* sources
(/root/reference/utils/model.py, models/networks/{__init__,layers,vgg}.py, models/detector_translator_model.py)
only use a small set of TF ops.  Registering this module as ``tensorflow`` lets those files execute eagerly on torch
CPU tensors, so that tests/golden/make_golden.py can record what the reference's wiring (layer order, scopes and
variable names, concat order, filter halving, (x,y) order, loss composition) produces.  The numeric semantics of each
op are the restatements in oracle/tf_ops.py — those stay "parity unpinned"; what this pins is everything ABOVE the ops.

Variables are looked up by their TF name in ``PARAMS`` (a dict installed by the caller): a name the reference's
scoping produces but the dict lacks raises KeyError, i.e. naming mismatches are caught, not papered over.
"""
import contextlib
import types

import numpy as np
import torch

from . import tf_ops as _T

PARAMS = {}          # name -> torch tensor (installed by the caller)
UPDATES = []         # (name, new value) batch-norm moving-average updates, in graph order
ACCESSED = set()     # variable names the reference's scopes asked for
DTYPE = torch.float64
AUTO_REUSE = object()
float32 = "float32"

_scope = []
_uid = [0]


class _Shape(tuple):
    def as_list(self):
        return list(self)


class Tensor:
    """Thin wrapper so that `.shape.as_list()`, slicing and arithmetic behave like tf.Tensor."""
    __array_priority__ = 1000

    def __init__(self, t, name=None):
        self.t = t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t), dtype=DTYPE)
        if self.t.is_floating_point() and self.t.dtype != DTYPE:
            self.t = self.t.to(DTYPE)
        self.name = name or "Tensor_%d:0" % _uid[0]
        _uid[0] += 1

    @property
    def shape(self):
        return _Shape(self.t.shape)

    def __getitem__(self, idx):
        return Tensor(self.t[idx])

    def _bin(self, other, fn, rev=False):
        o = other.t if isinstance(other, Tensor) else torch.as_tensor(np.asarray(other), dtype=DTYPE)
        return Tensor(fn(o, self.t) if rev else fn(self.t, o))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __neg__(self): return Tensor(-self.t)
    def __pow__(self, o): return Tensor(self.t ** o)


def _t(x):
    return x.t if isinstance(x, Tensor) else torch.as_tensor(np.asarray(x), dtype=DTYPE)


# ---- scoping ----------------------------------------------------------------------------------------
@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _scope.append(name)
    try:
        yield
    finally:
        _scope.pop()


def _full(name):
    return "/".join(_scope + [name])


def _var(name):
    ACCESSED.add(name)
    return PARAMS[name].to(DTYPE)


# ---- math ops ----------------------------------------------------------------------------------------
def to_float(x): return Tensor(_t(x).to(DTYPE))
def linspace(a, b, n): return Tensor(torch.tensor(a, dtype=DTYPE) + torch.arange(n, dtype=DTYPE) * ((b - a) / (n - 1) if n > 1 else 0.0))
def reshape(x, shape): return Tensor(_t(x).reshape([int(s) for s in shape]))
def expand_dims(x, axis): return Tensor(_t(x).unsqueeze(axis))
def square(x): return Tensor(_t(x) ** 2)
def exp(x): return Tensor(torch.exp(_t(x)))
def abs(x): return Tensor(torch.abs(_t(x)))          # noqa: A001
def transpose(x, perm): return Tensor(_t(x).permute(*perm))
def concat(values=None, axis=0, **kw):
    if isinstance(values, int):                       # tf.concat(axis=3, values=[...]) keyword form handled below
        values, axis = axis, values
    return Tensor(torch.cat([_t(v) for v in values], dim=axis))
def stack(values, axis=0): return Tensor(torch.stack([_t(v) for v in values], dim=axis))
def split(value=None, num_or_size_splits=None, axis=0, **kw):
    return [Tensor(c) for c in torch.chunk(_t(value), num_or_size_splits, dim=axis)]
def tile(x, multiples): return Tensor(_t(x).repeat(*multiples))
def shape(x): return list(_t(x).shape)
def ones_like(x): return Tensor(torch.ones_like(_t(x)))
def zeros_like(x): return Tensor(torch.zeros_like(_t(x)))
def clip_by_value(x, lo, hi): return Tensor(_t(x).clamp(lo, hi))
def pad(x, paddings):
    p = [int(v) for pair in reversed(paddings) for v in pair]
    return Tensor(torch.nn.functional.pad(_t(x), p))
def constant(value, name=None): return Tensor(torch.as_tensor(np.asarray(value), dtype=DTYPE))


def _reduce(fn):
    def f(x, axis=None, **kw):
        if isinstance(x, (list, tuple)):
            x = torch.stack([_t(v) for v in x], dim=0)
        else:
            x = _t(x)
        if axis is None:
            return Tensor(fn(x))
        r = fn(x, dim=axis)
        return Tensor(r[0] if isinstance(r, tuple) else r)
    return f


reduce_mean = _reduce(torch.mean)
reduce_sum = _reduce(torch.sum)
reduce_max = _reduce(torch.max)


# ---- namespaces ----------------------------------------------------------------------------------------
nn = types.SimpleNamespace(
    relu=lambda x: Tensor(torch.relu(_t(x))),
    sigmoid=lambda x: Tensor(torch.sigmoid(_t(x))),
    leaky_relu=lambda x, alpha=0.2: Tensor(_T.leaky_relu(_t(x), alpha)),
    softmax=lambda x, axis=-1: Tensor(torch.softmax(_t(x), dim=axis)),
    max_pool=lambda bottom, ksize, strides, padding, name=None: Tensor(_T.max_pool_2x2(_t(bottom))),
    conv2d=lambda bottom, filt, strides, padding: Tensor(_T.conv2d(_t(bottom), _t(filt), None, strides[1], 0)),
    bias_add=lambda x, b: Tensor(_t(x) + _t(b)),
    sigmoid_cross_entropy_with_logits=lambda labels=None, logits=None: Tensor(
        _T.sigmoid_cross_entropy_with_logits(_t(logits), _t(labels))),
)


def _layers_conv2d(inputs, filters, padding, kernel_size, kernel_initializer=None, strides=1, use_bias=True):
    assert padding == 'same'
    with variable_scope("conv2d"):
        w = _var(_full("kernel"))
        assert w.shape[0] == kernel_size and w.shape[3] == filters, (_full("kernel"), tuple(w.shape), kernel_size, filters)
        b = _var(_full("bias")) if use_bias else None
    return Tensor(_T.conv2d(_t(inputs), w, b, strides, 0))


layers = types.SimpleNamespace(conv2d=_layers_conv2d)


def _contrib_batch_norm(x, epsilon=1e-3, center=True, scale=True, scope=None, is_training=True, decay=0.999):
    with variable_scope(scope):
        names = [_full(n) for n in ("gamma", "beta", "moving_mean", "moving_variance")]
    g, b, mm, mv = [_var(n) for n in names]
    y, nmm, nmv = _T.batch_norm(_t(x), g, b, mm, mv, bool(is_training), eps=epsilon, decay=decay)
    if is_training:
        UPDATES.append((names[2], nmm))
        UPDATES.append((names[3], nmv))
    return Tensor(y)


contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
    batch_norm=_contrib_batch_norm, xavier_initializer=lambda: None, fully_connected=None))


def _resize_images(x, size):
    return Tensor(_T.resize_bilinear_legacy(_t(x), int(size[0]), int(size[1])))


image = types.SimpleNamespace(resize_images=_resize_images)


# ---- the training/summary plumbing of DetectorTranslatorModel.build(): inert stand-ins -----------------
class _Optimizer:
    def __init__(self, *a, **k): pass
    def minimize(self, *a, **k): return None


class _Var:
    def __init__(self, name): self.name = name


def trainable_variables():
    return [_Var(n + ":0") for n in PARAMS if "moving" not in n and not n.startswith("vgg")]


def get_collection(key): return []
@contextlib.contextmanager
def control_dependencies(deps): yield


train = types.SimpleNamespace(AdamOptimizer=_Optimizer,
                              exponential_decay=lambda lr, step, decay_steps, decay_rate: _T.exponential_decay(
                                  lr, float(step), decay_steps, decay_rate))
GraphKeys = types.SimpleNamespace(UPDATE_OPS="update_ops")
summary = types.SimpleNamespace(image=lambda *a, **k: None, scalar=lambda *a, **k: None, merge=lambda *a, **k: None,
                                FileWriter=lambda *a, **k: None)
logging = types.SimpleNamespace(info=lambda *a, **k: None)
