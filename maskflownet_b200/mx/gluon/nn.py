"""mxnet.gluon.nn stand-in: HybridBlock / HybridSequential / Conv2D / Conv2DTranspose / LeakyReLU / Activation with
gluon's construction protocol (name_scope, prefix, params.get, deferred input channels) on torch parameters.

Enough of the protocol is reproduced for network/MaskFlownet.py and network/layer.py of the reference to construct
and run unchanged; parameter names follow gluon's prefix rules so that shipped checkpoints can be matched by name."""
from __future__ import annotations

import contextlib
import math
from collections import OrderedDict

import torch

from .. import ndarray as F_ns
from ..ndarray import NDArray

_name_counters = {}
_scope_stack = []  # prefixes of enclosing name_scope()s


def _auto_prefix(hint):
    k = _name_counters.get(hint, 0)
    _name_counters[hint] = k + 1
    return f"{hint}{k}_"


def reset_name_counters():
    _name_counters.clear()


class Parameter:
    """A named, possibly shape-deferred tensor (dims given as 0 are inferred at first forward)."""

    def __init__(self, name, shape=None, init=None, allow_deferred_init=False):
        self.name = name
        self.shape = tuple(shape) if shape is not None else None
        self.init = init
        self.grad_req = "write"
        self._data = None

    @property
    def deferred(self):
        return self._data is None

    def set_data(self, t):
        self._data = t if isinstance(t, torch.Tensor) else torch.as_tensor(t)
        self.shape = tuple(self._data.shape)
        self._data.requires_grad_(self.grad_req != "null")

    def data(self, ctx=None):
        if self._data is None:
            raise RuntimeError(f"parameter {self.name} has not been initialised")
        return NDArray(self._data)

    def grad(self, ctx=None):
        return NDArray(self._data.grad)


class ParameterDict(OrderedDict):
    def __init__(self, prefix=""):
        super().__init__()
        self.prefix = prefix

    def get(self, name, shape=None, init=None, allow_deferred_init=False, **_):  # gluon: self.params.get('weight', ...)
        full = self.prefix + name
        if full not in self:
            self[full] = Parameter(full, shape, init, allow_deferred_init)
        return self[full]


class Block:
    def __init__(self, prefix=None, params=None):
        hint = type(self).__name__.lower()
        parent = _scope_stack[-1] if _scope_stack else ""
        self._prefix = parent + (prefix if prefix is not None else _auto_prefix(hint))
        self._params = ParameterDict(self._prefix)
        self._children = OrderedDict()
        self._reg = OrderedDict()  # attribute name -> Parameter (passed to hybrid_forward as keyword arguments)

    @property
    def prefix(self):
        return self._prefix

    @property
    def params(self):
        return self._params

    @contextlib.contextmanager
    def name_scope(self):
        _scope_stack.append(self._prefix)
        try:
            yield
        finally:
            _scope_stack.pop()

    def __setattr__(self, name, value):
        if isinstance(value, Block):
            self.__dict__.setdefault("_children", OrderedDict())[name] = value
        elif isinstance(value, Parameter):
            self.__dict__.setdefault("_reg", OrderedDict())[name] = value
        object.__setattr__(self, name, value)

    def register_child(self, block, name=None):
        self._children[name or str(len(self._children))] = block

    def collect_params(self):
        out = ParameterDict(self._prefix)
        seen = set()

        def visit(b):
            if id(b) in seen:
                return
            seen.add(id(b))
            for k, p in b._params.items():
                out[k] = p
            for c in b._children.values():
                visit(c)
        visit(self)
        return out

    def hybridize(self, *a, **k):
        return None  # nothing to trace: the shim executes eagerly (CUDA graphs are the caller's tool)

    # -- initialisation: MSRAPrelu(slope) like the reference (network/pipeline.py:26), seeded, on `device` --
    def initialize(self, init=None, ctx=None, seed=0, device=None, slope=0.1, **_):
        self._init_cfg = (seed, torch.device(device or ctx or F_ns.current_device()), slope)
        gen = torch.Generator().manual_seed(seed)
        for p in self.collect_params().values():
            p._gen, p._slope, p._device = gen, slope, self._init_cfg[1]
            if p.shape is not None and all(d > 0 for d in p.shape):
                _materialise(p)

    def load_dict(self, tensors, device=None, strict=True):
        """tensors: gluon parameter name -> array (see maskflownet_b200.params)."""
        dev = torch.device(device or F_ns.current_device())
        mine = self.collect_params()
        for k, p in mine.items():
            if k in tensors:
                p.set_data(torch.as_tensor(tensors[k], dtype=torch.float32).to(dev))
            elif strict:
                raise KeyError(f"missing parameter {k}")

    def __call__(self, *args):
        first = next((a for a in args if isinstance(a, NDArray)), None)
        if first is not None:
            F_ns.set_device(first.t.device)
        return self.forward(*args)

    def forward(self, *args):
        raise NotImplementedError


def _materialise(p):
    shape = p.shape
    gen, slope, dev = getattr(p, "_gen", None), getattr(p, "_slope", 0.1), getattr(p, "_device", F_ns.current_device())
    if p.name.endswith("bias") or len(shape) < 2:
        t = torch.zeros(shape)
    else:
        # MSRAPrelu: N(0, sqrt(2 / ((1 + slope^2) * fan_avg)))  with factor_type='avg'
        hw = 1
        for d in shape[2:]:
            hw *= d
        fan_in, fan_out = shape[1] * hw, shape[0] * hw
        std = math.sqrt(2.0 / ((1 + slope ** 2) * (fan_in + fan_out) / 2.0))
        t = torch.randn(shape, generator=gen) * std
    p.set_data(t.to(dev))


class HybridBlock(Block):
    def infer_param_shapes(self, x):
        """Resolve deferred input-channel dims (value 0) from the first input: weight (O, I, kh, kw) <- I = x.shape[1]."""
        for p in self._reg.values():
            if p.deferred:
                if p.shape is None:
                    raise RuntimeError(f"cannot infer shape of {p.name}")
                shp = list(p.shape)
                if len(shp) >= 2 and shp[1] == 0:
                    shp[1] = x.shape[1]
                if any(d <= 0 for d in shp):
                    raise RuntimeError(f"cannot infer shape {shp} of {p.name}")
                p.shape = tuple(shp)
                _materialise(p)

    def forward(self, x, *args):
        if self._reg:
            if any(p.deferred for p in self._reg.values()):
                self.infer_param_shapes(x)
            kw = {k: p.data() for k, p in self._reg.items()}
            return self.hybrid_forward(F_ns, x, *args, **kw)
        return self.hybrid_forward(F_ns, x, *args)

    def hybrid_forward(self, F, x, *args, **kwargs):
        raise NotImplementedError


class HybridSequential(HybridBlock):
    def __init__(self, prefix=None, params=None):
        super().__init__(prefix=prefix, params=params)
        self._seq = []

    def add(self, *blocks):
        for b in blocks:
            self._seq.append(b)
            self.register_child(b)

    def forward(self, x, *args):
        for b in self._seq:
            x = b(x)
        return x

    def __getitem__(self, i):
        return self._seq[i]

    def __len__(self):
        return len(self._seq)


Sequential = HybridSequential


def _tup(v, n=2):
    return (v,) * n if isinstance(v, int) else tuple(v)


class _ConvBase(HybridBlock):
    _hint = "conv"
    _transposed = False

    def __init__(self, channels, kernel_size, strides=1, padding=0, dilation=1, groups=1, layout="NCHW",
                 activation=None, use_bias=True, weight_initializer=None, bias_initializer="zeros", in_channels=0,
                 output_padding=0, prefix=None, params=None):
        if prefix is None:
            prefix = _auto_prefix(self._hint)
        super().__init__(prefix=prefix, params=params)
        self._channels, self._in = channels, in_channels
        self._k, self._s, self._p, self._d = _tup(kernel_size), _tup(strides), _tup(padding), _tup(dilation)
        self._groups, self._op = groups, _tup(output_padding)
        wshape = ((in_channels, channels // groups) if self._transposed else (channels, in_channels // groups if in_channels else 0)) + self._k
        self.weight = self.params.get("weight", shape=wshape, init=weight_initializer, allow_deferred_init=True)
        self.bias = self.params.get("bias", shape=(channels,), init=bias_initializer) if use_bias else None
        self._act = activation

    def infer_param_shapes(self, x):
        for p in self._reg.values():
            if p.deferred:
                shp = list(p.shape)
                if len(shp) == 4:
                    if self._transposed and shp[0] == 0:
                        shp[0] = x.shape[1]
                    if not self._transposed and shp[1] == 0:
                        shp[1] = x.shape[1] // self._groups
                p.shape = tuple(shp)
                _materialise(p)


class Conv2D(_ConvBase):
    """Dense convolution: delegated to cuDNN through torch (outside the hot-path scope)."""
    _hint = "conv"

    def hybrid_forward(self, F, x, weight, bias=None):
        y = torch.nn.functional.conv2d(x.t, weight.t, None if bias is None else bias.t, self._s, self._p, self._d,
                                       self._groups)
        return NDArray(y) if self._act is None else F.Activation(NDArray(y), self._act)


class Conv2DTranspose(_ConvBase):
    _hint = "conv"
    _transposed = True

    def hybrid_forward(self, F, x, weight, bias=None):
        y = torch.nn.functional.conv_transpose2d(x.t, weight.t, None if bias is None else bias.t, self._s, self._p,
                                                 self._op, self._groups, self._d)
        return NDArray(y) if self._act is None else F.Activation(NDArray(y), self._act)


class LeakyReLU(HybridBlock):
    def __init__(self, alpha, **kw):
        super().__init__(**kw)
        self._alpha = alpha

    def hybrid_forward(self, F, x):
        return F.LeakyReLU(x, act_type="leaky", slope=self._alpha)


class Activation(HybridBlock):
    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self._act_type = activation

    def hybrid_forward(self, F, x):
        return F.Activation(x, act_type=self._act_type)
