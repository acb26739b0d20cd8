"""The `F` namespace (mx.nd / mx.symbol) on torch tensors: an NDArray wrapper with MXNet method semantics plus the
operator functions the reference's model files call (inventory: SURVEY.md section 8b).

Hot operators -> maskflownet_b200.ops (hand-written CUDA):  Correlation, contrib.DeformableConvolution, GridGenerator,
BilinearSampler.  Everything else is shape plumbing / generic math mapped onto torch."""
from __future__ import annotations

import math
import threading
from typing import Sequence

import torch
import torch.nn.functional as tF

from .. import ops as _ops

_state = threading.local()


def current_device():
    return getattr(_state, "device", torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu"))


def set_device(dev):
    _state.device = torch.device(dev)


def _raw(x):
    return x.t if isinstance(x, NDArray) else x


def _wrap(t):
    return NDArray(t) if isinstance(t, torch.Tensor) else t


def _mx_reshape(shape_in: Sequence[int], codes: Sequence[int]):
    """MXNet reshape with the special codes 0 (copy), -1 (infer), -2 (copy the rest), -3 (merge two), -4 (split)."""
    out, i, codes = [], 0, list(codes)
    j = 0
    while j < len(codes):
        c = codes[j]
        if c > 0:
            out.append(c)
            i += 1
        elif c == 0:
            out.append(shape_in[i])
            i += 1
        elif c == -1:
            out.append(-1)
            i += 1
        elif c == -2:
            out.extend(shape_in[i:])
            i = len(shape_in)
        elif c == -3:
            out.append(shape_in[i] * shape_in[i + 1])
            i += 2
        elif c == -4:
            a, b = codes[j + 1], codes[j + 2]
            d = shape_in[i]
            if a == -1:
                a = d // b
            if b == -1:
                b = d // a
            out.extend([a, b])
            i += 1
            j += 2
        else:
            raise ValueError(f"unsupported reshape code {c}")
        j += 1
    return out


class NDArray:
    """Thin wrapper giving a torch.Tensor the method names / keyword conventions of mx.nd.NDArray."""
    __slots__ = ("t",)
    __array_priority__ = 100.0

    def __init__(self, t: torch.Tensor):
        self.t = t

    # --- basic protocol ---
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def dtype(self):
        return self.t.dtype

    @property
    def context(self):
        return self.t.device

    def asnumpy(self):
        return self.t.detach().cpu().numpy()

    def astype(self, dtype, copy=True):
        m = {"float32": torch.float32, "float16": torch.float16, "int32": torch.int32, "uint8": torch.uint8}
        return NDArray(self.t.to(m.get(dtype, dtype)))

    def __repr__(self):
        return f"NDArray{tuple(self.t.shape)}@{self.t.device}"

    def __len__(self):
        return self.t.shape[0]

    def __getitem__(self, idx):
        return NDArray(self.t[idx])

    def backward(self, out_grad=None):
        self.t.backward(_raw(out_grad) if out_grad is not None else torch.ones_like(self.t))

    # --- arithmetic ---
    def _bin(self, other, fn, rev=False):
        o = _raw(other)
        return NDArray(fn(o, self.t) if rev else fn(self.t, o))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return NDArray(_raw(o) - self.t)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return NDArray(_raw(o) / self.t)
    def __pow__(self, o): return NDArray(self.t ** _raw(o))
    def __neg__(self): return NDArray(-self.t)

    # --- MXNet-flavoured methods used by the reference ---
    def reshape(self, *shape, **kw):
        if "shape" in kw:
            shape = kw["shape"]
        elif len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = shape[0]
        return NDArray(self.t.reshape(_mx_reshape(self.t.shape, shape)))

    def flip(self, axis):
        return NDArray(self.t.flip(axis))

    def clip(self, a_min, a_max):
        return NDArray(self.t.clamp(a_min, a_max))

    def slice_axis(self, axis, begin, end):
        return NDArray(self.t.narrow(axis, begin, (self.t.shape[axis] if end is None else end) - begin))

    def squeeze(self, axis=None):
        return NDArray(self.t.squeeze() if axis is None else self.t.squeeze(axis))

    def expand_dims(self, axis):
        return NDArray(self.t.unsqueeze(axis))

    def mean(self, axis=None, exclude=False, keepdims=False):
        return mean(self, axis=axis, exclude=exclude, keepdims=keepdims)

    def sum(self, axis=None, exclude=False, keepdims=False):
        return sum(self, axis=axis, exclude=exclude, keepdims=keepdims)  # noqa: A001

    def tile(self, reps):
        return NDArray(self.t.repeat(*reps))


def array(a, ctx=None, dtype="float32"):
    t = torch.as_tensor(a, dtype=torch.float32)
    return NDArray(t.to(ctx if ctx is not None else current_device()))


# ------------------------------------------------------------------------------------------------------------
# generic helpers
# ------------------------------------------------------------------------------------------------------------
def _axes(x, axis, exclude):
    nd_ = x.dim()
    if axis is None:
        return list(range(nd_))
    ax = [axis] if isinstance(axis, int) else list(axis)
    ax = [a % nd_ for a in ax]
    return [a for a in range(nd_) if a not in ax] if exclude else ax


def sum(data, axis=None, exclude=False, keepdims=False):  # noqa: A001
    x = _raw(data)
    return NDArray(x.sum(dim=_axes(x, axis, exclude), keepdim=keepdims))


def mean(data, axis=None, exclude=False, keepdims=False):
    x = _raw(data)
    return NDArray(x.mean(dim=_axes(x, axis, exclude), keepdim=keepdims))


def concat(*arrays, dim=1):
    return NDArray(torch.cat([_raw(a) for a in arrays], dim=dim))


def expand_dims(data, axis):
    return NDArray(_raw(data).unsqueeze(axis))


def repeat(data, repeats, axis=None):
    x = _raw(data)
    return NDArray(x.flatten().repeat_interleave(repeats) if axis is None else x.repeat_interleave(repeats, dim=axis))


def reshape(data, shape):
    x = _raw(data)
    return NDArray(x.reshape(_mx_reshape(x.shape, shape)))


def reshape_like(lhs, rhs, lhs_begin=None, lhs_end=None, rhs_begin=None, rhs_end=None):
    a, b = _raw(lhs), _raw(rhs)
    lb, le = lhs_begin or 0, a.dim() if lhs_end is None else lhs_end
    rb, re_ = rhs_begin or 0, b.dim() if rhs_end is None else rhs_end
    new = list(a.shape[:lb]) + list(b.shape[rb:re_]) + list(a.shape[le:])
    return NDArray(a.reshape(new))


def pad(data, mode="constant", pad_width=None, constant_value=0.0):
    x = _raw(data)
    pw = list(pad_width)
    # MXNet: (before_0, after_0, before_1, after_1, ...); torch: last dimension first
    tp = []
    for d in range(len(pw) // 2 - 1, -1, -1):
        tp += [pw[2 * d], pw[2 * d + 1]]
    while len(tp) >= 2 and tp[-2:] == [0, 0] and len(tp) > 2 * max(1, x.dim() - 2):
        tp = tp[:-2]
    if mode == "edge":
        return NDArray(tF.pad(x, tp[:4], mode="replicate"))
    if mode == "reflect":
        return NDArray(tF.pad(x, tp[:4], mode="reflect"))
    return NDArray(tF.pad(x, tp, value=constant_value))


def slice(data, begin, end, step=None):  # noqa: A001
    x = _raw(data)
    idx = tuple(builtins_slice(b, e) for b, e in zip(begin, end))
    return NDArray(x[idx])


builtins_slice = __builtins__["slice"] if isinstance(__builtins__, dict) else __builtins__.slice


def slice_axis(data, axis, begin, end):
    return _wrap(data).slice_axis(axis, begin, end) if isinstance(data, NDArray) else NDArray(data).slice_axis(axis, begin, end)


def arange(start, stop=None, step=1.0, dtype="float32"):
    if stop is None:
        start, stop = 0, start
    return NDArray(torch.arange(start, stop, step, dtype=torch.float32, device=current_device()))


def abs(data): return NDArray(_raw(data).abs())  # noqa: A001
def sqrt(data): return NDArray(_raw(data).sqrt())
def square(data): return NDArray(_raw(data).square())
def sigmoid(data): return NDArray(torch.sigmoid(_raw(data)))
def zeros_like(data): return NDArray(torch.zeros_like(_raw(data)))
def ones_like(data): return NDArray(torch.ones_like(_raw(data)))
def broadcast_mul(lhs, rhs): return NDArray(_raw(lhs) * _raw(rhs))
def broadcast_div(lhs, rhs): return NDArray(_raw(lhs) / _raw(rhs))
def broadcast_add(lhs, rhs): return NDArray(_raw(lhs) + _raw(rhs))
def broadcast_sub(lhs, rhs): return NDArray(_raw(lhs) - _raw(rhs))
def BlockGrad(data): return NDArray(_raw(data).detach())
def flip(data, axis): return NDArray(_raw(data).flip(axis))


def add_n(*arrays):
    acc = _raw(arrays[0])
    for a in arrays[1:]:
        acc = acc + _raw(a)
    return NDArray(acc)


def LeakyReLU(data, act_type="leaky", slope=0.25):
    return NDArray(tF.leaky_relu(_raw(data), slope))


def Activation(data, act_type="relu"):
    x = _raw(data)
    fn = {"relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}[act_type]
    return NDArray(fn(x))


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def Convolution(data, weight, bias=None, kernel=None, stride=(1, 1), dilate=(1, 1), pad=(0, 0), num_filter=None,
                num_group=1, no_bias=False, layout="NCHW", **_):
    """Dense convolution (cuDNN through torch) -- out of the hot-path scope (last row)."""
    return NDArray(tF.conv2d(_raw(data), _raw(weight), None if no_bias else _raw(bias), _pair(stride), _pair(pad),
                             _pair(dilate), num_group))


def Deconvolution(data, weight, bias=None, kernel=None, stride=(1, 1), dilate=(1, 1), pad=(0, 0), adj=(0, 0),
                  num_filter=None, num_group=1, no_bias=True, layout="NCHW", **_):
    return NDArray(tF.conv_transpose2d(_raw(data), _raw(weight), None if no_bias else _raw(bias), _pair(stride),
                                       _pair(pad), _pair(adj), num_group, _pair(dilate)))


# ------------------------------------------------------------------------------------------------------------
# hot operators -> hand-written CUDA
# ------------------------------------------------------------------------------------------------------------
def Correlation(data1, data2, kernel_size=1, max_displacement=1, stride1=1, stride2=1, pad_size=0, is_multiply=True,
                **_):
    """mx.nd.Correlation (reference call: network/MaskFlownet.py:193-195)."""
    return NDArray(_ops.correlation(_raw(data1), _raw(data2), pad_size=pad_size, kernel_size=kernel_size,
                                    max_displacement=max_displacement, stride1=stride1, stride2=stride2,
                                    is_multiply=is_multiply))


def GridGenerator(data, transform_type="affine", target_shape=None, **_):
    """mx.nd.GridGenerator (reference call: network/layer.py:17); only transform_type='warp' is on the hot path."""
    if transform_type != "warp":
        raise _ops.MaskflowError("GridGenerator: only transform_type='warp' is implemented (the reference's hot path)")
    return NDArray(_ops.grid_generator_warp(_raw(data)))


def BilinearSampler(data, grid, **_):
    """mx.nd.BilinearSampler (reference call: network/layer.py:18)."""
    return NDArray(_ops.bilinear_sampler(_raw(data), _raw(grid)))


class _Contrib:
    @staticmethod
    def DeformableConvolution(data, offset, weight, bias=None, name=None, kernel=(3, 3), stride=(1, 1), dilate=(1, 1),
                              pad=(0, 0), num_filter=None, num_group=1, num_deformable_group=1, no_bias=False,
                              layout="NCHW", **_):
        """mx.nd.contrib.DeformableConvolution (reference call: network/layer.py:117-124)."""
        return NDArray(_ops.deformable_convolution(_raw(data), _raw(offset), _raw(weight),
                                                   None if bias is None else _raw(bias), kernel=tuple(kernel),
                                                   stride=tuple(stride), dilate=tuple(dilate), pad=tuple(pad),
                                                   num_filter=num_filter, num_group=num_group,
                                                   num_deformable_group=num_deformable_group, no_bias=no_bias,
                                                   layout=layout))

    @staticmethod
    def BilinearResize2D(data, height, width, **_):
        return NDArray(tF.interpolate(_raw(data), size=(height, width), mode="bilinear", align_corners=True))


contrib = _Contrib()
