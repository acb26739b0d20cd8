"""ctypes front-end of the C oracle (oracle/mfn_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (maskflownet_b200) never does.

All functions take / return contiguous float32 numpy arrays in NCHW layout.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmfn_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Returns the path of the shared object."""
    src = os.path.join(_HERE, "mfn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.mfn_ref_num_threads.restype = ctypes.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_f32p)


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def num_threads() -> int:
    return int(lib().mfn_ref_num_threads())


def correlation_forward(data1, data2, pad_size=4, kernel_size=1, max_displacement=4, stride1=1,
                        stride2=1, is_multiply=1, threads=1) -> np.ndarray:
    """MXNet F.Correlation (reference call sites network/MaskFlownet.py:193-195, 440-441)."""
    d1, d2 = _c(data1), _c(data2)
    assert d1.shape == d2.shape and d1.ndim == 4
    N, C, H, W = d1.shape
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib().mfn_ref_correlation_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1,
                                             stride2, ctypes.byref(oc), ctypes.byref(oh),
                                             ctypes.byref(ow))
    if rc:
        raise ValueError("invalid correlation parameters")
    out = np.empty((N, oc.value, oh.value, ow.value), np.float32)
    rc = lib().mfn_ref_correlation_forward(_p(d1), _p(d2), _p(out), N, C, H, W, pad_size,
                                           kernel_size, max_displacement, stride1, stride2,
                                           int(is_multiply), int(threads))
    if rc:
        raise RuntimeError(f"mfn_ref_correlation_forward failed: {rc}")
    return out


def correlation_backward(grad_out, data1, data2, max_displacement=4, threads=1):
    go, d1, d2 = _c(grad_out), _c(data1), _c(data2)
    N, C, H, W = d1.shape
    g1, g2 = np.empty_like(d1), np.empty_like(d2)
    rc = lib().mfn_ref_correlation_backward(_p(go), _p(d1), _p(d2), _p(g1), _p(g2), N, C, H, W,
                                            max_displacement, int(threads))
    if rc:
        raise RuntimeError(f"mfn_ref_correlation_backward failed: {rc}")
    return g1, g2


def deformable_conv_forward(data, offset, weight, bias=None, kernel=(3, 3), stride=(1, 1),
                            pad=(1, 1), dilate=(1, 1), num_group=1, num_deformable_group=1,
                            border_mode=0, threads=1) -> np.ndarray:
    """MXNet F.contrib.DeformableConvolution (reference call site network/layer.py:117-124)."""
    x, off, w = _c(data), _c(offset), _c(weight)
    b = _c(bias) if bias is not None else None
    N, C, H, W = x.shape
    F = w.shape[0]
    kh, kw = kernel
    OH = (H + 2 * pad[0] - (dilate[0] * (kh - 1) + 1)) // stride[0] + 1
    OW = (W + 2 * pad[1] - (dilate[1] * (kw - 1) + 1)) // stride[1] + 1
    assert off.shape == (N, 2 * kh * kw * num_deformable_group, OH, OW), off.shape
    assert w.shape == (F, C // num_group, kh, kw), w.shape
    out = np.empty((N, F, OH, OW), np.float32)
    rc = lib().mfn_ref_deformable_conv_forward(
        _p(x), _p(off), _p(w), _p(b) if b is not None else None, _p(out), N, C, H, W, F, kh, kw,
        stride[0], stride[1], pad[0], pad[1], dilate[0], dilate[1], num_group,
        num_deformable_group, int(border_mode), int(threads))
    if rc:
        raise RuntimeError(f"mfn_ref_deformable_conv_forward failed: {rc}")
    return out


def upsample(x, factor: int) -> np.ndarray:
    """Reference Upsample(f) block (network/MaskFlownet.py:35-62)."""
    a = _c(x)
    N, C, H, W = a.shape
    out = np.empty((N, C, H * factor, W * factor), np.float32)
    rc = lib().mfn_ref_upsample(_p(a), _p(out), N * C, H, W, int(factor))
    if rc:
        raise RuntimeError(f"mfn_ref_upsample failed: {rc}")
    return out


def grid_generator_warp(flow_xy) -> np.ndarray:
    """MXNet F.GridGenerator(transform_type='warp') (reference call site network/layer.py:17)."""
    f = _c(flow_xy)
    N, two, H, W = f.shape
    assert two == 2
    grid = np.empty_like(f)
    lib().mfn_ref_grid_generator_warp(_p(f), _p(grid), N, H, W)
    return grid


def bilinear_sampler(data, grid) -> np.ndarray:
    """MXNet F.BilinearSampler (reference call site network/layer.py:18)."""
    d, g = _c(data), _c(grid)
    N, C, H, W = d.shape
    _, _, OH, OW = g.shape
    out = np.empty((N, C, OH, OW), np.float32)
    lib().mfn_ref_bilinear_sampler(_p(d), _p(g), _p(out), N, C, H, W, OH, OW)
    return out


def reconstruction2d(x, flow_yx) -> np.ndarray:
    """layer.Reconstruction2D (network/layer.py:8-18): flip (y,x)->(x,y), GridGenerator, sampler."""
    f = _c(flow_yx)[:, ::-1].copy()
    return bilinear_sampler(x, grid_generator_warp(f))
