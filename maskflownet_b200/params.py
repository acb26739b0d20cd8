"""Reader for the reference's shipped MXNet checkpoints (`weights/*.params`, written by `save_parameters`,
network/pipeline.py:52-54) -> torch `state_dict` of maskflownet_b200.network models.

File format (NDArray-list, little endian): u64 0x112, u64 reserved, u64 n_arrays, then per array
u32 0xF993FAC9, i32 storage type (0 = dense), u32 ndim, i64 dims[ndim], i32 dev_type, i32 dev_id, i32 type flag
(0 = float32), raw data; then u64 n_names and per name u64 length + bytes.  Names are gluon parameter names such as
`hybridsequential3_conv2aweight`, `deform5weight`; the cascade's head parameters carry a `maskflownet_s0_` prefix.
"""
from __future__ import annotations

import re
import struct
from typing import Dict

import numpy as np
import torch

_MAGIC_LIST, _MAGIC_ND = 0x112, 0xF993FAC9
_DTYPES = {0: np.float32, 1: np.float64, 2: np.float16, 3: np.uint8, 4: np.int32, 5: np.int8, 6: np.int64}
_NAME_RE = re.compile(r"(dc_conv\d|conv\d_\d|conv\d[a-z]|upfeat\d|pred_flow\d|pred_mask\d|deform\d)_?(weight|bias)$")


def read_params(path: str) -> Dict[str, np.ndarray]:
    """Parse an MXNet .params file into {gluon name: array} (exact-length parse; raises on any inconsistency)."""
    buf = open(path, "rb").read()
    off = 0

    def take(fmt):
        nonlocal off
        vals = struct.unpack_from("<" + fmt, buf, off)
        off += struct.calcsize("<" + fmt)
        return vals if len(vals) > 1 else vals[0]

    if take("Q") != _MAGIC_LIST:
        raise ValueError(f"{path}: not an MXNet NDArray-list file")
    take("Q")
    n = take("Q")
    arrays = []
    for _ in range(n):
        if take("I") != _MAGIC_ND:
            raise ValueError(f"{path}: bad NDArray magic at byte {off - 4}")
        if take("i") != 0:
            raise ValueError(f"{path}: sparse storage is not supported")
        ndim = take("I")
        dims = [take("q") for _ in range(ndim)]
        take("ii")
        dt = _DTYPES[take("i")]
        cnt = int(np.prod(dims)) if dims else 1
        arrays.append(np.frombuffer(buf, dtype=dt, count=cnt, offset=off).reshape(dims).copy())
        off += cnt * np.dtype(dt).itemsize
    names = []
    if take("Q") != n:
        raise ValueError(f"{path}: name count does not match array count")
    for _ in range(n):
        ln = take("Q")
        names.append(buf[off:off + ln].decode())
        off += ln
    if off != len(buf):
        raise ValueError(f"{path}: {len(buf) - off} trailing bytes")
    return dict(zip(names, arrays))


def gluon_to_module_name(gluon_name: str, cascade: bool = False) -> str:
    """`[arg:][maskflownet_s0_]hybridsequential7_conv3bweight` -> `conv3b.weight` (`MaskFlownet_S.conv3b.weight` for
    head parameters inside a cascade checkpoint)."""
    name = gluon_name.split(":", 1)[-1]
    m = _NAME_RE.search(name)
    if not m:
        raise KeyError(f"unrecognised parameter name {gluon_name!r}")
    base = f"{m.group(1)}.{m.group(2)}"
    if cascade and "maskflownet_s" in name:
        return "MaskFlownet_S." + base
    return base


def load_checkpoint(model: torch.nn.Module, path: str, strict: bool = True) -> torch.nn.Module:
    """Load a shipped checkpoint into MaskFlownetS / MaskFlownet (layouts coincide: Conv2D (O,I,kh,kw),
    Conv2DTranspose (I,O,kh,kw), DeformableConv2D weight (F,C,3,3))."""
    raw = read_params(path)
    own = dict(model.named_parameters())
    cascade = any(k.startswith("MaskFlownet_S.") for k in own)
    seen = set()
    with torch.no_grad():
        for gname, arr in raw.items():
            key = gluon_to_module_name(gname, cascade)
            if key not in own:
                if strict:
                    raise KeyError(f"{gname} -> {key} has no counterpart in {type(model).__name__}")
                continue
            if tuple(own[key].shape) != tuple(arr.shape):
                raise ValueError(f"{key}: checkpoint shape {arr.shape} != model shape {tuple(own[key].shape)}")
            own[key].copy_(torch.from_numpy(arr.astype(np.float32)))
            seen.add(key)
    missing = set(own) - seen
    if strict and missing:
        raise KeyError(f"parameters missing from {path}: {sorted(missing)[:5]} ...")
    return model
