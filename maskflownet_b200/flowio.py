"""Flow file encoders of the reference's benchmark writers (host side, numpy): Middlebury/Sintel `.flo`
(reader/chairs/flo.py:4-25, reader/sintel.py Flo.save) and the KITTI 16-bit encoding (reader/kitti.py:61-62, predict.py:44-66).
Flow arrays are (H, W, 2) in (x, y) = (u, v) order -- what network.predict returns per sample."""
from __future__ import annotations

import numpy as np

FLO_MAGIC = b"PIEH"


def write_flo(path: str, flow_xy: np.ndarray) -> None:
    f = np.ascontiguousarray(flow_xy, dtype=np.float32)
    if f.ndim != 3 or f.shape[2] != 2:
        raise ValueError("write_flo: flow must be (H, W, 2)")
    with open(path, "wb") as fh:
        fh.write(FLO_MAGIC)
        np.array([f.shape[1], f.shape[0]], dtype="<i4").tofile(fh)
        f.astype("<f4").tofile(fh)


def read_flo(path: str) -> np.ndarray:
    with open(path, "rb") as fh:
        if fh.read(4) != FLO_MAGIC:
            raise ValueError(f"{path}: not a .flo file")
        w, h = np.fromfile(fh, dtype="<i4", count=2)
        return np.fromfile(fh, dtype="<f4", count=2 * int(w) * int(h)).reshape(int(h), int(w), 2)


def encode_kitti(flow_xy: np.ndarray, valid: np.ndarray | None = None) -> np.ndarray:
    """(H, W, 2) float flow -> (H, W, 3) uint16 [u*64 + 2^15, v*64 + 2^15, valid] (the PNG payload of the KITTI benchmark;
    inverse of reader/kitti.py:61-62)."""
    f = np.asarray(flow_xy, dtype=np.float64)
    out = np.zeros(f.shape[:2] + (3,), dtype=np.uint16)
    out[..., :2] = np.clip(np.rint(f * 64.0 + 32768.0), 0, 65535).astype(np.uint16)
    out[..., 2] = 1 if valid is None else (np.asarray(valid) > 0).astype(np.uint16)
    return out


def decode_kitti(png: np.ndarray):
    f = (png[..., :2].astype(np.float32) - 32768.0) / 64.0
    return f, png[..., 2] > 0
