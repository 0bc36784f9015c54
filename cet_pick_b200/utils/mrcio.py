"""Minimal MRC2014 writer/reader for the `<name>_hm.mrc` output of save_detection
(tomo_det.py:61-67, which uses the `mrcfile` package: mode 2 float32, data (nz, ny, nx))."""
from __future__ import annotations

import struct

import numpy as np


def write_mrc(path, data: np.ndarray):
    data = np.ascontiguousarray(data, dtype=np.float32)
    if data.ndim != 3:
        raise ValueError("write_mrc expects a 3-D array")
    nz, ny, nx = data.shape
    hdr = bytearray(1024)
    struct.pack_into("<10i", hdr, 0, nx, ny, nz, 2, 0, 0, 0, nx, ny, nz)
    struct.pack_into("<6f", hdr, 40, float(nx), float(ny), float(nz), 90.0, 90.0, 90.0)
    struct.pack_into("<3i", hdr, 64, 1, 2, 3)
    struct.pack_into("<3f", hdr, 76, float(data.min()), float(data.max()), float(data.mean()))
    struct.pack_into("<2i", hdr, 88, 0, 0)            # ispg, nsymbt
    hdr[104:108] = b"\x00\x00\x00\x00"
    struct.pack_into("<i", hdr, 108, 20140)           # nversion
    hdr[208:212] = b"MAP "
    hdr[212:216] = bytes([0x44, 0x44, 0x00, 0x00])    # little-endian machine stamp
    struct.pack_into("<f", hdr, 216, float(data.std()))
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(data.tobytes())


_MODES = {0: "i1", 1: "<i2", 2: "<f4", 6: "<u2"}


def read_mrc(path) -> np.ndarray:
    """(nz, ny, nx) array of an MRC file, like `mrcfile.open(path).data` (modes 0, 1, 2, 6)."""
    with open(path, "rb") as f:
        hdr = f.read(1024)
        nx, ny, nz, mode = struct.unpack_from("<4i", hdr, 0)
        nsymbt = struct.unpack_from("<i", hdr, 92)[0]
        if mode not in _MODES:
            raise ValueError(f"MRC mode {mode} is not supported (0, 1, 2, 6 are)")
        dt = np.dtype(_MODES[mode])
        f.seek(1024 + nsymbt)
        return np.frombuffer(f.read(nx * ny * nz * dt.itemsize), dtype=dt).reshape(nz, ny, nx)
