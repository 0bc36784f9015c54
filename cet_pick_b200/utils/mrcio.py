"""Minimal MRC2014 writer/reader for the `<name>_hm.mrc` output of save_detection
(tomo_det.py:61-67, which uses the `mrcfile` package: mode 2 float32, data (nz, ny, nx))."""
from __future__ import annotations

import struct

import numpy as np


def write_mrc(path, data: np.ndarray, stats=None):
    """stats = (min, max, mean, rms) of `data` when the caller already has them (the detector computes them on
    the device next to the heat-map); otherwise they are computed here like mrcfile does."""
    data = np.ascontiguousarray(data, dtype=np.float32)
    if data.ndim != 3:
        raise ValueError("write_mrc expects a 3-D array")
    nz, ny, nx = data.shape
    if stats is None:
        stats = (float(data.min()), float(data.max()), float(data.mean()), float(data.std()))
    hdr = bytearray(1024)
    struct.pack_into("<10i", hdr, 0, nx, ny, nz, 2, 0, 0, 0, nx, ny, nz)
    struct.pack_into("<6f", hdr, 40, float(nx), float(ny), float(nz), 90.0, 90.0, 90.0)
    struct.pack_into("<3i", hdr, 64, 1, 2, 3)
    struct.pack_into("<3f", hdr, 76, float(stats[0]), float(stats[1]), float(stats[2]))
    struct.pack_into("<2i", hdr, 88, 0, 0)            # ispg, nsymbt
    hdr[104:108] = b"\x00\x00\x00\x00"
    struct.pack_into("<i", hdr, 108, 20140)           # nversion
    hdr[208:212] = b"MAP "
    hdr[212:216] = bytes([0x44, 0x44, 0x00, 0x00])    # little-endian machine stamp
    struct.pack_into("<f", hdr, 216, float(stats[3]))
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(memoryview(data).cast("B"))


_MODES = {0: "i1", 1: "i2", 2: "f4", 6: "u2", 12: "f2"}


def read_mrc(path) -> np.ndarray:
    """(nz, ny, nx) array of an MRC file, like `mrcfile.open(path, permissive=True).data` (modes 0, 1, 2, 6, 12;
    either byte order, taken from the machine stamp at byte 212, falling back to whichever order gives a sane
    header like mrcfile's permissive mode)."""
    import os
    with open(path, "rb") as f:
        hdr = f.read(1024)
        if len(hdr) < 1024:
            raise ValueError(f"{path}: not an MRC file (header is {len(hdr)} bytes, expected 1024)")
        stamp = hdr[212]
        order = "<" if stamp == 0x44 else ">" if stamp == 0x11 else None
        if order is None:                              # bad stamp: take the order under which the mode is known
            order = "<" if struct.unpack_from("<i", hdr, 12)[0] in _MODES else ">"
        nx, ny, nz, mode = struct.unpack_from(order + "4i", hdr, 0)
        nsymbt = struct.unpack_from(order + "i", hdr, 92)[0]
        if mode not in _MODES:
            raise ValueError(f"{path}: MRC mode {mode} is not supported (0, 1, 2, 6, 12 are)")
        if min(nx, ny, nz) <= 0 or nsymbt < 0:
            raise ValueError(f"{path}: bad MRC header (nx, ny, nz, nsymbt = {nx}, {ny}, {nz}, {nsymbt})")
        dt = np.dtype(_MODES[mode]).newbyteorder(order) if mode != 0 else np.dtype("i1")
        need = 1024 + nsymbt + nx * ny * nz * dt.itemsize
        size = os.fstat(f.fileno()).st_size
        if size < need:
            raise ValueError(f"{path}: truncated MRC file ({size} bytes, header describes {need})")
        f.seek(1024 + nsymbt)
        a = np.frombuffer(f.read(nx * ny * nz * dt.itemsize), dtype=dt).reshape(nz, ny, nx)
        return a if order == "<" or mode == 0 else a.astype(dt.newbyteorder("<"))
