"""Minimal MRC2014 stack I/O (mode 2, float32) -- the two `mrcfile` calls the reference uses.

The reference only needs `mrcfile.mmap(path, 'r')` -> object with `.data` ([nz, ny, nx]
memmap) and `.close()` (Distances.py:75,150) and `mrcfile.new_mmap(path, shape, mrc_mode=2,
overwrite=True)` (Distances.py:85).  `mrcfile` itself is not a dependency of this package.
Header layout: SURVEY.md appendix B.
"""
from __future__ import annotations

import os
import numpy as np

_HEADER_BYTES = 1024


class MrcStack:
    def __init__(self, data, fh=None):
        self.data = data
        self._fh = fh

    def close(self):
        if isinstance(self.data, np.memmap):
            self.data.flush() if self.data.flags.writeable else None
        self.data = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _read_header(path):
    with open(path, 'rb') as f:
        raw = f.read(_HEADER_BYTES)
    if len(raw) < _HEADER_BYTES:
        raise ValueError('%s: shorter than an MRC header' % path)
    words = np.frombuffer(raw, dtype='<i4', count=56)
    nx, ny, nz, mode = (int(v) for v in words[:4])
    nsymbt = int(words[23])
    if mode != 2:
        raise ValueError('%s: MRC mode %d not supported (only mode 2, float32)' % (path, mode))
    if nx <= 0 or ny <= 0 or nz <= 0 or nsymbt < 0:
        raise ValueError('%s: bad MRC dimensions %s' % (path, (nx, ny, nz, nsymbt)))
    return nx, ny, nz, nsymbt


def mmap(path, mode='r'):
    """Open an MRC mode-2 stack as a read-only (or 'r+') memmap of shape [nz, ny, nx]."""
    nx, ny, nz, nsymbt = _read_header(path)
    need = _HEADER_BYTES + nsymbt + 4 * nx * ny * nz
    if os.path.getsize(path) < need:
        raise ValueError('%s: truncated (%d < %d bytes)' % (path, os.path.getsize(path), need))
    data = np.memmap(path, dtype='<f4', mode='r' if mode == 'r' else 'r+',
                     offset=_HEADER_BYTES + nsymbt, shape=(nz, ny, nx))
    return MrcStack(data)


def new_mmap(path, shape, mrc_mode=2, overwrite=False):
    """Create an MRC mode-2 stack of `shape` = (nz, ny, nx) and return a writable memmap."""
    if mrc_mode != 2:
        raise ValueError('only mrc_mode=2 (float32) is supported')
    if os.path.exists(path) and not overwrite:
        raise ValueError('%s exists' % path)
    nz, ny, nx = (int(v) for v in shape)
    hdr = np.zeros(_HEADER_BYTES // 4, dtype='<i4')
    hdr[0:3] = (nx, ny, nz)
    hdr[3] = 2
    hdr[7:10] = (nx, ny, nz)
    hdr[10:13] = np.array([nx, ny, nz], dtype='<f4').view('<i4')
    hdr[13:16] = np.array([90., 90., 90.], dtype='<f4').view('<i4')
    hdr[16:19] = (1, 2, 3)
    hdr[27] = 20140
    raw = bytearray(hdr.tobytes())
    raw[208:212] = b'MAP '
    raw[212:216] = bytes([0x44, 0x44, 0x00, 0x00])
    with open(path, 'wb') as f:
        f.write(raw)
        f.truncate(_HEADER_BYTES + 4 * nx * ny * nz)
    data = np.memmap(path, dtype='<f4', mode='r+', offset=_HEADER_BYTES, shape=(nz, ny, nx))
    return MrcStack(data)


def write_stack(path, stack):
    """Write a float32 [nz, ny, nx] array as an MRC mode-2 stack."""
    stack = np.ascontiguousarray(stack, dtype=np.float32)
    m = new_mmap(path, stack.shape, overwrite=True)
    m.data[:] = stack
    m.close()
