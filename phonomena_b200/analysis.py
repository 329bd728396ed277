"""Post-processing of the solver's HDF5 output without h5py: the Hann-windowed spectrum of
phonomena/simulation/analysis.py:44-96 (1-D along t, or 2-D along x and t), reading the file with
h5lite.H5Reader.  Same arguments and return values as the reference's `spectrum`; with a
surface-only file (z extent 1) `z_index` must be 0, as in the reference's consumers."""
from __future__ import annotations

import numpy as np

from .h5lite import H5Reader
from .hostmath import nonlinspace


def spectrum(file, u_id, z_index, y_index, x_index=None):
    """analysis.py:44-96.  `file`: path or H5Reader."""
    r = file if isinstance(file, H5Reader) else H5Reader(file)
    if u_id not in r.datasets:
        raise KeyError(u_id)
    shape = r.shape(u_id)
    x = nonlinspace(np.asarray(r.attrs["fdx"])[:, 0, 0]) if u_id == "ux" else np.array(r.attrs["x"])
    frames = int(r.attrs.get("frames_written", shape[3]))
    dt = float(r.attrs["dt"]) * int(r.attrs.get("record_every", 1))
    N = frames
    Nf = N // 2
    f = np.fft.fftfreq(N, d=dt)[:Nf]
    window = np.hanning(N)
    # (x, t) line at fixed y, z: read frame by frame (one chunk per frame)
    line = np.stack([r.read(u_id, frame=t)[:, y_index, z_index] for t in range(N)], axis=1)
    if x_index is None:
        dft = np.abs(np.fft.fft2(line * window, norm="ortho"))[:, :Nf]
    else:
        dft = np.abs(np.fft.fft(line[x_index] * window, norm="ortho"))[:Nf]
    return x, f, dft
