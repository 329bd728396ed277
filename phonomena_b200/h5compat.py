"""Read-only stand-in for the slice of the h5py API the reference's consumers use, backed by
h5lite.H5Reader -- for machines (like the B200 image) that have no h5py / libhdf5.

The reference's post-processing opens the solver's output with h5py:

    simulation/analysis.py:53-66      h5py.File(file, 'r'); hdf.get(u_id); hdf.attrs["fdx"][:,0,0]; u.shape[3];
                                      u[:, y_index, z_index, :]; u[x_index, y_index, z_index, :]; hdf.close()
    gui/widgets/analysis.py:128,250   with h5py.File(path, mode='r') as hdf: hdf.get('density'); hdf.attrs['x'] ...
    h5py2gif.py:16-24,44              hdf.keys(); u = hdf.get('uz'); u.shape; u[:,:,iz,:]; u[:,:,iz,it]

`sys.modules["h5py"] = phonomena_b200.h5compat` (or `analysis.h5py = h5compat`) makes exactly those calls
work on the files the plugin writes, without touching the reference.  With the real h5py installed this
module is not needed: the files are HDF5.
"""
from __future__ import annotations

import numpy as np

from .h5lite import H5Reader

__all__ = ["File", "Dataset"]


class Dataset:
    """A float64 dataset: `.shape`, `.dtype`, `.ndim`, `len()`, NumPy basic indexing.  Chunked (per-frame)
    datasets are read frame by frame where they lie in the file; only the selected frames are touched."""

    def __init__(self, reader, name):
        self._r, self.name = reader, name
        self.shape = tuple(int(v) for v in reader.shape(name))
        self.dtype = np.dtype("<f8")
        self.ndim = len(self.shape)
        self._chunked = reader.is_chunked(name)

    def __len__(self):
        return self.shape[0]

    def __ne__(self, other):          # `assert u != None` (analysis.py:60)
        return other is not self

    def __eq__(self, other):
        return other is self

    __hash__ = object.__hash__

    def __array__(self, dtype=None, copy=None):
        a = self[...]
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        if not self._chunked:
            return self._r.read(self.name)[idx]
        if not isinstance(idx, tuple):
            idx = (idx,)
        if any(i is Ellipsis for i in idx):
            k = next(n for n, i in enumerate(idx) if i is Ellipsis)
            idx = idx[:k] + (slice(None),) * (self.ndim - (len(idx) - 1)) + idx[k + 1:]
        idx = idx + (slice(None),) * (self.ndim - len(idx))
        if len(idx) != self.ndim:
            raise IndexError("too many indices for a %d-D dataset" % self.ndim)
        head, t = idx[:-1], idx[-1]
        nt = self.shape[-1]
        if isinstance(t, (int, np.integer)):
            tt = int(t) + (nt if t < 0 else 0)
            if not 0 <= tt < nt:
                raise IndexError("frame index %d out of range (%d frames)" % (t, nt))
            return self._r.frame_view(self.name, tt)[head].copy()
        frames = range(nt)[t] if isinstance(t, slice) else [int(v) + (nt if v < 0 else 0) for v in np.asarray(t).reshape(-1)]
        first = self._r.frame_view(self.name, frames[0])[head] if len(frames) else np.zeros(self.shape[:-1])[head]
        out = np.empty(np.shape(first) + (len(frames),), np.float64)
        for n, f in enumerate(frames):
            out[..., n] = self._r.frame_view(self.name, f)[head]
        return out


class File:
    """h5py.File(path, 'r') for an output file of the B200 plugin (read-only)."""

    def __init__(self, name, mode="r", **_kw):
        if mode not in ("r",):
            raise OSError("phonomena_b200.h5compat is read-only (mode %r)" % (mode,))
        self.filename = str(name)
        self._r = H5Reader(self.filename)
        self.attrs = dict(self._r.attrs)
        self._ds = {}

    def keys(self):
        return list(self._r.datasets.keys())

    def __iter__(self):
        return iter(self.keys())

    def __contains__(self, name):
        return name in self._r.datasets

    def get(self, name, default=None):
        if name not in self._r.datasets:
            return default
        if name not in self._ds:
            self._ds[name] = Dataset(self._r, name)
        return self._ds[name]

    def __getitem__(self, name):
        d = self.get(name)
        if d is None:
            raise KeyError(name)
        return d

    def close(self):
        self._r.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
