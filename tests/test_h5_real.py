"""Runs only where the real h5py / libhdf5 exist (not in the B200 image, which has neither; there the files
are checked by the specification-level validator tests/h5check.py): open a plugin file with h5py itself and
read it the way the reference's consumers do."""
import numpy as np
import pytest

h5py = pytest.importorskip("h5py")
if getattr(h5py, "__stub__", False) or not hasattr(h5py, "Dataset"):      # oracle/refshim.py's in-memory stand-in
    pytest.skip("h5py is the refshim stand-in, not the library", allow_module_level=True)


def test_libhdf5_opens_plugin_output(tmp_path):
    from tests.test_dropin_cpu import _write_like_plugin
    p = str(tmp_path / "real.h5")
    nx, ny, frames = 7, 6, 70
    _write_like_plugin(p, nx, ny, 5, frames)
    with h5py.File(p, "r") as hdf:
        assert sorted(hdf.keys()) == ["density", "ux", "uy", "uz"]
        u = hdf.get("uz")
        assert u.shape == (nx, ny, 1, frames) and u.chunks == (nx, ny, 1, 1) and u.dtype == np.float64
        n0 = (nx - 1) * ny + nx * (ny - 1)
        for t in (0, 33, frames - 1):
            assert np.array_equal(u[:, :, 0, t].reshape(-1), t * 1e6 + n0 + np.arange(nx * ny))
        assert hdf.attrs["dt"] == 2.5e-5 and hdf.attrs["fdx"][:, 0, 0].shape == (nx - 1,) and int(hdf.attrs["steps"]) == frames
        name = hdf.attrs["prim_material"]
        assert (name.decode() if isinstance(name, bytes) else name) == "Gallium Arsenide"
