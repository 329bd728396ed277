"""Line probes + on-device spectra (include/phb200.h "line probes") against the
reference's `spectrum` (simulation/analysis.py:44-96) run on the reference solver's own frames
(tests/golden/spectrum_default_json_128.npz).  Tolerance: probe lines bit-exact (fp64 exact
arithmetic), spectra <= 1e-12 relative L2 (direct DFT vs pocketfft)."""
import os

import numpy as np
import pytest

from tests import helpers as H
from tests.test_gpu_plugin import fake_from_golden, make_solver

pytestmark = pytest.mark.gpu
GOLD = os.path.join(H.GOLDEN_DIR, "spectrum_default_json_128.npz")


def test_device_spectrum_matches_reference(tmp_path):
    gold = np.load(GOLD)
    d = H.load_golden("default_json_1000")
    N, y, xi = int(gold["steps"]), int(gold["y_index"]), int(gold["x_index"])
    s = make_solver(d, tmp_path, record="off", write_mode="off",
                    probes=[{"u": "ux", "y": y, "z": 0}, {"u": "uz", "y": y, "z": 0}])
    g, m = fake_from_golden(d)
    s.init(g, m, N)
    s.run()
    for u_id in ("ux", "uz"):
        line = s.engine.probe_read(s._probes[(u_id, y, 0)])
        assert np.array_equal(line, gold["line_" + u_id]), u_id          # frame tt = state after step tt (App. B #10)
        x, f, d2 = s.spectrum(u_id, 0, y)
        _, _, d1 = s.spectrum(u_id, 0, y, x_index=xi)
        assert np.array_equal(x, gold["x_" + u_id]) and np.array_equal(f, gold["f"])
        assert d2.shape == gold["dft_%s_2d" % u_id].shape and d1.shape == gold["dft_%s_1d" % u_id].shape
        assert H.rel_l2([d2], [gold["dft_%s_2d" % u_id]]) <= 1e-12, u_id
        assert H.rel_l2([d1], [gold["dft_%s_1d" % u_id]]) <= 1e-12, u_id
        _, _, dm = s.spectrum(u_id, 0, y, x_index=-1)                    # negative index like u[x_index] in NumPy
        assert H.rel_l2([dm], [np.abs(np.fft.fft(line[-1] * np.hanning(N), norm="ortho"))[:N // 2]]) <= 1e-12
    with pytest.raises(KeyError):
        s.spectrum("uy", 0, y)


@pytest.mark.parametrize("dtype,record_every,steps", [("f64", 1, 250), ("f32", 3, 333)])
def test_device_spectrum_at_size(dtype, record_every, steps):
    """Odd lengths (250 / 111 frames, 95 / 96 rows), fp32 state, decimated sampling: the device transform
    of the probe's own samples against NumPy's FFT of the same samples."""
    from phonomena_b200.workloads import crystal_case
    case = crystal_case(96, 64, 32)
    e = case.make_engine(dtype=dtype, arith="fast", steps=steps, record_every=record_every)
    pu, pz = e.probe_add("ux", 20, 0, steps // record_every), e.probe_add(2, 33, 5, steps // record_every)
    e.run(steps)
    e.sync()
    for pid, rows in ((pu, 95), (pz, 96)):
        line = e.probe_read(pid)
        N = steps // record_every
        assert line.shape == (rows, N) and np.abs(line).max() > 0
        win = np.hanning(N)
        got = np.abs(e.probe_dft_xt(pid, win, N // 2, rows)) / np.sqrt(rows * N)
        ref = np.abs(np.fft.fft2(line * win, norm="ortho"))[:, :N // 2]
        assert H.rel_l2([got], [ref]) <= 1e-12
        a = e.probe_dft_t(pid, win, N // 2, 7, 3)
        assert H.rel_l2([np.abs(a)], [np.abs(np.fft.fft(line[7:10] * win, axis=1))[:, :N // 2]]) <= 1e-12
    with pytest.raises(Exception, match="full"):
        e.set_source_table(np.zeros(record_every))
        e.run(record_every)
    e.close()


def test_probe_argument_errors():
    from phonomena_b200.workloads import crystal_case
    e = crystal_case(32, 32, 16).make_engine(dtype="f64", arith="fast", steps=4)
    with pytest.raises(Exception, match="outside"):
        e.probe_add("uy", 31, 0, 4)           # uy has Ny-1 rows in y
    with pytest.raises(Exception, match="outside"):
        e.probe_add("uz", 0, 15, 4)           # uz has Nz-1 planes
    with pytest.raises(Exception, match="component"):
        e.probe_add(3, 0, 0, 4)
    with pytest.raises(Exception, match="no probe"):
        e.probe_shape(5)
    pid = e.probe_add("uz", 3, 0, 4)
    with pytest.raises(Exception):            # nothing sampled yet
        e.probe_dft_t(pid, np.zeros(0), 0)
    e.close()
