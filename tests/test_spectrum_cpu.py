"""Spectrum post-processing against the reference's own `spectrum`
(simulation/analysis.py:44-96), whose answers on the reference solver's frames are stored in
tests/golden/spectrum_default_json_128.npz (oracle/gen_golden.py case_spectrum).  CPU only."""
import os

import numpy as np
import pytest

from oracle import spectrum_numpy as osp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "spectrum_default_json_128.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("u_id", ["ux", "uz"])
def test_oracle_restatement_matches_reference(gold, u_id):
    line, dt, xi = gold["line_" + u_id], float(gold["dt"]), int(gold["x_index"])
    f, d2 = osp.spectrum_line(line, dt)
    _, d1 = osp.spectrum_line(line, dt, x_index=xi)
    assert np.array_equal(f, gold["f"])
    assert np.array_equal(d2, gold["dft_%s_2d" % u_id]) and np.array_equal(d1, gold["dft_%s_1d" % u_id])
    x = osp.nonlinspace(gold["fdx"]) if u_id == "ux" else gold["x"]
    assert np.array_equal(x, gold["x_" + u_id])


@pytest.mark.parametrize("u_id", ["ux", "uz"])
def test_file_based_spectrum_matches_reference(gold, u_id, tmp_path):
    """phonomena_b200.analysis.spectrum reading an h5lite file that holds the reference's frames."""
    from phonomena_b200 import analysis
    from phonomena_b200.h5lite import H5Writer
    line = gold["line_" + u_id]
    X, N = line.shape
    p = str(tmp_path / "g.h5")
    with H5Writer(p) as w:
        w.attrs.update({"x": gold["x"], "fdx": gold["fdx"].reshape(-1, 1, 1), "dt": float(gold["dt"]), "steps": N})
        d = w.create_chunked(u_id, (X, 1, 1, N))
        for t in range(N):
            w.write_frame(d, t, line[:, t].reshape(X, 1, 1))
    x, f, d2 = analysis.spectrum(p, u_id, 0, 0)
    _, _, d1 = analysis.spectrum(p, u_id, 0, 0, x_index=int(gold["x_index"]))
    assert np.array_equal(x, gold["x_" + u_id]) and np.array_equal(f, gold["f"])
    assert np.array_equal(d2, gold["dft_%s_2d" % u_id]) and np.array_equal(d1, gold["dft_%s_1d" % u_id])


def test_hostmath_nonlinspace(gold):
    from phonomena_b200 import hostmath as hm
    assert np.array_equal(hm.nonlinspace(gold["fdx"].reshape(-1, 1, 1)), gold["x_ux"])
