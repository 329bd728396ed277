"""Pin the oracle: the NumPy restatement must reproduce, BIT FOR BIT, what the
unmodified reference solver produced (tests/golden/*.npz, made by oracle/gen_golden.py),
and the known answers recorded in SURVEY.md App. D."""
import hashlib

import numpy as np
import pytest

from oracle import fdtd_numpy as onp
from tests import helpers as H


@pytest.mark.parametrize("name", H.golden_names())
def test_material_indexing_bit_exact(name):
    d = H.load_golden(name)
    ids = onp.material_id_map(d["x"], d["y"], d["z"], H.targets_of(d))
    assert ids.dtype == np.uint8 and np.array_equal(ids, d["ids"])


@pytest.mark.parametrize("name", H.golden_names())
def test_dt_bit_exact(name):
    d = H.load_golden(name)
    o = H.oracle_from_golden(d)
    assert o.dt == d["dt"]


@pytest.mark.parametrize("name", H.golden_names())
def test_fields_bit_exact(name):
    d = H.load_golden(name)
    o = H.oracle_from_golden(d)
    snaps = set(int(s) for s in d["snap_steps"])
    for n in range(1, d["steps"] + 1):
        o.step()
        if n in snaps:
            assert np.array_equal(o.ux[:, :, 0], d["snap_ux_%d" % n]), (name, n)
            assert np.array_equal(o.uy[:, :, 0], d["snap_uy_%d" % n]), (name, n)
            assert np.array_equal(o.uz[:, :, 0], d["snap_uz_%d" % n]), (name, n)
    for k in ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old", "T1", "T2", "T3", "T4", "T5", "T6"):
        assert np.array_equal(getattr(o, k), d[k]), (name, k)


def test_threaded_oracle_identical():
    d = H.load_golden("crystal_48x32x12")
    o = H.oracle_from_golden(d, threads=6).run(d["steps"])
    for k in ("ux", "uy", "uz"):
        assert np.array_equal(getattr(o, k), d[k])
    o.close()


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_survey_known_answers():
    """SURVEY.md App. D, captured from the live reference during the survey."""
    d = H.load_golden("testdefaults")
    assert d["dt"] == 1.0567842397069181e-09
    assert (_sha(d["ux"]), _sha(d["uy"]), _sha(d["uz"])) == ("5d571b6356ea995d", "1ad9aef8ad57196d", "ed80eec5a25f851a")
    d = H.load_golden("default_json_1000")
    assert d["dt"] == 2.1135684794138367e-05
    assert int(d["ids"].sum()) == 324
    assert (_sha(d["ux"]), _sha(d["uy"]), _sha(d["uz"])) == ("550fc2c9fa79dd8c", "e2b338b9acafdee2", "a4a7f181bcd15c68")
    o = H.oracle_from_golden(d).run(1000)
    assert (_sha(o.ux), _sha(o.uy), _sha(o.uz)) == ("550fc2c9fa79dd8c", "e2b338b9acafdee2", "a4a7f181bcd15c68")
    assert np.linalg.norm(H.load_golden("nonuniform_json_200")["uz"]) == 2.849693082748335
    assert np.linalg.norm(H.load_golden("fine_json_200")["uz"]) == 3.378846410806348


@pytest.mark.ref
def test_oracle_vs_live_reference_random_fields():
    """Run the live reference for a few steps from random fields on a mesh that is
    non-uniform in x, y AND z; compare every array bit for bit."""
    from oracle import refshim
    d = H.load_golden("nonuniform_json_200")
    rng = np.random.default_rng(0)
    z = np.cumsum(np.concatenate([[0.0], rng.uniform(0.5, 1.5, d["z"].size - 1)]))
    s = refshim.default_solver()
    from simulation import grid as rgrid, material as rmat

    class G(rgrid.Grid):  # keep the hand-made mesh: the reference re-runs buildMesh in init
        def buildMesh(self, *a, **k):
            pass
    g = G()
    g.size_x, g.size_y, g.size_z = 30, 20, 5
    g.x, g.y, g.z = d["x"].copy(), d["y"].copy(), z
    for t in d["targets"]:
        g.targets = np.append(g.targets, np.array([tuple(t)], dtype=g.trgt_dtype))
    g.update()
    m = rmat.Material()
    props = {"a": {"name": "a", "c": (d["prim_c"] / 1e10).tolist(), "p": d["prim_p"]},
             "b": {"name": "b", "c": (d["sec_c"] / 1e10).tolist(), "p": d["sec_p"]}}
    m.init(grid=g, properties=props)
    m.c_max = 0.2
    m.setPrimary("a"); m.setSecondary("b"); m.update()
    s.cfg.update({"wave": "ricker", "wave_args": {"f": 3000.0, "source_delay": 1e-6}})
    s.init(g, m, 4)
    t = onp.make_targets(d["targets"].tolist())
    C, P = onp.set_constants(g.x, g.y, g.z, t, d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    assert np.array_equal(C, s.m.C) and np.array_equal(P, s.m.P)
    o = onp.OracleSolver(g.x, g.y, g.z, C, P, s.m.dt, wave="ricker", wave_args=s.cfg["wave_args"])
    for k in ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old"):
        a = rng.standard_normal(getattr(o, k).shape) * 1e-3
        getattr(o, k)[...] = a
        getattr(s.g, k)[...] = a
    for k in ("ux", "uy", "uz"):  # invariant of the reference: u_new == u at step start
        getattr(o, k + "_new")[...] = getattr(o, k)
        getattr(s.g, k + "_new")[...] = getattr(s.g, k)
    s.run()
    o.run(4)
    for k in ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old", "T1", "T2", "T3", "T4", "T5", "T6"):
        assert np.array_equal(getattr(o, k), getattr(s.g, k)), k


@pytest.mark.parametrize("omp", [False, True])
@pytest.mark.parametrize("name", H.golden_names())
def test_c_oracle_bit_exact(name, omp):
    """The C restatement (oracle/fdtd_c.c) against the reference's golden outputs, bit for bit."""
    from oracle import fdtd_c
    d = H.load_golden(name)
    o = fdtd_c.COracle(d["x"], d["y"], d["z"], d["ids"], [d["prim_c"], d["sec_c"]], [d["prim_p"], d["sec_p"]], d["dt"],
                       wave=d["wave"], wave_args=d["wave_args"], omp=omp)
    o.run(d["steps"])
    for k in ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old", "T1", "T2", "T3", "T4", "T5", "T6"):
        assert np.array_equal(getattr(o, k), d[k]), (name, k)
    o.close()


@pytest.mark.parametrize("name", H.periodic_names())
def test_periodic_y_oracle_matches_reference_stubs(name):
    """bc_y = periodic: the oracle's restatement of the reference's ARCHIVED apply_T_pbc / apply_u_pbc
    (base_solver.py:383-400, 475-486) against fixtures made by the unmodified reference solver with those stubs
    switched on through its own update_T_BC / update_u_BC hooks (oracle/gen_golden.periodic_y_solver) -- bit for bit,
    stresses included."""
    d = H.load_golden(name)
    assert str(d["bc_y"]) == "periodic"
    o = H.oracle_from_golden(d, bc_y="periodic")
    assert o.dt == d["dt"]
    snaps = set(int(s) for s in d["snap_steps"])
    for n in range(1, d["steps"] + 1):
        o.step()
        if n in snaps:
            assert np.array_equal(o.uz[:, :, 0], d["snap_uz_%d" % n]) and np.array_equal(o.ux[:, :, 0], d["snap_ux_%d" % n]), (name, n)
    for k in ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old", "T1", "T2", "T3", "T4", "T5", "T6"):
        assert np.array_equal(getattr(o, k), d[k]), (name, k)
    # the mode does something: the absorbing oracle on the same inputs ends elsewhere, and the periodic rows are copies
    a = H.oracle_from_golden(d).run(d["steps"])
    assert not np.array_equal(a.uz, o.uz)
    assert np.array_equal(o.ux[:, 0, :-1], o.ux[:, -2, :-1]) and np.array_equal(o.uy[:, -1, :-1], o.uy[:, 1, :-1])
    assert np.array_equal(o.uz[:, 0, :-1], o.uz[:, -2, :-1])


@pytest.mark.parametrize("name", H.periodic_names())
def test_bloch_oracle_phase_zero_is_pinned_by_the_reference_stubs(name):
    """BlochOracle defines the phase != 0 semantics (no reference behaviour exists for it); at phase 0 it must collapse to
    the reference-with-stubs fixtures bit for bit, with an identically zero imaginary part -- and for phase != 0 the
    complex field obeys the Bloch relation on the periodic rows."""
    d = H.load_golden(name)
    t = H.targets_of(d)
    C, P = onp.set_constants(d["x"], d["y"], d["z"], t, d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    o = onp.BlochOracle(d["x"], d["y"], d["z"], C, P, d["dt"], 0.0, wave=d["wave"], wave_args=d["wave_args"]).run(d["steps"])
    for k in ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old"):
        assert np.array_equal(getattr(o.re, k), d[k]), (name, k)
        assert not getattr(o.im, k).any()
    phase = 0.9
    o = onp.BlochOracle(d["x"], d["y"], d["z"], C, P, d["dt"], phase, wave=d["wave"], wave_args=d["wave_args"]).run(min(d["steps"], 60))
    z = o.re.ux + 1j * o.im.ux
    assert np.abs(o.im.uz).max() > 0 and np.allclose(z[:, 0, :-1], np.exp(-1j * phase) * z[:, -2, :-1], rtol=1e-12, atol=1e-300)
    zy = o.re.uy + 1j * o.im.uy
    assert np.allclose(zy[:, -1, :-1], np.exp(1j * phase) * zy[:, 1, :-1], rtol=1e-12, atol=1e-300)
