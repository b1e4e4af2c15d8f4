"""Error behaviour of the C-ABI (include/mis.h: negative MIS_E_* codes + mis_last_error, no silent fallback) and edge-case inputs:
coincident particles, a scene wider than the Morton key range, calls out of order."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import make_oracle
from meshless_inflatable_softbody_b200 import SceneConfig, scenes, native

pytestmark = pytest.mark.gpu


def _sim(x0, **kw):
    from meshless_inflatable_softbody_b200 import Simulator
    return Simulator(np.asarray(x0, np.float32), SceneConfig(), **kw)


def test_call_order_violations_are_state_errors():
    x0, _ = scenes.jittered_sphere(300, seed=0, low_drop=True)
    sim = _sim(x0, apply_defaults=False)
    with pytest.raises(native.MisError, match=r"\(-3\)"):          # MIS_E_STATE: step before startup
        sim.step(1)
    sim.startup()
    with pytest.raises(native.MisError, match=r"\(-3\).*set_mass"):  # startup done, but mass / material never set
        sim.step(1)
    sim.set_all_external_force([0.0, 0.0, 0.0]); sim.set_youngs_modulus(1.5e5); sim.set_poisson_ratio(0.4); sim.set_mass(1e-4)
    sim.set_design(-1.0)
    sim.step(3)
    assert bool(torch.isfinite(sim.position()).all())


def test_bad_arguments_are_rejected_with_a_message():
    L = native.lib()
    x = torch.zeros((4, 3), device="cuda")
    p = native.MisParams(); p.h, p.dt = 0.007, 5e-5
    h = C.c_void_p()
    assert L.mis_create(0, x.data_ptr(), C.byref(p), None, C.byref(h)) == -1 and b"bad argument" in L.mis_last_error()
    p.h = 0.0
    assert L.mis_create(4, x.data_ptr(), C.byref(p), None, C.byref(h)) == -1 and b"positive" in L.mis_last_error()
    p.h, p.lanes_per_particle = 0.007, 5
    assert L.mis_create(4, x.data_ptr(), C.byref(p), None, C.byref(h)) == -1 and b"lanes_per_particle" in L.mis_last_error()
    p.lanes_per_particle, p.cluster_size = 0, 3
    assert L.mis_create(4, x.data_ptr(), C.byref(p), None, C.byref(h)) == -1 and b"cluster_size" in L.mis_last_error()
    assert L.mis_step(None, 1, None) == -1 and L.mis_destroy(None) == 0


def test_scene_wider_than_1024_cells_bins_with_per_axis_key_bits():
    """Two bodies 1 500 cells (21 m) apart: more than the 1 024 cells per axis a 3 x 10-bit Morton key holds.  The key takes a bit
    count per axis instead (11 + 4 + 4 here); lists and trajectory match the oracle as for any other scene."""
    a, _ = scenes.jittered_sphere(400, seed=6, low_drop=True)
    b = a.copy(); b[:, 0] += 1500 * 0.014
    x0 = np.concatenate([a, b], 0).astype(np.float32)
    sim, o, r = _sim(x0), make_oracle(x0), make_oracle(x0)
    off, nb = (t.cpu().numpy() for t in sim.neighbors())
    cnt, ooff, oflat = o.neighbor_lists()
    assert np.array_equal(np.diff(off), cnt)
    rows = np.repeat(np.arange(len(x0)), cnt)
    assert np.array_equal(nb[np.lexsort((nb, rows))], oflat)
    r.set_order(1)
    sim.startup(); o.startup(); r.startup()
    sim.step(50); o.step(50); r.step(50)
    x, v = sim.position_velocity()
    assert np.abs(x.cpu().numpy() - o.position()).max() <= 4 * np.abs(o.position() - r.position()).max() + 4e-9
    assert np.abs(v.cpu().numpy() - o.velocity()).max() <= 4 * np.abs(o.velocity() - r.velocity()).max() + 2e-5


def test_bounding_box_beyond_the_dense_cell_table_is_unsupported_not_wrong():
    # 701^3 cells of width 2h > 2^28: the dense cell table is the limit
    x0 = np.array([[0.0, 0.07, 0.0], [700 * 0.014, 0.07 + 700 * 0.014, 700 * 0.014]], np.float32)
    with pytest.raises(native.MisError, match=r"\(-4\).*2\^28"):      # MIS_E_UNSUPPORTED
        _sim(x0)


def test_coincident_particles_match_the_oracle():
    """Two particles at the same reference position are neighbours at distance 0 (q = 0 < 2, sim.py:137-141): W = sigma, the
    gradient term vanishes with x0_ij = 0.  No NaN, same neighbour lists and trajectory as the oracle."""
    x0, _ = scenes.jittered_sphere(400, seed=4, low_drop=True)
    x0 = np.concatenate([x0, x0[:5]], 0).astype(np.float32)            # five exact duplicates
    sim, o = _sim(x0), make_oracle(x0)
    off, nb = (t.cpu().numpy() for t in sim.neighbors())
    cnt, ooff, oflat = o.neighbor_lists()
    assert np.array_equal(np.diff(off), cnt)
    for i in list(range(5)) + list(range(len(x0) - 5, len(x0))):
        assert np.array_equal(np.sort(nb[off[i]:off[i + 1]]), oflat[ooff[i]:ooff[i + 1]])
    assert len(x0) - 5 in nb[off[0]:off[1]]                            # the duplicate of particle 0 is its neighbour
    sim.startup(); o.startup()
    sim.step(10); o.step(10)
    x, v = sim.position_velocity()
    assert bool(torch.isfinite(x).all()) and bool(torch.isfinite(v).all())
    assert np.abs(x.cpu().numpy() - o.position()).max() < 2e-7 and np.abs(v.cpu().numpy() - o.velocity()).max() < 2e-3
