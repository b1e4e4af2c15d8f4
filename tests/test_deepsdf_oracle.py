"""The DeepSDF oracle against golden vectors produced by the reference's own deepsdf.py
(tests/golden/make_deepsdf_golden.py) -- the one boundary of this path the reference pins."""
import os

import numpy as np
import pytest

from oracle import deepsdf_oracle as do

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deepsdf_seed0.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def state():
    return do.seeded_state(0)


def test_seeded_state_is_the_reference_state(gold, state):
    """Same keys, shapes and checksums as `torch.manual_seed(0); DeepSDFWithCode().state_dict()` of deepsdf.py."""
    keys = [str(k) for k in gold["keys"]]
    assert sorted(state) == keys
    for k, shp, s, a in zip(keys, gold["shapes"], gold["sums"], gold["abs_sums"]):
        assert str(tuple(state[k].shape)) == str(shp)
        assert state[k].astype(np.float64).sum() == pytest.approx(float(s), rel=1e-12, abs=1e-12)
        assert np.abs(state[k].astype(np.float64)).sum() == pytest.approx(float(a), rel=1e-12)
    assert np.array_equal(state["network.3.parametrizations.weight.original1"][0], gold["first_row"])


def test_forward_matches_reference_outputs(gold, state):
    got32 = do.forward(state, gold["points"], np.float32)
    got64 = do.forward(state, gold["points"], np.float64)
    assert got32.shape == gold["sdf"].shape == (256, 1)
    assert np.abs(got64 - gold["sdf64"]).max() < 1e-12                    # fp64 vs the reference in fp64
    scale = np.abs(gold["sdf64"]).max()
    assert np.abs(got32 - gold["sdf"]).max() < 2e-6 * scale + 1e-7        # fp32 vs the reference in fp32 (different GEMM blocking)
    assert np.abs(gold["sdf"] - gold["sdf64"]).max() < 2e-6 * scale + 1e-7


def test_state_dict_layout(state):
    """deepsdf.py:12-38: 9 weight-normalised Linear layers 3 -> 1024 x 8 -> 1 at Sequential indices 0,3,...,24."""
    lay = do.layers_from_state(state)
    assert len(lay) == 9
    assert lay[0][1].shape == (1024, 3) and lay[-1][1].shape == (1, 1024)
    assert all(v.shape == (1024, 1024) for _, v, _ in lay[1:-1])
    assert all(g.shape == (v.shape[0], 1) for g, v, _ in lay)
    # default weight_norm init: g = ||v|| so W == v
    W, _ = do.effective_weights(state)[3]
    assert np.abs(W - lay[3][1]).max() < 1e-6


def test_octahedron_weights_are_an_exact_sdf():
    st = do.octahedron_state(0.03, hidden=256, n_linear=4)
    rng = np.random.default_rng(1)
    p = rng.uniform(-0.08, 0.08, size=(500, 3))
    want = (np.abs(p).sum(1) - 0.03) / np.sqrt(3.0)
    got = do.forward(st, p, np.float64)[:, 0]
    assert np.abs(got - want).max() < 1e-8


def test_contact_plane_reproduces_ground_penalty():
    """sdf = y (a plane) must give the reference's ground penalty f_y = (range - y)^2 k (sim.py:241-243)."""
    hidden, n_linear = 256, 3
    st = do.octahedron_state(0.0, hidden=hidden, n_linear=n_linear)
    # rewire: layer 0 rows (+e_y, -e_y); last layer = u0 - u1 = y
    g0 = np.zeros((hidden, 1), np.float32); g0[:2] = 1.0
    v0 = np.zeros((hidden, 3), np.float32); v0[0, 1] = 1.0; v0[1, 1] = -1.0; v0[2:, 0] = 1.0
    st["network.0.parametrizations.weight.original0"] = g0
    st["network.0.parametrizations.weight.original1"] = v0
    vl = np.zeros((1, hidden), np.float32); vl[0, 0] = 1.0; vl[0, 1] = -1.0
    st[f"network.{3 * (n_linear - 1)}.parametrizations.weight.original1"] = vl
    st[f"network.{3 * (n_linear - 1)}.parametrizations.weight.original0"] = np.full((1, 1), np.sqrt(2.0), np.float32)
    st[f"network.{3 * (n_linear - 1)}.bias"] = np.zeros(1, np.float32)
    p = np.array([[0.01, 5e-5, 0.0], [0.0, 2e-4, 0.01], [0.0, -3e-5, 0.0]])
    s0, g, f = do.contact_force(st, p, np.eye(3), np.zeros(3), 3e5, 1e-4, 1e-3)
    assert np.allclose(s0, p[:, 1], atol=1e-12)
    want = np.where(p[:, 1] < 1e-4, (1e-4 - p[:, 1]) ** 2 * 3e5, 0.0)
    assert np.allclose(f[:, 1], want, rtol=1e-9) and np.abs(f[:, [0, 2]]).max() < 1e-12
