"""State export for the reference's consumers (export.py): the pbrt mesh block is formatted like PbrtRenderer.render does it
(pbrt_renderer.py:205-262), and frames recorded while the rollout runs equal the frames read back directly."""
import numpy as np
import pytest

from meshless_inflatable_softbody_b200 import export


def _reference_format(shape_properties):
    """pbrt_renderer.py:205-226,258-262 restated: the text PbrtRenderer.render writes for one shape's properties."""
    def convert(value):
        if isinstance(value, (float, int)):
            return str(value)
        value = list(value)
        is_float = any(type(v) in (float, np.float64, np.float32) for v in value)
        arr = np.asarray(value, dtype=np.float64 if is_float else np.int32).ravel()
        return "[" + " ".join(str(v) for v in arr) + "]"
    out = "   Shape \"trianglemesh\"\n"
    for k, v in shape_properties.items():
        out += "       \"{}\" {}\n".format(k, convert(v))
    return out


def test_trianglemesh_block_matches_the_renderer_format():
    rng = np.random.default_rng(0)
    v = rng.normal(size=(7, 3)).astype(np.float32)
    f = np.array([[0, 1, 2], [2, 3, 1], [4, 5, 6]])
    uv = rng.uniform(size=(7, 2))
    want = _reference_format({"integer indices": f.astype(np.int32).ravel(), "point3 P": v.astype(np.float64).ravel(),
                              "point2 uv": uv.ravel(), "float alpha": 1.0})
    assert export.trianglemesh_block(v, f, uv) == want
    want = _reference_format({"integer indices": f.astype(np.int32).ravel(), "point3 P": v.astype(np.float64).ravel(), "float alpha": 0.5})
    assert export.trianglemesh_block(v, f, None, alpha=0.5) == want


@pytest.mark.gpu
def test_recorded_frames_equal_direct_readback():
    from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes
    x0, out_num = scenes.jittered_sphere(1500, seed=0, low_drop=True)
    a = Simulator(x0, SceneConfig())
    a.startup()
    xs, vs = export.record_frames(a, frames=130, every=50, out_num=out_num, with_velocity=True)
    assert xs.shape == (3, out_num, 3) and a.frame == 130
    b = Simulator(x0, SceneConfig())
    b.startup()
    for k, f in enumerate((0, 50, 100)):
        b.step(f - b.frame)
        x, v = b.position_velocity()
        assert np.array_equal(xs[k], x.cpu().numpy()[:out_num])          # same kernels, same order: bit-identical
        assert np.array_equal(vs[k], v.cpu().numpy()[:out_num])
    assert np.array_equal(xs[0], x0[:out_num])                          # frame 0 = the reference configuration (sim.py:261-266)


@pytest.mark.gpu
def test_rollout_loss_matches_the_reference_formula(tmp_path):
    """compute_loss (sim.py:269-273) over the target frames, forward value: targets written by export_targets of one scene, loss of a
    second scene with a different design field against them, checked against the same sum formed in fp64 from the exported states;
    and a scene against its own targets has zero loss."""
    import torch
    from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes
    x0, out_num = scenes.jittered_sphere(1200, seed=1, low_drop=True)
    cfg = SceneConfig()
    a = Simulator(x0, cfg)
    a.startup()
    a.export_targets(str(tmp_path), every=20, count=3)                 # frames 20, 40, 60
    same = Simulator(x0, cfg)
    assert same.rollout_loss(str(tmp_path), frames=60) == 0.0
    b = Simulator(x0, cfg)
    xd = np.full(len(x0), -1.0, np.float32); xd[:out_num] = 0.5        # softer shell
    b.set_design(xd)
    got = b.rollout_loss(str(tmp_path), frames=60)
    c = Simulator(x0, cfg)
    c.set_design(xd); c.startup()
    want = 0.0
    for i in (1, 2, 3):
        c.step(20)
        x, v = (t.cpu().numpy().astype(np.float64) for t in c.position_velocity())
        tx, tv = np.load(tmp_path / f"position_{i}.npy").astype(np.float64), np.load(tmp_path / f"velocity_{i}.npy").astype(np.float64)
        want += ((x - tx) ** 2).sum() + float(np.float32(cfg.time_step)) * ((v - tv) ** 2).sum()
    assert got > 0.0 and abs(got - want) <= 1e-5 * want
    assert b.rollout_loss(str(tmp_path), frames=60) == got             # fixed summation order: reproducible to the bit
