"""-m gpu: whole-frame parity of the CUDA path (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

from tests.util import constants_bytes, image_metrics, setup_device

pytestmark = pytest.mark.gpu


def _render_both(device, scene, ltc_tables, ovar, gvar, width, height, frames, **ckw):
    from oracle import orc
    _, rgba, rg = ltc_tables
    osc = orc.OracleScene(scene, rgba, rg)
    cs = [orc.make_constants(scene, width, height, orc.frame_words(f)[0], ltc_res=rgba.shape[1], ltc_layers=rgba.shape[0], **ckw) for f in range(frames)]
    ref, ref_vis, ref_rays = osc.render(cs, ovar)
    setup_device(device, scene, rgba, rg, gvar, width, height, osc.records)
    device.render_frames(constants_bytes(cs))
    got = device.read_accum()
    vis = device.read_visibility()
    return ref, ref_vis, ref_rays, got, vis, device.counters()


@pytest.mark.parametrize("light_sampling,technique", [("uniform", "projected_solid_angle"), ("reservoir", "ltc_cp")])
def test_quad_over_plane(device, ltc_tables, light_sampling, technique):
    """BASELINE.json configs[0] at reduced size: single quad light over a diffuse plane."""
    from oracle import orc
    from risltc_b200 import api, scenes
    scene = scenes.quad_over_plane(320, 180)
    kw = dict(light_sampling=light_sampling, technique=technique, min_vertices=4, max_vertices=4)
    ref, ref_vis, ref_rays, got, vis, counters = _render_both(device, scene, ltc_tables, orc.variant(**kw), api.variant(**kw), 320, 180, 1)
    assert np.array_equal(vis, ref_vis), "primary visibility must be bit-exact"
    rmse, agree = image_metrics(got, ref)
    print(f"quad {light_sampling}/{technique}: rel_rmse={rmse:.3e} agree={agree:.5f} rays gpu={counters['shadow_rays']} cpu={ref_rays}")
    assert rmse <= 1e-3 and agree >= 0.99


@pytest.mark.parametrize("frames", [1, 4])
def test_room_default_variant(device, ltc_tables, frames):
    """The default estimator (RIS over LTC integrals -> PSA + LTC MIS) on a 64-light room."""
    from oracle import orc
    from risltc_b200 import api, scenes
    scene = scenes.many_light_room(64, 50, width=320, height=180)
    ref, ref_vis, ref_rays, got, vis, counters = _render_both(device, scene, ltc_tables, orc.variant(), api.variant(), 320, 180, frames)
    assert np.array_equal(vis, ref_vis), "primary visibility must be bit-exact"
    rmse, agree = image_metrics(got, ref)
    print(f"room frames={frames}: rel_rmse={rmse:.3e} agree={agree:.5f} rays gpu={counters['shadow_rays']} cpu={ref_rays}")
    assert agree >= 0.99
    assert rmse <= (1e-3 if frames > 1 else 5e-2)
