"""-m gpu: the offline drop-in executed as the reference is: `risltc -run_exp` with the EXP_* environment of the timing
experiment (experiment_list.c:354-396: five estimators, 1000 accumulated 1920x1080 frames each, one timings.txt and one
screenshot per experiment; main.c:2719-2790, frame_timer.c:49-51) on a generated scene stored under the reference's file
names. Checks the artefacts: timings.txt rows `accum_num,ms`, and the *.hdr of the default estimator decodes to the
frame the C ABI accumulates for the same settings."""
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PKG = Path(__file__).resolve().parent.parent / "risltc_b200"


def test_run_exp_timing_experiment(tmp_path):
    from risltc_b200 import api, formats, host, ltc_fit, scenes
    W, H, samples = 1920, 1080, 1000
    scene = scenes.many_light_room(64, 200, seed=2, width=W, height=H)
    fits = ltc_fit.fit_ggx_ltc(64, 51, 64)
    vks, tex, save = host.write_scene_files(scene, tmp_path, ltc_fits=fits, name="gen")
    # the experiment table names the reference's scenes: store the generated one under those names (SCENE=bistro_inside)
    shutil.move(vks, tmp_path / "Bistro_interior.vks")
    shutil.move(tex, tmp_path / "Bistro_textures")
    shutil.move(save, tmp_path / "quicksaves" / "Bistro_interior.save")
    env = {**os.environ, "RISLTC_DATA_DIR": str(tmp_path), "EXP_LO_ROUGH": "1", "EXP_TIMINGS": "1", "SCENE": "bistro_inside"}
    p = subprocess.run([str(PKG / "risltc"), "-run_exp"], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert "Defined 5 experiments to reproduce." in p.stdout and "All experiments finished" in p.stdout
    root = tmp_path / "experiments" / "lo_rough" / "bistro_inside"
    names = ["uniform_uniform_time", "uniform_cp_time", "uniform_area_time", "cp_cp_time", "ltc_cp_time"]
    mean_ms = {}
    for name in names:
        rows = [l.split(",") for l in (root / name / "timings.txt").read_text().splitlines()]
        # one row per rendered frame: accum_num counts up from 0, the frame that captures the screenshot is timed too
        assert len(rows) >= samples and [int(r[0]) for r in rows[:samples]] == list(range(samples)), name
        ms = np.array([float(r[1]) for r in rows])
        assert np.all(ms > 0.0) and np.all(ms < 1000.0), name
        mean_ms[name] = float(ms[10:].mean())
        shots = sorted((root / name).glob("*.hdr"))
        assert [s.name for s in shots] == [f"{samples:05d}.hdr"], (name, shots)
    print("mean frame times (ms):", {k: round(v, 3) for k, v in mean_ms.items()})
    # the screenshot of the default estimator ("ours": light_reservoir + ltc_cp) against the C ABI's accumulation buffer
    got = formats.read_hdr(root / "ltc_cp_time" / f"{samples:05d}.hdr")
    assert got.shape == (H, W, 3)
    shutil.move(tmp_path / "Bistro_interior.vks", tmp_path / "gen.vks")
    shutil.move(tmp_path / "Bistro_textures", tmp_path / "gen_textures")
    shutil.move(tmp_path / "quicksaves" / "Bistro_interior.save", tmp_path / "quicksaves" / "gen.save")
    app = host.Application(tmp_path)
    try:
        app.load(str(tmp_path / "gen.vks"), str(tmp_path / "gen_textures"), str(tmp_path / "quicksaves" / "gen.save"), W, H)
        app.settings(accum=1, roughness_factor=0.05, exposure_factor=1.5, light_sampling=api.LIGHT["reservoir"], polygon_sampling_technique=api.POLY["ltc_cp"])
        # the screenshot is read after the frame with accum_num == 1000 has been rendered: 1001 accumulated frames per experiment,
        # and the noise seed keeps counting across experiments (set_noise_constants, noise_table.c:24-28): this is the fifth
        app.reset(4 * (samples + 1))
        app.render_frames(samples + 1)
        want = app.device().read_accum()[..., :3]
    finally:
        app.close()
    want16 = want.astype(np.float16).astype(np.float32)      # the capture passes through fp16 (copy_pass.frag.glsl:37-53)
    # RGBE keeps 8 bits of the largest channel: |error| <= 2^-8 of it, plus one truncation step
    tolerance = np.max(want16, axis=-1, keepdims=True) * (2.0 ** -7) + 1e-6
    assert np.all(np.abs(got - want16) <= tolerance), float(np.max(np.abs(got - want16) / tolerance))
