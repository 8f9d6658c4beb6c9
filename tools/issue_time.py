import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import orc
from risltc_b200 import api, ltc_fit, scenes
from tests.util import constants_bytes, setup_device
W, H, frames = 1920, 1080, 64
fits = ltc_fit.fit_ggx_ltc(16, 6, 16); rgba, rg = ltc_fit.quantize_fits(fits)
scene = scenes.many_light_room(64, 200, seed=2, width=W, height=H)
osc = orc.OracleScene(scene, rgba, rg)
cs = constants_bytes([orc.make_constants(scene, W, H, orc.frame_words(f)[0], ltc_res=rgba.shape[1], ltc_layers=rgba.shape[0]) for f in range(frames)])
dev = api.Device(0)
for stripes in ((8, 0, 1), (8, 0, 8)):
    setup_device(dev, scene, rgba, rg, api.variant(), W, H, osc.records, stripes)
    for _ in range(2):
        dev.render_frames(cs); dev.synchronize()
    t0 = time.perf_counter(); dev.render_frames(cs); t1 = time.perf_counter(); dev.synchronize(); t2 = time.perf_counter()
    print(f"stripes {stripes}: host issues {frames} frames in {1e3 * (t1 - t0):.2f} ms ({1e6 * (t1 - t0) / frames:.1f} us per frame), device finishes after {1e3 * (t2 - t0):.2f} ms")
dev.close()
