import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from risltc_b200 import api
fn = np.load(Path(__file__).resolve().parent.parent / "tests/golden/functions.npz")
dev = api.Device(0)
for name, P in (("ris_ltc_v3", 4), ("uni_psa_v7", 8)):
    keep = fn[f"{name}.clipped_counts"] > 0
    polys, counts, rnd = fn[f"{name}.clipped"][keep], fn[f"{name}.clipped_counts"][keep], fn[f"{name}.randoms"][keep]
    want = fn[f"{name}.psa_polygon"][keep]
    got, _ = dev.kat_psa(polys, counts, rnd, P, 0, 0)
    gotf, _ = dev.kat_psa(polys, counts, rnd, P, 1, 0)
    for i in range(len(counts)):
        vc = int(counts[i]); n = vc if want[i, 33] > 0 else vc - 1
        d = got[i, 35:35 + n].view(np.uint32) != want[i, 35:35 + n].view(np.uint32)
        if d.any():
            eqf = got[i, 35:35 + n].view(np.uint32) == gotf[i, 35:35 + n].view(np.uint32)
            print(name, "polygon", i, "block", i // 128, "lane", i % 32, "warp", (i % 128) // 32, "vc", vc, "central" if want[i, 33] > 0 else "decentral", "bad sectors", np.nonzero(d)[0].tolist(), "equal to fast:", eqf.tolist())
dev.close()
