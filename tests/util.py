"""Shared helpers of the parity tests."""
import numpy as np


def image_metrics(a, b):
    """Frame relative RMSE and per-pixel agreement: a pixel agrees iff for all RGB channels
    |a - b| <= 1e-3 * max(|a|, |b|) + 1e-5 on the fp32 accumulation buffer."""
    a = np.asarray(a, dtype=np.float64)[..., :3]
    b = np.asarray(b, dtype=np.float64)[..., :3]
    rel_rmse = np.sqrt(np.sum((a - b) ** 2)) / max(np.sqrt(np.sum(b ** 2)), 1e-30)
    ok = np.all(np.abs(a - b) <= 1e-3 * np.maximum(np.abs(a), np.abs(b)) + 1e-5, axis=-1)
    return float(rel_rmse), float(ok.mean())


def setup_device(dev, scene, rgba, rg, var, width, height, records, stripes=(8, 0, 1)):
    from risltc_b200.scenes import material_constants
    dev.upload_mesh(scene["mesh"])
    if scene.get("textures") is not None:
        dev.upload_textures(scene["textures"])
    else:
        dev.upload_materials(material_constants(scene["materials"]))
    dev.upload_lights(records)
    dev.upload_ltc(rgba, rg)
    dev.set_variant(var)
    dev.resize(width, height, *stripes)


def constants_bytes(constants_list):
    import ctypes
    return b"".join(bytes(ctypes.string_at(ctypes.byref(c), 256)) for c in constants_list)


def parity_log(line):
    """Print a parity figure and append it to gpurun_out/parity_log.txt (copied to profiles/ after a GPU session, so that
    the numbers behind the gates are committed, not only their pass / fail)."""
    import os
    from pathlib import Path
    if os.environ.get("RISLTC_WINNER"):
        line = f"[RISLTC_WINNER={os.environ['RISLTC_WINNER']}] " + line
    print(line)
    out = Path(__file__).resolve().parent.parent / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        with open(out / "parity_log.txt", "a") as f:
            f.write(line + "\n")
    except OSError:
        pass
