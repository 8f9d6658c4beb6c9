#!/usr/bin/env python3
"""Build oracle/_ref/ from the reference's own sources where they lie under /root/reference.
TEST INFRASTRUCTURE; outputs (shared objects only) go to oracle/_ref/, which is git-ignored.

  libref_host.so            polygonal_light.c, camera.c and math_utilities.h compiled as-is with gcc
                            (GLFW calls of control_camera are satisfied by ref_host_stubs.c)
  libref_shading_<v>.so     shading_pass.frag.glsl + includes compiled as C++ through glsl_shim.hpp,
                            one library per shader variant (the -D table of main.c:962-991)

The reference's renderer itself cannot be built here (no Vulkan headers / loader / glslangValidator,
SURVEY.md 8c); its build system is not run. The GLSL is made palatable to g++ by token rewrites that
do not change any expression: see transform(). The rewritten text only ever exists in a temporary
directory."""
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("RISLTC_REFERENCE", "/root/reference"))
OUT = HERE / "_ref"
CXX = "/usr/bin/g++"
CC = "/usr/bin/gcc"

# name -> (light_sampling, technique, mis, S, L, fast_atan, biased, V_min, V_max)
VARIANTS = {
    "ris_ltc_v3": dict(light="reservoir", tech="ltc_cp", mis="optimal_clamped", S=1, L=1, vmin=3, vmax=3),
    "ris_ltc_v4": dict(light="reservoir", tech="ltc_cp", mis="optimal_clamped", S=1, L=1, vmin=4, vmax=4),
    "uni_ltc_v3": dict(light="uniform", tech="ltc_cp", mis="optimal_clamped", S=1, L=1, vmin=3, vmax=3),
    "uni_psa_v4": dict(light="uniform", tech="psa", mis="optimal_clamped", S=1, L=1, vmin=4, vmax=4),
    "ris_psa_v3": dict(light="reservoir", tech="psa", mis="optimal_clamped", S=1, L=1, vmin=3, vmax=3),
    "ris_turk_v3": dict(light="reservoir", tech="turk", mis="optimal_clamped", S=1, L=1, vmin=3, vmax=3),
    "uni_turk_v3": dict(light="uniform", tech="turk", mis="optimal_clamped", S=1, L=1, vmin=3, vmax=3),
    "ris_psa_s2l2_v3": dict(light="reservoir", tech="psa", mis="balance", S=2, L=2, vmin=3, vmax=3),
    "ris_ltc_weighted_v3": dict(light="reservoir", tech="ltc_cp", mis="weighted", S=1, L=1, vmin=3, vmax=3),
    "ris_ltc_optimal_v3": dict(light="reservoir", tech="ltc_cp", mis="optimal", S=1, L=1, vmin=3, vmax=3),
    "uni_psa_biased_fast_v5": dict(light="uniform", tech="psa_biased", mis="power", S=1, L=1, fast_atan=1, vmin=3, vmax=5),
    # the largest polygons the reference supports (MAX_POLYGONAL_LIGHT_VERTEX_COUNT up to 7, main.c:191-204)
    "ris_psa_v6": dict(light="reservoir", tech="psa", mis="optimal_clamped", S=1, L=1, vmin=6, vmax=6),
    "uni_psa_v7": dict(light="uniform", tech="psa", mis="optimal_clamped", S=1, L=1, vmin=3, vmax=7),
    # control: the default variant compiled the way a GLSL compiler may compile it (a*b+c contracted to fma). Its
    # distance to ris_ltc_v3 is the reference's own sensitivity to legal rounding changes (tests/test_gpu_frames.py).
    "ris_ltc_v3_fma": dict(light="reservoir", tech="ltc_cp", mis="optimal_clamped", S=1, L=1, vmin=3, vmax=3, contract=True),
}

EXCLUDED_INCLUDES = {"cubic_solver.glsl", "srgb_utility.glsl"}   # unreachable from the shading pass
# of polygon_sampling_related_work.glsl only the Turk sampler and its density are reachable (:34-64)
RELATED_WORK_LAST_LINE = 65


def inline_includes(path, seen_depth=0):
    out = []
    lines = path.read_text().splitlines()
    if path.name == "polygon_sampling_related_work.glsl":
        lines = lines[:RELATED_WORK_LAST_LINE]
    for line in lines:
        m = re.match(r'\s*#include\s+"([^"]+)"', line)
        if m:
            if m.group(1) in EXCLUDED_INCLUDES:
                continue
            out.append(inline_includes(path.parent / m.group(1), seen_depth + 1))
        elif re.match(r"\s*#(version|extension)\b", line):
            continue
        else:
            out.append(line)
    return "\n".join(out)


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def transform(text):
    """Token-level rewrites from GLSL to C++; no expression is altered."""
    text = strip_comments(text)
    # control-flow attributes
    text = re.sub(r"\[\[\s*(unroll|dont_unroll)\s*\]\]", "", text)
    # float literals without suffix are floats in GLSL
    text = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])", r"\1f", text)
    # interface blocks -> namespaces (members become plain globals)
    def block(m):
        return f"namespace {m.group(1)}_ns {{"
    text, n_blocks = re.subn(r"layout\s*\([^)]*\)\s*(?:uniform|buffer)\s+(\w+)\s*\{", block, text)
    # per-invocation inputs / outputs
    text = re.sub(r"layout\s*\([^)]*\)\s*in\s+vec4\s+gl_FragCoord\s*;", "thread_local vec4 gl_FragCoord;", text)
    text = re.sub(r"layout\s*\([^)]*\)\s*out\s+(\w+)\s+(\w+)\s*;", r"thread_local \1 \2;", text)
    text = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+usubpassInput\s+(\w+)\s*;", r"thread_local usubpassInput \1;", text)
    # other resources
    text = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+", "", text)
    # parameter qualifiers
    def param(m):
        qual, typ, name, arr = m.group(1), m.group(2), m.group(3), m.group(4)
        if arr:
            return f"{typ} {name}["
        return f"{typ}& {name}"
    text = re.sub(r"\b(inout|out)\s+(\w+)\s+(\w+)(\s*\[)?", param, text)
    text = re.sub(r"(?<=[(,])\s*in\s+(?=\w+\s+\w+)", " ", text)
    text = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", text)
    # GLSL evaluates function arguments left to right; C++ leaves the order open (g++ goes right to
    # left), which would swap the two draws of get_noise_2 (noise_utility.glsl:87-89). Brace
    # initialisation is sequenced left to right.
    text, n_noise2 = re.subn(r"return\s+vec2\s*\(\s*(get_noise_gen\(accessor\.seed\))\s*,\s*(get_noise_gen\(accessor\.seed\))\s*\)\s*;",
                             r"return vec2{\1, \2};", text)
    assert n_noise2 == 1, "get_noise_2 not found"
    text = text.replace("using namespace", "using namespace")
    # make the interface-block namespaces visible
    for name in re.findall(r"namespace (\w+_ns) \{", text):
        # insert the using-directive after the block's closing "};"
        idx = text.index(f"namespace {name} {{")
        end = text.index("};", idx) + 2
        text = text[:end] + f"\nusing namespace {name};\n" + text[end:]
    return text


def defines(v):
    tech = v["tech"]
    vmax, vmin = v["vmax"], v["vmin"]
    clipping = tech in ("psa", "psa_biased", "ltc_cp")
    S, L = v["S"], v["L"]
    d = {
        "MATERIAL_COUNT": 256, "POLYGONAL_LIGHT_COUNT": "g_rt_light_count", "POLYGONAL_LIGHT_ARRAY_SIZE": 16400,
        "POLYGONAL_LIGHT_COUNT_CLAMPED": 33, "LIGHT_SAMPLES": L, "LIGHT_SAMPLES_CLAMPED": min(L, 33), "LIGHT_TEXTURE_COUNT": 1,
        "MIN_POLYGON_VERTEX_COUNT_BEFORE_CLIPPING": vmin, "MAX_POLYGONAL_LIGHT_VERTEX_COUNT": vmax,
        "MAX_POLYGON_VERTEX_COUNT": vmax + 1 if clipping else vmax,
        "SAMPLE_COUNT": S, "SAMPLE_COUNT_CLAMPED": min(S, 33),
        "MIS_HEURISTIC_BALANCE": int(v["mis"] == "balance"), "MIS_HEURISTIC_POWER": int(v["mis"] == "power"),
        "MIS_HEURISTIC_WEIGHTED": int(v["mis"] == "weighted"), "MIS_HEURISTIC_OPTIMAL_CLAMPED": int(v["mis"] == "optimal_clamped"),
        "MIS_HEURISTIC_OPTIMAL": int(v["mis"] == "optimal"),
        "SAMPLE_LIGHT_UNIFORM": int(v["light"] == "uniform"), "SAMPLE_LIGHT_RIS": int(v["light"] == "reservoir"),
        "SAMPLE_POLYGON_BASELINE": 0, "SAMPLE_POLYGON_AREA_TURK": int(tech == "turk"),
        "SAMPLE_POLYGON_PROJECTED_SOLID_ANGLE": int(tech in ("psa", "psa_biased")), "SAMPLE_POLYGON_LTC_CP": int(tech == "ltc_cp"),
        "USE_FAST_ATAN": int(v.get("fast_atan", 0)), "ERROR_DISPLAY_DIFFUSE": 0, "ERROR_DISPLAY_SPECULAR": 0, "ERROR_INDEX": 0,
    }
    flags = [f"-D{k}={val}" for k, val in d.items()]
    flags.append("-DUSE_BIASED_PROJECTED_SOLID_ANGLE_SAMPLING" if tech == "psa_biased" else "-DDONT_USE_BIASED_PROJECTED_SOLID_ANGLE_SAMPLING")
    return flags


def build_shading(names=None, verbose=False):
    OUT.mkdir(exist_ok=True)
    src = REF / "src" / "shaders" / "shading_pass.frag.glsl"
    tu = transform(inline_includes(src))
    built = []
    with tempfile.TemporaryDirectory(prefix="risltc_ref_") as tmp:
        tu_path = Path(tmp) / "shading_tu.inc"
        tu_path.write_text(tu)
        for name in (names or VARIANTS):
            out = OUT / f"libref_shading_{name}.so"
            contract = ["-ffp-contract=fast", "-mfma"] if VARIANTS[name].get("contract") else ["-ffp-contract=off"]
            cmd = [CXX, "-std=gnu++17", "-O2"] + contract + ["-fno-fast-math", "-fopenmp", "-fPIC", "-shared", "-w",
                   "-I", str(HERE), f'-DREF_SHADER_TU="{tu_path}"'] + defines(VARIANTS[name]) + ["-o", str(out), str(HERE / "ref_harness.cpp")]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            built.append(out)
    return built


def build_host():
    OUT.mkdir(exist_ok=True)
    out = OUT / "libref_host.so"
    src = REF / "src"
    cmd = [CC, "-std=gnu99", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-DGLFW_INCLUDE_NONE",
           "-I", str(src), "-I", str(REF / "ext" / "glfw" / "include"), "-o", str(out),
           str(src / "polygonal_light.c"), str(src / "camera.c"), str(HERE / "ref_host_stubs.c"), "-lm"]
    subprocess.check_call(cmd)
    return out


def main():
    if not REF.exists():
        print(f"{REF} is not present: nothing to build (prebuilt oracle/_ref/*.so are used as they are)")
        return 0
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or None
    print("built", build_host())
    for b in build_shading(names, verbose="-v" in sys.argv):
        print("built", b)
    return 0


if __name__ == "__main__":
    sys.exit(main())
