#!/usr/bin/env python3
"""bench.py -- shaded light-samples/s of the risltc shading path on B200 (BASELINE.json metric).

A step = one pass of the hot path over one batch: all `spp` accumulated frames of the workload
(G-buffer -> RIS candidates -> winner's estimator -> shadow rays -> MIS sum + accumulation per frame), and for
N > 1 the single gather of the framebuffer stripes to rank 0. samples = W * H * spp * LIGHT_SAMPLES * 32 RIS
candidate evaluations (shading_pass.frag.glsl:726).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl reference]

Workload: N = 1 runs C2 (BASELINE.json configs[1], the configuration the metric is quoted on for one GPU); N > 1 runs C3
(configs[2]: 4K, 1024 lights, 64 spp -- the configuration the 8-GPU scaling target is stated on) at EVERY N > 1, and its
line also carries `same_workload_one_gpu`, the same C3 step rendered by rank 0 alone in the same run. --workload overrides.
c5 is the many-light sweep (configs[4]): one line whose `sweep` lists every light count.

value  : device-resident throughput (scene, BVH, lights, LTC tables in HBM; per-frame constants are kernel
         arguments), CUDA events on the library's stream, max over ranks.
e2e    : the same metric through the C99 host layer (librisltc_host.so: write_lights -> upload,
         write_constants per frame -> render, read_accumulation_buffer into pinned host memory).
roofline: FP32 roofline of the fused shading pass (kernels 2a + 2b; algorithmic flop of SURVEY.md 8d over their
         CUDA-event time, against 2 * 128 lanes * 148 SMs * the SM clock sampled during the run).
roofline_trace: the shadow-ray kernel: rays/s, node and triangle bytes per ray from device counters, achieved GB/s.
cpu_baseline / --impl reference: the reference's own shaders compiled for the CPU (oracle/_ref) -- or the
         C oracle when they are absent -- on ALL host cores, on a bounded sample of the same workload.
"""
import os

# The CPU legs (cpu_baseline, --impl reference) run OpenMP over image rows on every host core. torch.distributed.run
# exports OMP_NUM_THREADS=1 to its workers, which would time the reference on one core: set it before anything loads an
# OpenMP runtime. Only rank 0 ever runs those legs, so the ranks do not oversubscribe the host.
os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import argparse  # noqa: E402
import json  # noqa: E402
import statistics  # noqa: E402
import subprocess  # noqa: E402
import sys  # noqa: E402
import tempfile  # noqa: E402
import time  # noqa: E402
from pathlib import Path  # noqa: E402

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

# The contract is ONE JSON line on stdout. The C host layer (like the reference's main.c) printf()s progress messages, so
# file descriptor 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor.
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    os.write(_STDOUT_FD, (json.dumps(obj) + "\n").encode())


WORKLOADS = {
    # name: (description, lights, boxes, occluder triangles, W, H, spp, light vertices)
    "c1": ("single quad area light over a diffuse plane, 640x360, 1 spp, polygon_sampling only", 1, 0, 0, 640, 360, 1, 4),
    "c2": ("procedural 64-light textured-polygon scene, 1920x1080, RIS+LTC MIS, 16 spp", 64, 200, 0, 1920, 1080, 16, 3),
    "c3": ("procedural 1024-light scene, 3840x2160, 64 spp, image-tile sharded", 1024, 4000, 0, 3840, 2160, 64, 3),
    "c4": ("high-occlusion procedural scene with 5M triangles, 1920x1080, 32 spp", 64, 200, 5_000_000, 1920, 1080, 32, 3),
}
C5_LIGHTS = [16, 64, 256, 1024, 4096, 16384]     # configs[4]: N = 16 * 4^k, C2 geometry, 1920x1080, 8 spp
C5_SPP = 8
# algorithmic work per unit, SURVEY.md 8d (fma = 2 flop, MUFU op = 1 flop)
FLOP_PER_CANDIDATE = {3: 303.0, 4: 394.0}
FLOP_PER_SHADED_PIXEL = 1835.0 + 535.0 + 8.0
FP32_LANES_PER_SM, SM_COUNT, SM_MAX_MHZ = 128, 148, 1965.0
L2_BYTES = 126 * 1024 * 1024
NODE_BYTES, TRIANGLE_BYTES, RAY_RECORD_BYTES = 64, 48, 48   # Qbvh4Node, BvhTri, {origin, ray_a, ray_b} per ray


def measured_peaks():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except (OSError, ValueError):
        return {}


def ncu_metrics(workload):
    """Per-launch hardware counters of the dominant kernels from the committed `ncu --set full` capture of this bench
    (profiles/r2_ncu_metrics.json, written by tools/ncu_metrics_json.py from the .ncu-rep files): DRAM bytes and executed
    FP32 operations. None where no capture of that workload is committed."""
    try:
        return json.loads((ROOT / "profiles" / "r2_ncu_metrics.json").read_text()).get(workload)
    except (OSError, ValueError):
        return None


def make_workload(name, lights=None):
    from risltc_b200 import ltc_fit, scenes
    if name == "c5":
        desc, _, boxes, occluders, W, H, _, verts = WORKLOADS["c2"]
        desc = f"many-light sweep, {lights} polygonal lights, C2 geometry, 1920x1080, {C5_SPP} spp"
        spp = C5_SPP
    else:
        desc, lights, boxes, occluders, W, H, spp, verts = WORKLOADS[name]
    if LIGHT_VERTICES and name != "c1":
        verts = LIGHT_VERTICES
        desc += f"; lights with {verts} vertices"
    if name == "c1":
        scene = scenes.quad_over_plane(W, H)
    else:
        scene = scenes.many_light_room(lights, boxes, seed=2, occluder_triangles=occluders, width=W, height=H, vertex_count=verts)
        if TEXTURED:
            scenes.add_procedural_textures(scene)
            desc += "; materials with mip-mapped BC1 / BC5 textures (base colour, specular, normal), textureGrad per shading point"
    fits = ltc_fit.fit_ggx_ltc(64, 51, 64)
    rgba, rg = ltc_fit.quantize_fits(fits)
    return dict(name=name, desc=desc, scene=scene, fits=fits, rgba=rgba, rg=rg, W=W, H=H, spp=spp, verts=verts, lights=lights)


# the five estimators of the reference's timing experiment (experiment_list.c:354-396): name -> light sampling, polygon sampling
ESTIMATORS = {
    "uniform_uniform": ("uniform", "area_turk"), "uniform_cp": ("uniform", "projected_solid_angle"), "uniform_area": ("reservoir", "area_turk"),
    "cp_cp": ("reservoir", "projected_solid_angle"), "ltc_cp": ("reservoir", "ltc_cp"),
    "uniform_ltc_cp": ("uniform", "ltc_cp"),   # not in the timing experiment: our estimator without the reservoir (specialised kernels, too)
}
ESTIMATOR = "ltc_cp"
LIGHT_VERTICES = 0  # --light-vertices 4: quad lights instead of the workload's triangles (reference variant ris_ltc_v4)
TEXTURED = False    # --textured: procedural mip-mapped BC1 / BC5 material textures instead of flat-colour materials


def variant_kwargs(name, verts):
    if name == "c1":
        return dict(light_sampling="uniform", technique="projected_solid_angle", min_vertices=4, max_vertices=4)
    kw = dict(min_vertices=verts, max_vertices=verts)
    if ESTIMATOR != "ltc_cp":
        kw.update(light_sampling=ESTIMATORS[ESTIMATOR][0], technique=ESTIMATORS[ESTIMATOR][1])
    return kw


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc = None
        self.path = Path(tempfile.mkstemp(prefix="risltc_clocks_", suffix=".csv")[1])
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            sm, mx, power, reasons = [], [], [], set()
            for line in self.path.read_text().splitlines():
                p = [x.strip() for x in line.split(",")]
                if len(p) < 7:
                    continue
                try:
                    sm.append(float(p[0])); mx.append(float(p[1])); power.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out = dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm), power_w=statistics.median(power))
        try:
            self.path.unlink()
        except OSError:
            pass
        return out


def cpu_reference_run(wl, frames):
    """The reference's shading_pass.frag.glsl compiled for the CPU (oracle/_ref) -- or the oracle port --
    over `frames` frames of the workload with all host threads. Returns (seconds, samples, kind, cores)."""
    from oracle import orc, ref
    W, H = wl["W"], wl["H"]
    kw = variant_kwargs(wl["name"], wl["verts"])
    osc = wl.get("_oracle_scene")
    if osc is None:
        osc = wl["_oracle_scene"] = orc.OracleScene(wl["scene"], wl["rgba"], wl["rg"])   # BVH build: not part of the timed work
    cs = [orc.make_constants(wl["scene"], W, H, orc.frame_words(f)[0]) for f in range(frames)]
    ref_name = {"c1": "uni_psa_v4"}.get(wl["name"], "ris_ltc_v4" if wl["verts"] == 4 else "ris_ltc_v3")
    light_samples = 1
    if ref.available(ref_name):
        r = ref.RefShading(ref_name); r.bind(osc)
        t = time.perf_counter(); r.render(cs); dt = time.perf_counter() - t
        kind = "reference"
    else:
        t = time.perf_counter(); osc.render(cs, orc.variant(**kw)); dt = time.perf_counter() - t
        kind = "port"
    return dt, W * H * frames * light_samples * 32, kind, orc.thread_count()


def cpu_baseline_block(wl):
    """A bounded sample (about 10-20 s) of the workload on the host cores."""
    spp, W, H = wl["spp"], wl["W"], wl["H"]
    dt, samples, kind, cores = cpu_reference_run(wl, 1)
    n = 1
    if dt < 4.0:
        n = int(min(spp, max(1, round(12.0 / max(dt, 1e-3)))))
        dt, samples, kind, cores = cpu_reference_run(wl, n)
    return dict(value=samples / dt / 1e9, unit="Gsamples/s", cores=cores, kind=kind,
                sample=f"{n} of {spp} frames at {W}x{H}, all host threads (OpenMP over rows), {dt:.1f} s")


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    name = args.workload if args.workload != "c5" else "c2"
    wl = make_workload(name)
    frames = 1          # one frame per step: 1-10 s of CPU work on a 16-32 core host for every workload
    for _ in range(args.warmup):
        cpu_reference_run(wl, 1)
    total_t, total_s, kind, cores = 0.0, 0, "port", 1
    for _ in range(args.steps):
        dt, samples, kind, cores = cpu_reference_run(wl, frames)
        total_t += dt; total_s += samples
    value = total_s / total_t / 1e9
    line = dict(impl="reference", metric="shaded light-samples/sec", value=value, unit="Gsamples/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * total_t / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f32", data="synthetic", config=dict(workload=wl["desc"], sample=f"{frames} of {wl['spp']} frames per step, full resolution"),
                cpu_baseline=dict(value=value, unit="Gsamples/s", cores=cores, kind=kind,
                                  sample=f"{frames} frame(s) of {wl['spp']} at {wl['W']}x{wl['H']} per step, {cores} host threads (OpenMP over rows)"),
                e2e=dict(value=value, unit="Gsamples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


class Renderer:
    """The C99 host layer on generated scene files (the reference-shaped path) for one workload; its device object is
    also used for the device-resident measurement, so both legs run the same kernels on the same data."""

    def __init__(self, wl, args, local, stripe_index, stripe_count, tag):
        from risltc_b200 import api, host
        self.wl = wl
        kw = variant_kwargs(wl["name"], wl["verts"])
        self.tmp = tempfile.TemporaryDirectory(prefix=f"risltc_bench_{tag}_")
        vks, tex, save = host.write_scene_files(wl["scene"], self.tmp.name, ltc_fits=wl["fits"])
        self.app = host.Application(self.tmp.name, ordinal=local, stripe_height=args.stripe_height, stripe_index=stripe_index, stripe_count=stripe_count)
        self.app.load(vks, tex, save, wl["W"], wl["H"])
        self.app.settings(light_sampling=api.LIGHT[kw.get("light_sampling", "reservoir")],
                          polygon_sampling_technique=api.POLY[kw.get("technique", "ltc_cp")], accum=1)
        self.dev = self.app.device()
        self.dev.set_precision(args.precision)

    def step(self, upload_lights=False):
        self.app.reset(0)
        self.app.render_frames(self.wl["spp"], upload_lights=upload_lights)

    def close(self):
        self.dev.set_accum_buffer(0)
        self.app.close()
        self.tmp.cleanup()


def measure(args, wl, rank, world, local, want_cpu, same_workload_one_gpu=False):
    """Both legs (device-resident, end-to-end) of one workload on `world` GPUs. Returns the JSON line on rank 0, else None."""
    import torch
    import torch.distributed as dist
    from risltc_b200 import api, multi
    W, H, spp = wl["W"], wl["H"], wl["spp"]
    r = Renderer(wl, args, local, rank, world, f"{rank}")
    app, dev = r.app, r.dev
    gat = multi.StripeGather(W, H, args.stripe_height, rank, world, device=f"cuda:{local}")
    multi.attach(dev, gat)
    stream = torch.cuda.ExternalStream(int(api.lib().risltc_cuda_stream(dev.h)), device=local)
    host_frame = torch.empty((H, W, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pass_ms = np.zeros(6)

    def resident_step(collect):
        # constants for the spp frames are written on the host (256 B each, passed as kernel arguments)
        r.step()
        if collect:
            pass_ms[:] += dev.last_pass_ms()
        with torch.cuda.stream(stream):
            return gat.gather()

    def e2e_step():
        r.step(upload_lights=True)      # write_lights -> H2D, write_constants x spp -> launches
        with torch.cuda.stream(stream):
            full = gat.gather()
            if rank == 0:
                host_frame.copy_(full, non_blocking=True)   # D2H of the accumulated frame into pinned memory
        stream.synchronize()

    # ---- device-resident leg
    for _ in range(args.warmup):
        resident_step(False)
    with torch.cuda.stream(stream):
        flush.fill_(1)      # evict L2 before the timed region
    barrier()
    launches0 = dev.counters()["launches"]
    clocks = ClockSampler(local) if rank == 0 else None
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record(stream)
    for i in range(args.steps):
        # the per-pass times are read back (a host synchronisation) for the LAST step only, so that the host stays ahead of
        # the device for the whole timed region; they are scaled to `steps` below
        resident_step(i == args.steps - 1)
    end.record(stream)
    barrier()
    clock_info = clocks.stop() if clocks else None
    ms = start.elapsed_time(end)
    breakdown_note = "CUDA events around each pass, last step of the timed region"
    pass_ms[:] *= args.steps
    launches = dev.counters()["launches"] - launches0
    if dev.frame_overlap_active():
        # up to 4.5 M pixels a device overlaps consecutive frames on two streams (risltc_cuda_set_frame_overlap), so the
        # per-pass event intervals of the timed region overlap each other: take the breakdown from two extra, serial steps
        dev.set_frame_overlap("off")
        pass_ms[:] = 0.0
        for _ in range(2):
            resident_step(True)
        pass_ms[:] *= args.steps / 2.0
        dev.set_frame_overlap("auto")
        breakdown_note = "frames overlap in the timed region; per-pass times from two extra serial steps"
    counters = dev.counters()
    # ---- traversal statistics of the shadow-ray kernel: one extra step with its counting instantiation
    dev.traversal_counters(True)
    resident_step(False)
    trav = dev.traversal_counters(False)
    # ---- the collective alone: one gather of the slabs as they are, between events on the library's stream (after a barrier, so
    # that no rank's wait for another rank's frames is counted)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        g0.record(stream)
        gat.gather()
        g1.record(stream)
    barrier()
    gather_ms = g0.elapsed_time(g1)
    t = torch.tensor([ms, float(counters["shaded_pixels"]), float(counters["candidates"]), float(counters["shadow_rays"]), float(launches)] + list(pass_ms)
                     + [float(trav["rays"]), float(trav["node_visits"]), float(trav["triangle_tests"]), float(trav["occluded"]), gather_ms],
                     dtype=torch.float64, device=f"cuda:{local}")
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms_max = float(tmax[0])
    samples_per_step = W * H * spp * 1 * 32
    value = samples_per_step * args.steps / (ms_max * 1e-3) / 1e9

    # ---- end-to-end leg
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_step * args.steps / float(e2e_t[0]) / 1e9
    light_bytes = len(app.write_lights())
    triangles = int(wl["scene"]["mesh"]["material_indices"].shape[0])
    bvh = dev.bvh_stats()

    # ---- tear down in dependency order before anything else is created. torch's allocators record an event on every stream
    # a block was used on when the block is freed, and the stream here belongs to the library's device object: the tensors
    # (the pinned frame above all) have to go while that stream still exists; only then is the device object destroyed.
    barrier()
    dev.set_accum_buffer(0)         # the device object must not point into the gather slab any more
    del gat, flush, host_frame, start, end
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    del stream
    r.close()

    line = None
    if rank == 0:
        steps = args.steps
        shaded, cands = float(t[1]), float(t[2])       # counters hold the last step: per-step totals over all ranks
        per_step = {k: float(tmax[5 + i]) / steps for i, k in enumerate(("gbuffer_ms", "ris_ms", "winner_ms", "trace_ms", "resolve_ms", "render_call_ms"))}
        shade_ms = per_step["ris_ms"] + per_step["winner_ms"]     # slowest rank's shading kernels, per step
        flop_per_step = cands * FLOP_PER_CANDIDATE[wl["verts"]] + shaded * FLOP_PER_SHADED_PIXEL
        sm_mhz = clock_info["sm_mhz"] if clock_info and clock_info["sm_mhz"] else SM_MAX_MHZ
        peak = 2.0 * FP32_LANES_PER_SM * SM_COUNT * sm_mhz * 1e6 / 1e12 * world
        achieved = flop_per_step / (shade_ms * 1e-3) / 1e12 if shade_ms > 0 else 0.0
        fast_path = wl["name"] != "c1" and ESTIMATOR == "ltc_cp"
        hw = ncu_metrics(wl["name"]) or {}
        roofline = dict(bound="fp32", kernel=(("ris_ltc4_kernel" if wl["verts"] == 4 else "ris_ltc3_kernel") + " (2a: 32 RIS candidates per pixel) + winner_kernel (2b: the chosen light's PSA + LTC MIS estimator)" if fast_path
                                              else "shade_kernel<4, true> (generic fused RIS + shading kernel)"),
                        achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak if peak else None,
                        traffic=hw.get("shading_dram_bytes_per_launch"), traffic_source=hw.get("source"),
                        peak_source=("2*128 lanes*148 SMs*SM clock sampled by nvidia-smi during the timed region (MEASURED_PEAKS.json holds HBM and bf16 tensor peaks only)"
                                     if clock_info and clock_info["sm_mhz"] else "2*128 lanes*148 SMs*1965 MHz (nominal max clock; nvidia-smi sampling unavailable)"),
                        flop_per_launch=flop_per_step / spp / world, ms_per_launch=shade_ms / spp,
                        flop_model="303 flop per RIS candidate (V=3; 394 for V=4) + 2378 per shaded pixel-sample, SURVEY.md 8d",
                        frac_ris_kernel=(cands * FLOP_PER_CANDIDATE[wl["verts"]] + shaded * 535.0) / (per_step["ris_ms"] * 1e-3) / 1e12 / peak if per_step["ris_ms"] > 0 and fast_path else None,
                        executed_flop_frac=(hw["shading_executed_flop_per_launch"] * spp * world / (shade_ms * 1e-3) / 1e12 / peak) if hw.get("shading_executed_flop_per_launch") and shade_ms > 0 else None)
        # shadow rays: algorithmic bytes per ray = node visits * 64 + triangle tests * 48 + the 48-byte ray record, from the device counters
        rays, nodes, tris, occluded = float(t[11]), float(t[12]), float(t[13]), float(t[14])
        trace_s = per_step["trace_ms"] * 1e-3
        peaks = measured_peaks()
        bytes_per_ray = (nodes * NODE_BYTES + tris * TRIANGLE_BYTES) / rays + RAY_RECORD_BYTES if rays else None
        hbm_peak = peaks.get("hbm_gbs", 6650.0) * world
        trace_gbs = bytes_per_ray * rays / trace_s / 1e9 if rays and trace_s > 0 else None
        roofline_trace = dict(kernel="trace4p_kernel (3: any-hit traversal of the 4-wide, 8-bit BVH, one traversal per pair of rays of a pixel; node visits and triangle tests are counted per traversal)", bound="latency / ALU pipe (the working set of nodes and triangles is L1/L2-resident; HBM only streams the ray records)",
                              rays_per_step=rays, grays_per_s=rays / trace_s / 1e9 if trace_s > 0 else None, ms_per_launch=per_step["trace_ms"] / spp,
                              node_visits_per_ray=nodes / rays if rays else None, triangle_tests_per_ray=tris / rays if rays else None,
                              occluded_fraction=occluded / rays if rays else None, bytes_per_ray=bytes_per_ray,
                              achieved=trace_gbs, unit="GB/s", peak=hbm_peak, peak_source=("MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s") + " x GPUs",
                              frac=trace_gbs / hbm_peak if trace_gbs else None,
                              hbm_bytes_per_ray=RAY_RECORD_BYTES, hbm_frac=(RAY_RECORD_BYTES * rays / trace_s / 1e9 / hbm_peak) if rays and trace_s > 0 else None,
                              traffic=hw.get("trace_dram_bytes_per_launch"),
                              note="achieved = (node visits x 64 B + triangle tests x 48 B + 48 B ray record) x rays / kernel time: bytes the kernel requests from L1/L2, of which only the ray records (hbm_bytes_per_ray) must come from HBM")
        kernels = dict(per_step, shade_ms=shade_ms, gather_ms=float(tmax[15]), shadow_rays_per_step=float(t[3]), shaded_pixel_samples_per_step=shaded, note=breakdown_note)
        per_device_mb = (W * H // world) * 140 / 2 ** 20      # visibility, pick, origin, base, group, two ray slots, accumulation
        line = dict(metric="shaded light-samples/sec", value=value, unit="Gsamples/s", n_gpus=world, steps=steps, warmup=args.warmup,
                    ms_per_step=ms_max / steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=wl["desc"], variant=("light_reservoir (m=32) + sample_polygon_ltc_cp + mis_optimal_clamped, S=1, L=1" if fast_path else "light_uniform + projected_solid_angle" if wl["name"] == "c1"
                                                            else f"timing-experiment estimator {ESTIMATOR}: light_{ESTIMATORS[ESTIMATOR][0]} + sample_polygon_{ESTIMATORS[ESTIMATOR][1]} ({'specialised pick + winner kernels' if ESTIMATOR == 'uniform_ltc_cp' else 'generic kernel'})"),
                                lights=wl["lights"], triangles=triangles, width=W, height=H, spp=spp,
                                acceleration_structure=dict(builder=bvh["builder"], build_ms=round(bvh["build_ms"], 1), device_ms=[round(x, 2) for x in bvh["device_ms"]],
                                                            wide_nodes=bvh["wide_nodes"], depth=[bvh["binary_depth"], bvh["wide_depth"]],
                                                            note="built once at upload_scene, outside the timed region (scene.c:142-406 builds once per scene, too)"),
                                precision=args.precision, parallelism=f"image stripes of {args.stripe_height} rows x{world}, scene replicated, one gather per step",
                                l2=f"L2 flushed (192 MiB fill) before the timed region; every frame streams {per_device_mb:.0f} MB of per-pixel buffers per device through a 126 MB L2"
                                   + (" (larger than L2)" if per_device_mb * 2 ** 20 > L2_BYTES else " (smaller than L2: frames of a step reuse it, as they do in production)")),
                    clocks=clock_info, e2e=dict(value=e2e_value, unit="Gsamples/s", h2d_bytes_per_step=light_bytes + 256 * spp, d2h_bytes_per_step=W * H * 16,
                                                ms_per_step=1e3 * float(e2e_t[0]) / steps, api="librisltc_host.so: write_lights/write_constants -> risltc_cuda_render_frames -> read-back to pinned host memory"),
                    gpu_launches=int(float(t[4])), roofline=roofline, roofline_trace=roofline_trace, kernels=kernels, pixel_samples_per_s=W * H * spp * steps / (ms_max * 1e-3))
        if same_workload_one_gpu and world > 1:
            # the same workload rendered by this rank alone (whole frame, no gather), so that the line carries its own 1-GPU point
            solo = Renderer(wl, args, local, 0, 1, "solo")
            solo.step(); solo.dev.synchronize()
            t0 = time.perf_counter(); solo.step(); solo.dev.synchronize(); one_s = time.perf_counter() - t0
            solo_ms = solo.dev.last_pass_ms()[5]
            solo.close()
            line["same_workload_one_gpu"] = dict(value=samples_per_step / (solo_ms * 1e-3) / 1e9, unit="Gsamples/s", ms_per_step=solo_ms, wall_ms=one_s * 1e3,
                                                 note="one step of this workload rendered by rank 0 alone in the same run (whole frame, CUDA events around the render call)")
        if want_cpu:
            line["cpu_baseline"] = cpu_baseline_block(wl)
    barrier()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + ["c5"], help="default: c2 on one GPU, c3 on several")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fast", choices=["fast", "exact"])
    ap.add_argument("--stripe-height", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--estimator", default="ltc_cp", choices=sorted(ESTIMATORS), help="one of the five estimators of the reference's timing experiment (experiment_list.c:354-396); default: ours")
    ap.add_argument("--light-vertices", type=int, default=0, choices=[0, 3, 4], help="4: quad lights (the specialised kernels' second instantiation)")
    ap.add_argument("--textured", action="store_true", help="material textures (BC1 / BC5 .vkt files with mip chains through load_scene) instead of flat-colour materials")
    ap.add_argument("--emulate-stripes", type=int, default=0, help="profiling aid: render only stripe 0 of N on one GPU (the per-device share of an N-GPU run) and print its kernel times; not a bench line")
    args = ap.parse_args()
    global ESTIMATOR, TEXTURED, LIGHT_VERTICES
    ESTIMATOR, TEXTURED, LIGHT_VERTICES = args.estimator, args.textured, args.light_vertices
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        args.workload = "c2" if max(world, args.gpus) == 1 else "c3"
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.emulate_stripes:
        wl = make_workload(args.workload)
        r = Renderer(wl, args, local, 0, args.emulate_stripes, "emu")
        acc = np.zeros(6)
        for i in range(args.warmup + args.steps):
            r.step()
            if i >= args.warmup:
                acc += r.dev.last_pass_ms()
        emit(dict(emulated_share=f"stripe 0 of {args.emulate_stripes}", workload=args.workload,
                  per_step_ms=dict(zip(("gbuffer", "ris", "winner", "trace", "resolve", "call"), (acc / args.steps).round(3).tolist()))))
        r.close()
    elif args.workload == "c5":
        sweep, line = [], None
        for lights in C5_LIGHTS:
            wl = make_workload("c5", lights)
            one = measure(args, wl, rank, world, local, want_cpu=False)
            if rank == 0:
                sweep.append(dict(lights=lights, value=one["value"], ms_per_step=one["ms_per_step"], e2e=one["e2e"]["value"], frac=one["roofline"]["frac"],
                                  ris_ms=one["kernels"]["ris_ms"], winner_ms=one["kernels"]["winner_ms"], trace_ms=one["kernels"]["trace_ms"],
                                  light_table="shared memory" if lights <= 2048 else "global memory (L1/L2)"))
                if lights == 1024:
                    line = one
        if rank == 0:
            line["config"]["workload"] = f"many-light sweep from 16 to 16384 polygonal lights, C2 geometry, 1920x1080, {C5_SPP} spp; headline fields are the 1024-light point"
            line["sweep"] = sweep
            emit(line)
    else:
        wl = make_workload(args.workload)
        line = measure(args, wl, rank, world, local, want_cpu=not args.no_cpu_baseline, same_workload_one_gpu=True)
        if rank == 0:
            emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    torch.cuda.synchronize()
    sys.stdout.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
