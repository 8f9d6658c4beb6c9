#!/usr/bin/env python3
"""bench.py -- shaded light-samples/s of the risltc shading path on B200 (BASELINE.json metric).

A step = one pass of the hot path over one batch: all `spp` accumulated frames of the workload
(G-buffer -> fused RIS + shading -> shadow rays + MIS sum + accumulation per frame), and for N > 1 the
single gather of the framebuffer stripes to rank 0. samples = W * H * spp * LIGHT_SAMPLES * 32 RIS
candidate evaluations (shading_pass.frag.glsl:726).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c1|c4] [--impl reference]

value  : device-resident throughput (scene, BVH, lights, LTC tables in HBM; per-frame constants are kernel
         arguments), CUDA events on the library's stream, max over ranks.
e2e    : the same metric through the C99 host layer (librisltc_host.so: write_lights -> upload,
         write_constants per frame -> render, read_accumulation_buffer into pinned host memory).
roofline: FP32 roofline of the fused shading kernel (algorithmic flop of SURVEY.md 8d over the kernel's
         CUDA-event time, against 2 * 128 lanes * 148 SMs * the SM clock sampled during the run).
cpu_baseline / --impl reference: the reference's own shaders compiled for the CPU (oracle/_ref) -- or the
         C oracle when they are absent -- on the host cores, on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

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
# algorithmic work per unit, SURVEY.md 8d (fma = 2 flop, MUFU op = 1 flop)
FLOP_PER_CANDIDATE = {3: 303.0, 4: 394.0}
FLOP_PER_SHADED_PIXEL = 1835.0 + 535.0 + 8.0
FP32_LANES_PER_SM, SM_COUNT, SM_MAX_MHZ = 128, 148, 1965.0
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of ris_ltc3_kernel (the dominant kernel of the shading pass)
# from an `ncu --set full` capture of this bench (profiles/r1_ncu_final_kernels.txt); far below any HBM bound, as the model says
RIS_KERNEL_DRAM_BYTES = {"c2": 8.5e6}


def make_workload(name):
    from risltc_b200 import ltc_fit, scenes
    desc, lights, boxes, occluders, W, H, spp, verts = WORKLOADS[name]
    if name == "c1":
        scene = scenes.quad_over_plane(W, H)
    else:
        scene = scenes.many_light_room(lights, boxes, seed=2, occluder_triangles=occluders, width=W, height=H, vertex_count=verts)
    fits = ltc_fit.fit_ggx_ltc(64, 51, 64)
    rgba, rg = ltc_fit.quantize_fits(fits)
    return dict(name=name, desc=desc, scene=scene, fits=fits, rgba=rgba, rg=rg, W=W, H=H, spp=spp, verts=verts, lights=lights)


def variant_kwargs(name, verts):
    if name == "c1":
        return dict(light_sampling="uniform", technique="projected_solid_angle", min_vertices=4, max_vertices=4)
    return dict(min_vertices=verts, max_vertices=verts)


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


def cpu_reference_run(wl, frames, rows=None):
    """The reference's shading_pass.frag.glsl compiled for the CPU (oracle/_ref) -- or the oracle port --
    over `frames` frames of the workload with all host threads. Returns (seconds, samples, kind, cores)."""
    from oracle import orc, ref
    W, H = wl["W"], wl["H"]
    kw = variant_kwargs(wl["name"], wl["verts"])
    osc = orc.OracleScene(wl["scene"], wl["rgba"], wl["rg"])
    cs = [orc.make_constants(wl["scene"], W, H, orc.frame_words(f)[0]) for f in range(frames)]
    ref_name = {"c1": "uni_psa_v4"}.get(wl["name"], "ris_ltc_v3")
    light_samples = 1
    if ref.available(ref_name):
        r = ref.RefShading(ref_name); r.bind(osc)
        t = time.perf_counter(); r.render(cs); dt = time.perf_counter() - t
        kind = "reference"
    else:
        t = time.perf_counter(); osc.render(cs, orc.variant(**kw)); dt = time.perf_counter() - t
        kind = "port"
    return dt, W * H * frames * light_samples * 32, kind, orc.thread_count()


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    wl = make_workload(args.workload)
    frames = 1 if args.workload != "c1" else 1
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
                                  sample=f"{frames} frame(s) of {wl['spp']} at {wl['W']}x{wl['H']} per step, all host threads (OpenMP over rows)"),
                e2e=dict(value=value, unit="Gsamples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fast", choices=["fast", "exact"])
    ap.add_argument("--stripe-height", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--emulate-stripes", type=int, default=0, help="profiling aid: render only stripe 0 of N on one GPU (the per-device share of an N-GPU run) and print its kernel times; not a bench line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return 0

    import torch
    import torch.distributed as dist
    from risltc_b200 import api, host, multi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = make_workload(args.workload)
    W, H, spp = wl["W"], wl["H"], wl["spp"]
    kw = variant_kwargs(args.workload, wl["verts"])

    # ---- the C99 host layer on generated scene files (the reference-shaped path); its device object is
    # also used for the device-resident measurement, so both legs run the same kernels on the same data
    tmp = tempfile.TemporaryDirectory(prefix=f"risltc_bench_{rank}_")
    vks, tex, save = host.write_scene_files(wl["scene"], tmp.name, ltc_fits=wl["fits"])
    app = host.Application(tmp.name, ordinal=local, stripe_height=args.stripe_height, stripe_index=rank if not args.emulate_stripes else 0,
                           stripe_count=args.emulate_stripes or world)
    app.load(vks, tex, save, W, H)
    app.settings(light_sampling=api.LIGHT[kw.get("light_sampling", "reservoir")],
                 polygon_sampling_technique=api.POLY[kw.get("technique", "ltc_cp")], accum=1)
    dev = app.device()
    dev.set_precision(args.precision)
    if args.emulate_stripes:
        acc = np.zeros(4)
        for i in range(args.warmup + args.steps):
            app.reset(0)
            app.render_frames(spp, upload_lights=False)
            if i >= args.warmup:
                acc += dev.last_kernel_ms()
        emit((dict(emulated_share=f"stripe 0 of {args.emulate_stripes}", workload=args.workload, per_step_ms=dict(zip(("gbuffer", "shade", "trace_resolve", "call"), (acc / args.steps).round(3).tolist())))))
        app.close()
        sys.stdout.flush()
        os._exit(0)
    gat = multi.StripeGather(W, H, args.stripe_height, rank, world, device=f"cuda:{local}")
    multi.attach(dev, gat)
    stream = torch.cuda.ExternalStream(int(api.lib().risltc_cuda_stream(dev.h)), device=local)
    host_frame = torch.empty((H, W, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kernel_ms = np.zeros(4)

    def resident_step(collect):
        # constants for the spp frames are written on the host (256 B each, passed as kernel arguments)
        app.reset(0)
        app.render_frames(spp, upload_lights=False)
        if collect:
            kernel_ms[:] += dev.last_kernel_ms()
        with torch.cuda.stream(stream):
            return gat.gather()

    def e2e_step():
        app.reset(0)
        app.render_frames(spp, upload_lights=True)      # write_lights -> H2D, write_constants x spp -> launches
        with torch.cuda.stream(stream):
            full = gat.gather()
            if rank == 0:
                host_frame.copy_(full, non_blocking=True)   # D2H of the accumulated frame into pinned memory
        stream.synchronize()

    # ---- device-resident leg
    for _ in range(args.warmup):
        resident_step(False)
    with torch.cuda.stream(stream):
        flush.fill_(1)      # evict L2 (the per-frame working set, > 250 MB at 1080p, exceeds L2 anyway)
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
    kernel_ms[:] *= args.steps
    launches = dev.counters()["launches"] - launches0
    if world >= 4:
        # a device that renders a small share overlaps consecutive frames on two streams (risltc_cuda_set_frame_overlap), so the
        # per-pass event intervals of the timed region overlap each other: take the breakdown from two extra, serial steps
        dev.set_frame_overlap("off")
        kernel_ms[:] = 0.0
        for _ in range(2):
            resident_step(True)
        kernel_ms[:] *= args.steps / 2.0
        dev.set_frame_overlap("auto")
        breakdown_note = "frames overlap in the timed region; per-pass times from two extra serial steps"
    counters = dev.counters()
    t = torch.tensor([ms, float(counters["shaded_pixels"]), float(counters["candidates"]), float(counters["shadow_rays"]), float(launches)] + list(kernel_ms),
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

    if rank == 0:
        # counters hold the last step (reset at each render_frames call): per-step totals over all ranks
        shaded, cands = float(t[1]), float(t[2])
        shade_ms_per_step = float(tmax[6]) / args.steps     # slowest rank's shading kernels, per step
        flop_per_step = cands * FLOP_PER_CANDIDATE[wl["verts"]] + shaded * FLOP_PER_SHADED_PIXEL
        sm_mhz = clock_info["sm_mhz"] if clock_info and clock_info["sm_mhz"] else SM_MAX_MHZ
        peak = 2.0 * FP32_LANES_PER_SM * SM_COUNT * sm_mhz * 1e6 / 1e12 * world
        achieved = flop_per_step / (shade_ms_per_step * 1e-3) / 1e12 if shade_ms_per_step > 0 else 0.0
        frames_per_step = spp
        roofline = dict(bound="fp32", kernel="shade_kernel (fused RIS + shading)", achieved=achieved, peak=peak, unit="TFLOP/s",
                        frac=achieved / peak if peak else None, traffic=RIS_KERNEL_DRAM_BYTES.get(args.workload),
                        peak_source=("2*128 lanes*148 SMs*SM clock sampled by nvidia-smi during the timed region" if clock_info and clock_info["sm_mhz"]
                                     else "2*128 lanes*148 SMs*1965 MHz (nominal max clock; nvidia-smi sampling unavailable)"),
                        flop_per_launch=flop_per_step / frames_per_step / world, ms_per_launch=shade_ms_per_step / frames_per_step,
                        flop_model="303 flop per RIS candidate (V=3) + 2378 per shaded pixel-sample, SURVEY.md 8d")
        kernels = dict(gbuffer_ms=float(tmax[5]) / args.steps, shade_ms=float(tmax[6]) / args.steps, resolve_ms=float(tmax[7]) / args.steps,
                       render_call_ms=float(tmax[8]) / args.steps, shadow_rays_per_step=float(t[3]), shaded_pixel_samples_per_step=shaded, note=breakdown_note)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            dt, samples, kind, cores = cpu_reference_run(wl, 1)
            n = 1
            if dt < 4.0:   # aim for 10-30 s of CPU work in total
                n = int(min(spp, max(1, round(12.0 / max(dt, 1e-3)))))
                dt, samples, kind, cores = cpu_reference_run(wl, n)
            cpu = dict(value=samples / dt / 1e9, unit="Gsamples/s", cores=cores, kind=kind,
                       sample=f"{n} of {spp} frames at {W}x{H}, all host threads (OpenMP over rows), {dt:.1f} s")
        line = dict(metric="shaded light-samples/sec", value=value, unit="Gsamples/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=wl["desc"], variant="light_reservoir (m=32) + sample_polygon_ltc_cp + mis_optimal_clamped, S=1, L=1" if args.workload != "c1" else "light_uniform + projected_solid_angle",
                                lights=wl["lights"], triangles=int(wl["scene"]["mesh"]["material_indices"].shape[0]), width=W, height=H, spp=spp,
                                precision=args.precision, parallelism=f"image stripes of {args.stripe_height} rows x{world}, scene replicated, one gather per step",
                                l2="per-frame working set (visibility, ray and accumulation buffers, %d MB) exceeds the 126 MB L2; L2 flushed before the timed region" % (W * H * 132 // world >> 20)),
                    clocks=clock_info, e2e=dict(value=e2e_value, unit="Gsamples/s", h2d_bytes_per_step=light_bytes + 256 * spp, d2h_bytes_per_step=W * H * 16,
                                                ms_per_step=1e3 * float(e2e_t[0]) / args.steps, api="librisltc_host.so: write_lights/write_constants -> risltc_cuda_render_frames -> read-back to pinned host memory"),
                    gpu_launches=int(float(t[4])), roofline=roofline, kernels=kernels, pixel_samples_per_s=W * H * spp * args.steps / (ms_max * 1e-3))
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    barrier()
    # tear down in dependency order: the device object references the gather slab; the library's own CUDA runtime
    # instance must not outlive torch's tensors at interpreter exit
    dev.set_accum_buffer(0)
    torch.cuda.synchronize()
    del gat, flush, host_frame, start, end
    torch.cuda.empty_cache()
    del stream
    app.close()
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    sys.exit(main())
