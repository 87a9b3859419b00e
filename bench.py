#!/usr/bin/env python3
"""bench.py — measures BASELINE.json's metric (Msamples/s of the spectral path trace) on its
config 1: killeroo-simple, path integrator maxdepth 5, 32-band SampledSpectrum as the reference
ships it, LD sampler 64 spp, 700x700, box filter.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                       # the reference's own CPU renderer

A step is one whole render of the frame (every camera sample of the sample extent, 701 x 701 x 64 =
31 449 664, the same count the reference traces). `value` times spt_render with the scene resident
in HBM (CUDA events on the library's stream); `e2e` times the C-ABI call sequence a host makes with
HOST buffers: spt_scene_create (H2D of every scene table) -> spt_film_create -> spt_render ->
spt_film_download (D2H of the film). Multi-GPU: scene replicated, image tile sets per rank, film
reduced to rank 0 with NCCL; device-timed, max over ranks.
"""
import argparse
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "killeroo_path"
WORKLOAD_DESC = "scenes/killeroo-simple.pbrt, path integrator maxdepth 5, SampledSpectrum 32 bands (as shipped), LD 64 spp, 700x700, box filter"
SCENE_SPT = os.path.join(ROOT, "assets", "_lowered", WORKLOAD + ".spt")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt")
REF_SCENES = os.path.join(ROOT, "oracle", "_ref", "scenes")
CPU_SAMPLE_SPP = 16          # bounded sample of the workload for the in-line cpu_baseline leg: same frame, 16 of the 64 spp
REF_ARM_SPP = int(os.environ.get("SPT_REF_ARM_SPP", "64"))     # (the CPU test of this arm's output format lowers it)
# --impl reference renders the whole configuration (64 spp): ~9 s per step on 16 cores
# workloads that have a scene file the reference binary can run for the CPU leg: (xres, yres, spp, bounded-sample spp)
# synthetic workloads built through the reference's API (oracle/synth_scene.cpp): (triangles, xres, yres, spp)
SYNTH_API = {"synth_10m": (10000000, 3840, 2160, 1024), "synth_10m_1080p64": (10000000, 1920, 1080, 64)}
CPU_WORKLOADS = {"killeroo_path": (700, 700, 64, CPU_SAMPLE_SPP), "killeroo_direct": (700, 700, 64, CPU_SAMPLE_SPP),
                 "bunny_shipped": (640, 480, 256, 8), "metal_path": (400, 400, 512, 16), "ssenv_path": (1920, 1080, 1024, 8)}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([v.strip() for v in line.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
def reference_run(scene_text, ncores):
    """One run of the reference binary on scene_text; returns wall seconds."""
    d = tempfile.mkdtemp(prefix="sptref_")
    try:
        for sub in ("geometry", "spds", "brdfs", "textures"):
            os.symlink(os.path.join(REF_SCENES, sub), os.path.join(d, sub))
        with open(os.path.join(d, "scene.pbrt"), "w") as f:
            f.write(scene_text)
        t0 = time.perf_counter()
        subprocess.run([REF_BIN, "--quiet", "--ncores", str(ncores), "scene.pbrt"], cwd=d, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return time.perf_counter() - t0
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_scene(spp, res=None, workload=WORKLOAD):
    s = open(os.path.join(REF_SCENES, workload + ".pbrt")).read()
    s = re.sub(r'"integer pixelsamples" \[\d+\]', '"integer pixelsamples" [%d]' % spp, s)
    if res:
        s = re.sub(r'"integer xresolution" \[\d+\] "integer yresolution" \[\d+\]',
                   '"integer xresolution" [%d] "integer yresolution" [%d]' % (res, res), s, count=1)
    return s


def cpu_reference_msamples(runs=1, workload=WORKLOAD):
    """Msamples/s of the unmodified reference (oracle/_ref/bin/pbrt, all host cores) on the bounded
    sample: the workload's frame at a fraction of its spp. Parse + BVH build time (the same scene at 8x8, 1 spp)
    is subtracted, SURVEY.md 8d."""
    ncores = os.cpu_count() or 1
    xres, yres, _, cpu_spp = CPU_WORKLOADS[workload]
    setup = min(reference_run(reference_scene(1, 8, workload), ncores) for _ in range(2))
    best = min(reference_run(reference_scene(cpu_spp, None, workload), ncores) for _ in range(runs))
    n_samples = (xres + 1) * (yres + 1) * cpu_spp
    render = max(best - setup, 1e-6)
    return n_samples / render / 1e6, ncores, render, setup


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REF_BIN):
        print_json({"impl": "reference", "unavailable": "oracle/_ref/bin/pbrt not built (build() needs /root/reference)"})
        return
    ncores = os.cpu_count() or 1
    setup = min(reference_run(reference_scene(1, 8), ncores) for _ in range(2))
    times = []
    for i in range(args.warmup + args.steps):
        t = reference_run(reference_scene(REF_ARM_SPP), ncores)
        if i >= args.warmup:
            times.append(max(t - setup, 1e-6))
    n_samples = 701 * 701 * REF_ARM_SPP
    ms = 1e3 * sum(times) / len(times)
    value = n_samples / (ms / 1e3) / 1e6
    sample = "the whole configuration per step: 700x700 (sample extent 701x701) at all %d spp; parse+BVH build (%.2fs, 8x8 1spp run) subtracted" % (REF_ARM_SPP, setup)
    print_json({
        "impl": "reference", "metric": "Msamples/sec (%d-band spectral path trace)" % args.bands, "value": value, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "reference scene file (killeroo-simple) derived per SURVEY.md F9",
        "config": {"workload": WORKLOAD_DESC, "reference_binary": "oracle/_ref/bin/pbrt (unmodified reference, g++ -O2 -m64)" if args.bands == 32 else
                   "oracle/_ref/bin30/pbrt (the reference with the one line nSpectralSamples = 32 -> 30 of src/core/spectrum.h patched, g++ -O2 -m64)"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": ncores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wave-pixels", type=int, default=0)
    ap.add_argument("--frames-ahead", type=int, default=0, help="frames enqueued ahead of the one the host waits for (1..3; default 1 on one GPU, 2 on several)")
    ap.add_argument("--bands", type=int, default=32, choices=[32, 30], help="32: SampledSpectrum as the reference ships it (default); 30: the 30-band "
                    "variant BASELINE.json's metric names - libspt30.so against the reference built with nSpectralSamples = 30 (oracle/_ref/bin30)")
    ap.add_argument("--workload", default=None, help="lowered scene under assets/_lowered (default: BASELINE config 1); "
                    "synth_1m = config 5's recipe at 1 M triangles, BVH larger than L2 (no CPU arm)")
    args = ap.parse_args()
    global REF_BIN, WORKLOAD, WORKLOAD_DESC
    if args.bands == 30:
        os.environ["SPT_NBANDS"] = "30"                  # read by pbrt_v2_spectral_b200.ctypes_defs at import
        REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin30", "pbrt")
        WORKLOAD = "killeroo_path30"
        WORKLOAD_DESC = WORKLOAD_DESC.replace("SampledSpectrum 32 bands (as shipped)", "SampledSpectrum 30 bands (nSpectralSamples = 30, src/core/spectrum.h:43)")
        CPU_WORKLOADS[WORKLOAD] = CPU_WORKLOADS["killeroo_path"]
    if args.workload is None:
        args.workload = WORKLOAD
    # stdout carries exactly ONE JSON line: everything else written to file descriptor 1 by this process or by libraries
    # under it (NCCL's version banner, which it prints itself whatever torch's logging is set to) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global print_json

    def print_json(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())
    if args.impl == "reference":
        run_reference_arm(args)
        return

    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"        # keep stdout to the one JSON line
    import numpy as np
    import torch
    from pbrt_v2_spectral_b200 import capi, ctypes_defs as D, multi
    from pbrt_v2_spectral_b200.scene_io import LoweredScene

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    capi.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    scene_spt = os.path.join(ROOT, "assets", "_lowered", args.workload + ".spt")
    synth_tool = os.path.join(ROOT, "oracle", "_ref", "bin", "synth_scene")
    if not os.path.exists(scene_spt) and args.workload in SYNTH_API and os.path.exists(synth_tool):
        # BASELINE config 5 at full size: the scene is built through the reference's own API (oracle/synth_scene.cpp: its
        # Shape classes, its BVHAccel build) and lowered - scene preparation on the host, outside every timed region
        if rank == 0:
            ntris, xr, yr, spp_w = SYNTH_API[args.workload]
            os.makedirs(os.path.dirname(scene_spt), exist_ok=True)
            t_gen = time.perf_counter()
            subprocess.run([synth_tool, str(ntris), str(xr), str(yr), str(spp_w)], check=True, stdout=sys.stderr, cwd=os.path.dirname(scene_spt),
                           env=dict(os.environ, SPT_DUMP_PREFIX=scene_spt[:-4], SPT_DUMP_PIXELS="1", SPT_DUMP_LI="0", SPT_DUMP_NRNG="1"))
            os.remove(scene_spt[:-4] + ".golden")
            sys.stderr.write("[bench] %s built by the reference's API + BVHAccel and lowered in %.1f s\n" % (args.workload, time.perf_counter() - t_gen))
        if dist is not None:
            dist.barrier()
    if not os.path.exists(scene_spt) and args.workload.startswith("synth") and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bin", "oracle_dump")):
        # synthetic workloads are generated on the spot: .pbrt text -> the built reference's own parser + BVH build (the host
        # side of the boundary, which stays on the CPU by design) -> lowered scene. Scene preparation, outside every timed region.
        if rank == 0:
            subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "make_golden.py"), "--only", args.workload], check=True,
                           stdout=sys.stderr)
        if dist is not None:
            dist.barrier()
    if not os.path.exists(scene_spt):
        raise SystemExit("lowered workload scene %s missing: run __graft_entry__.build() where the reference tree is" % scene_spt)
    lowered = LoweredScene.load(scene_spt)
    workload_desc = WORKLOAD_DESC if args.workload == WORKLOAD else {
        "killeroo_direct": "scenes/killeroo-simple.pbrt AS SHIPPED: directlighting integrator (strategy all, sphere light nsamples 8), "
                           "LD 64 spp, 700x700, box filter",
        "bunny_shipped": "scenes/bunny.pbrt AS SHIPPED (BASELINE config 2's scene): directlighting integrator (point light + disk area light, "
                         "nsamples 4), measured BRDF brdfs/mystique.brdf on the bunny, 640x480, LD 256 spp, box filter",
        "metal_path": "scenes/metal.pbrt (BASELINE config 3) with its shipped floor - substrate, lines.exr as EWA-filtered Kd and as bump map - "
                      "Au teapot (measured eta/k SPDs, Blinn exponent 1000), grace environment map for the absent uffizi map, path maxdepth 5, "
                      "400x400, LD 512 spp, box filter",
        "ssenv_path": "scenes/ss-envmap.pbrt (BASELINE config 4) as shipped except the path integrator: subsurface teapot, substrate floor with "
                      "image-mapped Kd and bump map, grace environment map importance sampled, path maxdepth 5, 1920x1080, LD 1024 spp, box filter",
        "synth_10m": "BASELINE config 5 at full size: synthetic random-triangle scene, 10 000 000 triangles (BVH + vertices 1.9 GB in HBM), half matte / "
                     "half plastic, sphere area light + constant infinite light, path maxdepth 5, 3840x2160, LD 1024 spp, box filter; built "
                     "through the reference's API and BVHAccel (oracle/synth_scene.cpp)",
        "synth_1m": "synthetic random-triangle scene (BASELINE config 5 recipe, SURVEY 8d) at 1 000 000 triangles, matte + plastic, sphere area light + "
                    "constant infinite light, path maxdepth 5, 1024x576, LD 16 spp, box filter"}.get(args.workload, args.workload)
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
    rp.seed = 1
    rp = multi.rank_params(rp, rank, world)            # this rank's tile set (32x32 tiles, round-robin)
    rp.wave_pixels = args.wave_pixels
    fd = lowered.film
    n_samples_total = (rp.x_end - rp.x_start) * (rp.y_end - rp.y_start) * rp.spp

    scene = capi.Scene(lowered)
    # N = 1: the film is the library's. N > 1: ONE film per frame buffer lives on rank 0 (spt_film_create there, exported with
    # spt_film_ipc_export); the other ranks open it (spt_film_open_ipc) and their film kernel adds their tile sets' samples
    # straight into it over NVLink - the "gather" is fused into K7, what is left of it is a barrier at the end of the frame.
    # frames enqueued ahead of the one the host waits for: two where a rank's frame is small enough to run as ONE wave (up to 2^23
    # paths: the library then rotates consecutive frames over its lanes, and three whole frames overlap), else one
    one_wave = world > 1 and n_samples_total // world <= (1 << 23)
    ahead = min(max(args.frames_ahead, 1), 3) if args.frames_ahead else (2 if one_wave else 1)
    n_buf = ahead + 2 if world > 1 else 1
    if world == 1:
        films = [capi.Film(fd)]
    else:
        handles = [None] * n_buf
        if rank == 0:
            films = [capi.Film(fd) for _ in range(n_buf)]
            handles = [f.ipc_export() for f in films]
        dist.broadcast_object_list(handles, src=0)
        if rank != 0:
            films = [capi.Film(fd, ipc_handle=h) for h in handles]
    film = films[0]
    sync_t = torch.zeros(1, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.all_reduce(sync_t)
        torch.cuda.synchronize()

    def frame(k, params=None):
        """One whole frame of the job on N GPUs: every rank renders its tile set into film buffer k % n_buf (rank 0's memory),
        the end-of-frame barrier makes the film complete, and rank 0 clears it for the frame after next."""
        f = films[k % n_buf]
        scene.render_begin(f, params if params is not None else rp)  # the entry points of the timed region (same wave layout)
        scene.render_end()                                            # blocks until this rank's streams have drained
        if dist is not None:
            dist.all_reduce(sync_t)
            torch.cuda.current_stream().synchronize()
        return f

    # ---- untimed counted pass: BVH nodes visited / primitives tested per ray class (roofline's algorithmic bytes)
    scene.enable_counters(True)
    scene.render(film, rp)
    cst = scene.stats()
    scene.enable_counters(False)
    # the kernel counts per ray CLASS (closest-hit / any-hit query), whatever queue the ray came from
    n_closest, n_any = max(cst["closest_rays"], 1), max(cst["any_rays"], 1)
    nodes_per_closest = cst["node_visits_closest"] / n_closest
    prims_per_closest = cst["prim_tests_closest"] / n_closest
    nodes_per_any = cst["node_visits_any"] / n_any
    prims_per_any = cst["prim_tests_any"] / n_any
    # MIS rays traced as closest-hit queries (the others - towards an infinite light - are any-hit queries)
    mis_closest_frac = (cst["closest_rays"] - cst["class_rays"][D.K_TRACE_PATH]) / max(cst["class_rays"][D.K_TRACE_MIS], 1)

    # ---- warm-up, then the timed steps
    if rank == 0:
        film.clear()
    barrier()
    # (frames of a minute: the counted pass above is the first of the three warm-up frames)
    for k in range(max(args.warmup, 3) - (1 if args.workload.startswith("synth_10m") else 0)):
        f = frame(k)
        if rank == 0:
            f.clear()
    barrier()
    launches0 = scene.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    render_ms = 0.0
    lanes_used = 1
    barrier()
    e0.record()
    # The frames are pipelined: a rank ENQUEUES frames k + 1 .. k + A (spt_render_begin; A = 1 on one GPU, 2 on several) before
    # it waits for frame k (spt_render_end), so its GPU goes from frame to frame without a host bubble, and on a 1/N share of
    # the image - where a frame is one wave - consecutive frames run side by side on different lanes. N > 1: A + 2 film
    # buffers on rank 0. Frame k's end-of-frame barrier (a one-element all-reduce: once it completes every rank's samples are
    # in buffer k % (A + 2)) is enqueued behind the rank's render. Rank 0 clears buffer (k - 1) % (A + 2) - complete since
    # barrier k - 1 - BEFORE it enters barrier k; frame k + A + 1, the next to write that buffer, is enqueued on every rank only
    # once barrier k has completed. Every barrier has completed before the closing time stamp. N = 1: one film, the frames add
    # up in it.
    def _end(k):
        nonlocal render_ms
        scene.render_end()                                # blocks until frame k has drained on this rank
        render_ms += scene.render_ms()

    multi.pipelined_frames(args.steps, ahead, n_buf, lambda k, b: scene.render_begin(films[b], rp), _end,
                           clear=lambda b: films[b].clear_idle(), rank=rank, sync_t=sync_t)
    if dist is not None and rank == 0:
        films[(args.steps - 1) % n_buf].clear_idle()
    e1.record()
    barrier()
    lanes_used = scene.stats()["lanes_used"]
    if sampler:
        sampler.stop_flag = True
        sampler.join()
    launches = scene.stats()["kernel_launches"] - launches0
    # ---- per-kernel device time: in the timed region the waves of a frame overlap on two streams, so the
    # CUDA-event deltas of a kernel there include the other lane's work. The same frames are rendered
    # again with every wave on ONE stream and each kernel timed between its own events.
    class_ms = [0.0] * D.K_CLASSES
    class_launches = [0] * D.K_CLASSES
    class_rays = [0] * D.K_CLASSES
    elided = 0
    first_vertices = 0
    first_ms = [0.0] * D.K_CLASSES                   # bounce-0 launch of each class: mean device time and units over the profiled frames
    first_units = [0.0] * D.K_CLASSES
    serial_ms = 0.0
    prof_steps = max(1, min(args.steps, 5 if not args.workload.startswith("synth_10m") else 1))
    scene.set_lanes(1)
    if not args.workload.startswith("synth_10m"):
        scene.render(film, rp)
    for _ in range(prof_steps):
        scene.render(film, rp)
        st = scene.stats()
        serial_ms += st["render_ms"]
        elided += st["mis_rays_elided"]
        first_vertices += st["first_vertices"]
        for k in range(D.K_CLASSES):
            first_ms[k] += st["first_launch_ms"][k] / prof_steps; first_units[k] += st["first_launch_units"][k] / prof_steps
        for k in range(D.K_CLASSES):
            class_ms[k] += st["class_ms"][k]; class_launches[k] += st["class_launches"][k]; class_rays[k] += st["class_rays"][k]
    scene.set_lanes(int(os.environ.get("SPT_LANES", "4")))
    barrier()
    # ---- the same frames with the FAST traversal layout (include/spt.h: spt_scene_set_traversal): reported
    # beside the headline, which stays on the bit-exact walk
    fast_mode = None
    big = args.workload.startswith("synth_10m")         # frames of a minute: the extra passes are kept to one frame each
    if world == 1:
        t_build = time.perf_counter()
        scene.set_traversal(True)
        t_build = time.perf_counter() - t_build
        fsteps = max(1, min(args.steps, 5)) if not big else 1
        for _ in range(2 if not big else 1):
            scene.render(film, rp)
        fast_ms = 0.0
        for _ in range(fsteps):
            scene.render(film, rp)
            fast_ms += scene.render_ms()
        scene.set_traversal(False)
        fast_mode = {"ms_per_step": fast_ms / fsteps, "value": n_samples_total / (fast_ms / fsteps / 1e3) / 1e6, "unit": "Msamples/s",
                     "layout": "4-wide BVH collapsed from the reference's flattened tree, children entered nearest first (csrc/wide.h)",
                     "build_s": t_build,
                     "note": "not the parity path: same primitive and bit-identical distance wherever the closest hit is unique (tests/: no "
                             "difference on 2 x 1 048 576 camera rays of configs 1 and 2)"}
    # ---- the complete film of one more frame (N > 1: assembled on rank 0 over NVLink), and for N > 1 its check against a
    # one-GPU render of the same seed on rank 0
    if rank == 0:
        film.clear()
    barrier()
    frame(0)
    barrier()
    image_sum = 0.0
    film_check = None
    if rank == 0:
        c_all, w_all = film.download()
        image_sum = float(c_all.sum(dtype=np.float64))
        if world > 1:
            rp1 = multi.rank_params(rp, 0, 1)
            f1 = capi.Film(fd)
            scene.render(f1, rp1)
            c_one, w_one = f1.download()
            f1.close()
            scale = np.maximum(np.abs(c_one).max(axis=2, keepdims=True), 1e-20)
            rel = float((np.abs(c_all - c_one) / scale).max())
            film_check = {"max_rel_err_vs_1gpu_render": rel, "weights_equal": bool(np.array_equal(w_all, w_one)), "tolerance": 1e-5,
                          "ok": bool(rel <= 1e-5 and np.array_equal(w_all, w_one))}
            if not film_check["ok"]:
                raise SystemExit("N-GPU film differs from the 1-GPU render of the same seed: %r" % film_check)
        c_all = w_all = None
    barrier()
    step_ms = e0.elapsed_time(e1) / args.steps
    sys.stderr.write("[rank %d] render %.3f ms/step (library events), step incl. film zero + reduce %.3f ms, serialized kernels %.3f ms\n" % (rank, render_ms / args.steps, step_ms, serial_ms / prof_steps))
    if dist is not None:
        t = torch.tensor([step_ms, render_ms / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, render_only_ms = float(t[0]), float(t[1])
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt[0])
    else:
        render_only_ms = render_ms / args.steps
    # the K timed steps as one job: wall clock between the two time stamps / K, max over ranks (N > 1: incl. the end-of-frame barriers)
    ms_per_step = step_ms
    value = n_samples_total / (ms_per_step / 1e3) / 1e6

    # ---- e2e through the C ABI with host buffers (rank-local: each rank uploads, renders its tiles, downloads)
    scene_bytes = int(sum(v.nbytes for k, v in lowered.a.items() if k not in ("camera", "film", "params", "film_filename")))
    film_bytes = fd.x_pixel_count * fd.y_pixel_count * (D.NBANDS + 1) * 4
    # the host's film buffers are page-locked (spt_host_alloc), as a host that reads a film back every frame would hold them
    c_pin = capi.HostBuffer((fd.y_pixel_count, fd.x_pixel_count, D.NBANDS))
    w_pin = capi.HostBuffer((fd.y_pixel_count, fd.x_pixel_count))
    c_host, w_host = c_pin.array, w_pin.array
    e2e_times = []
    for i in range(1 + args.steps if not args.workload.startswith("synth_10m") else 1):
        barrier()
        t0 = time.perf_counter()
        sc2 = capi.Scene(lowered)                    # H2D: every scene table from host memory (every rank)
        if world == 1:
            f2 = capi.Film(fd)
        else:
            f2 = films[0]                            # the shared film on rank 0 (cleared below, inside the timed region)
            if rank == 0:
                f2.clear()
            dist.all_reduce(sync_t); torch.cuda.current_stream().synchronize()
        sc2.render(f2, rp)
        if world > 1:
            dist.all_reduce(sync_t); torch.cuda.current_stream().synchronize()      # end of frame: the film is complete on rank 0
        if rank == 0:
            f2.download((c_host, w_host))            # D2H: the COMPLETE film
        t1 = time.perf_counter()
        if world == 1:
            f2.close()
        sc2.close()
        if i > 0 or args.workload.startswith("synth_10m"):
            e2e_times.append(t1 - t0)
    e2e_s = sum(e2e_times) / len(e2e_times)
    e2e_checksum = float(c_host.sum(dtype=np.float64))
    c_host = w_host = None
    c_pin.close(); w_pin.close()
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = n_samples_total / e2e_s / 1e6

    if rank != 0:
        barrier()
        for f in films:
            f.close()
        scene.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    barrier()

    # ---- rooflines (algorithmic bytes per unit: SURVEY.md 8d for rays, DESIGN.md 3 for the other kernels)
    peak, peak_src = read_peaks()
    # ncu evidence of the committed kernels (profiles/r02_ncu_summary_v27.json, written on the GPU box by profiles/tools/ncu_summary.py from
    # one `ncu --set full --clock-control none` capture of this command's frame): DRAM bytes, issue-slot and pipe utilisation of the
    # bounce-0 launch of every kernel. bench.py cannot run under ncu itself; the capture it quotes is named in the line.
    ncu_all = {}
    ncu_path = os.path.join(ROOT, "profiles", "r02_ncu_summary_v27.json")
    if os.path.exists(ncu_path) and args.workload in (WORKLOAD, "killeroo_path", "killeroo_path30"):
        try:
            ncu_all = json.load(open(ncu_path))
        except Exception:
            ncu_all = {}
    closest_bytes = 32.0 * nodes_per_closest + 48.0 * prims_per_closest + 40.0
    any_bytes = 32.0 * nodes_per_any + 48.0 * prims_per_any + 40.0
    v_all, v_first = class_rays[D.K_ACCUMULATE], first_vertices
    v_adv = class_rays[D.K_ADVANCE]
    rays_mis_closest_steps = mis_closest_frac * class_rays[D.K_TRACE_MIS]
    rays_closest = class_rays[D.K_TRACE_PATH] + rays_mis_closest_steps
    rays_any = class_rays[D.K_TRACE_SHADOW] + (class_rays[D.K_TRACE_MIS] - rays_mis_closest_steps)
    unit_bytes = {                                   # class -> (unit, algorithmic bytes moved by the class over the profiled steps, units)
        D.K_GEN: ("camera sample", 44.0 * class_rays[D.K_GEN], class_rays[D.K_GEN]),
        # ONE traversal launch per bounce: its path rays + the previous bounce's shadow / MIS rays
        D.K_TRACE_PATH: ("ray", closest_bytes * rays_closest + any_bytes * rays_any, rays_closest + rays_any),
        D.K_SHADE: ("path vertex", 250.0 * class_rays[D.K_SHADE], class_rays[D.K_SHADE]),
        # k_addlight: records 48 + verdicts 8 + (T 128 + L 256 | first vertex: L 128 written)
        D.K_ACCUMULATE: ("path vertex", (56.0 + 128.0) * v_first + (56.0 + 384.0) * max(v_all - v_first, 0), v_all),
        # k_advance: records 32 + g0/g3 32 + next ray 32 + queue 8 + (T 256 | first vertex: T 128 written)
        D.K_ADVANCE: ("path vertex", 104.0 * v_adv + 128.0 * min(v_first, v_adv) + 256.0 * max(v_adv - v_first, 0), v_adv),
        D.K_FILM: ("camera sample", 136.0 * class_rays[D.K_FILM] + 132.0 * class_rays[D.K_FILM] / max(rp.spp, 1), class_rays[D.K_FILM]),
    }
    total_class = sum(class_ms)
    # what bounds each kernel, read off the ncu capture (issue slots busy / lanes per instruction / DRAM and L2 throughput)
    LIMITER = {
        D.K_GEN: "instruction issue (80 % of the issue slots busy: hashing + camera arithmetic per sample)",
        D.K_TRACE_PATH: "instruction issue at 12-24 of 32 lanes per instruction; the BVH is served by L1/L2 (DRAM 1-10 % of peak even on the 10 M-triangle "
                        "scene: profiles/r02_ncu_trace_synth10m.csv), so the HBM fraction below is algorithmic bytes, not DRAM traffic",
        D.K_SHADE: "instruction latency (51 % of the issue slots busy at 16 warps/SM: 128 registers; 12.2 k SASS instructions of exact fp32/fp64 arithmetic, ~3.8 k executed per vertex)",
        D.K_ACCUMULATE: "HBM streaming (bounce 0: 68 % of the measured copy bandwidth in DRAM traffic with 67 % of the issue slots busy; later bounces 84 %)",
        D.K_ADVANCE: "HBM streaming (76-79 % of the measured copy bandwidth in DRAM traffic)",
        D.K_FILM: "HBM reads of the radiance rows + film atomics",
    }

    # the same in one word per kernel class: "issue" (instruction issue slots), "latency" (dependent instructions at low occupancy), "hbm"
    LIMITED_BY = {D.K_GEN: "issue", D.K_TRACE_PATH: "issue", D.K_SHADE: "latency", D.K_ACCUMULATE: "hbm", D.K_ADVANCE: "hbm", D.K_FILM: "hbm"}

    def roof(k):
        """Two views of a kernel class: over ALL its launches of the profiled frames (share of the step), and its bounce-0 launch alone
        - one well-defined launch (first vertices / camera rays) whose DRAM traffic the ncu capture gives for comparison."""
        ms = class_ms[k]
        unit, nbytes, units = unit_bytes[k]
        per_unit = nbytes / max(units, 1)
        ach = nbytes / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        f_ms, f_units = first_ms[k], first_units[k]
        if k == D.K_ACCUMULATE:
            f_bytes = (56.0 + 128.0) * f_units
        elif k == D.K_ADVANCE:
            f_bytes = (104.0 + 128.0) * f_units
        elif k == D.K_TRACE_PATH:
            f_bytes = closest_bytes * f_units
        else:
            f_bytes = per_unit * f_units
        f_ach = f_bytes / (f_ms / 1e3) / 1e9 if f_ms > 0 else 0.0
        nc = ncu_all.get(D.K_NAMES[k]) or {}
        return {"bound": "hbm", "limited_by": LIMITED_BY.get(k), "limiter": LIMITER.get(k), "kernel": D.K_KERNELS[k] + " (" + D.K_NAMES[k] + "), bounce-0 launch",
                "achieved": f_ach, "peak": peak, "unit": "GB/s", "frac": f_ach / peak,
                "traffic": nc.get("dram_bytes"), "peak_source": peak_src,
                "bytes_per_unit": f_bytes / max(f_units, 1), "unit_of_work": unit, "units_per_launch": f_units, "avg_launch_ms": f_ms,
                "all_launches": {"achieved": ach, "frac": ach / peak, "bytes_per_unit": per_unit, "units_per_launch": units / max(class_launches[k], 1),
                                 "avg_launch_ms": ms / max(class_launches[k], 1), "launches_per_step": class_launches[k] / prof_steps},
                "share_of_step": ms / total_class if total_class else None,
                "ncu": {q: nc.get(q) for q in ("ms", "dram_gbs", "l2_gbs", "l1_gbs", "dram_throughput_pct", "l2_throughput_pct", "l1_hit_pct", "l2_hit_pct", "issue_active_pct",
                                               "fma_pipe_pct", "alu_pipe_pct", "fp64_pipe_pct", "lanes_per_inst", "warps_active_pct", "registers")} if nc else None}
    dom = max((k for k in unit_bytes if class_launches[k]), key=lambda k: class_ms[k])
    roofline = roof(dom)
    roofline.update({
        "nodes_per_closest_ray": nodes_per_closest, "prim_tests_per_closest_ray": prims_per_closest,
        "nodes_per_shadow_ray": nodes_per_any, "prim_tests_per_shadow_ray": prims_per_any,
        "ncu_capture": ncu_all.get("_source"),
        "note": "achieved = algorithmic bytes of the kernel's bounce-0 launch (bytes_per_unit x units_per_launch, DESIGN.md 3) / its CUDA-event time in "
                "this run; traffic = DRAM bytes ncu measured for the same launch of the same command (profiles/r02_ncu_summary_v27.json); `limiter` says "
                "what actually bounds the kernel - only k_advance / k_addlight are HBM-bound; per-kernel lines in roofline_by_kernel"})
    roofline_by_kernel = {D.K_NAMES[k]: roof(k) for k in unit_bytes if class_launches[k]}
    rays_total = class_rays[D.K_TRACE_PATH] + class_rays[D.K_TRACE_MIS] + class_rays[D.K_TRACE_SHADOW]
    kernel_launch_counts = {D.K_NAMES[k]: class_launches[k] / prof_steps for k in range(D.K_CLASSES) if class_launches[k]}
    out = {
        "metric": "Msamples/sec (%d-band spectral path trace)" % args.bands, "value": value, "unit": "Msamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "reference scene file (killeroo-simple) lowered by the host side; no synthetic substitution" if args.workload == WORKLOAD
                else "reference scene file lowered by the host side (substitutions listed in config.workload)" if not args.workload.startswith("synth")
                else "synthetic scene built through the reference's own API and BVHAccel (oracle/synth_scene.cpp) and lowered by the host side" if args.workload in SYNTH_API
                else "synthetic scene written as .pbrt text, parsed, BVH-built and lowered by the reference's own code (oracle/make_golden.py)",
        "config": {"workload": workload_desc, "camera_samples_per_step": n_samples_total,
                   "parallelism": "image tile sets (32x32, round-robin) x%d, scene replicated; every rank's film kernel adds its samples straight into ONE film "
                                  "on rank 0 through a peer mapping (NVLink), a NCCL barrier ends the frame" % world,
                   "l2": "per-step wave state (>2 GB) and film are larger than L2; no explicit flush",
                   "pipelining": "a rank enqueues frames k+1..k+A (spt_render_begin; A = %d here: 2 where a rank's frame runs as one wave, else 1) before it waits for frame k "
                                 "(spt_render_end). N > 1: A + 2 film buffers on rank 0; frame k's end-of-frame barrier is waited for before "
                                 "frame k+A+1 is enqueued, rank 0 clears a buffer behind its frame's barrier and before it enters the next "
                                 "one; every frame's film is complete on rank 0 (every barrier has completed) inside the timed region" % ahead},
        "mrays_per_s": rays_total / prof_steps / (ms_per_step / 1e3) / 1e6 if world == 1 else None,
        "rays_per_sample": rays_total / prof_steps / (n_samples_total / world) if world else None,
        "rays_per_sample_reference": (rays_total + elided) / prof_steps / (n_samples_total / world) if world else None,
        "rays_note": "rays_per_sample counts rays traced; the reference also traces the BSDF-sampled MIS rays of EstimateDirect that cannot "
                     "reach the sampled sphere light (contribution exactly zero) - counted in rays_per_sample_reference, not traced here",
        "kernel_ms_per_step": {D.K_NAMES[k]: class_ms[k] / prof_steps for k in range(D.K_CLASSES) if class_launches[k]},
        "kernel_ms_note": "timed region: the frame's waves overlap on %d streams (ms_per_step); per-kernel times, rooflines and ray counts come "
                          "from %d more frames rendered with every wave on one stream (%.2f ms per frame), each kernel between its own CUDA events" % (lanes_used, prof_steps, serial_ms / prof_steps),
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": scene_bytes * world, "d2h_bytes_per_step": film_bytes,
                "ms_per_step": e2e_s * 1e3,
                "path": "spt_scene_create+spt_film_create+spt_render+spt_film_download, host buffers (film read into page-locked host memory)" if world == 1 else
                        "per rank spt_scene_create from host buffers + spt_render of its tile set into rank 0's film (cleared inside the timed region), "
                        "end-of-frame barrier, rank 0 spt_film_download of the COMPLETE film into page-locked host memory; max over ranks",
                "image_checksum": e2e_checksum},
        "gpu_launches": launches,
        "render_ms_max_rank": render_only_ms, "step_ms_max_rank": step_ms,
        "roofline": roofline,
        "roofline_by_kernel": roofline_by_kernel,
        "clocks": sampler.summary() if sampler else None,
        "image_checksum": image_sum,
        "film_check": film_check,
        "fast_mode": fast_mode,
        "kernel_launches_per_step": kernel_launch_counts,
    }
    if world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN) and args.workload in CPU_WORKLOADS and \
            os.path.exists(os.path.join(REF_SCENES, args.workload + ".pbrt")):
        try:
            v, cores, render_s, setup_s = cpu_reference_msamples(workload=args.workload)
            out["cpu_baseline"] = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "reference",
                                   "sample": "reference pbrt on the same frame at %d of %d spp (%.1f s render, %.2f s parse+BVH subtracted)" % (
                                       CPU_WORKLOADS[args.workload][3], CPU_WORKLOADS[args.workload][2], render_s, setup_s)}
        except (subprocess.CalledProcessError, OSError) as e:
            # the GPU measurement stands; a workload whose scene files did not travel (the texture directory of the
            # environment-map scenes is not shipped to the GPU box) has no CPU number here
            out["cpu_baseline"] = {"value": None, "unit": "Msamples/s", "kind": "reference", "unavailable": "reference run failed: %s" % e}
    else:
        out["cpu_baseline"] = None
    print_json(out)
    for f in films:
        f.close()
    scene.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
