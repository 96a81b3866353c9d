#!/usr/bin/env python3
"""bench.py -- the headline measurement of the B200 rendering core.

Metric (BASELINE.json): Mrays/s (and ms/frame) of the hot path -- camera rays, scene-graph/BVH
traversal, Blinn path-tracing estimator with shadow, refraction and GI rays (and, for config 5, photon
emission + kNN gather).  A "ray" is one closest-hit query (reference recursive(), Main.cpp:389) or one
any-hit query (reference GenLight::Shadow, GenLight.cpp:10): the definition the instrumented reference
counts (oracle/build_ref.py, patch O-5).  Photon paths are reported separately, never in `value`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--exact] [--impl ours|reference]

Default workload = BASELINE config 2 (`scenes/cfg2_mesh_refract.xml`, 1920x1080, 16 spp, H-mode, GI
depth 3), the configuration the metric is quoted on; `--config` selects any of the five.

A step = ONE FRAME of the configuration.  N > 1 (launched under torchrun, one rank per GPU): STRONG
scaling -- the same frame split over the ranks by the C++ multi-GPU host inside libbhrt.so
(include/bhrt_dist.h): sample ranges + NCCL all-reduce of the float4 accumulation buffers for the
stochastic configs, interleaved rows for the deterministic config 1, sharded emission + all-gather of
the 24-byte photon records + replicated build for config 5.  torch.distributed only carries the NCCL
rendezvous id and the final max-over-ranks of the timings.
`value` = rays of the frame / device time (CUDA events on the one stream that carries the render kernel,
the collectives and the resolve; max over ranks), scene tables resident in HBM.
`e2e` = the same frame through the reference-facing C-ABI calls with HOST buffers: scene tables uploaded
H2D and the RGB8 image read back D2H (into pinned memory) inside the timed region, wall clock.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/bhrt_ref[_pm], built from
/root/reference by oracle/build_ref.py; it travels to the GPU box as a prebuilt binary) on all host
threads, same scene / spp / mode; the frame is rendered at full resolution when K of them fit in the time
budget, else at the largest reduced resolution that does (`config.sample` says which).  Only this leg and
the `cpu_baseline` leg execute anything under oracle/.
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GI_DEPTH = 3
# BASELINE.json configs (SURVEY.md 8(d) restates them as concrete inputs)
CONFIGS = {
    1: dict(name="C1", scene="scenes/cfg1_spheres_plane.xml", w=800, h=600, spp=1, det=True, photons=0,
            what="cfg1_spheres_plane 800x600 1spp D-mode (BASELINE config 1)"),
    2: dict(name="C2", scene="scenes/cfg2_mesh_refract.xml", w=1920, h=1080, spp=16, det=False, photons=0,
            what="cfg2_mesh_refract 1920x1080 16spp H-mode gi3 (BASELINE config 2)"),
    3: dict(name="C3", scene="scenes/cfg3_textured_arealight.xml", w=1920, h=1080, spp=64, det=False, photons=0,
            what="cfg3_textured_arealight 1920x1080 64spp H-mode gi3 (BASELINE config 3)"),
    4: dict(name="C4", scene="scenes/cfg4_gi.xml", w=3840, h=2160, spp=256, det=False, photons=0,
            what="cfg4_gi 3840x2160 256spp H-mode gi3 (BASELINE config 4)"),
    5: dict(name="C5", scene="scenes/cfg5_caustics.xml", w=1920, h=1080, spp=32, det=False, photons=10_000_000,
            what="cfg5_caustics 1920x1080 32spp H-mode gi3, 10M caustic photons k=1000 r=0.5 (BASELINE config 5)"),
}
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref")
REF_PM_BIN = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref_pm")
REF_CNT_BIN = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref_cnt")
CPU_BUDGET_S = 150.0        # wall-clock budget of the whole reference arm; ~25 s for the cpu_baseline leg


def sized_scene(scene, w, h):
    """The workload scene with the camera resolution replaced (the reference reads it from the XML)."""
    src = open(os.path.join(ROOT, scene)).read()
    src = re.sub(r'<width value="\d+"/>', '<width value="%d"/>' % w, src)
    src = re.sub(r'<height value="\d+"/>', '<height value="%d"/>' % h, src)
    fd, path = tempfile.mkstemp(prefix="_tmp_bench_", suffix=".xml", dir=os.path.join(ROOT, "scenes"))
    os.write(fd, src.encode())
    os.close(fd)
    return path


def run_reference(cfg, w, h, spp, threads=None, seed=1, binary=None, photons=None, libc_rand=False):
    """One frame of the reference binary; returns its JSON line (ms = BeginRender only)."""
    path = sized_scene(cfg["scene"], w, h)
    photons = cfg["photons"] if photons is None else photons
    binary = binary or (REF_PM_BIN if photons else REF_BIN)
    try:
        with tempfile.TemporaryDirectory() as tmp:
            cmd = [binary, path, os.path.join(tmp, "out"), "--gi", str(GI_DEPTH), "--seed", str(seed)]
            cmd += ["--det"] if cfg["det"] else ["--spp", str(spp)]
            if threads:
                cmd += ["--threads", str(threads)]
            if photons:
                cmd += ["--photons", str(photons)]
            if libc_rand:
                cmd += ["--libc-rand"]
            out = subprocess.check_output(cmd, cwd=ROOT)
        return json.loads(out.decode().strip().splitlines()[-1])
    finally:
        os.remove(path)


def algorithmic_bytes(info, w, h):
    """bytes the reference's own data structures force per frame, from the
    instrumented reference's counters."""
    rays = info["rays_closest"] + info["rays_shadow"]
    b = (24 * rays + 64 * info["inst_tested"] + 28 * info["boxes_tested"] + 48 * info["tris_tested"] +
         96 * info["tris_accepted"] + 40 * info["rays_closest"] + 12 * info["texture_taps"] + 15 * w * h)
    return b, rays


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def scene_bytes(desc):
    """Bytes bhrt_upload_scene copies host -> device for this description."""
    d = desc.contents
    sz = {"node": 104, "material": 72, "light": 32, "texmap": 88, "texture": 48, "mesh": 88}
    return (d.n_nodes * sz["node"] + d.n_materials * sz["material"] + d.n_lights * sz["light"] + d.n_texmaps * sz["texmap"] +
            d.n_textures * sz["texture"] + d.n_meshes * sz["mesh"] + d.texel_bytes + d.mesh_bytes + 31 * 8)


def bench_ours(args):
    import torch
    import bhraytracer_b200 as b

    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched under torch.distributed.run with %d ranks" % (args.gpus, args.gpus))
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    os.chdir(ROOT)
    W, H, SPP = cfg["w"], cfg["h"], cfg["spp"]
    if args.size:
        W, H = (int(x) for x in args.size.lower().split("x"))
    if args.spp:
        SPP = args.spp
    mode = b.MODE_DET if cfg["det"] else b.MODE_PATH
    scene = b.LoadScene(cfg["scene"], W, H)
    r = b.Renderer(local_rank)
    r.upload(scene)
    stream = torch.cuda.Stream()            # one stream carries the render kernel, the collectives and the resolve
    torch.cuda.set_stream(stream)
    r.set_stream(stream.cuda_stream)
    dd = None
    if world > 1:
        # the C++ host's NCCL communicator: rank 0 makes the id, torch.distributed carries the 128 bytes
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(b.dist_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        torch.cuda.synchronize()
        dd = b.Dist(r, rank, world, bytes(idt.cpu().numpy().tobytes()))
    pm = cfg["photons"] > 0
    params = r.params(mode=mode, spp=SPP, gi_depth=GI_DEPTH, seed=1, use_photon_map=pm, exact_math=args.exact, wide_bvh=args.wide)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB of L2

    photon_info = {}

    def step():
        """One frame, everything on the device: (photon emission + exchange + build,) render, reduce, resolve."""
        if pm:
            if dd is not None:
                stored, prays = dd.photons(cfg["photons"], seed=1)
            else:
                stored, prays = r.photons_emit(cfg["photons"], seed=1)
                r.photons_build(stored)
            photon_info.update(stored=stored, photon_rays=prays)
        if dd is not None:
            return dd.render(params, dst=0, want=())
        st = r.render_device(params)
        r.resolve(1 if cfg["det"] else SPP, want=())
        return st

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = r.launch_count()
    ms_steps, ms_kernel, rays = [], [], 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                       # evict the scene from L2 between timed frames (untimed)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = step()
        e1.record(stream)
        e1.synchronize()
        ms_steps.append(e0.elapsed_time(e1))
        ms_kernel.append(st["kernel_ms"])
        rays += st["rays_closest"] + st["rays_shadow"]      # (multi-GPU: already the total over the ranks)
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    launches = r.launch_count() - launches0

    t = torch.tensor([sum(ms_steps), sum(ms_kernel), float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, kernel_ms_sum, total_launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
    else:
        total_ms, kernel_ms_sum, total_launches = float(t[0]), float(t[1]), int(t[2])
    total_rays = float(rays)

    # ---- end to end through the C ABI with host buffers (scene upload H2D + RGB8 readback D2H)
    pinned = torch.empty(W * H * 3, dtype=torch.uint8).pin_memory()
    h2d = scene_bytes(scene.desc)
    d2h = W * H * 3 + 16

    def e2e_step():
        r.upload(scene)
        if pm:
            if dd is not None:
                dd.photons(cfg["photons"], seed=1)
            else:
                stored, _ = r.photons_emit(cfg["photons"], seed=1)
                r.photons_build(stored)
        if dd is not None:
            return dd.render_into(params, dst=0, rgb8=pinned.data_ptr())
        return r.render_into(params, rgb8=pinned.data_ptr())

    e2e_n = args.steps if total_ms / args.steps < 3000 else max(1, min(args.steps, 2))     # long frames: fewer repeats
    e2e_step()
    barrier()
    e2e_rays = 0
    t0 = time.perf_counter()
    for _ in range(e2e_n):
        st = e2e_step()
        e2e_rays += st["rays_closest"] + st["rays_shadow"]
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = total_rays / (total_ms * 1e-3) / 1e6
        workload = cfg["what"] if not (args.size or args.spp) else "PROFILE RUN (not a bench value): %s at %dx%d %dspp" % (cfg["name"], W, H, SPP)
        build = "exact (IEEE, no FMA contraction)" if (cfg["det"] or args.exact) else "fast-math"
        sharding = "none"
        if world > 1:
            sharding = ("interleaved rows + NCCL sum all-reduce (frame byte-identical to 1 GPU)" if cfg["det"] else
                        "sample ranges + NCCL all-reduce of the float4 accumulation" + ("; photons: sharded emission + all-gather + replicated build" if pm else ""))
        line = {
            "metric": "Mrays/s", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 3), "ms_per_frame": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "scene": cfg["scene"], "width": W, "height": H, "frame_spp": 1 if cfg["det"] else SPP,
                       "mode": "D" if cfg["det"] else "H", "gi_depth": GI_DEPTH, "build": build,
                       "traversal": "8 lanes per ray, 8-wide BVH" if args.wide else "1 ray per thread, binary BVH",
                       "sharding": sharding, "host": "C++ (libbhrt.so: bhrt_dist_*, NCCL)" if world > 1 else "C ABI (libbhrt.so)",
                       "rays_per_step": int(total_rays / args.steps), "l2": "512 MiB buffer written between timed frames (flush)",
                       "timing": "CUDA events on the stream carrying render kernel, collectives and resolve; max over ranks"},
            "wall_s_timed_region": round(wall_s, 3),
            "e2e": {"value": round(e2e_rays / e2e_s / 1e6, 2), "unit": "Mrays/s", "ms_per_frame": round(e2e_s / e2e_n * 1e3, 3), "frames": e2e_n,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "how": "bhrt_upload_scene (host tables -> HBM) + bhrt_render into pinned host RGB8, wall clock, every step"},
            "gpu_launches": total_launches,
            "clocks": clocks,
        }
        if pm:
            line["photons"] = dict(photon_info, note="emission + exchange + build are inside every timed step; photon paths are not in `value`")
        # ---- roofline of the dominant kernel (wf_render), algorithmic bytes per SURVEY.md 8(d)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        key = os.path.splitext(os.path.basename(cfg["scene"]))[0]
        bytes_per_ray, src = None, None
        cw, ch = cpu_sample_size(cfg, 1 / 36.0)
        if world == 1 and os.path.exists(REF_CNT_BIN) and not args.no_cpu and not pm:
            cnt_info = run_reference(cfg, cw, ch, min(SPP, 16), binary=REF_CNT_BIN, photons=0)
            bts, rr = algorithmic_bytes(cnt_info, cw, ch)
            bytes_per_ray = bts / rr
            src = "instrumented reference (oracle/_ref/bhrt_ref_cnt) on the same scene at %dx%d" % (cw, ch)
        if bytes_per_ray is None:
            fx = json.load(open(os.path.join(ROOT, "profiles", "algorithmic_bytes.json")))
            bytes_per_ray = fx[key]["bytes_per_ray"]
            src = "committed fixture profiles/algorithmic_bytes.json (instrumented reference)" + (
                "; ray part only: the 24 B per photon examined by the gather are not in it" if pm else "")
        kernel_ms = kernel_ms_sum / args.steps
        rays_per_launch = total_rays / args.steps / world
        frame_bytes = bytes_per_ray * rays_per_launch
        gather_note = None
        if pm:
            # + 24 B per photon the reference's kd-tree search examines, per gather query (counted by bhrt_ref_pm_cnt with the full map)
            ga = json.load(open(os.path.join(ROOT, "profiles", "algorithmic_bytes.json")))[key].get("gather")
            if ga:
                queries = ga["gather_queries_per_camera_sample"] * W * H * SPP / world
                frame_bytes += queries * ga["photons_examined_per_query"] * ga["bytes_per_photon_examined"]
                src = src.replace("; ray part only: the 24 B per photon examined by the gather are not in it", "")
                gather_note = "%.2f gather queries per camera sample x %.0f photons examined x 24 B (reference kd-tree, 10 M photons) added to the ray part" % (
                    ga["gather_queries_per_camera_sample"], ga["photons_examined_per_query"])
        achieved = frame_bytes / (kernel_ms * 1e-3) / 1e9
        RK = "wf_render (one persistent launch per frame)" if not pm else (
            "photon-mapped frame = pm_emit + build + rounds of [wf_render until every lane is parked at a gather, pm_query_keys, radix sort, "
            "pm_gather_queries, pm_gather_heavy] (split gather; pm_gather_queries is the dominant kernel, profiles/r2_cfg5_pm_gather_queries_ncu_summary.json)")
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key, {}).get("wf_render_dram_bytes_per_launch")
        except (OSError, ValueError, AttributeError):
            pass
        line["roofline"] = {"bound": "hbm", "kernel": RK, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(achieved / peak, 4), "traffic": traffic, "peak_kind": peak_kind,
                            "algorithmic_bytes_per_ray": round(bytes_per_ray, 1), "bytes_source": src,
                            "kernel_ms_per_launch": round(kernel_ms, 3), "rays_per_launch": int(rays_per_launch), "gather_bytes": gather_note,
                            "note": "algorithmic bytes are defined on the reference's data structures (SURVEY.md 8d); the scene is L2-resident, the measured DRAM traffic is the kernel's own lane state (profiles/)"}
        line["cpu_baseline"] = cpu_baseline(cfg, SPP) if (world == 1 and os.path.exists(REF_BIN) and not args.no_cpu) else None
        print(json.dumps(line), flush=True)
    if dd is not None:
        dd.close()
    r.close()
    if dist is not None:
        dist.destroy_process_group()


def cpu_sample_size(cfg, frac):
    """Resolution with about `frac` of the frame's pixels, same aspect ratio (multiples of 8)."""
    s = frac ** 0.5
    return max(16, int(cfg["w"] * s) // 8 * 8), max(16, int(cfg["h"] * s) // 8 * 8)


def cpu_baseline(cfg, spp):
    """The reference itself on this box's host cores, on a bounded sample of the workload: (ii) fair multi-core
    (per-thread rand() interposed, all host threads) and (i) as shipped (libc rand(), 1 thread) -- BASELINE.md 3."""
    threads = os.cpu_count() or 1
    photons = min(cfg["photons"], 200_000)          # the reference emits ~20 k photons/s, serially
    cspp = min(spp, 16)
    w, h = cpu_sample_size(cfg, 1 / 36.0)
    info = run_reference(cfg, w, h, cspp, threads, photons=photons)
    cr = info["rays_closest"] + info["rays_shadow"]
    sample = "same scene and mode, %dx%d, %d spp, gi %d" % (w, h, 1 if cfg["det"] else cspp, GI_DEPTH) + (", %d photons" % photons if photons else "")
    out = {"value": round(cr / (info["ms"] * 1e-3) / 1e6, 3), "unit": "Mrays/s", "cores": info["threads"], "kind": "reference",
           "sample": sample + "; per-thread rand() interposed (glibc rand() lock, SURVEY.md 6); BeginRender() only",
           "ms": info["ms"], "host_cpus": threads}
    w1, h1 = cpu_sample_size(cfg, 1 / 144.0)
    one = run_reference(cfg, w1, h1, cspp, 1, photons=min(photons, 50_000), libc_rand=True)
    r1 = one["rays_closest"] + one["rays_shadow"]
    out["as_shipped"] = {"value": round(r1 / (one["ms"] * 1e-3) / 1e6, 3), "unit": "Mrays/s", "cores": 1, "rand": "libc rand()",
                         "sample": "same scene and mode, %dx%d, %d spp" % (w1, h1, 1 if cfg["det"] else cspp), "ms": one["ms"]}
    return out


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    binary = REF_PM_BIN if cfg["photons"] else REF_BIN
    if not os.path.exists(binary):
        # the oracle is always built by __graft_entry__.build(); this only happens on a broken snapshot
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/%s missing (run __graft_entry__.build())" % os.path.basename(binary)}))
        return
    threads = os.cpu_count() or 1
    spp = args.spp or cfg["spp"]
    photons = min(cfg["photons"], 200_000)
    # probe: a small frame gives the rate, from which the largest sample that fits the budget follows
    pw, ph = cpu_sample_size(cfg, 1 / 144.0)
    t0 = time.perf_counter()
    probe = run_reference(cfg, pw, ph, spp, threads, photons=min(photons, 20_000))
    probe_wall = time.perf_counter() - t0
    for _ in range(max(0, args.warmup - 1)):
        run_reference(cfg, pw, ph, min(spp, 4), threads, photons=min(photons, 20_000))
    per_pixel_s = max(probe["ms"] * 1e-3, 1e-3) / (pw * ph)
    overhead = max(0.0, probe_wall - probe["ms"] * 1e-3) + (photons / 20_000.0 if photons else 0.0)
    budget = max(10.0, CPU_BUDGET_S - probe_wall * max(1, args.warmup))
    frac = 1.0
    for frac in (1.0, 1 / 4.0, 1 / 9.0, 1 / 16.0, 1 / 36.0, 1 / 64.0, 1 / 144.0):
        w, h = (cfg["w"], cfg["h"]) if frac == 1.0 else cpu_sample_size(cfg, frac)
        if args.steps * (per_pixel_s * w * h + overhead) <= budget:
            break
    ms, rays = 0.0, 0
    for k in range(args.steps):
        info = run_reference(cfg, w, h, spp, threads, seed=1 + k, photons=photons)
        ms += info["ms"]
        rays += info["rays_closest"] + info["rays_shadow"]
    value = rays / (ms * 1e-3) / 1e6
    same = (w, h) == (cfg["w"], cfg["h"]) and photons == cfg["photons"]
    sample = ("the full frame" if (w, h) == (cfg["w"], cfg["h"]) else "same scene, spp and mode at %dx%d (1/%d of the frame's pixels: what fits %d steps in %d s)" % (
        w, h, round(cfg["w"] * cfg["h"] / float(w * h)), args.steps, int(CPU_BUDGET_S)))
    if cfg["photons"]:
        sample += "; %d photons (the reference emits them serially, ~20 k/s)" % photons
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": round(value, 3), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["what"], "scene": cfg["scene"], "sample": sample, "same_config": same, "width": w, "height": h,
                   "frame_spp": 1 if cfg["det"] else spp, "rays_per_step": rays // args.steps,
                   "timing": "omp_get_wtime() around the reference's BeginRender(); load and file output excluded"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": threads, "kind": "reference",
                         "sample": sample + "; per-thread rand() interposed (glibc rand() lock, SURVEY.md 6)"},
        "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (default 2: the one the metric is quoted on)")
    ap.add_argument("--exact", action="store_true", help="stochastic frames on the exact build (BHRT_FLAG_EXACT_MATH) instead of the fast-math one")
    ap.add_argument("--wide", action="store_true", help="fast build with the 8-lanes-per-ray traversal of the 8-wide BVH (BHRT_FLAG_WIDE_BVH)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs (profiling runs)")
    ap.add_argument("--size", default=None, help="WxH override for ncu runs (marks the line as a profile run; not a bench value)")
    ap.add_argument("--spp", type=int, default=None, help="samples per pixel override for ncu runs")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
