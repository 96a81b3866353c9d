#!/usr/bin/env python3
"""bench.py -- the headline measurement of the B200 rendering core.

Metric (BASELINE.json): Mrays/s (and ms/frame) of the hot path -- camera rays, scene-graph/BVH
traversal, Blinn path-tracing estimator with shadow, refraction and GI rays -- on BASELINE
config 2: `scenes/cfg2_mesh_refract.xml`, 1920x1080, 16 spp, H-mode (HEAD's stochastic
estimator), GI depth 3.  A "ray" is one closest-hit query (reference recursive(), Main.cpp:389) or
one any-hit query (reference GenLight::Shadow, GenLight.cpp:10): the same definition the
instrumented reference counts (oracle/build_ref.py, patch O-5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one frame.  N > 1 (launched under torchrun, one rank per GPU): WEAK scaling by sample
range -- every rank renders its own 16 samples of every pixel, the frame has 16*N spp, the float4
accumulation buffers are all-reduced over NCCL inside the timed step and resolved on each rank.
`value` = rays traced by all ranks / device time (CUDA events on the one stream that carries the
render kernel, the all-reduce and the resolve; max over ranks), scene tables resident in HBM.
`e2e` = the same frame through the reference-facing C-ABI calls with HOST buffers: scene tables
uploaded H2D and the RGB8 image read back D2H (into pinned memory) inside the timed region.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/bhrt_ref, built from
/root/reference by oracle/build_ref.py; it travels to the GPU box as a prebuilt binary) on all host
threads, each step a bounded sample of the same workload (same scene, same spp, reduced
resolution).  Only this leg and the `cpu_baseline` leg execute anything under oracle/.
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

SCENE = "scenes/cfg2_mesh_refract.xml"
WIDTH, HEIGHT, SPP, GI_DEPTH = 1920, 1080, 16, 3
WORKLOAD = "cfg2_mesh_refract 1920x1080 16spp H-mode gi3 (BASELINE config 2)"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref")
REF_CNT_BIN = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref_cnt")
# bounded sample of the workload for the CPU legs: same scene and spp, 1/36 of the pixels
CPU_W, CPU_H = 320, 180
CPU_SAMPLE = "same scene, %dx%d, %d spp, gi %d (1/36 of the frame's pixels)" % (CPU_W, CPU_H, SPP, GI_DEPTH)


def sized_scene(w, h):
    """The workload scene with the camera resolution replaced (the reference reads it from the XML)."""
    src = open(os.path.join(ROOT, SCENE)).read()
    src = re.sub(r'<width value="\d+"/>', '<width value="%d"/>' % w, src)
    src = re.sub(r'<height value="\d+"/>', '<height value="%d"/>' % h, src)
    fd, path = tempfile.mkstemp(prefix="_tmp_bench_", suffix=".xml", dir=os.path.join(ROOT, "scenes"))
    os.write(fd, src.encode())
    os.close(fd)
    return path


def run_reference(binary, w, h, spp, threads=None, seed=1):
    path = sized_scene(w, h)
    try:
        with tempfile.TemporaryDirectory() as tmp:
            cmd = [binary, path, os.path.join(tmp, "out"), "--spp", str(spp), "--gi", str(GI_DEPTH), "--seed", str(seed)]
            if threads:
                cmd += ["--threads", str(threads)]
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


class _CudaArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


def scene_bytes(desc):
    """Bytes bhrt_upload_scene copies host -> device for this description."""
    import bhraytracer_b200 as b
    d = desc.contents
    sz = {"node": 104, "material": 72, "light": 32, "texmap": 88, "texture": 48, "mesh": 88}
    return (d.n_nodes * sz["node"] + d.n_materials * sz["material"] + d.n_lights * sz["light"] + d.n_texmaps * sz["texmap"] +
            d.n_textures * sz["texture"] + d.n_meshes * sz["mesh"] + d.texel_bytes + d.mesh_bytes + 31 * 8)


def bench_ours(args):
    import torch
    import bhraytracer_b200 as b

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
    scene = b.LoadScene(SCENE, WIDTH, HEIGHT)
    r = b.Renderer(local_rank)
    r.upload(scene)
    stream = torch.cuda.Stream()            # one stream carries the render kernel, the all-reduce and the resolve
    torch.cuda.set_stream(stream)
    r.set_stream(stream.cuda_stream)
    total_spp = SPP * world
    params = r.params(mode=b.MODE_PATH, spp=total_spp, gi_depth=GI_DEPTH, seed=1, sample_range=(rank * SPP, (rank + 1) * SPP))
    acc_ptr, acc_bytes = r.accum_buffer()
    accum = torch.as_tensor(_CudaArray(acc_ptr, acc_bytes // 4), device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB of L2

    def step():
        st = r.render_device(params)
        if dist is not None:
            dist.all_reduce(accum)
        r.resolve(total_spp, want=())
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
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = step()
        e1.record(stream)
        e1.synchronize()
        ms_steps.append(e0.elapsed_time(e1))
        ms_kernel.append(st["kernel_ms"])
        rays += st["rays_closest"] + st["rays_shadow"]
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    launches = r.launch_count() - launches0

    t = torch.tensor([sum(ms_steps), float(rays), sum(ms_kernel), float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, total_rays, kernel_ms_sum, total_launches = float(tmax[0]), float(tsum[1]), float(tmax[2]), int(tsum[3])
    else:
        total_ms, total_rays, kernel_ms_sum, total_launches = float(t[0]), float(t[1]), float(t[2]), int(t[3])

    # ---- end to end through the C ABI with host buffers (scene upload H2D + RGB8 readback D2H)
    pinned = torch.empty(WIDTH * HEIGHT * 3, dtype=torch.uint8).pin_memory()
    h2d = scene_bytes(scene.desc)
    d2h = WIDTH * HEIGHT * 3 + 16
    e2e_params = r.params(mode=b.MODE_PATH, spp=total_spp, gi_depth=GI_DEPTH, seed=1, sample_range=(rank * SPP, (rank + 1) * SPP))

    def e2e_step():
        r.upload(scene)
        if dist is None:
            return r.render_into(e2e_params, rgb8=pinned.data_ptr())
        st = r.render_device(e2e_params)
        dist.all_reduce(accum)
        o = b.Outputs()
        o.rgb8 = pinned.data_ptr() if rank == 0 else None
        r._check(r._L.bhrt_resolve(r._ctx, total_spp, C.byref(o)))
        return st

    e2e_step()
    barrier()
    e2e_rays = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = e2e_step()
        e2e_rays += st["rays_closest"] + st["rays_shadow"]
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, float(e2e_rays)], dtype=torch.float64, device="cuda")
    if dist is not None:
        a = te.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        s2 = te.clone()
        dist.all_reduce(s2, op=dist.ReduceOp.SUM)
        e2e_s, e2e_rays = float(a[0]), float(s2[1])

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = total_rays / (total_ms * 1e-3) / 1e6
        line = {
            "metric": "Mrays/s", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 3), "ms_per_frame": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "scene": SCENE, "width": WIDTH, "height": HEIGHT, "spp_per_gpu": SPP, "frame_spp": total_spp,
                       "mode": "H", "gi_depth": GI_DEPTH, "sharding": "sample ranges + NCCL all-reduce of float4 accumulation" if world > 1 else "none",
                       "rays_per_step": int(total_rays / args.steps), "l2": "512 MiB buffer written between timed frames (flush)",
                       "timing": "CUDA events on the stream carrying render kernel, all-reduce and resolve; max over ranks"},
            "wall_s_timed_region": round(wall_s, 3),
            "e2e": {"value": round(e2e_rays / e2e_s / 1e6, 2), "unit": "Mrays/s", "ms_per_frame": round(e2e_s / args.steps * 1e3, 3),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "how": "bhrt_upload_scene (host tables -> HBM) + bhrt_render into pinned host RGB8, wall clock, every step"},
            "gpu_launches": total_launches,
            "clocks": clocks,
        }
        # ---- roofline of the dominant kernel (wf_render), algorithmic bytes per SURVEY.md 8(d)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        bytes_per_ray, cnt_info = None, None
        if world == 1 and os.path.exists(REF_CNT_BIN) and not args.no_cpu:
            cnt_info = run_reference(REF_CNT_BIN, CPU_W, CPU_H, SPP)
            bts, rr = algorithmic_bytes(cnt_info, CPU_W, CPU_H)
            bytes_per_ray = bts / rr
            src = "instrumented reference (oracle/_ref/bhrt_ref_cnt) on " + CPU_SAMPLE
        if bytes_per_ray is None:
            fx = json.load(open(os.path.join(ROOT, "profiles", "algorithmic_bytes.json")))
            bytes_per_ray = fx["cfg2_mesh_refract"]["bytes_per_ray"]
            src = "committed fixture profiles/algorithmic_bytes.json (instrumented reference)"
        kernel_ms = kernel_ms_sum / args.steps
        rays_per_launch = total_rays / args.steps / world
        achieved = bytes_per_ray * rays_per_launch / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("wf_render_dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        line["roofline"] = {"bound": "hbm", "kernel": "wf_render (one persistent launch per frame)", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(achieved / peak, 4), "traffic": traffic, "peak_kind": peak_kind,
                            "algorithmic_bytes_per_ray": round(bytes_per_ray, 1), "bytes_source": src,
                            "kernel_ms_per_launch": round(kernel_ms, 3), "rays_per_launch": int(rays_per_launch),
                            "note": "algorithmic bytes are defined on the reference's data structures (SURVEY.md 8d); the scene (<= 20 MB) is L2-resident, the measured DRAM traffic is the kernel's own lane state (profiles/r1u_wf_render_ncu_summary.json)"}
        # ---- CPU baseline: the reference itself on this box's host cores, bounded sample
        if world == 1 and os.path.exists(REF_BIN) and not args.no_cpu:
            info = run_reference(REF_BIN, CPU_W, CPU_H, SPP)
            cr = info["rays_closest"] + info["rays_shadow"]
            line["cpu_baseline"] = {"value": round(cr / (info["ms"] * 1e-3) / 1e6, 3), "unit": "Mrays/s", "cores": info["threads"], "kind": "reference",
                                    "sample": CPU_SAMPLE + "; per-thread rand() interposed (glibc rand() lock, SURVEY.md 6)",
                                    "ms": info["ms"], "host_cpus": os.cpu_count()}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    r.close()
    if dist is not None:
        dist.destroy_process_group()


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REF_BIN):
        # the oracle is always built by __graft_entry__.build(); this only happens on a broken snapshot
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bhrt_ref missing (run __graft_entry__.build())"}))
        return
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        run_reference(REF_BIN, CPU_W // 2, CPU_H // 2, SPP, threads)
    ms, rays = 0.0, 0
    for k in range(args.steps):
        info = run_reference(REF_BIN, CPU_W, CPU_H, SPP, threads, seed=1 + k)
        ms += info["ms"]
        rays += info["rays_closest"] + info["rays_shadow"]
    value = rays / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": round(value, 3), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scene": SCENE, "sample": CPU_SAMPLE, "rays_per_step": rays // args.steps,
                   "timing": "omp_get_wtime() around the reference's BeginRender(); load and file output excluded"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": threads, "kind": "reference",
                         "sample": CPU_SAMPLE + "; per-thread rand() interposed (glibc rand() lock, SURVEY.md 6)"},
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs (profiling runs)")
    ap.add_argument("--size", default=None, help="WxH override for ncu runs (marks the line as a profile run; not a bench value)")
    ap.add_argument("--spp", type=int, default=None, help="samples per pixel override for ncu runs")
    args = ap.parse_args()
    global WIDTH, HEIGHT, SPP, WORKLOAD
    if args.size or args.spp:
        if args.size:
            WIDTH, HEIGHT = (int(x) for x in args.size.lower().split("x"))
        if args.spp:
            SPP = args.spp
        WORKLOAD = "PROFILE RUN (not a bench value): cfg2_mesh_refract %dx%d %dspp H-mode gi3" % (WIDTH, HEIGHT, SPP)
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
