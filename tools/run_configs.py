#!/usr/bin/env python3
"""Render the five BASELINE.json configurations on cuda:0 through the C ABI and print one JSON line
each: rays, device ms, Mrays/s.  `--scale f` shrinks resolution (and photon count) for quick runs."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import bhraytracer_b200 as b  # noqa: E402

CONFIGS = [
    # name, scene, width, height, spp, mode, photons
    ("C1", "scenes/cfg1_spheres_plane.xml", 800, 600, 1, b.MODE_DET, 0),
    ("C2", "scenes/cfg2_mesh_refract.xml", 1920, 1080, 16, b.MODE_PATH, 0),
    ("C3", "scenes/cfg3_textured_arealight.xml", 1920, 1080, 64, b.MODE_PATH, 0),
    ("C4", "scenes/cfg4_gi.xml", 3840, 2160, 256, b.MODE_PATH, 0),
    ("C5", "scenes/cfg5_caustics.xml", 1920, 1080, 32, b.MODE_PATH, 10_000_000),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="")
    ap.add_argument("--png", default="")
    args = ap.parse_args()
    os.chdir(ROOT)
    import gen_assets
    gen_assets.ensure_assets()
    r = b.Renderer(0)
    for name, scene, w, h, spp, mode, photons in CONFIGS:
        if args.only and name not in args.only.split(","):
            continue
        w2, h2 = max(16, int(w * args.scale)), max(16, int(h * args.scale))
        spp2 = spp if args.scale >= 1 or mode == b.MODE_DET else max(1, int(spp * args.scale))
        sc = b.LoadScene(scene, w2, h2)
        r.upload(sc)
        line = {"config": name, "scene": scene, "width": w2, "height": h2, "spp": spp2, "mode": "D" if mode == b.MODE_DET else "H"}
        if photons:
            n = max(1000, int(photons * args.scale * args.scale))
            t0 = time.perf_counter()
            stored, prays = r.photons_emit(n, seed=1)
            t1 = time.perf_counter()
            r.photons_build(stored)
            t2 = time.perf_counter()
            line.update(photons=stored, photon_rays=prays, emit_ms=round((t1 - t0) * 1e3, 2), build_ms=round((t2 - t1) * 1e3, 2))
        r.render(mode=mode, spp=min(spp2, 2), use_photon_map=bool(photons), want=())       # warm-up
        out = r.render(mode=mode, spp=spp2, use_photon_map=bool(photons), want=("rgb8",))
        rays = out["rays_closest"] + out["rays_shadow"]
        line.update(rays=rays, kernel_ms=round(out["kernel_ms"], 3), mrays_per_s=round(rays / out["kernel_ms"] / 1e3, 1),
                    rays_per_sample=round(rays / float(w2 * h2 * spp2), 2), mean_rgb8=round(float(out["rgb8"].mean()), 2))
        print(json.dumps(line), flush=True)
        if args.png:
            b.save_png(os.path.join(args.png, name + ".png"), out["rgb8"])
    r.close()


if __name__ == "__main__":
    main()
