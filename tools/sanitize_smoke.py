#!/usr/bin/env python3
"""Small frames through every kernel path (deterministic, stochastic exact + fast, photon emission /
build / gather) -- meant to be run under compute-sanitizer where the pool allows it (it is closed on
the round-1 pool; the same paths are covered by the parity tests):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
os.chdir(ROOT)
import bhraytracer_b200 as b  # noqa: E402
import gen_assets  # noqa: E402

gen_assets.ensure_assets()
r = b.Renderer(0)
for scene, w, h in (("scenes/det_mesh.xml", 48, 36), ("scenes/det_textured.xml", 40, 30), ("scenes/det_refract.xml", 40, 30)):
    sc = b.LoadScene(scene, w, h)
    r.upload(sc)
    r.render(mode=b.MODE_DET)
    r.render(mode=b.MODE_PATH, spp=2, seed=1)
    r.render(mode=b.MODE_PATH, spp=2, seed=1, exact_math=True)
sc = b.LoadScene("scenes/det_caustics.xml", 40, 30)
r.upload(sc)
stored, rays = r.photons_emit(3000, seed=2)
r.photons_build(stored)
r.render(mode=b.MODE_DET, use_photon_map=True, photon_k=50)
r.render(mode=b.MODE_PATH, spp=2, use_photon_map=True, photon_k=50)
r.close()
print("sanitize_smoke: done")
