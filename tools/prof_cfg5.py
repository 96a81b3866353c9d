"""Profiling driver for the photon-mapped frame (BASELINE config 5 at a reduced frame): one wf_render launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import bhraytracer_b200 as b
os.chdir(ROOT)
w, h, spp = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (960, 540, 2)))
r = b.Renderer(0)
sc = b.LoadScene("scenes/cfg5_caustics.xml", w, h); r.upload(sc)
stored, _ = r.photons_emit(10_000_000, 1); r.photons_build(stored)
o = r.render(mode=b.MODE_PATH, spp=spp, seed=5, use_photon_map=True, want=("rgb8",))
print("cfg5 %dx%d %d spp: rays %d+%d ms %.1f mean %.2f" % (w, h, spp, o["rays_closest"], o["rays_shadow"], o["kernel_ms"], o["rgb8"].mean()))
