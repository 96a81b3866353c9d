import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import bhraytracer_b200 as b
os.chdir(ROOT)
r = b.Renderer(0)
sc = b.LoadScene("scenes/cfg5_caustics.xml", 1920, 1080); r.upload(sc)
stored, _ = r.photons_emit(10_000_000, 1); r.photons_build(stored)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
o = r.render(mode=b.MODE_PATH, spp=spp, seed=5, use_photon_map=True, want=("rgb8",))
print("cfg5 1920x1080 %d spp: rays %d+%d ms %.1f mean %.2f" % (spp, o["rays_closest"], o["rays_shadow"], o["kernel_ms"], o["rgb8"].mean()))
