import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import bhraytracer_b200 as b
os.chdir(ROOT)
r = b.Renderer(0)
for scene in ("scenes/cfg5_caustics.xml", "scenes/det_caustics.xml"):
    sc = b.LoadScene(scene, 240, 135); r.upload(sc)
    stored, _ = r.photons_emit(200000, 77); r.photons_build(stored)
    res = {}
    for name, kw in (("exact+pm", dict(exact_math=True, use_photon_map=True)), ("fast+pm", dict(use_photon_map=True)),
                     ("exact", dict(exact_math=True)), ("fast", dict()), ("fast+pm k100", dict(use_photon_map=True, photon_k=100)),
                     ("exact+pm k100", dict(exact_math=True, use_photon_map=True, photon_k=100))):
        o = r.render(mode=b.MODE_PATH, spp=16, seed=5, **kw)
        L = np.clip(o["linear"], 0, 1)
        res[name] = L
        print(scene, name, "mean %.4f nan %d rays %d+%d ms %.1f" % (L.mean(), int(np.isnan(o["linear"]).sum()), o["rays_closest"], o["rays_shadow"], o["kernel_ms"]))
    d = res["fast+pm"].mean(axis=2) - res["exact+pm"].mean(axis=2)
    H, W = d.shape
    for ty in range(6):
        print(" ".join("%+.2f" % d[ty*H//6:(ty+1)*H//6, tx*W//8:(tx+1)*W//8].mean() for tx in range(8)))
