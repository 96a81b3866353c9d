"""The two schedules of the caustic gather side by side on BASELINE config 5's scene: inside wf_render (BHRT_GATHER=inkernel)
and split (device-wide, queries sorted by photon cell; the default).  Same per-query routine, same samples: the frames must agree
(to the run-to-run noise of the float sums); prints the kernel time of each.  usage: gather_modes.py [width height spp [photons]]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bhraytracer_b200 as b
os.chdir(ROOT)
w, h, spp = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (960, 540, 2)))
n_ph = int(sys.argv[4]) if len(sys.argv) > 4 else 10_000_000
frames = {}
for mode in ("inkernel", "split"):
    os.environ["BHRT_GATHER"] = mode
    r = b.Renderer(0)
    sc = b.LoadScene("scenes/cfg5_caustics.xml", w, h); r.upload(sc)
    stored, _ = r.photons_emit(n_ph, 1); r.photons_build(stored)
    prev = None
    for rep in range(2):
        o = r.render(mode=b.MODE_PATH, spp=spp, seed=5, use_photon_map=True, want=("rgb8", "linear"))
        if prev is not None:    # (the candidates of a selection's crossing bin are ranked in the order shared-memory atomics admitted them: float sums may differ in the last bit from run to run)
            print("%-8s run to run: max |linear difference| %.3g, rgb8 differing channels %d" % (mode, np.abs(o["linear"] - prev["linear"]).max(), int((o["rgb8"] != prev["rgb8"]).sum())))
        prev = o
    frames[mode] = o
    print("%-8s %dx%d %d spp: rays %d+%d kernel %.1f ms, mean %.3f" % (mode, w, h, spp, o["rays_closest"], o["rays_shadow"], o["kernel_ms"], o["rgb8"].mean()), flush=True)
    r.close()
a, c = frames["inkernel"], frames["split"]
d = np.abs(a["linear"] - c["linear"]).max()
print("max |linear difference| %.3g, rgb8 differing channels %d, ray counts equal %s" % (d, int((a["rgb8"] != c["rgb8"]).sum()),
      (a["rays_closest"], a["rays_shadow"]) == (c["rays_closest"], c["rays_shadow"])))
assert np.abs(a["linear"] - c["linear"]).mean() < 1e-5, "the two gather schedules must produce the same frame"
