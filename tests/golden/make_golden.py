#!/usr/bin/env python3
"""Generate the golden fixtures in this directory by running THE REFERENCE ITSELF
(oracle/_ref/bhrt_ref, built from /root/reference by oracle/build_ref.py) on the scenes of
scenes/.  Run from the repository root, in the container that has /root/reference:

    python tests/golden/make_golden.py

Each deterministic fixture <name>_det.npz holds the reference's D-mode outputs:
hit_id (int16, DFS node index, -1 = miss), rgb8 (uint8), z (float32 bit pattern as uint32),
plus the exact ray counts.  Stochastic fixtures <name>_h<spp>.npz hold a high-spp reference
mean (float16 linear RGB) and the reference's own seed-to-seed RMSE at the test spp.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aovio import read_aov  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref")

# (scene, width, height) rendered in D-mode
DET = [
    ("cfg1_spheres_plane", 800, 600),
    ("det_refract", 400, 300),
    ("det_textured", 400, 300),
    ("det_mesh", 400, 300),
    ("det_objmtl", 320, 240),
]
# (scene, width, height, test spp, reference spp) rendered in H-mode
STOCH = [
    ("cfg1_spheres_plane", 200, 150, 16, 1024),
    ("det_refract", 200, 150, 16, 1024),
    ("det_textured", 200, 150, 16, 1024),
    ("det_mesh", 160, 120, 16, 512),
]


def sized_scene(name, w, h, tmp):
    """Scene file with the camera resolution replaced (the reference reads it from the XML)."""
    src = open(os.path.join(ROOT, "scenes", name + ".xml")).read()
    import re
    src = re.sub(r'<width value="\d+"/>', '<width value="%d"/>' % w, src)
    src = re.sub(r'<height value="\d+"/>', '<height value="%d"/>' % h, src)
    path = os.path.join(ROOT, "scenes", "_tmp_%s_%dx%d.xml" % (name, w, h))
    open(path, "w").write(src)
    return path


def run_ref(scene_path, prefix, *args):
    out = subprocess.check_output([REF, scene_path, prefix] + list(args), cwd=ROOT)
    return json.loads(out.decode().strip().splitlines()[-1]), read_aov(prefix + ".aov")


def main():
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for name, w, h in DET:
            if only and name not in only:
                continue
            sp = sized_scene(name, w, h, tmp)
            try:
                info, a = run_ref(sp, os.path.join(tmp, name), "--det")
                info1, a1 = run_ref(sp, os.path.join(tmp, name + "_t1"), "--det", "--threads", "1")
            finally:
                os.remove(sp)
            assert (a["rgb8"] == a1["rgb8"]).all() and (a["hit_id"] == a1["hit_id"]).all(), name + ": D-mode not deterministic"
            out = os.path.join(ROOT, "tests", "golden", name + "_det.npz")
            np.savez_compressed(out, hit_id=a["hit_id"].astype(np.int16), rgb8=a["rgb8"], z_bits=a["z"].view(np.uint32),
                                rays_closest=info["rays_closest"], rays_shadow=info["rays_shadow"], width=w, height=h)
            print("%-28s %dx%d rays %d+%d -> %d bytes" % (name, w, h, info["rays_closest"], info["rays_shadow"], os.path.getsize(out)))
        for name, w, h, spp, ref_spp in STOCH:
            if only and name not in only:
                continue
            sp = sized_scene(name, w, h, tmp)
            try:
                _, hi = run_ref(sp, os.path.join(tmp, name + "_hi"), "--spp", str(ref_spp), "--seed", "7")
                ia, a = run_ref(sp, os.path.join(tmp, name + "_a"), "--spp", str(spp), "--seed", "11")
                ib, b = run_ref(sp, os.path.join(tmp, name + "_b"), "--spp", str(spp), "--seed", "12")
            finally:
                os.remove(sp)
            ref = np.clip(hi["linear"].astype(np.float64), 0, 1)
            rm_a = float(np.sqrt(np.mean((np.clip(a["linear"], 0, 1) - ref) ** 2)))
            rm_b = float(np.sqrt(np.mean((np.clip(b["linear"], 0, 1) - ref) ** 2)))
            out = os.path.join(ROOT, "tests", "golden", "%s_h%d.npz" % (name, spp))
            np.savez_compressed(out, mean=hi["linear"].astype(np.float16), ref_spp=ref_spp, spp=spp, width=w, height=h,
                                rmse_ref_a=rm_a, rmse_ref_b=rm_b,
                                rays_per_sample=(ia["rays_closest"] + ia["rays_shadow"]) / float(w * h * spp))
            print("%-28s %dx%d H-mode spp %d vs %d: reference RMSE %.5f / %.5f -> %d bytes" % (name, w, h, spp, ref_spp, rm_a, rm_b, os.path.getsize(out)))


if __name__ == "__main__":
    main()
