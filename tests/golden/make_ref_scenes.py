#!/usr/bin/env python3
"""Golden fixtures from the reference's OWN scene files (Resource/Data/proj*.xml), rendered by the reference itself
(oracle/_ref/bhrt_ref) -- run in the container that has /root/reference:

    python tests/golden/make_ref_scenes.py

For each scene: the XML as the reference ships it, with only the camera resolution replaced (160x120), is written
to tests/golden/ref_scenes/<name>.xml (test data), and <name>_det.npz holds the reference's D-mode outputs (hit ids,
z bits, RGB8, ray counts).  Two of them also get a stochastic fixture <name>_conv.npz like the BASELINE configs
(two 512-spp reference runs).  The scenes name their assets `Resource/Data/...` relative to the working directory:
both the reference (here) and the tests render from a directory whose Resource/Data holds THIS repository's
generated bricks.png / clouds.png (scenes/assets), so no reference asset is copied; OBJ meshes the reference tree does
not contain are dropped by both loaders alike."""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scenes"))
from aovio import read_aov  # noqa: E402

REF_ROOT = "/root/reference/BHRayTracer"
REF = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref")
OUT = os.path.join(ROOT, "tests", "golden", "ref_scenes")
SCENES = ["proj2", "proj3", "proj3Test", "proj4", "proj5", "proj7", "proj9", "proj10", "proj11_0", "proj11_2",
          "proj12", "proj12_0", "proj12_backfaceTest", "proj13"]
STOCH = ["proj12_backfaceTest", "proj13"]
W, H = 160, 120


def workdir(tmp):
    """A directory laid out like the reference's working directory, with this repository's generated textures."""
    import gen_assets
    gen_assets.ensure_assets()
    d = os.path.join(tmp, "Resource", "Data")
    os.makedirs(d)
    for f in ("bricks.png", "clouds.png"):
        shutil.copy(os.path.join(ROOT, "scenes", "assets", f), os.path.join(d, f))
    return tmp


def fix16(lin):
    return np.clip(np.rint(lin.astype(np.float64) * 16384.0), 0, 65535).astype(np.uint16)


def main():
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        wd = workdir(tmp)
        for name in SCENES:
            src = open(os.path.join(REF_ROOT, "Resource", "Data", name + ".xml")).read()
            small = re.sub(r'<width value="\d+"\s*/>', '<width value="%d"/>' % W, src)
            small = re.sub(r'<height value="\d+"\s*/>', '<height value="%d"/>' % H, small)
            scene = os.path.join(OUT, name + ".xml")
            open(scene, "w").write(small)
            prefix = os.path.join(tmp, name)
            out = subprocess.check_output([REF, scene, prefix, "--det", "--threads", "4", "--seed", "1"], cwd=wd, stderr=subprocess.DEVNULL)
            info = json.loads(out.decode().splitlines()[-1])
            a = read_aov(prefix + ".aov")
            area = re.search(r"<size\s", src) is not None
            np.savez_compressed(os.path.join(OUT, name + "_det.npz"), hit_id=a["hit_id"].astype(np.int16), rgb8=a["rgb8"], z_bits=a["z"].view(np.uint32),
                                rays_closest=info["rays_closest"], rays_shadow=info["rays_shadow"], width=W, height=H, area_light=area)
            print("%-22s D-mode rays %d+%d area_light=%s" % (name, info["rays_closest"], info["rays_shadow"], area))
        for name in STOCH:
            scene = os.path.join(OUT, name + ".xml")
            runs = []
            for seed, spp in ((21, 512), (22, 512), (31, 16), (32, 16)):
                prefix = os.path.join(tmp, "%s_s%d" % (name, seed))
                out = subprocess.check_output([REF, scene, prefix, "--spp", str(spp), "--seed", str(seed)], cwd=wd, stderr=subprocess.DEVNULL)
                runs.append((json.loads(out.decode().splitlines()[-1]), read_aov(prefix + ".aov")["linear"].astype(np.float64)))
            A, B = runs[0][1], runs[1][1]
            M = np.clip((A + B) / 2, 0, 1)
            pair = float(np.sqrt(np.mean((np.clip(A, 0, 1) - np.clip(B, 0, 1)) ** 2)))
            r16 = [float(np.sqrt(np.mean((np.clip(x[1], 0, 1) - M) ** 2))) for x in runs[2:]]
            ia = runs[0][0]
            np.savez_compressed(os.path.join(OUT, name + "_conv.npz"), mean_a=fix16(A), mean_b=fix16(B), ref_spp=512, spp=16, width=W, height=H, photons=0,
                                rmse_pair=pair, rmse_ref_a=r16[0], rmse_ref_b=r16[1],
                                rays_per_sample=(ia["rays_closest"] + ia["rays_shadow"]) / float(W * H * 512), closest_per_sample=ia["rays_closest"] / float(W * H * 512))
            print("%-22s H-mode pair RMSE at 512 spp %.5f, 16-spp RMSE %.5f / %.5f" % (name, pair, r16[0], r16[1]))


if __name__ == "__main__":
    main()
