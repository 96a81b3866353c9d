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
    ("det_facemtl", 320, 240),
]
# (scene, width, height, fixture tag, extra arguments of the reference driver) more D-mode fixtures.  det_facemtl with
# --face-materials = the reference with oracle patch Q-1 (hInfo.mtlID set to the face's usemtl group): what
# BHRT_FLAG_FACE_MATERIALS must reproduce; the plain det_facemtl fixture above is the reference as it is.
DET_EXTRA = [
    ("det_facemtl", 320, 240, "facemtl", ["--face-materials"]),
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


def sized_scene_keep(path, w, h):
    import re
    src = open(path).read()
    src = re.sub(r'<width value="\d+"/>', '<width value="%d"/>' % w, src)
    src = re.sub(r'<height value="\d+"/>', '<height value="%d"/>' % h, src)
    p2 = path + ".small.xml"
    open(p2, "w").write(src)
    return p2


def run_ref(scene_path, prefix, *args):
    out = subprocess.check_output([REF, scene_path, prefix] + list(args), cwd=ROOT)
    return json.loads(out.decode().strip().splitlines()[-1]), read_aov(prefix + ".aov")


# (scene, width, height, photons) D-mode WITH the caustic photon map (bhrt_ref_pm): the fixture keeps
# the reference's own photons (its .dat: 24-byte records, power already scaled by 1/N) so that the
# gather can be compared deterministically, plus summary statistics of the emission.
PHOTON = [("det_caustics", 200, 150, 40000)]
REF_PM = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref_pm")
from photon_util import PHOTON_DT, photon_stats  # noqa: E402


# (scene, width, height, test spp, reference spp, photons) H-mode CONVERGENCE fixtures for the BASELINE configs that
# are benchmarked in the stochastic mode (area lights: cfg2 size 1, cfg3 / cfg4 size 5; cfg5 with the caustic photon
# map, bhrt_ref_pm).  TWO independent reference runs at the high spp are kept (uint16 fixed point, 1/16384), so the
# reference's own seed-to-seed RMSE AT THAT SPP is known and a bias far below the 16-spp noise is visible.
CONV = [
    ("cfg2_mesh_refract", 320, 180, 16, 1024, 0),
    ("cfg3_textured_arealight", 320, 180, 16, 1024, 0),
    ("cfg4_gi", 320, 180, 16, 1024, 0),
    ("cfg5_caustics", 240, 135, 16, 1024, 200000),
]


def fix16(lin):
    return np.clip(np.rint(lin.astype(np.float64) * 16384.0), 0, 65535).astype(np.uint16)


def make_conv(name, w, h, spp, ref_spp, photons, tmp):
    sp = sized_scene(name, w, h, tmp)
    binary = REF_PM if photons else REF
    extra = ["--photons", str(photons)] if photons else []

    def run(tag, n, seed):
        prefix = os.path.join(tmp, name + tag)
        out = subprocess.check_output([binary, sp, prefix, "--spp", str(n), "--seed", str(seed)] + extra, cwd=ROOT)
        return json.loads(out.decode().strip().splitlines()[-1]), read_aov(prefix + ".aov")
    try:
        ia, a = run("_ca", ref_spp, 21)
        ib, b = run("_cb", ref_spp, 22)
        i1, l1 = run("_c1", spp, 31)
        i2, l2 = run("_c2", spp, 32)
    finally:
        os.remove(sp)
    A, B = a["linear"].astype(np.float64), b["linear"].astype(np.float64)
    M = np.clip((A + B) / 2, 0, 1)
    pair = float(np.sqrt(np.mean((np.clip(A, 0, 1) - np.clip(B, 0, 1)) ** 2)))
    r1 = float(np.sqrt(np.mean((np.clip(l1["linear"], 0, 1) - M) ** 2)))
    r2 = float(np.sqrt(np.mean((np.clip(l2["linear"], 0, 1) - M) ** 2)))
    out = os.path.join(ROOT, "tests", "golden", "%s_conv.npz" % name)
    np.savez_compressed(out, mean_a=fix16(A), mean_b=fix16(B), ref_spp=ref_spp, spp=spp, width=w, height=h, photons=photons,
                        rmse_pair=pair, rmse_ref_a=r1, rmse_ref_b=r2,
                        rays_per_sample=(ia["rays_closest"] + ia["rays_shadow"]) / float(w * h * ref_spp),
                        closest_per_sample=ia["rays_closest"] / float(w * h * ref_spp))
    print("%-28s %dx%d H-mode convergence: pair RMSE at %d spp %.5f, %d-spp RMSE %.5f / %.5f -> %d bytes" % (
        name, w, h, ref_spp, pair, spp, r1, r2, os.path.getsize(out)))


def main():
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for name, w, h, spp, ref_spp, photons in CONV:
            if only and name + "_conv" not in only:
                continue
            make_conv(name, w, h, spp, ref_spp, photons, tmp)
    if only and all(o.endswith("_conv") for o in only):
        return
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
        for name, w, h, tag, extra in DET_EXTRA:
            if only and name + "_" + tag not in only:
                continue
            sp = sized_scene(name, w, h, tmp)
            try:
                info, a = run_ref(sp, os.path.join(tmp, name + tag), "--det", *extra)
            finally:
                os.remove(sp)
            out = os.path.join(ROOT, "tests", "golden", "%s_%s_det.npz" % (name, tag))
            np.savez_compressed(out, hit_id=a["hit_id"].astype(np.int16), rgb8=a["rgb8"], z_bits=a["z"].view(np.uint32),
                                rays_closest=info["rays_closest"], rays_shadow=info["rays_shadow"], width=w, height=h)
            print("%-28s %dx%d %s rays %d+%d -> %d bytes" % (name, w, h, " ".join(extra), info["rays_closest"], info["rays_shadow"], os.path.getsize(out)))
        for name, w, h, nph in PHOTON:
            if only and name not in only:
                continue
            sp = sized_scene(name, w, h, tmp)
            try:
                prefix = os.path.join(tmp, name + "_pm")
                out = subprocess.check_output([REF_PM, sp, prefix, "--det", "--photons", str(nph), "--seed", "3"], cwd=ROOT)
                info = json.loads(out.decode().strip().splitlines()[-1])
                a = read_aov(prefix + ".aov")
                rec = np.fromfile(prefix + ".photons.dat", dtype=PHOTON_DT)
                # a larger, independent emission only for the statistics
                prefix2 = os.path.join(tmp, name + "_pm2")
                subprocess.check_output([REF_PM, sized_scene_keep(sp, 16, 12), prefix2, "--det", "--photons", "300000", "--seed", "5"], cwd=ROOT)
                big = np.fromfile(prefix2 + ".photons.dat", dtype=PHOTON_DT)
            finally:
                os.remove(sp)
                if os.path.exists(sp + ".small.xml"):
                    os.remove(sp + ".small.xml")
            assert len(rec) == nph
            st = photon_stats(big, len(big))
            # the GLOBAL map (BuildPhotonMap, Main.cpp:251-295; oracle patch O-7 runs it and exits after its dump): statistics only
            sp3 = sized_scene(name, 16, 12, tmp)
            try:
                prefix3 = os.path.join(tmp, name + "_gpm")
                subprocess.check_output([REF_PM, sp3, prefix3, "--photons", "300000", "--seed", "5", "--global-photons"], cwd=ROOT)
                glob = np.fromfile(prefix3 + ".photons.dat", dtype=PHOTON_DT)
            finally:
                os.remove(sp3)
            gst = {"global_" + k: v for k, v in photon_stats(glob, len(glob)).items()}
            outp = os.path.join(ROOT, "tests", "golden", name + "_pm.npz")
            np.savez_compressed(outp, photons=rec.view(np.uint8).reshape(-1, 24), hit_id=a["hit_id"].astype(np.int16), rgb8=a["rgb8"],
                                width=w, height=h, k=1000, radius=0.5, stats_photons=len(big), global_stats_photons=len(glob), **st, **gst)
            print("%-28s %dx%d photon map %d photons -> %d bytes" % (name, w, h, nph, os.path.getsize(outp)))
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
