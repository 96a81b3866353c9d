#!/usr/bin/env python3
"""Random-scene parity fuzzer (CPU only, test infrastructure).

Generates random scenes in the reference's XML dialect -- nested transformed groups, spheres, planes,
OBJ meshes, Blinn materials with refraction / absorption / checker and bitmap textures, point /
direct / ambient lights -- and renders each in deterministic mode with

  * the C restatement of the reference  (oracle/bhrt_oracle.c),
  * our device code compiled for the CPU (build/emu_render, the source nvcc builds into the kernels),
  * the reference itself                (oracle/_ref/bhrt_ref, when it has been built),

then compares primary hit ids, z bits, 8-bit pixels and the ray counters.

usage: fuzz_parity.py [--seeds A:B] [--size WxH] [--keep DIR] [--no-ref] [--photon-map N] [--face-materials]

--face-materials: every scene gets OBJs whose usemtl groups become a MultiMtl (no material= attribute) and all three
renderers run with per-face sub-materials on (the reference with oracle patch Q-1, ours with BHRT_FLAG_FACE_MATERIALS).

--photon-map N: the caustic gather instead -- the reference emits N photons and renders the gather-lit
frame, the restatement and our device code render the same frame from its photon dump.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref")
EMU_BIN = os.path.join(ROOT, "build", "emu_render")
MESHES = ["scenes/assets/torus_small.obj", "scenes/assets/blob_small.obj", "scenes/assets/cube_mtl.obj"]
BITMAPS = ["scenes/assets/bricks.png", "scenes/assets/clouds.png"]
MAX_LEVELS = 4096       # BHRT_MAX_FRAMES_LIMIT (csrc/device/core.cuh)
FACE_MATERIALS = False  # --face-materials


def f(x):
    return "%.4g" % x


def transform(rng, indent, scale=(0.5, 4.0), spread=8.0, lift=True, dialect=1):
    out = []
    if rng.random() < 0.7:
        if rng.random() < 0.6:
            out.append('<scale value="%s"/>' % f(rng.uniform(*scale)))
        else:
            out.append('<scale x="%s" y="%s" z="%s"/>' % tuple(f(rng.uniform(*scale)) for _ in range(3)))
    for _ in range(rng.choice([0, 0, 1, 2])):
        if dialect < 2 or rng.random() < 0.7:
            axis = rng.choice(["x", "y", "z"])
            out.append('<rotate angle="%s" %s="1"/>' % (f(rng.uniform(-180, 180)), axis))
        else:       # arbitrary (unnormalised) axis
            out.append('<rotate angle="%s" x="%s" y="%s" z="%s"/>' % (f(rng.uniform(-180, 180)), f(rng.uniform(-1, 1)), f(rng.uniform(-1, 1)),
                                                                       f(rng.uniform(0.1, 1))))
    if rng.random() < 0.9:
        out.append('<translate x="%s" y="%s" z="%s"/>' % (f(rng.uniform(-spread, spread)), f(rng.uniform(-spread, spread)),
                                                          f(rng.uniform(0, spread) if lift else rng.uniform(-spread, spread))))
    return "".join(indent + t + "\n" for t in out)


def texture(rng, indent):
    kind = rng.choice(["checkerboard"] + BITMAPS)
    s = ' texture="%s">\n' % kind
    if kind == "checkerboard":
        s += indent + '  <color1 r="%s" g="%s" b="%s"/>\n' % tuple(f(rng.random()) for _ in range(3))
        s += indent + '  <color2 r="%s" g="%s" b="%s"/>\n' % tuple(f(rng.random()) for _ in range(3))
    if rng.random() < 0.8:
        s += indent + '  <scale value="%s"/>\n' % f(rng.choice([0.05, 0.1, 0.25, 0.5, 1.0, 2.0]))
    if rng.random() < 0.3:
        s += indent + '  <rotate angle="%s" z="1"/>\n' % f(rng.uniform(-90, 90))
    if rng.random() < 0.3:
        s += indent + '  <translate x="%s" y="%s"/>\n' % (f(rng.uniform(-1, 1)), f(rng.uniform(-1, 1)))
    return s


def material(rng, name):
    s = '    <material type="blinn" name="%s">\n' % name
    col = 'r="%s" g="%s" b="%s"' % tuple(f(rng.random()) for _ in range(3))
    if rng.random() < 0.35:
        s += "      <diffuse " + col + texture(rng, "      ") + "      </diffuse>\n"
    else:
        s += "      <diffuse " + col + "/>\n"
    s += '      <specular value="%s"/>\n' % f(rng.choice([0, 0.2, 0.5, 0.8, 1.0]))
    s += '      <glossiness value="%s"/>\n' % f(rng.choice([5, 20, 60, 200]))
    if rng.random() < 0.3:
        s += '      <refraction value="%s" index="%s"/>\n' % (f(rng.uniform(0.5, 1.0)), f(rng.choice([1.1, 1.33, 1.52, 2.0])))
        if rng.random() < 0.5:
            s += '      <absorption r="%s" g="%s" b="%s"/>\n' % tuple(f(rng.uniform(0, 0.3)) for _ in range(3))
    s += "    </material>\n"
    return s


def obj(rng, depth, mats, counter, meshes, dialect):
    ind = "    " + "  " * depth
    kinds = ["sphere", "sphere", "plane"] + (["obj"] if meshes else [])
    kind = rng.choice(kinds + ([None] if depth < 2 else []))
    counter[0] += 1
    name = "n%d" % counter[0]
    if kind is None:
        s = ind + '<object name="%s">\n' % name
    elif kind == "obj" and FACE_MATERIALS and not counter[1] and rng.random() < 0.7:
        # usemtl groups -> MultiMtl (xmlload.cpp:219-250).  Only ONE such node per scene: the loader creates the MultiMtl and
        # its node binding for the first instance of the file only, a second one is left without a material and the
        # reference dereferences null when it is hit (Main.cpp:161; our loader refuses such a scene)
        counter[1] = 1
        s = ind + '<object type="obj" name="scenes/assets/cube_mtl.obj">\n'
    elif kind == "obj":
        # (with --face-materials the cube only appears WITHOUT a material attribute: a file already loaded for a node that has
        # one is found in the object list and never gets its MultiMtl, xmlload.cpp:208-217)
        s = ind + '<object type="obj" name="%s" material="%s">\n' % (rng.choice(MESHES[:2] if FACE_MATERIALS else MESHES), rng.choice(mats))
    else:
        s = ind + '<object type="%s" name="%s" material="%s">\n' % (kind, name, rng.choice(mats))
    s += transform(rng, ind + "  ", scale=(3, 12) if kind == "plane" else (0.5, 4.0), dialect=dialect)
    if depth < 3:
        for _ in range(rng.choice([0, 0, 0, 1, 2] if kind else [1, 2, 3])):
            s += obj(rng, depth + 1, mats, counter, meshes, dialect)
    s += ind + "</object>\n"
    return s


def scene_xml(seed, w, h, meshes=True, dialect=None):
    """dialect 1: the generator the recorded seeds of tests/test_fuzz_parity.py were found with; dialect 2
    adds rotations about arbitrary axes and texture transforms on the background (default for
    seeds >= 2000)."""
    if dialect is None:
        dialect = 2 if seed >= 2000 else 1
    rng = random.Random(seed)
    mats = ["m%d" % i for i in range(rng.randint(2, 5))]
    s = "<xml>\n  <scene>\n"
    if rng.random() < 0.5:
        if dialect < 2:
            s += '    <background r="1" g="1" b="1" texture="%s"/>\n' % rng.choice(BITMAPS)
        else:
            s += '    <background r="1" g="1" b="1" texture="%s">\n' % rng.choice(BITMAPS)
            if rng.random() < 0.5:
                s += '      <scale x="%s" y="%s"/>\n' % (f(rng.choice([0.5, 1, 2])), f(rng.choice([0.5, 1, 2])))
            if rng.random() < 0.3:
                s += '      <translate x="%s"/>\n' % f(rng.uniform(-1, 1))
            s += '    </background>\n'
    else:
        s += '    <background r="%s" g="%s" b="%s"/>\n' % tuple(f(rng.random()) for _ in range(3))
    if rng.random() < 0.5:
        s += '    <environment value="%s" texture="%s"/>\n' % (f(rng.uniform(0.2, 1)), rng.choice(BITMAPS))
    else:
        s += '    <environment value="%s"/>\n' % f(rng.uniform(0, 1))
    # a ground plane most of the time, then random objects
    counter = [0, 0]       # objects so far; a material-less OBJ was placed (--face-materials)
    if rng.random() < 0.8:
        s += '    <object type="plane" name="ground" material="%s">\n      <scale value="30"/>\n    </object>\n' % mats[0]
    for _ in range(rng.randint(2, 6)):
        s += obj(rng, 0, mats, counter, meshes, dialect)
    for m in mats:
        s += material(rng, m)
    for i in range(rng.randint(1, 3)):
        kind = rng.choice(["point", "point", "direct", "ambient"])
        s += '    <light type="%s" name="l%d">\n' % (kind, i)
        if kind == "point":
            s += '      <intensity value="%s"/>\n      <position x="%s" y="%s" z="%s"/>\n' % (
                f(rng.uniform(50, 400)), f(rng.uniform(-12, 12)), f(rng.uniform(-12, 12)), f(rng.uniform(8, 25)))
        elif kind == "direct":
            s += '      <intensity value="%s"/>\n      <direction x="%s" y="%s" z="%s"/>\n' % (
                f(rng.uniform(0.2, 0.9)), f(rng.uniform(-1, 1)), f(rng.uniform(-1, 1)), f(rng.uniform(-1, -0.2)))
        else:
            s += '      <intensity value="%s"/>\n' % f(rng.uniform(0.05, 0.3))
        s += "    </light>\n"
    s += "  </scene>\n  <camera>\n"
    s += '    <position x="%s" y="%s" z="%s"/>\n' % (f(rng.uniform(-10, 10)), f(rng.uniform(-45, -25)), f(rng.uniform(3, 20)))
    s += '    <target x="%s" y="%s" z="%s"/>\n    <up x="0" y="0" z="1"/>\n' % (f(rng.uniform(-3, 3)), f(rng.uniform(-3, 3)), f(rng.uniform(0, 6)))
    s += '    <fov value="%s"/>\n    <width value="%d"/>\n    <height value="%d"/>\n  </camera>\n</xml>\n' % (f(rng.uniform(25, 60)), w, h)
    return s


def compare(a, b, what):
    """-> list of human-readable differences between two AOV dicts"""
    diffs = []
    n = int((a["hit_id"] != b["hit_id"]).sum())
    if n:
        diffs.append("%s: %d primary hit ids differ" % (what, n))
    n = int((a["z"].view(np.uint32) != b["z"].view(np.uint32)).sum())
    if n:
        diffs.append("%s: %d z values differ in bits" % (what, n))
    d = np.abs(a["rgb8"].astype(np.int32) - b["rgb8"].astype(np.int32))
    if d.max() > 0:
        diffs.append("%s: %d 8-bit channels differ (max %d)" % (what, int((d > 0).sum()), int(d.max())))
    if (a["rays_closest"], a["rays_shadow"]) != (b["rays_closest"], b["rays_shadow"]):
        diffs.append("%s: ray counts %s vs %s" % (what, (a["rays_closest"], a["rays_shadow"]), (b["rays_closest"], b["rays_shadow"])))
    return diffs


def run_binary(binary, path, out_prefix):
    from aovio import read_aov
    out = subprocess.check_output([binary, path, out_prefix, "--det"] + (["--threads", "2"] if binary == REF_BIN else []) + (["--face-materials"] if FACE_MATERIALS else []),
                                  cwd=ROOT, stderr=subprocess.DEVNULL)
    info = json.loads(out.decode().splitlines()[-1])
    r = read_aov(out_prefix + ".aov")
    r["rays_closest"], r["rays_shadow"] = info["rays_closest"], info["rays_shadow"]
    return r


def check_seed(seed, w, h, workdir, use_ref=True, meshes=True):
    """-> (list of differences, scene path)"""
    import bhraytracer_b200 as b
    import oracle_c
    path = os.path.join(workdir, "fuzz_%d.xml" % seed)
    open(path, "w").write(scene_xml(seed, w, h, meshes))
    cwd = os.getcwd()
    os.chdir(ROOT)                      # asset paths in the scenes are relative to the repository root
    try:
        scene = b.LoadScene(path)
        o = oracle_c.render_det(scene, face_materials=FACE_MATERIALS)
    finally:
        os.chdir(cwd)
    e = run_binary(EMU_BIN, path, os.path.join(workdir, "emu_%d" % seed))
    if o["max_depth"] > MAX_LEVELS:
        # the reference's Shade() recursion has no bound; the estimator's stack grows to MAX_LEVELS at most
        # (render.cu: launch_render): beyond it the frames are comparable only up to the cut-off tail
        return ["DEEP: the reference recurses %d levels (estimator stack: %d)" % (o["max_depth"], MAX_LEVELS)], path
    diffs = compare(e, o, "device code vs restatement")
    if use_ref and os.path.exists(REF_BIN):
        r = run_binary(REF_BIN, path, os.path.join(workdir, "ref_%d" % seed))
        diffs += compare(o, r, "restatement vs reference")
        diffs += compare(e, r, "device code vs reference")
    return diffs, path


REF_PM_BIN = os.path.join(ROOT, "oracle", "_ref", "bhrt_ref_pm")


def check_seed_photon_map(seed, w, h, workdir, n_photons=20000):
    """The caustic gather (MtlBlinn.cpp:329-342, cyPhotonMap.h:332-498): the REFERENCE emits the photons
    (its own serial loop, dumped as its .dat) and renders the gather-lit deterministic frame; our device
    code renders the same frame from the same photons.  -> (differences, path); needs oracle/_ref."""
    from aovio import read_aov
    path = os.path.join(workdir, "fuzz_pm_%d.xml" % seed)
    xml = scene_xml(seed, w, h)
    if 'type="point"' not in xml or "<refraction" not in xml:
        return ["SKIP: no point light or no refractive material"], path
    open(path, "w").write(xml)
    prefix = os.path.join(workdir, "refpm_%d" % seed)
    try:
        out = subprocess.check_output([REF_PM_BIN, path, prefix, "--det", "--photons", str(n_photons), "--seed", "3", "--threads", "2"],
                                      cwd=ROOT, stderr=subprocess.DEVNULL, timeout=120)
    except subprocess.TimeoutExpired:
        return ["SKIP: the reference's emission loop did not store %d photons in 120 s" % n_photons], path
    info = json.loads(out.decode().splitlines()[-1])
    r = read_aov(prefix + ".aov")
    r["rays_closest"], r["rays_shadow"] = info["rays_closest"], info["rays_shadow"]
    # The reference's kd-tree never visits the last photons of its balanced array: a node is expanded only
    # while index < halfStoredPhotons = N/2 - 1 (cyPhotonMap.h:257,427), so the children of nodes N/2-1 and
    # N/2 -- 3 photons for even N, 4 for odd -- are unreachable.  Our gather gets the reachable ones.
    ph = np.fromfile(prefix + ".photons.dat", dtype=np.uint8).reshape(-1, 24)
    reachable = 2 * (len(ph) // 2 - 2) + 1
    ph[:reachable].tofile(prefix + ".reachable.dat")
    eprefix = os.path.join(workdir, "emupm_%d" % seed)
    out = subprocess.check_output([EMU_BIN, path, eprefix, "--det", "--photons", prefix + ".reachable.dat", "1", "1000", "0.5"],
                                  cwd=ROOT, stderr=subprocess.DEVNULL)
    einfo = json.loads(out.decode().splitlines()[-1])
    e = read_aov(eprefix + ".aov")
    diffs = []
    # the restatement of the reference's gather (kd-tree heap order, LocatePhotons, its k-nearest heap) on ALL photons
    import bhraytracer_b200 as b
    import oracle_c
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        o = oracle_c.render_det_pm(b.LoadScene(path), ph)
    finally:
        os.chdir(cwd)
    if (o["rgb8"] != r["rgb8"]).any() or (o["hit_id"] != r["hit_id"]).any():
        diffs.append("restatement vs reference: %d 8-bit channels differ" % int((o["rgb8"] != r["rgb8"]).sum()))
    if (e["hit_id"] != r["hit_id"]).any():
        diffs.append("%d primary hit ids differ" % int((e["hit_id"] != r["hit_id"]).sum()))
    d = np.abs(e["rgb8"].astype(np.int32) - r["rgb8"].astype(np.int32))
    if d.max() > 1:         # the k nearest are summed in a different order (kd-tree heap vs grid): <= 1/255, not bit-exact
        diffs.append("%d 8-bit channels differ by more than 1 (max %d)" % (int((d > 1).sum()), int(d.max())))
    if einfo["rays_shadow"] != r["rays_shadow"]:         # (the reference's closest-hit counter includes its photon paths)
        diffs.append("shadow ray counts %d vs %d" % (einfo["rays_shadow"], r["rays_shadow"]))
    return diffs, path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="0:20")
    ap.add_argument("--size", default="96x64")
    ap.add_argument("--keep", default=None)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--no-meshes", action="store_true")
    ap.add_argument("--photon-map", type=int, default=0, metavar="N", help="gather-lit frames from N photons emitted by the reference")
    ap.add_argument("--face-materials", action="store_true", help="per-face sub-materials on in all three renderers (patch Q-1 / BHRT_FLAG_FACE_MATERIALS)")
    a = ap.parse_args()
    global FACE_MATERIALS
    FACE_MATERIALS = a.face_materials
    lo, hi = (int(x) for x in a.seeds.split(":"))
    w, h = (int(x) for x in a.size.split("x"))
    workdir = a.keep or tempfile.mkdtemp(prefix="bhrt_fuzz_")
    os.makedirs(workdir, exist_ok=True)
    bad = deep = 0
    for seed in range(lo, hi):
        try:
            if a.photon_map:
                diffs, path = check_seed_photon_map(seed, w, h, workdir, a.photon_map)
            else:
                diffs, path = check_seed(seed, w, h, workdir, not a.no_ref, not a.no_meshes)
        except subprocess.CalledProcessError as ex:
            print("seed %d: %s exited with %s" % (seed, os.path.basename(ex.cmd[0]), ex.returncode))
            bad += 1
            continue
        if diffs and diffs[0].startswith("SKIP"):
            deep += 1
            print("seed %d: %s" % (seed, diffs[0]))
        elif diffs and diffs[0].startswith("DEEP"):
            deep += 1
            print("seed %d (%s): %s" % (seed, path, diffs[0]))
        elif diffs:
            bad += 1
            print("seed %d (%s):\n  " % (seed, path) + "\n  ".join(diffs))
        else:
            print("seed %d ok" % seed)
    print("%d of %d seeds differ; %d skipped" % (bad, hi - lo, deep))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
