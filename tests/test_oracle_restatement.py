"""oracle/bhrt_oracle.c -- the plain-C restatement of the reference's deterministic path for analytic
scenes (spheres + planes, point/direct lights, Blinn-Phong, hard shadows, Fresnel refraction with
absorption and total internal reflection, checker/bitmap textures, textured background/environment,
triangle meshes through the reference's own BVH topology).

 1. PIN: against the reference's own output (tests/golden, made by oracle/_ref/bhrt_ref --det) and,
    where /root/reference exists, against the reference run live on its own sphere scenes.
 2. USE: our device code (tests/emu on CPU, the kernels through the C ABI on the GPU) against the
    restatement at sizes and cameras no fixture covers."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import bhraytracer_b200 as b
import oracle_c
from aovio import read_aov
from conftest import EMU_BIN, REF_BIN, ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
REF_ROOT = "/root/reference/BHRayTracer"


def hard_lights(name, tmp_path, w, h):
    """A copy of scenes/<name>.xml with its area lights shrunk to points (the D-mode switches do not
    cover PointLight::Illuminate's random sample, PointLight.cpp:13-15) at w x h."""
    src = open(os.path.join(ROOT, "scenes", name + ".xml")).read()
    src, n = re.subn(r'<size value="[^"]*"\s*/>', "", src)
    assert n > 0
    src = re.sub(r'<width value="\d+"\s*/>', '<width value="%d"/>' % w, src)
    src = re.sub(r'<height value="\d+"\s*/>', '<height value="%d"/>' % h, src)
    path = str(tmp_path / (name + "_hard.xml"))
    open(path, "w").write(src)
    return path


def same(a, o):
    assert (a["hit_id"] == o["hit_id"]).all(), "primary hit ids differ"
    assert (a["z"].view(np.uint32) == o["z"].view(np.uint32)).all(), "primary hit distances differ in bits"
    assert np.abs(a["rgb8"].astype(np.int32) - o["rgb8"].astype(np.int32)).max() <= 1


def test_restatement_with_face_materials_is_pinned_by_the_patched_reference(built):
    """hInfo.mtlID = the usemtl group of the face (oracle patch Q-1 of build_ref.py): the restatement with that switch on
    reproduces the patched reference's fixture in every bit, and its default stays the unpatched reference's."""
    g = np.load(os.path.join(GOLD, "det_facemtl_facemtl_det.npz"))
    plain = np.load(os.path.join(GOLD, "det_facemtl_det.npz"))
    scene = b.LoadScene("scenes/det_facemtl.xml", int(g["width"]), int(g["height"]))
    for gold, flag in ((g, True), (plain, False)):
        o = oracle_c.render_det(scene, face_materials=flag)
        assert (o["hit_id"] == gold["hit_id"]).all() and (o["z"].view(np.uint32) == gold["z_bits"]).all()
        assert (o["rgb8"] == gold["rgb8"]).all()
        assert (o["rays_closest"], o["rays_shadow"]) == (int(gold["rays_closest"]), int(gold["rays_shadow"]))


@pytest.mark.parametrize("name", ["cfg1_spheres_plane", "det_refract", "det_textured", "det_mesh", "det_objmtl", "det_facemtl"])
def test_restatement_is_pinned_by_the_reference_golden(built, name):
    g = np.load(os.path.join(GOLD, name + "_det.npz"))
    scene = b.LoadScene("scenes/%s.xml" % name, int(g["width"]), int(g["height"]))
    assert oracle_c.in_scope(scene)
    o = oracle_c.render_det(scene)
    assert (o["hit_id"] == g["hit_id"]).all()
    assert (o["z"].view(np.uint32) == g["z_bits"]).all()
    assert (o["rgb8"] == g["rgb8"]).all()
    assert (o["rays_closest"], o["rays_shadow"]) == (int(g["rays_closest"]), int(g["rays_shadow"]))


@pytest.mark.skipif(not (os.path.isdir(REF_ROOT) and os.path.exists(REF_BIN)), reason="needs /root/reference and oracle/_ref")
@pytest.mark.parametrize("name", ["proj2", "proj3", "proj3Test", "proj4", "proj5", "proj7", "proj9", "proj10", "proj11_0", "proj11_1",
                                  "proj11_2", "proj12", "proj12_0", "proj12_backfaceTest", "proj12_metals", "proj13"])
def test_restatement_matches_the_reference_live_on_its_own_scenes(built, tmp_path, monkeypatch, name):
    src = open(os.path.join(REF_ROOT, "Resource", "Data", name + ".xml")).read()
    small = re.sub(r'<width value="\d+"\s*/>', '<width value="160"/>', src)
    small = re.sub(r'<height value="\d+"\s*/>', '<height value="120"/>', small)
    small = re.sub(r'<size value="[^"]*"\s*/>', "", small)      # area lights -> points (see hard_lights)
    path = str(tmp_path / (name + ".xml"))
    open(path, "w").write(small)
    info = json.loads(subprocess.check_output([REF_BIN, path, str(tmp_path / "ref"), "--det", "--threads", "2"], cwd=REF_ROOT,
                                              stderr=subprocess.DEVNULL).decode().splitlines()[-1])
    ref = read_aov(str(tmp_path / "ref.aov"))
    monkeypatch.chdir(REF_ROOT)                 # texture paths in the reference's scenes are relative to its root
    scene = b.LoadScene(path)
    if not oracle_c.in_scope(scene):
        pytest.skip("outside the restatement's scope")
    o = oracle_c.render_det(scene)
    same(o, ref)
    assert (o["rays_closest"], o["rays_shadow"]) == (info["rays_closest"], info["rays_shadow"])


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="needs oracle/_ref")
@pytest.mark.parametrize("name,w,h", [("cfg2_mesh_refract", 320, 180), ("cfg3_textured_arealight", 160, 90), ("cfg4_gi", 160, 90),
                                      ("cfg5_caustics", 160, 90)])
def test_restatement_matches_the_reference_live_on_the_config_scenes(built, tmp_path, name, w, h):
    path = hard_lights(name, tmp_path, w, h)
    info = json.loads(subprocess.check_output([REF_BIN, path, str(tmp_path / "ref"), "--det", "--threads", "2"], cwd=ROOT,
                                              stderr=subprocess.DEVNULL).decode().splitlines()[-1])
    ref = read_aov(str(tmp_path / "ref.aov"))
    o = oracle_c.render_det(b.LoadScene(path))
    assert (o["hit_id"] == ref["hit_id"]).all() and (o["z"].view(np.uint32) == ref["z"].view(np.uint32)).all()
    assert (o["rgb8"] == ref["rgb8"]).all()
    assert (o["rays_closest"], o["rays_shadow"]) == (info["rays_closest"], info["rays_shadow"])


@pytest.mark.parametrize("name,w,h", [("cfg1_spheres_plane", 97, 61), ("cfg1_spheres_plane", 33, 200), ("det_refract", 131, 77),
                                      ("det_textured", 90, 141), ("det_mesh", 75, 113), ("det_objmtl", 101, 67)])
def test_device_code_on_cpu_matches_the_restatement(built, tmp_path, name, w, h):
    scene = b.LoadScene("scenes/%s.xml" % name, w, h)
    o = oracle_c.render_det(scene)
    out = subprocess.check_output([EMU_BIN, "scenes/%s.xml" % name, str(tmp_path / "e"), "--det", "--size", str(w), str(h)],
                                  cwd=ROOT, stderr=subprocess.DEVNULL)
    info = json.loads(out.decode().splitlines()[-1])
    same(read_aov(str(tmp_path / "e.aov")), o)
    assert (info["rays_closest"], info["rays_shadow"]) == (o["rays_closest"], o["rays_shadow"])


@pytest.mark.parametrize("name,w,h", [("cfg2_mesh_refract", 240, 135), ("cfg3_textured_arealight", 120, 67)])
def test_device_code_on_cpu_matches_the_restatement_on_config_scenes(built, tmp_path, name, w, h):
    path = hard_lights(name, tmp_path, w, h)
    o = oracle_c.render_det(b.LoadScene(path))
    out = subprocess.check_output([EMU_BIN, path, str(tmp_path / "e"), "--det"], cwd=ROOT, stderr=subprocess.DEVNULL)
    info = json.loads(out.decode().splitlines()[-1])
    same(read_aov(str(tmp_path / "e.aov")), o)
    assert (info["rays_closest"], info["rays_shadow"]) == (o["rays_closest"], o["rays_shadow"])


def test_scope_check_refuses_what_the_restatement_does_not_cover(built):
    for name in ("cfg2_mesh_refract", "cfg3_textured_arealight"):
        scene = b.LoadScene(os.path.join("scenes", name + ".xml"), 32, 24)
        assert not oracle_c.in_scope(scene)
        with pytest.raises(ValueError):
            oracle_c.render_det(scene)


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h", [("cfg1_spheres_plane", 800, 600), ("cfg1_spheres_plane", 251, 97), ("det_refract", 403, 301),
                                      ("det_textured", 333, 222), ("det_mesh", 301, 211), ("det_objmtl", 320, 240)])
def test_kernels_match_the_restatement(built, name, w, h):
    scene = b.LoadScene("scenes/%s.xml" % name, w, h)
    o = oracle_c.render_det(scene)
    r = b.Renderer(0)
    try:
        r.upload(scene)
        out = r.render(mode=b.MODE_DET)
    finally:
        r.close()
    same(out, o)
    assert (out["rays_closest"], out["rays_shadow"]) == (o["rays_closest"], o["rays_shadow"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h", [("cfg2_mesh_refract", 640, 360), ("cfg4_gi", 320, 180), ("cfg5_caustics", 320, 180)])
def test_kernels_match_the_restatement_on_config_scenes(built, tmp_path, name, w, h):
    scene = b.LoadScene(hard_lights(name, tmp_path, w, h))
    o = oracle_c.render_det(scene)
    r = b.Renderer(0)
    try:
        r.upload(scene)
        out = r.render(mode=b.MODE_DET)
    finally:
        r.close()
    same(out, o)
    assert (out["rays_closest"], out["rays_shadow"]) == (o["rays_closest"], o["rays_shadow"])


def test_photon_gather_restatement_is_pinned_by_the_reference_golden(built):
    """EstimateIrradiance / LocatePhotons (cyPhotonMap.h:332-498) over the reference's own photon dump: the
    gather-lit deterministic frame of the reference, identical in every 8-bit channel."""
    g = np.load(os.path.join(GOLD, "det_caustics_pm.npz"))
    scene = b.LoadScene("scenes/det_caustics.xml", int(g["width"]), int(g["height"]))
    o = oracle_c.render_det_pm(scene, g["photons"])
    assert (o["hit_id"] == g["hit_id"]).all()
    assert (o["rgb8"] == g["rgb8"]).all()
