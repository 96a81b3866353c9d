"""CPU tier: the device code of the renderer (csrc/device/core.cuh -- the source nvcc compiles into
the sm_100a kernels) compiled with g++ by tests/emu and compared with the reference's own outputs
(tests/golden, generated from oracle/_ref by tests/golden/make_golden.py).  The GPU tier
(test_gpu_parity.py) runs the same comparisons through the C ABI on the CUDA build."""
import json
import os
import subprocess

import numpy as np
import pytest

from aovio import read_aov
from conftest import EMU_BIN, REF_BIN, ROOT, have_ref

GOLD = os.path.join(ROOT, "tests", "golden")


def emu(scene, prefix, *args):
    out = subprocess.check_output([EMU_BIN, scene, prefix] + [str(a) for a in args], cwd=ROOT, stderr=subprocess.DEVNULL)
    return json.loads(out.decode().splitlines()[-1]), read_aov(prefix + ".aov")


@pytest.mark.parametrize("name", ["cfg1_spheres_plane", "det_refract", "det_textured", "det_mesh", "det_objmtl", "det_facemtl"])
def test_det_mode_matches_reference_golden(built, tmp_path, name):
    g = np.load(os.path.join(GOLD, name + "_det.npz"))
    w, h = int(g["width"]), int(g["height"])
    info, a = emu(os.path.join("scenes", name + ".xml"), str(tmp_path / name), "--det", "--size", w, h)
    assert (a["hit_id"] == g["hit_id"]).all(), "primary hit ids differ from the reference"
    assert (a["z"].view(np.uint32) == g["z_bits"]).all()
    d = np.abs(a["rgb8"].astype(np.int32) - g["rgb8"].astype(np.int32))
    assert d.max() <= 1
    assert info["rays_closest"] == int(g["rays_closest"]) and info["rays_shadow"] == int(g["rays_shadow"])


def test_face_materials_flag_matches_the_patched_reference(built, tmp_path):
    """BHRT_FLAG_FACE_MATERIALS (bhrt.h; SURVEY.md H4 quirk switch): every face shaded with the sub-material of its usemtl
    group.  The fixture is the reference with ONE assignment added -- hInfo.mtlID = GetMaterialIndex(faceID) in
    TriObj::IntersectTriangle (oracle/build_ref.py patch Q-1), which is what its own MultiMtl::Shade expects
    (materials.h:71).  Without the flag the frame is the unpatched reference's (det_facemtl in the test above), and the
    two fixtures differ in thousands of pixels."""
    g = np.load(os.path.join(GOLD, "det_facemtl_facemtl_det.npz"))
    plain = np.load(os.path.join(GOLD, "det_facemtl_det.npz"))
    assert int((g["rgb8"] != plain["rgb8"]).any(axis=2).sum()) > 1000
    w, h = int(g["width"]), int(g["height"])
    info, a = emu(os.path.join("scenes", "det_facemtl.xml"), str(tmp_path / "fm"), "--det", "--size", w, h, "--face-materials")
    assert (a["hit_id"] == g["hit_id"]).all() and (a["z"].view(np.uint32) == g["z_bits"]).all()
    assert np.abs(a["rgb8"].astype(np.int32) - g["rgb8"].astype(np.int32)).max() <= 1
    assert info["rays_closest"] == int(g["rays_closest"]) and info["rays_shadow"] == int(g["rays_shadow"])


@pytest.mark.parametrize("name", ["cfg1_spheres_plane", "det_refract", "det_textured", "det_mesh"])
def test_path_mode_rmse_vs_reference_golden(built, tmp_path, name):
    g = np.load(os.path.join(GOLD, name + "_h16.npz"))
    w, h, spp = int(g["width"]), int(g["height"]), int(g["spp"])
    info, a = emu(os.path.join("scenes", name + ".xml"), str(tmp_path / name), "--spp", spp, "--seed", 5, "--size", w, h)
    ref = np.clip(g["mean"].astype(np.float64), 0, 1)
    rmse = float(np.sqrt(np.mean((np.clip(a["linear"], 0, 1) - ref) ** 2)))
    bound = 1.25 * max(float(g["rmse_ref_a"]), float(g["rmse_ref_b"]))
    assert rmse <= bound, "RMSE %.5f > %.5f" % (rmse, bound)
    assert abs(np.clip(a["linear"], 0, 1).mean() - ref.mean()) < 0.01
    rps = (info["rays_closest"] + info["rays_shadow"]) / float(w * h * spp)
    assert abs(rps - float(g["rays_per_sample"])) / float(g["rays_per_sample"]) < 0.05


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_golden_is_reproducible_from_the_reference(tmp_path):
    """The committed fixture really is what the reference outputs (pins the fixture, not us)."""
    g = np.load(os.path.join(GOLD, "det_refract_det.npz"))
    import re
    src = open(os.path.join(ROOT, "scenes", "det_refract.xml")).read()
    src = re.sub(r'<width value="\d+"/>', '<width value="%d"/>' % int(g["width"]), src)
    src = re.sub(r'<height value="\d+"/>', '<height value="%d"/>' % int(g["height"]), src)
    p = os.path.join(ROOT, "scenes", "_tmp_pin.xml")
    open(p, "w").write(src)
    try:
        subprocess.check_output([REF_BIN, p, str(tmp_path / "r"), "--det", "--threads", "2"], cwd=ROOT)
    finally:
        os.remove(p)
    a = read_aov(str(tmp_path / "r.aov"))
    assert (a["hit_id"] == g["hit_id"]).all() and (a["rgb8"] == g["rgb8"]).all()
