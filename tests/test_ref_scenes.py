"""north_star: "every repo scene renders within tolerance of the CPU reference".  The reference's OWN scene files
(Resource/Data/proj*.xml, committed unchanged apart from the camera resolution under tests/golden/ref_scenes/) against
what the reference itself rendered from them (tests/golden/make_ref_scenes.py): on the GPU through the C ABI
(`-m gpu`), and through the CPU build of the same device code (tests/emu) so the CPU tier covers them wherever it
runs.  Deterministic mode: bit-exact primary hit ids and z, <= 1/255 per 8-bit channel, identical ray counts.
Scenes whose point lights have a <size> keep one random number in D-mode (the shadow ray of PointLight::Illuminate
goes to a random point of the light's disk, PointLight.cpp:13-14): hit ids, z and closest-hit counts must still be
exact, the image may only differ in penumbrae.  Two scenes are also checked in the stochastic mode against two
512-spp reference runs (same bars as tests/test_gpu_convergence.py)."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from aovio import read_aov
from conftest import EMU_BIN, ROOT

FIX = os.path.join(ROOT, "tests", "golden", "ref_scenes")
SCENES = ["proj2", "proj3", "proj3Test", "proj4", "proj5", "proj7", "proj9", "proj10", "proj11_0", "proj11_2",
          "proj12", "proj12_0", "proj12_backfaceTest", "proj13"]
STOCH = ["proj12_backfaceTest", "proj13"]


@pytest.fixture(scope="module")
def workdir(tmp_path_factory, assets):
    """A directory laid out like the reference's working directory (the scenes name `Resource/Data/...`), holding this
    repository's generated textures -- the same ones the fixtures were rendered with."""
    d = tmp_path_factory.mktemp("refwd")
    os.makedirs(d / "Resource" / "Data")
    for f in ("bricks.png", "clouds.png"):
        shutil.copy(os.path.join(ROOT, "scenes", "assets", f), d / "Resource" / "Data" / f)
    return str(d)


def check_det(out, g, name):
    assert (out["hit_id"] == g["hit_id"].astype(np.int32)).all(), "%s: primary hit ids differ on %d pixels" % (name, int((out["hit_id"] != g["hit_id"]).sum()))
    assert (out["z"].view(np.uint32) == g["z_bits"]).all(), name + ": primary hit distances differ in bits"
    assert out["rays_closest"] == int(g["rays_closest"]), name
    d = np.abs(out["rgb8"].astype(np.int32) - g["rgb8"].astype(np.int32))
    if not bool(g["area_light"]):
        assert d.max() <= 1, "%s: 8-bit error %d on %d pixels" % (name, d.max(), int((d.max(axis=2) > 1).sum()))
        assert out["rays_shadow"] == int(g["rays_shadow"]), name
    else:
        assert (d.max(axis=2) > 1).mean() < 0.15, name


@pytest.mark.parametrize("name", SCENES)
def test_reference_scene_on_the_cpu_build_of_the_device_code(built, workdir, tmp_path, name):
    g = np.load(os.path.join(FIX, name + "_det.npz"))
    out = subprocess.check_output([EMU_BIN, os.path.join(FIX, name + ".xml"), str(tmp_path / "emu"), "--det"], cwd=workdir, stderr=subprocess.DEVNULL)
    info = json.loads(out.decode().splitlines()[-1])
    a = read_aov(str(tmp_path / "emu.aov"))
    a.update(rays_closest=info["rays_closest"], rays_shadow=info["rays_shadow"])
    check_det(a, g, name)


@pytest.fixture(scope="module")
def renderer(built):
    import bhraytracer_b200 as b
    r = b.Renderer(0)
    yield r
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_reference_scene_on_the_gpu(renderer, workdir, name):
    import bhraytracer_b200 as b
    g = np.load(os.path.join(FIX, name + "_det.npz"))
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        scene = b.LoadScene(os.path.join(FIX, name + ".xml"))
    finally:
        os.chdir(cwd)
    assert (scene.width, scene.height) == (int(g["width"]), int(g["height"]))
    renderer.upload(scene)
    out = renderer.render(mode=b.MODE_DET, spp=1)
    check_det(out, g, name)


@pytest.mark.gpu
@pytest.mark.parametrize("exact", [False, True], ids=["fast", "exact"])
@pytest.mark.parametrize("name", STOCH)
def test_reference_scene_stochastic_mode_on_the_gpu(renderer, workdir, name, exact):
    import bhraytracer_b200 as b
    from test_gpu_convergence import check_ray_counts, conv_stats
    g = np.load(os.path.join(FIX, name + "_conv.npz"))
    A = np.clip(g["mean_a"].astype(np.float64) / 16384.0, 0, 1)
    B = np.clip(g["mean_b"].astype(np.float64) / 16384.0, 0, 1)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        scene = b.LoadScene(os.path.join(FIX, name + ".xml"))
    finally:
        os.chdir(cwd)
    renderer.upload(scene)
    w, h = scene.width, scene.height
    out16 = renderer.render(mode=b.MODE_PATH, spp=int(g["spp"]), seed=5, exact_math=exact)
    M = (A + B) / 2
    rmse = float(np.sqrt(np.mean((np.clip(out16["linear"], 0, 1) - M) ** 2)))
    assert rmse <= 1.25 * max(float(g["rmse_ref_a"]), float(g["rmse_ref_b"]))
    spp = int(g["ref_spp"])
    out = renderer.render(mode=b.MODE_PATH, spp=spp, seed=6, exact_math=exact)
    st = conv_stats(np.clip(out["linear"].astype(np.float64), 0, 1), A, B)
    msg = "%s %s: %r" % (name, "exact" if exact else "fast", st)
    print(msg)
    assert st["ours"] <= 1.05 * st["pair"], msg
    assert st["tile_excess"] <= 0.1 / 255 and abs(st["mean_diff"]) <= 0.1 / 255, msg
    check_ray_counts(out, g, w * h * spp, False, 0.01)
