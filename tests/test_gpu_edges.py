"""GPU tier, edge cases of the hot path through the C ABI: frame sizes that are not multiples of the
8x8 work tiles (ragged), a 1x1 frame, more lanes than work, empty shares, an empty photon map, limits."""
import json
import os
import subprocess

import numpy as np
import pytest

import bhraytracer_b200 as b
from aovio import read_aov
from conftest import EMU_BIN, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer(built):
    r = b.Renderer(0)
    yield r
    r.close()


def emu(scene, prefix, *args):
    out = subprocess.check_output([EMU_BIN, scene, prefix] + [str(a) for a in args], cwd=ROOT, stderr=subprocess.DEVNULL)
    return json.loads(out.decode().splitlines()[-1]), read_aov(prefix + ".aov")


@pytest.mark.parametrize("w,h", [(1, 1), (7, 5), (37, 23), (129, 65), (8, 200)])
def test_ragged_frame_sizes_match_the_emulated_kernel(renderer, tmp_path, w, h):
    scene = b.LoadScene("scenes/det_refract.xml", w, h)
    renderer.upload(scene)
    out = renderer.render(mode=b.MODE_DET)
    info, e = emu("scenes/det_refract.xml", str(tmp_path / "e"), "--det", "--size", w, h)
    assert out["hit_id"].shape == (h, w)
    assert (out["hit_id"] == e["hit_id"]).all()
    assert (out["z"].view(np.uint32) == e["z"].view(np.uint32)).all()
    assert np.abs(out["rgb8"].astype(int) - e["rgb8"].astype(int)).max() <= 1
    assert (out["rays_closest"], out["rays_shadow"]) == (info["rays_closest"], info["rays_shadow"])


def test_path_mode_every_pixel_gets_every_sample_once(renderer):
    """Sample bookkeeping at a ragged size with more lanes in the pools than camera samples: the
    accumulation buffer's count channel must be exactly spp everywhere."""
    import ctypes as C
    import torch
    scene = b.LoadScene("scenes/cfg1_spheres_plane.xml", 61, 43)
    renderer.upload(scene)
    renderer.render_device(renderer.params(mode=b.MODE_PATH, spp=5, seed=2))
    ptr, nbytes = renderer.accum_buffer()

    class _A:
        __cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
    acc = torch.as_tensor(_A(), device="cuda").cpu().numpy().reshape(43, 61, 4)
    assert (acc[:, :, 3] == 5).all()
    assert np.isfinite(acc).all() and (acc[:, :, :3] >= 0).all() and (acc[:, :, :3] <= 3 * 5 + 1e-4).all()  # a sample is a sum of three terms clamped to 1 each (MtlBlinn.cpp:117-131)


def test_empty_sample_range_and_empty_rows_clear_the_buffers(renderer):
    scene = b.LoadScene("scenes/cfg1_spheres_plane.xml", 64, 48)
    renderer.upload(scene)
    renderer.render_device(renderer.params(mode=b.MODE_PATH, spp=4, sample_range=(4, 4)))
    out = renderer.resolve(1)
    assert (out["linear"] == 0).all()
    st = renderer.render_device(renderer.params(mode=b.MODE_DET, rows=(10, 10)))
    assert st["rays_closest"] == 0 and st["rays_shadow"] == 0
    assert (renderer.resolve(1)["linear"] == 0).all()


def test_same_seed_same_bits_and_different_seed_differs(renderer):
    scene = b.LoadScene("scenes/det_mesh.xml", 120, 90)
    renderer.upload(scene)
    a = renderer.render(mode=b.MODE_PATH, spp=4, seed=11)
    c = renderer.render(mode=b.MODE_PATH, spp=4, seed=11)
    d = renderer.render(mode=b.MODE_PATH, spp=4, seed=12)
    assert (a["linear"].view(np.uint32) == c["linear"].view(np.uint32)).all()       # independent of scheduling
    assert (a["linear"] != d["linear"]).any()


def test_render_with_empty_photon_map_and_without_build(renderer):
    scene = b.LoadScene("scenes/det_caustics.xml", 80, 60)
    renderer.upload(scene)
    with pytest.raises(b.BhrtError):
        renderer.render(mode=b.MODE_DET, use_photon_map=True)          # no map built for this scene
    renderer.photons_set(np.zeros((0, 24), np.uint8))
    renderer.photons_build(1)
    with_empty = renderer.render(mode=b.MODE_DET, use_photon_map=True)
    without = renderer.render(mode=b.MODE_DET)
    assert (with_empty["rgb8"] == without["rgb8"]).all()


def test_bad_arguments_return_error_codes(renderer):
    scene = b.LoadScene("scenes/cfg1_spheres_plane.xml", 32, 24)
    renderer.upload(scene)
    for kw in (dict(spp=0), dict(sample_range=(3, 2), spp=4), dict(rows=(0, 25)), dict(mode=7)):
        with pytest.raises(b.BhrtError):
            renderer.render(**kw)
    r2 = b.Renderer(0)
    with pytest.raises(b.BhrtError):
        r2.render_device(r2.params(mode=b.MODE_DET))          # no scene uploaded
    r2.close()


def test_sample_count_aov(built):
    """RenderImage::GetSampleCount (scene.h:534,570): the camera samples actually accumulated per pixel -- all spp after a
    full frame, only the rendered rows / samples after a partial one -- and its stretched 8-bit view (scene.h:602-626)."""
    r = b.Renderer(0)
    scene = b.LoadScene("scenes/cfg1_spheres_plane.xml", 96, 64)
    r.upload(scene)
    r.render(mode=b.MODE_PATH, spp=5, seed=1, want=())
    assert (r.sample_counts() == 5).all()
    r.render_device(r.params(mode=b.MODE_PATH, spp=5, seed=1, sample_range=(1, 3)))
    assert (r.sample_counts() == 2).all()
    r.render_device(r.params(mode=b.MODE_DET, rows=(1, 64), row_stride=2))
    c = r.sample_counts()
    assert (c[1::2] == 1).all() and (c[0::2] == 0).all()
    img, mx = b.sample_count_image(c)
    assert mx == 1 and (img[1::2] == 255).all() and (img[0::2] == 0).all()
    r.close()


def test_two_odd_sized_file_textures_upload_and_render(built, tmp_path):
    """the loader pads the texel blob between textures and the device library used to refuse offsets that are
    not multiples of 3.  Two PPM textures of 5x5 and 5x7 texels (75 and 105 bytes) through loader -> upload -> render,
    against the CPU build of the device code."""
    rng = np.random.default_rng(3)
    for name, (w, h) in (("a.ppm", (5, 5)), ("b.ppm", (5, 7))):
        with open(tmp_path / name, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (w, h) + rng.integers(0, 256, (h, w, 3), dtype=np.uint8).tobytes())
    scene = tmp_path / "two_tex.xml"
    scene.write_text('''<xml><scene>
      <object type="plane" name="floor" material="m1"><scale value="10"/></object>
      <object type="plane" name="wall" material="m2"><scale value="10"/><rotate angle="90" x="1"/><translate y="8" z="5"/></object>
      <material type="blinn" name="m1"><diffuse texture="a.ppm"/><specular value="0"/></material>
      <material type="blinn" name="m2"><diffuse texture="b.ppm"/><specular value="0"/></material>
      <light type="point" name="key"><intensity value="250"/><position x="-3" y="-6" z="12"/></light>
    </scene>
    <camera><position x="0" y="-20" z="7"/><target x="0" y="0" z="3"/><up x="0" y="0" z="1"/><fov value="40"/><width value="96"/><height value="64"/></camera></xml>''')
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        sc = b.LoadScene(str(scene))
        r = b.Renderer(0)
        r.upload(sc)                    # used to fail: "texel_offset is not a multiple of 3"
        out = r.render(mode=b.MODE_DET)
        r.close()
        subprocess.check_output([EMU_BIN, str(scene), str(tmp_path / "emu"), "--det"], cwd=str(tmp_path), stderr=subprocess.DEVNULL)
    finally:
        os.chdir(cwd)
    emu = read_aov(str(tmp_path / "emu.aov"))
    assert (out["hit_id"] == emu["hit_id"]).all()
    assert np.abs(out["rgb8"].astype(int) - emu["rgb8"].astype(int)).max() <= 1
    assert out["rgb8"][out["hit_id"] == 0].std() > 5 and out["rgb8"][out["hit_id"] == 1].std() > 5      # both textures are in the picture


def test_photon_emission_ends_when_nothing_can_be_stored(built, tmp_path):
    """BuildCausticPhotonMap's loop (Main.cpp:363) never ends in a scene that cannot store a caustic photon;
    the emission kernel is bounded and reports the short count instead of hanging the GPU."""
    path = tmp_path / "diffuse_only.xml"
    path.write_text('''<xml><scene>
      <object type="plane" name="floor" material="m"><scale value="10"/></object>
      <object type="sphere" name="ball" material="m"><translate z="1"/></object>
      <material type="blinn" name="m"><diffuse value="0.7"/><specular value="0"/></material>
      <light type="point" name="key"><intensity value="100"/><position x="0" y="0" z="8"/></light>
    </scene>
    <camera><position x="0" y="-12" z="4"/><target x="0" y="0" z="1"/><up x="0" y="0" z="1"/><fov value="40"/><width value="64"/><height value="48"/></camera></xml>''')
    r = b.Renderer(0)
    scene = b.LoadScene(str(path))
    r.upload(scene)
    os.environ["BHRT_PHOTON_MAX_PATHS"] = "2000000"
    try:
        stored, rays = r.photons_emit(5000, seed=1)
    finally:
        del os.environ["BHRT_PHOTON_MAX_PATHS"]
    assert stored == 0 and rays >= 2000000          # the path budget ran out; the call returned
    r.photons_build(max(stored, 1))
    out = r.render(mode=b.MODE_DET, use_photon_map=True)
    assert np.isfinite(out["linear"]).all()
    r.close()
