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
