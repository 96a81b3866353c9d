"""The 8-wide regrouping of the reference's BVH (csrc/device/accel.hpp: build_wide) and the traversal the
kernels' fast build runs over it, checked on the CPU through its single-thread statement
(core.cuh: wide_trav_run, `emu_render --wide`): the frame must be the one the binary traversal produces --
which is the reference's (tests/test_emu_parity.py) -- including the mesh-shadow first-leaf rule
(TriObj.cpp:272-291), which is why leaf groups remember where the reference's own leaves start."""
import json
import os
import subprocess

import numpy as np
import pytest

from aovio import read_aov
from conftest import EMU_BIN, ROOT

GOLD = os.path.join(ROOT, "tests", "golden")


def emu(scene, prefix, *args):
    out = subprocess.check_output([EMU_BIN, scene, prefix] + [str(a) for a in args], cwd=ROOT, stderr=subprocess.DEVNULL)
    return json.loads(out.decode().splitlines()[-1]), read_aov(prefix + ".aov")


@pytest.mark.parametrize("name", ["det_mesh", "det_objmtl"])
def test_wide_traversal_reproduces_the_reference_golden(built, tmp_path, name):
    g = np.load(os.path.join(GOLD, name + "_det.npz"))
    info, a = emu(os.path.join("scenes", name + ".xml"), str(tmp_path / "w"), "--det", "--size", int(g["width"]), int(g["height"]), "--wide")
    assert (a["hit_id"] == g["hit_id"].astype(np.int32)).all()
    assert (a["z"].view(np.uint32) == g["z_bits"]).all()
    assert np.abs(a["rgb8"].astype(int) - g["rgb8"].astype(int)).max() <= 1
    assert (info["rays_closest"], info["rays_shadow"]) == (int(g["rays_closest"]), int(g["rays_shadow"]))


@pytest.mark.parametrize("name,size", [("cfg2_mesh_refract", (240, 135)), ("cfg3_textured_arealight", (240, 135)), ("cfg4_gi", (192, 108))])
def test_wide_equals_binary_on_the_config_meshes(built, tmp_path, name, size):
    """100 k / 50 k triangle meshes (BVH depth 27 / 17): same hit ids, z bits, 8-bit image and ray counts."""
    scene = os.path.join("scenes", name + ".xml")
    ib, b = emu(scene, str(tmp_path / "b"), "--det", "--size", *size)
    iw, w = emu(scene, str(tmp_path / "w"), "--det", "--size", *size, "--wide")
    assert (b["hit_id"] == w["hit_id"]).all()
    assert (b["z"].view(np.uint32) == w["z"].view(np.uint32)).all()
    assert (b["rgb8"] == w["rgb8"]).all()
    assert (ib["rays_closest"], ib["rays_shadow"]) == (iw["rays_closest"], iw["rays_shadow"])


def test_wide_stochastic_frame_matches_binary_statistically(built, tmp_path):
    """Same seeds, same estimator: the two traversals may only differ where a ray meets an exact tie."""
    scene = os.path.join("scenes", "det_mesh.xml")
    ib, b = emu(scene, str(tmp_path / "b"), "--spp", 8, "--seed", 3, "--size", 96, 72)
    iw, w = emu(scene, str(tmp_path / "w"), "--spp", 8, "--seed", 3, "--size", 96, 72, "--wide")
    d = np.abs(b["linear"] - w["linear"]).max(axis=2)
    assert (d > 1e-4).mean() < 0.002
    assert abs(ib["rays_closest"] - iw["rays_closest"]) <= 0.001 * ib["rays_closest"]
