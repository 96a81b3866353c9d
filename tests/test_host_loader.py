"""Host scene pipeline (libbhrt_host.so): the reference's LoadScene semantics (reference
xmlload.cpp:65-583) and the flattening pass."""
import ctypes as C
import os

import numpy as np
import pytest

import bhraytracer_b200 as b
from conftest import ROOT


def write_scene(tmp_path, body, camera=""):
    p = tmp_path / "s.xml"
    p.write_text("<xml><scene>%s</scene><camera>%s</camera></xml>" % (body, camera))
    return str(p)


def test_cfg1_flattening(built):
    s = b.LoadScene("scenes/cfg1_spheres_plane.xml")
    d = s.desc.contents
    assert (s.width, s.height) == (800, 600)
    names = [s.node_name(i) for i in range(d.n_nodes)]
    # DFS pre-order, the order recursive() visits (reference Main.cpp:389-413)
    assert names[:4] == ["floor", "redBall", "greenBall", "group"]
    assert names[4:6] == ["blueBall", "smallBall"]
    nodes = [d.nodes[i] for i in range(d.n_nodes)]
    assert nodes[3].kind == 0 and nodes[3].depth == 1 and nodes[3].parent == -1
    assert nodes[4].parent == 3 and nodes[4].depth == 2 and nodes[4].kind == 1
    assert nodes[0].kind == 2
    # itm is the inverse of tm
    tm = np.array(list(nodes[2].xf.tm), np.float64).reshape(3, 3).T
    itm = np.array(list(nodes[2].xf.itm), np.float64).reshape(3, 3).T
    assert np.allclose(tm @ itm, np.eye(3), atol=1e-5)


def test_defaults_and_case_insensitive_tags(built, tmp_path):
    """Material defaults (materials.h:23-25), camera defaults (scene.h:513-523), COMPARE is
    case-insensitive (xmlload.cpp:34-38), `value` multiplies (xmlload.cpp:478-512)."""
    path = write_scene(tmp_path, '<OBJECT type="SPHERE" name="a" material="m"><Scale value="2"/></OBJECT>'
                                 '<material type="blinn" name="m"><diffuse r="0.5" g="1" b="1" value="0.5"/></material>'
                                 '<light type="point" name="l"><intensity value="3"/><position x="1" y="2" z="3"/></light>')
    s = b.LoadScene(path)
    d = s.desc.contents
    assert d.n_nodes == 1 and d.nodes[0].kind == 1
    assert (d.camera.width, d.camera.height) == (200, 150)
    assert abs(d.camera.fov - 40) < 1e-6
    assert list(d.camera.dir) == [0, 0, -1]
    mats = (C.c_float * 18).from_address(d.materials)
    assert np.allclose(list(mats[0:3]), [0.25, 0.5, 0.5])       # diffuse * value
    assert np.allclose(list(mats[4:7]), [0.7, 0.7, 0.7])        # default specular
    assert abs(mats[14] - 20) < 1e-6 and abs(mats[15] - 1) < 1e-6   # glossiness 20, ior 1
    assert list(d.nodes[0].xf.tm)[0] == 2.0


def test_missing_obj_keeps_node_without_object(built, tmp_path):
    """xmlload.cpp:209-214,253: a mesh that fails to load leaves the node with no object."""
    path = write_scene(tmp_path, '<object type="obj" name="does/not/exist.obj" material="m"/>'
                                 '<object type="plane" name="p" material="m"/>'
                                 '<material type="blinn" name="m"/>')
    s = b.LoadScene(path)
    d = s.desc.contents
    assert d.n_nodes == 2 and d.nodes[0].kind == 0 and d.nodes[1].kind == 2
    assert "exist.obj" in s.warnings


def test_missing_file_fails_like_the_reference(built):
    with pytest.raises(b.BhrtError):
        b.LoadScene("scenes/no_such_scene.xml")


def test_mesh_tables_and_bvh_topology(built):
    """The flattened mesh keeps the reference's BVH layout (cyBVH.h:76,87-100,187-203): root = 1,
    children adjacent, leaves of at most 4 triangles (objects.h:59), every face in exactly one leaf."""
    s = b.LoadScene("scenes/det_mesh.xml")
    d = s.desc.contents
    assert d.n_meshes == 2          # torus_small is instanced twice but loaded once (xmlload.cpp:208-217)
    meshes = (C.c_int32 * (22 * d.n_meshes)).from_address(d.meshes)
    blob = np.ctypeslib.as_array((C.c_uint8 * d.mesh_bytes).from_address(d.mesh_data))
    for mi in range(d.n_meshes):
        raw = np.frombuffer(bytes(meshes)[mi * 88:(mi + 1) * 88], dtype=np.uint8)
        nv, nf, nvn, nvt, nb, _ = np.frombuffer(raw[:24].tobytes(), np.int32)
        offs = np.frombuffer(raw[24:].tobytes(), np.int64)
        assert all(o % 16 == 0 for o in offs)
        bvh = np.frombuffer(blob[offs[6]:offs[6] + 32 * (nb + 1)].tobytes(), np.uint32).reshape(-1, 8)
        elems = np.frombuffer(blob[offs[7]:offs[7] + 4 * nf].tobytes(), np.uint32)
        seen = np.zeros(nf, np.int32)
        stack = [1]
        while stack:
            n = stack.pop()
            data = int(bvh[n, 6])
            if data & 0x80000000:
                cnt, off = ((data >> 28) & 7) + 1, data & 0x0FFFFFFF
                assert cnt <= 8      # MeanSplit may fail to split coincident centroids (cyBVH.h:295-328); 3-bit count
                seen[elems[off:off + cnt]] += 1
            else:
                c = data & 0x7FFFFFFF
                stack += [c, c + 1]
                box = bvh[n, :6].view(np.float32)
                for ch in (c, c + 1):
                    cb = bvh[ch, :6].view(np.float32)
                    assert (cb[:3] >= box[:3]).all() and (cb[3:] <= box[3:]).all()
        assert (seen == 1).all()


def test_set_resolution_and_png_roundtrip(built, tmp_path):
    s = b.LoadScene("scenes/cfg1_spheres_plane.xml", 64, 48)
    assert (s.width, s.height) == (64, 48)
    img = (np.arange(64 * 48 * 3) % 251).astype(np.uint8).reshape(48, 64, 3)
    p = str(tmp_path / "x.png")
    b.save_png(p, img)
    # read it back through the loader's own PNG decoder: use it as a texture
    path = write_scene(tmp_path, '<object type="plane" name="p" material="m"/>'
                                 '<material type="blinn" name="m"><diffuse texture="%s"/></material>' % p)
    s2 = b.LoadScene(path)
    d = s2.desc.contents
    assert d.n_textures == 1 and d.texel_bytes == 64 * 48 * 3
    tex = np.ctypeslib.as_array((C.c_uint8 * d.texel_bytes).from_address(d.texels)).reshape(48, 64, 3)
    assert (tex == img).all()


def test_render_image_side_buffers(built):
    """RenderImage::ComputeZBufferImage / ComputeSampleCountImage (reference scene.h:579-626) restated in numpy."""
    rng = np.random.default_rng(3)
    z = rng.uniform(2.0, 90.0, (40, 50)).astype(np.float32)
    z[rng.random(z.shape) < 0.2] = np.float32(1.0e30)
    img = b.zbuffer_image(z)
    hit = z != np.float32(1.0e30)
    zmin, zmax = z[hit].min(), z[hit].max()
    want = np.zeros(z.shape, np.uint8)
    f = (zmax - z[hit]) / (zmax - zmin)
    want[hit] = np.clip((f * np.float32(255)).astype(np.int32), 0, 255)
    assert (img == want).all() and img[~hit].max() == 0 and img.max() == 255
    cnt = rng.integers(3, 40, (40, 50)).astype(np.uint8)
    simg, smax = b.sample_count_image(cnt)
    assert smax == cnt.max()
    assert (simg == (255 * (cnt.astype(np.int32) - cnt.min())) // (int(cnt.max()) - int(cnt.min()))).all()
    flat, m = b.sample_count_image(np.full((4, 4), 7, np.uint8))
    assert m == 7 and (flat == 0).all()


def test_usemtl_groups_and_sub_materials_are_flattened(built):
    """An OBJ's usemtl groups (cyTriMesh.h:467-490: faces regrouped by material, cumulative counts in mcfc) reach the ABI as
    bhrt_mesh::n_groups + int32 group_end[] behind the elems array, and ALL sub-materials of the auto-MultiMtl
    (xmlload.cpp:219-250) follow the first one in the material table (bhrt.h: BHRT_MTL_MULTI / BHRT_MTL_SUBCOUNT) -- what
    BHRT_FLAG_FACE_MATERIALS reads; the default frame only ever uses sub-material 0, like the reference."""
    s = b.LoadScene("scenes/det_facemtl.xml")
    d = s.desc.contents
    assert d.n_meshes == 1
    raw = bytes((C.c_uint8 * 88).from_address(d.meshes))
    nv, nf, nvn, nvt, nb, n_groups = np.frombuffer(raw[:24], np.int32)
    offs = np.frombuffer(raw[24:], np.int64)
    assert (nf, n_groups) == (12, 2)                # six quads fanned into triangles, two usemtl groups
    blob = np.ctypeslib.as_array((C.c_uint8 * d.mesh_bytes).from_address(d.mesh_data))
    at = (int(offs[7]) + 4 * int(nf) + 15) & ~15
    group_end = np.frombuffer(blob[at:at + 4 * n_groups].tobytes(), np.int32)
    assert list(group_end) == [6, 12]
    # the cube's node -> the MultiMtl's first entry, flagged with its sub-material count; matB right behind it
    cube = [d.nodes[i] for i in range(d.n_nodes) if d.nodes[i].kind == 3][0]
    mraw = bytes((C.c_uint8 * (72 * d.n_materials)).from_address(d.materials))      # bhrt_material: 72 bytes, flags at offset 64
    def mtl(i):
        r = mraw[72 * i:72 * (i + 1)]
        return np.frombuffer(r[:12], np.float32), int(np.frombuffer(r[64:68], np.int32)[0])
    (kd0, flags0), (kd1, flags1) = mtl(cube.material), mtl(cube.material + 1)
    assert flags0 & 1 and (flags0 >> 8) & 0xff == 2 and flags1 == 0
    assert np.allclose(kd0, [0.8, 0.3, 0.2]) and np.allclose(kd1, [0.2, 0.3, 0.8])
