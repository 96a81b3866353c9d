"""The kernels' two mesh traversals, ray by ray on the GPU (bhrt_debug_trace_mesh, exact arithmetic): eight
lanes per ray over the 8-wide regrouping of the BVH (csrc/device/wide.cuh) against one ray per thread over the
binary tree (TriObj::TraceBVHNode / TraceBVHShadow made iterative, reference TriObj.cpp:192-307) -- and whole
frames of the fast build with either traversal."""
import os

import numpy as np
import pytest

import bhraytracer_b200 as b
from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer(built):
    r = b.Renderer(0)
    yield r
    r.close()


def obj_bounds(path):
    v = np.array([[float(x) for x in l.split()[1:4]] for l in open(path) if l.startswith("v ")], dtype=np.float64)
    return v.min(axis=0), v.max(axis=0)


def random_rays(lo, hi, n, rng):
    """Closest-hit rays from outside the box through it and from inside it; shadow segments across it."""
    c, e = (lo + hi) / 2, (hi - lo) / 2
    r = np.zeros((n, 8), np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    target = c + e * rng.uniform(-1, 1, size=(n, 3))
    outside = rng.random(n) < 0.7
    origin = np.where(outside[:, None], c + 2.5 * np.linalg.norm(e) * d, c + e * rng.uniform(-1, 1, size=(n, 3)))
    dirv = target - origin
    scale = np.where(rng.random(n) < 0.5, 1.0, rng.uniform(0.2, 3.0, size=n))           # unnormalised directions too
    r[:, 0:3] = origin
    r[:, 3:6] = dirv * scale[:, None]
    shadow = rng.random(n) < 0.5
    r[:, 6] = np.where(shadow, rng.choice([1.0, 0.5, 2.0], size=n), 1e30)
    r[:, 7] = np.where(shadow, 2.0, 1.0)
    return r


@pytest.mark.parametrize("scene,obj", [("det_mesh", "blob_small.obj"), ("cfg2_mesh_refract", "blob_100k.obj"), ("cfg3_textured_arealight", "torus_50k.obj")])
def test_wide_traversal_equals_binary_ray_by_ray(renderer, scene, obj):
    sc = b.LoadScene(os.path.join("scenes", scene + ".xml"), 64, 36)
    renderer.upload(sc)
    lo, hi = obj_bounds(os.path.join(ROOT, "scenes", "assets", obj))
    rays = random_rays(lo, hi, 400_000, np.random.default_rng(5))
    a, w = renderer.debug_trace_mesh(0, rays)
    a, w = renderer.debug_trace_mesh(0, rays)       # (second call: warm caches for the timing printed below)
    ms = renderer.debug_trace_ms
    print("%s: %d random rays, binary (1 ray/thread) %.3f ms = %.0f Mrays/s, wide (8 lanes/ray) %.3f ms = %.0f Mrays/s" % (
        scene, len(rays), ms[0], len(rays) / ms[0] / 1e3, ms[1], len(rays) / ms[1] / 1e3))
    shadow = rays[:, 7] == 2.0
    assert a[:, 0].sum() > 0.2 * len(rays), "the rays hardly hit the mesh: bad test geometry"
    # shadow: the reference's first-leaf answer, identical (the exact build's binary traversal also visits children
    # whose box the ray misses, which can only matter through rounding: allow a handful)
    bad_shadow = int((a[shadow, 0] != w[shadow, 0]).sum())
    assert bad_shadow <= 4, "%d of %d shadow answers differ" % (bad_shadow, int(shadow.sum()))
    cl = ~shadow
    assert (a[cl, 0] == w[cl, 0]).mean() > 0.99999
    both = cl & (a[:, 0] == 1) & (w[:, 0] == 1)
    zdiff = a[both, 1].view(np.uint32) != w[both, 1].view(np.uint32)
    assert zdiff.sum() <= 4, "%d closest hits differ in z" % int(zdiff.sum())
    # same triangle except on exact ties (shared edges): same z there
    same_face = a[both, 2] == w[both, 2]
    assert same_face.mean() > 0.999


@pytest.mark.parametrize("scene", ["det_mesh", "cfg2_mesh_refract"])
def test_fast_build_frames_agree_between_the_two_traversals(renderer, scene):
    """Same seeds -> same samples: the fast build with the wide traversal and with the binary one differ only where a
    ray meets a tie or a rounding case; ray counts within 0.1 %."""
    sc = b.LoadScene(os.path.join("scenes", scene + ".xml"), 320, 180)
    renderer.upload(sc)
    w = renderer.render(mode=b.MODE_PATH, spp=8, seed=4, wide_bvh=True)
    a = renderer.render(mode=b.MODE_PATH, spp=8, seed=4)
    d = np.abs(w["linear"] - a["linear"]).max(axis=2)
    assert (d > 1e-3).mean() < 0.01, "%.4f of the pixels differ" % float((d > 1e-3).mean())
    assert abs(w["linear"].mean() - a["linear"].mean()) < 2e-4
    rw, ra = w["rays_closest"] + w["rays_shadow"], a["rays_closest"] + a["rays_shadow"]
    assert abs(rw - ra) / ra < 0.001
