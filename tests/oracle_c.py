"""ctypes binding of oracle/_build/libbhrt_oracle.so (oracle/bhrt_oracle.c: the plain-C restatement of
the reference's deterministic path for analytic scenes).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "_build", "libbhrt_oracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"])
        _LIB = C.CDLL(path)
        _LIB.bhrt_oracle_render_det.argtypes = [C.c_void_p] * 6
        _LIB.bhrt_oracle_render_det.restype = C.c_int
        _LIB.bhrt_oracle_in_scope.argtypes = [C.c_void_p]
        _LIB.bhrt_oracle_in_scope.restype = C.c_int
    return _LIB


def in_scope(scene):
    return bool(lib().bhrt_oracle_in_scope(C.cast(scene.desc, C.c_void_p)))


def render_det(scene, tir_budget=16, face_materials=False):
    """scene: bhraytracer_b200.Scene (the host loader's flattened description).  Also returns the
    per-pixel (closest, shadow) ray counts as pixel_rays[h, w, 2]."""
    w, h = scene.width, scene.height
    n = w * h
    hit, z = np.empty(n, np.int32), np.empty(n, np.float32)
    lin, rgb = np.empty(n * 3, np.float32), np.empty(n * 3, np.uint8)
    pix = np.zeros(n * 2, np.uint32)
    rays = (C.c_uint64 * 2)()
    L = lib()
    L.bhrt_oracle_set_face_materials(1 if face_materials else 0)       # (oracle patch Q-1 / BHRT_FLAG_FACE_MATERIALS)
    L.bhrt_oracle_render_det3.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
    L.bhrt_oracle_render_det3.restype = C.c_int
    rc = L.bhrt_oracle_render_det3(C.cast(scene.desc, C.c_void_p), tir_budget, hit.ctypes.data, z.ctypes.data, lin.ctypes.data, rgb.ctypes.data,
                                   C.cast(rays, C.c_void_p), pix.ctypes.data)
    L.bhrt_oracle_set_face_materials(0)
    if rc != 0:
        raise ValueError("scene outside the restatement's scope (rc %d)" % rc)
    return dict(hit_id=hit.reshape(h, w), z=z.reshape(h, w), linear=lin.reshape(h, w, 3), rgb8=rgb.reshape(h, w, 3),
                rays_closest=int(rays[0]), rays_shadow=int(rays[1]), pixel_rays=pix.reshape(h, w, 2), max_depth=int(L.bhrt_oracle_last_max_depth()))


def render_det_pm(scene, photons, tir_budget=16):
    """The gather-lit deterministic frame; photons: uint8[n, 24], the reference's own .dat dump."""
    w, h = scene.width, scene.height
    n = w * h
    hit, z = np.empty(n, np.int32), np.empty(n, np.float32)
    lin, rgb = np.empty(n * 3, np.float32), np.empty(n * 3, np.uint8)
    rays = (C.c_uint64 * 2)()
    ph = np.ascontiguousarray(photons, dtype=np.uint8).reshape(-1, 24)
    L = lib()
    L.bhrt_oracle_render_det_pm.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5
    L.bhrt_oracle_render_det_pm.restype = C.c_int
    rc = L.bhrt_oracle_render_det_pm(C.cast(scene.desc, C.c_void_p), tir_budget, ph.ctypes.data, len(ph), hit.ctypes.data, z.ctypes.data,
                                     lin.ctypes.data, rgb.ctypes.data, C.cast(rays, C.c_void_p))
    if rc != 0:
        raise ValueError("scene outside the restatement's scope (rc %d)" % rc)
    return dict(hit_id=hit.reshape(h, w), z=z.reshape(h, w), linear=lin.reshape(h, w, 3), rgb8=rgb.reshape(h, w, 3),
                rays_closest=int(rays[0]), rays_shadow=int(rays[1]))
