"""bhraytracer_b200 -- Python harness over the C ABI of the B200 rendering core.

The product is two in-tree shared libraries built from ``csrc/`` (see ``__graft_entry__.build``):

* ``lib/libbhrt_host.so`` -- host C++ scene pipeline: the reference's ``LoadScene`` XML semantics
  (reference ``xmlload.cpp:65``) + the flattening pass (``include/bhrt_host.h``);
* ``lib/libbhrt.so``      -- hand-written sm_100a CUDA kernels behind ``include/bhrt.h``; replaces
  ``BeginRender()`` (reference ``Main.cpp:178``).

This module only binds them with ``ctypes`` (plain pointers and sizes; no torch types cross the
boundary) for the tests and ``bench.py``.  There is no CPU fallback: if ``libbhrt.so`` is missing
or no CUDA device is present, construction of :class:`Renderer` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("BHRT_LIB_DIR") or os.path.join(_HERE, "lib")

MODE_PATH = 0   # "H-mode": HEAD's stochastic estimator
MODE_DET = 1    # "D-mode": the deterministic switches of SURVEY.md 8(c)


class BhrtError(RuntimeError):
    pass


class _XForm(C.Structure):
    _fields_ = [("tm", C.c_float * 9), ("itm", C.c_float * 9), ("pos", C.c_float * 3)]


class _Node(C.Structure):
    _fields_ = [("xf", _XForm), ("parent", C.c_int32), ("depth", C.c_int32), ("kind", C.c_int32),
                ("material", C.c_int32), ("mesh", C.c_int32)]


class _TexColor(C.Structure):
    _fields_ = [("color", C.c_float * 3), ("texmap", C.c_int32)]


class _Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("dir", C.c_float * 3), ("up", C.c_float * 3),
                ("fov", C.c_float), ("focaldist", C.c_float), ("dof", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32)]


class SceneDesc(C.Structure):
    """bhrt_scene_desc (include/bhrt.h)."""
    _fields_ = [("abi_version", C.c_int32),
                ("n_nodes", C.c_int32), ("nodes", C.POINTER(_Node)),
                ("n_materials", C.c_int32), ("materials", C.c_void_p),
                ("n_lights", C.c_int32), ("lights", C.c_void_p),
                ("n_texmaps", C.c_int32), ("texmaps", C.c_void_p),
                ("n_textures", C.c_int32), ("textures", C.c_void_p),
                ("n_meshes", C.c_int32), ("meshes", C.c_void_p),
                ("texels", C.c_void_p), ("texel_bytes", C.c_size_t),
                ("mesh_data", C.c_void_p), ("mesh_bytes", C.c_size_t),
                ("background", _TexColor), ("environment", _TexColor),
                ("camera", _Camera)]


FLAG_EXACT_MATH = 1      # BHRT_FLAG_EXACT_MATH
FLAG_WIDE_BVH = 2        # BHRT_FLAG_WIDE_BVH
FLAG_FACE_MATERIALS = 4  # BHRT_FLAG_FACE_MATERIALS


class Params(C.Structure):
    """bhrt_params (include/bhrt.h)."""
    _fields_ = [("mode", C.c_int32), ("spp", C.c_int32), ("gi_depth", C.c_int32), ("tir_budget", C.c_int32),
                ("seed", C.c_uint64),
                ("sample_begin", C.c_int32), ("sample_end", C.c_int32),
                ("row_begin", C.c_int32), ("row_end", C.c_int32), ("row_stride", C.c_int32),
                ("use_photon_map", C.c_int32), ("photon_k", C.c_int32), ("photon_radius", C.c_float),
                ("flags", C.c_uint32)]


class Outputs(C.Structure):
    """bhrt_outputs (include/bhrt.h)."""
    _fields_ = [("rgb8", C.c_void_p), ("linear", C.c_void_p), ("hit_id", C.c_void_p), ("z", C.c_void_p),
                ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("kernel_ms", C.c_double)]


# every symbol include/bhrt.h and include/bhrt_host.h declare
DEVICE_SYMBOLS = ["bhrt_abi_version", "bhrt_last_error", "bhrt_create", "bhrt_destroy", "bhrt_upload_scene",
                  "bhrt_render", "bhrt_render_device", "bhrt_accum_buffer", "bhrt_aov_buffers", "bhrt_resolve",
                  "bhrt_set_stream", "bhrt_launch_count",
                  "bhrt_photons_emit", "bhrt_photons_buffer", "bhrt_photons_set", "bhrt_photons_get",
                  "bhrt_photons_build", "bhrt_debug_trace_mesh", "bhrt_device_and_stream", "bhrt_frame_size", "bhrt_set_last_error",
                  "bhrt_render_async", "bhrt_progress", "bhrt_cancel", "bhrt_wait", "bhrt_sample_counts", "bhrt_photons_emit_global"]
DIST_SYMBOLS = ["bhrt_dist_shard_rows", "bhrt_dist_shard_samples", "bhrt_dist_shard_photons", "bhrt_dist_unique_id", "bhrt_dist_create",
                "bhrt_dist_destroy", "bhrt_dist_render", "bhrt_dist_photons"]
HOST_SYMBOLS = ["bhrt_host_load_scene", "bhrt_host_free_scene", "bhrt_host_last_error", "bhrt_host_scene_desc",
                "bhrt_host_node_name", "bhrt_host_warnings", "bhrt_host_set_resolution", "bhrt_host_save_png",
                "bhrt_host_zbuffer_image", "bhrt_host_sample_count_image"]

_host = None
_dev = None


def host_lib():
    global _host
    if _host is None:
        path = os.path.join(LIB_DIR, "libbhrt_host.so")
        if not os.path.exists(path):
            raise BhrtError("%s not built -- run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
        L = C.CDLL(path)
        L.bhrt_host_load_scene.restype = C.c_void_p
        L.bhrt_host_load_scene.argtypes = [C.c_char_p]
        L.bhrt_host_free_scene.argtypes = [C.c_void_p]
        L.bhrt_host_last_error.restype = C.c_char_p
        L.bhrt_host_scene_desc.restype = C.POINTER(SceneDesc)
        L.bhrt_host_scene_desc.argtypes = [C.c_void_p]
        L.bhrt_host_node_name.restype = C.c_char_p
        L.bhrt_host_node_name.argtypes = [C.c_void_p, C.c_int32]
        L.bhrt_host_warnings.restype = C.c_char_p
        L.bhrt_host_warnings.argtypes = [C.c_void_p]
        L.bhrt_host_set_resolution.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.bhrt_host_save_png.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.bhrt_host_zbuffer_image.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.bhrt_host_zbuffer_image.restype = None
        L.bhrt_host_sample_count_image.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.bhrt_host_sample_count_image.restype = C.c_int
        _host = L
    return _host


def device_lib():
    """libbhrt.so.  Loading works without a GPU (static cudart); creating a ctx does not."""
    global _dev
    if _dev is None:
        path = os.path.join(LIB_DIR, "libbhrt.so")
        if not os.path.exists(path):
            raise BhrtError("%s not built -- the CUDA extension is required, there is no fallback" % path)
        L = C.CDLL(path)
        L.bhrt_last_error.restype = C.c_char_p
        L.bhrt_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.bhrt_destroy.argtypes = [C.c_void_p]
        L.bhrt_upload_scene.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        L.bhrt_render.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Outputs)]
        L.bhrt_render_device.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Outputs)]
        L.bhrt_accum_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.bhrt_resolve.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Outputs)]
        L.bhrt_aov_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.bhrt_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.bhrt_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.bhrt_photons_emit.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.bhrt_photons_emit_global.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.bhrt_photons_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
        L.bhrt_photons_set.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int]
        L.bhrt_photons_get.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        L.bhrt_photons_build.argtypes = [C.c_void_p, C.c_int32]
        L.bhrt_sample_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.bhrt_render_async.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Outputs)]
        L.bhrt_progress.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.bhrt_cancel.argtypes = [C.c_void_p]
        L.bhrt_wait.argtypes = [C.c_void_p]
        L.bhrt_dist_shard_rows.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.bhrt_dist_shard_samples.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.bhrt_dist_shard_photons.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.bhrt_dist_shard_photons.restype = C.c_int32
        L.bhrt_dist_unique_id.argtypes = [C.c_void_p]
        L.bhrt_dist_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
        L.bhrt_dist_destroy.argtypes = [C.c_void_p]
        L.bhrt_dist_render.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int32, C.POINTER(Outputs)]
        L.bhrt_dist_photons.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.bhrt_debug_trace_mesh.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _dev = L
    return _dev


class Scene:
    """A loaded + flattened scene (host side).  ``Scene(path)`` == the reference's LoadScene(path)."""

    def __init__(self, xml_path, width=None, height=None):
        L = host_lib()
        self._h = L.bhrt_host_load_scene(os.fsencode(xml_path))
        if not self._h:
            raise BhrtError("LoadScene(%s): %s" % (xml_path, L.bhrt_host_last_error().decode()))
        if width and height:
            if L.bhrt_host_set_resolution(self._h, int(width), int(height)) != 0:
                raise BhrtError("bad resolution")
        self.path = xml_path

    @property
    def desc(self):
        return host_lib().bhrt_host_scene_desc(self._h)

    @property
    def width(self):
        return self.desc.contents.camera.width

    @property
    def height(self):
        return self.desc.contents.camera.height

    @property
    def n_nodes(self):
        return self.desc.contents.n_nodes

    def node_name(self, i):
        return host_lib().bhrt_host_node_name(self._h, i).decode()

    @property
    def warnings(self):
        return host_lib().bhrt_host_warnings(self._h).decode()

    def close(self):
        if getattr(self, "_h", None):
            host_lib().bhrt_host_free_scene(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def LoadScene(filename, width=None, height=None):
    """Reference-named entry point (xmlload.cpp:65)."""
    return Scene(filename, width, height)


def zbuffer_image(z):
    """RenderImage::ComputeZBufferImage (scene.h:579-600) of a float z buffer (bhrt_outputs.z)."""
    z = np.ascontiguousarray(z, dtype=np.float32)
    out = np.empty(z.shape, np.uint8)
    host_lib().bhrt_host_zbuffer_image(z.ctypes.data, z.size, out.ctypes.data)
    return out


def sample_count_image(count):
    """RenderImage::ComputeSampleCountImage (scene.h:602-626); returns (image, max count)."""
    count = np.ascontiguousarray(count, dtype=np.uint8)
    out = np.empty(count.shape, np.uint8)
    smax = host_lib().bhrt_host_sample_count_image(count.ctypes.data, count.size, out.ctypes.data)
    return out, smax


def save_png(path, img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    comp = 1 if img.ndim == 2 else img.shape[2]
    if host_lib().bhrt_host_save_png(os.fsencode(path), img.ctypes.data, w, h, comp) != 0:
        raise BhrtError("cannot write %s" % path)


class Renderer:
    """One GPU's render context (bhrt_ctx).  ``render()`` == the reference's BeginRender()."""

    def __init__(self, device=0):
        self._L = device_lib()
        self._ctx = C.c_void_p()
        self._check(self._L.bhrt_create(int(device), C.byref(self._ctx)))
        self.device = device
        self.width = self.height = 0

    def _check(self, rc):
        if rc != 0:
            raise BhrtError("libbhrt error %d: %s" % (rc, self._L.bhrt_last_error().decode()))

    def upload(self, scene):
        self._check(self._L.bhrt_upload_scene(self._ctx, scene.desc))
        self.width, self.height = scene.width, scene.height
        self._scene = scene      # keep the host tables alive

    @staticmethod
    def params(mode=MODE_PATH, spp=32, gi_depth=3, tir_budget=16, seed=1, sample_range=None, rows=None,
               row_stride=1, use_photon_map=False, photon_k=1000, photon_radius=0.5, exact_math=False, wide_bvh=False, face_materials=False):
        p = Params()
        p.mode, p.spp, p.gi_depth, p.tir_budget, p.seed = mode, spp, gi_depth, tir_budget, seed
        if sample_range:
            p.sample_begin, p.sample_end = sample_range
        if rows:
            p.row_begin, p.row_end = rows
        p.row_stride = row_stride
        p.use_photon_map, p.photon_k, p.photon_radius = int(use_photon_map), photon_k, photon_radius
        p.flags = (FLAG_EXACT_MATH if exact_math else 0) | (FLAG_WIDE_BVH if wide_bvh else 0) | (FLAG_FACE_MATERIALS if face_materials else 0)
        return p

    def _alloc_outputs(self, want):
        n = self.width * self.height
        o = Outputs()
        bufs = {}
        if "rgb8" in want:
            bufs["rgb8"] = np.empty((self.height, self.width, 3), np.uint8)
            o.rgb8 = bufs["rgb8"].ctypes.data
        if "linear" in want:
            bufs["linear"] = np.empty((self.height, self.width, 3), np.float32)
            o.linear = bufs["linear"].ctypes.data
        if "hit_id" in want:
            bufs["hit_id"] = np.empty((self.height, self.width), np.int32)
            o.hit_id = bufs["hit_id"].ctypes.data
        if "z" in want:
            bufs["z"] = np.empty((self.height, self.width), np.float32)
            o.z = bufs["z"].ctypes.data
        assert n > 0
        return o, bufs

    def render(self, params=None, want=("rgb8", "linear", "hit_id", "z"), **kw):
        """Render + resolve + copy the requested outputs to host numpy arrays."""
        p = params or self.params(**kw)
        o, bufs = self._alloc_outputs(want)
        self._check(self._L.bhrt_render(self._ctx, C.byref(p), C.byref(o)))
        bufs.update(rays_closest=int(o.rays_closest), rays_shadow=int(o.rays_shadow), kernel_ms=float(o.kernel_ms))
        return bufs

    def render_into(self, params, rgb8=None, linear=None, hit_id=None, z=None):
        """bhrt_render with caller-owned HOST buffers given as raw addresses (e.g. pinned memory)."""
        o = Outputs()
        o.rgb8, o.linear, o.hit_id, o.z = rgb8, linear, hit_id, z
        self._check(self._L.bhrt_render(self._ctx, C.byref(params), C.byref(o)))
        return dict(rays_closest=int(o.rays_closest), rays_shadow=int(o.rays_shadow), kernel_ms=float(o.kernel_ms))

    def render_device(self, params=None, stats=True, **kw):
        """Render into the device accumulation buffer only (nothing crosses PCIe except counters)."""
        p = params or self.params(**kw)
        o = Outputs()
        self._check(self._L.bhrt_render_device(self._ctx, C.byref(p), C.byref(o) if stats else None))
        return dict(rays_closest=int(o.rays_closest), rays_shadow=int(o.rays_shadow), kernel_ms=float(o.kernel_ms))

    def accum_buffer(self):
        """(device pointer, bytes) of the float4 accumulation buffer, for an NCCL all-reduce."""
        ptr, n = C.c_void_p(), C.c_size_t()
        self._check(self._L.bhrt_accum_buffer(self._ctx, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def aov_buffers(self):
        """Device pointers of the int32 hit-id and float z buffers (W*H each)."""
        a, z = C.c_void_p(), C.c_void_p()
        self._check(self._L.bhrt_aov_buffers(self._ctx, C.byref(a), C.byref(z)))
        return a.value, z.value

    def resolve(self, total_spp, want=("rgb8", "linear", "hit_id", "z")):
        if not want:
            self._check(self._L.bhrt_resolve(self._ctx, int(total_spp), None))
            return {}
        o, bufs = self._alloc_outputs(want)
        self._check(self._L.bhrt_resolve(self._ctx, int(total_spp), C.byref(o)))
        return bufs

    def set_stream(self, cuda_stream):
        """Run this ctx's work on the given cudaStream_t (int handle; 0/None = its own stream)."""
        self._check(self._L.bhrt_set_stream(self._ctx, C.c_void_p(int(cuda_stream) if cuda_stream else None)))

    def launch_count(self):
        n = C.c_uint64()
        self._check(self._L.bhrt_launch_count(self._ctx, C.byref(n)))
        return n.value

    # ---- the viewer's asynchronous contract (reference viewport.cpp:37-38,395-415,449-453)
    def render_async(self, params=None, want=("rgb8", "linear", "hit_id", "z"), **kw):
        """BeginRender() on the library's worker thread; returns at once.  Poll progress(), then wait()."""
        p = params or self.params(**kw)
        o, bufs = self._alloc_outputs(want)
        self._check(self._L.bhrt_render_async(self._ctx, C.byref(p), C.byref(o)))
        # (kept alive until wait(): the library's worker thread reads params and writes the outputs)
        self._async_p, self._async_o, self._async_bufs = p, o, bufs

    def progress(self):
        """(pixels done, pixels total, finished) -- RenderImage::GetNumRenderedPixels / IsRenderDone."""
        d, t, f = C.c_int64(), C.c_int64(), C.c_int32()
        self._check(self._L.bhrt_progress(self._ctx, C.byref(d), C.byref(t), C.byref(f)))
        return d.value, t.value, bool(f.value)

    def cancel(self):
        """StopRender()."""
        self._check(self._L.bhrt_cancel(self._ctx))

    def wait(self):
        """Joins the frame started by render_async; returns its outputs (raises BhrtError -5 after a cancel)."""
        self._check(self._L.bhrt_wait(self._ctx))
        o, bufs = self._async_o, self._async_bufs
        bufs.update(rays_closest=int(o.rays_closest), rays_shadow=int(o.rays_shadow), kernel_ms=float(o.kernel_ms))
        return bufs

    def sample_counts(self):
        """bhrt_sample_counts: uint8[H, W] camera samples accumulated per pixel by the last render."""
        out = np.empty((self.height, self.width), np.uint8)
        self._check(self._L.bhrt_sample_counts(self._ctx, out.ctypes.data))
        return out

    def debug_trace_mesh(self, mesh, rays):
        """bhrt_debug_trace_mesh: rays float32[n, 8] -> (binary, wide), float32[n, 4] each (result, z, face, front)."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = rays.shape[0]
        a, w = np.empty((n, 4), np.float32), np.empty((n, 4), np.float32)
        ms = np.zeros(2, np.float32)
        self._check(self._L.bhrt_debug_trace_mesh(self._ctx, int(mesh), n, rays.ctypes.data, a.ctypes.data, w.ctypes.data, ms.ctypes.data))
        self.debug_trace_ms = (float(ms[0]), float(ms[1]))      # device time of the binary / the wide pass
        return a, w

    # ---- caustic photon map -------------------------------------------------------------------
    def photons_emit(self, target_stored, seed=1, global_map=False):
        """BuildCausticPhotonMap's emission loop (or, global_map=True, BuildPhotonMap's); returns (stored, photon rays)."""
        stored, rays = C.c_int32(), C.c_uint64()
        fn = self._L.bhrt_photons_emit_global if global_map else self._L.bhrt_photons_emit
        self._check(fn(self._ctx, int(target_stored), int(seed), C.byref(stored), C.byref(rays)))
        return stored.value, rays.value

    def photons_buffer(self):
        ptr, n = C.c_void_p(), C.c_int32()
        self._check(self._L.bhrt_photons_buffer(self._ctx, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def photons_get(self):
        _, n = self.photons_buffer()
        buf = np.empty((max(n, 1), 24), np.uint8)
        got = C.c_int32()
        self._check(self._L.bhrt_photons_get(self._ctx, buf.ctypes.data, n, C.byref(got)))
        return buf[:got.value]

    def photons_set(self, records, from_device=False, count=None):
        if from_device:
            self._check(self._L.bhrt_photons_set(self._ctx, C.c_void_p(int(records)), int(count), 1))
        else:
            records = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, 24)
            self._check(self._L.bhrt_photons_set(self._ctx, records.ctypes.data, records.shape[0], 0))

    def photons_build(self, total_stored):
        self._check(self._L.bhrt_photons_build(self._ctx, int(total_stored)))

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.bhrt_destroy(self._ctx)
            self._ctx = C.c_void_p()


def dist_unique_id():
    """bhrt_dist_unique_id: the 128-byte rendezvous id rank 0 makes (bytes)."""
    L = device_lib()
    buf = (C.c_uint8 * 128)()
    rc = L.bhrt_dist_unique_id(buf)
    if rc != 0:
        raise BhrtError("libbhrt error %d: %s" % (rc, L.bhrt_last_error().decode()))
    return bytes(buf)


def dist_shards(kind, n, rank, world):
    """The C++ host's partitions (bhrt_dist_shard_*): 'rows' -> (begin, end, stride), 'samples' -> (begin, end),
    'photons' -> quota.  Pure host arithmetic: works without a GPU."""
    L = device_lib()
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    if kind == "rows":
        rc = L.bhrt_dist_shard_rows(n, rank, world, C.byref(a), C.byref(b), C.byref(c))
        res = (a.value, b.value, c.value)
    elif kind == "samples":
        rc = L.bhrt_dist_shard_samples(n, rank, world, C.byref(a), C.byref(b))
        res = (a.value, b.value)
    else:
        res = L.bhrt_dist_shard_photons(n, rank, world)
        rc = 0 if res >= 0 else -1
    if rc != 0:
        raise ValueError(L.bhrt_last_error().decode())
    return res


class Dist:
    """bhrt_dist (include/bhrt_dist.h): the C++ multi-GPU host -- one per rank, over that rank's Renderer.  The ranks
    exchange over NCCL inside the library; Python only carries the rendezvous id between the processes."""

    def __init__(self, renderer, rank, world, unique_id):
        self.r = renderer
        self._L = renderer._L
        self._d = C.c_void_p()
        self.rank, self.world = rank, world
        idbuf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        renderer._check(self._L.bhrt_dist_create(renderer._ctx, rank, world, idbuf, C.byref(self._d)))

    def render(self, params=None, dst=0, want=("rgb8", "linear", "hit_id", "z"), **kw):
        p = params or self.r.params(**kw)
        o, bufs = self.r._alloc_outputs(want if self.rank == dst else ())
        self.r._check(self._L.bhrt_dist_render(self._d, C.byref(p), dst, C.byref(o)))
        bufs.update(rays_closest=int(o.rays_closest), rays_shadow=int(o.rays_shadow), kernel_ms=float(o.kernel_ms))
        return bufs

    def render_into(self, params, dst=0, rgb8=None):
        """bhrt_dist_render with a caller-owned HOST RGB8 buffer (raw address, e.g. pinned memory) on rank dst."""
        o = Outputs()
        o.rgb8 = rgb8 if self.rank == dst else None
        self.r._check(self._L.bhrt_dist_render(self._d, C.byref(params), dst, C.byref(o)))
        return dict(rays_closest=int(o.rays_closest), rays_shadow=int(o.rays_shadow), kernel_ms=float(o.kernel_ms))

    def photons(self, total_stored, seed=1):
        stored, rays = C.c_int32(), C.c_uint64()
        self.r._check(self._L.bhrt_dist_photons(self._d, int(total_stored), int(seed), C.byref(stored), C.byref(rays)))
        return stored.value, rays.value

    def close(self):
        if getattr(self, "_d", None) and self._d.value:
            self._L.bhrt_dist_destroy(self._d)
            self._d = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
