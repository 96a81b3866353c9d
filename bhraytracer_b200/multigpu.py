"""Multi-GPU host logic: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch) for
the plumbing.  The reference is a single process with an OpenMP loop over image columns
(reference Main.cpp:204-234); nothing here has a counterpart there.  SURVEY.md 8(e):

* deterministic frames (D-mode)  -> screen rows interleaved across ranks (rank r renders rows
  r, r+N, r+2N, ...: balanced without knowing the scene), no collective during the render; the
  per-rank buffers hold zeros outside the rank's rows, so ONE sum all-reduce of the float4
  accumulation buffer (and of the hit-id / z AOVs) assembles the frame, bit-identical to a
  single-GPU render because x + 0 is exact;
* stochastic frames (H-mode)     -> sample ranges [s0, s1) of every pixel per rank (the counter-based
  RNG is keyed by (seed, pixel, sample), so the union of the ranges IS the single-GPU sample set),
  then a sum all-reduce of the accumulation buffers, then resolve with the total spp;
* caustic photons                -> per-rank emission quota, all-gather of the 24-byte records, every
  rank builds the full gather structure (`build_photon_map`).

The orchestration is written against a small engine interface so the CPU-only test tier can run it
under `gloo` with world_size 2 (tests/test_multigpu_host.py); on a GPU the engine is
:class:`GpuEngine` (a :class:`bhraytracer_b200.Renderer` plus zero-copy torch views of its device
buffers).
"""
import numpy as np

from . import MODE_DET, MODE_PATH, Renderer


def shard_rows(height, rank, world):
    """Interleaved rows of `rank`: (row_begin, row_end, row_stride)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return rank, height, world


def shard_samples(spp, rank, world):
    """Contiguous, balanced sample range [begin, end) of `rank`; ranges tile [0, spp) exactly."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(spp, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_photons(total, rank, world):
    """Stored-photon quota of `rank` (sums to `total`)."""
    b, e = shard_samples(total, rank, world)
    return e - b


class _CudaArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v3)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class GpuEngine:
    """The real engine: libbhrt.so on this rank's GPU, its buffers seen as torch tensors."""

    def __init__(self, scene, device):
        import torch
        self.torch = torch
        self.r = Renderer(device)
        self.r.upload(scene)
        self.width, self.height = scene.width, scene.height
        self.stream = torch.cuda.Stream(device=device)
        self.r.set_stream(self.stream.cuda_stream)
        n = self.width * self.height
        ptr, nbytes = self.r.accum_buffer()
        dev = torch.device("cuda", device)
        self.accum = torch.as_tensor(_CudaArray(ptr, nbytes // 4, "<f4"), device=dev)
        hid, z = self.r.aov_buffers()
        self.hit_id = torch.as_tensor(_CudaArray(hid, n, "<i4"), device=dev)
        self.z = torch.as_tensor(_CudaArray(z, n, "<f4"), device=dev)

    def params(self, **kw):
        return self.r.params(**kw)

    def render_share(self, **kw):
        return self.r.render_device(self.r.params(**kw))

    def reduce_buffers(self, det):
        # hit_id / z are the AOVs of sample 0: in the stochastic mode only the rank that owns sample 0 writes
        # them (zeros elsewhere, render.cu), in the deterministic mode every rank writes its own rows -- a sum
        # all-reduce assembles them on every rank either way (x + 0 is exact)
        return [self.accum, self.hit_id, self.z]

    def resolve(self, total_spp, want):
        return self.r.resolve(total_spp, want=want)

    # ---- caustic photon map
    def photons_emit(self, quota, seed):
        return self.r.photons_emit(quota, seed)

    def photon_records(self):
        """uint8 view [count * 24] of this rank's records in device memory (zero-copy)."""
        ptr, n = self.r.photons_buffer()
        dev = self.torch.device("cuda", self.r.device)
        if n == 0:
            return self.torch.zeros(0, dtype=self.torch.uint8, device=dev)
        return self.torch.as_tensor(_CudaArray(ptr, n * 24, "|u1"), device=dev)

    def photons_set(self, records, count):
        self.r.photons_set(records.data_ptr(), from_device=True, count=count)

    def photons_build(self, total_stored):
        self.r.photons_build(total_stored)

    def close(self):
        self.r.close()


class DistributedRenderer:
    """BeginRender() across the ranks of a process group.  Every rank calls `render` with the same
    arguments; every rank ends up with the full frame (accumulation buffer and the sample-0 AOVs hit_id / z)
    in its device buffers, rank `dst` gets host copies of the requested outputs (other ranks get `{}` plus
    the counters).  Deterministic frames are bit-identical whatever the world size; stochastic frames hold
    the same SAMPLES whatever the world size, but the per-rank partial sums and NCCL's reduction change the
    fp32 addition order, so the pixel sums agree to rounding (a few 1e-7), not in bits."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def _allreduce(self, tensors):
        if self.world == 1:
            return
        stream = getattr(self.engine, "stream", None)
        if stream is not None:
            with self.engine.torch.cuda.stream(stream):
                for t in tensors:
                    self.dist.all_reduce(t, group=self.group)
            stream.synchronize()
        else:
            for t in tensors:
                self.dist.all_reduce(t, group=self.group)

    def render(self, mode=MODE_PATH, spp=32, gi_depth=3, tir_budget=16, seed=1, want=("rgb8", "linear", "hit_id", "z"), dst=0,
               use_photon_map=False, photon_k=1000, photon_radius=0.5):
        det = mode == MODE_DET
        common = dict(mode=mode, gi_depth=gi_depth, tir_budget=tir_budget, seed=seed,
                      use_photon_map=use_photon_map, photon_k=photon_k, photon_radius=photon_radius)
        if det:
            rb, re_, rs = shard_rows(self.engine.height, self.rank, self.world)
            stats = self.engine.render_share(spp=1, rows=(rb, re_), row_stride=rs, **common)
            total = 1
        else:
            sb, se = shard_samples(spp, self.rank, self.world)
            # a rank with an empty range still takes part in the collective with a zero buffer
            stats = self.engine.render_share(spp=spp, sample_range=(sb, se), **common) if se > sb else self._empty_share(spp, common)
            total = spp
        self._allreduce(self.engine.reduce_buffers(det))
        out = self.engine.resolve(total, want if self.rank == dst else ())
        counts = np.array([stats["rays_closest"], stats["rays_shadow"]], dtype=np.int64)
        if self.world > 1:
            import torch
            t = torch.from_numpy(counts)
            dev = getattr(self.engine, "accum", None)
            if dev is not None and dev.is_cuda:
                t = t.to(dev.device)
            self.dist.all_reduce(t, group=self.group)
            counts = t.cpu().numpy()
        out = dict(out)
        out.update(rays_closest=int(counts[0]), rays_shadow=int(counts[1]), kernel_ms=stats["kernel_ms"])
        return out

    def _empty_share(self, spp, common):
        # sample_range (spp, spp) is an empty, valid range: the library clears the buffers and launches nothing
        return self.engine.render_share(spp=spp, sample_range=(spp, spp), **common)

    # ---- caustic photon map (reference Main.cpp:342-386, serial there) ---------------------------
    def build_photon_map(self, total_stored, seed=1):
        """Sharded emission -> all-gather of the 24-byte records -> every rank builds the full
        gather structure.  Power scale uses the GLOBAL stored count (reference Main.cpp:380)."""
        import torch
        eng = self.engine
        quota = shard_photons(total_stored, self.rank, self.world)
        stored, rays = eng.photons_emit(quota, seed + 7919 * self.rank)
        if self.world == 1:
            eng.photons_build(stored)
            return stored, rays
        # Everything between the emission and photons_set runs ON THE ENGINE'S STREAM: photons_set copies
        # out of `allrec` with a cudaMemcpyAsync on that stream, so the pad / all_gather / cat that produce
        # it must be ordered before it (on torch's default stream nothing would order them).
        import contextlib
        stream = getattr(eng, "stream", None)
        ctx = eng.torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
        with ctx:
            rec = eng.photon_records()                     # uint8 tensor [stored * 24] on the engine's device
            cnt = torch.tensor([stored, rays], dtype=torch.int64, device=rec.device)
            cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
            self.dist.all_gather(cnts, cnt, group=self.group)
            counts = [int(c[0]) for c in cnts]
            mx = max(counts)
            pad = torch.zeros(mx * 24, dtype=torch.uint8, device=rec.device)
            pad[:stored * 24] = rec[:stored * 24]
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            self.dist.all_gather(parts, pad, group=self.group)
            allrec = torch.cat([p[:c * 24] for p, c in zip(parts, counts)])
            total = sum(counts)
            eng.photons_set(allrec, total)                 # synchronises the stream before it returns
        eng.photons_build(total)
        del allrec, parts, pad
        return total, int(sum(int(c[1]) for c in cnts))
