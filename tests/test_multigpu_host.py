"""Multi-GPU host logic (bhraytracer_b200/multigpu.py) on CPU: world_size 2 under `gloo`.

The orchestration (row / sample-range / photon sharding, all-reduce of the accumulation buffers,
all-gather of the photon records) is the code the GPU ranks run; only the engine differs: here it is
`EmuEngine`, which renders a rank's share with tests/emu (the device code compiled by g++) instead of
libbhrt.so.  Checks: N ranks assemble exactly the frame one process renders."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from aovio import read_aov
from conftest import EMU_BIN, ROOT

W, H = 96, 72


class EmuEngine:
    def __init__(self, scene, tmp, tag):
        import torch
        self.torch = torch
        self.scene, self.tmp, self.tag = scene, tmp, tag
        self.width, self.height = W, H
        self.accum = torch.zeros(W * H * 4, dtype=torch.float32)
        self.hit_id = torch.zeros(W * H, dtype=torch.int32)
        self.z = torch.zeros(W * H, dtype=torch.float32)
        self.photon_file, self.photon_total = None, 1
        self.records = torch.zeros(0, dtype=torch.uint8)

    def _run(self, prefix, *args):
        out = subprocess.check_output([EMU_BIN, self.scene, prefix, "--size", str(W), str(H)] + [str(a) for a in args], cwd=ROOT, stderr=subprocess.DEVNULL)
        return json.loads(out.decode().splitlines()[-1])

    def render_share(self, mode=0, spp=1, gi_depth=3, tir_budget=16, seed=1, sample_range=None, rows=None, row_stride=1,
                     use_photon_map=False, photon_k=1000, photon_radius=0.5):
        prefix = os.path.join(self.tmp, "share_" + self.tag)
        args = ["--gi", gi_depth, "--seed", seed]
        args += ["--det"] if mode == 1 else ["--spp", spp]
        if rows:
            args += ["--rows", rows[0], rows[1], row_stride]
        if sample_range:
            if sample_range[0] == sample_range[1]:        # empty share: zero buffers, nothing traced
                self.accum.zero_(); self.hit_id.zero_(); self.z.zero_()
                return dict(rays_closest=0, rays_shadow=0, kernel_ms=0.0)
            args += ["--samples", sample_range[0], sample_range[1]]
        if use_photon_map:
            args += ["--photons", self.photon_file, self.photon_total, photon_k, photon_radius]
        info = self._run(prefix, *args)
        a = read_aov(prefix + ".aov")
        self.accum.copy_(self.torch.from_numpy(np.fromfile(prefix + ".acc", dtype=np.float32)))
        self.hit_id.copy_(self.torch.from_numpy(a["hit_id"].reshape(-1).astype(np.int32)))
        self.z.copy_(self.torch.from_numpy(a["z"].reshape(-1)))
        return dict(rays_closest=info["rays_closest"], rays_shadow=info["rays_shadow"], kernel_ms=info["ms"])

    def reduce_buffers(self, det):
        return [self.accum, self.hit_id, self.z]

    def resolve(self, total_spp, want):
        acc = self.accum.numpy().reshape(H, W, 4)
        out = dict(accum=acc.copy(), linear=acc[:, :, :3] / np.float32(total_spp))
        out.update(hit_id=self.hit_id.numpy().reshape(H, W).copy(), z=self.z.numpy().reshape(H, W).copy())
        return out if want else {}

    # caustic photons
    def photons_emit(self, quota, seed):
        f = os.path.join(self.tmp, "emit_%s.dat" % self.tag)
        info = self._run(os.path.join(self.tmp, "e_" + self.tag), "--emit", quota, f, "--seed", seed)
        self.records = self.torch.from_numpy(np.fromfile(f, dtype=np.uint8))
        return info["photons"], info["photon_rays"]

    def photon_records(self):
        return self.records

    def photons_set(self, records, count):
        self.records = records.clone()
        self.photon_file = os.path.join(self.tmp, "all_%s.dat" % self.tag)
        self.records.numpy()[:count * 24].tofile(self.photon_file)

    def photons_build(self, total_stored):
        if self.photon_file is None:
            self.photons_set(self.records, self.records.numel() // 24)
        self.photon_total = total_stored


def _worker(rank, world, port, tmp, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from bhraytracer_b200 import multigpu as mg
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {}
    try:
        eng = EmuEngine(os.path.join("scenes", "det_refract.xml"), tmp, "r%d" % rank)
        dr = mg.DistributedRenderer(eng)
        det = dr.render(mode=1, want=("linear",), dst=0)
        res["det_rays"] = (det["rays_closest"], det["rays_shadow"])
        if rank == 0:
            res["det_accum"], res["det_hit"], res["det_z"] = det["accum"], det["hit_id"], det["z"]
        pth = dr.render(mode=0, spp=5, seed=9, want=("linear",), dst=0)        # 5 samples over 2 ranks: 3 + 2
        res["path_rays"] = (pth["rays_closest"], pth["rays_shadow"])
        if rank == 0:
            res["path_accum"] = pth["accum"]
        # the sample-0 AOVs reach a destination rank that does not own sample 0 (they were all zero there)
        pth1 = dr.render(mode=0, spp=5, seed=9, want=("linear",), dst=1)
        if rank == 1:
            res["path_hit_dst1"], res["path_z_dst1"] = pth1["hit_id"], pth1["z"]
        # photons: sharded emission, all-gather, same full map on every rank
        eng2 = EmuEngine(os.path.join("scenes", "det_caustics.xml"), tmp, "p%d" % rank)
        dr2 = mg.DistributedRenderer(eng2)
        total, rays = dr2.build_photon_map(3001, seed=4)
        res["photon_total"], res["photon_rays"] = total, rays
        res["photon_bytes"] = eng2.records.numpy()[:total * 24].copy()
        res["photon_scale_total"] = eng2.photon_total
        pm = dr2.render(mode=1, use_photon_map=True, photon_k=100, want=("linear",), dst=0)
        if rank == 0:
            res["pm_accum"] = pm["accum"]
    finally:
        dist.barrier()
        dist.destroy_process_group()
    q.put((rank, res))


@pytest.fixture(scope="module")
def two_ranks(built, tmp_path_factory):
    import torch.multiprocessing as mp
    tmp = str(tmp_path_factory.mktemp("mg"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tmp, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in procs:
        rank, res = q.get(timeout=600)
        out[rank] = res
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return out, tmp


def _single(tmp, scene, *args):
    prefix = os.path.join(tmp, "single")
    out = subprocess.check_output([EMU_BIN, scene, prefix, "--size", str(W), str(H)] + [str(a) for a in args], cwd=ROOT, stderr=subprocess.DEVNULL)
    info = json.loads(out.decode().splitlines()[-1])
    return info, read_aov(prefix + ".aov"), np.fromfile(prefix + ".acc", dtype=np.float32).reshape(H, W, 4)


def test_shard_functions():
    sys.path.insert(0, ROOT)
    from bhraytracer_b200 import multigpu as mg
    for spp in (1, 5, 16, 255):
        for world in (1, 2, 3, 8):
            r = [mg.shard_samples(spp, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
            assert sum(mg.shard_photons(spp, k, world) for k in range(world)) == spp
    rows = sorted(j for k in range(3) for j in range(*mg.shard_rows(10, k, 3)))
    assert rows == list(range(10))
    with pytest.raises(ValueError):
        mg.shard_rows(10, 3, 3)


def test_two_ranks_deterministic_rows_assemble_the_single_process_frame(two_ranks):
    out, tmp = two_ranks
    info, aov, acc = _single(tmp, os.path.join("scenes", "det_refract.xml"), "--det")
    r0 = out[0]
    assert (r0["det_accum"].view(np.uint32) == acc.view(np.uint32)).all()       # x + 0 is exact: bit-identical
    assert (r0["det_hit"] == aov["hit_id"]).all()
    assert (r0["det_z"].view(np.uint32) == aov["z"].view(np.uint32)).all()
    assert out[0]["det_rays"] == out[1]["det_rays"] == (info["rays_closest"], info["rays_shadow"])


def test_two_ranks_sample_ranges_sum_to_the_single_process_frame(two_ranks):
    out, tmp = two_ranks
    info, aov, acc = _single(tmp, os.path.join("scenes", "det_refract.xml"), "--spp", 5, "--seed", 9)
    a = out[0]["path_accum"]
    assert (a[:, :, 3] == 5).all()                       # every pixel got all 5 samples, once
    assert np.allclose(a[:, :, :3], acc[:, :, :3], rtol=0, atol=4e-6)           # (s0+s1+s2) + (s3+s4) vs s0+..+s4
    assert out[0]["path_rays"] == out[1]["path_rays"] == (info["rays_closest"], info["rays_shadow"])
    assert (out[1]["path_hit_dst1"] == aov["hit_id"]).all()
    assert (out[1]["path_z_dst1"].view(np.uint32) == aov["z"].view(np.uint32)).all()


def test_two_ranks_photon_all_gather(two_ranks):
    out, tmp = two_ranks
    assert out[0]["photon_total"] == out[1]["photon_total"] == 3001
    assert out[0]["photon_scale_total"] == out[1]["photon_scale_total"] == 3001      # global stored count (Main.cpp:380)
    assert (out[0]["photon_bytes"] == out[1]["photon_bytes"]).all()                  # the same full map on every rank
    assert out[0]["photon_rays"] == out[1]["photon_rays"] > 3001
    # and the gathered map renders like one process with the same records
    f = os.path.join(tmp, "gathered.dat")
    out[0]["photon_bytes"].tofile(f)
    info, aov, acc = _single(tmp, os.path.join("scenes", "det_caustics.xml"), "--det", "--photons", f, 3001, 100, 0.5)
    assert (out[0]["pm_accum"].view(np.uint32) == acc.view(np.uint32)).all()
