#!/usr/bin/env python3
"""Run under torchrun (one rank per GPU): checks the multi-GPU paths on real GPUs over NCCL against a single-GPU
render done by rank 0 -- first the C++ host of libbhrt.so (include/bhrt_dist.h, `b.Dist`; torch.distributed only
carries the 128-byte rendezvous id), then the Python orchestration of bhraytracer_b200/multigpu.py that the CPU
test tier exercises under gloo.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multigpu_check.py

 * deterministic frame, interleaved rows + sum all-reduce  -> bit-identical to one GPU
 * stochastic frame, sample ranges + sum all-reduce         -> the same accumulation up to fp32 addition order
 * caustic photons, per-rank quota + all-gather             -> every rank holds the same full map; the
                                                               gather-lit frame equals one GPU rendering that map
Prints one JSON line on rank 0 and exits non-zero on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import bhraytracer_b200 as b  # noqa: E402
from bhraytracer_b200 import multigpu as mg  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.chdir(ROOT)
    import gen_assets
    gen_assets.ensure_assets()
    res = {"world": world}
    ok = True

    # ---- the C++ host (bhrt_dist_*)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(b.dist_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    uid = bytes(idt.cpu().numpy().tobytes())
    scene = b.LoadScene("scenes/det_mesh.xml", 640, 360)
    rr = b.Renderer(local)
    rr.upload(scene)
    dd = b.Dist(rr, rank, world, uid)
    c_det = dd.render(mode=b.MODE_DET, dst=0)
    c_pth = dd.render(mode=b.MODE_PATH, spp=8, seed=3, want=("linear", "hit_id", "z"), dst=world - 1)
    if rank == 0:
        single = b.Renderer(local)
        single.upload(scene)
        s_det = single.render(mode=b.MODE_DET)
        res["cxx_det_bit_identical"] = bool((c_det["linear"].view(np.uint32) == s_det["linear"].view(np.uint32)).all() and (c_det["hit_id"] == s_det["hit_id"]).all() and
                                            (c_det["z"].view(np.uint32) == s_det["z"].view(np.uint32)).all() and (c_det["rgb8"] == s_det["rgb8"]).all())
        res["cxx_det_rays_equal"] = (c_det["rays_closest"], c_det["rays_shadow"]) == (s_det["rays_closest"], s_det["rays_shadow"])
        ok &= res["cxx_det_bit_identical"] and res["cxx_det_rays_equal"]
        single.close()
    if rank == world - 1:       # the stochastic frame lands on the LAST rank, which does not own sample 0: AOVs must still be there
        single = b.Renderer(local)
        single.upload(scene)
        s_pth = single.render(mode=b.MODE_PATH, spp=8, seed=3, want=("linear", "hit_id", "z"))
        good = (float(np.abs(c_pth["linear"] - s_pth["linear"]).max()) < 2e-6 and (c_pth["hit_id"] == s_pth["hit_id"]).all() and
                (c_pth["z"].view(np.uint32) == s_pth["z"].view(np.uint32)).all() and
                (c_pth["rays_closest"] + c_pth["rays_shadow"]) == (s_pth["rays_closest"] + s_pth["rays_shadow"]))
        single.close()
        flag_p = torch.tensor([1 if good else 0], device="cuda")
    else:
        flag_p = torch.tensor([1], device="cuda")
    dist.all_reduce(flag_p, op=dist.ReduceOp.MIN)
    res["cxx_path_ok"] = bool(int(flag_p[0]))
    ok &= res["cxx_path_ok"]
    # photons through the C++ host: sharded emission, all-gather, replicated build
    scene2 = b.LoadScene("scenes/det_caustics.xml", 320, 240)
    rr.upload(scene2)
    c_total, c_rays = dd.photons(100003, seed=5)
    c_rec = rr.photons_get()
    digest = torch.tensor([int(np.frombuffer(c_rec.tobytes(), dtype=np.uint32).astype(np.uint64).sum() % (1 << 62)), c_rec.shape[0]], dtype=torch.int64, device="cuda")
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    c_same = all(bool((d == all_d[0]).all()) for d in all_d)
    c_pm = dd.render(mode=b.MODE_DET, use_photon_map=True, dst=0)
    if rank == 0:
        single = b.Renderer(local)
        single.upload(scene2)
        single.photons_set(c_rec)
        single.photons_build(c_total)
        s_pm = single.render(mode=b.MODE_DET, use_photon_map=True)
        res.update(cxx_photons_total=c_total, cxx_same_map_on_every_rank=c_same,
                   cxx_photon_frame_bit_identical=bool((c_pm["linear"].view(np.uint32) == s_pm["linear"].view(np.uint32)).all()))
        ok &= c_total == 100003 and c_same and res["cxx_photon_frame_bit_identical"]
        single.close()
    dd.close()
    rr.close()

    # ---- the Python orchestration (what tests/test_multigpu_host.py runs under gloo)
    eng = mg.GpuEngine(scene, local)
    dr = mg.DistributedRenderer(eng)
    det = dr.render(mode=b.MODE_DET, dst=0)
    pth = dr.render(mode=b.MODE_PATH, spp=8, seed=3, want=("linear",), dst=0)
    if rank == 0:
        single = b.Renderer(local)
        single.upload(scene)
        s_det = single.render(mode=b.MODE_DET)
        s_pth = single.render(mode=b.MODE_PATH, spp=8, seed=3, want=("linear",))
        res["det_bit_identical"] = bool((det["linear"].view(np.uint32) == s_det["linear"].view(np.uint32)).all() and
                                        (det["hit_id"] == s_det["hit_id"]).all() and (det["z"].view(np.uint32) == s_det["z"].view(np.uint32)).all() and
                                        (det["rgb8"] == s_det["rgb8"]).all())
        res["det_rays_equal"] = (det["rays_closest"], det["rays_shadow"]) == (s_det["rays_closest"], s_det["rays_shadow"])
        res["path_max_abs_diff"] = float(np.abs(pth["linear"] - s_pth["linear"]).max())
        res["path_rays_equal"] = (pth["rays_closest"] + pth["rays_shadow"]) == (s_pth["rays_closest"] + s_pth["rays_shadow"])
        ok &= res["det_bit_identical"] and res["det_rays_equal"] and res["path_max_abs_diff"] < 2e-6 and res["path_rays_equal"]
        single.close()
    eng.close()

    scene2 = b.LoadScene("scenes/det_caustics.xml", 320, 240)
    eng2 = mg.GpuEngine(scene2, local)
    dr2 = mg.DistributedRenderer(eng2)
    total, rays = dr2.build_photon_map(100003, seed=5)
    rec = eng2.r.photons_get()
    digest = torch.tensor([int(np.frombuffer(rec.tobytes(), dtype=np.uint32).astype(np.uint64).sum() % (1 << 62)), rec.shape[0]], dtype=torch.int64, device="cuda")
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same_map = all(bool((d == all_d[0]).all()) for d in all_d)
    pm = dr2.render(mode=b.MODE_DET, use_photon_map=True, dst=0)
    if rank == 0:
        single = b.Renderer(local)
        single.upload(scene2)
        single.photons_set(rec)
        single.photons_build(total)
        s_pm = single.render(mode=b.MODE_DET, use_photon_map=True)
        res.update(photons_total=total, photon_rays=rays, same_map_on_every_rank=same_map,
                   photon_frame_bit_identical=bool((pm["linear"].view(np.uint32) == s_pm["linear"].view(np.uint32)).all()))
        ok &= total == 100003 and same_map and res["photon_frame_bit_identical"]
        single.close()
    eng2.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0:
        res["ok"] = bool(ok)
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == "__main__":
    main()
