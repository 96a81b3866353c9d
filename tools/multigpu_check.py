#!/usr/bin/env python3
"""Run under torchrun (one rank per GPU): checks the multi-GPU paths of bhraytracer_b200/multigpu.py on
real GPUs over NCCL against a single-GPU render done by rank 0.

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

    scene = b.LoadScene("scenes/det_mesh.xml", 640, 360)
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
