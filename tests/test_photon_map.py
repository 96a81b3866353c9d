"""Caustic photon map: emission, the 24-byte record, the gather structure
and the k-nearest irradiance estimate, against the reference built with -DUSE_PhotonMap
(oracle/_ref/bhrt_ref_pm; fixture tests/golden/det_caustics_pm.npz made by tests/golden/make_golden.py).

 * gather (P3/P4): the fixture holds the REFERENCE's own photons (its .dat dump, powers already
   scaled, Main.cpp:380-385) and its D-mode render with the gather term (k = 1000, r = 0.5).
   Gathering the same photons must reproduce that image to <= 1/255 -- deterministic.
 * emission (P1/P2): stochastic, so distributions are compared with the reference's (300 k photons):
   where the caustic photons land, their mean power and direction signs.
CPU tier = the device code compiled by tests/emu; GPU tier = the kernels through the C ABI.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from aovio import read_aov
from conftest import EMU_BIN, ROOT
from photon_util import as_records, photon_stats

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "det_caustics_pm.npz"))
SCENE = os.path.join("scenes", "det_caustics.xml")
W, H, K, RADIUS = int(GOLD["width"]), int(GOLD["height"]), int(GOLD["k"]), float(GOLD["radius"])


def check_gather_image(rgb8, hit_id):
    assert (hit_id == GOLD["hit_id"]).all(), "primary hit ids differ"
    d = np.abs(rgb8.astype(np.int32) - GOLD["rgb8"].astype(np.int32))
    assert d.max() <= 1, "8-bit error %d > 1/255 on %d pixels" % (d.max(), int((d.max(axis=2) > 1).sum()))


def check_emission_stats(raw, n_target, prefix=""):
    """Distribution of OUR photons (unscaled power) vs the reference's 300 k-photon statistics (prefix "global_": of the
    global map, BuildPhotonMap)."""
    rec = as_records(raw)
    assert len(rec) == n_target
    st = photon_stats(rec, 1.0)
    G = GOLD
    if prefix:
        G = {k[len(prefix):]: GOLD[k] for k in GOLD.files if k.startswith(prefix)}
        return _compare_stats(st, G, len(rec), caustic=False)
    return _compare_stats(st, GOLD, len(rec), caustic=True)


def _compare_stats(st, GOLD, n_rec, caustic):
    rec = range(n_rec)
    n_ref = float(GOLD["stats_photons"])
    # binomial standard error of a bin share with both sample sizes, 5 sigma + a small floor
    def tol(p):
        return 5.0 * np.sqrt(p * (1 - p) * (1.0 / len(rec) + 1.0 / n_ref)) + 2e-4
    assert abs(st["floor_frac"] - float(GOLD["floor_frac"])) < tol(float(GOLD["floor_frac"]))
    assert abs(st["dirz_neg_frac"] - float(GOLD["dirz_neg_frac"])) < tol(float(GOLD["dirz_neg_frac"]))
    assert abs(st["mean_power"] / float(GOLD["mean_power"]) - 1.0) < 0.02
    ref_axis = GOLD["axis_hist"].astype(np.float64)
    assert (np.abs(st["axis_hist"] - ref_axis) < tol(ref_axis)).all(), "per-axis position histograms differ"
    ref_floor = GOLD["floor_hist"].astype(np.float64)
    assert (np.abs(st["floor_hist"] - ref_floor) < tol(ref_floor)).all(), "caustic footprint on the floor differs"
    # the caustic must actually be a concentration: the busiest floor cell holds far more than the uniform share
    if caustic:
        assert st["floor_hist"].max() > 20.0 / (24 * 24)


def check_record_format(raw):
    """Photon::SetPower / SetDirection (cyPhotonMap.h:172-191): power = max channel, 8-bit colour
    ratio with at least one channel at 255, 16-bit fixed-point direction inside the unit disk."""
    rec = as_records(raw)
    assert (rec["power"] > 0).all()
    assert (rec["rgb"].max(axis=1) == 255).all()
    dx, dy = rec["dx"].astype(np.float64) / 0x7FFF, rec["dy"].astype(np.float64) / 0x7FFF
    assert (dx * dx + dy * dy <= 1.0 + 1e-4).all()
    assert ((rec["flag"] & ~np.uint8(0x8)) == 0).all()      # only the dirZ sign bit is set before the kd-tree build


# ---- CPU tier -----------------------------------------------------------------------------------------
def test_emu_gather_matches_reference(built, tmp_path):
    ph = tmp_path / "ref_photons.dat"
    GOLD["photons"].tofile(str(ph))
    prefix = str(tmp_path / "pm")
    subprocess.check_output([EMU_BIN, SCENE, prefix, "--det", "--size", str(W), str(H), "--photons", str(ph), "1", str(K), str(RADIUS)],
                            cwd=ROOT, stderr=subprocess.DEVNULL)
    a = read_aov(prefix + ".aov")
    check_gather_image(a["rgb8"], a["hit_id"])
    # and the photon map does matter: without it thousands of pixels differ
    subprocess.check_output([EMU_BIN, SCENE, prefix + "0", "--det", "--size", str(W), str(H)], cwd=ROOT, stderr=subprocess.DEVNULL)
    b = read_aov(prefix + "0.aov")
    assert (np.abs(b["rgb8"].astype(int) - GOLD["rgb8"].astype(int)).max(axis=2) > 1).sum() > 1000


def decode_dirs(rec):
    """Photon::GetDirection (cyPhotonMap.h:193-214): integer square root of 0x3FFF0001 - dirX^2
    (its `+ dirY - dirY` slip drops dirY^2), sign from planeAndDirZ bit 3."""
    dxi, dyi = rec["dx"].astype(np.int64), rec["dy"].astype(np.int64)
    z2 = 0x3FFF0001 - np.minimum(dxi * dxi, 0x3FFF0001)
    dz = np.floor(np.sqrt(z2.astype(np.float64))).astype(np.int64)
    dz = np.where((dz + 1) * (dz + 1) <= z2, dz + 1, dz)
    dz = np.where(dz * dz > z2, dz - 1, dz)
    f = np.float32(0x7FFF)
    return np.stack([dxi.astype(np.float32) / f, dyi.astype(np.float32) / f,
                     np.where(rec["flag"] & 8, -1, 1).astype(np.float32) * (dz.astype(np.float32) / f)], axis=1)


@pytest.mark.parametrize("k,radius", [(1000, 0.5), (50, 0.5), (7, 0.5), (1000, 0.05), (300, 2.0)])
def test_emu_irradiance_estimate_vs_brute_force(built, tmp_path, k, radius):
    """PhotonMap::EstimateIrradiance<k> + LocatePhotons (cyPhotonMap.h:332-382,421-498) with the
    constant filter, restated as a brute-force numpy selection over the reference's own photons:
    the k nearest photons within `radius` with dir.N < 0; area = pi r^2 with r^2 = radius^2 when
    at most k qualified, else the k-th smallest squared distance.  The grid/selection estimator
    must return the same irradiance and mean direction."""
    rec = as_records(GOLD["photons"])
    pos, dirs = rec["pos"].astype(np.float32), decode_dirs(rec)
    maxp = rec["power"].astype(np.float32)
    power = maxp[:, None] * (rec["rgb"].astype(np.float32) / np.float32(255.0))
    rng = np.random.default_rng(5)
    # queries: photon positions themselves (dense caustic), jittered copies, and empty space
    pick = rng.choice(len(rec), 160, replace=False)
    qpos = np.concatenate([pos[pick[:80]], pos[pick[80:]] + rng.normal(0, 0.2, (80, 3)).astype(np.float32),
                           rng.uniform(-14, 14, (40, 3)).astype(np.float32)])
    qn = np.tile(np.array([0, 0, 1], np.float32), (len(qpos), 1))
    qn[1::3] = -dirs[pick[0]] / np.linalg.norm(dirs[pick[0]])
    qn[2::3] = np.array([0, 1, 0], np.float32)
    q = np.concatenate([qpos, qn], axis=1).astype(np.float32)
    ph, qi, qo = tmp_path / "p.dat", tmp_path / "q.bin", tmp_path / "r.bin"
    GOLD["photons"].tofile(str(ph))
    q.tofile(str(qi))
    subprocess.check_output([EMU_BIN, SCENE, str(tmp_path / "x"), "--det", "--photons", str(ph), "1", str(k), str(radius),
                             "--queries", str(qi), str(qo)], cwd=ROOT, stderr=subprocess.DEVNULL)
    both = np.fromfile(str(qo), dtype=np.float32).reshape(-1, 12)
    res, scan = both[:, :6], both[:, 6:]
    # (1) the grid + radix-select estimator against a plain scan + sort of the same records in C++
    assert np.allclose(res[:, :3], scan[:, :3], rtol=2e-5, atol=1e-9), np.abs(res[:, :3] - scan[:, :3]).max()
    assert np.allclose(res[:, 3:], scan[:, 3:], atol=2e-5)
    # (2) and both against the numpy restatement of the reference's estimate
    n_multi = 0
    for i in range(len(q)):
        df = pos - q[i, :3]                                 # float32, then x*x + y*y + z*z left to right like the C++
        d2 = df[:, 0] * df[:, 0] + df[:, 1] * df[:, 1] + df[:, 2] * df[:, 2]
        dn = dirs[:, 0] * q[i, 3] + dirs[:, 1] * q[i, 4] + dirs[:, 2] * q[i, 5]
        ok = (d2 < np.float32(radius) * np.float32(radius)) & (dn < 0)
        idx = np.nonzero(ok)[0]
        if len(idx) == 0:
            assert (res[i, :3] == 0).all()
            continue
        if len(idx) > k:
            order = np.argsort(d2[idx], kind="stable")
            kth = d2[idx][order[k - 1]]
            if k < len(idx) and d2[idx][order[k]] == kth:
                continue                                    # a tie at the k-th distance: the heap's choice is order dependent
            idx = idx[order[:k]]
            r2 = kth
            n_multi += 1
        else:
            r2 = np.float32(radius * radius)
        want = power[idx].astype(np.float64).sum(axis=0) / (np.pi * float(r2))
        assert np.allclose(res[i, :3], want, rtol=2e-4, atol=1e-7), (i, len(idx), res[i, :3], want)
        wdir = (dirs[idx].astype(np.float64) * maxp[idx, None]).sum(axis=0)
        wdir /= np.linalg.norm(wdir)
        assert np.allclose(res[i, 3:], wdir, atol=2e-4)
    if k <= 50 and radius >= 0.5:
        assert n_multi > 20         # the k-th-distance selection path was really exercised


def test_emu_emission_statistics(built, tmp_path):
    out = tmp_path / "emit.dat"
    n = 60000
    info = json.loads(subprocess.check_output([EMU_BIN, SCENE, str(tmp_path / "x"), "--emit", str(n), str(out), "--seed", "11"],
                                              cwd=ROOT, stderr=subprocess.DEVNULL).decode().splitlines()[-1])
    raw = np.fromfile(str(out), dtype=np.uint8).reshape(-1, 24)
    assert info["photons"] == n and info["paths"] > n
    check_record_format(raw)
    check_emission_stats(raw, n)


# ---- GPU tier -----------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def renderer(built):
    import bhraytracer_b200 as b
    r = b.Renderer(0)
    yield r
    r.close()


@pytest.mark.gpu
def test_gpu_gather_matches_reference(renderer):
    import bhraytracer_b200 as b
    scene = b.LoadScene(SCENE, W, H)
    renderer.upload(scene)
    renderer.photons_set(GOLD["photons"])
    renderer.photons_build(1)            # the reference's dump is already scaled by 1/N (Main.cpp:380)
    out = renderer.render(mode=b.MODE_DET, use_photon_map=True, photon_k=K, photon_radius=RADIUS)
    check_gather_image(out["rgb8"], out["hit_id"])
    # round trip of the record buffer (P2: the .dat wire format)
    back = renderer.photons_get()
    assert (back == GOLD["photons"]).all()


@pytest.mark.gpu
def test_gpu_gather_equals_emulated_gather(renderer, tmp_path):
    """Device-built grid (radix sort) vs host-built grid (std::sort), other k / radius: same linear image."""
    import bhraytracer_b200 as b
    scene = b.LoadScene(SCENE, W, H)
    renderer.upload(scene)
    renderer.photons_set(GOLD["photons"])
    renderer.photons_build(1)
    ph = tmp_path / "p.dat"
    GOLD["photons"].tofile(str(ph))
    for k, r in ((50, 0.5), (1000, 0.05), (300, 1.0)):
        out = renderer.render(mode=b.MODE_DET, use_photon_map=True, photon_k=k, photon_radius=r)
        prefix = str(tmp_path / ("e_%d_%g" % (k, r)))
        subprocess.check_output([EMU_BIN, SCENE, prefix, "--det", "--size", str(W), str(H), "--photons", str(ph), "1", str(k), str(r)],
                                cwd=ROOT, stderr=subprocess.DEVNULL)
        e = read_aov(prefix + ".aov")
        d = np.abs(out["rgb8"].astype(int) - e["rgb8"].astype(int))
        assert d.max() <= 1, "k=%d r=%g: GPU and emulated gather differ by %d" % (k, r, d.max())


@pytest.mark.gpu
def test_gpu_emission_statistics(renderer):
    import bhraytracer_b200 as b
    scene = b.LoadScene(SCENE, W, H)
    renderer.upload(scene)
    n = 200000
    stored, rays = renderer.photons_emit(n, seed=7)
    assert stored == n and rays > n
    raw = renderer.photons_get()
    check_record_format(raw)
    check_emission_stats(raw, n)
    # emission is keyed by (seed, path index): the same call reproduces the same SET of photons
    renderer.photons_emit(n, seed=7)
    again = renderer.photons_get()
    a = set(np.ascontiguousarray(raw).view(np.dtype((np.void, 24))).reshape(-1).tolist())
    c = set(np.ascontiguousarray(again).view(np.dtype((np.void, 24))).reshape(-1).tolist())
    assert len(a & c) > 0.97 * n        # (the paths in flight when the map filled up race for the last slots)


def test_global_photon_map_statistics_on_the_cpu_build(built, tmp_path):
    """The device code's global-map path (photon_advance_global) through the CPU harness against the reference's statistics."""
    f = str(tmp_path / "g.dat")
    subprocess.check_output([EMU_BIN, SCENE, str(tmp_path / "e"), "--size", "16", "12", "--emit-global", "120000", f, "--seed", "3"], cwd=ROOT, stderr=subprocess.DEVNULL)
    raw = np.fromfile(f, dtype=np.uint8).reshape(-1, 24)
    check_record_format(raw)
    check_emission_stats(raw, 120000, prefix="global_")


@pytest.mark.gpu
def test_gpu_global_photon_map_statistics(renderer):
    """the GLOBAL photon map (BuildPhotonMap / TracePhotonRay, Main.cpp:251-317; MtlBlinn::RandomPhotonBounce,
    MtlBlinn.cpp:140-202) against the reference's own 300 k-photon run of that code (oracle patch O-7)."""
    import bhraytracer_b200 as b
    scene = b.LoadScene(SCENE, W, H)
    renderer.upload(scene)
    n = 200000
    stored, rays = renderer.photons_emit(n, seed=11, global_map=True)
    assert stored == n and rays > n
    raw = renderer.photons_get()
    check_record_format(raw)
    check_emission_stats(raw, n, prefix="global_")
    # and it is a different map from the caustic one: photons everywhere, not only under the glass
    assert as_records(raw)["pos"][:, 2].max() > 5.0


@pytest.mark.gpu
def test_gpu_emit_build_render_caustic(renderer):
    """End to end on the GPU: emit -> build -> stochastic render with the gather term brightens the
    floor under the glass sphere like the reference's photon-mapped render does."""
    import bhraytracer_b200 as b
    scene = b.LoadScene(SCENE, W, H)
    renderer.upload(scene)
    stored, _ = renderer.photons_emit(40000, seed=3)
    renderer.photons_build(stored)
    with_pm = renderer.render(mode=b.MODE_DET, use_photon_map=True)
    without = renderer.render(mode=b.MODE_DET, use_photon_map=False)
    ref = GOLD["rgb8"].astype(np.float64)
    # same photon count as the fixture: our caustic image is statistically the reference's
    err_with = np.abs(with_pm["rgb8"].astype(np.float64) - ref).mean()
    err_without = np.abs(without["rgb8"].astype(np.float64) - ref).mean()
    assert err_with < 0.5 * err_without
    gain = with_pm["linear"].sum() - without["linear"].sum()
    ref_gain_8bit = (ref.sum() - without["rgb8"].astype(np.float64).sum())
    assert gain > 0 and ref_gain_8bit > 0


@pytest.mark.gpu
def test_cli_photon_dat_round_trip(built, tmp_path):
    """bhrt_render --photons-out writes the reference's on-disk map (24-byte records, scaled powers,
    Main.cpp:380-385); --photons-in renders with such a file -- here the REFERENCE's own dump from the
    fixture, which must give the reference's image."""
    cli = os.path.join(ROOT, "bhraytracer_b200", "lib", "bhrt_render")
    ref_dat = tmp_path / "ref.dat"
    GOLD["photons"].tofile(str(ref_dat))
    subprocess.check_output([cli, SCENE, str(tmp_path / "in"), "--det", "--size", str(W), str(H), "--photons-in", str(ref_dat), "--aov"], cwd=ROOT)
    a = read_aov(str(tmp_path / "in.aov"))
    check_gather_image(a["rgb8"], a["hit_id"])
    out_dat = tmp_path / "ours.dat"
    info = json.loads(subprocess.check_output([cli, SCENE, str(tmp_path / "out"), "--det", "--size", str(W), str(H), "--photons", "20000",
                                               "--photons-out", str(out_dat)], cwd=ROOT).decode().splitlines()[-1])
    rec = as_records(np.fromfile(str(out_dat), dtype=np.uint8).reshape(-1, 24))
    assert len(rec) == info["photons_stored"] == 20000
    # scaled like the reference's file: mean power * N is the light's emitted power scale
    assert abs(float(rec["power"].astype(np.float64).sum()) / float(GOLD["mean_power"]) - 1.0) < 0.05


@pytest.mark.gpu
def test_gpu_photon_mapped_frame_loses_no_sample(renderer):
    """Lanes parked for a caustic gather must all come back: every pixel's sample count equals spp (a
    frame much smaller than the lane pools, so the pools drain with lanes still parked)."""
    import torch
    import bhraytracer_b200 as b
    scene = b.LoadScene(SCENE, 53, 37)
    renderer.upload(scene)
    renderer.photons_set(GOLD["photons"])
    renderer.photons_build(1)
    renderer.render_device(renderer.params(mode=b.MODE_PATH, spp=3, seed=2, use_photon_map=True, photon_k=100))
    ptr, nbytes = renderer.accum_buffer()

    class _A:
        __cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
    acc = torch.as_tensor(_A(), device="cuda").cpu().numpy().reshape(37, 53, 4)
    assert (acc[:, :, 3] == 3).all()
    assert np.isfinite(acc).all()


# ---- the schedules of the gather: inside wf_render, split (the default), split with every query that scans
# more than a handful of records handed to the CTA-wide routine -- same photons, same queries, same frame
def _render_with_env(env, mode, **kw):
    import bhraytracer_b200 as b
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        r = b.Renderer(0)               # (the knobs are read by bhrt_create)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    try:
        scene = b.LoadScene(SCENE, W, H)
        r.upload(scene)
        r.photons_set(GOLD["photons"])
        r.photons_build(1)
        return r.render(mode=mode, use_photon_map=True, photon_k=K, photon_radius=RADIUS, **kw)
    finally:
        r.close()


@pytest.mark.gpu
def test_gpu_gather_schedules_agree_in_the_deterministic_mode(built):
    """Exact build: the in-kernel and the split schedule run the same routine on the same queries -- identical frames, both
    equal to the reference's; the CTA-wide routine adds its sums in another order -- within the 1/255 bar of the fixture."""
    import bhraytracer_b200 as b
    inker = _render_with_env({"BHRT_GATHER": "inkernel"}, b.MODE_DET)
    split = _render_with_env({"BHRT_GATHER": "split"}, b.MODE_DET)
    cta = _render_with_env({"BHRT_GATHER": "split", "BHRT_GATHER_DEFER": "64"}, b.MODE_DET)
    for o in (inker, split, cta):
        check_gather_image(o["rgb8"], o["hit_id"])
    assert (inker["rgb8"] == split["rgb8"]).all() and (inker["linear"] == split["linear"]).all()
    assert (inker["rays_closest"], inker["rays_shadow"]) == (split["rays_closest"], split["rays_shadow"]) == (cta["rays_closest"], cta["rays_shadow"])
    assert np.abs(cta["linear"] - split["linear"]).max() < 1e-4


@pytest.mark.gpu
def test_gpu_gather_schedules_agree_in_the_stochastic_mode(built):
    """Fast-math build, 8 spp: same samples (the RNG is keyed by pixel and sample), same photons; the schedules may differ
    in the last bits of a query's float sums (the inlined and the out-of-line copy of the routine are compiled apart),
    which moves single samples, not the image."""
    import bhraytracer_b200 as b
    frames = [_render_with_env(env, b.MODE_PATH, spp=8, seed=3) for env in
              ({"BHRT_GATHER": "inkernel"}, {"BHRT_GATHER": "split"}, {"BHRT_GATHER": "split", "BHRT_GATHER_DEFER": "64"})]
    ref = frames[0]
    for o in frames[1:]:
        assert abs(o["rays_closest"] + o["rays_shadow"] - ref["rays_closest"] - ref["rays_shadow"]) <= 1e-4 * (ref["rays_closest"] + ref["rays_shadow"])
        assert np.abs(o["linear"] - ref["linear"]).mean() < 2e-4
        assert (np.abs(o["rgb8"].astype(int) - ref["rgb8"].astype(int)) > 1).mean() < 1e-3
