"""The asynchronous seam the reference's viewer expects (viewport.cpp:37-38 "renderer must run in a separate thread",
:395-415 progress polling through RenderImage::IsRenderDone / GetNumRenderedPixels, :449-453 BeginRender / StopRender):
bhrt_render_async / bhrt_progress / bhrt_cancel / bhrt_wait."""
import time

import numpy as np
import pytest

import bhraytracer_b200 as b

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer(built):
    r = b.Renderer(0)
    yield r
    r.close()


def test_async_frame_equals_the_blocking_one_and_reports_progress(renderer):
    scene = b.LoadScene("scenes/cfg4_gi.xml", 960, 540)
    renderer.upload(scene)
    ref = renderer.render(mode=b.MODE_PATH, spp=32, seed=9)
    renderer.render_async(mode=b.MODE_PATH, spp=32, seed=9)
    seen = []
    while True:
        done, total, finished = renderer.progress()
        assert total == 960 * 540
        seen.append(done)
        if finished:
            break
        time.sleep(0.002)
    out = renderer.wait()
    assert seen == sorted(seen), "progress went backwards"
    assert seen[-1] == 960 * 540
    assert any(0 < s < 960 * 540 for s in seen), "never saw the frame half way (polls: %d)" % len(seen)
    # same seed, same samples, same reduction order: the same bits
    assert (out["linear"].view(np.uint32) == ref["linear"].view(np.uint32)).all()
    assert (out["rays_closest"], out["rays_shadow"]) == (ref["rays_closest"], ref["rays_shadow"])
    with pytest.raises(b.BhrtError):
        renderer.wait()                 # nothing in flight any more


def test_cancel_stops_a_long_frame_early(renderer):
    scene = b.LoadScene("scenes/cfg4_gi.xml", 1920, 1080)
    renderer.upload(scene)
    t0 = time.perf_counter()
    full = renderer.render(mode=b.MODE_PATH, spp=128, seed=2, want=())
    t_full = time.perf_counter() - t0
    t0 = time.perf_counter()
    renderer.render_async(mode=b.MODE_PATH, spp=128, seed=2, want=("rgb8",))
    while True:
        done, total, finished = renderer.progress()
        if finished or done > total // 10:
            break
        time.sleep(0.001)
    assert not finished, "the frame finished before it could be cancelled"
    renderer.cancel()
    with pytest.raises(b.BhrtError) as e:
        renderer.wait()
    t_cancelled = time.perf_counter() - t0
    assert "-5" in str(e.value) and "cancel" in str(e.value)
    assert t_cancelled < 0.6 * t_full, "cancelled frame took %.3f s of %.3f s" % (t_cancelled, t_full)
    # the ctx is usable again, and a second async frame may start
    again = renderer.render(mode=b.MODE_PATH, spp=2, seed=2, want=())
    assert again["rays_closest"] > 0
    assert full["rays_closest"] > again["rays_closest"]
    with pytest.raises(b.BhrtError):
        renderer.render_async(mode=b.MODE_PATH, spp=1)
        renderer.render_async(mode=b.MODE_PATH, spp=1)        # second one while the first is in flight: refused
    renderer.wait()


def test_cancel_stops_a_photon_mapped_frame_between_its_gather_rounds(renderer):
    """A photon-mapped frame is hundreds of launches (render kernel / caustic gather in turns): progress moves
    across them, a cancel ends the frame after the samples in flight have had their remaining gathers, and the ctx renders
    the same frame completely afterwards."""
    scene = b.LoadScene("scenes/cfg5_caustics.xml", 960, 540)
    renderer.upload(scene)
    stored, _ = renderer.photons_emit(1_000_000, seed=1)
    renderer.photons_build(stored)
    t0 = time.perf_counter()
    full = renderer.render(mode=b.MODE_PATH, spp=64, seed=4, use_photon_map=True, want=("rgb8",))
    t_full = time.perf_counter() - t0
    t0 = time.perf_counter()
    renderer.render_async(mode=b.MODE_PATH, spp=64, seed=4, use_photon_map=True, want=("rgb8",))
    seen = []
    while True:
        done, total, finished = renderer.progress()
        seen.append(done)
        if finished or done > total // 8:
            break
        time.sleep(0.001)
    assert not finished, "the frame finished before it could be cancelled"
    assert seen == sorted(seen)
    renderer.cancel()
    with pytest.raises(b.BhrtError) as e:
        renderer.wait()
    t_cancelled = time.perf_counter() - t0
    assert "cancel" in str(e.value)
    assert t_cancelled < 0.6 * t_full, "cancelled frame took %.3f s of %.3f s" % (t_cancelled, t_full)
    again = renderer.render(mode=b.MODE_PATH, spp=64, seed=4, use_photon_map=True, want=("rgb8",))
    assert (again["rays_closest"], again["rays_shadow"]) == (full["rays_closest"], full["rays_shadow"])
    assert (again["rgb8"] == full["rgb8"]).all()
