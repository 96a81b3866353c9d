"""The drop-in boundary: both shared libraries load WITHOUT a GPU and export every function that
include/bhrt.h and include/bhrt_host.h declare; the ctypes mirrors have the C struct sizes; and
the product fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

import bhraytracer_b200 as b
from conftest import ROOT


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bhrt_[a-z_0-9]+)\s*\(", src)))


def test_device_library_exports_every_declared_symbol(built):
    lib = b.device_lib()
    names = declared_functions("bhrt.h")
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), "libbhrt.so does not export " + n
    assert sorted(names) == sorted(b.DEVICE_SYMBOLS), "python binding list out of date"
    assert lib.bhrt_abi_version() == 1
    # the multi-GPU host lives in the same library (include/bhrt_dist.h)
    dnames = [n for n in declared_functions("bhrt_dist.h") if n.startswith("bhrt_dist_")]
    assert len(dnames) >= 8
    for n in dnames:
        assert hasattr(lib, n), "libbhrt.so does not export " + n
    assert sorted(dnames) == sorted(b.DIST_SYMBOLS)


def test_dist_partitions_tile_exactly(built):
    """bhrt_dist_shard_* (the C++ host's partitions) without a GPU: ranges tile, quotas sum, rows interleave."""
    for spp in (1, 5, 16, 255):
        for world in (1, 2, 3, 8):
            r = [b.dist_shards("samples", spp, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(e - s for s, e in r) - min(e - s for s, e in r) <= 1
            assert sum(b.dist_shards("photons", spp, k, world) for k in range(world)) == spp
    rows = sorted(j for k in range(3) for j in range(*b.dist_shards("rows", 10, k, 3)))
    assert rows == list(range(10))
    with pytest.raises(ValueError):
        b.dist_shards("rows", 10, 3, 3)
    # and they are the partitions the Python orchestration of the CPU test tier uses
    from bhraytracer_b200 import multigpu as mg
    for world in (1, 2, 3, 8):
        for k in range(world):
            assert b.dist_shards("samples", 37, k, world) == mg.shard_samples(37, k, world)
            assert b.dist_shards("rows", 37, k, world) == mg.shard_rows(37, k, world)
            assert b.dist_shards("photons", 3001, k, world) == mg.shard_photons(3001, k, world)


def test_host_library_exports_every_declared_symbol(built):
    lib = b.host_lib()
    names = declared_functions("bhrt_host.h")
    for n in names:
        assert hasattr(lib, n), "libbhrt_host.so does not export " + n
    assert sorted(names) == sorted(b.HOST_SYMBOLS)


def test_struct_sizes_match_the_header(built, tmp_path):
    """Compile a tiny C program against include/bhrt.h and compare sizeof() with the ctypes mirrors."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "bhrt.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(bhrt_scene_desc),sizeof(bhrt_params),sizeof(bhrt_outputs),sizeof(bhrt_node),sizeof(bhrt_material),'
                   'sizeof(bhrt_light),sizeof(bhrt_texmap),sizeof(bhrt_texture),sizeof(bhrt_mesh),sizeof(bhrt_bvh_node));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes[0] == C.sizeof(b.SceneDesc)
    assert sizes[1] == C.sizeof(b.Params)
    assert sizes[2] == C.sizeof(b.Outputs)
    assert sizes[3] == C.sizeof(b._Node)
    assert sizes[4:] == [72, 32, 88, 48, 88, 32]


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="this is the no-GPU behaviour")
def test_no_cpu_fallback(built):
    """Without a CUDA device the product must refuse to run rather than fall back."""
    with pytest.raises(b.BhrtError) as e:
        b.Renderer(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_error_codes_without_device(built):
    lib = b.device_lib()
    assert lib.bhrt_create(0, None) == -1                      # BHRT_ERR_ARG
    assert b"null" in lib.bhrt_last_error()
    assert lib.bhrt_destroy(None) == 0
    assert lib.bhrt_upload_scene(None, None) == -1
