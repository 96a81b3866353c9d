"""Read the .aov dumps written by oracle/_ref (driver.cpp), tests/emu and the CLI."""
import numpy as np


def read_aov(path):
    with open(path, "rb") as f:
        hdr = np.frombuffer(f.read(16), dtype="<i4")
        assert hdr[0] == 0x54524842, "bad magic in %s" % path
        w, h, spp = int(hdr[1]), int(hdr[2]), int(hdr[3])
        n = w * h
        linear = np.frombuffer(f.read(n * 12), dtype="<f4").reshape(h, w, 3)
        rgb8 = np.frombuffer(f.read(n * 3), dtype=np.uint8).reshape(h, w, 3)
        hit = np.frombuffer(f.read(n * 4), dtype="<i4").reshape(h, w)
        z = np.frombuffer(f.read(n * 4), dtype="<f4").reshape(h, w)
    return dict(width=w, height=h, spp=spp, linear=linear, rgb8=rgb8, hit_id=hit, z=z)


def compare(a, b):
    """Summary of the differences between two AOV sets (dicts from read_aov)."""
    out = {}
    out["hit_id_mismatch"] = int((a["hit_id"] != b["hit_id"]).sum())
    out["z_bits_mismatch"] = int((a["z"].view(np.uint32) != b["z"].view(np.uint32)).sum())
    d8 = np.abs(a["rgb8"].astype(np.int32) - b["rgb8"].astype(np.int32))
    out["rgb8_max_diff"] = int(d8.max())
    out["rgb8_pixels_gt1"] = int((d8.max(axis=2) > 1).sum())
    out["rgb8_pixels_ne"] = int((d8.max(axis=2) > 0).sum())
    la, lb = a["linear"].astype(np.float64), b["linear"].astype(np.float64)
    ok = np.isfinite(la) & np.isfinite(lb)
    out["linear_max_abs"] = float(np.abs(la - lb)[ok].max()) if ok.any() else 0.0
    out["linear_bits_mismatch"] = int((a["linear"].view(np.uint32) != b["linear"].view(np.uint32)).sum())
    out["rmse"] = float(np.sqrt(np.mean((np.clip(la, 0, 1) - np.clip(lb, 0, 1))[ok] ** 2)))
    return out


if __name__ == "__main__":
    import json
    import sys
    print(json.dumps(compare(read_aov(sys.argv[1]), read_aov(sys.argv[2]))))
