"""Shared by tests/golden/make_golden.py and the photon-map tests: the reference's 24-byte photon
record (cyPhotonMap.h:72-80; PhotonMapViewer/PhotonMapViz.cpp:30-36 documents the same layout) and
the distribution summary the emission tests compare."""
import numpy as np

PHOTON_DT = np.dtype([("pos", "<f4", 3), ("power", "<f4"), ("rgb", "u1", 3), ("flag", "u1"), ("dx", "<i2"), ("dy", "<i2")])


def as_records(raw):
    """(N,24) uint8 -> structured view"""
    return np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1, 24).view(PHOTON_DT).reshape(-1)


def photon_stats(rec, scaled_by):
    """Positions on the floor (z = 0 plane of the Cornell scenes), per-axis histograms, mean power
    (`scaled_by` undoes ScalePhotonPowers, Main.cpp:380) and the share of downward photons."""
    pos = rec["pos"]
    floor = rec[np.abs(pos[:, 2]) < 1e-3]
    h2, _, _ = np.histogram2d(floor["pos"][:, 0], floor["pos"][:, 1], bins=24, range=[[-15, 15], [-30, 20]])
    ax = [np.histogram(pos[:, a], bins=16, range=(-16, 25))[0] / float(len(rec)) for a in range(3)]
    return dict(floor_hist=(h2 / max(1, len(rec))).astype(np.float32), axis_hist=np.array(ax, np.float32),
                floor_frac=len(floor) / float(len(rec)), mean_power=float((rec["power"].astype(np.float64) * scaled_by).mean()),
                dirz_neg_frac=float(((rec["flag"] & 8) != 0).mean()))
