// photon_build.hpp -- grid parameters of the photon gather structure, and its HOST builder.
//
// The reference balances its photons into a left-balanced kd-tree
// (PhotonMap::PrepareForIrradianceEstimation, reference cyPhotonMap.h:236-328) after scaling their
// powers by 1/stored (Main.cpp:380).  The GPU gather (core.cuh estimate_irradiance) uses a sorted
// multi-resolution grid instead; this header holds what both builders share -- the choice of the
// grid (origin, coarse cell, dimensions) -- plus a plain std::sort builder used by the CPU test
// harness (tests/emu) and as the cross-check of the device builder in photon.cuh.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>
#include "core.cuh"

namespace bhrt {

struct PhotonGrid {
    float origin[3];
    float cell;
    int dim[3];
};

// Grid over the bounding box [lo,hi] of the photons.  The coarse cell is slightly larger than HALF the
// gather radius (so the radius itself is a safe bound for the 5x5x5 block of search level 0) and grows if the box would
// need more coarse cells than the key has room for (key layout: 32 - BHRT_PM_FINE_BITS bits coarse | BHRT_PM_FINE_BITS bits Morton).
inline PhotonGrid photon_grid(const float lo[3], const float hi[3], float radius)
{
    PhotonGrid g;
    float cell = radius * 1.001f * 0.5f;
    for (;;) {
        double n = 1;
        for (int a = 0; a < 3; a++) {
            g.origin[a] = lo[a] - 0.01f * cell;
            g.dim[a] = (int)floorf((hi[a] - g.origin[a]) / cell) + 1;
            if (g.dim[a] < 1) g.dim[a] = 1;
            n *= g.dim[a];
        }
        if (n < (double)(1u << (32 - BHRT_PM_FINE_BITS))) break;
        cell *= 1.26f;
    }
    g.cell = cell;
    return g;
}

struct HostPhotonMap {
    PhotonGrid grid;
    std::vector<uint32_t> keys;
    std::vector<GatherRec> recs;
    std::vector<uint32_t> coarse_start;
    PhotonGather view(int k, float radius) const
    {
        PhotonGather g;
        memset(&g, 0, sizeof(g));
        g.use = 1; g.k = k; g.radius = radius; g.n_photons = (int)keys.size();
        g.recs = recs.data(); g.keys = keys.data(); g.coarse_start = coarse_start.data();
        for (int a = 0; a < 3; a++) { g.origin[a] = grid.origin[a]; g.dim[a] = grid.dim[a]; }
        g.cell = grid.cell;
        return g;
    }
};

// records: n photons with UNSCALED power; total_stored: the global stored count (Main.cpp:380)
inline void build_photon_map_host(const PhotonRecord *records, int n, int total_stored, float radius, HostPhotonMap &out)
{
    float lo[3] = { 1e30f, 1e30f, 1e30f }, hi[3] = { -1e30f, -1e30f, -1e30f };
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], records[i].pos[a]); hi[a] = std::max(hi[a], records[i].pos[a]); }
    if (n == 0) for (int a = 0; a < 3; a++) lo[a] = hi[a] = 0;
    out.grid = photon_grid(lo, hi, radius);
    const float scale = 1.f / (float)(total_stored > 0 ? total_stored : 1);
    std::vector<uint32_t> key(n), order(n);
    for (int i = 0; i < n; i++) {
        uint32_t k = 0;
        photon_key(out.grid.origin, out.grid.cell, out.grid.dim, ld3(records[i].pos), k);   // inside by construction
        key[i] = k;
    }
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    out.keys.resize(n); out.recs.resize(n);
    for (int i = 0; i < n; i++) { out.keys[i] = key[order[i]]; out.recs[i] = photon_gather_rec(records[order[i]], scale); }
    const size_t nc = (size_t)out.grid.dim[0] * out.grid.dim[1] * out.grid.dim[2];
    out.coarse_start.assign(nc + 1, 0);
    for (int i = 0; i < n; i++) out.coarse_start[(out.keys[i] >> BHRT_PM_FINE_BITS) + 1]++;
    for (size_t c = 0; c < nc; c++) out.coarse_start[c + 1] += out.coarse_start[c];
}

}  // namespace bhrt
