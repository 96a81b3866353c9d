// photon.cuh -- caustic photon map on the GPU: sharded emission and the gather-structure build.
//
// Replaces BuildCausticPhotonMap (reference Main.cpp:342-386; a SERIAL while-loop there, ~49 s per
// million stored photons) and PhotonMap::PrepareForIrradianceEstimation (cyPhotonMap.h:236-328).
//
//   pm_emit      persistent lanes; each runs photon paths (core.cuh photon_begin / photon_advance,
//                one ray query per step, the same scene_trace as the renderer) until the map holds
//                `target` photons; path indices come from a warp-aggregated global counter and key
//                the RNG, so a photon does not depend on which lane traced it.  Records are the
//                reference's 24-byte Photon (cyPhotonMap.h:72-80) with UNSCALED power.
//   build        bounding box (atomics on order-preserving ints) -> sort keys (coarse cell | Morton
//                sub-cell) -> radix sort of (key, index) -> 32-byte gather records in key order with
//                power scaled by 1/total_stored (Main.cpp:380) -> coarse-cell start table.  The
//                radix sort is cub::DeviceRadixSort (CUDA toolkit); everything else is hand-written.
//
// The gather itself (estimate_irradiance) is device code in core.cuh, called from the estimator.
#pragma once
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <string>
#include <vector>
#include "core.cuh"
#include "photon_build.hpp"

namespace bhrt {

struct PmEmitCounters {
    unsigned long long next_path;       // photon paths started
    unsigned long long rays;            // closest-hit queries traced for photons
    unsigned int stored;                // photons appended (may run past `target`; clamped on the host)
    unsigned int pad;
};

__global__ void __launch_bounds__(128)
pm_emit(SceneView s, PhotonEmitCfg cfg, unsigned long long seed, unsigned target, unsigned long long max_paths, PhotonRecord *__restrict__ records, PmEmitCounters *cnt)
{
    const unsigned lane = threadIdx.x & 31u;
    PhotonLane P;
    bool active = false, exhausted = false;
    unsigned long long rays = 0;
    for (;;) {
        // ---- start new paths on idle lanes while the map is not full
        const bool need = !active && !exhausted;
        const unsigned nm = __ballot_sync(0xffffffffu, need);
        if (nm) {
            const int leader = __ffs(nm) - 1;
            unsigned long long base = 0;
            unsigned stored_now = 0;
            if ((int)lane == leader) {
                stored_now = *(volatile unsigned *)&cnt->stored;
                if (stored_now < target) base = atomicAdd(&cnt->next_path, (unsigned long long)__popc(nm));
            }
            stored_now = __shfl_sync(0xffffffffu, stored_now, leader);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (need) {
                // (the reference's while-loop, Main.cpp:363, never ends when the scene cannot store a caustic photon;
                // here the number of paths is bounded: the map then simply holds fewer photons than asked for)
                if (stored_now >= target || base >= max_paths) exhausted = true;
                else {
                    P.rng.seed(seed, base + __popc(nm & ((1u << lane) - 1u)), 0x50484f54u);
                    photon_begin(P, s, cfg, rays);
                    active = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) {
            if (!__any_sync(0xffffffffu, !exhausted)) break;
            continue;
        }
        if (active) {
            P.hit_ok = scene_trace(s, P.q, P.hit);
            PhotonRecord r; bool store;
            const bool ended = cfg.global_map ? photon_advance_global(P, s, r, store, rays) : photon_advance(P, s, r, store, rays);
            if (store) {
                const unsigned slot = atomicAdd(&cnt->stored, 1u);
                if (slot < target) records[slot] = r;       // PhotonMap::AddPhoton refuses beyond its size
            }
            if (ended) active = false;
        }
    }
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_down_sync(0xffffffffu, rays, o);
    if (lane == 0 && rays) atomicAdd(&cnt->rays, rays);
}

// ---- build kernels ----------------------------------------------------------------------------------
__device__ __forceinline__ int pm_f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float pm_ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void pm_bbox(const PhotonRecord *__restrict__ rec, int n, int *bb /* [6] ordered ints: min xyz, max xyz */)
{
    int lo[3] = { 0x7fffffff, 0x7fffffff, 0x7fffffff }, hi[3] = { (int)0x80000000, (int)0x80000000, (int)0x80000000 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int a = 0; a < 3; a++) { int v = pm_f2ord(rec[i].pos[a]); lo[a] = min(lo[a], v); hi[a] = max(hi[a], v); }
    for (int a = 0; a < 3; a++) {
        for (int o = 16; o > 0; o >>= 1) { lo[a] = min(lo[a], __shfl_down_sync(0xffffffffu, lo[a], o)); hi[a] = max(hi[a], __shfl_down_sync(0xffffffffu, hi[a], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(&bb[a], lo[a]); atomicMax(&bb[3 + a], hi[a]); }
    }
}

__global__ void pm_keys(const PhotonRecord *__restrict__ rec, int n, PhotonGrid g, uint32_t *__restrict__ keys, uint32_t *__restrict__ idx)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t k = 0;
        photon_key(g.origin, g.cell, g.dim, ld3(rec[i].pos), k);
        keys[i] = k; idx[i] = (uint32_t)i;
    }
}

__global__ void pm_gather_records(const PhotonRecord *__restrict__ rec, const uint32_t *__restrict__ keys, const uint32_t *__restrict__ idx, int n,
                                  float scale, GatherRec *__restrict__ out, uint32_t *__restrict__ coarse_count)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        out[i] = photon_gather_rec(rec[idx[i]], scale);
        atomicAdd(&coarse_count[keys[i] >> BHRT_PM_FINE_BITS], 1u);
    }
}

struct PhotonState {
    PhotonRecord *records = nullptr;    // device, unscaled power, emission/arrival order
    int count = 0, capacity = 0;
    bool built = false;
    float built_radius = 0;
    int built_total = 0;
    PhotonGrid grid;
    // gather structure
    GatherRec *recs = nullptr; uint32_t *keys = nullptr, *keys_tmp = nullptr, *idx = nullptr, *idx_tmp = nullptr, *coarse = nullptr;
    void *cub_tmp = nullptr; size_t cub_bytes = 0;
    int built_cap = 0; size_t coarse_cap = 0;
    PmEmitCounters *counters = nullptr;
    int *bbox = nullptr;

    void free_gather()
    {
        cudaFree(recs); cudaFree(keys); cudaFree(keys_tmp); cudaFree(idx); cudaFree(idx_tmp); cudaFree(coarse); cudaFree(cub_tmp);
        recs = nullptr; keys = keys_tmp = idx = idx_tmp = coarse = nullptr; cub_tmp = nullptr; cub_bytes = 0; built_cap = 0; coarse_cap = 0;
        built = false;
    }
    void release()
    {
        free_gather();
        cudaFree(records); cudaFree(counters); cudaFree(bbox);
        records = nullptr; counters = nullptr; bbox = nullptr; count = capacity = 0;
    }
    int reserve(int n, std::string &err)
    {
        if (n <= capacity) return 0;
        PhotonRecord *p = nullptr;
        if (cudaMalloc(&p, (size_t)n * sizeof(PhotonRecord)) != cudaSuccess) { err = "cudaMalloc(photon records) failed"; return BHRT_ERR_CUDA; }
        if (records && count > 0) cudaMemcpy(p, records, (size_t)count * sizeof(PhotonRecord), cudaMemcpyDeviceToDevice);
        cudaFree(records);
        records = p; capacity = n;
        return 0;
    }
    int ensure_small(std::string &err)
    {
        if (!counters && cudaMalloc(&counters, sizeof(PmEmitCounters)) != cudaSuccess) { err = "cudaMalloc failed"; return BHRT_ERR_CUDA; }
        if (!bbox && cudaMalloc(&bbox, 6 * sizeof(int)) != cudaSuccess) { err = "cudaMalloc failed"; return BHRT_ERR_CUDA; }
        return 0;
    }

    PhotonGather gather_view(int k, float radius) const
    {
        PhotonGather g;
        memset(&g, 0, sizeof(g));
        g.use = 1; g.k = k; g.radius = radius; g.n_photons = count;
        g.recs = recs; g.keys = keys; g.coarse_start = coarse;
        for (int a = 0; a < 3; a++) { g.origin[a] = grid.origin[a]; g.dim[a] = grid.dim[a]; }
        g.cell = grid.cell;
        return g;
    }

    // BuildCausticPhotonMap's emission loop.  The emitted records replace the current contents.
    int emit(const SceneView &view, const std::vector<bhrt_light> &sorted_lights, cudaStream_t stream, int sm_count, int target, unsigned long long seed,
             int *stored, unsigned long long *photon_rays, unsigned long long &launches, std::string &err, bool global_map = false)
    {
        if (target <= 0) { err = "bhrt_photons_emit: target_stored must be > 0"; return BHRT_ERR_ARG; }
        PhotonEmitCfg cfg = photon_emit_cfg(sorted_lights);
        cfg.global_map = global_map ? 1 : 0;
        if (cfg.n_lights == 0) { err = "bhrt_photons_emit: the scene has no point light (Main.cpp:354 returns false)"; return BHRT_ERR_STATE; }
        int r;
        count = 0; built = false;
        if ((r = reserve(target, err))) return r;
        if ((r = ensure_small(err))) return r;
        cudaMemsetAsync(counters, 0, sizeof(PmEmitCounters), stream);
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pm_emit, 128, 0);
        if (per_sm < 1) per_sm = 1;
        // path budget: 4096 paths per photon asked for, at least 2^26 (config 5 needs ~30 paths per stored photon)
        unsigned long long max_paths = 4096ull * (unsigned long long)target;
        if (max_paths < (1ull << 26)) max_paths = 1ull << 26;
        if (const char *e = getenv("BHRT_PHOTON_MAX_PATHS")) { const long long v = atoll(e); if (v > 0) max_paths = (unsigned long long)v; }
        // The reference's loop is serial: every path but the last one deposits ALL its photons.  Here the paths in
        // flight when the map fills up are cut short, which under-represents late bounces by (lanes in flight) /
        // (paths of the whole map): keep that share small -- at most one lane per 64 photons asked for.
        int blocks = sm_count * per_sm;
        const int cap = target / (64 * 128) > 1 ? target / (64 * 128) : 1;
        if (blocks > cap) blocks = cap;
        pm_emit<<<blocks, 128, 0, stream>>>(view, cfg, seed, (unsigned)target, max_paths, records, counters);
        launches++;
        PmEmitCounters h;
        if (cudaMemcpyAsync(&h, counters, sizeof(h), cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
            err = std::string("pm_emit: ") + cudaGetErrorString(cudaGetLastError()); return BHRT_ERR_CUDA;
        }
        count = (int)(h.stored < (unsigned)target ? h.stored : (unsigned)target);
        if (stored) *stored = count;
        if (photon_rays) *photon_rays = h.rays;
        return 0;
    }

    int set(const void *src, int n, bool from_device, cudaStream_t stream, std::string &err)
    {
        int r;
        count = 0; built = false;
        if ((r = reserve(n > 0 ? n : 1, err))) return r;
        if (n > 0 && cudaMemcpyAsync(records, src, (size_t)n * sizeof(PhotonRecord), from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream) != cudaSuccess) {
            err = "bhrt_photons_set: copy failed"; return BHRT_ERR_CUDA;
        }
        cudaStreamSynchronize(stream);
        count = n;
        return 0;
    }

    // ScalePhotonPowers(1/total) + the gather structure for a given gather radius
    int build(int total_stored, float radius, cudaStream_t stream, unsigned long long &launches, std::string &err)
    {
        int r;
        if ((r = ensure_small(err))) return r;
        const int n = count;
        float lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
        if (n > 0) {
            int init[6] = { 0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000 };
            cudaMemcpyAsync(bbox, init, sizeof(init), cudaMemcpyHostToDevice, stream);
            pm_bbox<<<256, 256, 0, stream>>>(records, n, bbox);
            launches++;
            int hb[6];
            cudaMemcpyAsync(hb, bbox, sizeof(hb), cudaMemcpyDeviceToHost, stream);
            if (cudaStreamSynchronize(stream) != cudaSuccess) { err = "pm_bbox failed"; return BHRT_ERR_CUDA; }
            for (int a = 0; a < 3; a++) {
                int l = hb[a] >= 0 ? hb[a] : hb[a] ^ 0x7fffffff;
                int h = hb[3 + a] >= 0 ? hb[3 + a] : hb[3 + a] ^ 0x7fffffff;
                memcpy(&lo[a], &l, 4);
                memcpy(&hi[a], &h, 4);
            }
        }
        grid = photon_grid(lo, hi, radius);
        const size_t nc = (size_t)grid.dim[0] * grid.dim[1] * grid.dim[2];
        const int cap = n > 0 ? n : 1;
        if (cap > built_cap || nc + 1 > coarse_cap) {
            free_gather();
            if (cudaMalloc(&recs, (size_t)cap * sizeof(GatherRec)) != cudaSuccess || cudaMalloc(&keys, (size_t)cap * 4) != cudaSuccess ||
                cudaMalloc(&keys_tmp, (size_t)cap * 4) != cudaSuccess || cudaMalloc(&idx, (size_t)cap * 4) != cudaSuccess ||
                cudaMalloc(&idx_tmp, (size_t)cap * 4) != cudaSuccess || cudaMalloc(&coarse, (nc + 2) * 4) != cudaSuccess) {
                err = "cudaMalloc(photon gather structure) failed"; return BHRT_ERR_CUDA;
            }
            built_cap = cap; coarse_cap = nc + 1;
        }
        cudaMemsetAsync(coarse, 0, (nc + 2) * 4, stream);
        if (n > 0) {
            pm_keys<<<512, 256, 0, stream>>>(records, n, grid, keys_tmp, idx_tmp);
            launches++;
            size_t need = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, need, keys_tmp, keys, idx_tmp, idx, n, 0, 32, stream);
            size_t need2 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, need2, coarse, coarse, (int)(nc + 1), stream);
            if (need2 > need) need = need2;
            if (need > cub_bytes) {
                cudaFree(cub_tmp); cub_tmp = nullptr; cub_bytes = 0;
                if (cudaMalloc(&cub_tmp, need) != cudaSuccess) { err = "cudaMalloc(sort scratch) failed"; return BHRT_ERR_CUDA; }
                cub_bytes = need;
            }
            size_t tmp = cub_bytes;
            cub::DeviceRadixSort::SortPairs(cub_tmp, tmp, keys_tmp, keys, idx_tmp, idx, n, 0, 32, stream);
            const float scale = 1.f / (float)(total_stored > 0 ? total_stored : 1);
            // counts go to coarse[c + 1]... the exclusive scan below turns counts at [c] into starts
            pm_gather_records<<<512, 256, 0, stream>>>(records, keys, idx, n, scale, recs, coarse);
            launches++;
            tmp = cub_bytes;
            cub::DeviceScan::ExclusiveSum(cub_tmp, tmp, coarse, coarse, (int)(nc + 1), stream);
        }
        if (cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) { err = "photon map build failed"; return BHRT_ERR_CUDA; }
        built = true; built_radius = radius; built_total = total_stored;
        return 0;
    }
};

}  // namespace bhrt
