// wavefront.cuh -- the headless wavefront renderer that replaces the pixel loop of BeginRender()
// (reference Main.cpp:204-234: `#pragma omp parallel for` over image columns, PathTracing(i,j)
// per pixel).
//
// Why a wavefront: the estimator (MtlBlinn::Shade, reference MtlBlinn.cpp:89-138) is a clamped
// recursion tree with nine distinct stages.  Run as one big per-pixel kernel it executed with
// ~3 of 32 lanes active per issued instruction and spent most issue slots waiting on instruction
// fetch (profiles/r1a_*, r1b_*: 64 KB of hot code, stall_no_instruction dominant).  Here every
// camera sample is a LANE whose state (core.cuh `Lane`: query, hit, frame stack, RNG) lives in
// HBM, and five small kernels advance all live lanes:
//
//   wf_trace_graph  the scene-graph node loop for the lanes that wait for a ray (warp-uniform control
//                   flow, analytic objects, mesh root boxes); lanes whose ray entered a mesh are
//                   compacted into a queue (warp-aggregated atomic), the rest are finished;
//   wf_trace_mesh   BVH traversal for the queued lanes only, persistent lanes with per-lane refill:
//                   every lane of every warp steps through BVH nodes; closest-hit and shadow
//                   queries share the stepper (core.cuh mesh_trav_step);
//   wf_trace_finish winner-only work of the closest hits (uv with atan2/asin, ray differentials,
//                   normal interpolation, FromNodeCoords);
//   wf_step         ONE step of the estimator coroutine (sample_step) per live lane, lanes sorted BY
//                   the step they run, so a warp executes one case of the state machine and a CTA's
//                   instruction footprint is one or two cases; a lane whose sample finished writes
//                   the sample colour and immediately starts the next camera sample (regeneration,
//                   one warp-aggregated atomic per refill), so the pool stays full until the frame
//                   runs out of samples;
//   wf_scatter      counting sort of the live lanes by next step into the next live list
//                   (ray-waiting steps first; dead lanes dropped); warp-aggregated with
//                   __match_any_sync.
//
// Sample colours go to a per-chunk buffer and are added to the accumulation buffer IN SAMPLE ORDER
// by wf_reduce, so a pixel's sum is bit-identical whatever the pool size, launch geometry or GPU
// count (same guarantee the first, per-pixel kernel gave).
#pragma once
#include <cuda_runtime.h>
#include "core.cuh"

namespace bhrt {

#define WF_BLOCK 256
#define WF_NKEYS 16
// Sort keys = the step a lane runs next, ordered so that the lanes that need a ray first come
// first: [0, n_trace) of the sorted live list is exactly the input of the trace kernels.
enum { WF_KEY_PRIMARY = 0, WF_KEY_REFR_IN = 1, WF_KEY_REFR_OUT = 2, WF_KEY_GI_TRACE = 3, WF_KEY_SHADOW = 4,
       WF_KEY_REFR = 5, WF_KEY_REFR_RET = 6, WF_KEY_GI = 7, WF_KEY_GI_RET = 8, WF_KEY_DIRECT = 9, WF_KEY_DEAD = 10,
       WF_N_TRACE_KEYS = 5 };
__device__ __forceinline__ unsigned wf_key_of_stage(int stage)
{
    // ST_REFR 0, ST_REFR_IN_WAIT 1, ST_REFR_OUT_WAIT 2, ST_REFR_WAIT 3, ST_GI 4, ST_GI_TRACE_WAIT 5, ST_GI_WAIT 6, ST_DIRECT 7, ST_SHADOW_WAIT 8
    const unsigned long long table = 0x498376215ull;       // nibble i = key of stage i
    return (unsigned)((table >> (4 * stage)) & 15ull);
}

struct WaveCounters {
    unsigned long long rays_closest, rays_shadow;
    unsigned long long next_work;           // next camera-sample work item of the chunk
    unsigned long long pad;
    unsigned n_active[2];                   // live lanes of each parity's list
    unsigned n_trace[2];                    // ... of which the first n_trace wait for a ray
    unsigned n_mesh[2];                     // lanes queued for BVH traversal this iteration
    unsigned mesh_cursor[2];                // ... and how many of them the mesh kernel has taken
    unsigned key_count[2][WF_NKEYS];
    unsigned key_cursor[2][WF_NKEYS];
};

struct WaveArgs {
    ShadeCfg cfg;
    unsigned long long seed;
    int sample_first;           // global index of the chunk's first sample
    int chunk_spp;              // samples per pixel in this chunk
    int row_begin, row_stride, n_rows, tiles_x;
    unsigned n_pix_work;        // tiles * 64 (work items outside the image are skipped)
    unsigned long long n_work;  // n_pix_work * chunk_spp
    PhotonGather gather;
    Lane *lanes;                // [pool] lane headers
    Frame *frames;              // [pool][BHRT_MAX_FRAMES] recursion levels
    unsigned *list[2];          // live lists
    unsigned *mesh_queue;       // lanes whose ray entered a mesh's root box
    unsigned char *keys;        // key of live-list slot i after wf_step
    float4 *samples;            // [chunk_spp][n_pix_work]
    int *hit_id; float *zbuf;
    WaveCounters *cnt;
    unsigned pool;              // number of lanes
    int smem_nodes, smem_materials, smem_lights, smem_offsets;   // byte offsets of the staged tables
};

// Contiguous, warp-uniform partition of [0,n) over the grid: block b owns one range (so it stays
// inside one or two sort keys -- one case of the estimator -- and its instruction footprint
// stays small), every thread of a warp runs the same number of trips.
struct WfRange { unsigned begin, trips; };
__device__ __forceinline__ WfRange wf_block_range(unsigned n)
{
    unsigned per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + blockDim.x - 1) / blockDim.x * blockDim.x;
    WfRange r; r.begin = blockIdx.x * per; r.trips = per / blockDim.x;
    return r;
}

// Stage the small, uniformly-read scene tables in shared memory (node table, materials, lights,
// texture-filter taps) and redirect the view to them.
__device__ __forceinline__ void stage_tables(SceneView &s, const WaveArgs &a, unsigned char *smem)
{
    const uint32_t *src; uint32_t *dst; int n;
    src = (const uint32_t *)s.nodes; dst = (uint32_t *)(smem + a.smem_nodes); n = s.n_nodes * (int)(sizeof(bhrt_node) / 4);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    src = (const uint32_t *)s.materials; dst = (uint32_t *)(smem + a.smem_materials); n = s.n_materials * (int)(sizeof(bhrt_material) / 4);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    src = (const uint32_t *)s.lights; dst = (uint32_t *)(smem + a.smem_lights); n = s.n_lights * (int)(sizeof(bhrt_light) / 4);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    src = (const uint32_t *)s.tex_offsets; dst = (uint32_t *)(smem + a.smem_offsets); n = 2 * (BHRT_TEX_SAMPLES - 1);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    s.nodes = (const bhrt_node *)(smem + a.smem_nodes);
    s.materials = (const bhrt_material *)(smem + a.smem_materials);
    s.lights = (const bhrt_light *)(smem + a.smem_lights);
    s.tex_offsets = (const float *)(smem + a.smem_offsets);
}

// 128-bit copies between a lane's records in HBM and the thread's own copy.  All loads of a record
// are issued back to back (memory-level parallelism) instead of one dependent scalar load per
// field; the thread-local copy is interleaved across the warp by the hardware, i.e. coalesced.
template <typename T>
__device__ __forceinline__ void wf_load(T &dst, const T *src)
{
    static_assert(sizeof(T) % 16 == 0, "record size must be a multiple of 16 bytes");
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(&dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d4[i] = s4[i];
}
template <typename T>
__device__ __forceinline__ void wf_store(T *dst, const T &src)
{
    const uint4 *s4 = reinterpret_cast<const uint4 *>(&src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d4[i] = s4[i];
}

// ---- wf_trace_graph / wf_trace_mesh / wf_trace_finish ----------------------------------------------
// Phase 1 of scene_trace for the lanes that wait for a ray: the warp-uniform node loop.  Shadow rays
// blocked by an analytic object are answered here; lanes whose ray entered a mesh's root box are
// appended to the mesh queue (warp-aggregated atomic).
__global__ void __launch_bounds__(WF_BLOCK)
wf_trace_graph(SceneView s, WaveArgs a, int parity)
{
    extern __shared__ __align__(16) unsigned char smem[];
    stage_tables(s, a, smem);
    const unsigned n = a.cnt->n_trace[parity];
    const unsigned *list = a.list[parity];
    const unsigned lane = threadIdx.x & 31u;
    const WfRange R = wf_block_range(n);
    for (unsigned t = 0; t < R.trips; t++) {
        const unsigned i = R.begin + t * blockDim.x + threadIdx.x;
        const bool valid = i < n;
        bool to_mesh = false;
        unsigned id = 0;
        if (valid) {
            id = list[i];
            Lane &L = a.lanes[id];
            Query q;
            wf_load(q, &L.q);
            Hit h;
            uint32_t mask;
            const bool blocked = scene_trace_graph(s, q, h, mask);
            L.hit_ok = blocked;
            L.hit.z = h.z; L.hit.node = h.node; L.hit.front = h.front;
            L.mesh_mask = 0;
            if (!blocked && (mask || s.n_mesh_nodes > BHRT_MESH_MASK_BITS)) { L.mesh_mask = mask; to_mesh = true; }
        }
        const unsigned m = __ballot_sync(0xffffffffu, to_mesh);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(&a.cnt->n_mesh[parity], (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (to_mesh) a.mesh_queue[base + __popc(m & ((1u << lane) - 1u))] = id;
        }
    }
}

// Phase 2: BVH traversal for the queued lanes only.  Persistent lanes with PER-LANE refill: a lane
// whose traversal finished takes the next queued ray at once (one warp-aggregated atomic per
// refill) instead of idling until the longest traversal of its warp ends, so the warp keeps ~all
// lanes stepping through BVH nodes (traversal lengths are heavily skewed: most rays leave the
// mesh after a few nodes, a few visit hundreds).
__global__ void __launch_bounds__(WF_BLOCK)
wf_trace_mesh(SceneView s, WaveArgs a, int parity)
{
    extern __shared__ __align__(16) unsigned char smem[];
    stage_tables(s, a, smem);
    const unsigned n = a.cnt->n_mesh[parity];
    const unsigned lane = threadIdx.x & 31u;
    MeshJob J;
    bool have = false, exhausted = false;
    unsigned id = 0;
    for (;;) {
        const bool need = !have && !exhausted;
        const unsigned nm = __ballot_sync(0xffffffffu, need);
        if (nm) {
            const int leader = __ffs(nm) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(&a.cnt->mesh_cursor[parity], (unsigned)__popc(nm));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (need) {
                const unsigned i = base + __popc(nm & ((1u << lane) - 1u));
                if (i >= n) exhausted = true;
                else {
                    id = a.mesh_queue[i];
                    const Lane &L = a.lanes[id];
                    Query q;
                    wf_load(q, &L.q);
                    mesh_job_begin(J, s, q, L.hit.z, L.hit.node, L.hit.front, L.mesh_mask);
                    have = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, have)) {
            if (!__any_sync(0xffffffffu, !exhausted)) break;
            continue;
        }
        if (have && mesh_job_step(J, s)) {
            Lane &L = a.lanes[id];
            if (J.blocked) L.hit_ok = true;
            else if (J.q.kind == Q_CLOSEST) {
                L.hit.z = J.h.z; L.hit.node = J.h.node; L.hit.front = J.h.front;
                L.hit.uvw = J.h.uvw; L.hit.duvw0.x = J.h.duvw0.x;
            }
            have = false;
        }
    }
}

// Phase 3: complete the closest hits (winner-only transcendental work, FromNodeCoords)
__global__ void __launch_bounds__(WF_BLOCK)
wf_trace_finish(SceneView s, WaveArgs a, int parity)
{
    extern __shared__ __align__(16) unsigned char smem[];
    stage_tables(s, a, smem);
    const unsigned n = a.cnt->n_trace[parity];
    const unsigned *list = a.list[parity];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Lane &L = a.lanes[list[i]];
        Query q;
        wf_load(q, &L.q);
        if (q.kind != Q_CLOSEST) continue;
        Hit h;
        wf_load(h, &L.hit);
        L.hit_ok = scene_trace_finish(s, q, h);
        wf_store(&L.hit, h);
    }
}

// ---- wf_step --------------------------------------------------------------------------------------
// Work item g of the chunk -> (pixel, sample).  Sample-major: consecutive items are the pixels of
// one 8x8 tile, so the lanes of a warp start on neighbouring primary rays.
__device__ __forceinline__ bool work_to_pixel(const SceneView &s, const WaveArgs &a, unsigned long long g, int &pi, int &pj, int &sidx)
{
    const unsigned w = (unsigned)(g % a.n_pix_work);
    sidx = a.sample_first + (int)(g / a.n_pix_work);
    const unsigned tile = w >> 6, in = w & 63u;
    const int tx = (int)(tile % (unsigned)a.tiles_x), ty = (int)(tile / (unsigned)a.tiles_x);
    const int i = tx * 8 + (int)(in & 7u), trow = ty * 8 + (int)(in >> 3);
    if (i >= s.width || trow >= a.n_rows) return false;
    pi = i; pj = a.row_begin + trow * a.row_stride;
    return true;
}

// One estimator step per live lane (sample_step): lanes arrive sorted by the step they run, so a
// warp executes one case.  The lane header and the top frame are copied in with 128-bit loads,
// stepped in the thread's own copy and written back (plus the child frame when the step
// recursed).  A finished sample is written out and the lane immediately starts the next camera
// sample of the chunk (warp-aggregated fetch).  `first` = fill the pool.
__global__ void __launch_bounds__(WF_BLOCK)
wf_step(SceneView s, WaveArgs a, int parity, int first)
{
    extern __shared__ __align__(16) unsigned char smem[];
    stage_tables(s, a, smem);
    const unsigned n = first ? a.pool : a.cnt->n_active[parity];
    const unsigned *list = a.list[parity];
    const bool det = a.cfg.mode == BHRT_MODE_DET;
    const unsigned lane = threadIdx.x & 31u;
    // the counters of this parity were consumed by the previous iteration's wf_scatter / trace
    // kernels and are filled again by the next iteration: clear them here
    if (!first && blockIdx.x == 0 && threadIdx.x < WF_NKEYS) {
        a.cnt->key_count[parity][threadIdx.x] = 0; a.cnt->key_cursor[parity][threadIdx.x] = 0;
        if (threadIdx.x == 0) { a.cnt->n_mesh[parity] = 0; a.cnt->mesh_cursor[parity] = 0; }
    }
    RayCounters rc; rc.closest = 0; rc.shadow = 0;
    const WfRange R = wf_block_range(n);
    for (unsigned t = 0; t < R.trips; t++) {
        const unsigned i = R.begin + t * blockDim.x + threadIdx.x;
        const bool valid = i < n;
        const unsigned id = valid ? (first ? i : list[i]) : 0u;
        Lane L;
        Frame top, child;
        Frame *frames = a.frames + (size_t)id * BHRT_MAX_FRAMES;
        bool want_new = valid && first;
        int sp_in = 0;
        if (valid && !first) {
            wf_load(L, &a.lanes[id]);
            sp_in = L.sp;
            if (sp_in > 0) wf_load(top, &frames[sp_in - 1]);
            V3 c;
            if (sample_step(L, top, child, s, a.cfg, a.gather, L.bgu, L.bgv, rc, c) == STEP_DONE) {
                a.samples[L.work] = make_float4(c.x, c.y, c.z, 1.f);
                if (L.sidx == 0) { const size_t p = (size_t)L.pj * s.width + L.pi; a.hit_id[p] = L.hit_id; a.zbuf[p] = L.hit_z; }
                want_new = true;
            } else {
                if (sp_in > 0 && L.sp >= sp_in) wf_store(&frames[sp_in - 1], top);    // (a popped frame is dead)
                if (L.sp > sp_in) wf_store(&frames[sp_in], child);
            }
        }
        // ---- regeneration: warp-aggregated fetch of new camera samples
        bool dead = false;
        for (;;) {
            const unsigned m = __ballot_sync(0xffffffffu, want_new);
            if (!m) break;
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(&a.cnt->next_work, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want_new) {
                const unsigned long long g = base + __popc(m & ((1u << lane) - 1u));
                if (g >= a.n_work) { dead = true; want_new = false; }
                else {
                    int pi, pj, sidx;
                    if (work_to_pixel(s, a, g, pi, pj, sidx)) {
                        L.work = g; L.pi = pi; L.pj = pj; L.sidx = sidx;
                        L.bgu = (float)pi / s.width; L.bgv = (float)pj / s.height;
                        L.rng.seed(a.seed, (uint64_t)pj * s.width + pi, (uint64_t)sidx);
                        Ray pr = primary_ray(s, pi, pj, !det, L.rng);
                        sample_begin(L, pr, rc);
                        want_new = false;
                    }
                }
            }
        }
        // ---- key of this slot for the counting sort
        unsigned key = WF_KEY_DEAD;
        if (valid && !dead) {
            key = (L.sp == 0) ? (unsigned)WF_KEY_PRIMARY : wf_key_of_stage(L.stage[L.sp - 1]);
            wf_store(&a.lanes[id], L);
        }
        if (valid) a.keys[i] = (unsigned char)key;
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned peers = __match_any_sync(vm, key);
            if ((int)lane == __ffs(peers) - 1) atomicAdd(&a.cnt->key_count[parity ^ 1][key], (unsigned)__popc(peers));
        }
    }
    unsigned long long c0 = rc.closest, c1 = rc.shadow;
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_down_sync(0xffffffffu, c0, o);
        c1 += __shfl_down_sync(0xffffffffu, c1, o);
    }
    if (lane == 0 && (c0 | c1)) { atomicAdd(&a.cnt->rays_closest, c0); atomicAdd(&a.cnt->rays_shadow, c1); }
}

// ---- wf_scatter -----------------------------------------------------------------------------------
// Counting sort of the slots of list[parity] (keys[] from wf_advance) into list[parity^1]; dead
// lanes are dropped.  Also publishes n_active[parity^1].
__global__ void __launch_bounds__(WF_BLOCK)
wf_scatter(WaveArgs a, int parity, int first)
{
    const unsigned n = first ? a.pool : a.cnt->n_active[parity];
    const unsigned *list = a.list[parity];
    unsigned *out = a.list[parity ^ 1];
    const unsigned *count = a.cnt->key_count[parity ^ 1];
    unsigned *cursor = a.cnt->key_cursor[parity ^ 1];
    __shared__ unsigned offs[WF_NKEYS];
    if (threadIdx.x == 0) {
        unsigned acc = 0, ntr = 0;
        for (int k = 0; k < WF_NKEYS; k++) { offs[k] = acc; if (k != WF_KEY_DEAD) acc += count[k]; if (k == WF_N_TRACE_KEYS - 1) ntr = acc; }
        if (blockIdx.x == 0) { a.cnt->n_active[parity ^ 1] = acc; a.cnt->n_trace[parity ^ 1] = ntr; }
    }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_round = (n + 31u) & ~31u;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n;
        const unsigned key = valid ? a.keys[i] : (unsigned)WF_KEY_DEAD;
        const bool live = valid && key != WF_KEY_DEAD;
        const unsigned lm = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned peers = __match_any_sync(lm, key);
            const int leader = __ffs(peers) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(&cursor[key], (unsigned)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            out[offs[key] + base + __popc(peers & ((1u << lane) - 1u))] = first ? i : list[i];
        }
    }
}

// ---- wf_reduce ------------------------------------------------------------------------------------
// accum[p] (+)= sum over the chunk's samples IN SAMPLE ORDER (Main.cpp:150-170 adds them in order)
__global__ void wf_reduce(SceneView s, WaveArgs a, float4 *__restrict__ accum, int overwrite)
{
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.n_pix_work) return;
    int pi, pj, sidx;
    if (!work_to_pixel(s, a, w, pi, pj, sidx)) return;
    const size_t p = (size_t)pj * s.width + pi;
    float4 v = overwrite ? make_float4(0.f, 0.f, 0.f, 0.f) : accum[p];
    for (int k = 0; k < a.chunk_spp; k++) {
        const float4 c = a.samples[(size_t)k * a.n_pix_work + w];
        v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
    }
    accum[p] = v;
}

}  // namespace bhrt
