// wavefront.cuh -- the headless renderer that replaces the pixel loop of BeginRender()
// (reference Main.cpp:204-234: `#pragma omp parallel for` over image columns, PathTracing(i,j)
// per pixel): ONE persistent kernel in which every CTA runs its own small wavefront.
//
// Why this shape (profiles/r1a..r1d document the three designs tried before it):
//  * the estimator (MtlBlinn::Shade, reference MtlBlinn.cpp:89-138) is a clamped recursion tree
//    with nine distinct stages.  One thread per pixel executed it with ~3 of 32 lanes active per
//    issued instruction and starved on instruction fetch (r1a, r1b);
//  * a device-wide wavefront (five kernels per iteration over a lane pool in HBM, r1c/r1d) fixed
//    the divergence but paid a device-wide barrier per iteration: BVH traversal lengths are
//    heavily skewed (88 % of the rays of config 2 never enter the mesh, 1 % need > 256 node steps,
//    a few need thousands), so every launch ended in a long, almost empty tail, and the lane
//    state made five round trips to HBM per ray.
//
// Here a CTA owns a private pool of lanes (camera samples in flight; state in global memory, but
// touched by this CTA only, so it lives in L1/L2) and loops over three phases separated by
// __syncthreads() -- the only synchronisation there is:
//
//   P1 lanes   one pass over the CTA's live lanes, SORTED BY WHAT THEY WILL EXECUTE NEXT (stage x
//              class of the pending ray's answer, wf_key): warps take 32-lane chunks of the sorted
//              list from a shared cursor, so a warp executes one case of the estimator and the
//              whole CTA walks the cases in step.  Per chunk: complete the pending closest hit
//              (winner-only transcendental work), run the estimator coroutine up to its next ray
//              (min-stage scheduling inside the warp), write finished samples and start the next
//              camera sample (warp-aggregated fetch from the global work counter), then answer
//              the new ray against the scene graph (warp-uniform node loop, analytic objects,
//              mesh root boxes).  Rays that entered a mesh are queued; the rest is answered.
//   P3 sort    counting sort of the live lanes by key (shared memory, no atomics).
//   P2 meshes  BVH traversal of the queued rays, run only when enough are queued to keep every
//              thread busy.  Each THREAD keeps its traversal (stack included) across iterations
//              and gets a few more trips after the queue has drained, so a ray that needs
//              thousands of steps occupies one thread for many iterations while its lane simply
//              stays out of the live list -- nobody waits for it.  A second sort merges the lanes
//              it completed.
//
// Sample colours go to a per-chunk buffer and are added to the accumulation buffer IN SAMPLE ORDER
// by wf_reduce, so on ONE GPU a pixel's sum is bit-identical whatever the pool size or launch geometry.
// (Across GPUs the deterministic mode stays bit-identical -- each pixel is written by one rank -- while the
// stochastic mode adds per-rank partial sums through NCCL: same samples, fp32 addition order differs.)
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "core.cuh"
// BHRT_WIDE_P2: the traversal phase runs the 8-lanes-per-ray traversal of the wide BVH (wide.cuh) instead of one
// ray per thread over the binary tree.  Set by render_fast.cu; the exact build keeps the binary tree, whose
// visiting order is part of the deterministic mode's bit-exactness.
#ifndef BHRT_WIDE_P2
#define BHRT_WIDE_P2 0
#endif
#if BHRT_WIDE_P2
#include "wide.cuh"
#endif

namespace BHRT_NS {

// WF_LOCAL_FRAMES: the estimator works on THREAD-LOCAL copies of the top recursion level and of the slot above it
// (128-bit copies in and out, like the lane record) instead of reaching into the frames in global memory field by
// field: a warp's lanes sit in 32 different 128-byte frames, so every scalar field access there costs 32 sectors.
// MEASURED (round 2, BASELINE config 2): 1338 Mrays/s with the copies against 1541 in place -- the sixteen extra
// 128-bit transfers per lane and pass and the spills they cause cost more than the scattered field accesses,
// most of which hit L1 right after the first touch.  Kept as a compile-time variant; off.
#ifndef WF_LOCAL_FRAMES
#define WF_LOCAL_FRAMES 0
#endif
#ifndef WF_BLOCK
#define WF_BLOCK 1024
#endif
#ifndef WF_MIN_BLOCKS
#define WF_MIN_BLOCKS 1      // resident CTAs per SM the register budget is sized for (1024 x 1 and 512 x 2 both mean 64 registers;
                            // 2 x 512 with pools of 1536 / 2048 / 3072 lanes measured 1242 / 1269 / 1293 Mrays/s against 1506)
#endif
// WF_PREFETCH_FRAME: the depth of every lane's frame stack is mirrored in shared memory, so a warp can ask for the top
// frame's line (prefetch.global.L1) in the same breath as the lane record instead of after it has arrived.
// MEASURED (round 2, config 2): 1495 Mrays/s with it, 1503 without -- no effect; off.
#ifndef WF_PREFETCH_FRAME
#define WF_PREFETCH_FRAME 0
#endif
#define WF_NKEYS 49
#define WF_MESH_BUDGET 4
// Sort keys.  A lane's key says what its next pass through the estimator will execute, as far as that
// is known when the lanes are sorted:
//   bits 5..3  the stage that consumes the pending ray's answer (6 slots)
//   bits 2..1  class of the answer: closest-hit stages -> 0 = miss, else the behaviour class of the
//              hit node's material; shadow stage -> class of the shading point's material
//              (1 diffuse only, 2 has a specular lobe, 3 refractive -- wf_material_class)
//   bit  0     the GI budget of the frame is used up (the next level shades direct light only)
// so a warp's 32 lanes not only run the same case of the state machine, they take the same branches
// inside it (refractive / glossy / diffuse material, miss -> environment, ...).  48 keys keep the
// sort's tables -- and with them the CTA's shared memory -- under 32 KB, which leaves L1 (where the
// thread-local lane copies live) the largest configuration.
enum { WF_ST_PRIMARY = 0, WF_ST_REFR_IN = 1, WF_ST_REFR_OUT = 2, WF_ST_GI_TRACE = 3, WF_ST_SHADOW = 4, WF_ST_OTHER = 5, WF_KEY_DEAD = 63 };
// a lane without a camera sample that found the CTA's claimed work used up in this pass: it asks again in the next one
#define WF_KEY_IDLE 48
__device__ __forceinline__ unsigned wf_stage_slot(int stage)
{
    // ST_REFR 0, ST_REFR_IN_WAIT 1, ST_REFR_OUT_WAIT 2, ST_REFR_WAIT 3, ST_GI 4, ST_GI_TRACE_WAIT 5, ST_GI_WAIT 6, ST_DIRECT 7, ST_SHADOW_WAIT 8
    const unsigned long long table = 0x5455355215ull;      // nibble i = slot of stage i (ST_GATHER_WAIT 9 never waits across a sort)
    return (unsigned)((table >> (4 * stage)) & 15ull);
}
__device__ __forceinline__ unsigned wf_material_class(const bhrt_material &m)
{
    if (m.refraction[0] != 0.f || m.refraction[1] != 0.f || m.refraction[2] != 0.f) return 3u;
    if (m.specular.color[0] != 0.f || m.specular.color[1] != 0.f || m.specular.color[2] != 0.f) return 2u;
    return 1u;
}
// key of a lane whose pending ray has been answered (hit_ok / hit node known)
__device__ __forceinline__ unsigned wf_key(const SceneView &s, const Lane &L, const Frame &top)
{
    if (L.sp == 0) {
        const unsigned cls = (L.hit.node >= 0) ? wf_material_class(s.materials[s.nodes[L.hit.node].material]) : 0u;
        return (WF_ST_PRIMARY << 3) | (cls << 1);
    }
    const unsigned slot = wf_stage_slot(lane_stage(L));
    unsigned cls;
    if (slot == WF_ST_SHADOW || slot == WF_ST_OTHER) cls = wf_material_class(s.materials[top.mtl]);
    else cls = (L.q.kind == Q_CLOSEST && L.hit.node >= 0) ? wf_material_class(s.materials[s.nodes[L.hit.node].material]) : 0u;
    return (slot << 3) | (cls << 1) | (top.gi <= 0 ? 1u : 0u);
}
// the same after a BVH traversal changed the answer: only the class bits of a closest-hit stage move
__device__ __forceinline__ unsigned wf_key_patch(const SceneView &s, unsigned key, int hit_node)
{
    const unsigned slot = key >> 3;
    if (slot == WF_ST_SHADOW || slot == WF_ST_OTHER) return key;
    const unsigned cls = hit_node >= 0 ? wf_material_class(s.materials[s.nodes[hit_node].material]) : 0u;
    return (key & ~0x6u) | (cls << 1);
}

struct WaveCounters {
    unsigned long long rays_closest, rays_shadow;
    unsigned long long next_work;           // next camera-sample work item of the chunk
    unsigned long long depth_overflow;      // Shade() levels refused because the lanes' frame stacks were full (the host retries deeper)
    // clock64() cycles summed over the CTAs (warp 0): lanes pass, sort, BVH traversal; iterations, traversal phases
    unsigned long long prof[6];
    unsigned long long progress;            // camera samples claimed by the CTAs so far in this frame (all chunks; bhrt_progress)
    unsigned long long pmprof[16];          // -DBHRT_PM_PROF: per-phase clocks / counts of the caustic gather (see estimate_irradiance_warp)
};

struct WaveArgs {
    ShadeCfg cfg;
    unsigned long long seed;
    int sample_first;           // global index of the chunk's first sample
    int chunk_spp;              // samples per pixel in this chunk
    int row_begin, row_stride, n_rows, tiles_x;
    int stiles_x;               // super-tiles (WF_ST x WF_ST tiles of 8x8 pixels) per row of super-tiles
    unsigned batch;             // work items a CTA claims from the global counter at a time (a multiple of one super-tile)
    unsigned n_pix_work;        // tiles * 64 (work items outside the image are skipped)
    unsigned long long n_work;  // n_pix_work * chunk_spp
    PhotonGather gather;
    Lane *lanes;                // [gridDim.x][pool] lane headers
    Frame *frames;              // [gridDim.x][cfg.max_frames][pool] recursion levels (level-major: only the low levels are ever touched)
    float4 *samples;            // [chunk_spp][n_pix_work]
    int *hit_id; float *zbuf;
    WaveCounters *cnt;
    const unsigned *cancel;     // set by bhrt_cancel: stop claiming camera samples
    unsigned pool;              // lanes per CTA (<= 65535)
    int mesh_budget;            // traversal trips a thread may still run after the last warp left the lanes pass
    unsigned mesh_batch;        // queued rays that switch the traversal phase on
    unsigned gather_batch;      // parked caustic queries that switch the gather phase on
    int trav_min, trav_max, trav_pct;   // internal-node steps per trip: at least / at most / until this % of the job lanes is still at an internal node
    int smem_nodes, smem_materials, smem_lights, smem_offsets, smem_wave, smem_gather, smem_oct, smem_hot;   // byte offsets into dynamic shared memory
    // ---- split gather: the caustic queries are NOT evaluated inside this kernel.  A CTA runs until every one of
    // its lanes is parked at a query (or dead), appends its queries to gq_list, saves what it needs to carry on and EXITS;
    // the host sorts the queries of ALL CTAs by photon cell, pm_gather_queries evaluates them (neighbouring warps stream
    // the same photon rows, which the in-kernel phase could only arrange for the primary hits of a CTA's image patch),
    // and the kernel is launched again with `resume` set.
    int ext_gather, resume, gather_prof;
    unsigned key_mask;          // bits of the sort key in use (0x3f all; 0x38 = the stage only -- fewer, fuller chunks, less coherent branches)
    unsigned exit_live;         // split gather: a CTA leaves when no more than this many of its lanes are still live
    unsigned gather_defer;      // records per pass above which a query goes to the CTA-wide routine (estimate_irradiance_cta)
    unsigned *gq_list;          // [gridDim.x * pool] global lane index (CTA * pool + lane) of every parked query of this round
    unsigned *gq_total;         // number of entries
    unsigned char *cta_state;   // [gridDim.x][wf_state_bytes(pool)] per-CTA state carried from one launch to the next
};
struct WfClaim;
__host__ __device__ static inline size_t wf_state_bytes(unsigned pool) { return (((size_t)pool + 15u) & ~(size_t)15u) + 256u; }     // lane keys + the work ring

// dynamic shared memory needed by the CTA-local wavefront for `pool` lanes
static inline int wf_wave_smem_bytes(unsigned pool, bool gather)
{
    unsigned counts = (pool / 32 + 1) * WF_NKEYS * 2;       // group counts of the sort; with a photon map also the 4096 cell counters of the gather's query sort
    if (gather && counts < 4096u * 4u) counts = 4096u * 4u;
    return (int)(pool * 2 /*list*/ + pool * 2 /*mesh queue*/ + pool /*key*/ + pool /*pending*/ + counts + 16 + (WF_PREFETCH_FRAME ? pool : 0u) /*frame stack depths*/);
}

// Stage the small, uniformly-read scene tables in shared memory (node table, materials, lights, texture-filter
// taps) and redirect the view to them.  The four tables arrive as TMA bulk copies (cp.async.bulk, global -> shared,
// completion counted in bytes on an mbarrier): one thread issues four instructions and the copy engine does the rest
// while the CTA initialises its lists -- with the split gather a photon-mapped frame launches this kernel a couple of
// hundred times, so the prologue is no longer paid once per frame.  Sizes are rounded up to the 16 bytes a bulk copy
// moves (the device buffers and the shared-memory regions are padded accordingly: render.cu upload(), al16()).
__device__ __forceinline__ uint32_t wf_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wf_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(wf_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void stage_tables_begin(const SceneView &s, const WaveArgs &a, unsigned char *smem, unsigned long long *bar_storage)
{
    const uint32_t bar = wf_smem_u32(bar_storage);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t nb[4] = { ((uint32_t)s.n_nodes * (uint32_t)sizeof(bhrt_node) + 15u) & ~15u, ((uint32_t)s.n_materials * (uint32_t)sizeof(bhrt_material) + 15u) & ~15u,
                                 ((uint32_t)s.n_lights * (uint32_t)sizeof(bhrt_light) + 15u) & ~15u, (2u * (BHRT_TEX_SAMPLES - 1) * 4u + 15u) & ~15u };
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(nb[0] + nb[1] + nb[2] + nb[3]) : "memory");
        if (nb[0]) wf_bulk_g2s(smem + a.smem_nodes, s.nodes, nb[0], bar);
        if (nb[1]) wf_bulk_g2s(smem + a.smem_materials, s.materials, nb[1], bar);
        if (nb[2]) wf_bulk_g2s(smem + a.smem_lights, s.lights, nb[2], bar);
        if (nb[3]) wf_bulk_g2s(smem + a.smem_offsets, s.tex_offsets, nb[3], bar);
    }
}
// every thread: wait for the bytes (phase 0 of the barrier), then read the tables from shared memory
__device__ __forceinline__ void stage_tables_end(SceneView &s, const WaveArgs &a, unsigned char *smem, unsigned long long *bar_storage)
{
    const uint32_t bar = wf_smem_u32(bar_storage);
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    } while (!ok);
    s.nodes = (const bhrt_node *)(smem + a.smem_nodes);
    s.materials = (const bhrt_material *)(smem + a.smem_materials);
    s.lights = (const bhrt_light *)(smem + a.smem_lights);
    s.tex_offsets = (const float *)(smem + a.smem_offsets);
}

// 128-bit copies between a lane's records in global memory and the thread's own copy.  All loads
// of a record are issued back to back (memory-level parallelism) instead of one dependent scalar
// load per field.
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
    static_assert(sizeof(T) % 16 == 0, "record size must be a multiple of 16 bytes");
    const uint4 *s4 = reinterpret_cast<const uint4 *>(&src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d4[i] = s4[i];
}
// the persistent part of a lane record only (core.cuh: BHRT_LANE_PERSIST)
__device__ __forceinline__ void wf_load_lane(Lane &dst, const Lane *src)
{
    static_assert(offsetof(Lane, hit) + 32 == BHRT_LANE_PERSIST && BHRT_LANE_PERSIST % 16 == 0 && offsetof(Hit, p) >= 28, "Lane layout");
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(&dst);
    // (streaming loads / stores: a record is read once and written once per pass; it should not push
    // the thread-local copies out of L1)
#pragma unroll
    for (int i = 0; i < BHRT_LANE_PERSIST / 16; i++) d4[i] = __ldcs(s4 + i);
}
__device__ __forceinline__ void wf_store_lane(Lane *dst, const Lane &src)
{
    const uint4 *s4 = reinterpret_cast<const uint4 *>(&src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < BHRT_LANE_PERSIST / 16; i++) __stcs(d4 + i, s4[i]);
}
// Work item g of the chunk -> (pixel, sample).  Sample-major; within a sample the items walk SUPER-TILES of
// WF_ST x WF_ST tiles of 8x8 pixels (64 x 64 pixels, 4096 items), tile by tile: consecutive items are the pixels of
// one tile, so the lanes of a warp start on neighbouring primary rays, and a CTA -- which claims whole batches of
// consecutive items, see wf_claim -- works on a compact patch of the image: its lanes meet the same textures, BVH
// subtrees and (above all) photon cells, which is what makes the caustic gather cacheable.
#define WF_ST 8
__device__ __forceinline__ bool work_to_pixel(const SceneView &s, const WaveArgs &a, unsigned long long g, int &pi, int &pj, int &sidx)
{
    const unsigned w = (unsigned)(g % a.n_pix_work);
    sidx = a.sample_first + (int)(g / a.n_pix_work);
    const unsigned tile = w >> 6, in = w & 63u;
    const unsigned st = tile / (WF_ST * WF_ST), lt = tile % (WF_ST * WF_ST);
    const int tx = (int)(st % (unsigned)a.stiles_x) * WF_ST + (int)(lt % WF_ST), ty = (int)(st / (unsigned)a.stiles_x) * WF_ST + (int)(lt / WF_ST);
    const int i = tx * 8 + (int)(in & 7u), trow = ty * 8 + (int)(in >> 3);
    if (i >= s.width || trow >= a.n_rows) return false;
    pi = i; pj = a.row_begin + trow * a.row_stride;
    return true;
}

// A CTA takes its work from the global counter in batches of a.batch consecutive items and hands them to its warps
// from shared memory.  WF_RING batches can be claimed ahead; thread 0 tops the ring up between passes so that a
// pass never runs dry while the frame still has work (at most `pool` lanes ask for an item in one pass).
#define WF_RING 8
struct WfClaim {
    unsigned long long base[WF_RING];   // first global item of the batch in ring slot k
    unsigned long long taken;           // virtual items handed to warps so far (monotonic)
    unsigned long long claimed;         // virtual items claimed from the global counter so far (multiple of a.batch)
};
static_assert(sizeof(WfClaim) <= 256, "wf_state_bytes reserves 256 bytes for the work ring");
__device__ __forceinline__ void wf_claim_topup(WfClaim &c, const WaveArgs &a)
{
    // (thread 0 only, between passes)
    while (c.claimed - c.taken < (unsigned long long)a.pool + 32ull && c.claimed - c.taken + a.batch <= (unsigned long long)WF_RING * a.batch) {
        // (a cancelled frame -- bhrt_cancel, the reference's StopRender -- hands out nothing any more)
        const bool cancelled = *(const volatile unsigned *)a.cancel != 0u;
        const unsigned long long g = cancelled ? a.n_work : atomicAdd(&a.cnt->next_work, (unsigned long long)a.batch);
        c.base[(c.claimed / a.batch) % WF_RING] = g;
        c.claimed += a.batch;
        if (g >= a.n_work) break;       // the frame has no work left: everything claimed from here on is past the end
        atomicAdd(&a.cnt->progress, g + a.batch <= a.n_work ? (unsigned long long)a.batch : a.n_work - g);
    }
}
// virtual item v of this CTA -> global item (>= a.n_work: no work)
__device__ __forceinline__ unsigned long long wf_claim_item(const WfClaim &c, const WaveArgs &a, unsigned long long v)
{
    if (v >= c.claimed) return ~0ull;
    return c.base[(v / a.batch) % WF_RING] + v % a.batch;
}

// ---- caustic irradiance estimate, one query per WARP ------------------------------------------------
// estimate_irradiance (core.cuh) restated for 32 cooperating lanes: the same selection -- the k nearest
// photons within the radius with dir.N < 0, constant filter (cyPhotonMap.h:332-382,421-498) -- but the
// lanes stride together over the photons of the 27 cells (32-byte records, coalesced), the 32-bin
// histograms of the radix select live in shared memory, and the sums are warp reductions.  A thread
// running the scalar version visits the same tens of thousands of photons alone, with its histogram
// in thread-local memory; the reference spends 85 % of a photon-mapped render here.
// All arguments are warp-uniform; every lane returns the same result.  `wsm` = this warp's scratch.
struct PmWarpScratch {
    int hist[256]; float cd2[BHRT_PM_CAND]; uint32_t cidx[BHRT_PM_CAND]; int ncand; int pad[3];
    uint32_t qidx[64]; float qd2[64];       // ring of photons that passed the distance test (pm_scan)
};

// One pass over the photons of the query's 3x3x3 block with WARP-LEVEL COMPACTION.  Only about a third of the block's
// photons lie within the search bound; testing that costs a load and seven instructions, everything after it (the
// direction test, the power unpacking, sums, histogram) five times as much.  So the lanes stride over the records doing
// only the distance test, the survivors' indices go into a 64-entry ring in shared memory, and whenever the ring holds
// 32 of them `body(index, d2, true)` runs for a FULL warp (and once more, partially filled, at the end).
// my_b / my_e: lane n < 27 holds the record range of cell n.  Warp-convergent.
// STREAM = false (what every caller uses): four records per lane and trip, the loads issued back to back.
// STREAM = true: one group of 32 records per iteration with the next group's load in flight -- fewer instructions (the four
// predicated loads, their addresses and fill values are half of a trip's), same groups, same order, same result; tried for
// the dedicated kernels, whose limit is instruction issue.  MEASURED on BASELINE config 5: 3.22 s per frame against 2.88 s
// -- even at a 90 % L1 hit rate the four loads in flight are worth more than the instructions they cost.  Kept, unused.
template <bool STREAM, typename Body>
__device__ __forceinline__ void pm_scan(const PhotonGather &pg, uint32_t my_b, uint32_t my_e, int n_step, V3 pos, float bound2, PmWarpScratch *wsm, unsigned lane, Body body)
{
    unsigned head = 0, fill = 0;        // ring state (warp-uniform)
    // one group: distance test, survivors into the ring, a full ring runs the body
    auto group = [&](uint32_t i, bool in_range, const F4 &A) {
        const float d2 = length2(V3(A.x - pos.x, A.y - pos.y, A.z - pos.z));
        const bool ok = in_range && d2 < bound2;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) { const unsigned slot = (head + fill + __popc(m & ((1u << lane) - 1u))) & 63u; wsm->qidx[slot] = i; wsm->qd2[slot] = d2; }
        fill += __popc(m);
        if (fill >= 32u) {
            __syncwarp();
            const unsigned slot = (head + lane) & 63u;
            body(wsm->qidx[slot], wsm->qd2[slot], true);
            __syncwarp();
            head += 32u; fill -= 32u;
        }
    };
    for (int n = 0; n < 27; n += n_step) {
        const uint32_t b = __shfl_sync(0xffffffffu, my_b, n), e = __shfl_sync(0xffffffffu, my_e, n);
        if (STREAM) {
            if (b >= e) continue;
            uint32_t base = b;
            F4 A; A.x = A.y = A.z = A.w = 0.f;
            if (base + lane < e) A = ld4(&pg.recs[base + lane].a);
            for (;;) {
                const uint32_t next = base + 32u;
                const bool more = next < e;         // (warp-uniform)
                F4 An; An.x = An.y = An.z = An.w = 0.f;
                if (more && next + lane < e) An = ld4(&pg.recs[next + lane].a);
                group(base + lane, base + lane < e, A);
                if (!more) break;
                A = An; base = next;
            }
        } else {
            for (uint32_t base = b; base < e; base += 128u) {
                F4 A[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t i = base + 32u * k + lane;
                    if (i < e) A[k] = ld4(&pg.recs[i].a); else { A[k].x = A[k].y = A[k].z = 3.0e18f; A[k].w = 0.f; }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (base + 32u * k >= e) break;         // (warp-uniform)
                    group(base + 32u * k + lane, true, A[k]);       // (records past the end sit at 3e18: never inside)
                }
            }
        }
    }
    __syncwarp();
    if (fill) {     // (warp-uniform)
        const unsigned slot = (head + lane) & 63u;
        const bool act = lane < fill;
        body(act ? wsm->qidx[slot] : 0u, act ? wsm->qd2[slot] : 0.f, act);
    }
    __syncwarp();
}

// After pass A the k-th nearest photon is known to lie in distance bin `top` of 256 (the top 8 bits of the 24-bit key
// u = d2 / bound2 * 2^24): the later passes only need the photons up to that bin, so their distance test -- all that
// most records ever see -- can use (top + 1) / 256 of the bound (with a margin for the rounding of u; the body still
// decides on the exact key).  In a caustic's focus that turns three of the four passes into plain scans.
__device__ __forceinline__ float pm_tight_bound(float bound2, uint32_t top) { return fminf(bound2, bound2 * (((float)top + 1.01f) * (1.0f / 256.0f))); }
__device__ __forceinline__ float wf_sum(float v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
__device__ __forceinline__ float wf_max(float v) { for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ V3 wf_sum3(V3 v) { return V3(wf_sum(v.x), wf_sum(v.y), wf_sum(v.z)); }

#ifdef BHRT_PM_PROF
#define PM_PROF_ADD(i, v) do { if (lane == 0) atomicAdd(&prof[i], (unsigned long long)(v)); } while (0)
#define PM_PROF_MARK(i) do { const long long c_ = clock64(); PM_PROF_ADD(i, c_ - pm_t); pm_t = c_; } while (0)
#else
#define PM_PROF_ADD(i, v)
#define PM_PROF_MARK(i)
#endif
// The record ranges of one search level's block (pm_ranges of core.cuh, one range per lane): level 0 -> lane n < 25 owns ROW n
// of the 5x5x5 block of coarse cells (five x-adjacent cells = one contiguous range); level >= 1 -> lane n < 27 owns cell n of the
// 3x3x3 block at resolution level - 1.  bound2 = the squared distance up to which the block is complete.
__device__ __forceinline__ void pm_level_ranges(const PhotonGather &pg, int level, const int (&c)[3], const int (&sub)[3], V3 pos, float r2, unsigned lane,
                                                uint32_t &my_b_out, uint32_t &my_e_out, float &bound2_out)
{
    uint32_t my_b = 0, my_e = 0;
    const int res = BHRT_PM_RES(level);
    const float cell_l = pg.cell / (float)(1 << res);
    // every photon this close lies in the block: level 0 reaches two coarse cells (>= the radius), so the radius itself is
    // the bound there; a 3x3x3 block is complete up to one cell edge
    float bound2 = level == 0 ? r2 : 0.998f * cell_l * cell_l;
    if (bound2 > r2) bound2 = r2;
    const float q3[3] = { pos.x, pos.y, pos.z };
    if (level == 0) {
        if (lane < 25u) {
            const int dz = (int)lane / 5 - 2, dy = (int)lane % 5 - 2;
            const int y = c[1] + dy, z = c[2] + dz;
            int x0 = c[0] - 2, x1 = c[0] + 2;
            if (x0 < 0) x0 = 0;
            if (x1 > pg.dim[0] - 1) x1 = pg.dim[0] - 1;
            if (!(y < 0 || z < 0 || y >= pg.dim[1] || z >= pg.dim[2] || x0 > x1)) {
                // a row farther from the query than the radius holds nothing of interest (the edges of the block), and of a
                // row that passes by at distance d only the cells within sqrt(r^2 - d^2) along x do (a third fewer records)
                float dyz2 = 0;
                {
                    const int lo2[2] = { y, z };
                    for (int ax = 1; ax < 3; ax++) {
                        const float lo = pg.origin[ax] + (float)lo2[ax - 1] * pg.cell, hi = lo + pg.cell;
                        const float d = q3[ax] < lo ? lo - q3[ax] : (q3[ax] > hi ? q3[ax] - hi : 0.f);
                        dyz2 += d * d;
                    }
                }
                if (!(dyz2 * 0.999f >= bound2)) {
                    const float xr = sqrtf(fmaxf(bound2 - dyz2 * 0.999f, 0.f)) * 1.001f + 1.0e-6f * pg.cell;
                    const int xa = (int)floorf((q3[0] - xr - pg.origin[0]) / pg.cell), xb = (int)floorf((q3[0] + xr - pg.origin[0]) / pg.cell);
                    if (xa > x0) x0 = xa;
                    if (xb < x1) x1 = xb;
                    if (x0 <= x1) {
                        const uint32_t row = (uint32_t)pg.dim[0] * ((uint32_t)y + (uint32_t)pg.dim[1] * (uint32_t)z);
                        my_b = pg.coarse_start[row + (uint32_t)x0]; my_e = pg.coarse_start[row + (uint32_t)x1 + 1u];
                    }
                }
            }
        }
    } else if (lane < 27u) {
        const int sh = BHRT_PM_SUB_BITS - res, span = 1 << res;
        const int dz = (int)lane / 9 - 1, dy = ((int)lane / 3) % 3 - 1, dx = (int)lane % 3 - 1;
        const int x = ((c[0] << res) | (sub[0] >> sh)) + dx, y = ((c[1] << res) | (sub[1] >> sh)) + dy, z = ((c[2] << res) | (sub[2] >> sh)) + dz;
        if (!(x < 0 || y < 0 || z < 0 || x >= pg.dim[0] * span || y >= pg.dim[1] * span || z >= pg.dim[2] * span)) {
            const uint32_t cx = (uint32_t)(x >> res), cy = (uint32_t)(y >> res), cz = (uint32_t)(z >> res);
            const uint32_t coarse = cx + (uint32_t)pg.dim[0] * (cy + (uint32_t)pg.dim[1] * cz);
            const uint32_t lo = pg.coarse_start[coarse], hi = pg.coarse_start[coarse + 1];
            // a cell farther from the query than the search bound holds nothing of interest (the corners of the block, mostly)
            const int g3[3] = { x, y, z };
            float md2 = 0;
            for (int ax = 0; ax < 3; ax++) {
                const float l_ = pg.origin[ax] + (float)g3[ax] * cell_l, h_ = l_ + cell_l;
                const float d = q3[ax] < l_ ? l_ - q3[ax] : (q3[ax] > h_ ? q3[ax] - h_ : 0.f);
                md2 += d * d;
            }
            if (lo != hi && !(md2 * 0.999f >= bound2)) {
                if (res == 0) { my_b = lo; my_e = hi; }
                else {
                    const uint32_t m = photon_morton((uint32_t)(x & (span - 1)) << sh, (uint32_t)(y & (span - 1)) << sh, (uint32_t)(z & (span - 1)) << sh);
                    const uint32_t k0 = (coarse << BHRT_PM_FINE_BITS) | m, k1 = k0 + (1u << (3 * sh));
                    my_b = pm_lower_bound(pg.keys, lo, hi, k0);
                    my_e = pm_lower_bound(pg.keys, my_b, hi, k1);
                }
            }
        }
    }
    my_b_out = my_b; my_e_out = my_e; bound2_out = bound2;
}

// defer_visit > 0: a query that would have to scan more than that many records per pass is NOT evaluated (returns true): the
// caller hands it to estimate_irradiance_cta, where a whole CTA shares the scan -- one warp alone needs tens of milliseconds
// for a query in a caustic's focus, and everybody else ends up waiting for it.
// (inlined into the dedicated gather kernel, where the structure's fields are kernel parameters -- uniform registers, no
// reloads of the record pointer from a local copy; wf_render calls it out of line, below)
template <bool STREAM>
__device__ __forceinline__ bool estimate_irradiance_warp_inl(const PhotonGather &pg, V3 pos, V3 N, V3 &irrad, V3 &direction, PmWarpScratch *wsm, unsigned long long *prof, unsigned defer_visit)
{
#ifdef BHRT_PM_PROF
    long long pm_t = clock64();
#endif
    const unsigned lane = threadIdx.x & 31u;
    irrad = V3(0, 0, 0); direction = V3(0, 0, 0);
    if (pg.n_photons <= 0) return false;
    const float r2 = pg.radius * pg.radius;
    float f[3] = { (pos.x - pg.origin[0]) / pg.cell, (pos.y - pg.origin[1]) / pg.cell, (pos.z - pg.origin[2]) / pg.cell };
    int c[3], sub[3];
    for (int a = 0; a < 3; a++) {
        c[a] = (int)floorf(f[a]);
        if (c[a] < -2 || c[a] > pg.dim[a] + 1) return false;      // farther than two cells (>= the radius) from the grid
        int s_ = (int)((f[a] - c[a]) * (float)(1 << BHRT_PM_SUB_BITS));
        sub[a] = s_ < 0 ? 0 : (s_ > (1 << BHRT_PM_SUB_BITS) - 1 ? (1 << BHRT_PM_SUB_BITS) - 1 : s_);
    }
    // start at the coarse level; the loop below moves to finer levels, before reading any photon, while
    // the 3x3x3 block holds far more photons than a query needs
    int level = 0;
    bool coarsened = false;         // a finer level held too few photons: do not try it again
    PM_PROF_ADD(0, 1);
    for (;;) {
        PM_PROF_ADD(6, 1);
        uint32_t my_b, my_e; float bound2;
        pm_level_ranges(pg, level, c, sub, pos, r2, lane, my_b, my_e, bound2);
        const int n_step = 1;       // (walking four ranges at a time, eight lanes each, was measured: 12.8 s instead of 6.2 s per frame)
        // far more photons in the block than the k a query needs: the next finer level first
        {
            unsigned visit = my_e - my_b;
            for (int o = 16; o > 0; o >>= 1) visit += __shfl_xor_sync(0xffffffffu, visit, o);
            if (!coarsened && level < BHRT_PM_LEVELS - 1 && visit > 12u * (unsigned)pg.k) { PM_PROF_MARK(1); level++; continue; }
            if (defer_visit && visit > defer_visit) return true;
            PM_PROF_ADD(5, visit);
        }
        PM_PROF_MARK(1);
        const float scale = 32.f / bound2;
        // ---- pass A (256-bin histogram of a 24-bit distance key: the crossing bin then usually holds
        // few enough photons for the exact ranking, so a query is two passes over its photons)
        for (int j = (int)lane; j < 256; j += 32) wsm->hist[j] = 0;
        __syncwarp();
        int total = 0;
        V3 sumP(0, 0, 0), sumD(0, 0, 0);
        pm_scan<STREAM>(pg, my_b, my_e, n_step, pos, bound2, wsm, lane, [&](uint32_t i, float d2, bool act) {
            if (!act) return;
            const F4 B = ld4(&pg.recs[i].b);
            if (dot(V3(B.x, B.y, B.z), N) >= 0) return;
            uint32_t u = (uint32_t)((d2 * scale) * 524288.f); if (u > 0xffffffu) u = 0xffffffu;
            atomicAdd(&wsm->hist[u >> 16], 1);
            total++;
            if (level == 0) { const float pw = ld4(&pg.recs[i].a).w; sumP = sumP + pm_power(&pg.recs[i], pw); sumD = sumD + V3(B.x, B.y, B.z) * pw; }
        });
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        __syncwarp();
        PM_PROF_MARK(2);
        if (total <= pg.k) {
            PM_PROF_ADD(7, 1);
            if (level > 0) { level--; coarsened = true; continue; }
            if (total == 0) return false;
            sumP = wf_sum3(sumP); sumD = wf_sum3(sumD);
            irrad = sumP * (1.0f / (3.14159265358979323846f * r2));
            direction = normalized(sumD);
            return false;
        }
        if (!coarsened && level < BHRT_PM_LEVELS - 1 && total > 8 * pg.k) { PM_PROF_ADD(10, 1); level++; continue; }
        PM_PROF_ADD(11, total);
        // ---- radix-256 select of the k-th smallest squared distance (u = 24-bit key, top digit = pass A's bin)
        uint32_t prefix = 0;
        int shift = 16, need = pg.k;
        for (;;) {
            // each lane owns 8 consecutive bins
            int hb8[8], mine = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { hb8[j] = wsm->hist[lane * 8 + j]; mine += hb8[j]; }
            int incl = mine;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
            const unsigned m = __ballot_sync(0xffffffffu, incl >= need);
            const int owner = m ? __ffs(m) - 1 : 31;
            int b = 0, cum = 0, hb = 0;
            if ((int)lane == owner) {
                cum = incl - mine;
                int j = 0;
                for (; j < 7; j++) { if (cum + hb8[j] >= need) break; cum += hb8[j]; }
                b = (int)lane * 8 + j; hb = hb8[j];
            }
            b = __shfl_sync(0xffffffffu, b, owner); cum = __shfl_sync(0xffffffffu, cum, owner); hb = __shfl_sync(0xffffffffu, hb, owner);
            need -= cum;
            prefix = (prefix << 8) | (uint32_t)b;
            if (hb <= BHRT_PM_CAND || shift == 0) break;
            PM_PROF_ADD(9, 1);
            shift -= 8;
            __syncwarp();
            for (int j = (int)lane; j < 256; j += 32) wsm->hist[j] = 0;
            __syncwarp();
            pm_scan<STREAM>(pg, my_b, my_e, n_step, pos, pm_tight_bound(bound2, prefix >> (16 - shift - 8)), wsm, lane, [&](uint32_t i, float d2, bool act) {
                if (!act) return;
                uint32_t u = (uint32_t)((d2 * scale) * 524288.f); if (u > 0xffffffu) u = 0xffffffu;
                if ((u >> (shift + 8)) != prefix) return;
                const F4 B = ld4(&pg.recs[i].b);
                if (dot(V3(B.x, B.y, B.z), N) >= 0) return;
                atomicAdd(&wsm->hist[(u >> shift) & 255u], 1);
            });
            __syncwarp();
        }
        // ---- last pass: everything below the crossing bin is in; the crossing bin's photons are candidates
        PM_PROF_MARK(3);
        if (lane == 0) wsm->ncand = 0;
        __syncwarp();
        sumP = V3(0, 0, 0); sumD = V3(0, 0, 0);
        V3 subP(0, 0, 0), subD(0, 0, 0);
        float kth2 = 0, sub_max = 0;
        pm_scan<STREAM>(pg, my_b, my_e, n_step, pos, pm_tight_bound(bound2, prefix >> (16 - shift)), wsm, lane, [&](uint32_t i, float d2, bool act) {
            if (!act) return;
            uint32_t u = (uint32_t)((d2 * scale) * 524288.f); if (u > 0xffffffu) u = 0xffffffu;
            const uint32_t top = u >> shift;
            if (top > prefix) return;
            const F4 B = ld4(&pg.recs[i].b);
            if (dot(V3(B.x, B.y, B.z), N) >= 0) return;
            const float pw = ld4(&pg.recs[i].a).w;
            if (top < prefix) {
                sumP = sumP + pm_power(&pg.recs[i], pw); sumD = sumD + V3(B.x, B.y, B.z) * pw;
                kth2 = fmaxf(kth2, d2);
            } else {
                subP = subP + pm_power(&pg.recs[i], pw); subD = subD + V3(B.x, B.y, B.z) * pw;
                sub_max = fmaxf(sub_max, d2);
                const int slot = atomicAdd(&wsm->ncand, 1);
                if (slot < BHRT_PM_CAND) { wsm->cd2[slot] = d2; wsm->cidx[slot] = i; }
            }
        });
        __syncwarp();
        const int ncand = wsm->ncand;
        if (ncand > BHRT_PM_CAND) {
            // more than 48 photons share all 20 key bits (coincident positions): keep them all
            sumP = sumP + subP; sumD = sumD + subD;
            kth2 = fmaxf(kth2, sub_max);
        } else {
            // exact: a candidate is in iff fewer than `need` candidates are nearer (ties: the lower record index).  The slots
            // were handed out in the order the atomics arrived; ranks are unique, and lane r adds the candidate of RANK r, so
            // neither the chosen set nor the order of the float sums depends on that order
            int *const order = wsm->hist;       // (the histogram is not needed any more)
            for (int j = (int)lane; j < ncand; j += 32) {
                const float dj = wsm->cd2[j]; const uint32_t xj = wsm->cidx[j];
                int rank = 0;
                for (int i = 0; i < ncand; i++) { const float di = wsm->cd2[i]; rank += (di < dj || (di == dj && wsm->cidx[i] < xj)) ? 1 : 0; }
                order[rank] = j;
            }
            __syncwarp();
            for (int r = (int)lane; r < ncand && r < need; r += 32) {
                const int j = order[r];
                const uint32_t ix = wsm->cidx[j];
                F4 A = ld4(&pg.recs[ix].a); F4 B = ld4(&pg.recs[ix].b);
                sumP = sumP + pm_power(&pg.recs[ix], A.w); sumD = sumD + V3(B.x, B.y, B.z) * A.w;
                kth2 = fmaxf(kth2, wsm->cd2[j]);
            }
        }
        sumP = wf_sum3(sumP); sumD = wf_sum3(sumD); kth2 = wf_max(kth2);
        __syncwarp();
        const float area = 3.14159265358979323846f * kth2;
        irrad = area > 0 ? sumP * (1.0f / area) : sumP;
        direction = normalized(sumD);
        PM_PROF_MARK(4);
        PM_PROF_ADD(8, 1);
        return false;
    }
}

__device__ __noinline__ bool estimate_irradiance_warp(const PhotonGather &pg, V3 pos, V3 N, V3 &irrad, V3 &direction, PmWarpScratch *wsm, unsigned long long *prof, unsigned defer_visit = 0u)
{
    return estimate_irradiance_warp_inl<false>(pg, pos, N, irrad, direction, wsm, prof, defer_visit);
}

// ---- the same estimate for a WHOLE CTA: every warp scans its own slice of each record range, the histogram and the
// candidate list are shared, and the per-lane sums meet in a cross-warp reduction after every pass.  For the queries the
// warp routine defers (a caustic's focus: hundreds of thousands of photons within the finest block).  Every decision is
// taken from CTA-uniform values, so all warps follow the same path and the barriers line up; all threads return the same.
// scr[0] holds the shared tables, scr[warp] the warp's compaction ring; red = [warps][16] floats.
__device__ __forceinline__ void pm_cta_reduce(float *v, int n_sum, int n_max, float (*red)[16], unsigned w, unsigned W, unsigned lane)
{
    // v[0 .. n_sum) are summed, v[n_sum .. n_sum + n_max) maximised: over the lanes, then over the warps (in warp order)
    for (int i = 0; i < n_sum; i++) v[i] = wf_sum(v[i]);
    for (int i = n_sum; i < n_sum + n_max; i++) v[i] = wf_max(v[i]);
    __syncthreads();
    if (lane == 0) for (int i = 0; i < n_sum + n_max; i++) red[w][i] = v[i];
    __syncthreads();
    for (int i = 0; i < n_sum + n_max; i++) {
        float acc = red[0][i];
        for (unsigned k = 1; k < W; k++) acc = i < n_sum ? acc + red[k][i] : fmaxf(acc, red[k][i]);
        v[i] = acc;
    }
}
__device__ __forceinline__ void estimate_irradiance_cta(const PhotonGather &pg, V3 pos, V3 N, V3 &irrad, V3 &direction, PmWarpScratch *scr, float (*red)[16])
{
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    PmWarpScratch *const sh = scr, *const ring = scr + w;
    irrad = V3(0, 0, 0); direction = V3(0, 0, 0);
    if (pg.n_photons <= 0) return;
    const float r2 = pg.radius * pg.radius;
    float f[3] = { (pos.x - pg.origin[0]) / pg.cell, (pos.y - pg.origin[1]) / pg.cell, (pos.z - pg.origin[2]) / pg.cell };
    int c[3], sub[3];
    for (int a = 0; a < 3; a++) {
        c[a] = (int)floorf(f[a]);
        if (c[a] < -2 || c[a] > pg.dim[a] + 1) return;
        int s_ = (int)((f[a] - c[a]) * (float)(1 << BHRT_PM_SUB_BITS));
        sub[a] = s_ < 0 ? 0 : (s_ > (1 << BHRT_PM_SUB_BITS) - 1 ? (1 << BHRT_PM_SUB_BITS) - 1 : s_);
    }
    int level = 0;
    bool coarsened = false;
    for (;;) {
        uint32_t my_b, my_e; float bound2;
        pm_level_ranges(pg, level, c, sub, pos, r2, lane, my_b, my_e, bound2);
        {
            unsigned visit = my_e - my_b;
            for (int o = 16; o > 0; o >>= 1) visit += __shfl_xor_sync(0xffffffffu, visit, o);
            if (!coarsened && level < BHRT_PM_LEVELS - 1 && visit > 12u * (unsigned)pg.k) { level++; continue; }
        }
        // this warp's slice of every range
        {
            const uint32_t len = my_e - my_b;
            const uint32_t lo = (uint32_t)(((unsigned long long)len * w) / W), hi = (uint32_t)(((unsigned long long)len * (w + 1u)) / W);
            my_e = my_b + hi; my_b = my_b + lo;
        }
        const float scale = 32.f / bound2;
        __syncthreads();
        for (int j = (int)threadIdx.x; j < 256; j += (int)blockDim.x) sh->hist[j] = 0;
        __syncthreads();
        float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };         // total, sumP, sumD
        {
            int total = 0;
            V3 sumP(0, 0, 0), sumD(0, 0, 0);
            pm_scan<false>(pg, my_b, my_e, 1, pos, bound2, ring, lane, [&](uint32_t i, float d2, bool act) {
                if (!act) return;
                const F4 B = ld4(&pg.recs[i].b);
                if (dot(V3(B.x, B.y, B.z), N) >= 0) return;
                uint32_t u = (uint32_t)((d2 * scale) * 524288.f); if (u > 0xffffffu) u = 0xffffffu;
                atomicAdd(&sh->hist[u >> 16], 1);
                total++;
                if (level == 0) { const float pw = ld4(&pg.recs[i].a).w; sumP = sumP + pm_power(&pg.recs[i], pw); sumD = sumD + V3(B.x, B.y, B.z) * pw; }
            });
            acc[0] = (float)total; acc[1] = sumP.x; acc[2] = sumP.y; acc[3] = sumP.z; acc[4] = sumD.x; acc[5] = sumD.y; acc[6] = sumD.z;
        }
        pm_cta_reduce(acc, 7, 0, red, w, W, lane);
        const int total = (int)(acc[0] + 0.5f);             // (counts below 2^24: exact in float)
        if (total <= pg.k) {
            if (level > 0) { level--; coarsened = true; continue; }
            if (total == 0) return;
            irrad = V3(acc[1], acc[2], acc[3]) * (1.0f / (3.14159265358979323846f * r2));
            direction = normalized(V3(acc[4], acc[5], acc[6]));
            return;
        }
        if (!coarsened && level < BHRT_PM_LEVELS - 1 && total > 8 * pg.k) { level++; continue; }
        // ---- radix-256 select of the k-th smallest squared distance: every warp reads the shared histogram
        uint32_t prefix = 0;
        int shift = 16, need = pg.k;
        for (;;) {
            int hb8[8], mine = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { hb8[j] = sh->hist[lane * 8 + j]; mine += hb8[j]; }
            int incl = mine;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
            const unsigned m = __ballot_sync(0xffffffffu, incl >= need);
            const int owner = m ? __ffs(m) - 1 : 31;
            int b = 0, cum = 0, hb = 0;
            if ((int)lane == owner) {
                cum = incl - mine;
                int j = 0;
                for (; j < 7; j++) { if (cum + hb8[j] >= need) break; cum += hb8[j]; }
                b = (int)lane * 8 + j; hb = hb8[j];
            }
            b = __shfl_sync(0xffffffffu, b, owner); cum = __shfl_sync(0xffffffffu, cum, owner); hb = __shfl_sync(0xffffffffu, hb, owner);
            need -= cum;
            prefix = (prefix << 8) | (uint32_t)b;
            if (hb <= BHRT_PM_CAND || shift == 0) break;
            shift -= 8;
            __syncthreads();
            for (int j = (int)threadIdx.x; j < 256; j += (int)blockDim.x) sh->hist[j] = 0;
            __syncthreads();
            pm_scan<false>(pg, my_b, my_e, 1, pos, pm_tight_bound(bound2, prefix >> (16 - shift - 8)), ring, lane, [&](uint32_t i, float d2, bool act) {
                if (!act) return;
                uint32_t u = (uint32_t)((d2 * scale) * 524288.f); if (u > 0xffffffu) u = 0xffffffu;
                if ((u >> (shift + 8)) != prefix) return;
                const F4 B = ld4(&pg.recs[i].b);
                if (dot(V3(B.x, B.y, B.z), N) >= 0) return;
                atomicAdd(&sh->hist[(u >> shift) & 255u], 1);
            });
            __syncthreads();
        }
        // ---- last pass: everything below the crossing bin is in; the crossing bin's photons are candidates
        __syncthreads();
        if (threadIdx.x == 0) sh->ncand = 0;
        __syncthreads();
        V3 sumP(0, 0, 0), sumD(0, 0, 0), subP(0, 0, 0), subD(0, 0, 0);
        float kth2 = 0, sub_max = 0;
        pm_scan<false>(pg, my_b, my_e, 1, pos, pm_tight_bound(bound2, prefix >> (16 - shift)), ring, lane, [&](uint32_t i, float d2, bool act) {
            if (!act) return;
            uint32_t u = (uint32_t)((d2 * scale) * 524288.f); if (u > 0xffffffu) u = 0xffffffu;
            const uint32_t top = u >> shift;
            if (top > prefix) return;
            const F4 B = ld4(&pg.recs[i].b);
            if (dot(V3(B.x, B.y, B.z), N) >= 0) return;
            const float pw = ld4(&pg.recs[i].a).w;
            if (top < prefix) {
                sumP = sumP + pm_power(&pg.recs[i], pw); sumD = sumD + V3(B.x, B.y, B.z) * pw;
                kth2 = fmaxf(kth2, d2);
            } else {
                subP = subP + pm_power(&pg.recs[i], pw); subD = subD + V3(B.x, B.y, B.z) * pw;
                sub_max = fmaxf(sub_max, d2);
                const int slot = atomicAdd(&sh->ncand, 1);
                if (slot < BHRT_PM_CAND) { sh->cd2[slot] = d2; sh->cidx[slot] = i; }
            }
        });
        __syncthreads();
        const int ncand = sh->ncand;
        if (ncand > BHRT_PM_CAND) {
            sumP = sumP + subP; sumD = sumD + subD;
            kth2 = fmaxf(kth2, sub_max);
        } else if (w == 0) {
            // exact, by one warp: a candidate is in iff fewer than `need` candidates are nearer (ties: the lower record index);
            // lane r adds the candidate of rank r (see estimate_irradiance_warp: independent of the order of the atomics)
            int *const order = sh->hist;
            for (int j = (int)lane; j < ncand; j += 32) {
                const float dj = sh->cd2[j]; const uint32_t xj = sh->cidx[j];
                int rank = 0;
                for (int i = 0; i < ncand; i++) { const float di = sh->cd2[i]; rank += (di < dj || (di == dj && sh->cidx[i] < xj)) ? 1 : 0; }
                order[rank] = j;
            }
            __syncwarp();
            for (int r = (int)lane; r < ncand && r < need; r += 32) {
                const int j = order[r];
                const uint32_t ix = sh->cidx[j];
                F4 A = ld4(&pg.recs[ix].a); F4 B = ld4(&pg.recs[ix].b);
                sumP = sumP + pm_power(&pg.recs[ix], A.w); sumD = sumD + V3(B.x, B.y, B.z) * A.w;
                kth2 = fmaxf(kth2, sh->cd2[j]);
            }
        }
        float fin[7] = { sumP.x, sumP.y, sumP.z, sumD.x, sumD.y, sumD.z, kth2 };
        pm_cta_reduce(fin, 6, 1, red, w, W, lane);
        const float area = 3.14159265358979323846f * fin[6];
        const V3 sp(fin[0], fin[1], fin[2]);
        irrad = area > 0 ? sp * (1.0f / area) : sp;
        direction = normalized(V3(fin[3], fin[4], fin[5]));
        return;
    }
}

// ---- the persistent kernel ------------------------------------------------------------------------
__global__ void __launch_bounds__(WF_BLOCK, WF_MIN_BLOCKS)
wf_render(SceneView s, WaveArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) unsigned long long tma_bar;
    stage_tables_begin(s, a, smem, &tma_bar);       // (the copies run while the lists below are initialised)
    const unsigned P = a.pool;
    unsigned short *list = (unsigned short *)(smem + a.smem_wave);         // [P] live lanes sorted by key
    unsigned short *mq = list + P;                                         // [P] mesh queue (ring)
    unsigned char *lkey = (unsigned char *)(mq + P);                       // [P] key of lane id
    unsigned char *pend = lkey + P;                                        // [P] lane id is queued / being traversed
    PmWarpScratch *pm_scratch = (PmWarpScratch *)(smem + a.smem_gather);   // [WF_BLOCK / 32]   (only with a photon map)
    // [P] lanes waiting for a caustic gather (split gather: the only thing this kernel keeps of the gather)
    unsigned short *gq = a.ext_gather ? (unsigned short *)(smem + a.smem_gather) : (unsigned short *)(pm_scratch + WF_BLOCK / 32);
    unsigned short *gcount = (unsigned short *)(pend + P);                 // [P/32][WF_NKEYS] per-group key counts / offsets (P is a multiple of 32)
#if WF_PREFETCH_FRAME
    unsigned char *lsp = smem + a.smem_gather - P - 16;                     // [P] min(frames in use, 255) of lane id
#endif
    __shared__ unsigned key_count[WF_NKEYS], key_base[WF_NKEYS];
    __shared__ unsigned ray_count_s[3];      // rays_closest, rays_shadow, depth_overflow of this CTA and launch
    __shared__ unsigned n_live_s, n_pending_s, mq_head, mq_tail, run_mesh_s, done_s, p1_cursor, gq_count, gq_cursor, run_gather_s, n_parked_s, exit_s, gq_base_s;
    __shared__ WfClaim claim;
    __shared__ unsigned scan_part[32];

    Lane *const lanes = a.lanes + (size_t)blockIdx.x * P;
    Frame *const frames_all = a.frames + (size_t)blockIdx.x * P * (size_t)a.cfg.max_frames;
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const bool det = BHRT_FAST ? false : a.cfg.mode == BHRT_MODE_DET;

    // (split gather, resumed launch: the lanes' keys and the work ring come from the previous launch; every lane that was
    // parked has its answer now.  The first pass finds an empty live list, and the sort that follows rebuilds it.)
    unsigned char *const state = a.ext_gather ? a.cta_state + (size_t)blockIdx.x * wf_state_bytes(P) : (unsigned char *)0;
    const bool resumed = a.ext_gather && a.resume;
    for (unsigned i = tid; i < P; i += WF_BLOCK) { list[i] = (unsigned short)i; lkey[i] = resumed ? state[i] : (unsigned char)WF_KEY_DEAD; pend[i] = 0; }
#if WF_PREFETCH_FRAME
    for (unsigned i = tid; i < P; i += WF_BLOCK) lsp[i] = 0;
#endif
    if (tid < WF_NKEYS) key_count[tid] = 0;
    if (tid < 3) ray_count_s[tid] = 0;
    if (tid == 0) {
        n_live_s = resumed ? 0u : P; n_pending_s = 0; mq_head = 0; mq_tail = 0; p1_cursor = 0; run_mesh_s = 0; done_s = 0; gq_count = 0; gq_cursor = 0; n_parked_s = 0;
        exit_s = 0;
        if (resumed) claim = *(const WfClaim *)(state + wf_state_bytes(P) - 256u);
        else { claim.taken = 0; claim.claimed = 0; }
        wf_claim_topup(claim, a);
    }
    __syncthreads();        // (also publishes the barrier's initialisation to the threads that wait on it)
    stage_tables_end(s, a, smem, &tma_bar);

    // ray counters of the CTA in shared memory (one ATOMS per ray against five registers per thread)
    struct SharedCount { unsigned *p; __device__ __forceinline__ void operator++(int) { atomicAdd(p, 1u); } };
    struct { SharedCount closest, shadow, overflow; } rc = { { &ray_count_s[0] }, { &ray_count_s[1] }, { &ray_count_s[2] } };
#if BHRT_WIDE_P2
    // this OCTET's BVH traversal, kept across iterations (identical in its 8 lanes)
    OctTrav OT; OT.done = true; OT.cur = 0; OT.sp = 0;
    OctShared *const OS = (OctShared *)(smem + a.smem_oct) + (tid >> 3);
    const unsigned sub = lane & 7u, oshift = lane & 24u;
    uint32_t jmask = 0; int jextra = 0, jk = -1, jkind = Q_NONE;
    bool in_trav = false, jblocked = false, jfirst = false;
#else
    MeshJob J;                      // this thread's BVH traversal, kept across iterations: the job, the cold part, the stack ...
    // ... and the part every BVH step touches, in SHARED memory (44 bytes per thread, stride 11 words: conflict-free)
    // instead of local memory (a dozen reloads per step through the L1 that a thousand threads' lane copies and stacks
    // share).  Measured neutral on config 2: traversal phase -4 %, lanes pass +2 % for the 44 KB of L1.
    MeshTravHot &H = ((MeshTravHot *)(smem + a.smem_hot))[tid];
#endif
    bool have = false;
    unsigned jid = 0;
    bool first = !resumed;

    // phase clocks (BHRT_DEBUG's breakdown): kept by thread 0 in shared memory, not in eight registers of every thread
    __shared__ long long prof_s[5];         // lanes pass, sort, traversal, iterations, last mark
    if (tid == 0) { prof_s[0] = prof_s[1] = prof_s[2] = prof_s[3] = 0; }
    for (;;) {
        if (tid == 0) { prof_s[4] = clock64(); prof_s[3]++; }
        // ================= P1: one pass over the live lanes =================
        // Warps take 32-lane chunks of the sorted list in order (shared cursor): the CTA walks the
        // stages together and no warp waits at the barrier for a slower one's fixed share.
        const unsigned n_live = n_live_s;
        for (;;) {
            unsigned chunk = 0;
            if (lane == 0) chunk = atomicAdd(&p1_cursor, 32u);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            if (chunk >= n_live) break;
            const unsigned idx = chunk + lane;
            const bool valid = idx < n_live;
            const unsigned id = valid ? (unsigned)list[idx] : 0u;
            Lane L;
            // the recursion levels are used in place (global memory, cached in L1 while this lane is
            // being stepped): a step touches the top frame and, when it recurses, the slot above it
            Frame *const frames = frames_all + id;      // level l of this lane: frames[l * P]
            const bool was_idle = valid && !first && lkey[id] == WF_KEY_IDLE;
            bool want_new = valid && (first || was_idle);
            bool running = valid && !first && !was_idle;
            bool idle = false;
#if WF_PREFETCH_FRAME
            if (running) {
                const unsigned d = lsp[id];
                if (d > 0u && d < 255u) asm volatile("prefetch.global.L1 [%0];" ::"l"(&frames[(size_t)(d - 1u) * P]));
            }
#endif
            if (running) {
                wf_load_lane(L, &lanes[id]);
                // the answer of the pending query: winner-only work of a closest hit (a lane that comes
                // back from a caustic gather has no new answer: its last query was consumed before)
                const bool from_gather = L.sp > 0 && lane_stage(L) == ST_GATHER_WAIT;
                if (L.q.kind == Q_CLOSEST && !from_gather) L.hit_ok = scene_trace_finish(s, L.q, L.hit);
            }
            // The estimator coroutine up to its next ray.  Lanes of a warp may be in different
            // stages (and reach a stage after different numbers of steps): each trip runs ONE
            // stage -- the earliest any lane is in -- for all lanes that are in it, so later
            // stages wait until the stragglers have caught up and then run together.
            // (a caustic gather asked for by a step parks the lane: the queries of the whole CTA are
            // evaluated after the pass, one per warp at a time -- see the gather phase below)
            bool to_gather = false;
#if WF_LOCAL_FRAMES
            Frame F, Fc;                    // the top level and the slot above it
            if (running && L.sp > 0) wf_load(F, &frames[(size_t)(L.sp - 1) * P]);
#endif
            for (;;) {
                const int st = running ? (L.sp == 0 ? -1 : lane_stage(L)) : 127;
                const int mn = __reduce_min_sync(0xffffffffu, st);
                if (mn == 127) break;
                if (st != mn) continue;
                V3 c;
#if WF_LOCAL_FRAMES
                const int sp0 = L.sp;
                const int r = sample_step(L, F, Fc, s, a.cfg, a.gather, L.sp == 0 ? lanes[id].cold.bgu : 0.f, L.sp == 0 ? lanes[id].cold.bgv : 0.f, rc, c);
                if (L.sp > sp0) {           // a level was pushed: the parent goes back to its slot, the child is the top now
                    if (sp0 > 0) wf_store(&frames[(size_t)(sp0 - 1) * P], F);
                    F = Fc;
                } else if (L.sp < sp0 && L.sp > 0) wf_load(F, &frames[(size_t)(L.sp - 1) * P]);     // popped: the parent is the top again
                if (r == STEP_GATHER || r == STEP_RAY) wf_store(&frames[(size_t)(L.sp - 1) * P], F);     // the lane leaves the pass with a level on top
#else
                const int r = sample_step(L, frames[(size_t)(L.sp > 0 ? L.sp - 1 : 0) * P], frames[(size_t)(L.sp < a.cfg.max_frames ? L.sp : a.cfg.max_frames - 1) * P],
                                          s, a.cfg, a.gather, L.sp == 0 ? lanes[id].cold.bgu : 0.f, L.sp == 0 ? lanes[id].cold.bgv : 0.f, rc, c);
#endif
                if (r == STEP_GATHER) { to_gather = true; running = false; continue; }
                if (r == STEP_RAY) { running = false; continue; }
                if (r == STEP_DONE) {
                    const LaneCold &lc = lanes[id].cold;
                    a.samples[lc.work] = make_float4(c.x, c.y, c.z, 1.f);
                    if (lc.sidx == 0) { const size_t p = (size_t)lc.pj * s.width + lc.pi; a.hit_id[p] = L.hit_id; a.zbuf[p] = L.hit_z; }
                    want_new = true;
                    running = false;
                }
            }
            // ---- regeneration: warp-aggregated fetch of new camera samples
            bool dead = false;
            for (;;) {
                const unsigned m = __ballot_sync(0xffffffffu, want_new);
                if (!m) break;
                const int leader = __ffs(m) - 1;
                unsigned long long base = 0;
                // the CTA's claimed items are handed out without ever running past what is claimed (an item index
                // beyond it could not be given back); a lane that gets none stays idle until the next pass
                unsigned got = 0;
                if ((int)lane == leader) {
                    const unsigned want = (unsigned)__popc(m);
                    for (;;) {
                        const unsigned long long old = *(volatile unsigned long long *)&claim.taken;
                        const unsigned long long avail = claim.claimed - old;
                        got = avail < (unsigned long long)want ? (unsigned)avail : want;
                        if (got == 0u) break;
                        if (atomicCAS(&claim.taken, old, old + got) == old) { base = old; break; }
                    }
                }
                base = __shfl_sync(0xffffffffu, base, leader);
                got = __shfl_sync(0xffffffffu, got, leader);
                if (want_new) {
                    const unsigned rnk = (unsigned)__popc(m & ((1u << lane) - 1u));
                    const unsigned long long g = rnk < got ? wf_claim_item(claim, a, base + rnk) : 0ull;
                    if (rnk >= got) { idle = true; want_new = false; }
                    else if (g >= a.n_work) { dead = true; want_new = false; }
                    else {
                        int pi, pj, sidx;
                        if (work_to_pixel(s, a, g, pi, pj, sidx)) {
                            LaneCold lc;
                            lc.work = g; lc.pi = pi; lc.pj = pj; lc.sidx = sidx;
                            lc.bgu = (float)pi / s.width; lc.bgv = (float)pj / s.height; lc.pad_ = 0;
                            lanes[id].cold = lc;
                            L.rng.seed(a.seed, (uint64_t)pj * s.width + pi, (uint64_t)sidx);
                            Ray pr = primary_ray(s, pi, pj, !det, L.rng);
                            sample_begin(L, pr, rc);
                            want_new = false;
                        }
                    }
                }
            }
            // ---- the new ray against the scene graph (every surviving lane has one)
            const bool alive = valid && !dead && !idle;
            bool to_mesh = false;
            if (alive && !to_gather) {
                Hit h;
                uint32_t mask;
#if BHRT_WIDE_P2
                Ray mray;
                const bool blocked = scene_trace_graph(s, L.q, h, mask, &mray);
                if (!blocked && mask) {
                    const uint4 m0 = make_uint4(f2u(mray.p.x), f2u(mray.p.y), f2u(mray.p.z), f2u(mray.dir.x));
                    const uint4 m1 = make_uint4(f2u(mray.dir.y), f2u(mray.dir.z), (uint32_t)(L.q.kind | (L.q.hitSide << 8)), f2u(L.q.t_max));
                    uint4 *dst = reinterpret_cast<uint4 *>(&lanes[id].mesh_ray);
                    dst[0] = m0; dst[1] = m1;
                }
#else
                const bool blocked = scene_trace_graph(s, L.q, h, mask);
#endif
                L.hit_ok = blocked;
                L.hit.z = h.z; L.hit.node = h.node; L.hit.front = h.front;
                L.mesh_mask = 0;
                if (!blocked && (mask || s.n_mesh_nodes > BHRT_MESH_MASK_BITS)) { L.mesh_mask = mask; to_mesh = true; }
            }
            if (alive) {
                wf_store_lane(&lanes[id], L);
#if WF_PREFETCH_FRAME
                lsp[id] = (unsigned char)(L.sp < 255 ? L.sp : 255);
#endif
#if WF_LOCAL_FRAMES
                lkey[id] = (unsigned char)wf_key(s, L, F);          // (F is the top level whenever L.sp > 0)
#else
                lkey[id] = (unsigned char)(wf_key(s, L, frames[(size_t)(L.sp > 0 ? L.sp - 1 : 0) * P]) & a.key_mask);
#endif
            } else if (valid) lkey[id] = idle ? (unsigned char)WF_KEY_IDLE : (unsigned char)WF_KEY_DEAD;
            // park the lanes that wait for a caustic gather
            const unsigned gmask = __ballot_sync(0xffffffffu, to_gather);
            if (gmask) {
                const int leader = __ffs(gmask) - 1;
                unsigned base = 0;
                if ((int)lane == leader) { base = atomicAdd(&gq_count, (unsigned)__popc(gmask)); atomicAdd(&n_parked_s, (unsigned)__popc(gmask)); }
                base = __shfl_sync(0xffffffffu, base, leader);
                if (to_gather) { gq[base + __popc(gmask & ((1u << lane) - 1u))] = (unsigned short)id; pend[id] = 1; }
            }
            // queue the rays that entered a mesh
            const unsigned mm = __ballot_sync(0xffffffffu, to_mesh);
            if (mm) {
                const int leader = __ffs(mm) - 1;
                unsigned base = 0;
                if ((int)lane == leader) { base = atomicAdd(&mq_tail, (unsigned)__popc(mm)); atomicAdd(&n_pending_s, (unsigned)__popc(mm)); }
                base = __shfl_sync(0xffffffffu, base, leader);
                if (to_mesh) { mq[(base + __popc(mm & ((1u << lane) - 1u))) % P] = (unsigned short)id; pend[id] = 1; }
            }
        }
        first = false;
        __syncthreads();
        // ================= caustic gathers of the parked lanes: one query per warp at a time =================
        // (evaluated here rather than inside the pass: the queries of a pass sit in a few chunks, and the
        // warps that own those chunks would keep all the others waiting at the barrier)
        if (a.gather.use && !a.ext_gather) {
            // (batched like the BVH traversal: a phase with few queries cannot keep 32 warps evenly busy)
            if (tid == 0) run_gather_s = gq_count > 0 && (gq_count >= a.gather_batch || n_live <= P / 4);
            __syncthreads();
        }
        if (a.gather.use && !a.ext_gather && run_gather_s) {
            const unsigned ng = gq_count;
            // ---- order the parked queries by photon cell (counting sort on a 12-bit key: the low 4 bits of each coarse
            // cell coordinate; putting the queries of the most populated cells first was measured too: no gain).  The gather is bound by the L2 -> SM bandwidth of the records it streams (77 KB per
            // query on BASELINE config 5, 32 warps at once); with the queries of one cell next to each other the 32
            // warps stream the SAME block at the same time and all but the first find it in L1.  (`list` and the
            // sort's group counters are free between the lanes pass and the next sort.)
            unsigned short *const gq2 = gq + P;
            {
                unsigned *const cnt = (unsigned *)gcount;               // [4096]
                unsigned short *const k16 = list;                       // key of entry i
                for (unsigned i = tid; i < 4096u; i += WF_BLOCK) cnt[i] = 0u;
                __syncthreads();
                for (unsigned i = tid; i < ng; i += WF_BLOCK) {
                    const unsigned gid = gq[i];
                    const V3 qp = frames_all[(size_t)(lanes[gid].sp - 1) * P + gid].p;
                    const int cx = (int)floorf((qp.x - a.gather.origin[0]) / a.gather.cell), cy = (int)floorf((qp.y - a.gather.origin[1]) / a.gather.cell),
                              cz = (int)floorf((qp.z - a.gather.origin[2]) / a.gather.cell);
                    const unsigned key = ((unsigned)cx & 15u) | (((unsigned)cy & 15u) << 4) | (((unsigned)cz & 15u) << 8);
                    k16[i] = (unsigned short)key;
                    atomicAdd(&cnt[key], 1u);
                }
                __syncthreads();
                // exclusive scan of the 4096 counters: each thread owns WF_SCAN_PER consecutive ones
                const unsigned per = (4096u + WF_BLOCK - 1u) / WF_BLOCK, c0 = tid * per;
                unsigned mine = 0;
                for (unsigned j = 0; j < per; j++) if (c0 + j < 4096u) mine += cnt[c0 + j];
                unsigned incl = mine;
                for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
                if (lane == 31u) scan_part[tid >> 5] = incl;
                __syncthreads();
                if (tid < 32u) {
                    const unsigned v = tid < WF_BLOCK / 32 ? scan_part[tid] : 0u;
                    unsigned inc2 = v;
                    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc2, o); if ((int)lane >= o) inc2 += t; }
                    scan_part[tid] = inc2 - v;
                }
                __syncthreads();
                unsigned run = scan_part[tid >> 5] + incl - mine;
                for (unsigned j = 0; j < per; j++) if (c0 + j < 4096u) { const unsigned t = cnt[c0 + j]; cnt[c0 + j] = run; run += t; }
                __syncthreads();
                for (unsigned i = tid; i < ng; i += WF_BLOCK) gq2[atomicAdd(&cnt[k16[i]], 1u)] = gq[i];
                __syncthreads();
            }
            for (;;) {
                unsigned i = 0;
                if (lane == 0) i = atomicAdd(&gq_cursor, 1u);
                i = __shfl_sync(0xffffffffu, i, 0);
                if (i >= ng) break;
                const unsigned gid = gq2[i];
                Lane &Lg = lanes[gid];
                const Frame &fq = frames_all[(size_t)(Lg.sp - 1) * P + gid];
                V3 irr, dirv;
                estimate_irradiance_warp(a.gather, fq.p, fq.vN, irr, dirv, pm_scratch + (tid >> 5), a.cnt->pmprof);
                if (lane == 0) { Lg.ret = irr; Lg.sw = dirv; pend[gid] = 0; atomicSub(&n_parked_s, 1u); }
            }
            __syncthreads();
            if (tid == 0) { gq_count = 0; gq_cursor = 0; }
        }
        if (tid == 0) { const long long c = clock64(); prof_s[0] += c - prof_s[4]; prof_s[4] = c; }

        // Sort first: the lanes that did not go to a mesh are ready for the next pass.  The traversal
        // phase runs only when it has enough rays to keep every thread busy (or the live list has run
        // low); the lanes it completes are merged by a second sort.
        for (int round = 0; round < 2; round++) {
            // ================= P3: counting sort of the live lanes by key =================
            // No atomics: per 32-lane group, __match_any_sync gives each key's count;
            // one thread per key scans its column over the groups; a lane's slot is then
            // key base + its group's offset + its rank among the group's lanes of that key.
            const unsigned n_groups = (P + 31u) / 32u;
            for (unsigned g = tid >> 5; g < n_groups; g += WF_BLOCK / 32) {
                const unsigned i = g * 32u + lane;
                const unsigned k = (i < P && !pend[i]) ? (unsigned)lkey[i] : (unsigned)WF_KEY_DEAD;
                unsigned short *row = gcount + g * WF_NKEYS;
                for (unsigned j = lane; j < WF_NKEYS; j += 32u) row[j] = 0;
                __syncwarp();
                const unsigned peers = __match_any_sync(0xffffffffu, k);
                if (k != WF_KEY_DEAD && (int)lane == __ffs(peers) - 1) row[k] = (unsigned short)__popc(peers);
            }
            __syncthreads();
            if (tid < WF_NKEYS) {
                unsigned run = 0;
                for (unsigned g = 0; g < n_groups; g++) { const unsigned t = gcount[g * WF_NKEYS + tid]; gcount[g * WF_NKEYS + tid] = (unsigned short)run; run += t; }
                key_count[tid] = run;
            }
            __syncthreads();
            if (tid == 0) {
                unsigned acc = 0;
                for (int k = 0; k < WF_NKEYS; k++) { key_base[k] = acc; acc += key_count[k]; }
                n_live_s = acc;
                p1_cursor = 0;
                wf_claim_topup(claim, a);
                // decisions every thread must agree on, taken while nobody changes their inputs
                run_mesh_s = round == 0 && n_pending_s > 0 && (mq_tail - mq_head >= a.mesh_batch || acc < WF_BLOCK);
                done_s = acc == 0 && n_pending_s == 0 && n_parked_s == 0;
                // split gather: nothing left to do but the parked queries -> hand them over and leave
                // (not waiting for the last few live lanes -- a pass over a handful of lanes costs a sort and four barriers like any
                // other: they keep their keys and carry on after the gather, with everybody else)
                exit_s = a.ext_gather && acc <= a.exit_live && n_pending_s == 0 && n_parked_s > 0;
                if (exit_s) gq_base_s = atomicAdd(a.gq_total, gq_count);
            }
            __syncthreads();
            for (unsigned g = tid >> 5; g < n_groups; g += WF_BLOCK / 32) {
                const unsigned i = g * 32u + lane;
                const unsigned k = (i < P && !pend[i]) ? (unsigned)lkey[i] : (unsigned)WF_KEY_DEAD;
                const unsigned peers = __match_any_sync(0xffffffffu, k);
                if (k != WF_KEY_DEAD) list[key_base[k] + gcount[g * WF_NKEYS + k] + __popc(peers & ((1u << lane) - 1u))] = (unsigned short)i;
            }
            __syncthreads();
            if (tid == 0) { const long long c = clock64(); prof_s[1] += c - prof_s[4]; prof_s[4] = c; }
            if (!run_mesh_s) break;

            // ================= P2: BVH traversal of the queued rays =================
            // per-lane refill while the queue lasts, then at most `mesh_budget` more trips: what
            // is still running after that stays in its thread until the next traversal phase
            const unsigned tail = mq_tail;
            int budget = n_live_s == 0 ? 64 : a.mesh_budget;    // nothing else to do: stay in the traversal
#if BHRT_WIDE_P2
            // ---- eight lanes per ray over the wide tree (wide.cuh).  A job is owned by an OCTET: everything below
            // is octet-uniform except the per-child / per-triangle work inside oct_internal / oct_leaf.
            for (;;) {
                const bool need = !have;
                const unsigned nm = __ballot_sync(0xffffffffu, need && sub == 0u);
                unsigned head = 0;
                if (lane == 0) head = *(volatile unsigned *)&mq_head;
                head = __shfl_sync(0xffffffffu, head, 0);
                const bool drained = head >= tail;
                if (nm && !drained) {
                    const int leader = __ffs(nm) - 1;
                    unsigned base = 0;
                    if ((int)lane == leader) base = atomicAdd(&mq_head, (unsigned)__popc(nm));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (need) {
                        const unsigned i = base + __popc(nm & ((1u << oshift) - 1u));
                        if (i < tail) {
                            jid = mq[i % P];
                            const Lane &Lg = lanes[jid];
                            // mesh_job_begin (core.cuh)
                            OS->hz = Lg.hit.z; OS->hnode = Lg.hit.node; OS->hfront = Lg.hit.front ? 1 : 0;
                            OS->hface = 0; OS->ha0 = OS->ha1 = OS->ha2 = 0.f;
                            jmask = Lg.mesh_mask;
                            jextra = s.n_mesh_nodes > BHRT_MESH_MASK_BITS ? s.mesh_nodes[BHRT_MESH_MASK_BITS - 1] + 1 : s.n_nodes;
                            have = true; in_trav = false; jblocked = false; jfirst = jmask != 0u;
                        }
                    }
                }
                if (!__any_sync(0xffffffffu, have)) break;
                // job control: the next mesh node of the ray (mesh_job_next), or none left: finished
                bool finished = false;
                if (have && !in_trav) {
                    int k = -1;
                    if (jmask) { k = s.mesh_nodes[__ffs(jmask) - 1]; jmask &= jmask - 1u; }
                    else {
                        while (jextra < s.n_nodes && s.nodes[jextra].kind != BHRT_OBJ_MESH) jextra++;
                        if (jextra < s.n_nodes) k = jextra++;
                    }
                    if (k < 0) finished = true;
                    else {
                        jk = k;
                        const MeshAccel &A = s.accels[s.nodes[k].mesh];
                        Ray r; int hitSide; float t_max;
                        if (jfirst) {
                            // the lanes pass left the ray in this mesh's space (and its root box is known to be entered)
                            const uint4 *src = reinterpret_cast<const uint4 *>(&lanes[jid].mesh_ray);
                            const uint4 m0 = src[0], m1 = src[1];
                            r.p = V3(u2f(m0.x), u2f(m0.y), u2f(m0.z)); r.dir = V3(u2f(m0.w), u2f(m1.x), u2f(m1.y));
                            jkind = (int)(m1.z & 255u); hitSide = (int)(m1.z >> 8); t_max = u2f(m1.w);
                            jfirst = false;
                        } else {
                            Query q;
                            wf_load(q, &lanes[jid].q);
                            jkind = q.kind; hitSide = q.hitSide; t_max = q.t_max;
                            r = node_local_ray(s, k, q.ray);
                        }
                        if (jkind == Q_SHADOW) oct_begin(OT, s, A, r, true, t_max, BHRT_BIGFLOAT, true, hitSide);
                        else {
                            // an equal-t candidate of an EARLIER node loses against the current winner
                            const float hz = OS->hz;
                            oct_begin(OT, s, A, r, false, 0.f, OS->hnode > k ? prev_float(hz) : hz, OS->hfront != 0, hitSide);
                        }
                        in_trav = true;
                    }
                }
                // internal nodes while most octets are at one, then the leaf groups of all octets that reached one
                for (int r = 0; r < a.trav_max; r++) {
                    const bool in_job = have && in_trav && !OT.done;
                    const bool go = in_job && !BHRT_BVH_LEAF(OT.cur);
                    const unsigned gm = __ballot_sync(0xffffffffu, go), jm = __ballot_sync(0xffffffffu, in_job);
                    if (!gm || (r >= a.trav_min && 100 * __popc(gm) <= a.trav_pct * __popc(jm))) break;
                    oct_internal(OT, OS, go, sub, oshift);
                }
                {
                    const bool at_leaf = have && in_trav && !OT.done && BHRT_BVH_LEAF(OT.cur);
                    if (__any_sync(0xffffffffu, at_leaf)) oct_leaf(OT, OS, at_leaf, sub, oshift);
                }
                // mesh_job_harvest
                if (have && in_trav && OT.done) {
                    in_trav = false;
                    if (OT.result) {
                        if (jkind == Q_SHADOW) { jblocked = true; finished = true; }
                        else { OS->hz = OT.z; OS->hfront = OT.front ? 1 : 0; OS->hnode = jk; OS->hface = OT.face; OS->ha0 = OT.a0; OS->ha1 = OT.a1; OS->ha2 = OT.a2; }
                    }
                }
                if (finished) {
                    if (sub == 0u) {
                        Lane &Lg = lanes[jid];
                        if (jblocked) Lg.hit_ok = true;
                        else if (jkind == Q_CLOSEST) {
                            const int hn = OS->hnode;
                            Lg.hit.z = OS->hz; Lg.hit.node = hn; Lg.hit.front = OS->hfront != 0;
                            Lg.hit.uvw = V3(OS->ha0, OS->ha1, OS->ha2); Lg.hit.duvw0.x = u2f(OS->hface);
                            lkey[jid] = (unsigned char)(wf_key_patch(s, lkey[jid], hn) & a.key_mask);
                        }
                        pend[jid] = 0;
                        atomicSub(&n_pending_s, 1u);
                    }
                    have = false;
                }
                if (drained && --budget <= 0) break;
            }
#else
            // (MEASURED, round 2: a register copy of the traversal's scalars for the length of the phase -- MeshTravHot,
            // which exists so that it can be tried -- costs more in spills elsewhere than the per-step reloads it
            // removes: 1437 against 1500 Mrays/s on config 2.  The 64-register budget of a 1024-thread CTA is the wall.)
            for (;;) {
                const bool need = !have;
                const unsigned nm = __ballot_sync(0xffffffffu, need);
                unsigned head = 0;
                if (lane == 0) head = *(volatile unsigned *)&mq_head;
                head = __shfl_sync(0xffffffffu, head, 0);
                const bool drained = head >= tail;
                if (nm && !drained) {
                    const int leader = __ffs(nm) - 1;
                    unsigned base = 0;
                    if ((int)lane == leader) base = atomicAdd(&mq_head, (unsigned)__popc(nm));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (need) {
                        const unsigned i = base + __popc(nm & ((1u << lane) - 1u));
                        if (i < tail) {
                            jid = mq[i % P];
                            const Lane &Lg = lanes[jid];
                            Query q;
                            wf_load(q, &Lg.q);
                            mesh_job_begin(J, s, q, Lg.hit.z, Lg.hit.node, Lg.hit.front, Lg.mesh_mask);
                            have = true;
                        }
                    }
                }
                if (!__any_sync(0xffffffffu, have)) break;
                // job control: next mesh node of the ray (or none left: finished)
                bool finished = false;
                if (have && !J.in_trav) finished = !mesh_job_next(J, H, s);
                // internal nodes: lanes that reach a leaf wait here for the others (a few trips),
                // then all leaves are processed together -- the triangle tests run converged
                for (int r = 0; r < a.trav_max; r++) {
                    const bool in_job = have && !finished && !H.done;
                    const bool go = in_job && !BHRT_BVH_LEAF(H.cur);
                    const unsigned gm = __ballot_sync(0xffffffffu, go), jm = __ballot_sync(0xffffffffu, in_job);
                    // stop when no lane is at an internal node, or (after a few steps) when most are waiting at a leaf
                    if (!gm || (r >= a.trav_min && 100 * __popc(gm) <= a.trav_pct * __popc(jm))) break;
                    if (go) mesh_trav_internal(H, J.stk, s.accel_data);
                }
                if (have && !finished && !H.done && BHRT_BVH_LEAF(H.cur)) mesh_trav_leaf(H, J.C, J.stk, s.accel_data);
                if (have && !finished && H.done) finished = mesh_job_harvest(J, H);
                if (finished) {
                    Lane &Lg = lanes[jid];
                    if (J.blocked) Lg.hit_ok = true;
                    else if (J.q.kind == Q_CLOSEST) {
                        Lg.hit.z = J.h.z; Lg.hit.node = J.h.node; Lg.hit.front = J.h.front;
                        Lg.hit.uvw = J.h.uvw; Lg.hit.duvw0.x = J.h.duvw0.x;
                        lkey[jid] = (unsigned char)(wf_key_patch(s, lkey[jid], J.h.node) & a.key_mask);
                    }
                    pend[jid] = 0;
                    atomicSub(&n_pending_s, 1u);
                    have = false;
                }
                if (drained && --budget <= 0) break;
            }
#endif
            __syncthreads();
            if (tid == 0 && mq_head > mq_tail) mq_head = mq_tail;       // over-claimed pops of an empty queue
            if (tid == 0) { const long long c = clock64(); prof_s[2] += c - prof_s[4]; prof_s[4] = c; }
        }
        if (done_s || exit_s) break;
    }
    if (a.ext_gather) {
        // what the next launch needs: the lanes' keys (dead lanes stay dead, parked ones come back with the key they were
        // parked with) and the work ring; the parked queries go to the device-wide list
        for (unsigned i = tid; i < P; i += WF_BLOCK) state[i] = lkey[i];
        if (tid == 0) *(WfClaim *)(state + wf_state_bytes(P) - 256u) = claim;
        if (exit_s) for (unsigned i = tid; i < gq_count; i += WF_BLOCK) a.gq_list[gq_base_s + i] = blockIdx.x * P + (unsigned)gq[i];
    }
    __syncthreads();
    if (tid == 0) {
        if (ray_count_s[0]) atomicAdd(&a.cnt->rays_closest, (unsigned long long)ray_count_s[0]);
        if (ray_count_s[1]) atomicAdd(&a.cnt->rays_shadow, (unsigned long long)ray_count_s[1]);
        if (ray_count_s[2]) atomicAdd(&a.cnt->depth_overflow, (unsigned long long)ray_count_s[2]);
    }
    if (tid == 0) {
        atomicAdd(&a.cnt->prof[0], (unsigned long long)prof_s[0]); atomicAdd(&a.cnt->prof[1], (unsigned long long)prof_s[1]);
        atomicAdd(&a.cnt->prof[2], (unsigned long long)prof_s[2]); atomicAdd(&a.cnt->prof[3], (unsigned long long)prof_s[3]);
    }
}

// ---- split gather: the parked queries of ALL CTAs, sorted by photon cell --------------------------
// pm_query_keys: sort key of every parked query = its coarse photon cell (x fastest: a level-0 search reads ROWS of
// five x-adjacent cells, so queries that are neighbours in the sorted order stream nearly the same records).
__device__ __forceinline__ const Frame &wf_query_frame(const WaveArgs &a, unsigned entry)
{
    const unsigned cta = entry / a.pool, id = entry % a.pool;
    return a.frames[(size_t)cta * a.pool * (size_t)a.cfg.max_frames + (size_t)(a.lanes[entry].sp - 1) * a.pool + id];
}
__global__ void pm_query_keys(WaveArgs a, unsigned n, int cell_bits, uint32_t *__restrict__ keys)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 qp = wf_query_frame(a, a.gq_list[i]).p;
    const float q3[3] = { qp.x, qp.y, qp.z };
    const PhotonGather &pg = a.gather;
    uint32_t key = 0, mul = 1;
    int c[3];
    for (int ax = 0; ax < 3; ax++) {
        // (cells up to two beyond the grid still reach photons; everything farther away shares the border key and ends at once)
        float f = floorf((q3[ax] - pg.origin[ax]) / pg.cell);
        const float hi = (float)(pg.dim[ax] + 1);
        f = f < -2.f ? -2.f : (f > hi ? hi : f);
        c[ax] = (int)f;
        key += mul * (uint32_t)(c[ax] + 2);
        mul *= (uint32_t)(pg.dim[ax] + 4);
    }
    // Cost class of the query = log2 of the photons in its level-0 block (what a query costs follows it: a sparse block is
    // summed once, a dense one is scanned several times at a finer level).  The EXPENSIVE classes come first in the sorted
    // order: the queries of a caustic's focus cost a hundred times the average, and handed out last a few CTAs would still
    // be working on them long after the rest of the device has run out of work (measured: SMs busy 27 % of the kernel's time).
    if (cell_bits + 3 <= 32) {
        uint32_t visit = 0;
        int x0 = c[0] - 2, x1 = c[0] + 2;
        if (x0 < 0) x0 = 0;
        if (x1 > pg.dim[0] - 1) x1 = pg.dim[0] - 1;
        if (x0 <= x1)
            for (int dz = -2; dz <= 2; dz++)
                for (int dy = -2; dy <= 2; dy++) {
                    const int y = c[1] + dy, z = c[2] + dz;
                    if (y < 0 || z < 0 || y >= pg.dim[1] || z >= pg.dim[2]) continue;
                    const uint32_t row = (uint32_t)pg.dim[0] * ((uint32_t)y + (uint32_t)pg.dim[1] * (uint32_t)z);
                    visit += pg.coarse_start[row + (uint32_t)x1 + 1u] - pg.coarse_start[row + (uint32_t)x0];
                }
        int cls = 23 - __clz(visit | 1u);       // < 512 photons: 0 ... >= 32768: 7
        cls = cls < 0 ? 0 : (cls > 7 ? 7 : cls);
        key |= (uint32_t)(7 - cls) << cell_bits;
    }
    keys[i] = key;
}
// pm_gather_queries: one query per warp at a time, taken IN SORTED ORDER -- a CTA claims PMQ_CLAIM consecutive queries (two per warp: small claims keep the device balanced) and
// its warps walk them together, so the warps of an SM stream the same photon rows at about the same time: the first one
// brings a line in, the others find it in L1 (hit rate of the in-kernel phase, where three quarters of a CTA's queries
// are GI hits scattered over the scene: 25 %).  The per-query routine is the one the in-kernel phase runs: same result.
#define PMQ_BLOCK 512
#define PMQ_CLAIM 32
#define PMQ_DEFER 16384      // default of WaveArgs::gather_defer
__global__ void __launch_bounds__(PMQ_BLOCK, 2)
pm_gather_queries(WaveArgs a, const unsigned *__restrict__ sorted, unsigned n, unsigned *cursor, unsigned *__restrict__ heavy, unsigned *n_heavy)
{
    __shared__ PmWarpScratch scratch[PMQ_BLOCK / 32];
    __shared__ unsigned base_s, next_s;
    const unsigned lane = threadIdx.x & 31u;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { base_s = atomicAdd(cursor, (unsigned)PMQ_CLAIM); next_s = 0; }
        __syncthreads();
        const unsigned base = base_s;
        if (base >= n) break;
        const unsigned cnt = n - base < (unsigned)PMQ_CLAIM ? n - base : (unsigned)PMQ_CLAIM;
        for (;;) {
            unsigned i = 0;
            if (lane == 0) i = atomicAdd(&next_s, 1u);
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i >= cnt) break;
            const unsigned entry = sorted[base + i];
            const Frame &fq = wf_query_frame(a, entry);
            V3 irr, dirv;
            const long long t0 = clock64();
            const bool deferred = estimate_irradiance_warp_inl<false>(a.gather, fq.p, fq.vN, irr, dirv, scratch + (threadIdx.x >> 5), a.cnt->pmprof, a.gather_defer);
            if (deferred) { if (lane == 0) heavy[atomicAdd(n_heavy, 1u)] = entry; }
            else if (lane == 0) {
                Lane &Lg = a.lanes[entry]; Lg.ret = irr; Lg.sw = dirv;
                if (a.gather_prof) {        // BHRT_DEBUG: the cost distribution of the queries (the slowest one bounds the kernel)
                    const unsigned long long dt = (unsigned long long)(clock64() - t0);
                    atomicMax(&a.cnt->pmprof[12], dt); atomicAdd(&a.cnt->pmprof[14], dt); atomicAdd(&a.cnt->pmprof[15], 1ull);
                    if (dt > 1000000ull) atomicAdd(&a.cnt->pmprof[13], 1ull);
                }
            }
        }
    }
}

// the deferred queries: one per CTA at a time
__global__ void __launch_bounds__(PMQ_BLOCK, 2)
pm_gather_heavy(WaveArgs a, const unsigned *__restrict__ heavy, const unsigned *__restrict__ n_heavy, unsigned *cursor)
{
    __shared__ PmWarpScratch scratch[PMQ_BLOCK / 32];
    __shared__ float red[PMQ_BLOCK / 32][16];
    __shared__ unsigned next_s;
    const unsigned n = *n_heavy;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) next_s = atomicAdd(cursor, 1u);
        __syncthreads();
        const unsigned i = next_s;
        if (i >= n) break;
        const unsigned entry = heavy[i];
        const Frame &fq = wf_query_frame(a, entry);
        V3 irr, dirv;
        estimate_irradiance_cta(a.gather, fq.p, fq.vN, irr, dirv, scratch, red);
        if (threadIdx.x == 0) { Lane &Lg = a.lanes[entry]; Lg.ret = irr; Lg.sw = dirv; }
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

}  // namespace BHRT_NS
