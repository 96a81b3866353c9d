// render.cu -- libbhrt.so: the sm_100a render kernels and the C ABI of include/bhrt.h.
//
// Replaces BeginRender() (reference Main.cpp:178-242): the OpenMP loop over image columns becomes
// ONE launch of the persistent kernel wf_render (wavefront.cuh: every CTA runs its own wavefront
// over a pool of camera samples; exact build here, fast-math build in render_fast.cu), wf_reduce
// adds the samples of a pixel in sample order, and resolve_kernel does sum/spp -> gamma -> Color24.
// The photon kernels live in photon.cuh.
//
// There is NO CPU fallback in this library: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <string>
#include <thread>
#include <vector>
#include "core.cuh"
#include "setup.hpp"
#include "accel.hpp"
#include "photon.cuh"
#include "wavefront.cuh"
#include "wide.cuh"

using namespace bhrt;

// render_fast.cu: the same render kernel built with --use_fast_math (stochastic mode)
extern "C" int bhrt_fast_prepare(int smem_bytes, int *blocks_per_sm);
extern "C" int bhrt_fast_launch(const void *view, size_t view_bytes, const void *args, size_t args_bytes, unsigned grid, int smem_bytes, cudaStream_t stream);
// the same with the round-1 traversal phase (one ray per thread, binary tree)
extern "C" int bhrt_fast_gather_keys(const void *args, size_t args_bytes, unsigned n, int cell_bits, uint32_t *keys, cudaStream_t stream);
extern "C" int bhrt_fast_gather_queries(const void *args, size_t args_bytes, const unsigned *sorted, unsigned n, unsigned *small, unsigned *heavy, unsigned grid, cudaStream_t stream);
extern "C" int bhrt_fastb_gather_keys(const void *args, size_t args_bytes, unsigned n, int cell_bits, uint32_t *keys, cudaStream_t stream);
extern "C" int bhrt_fastb_gather_queries(const void *args, size_t args_bytes, const unsigned *sorted, unsigned n, unsigned *small, unsigned *heavy, unsigned grid, cudaStream_t stream);
extern "C" int bhrt_fastb_prepare(int smem_bytes, int *blocks_per_sm);
extern "C" int bhrt_fastb_launch(const void *view, size_t view_bytes, const void *args, size_t args_bytes, unsigned grid, int smem_bytes, cudaStream_t stream);

static thread_local std::string g_err;
static int set_err(int code, const std::string &m) { g_err = m; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return set_err(BHRT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct DevBuf {
    void *p = nullptr; size_t bytes = 0;
    int ensure(size_t n)
    {
        if (n <= bytes && p) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        if (n == 0) n = 16;
        if (cudaMalloc(&p, n) != cudaSuccess) return -1;
        bytes = n;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct bhrt_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    unsigned long long launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool have_scene = false;
    SceneView view;
    int width = 0, height = 0;
    DevBuf nodes, materials, lights, texmaps, textures, texels, meshes, mesh_data, tex_offsets;
    DevBuf accum, hit_id, z, rgb8, linear, counters;
    PhotonState photons;
    std::vector<bhrt_light> sorted_lights;      // host copy of the device light table (photon emission setup)
    int rendered_spp = 0;
    int n_materials = 0;
    DevBuf accels, accel_data;
    // the 8-wide regrouping of the BVHs is built on first use (it is two thirds of the host time of an upload and only the
    // BHRT_FLAG_WIDE_BVH traversal reads it): what ensure_wide() needs to build it later
    std::vector<bhrt_mesh> host_meshes; size_t host_mesh_bytes = 0; bool wide_built = false;
    // per kernel variant (0 exact, 1 fast + wide traversal, 2 fast + binary traversal): prepared for this much shared memory
    int kernel_smem[3] = { -1, -1, -1 }, render_blocks[3] = { 1, 1, 1 };
    int traversal = 0;                          // env BHRT_TRAVERSAL: 1 = wide, 2 = binary, 0 = what the flags say
    unsigned pool_lanes = 3072;                 // lanes (camera samples in flight) per CTA, 2.3 KB of state each
    size_t last_lanes_total = 0;
    int max_frames = BHRT_MAX_FRAMES;   // depth of the estimator's stack; doubled when a frame overflowed it (launch_render)
    int mesh_budget = WF_MESH_BUDGET;
    unsigned mesh_batch = WF_BLOCK;
    unsigned gather_batch = 0;                  // 0 = 11/12 of the pool
    bool pool_lanes_set = false;
    int trav_min = 3, trav_max = 12, trav_pct = 50;
    DevBuf lanes, frames, samples, wave_counters;
    // split gather (wavefront.cuh): the parked queries of a round, their sort keys (double-buffered), per-CTA carry-over state
    DevBuf gq_list, gq_sorted, gq_keys, gq_keys_sorted, gq_small, cta_state, gq_cub;
    int gather_mode = 0;                        // env BHRT_GATHER: 1 = inside wf_render, 2 = split, 0 = default (split)
    // asynchronous frames (bhrt_render_async): worker thread, a second stream for progress / cancel, the cancel flag
    DevBuf cancel_flag;
    cudaStream_t aux_stream = nullptr;
    std::thread worker;
    std::atomic<int> async_state{0};            // 0 idle, 1 running, 2 finished (not yet waited for)
    std::atomic<int> cancel_requested{0};
    int async_status = BHRT_OK;
    std::string async_error;
    long long frame_items = 0, frame_pixels = 0; int frame_spp = 1;
};

// sum/spp (Main.cpp:170), gamma (Main.cpp:220-227), Color24 (Main.cpp:230)
__global__ void resolve_kernel(const float4 *__restrict__ accum, int n, float inv_is_div_by, float *__restrict__ linear, uint8_t *__restrict__ rgb8)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float4 a = accum[p];
    V3 out = V3(a.x, a.y, a.z) / inv_is_div_by;
    linear[3 * p] = out.x; linear[3 * p + 1] = out.y; linear[3 * p + 2] = out.z;
    V3 g = gamma_correct(out);
    rgb8[3 * p] = float_to_byte(g.x); rgb8[3 * p + 1] = float_to_byte(g.y); rgb8[3 * p + 2] = float_to_byte(g.z);
}

// ---- test hook: the two mesh traversals side by side on caller-supplied rays (bhrt_debug_trace_mesh) ----
// rays: 8 floats each -- p.xyz, dir.xyz (mesh-local), bound (closest: incoming z; shadow: t_max), kind (1 closest,
// 2 shadow).  out: 4 floats per ray -- result (0/1), z, face id, front.
__global__ void dbg_trace_binary(SceneView s, int mesh, int n, const float *__restrict__ rays, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *q = rays + 8 * (size_t)i;
    Ray r; r.p = V3(q[0], q[1], q[2]); r.dir = V3(q[3], q[4], q[5]);
    const bool shadow = q[7] == 2.f;
    MeshTrav T;
    mesh_trav_begin(T.hot, T.cold, s, s.accels[mesh], r, shadow, q[6], shadow ? BHRT_BIGFLOAT : q[6], true, shadow ? (int)HIT_FRONT : (int)HIT_FRONT_AND_BACK);
    while (!T.hot.done) mesh_trav_step(T, s.accel_data);
    float *o = out + 4 * (size_t)i;
    o[0] = T.hot.result ? 1.f : 0.f; o[1] = T.hot.z; o[2] = (float)T.cold.mh.face; o[3] = T.hot.front ? 1.f : 0.f;
}
__global__ void __launch_bounds__(256)
dbg_trace_wide(SceneView s, int mesh, int n, const float *__restrict__ rays, float *__restrict__ out)
{
    __shared__ OctShared shared[256 / 8];
    const unsigned lane = threadIdx.x & 31u, sub = lane & 7u, oshift = lane & 24u;
    const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 3);
    const bool valid = i < n;
    OctShared *S = shared + (threadIdx.x >> 3);
    OctTrav T; T.done = true; T.cur = 0; T.sp = 0; T.result = false; T.z = 0; T.face = 0; T.front = true;
    if (valid) {
        const float *q = rays + 8 * (size_t)i;
        Ray r; r.p = V3(q[0], q[1], q[2]); r.dir = V3(q[3], q[4], q[5]);
        const bool shadow = q[7] == 2.f;
        oct_begin(T, s, s.accels[mesh], r, shadow, q[6], shadow ? BHRT_BIGFLOAT : q[6], true, shadow ? (int)HIT_FRONT : (int)HIT_FRONT_AND_BACK);
    }
    oct_trace_mesh(T, S, valid, sub, oshift);
    if (valid && sub == 0u) {
        float *o = out + 4 * (size_t)i;
        o[0] = T.result ? 1.f : 0.f; o[1] = T.z; o[2] = (float)T.face; o[3] = T.front ? 1.f : 0.f;
    }
}

// RenderImage::GetSampleCount (reference scene.h:534,570): samples accumulated per pixel, clamped to the reference's uint8
__global__ void sample_count_kernel(const float4 *__restrict__ accum, int n, uint8_t *__restrict__ out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float w = accum[p].w;
    out[p] = (uint8_t)(w < 0.f ? 0.f : (w > 255.f ? 255.f : w));
}

// ------------------------------------------------------------------------------------------
extern "C" {

int bhrt_abi_version(void) { return BHRT_ABI_VERSION; }
const char *bhrt_last_error(void) { return g_err.c_str(); }

int bhrt_create(int device, bhrt_ctx **out)
{
    if (!out) return set_err(BHRT_ERR_ARG, "bhrt_create: null out");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_err(BHRT_ERR_CUDA, std::string("bhrt_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
    if (device < 0 || device >= n) return set_err(BHRT_ERR_ARG, "bhrt_create: bad device index");
    CU(cudaSetDevice(device));
    bhrt_ctx *c = new bhrt_ctx;
    c->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    if (c->cancel_flag.ensure(16)) { delete c; return set_err(BHRT_ERR_CUDA, "cudaMalloc(cancel flag)"); }
    CU(cudaMemset(c->cancel_flag.p, 0, 16));
    if (c->counters.ensure(8 * sizeof(unsigned long long))) { delete c; return set_err(BHRT_ERR_CUDA, "cudaMalloc(counters)"); }
    // tuning knobs
    if (const char *e = getenv("BHRT_POOL_LANES")) { long v = atol(e); if (v >= 32 && v <= 32768) { c->pool_lanes = (unsigned)v; c->pool_lanes_set = true; } }
    if (const char *e = getenv("BHRT_MESH_BUDGET")) { long v = atol(e); if (v >= 1) c->mesh_budget = (int)v; }
    if (const char *e = getenv("BHRT_TRAV")) { int x, y, z; if (sscanf(e, "%d,%d,%d", &x, &y, &z) == 3 && x >= 0 && y >= 1) { c->trav_min = x; c->trav_max = y; c->trav_pct = z; } }
    if (const char *e = getenv("BHRT_GATHER_BATCH")) { long v = atol(e); if (v >= 1) c->gather_batch = (unsigned)v; }
    if (const char *e = getenv("BHRT_MESH_BATCH")) { long v = atol(e); if (v >= 1) c->mesh_batch = (unsigned)v; }
    if (const char *e = getenv("BHRT_GATHER")) c->gather_mode = !strcmp(e, "inkernel") ? 1 : (!strcmp(e, "split") ? 2 : 0);
    if (const char *e = getenv("BHRT_TRAVERSAL")) c->traversal = !strcmp(e, "wide") ? 1 : (!strcmp(e, "binary") ? 2 : 0);
    *out = c;
    return BHRT_OK;
}

int bhrt_destroy(bhrt_ctx *c)
{
    if (!c) return BHRT_OK;
    if (c->worker.joinable()) { bhrt_cancel(c); c->worker.join(); }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->cancel_flag.release();
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    DevBuf *bufs[] = { &c->nodes, &c->materials, &c->lights, &c->texmaps, &c->textures, &c->texels, &c->meshes, &c->mesh_data,
                       &c->tex_offsets, &c->accels, &c->accel_data, &c->lanes, &c->frames, &c->samples, &c->wave_counters, &c->accum, &c->hit_id, &c->z, &c->rgb8, &c->linear, &c->counters,
                       &c->gq_list, &c->gq_sorted, &c->gq_keys, &c->gq_keys_sorted, &c->gq_small, &c->cta_state, &c->gq_cub };
    for (auto b : bufs) b->release();
    c->photons.release();
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return BHRT_OK;
}

static int upload(bhrt_ctx *c, DevBuf &b, const void *src, size_t bytes)
{
    if (b.ensure(bytes + 16)) return set_err(BHRT_ERR_CUDA, "cudaMalloc failed");    // (+16: the kernels' bulk copies move whole 16-byte units)
    if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return BHRT_OK;
}

int bhrt_upload_scene(bhrt_ctx *c, const bhrt_scene_desc *d)
{
    if (!c || !d) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: null argument");
    if (d->abi_version != BHRT_ABI_VERSION) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: abi_version mismatch");
    if (d->n_nodes < 0 || d->n_nodes > BHRT_MAX_NODES || d->n_materials > BHRT_MAX_MATERIALS || d->n_lights > BHRT_MAX_LIGHTS ||
        d->n_texmaps > BHRT_MAX_TEXMAPS || d->n_textures > BHRT_MAX_TEXTURES || d->n_meshes > BHRT_MAX_MESHES)
        return set_err(BHRT_ERR_LIMIT, "bhrt_upload_scene: table larger than BHRT_MAX_*");
    if (d->camera.width <= 0 || d->camera.height <= 0) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: bad camera resolution");
    int max_depth = 1;
    for (int i = 0; i < d->n_nodes; i++) {
        const bhrt_node &n = d->nodes[i];
        if (n.depth < 1 || n.depth > BHRT_MAX_DEPTH) return set_err(BHRT_ERR_LIMIT, "bhrt_upload_scene: node depth out of range");
        if (n.parent >= i) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: nodes are not in DFS pre-order");
        if (n.parent >= 0 && d->nodes[n.parent].depth != n.depth - 1) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: inconsistent node depth");
        if (n.parent < 0 && n.depth != 1) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: inconsistent node depth");
        if (n.kind != BHRT_OBJ_NONE && (n.material < 0 || n.material >= d->n_materials)) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: node material out of range");
        if (n.kind == BHRT_OBJ_MESH && (n.mesh < 0 || n.mesh >= d->n_meshes)) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: node mesh out of range");
        if (n.depth > max_depth) max_depth = n.depth;
    }
    for (int i = 0; i < d->n_meshes; i++) {
        const bhrt_mesh &m = d->meshes[i];
        if (m.nf <= 0 || m.n_bvh_nodes <= 0) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: empty mesh");
        if ((size_t)m.bvh_off + (size_t)(m.n_bvh_nodes + 1) * sizeof(bhrt_bvh_node) > d->mesh_bytes) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: mesh blob too small");
        if (m.n_groups < 0 || (m.n_groups > 0 && (size_t)((m.elems_off + 4 * (int64_t)m.nf + 15) & ~(int64_t)15) + 4 * (size_t)m.n_groups > d->mesh_bytes))
            return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: mesh blob too small for its usemtl group table");
    }
    CU(cudaSetDevice(c->device));
    std::vector<bhrt_light> lights(d->lights, d->lights + d->n_lights);
    float allI = 0;
    sort_lights(lights, allI);
    std::vector<float> offs = texture_filter_offsets();
    int r;
    if ((r = upload(c, c->nodes, d->nodes, sizeof(bhrt_node) * d->n_nodes))) return r;
    if ((r = upload(c, c->materials, d->materials, sizeof(bhrt_material) * d->n_materials))) return r;
    if ((r = upload(c, c->lights, lights.data(), sizeof(bhrt_light) * lights.size()))) return r;
    c->sorted_lights = lights;
    c->photons.count = 0; c->photons.built = false;     // a photon map belongs to the scene it was traced in
    if ((r = upload(c, c->texmaps, d->texmaps, sizeof(bhrt_texmap) * d->n_texmaps))) return r;
    // texels: RGB8 -> RGBX8 (one aligned 32-bit load per texel in the kernels).  Every file texture is expanded
    // from its OWN offset (any byte offset / padding between textures is fine in the descriptor); the device copy
    // of the texture table gets texel_offset = 3 * (first RGBX8 texel), which is what texture_sample() divides by 3.
    std::vector<bhrt_texture> dev_textures(d->textures, d->textures + d->n_textures);
    std::vector<uint8_t> rgbx;
    for (int i = 0; i < d->n_textures; i++) {
        bhrt_texture &t = dev_textures[i];
        if (t.kind != BHRT_TEX_FILE) continue;
        const size_t n = (size_t)(t.width > 0 ? t.width : 0) * (size_t)(t.height > 0 ? t.height : 0);
        if (t.texel_offset < 0 || (size_t)t.texel_offset + 3 * n > d->texel_bytes) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: texture texels outside the texel blob");
        const uint8_t *src = d->texels + t.texel_offset;
        const size_t first = rgbx.size() / 4;
        rgbx.resize(rgbx.size() + 4 * n);
        uint8_t *dst = rgbx.data() + 4 * first;
        for (size_t k = 0; k < n; k++) { dst[4 * k] = src[3 * k]; dst[4 * k + 1] = src[3 * k + 1]; dst[4 * k + 2] = src[3 * k + 2]; dst[4 * k + 3] = 0; }
        t.texel_offset = (int64_t)(3 * first);
    }
    if ((r = upload(c, c->textures, dev_textures.data(), sizeof(bhrt_texture) * dev_textures.size()))) return r;
    if ((r = upload(c, c->texels, rgbx.data(), rgbx.size()))) return r;
    if ((r = upload(c, c->meshes, d->meshes, sizeof(bhrt_mesh) * d->n_meshes))) return r;
    if ((r = upload(c, c->mesh_data, d->mesh_data, d->mesh_bytes))) return r;
    if ((r = upload(c, c->tex_offsets, offs.data(), sizeof(float) * offs.size()))) return r;
    AccelBlob accel;
    std::string aerr;
    const bool wide_now = c->traversal == 1;             // env BHRT_TRAVERSAL=wide: every frame will want it
    if (!build_accel(*d, accel, &aerr, wide_now)) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: " + aerr);
    c->host_meshes.assign(d->meshes, d->meshes + d->n_meshes); c->host_mesh_bytes = d->mesh_bytes; c->wide_built = wide_now;
    if ((r = upload(c, c->accels, accel.meshes.data(), sizeof(MeshAccel) * accel.meshes.size()))) return r;
    if ((r = upload(c, c->accel_data, accel.data.data(), accel.data.size()))) return r;
    CU(cudaStreamSynchronize(c->stream));      // the staging vectors above die with this scope

    FrameConstants fc = camera_setup(d->camera);
    SceneView &s = c->view;
    memset(&s, 0, sizeof(s));
    s.n_nodes = d->n_nodes; s.n_lights = d->n_lights; s.max_depth = max_depth; s.n_materials = d->n_materials;
    s.nodes = (const bhrt_node *)c->nodes.p;
    s.materials = (const bhrt_material *)c->materials.p;
    s.lights = (const bhrt_light *)c->lights.p;
    s.texmaps = (const bhrt_texmap *)c->texmaps.p;
    s.textures = (const bhrt_texture *)c->textures.p;
    s.texels = (const uint8_t *)c->texels.p;
    s.meshes = (const bhrt_mesh *)c->meshes.p;
    s.mesh_data = (const uint8_t *)c->mesh_data.p;
    s.tex_offsets = (const float *)c->tex_offsets.p;
    s.accels = (const MeshAccel *)c->accels.p;
    s.accel_data = (const uint8_t *)c->accel_data.p;
    c->n_materials = d->n_materials;
    fill_mesh_nodes(s, d->nodes, d->n_nodes);
    s.use_wide = wide_now ? 1 : 0;
    for (const MeshAccel &A : accel.meshes) if (!A.wide_stack) s.use_wide = 0;
    s.background = d->background; s.environment = d->environment;
    s.all_light_intensity = allI;
    s.cam_pos = fc.cam_pos; s.top_left = fc.top_left; s.dd_x = fc.dd_x; s.dd_y = fc.dd_y;
    s.width = d->camera.width; s.height = d->camera.height;
    c->width = s.width; c->height = s.height;
    size_t n = (size_t)s.width * s.height;
    if (c->accum.ensure(n * sizeof(float4)) || c->hit_id.ensure(n * 4) || c->z.ensure(n * 4) || c->rgb8.ensure(n * 3) || c->linear.ensure(n * 12))
        return set_err(BHRT_ERR_CUDA, "cudaMalloc(frame buffers) failed");
    c->have_scene = true;
    c->rendered_spp = 0;
    return BHRT_OK;
}

// The wide BVHs, on first use: the mesh blob comes back from the device, the tables are built again with the wide
// trees appended and replace the ones uploaded with the scene.
static int ensure_wide(bhrt_ctx *c)
{
    if (c->wide_built || !c->have_scene) return BHRT_OK;
    c->wide_built = true;
    if (c->host_meshes.empty()) return BHRT_OK;
    CU(cudaSetDevice(c->device));
    std::vector<uint8_t> blob(c->host_mesh_bytes);
    CU(cudaMemcpyAsync(blob.data(), c->mesh_data.p, blob.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    bhrt_scene_desc d;
    memset(&d, 0, sizeof(d));
    d.n_meshes = (int32_t)c->host_meshes.size(); d.meshes = c->host_meshes.data(); d.mesh_data = blob.data(); d.mesh_bytes = blob.size();
    AccelBlob accel;
    std::string aerr;
    if (!build_accel(d, accel, &aerr, true)) return set_err(BHRT_ERR_ARG, "wide BVH: " + aerr);
    int r;
    if ((r = upload(c, c->accels, accel.meshes.data(), sizeof(MeshAccel) * accel.meshes.size()))) return r;
    if ((r = upload(c, c->accel_data, accel.data.data(), accel.data.size()))) return r;
    CU(cudaStreamSynchronize(c->stream));
    c->view.accels = (const MeshAccel *)c->accels.p;
    c->view.accel_data = (const uint8_t *)c->accel_data.p;
    c->view.use_wide = 1;
    for (const MeshAccel &A : accel.meshes) if (!A.wide_stack) c->view.use_wide = 0;
    return BHRT_OK;
}

// One frame (or one rank's share of it) through the wavefront kernels of wavefront.cuh, with c->max_frames
// levels of Shade() recursion per lane; *overflow = levels that were refused for lack of frames.
static int launch_render_once(bhrt_ctx *c, const bhrt_params *p, bhrt_outputs *out, unsigned long long *overflow)
{
    if (!c || !p) return set_err(BHRT_ERR_ARG, "bhrt_render: null argument");
    if (!c->have_scene) return set_err(BHRT_ERR_STATE, "bhrt_render: no scene uploaded");
    if (p->mode != BHRT_MODE_PATH && p->mode != BHRT_MODE_DET) return set_err(BHRT_ERR_ARG, "bhrt_render: bad mode");
    CU(cudaSetDevice(c->device));
    WaveArgs a;
    memset(&a, 0, sizeof(a));
    a.cfg.mode = p->mode;
    a.cfg.gi_depth = p->gi_depth;
    a.cfg.tir_budget = p->tir_budget > 0 ? p->tir_budget : 16;
    a.cfg.max_frames = c->max_frames;
    c->view.face_materials = (p->flags & BHRT_FLAG_FACE_MATERIALS) ? 1 : 0;
    a.seed = p->seed;
    int spp = p->mode == BHRT_MODE_DET ? 1 : p->spp;
    if (spp <= 0) return set_err(BHRT_ERR_ARG, "bhrt_render: spp must be > 0");
    int sample_begin = p->sample_begin, sample_end = p->sample_end;
    if (sample_begin == 0 && sample_end == 0) sample_end = spp;
    if (sample_begin < 0 || sample_end > spp || sample_begin > sample_end) return set_err(BHRT_ERR_ARG, "bhrt_render: bad sample range");
    int row_begin = p->row_begin, row_end = p->row_end;
    if (row_begin == 0 && row_end == 0) row_end = c->height;
    if (row_begin < 0 || row_end > c->height || row_begin > row_end) return set_err(BHRT_ERR_ARG, "bhrt_render: bad row range");
    a.row_begin = row_begin;
    a.row_stride = p->row_stride > 1 ? p->row_stride : 1;
    a.n_rows = (row_end - row_begin + a.row_stride - 1) / a.row_stride;
    a.tiles_x = (c->width + 7) / 8;
    // work items walk super-tiles of WF_ST x WF_ST tiles (wavefront.cuh: work_to_pixel); the padding tiles of the last
    // row / column of super-tiles are items without a pixel (skipped)
    a.stiles_x = (a.tiles_x + WF_ST - 1) / WF_ST;
    a.n_pix_work = (unsigned)a.stiles_x * (unsigned)(((a.n_rows + 7) / 8 + WF_ST - 1) / WF_ST) * (unsigned)(WF_ST * WF_ST * 64);
    a.gather.use = 0;
    if (p->use_photon_map) {
        if (!c->photons.built) return set_err(BHRT_ERR_STATE, "bhrt_render: use_photon_map set but no photon map built");
        const float radius = p->photon_radius > 0 ? p->photon_radius : 0.5f;
        if (radius != c->photons.built_radius) {            // the grid cell follows the gather radius
            std::string perr;
            int pr = c->photons.build(c->photons.built_total, radius, c->stream, c->launches, perr);
            if (pr) return set_err(pr, perr);
        }
        a.gather = c->photons.gather_view(p->photon_k > 0 ? p->photon_k : 1000, radius);
    }
    // ---- lane pools (one per CTA) and the shared-memory layout
    const size_t n_pix_share = (size_t)a.n_pix_work;
    const int n_spp_share = sample_end - sample_begin;
    int chunk = n_spp_share;
    const size_t max_samples = (size_t)1 << 27;                 // 2 GiB of float4 sample colours per chunk
    if (chunk > 0 && n_pix_share * (size_t)chunk > max_samples) chunk = (int)(max_samples / n_pix_share);
    if (chunk < 1) chunk = 1;
    a.pool = (c->pool_lanes + 31u) & ~31u;
    // photon-mapped frames are dominated by the caustic gathers, whose phase balances its 32 warps the
    // better the more queries it has: four times the lanes, gathered when nearly all of them are parked
    if (a.gather.use && !c->pool_lanes_set) a.pool = 12288;
    auto al16 = [](int v) { return (v + 15) & ~15; };
    // the shared-memory layout; a pool whose tables do not fit next to the scene's (large node / material tables) shrinks
    const int smem_limit = 227 * 1024 - 2048;           // (static shared memory of the kernel included)
    // (split gather: the render kernel only parks the queries -- no per-warp scratch, one list instead of two)
    a.ext_gather = a.gather.use && c->gather_mode != 1 ? 1 : 0;
    // the hot part of every thread's BVH traversal (binary traversal; nothing for a scene without meshes)
    const int hot_bytes = c->view.n_mesh_nodes > 0 ? WF_BLOCK * (int)sizeof(MeshTravHot) : 0;
    for (;;) {
        a.smem_nodes = 0;
        a.smem_materials = al16(a.smem_nodes + c->view.n_nodes * (int)sizeof(bhrt_node));
        a.smem_lights = al16(a.smem_materials + c->view.n_materials * (int)sizeof(bhrt_material));
        a.smem_offsets = al16(a.smem_lights + c->view.n_lights * (int)sizeof(bhrt_light));
        a.smem_wave = al16(a.smem_offsets + 2 * (BHRT_TEX_SAMPLES - 1) * (int)sizeof(float));
        a.smem_gather = al16(a.smem_wave + wf_wave_smem_bytes(a.pool, a.gather.use != 0 && !a.ext_gather));
        // (the per-warp gather scratch only when a photon map is gathered in this kernel: without it the CTA stays small)
        a.smem_oct = al16(a.smem_gather + (!a.gather.use ? 0 : (a.ext_gather ? (int)a.pool * 2 : (WF_BLOCK / 32) * (int)sizeof(PmWarpScratch) + (int)a.pool * 4)));
        if (a.smem_oct + hot_bytes <= smem_limit || a.pool <= 1024u) break;
        a.pool -= 1024u;
    }
    if (a.smem_oct + hot_bytes > smem_limit) return set_err(BHRT_ERR_LIMIT, "bhrt_render: scene tables do not fit in shared memory");
    // a CTA claims its work in batches of whole super-tiles: one (4096 items) normally, two for photon-mapped frames
    // (measured on BASELINE config 5: 8192 items 7.6 s, 16384 items 8.1 s -- larger batches put whole caustic regions
    // on a few CTAs and the others wait for them)
    a.batch = (unsigned)(WF_ST * WF_ST * 64);
    if (a.gather.use) a.batch *= 2;
    {
        // ... but a share with little work (one GPU's 2 spp of a 16-spp frame split over 8) would leave each CTA only a
        // handful of such batches and the frame would end with most CTAs waiting for the last ones (measured: 29.7 ms
        // for 1/8 of a 188 ms frame): keep at least ~32 batches per CTA, down to 512 items
        const unsigned long long per_cta = (unsigned long long)a.n_pix_work * (unsigned long long)(chunk > 0 ? chunk : 1) / (unsigned long long)(c->sm_count > 0 ? c->sm_count : 1) / 32ull;
        while (a.batch > 512u && (unsigned long long)a.batch > per_cta) a.batch /= 2;
    }
    while ((unsigned long long)(WF_RING - 1) * a.batch < (unsigned long long)a.pool + 32ull) a.batch *= 2;     // a pass must never run dry (wf_claim_topup)
    if (const char *e = getenv("BHRT_WORK_BATCH")) { const long v = atol(e); if (v >= 1) a.batch = (unsigned)v * (unsigned)(WF_ST * WF_ST * 64); }
    a.mesh_budget = c->mesh_budget;
    a.mesh_batch = c->mesh_batch;
    a.gather_batch = c->gather_batch ? c->gather_batch : a.pool - a.pool / 12;
    a.trav_min = c->trav_min; a.trav_max = c->trav_max; a.trav_pct = c->trav_pct;
    // stochastic frames run the fast-math build of the kernel unless the caller asks for the exact one; its
    // traversal phase is one ray per thread over the binary tree, or -- BHRT_FLAG_WIDE_BVH -- 8 lanes per ray over
    // the wide tree (per-octet stacks in shared memory) when every mesh has one and the stacks fit
    const bool fast = p->mode == BHRT_MODE_PATH && !(p->flags & BHRT_FLAG_EXACT_MATH);
    const bool want_wide = c->traversal ? c->traversal == 1 : (p->flags & BHRT_FLAG_WIDE_BVH) != 0;
    if (fast && want_wide) if (int wr = ensure_wide(c)) return wr;
    const int smem_wide = al16(a.smem_oct + (WF_BLOCK / 8) * (int)sizeof(OctShared));
    a.smem_hot = a.smem_oct;                            // (binary traversal: the hot part of every thread's BVH traversal)
    const int smem_binary = al16(a.smem_oct + hot_bytes);
    const int variant = !fast ? 0 : ((want_wide && c->view.use_wide && smem_wide <= 200 * 1024) ? 1 : 2);
    // (BHRT_EXTRA_SMEM: unused dynamic shared memory on top, i.e. that much less L1 -- the experiment behind DESIGN.md 4.1's
    // "what would a shared-memory home for the traversal state cost the rest of the kernel")
    const int smem_bytes = (variant == 1 ? smem_wide : smem_binary) + (getenv("BHRT_EXTRA_SMEM") ? atoi(getenv("BHRT_EXTRA_SMEM")) : 0);
    if (c->kernel_smem[variant] != smem_bytes) {
        int per_sm = 0;
        if (variant == 0) {
            CU(cudaFuncSetAttribute(wf_render, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
            CU(cudaFuncSetAttribute(wf_render, cudaFuncAttributePreferredSharedMemoryCarveout, WF_MIN_BLOCKS * (smem_bytes + 2048) * 100 / (228 * 1024) + 1));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_render, WF_BLOCK, smem_bytes));
        } else if (int e = (variant == 1 ? bhrt_fast_prepare : bhrt_fastb_prepare)(smem_bytes, &per_sm))
            return set_err(BHRT_ERR_CUDA, std::string("bhrt_render: ") + cudaGetErrorString((cudaError_t)e));
        if (per_sm < 1) return set_err(BHRT_ERR_LIMIT, "bhrt_render: scene tables do not fit in shared memory");
        c->render_blocks[variant] = per_sm;
        c->kernel_smem[variant] = smem_bytes;
    }
    // grid: one resident CTA per slot, but no more CTAs than there is work for their pools
    unsigned grid = (unsigned)(c->sm_count * c->render_blocks[variant]);
    {
        const unsigned long long work = (unsigned long long)n_pix_share * (unsigned long long)chunk;
        const unsigned long long need = (work + a.pool - 1) / a.pool;
        if (need < grid) grid = need < 1 ? 1u : (unsigned)need;
    }
    const size_t lanes_total = (size_t)grid * a.pool;
    if (c->lanes.ensure(lanes_total * sizeof(Lane)) || c->frames.ensure(lanes_total * sizeof(Frame) * (size_t)c->max_frames) ||
        c->samples.ensure(n_pix_share * (size_t)chunk * sizeof(float4)) || c->wave_counters.ensure(sizeof(WaveCounters)))
        return set_err(BHRT_ERR_CUDA, "cudaMalloc(wavefront buffers) failed");
    a.lanes = (Lane *)c->lanes.p; a.frames = (Frame *)c->frames.p; a.samples = (float4 *)c->samples.p;
    a.hit_id = (int *)c->hit_id.p; a.zbuf = (float *)c->z.p;
    a.cnt = (WaveCounters *)c->wave_counters.p;
    a.cancel = (const unsigned *)c->cancel_flag.p;
    // split gather: the caustic queries are evaluated between launches of wf_render, device-wide and sorted by photon cell
    a.gather_prof = getenv("BHRT_DEBUG") ? 1 : 0;
    a.exit_live = 0;            // (measured on config 5: leaving with up to 1/32 ... 1/4 of the lanes still live is no faster than waiting for all)
    if (const char *e = getenv("BHRT_EXIT_LIVE")) { const long v = atol(e); a.exit_live = v > 0 ? a.pool / (unsigned)v : 0u; }
    a.key_mask = 0x3fu;
    if (const char *e = getenv("BHRT_KEY_MASK")) a.key_mask = (unsigned)strtoul(e, nullptr, 0) & 0x3fu;
    a.gather_defer = PMQ_DEFER;
    if (const char *e = getenv("BHRT_GATHER_DEFER")) { const long v = atol(e); if (v >= 1) a.gather_defer = (unsigned)v; }
    if (a.ext_gather) {
        if (c->gq_list.ensure(lanes_total * 4) || c->gq_sorted.ensure(lanes_total * 4) || c->gq_keys.ensure(lanes_total * 4) || c->gq_keys_sorted.ensure(lanes_total * 4) ||
            c->gq_small.ensure(64) || c->cta_state.ensure((size_t)grid * wf_state_bytes(a.pool)))
            return set_err(BHRT_ERR_CUDA, "cudaMalloc(split gather buffers) failed");
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, (const uint32_t *)c->gq_keys.p, (uint32_t *)c->gq_keys_sorted.p, (const unsigned *)c->gq_list.p, (unsigned *)c->gq_sorted.p,
                                        (int)lanes_total, 0, 32, c->stream);
        if (c->gq_cub.ensure(need)) return set_err(BHRT_ERR_CUDA, "cudaMalloc(split gather sort) failed");
        a.gq_list = (unsigned *)c->gq_list.p; a.gq_total = (unsigned *)c->gq_small.p; a.cta_state = (unsigned char *)c->cta_state.p;
        CU(cudaMemsetAsync(c->gq_small.p, 0, 64, c->stream));
    }
    c->frame_spp = spp; c->frame_pixels = (long long)c->width * a.n_rows;
    c->frame_items = (long long)a.n_pix_work * (long long)n_spp_share;

    // rows / samples outside this launch's share must read as zero for the multi-GPU sum
    const bool partial_rows = a.n_rows != c->height;
    if (partial_rows || n_spp_share == 0)
        CU(cudaMemsetAsync(c->accum.p, 0, (size_t)c->width * c->height * sizeof(float4), c->stream));
    if (partial_rows || sample_begin != 0) {
        CU(cudaMemsetAsync(c->hit_id.p, 0, (size_t)c->width * c->height * 4, c->stream));
        CU(cudaMemsetAsync(c->z.p, 0, (size_t)c->width * c->height * 4, c->stream));
    }
    CU(cudaEventRecord(c->ev0, c->stream));
    CU(cudaMemsetAsync(c->wave_counters.p, 0, sizeof(WaveCounters), c->stream));
    for (int s0 = sample_begin; s0 < sample_end && a.n_rows > 0; s0 += chunk) {
        a.sample_first = s0;
        a.chunk_spp = (sample_end - s0 < chunk) ? (sample_end - s0) : chunk;
        a.n_work = (unsigned long long)a.n_pix_work * (unsigned long long)a.chunk_spp;
        CU(cudaMemsetAsync(&((WaveCounters *)c->wave_counters.p)->next_work, 0, sizeof(unsigned long long), c->stream));
        a.resume = 0;
        for (;;) {
            if (variant) {
                if (int e = (variant == 1 ? bhrt_fast_launch : bhrt_fastb_launch)(&c->view, sizeof(c->view), &a, sizeof(a), grid, smem_bytes, c->stream))
                    return set_err(BHRT_ERR_CUDA, std::string("wf_render (fast-math build): ") + cudaGetErrorString((cudaError_t)e));
            } else wf_render<<<grid, WF_BLOCK, smem_bytes, c->stream>>>(c->view, a);
            c->launches++;
            if (!a.ext_gather) break;
            // the round's parked queries: sort by photon cell, evaluate, carry on (until a round parks nothing: the chunk is done)
            unsigned nq = 0;
            CU(cudaMemcpyAsync(&nq, a.gq_total, 4, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            if (nq == 0) break;
            if ((size_t)nq > lanes_total) return set_err(BHRT_ERR_STATE, "split gather: more queries than lanes");
            unsigned *const small = (unsigned *)c->gq_small.p;         // [0] query count, [1] cursor of pm_gather_queries
            uint32_t key_max = 1;
            for (int ax = 0; ax < 3; ax++) key_max *= (uint32_t)(a.gather.dim[ax] + 4);
            int bits = 1;
            while (bits < 32 && (1ull << bits) < (unsigned long long)key_max) bits++;
            const unsigned ggrid = (unsigned)(c->sm_count * 2);
            if (variant) {
                if (int e = (variant == 1 ? bhrt_fast_gather_keys : bhrt_fastb_gather_keys)(&a, sizeof(a), nq, bits, (uint32_t *)c->gq_keys.p, c->stream))
                    return set_err(BHRT_ERR_CUDA, std::string("pm_query_keys: ") + cudaGetErrorString((cudaError_t)e));
            } else pm_query_keys<<<(nq + 255u) / 256u, 256, 0, c->stream>>>(a, nq, bits, (uint32_t *)c->gq_keys.p);
            size_t tmp_bytes = c->gq_cub.bytes;
            CU(cub::DeviceRadixSort::SortPairs(c->gq_cub.p, tmp_bytes, (const uint32_t *)c->gq_keys.p, (uint32_t *)c->gq_keys_sorted.p, (const unsigned *)c->gq_list.p,
                                               (unsigned *)c->gq_sorted.p, (int)nq, 0, bits + 3 <= 32 ? bits + 3 : bits, c->stream));
            CU(cudaMemsetAsync(small, 0, 16, c->stream));
            // (the sorted keys are not needed any more: their buffer takes the queries the warp routine defers to the CTA-wide one)
            unsigned *const heavy = (unsigned *)c->gq_keys_sorted.p;
            if (variant) {
                if (int e = (variant == 1 ? bhrt_fast_gather_queries : bhrt_fastb_gather_queries)(&a, sizeof(a), (const unsigned *)c->gq_sorted.p, nq, small, heavy, ggrid, c->stream))
                    return set_err(BHRT_ERR_CUDA, std::string("pm_gather_queries: ") + cudaGetErrorString((cudaError_t)e));
            } else {
                pm_gather_queries<<<ggrid, PMQ_BLOCK, 0, c->stream>>>(a, (const unsigned *)c->gq_sorted.p, nq, small + 1, heavy, small + 2);
                pm_gather_heavy<<<ggrid, PMQ_BLOCK, 0, c->stream>>>(a, heavy, small + 2, small + 3);
            }
            c->launches += 5;
            a.resume = 1;
            if (a.gather_prof) {
                unsigned long long pp[4]; unsigned sm4[4];
                CU(cudaMemcpyAsync(pp, &((WaveCounters *)c->wave_counters.p)->pmprof[12], sizeof(pp), cudaMemcpyDeviceToHost, c->stream));
                CU(cudaMemcpyAsync(sm4, small, sizeof(sm4), cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                fprintf(stderr, "[bhrt] gather round: %u queries, %u of them CTA-wide; so far: slowest warp query %.2f Mcycles, %llu above 1 Mcycle, mean %.1f kcycles\n", nq, sm4[2], pp[0] / 1e6, pp[1], pp[3] ? pp[2] / (double)pp[3] / 1e3 : 0.0);
            }
        }
        wf_reduce<<<(a.n_pix_work + 255) / 256, 256, 0, c->stream>>>(c->view, a, (float4 *)c->accum.p, s0 == sample_begin ? 1 : 0);
        c->launches += 1;
    }
    unsigned long long rays[4] = { 0, 0, 0, 0 };      // rays_closest, rays_shadow, next_work, depth_overflow
    CU(cudaMemcpyAsync(rays, &((WaveCounters *)c->wave_counters.p)->rays_closest, sizeof(rays), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->rendered_spp = n_spp_share;
    *overflow = rays[3];
    c->last_lanes_total = lanes_total;
    if (getenv("BHRT_DEBUG")) {
        WaveCounters h;
        CU(cudaMemcpy(&h, c->wave_counters.p, sizeof(h), cudaMemcpyDeviceToHost));
        const double tot = (double)(h.prof[0] + h.prof[1] + h.prof[2]);
        fprintf(stderr, "[bhrt] grid %u pool %u: lanes pass %.1f%%  sort %.1f%%  traversal %.1f%%  | %.0f iterations, %.0f kcycles per CTA\n", grid, a.pool,
                100.0 * h.prof[0] / tot, 100.0 * h.prof[1] / tot, 100.0 * h.prof[2] / tot, (double)h.prof[3] / grid, tot / grid / 1e3);
        if (h.pmprof[0]) {
            const double q = (double)h.pmprof[0];
            fprintf(stderr, "[bhrt] gather: %.0f queries; per query: %.2f level iterations, %.0f records in the block at the level used, %.0f qualifying when selected; "
                            "warp cycles: setup %.0f  pass A %.0f  radix passes %.0f  final %.0f; <=k exits %.3f, refined after pass A %.3f, extra radix passes %.3f, selections %.3f\n",
                    q, h.pmprof[6] / q, h.pmprof[5] / q, h.pmprof[11] / (h.pmprof[8] ? (double)h.pmprof[8] : 1.0), h.pmprof[1] / q, h.pmprof[2] / q, h.pmprof[3] / q, h.pmprof[4] / q,
                    h.pmprof[7] / q, h.pmprof[10] / q, h.pmprof[9] / q, h.pmprof[8] / q);
        }
    }
    if (out) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        out->rays_closest = rays[0]; out->rays_shadow = rays[1]; out->kernel_ms = ms;
    }
    return BHRT_OK;
}

// The reference's Shade() recursion has no bound of its own (a plain refraction consumes nothing of the
// internal-reflection budget, MtlBlinn.cpp:476-519).  Ours is bounded by the frames a lane owns: when a
// frame ran into that bound, render it again with twice the depth (same seed, same samples), as long as
// the frame stacks fit in half of the free device memory.
static int launch_render(bhrt_ctx *c, const bhrt_params *p, bhrt_outputs *out)
{
    double ms_all = 0;      // device time of EVERY pass: a frame that had to be rendered again is not reported faster than it was
    for (;;) {
        unsigned long long overflow = 0;
        const int r = launch_render_once(c, p, out, &overflow);
        if (out) { ms_all += out->kernel_ms; out->kernel_ms = ms_all; }
        if (r != BHRT_OK || overflow == 0 || c->max_frames * 2 > BHRT_MAX_FRAMES_LIMIT) return r;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return r;
        const size_t need = c->last_lanes_total * sizeof(Frame) * (size_t)c->max_frames * 2;
        if (need > (free_b + c->frames.bytes) / 2) return r;
        c->max_frames *= 2;
        if (getenv("BHRT_DEBUG")) fprintf(stderr, "[bhrt] %llu Shade() levels did not fit: rendering again with %d levels per lane\n", overflow, c->max_frames);
    }
}

int bhrt_render_device(bhrt_ctx *c, const bhrt_params *p, bhrt_outputs *out) { return launch_render(c, p, out); }

// ---- asynchronous frames: the reference viewer's BeginRender / IsRenderDone / StopRender contract ----
int bhrt_render_async(bhrt_ctx *c, const bhrt_params *p, bhrt_outputs *out)
{
    if (!c || !p) return set_err(BHRT_ERR_ARG, "bhrt_render_async: null argument");
    if (!c->have_scene) return set_err(BHRT_ERR_STATE, "bhrt_render_async: no scene uploaded");
    if (c->async_state.load() != 0) return set_err(BHRT_ERR_STATE, "bhrt_render_async: a frame is already in flight (bhrt_wait first)");
    CU(cudaSetDevice(c->device));
    CU(cudaMemset(c->cancel_flag.p, 0, 16));
    c->cancel_requested = 0;
    c->frame_items = 0; c->frame_pixels = (long long)c->width * c->height;
    c->async_status = BHRT_OK; c->async_error.clear();
    c->async_state = 1;
    c->worker = std::thread([c, p, out]() {
        int r = bhrt_render(c, p, out);
        cudaMemset(c->cancel_flag.p, 0, 16);                    // a cancel ends with its frame (the stream is idle again here)
        if (r != BHRT_OK) c->async_error = g_err;               // (the worker's thread-local message)
        else if (c->cancel_requested.load()) { r = BHRT_ERR_CANCELLED; c->async_error = "the frame was cancelled (bhrt_cancel)"; }
        c->async_status = r;
        c->async_state = 2;
    });
    return BHRT_OK;
}

int bhrt_progress(bhrt_ctx *c, int64_t *pixels_done, int64_t *pixels_total, int32_t *finished)
{
    if (!c) return set_err(BHRT_ERR_ARG, "bhrt_progress: null ctx");
    const int state = c->async_state.load();
    long long done_items = 0;
    if (state == 1 && c->wave_counters.p) {
        // read the claim counter next to the running kernel, on the second stream
        cudaSetDevice(c->device);
        unsigned long long v = 0;
        if (cudaMemcpyAsync(&v, &((WaveCounters *)c->wave_counters.p)->progress, sizeof(v), cudaMemcpyDeviceToHost, c->aux_stream) == cudaSuccess &&
            cudaStreamSynchronize(c->aux_stream) == cudaSuccess) done_items = (long long)v;
    }
    const long long total = c->frame_pixels;
    long long done = state == 2 ? total : 0;
    if (state == 1 && c->frame_items > 0) { done = (long long)((double)done_items / (double)c->frame_items * (double)total); if (done > total) done = total; }
    if (state == 2 && c->cancel_requested.load()) done = -1;        // cancelled: the count is not known any more
    if (pixels_done) *pixels_done = done;
    if (pixels_total) *pixels_total = total;
    if (finished) *finished = state != 1;
    return BHRT_OK;
}

int bhrt_cancel(bhrt_ctx *c)
{
    if (!c) return set_err(BHRT_ERR_ARG, "bhrt_cancel: null ctx");
    if (c->async_state.load() != 1) return BHRT_OK;
    c->cancel_requested = 1;
    cudaSetDevice(c->device);
    const unsigned one = 1u;
    if (cudaMemcpyAsync(c->cancel_flag.p, &one, sizeof(one), cudaMemcpyHostToDevice, c->aux_stream) != cudaSuccess || cudaStreamSynchronize(c->aux_stream) != cudaSuccess)
        return set_err(BHRT_ERR_CUDA, "bhrt_cancel: could not reach the device");
    return BHRT_OK;
}

int bhrt_wait(bhrt_ctx *c)
{
    if (!c) return set_err(BHRT_ERR_ARG, "bhrt_wait: null ctx");
    if (c->async_state.load() == 0) return set_err(BHRT_ERR_STATE, "bhrt_wait: no frame in flight");
    if (c->worker.joinable()) c->worker.join();
    cudaSetDevice(c->device);
    cudaMemset(c->cancel_flag.p, 0, 16);        // (a cancel that arrived just as the frame ended)
    c->async_state = 0;
    if (c->async_status != BHRT_OK) return set_err(c->async_status, c->async_error);
    return BHRT_OK;
}

int bhrt_accum_buffer(bhrt_ctx *c, void **ptr, size_t *bytes)
{
    if (!c || !c->have_scene) return set_err(BHRT_ERR_STATE, "bhrt_accum_buffer: no scene uploaded");
    if (ptr) *ptr = c->accum.p;
    if (bytes) *bytes = (size_t)c->width * c->height * sizeof(float4);
    return BHRT_OK;
}

int bhrt_aov_buffers(bhrt_ctx *c, void **hit_id, void **z)
{
    if (!c || !c->have_scene) return set_err(BHRT_ERR_STATE, "bhrt_aov_buffers: no scene uploaded");
    if (hit_id) *hit_id = c->hit_id.p;
    if (z) *z = c->z.p;
    return BHRT_OK;
}

int bhrt_set_stream(bhrt_ctx *c, void *stream)
{
    if (!c) return set_err(BHRT_ERR_ARG, "bhrt_set_stream: null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    return BHRT_OK;
}

int bhrt_set_last_error(int code, const char *msg) { return set_err(code, msg ? msg : ""); }

int bhrt_device_and_stream(bhrt_ctx *c, int32_t *device, void **cuda_stream)
{
    if (!c) return set_err(BHRT_ERR_ARG, "bhrt_device_and_stream: null ctx");
    if (device) *device = c->device;
    if (cuda_stream) *cuda_stream = (void *)c->stream;
    return BHRT_OK;
}

int bhrt_frame_size(bhrt_ctx *c, int32_t *width, int32_t *height)
{
    if (!c || !c->have_scene) return set_err(BHRT_ERR_STATE, "bhrt_frame_size: no scene uploaded");
    if (width) *width = c->width;
    if (height) *height = c->height;
    return BHRT_OK;
}

int bhrt_launch_count(bhrt_ctx *c, uint64_t *launches)
{
    if (!c || !launches) return set_err(BHRT_ERR_ARG, "bhrt_launch_count: null argument");
    *launches = c->launches;
    return BHRT_OK;
}

int bhrt_resolve(bhrt_ctx *c, int32_t total_spp, bhrt_outputs *out)
{
    if (!c || !c->have_scene) return set_err(BHRT_ERR_STATE, "bhrt_resolve: no scene uploaded");
    if (total_spp <= 0) return set_err(BHRT_ERR_ARG, "bhrt_resolve: total_spp must be > 0");
    CU(cudaSetDevice(c->device));
    int n = c->width * c->height;
    resolve_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>((const float4 *)c->accum.p, n, (float)total_spp, (float *)c->linear.p, (uint8_t *)c->rgb8.p);
    c->launches++;
    CU(cudaGetLastError());
    if (out) {
        if (out->rgb8) CU(cudaMemcpyAsync(out->rgb8, c->rgb8.p, (size_t)n * 3, cudaMemcpyDeviceToHost, c->stream));
        if (out->linear) CU(cudaMemcpyAsync(out->linear, c->linear.p, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
        if (out->hit_id) CU(cudaMemcpyAsync(out->hit_id, c->hit_id.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (out->z) CU(cudaMemcpyAsync(out->z, c->z.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return BHRT_OK;
}

int bhrt_sample_counts(bhrt_ctx *c, uint8_t *host_counts)
{
    if (!c || !c->have_scene || !host_counts) return set_err(BHRT_ERR_ARG, "bhrt_sample_counts: bad argument");
    CU(cudaSetDevice(c->device));
    const int n = c->width * c->height;
    // (the RGB8 image buffer is rewritten by every resolve: borrow its first n bytes)
    sample_count_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>((const float4 *)c->accum.p, n, (uint8_t *)c->rgb8.p);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host_counts, c->rgb8.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return BHRT_OK;
}

int bhrt_render(bhrt_ctx *c, const bhrt_params *p, bhrt_outputs *out)
{
    int r = launch_render(c, p, out);
    if (r) return r;
    int spp = p->mode == BHRT_MODE_DET ? 1 : p->spp;
    return bhrt_resolve(c, spp, out);
}

int bhrt_debug_trace_mesh(bhrt_ctx *c, int32_t mesh, int32_t n, const float *rays, float *out_binary, float *out_wide, float *kernel_ms)
{
    if (!c || !c->have_scene || !rays || n <= 0) return set_err(BHRT_ERR_ARG, "bhrt_debug_trace_mesh: bad argument");
    if (mesh < 0 || (size_t)mesh * sizeof(MeshAccel) >= c->accels.bytes) return set_err(BHRT_ERR_ARG, "bhrt_debug_trace_mesh: bad mesh index");
    if (int wr = ensure_wide(c)) return wr;
    if (!c->view.use_wide) return set_err(BHRT_ERR_STATE, "bhrt_debug_trace_mesh: the scene has a mesh without a wide tree");
    CU(cudaSetDevice(c->device));
    DevBuf dr, d0, d1;
    if (dr.ensure((size_t)n * 32) || d0.ensure((size_t)n * 16) || d1.ensure((size_t)n * 16)) { dr.release(); d0.release(); d1.release(); return set_err(BHRT_ERR_CUDA, "cudaMalloc failed"); }
    cudaMemcpyAsync(dr.p, rays, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream);
    cudaEvent_t ev[3];
    for (auto &e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], c->stream);
    dbg_trace_binary<<<(n + 127) / 128, 128, 0, c->stream>>>(c->view, mesh, n, (const float *)dr.p, (float *)d0.p);
    cudaEventRecord(ev[1], c->stream);
    dbg_trace_wide<<<(unsigned)(((size_t)n * 8 + 255) / 256), 256, 0, c->stream>>>(c->view, mesh, n, (const float *)dr.p, (float *)d1.p);
    cudaEventRecord(ev[2], c->stream);
    c->launches += 2;
    if (out_binary) cudaMemcpyAsync(out_binary, d0.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream);
    if (out_wide) cudaMemcpyAsync(out_wide, d1.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess && kernel_ms) { cudaEventElapsedTime(&kernel_ms[0], ev[0], ev[1]); cudaEventElapsedTime(&kernel_ms[1], ev[1], ev[2]); }
    for (auto &x : ev) cudaEventDestroy(x);
    dr.release(); d0.release(); d1.release();
    if (e != cudaSuccess) return set_err(BHRT_ERR_CUDA, std::string("bhrt_debug_trace_mesh: ") + cudaGetErrorString(e));
    return BHRT_OK;
}

// ---- photon map ------------------------------------------------------------------------------
int bhrt_photons_emit(bhrt_ctx *c, int32_t target_stored, uint64_t seed, int32_t *stored, uint64_t *photon_rays)
{
    if (!c || !c->have_scene) return set_err(BHRT_ERR_STATE, "bhrt_photons_emit: no scene uploaded");
    CU(cudaSetDevice(c->device));
    std::string err;
    int r = c->photons.emit(c->view, c->sorted_lights, c->stream, c->sm_count, target_stored, seed, stored, (unsigned long long *)photon_rays, c->launches, err);
    if (r) return set_err(r, err);
    return BHRT_OK;
}
int bhrt_photons_emit_global(bhrt_ctx *c, int32_t target_stored, uint64_t seed, int32_t *stored, uint64_t *photon_rays)
{
    if (!c || !c->have_scene) return set_err(BHRT_ERR_STATE, "bhrt_photons_emit_global: no scene uploaded");
    CU(cudaSetDevice(c->device));
    std::string err;
    int r = c->photons.emit(c->view, c->sorted_lights, c->stream, c->sm_count, target_stored, seed, stored, (unsigned long long *)photon_rays, c->launches, err, true);
    if (r) return set_err(r, err);
    return BHRT_OK;
}
int bhrt_photons_buffer(bhrt_ctx *c, void **ptr, int32_t *count)
{
    if (!c) return set_err(BHRT_ERR_ARG, "null ctx");
    if (ptr) *ptr = c->photons.records;
    if (count) *count = c->photons.count;
    return BHRT_OK;
}
int bhrt_photons_set(bhrt_ctx *c, const void *records, int32_t count, int from_device)
{
    if (!c || !records || count < 0) return set_err(BHRT_ERR_ARG, "bhrt_photons_set: bad argument");
    CU(cudaSetDevice(c->device));
    std::string err;
    int r = c->photons.set(records, count, from_device != 0, c->stream, err);
    if (r) return set_err(r, err);
    return BHRT_OK;
}
int bhrt_photons_get(bhrt_ctx *c, void *host_records, int32_t max_count, int32_t *count)
{
    if (!c || !host_records) return set_err(BHRT_ERR_ARG, "bhrt_photons_get: bad argument");
    CU(cudaSetDevice(c->device));
    int n = c->photons.count < max_count ? c->photons.count : max_count;
    if (n > 0) CU(cudaMemcpy(host_records, c->photons.records, (size_t)n * sizeof(PhotonRecord), cudaMemcpyDeviceToHost));
    if (count) *count = n;
    return BHRT_OK;
}
int bhrt_photons_build(bhrt_ctx *c, int32_t total_stored)
{
    if (!c) return set_err(BHRT_ERR_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    std::string err;
    int r = c->photons.build(total_stored, 0.5f /* MAX_Area, MtlBlinn.cpp:29; rebuilt if bhrt_params asks for another radius */, c->stream, c->launches, err);
    if (r) return set_err(r, err);
    return BHRT_OK;
}

}  // extern "C"
