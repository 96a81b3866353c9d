// render.cu -- libbhrt.so: the sm_100a render kernels and the C ABI of include/bhrt.h.
//
// Replaces BeginRender() (reference Main.cpp:178-242): the OpenMP loop over image columns
// becomes a grid of warps, each owning an 8x4 pixel tile (coherent primary rays); every
// thread evaluates its pixel's samples with the explicit-stack estimator of core.cuh and
// adds them in sample order, so a pixel's value is independent of the launch geometry.
// A second kernel resolves sum/spp -> gamma -> Color24.
//
// There is NO CPU fallback in this library: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "core.cuh"
#include "setup.hpp"
#include "accel.hpp"
#include "photon.cuh"

using namespace bhrt;

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
    int rendered_spp = 0;
    int n_materials = 0;
    DevBuf accels, accel_data;
    bool kernel_ready = false;
    int kernel_smem = 0, blocks_per_sm = 1;
};

// ------------------------------------------------------------------------------------------
struct RenderArgs {
    ShadeCfg cfg;
    unsigned long long seed;
    int sample_begin, sample_end;
    int row_begin, row_stride;
    int n_rows;              // number of rows this launch covers
    int tiles_x;             // 8x8 pixel tiles per row of tiles
    unsigned int n_work;     // work items = tiles * 64 (items outside the image are skipped)
    PhotonGather gather;     // caustic photon map (use = 0 when disabled)
    int smem_nodes, smem_materials, smem_lights, smem_offsets;   // byte offsets of the staged tables
};

#define BHRT_BLOCK 128
#define BHRT_FULL 0xffffffffu

// Persistent-thread renderer.  The grid is sized to fill the machine once (SMs x resident CTAs);
// every lane owns ONE pixel at a time and evaluates its samples in order (so a pixel's sum is
// independent of launch geometry and GPU count), then pulls the next pixel from a global counter
// -- one warp-aggregated atomic per refill.  Pixels are numbered in 8x8 tiles so the lanes of a
// warp start on neighbouring primary rays.
//
// Each loop iteration has exactly one traversal (scene_trace) that all lanes with work enter
// together, whatever their sample's recursion level or stage; sample_advance() then runs each
// lane's estimator coroutine to its next ray.  The scene-graph node table, materials, lights and
// the texture-filter taps are staged in shared memory once per CTA: the node loop of scene_trace
// reads them at warp-uniform addresses (one broadcast per access).
__global__ void __launch_bounds__(BHRT_BLOCK)
render_kernel(SceneView s, RenderArgs a, float4 *__restrict__ accum, int *__restrict__ hit_id, float *__restrict__ zbuf,
              unsigned long long *__restrict__ counters)
{
    extern __shared__ __align__(16) unsigned char smem[];
    {
        // cooperative 32-bit copies of the small tables (all sizes are multiples of 4)
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
    const int lane = threadIdx.x & 31;
    const bool det = a.cfg.mode == BHRT_MODE_DET;
    RayCounters rc; rc.closest = 0; rc.shadow = 0;
    Lane L;
    int pi = 0, pj = 0;             // the lane's pixel
    bool have_pixel = false, in_sample = false, exhausted = false;
    int sidx = 0, id0 = -1;
    float z0 = BHRT_BIGFLOAT;
    V3 sum(0, 0, 0);
    for (;;) {
        // ---- refill idle lanes (warp-aggregated fetch)
        const bool need = !have_pixel && !exhausted;
        const unsigned nm = __ballot_sync(BHRT_FULL, need);
        if (nm) {
            const int leader = __ffs(nm) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd((unsigned int *)&counters[2], (unsigned)__popc(nm));
            base = __shfl_sync(BHRT_FULL, base, leader);
            if (need) {
                const unsigned idx = base + __popc(nm & ((1u << lane) - 1u));
                if (idx >= a.n_work) exhausted = true;
                else {
                    const unsigned tile = idx >> 6, w = idx & 63u;
                    const int tx = (int)(tile % (unsigned)a.tiles_x), ty = (int)(tile / (unsigned)a.tiles_x);
                    const int i = tx * 8 + (int)(w & 7u), trow = ty * 8 + (int)(w >> 3);
                    if (i < s.width && trow < a.n_rows) {
                        pi = i; pj = a.row_begin + trow * a.row_stride;
                        have_pixel = true; in_sample = false;
                        sidx = a.sample_begin; sum = V3(0, 0, 0);
                    }
                }
            }
        }
        if (!__any_sync(BHRT_FULL, have_pixel)) {
            if (!__any_sync(BHRT_FULL, !exhausted)) break;
            continue;
        }
        if (have_pixel) {
            if (!in_sample) {
                L.rng.seed(a.seed, (uint64_t)pj * s.width + pi, (uint64_t)sidx);
                Ray pr = primary_ray(s, pi, pj, !det, L.rng);
                sample_begin(L, pr, rc);
                in_sample = true;
            }
            L.hit_ok = scene_trace(s, L.q, L.hit);
            V3 c;
            if (sample_advance(L, s, a.cfg, a.gather, (float)pi / s.width, (float)pj / s.height, rc, c)) {
                if (sidx == 0) { id0 = L.hit_id; z0 = L.hit_z; }
                sum = sum + c;
                in_sample = false;
                if (++sidx == a.sample_end) {
                    const size_t p = (size_t)pj * s.width + pi;
                    accum[p] = make_float4(sum.x, sum.y, sum.z, (float)(a.sample_end - a.sample_begin));
                    if (a.sample_begin == 0) { hit_id[p] = id0; zbuf[p] = z0; }
                    have_pixel = false;
                }
            }
        }
    }
    // ray counters: warp-reduce, one atomic per warp
    unsigned long long c0 = rc.closest, c1 = rc.shadow;
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_down_sync(BHRT_FULL, c0, o);
        c1 += __shfl_down_sync(BHRT_FULL, c1, o);
    }
    if (lane == 0 && (c0 | c1)) { atomicAdd(&counters[0], c0); atomicAdd(&counters[1], c1); }
}

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
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    if (c->counters.ensure(8 * sizeof(unsigned long long))) { delete c; return set_err(BHRT_ERR_CUDA, "cudaMalloc(counters)"); }
    // the estimator keeps a frame stack + BVH stacks in local memory
    cudaDeviceSetLimit(cudaLimitStackSize, 16 * 1024);
    *out = c;
    return BHRT_OK;
}

int bhrt_destroy(bhrt_ctx *c)
{
    if (!c) return BHRT_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf *bufs[] = { &c->nodes, &c->materials, &c->lights, &c->texmaps, &c->textures, &c->texels, &c->meshes, &c->mesh_data,
                       &c->tex_offsets, &c->accels, &c->accel_data, &c->accum, &c->hit_id, &c->z, &c->rgb8, &c->linear, &c->counters };
    for (auto b : bufs) b->release();
    c->photons.release();
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return BHRT_OK;
}

static int upload(bhrt_ctx *c, DevBuf &b, const void *src, size_t bytes)
{
    if (b.ensure(bytes)) return set_err(BHRT_ERR_CUDA, "cudaMalloc failed");
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
    if ((r = upload(c, c->texmaps, d->texmaps, sizeof(bhrt_texmap) * d->n_texmaps))) return r;
    if ((r = upload(c, c->textures, d->textures, sizeof(bhrt_texture) * d->n_textures))) return r;
    if ((r = upload(c, c->texels, d->texels, d->texel_bytes))) return r;
    if ((r = upload(c, c->meshes, d->meshes, sizeof(bhrt_mesh) * d->n_meshes))) return r;
    if ((r = upload(c, c->mesh_data, d->mesh_data, d->mesh_bytes))) return r;
    if ((r = upload(c, c->tex_offsets, offs.data(), sizeof(float) * offs.size()))) return r;
    AccelBlob accel;
    std::string aerr;
    if (!build_accel(*d, accel, &aerr)) return set_err(BHRT_ERR_ARG, "bhrt_upload_scene: " + aerr);
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

static int launch_render(bhrt_ctx *c, const bhrt_params *p, bhrt_outputs *out)
{
    if (!c || !p) return set_err(BHRT_ERR_ARG, "bhrt_render: null argument");
    if (!c->have_scene) return set_err(BHRT_ERR_STATE, "bhrt_render: no scene uploaded");
    if (p->mode != BHRT_MODE_PATH && p->mode != BHRT_MODE_DET) return set_err(BHRT_ERR_ARG, "bhrt_render: bad mode");
    CU(cudaSetDevice(c->device));
    RenderArgs a;
    memset(&a, 0, sizeof(a));
    a.cfg.mode = p->mode;
    a.cfg.gi_depth = p->gi_depth;
    a.cfg.tir_budget = p->tir_budget > 0 ? p->tir_budget : 16;
    a.seed = p->seed;
    int spp = p->mode == BHRT_MODE_DET ? 1 : p->spp;
    if (spp <= 0) return set_err(BHRT_ERR_ARG, "bhrt_render: spp must be > 0");
    a.sample_begin = p->sample_begin; a.sample_end = p->sample_end;
    if (a.sample_begin == 0 && a.sample_end == 0) a.sample_end = spp;
    if (a.sample_begin < 0 || a.sample_end > spp || a.sample_begin > a.sample_end) return set_err(BHRT_ERR_ARG, "bhrt_render: bad sample range");
    int row_begin = p->row_begin, row_end = p->row_end;
    if (row_begin == 0 && row_end == 0) row_end = c->height;
    if (row_begin < 0 || row_end > c->height || row_begin > row_end) return set_err(BHRT_ERR_ARG, "bhrt_render: bad row range");
    a.row_begin = row_begin;
    a.row_stride = p->row_stride > 1 ? p->row_stride : 1;
    a.n_rows = (row_end - row_begin + a.row_stride - 1) / a.row_stride;
    a.tiles_x = (c->width + 7) / 8;
    a.n_work = (unsigned)a.tiles_x * (unsigned)((a.n_rows + 7) / 8) * 64u;
    a.gather.use = 0;
    if (p->use_photon_map) {
        if (!c->photons.built) return set_err(BHRT_ERR_STATE, "bhrt_render: use_photon_map set but no photon map built");
        a.gather = c->photons.gather_view(p->photon_k > 0 ? p->photon_k : 1000, p->photon_radius > 0 ? p->photon_radius : 0.5f);
    }
    // shared-memory staging of the small scene tables
    auto al16 = [](int v) { return (v + 15) & ~15; };
    a.smem_nodes = 0;
    a.smem_materials = al16(a.smem_nodes + c->view.n_nodes * (int)sizeof(bhrt_node));
    a.smem_lights = al16(a.smem_materials + c->n_materials * (int)sizeof(bhrt_material));
    a.smem_offsets = al16(a.smem_lights + c->view.n_lights * (int)sizeof(bhrt_light));
    const int smem_bytes = al16(a.smem_offsets + 2 * (BHRT_TEX_SAMPLES - 1) * (int)sizeof(float));
    if (!c->kernel_ready || c->kernel_smem < smem_bytes) {
        CU(cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_kernel, BHRT_BLOCK, smem_bytes));
        if (per_sm < 1) return set_err(BHRT_ERR_LIMIT, "bhrt_render: scene tables do not fit in shared memory");
        c->blocks_per_sm = per_sm;
        c->kernel_smem = smem_bytes;
        c->kernel_ready = true;
    }
    CU(cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(unsigned long long), c->stream));
    // rows / samples outside this launch's share must read as zero for the multi-GPU sum
    if (a.n_rows != c->height || a.sample_begin != 0 || a.sample_end == a.sample_begin)
        CU(cudaMemsetAsync(c->accum.p, 0, (size_t)c->width * c->height * sizeof(float4), c->stream));
    if (a.n_rows != c->height || a.sample_begin != 0) {
        CU(cudaMemsetAsync(c->hit_id.p, 0, (size_t)c->width * c->height * 4, c->stream));
        CU(cudaMemsetAsync(c->z.p, 0, (size_t)c->width * c->height * 4, c->stream));
    }
    // persistent grid: fill the machine once; never more CTAs than there is work for
    unsigned grid = (unsigned)(c->sm_count * c->blocks_per_sm);
    const unsigned need_blocks = (a.n_work + BHRT_BLOCK - 1) / BHRT_BLOCK;
    if (grid > need_blocks) grid = need_blocks;
    CU(cudaEventRecord(c->ev0, c->stream));
    if (a.n_rows > 0 && a.sample_end > a.sample_begin && grid > 0) {
        render_kernel<<<grid, BHRT_BLOCK, smem_bytes, c->stream>>>(c->view, a, (float4 *)c->accum.p, (int *)c->hit_id.p, (float *)c->z.p,
                                                                    (unsigned long long *)c->counters.p);
        c->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev1, c->stream));
    c->rendered_spp = a.sample_end - a.sample_begin;
    if (out) {
        unsigned long long cnt[2];
        CU(cudaMemcpyAsync(cnt, c->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        out->rays_closest = cnt[0]; out->rays_shadow = cnt[1]; out->kernel_ms = ms;
    } else CU(cudaStreamSynchronize(c->stream));
    return BHRT_OK;
}

int bhrt_render_device(bhrt_ctx *c, const bhrt_params *p, bhrt_outputs *out) { return launch_render(c, p, out); }

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

int bhrt_render(bhrt_ctx *c, const bhrt_params *p, bhrt_outputs *out)
{
    int r = launch_render(c, p, out);
    if (r) return r;
    int spp = p->mode == BHRT_MODE_DET ? 1 : p->spp;
    return bhrt_resolve(c, spp, out);
}

// ---- photon map ------------------------------------------------------------------------------
int bhrt_photons_emit(bhrt_ctx *c, int32_t target_stored, uint64_t seed, int32_t *stored, uint64_t *photon_rays)
{
    if (!c || !c->have_scene) return set_err(BHRT_ERR_STATE, "bhrt_photons_emit: no scene uploaded");
    CU(cudaSetDevice(c->device));
    std::string err;
    int r = c->photons.emit(c->view, c->stream, c->sm_count, target_stored, seed, stored, (unsigned long long *)photon_rays, err);
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
    int r = c->photons.build(total_stored, c->stream, err);
    if (r) return set_err(r, err);
    return BHRT_OK;
}

}  // extern "C"
