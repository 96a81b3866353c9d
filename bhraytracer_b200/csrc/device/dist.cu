// dist.cu -- the multi-GPU host of libbhrt.so (include/bhrt_dist.h): shards BeginRender() over ranks and
// exchanges over NCCL.  Host code only (compiled by nvcc for the CUDA runtime calls); it drives the kernels
// through the single-GPU C ABI of bhrt.h, so there is exactly one implementation of the render path.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../../include/bhrt_dist.h"

// ---- NCCL, bound at run time ------------------------------------------------------------------------
// (the few declarations used, as in nccl.h 2.x: the ABI of these entry points has been stable since 2.0)
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint8 = 1, ncclInt32 = 2, ncclInt64 = 4, ncclFloat32 = 7 };
enum { ncclSum = 0 };
struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
    bool load()
    {
        if (lib) return true;
        const char *names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char *n : names) if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
        bool ok = true;
        auto sym = [&](const char *n) { void *p = dlsym(lib, n); if (!p) { ok = false; err = std::string("libnccl lacks ") + n; } return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(lib); lib = nullptr; }
        return ok;
    }
} g_nccl;
}  // namespace

extern "C" int bhrt_set_last_error(int code, const char *msg);      // render.cu
static int fail(int code, const std::string &m) { return bhrt_set_last_error(code, m.c_str()); }
#define NC(call) do { int r_ = (call); if (r_ != ncclSuccess) return fail(BHRT_ERR_CUDA, std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "NCCL error")); } while (0)
#define CUD(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(BHRT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct bhrt_dist {
    bhrt_ctx *ctx = nullptr;
    int rank = 0, world = 1, device = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    long long *counts_dev = nullptr;        // [4 * world] scratch for the small all-gathers
    void *gather_dev = nullptr; size_t gather_bytes = 0;
};

extern "C" {

int bhrt_dist_shard_rows(int32_t height, int32_t rank, int32_t world, int32_t *b, int32_t *e, int32_t *s)
{
    if (world < 1 || rank < 0 || rank >= world || height < 0) return fail(BHRT_ERR_ARG, "bhrt_dist_shard_rows: rank outside the world");
    if (b) *b = rank; if (e) *e = height; if (s) *s = world;
    return BHRT_OK;
}
int bhrt_dist_shard_samples(int32_t spp, int32_t rank, int32_t world, int32_t *b, int32_t *e)
{
    if (world < 1 || rank < 0 || rank >= world || spp < 0) return fail(BHRT_ERR_ARG, "bhrt_dist_shard_samples: rank outside the world");
    const int base = spp / world, extra = spp % world;
    const int begin = rank * base + (rank < extra ? rank : extra);
    if (b) *b = begin; if (e) *e = begin + base + (rank < extra ? 1 : 0);
    return BHRT_OK;
}
int32_t bhrt_dist_shard_photons(int32_t total, int32_t rank, int32_t world)
{
    int32_t b = 0, e = 0;
    if (bhrt_dist_shard_samples(total, rank, world, &b, &e)) return -1;
    return e - b;
}

int bhrt_dist_unique_id(uint8_t id[BHRT_DIST_ID_BYTES])
{
    if (!id) return fail(BHRT_ERR_ARG, "bhrt_dist_unique_id: null id");
    if (!g_nccl.load()) return fail(BHRT_ERR_STATE, g_nccl.err);
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    static_assert(sizeof(u) == BHRT_DIST_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &u, sizeof(u));
    return BHRT_OK;
}

int bhrt_dist_create(bhrt_ctx *ctx, int32_t rank, int32_t world, const uint8_t id[BHRT_DIST_ID_BYTES], bhrt_dist **out)
{
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return fail(BHRT_ERR_ARG, "bhrt_dist_create: bad argument");
    *out = nullptr;
    if (!g_nccl.load()) return fail(BHRT_ERR_STATE, g_nccl.err);
    int device = 0; void *stream = nullptr;
    if (int r = bhrt_device_and_stream(ctx, &device, &stream)) return r;
    CUD(cudaSetDevice(device));
    bhrt_dist *d = new bhrt_dist;
    d->ctx = ctx; d->rank = rank; d->world = world; d->device = device; d->stream = (cudaStream_t)stream;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    if (int r = g_nccl.CommInitRank(&d->comm, world, u, rank)) { delete d; return fail(BHRT_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); }
    if (cudaMalloc(&d->counts_dev, sizeof(long long) * 4 * (size_t)world) != cudaSuccess) { g_nccl.CommDestroy(d->comm); delete d; return fail(BHRT_ERR_CUDA, "cudaMalloc failed"); }
    *out = d;
    return BHRT_OK;
}

int bhrt_dist_destroy(bhrt_dist *d)
{
    if (!d) return BHRT_OK;
    cudaSetDevice(d->device);
    cudaStreamSynchronize(d->stream);
    if (d->comm) g_nccl.CommDestroy(d->comm);
    cudaFree(d->counts_dev); cudaFree(d->gather_dev);
    delete d;
    return BHRT_OK;
}

// sum of `n` 64-bit counters over the ranks (in place, host values)
static int sum_counts(bhrt_dist *d, long long *v, int n)
{
    if (d->world == 1) return BHRT_OK;
    CUD(cudaMemcpyAsync(d->counts_dev, v, sizeof(long long) * n, cudaMemcpyHostToDevice, d->stream));
    NC(g_nccl.AllReduce(d->counts_dev, d->counts_dev, (size_t)n, ncclInt64, ncclSum, d->comm, d->stream));
    CUD(cudaMemcpyAsync(v, d->counts_dev, sizeof(long long) * n, cudaMemcpyDeviceToHost, d->stream));
    CUD(cudaStreamSynchronize(d->stream));
    return BHRT_OK;
}

int bhrt_dist_render(bhrt_dist *d, const bhrt_params *params, int32_t dst, bhrt_outputs *out)
{
    if (!d || !params) return fail(BHRT_ERR_ARG, "bhrt_dist_render: null argument");
    if (dst < 0 || dst >= d->world) return fail(BHRT_ERR_ARG, "bhrt_dist_render: dst outside the world");
    CUD(cudaSetDevice(d->device));
    bhrt_params p = *params;
    const bool det = p.mode == BHRT_MODE_DET;
    const int total_spp = det ? 1 : p.spp;
    int32_t width = 0, height = 0;
    if (int r = bhrt_frame_size(d->ctx, &width, &height)) return r;
    p.sample_begin = p.sample_end = 0; p.row_begin = p.row_end = 0; p.row_stride = 0;
    bool empty = false;
    if (det) bhrt_dist_shard_rows(height, d->rank, d->world, &p.row_begin, &p.row_end, &p.row_stride);
    else {
        bhrt_dist_shard_samples(total_spp, d->rank, d->world, &p.sample_begin, &p.sample_end);
        // a rank with an empty range still takes part in the collectives, with zero buffers
        empty = p.sample_end == p.sample_begin;
        if (empty) { p.sample_begin = p.sample_end = total_spp; }
    }
    bhrt_outputs st;
    memset(&st, 0, sizeof(st));
    if (int r = bhrt_render_device(d->ctx, &p, &st)) return r;
    void *accum = nullptr, *hit = nullptr, *z = nullptr; size_t accum_bytes = 0;
    if (int r = bhrt_accum_buffer(d->ctx, &accum, &accum_bytes)) return r;
    if (int r = bhrt_aov_buffers(d->ctx, &hit, &z)) return r;
    if (d->world > 1) {
        const size_t n = (size_t)width * height;
        // one group: the accumulation buffer and the sample-0 AOVs (zero off the rank / rows that own them)
        NC(g_nccl.GroupStart());
        NC(g_nccl.AllReduce(accum, accum, accum_bytes / 4, ncclFloat32, ncclSum, d->comm, d->stream));
        NC(g_nccl.AllReduce(hit, hit, n, ncclInt32, ncclSum, d->comm, d->stream));
        NC(g_nccl.AllReduce(z, z, n, ncclFloat32, ncclSum, d->comm, d->stream));
        NC(g_nccl.GroupEnd());
    }
    bhrt_outputs none;
    memset(&none, 0, sizeof(none));
    bhrt_outputs *ro = (out && d->rank == dst) ? out : &none;
    if (int r = bhrt_resolve(d->ctx, total_spp, ro)) return r;
    long long cnt[2] = { (long long)st.rays_closest, (long long)st.rays_shadow };
    if (int r = sum_counts(d, cnt, 2)) return r;
    if (out) { out->rays_closest = (uint64_t)cnt[0]; out->rays_shadow = (uint64_t)cnt[1]; out->kernel_ms = st.kernel_ms; }
    return BHRT_OK;
}

int bhrt_dist_photons(bhrt_dist *d, int32_t total_stored, uint64_t seed, int32_t *stored_out, uint64_t *rays_out)
{
    if (!d || total_stored <= 0) return fail(BHRT_ERR_ARG, "bhrt_dist_photons: bad argument");
    CUD(cudaSetDevice(d->device));
    const int32_t quota = bhrt_dist_shard_photons(total_stored, d->rank, d->world);
    int32_t stored = 0; uint64_t rays = 0;
    if (quota > 0) { if (int r = bhrt_photons_emit(d->ctx, quota, seed + 7919ull * (uint64_t)d->rank, &stored, &rays)) return r; }
    if (d->world == 1) {
        if (int r = bhrt_photons_build(d->ctx, stored)) return r;
        if (stored_out) *stored_out = stored; if (rays_out) *rays_out = rays;
        return BHRT_OK;
    }
    // the stored counts of all ranks (the emission may fall short of its quota)
    std::vector<long long> mine(2), all(2 * (size_t)d->world);
    mine[0] = stored; mine[1] = (long long)rays;
    CUD(cudaMemcpyAsync(d->counts_dev + 2 * (size_t)d->world, mine.data(), 16, cudaMemcpyHostToDevice, d->stream));
    NC(g_nccl.AllGather(d->counts_dev + 2 * (size_t)d->world, d->counts_dev, 2, ncclInt64, d->comm, d->stream));
    CUD(cudaMemcpyAsync(all.data(), d->counts_dev, 16 * (size_t)d->world, cudaMemcpyDeviceToHost, d->stream));
    CUD(cudaStreamSynchronize(d->stream));
    long long total = 0, total_rays = 0, mx = 0;
    for (int k = 0; k < d->world; k++) { total += all[2 * k]; total_rays += all[2 * k + 1]; if (all[2 * k] > mx) mx = all[2 * k]; }
    // all-gather of the 24-byte records, padded to the largest share; then pack the shares back to back
    const size_t slot = (size_t)mx * 24;
    const size_t need = slot * ((size_t)d->world + 1) + (size_t)total * 24 + 64;
    if (need > d->gather_bytes) {
        cudaFree(d->gather_dev); d->gather_dev = nullptr; d->gather_bytes = 0;
        CUD(cudaMalloc(&d->gather_dev, need));
        d->gather_bytes = need;
    }
    uint8_t *send = (uint8_t *)d->gather_dev, *recv = send + slot, *packed = recv + slot * (size_t)d->world;
    void *recs = nullptr; int32_t n_recs = 0;
    if (int r = bhrt_photons_buffer(d->ctx, &recs, &n_recs)) return r;
    if (slot) {
        CUD(cudaMemsetAsync(send, 0, slot, d->stream));
        if (stored > 0) CUD(cudaMemcpyAsync(send, recs, (size_t)stored * 24, cudaMemcpyDeviceToDevice, d->stream));
        NC(g_nccl.AllGather(send, recv, slot, ncclUint8, d->comm, d->stream));
        size_t off = 0;
        for (int k = 0; k < d->world; k++) {
            const size_t bytes = (size_t)all[2 * k] * 24;
            if (bytes) CUD(cudaMemcpyAsync(packed + off, recv + slot * (size_t)k, bytes, cudaMemcpyDeviceToDevice, d->stream));
            off += bytes;
        }
        CUD(cudaStreamSynchronize(d->stream));
    }
    if (int r = bhrt_photons_set(d->ctx, packed, (int32_t)total, 1)) return r;
    if (int r = bhrt_photons_build(d->ctx, (int32_t)total)) return r;
    if (stored_out) *stored_out = (int32_t)total; if (rays_out) *rays_out = (uint64_t)total_rays;
    return BHRT_OK;
}

}  // extern "C"
