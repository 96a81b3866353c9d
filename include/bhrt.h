/* bhrt.h -- C ABI of the B200 rendering core (libbhrt.so).
 *
 * This is the drop-in boundary for the reference's one data-parallel hot path.  The
 * reference (BosonHBC/BHRayTracer) has no plugin/FFI API; its de-facto seam is
 *     int  LoadScene(char const *filename);     reference xmlload.cpp:65
 *     void BeginRender(); void StopRender();     reference Main.cpp:178,243
 * plus the global RenderImage the viewport polls (reference Scenes/scene.h:528-645,
 * viewport.cpp:37-44,395-415).  LoadScene stays on the host (see bhrt_host.h); the
 * functions below replace BeginRender(): the host flattens its scene graph into the POD
 * tables declared here, uploads them once, and asks for frames.
 *
 * Conventions: every function returns 0 on success and a negative bhrt_status on error
 * (bhrt_last_error() gives the text); the caller owns every host buffer; the library owns
 * all device memory; one ctx is used from one thread at a time; bhrt_render* block until
 * the frame is complete.  There is no CPU fallback: bhrt_create fails without a CUDA device.
 *
 * All 3x3 matrices are column-major exactly like cy::Matrix3f (reference
 * DataStructure/cyMatrix.h:388; cell[3*c+r]).
 */
#ifndef BHRT_H_INCLUDED
#define BHRT_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BHRT_ABI_VERSION 1

typedef enum bhrt_status {
    BHRT_OK = 0,
    BHRT_ERR_ARG = -1,      /* bad argument / inconsistent scene description */
    BHRT_ERR_CUDA = -2,     /* CUDA runtime error (no device, out of memory, launch failure) */
    BHRT_ERR_STATE = -3,    /* call order (render before upload, ...) */
    BHRT_ERR_LIMIT = -4,    /* scene exceeds a compiled-in limit (BHRT_MAX_*) */
    BHRT_ERR_CANCELLED = -5 /* the frame was stopped by bhrt_cancel (StopRender): the outputs hold a partial image */
} bhrt_status;

/* limits of the constant-memory tables */
#define BHRT_MAX_NODES 256
#define BHRT_MAX_MATERIALS 128
#define BHRT_MAX_LIGHTS 32
#define BHRT_MAX_TEXMAPS 128
#define BHRT_MAX_TEXTURES 64
#define BHRT_MAX_MESHES 32
#define BHRT_MAX_DEPTH 6

/* reference Scenes/scene.h:208-246 (class Transformation): tm, its cached inverse, pos */
typedef struct bhrt_xform {
    float tm[9];
    float itm[9];
    float pos[3];
} bhrt_xform;

enum { BHRT_OBJ_NONE = 0, BHRT_OBJ_SPHERE = 1, BHRT_OBJ_PLANE = 2, BHRT_OBJ_MESH = 3 };

/* One scene-graph Node (reference Scenes/scene.h:426-502), in DFS pre-order: the order in
 * which recursive() (reference Main.cpp:389-413) visits them.  depth 1 = child of rootNode. */
typedef struct bhrt_node {
    bhrt_xform xf;
    int32_t parent;     /* index of the parent node, -1 = rootNode */
    int32_t depth;      /* 1..BHRT_MAX_DEPTH */
    int32_t kind;       /* BHRT_OBJ_* */
    int32_t material;   /* index into materials, -1 only when kind == BHRT_OBJ_NONE */
    int32_t mesh;       /* index into meshes when kind == BHRT_OBJ_MESH */
} bhrt_node;

/* reference Scenes/scene.h:394-422 (TexturedColor): colour, optionally times a TextureMap */
typedef struct bhrt_texcolor {
    float color[3];
    int32_t texmap;     /* index into texmaps, -1 = none */
} bhrt_texcolor;

/* reference Scenes/scene.h:364-386 (TextureMap = Transformation + Texture*) */
typedef struct bhrt_texmap {
    bhrt_xform xf;
    int32_t texture;    /* index into textures, -1 = null texture (samples black) */
} bhrt_texmap;

enum { BHRT_TEX_FILE = 0, BHRT_TEX_CHECKER = 1 };

/* reference Textures/texture.h:20-46 (TextureFile: RGB8 texels; TextureChecker: 2 colours) */
typedef struct bhrt_texture {
    int32_t kind;
    int32_t width, height;
    int64_t texel_offset;   /* byte offset of the first RGB8 texel in bhrt_scene_desc.texels */
    float color1[3], color2[3];
} bhrt_texture;

/* bhrt_material.flags.  BHRT_MTL_MULTI: flattened from a MultiMtl (xmlload.cpp:219-250) -- this entry is its sub-material 0,
 * which shades every face in the reference (hInfo.mtlID is never set, TriObj.cpp:68-189); bits 8..15 hold the number of
 * sub-materials, the others follow at the next indices of the material table (read only with BHRT_FLAG_FACE_MATERIALS). */
enum { BHRT_MTL_MULTI = 1 };
#define BHRT_MTL_SUBCOUNT(flags) ((int)(((uint32_t)(flags) >> 8) & 0xffu))

/* reference Materials/materials.h:20-61 (MtlBlinn).  emission / reflection /
 * reflectionGlossiness are parsed by the reference loader but never read by its renderer
 * (MtlBlinn.cpp:89-138), so they are not part of the device table.  refractionGlossiness is
 * ignored by Shade() too (MtlBlinn.cpp:115-116) but IS used where a caustic photon leaves a
 * refractive object (MtlBlinn.cpp:236). */
typedef struct bhrt_material {
    bhrt_texcolor diffuse;
    bhrt_texcolor specular;
    float refraction[3];
    float absorption[3];
    float glossiness;
    float ior;
    int32_t flags;
    float refraction_glossiness;
} bhrt_material;

enum { BHRT_LIGHT_AMBIENT = 0, BHRT_LIGHT_DIRECT = 1, BHRT_LIGHT_POINT = 2 };

/* reference Lights/lights.h:29-87 */
typedef struct bhrt_light {
    int32_t kind;
    float intensity[3];
    float v[3];         /* direction (normalised, DIRECT) or position (POINT) */
    float size;         /* POINT only */
} bhrt_light;

/* One triangle mesh (reference Objects/objects.h:45-65 TriObj = cyTriMesh + cyBVHTriMesh).
 * All offsets are BYTE offsets into bhrt_scene_desc.mesh_data and must be 16-byte aligned.
 *   v/vn/vt : float[3] per element          (cyTriMesh.h:106-111)
 *   f/fn/ft : uint32[3] per face
 *   bvh     : bhrt_bvh_node[n_bvh_nodes+1], node 0 unused, root = 1, the two children of an
 *             internal node are adjacent (cyBVH.h:76,87-100,187-203) -- the reference's own
 *             topology, needed because its mesh shadow test depends on it (TriObj.cpp:272-307)
 *   elems   : uint32 face ids referenced by the leaves */
typedef struct bhrt_bvh_node {
    float box[6];       /* xmin ymin zmin xmax ymax zmax */
    uint32_t data;      /* cyBVH.h:190-195 encoding: bit31 leaf | (count-1)<<28 | offset, or child index */
    uint32_t pad;
} bhrt_bvh_node;

typedef struct bhrt_mesh {
    int32_t nv, nf, nvn, nvt;
    int32_t n_bvh_nodes;
    int32_t n_groups;   /* usemtl groups of the OBJ (cyTriMesh mcfc, cyTriMesh.h:467-490; faces are stored group by group); when
                         * > 0, int32 group_end[n_groups] -- face f belongs to the first group g with f < group_end[g] -- follows
                         * the elems array, at elems_off + 4 * nf rounded up to 16 bytes */
    int64_t v_off, vn_off, vt_off;
    int64_t f_off, fn_off, ft_off;
    int64_t bvh_off, elems_off;
} bhrt_mesh;

/* reference Scenes/scene.h:506-524 (Camera), after LoadScene's target->dir conversion and
 * up re-orthogonalisation (xmlload.cpp:124-127) */
typedef struct bhrt_camera {
    float pos[3], dir[3], up[3];
    float fov, focaldist, dof;
    int32_t width, height;
} bhrt_camera;

typedef struct bhrt_scene_desc {
    int32_t abi_version;            /* BHRT_ABI_VERSION */
    int32_t n_nodes;     const bhrt_node *nodes;
    int32_t n_materials; const bhrt_material *materials;
    int32_t n_lights;    const bhrt_light *lights;     /* document order; the renderer sorts them like Main.cpp:116-123 */
    int32_t n_texmaps;   const bhrt_texmap *texmaps;
    int32_t n_textures;  const bhrt_texture *textures;
    int32_t n_meshes;    const bhrt_mesh *meshes;
    const uint8_t *texels;    size_t texel_bytes;
    const uint8_t *mesh_data; size_t mesh_bytes;
    bhrt_texcolor background;       /* reference Main.cpp:28 */
    bhrt_texcolor environment;      /* reference Main.cpp:29 */
    bhrt_camera camera;
} bhrt_scene_desc;

enum {
    BHRT_MODE_PATH = 0,     /* "H-mode": HEAD's stochastic estimator (Main.cpp:143-172, MtlBlinn.cpp) */
    BHRT_MODE_DET = 1       /* "D-mode": HEAD with the four determinism switches of SURVEY.md 8(c) */
};

typedef struct bhrt_params {
    int32_t mode;           /* BHRT_MODE_* */
    int32_t spp;            /* PT_SampleCount (Main.cpp:141); forced to 1 in D-mode */
    int32_t gi_depth;       /* GIBounceCount (Main.cpp:130), default 3 */
    int32_t tir_budget;     /* INTERNAL_REFLECTION_BOUNCE (Main.cpp:41), default 16 */
    uint64_t seed;          /* counter-based RNG key: results depend on (seed,pixel,sample) only */
    int32_t sample_begin;   /* this call renders samples [sample_begin, sample_end) of every pixel ... */
    int32_t sample_end;     /* ... (0,0 = all spp); used for sample-range sharding across GPUs */
    int32_t row_begin;      /* and only rows [row_begin,row_end) (0,0 = all); tile sharding */
    int32_t row_end;
    int32_t row_stride;     /* >1: rows row_begin, row_begin+stride, ... (interleaved tiles); 0/1 = contiguous */
    int32_t use_photon_map; /* add the caustic gather term (MtlBlinn.cpp:329-342) */
    int32_t photon_k;       /* MAX_PhotonCountInArea, default 1000 */
    float photon_radius;    /* MAX_Area, default 0.5 */
    uint32_t flags;         /* BHRT_FLAG_*; 0 = defaults */
} bhrt_params;

/* The stochastic mode (BHRT_MODE_PATH) is judged statistically (RMSE against the reference's
 * high-spp render), so by default it runs a build of the kernels with approximate division, square
 * root and transcendentals and with FMA contraction.  This flag selects the exact build (IEEE
 * operations in the reference's order) for it as well.  The deterministic mode is always exact. */
#define BHRT_FLAG_EXACT_MATH 1u
/* Triangle meshes are traversed one ray per thread over the reference's binary BVH (TriObj.cpp:192-307 made
 * iterative).  With this flag the fast build instead runs EIGHT LANES PER RAY over an 8-wide regrouping of the same
 * BVH (same boxes, same leaf order; per-ray stacks in shared memory; csrc/device/wide.cuh).  Both give the same
 * answers ray by ray (tests/test_gpu_wide_bvh.py); the flag exists because which one is faster depends on the
 * scene (DESIGN.md 4.1 has the measurements).  Ignored by the exact build, by scenes with a mesh too deep for
 * the wide stack, and when the scene tables plus the stacks do not fit in shared memory.
 * Environment override for all renders of a ctx: BHRT_TRAVERSAL=wide|binary. */
#define BHRT_FLAG_WIDE_BVH 2u
/* Quirk switch: shade every face of an OBJ with the sub-material of ITS usemtl group -- what MultiMtl::Shade
 * (materials.h:71) would do if TriObj::IntersectTriangle set hInfo.mtlID, which the reference at HEAD never does
 * (TriObj.cpp:68-189), so that sub-material 0 shades the whole mesh.  Off by default: parity is against the reference as
 * it is; the switched-on behaviour is pinned against the reference with that one assignment added (oracle patch Q-1). */
#define BHRT_FLAG_FACE_MATERIALS 4u

/* Any pointer may be NULL to skip that output.  Layouts follow the reference's
 * colorArray / RenderImage (Main.cpp:202,229-230): row-major, j*W+i, row 0 at the top. */
typedef struct bhrt_outputs {
    uint8_t *rgb8;          /* [H*W*3] Color24 after gamma (cyColor.h:271) */
    float *linear;          /* [H*W*3] pre-gamma pixel colour (sum/spp) */
    int32_t *hit_id;        /* [H*W]   node index of the primary hit of sample 0, -1 = miss */
    float *z;               /* [H*W]   primary hit distance of sample 0 (1e30 on miss) */
    uint64_t rays_closest;  /* closest-hit rays traced (one per reference recursive() top-level call) */
    uint64_t rays_shadow;   /* any-hit rays traced (one per reference GenLight::Shadow call) */
    double kernel_ms;       /* device time of the render kernels (CUDA events), summed over every pass when a frame had to be
                             * rendered again with deeper Shade() stacks; the ray counters are those of the final pass */
} bhrt_outputs;

typedef struct bhrt_ctx bhrt_ctx;

int bhrt_abi_version(void);
const char *bhrt_last_error(void);

int bhrt_create(int device, bhrt_ctx **out);
int bhrt_destroy(bhrt_ctx *ctx);

/* Replaces the global scene state BeginRender() reads (Main.cpp:17-37). */
int bhrt_upload_scene(bhrt_ctx *ctx, const bhrt_scene_desc *scene);

/* BeginRender() (Main.cpp:178-242) with host outputs: renders, resolves (divide by spp,
 * gamma, 8-bit) and copies the requested outputs to the caller's HOST buffers.
 * MtlBlinn::Shade recursion is as deep as the scene makes it (the reference has no bound,
 * MtlBlinn.cpp:476-519): a ctx starts with 64 levels per camera sample in flight and renders a
 * frame again with twice as many while any level was refused (up to 4096, device memory
 * permitting); kernel_ms is the time of ALL passes, the ray counters are those of the last one.
 * With use_photon_map the frame is many launches on the ctx's stream (the render kernel alternates with the kernels of the
 * caustic gather) and the call reads a 4-byte counter back between them; kernel_ms spans all of it. */
int bhrt_render(bhrt_ctx *ctx, const bhrt_params *params, bhrt_outputs *out);

/* The same frame, outputs left in device memory (accumulation buffer = sum over the
 * rendered samples, float[H*W*4]); nothing crosses PCIe.  out may be NULL; only the
 * counters / kernel_ms fields are filled. */
int bhrt_render_device(bhrt_ctx *ctx, const bhrt_params *params, bhrt_outputs *out);

/* ---- the asynchronous seam of the reference's viewer (viewport.cpp:37-38: "Called to start rendering (renderer
 * must run in a separate thread)"; :395-415 polls RenderImage::IsRenderDone / GetNumRenderedPixels; :449-453 calls
 * BeginRender / StopRender):
 *   bhrt_render_async  BeginRender() on a worker thread of the library: returns at once; params / out must stay valid
 *                      until bhrt_wait.  One frame at a time per ctx; until bhrt_wait returns only bhrt_progress,
 *                      bhrt_cancel and bhrt_wait may be called on the ctx.
 *   bhrt_progress      RenderImage::GetNumRenderedPixels / IsRenderDone: pixels whose samples have been handed to the
 *                      GPU so far (camera samples claimed / spp), the frame's pixel count, and whether it is finished.
 *   bhrt_cancel        StopRender() (empty in the reference, Main.cpp:243): the kernels stop taking new camera samples;
 *                      samples in flight finish, so the frame ends within milliseconds.
 *   bhrt_wait          joins the worker: the frame's status (BHRT_ERR_CANCELLED after a cancel; the outputs then hold
 *                      what was resolved from the partial accumulation). */
int bhrt_render_async(bhrt_ctx *ctx, const bhrt_params *params, bhrt_outputs *out);
int bhrt_progress(bhrt_ctx *ctx, int64_t *pixels_done, int64_t *pixels_total, int32_t *finished);
int bhrt_cancel(bhrt_ctx *ctx);
int bhrt_wait(bhrt_ctx *ctx);

/* Device pointer of the accumulation buffer (float[H*W*4]: r,g,b sums + sample count) so a
 * multi-GPU host can all-reduce it (NCCL) between bhrt_render_device and bhrt_resolve. */
int bhrt_accum_buffer(bhrt_ctx *ctx, void **device_ptr, size_t *bytes);

/* Device pointers of the sample-0 AOV buffers (int32 hit_id[H*W], float z[H*W]); rows a
 * launch did not cover hold zeros, so a multi-GPU host can sum-reduce them like the
 * accumulation buffer (tile sharding). */
int bhrt_aov_buffers(bhrt_ctx *ctx, void **hit_id_device_ptr, void **z_device_ptr);

/* Divide by total_spp, gamma, 8-bit; copy requested outputs to HOST buffers (out may be NULL:
 * the resolved image then stays in device memory). */
int bhrt_resolve(bhrt_ctx *ctx, int32_t total_spp, bhrt_outputs *out);

/* RenderImage::GetSampleCount (reference scene.h:534,570; HEAD itself only ever stores 0 there, Main.cpp:214): the number
 * of camera samples accumulated in each pixel by the last render (uint8[H*W], clamped at 255) -- after a multi-GPU
 * reduce the frame's total, after a cancelled or partial (row / sample range) render what was actually taken.  Feed it
 * to bhrt_host_sample_count_image for the reference's stretched 8-bit view (scene.h:602-626).  Call it before the next
 * bhrt_resolve / bhrt_render re-uses the image buffers. */
int bhrt_sample_counts(bhrt_ctx *ctx, uint8_t *host_counts);

/* Run all further work of this ctx on the caller's CUDA stream (a cudaStream_t; NULL = the
 * ctx's own stream), so a host that interleaves collectives can order and time everything on
 * one stream. */
int bhrt_set_stream(bhrt_ctx *ctx, void *cuda_stream);

/* The CUDA device index of a ctx and the stream its work currently runs on (a cudaStream_t), for a host that
 * enqueues its own work (collectives) behind the render; and the resolution of the uploaded scene. */
int bhrt_device_and_stream(bhrt_ctx *ctx, int32_t *device, void **cuda_stream);
int bhrt_frame_size(bhrt_ctx *ctx, int32_t *width, int32_t *height);
/* Sets the calling thread's bhrt_last_error() text and returns `code` (for the layers built on this ABI). */
int bhrt_set_last_error(int code, const char *message);

/* Number of kernels this ctx has launched since bhrt_create (bench.py's gpu_launches). */
int bhrt_launch_count(bhrt_ctx *ctx, uint64_t *launches);

/* TEST HOOK (tests/test_gpu_wide_bvh.py): trace n caller-supplied rays against triangle mesh `mesh` of the
 * uploaded scene with BOTH mesh traversals of the library, in exact arithmetic -- one ray per thread over the
 * binary tree (TriObj::TraceBVHNode / TraceBVHShadow made iterative, reference TriObj.cpp:192-307) and eight
 * lanes per ray over its 8-wide regrouping -- so that a test can compare them ray by ray.
 * rays: float[n*8] = origin xyz, direction xyz (mesh-local space), bound (closest: the incoming hit distance,
 * shadow: t_max), kind (1 closest, 2 shadow).  out_*: float[n*4] = result (0/1), z, face id, front.
 * kernel_ms (may be NULL): float[2], device time of the binary and of the wide pass. */
int bhrt_debug_trace_mesh(bhrt_ctx *ctx, int32_t mesh, int32_t n, const float *rays, float *out_binary, float *out_wide, float *kernel_ms);

/* Caustic photon map (reference Main.cpp:342-386, cyPhotonMap.h).
 * emit: traces photons until `target_stored` are stored on this GPU; records are the
 * reference's 24-byte Photon (cyPhotonMap.h:72-80) with UNSCALED power.  The reference's loop
 * (Main.cpp:363) never ends when the scene cannot store a caustic photon; here the number of photon
 * paths is bounded (4096 per photon asked for, at least 2^26; env BHRT_PHOTON_MAX_PATHS), and *stored
 * then comes back smaller than target_stored (possibly 0) with status BHRT_OK.
 * photons_get/set: device<->host copies of the raw records (for an NCCL all-gather done by
 * the host).  build: scales powers by 1/total_stored (Main.cpp:380) and builds the gather
 * structure over the photons currently set. */
int bhrt_photons_emit(bhrt_ctx *ctx, int32_t target_stored, uint64_t seed, int32_t *stored, uint64_t *photon_rays);
/* The GLOBAL photon map the reference also knows how to build (BuildPhotonMap / TracePhotonRay, Main.cpp:251-317, with
 * MtlBlinn::RandomPhotonBounce, MtlBlinn.cpp:140-202; its only call is commented out at Main.cpp:196 and HEAD never
 * gathers from it): same records, same get / set / build calls afterwards -- a photon is stored on every diffuse
 * surface after its first hit, refractive surfaces end the path, 30 % of the bounces are absorbed. */
int bhrt_photons_emit_global(bhrt_ctx *ctx, int32_t target_stored, uint64_t seed, int32_t *stored, uint64_t *photon_rays);
int bhrt_photons_buffer(bhrt_ctx *ctx, void **device_ptr, int32_t *count);
int bhrt_photons_set(bhrt_ctx *ctx, const void *records, int32_t count, int from_device);
int bhrt_photons_get(bhrt_ctx *ctx, void *host_records, int32_t max_count, int32_t *count);
int bhrt_photons_build(bhrt_ctx *ctx, int32_t total_stored);

#ifdef __cplusplus
}
#endif
#endif
