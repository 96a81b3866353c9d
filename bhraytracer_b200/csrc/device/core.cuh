// core.cuh -- per-ray device code of the rendering core: vector math in the reference's
// operation order, sphere/plane/box/triangle tests, scene-graph closest-hit and any-hit,
// texture sampling, and the Blinn path-tracing estimator as an explicit-stack state machine.
//
// Everything here is scalar per-ray code marked BHRT_HD so that the SAME source is compiled
// (a) by nvcc into the sm_100a kernels of render.cu -- the product -- and (b) by g++ into
// tests/emu (a test-only harness that lets the CPU-only CI compare this logic with the
// oracle).  libbhrt.so contains only (a); there is no CPU path in the library.
//
// Bit-exactness contract: every hit/miss/closest DECISION uses only
// + - * / sqrt in the reference's order, compiled without FMA contraction (-fmad=false),
// so primary hit ids are bit-identical to the reference's CPU render.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include "../../../include/bhrt.h"

#ifdef __CUDACC__
#define BHRT_HD __host__ __device__ __forceinline__
#define BHRT_HD_NOINLINE __host__ __device__ __noinline__
#else
#define BHRT_HD inline
#define BHRT_HD_NOINLINE inline
#endif
// BHRT_FN: a function kept OUT OF LINE in the kernels.  The estimator calls the samplers, the
// texture filter and the libm-grade transcendentals from dozens of sites; inlined everywhere the
// render kernel was 935 KB of SASS and spent most issue slots waiting for instruction fetch.
#define BHRT_FN BHRT_HD_NOINLINE
#define BHRT_RESTRICT __restrict__

// The whole device code lives in namespace BHRT_NS: `bhrt` in the exact build (render.cu, tests/emu)
// and `bhrt_fast` in render_fast.cu, which compiles the same source with --use_fast_math for the
// stochastic mode (see render_fast.cu).
#ifndef BHRT_NS
#define BHRT_NS bhrt
#endif
#ifndef BHRT_FAST
#define BHRT_FAST 0
#endif
namespace BHRT_NS {

#ifdef __CUDA_ARCH__
#define BHRT_ON_DEVICE 1
#else
#define BHRT_ON_DEVICE 0
#endif
#define BHRT_BIGFLOAT 1.0e30f          // reference Scenes/scene.h:38
#define BHRT_BIAS 0.0001f              // reference MtlBlinn.cpp:10, TriObj.cpp:9
#define BHRT_SHADOW_BIAS 0.00001f      // reference GenLight.cpp:5
#define BHRT_PERP_DET 0.001745f        // reference TriObj.cpp:12
#define BHRT_PI_D 3.14159265           // reference "#define PI 3.14159265" (a double literal)
#define BHRT_EULER 2.7182818f          // reference MtlBlinn.cpp:11
#define BHRT_TEX_SAMPLES 32            // reference Scenes/scene.h:18

enum { HIT_FRONT = 1, HIT_BACK = 2, HIT_FRONT_AND_BACK = 3 };   // reference scene.h:57-60

// ------------------------------------------------------------------------------------------
// Vec3f restated (reference DataStructure/cyVector.h:273-386). Operation order matters.
struct V3 {
    float x, y, z;
    BHRT_HD V3() {}
    BHRT_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
};
BHRT_HD V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
BHRT_HD V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
BHRT_HD V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
BHRT_HD V3 operator*(V3 a, float v) { return V3(a.x * v, a.y * v, a.z * v); }
BHRT_HD V3 operator*(float v, V3 a) { return V3(a.x * v, a.y * v, a.z * v); }
BHRT_HD V3 operator/(V3 a, float v) { return V3(a.x / v, a.y / v, a.z / v); }
BHRT_HD V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
BHRT_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
BHRT_HD V3 cross(V3 a, V3 p) { return V3(a.y * p.z - a.z * p.y, a.z * p.x - a.x * p.z, a.x * p.y - a.y * p.x); }
BHRT_HD float length2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
BHRT_HD float length(V3 a) { return sqrtf(length2(a)); }          // cy::Sqrt<float> = sqrtss
BHRT_HD V3 normalized(V3 a) { return a / length(a); }
BHRT_HD bool is_zero(V3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
BHRT_HD float fmaxr(float a, float b) { return a >= b ? a : b; }   // cy::Max (cyCore.h:187)
BHRT_HD float fminr(float a, float b) { return a <= b ? a : b; }   // cy::Min (cyCore.h:188)
BHRT_HD V3 ld3(const float *p) { return V3(p[0], p[1], p[2]); }
BHRT_HD float gray(V3 c) { return (c.x + c.y + c.z) / 3.0f; }     // cyColor.h:83
BHRT_HD void clamp_white(V3 &c) { if (c.x > 1) c.x = 1.f; if (c.y > 1) c.y = 1.f; if (c.z > 1) c.z = 1.f; }  // MtlBlinn.cpp:79-83
BHRT_HD bool saturated(V3 c) { return c.x >= 1 && c.y >= 1 && c.z >= 1; }
BHRT_HD bool is_nan(float v) { return v != v; }
// exact build: one out-of-line copy of each libm-grade transcendental; fast build (--use_fast_math):
// powf / sinf / cosf / tanf become a few SFU instructions and are inlined
#if BHRT_FAST
#define BHRT_MATH_FN BHRT_HD
#else
#define BHRT_MATH_FN BHRT_FN
#endif
// the estimator's helpers (samplers, light / refraction / direct-term pieces): inline in both builds
// (+7 % in the exact build, +15 % in the fast one); what must stay out of line in the exact build is
// the libm-grade transcendentals and the texture filter
#define BHRT_EST_FN BHRT_HD
#if BHRT_FAST && defined(__CUDA_ARCH__)
// powf under --use_fast_math is exp2(b * log2(a)): NaN for every negative base, where powf gives +-|a|^b when b is
// an integer -- and glossiness values are (20, 100, 20000 ...).  The Blinn lobe pow(vH.vN, glossiness) meets
// negative bases all the time in the caustic term (vL = mean photon direction points INTO the surface,
// MtlBlinn.cpp:333-338): the NaN then wiped out the whole direct term of diffuse materials (0 * NaN), and the
// photon-mapped frame of the fast build came out at less than half its brightness (found by the round-2
// convergence test on BASELINE config 5).  Same special cases as powf for the arguments this code can see.
BHRT_HD float m_pow(float a, float b)
{
    if (b == 0.f) return 1.f;
    float r = __powf(fabsf(a), b);
    if (a < 0.f) {
        const float t = truncf(b);
        if (t != b) r = __int_as_float(0x7fffffff);
        else if (fmodf(t, 2.f) != 0.f) r = -r;
    }
    return r;
}
#else
BHRT_MATH_FN float m_pow(float a, float b) { return powf(a, b); }
#endif
BHRT_MATH_FN float m_sin(float a) { return sinf(a); }
BHRT_MATH_FN float m_cos(float a) { return cosf(a); }
BHRT_MATH_FN float m_tan(float a) { return tanf(a); }
BHRT_FN float m_acos(float a) { return acosf(a); }
BHRT_FN float m_asin(float a) { return asinf(a); }
BHRT_FN float m_atan2(float a, float b) { return atan2f(a, b); }

// Matrix3f * Vec3f (reference cyMatrix.h:297-302), column-major cells
BHRT_HD V3 mat_mul(const float *m, V3 p)
{
    return V3(p.x * m[0] + p.y * m[3] + p.z * m[6],
              p.x * m[1] + p.y * m[4] + p.z * m[7],
              p.x * m[2] + p.y * m[5] + p.z * m[8]);
}
// Transformation::TransposeMult (reference scene.h:238-245): d.k = column(k) . dir
BHRT_HD V3 mat_tmul(const float *m, V3 d)
{
    return V3(m[0] * d.x + m[1] * d.y + m[2] * d.z,
              m[3] * d.x + m[4] * d.y + m[5] * d.z,
              m[6] * d.x + m[7] * d.y + m[8] * d.z);
}
BHRT_HD V3 xf_to(const bhrt_xform &t, V3 p) { return mat_mul(t.itm, p - ld3(t.pos)); }       // scene.h:220
BHRT_HD V3 xf_from(const bhrt_xform &t, V3 p) { return mat_mul(t.tm, p) + ld3(t.pos); }      // scene.h:221
BHRT_HD V3 xf_vec_from(const bhrt_xform &t, V3 d) { return mat_tmul(t.itm, d); }             // scene.h:227

struct Ray { V3 p, dir; };

struct alignas(16) Hit {    // reference scene.h:62-75 (HitInfo)
    // (the first 32 bytes are what a traversal leaves behind -- z, node, front and, for a mesh, the
    // winning triangle in uvw / duvw0.x; scene_trace_finish recomputes everything else)
    float z;
    int node;               // index into nodes, -1 = none
    bool front;
    V3 uvw, duvw0;
    V3 p, N, duvw1;
    int sub;                // sub-material of the hit face (HitInfo::mtlID, scene.h:71): 0 unless BHRT_FLAG_FACE_MATERIALS (scene_trace_finish)
    BHRT_HD void init()
    {
        z = BHRT_BIGFLOAT; node = -1; front = true; sub = 0;
        uvw = V3(0.5f, 0.5f, 0.5f); duvw0 = V3(0, 0, 0); duvw1 = V3(0, 0, 0);
        p = V3(0, 0, 0); N = V3(0, 0, 0);
    }
};

// Node::ToNodeCoords (reference scene.h:490-496)
BHRT_HD Ray to_node(const bhrt_xform &t, const Ray &r)
{
    Ray o;
    o.p = xf_to(t, r.p);
    o.dir = xf_to(t, r.p + r.dir) - o.p;
    return o;
}

#define BHRT_MESH_MASK_BITS 32
// ---- device traversal tables (built at upload by accel.hpp from the bhrt_mesh tables) --------
struct alignas(16) F4 { float x, y, z, w; };
BHRT_HD F4 ld4(const F4 *p)
{
#ifdef __CUDA_ARCH__
    float4 v = __ldg(reinterpret_cast<const float4 *>(p));      // one 128-bit read-only load
    F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
#else
    return *p;
#endif
}
BHRT_HD float u2f(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
BHRT_HD uint32_t f2u(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
// a = (vN.xyz, |vN|)   b = (vN.v0, p0.x, p0.y, p1.x)   c = (p1.y, p2.x, p2.y, bits: face | axis << 30)
struct alignas(16) TriRec { F4 a, b, c; };
// both child boxes of one internal cyBVH node + the two child words (leaf: bit31 | (count-1) << 28 |
// first TriRec; internal: PairNode index)
struct alignas(64) PairNode { float box1[6], box2[6]; uint32_t word1, word2, pad0, pad1; };
// ---- the WIDE tree (fast build, wavefront.cuh / wide.cuh): the same cyBVH collapsed to up to 8 children per
// node (accel.hpp: build_wide).  One child = 32 bytes = two 128-bit loads: its box, its word (leaf group:
// bit31 | (count-1) << 28 | first TriRec; internal: WideNode index; BHRT_WIDE_EMPTY: unused slot) and, for a
// leaf group -- up to 8 triangles, the contiguous records of a whole cyBVH subtree -- `aux` = bit i set when
// record i starts one of the reference's own leaves (the mesh shadow rule needs them, TriObj.cpp:272-291).
// Children keep the DFS order of the binary tree, so visiting slots in order visits the reference's leaves in
// the reference's order.
struct alignas(32) WideChild { float box[6]; uint32_t word, aux; };
struct alignas(256) WideNode { WideChild c[8]; };
#define BHRT_WIDE_EMPTY 0xffffffffu
#define BHRT_WIDE_STACK 64              // entries per traversal; a tree whose worst case needs more is not built (wide_stack = 0)
struct MeshAccel {
    float root_box[6];
    uint32_t root_word, max_depth;          // max_depth: levels below the root (checked against BHRT_BVH_STACK at upload)
    int64_t pairs_off, tris_off;            // byte offsets into SceneView::accel_data
    int64_t vn_off, vt_off, fn_off, ft_off; // byte offsets into SceneView::mesh_data (attributes of the final hit)
    int64_t wide_off;                       // WideNode array (accel_data)
    uint32_t wide_root, wide_root_aux;      // word / aux of the root (a leaf group when the mesh has <= 8 triangles)
    uint32_t wide_stack, wide_nodes;        // worst-case deferred entries of a traversal (0: no wide tree), node count
    int64_t groups_off;                     // int32 group_end[n_groups] in SceneView::mesh_data (bhrt.h: bhrt_mesh::n_groups)
    int32_t n_groups, pad_;
};

// ------------------------------------------------------------------------------------------
// Device-side view of an uploaded scene.  Pointers are device pointers in render.cu and
// host pointers in tests/emu.
struct SceneView {
    int n_nodes, n_lights, max_depth, n_materials;
    const bhrt_node *nodes;
    const bhrt_material *materials;
    const bhrt_light *lights;           // sorted ascending by grey intensity (Main.cpp:116-118)
    const bhrt_texmap *texmaps;
    const bhrt_texture *textures;
    const uint8_t *texels;
    const bhrt_mesh *meshes;
    const uint8_t *mesh_data;
    const MeshAccel *accels;            // one per mesh
    const uint8_t *accel_data;
    int mesh_nodes[BHRT_MESH_MASK_BITS]; // node indices of the first 32 mesh nodes in DFS order
    int use_wide;                       // every mesh has a wide tree (accel.hpp); CPU harness: traverse it (--wide)
    int n_mesh_nodes;                   // all mesh nodes of the scene graph
    const float *tex_offsets;           // 31 (x,y) pairs of the elliptical Halton filter (scene.h:322-329)
    bhrt_texcolor background, environment;
    float all_light_intensity;          // Main.cpp:119-122
    V3 cam_pos, top_left, dd_x, dd_y;   // Main.cpp:179-192, computed on the host
    int width, height;
    int face_materials;                 // BHRT_FLAG_FACE_MATERIALS of the frame being rendered
};

struct RayCounters { unsigned long long closest, shadow; unsigned overflow = 0; };   // overflow: Shade() levels refused for lack of frames

// The reference's 24-byte photon (cyPhotonMap.h:72-80): position, max-channel power, 8-bit
// colour ratio, kd-tree plane bits + sign of dir.z, 16-bit fixed-point dir.x / dir.y.
struct PhotonRecord {
    float pos[3];
    float power;
    uint8_t rgb[3];
    uint8_t planeAndDirZ;
    int16_t dirX, dirY;
};

// Gather-side photon: 32 bytes, two 128-bit loads.  a = (position, max-channel power);
// b = (direction as Photon::GetDirection decodes it -- including its dirX*dirX + dirY - dirY
// slip, cyPhotonMap.h:196 -- and the 8-bit colour ratio in the bits of b.w).
struct alignas(16) GatherRec { F4 a, b; };

// The coarse cell is HALF the gather radius.  Search level 0 looks at the 5x5x5 block of coarse cells (which covers
// the whole radius: bound = r) -- on a surface that is 6.25 r^2 of photons instead of the 9 r^2 of a 3x3x3 block of
// radius-sized cells, and rows of it beyond the radius are skipped; 88 % of config 5's queries end at this level.
// Search level l >= 1 is the 3x3x3 block at grid resolution l - 1 (cell, cell/2, cell/4, cell/8 = r/16) with
// bound = that cell edge.  (Two more resolutions were measured in round 1: 3x SLOWER on config 5.)
#define BHRT_PM_SUB_BITS 3                          // octree levels below a coarse cell
#define BHRT_PM_LEVELS (BHRT_PM_SUB_BITS + 2)       // search levels: 5x5x5 @ cell, then 3x3x3 @ cell, cell/2, cell/4, cell/8
#define BHRT_PM_FINE_BITS (3 * BHRT_PM_SUB_BITS)    // Morton bits of the finest sub-cell
#define BHRT_PM_RES(level) ((level) > 0 ? (level) - 1 : 0)      // grid resolution of a search level
// Read-only view of the gather structure (photon_build.hpp / photon.cuh build it): photons sorted
// by key = coarse cell << 12 | Morton code of the 1/16 sub-cell.  Every cell of every level is a
// contiguous key range, so ONE sorted array serves all five resolutions; a query picks the level
// at which the k nearest photons fit in its 3x3x3 block (cyPhotonMap's kd-tree adapts to the
// density the same way, cyPhotonMap.h:421-498).
struct PhotonGather {
    int use;                    // 0 = no caustic term (reference built without USE_PhotonMap)
    int k;                      // MAX_PhotonCountInArea (MtlBlinn.cpp:28)
    float radius;               // MAX_Area (MtlBlinn.cpp:29)
    int n_photons;
    const GatherRec *recs;      // sorted by key
    const uint32_t *keys;       // sorted keys
    const uint32_t *coarse_start;   // [dim.x*dim.y*dim.z + 1]
    float origin[3];
    float cell;                 // coarse cell edge (>= radius)
    int dim[3];
};

// Photon::SetDirection / SetPower (cyPhotonMap.h:172-191) and the Color24 conversion (cyColor.h:271)
BHRT_HD void photon_pack(PhotonRecord &r, V3 pos, V3 dir, V3 power)
{
    r.pos[0] = pos.x; r.pos[1] = pos.y; r.pos[2] = pos.z;
    r.dirX = (int16_t)(dir.x * 0x7FFF);
    r.dirY = (int16_t)(dir.y * 0x7FFF);
    r.planeAndDirZ = dir.z > 0 ? 0 : 0x8;
    float pw = power.x;
    if (pw < power.y) pw = power.y;
    if (pw < power.z) pw = power.z;
    r.power = pw;
    V3 c = power / pw;
    int v;
    v = (int)(c.x * 255 + 0.5f); r.rgb[0] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    v = (int)(c.y * 255 + 0.5f); r.rgb[1] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    v = (int)(c.z * 255 + 0.5f); r.rgb[2] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
// Photon::GetDirection (cyPhotonMap.h:193-214), integer square root and all
BHRT_HD V3 photon_direction(const PhotonRecord &r)
{
    V3 dir;
    dir.x = float(r.dirX) / float(0x7FFF);
    dir.y = float(r.dirY) / float(0x7FFF);
    int dirXY2 = r.dirX * r.dirX + r.dirY - r.dirY;
    if (dirXY2 > 0x3FFF0001) dirXY2 = 0x3FFF0001;
    int dirZ2 = 0x3FFF0001 - dirXY2;
    int dirZ = 0;
    int place = 0x40000000;
    int remainder = dirZ2;
    while (place > remainder) place = place >> 2;
    while (place) {
        if (remainder >= dirZ + place) {
            remainder = remainder - dirZ - place;
            dirZ = dirZ + (place << 1);
        }
        dirZ = dirZ >> 1;
        place = place >> 2;
    }
    dir.z = float(dirZ) / float(0x7FFF);
    if (r.planeAndDirZ & 0x8) dir.z = -dir.z;
    return dir;
}
BHRT_HD GatherRec photon_gather_rec(const PhotonRecord &r, float power_scale)
{
    GatherRec g;
    V3 d = photon_direction(r);
    g.a.x = r.pos[0]; g.a.y = r.pos[1]; g.a.z = r.pos[2]; g.a.w = r.power * power_scale;      // ScalePhotonPowers (Main.cpp:380)
    g.b.x = d.x; g.b.y = d.y; g.b.z = d.z;
    uint32_t rgb = (uint32_t)r.rgb[0] | ((uint32_t)r.rgb[1] << 8) | ((uint32_t)r.rgb[2] << 16);
    memcpy(&g.b.w, &rgb, 4);
    return g;
}
// Morton code (x,y,z interleaved, most significant level first) of a BHRT_PM_SUB_BITS-bit sub-cell coordinate
BHRT_HD uint32_t photon_morton(uint32_t x, uint32_t y, uint32_t z)
{
    uint32_t m = 0;
    for (int j = BHRT_PM_SUB_BITS - 1; j >= 0; j--) m = (m << 3) | (((x >> j) & 1u) << 2) | (((y >> j) & 1u) << 1) | ((z >> j) & 1u);
    return m;
}
// Sort key of a position (shared by the host and the device builder); false = outside the grid
BHRT_HD bool photon_key(const float *origin, float cell, const int *dim, V3 p, uint32_t &key)
{
    float fx = (p.x - origin[0]) / cell, fy = (p.y - origin[1]) / cell, fz = (p.z - origin[2]) / cell;
    int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= dim[0] || cy >= dim[1] || cz >= dim[2]) return false;
    const float sc = (float)(1 << BHRT_PM_SUB_BITS);
    const int smax = (1 << BHRT_PM_SUB_BITS) - 1;
    int sx = (int)((fx - cx) * sc), sy = (int)((fy - cy) * sc), sz = (int)((fz - cz) * sc);
    sx = sx < 0 ? 0 : (sx > smax ? smax : sx); sy = sy < 0 ? 0 : (sy > smax ? smax : sy); sz = sz < 0 ? 0 : (sz > smax ? smax : sz);
    uint32_t coarse = (uint32_t)cx + (uint32_t)dim[0] * ((uint32_t)cy + (uint32_t)dim[1] * (uint32_t)cz);
    key = (coarse << BHRT_PM_FINE_BITS) | photon_morton((uint32_t)sx, (uint32_t)sy, (uint32_t)sz);
    return true;
}

// ------------------------------------------------------------------------------------------
// Box::IntersectRay (reference Objects/Box/Box.cpp:3-46)
BHRT_HD bool intersect_box6(float b0, float b1, float b2, float b3, float b4, float b5, const Ray &r, float t_max, float &t_min)
{
    float tz1 = (r.dir.z != 0) ? (b2 - r.p.z) / r.dir.z : BHRT_BIGFLOAT;
    float tz2 = (r.dir.z != 0) ? (b5 - r.p.z) / r.dir.z : -BHRT_BIGFLOAT;
    float ty1 = (r.dir.y != 0) ? (b1 - r.p.y) / r.dir.y : BHRT_BIGFLOAT;
    float ty2 = (r.dir.y != 0) ? (b4 - r.p.y) / r.dir.y : -BHRT_BIGFLOAT;
    float tx1 = (r.dir.x != 0) ? (b0 - r.p.x) / r.dir.x : BHRT_BIGFLOAT;
    float tx2 = (r.dir.x != 0) ? (b3 - r.p.x) / r.dir.x : -BHRT_BIGFLOAT;
    float tMin = fmaxr(fmaxr(fminr(tx1, tx2), fminr(ty1, ty2)), fminr(tz1, tz2));
    float tMax = fminr(fminr(fmaxr(tx1, tx2), fmaxr(ty1, ty2)), fmaxr(tz1, tz2));
    if (tMin <= tMax && tMin < t_max) { t_min = tMin; return true; }
    return false;
}
BHRT_HD bool intersect_box(const float *b, const Ray &r, float t_max, float &t_min)
{
    return intersect_box6(b[0], b[1], b[2], b[3], b[4], b[5], r, t_max, t_min);
}

// ------------------------------------------------------------------------------------------
// Mesh attribute fetch (only for the FINAL closest hit of a ray: the reference interpolates vn/vt
// for every accepted candidate, TriObj.cpp:169-177, and the last accepted one wins -- the same
// value as interpolating once for the winner).
BHRT_HD void mesh_attributes(const SceneView &s, const MeshAccel &A, uint32_t face, float a0, float a1, float a2, V3 &N, V3 &tc)
{
    const float *vn = (const float *)(s.mesh_data + A.vn_off);
    const float *vt = (const float *)(s.mesh_data + A.vt_off);
    const uint32_t *fn = (const uint32_t *)(s.mesh_data + A.fn_off);
    const uint32_t *ft = (const uint32_t *)(s.mesh_data + A.ft_off);
    float a = a0 + a1 + a2;
    V3 bc(a0 / a, a1 / a, a2 / a);
    // TriMesh::Interpolate (cyTriMesh.h:191): v[f.v[0]]*bc.x + v[f.v[1]]*bc.y + v[f.v[2]]*bc.z
    uint32_t n0 = fn[3 * face], n1 = fn[3 * face + 1], n2 = fn[3 * face + 2];
    N = ld3(vn + 3 * n0) * bc.x + ld3(vn + 3 * n1) * bc.y + ld3(vn + 3 * n2) * bc.z;
    uint32_t t0 = ft[3 * face], t1 = ft[3 * face + 1], t2 = ft[3 * face + 2];
    tc = ld3(vt + 3 * t0) * bc.x + ld3(vt + 3 * t1) * bc.y + ld3(vt + 3 * t2) * bc.z;
}

// Candidate hit inside a mesh: enough to rebuild the HitInfo of TriObj::IntersectTriangle
struct MeshHit { uint32_t face; float a0, a1, a2; };

// TriObj::IntersectTriangle (reference Objects/TriObj/TriObj.cpp:68-160) on a precomputed
// TriRec.  `z` is hInfo.z (in/out); `len_dir` = |ray.dir| (hoisted; same value every call).
BHRT_HD bool intersect_trirec(const TriRec *rec, const Ray &ray, float len_dir, float &z, bool &front_out, int hitSide, MeshHit &mh)
{
    F4 A = ld4(&rec->a), B = ld4(&rec->b), C = ld4(&rec->c);
    V3 vN(A.x, A.y, A.z);
    float t_div = dot(vN, ray.dir);
    if (t_div == 0) return false;
    float perp = t_div / (A.w * len_dir);
    if (perp > -BHRT_PERP_DET && perp < BHRT_PERP_DET) return false;
    float t = (B.x - dot(vN, ray.p)) / t_div;
    if (t <= 0 || t > z) return false;
    bool front = t_div < 0;
    if (!front && hitSide == HIT_FRONT) return false;
    else if (front && hitSide == HIT_BACK) return false;
    V3 vX = ray.p + t * ray.dir;
    uint32_t tag = f2u(C.w);
    uint32_t axis = tag >> 30;
    float pxx = axis == 0 ? vX.y : vX.x;
    float pxy = axis == 2 ? vX.y : vX.z;
    float p0x = B.y, p0y = B.z, p1x = B.w, p1y = C.x, p2x = C.y, p2y = C.z;
    // Vec2::Cross(a,b) = a.x*b.y - a.y*b.x (cyVector.h)
    float a0 = ((p1x - pxx) * (p2y - pxy) - (p1y - pxy) * (p2x - pxx)) / 2.f;
    float a1 = ((p2x - pxx) * (p0y - pxy) - (p2y - pxy) * (p0x - pxx)) / 2.f;
    float a2 = ((p0x - pxx) * (p1y - pxy) - (p0y - pxy) * (p1x - pxx)) / 2.f;
    if ((a2 < 0 || a1 < 0 || a0 < 0) && !(a0 < 0 && a1 < 0 && a2 < 0)) return false;
    z = t; front_out = front;
    mh.face = tag & 0x3fffffffu; mh.a0 = a0; mh.a1 = a1; mh.a2 = a2;
    return true;
}

// Both child boxes of one internal node (Box::IntersectRay x2, reference Box.cpp:3-46)
BHRT_HD void intersect_pair(const PairNode *pn, const Ray &r, float t_max, bool &h1, float &t1, bool &h2, float &t2, uint32_t &w1, uint32_t &w2)
{
    const F4 *q = reinterpret_cast<const F4 *>(pn);
    F4 q0 = ld4(q), q1 = ld4(q + 1), q2 = ld4(q + 2), q3 = ld4(q + 3);
    t1 = BHRT_BIGFLOAT; t2 = BHRT_BIGFLOAT;
    h1 = intersect_box6(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, r, t_max, t1);
    h2 = intersect_box6(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, r, t_max, t2);
    w1 = f2u(q3.x); w2 = f2u(q3.y);
}

#define BHRT_BVH_LEAF(d) (((d) & 0x80000000u) != 0)
#define BHRT_BVH_COUNT(d) ((((d) >> 28) & 0x7u) + 1)
#define BHRT_BVH_OFFSET(d) ((d) & 0x0fffffffu)
#define BHRT_BVH_CHILD(d) ((d) & 0x7fffffffu)
#ifndef BHRT_BVH_STACK
#define BHRT_BVH_STACK 64
#endif

// One traversal for both query kinds (the two heavy parts -- pair slab tests and triangle tests --
// are shared, so a warp whose lanes mix closest-hit and shadow queries stays converged), written
// as a STEPPER (one BVH node per step) so that the mesh kernel can hand a finished lane its next
// ray while its warp neighbours are still traversing:
//
//  closest (any_hit == false): TriObj::IntersectRay + TraceBVHNode (reference TriObj.cpp:17-39,
//    192-270) made iterative.  The visiting order, the near/far rule, "the far child is skipped
//    only when the near child produced a hit and z <= tmin_far" and the traversal of box-missed
//    far children are all kept, so ties and rounding cases resolve as in the reference.
//  shadow (any_hit == true): TriObj::ShadowRecursive + TraceBVHShadow (TriObj.cpp:41-66,272-307):
//    child 1 before child 2, both visited when EITHER box is hit (boxes tested against BIGFLOAT),
//    stop at the FIRST leaf containing any front-face hit and range-check only that leaf's t.
//
// Result: closest -> number of accepted triangles > 0 (z, front, mh updated);
//         shadow  -> the reference's bool.
// (the scalars and the deferred-children stack are separate objects on purpose: the traversal phase of the render
// kernel keeps a COPY of the scalars in registers while it steps -- as members of one struct next to dynamically
// indexed arrays the compiler has to assume every push may have changed them and re-reads the ray from local memory on
// every step, which was 37 % of the kernel's local-memory traffic, profiles/r2_cfg2_wf_render_ncu_summary.json)
struct MeshTravHot {        // what EVERY BVH step reads and writes: 44 bytes = 11 words (odd: one per thread in shared memory is conflict-free)
    Ray ray;
    float z;                // running closest hit distance (in/out)
    uint32_t cur;
    uint16_t sp, hits;      // deferred children on the stack; triangles accepted so far
    uint32_t pairs_off;     // the mesh's PairNode array: byte offset into SceneView::accel_data
    bool any_hit, done, result, front;
};
static_assert(sizeof(MeshTravHot) == 44, "MeshTravHot: 11 words");
struct MeshTravCold {       // leaves and accepted hits only
    float len_dir;
    uint32_t tris_off;      // the mesh's TriRec array (accel_data)
    MeshHit mh;             // the winning triangle (in/out)
    float t_max_shadow;
    int hitSide;
#ifndef __CUDACC__
    const WideNode *wide; uint32_t wide_root, wide_root_aux;    // CPU harness only (wide_trav_run)
#endif
};
struct MeshTravStack {
    uint32_t st_word[BHRT_BVH_STACK];
    float st_tmin[BHRT_BVH_STACK];
#if !BHRT_FAST
    int st_hits[BHRT_BVH_STACK];        // (fast build: only box-hit children are deferred, see mesh_trav_pop)
#endif
};
// the three parts of one traversal.  They are separate objects on purpose: the render kernel keeps the hot part of every
// thread in SHARED memory (wavefront.cuh) instead of local memory, where its dozen reloads per step competed with a
// thousand threads' lane copies and stacks for L1.  MEASURED: the node step became pure LDS, the traversal phase 4 %
// shorter, the lanes pass 2 % longer for the 44 KB of L1 -- neutral on BASELINE config 2.
struct MeshTrav { MeshTravHot hot; MeshTravCold cold; MeshTravStack stk; };

BHRT_HD void mesh_trav_begin(MeshTravHot &T, MeshTravCold &C, const SceneView &s, const MeshAccel &A, const Ray &ray, bool any_hit, float t_max_shadow,
                             float z, bool front, int hitSide)
{
    T.ray = ray; T.any_hit = any_hit; C.t_max_shadow = t_max_shadow; T.z = z; T.front = front; C.hitSide = hitSide;
    C.mh.face = 0; C.mh.a0 = C.mh.a1 = C.mh.a2 = 0;
    T.sp = 0; T.hits = 0; T.result = false;
    float tmin = -1;
    if (!intersect_box(A.root_box, ray, any_hit ? BHRT_BIGFLOAT : z, tmin)) { T.done = true; return; }
    T.pairs_off = (uint32_t)A.pairs_off;
    C.tris_off = (uint32_t)A.tris_off;
#ifndef __CUDACC__
    C.wide = (const WideNode *)(s.accel_data + A.wide_off); C.wide_root = A.wide_root; C.wide_root_aux = A.wide_root_aux;
#endif
    C.len_dir = length(ray.dir);
    T.cur = A.root_word;
    T.done = false;
}

#ifndef BHRT_TRAV_HOOK
#define BHRT_TRAV_HOOK(is_leaf, is_shadow)
#endif
// after a leaf, or an internal node neither of whose boxes was hit: resume at the deferred child
BHRT_HD void mesh_trav_pop(MeshTravHot &T, MeshTravStack &S)
{
    while (T.sp > 0) {
        T.sp--;
        if (!T.any_hit) {
#if BHRT_FAST
            // only children whose box was hit (tmin < z at that time) are deferred in this build, and z can only
            // have shrunk through a hit in the near child since: the reference's "near child produced a hit and
            // z <= tmin_far" is then the plain distance cull, and no hit count needs to be remembered
            if (!(T.z > S.st_tmin[T.sp])) continue;
#else
            bool nearHit = ((int)T.hits != S.st_hits[T.sp]);
            if (nearHit && !(T.z > S.st_tmin[T.sp])) continue;
#endif
        }
        T.cur = S.st_word[T.sp];
        return;
    }
    T.result = T.hits > 0;
    T.done = true;
}
// T.cur is an internal node: both child boxes, near child first.  accel = SceneView::accel_data
BHRT_HD void mesh_trav_internal(MeshTravHot &T, MeshTravStack &S, const uint8_t *accel)
{
    BHRT_TRAV_HOOK(false, T.any_hit);
    bool h1, h2; float t1, t2; uint32_t w1, w2;
    intersect_pair((const PairNode *)(accel + T.pairs_off) + T.cur, T.ray, T.any_hit ? BHRT_BIGFLOAT : T.z, h1, t1, h2, t2, w1, w2);
    if (h1 || h2) {
        uint32_t first, second; float tsecond;
        if (T.any_hit || t1 < t2) { first = w1; second = w2; tsecond = t2; }
        else { first = w2; second = w1; tsecond = t1; }
#if BHRT_FAST
        // The reference descends into BOTH children as soon as one box is hit (TriObj.cpp:221-223): the
        // missed child costs it a leaf's triangle tests or one more pair of (missed) boxes and can only
        // produce a hit through rounding.  The fast build (stochastic mode) does not visit it: a third
        // fewer node steps on BASELINE config 2, same image statistics.
        if (!(h1 && h2)) { T.cur = h1 ? w1 : w2; return; }
#endif
        if (T.sp < BHRT_BVH_STACK) {
            S.st_word[T.sp] = second; S.st_tmin[T.sp] = tsecond;
#if !BHRT_FAST
            S.st_hits[T.sp] = (int)T.hits;
#endif
            T.sp++;
        }
        T.cur = first;
        return;
    }
    mesh_trav_pop(T, S);
}
// T.cur is a leaf: its triangles
BHRT_HD void mesh_trav_leaf(MeshTravHot &T, MeshTravCold &C, MeshTravStack &S, const uint8_t *accel)
{
    BHRT_TRAV_HOOK(true, T.any_hit);
    const uint32_t cur = T.cur;
    const uint32_t cnt = BHRT_BVH_COUNT(cur), off = BHRT_BVH_OFFSET(cur);
    const TriRec *tris = (const TriRec *)(accel + C.tris_off) + off;
    if (T.any_hit) {
        float lz = BHRT_BIGFLOAT; bool lf; MeshHit lm;
        bool any = false;
        for (uint32_t i = 0; i < cnt; i++)
            if (intersect_trirec(tris + i, T.ray, C.len_dir, lz, lf, HIT_FRONT, lm)) any = true;
        if (any) { T.result = lz > BHRT_BIAS && lz < C.t_max_shadow; T.done = true; return; }
    } else {
        for (uint32_t i = 0; i < cnt; i++)
            if (intersect_trirec(tris + i, T.ray, C.len_dir, T.z, T.front, C.hitSide, C.mh)) T.hits++;
    }
    mesh_trav_pop(T, S);
}
BHRT_HD void mesh_trav_step(MeshTrav &M, const uint8_t *accel)
{
    if (BHRT_BVH_LEAF(M.hot.cur)) mesh_trav_leaf(M.hot, M.cold, M.stk, accel); else mesh_trav_internal(M.hot, M.stk, accel);
}

#ifndef __CUDACC__
// The traversal of the WIDE tree, stated for one thread (CPU harness, `emu_render --wide`): this is the
// definition the kernels' 8-lanes-per-ray version (wide.cuh) evaluates in parallel -- lane j does the j-th
// iteration of each `for (j ...)` loop below and ballots / shuffles replace the scalar bookkeeping.
//   closest: children whose box is hit (tmin < z) are visited nearest first (ties: lower slot), deferred ones
//            are dropped when tmin >= z at the time they are popped; a leaf group's records are all tested
//            against the z the group was entered with and the nearest accepted one wins (ties: the later
//            record, like TriObj::IntersectTriangle's `t > z` rejection met them in order);
//   shadow:  children whose box is hit are visited in SLOT order, i.e. the reference's leaves in the
//            reference's order; the first of the reference's leaves that holds a front-face hit decides
//            (TraceBVHShadow, TriObj.cpp:272-291).  Like the fast build of the binary traversal, a child whose
//            box the ray misses is not visited.
inline void wide_trav_run(MeshTravHot &T, MeshTravCold &C, const uint8_t *accel)
{
    const TriRec *T_tris = (const TriRec *)(accel + C.tris_off);
    uint32_t st_w[BHRT_WIDE_STACK]; float st_t[BHRT_WIDE_STACK]; uint32_t st_a[BHRT_WIDE_STACK];
    int sp = 0;
    uint32_t cur = C.wide_root, aux = C.wide_root_aux;
    for (;;) {
        if (BHRT_BVH_LEAF(cur)) {
            const uint32_t cnt = BHRT_BVH_COUNT(cur), off = BHRT_BVH_OFFSET(cur);
            bool hit[8]; float t[8]; bool fr[8]; MeshHit mh[8];
            for (uint32_t j = 0; j < cnt; j++) {
                t[j] = T.any_hit ? BHRT_BIGFLOAT : T.z;
                hit[j] = intersect_trirec(T_tris + off + j, T.ray, C.len_dir, t[j], fr[j], T.any_hit ? (int)HIT_FRONT : C.hitSide, mh[j]);
            }
            if (T.any_hit) {
                int f = -1;
                for (uint32_t j = 0; j < cnt && f < 0; j++) if (hit[j]) f = (int)j;
                if (f >= 0) {
                    int b = f; while (b > 0 && !((aux >> b) & 1u)) b--;
                    uint32_t e = (uint32_t)f + 1; while (e < cnt && !((aux >> e) & 1u)) e++;
                    float lz = BHRT_BIGFLOAT;
                    for (uint32_t j = (uint32_t)b; j < e; j++) if (hit[j] && t[j] < lz) lz = t[j];
                    T.result = lz > BHRT_BIAS && lz < C.t_max_shadow; T.done = true; return;
                }
            } else {
                int w = -1;
                for (uint32_t j = 0; j < cnt; j++) if (hit[j] && (w < 0 || t[j] <= t[w])) w = (int)j;
                if (w >= 0) { T.z = t[w]; T.front = fr[w]; C.mh = mh[w]; T.hits++; }
            }
        } else {
            const WideNode &nd = C.wide[cur];
            int order[8], n = 0; float tm[8];
            for (int j = 0; j < 8; j++) {
                if (nd.c[j].word == BHRT_WIDE_EMPTY) continue;
                float tmin = -1;
                if (!intersect_box(nd.c[j].box, T.ray, T.any_hit ? BHRT_BIGFLOAT : T.z, tmin)) continue;
                tm[j] = tmin;
                int k = n++;
                if (!T.any_hit) while (k > 0 && tm[order[k - 1]] > tmin) { order[k] = order[k - 1]; k--; }
                order[k] = j;
            }
            if (n > 0) {
                for (int k = n - 1; k >= 1; k--) { st_w[sp] = nd.c[order[k]].word; st_t[sp] = tm[order[k]]; st_a[sp] = nd.c[order[k]].aux; sp++; }
                cur = nd.c[order[0]].word; aux = nd.c[order[0]].aux;
                continue;
            }
        }
        bool found = false;
        while (sp > 0) {
            sp--;
            if (!T.any_hit && !(st_t[sp] < T.z)) continue;
            cur = st_w[sp]; aux = st_a[sp]; found = true;
            break;
        }
        if (!found) { T.result = T.hits > 0; T.done = true; return; }
    }
}
#endif

// ------------------------------------------------------------------------------------------
// A ray query: the ONE place the estimator meets the geometry.  kind CLOSEST = recursive()
// (reference Main.cpp:389-413, linear DFS over the scene graph); kind SHADOW = GenLight::Shadow /
// ShadowRayRecursive (reference Lights/GenLight.cpp:10-69, any-hit; the result is the OR over
// all nodes of the node's own test, so DFS pre-order is as good as the reference's post-order).
enum { Q_NONE = 0, Q_CLOSEST = 1, Q_SHADOW = 2 };
struct alignas(16) Query {
    Ray ray;
    int kind;
    int hitSide;        // CLOSEST
    float t_max;        // SHADOW
};

// The query is answered in three phases so that the wavefront can give each its own, converged,
// kernel (wavefront.cuh); scene_trace() below simply runs them back to back.
//
//  1 scene_trace_graph   the node loop: ray transform + analytic object per node (warp-uniform
//                        control flow: every lane visits node k at the same time).  A mesh node is
//                        only tested against its root box; hits are recorded in `mesh_mask` (bit i
//                        = i-th mesh node of the DFS order, SceneView::mesh_nodes).
//  2 scene_trace_meshes  BVH traversal of the recorded mesh nodes -- only lanes that need it.
//  3 scene_trace_finish  winner-only work: sphere uv (atan2/asin), plane ray differentials (pow),
//                        mesh normal/uv interpolation, Node::FromNodeCoords.
//
// Equivalence with recursive() (Main.cpp:389-413), which tests the nodes strictly in DFS order
// with one running HitInfo: every candidate's t is computed by the same expression whatever the
// running z is (z only rejects), so the winner is the candidate with the smallest t and, between
// equal t, the LATER node (the reference rejects only t > z).  Phase 2 therefore starts each mesh
// with the best z found so far and applies that tie rule explicitly (`strict`).
// nodes[] is in DFS pre-order; rs[d] is the ray in the space of the current node at depth d.
// Shadow quirks kept: sphere ignored when the origin is inside (GenLight.cpp:36-37), plane
// bounds tested with the PARENT-space ray (:54), no rayDz == 0 test (:50).
// returns true when a SHADOW query is already answered "blocked"
BHRT_HD bool scene_trace_graph(const SceneView &s, const Query &q, Hit &h, uint32_t &mesh_mask, Ray *first_mesh_ray = 0)
{
    const bool shadow = (q.kind == Q_SHADOW);
    // the rays in the space of the current ancestors: depths 1 and 2 (every repository scene) stay in
    // registers, deeper graphs use the array
    Ray rs[BHRT_MAX_DEPTH + 1];
    Ray r1 = q.ray, r2 = q.ray;
    h.init();
    mesh_mask = 0;
    int mesh_ordinal = 0;
    for (int k = 0; k < s.n_nodes; k++) {
        const bhrt_node &nd = s.nodes[k];
        const int d = nd.depth;
        const Ray pr = d == 1 ? q.ray : (d == 2 ? r1 : (d == 3 ? r2 : rs[d - 1]));
        Ray r = to_node(nd.xf, pr);
        if (d == 1) r1 = r; else if (d == 2) r2 = r; else if (d < s.max_depth) rs[d] = r;
        bool hit = false;
        if (nd.kind == BHRT_OBJ_SPHERE) {
            // Sphere::IntersectRay (Sphere.cpp:8-49) / the inline quadratic of GenLight.cpp:24-43
            float A = dot(r.dir, r.dir);
            float B = 2 * dot(r.dir, r.p);
            float C = dot(r.p, r.p) - 1;
            float DD = B * B - 4 * A * C;
            if (DD > 0) {
                float sq = sqrtf(DD);
                float t1 = (-B + sq) / (2 * A);
                float t2 = (-B - sq) / (2 * A);
                if (shadow) {
                    float t = fminr(t1, t2);
                    if (!(t < 0) && t < q.t_max && t > BHRT_SHADOW_BIAS) return true;
                } else {
                    float t = BHRT_BIGFLOAT;
                    bool front = true, cand = true;
                    if (t1 < 0 && t2 < 0) cand = false;
                    else if (t1 * t2 <= 0) {
                        if (q.hitSide == HIT_FRONT) cand = false;
                        t = t1; front = false;
                    } else if (t1 > 0 && t2 > 0) {
                        if (q.hitSide == HIT_FRONT || q.hitSide == HIT_FRONT_AND_BACK) { t = t2; front = true; }
                        else if (q.hitSide == HIT_BACK) { t = t1; front = false; }
                    }
                    if (cand && !(h.z < t || t <= 0)) { h.z = t; h.front = front; hit = true; }
                }
            }
        } else if (nd.kind == BHRT_OBJ_PLANE) {
            if (shadow) {
                float t = -r.p.z / r.dir.z;
                if (!(t < 0)) {
                    V3 x = pr.p + t * pr.dir;
                    if (!(x.x < -1 || x.x > 1 || x.y < -1 || x.y > 1) && t < q.t_max && t > BHRT_SHADOW_BIAS) return true;
                }
            } else if (r.dir.z != 0.0f) {
                // Plane::IntersectRay (Plane.cpp:8-49)
                float t = -r.p.z / r.dir.z;
                if (!(t <= 0 || t > h.z)) {
                    V3 x = r.p + t * r.dir;
                    if (!(x.x < -1 || x.x > 1 || x.y < -1 || x.y > 1)) {
                        bool front = (dot(-r.dir, V3(0, 0, 1)) > 0);
                        if (!((!front && q.hitSide == HIT_FRONT) || (front && q.hitSide == HIT_BACK))) { h.z = t; h.front = front; hit = true; }
                    }
                }
            }
        } else if (nd.kind == BHRT_OBJ_MESH) {
            const MeshAccel &A = s.accels[nd.mesh];
            // only the root box here; nodes beyond the mask are picked up by scene_trace_meshes
            if (mesh_ordinal < BHRT_MESH_MASK_BITS) {
                float tmin = -1;
                if (intersect_box(A.root_box, r, shadow ? BHRT_BIGFLOAT : h.z, tmin)) {
                    if (first_mesh_ray && !mesh_mask) *first_mesh_ray = r;      // the ray in the space of the first mesh it entered
                    mesh_mask |= 1u << mesh_ordinal;
                }
            }
            mesh_ordinal++;
        }
        if (hit) h.node = k;
    }
    return false;
}

// The ray in the local space of node k: Node::ToNodeCoords down the ancestor chain (what rs[] held)
BHRT_HD Ray node_local_ray(const SceneView &s, int k, const Ray &world)
{
    int chain[BHRT_MAX_DEPTH + 1];
    int n = 0;
    for (int i = k; i >= 0 && n <= BHRT_MAX_DEPTH; i = s.nodes[i].parent) chain[n++] = i;
    Ray r = world;
    for (int i = n - 1; i >= 0; i--) r = to_node(s.nodes[chain[i]].xf, r);
    return r;
}

BHRT_HD float prev_float(float z)     // the next float below a positive finite z
{
    uint32_t u; memcpy(&u, &z, 4); u -= 1u;
    float r; memcpy(&r, &u, 4);
    return r;
}

// Phase 2 as a stepper: one BVH node (or one "pick the next mesh node" action) per step, so the
// mesh kernel can refill a finished lane while its warp neighbours still traverse.  On completion
// `blocked` answers a SHADOW query; for CLOSEST h holds z, front, node and the winning triangle in
// h.uvw (a0,a1,a2) / h.duvw0.x (face bits) for phase 3.
struct MeshJob {
    Query q;
    Hit h;
    uint32_t mask;
    int extra;              // next node index to scan for mesh nodes beyond the mask
    int k;                  // mesh node being traversed
    bool in_trav, strict, blocked;
    MeshTravCold C;         // (the hot part of the traversal is the caller's: shared memory in the render kernel)
    MeshTravStack stk;
};

BHRT_HD void mesh_job_begin(MeshJob &J, const SceneView &s, const Query &q, float z, int node, bool front, uint32_t mask)
{
    J.q = q;
    J.h.init();
    J.h.z = z; J.h.node = node; J.h.front = front;
    J.mask = mask;
    // mesh nodes beyond the 32 the mask covers (s.n_mesh_nodes > 32) are visited unconditionally:
    // mesh_trav_begin tests the root box itself
    J.extra = s.n_mesh_nodes > BHRT_MESH_MASK_BITS ? s.mesh_nodes[BHRT_MESH_MASK_BITS - 1] + 1 : s.n_nodes;
    J.in_trav = false; J.blocked = false; J.strict = false; J.k = -1;
}

// Pick the next mesh node the ray has to visit and start its traversal (root box included).
// false = no mesh node left: the job is finished.
BHRT_HD bool mesh_job_next(MeshJob &J, MeshTravHot &T, const SceneView &s)
{
    const bool shadow = (J.q.kind == Q_SHADOW);
    int k;
    if (J.mask) {
        int bit = 0;
        while (!((J.mask >> bit) & 1u)) bit++;
        J.mask &= J.mask - 1u;
        k = s.mesh_nodes[bit];
    } else {
        while (J.extra < s.n_nodes && s.nodes[J.extra].kind != BHRT_OBJ_MESH) J.extra++;
        if (J.extra >= s.n_nodes) return false;
        k = J.extra++;
    }
    J.k = k;
    const bhrt_node &nd = s.nodes[k];
    Ray r = node_local_ray(s, k, J.q.ray);
    if (shadow) mesh_trav_begin(T, J.C, s, s.accels[nd.mesh], r, true, J.q.t_max, BHRT_BIGFLOAT, true, J.q.hitSide);
    else {
        // an equal-t candidate of an EARLIER node loses against the current winner
        J.strict = J.h.node > k;
        mesh_trav_begin(T, J.C, s, s.accels[nd.mesh], r, false, 0.f, J.strict ? prev_float(J.h.z) : J.h.z, J.h.front, J.q.hitSide);
    }
    J.in_trav = true;
    return true;
}
// Collect a finished traversal (J.T.done).  true = the job is finished (shadow ray blocked).
BHRT_HD bool mesh_job_harvest(MeshJob &J, const MeshTravHot &T)
{
    J.in_trav = false;
    if (T.result) {
        if (J.q.kind == Q_SHADOW) { J.blocked = true; return true; }
        J.h.z = T.z; J.h.front = T.front; J.h.node = J.k;
        J.h.uvw = V3(J.C.mh.a0, J.C.mh.a1, J.C.mh.a2); J.h.duvw0.x = u2f(J.C.mh.face);
    }
    return false;
}
// one action per call; returns true when the job is finished
BHRT_HD bool mesh_job_step(MeshJob &J, MeshTravHot &T, const SceneView &s)
{
    if (!J.in_trav) {
        if (!mesh_job_next(J, T, s)) return true;
#ifndef __CUDACC__
        if (s.use_wide && !T.done) wide_trav_run(T, J.C, s.accel_data);
#endif
    }
    else if (BHRT_BVH_LEAF(T.cur)) mesh_trav_leaf(T, J.C, J.stk, s.accel_data);
    else mesh_trav_internal(T, J.stk, s.accel_data);
    if (T.done) return mesh_job_harvest(J, T);
    return false;
}

BHRT_HD bool scene_trace_meshes(const SceneView &s, const Query &q, Hit &h, uint32_t mesh_mask)
{
    MeshJob J;
    MeshTravHot T;
    mesh_job_begin(J, s, q, h.z, h.node, h.front, mesh_mask);
    while (!mesh_job_step(J, T, s)) {}
    if (J.blocked) return true;
    h = J.h;
    return false;
}

// Winner-only part of the closest hit: complete HitInfo in local space, then FromNodeCoords
// (kept out of line in both builds.  Measured in round 2: inlining it in the fast build -- which lets the compiler
// keep the stepping thread's Lane in registers instead of local memory, since no reference to it escapes any more
// -- LOWERED config 2 from 1570 to 1440 Mrays/s: at the 64-register cap the promoted fields come back as spills,
// 1322 instead of 390 bytes, and the compiler's spill placement is worse than the plain local-memory copy.)
BHRT_FN bool scene_trace_finish(const SceneView &s, const Query &q, Hit &h)
{
    h.sub = 0;
    if (h.node < 0) return false;
    const bhrt_node &nd = s.nodes[h.node];
    const Ray wr = node_local_ray(s, h.node, q.ray);
    V3 lp, lN;
    // uvw / duvw are read by texture lookups only (scene.h:371-380): an untextured material never
    // looks at them, so the atan2/asin/pow work is skipped there (they are defined as 0)
    const bhrt_material &hm = s.materials[nd.material];
    const bool textured = hm.diffuse.texmap >= 0 || hm.specular.texmap >= 0;
    if (!textured && nd.kind != BHRT_OBJ_MESH) {
        lp = wr.p + h.z * wr.dir;
        lN = nd.kind == BHRT_OBJ_SPHERE ? lp : V3(0, 0, 1);
        h.uvw = V3(0, 0, 0); h.duvw0 = V3(0, 0, 0); h.duvw1 = V3(0, 0, 0);
    } else if (nd.kind == BHRT_OBJ_SPHERE) {
        lp = wr.p + h.z * wr.dir;                           // Sphere.cpp:52-54
        lN = lp;
        V3 d = normalized(lN);
        // 0.5f + atan2(d.y,d.x)/(2*PI): float atan2f promoted to double, PI is a double literal
        h.uvw.x = (float)(0.5f + m_atan2(d.y, d.x) / (2 * BHRT_PI_D));
        h.uvw.y = (float)(0.5f - m_asin(d.z) / (BHRT_PI_D));
        // uvw.z keeps whatever the default-constructed Vec3f held in the reference (uninitialised
        // there; never read by 2-D textures).  We define it as 0.
        h.uvw.z = 0.f;
        h.duvw0 = V3(0, 0, 0); h.duvw1 = V3(0, 0, 0);
    } else if (nd.kind == BHRT_OBJ_PLANE) {
        const float t = h.z;
        lp = wr.p + t * wr.dir;
        lN = V3(0, 0, 1);
        h.uvw = V3((1 + lp.x) / 2.f, (1 + lp.y) / 2.f, 0.f);
        // ray differentials from the PRIMARY-ray pixel steps dd_x, dd_y, for every ray (Plane.cpp:3-4,51-70)
        V3 nd_ = normalized(wr.dir);
        float scaled_t = length(t * wr.dir);
        float ndnd = dot(nd_, nd_);
        float den = m_pow(ndnd, 1.5f);
        V3 dDX = (ndnd * s.dd_x - dot(nd_, s.dd_x) * nd_) / den;
        V3 dDY = (ndnd * s.dd_y - dot(nd_, s.dd_y) * nd_) / den;
        float ndN = dot(nd_, lN);
        float d_t_x = -(0 + scaled_t * dot(dDX, lN) / ndN);
        float d_t_y = -(0 + scaled_t * dot(dDY, lN) / ndN);
        // "0 + scaled_t * d + d_t * nd": (0 + Vec) is Vec + 0.0f component-wise
        V3 hx = (scaled_t * dDX + V3(0.f, 0.f, 0.f)) + d_t_x * nd_;
        V3 hy = (scaled_t * dDY + V3(0.f, 0.f, 0.f)) + d_t_y * nd_;
        h.duvw0 = hx / 2.f;
        h.duvw1 = hy / 2.f;
    } else {
        lp = wr.p + h.z * wr.dir;                           // TriObj.cpp:107
        const float a0 = h.uvw.x, a1 = h.uvw.y, a2 = h.uvw.z;
        const uint32_t face = f2u(h.duvw0.x);
        mesh_attributes(s, s.accels[nd.mesh], face, a0, a1, a2, lN, h.uvw);
        if (s.face_materials && (hm.flags & BHRT_MTL_MULTI)) {
            // hInfo.mtlID = the usemtl group of the face (what the reference's MultiMtl::Shade expects and
            // TriObj::IntersectTriangle never sets, materials.h:71 / TriObj.cpp:68-189; cyTriMesh::GetMaterialIndex,
            // cyTriMesh.h:221-227: the first group whose cumulative face count exceeds the face index, none -> 0)
            const MeshAccel &MA = s.accels[nd.mesh];
            const int32_t *group_end = (const int32_t *)(s.mesh_data + MA.groups_off);
            int g = 0;
            while (g < MA.n_groups && (int)face >= group_end[g]) g++;
            if (g < MA.n_groups && g < BHRT_MTL_SUBCOUNT(hm.flags)) h.sub = g;
        }
        h.duvw0 = V3(0, 0, 0); h.duvw1 = V3(0, 0, 0);
    }
    // Node::FromNodeCoords of the hit node (Main.cpp:401) and then of its parent
    // (Main.cpp:407-412; the rootNode's identity transform for depth-1 nodes).  Deeper
    // ancestors are NOT applied by the reference; neither are they here.
    h.p = xf_from(nd.xf, lp);
    h.N = normalized(xf_vec_from(nd.xf, lN));
    if (nd.parent >= 0) {
        const bhrt_node &pn = s.nodes[nd.parent];
        h.p = xf_from(pn.xf, h.p);
        h.N = normalized(xf_vec_from(pn.xf, h.N));
    } else {
        // identity: tm*p+pos and itm^T*N are exact, the second normalisation is not a no-op
        h.p = V3(h.p.x * 1.f + h.p.y * 0.f + h.p.z * 0.f + 0.f, h.p.x * 0.f + h.p.y * 1.f + h.p.z * 0.f + 0.f, h.p.x * 0.f + h.p.y * 0.f + h.p.z * 1.f + 0.f);
        h.N = normalized(V3(1.f * h.N.x + 0.f * h.N.y + 0.f * h.N.z, 0.f * h.N.x + 1.f * h.N.y + 0.f * h.N.z, 0.f * h.N.x + 0.f * h.N.y + 1.f * h.N.z));
    }
    return true;
}

// All three phases for one ray (CPU harness, photon emission)
#ifndef BHRT_QUERY_HOOK
#define BHRT_QUERY_HOOK(kind)
#endif
BHRT_HD bool scene_trace(const SceneView &s, const Query &q, Hit &h)
{
    uint32_t mask;
    if (scene_trace_graph(s, q, h, mask)) { BHRT_QUERY_HOOK(q.kind); return true; }
    const bool blocked = (mask || s.n_mesh_nodes > BHRT_MESH_MASK_BITS) && scene_trace_meshes(s, q, h, mask);
    BHRT_QUERY_HOOK(q.kind);
    if (blocked) return true;
    if (q.kind == Q_SHADOW) return false;
    return scene_trace_finish(s, q, h);
}

// ------------------------------------------------------------------------------------------
// Textures (reference Textures/Texture.cpp:97-136, Scenes/scene.h:310-422)
BHRT_HD V3 tile_clamp(V3 uvw)
{
    V3 u(uvw.x - (int)uvw.x, uvw.y - (int)uvw.y, uvw.z - (int)uvw.z);
    if (u.x < 0) u.x += 1;
    if (u.y < 0) u.y += 1;
    if (u.z < 0) u.z += 1;
    return u;
}
// Color24::ToColor (cyColor.h:236).  In the kernels the texel table is RGBX8 (expanded at upload:
// one aligned 32-bit load per texel instead of three byte loads); the CPU harness reads the host's RGB8.
BHRT_HD V3 texel(const uint8_t *d, int index)
{
#ifdef __CUDA_ARCH__
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(d) + index);
    return V3((float)(v & 255u) / 255.0f, (float)((v >> 8) & 255u) / 255.0f, (float)((v >> 16) & 255u) / 255.0f);
#else
    const uint8_t *p = d + 3 * index;
    return V3(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f);
#endif
}

BHRT_MATH_FN V3 texture_sample(const SceneView &s, const bhrt_texture &t, V3 uvw)
{
    V3 u = tile_clamp(uvw);
    if (t.kind == BHRT_TEX_CHECKER) {
        if (u.x <= 0.5f) return u.y <= 0.5f ? ld3(t.color1) : ld3(t.color2);
        else return u.y <= 0.5f ? ld3(t.color2) : ld3(t.color1);
    }
    int width = t.width, height = t.height;
    if (width + height == 0) return V3(0, 0, 0);
    float x = width * u.x, y = height * u.y;
    int ix = (int)x, iy = (int)y;
    float fx = x - ix, fy = y - iy;
    if (ix < 0) ix -= (ix / width - 1) * width;
    if (ix >= width) ix -= (ix / width) * width;
    int ixp = ix + 1; if (ixp >= width) ixp -= width;
    if (iy < 0) iy -= (iy / height - 1) * height;
    if (iy >= height) iy -= (iy / height) * height;
    int iyp = iy + 1; if (iyp >= height) iyp -= height;
#ifdef __CUDA_ARCH__
    const uint8_t *d = s.texels + (t.texel_offset / 3) * 4;        // RGBX8 on the device (texel_offset counts RGB8 bytes)
#else
    const uint8_t *d = s.texels + t.texel_offset;
#endif
    return texel(d, iy * width + ix) * ((1 - fx) * (1 - fy)) +
           texel(d, iy * width + ixp) * (fx * (1 - fy)) +
           texel(d, iyp * width + ix) * ((1 - fx) * fy) +
           texel(d, iyp * width + ixp) * (fx * fy);
}

// Texture::Sample(uvw, duvw) (scene.h:318-337): 32-tap elliptical Halton filter when duvw != 0
// (out of line in BOTH builds: the estimator reaches it from ten places, and ten inlined copies of the tap loop made the
// fast kernel a third larger -- on BASELINE config 3, 20 % of its stall samples were instruction-cache misses)
BHRT_FN V3 texture_sample_filtered(const SceneView &s, const bhrt_texture &t, V3 uvw, V3 d0, V3 d1)
{
    V3 c = texture_sample(s, t, uvw);
    if (length2(d0) + length2(d1) == 0) return c;
    for (int i = 1; i < BHRT_TEX_SAMPLES; i++) {
        float x = s.tex_offsets[2 * (i - 1)], y = s.tex_offsets[2 * (i - 1) + 1];
        c = c + texture_sample(s, t, uvw + x * d0 + y * d1);
    }
    return c / float(BHRT_TEX_SAMPLES);
}

// TexturedColor::Sample(uvw) (scene.h:410) -> TextureMap::Sample(uvw) (:371)
BHRT_HD V3 texcolor_sample(const SceneView &s, const bhrt_texcolor &tc, V3 uvw)
{
    V3 color = ld3(tc.color);
    if (tc.texmap < 0) return color;
    const bhrt_texmap &m = s.texmaps[tc.texmap];
    if (m.texture < 0) return color * V3(0, 0, 0);
    return color * texture_sample(s, s.textures[m.texture], xf_to(m.xf, uvw));
}
// TexturedColor::Sample(uvw,duvw) (scene.h:411) -> TextureMap::Sample(uvw,duvw) (:372-380)
BHRT_EST_FN V3 texcolor_sample_filtered(const SceneView &s, V3 color, int texmap, V3 uvw, V3 d0, V3 d1)
{
    if (texmap < 0) return color;
    const bhrt_texmap &m = s.texmaps[texmap];
    if (m.texture < 0) return color * V3(0, 0, 0);
    V3 u = xf_to(m.xf, uvw);
    V3 e0 = xf_to(m.xf, d0 + uvw) - u;
    V3 e1 = xf_to(m.xf, d1 + uvw) - u;
    return color * texture_sample_filtered(s, s.textures[m.texture], u, e0, e1);
}
// TexturedColor::SampleEnvironment (scene.h:414-420); dir is NOT normalised by the callers
BHRT_EST_FN V3 sample_environment(const SceneView &s, V3 dir)
{
    float z = m_asin(-dir.z) / 3.14159265358979323846f + 0.5f;
    float x = dir.x / (fabsf(dir.x) + fabsf(dir.y));
    float y = dir.y / (fabsf(dir.x) + fabsf(dir.y));
    V3 uvw = V3(0.5f, 0.5f, 0.0f) + z * (x * V3(0.5f, 0.5f, 0) + y * V3(-0.5f, 0.5f, 0));
    return texcolor_sample(s, s.environment, uvw);
}

// ------------------------------------------------------------------------------------------
// Counter-based RNG.  The reference draws from libc rand(); parity there is statistical, so
// any good generator does.  PCG32 keyed by (seed, pixel, sample): a pixel's value does not
// depend on which GPU, tile or launch produced it.
struct Rng {
    uint64_t state, inc;
    BHRT_HD void seed(uint64_t seed, uint64_t pixel, uint64_t sample)
    {
        uint64_t k = seed * 0x9E3779B97F4A7C15ull + pixel * 0xBF58476D1CE4E5B9ull + sample * 0x94D049BB133111EBull;
        k ^= k >> 31; k *= 0xD6E8FEB86659FD93ull; k ^= k >> 32;
        inc = ((pixel * 0x100000001B3ull + sample) << 1) | 1u;
        state = 0; next(); state += k; next();
    }
    BHRT_HD uint32_t next()
    {
        uint64_t old = state;
        state = old * 6364136223846793005ull + inc;
        uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((32u - rot) & 31u));
    }
    // (double)rand()/RAND_MAX (Main.cpp:136) is a uniform variate on a 2^31-point lattice; parity with
    // the reference's stream is statistical, so this is a uniform variate on the 2^24-point lattice
    // every float in [0,1) can hold, with no double-precision division.
    BHRT_HD float unit() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }
    // Rnd01 (MtlBlinn.cpp:42-49): float, redrawn while exactly 0 or 1
    BHRT_HD float rnd01()
    {
        float r = unit();
        while (r == 0.0f) r = unit();
        return r;
    }
};

BHRT_HD float acos_safe(float v) { return m_acos(fminr(1.f, fmaxr(-1.f, v))); }   // cyCore.h:193

// GetRandomCrossingVector (MtlBlinn.cpp:591-600)
BHRT_HD V3 random_crossing_vector(V3 v, Rng &rng)
{
    V3 r(0, 0, 1);
    while (is_zero(cross(v, r))) { float a = rng.rnd01(), b = rng.rnd01(), c = rng.rnd01(); r = V3(a, b, c); }
    return r;
}
// GetSampleAlongNormal (MtlBlinn.cpp:602-617): uniform disk of radius R at the tip of N
BHRT_EST_FN V3 sample_along_normal(V3 N, float R, Rng &rng)
{
    float r = rng.rnd01();
    r = sqrtf(r) * R;
    float theta = (float)(rng.rnd01() * 2 * BHRT_PI_D);
    float x = r * m_cos(theta), y = r * m_sin(theta);
    V3 a1 = cross(random_crossing_vector(N, rng), N);
    V3 a2 = cross(a1, N);
    return N + normalized(a1) * x + normalized(a2) * y;
}
// GetSampleAlongLightDirection (MtlBlinn.cpp:619-635): Phong lobe about N, unnormalised result
BHRT_EST_FN V3 sample_along_direction(V3 N, float gloss, float &theta_out, Rng &rng)
{
    float wt = acos_safe(m_pow(rng.rnd01(), 1.f / (gloss + 1.f)));
    theta_out = wt;
    float R = m_tan(wt);
    float phi = (float)(rng.rnd01() * 2 * BHRT_PI_D);
    float x = R * m_cos(phi), y = R * m_sin(phi);
    V3 a1 = cross(random_crossing_vector(N, rng), N);
    V3 a2 = cross(a1, N);
    return N + normalized(a1) * x + normalized(a2) * y;
}
// GetSampleInSemiSphere (MtlBlinn.cpp:697-716): cosine-weighted hemisphere, retried while N.d <= 0
BHRT_EST_FN V3 sample_semisphere(V3 N, float &theta_out, Rng &rng)
{
    for (int guard = 0; guard < 64; guard++) {
        V3 axisY = normalized(cross(N, random_crossing_vector(N, rng)));
        V3 axisX = cross(N, axisY);
        float phi = (float)(rng.rnd01() * 2 * BHRT_PI_D);
        float rnd = rng.rnd01();
        float theta = 0.5f * acos_safe(1 - 2 * rnd);
        theta_out = theta;
        float st = m_sin(theta);
        V3 ret = st * m_cos(phi) * axisX + st * m_sin(phi) * axisY + m_cos(theta) * N;
        if (dot(N, ret) > 0) return ret;
    }
    return N;
}

// ------------------------------------------------------------------------------------------
// Light::Illuminate (PointLight.cpp:7-18, lights.h:33,50), split around its one shadow ray:
// illuminate_begin builds the shadow query (returns false when the light needs none and the
// irradiance is already known), illuminate_end turns the any-hit answer into irradiance.
BHRT_HD bool illuminate_begin(const bhrt_light &L, V3 p, Rng *rng, Query &q, V3 &irrad)
{
    V3 I = ld3(L.intensity);
    if (L.kind == BHRT_LIGHT_AMBIENT) { irrad = I; return false; }
    q.kind = Q_SHADOW; q.hitSide = HIT_FRONT; q.ray.p = p;
    if (L.kind == BHRT_LIGHT_DIRECT) { q.ray.dir = -ld3(L.v); q.t_max = BHRT_BIGFLOAT; return true; }
    V3 centerDir = ld3(L.v) - p;
    float r = length(centerDir);
    float rr = r * r;
    if (rr == 0) { irrad = V3(1, 1, 1) * BHRT_BIGFLOAT; return false; }
    if (L.size > 0 && rng) q.ray.dir = sample_along_normal(centerDir, L.size, *rng);
    else q.ray.dir = centerDir;
    q.t_max = 1.f;
    return true;
}
BHRT_HD V3 illuminate_end(const bhrt_light &L, V3 p, bool shadowed)
{
    V3 I = ld3(L.intensity);
    float sh = shadowed ? 0.f : 1.f;
    if (L.kind == BHRT_LIGHT_DIRECT) return sh * I;
    V3 centerDir = ld3(L.v) - p;
    float r = length(centerDir);
    float rr = r * r;
    return (sh * I) / rr;
}
// Light::Direction (lights.h:34,51,69)
BHRT_HD V3 light_direction(const bhrt_light &L, V3 p)
{
    if (L.kind == BHRT_LIGHT_AMBIENT) return V3(0, 0, 0);
    if (L.kind == BHRT_LIGHT_DIRECT) return ld3(L.v);
    return normalized(p - ld3(L.v));
}

// GetSampleInLight (MtlBlinn.cpp:637-695)
BHRT_EST_FN V3 sample_in_light(const bhrt_light &L, V3 p, float kd, float ks, float gloss, Rng &rng)
{
    if (L.kind != BHRT_LIGHT_POINT) return -normalized(light_direction(L, p));
    V3 vL = ld3(L.v) - p;
    float dtheta = 0;
    V3 diffuse_vL = sample_along_direction(normalized(vL), gloss, dtheta, rng);
    float p_diffuse = m_pow(m_cos(dtheta), gloss);
    if (ks == 0 && kd != 0) return normalized(diffuse_vL);
    float r = rng.rnd01();
    float R = sqrtf(r) * (float)(int)L.size;         // PointLight::GetSize() returns int (lights.h:76)
    float st = (float)(rng.rnd01() * 2 * BHRT_PI_D);
    float x = R * m_cos(st), y = R * m_sin(st);
    V3 a1 = cross(random_crossing_vector(vL, rng), vL);
    V3 a2 = cross(a1, vL);
    V3 specular_vL = vL + normalized(a1) * x + normalized(a2) * y;
    float p_specular = 2 * r / (R * R);
    if (ks != 0 && kd == 0) return normalized(specular_vL);
    float P_d = kd * p_diffuse, P_s = ks * p_specular;
    float P_dn = P_d / (P_d + P_s);
    float rnd = rng.rnd01();
    bool useSpecular = rnd >= P_dn;
    return useSpecular ? normalized(specular_vL) : normalized(diffuse_vL);
}

// HandleRayWhenRefractionRayOut (MtlBlinn.cpp:543-589)
BHRT_EST_FN Ray handle_refraction_out(const Ray &inRay, const Hit &ih, float ior, bool &toOut, float rgloss, Rng *rng)
{
    V3 vN = ih.N;
    V3 vV = -inRay.dir;
    float cosPhi1 = dot(vV, -vN);
    float sinPhi1 = sqrtf(1 - cosPhi1 * cosPhi1);
    float sinPhi2 = ior * sinPhi1;
    Ray o;
    if (sinPhi2 <= 1) {
        float cosPhi2 = sqrtf(1 - sinPhi2 * sinPhi2);
        V3 vTn = vN * cosPhi2;
        V3 vNxV = cross(vN, vV);
        V3 vTp = normalized(cross(vN, vNxV)) * sinPhi2;
        V3 vT = vTn + vTp;
        V3 vTs = normalized(vT);
        if (rgloss > 0 && rng) {
            float ds = 0; int guard = 0;
            while (ds <= 0 && guard++ < 64) { float th = 0; vTs = sample_along_direction(vT, rgloss, th, *rng); ds = dot(vTs, vN); }
        }
        o.dir = normalized(vTs);
        o.p = ih.p + vN * BHRT_BIAS;
        toOut = true;
    } else {
        o.dir = (-2 * cosPhi1 * vN - vV);
        o.p = ih.p - vN * BHRT_BIAS;
        toOut = false;
    }
    return o;
}

// ------------------------------------------------------------------------------------------
// PhotonMap::EstimateIrradiance<k> + LocatePhotons (reference cyPhotonMap.h:332-382,421-498),
// FILTER_TYPE_CONSTANT, ellipticity 1.  With the constant filter the result depends only on the
// SET of photons the max-heap ends up holding: the k nearest photons within `radius` whose
// direction satisfies dir.N < 0; irrad = sum(power) / (pi * r2), r2 = radius^2 when at most k
// qualified, else the k-th smallest squared distance; direction = normalize(sum(dir * maxpower)).
// A 1000-entry heap per thread is hopeless on a GPU, so the set is found by
// selection instead: pick a grid level whose 3x3x3 block holds more than k qualifying photons
// within one cell edge, then a radix-32 select of the k-th smallest squared distance (32-bin
// histograms of a 20-bit distance key, narrowing to the crossing bin until it holds a handful of
// photons) and an exact selection among those.
struct PmRange { uint32_t b[27], e[27]; };

BHRT_HD uint32_t pm_lower_bound(const uint32_t *keys, uint32_t lo, uint32_t hi, uint32_t key)
{
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// key ranges of the block of search level `level` around (coarse cell c, sub-cell s): level 0 -> the 25 ROWS of the
// 5x5x5 block of coarse cells (the five x-adjacent cells of a row are consecutive coarse cells, i.e. one contiguous
// range of records; entries 25, 26 stay empty); level >= 1 -> the 27 cells of the 3x3x3 block at resolution level - 1
BHRT_HD void pm_ranges(const PhotonGather &pg, int level, const int *c, const int *sub, PmRange &R)
{
    if (level == 0) {
        int n = 0;
        for (int dz = -2; dz <= 2; dz++)
            for (int dy = -2; dy <= 2; dy++, n++) {
                R.b[n] = 0; R.e[n] = 0;
                const int y = c[1] + dy, z = c[2] + dz;
                if (y < 0 || z < 0 || y >= pg.dim[1] || z >= pg.dim[2]) continue;
                int x0 = c[0] - 2, x1 = c[0] + 2;
                if (x0 < 0) x0 = 0;
                if (x1 > pg.dim[0] - 1) x1 = pg.dim[0] - 1;
                if (x0 > x1) continue;
                const uint32_t row = (uint32_t)pg.dim[0] * ((uint32_t)y + (uint32_t)pg.dim[1] * (uint32_t)z);
                R.b[n] = pg.coarse_start[row + (uint32_t)x0]; R.e[n] = pg.coarse_start[row + (uint32_t)x1 + 1u];
            }
        R.b[25] = R.e[25] = R.b[26] = R.e[26] = 0;
        return;
    }
    const int res = level - 1;
    const int sh = BHRT_PM_SUB_BITS - res;            // sub-cell bits dropped at this resolution
    const int span = 1 << res;                        // cells per coarse cell edge
    int g[3] = { (c[0] << res) | (sub[0] >> sh), (c[1] << res) | (sub[1] >> sh), (c[2] << res) | (sub[2] >> sh) };
    int n = 0;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++, n++) {
                int x = g[0] + dx, y = g[1] + dy, z = g[2] + dz;
                R.b[n] = 0; R.e[n] = 0;
                if (x < 0 || y < 0 || z < 0 || x >= pg.dim[0] * span || y >= pg.dim[1] * span || z >= pg.dim[2] * span) continue;
                uint32_t cx = (uint32_t)(x >> res), cy = (uint32_t)(y >> res), cz = (uint32_t)(z >> res);
                uint32_t coarse = cx + (uint32_t)pg.dim[0] * (cy + (uint32_t)pg.dim[1] * cz);
                uint32_t lo = pg.coarse_start[coarse], hi = pg.coarse_start[coarse + 1];
                if (lo == hi) continue;
                if (res == 0) { R.b[n] = lo; R.e[n] = hi; continue; }
                uint32_t m = photon_morton((uint32_t)(x & (span - 1)) << sh, (uint32_t)(y & (span - 1)) << sh, (uint32_t)(z & (span - 1)) << sh);
                uint32_t k0 = (coarse << BHRT_PM_FINE_BITS) | m;
                uint32_t k1 = k0 + (1u << (3 * sh));
                R.b[n] = pm_lower_bound(pg.keys, lo, hi, k0);
                R.e[n] = pm_lower_bound(pg.keys, R.b[n], hi, k1);
            }
}

BHRT_HD V3 pm_power(const GatherRec *rec, float maxpower)
{
    uint32_t rgb = f2u(ld4(&rec->b).w);
    // Photon::GetPower: color.ToColor() * power (cyPhotonMap.h:83, cyColor.h:236)
    return V3((float)(rgb & 255u) / 255.0f, (float)((rgb >> 8) & 255u) / 255.0f, (float)((rgb >> 16) & 255u) / 255.0f) * maxpower;
}

#define BHRT_PM_CAND 48
BHRT_FN void estimate_irradiance(const PhotonGather &pg, V3 pos, V3 N, V3 &irrad, V3 &direction)
{
    irrad = V3(0, 0, 0); direction = V3(0, 0, 0);
    if (pg.n_photons <= 0) return;
    const float r2 = pg.radius * pg.radius;
    // the query's coarse cell and finest sub-cell, computed like photon_key does
    float f[3] = { (pos.x - pg.origin[0]) / pg.cell, (pos.y - pg.origin[1]) / pg.cell, (pos.z - pg.origin[2]) / pg.cell };
    int c[3], sub[3];
    bool inside = true;
    for (int a = 0; a < 3; a++) {
        c[a] = (int)floorf(f[a]);
        if (c[a] < -2 || c[a] > pg.dim[a] + 1) return;      // farther than two cells (>= the radius) from the grid: nothing in range
        int s_ = (int)((f[a] - c[a]) * (float)(1 << BHRT_PM_SUB_BITS));
        sub[a] = s_ < 0 ? 0 : (s_ > (1 << BHRT_PM_SUB_BITS) - 1 ? (1 << BHRT_PM_SUB_BITS) - 1 : s_);
        if (c[a] < 0 || c[a] >= pg.dim[a]) inside = false;
    }
    // starting level from the population of the query's coarse cell (surface density guess):
    // expect ~ n_c * pi / 4^l qualifying photons within one level-l cell edge
    int level = 0;
    if (inside) {
        uint32_t coarse = (uint32_t)c[0] + (uint32_t)pg.dim[0] * ((uint32_t)c[1] + (uint32_t)pg.dim[1] * (uint32_t)c[2]);
        float n_c = (float)(pg.coarse_start[coarse + 1] - pg.coarse_start[coarse]) * 3.14159265f;
        float need = 2.f * pg.k;                            // enough within one coarse cell edge: the 3x3x3 block at that resolution
        while (level < BHRT_PM_LEVELS - 1 && n_c > need) { level++; need *= 4.f; }
    }
    bool coarsened = false;         // a finer level held too few photons: do not try it again
    PmRange R;
    int hist[32];
    for (;;) {
        pm_ranges(pg, level, c, sub, R);
        const float cell_l = pg.cell / (float)(1 << BHRT_PM_RES(level));
        // every photon this close lies in the block: level 0 reaches two coarse cells (photon_grid: 2 * cell >= 1.001 r), so
        // the radius itself is the bound there; a 3x3x3 block is complete up to one cell edge
        float bound2 = level == 0 ? r2 : 0.998f * cell_l * cell_l;
        if (bound2 > r2) bound2 = r2;
        const float scale = 32.f / bound2;
        // ---- pass A: 32-bin histogram of squared distances (and, at level 0, the plain sums)
        for (int i = 0; i < 32; i++) hist[i] = 0;
        int total = 0;
        V3 sumP(0, 0, 0), sumD(0, 0, 0);
        for (int n = 0; n < 27; n++)
            for (uint32_t i = R.b[n]; i < R.e[n]; i++) {
                F4 A = ld4(&pg.recs[i].a);
                V3 dif(A.x - pos.x, A.y - pos.y, A.z - pos.z);
                float d2 = length2(dif);
                if (!(d2 < bound2)) continue;
                F4 B = ld4(&pg.recs[i].b);
                if (dot(V3(B.x, B.y, B.z), N) >= 0) continue;       // cyPhotonMap.h:450
                uint32_t u = (uint32_t)((d2 * scale) * 32768.f); if (u > 0xfffffu) u = 0xfffffu;
                hist[u >> 15]++; total++;
                if (level == 0) { sumP = sumP + pm_power(&pg.recs[i], A.w); sumD = sumD + V3(B.x, B.y, B.z) * A.w; }
            }
        if (total <= pg.k) {
            if (level > 0) { level--; coarsened = true; continue; }           // the k nearest do not fit in this block: one level up
            if (total == 0) return;
            // at most k photons qualified within the radius: area = pi * radius^2 (dist2[0] never shrank)
            float area = 3.14159265358979323846f * r2;
            irrad = sumP * (1.0f / area);
            direction = normalized(sumD);
            return;
        }
        // far more than k: a finer level answers the same question from a quarter of the photons
        if (!coarsened && level < BHRT_PM_LEVELS - 1 && total > 8 * pg.k) { level++; continue; }
        // ---- more than k qualify: find the k-th smallest squared distance by a radix-32 select on
        // u = (int)(d2 * scale * 2^15), a 20-bit key whose top digit is the bin of pass A (the
        // products are exact scalings by powers of two, so every pass sees the same digits)
        const float qscale = 32768.f;
        uint32_t prefix = 0;            // digits of the crossing bin found so far
        int shift = 15;                 // position of the digit `hist` currently holds
        int need = pg.k;                // 1-based rank of the k-th photon among those matching `prefix`
        for (;;) {
            int cum = 0, b = 0;
            for (; b < 31; b++) { if (cum + hist[b] >= need) break; cum += hist[b]; }
            need -= cum;
            prefix = (prefix << 5) | (uint32_t)b;
            if (hist[b] <= BHRT_PM_CAND || shift == 0) break;
            shift -= 5;
            for (int i = 0; i < 32; i++) hist[i] = 0;
            for (int n = 0; n < 27; n++)
                for (uint32_t i = R.b[n]; i < R.e[n]; i++) {
                    F4 A = ld4(&pg.recs[i].a);
                    V3 dif(A.x - pos.x, A.y - pos.y, A.z - pos.z);
                    float d2 = length2(dif);
                    if (!(d2 < bound2)) continue;
                    uint32_t u = (uint32_t)((d2 * scale) * qscale); if (u > 0xfffffu) u = 0xfffffu;
                    if ((u >> (shift + 5)) != prefix) continue;
                    F4 B = ld4(&pg.recs[i].b);
                    if (dot(V3(B.x, B.y, B.z), N) >= 0) continue;
                    hist[(u >> shift) & 31u]++;
                }
        }
        // ---- last pass: sum everything below the crossing bin, collect the crossing bin's photons
        sumP = V3(0, 0, 0); sumD = V3(0, 0, 0);
        float cd2[BHRT_PM_CAND]; uint32_t cidx[BHRT_PM_CAND];
        int ncand = 0;
        bool overflow = false;
        V3 subP(0, 0, 0), subD(0, 0, 0);
        float kth2 = 0, sub_max = 0;
        for (int n = 0; n < 27; n++)
            for (uint32_t i = R.b[n]; i < R.e[n]; i++) {
                F4 A = ld4(&pg.recs[i].a);
                V3 dif(A.x - pos.x, A.y - pos.y, A.z - pos.z);
                float d2 = length2(dif);
                if (!(d2 < bound2)) continue;
                uint32_t u = (uint32_t)((d2 * scale) * qscale); if (u > 0xfffffu) u = 0xfffffu;
                const uint32_t top = u >> shift;
                if (top > prefix) continue;
                F4 B = ld4(&pg.recs[i].b);
                if (dot(V3(B.x, B.y, B.z), N) >= 0) continue;
                if (top < prefix) {
                    sumP = sumP + pm_power(&pg.recs[i], A.w); sumD = sumD + V3(B.x, B.y, B.z) * A.w;
                    if (d2 > kth2) kth2 = d2;
                } else {
                    subP = subP + pm_power(&pg.recs[i], A.w); subD = subD + V3(B.x, B.y, B.z) * A.w;
                    if (d2 > sub_max) sub_max = d2;
                    if (ncand < BHRT_PM_CAND) { cd2[ncand] = d2; cidx[ncand] = i; ncand++; } else overflow = true;
                }
            }
        if (overflow) {
            // more than 48 photons share all 20 key bits (coincident positions): keep them all
            sumP = sumP + subP; sumD = sumD + subD;
            kth2 = sub_max;
        } else {
            // exact: the `need` nearest photons of the crossing bin (insertion sort, <= 48 entries)
            for (int i = 1; i < ncand; i++) {
                float d = cd2[i]; uint32_t ix = cidx[i]; int j = i - 1;
                while (j >= 0 && cd2[j] > d) { cd2[j + 1] = cd2[j]; cidx[j + 1] = cidx[j]; j--; }
                cd2[j + 1] = d; cidx[j + 1] = ix;
            }
            for (int i = 0; i < need && i < ncand; i++) {
                F4 A = ld4(&pg.recs[cidx[i]].a); F4 B = ld4(&pg.recs[cidx[i]].b);
                sumP = sumP + pm_power(&pg.recs[cidx[i]], A.w); sumD = sumD + V3(B.x, B.y, B.z) * A.w;
                kth2 = cd2[i];
            }
        }
        float area = 3.14159265358979323846f * kth2;        // (float)M_PI * np.dist2[0]
        if (area > 0) irrad = sumP * (1.0f / area); else irrad = sumP;
        direction = normalized(sumD);
        return;
    }
}

// ------------------------------------------------------------------------------------------
// The estimator.  MtlBlinn::Shade (MtlBlinn.cpp:89-138) is a recursion TREE with a clamp and
// an early-out at every level, not a linear path, so it is evaluated with an
// explicit frame stack and a post-order resolve.  It is written as a coroutine around ONE ray
// query: sample_advance() runs until the estimator needs a ray, leaves it in Lane::q and
// returns; the caller traces it (scene_trace) and calls sample_advance() again.  In the kernel
// every lane of a warp therefore meets the traversal code together, whatever stage of whatever
// recursion level its sample is in.
struct alignas(128) Frame {              // one 128-byte line per recursion level
    V3 p, vN, vV, uvw, duvw0, duvw1;    // shading point
    V3 out;                              // colour accumulated so far at this level
    V3 mult;                             // factor for the pending child's value
    V3 newSpec;                          // Fresnel-boosted specular colour (MtlBlinn.cpp:110-113)
    int mtl, bounce, gi;
    int useSpec;                         // lobe chosen by the GI sample
    int pad_;
};
enum { ST_REFR = 0, ST_REFR_IN_WAIT = 1, ST_REFR_OUT_WAIT = 2, ST_REFR_WAIT = 3, ST_GI = 4, ST_GI_TRACE_WAIT = 5, ST_GI_WAIT = 6,
       ST_DIRECT = 7, ST_SHADOW_WAIT = 8, ST_GATHER_WAIT = 9 };
#define BHRT_MAX_FRAMES 64          // default depth of the estimator's stack (levels of MtlBlinn::Shade recursion)
#define BHRT_MAX_FRAMES_LIMIT 4096  // the host doubles the depth up to here when a frame overflows it (render.cu)

struct ShadeCfg {
    int mode;           // BHRT_MODE_*
    int gi_depth, tir_budget;
    int max_frames;     // frames available per lane: a Shade() level beyond it is refused (black) and counted
};

// camera sample a lane is evaluating: written when the sample starts, read when it ends
struct alignas(16) LaneCold {
    unsigned long long work;
    int pi, pj, sidx;
    float bgu, bgv;
    int pad_;
};
struct alignas(16) Lane {
    // ---- persistent part: the first BHRT_LANE_PERSIST bytes travel between the lane pool and the
    // thread on every pass
    Query q;                // the pending ray query
    Rng rng;
    V3 ret;                 // value returned by the child frame that just completed
    // scratch of the top frame's wait states (a child frame runs to completion in between, so
    // one set per lane is enough)
    V3 sv;                  // refraction: (1-F)*refraction      direct: colour summed over lights
    V3 sw;                  // direct: the sampled light direction vL
    float sf;               // refraction: glossiness of the refraction lobe
    int si;                 // refraction: remaining internal-reflection budget      direct: light index
    int sp;                 // frames in use; 0 = waiting for the primary ray
    bool hit_ok;            // anything was hit / the shadow ray is blocked
    uint32_t mesh_mask;     // wavefront: mesh nodes whose root box the pending ray entered
    int hit_id; float hit_z;    // primary AOVs of this sample
    // stage of the top recursion level; a level below the top waits for the child above it, either in
    // ST_REFR_WAIT or in ST_GI_WAIT: one bit each (the frames themselves live apart)
    uint8_t top_stage;
    uint8_t pad_stage[3];
    uint32_t gi_wait[3];    // bit i: level i is in ST_GI_WAIT (else ST_REFR_WAIT).  Three words addressed by selects, never
                            // by a run-time index: a dynamically indexed member would pin the whole Lane in local memory
    Hit hit;                // the query's answer; only its first 32 bytes are persistent
    // ---- not copied
    LaneCold cold;
    // what the 8-lanes-per-ray traversal (wide.cuh) needs to start on the pending query, left by the lanes pass: the
    // ray in the space of the first mesh whose root box it entered, kind | hitSide << 8, t_max -- two 128-bit loads
    Ray mesh_ray; int mesh_kind; float mesh_tmax;
};
#define BHRT_LANE_PERSIST 176
// The recursion levels of one lane: the estimator only ever touches the top frame and, when it
// recurses, the slot above it.
struct LaneStack { Frame *fr; };     // cfg.max_frames frames, owned by the caller (CPU harness)
// the stage of the top recursion level, and the push / pop of a level
BHRT_HD int lane_stage(const Lane &L) { return L.top_stage; }
BHRT_HD void lane_push(Lane &L, int child_stage)
{
    const int i = L.sp - 1;         // the level that now waits for the child
    if (i >= 0 && i < 96) {         // (a GI child exists only while f.gi >= 0, i.e. on the first gi_depth + 1 levels)
        const uint32_t bit = 1u << (i & 31), set = L.top_stage == ST_GI_WAIT ? bit : 0u;
        const int w = i >> 5;
        if (w == 0) L.gi_wait[0] = (L.gi_wait[0] & ~bit) | set;
        else if (w == 1) L.gi_wait[1] = (L.gi_wait[1] & ~bit) | set;
        else L.gi_wait[2] = (L.gi_wait[2] & ~bit) | set;
    }
    L.top_stage = (uint8_t)child_stage;
    L.sp++;
}
BHRT_HD void lane_pop(Lane &L)
{
    L.sp--;
    if (L.sp > 0) {
        const int i = L.sp - 1, w = i >> 5;
        const uint32_t word = w == 0 ? L.gi_wait[0] : (w == 1 ? L.gi_wait[1] : (w == 2 ? L.gi_wait[2] : 0u));
        L.top_stage = ((word >> (i & 31)) & 1u) ? ST_GI_WAIT : ST_REFR_WAIT;
    }
}

BHRT_HD void frame_begin(Frame &f, const SceneView &s, V3 ray_p, const Hit &h, int bounce, int gi)
{
    f.p = h.p; f.uvw = h.uvw; f.duvw0 = h.duvw0; f.duvw1 = h.duvw1;
    f.vN = normalized(h.N);
    f.vV = normalized(ray_p - h.p);
    f.mtl = s.nodes[h.node].material + h.sub;
    f.bounce = bounce; f.gi = gi;
    f.out = V3(0, 0, 0);
}

template <class RC> BHRT_HD void issue_closest(Lane &L, const Ray &r, int hitSide, RC &rc)
{
    L.q.ray = r; L.q.kind = Q_CLOSEST; L.q.hitSide = hitSide; L.q.t_max = 0;
    rc.closest++;
}

// Start one camera sample (the body of the loop in PathTracing(), Main.cpp:150-169)
template <class RC> BHRT_HD void sample_begin(Lane &L, const Ray &primary, RC &rc)
{
    L.sp = 0;
    issue_closest(L, primary, HIT_FRONT, rc);
}

// Blinn BRDF * cos of the direct term (MtlBlinn.cpp:319-327) for light direction vL
BHRT_EST_FN V3 direct_brdf(const SceneView &s, const Frame &f, const bhrt_material &M, V3 vL)
{
    float cosTheta = dot(vL, f.vN);
    V3 vH = normalized(vL + f.vV);
    V3 kdT = texcolor_sample_filtered(s, ld3(M.diffuse.color), M.diffuse.texmap, f.uvw, f.duvw0, f.duvw1);
    V3 ksT = texcolor_sample_filtered(s, f.newSpec, M.specular.texmap, f.uvw, f.duvw0, f.duvw1);
    return kdT * cosTheta + ksT * m_pow(dot(vH, f.vN), M.glossiness);
}

// End of PathTracing_DiffuseNSpecular: the caustic gather term (MtlBlinn.cpp:329-342, only when a
// photon map is in use), then the clamp / NaN epilogue (MtlBlinn.cpp:343-350,131-136)
BHRT_EST_FN void direct_finish(const SceneView &s, bool have_gather, V3 irr, V3 vL, Frame &f, const bhrt_material &M, V3 D)
{
    if (have_gather) {
        float cosTheta = -dot(vL, f.vN);
        if (cosTheta > 0) {
            V3 vH = normalized(vL + f.vV);
            V3 kdT = texcolor_sample_filtered(s, ld3(M.diffuse.color), M.diffuse.texmap, f.uvw, f.duvw0, f.duvw1);
            V3 ksT = texcolor_sample_filtered(s, f.newSpec, M.specular.texmap, f.uvw, f.duvw0, f.duvw1);
            V3 brdf = kdT + ksT * m_pow(dot(vH, f.vN), M.glossiness) / cosTheta;
            D = D + brdf * irr;
        }
    }
    clamp_white(D);
    if (is_nan(D.x)) D = V3(0, 0, 0);
    f.out = f.out + D;
    if (!saturated(f.out) && is_nan(f.out.x)) f.out = V3(1.0f, 0.0f, 1.0f);
}

// ONE step of the estimator coroutine: the case of the top frame's stage (or the consumption of
// the primary hit when no frame exists yet).  STEP_RAY: a query was left in L.q and the stage that
// consumes its answer is on top; STEP_CONT: no ray needed, call again; STEP_DONE: the sample is
// finished, colour in `color`.  Stages that consume a ray answer read L.hit / L.hit_ok.
// The wavefront sorts lanes by the stage they are about to run, so a warp executes one case.
// STEP_GATHER: the estimator needs the caustic irradiance estimate at the top frame's (p, vN): the
// caller evaluates it (estimate_irradiance; a whole warp per query in the kernel), leaves the
// irradiance in L.ret and the mean photon direction in L.sw (neither is live in this stage), and
// calls again.
enum { STEP_CONT = 0, STEP_RAY = 1, STEP_DONE = 2, STEP_GATHER = 3 };
// `f` = the top frame (fr[sp-1]; unused while sp == 0), `child` = the slot above it (fr[sp]).
// (RC = RayCounters, or the render kernel's shared-memory counters: five registers fewer in a kernel that has none to spare)
template <class RC> BHRT_HD int sample_step(Lane &BHRT_RESTRICT L, Frame &BHRT_RESTRICT f, Frame &BHRT_RESTRICT child, const SceneView &s, const ShadeCfg &cfg, const PhotonGather &pg, float bgu, float bgv,
                        RC &rc, V3 &color)
{
    // (the fast-math kernels only ever render the stochastic mode -- render.cu picks the exact build for deterministic
    // frames -- so the deterministic branches are compiled out of them)
    const bool det = BHRT_FAST && BHRT_ON_DEVICE ? false : (cfg.mode == BHRT_MODE_DET);
    Rng &rng = L.rng;
    if (L.sp == 0) {
        L.hit_id = L.hit_ok ? L.hit.node : -1;
        L.hit_z = L.hit.z;
        if (!L.hit_ok) { color = texcolor_sample(s, s.background, V3(bgu, bgv, 0.0f)); return STEP_DONE; }   // Main.cpp:166-167
        frame_begin(child, s, L.q.ray.p, L.hit, cfg.tir_budget, cfg.gi_depth);
        lane_push(L, ST_REFR);
        return STEP_CONT;
    }
    {
        uint8_t &stage = L.top_stage;
        const bhrt_material &M = s.materials[f.mtl];
        bool push_child = false;    // a child frame was pushed at fr[sp]
        bool done = false;          // this frame finished, value in f.out
        bool refr_resolved = false; V3 R(0, 0, 0);      // refraction term known
        bool gi_resolved = false; V3 G(0, 0, 0);        // GI term known

        switch (stage) {
        case ST_REFR: {
            // ---- Shade prologue (MtlBlinn.cpp:91-117)
            float cosPhi1 = dot(f.vN, f.vV);
            if (cosPhi1 > 1) cosPhi1 = 1;
            if (cosPhi1 <= 0) cosPhi1 = 0;
            // float R0 = pow((1-ior)/(1+ior), 2); F = R0 + (1-R0)*pow(1-cosPhi1, 5): double pow() in the
            // reference (MtlBlinn.cpp:107-109).  x*x of a float is exact in double; x^5 by three double
            // products is within 1 double-ulp of the correctly rounded power, i.e. the same float
            // except on ~2^-28 of the inputs (shading only, never a hit decision).
            const double r0d = (double)((1 - M.ior) / (1 + M.ior));
            float R0 = (float)(r0d * r0d);
            const double omc = (double)(1 - cosPhi1);
            const double omc2 = omc * omc;
            float F = (float)(R0 + (1 - R0) * (omc2 * omc2 * omc));
            V3 refr = ld3(M.refraction);
            V3 fs = ld3(M.specular.color) + F * refr;
            clamp_white(fs);
            f.newSpec = fs;
            float rgloss = 0;
            if (!det && M.glossiness > 50) rgloss = M.glossiness;
            V3 refrC = (1 - F) * refr;
            // ---- PathTracing_Refraction (MtlBlinn.cpp:437-473)
            if (f.bounce > 0 && !is_zero(refrC)) {
                float sinPhi1 = sqrtf(1 - cosPhi1 * cosPhi1);
                float sinPhi2 = sinPhi1 / M.ior;
                float cosPhi2 = sqrtf(1 - sinPhi2 * sinPhi2);
                V3 vTn = -cosPhi2 * f.vN;
                V3 vNxV = cross(f.vN, f.vV);
                V3 vTp = normalized(cross(f.vN, vNxV)) * sinPhi2;
                V3 vT = vTn + vTp;
                V3 vTs = normalized(vT);
                if (rgloss > 0) {
                    float ds = 0; int guard = 0;
                    while (ds >= 0 && guard++ < 64) { float th = 0; vTs = sample_along_direction(vT, rgloss, th, rng); ds = dot(vTs, f.vN); }
                }
                // ---- RefractionRecusive (MtlBlinn.cpp:476-519); its TIR recursion is the loop
                // ST_REFR_IN_WAIT -> ST_REFR_IN_WAIT below
                L.sv = refrC; L.sf = rgloss; L.si = f.bounce;
                Ray rin; rin.dir = normalized(vTs); rin.p = f.p - f.vN * BHRT_BIAS;
                issue_closest(L, rin, HIT_FRONT_AND_BACK, rc);
                stage = ST_REFR_IN_WAIT;
                return STEP_RAY;
            }
            refr_resolved = true;
            break;
        }
        case ST_REFR_IN_WAIT: {
            const Ray rin = L.q.ray;
            if (!L.hit_ok) { R = V3(1.0f, 0.0f, 1.0f); refr_resolved = true; break; }            // NANPurple (:517)
            if (!L.hit.front) {
                bool goingOut;
                Ray nx = handle_refraction_out(rin, L.hit, M.ior, goingOut, L.sf, det ? (Rng *)0 : &rng);
                if (goingOut) {
                    // ---- RefractionOut (MtlBlinn.cpp:521-541)
                    issue_closest(L, nx, HIT_FRONT, rc);
                    stage = ST_REFR_OUT_WAIT;
                    return STEP_RAY;
                }
                if (L.si <= 0) { refr_resolved = true; break; }
                L.si--;
                Ray r2; r2.dir = nx.dir; r2.p = L.hit.p - L.hit.N * BHRT_BIAS;
                issue_closest(L, r2, HIT_FRONT_AND_BACK, rc);
                return STEP_RAY;
            }
            f.mult = V3(1, 1, 1);
            if (L.sp < cfg.max_frames) {
                frame_begin(child, s, rin.p, L.hit, L.si, f.gi - 1);
                push_child = true;
                stage = ST_REFR_WAIT;
            } else { rc.overflow++; refr_resolved = true; }
            break;
        }
        case ST_REFR_OUT_WAIT: {
            const Ray nx = L.q.ray;
            if (L.hit_ok) {
                V3 ab = ld3(M.absorption);
                float hz = L.hit.z;
                V3 af(m_pow(BHRT_EULER, -ab.x * hz), m_pow(BHRT_EULER, -ab.y * hz), m_pow(BHRT_EULER, -ab.z * hz));
                f.mult = L.sv * af;
                if (L.sp < cfg.max_frames) {
                    frame_begin(child, s, nx.p, L.hit, L.si, f.gi - 1);
                    push_child = true;
                    stage = ST_REFR_WAIT;
                } else { rc.overflow++; refr_resolved = true; }
            } else {
                R = L.sv * sample_environment(s, nx.dir);
                clamp_white(R);
                refr_resolved = true;
            }
            break;
        }
        case ST_REFR_WAIT:
            R = f.mult * L.ret;         // refraction*absorption*Shade(...) (:533) or Shade (:509)
            refr_resolved = true;
            break;
        case ST_GI: {
            // ---- PathTracing_GlobalIllumination (MtlBlinn.cpp:383-433)
            if (det || f.gi < 0) { stage = ST_DIRECT; L.si = 0; L.sv = V3(0, 0, 0); break; }
            // GIUseSpecularDirOrDiffuseDir (:354-378)
            V3 dcol = ld3(M.diffuse.color);
            float kd = fmaxr(fmaxr(dcol.x, dcol.y), dcol.z);
            float ks = fmaxr(fmaxr(f.newSpec.x, f.newSpec.y), f.newSpec.z);
            float dTheta = 0;
            V3 ddir = normalized(sample_semisphere(f.vN, dTheta, rng));
            float p_d = m_sin(2 * dTheta);
            float sTheta = 0;
            float cvn = dot(f.vN, f.vV);
            V3 vR = 2 * cvn * f.vN - f.vV;
            V3 sdir = sample_along_direction(vR, M.glossiness, sTheta, rng);
            float p_s = m_pow(m_cos(sTheta), M.glossiness);
            float P_d = kd * p_d;
            float P_sum = P_d + ks * p_s;
            float P_dn = P_d / P_sum;
            float rnd = rng.rnd01();
            f.useSpec = rnd >= P_dn ? 1 : 0;
            Ray g; g.dir = f.useSpec ? sdir : ddir; g.p = f.p + f.vN * BHRT_BIAS;
            issue_closest(L, g, HIT_FRONT, rc);
            stage = ST_GI_TRACE_WAIT;
            return STEP_RAY;
        }
        case ST_GI_TRACE_WAIT: {
            const Ray g = L.q.ray;
            if (L.hit_ok) {
                if (fabsf(L.hit.z) > BHRT_BIAS && L.sp < cfg.max_frames) {
                    frame_begin(child, s, g.p, L.hit, f.bounce, f.gi - 1);
                    push_child = true;
                    stage = ST_GI_WAIT;
                } else {                         // indirect = black, black * albedo = black
                    if (fabsf(L.hit.z) > BHRT_BIAS) rc.overflow++;
                    gi_resolved = true;
                }
            } else {
                if (g.dir.x == g.dir.y && g.dir.x == 0) G = V3(1.0f, 0.0f, 1.0f);
                else {
                    V3 alb = f.useSpec ? texcolor_sample_filtered(s, f.newSpec, M.specular.texmap, f.uvw, f.duvw0, f.duvw1)
                                       : texcolor_sample_filtered(s, ld3(M.diffuse.color), M.diffuse.texmap, f.uvw, f.duvw0, f.duvw1);
                    V3 env = sample_environment(s, g.dir) * alb;
                    if (!(is_nan(env.x) || is_nan(env.y) || is_nan(env.z))) G = env;
                }
                gi_resolved = true;
            }
            break;
        }
        case ST_GI_WAIT: {
            V3 alb = f.useSpec ? texcolor_sample_filtered(s, f.newSpec, M.specular.texmap, f.uvw, f.duvw0, f.duvw1)
                               : texcolor_sample_filtered(s, ld3(M.diffuse.color), M.diffuse.texmap, f.uvw, f.duvw0, f.duvw1);
            G = L.ret * alb;
            gi_resolved = true;
            break;
        }
        case ST_DIRECT: {
            // ---- PathTracing_DiffuseNSpecular (MtlBlinn.cpp:304-351); L.si / L.sv were reset on entry
            if (det) {
                // D-3 of SURVEY.md 8(c): every light, exact direction
                bool issued = false;
                while (L.si < s.n_lights) {
                    const bhrt_light &Lt = s.lights[L.si];
                    V3 vL = -normalized(light_direction(Lt, f.p));
                    float cosTheta = dot(vL, f.vN);
                    if (cosTheta > 0) {
                        V3 irrad;
                        L.sw = vL;
                        if (illuminate_begin(Lt, f.p, (Rng *)0, L.q, irrad)) { rc.shadow++; stage = ST_SHADOW_WAIT; issued = true; break; }
                        L.sv = L.sv + irrad * direct_brdf(s, f, M, vL);
                    }
                    L.si++;
                }
                if (issued) return STEP_RAY;
                if (pg.use) { stage = ST_GATHER_WAIT; return STEP_GATHER; }
                direct_finish(s, false, V3(0, 0, 0), V3(0, 0, 0), f, M, L.sv);
                done = true;
            } else {
                if (s.n_lights > 0) {
                    float rnd = rng.rnd01();
                    int i = 0;
                    while (rnd > gray(ld3(s.lights[i].intensity)) / s.all_light_intensity && i < s.n_lights - 1) i++;
                    const bhrt_light &Lt = s.lights[i];
                    V3 dcol = ld3(M.diffuse.color);
                    float kd = fmaxr(fmaxr(dcol.x, dcol.y), dcol.z);
                    float ks = fmaxr(fmaxr(f.newSpec.x, f.newSpec.y), f.newSpec.z);
                    V3 vL = sample_in_light(Lt, f.p, kd, ks, M.glossiness, rng);
                    float cosTheta = dot(vL, f.vN);
                    if (cosTheta > 0) {
                        V3 irrad;
                        L.sw = vL; L.si = i;
                        if (illuminate_begin(Lt, f.p, &rng, L.q, irrad)) { rc.shadow++; stage = ST_SHADOW_WAIT; return STEP_RAY; }
                        L.sv = L.sv + irrad * direct_brdf(s, f, M, vL);
                    }
                }
                if (pg.use) { stage = ST_GATHER_WAIT; return STEP_GATHER; }
                direct_finish(s, false, V3(0, 0, 0), V3(0, 0, 0), f, M, L.sv);
                done = true;
            }
            break;
        }
        case ST_SHADOW_WAIT: {
            const bhrt_light &Lt = s.lights[L.si];
            V3 irrad = illuminate_end(Lt, f.p, L.hit_ok);
            L.sv = L.sv + irrad * direct_brdf(s, f, M, L.sw);
            if (det) { L.si++; stage = ST_DIRECT; }
            else if (pg.use) { stage = ST_GATHER_WAIT; return STEP_GATHER; }
            else { direct_finish(s, false, V3(0, 0, 0), V3(0, 0, 0), f, M, L.sv); done = true; }
            break;
        }
        case ST_GATHER_WAIT:
            // the caustic term (MtlBlinn.cpp:329-342) with the estimate the caller left in L.ret / L.sw
            direct_finish(s, true, L.ret, L.sw, f, M, L.sv);
            done = true;
            break;
        }

        if (refr_resolved) {
            clamp_white(R);
            f.out = f.out + R;
            if (saturated(f.out)) done = true; else stage = ST_GI;
        }
        if (gi_resolved) {
            if (is_nan(G.x)) G = V3(1.0f, 0.0f, 1.0f);
            clamp_white(G);
            f.out = f.out + G;
            if (saturated(f.out)) done = true;
            else { stage = ST_DIRECT; L.si = 0; L.sv = V3(0, 0, 0); }
        }
        if (push_child) { lane_push(L, ST_REFR); return STEP_CONT; }
        if (done) {
            L.ret = f.out;
            lane_pop(L);
            if (L.sp == 0) { color = L.ret; return STEP_DONE; }
        }
    }
    return STEP_CONT;
}

// Runs steps until the estimator needs a ray (false, query in L.q) or the sample is finished (true)
BHRT_HD bool sample_advance(Lane &L, LaneStack &S, const SceneView &s, const ShadeCfg &cfg, const PhotonGather &pg, float bgu, float bgv,
                            RayCounters &rc, V3 &color)
{
    for (;;) {
        const int top = L.sp > 0 ? L.sp - 1 : 0, up = L.sp < cfg.max_frames ? L.sp : cfg.max_frames - 1;
        int r = sample_step(L, S.fr[top], S.fr[up], s, cfg, pg, bgu, bgv, rc, color);
        if (r == STEP_GATHER) { estimate_irradiance(pg, S.fr[top].p, S.fr[top].vN, L.ret, L.sw); continue; }
        if (r == STEP_RAY) return false;
        if (r == STEP_DONE) return true;
    }
}


// One camera sample, synchronously: the reference's PathTracing() loop body.  Used by the CPU
// harness (tests/emu); the kernel interleaves the same two calls across a warp.
BHRT_HD V3 trace_sample(Lane &L, LaneStack &S, const SceneView &s, const ShadeCfg &cfg, const PhotonGather &pg, const Ray &primary, float bgu, float bgv,
                        RayCounters &rc, int &hit_id, float &hit_z)
{
    V3 color(0, 0, 0);
    sample_begin(L, primary, rc);
    for (;;) {
        L.hit_ok = scene_trace(s, L.q, L.hit);
        if (sample_advance(L, S, s, cfg, pg, bgu, bgv, rc, color)) break;
    }
    hit_id = L.hit_id; hit_z = L.hit_z;
    return color;
}

// ------------------------------------------------------------------------------------------
// Caustic photon emission: BuildCausticPhotonMap / TraceCausticPhotonRay (reference Main.cpp:319-386),
// PointLight::RandomPhoton (PointLight.cpp:20-34), MtlBlinn::RandomPhotonBounceForCaustic
// (MtlBlinn.cpp:203-303) -- again a coroutine around one ray query per step.
struct PhotonEmitCfg {
    int n_lights;               // point lights only (the reference reinterpret_casts every light to
                                // PointLight, Main.cpp:350 -- undefined for the others), sorted like
                                // ComparePointLight (Main.cpp:247): ascending intensity.Gray() * int(size)
    int light[BHRT_MAX_LIGHTS]; // indices into SceneView::lights
    float sum;                  // sum of Gray() * int(size)  (Main.cpp:357-361)
    int global_map;             // 0: caustic map (BuildCausticPhotonMap, Main.cpp:342-386); 1: the global map
                                // (BuildPhotonMap / TracePhotonRay / MtlBlinn::RandomPhotonBounce, Main.cpp:251-317, MtlBlinn.cpp:140-202)
};
enum { PH_TRACE_WAIT = 0, PH_REFR_IN_WAIT = 1 };
#define BHRT_PHOTON_ABSORB 0.3f     // Photon_AbsorbChance (MtlBlinn.cpp:27)
struct PhotonLane {
    Query q;
    Hit hit;
    bool hit_ok;
    Rng rng;
    V3 c;                       // photon power carried along the path
    int stage;
    bool first;                 // no bounce yet (direct light is not stored, Main.cpp:328)
    int mtl;                    // material of the refractive object being crossed
};

BHRT_HD void photon_begin(PhotonLane &P, const SceneView &s, const PhotonEmitCfg &cfg, unsigned long long &rays)
{
    Rng &rng = P.rng;
    float rnd = rng.rnd01();
    int i = 0;
    // PointLight::GetProbability: Gray() * size (the float) / sum (Lights/lights.h:81)
    while (i < cfg.n_lights - 1 && rnd > gray(ld3(s.lights[cfg.light[i]].intensity)) * s.lights[cfg.light[i]].size / cfg.sum) i++;
    const bhrt_light &Lt = s.lights[cfg.light[i]];
    float phi = (float)(rng.rnd01() * 2 * BHRT_PI_D);
    float theta = acos_safe(1 - 2 * rng.rnd01());
    Ray r;
    r.dir = m_sin(theta) * (V3(1, 0, 0) * m_cos(phi) + V3(0, 1, 0) * m_sin(phi)) + V3(0, 0, 1) * m_cos(theta);
    r.p = ld3(Lt.v);
    P.c = ld3(Lt.intensity);
    P.first = true;
    P.stage = PH_TRACE_WAIT;
    P.q.ray = r; P.q.kind = Q_CLOSEST; P.q.hitSide = HIT_FRONT; P.q.t_max = 0;
    rays++;
}

// The GLOBAL photon map's path (TracePhotonRay, Main.cpp:296-317 + MtlBlinn::RandomPhotonBounce, MtlBlinn.cpp:140-202):
// stored on every photon surface after the first hit; a refractive surface ends the path, 30 % are absorbed, the
// rest bounce diffusely or specularly with the colour divided by the lobe's probability.
BHRT_HD bool photon_advance_global(PhotonLane &P, const SceneView &s, PhotonRecord &rec, bool &store, unsigned long long &rays)
{
    Rng &rng = P.rng;
    store = false;
    if (!P.hit_ok) return true;                                             // Main.cpp:314-316
    const Ray ray = P.q.ray;
    const Hit &h = P.hit;
    const bhrt_material &M = s.materials[s.nodes[h.node].material];
    if (!P.first && gray(ld3(M.diffuse.color)) > 0) {                       // IsPhotonSurface (materials.h:49)
        photon_pack(rec, h.p, normalized(ray.dir), P.c);
        store = true;
    }
    float rnd = rng.rnd01();
    V3 vN = normalized(h.N);
    V3 vV = -normalized(ray.dir);
    if (gray(ld3(M.refraction)) > 0) return true;                           // MtlBlinn.cpp:149-151
    if (rnd < BHRT_PHOTON_ABSORB) return true;
    V3 dcol = ld3(M.diffuse.color), scol = ld3(M.specular.color);
    float kd = fmaxr(fmaxr(dcol.x, dcol.y), dcol.z);
    float ks = fmaxr(fmaxr(scol.x, scol.y), scol.z);
    float dTheta = 0;
    V3 ddir = normalized(sample_semisphere(vN, dTheta, rng));
    float p_d = m_sin(2 * dTheta);
    float sTheta = 0;
    float cvn = dot(vN, vV);
    V3 vR = 2 * cvn * vN - vV;
    V3 sdir = sample_along_direction(vR, M.glossiness, sTheta, rng);
    float p_s = m_pow(m_cos(sTheta), M.glossiness);
    float P_d = kd * p_d;
    float P_sum = P_d + ks * p_s;
    float p_Diff = (P_d / P_sum) * (1 - BHRT_PHOTON_ABSORB) + BHRT_PHOTON_ABSORB;
    float p_Spec = (1 - p_Diff) * (1 - BHRT_PHOTON_ABSORB) + BHRT_PHOTON_ABSORB;
    const bool useSpecular = rnd >= p_Diff;
    P.c = P.c * (useSpecular ? scol / p_Spec : dcol / p_Diff);
    Ray nr; nr.dir = useSpecular ? sdir : ddir; nr.p = h.p + h.N * BHRT_BIAS;
    P.first = false;
    P.stage = PH_TRACE_WAIT;
    P.q.ray = nr; P.q.kind = Q_CLOSEST; P.q.hitSide = HIT_FRONT;
    rays++;
    return false;
}

// Consumes the answer to P.q.  Returns true when the path ended; `store` tells the caller to
// append `rec` to the photon map (PhotonMap::AddPhoton, cyPhotonMap.h:218-232).
BHRT_HD bool photon_advance(PhotonLane &P, const SceneView &s, PhotonRecord &rec, bool &store, unsigned long long &rays)
{
    Rng &rng = P.rng;
    store = false;
    if (P.stage == PH_REFR_IN_WAIT) {
        // inside a refractive object (MtlBlinn.cpp:232-248)
        if (!(P.hit_ok && P.hit.node >= 0)) return true;
        const bhrt_material &M = s.materials[P.mtl];
        bool goingOut;
        Ray nx = handle_refraction_out(P.q.ray, P.hit, M.ior, goingOut, M.refraction_glossiness, &rng);
        if (!goingOut) return true;
        P.first = false;
        P.stage = PH_TRACE_WAIT;
        P.q.ray = nx; P.q.kind = Q_CLOSEST; P.q.hitSide = HIT_FRONT;
        rays++;
        return false;
    }
    if (!P.hit_ok) return true;                                             // Main.cpp:337-339
    const Ray ray = P.q.ray;
    const Hit &h = P.hit;
    const int mtl = s.nodes[h.node].material;
    const bhrt_material &M = s.materials[mtl];
    if (!P.first && gray(ld3(M.diffuse.color)) > 0) {                       // IsPhotonSurface (materials.h:49)
        photon_pack(rec, h.p, normalized(ray.dir), P.c);
        store = true;
    }
    // ---- RandomPhotonBounceForCaustic (MtlBlinn.cpp:203-303)
    float rnd = rng.rnd01();
    V3 vN = normalized(h.N);
    V3 vV = -normalized(ray.dir);
    if (gray(ld3(M.refraction)) > 0) {
        float cosPhi1 = dot(vN, vV);
        float sinPhi1 = sqrtf(1 - cosPhi1 * cosPhi1);
        float sinPhi2 = sinPhi1 / M.ior;
        float cosPhi2 = sqrtf(1 - sinPhi2 * sinPhi2);
        V3 vTn = -cosPhi2 * vN;
        V3 vNxV = cross(vN, vV);
        V3 vTp = normalized(cross(vN, vNxV)) * sinPhi2;
        Ray rin; rin.dir = vTn + vTp; rin.p = h.p - vN * BHRT_BIAS;
        P.mtl = mtl;
        P.stage = PH_REFR_IN_WAIT;
        P.q.ray = rin; P.q.kind = Q_CLOSEST; P.q.hitSide = HIT_BACK;
        rays++;
        return false;
    }
    if (rnd < BHRT_PHOTON_ABSORB) return true;
    V3 dcol = ld3(M.diffuse.color), scol = ld3(M.specular.color);
    float kd = fmaxr(fmaxr(dcol.x, dcol.y), dcol.z);
    float ks = fmaxr(fmaxr(scol.x, scol.y), scol.z);
    float dTheta = 0;
    V3 ddir = normalized(sample_semisphere(vN, dTheta, rng));
    (void)ddir;
    float p_d = m_sin(2 * dTheta);
    float sTheta = 0;
    float cvn = dot(vN, vV);
    V3 vR = 2 * cvn * vN - vV;
    V3 sdir = sample_along_direction(vR, M.glossiness, sTheta, rng);
    float p_s = m_pow(m_cos(sTheta), M.glossiness);
    float P_d = kd * p_d;
    float P_sum = P_d + ks * p_s;
    float p_Diff = (P_d / P_sum) * (1 - BHRT_PHOTON_ABSORB) + BHRT_PHOTON_ABSORB;
    float p_Spec = (1 - p_Diff) * (1 - BHRT_PHOTON_ABSORB) + BHRT_PHOTON_ABSORB;
    if (!(rnd >= p_Diff)) return true;                                      // a diffuse bounce ends a caustic path
    P.c = P.c * (scol / p_Spec);
    Ray nr; nr.dir = sdir; nr.p = h.p + h.N * BHRT_BIAS;
    P.first = false;
    P.stage = PH_TRACE_WAIT;
    P.q.ray = nr; P.q.kind = Q_CLOSEST; P.q.hitSide = HIT_FRONT;
    rays++;
    return false;
}

// RandomPositionInPixel (Main.cpp:132-139) + the primary ray of PathTracing (Main.cpp:145-153)
BHRT_HD Ray primary_ray(const SceneView &s, int i, int j, bool jitter, Rng &rng)
{
    V3 pc = s.top_left + (float)i * s.dd_x - (float)j * s.dd_y;    // (i + 1/2) with 1/2 == 0
    if (jitter) {
        float pixelLen = length(s.dd_x);
        V3 ux = normalized(s.dd_x), uy = normalized(s.dd_y);
        float a = rng.unit() * 2 - 1;
        pc = pc + ux * a * pixelLen / 2;
        float b = rng.unit() * 2 - 1;
        pc = pc + uy * b * pixelLen / 2;
    }
    Ray r; r.p = s.cam_pos; r.dir = pc - s.cam_pos;
    return r;
}

// Color24(Color) (cyColor.h:231,271-272) after pow(c, 1/2.2f) (Main.cpp:220-227)
BHRT_HD uint8_t float_to_byte(float r)
{
    int v = (int)(r * 255 + 0.5f);
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
BHRT_HD V3 gamma_correct(V3 c)
{
    const float ig = 1 / 2.2f;
    return V3(m_pow(c.x, ig), m_pow(c.y, ig), m_pow(c.z, ig));
}

}  // namespace BHRT_NS
