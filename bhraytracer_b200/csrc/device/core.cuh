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
#include "../../../include/bhrt.h"

#ifdef __CUDACC__
#define BHRT_HD __host__ __device__ __forceinline__
#define BHRT_HD_NOINLINE __host__ __device__ __noinline__
#else
#define BHRT_HD inline
#define BHRT_HD_NOINLINE
#endif

namespace bhrt {

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

struct Hit {                // reference scene.h:62-75 (HitInfo)
    float z;
    V3 p, N, uvw, duvw0, duvw1;
    int node;               // index into nodes, -1 = none
    bool front;
    BHRT_HD void init()
    {
        z = BHRT_BIGFLOAT; node = -1; front = true;
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

// ------------------------------------------------------------------------------------------
// Device-side view of an uploaded scene.  Pointers are device pointers in render.cu and
// host pointers in tests/emu.
struct SceneView {
    int n_nodes, n_lights, max_depth, pad0;
    const bhrt_node *nodes;
    const bhrt_material *materials;
    const bhrt_light *lights;           // sorted ascending by grey intensity (Main.cpp:116-118)
    const bhrt_texmap *texmaps;
    const bhrt_texture *textures;
    const uint8_t *texels;
    const bhrt_mesh *meshes;
    const uint8_t *mesh_data;
    const float *tex_offsets;           // 31 (x,y) pairs of the elliptical Halton filter (scene.h:322-329)
    bhrt_texcolor background, environment;
    float all_light_intensity;          // Main.cpp:119-122
    V3 cam_pos, top_left, dd_x, dd_y;   // Main.cpp:179-192, computed on the host
    int width, height;
};

struct RayCounters { unsigned long long closest, shadow; };

// The reference's 24-byte photon (cyPhotonMap.h:72-80): position, max-channel power, 8-bit
// colour ratio, kd-tree plane bits + sign of dir.z, 16-bit fixed-point dir.x / dir.y.
struct PhotonRecord {
    float pos[3];
    float power;
    uint8_t rgb[3];
    uint8_t planeAndDirZ;
    int16_t dirX, dirY;
};

// Read-only view of the gather structure (photon.cuh builds it): photons sorted by the cell
// of a uniform grid whose cell edge is the gather radius, so a query touches 27 cells.
struct PhotonGather {
    int use;                    // 0 = no caustic term (reference built without USE_PhotonMap)
    int k;                      // MAX_PhotonCountInArea (MtlBlinn.cpp:28)
    float radius;               // MAX_Area (MtlBlinn.cpp:29)
    int n_photons;
    const PhotonRecord *photons;
    const uint32_t *cell_start; // [n_cells + 1]
    float origin[3];
    float inv_cell;
    int dim[3];
};

// ------------------------------------------------------------------------------------------
// Sphere::IntersectRay (reference Objects/Sphere/Sphere.cpp:8-75)
BHRT_HD bool intersect_sphere(const Ray &ray, Hit &h, int hitSide)
{
    V3 dir = ray.dir, oc = ray.p;
    float A = dot(dir, dir);
    float B = 2 * dot(dir, oc);
    float C = dot(oc, oc) - 1;
    float DD = B * B - 4 * A * C;
    if (!(DD > 0)) return false;
    float sq = sqrtf(DD);
    float t1 = (-B + sq) / (2 * A);
    float t2 = (-B - sq) / (2 * A);
    float t = BHRT_BIGFLOAT;
    bool front = true;
    if (t1 < 0 && t2 < 0) return false;
    else if (t1 * t2 <= 0) {
        if (hitSide == HIT_FRONT) return false;
        t = t1; front = false;
    } else if (t1 > 0 && t2 > 0) {
        if (hitSide == HIT_FRONT || hitSide == HIT_FRONT_AND_BACK) { t = t2; front = true; }
        else if (hitSide == HIT_BACK) { t = t1; front = false; }
    }
    if (h.z < t || t <= 0) return false;
    h.z = t;
    h.p = oc + h.z * dir;
    h.N = h.p;
    h.front = front;
    V3 d = normalized(h.N);
    // 0.5f + atan2(d.y,d.x)/(2*PI): float atan2f promoted to double, PI is a double literal
    h.uvw.x = (float)(0.5f + atan2f(d.y, d.x) / (2 * BHRT_PI_D));
    h.uvw.y = (float)(0.5f - asinf(d.z) / (BHRT_PI_D));
    // uvw.z keeps whatever the default-constructed Vec3f held in the reference (uninitialised
    // there; never read by 2-D textures).  We define it as 0.
    h.uvw.z = 0.f;
    h.duvw0 = V3(0, 0, 0); h.duvw1 = V3(0, 0, 0);
    return true;
}

// Plane::IntersectRay (reference Objects/Plane/Plane.cpp:8-77) incl. its ray differentials,
// which use the PRIMARY-ray pixel steps dd_x, dd_y for every ray (Plane.cpp:3-4,51-70).
BHRT_HD bool intersect_plane(const Ray &ray, Hit &h, int hitSide, V3 dd_x, V3 dd_y)
{
    float rayPz = ray.p.z, rayDz = ray.dir.z;
    if (rayDz == 0.0f) return false;
    float t = -rayPz / rayDz;
    if (t <= 0 || t > h.z) return false;
    V3 x = ray.p + t * ray.dir;
    if (x.x < -1 || x.x > 1 || x.y < -1 || x.y > 1) return false;
    V3 faceNormal(0, 0, 1);
    bool front = (dot(-ray.dir, faceNormal) > 0);
    if (!front && hitSide == HIT_FRONT) return false;
    else if (front && hitSide == HIT_BACK) return false;
    h.p = x; h.N = faceNormal; h.z = t; h.front = front;
    h.uvw = V3((1 + h.p.x) / 2.f, (1 + h.p.y) / 2.f, 0.f);
    {
        V3 nd = normalized(ray.dir);
        float scaled_t = length(t * ray.dir);
        float ndnd = dot(nd, nd);
        float den = powf(ndnd, 1.5f);
        V3 dDX = (ndnd * dd_x - dot(nd, dd_x) * nd) / den;
        V3 dDY = (ndnd * dd_y - dot(nd, dd_y) * nd) / den;
        float ndN = dot(nd, h.N);
        float d_t_x = -(0 + scaled_t * dot(dDX, h.N) / ndN);
        float d_t_y = -(0 + scaled_t * dot(dDY, h.N) / ndN);
        // "0 + scaled_t * d + d_t * nd": (0 + Vec) is Vec + 0.0f component-wise
        V3 hx = (scaled_t * dDX + V3(0.f, 0.f, 0.f)) + d_t_x * nd;
        V3 hy = (scaled_t * dDY + V3(0.f, 0.f, 0.f)) + d_t_y * nd;
        h.duvw0 = hx / 2.f;
        h.duvw1 = hy / 2.f;
    }
    return true;
}

// Box::IntersectRay (reference Objects/Box/Box.cpp:3-46)
BHRT_HD bool intersect_box(const float *b, const Ray &r, float t_max, float &t_min)
{
    float tz1 = (r.dir.z != 0) ? (b[2] - r.p.z) / r.dir.z : BHRT_BIGFLOAT;
    float tz2 = (r.dir.z != 0) ? (b[5] - r.p.z) / r.dir.z : -BHRT_BIGFLOAT;
    float ty1 = (r.dir.y != 0) ? (b[1] - r.p.y) / r.dir.y : BHRT_BIGFLOAT;
    float ty2 = (r.dir.y != 0) ? (b[4] - r.p.y) / r.dir.y : -BHRT_BIGFLOAT;
    float tx1 = (r.dir.x != 0) ? (b[0] - r.p.x) / r.dir.x : BHRT_BIGFLOAT;
    float tx2 = (r.dir.x != 0) ? (b[3] - r.p.x) / r.dir.x : -BHRT_BIGFLOAT;
    float tMin = fmaxr(fmaxr(fminr(tx1, tx2), fminr(ty1, ty2)), fminr(tz1, tz2));
    float tMax = fminr(fminr(fmaxr(tx1, tx2), fmaxr(ty1, ty2)), fmaxr(tz1, tz2));
    if (tMin <= tMax && tMin < t_max) { t_min = tMin; return true; }
    return false;
}

// ------------------------------------------------------------------------------------------
struct MeshView {
    const float *v, *vn, *vt;
    const uint32_t *f, *fn, *ft;
    const bhrt_bvh_node *bvh;
    const uint32_t *elems;
};
BHRT_HD MeshView mesh_view(const SceneView &s, int mesh)
{
    const bhrt_mesh &m = s.meshes[mesh];
    MeshView mv;
    mv.v = (const float *)(s.mesh_data + m.v_off);
    mv.vn = (const float *)(s.mesh_data + m.vn_off);
    mv.vt = (const float *)(s.mesh_data + m.vt_off);
    mv.f = (const uint32_t *)(s.mesh_data + m.f_off);
    mv.fn = (const uint32_t *)(s.mesh_data + m.fn_off);
    mv.ft = (const uint32_t *)(s.mesh_data + m.ft_off);
    mv.bvh = (const bhrt_bvh_node *)(s.mesh_data + m.bvh_off);
    mv.elems = (const uint32_t *)(s.mesh_data + m.elems_off);
    return mv;
}

// TriObj::IntersectTriangle (reference Objects/TriObj/TriObj.cpp:68-189)
BHRT_HD bool intersect_triangle(const MeshView &m, const Ray &ray, Hit &h, int hitSide, uint32_t face)
{
    uint32_t i0 = m.f[3 * face], i1 = m.f[3 * face + 1], i2 = m.f[3 * face + 2];
    V3 v0 = ld3(m.v + 3 * i0), v1 = ld3(m.v + 3 * i1), v2 = ld3(m.v + 3 * i2);
    V3 vN = cross(v1 - v0, v2 - v0);
    float t_div = dot(vN, ray.dir);
    if (t_div == 0) return false;
    float perp = t_div / (length(vN) * length(ray.dir));
    if (perp > -BHRT_PERP_DET && perp < BHRT_PERP_DET) return false;
    float t = (dot(vN, v0) - dot(vN, ray.p)) / t_div;
    if (t <= 0 || t > h.z) return false;
    bool front = t_div < 0;
    if (!front && hitSide == HIT_FRONT) return false;
    else if (front && hitSide == HIT_BACK) return false;
    V3 vX = ray.p + t * ray.dir;
    float ax = fabsf(vN.x), ay = fabsf(vN.y), az = fabsf(vN.z);
    float p0x, p0y, p1x, p1y, p2x, p2y, pxx, pxy;
    if (ax >= ay && ax >= az) { p0x = v0.y; p0y = v0.z; p1x = v1.y; p1y = v1.z; p2x = v2.y; p2y = v2.z; pxx = vX.y; pxy = vX.z; }
    else if (ay >= ax && ay >= az) { p0x = v0.x; p0y = v0.z; p1x = v1.x; p1y = v1.z; p2x = v2.x; p2y = v2.z; pxx = vX.x; pxy = vX.z; }
    else { p0x = v0.x; p0y = v0.y; p1x = v1.x; p1y = v1.y; p2x = v2.x; p2y = v2.y; pxx = vX.x; pxy = vX.y; }
    // Vec2::Cross(a,b) = a.x*b.y - a.y*b.x (cyVector.h)
    float a0 = ((p1x - pxx) * (p2y - pxy) - (p1y - pxy) * (p2x - pxx)) / 2.f;
    float a1 = ((p2x - pxx) * (p0y - pxy) - (p2y - pxy) * (p0x - pxx)) / 2.f;
    float a2 = ((p0x - pxx) * (p1y - pxy) - (p0y - pxy) * (p1x - pxx)) / 2.f;
    if ((a2 < 0 || a1 < 0 || a0 < 0) && !(a0 < 0 && a1 < 0 && a2 < 0)) return false;
    float a = a0 + a1 + a2;
    V3 bc(a0 / a, a1 / a, a2 / a);
    // TriMesh::Interpolate (cyTriMesh.h:191): v[f.v[0]]*bc.x + v[f.v[1]]*bc.y + v[f.v[2]]*bc.z
    uint32_t n0 = m.fn[3 * face], n1 = m.fn[3 * face + 1], n2 = m.fn[3 * face + 2];
    V3 N = ld3(m.vn + 3 * n0) * bc.x + ld3(m.vn + 3 * n1) * bc.y + ld3(m.vn + 3 * n2) * bc.z;
    uint32_t t0 = m.ft[3 * face], t1 = m.ft[3 * face + 1], t2 = m.ft[3 * face + 2];
    V3 tc = ld3(m.vt + 3 * t0) * bc.x + ld3(m.vt + 3 * t1) * bc.y + ld3(m.vt + 3 * t2) * bc.z;
    h.z = t; h.N = N; h.p = vX; h.front = front; h.uvw = tc;
    h.duvw0 = V3(0, 0, 0); h.duvw1 = V3(0, 0, 0);
    return true;
}

#define BHRT_BVH_LEAF(d) (((d) & 0x80000000u) != 0)
#define BHRT_BVH_COUNT(d) ((((d) >> 28) & 0x7u) + 1)
#define BHRT_BVH_OFFSET(d) ((d) & 0x0fffffffu)
#define BHRT_BVH_CHILD(d) ((d) & 0x7fffffffu)
#define BHRT_BVH_STACK 64

// TriObj::IntersectRay + TraceBVHNode (reference TriObj.cpp:17-39,192-270), made iterative.
// The visiting order, the near/far rule, the "far child is skipped only when the near child
// produced a hit and z <= tmin_far" rule and the traversal of box-missed far children are
// all kept, so ties and rounding cases resolve as in the reference.
BHRT_HD bool mesh_closest(const MeshView &m, const Ray &ray, Hit &h, int hitSide)
{
    float tmin = -1;
    if (!intersect_box(m.bvh[1].box, ray, h.z, tmin)) return false;
    uint32_t st_node[BHRT_BVH_STACK];
    float st_tmin[BHRT_BVH_STACK];
    int st_hits[BHRT_BVH_STACK];
    int sp = 0, hits = 0;
    uint32_t cur = 1;
    for (;;) {
        uint32_t d = m.bvh[cur].data;
        bool descend = false;
        if (BHRT_BVH_LEAF(d)) {
            uint32_t cnt = BHRT_BVH_COUNT(d), off = BHRT_BVH_OFFSET(d);
            for (uint32_t i = 0; i < cnt; i++)
                if (intersect_triangle(m, ray, h, hitSide, m.elems[off + i])) hits++;
        } else {
            uint32_t c1 = BHRT_BVH_CHILD(d), c2 = c1 + 1;
            float tmin1 = BHRT_BIGFLOAT, tmin2 = BHRT_BIGFLOAT;
            bool h1 = intersect_box(m.bvh[c1].box, ray, h.z, tmin1);
            bool h2 = intersect_box(m.bvh[c2].box, ray, h.z, tmin2);
            if (h1 || h2) {
                uint32_t first, second; float tsecond;
                if (tmin1 < tmin2) { first = c1; second = c2; tsecond = tmin2; }
                else { first = c2; second = c1; tsecond = tmin1; }
                if (sp < BHRT_BVH_STACK) { st_node[sp] = second; st_tmin[sp] = tsecond; st_hits[sp] = hits; sp++; }
                cur = first;
                descend = true;
            }
        }
        if (descend) continue;
        // pop
        bool found = false;
        while (sp > 0) {
            sp--;
            bool nearHit = (hits != st_hits[sp]);
            if (nearHit && !(h.z > st_tmin[sp])) continue;
            cur = st_node[sp];
            found = true;
            break;
        }
        if (!found) break;
    }
    return hits > 0;
}

// TriObj::ShadowRecursive + TraceBVHShadow (reference TriObj.cpp:41-66,272-307), iterative.
// Reproduces the reference's behaviour of stopping at the FIRST leaf (depth-first, child1
// before child2) that contains any front-face hit and range-checking only that leaf's t.
BHRT_HD bool mesh_shadow(const MeshView &m, const Ray &ray, float t_max)
{
    float tmin = -1;
    if (!intersect_box(m.bvh[1].box, ray, BHRT_BIGFLOAT, tmin)) return false;
    uint32_t st[BHRT_BVH_STACK];
    int sp = 0;
    uint32_t cur = 1;
    for (;;) {
        uint32_t d = m.bvh[cur].data;
        if (BHRT_BVH_LEAF(d)) {
            uint32_t cnt = BHRT_BVH_COUNT(d), off = BHRT_BVH_OFFSET(d);
            Hit h; h.init();
            bool any = false;
            for (uint32_t i = 0; i < cnt; i++)
                if (intersect_triangle(m, ray, h, HIT_FRONT, m.elems[off + i])) any = true;
            if (any) return h.z > BHRT_BIAS && h.z < t_max;
        } else {
            uint32_t c1 = BHRT_BVH_CHILD(d), c2 = c1 + 1;
            float t1 = BHRT_BIGFLOAT, t2 = BHRT_BIGFLOAT;
            bool h1 = intersect_box(m.bvh[c1].box, ray, BHRT_BIGFLOAT, t1);
            bool h2 = intersect_box(m.bvh[c2].box, ray, BHRT_BIGFLOAT, t2);
            if (h1 || h2) {
                if (sp < BHRT_BVH_STACK) st[sp++] = c2;
                cur = c1;
                continue;
            }
        }
        if (sp == 0) return false;
        cur = st[--sp];
    }
}

// ------------------------------------------------------------------------------------------
// recursive() (reference Main.cpp:389-413): linear DFS over the scene graph, closest hit.
// nodes[] is in DFS pre-order; rs[d] is the ray in the space of the current node at depth d.
BHRT_HD bool scene_closest(const SceneView &s, const Ray &ray, Hit &h, int hitSide)
{
    Ray rs[BHRT_MAX_DEPTH + 1];
    rs[0] = ray;
    V3 lp(0, 0, 0), lN(0, 0, 0);
    for (int k = 0; k < s.n_nodes; k++) {
        const bhrt_node &nd = s.nodes[k];
        int d = nd.depth;
        Ray r = to_node(nd.xf, rs[d - 1]);
        if (d < s.max_depth) rs[d] = r;
        bool hit = false;
        if (nd.kind == BHRT_OBJ_SPHERE) hit = intersect_sphere(r, h, hitSide);
        else if (nd.kind == BHRT_OBJ_PLANE) hit = intersect_plane(r, h, hitSide, s.dd_x, s.dd_y);
        else if (nd.kind == BHRT_OBJ_MESH) hit = mesh_closest(mesh_view(s, nd.mesh), r, h, hitSide);
        if (hit) { h.node = k; lp = h.p; lN = h.N; }
    }
    if (h.node < 0) return false;
    // Node::FromNodeCoords of the hit node (Main.cpp:401) and then of its parent
    // (Main.cpp:407-412; the rootNode's identity transform for depth-1 nodes).  Deeper
    // ancestors are NOT applied by the reference; neither are they here.
    const bhrt_node &nd = s.nodes[h.node];
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

// GenLight::Shadow / ShadowRayRecursive (reference Lights/GenLight.cpp:10-69): any-hit.
// The result is the OR over all nodes of the node's own test, so DFS pre-order is as good as
// the reference's post-order.  Quirks kept: sphere ignored when the origin is inside
// (:36-37), plane bounds tested with the PARENT-space ray (:54), no rayDz==0 test (:50).
BHRT_HD bool scene_shadow(const SceneView &s, const Ray &ray, float t_max)
{
    Ray rs[BHRT_MAX_DEPTH + 1];
    rs[0] = ray;     // rootNode.ToNodeCoords is the identity and exact
    for (int k = 0; k < s.n_nodes; k++) {
        const bhrt_node &nd = s.nodes[k];
        int d = nd.depth;
        const Ray &pr = rs[d - 1];
        Ray r = to_node(nd.xf, pr);
        if (d < s.max_depth) rs[d] = r;
        if (nd.kind == BHRT_OBJ_SPHERE) {
            float A = dot(r.dir, r.dir);
            float B = 2 * dot(r.dir, r.p);
            float C = dot(r.p, r.p) - 1;
            float DD = B * B - 4 * A * C;
            if (DD > 0) {
                float sq = sqrtf(DD);
                float t1 = (-B + sq) / (2 * A);
                float t2 = (-B - sq) / (2 * A);
                float t = fminr(t1, t2);
                if (!(t < 0) && t < t_max && t > BHRT_SHADOW_BIAS) return true;
            }
        } else if (nd.kind == BHRT_OBJ_PLANE) {
            float t = -r.p.z / r.dir.z;
            if (!(t < 0)) {
                V3 x = pr.p + t * pr.dir;
                if (!(x.x < -1 || x.x > 1 || x.y < -1 || x.y > 1) && t < t_max && t > BHRT_SHADOW_BIAS) return true;
            }
        } else if (nd.kind == BHRT_OBJ_MESH) {
            if (mesh_shadow(mesh_view(s, nd.mesh), r, t_max)) return true;
        }
    }
    return false;
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
BHRT_HD V3 texel(const uint8_t *p) { return V3(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f); }

BHRT_HD V3 texture_sample(const SceneView &s, const bhrt_texture &t, V3 uvw)
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
    const uint8_t *d = s.texels + t.texel_offset;
    return texel(d + 3 * (iy * width + ix)) * ((1 - fx) * (1 - fy)) +
           texel(d + 3 * (iy * width + ixp)) * (fx * (1 - fy)) +
           texel(d + 3 * (iyp * width + ix)) * ((1 - fx) * fy) +
           texel(d + 3 * (iyp * width + ixp)) * (fx * fy);
}

// Texture::Sample(uvw, duvw) (scene.h:318-337): 32-tap elliptical Halton filter when duvw != 0
BHRT_HD V3 texture_sample_filtered(const SceneView &s, const bhrt_texture &t, V3 uvw, V3 d0, V3 d1)
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
BHRT_HD V3 texcolor_sample_filtered(const SceneView &s, V3 color, int texmap, V3 uvw, V3 d0, V3 d1)
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
BHRT_HD V3 sample_environment(const SceneView &s, V3 dir)
{
    float z = asinf(-dir.z) / 3.14159265358979323846f + 0.5f;
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
    // (double)rand()/RAND_MAX with glibc's RAND_MAX = 2^31-1 (Main.cpp:136)
    BHRT_HD double unit() { return (double)(next() >> 1) / 2147483647.0; }
    // Rnd01 (MtlBlinn.cpp:42-49): float, redrawn while exactly 0 or 1
    BHRT_HD float rnd01()
    {
        float r = (float)unit();
        while (r == 0.0f || r == 1.0f) r = (float)unit();
        return r;
    }
};

BHRT_HD float acos_safe(float v) { return acosf(fminr(1.f, fmaxr(-1.f, v))); }   // cyCore.h:193

// GetRandomCrossingVector (MtlBlinn.cpp:591-600)
BHRT_HD V3 random_crossing_vector(V3 v, Rng &rng)
{
    V3 r(0, 0, 1);
    while (is_zero(cross(v, r))) { float a = rng.rnd01(), b = rng.rnd01(), c = rng.rnd01(); r = V3(a, b, c); }
    return r;
}
// GetSampleAlongNormal (MtlBlinn.cpp:602-617): uniform disk of radius R at the tip of N
BHRT_HD V3 sample_along_normal(V3 N, float R, Rng &rng)
{
    float r = rng.rnd01();
    r = sqrtf(r) * R;
    float theta = (float)(rng.rnd01() * 2 * BHRT_PI_D);
    float x = r * cosf(theta), y = r * sinf(theta);
    V3 a1 = cross(random_crossing_vector(N, rng), N);
    V3 a2 = cross(a1, N);
    return N + normalized(a1) * x + normalized(a2) * y;
}
// GetSampleAlongLightDirection (MtlBlinn.cpp:619-635): Phong lobe about N, unnormalised result
BHRT_HD V3 sample_along_direction(V3 N, float gloss, float &theta_out, Rng &rng)
{
    float wt = acos_safe(powf(rng.rnd01(), 1.f / (gloss + 1.f)));
    theta_out = wt;
    float R = tanf(wt);
    float phi = (float)(rng.rnd01() * 2 * BHRT_PI_D);
    float x = R * cosf(phi), y = R * sinf(phi);
    V3 a1 = cross(random_crossing_vector(N, rng), N);
    V3 a2 = cross(a1, N);
    return N + normalized(a1) * x + normalized(a2) * y;
}
// GetSampleInSemiSphere (MtlBlinn.cpp:697-716): cosine-weighted hemisphere, retried while N.d <= 0
BHRT_HD V3 sample_semisphere(V3 N, float &theta_out, Rng &rng)
{
    for (int guard = 0; guard < 64; guard++) {
        V3 axisY = normalized(cross(N, random_crossing_vector(N, rng)));
        V3 axisX = cross(N, axisY);
        float phi = (float)(rng.rnd01() * 2 * BHRT_PI_D);
        float rnd = rng.rnd01();
        float theta = 0.5f * acos_safe(1 - 2 * rnd);
        theta_out = theta;
        float st = sinf(theta);
        V3 ret = st * cosf(phi) * axisX + st * sinf(phi) * axisY + cosf(theta) * N;
        if (dot(N, ret) > 0) return ret;
    }
    return N;
}

// ------------------------------------------------------------------------------------------
// Light::Illuminate (PointLight.cpp:7-18, lights.h:33,50).  One shadow ray.
BHRT_HD V3 illuminate(const SceneView &s, const bhrt_light &L, V3 p, Rng *rng, RayCounters &rc)
{
    V3 I = ld3(L.intensity);
    if (L.kind == BHRT_LIGHT_AMBIENT) return I;
    if (L.kind == BHRT_LIGHT_DIRECT) {
        Ray r; r.p = p; r.dir = -ld3(L.v);
        rc.shadow++;
        float sh = scene_shadow(s, r, BHRT_BIGFLOAT) ? 0.f : 1.f;
        return sh * I;
    }
    V3 centerDir = ld3(L.v) - p;
    float r = length(centerDir);
    float rr = r * r;
    if (rr == 0) return V3(1, 1, 1) * BHRT_BIGFLOAT;
    Ray sr; sr.p = p;
    if (L.size > 0 && rng) sr.dir = sample_along_normal(centerDir, L.size, *rng);
    else sr.dir = centerDir;
    rc.shadow++;
    float sh = scene_shadow(s, sr, 1.f) ? 0.f : 1.f;
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
BHRT_HD V3 sample_in_light(const bhrt_light &L, V3 p, float kd, float ks, float gloss, Rng &rng)
{
    if (L.kind != BHRT_LIGHT_POINT) return -normalized(light_direction(L, p));
    V3 vL = ld3(L.v) - p;
    float dtheta = 0;
    V3 diffuse_vL = sample_along_direction(normalized(vL), gloss, dtheta, rng);
    float p_diffuse = powf(cosf(dtheta), gloss);
    if (ks == 0 && kd != 0) return normalized(diffuse_vL);
    float r = rng.rnd01();
    float R = sqrtf(r) * (float)(int)L.size;         // PointLight::GetSize() returns int (lights.h:76)
    float st = (float)(rng.rnd01() * 2 * BHRT_PI_D);
    float x = R * cosf(st), y = R * sinf(st);
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
BHRT_HD Ray handle_refraction_out(const Ray &inRay, const Hit &ih, float ior, bool &toOut, float rgloss, Rng *rng)
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
// The estimator.  MtlBlinn::Shade (MtlBlinn.cpp:89-138) is a recursion TREE with a clamp and
// an early-out at every level, not a linear path, so it is evaluated here
// with an explicit frame stack and a post-order resolve.
struct Frame {
    V3 p, vN, vV, uvw, duvw0, duvw1;    // shading point
    V3 out;                              // colour accumulated so far at this level
    V3 mult;                             // factor for the pending child's value
    V3 newSpec;                          // Fresnel-boosted specular colour (MtlBlinn.cpp:110-113)
    int mtl, bounce, gi;
    int stage;
    bool useSpec;                        // lobe chosen by the GI sample
    bool z_ok;
};
enum { ST_REFR = 0, ST_REFR_WAIT = 1, ST_GI = 2, ST_GI_WAIT = 3, ST_DIRECT = 4 };
#define BHRT_MAX_FRAMES 24

struct ShadeCfg {
    int mode;           // BHRT_MODE_*
    int gi_depth, tir_budget;
};

BHRT_HD void frame_begin(Frame &f, const SceneView &s, const Ray &ray, const Hit &h, int bounce, int gi)
{
    f.p = h.p; f.uvw = h.uvw; f.duvw0 = h.duvw0; f.duvw1 = h.duvw1;
    f.vN = normalized(h.N);
    f.vV = normalized(ray.p - h.p);
    f.mtl = s.nodes[h.node].material;
    f.bounce = bounce; f.gi = gi;
    f.out = V3(0, 0, 0);
    f.stage = ST_REFR;
}

// One camera sample (the body of the loop in PathTracing(), Main.cpp:150-169).
// Returns the sample's colour; fills the primary hit id / z.
BHRT_HD_NOINLINE V3 trace_sample(const SceneView &s, const ShadeCfg &cfg, const PhotonGather &pg, const Ray &primary, float bgu, float bgv,
                                 Rng &rng, RayCounters &rc, int &hit_id, float &hit_z)
{
    const bool det = (cfg.mode == BHRT_MODE_DET);
    Hit h; h.init();
    rc.closest++;
    bool bHit = scene_closest(s, primary, h, HIT_FRONT);
    hit_id = bHit ? h.node : -1;
    hit_z = h.z;
    if (!bHit) return texcolor_sample(s, s.background, V3(bgu, bgv, 0.0f));   // Main.cpp:166-167

    Frame fr[BHRT_MAX_FRAMES];
    int sp = 1;
    frame_begin(fr[0], s, primary, h, cfg.tir_budget, cfg.gi_depth);
    V3 ret(0, 0, 0);        // value returned by the child that just completed
    bool have_ret = false;

    for (;;) {
        Frame &f = fr[sp - 1];
        const bhrt_material &M = s.materials[f.mtl];
        bool push_child = false;    // a child frame was pushed at fr[sp]
        bool done = false;          // this frame finished, value in f.out

        if (f.stage == ST_REFR) {
            // ---- Shade prologue (MtlBlinn.cpp:91-117)
            float cosPhi1 = dot(f.vN, f.vV);
            if (cosPhi1 > 1) cosPhi1 = 1;
            if (cosPhi1 <= 0) cosPhi1 = 0;
            float R0 = (float)pow((double)((1 - M.ior) / (1 + M.ior)), 2.0);
            float F = (float)(R0 + (1 - R0) * pow((double)(1 - cosPhi1), 5.0));
            V3 refr = ld3(M.refraction);
            V3 fs = ld3(M.specular.color) + F * refr;
            clamp_white(fs);
            f.newSpec = fs;
            float rgloss = 0;
            if (!det && M.glossiness > 50) rgloss = M.glossiness;
            V3 refrC = (1 - F) * refr;
            // ---- PathTracing_Refraction (MtlBlinn.cpp:437-473)
            V3 R(0, 0, 0);
            bool resolved = true;
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
                // ---- RefractionRecusive (MtlBlinn.cpp:476-519); the TIR recursion is a loop
                V3 cur_p = f.p, cur_N = f.vN, dir = normalized(vTs);
                int b = f.bounce;
                for (;;) {
                    Ray rin; rin.dir = dir; rin.p = cur_p - cur_N * BHRT_BIAS;
                    Hit hin; hin.init();
                    rc.closest++;
                    bool hitIn = scene_closest(s, rin, hin, HIT_FRONT_AND_BACK);
                    if (!hitIn) { R = V3(1.0f, 0.0f, 1.0f); break; }           // NANPurple (:517)
                    if (!hin.front) {
                        bool goingOut;
                        Ray nx = handle_refraction_out(rin, hin, M.ior, goingOut, rgloss, det ? (Rng *)0 : &rng);
                        if (goingOut) {
                            // ---- RefractionOut (MtlBlinn.cpp:521-541)
                            Hit ho; ho.init();
                            rc.closest++;
                            bool hitOut = scene_closest(s, nx, ho, HIT_FRONT);
                            if (hitOut) {
                                V3 ab = ld3(M.absorption);
                                V3 af(powf(BHRT_EULER, -ab.x * ho.z), powf(BHRT_EULER, -ab.y * ho.z), powf(BHRT_EULER, -ab.z * ho.z));
                                f.mult = refrC * af;
                                if (sp < BHRT_MAX_FRAMES) {
                                    frame_begin(fr[sp], s, nx, ho, b, f.gi - 1);
                                    push_child = true; resolved = false;
                                } else R = V3(0, 0, 0);
                            } else {
                                R = refrC * sample_environment(s, nx.dir);
                                clamp_white(R);
                            }
                            break;
                        } else {
                            if (b <= 0) { R = V3(0, 0, 0); break; }
                            b--;
                            cur_p = hin.p; cur_N = hin.N; dir = nx.dir;
                            continue;
                        }
                    } else {
                        f.mult = V3(1, 1, 1);
                        if (sp < BHRT_MAX_FRAMES) {
                            frame_begin(fr[sp], s, rin, hin, b, f.gi - 1);
                            push_child = true; resolved = false;
                        } else R = V3(0, 0, 0);
                        break;
                    }
                }
            }
            if (resolved) {
                clamp_white(R);
                f.out = f.out + R;
                if (saturated(f.out)) done = true; else f.stage = ST_GI;
            } else f.stage = ST_REFR_WAIT;
        } else if (f.stage == ST_REFR_WAIT) {
            V3 R = f.mult * ret;        // refraction*absorption*Shade(...) (:533) or Shade (:509)
            clamp_white(R);
            f.out = f.out + R;
            have_ret = false;
            if (saturated(f.out)) done = true; else f.stage = ST_GI;
        } else if (f.stage == ST_GI) {
            // ---- PathTracing_GlobalIllumination (MtlBlinn.cpp:383-433)
            if (det || f.gi < 0) f.stage = ST_DIRECT;
            else {
                // GIUseSpecularDirOrDiffuseDir (:354-378)
                V3 dcol = ld3(M.diffuse.color);
                float kd = fmaxr(fmaxr(dcol.x, dcol.y), dcol.z);
                float ks = fmaxr(fmaxr(f.newSpec.x, f.newSpec.y), f.newSpec.z);
                float dTheta = 0;
                V3 ddir = normalized(sample_semisphere(f.vN, dTheta, rng));
                float p_d = sinf(2 * dTheta);
                float sTheta = 0;
                float cvn = dot(f.vN, f.vV);
                V3 vR = 2 * cvn * f.vN - f.vV;
                V3 sdir = sample_along_direction(vR, M.glossiness, sTheta, rng);
                float p_s = powf(cosf(sTheta), M.glossiness);
                float P_d = kd * p_d;
                float P_sum = P_d + ks * p_s;
                float P_dn = P_d / P_sum;
                float rnd = rng.rnd01();
                f.useSpec = rnd >= P_dn;
                Ray g; g.dir = f.useSpec ? sdir : ddir; g.p = f.p + f.vN * BHRT_BIAS;
                Hit hg; hg.init();
                rc.closest++;
                bool hitG = scene_closest(s, g, hg, HIT_FRONT);
                V3 G(0, 0, 0);
                bool resolved = true;
                if (hitG) {
                    if (fabsf(hg.z) > BHRT_BIAS && sp < BHRT_MAX_FRAMES) {
                        frame_begin(fr[sp], s, g, hg, f.bounce, f.gi - 1);
                        push_child = true; resolved = false;
                    }
                    // else: indirect = black, black * albedo = black
                } else {
                    if (g.dir.x == g.dir.y && g.dir.x == 0) G = V3(1.0f, 0.0f, 1.0f);
                    else {
                        V3 alb = f.useSpec ? texcolor_sample_filtered(s, f.newSpec, M.specular.texmap, f.uvw, f.duvw0, f.duvw1)
                                           : texcolor_sample_filtered(s, ld3(M.diffuse.color), M.diffuse.texmap, f.uvw, f.duvw0, f.duvw1);
                        V3 env = sample_environment(s, g.dir) * alb;
                        if (!(is_nan(env.x) || is_nan(env.y) || is_nan(env.z))) G = env;
                    }
                }
                if (resolved) {
                    if (is_nan(G.x)) G = V3(1.0f, 0.0f, 1.0f);
                    clamp_white(G);
                    f.out = f.out + G;
                    if (saturated(f.out)) done = true; else f.stage = ST_DIRECT;
                } else f.stage = ST_GI_WAIT;
            }
        } else if (f.stage == ST_GI_WAIT) {
            V3 alb = f.useSpec ? texcolor_sample_filtered(s, f.newSpec, M.specular.texmap, f.uvw, f.duvw0, f.duvw1)
                               : texcolor_sample_filtered(s, ld3(M.diffuse.color), M.diffuse.texmap, f.uvw, f.duvw0, f.duvw1);
            V3 G = ret * alb;
            have_ret = false;
            if (is_nan(G.x)) G = V3(1.0f, 0.0f, 1.0f);
            clamp_white(G);
            f.out = f.out + G;
            if (saturated(f.out)) done = true; else f.stage = ST_DIRECT;
        } else {
            // ---- PathTracing_DiffuseNSpecular (MtlBlinn.cpp:304-351)
            V3 D(0, 0, 0);
            V3 dcol = ld3(M.diffuse.color);
            if (det) {
                // D-3 of SURVEY.md 8(c): every light, exact direction
                for (int li = 0; li < s.n_lights; li++) {
                    const bhrt_light &L = s.lights[li];
                    V3 vL = -normalized(light_direction(L, f.p));
                    float cosTheta = dot(vL, f.vN);
                    if (cosTheta > 0) {
                        V3 vH = normalized(vL + f.vV);
                        V3 irrad = illuminate(s, L, f.p, (Rng *)0, rc);
                        V3 kdT = texcolor_sample_filtered(s, dcol, M.diffuse.texmap, f.uvw, f.duvw0, f.duvw1);
                        V3 ksT = texcolor_sample_filtered(s, f.newSpec, M.specular.texmap, f.uvw, f.duvw0, f.duvw1);
                        V3 brdf = kdT * cosTheta + ksT * powf(dot(vH, f.vN), M.glossiness);
                        D = D + irrad * brdf;
                    }
                }
            } else if (s.n_lights > 0) {
                float rnd = rng.rnd01();
                int i = 0;
                while (rnd > gray(ld3(s.lights[i].intensity)) / s.all_light_intensity && i < s.n_lights - 1) i++;
                const bhrt_light &L = s.lights[i];
                float kd = fmaxr(fmaxr(dcol.x, dcol.y), dcol.z);
                float ks = fmaxr(fmaxr(f.newSpec.x, f.newSpec.y), f.newSpec.z);
                V3 vL = sample_in_light(L, f.p, kd, ks, M.glossiness, rng);
                float cosTheta = dot(vL, f.vN);
                if (cosTheta > 0) {
                    V3 vH = normalized(vL + f.vV);
                    V3 irrad = illuminate(s, L, f.p, &rng, rc);
                    V3 kdT = texcolor_sample_filtered(s, dcol, M.diffuse.texmap, f.uvw, f.duvw0, f.duvw1);
                    V3 ksT = texcolor_sample_filtered(s, f.newSpec, M.specular.texmap, f.uvw, f.duvw0, f.duvw1);
                    V3 brdf = kdT * cosTheta + ksT * powf(dot(vH, f.vN), M.glossiness);
                    D = D + irrad * brdf;
                }
            }
            clamp_white(D);
            if (is_nan(D.x)) D = V3(0, 0, 0);
            f.out = f.out + D;
            if (!saturated(f.out) && is_nan(f.out.x)) f.out = V3(1.0f, 0.0f, 1.0f);   // Shade epilogue (:131-136)
            done = true;
        }

        if (push_child) { sp++; continue; }
        if (done) {
            ret = fr[sp - 1].out;
            have_ret = true;
            sp--;
            if (sp == 0) break;
        }
    }
    (void)have_ret;
    return ret;
}

// RandomPositionInPixel (Main.cpp:132-139) + the primary ray of PathTracing (Main.cpp:145-153)
BHRT_HD Ray primary_ray(const SceneView &s, int i, int j, bool jitter, Rng &rng)
{
    V3 pc = s.top_left + (float)i * s.dd_x - (float)j * s.dd_y;    // (i + 1/2) with 1/2 == 0
    if (jitter) {
        float pixelLen = length(s.dd_x);
        V3 ux = normalized(s.dd_x), uy = normalized(s.dd_y);
        float a = (float)(rng.unit() * 2 - 1);
        pc = pc + ux * a * pixelLen / 2;
        float b = (float)(rng.unit() * 2 - 1);
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
    return V3(powf(c.x, ig), powf(c.y, ig), powf(c.z, ig));
}

}  // namespace bhrt
