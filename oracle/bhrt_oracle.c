/* bhrt_oracle.c -- TEST INFRASTRUCTURE, never linked into the product.
 *
 * A plain-C restatement of the reference's deterministic hot path (spheres, planes and triangle
 * meshes traversed through the reference's own BVH topology; point / direct / ambient lights;
 * Blinn-Phong with hard shadows; Fresnel refraction with absorption and total internal reflection;
 * checker and bitmap textures with the plane's ray differentials; textured background and
 * environment; the caustic gather EstimateIrradiance / LocatePhotons over the reference's own
 * balanced photon array; 1 spp at the pixel corner), written from the
 * reference sources cited at every function (paths relative to /root/reference/BHRayTracer).  It is
 * independent of csrc/device: the only thing it shares with the product is the POD scene
 * description of include/bhrt.h.
 *
 * PINNED: tests/test_oracle_restatement.py checks it against the committed outputs of the
 * reference itself (all five tests/golden/*_det.npz, made by oracle/_ref/bhrt_ref --det): primary
 * hit ids and z bit-exact, RGB8 identical, ray counts identical; and, where /root/reference is
 * present, against the reference run live on 16 of its own scenes and on this repository's five
 * config scenes.  The real reference (oracle/_ref, built by oracle/build_ref.py) remains the
 * oracle for what is outside this file's scope -- the stochastic estimator (GI, area lights,
 * jitter) and the photon emission / kd-tree balancing -- where parity is statistical and a
 * restatement would only repeat the reference with more room for error.  The gather itself is
 * restated (bhrt_oracle_render_det_pm) and pinned by tests/golden/det_caustics_pm.npz: the
 * reference's gather-lit frame from its own photon dump, identical in every 8-bit channel.
 *
 * Scope check: bhrt_oracle_render_det returns -2 for scenes with area lights (the D-mode switches
 * leave PointLight::Illuminate's random sample in place, PointLight.cpp:13-15).
 * "D-mode" = the four switches of SURVEY.md 8(c): no jitter, GI off, every light with its exact
 * direction, no glossy refraction.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/Makefile).  x86-64 without -mfma
 * contracts nothing, like the reference's own build.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/bhrt.h"

#define BIGFLOAT 1.0e30f          /* Scenes/scene.h:38 */
#define SHADOW_BIAS 0.00001f      /* Lights/GenLight.cpp:5 */
#define BIAS 0.0001f              /* Materials/Blinn/MtlBlinn.cpp:10 */
#define EULER_N 2.7182818f        /* MtlBlinn.cpp:11 */
#define TEXTURE_SAMPLE_COUNT 32   /* Scenes/scene.h:18 */
#define PI_F ((float)3.14159265358979323846)   /* (float)M_PI */
#define PI_D 3.14159265           /* "#define PI 3.14159265" (Main.cpp, a double literal) */
enum { HIT_FRONT = 1, HIT_BACK = 2, HIT_FRONT_AND_BACK = 3 };      /* scene.h:57-60 */

typedef struct { float x, y, z; } v3;
static v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
/* cyVector.h:273-386 -- component-wise, in this order */
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 mulf(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static v3 divf(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static float len(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }     /* cy::Sqrt = sqrtss */
static v3 norm(v3 a) { return divf(a, len(a)); }                                /* GetNormalized: *this / Length() */
static v3 ld(const float *p) { return V(p[0], p[1], p[2]); }

/* Matrix3f * Vec3f, column-major cells (cyMatrix.h:297-302) */
static v3 mat_mul(const float *m, v3 p)
{
    return V(p.x * m[0] + p.y * m[3] + p.z * m[6], p.x * m[1] + p.y * m[4] + p.z * m[7], p.x * m[2] + p.y * m[5] + p.z * m[8]);
}
/* Transformation::TransposeMult (scene.h:238-245) */
static v3 mat_tmul(const float *m, v3 d)
{
    return V(m[0] * d.x + m[1] * d.y + m[2] * d.z, m[3] * d.x + m[4] * d.y + m[5] * d.z, m[6] * d.x + m[7] * d.y + m[8] * d.z);
}
static v3 transform_to(const bhrt_xform *t, v3 p) { return mat_mul(t->itm, sub(p, ld(t->pos))); }      /* scene.h:220 */
static v3 transform_from(const bhrt_xform *t, v3 p) { return add(mat_mul(t->tm, p), ld(t->pos)); }     /* scene.h:221 */
static v3 vector_transform_from(const bhrt_xform *t, v3 d) { return mat_tmul(t->itm, d); }             /* scene.h:227 */

typedef struct { v3 p, dir; } ray_t;
/* Node::ToNodeCoords (scene.h:490-496): the direction is the difference of two transformed points */
static ray_t to_node_coords(const bhrt_xform *t, ray_t r)
{
    ray_t o;
    o.p = transform_to(t, r.p);
    o.dir = sub(transform_to(t, add(r.p, r.dir)), o.p);
    return o;
}

typedef struct { float z; v3 p, N, uvw, duvw[2]; int node; int front; int mtlID; } hit_t;     /* scene.h:62-75 */
/* oracle patch Q-1 of build_ref.py / BHRT_FLAG_FACE_MATERIALS of the product: TriObj::IntersectTriangle sets hInfo.mtlID to the
 * usemtl group of the face (HEAD never does, TriObj.cpp:68-189); 0 = the reference as it is */
static int g_face_materials = 0;
void bhrt_oracle_set_face_materials(int on) { g_face_materials = on; }
/* HitInfo::Init (scene.h:74) */
static void hit_init(hit_t *h) { h->z = BIGFLOAT; h->node = -1; h->front = 1; h->mtlID = 0; h->p = V(0, 0, 0); h->N = V(0, 0, 0); h->uvw = V(0.5f, 0.5f, 0.5f); h->duvw[0] = V(0, 0, 0); h->duvw[1] = V(0, 0, 0); }

/* Sphere::IntersectRay (Objects/Sphere/Sphere.cpp:8-75), unit sphere at the origin */
static int sphere_intersect(ray_t ray, hit_t *h, int hitSide)
{
    v3 dir = ray.dir, oc = ray.p;
    float A = dot(dir, dir);
    float B = 2 * dot(dir, oc);
    float C = dot(oc, oc) - 1;
    float DD = B * B - 4 * A * C;
    if (DD > 0) {
        float t1 = (-B + sqrtf(DD)) / (2 * A);
        float t2 = (-B - sqrtf(DD)) / (2 * A);
        float t = BIGFLOAT;
        int front = 1;
        if (t1 < 0 && t2 < 0) return 0;
        else if (t1 * t2 <= 0) {
            if (hitSide == HIT_FRONT) return 0;
            t = t1; front = 0;
        } else if (t1 > 0 && t2 > 0) {
            if (hitSide == HIT_FRONT || hitSide == HIT_FRONT_AND_BACK) { t = t2; front = 1; }
            else if (hitSide == HIT_BACK) { t = t1; front = 0; }
        }
        if (h->z < t || t <= 0) return 0;
        h->z = t;
        h->p = add(oc, mulf(dir, h->z));
        h->N = h->p;
        h->front = front;
        /* :57-70 -- float atan2/asin promoted to double by the double literal PI; uvw.z is never
         * assigned in the reference (an uninitialised local), 2-D textures do not read it */
        v3 d = norm(h->N);
        h->uvw.x = (float)(0.5f + atan2f(d.y, d.x) / (2 * PI_D));
        h->uvw.y = (float)(0.5f - asinf(d.z) / (PI_D));
        h->uvw.z = 0;
        h->duvw[0] = V(0, 0, 0); h->duvw[1] = V(0, 0, 0);
        return 1;
    }
    return 0;
}

/* Plane::IntersectRay (Objects/Plane/Plane.cpp:8-49), the z = 0 plane within [-1,1]^2 */
static int plane_intersect(ray_t ray, hit_t *h, int hitSide, v3 dd_x, v3 dd_y)
{
    float rayPz = ray.p.z, rayDz = ray.dir.z;
    if (rayDz == 0.0f) return 0;
    float t = -rayPz / rayDz;
    if (t <= 0 || t > h->z) return 0;
    v3 x = add(ray.p, mulf(ray.dir, t));
    if (x.x < -1 || x.x > 1 || x.y < -1 || x.y > 1) return 0;
    int front = dot(neg(ray.dir), V(0, 0, 1)) > 0;
    if (!front && hitSide == HIT_FRONT) return 0;
    else if (front && hitSide == HIT_BACK) return 0;
    h->p = x; h->N = V(0, 0, 1); h->z = t; h->front = front;
    h->uvw = V((1 + h->p.x) / 2.f, (1 + h->p.y) / 2.f, 0);
    /* RAY_DIFFERENTIAL block (:51-70): the globals dd_x, dd_y are the PRIMARY rays' pixel steps, used
     * for every ray; "0 + vec" is a component-wise + 0.0f */
    v3 nd = norm(ray.dir);
    float scaled_t = len(mulf(ray.dir, t));
    float ndnd = dot(nd, nd);
    float den = powf(ndnd, 1.5f);
    v3 dDX = divf(sub(mulf(dd_x, ndnd), mulf(nd, dot(nd, dd_x))), den);
    v3 dDY = divf(sub(mulf(dd_y, ndnd), mulf(nd, dot(nd, dd_y))), den);
    float d_t_x = -(0 + scaled_t * dot(dDX, h->N) / dot(nd, h->N));
    float d_t_y = -(0 + scaled_t * dot(dDY, h->N) / dot(nd, h->N));
    v3 zero = V(0, 0, 0);
    v3 hx = add(add(mulf(dDX, scaled_t), zero), mulf(nd, d_t_x));
    v3 hy = add(add(mulf(dDY, scaled_t), zero), mulf(nd, d_t_y));
    h->duvw[0] = divf(hx, 2.f);
    h->duvw[1] = divf(hy, 2.f);
    return 1;
}

/* ---- triangle meshes ------------------------------------------------------------------------- */
#define PERP_FACE_DET 0.001745f        /* Objects/TriObj/TriObj.cpp:12 */
static float fmin_cy(float a, float b) { return a <= b ? a : b; }      /* cyCore.h:187-188 */
static float fmax_cy(float a, float b) { return a >= b ? a : b; }
typedef struct {
    const float *v, *vn, *vt;
    const uint32_t *f, *fn, *ft;
    const bhrt_bvh_node *bvh;
    const uint32_t *elems;
    const int32_t *group_end; int n_groups;   /* cyTriMesh mcfc (cyTriMesh.h:113,467-490): bhrt_mesh::n_groups */
} mesh_t;
static mesh_t mesh_view(const bhrt_scene_desc *d, int i)
{
    const bhrt_mesh *m = &d->meshes[i];
    const uint8_t *b = d->mesh_data;
    mesh_t r = { (const float *)(b + m->v_off), (const float *)(b + m->vn_off), (const float *)(b + m->vt_off),
                 (const uint32_t *)(b + m->f_off), (const uint32_t *)(b + m->fn_off), (const uint32_t *)(b + m->ft_off),
                 (const bhrt_bvh_node *)(b + m->bvh_off), (const uint32_t *)(b + m->elems_off),
                 (const int32_t *)(b + ((m->elems_off + 4 * (int64_t)m->nf + 15) & ~(int64_t)15)), m->n_groups };
    return r;
}
/* Box::IntersectRay (Objects/Box/Box.cpp:3-46); Corner(0) = pmin, Corner(7) = pmax (scene.h:100-107) */
static int box_intersect(const float *box, ray_t r, float t_max, float *t_min)
{
    v3 v0 = V(box[0], box[1], box[2]), v7 = V(box[3], box[4], box[5]);
    v3 nz = V(0, 0, 1), ny = V(0, 1, 0), nx = V(1, 0, 0);
    float nDotrr = dot(nz, r.dir), nDotrp = dot(nz, r.p);
    float tz1 = (nDotrr != 0) ? (dot(nz, v0) - nDotrp) / nDotrr : BIGFLOAT;
    float tz2 = (nDotrr != 0) ? (dot(nz, v7) - nDotrp) / nDotrr : -BIGFLOAT;
    nDotrr = dot(ny, r.dir); nDotrp = dot(ny, r.p);
    float ty1 = (nDotrr != 0) ? (dot(ny, v0) - nDotrp) / nDotrr : BIGFLOAT;
    float ty2 = (nDotrr != 0) ? (dot(ny, v7) - nDotrp) / nDotrr : -BIGFLOAT;
    nDotrr = dot(nx, r.dir); nDotrp = dot(nx, r.p);
    float tx1 = (nDotrr != 0) ? (dot(nx, v0) - nDotrp) / nDotrr : BIGFLOAT;
    float tx2 = (nDotrr != 0) ? (dot(nx, v7) - nDotrp) / nDotrr : -BIGFLOAT;
    float tMin = fmax_cy(fmax_cy(fmin_cy(tx1, tx2), fmin_cy(ty1, ty2)), fmin_cy(tz1, tz2));
    float tMax = fmin_cy(fmin_cy(fmax_cy(tx1, tx2), fmax_cy(ty1, ty2)), fmax_cy(tz1, tz2));
    if (tMin <= tMax && tMin < t_max) { *t_min = tMin; return 1; }
    return 0;
}
static float cross2(float ax, float ay, float bx, float by) { return (-ay) * bx + ax * by; }     /* cyVector.h:260 */
/* TriMesh::Interpolate (cyTriMesh.h:191) */
static v3 interpolate(const float *v, const uint32_t *f, uint32_t face, v3 bc)
{
    return add(add(mulf(ld(v + 3 * f[3 * face]), bc.x), mulf(ld(v + 3 * f[3 * face + 1]), bc.y)), mulf(ld(v + 3 * f[3 * face + 2]), bc.z));
}
/* TriObj::IntersectTriangle (TriObj.cpp:68-189) */
static int triangle_intersect(const mesh_t *m, ray_t ray, hit_t *h, int hitSide, uint32_t face)
{
    v3 v0 = ld(m->v + 3 * m->f[3 * face]), v1 = ld(m->v + 3 * m->f[3 * face + 1]), v2 = ld(m->v + 3 * m->f[3 * face + 2]);
    v3 vN = cross(sub(v1, v0), sub(v2, v0));
    float t_divisor = dot(vN, ray.dir);
    if (t_divisor == 0) return 0;
    float perp = t_divisor / (len(vN) * len(ray.dir));
    if (perp > -PERP_FACE_DET && perp < PERP_FACE_DET) return 0;
    float t = (dot(vN, v0) - dot(vN, ray.p)) / t_divisor;
    if (t <= 0 || t > h->z) return 0;
    int front = t_divisor < 0;
    if (!front && hitSide == HIT_FRONT) return 0;
    else if (front && hitSide == HIT_BACK) return 0;
    v3 vX = add(ray.p, mulf(ray.dir, t));
    v3 a = V(fabsf(vN.x), fabsf(vN.y), fabsf(vN.z));
    float p0x = 0, p0y = 0, p1x = 0, p1y = 0, p2x = 0, p2y = 0, px = 0, py = 0;
    if (a.x >= a.y && a.x >= a.z) { p0x = v0.y; p0y = v0.z; p1x = v1.y; p1y = v1.z; p2x = v2.y; p2y = v2.z; px = vX.y; py = vX.z; }
    else if (a.y >= a.x && a.y >= a.z) { p0x = v0.x; p0y = v0.z; p1x = v1.x; p1y = v1.z; p2x = v2.x; p2y = v2.z; px = vX.x; py = vX.z; }
    else if (a.z >= a.y && a.z >= a.x) { p0x = v0.x; p0y = v0.y; p1x = v1.x; p1y = v1.y; p2x = v2.x; p2y = v2.y; px = vX.x; py = vX.y; }
    float a0 = cross2(p1x - px, p1y - py, p2x - px, p2y - py) / 2.f;
    float a1 = cross2(p2x - px, p2y - py, p0x - px, p0y - py) / 2.f;
    float a2 = cross2(p0x - px, p0y - py, p1x - px, p1y - py) / 2.f;
    if ((a2 < 0 || a1 < 0 || a0 < 0) && !(a0 < 0 && a1 < 0 && a2 < 0)) return 0;
    float area = a0 + a1 + a2;
    v3 bc = V(a0 / area, a1 / area, a2 / area);
    h->z = t;
    h->N = interpolate(m->vn, m->fn, face, bc);            /* GetNormal: not normalised */
    h->p = vX;
    h->front = front;
    h->uvw = interpolate(m->vt, m->ft, face, bc);
    h->duvw[0] = V(0, 0, 0); h->duvw[1] = V(0, 0, 0);
    if (g_face_materials) {                                /* TriMesh::GetMaterialIndex (cyTriMesh.h:221-227); none -> 0 */
        int g = 0;
        while (g < m->n_groups && (int)face >= m->group_end[g]) g++;
        h->mtlID = g < m->n_groups ? g : 0;
    }
    return 1;
}
#define BVH_LEAF(d) ((d) >> 31)                            /* cyBVH.h:187-203 */
#define BVH_COUNT(d) ((((d) >> 28) & 7u) + 1u)
#define BVH_OFFSET(d) ((d) & 0x0fffffffu)
#define BVH_CHILD(d) ((d) & 0x7fffffffu)
/* TriObj::TraceBVHNode (TriObj.cpp:192-270) */
static int trace_bvh_node(const mesh_t *m, ray_t ray, hit_t *h, int hitSide, uint32_t node)
{
    uint32_t data = m->bvh[node].data;
    if (BVH_LEAF(data)) {
        int bHit = 0;
        for (uint32_t i = 0; i < BVH_COUNT(data); i++)
            if (triangle_intersect(m, ray, h, hitSide, m->elems[BVH_OFFSET(data) + i])) bHit = 1;
        return bHit;
    }
    uint32_t c1 = BVH_CHILD(data), c2 = c1 + 1;
    float tmin1 = BIGFLOAT, tmin2 = BIGFLOAT;
    int hit1 = box_intersect(m->bvh[c1].box, ray, h->z, &tmin1);
    int hit2 = box_intersect(m->bvh[c2].box, ray, h->z, &tmin2);
    if (!hit1 && !hit2) return 0;
    uint32_t nearc = tmin1 < tmin2 ? c1 : c2, farc = tmin1 < tmin2 ? c2 : c1;
    float tmin_far = tmin1 < tmin2 ? tmin2 : tmin1;
    if (trace_bvh_node(m, ray, h, hitSide, nearc)) {
        hit_t tmp = *h;
        if (tmp.z > tmin_far && trace_bvh_node(m, ray, &tmp, hitSide, farc)) *h = tmp;
        return 1;
    }
    return trace_bvh_node(m, ray, h, hitSide, farc);
}
/* TriObj::IntersectRay (TriObj.cpp:17-39) */
static int mesh_intersect(const mesh_t *m, ray_t ray, hit_t *h, int hitSide)
{
    float tmin = -1;
    if (box_intersect(m->bvh[1].box, ray, h->z, &tmin)) return trace_bvh_node(m, ray, h, hitSide, 1);
    return 0;
}
/* TriObj::TraceBVHShadow (TriObj.cpp:272-307); `a | b` with both operands evaluated, left first (gcc) */
static int trace_bvh_shadow(const mesh_t *m, ray_t ray, float *t_min, int *hitOnce, uint32_t node)
{
    if (*hitOnce) return *hitOnce;
    uint32_t data = m->bvh[node].data;
    if (BVH_LEAF(data)) {
        hit_t h;
        hit_init(&h);
        for (uint32_t i = 0; i < BVH_COUNT(data); i++)
            if (triangle_intersect(m, ray, &h, HIT_FRONT, m->elems[BVH_OFFSET(data) + i])) { *hitOnce = 1; *t_min = h.z; }
        return *hitOnce;
    }
    uint32_t c1 = BVH_CHILD(data), c2 = c1 + 1;
    float tmin1 = BIGFLOAT, tmin2 = BIGFLOAT;
    int hit1 = box_intersect(m->bvh[c1].box, ray, BIGFLOAT, &tmin1);
    int hit2 = box_intersect(m->bvh[c2].box, ray, BIGFLOAT, &tmin2);
    if (!hit1 && !hit2) return 0;
    int r1 = trace_bvh_shadow(m, ray, t_min, hitOnce, c1);
    int r2 = trace_bvh_shadow(m, ray, t_min, hitOnce, c2);
    return r1 | r2;
}
/* TriObj::ShadowRecursive (TriObj.cpp:41-66) */
static int mesh_shadow(const mesh_t *m, ray_t ray, float t_max)
{
    float tmin = -1, t_min = BIGFLOAT;
    int bHit = 0;
    if (box_intersect(m->bvh[1].box, ray, BIGFLOAT, &tmin)) trace_bvh_shadow(m, ray, &t_min, &bHit, 1);
    return bHit && t_min > BIAS && t_min < t_max;
}

/* Node::FromNodeCoords(HitInfo) (scene.h:497-501) */
static void from_node_coords(const bhrt_xform *t, hit_t *h)
{
    h->p = transform_from(t, h->p);
    h->N = norm(vector_transform_from(t, h->N));
}

typedef struct {
    const bhrt_scene_desc *d;
    int *first_child, *next_sibling;      /* the Node tree rebuilt from the DFS table */
    uint64_t rays_closest, rays_shadow;
    int depth, max_depth;                 /* Shade() recursion: current / deepest level of the frame */
    const uint8_t *photons; int n_photons, half_photons;    /* the caustic map: balanced kd-tree in heap order (the reference's .dat dump), or NULL */
    v3 dd_x, dd_y;                        /* Main.cpp:31 globals, read by Plane.cpp */
    const bhrt_light *lights; int n_lights;       /* sorted (Main.cpp:112-123) */
    float tex_x[TEXTURE_SAMPLE_COUNT], tex_y[TEXTURE_SAMPLE_COUNT];   /* scene.h:323-329 offsets */
} scene_t;
static const bhrt_xform IDENTITY = { { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, { 0, 0, 0 } };

/* recursive() (Main.cpp:389-413); `root` = -1 is rootNode (identity transform) */
static void recursive(scene_t *s, int root, ray_t ray, hit_t *out, int *bHit, int hitSide)
{
    int first = root < 0 ? (s->d->n_nodes > 0 ? 0 : -1) : s->first_child[root];
    if (first < 0) return;
    for (int c = first; c >= 0; c = (root < 0 ? s->next_sibling[c] : s->next_sibling[c])) {
        const bhrt_node *n = &s->d->nodes[c];
        ray_t tr = to_node_coords(&n->xf, ray);
        if (n->kind != BHRT_OBJ_NONE) {
            int got;
            if (n->kind == BHRT_OBJ_SPHERE) got = sphere_intersect(tr, out, hitSide);
            else if (n->kind == BHRT_OBJ_PLANE) got = plane_intersect(tr, out, hitSide, s->dd_x, s->dd_y);
            else { mesh_t m = mesh_view(s->d, n->mesh); got = mesh_intersect(&m, tr, out, hitSide); }
            if (got) { out->node = c; *bHit = 1; from_node_coords(&n->xf, out); }
        }
        recursive(s, c, tr, out, bHit, hitSide);
    }
    for (int c = first; c >= 0; c = s->next_sibling[c])
        if (c == out->node) { from_node_coords(root < 0 ? &IDENTITY : &s->d->nodes[root].xf, out); break; }
}

/* ShadowRayRecursive (Lights/GenLight.cpp:15-69): children first, then the node's own object; the
 * plane test bounds-checks with the PARENT-space ray (:54), the sphere test ignores origins inside (:36) */
static int shadow_recursive(scene_t *s, int root, ray_t ray, float t_max)
{
    ray_t tr = to_node_coords(root < 0 ? &IDENTITY : &s->d->nodes[root].xf, ray);
    int first = root < 0 ? (s->d->n_nodes > 0 ? 0 : -1) : s->first_child[root];
    for (int c = first; c >= 0; c = s->next_sibling[c])
        if (shadow_recursive(s, c, tr, t_max)) return 1;
    if (root >= 0 && s->d->nodes[root].kind == BHRT_OBJ_SPHERE) {
        v3 dir = tr.dir, oc = tr.p;
        float A = dot(dir, dir), B = 2 * dot(dir, oc), C = dot(oc, oc) - 1;
        float DD = B * B - 4 * A * C;
        if (DD > 0) {
            float t1 = (-B + sqrtf(DD)) / (2 * A), t2 = (-B - sqrtf(DD)) / (2 * A);
            float t = t1 <= t2 ? t1 : t2;                       /* cy::Min */
            if (t < 0) return 0;
            if (t < t_max && t > SHADOW_BIAS) return 1;
        }
    } else if (root >= 0 && s->d->nodes[root].kind == BHRT_OBJ_PLANE) {
        float t = -tr.p.z / tr.dir.z;
        if (t < 0) return 0;
        v3 x = add(ray.p, mulf(ray.dir, t));                    /* (sic) the untransformed ray */
        if (x.x < -1 || x.x > 1 || x.y < -1 || x.y > 1) return 0;
        if (t < t_max && t > SHADOW_BIAS) return 1;
        return 0;
    } else if (root >= 0 && s->d->nodes[root].kind == BHRT_OBJ_MESH) {
        mesh_t m = mesh_view(s->d, s->d->nodes[root].mesh);
        return mesh_shadow(&m, tr, t_max);
    }
    return 0;
}
/* GenLight::Shadow (GenLight.cpp:10-13) */
static float shadow(scene_t *s, ray_t ray, float t_max)
{
    s->rays_shadow++;
    return shadow_recursive(s, -1, ray, t_max) ? 0.f : 1.f;
}

/* Light::Illuminate / Direction (Lights/lights.h:29-87, PointLight.cpp:7-18); size 0 in D-mode scenes */
static v3 illuminate(scene_t *s, const bhrt_light *L, v3 p)
{
    v3 I = ld(L->intensity);
    if (L->kind == BHRT_LIGHT_AMBIENT) return I;
    if (L->kind == BHRT_LIGHT_DIRECT) { ray_t r = { p, neg(ld(L->v)) }; return mulf(I, shadow(s, r, BIGFLOAT)); }
    v3 centerDir = sub(ld(L->v), p);
    float r = len(centerDir), rr = r * r;
    if (rr == 0) return V(BIGFLOAT, BIGFLOAT, BIGFLOAT);
    ray_t sr = { p, centerDir };
    return divf(mulf(I, shadow(s, sr, 1)), rr);
}
static v3 light_direction(const bhrt_light *L, v3 p)
{
    if (L->kind == BHRT_LIGHT_AMBIENT) return V(0, 0, 0);
    if (L->kind == BHRT_LIGHT_DIRECT) return ld(L->v);
    return norm(sub(p, ld(L->v)));
}
static float gray(const bhrt_light *L) { return (L->intensity[0] + L->intensity[1] + L->intensity[2]) / 3.0f; }   /* cyColor.h:83 */
static void clamp_white(v3 *c) { if (c->x > 1) c->x = 1.f; if (c->y > 1) c->y = 1.f; if (c->z > 1) c->z = 1.f; }   /* MtlBlinn.cpp:79-83 */

/* ---- textures ------------------------------------------------------------------------------ */
/* Texture::TileClamp (scene.h:344-354) */
static v3 tile_clamp(v3 uvw)
{
    v3 u = V(uvw.x - (int)uvw.x, uvw.y - (int)uvw.y, uvw.z - (int)uvw.z);
    if (u.x < 0) u.x += 1;
    if (u.y < 0) u.y += 1;
    if (u.z < 0) u.z += 1;
    return u;
}
static v3 texel(const scene_t *s, const bhrt_texture *T, int i)        /* Color24::ToColor (cyColor.h:236) */
{
    const uint8_t *t = s->d->texels + T->texel_offset + 3 * (size_t)i;
    return V(t[0] / 255.0f, t[1] / 255.0f, t[2] / 255.0f);
}
/* TextureFile::Sample (Textures/Texture.cpp:100-123), TextureChecker::Sample (:127-136) */
static v3 texture_sample(const scene_t *s, const bhrt_texture *T, v3 uvw)
{
    v3 u = tile_clamp(uvw);
    if (T->kind == BHRT_TEX_CHECKER) {
        if (u.x <= 0.5f) return u.y <= 0.5f ? ld(T->color1) : ld(T->color2);
        return u.y <= 0.5f ? ld(T->color2) : ld(T->color1);
    }
    int width = T->width, height = T->height;
    if (width + height == 0) return V(0, 0, 0);
    float x = width * u.x, y = height * u.y;
    int ix = (int)x, iy = (int)y;
    float fx = x - ix, fy = y - iy;
    if (ix < 0) ix -= (ix / width - 1) * width;
    if (ix >= width) ix -= (ix / width) * width;
    int ixp = ix + 1;
    if (ixp >= width) ixp -= width;
    if (iy < 0) iy -= (iy / height - 1) * height;
    if (iy >= height) iy -= (iy / height) * height;
    int iyp = iy + 1;
    if (iyp >= height) iyp -= height;
    return add(add(add(mulf(texel(s, T, iy * width + ix), (1 - fx) * (1 - fy)), mulf(texel(s, T, iy * width + ixp), fx * (1 - fy))),
                   mulf(texel(s, T, iyp * width + ix), (1 - fx) * fy)), mulf(texel(s, T, iyp * width + ixp), fx * fy));
}
/* Halton (scene.h:137-146) */
static float halton(int index, int base)
{
    float r = 0, f = 1.0f / (float)base;
    for (int i = index; i > 0; i /= base) { r += f * (i % base); f /= (float)base; }
    return r;
}
/* Texture::Sample(uvw, duvw, elliptic = true) (scene.h:318-337) */
static v3 texture_sample_filtered(const scene_t *s, const bhrt_texture *T, v3 uvw, const v3 *duvw)
{
    v3 c = texture_sample(s, T, uvw);
    if (dot(duvw[0], duvw[0]) + dot(duvw[1], duvw[1]) == 0) return c;
    for (int i = 1; i < TEXTURE_SAMPLE_COUNT; i++)
        c = add(c, texture_sample(s, T, add(add(uvw, mulf(duvw[0], s->tex_x[i])), mulf(duvw[1], s->tex_y[i]))));
    return divf(c, (float)TEXTURE_SAMPLE_COUNT);
}
/* TexturedColor::Sample(uvw) / TextureMap::Sample(uvw) (scene.h:371,410) */
static v3 texcolor_sample(const scene_t *s, const bhrt_texcolor *tc, v3 uvw)
{
    if (tc->texmap < 0) return ld(tc->color);
    const bhrt_texmap *m = &s->d->texmaps[tc->texmap];
    v3 t = m->texture >= 0 ? texture_sample(s, &s->d->textures[m->texture], transform_to(&m->xf, uvw)) : V(0, 0, 0);
    return mulv(ld(tc->color), t);
}
/* TexturedColor::Sample(uvw, duvw) / TextureMap::Sample(uvw, duvw) (scene.h:372-380,411); `color` is
 * passed separately because Shade() swaps in the Fresnel-boosted specular colour (MtlBlinn.cpp:111-113) */
static v3 texcolor_sample_filtered(const scene_t *s, v3 color, int texmap, v3 uvw, const v3 *duvw)
{
    if (texmap < 0) return color;
    const bhrt_texmap *m = &s->d->texmaps[texmap];
    if (m->texture < 0) return mulv(color, V(0, 0, 0));
    v3 u = transform_to(&m->xf, uvw);
    v3 dd[2];
    dd[0] = sub(transform_to(&m->xf, add(duvw[0], uvw)), u);
    dd[1] = sub(transform_to(&m->xf, add(duvw[1], uvw)), u);
    return mulv(color, texture_sample_filtered(s, &s->d->textures[m->texture], u, dd));
}
/* TexturedColor::SampleEnvironment (scene.h:414-420) */
static v3 sample_environment(const scene_t *s, v3 dir)
{
    float z = asinf(-dir.z) / PI_F + 0.5f;
    float x = dir.x / (fabsf(dir.x) + fabsf(dir.y));
    float y = dir.y / (fabsf(dir.x) + fabsf(dir.y));
    v3 uvw = add(V(0.5f, 0.5f, 0.0f), mulf(add(mulf(V(0.5f, 0.5f, 0), x), mulf(V(-0.5f, 0.5f, 0), y)), z));
    return texcolor_sample(s, &s->d->environment, uvw);
}

/* ---- caustic photon map: gather (cyPhotonMap.h) ---------------------------------------------- */
#define MAX_PHOTONS_IN_AREA 1000      /* MtlBlinn.cpp:28 */
#define MAX_AREA 0.5f                 /* MtlBlinn.cpp:29 (a double literal, passed as float radius) */
/* PhotonMap::Photon (cyPhotonMap.h:72-90), 24 bytes; photon i (1-based, photons[i]) = record i - 1 of the dump */
typedef struct { float pos[3]; float power; uint8_t rgb[3]; uint8_t planeAndDirZ; int16_t dirX, dirY; } photon_t;
static const photon_t *photon_at(const scene_t *s, int index) { return (const photon_t *)(s->photons + 24 * (size_t)(index - 1)); }
/* Photon::GetDirection (cyPhotonMap.h:192-214), with its dirX*dirX + dirY - dirY slip (:196) */
static v3 photon_direction(const photon_t *p)
{
    v3 dir;
    dir.x = (float)p->dirX / (float)0x7FFF;
    dir.y = (float)p->dirY / (float)0x7FFF;
    int dirXY2 = p->dirX * p->dirX + p->dirY - p->dirY;
    if (dirXY2 > 0x3FFF0001) dirXY2 = 0x3FFF0001;
    int dirZ2 = 0x3FFF0001 - dirXY2;
    int dirZ = 0, place = 0x40000000, remainder = dirZ2;
    while (place > remainder) place = place >> 2;
    while (place) {
        if (remainder >= dirZ + place) { remainder = remainder - dirZ - place; dirZ = dirZ + (place << 1); }
        dirZ = dirZ >> 1;
        place = place >> 2;
    }
    dir.z = (float)dirZ / (float)0x7FFF;
    if (p->planeAndDirZ & 0x8) dir.z = -dir.z;
    return dir;
}
typedef struct { v3 pos, normal; int maxPhotons, found; float *dist2; const photon_t **photon; } nearest_t;
static float comp(v3 a, int axis) { return axis == 0 ? a.x : (axis == 1 ? a.y : a.z); }
/* PhotonMap::LocatePhotons (cyPhotonMap.h:421-498), ellipticity 1 (normScale 0) */
static void locate_photons(const scene_t *s, nearest_t *np, int index)
{
    const photon_t *p = photon_at(s, index);
    int axis = p->planeAndDirZ & 0x3;
    if (index < s->half_photons) {
        float dist = comp(np->pos, axis) - p->pos[axis];
        if (dist > 0) {
            locate_photons(s, np, 2 * index + 1);
            if (dist * dist < np->dist2[0]) locate_photons(s, np, 2 * index);
        } else {
            locate_photons(s, np, 2 * index);
            if (dist * dist < np->dist2[0]) locate_photons(s, np, 2 * index + 1);
        }
    }
    v3 dif = sub(ld(p->pos), np->pos);
    float dist2 = dot(dif, dif);
    if (dist2 < np->dist2[0]) {
        if (dot(photon_direction(p), np->normal) >= 0) return;
        if (np->found < np->maxPhotons) {
            np->found++;
            np->dist2[np->found] = dist2;
            np->photon[np->found] = p;
            if (np->found == np->maxPhotons) {          /* build a heap; dist2[0] is NOT updated here (:462-478) */
                int half_found = np->found >> 1;
                for (int k = half_found; k >= 1; k--) {
                    int parent = k;
                    const photon_t *tp = np->photon[k];
                    float td2 = np->dist2[k];
                    while (parent <= half_found) {
                        int j = parent + parent;
                        if (j < np->found && np->dist2[j] < np->dist2[j + 1]) j++;
                        if (td2 >= np->dist2[j]) break;
                        np->dist2[parent] = np->dist2[j];
                        np->photon[parent] = np->photon[j];
                        parent = j;
                    }
                    np->photon[parent] = tp;
                    np->dist2[parent] = td2;
                }
            }
        } else {
            int parent = 1, j = 2;
            while (j <= np->found) {
                if (j < np->found && np->dist2[j] < np->dist2[j + 1]) j++;
                if (dist2 > np->dist2[j]) break;
                np->dist2[parent] = np->dist2[j];
                np->photon[parent] = np->photon[j];
                parent = j;
                j <<= 1;
            }
            np->photon[parent] = p;
            np->dist2[parent] = dist2;
            np->dist2[0] = np->dist2[1];
        }
    }
}
/* PhotonMap::EstimateIrradiance<1000> (cyPhotonMap.h:332-382), constant filter */
static void estimate_irradiance(const scene_t *s, v3 pos, v3 normal, float radius, v3 *irrad, v3 *direction)
{
    *irrad = V(0, 0, 0); *direction = V(0, 0, 0);
    float found_dist2[MAX_PHOTONS_IN_AREA + 1];
    const photon_t *found_photon[MAX_PHOTONS_IN_AREA + 1];
    nearest_t np;
    np.pos = pos; np.normal = normal; np.maxPhotons = MAX_PHOTONS_IN_AREA; np.found = 0;
    np.dist2 = found_dist2; np.photon = found_photon;
    np.dist2[0] = radius * radius;
    locate_photons(s, &np, 1);
    for (int i = 1; i <= np.found; i++) {
        const photon_t *p = np.photon[i];
        v3 power = mulf(V(p->rgb[0] / 255.0f, p->rgb[1] / 255.0f, p->rgb[2] / 255.0f), p->power);     /* GetPower (:83) */
        float filter = 1;
        *irrad = add(*irrad, mulf(power, filter));
        *direction = add(*direction, mulf(photon_direction(p), filter * p->power));
    }
    if (np.found > 0) {
        float area = PI_F * np.dist2[0];
        if (area > 0) { float one_over_area = 1.0f / area; *irrad = mulf(*irrad, one_over_area); }
        *direction = divf(*direction, len(*direction));
    }
}

/* ---- shading -------------------------------------------------------------------------------- */
static int all_white(v3 c) { return c.x >= 1 && c.y >= 1 && c.z >= 1; }
static const v3 NAN_PURPLE = { 1.0f, 0.0f, 1.0f };                  /* cyColor.h:130 */
static v3 shade(scene_t *s, ray_t ray, const hit_t *h, int bounceCount);
/* one top-level recursive() call (Main.cpp:389) */
static int trace(scene_t *s, ray_t ray, hit_t *h, int hitSide)
{
    int bHit = 0;
    hit_init(h);
    s->rays_closest++;
    recursive(s, -1, ray, h, &bHit, hitSide);
    return bHit && h->node >= 0;
}

/* HandleRayWhenRefractionRayOut (MtlBlinn.cpp:544-596), refractionGlossiness = 0 (D-4) */
static ray_t ray_leaving(ray_t inRay, const hit_t *in, float ior, int *toOut)
{
    v3 vN = in->N, vV = neg(inRay.dir);
    float cosPhi1 = dot(vV, neg(vN));
    float sinPhi1 = sqrtf(1 - cosPhi1 * cosPhi1);
    float sinPhi2 = ior * sinPhi1;
    ray_t o;
    if (sinPhi2 <= 1) {
        float cosPhi2 = sqrtf(1 - sinPhi2 * sinPhi2);
        v3 vTn = mulf(vN, cosPhi2);
        v3 vTp = mulf(norm(cross(vN, cross(vN, vV))), sinPhi2);
        v3 vT = add(vTn, vTp);
        o.dir = norm(norm(vT));
        o.p = add(in->p, mulf(vN, BIAS));
        *toOut = 1;
    } else {
        o.dir = sub(mulf(vN, -2 * cosPhi1), vV);
        o.p = sub(in->p, mulf(vN, BIAS));
        *toOut = 0;
    }
    return o;
}
/* RefractionOut (MtlBlinn.cpp:520-542) */
static v3 refraction_out(scene_t *s, ray_t outRay, v3 absorption, v3 refraction, int bounceCount)
{
    v3 out;
    hit_t h;
    if (trace(s, outRay, &h, HIT_FRONT)) {
        v3 af = V(powf(EULER_N, -absorption.x * h.z), powf(EULER_N, -absorption.y * h.z), powf(EULER_N, -absorption.z * h.z));
        out = mulv(mulv(refraction, af), shade(s, outRay, &h, bounceCount));
    } else out = mulv(refraction, sample_environment(s, outRay.dir));
    clamp_white(&out);
    return out;
}
/* RefractionRecusive (MtlBlinn.cpp:476-518) */
static v3 refraction_recursive(scene_t *s, v3 refraction, float ior, v3 vT, v3 p, v3 vN, v3 absorption, int bounceCount)
{
    ray_t in = { sub(p, mulf(vN, BIAS)), vT };
    hit_t h;
    if (!trace(s, in, &h, HIT_FRONT_AND_BACK)) return NAN_PURPLE;
    v3 c = V(0, 0, 0);
    if (!h.front) {
        int out;
        ray_t next = ray_leaving(in, &h, ior, &out);
        if (out) c = refraction_out(s, next, absorption, refraction, bounceCount);
        else if (bounceCount <= 0) c = V(0, 0, 0);                  /* total internal reflection, budget spent */
        else c = refraction_recursive(s, refraction, ior, next.dir, h.p, h.N, absorption, bounceCount - 1);
    } else c = shade(s, in, &h, bounceCount);
    clamp_white(&c);
    return c;
}
/* PathTracing_Refraction (MtlBlinn.cpp:438-473), refractionGlossiness = 0 */
static v3 refraction_term(scene_t *s, v3 refraction, v3 absorption, float ior, const hit_t *h, float cosPhi1, v3 vN, v3 vV, int bounceCount)
{
    v3 c = V(0, 0, 0);
    if (bounceCount <= 0) return c;
    if (!(refraction.x == 0.0f && refraction.y == 0.0f && refraction.z == 0.0f)) {
        float sinPhi1 = sqrtf(1 - cosPhi1 * cosPhi1);
        float sinPhi2 = sinPhi1 / ior;
        float cosPhi2 = sqrtf(1 - sinPhi2 * sinPhi2);
        v3 vTn = mulf(vN, -cosPhi2);
        v3 vTp = mulf(norm(cross(vN, cross(vN, vV))), sinPhi2);
        v3 vT = norm(add(vTn, vTp));
        c = refraction_recursive(s, refraction, ior, norm(vT), h->p, vN, absorption, bounceCount);
    }
    clamp_white(&c);
    return c;
}

/* MtlBlinn::Shade (Materials/Blinn/MtlBlinn.cpp:89-138) in D-mode: GI is switched off (D-2),
 * PathTracing_DiffuseNSpecular (:304-351) loops over every light with its exact direction (D-3) */
static v3 shade_body(scene_t *s, ray_t ray, const hit_t *h, int bounceCount);
static v3 shade(scene_t *s, ray_t ray, const hit_t *h, int bounceCount)
{
    if (++s->depth > s->max_depth) s->max_depth = s->depth;
    v3 c = shade_body(s, ray, h, bounceCount);
    s->depth--;
    return c;
}
static v3 shade_body(scene_t *s, ray_t ray, const hit_t *h, int bounceCount)
{
    /* Node::GetMaterial()->Shade: a MultiMtl dispatches on hInfo.mtlID (materials.h:71; its sub-materials are contiguous in the
     * flattened table, bhrt.h), which stays 0 in the reference as it is */
    const bhrt_material *M = &s->d->materials[s->d->nodes[h->node].material];
    if ((M->flags & BHRT_MTL_MULTI) && h->mtlID > 0 && h->mtlID < BHRT_MTL_SUBCOUNT(M->flags)) M += h->mtlID;
    v3 vN = norm(h->N);
    v3 vV = norm(sub(ray.p, h->p));
    float cosPhi1 = dot(vN, vV);
    if (cosPhi1 > 1) cosPhi1 = 1;
    if (cosPhi1 <= 0) cosPhi1 = 0;
    float R0 = (float)pow((double)((1 - M->ior) / (1 + M->ior)), 2.0);
    float F = (float)(R0 + (1 - R0) * pow((double)(1 - cosPhi1), 5.0));
    v3 newSpecular = add(ld(M->specular.color), mulf(ld(M->refraction), F));
    clamp_white(&newSpecular);
    v3 out = V(0, 0, 0);
    out = add(out, refraction_term(s, mulf(ld(M->refraction), 1 - F), ld(M->absorption), M->ior, h, cosPhi1, vN, vV, bounceCount));
    if (all_white(out)) return out;
    v3 D = V(0, 0, 0);
    for (int li = 0; li < s->n_lights; li++) {
        v3 vL = neg(norm(light_direction(&s->lights[li], h->p)));
        float cosTheta = dot(vL, vN);
        if (cosTheta > 0) {
            v3 vH = norm(add(vL, vV));
            v3 irrad = illuminate(s, &s->lights[li], h->p);
            v3 kd = texcolor_sample_filtered(s, ld(M->diffuse.color), M->diffuse.texmap, h->uvw, h->duvw);
            v3 ks = texcolor_sample_filtered(s, newSpecular, M->specular.texmap, h->uvw, h->duvw);
            v3 brdf = add(mulf(kd, cosTheta), mulf(ks, powf(dot(vH, vN), M->glossiness)));
            D = add(D, mulv(irrad, brdf));
        }
    }
    if (s->photons && s->n_photons > 0) {
        /* "Adding caustics for direct color" (:329-342) */
        v3 vL, irr;
        estimate_irradiance(s, h->p, h->N, MAX_AREA, &irr, &vL);
        float cosTheta = -dot(vL, vN);
        if (cosTheta > 0) {
            v3 vH = norm(add(vL, vV));
            v3 kd = texcolor_sample_filtered(s, ld(M->diffuse.color), M->diffuse.texmap, h->uvw, h->duvw);
            v3 ks = texcolor_sample_filtered(s, newSpecular, M->specular.texmap, h->uvw, h->duvw);
            v3 brdf = add(kd, divf(mulf(ks, powf(dot(vH, vN), M->glossiness)), cosTheta));
            D = add(D, mulv(brdf, irr));
        }
    }
    clamp_white(&D);                                            /* :343 */
    if (D.x != D.x) D = V(0, 0, 0);                             /* :344-347 */
    out = add(out, D);
    if (all_white(out)) return out;
    if (out.x != out.x) out = NAN_PURPLE;                       /* :133-136 */
    return out;
}

/* Color24::FloatToByte (cyColor.h:271-272) */
static uint8_t float_to_byte(float r)
{
    int v = (int)(r * 255 + 0.5f);
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

int bhrt_oracle_in_scope(const bhrt_scene_desc *d)
{
    for (int i = 0; i < d->n_lights; i++) if (d->lights[i].kind == BHRT_LIGHT_POINT && d->lights[i].size != 0) return 0;
    return 1;
}

/* BeginRender() + PathTracing() (Main.cpp:178-242,143-172) in D-mode.  Outputs: hit_id[H*W] (node
 * index in DFS order, -1 = miss), z[H*W], linear[H*W*3], rgb8[H*W*3], rays[2] = closest, shadow.
 * tir_budget = INTERNAL_REFLECTION_BOUNCE (Main.cpp:41; 16 in the reference). */
static int g_last_max_depth;
/* deepest Shade() recursion of the last frame rendered (not thread-safe; the reference has no bound on it) */
int bhrt_oracle_last_max_depth(void) { return g_last_max_depth; }

static int render_det(const bhrt_scene_desc *d, int tir_budget, int32_t *hit_id, float *z, float *linear, uint8_t *rgb8, uint64_t *rays, uint32_t *pixel_rays,
                      const uint8_t *photons, int n_photons)
{
    if (!d || !bhrt_oracle_in_scope(d)) return -2;
    scene_t s;
    memset(&s, 0, sizeof(s));
    s.d = d;
    s.photons = photons; s.n_photons = n_photons; s.half_photons = n_photons / 2 - 1;       /* cyPhotonMap.h:257 */
    int n = d->n_nodes;
    s.first_child = (int *)malloc(sizeof(int) * (n + 1));
    s.next_sibling = (int *)malloc(sizeof(int) * (n + 1));
    int *last_child = (int *)malloc(sizeof(int) * (n + 1));
    int last_root = -1;
    for (int i = 0; i < n; i++) { s.first_child[i] = -1; s.next_sibling[i] = -1; last_child[i] = -1; }
    for (int i = 0; i < n; i++) {             /* children in document order = DFS pre-order of the table */
        int p = d->nodes[i].parent;
        if (p < 0) { if (last_root >= 0) s.next_sibling[last_root] = i; last_root = i; }
        else { if (last_child[p] >= 0) s.next_sibling[last_child[p]] = i; else s.first_child[p] = i; last_child[p] = i; }
    }
    free(last_child);

    /* camera frame (Main.cpp:179-192), float with a double tan() */
    const bhrt_camera *c = &d->camera;
    int W = c->width, H = c->height;
    float aor = W / (float)H;
    float tan_h_pov = (float)tan(c->fov / 2 * PI_D / 180.0);
    float l = c->focaldist;
    float h = 2 * l * tan_h_pov;
    float w = aor * h;
    v3 camZ = neg(ld(c->dir)), camY = ld(c->up), camX = cross(camY, camZ), pos = ld(c->pos);
    v3 topLeft = sub(add(sub(pos, mulf(camZ, l)), divf(mulf(camY, h), 2)), divf(mulf(camX, w), 2));
    v3 dd_x = divf(mulf(camX, w), (float)W), dd_y = divf(mulf(camY, h), (float)H);
    s.dd_x = dd_x; s.dd_y = dd_y;

    /* the elliptic filter offsets of Texture::Sample (scene.h:323-329) */
    for (int i = 1; i < TEXTURE_SAMPLE_COUNT; i++) {
        float x = halton(i, 2), y = halton(i, 3);
        float r = sqrtf(x) * 0.5f;
        s.tex_x[i] = r * sinf(y * PI_F * 2);
        s.tex_y[i] = r * cosf(y * PI_F * 2);
    }

    /* CalculateLightsIntensity (Main.cpp:112-123): lights sorted ascending by Gray() (insertion sort = std::sort below 16 elements) */
    bhrt_light lights[BHRT_MAX_LIGHTS];
    int nl = d->n_lights;
    for (int i = 0; i < nl; i++) {
        bhrt_light v = d->lights[i];
        int k = i;
        if (i > 0 && gray(&v) < gray(&lights[0])) { for (k = i; k > 0; k--) lights[k] = lights[k - 1]; lights[0] = v; continue; }
        while (k > 0 && gray(&v) < gray(&lights[k - 1])) { lights[k] = lights[k - 1]; k--; }
        lights[k] = v;
    }
    s.lights = lights; s.n_lights = nl;

    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            /* pixel corner: (i + 1/2) with integer 1/2 == 0 (Main.cpp:145); direction not normalised (:153) */
            v3 pc = sub(add(topLeft, mulf(dd_x, (float)i)), mulf(dd_y, (float)j));
            ray_t ray = { pos, sub(pc, pos) };
            hit_t hit;
            const uint64_t c0 = s.rays_closest, s0 = s.rays_shadow;
            int bHit = trace(&s, ray, &hit, HIT_FRONT);
            v3 col;
            if (bHit) col = shade(&s, ray, &hit, tir_budget);
            else col = texcolor_sample(&s, &d->background, V((float)i / W, (float)j / H, 0.0f));      /* Main.cpp:165-167 */
            size_t p = (size_t)j * W + i;
            hit_id[p] = bHit ? hit.node : -1;
            z[p] = hit.z;
            linear[3 * p] = col.x; linear[3 * p + 1] = col.y; linear[3 * p + 2] = col.z;
            const float ig = 1 / 2.2f;                          /* Main.cpp:128,220-227 */
            rgb8[3 * p] = float_to_byte(powf(col.x, ig));
            rgb8[3 * p + 1] = float_to_byte(powf(col.y, ig));
            rgb8[3 * p + 2] = float_to_byte(powf(col.z, ig));
            if (pixel_rays) { pixel_rays[2 * p] = (uint32_t)(s.rays_closest - c0); pixel_rays[2 * p + 1] = (uint32_t)(s.rays_shadow - s0); }
        }
    rays[0] = s.rays_closest; rays[1] = s.rays_shadow;
    g_last_max_depth = s.max_depth;
    free(s.first_child); free(s.next_sibling);
    return 0;
}

int bhrt_oracle_render_det2(const bhrt_scene_desc *d, int tir_budget, int32_t *hit_id, float *z, float *linear, uint8_t *rgb8, uint64_t *rays)
{
    return render_det(d, tir_budget, hit_id, z, linear, rgb8, rays, NULL, NULL, 0);
}
/* the same, plus the (closest, shadow) ray counts of every pixel: pixel_rays[H*W*2] */
int bhrt_oracle_render_det3(const bhrt_scene_desc *d, int tir_budget, int32_t *hit_id, float *z, float *linear, uint8_t *rgb8, uint64_t *rays,
                            uint32_t *pixel_rays)
{
    return render_det(d, tir_budget, hit_id, z, linear, rgb8, rays, pixel_rays, NULL, 0);
}

/* the deterministic frame WITH the caustic gather (-DUSE_PhotonMap build of the reference): `photons` =
 * n_photons 24-byte records exactly as the reference dumps them (Main.cpp:380-385: powers scaled, balanced
 * kd-tree in heap order with the split planes in planeAndDirZ) */
int bhrt_oracle_render_det_pm(const bhrt_scene_desc *d, int tir_budget, const uint8_t *photons, int n_photons, int32_t *hit_id, float *z,
                              float *linear, uint8_t *rgb8, uint64_t *rays)
{
    return render_det(d, tir_budget, hit_id, z, linear, rgb8, rays, NULL, photons, n_photons);
}

int bhrt_oracle_render_det(const bhrt_scene_desc *d, int32_t *hit_id, float *z, float *linear, uint8_t *rgb8, uint64_t *rays)
{
    return bhrt_oracle_render_det2(d, 16, hit_id, z, linear, rgb8, rays);
}
