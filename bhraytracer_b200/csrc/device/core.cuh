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
struct MeshAccel {
    float root_box[6];
    uint32_t root_word, pad;
    int64_t pairs_off, tris_off;            // byte offsets into SceneView::accel_data
    int64_t vn_off, vt_off, fn_off, ft_off; // byte offsets into SceneView::mesh_data (attributes of the final hit)
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
    float b1[6] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y };
    float b2[6] = { q1.z, q1.w, q2.x, q2.y, q2.z, q2.w };
    t1 = BHRT_BIGFLOAT; t2 = BHRT_BIGFLOAT;
    h1 = intersect_box(b1, r, t_max, t1);
    h2 = intersect_box(b2, r, t_max, t2);
    w1 = f2u(q3.x); w2 = f2u(q3.y);
}

#define BHRT_BVH_LEAF(d) (((d) & 0x80000000u) != 0)
#define BHRT_BVH_COUNT(d) ((((d) >> 28) & 0x7u) + 1)
#define BHRT_BVH_OFFSET(d) ((d) & 0x0fffffffu)
#define BHRT_BVH_CHILD(d) ((d) & 0x7fffffffu)
#define BHRT_BVH_STACK 64

// One traversal loop for both query kinds (the two heavy parts -- pair slab tests and triangle
// tests -- are shared, so a warp whose lanes mix closest-hit and shadow queries stays converged):
//
//  closest (any_hit == false): TriObj::IntersectRay + TraceBVHNode (reference TriObj.cpp:17-39,
//    192-270) made iterative.  The visiting order, the near/far rule, "the far child is skipped
//    only when the near child produced a hit and z <= tmin_far" and the traversal of box-missed
//    far children are all kept, so ties and rounding cases resolve as in the reference.
//  shadow (any_hit == true): TriObj::ShadowRecursive + TraceBVHShadow (TriObj.cpp:41-66,272-307):
//    child 1 before child 2, both visited when EITHER box is hit (boxes tested against BIGFLOAT),
//    stop at the FIRST leaf containing any front-face hit and range-check only that leaf's t.
//
// Returns: closest -> number of accepted triangles > 0 (z, front, mh updated);
//          shadow  -> the reference's bool.
BHRT_HD bool mesh_trace(const SceneView &s, const MeshAccel &A, const Ray &ray, bool any_hit, float t_max_shadow,
                        float &z, bool &front, int hitSide, MeshHit &mh)
{
    float tmin = -1;
    if (!intersect_box(A.root_box, ray, any_hit ? BHRT_BIGFLOAT : z, tmin)) return false;
    const PairNode *pairs = (const PairNode *)(s.accel_data + A.pairs_off);
    const TriRec *tris = (const TriRec *)(s.accel_data + A.tris_off);
    const float len_dir = length(ray.dir);
    uint32_t st_word[BHRT_BVH_STACK];
    float st_tmin[BHRT_BVH_STACK];
    int st_hits[BHRT_BVH_STACK];
    int sp = 0, hits = 0;
    uint32_t cur = A.root_word;
    for (;;) {
        bool descend = false;
        if (BHRT_BVH_LEAF(cur)) {
            uint32_t cnt = BHRT_BVH_COUNT(cur), off = BHRT_BVH_OFFSET(cur);
            if (any_hit) {
                float lz = BHRT_BIGFLOAT; bool lf; MeshHit lm;
                bool any = false;
                for (uint32_t i = 0; i < cnt; i++)
                    if (intersect_trirec(tris + off + i, ray, len_dir, lz, lf, HIT_FRONT, lm)) any = true;
                if (any) return lz > BHRT_BIAS && lz < t_max_shadow;
            } else {
                for (uint32_t i = 0; i < cnt; i++)
                    if (intersect_trirec(tris + off + i, ray, len_dir, z, front, hitSide, mh)) hits++;
            }
        } else {
            bool h1, h2; float t1, t2; uint32_t w1, w2;
            intersect_pair(pairs + cur, ray, any_hit ? BHRT_BIGFLOAT : z, h1, t1, h2, t2, w1, w2);
            if (h1 || h2) {
                uint32_t first, second; float tsecond;
                if (any_hit || t1 < t2) { first = w1; second = w2; tsecond = t2; }
                else { first = w2; second = w1; tsecond = t1; }
                if (sp < BHRT_BVH_STACK) { st_word[sp] = second; st_tmin[sp] = tsecond; st_hits[sp] = hits; sp++; }
                cur = first;
                descend = true;
            }
        }
        if (descend) continue;
        bool found = false;
        while (sp > 0) {
            sp--;
            if (!any_hit) {
                bool nearHit = (hits != st_hits[sp]);
                if (nearHit && !(z > st_tmin[sp])) continue;
            }
            cur = st_word[sp];
            found = true;
            break;
        }
        if (!found) break;
    }
    return hits > 0;
}

// ------------------------------------------------------------------------------------------
// A ray query: the ONE place the estimator meets the geometry.  kind CLOSEST = recursive()
// (reference Main.cpp:389-413, linear DFS over the scene graph); kind SHADOW = GenLight::Shadow /
// ShadowRayRecursive (reference Lights/GenLight.cpp:10-69, any-hit; the result is the OR over
// all nodes of the node's own test, so DFS pre-order is as good as the reference's post-order).
enum { Q_NONE = 0, Q_CLOSEST = 1, Q_SHADOW = 2 };
struct Query {
    Ray ray;
    int kind;
    int hitSide;        // CLOSEST
    float t_max;        // SHADOW
};

// nodes[] is in DFS pre-order; rs[d] is the ray in the space of the current node at depth d.
// Shadow quirks kept: sphere ignored when the origin is inside (GenLight.cpp:36-37), plane
// bounds tested with the PARENT-space ray (:54), no rayDz == 0 test (:50).
// Closest: per-candidate HitInfo fields that only the winner needs (sphere uv with atan2/asin,
// plane ray differentials with pow, mesh normals/uv) are derived once after the loop from the
// winner's local ray -- the same expressions on the same inputs as the reference evaluates at
// accept time.
BHRT_HD bool scene_trace(const SceneView &s, const Query &q, Hit &h)
{
    const bool shadow = (q.kind == Q_SHADOW);
    Ray rs[BHRT_MAX_DEPTH + 1];
    rs[0] = q.ray;
    h.init();
    Ray wr; wr.p = V3(0, 0, 0); wr.dir = V3(0, 0, 0);   // winner's local ray
    MeshHit mh; mh.face = 0; mh.a0 = mh.a1 = mh.a2 = 0;
    for (int k = 0; k < s.n_nodes; k++) {
        const bhrt_node &nd = s.nodes[k];
        const int d = nd.depth;
        const Ray &pr = rs[d - 1];
        Ray r = to_node(nd.xf, pr);
        if (d < s.max_depth) rs[d] = r;
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
            bool r_ = mesh_trace(s, s.accels[nd.mesh], r, shadow, q.t_max, h.z, h.front, q.hitSide, mh);
            if (shadow) { if (r_) return true; }
            else hit = r_;
        }
        if (hit) { h.node = k; wr = r; }
    }
    if (shadow || h.node < 0) return false;

    // ---- finish the winner's HitInfo in its local space
    const bhrt_node &nd = s.nodes[h.node];
    V3 lp, lN;
    h.duvw0 = V3(0, 0, 0); h.duvw1 = V3(0, 0, 0);
    if (nd.kind == BHRT_OBJ_SPHERE) {
        lp = wr.p + h.z * wr.dir;                           // Sphere.cpp:52-54
        lN = lp;
        V3 d = normalized(lN);
        // 0.5f + atan2(d.y,d.x)/(2*PI): float atan2f promoted to double, PI is a double literal
        h.uvw.x = (float)(0.5f + atan2f(d.y, d.x) / (2 * BHRT_PI_D));
        h.uvw.y = (float)(0.5f - asinf(d.z) / (BHRT_PI_D));
        // uvw.z keeps whatever the default-constructed Vec3f held in the reference (uninitialised
        // there; never read by 2-D textures).  We define it as 0.
        h.uvw.z = 0.f;
    } else if (nd.kind == BHRT_OBJ_PLANE) {
        const float t = h.z;
        lp = wr.p + t * wr.dir;
        lN = V3(0, 0, 1);
        h.uvw = V3((1 + lp.x) / 2.f, (1 + lp.y) / 2.f, 0.f);
        // ray differentials from the PRIMARY-ray pixel steps dd_x, dd_y, for every ray (Plane.cpp:3-4,51-70)
        V3 nd_ = normalized(wr.dir);
        float scaled_t = length(t * wr.dir);
        float ndnd = dot(nd_, nd_);
        float den = powf(ndnd, 1.5f);
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
        mesh_attributes(s, s.accels[nd.mesh], mh.face, mh.a0, mh.a1, mh.a2, lN, h.uvw);
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
// an early-out at every level, not a linear path, so it is evaluated with an
// explicit frame stack and a post-order resolve.  It is written as a coroutine around ONE ray
// query: sample_advance() runs until the estimator needs a ray, leaves it in Lane::q and
// returns; the caller traces it (scene_trace) and calls sample_advance() again.  In the kernel
// every lane of a warp therefore meets the traversal code together, whatever stage of whatever
// recursion level its sample is in.
struct Frame {
    V3 p, vN, vV, uvw, duvw0, duvw1;    // shading point
    V3 out;                              // colour accumulated so far at this level
    V3 mult;                             // factor for the pending child's value
    V3 newSpec;                          // Fresnel-boosted specular colour (MtlBlinn.cpp:110-113)
    int mtl, bounce, gi;
    int stage;
    bool useSpec;                        // lobe chosen by the GI sample
};
enum { ST_REFR = 0, ST_REFR_IN_WAIT = 1, ST_REFR_OUT_WAIT = 2, ST_REFR_WAIT = 3, ST_GI = 4, ST_GI_TRACE_WAIT = 5, ST_GI_WAIT = 6,
       ST_DIRECT = 7, ST_SHADOW_WAIT = 8 };
#define BHRT_MAX_FRAMES 24

struct ShadeCfg {
    int mode;           // BHRT_MODE_*
    int gi_depth, tir_budget;
};

struct Lane {
    Frame fr[BHRT_MAX_FRAMES];
    int sp;                 // frames in use; 0 = waiting for the primary ray
    Query q;                // the pending ray query
    Hit hit;                // its answer (closest) ...
    bool hit_ok;            // ... and whether anything was hit / the shadow ray is blocked
    Rng rng;
    V3 ret;                 // value returned by the child frame that just completed
    // scratch of the top frame's wait states (a child frame runs to completion in between, so
    // one set per lane is enough)
    V3 sv;                  // refraction: (1-F)*refraction      direct: colour summed over lights
    V3 sw;                  // direct: the sampled light direction vL
    float sf;               // refraction: glossiness of the refraction lobe
    int si;                 // refraction: remaining internal-reflection budget      direct: light index
    int hit_id; float hit_z;    // primary AOVs of this sample
};

BHRT_HD void frame_begin(Frame &f, const SceneView &s, V3 ray_p, const Hit &h, int bounce, int gi)
{
    f.p = h.p; f.uvw = h.uvw; f.duvw0 = h.duvw0; f.duvw1 = h.duvw1;
    f.vN = normalized(h.N);
    f.vV = normalized(ray_p - h.p);
    f.mtl = s.nodes[h.node].material;
    f.bounce = bounce; f.gi = gi;
    f.out = V3(0, 0, 0);
    f.stage = ST_REFR;
}

BHRT_HD void issue_closest(Lane &L, const Ray &r, int hitSide, RayCounters &rc)
{
    L.q.ray = r; L.q.kind = Q_CLOSEST; L.q.hitSide = hitSide; L.q.t_max = 0;
    rc.closest++;
}

// Start one camera sample (the body of the loop in PathTracing(), Main.cpp:150-169)
BHRT_HD void sample_begin(Lane &L, const Ray &primary, RayCounters &rc)
{
    L.sp = 0;
    issue_closest(L, primary, HIT_FRONT, rc);
}

// Shade epilogue shared by the two ways the direct term ends (MtlBlinn.cpp:343-350,131-136)
BHRT_HD void direct_finish(Frame &f, V3 D)
{
    clamp_white(D);
    if (is_nan(D.x)) D = V3(0, 0, 0);
    f.out = f.out + D;
    if (!saturated(f.out) && is_nan(f.out.x)) f.out = V3(1.0f, 0.0f, 1.0f);
}

// Blinn BRDF * cos of the direct term (MtlBlinn.cpp:319-327) for light direction vL
BHRT_HD V3 direct_brdf(const SceneView &s, const Frame &f, const bhrt_material &M, V3 vL)
{
    float cosTheta = dot(vL, f.vN);
    V3 vH = normalized(vL + f.vV);
    V3 kdT = texcolor_sample_filtered(s, ld3(M.diffuse.color), M.diffuse.texmap, f.uvw, f.duvw0, f.duvw1);
    V3 ksT = texcolor_sample_filtered(s, f.newSpec, M.specular.texmap, f.uvw, f.duvw0, f.duvw1);
    return kdT * cosTheta + ksT * powf(dot(vH, f.vN), M.glossiness);
}

// Consume the answer to L.q (L.hit / L.hit_ok) and run the estimator until it needs the next
// ray (returns false, query in L.q) or the sample is finished (returns true, colour in `color`).
BHRT_HD_NOINLINE bool sample_advance(Lane &L, const SceneView &s, const ShadeCfg &cfg, const PhotonGather &pg, float bgu, float bgv,
                                     RayCounters &rc, V3 &color)
{
    const bool det = (cfg.mode == BHRT_MODE_DET);
    Rng &rng = L.rng;
    if (L.sp == 0) {
        L.hit_id = L.hit_ok ? L.hit.node : -1;
        L.hit_z = L.hit.z;
        if (!L.hit_ok) { color = texcolor_sample(s, s.background, V3(bgu, bgv, 0.0f)); return true; }   // Main.cpp:166-167
        frame_begin(L.fr[0], s, L.q.ray.p, L.hit, cfg.tir_budget, cfg.gi_depth);
        L.sp = 1;
    }
    for (;;) {
        Frame &f = L.fr[L.sp - 1];
        const bhrt_material &M = s.materials[f.mtl];
        bool push_child = false;    // a child frame was pushed at fr[sp]
        bool done = false;          // this frame finished, value in f.out
        bool refr_resolved = false; V3 R(0, 0, 0);      // refraction term known
        bool gi_resolved = false; V3 G(0, 0, 0);        // GI term known

        switch (f.stage) {
        case ST_REFR: {
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
                f.stage = ST_REFR_IN_WAIT;
                return false;
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
                    f.stage = ST_REFR_OUT_WAIT;
                    return false;
                }
                if (L.si <= 0) { refr_resolved = true; break; }
                L.si--;
                Ray r2; r2.dir = nx.dir; r2.p = L.hit.p - L.hit.N * BHRT_BIAS;
                issue_closest(L, r2, HIT_FRONT_AND_BACK, rc);
                return false;
            }
            f.mult = V3(1, 1, 1);
            if (L.sp < BHRT_MAX_FRAMES) {
                frame_begin(L.fr[L.sp], s, rin.p, L.hit, L.si, f.gi - 1);
                push_child = true;
                f.stage = ST_REFR_WAIT;
            } else refr_resolved = true;
            break;
        }
        case ST_REFR_OUT_WAIT: {
            const Ray nx = L.q.ray;
            if (L.hit_ok) {
                V3 ab = ld3(M.absorption);
                float hz = L.hit.z;
                V3 af(powf(BHRT_EULER, -ab.x * hz), powf(BHRT_EULER, -ab.y * hz), powf(BHRT_EULER, -ab.z * hz));
                f.mult = L.sv * af;
                if (L.sp < BHRT_MAX_FRAMES) {
                    frame_begin(L.fr[L.sp], s, nx.p, L.hit, L.si, f.gi - 1);
                    push_child = true;
                    f.stage = ST_REFR_WAIT;
                } else refr_resolved = true;
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
            if (det || f.gi < 0) { f.stage = ST_DIRECT; L.si = 0; L.sv = V3(0, 0, 0); break; }
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
            issue_closest(L, g, HIT_FRONT, rc);
            f.stage = ST_GI_TRACE_WAIT;
            return false;
        }
        case ST_GI_TRACE_WAIT: {
            const Ray g = L.q.ray;
            if (L.hit_ok) {
                if (fabsf(L.hit.z) > BHRT_BIAS && L.sp < BHRT_MAX_FRAMES) {
                    frame_begin(L.fr[L.sp], s, g.p, L.hit, f.bounce, f.gi - 1);
                    push_child = true;
                    f.stage = ST_GI_WAIT;
                } else gi_resolved = true;      // indirect = black, black * albedo = black
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
                        if (illuminate_begin(Lt, f.p, (Rng *)0, L.q, irrad)) { rc.shadow++; f.stage = ST_SHADOW_WAIT; issued = true; break; }
                        L.sv = L.sv + irrad * direct_brdf(s, f, M, vL);
                    }
                    L.si++;
                }
                if (issued) return false;
                direct_finish(f, L.sv);
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
                        if (illuminate_begin(Lt, f.p, &rng, L.q, irrad)) { rc.shadow++; f.stage = ST_SHADOW_WAIT; return false; }
                        L.sv = L.sv + irrad * direct_brdf(s, f, M, vL);
                    }
                }
                direct_finish(f, L.sv);
                done = true;
            }
            break;
        }
        case ST_SHADOW_WAIT: {
            const bhrt_light &Lt = s.lights[L.si];
            V3 irrad = illuminate_end(Lt, f.p, L.hit_ok);
            L.sv = L.sv + irrad * direct_brdf(s, f, M, L.sw);
            if (det) { L.si++; f.stage = ST_DIRECT; }
            else { direct_finish(f, L.sv); done = true; }
            break;
        }
        }

        if (refr_resolved) {
            clamp_white(R);
            f.out = f.out + R;
            if (saturated(f.out)) done = true; else f.stage = ST_GI;
        }
        if (gi_resolved) {
            if (is_nan(G.x)) G = V3(1.0f, 0.0f, 1.0f);
            clamp_white(G);
            f.out = f.out + G;
            if (saturated(f.out)) done = true;
            else { f.stage = ST_DIRECT; L.si = 0; L.sv = V3(0, 0, 0); }
        }
        if (push_child) { L.sp++; continue; }
        if (done) {
            L.ret = f.out;
            L.sp--;
            if (L.sp == 0) { color = L.ret; return true; }
        }
    }
}

// One camera sample, synchronously: the reference's PathTracing() loop body.  Used by the CPU
// harness (tests/emu); the kernel interleaves the same two calls across a warp.
BHRT_HD V3 trace_sample(Lane &L, const SceneView &s, const ShadeCfg &cfg, const PhotonGather &pg, const Ray &primary, float bgu, float bgv,
                        RayCounters &rc, int &hit_id, float &hit_z)
{
    V3 color(0, 0, 0);
    sample_begin(L, primary, rc);
    for (;;) {
        L.hit_ok = scene_trace(s, L.q, L.hit);
        if (sample_advance(L, s, cfg, pg, bgu, bgv, rc, color)) break;
    }
    hit_id = L.hit_id; hit_z = L.hit_z;
    return color;
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
