// wide.cuh -- warp-cooperative traversal of the 8-wide BVH: EIGHT LANES PER RAY.
//
// Replaces, for the mesh part of a ray query, the per-thread traversal of core.cuh (MeshTrav: one ray per
// thread over the binary PairNode tree -- TriObj::TraceBVHNode / TraceBVHShadow made iterative, reference
// Objects/TriObj/TriObj.cpp:192-307) in the fast build of the render kernel.  The tree is the reference's own
// cyBVH regrouped to <= 8 children per node and <= 8 triangles per leaf group (accel.hpp: build_wide); the
// semantics are the ones `wide_trav_run` (core.cuh) states for a single thread, and the CPU harness checks
// that statement against the binary traversal and the reference (tests/test_wide_bvh.py).
//
// A warp holds four rays ("octets": lanes 8k .. 8k+7).  Per step
//   internal node  lane j of the octet fetches child j -- box + word + aux = 32 bytes, two 128-bit loads, the
//                  eight lanes together read the node's 256 bytes -- and runs the slab test; one ballot gives
//                  the hit mask; closest-hit rays rank the hit children by entry distance with 7 shuffles
//                  (nearest first), shadow rays by slot (= the reference's leaf order); the best child becomes
//                  the current node, the others go on the octet's stack in SHARED memory in one go;
//   leaf group     lane j tests triangle j (48-byte TriRec, three 128-bit loads); closest: a 3-step butterfly
//                  picks the nearest accepted triangle; shadow: ballot -> first of the reference's leaves with a
//                  front-face hit -> butterfly min of that leaf's t (TriObj.cpp:272-291);
//   pop            the eight lanes look at the eight topmost stack entries at once and a ballot finds the
//                  first one that is still nearer than the current hit.
// Per-ray state is therefore shared by 8 lanes: a 64-entry stack per octet in shared memory (64 KB per CTA)
// instead of three 64-entry arrays per thread in local memory, and the dependent chain per step is one node
// fetch for up to 8 children.
#pragma once
#include <cuda_runtime.h>
#include "core.cuh"

namespace BHRT_NS {

#define WIDE_FULL 0xffffffffu

// per-octet record in shared memory: the deferred children and the rarely touched part of the job.
// Every lane of the octet writes the same values (the computation is octet-uniform), so no lane waits for another.
struct alignas(16) OctShared {
    uint2 stack[BHRT_WIDE_STACK];       // (word, closest: entry distance bits | shadow: aux)
    // the query's answer so far (MeshJob::h of core.cuh)
    float hz; int hnode; int hfront; uint32_t hface;
    float ha0, ha1, ha2; int pad_;
};

// the hot part of an octet's traversal, identical in its 8 lanes
struct OctTrav {
    Ray r;                  // the ray in the mesh node's space
    float len_dir;
    float z;                // closest: nearest accepted t so far (in/out); shadow: unused
    float t_max_shadow;
    uint32_t cur, aux;
    int sp;
    int hitSide;
    const WideNode *wide;
    const TriRec *tris;
    // winner so far inside this mesh
    uint32_t face; float a0, a1, a2;
    bool front, any_hit, done, result, hit_any;
};

__device__ __forceinline__ unsigned oct_bits(unsigned ballot, unsigned oshift) { return (ballot >> oshift) & 0xffu; }

// Pop for the octets with `want`: the first deferred entry (from the top) that is still worth visiting becomes
// the current node; none left -> the traversal is done.  Warp-convergent (all 32 lanes call it).
__device__ __forceinline__ void oct_pop(OctTrav &T, OctShared *S, bool want, unsigned sub, unsigned oshift)
{
    for (;;) {
        if (!__any_sync(WIDE_FULL, want)) return;
        uint2 e = make_uint2(0u, 0u);
        bool ok = false;
        __syncwarp();           // entries pushed by the other lanes of the octet
        if (want && (int)sub < T.sp) {
            e = S->stack[T.sp - 1 - (int)sub];
            ok = T.any_hit || __uint_as_float(e.y) < T.z;
        }
        const unsigned pm = oct_bits(__ballot_sync(WIDE_FULL, ok), oshift);
        const int f = pm ? __ffs(pm) - 1 : 0;
        const uint32_t w = __shfl_sync(WIDE_FULL, e.x, (int)(oshift + f)), a = __shfl_sync(WIDE_FULL, e.y, (int)(oshift + f));
        if (want) {
            if (pm) { T.cur = w; T.aux = a; T.sp -= f + 1; want = false; }
            else if (T.sp > 8) T.sp -= 8;
            else { T.sp = 0; T.done = true; T.result = T.hit_any; want = false; }
        }
    }
}

// One internal-node step for the octets with `act` (T.cur is a WideNode index).  Warp-convergent.
__device__ __forceinline__ void oct_internal(OctTrav &T, OctShared *S, bool act, unsigned sub, unsigned oshift)
{
    uint32_t w = BHRT_WIDE_EMPTY, ax = 0;
    float tmin = 0.f;
    bool hit = false;
    if (act) {
        const F4 *ch = reinterpret_cast<const F4 *>(&T.wide[T.cur].c[sub]);
        const F4 q0 = ld4(ch), q1 = ld4(ch + 1);
        w = f2u(q1.z); ax = f2u(q1.w);
        if (w != BHRT_WIDE_EMPTY) hit = intersect_box6(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, T.r, T.any_hit ? BHRT_BIGFLOAT : T.z, tmin);
    }
    const unsigned hm = oct_bits(__ballot_sync(WIDE_FULL, hit), oshift);
    // order key: entry distance (clamped at 0) with the low 3 mantissa bits replaced by the slot -> unique, ties by
    // slot; 8 ulps are taken off first, so the key never exceeds the distance and the cull `key < z` on pop stays
    // conservative
    uint32_t kb = f2u(tmin > 0.f ? tmin : 0.f);
    kb = kb >= 8u ? kb - 8u : 0u;
    const uint32_t key = (kb & ~7u) | sub;
    int rank = 0;
    if (T.any_hit) rank = __popc(hm & ((1u << sub) - 1u));
#pragma unroll
    for (int d = 1; d < 8; d++) {
        const uint32_t ok = __shfl_xor_sync(WIDE_FULL, key, d, 8);
        if (!T.any_hit && ((hm >> (sub ^ (unsigned)d)) & 1u) && ok < key) rank++;
    }
    const int n = __popc(hm);
    const unsigned first = oct_bits(__ballot_sync(WIDE_FULL, hit && rank == 0), oshift);
    const int src = (int)oshift + (first ? __ffs(first) - 1 : 0);
    const uint32_t nw = __shfl_sync(WIDE_FULL, w, src), na = __shfl_sync(WIDE_FULL, ax, src);
    bool want_pop = false;
    if (act) {
        if (n == 0) want_pop = true;
        else {
            if (hit && rank > 0) S->stack[T.sp + (n - 1 - rank)] = make_uint2(w, T.any_hit ? ax : key);
            T.sp += n - 1;
            T.cur = nw; T.aux = na;
        }
    }
    __syncwarp();
    oct_pop(T, S, want_pop, sub, oshift);
}

// One leaf-group step for the octets with `act` (T.cur is a leaf word).  Warp-convergent.
__device__ __forceinline__ void oct_leaf(OctTrav &T, OctShared *S, bool act, unsigned sub, unsigned oshift)
{
    const uint32_t cnt = BHRT_BVH_COUNT(T.cur), off = BHRT_BVH_OFFSET(T.cur);
    float t = BHRT_BIGFLOAT;
    bool th = false, fr = true;
    MeshHit mh; mh.face = 0; mh.a0 = mh.a1 = mh.a2 = 0.f;
    if (act && sub < cnt) {
        t = T.any_hit ? BHRT_BIGFLOAT : T.z;
        th = intersect_trirec(T.tris + off + sub, T.r, T.len_dir, t, fr, T.any_hit ? (int)HIT_FRONT : T.hitSide, mh);
    }
    const unsigned tm = oct_bits(__ballot_sync(WIDE_FULL, th), oshift);
    // shadow: only the first of the reference's leaves that holds a hit counts (TriObj.cpp:272-291)
    unsigned seg = 0xffu;
    if (T.any_hit && tm) {
        const int f = __ffs(tm) - 1;
        const unsigned starts = T.aux | 1u;
        const int b = 31 - __clz(starts & ((2u << f) - 1u));
        const unsigned rest = T.aux >> (f + 1);
        const int e = rest ? f + __ffs(rest) : (int)cnt;
        seg = ((1u << e) - 1u) & ~((1u << b) - 1u);
    }
    // butterfly: nearest hit of the octet (closest: ties go to the later record; shadow: inside the segment)
    float bt = (th && ((seg >> sub) & 1u)) ? t : BHRT_BIGFLOAT;
    int bs = (int)sub;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        const float ot = __shfl_xor_sync(WIDE_FULL, bt, d, 8);
        const int os = __shfl_xor_sync(WIDE_FULL, bs, d, 8);
        if (ot < bt || (ot == bt && os > bs)) { bt = ot; bs = os; }
    }
    const int src = (int)oshift + bs;
    const bool wfr = __shfl_sync(WIDE_FULL, (int)fr, src) != 0;
    const uint32_t wface = __shfl_sync(WIDE_FULL, mh.face, src);
    const float wa0 = __shfl_sync(WIDE_FULL, mh.a0, src), wa1 = __shfl_sync(WIDE_FULL, mh.a1, src), wa2 = __shfl_sync(WIDE_FULL, mh.a2, src);
    bool want_pop = false;
    if (act) {
        if (tm) {
            if (T.any_hit) { T.result = bt > BHRT_BIAS && bt < T.t_max_shadow; T.done = true; }
            else { T.z = bt; T.front = wfr; T.face = wface; T.a0 = wa0; T.a1 = wa1; T.a2 = wa2; T.hit_any = true; }
        }
        want_pop = !T.done;
    }
    oct_pop(T, S, want_pop, sub, oshift);
}

// Start the traversal of one mesh for the octet (octet-uniform): root box, then the root of the wide tree.
__device__ __forceinline__ void oct_begin(OctTrav &T, const SceneView &s, const MeshAccel &A, const Ray &ray, bool any_hit, float t_max_shadow,
                                          float z, bool front, int hitSide)
{
    T.r = ray; T.any_hit = any_hit; T.t_max_shadow = t_max_shadow; T.z = z; T.front = front; T.hitSide = hitSide;
    T.face = 0; T.a0 = T.a1 = T.a2 = 0.f;
    T.sp = 0; T.hit_any = false; T.result = false;
    float tmin = -1;
    if (A.wide_root == BHRT_WIDE_EMPTY || !intersect_box(A.root_box, ray, any_hit ? BHRT_BIGFLOAT : z, tmin)) { T.done = true; return; }
    T.wide = (const WideNode *)(s.accel_data + A.wide_off);
    T.tris = (const TriRec *)(s.accel_data + A.tris_off);
    T.len_dir = length(ray.dir);
    T.cur = A.wide_root; T.aux = A.wide_root_aux;
    T.done = false;
}

// The whole mesh part of ONE query per octet, run to completion (test kernel; the render kernel interleaves
// the same steps with its queue, wavefront.cuh).  `valid` is octet-uniform.  Returns the MeshJob-style answer.
__device__ __forceinline__ void oct_trace_mesh(OctTrav &T, OctShared *S, bool valid, unsigned sub, unsigned oshift)
{
    for (;;) {
        const bool in_job = valid && !T.done;
        if (!__any_sync(WIDE_FULL, in_job)) break;
        const bool at_leaf = in_job && BHRT_BVH_LEAF(T.cur);
        if (__any_sync(WIDE_FULL, in_job && !at_leaf)) oct_internal(T, S, in_job && !at_leaf, sub, oshift);
        else oct_leaf(T, S, at_leaf, sub, oshift);
    }
}

}  // namespace BHRT_NS
