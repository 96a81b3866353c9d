// accel.hpp -- host-side builder of the device traversal tables.
//
// The C ABI hands over each mesh in the reference's own layout (indexed v/vn/vt + the cyBVH
// node array, bhrt.h).  The kernels do not chase those indices: at upload the library repacks
// every mesh into
//   * PairNode  (64 B, four 128-bit loads): BOTH child boxes of an internal cyBVH node plus the
//     two child words -- one fetch yields the two slab tests TraceBVHNode / TraceBVHShadow do
//     per step (reference Objects/TriObj/TriObj.cpp:192-307);
//   * TriRec    (48 B, three 128-bit loads), stored in LEAF ORDER so a leaf's triangles are
//     contiguous: the geometric normal vN = (v1-v0)x(v2-v0), |vN|, vN.v0 and the three vertices
//     already projected on the triangle's dominant axis pair -- everything
//     TriObj::IntersectTriangle (TriObj.cpp:68-160) derives from the vertices alone.
// The topology is untouched (same nodes, same child order, same leaves): the reference's mesh
// shadow test depends on it (first-leaf rule, TriObj.cpp:272-307).
//
// Bit-exactness: the precomputed values are produced by the SAME float expressions, in the same
// order, that the per-ray code evaluated before (core.cuh V3 helpers, no FMA contraction on the
// host: g++ x86-64 baseline has no FMA and tests/emu uses -ffp-contract=off; sqrtf is correctly
// rounded on both sides), so precomputing them cannot change a hit/miss decision.
//
// Shared by render.cu (the library) and tests/emu (the CPU harness of the same device code).
#pragma once
#include <cstring>
#include <string>
#include <vector>
#include "core.cuh"

namespace bhrt {

struct AccelBlob {
    std::vector<MeshAccel> meshes;      // one per bhrt_mesh, offsets into `data`
    std::vector<uint8_t> data;          // PairNode / TriRec arrays, 64-byte aligned
};

inline size_t accel_align(std::vector<uint8_t> &d)
{
    size_t o = (d.size() + 63) & ~(size_t)63;
    d.resize(o);
    return o;
}

inline bool build_accel(const bhrt_scene_desc &D, AccelBlob &out, std::string *err)
{
    out.meshes.clear();
    out.data.clear();
    for (int mi = 0; mi < D.n_meshes; mi++) {
        const bhrt_mesh &m = D.meshes[mi];
        const float *v = (const float *)(D.mesh_data + m.v_off);
        const uint32_t *f = (const uint32_t *)(D.mesh_data + m.f_off);
        const bhrt_bvh_node *bvh = (const bhrt_bvh_node *)(D.mesh_data + m.bvh_off);
        const uint32_t *elems = (const uint32_t *)(D.mesh_data + m.elems_off);
        MeshAccel A;
        memset(&A, 0, sizeof(A));
        A.vn_off = m.vn_off; A.vt_off = m.vt_off; A.fn_off = m.fn_off; A.ft_off = m.ft_off;

        // ---- pair nodes: DFS pre-order numbering of the internal nodes (root first)
        std::vector<int> pair_of(m.n_bvh_nodes + 2, -1);
        std::vector<uint32_t> order;           // internal node ids in pair order
        {
            std::vector<uint32_t> st, depth;
            st.push_back(1); depth.push_back(0);
            while (!st.empty()) {
                uint32_t n = st.back(); st.pop_back();
                const uint32_t dn = depth.back(); depth.pop_back();
                // the traversal defers one sibling per level on a fixed stack: a deeper tree cannot be
                // traversed correctly, so it is refused here instead of dropping nodes silently
                if (dn > BHRT_BVH_STACK) { if (err) *err = "mesh BVH deeper than BHRT_BVH_STACK levels"; return false; }
                if (dn > A.max_depth) A.max_depth = dn;
                if (n == 0 || n > (uint32_t)m.n_bvh_nodes) { if (err) *err = "mesh BVH child index out of range"; return false; }
                uint32_t d = bvh[n].data;
                if (BHRT_BVH_LEAF(d)) continue;
                if (pair_of[n] >= 0) { if (err) *err = "mesh BVH is not a tree"; return false; }
                pair_of[n] = (int)order.size();
                order.push_back(n);
                uint32_t c1 = BHRT_BVH_CHILD(d);
                st.push_back(c1 + 1); depth.push_back(dn + 1);
                st.push_back(c1); depth.push_back(dn + 1);          // child 1 is numbered right after its parent
            }
        }
        auto word = [&](uint32_t n) -> uint32_t {
            uint32_t d = bvh[n].data;
            if (BHRT_BVH_LEAF(d)) return d;                 // leaf bit | (count-1)<<28 | first record (= elems offset)
            return (uint32_t)pair_of[n];
        };
        for (int k = 0; k < 6; k++) A.root_box[k] = bvh[1].box[k];
        A.root_word = word(1);
        A.pairs_off = (int64_t)accel_align(out.data);
        out.data.resize(out.data.size() + sizeof(PairNode) * (order.size() ? order.size() : 1));
        PairNode *pairs = (PairNode *)(out.data.data() + A.pairs_off);
        for (size_t i = 0; i < order.size(); i++) {
            uint32_t c1 = BHRT_BVH_CHILD(bvh[order[i]].data), c2 = c1 + 1;
            PairNode &p = pairs[i];
            for (int k = 0; k < 6; k++) { p.box1[k] = bvh[c1].box[k]; p.box2[k] = bvh[c2].box[k]; }
            p.word1 = word(c1); p.word2 = word(c2); p.pad0 = 0; p.pad1 = 0;
        }

        // ---- triangle records in leaf (elems) order
        A.tris_off = (int64_t)accel_align(out.data);
        out.data.resize(out.data.size() + sizeof(TriRec) * (size_t)m.nf);
        TriRec *tris = (TriRec *)(out.data.data() + A.tris_off);
        // (re-take the pointer: the vector may have moved)
        pairs = (PairNode *)(out.data.data() + A.pairs_off);
        (void)pairs;
        for (int i = 0; i < m.nf; i++) {
            uint32_t face = elems[i];
            if (face >= (uint32_t)m.nf) { if (err) *err = "mesh BVH element out of range"; return false; }
            uint32_t i0 = f[3 * face], i1 = f[3 * face + 1], i2 = f[3 * face + 2];
            V3 v0 = ld3(v + 3 * i0), v1 = ld3(v + 3 * i1), v2 = ld3(v + 3 * i2);
            V3 vN = cross(v1 - v0, v2 - v0);
            float ax = fabsf(vN.x), ay = fabsf(vN.y), az = fabsf(vN.z);
            uint32_t axis;
            float p[6];
            if (ax >= ay && ax >= az) { axis = 0; p[0] = v0.y; p[1] = v0.z; p[2] = v1.y; p[3] = v1.z; p[4] = v2.y; p[5] = v2.z; }
            else if (ay >= ax && ay >= az) { axis = 1; p[0] = v0.x; p[1] = v0.z; p[2] = v1.x; p[3] = v1.z; p[4] = v2.x; p[5] = v2.z; }
            else { axis = 2; p[0] = v0.x; p[1] = v0.y; p[2] = v1.x; p[3] = v1.y; p[4] = v2.x; p[5] = v2.y; }
            TriRec &t = tris[i];
            t.a.x = vN.x; t.a.y = vN.y; t.a.z = vN.z; t.a.w = length(vN);
            t.b.x = dot(vN, v0); t.b.y = p[0]; t.b.z = p[1]; t.b.w = p[2];
            t.c.x = p[3]; t.c.y = p[4]; t.c.z = p[5];
            uint32_t tag = face | (axis << 30);
            memcpy(&t.c.w, &tag, 4);
        }
        out.meshes.push_back(A);
    }
    accel_align(out.data);
    return true;
}

}  // namespace bhrt
