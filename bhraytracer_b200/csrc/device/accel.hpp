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

// Collapse the reference's binary cyBVH into the 8-wide tree of core.cuh (WideNode):
//  * a subtree that holds at most 8 triangles becomes ONE leaf group -- its records are contiguous because
//    cy::BVH partitions its element array in place (cyBVH.h:122-142), which is checked here, not assumed;
//    `aux` marks where each of the reference's own leaves starts inside the group;
//  * every other internal node starts with its two children and keeps replacing the child with the largest
//    box surface -- in place, so the left-to-right (= the reference's DFS) order survives -- by that child's
//    two children until it has 8 or nothing is left to open.
// The topology of the reference tree is only regrouped, never rebuilt: every box is one of cyBVH's boxes.
// A.wide_stack = the exact worst case of deferred entries of one traversal; 0 = no wide tree (too deep for
// BHRT_WIDE_STACK, or element ranges that are not contiguous): the kernels then keep the binary traversal.
inline bool build_wide(const bhrt_mesh &m, const bhrt_bvh_node *bvh, MeshAccel &A, std::vector<uint8_t> &data, std::string *err)
{
    (void)err;
    const int nn = m.n_bvh_nodes;
    // triangle range of every binary node (post-order over an explicit stack)
    std::vector<uint32_t> first(nn + 2, 0), count(nn + 2, 0);
    std::vector<uint8_t> contiguous(nn + 2, 1);
    {
        std::vector<std::pair<uint32_t, int>> st;
        st.push_back(std::make_pair(1u, 0));
        while (!st.empty()) {
            const uint32_t n = st.back().first; const int phase = st.back().second;
            const uint32_t d = bvh[n].data;
            if (BHRT_BVH_LEAF(d)) { first[n] = BHRT_BVH_OFFSET(d); count[n] = BHRT_BVH_COUNT(d); st.pop_back(); continue; }
            const uint32_t c1 = BHRT_BVH_CHILD(d), c2 = c1 + 1;
            if (phase == 0) { st.back().second = 1; st.push_back(std::make_pair(c2, 0)); st.push_back(std::make_pair(c1, 0)); continue; }
            first[n] = first[c1]; count[n] = count[c1] + count[c2];
            contiguous[n] = contiguous[c1] && contiguous[c2] && first[c1] + count[c1] == first[c2];
            st.pop_back();
        }
    }
    A.wide_stack = 0; A.wide_nodes = 0; A.wide_root = BHRT_WIDE_EMPTY; A.wide_root_aux = 0;
    A.wide_off = (int64_t)accel_align(data);
    auto is_group = [&](uint32_t n) { return BHRT_BVH_LEAF(bvh[n].data) || (count[n] <= 8 && contiguous[n]); };
    auto group_word = [&](uint32_t n, uint32_t &aux) -> uint32_t {
        aux = 0;
        std::vector<uint32_t> st(1, n);
        while (!st.empty()) {
            const uint32_t k = st.back(); st.pop_back();
            const uint32_t d = bvh[k].data;
            if (BHRT_BVH_LEAF(d)) { aux |= 1u << (BHRT_BVH_OFFSET(d) - first[n]); continue; }
            st.push_back(BHRT_BVH_CHILD(d) + 1); st.push_back(BHRT_BVH_CHILD(d));
        }
        return 0x80000000u | ((count[n] - 1u) << 28) | first[n];
    };
    auto area = [&](uint32_t n) {
        const float *b = bvh[n].box;
        const float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
        return dx * dy + dy * dz + dz * dx;
    };
    if (is_group(1)) {                       // the whole mesh is one leaf group
        A.wide_root = group_word(1, A.wide_root_aux);
        A.wide_stack = 1;
        data.resize(data.size() + sizeof(WideNode));     // (never read)
        return true;
    }
    std::vector<WideNode> nodes;
    std::vector<uint32_t> need;              // worst-case deferred entries of the subtree of each wide node
    bool ok = true;
    // recursive descent with an explicit stack of (binary node, wide index) pairs
    struct Item { uint32_t bin, idx; };
    std::vector<Item> work;
    nodes.emplace_back(); need.push_back(0);
    work.push_back(Item{ 1u, 0u });
    std::vector<std::vector<uint32_t>> kids;         // internal wide children of each wide node (for `need`)
    kids.emplace_back();
    std::vector<uint32_t> fan(1, 0);
    while (!work.empty()) {
        const Item it = work.back(); work.pop_back();
        std::vector<uint32_t> ch;
        ch.push_back(BHRT_BVH_CHILD(bvh[it.bin].data)); ch.push_back(BHRT_BVH_CHILD(bvh[it.bin].data) + 1);
        while (ch.size() < 8) {
            int best = -1; float best_a = -1.f;
            for (size_t i = 0; i < ch.size(); i++) if (!is_group(ch[i]) && area(ch[i]) > best_a) { best_a = area(ch[i]); best = (int)i; }
            if (best < 0) break;
            const uint32_t c1 = BHRT_BVH_CHILD(bvh[ch[best]].data);
            ch[best] = c1; ch.insert(ch.begin() + best + 1, c1 + 1);
        }
        WideNode wn;
        memset(&wn, 0, sizeof(wn));
        for (int j = 0; j < 8; j++) wn.c[j].word = BHRT_WIDE_EMPTY;
        fan[it.idx] = (uint32_t)ch.size();
        for (size_t j = 0; j < ch.size(); j++) {
            for (int k = 0; k < 6; k++) wn.c[j].box[k] = bvh[ch[j]].box[k];
            if (is_group(ch[j])) wn.c[j].word = group_word(ch[j], wn.c[j].aux);
            else {
                const uint32_t idx = (uint32_t)nodes.size();
                nodes.emplace_back(); need.push_back(0); kids.emplace_back(); fan.push_back(0);
                wn.c[j].word = idx;
                kids[it.idx].push_back(idx);
                work.push_back(Item{ ch[j], idx });
            }
        }
        nodes[it.idx] = wn;
    }
    // need(node) = (children - 1) + max over internal children; children were created after their parents
    for (size_t i = nodes.size(); i-- > 0;) {
        uint32_t mx = 0;
        for (uint32_t k : kids[i]) if (need[k] > mx) mx = need[k];
        need[i] = fan[i] - 1 + mx;
    }
    if (need[0] + 1 > BHRT_WIDE_STACK || nodes.size() >= 0x7fffffffu) ok = false;
    if (!ok) { data.resize(data.size() + sizeof(WideNode)); return true; }
    A.wide_root = 0; A.wide_root_aux = 0;
    A.wide_stack = need[0] + 1; A.wide_nodes = (uint32_t)nodes.size();
    const size_t o = data.size();
    data.resize(o + nodes.size() * sizeof(WideNode));
    memcpy(data.data() + o, nodes.data(), nodes.size() * sizeof(WideNode));
    return true;
}

inline bool build_accel(const bhrt_scene_desc &D, AccelBlob &out, std::string *err, bool with_wide = true)
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
        A.n_groups = m.n_groups; A.groups_off = (m.elems_off + 4 * (int64_t)m.nf + 15) & ~(int64_t)15;

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
        // ---- the wide tree: the same cyBVH collapsed to <= 8 children per node (fast build of the kernels)
        if (with_wide && !build_wide(m, bvh, A, out.data, err)) return false;
        out.meshes.push_back(A);
    }
    accel_align(out.data);
    return true;
}

}  // namespace bhrt
