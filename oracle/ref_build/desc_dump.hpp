// desc_dump.hpp -- TEST INFRASTRUCTURE: canonical serialisation of a bhrt_scene_desc (include/bhrt.h), independent of
// where the tables sit in their blobs, so two descriptions of the same scene -- one flattened from the REFERENCE's own
// loaded scene (GpuRender.cpp), one from this repository's loader (libbhrt_host.so) -- can be compared byte for byte.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "bhrt.h"

static inline void dd_put(FILE *f, const char *tag, const void *p, size_t bytes)
{
    char t[8] = { 0 };
    strncpy(t, tag, 7);
    uint64_t n = bytes;
    fwrite(t, 1, 8, f); fwrite(&n, 8, 1, f);
    if (bytes) fwrite(p, 1, bytes, f);
}

static inline bool desc_dump(const bhrt_scene_desc &D, const char *path)
{
    FILE *f = fopen(path, "wb");
    if (!f) return false;
    dd_put(f, "nodes", D.nodes, sizeof(bhrt_node) * D.n_nodes);
    dd_put(f, "mtls", D.materials, sizeof(bhrt_material) * D.n_materials);
    dd_put(f, "lights", D.lights, sizeof(bhrt_light) * D.n_lights);
    dd_put(f, "texmaps", D.texmaps, sizeof(bhrt_texmap) * D.n_texmaps);
    for (int i = 0; i < D.n_textures; i++) {
        bhrt_texture t = D.textures[i];
        const int64_t off = t.texel_offset;
        t.texel_offset = 0;
        dd_put(f, "texture", &t, sizeof(t));
        dd_put(f, "texels", t.kind == BHRT_TEX_FILE ? D.texels + off : nullptr, t.kind == BHRT_TEX_FILE ? (size_t)t.width * t.height * 3 : 0);
    }
    for (int i = 0; i < D.n_meshes; i++) {
        const bhrt_mesh &m = D.meshes[i];
        int32_t counts[5] = { m.nv, m.nf, m.nvn, m.nvt, m.n_bvh_nodes };
        dd_put(f, "mesh", counts, sizeof(counts));
        dd_put(f, "v", D.mesh_data + m.v_off, (size_t)m.nv * 12);
        dd_put(f, "vn", D.mesh_data + m.vn_off, (size_t)m.nvn * 12);
        dd_put(f, "vt", D.mesh_data + m.vt_off, (size_t)m.nvt * 12);
        dd_put(f, "f", D.mesh_data + m.f_off, (size_t)m.nf * 12);
        dd_put(f, "fn", D.mesh_data + m.fn_off, (size_t)m.nf * 12);
        dd_put(f, "ft", D.mesh_data + m.ft_off, (size_t)m.nf * 12);
        // (node 0 of the cyBVH array is never initialised by the reference: skip it)
        dd_put(f, "bvh", D.mesh_data + m.bvh_off + sizeof(bhrt_bvh_node), sizeof(bhrt_bvh_node) * (size_t)m.n_bvh_nodes);
        dd_put(f, "elems", D.mesh_data + m.elems_off, (size_t)m.nf * 4);
        // usemtl groups: cumulative face counts behind the elems array (bhrt.h: bhrt_mesh::n_groups)
        dd_put(f, "groups", &m.n_groups, sizeof(m.n_groups));
        if (m.n_groups > 0) dd_put(f, "group_end", D.mesh_data + ((m.elems_off + 4 * (int64_t)m.nf + 15) & ~(int64_t)15), (size_t)m.n_groups * 4);
    }
    dd_put(f, "bg", &D.background, sizeof(D.background));
    dd_put(f, "env", &D.environment, sizeof(D.environment));
    dd_put(f, "camera", &D.camera, sizeof(D.camera));
    fclose(f);
    return true;
}
