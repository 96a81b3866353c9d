// flatten.cpp -- the flattening pass: host scene graph -> SoA device tables (bhrt.h), and
// the C ABI of libbhrt_host.so (bhrt_host.h).
//
// Walks the Node tree in the order recursive() visits it (reference Main.cpp:389-413: for
// each child, test the child's object, then recurse into the child), so the node index is
// also the reference's DFS pre-order id used for hit-id parity.
#include <cstring>
#include <map>
#include "../../../include/bhrt_host.h"
#include "png.hpp"
#include "scene.hpp"

namespace bhrt { namespace host {

namespace {

struct Flattener {
    Scene &sc;
    FlatScene &out;
    std::string *err;
    std::map<const Material *, int> mtlIndex;
    std::map<const TriMeshObject *, int> meshIndex;
    std::map<const Texture *, int> texIndex;

    Flattener(Scene &s, FlatScene &o, std::string *e) : sc(s), out(o), err(e) {}

    bool Fail(const std::string &m) { if (err) *err = m; return false; }

    int AddTexture(const Texture *t)
    {
        if (!t) return -1;
        auto it = texIndex.find(t);
        if (it != texIndex.end()) return it->second;
        bhrt_texture d;
        memset(&d, 0, sizeof(d));
        d.kind = t->kind == Texture::Checker ? BHRT_TEX_CHECKER : BHRT_TEX_FILE;
        d.width = t->width; d.height = t->height;
        d.color1[0] = t->color1.r; d.color1[1] = t->color1.g; d.color1[2] = t->color1.b;
        d.color2[0] = t->color2.r; d.color2[1] = t->color2.g; d.color2[2] = t->color2.b;
        d.texel_offset = 0;
        if (t->kind == Texture::File) {
            while (out.texels.size() % 16) out.texels.push_back(0);
            d.texel_offset = (int64_t)out.texels.size();
            out.texels.insert(out.texels.end(), t->rgb.begin(), t->rgb.end());
        }
        int idx = (int)out.textures.size();
        out.textures.push_back(d);
        texIndex[t] = idx;
        return idx;
    }
    bhrt_texcolor AddTexColor(const TexturedColor &tc)
    {
        bhrt_texcolor d;
        d.color[0] = tc.color.r; d.color[1] = tc.color.g; d.color[2] = tc.color.b;
        d.texmap = -1;
        if (tc.map) {
            bhrt_texmap m;
            memset(&m, 0, sizeof(m));
            tc.map->FillXform(m.xf);
            m.texture = AddTexture(tc.map->texture);
            d.texmap = (int)out.texmaps.size();
            out.texmaps.push_back(m);
        }
        return d;
    }
    int AddBlinn(const MtlBlinn &b, int flags)
    {
        bhrt_material d;
        memset(&d, 0, sizeof(d));
        d.diffuse = AddTexColor(b.diffuse);
        d.specular = AddTexColor(b.specular);
        d.refraction[0] = b.refraction.color.r; d.refraction[1] = b.refraction.color.g; d.refraction[2] = b.refraction.color.b;
        d.absorption[0] = b.absorption.r; d.absorption[1] = b.absorption.g; d.absorption[2] = b.absorption.b;
        d.glossiness = b.glossiness;
        d.refraction_glossiness = b.refractionGlossiness;
        d.ior = b.ior;
        d.flags = flags;
        out.materials.push_back(d);
        return (int)out.materials.size() - 1;
    }
    int AddMaterial(const Material *m, bool &ok)
    {
        auto it = mtlIndex.find(m);
        if (it != mtlIndex.end()) return it->second;
        int idx;
        if (m->kind == Material::Blinn) idx = AddBlinn(*static_cast<const MtlBlinn *>(m), 0);
        else {
            // MultiMtl::Shade uses mtls[hInfo.mtlID] and mtlID is never set (materials.h:71,
            // TriObj.cpp:68-189), so sub-material 0 shades every face.
            const MultiMtl *mm = static_cast<const MultiMtl *>(m);
            if (mm->mtls.empty() || mm->mtls.size() > 255) { ok = false; return -1; }
            idx = AddBlinn(*mm->mtls[0], BHRT_MTL_MULTI | ((uint32_t)mm->mtls.size() << 8));
            // the other sub-materials follow contiguously (BHRT_FLAG_FACE_MATERIALS; bhrt.h)
            for (size_t i = 1; i < mm->mtls.size(); i++) AddBlinn(*mm->mtls[i], 0);
        }
        mtlIndex[m] = idx;
        return idx;
    }
    template <class T> int64_t AppendBlob(const T *data, size_t count)
    {
        while (out.meshData.size() % 16) out.meshData.push_back(0);
        int64_t off = (int64_t)out.meshData.size();
        const uint8_t *p = (const uint8_t *)data;
        out.meshData.insert(out.meshData.end(), p, p + count * sizeof(T));
        return off;
    }
    int AddMesh(const TriMeshObject *t)
    {
        auto it = meshIndex.find(t);
        if (it != meshIndex.end()) return it->second;
        bhrt_mesh d;
        memset(&d, 0, sizeof(d));
        d.nv = (int)t->v.size(); d.nf = (int)t->f.size(); d.nvn = (int)t->vn.size(); d.nvt = (int)t->vt.size();
        d.n_bvh_nodes = (int)t->bvhNodes.size() - 1;
        d.v_off = AppendBlob(t->v.data(), t->v.size());
        d.vn_off = AppendBlob(t->vn.data(), t->vn.size());
        d.vt_off = AppendBlob(t->vt.data(), t->vt.size());
        d.f_off = AppendBlob(t->f.data(), t->f.size());
        d.fn_off = AppendBlob(t->fn.data(), t->fn.size());
        d.ft_off = AppendBlob(t->ft.data(), t->ft.size());
        std::vector<bhrt_bvh_node> nodes(t->bvhNodes.size());
        for (size_t i = 0; i < nodes.size(); i++) {
            memcpy(nodes[i].box, t->bvhNodes[i].box, sizeof(float) * 6);
            nodes[i].data = t->bvhNodes[i].data;
            nodes[i].pad = 0;
        }
        d.bvh_off = AppendBlob(nodes.data(), nodes.size());
        d.elems_off = AppendBlob(t->bvhElements.data(), t->bvhElements.size());
        // usemtl groups: cumulative face counts (cyTriMesh mcfc), right behind the elems array (bhrt.h: bhrt_mesh::n_groups)
        d.n_groups = (int)t->mcfc.size();
        if (d.n_groups > 0) AppendBlob(t->mcfc.data(), t->mcfc.size());
        int idx = (int)out.meshes.size();
        out.meshes.push_back(d);
        meshIndex[t] = idx;
        return idx;
    }
    bool Walk(const Node *parent, int parentIndex, int depth)
    {
        for (auto &cp : parent->children) {
            const Node *c = cp.get();
            if (depth > BHRT_MAX_DEPTH) return Fail("scene graph deeper than BHRT_MAX_DEPTH");
            bhrt_node d;
            memset(&d, 0, sizeof(d));
            c->FillXform(d.xf);
            d.parent = parentIndex;
            d.depth = depth;
            d.kind = BHRT_OBJ_NONE;
            d.material = -1;
            d.mesh = -1;
            if (c->obj) {
                if (!c->mtl)
                    return Fail("object node \"" + c->name + "\" has no material: the reference dereferences a null Material here (Main.cpp:161)");
                bool ok = true;
                d.material = AddMaterial(c->mtl, ok);
                if (!ok) return Fail("MultiMtl without sub-materials on node \"" + c->name + "\"");
                switch (c->obj->kind) {
                case Object::Sphere: d.kind = BHRT_OBJ_SPHERE; break;
                case Object::Plane: d.kind = BHRT_OBJ_PLANE; break;
                case Object::TriObj: d.kind = BHRT_OBJ_MESH; d.mesh = AddMesh(static_cast<const TriMeshObject *>(c->obj)); break;
                }
            }
            int idx = (int)out.nodes.size();
            out.nodes.push_back(d);
            out.nodeNames.push_back(c->name);
            if (!Walk(c, idx, depth + 1)) return false;
        }
        return true;
    }
};

}  // namespace

bool Flatten(Scene &sc, FlatScene &out, std::string *err)
{
    Flattener fl(sc, out, err);
    if (!fl.Walk(&sc.rootNode, -1, 1)) return false;
    for (auto &l : sc.lights) {
        bhrt_light d;
        memset(&d, 0, sizeof(d));
        d.intensity[0] = l->intensity.r; d.intensity[1] = l->intensity.g; d.intensity[2] = l->intensity.b;
        switch (l->kind) {
        case Light::Ambient: d.kind = BHRT_LIGHT_AMBIENT; break;
        case Light::Direct: d.kind = BHRT_LIGHT_DIRECT; d.v[0] = l->direction.x; d.v[1] = l->direction.y; d.v[2] = l->direction.z; break;
        case Light::Point: d.kind = BHRT_LIGHT_POINT; d.v[0] = l->position.x; d.v[1] = l->position.y; d.v[2] = l->position.z; d.size = l->size; break;
        }
        out.lights.push_back(d);
    }
    bhrt_scene_desc &D = out.desc;
    memset(&D, 0, sizeof(D));
    D.background = fl.AddTexColor(sc.background);
    D.environment = fl.AddTexColor(sc.environment);
    if (out.nodes.size() > BHRT_MAX_NODES) return fl.Fail("too many nodes");
    if (out.materials.size() > BHRT_MAX_MATERIALS) return fl.Fail("too many materials");
    if (out.lights.size() > BHRT_MAX_LIGHTS) return fl.Fail("too many lights");
    if (out.texmaps.size() > BHRT_MAX_TEXMAPS) return fl.Fail("too many texture maps");
    if (out.textures.size() > BHRT_MAX_TEXTURES) return fl.Fail("too many textures");
    if (out.meshes.size() > BHRT_MAX_MESHES) return fl.Fail("too many meshes");
    D.abi_version = BHRT_ABI_VERSION;
    D.n_nodes = (int)out.nodes.size(); D.nodes = out.nodes.data();
    D.n_materials = (int)out.materials.size(); D.materials = out.materials.data();
    D.n_lights = (int)out.lights.size(); D.lights = out.lights.data();
    D.n_texmaps = (int)out.texmaps.size(); D.texmaps = out.texmaps.data();
    D.n_textures = (int)out.textures.size(); D.textures = out.textures.data();
    D.n_meshes = (int)out.meshes.size(); D.meshes = out.meshes.data();
    D.texels = out.texels.data(); D.texel_bytes = out.texels.size();
    D.mesh_data = out.meshData.data(); D.mesh_bytes = out.meshData.size();
    const Camera &c = sc.camera;
    D.camera.pos[0] = c.pos.x; D.camera.pos[1] = c.pos.y; D.camera.pos[2] = c.pos.z;
    D.camera.dir[0] = c.dir.x; D.camera.dir[1] = c.dir.y; D.camera.dir[2] = c.dir.z;
    D.camera.up[0] = c.up.x; D.camera.up[1] = c.up.y; D.camera.up[2] = c.up.z;
    D.camera.fov = c.fov; D.camera.focaldist = c.focaldist; D.camera.dof = c.dof;
    D.camera.width = c.imgWidth; D.camera.height = c.imgHeight;
    return true;
}

}}  // namespace

// ---------------------------------------------------------------------------------------------
struct bhrt_host_scene {
    bhrt::host::Scene scene;
    bhrt::host::FlatScene flat;
    std::string warnings;
};

static thread_local std::string g_host_error;

extern "C" {

bhrt_host_scene *bhrt_host_load_scene(const char *xml_path)
{
    g_host_error.clear();
    if (!xml_path) { g_host_error = "null path"; return nullptr; }
    bhrt_host_scene *s = new bhrt_host_scene;
    std::string err;
    if (!bhrt::host::LoadScene(xml_path, s->scene, &err)) { g_host_error = err; delete s; return nullptr; }
    if (!bhrt::host::Flatten(s->scene, s->flat, &err)) { g_host_error = err; delete s; return nullptr; }
    for (auto &w : s->scene.warnings) { s->warnings += w; s->warnings += "\n"; }
    return s;
}
void bhrt_host_free_scene(bhrt_host_scene *scene) { delete scene; }
const char *bhrt_host_last_error(void) { return g_host_error.c_str(); }
const bhrt_scene_desc *bhrt_host_scene_desc(const bhrt_host_scene *scene) { return scene ? &scene->flat.desc : nullptr; }
const char *bhrt_host_node_name(const bhrt_host_scene *scene, int32_t node)
{
    if (!scene || node < 0 || node >= (int32_t)scene->flat.nodeNames.size()) return "";
    return scene->flat.nodeNames[node].c_str();
}
const char *bhrt_host_warnings(const bhrt_host_scene *scene) { return scene ? scene->warnings.c_str() : ""; }
int bhrt_host_set_resolution(bhrt_host_scene *scene, int32_t width, int32_t height)
{
    if (!scene || width <= 0 || height <= 0) return BHRT_ERR_ARG;
    scene->scene.camera.imgWidth = width; scene->scene.camera.imgHeight = height;
    scene->flat.desc.camera.width = width; scene->flat.desc.camera.height = height;
    return BHRT_OK;
}
// RenderImage::ComputeZBufferImage (reference Scenes/scene.h:579-600)
void bhrt_host_zbuffer_image(const float *z, int32_t n, uint8_t *out)
{
    const float BIG = 1.0e30f;
    float zmin = BIG, zmax = 0;
    for (int i = 0; i < n; i++) {
        if (z[i] == BIG) continue;
        if (zmin > z[i]) zmin = z[i];
        if (zmax < z[i]) zmax = z[i];
    }
    for (int i = 0; i < n; i++) {
        if (z[i] == BIG) { out[i] = 0; continue; }
        float f = (zmax - z[i]) / (zmax - zmin);
        int c = int(f * 255);
        out[i] = (uint8_t)(c < 0 ? 0 : (c > 255 ? 255 : c));
    }
}

// RenderImage::ComputeSampleCountImage (reference Scenes/scene.h:602-626)
int bhrt_host_sample_count_image(const uint8_t *count, int32_t n, uint8_t *out)
{
    uint8_t smin = 255, smax = 0;
    for (int i = 0; i < n; i++) {
        if (smin > count[i]) smin = count[i];
        if (smax < count[i]) smax = count[i];
    }
    for (int i = 0; i < n; i++) {
        if (smax == smin) { out[i] = 0; continue; }
        int c = (255 * (count[i] - smin)) / (smax - smin);
        out[i] = (uint8_t)(c < 0 ? 0 : (c > 255 ? 255 : c));
    }
    return smax;
}

int bhrt_host_save_png(const char *path, const uint8_t *data, int32_t width, int32_t height, int32_t comp)
{
    if (!path || !data || width <= 0 || height <= 0) return BHRT_ERR_ARG;
    return bhrt::host::png_save(path, data, width, height, comp) ? BHRT_OK : BHRT_ERR_ARG;
}

}  // extern "C"
