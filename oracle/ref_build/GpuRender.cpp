// GpuRender.cpp -- the reference-side binding of INTEGRATION.md section 1, made real (TEST INFRASTRUCTURE, built by
// oracle/build_ref.py into oracle/_ref/bhrt_ref_gpu together with the reference's OWN xmlload.cpp and scene classes):
// the reference's LoadScene() fills its globals, this file walks them into the POD tables of include/bhrt.h -- the
// translation unit a maintainer of BHRayTracer would add next to Main.cpp -- and either
//   --dump-desc FILE   writes the canonical serialisation of that description (desc_dump.hpp), which the tests compare
//                      byte for byte with the one libbhrt_host.so produces from the same scene file, or
//   (default)          hands it to libbhrt.so (dlopen, the C ABI): bhrt_create / bhrt_upload_scene / bhrt_render, and
//                      writes the frame as the same .aov the reference driver writes -- BeginRender() on the GPU with
//                      the scene loaded by the reference itself.
// The reference keeps the fields it never meant to export private; a maintainer would add getters, this file
// re-declares `private` instead (after the standard headers).
#include <dlfcn.h>
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#include <iostream>
#include <sstream>
#include <fstream>
#include <atomic>
#include <math.h>
#define private public
#define protected public
#include "Scenes/scene.h"
#include "Objects/objects.h"
#include "Materials/materials.h"
#include "Lights/lights.h"
#include "Textures/texture.h"
#undef private
#undef protected
#include "bhrt.h"
#include "desc_dump.hpp"

extern Node rootNode; extern Camera camera; extern RenderImage renderImage;
extern MaterialList materials; extern LightList lights; extern TexturedColor background, environment;
int LoadScene(char const *filename);

namespace {
struct Flat {
    std::vector<bhrt_node> nodes; std::vector<bhrt_material> mtls; std::vector<bhrt_light> lts;
    std::vector<bhrt_texmap> texmaps; std::vector<bhrt_texture> textures; std::vector<bhrt_mesh> meshes;
    std::vector<uint8_t> texels, meshData;
    std::map<const Texture *, int> texIndex; std::map<const Material *, int> mtlIndex; std::map<const TriObj *, int> meshIndex;
    bhrt_scene_desc D;
    std::string err;
};

void fillXform(const Transformation &t, bhrt_xform &x)
{
    memcpy(x.tm, &t.GetTransform().cell[0], 36);
    memcpy(x.itm, &t.GetInverseTransform().cell[0], 36);
    memcpy(x.pos, &t.GetPosition().x, 12);
}
int addTexture(Flat &F, const Texture *t)
{
    if (!t) return -1;
    auto it = F.texIndex.find(t);
    if (it != F.texIndex.end()) return it->second;
    bhrt_texture d; memset(&d, 0, sizeof(d));
    if (const TextureChecker *c = dynamic_cast<const TextureChecker *>(t)) {
        d.kind = BHRT_TEX_CHECKER;
        d.color1[0] = c->color1.r; d.color1[1] = c->color1.g; d.color1[2] = c->color1.b;
        d.color2[0] = c->color2.r; d.color2[1] = c->color2.g; d.color2[2] = c->color2.b;
    } else {
        const TextureFile *f = static_cast<const TextureFile *>(t);
        d.kind = BHRT_TEX_FILE; d.width = f->width; d.height = f->height;
        d.color2[0] = d.color2[1] = d.color2[2] = 1.f;            // (the checker defaults our loader's Texture carries too)
        d.texel_offset = (int64_t)F.texels.size();
        const uint8_t *p = (const uint8_t *)f->data.data();       // Color24 = 3 bytes r,g,b (cyColor.h:222-240)
        F.texels.insert(F.texels.end(), p, p + f->data.size() * 3);
    }
    int idx = (int)F.textures.size();
    F.textures.push_back(d); F.texIndex[t] = idx;
    return idx;
}
bhrt_texcolor addTexColor(Flat &F, const TexturedColor &tc)
{
    bhrt_texcolor d;
    const Color c = tc.GetColor();
    d.color[0] = c.r; d.color[1] = c.g; d.color[2] = c.b; d.texmap = -1;
    if (const TextureMap *m = tc.GetTexture()) {
        bhrt_texmap tm; memset(&tm, 0, sizeof(tm));
        fillXform(*m, tm.xf);
        tm.texture = addTexture(F, m->texture);
        d.texmap = (int)F.texmaps.size();
        F.texmaps.push_back(tm);
    }
    return d;
}
int addBlinn(Flat &F, const MtlBlinn &b, int flags)
{
    bhrt_material d; memset(&d, 0, sizeof(d));
    d.diffuse = addTexColor(F, b.diffuse);
    d.specular = addTexColor(F, b.specular);
    const Color r = b.refraction.GetColor();
    d.refraction[0] = r.r; d.refraction[1] = r.g; d.refraction[2] = r.b;
    d.absorption[0] = b.absorption.r; d.absorption[1] = b.absorption.g; d.absorption[2] = b.absorption.b;
    d.glossiness = b.glossiness; d.refraction_glossiness = b.refractionGlossiness; d.ior = b.ior; d.flags = flags;
    F.mtls.push_back(d);
    return (int)F.mtls.size() - 1;
}
int addMaterial(Flat &F, const Material *m)
{
    auto it = F.mtlIndex.find(m);
    if (it != F.mtlIndex.end()) return it->second;
    int idx;
    if (const MtlBlinn *b = dynamic_cast<const MtlBlinn *>(m)) idx = addBlinn(F, *b, 0);
    else {
        // MultiMtl::Shade uses mtls[hInfo.mtlID] and mtlID is never set (materials.h:71, TriObj.cpp:68-189)
        const MultiMtl *mm = static_cast<const MultiMtl *>(m);
        if (mm->mtls.empty() || mm->mtls.size() > 255) return -1;
        idx = addBlinn(F, *static_cast<const MtlBlinn *>(mm->mtls[0]), BHRT_MTL_MULTI | ((uint32_t)mm->mtls.size() << 8));
        for (size_t i = 1; i < mm->mtls.size(); i++) addBlinn(F, *static_cast<const MtlBlinn *>(mm->mtls[i]), 0);     // (BHRT_FLAG_FACE_MATERIALS)
    }
    F.mtlIndex[m] = idx;
    return idx;
}
template <class T> int64_t appendBlob(Flat &F, const T *data, size_t count)
{
    while (F.meshData.size() % 16) F.meshData.push_back(0);
    int64_t off = (int64_t)F.meshData.size();
    const uint8_t *p = (const uint8_t *)data;
    F.meshData.insert(F.meshData.end(), p, p + count * sizeof(T));
    return off;
}
int addMesh(Flat &F, const TriObj *t)
{
    auto it = F.meshIndex.find(t);
    if (it != F.meshIndex.end()) return it->second;
    bhrt_mesh d; memset(&d, 0, sizeof(d));
    d.nv = (int)t->nv; d.nf = (int)t->nf; d.nvn = (int)t->nvn; d.nvt = (int)t->nvt;
    d.v_off = appendBlob(F, t->v, t->nv); d.vn_off = appendBlob(F, t->vn, t->nvn); d.vt_off = appendBlob(F, t->vt, t->nvt);
    d.f_off = appendBlob(F, t->f, t->nf); d.fn_off = appendBlob(F, t->fn, t->nf); d.ft_off = appendBlob(F, t->ft, t->nf);
    // cyBVH keeps no node count: the tree is walked from the root (node 1; children are adjacent, cyBVH.h:87-100)
    const cyBVHTriMesh &bvh = t->bvh;
    unsigned maxNode = 1;
    std::vector<unsigned> st(1, 1u);
    while (!st.empty()) {
        unsigned n = st.back(); st.pop_back();
        if (n > maxNode) maxNode = n;
        if (!bvh.IsLeafNode(n)) { st.push_back(bvh.GetFirstChildNode(n)); st.push_back(bvh.GetSecondChildNode(n)); }
    }
    std::vector<bhrt_bvh_node> nodes(maxNode + 1);
    memset(nodes.data(), 0, nodes.size() * sizeof(bhrt_bvh_node));
    for (unsigned n = 1; n <= maxNode; n++) { memcpy(nodes[n].box, bvh.GetNodeBounds(n), 24); nodes[n].data = bvh.nodes[n].data; }
    d.n_bvh_nodes = (int)maxNode;
    d.bvh_off = appendBlob(F, nodes.data(), nodes.size());
    d.elems_off = appendBlob(F, bvh.elements, t->nf);
    d.n_groups = (int)t->nm;                                    // usemtl groups: cyTriMesh's cumulative face counts behind the elems
    if (t->nm > 0) appendBlob(F, t->mcfc, t->nm);
    int idx = (int)F.meshes.size();
    F.meshes.push_back(d); F.meshIndex[t] = idx;
    return idx;
}
bool walk(Flat &F, const Node *parent, int parentIndex, int depth)
{
    for (int i = 0; i < parent->GetNumChild(); i++) {           // DFS pre-order == recursive() order (Main.cpp:389-413)
        const Node *c = parent->GetChild(i);
        if (depth > BHRT_MAX_DEPTH) { F.err = "scene graph deeper than BHRT_MAX_DEPTH"; return false; }
        bhrt_node d; memset(&d, 0, sizeof(d));
        fillXform(*c, d.xf);
        d.parent = parentIndex; d.depth = depth; d.kind = BHRT_OBJ_NONE; d.material = -1; d.mesh = -1;
        if (const Object *o = c->GetNodeObj()) {
            if (!c->GetMaterial()) { F.err = "object node without a material"; return false; }
            d.material = addMaterial(F, c->GetMaterial());
            if (dynamic_cast<const Sphere *>(o)) d.kind = BHRT_OBJ_SPHERE;
            else if (dynamic_cast<const Plane *>(o)) d.kind = BHRT_OBJ_PLANE;
            else { d.kind = BHRT_OBJ_MESH; d.mesh = addMesh(F, static_cast<const TriObj *>(o)); }
        }
        int idx = (int)F.nodes.size();
        F.nodes.push_back(d);
        if (!walk(F, c, idx, depth + 1)) return false;
    }
    return true;
}
bool buildDesc(Flat &F)
{
    if (!walk(F, &rootNode, -1, 1)) return false;
    for (Light *l : lights) {
        bhrt_light d; memset(&d, 0, sizeof(d));
        if (PointLight *p = dynamic_cast<PointLight *>(l)) {
            d.kind = BHRT_LIGHT_POINT; d.v[0] = p->position.x; d.v[1] = p->position.y; d.v[2] = p->position.z; d.size = p->size;
            d.intensity[0] = p->intensity.r; d.intensity[1] = p->intensity.g; d.intensity[2] = p->intensity.b;
        } else if (DirectLight *dl = dynamic_cast<DirectLight *>(l)) {
            d.kind = BHRT_LIGHT_DIRECT; d.v[0] = dl->direction.x; d.v[1] = dl->direction.y; d.v[2] = dl->direction.z;
            d.intensity[0] = dl->intensity.r; d.intensity[1] = dl->intensity.g; d.intensity[2] = dl->intensity.b;
        } else {
            AmbientLight *a = static_cast<AmbientLight *>(l);
            d.kind = BHRT_LIGHT_AMBIENT;
            d.intensity[0] = a->intensity.r; d.intensity[1] = a->intensity.g; d.intensity[2] = a->intensity.b;
        }
        F.lts.push_back(d);
    }
    bhrt_scene_desc &D = F.D;
    memset(&D, 0, sizeof(D));
    D.background = addTexColor(F, background);
    D.environment = addTexColor(F, environment);
    D.abi_version = BHRT_ABI_VERSION;
    D.n_nodes = (int)F.nodes.size(); D.nodes = F.nodes.data();
    D.n_materials = (int)F.mtls.size(); D.materials = F.mtls.data();
    D.n_lights = (int)F.lts.size(); D.lights = F.lts.data();
    D.n_texmaps = (int)F.texmaps.size(); D.texmaps = F.texmaps.data();
    D.n_textures = (int)F.textures.size(); D.textures = F.textures.data();
    D.n_meshes = (int)F.meshes.size(); D.meshes = F.meshes.data();
    D.texels = F.texels.data(); D.texel_bytes = F.texels.size();
    D.mesh_data = F.meshData.data(); D.mesh_bytes = F.meshData.size();
    memcpy(D.camera.pos, &camera.pos.x, 12); memcpy(D.camera.dir, &camera.dir.x, 12); memcpy(D.camera.up, &camera.up.x, 12);
    D.camera.fov = camera.fov; D.camera.focaldist = camera.focaldist; D.camera.dof = camera.dof;
    D.camera.width = camera.imgWidth; D.camera.height = camera.imgHeight;
    return true;
}
}  // namespace

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.xml out_prefix [--dump-desc FILE] [--det] [--spp N] [--seed S] [--lib PATH]\n", argv[0]); return 2; }
    const char *dump = nullptr, *lib = nullptr;
    int det = 0, spp = 16; unsigned long long seed = 1;
    for (int a = 3; a < argc; a++) {
        if (!strcmp(argv[a], "--dump-desc") && a + 1 < argc) dump = argv[++a];
        else if (!strcmp(argv[a], "--det")) det = 1;
        else if (!strcmp(argv[a], "--spp") && a + 1 < argc) spp = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--seed") && a + 1 < argc) seed = strtoull(argv[++a], nullptr, 0);
        else if (!strcmp(argv[a], "--lib") && a + 1 < argc) lib = argv[++a];
        else { fprintf(stderr, "unknown arg %s\n", argv[a]); return 2; }
    }
    // the reference's loader printf()s the whole scene: keep stdout for our one JSON line
    FILE *realOut = fdopen(dup(fileno(stdout)), "w");
    if (!freopen("/dev/null", "w", stdout)) return 3;
    if (!LoadScene(argv[1])) { fprintf(stderr, "the reference's LoadScene failed on %s\n", argv[1]); return 3; }
    Flat F;
    if (!buildDesc(F)) { fprintf(stderr, "flatten: %s\n", F.err.c_str()); return 3; }
    if (dump) {
        if (!desc_dump(F.D, dump)) return 4;
        fprintf(realOut, "{\"dumped\": \"%s\", \"nodes\": %d, \"materials\": %d, \"meshes\": %d, \"textures\": %d}\n", dump, F.D.n_nodes, F.D.n_materials, F.D.n_meshes, F.D.n_textures);
        fflush(realOut);
        return 0;
    }
    // ---- BeginRender() on the GPU through the C ABI
    std::string path = lib ? lib : "libbhrt.so";
    void *h = dlopen(path.c_str(), RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen(%s): %s\n", path.c_str(), dlerror()); return 5; }
    auto p_create = (int (*)(int, bhrt_ctx **))dlsym(h, "bhrt_create");
    auto p_upload = (int (*)(bhrt_ctx *, const bhrt_scene_desc *))dlsym(h, "bhrt_upload_scene");
    auto p_render = (int (*)(bhrt_ctx *, const bhrt_params *, bhrt_outputs *))dlsym(h, "bhrt_render");
    auto p_destroy = (int (*)(bhrt_ctx *))dlsym(h, "bhrt_destroy");
    auto p_err = (const char *(*)())dlsym(h, "bhrt_last_error");
    if (!p_create || !p_upload || !p_render || !p_destroy || !p_err) { fprintf(stderr, "libbhrt.so lacks an entry point\n"); return 5; }
    bhrt_ctx *ctx = nullptr;
    if (p_create(0, &ctx)) { fprintf(stderr, "bhrt: %s\n", p_err()); return 6; }
    if (p_upload(ctx, &F.D)) { fprintf(stderr, "bhrt: %s\n", p_err()); return 6; }
    const int W = camera.imgWidth, H = camera.imgHeight;
    std::vector<float> linear((size_t)W * H * 3), z((size_t)W * H);
    std::vector<uint8_t> rgb8((size_t)W * H * 3);
    std::vector<int32_t> hit((size_t)W * H);
    bhrt_params p; memset(&p, 0, sizeof(p));
    p.mode = det ? BHRT_MODE_DET : BHRT_MODE_PATH; p.spp = det ? 1 : spp; p.gi_depth = 3; p.tir_budget = 16; p.seed = seed;
    bhrt_outputs o; memset(&o, 0, sizeof(o));
    o.rgb8 = rgb8.data(); o.linear = linear.data(); o.hit_id = hit.data(); o.z = z.data();
    if (p_render(ctx, &p, &o)) { fprintf(stderr, "bhrt: %s\n", p_err()); return 6; }
    // what BeginRender() leaves in the reference's RenderImage (Main.cpp:229-230; viewport.cpp:395-415 reads it)
    memcpy(renderImage.GetPixels(), rgb8.data(), rgb8.size());
    memcpy(renderImage.GetZBuffer(), z.data(), z.size() * 4);
    renderImage.IncrementNumRenderPixel(W * H);
    std::string aov = std::string(argv[2]) + ".aov";
    FILE *fp = fopen(aov.c_str(), "wb");
    if (!fp) return 4;
    int32_t hdr[4] = { 0x54524842, W, H, p.spp };
    fwrite(hdr, sizeof(hdr), 1, fp);
    fwrite(linear.data(), 4, linear.size(), fp); fwrite(rgb8.data(), 1, rgb8.size(), fp);
    fwrite(hit.data(), 4, hit.size(), fp); fwrite(z.data(), 4, z.size(), fp);
    fclose(fp);
    fprintf(realOut, "{\"gpu\": true, \"loader\": \"reference xmlload.cpp\", \"width\": %d, \"height\": %d, \"spp\": %d, \"det\": %d, \"kernel_ms\": %.3f, \"rays_closest\": %llu, \"rays_shadow\": %llu, \"render_done\": %d}\n",
            W, H, p.spp, det, o.kernel_ms, (unsigned long long)o.rays_closest, (unsigned long long)o.rays_shadow, renderImage.IsRenderDone() ? 1 : 0);
    fflush(realOut);
    p_destroy(ctx);
    return 0;
}
