// loader.cpp -- scene ingestion with the reference's XML semantics.
//
// Restates LoadScene / LoadNode / LoadTransform / LoadMaterial / LoadLight / ReadTexture /
// ReadVector / ReadColor / ReadFloat of reference BHRayTracer/xmlload.cpp (line numbers cited
// per function) on top of xml.hpp / png.hpp, filling a bhrt::host::Scene instead of globals.
// Schema: SURVEY.md Appendix B.
#include <strings.h>
#include <cctype>
#include <cstdio>
#include <cstring>
#include "png.hpp"
#include "scene.hpp"
#include "xml.hpp"

namespace bhrt { namespace host {

static bool Compare(const char *a, const char *b) { return strcasecmp(a, b) == 0; }   // xmlload.cpp:34-38

// xmlload.cpp:516-521
static void ReadFloat(const XmlElement *e, float &f, const char *name = "value")
{
    double d = (double)f;
    e->QueryDouble(name, &d);
    f = (float)d;
}
// xmlload.cpp:478-493
static void ReadVector(const XmlElement *e, Vec3 &v)
{
    double x = (double)v.x, y = (double)v.y, z = (double)v.z;
    e->QueryDouble("x", &x);
    e->QueryDouble("y", &y);
    e->QueryDouble("z", &z);
    v.x = (float)x; v.y = (float)y; v.z = (float)z;
    float f = 1;
    ReadFloat(e, f);
    v = v * f;
}
// xmlload.cpp:497-512
static void ReadColor(const XmlElement *e, Color &c)
{
    double r = (double)c.r, g = (double)c.g, b = (double)c.b;
    e->QueryDouble("r", &r);
    e->QueryDouble("g", &g);
    e->QueryDouble("b", &b);
    c.r = (float)r; c.g = (float)g; c.b = (float)b;
    float f = 1;
    ReadFloat(e, f);
    c.r *= f; c.g *= f; c.b *= f;
}

// xmlload.cpp:275-303
static void LoadTransform(Transformation *t, const XmlElement *e)
{
    for (auto &ch : e->children) {
        const char *tag = ch->name.c_str();
        if (Compare(tag, "scale")) {
            Vec3 s(1, 1, 1);
            ReadVector(ch.get(), s);
            t->Scale(s.x, s.y, s.z);
        } else if (Compare(tag, "rotate")) {
            Vec3 s(0, 0, 0);
            ReadVector(ch.get(), s);
            s = Normalized(s);
            float a = 0;       // uninitialised in the reference when the attribute is missing
            ReadFloat(ch.get(), a, "angle");
            t->Rotate(s, a);
        } else if (Compare(tag, "translate")) {
            Vec3 v(0, 0, 0);
            ReadVector(ch.get(), v);
            t->Translate(v);
        }
    }
}

// LoadPPM (Textures/Texture.cpp:18-54)
static int PpmReadLine(FILE *fp, int size, char *buffer)
{
    int i;
    for (i = 0; i < size; i++) {
        buffer[i] = (char)fgetc(fp);
        if (feof(fp) || buffer[i] == '\n' || buffer[i] == '\r') { buffer[i] = '\0'; return i + 1; }
    }
    return i;
}
static bool LoadPPM(const char *name, int &w, int &h, std::vector<uint8_t> &rgb)
{
    FILE *fp = fopen(name, "rb");
    if (!fp) return false;
    char buffer[1024];
    PpmReadLine(fp, 1024, buffer);
    if (buffer[0] != 'P' && buffer[1] != '6') { fclose(fp); return false; }
    PpmReadLine(fp, 1024, buffer);
    while (buffer[0] == '#') PpmReadLine(fp, 1024, buffer);
    w = h = 0;
    sscanf(buffer, "%d %d", &w, &h);
    PpmReadLine(fp, 1024, buffer);
    while (buffer[0] == '#') PpmReadLine(fp, 1024, buffer);
    if (w <= 0 || h <= 0) { fclose(fp); return false; }
    rgb.resize((size_t)w * h * 3);
    size_t got = fread(rgb.data(), 3, (size_t)w * h, fp);
    (void)got;
    fclose(fp);
    return true;
}

// TextureFile::Load (Textures/Texture.cpp:58-93)
static bool LoadTextureFile(Texture &t)
{
    const char *name = t.name.c_str();
    int len = (int)t.name.size();
    if (len < 3) return false;
    char ext[4] = { (char)tolower(name[len - 3]), (char)tolower(name[len - 2]), (char)tolower(name[len - 1]), 0 };
    if (!strcmp(ext, "png")) return png_load_rgb8(name, t.rgb, t.width, t.height);
    if (!strcmp(ext, "ppm")) return LoadPPM(name, t.width, t.height, t.rgb);
    return false;
}

// ReadTexture(char const*) (xmlload.cpp:562-582): cached by name
static Texture *ReadTextureFile(Scene &sc, const char *texName)
{
    for (auto &t : sc.textureList) if (t->kind == Texture::File && t->name == texName) return t.get();
    std::unique_ptr<Texture> t(new Texture);
    t->kind = Texture::File;
    t->name = texName;
    if (!LoadTextureFile(*t)) { sc.warnings.push_back(std::string("Error loading texture file ") + texName); return nullptr; }
    sc.textureList.push_back(std::move(t));
    return sc.textureList.back().get();
}

// ReadTexture(XMLElement*) (xmlload.cpp:525-558)
static std::unique_ptr<TextureMap> ReadTexture(Scene &sc, const XmlElement *e)
{
    const char *texName = e->Attribute("texture");
    if (!texName) return nullptr;
    Texture *tex = nullptr;
    if (Compare(texName, "checkerboard")) {
        std::unique_ptr<Texture> c(new Texture);
        c->kind = Texture::Checker;
        c->name = texName;
        for (auto &ch : e->children) {
            if (Compare(ch->name.c_str(), "color1")) { Color col(0, 0, 0); ReadColor(ch.get(), col); c->color1 = col; }
            else if (Compare(ch->name.c_str(), "color2")) { Color col(0, 0, 0); ReadColor(ch.get(), col); c->color2 = col; }
        }
        sc.textureList.push_back(std::move(c));
        tex = sc.textureList.back().get();
    } else tex = ReadTextureFile(sc, texName);
    std::unique_ptr<TextureMap> map(new TextureMap);
    map->texture = tex;
    LoadTransform(map.get(), e);
    return map;
}

struct NodeMtl { Node *node; std::string mtlName; };

// LoadNode (xmlload.cpp:172-271)
static void LoadNode(Scene &sc, std::vector<NodeMtl> &nodeMtl, Node *parent, const XmlElement *e)
{
    parent->children.emplace_back(new Node);
    Node *node = parent->children.back().get();
    const char *name = e->Attribute("name");
    if (name) { node->name = name; node->hasName = true; }
    const char *mtlName = e->Attribute("material");
    if (mtlName) nodeMtl.push_back({node, mtlName});
    const char *type = e->Attribute("type");
    if (type) {
        if (Compare(type, "sphere")) node->obj = &sc.theSphere;
        else if (Compare(type, "plane")) node->obj = &sc.thePlane;
        else if (Compare(type, "obj")) {
            TriMeshObject *obj = nullptr;
            std::string file = name ? name : "";
            for (auto &o : sc.objList) if (o->name == file) obj = o.get();
            if (!obj) {
                std::unique_ptr<TriMeshObject> t(new TriMeshObject);
                t->name = file;
                // O-6 of SURVEY.md 8(c): the repo's scenes use Windows separators
                std::string path = file;
                for (auto &ch : path) if (ch == '\\') ch = '/';
                std::string err;
                if (!t->LoadFromFileObj(path.c_str(), mtlName == nullptr, &err) || t->f.empty()) {
                    sc.warnings.push_back("ERROR: Cannot load file \"" + file + "\"" + (err.empty() ? "" : ": " + err));
                } else {
                    if (t->vn.empty()) t->ComputeNormals();       // objects.h:57
                    t->BuildBVH(4);                               // objects.h:59
                    obj = t.get();
                    // auto MultiMtl from the OBJ's usemtl groups (xmlload.cpp:219-250)
                    if (!t->m.empty() && sc.FindMaterial(file.c_str()) == nullptr) {
                        std::unique_ptr<MultiMtl> mm(new MultiMtl);
                        for (auto &om : t->m) {
                            std::unique_ptr<MtlBlinn> m(new MtlBlinn);
                            m->diffuse.color = Color(om.Kd[0], om.Kd[1], om.Kd[2]);
                            m->specular.color = Color(om.Ks[0], om.Ks[1], om.Ks[2]);
                            m->glossiness = om.Ns;
                            m->ior = om.Ni;
                            auto mapOf = [&](const std::string &f) {
                                std::unique_ptr<TextureMap> tm(new TextureMap);
                                tm->texture = ReadTextureFile(sc, f.c_str());
                                return tm;
                            };
                            if (!om.map_Kd.empty()) m->diffuse.map = mapOf(om.map_Kd);
                            if (!om.map_Ks.empty()) m->diffuse.map = mapOf(om.map_Ks);   // sic: xmlload.cpp:230 sets the DIFFUSE texture
                            if (om.illum > 2 && om.illum <= 7) {
                                m->reflection.color = Color(om.Ks[0], om.Ks[1], om.Ks[2]);
                                float gloss = acosf(powf(2, 1 / om.Ns));
                                m->reflectionGlossiness = gloss;
                                if (om.illum >= 6) {
                                    m->refraction.color = Color(1 - om.Tf[0], 1 - om.Tf[1], 1 - om.Tf[2]);
                                    m->refractionGlossiness = gloss;
                                }
                            }
                            mm->mtls.push_back(std::move(m));
                        }
                        mm->name = file;
                        sc.materials.push_back(std::move(mm));
                        nodeMtl.push_back({node, file});
                    }
                    sc.objList.push_back(std::move(t));
                }
            }
            node->obj = obj;
        }
    }
    for (auto &ch : e->children)
        if (Compare(ch->name.c_str(), "object")) LoadNode(sc, nodeMtl, node, ch.get());
    LoadTransform(node, e);
}

// LoadMaterial (xmlload.cpp:307-390)
static void LoadMaterial(Scene &sc, const XmlElement *e)
{
    const char *name = e->Attribute("name");
    const char *type = e->Attribute("type");
    if (!type || !Compare(type, "blinn")) return;
    std::unique_ptr<MtlBlinn> m(new MtlBlinn);
    for (auto &chp : e->children) {
        const XmlElement *ch = chp.get();
        const char *tag = ch->name.c_str();
        Color c(1, 1, 1);
        float f = 1;
        if (Compare(tag, "diffuse")) { ReadColor(ch, c); m->diffuse.color = c; m->diffuse.map = ReadTexture(sc, ch); }
        else if (Compare(tag, "specular")) { ReadColor(ch, c); m->specular.color = c; m->specular.map = ReadTexture(sc, ch); }
        else if (Compare(tag, "glossiness")) { ReadFloat(ch, f); m->glossiness = f; }
        else if (Compare(tag, "emission")) { ReadColor(ch, c); m->emission.color = c; m->emission.map = ReadTexture(sc, ch); }
        else if (Compare(tag, "reflection")) {
            ReadColor(ch, c); m->reflection.color = c; m->reflection.map = ReadTexture(sc, ch);
            f = 0; ReadFloat(ch, f, "glossiness"); m->reflectionGlossiness = f;
        } else if (Compare(tag, "refraction")) {
            ReadColor(ch, c); m->refraction.color = c;
            ReadFloat(ch, f, "index"); m->ior = f;
            m->refraction.map = ReadTexture(sc, ch);
            f = 0; ReadFloat(ch, f, "glossiness"); m->refractionGlossiness = f;
        } else if (Compare(tag, "absorption")) { ReadColor(ch, c); m->absorption = c; }
    }
    m->name = name ? name : "";
    sc.materials.push_back(std::move(m));
}

// LoadLight (xmlload.cpp:394-474)
static void LoadLight(Scene &sc, const XmlElement *e)
{
    const char *name = e->Attribute("name");
    const char *type = e->Attribute("type");
    if (!type) return;
    std::unique_ptr<Light> l;
    if (Compare(type, "ambient")) l.reset(new Light(Light::Ambient));
    else if (Compare(type, "direct")) l.reset(new Light(Light::Direct));
    else if (Compare(type, "point")) l.reset(new Light(Light::Point));
    else return;
    for (auto &chp : e->children) {
        const XmlElement *ch = chp.get();
        const char *tag = ch->name.c_str();
        if (Compare(tag, "intensity")) { Color c(1, 1, 1); ReadColor(ch, c); l->intensity = c; }
        else if (l->kind == Light::Direct && Compare(tag, "direction")) { Vec3 v(1, 1, 1); ReadVector(ch, v); l->direction = Normalized(v); }
        else if (l->kind == Light::Point && Compare(tag, "position")) { Vec3 v(0, 0, 0); ReadVector(ch, v); l->position = v; }
        else if (l->kind == Light::Point && Compare(tag, "size")) { float f = 0; ReadFloat(ch, f); l->size = f; }
    }
    l->name = name ? name : "";
    sc.lights.push_back(std::move(l));
}

// LoadScene (xmlload.cpp:65-132, 140-168)
int LoadScene(const char *filename, Scene &sc, std::string *err)
{
    auto fail = [&](const std::string &m) { if (err) *err = m; return 0; };
    XmlDocument doc;
    if (!doc.LoadFile(filename)) return fail(std::string("Failed to load the file \"") + filename + "\": " + doc.error);
    const XmlElement *xml = doc.root->FirstChildElement("xml");
    if (!xml) return fail("No \"xml\" tag found.");
    const XmlElement *scene = xml->FirstChildElement("scene");
    if (!scene) return fail("No \"scene\" tag found.");
    const XmlElement *cam = xml->FirstChildElement("camera");
    if (!cam) return fail("No \"camera\" tag found.");

    std::vector<NodeMtl> nodeMtl;
    for (auto &chp : scene->children) {
        const XmlElement *ch = chp.get();
        const char *tag = ch->name.c_str();
        if (Compare(tag, "background")) { Color c(1, 1, 1); ReadColor(ch, c); sc.background.color = c; sc.background.map = ReadTexture(sc, ch); }
        else if (Compare(tag, "environment")) { Color c(1, 1, 1); ReadColor(ch, c); sc.environment.color = c; sc.environment.map = ReadTexture(sc, ch); }
        else if (Compare(tag, "object")) LoadNode(sc, nodeMtl, &sc.rootNode, ch);
        else if (Compare(tag, "material")) LoadMaterial(sc, ch);
        else if (Compare(tag, "light")) LoadLight(sc, ch);
    }
    for (auto &nm : nodeMtl) {
        Material *m = sc.FindMaterial(nm.mtlName.c_str());
        if (m) nm.node->mtl = m;
    }
    // camera (xmlload.cpp:109-127)
    Camera &c = sc.camera;
    c = Camera();
    c.dir = c.dir + c.pos;
    for (auto &chp : cam->children) {
        const XmlElement *ch = chp.get();
        const char *tag = ch->name.c_str();
        if (Compare(tag, "position")) ReadVector(ch, c.pos);
        else if (Compare(tag, "target")) ReadVector(ch, c.dir);
        else if (Compare(tag, "up")) ReadVector(ch, c.up);
        else if (Compare(tag, "fov")) ReadFloat(ch, c.fov);
        else if (Compare(tag, "focaldist")) ReadFloat(ch, c.focaldist);
        else if (Compare(tag, "dof")) ReadFloat(ch, c.dof);
        else if (Compare(tag, "width")) ch->QueryInt("value", &c.imgWidth);
        else if (Compare(tag, "height")) ch->QueryInt("value", &c.imgHeight);
    }
    c.dir = c.dir - c.pos;
    c.dir = Normalized(c.dir);
    Vec3 x = Cross(c.dir, c.up);
    c.up = Normalized(Cross(x, c.dir));
    return 1;
}

}}  // namespace
