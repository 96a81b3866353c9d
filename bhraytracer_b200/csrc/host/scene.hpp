// scene.hpp -- host-side scene model of the B200 renderer.
//
// Mirrors the reference's Object / Material / Light / Texture / Node abstractions (reference
// BHRayTracer/Scenes/scene.h, Objects/objects.h, Materials/materials.h, Lights/lights.h,
// Textures/texture.h) as plain data: on this side of the C ABI they only describe a scene --
// intersection and shading live in the CUDA kernels.  An existing scene XML loads into this
// model unchanged (loader.cpp) and is then flattened into device tables (flatten.cpp).
//
// The small float math below restates cy::Vec3f / cy::Matrix3f (reference
// DataStructure/cyVector.h, cyMatrix.h) in the SAME operation order, because node transforms
// feed the bit-exact hit-id contract.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include "../../../include/bhrt.h"

namespace bhrt { namespace host {

struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() {}
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(Vec3 a, float v) { return Vec3(a.x * v, a.y * v, a.z * v); }
inline Vec3 operator/(Vec3 a, float v) { return Vec3(a.x / v, a.y / v, a.z / v); }
inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 Cross(Vec3 a, Vec3 p) { return Vec3(a.y * p.z - a.z * p.y, a.z * p.x - a.x * p.z, a.x * p.y - a.y * p.x); }
inline float Length(Vec3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
inline Vec3 Normalized(Vec3 a) { return a / Length(a); }

// column-major 3x3, cell[3*c+r] (cyMatrix.h:388)
struct Mat3 {
    float cell[9];
    void SetIdentity() { for (int i = 0; i < 9; i++) cell[i] = 0; cell[0] = cell[4] = cell[8] = 1; }
    // Matrix3::operator*(Matrix3) (cyMatrix.h:671-683)
    Mat3 operator*(const Mat3 &right) const
    {
        Mat3 r;
        for (int i = 0; i < 9; i += 3)
            for (int j = 0; j < 3; j++) {
                float a = cell[j] * right.cell[i], b = cell[3 + j] * right.cell[i + 1], c = cell[6 + j] * right.cell[i + 2];
                r.cell[i + j] = a + b + c;
            }
        return r;
    }
    // Matrix3::operator*(Vec3) (cyMatrix.h:684-688)
    Vec3 operator*(Vec3 p) const
    {
        return Vec3(p.x * cell[0] + p.y * cell[3] + p.z * cell[6],
                    p.x * cell[1] + p.y * cell[4] + p.z * cell[7],
                    p.x * cell[2] + p.y * cell[5] + p.z * cell[8]);
    }
    // Matrix3::GetInverse (cyMatrix.h:799-821): cofactors, then every cell divided by det
    Mat3 GetInverse() const
    {
        Mat3 inv;
        inv.cell[0] = (cell[4] * cell[8] - cell[5] * cell[7]);
        inv.cell[1] = (cell[2] * cell[7] - cell[1] * cell[8]);
        inv.cell[2] = (cell[1] * cell[5] - cell[2] * cell[4]);
        inv.cell[3] = (cell[5] * cell[6] - cell[3] * cell[8]);
        inv.cell[4] = (cell[0] * cell[8] - cell[2] * cell[6]);
        inv.cell[5] = (cell[2] * cell[3] - cell[0] * cell[5]);
        inv.cell[6] = (cell[3] * cell[7] - cell[4] * cell[6]);
        inv.cell[7] = (cell[1] * cell[6] - cell[0] * cell[7]);
        inv.cell[8] = (cell[0] * cell[4] - cell[1] * cell[3]);
        float det = cell[0] * inv.cell[0] + cell[1] * inv.cell[3] + cell[2] * inv.cell[6];
        for (int i = 0; i < 9; i++) inv.cell[i] /= det;
        return inv;
    }
    // Matrix3::SetRotation(axis, sin, cos) (cyMatrix.h:522-533)
    void SetRotation(Vec3 axis, float sinAngle, float cosAngle)
    {
        float t = 1.0f - cosAngle;
        Vec3 a = axis * t;
        float txy = a.x * axis.y, txz = a.x * axis.z, tyz = a.y * axis.z;
        Vec3 s = axis * sinAngle;
        cell[0] = a.x * axis.x + cosAngle; cell[1] = txy + s.z;               cell[2] = txz - s.y;
        cell[3] = txy - s.z;               cell[4] = a.y * axis.y + cosAngle; cell[5] = tyz + s.x;
        cell[6] = txz + s.y;               cell[7] = tyz - s.x;               cell[8] = a.z * axis.z + cosAngle;
    }
};

// reference Scenes/scene.h:208-246
class Transformation {
public:
    Transformation() { tm.SetIdentity(); itm.SetIdentity(); }
    const Mat3 &GetTransform() const { return tm; }
    const Vec3 &GetPosition() const { return pos; }
    const Mat3 &GetInverseTransform() const { return itm; }
    Vec3 TransformFrom(Vec3 p) const { return tm * p + pos; }
    void Translate(Vec3 p) { pos = pos + p; }
    void Rotate(Vec3 axis, float degrees)
    {
        Mat3 m;
        float angle = degrees * (float)M_PI / 180.0f;
        m.SetRotation(axis, std::sin(angle), std::cos(angle));
        Transform(m);
    }
    void Scale(float sx, float sy, float sz)
    {
        Mat3 m;
        for (int i = 0; i < 9; i++) m.cell[i] = 0;
        m.cell[0] = sx; m.cell[4] = sy; m.cell[8] = sz;
        Transform(m);
    }
    void Transform(const Mat3 &m) { tm = m * tm; pos = m * pos; itm = tm.GetInverse(); }
    void FillXform(bhrt_xform &x) const
    {
        for (int i = 0; i < 9; i++) { x.tm[i] = tm.cell[i]; x.itm[i] = itm.cell[i]; }
        x.pos[0] = pos.x; x.pos[1] = pos.y; x.pos[2] = pos.z;
    }
private:
    Mat3 tm, itm;
    Vec3 pos;
};

struct Color { float r = 0, g = 0, b = 0; Color() {} Color(float a, float c, float d) : r(a), g(c), b(d) {} };

// ---- textures (reference Textures/texture.h) ---------------------------------------------
struct Texture {
    enum Kind { File, Checker } kind = File;
    std::string name;
    int width = 0, height = 0;
    std::vector<uint8_t> rgb;           // File: RGB8 texels (TextureFile::data)
    Color color1{0, 0, 0}, color2{1, 1, 1};   // Checker defaults (texture.h:38)
    int index = -1;                     // slot in the flattened table
};
struct TextureMap : Transformation {     // scene.h:364-386
    Texture *texture = nullptr;
};
struct TexturedColor {                   // scene.h:394-422
    Color color{0, 0, 0};
    std::unique_ptr<TextureMap> map;
};

// ---- objects (reference Objects/objects.h) --------------------------------------------------
struct TriFace { uint32_t v[3]; };
struct ObjMtl {                          // cyTriMesh.h:74-103
    std::string name;
    float Ka[3] = {0, 0, 0}, Kd[3] = {1, 1, 1}, Ks[3] = {0, 0, 0}, Tf[3] = {0, 0, 0};
    float Ns = 0, Ni = 1;
    int illum = 2;
    std::string map_Kd, map_Ks;
};
struct BvhNode { float box[6]; uint32_t data; };

struct Object {
    enum Kind { Sphere, Plane, TriObj } kind;
    explicit Object(Kind k) : kind(k) {}
    virtual ~Object() {}
};
struct TriMeshObject : Object {          // TriObj = cyTriMesh + cyBVHTriMesh
    TriMeshObject() : Object(TriObj) {}
    std::string name;
    std::vector<Vec3> v, vn, vt;
    std::vector<TriFace> f, fn, ft;
    std::vector<ObjMtl> m;
    std::vector<int> mcfc;
    std::vector<BvhNode> bvhNodes;       // [0] unused, root = 1 (cyBVH.h:76)
    std::vector<uint32_t> bvhElements;
    bool vtSynthesised = false;          // OBJ had no vt: the reference would crash (TriObj.cpp:177)
    int index = -1;
    bool LoadFromFileObj(const char *filename, bool loadMtl, std::string *err);
    void ComputeNormals();
    void BuildBVH(unsigned maxElementsPerNode = 4);   // objects.h:59
};

// ---- materials (reference Materials/materials.h) ---------------------------------------------
struct Material {
    enum Kind { Blinn, Multi } kind;
    std::string name;
    explicit Material(Kind k) : kind(k) {}
    virtual ~Material() {}
};
struct MtlBlinn : Material {
    MtlBlinn() : Material(Blinn) { diffuse.color = Color(0.5f, 0.5f, 0.5f); specular.color = Color(0.7f, 0.7f, 0.7f); }
    TexturedColor diffuse, specular, reflection, refraction, emission;
    float glossiness = 20.0f;
    Color absorption{0, 0, 0};
    float ior = 1;
    float reflectionGlossiness = 0, refractionGlossiness = 0;
};
struct MultiMtl : Material {
    MultiMtl() : Material(Multi) {}
    std::vector<std::unique_ptr<MtlBlinn>> mtls;
};

// ---- lights (reference Lights/lights.h) --------------------------------------------------------
struct Light {
    enum Kind { Ambient, Direct, Point } kind;
    std::string name;
    Color intensity{0, 0, 0};
    Vec3 direction{0, 0, 1};
    Vec3 position{0, 0, 0};
    float size = 0;
    explicit Light(Kind k) : kind(k) {}
};

// ---- scene graph (reference Scenes/scene.h:426-502) --------------------------------------------
struct Node : Transformation {
    std::string name;
    bool hasName = false;
    std::vector<std::unique_ptr<Node>> children;
    Object *obj = nullptr;               // not owned
    Material *mtl = nullptr;             // not owned
};

struct Camera {                           // scene.h:506-524
    Vec3 pos{0, 0, 0}, dir{0, 0, -1}, up{0, 1, 0};
    float fov = 40, focaldist = 1, dof = 0;
    int imgWidth = 200, imgHeight = 150;
};

// Everything the reference keeps in globals (Main.cpp:17-30)
struct Scene {
    Node rootNode;
    Camera camera;
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<std::unique_ptr<Light>> lights;
    std::vector<std::unique_ptr<TriMeshObject>> objList;
    std::vector<std::unique_ptr<Texture>> textureList;
    TexturedColor background, environment;
    Object theSphere{Object::Sphere}, thePlane{Object::Plane};
    std::vector<std::string> warnings;

    Material *FindMaterial(const char *name)
    {
        for (auto &m : materials) if (m && m->name == name) return m.get();
        return nullptr;
    }
};

// LoadScene (reference xmlload.cpp:65-132).  Returns 1 on success, 0 on failure like the
// reference; *err receives the message the reference would printf.
int LoadScene(const char *filename, Scene &scene, std::string *err);

// Flattened, upload-ready form.  Owns every table the bhrt_scene_desc points into.
struct FlatScene {
    std::vector<bhrt_node> nodes;
    std::vector<std::string> nodeNames;
    std::vector<bhrt_material> materials;
    std::vector<bhrt_light> lights;
    std::vector<bhrt_texmap> texmaps;
    std::vector<bhrt_texture> textures;
    std::vector<bhrt_mesh> meshes;
    std::vector<uint8_t> texels;
    std::vector<uint8_t> meshData;
    bhrt_scene_desc desc;
};
// Returns false (with *err) when the scene cannot be rendered (e.g. an object without material:
// the reference dereferences a null Material there, Main.cpp:161).
bool Flatten(Scene &scene, FlatScene &flat, std::string *err);

}}  // namespace
