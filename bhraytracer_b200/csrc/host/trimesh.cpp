// trimesh.cpp -- OBJ/MTL ingestion, vertex normals and the reference-topology BVH.
//
// Restates, on the host: cy::TriMesh::LoadFromFileObj / ComputeNormals (reference
// Objects/TriObj/cyTriMesh.h:248-547) and cy::BVH::Build / MeanSplit / ConvertTempData with
// the BVHTriMesh element callbacks (reference DataStructure/cyBVH.h:122-142,242-328,356-375).
// The BVH must have the reference's exact topology and node numbering: its mesh shadow test
// stops at the first leaf, in child order, that holds a front-face hit (TriObj.cpp:272-307),
// so the image depends on it.
#include <cctype>
#include <cstdio>
#include <cstring>
#include "scene.hpp"

namespace bhrt { namespace host {

namespace {

// Buffer::ReadLine (cyTriMesh.h:278-302): skips blank space and '#' lines, collapses runs of
// white space into one ' ', at most 1023 characters.
struct LineReader {
    char data[1024];
    int ReadLine(FILE *fp)
    {
        int c = fgetc(fp);
        while (!feof(fp)) {
            while (c != EOF && isspace(c)) c = fgetc(fp);
            if (c == '#') { while (!feof(fp) && c != '\n' && c != '\r' && c != '\0') c = fgetc(fp); }
            else break;
        }
        int i = 0;
        bool inspace = false;
        while (i < 1024 - 1) {
            if (feof(fp) || c == EOF || c == '\n' || c == '\r' || c == '\0') break;
            if (isspace(c)) inspace = true;
            else {
                if (inspace) data[i++] = ' ';
                inspace = false;
                if (i < 1024 - 1) data[i++] = (char)c;
            }
            c = fgetc(fp);
        }
        data[i] = '\0';
        return i;
    }
    bool IsCommand(const char *cmd) const
    {
        int i = 0;
        while (cmd[i] != '\0') { if (cmd[i] != data[i]) return false; i++; }
        return data[i] == '\0' || data[i] == ' ';
    }
    void ReadVertex(Vec3 &v) const { v = Vec3(0, 0, 0); sscanf(data + 2, "%f %f %f", &v.x, &v.y, &v.z); }
    void ReadFloat3(float f[3]) const { f[2] = f[1] = f[0] = 0; int n = sscanf(data + 2, "%f %f %f", &f[0], &f[1], &f[2]); if (n == 1) f[2] = f[1] = f[0]; }
    std::string Rest(int start) const
    {
        if ((int)strlen(data) < start) return "";
        const char *p = data + start;
        while (*p != '\0' && *p <= ' ') p++;
        return p;
    }
};

struct MtlData { std::string name; unsigned firstFace = 0, faceCount = 0; };

}  // namespace

bool TriMeshObject::LoadFromFileObj(const char *filename, bool loadMtl, std::string *err)
{
    FILE *fp = fopen(filename, "r");
    if (!fp) { if (err) *err = std::string("Cannot open file ") + filename; return false; }
    LineReader buffer;
    std::vector<MtlData> mtlData;
    std::vector<Vec3> _v, _vn, _vt;
    std::vector<TriFace> _f, _fn, _ft;
    std::vector<std::string> mtlFiles;
    std::vector<int> faceMtlIndex;
    int currentMtlIndex = -1;
    bool hasTextures = false, hasNormals = false;

    while (int rb = buffer.ReadLine(fp)) {
        if (buffer.IsCommand("v")) { Vec3 p; buffer.ReadVertex(p); _v.push_back(p); }
        else if (buffer.IsCommand("vt")) { Vec3 p; buffer.ReadVertex(p); _vt.push_back(p); hasTextures = true; }
        else if (buffer.IsCommand("vn")) { Vec3 p; buffer.ReadVertex(p); _vn.push_back(p); hasNormals = true; }
        else if (buffer.IsCommand("f")) {
            // polygon -> triangle fan, indices 1-based or negative (cyTriMesh.h:379-438)
            int facevert = -1;
            bool inspace = true, negative = false;
            int type = 0;
            unsigned index = 0;
            TriFace face = {{0, 0, 0}}, tf = {{0, 0, 0}}, nf = {{0, 0, 0}};
            unsigned nFacesBefore = (unsigned)_f.size();
            for (int i = 2; i < rb; i++) {
                char ch = buffer.data[i];
                if (ch == ' ') { inspace = true; continue; }
                if (inspace) {
                    inspace = false; negative = false; type = 0; index = 0;
                    if (facevert < 2) facevert++;
                    else {
                        _f.push_back(face); face.v[1] = face.v[2];
                        if (hasTextures) { _ft.push_back(tf); tf.v[1] = tf.v[2]; }
                        if (hasNormals) { _fn.push_back(nf); nf.v[1] = nf.v[2]; }
                        faceMtlIndex.push_back(currentMtlIndex);
                    }
                }
                if (ch == '/') { type++; index = 0; }
                if (ch == '-') negative = true;
                if (ch >= '0' && ch <= '9') {
                    index = index * 10 + (unsigned)(ch - '0');
                    switch (type) {
                    case 0: face.v[facevert] = negative ? (unsigned)_v.size() - index : index - 1; break;
                    case 1: tf.v[facevert] = negative ? (unsigned)_vt.size() - index : index - 1; hasTextures = true; break;
                    case 2: nf.v[facevert] = negative ? (unsigned)_vn.size() - index : index - 1; hasNormals = true; break;
                    }
                }
            }
            _f.push_back(face);
            if (hasTextures) _ft.push_back(tf);
            if (hasNormals) _fn.push_back(nf);
            faceMtlIndex.push_back(currentMtlIndex);
            if (currentMtlIndex >= 0) mtlData[currentMtlIndex].faceCount += (unsigned)_f.size() - nFacesBefore;
        } else if (loadMtl) {
            if (buffer.IsCommand("usemtl")) {
                std::string nm = (int)strlen(buffer.data) >= 7 ? std::string(buffer.data + 7) : std::string();
                if (nm.empty()) currentMtlIndex = 0;
                else {
                    int idx = -1;
                    for (size_t k = 0; k < mtlData.size(); k++) if (mtlData[k].name == nm) idx = (int)k;
                    if (idx < 0) { MtlData md; md.name = nm; md.firstFace = (unsigned)_f.size(); mtlData.push_back(md); idx = (int)mtlData.size() - 1; }
                    currentMtlIndex = idx;
                }
                if (currentMtlIndex >= (int)mtlData.size()) currentMtlIndex = -1;
            }
            if (buffer.IsCommand("mtllib") && (int)strlen(buffer.data) >= 7) mtlFiles.push_back(buffer.data + 7);
        }
        if (feof(fp)) break;
    }
    fclose(fp);
    if (_f.empty()) return true;

    // every index must be in range: the reference would read out of bounds instead
    for (auto &t : _f) for (int k = 0; k < 3; k++) if (t.v[k] >= _v.size()) { if (err) *err = "face index out of range"; return false; }
    for (auto &t : _ft) for (int k = 0; k < 3; k++) if (t.v[k] >= _vt.size()) { if (err) *err = "texture index out of range"; return false; }
    for (auto &t : _fn) for (int k = 0; k < 3; k++) if (t.v[k] >= _vn.size()) { if (err) *err = "normal index out of range"; return false; }

    v = _v; vt = _vt; vn = _vn;
    size_t nfaces = _f.size();
    f.assign(nfaces, TriFace{{0, 0, 0}});
    bool haveFt = !_vt.empty(), haveFn = !_vn.empty();
    if (haveFt) { ft.assign(nfaces, TriFace{{0, 0, 0}}); _ft.resize(nfaces, TriFace{{0, 0, 0}}); }
    if (haveFn) { fn.assign(nfaces, TriFace{{0, 0, 0}}); _fn.resize(nfaces, TriFace{{0, 0, 0}}); }
    if (loadMtl) { m.assign(mtlData.size(), ObjMtl()); mcfc.assign(mtlData.size(), 0); }

    if (!mtlData.empty()) {
        // faces regrouped by material (cyTriMesh.h:467-490)
        unsigned fid = 0;
        for (int mi = 0; mi < (int)mtlData.size(); mi++) {
            for (unsigned i = mtlData[mi].firstFace, j = 0; j < mtlData[mi].faceCount && i < nfaces; i++) {
                if (faceMtlIndex[i] == mi) {
                    f[fid] = _f[i];
                    if (haveFn) fn[fid] = _fn[i];
                    if (haveFt) ft[fid] = _ft[i];
                    fid++; j++;
                }
            }
            mcfc[mi] = (int)fid;
        }
        if (fid < nfaces) {
            for (unsigned i = 0; i < nfaces && fid < nfaces; i++) {
                if (faceMtlIndex[i] < 0) {
                    f[fid] = _f[i];
                    if (haveFn) fn[fid] = _fn[i];
                    if (haveFt) ft[fid] = _ft[i];
                    fid++;
                }
            }
        }
    } else {
        f = _f;
        if (haveFt) ft = _ft;
        if (haveFn) fn = _fn;
    }

    if (loadMtl) {
        std::string dir;
        const char *pathEnd = strrchr(filename, '\\');
        if (!pathEnd) pathEnd = strrchr(filename, '/');
        if (pathEnd) dir.assign(filename, pathEnd - filename + 1);
        for (auto &lib : mtlFiles) {
            std::string mf = dir + lib;
            FILE *mp = fopen(mf.c_str(), "r");
            if (!mp) continue;
            int mtlID = -1;
            while (buffer.ReadLine(mp)) {
                if (buffer.IsCommand("newmtl")) {
                    std::string nm = (int)strlen(buffer.data) >= 7 ? std::string(buffer.data + 7) : std::string();
                    mtlID = -1;
                    for (size_t k = 0; k < mtlData.size(); k++) if (mtlData[k].name == nm) mtlID = (int)k;
                    if (mtlID >= 0) m[mtlID].name = buffer.Rest(7);
                } else if (mtlID >= 0) {
                    ObjMtl &mm = m[mtlID];
                    if (buffer.IsCommand("Ka")) buffer.ReadFloat3(mm.Ka);
                    else if (buffer.IsCommand("Kd")) buffer.ReadFloat3(mm.Kd);
                    else if (buffer.IsCommand("Ks")) buffer.ReadFloat3(mm.Ks);
                    else if (buffer.IsCommand("Tf")) buffer.ReadFloat3(mm.Tf);
                    else if (buffer.IsCommand("Ns")) sscanf(buffer.data + 2, "%f", &mm.Ns);
                    else if (buffer.IsCommand("Ni")) sscanf(buffer.data + 2, "%f", &mm.Ni);
                    else if (buffer.IsCommand("illum")) sscanf(buffer.data + 5, "%d", &mm.illum);
                    else if (buffer.IsCommand("map_Kd")) mm.map_Kd = buffer.Rest(7);
                    else if (buffer.IsCommand("map_Ks")) mm.map_Ks = buffer.Rest(7);
                }
                if (feof(mp)) break;
            }
            fclose(mp);
        }
    }

    if (vt.empty()) {
        // TriObj::IntersectTriangle dereferences ft unconditionally (TriObj.cpp:177); the
        // reference segfaults on such a mesh.  We give it a single (0,0,0) texture vertex.
        vt.push_back(Vec3(0, 0, 0));
        ft.assign(nfaces, TriFace{{0, 0, 0}});
        vtSynthesised = true;
    }
    return true;
}

// TriMesh::ComputeNormals (cyTriMesh.h:248-261)
void TriMeshObject::ComputeNormals()
{
    vn.assign(v.size(), Vec3(0, 0, 0));
    fn.assign(f.size(), TriFace{{0, 0, 0}});
    for (size_t i = 0; i < f.size(); i++) {
        Vec3 N = Cross(v[f[i].v[1]] - v[f[i].v[0]], v[f[i].v[2]] - v[f[i].v[0]]);
        vn[f[i].v[0]] = vn[f[i].v[0]] + N;
        vn[f[i].v[1]] = vn[f[i].v[1]] + N;
        vn[f[i].v[2]] = vn[f[i].v[2]] + N;
        fn[i] = f[i];
    }
    for (auto &n : vn) n = n / Length(n);
}

namespace {
struct Box6 {
    float b[6];
    Box6() { b[0] = b[1] = b[2] = 1e30f; b[3] = b[4] = b[5] = -1e30f; }
    void operator+=(const Box6 &o) { for (int i = 0; i < 3; i++) { if (b[i] > o.b[i]) b[i] = o.b[i]; if (b[i + 3] < o.b[i + 3]) b[i + 3] = o.b[i + 3]; } }
};
struct TempNode {
    int child1 = -1, child2 = -1;
    Box6 box;
    unsigned count = 0, offset = 0;
};

struct Builder {
    const TriMeshObject &mesh;
    std::vector<uint32_t> &elements;
    std::vector<TempNode> temp;
    explicit Builder(const TriMeshObject &m, std::vector<uint32_t> &e) : mesh(m), elements(e) {}

    // BVHTriMesh::GetElementBounds (cyBVH.h:356-368)
    Box6 ElementBounds(unsigned i) const
    {
        Box6 bx;
        const TriFace &fc = mesh.f[i];
        Vec3 p = mesh.v[fc.v[0]];
        bx.b[0] = bx.b[3] = p.x; bx.b[1] = bx.b[4] = p.y; bx.b[2] = bx.b[5] = p.z;
        for (int j = 1; j < 3; j++) {
            Vec3 q = mesh.v[fc.v[j]];
            for (int k = 0; k < 3; k++) { if (bx.b[k] > q[k]) bx.b[k] = q[k]; if (bx.b[k + 3] < q[k]) bx.b[k + 3] = q[k]; }
        }
        return bx;
    }
    // BVHTriMesh::GetElementCenter (cyBVH.h:371-375)
    float ElementCenter(unsigned i, int dim) const
    {
        const TriFace &fc = mesh.f[i];
        return (mesh.v[fc.v[0]][dim] + mesh.v[fc.v[1]][dim] + mesh.v[fc.v[2]][dim]) / 3.0f;
    }
    // BVH::MeanSplit (cyBVH.h:295-328)
    unsigned MeanSplit(unsigned elementCount, uint32_t *nodeElements, const float *box, unsigned maxPerNode) const
    {
        if (elementCount <= maxPerNode) return 0;
        float d[3] = { box[3] - box[0], box[4] - box[1], box[5] - box[2] };
        unsigned sd[3];
        sd[0] = d[0] >= d[1] ? (d[0] >= d[2] ? 0 : 2) : (d[1] >= d[2] ? 1 : 2);
        sd[1] = (sd[0] + 1) % 3;
        sd[2] = (sd[0] + 2) % 3;
        if (d[sd[1]] < d[sd[2]]) { unsigned t = sd[1]; sd[1] = sd[2]; sd[2] = t; }
        unsigned child1 = 0;
        for (int s = 0; s < 3; s++) {
            unsigned dim = sd[s];
            float splitPos = 0.5f * (box[dim] + box[dim + 3]);
            unsigned i = 0, j = elementCount;
            while (i < j) {
                float center = ElementCenter(nodeElements[i], (int)dim);
                if (center <= splitPos) i++;
                else { j--; uint32_t t = nodeElements[i]; nodeElements[i] = nodeElements[j]; nodeElements[j] = t; }
            }
            if (i < elementCount && i > 0) { child1 = i; break; }
        }
        return child1;
    }
    // BVH::SplitTempNode (cyBVH.h:242-278), CY_BVH_MAX_ELEMENT_COUNT = 8
    void Split(int ti, unsigned maxPerNode)
    {
        unsigned count = temp[ti].count, offset = temp[ti].offset;
        uint32_t *nodeElements = &elements[offset];
        Box6 box = temp[ti].box;
        unsigned c1 = MeanSplit(count, nodeElements, box.b, maxPerNode);
        if (c1 == 0 || c1 >= count) {
            if (count > 8) c1 = count / 2;
            else return;
        }
        Box6 b1, b2;
        for (unsigned i = 0; i < c1; i++) b1 += ElementBounds(nodeElements[i]);
        for (unsigned i = c1; i < count; i++) b2 += ElementBounds(nodeElements[i]);
        TempNode n1, n2;
        n1.count = c1; n1.offset = offset; n1.box = b1;
        n2.count = count - c1; n2.offset = offset + c1; n2.box = b2;
        int i1 = (int)temp.size(); temp.push_back(n1);
        int i2 = (int)temp.size(); temp.push_back(n2);
        temp[ti].child1 = i1; temp[ti].child2 = i2;
        Split(i1, maxPerNode);
        Split(i2, maxPerNode);
    }
    // BVH::ConvertTempData (cyBVH.h:281-291)
    unsigned Convert(std::vector<BvhNode> &nodes, unsigned nodeID, int ti, unsigned childIndex) const
    {
        const TempNode &t = temp[ti];
        if (nodes.size() <= nodeID) nodes.resize(nodeID + 1);
        for (int k = 0; k < 6; k++) nodes[nodeID].box[k] = t.box.b[k];
        if (t.child1 < 0) {
            nodes[nodeID].data = (t.offset & 0x0fffffffu) | ((t.count - 1) << 28) | 0x80000000u;
            return childIndex;
        }
        nodes[nodeID].data = childIndex & 0x7fffffffu;
        unsigned next = Convert(nodes, childIndex, t.child1, childIndex + 2);
        return Convert(nodes, childIndex + 1, t.child2, next);
    }
};
}  // namespace

// BVH::Build (cyBVH.h:122-142) via BVHTriMesh::SetMesh(this, 4) (objects.h:59)
void TriMeshObject::BuildBVH(unsigned maxElementsPerNode)
{
    bvhNodes.clear();
    bvhElements.clear();
    unsigned n = (unsigned)f.size();
    if (n == 0) return;
    if (maxElementsPerNode > 8) maxElementsPerNode = 8;
    bvhElements.resize(n);
    for (unsigned i = 0; i < n; i++) bvhElements[i] = i;
    Builder b(*this, bvhElements);
    Box6 box;
    for (unsigned i = 0; i < n; i++) box += b.ElementBounds(i);
    TempNode root;
    root.count = n; root.offset = 0; root.box = box;
    b.temp.reserve(2 * n + 2);
    b.temp.push_back(root);
    b.Split(0, maxElementsPerNode);
    bvhNodes.assign(b.temp.size() + 1, BvhNode());
    for (auto &nd : bvhNodes) { for (int k = 0; k < 6; k++) nd.box[k] = 0; nd.data = 0x80000000u; }
    b.Convert(bvhNodes, 1, 0, 2);
}

}}  // namespace
