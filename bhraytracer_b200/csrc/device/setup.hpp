// setup.hpp -- host-side frame setup shared by render.cu (the library) and tests/emu.
//
// Restates the scalar prologue of BeginRender() (reference Main.cpp:178-200): camera basis,
// topLeft, dd_x, dd_y in the reference's float/double mix; CalculateLightsIntensity()
// (Main.cpp:112-123); and the constant table of the 31 extra texture-filter taps
// (Scenes/scene.h:322-329).  Runs on the CPU in both builds so that these constants are the
// same bits the reference computes ("computed on the host ... and passed in").
#pragma once
#include <cmath>
#include <vector>
#include "core.cuh"

namespace bhrt {

struct FrameConstants {
    V3 cam_pos, top_left, dd_x, dd_y;
};

inline FrameConstants camera_setup(const bhrt_camera &c)
{
    FrameConstants fc;
    V3 pos = ld3(c.pos), dir = ld3(c.dir), up = ld3(c.up);
    float aor = c.width / (float)c.height;
    float tan_h_pov = (float)tan(c.fov / 2 * BHRT_PI_D / 180.0);
    float l = c.focaldist;
    float h = 2 * l * tan_h_pov;
    float w = aor * h;
    V3 camZ = -dir;
    V3 camY = up;
    V3 camX = cross(camY, camZ);
    fc.top_left = pos - camZ * l + camY * h / 2 - camX * w / 2;
    fc.dd_x = camX * w / (float)c.width;
    fc.dd_y = camY * h / (float)c.height;
    fc.cam_pos = pos;
    return fc;
}

// std::sort with CompareGenLight (Main.cpp:112-118).  libstdc++ sorts ranges of <= 16
// elements with a plain insertion sort, which is what this is (BHRT_MAX_LIGHTS = 32 > 16 is
// accepted but then ties may order differently from the reference).
inline void sort_lights(std::vector<bhrt_light> &lights, float &all_intensity)
{
    for (size_t i = 1; i < lights.size(); i++) {
        bhrt_light v = lights[i];
        float gv = gray(ld3(v.intensity));
        if (gv < gray(ld3(lights[0].intensity))) {
            for (size_t k = i; k > 0; k--) lights[k] = lights[k - 1];
            lights[0] = v;
        } else {
            size_t k = i;
            while (gv < gray(ld3(lights[k - 1].intensity))) { lights[k] = lights[k - 1]; k--; }
            lights[k] = v;
        }
    }
    all_intensity = 0;     // the reference's global starts at 0 and BeginRender() runs once
    for (auto &l : lights) all_intensity += gray(ld3(l.intensity));
}

// SceneView::mesh_nodes: node indices of the first BHRT_MESH_MASK_BITS mesh nodes in DFS order
inline void fill_mesh_nodes(SceneView &s, const bhrt_node *host_nodes, int n_nodes)
{
    int n = 0;
    for (int i = 0; i < BHRT_MESH_MASK_BITS; i++) s.mesh_nodes[i] = 0;
    s.n_mesh_nodes = 0;
    for (int k = 0; k < n_nodes; k++)
        if (host_nodes[k].kind == BHRT_OBJ_MESH) { if (n < BHRT_MESH_MASK_BITS) s.mesh_nodes[n++] = k; s.n_mesh_nodes++; }
}

// Point lights in the order BuildCausticPhotonMap uses them (reference Main.cpp:347-361): sorted
// ascending by Gray() * GetSize(), GetSize() returning int (Lights/lights.h:76); `lights` must be
// the already intensity-sorted device table.
inline PhotonEmitCfg photon_emit_cfg(const std::vector<bhrt_light> &lights)
{
    PhotonEmitCfg cfg;
    cfg.n_lights = 0; cfg.sum = 0; cfg.global_map = 0;
    for (size_t i = 0; i < lights.size() && cfg.n_lights < BHRT_MAX_LIGHTS; i++)
        if (lights[i].kind == BHRT_LIGHT_POINT) cfg.light[cfg.n_lights++] = (int)i;
    auto w = [&](int li) { return gray(ld3(lights[li].intensity)) * (float)(int)lights[li].size; };
    for (int i = 1; i < cfg.n_lights; i++) {           // insertion sort = std::sort for <= 16 elements
        int v = cfg.light[i]; int k = i;
        while (k > 0 && w(v) < w(cfg.light[k - 1])) { cfg.light[k] = cfg.light[k - 1]; k--; }
        cfg.light[k] = v;
    }
    for (int i = 0; i < cfg.n_lights; i++) cfg.sum += w(cfg.light[i]);
    return cfg;
}

// Halton(i,2), Halton(i,3) mapped to the elliptical footprint (scene.h:137-146,322-329)
inline float halton(int index, int base)
{
    float r = 0;
    float f = 1.0f / (float)base;
    for (int i = index; i > 0; i /= base) { r += f * (i % base); f /= (float)base; }
    return r;
}
inline std::vector<float> texture_filter_offsets()
{
    std::vector<float> o;
    for (int i = 1; i < BHRT_TEX_SAMPLES; i++) {
        float x = halton(i, 2), y = halton(i, 3);
        float r = sqrtf(x) * 0.5f;
        x = r * sinf(y * (float)M_PI * 2);
        float y2 = r * cosf(y * (float)M_PI * 2);
        o.push_back(x); o.push_back(y2);
    }
    return o;
}

}  // namespace bhrt
