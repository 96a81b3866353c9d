// TEST-ONLY harness: compiles the device code of the renderer (csrc/device/core.cuh, the very
// source nvcc builds into the sm_100a kernels) with g++ and runs it per pixel on the CPU, so
// that the CPU-only CI (`pytest -m "not gpu"`) can compare the kernel LOGIC with the oracle.
// It is never linked into libbhrt.so / libbhrt_host.so and nothing in the product loads it.
//
// usage: emu_render <scene.xml> <out_prefix> [--det] [--spp N] [--gi N] [--seed S] [--size W H]
//                   [--rows B E STRIDE] [--samples B E]
//                   [--emit N FILE]                 emit N caustic photons (unscaled power) into FILE and exit
//                   [--emit-global N FILE]          the same for the GLOBAL photon map (BuildPhotonMap, Main.cpp:251-317)
//                   [--photons FILE TOTAL K RADIUS] render with the caustic gather term over the
//                                                   24-byte records of FILE, powers scaled by 1/TOTAL
//                   [--queries IN OUT]              with --photons: evaluate the irradiance estimate at the
//                                                   (pos, normal) float[6] records of IN, write float[6]
//                                                   (irradiance rgb, mean direction) per query to OUT, exit
// writes <out_prefix>.aov in the layout of oracle/ref_build/driver.cpp and <out_prefix>.acc, the
// float4 accumulation buffer (r,g,b sums + sample count; zeros outside the rendered share) a
// GPU rank would hold after bhrt_render_device with the same row / sample share.
#include <omp.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <vector>
#ifdef BHRT_EMU_STATS
static thread_local unsigned long long g_trav[4];
#define BHRT_TRAV_HOOK(is_leaf, is_shadow) g_trav[((is_shadow) ? 2 : 0) + ((is_leaf) ? 1 : 0)]++
// per-query histogram of BVH steps (log2 buckets), [kind-1][bucket]; merged at exit
static unsigned long long g_hist[2][24], g_sum[2][2], g_nq[2];
static inline void emu_query_done(int kind)
{
    const int k = kind - 1;
    unsigned long long n = g_trav[2 * k] + g_trav[2 * k + 1];
    int b = 0; while ((1ull << b) <= n && b < 23) b++;
#pragma omp atomic
    g_hist[k][b]++;
#pragma omp atomic
    g_sum[k][0] += g_trav[2 * k];
#pragma omp atomic
    g_sum[k][1] += g_trav[2 * k + 1];
#pragma omp atomic
    g_nq[k]++;
    g_trav[0] = g_trav[1] = g_trav[2] = g_trav[3] = 0;
}
#define BHRT_QUERY_HOOK(kind) emu_query_done(kind)
#endif
#include "../../bhraytracer_b200/csrc/device/core.cuh"
#include "../../bhraytracer_b200/csrc/device/setup.hpp"
#include "../../bhraytracer_b200/csrc/device/accel.hpp"
#include "../../bhraytracer_b200/csrc/device/photon_build.hpp"
#include "../../bhraytracer_b200/csrc/host/scene.hpp"

using namespace bhrt;

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.xml out_prefix [--det] [--spp N] [--gi N] [--seed S] [--size W H]\n", argv[0]); return 2; }
    int det = 0, spp = 32, gi = 3, W = 0, H = 0;
    int row_b = 0, row_e = 0, row_s = 1, smp_b = 0, smp_e = 0;
    int emit_n = 0, emit_global = 0; const char *emit_file = nullptr;
    const char *photon_file = nullptr; int photon_total = 1, photon_k = 1000; float photon_radius = 0.5f;
    const char *query_in = nullptr, *query_out = nullptr;
    const char *pixel_rays = nullptr;
    int max_frames = 0, wide = 0, face_materials = 0;
    unsigned long long seed = 1;
    for (int a = 3; a < argc; a++) {
        if (!strcmp(argv[a], "--det")) det = 1;
        else if (!strcmp(argv[a], "--spp") && a + 1 < argc) spp = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--gi") && a + 1 < argc) gi = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--seed") && a + 1 < argc) seed = strtoull(argv[++a], nullptr, 0);
        else if (!strcmp(argv[a], "--size") && a + 2 < argc) { W = atoi(argv[++a]); H = atoi(argv[++a]); }
        else if (!strcmp(argv[a], "--rows") && a + 3 < argc) { row_b = atoi(argv[++a]); row_e = atoi(argv[++a]); row_s = atoi(argv[++a]); }
        else if (!strcmp(argv[a], "--samples") && a + 2 < argc) { smp_b = atoi(argv[++a]); smp_e = atoi(argv[++a]); }
        else if (!strcmp(argv[a], "--emit") && a + 2 < argc) { emit_n = atoi(argv[++a]); emit_file = argv[++a]; }
        else if (!strcmp(argv[a], "--emit-global") && a + 2 < argc) { emit_n = atoi(argv[++a]); emit_file = argv[++a]; emit_global = 1; }
        else if (!strcmp(argv[a], "--photons") && a + 4 < argc) { photon_file = argv[++a]; photon_total = atoi(argv[++a]); photon_k = atoi(argv[++a]); photon_radius = (float)atof(argv[++a]); }
        else if (!strcmp(argv[a], "--pixel-rays") && a + 1 < argc) pixel_rays = argv[++a];
        else if (!strcmp(argv[a], "--max-frames") && a + 1 < argc) max_frames = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--wide")) wide = 1;
        else if (!strcmp(argv[a], "--face-materials")) face_materials = 1;
        else if (!strcmp(argv[a], "--queries") && a + 2 < argc) { query_in = argv[++a]; query_out = argv[++a]; }
        else { fprintf(stderr, "unknown arg %s\n", argv[a]); return 2; }
    }
    if (det) spp = 1;
    host::Scene sc;
    host::FlatScene flat;
    std::string err;
    if (!host::LoadScene(argv[1], sc, &err)) { fprintf(stderr, "LoadScene: %s\n", err.c_str()); return 3; }
    if (!host::Flatten(sc, flat, &err)) { fprintf(stderr, "Flatten: %s\n", err.c_str()); return 3; }
    for (auto &w : sc.warnings) fprintf(stderr, "warning: %s\n", w.c_str());
    bhrt_scene_desc &D = flat.desc;
    if (W > 0) { D.camera.width = W; D.camera.height = H; }
    W = D.camera.width; H = D.camera.height;

    std::vector<bhrt_light> lights(D.lights, D.lights + D.n_lights);
    float allI = 0;
    sort_lights(lights, allI);
    std::vector<float> offs = texture_filter_offsets();
    FrameConstants fc = camera_setup(D.camera);

    SceneView s;
    memset(&s, 0, sizeof(s));
    s.n_nodes = D.n_nodes; s.n_lights = D.n_lights;
    s.max_depth = 1;
    for (int i = 0; i < D.n_nodes; i++) if (D.nodes[i].depth > s.max_depth) s.max_depth = D.nodes[i].depth;
    s.nodes = D.nodes; s.materials = D.materials; s.lights = lights.data();
    s.texmaps = D.texmaps; s.textures = D.textures; s.texels = D.texels;
    s.meshes = D.meshes; s.mesh_data = D.mesh_data; s.tex_offsets = offs.data();
    AccelBlob accel;
    if (!build_accel(D, accel, &err)) { fprintf(stderr, "accel: %s\n", err.c_str()); return 3; }
    s.accels = accel.meshes.data(); s.accel_data = accel.data.data();
    s.face_materials = face_materials;          // BHRT_FLAG_FACE_MATERIALS (bhrt.h)
    if (wide) {
        // traverse the 8-wide collapse of the BVH (the structure the kernels' fast build uses) instead of the binary tree
        s.use_wide = 1;
        for (auto &A : accel.meshes) {
            if (!A.wide_stack) { fprintf(stderr, "--wide: a mesh has no wide tree\n"); return 3; }
            fprintf(stderr, "wide tree: %u nodes, worst-case stack %u\n", A.wide_nodes, A.wide_stack);
        }
    }
    s.background = D.background; s.environment = D.environment;
    s.all_light_intensity = allI;
    s.cam_pos = fc.cam_pos; s.top_left = fc.top_left; s.dd_x = fc.dd_x; s.dd_y = fc.dd_y;
    s.width = W; s.height = H;
    fill_mesh_nodes(s, D.nodes, D.n_nodes);

    if (emit_n > 0) {
        // BuildCausticPhotonMap's serial while loop (reference Main.cpp:363-379) over the same coroutine
        // the emission kernel runs
        PhotonEmitCfg ecfg = photon_emit_cfg(lights);
        if (ecfg.n_lights == 0) { fprintf(stderr, "no point light\n"); return 3; }
        std::vector<PhotonRecord> recs;
        unsigned long long rays = 0, paths = 0;
        PhotonLane P;
        while ((int)recs.size() < emit_n) {
            P.rng.seed(seed, paths, 0x50484f54u);
            paths++;
            photon_begin(P, s, ecfg, rays);
            for (;;) {
                P.hit_ok = scene_trace(s, P.q, P.hit);
                PhotonRecord r; bool store;
                bool ended = emit_global ? photon_advance_global(P, s, r, store, rays) : photon_advance(P, s, r, store, rays);
                if (store && (int)recs.size() < emit_n) recs.push_back(r);
                if (ended) break;
            }
        }
        FILE *fp = fopen(emit_file, "wb");
        if (!fp) return 4;
        fwrite(recs.data(), sizeof(PhotonRecord), recs.size(), fp);
        fclose(fp);
        printf("{\"emu\": true, \"photons\": %d, \"paths\": %llu, \"photon_rays\": %llu}\n", (int)recs.size(), paths, rays);
        return 0;
    }
    PhotonGather pg; memset(&pg, 0, sizeof(pg));
    HostPhotonMap pmap;
    if (photon_file) {
        FILE *fp = fopen(photon_file, "rb");
        if (!fp) { fprintf(stderr, "cannot read %s\n", photon_file); return 3; }
        std::vector<PhotonRecord> recs;
        PhotonRecord r;
        while (fread(&r, sizeof(r), 1, fp) == 1) recs.push_back(r);
        fclose(fp);
        build_photon_map_host(recs.data(), (int)recs.size(), photon_total, photon_radius, pmap);
        pg = pmap.view(photon_k, photon_radius);
        if (query_in) {
            FILE *fq = fopen(query_in, "rb");
            if (!fq) { fprintf(stderr, "cannot read %s\n", query_in); return 3; }
            std::vector<float> q;
            float rec6[6];
            while (fread(rec6, sizeof(float), 6, fq) == 6) q.insert(q.end(), rec6, rec6 + 6);
            fclose(fq);
            // per query: the estimator's (irradiance, direction), then the same from a plain scan of all
            // gather records (same arithmetic, std::sort instead of the grid + histogram selection)
            std::vector<float> res(q.size() * 2);
            for (size_t i = 0, o = 0; i + 6 <= q.size(); i += 6, o += 12) {
                V3 irr, dir;
                const V3 qp(q[i], q[i + 1], q[i + 2]), qn(q[i + 3], q[i + 4], q[i + 5]);
                estimate_irradiance(pg, qp, qn, irr, dir);
                res[o] = irr.x; res[o + 1] = irr.y; res[o + 2] = irr.z; res[o + 3] = dir.x; res[o + 4] = dir.y; res[o + 5] = dir.z;
                std::vector<std::pair<float, int>> cand;
                for (int j = 0; j < pg.n_photons; j++) {
                    const GatherRec &g = pg.recs[j];
                    V3 dif(g.a.x - qp.x, g.a.y - qp.y, g.a.z - qp.z);
                    float d2 = length2(dif);
                    if (!(d2 < pg.radius * pg.radius)) continue;
                    if (dot(V3(g.b.x, g.b.y, g.b.z), qn) >= 0) continue;
                    cand.push_back(std::make_pair(d2, j));
                }
                std::sort(cand.begin(), cand.end());
                float r2 = pg.radius * pg.radius;
                if ((int)cand.size() > pg.k) { cand.resize(pg.k); r2 = cand.back().first; }
                V3 sp(0, 0, 0), sd(0, 0, 0);
                for (auto &c : cand) { const GatherRec &g = pg.recs[c.second]; sp = sp + pm_power(&g, g.a.w); sd = sd + V3(g.b.x, g.b.y, g.b.z) * g.a.w; }
                if (!cand.empty()) { sp = sp * (1.0f / (3.14159265358979323846f * r2)); sd = normalized(sd); }
                res[o + 6] = sp.x; res[o + 7] = sp.y; res[o + 8] = sp.z; res[o + 9] = sd.x; res[o + 10] = sd.y; res[o + 11] = sd.z;
            }
            FILE *fo = fopen(query_out, "wb");
            if (!fo) return 4;
            fwrite(res.data(), sizeof(float), res.size(), fo);
            fclose(fo);
            printf("{\"emu\": true, \"queries\": %d}\n", (int)(q.size() / 6));
            return 0;
        }
    }
    ShadeCfg cfg; cfg.mode = det ? BHRT_MODE_DET : BHRT_MODE_PATH; cfg.gi_depth = gi; cfg.tir_budget = 16;
    // the depth the library's host side converges to by doubling (render.cu: launch_render) when memory allows;
    // --max-frames N pins it (the estimator then cuts deeper chains off, counted in "depth_overflow")
    cfg.max_frames = max_frames > 0 ? max_frames : BHRT_MAX_FRAMES_LIMIT;
    if (row_e == 0) row_e = H;
    if (row_s < 1) row_s = 1;
    if (smp_e == 0) smp_e = spp;
    std::vector<float> linear((size_t)W * H * 3, 0.f), z((size_t)W * H, 0.f), acc((size_t)W * H * 4, 0.f);
    std::vector<uint8_t> rgb8((size_t)W * H * 3, 0);
    std::vector<int32_t> hitid((size_t)W * H, 0);
    std::vector<uint32_t> pixrays((size_t)W * H * 2, 0);     // per-pixel (closest, shadow) ray counts, written with --pixel-rays
    unsigned long long closest = 0, shadow = 0, overflowed = 0;
    std::vector<std::vector<Frame>> stacks(omp_get_max_threads(), std::vector<Frame>(cfg.max_frames));
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : closest, shadow, overflowed)
    for (int j = row_b; j < row_e; j += row_s) {
        for (int i = 0; i < W; i++) {
            RayCounters rc; rc.closest = 0; rc.shadow = 0;
            V3 sum(0, 0, 0);
            int id0 = -1; float z0 = BHRT_BIGFLOAT;
            Lane L; LaneStack S; S.fr = stacks[omp_get_thread_num()].data();
            for (int sidx = smp_b; sidx < smp_e; sidx++) {
                L.rng.seed(seed, (uint64_t)j * W + i, sidx);
                Ray pr = primary_ray(s, i, j, !det, L.rng);
                int id; float zz;
                V3 c = trace_sample(L, S, s, cfg, pg, pr, (float)i / W, (float)j / H, rc, id, zz);
                if (sidx == 0) { id0 = id; z0 = zz; }
                sum = sum + c;
            }
            V3 out = sum / (float)spp;
            size_t p = (size_t)j * W + i;
            acc[4 * p] = sum.x; acc[4 * p + 1] = sum.y; acc[4 * p + 2] = sum.z; acc[4 * p + 3] = (float)(smp_e - smp_b);
            linear[3 * p] = out.x; linear[3 * p + 1] = out.y; linear[3 * p + 2] = out.z;
            V3 g = gamma_correct(out);
            rgb8[3 * p] = float_to_byte(g.x); rgb8[3 * p + 1] = float_to_byte(g.y); rgb8[3 * p + 2] = float_to_byte(g.z);
            // like render.cu: the AOVs of sample 0 are written by the share that owns sample 0, zeros elsewhere
            if (smp_b == 0) { hitid[p] = id0; z[p] = z0; }
            closest += rc.closest; shadow += rc.shadow; overflowed += rc.overflow;
            pixrays[2 * p] = (uint32_t)rc.closest; pixrays[2 * p + 1] = (uint32_t)rc.shadow;
        }
    }
    double ms = (omp_get_wtime() - t0) * 1e3;
#ifdef BHRT_EMU_STATS
    for (int k = 0; k < 2; k++) {
        fprintf(stderr, "%s queries %llu  internal steps %.2f  leaf steps %.2f per query\n  hist(steps < 2^b):", k ? "shadow" : "closest", g_nq[k],
                (double)g_sum[k][0] / g_nq[k], (double)g_sum[k][1] / g_nq[k]);
        for (int b = 0; b < 24; b++) fprintf(stderr, " %llu", g_hist[k][b]);
        fprintf(stderr, "\n");
    }
#endif
    std::string path = std::string(argv[2]) + ".aov";
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) return 4;
    int32_t hdr[4] = { 0x54524842, W, H, spp };
    fwrite(hdr, sizeof(hdr), 1, fp);
    fwrite(linear.data(), 4, linear.size(), fp);
    fwrite(rgb8.data(), 1, rgb8.size(), fp);
    fwrite(hitid.data(), 4, hitid.size(), fp);
    fwrite(z.data(), 4, z.size(), fp);
    fclose(fp);
    if (pixel_rays) {
        fp = fopen(pixel_rays, "wb");
        if (!fp) return 4;
        fwrite(pixrays.data(), 4, pixrays.size(), fp);
        fclose(fp);
    }
    fp = fopen((std::string(argv[2]) + ".acc").c_str(), "wb");
    if (!fp) return 4;
    fwrite(acc.data(), 4, acc.size(), fp);
    fclose(fp);
    printf("{\"emu\": true, \"width\": %d, \"height\": %d, \"spp\": %d, \"ms\": %.3f, \"rays_closest\": %llu, \"rays_shadow\": %llu, \"depth_overflow\": %llu, \"nodes\": %d}\n",
           W, H, spp, ms, closest, shadow, overflowed, D.n_nodes);
    return 0;
}
