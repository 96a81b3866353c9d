// bhrt_render -- headless C++ host program: the replacement for the reference's main()
// (reference Main.cpp:418-431: LoadScene(literal path) -> ShowViewport() -> BeginRender()).
// It stays C++ and reaches the GPU only through the C ABI of include/bhrt.h.
//
// usage: bhrt_render <scene.xml> <out_prefix> [--det] [--spp N] [--gi N] [--seed S] [--size W H]
//                    [--device D] [--gpus N] [--aov] [--photons N] [--photons-out F.dat] [--photons-in F.dat]
// --gpus N renders ONE frame on N GPUs of the box from this one process: a host thread per GPU, each with its own
// bhrt_ctx, joined by the C++ multi-GPU host of libbhrt.so (include/bhrt_dist.h: interleaved rows or sample ranges,
// NCCL all-reduce, sharded photon emission + all-gather) -- the host stays C++ at any GPU count.
// --photons-out writes the caustic map like the reference does (Main.cpp:380-385: 24-byte records,
// powers already scaled by 1/stored; PhotonMapViewer/PhotonMapViz.cpp reads this file); --photons-in
// renders with such a file (e.g. one the reference wrote) instead of emitting photons.
// writes <out_prefix>.png (RenderImage::SaveImage equivalent) and, with --aov, <out_prefix>.aov in
// the layout of oracle/ref_build/driver.cpp; prints one JSON line.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../../include/bhrt.h"
#include "../../../include/bhrt_host.h"
#include "../../../include/bhrt_dist.h"
#include <thread>

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.xml out_prefix [--det] [--spp N] [--gi N] [--seed S] [--size W H] [--device D] [--aov] [--photons N]\n", argv[0]); return 2; }
    bhrt_params p;
    memset(&p, 0, sizeof(p));
    p.mode = BHRT_MODE_PATH; p.spp = 32; p.gi_depth = 3; p.tir_budget = 16; p.seed = 1;
    int W = 0, H = 0, device = 0, aov = 0, photons = 0, gpus = 1;
    const char *photons_out = nullptr, *photons_in = nullptr;
    for (int a = 3; a < argc; a++) {
        if (!strcmp(argv[a], "--det")) p.mode = BHRT_MODE_DET;
        else if (!strcmp(argv[a], "--spp") && a + 1 < argc) p.spp = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--gi") && a + 1 < argc) p.gi_depth = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--seed") && a + 1 < argc) p.seed = strtoull(argv[++a], nullptr, 0);
        else if (!strcmp(argv[a], "--size") && a + 2 < argc) { W = atoi(argv[++a]); H = atoi(argv[++a]); }
        else if (!strcmp(argv[a], "--device") && a + 1 < argc) device = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--gpus") && a + 1 < argc) gpus = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--photons") && a + 1 < argc) photons = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--photons-out") && a + 1 < argc) photons_out = argv[++a];
        else if (!strcmp(argv[a], "--photons-in") && a + 1 < argc) photons_in = argv[++a];
        else if (!strcmp(argv[a], "--aov")) aov = 1;
        else { fprintf(stderr, "unknown arg %s\n", argv[a]); return 2; }
    }
    if (p.mode == BHRT_MODE_DET) p.spp = 1;
    bhrt_host_scene *scene = bhrt_host_load_scene(argv[1]);
    if (!scene) { fprintf(stderr, "LoadScene failed: %s\n", bhrt_host_last_error()); return 3; }
    if (*bhrt_host_warnings(scene)) fprintf(stderr, "%s", bhrt_host_warnings(scene));
    if (W > 0 && bhrt_host_set_resolution(scene, W, H)) { fprintf(stderr, "bad --size\n"); return 2; }
    const bhrt_scene_desc *d = bhrt_host_scene_desc(scene);
    W = d->camera.width; H = d->camera.height;

    if (gpus > 1) {
        // ---- one frame on `gpus` GPUs: a thread per GPU around the C++ multi-GPU host
        if (photons_in || photons_out) { fprintf(stderr, "--photons-in / --photons-out are single-GPU options\n"); return 2; }
        uint8_t id[BHRT_DIST_ID_BYTES];
        if (bhrt_dist_unique_id(id)) { fprintf(stderr, "bhrt_dist_unique_id: %s\n", bhrt_last_error()); return 4; }
        const size_t n = (size_t)W * H;
        std::vector<uint8_t> rgb8(n * 3);
        std::vector<float> linear(n * 3), z(n);
        std::vector<int32_t> hit(n);
        bhrt_outputs o;
        memset(&o, 0, sizeof(o));
        o.rgb8 = rgb8.data(); o.linear = linear.data(); o.hit_id = hit.data(); o.z = z.data();
        std::vector<int> status(gpus, 0);
        std::vector<std::string> errors(gpus);
        int stored_all = 0;
        std::vector<std::thread> th;
        for (int k = 0; k < gpus; k++)
            th.emplace_back([&, k]() {
                bhrt_ctx *c = nullptr; bhrt_dist *dd = nullptr;
                bhrt_params pk = p;
                auto bail = [&](const char *what) { status[k] = 1; errors[k] = std::string(what) + ": " + bhrt_last_error(); };
                if (bhrt_create(device + k, &c)) return bail("bhrt_create");
                if (bhrt_upload_scene(c, d)) return bail("bhrt_upload_scene");
                if (bhrt_dist_create(c, k, gpus, id, &dd)) return bail("bhrt_dist_create");
                if (photons > 0) {
                    int32_t st = 0; uint64_t pr = 0;
                    if (bhrt_dist_photons(dd, photons, pk.seed, &st, &pr)) return bail("bhrt_dist_photons");
                    if (k == 0) stored_all = st;
                    pk.use_photon_map = 1; pk.photon_k = 1000; pk.photon_radius = 0.5f;
                }
                bhrt_outputs ok_;
                memset(&ok_, 0, sizeof(ok_));
                if (bhrt_dist_render(dd, &pk, 0, k == 0 ? &o : &ok_)) return bail("bhrt_dist_render");
                bhrt_dist_destroy(dd);
                bhrt_destroy(c);
            });
        for (auto &t : th) t.join();
        for (int k = 0; k < gpus; k++) if (status[k]) { fprintf(stderr, "GPU %d: %s\n", device + k, errors[k].c_str()); return 5; }
        std::string prefix = argv[2];
        if (bhrt_host_save_png((prefix + ".png").c_str(), rgb8.data(), W, H, 3)) fprintf(stderr, "cannot write %s.png\n", prefix.c_str());
        if (aov) {
            FILE *fp = fopen((prefix + ".aov").c_str(), "wb");
            if (fp) {
                int32_t hdr[4] = { 0x54524842, W, H, p.spp };
                fwrite(hdr, sizeof(hdr), 1, fp);
                fwrite(linear.data(), 4, linear.size(), fp); fwrite(rgb8.data(), 1, rgb8.size(), fp);
                fwrite(hit.data(), 4, hit.size(), fp); fwrite(z.data(), 4, z.size(), fp);
                fclose(fp);
            }
        }
        const double rays = (double)o.rays_closest + (double)o.rays_shadow;
        printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"mode\": \"%s\", \"gpus\": %d, \"kernel_ms\": %.4f, "
               "\"rays_closest\": %llu, \"rays_shadow\": %llu, \"mrays_per_s\": %.2f, \"photons_stored\": %d}\n",
               argv[1], W, H, p.spp, p.mode == BHRT_MODE_DET ? "det" : "path", gpus, o.kernel_ms,
               (unsigned long long)o.rays_closest, (unsigned long long)o.rays_shadow, rays / (o.kernel_ms * 1e3), stored_all);
        bhrt_host_free_scene(scene);
        return 0;
    }
    bhrt_ctx *ctx = nullptr;
    if (bhrt_create(device, &ctx)) { fprintf(stderr, "bhrt_create: %s\n", bhrt_last_error()); return 4; }
    if (bhrt_upload_scene(ctx, d)) { fprintf(stderr, "bhrt_upload_scene: %s\n", bhrt_last_error()); return 4; }
    int stored = 0;
    uint64_t photon_rays = 0;
    if (photons > 0) {
        if (bhrt_photons_emit(ctx, photons, p.seed, &stored, &photon_rays) || bhrt_photons_build(ctx, stored)) {
            fprintf(stderr, "photon map: %s\n", bhrt_last_error()); return 4;
        }
        p.use_photon_map = 1; p.photon_k = 1000; p.photon_radius = 0.5f;
    }
    if (photons_in) {
        FILE *fp = fopen(photons_in, "rb");
        if (!fp) { fprintf(stderr, "cannot read %s\n", photons_in); return 3; }
        std::vector<uint8_t> rec;
        uint8_t buf[24 * 4096];
        size_t got;
        while ((got = fread(buf, 1, sizeof(buf), fp)) > 0) rec.insert(rec.end(), buf, buf + got);
        fclose(fp);
        stored = (int)(rec.size() / 24);
        // the file holds scaled powers already: total_stored = 1 leaves them as they are
        if (bhrt_photons_set(ctx, rec.data(), stored, 0) || bhrt_photons_build(ctx, 1)) { fprintf(stderr, "photon map: %s\n", bhrt_last_error()); return 4; }
        p.use_photon_map = 1; p.photon_k = 1000; p.photon_radius = 0.5f;
    }
    if (photons_out && stored > 0 && !photons_in) {
        std::vector<uint8_t> rec((size_t)stored * 24);
        int32_t got = 0;
        if (bhrt_photons_get(ctx, rec.data(), stored, &got)) { fprintf(stderr, "photon map: %s\n", bhrt_last_error()); return 4; }
        for (int i = 0; i < got; i++) {                 // ScalePhotonPowers(1 / stored) (Main.cpp:380, cyPhotonMap.h:149-152)
            float pw; memcpy(&pw, &rec[(size_t)i * 24 + 12], 4);
            pw *= 1.0f / (float)got;
            memcpy(&rec[(size_t)i * 24 + 12], &pw, 4);
        }
        FILE *fp = fopen(photons_out, "wb");
        if (!fp || fwrite(rec.data(), 24, (size_t)got, fp) != (size_t)got) { fprintf(stderr, "cannot write %s\n", photons_out); return 3; }
        fclose(fp);
    }
    size_t n = (size_t)W * H;
    std::vector<uint8_t> rgb8(n * 3);
    std::vector<float> linear(n * 3), z(n);
    std::vector<int32_t> hit(n);
    bhrt_outputs o;
    memset(&o, 0, sizeof(o));
    o.rgb8 = rgb8.data(); o.linear = linear.data(); o.hit_id = hit.data(); o.z = z.data();
    if (bhrt_render(ctx, &p, &o)) { fprintf(stderr, "bhrt_render: %s\n", bhrt_last_error()); return 5; }
    std::string prefix = argv[2];
    if (bhrt_host_save_png((prefix + ".png").c_str(), rgb8.data(), W, H, 3)) fprintf(stderr, "cannot write %s.png\n", prefix.c_str());
    if (aov) {
        FILE *fp = fopen((prefix + ".aov").c_str(), "wb");
        if (fp) {
            int32_t hdr[4] = { 0x54524842, W, H, p.spp };
            fwrite(hdr, sizeof(hdr), 1, fp);
            fwrite(linear.data(), 4, linear.size(), fp);
            fwrite(rgb8.data(), 1, rgb8.size(), fp);
            fwrite(hit.data(), 4, hit.size(), fp);
            fwrite(z.data(), 4, z.size(), fp);
            fclose(fp);
        }
    }
    double rays = (double)o.rays_closest + (double)o.rays_shadow;
    printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"mode\": \"%s\", \"kernel_ms\": %.4f, "
           "\"rays_closest\": %llu, \"rays_shadow\": %llu, \"mrays_per_s\": %.2f, \"photons_stored\": %d}\n",
           argv[1], W, H, p.spp, p.mode == BHRT_MODE_DET ? "det" : "path", o.kernel_ms,
           (unsigned long long)o.rays_closest, (unsigned long long)o.rays_shadow, rays / (o.kernel_ms * 1e3), stored);
    bhrt_destroy(ctx);
    bhrt_host_free_scene(scene);
    return 0;
}
