// TEST INFRASTRUCTURE (oracle/_ref build only). Not part of the product library.
//
// Headless driver for the UNMODIFIED reference renderer (sources compiled where they lie
// under /root/reference, plus the enumerated sed patches in build_ref.sh).  It replaces the
// reference's main() (Main.cpp:418-431), whose scene path / thread count are literals, and
// provides the hooks the patches call: AOV recording, ray counters, per-thread rand().
//
// usage: bhrt_ref <scene.xml> <out_prefix> [--det] [--spp N] [--gi N] [--threads N]
//                 [--seed S] [--photons N] [--png] [--libc-rand]
// writes <out_prefix>.aov  (binary, see below) and prints one JSON line on stdout.
//
// .aov layout (little endian):
//   int32 magic 'BHRT' (0x54524842), int32 W, int32 H, int32 spp,
//   float  linear[H*W*3]   pre-gamma pixel colour (the value PathTracing(i,j) returned)
//   uint8  rgb8  [H*W*3]   renderImage.GetPixels() (post gamma, Color24)
//   int32  hitid [H*W]     DFS pre-order index of the primary hit node of sample 0, -1 = miss
//   float  z     [H*W]     primary hit z of sample 0 (BIGFLOAT on miss)
#include "Scenes/scene.h"
#include "Objects/objects.h"
#include "Lights/lights.h"
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <string>
#include <map>
#include <vector>

extern Node rootNode;
extern Camera camera;
extern RenderImage renderImage;
int LoadScene(char const *filename);
void BeginRender();

// knobs the patched reference reads (see build_ref.sh)
int g_bhrt_det = 0;
int g_bhrt_face_mtl = 0;      // patch Q-1 (oracle/build_ref.py): per-face sub-materials
int g_bhrt_spp = 32;      // reference Main.cpp:141
int g_bhrt_gi = 3;        // reference Main.cpp:130
int g_bhrt_photons = 1000000;  // reference Main.cpp:52-53
int g_bhrt_global_pm = 0;      // --global-photons: BuildPhotonMap instead of BuildCausticPhotonMap, then exit (patch O-7)
const char *g_bhrt_photon_path = nullptr;

static std::vector<float> s_linear;
static std::vector<int32_t> s_hitid;
static std::vector<float> s_z;
static std::map<const Node *, int> s_nodeId;
static int s_W = 0, s_H = 0;

struct alignas(64) Counters { long long closest, shadow; long long pad[6]; };
static Counters s_cnt[512];
static double s_photon_t0 = 0, s_photon_ms = 0;

void bhrt_record_primary(int i, int j, const Node *node, float z)
{
	int idx = j * s_W + i;
	int id = -1;
	if (node) { auto it = s_nodeId.find(node); if (it != s_nodeId.end()) id = it->second; }
	s_hitid[idx] = id;
	s_z[idx] = z;
}
void bhrt_record_linear(int i, int j, float r, float g, float b)
{
	int idx = (j * s_W + i) * 3;
	s_linear[idx] = r; s_linear[idx + 1] = g; s_linear[idx + 2] = b;
}
#ifdef BHRT_COUNT_ALL
// work counters of the instrumented build (see build_ref.py): 0 instances tested, 1 BVH boxes
// tested, 2 triangles tested, 3 triangles accepted (vn+vt fetched), 4 file-texture taps, 5 photons examined by the gather
// (LocatePhotons calls, cyPhotonMap.h:421; photon-map build only), 6 gather queries (EstimateIrradiance calls)
struct alignas(64) Work { long long c[8]; };
static Work s_work[512];
void bhrt_count(int k) { s_work[omp_get_thread_num() & 511].c[k]++; }
#endif
void bhrt_count_closest() { s_cnt[omp_get_thread_num() & 511].closest++; }
void bhrt_count_shadow() { s_cnt[omp_get_thread_num() & 511].shadow++; }
void bhrt_photon_begin() { s_photon_t0 = omp_get_wtime(); }
static FILE *s_real_out = nullptr;
void bhrt_global_done() { FILE *o = s_real_out ? s_real_out : stdout; fprintf(o, "{\"global_photon_map\": true, \"photons\": %d, \"photon_build_ms\": %.3f}\n", g_bhrt_photons, (omp_get_wtime() - s_photon_t0) * 1e3); fflush(o); _exit(0); }
void bhrt_photon_end() { s_photon_ms = (omp_get_wtime() - s_photon_t0) * 1e3; }

// ---- per-thread rand(): glibc's rand() takes a global lock (8 threads run 4x slower than 1,
// SURVEY.md section 6).  MSVC's rand() is per-thread; this restores that behaviour.  The
// executable's definition pre-empts libc's.  Disabled with --libc-rand.
static int s_use_libc_rand = 0;
static uint64_t s_seed = 0x9E3779B97F4A7C15ull;
static thread_local uint64_t t_state = 0;
extern "C" int __bhrt_libc_rand(void);
static inline uint64_t splitmix(uint64_t &x)
{
	uint64_t z = (x += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
extern "C" int rand(void) noexcept
{
	if (s_use_libc_rand) return (int)(random() & 0x7fffffff);  // random() == glibc rand(), same lock
	if (t_state == 0) {
		uint64_t s = s_seed + 0x51ED270B7F4A7C15ull * (uint64_t)(omp_get_thread_num() + 1);
		t_state = splitmix(s) | 1;
	}
	// xorshift64*
	uint64_t x = t_state;
	x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
	t_state = x;
	return (int)(((x * 0x2545F4914F6CDD1Dull) >> 33) & 0x7fffffff);  // [0, RAND_MAX]
}

static void number_nodes(const Node *n, int &next)
{
	for (int i = 0; i < n->GetNumChild(); i++) {
		const Node *c = n->GetChild(i);
		s_nodeId[c] = next++;
		number_nodes(c, next);
	}
}

int main(int argc, char **argv)
{
	if (argc < 3) { fprintf(stderr, "usage: %s scene.xml out_prefix [--det] [--spp N] [--gi N] [--threads N] [--seed S] [--photons N] [--png] [--libc-rand]\n", argv[0]); return 2; }
	const char *scene = argv[1];
	std::string prefix = argv[2];
	int threads = omp_get_max_threads();
	int writePng = 0;
	for (int a = 3; a < argc; a++) {
		if (!strcmp(argv[a], "--det")) g_bhrt_det = 1;
		else if (!strcmp(argv[a], "--spp") && a + 1 < argc) g_bhrt_spp = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--gi") && a + 1 < argc) g_bhrt_gi = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--threads") && a + 1 < argc) threads = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--seed") && a + 1 < argc) s_seed = strtoull(argv[++a], nullptr, 0) * 0x9E3779B97F4A7C15ull + 12345;
		else if (!strcmp(argv[a], "--photons") && a + 1 < argc) g_bhrt_photons = atoi(argv[++a]);
		else if (!strcmp(argv[a], "--png")) writePng = 1;
		else if (!strcmp(argv[a], "--libc-rand")) s_use_libc_rand = 1;
		else if (!strcmp(argv[a], "--global-photons")) g_bhrt_global_pm = 1;
		else if (!strcmp(argv[a], "--face-materials")) g_bhrt_face_mtl = 1;
		else { fprintf(stderr, "unknown arg %s\n", argv[a]); return 2; }
	}
	if (g_bhrt_det) g_bhrt_spp = 1;
	std::string photonPath = prefix + ".photons.dat";
	g_bhrt_photon_path = photonPath.c_str();
	omp_set_num_threads(threads);

	// the loader printf()s the whole scene; keep stdout for the JSON line only
	FILE *realOut = fdopen(dup(fileno(stdout)), "w");
	s_real_out = realOut;
	if (!freopen("/dev/null", "w", stdout)) return 3;
	if (!LoadScene(scene)) { fprintf(stderr, "LoadScene failed: %s\n", scene); return 3; }
	s_W = camera.imgWidth; s_H = camera.imgHeight;
	s_linear.assign((size_t)s_W * s_H * 3, 0.f);
	s_hitid.assign((size_t)s_W * s_H, -1);
	s_z.assign((size_t)s_W * s_H, BIGFLOAT);
	int next = 0;
	number_nodes(&rootNode, next);

	double t0 = omp_get_wtime();
	BeginRender();
	double ms = (omp_get_wtime() - t0) * 1e3;

	long long closest = 0, shadow = 0;
	for (int i = 0; i < 512; i++) { closest += s_cnt[i].closest; shadow += s_cnt[i].shadow; }

	std::string aov = prefix + ".aov";
	FILE *fp = fopen(aov.c_str(), "wb");
	if (!fp) { fprintf(stderr, "cannot write %s\n", aov.c_str()); return 4; }
	int32_t hdr[4] = { 0x54524842, s_W, s_H, g_bhrt_spp };
	fwrite(hdr, sizeof(hdr), 1, fp);
	fwrite(s_linear.data(), sizeof(float), s_linear.size(), fp);
	fwrite(renderImage.GetPixels(), 3, (size_t)s_W * s_H, fp);
	fwrite(s_hitid.data(), sizeof(int32_t), s_hitid.size(), fp);
	fwrite(s_z.data(), sizeof(float), s_z.size(), fp);
	fclose(fp);
	if (writePng) renderImage.SaveImage((prefix + ".png").c_str());

	fprintf(realOut, "{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"gi\": %d, \"det\": %d, "
		"\"threads\": %d, \"rand\": \"%s\", \"ms\": %.3f, \"photon_build_ms\": %.3f, \"rays_closest\": %lld, \"rays_shadow\": %lld, \"nodes\": %d",
		scene, s_W, s_H, g_bhrt_spp, g_bhrt_gi, g_bhrt_det, threads, s_use_libc_rand ? "libc" : "per-thread",
		ms - s_photon_ms, s_photon_ms, closest, shadow, next);
#ifdef BHRT_COUNT_ALL
	long long w[7] = { 0, 0, 0, 0, 0, 0, 0 };
	for (int i = 0; i < 512; i++) for (int k = 0; k < 7; k++) w[k] += s_work[i].c[k];
	fprintf(realOut, ", \"inst_tested\": %lld, \"boxes_tested\": %lld, \"tris_tested\": %lld, \"tris_accepted\": %lld, \"texture_taps\": %lld, \"photons_examined\": %lld, \"gather_queries\": %lld",
		w[0], w[1], w[2], w[3], w[4], w[5], w[6]);
#endif
	fprintf(realOut, "}\n");
	fclose(realOut);
	return 0;
}
