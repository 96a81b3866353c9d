#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: build the reference renderer headless into oracle/_ref/.

Compiles the reference's own sources WHERE THEY LIE under /root/reference (nothing is
copied into this repository).  Three of its files carry compile-time literals that the
BASELINE configs need to vary (spp, GI depth, photon counts, scene/output path) and have no
deterministic mode, so those three are piped through the guarded line patches below into a
throw-away temp directory and compiled from there; only the resulting binaries land in
oracle/_ref/ (git-ignored, NOT gpurun-ignored, so they travel to the GPU box).

Patch list (numbers = SURVEY.md section 8(c)):
  O-1  main() renamed away (-Dmain=...), driver.cpp supplies scene path / threads / output
  O-2  PT_SampleCount, GIBounceCount, MAX_*PhotonCount -> runtime ints
  O-3  `return true;` at the end of both photon builders (UB otherwise)
  O-4  per-thread rand() (driver.cpp, link-time; no source edit)
  O-5  AOV + ray-counter hooks
  O-7  --global-photons: call BuildPhotonMap (Main.cpp:251-295; its call at Main.cpp:196 is commented out at HEAD), dump, exit
  Q-1  --face-materials: TriObj::IntersectTriangle sets hInfo.mtlID = the usemtl group of the face, which HEAD never does
       (TriObj.cpp:68-189; MultiMtl::Shade dispatches on it, materials.h:71) -- the reference behaviour behind the product's
       optional BHRT_FLAG_FACE_MATERIALS; off unless asked for, so every other fixture is the unmodified behaviour
  D-1  det mode: no jitter              (Main.cpp:153)
  D-2  det mode: GI term off            (MtlBlinn.cpp:124)
  D-3  det mode: all lights, exact L    (MtlBlinn.cpp:308)
  D-4  det mode: refraction gloss = 0   (MtlBlinn.cpp:116)
All D-* patches are runtime-switched by g_bhrt_det, so one binary serves both modes.
`bhrt_ref_pm` is the same build with -DUSE_PhotonMap (reference Main.cpp:51, MtlBlinn.cpp:25).
`bhrt_ref_gpu` links the reference's loader and scene classes with ref_build/GpuRender.cpp: the reference-side
binding of INTEGRATION.md 1 (flattens the reference's loaded globals into include/bhrt.h tables; renders through libbhrt.so).
`bhrt_ref_cnt` is bhrt_ref with -DBHRT_COUNT_ALL: five more per-thread counters (instances
tested, BVH boxes tested, triangles tested, triangles accepted, file-texture taps) that define
the ALGORITHMIC bytes per ray of SURVEY.md 8(d); it is never used for timing.
"""
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BHRT_REFERENCE", "/root/reference") + "/BHRayTracer"
OUT = os.path.join(HERE, "_ref")

HOOKS = (
    "extern int g_bhrt_det, g_bhrt_spp, g_bhrt_gi, g_bhrt_photons, g_bhrt_global_pm; extern const char* g_bhrt_photon_path; void bhrt_global_done();\n"
    "class Node; void bhrt_record_primary(int,int,const Node*,float); void bhrt_record_linear(int,int,float,float,float);\n"
    "void bhrt_count_closest(); void bhrt_count_shadow(); void bhrt_photon_begin(); void bhrt_photon_end();\n"
)
# work counters of the -DBHRT_COUNT_ALL build (0 inst, 1 box, 2 tri, 3 tri accepted, 4 texture tap)
CNT = "#ifdef BHRT_COUNT_ALL\nvoid bhrt_count(int);\n#define BHRT_CNT(k) bhrt_count(k)\n#else\n#define BHRT_CNT(k)\n#endif\n"

DET_DIRECT = (
    "if (g_bhrt_det) { for (int li = 0; li < (int)lights.size(); ++li) { Light* light = lights[li];"
    " Vec3f vL = -light->Direction(hInfo.p).GetNormalized(); float cosTheta = vL.Dot(vN);"
    " if (cosTheta > 0) { Vec3f vH = (vL + vV).GetNormalized(); Color irrad = (light)->Illuminate(hInfo.p, vN);"
    " Color brdfXCosTheta = diffuse.Sample(hInfo.uvw, hInfo.duvw)  * cosTheta + specular.Sample(hInfo.uvw, hInfo.duvw) * pow(vH.Dot(vN), glossiness);"
    " outColor += irrad * brdfXCosTheta; } } } else\n"
)

# (line number, substring that must be on that line, replacement function)
PATCHES = {
    "Main.cpp": [
        (1, "", lambda s: HOOKS + CNT + s),
        (393, "Ray transformedRay = root->GetChild(i)->ToNodeCoords(ray);", lambda s: s + "\t\tBHRT_CNT(0);\n"),
        (52, "#define MAX_PhotonCount 1000000", lambda s: "#define MAX_PhotonCount g_bhrt_photons\n"),
        (53, "#define MAX_CausticPhotonCount 1000000", lambda s: "#define MAX_CausticPhotonCount g_bhrt_photons\n"),
        (130, "#define GIBounceCount 3", lambda s: "#define GIBounceCount g_bhrt_gi\n"),
        (141, "#define PT_SampleCount 32", lambda s: "#define PT_SampleCount g_bhrt_spp\n"),
        (153, "RandomPositionInPixel(pixelCenter, pixelLen) - camera.pos",
         lambda s: s.replace("RandomPositionInPixel(pixelCenter, pixelLen)",
                             "(g_bhrt_det ? pixelCenter : RandomPositionInPixel(pixelCenter, pixelLen))")),
        (158, "recursive(&rootNode, ray, tHitInfo, bHit, HIT_FRONT);",
         lambda s: s + "\t\tif (i == 0) bhrt_record_primary(i_i, i_j, bHit ? tHitInfo.node : nullptr, tHitInfo.z);\n"),
        (197, "BuildCausticPhotonMap(lights, &rootNode);",
         lambda s: "\tbhrt_photon_begin(); if (g_bhrt_global_pm) { BuildPhotonMap(lights, &rootNode); bhrt_global_done(); } BuildCausticPhotonMap(lights, &rootNode); bhrt_photon_end();\n"),
        (218, "outColor = PathTracing(i, j);",
         lambda s: s + "\t\t\tbhrt_record_linear(i, j, outColor.r, outColor.g, outColor.b);\n"),
        (240, "SaveImages();", lambda s: "\t;\n"),
        (292, 'fopen("Resource/photonmap.dat", "wb")', lambda s: s.replace('"Resource/photonmap.dat"', "g_bhrt_photon_path")),
        (294, "fclose(fp);", lambda s: s.rstrip("\n") + " return true;\n"),
        (383, 'fopen("Resource/causticPhotonMap.dat", "wb")', lambda s: s.replace('"Resource/causticPhotonMap.dat"', "g_bhrt_photon_path")),
        (385, "fclose(fp);", lambda s: s.rstrip("\n") + " return true;\n"),
        (390, "if (root->GetNumChild() <= 0) return;", lambda s: "\tif (root == &rootNode) bhrt_count_closest();\n" + s),
    ],
    "DataStructure/cyPhotonMap.h": [
        (423, "Photon const &p = photons[index];", lambda s: s + "\t\tBHRT_CNT(5);\n"),
    ],
    "Materials/Blinn/MtlBlinn.cpp": [
        (1, "", lambda s: "extern int g_bhrt_det;\n" + CNT + s),
        (116, "if (glossiness > 50) refraction_glossiness = glossiness;",
         lambda s: s.replace("if (glossiness > 50)", "if (!g_bhrt_det && glossiness > 50)")),
        (124, "outColor += PathTracing_GlobalIllumination(", lambda s: "\tif (!g_bhrt_det) " + s.lstrip()),
        (308, "{", lambda s: DET_DIRECT + s),
        (334, "causticPhotonMap->EstimateIrradiance<MAX_PhotonCountInArea>(", lambda s: "\t\t\tBHRT_CNT(6);\n" + s),
    ],
    "Lights/GenLight.cpp": [
        (1, "", lambda s: "void bhrt_count_shadow();\n" + CNT + s),
        (17, "Ray transformedRay = root->ToNodeCoords(ray);", lambda s: s + "\tif (root != &rootNode) BHRT_CNT(0);\n"),
        (12, "return ShadowRayRecursive(&rootNode, ray, t_max) ? 0.f : 1.f;", lambda s: "\tbhrt_count_shadow();\n" + s),
    ],
    "Objects/Box/Box.cpp": [
        (1, "", lambda s: CNT + s),
        (4, "{", lambda s: "{ BHRT_CNT(1);\n"),
    ],
    "Objects/TriObj/TriObj.cpp": [
        (1, "", lambda s: "extern int g_bhrt_face_mtl;\n" + CNT + s),
        (69, "{", lambda s: "{ BHRT_CNT(2);\n"),
        (169, "Vec3f hitPointNormal = GetNormal(faceID, bc);",
         lambda s: "\tBHRT_CNT(3);\n\tif (g_bhrt_face_mtl) { int m_ = GetMaterialIndex((int)faceID); hInfo.mtlID = m_ > 0 ? m_ : 0; }\n" + s),
    ],
    "Textures/Texture.cpp": [
        (1, "", lambda s: CNT + s),
        (99, "if (width + height == 0) return Color(0, 0, 0);", lambda s: "\tBHRT_CNT(4);\n" + s),
    ],
}

UNPATCHED = [
    "xmlload.cpp", "Lights/PointLight.cpp", "Objects/Plane/Plane.cpp", "Objects/Sphere/Sphere.cpp",
    "Tinyxml2/tinyxml2.cpp", "LoadPNG/lodepng.cpp",
]


def patched_copy(rel, tmp):
    with open(os.path.join(REF, rel), "r", newline="") as f:
        lines = f.readlines()
    for lineno, expect, fn in PATCHES[rel]:
        cur = lines[lineno - 1]
        if expect not in cur:
            raise SystemExit(f"{rel}:{lineno}: expected {expect!r}, found {cur!r} -- reference changed?")
        lines[lineno - 1] = fn(cur)
    dst = os.path.join(tmp, rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w", newline="") as f:
        f.writelines(lines)
    return dst


def main():
    if not os.path.isdir(REF):
        print(f"build_ref: {REF} not present (GPU box?) -- keeping prebuilt oracle/_ref", file=sys.stderr)
        return 0
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="bhrt_ref_build_")
    try:
        flags = ["-O2", "-fopenmp", "-w", "-D_stricmp=strcasecmp", "-include", "strings.h",
                 "-I" + REF, "-I" + REF + "/DataStructure", "-I" + REF + "/LoadPNG"]
        jobs = []  # (source, object, extra flags)
        for rel in UNPATCHED:
            jobs.append((os.path.join(REF, rel), os.path.join(tmp, rel.replace("/", "_") + ".o"), []))
        jobs.append((os.path.join(HERE, "ref_build/gl_stubs.cpp"), os.path.join(tmp, "gl_stubs.o"), []))
        main_src = patched_copy("Main.cpp", tmp)
        mtl_src = patched_copy("Materials/Blinn/MtlBlinn.cpp", tmp)
        # cyPhotonMap.h with the photons-examined counter: a quoted include looks in the including file's directory first,
        # so the patched copy sits next to each patched source that includes it
        hdr = patched_copy("DataStructure/cyPhotonMap.h", tmp)
        shutil.copy(hdr, os.path.join(tmp, "cyPhotonMap.h"))
        shutil.copy(hdr, os.path.join(tmp, "Materials", "Blinn", "cyPhotonMap.h"))
        # sources whose object differs between the plain and the -DBHRT_COUNT_ALL build
        counted = [(patched_copy("Lights/GenLight.cpp", tmp), "GenLight", ["-I" + REF + "/Lights"]),
                   (patched_copy("Objects/Box/Box.cpp", tmp), "Box", ["-I" + REF + "/Objects/Box"]),
                   (patched_copy("Objects/TriObj/TriObj.cpp", tmp), "TriObj", ["-I" + REF + "/Objects/TriObj"]),
                   (patched_copy("Textures/Texture.cpp", tmp), "Texture", ["-I" + REF + "/Textures"]),
                   (main_src, "Main", ["-Dmain=ref_main_unused"]),
                   (os.path.join(HERE, "ref_build/driver.cpp"), "driver", [])]
        for src, stem, extra in counted:
            jobs.append((src, os.path.join(tmp, "V_%s.o" % stem), extra))
            jobs.append((src, os.path.join(tmp, "V_%s_cnt.o" % stem), extra + ["-DBHRT_COUNT_ALL"]))
        jobs.append((main_src, os.path.join(tmp, "V_Main_pm.o"), ["-Dmain=ref_main_unused", "-DUSE_PhotonMap"]))
        jobs.append((main_src, os.path.join(tmp, "V_Main_pm_cnt.o"), ["-Dmain=ref_main_unused", "-DUSE_PhotonMap", "-DBHRT_COUNT_ALL"]))
        jobs.append((mtl_src, os.path.join(tmp, "V_MtlBlinn_pm_cnt.o"), ["-I" + REF + "/Materials/Blinn", "-DUSE_PhotonMap", "-DBHRT_COUNT_ALL"]))
        # the reference-side binding of INTEGRATION.md 1 (GpuRender.cpp): the reference's loader + this repository's C ABI
        ROOT_INC = os.path.join(os.path.dirname(HERE), "include")
        jobs.append((os.path.join(HERE, "ref_build/GpuRender.cpp"), os.path.join(tmp, "V_GpuRender.o"),
                     ["-I" + ROOT_INC, "-I" + os.path.join(HERE, "ref_build"), "-std=c++17"]))
        jobs.append((os.path.join(HERE, "ref_build/driver.cpp"), os.path.join(tmp, "V_driver_nomain.o"), ["-Dmain=bhrt_ref_driver_main_unused"]))
        for tag, extra in (("", []), ("_pm", ["-DUSE_PhotonMap"])):
            jobs.append((mtl_src, os.path.join(tmp, "V_MtlBlinn%s.o" % tag), ["-I" + REF + "/Materials/Blinn"] + extra))

        def cc(job):
            src, obj, extra = job
            subprocess.check_call(["g++"] + flags + extra + ["-c", src, "-o", obj])
            return obj

        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            objs = list(ex.map(cc, jobs))
        common = [o for o in objs if "/V_" not in o]
        V = lambda stem: os.path.join(tmp, "V_%s.o" % stem)
        plain = [V("GenLight"), V("Box"), V("TriObj"), V("Texture"), V("driver")]
        variants = {
            "bhrt_ref": plain + [V("Main"), V("MtlBlinn")],
            "bhrt_ref_pm": plain + [V("Main_pm"), V("MtlBlinn_pm")],
            "bhrt_ref_cnt": [V(st + "_cnt") for st in ("GenLight", "Box", "TriObj", "Texture", "driver", "Main")] + [V("MtlBlinn")],
            "bhrt_ref_pm_cnt": [V(st + "_cnt") for st in ("GenLight", "Box", "TriObj", "Texture", "driver")] + [V("Main_pm_cnt"), V("MtlBlinn_pm_cnt")],
            "bhrt_ref_gpu": [V("GenLight"), V("Box"), V("TriObj"), V("Texture"), V("driver_nomain"), V("Main"), V("MtlBlinn"), V("GpuRender")],
        }
        for name, extra_objs in variants.items():
            exe = os.path.join(OUT, name)
            subprocess.check_call(["g++", "-fopenmp", "-o", exe] + common + extra_objs + ["-ldl"])
            print("built", exe)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
