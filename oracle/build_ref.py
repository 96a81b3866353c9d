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
  D-1  det mode: no jitter              (Main.cpp:153)
  D-2  det mode: GI term off            (MtlBlinn.cpp:124)
  D-3  det mode: all lights, exact L    (MtlBlinn.cpp:308)
  D-4  det mode: refraction gloss = 0   (MtlBlinn.cpp:116)
All D-* patches are runtime-switched by g_bhrt_det, so one binary serves both modes.
`bhrt_ref_pm` is the same build with -DUSE_PhotonMap (reference Main.cpp:51, MtlBlinn.cpp:25).
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
    "extern int g_bhrt_det, g_bhrt_spp, g_bhrt_gi, g_bhrt_photons; extern const char* g_bhrt_photon_path;\n"
    "class Node; void bhrt_record_primary(int,int,const Node*,float); void bhrt_record_linear(int,int,float,float,float);\n"
    "void bhrt_count_closest(); void bhrt_count_shadow(); void bhrt_photon_begin(); void bhrt_photon_end();\n"
)

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
        (1, "", lambda s: HOOKS + s),
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
         lambda s: "\tbhrt_photon_begin(); BuildCausticPhotonMap(lights, &rootNode); bhrt_photon_end();\n"),
        (218, "outColor = PathTracing(i, j);",
         lambda s: s + "\t\t\tbhrt_record_linear(i, j, outColor.r, outColor.g, outColor.b);\n"),
        (240, "SaveImages();", lambda s: "\t;\n"),
        (292, 'fopen("Resource/photonmap.dat", "wb")', lambda s: s.replace('"Resource/photonmap.dat"', "g_bhrt_photon_path")),
        (294, "fclose(fp);", lambda s: s.rstrip("\n") + " return true;\n"),
        (383, 'fopen("Resource/causticPhotonMap.dat", "wb")', lambda s: s.replace('"Resource/causticPhotonMap.dat"', "g_bhrt_photon_path")),
        (385, "fclose(fp);", lambda s: s.rstrip("\n") + " return true;\n"),
        (390, "if (root->GetNumChild() <= 0) return;", lambda s: "\tif (root == &rootNode) bhrt_count_closest();\n" + s),
    ],
    "Materials/Blinn/MtlBlinn.cpp": [
        (1, "", lambda s: "extern int g_bhrt_det;\n" + s),
        (116, "if (glossiness > 50) refraction_glossiness = glossiness;",
         lambda s: s.replace("if (glossiness > 50)", "if (!g_bhrt_det && glossiness > 50)")),
        (124, "outColor += PathTracing_GlobalIllumination(", lambda s: "\tif (!g_bhrt_det) " + s.lstrip()),
        (308, "{", lambda s: DET_DIRECT + s),
    ],
    "Lights/GenLight.cpp": [
        (1, "", lambda s: "void bhrt_count_shadow();\n" + s),
        (12, "return ShadowRayRecursive(&rootNode, ray, t_max) ? 0.f : 1.f;", lambda s: "\tbhrt_count_shadow();\n" + s),
    ],
}

UNPATCHED = [
    "xmlload.cpp", "Lights/PointLight.cpp", "Objects/Box/Box.cpp", "Objects/Plane/Plane.cpp",
    "Objects/Sphere/Sphere.cpp", "Objects/TriObj/TriObj.cpp", "Textures/Texture.cpp",
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
        jobs.append((os.path.join(HERE, "ref_build/driver.cpp"), os.path.join(tmp, "driver.o"), []))
        jobs.append((patched_copy("Lights/GenLight.cpp", tmp), os.path.join(tmp, "GenLight.o"), ["-I" + REF + "/Lights"]))
        main_src = patched_copy("Main.cpp", tmp)
        mtl_src = patched_copy("Materials/Blinn/MtlBlinn.cpp", tmp)
        for tag, extra in (("", []), ("_pm", ["-DUSE_PhotonMap"])):
            jobs.append((main_src, os.path.join(tmp, "Main%s.o" % tag), ["-Dmain=ref_main_unused"] + extra))
            jobs.append((mtl_src, os.path.join(tmp, "MtlBlinn%s.o" % tag), ["-I" + REF + "/Materials/Blinn"] + extra))

        def cc(job):
            src, obj, extra = job
            subprocess.check_call(["g++"] + flags + extra + ["-c", src, "-o", obj])
            return obj

        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            objs = list(ex.map(cc, jobs))
        common = [o for o in objs if "/Main" not in o and "/MtlBlinn" not in o]
        for tag in ("", "_pm"):
            exe = os.path.join(OUT, "bhrt_ref" + tag)
            subprocess.check_call(["g++", "-fopenmp", "-o", exe] + common +
                                  [os.path.join(tmp, "Main%s.o" % tag), os.path.join(tmp, "MtlBlinn%s.o" % tag)])
            print("built", exe)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
