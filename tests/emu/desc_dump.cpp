// desc_dump -- TEST INFRASTRUCTURE: loads a scene with this repository's loader (libbhrt_host.so, include/bhrt_host.h) and
// writes the canonical serialisation of its flattened description (oracle/ref_build/desc_dump.hpp), to be compared with
// the one oracle/_ref/bhrt_ref_gpu writes from the REFERENCE's loader for the same file.
#include <cstdio>
#include "bhrt_host.h"
#include "../../oracle/ref_build/desc_dump.hpp"

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.xml out.bin\n", argv[0]); return 2; }
    bhrt_host_scene *s = bhrt_host_load_scene(argv[1]);
    if (!s) { fprintf(stderr, "%s\n", bhrt_host_last_error()); return 3; }
    const bool ok = desc_dump(*bhrt_host_scene_desc(s), argv[2]);
    bhrt_host_free_scene(s);
    return ok ? 0 : 4;
}
