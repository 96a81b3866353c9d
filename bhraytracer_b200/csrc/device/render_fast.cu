// render_fast.cu -- the render kernel of wavefront.cuh compiled a second time, with
// --use_fast_math (approximate division / sqrt / sin / cos / pow, FMA contraction), into namespace
// bhrt_fast.  libbhrt.so uses it for the STOCHASTIC mode only (BHRT_MODE_PATH, unless the caller sets
// BHRT_FLAG_EXACT_MATH): there parity with the reference is statistical (RMSE against its high-spp
// render, BASELINE.json), and ulp-level differences in a sample's arithmetic are far below the
// Monte-Carlo noise.  The deterministic mode always runs the exact build of render.cu
// (-fmad=false, IEEE division and square root, libm-grade transcendentals), because its bar is
// bit-exact primary hits and <= 1/255 per pixel.
//
// Same source, same data layout: the launcher in render.cu hands over its SceneView / WaveArgs as
// plain bytes.
//
// This file is compiled TWICE (csrc/Makefile):
//   render_fast_bin.o  -DBHRT_FAST_WIDE=0  namespace bhrt_fastb  traversal phase = one ray per thread over
//                      the binary tree -- the default;
//   render_fast.o      -DBHRT_FAST_WIDE=1  namespace bhrt_fast   traversal phase = 8 lanes per ray over the
//                      8-wide regrouping of the BVH with the stacks in shared memory (wide.cuh), selected by
//                      BHRT_FLAG_WIDE_BVH / env BHRT_TRAVERSAL=wide (DESIGN.md 4.1 compares the two).
#ifndef BHRT_FAST_WIDE
#define BHRT_FAST_WIDE 1
#endif
#if BHRT_FAST_WIDE
#define BHRT_NS bhrt_fast
#define BHRT_FAST_ENTRY(name) bhrt_fast_##name
#define BHRT_WIDE_P2 1
#else
#define BHRT_NS bhrt_fastb
#define BHRT_FAST_ENTRY(name) bhrt_fastb_##name
#define BHRT_WIDE_P2 0
#endif
#define BHRT_FAST 1
#include <cuda_runtime.h>
#include <cstring>
#include "wavefront.cuh"

extern "C" {

__attribute__((visibility("hidden")))
int BHRT_FAST_ENTRY(prepare)(int smem_bytes, int *blocks_per_sm)
{
    cudaError_t e = cudaFuncSetAttribute(BHRT_NS::wf_render, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    // one CTA per SM: leave L1 (where the thread-local lane copies live) all the rest
    cudaFuncSetAttribute(BHRT_NS::wf_render, cudaFuncAttributePreferredSharedMemoryCarveout, WF_MIN_BLOCKS * (smem_bytes + 2048) * 100 / (228 * 1024) + 1);
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, BHRT_NS::wf_render, WF_BLOCK, smem_bytes);
}

__attribute__((visibility("hidden")))
int BHRT_FAST_ENTRY(launch)(const void *view, size_t view_bytes, const void *args, size_t args_bytes, unsigned grid, int smem_bytes, cudaStream_t stream)
{
    if (view_bytes != sizeof(BHRT_NS::SceneView) || args_bytes != sizeof(BHRT_NS::WaveArgs)) return (int)cudaErrorInvalidValue;
    BHRT_NS::SceneView v;
    BHRT_NS::WaveArgs a;
    memcpy(&v, view, sizeof(v));
    memcpy(&a, args, sizeof(a));
    BHRT_NS::wf_render<<<grid, WF_BLOCK, smem_bytes, stream>>>(v, a);
    return (int)cudaGetLastError();
}

// split gather (wavefront.cuh): sort keys of the parked queries; the queries in sorted order
__attribute__((visibility("hidden")))
int BHRT_FAST_ENTRY(gather_keys)(const void *args, size_t args_bytes, unsigned n, int cell_bits, uint32_t *keys, cudaStream_t stream)
{
    if (args_bytes != sizeof(BHRT_NS::WaveArgs)) return (int)cudaErrorInvalidValue;
    BHRT_NS::WaveArgs a;
    memcpy(&a, args, sizeof(a));
    BHRT_NS::pm_query_keys<<<(n + 255u) / 256u, 256, 0, stream>>>(a, n, cell_bits, keys);
    return (int)cudaGetLastError();
}
__attribute__((visibility("hidden")))
int BHRT_FAST_ENTRY(gather_queries)(const void *args, size_t args_bytes, const unsigned *sorted, unsigned n, unsigned *small, unsigned *heavy, unsigned grid, cudaStream_t stream)
{
    if (args_bytes != sizeof(BHRT_NS::WaveArgs)) return (int)cudaErrorInvalidValue;
    BHRT_NS::WaveArgs a;
    memcpy(&a, args, sizeof(a));
    // small[1] = cursor of the sorted queries, small[2] = number of deferred queries, small[3] = their cursor
    BHRT_NS::pm_gather_queries<<<grid, PMQ_BLOCK, 0, stream>>>(a, sorted, n, small + 1, heavy, small + 2);
    BHRT_NS::pm_gather_heavy<<<grid, PMQ_BLOCK, 0, stream>>>(a, heavy, small + 2, small + 3);
    return (int)cudaGetLastError();
}

}  // extern "C"
