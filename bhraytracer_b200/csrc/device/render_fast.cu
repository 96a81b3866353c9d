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
#define BHRT_NS bhrt_fast
#define BHRT_FAST 1
#include <cuda_runtime.h>
#include <cstring>
#include "wavefront.cuh"

extern "C" {

__attribute__((visibility("hidden")))
int bhrt_fast_prepare(int smem_bytes, int *blocks_per_sm)
{
    cudaError_t e = cudaFuncSetAttribute(bhrt_fast::wf_render, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    // one CTA per SM: leave L1 (where the thread-local lane copies live) all the rest
    cudaFuncSetAttribute(bhrt_fast::wf_render, cudaFuncAttributePreferredSharedMemoryCarveout, (smem_bytes + 2048) * 100 / (228 * 1024) + 1);
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, bhrt_fast::wf_render, WF_BLOCK, smem_bytes);
}

__attribute__((visibility("hidden")))
int bhrt_fast_launch(const void *view, size_t view_bytes, const void *args, size_t args_bytes, unsigned grid, int smem_bytes, cudaStream_t stream)
{
    if (view_bytes != sizeof(bhrt_fast::SceneView) || args_bytes != sizeof(bhrt_fast::WaveArgs)) return (int)cudaErrorInvalidValue;
    bhrt_fast::SceneView v;
    bhrt_fast::WaveArgs a;
    memcpy(&v, view, sizeof(v));
    memcpy(&a, args, sizeof(a));
    bhrt_fast::wf_render<<<grid, WF_BLOCK, smem_bytes, stream>>>(v, a);
    return (int)cudaGetLastError();
}

}  // extern "C"
