// photon.cuh -- caustic photon map on the GPU (emission, gather structure).  Host-side state
// object used by render.cu; kernels live in photon_kernels.cuh.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "core.cuh"

namespace bhrt {

struct PhotonState {
    PhotonRecord *records = nullptr;   // device
    int count = 0, capacity = 0;
    bool built = false;
    void release() { if (records) cudaFree(records); records = nullptr; count = capacity = 0; built = false; }
    PhotonGather gather_view(int k, float radius) const { PhotonGather g; memset(&g, 0, sizeof(g)); g.k = k; g.radius = radius; return g; }
    int emit(const SceneView &, cudaStream_t, int, int, unsigned long long, int *, unsigned long long *, std::string &err) { err = "photon emission not built yet"; return BHRT_ERR_STATE; }
    int set(const void *, int, bool, cudaStream_t, std::string &err) { err = "photon map not built yet"; return BHRT_ERR_STATE; }
    int build(int, cudaStream_t, std::string &err) { err = "photon map not built yet"; return BHRT_ERR_STATE; }
};

}  // namespace bhrt
