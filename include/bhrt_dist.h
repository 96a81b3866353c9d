/* bhrt_dist.h -- C ABI of the multi-GPU host of the B200 rendering core (part of libbhrt.so).
 *
 * One PROCESS (or thread) per GPU, each with its own bhrt_ctx holding the same uploaded scene; this layer
 * shards one BeginRender() (reference Main.cpp:178-242) over the ranks and exchanges over NCCL (NVLink 5 /
 * NVSwitch inside a box) -- SURVEY.md 8(e):
 *
 *   deterministic frames (BHRT_MODE_DET)   interleaved rows per rank, one sum all-reduce of the accumulation and
 *                                          sample-0 AOV buffers (x + 0 is exact: the frame is byte-identical to
 *                                          the single-GPU one);
 *   stochastic frames (BHRT_MODE_PATH)     a sample range per rank over the full frame (the RNG is keyed by
 *                                          (seed, pixel, sample): the union of the ranges IS the single-GPU
 *                                          sample set), sum all-reduce of the float4 accumulation buffers;
 *   caustic photons                        emission quota per rank, all-gather of the reference's 24-byte Photon
 *                                          records, every rank builds the full gather structure; the power scale
 *                                          uses the GLOBAL stored count (reference Main.cpp:380).
 *
 * There is no exchange inside a render, so nothing to fuse into the render kernel: the collectives run on the
 * ctx's stream right behind it.  NCCL is bound at run time (dlopen of libnccl.so.2): a process that already
 * loaded one (PyTorch) shares it, a plain C++ host gets the system library.  The rendezvous is the caller's: rank 0
 * makes the 128-byte id, every rank passes the same bytes to bhrt_dist_create (MPI / a socket / a file /
 * torch.distributed's store -- bench.py uses the latter).  In one process with several GPUs (csrc/cli/bhrt_render
 * --gpus N) create the communicators from one thread per GPU.
 */
#ifndef BHRT_DIST_H_INCLUDED
#define BHRT_DIST_H_INCLUDED

#include "bhrt.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BHRT_DIST_ID_BYTES 128
typedef struct bhrt_dist bhrt_dist;

/* ---- the partitions (pure host arithmetic; usable without a GPU) ---- */
/* interleaved rows of `rank`: row_begin, row_begin + stride, ... < row_end */
int bhrt_dist_shard_rows(int32_t height, int32_t rank, int32_t world, int32_t *row_begin, int32_t *row_end, int32_t *row_stride);
/* contiguous, balanced sample range [begin, end) of `rank`; the ranges tile [0, spp) exactly */
int bhrt_dist_shard_samples(int32_t spp, int32_t rank, int32_t world, int32_t *begin, int32_t *end);
/* stored-photon quota of `rank` (the quotas sum to `total`); negative = bad arguments */
int32_t bhrt_dist_shard_photons(int32_t total, int32_t rank, int32_t world);

/* ---- the communicator ---- */
int bhrt_dist_unique_id(uint8_t id[BHRT_DIST_ID_BYTES]);
/* Collective over all ranks.  ctx must have a scene uploaded (the frame buffers' size is taken from it). */
int bhrt_dist_create(bhrt_ctx *ctx, int32_t rank, int32_t world, const uint8_t id[BHRT_DIST_ID_BYTES], bhrt_dist **out);
int bhrt_dist_destroy(bhrt_dist *d);

/* BeginRender() across the ranks: every rank calls it with the same params (sample / row ranges in params are
 * ignored: they are what this call fills in).  On return every rank holds the full frame in its device buffers;
 * rank `dst` also gets host copies of the outputs it asked for.  out->rays_* are the totals over all ranks,
 * out->kernel_ms this rank's render kernels. */
int bhrt_dist_render(bhrt_dist *d, const bhrt_params *params, int32_t dst, bhrt_outputs *out);

/* BuildCausticPhotonMap (reference Main.cpp:342-386) across the ranks: sharded emission, all-gather, replicated
 * build.  *stored = photons in the map (all ranks), *photon_rays = closest-hit queries traced for them. */
int bhrt_dist_photons(bhrt_dist *d, int32_t total_stored, uint64_t seed, int32_t *stored, uint64_t *photon_rays);

#ifdef __cplusplus
}
#endif
#endif
