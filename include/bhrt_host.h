/* bhrt_host.h -- C ABI of the host-side scene pipeline (libbhrt_host.so).
 *
 * Keeps the reference's scene ingestion contract
 *     int LoadScene(char const *filename);          reference xmlload.cpp:65 (1 = ok, 0 = failed)
 * -- same XML schema, same Object/Material/Light/Texture model -- and
 * adds the flattening pass that packs the loaded scene into the device tables of bhrt.h.
 * Pure host code: no CUDA, no torch.
 */
#ifndef BHRT_HOST_H_INCLUDED
#define BHRT_HOST_H_INCLUDED

#include "bhrt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bhrt_host_scene bhrt_host_scene;

/* LoadScene + flatten.  Returns NULL on failure (bhrt_host_last_error() has the message the
 * reference would have printed).  Relative asset paths inside the XML are resolved against
 * the current working directory, exactly like the reference. */
bhrt_host_scene *bhrt_host_load_scene(const char *xml_path);
void bhrt_host_free_scene(bhrt_host_scene *scene);
const char *bhrt_host_last_error(void);

/* The flattened description; valid until bhrt_host_free_scene. */
const bhrt_scene_desc *bhrt_host_scene_desc(const bhrt_host_scene *scene);

/* Node names in the same DFS order as bhrt_scene_desc.nodes ("" for unnamed nodes). */
const char *bhrt_host_node_name(const bhrt_host_scene *scene, int32_t node);
/* Non-fatal loader messages (missing OBJ/texture files ...), '\n'-separated. */
const char *bhrt_host_warnings(const bhrt_host_scene *scene);

/* Override the camera resolution after loading (BASELINE configs re-use repo scenes at other sizes). */
int bhrt_host_set_resolution(bhrt_host_scene *scene, int32_t width, int32_t height);

/* RenderImage::SaveImage (reference Scenes/scene.h:628,634-644): 8-bit RGB (comp=3) or grey (comp=1) PNG. */
int bhrt_host_save_png(const char *path, const uint8_t *data, int32_t width, int32_t height, int32_t comp);

/* RenderImage::ComputeZBufferImage (reference Scenes/scene.h:579-600): 8-bit depth image, nearest = 255,
 * farthest = 0, misses (z == 1e30) = 0.  `z` is bhrt_outputs.z. */
void bhrt_host_zbuffer_image(const float *z, int32_t n, uint8_t *out);
/* RenderImage::ComputeSampleCountImage (scene.h:602-626): counts stretched to 0..255; returns the maximum count. */
int bhrt_host_sample_count_image(const uint8_t *count, int32_t n, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
