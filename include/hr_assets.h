/*
 * hr_assets.h — C ABI of libhr_assets.so: scene / asset ingestion in the reference's formats (C interface, host only, no GPU).
 *
 * The reference reads its inputs through third-party libraries that are not vendored in /root/reference:
 *   meshes        dw::Mesh::load -> assimp (aiProcess_Triangulate | GenSmoothNormals | FlipUVs | CalcTangentSpace),
 *                 external/dwSampleFramework/src/mesh.cpp:244-613; files: the .gltf test scenes and meshes/sponza.obj (src/common.cpp:347-513)
 *   images        dw::vk::Image::create_from_file -> stb_image (stbi_load / stbi_loadf), src/vk.cpp:136-190:
 *                 the blue-noise PNGs (src/blue_noise.cpp:5-33), the .hdr environment maps (src/common.cpp:11)
 *   BRDF LUT      raw 512 x 512 RG16F file, extras/brdf_preintegrate_lut.cpp:8-31
 *   scene tables  dw::RayTracedScene (instances -> meshes -> sub-meshes -> materials), extras/ray_traced_scene.cpp:269-613,
 *                 mesh id = running (instance, sub-mesh) counter of the G-buffer pass, src/g_buffer.cpp:141-175
 * This library restates those code paths from the published formats (Wavefront OBJ / MTL, glTF 2.0 incl. .glb and data URIs,
 * PNG / zlib-deflate, Radiance RGBE) and produces exactly the arrays hr_scene_build / hr_bluenoise_set / hr_brdf_lut_set take.
 * Everything returns 0 or a negative code; hra_last_error() holds the message (thread-local).
 */
#ifndef HR_ASSETS_H
#define HR_ASSETS_H
#include "hr_api.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { HRA_OK = 0, HRA_ERR_IO = -1, HRA_ERR_FORMAT = -2, HRA_ERR_UNSUPPORTED = -3, HRA_ERR_INVALID_ARG = -4 };
HR_API const char* hra_last_error(void);

/* ---- images (stb_image semantics of Image::create_from_file, vk.cpp:136-190) -------------------------------------------
 * 8-bit path: PNG (all colour types, bit depths 1-16, Adam7, tRNS); channels = the file's (grey 1, grey+alpha 2, RGBA 4) except
 * RGB which is expanded to RGBA with alpha 255 (vk.cpp:163-168).  16-bit samples keep their high byte (stb).  flip_vertical =
 * stbi_set_flip_vertically_on_load.  float path: Radiance .hdr (RGBE, flat or new-style RLE), always 4 channels (alpha 1). */
HR_API int  hra_image_load(const char* path, int flip_vertical, int* width, int* height, int* channels, uint8_t** data);
HR_API int  hra_image_load_memory(const uint8_t* bytes, size_t n, int flip_vertical, int* width, int* height, int* channels, uint8_t** data);
HR_API int  hra_image_loadf(const char* path, int flip_vertical, int* width, int* height, float** rgba);
HR_API void hra_image_free(void* data);
/* Write an 8-bit image (1, 2, 3 or 4 channels: grey, grey + alpha, RGB, RGBA) as a PNG — e.g. the RGBA8 output of hr_tonemap_render read back
 * with hr_pass_download (the reference has no screenshot function; its data/screenshot_*.jpg were taken by hand).  Filter type 0,
 * deflate "stored" blocks: valid for every decoder, no compression. */
HR_API int  hra_image_save_png(const char* path, int width, int height, int channels, const uint8_t* data);

/* BlueNoise::BlueNoise (src/blue_noise.cpp:21-33): <dir>/sobol_256_4d.png and <dir>/scrambling_ranking_128x128_2d_{1..256}spp.png.
 * sobol: 256 x 1 RGBA8; scrambling_ranking[slot]: 128 x 128 RGBA8, slot = log2(spp).  slots_loaded: bit s set when table s was
 * present (the 1-spp table and the Sobol' table are required).  Feed hr_bluenoise_set / hr_bluenoise_set_slot. */
HR_API int hra_bluenoise_load(const char* dir, uint8_t* sobol_256x4, uint8_t* scrambling_ranking_9x128x128x4, uint32_t* slots_loaded);
/* dw::BRDFIntegrateLUT (extras/brdf_preintegrate_lut.cpp:24-31): 512 * 512 * 2 halves, raw. */
HR_API int hra_brdf_lut_load(const char* path, uint16_t* rg16f_512x512);
/* The library's environment is one constant colour: the solid-angle weighted mean radiance of an equirectangular
 * .hdr map = the band-0 term of the irradiance SH the reference projects it to (extras/cubemap_sh_projection). */
HR_API int hra_environment_constant(const char* hdr_path, float rgb[3]);

/* ---- meshes (dw::Mesh, include/mesh.h:16-140) ----------------------------------------------------------------------------- */
typedef struct hra_mesh hra_mesh;
typedef struct hra_submesh { /* dw::SubMesh, mesh.h:26-36 (base_vertex is already folded into the indices, mesh.cpp:593-601) */
    uint32_t mat_idx;
    uint32_t index_count;
    uint32_t base_vertex;
    uint32_t base_index;
    uint32_t vertex_count;
    float    max_extents[3];
    float    min_extents[3];
} hra_submesh;
enum { HRA_TEX_ALBEDO = 0, HRA_TEX_NORMAL = 1, HRA_TEX_ROUGHNESS = 2, HRA_TEX_METALLIC = 3, HRA_TEX_EMISSIVE = 4 };

/* Mesh::load(path, load_materials = true, is_orca_mesh = false).  .obj (+ .mtl), .gltf (+ .bin / data URIs), .glb.
 * Like the reference's loader: one sub-mesh per assimp mesh (OBJ: per object / group / material run; glTF: per primitive, node
 * transforms are NOT applied, mesh.cpp:283-291 iterates aiScene::mMeshes); triangulated (fans); missing normals are generated
 * smooth per sub-mesh; v texture coordinates flipped; tangent frame computed from the UVs when the file has none; the tangent is
 * flipped for right-handedness (mesh.cpp:553-556); vertex.position.w = the sub-mesh's material index (mesh.cpp:544). */
HR_API int  hra_mesh_load(const char* path, hra_mesh** out);
HR_API void hra_mesh_destroy(hra_mesh* m);
HR_API void hra_mesh_counts(const hra_mesh* m, uint64_t* n_vertices, uint64_t* n_indices, uint64_t* n_submeshes, uint64_t* n_materials);
HR_API const hr_vertex*   hra_mesh_vertices(const hra_mesh* m);
HR_API const uint32_t*    hra_mesh_indices(const hra_mesh* m);
HR_API const hr_material* hra_mesh_materials(const hra_mesh* m); /* constant factors (textures are listed, not sampled: DESIGN.md §7) */
HR_API int                hra_mesh_submesh(const hra_mesh* m, uint32_t i, hra_submesh* out);
HR_API void               hra_mesh_extents(const hra_mesh* m, float mn[3], float mx[3]);
/* texture path a material references ("" when it uses the constant), resolved relative to the mesh file (mesh.cpp:318-489) */
HR_API const char*        hra_mesh_material_texture(const hra_mesh* m, uint32_t material, int kind);

/* ---- scenes (dw::RayTracedScene::create(backend, instances), extras/ray_traced_scene.h:15-21) ------------------------------ */
typedef struct hra_scene hra_scene;
HR_API hra_scene* hra_scene_create(void);
HR_API void       hra_scene_destroy(hra_scene* s);
/* one RayTracedScene::Instance { transform, mesh }; the mesh must outlive hra_scene_finalize */
HR_API int hra_scene_add_instance(hra_scene* s, const hra_mesh* mesh, const float model16[16]);
/* Flatten into hr_scene_build's arguments: meshes deduplicated and concatenated (ray_traced_scene.cpp:283-345), materials made
 * global per mesh, one hr_instance per (instance, sub-mesh) in the G-buffer pass's draw order so that the instance index is the
 * reference's mesh id (g_buffer.cpp:141-175). */
HR_API int hra_scene_finalize(hra_scene* s);
HR_API void               hra_scene_counts(const hra_scene* s, uint64_t* n_vertices, uint64_t* n_indices, uint64_t* n_instances, uint64_t* n_materials);
HR_API const hr_vertex*   hra_scene_vertices(const hra_scene* s);
HR_API const uint32_t*    hra_scene_indices(const hra_scene* s);
HR_API const hr_instance* hra_scene_instances(const hra_scene* s);
HR_API const hr_material* hra_scene_materials(const hra_scene* s);
/* The scene's material textures for hr_scene_set_textures (Material::load, material.cpp:100-190: albedo sRGB, the others linear; roughness /
 * metallic channel 1 / 2 for glTF's packed image, 0 otherwise, mesh.cpp:404-447).  hra_scene_finalize loads every referenced image once
 * (PNG; a file that is missing or in another format leaves the material constant and is listed in hra_scene_texture_warnings).  The
 * hr_texture::data pointers stay valid until the scene is finalized again or destroyed. */
HR_API void hra_scene_texture_counts(const hra_scene* s, uint64_t* n_textures, uint64_t* n_material_bindings);
HR_API const hr_texture*           hra_scene_textures(const hra_scene* s);
HR_API const hr_material_textures* hra_scene_material_textures(const hra_scene* s);
HR_API const char*                 hra_scene_texture_warnings(const hra_scene* s); /* one line per image that could not be used */

#ifdef __cplusplus
}
#endif
#endif
