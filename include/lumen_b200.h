/*
 * lumen_b200.h — C ABI of the B200-native wavefront path-tracing core.
 *
 * This is the drop-in boundary for the reference's `LumenRenderer` / `WaveFront::WaveFrontRenderer`
 * hot path (scene/mesh/material/light upload, camera, frame settings, TraceFrame, output read-back).
 * Every entry point cites the reference interface it replaces. Paths are relative to
 * /root/reference/Lumen_Engine/ :
 *   LM/ = Lumen/src/Lumen/            PT/ = LumenPT/src/
 *
 * Conventions
 *   - plain C, opaque renderer pointer, integer resource handles (>= 0), plain pointers + sizes;
 *   - every call returns LB_OK (0) or a negative LB_ERR_* code; lb_last_error() gives the text;
 *   - inputs are copied before the call returns (reference uploads are synchronous as well,
 *     PT/Framework/MemoryBuffer.cpp:50-56), so callers may free host memory immediately;
 *   - one renderer instance per GPU, not thread-safe across concurrent calls on the same instance;
 *   - there is NO CPU fallback: lb_create fails with LB_ERR_CUDA when no sm_100 device is usable.
 *
 * The same declarations, prefixed `lo_` instead of `lb_`, are exported by the CPU oracle
 * (oracle/liblumen_oracle.so) — test infrastructure only, never linked by the product.
 */
#ifndef LUMEN_B200_H
#define LUMEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef LB_API
#define LB_API __attribute__((visibility("default")))
#endif

typedef struct LbRendererOpaque* LbRenderer;
typedef int32_t LbHandle;
#define LB_NO_HANDLE (-1)

enum {
    LB_OK = 0,
    LB_ERR_INVALID_ARGUMENT = -1,
    LB_ERR_INVALID_HANDLE = -2,
    LB_ERR_CUDA = -3,
    LB_ERR_OUT_OF_MEMORY = -4,
    LB_ERR_UNSUPPORTED = -5,
    LB_ERR_STATE = -6
};

/* Lumen::EmissionMode, LM/ModelLoading/MeshInstance.h:14-19 */
enum { LB_EMISSION_ENABLED = 0, LB_EMISSION_DISABLED = 1, LB_EMISSION_OVERRIDE = 2 };

/* WaveFront::LightChannel, PT/Shaders/CppCommon/WaveFrontDataStructs/LightData.h:12-19 */
enum { LB_CHANNEL_DIRECT = 0, LB_CHANNEL_INDIRECT = 1, LB_CHANNEL_SPECULAR = 2, LB_CHANNEL_VOLUMETRIC = 3, LB_NUM_CHANNELS = 4 };

/* Volume shading model. COMPAT = the reference's fixed 5-step constant-density march
 * (PT/CUDAKernels/VolumetricKernels/GPUVolumetricShadeDirect.cu:8-101); DELTA = delta tracking with a
 * per-grid majorant over the (procedural / NanoVDB-style dense) density grid (north_star item 4). */
enum { LB_VOLUME_COMPAT = 0, LB_VOLUME_DELTA = 1 };

/* WaveFront::WaveFrontSettings, PT/Framework/WaveFrontRenderer.h:31-48 (+ ReSTIRSettings toggles,
 * PT/Shaders/CppCommon/ReSTIRData.h:25-66). depth counts extend waves (= bounces + 1). */
typedef struct LbSettings {
    uint32_t width;            /* renderResolution.x */
    uint32_t height;           /* renderResolution.y */
    uint32_t depth;            /* WaveFrontSettings::depth */
    uint32_t blend_output;     /* WaveFrontSettings::blendOutput (progressive accumulation) */
    uint32_t restir;           /* 1: ReSTIR at depth 0 (reference behaviour); 0: NEE at depth 0 (SURVEY hazard 11) */
    uint32_t restir_temporal;  /* ReSTIRSettings::enableTemporal */
    uint32_t restir_spatial;   /* ReSTIRSettings::enableSpatial */
    int32_t  device;           /* CUDA device ordinal (ignored by the oracle) */
    uint32_t volume_mode;      /* LB_VOLUME_* */
    uint32_t first_frame_count;/* value of the reference's static frameCount for the first frame minus 1 (0 = reference);
                                  sample sharding sets 2*rank so that streams do not overlap */
    uint32_t frame_count_stride;/* frameCount advance per frame; 0 means the reference's 2 (SURVEY hazard 10) */
    /* Row band of a larger frame (image-tile sharding across GPUs): this renderer produces rows
     * [band_row0, band_row0 + height) of a frame that is band_full_height rows tall (0 = not a band: the frame is `height` rows).
     * Camera, jitter, motion vectors and every per-pixel random stream use the pixel's position in the FULL frame, so a band
     * pixel whose ReSTIR neighbourhood lies inside the band equals the same pixel of the full-frame render bit for bit.
     * band_row0 * width must be a multiple of 256 (the RIS light-bag group, ReSTIRKernels.cu:423). */
    uint32_t band_row0;
    uint32_t band_full_height;
    uint32_t restir_unbiased;  /* !ReSTIRSettings::enableBiased (ReSTIRData.h:63; the reference ships `enableBiased = true`, so 0 is its behaviour):
                                  1 = temporal and spatial reuse take the CombineUnbiased branches, ReSTIRKernels.cu:905-970, :1123-1198 */
    /* Rows of the FULL frame whose radiance is wanted from this band renderer (band_own_rows = 0: every rendered row). The other rendered
     * rows are the ReSTIR halo: they are traced and shaded up to the primary surface record and take part in RIS / temporal / spatial reuse
     * (that is all a neighbour needs of them), but spawn no NEE shadow rays and no bounce rays. Ignored when the scene holds media. */
    uint32_t band_own_row0;
    uint32_t band_own_rows;
} LbSettings;

/* LumenRenderer::MaterialData, LM/Renderer/LumenRenderer.h:64-112 (defaults :66-82).
 * Texture handles may be LB_NO_HANDLE = the renderer's default 1x1 texture
 * (white; normal map default (128,128,255), LM/Renderer/LumenRenderer.cpp:50-58). */
typedef struct LbMaterialDesc {
    float diffuse_color[4];
    float emission[3];
    float transmission_factor;
    float clear_coat_factor;
    float clear_coat_roughness_factor;
    float index_of_refraction;
    float specular_factor;
    float specular_tint_factor;
    float subsurface_factor;
    float luminance;
    float anisotropic;
    float sheen_factor;
    float sheen_tint_factor;
    float metallic_factor;
    float roughness_factor;
    float tint_factor[3];
    float transmittance[3];
    LbHandle diffuse_texture;
    LbHandle normal_texture;
    LbHandle metallic_roughness_texture;
    LbHandle emissive_texture;
    LbHandle transmission_texture;
    LbHandle clear_coat_texture;
    LbHandle clear_coat_roughness_texture;
    LbHandle tint_texture;
} LbMaterialDesc;

/* LumenRenderer::PrimitiveData, LM/Renderer/LumenRenderer.h:44-61. Attribute streams are given as base
 * pointer + byte stride, which covers both the interleaved 48-byte `Vertex` (PT/Shaders/CppCommon/ModelStructs.h:21-28)
 * and the non-interleaved VectorViews. uvs/normals/tangents may be NULL (zeros / +Z / +X,w=1 are used). */
typedef struct LbPrimitiveDesc {
    const void* positions;  uint32_t position_stride;   /* float3 */
    const void* uvs;        uint32_t uv_stride;         /* float2 */
    const void* normals;    uint32_t normal_stride;     /* float3 */
    const void* tangents;   uint32_t tangent_stride;    /* float4, w = bitangent sign */
    uint32_t vertex_count;
    const void* indices;    uint32_t index_size;        /* 2 or 4 bytes, m_IndexSize */
    uint32_t index_count;                               /* 3 * triangles */
    LbHandle material;
} LbPrimitiveDesc;

/* MeshInstance::Emissiveness, LM/ModelLoading/MeshInstance.h:24-34 */
typedef struct LbEmissiveness {
    int32_t mode;               /* LB_EMISSION_* */
    float override_radiance[3];
    float scale;
} LbEmissiveness;

/* Procedural / dense float density grid standing in for nanovdb::FloatGrid
 * (PT/Framework/PTVolume.cpp:47-108 loads .vdb/.vndb; NanoVDB files: lb_volume_create_file / lb_nanovdb_* below).
 * density[(z*ny + y)*nx + x], world bbox = bbox_min..bbox_max in the volume's object space. */
typedef struct LbVolumeDesc {
    const float* density;       /* may be NULL: homogeneous medium of value 1 */
    uint32_t nx, ny, nz;
    float bbox_min[3];
    float bbox_max[3];
} LbVolumeDesc;

/* ---- lifetime: WaveFrontRenderer::Init, PT/Framework/WaveFrontRenderer.cpp:70-322 ---- */
LB_API int lb_create(const LbSettings* settings, LbRenderer* out);
LB_API int lb_destroy(LbRenderer r);
LB_API const char* lb_last_error(void);
LB_API const char* lb_version(void);

/* ---- resources ---- */
/* LumenRenderer::CreateTexture, LM/Renderer/LumenRenderer.h:161; PT/Framework/PTTexture.cpp:35-74
 * (RGBA8, wrap addressing, bilinear, normalised coords, optional sRGB decode). */
LB_API int lb_texture_create(LbRenderer r, const uint8_t* rgba8, uint32_t width, uint32_t height, int srgb, LbHandle* out);
/* LumenRenderer::CreateMaterial, LM/Renderer/LumenRenderer.h:164; PT/Framework/WaveFrontRenderer.cpp:1260-1311 */
LB_API int lb_material_create(LbRenderer r, const LbMaterialDesc* desc, LbHandle* out);
/* ILumenMaterial setters, LM/Renderer/ILumenResources.h:18-87 (re-uploads on next frame, PT/Framework/PTMaterial.cpp:21-36) */
LB_API int lb_material_update(LbRenderer r, LbHandle material, const LbMaterialDesc* desc);
/* LumenRenderer::CreatePrimitive, LM/Renderer/LumenRenderer.h:157; PT/Framework/WaveFrontRenderer.cpp:1148-1252 */
LB_API int lb_primitive_create(LbRenderer r, const LbPrimitiveDesc* desc, LbHandle* out);
/* LumenRenderer::CreateMesh, LM/Renderer/LumenRenderer.h:159 */
LB_API int lb_mesh_create(LbRenderer r, const LbHandle* primitives, uint32_t count, LbHandle* out);
/* LumenRenderer::CreateVolume, LM/Renderer/LumenRenderer.h:168 (file path replaced by an in-memory grid) */
LB_API int lb_volume_create(LbRenderer r, const LbVolumeDesc* desc, LbHandle* out);

/* ---- scene: ILumenScene::AddMesh / AddVolume, LM/ModelLoading/ILumenScene.h:11-71; PT/Framework/PTScene.cpp:67-171 ---- */
/* transform: row-major 4x4 world matrix (what PT/Framework/PTMeshInstance.cpp:143-147 uploads). */
LB_API int lb_scene_add_mesh_instance(LbRenderer r, LbHandle mesh, const float* transform16,
                                      const LbEmissiveness* emissiveness, LbHandle override_material, LbHandle* out);
/* Transform::Set*, MeshInstance::SetEmissiveness / SetOverrideMaterial, PT/Framework/PTMeshInstance.cpp:36-40,110-115 */
LB_API int lb_instance_set_transform(LbRenderer r, LbHandle instance, const float* transform16);
LB_API int lb_instance_set_emissiveness(LbRenderer r, LbHandle instance, const LbEmissiveness* emissiveness);
LB_API int lb_instance_set_override_material(LbRenderer r, LbHandle instance, LbHandle material);
/* ILumenScene::AddVolume + VolumeInstance::m_Density, LM/ModelLoading/VolumeInstance.h:24 */
LB_API int lb_scene_add_volume_instance(LbRenderer r, LbHandle volume, const float* transform16, float density, LbHandle* out);
LB_API int lb_scene_clear(LbRenderer r);

/* ---- camera: Camera::GetVectorData, LM/Renderer/Camera.cpp:79-93,122-140 ---- */
/* position + rotation quaternion (w,x,y,z); fovY is the reference's hard-coded 90 degrees (Camera.h:63) unless overridden. */
LB_API int lb_camera_set_pose(LbRenderer r, const float* position3, const float* rotation_wxyz);
/* The camera's world matrix as the reference keeps it (Camera::GetMatrixData's current matrix, LM/Renderer/Camera.cpp:95-104,128-140:
 * columns right / up / forward / position), given ROW-major (= glm::transpose of it), float precision preserved. Replaces the pose
 * until the next lb_camera_set_pose. This is what an adapter holding only a `Camera` object passes (it has no rotation getter). */
LB_API int lb_camera_set_matrix(LbRenderer r, const float* world16_row_major);
LB_API int lb_camera_set_fov_y(LbRenderer r, float degrees);
/* Camera::SetMinMaxRenderDistance, LM/Renderer/Camera.h:36-37,60 (default 0.1, 1000): normalisation range of the depth side output */
LB_API int lb_camera_set_min_max_distance(LbRenderer r, float min_distance, float max_distance);

/* ---- frame settings: LumenRenderer::Set/GetRenderResolution, SetBlendMode, LM/Renderer/LumenRenderer.h:178-196 ---- */
LB_API int lb_set_render_resolution(LbRenderer r, uint32_t width, uint32_t height);
LB_API int lb_get_render_resolution(LbRenderer r, uint32_t* width, uint32_t* height);
/* The settings the renderer currently runs with (WaveFrontRenderer::m_Settings). */
LB_API int lb_get_settings(LbRenderer r, LbSettings* out);
LB_API int lb_set_depth(LbRenderer r, uint32_t depth);
LB_API int lb_set_blend_mode(LbRenderer r, int blend);
LB_API int lb_get_blend_mode(LbRenderer r, int* blend);
/* Discards progressive accumulation and ReSTIR history (what a ResizeBuffers does, PT/Framework/WaveFrontRenderer.cpp:1424-1540). */
LB_API int lb_reset_history(LbRenderer r);

/* ---- the hot path: WaveFrontRenderer::TraceFrame, PT/Framework/WaveFrontRenderer.cpp:435-1089 ---- */
/* Renders `frames` frames back to back, asynchronously on the renderer's stream; read-backs synchronise. */
LB_API int lb_render_frames(LbRenderer r, uint32_t frames);
LB_API int lb_synchronize(LbRenderer r);
/* LumenRenderer::StartRendering (render thread, PT/Framework/WaveFrontRenderer.cpp:1109-1117) */
LB_API int lb_start_rendering(LbRenderer r);
LB_API int lb_stop_rendering(LbRenderer r);

/* ---- output: GetOutputTexturePixels, PT/Framework/WaveFrontRenderer.cpp:1379-1394 (+ fp32 HDR, north_star) ---- */
LB_API int lb_read_hdr(LbRenderer r, float* rgba32f, size_t capacity_bytes);      /* width*height*4 floats, merged/blended image */
/* The same read-back without blocking the caller (no reference counterpart: GetOutputTexturePixels, WaveFrontRenderer.cpp:1379-1394, is
 * synchronous). The copy of the frame just rendered runs on a copy engine while the next frame renders; `rgba32f` should be pinned host
 * memory and must not be touched until lb_readback_wait() returns. One read-back can be pending per renderer. */
LB_API int lb_read_hdr_async(LbRenderer r, float* rgba32f, size_t capacity_bytes);
LB_API int lb_readback_wait(LbRenderer r);
LB_API int lb_read_ldr(LbRenderer r, uint8_t* rgba8, size_t capacity_bytes);      /* clamp + sRGB OETF + 8 bit, GPUShadingKernels.cu:28-56 */
LB_API int lb_read_channel(LbRenderer r, int channel, float* rgba32f, size_t capacity_bytes);
LB_API int lb_read_motion_vectors(LbRenderer r, float* xy32f, size_t capacity_bytes); /* MotionVectors.cu:8-55 (fp16-rounded values) */

/* G-buffer side outputs of the frame just rendered, for a denoiser / upscaler behind the path (any pointer may be NULL):
 *   depth             float[N]   ExtractDepthDataGpu / ExtractNRD_DLSSdataGpu, PT/CUDAKernels/WaveFrontKernels/GPUExtractDepthData.cu:6-72,
 *                                GPUExtractNRD_DLSSdata.cu:6-89: (t - min(minD, t)) / (max(maxD, t) - min(minD, t)), 0 where t < 0
 *   normal_roughness  float4[N]  shading normal + MaterialData::GetRoughness(); the reference writes a half4 surface, values are rounded
 *                                through fp16 accordingly (GPUExtractNRD_DLSSdata.cu:77-86)
 *   albedo            float4[N]  m_MaterialData.m_Color (PrepareOptixDenoisingGPU, GPUPostProcessingEffects.cu:13-50)
 * Motion vectors: lb_read_motion_vectors. */
LB_API int lb_read_gbuffer(LbRenderer r, float* depth, float* normal_roughness4, float* albedo4, size_t pixel_capacity);

/* FrameStats, LM/Renderer/LumenRenderer.h:29-34: per-stage device time (CUDA events) of the last frame, microseconds.
 * names are returned as a single ';'-separated string valid until the next call. */
LB_API int lb_frame_stats(LbRenderer r, const char** names, float* micros, uint32_t capacity, uint32_t* count);
/* counters of the last frame: [0]=extend rays, [1]=shadow rays, [2]=ReSTIR visibility rays, [3]=kernel launches,
 * [4]=lights, [5]=triangles, [6]=bvh nodes, [7]=bvh bytes, [8]=bvh build time (us, device time of the build proper: what a re-commit costs),
 * [9]=bvh levels, [10]=PLOC rounds, [11]=traversal-stack overflows (must be 0), [12]=last refit (us), [13]=refits since the last build,
 * [14]=host time of the build's allocations (us; paid by the first commit of a renderer only) */
LB_API int lb_frame_counters(LbRenderer r, uint64_t* values, uint32_t capacity, uint32_t* count);

/* ---- output stage behind the path ---- */
/* Screenshot: the 8-bit output image (lb_read_ldr) as a PNG file — colour type 6, 8 bit, rows top to bottom. Replaces
 * OutputLayer::MakeScreenshot = GetOutputTexturePixels + stbi_write_png(w, h, 4, pixels, 0) (Sandbox/src/OutputLayer.cpp:882-896).
 * LB_ERR_INVALID_ARGUMENT when the file cannot be written. */
LB_API int lb_save_png(LbRenderer r, const char* path);
/* FrameStats export (LM/Renderer/LumenRenderer.h:29-34, consumed by the profiler pane Sandbox/src/OutputLayer.cpp:377-420) as one
 * JSON object: {"frame_id", "resolution", "times_us": {stage: microseconds, ...}, "counters": {name: value, ...}}. Writes at most
 * `capacity` bytes including the terminating 0; `*needed` (may be NULL) receives the full length + 1. LB_ERR_INVALID_ARGUMENT when the
 * buffer is too small (nothing is written then). */
LB_API int lb_frame_stats_json(LbRenderer r, char* json, size_t capacity, size_t* needed);

/* ---- asset ingest in front of the path: glTF 2.0 (.gltf + .bin / data URIs, .glb) ----
 * Restates LumenPTModelConverter (PT/Tools/LumenPTModelConverter.cpp): material mapping :347-531, accessor extraction :1027-1059,
 * tangent generation :734-900, node hierarchy :953-1025 + :275-317. Host-only: lb_gltf_open needs no renderer and no GPU, the
 * parsed document can be inspected (what the parity tests do) and uploaded any number of times. */
typedef struct LbGltfOpaque* LbGltf;
/* Decoder for image formats other than PNG (the reference links stb_image, LumenPTModelConverter.cpp:121). Returns 0 on success with
 * *rgba8 = malloc()ed width*height*4 bytes, which the library frees. */
typedef int (*LbImageDecodeFn)(const uint8_t* bytes, size_t size, uint8_t** rgba8, uint32_t* width, uint32_t* height, void* user);
typedef struct LbGltfInfo {
    uint32_t images, undecoded_images;   /* undecoded images become the renderer's 1x1 default texture */
    uint32_t materials, meshes, primitives, instances, triangles, vertices;
} LbGltfInfo;
LB_API int lb_gltf_open(const char* path, LbImageDecodeFn decoder /* may be NULL */, void* user, LbGltf* out);
LB_API int lb_gltf_close(LbGltf g);
/* The model converter's cache (LumenPTModelConverter::ConvertGLTF, PT/Tools/LumenPTModelConverter.cpp:27-68: GenerateHeader :533-576 +
 * OutputToFile :588-598; on-disk records LumenPTModelConverter.h:79-186): writes the opened document as an `.ollad` file, byte for byte
 * the file the reference writes for the same asset. lb_gltf_open reads a path ending in ".ollad" back (LoadFile :70-273) instead of
 * parsing glTF, as SceneManager does when the cache exists. */
LB_API int lb_gltf_save_ollad(LbGltf g, const char* path);
/* SceneManager::LoadGLTF's cache protocol (LM/ModelLoading/SceneManager.cpp:55-76 over WaveFrontRenderer::OpenCustomFileFormat /
 * CreateCustomFileFormat, PT/Framework/WaveFrontRenderer.cpp:1135-1146): reads `<path without extension>.ollad` when it exists, otherwise
 * parses the glTF source and writes that cache next to it (a directory that cannot be written is not an error). */
LB_API int lb_gltf_open_cached(const char* path, LbImageDecodeFn decoder /* may be NULL */, void* user, LbGltf* out);
LB_API const char* lb_gltf_last_error(void);
LB_API int lb_gltf_info(LbGltf g, LbGltfInfo* out);
/* Inspection; returned pointers stay valid until lb_gltf_close. Texture members of the material are IMAGE indices of this document
 * (LB_NO_HANDLE = none), `material` of the primitive is a glTF material index (-1 = glTF default material). Indices are 32 bit. */
LB_API int lb_gltf_image(LbGltf g, uint32_t image, const uint8_t** rgba8, uint32_t* width, uint32_t* height, int* srgb, int* decoded);
LB_API int lb_gltf_material(LbGltf g, uint32_t material, LbMaterialDesc* out);
LB_API int lb_gltf_mesh_primitive_count(LbGltf g, uint32_t mesh, uint32_t* count);
LB_API int lb_gltf_primitive(LbGltf g, uint32_t mesh, uint32_t primitive, LbPrimitiveDesc* out);
LB_API int lb_gltf_instance(LbGltf g, uint32_t instance, uint32_t* mesh, float* transform16 /* row-major world matrix */);
/* CreateTexture / CreateMaterial / CreatePrimitive / CreateMesh / AddMesh for the whole document, in the order
 * LumenPTModelConverter::LoadFile issues them (:105-268). root_transform16 (row-major, may be NULL) is applied on the left of
 * every instance transform (SceneManager::LoadGLTF's a_TransformMat, LM/ModelLoading/SceneManager.cpp:41). */
LB_API int lb_gltf_upload(LbRenderer r, LbGltf g, const float* root_transform16, LbHandle* first_instance, uint32_t* instance_count);

/* ---- asset ingest in front of the path: NanoVDB files (.vndb / .nvdb) ----
 * Restates nanovdb::io::readGrid + the grid accessors of the NanoVDB the reference vendors (ABI 29.3.0,
 * LumenPT/vendor/openvdb/nanovdb/nanovdb/util/IO.h:107-160,301-352,573-592; NanoVDB.h:1890-1905,2184-2190,2394-2456,2733-2766,3022-3040),
 * which PTVolume::Load calls for ".vndb" (PT/Framework/PTVolume.cpp:93-98). Host-only: opening and inspecting a file needs no renderer
 * and no GPU. Float grids only (the reference reads nanovdb::FloatGrid); codecs NONE and ZIP (BLOSC: LB_ERR_UNSUPPORTED). */
typedef struct LbNanoVdbOpaque* LbNanoVdb;
enum { LB_NANOVDB_TYPE_FLOAT = 1 };                                   /* nanovdb::GridType::Float, NanoVDB.h:275-288 */
enum { LB_NANOVDB_CLASS_UNKNOWN = 0, LB_NANOVDB_CLASS_LEVEL_SET = 1, LB_NANOVDB_CLASS_FOG_VOLUME = 2 };   /* nanovdb::GridClass, NanoVDB.h:305-313 */
typedef struct LbNanoVdbInfo {
    uint32_t grid_type, grid_class;     /* GridData::mGridType / mGridClass */
    uint32_t version[3];                /* major (ABI), minor, patch */
    uint32_t codec;                     /* of the segment the grid came from: 0 none, 1 ZIP */
    uint32_t grid_count;                /* grids in the whole file */
    uint32_t node_count[4];             /* leaf, lower, upper, root (TreeData::mCount) */
    int32_t index_min[3], index_max[3]; /* Tree::bbox(), inclusive; max < min = no active voxel */
    double world_min[3], world_max[3];  /* Grid::worldBBox() */
    double voxel_size[3];
    double map_matrix[9];               /* row-major 3x3 of the index -> world map (Map::mMatD) */
    double map_translation[3];          /* Map::mVecD */
    uint64_t active_voxels, grid_bytes;
    float background, value_min, value_max;
    char name[256];
} LbNanoVdbInfo;
LB_API int lb_nanovdb_open(const char* path, uint32_t grid_index, LbNanoVdb* out);                               /* io::readGrid(fileName, n) */
LB_API int lb_nanovdb_open_memory(const void* bytes, size_t size, uint32_t grid_index, LbNanoVdb* out);         /* io::readGrid(istream, n) */
LB_API int lb_nanovdb_close(LbNanoVdb g);
LB_API const char* lb_nanovdb_last_error(void);
LB_API int lb_nanovdb_info(LbNanoVdb g, LbNanoVdbInfo* out);
/* ReadAccessor::getValue / isActive for n index coordinates (ijk3 = n x {i, j, k}); `active` may be NULL. */
LB_API int lb_nanovdb_values(LbNanoVdb g, const int32_t* ijk3, uint32_t n, float* values, uint8_t* active);
/* All values of the box index_min..index_max, out[((k - kmin) * ny + (j - jmin)) * nx + (i - imin)] — the layout of LbVolumeDesc::density.
 * as_density = 0: the stored values; 1: the density the renderer uses (fog volume: max(value, 0); level set: OpenVDB's sdfToFogVolume
 * ramp, inside min(1, -value / background), outside 0). */
LB_API int lb_nanovdb_dense(LbNanoVdb g, int as_density, float* out, size_t capacity_floats);
/* LumenRenderer::CreateVolume for a parsed grid: object-space box = the world-space box of the stored voxels (= Grid::worldBBox(), what
 * Shaders/volumetric_wavefront.cu:87 intersects), density field = lb_nanovdb_dense(as_density = 1). The index -> world map must be
 * scale + translation. */
LB_API int lb_volume_create_nanovdb(LbRenderer r, LbNanoVdb g, LbHandle* out);
/* LumenRenderer::CreateVolume(const std::string& path), LM/Renderer/LumenRenderer.h:168: dispatches on the extension like
 * PTVolume::Load (PT/Framework/PTVolume.cpp:69-107): ".vndb" / ".nvdb" are read here; ".vdb" (OpenVDB's own container, decoded by the
 * OpenVDB library in the reference) and anything else return LB_ERR_UNSUPPORTED. Error text: lb_nanovdb_last_error(). */
LB_API int lb_volume_create_file(LbRenderer r, const char* path, LbHandle* out);

/* ---- multi-GPU / framework interop ---- */
/* Device pointer of the fp32 RGBA accumulation buffer (sum over blended frames) and its frame count, for an
 * external NCCL reduce; lb_resolve_accum divides by `total_frames` and refreshes HDR/LDR. */
/* Device pointer of the merged fp32 RGBA frame (what lb_read_hdr copies): lets a multi-GPU host gather row bands device to device
 * (NCCL) on the renderer's stream. Valid until the resolution changes. The oracle returns its host buffer. */
LB_API int lb_hdr_buffer(LbRenderer r, void** device_ptr, size_t* bytes);
LB_API int lb_accum_buffer(LbRenderer r, void** device_ptr, size_t* bytes, uint32_t* frames);
LB_API int lb_resolve_accum(LbRenderer r, uint32_t total_frames);
/* Overlap mode, a mask (environment LB_OVERLAP sets it at creation). The reference serialises every kernel with cudaDeviceSynchronize
 * (PT/Framework/WaveFrontRenderer.cpp:604-850); here independent launches of a frame may share the device:
 *   bit 0 (default ON)  the shadow rays of bounce wave d run on a side stream under the extend launch of wave d + 1;
 *   bit 1 (default OFF — measured slower on B200) the ReSTIR passes run on the side stream beside ALL bounce waves;
 *   bit 2 (default ON)  the ReSTIR passes are launched after the first bounce wave; the later, latency-bound waves (1e5 .. 1e4 rays) run
 *                       on the side stream beside them (takes precedence over bit 0 where it applies: ReSTIR on, depth >= 3, no media).
 *   bit 3 (default OFF — measured 0.07 ms slower on C2) from the third wave on (1e5 .. 1e4 rays on C2) the rest of the bounce chain — extend,
 *                       shade, shadow of every remaining wave — is ONE launch in which a lane follows a path to its end (no media).
 * Results are identical in every mode; with mode 0 every stage time of lb_frame_stats is an exclusive device time (what bench.py uses for
 * its per-kernel roofline table). */
LB_API int lb_set_overlap(LbRenderer r, int mode);
/* Run all work on an externally owned CUDA stream (e.g. torch's current stream); 0/NULL = the renderer's own. */
LB_API int lb_set_stream(LbRenderer r, void* cuda_stream);
LB_API int lb_get_stream(LbRenderer r, void** cuda_stream);

/* ---- multi-GPU inside the library (csrc/lb_multigpu.cpp). The reference is single-GPU: WaveFrontRenderer owns one CUDA context
 * (PT/Framework/WaveFrontRenderer.cpp:70-322). NCCL (libnccl.so.2) is loaded at run time by the first call below — a single-GPU application
 * does not need it installed. Two partitionings, both without any exchange DURING a frame:
 *   samples  every GPU renders its own frames with a disjoint slice of the reference's frameCount sequence (first_frame_count = 2 * rank,
 *            frame_count_stride = 2 * ranks) into its fp32 accumulation buffer; ONE ncclReduce(sum) of that buffer on the renderers' streams
 *            produces the image on the root, which divides by the total frame count;
 *   bands    every GPU renders a row band of ONE frame plus a 60-row ReSTIR halo (lb_band_settings); the owned rows are gathered on the root
 *            (ncclSend / ncclRecv on the renderers' streams).
 * Errors: LB_ERR_* codes; text in lb_multigpu_last_error(). */
#define LB_RESTIR_HALO 60          /* spatial radius 30 px x 2 iterations, ReSTIRData.h:49,56 */
/* Settings of the renderer that produces rank `rank`'s band (+ halo) of the frame described by `full`: height / band_row0 / band_full_height /
 * band_own_row0 / band_own_rows are set; own_y0 / own_y1 (optional) receive the owned rows. band_row0 is lowered until band_row0 * width is a
 * multiple of 256 (the RIS light-bag group). */
LB_API int lb_band_settings(const LbSettings* full, uint32_t rank, uint32_t ranks, LbSettings* out, uint32_t* own_y0, uint32_t* own_y1);
/* Settings of sample-sharding rank `rank`: blend mode + the rank's frameCount stream. */
LB_API int lb_shard_settings(const LbSettings* base, uint32_t rank, uint32_t ranks, LbSettings* out);

/* Hooks the two layers below are built on (a host with its own collective library can use them the same way): lb_reduce_begin hands out a
 * side stream that has waited for the frames rendered so far, the accumulation buffer to send and — for the root — a separate buffer to
 * receive the sum in; lb_reduce_end(is_root, total_frames) resolves the image from that sum on the side stream (root) and arms a fence: the
 * renderer's next frame waits for the collective only before its merge kernel, read-backs wait on their copy stream — so the exchange runs
 * under the next frame. lb_reduce_wait makes the renderer's own stream wait for it (e.g. before timing with events on that stream). */
LB_API int lb_reduce_begin(LbRenderer r, void** side_stream, void** send_device, void** recv_device_root, size_t* bytes);
LB_API int lb_reduce_end(LbRenderer r, int is_root, uint32_t total_frames);
LB_API int lb_reduce_wait(LbRenderer r);

/* -- one rank per process (any launcher: the 128-byte id is created on one rank and handed to the others by the host application) */
#define LB_COMM_ID_BYTES 128
LB_API int lb_comm_unique_id(uint8_t* id128);
LB_API int lb_comm_init(LbRenderer r, const uint8_t* id128, int rank, int ranks);
/* samples: sum-reduce the accumulation buffers onto `root` (in place, on the renderer's stream); the root then resolves its HDR / LDR buffers as
 * sum / total_frames (total_frames = frames accumulated by ALL ranks). Asynchronous, and off the renderer's stream (lb_reduce_begin / _end): the
 * next frames overlap the collective. Every rank's accumulation buffer is left untouched (the root receives into a separate buffer);
 * lb_set_blend_mode(r, 1) starts the next image. */
LB_API int lb_comm_reduce_accum(LbRenderer r, int root, uint32_t total_frames);
/* bands: every rank sends the rows it owns of its merged frame to `root`; on the root `full_frame_device` receives the band_full_height x width
 * float4 image (device memory of the root's GPU; ignored on the other ranks). The renderer must have been created from lb_band_settings. */
LB_API int lb_comm_gather_bands(LbRenderer r, int root, void* full_frame_device);
LB_API int lb_comm_destroy(LbRenderer r);

/* -- one process driving n GPUs (ncclCommInitAll): what a C++ application on the adapter uses */
typedef struct LbGroup_t* LbGroup;
enum { LB_GROUP_SAMPLES = 0, LB_GROUP_BANDS = 1 };
/* Creates one renderer per device (settings derived from `settings` by lb_shard_settings / lb_band_settings) and one communicator over them.
 * The scene is loaded into every member through the ordinary entry points (lb_group_member). */
LB_API int lb_group_create(const int* devices, uint32_t n, const LbSettings* settings, int mode, LbGroup* out);
LB_API int lb_group_size(LbGroup g, uint32_t* n);
LB_API int lb_group_member(LbGroup g, uint32_t i, LbRenderer* out);
/* samples: `frames` frames on every member (n * frames samples); bands: `frames` complete frames, each gathered on member 0. Asynchronous. */
LB_API int lb_group_render(LbGroup g, uint32_t frames);
/* samples: the one collective + resolve on member 0 (bands: no-op). Asynchronous; lb_group_read_hdr synchronises. The accumulation buffers keep
 * accumulating afterwards (more frames + another reduce refine the same image); lb_group_reset starts the next progressive image. */
LB_API int lb_group_reduce(LbGroup g);
LB_API int lb_group_reset(LbGroup g);
/* The image on member 0: width x full height float4. */
LB_API int lb_group_read_hdr(LbGroup g, float* rgba, size_t capacity_bytes);
LB_API int lb_group_synchronize(LbGroup g);
LB_API int lb_group_destroy(LbGroup g);
LB_API const char* lb_multigpu_last_error(void);

/* ---- debug taps used by the parity tests ---- */
/* Trace caller-supplied rays with the extend kernel (closest hit) or the any-hit kernel.
 * rays: n x {ox,oy,oz,dx,dy,dz}; hits: n x {u32 instance, u32 primitive, f32 bary_u, f32 bary_v, f32 t} (t=-1: miss). */
LB_API int lb_debug_trace_closest(LbRenderer r, const float* rays6, uint32_t n, float tmin, float tmax, void* hits20);
LB_API int lb_debug_trace_any(LbRenderer r, const float* rays6, const float* tmax_per_ray, uint32_t n, float tmin, uint8_t* occluded);
/* Sorted emissive triangle list + CDF after the last frame: lights n x 16 floats (p0,p1,p2,normal,radiance,area). */
LB_API int lb_debug_read_lights(LbRenderer r, float* lights16, float* cdf, uint32_t capacity, uint32_t* count);
/* Primary-hit records of the last frame, one per pixel, same layout as lb_debug_trace_closest. */
LB_API int lb_debug_read_primary_hits(LbRenderer r, void* hits20, size_t capacity_bytes);
/* Primary surface data of the last frame, per pixel 24 floats:
 * position3,t, normal3,flags, tangent3,0, incoming3,0, transport3,0, color4 (SurfaceData.h:49-104). */
LB_API int lb_debug_read_surface(LbRenderer r, float* surf24, size_t capacity_bytes);
/* Current-frame reservoirs after ReSTIR::Run, per pixel 20 floats: weightSum, weight, sampleCount, pdf,
 * position3, area, normal3, 0, radiance3, 0, contribution3, 0 (ReSTIRData.h:107-183). */
LB_API int lb_debug_read_reservoirs(LbRenderer r, float* res20, size_t capacity_bytes);
/* Stand-alone BSDF evaluation on the device for parity against the reference headers
 * (disney.cuh:173-405). mat24: color4, transmittance3, ior(eta), tint3, luminance, metallic, subsurface, specular, roughness,
 * spectint, anisotropic, sheen, sheentint, clearcoat, clearcoatgloss, transmission, pad (params are byte-quantised like MaterialStructs.h:84-260). */
LB_API int lb_debug_eval_bsdf(LbRenderer r, const float* mat24, const float* n_t_wo_wi12, uint32_t n, float* bsdf_pdf4);
LB_API int lb_debug_sample_bsdf(LbRenderer r, const float* mat24, const float* n_t_wo_r12, uint32_t n, float* bsdf_wi_pdf_spec8);

#ifdef __cplusplus
}
#endif
#endif /* LUMEN_B200_H */
