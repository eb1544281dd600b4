// lb_kernels.h — host-callable launchers of the wavefront, ReSTIR, volume and scene-preparation kernels.
// One stream, no host round trips inside a frame: queue sizes live in device counters and every kernel is a
// persistent grid (a multiple of the SM count) that strides over the device-side count.
#pragma once
#include "lb_host.h"

namespace lb {

// device counters (uint32): wavefront queue sizes + per-launch tickets for the dynamic ray fetch
enum : uint32_t {
    CNT_RAYS_A = 0, CNT_RAYS_B = 1, CNT_SHADOW = 2, CNT_VIS = 3, CNT_VOL_SHADOW = 4,
    CNT_STACK_OVERFLOW = 5, CNT_VIS2 = 6,   // CNT_VIS / CNT_VIS2: sizes of the binned visibility-ray queue of the two ReSTIR visibility passes
             // traversal-stack entries dropped this frame (Tracer::push at kTraceStack): must stay 0, surfaced in lb_frame_counters
    CNT_TICKET0 = 8,                 // tickets CNT_TICKET0 .. CNT_TICKET0+kMaxTickets-1, one per trace launch of a frame
    // worst case of the schedule lb_create accepts (depth 24, LB_VOLUME_COMPAT: extend + shadow + volume shadow per wave, 5 ReSTIR tickets) is 77
    kMaxTickets = 120, kNumCounters = 128
};
// device statistics (uint64): rays traced per kind this frame
enum : uint32_t { STAT_EXTEND = 0, STAT_SHADOW = 1, STAT_VIS = 2, kNumStats = 4 };

struct CameraBasis { float3 eye, U, V, W; };

struct LaunchCfg { int sms = 148; cudaStream_t stream = nullptr; TraceTuning trace, trace_any; };      // warp-scheduling knobs of closest-hit / any-hit launches

struct FrameView {
    uint32_t width = 0, height = 0, npix = 0;
    // row band of a larger frame (LbSettings::band_*): first row, rows of the full frame, index of the band's first pixel in the full frame
    uint32_t row0 = 0, full_height = 0, pix0 = 0;
    // pixels [own_pix0, own_pix1) of this renderer are the ones whose radiance is wanted (LbSettings::band_own_*): the rest is ReSTIR halo and
    // spawns neither NEE shadow rays nor bounce rays
    uint32_t own_pix0 = 0, own_pix1 = 0xFFFFFFFFu;
    RayQueue rays[2];
    uint4* hits = nullptr;            // per queue slot (depth > 0)
    uint4* primary_hits = nullptr;    // per pixel == per queue slot at depth 0
    ShadowQueue shadow;
    float4* surf_cur = nullptr; float4* surf_prev = nullptr;      // kSurfPlanes planes each
    float4* res_cur = nullptr; float4* res_prev = nullptr; float4* res_tmp_a = nullptr; float4* res_tmp_b = nullptr;   // kResPlanes planes each
    float4* channels = nullptr;       // LB_NUM_CHANNELS planes
    float4* combined = nullptr; float4* accum = nullptr; float2* motion = nullptr; uchar4* ldr = nullptr;
    float4* vol_hits = nullptr;       // per queue slot: t0, t1, density, volume-instance (int bits, < 0 = none)
    ShadowQueue vol_shadow;           // compat-mode volumetric shadow rays, 5 per ray and wave
    uint32_t* counters = nullptr; unsigned long long* stats = nullptr;
};

struct DevVolume {                    // dense density grid standing in for nanovdb::FloatGrid + its instance
    float inv[12];                    // world -> volume object space (row-major 3x4)
    float3 lo, hi;                    // object-space bounding box
    const float* density;             // nx*ny*nz floats or nullptr (homogeneous 1)
    uint32_t nx, ny, nz;
    float instance_density, majorant;
};

struct ShadeArgs {
    uint32_t depth, max_depth, seed;
    int do_nee, nee_channel, do_bounce;
    float prev_view_proj[16];         // projection * inverse(previous camera), MotionVectors.cu:8-55
    const DevVolume* volumes; uint32_t num_volumes;
    int volume_mode;                  // LB_VOLUME_COMPAT: the reference's 5-step march; LB_VOLUME_DELTA: delta / ratio tracking
};

struct RestirArgs {
    uint32_t seed; int temporal, spatial;
    int unbiased = 0;                 // LbSettings::restir_unbiased: the CombineUnbiased branches of temporal / spatial reuse
    int ris_simple = 1;               // RIS may take the lean BSDF evaluation for rows of simple materials (LB_RIS_SIMPLE=0: always the general one)
    void (*lap)(void* user, const char* stage) = nullptr; void* lap_user = nullptr;      // per-kernel timing marks (CUDA events of the renderer)
};

// ---- wavefront (lb_wavefront.cu)
void launch_raygen(const LaunchCfg&, const FrameView&, const CameraBasis&, uint32_t frame_count);
void launch_extend(const LaunchCfg&, const FrameView&, const BvhView&, int queue, uint32_t ticket, bool primary, float tmin, float tmax);
// ---- volumes (lb_volume.cu)
void launch_volume_extend(const LaunchCfg&, const FrameView&, int queue, bool primary, const DevVolume* volumes, uint32_t num_volumes, float tmin, float tmax);
void launch_volume_delta(const LaunchCfg&, const FrameView&, const SceneView&, int queue, bool primary, const ShadeArgs&);
void launch_shade(const LaunchCfg&, const FrameView&, const SceneView&, int queue, const ShadeArgs&);
void launch_shadow(const LaunchCfg&, const FrameView&, const BvhView&, uint32_t ticket, float tmin);
// waves first_depth .. max_depth - 1 of the bounce chain in one launch, a lane per path (no media): rays[queue] holds the input of wave first_depth
void launch_tail(const LaunchCfg&, const FrameView&, const SceneView&, const BvhView& closest, const BvhView& any_hit, int queue, uint32_t first_depth, uint32_t max_depth,
                 uint32_t seed, float tmin, float tmax);
void launch_volume_shadow(const LaunchCfg&, const FrameView&, const BvhView&, uint32_t ticket, float tmin);
void launch_merge(const LaunchCfg&, const FrameView&, int blend, uint32_t blend_count);
void launch_resolve(const LaunchCfg&, const FrameView&, float inv_frames);
void launch_gbuffer(const LaunchCfg&, const float4* surf_planes, uint32_t npix, float min_d, float max_d, float* depth, float4* normal_rough, float4* albedo);
void launch_debug_trace(const LaunchCfg&, const BvhView&, const float* rays6, const float* tmax_per_ray, uint32_t n, float tmin, float tmax, void* hits20, uint8_t* occluded);
void launch_debug_bsdf(const LaunchCfg&, const float* mat24, const float* v12, uint32_t n, float* out, bool sample);
void launch_debug_surface(const LaunchCfg&, const float4* planes, uint32_t npix, float* out24);
void launch_debug_reservoirs(const LaunchCfg&, const float4* planes, uint32_t npix, float* out20);
void launch_debug_hits(const LaunchCfg&, const uint4* hits, uint32_t n, void* hits20);

// ---- ReSTIR (lb_restir.cu)
struct RestirBuffers { uint2* bags = nullptr; uint2* ris_order = nullptr; float4* vis_ray_o = nullptr; float4* vis_ray_d = nullptr;
                       const void* tmap_geom = nullptr; unsigned long long* spatial_nb = nullptr; };      // spatial_nb: per pixel, the accepted neighbours of the first spatial pass (the second pass draws the same ones)      // CUtensorMap of the current frame's surface plane 1 (k_spatial_tma), nullptr = gather from global memory      // vis_ray_*: binned visibility-ray queue (o.xyz, tmax | d.xyz, pixel), nullptr = trace straight from the reservoirs      // kNumBags*kLightsPerBag entries {light index, pdf bits}; ceil(npix/256) {pixel group, bag} sorted by bag
void launch_restir(const LaunchCfg&, const FrameView&, const SceneView&, const BvhView&, const RestirBuffers&, const RestirArgs&, uint32_t& ticket);

// ---- scene preparation (lb_scene.cu)
struct ScenePrepIn {
    const DevEntry* entries; uint32_t num_entries; uint32_t total_tris;
    const uint32_t* indices; const float4* vtx_pos;
};
void launch_flatten(const LaunchCfg&, const ScenePrepIn&, DevTri* out);
// per (primitive, triangle) emissive flag: FindEmissivesGpu, GPUEmissiveLookup.cu:13-109
struct DevPrimRange { uint32_t index_base, vertex_base, tri_count, material, flag_offset; };
void launch_find_emissives(const LaunchCfg&, const SceneView&, const DevPrimRange* prims, uint32_t num_prims, uint32_t total_prim_tris,
                           uint8_t* flags, uint32_t* per_prim_counts);
// emissive triangle list + sorted CDF: LightDataBuffer.cpp:37-125, GPUDataBufferKernels.cu:66-186, ReSTIRKernels.cu:49-130
struct LightBuild {
    DevBuf<DevLight> lights; DevBuf<float> cdf; uint32_t num_lights = 0; float cdf_sum = 0.f;
};
void build_lights(const LaunchCfg&, const SceneView&, const ScenePrepIn&, const uint8_t* prim_flags, LightBuild& out);

#ifdef LB_RIS_STATS
void dump_ris_stats();      // debug build: prints and clears the survivor statistics of k_ris
#endif
#ifdef LB_TRACE_STATS
void dump_trace_stats_wavefront(); void dump_trace_stats_restir();      // debug build: print and clear the traversal statistics
#endif
} // namespace lb
