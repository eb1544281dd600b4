// lb_api.cu — host side of liblumen_b200.so: renderer state, scene commit, the per-frame wavefront schedule and the
// C ABI declared in include/lumen_b200.h.
//
// Mirrors WaveFront::WaveFrontRenderer (under /root/reference/Lumen_Engine/LumenPT/src/Framework/):
//   Init / ResizeBuffers      WaveFrontRenderer.cpp:70-322, :1424-1540
//   TraceFrame                WaveFrontRenderer.cpp:435-1089   (seed handling :685,:830; frameCount :593,:1052)
//   Shade dispatch            ../CUDAKernels/WaveFrontKernels/CPUShadingKernels.cu:89-193
//   resource factories        WaveFrontRenderer.cpp:1148-1325, PTMaterial.cpp:97-266, PTMeshInstance.cpp:123-178
//   camera                    ../../../Lumen/src/Lumen/Renderer/Camera.cpp:79-140
// What is different by design: one CUDA stream, no cudaDeviceSynchronize / counter read-backs inside a frame (the
// reference syncs after every launch, WaveFrontRenderer.cpp:604-850), scene flattened to world space, shadow rays
// resolved per wave, fp32 channels. There is no CPU fallback anywhere in this file.
#include "../../include/lumen_b200.h"
#include "lb_kernels.h"
#include <cstdlib>
#include "lb_bsdf.cuh"
#include "lb_png.h"
#include <cuda.h>                     // CUtensorMap + the cuTensorMapEncodeTiled prototype (resolved at run time: no link against libcuda)
#include <vector>
#include <string>
#include <memory>
#include <mutex>
#include <thread>
#include <atomic>
#include <cstring>
#include <cmath>
#include <algorithm>

namespace lb {

struct HostTexture { uint32_t w = 1, h = 1; bool srgb = false; std::vector<uint8_t> px; };
struct HostMaterial { LbMaterialDesc desc; DevMaterial dev; };
struct HostPrimitive { std::vector<float4> pos, nu, tv; std::vector<float> tw; std::vector<uint32_t> idx; int material = 0; uint32_t num_lights = 0; };
struct HostMesh { std::vector<int> prims; };
struct HostInstance { int mesh; float m[16]; LbEmissiveness em; int override_mat; };
struct HostVolume { std::vector<float> density; uint32_t nx = 0, ny = 0, nz = 0; float3 lo, hi; float majorant = 1.f; };
struct HostVolumeInstance { int volume; float m[16]; float inv[16]; float density; };

static void invert_affine(const float* m, float* inv) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g), id = 1.0 / det;
    const double r[9] = {(e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id, (f * g - d * i) * id, (a * i - c * g) * id, (c * d - a * f) * id, (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id};
    for (int k = 0; k < 3; ++k) {
        inv[k * 4] = (float)r[k * 3]; inv[k * 4 + 1] = (float)r[k * 3 + 1]; inv[k * 4 + 2] = (float)r[k * 3 + 2];
        inv[k * 4 + 3] = (float)-(r[k * 3] * m[3] + r[k * 3 + 1] * m[7] + r[k * 3 + 2] * m[11]);
    }
    inv[12] = inv[13] = inv[14] = 0; inv[15] = 1;
}

struct Renderer {
    LbSettings st{};
    int device = 0; int sms = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    float cam_min_d = 0.1f, cam_max_d = 1000.f;           // Camera::m_MinMaxRenderDistance, Camera.h:60
    DevBuf<float> d_gbuffer;
    // asynchronous read-back (lb_read_hdr_async): a copy stream + two events; the next frame's merge waits for a pending copy
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_rendered = nullptr, ev_copied = nullptr; bool copy_pending = false;
    // asynchronous multi-GPU reduce (lb_reduce_begin / lb_reduce_end, used by csrc/lb_multigpu.cpp): the collective runs on a side stream behind
    // the frames rendered so far and — on the root — lands in `d_reduced`, from which the image is resolved; the next frame only waits for it
    // before its merge kernel (the one kernel that writes the accumulation buffer), so the reduce hides under the next frame
    cudaStream_t reduce_stream = nullptr; cudaEvent_t ev_frames = nullptr, ev_reduced = nullptr; bool reduce_pending = false; DevBuf<float4> d_reduced;
    std::mutex mu;

    // ---- host-side scene (the reference keeps the same tables in PTScene / SceneDataTable)
    std::vector<HostTexture> textures; std::vector<HostMaterial> materials; std::vector<HostPrimitive> prims; std::vector<HostMesh> meshes;
    std::vector<HostInstance> instances; std::vector<HostVolume> volumes; std::vector<HostVolumeInstance> vinstances;
    bool resources_dirty = true, scene_dirty = true;
    // only instance transforms changed since the last commit: the hierarchies are refitted instead of rebuilt (refit_scene)
    bool transforms_dirty = false; std::vector<uint32_t> inst_entry_begin;

    // ---- device-side scene
    DevBuf<uchar4> d_texels; DevBuf<DevTexture> d_textures; DevBuf<DevMaterial> d_materials; DevBuf<float> d_srgb_lut;
    DevBuf<uint32_t> d_indices; DevBuf<float4> d_vtx_pos, d_vtx_nu, d_vtx_tv; DevBuf<float> d_vtx_tw;
    DevBuf<DevPrimRange> d_prim_ranges; DevBuf<uint8_t> d_prim_flags; DevBuf<uint32_t> d_prim_counts;
    DevBuf<DevEntry> d_entries; DevBuf<DevTri> d_flat;
    DevBuf<DevVolume> d_volumes; std::vector<std::unique_ptr<DevBuf<float>>> d_volume_grids;
    std::vector<uint32_t> prim_index_base, prim_vertex_base, prim_flag_offset;
    std::vector<DevEntry> h_entries;
    // Two hierarchies over the same triangles (hits are a pure function of ray and triangle set, so which hierarchy answers is free):
    // `bvh` serves closest-hit rays (extend), `bvh_any` any-hit rays (shadow, ReSTIR visibility). DESIGN.md "Two hierarchies".
    DeviceBvh bvh, bvh_any; bool dual_bvh = false; LightBuild lights;
    uint32_t total_tris = 0;

    // ---- frame state
    float3 cam_pos = f3(0.f); float cam_q[4] = {1, 0, 0, 0}; float fov_y = 90.f;
    bool cam_from_matrix = false; float cam_m[16];        // lb_camera_set_matrix: the caller's float matrix, used as is
    double prev_cam[16]; bool have_prev_cam = false;
    uint32_t frame_index = 0, surf_cur = 0, res_cur = 0, blend_count = 0;
    uint32_t launches_last_frame = 0;

    // ---- per-resolution buffers
    DevBuf<float4> d_rays[2][3], d_shadow[3], d_surf[2], d_res[4], d_channels, d_combined, d_accum, d_vol_hits, d_vol_shadow[3];
    DevBuf<uint4> d_hits, d_primary_hits; DevBuf<float2> d_motion; DevBuf<uchar4> d_ldr;
    DevBuf<uint32_t> d_counters; DevBuf<unsigned long long> d_stats, d_spatial_nb; DevBuf<uint2> d_bags, d_ris_order;
    // direction-binned queue of the ReSTIR visibility rays (lb_restir.cu k_vis_bin), LB_VIS_SORT=1. Off by default: measured on C2 the binned
    // trace is 5.5 % faster (1.586 -> 1.499 ms for both passes) but the binning pre-pass costs 0.230 ms (profiles/r02_a_ab.md)
    // TMA descriptors of surface plane 1 (normal + signed depth) of both surface buffers: the spatial-reuse pass stages a 92 x 76 box of it per
    // 32 x 16-pixel tile in shared memory (lb_restir.cu k_spatial_tma), LB_SPATIAL_TMA=1. Off by default: measured on C2 the staged pass takes
    // 1.22 ms against 0.53 ms per pass — the probes are 15 % of the pass's gathers, and the 224 KB of tiles leave the reservoir / surface
    // gathers 1/8 of the L1 (profiles/r02_k_spatial_tma.md)
    CUtensorMap tmap_geom[2]; bool have_tmap = false;
    bool spatial_tma = []() { const char* e = getenv("LB_SPATIAL_TMA"); return e && atoi(e) != 0; }();
    void make_tensor_maps() {
        have_tmap = false;
        if (!spatial_tma) return;
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = []() -> EncodeFn {
            void* p = nullptr; cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
            return (EncodeFn)p;
        }();
        if (!encode) throw CudaError("LB_SPATIAL_TMA=1: the driver does not export cuTensorMapEncodeTiled");
        // plane 1 = the second float4 of every 32-byte record of plane pair 0 (lb_device.cuh): a 3-D tensor {2 eight-byte elements, W records
        // with a 32-byte stride, H rows}, based 16 bytes into the buffer
        const cuuint64_t dims[3] = {2, st.width, st.height}, strides[2] = {32, (cuuint64_t)st.width * 32};
        const cuuint32_t box[3] = {2, 32 + 2 * 30, 16 + 2 * 30}, estr[3] = {1, 1, 1};
        for (int k = 0; k < 2; ++k)
            if (encode(&tmap_geom[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d_surf[k].p + 1, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                throw CudaError("LB_SPATIAL_TMA=1: cuTensorMapEncodeTiled rejected the descriptor of surface plane 1");
        have_tmap = true;
    }
    DevBuf<float4> d_vis_rays[2];
    bool vis_sort = vis_sort_default();
    static bool vis_sort_default() { const char* e = getenv("LB_VIS_SORT"); return e && atoi(e) != 0; }
    uint64_t counters[16]{};

    // ---- FrameStats (LumenRenderer.h:29-34): CUDA events instead of host wall clock around forced syncs
    // A lap's time is measured from `prev`, the preceding lap of the same stream (overlap mode has two chains; the ReSTIR chain starts at
    // the fork lap)
    struct Lap { const char* name; cudaEvent_t ev; size_t prev; };
    std::vector<Lap> laps; std::vector<cudaEvent_t> event_pool; size_t events_used = 0;
    size_t last_lap[2] = {0, 0};
    // ---- overlap mode (lb_set_overlap): the ReSTIR passes depend only on the primary surface records and write only the DIRECT channel;
    // the bounce waves (extend / shade / shadow at depth > 0) write only the other channels. They run as two chains that fork after the
    // primary shade and join before the merge, so that the latency-bound tail waves can execute under the ReSTIR kernels. Off by default:
    // on B200 the two chains of persistent grids interfere (8.45 -> 8.87..9.45 ms/frame, profiles/r01_p_experiments.md).
    // bit 3 (default on) = from the third wave on the bounce chain is one launch, a lane per path (k_tail).
    // `overlap` is a mask: bit 0 (default on) = the shadow rays of bounce wave d run on the side stream under the extend of wave d + 1 — two
    // small latency-bound launches that share nothing but read-only data; bit 1 (default off) = the ReSTIR chain on the side stream beside ALL
    // bounce waves; bit 2 (default on) = the ReSTIR chain is launched after the first bounce wave and the later waves run beside it.
    int overlap = overlap_default(); cudaStream_t restir_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_shadow = nullptr;
    std::string stats_names;

    // ---- render thread (StartRendering, WaveFrontRenderer.cpp:1109-1117)
    std::thread render_thread; std::atomic<bool> stop_flag{false}; std::string thread_error;

    uint32_t npix() const { return st.width * st.height; }
    uint32_t full_height() const { return st.band_full_height ? st.band_full_height : st.height; }
    // LB_TRACE_REFILL_MIN / LB_TRACE_TRI_QUARTER: warp-scheduling knobs of trace_queue (profiling experiments; defaults in TraceTuning)
    static int overlap_default() { const char* e = getenv("LB_OVERLAP"); return e ? (atoi(e) & 15) : 5; }       // whole-chain overlap (bit 1) and the fused tail (bit 3) off: measured slower
    static TraceTuning trace_tuning(bool any) {
        TraceTuning t;
        if (const char* e = getenv(any ? "LB_TRACE_ANY_REFILL_MIN" : "LB_TRACE_REFILL_MIN")) t.refill_min = atoi(e);
        if (const char* e = getenv(any ? "LB_TRACE_ANY_TRI_QUARTER" : "LB_TRACE_TRI_QUARTER")) t.tri_quarter = atoi(e);
        return t;
    }
    LaunchCfg cfg() const { static const TraceTuning tune = trace_tuning(false), tune_any = trace_tuning(true); LaunchCfg c; c.sms = sms; c.stream = stream; c.trace = tune; c.trace_any = tune_any; return c; }

    ~Renderer() {
        stop_thread();
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        for (cudaEvent_t e : event_pool) cudaEventDestroy(e);
        if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
        if (reduce_stream) { cudaStreamSynchronize(reduce_stream); cudaStreamDestroy(reduce_stream); }
        if (ev_frames) cudaEventDestroy(ev_frames);
        if (ev_reduced) cudaEventDestroy(ev_reduced);
        if (ev_rendered) cudaEventDestroy(ev_rendered);
        if (ev_copied) cudaEventDestroy(ev_copied);
        if (own_stream) cudaStreamDestroy(own_stream);
        if (restir_stream) { cudaStreamSynchronize(restir_stream); cudaStreamDestroy(restir_stream); }
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_shadow) cudaEventDestroy(ev_shadow);
    }
    void stop_thread() {
        if (render_thread.joinable()) { stop_flag = true; render_thread.join(); }
        stop_flag = false;
    }

    void init() {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) throw CudaError("no CUDA device: liblumen_b200 has no CPU fallback");
        device = st.device;
        if (device < 0 || device >= count) throw CudaError("CUDA device ordinal out of range");
        LB_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop; LB_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) throw CudaError("liblumen_b200 is built for sm_100a (B200) only; no usable device and no fallback path");
        sms = prop.multiProcessorCount;
        LB_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
        {   // build scratch comes from the default stream-ordered pool (StreamBuf, lb_host.h): keep up to 4 GiB cached between scene commits
            cudaMemPool_t pool = nullptr; LB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
            uint64_t keep = 4ull << 30; LB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        stream = own_stream;
        LB_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        LB_CUDA(cudaEventCreateWithFlags(&ev_rendered, cudaEventDisableTiming)); LB_CUDA(cudaEventCreateWithFlags(&ev_copied, cudaEventDisableTiming));
        // sRGB decode table, computed in double exactly like the oracle does per texel
        float lut[256];
        for (int b = 0; b < 256; ++b) { const double c = b / 255.0; lut[b] = (float)(c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4)); }
        d_srgb_lut.upload(lut, 256, stream);
        LB_CUDA(cudaStreamSynchronize(stream));
        HostTexture white; white.px = {255, 255, 255, 255}; HostTexture nrm; nrm.px = {128, 128, 255, 255};   // LumenRenderer.cpp:50-58
        textures.push_back(white); textures.push_back(nrm);
        d_counters.reserve(kNumCounters); d_counters.zero(stream); d_stats.reserve(kNumStats); d_stats.zero(stream); d_bags.reserve(50 * 1000);
        resize();
    }

    void resize() {
        const size_t n = npix();
        d_ris_order.reserve((n + 255) / 256 + 128); d_spatial_nb.reserve(n);
        if (vis_sort) for (auto& p : d_vis_rays) p.reserve(n);          // + per-bag {start, count} and work tickets (lb_restir.cu k_ris_order)
        for (auto& q : d_rays) for (auto& p : q) p.reserve(n);
        for (auto& p : d_shadow) p.reserve(n);
        for (auto& p : d_surf) { p.reserve(n * kSurfPlanes); p.zero(stream); }
        for (auto& p : d_res) { p.reserve(n * kResPlanes); p.zero(stream); }
        d_channels.reserve(n * LB_NUM_CHANNELS); d_channels.zero(stream);
        d_combined.reserve(n); d_combined.zero(stream); d_accum.reserve(n); d_accum.zero(stream);
        d_vol_hits.reserve(n); for (auto& p : d_vol_shadow) p.reserve(n * 5);
        d_hits.reserve(n); d_primary_hits.reserve(n); d_primary_hits.zero(stream); d_motion.reserve(n); d_motion.zero(stream); d_ldr.reserve(n); d_ldr.zero(stream);
        blend_count = 0; frame_index = 0; surf_cur = 0; res_cur = 0; have_prev_cam = false;
        make_tensor_maps();
    }

    // ---- materials: WaveFrontRenderer::CreateMaterial (WaveFrontRenderer.cpp:1260-1311) + PTMaterial setters (PTMaterial.cpp:160-266)
    int tex_or(LbHandle h, int def) const { return h < 0 ? def : h; }
    bool fill_material(HostMaterial& m, const LbMaterialDesc& d) const {
        const LbHandle hs[8] = {d.diffuse_texture, d.normal_texture, d.metallic_roughness_texture, d.emissive_texture, d.transmission_texture, d.clear_coat_texture, d.clear_coat_roughness_texture, d.tint_texture};
        for (LbHandle h : hs) if (h >= (LbHandle)textures.size()) return false;
        m.desc = d;
        Material& p = m.dev.mat; memset(&p, 0, sizeof p);
        pack8(p.params.x, 1.f, 24);
        p.color = make_float4(d.diffuse_color[0], d.diffuse_color[1], d.diffuse_color[2], d.diffuse_color[3]);
        p.emissive = make_float4(d.emission[0], d.emission[1], d.emission[2], 0.f);
        pack8(p.params.z, d.transmission_factor, 16);
        pack8(p.params.z, d.clear_coat_factor, 0);
        pack8(p.params.z, 1.f - d.clear_coat_roughness_factor, 8);
        pack8(p.params.x, d.specular_factor, 16);
        pack8(p.params.y, d.specular_tint_factor, 0);
        pack8(p.params.x, d.subsurface_factor, 8);
        pack8(p.params.y, d.anisotropic, 8);
        pack8(p.params.y, d.sheen_factor, 16);
        pack8(p.params.y, d.sheen_tint_factor, 24);
        p.tint = make_float4(d.tint_factor[0], d.tint_factor[1], d.tint_factor[2], d.luminance);
        p.transmittance = make_float4(d.transmittance[0], d.transmittance[1], d.transmittance[2], d.index_of_refraction);
        pack8(p.params.x, d.roughness_factor, 24);
        pack8(p.params.x, d.metallic_factor, 0);
        m.dev.tex_diffuse = tex_or(d.diffuse_texture, 0); m.dev.tex_normal = tex_or(d.normal_texture, 1); m.dev.tex_mr = tex_or(d.metallic_roughness_texture, 0);
        m.dev.tex_emissive = tex_or(d.emissive_texture, 0); m.dev.tex_transmission = tex_or(d.transmission_texture, 0); m.dev.tex_coat = tex_or(d.clear_coat_texture, 0);
        m.dev.tex_coat_rough = 0;      // the reference never binds this slot (PTMaterial.cpp:121-129, SURVEY hazard 8): white
        m.dev.tex_tint = tex_or(d.tint_texture, 0);
        return true;
    }

    SceneView scene_view() const {
        SceneView v{};
        v.entries = d_entries.p; v.materials = d_materials.p; v.textures = d_textures.p; v.texels = d_texels.p; v.srgb_lut = d_srgb_lut.p;
        v.indices = d_indices.p; v.vtx_nu = d_vtx_nu.p; v.vtx_tv = d_vtx_tv.p; v.vtx_tw = d_vtx_tw.p;
        v.lights = lights.lights.p; v.cdf = lights.cdf.p; v.num_lights = lights.num_lights; v.cdf_sum = lights.cdf_sum;
        return v;
    }

    // ---- upload of textures / materials / vertex data (only when a resource was created or changed)
    void upload_resources() {
        std::vector<uchar4> texels; std::vector<DevTexture> tt;
        for (const HostTexture& t : textures) {
            DevTexture d{(uint32_t)texels.size(), t.w, t.h, t.srgb ? 1u : 0u}; tt.push_back(d);
            const uchar4* src = reinterpret_cast<const uchar4*>(t.px.data());
            texels.insert(texels.end(), src, src + (size_t)t.w * t.h);
        }
        d_texels.upload(texels.data(), texels.size(), stream); d_textures.upload(tt.data(), tt.size(), stream);
        std::vector<DevMaterial> mm; for (const HostMaterial& m : materials) mm.push_back(m.dev);
        d_materials.upload(mm.data(), mm.size(), stream);
        std::vector<uint32_t> idx; std::vector<float4> pos, nu, tv; std::vector<float> tw; std::vector<DevPrimRange> ranges;
        prim_index_base.clear(); prim_vertex_base.clear(); prim_flag_offset.clear();
        uint32_t flag_off = 0;
        for (const HostPrimitive& p : prims) {
            prim_index_base.push_back((uint32_t)idx.size()); prim_vertex_base.push_back((uint32_t)pos.size()); prim_flag_offset.push_back(flag_off);
            ranges.push_back(DevPrimRange{(uint32_t)idx.size(), (uint32_t)pos.size(), (uint32_t)p.idx.size() / 3u, (uint32_t)p.material, flag_off});
            idx.insert(idx.end(), p.idx.begin(), p.idx.end()); pos.insert(pos.end(), p.pos.begin(), p.pos.end());
            nu.insert(nu.end(), p.nu.begin(), p.nu.end()); tv.insert(tv.end(), p.tv.begin(), p.tv.end()); tw.insert(tw.end(), p.tw.begin(), p.tw.end());
            flag_off += (uint32_t)p.idx.size() / 3u;
        }
        d_indices.upload(idx.data(), idx.size(), stream); d_vtx_pos.upload(pos.data(), pos.size(), stream);
        d_vtx_nu.upload(nu.data(), nu.size(), stream); d_vtx_tv.upload(tv.data(), tv.size(), stream); d_vtx_tw.upload(tw.data(), tw.size(), stream);
        d_prim_ranges.upload(ranges.data(), ranges.size(), stream);
        d_prim_flags.reserve(std::max<uint32_t>(flag_off, 1u)); d_prim_counts.reserve(std::max<size_t>(prims.size(), 1)); d_prim_counts.zero(stream);
        if (!prims.empty()) {
            // FindEmissives per primitive (WaveFrontRenderer.cpp:1192-1209), all primitives in one parallel launch
            launch_find_emissives(cfg(), scene_view(), d_prim_ranges.p, (uint32_t)prims.size(), flag_off, d_prim_flags.p, d_prim_counts.p);
            std::vector<uint32_t> counts(prims.size());
            LB_CUDA(cudaMemcpyAsync(counts.data(), d_prim_counts.p, counts.size() * 4, cudaMemcpyDeviceToHost, stream));
            LB_CUDA(cudaStreamSynchronize(stream));
            for (size_t i = 0; i < prims.size(); ++i) prims[i].num_lights = counts[i];
        } else LB_CUDA(cudaStreamSynchronize(stream));
        resources_dirty = false;
    }

    // ---- scene commit: scene data table, world-space flattening, BVH, light list + CDF
    void commit_scene() {
        LB_CUDA(cudaStreamSynchronize(stream));        // frames in flight read the buffers replaced below
        if (resources_dirty) upload_resources();
        h_entries.clear(); total_tris = 0; inst_entry_begin.assign(instances.size() + 1, 0u);
        for (size_t i = 0; i < instances.size(); ++i) {
            const HostInstance& in = instances[i];
            inst_entry_begin[i] = (uint32_t)h_entries.size();
            bool mesh_emissive = false; for (int q : meshes[in.mesh].prims) mesh_emissive |= prims[q].num_lights > 0;
            for (int p : meshes[in.mesh].prims) {
                DevEntry e{}; memcpy(e.m, in.m, sizeof e.m);
                e.index_base = prim_index_base[p]; e.vertex_base = prim_vertex_base[p]; e.tri_count = (uint32_t)prims[p].idx.size() / 3u;
                e.material = (uint32_t)(in.override_mat >= 0 ? in.override_mat : prims[p].material);
                e.em_mode = in.em.mode; e.em_r = in.em.override_radiance[0]; e.em_g = in.em.override_radiance[1]; e.em_b = in.em.override_radiance[2]; e.em_scale = in.em.scale;
                e.tri_offset = total_tris; e.flag_offset = prim_flag_offset[p];
                // LightDataBuffer.cpp:37-125: which scene-table rows feed the light list
                bool on = in.em.mode != LB_EMISSION_DISABLED;
                if (in.em.mode == LB_EMISSION_ENABLED && !(mesh_emissive && prims[p].num_lights > 0)) on = false;
                e.lights_on = on ? 1u : 0u;
                total_tris += e.tri_count; h_entries.push_back(e);
            }
        }
        inst_entry_begin[instances.size()] = (uint32_t)h_entries.size();
        d_entries.upload(h_entries.data(), h_entries.size(), stream);
        LB_CUDA(cudaStreamSynchronize(stream));
        ScenePrepIn in{d_entries.p, (uint32_t)h_entries.size(), total_tris, d_indices.p, d_vtx_pos.p};
        d_flat.reserve(std::max<uint32_t>(total_tris, 1u));
        launch_flatten(cfg(), in, d_flat.p);
        {   // LB_BVH_BUILDER=lbvh selects the fastest build (Karras radix tree); default is the SAH-quality PLOC hierarchy
            const char* e = getenv("LB_BVH_BUILDER");
            // LB_BVH_SPLIT=k: early split clipping on a grid of (scene extent / k) for triangles larger than a cell (0: off)
            const char* sp = getenv("LB_BVH_SPLIT");
            const float split = sp ? (atof(sp) > 0.0 ? 1.f / (float)atof(sp) : 0.f) : 0.f;
            bvh_build(stream, d_flat.p, total_tris, bvh, (e && !strcmp(e, "lbvh")) ? BvhBuilder::LBVH : BvhBuilder::PLOC, getenv("LB_PLOC_RADIUS") ? atoi(getenv("LB_PLOC_RADIUS")) : 16, split);
            // LB_BVH_ANYHIT=same: one hierarchy for every ray (halves the build); default: a second PLOC hierarchy with a 128-wide search window
            const char* a = getenv("LB_BVH_ANYHIT");
            dual_bvh = !(a && !strcmp(a, "same")) && total_tris > 1u;
            if (dual_bvh) bvh_build(stream, d_flat.p, total_tris, bvh_any, BvhBuilder::PLOC, a && !strcmp(a, "ploc64") ? 64 : 128, split);
        }
        // the traversal keeps one pending sibling group per level (+ one transient entry) on a kTraceStack-entry stack: refuse a hierarchy it
        // cannot walk without dropping groups instead of rendering wrong hits
        if (std::max(bvh.levels, dual_bvh ? bvh_any.levels : 0u) + 2u > 64u) throw std::runtime_error("BVH deeper than the traversal stack (62 levels)");
        lights.num_lights = 0; lights.cdf_sum = 0.f;
        if (total_tris) build_lights(cfg(), scene_view(), in, d_prim_flags.p, lights);
        // volumes
        std::vector<DevVolume> dv; d_volume_grids.clear();
        std::vector<int> grid_of(volumes.size(), -1);
        for (const HostVolumeInstance& vi : vinstances) {
            const HostVolume& hv = volumes[vi.volume];
            DevVolume v{}; memcpy(v.inv, vi.inv, sizeof v.inv); v.lo = hv.lo; v.hi = hv.hi; v.nx = hv.nx; v.ny = hv.ny; v.nz = hv.nz;
            v.instance_density = vi.density; v.majorant = hv.majorant; v.density = nullptr;
            if (!hv.density.empty()) {
                if (grid_of[vi.volume] < 0) { d_volume_grids.emplace_back(new DevBuf<float>()); d_volume_grids.back()->upload(hv.density.data(), hv.density.size(), stream); grid_of[vi.volume] = (int)d_volume_grids.size() - 1; }
                v.density = d_volume_grids[grid_of[vi.volume]]->p;
            }
            dv.push_back(v);
        }
        d_volumes.upload(dv.data(), dv.size(), stream);
        LB_CUDA(cudaStreamSynchronize(stream));
        counters[4] = lights.num_lights; counters[5] = total_tris; counters[6] = bvh.num_nodes + (dual_bvh ? bvh_any.num_nodes : 0u); counters[7] = bvh.bytes() + (dual_bvh ? bvh_any.bytes() : 0u);
        counters[8] = (uint64_t)((bvh.build_ms + (dual_bvh ? bvh_any.build_ms : 0.f)) * 1000.f); counters[9] = bvh.levels; counters[10] = bvh.ploc_rounds;
        counters[14] = (uint64_t)((bvh.alloc_ms + (dual_bvh ? bvh_any.alloc_ms : 0.f)) * 1000.f);   // host time of the builds' allocations (first commit only)
        counters[12] = 0; counters[13] = 0;
        scene_dirty = false; transforms_dirty = false;
    }

    // ---- instances moved, nothing else changed (lb_instance_set_transform): the scene table gets the new matrices, the world-space triangles
    // are flattened again, and both hierarchies are REFITTED (bvh_refit: same topology, new boxes) instead of rebuilt — the reference rebuilds
    // its instance acceleration structures on every transform change (PTScene.cpp:74-156, PTMeshInstance.cpp:51-103). Everything is enqueued
    // on the renderer's stream behind the frames in flight. LB_REFIT_MAX bounds how many refits may follow a build (boxes only ever loosen).
    void refit_scene() {
        static const uint32_t max_refits = []() { const char* e = getenv("LB_REFIT_MAX"); return e ? (uint32_t)atoi(e) : 4096u; }();
        if (!total_tris || bvh.src_tris != total_tris || bvh.refits >= max_refits) { scene_dirty = true; commit_scene(); return; }
        for (size_t i = 0; i < instances.size(); ++i)
            for (uint32_t e = inst_entry_begin[i]; e < inst_entry_begin[i + 1]; ++e) memcpy(h_entries[e].m, instances[i].m, sizeof h_entries[e].m);
        d_entries.upload(h_entries.data(), h_entries.size(), stream);
        ScenePrepIn in{d_entries.p, (uint32_t)h_entries.size(), total_tris, d_indices.p, d_vtx_pos.p};
        launch_flatten(cfg(), in, d_flat.p);
        bvh_refit(stream, d_flat.p, total_tris, bvh);
        if (dual_bvh) bvh_refit(stream, d_flat.p, total_tris, bvh_any);
        lights.num_lights = 0; lights.cdf_sum = 0.f;
        build_lights(cfg(), scene_view(), in, d_prim_flags.p, lights);
        counters[4] = lights.num_lights;
        counters[12] = (uint64_t)((bvh.refit_ms + (dual_bvh ? bvh_any.refit_ms : 0.f)) * 1000.f); counters[13] = bvh.refits;
        transforms_dirty = false;
    }

    // ---- camera (Camera.cpp:79-93,122-140): row-major world matrix, columns right/up/forward/position
    void camera_matrix(double m[16]) const {
        if (cam_from_matrix) { for (int k = 0; k < 16; ++k) m[k] = cam_m[k]; return; }
        const double w = cam_q[0], x = cam_q[1], y = cam_q[2], z = cam_q[3];
        const double c0[3] = {1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)};
        const double c1[3] = {2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)};
        const double c2[3] = {2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)};
        for (int r = 0; r < 3; ++r) { m[r * 4 + 0] = c0[r]; m[r * 4 + 1] = c1[r]; m[r * 4 + 2] = c2[r]; }
        m[3] = cam_pos.x; m[7] = cam_pos.y; m[11] = cam_pos.z; m[12] = m[13] = m[14] = 0; m[15] = 1;
    }
    CameraBasis camera_basis() const {
        double m[16]; camera_matrix(m);
        const float half_y = 1.0f * tanf((fov_y * 0.01745329251994329576923690768489f) * 0.5f);
        const float half_x = half_y * ((float)st.width / (float)full_height());
        CameraBasis c;
        c.eye = cam_pos;
        c.U = f3((float)m[0], (float)m[4], (float)m[8]) * half_x;
        c.V = f3((float)m[1], (float)m[5], (float)m[9]) * half_y;
        c.W = f3((float)m[2], (float)m[6], (float)m[10]) * 1.0f;
        return c;
    }
    // projection * inverse(previous camera matrix): WaveFrontRenderer.cpp:760-781 + CPUShadingKernels.cu:27-54
    void prev_view_proj(float out[16]) const {
        double cur[16]; camera_matrix(cur);
        const double* c = have_prev_cam ? prev_cam : cur;
        double view[16];
        for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) view[r * 4 + k] = c[k * 4 + r]; view[r * 4 + 3] = -(c[0 * 4 + r] * c[3] + c[1 * 4 + r] * c[7] + c[2 * 4 + r] * c[11]); }
        view[12] = view[13] = view[14] = 0; view[15] = 1;
        const double aspect = (double)st.width / (double)full_height(), zn = 0.5, zf = 10000.0, th = tan((fov_y * 0.01745329251994329576923690768489) / 2.0);
        double P[16] = {1.0 / (aspect * th), 0, 0, 0, 0, 1.0 / th, 0, 0, 0, 0, -(zf + zn) / (zf - zn), -(2.0 * zf * zn) / (zf - zn), 0, 0, -1, 0};
        for (int r = 0; r < 4; ++r) for (int k = 0; k < 4; ++k) { double s = 0; for (int j = 0; j < 4; ++j) s += P[r * 4 + j] * view[j * 4 + k]; out[r * 4 + k] = (float)s; }
    }

    FrameView frame_view() {
        FrameView fv;
        fv.width = st.width; fv.height = st.height; fv.npix = npix();
        fv.row0 = st.band_row0; fv.full_height = full_height(); fv.pix0 = st.band_row0 * st.width;
        if (st.band_own_rows && vinstances.empty()) { fv.own_pix0 = (st.band_own_row0 - st.band_row0) * st.width; fv.own_pix1 = fv.own_pix0 + st.band_own_rows * st.width; }
        for (int q = 0; q < 2; ++q) fv.rays[q] = RayQueue{d_rays[q][0].p, d_rays[q][1].p, d_rays[q][2].p};
        fv.hits = d_hits.p; fv.primary_hits = d_primary_hits.p;
        fv.shadow = ShadowQueue{d_shadow[0].p, d_shadow[1].p, d_shadow[2].p};
        fv.surf_cur = d_surf[surf_cur].p; fv.surf_prev = d_surf[surf_cur ^ 1u].p;
        fv.res_cur = d_res[res_cur].p; fv.res_prev = d_res[res_cur ^ 1u].p; fv.res_tmp_a = d_res[2].p; fv.res_tmp_b = d_res[3].p;
        fv.channels = d_channels.p; fv.combined = d_combined.p; fv.accum = d_accum.p; fv.motion = d_motion.p; fv.ldr = d_ldr.p;
        fv.vol_hits = d_vol_hits.p; fv.vol_shadow = ShadowQueue{d_vol_shadow[0].p, d_vol_shadow[1].p, d_vol_shadow[2].p}; fv.counters = d_counters.p; fv.stats = d_stats.p;
        return fv;
    }

    void need_side_stream() {
        if (restir_stream) return;
        int lo = 0, hi = 0; LB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        static const bool high = []() { const char* e = getenv("LB_OVERLAP_PRIORITY"); return !e || atoi(e) != 0; }();
        LB_CUDA(cudaStreamCreateWithPriority(&restir_stream, cudaStreamNonBlocking, high ? hi : lo));
        LB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming)); LB_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        LB_CUDA(cudaEventCreateWithFlags(&ev_shadow, cudaEventDisableTiming));
    }
    void lap(const char* name, int chain = 0) {
        if (events_used == event_pool.size()) { cudaEvent_t e; LB_CUDA(cudaEventCreate(&e)); event_pool.push_back(e); }
        cudaEvent_t e = event_pool[events_used++];
        LB_CUDA(cudaEventRecord(e, chain ? restir_stream : stream));
        laps.push_back(Lap{name, e, last_lap[chain]});
        last_lap[chain] = laps.size() - 1;
    }

    // ---- WaveFrontRenderer::TraceFrame
    void render_frame() {
        if (scene_dirty || resources_dirty) commit_scene();
        else if (transforms_dirty) refit_scene();
        laps.clear(); events_used = 0; last_lap[0] = last_lap[1] = 0;
        lap("begin");
        const LaunchCfg c = cfg();
        FrameView fv = frame_view();
        const SceneView sc = scene_view();
        uint32_t* const ovf = d_counters.p + CNT_STACK_OVERFLOW;
        const BvhView bv = bvh.view(ovf), bva = dual_bvh ? bvh_any.view(ovf) : bvh.view(ovf);
        const uint32_t stride = st.frame_count_stride ? st.frame_count_stride : 2u;
        const uint32_t frame_count = st.first_frame_count + 1u + stride * frame_index;      // the reference's counter advances twice per frame
        uint32_t launches = 0;
        LB_CUDA(cudaMemsetAsync(d_counters.p, 0, kNumCounters * sizeof(uint32_t), stream));
        LB_CUDA(cudaMemsetAsync(d_stats.p, 0, kNumStats * sizeof(unsigned long long), stream));
        launch_raygen(c, fv, camera_basis(), frame_count); ++launches;
        lap("raygen");
        uint32_t seed = wang_hash(frame_count);
        uint32_t ticket = 0;
        // every trace launch of a frame owns one device ticket; a schedule that needs more than the counter block holds is refused here, not
        // left to index past d_counters
        auto take_ticket = [&]() { if (ticket >= kMaxTickets) throw std::runtime_error("frame schedule needs more than kMaxTickets device tickets"); return ticket++; };
        ShadeArgs a{}; a.max_depth = st.depth;
        a.volumes = d_volumes.p; a.num_volumes = (uint32_t)vinstances.size(); a.volume_mode = (int)st.volume_mode;
        prev_view_proj(a.prev_view_proj);
        bool forked = false, shadow_in_flight = false;
        // overlap bit 2: the ReSTIR chain is launched after the FIRST bounce wave (the only bounce wave that fills the machine); the later,
        // latency-bound waves (1e5 .. 1e4 rays) then run on the side stream beside it
        const bool tail_mode = (overlap & 4) && !(overlap & 2) && st.restir && st.depth >= 3u && sc.num_lights != 0u && a.num_volumes == 0u;
        bool tail_forked = false;
        const uint32_t seed0 = seed;
        auto run_restir = [&](bool on_side) {
            RestirArgs ra{seed0, (int)st.restir_temporal, (int)st.restir_spatial};
            static const int ris_simple = []() { const char* e = getenv("LB_RIS_SIMPLE"); return (!e || atoi(e) != 0) ? 1 : 0; }();
            ra.ris_simple = ris_simple; ra.unbiased = st.restir_unbiased ? 1 : 0;
            RestirBuffers rb{d_bags.p, d_ris_order.p, vis_sort ? d_vis_rays[0].p : nullptr, vis_sort ? d_vis_rays[1].p : nullptr, have_tmap ? &tmap_geom[surf_cur] : nullptr, d_spatial_nb.p};
            LaunchCfg cr = c;
            if (on_side) {
                need_side_stream();
                LB_CUDA(cudaEventRecord(ev_fork, stream)); LB_CUDA(cudaStreamWaitEvent(restir_stream, ev_fork, 0));
                cr.stream = restir_stream; last_lap[1] = last_lap[0];
                ra.lap = [](void* user, const char* stage) { static_cast<Renderer*>(user)->lap(stage, 1); };
            } else ra.lap = [](void* user, const char* stage) { static_cast<Renderer*>(user)->lap(stage, 0); };
            ra.lap_user = this;
            if (ticket + 6u > kMaxTickets) throw std::runtime_error("frame schedule needs more than kMaxTickets device tickets");
            launch_restir(cr, fv, sc, bva, rb, ra, ticket);
            if (on_side) LB_CUDA(cudaEventRecord(ev_join, restir_stream));
            if (sc.num_lights) launches += 4u + (st.restir_temporal ? 1u : 0u) + (st.restir_spatial ? 4u : 0u) + (vis_sort ? (st.restir_spatial ? 2u : 1u) : 0u);
        };
        for (uint32_t depth = 0; depth < st.depth; ++depth) {
            const int queue = (int)(depth & 1u);
            LaunchCfg cb = c; if (tail_forked) cb.stream = restir_stream;         // stream of the bounce chain
            const int chain = tail_forked ? 1 : 0;
            // overlap bit 3: from the third wave on, the rest of the bounce chain is ONE launch, a lane per path (lb_wavefront.cu k_tail)
            if (depth >= 2u && (overlap & 8) && a.num_volumes == 0u) {
                // wave (depth - 1)'s shadow rays on the side stream add into the same INDIRECT channel: they come first
                if (shadow_in_flight) { LB_CUDA(cudaStreamWaitEvent(cb.stream, ev_shadow, 0)); shadow_in_flight = false; }
                launch_tail(cb, fv, sc, bv, bva, queue, depth, st.depth, seed, 0.01f, 5000.f); ++launches;
                lap("tail", chain);
                break;
            }
            launch_extend(cb, fv, bv, queue, take_ticket(), depth == 0, 0.01f, 5000.f); ++launches;
            lap("extend", chain);
            // the previous wave's shadow rays (side stream) read the shadow queue this wave's shade kernel is about to refill
            if (shadow_in_flight) { LB_CUDA(cudaStreamWaitEvent(stream, ev_shadow, 0)); shadow_in_flight = false; }
            // the next wave's queue and this wave's shadow queues start empty
            LB_CUDA(cudaMemsetAsync(d_counters.p + (queue ? CNT_RAYS_A : CNT_RAYS_B), 0, sizeof(uint32_t), cb.stream));
            LB_CUDA(cudaMemsetAsync(d_counters.p + CNT_SHADOW, 0, sizeof(uint32_t), cb.stream));
            a.depth = depth; a.seed = seed;
            a.do_nee = (depth > 0 || !st.restir) ? 1 : 0;
            a.nee_channel = depth == 0 ? LB_CHANNEL_DIRECT : LB_CHANNEL_INDIRECT;
            a.do_bounce = depth + 1u < st.depth ? 1 : 0;
            if (a.num_volumes) {
                LB_CUDA(cudaMemsetAsync(d_counters.p + CNT_VOL_SHADOW, 0, sizeof(uint32_t), stream));
                launch_volume_extend(c, fv, queue, depth == 0, d_volumes.p, a.num_volumes, 0.01f, 5000.f); ++launches;
                if (st.volume_mode == LB_VOLUME_DELTA) { launch_volume_delta(c, fv, sc, queue, depth == 0, a); ++launches; }
                lap("volume");
            }
            launch_shade(cb, fv, sc, queue, a); ++launches;
            lap("shade", chain);
            if (depth == 0 && st.restir && !tail_mode) {
                forked = (overlap & 2) && st.depth > 1 && sc.num_lights != 0u && a.num_volumes == 0u;      // media: volume shadow rays also write DIRECT at depth 0
                run_restir(forked);
            }
            if (depth == 1u && tail_mode) {
                // fork: the rest of the bounce chain continues on the side stream, the ReSTIR chain takes the main stream
                need_side_stream();
                LB_CUDA(cudaEventRecord(ev_fork, stream)); LB_CUDA(cudaStreamWaitEvent(restir_stream, ev_fork, 0));
                last_lap[1] = last_lap[0]; tail_forked = true; cb.stream = restir_stream;
                run_restir(false);
            }
            const int chain_s = tail_forked ? 1 : 0;
            if (a.do_nee || (a.num_volumes && st.volume_mode == LB_VOLUME_DELTA)) {
                // bounce waves: this launch and the next wave's extend are both small and latency-bound (each lasts as long as its slowest
                // ray) and touch disjoint buffers — the shadow rays go to the side stream and are joined before the next shade
                const bool side = (overlap & 1) && !forked && !tail_forked && depth >= 1u && depth + 1u < st.depth && a.num_volumes == 0u;
                if (side) {
                    need_side_stream();
                    LB_CUDA(cudaEventRecord(ev_fork, stream)); LB_CUDA(cudaStreamWaitEvent(restir_stream, ev_fork, 0));
                    LaunchCfg cs = c; cs.stream = restir_stream; last_lap[1] = last_lap[0];
                    launch_shadow(cs, fv, bva, take_ticket(), 0.01f); ++launches; lap("shadow", 1);
                    LB_CUDA(cudaEventRecord(ev_shadow, restir_stream)); shadow_in_flight = true;
                } else { launch_shadow(cb, fv, bva, take_ticket(), 0.01f); ++launches; lap("shadow", chain_s); }
            }
            if (a.do_nee && a.num_volumes && st.volume_mode == LB_VOLUME_COMPAT) { launch_volume_shadow(c, fv, bva, take_ticket(), 0.01f); ++launches; lap("volume_shadow"); }
            seed = wang_hash(seed);
        }
        if (tail_forked) { LB_CUDA(cudaEventRecord(ev_join, restir_stream)); forked = true; }
        if (shadow_in_flight) { LB_CUDA(cudaStreamWaitEvent(stream, ev_shadow, 0)); shadow_in_flight = false; }
        if (forked) { LB_CUDA(cudaStreamWaitEvent(stream, ev_join, 0)); lap("restir_join"); }    // the time the bounce chain waited for the ReSTIR chain
        if (copy_pending) LB_CUDA(cudaStreamWaitEvent(stream, ev_copied, 0));      // an asynchronous read-back still owns the combined buffer
        if (reduce_pending) { LB_CUDA(cudaStreamWaitEvent(stream, ev_reduced, 0)); reduce_pending = false; }      // a reduce still reads the accumulation buffer
        launch_merge(c, fv, (int)st.blend_output, blend_count); ++launches;
        lap("merge");
        if (st.blend_output) ++blend_count; else blend_count = 1;
        double m[16]; camera_matrix(m); memcpy(prev_cam, m, sizeof m); have_prev_cam = true;      // Camera::UpdatePreviousFrameMatrix
        surf_cur ^= 1u; res_cur ^= 1u;                                                              // ReSTIR::SwapBuffers once per frame (SURVEY hazard 13)
        ++frame_index;
        launches_last_frame = launches;
    }
};

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

template <class F>
static int guarded(Renderer* r, F&& f) {
    if (!r) return fail(LB_ERR_INVALID_ARGUMENT, "null renderer");
    try {
        std::lock_guard<std::mutex> lock(r->mu);
        LB_CUDA(cudaSetDevice(r->device));
        return f();
    } catch (const CudaError& e) { return fail(LB_ERR_CUDA, e.what()); }
    catch (const std::bad_alloc&) { return fail(LB_ERR_OUT_OF_MEMORY, "out of host memory"); }
    catch (const std::exception& e) { return fail(LB_ERR_STATE, e.what()); }
}

} // namespace lb

using namespace lb;
#define R_ (reinterpret_cast<lb::Renderer*>(r))

extern "C" {

LB_API int lb_create(const LbSettings* s, LbRenderer* out) {
    if (!s || !out || !s->width || !s->height || !s->depth) return fail(LB_ERR_INVALID_ARGUMENT, "bad settings");
    if (s->depth > 24) return fail(LB_ERR_INVALID_ARGUMENT, "depth > 24 is not supported");
    if (s->band_full_height && (s->band_row0 + s->height > s->band_full_height || ((uint64_t)s->band_row0 * s->width) % 256u))
        return fail(LB_ERR_INVALID_ARGUMENT, "row band: band_row0 + height must fit band_full_height and band_row0 * width must be a multiple of 256");
    if (!s->band_full_height && s->band_row0) return fail(LB_ERR_INVALID_ARGUMENT, "band_row0 without band_full_height");
    if (s->band_own_rows && (s->band_own_row0 < s->band_row0 || s->band_own_row0 + s->band_own_rows > s->band_row0 + s->height))
        return fail(LB_ERR_INVALID_ARGUMENT, "band_own_row0 / band_own_rows must lie inside the rendered rows");
    std::unique_ptr<lb::Renderer> r(new lb::Renderer());
    r->st = *s;
    try { r->init(); }
    catch (const CudaError& e) { return fail(LB_ERR_CUDA, e.what()); }
    catch (const std::exception& e) { return fail(LB_ERR_STATE, e.what()); }
    *out = reinterpret_cast<LbRenderer>(r.release());
    return LB_OK;
}
LB_API int lb_destroy(LbRenderer r) { if (!r) return fail(LB_ERR_INVALID_ARGUMENT, "null renderer"); delete R_; return LB_OK; }
LB_API const char* lb_last_error(void) { return g_err.c_str(); }
LB_API const char* lb_version(void) { return "lumen-b200 0.1 (sm_100a)"; }

LB_API int lb_texture_create(LbRenderer r, const uint8_t* rgba8, uint32_t w, uint32_t h, int srgb, LbHandle* out) {
    return guarded(R_, [&]() {
        if (!rgba8 || !w || !h || !out) return fail(LB_ERR_INVALID_ARGUMENT, "bad texture");
        HostTexture t; t.w = w; t.h = h; t.srgb = srgb != 0; t.px.assign(rgba8, rgba8 + (size_t)w * h * 4);
        R_->textures.push_back(std::move(t)); R_->resources_dirty = true; *out = (LbHandle)R_->textures.size() - 1; return (int)LB_OK;
    });
}
LB_API int lb_material_create(LbRenderer r, const LbMaterialDesc* d, LbHandle* out) {
    return guarded(R_, [&]() {
        if (!d || !out) return fail(LB_ERR_INVALID_ARGUMENT, "null");
        if (!(d->roughness_factor > 0.f && d->roughness_factor <= 1.f)) return fail(LB_ERR_INVALID_ARGUMENT, "roughness must be in (0,1]");   // WaveFrontRenderer.cpp:1274-1284
        HostMaterial m; if (!R_->fill_material(m, *d)) return fail(LB_ERR_INVALID_HANDLE, "texture handle");
        R_->materials.push_back(m); R_->resources_dirty = true; *out = (LbHandle)R_->materials.size() - 1; return (int)LB_OK;
    });
}
LB_API int lb_material_update(LbRenderer r, LbHandle h, const LbMaterialDesc* d) {
    return guarded(R_, [&]() {
        if (h < 0 || h >= (LbHandle)R_->materials.size() || !d) return fail(LB_ERR_INVALID_HANDLE, "material");
        if (!(d->roughness_factor > 0.f && d->roughness_factor <= 1.f)) return fail(LB_ERR_INVALID_ARGUMENT, "roughness must be in (0,1]");
        if (!R_->fill_material(R_->materials[h], *d)) return fail(LB_ERR_INVALID_HANDLE, "texture handle");
        R_->resources_dirty = true; R_->scene_dirty = true; return (int)LB_OK;
    });
}
LB_API int lb_primitive_create(LbRenderer r, const LbPrimitiveDesc* d, LbHandle* out) {
    return guarded(R_, [&]() {
        if (!d || !out || !d->positions || !d->indices || !d->vertex_count || d->index_count % 3) return fail(LB_ERR_INVALID_ARGUMENT, "bad primitive");
        if (d->index_size != 2 && d->index_size != 4) return fail(LB_ERR_INVALID_ARGUMENT, "index size");
        if (d->material < 0 || d->material >= (LbHandle)R_->materials.size()) return fail(LB_ERR_INVALID_HANDLE, "material");
        HostPrimitive p; p.material = d->material; const uint32_t n = d->vertex_count;
        p.pos.resize(n); p.nu.assign(n, make_float4(0, 0, 1, 0)); p.tv.assign(n, make_float4(1, 0, 0, 0)); p.tw.assign(n, 1.f);
        auto at = [](const void* base, uint32_t stride, uint32_t i) { return (const float*)((const char*)base + (size_t)stride * i); };
        for (uint32_t i = 0; i < n; ++i) {
            const float* q = at(d->positions, d->position_stride ? d->position_stride : 12, i); p.pos[i] = make_float4(q[0], q[1], q[2], 0.f);
            if (d->uvs) { q = at(d->uvs, d->uv_stride ? d->uv_stride : 8, i); p.nu[i].w = q[0]; p.tv[i].w = q[1]; }
            if (d->normals) { q = at(d->normals, d->normal_stride ? d->normal_stride : 12, i); p.nu[i].x = q[0]; p.nu[i].y = q[1]; p.nu[i].z = q[2]; }
            if (d->tangents) { q = at(d->tangents, d->tangent_stride ? d->tangent_stride : 16, i); p.tv[i].x = q[0]; p.tv[i].y = q[1]; p.tv[i].z = q[2]; p.tw[i] = q[3]; }
        }
        p.idx.resize(d->index_count);
        for (uint32_t i = 0; i < d->index_count; ++i) {
            p.idx[i] = d->index_size == 2 ? ((const uint16_t*)d->indices)[i] : ((const uint32_t*)d->indices)[i];
            if (p.idx[i] >= n) return fail(LB_ERR_INVALID_ARGUMENT, "index out of range");
        }
        R_->prims.push_back(std::move(p)); R_->resources_dirty = true; *out = (LbHandle)R_->prims.size() - 1; return (int)LB_OK;
    });
}
LB_API int lb_mesh_create(LbRenderer r, const LbHandle* prims, uint32_t count, LbHandle* out) {
    return guarded(R_, [&]() {
        if (!prims || !count || !out) return fail(LB_ERR_INVALID_ARGUMENT, "bad mesh");
        HostMesh m;
        for (uint32_t i = 0; i < count; ++i) { if (prims[i] < 0 || prims[i] >= (LbHandle)R_->prims.size()) return fail(LB_ERR_INVALID_HANDLE, "primitive"); m.prims.push_back(prims[i]); }
        R_->meshes.push_back(m); *out = (LbHandle)R_->meshes.size() - 1; return (int)LB_OK;
    });
}
LB_API int lb_volume_create(LbRenderer r, const LbVolumeDesc* d, LbHandle* out) {
    return guarded(R_, [&]() {
        if (!d || !out) return fail(LB_ERR_INVALID_ARGUMENT, "null");
        HostVolume v; v.nx = d->nx; v.ny = d->ny; v.nz = d->nz; v.lo = f3(d->bbox_min[0], d->bbox_min[1], d->bbox_min[2]); v.hi = f3(d->bbox_max[0], d->bbox_max[1], d->bbox_max[2]);
        if (d->density) {
            if (!d->nx || !d->ny || !d->nz) return fail(LB_ERR_INVALID_ARGUMENT, "grid size");
            v.density.assign(d->density, d->density + (size_t)d->nx * d->ny * d->nz); v.majorant = 0.f; for (float x : v.density) v.majorant = fmaxf(v.majorant, x);
        }
        R_->volumes.push_back(std::move(v)); *out = (LbHandle)R_->volumes.size() - 1; return (int)LB_OK;
    });
}
static const float kIdentity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
LB_API int lb_scene_add_mesh_instance(LbRenderer r, LbHandle mesh, const float* m16, const LbEmissiveness* em, LbHandle ov, LbHandle* out) {
    return guarded(R_, [&]() {
        if (mesh < 0 || mesh >= (LbHandle)R_->meshes.size()) return fail(LB_ERR_INVALID_HANDLE, "mesh");
        if (ov >= (LbHandle)R_->materials.size()) return fail(LB_ERR_INVALID_HANDLE, "material");
        HostInstance in{}; in.mesh = mesh; in.override_mat = ov;
        memcpy(in.m, m16 ? m16 : kIdentity, sizeof in.m);
        in.em = em ? *em : LbEmissiveness{LB_EMISSION_ENABLED, {0, 0, 0}, 1.f};
        R_->instances.push_back(in); R_->scene_dirty = true; if (out) *out = (LbHandle)R_->instances.size() - 1; return (int)LB_OK;
    });
}
LB_API int lb_instance_set_transform(LbRenderer r, LbHandle i, const float* m16) {
    return guarded(R_, [&]() { if (i < 0 || i >= (LbHandle)R_->instances.size() || !m16) return fail(LB_ERR_INVALID_HANDLE, "instance"); memcpy(R_->instances[i].m, m16, 64); R_->transforms_dirty = true; return (int)LB_OK; });
}
LB_API int lb_instance_set_emissiveness(LbRenderer r, LbHandle i, const LbEmissiveness* em) {
    return guarded(R_, [&]() { if (i < 0 || i >= (LbHandle)R_->instances.size() || !em) return fail(LB_ERR_INVALID_HANDLE, "instance"); R_->instances[i].em = *em; R_->scene_dirty = true; return (int)LB_OK; });
}
LB_API int lb_instance_set_override_material(LbRenderer r, LbHandle i, LbHandle m) {
    return guarded(R_, [&]() { if (i < 0 || i >= (LbHandle)R_->instances.size() || m >= (LbHandle)R_->materials.size()) return fail(LB_ERR_INVALID_HANDLE, "instance"); R_->instances[i].override_mat = m; R_->scene_dirty = true; return (int)LB_OK; });
}
LB_API int lb_scene_add_volume_instance(LbRenderer r, LbHandle vol, const float* m16, float density, LbHandle* out) {
    return guarded(R_, [&]() {
        if (vol < 0 || vol >= (LbHandle)R_->volumes.size()) return fail(LB_ERR_INVALID_HANDLE, "volume");
        HostVolumeInstance vi{}; vi.volume = vol; vi.density = density;
        memcpy(vi.m, m16 ? m16 : kIdentity, 64); invert_affine(vi.m, vi.inv);
        R_->vinstances.push_back(vi); R_->scene_dirty = true; if (out) *out = (LbHandle)R_->vinstances.size() - 1; return (int)LB_OK;
    });
}
LB_API int lb_scene_clear(LbRenderer r) { return guarded(R_, [&]() { R_->instances.clear(); R_->vinstances.clear(); R_->scene_dirty = true; return (int)LB_OK; }); }
LB_API int lb_camera_set_pose(LbRenderer r, const float* p, const float* q) {
    return guarded(R_, [&]() { if (!p || !q) return fail(LB_ERR_INVALID_ARGUMENT, "null"); R_->cam_pos = f3(p[0], p[1], p[2]); memcpy(R_->cam_q, q, 16); R_->cam_from_matrix = false; return (int)LB_OK; });
}
LB_API int lb_camera_set_matrix(LbRenderer r, const float* m) {
    return guarded(R_, [&]() { if (!m) return fail(LB_ERR_INVALID_ARGUMENT, "null"); memcpy(R_->cam_m, m, 64); R_->cam_pos = f3(m[3], m[7], m[11]); R_->cam_from_matrix = true; return (int)LB_OK; });
}
LB_API int lb_camera_set_fov_y(LbRenderer r, float deg) {
    return guarded(R_, [&]() { if (!(deg > 0.f && deg < 180.f)) return fail(LB_ERR_INVALID_ARGUMENT, "fov"); R_->fov_y = deg; return (int)LB_OK; });
}
LB_API int lb_camera_set_min_max_distance(LbRenderer r, float mn, float mx) {
    return guarded(R_, [&]() { if (!(mx > mn)) return fail(LB_ERR_INVALID_ARGUMENT, "min/max distance"); R_->cam_min_d = mn; R_->cam_max_d = mx; return (int)LB_OK; });
}
LB_API int lb_set_render_resolution(LbRenderer r, uint32_t w, uint32_t h) {
    return guarded(R_, [&]() { if (!w || !h) return fail(LB_ERR_INVALID_ARGUMENT, "resolution"); LB_CUDA(cudaStreamSynchronize(R_->stream)); R_->st.width = w; R_->st.height = h; R_->resize(); return (int)LB_OK; });
}
LB_API int lb_get_settings(LbRenderer r, LbSettings* out) { return guarded(R_, [&]() { if (!out) return fail(LB_ERR_INVALID_ARGUMENT, "null"); *out = R_->st; return (int)LB_OK; }); }
LB_API int lb_get_render_resolution(LbRenderer r, uint32_t* w, uint32_t* h) { return guarded(R_, [&]() { *w = R_->st.width; *h = R_->st.height; return (int)LB_OK; }); }
LB_API int lb_set_depth(LbRenderer r, uint32_t d) { return guarded(R_, [&]() { if (!d || d > 24) return fail(LB_ERR_INVALID_ARGUMENT, "depth"); R_->st.depth = d; return (int)LB_OK; }); }
LB_API int lb_set_blend_mode(LbRenderer r, int b) {
    // the first frame after this call overwrites the accumulation buffer (k_merge with blend_count == 0): no memset here, which would have to
    // wait for a reduce that may still be reading the buffer (lb_reduce_begin)
    return guarded(R_, [&]() { R_->st.blend_output = b != 0; R_->blend_count = 0; return (int)LB_OK; });
}
LB_API int lb_get_blend_mode(LbRenderer r, int* b) { return guarded(R_, [&]() { *b = (int)R_->st.blend_output; return (int)LB_OK; }); }
LB_API int lb_reset_history(LbRenderer r) { return guarded(R_, [&]() { LB_CUDA(cudaStreamSynchronize(R_->stream)); R_->resize(); return (int)LB_OK; }); }
LB_API int lb_render_frames(LbRenderer r, uint32_t frames) {
    return guarded(R_, [&]() { for (uint32_t i = 0; i < frames; ++i) R_->render_frame(); return (int)LB_OK; });
}
LB_API int lb_synchronize(LbRenderer r) {
    return guarded(R_, [&]() { LB_CUDA(cudaStreamSynchronize(R_->stream)); if (R_->reduce_stream) LB_CUDA(cudaStreamSynchronize(R_->reduce_stream)); return (int)LB_OK; });
}
LB_API int lb_start_rendering(LbRenderer r) {
    if (!r) return fail(LB_ERR_INVALID_ARGUMENT, "null renderer");
    lb::Renderer* R = R_;
    if (R->render_thread.joinable()) return fail(LB_ERR_STATE, "already rendering");
    R->stop_flag = false;
    R->render_thread = std::thread([R]() {
        while (!R->stop_flag) {
            try {
                std::lock_guard<std::mutex> lock(R->mu);
                LB_CUDA(cudaSetDevice(R->device));
                R->render_frame();
                LB_CUDA(cudaStreamSynchronize(R->stream));
            } catch (const std::exception& e) { R->thread_error = e.what(); break; }
            std::this_thread::yield();
        }
    });
    return LB_OK;
}
LB_API int lb_stop_rendering(LbRenderer r) {
    if (!r) return fail(LB_ERR_INVALID_ARGUMENT, "null renderer");
    R_->stop_thread();
    if (!R_->thread_error.empty()) { const std::string e = R_->thread_error; R_->thread_error.clear(); return fail(LB_ERR_CUDA, e); }
    return LB_OK;
}

static int read_back(lb::Renderer* R, const void* src, size_t bytes, void* dst, size_t cap) {
    if (!dst || cap < bytes) return fail(LB_ERR_INVALID_ARGUMENT, "buffer too small");
    if (R->reduce_pending) LB_CUDA(cudaStreamWaitEvent(R->stream, R->ev_reduced, 0));          // the image may be the one a reduce is resolving
    LB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, R->stream));
    LB_CUDA(cudaStreamSynchronize(R->stream));
    return LB_OK;
}
LB_API int lb_read_hdr(LbRenderer r, float* out, size_t cap) { return guarded(R_, [&]() { return read_back(R_, R_->d_combined.p, (size_t)R_->npix() * 16, out, cap); }); }
LB_API int lb_read_hdr_async(LbRenderer r, float* out, size_t cap) {
    return guarded(R_, [&]() {
        const size_t bytes = (size_t)R_->npix() * 16;
        if (!out || cap < bytes) return fail(LB_ERR_INVALID_ARGUMENT, "buffer too small");
        LB_CUDA(cudaEventRecord(R_->ev_rendered, R_->stream));
        LB_CUDA(cudaStreamWaitEvent(R_->copy_stream, R_->ev_rendered, 0));
        if (R_->reduce_pending) LB_CUDA(cudaStreamWaitEvent(R_->copy_stream, R_->ev_reduced, 0));      // the copy engine, not the render stream, waits for the resolve
        LB_CUDA(cudaMemcpyAsync(out, R_->d_combined.p, bytes, cudaMemcpyDeviceToHost, R_->copy_stream));
        LB_CUDA(cudaEventRecord(R_->ev_copied, R_->copy_stream));
        R_->copy_pending = true;
        return (int)LB_OK;
    });
}
LB_API int lb_readback_wait(LbRenderer r) {
    return guarded(R_, [&]() { if (R_->copy_pending) { LB_CUDA(cudaEventSynchronize(R_->ev_copied)); R_->copy_pending = false; } return (int)LB_OK; });
}
LB_API int lb_read_gbuffer(LbRenderer r, float* depth, float* normal_rough, float* albedo, size_t pixel_capacity) {
    return guarded(R_, [&]() {
        const size_t n = R_->npix();
        if (pixel_capacity < n) return fail(LB_ERR_INVALID_ARGUMENT, "buffer too small");
        R_->d_gbuffer.reserve(n * 9);                                                   // depth | normal+roughness (4) | albedo (4)
        float* d_depth = R_->d_gbuffer.p; float4* d_nr = reinterpret_cast<float4*>(R_->d_gbuffer.p + n); float4* d_al = d_nr + n;
        launch_gbuffer(R_->cfg(), R_->d_surf[R_->surf_cur ^ 1u].p, (uint32_t)n, R_->cam_min_d, R_->cam_max_d, depth ? d_depth : nullptr, normal_rough ? d_nr : nullptr, albedo ? d_al : nullptr);
        if (depth) LB_CUDA(cudaMemcpyAsync(depth, d_depth, n * 4, cudaMemcpyDeviceToHost, R_->stream));
        if (normal_rough) LB_CUDA(cudaMemcpyAsync(normal_rough, d_nr, n * 16, cudaMemcpyDeviceToHost, R_->stream));
        if (albedo) LB_CUDA(cudaMemcpyAsync(albedo, d_al, n * 16, cudaMemcpyDeviceToHost, R_->stream));
        LB_CUDA(cudaStreamSynchronize(R_->stream));
        return (int)LB_OK;
    });
}
LB_API int lb_read_ldr(LbRenderer r, uint8_t* out, size_t cap) { return guarded(R_, [&]() { return read_back(R_, R_->d_ldr.p, (size_t)R_->npix() * 4, out, cap); }); }
LB_API int lb_read_channel(LbRenderer r, int c, float* out, size_t cap) {
    return guarded(R_, [&]() { if (c < 0 || c >= LB_NUM_CHANNELS) return fail(LB_ERR_INVALID_ARGUMENT, "channel"); return read_back(R_, R_->d_channels.p + (size_t)c * R_->npix(), (size_t)R_->npix() * 16, out, cap); });
}
LB_API int lb_read_motion_vectors(LbRenderer r, float* out, size_t cap) { return guarded(R_, [&]() { return read_back(R_, R_->d_motion.p, (size_t)R_->npix() * 8, out, cap); }); }
LB_API int lb_frame_stats(LbRenderer r, const char** names, float* micros, uint32_t cap, uint32_t* count) {
    return guarded(R_, [&]() {
        LB_CUDA(cudaStreamSynchronize(R_->stream));
        R_->stats_names.clear(); uint32_t n = 0;
        for (size_t i = 1; i < R_->laps.size() && n < cap; ++i) {
            float ms = 0.f; LB_CUDA(cudaEventElapsedTime(&ms, R_->laps[R_->laps[i].prev].ev, R_->laps[i].ev));
            if (n) R_->stats_names += ';';
            R_->stats_names += R_->laps[i].name; if (micros) micros[n] = ms * 1000.f; ++n;
        }
        if (names) *names = R_->stats_names.c_str();
        if (count) *count = n;
        return (int)LB_OK;
    });
}
static int frame_counters_locked(lb::Renderer* R, uint64_t* v, uint32_t cap, uint32_t* count) {     // caller holds R->mu
#ifdef LB_RIS_STATS
    cudaStreamSynchronize(R->stream); lb::dump_ris_stats();
#endif
#ifdef LB_TRACE_STATS
    cudaDeviceSynchronize(); lb::dump_trace_stats_wavefront(); lb::dump_trace_stats_restir();
#endif
    unsigned long long s[kNumStats]; uint32_t overflows = 0;
    LB_CUDA(cudaMemcpyAsync(s, R->d_stats.p, sizeof s, cudaMemcpyDeviceToHost, R->stream));
    LB_CUDA(cudaMemcpyAsync(&overflows, R->d_counters.p + CNT_STACK_OVERFLOW, sizeof overflows, cudaMemcpyDeviceToHost, R->stream));
    LB_CUDA(cudaStreamSynchronize(R->stream));
    R->counters[11] = overflows;              // traversal-stack overflows since the last frame began (debug traces included): must be 0
    R->counters[0] = s[STAT_EXTEND]; R->counters[1] = s[STAT_SHADOW]; R->counters[2] = s[STAT_VIS]; R->counters[3] = R->launches_last_frame;
    const uint32_t n = cap < 15 ? cap : 15; memcpy(v, R->counters, n * 8); if (count) *count = n; return (int)LB_OK;
}
LB_API int lb_frame_counters(LbRenderer r, uint64_t* v, uint32_t cap, uint32_t* count) {
    return guarded(R_, [&]() { return frame_counters_locked(R_, v, cap, count); });
}
// ---- output stage: screenshot + FrameStats export
LB_API int lb_save_png(LbRenderer r, const char* path) {
    return guarded(R_, [&]() {
        if (!path || !*path) return fail(LB_ERR_INVALID_ARGUMENT, "path");
        std::vector<uint8_t> px((size_t)R_->npix() * 4);
        const int rc = read_back(R_, R_->d_ldr.p, px.size(), px.data(), px.size());
        if (rc) return rc;
        if (!png::write_file(path, png::encode_rgba8(px.data(), R_->st.width, R_->st.height))) return fail(LB_ERR_INVALID_ARGUMENT, std::string("cannot write ") + path);
        return (int)LB_OK;
    });
}
LB_API int lb_frame_stats_json(LbRenderer r, char* json, size_t cap, size_t* needed) {
    return guarded(R_, [&]() {
        static const char* kCounterNames[14] = {"extend_rays", "shadow_rays", "visibility_rays", "kernel_launches", "lights", "triangles", "bvh_nodes",
                                                "bvh_bytes", "bvh_build_us", "bvh_levels", "bvh_build_rounds", "stack_overflows", "bvh_refit_us", "bvh_refits"};
        uint64_t cnt[14]; uint32_t n = 0;
        int rc = frame_counters_locked(R_, cnt, 14, &n); if (rc) return rc;
        LB_CUDA(cudaStreamSynchronize(R_->stream));
        // stages that run several times per frame (extend, shade, shadow per wave) are summed, as FrameStats::m_Times does with its map
        std::vector<std::pair<std::string, double>> times;
        for (size_t i = 1; i < R_->laps.size(); ++i) {
            float ms = 0.f; LB_CUDA(cudaEventElapsedTime(&ms, R_->laps[R_->laps[i].prev].ev, R_->laps[i].ev));
            auto it = std::find_if(times.begin(), times.end(), [&](const std::pair<std::string, double>& t) { return t.first == R_->laps[i].name; });
            if (it == times.end()) times.emplace_back(R_->laps[i].name, (double)ms * 1000.0); else it->second += (double)ms * 1000.0;
        }
        char num[64];
        std::string o = "{\"frame_id\": " + std::to_string(R_->frame_index) + ", \"resolution\": [" + std::to_string(R_->st.width) + ", " + std::to_string(R_->st.height) + "], \"times_us\": {";
        for (size_t i = 0; i < times.size(); ++i) { snprintf(num, sizeof num, "%.3f", times[i].second); o += (i ? ", \"" : "\"") + times[i].first + "\": " + num; }
        o += "}, \"counters\": {";
        for (uint32_t i = 0; i < n && i < 14; ++i) o += std::string(i ? ", \"" : "\"") + kCounterNames[i] + "\": " + std::to_string(cnt[i]);
        o += "}}";
        if (needed) *needed = o.size() + 1;
        if (!json || cap < o.size() + 1) return (json || cap) ? fail(LB_ERR_INVALID_ARGUMENT, "buffer too small") : (needed ? (int)LB_OK : fail(LB_ERR_INVALID_ARGUMENT, "null"));
        memcpy(json, o.c_str(), o.size() + 1);
        return (int)LB_OK;
    });
}
LB_API int lb_hdr_buffer(LbRenderer r, void** p, size_t* bytes) {
    return guarded(R_, [&]() { if (!p || !bytes) return fail(LB_ERR_INVALID_ARGUMENT, "null"); *p = R_->d_combined.p; *bytes = (size_t)R_->npix() * 16; return (int)LB_OK; });
}
LB_API int lb_accum_buffer(LbRenderer r, void** p, size_t* bytes, uint32_t* frames) {
    return guarded(R_, [&]() { *p = R_->d_accum.p; *bytes = (size_t)R_->npix() * 16; *frames = R_->blend_count; return (int)LB_OK; });
}
LB_API int lb_resolve_accum(LbRenderer r, uint32_t total) {
    return guarded(R_, [&]() { if (!total) return fail(LB_ERR_INVALID_ARGUMENT, "frames"); FrameView fv = R_->frame_view(); if (R_->copy_pending) LB_CUDA(cudaStreamWaitEvent(R_->stream, R_->ev_copied, 0)); launch_resolve(R_->cfg(), fv, 1.0f / (float)total); return (int)LB_OK; });
}
// ---- hooks of the multi-GPU layer (csrc/lb_multigpu.cpp): a collective over the accumulation buffers that overlaps the next frame
LB_API int lb_reduce_begin(LbRenderer r, void** side_stream, void** send, void** recv_root, size_t* bytes) {
    return guarded(R_, [&]() {
        if (!side_stream || !send || !recv_root || !bytes) return fail(LB_ERR_INVALID_ARGUMENT, "null");
        if (!R_->reduce_stream) {
            LB_CUDA(cudaStreamCreateWithFlags(&R_->reduce_stream, cudaStreamNonBlocking));
            LB_CUDA(cudaEventCreateWithFlags(&R_->ev_frames, cudaEventDisableTiming)); LB_CUDA(cudaEventCreateWithFlags(&R_->ev_reduced, cudaEventDisableTiming));
        }
        R_->d_reduced.reserve(R_->npix());
        LB_CUDA(cudaEventRecord(R_->ev_frames, R_->stream)); LB_CUDA(cudaStreamWaitEvent(R_->reduce_stream, R_->ev_frames, 0));
        if (R_->copy_pending) LB_CUDA(cudaStreamWaitEvent(R_->reduce_stream, R_->ev_copied, 0));      // the resolve will overwrite the image a read-back may still be copying
        *side_stream = (void*)R_->reduce_stream; *send = R_->d_accum.p; *recv_root = R_->d_reduced.p; *bytes = (size_t)R_->npix() * 16;
        return (int)LB_OK;
    });
}
LB_API int lb_reduce_end(LbRenderer r, int is_root, uint32_t total_frames) {
    return guarded(R_, [&]() {
        if (!R_->reduce_stream) return fail(LB_ERR_STATE, "lb_reduce_begin first");
        if (is_root) {
            if (!total_frames) return fail(LB_ERR_INVALID_ARGUMENT, "frames");
            FrameView fv = R_->frame_view(); fv.accum = R_->d_reduced.p;
            LaunchCfg c = R_->cfg(); c.stream = R_->reduce_stream;
            launch_resolve(c, fv, 1.0f / (float)total_frames);
        }
        LB_CUDA(cudaEventRecord(R_->ev_reduced, R_->reduce_stream));
        R_->reduce_pending = true;
        return (int)LB_OK;
    });
}
LB_API int lb_reduce_wait(LbRenderer r) {
    return guarded(R_, [&]() { if (R_->reduce_pending) { LB_CUDA(cudaStreamWaitEvent(R_->stream, R_->ev_reduced, 0)); R_->reduce_pending = false; } return (int)LB_OK; });
}
LB_API int lb_set_overlap(LbRenderer r, int enabled) {
    return guarded(R_, [&]() { if (enabled < 0 || enabled > 15) return fail(LB_ERR_INVALID_ARGUMENT, "overlap mode"); LB_CUDA(cudaStreamSynchronize(R_->stream)); R_->overlap = enabled; return (int)LB_OK; });
}
LB_API int lb_get_stream(LbRenderer r, void** s) { return guarded(R_, [&]() { if (!s) return fail(LB_ERR_INVALID_ARGUMENT, "null"); *s = (void*)R_->stream; return (int)LB_OK; }); }
LB_API int lb_set_stream(LbRenderer r, void* s) {
    return guarded(R_, [&]() { LB_CUDA(cudaStreamSynchronize(R_->stream)); R_->stream = s ? (cudaStream_t)s : R_->own_stream; return (int)LB_OK; });
}

// ---- debug taps
LB_API int lb_debug_trace_closest(LbRenderer r, const float* rays6, uint32_t n, float tmin, float tmax, void* hits20) {
    return guarded(R_, [&]() {
        if (R_->scene_dirty || R_->resources_dirty) R_->commit_scene(); else if (R_->transforms_dirty) R_->refit_scene();
        if (!n) return (int)LB_OK;
        DevBuf<float> d_rays; DevBuf<unsigned char> d_hits;
        d_rays.upload(rays6, (size_t)n * 6, R_->stream); d_hits.reserve((size_t)n * 20);
        launch_debug_trace(R_->cfg(), R_->bvh.view(R_->d_counters.p + CNT_STACK_OVERFLOW), d_rays.p, nullptr, n, tmin, tmax, d_hits.p, nullptr);
        return read_back(R_, d_hits.p, (size_t)n * 20, hits20, (size_t)n * 20);
    });
}
LB_API int lb_debug_trace_any(LbRenderer r, const float* rays6, const float* tmaxs, uint32_t n, float tmin, uint8_t* occ) {
    return guarded(R_, [&]() {
        if (R_->scene_dirty || R_->resources_dirty) R_->commit_scene(); else if (R_->transforms_dirty) R_->refit_scene();
        if (!n) return (int)LB_OK;
        DevBuf<float> d_rays, d_tmax; DevBuf<uint8_t> d_occ;
        d_rays.upload(rays6, (size_t)n * 6, R_->stream); d_tmax.upload(tmaxs, n, R_->stream); d_occ.reserve(n);
        launch_debug_trace(R_->cfg(), R_->dual_bvh ? R_->bvh_any.view(R_->d_counters.p + CNT_STACK_OVERFLOW) : R_->bvh.view(R_->d_counters.p + CNT_STACK_OVERFLOW), d_rays.p, d_tmax.p, n, tmin, 0.f, nullptr, d_occ.p);
        return read_back(R_, d_occ.p, n, occ, n);
    });
}
LB_API int lb_debug_read_lights(LbRenderer r, float* l16, float* cdf, uint32_t cap, uint32_t* count) {
    return guarded(R_, [&]() {
        if (R_->scene_dirty || R_->resources_dirty) R_->commit_scene(); else if (R_->transforms_dirty) R_->refit_scene();
        const uint32_t n = R_->lights.num_lights; if (count) *count = n;
        if (cap < n) return fail(LB_ERR_INVALID_ARGUMENT, "capacity");
        if (n && l16) { const int rc = read_back(R_, R_->lights.lights.p, (size_t)n * 64, l16, (size_t)cap * 64); if (rc) return rc; }
        if (n && cdf) { const int rc = read_back(R_, R_->lights.cdf.p, (size_t)n * 4, cdf, (size_t)cap * 4); if (rc) return rc; }
        return (int)LB_OK;
    });
}
LB_API int lb_debug_read_primary_hits(LbRenderer r, void* hits, size_t cap) {
    return guarded(R_, [&]() {
        const uint32_t n = R_->npix();
        DevBuf<unsigned char> tmp; tmp.reserve((size_t)n * 20);
        launch_debug_hits(R_->cfg(), R_->d_primary_hits.p, n, tmp.p);
        return read_back(R_, tmp.p, (size_t)n * 20, hits, cap);
    });
}
LB_API int lb_debug_read_surface(LbRenderer r, float* out, size_t cap) {
    return guarded(R_, [&]() {
        const uint32_t n = R_->npix();
        DevBuf<float> tmp; tmp.reserve((size_t)n * 24);
        launch_debug_surface(R_->cfg(), R_->d_surf[R_->surf_cur ^ 1u].p, n, tmp.p);      // the frame just rendered
        return read_back(R_, tmp.p, (size_t)n * 96, out, cap);
    });
}
LB_API int lb_debug_read_reservoirs(LbRenderer r, float* out, size_t cap) {
    return guarded(R_, [&]() {
        const uint32_t n = R_->npix();
        DevBuf<float> tmp; tmp.reserve((size_t)n * 20);
        launch_debug_reservoirs(R_->cfg(), R_->d_res[R_->res_cur ^ 1u].p, n, tmp.p);
        return read_back(R_, tmp.p, (size_t)n * 80, out, cap);
    });
}
static int debug_bsdf(lb::Renderer* R, const float* mat24, const float* v12, uint32_t n, float* out, bool sample) {
    if (!mat24 || !v12 || !out) return fail(LB_ERR_INVALID_ARGUMENT, "null");
    if (!n) return LB_OK;
    const size_t width = sample ? 8 : 4;
    DevBuf<float> d_m, d_v, d_o;
    d_m.upload(mat24, 24, R->stream); d_v.upload(v12, (size_t)n * 12, R->stream); d_o.reserve((size_t)n * width);
    launch_debug_bsdf(R->cfg(), d_m.p, d_v.p, n, d_o.p, sample);
    return read_back(R, d_o.p, (size_t)n * width * 4, out, (size_t)n * width * 4);
}
LB_API int lb_debug_eval_bsdf(LbRenderer r, const float* mat24, const float* v12, uint32_t n, float* out4) { return guarded(R_, [&]() { return debug_bsdf(R_, mat24, v12, n, out4, false); }); }
LB_API int lb_debug_sample_bsdf(LbRenderer r, const float* mat24, const float* v12, uint32_t n, float* out8) { return guarded(R_, [&]() { return debug_bsdf(R_, mat24, v12, n, out8, true); }); }

}
