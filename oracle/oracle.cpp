// CPU ORACLE — TEST INFRASTRUCTURE ONLY.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
// The product (lumenrenderer_b200/, liblumen_b200.so) never links, imports or executes anything in oracle/.
//
// Scalar (OpenMP-parallel over pixels) restatement of the reference's wavefront path: every function cites
// the file:line it follows under /root/reference/Lumen_Engine/LumenPT/src/ (abbreviated PT/).
// Canonical choices for the reference's non-deterministic spots are listed in DESIGN.md ("parity hazards").
// Parity status: the BSDF is pinned bit-exactly against the reference's own headers compiled for the host
// (oracle/_ref); RNG/packing are pinned the same way; ray/triangle arithmetic is OptiX-internal in the
// reference and therefore UNPINNED (see lo_trace.h).
#include "../include/lumen_b200.h"
#include "lo_math.h"
#include "lo_bsdf.h"
#include "lo_trace.h"
#include <cfloat>
#include <vector>
#include <string>
#include <memory>
#include <numeric>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace lo {

enum : uint32_t { SURF_EMISSIVE = 1, SURF_ALPHA = 2, SURF_MISS = 4 };      // SurfaceData.h:17-23

struct Texture { uint32_t w = 1, h = 1; bool srgb = false; std::vector<uint8_t> px; };
struct Material { LbMaterialDesc desc; Mat mat; int tex_diffuse, tex_normal, tex_mr, tex_emissive, tex_transmission, tex_coat, tex_coat_rough, tex_tint; };
struct Primitive { std::vector<V3> pos, nrm; std::vector<V2> uv; std::vector<V4> tan; std::vector<uint32_t> idx; std::vector<uint8_t> emissive; uint32_t num_lights = 0; int material = 0; };
struct Mesh { std::vector<int> prims; };
struct Instance { int mesh; float m[16]; LbEmissiveness em; int override_mat; uint32_t first_entry; };
struct Entry { int inst, prim; };
struct Volume { std::vector<float> density; uint32_t nx = 0, ny = 0, nz = 0; V3 lo, hi; float majorant = 1.f; };
struct VolumeInstance { int volume; float m[16]; float inv[16]; float density; };

struct Ray { uint32_t px, py; V3 o, d, contrib; };                           // IntersectionRayData.h:24-78
struct ShadowRay { uint32_t px, py; V3 o, d; float tmax; V3 radiance; int channel; };   // ShadowRayData.h:13-62
struct Surface {                                                              // SurfaceData.h:49-104
    uint32_t px = 0, py = 0; V3 pos{0, 0, 0}, normal{0, 0, 0}, gnormal{0, 0, 0}, tangent{0, 0, 0}; float t = 0; V3 incoming{0, 0, 0};
    Mat mat{}; uint32_t flags = 0; V3 transport{0, 0, 0};
};
struct LightTri { V3 p0, p1, p2, normal, radiance; float area; };            // LightData.h:21-27
struct LightSample { V3 radiance{0, 0, 0}, normal{0, 0, 0}, position{0, 0, 0}; float area = 0; V3 contribution{0, 0, 0}; float pdf = 0; };   // ReSTIRData.h:90-102
struct Reservoir {                                                            // ReSTIRData.h:107-183
    float weight_sum = 0; long long count = 0; float weight = 0; LightSample sample;
    bool update(const LightSample& s, float w, uint32_t seed /* by value: hazard 14 */) {
        weight_sum += w; ++count;
        const float r = rand_f(seed);
        if (r <= (w / weight_sum)) { sample = s; return true; }
        return false;
    }
    void update_weight() {
        if (count == 0 || weight_sum <= 0.f) { weight = 0; return; }
        weight = (1.f / fmaxf(sample.pdf, FLT_EPSILON)) * ((1.f / (float)count) * weight_sum);
    }
    void reset() { weight_sum = 0; count = 0; weight = 0; }
};
// CDF::Get / BinarySearch (ReSTIRData.h:232-302) over the accumulated sums cdf[0..n): index of the entry whose interval holds value * sum
static void cdf_lookup(const float* cdf, int n, float cdf_sum, float value, uint32_t& index, float& pdf) {
    const float required = cdf_sum * value;
    int first = 0, last = n - 1, center = 0;
    for (;;) {
        center = (last + first) / 2;
        const float higher = cdf[center], lower = center ? cdf[center - 1] : 0.f;
        if (required < lower && center - 1 >= first) { last = center - 1; continue; }
        if (required > higher && center + 1 <= last) { first = center + 1; continue; }
        break;
    }
    const float higher = cdf[center], lower = center ? cdf[center - 1] : 0.f;
    index = (uint32_t)center; pdf = (higher - lower) / cdf_sum;
}
struct BagEntry { uint32_t light; float pdf; };                              // LightBagEntry, ReSTIRData.h:308-312 (light by index)
struct VolumeHit { float t0 = -1, t1 = -1; float density = 0; int vinst = -1; };   // VolumetricData.h

constexpr uint32_t kNumBags = 50, kLightsPerBag = 1000, kPrimarySamples = 32, kSpatialSamples = 5, kSpatialRadius = 30, kSpatialIterations = 2;   // ReSTIRData.h:34-56
constexpr float kSimilarCos = 0.72222222223f;

// lo_kat_rand_right_to_left(1) (tests only): the three RandomFloat(seed) ARGUMENTS of the reference's SampleBSDF call (GPUShadeIndirect.cu:99-101)
// are evaluated in unspecified order (hazard 3) — the canonical choice is left to right, the host build of the reference (g++, x86-64) evaluates
// them right to left. With the switch the oracle's ShadeIndirect is bit-identical to that host build, which pins everything around the call.
static bool g_kat_rand_right_to_left = false;

struct Renderer {
    LbSettings st{};
    std::vector<Texture> textures; std::vector<Material> materials; std::vector<Primitive> prims; std::vector<Mesh> meshes;
    std::vector<Instance> instances; std::vector<Entry> entries; std::vector<Volume> volumes; std::vector<VolumeInstance> vinstances;
    V3 cam_pos{0, 0, 0}; float cam_q[4]{1, 0, 0, 0}; float fov_y = 90.f;
    bool cam_from_matrix = false; float cam_m[16];
    double prev_cam[16]; bool have_prev_cam = false;
    bool scene_dirty = true;
    // derived scene
    std::vector<Tri> tris; Bvh2 bvh; std::vector<LightTri> lights; std::vector<float> cdf; float cdf_sum = 0;
    // frame state
    uint32_t frame_index = 0, surf_cur = 0, res_cur = 0, blend_count = 0;
    float cam_min_d = 0.1f, cam_max_d = 1000.f;           // Camera::m_MinMaxRenderDistance, Camera.h:60
    std::vector<Surface> surface[3]; std::vector<Reservoir> reservoirs[4];
    std::vector<V4> channel[4], combined, accum; std::vector<V2> motion; std::vector<HitRec> primary_hits; std::vector<uint8_t> ldr;
    std::vector<VolumeHit> volhits;
    std::vector<BagEntry> bags;
    uint64_t counters[8]{};
    std::vector<std::pair<std::string, float>> stats; std::string stats_names;

    uint32_t npix() const { return st.width * st.height; }
    // row band of a larger frame (LbSettings::band_*): random streams, camera and motion vectors are keyed on the full-frame position
    uint32_t full_height() const { return st.band_full_height ? st.band_full_height : st.height; }
    uint32_t pix0() const { return st.band_row0 * st.width; }
    // LbSettings::band_own_*: pixels outside the owned rows are ReSTIR halo — no NEE shadow rays, no bounce rays (ignored with media)
    bool owned(uint32_t pixel_index) const {
        if (!st.band_own_rows || !vinstances.empty()) return true;
        const uint32_t row = pixel_index / st.width + st.band_row0;
        return row >= st.band_own_row0 && row < st.band_own_row0 + st.band_own_rows;
    }
    void resize() {
        const size_t n = npix();
        for (auto& s : surface) s.assign(n, Surface{});
        for (auto& r : reservoirs) r.assign(n, Reservoir{});
        for (auto& c : channel) c.assign(n, V4{0, 0, 0, 0});
        combined.assign(n, V4{0, 0, 0, 0}); accum.assign(n, V4{0, 0, 0, 0}); motion.assign(n, V2{0, 0});
        primary_hits.assign(n, HitRec{0, 0, 0, 0, -1.f}); ldr.assign(n * 4, 0); volhits.assign(n, VolumeHit{});
        blend_count = 0; frame_index = 0; surf_cur = 0; res_cur = 0; have_prev_cam = false;
    }

    // ------------------------------------------------------------------ textures (PTTexture.cpp:35-74; exact fp32 bilinear, hazard 6)
    static float srgb_to_linear(uint8_t b) { const double c = b / 255.0; return (float)(c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4)); }
    V4 texel(const Texture& t, int x, int y) const {
        const uint8_t* p = &t.px[(size_t(y) * t.w + x) * 4];
        if (t.srgb) return {srgb_to_linear(p[0]), srgb_to_linear(p[1]), srgb_to_linear(p[2]), (float)p[3] * (1.0f / 255.0f)};
        return {(float)p[0] * (1.0f / 255.0f), (float)p[1] * (1.0f / 255.0f), (float)p[2] * (1.0f / 255.0f), (float)p[3] * (1.0f / 255.0f)};
    }
    V4 tex2d(int handle, float u, float v) const {
        const Texture& t = textures[handle];
        if (t.w == 1 && t.h == 1) return texel(t, 0, 0);
        const float fu = u - floorf(u), fv = v - floorf(v);
        const float x = fu * (float)t.w - 0.5f, y = fv * (float)t.h - 0.5f;
        const float x0f = floorf(x), y0f = floorf(y);
        const float ax = x - x0f, ay = y - y0f;
        const int x0 = (((int)x0f % (int)t.w) + (int)t.w) % (int)t.w, y0 = (((int)y0f % (int)t.h) + (int)t.h) % (int)t.h;
        const int x1 = (x0 + 1) % (int)t.w, y1 = (y0 + 1) % (int)t.h;
        const V4 a = texel(t, x0, y0), b = texel(t, x1, y0), c = texel(t, x0, y1), d = texel(t, x1, y1);
        auto mix4 = [](const V4& p, const V4& q, float s) { return V4{mixf(p.x, q.x, s), mixf(p.y, q.y, s), mixf(p.z, q.z, s), mixf(p.w, q.w, s)}; };
        return mix4(mix4(a, b, ax), mix4(c, d, ax), ay);
    }

    // ------------------------------------------------------------------ materials (WaveFrontRenderer.cpp:1260-1311, PTMaterial.cpp:160-266)
    int tex_or(LbHandle h, int def) const { return h < 0 ? def : h; }
    bool fill_material(Material& m, const LbMaterialDesc& d) {
        m.desc = d;
        Mat& p = m.mat; memset(&p, 0, sizeof p);
        pack8(p.params[0], 1.f, 24);                               // PTMaterial ctor sets roughness 1
        p.color = {d.diffuse_color[0], d.diffuse_color[1], d.diffuse_color[2], d.diffuse_color[3]};
        p.emissive = {d.emission[0], d.emission[1], d.emission[2], 0.f};
        pack8(p.params[2], d.transmission_factor, 16);
        pack8(p.params[2], d.clear_coat_factor, 0);
        pack8(p.params[2], 1.f - d.clear_coat_roughness_factor, 8);
        p.transmittance.w = d.index_of_refraction;
        pack8(p.params[0], d.specular_factor, 16);
        pack8(p.params[1], d.specular_tint_factor, 0);
        pack8(p.params[0], d.subsurface_factor, 8);
        p.tint.w = d.luminance;
        pack8(p.params[1], d.anisotropic, 8);
        pack8(p.params[1], d.sheen_factor, 16);
        pack8(p.params[1], d.sheen_tint_factor, 24);
        p.tint = {d.tint_factor[0], d.tint_factor[1], d.tint_factor[2], p.tint.w};
        p.transmittance = {d.transmittance[0], d.transmittance[1], d.transmittance[2], p.transmittance.w};
        pack8(p.params[0], d.roughness_factor, 24);
        pack8(p.params[0], d.metallic_factor, 0);
        const LbHandle hs[8] = {d.diffuse_texture, d.normal_texture, d.metallic_roughness_texture, d.emissive_texture, d.transmission_texture, d.clear_coat_texture, d.clear_coat_roughness_texture, d.tint_texture};
        for (LbHandle h : hs) if (h >= (LbHandle)textures.size()) return false;
        m.tex_diffuse = tex_or(d.diffuse_texture, 0); m.tex_normal = tex_or(d.normal_texture, 1); m.tex_mr = tex_or(d.metallic_roughness_texture, 0);
        m.tex_emissive = tex_or(d.emissive_texture, 0); m.tex_transmission = tex_or(d.transmission_texture, 0); m.tex_coat = tex_or(d.clear_coat_texture, 0);
        m.tex_coat_rough = 0;   // hazard 8: the reference never binds this slot (PTMaterial.cpp:121-129): canonical = white
        m.tex_tint = tex_or(d.tint_texture, 0);
        return true;
    }
    // FindEmissivesGpu, PT/CUDAKernels/WaveFrontKernels/GPUEmissiveLookup.cu:13-109 (runs only for materials with emission, WaveFrontRenderer.cpp:1192-1209)
    void find_emissives(Primitive& p) {
        const Material& m = materials[p.material];
        const size_t nt = p.idx.size() / 3;
        p.emissive.assign(nt, 0); p.num_lights = 0;
        if (m.mat.emissive.x == 0.f && m.mat.emissive.y == 0.f && m.mat.emissive.z == 0.f) return;
        for (size_t t = 0; t < nt; ++t) {
            const V2 c = (p.uv[p.idx[3 * t]] + p.uv[p.idx[3 * t + 1]] + p.uv[p.idx[3 * t + 2]]) * (1.f / 3.f);
            const V4 e = m.mat.emissive * tex2d(m.tex_emissive, c.x, c.y);
            if (e.x > 0.f || e.y > 0.f || e.z > 0.f) { p.emissive[t] = 1; p.num_lights++; }
        }
    }

    // ------------------------------------------------------------------ scene commit
    const Material& entry_material(const Entry& e) const {
        const Instance& in = instances[e.inst];
        return materials[in.override_mat >= 0 ? in.override_mat : prims[e.prim].material];
    }
    void commit_scene() {
        // scene data table: one entry per (mesh instance, primitive), PTMeshInstance.cpp:123-178
        entries.clear(); tris.clear();
        for (size_t i = 0; i < instances.size(); ++i) {
            instances[i].first_entry = (uint32_t)entries.size();
            for (int p : meshes[instances[i].mesh].prims) entries.push_back({(int)i, p});
        }
        for (size_t e = 0; e < entries.size(); ++e) {
            const Instance& in = instances[entries[e].inst]; const Primitive& p = prims[entries[e].prim];
            for (size_t t = 0; t < p.idx.size() / 3; ++t)
                tris.push_back({xform_point(in.m, p.pos[p.idx[3 * t]]), xform_point(in.m, p.pos[p.idx[3 * t + 1]]), xform_point(in.m, p.pos[p.idx[3 * t + 2]]), (uint32_t)e, (uint32_t)t});
        }
        bvh.build(tris);
        build_lights();
        scene_dirty = false;
        counters[5] = tris.size(); counters[6] = bvh.nodes.size(); counters[7] = bvh.nodes.size() * sizeof(Bvh2::Node);
    }
    // LightDataBuffer::BuildLightDataBuffer PT/Framework/LightDataBuffer.cpp:37-125 + BuildLightDataInstance
    // PT/CUDAKernels/WaveFrontKernels/GPUDataBufferKernels.cu:66-186, then FillCDF PT/CUDAKernels/ReSTIRKernels.cu:49-130.
    // Canonical: compact list in (table entry, triangle) order, stable sort by mean radiance (hazard 4), blocked fp32 scan.
    void build_lights() {
        lights.clear();
        for (size_t e = 0; e < entries.size(); ++e) {
            const Instance& in = instances[entries[e].inst]; const Primitive& p = prims[entries[e].prim];
            if (in.em.mode == LB_EMISSION_DISABLED) continue;
            bool mesh_emissive = false; for (int q : meshes[in.mesh].prims) mesh_emissive |= prims[q].num_lights > 0;
            if (in.em.mode == LB_EMISSION_ENABLED && !(mesh_emissive && p.num_lights > 0)) continue;
            const Material& m = entry_material(entries[e]);
            for (size_t t = 0; t < p.idx.size() / 3; ++t) {
                if (!(in.em.mode == LB_EMISSION_OVERRIDE || p.emissive[t])) continue;
                const uint32_t i0 = p.idx[3 * t], i1 = p.idx[3 * t + 1], i2 = p.idx[3 * t + 2];
                V4 em{0, 0, 0, 0};
                if (in.em.mode == LB_EMISSION_ENABLED) {
                    const V2 c = (p.uv[i0] + p.uv[i1] + p.uv[i2]) * (1.f / 3.f);
                    em = tex2d(m.tex_emissive, c.x, c.y); em = em * (m.mat.emissive * in.em.scale);
                } else {
                    em = V4{in.em.override_radiance[0], in.em.override_radiance[1], in.em.override_radiance[2], in.em.scale} * in.em.scale;
                }
                if (!(em.x > 0.f || em.y > 0.f || em.z > 0.f)) continue;
                LightTri l; l.p0 = xform_point(in.m, p.pos[i0]); l.p1 = xform_point(in.m, p.pos[i1]); l.p2 = xform_point(in.m, p.pos[i2]);
                l.radiance = v3(em);
                l.normal = normalize(xform_vector(in.m, (p.nrm[i0] + p.nrm[i1] + p.nrm[i2]) * (1.f / 3.f)));
                const V3 a = l.p0 - l.p1, b = l.p0 - l.p2;
                const float cx = a.y * b.z - b.y * a.z, cy = a.x * b.z - b.x * a.z, cz = a.x * b.y - b.x * a.y;
                l.area = sqrtf(cx * cx + cy * cy + cz * cz) / 2.0f;
                lights.push_back(l);
            }
        }
        auto key = [](const LightTri& l) { return (l.radiance.x + l.radiance.y + l.radiance.z) / 3.f; };
        std::stable_sort(lights.begin(), lights.end(), [&](const LightTri& a, const LightTri& b) { return key(a) < key(b); });
        // blocked inclusive scan: 256-element blocks summed sequentially, block totals scanned sequentially
        const size_t n = lights.size(); cdf.assign(n, 0.f);
        const size_t nb = (n + 255) / 256; std::vector<float> tot(nb, 0.f);
        for (size_t b = 0; b < nb; ++b) { float s = 0.f; for (size_t i = b * 256; i < std::min(n, (b + 1) * 256); ++i) { s += key(lights[i]); cdf[i] = s; } tot[b] = s; }
        float run = 0.f;
        for (size_t b = 0; b < nb; ++b) { if (b) for (size_t i = b * 256; i < std::min(n, (b + 1) * 256); ++i) cdf[i] = run + cdf[i]; run = b ? run + tot[b] : tot[b]; }
        cdf_sum = n ? cdf[n - 1] : 0.f;
        counters[4] = n;
    }
    // CDF::Get / BinarySearch, PT/Shaders/CppCommon/ReSTIRData.h:232-302
    void cdf_get(float value, uint32_t& index, float& pdf) const { cdf_lookup(cdf.data(), (int)cdf.size(), cdf_sum, value, index, pdf); }

    // ------------------------------------------------------------------ camera (LM/Renderer/Camera.cpp:79-93,122-140)
    void camera_matrix(double m[16]) const {   // row-major world matrix (columns right, up, forward, position)
        if (cam_from_matrix) { for (int k = 0; k < 16; ++k) m[k] = cam_m[k]; return; }
        const double w = cam_q[0], x = cam_q[1], y = cam_q[2], z = cam_q[3];
        const double c0[3] = {1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)};
        const double c1[3] = {2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)};
        const double c2[3] = {2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)};
        for (int r = 0; r < 3; ++r) { m[r * 4 + 0] = c0[r]; m[r * 4 + 1] = c1[r]; m[r * 4 + 2] = c2[r]; }
        m[3] = cam_pos.x; m[7] = cam_pos.y; m[11] = cam_pos.z; m[12] = m[13] = m[14] = 0; m[15] = 1;
    }
    void camera_vectors(V3& eye, V3& U, V3& V, V3& W) const {
        double m[16]; camera_matrix(m);
        const float half_y = 1.0f * tanf((fov_y * 0.01745329251994329576923690768489f) * 0.5f);
        const float half_x = half_y * ((float)st.width / (float)full_height());
        eye = cam_pos;
        U = v3((float)m[0], (float)m[4], (float)m[8]) * half_x;
        V = v3((float)m[1], (float)m[5], (float)m[9]) * half_y;
        W = v3((float)m[2], (float)m[6], (float)m[10]) * 1.0f;
    }
    // projection * inverse(previous camera matrix), WaveFrontRenderer.cpp:760-781 + CPUShadingKernels.cu:27-54
    void prev_view_proj(float out[16]) const {
        double cur[16]; camera_matrix(cur);
        const double* c = have_prev_cam ? prev_cam : cur;
        double view[16];   // inverse of rigid transform
        for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) view[r * 4 + k] = c[k * 4 + r]; view[r * 4 + 3] = -(c[0 * 4 + r] * c[3] + c[1 * 4 + r] * c[7] + c[2 * 4 + r] * c[11]); }
        view[12] = view[13] = view[14] = 0; view[15] = 1;
        const double aspect = (double)st.width / (double)full_height(), zn = 0.5, zf = 10000.0, th = tan((fov_y * 0.01745329251994329576923690768489) / 2.0);
        double P[16] = {1.0 / (aspect * th), 0, 0, 0, 0, 1.0 / th, 0, 0, 0, 0, -(zf + zn) / (zf - zn), -(2.0 * zf * zn) / (zf - zn), 0, 0, -1, 0};
        for (int r = 0; r < 4; ++r) for (int k = 0; k < 4; ++k) { double s = 0; for (int j = 0; j < 4; ++j) s += P[r * 4 + j] * view[j * 4 + k]; out[r * 4 + k] = (float)s; }
    }

    // ------------------------------------------------------------------ K1 ray generation, GPUGeneratePrimRay.cu:8-82
    static float halton(uint32_t index, uint32_t base) {
        ++index; float f = 1.f, r = 0.f;
        while (index > 0) { f = f / (float)base; r = r + f * (float)(index % base); index = index / base; }
        return r;
    }
    void raygen(uint32_t frame_count, std::vector<Ray>& rays) const {
        V3 eye, U, V, W; camera_vectors(eye, U, V, W);
        rays.resize(npix());
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)npix(); ++i) {
            const int sy = (int)(i / st.width), sx = (int)(i - (int64_t)sy * st.width);
            const uint32_t gi = (uint32_t)i + pix0();
            const float jx = halton(frame_count + gi, 2), jy = halton(frame_count + gi, 3);
            float dx = ((float)sx + jx) / (float)st.width, dy = ((float)((uint32_t)sy + st.band_row0) + jy) / (float)full_height();
            dx = -(dx * 2.0f - 1.0f); dy = -(dy * 2.0f - 1.0f);
            // canonical fused order (identical on the GPU): dx*U + (dy*V + W)
            const V3 d = v3(fmaf(dx, U.x, fmaf(dy, V.x, W.x)), fmaf(dx, U.y, fmaf(dy, V.y, W.y)), fmaf(dx, U.z, fmaf(dy, V.z, W.z)));
            const float len2 = fmaf(d.x, d.x, fmaf(d.y, d.y, d.z * d.z));
            const float inv = 1.0f / sqrtf(len2);
            rays[i] = {(uint32_t)sx, (uint32_t)sy, eye, v3(d.x * inv, d.y * inv, d.z * inv), v3(1.f, 1.f, 1.f)};
        }
    }

    // ------------------------------------------------------------------ K2 extend, WaveFrontShaders.cu:42-112,301-328
    void extend(const std::vector<Ray>& rays, std::vector<HitRec>& hits) const {
        hits.resize(rays.size());
        #pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)rays.size(); ++i) {
            HitRec h;
            if (bvh.closest(rays[i].o, rays[i].d, 0.01f, 5000.f, h)) { h.u = half_round(h.u); h.v = half_round(h.v); hits[i] = h; }   // fp16 barycentrics, hazard 7
            else hits[i] = {0, 0, 0, 0, -1.f};
        }
    }
    // volume bbox slab (K5, PT/Shaders/volumetric_wavefront.cu:52-95): nearest volume instance in front of the surface hit
    void extend_volumes(const std::vector<Ray>& rays, const std::vector<HitRec>& hits, std::vector<VolumeHit>& out) const {
        out.assign(rays.size(), VolumeHit{});
        if (vinstances.empty()) return;
        for (size_t i = 0; i < rays.size(); ++i) {
            const float tmax = hits[i].t > 0.f ? fminf(5000.f, hits[i].t) : 5000.f;
            for (size_t k = 0; k < vinstances.size(); ++k) {
                const VolumeInstance& vi = vinstances[k]; const Volume& vol = volumes[vi.volume];
                const V3 o = xform_point(vi.inv, rays[i].o), d = xform_vector(vi.inv, rays[i].d);
                float t0 = 0.01f, t1 = tmax; bool ok = true;
                for (int a = 0; a < 3 && ok; ++a) {
                    const float inv = 1.0f / comp(d, a);
                    float ta = (comp(vol.lo, a) - comp(o, a)) * inv, tb = (comp(vol.hi, a) - comp(o, a)) * inv;
                    if (ta > tb) std::swap(ta, tb);
                    t0 = fmaxf(t0, ta); t1 = fminf(t1, tb); ok = t0 <= t1;
                }
                if (ok && (out[i].vinst < 0 || t0 < out[i].t0)) out[i] = {t0, t1, vi.density, (int)k};
            }
        }
    }

    // ------------------------------------------------------------------ K6 surface extraction, GPUExtractSurfaceData.cu:8-228
    void extract(const std::vector<Ray>& rays, const std::vector<HitRec>& hits, std::vector<Surface>& out) const {
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)rays.size(); ++i) {
            const Ray& ray = rays[i]; const HitRec& hit = hits[i];
            Surface& dst = out[(size_t)ray.py * st.width + ray.px];
            if (!(hit.t > 0.f)) { dst.flags = SURF_MISS; continue; }
            const Entry& e = entries[hit.inst]; const Instance& in = instances[e.inst]; const Primitive& p = prims[e.prim]; const Material& m = entry_material(e);
            const uint32_t ia = p.idx[3 * hit.prim], ib = p.idx[3 * hit.prim + 1], ic = p.idx[3 * hit.prim + 2];
            const float U = hit.u, V = hit.v, W = 1.f - (U + V);
            const V2 uv = p.uv[ia] * W + p.uv[ib] * U + p.uv[ic] * V;
            const float flip = p.tan[ia].w;
            const V4 nmap = tex2d(m.tex_normal, uv.x, uv.y), tcol = tex2d(m.tex_diffuse, uv.x, uv.y);
            V4 em{0, 0, 0, 0};
            if (in.em.mode == LB_EMISSION_ENABLED) { em = m.mat.emissive * in.em.scale; em = em * tex2d(m.tex_emissive, uv.x, uv.y); }
            else if (in.em.mode == LB_EMISSION_OVERRIDE) em = V4{in.em.override_radiance[0], in.em.override_radiance[1], in.em.override_radiance[2], in.em.scale} * in.em.scale;
            Surface s; s.flags = 0;
            const V3 ln = normalize(p.nrm[ia] * W + p.nrm[ib] * U + p.nrm[ic] * V);
            const V3 lt0 = v3(p.tan[ia]) * W + v3(p.tan[ib]) * U + v3(p.tan[ic]) * V;
            const V3 lt = normalize(lt0);
            const V3 nw = normalize(xform_vector(in.m, ln)), tw = normalize(xform_vector(in.m, lt));
            const V3 bw = cross(nw, tw) * flip;
            V3 nm = v3(nmap.x, nmap.y, nmap.z) * 2.f - v3(1.f);
            nm = normalize(nm);
            nm = normalize(v3(nm.x * tw.x + nm.y * bw.x + nm.z * nw.x, nm.x * tw.y + nm.y * bw.y + nm.z * nw.y, nm.x * tw.z + nm.y * bw.z + nm.z * nw.z));
            s.px = ray.px; s.py = ray.py; s.t = hit.t; s.normal = nm;
            if (em.x > 0.f || em.y > 0.f || em.z > 0.f) {
                const float mx = fmaxf(em.x, fmaxf(em.y, em.z)); const float inv = 1.0f / mx;     // float4 /= float multiplies by the reciprocal
                s.mat.color = em * inv; s.flags |= SURF_EMISSIVE; dst = s; continue;
            }
            s.pos = ray.o + ray.d * hit.t; s.incoming = ray.d; s.transport = ray.contrib;
            if (tcol.w < 0.51f) { s.flags |= SURF_ALPHA; dst = s; continue; }
            const float eta = 1.f / m.mat.transmittance.w;
            s.gnormal = nw; s.tangent = tw; s.mat = m.mat;
            const MatView mv(m.mat);
            const V4 mr = tex2d(m.tex_mr, uv.x, uv.y);
            pack8(s.mat.params[0], mr.z * mv.metallic, 0);
            pack8(s.mat.params[0], mr.y * mv.roughness, 24);
            s.mat.color = tcol * m.mat.color;
            const V4 cc = tex2d(m.tex_coat, uv.x, uv.y), ccr = tex2d(m.tex_coat_rough, uv.x, uv.y), tr = tex2d(m.tex_transmission, uv.x, uv.y), ti = tex2d(m.tex_tint, uv.x, uv.y);
            const V3 tint = v3(ti.x, ti.y, ti.z) * mv.tint;
            pack8(s.mat.params[2], mv.clearcoat * cc.x, 0);
            pack8(s.mat.params[2], mv.clearcoatgloss * (1.f - ccr.x), 8);
            s.mat.tint = v4(tint, s.mat.tint.w);
            pack8(s.mat.params[2], mv.transmission * tr.x, 16);
            s.mat.transmittance.w = eta;
            dst = s;
        }
    }

    // ------------------------------------------------------------------ K8 motion vectors, MotionVectors.cu:8-55
    void motion_vectors(const std::vector<Surface>& surf) {
        float M[16]; prev_view_proj(M);
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)npix(); ++i) {
            const uint32_t y = (uint32_t)(i / st.width), x = (uint32_t)(i - (int64_t)y * st.width);
            V2 mv{0.f, 0.f}; const Surface& s = surf[i];
            if (s.t > 0.f) {
                const float cx = ((float)x + 0.5f) / (float)st.width, cy = ((float)(y + st.band_row0) + 0.5f) / (float)full_height();
                const float px = fmaf(M[0], s.pos.x, fmaf(M[1], s.pos.y, fmaf(M[2], s.pos.z, M[3])));
                const float py = fmaf(M[4], s.pos.x, fmaf(M[5], s.pos.y, fmaf(M[6], s.pos.z, M[7])));
                const float pw = fmaf(M[12], s.pos.x, fmaf(M[13], s.pos.y, fmaf(M[14], s.pos.z, M[15])));
                const float inv = 1.0f / pw;
                mv = {half_round((px * inv * 0.5f + 0.5f) - cx), half_round((py * inv * 0.5f + 0.5f) - cy)};
            }
            motion[i] = mv;
        }
    }

    // ------------------------------------------------------------------ Resample, ReSTIRKernels.cu:1259-1325
    static void resample(const LightSample& in, const Surface& px, LightSample& out) {
        out = in;
        V3 dir = in.position - px.pos; const float dist = length(dir); dir /= dist;
        const float cos_in = fmaxf(dot(dir, px.normal), 0.f), cos_out = fmaxf(dot(in.normal, -dir), 0.f);
        if (cos_in <= 0 || cos_out <= 0 || dist <= 0.01f) { out.pdf = 0; return; }
        const float solid = (cos_out * in.area) / (dist * dist);
        float pdf = 0.f; const V3 bsdf = disney_eval(px.mat, px.normal, px.tangent, -px.incoming, dir, pdf);
        const float added = pdf + bsdf.x + bsdf.y + bsdf.z;
        if (pdf <= kBsdfEps || std::isnan(added) || std::isinf(added)) { out.contribution = v3(0); out.pdf = 0; return; }
        const V3 c = (bsdf / pdf) * solid * cos_in * out.radiance;
        out.contribution = c; out.pdf = (c.x + c.y + c.z) / 3.f;
    }
    // CombineBiased, ReSTIRKernels.cu:1200-1257
    static void combine_biased(Reservoir& dst, int n, const Reservoir* in, const Surface& px, uint32_t seed) {
        Reservoir out; long long total = 0;
        for (int i = 0; i < n; ++i) {
            LightSample rs; resample(in[i].sample, px, rs);
            const float w = (float)in[i].count * in[i].weight * rs.pdf;
            out.update(rs, w, seed); total += in[i].count;
        }
        out.count = total; out.update_weight(); dst = out;
    }

    // CombineUnbiased, ReSTIRKernels.cu:1123-1198 (dead in the reference's build: `constexpr bool enableBiased = true`, ReSTIRKernels.cu:880 —
    // restated and pinned for completeness, selected by LbSettings::restir_unbiased). `from[i]` is the pixel reservoir i was generated at:
    // the selected sample is re-evaluated there, and only reservoirs for whose pixel it has a non-zero target pdf count towards M.
    static void combine_unbiased(Reservoir& dst, const Surface& px, int n, const Reservoir* in, const Surface* from, uint32_t seed) {
        Reservoir out; int total = 0;
        for (int i = 0; i < n; ++i) {
            LightSample rs; resample(in[i].sample, px, rs);
            const float w = (float)in[i].count * in[i].weight * rs.pdf;
            out.update(rs, w, seed); total += (int)in[i].count;
        }
        out.count = total;
        int correction = 0;
        for (int i = 0; i < n; ++i) {
            LightSample rs; resample(out.sample, from[i], rs);
            if (rs.pdf > 0) correction += (int)in[i].count;
        }
        const float m = 1.f / fmaxf((float)correction, FLT_EPSILON);                 // MINFLOAT = std::numeric_limits<float>::epsilon(), ReSTIRData.h:12
        out.weight = (1.f / fmaxf(out.sample.pdf, FLT_EPSILON)) * (m * out.weight_sum);
        dst = out;
    }

    // ------------------------------------------------------------------ NEE: ShadeDirect, GPUShadeDirect.cu:42-153
    bool shade_direct_pixel(const Surface& s, uint32_t pixel_index, uint32_t seed_in, int chan, const VolumeHit* vh, std::vector<ShadowRay>* vol_rays, ShadowRay& out) {
        uint32_t seed = wang_hash(seed_in + pixel_index + pix0());
        if (vh && vh->vinst >= 0 && vh->t1 > vh->t0 && st.volume_mode == LB_VOLUME_COMPAT && !lights.empty()) volume_compat_pixel(s, pixel_index, *vh, seed, *vol_rays);
        if (s.flags || lights.empty()) return false;
        uint32_t li; float lpdf; cdf_get(rand_f(seed), li, lpdf);
        const LightTri& l = lights[li];
        const float u = rand_f(seed), v = rand_f(seed) * (1.f - u);
        const V3 point = l.p0 + ((l.p1 - l.p0) * u) + ((l.p2 - l.p0) * v);
        V3 dir = point - s.pos; const float dist = length(dir); dir /= dist;
        const float cos_in = fmaxf(dot(dir, s.normal), 0.f), cos_out = fmaxf(0.f, dot(l.normal, -dir));
        if (cos_in <= 0.f || dist <= 0.01f) return false;
        const float solid = (cos_out * l.area) / (dist * dist);
        float bpdf = 0.f; const V3 bsdf = disney_eval(s.mat, s.normal, s.tangent, -s.incoming, dir, bpdf);
        if (bpdf <= kBsdfEps) return false;
        V3 c = (bsdf / bpdf) * solid * cos_in * l.radiance;
        c *= ((1.f / lpdf) * s.transport);
        out = {pixel_index % st.width, pixel_index / st.width, s.pos, dir, dist - 0.2f, c, chan};
        if (!vinstances.empty() && st.volume_mode == LB_VOLUME_DELTA) {          // shadow rays cross the media: ratio-tracked transmittance
            uint32_t vseed = wang_hash((seed_in ^ 0x85ebca6bu) + pixel_index + pix0());
            out.radiance *= ratio_transmittance(out.o, out.d, 0.01f, out.tmax, vseed);
        }
        return true;
    }
    // VolumetricShadeDirect (compat mode), GPUVolumetricShadeDirect.cu:8-101; ray data reconstructed from the pixel's ray
    std::vector<Ray> const* cur_rays_by_pixel = nullptr; std::vector<int> ray_of_pixel;
    void volume_compat_pixel(const Surface&, uint32_t pixel_index, const VolumeHit& vh, uint32_t& seed, std::vector<ShadowRay>& vol_rays) {
        const Ray& ray = (*cur_rays_by_pixel)[ray_of_pixel[pixel_index]];
        const V3 entry = ray.o + ray.d * vh.t0;
        const float distance = vh.t1 - vh.t0; float acc = 0.f; const float step = distance / 5;
        V3 prev = entry; const float offset = rand_f(seed) * step;
        for (int i = 0; i < 5 && acc < 1.0f && (float)i * step < distance; i++) {
            const float ts = (float)i * step + offset; const V3 p = entry + ray.d * ts;
            const float dprev = length(p - prev); prev = p;
            uint32_t li; float lpdf; cdf_get(rand_f(seed), li, lpdf); const LightTri& l = lights[li];
            const float u = rand_f(seed), v = rand_f(seed) * (1.f - u);
            const V3 point = l.p0 + ((l.p1 - l.p0) * u) + ((l.p2 - l.p0) * v);
            V3 dir = point - p; const float ld = length(dir); dir /= ld;
            vol_rays.push_back({ray.px, ray.py, p, dir, ld - 0.2f, v3(1.f, 1.f, 1.f) * 0.01f, LB_CHANNEL_VOLUMETRIC});
            acc += vh.density * dprev;
        }
        channel[LB_CHANNEL_VOLUMETRIC][pixel_index] = {0.f, 0.f, 0.f, acc};
    }
    // shadow rays: ShadowRaysRayGen, WaveFrontShaders.cu:114-179 (fp32 accumulate, one ray per pixel per launch: hazard 2)
    void resolve_shadow(const std::vector<ShadowRay>& rays) {
        std::vector<uint8_t> occ(rays.size());
        #pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)rays.size(); ++i) occ[i] = bvh.any(rays[i].o, rays[i].d, 0.01f, rays[i].tmax) ? 1 : 0;
        for (size_t i = 0; i < rays.size(); ++i) if (!occ[i]) {
            V4& c = channel[rays[i].channel][(size_t)rays[i].py * st.width + rays[i].px];
            c.x += rays[i].radiance.x; c.y += rays[i].radiance.y; c.z += rays[i].radiance.z;
        }
        counters[1] += rays.size();
    }

    // ------------------------------------------------------------------ ShadeIndirect, GPUShadeIndirect.cu:7-146
    bool shade_indirect_pixel(const Surface& s, uint32_t pixel_index, uint32_t seed_in, Ray& out) const {
        uint32_t seed = wang_hash(seed_in + wang_hash(pixel_index + pix0()));
        const uint32_t px = pixel_index % st.width, py = pixel_index / st.width;
        if (s.flags & SURF_ALPHA) { out = {px, py, s.pos, s.incoming, s.transport}; return true; }
        if (s.flags) return false;
        if (fabsf(dot(s.normal, s.incoming)) < 3.f * kBsdfEps) return false;
        V3 wi = v3(0); float pdf = 0.f; bool specular = false;
        float r0, r1, r2;
        if (!g_kat_rand_right_to_left) { r0 = rand_f(seed); r1 = rand_f(seed); r2 = rand_f(seed); }      // left to right, hazard 3 (canonical; what nvcc's device front end does)
        else { r2 = rand_f(seed); r1 = rand_f(seed); r0 = rand_f(seed); }                                // what g++ makes of the same call in the host build of the reference (tests only)
        const V3 bsdf = disney_sample(s.mat, s.normal, s.normal, s.tangent, -s.incoming, 1.f, r0, r1, r2, wi, pdf, specular);
        if (pdf <= kBsdfEps || std::isnan(pdf + bsdf.x + bsdf.y + bsdf.z)) return false;
        const float rr = specular ? 1.f : fminf(fmaxf(bsdf.x, fmaxf(bsdf.y, bsdf.z)), 1.f);
        const float rnd = rand_f(seed);
        if (rr < rnd) return false;
        V3 c = s.transport * (1.f / rr);
        c *= bsdf * fabsf(dot(s.normal, wi)) * (1.f / pdf);
        out = {px, py, s.pos, wi, c};
        return true;
    }

    // ------------------------------------------------------------------ ReSTIR::Run, PT/Framework/ReSTIR.cpp:65-233
    // GenerateShadowRay, ReSTIRKernels.cu:546-582: the ray of a reservoir's sample (origin = the surface position); false = no ray
    static bool visibility_ray(const Surface& s, const Reservoir& r, V3& d, float& tmax) {
        if (s.flags || !(r.weight > 0.f)) return false;
        d = r.sample.position - s.pos; const float l = length(d); d /= l;
        tmax = l - 0.05f;
        return true;
    }
    void visibility_and_shade(std::vector<Reservoir>& res, const std::vector<Surface>& surf) {
        // GenerateShadowRay ReSTIRKernels.cu:546-582 + ReSTIRRayGen WaveFrontShaders.cu:181-216 + ShadeReservoirs :600-665
        const float shaded = 1.f + (st.restir_temporal ? 1.f : 0.f) + (st.restir_spatial ? 1.f : 0.f);
        uint64_t nrays = 0;
        #pragma omp parallel for schedule(dynamic, 256) reduction(+ : nrays)
        for (int64_t i = 0; i < (int64_t)npix(); ++i) {
            const Surface& s = surf[i]; Reservoir& r = res[i];
            V3 d; float tmax;
            if (visibility_ray(s, r, d, tmax)) {
                ++nrays;
                if (bvh.any(s.pos, d, 0.1f, tmax)) r.weight = 0.f;
            }
            if (r.weight > 0.f) {
                const V3 c = r.sample.contribution * (r.weight / shaded);
                V4& o = channel[LB_CHANNEL_DIRECT][i]; o.x += c.x; o.y += c.y; o.z += c.z;
            }
        }
        counters[2] += nrays;
    }
    // ---- the passes of ReSTIR::Run as separate stages over explicit buffers (restir_run below chains them; the known-answer taps at the end
    //      of this file call them one by one against the reference's kernels compiled in place)
    // FillLightBagsInternal, ReSTIRKernels.cu:343-370
    void fill_bags(uint32_t a_seed) {
        bags.resize(kNumBags * kLightsPerBag);
        for (uint32_t i = 0; i < kNumBags * kLightsPerBag; ++i) { uint32_t s = wang_hash(a_seed + wang_hash(i)); const float r = rand_f(s); cdf_get(r, bags[i].light, bags[i].pdf); }
    }
    // PickPrimarySamplesInternal, ReSTIRKernels.cu:402-522
    void ris_stage(const std::vector<Surface>& cur, std::vector<Reservoir>& R, uint32_t seed) const {
        const uint32_t n = (uint32_t)cur.size();
        #pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            uint32_t bag_seed = wang_hash(seed + ((uint32_t)i + pix0()) / 256u);                       // hazard 1: block index instead of %smid
            const int bag = (int)roundf((float)(kNumBags - 1) * rand_f(bag_seed));
            const BagEntry* picked = &bags[(size_t)bag * kLightsPerBag];
            const Surface& px = cur[i];
            if (px.flags) { R[i] = Reservoir{}; continue; }
            uint32_t s = wang_hash(seed + wang_hash((uint32_t)i + pix0()));
            Reservoir fresh;
            for (uint32_t k = 0; k < kPrimarySamples; ++k) {
                const float r = rand_f(s);
                const BagEntry& be = picked[(int)roundf((float)(kLightsPerBag - 1) * r)];
                const LightTri& l = lights[be.light];
                const float u = rand_f(s), v = rand_f(s) * (1.f - u);
                LightSample ls; ls.radiance = l.radiance; ls.normal = l.normal; ls.area = l.area;
                ls.position = l.p0 + ((l.p1 - l.p0) * u) + ((l.p2 - l.p0) * v);
                LightSample rs; resample(ls, px, rs);
                fresh.update(rs, rs.pdf / be.pdf, s);
            }
            fresh.update_weight(); R[i] = fresh;
        }
    }
    // CombineTemporalSamplesInternal, ReSTIRKernels.cu:1015-1121. `direct` receives the shading of the previous reservoir (:1105, ShadeReservoirs)
    void temporal_stage(const std::vector<Surface>& cur, const std::vector<Surface>& prev, std::vector<Reservoir>& R, const std::vector<Reservoir>& Rprev,
                        const std::vector<V2>& mv, std::vector<V4>& direct, uint32_t seed, float shaded) const {
        const uint32_t n = (uint32_t)cur.size();
        #pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            const int cy = (int)(i / st.width), cx = (int)(i - (int64_t)cy * st.width);
            const int mx = (int)roundf((float)st.width * mv[i].x), my = (int)roundf((float)full_height() * mv[i].y);
            int ty = cy + my, tx = cx + mx; int64_t ti = i;
            if (ty >= 0 && ty < (int)st.height && tx >= 0 && tx < (int)st.width) ti = (int64_t)ty * st.width + tx;
            const Surface& sp = prev[ti]; const Surface& sc = cur[i];
            if (sp.flags || sc.flags) continue;
            Reservoir pair[2] = {Rprev[ti], R[i]};
            const float d1 = sp.t, d2 = sc.t; const float pct = fabsf(d1 - d2) / ((d1 + d2) / 2.f);
            const float ang = dot(sp.normal, sc.normal);
            if (!(pct < 0.10f && ang > kSimilarCos)) continue;
            if (Rprev[ti].weight > 0.f) { const V3 c = Rprev[ti].sample.contribution * (Rprev[ti].weight / shaded); V4& o = direct[i]; o.x += c.x; o.y += c.y; o.z += c.z; }
            pair[0].count = std::min(pair[0].count, pair[1].count * 20);
            if (st.restir_unbiased) { const Surface from[2] = {sp, sc}; combine_unbiased(R[i], sc, 2, pair, from, wang_hash(seed + (uint32_t)i + pix0())); }
            else combine_biased(R[i], 2, pair, sc, wang_hash(seed + (uint32_t)i + pix0()));
        }
    }
    // SpatialNeighbourSamplingInternal, ReSTIRKernels.cu:787-980 (one iteration; driver :745-785)
    void spatial_pass(const std::vector<Surface>& cur, const std::vector<Reservoir>& In, std::vector<Reservoir>& Out, uint32_t seed) const {
        const uint32_t n = (uint32_t)cur.size();
        #pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            const Surface& sc = cur[i]; if (sc.flags) continue;
            uint32_t s = wang_hash(seed + (uint32_t)i + pix0());
            const int y = (int)(i / st.width), x = (int)(i - (int64_t)y * st.width);
            const Surface* pd[kSpatialSamples]; const Reservoir* pr[kSpatialSamples]; int count = 0;
            for (uint32_t k = 0; k < kSpatialSamples; ++k) {
                const int ny = (int)roundf((rand_f(s) * 2.f - 1.f) * (float)kSpatialRadius) + y;
                const int nx = (int)roundf((rand_f(s) * 2.f - 1.f) * (float)kSpatialRadius) + x;
                if (nx < 0 || nx >= (int)st.width || ny < 0 || ny >= (int)st.height) continue;
                const int64_t ni = (int64_t)ny * st.width + nx;
                pd[count] = &cur[ni];
                if (pd[count]->flags) continue;
                pr[count] = &In[ni];
                const float d1 = pd[count]->t, d2 = sc.t; const float pct = fabsf(d1 - d2) / ((d1 + d2) / 2.f);
                const float ang = dot(pd[count]->normal, sc.normal);
                if (pct < 0.10f && ang > kSimilarCos) ++count;
            }
            if (count > 1) {
                Reservoir out; long long total = 0;
                for (int k = 0; k < count; ++k) {
                    LightSample rs; resample(pr[k]->sample, *pd[0], rs);   // against the FIRST accepted neighbour
                    const float w = (float)pr[k]->count * pr[k]->weight * rs.pdf;
                    out.update(rs, w, seed); total += pr[k]->count;        // kernel-wide seed by value (hazard 14)
                }
                if (!st.restir_unbiased) { out.count = total; out.update_weight(); }
                else {
                    // the unbiased branch (:905-970, dead at the reference's enableBiased = true): the selected sample is re-evaluated at every
                    // accepted neighbour; the reference adds the sample count of the OUTPUT buffer's stale reservoir of this pixel
                    // (a_ReservoirsOut[index].sampleCount, :951), not the neighbour's — restated as written
                    out.count = (int)total; int correction = 0;
                    for (int k = 0; k < count; ++k) { LightSample rs; resample(out.sample, *pd[k], rs); if (rs.pdf > 0) correction += (int)Out[i].count; }
                    const float m = 1.f / fmaxf((float)correction, FLT_EPSILON);
                    out.weight = (1.f / fmaxf(out.sample.pdf, FLT_EPSILON)) * (m * out.weight_sum);
                }
                Out[i] = out;
            } else Out[i].reset();
        }
    }
    // CombineReservoirBuffersInternal, ReSTIRKernels.cu:1407-1436 (always CombineBiased, :1429-1433)
    void combine_buffers(const std::vector<Surface>& cur, std::vector<Reservoir>& R, const std::vector<Reservoir>& Nb, uint32_t cseed) const {
        const uint32_t n = (uint32_t)cur.size();
        #pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            if (cur[i].flags) continue;
            Reservoir pair[2] = {R[i], Nb[i]};
            combine_biased(R[i], 2, pair, cur[i], wang_hash(cseed + (uint32_t)i + pix0()));
        }
    }
    void restir_run(const std::vector<Surface>& cur, const std::vector<Surface>& prev, uint32_t a_seed) {
        std::vector<Reservoir>& R = reservoirs[res_cur]; std::vector<Reservoir>& Rprev = reservoirs[res_cur == 1 ? 0 : 1];
        uint32_t seed = wang_hash(a_seed);
        fill_bags(a_seed);
        seed = wang_hash(seed);
        ris_stage(cur, R, seed);
        visibility_and_shade(R, cur);
        if (st.restir_temporal) {
            seed = wang_hash(seed);
            temporal_stage(cur, prev, R, Rprev, motion, channel[LB_CHANNEL_DIRECT], seed, 1.f + 1.f + (st.restir_spatial ? 1.f : 0.f));
        }
        if (st.restir_spatial) {
            seed = wang_hash(seed);
            std::vector<Reservoir>* from = &R; std::vector<Reservoir>* to = &reservoirs[2];
            for (uint32_t it = 0; it < kSpatialIterations; ++it) {
                spatial_pass(cur, *from, *to, seed);
                if (it == 0) { from = &reservoirs[2]; to = &reservoirs[3]; } else std::swap(from, to);
            }
            visibility_and_shade(R, cur);                                           // on the CURRENT reservoirs again, ReSTIR.cpp:211-212
            combine_buffers(cur, R, *from, wang_hash(seed));
        }
    }

    // ------------------------------------------------------------------ delta tracking (north_star item 4; no reference counterpart)
    float volume_density(const Volume& v, const V3& p) const {
        if (v.density.empty()) return 1.f;
        const V3 e = v.hi - v.lo;
        int x = (int)((p.x - v.lo.x) / e.x * (float)v.nx), y = (int)((p.y - v.lo.y) / e.y * (float)v.ny), z = (int)((p.z - v.lo.z) / e.z * (float)v.nz);
        x = x < 0 ? 0 : (x >= (int)v.nx ? (int)v.nx - 1 : x); y = y < 0 ? 0 : (y >= (int)v.ny ? (int)v.ny - 1 : y); z = z < 0 ? 0 : (z >= (int)v.nz ? (int)v.nz - 1 : z);
        return v.density[((size_t)z * v.ny + y) * v.nx + x];
    }
    // returns true when the ray scatters inside the medium before t1; RNG stream: WangHash(seed ^ 0x9e3779b9 + pixel)
    bool delta_track(const VolumeHit& vh, const Ray& ray, uint32_t& seed, float& t_scatter) const {
        const VolumeInstance& vi = vinstances[vh.vinst]; const Volume& vol = volumes[vi.volume];
        const float sigma_max = vi.density * vol.majorant;
        if (!(sigma_max > 0.f)) return false;
        const V3 o = xform_point(vi.inv, ray.o), d = xform_vector(vi.inv, ray.d);
        float t = vh.t0;
        for (int it = 0; it < 1024; ++it) {
            t -= logf(1.0f - rand_f(seed) * 0.99999994f) / sigma_max;
            if (t >= vh.t1) return false;
            const float dens = vi.density * volume_density(vol, o + d * t);
            if (rand_f(seed) * sigma_max < dens) { t_scatter = t; return true; }
        }
        return false;
    }

    // ratio-tracking estimate of the transmittance of [tmin, tmax] through every volume instance (homogeneous media: analytic)
    float ratio_transmittance(const V3& ro, const V3& rd, float tmin, float tmax, uint32_t& seed) const {
        float tr = 1.f;
        for (size_t k = 0; k < vinstances.size(); ++k) {
            const VolumeInstance& vi = vinstances[k]; const Volume& vol = volumes[vi.volume];
            const V3 o = xform_point(vi.inv, ro), d = xform_vector(vi.inv, rd);
            float t0 = tmin, t1 = tmax; bool ok = true;
            for (int a = 0; a < 3 && ok; ++a) {
                const float inv = 1.0f / comp(d, a);
                float ta = (comp(vol.lo, a) - comp(o, a)) * inv, tb = (comp(vol.hi, a) - comp(o, a)) * inv;
                if (ta > tb) std::swap(ta, tb);
                t0 = fmaxf(t0, ta); t1 = fminf(t1, tb); ok = t0 <= t1;
            }
            if (!ok) continue;
            const float sigma_max = vi.density * vol.majorant;
            if (!(sigma_max > 0.f)) continue;
            if (vol.density.empty()) { tr *= expf(-sigma_max * (t1 - t0)); continue; }
            float t = t0;
            for (int it = 0; it < 1024; ++it) {
                t -= logf(1.0f - rand_f(seed) * 0.99999994f) / sigma_max;
                if (t >= t1) break;
                const float dens = vi.density * volume_density(vol, o + d * t);
                tr *= 1.0f - dens / sigma_max;
                if (tr <= 0.f) return 0.f;
            }
        }
        return tr;
    }
    // LB_VOLUME_DELTA: a real collision inside the nearest medium replaces this wave's surface interaction by an isotropic
    // scattering event (albedo 0.8): NEE with ratio-tracked transmittance + continuation ray. RNG stream WangHash((seed ^ 0x9e3779b9) + pixel).
    void delta_scatter_all(std::vector<Ray>& rays, std::vector<HitRec>& hits, const std::vector<VolumeHit>& vhits, uint32_t seed_in, int chan, bool bounce,
                           std::vector<ShadowRay>& srays, std::vector<Ray>& next) const {
        for (size_t i = 0; i < rays.size(); ++i) {
            const VolumeHit& vh = vhits[i];
            if (vh.vinst < 0) continue;
            const Ray& ray = rays[i];
            const uint32_t pixel = ray.py * st.width + ray.px;
            uint32_t seed = wang_hash((seed_in ^ 0x9e3779b9u) + pixel + pix0());
            float ts;
            if (!delta_track(vh, ray, seed, ts)) continue;
            hits[i].t = -2.f;
            const V3 p = ray.o + ray.d * ts; const V3 T = ray.contrib;
            if (!lights.empty()) {
                uint32_t li; float lpdf; cdf_get(rand_f(seed), li, lpdf);
                const LightTri& l = lights[li];
                const float u = rand_f(seed), v = rand_f(seed) * (1.f - u);
                const V3 point = l.p0 + ((l.p1 - l.p0) * u) + ((l.p2 - l.p0) * v);
                V3 dir = point - p; const float dist = length(dir); dir /= dist;
                const float cos_out = fmaxf(0.f, dot(l.normal, -dir));
                if (dist > 0.01f && cos_out > 0.f) {
                    const float solid = (cos_out * l.area) / (dist * dist);
                    const float tr = ratio_transmittance(p, dir, 0.f, dist - 0.2f, seed);
                    const V3 c = T * (0.8f * (0.25f * kInvPi) * solid * (1.f / lpdf) * tr) * l.radiance;
                    srays.push_back({ray.px, ray.py, p, dir, dist - 0.2f, c, chan});
                }
            }
            if (bounce) {
                const float z = 1.f - 2.f * rand_f(seed);
                const float r = sqrtf(fmaxf(0.f, 1.f - z * z));
                const float phi = kTwoPi * rand_f(seed);
                float sp, cp; det_sincos(phi, sp, cp);
                next.push_back({ray.px, ray.py, p, v3(r * cp, r * sp, z), T * 0.8f});
            }
        }
    }

    // ------------------------------------------------------------------ frame: WaveFrontRenderer::TraceFrame, WaveFrontRenderer.cpp:435-1089
    void render_frame() {
        auto tic = std::chrono::steady_clock::now();
        auto lap = [&](const char* name) { auto now = std::chrono::steady_clock::now(); stats.push_back({name, std::chrono::duration<float, std::micro>(now - tic).count()}); tic = now; };
        stats.clear();
        if (scene_dirty) commit_scene();
        lap("scene");
        const uint32_t n = npix();
        const uint32_t stride = st.frame_count_stride ? st.frame_count_stride : 2u;
        const uint32_t frame_count = st.first_frame_count + 1u + stride * frame_index;      // hazard 10
        for (auto& c : channel) std::fill(c.begin(), c.end(), V4{0, 0, 0, 0});
        counters[0] = counters[1] = counters[2] = 0;
        std::vector<Ray> rays, next; std::vector<HitRec> hits; std::vector<VolumeHit> vhits;
        std::vector<std::vector<ShadowRay>> vol_groups;     // volumetric shadow rays resolve after the last wave (WaveFrontRenderer.cpp:855-871)
        raygen(frame_count, rays);
        lap("raygen");
        const uint32_t cur = surf_cur, prv = surf_cur == 1 ? 0 : 1;
        std::fill(surface[cur].begin(), surface[cur].end(), Surface{});
        uint32_t seed = wang_hash(frame_count);
        for (uint32_t depth = 0; depth < st.depth && !rays.empty(); ++depth) {
            extend(rays, hits); counters[0] += rays.size();
            extend_volumes(rays, hits, vhits);
            lap("extend");
            next.clear();
            std::vector<ShadowRay> srays, vrays;
            const bool delta = !vinstances.empty() && st.volume_mode == LB_VOLUME_DELTA;
            if (delta) { delta_scatter_all(rays, hits, vhits, seed, depth == 0 ? LB_CHANNEL_DIRECT : LB_CHANNEL_INDIRECT, depth < st.depth - 1, srays, next); lap("volume"); }
            if (depth == 0) { primary_hits = hits; for (auto& h : primary_hits) if (!(h.t > 0.f)) h = {0, 0, 0, 0, -1.f}; }
            std::vector<Surface>& S = surface[depth == 0 ? cur : 2];
            extract(rays, hits, S);
            lap("extract");
            ray_of_pixel.assign(n, -1); for (size_t i = 0; i < rays.size(); ++i) ray_of_pixel[(size_t)rays[i].py * st.width + rays[i].px] = (int)i;
            cur_rays_by_pixel = &rays;
            std::fill(volhits.begin(), volhits.end(), VolumeHit{});
            for (size_t i = 0; i < rays.size(); ++i) if (vhits[i].vinst >= 0) volhits[(size_t)rays[i].py * st.width + rays[i].px] = vhits[i];
            if (depth == 0) { motion_vectors(S); lap("motion"); }
            if (depth == 0) {
                // ResolveDirectLightHits, GPUShadeDirect.cu:11-40
                for (uint32_t i = 0; i < n; ++i) if (S[i].flags & SURF_EMISSIVE) channel[LB_CHANNEL_DIRECT][i] = S[i].mat.color;
                if (st.restir) {
                    if (!lights.empty()) restir_run(S, surface[prv], seed);
                    lap("restir");
                    if (delta) resolve_shadow(srays);
                } else {
                    shade_direct_all(S, rays, seed, LB_CHANNEL_DIRECT, srays, vrays);
                    resolve_shadow(srays); vol_groups.push_back(vrays);
                    lap("nee");
                }
            } else {
                shade_direct_all(S, rays, seed, LB_CHANNEL_INDIRECT, srays, vrays);
                resolve_shadow(srays); vol_groups.push_back(vrays);
                lap("nee");
            }
            if (depth < st.depth - 1) {
                const uint32_t s2 = wang_hash(seed);
                std::vector<Ray> cand(rays.size()); std::vector<uint8_t> ok(rays.size());
                #pragma omp parallel for schedule(dynamic, 256)
                for (int64_t i = 0; i < (int64_t)rays.size(); ++i) {
                    const uint32_t pi = rays[i].py * st.width + rays[i].px;
                    ok[i] = owned(pi) && shade_indirect_pixel(S[pi], pi, s2, cand[i]) ? 1 : 0;
                }
                for (size_t i = 0; i < rays.size(); ++i) if (ok[i]) next.push_back(cand[i]);
                lap("bounce");
            }
            if (depth > 0) for (const Ray& r : rays) surface[2][(size_t)r.py * st.width + r.px] = Surface{};
            rays.swap(next);
            seed = wang_hash(seed);
        }
        for (auto& g : vol_groups) resolve_shadow(g);
        res_cur = res_cur == 1 ? 0 : 1;          // ReSTIR::SwapBuffers once per frame (hazard 13)
        // MergeOutputChannels, GPUMergeOutputChannels.cu:5-88 (fp32; progressive = fp32 sum / count)
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            const V4 d = channel[0][i], in = channel[1][i], sp = channel[2][i], vo = channel[3][i];
            V4 m = {(d.x + in.x) + sp.x, (d.y + in.y) + sp.y, (d.z + in.z) + sp.z, (d.w + in.w) + sp.w};
            const float a = vo.w;
            m = {m.x * (1.0f - a) + vo.x * a, m.y * (1.0f - a) + vo.y * a, m.z * (1.0f - a) + vo.z * a, m.w * (1.0f - a) + vo.w * a};
            if (st.blend_output) {
                V4& acc = accum[i]; acc = {acc.x + m.x, acc.y + m.y, acc.z + m.z, acc.w + m.w};
                const float inv = 1.0f / (float)(blend_count + 1);
                combined[i] = {acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv};
            } else { combined[i] = m; accum[i] = m; }
        }
        if (st.blend_output) ++blend_count; else blend_count = 1;
        write_ldr();
        lap("merge");
        double m[16]; camera_matrix(m); memcpy(prev_cam, m, sizeof m); have_prev_cam = true;   // Camera::UpdatePreviousFrameMatrix
        surf_cur = surf_cur == 1 ? 0 : 1; ++frame_index;
    }
    void shade_direct_all(const std::vector<Surface>& S, const std::vector<Ray>& rays, uint32_t seed, int chan, std::vector<ShadowRay>& srays, std::vector<ShadowRay>& vrays) {
        std::vector<ShadowRay> cand(rays.size()); std::vector<uint8_t> ok(rays.size());
        const bool has_vol = !vinstances.empty() && st.volume_mode == LB_VOLUME_COMPAT;
        #pragma omp parallel for schedule(dynamic, 256) if (!has_vol)
        for (int64_t i = 0; i < (int64_t)rays.size(); ++i) {
            const uint32_t pi = rays[i].py * st.width + rays[i].px;
            ok[i] = owned(pi) && shade_direct_pixel(S[pi], pi, seed, chan, has_vol ? &volhits[pi] : nullptr, &vrays, cand[i]) ? 1 : 0;
        }
        for (size_t i = 0; i < rays.size(); ++i) if (ok[i]) srays.push_back(cand[i]);
    }
    // WriteToOutput, GPUShadingKernels.cu:28-56 + make_color, LumenPT/vendor/Include/Cuda/cuda/helpers.h:35-66
    void write_ldr() {
        auto q = [](float c) { c = clampf(c, 0.f, 1.f); const float s = c < 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
                               const float x = clampf(s, 0.f, 1.f); const uint32_t v = (uint32_t)(x * 256.f); return (uint8_t)(v > 255u ? 255u : v); };
        for (size_t i = 0; i < combined.size(); ++i) { ldr[4 * i] = q(combined[i].x); ldr[4 * i + 1] = q(combined[i].y); ldr[4 * i + 2] = q(combined[i].z); ldr[4 * i + 3] = 255; }
    }
};

static thread_local std::string g_err;
static int fail(int code, const char* msg) { g_err = msg; return code; }
static void invert_affine(const float* m, float* inv) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g), id = 1.0 / det;
    const double r[9] = {(e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id, (f * g - d * i) * id, (a * i - c * g) * id, (c * d - a * f) * id, (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id};
    for (int k = 0; k < 3; ++k) { inv[k * 4] = (float)r[k * 3]; inv[k * 4 + 1] = (float)r[k * 3 + 1]; inv[k * 4 + 2] = (float)r[k * 3 + 2];
        inv[k * 4 + 3] = (float)-(r[k * 3] * m[3] + r[k * 3 + 1] * m[7] + r[k * 3 + 2] * m[11]); }
    inv[12] = inv[13] = inv[14] = 0; inv[15] = 1;
}

} // namespace lo

using namespace lo;
#define R_ ((lo::Renderer*)r)
#define CHECK_R if (!r) return fail(LB_ERR_INVALID_ARGUMENT, "null renderer")

extern "C" {

LB_API int lo_create(const LbSettings* s, LbRenderer* out) {
    if (!s || !out || !s->width || !s->height || !s->depth) return fail(LB_ERR_INVALID_ARGUMENT, "bad settings");
    if (s->band_full_height && (s->band_row0 + s->height > s->band_full_height || ((uint64_t)s->band_row0 * s->width) % 256u))
        return fail(LB_ERR_INVALID_ARGUMENT, "row band: band_row0 + height must fit band_full_height and band_row0 * width must be a multiple of 256");
    if (!s->band_full_height && s->band_row0) return fail(LB_ERR_INVALID_ARGUMENT, "band_row0 without band_full_height");
    if (s->band_own_rows && (s->band_own_row0 < s->band_row0 || s->band_own_row0 + s->band_own_rows > s->band_row0 + s->height))
        return fail(LB_ERR_INVALID_ARGUMENT, "band_own_row0 / band_own_rows must lie inside the rendered rows");
    auto* r = new lo::Renderer(); r->st = *s;
    Texture white; white.px = {255, 255, 255, 255}; Texture nrm; nrm.px = {128, 128, 255, 255};   // LM/Renderer/LumenRenderer.cpp:50-58
    r->textures.push_back(white); r->textures.push_back(nrm);
    r->resize(); *out = (LbRenderer)r; return LB_OK;
}
LB_API int lo_destroy(LbRenderer r) { delete R_; return LB_OK; }
LB_API const char* lo_last_error(void) { return g_err.c_str(); }
LB_API const char* lo_version(void) { return "lumen-oracle 0.1 (cpu)"; }

LB_API int lo_texture_create(LbRenderer r, const uint8_t* rgba8, uint32_t w, uint32_t h, int srgb, LbHandle* out) {
    CHECK_R; if (!rgba8 || !w || !h || !out) return fail(LB_ERR_INVALID_ARGUMENT, "bad texture");
    Texture t; t.w = w; t.h = h; t.srgb = srgb != 0; t.px.assign(rgba8, rgba8 + (size_t)w * h * 4);
    R_->textures.push_back(std::move(t)); *out = (LbHandle)R_->textures.size() - 1; return LB_OK;
}
LB_API int lo_material_create(LbRenderer r, const LbMaterialDesc* d, LbHandle* out) {
    CHECK_R; if (!d || !out) return fail(LB_ERR_INVALID_ARGUMENT, "null");
    if (!(d->roughness_factor > 0.f && d->roughness_factor <= 1.f)) return fail(LB_ERR_INVALID_ARGUMENT, "roughness must be in (0,1]");   // WaveFrontRenderer.cpp:1274-1284
    Material m; if (!R_->fill_material(m, *d)) return fail(LB_ERR_INVALID_HANDLE, "texture handle");
    R_->materials.push_back(m); *out = (LbHandle)R_->materials.size() - 1; return LB_OK;
}
LB_API int lo_material_update(LbRenderer r, LbHandle h, const LbMaterialDesc* d) {
    CHECK_R; if (h < 0 || h >= (LbHandle)R_->materials.size() || !d) return fail(LB_ERR_INVALID_HANDLE, "material");
    if (!(d->roughness_factor > 0.f && d->roughness_factor <= 1.f)) return fail(LB_ERR_INVALID_ARGUMENT, "roughness must be in (0,1]");
    if (!R_->fill_material(R_->materials[h], *d)) return fail(LB_ERR_INVALID_HANDLE, "texture handle");
    for (auto& p : R_->prims) if (p.material == h) R_->find_emissives(p);
    R_->scene_dirty = true; return LB_OK;
}
LB_API int lo_primitive_create(LbRenderer r, const LbPrimitiveDesc* d, LbHandle* out) {
    CHECK_R; if (!d || !out || !d->positions || !d->indices || !d->vertex_count || d->index_count % 3) return fail(LB_ERR_INVALID_ARGUMENT, "bad primitive");
    if (d->index_size != 2 && d->index_size != 4) return fail(LB_ERR_INVALID_ARGUMENT, "index size");
    if (d->material < 0 || d->material >= (LbHandle)R_->materials.size()) return fail(LB_ERR_INVALID_HANDLE, "material");
    Primitive p; p.material = d->material; const uint32_t n = d->vertex_count;
    p.pos.resize(n); p.uv.assign(n, V2{0, 0}); p.nrm.assign(n, V3{0, 0, 1}); p.tan.assign(n, V4{1, 0, 0, 1});
    auto at = [](const void* base, uint32_t stride, uint32_t i) { return (const float*)((const char*)base + (size_t)stride * i); };
    for (uint32_t i = 0; i < n; ++i) {
        const float* q = at(d->positions, d->position_stride ? d->position_stride : 12, i); p.pos[i] = {q[0], q[1], q[2]};
        if (d->uvs) { q = at(d->uvs, d->uv_stride ? d->uv_stride : 8, i); p.uv[i] = {q[0], q[1]}; }
        if (d->normals) { q = at(d->normals, d->normal_stride ? d->normal_stride : 12, i); p.nrm[i] = {q[0], q[1], q[2]}; }
        if (d->tangents) { q = at(d->tangents, d->tangent_stride ? d->tangent_stride : 16, i); p.tan[i] = {q[0], q[1], q[2], q[3]}; }
    }
    p.idx.resize(d->index_count);
    for (uint32_t i = 0; i < d->index_count; ++i) { p.idx[i] = d->index_size == 2 ? ((const uint16_t*)d->indices)[i] : ((const uint32_t*)d->indices)[i]; if (p.idx[i] >= n) return fail(LB_ERR_INVALID_ARGUMENT, "index out of range"); }
    R_->find_emissives(p);
    R_->prims.push_back(std::move(p)); *out = (LbHandle)R_->prims.size() - 1; return LB_OK;
}
LB_API int lo_mesh_create(LbRenderer r, const LbHandle* prims, uint32_t count, LbHandle* out) {
    CHECK_R; if (!prims || !count || !out) return fail(LB_ERR_INVALID_ARGUMENT, "bad mesh");
    Mesh m; for (uint32_t i = 0; i < count; ++i) { if (prims[i] < 0 || prims[i] >= (LbHandle)R_->prims.size()) return fail(LB_ERR_INVALID_HANDLE, "primitive"); m.prims.push_back(prims[i]); }
    R_->meshes.push_back(m); *out = (LbHandle)R_->meshes.size() - 1; return LB_OK;
}
LB_API int lo_volume_create(LbRenderer r, const LbVolumeDesc* d, LbHandle* out) {
    CHECK_R; if (!d || !out) return fail(LB_ERR_INVALID_ARGUMENT, "null");
    Volume v; v.nx = d->nx; v.ny = d->ny; v.nz = d->nz; v.lo = {d->bbox_min[0], d->bbox_min[1], d->bbox_min[2]}; v.hi = {d->bbox_max[0], d->bbox_max[1], d->bbox_max[2]};
    if (d->density) { if (!d->nx || !d->ny || !d->nz) return fail(LB_ERR_INVALID_ARGUMENT, "grid size"); v.density.assign(d->density, d->density + (size_t)d->nx * d->ny * d->nz); v.majorant = 0.f; for (float x : v.density) v.majorant = fmaxf(v.majorant, x); }
    R_->volumes.push_back(std::move(v)); *out = (LbHandle)R_->volumes.size() - 1; return LB_OK;
}
LB_API int lo_scene_add_mesh_instance(LbRenderer r, LbHandle mesh, const float* m16, const LbEmissiveness* em, LbHandle ov, LbHandle* out) {
    CHECK_R; if (mesh < 0 || mesh >= (LbHandle)R_->meshes.size()) return fail(LB_ERR_INVALID_HANDLE, "mesh");
    if (ov >= (LbHandle)R_->materials.size()) return fail(LB_ERR_INVALID_HANDLE, "material");
    Instance in{}; in.mesh = mesh; in.override_mat = ov;
    static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(in.m, m16 ? m16 : ident, sizeof in.m);
    in.em = em ? *em : LbEmissiveness{LB_EMISSION_ENABLED, {0, 0, 0}, 1.f};
    R_->instances.push_back(in); R_->scene_dirty = true; if (out) *out = (LbHandle)R_->instances.size() - 1; return LB_OK;
}
LB_API int lo_instance_set_transform(LbRenderer r, LbHandle i, const float* m16) { CHECK_R; if (i < 0 || i >= (LbHandle)R_->instances.size() || !m16) return fail(LB_ERR_INVALID_HANDLE, "instance"); memcpy(R_->instances[i].m, m16, 64); R_->scene_dirty = true; return LB_OK; }
LB_API int lo_instance_set_emissiveness(LbRenderer r, LbHandle i, const LbEmissiveness* em) { CHECK_R; if (i < 0 || i >= (LbHandle)R_->instances.size() || !em) return fail(LB_ERR_INVALID_HANDLE, "instance"); R_->instances[i].em = *em; R_->scene_dirty = true; return LB_OK; }
LB_API int lo_instance_set_override_material(LbRenderer r, LbHandle i, LbHandle m) { CHECK_R; if (i < 0 || i >= (LbHandle)R_->instances.size() || m >= (LbHandle)R_->materials.size()) return fail(LB_ERR_INVALID_HANDLE, "instance"); R_->instances[i].override_mat = m; R_->scene_dirty = true; return LB_OK; }
LB_API int lo_scene_add_volume_instance(LbRenderer r, LbHandle vol, const float* m16, float density, LbHandle* out) {
    CHECK_R; if (vol < 0 || vol >= (LbHandle)R_->volumes.size()) return fail(LB_ERR_INVALID_HANDLE, "volume");
    VolumeInstance vi{}; vi.volume = vol; vi.density = density;
    static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(vi.m, m16 ? m16 : ident, 64); invert_affine(vi.m, vi.inv);
    R_->vinstances.push_back(vi); if (out) *out = (LbHandle)R_->vinstances.size() - 1; return LB_OK;
}
LB_API int lo_scene_clear(LbRenderer r) { CHECK_R; R_->instances.clear(); R_->vinstances.clear(); R_->scene_dirty = true; return LB_OK; }
LB_API int lo_camera_set_pose(LbRenderer r, const float* p, const float* q) { CHECK_R; if (!p || !q) return fail(LB_ERR_INVALID_ARGUMENT, "null"); R_->cam_pos = {p[0], p[1], p[2]}; memcpy(R_->cam_q, q, 16); R_->cam_from_matrix = false; return LB_OK; }
LB_API int lo_camera_set_matrix(LbRenderer r, const float* m) { CHECK_R; if (!m) return fail(LB_ERR_INVALID_ARGUMENT, "null"); memcpy(R_->cam_m, m, 64); R_->cam_pos = {m[3], m[7], m[11]}; R_->cam_from_matrix = true; return LB_OK; }
LB_API int lo_camera_set_fov_y(LbRenderer r, float deg) { CHECK_R; if (!(deg > 0.f && deg < 180.f)) return fail(LB_ERR_INVALID_ARGUMENT, "fov"); R_->fov_y = deg; return LB_OK; }
LB_API int lo_camera_set_min_max_distance(LbRenderer r, float mn, float mx) { CHECK_R; if (!(mx > mn)) return fail(LB_ERR_INVALID_ARGUMENT, "min/max distance"); R_->cam_min_d = mn; R_->cam_max_d = mx; return LB_OK; }
// G-buffer side outputs: GPUExtractDepthData.cu:6-72, GPUExtractNRD_DLSSdata.cu:6-89 (half4 normal+roughness), GPUPostProcessingEffects.cu:13-50 (albedo)
LB_API int lo_read_gbuffer(LbRenderer r, float* depth, float* nr, float* albedo, size_t pixel_capacity) {
    CHECK_R; const auto& S = R_->surface[R_->surf_cur == 1 ? 0 : 1];
    if (pixel_capacity < S.size()) return fail(LB_ERR_INVALID_ARGUMENT, "buffer too small");
    const float mn = R_->cam_min_d, mx = R_->cam_max_d;
    for (size_t i = 0; i < S.size(); ++i) {
        const Surface& s = S[i]; const float t = s.t;
        if (depth) depth[i] = t < 0.f ? 0.f : (t - fminf(mn, t)) / (fmaxf(mx, t) - fminf(mn, t));
        if (nr) { nr[4 * i] = half_round(s.normal.x); nr[4 * i + 1] = half_round(s.normal.y); nr[4 * i + 2] = half_round(s.normal.z); nr[4 * i + 3] = half_round(unpack8(s.mat.params[0], 24)); }
        if (albedo) { albedo[4 * i] = s.mat.color.x; albedo[4 * i + 1] = s.mat.color.y; albedo[4 * i + 2] = s.mat.color.z; albedo[4 * i + 3] = s.mat.color.w; }
    }
    return LB_OK;
}
LB_API int lo_set_render_resolution(LbRenderer r, uint32_t w, uint32_t h) { CHECK_R; if (!w || !h) return fail(LB_ERR_INVALID_ARGUMENT, "resolution"); R_->st.width = w; R_->st.height = h; R_->resize(); return LB_OK; }
LB_API int lo_get_settings(LbRenderer r, LbSettings* out) { CHECK_R; if (!out) return fail(LB_ERR_INVALID_ARGUMENT, "null"); *out = R_->st; return LB_OK; }
LB_API int lo_get_render_resolution(LbRenderer r, uint32_t* w, uint32_t* h) { CHECK_R; *w = R_->st.width; *h = R_->st.height; return LB_OK; }
LB_API int lo_set_depth(LbRenderer r, uint32_t d) { CHECK_R; if (!d) return fail(LB_ERR_INVALID_ARGUMENT, "depth"); R_->st.depth = d; return LB_OK; }
LB_API int lo_set_blend_mode(LbRenderer r, int b) { CHECK_R; R_->st.blend_output = b != 0; R_->blend_count = 0; std::fill(R_->accum.begin(), R_->accum.end(), V4{0, 0, 0, 0}); return LB_OK; }
LB_API int lo_get_blend_mode(LbRenderer r, int* b) { CHECK_R; *b = (int)R_->st.blend_output; return LB_OK; }
LB_API int lo_reset_history(LbRenderer r) { CHECK_R; R_->resize(); return LB_OK; }
LB_API int lo_render_frames(LbRenderer r, uint32_t frames) { CHECK_R; for (uint32_t i = 0; i < frames; ++i) R_->render_frame(); return LB_OK; }
LB_API int lo_synchronize(LbRenderer r) { CHECK_R; return LB_OK; }
LB_API int lo_start_rendering(LbRenderer) { return fail(LB_ERR_UNSUPPORTED, "oracle has no render thread"); }
LB_API int lo_stop_rendering(LbRenderer) { return fail(LB_ERR_UNSUPPORTED, "oracle has no render thread"); }

static int copy_out(const void* src, size_t bytes, void* dst, size_t cap) { if (!dst || cap < bytes) return fail(LB_ERR_INVALID_ARGUMENT, "buffer too small"); memcpy(dst, src, bytes); return LB_OK; }
LB_API int lo_read_hdr(LbRenderer r, float* out, size_t cap) { CHECK_R; return copy_out(R_->combined.data(), R_->combined.size() * 16, out, cap); }
// the CPU restatement has nothing to overlap: the asynchronous read-back is the synchronous one, the wait a no-op
LB_API int lo_read_hdr_async(LbRenderer r, float* out, size_t cap) { return lo_read_hdr(r, out, cap); }
LB_API int lo_readback_wait(LbRenderer r) { CHECK_R; return LB_OK; }
LB_API int lo_read_ldr(LbRenderer r, uint8_t* out, size_t cap) { CHECK_R; return copy_out(R_->ldr.data(), R_->ldr.size(), out, cap); }
LB_API int lo_read_channel(LbRenderer r, int c, float* out, size_t cap) { CHECK_R; if (c < 0 || c >= 4) return fail(LB_ERR_INVALID_ARGUMENT, "channel"); return copy_out(R_->channel[c].data(), R_->channel[c].size() * 16, out, cap); }
LB_API int lo_read_motion_vectors(LbRenderer r, float* out, size_t cap) { CHECK_R; return copy_out(R_->motion.data(), R_->motion.size() * 8, out, cap); }
LB_API int lo_frame_stats(LbRenderer r, const char** names, float* micros, uint32_t cap, uint32_t* count) {
    CHECK_R; R_->stats_names.clear(); uint32_t n = 0;
    for (auto& s : R_->stats) { if (n >= cap) break; if (n) R_->stats_names += ';'; R_->stats_names += s.first; micros[n++] = s.second; }
    if (names) *names = R_->stats_names.c_str(); if (count) *count = n; return LB_OK;
}
LB_API int lo_frame_counters(LbRenderer r, uint64_t* v, uint32_t cap, uint32_t* count) { CHECK_R; const uint32_t n = cap < 8 ? cap : 8; memcpy(v, R_->counters, n * 8); if (count) *count = n; return LB_OK; }
// ---- output stage mirror (test infrastructure): PNG with STORED deflate blocks — an encoder independent of the product's, so that the
// decoded pixels of both files can be compared; FrameStats JSON in the same shape.
static uint32_t crc32_bytes(const uint8_t* p, size_t n) { uint32_t c = ~0u; for (size_t i = 0; i < n; ++i) { c ^= p[i]; for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u))); } return ~c; }
static void be32(std::vector<uint8_t>& f, uint32_t v) { f.push_back(v >> 24); f.push_back(v >> 16); f.push_back(v >> 8); f.push_back(v); }
static void chunk(std::vector<uint8_t>& f, const char* type, const std::vector<uint8_t>& d) {
    be32(f, (uint32_t)d.size()); const size_t at = f.size(); f.insert(f.end(), type, type + 4); f.insert(f.end(), d.begin(), d.end()); be32(f, crc32_bytes(f.data() + at, d.size() + 4));
}
LB_API int lo_save_png(LbRenderer r, const char* path) {
    CHECK_R; if (!path || !*path) return fail(LB_ERR_INVALID_ARGUMENT, "path");
    const uint32_t w = R_->st.width, h = R_->st.height; const size_t stride = (size_t)w * 4;
    std::vector<uint8_t> raw; raw.reserve((stride + 1) * h);
    for (uint32_t y = 0; y < h; ++y) { raw.push_back(0); raw.insert(raw.end(), R_->ldr.begin() + (size_t)y * stride, R_->ldr.begin() + (size_t)(y + 1) * stride); }
    std::vector<uint8_t> z = {0x78, 0x01}; uint32_t a = 1, b = 0;
    for (size_t at = 0; at < raw.size() || at == 0; at += 65535) {
        const size_t n = std::min<size_t>(65535, raw.size() - at); const bool last = at + n >= raw.size();
        z.push_back(last ? 1 : 0); z.push_back(n & 255); z.push_back(n >> 8); z.push_back(~n & 255); z.push_back((~n >> 8) & 255);
        z.insert(z.end(), raw.begin() + at, raw.begin() + at + n);
        if (last) break;
    }
    for (uint8_t v : raw) { a = (a + v) % 65521u; b = (b + a) % 65521u; }
    be32(z, (b << 16) | a);
    std::vector<uint8_t> f = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, ihdr;
    be32(ihdr, w); be32(ihdr, h); ihdr.insert(ihdr.end(), {8, 6, 0, 0, 0});
    chunk(f, "IHDR", ihdr); chunk(f, "IDAT", z); chunk(f, "IEND", {});
    FILE* fp = fopen(path, "wb"); if (!fp) return fail(LB_ERR_INVALID_ARGUMENT, "cannot write file");
    const bool ok = fwrite(f.data(), 1, f.size(), fp) == f.size(); fclose(fp);
    return ok ? LB_OK : fail(LB_ERR_INVALID_ARGUMENT, "short write");
}
LB_API int lo_frame_stats_json(LbRenderer r, char* json, size_t cap, size_t* needed) {
    CHECK_R;
    static const char* names[8] = {"extend_rays", "shadow_rays", "visibility_rays", "kernel_launches", "lights", "triangles", "bvh_nodes", "bvh_bytes"};
    std::vector<std::pair<std::string, double>> times;
    for (auto& s : R_->stats) { bool found = false; for (auto& t : times) if (t.first == s.first) { t.second += s.second; found = true; } if (!found) times.push_back({s.first, s.second}); }
    char num[64];
    std::string o = "{\"frame_id\": " + std::to_string(R_->frame_index) + ", \"resolution\": [" + std::to_string(R_->st.width) + ", " + std::to_string(R_->st.height) + "], \"times_us\": {";
    for (size_t i = 0; i < times.size(); ++i) { snprintf(num, sizeof num, "%.3f", times[i].second); o += std::string(i ? ", \"" : "\"") + times[i].first + "\": " + num; }
    o += "}, \"counters\": {";
    for (int i = 0; i < 8; ++i) o += std::string(i ? ", \"" : "\"") + names[i] + "\": " + std::to_string(R_->counters[i]);
    o += "}}";
    if (needed) *needed = o.size() + 1;
    if (!json || cap < o.size() + 1) return (json || cap) ? fail(LB_ERR_INVALID_ARGUMENT, "buffer too small") : (needed ? (int)LB_OK : fail(LB_ERR_INVALID_ARGUMENT, "null"));
    memcpy(json, o.c_str(), o.size() + 1); return LB_OK;
}
LB_API int lo_hdr_buffer(LbRenderer r, void** p, size_t* bytes) { CHECK_R; if (!p || !bytes) return fail(LB_ERR_INVALID_ARGUMENT, "null"); *p = R_->combined.data(); *bytes = R_->combined.size() * 16; return LB_OK; }
LB_API int lo_accum_buffer(LbRenderer r, void** p, size_t* bytes, uint32_t* frames) { CHECK_R; *p = R_->accum.data(); *bytes = R_->accum.size() * 16; *frames = R_->blend_count; return LB_OK; }
LB_API int lo_resolve_accum(LbRenderer r, uint32_t total) { CHECK_R; if (!total) return fail(LB_ERR_INVALID_ARGUMENT, "frames"); const float inv = 1.0f / (float)total;
    for (size_t i = 0; i < R_->accum.size(); ++i) R_->combined[i] = {R_->accum[i].x * inv, R_->accum[i].y * inv, R_->accum[i].z * inv, R_->accum[i].w * inv}; R_->write_ldr(); return LB_OK; }
// ---- known-answer taps of the ReSTIR data structures (oracle only; checked against the reference's own ReSTIRData.h compiled for the host,
// oracle/ref_shim/ref_restir.cpp -> tests/golden/restir_reference.npz). Same signatures as ref_kat_reservoir / ref_kat_cdf.
LB_API void lo_kat_use_libm_sincos(int on) { lo::g_libm_sincos = on != 0; }
LB_API void lo_kat_rand_right_to_left(int on) { g_kat_rand_right_to_left = on != 0; }
// ---- known-answer taps of the radiance-deciding functions (oracle only; checked against the reference's own Resample / CombineBiased /
// CombineUnbiased / ShadeIndirect / ShadeDirect compiled for the host in place, oracle/ref_shim/ref_kernels.cpp ->
// tests/golden/kernels_reference.npz). Same signatures and flat layouts as ref_kat_* there.
static Mat unpack_mat24(const float* m);
static Surface kat_surface(const float* s, uint32_t px, uint32_t py) {
    Surface d; d.px = px; d.py = py;
    d.pos = {s[0], s[1], s[2]}; d.normal = {s[3], s[4], s[5]}; d.gnormal = d.normal; d.tangent = {s[6], s[7], s[8]}; d.incoming = {s[9], s[10], s[11]};
    d.transport = {s[12], s[13], s[14]}; d.t = s[15]; d.flags = (uint32_t)s[16]; d.mat = unpack_mat24(s + 20);
    return d;
}
static LightSample kat_sample(const float* s) {
    LightSample l; l.radiance = {s[0], s[1], s[2]}; l.normal = {s[3], s[4], s[5]}; l.position = {s[6], s[7], s[8]}; l.area = s[9];
    l.contribution = {s[10], s[11], s[12]}; l.pdf = s[13];
    return l;
}
static void kat_put_sample(const LightSample& l, float* o) {
    o[0] = l.radiance.x; o[1] = l.radiance.y; o[2] = l.radiance.z; o[3] = l.normal.x; o[4] = l.normal.y; o[5] = l.normal.z;
    o[6] = l.position.x; o[7] = l.position.y; o[8] = l.position.z; o[9] = l.area; o[10] = l.contribution.x; o[11] = l.contribution.y; o[12] = l.contribution.z; o[13] = l.pdf;
}
LB_API void lo_kat_resample(const float* samples14, unsigned n, const float* surf44, float* out14) {
    const Surface px = kat_surface(surf44, 0, 0);
    for (unsigned k = 0; k < n; ++k) { LightSample out; Renderer::resample(kat_sample(samples14 + 14 * k), px, out); kat_put_sample(out, out14 + 14 * k); }
}
LB_API void lo_kat_combine(const float* res17, unsigned n, const float* surf44, const float* surfs44, unsigned seed, int unbiased, float* out17) {
    const Surface px = kat_surface(surf44, 0, 0);
    std::vector<Reservoir> in(n); std::vector<Surface> from(n);
    for (unsigned k = 0; k < n; ++k) {
        const float* r = res17 + 17 * k; in[k].weight_sum = r[0]; in[k].count = (long long)r[1]; in[k].weight = r[2]; in[k].sample = kat_sample(r + 3);
        if (unbiased) from[k] = kat_surface(surfs44 + 44 * k, 0, 0);
    }
    Reservoir out;
    if (unbiased) Renderer::combine_unbiased(out, px, (int)n, in.data(), from.data(), seed); else Renderer::combine_biased(out, (int)n, in.data(), px, seed);
    out17[0] = out.weight_sum; out17[1] = (float)out.count; out17[2] = out.weight; kat_put_sample(out.sample, out17 + 3);
}
static void kat_put_reservoir(const Reservoir& r, float* o) { o[0] = r.weight_sum; o[1] = (float)r.count; o[2] = r.weight; kat_put_sample(r.sample, o + 3); }
static Reservoir kat_reservoir(const float* r) { Reservoir q; q.weight_sum = r[0]; q.count = (long long)r[1]; q.weight = r[2]; q.sample = kat_sample(r + 3); return q; }
static std::vector<Surface> kat_surfaces(const float* surfs44, unsigned w, unsigned h) {
    std::vector<Surface> s((size_t)w * h);
    for (unsigned i = 0; i < w * h; ++i) s[i] = kat_surface(surfs44 + 44 * (size_t)i, i % w, i / w);
    return s;
}
static std::vector<Reservoir> kat_reservoirs(const float* r17, unsigned n) { std::vector<Reservoir> r(n); for (unsigned i = 0; i < n; ++i) r[i] = kat_reservoir(r17 + 17 * (size_t)i); return r; }
static void kat_lights(Renderer& R, const float* lights16, const float* cdf_weights, unsigned nlights) {
    float run = 0.f;
    for (unsigned k = 0; k < nlights; ++k) {
        const float* l = lights16 + 16 * k; LightTri t;
        t.p0 = {l[0], l[1], l[2]}; t.p1 = {l[3], l[4], l[5]}; t.p2 = {l[6], l[7], l[8]}; t.normal = {l[9], l[10], l[11]}; t.radiance = {l[12], l[13], l[14]}; t.area = l[15];
        R.lights.push_back(t); run += cdf_weights[k]; R.cdf.push_back(run);
    }
    R.cdf_sum = run;
}
// whole ReSTIR stages over a w x h frame (signatures of ref_kat_ris / _visibility_rays / _temporal / _spatial / _combine_buffers; `unbiased` selects
// LbSettings::restir_unbiased where the reference build has a branch for it)
LB_API void lo_kat_ris(const float* surfs44, unsigned w, unsigned h, unsigned bag_seed, unsigned ris_seed, const float* lights16, const float* cdf_weights, unsigned nlights,
                       float* bag_pdf_out, float* bag_p0x_out, float* reservoirs_out) {
    Renderer R; R.st.width = w; R.st.height = h; kat_lights(R, lights16, cdf_weights, nlights);
    R.fill_bags(bag_seed);
    for (size_t i = 0; i < R.bags.size(); ++i) { bag_pdf_out[i] = R.bags[i].pdf; bag_p0x_out[i] = R.lights[R.bags[i].light].p0.x; }
    const std::vector<Surface> s = kat_surfaces(surfs44, w, h); std::vector<Reservoir> res((size_t)w * h);
    R.ris_stage(s, res, ris_seed);
    for (size_t i = 0; i < res.size(); ++i) kat_put_reservoir(res[i], reservoirs_out + 17 * i);
}
LB_API unsigned lo_kat_visibility_rays(const float* surfs44, const float* reservoirs17, unsigned w, unsigned h, float* rays8) {
    const std::vector<Surface> s = kat_surfaces(surfs44, w, h); const std::vector<Reservoir> res = kat_reservoirs(reservoirs17, w * h);
    unsigned n = 0;
    for (unsigned i = 0; i < w * h; ++i) {
        V3 d; float tmax;
        if (!Renderer::visibility_ray(s[i], res[i], d, tmax)) continue;
        float* o = rays8 + 8 * (size_t)n++; o[0] = (float)i; o[1] = s[i].pos.x; o[2] = s[i].pos.y; o[3] = s[i].pos.z; o[4] = d.x; o[5] = d.y; o[6] = d.z; o[7] = tmax;
    }
    return n;
}
LB_API void lo_kat_temporal(const float* cur44, const float* prev44, const float* cur17, const float* prev17, const float* motion2, unsigned w, unsigned h, unsigned seed, int unbiased,
                            float* cur_out17, float* direct4) {
    Renderer R; R.st.width = w; R.st.height = h; R.st.restir_unbiased = unbiased ? 1u : 0u;
    const std::vector<Surface> sc = kat_surfaces(cur44, w, h), sp = kat_surfaces(prev44, w, h);
    std::vector<Reservoir> rc = kat_reservoirs(cur17, w * h); const std::vector<Reservoir> rp = kat_reservoirs(prev17, w * h);
    std::vector<V2> mv((size_t)w * h); for (size_t i = 0; i < mv.size(); ++i) mv[i] = {motion2[2 * i], motion2[2 * i + 1]};
    std::vector<V4> direct((size_t)w * h); for (size_t i = 0; i < direct.size(); ++i) direct[i] = {direct4[4 * i], direct4[4 * i + 1], direct4[4 * i + 2], direct4[4 * i + 3]};
    R.temporal_stage(sc, sp, rc, rp, mv, direct, seed, 3.f);
    for (size_t i = 0; i < rc.size(); ++i) { kat_put_reservoir(rc[i], cur_out17 + 17 * i); direct4[4 * i] = direct[i].x; direct4[4 * i + 1] = direct[i].y; direct4[4 * i + 2] = direct[i].z; direct4[4 * i + 3] = direct[i].w; }
}
LB_API void lo_kat_spatial(const float* surfs44, const float* in17, unsigned w, unsigned h, unsigned seed, int unbiased, float* out17) {
    Renderer R; R.st.width = w; R.st.height = h; R.st.restir_unbiased = unbiased ? 1u : 0u;
    const std::vector<Surface> s = kat_surfaces(surfs44, w, h); const std::vector<Reservoir> in = kat_reservoirs(in17, w * h); std::vector<Reservoir> out = kat_reservoirs(out17, w * h);
    R.spatial_pass(s, in, out, seed);
    for (size_t i = 0; i < out.size(); ++i) kat_put_reservoir(out[i], out17 + 17 * i);
}
LB_API void lo_kat_combine_buffers(const float* surfs44, float* a17, const float* b17, unsigned w, unsigned h, unsigned seed) {
    Renderer R; R.st.width = w; R.st.height = h;
    const std::vector<Surface> s = kat_surfaces(surfs44, w, h); std::vector<Reservoir> a = kat_reservoirs(a17, w * h); const std::vector<Reservoir> b = kat_reservoirs(b17, w * h);
    R.combine_buffers(s, a, b, seed);
    for (size_t i = 0; i < a.size(); ++i) kat_put_reservoir(a[i], a17 + 17 * i);
}
LB_API unsigned lo_kat_shade_indirect(const float* surfs44, unsigned w, unsigned h, unsigned seed, float* rays11, unsigned cap) {
    Renderer R; R.st.width = w; R.st.height = h;
    unsigned n = 0;
    for (unsigned i = 0; i < w * h; ++i) {
        Ray ray;
        if (!R.shade_indirect_pixel(kat_surface(surfs44 + 44 * (size_t)i, i % w, i / w), i, seed, ray)) continue;
        if (n < cap) { float* o = rays11 + 11 * n; o[0] = (float)ray.px; o[1] = (float)ray.py; o[2] = ray.o.x; o[3] = ray.o.y; o[4] = ray.o.z; o[5] = ray.d.x; o[6] = ray.d.y; o[7] = ray.d.z;
            o[8] = ray.contrib.x; o[9] = ray.contrib.y; o[10] = ray.contrib.z; }
        ++n;
    }
    return n;
}
LB_API unsigned lo_kat_shade_direct(const float* surfs44, unsigned w, unsigned h, unsigned seed, const float* lights16, const float* cdf_weights, unsigned nlights,
                                    float* rays12, unsigned cap) {
    Renderer R; R.st.width = w; R.st.height = h;
    float run = 0.f;
    for (unsigned k = 0; k < nlights; ++k) {
        const float* l = lights16 + 16 * k; LightTri t;
        t.p0 = {l[0], l[1], l[2]}; t.p1 = {l[3], l[4], l[5]}; t.p2 = {l[6], l[7], l[8]}; t.normal = {l[9], l[10], l[11]}; t.radiance = {l[12], l[13], l[14]}; t.area = l[15];
        R.lights.push_back(t); run += cdf_weights[k]; R.cdf.push_back(run);
    }
    R.cdf_sum = run;
    unsigned n = 0;
    for (unsigned i = 0; i < w * h; ++i) {
        ShadowRay ray;
        if (!R.shade_direct_pixel(kat_surface(surfs44 + 44 * (size_t)i, i % w, i / w), i, seed, LB_CHANNEL_INDIRECT, nullptr, nullptr, ray)) continue;
        if (n < cap) { float* o = rays12 + 12 * n; o[0] = (float)ray.px; o[1] = (float)ray.py; o[2] = ray.o.x; o[3] = ray.o.y; o[4] = ray.o.z; o[5] = ray.d.x; o[6] = ray.d.y; o[7] = ray.d.z;
            o[8] = ray.tmax; o[9] = ray.radiance.x; o[10] = ray.radiance.y; o[11] = ray.radiance.z; }
        ++n;
    }
    return n;
}
LB_API void lo_kat_reservoir(const float* weights, const unsigned* seeds, const float* pdfs, unsigned n, float* out5, unsigned char* selected) {
    Reservoir r; r.sample.radiance.x = -1.f;
    for (unsigned k = 0; k < n; ++k) { LightSample s; s.pdf = pdfs[k]; s.radiance.x = (float)k; selected[k] = r.update(s, weights[k], seeds[k]) ? 1 : 0; }
    r.update_weight();
    out5[0] = r.weight_sum; out5[1] = (float)r.count; out5[2] = r.weight; out5[3] = r.sample.radiance.x; out5[4] = r.sample.pdf;
}
LB_API void lo_kat_cdf(const float* weights, unsigned n, const float* values, unsigned m, float* cdf_out, unsigned* index, float* pdf) {
    float sum = 0.f;
    for (unsigned k = 0; k < n; ++k) { sum += weights[k]; cdf_out[k] = sum; }          // CDF::Insert, ReSTIRData.h:194-203
    for (unsigned k = 0; k < m; ++k) { uint32_t i; float p; cdf_lookup(cdf_out, (int)n, sum, values[k], i, p); index[k] = i; pdf[k] = p; }
}
LB_API int lo_set_stream(LbRenderer, void*) { return LB_OK; }
LB_API int lo_set_overlap(LbRenderer r, int) { CHECK_R; return LB_OK; }      // the CPU restatement is one sequence of passes

LB_API int lo_debug_trace_closest(LbRenderer r, const float* rays6, uint32_t n, float tmin, float tmax, void* hits20) {
    CHECK_R; if (R_->scene_dirty) R_->commit_scene();
    HitRec* out = (HitRec*)hits20;
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        HitRec h; const V3 o = {rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]}, d = {rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]};
        if (R_->bvh.closest(o, d, tmin, tmax, h)) out[i] = h; else out[i] = {0, 0, 0, 0, -1.f};
    }
    return LB_OK;
}
LB_API int lo_debug_trace_any(LbRenderer r, const float* rays6, const float* tmaxs, uint32_t n, float tmin, uint8_t* occ) {
    CHECK_R; if (R_->scene_dirty) R_->commit_scene();
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; ++i) { const V3 o = {rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]}, d = {rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]}; occ[i] = R_->bvh.any(o, d, tmin, tmaxs[i]) ? 1 : 0; }
    return LB_OK;
}
// brute-force closest hit (oracle only): validates the oracle's own BVH
LB_API int lo_debug_trace_closest_brute(LbRenderer r, const float* rays6, uint32_t n, float tmin, float tmax, void* hits20) {
    CHECK_R; if (R_->scene_dirty) R_->commit_scene();
    HitRec* out = (HitRec*)hits20;
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        HitRec h; const V3 o = {rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]}, d = {rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]};
        if (closest_brute(R_->tris, o, d, tmin, tmax, h)) out[i] = h; else out[i] = {0, 0, 0, 0, -1.f};
    }
    return LB_OK;
}
LB_API int lo_debug_read_lights(LbRenderer r, float* l16, float* cdf, uint32_t cap, uint32_t* count) {
    CHECK_R; if (R_->scene_dirty) R_->commit_scene();
    const uint32_t n = (uint32_t)R_->lights.size(); if (count) *count = n; if (cap < n) return fail(LB_ERR_INVALID_ARGUMENT, "capacity");
    for (uint32_t i = 0; i < n; ++i) { if (l16) memcpy(l16 + 16 * i, &R_->lights[i], 64); if (cdf) cdf[i] = R_->cdf[i]; }
    return LB_OK;
}
LB_API int lo_debug_read_primary_hits(LbRenderer r, void* hits, size_t cap) { CHECK_R; return copy_out(R_->primary_hits.data(), R_->primary_hits.size() * 20, hits, cap); }
LB_API int lo_debug_read_surface(LbRenderer r, float* out, size_t cap) {
    CHECK_R; const auto& S = R_->surface[R_->surf_cur == 1 ? 0 : 1];   // the frame just rendered
    if (cap < S.size() * 96) return fail(LB_ERR_INVALID_ARGUMENT, "capacity");
    for (size_t i = 0; i < S.size(); ++i) { float* o = out + 24 * i; const Surface& s = S[i];
        o[0] = s.pos.x; o[1] = s.pos.y; o[2] = s.pos.z; o[3] = s.t; o[4] = s.normal.x; o[5] = s.normal.y; o[6] = s.normal.z; o[7] = (float)s.flags;
        o[8] = s.tangent.x; o[9] = s.tangent.y; o[10] = s.tangent.z; o[11] = 0; o[12] = s.incoming.x; o[13] = s.incoming.y; o[14] = s.incoming.z; o[15] = 0;
        o[16] = s.transport.x; o[17] = s.transport.y; o[18] = s.transport.z; o[19] = 0; o[20] = s.mat.color.x; o[21] = s.mat.color.y; o[22] = s.mat.color.z; o[23] = s.mat.color.w; }
    return LB_OK;
}
LB_API int lo_debug_read_reservoirs(LbRenderer r, float* out, size_t cap) {
    CHECK_R; const auto& Rv = R_->reservoirs[R_->res_cur == 1 ? 0 : 1];
    if (cap < Rv.size() * 80) return fail(LB_ERR_INVALID_ARGUMENT, "capacity");
    for (size_t i = 0; i < Rv.size(); ++i) { float* o = out + 20 * i; const Reservoir& q = Rv[i];
        o[0] = q.weight_sum; o[1] = q.weight; o[2] = (float)q.count; o[3] = q.sample.pdf; o[4] = q.sample.position.x; o[5] = q.sample.position.y; o[6] = q.sample.position.z; o[7] = q.sample.area;
        o[8] = q.sample.normal.x; o[9] = q.sample.normal.y; o[10] = q.sample.normal.z; o[11] = 0; o[12] = q.sample.radiance.x; o[13] = q.sample.radiance.y; o[14] = q.sample.radiance.z; o[15] = 0;
        o[16] = q.sample.contribution.x; o[17] = q.sample.contribution.y; o[18] = q.sample.contribution.z; o[19] = 0; }
    return LB_OK;
}
// mat24: color4, transmittance3, ior, tint3, luminance, metallic, subsurface, specular, roughness, spectint, anisotropic,
//        sheen, sheentint, clearcoat, clearcoatgloss, transmission, pad
static Mat unpack_mat24(const float* m) {
    Mat p; memset(&p, 0, sizeof p);
    p.color = {m[0], m[1], m[2], m[3]}; p.transmittance = {m[4], m[5], m[6], m[7]}; p.tint = {m[8], m[9], m[10], m[11]};
    pack8(p.params[0], m[12], 0); pack8(p.params[0], m[13], 8); pack8(p.params[0], m[14], 16); pack8(p.params[0], m[15], 24);
    pack8(p.params[1], m[16], 0); pack8(p.params[1], m[17], 8); pack8(p.params[1], m[18], 16); pack8(p.params[1], m[19], 24);
    pack8(p.params[2], m[20], 0); pack8(p.params[2], m[21], 8); pack8(p.params[2], m[22], 16);
    return p;
}
LB_API int lo_debug_eval_bsdf(LbRenderer, const float* mat24, const float* v12, uint32_t n, float* out4) {
    const Mat m = unpack_mat24(mat24);
    for (uint32_t i = 0; i < n; ++i) { const float* v = v12 + 12 * i; float pdf = 0;
        const V3 b = disney_eval(m, {v[0], v[1], v[2]}, {v[3], v[4], v[5]}, {v[6], v[7], v[8]}, {v[9], v[10], v[11]}, pdf);
        out4[4 * i] = b.x; out4[4 * i + 1] = b.y; out4[4 * i + 2] = b.z; out4[4 * i + 3] = pdf; }
    return LB_OK;
}
LB_API int lo_debug_sample_bsdf(LbRenderer, const float* mat24, const float* v12, uint32_t n, float* out8) {
    const Mat m = unpack_mat24(mat24);
    for (uint32_t i = 0; i < n; ++i) { const float* v = v12 + 12 * i; float pdf = 0; bool spec = false; V3 wi = v3(0);
        const V3 nrm = {v[0], v[1], v[2]};
        const V3 b = disney_sample(m, nrm, nrm, {v[3], v[4], v[5]}, {v[6], v[7], v[8]}, 1.f, v[9], v[10], v[11], wi, pdf, spec);
        float* o = out8 + 8 * i; o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = wi.x; o[4] = wi.y; o[5] = wi.z; o[6] = pdf; o[7] = spec ? 1.f : 0.f; }
    return LB_OK;
}
LB_API int lo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
LB_API void lo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}
