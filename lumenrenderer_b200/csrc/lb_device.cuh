// lb_device.cuh — device-side vocabulary of the B200 wavefront path tracer: vector maths, RNG, buffer layouts.
//
// Arithmetic contract (DESIGN.md "parity"): this library is compiled with -fmad=false, IEEE division and
// square root, so that every expression below evaluates exactly as written (no implicit contraction); fused
// operations appear only where fmaf() is written out. The same contract holds for the CPU oracle
// (-ffp-contract=off), which is what makes hit records bit-comparable.
//
// Reference semantics cited per item (paths relative to /root/reference/Lumen_Engine/LumenPT/):
//   vector helpers     vendor/Include/sutil/vec_math.h:454-561 (dot = left-to-right sum, normalize = v * (1/sqrt))
//   RNG                src/CUDAKernels/RandomUtilities.cuh:5-18
//   surface flags      src/Shaders/CppCommon/WaveFrontDataStructs/SurfaceData.h:17-23
//   material packing   src/Shaders/CppCommon/MaterialStructs.h:13-261
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define LB_HD __host__ __device__ __forceinline__
#define LB_D __device__ __forceinline__

namespace lb {

// ---------------------------------------------------------------- float3 algebra
LB_HD float3 f3(float a) { return make_float3(a, a, a); }
LB_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
LB_HD float3 f3(const float4& a) { return make_float3(a.x, a.y, a.z); }
LB_HD float4 f4(const float3& a, float w) { return make_float4(a.x, a.y, a.z, w); }
LB_HD float3 operator+(const float3& a, const float3& b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
LB_HD float3 operator-(const float3& a, const float3& b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
LB_HD float3 operator-(const float3& a) { return f3(-a.x, -a.y, -a.z); }
LB_HD float3 operator*(const float3& a, const float3& b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
LB_HD float3 operator*(const float3& a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
LB_HD float3 operator*(float s, const float3& a) { return f3(s * a.x, s * a.y, s * a.z); }
LB_HD float3 operator/(const float3& a, float s) { const float inv = 1.0f / s; return a * inv; }
LB_HD float3 operator+(const float3& a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
LB_HD float3 operator+(float s, const float3& a) { return f3(s + a.x, s + a.y, s + a.z); }
LB_HD float3& operator+=(float3& a, const float3& b) { a = a + b; return a; }
LB_HD float3& operator*=(float3& a, const float3& b) { a = a * b; return a; }
LB_HD float3& operator*=(float3& a, float s) { a = a * s; return a; }
LB_HD float3& operator/=(float3& a, float s) { a = a / s; return a; }
LB_HD float4 operator+(const float4& a, const float4& b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
LB_HD float4 operator*(const float4& a, const float4& b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
LB_HD float4 operator*(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
LB_HD float2 operator+(const float2& a, const float2& b) { return make_float2(a.x + b.x, a.y + b.y); }
LB_HD float2 operator*(const float2& a, float s) { return make_float2(a.x * s, a.y * s); }

LB_HD float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
LB_HD float3 cross(const float3& a, const float3& b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
LB_HD float length(const float3& v) { return sqrtf(dot(v, v)); }
LB_HD float3 normalize(const float3& v) { const float inv = 1.0f / sqrtf(dot(v, v)); return v * inv; }
LB_HD float3 reflect(const float3& i, const float3& n) { return i - 2.0f * n * dot(n, i); }
LB_HD float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
LB_HD float mixf(float a, float b, float t) { return a + t * (b - a); }
LB_HD float sq(float a) { return a * a; }
LB_HD float comp(const float3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// IEEE division, square root and normalisation spelled out. Exact-class code (ray generation, traversal, hit records, surface extraction,
// motion vectors: bit-compared with the oracle, DESIGN.md "Arithmetic contract") uses these, so that it does not depend on the
// -prec-div / -prec-sqrt setting of the translation unit it is inlined into (the fused shade kernel lives in a fast-math unit).
LB_D float xdiv(float a, float b) { return __fdiv_rn(a, b); }
LB_D float xsqrt(float a) { return __fsqrt_rn(a); }
LB_D float3 xnormalize(const float3& v) { const float inv = xdiv(1.0f, xsqrt(dot(v, v))); return v * inv; }
// sin and cos of an angle in [0, 2 pi] by Cody-Waite reduction to [-pi/4, pi/4] and the Cephes single-precision minimax polynomials
// (about 1 ulp), written with explicit fmaf only: the SAME sequence of IEEE operations in the CUDA library (lb_device.cuh) and in the
// oracle (lo_math.h), so the direction a bounce ray leaves in is bit-identical on both sides. A libm call is not: glibc's and libdevice's
// sinf differ in the last ulp, and a bounce direction is amplified at the next near-mirror vertex (canonical choice 16).
LB_D void det_sincos(float x, float& s, float& c) {
    const float j = floorf(fmaf(x, 0.636619772367581343f, 0.5f));                   // nearest multiple of pi/2: 0 .. 4
    float r = fmaf(-j, 1.5703125f, x);
    r = fmaf(-j, 4.837512969970703125e-4f, r);
    r = fmaf(-j, 7.54978995489188216e-8f, r);
    const float r2 = r * r;
    const float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
    const float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2 * r2, fmaf(-0.5f, r2, 1.0f));
    const int q = (int)j & 3;
    s = q == 0 ? sp : (q == 1 ? cp : (q == 2 ? -sp : -cp));
    c = q == 0 ? cp : (q == 1 ? -sp : (q == 2 ? -cp : sp));
}
// pow behind the clear-coat lobe's sampled direction: evaluated in double and rounded to float. Double-precision pow is accurate to well
// under one double ulp in both libdevice and glibc, so the float results agree on CPU and GPU except when the exact value lies within
// ~1e-16 (relative) of a float rounding boundary — and --use_fast_math does not touch double precision.
LB_D float xpow(float a, float b) { return (float)pow((double)a, (double)b); }

// fp16 round trip: barycentrics (IntersectionData.h:90) and motion vectors (MotionVectors.cu:44) are stored as half.
LB_D float half_round(float f) { return __half2float(__float2half_rn(f)); }

// ---------------------------------------------------------------- RNG (RandomUtilities.cuh:5-18)
LB_HD uint32_t wang_hash(uint32_t s) { s = (s ^ 61u) ^ (s >> 16); s *= 9u; s = s ^ (s >> 4); s *= 0x27d4eb2du; s = s ^ (s >> 15); return s; }
LB_HD uint32_t rand_u32(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
LB_HD float rand_f(uint32_t& s) { return (float)rand_u32(s) * 2.3283064365387e-10f; }

// ---------------------------------------------------------------- affine transforms: row-major 3x4, explicit fused chain
LB_HD float3 xform_point(const float* m, const float3& p) {
    return f3(fmaf(m[0], p.x, fmaf(m[1], p.y, fmaf(m[2], p.z, m[3]))),
              fmaf(m[4], p.x, fmaf(m[5], p.y, fmaf(m[6], p.z, m[7]))),
              fmaf(m[8], p.x, fmaf(m[9], p.y, fmaf(m[10], p.z, m[11]))));
}
LB_HD float3 xform_vector(const float* m, const float3& v) {
    return f3(fmaf(m[0], v.x, fmaf(m[1], v.y, m[2] * v.z)),
              fmaf(m[4], v.x, fmaf(m[5], v.y, m[6] * v.z)),
              fmaf(m[8], v.x, fmaf(m[9], v.y, m[10] * v.z)));
}

// ---------------------------------------------------------------- scene records (device)
enum : uint32_t { SURF_EMISSIVE = 1u, SURF_ALPHA = 2u, SURF_MISS = 4u };

// MaterialData (MaterialStructs.h:13-21): params.x = metallic|subsurface|specular|roughness (8 bit each, LSB first),
// params.y = spectint|anisotropic|sheen|sheentint, params.z = clearcoat|clearcoatgloss|transmission.
struct Material {
    float4 color, emissive, transmittance /* w = ior or eta */, tint /* w = luminance */;
    uint4 params;
};
struct DevMaterial {          // DeviceMaterial (ModelStructs.h) — material constants + 8 texture slots
    Material mat;
    int32_t tex_diffuse, tex_normal, tex_mr, tex_emissive, tex_transmission, tex_coat, tex_coat_rough, tex_tint;
};
struct DevTexture { uint32_t offset, w, h, srgb; };   // texels: RGBA8 in one pool (PTTexture.cpp:35-74: wrap, bilinear)

// One row of the scene data table = (mesh instance, primitive): DevicePrimitiveInstance (ModelStructs.h:70-77,
// PTMeshInstance.cpp:123-178). Vertex attributes stay in object space; m is the row-major world matrix.
struct DevEntry {
    float m[12];
    uint32_t index_base, vertex_base, tri_count, material;
    int32_t em_mode; float em_r, em_g, em_b, em_scale;
    uint32_t tri_offset;     // first global (entry-major) triangle number of this entry
    uint32_t lights_on;      // host decision of LightDataBuffer.cpp:37-125: does this entry contribute lights
    uint32_t flag_offset;    // first per-(primitive, triangle) emissive flag of this entry's primitive
};

// World-space triangle in BVH leaf order: 48 B = 3 x LDG.128. ids ride in the w lanes.
struct DevTri { float4 v0 /* w = instance(entry) bits */, v1 /* w = primitive index bits */, v2; };

// TriangleLight (LightData.h:21-27), 64 B
struct DevLight { float3 p0, p1, p2, normal, radiance; float area; };

// 8-wide compressed BVH node, 80 B (layout after Ylitie, Karras, Laine 2017):
//  q0: origin.xyz (f32), {ex, ey, ez, imask} bytes
//  q1: child_base, tri_base, meta[0..3], meta[4..7]
//  q2: qlo_x[0..7] | qlo_y[0..7]      q3: qlo_z[0..7] | qhi_x[0..7]      q4: qhi_y[0..7] | qhi_z[0..7]
struct Bvh8Node { uint4 q0, q1, q2, q3, q4; };

// warp scheduling of trace_queue (lb_trace.cuh): refill when >= refill_min lanes are idle; triangle round when pending * tri_quarter >= live
struct TraceTuning { int refill_min = 12; int tri_quarter = 4; };

struct BvhView {
    const Bvh8Node* nodes;
    const DevTri* tris;
    uint32_t num_tris;
    uint32_t* overflow;               // device counter of traversal-stack overflows (nullptr = not counted)
};

struct SceneView {
    const DevEntry* entries;
    const DevMaterial* materials;
    const DevTexture* textures;
    const uchar4* texels;
    const float* srgb_lut;            // 256 entries, computed on the host in double (bit-identical to the oracle)
    const uint32_t* indices;
    const float4* vtx_nu;             // normal.xyz, u
    const float4* vtx_tv;             // tangent.xyz, v
    const float* vtx_tw;              // tangent.w (bitangent sign)
    const DevLight* lights;
    const float* cdf;
    uint32_t num_lights;
    float cdf_sum;
};

// ---------------------------------------------------------------- per-pixel SoA planes (each plane: float4[npix])
// Surface (SurfaceData.h:49-104, 176 B AoS in the reference) as 9 coalesced 16-B planes:
//  0 position.xyz, flags(bits)   1 normal.xyz, t (sign bit set <=> flags != 0)   2 tangent.xyz, -   3 incoming.xyz, -
//  4 transport.xyz, -            5 color.rgba   6 transmittance.xyz, eta   7 tint.xyz, luminance   8 params (uint4 bits)
// Plane 1 alone answers the similarity test of temporal / spatial reuse (normal, depth, "is a plain surface"): a neighbour
// probe is ONE 16-byte gather. t is never negative (0 on a miss), so its sign bit is free to carry "flagged".
constexpr int kSurfPlanes = 9;
// Reservoir (ReSTIRData.h:107-183, 80 B AoS) as 5 planes:
//  0 weightSum, weight, sampleCount(int bits), sample.pdf   1 position.xyz, area   2 normal.xyz, -   3 radiance.xyz, -
//  4 unshadowed contribution.xyz, -
constexpr int kResPlanes = 5;

struct Surface {
    float3 pos, normal, tangent, incoming, transport; float t; uint32_t flags; Material mat;
};
struct LightSample { float3 radiance, normal, position, contribution; float area, pdf; };
struct Reservoir { float weight_sum, weight; int count; LightSample s; };

// PHYSICAL LAYOUT of the surface and reservoir planes: planes are stored in PAIRS — {0,1} {2,3} {5,6} {7,8} of a surface, {0,1} {2,3} of a
// reservoir — as 32-byte records (the two float4 of ONE pixel side by side, pair k of pixel i at float4 index k * 2n + 2i), the odd plane
// (surface 4 = transport, reservoir 4 = unshadowed contribution) alone at the end. A pair is read and written with one 256-bit access
// (LDG.E.256 / STG.E.256, new with Blackwell). Why: the reuse kernels gather these records per lane from random neighbours; a 16-byte gather
// pulls a 32-byte sector and uses half of it, and the spatial pass turned out to be bound by the NUMBER of such requests (33 per pixel,
// profiles/r02_a_ab.md). Paired, a neighbour's reservoir is 2 requests instead of 4, its shading record 4 instead of 8, every sector
// fully used; the streaming kernels lose nothing (a warp's 32 pair reads are 1 KB contiguous).
struct __align__(32) Float8 { float4 a, b; };
LB_D Float8 ld2(const float4* p) { return *reinterpret_cast<const Float8*>(p); }
LB_D void st2(float4* p, const float4& a, const float4& b) { Float8 v; v.a = a; v.b = b; *reinterpret_cast<Float8*>(p) = v; }
LB_D size_t surf_pair(size_t n, int pair, size_t i) { return (size_t)pair * 2u * n + 2u * i; }     // pairs 0..3 = planes {0,1} {2,3} {5,6} {7,8}
LB_D size_t surf_at(size_t n, int plane, size_t i) {
    if (plane == 4) return 8u * n + i;
    const int q = plane < 4 ? plane : plane - 1;
    return (size_t)(q >> 1) * 2u * n + 2u * i + (size_t)(q & 1);
}
LB_D size_t res_pair(size_t n, int pair, size_t i) { return (size_t)pair * 2u * n + 2u * i; }      // pairs 0..1 = planes {0,1} {2,3}
LB_D size_t res_at(size_t n, int plane, size_t i) { return plane == 4 ? 4u * n + i : (size_t)(plane >> 1) * 2u * n + 2u * i + (size_t)(plane & 1); }

LB_D void surface_store(float4* planes, size_t n, size_t i, const Surface& s) {
    st2(planes + surf_pair(n, 0, i), f4(s.pos, __uint_as_float(s.flags)), f4(s.normal, s.flags ? __uint_as_float(__float_as_uint(s.t) | 0x80000000u) : s.t));
    st2(planes + surf_pair(n, 1, i), f4(s.tangent, 0.f), f4(s.incoming, 0.f));
    planes[surf_at(n, 4, i)] = f4(s.transport, 0.f);
    st2(planes + surf_pair(n, 2, i), s.mat.color, s.mat.transmittance);
    st2(planes + surf_pair(n, 3, i), s.mat.tint, make_float4(__uint_as_float(s.mat.params.x), __uint_as_float(s.mat.params.y), __uint_as_float(s.mat.params.z), __uint_as_float(s.mat.params.w)));
}
LB_D uint32_t surface_flags(const float4* planes, size_t n, size_t i) { return __float_as_uint(planes[surf_at(n, 0, i)].w); }
// the similarity record of a pixel: normal, depth, flagged — one 16-byte load
struct SurfGeom { float3 normal; float t; bool flagged; };
LB_D SurfGeom surf_geom_unpack(const float4& b) {
    SurfGeom g; g.normal = f3(b); g.flagged = (__float_as_uint(b.w) >> 31) != 0u; g.t = fabsf(b.w);
    return g;
}
LB_D SurfGeom surface_geom(const float4* planes, size_t n, size_t i) { return surf_geom_unpack(planes[surf_at(n, 1, i)]); }
// everything a BSDF evaluation at the pixel needs (no path throughput): four 32-byte reads
LB_D void surface_load_shading(const float4* planes, size_t n, size_t i, Surface& s) {
    const Float8 p01 = ld2(planes + surf_pair(n, 0, i)), p23 = ld2(planes + surf_pair(n, 1, i)), p56 = ld2(planes + surf_pair(n, 2, i)), p78 = ld2(planes + surf_pair(n, 3, i));
    s.pos = f3(p01.a); s.flags = __float_as_uint(p01.a.w); s.normal = f3(p01.b); s.t = fabsf(p01.b.w);
    s.tangent = f3(p23.a); s.incoming = f3(p23.b);
    s.mat.color = p56.a; s.mat.transmittance = p56.b; s.mat.tint = p78.a;
    const float4 p = p78.b;
    s.mat.params = make_uint4(__float_as_uint(p.x), __float_as_uint(p.y), __float_as_uint(p.z), __float_as_uint(p.w));
    s.mat.emissive = make_float4(0.f, 0.f, 0.f, 0.f);
    s.transport = f3(0.f);
}
LB_D void surface_load(const float4* planes, size_t n, size_t i, Surface& s) {
    surface_load_shading(planes, n, i, s);
    s.transport = f3(planes[surf_at(n, 4, i)]);
}
LB_D void reservoir_store(float4* planes, size_t n, size_t i, const Reservoir& r) {
    st2(planes + res_pair(n, 0, i), make_float4(r.weight_sum, r.weight, __int_as_float(r.count), r.s.pdf), f4(r.s.position, r.s.area));
    st2(planes + res_pair(n, 1, i), f4(r.s.normal, 0.f), f4(r.s.radiance, 0.f));
    planes[res_at(n, 4, i)] = f4(r.s.contribution, 0.f);
}
LB_D void reservoir_load(const float4* planes, size_t n, size_t i, Reservoir& r) {
    const Float8 ab = ld2(planes + res_pair(n, 0, i)), cd = ld2(planes + res_pair(n, 1, i));
    r.weight_sum = ab.a.x; r.weight = ab.a.y; r.count = __float_as_int(ab.a.z); r.s.pdf = ab.a.w;
    r.s.position = f3(ab.b); r.s.area = ab.b.w;
    r.s.normal = f3(cd.a); r.s.radiance = f3(cd.b); r.s.contribution = f3(planes[res_at(n, 4, i)]);
}
LB_D Reservoir reservoir_zero() {
    Reservoir r; r.weight_sum = 0.f; r.weight = 0.f; r.count = 0;
    r.s.radiance = f3(0.f); r.s.normal = f3(0.f); r.s.position = f3(0.f); r.s.contribution = f3(0.f); r.s.area = 0.f; r.s.pdf = 0.f;
    return r;
}

// ---------------------------------------------------------------- wavefront queues
// Ray queue (IntersectionRayData.h:24-78, 40 B AoS) as 3 planes: o.xyz,- | d.xyz,pixel(bits) | throughput.xyz,-
// Shadow queue (ShadowRayData.h:13-62, 48 B AoS) as 3 planes: o.xyz,tmax | d.xyz,pixel(bits) | radiance.xyz,channel(bits)
// Hit record (IntersectionData.h:82-108, 16 B): {instance(entry), primitive, half2 barycentrics, t}; t < 0 = miss.
struct RayQueue { float4* o; float4* d; float4* T; };
struct ShadowQueue { float4* o; float4* d; float4* L; };

// Warp-aggregated queue append: one atomic per warp instead of one per ray (the reference's AtomicBuffer::Add does
// one atomicAdd per element, AtomicBuffer.h:26-28).
LB_D uint32_t queue_append_slot(uint32_t* counter) {
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

} // namespace lb
