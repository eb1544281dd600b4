// CPU ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path.
//
// Minimal vector maths with the same operation order as the reference's sutil/vec_math.h
// (/root/reference/Lumen_Engine/LumenPT/vendor/Include/sutil/vec_math.h:454-561): dot is a left-to-right
// sum of products, normalize multiplies by 1/sqrt(dot), vector / scalar multiplies by the reciprocal.
// Built with -ffp-contract=off so that no product/sum is fused unless fmaf() is written explicitly.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace lo {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

static inline V3 v3(float a) { return {a, a, a}; }
static inline V3 v3(float x, float y, float z) { return {x, y, z}; }
static inline V3 v3(const V4& a) { return {a.x, a.y, a.z}; }
static inline V4 v4(const V3& a, float w) { return {a.x, a.y, a.z, w}; }
static inline V4 v4(float a) { return {a, a, a, a}; }

static inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(const V3& a) { return {-a.x, -a.y, -a.z}; }
static inline V3 operator*(const V3& a, const V3& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator*(const V3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 operator/(const V3& a, float s) { const float inv = 1.0f / s; return a * inv; }
static inline V3 operator+(const V3& a, float s) { return {a.x + s, a.y + s, a.z + s}; }
static inline V3 operator+(float s, const V3& a) { return {s + a.x, s + a.y, s + a.z}; }
static inline V3& operator+=(V3& a, const V3& b) { a = a + b; return a; }
static inline V3& operator*=(V3& a, const V3& b) { a = a * b; return a; }
static inline V3& operator*=(V3& a, float s) { a = a * s; return a; }
static inline V3& operator/=(V3& a, float s) { a = a / s; return a; }

static inline V4 operator+(const V4& a, const V4& b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
static inline V4 operator*(const V4& a, const V4& b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
static inline V4 operator*(const V4& a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
static inline V2 operator+(const V2& a, const V2& b) { return {a.x + b.x, a.y + b.y}; }
static inline V2 operator*(const V2& a, float s) { return {a.x * s, a.y * s}; }

static inline float dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(const V3& a, const V3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline float length(const V3& v) { return sqrtf(dot(v, v)); }
static inline V3 normalize(const V3& v) { const float inv = 1.0f / sqrtf(dot(v, v)); return v * inv; }
static inline V3 reflect(const V3& i, const V3& n) { return i - 2.0f * n * dot(n, i); }
static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
static inline float mixf(float a, float b, float t) { return a + t * (b - a); }   // bsdf_math.cuh:16-19
static inline float sq(float a) { return a * a; }
static inline float comp(const V3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// fp16 round trip (reference stores barycentrics, motion vectors as half: IntersectionData.h:90, MotionVectors.cu:44)
static inline float half_round(float f) { return (float)(_Float16)f; }

// RNG, PT/CUDAKernels/RandomUtilities.cuh:5-18
static inline uint32_t wang_hash(uint32_t s) { s = (s ^ 61u) ^ (s >> 16); s *= 9u; s = s ^ (s >> 4); s *= 0x27d4eb2du; s = s ^ (s >> 15); return s; }
static inline uint32_t rand_u32(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
static inline float rand_f(uint32_t& s) { return (float)rand_u32(s) * 2.3283064365387e-10f; }

// Canonical affine transform (explicit fused chain, identical on the GPU): row-major 3x4 of a 4x4.
static inline V3 xform_point(const float* m, const V3& p) {
    return { fmaf(m[0], p.x, fmaf(m[1], p.y, fmaf(m[2], p.z, m[3]))),
             fmaf(m[4], p.x, fmaf(m[5], p.y, fmaf(m[6], p.z, m[7]))),
             fmaf(m[8], p.x, fmaf(m[9], p.y, fmaf(m[10], p.z, m[11]))) };
}
static inline V3 xform_vector(const float* m, const V3& v) {
    return { fmaf(m[0], v.x, fmaf(m[1], v.y, m[2] * v.z)),
             fmaf(m[4], v.x, fmaf(m[5], v.y, m[6] * v.z)),
             fmaf(m[8], v.x, fmaf(m[9], v.y, m[10] * v.z)) };
}

// sin and cos of an angle in [0, 2 pi]: Cody-Waite reduction to [-pi/4, pi/4] + the Cephes single-precision minimax polynomials (about
// 1 ulp), explicit fmaf only — operation for operation the sequence of lumenrenderer_b200/csrc/lb_device.cuh det_sincos, so that sampled
// bounce directions are bit-identical on CPU and GPU (glibc's and libdevice's sinf differ in the last ulp; canonical choice 16). The
// reference calls CUDA's sinf / cosf here (ggxmdf.cuh:90, disney.cuh cosine sampling); agreement with it is checked by the BSDF golden vectors.
// lo_kat_use_libm_sincos(1) (tests only) switches to glibc's sinf / cosf: with it the oracle's SampleBSDF is bit-identical to the host build
// of the reference headers, which pins everything around the two calls.
static bool g_libm_sincos = false;
static inline void det_sincos(float x, float& s, float& c) {
    if (g_libm_sincos) { s = sinf(x); c = cosf(x); return; }
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

// pow behind the clear-coat lobe's sampled direction, like lb_device.cuh xpow: double precision, rounded to float (glibc's and libdevice's
// double pow agree after that rounding); glibc's powf under the lo_kat_use_libm_sincos switch (what the reference headers' host build calls)
static inline float det_pow(float a, float b) { return g_libm_sincos ? powf(a, b) : (float)pow((double)a, (double)b); }

} // namespace lo
