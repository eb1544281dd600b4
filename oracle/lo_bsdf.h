// CPU ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path.
//
// Scalar restatement of the reference's Disney BSDF (appleseed -> Lighthouse2 lineage):
//   sample  : /root/reference/Lumen_Engine/LumenPT/src/CUDAKernels/disney.cuh:173-304
//   evaluate: disney.cuh:320-405; lobes disney.cuh:33-150
//   GGX / GTR1 microfacet distributions: ggxmdf.cuh:43-228
//   rough dielectric helpers: frosted.cuh:28-120
//   frame / cosine sampling: bsdf_math.cuh:57-176
//   8-bit parameter packing: Shaders/CppCommon/MaterialStructs.h:84-260
// Pinned against the reference headers themselves compiled for the host (oracle/_ref/libref_bsdf.so built by
// oracle/Makefile with the headers' DEVICE branches enabled; golden vectors in tests/golden/bsdf_reference.npz,
// tests/test_oracle_golden.py): EvaluateBSDF and SampleBSDF are bit-identical on 24 materials x 160 directions.
#pragma once
#include "lo_math.h"

namespace lo {

constexpr float kPi = 3.14159265358979323846264f;
constexpr float kInvPi = 0.31830988618379067153777f;
constexpr float kTwoPi = 6.28318530717958647692528f;
constexpr float kBsdfEps = 0.0001f;       // the EPSILON macro of bsdf_math.cuh:39-41

// MaterialData, MaterialStructs.h:13-21. params.x = metallic|subsurface|specular|roughness,
// params.y = spectint|anisotropic|sheen|sheentint, params.z = clearcoat|clearcoatgloss|transmission.
struct Mat {
    V4 color, emissive, transmittance /* w = ior */, tint /* w = luminance */;
    uint32_t params[4];
};

static inline float unpack8(uint32_t word, int shift) { return ((float)((word >> shift) & 255u)) * (1.0f / 255.0f); }
static inline void pack8(uint32_t& word, float v, int shift) {
    const uint32_t c = (uint32_t)(v * 255.f);
    word &= ~(255u << shift);
    word |= c << shift;
}
struct MatView {   // decoded once; every getter of the reference recomputes the same value
    V3 color; V3 transmittance; float ior; V3 tint; float luminance;
    float metallic, subsurface, specular, roughness, spectint, anisotropic, sheen, sheentint, clearcoat, clearcoatgloss, transmission;
    explicit MatView(const Mat& m) {
        color = v3(m.color); transmittance = v3(m.transmittance); ior = m.transmittance.w; tint = v3(m.tint); luminance = m.tint.w;
        metallic = unpack8(m.params[0], 0); subsurface = unpack8(m.params[0], 8); specular = unpack8(m.params[0], 16); roughness = unpack8(m.params[0], 24);
        spectint = unpack8(m.params[1], 0); anisotropic = unpack8(m.params[1], 8); sheen = unpack8(m.params[1], 16); sheentint = unpack8(m.params[1], 24);
        clearcoat = unpack8(m.params[2], 0); clearcoatgloss = unpack8(m.params[2], 8); transmission = unpack8(m.params[2], 16);
    }
};

// ---------------------------------------------------------------- microfacet distributions (ggxmdf.cuh)
static inline void alpha_from_roughness(float roughness, float anisotropy, float& ax, float& ay) {   // ggxmdf.cuh:221-227
    const float r2 = roughness * roughness;
    const float aspect = sqrtf(1.0f + anisotropy * (anisotropy < 0 ? 0.9f : -0.9f));
    ax = fmaxf(0.001f, r2 / aspect);
    ay = fmaxf(0.001f, r2 * aspect);
}
static inline float ggx_d(const V3& m, float ax, float ay) {                                          // ggxmdf.cuh:43-53
    if (m.z == 0) return sq(ax) * kInvPi;
    const float c2 = sq(m.z);
    const float s = sqrtf(fmaxf(0.0f, 1 - c2));
    const float t2 = (1.0f - c2) / c2;
    float stretched;
    if (ax == ay || s == 0.0f) stretched = 1.0f / sq(ax);
    else stretched = sq(m.x / (s * ax)) + sq(m.y / (s * ay));
    return 1.0f / (kPi * ax * ay * sq(c2) * sq(1.0f + t2 * stretched));
}
static inline float ggx_lambda(const V3& v, float ax, float ay) {                                     // ggxmdf.cuh:55-66
    if (v.z == 0) return 0;
    const float c2 = v.z * v.z;
    const float s = sqrtf(fmaxf(0.0f, 1 - c2));
    float projected;
    if (ax == ay || s == 0.0f) projected = ax;
    else projected = sqrtf(sq((v.x * ax) / s) + sq((v.y * ay) / s));
    const float t2 = sq(s) / c2;
    const float a2_rcp = sq(projected) * t2;
    return (-1.0f + sqrtf(1.0f + a2_rcp)) * 0.5f;
}
static inline float ggx_g(const V3& wi, const V3& wo, float ax, float ay) { return 1.0f / (1.0f + ggx_lambda(wo, ax, ay) + ggx_lambda(wi, ax, ay)); }
static inline float ggx_g1(const V3& v, float ax, float ay) { return 1.0f / (1.0f + ggx_lambda(v, ax, ay)); }
static inline float ggx_pdf(const V3& v, const V3& m, float ax, float ay) {                           // ggxmdf.cuh:142-152
    if (v.z == 0.0f) return 0;
    return ggx_g1(v, ax, ay) * fabsf(dot(v, m)) * ggx_d(m, ax, ay) / fabsf(v.z);
}
static inline V3 ggx_sample(const V3& v, float r0, float r1, float ax, float ay) {                    // ggxmdf.cuh:78-107 (device branch)
    const float sgn = v.z < 0.0f ? -1.0f : 1.0f;
    const V3 st = normalize(v3(sgn * v.x * ax, sgn * v.y * ay, sgn * v.z));
    const V3 t1 = v.z < 0.9999f ? normalize(cross(st, v3(0, 0, 1))) : v3(1, 0, 0);
    const V3 t2 = cross(t1, st);
    const float a = 1.0f / (1.0f + st.z);
    const float r = sqrtf(r0);
    const float phi = r1 < a ? (r1 / a * kPi) : (kPi + (r1 - a) / (1.0f - a) * kPi);
    float p1, p2; det_sincos(phi, p2, p1);
    p1 *= r;
    p2 *= r * (r1 < a ? 1.0f : st.z);
    const V3 h = p1 * t1 + p2 * t2 + sqrtf(fmaxf(0.0f, 1.0f - p1 * p1 - p2 * p2)) * st;
    return normalize(v3(h.x * ax, h.y * ay, fmaxf(0.0f, h.z)));
}
static inline float gtr1_alpha(float a) { return clampf(a, 0.001f, 0.999f); }
static inline float gtr1_d(const V3& m, float ax) {                                                   // ggxmdf.cuh:161-168
    const float a2 = sq(gtr1_alpha(ax));
    const float a = (a2 - 1.0f) / (kPi * logf(a2));
    const float b = (1 / (1 + (a2 - 1) * sq(m.z)));
    return a * b;
}
static inline float gtr1_lambda(const V3& v, float ax) {                                              // ggxmdf.cuh:170-186
    if (v.z == 0) return 0;
    const float c2 = sq(v.z);
    const float s = sqrtf(fmaxf(0.0f, 1.0f - c2));
    if (s == 0) return 0;
    const float cot2 = c2 / sq(s);
    const float cot = sqrtf(cot2);
    const float a2 = sq(gtr1_alpha(ax));
    const float a = sqrtf(cot2 + a2);
    const float b = sqrtf(cot2 + 1.0f);
    const float c = logf(cot + b);
    const float d = logf(cot + a);
    return (a - b + cot * (c - d)) / (cot * logf(a2));
}
static inline float gtr1_g(const V3& wi, const V3& wo, float ax) { return 1.0f / (1.0f + gtr1_lambda(wo, ax) + gtr1_lambda(wi, ax)); }
static inline float gtr1_pdf(const V3& m, float ax) { return gtr1_d(m, ax) * fabsf(m.z); }
static inline V3 gtr1_sample(float r0, float r1, float ax) {                                          // ggxmdf.cuh:198-213
    const float a2 = sq(gtr1_alpha(ax));
    const float c2 = (1.0f - det_pow(a2, 1.0f - r0)) / (1.0f - a2);
    const float s = sqrtf(fmaxf(0.0f, 1.0f - c2));
    const float phi = kTwoPi * r1;
    float cp, sp; det_sincos(phi, sp, cp);
    return v3(cp * s, sp * s, sqrtf(c2));
}

// ---------------------------------------------------------------- lobes (disney.cuh:33-150)
static inline float schlick_w(float u) { const float m = clampf(1.0f - u, 0.f, 1.f), m2 = sq(m), m4 = sq(m2); return m4 * m; }
static inline V3 spec_fresnel(const MatView& s, const V3& o, const V3& h) {                           // disney.cuh:38-45
    V3 v = (1.0f - s.spectint) + s.spectint * s.tint;
    v *= s.specular * 0.08f;
    v = (1.0f - s.metallic) * v + s.metallic * s.color;
    const float f = schlick_w(fabsf(dot(o, h)));
    return (1.0f - f) * v + f;
}
static inline V3 coat_fresnel(const MatView& s, const V3& o, const V3& h) {                           // disney.cuh:46-50
    return v3(mixf(0.04f, 1.0f, schlick_w(fabsf(dot(o, h)))) * 0.25f * s.clearcoat);
}
static inline float coat_roughness(const MatView& s) { return mixf(0.1f, 0.001f, s.clearcoatgloss); }

enum Mdf { MDF_GGX, MDF_GTR1 };
static inline float mf_d(Mdf k, const V3& m, float ax, float ay) { return k == MDF_GGX ? ggx_d(m, ax, ay) : gtr1_d(m, ax); }
static inline float mf_g(Mdf k, const V3& wi, const V3& wo, float ax, float ay) { return k == MDF_GGX ? ggx_g(wi, wo, ax, ay) : gtr1_g(wi, wo, ax); }
static inline float mf_pdf(Mdf k, const V3& v, const V3& m, float ax, float ay) { return k == MDF_GGX ? ggx_pdf(v, m, ax, ay) : gtr1_pdf(m, ax); }

// disney.cuh:78-99. component pdf is 0 when the reference leaves it unset (wil.z == 0; documented deviation).
static inline void mf_sample(Mdf k, const MatView& s, float r0, float r1, float ax, float ay, const V3& wol, V3& wil, float& pdf, V3& value) {
    if (wol.z == 0) { value = v3(0); pdf = 0; return; }
    const V3 m = k == MDF_GGX ? ggx_sample(wol, r0, r1, ax, ay) : gtr1_sample(r0, r1, ax);
    wil = reflect(wol * -1.0f, m);
    pdf = 0;
    if (wil.z == 0) return;
    const float cos_oh = dot(wol, m);
    pdf = mf_pdf(k, wol, m, ax, ay) / fabsf(4.0f * cos_oh);
    if (pdf < 1.0e-6f) return;
    const float D = mf_d(k, m, ax, ay), G = mf_g(k, wil, wol, ax, ay);
    value = k == MDF_GGX ? spec_fresnel(s, wol, m) : coat_fresnel(s, wol, m);
    value *= D * G;
}
// disney.cuh:101-113
static inline float mf_eval(Mdf k, const MatView& s, float ax, float ay, const V3& wol, const V3& wil, const V3& m, V3& bsdf) {
    if (wol.z == 0 || wil.z == 0) return 0;
    const float cos_oh = dot(wol, m);
    if (cos_oh == 0) return 0;
    const float D = mf_d(k, m, ax, ay), G = mf_g(k, wil, wol, ax, ay);
    bsdf = k == MDF_GGX ? spec_fresnel(s, wol, m) : coat_fresnel(s, wol, m);
    bsdf *= D * G / fabsf(4.0f * wol.z * wil.z);
    return mf_pdf(k, wol, m, ax, ay) / fabsf(4.0f * cos_oh);
}
// disney.cuh:115-138
static inline float diffuse_eval(const MatView& s, const V3& n, const V3& wo, const V3& wi, const V3& m, V3& value) {
    const float cos_on = dot(n, wo), cos_in = dot(n, wi), cos_ih = dot(wi, m);
    const float fl = schlick_w(cos_in), fv = schlick_w(cos_on);
    float fd = 0;
    if (s.subsurface != 1.0f) {
        const float fd90 = 0.5f + 2.0f * sq(cos_ih) * s.roughness;
        fd = mixf(1.f, fd90, fl) * mixf(1.f, fd90, fv);
    }
    if (s.subsurface > 0) {
        const float fss90 = sq(cos_ih) * s.roughness;
        const float fss = mixf(1.0f, fss90, fl) * mixf(1.0f, fss90, fv);
        const float ss = 1.25f * (fss * (1.0f / (fabsf(cos_on) + fabsf(cos_in)) - 0.5f) + 0.5f);
        fd = mixf(fd, ss, s.subsurface);
    }
    value = s.color * fd * kInvPi * (1.0f - s.metallic);
    return fabsf(cos_in) * kInvPi;
}
// disney.cuh:140-150
static inline float sheen_eval(const MatView& s, const V3& wi, const V3& m, V3& value) {
    const float fh = schlick_w(dot(wi, m));
    value = (1.0f - s.sheentint) + s.sheentint * s.tint;
    value *= fh * s.sheen * (1.0f - s.metallic);
    return 1.0f / (2 * kPi);
}

// ---------------------------------------------------------------- rough dielectric (frosted.cuh:28-120)
static inline float fresnel_dielectric(float cos_i, float eta, float& cos_t) {
    const float s2 = (1 - sq(cos_i)) * sq(eta);
    if (s2 > 1) { cos_t = 0; return 1; }
    cos_t = fminf(sqrtf(fmaxf(1 - s2, 0.0f)), 1.0f);
    const float ci = fabsf(cos_i);
    if (ci == 0 && cos_t == 0) return 1;
    const float k0 = eta * cos_t, k1 = eta * ci;
    return 0.5f * (sq((ci - k0) / (ci + k0)) + sq((cos_t - k1) / (cos_t + k1)));
}
static inline V3 refract_dir(const V3& wo, const V3& m, float cos_wom, float cos_t, float rcp_eta) {
    const V3 wi = cos_wom > 0 ? (rcp_eta * cos_wom - cos_t) * m - rcp_eta * wo
                              : (rcp_eta * cos_wom + cos_t) * m - rcp_eta * wo;
    return wi * ((3 - dot(wi, wi)) * 0.5f);
}
static inline float choose_reflection(float rw, float tw, float F) {
    const float r = F * rw, t = (1 - F) * tw, sum = r + t;
    return sum != 0 ? r / sum : 1;
}
static inline V3 half_reflect(const V3& wo, const V3& wi) { const V3 h = normalize(wi + wo); return h.z < 0 ? (h * -1.f) : h; }
static inline V3 half_refract(const V3& wo, const V3& wi, float eta) { const V3 h = normalize(wo + eta * wi); return h.z < 0 ? (h * -1.f) : h; }
static inline void reflection_eval(const V3& color, const V3& wo, const V3& wi, const V3& m, float ax, float ay, float F, V3& value) {
    const float denom = fabsf(4 * wo.z * wi.z);
    if (denom == 0) { value = v3(0); return; }
    const float D = ggx_d(m, ax, ay), G = ggx_g(wi, wo, ax, ay);
    value = color * (F * D * G / denom);
}
static inline void refraction_eval(float eta, const V3& color, bool adjoint, const V3& wo, const V3& wi, const V3& m, float ax, float ay, float T, V3& value) {
    if (wo.z == 0 || wi.z == 0) { value = v3(0); return; }
    const float cos_ih = dot(m, wi), cos_oh = dot(m, wo);
    const float dots = (cos_ih * cos_oh) / (wi.z * wo.z);
    const float sd = cos_oh + eta * cos_ih;
    if (fabsf(sd) < 1.0e-6f) { value = v3(0); return; }
    const float D = ggx_d(m, ax, ay), G = ggx_g(wi, wo, ax, ay);
    float mult = fabsf(dots) * T * D * G / sq(sd);
    if (!adjoint) mult *= sq(eta);
    value = color * mult;
}
static inline float reflection_jacobian(float cos_oh) { return cos_oh == 0 ? 0 : 1 / (4 * fabsf(cos_oh)); }
static inline float refraction_jacobian(const V3& wo, const V3& wi, const V3& m, float eta) {
    const float cos_ih = dot(m, wi), cos_oh = dot(m, wo);
    const float sd = cos_oh + eta * cos_ih;
    if (fabsf(sd) < 1.0e-6f) return 0;
    return fabsf(cos_ih) * sq(eta / sd);
}

static inline V3 to_local(const V3& v, const V3& n, const V3& t, const V3& b) { return v3(dot(v, t), dot(v, b), dot(v, n)); }
static inline V3 to_world(const V3& v, const V3& n, const V3& t, const V3& b) { return v.x * t + v.y * b + v.z * n; }
static inline V3 cos_weighted(float r0, float r1, const V3& n, const V3& t, const V3& b) {             // bsdf_math.cuh:133-139
    const float term1 = kTwoPi * r0, term2 = sqrtf(1 - r1);
    float s, c; det_sincos(term1, s, c);
    return (c * term2 * t) + (s * term2) * b + sqrtf(r1) * n;
}
static inline void lobe_weights(const MatView& s, float w[4]) {                                        // disney.cuh:228-229, 368-369
    w[0] = mixf(s.luminance, 0.f, s.metallic); w[1] = mixf(s.sheen, 0.f, s.metallic);
    w[2] = mixf(s.specular, 1.f, s.metallic);  w[3] = s.clearcoat * 0.25f;
    const float inv = 1.0f / (w[0] + w[1] + w[2] + w[3]);
    for (int i = 0; i < 4; ++i) w[i] *= inv;
}

// ---------------------------------------------------------------- EvaluateBSDF, disney.cuh:320-405
static inline V3 disney_eval(const Mat& mat, const V3& iN, const V3& iT, const V3& wow, const V3& wiw, float& pdf) {
    const MatView s(mat);
    V3 trans_bsdf = v3(0); float trans_pdf = 0.f;
    if (s.transmission > 0.f) {
        const V3 B = normalize(cross(iN, iT)), T = normalize(cross(iN, B));
        const V3 wol = to_local(wow, iN, T, B), wil = to_local(wiw, iN, T, B);
        const float eta = wol.z > 0 ? s.ior : (1.0f / s.ior);
        if (eta == 1) { pdf = 0; return v3(0); }
        float ax, ay, jac; alpha_from_roughness(s.roughness, s.anisotropic, ax, ay);
        V3 m;
        if (wil.z * wol.z >= 0) {
            m = half_reflect(wol, wil);
            const float cos_wom = dot(wol, m); float ct;
            const float F = fresnel_dielectric(cos_wom, 1 / eta, ct);
            reflection_eval(s.color, wol, wil, m, ax, ay, F, trans_bsdf);
            trans_pdf = choose_reflection(1, 1, F); jac = reflection_jacobian(cos_wom);
        } else {
            m = half_refract(wol, wil, eta);
            const float cos_wom = dot(wol, m); float ct;
            const float F = fresnel_dielectric(cos_wom, 1 / eta, ct);
            refraction_eval(eta, s.color, false, wol, wil, m, ax, ay, 1 - F, trans_bsdf);
            trans_pdf = 1 - choose_reflection(1, 1, F); jac = refraction_jacobian(wol, wil, m, eta);
        }
        trans_pdf *= jac * ggx_pdf(wol, m, ax, ay);
    }
    if (s.roughness <= 0.001f) { pdf = trans_pdf; return trans_bsdf; }
    const V3 B = normalize(cross(iN, iT)), T = normalize(cross(iN, B));
    float w[4]; lobe_weights(s, w);
    pdf = 0; V3 value = v3(0);
    if (w[0] + w[1] > 0) {
        const V3 m = normalize(wiw + wow);
        if (w[0] > 0) pdf += w[0] * diffuse_eval(s, iN, wow, wiw, m, value);
        if (w[1] > 0) pdf += w[1] * sheen_eval(s, wiw, m, value);      // overwrites the diffuse value, as the reference does (:377)
    }
    if (w[2] + w[3] > 0) {
        const V3 wol = to_local(wow, iN, T, B), wil = to_local(wiw, iN, T, B);
        const V3 m = normalize(wol + wil);
        if (w[2] > 0) {
            float ax, ay; alpha_from_roughness(s.roughness, s.anisotropic, ax, ay);
            V3 c = v3(0); const float p = mf_eval(MDF_GGX, s, ax, ay, wol, wil, m, c);
            if (p > 0) { pdf += w[2] * p; value += c; }
        }
        if (w[3] > 0) {
            const float a = coat_roughness(s);
            V3 c = v3(0); const float p = mf_eval(MDF_GTR1, s, a, a, wol, wil, m, c);
            if (p > 0) { pdf += w[3] * p; value += c; }
        }
    }
    pdf = (pdf * (1.f - s.transmission));
    pdf += (trans_pdf * s.transmission);
    return (trans_bsdf * s.transmission) + (value * (1.f - s.transmission));
}

// ---------------------------------------------------------------- SampleBSDF, disney.cuh:173-304
static inline V3 disney_sample(const Mat& mat, V3 iN, const V3& N, const V3& iT, const V3& wow, float distance,
                               float r0, float r1, float r2, V3& wiw, float& pdf, bool& specular) {
    const MatView s(mat);
    const float flip = (dot(wow, N) < 0) ? -1.f : 1.f;
    iN *= flip;
    const V3 B = normalize(cross(iN, iT)), T = normalize(cross(iN, B));
    if (r0 < s.transmission) {
        specular = true;
        const float r3 = r0 / s.transmission;
        const V3 wol = to_local(wow, iN, T, B);
        const float eta = flip < 0 ? (1 / s.ior) : s.ior;
        if (eta == 1) return v3(0);
        const V3 beer = v3(expf(-s.transmittance.x * distance * 2.0f), expf(-s.transmittance.y * distance * 2.0f), expf(-s.transmittance.z * distance * 2.0f));
        float ax, ay; alpha_from_roughness(s.roughness, s.anisotropic, ax, ay);
        const V3 m = ggx_sample(wol, r1, r3, ax, ay);
        const float rcp_eta = 1 / eta, cos_wom = clampf(dot(wol, m), -1.0f, 1.0f);
        float ct, jac;
        const float F = fresnel_dielectric(cos_wom, eta, ct);
        V3 wil, ret = v3(0);
        if (r2 < F) {
            wil = reflect(wol * -1.0f, m);
            if (wil.z * wol.z <= 0) return v3(0);
            reflection_eval(s.color, wol, wil, m, ax, ay, F, ret);
            pdf = F; jac = reflection_jacobian(cos_wom);
        } else {
            wil = refract_dir(wol, m, cos_wom, ct, eta);
            if (wil.z * wol.z > 0) return v3(0);
            refraction_eval(rcp_eta, s.color, false, wol, wil, m, ax, ay, 1 - F, ret);
            pdf = 1 - F; jac = refraction_jacobian(wol, wil, m, rcp_eta);
        }
        pdf *= jac * ggx_pdf(wol, m, ax, ay);
        if (pdf > 1.0e-6f) wiw = to_world(wil, iN, T, B);
        return ret * beer;
    }
    const float r3 = (r0 - s.transmission) / (1 - s.transmission);
    float w[4]; lobe_weights(s, w);
    const float cdf_x = w[0], cdf_y = w[0] + w[1], cdf_z = w[0] + w[1] + w[2];
    float probability, component_pdf = 0;
    V3 contrib = v3(0), value = v3(0);
    if (r3 < cdf_y) {
        const float rr = r3 / cdf_y;
        wiw = cos_weighted(rr, r1, iN, T, B);
        const V3 m = normalize(wiw + wow);
        if (r3 < cdf_x) { component_pdf = diffuse_eval(s, iN, wow, wiw, m, value); probability = w[0] * component_pdf; w[0] = 0; }
        else            { component_pdf = sheen_eval(s, wiw, m, value);           probability = w[1] * component_pdf; w[1] = 0; }
    } else {
        const V3 wol = to_local(wow, iN, T, B);
        V3 wil = v3(0);
        if (r3 < cdf_z) {
            const float rr = (r3 - cdf_y) / (cdf_z - cdf_y);
            float ax, ay; alpha_from_roughness(s.roughness, s.anisotropic, ax, ay);
            mf_sample(MDF_GGX, s, rr, r1, ax, ay, wol, wil, component_pdf, value);
            probability = w[2] * component_pdf; w[2] = 0;
        } else {
            const float rr = (r3 - cdf_z) / (1 - cdf_z);
            const float a = coat_roughness(s);
            mf_sample(MDF_GTR1, s, rr, r1, a, a, wol, wil, component_pdf, value);
            probability = w[3] * component_pdf; w[3] = 0;
        }
        value *= 1.0f / fabsf(4.0f * wol.z * wil.z);
        wiw = to_world(wil, iN, T, B);
    }
    if (w[0] + w[1] > 0) {
        const V3 m = normalize(wiw + wow);
        if (w[0] > 0) { contrib = v3(0); probability += w[0] * diffuse_eval(s, iN, wow, wiw, m, contrib); value += contrib; }
        if (w[1] > 0) { contrib = v3(0); probability += w[1] * sheen_eval(s, wiw, m, contrib); value += contrib; }
    }
    if (w[2] + w[3] > 0) {
        const V3 wol = to_local(wow, iN, T, B), wil = to_local(wiw, iN, T, B);
        const V3 m = normalize(wol + wil);
        if (w[2] > 0) {
            float ax, ay; alpha_from_roughness(s.roughness, s.anisotropic, ax, ay);
            contrib = v3(0); probability += w[2] * mf_eval(MDF_GGX, s, ax, ay, wol, wil, m, contrib); value += contrib;
        }
        if (w[3] > 0) {
            const float a = coat_roughness(s);
            contrib = v3(0); probability += w[3] * mf_eval(MDF_GTR1, s, a, a, wol, wil, m, contrib); value += contrib;
        }
    }
    pdf = probability > 1.0e-6f ? probability : 0;
    return value;
}

} // namespace lo
