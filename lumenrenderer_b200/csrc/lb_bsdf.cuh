// lb_bsdf.cuh — Disney "principled" BSDF for the shade / RIS / reuse kernels.
//
// Behaviour follows the reference's device functions (paths under /root/reference/Lumen_Engine/LumenPT/src/CUDAKernels/):
//   SampleBSDF    disney.cuh:173-304        EvaluateBSDF  disney.cuh:320-405     lobes disney.cuh:33-150
//   GGX / GTR1    ggxmdf.cuh:43-228         dielectric    frosted.cuh:28-120     frames/sampling bsdf_math.cuh:57-176
//   8-bit material parameter decode  Shaders/CppCommon/MaterialStructs.h:84-260
// Evaluation order is kept expression by expression so that results differ from the host-compiled reference
// headers only through the device libm (sinf/cosf/logf/expf/powf), see tests/test_gpu_bsdf.py for the tolerance.
#pragma once
#include "lb_device.cuh"

namespace lb {

constexpr float kPi = 3.14159265358979323846264f;
constexpr float kInvPi = 0.31830988618379067153777f;
constexpr float kTwoPi = 6.28318530717958647692528f;
constexpr float kBsdfEps = 0.0001f;      // EPSILON macro, bsdf_math.cuh:39-41 (shadows RenderingUtility.h's FLT_EPSILON)

LB_HD float unpack8(uint32_t word, int shift) { return ((float)((word >> shift) & 255u)) * (1.0f / 255.0f); }
LB_HD void pack8(uint32_t& word, float v, int shift) {
    const uint32_t c = (uint32_t)(v * 255.f);
    word &= ~(255u << shift);
    word |= c << shift;
}

// Material decoded once into registers.
struct Shading {
    float3 color, transmittance, tint;
    float ior, luminance;
    float metallic, subsurface, specular, roughness, spectint, anisotropic, sheen, sheentint, clearcoat, clearcoatgloss, transmission;
    LB_D explicit Shading(const Material& m) {
        color = f3(m.color); transmittance = f3(m.transmittance); ior = m.transmittance.w; tint = f3(m.tint); luminance = m.tint.w;
        metallic = unpack8(m.params.x, 0); subsurface = unpack8(m.params.x, 8); specular = unpack8(m.params.x, 16); roughness = unpack8(m.params.x, 24);
        spectint = unpack8(m.params.y, 0); anisotropic = unpack8(m.params.y, 8); sheen = unpack8(m.params.y, 16); sheentint = unpack8(m.params.y, 24);
        clearcoat = unpack8(m.params.z, 0); clearcoatgloss = unpack8(m.params.z, 8); transmission = unpack8(m.params.z, 16);
    }
};

// Two kinds of arithmetic live in this header (DESIGN.md "Arithmetic contract"). EVALUATION (value and pdf of a given pair of directions)
// is well conditioned — a relative error of 1e-6 in, 1e-6 out — and is written with plain operators: it takes the unit's division /
// square root / transcendentals (fast in lb_wavefront.cu and lb_restir.cu). SAMPLING A DIRECTION is not: the direction a bounce ray leaves
// in decides what it hits, and one more ulp of error at a near-mirror surface (alpha down to 1e-4) is a different radiance at the next
// vertex. Everything on the way from the random numbers to the sampled direction (lobe weights and thresholds, alphas, frames, the lobe
// samplers, Fresnel split, refraction) therefore spells IEEE division / square root (xdiv, xsqrt), the accurate pow (xpow) and the portable sin / cos of lb_device.cuh (det_sincos).
// ---------------------------------------------------------------- microfacet distributions
LB_D void mf_alpha(float roughness, float anisotropy, float& ax, float& ay) {
    const float r2 = roughness * roughness;
    const float aspect = xsqrt(1.0f + anisotropy * (anisotropy < 0 ? 0.9f : -0.9f));
    ax = fmaxf(0.001f, xdiv(r2, aspect));
    ay = fmaxf(0.001f, r2 * aspect);
}
// ISO: the caller has established ax == ay (isotropic roughness, every material without the `anisotropic` parameter). The anisotropic
// expressions are then not even generated: left to the compiler they are if-converted and executed predicated-off — 3 % of the instructions
// and two MUFU divisions per evaluation for a material that never uses them (profiles/r02_a_ab.md).
template <bool ISO = false>
LB_D float ggx_D(const float3& m, float ax, float ay) {
    if (m.z == 0) return sq(ax) * kInvPi;
    const float c2 = sq(m.z);
    const float s = sqrtf(fmaxf(0.0f, 1 - c2));
    const float t2 = (1.0f - c2) / c2;
    float stretched;
    if (ISO || ax == ay || s == 0.0f) stretched = 1.0f / sq(ax);
    else stretched = sq(m.x / (s * ax)) + sq(m.y / (s * ay));
    return 1.0f / (kPi * ax * ay * sq(c2) * sq(1.0f + t2 * stretched));
}
template <bool ISO = false>
LB_D float ggx_Lambda(const float3& v, float ax, float ay) {
    if (v.z == 0) return 0;
    const float c2 = v.z * v.z;
    const float s = sqrtf(fmaxf(0.0f, 1 - c2));
    float projected;
    if (ISO || ax == ay || s == 0.0f) projected = ax;
    else projected = sqrtf(sq((v.x * ax) / s) + sq((v.y * ay) / s));
    const float t2 = sq(s) / c2;
    const float a2_rcp = sq(projected) * t2;
    return (-1.0f + sqrtf(1.0f + a2_rcp)) * 0.5f;
}
LB_D float ggx_G(const float3& wi, const float3& wo, float ax, float ay) { return 1.0f / (1.0f + ggx_Lambda(wo, ax, ay) + ggx_Lambda(wi, ax, ay)); }
LB_D float ggx_G1(const float3& v, float ax, float ay) { return 1.0f / (1.0f + ggx_Lambda(v, ax, ay)); }
LB_D float ggx_pdf(const float3& v, const float3& m, float ax, float ay) {
    if (v.z == 0.0f) return 0;
    return ggx_G1(v, ax, ay) * fabsf(dot(v, m)) * ggx_D(m, ax, ay) / fabsf(v.z);
}
// visible-normal sampling, device branch of ggxmdf.cuh:78-107
LB_D float3 ggx_sample(const float3& v, float r0, float r1, float ax, float ay) {
    const float sgn = v.z < 0.0f ? -1.0f : 1.0f;
    const float3 st = xnormalize(f3(sgn * v.x * ax, sgn * v.y * ay, sgn * v.z));
    const float3 t1 = v.z < 0.9999f ? xnormalize(cross(st, f3(0, 0, 1))) : f3(1, 0, 0);
    const float3 t2 = cross(t1, st);
    const float a = xdiv(1.0f, 1.0f + st.z);
    const float r = xsqrt(r0);
    const float phi = r1 < a ? (xdiv(r1, a) * kPi) : (kPi + xdiv(r1 - a, 1.0f - a) * kPi);
    float p1, p2; det_sincos(phi, p2, p1);
    p1 *= r;
    p2 *= r * (r1 < a ? 1.0f : st.z);
    const float3 h = p1 * t1 + p2 * t2 + xsqrt(fmaxf(0.0f, 1.0f - p1 * p1 - p2 * p2)) * st;
    return xnormalize(f3(h.x * ax, h.y * ay, fmaxf(0.0f, h.z)));
}
LB_D float gtr1_clamp(float a) { return clampf(a, 0.001f, 0.999f); }
LB_D float gtr1_D(const float3& m, float ax) {
    const float a2 = sq(gtr1_clamp(ax));
    const float a = (a2 - 1.0f) / (kPi * logf(a2));
    const float b = (1 / (1 + (a2 - 1) * sq(m.z)));
    return a * b;
}
LB_D float gtr1_Lambda(const float3& v, float ax) {
    if (v.z == 0) return 0;
    const float c2 = sq(v.z);
    const float s = sqrtf(fmaxf(0.0f, 1.0f - c2));
    if (s == 0) return 0;
    const float cot2 = c2 / sq(s);
    const float cot = sqrtf(cot2);
    const float a2 = sq(gtr1_clamp(ax));
    const float a = sqrtf(cot2 + a2);
    const float b = sqrtf(cot2 + 1.0f);
    const float c = logf(cot + b);
    const float d = logf(cot + a);
    return (a - b + cot * (c - d)) / (cot * logf(a2));
}
LB_D float gtr1_G(const float3& wi, const float3& wo, float ax) { return 1.0f / (1.0f + gtr1_Lambda(wo, ax) + gtr1_Lambda(wi, ax)); }
LB_D float gtr1_pdf(const float3& m, float ax) { return gtr1_D(m, ax) * fabsf(m.z); }
LB_D float3 gtr1_sample(float r0, float r1, float ax) {
    const float a2 = sq(gtr1_clamp(ax));
    const float c2 = xdiv(1.0f - xpow(a2, 1.0f - r0), 1.0f - a2);
    const float s = xsqrt(fmaxf(0.0f, 1.0f - c2));
    const float phi = kTwoPi * r1;
    float cp, sp; det_sincos(phi, sp, cp);
    return f3(cp * s, sp * s, xsqrt(c2));
}

// ---------------------------------------------------------------- lobes
LB_D float schlick_w(float u) { const float m = clampf(1.0f - u, 0.f, 1.f), m2 = sq(m), m4 = sq(m2); return m4 * m; }
LB_D float3 fresnel_spec(const Shading& s, const float3& o, const float3& h) {
    float3 v = (1.0f - s.spectint) + s.spectint * s.tint;
    v *= s.specular * 0.08f;
    v = (1.0f - s.metallic) * v + s.metallic * s.color;
    const float f = schlick_w(fabsf(dot(o, h)));
    return (1.0f - f) * v + f;
}
LB_D float3 fresnel_coat(const Shading& s, const float3& o, const float3& h) {
    return f3(mixf(0.04f, 1.0f, schlick_w(fabsf(dot(o, h)))) * 0.25f * s.clearcoat);
}
LB_D float coat_alpha(const Shading& s) { return mixf(0.1f, 0.001f, s.clearcoatgloss); }

template <bool GGX> LB_D float mdf_D(const float3& m, float ax, float ay) { return GGX ? ggx_D(m, ax, ay) : gtr1_D(m, ax); }
template <bool GGX> LB_D float mdf_G(const float3& wi, const float3& wo, float ax, float ay) { return GGX ? ggx_G(wi, wo, ax, ay) : gtr1_G(wi, wo, ax); }
template <bool GGX> LB_D float mdf_pdf(const float3& v, const float3& m, float ax, float ay) { return GGX ? ggx_pdf(v, m, ax, ay) : gtr1_pdf(m, ax); }

// sample one microfacet lobe (disney.cuh:78-99). The component pdf stays 0 where the reference leaves it unset.
template <bool GGX>
LB_D void lobe_sample(const Shading& s, float r0, float r1, float ax, float ay, const float3& wol, float3& wil, float& pdf, float3& value) {
    if (wol.z == 0) { value = f3(0.f); pdf = 0; return; }
    const float3 m = GGX ? ggx_sample(wol, r0, r1, ax, ay) : gtr1_sample(r0, r1, ax);
    wil = reflect(wol * -1.0f, m);
    pdf = 0;
    if (wil.z == 0) return;
    const float cos_oh = dot(wol, m);
    pdf = mdf_pdf<GGX>(wol, m, ax, ay) / fabsf(4.0f * cos_oh);
    if (pdf < 1.0e-6f) return;
    const float D = mdf_D<GGX>(m, ax, ay), G = mdf_G<GGX>(wil, wol, ax, ay);
    value = GGX ? fresnel_spec(s, wol, m) : fresnel_coat(s, wol, m);
    value *= D * G;
}
// evaluate one microfacet lobe (disney.cuh:101-113)
template <bool GGX>
LB_D float lobe_eval(const Shading& s, float ax, float ay, const float3& wol, const float3& wil, const float3& m, float3& bsdf) {
    if (wol.z == 0 || wil.z == 0) return 0;
    const float cos_oh = dot(wol, m);
    if (cos_oh == 0) return 0;
    const float D = mdf_D<GGX>(m, ax, ay), G = mdf_G<GGX>(wil, wol, ax, ay);
    bsdf = GGX ? fresnel_spec(s, wol, m) : fresnel_coat(s, wol, m);
    bsdf *= D * G / fabsf(4.0f * wol.z * wil.z);
    return mdf_pdf<GGX>(wol, m, ax, ay) / fabsf(4.0f * cos_oh);
}
LB_D float diffuse_lobe(const Shading& s, const float3& n, const float3& wo, const float3& wi, const float3& m, float3& value) {
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
LB_D float sheen_lobe(const Shading& s, const float3& wi, const float3& m, float3& value) {
    const float fh = schlick_w(dot(wi, m));
    value = (1.0f - s.sheentint) + s.sheentint * s.tint;
    value *= fh * s.sheen * (1.0f - s.metallic);
    return 1.0f / (2 * kPi);
}

// ---------------------------------------------------------------- rough dielectric (transmission branch)
LB_D float fresnel_dielectric(float cos_i, float eta, float& cos_t) {
    const float s2 = (1 - sq(cos_i)) * sq(eta);
    if (s2 > 1) { cos_t = 0; return 1; }
    cos_t = fminf(xsqrt(fmaxf(1 - s2, 0.0f)), 1.0f);            // cos_t is a component of the refracted direction, F splits reflect / refract
    const float ci = fabsf(cos_i);
    if (ci == 0 && cos_t == 0) return 1;
    const float k0 = eta * cos_t, k1 = eta * ci;
    return 0.5f * (sq(xdiv(ci - k0, ci + k0)) + sq(xdiv(cos_t - k1, cos_t + k1)));
}
LB_D float3 refracted(const float3& wo, const float3& m, float cos_wom, float cos_t, float rcp_eta) {
    const float3 wi = cos_wom > 0 ? (rcp_eta * cos_wom - cos_t) * m - rcp_eta * wo
                                  : (rcp_eta * cos_wom + cos_t) * m - rcp_eta * wo;
    return wi * ((3 - dot(wi, wi)) * 0.5f);
}
LB_D float pick_reflection(float rw, float tw, float F) {
    const float r = F * rw, t = (1 - F) * tw, sum = r + t;
    return sum != 0 ? r / sum : 1;
}
LB_D float3 half_reflect(const float3& wo, const float3& wi) { const float3 h = normalize(wi + wo); return h.z < 0 ? (h * -1.f) : h; }
LB_D float3 half_refract(const float3& wo, const float3& wi, float eta) { const float3 h = normalize(wo + eta * wi); return h.z < 0 ? (h * -1.f) : h; }
LB_D void eval_reflection(const float3& color, const float3& wo, const float3& wi, const float3& m, float ax, float ay, float F, float3& value) {
    const float denom = fabsf(4 * wo.z * wi.z);
    if (denom == 0) { value = f3(0.f); return; }
    const float D = ggx_D(m, ax, ay), G = ggx_G(wi, wo, ax, ay);
    value = color * (F * D * G / denom);
}
LB_D void eval_refraction(float eta, const float3& color, const float3& wo, const float3& wi, const float3& m, float ax, float ay, float T, float3& value) {
    if (wo.z == 0 || wi.z == 0) { value = f3(0.f); return; }
    const float cos_ih = dot(m, wi), cos_oh = dot(m, wo);
    const float dots = (cos_ih * cos_oh) / (wi.z * wo.z);
    const float sd = cos_oh + eta * cos_ih;
    if (fabsf(sd) < 1.0e-6f) { value = f3(0.f); return; }
    const float D = ggx_D(m, ax, ay), G = ggx_G(wi, wo, ax, ay);
    float mult = fabsf(dots) * T * D * G / sq(sd);
    mult *= sq(eta);                 // radiance transport (adjoint == false)
    value = color * mult;
}
LB_D float jacobian_reflection(float cos_oh) { return cos_oh == 0 ? 0 : 1 / (4 * fabsf(cos_oh)); }
LB_D float jacobian_refraction(const float3& wo, const float3& wi, const float3& m, float eta) {
    const float cos_ih = dot(m, wi), cos_oh = dot(m, wo);
    const float sd = cos_oh + eta * cos_ih;
    if (fabsf(sd) < 1.0e-6f) return 0;
    return fabsf(cos_ih) * sq(eta / sd);
}

LB_D float3 to_local(const float3& v, const float3& n, const float3& t, const float3& b) { return f3(dot(v, t), dot(v, b), dot(v, n)); }
LB_D float3 to_world(const float3& v, const float3& n, const float3& t, const float3& b) { return v.x * t + v.y * b + v.z * n; }
LB_D float3 cosine_hemisphere(float r0, float r1, const float3& n, const float3& t, const float3& b) {
    const float term1 = kTwoPi * r0, term2 = xsqrt(1 - r1);
    float s, c; det_sincos(term1, s, c);
    return (c * term2 * t) + (s * term2) * b + xsqrt(r1) * n;
}
LB_D void lobe_weights(const Shading& s, float w[4]) {
    w[0] = mixf(s.luminance, 0.f, s.metallic); w[1] = mixf(s.sheen, 0.f, s.metallic);
    w[2] = mixf(s.specular, 1.f, s.metallic);  w[3] = s.clearcoat * 0.25f;
    const float inv = xdiv(1.0f, w[0] + w[1] + w[2] + w[3]);     // the weights are the lobe-selection thresholds of bsdf_sample
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] *= inv;
}

// ---------------------------------------------------------------- EvaluateBSDF (disney.cuh:320-405)
// The reference evaluates the whole BSDF from scratch per (surface, light direction) pair. RIS evaluates 32 light
// directions and spatial reuse up to 5 against ONE surface, so everything that depends only on the surface and the
// outgoing direction is computed once here (tangent frame, local outgoing direction, lobe weights, microfacet alphas,
// the outgoing-side Lambda / G1 / Schlick terms, the GTR1 normalisation with its logf). Every hoisted value is the
// same expression the per-call code evaluated, so results are bit-identical to the un-hoisted form.
struct BsdfCtx {
    Shading s;
    float3 N, T, B, wow, wol;
    float w[4];
    float ax, ay;                       // GGX alphas (microfacet_alpha_from_roughness)
    float3 spec_v, sheen_v;             // view-independent parts of the specular / sheen Fresnel colour
    float lam_wo_ggx, g1_wo_ggx;        // Lambda(wol), G1(wol) of the specular lobe
    float coat_a2, coat_norm, coat_log_a2, lam_wo_gtr;   // GTR1 clear-coat: alpha^2, (a2-1)/(pi log a2), log a2, Lambda(wol)
    float cos_on, fv;                   // diffuse lobe: N.wo and its Schlick weight

    LB_D BsdfCtx(const Material& mat, const float3& iN, const float3& iT, const float3& wow_) : s(mat) {
        N = iN; wow = wow_;
        B = normalize(cross(iN, iT)); T = normalize(cross(iN, B));
        wol = to_local(wow, iN, T, B);
        mf_alpha(s.roughness, s.anisotropic, ax, ay);
        lobe_weights(s, w);
        spec_v = (1.0f - s.spectint) + s.spectint * s.tint;
        spec_v *= s.specular * 0.08f;
        spec_v = (1.0f - s.metallic) * spec_v + s.metallic * s.color;
        sheen_v = (1.0f - s.sheentint) + s.sheentint * s.tint;
        lam_wo_ggx = ggx_Lambda(wol, ax, ay);
        g1_wo_ggx = 1.0f / (1.0f + lam_wo_ggx);
        const float ca = coat_alpha(s);
        coat_a2 = sq(gtr1_clamp(ca));
        coat_log_a2 = logf(coat_a2);
        coat_norm = (coat_a2 - 1.0f) / (kPi * coat_log_a2);
        lam_wo_gtr = (w[3] > 0) ? gtr1_Lambda(wol, ca) : 0.f;
        cos_on = dot(iN, wow);
        fv = schlick_w(cos_on);
    }

    LB_D float3 fresnel_spec_at(const float3& h) const {
        const float f = schlick_w(fabsf(dot(wol, h)));
        return (1.0f - f) * spec_v + f;
    }
    template <bool ISO = false>
    LB_D float ggx_pdf_wo(const float3& m) const {               // ggx_pdf(wol, m)
        if (wol.z == 0.0f) return 0;
        return g1_wo_ggx * fabsf(dot(wol, m)) * ggx_D<ISO>(m, ax, ay) / fabsf(wol.z);
    }
    LB_D float gtr1_D_m(const float3& m) const { return coat_norm * (1 / (1 + (coat_a2 - 1) * sq(m.z))); }
    LB_D float gtr1_Lambda_wi(const float3& v) const {
        if (v.z == 0) return 0;
        const float c2 = sq(v.z);
        const float sn = sqrtf(fmaxf(0.0f, 1.0f - c2));
        if (sn == 0) return 0;
        const float cot2 = c2 / sq(sn);
        const float cot = sqrtf(cot2);
        const float a = sqrtf(cot2 + coat_a2);
        const float b = sqrtf(cot2 + 1.0f);
        const float c = logf(cot + b);
        const float d = logf(cot + a);
        return (a - b + cot * (c - d)) / (cot * coat_log_a2);
    }
    // lobe_eval<true> (GGX specular) and lobe_eval<false> (GTR1 clear coat) against the hoisted outgoing side
    template <bool ISO = false>
    LB_D float spec_eval(const float3& wil, const float3& m, float3& bsdf) const {
        if (wol.z == 0 || wil.z == 0) return 0;
        const float cos_oh = dot(wol, m);
        if (cos_oh == 0) return 0;
        const float D = ggx_D<ISO>(m, ax, ay), G = 1.0f / (1.0f + lam_wo_ggx + ggx_Lambda<ISO>(wil, ax, ay));
        bsdf = fresnel_spec_at(m);
        bsdf *= D * G / fabsf(4.0f * wol.z * wil.z);
        return ggx_pdf_wo<ISO>(m) / fabsf(4.0f * cos_oh);
    }
    LB_D float coat_eval(const float3& wil, const float3& m, float3& bsdf) const {
        if (wol.z == 0 || wil.z == 0) return 0;
        const float cos_oh = dot(wol, m);
        if (cos_oh == 0) return 0;
        const float D = gtr1_D_m(m), G = 1.0f / (1.0f + lam_wo_gtr + gtr1_Lambda_wi(wil));
        bsdf = f3(mixf(0.04f, 1.0f, schlick_w(fabsf(dot(wol, m)))) * 0.25f * s.clearcoat);
        bsdf *= D * G / fabsf(4.0f * wol.z * wil.z);
        return (D * fabsf(m.z)) / fabsf(4.0f * cos_oh);
    }
    LB_D float diffuse_eval(const float3& wi, const float3& m, float3& value) const {
        const float cos_in = dot(N, wi), cos_ih = dot(wi, m);
        const float fl = schlick_w(cos_in);
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
    LB_D float sheen_eval(const float3& wi, const float3& m, float3& value) const {
        const float fh = schlick_w(dot(wi, m));
        value = sheen_v;
        value *= fh * s.sheen * (1.0f - s.metallic);
        return 1.0f / (2 * kPi);
    }

    // Materials without transmission, sheen, clear coat, anisotropy and subsurface scattering — most of a typical scene — keep only the diffuse
    // and the isotropic GGX lobe. eval_simple() is eval() with the branches such a material never takes removed: the same expressions in the
    // same order for what is left (bit-identical to eval() in the exact-arithmetic unit: the debug tap of lb_debug.cu poisons its output on a
    // mismatch, so the golden-vector tests of tests/test_gpu_bsdf.py check it for every simple material they hold). The callers
    // that evaluate many samples per pixel (RIS: 32) pick it per WARP, so the warp runs a loop body without the predicated-off lobes.
    LB_D bool is_simple() const { return s.transmission == 0.f && w[1] == 0.f && w[3] == 0.f && ax == ay && s.subsurface == 0.f && s.roughness > 0.001f; }
    LB_D float3 eval_simple(const float3& wiw, float& pdf) const {
        pdf = 0; float3 value = f3(0.f);
        if (w[0] > 0) {
            const float3 m = normalize(wiw + wow);
            const float cos_in = dot(N, wiw), cos_ih = dot(wiw, m);
            const float fl = schlick_w(cos_in);
            const float fd90 = 0.5f + 2.0f * sq(cos_ih) * s.roughness;
            const float fd = mixf(1.f, fd90, fl) * mixf(1.f, fd90, fv);
            value = s.color * fd * kInvPi * (1.0f - s.metallic);
            pdf += w[0] * (fabsf(cos_in) * kInvPi);
        }
        if (w[2] > 0) {
            const float3 wil = to_local(wiw, N, T, B);
            const float3 m = normalize(wol + wil);
            if (wol.z != 0 && wil.z != 0) {
                const float cos_oh = dot(wol, m);
                if (cos_oh != 0) {
                    const float D = ggx_D(m, ax, ax), G = 1.0f / (1.0f + lam_wo_ggx + ggx_Lambda(wil, ax, ax));
                    float3 c = fresnel_spec_at(m);
                    c *= D * G / fabsf(4.0f * wol.z * wil.z);
                    const float p = (g1_wo_ggx * fabsf(cos_oh) * D / fabsf(wol.z)) / fabsf(4.0f * cos_oh);
                    if (p > 0) { pdf += w[2] * p; value += c; }
                }
            }
        }
        return value;
    }

    LB_D bool is_isotropic() const { return ax == ay; }
    // ISO: the caller has established is_isotropic() (see ggx_D)
    template <bool ISO = false>
    LB_D float3 eval(const float3& wiw, float& pdf) const {
        float3 trans_bsdf = f3(0.f); float trans_pdf = 0.f;
        if (s.transmission > 0.f) {
            const float3 wil = to_local(wiw, N, T, B);
            const float eta = wol.z > 0 ? s.ior : (1.0f / s.ior);
            if (eta == 1) { pdf = 0; return f3(0.f); }
            float jac;
            float3 m;
            if (wil.z * wol.z >= 0) {
                m = half_reflect(wol, wil);
                const float cos_wom = dot(wol, m); float ct;
                const float F = fresnel_dielectric(cos_wom, 1 / eta, ct);
                eval_reflection(s.color, wol, wil, m, ax, ay, F, trans_bsdf);
                trans_pdf = pick_reflection(1, 1, F); jac = jacobian_reflection(cos_wom);
            } else {
                m = half_refract(wol, wil, eta);
                const float cos_wom = dot(wol, m); float ct;
                const float F = fresnel_dielectric(cos_wom, 1 / eta, ct);
                eval_refraction(eta, s.color, wol, wil, m, ax, ay, 1 - F, trans_bsdf);
                trans_pdf = 1 - pick_reflection(1, 1, F); jac = jacobian_refraction(wol, wil, m, eta);
            }
            trans_pdf *= jac * ggx_pdf_wo<ISO>(m);
        }
        if (s.roughness <= 0.001f) { pdf = trans_pdf; return trans_bsdf; }
        pdf = 0; float3 value = f3(0.f);
        if (w[0] + w[1] > 0) {
            const float3 m = normalize(wiw + wow);
            if (w[0] > 0) pdf += w[0] * diffuse_eval(wiw, m, value);
            if (w[1] > 0) pdf += w[1] * sheen_eval(wiw, m, value);   // replaces the diffuse value, as the reference does (disney.cuh:377)
        }
        if (w[2] + w[3] > 0) {
            const float3 wil = to_local(wiw, N, T, B);
            const float3 m = normalize(wol + wil);
            if (w[2] > 0) {
                float3 c = f3(0.f); const float p = spec_eval<ISO>(wil, m, c);
                if (p > 0) { pdf += w[2] * p; value += c; }
            }
            if (w[3] > 0) {
                float3 c = f3(0.f); const float p = coat_eval(wil, m, c);
                if (p > 0) { pdf += w[3] * p; value += c; }
            }
        }
        pdf = (pdf * (1.f - s.transmission));
        pdf += (trans_pdf * s.transmission);
        return (trans_bsdf * s.transmission) + (value * (1.f - s.transmission));
    }
};

LB_D float3 bsdf_eval(const Material& mat, const float3& iN, const float3& iT, const float3& wow, const float3& wiw, float& pdf) {
    const BsdfCtx c(mat, iN, iT, wow);
#if !defined(LB_ISO_PER_LANE) || LB_ISO_PER_LANE
    if (c.is_isotropic()) return c.eval<true>(wiw, pdf);
#endif
    return c.eval<false>(wiw, pdf);
}

// ---------------------------------------------------------------- SampleBSDF (disney.cuh:173-304)
LB_D float3 bsdf_sample(const Material& mat, float3 iN, const float3& N, const float3& iT, const float3& wow, float distance,
                        float r0, float r1, float r2, float3& wiw, float& pdf, bool& specular) {
    const Shading s(mat);
    const float flip = (dot(wow, N) < 0) ? -1.f : 1.f;
    iN *= flip;
    const float3 B = xnormalize(cross(iN, iT)), T = xnormalize(cross(iN, B));
    if (r0 < s.transmission) {
        specular = true;
        const float r3 = xdiv(r0, s.transmission);
        const float3 wol = to_local(wow, iN, T, B);
        const float eta = flip < 0 ? xdiv(1.f, s.ior) : s.ior;
        if (eta == 1) return f3(0.f);
        const float3 beer = f3(expf(-s.transmittance.x * distance * 2.0f), expf(-s.transmittance.y * distance * 2.0f), expf(-s.transmittance.z * distance * 2.0f));
        float ax, ay; mf_alpha(s.roughness, s.anisotropic, ax, ay);
        const float3 m = ggx_sample(wol, r1, r3, ax, ay);
        const float rcp_eta = xdiv(1.f, eta), cos_wom = clampf(dot(wol, m), -1.0f, 1.0f);
        float ct, jac;
        const float F = fresnel_dielectric(cos_wom, eta, ct);
        float3 wil, ret = f3(0.f);
        if (r2 < F) {
            wil = reflect(wol * -1.0f, m);
            if (wil.z * wol.z <= 0) return f3(0.f);
            eval_reflection(s.color, wol, wil, m, ax, ay, F, ret);
            pdf = F; jac = jacobian_reflection(cos_wom);
        } else {
            wil = refracted(wol, m, cos_wom, ct, eta);
            if (wil.z * wol.z > 0) return f3(0.f);
            eval_refraction(rcp_eta, s.color, wol, wil, m, ax, ay, 1 - F, ret);
            pdf = 1 - F; jac = jacobian_refraction(wol, wil, m, rcp_eta);
        }
        pdf *= jac * ggx_pdf(wol, m, ax, ay);
        if (pdf > 1.0e-6f) wiw = to_world(wil, iN, T, B);
        return ret * beer;
    }
    const float r3 = xdiv(r0 - s.transmission, 1 - s.transmission);
    float w[4]; lobe_weights(s, w);
    const float cdf_x = w[0], cdf_y = w[0] + w[1], cdf_z = w[0] + w[1] + w[2];
    float probability, component_pdf = 0;
    float3 contrib = f3(0.f), value = f3(0.f);
    if (r3 < cdf_y) {
        const float rr = xdiv(r3, cdf_y);
        wiw = cosine_hemisphere(rr, r1, iN, T, B);
        const float3 m = normalize(wiw + wow);
        if (r3 < cdf_x) { component_pdf = diffuse_lobe(s, iN, wow, wiw, m, value); probability = w[0] * component_pdf; w[0] = 0; }
        else            { component_pdf = sheen_lobe(s, wiw, m, value);           probability = w[1] * component_pdf; w[1] = 0; }
    } else {
        const float3 wol = to_local(wow, iN, T, B);
        float3 wil = f3(0.f);
        if (r3 < cdf_z) {
            const float rr = xdiv(r3 - cdf_y, cdf_z - cdf_y);
            float ax, ay; mf_alpha(s.roughness, s.anisotropic, ax, ay);
            lobe_sample<true>(s, rr, r1, ax, ay, wol, wil, component_pdf, value);
            probability = w[2] * component_pdf; w[2] = 0;
        } else {
            const float rr = xdiv(r3 - cdf_z, 1 - cdf_z);
            const float a = coat_alpha(s);
            lobe_sample<false>(s, rr, r1, a, a, wol, wil, component_pdf, value);
            probability = w[3] * component_pdf; w[3] = 0;
        }
        value *= 1.0f / fabsf(4.0f * wol.z * wil.z);
        wiw = to_world(wil, iN, T, B);
    }
    if (w[0] + w[1] > 0) {
        const float3 m = normalize(wiw + wow);
        if (w[0] > 0) { contrib = f3(0.f); probability += w[0] * diffuse_lobe(s, iN, wow, wiw, m, contrib); value += contrib; }
        if (w[1] > 0) { contrib = f3(0.f); probability += w[1] * sheen_lobe(s, wiw, m, contrib); value += contrib; }
    }
    if (w[2] + w[3] > 0) {
        const float3 wol = to_local(wow, iN, T, B), wil = to_local(wiw, iN, T, B);
        const float3 m = normalize(wol + wil);
        if (w[2] > 0) {
            float ax, ay; mf_alpha(s.roughness, s.anisotropic, ax, ay);
            contrib = f3(0.f); probability += w[2] * lobe_eval<true>(s, ax, ay, wol, wil, m, contrib); value += contrib;
        }
        if (w[3] > 0) {
            const float a = coat_alpha(s);
            contrib = f3(0.f); probability += w[3] * lobe_eval<false>(s, a, a, wol, wil, m, contrib); value += contrib;
        }
    }
    pdf = probability > 1.0e-6f ? probability : 0;
    return value;
}

} // namespace lb
