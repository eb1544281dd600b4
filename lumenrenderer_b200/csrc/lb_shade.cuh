// lb_shade.cuh — per-ray shading building blocks: texture fetch, surface extraction, next-event estimation,
// BSDF bounce, light-CDF lookup. Device-inline so that one fused kernel per wave keeps the surface record in
// registers instead of writing/reading the reference's 176-B SurfaceData three times per wave.
// Reference behaviour (paths under /root/reference/Lumen_Engine/LumenPT/src/):
//   surface extraction   CUDAKernels/WaveFrontKernels/GPUExtractSurfaceData.cu:8-228
//   NEE                  CUDAKernels/WaveFrontKernels/GPUShadeDirect.cu:42-153
//   bounce               CUDAKernels/WaveFrontKernels/GPUShadeIndirect.cu:7-146
//   CDF                  Shaders/CppCommon/ReSTIRData.h:232-302
//   textures             Framework/PTTexture.cpp:35-74 (wrap, bilinear, optional sRGB) — evaluated here in exact fp32
//                        arithmetic on raw RGBA8 texels so results do not depend on the 9-bit texture-unit filter
#pragma once
#include "lb_bsdf.cuh"
#ifndef LB_ISO_PER_LANE
#define LB_ISO_PER_LANE 1
#endif

namespace lb {

LB_D float4 texel_at(const SceneView& sc, const DevTexture& t, int x, int y) {
    const uchar4 p = __ldg(&sc.texels[t.offset + (uint32_t)y * t.w + (uint32_t)x]);
    if (t.srgb) return make_float4(__ldg(&sc.srgb_lut[p.x]), __ldg(&sc.srgb_lut[p.y]), __ldg(&sc.srgb_lut[p.z]), (float)p.w * (1.0f / 255.0f));
    return make_float4((float)p.x * (1.0f / 255.0f), (float)p.y * (1.0f / 255.0f), (float)p.z * (1.0f / 255.0f), (float)p.w * (1.0f / 255.0f));
}
LB_D float4 mix4(const float4& p, const float4& q, float s) { return make_float4(mixf(p.x, q.x, s), mixf(p.y, q.y, s), mixf(p.z, q.z, s), mixf(p.w, q.w, s)); }
LB_D float4 tex2d(const SceneView& sc, int handle, float u, float v) {
    const DevTexture t = sc.textures[handle];
    if (t.w == 1u && t.h == 1u) return texel_at(sc, t, 0, 0);
    const float fu = u - floorf(u), fv = v - floorf(v);
    const float x = fu * (float)t.w - 0.5f, y = fv * (float)t.h - 0.5f;
    const float x0f = floorf(x), y0f = floorf(y);
    const float ax = x - x0f, ay = y - y0f;
    const int w = (int)t.w, h = (int)t.h;
    // wrap addressing without integer division: fu, fv are in [0, 1], so x0f is in [-1, w - 1] (and y0f in [-1, h - 1]) — the general
    // ((i % w) + w) % w of the oracle reduces to one compare each
    const int xi = (int)x0f, yi = (int)y0f;
    const int x0 = xi < 0 ? xi + w : (xi >= w ? xi - w : xi), y0 = yi < 0 ? yi + h : (yi >= h ? yi - h : yi);
    const int x1 = x0 + 1 == w ? 0 : x0 + 1, y1 = y0 + 1 == h ? 0 : y0 + 1;
    const float4 a = texel_at(sc, t, x0, y0), b = texel_at(sc, t, x1, y0), c = texel_at(sc, t, x0, y1), d = texel_at(sc, t, x1, y1);
    return mix4(mix4(a, b, ax), mix4(c, d, ax), ay);
}

// CDF::Get / BinarySearch (ReSTIRData.h:232-302): index of the light whose interval holds value*sum, and its pdf
LB_D void cdf_get(const SceneView& sc, float value, uint32_t& index, float& pdf) {
    const float required = sc.cdf_sum * value;
    int first = 0, last = (int)sc.num_lights - 1, center = 0;
    for (;;) {
        center = (last + first) / 2;
        const float higher = __ldg(&sc.cdf[center]), lower = center ? __ldg(&sc.cdf[center - 1]) : 0.f;
        if (required < lower && center - 1 >= first) { last = center - 1; continue; }
        if (required > higher && center + 1 <= last) { first = center + 1; continue; }
        break;
    }
    const float higher = __ldg(&sc.cdf[center]), lower = center ? __ldg(&sc.cdf[center - 1]) : 0.f;
    index = (uint32_t)center; pdf = (higher - lower) / sc.cdf_sum;
}

LB_D DevLight load_light(const SceneView& sc, uint32_t i) {
    const float4* p = reinterpret_cast<const float4*>(sc.lights + i);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    DevLight l;
    l.p0 = f3(a.x, a.y, a.z); l.p1 = f3(a.w, b.x, b.y); l.p2 = f3(b.z, b.w, c.x);
    l.normal = f3(c.y, c.z, c.w); l.radiance = f3(d.x, d.y, d.z); l.area = d.w;
    return l;
}

// GPUExtractSurfaceData.cu:8-228. hit_t <= 0 = miss. Fields not written by the reference for a branch stay zero.
LB_D Surface extract_surface(const SceneView& sc, const float3& ro, const float3& rd, const float3& throughput,
                             uint32_t inst, uint32_t prim, float hu, float hv, float hit_t) {
    Surface s;
    s.pos = f3(0.f); s.normal = f3(0.f); s.tangent = f3(0.f); s.incoming = f3(0.f); s.transport = f3(0.f); s.t = 0.f; s.flags = 0u;
    s.mat.color = make_float4(0.f, 0.f, 0.f, 0.f); s.mat.emissive = s.mat.color; s.mat.transmittance = s.mat.color; s.mat.tint = s.mat.color;
    s.mat.params = make_uint4(0u, 0u, 0u, 0u);
    if (!(hit_t > 0.f)) { s.flags = SURF_MISS; return s; }
    const DevEntry& e = sc.entries[inst];
    const DevMaterial& dm = sc.materials[e.material];
    const uint32_t ia = __ldg(&sc.indices[e.index_base + 3u * prim]) + e.vertex_base;
    const uint32_t ib = __ldg(&sc.indices[e.index_base + 3u * prim + 1u]) + e.vertex_base;
    const uint32_t ic = __ldg(&sc.indices[e.index_base + 3u * prim + 2u]) + e.vertex_base;
    const float U = hu, V = hv, W = 1.f - (U + V);
    const float4 na = __ldg(&sc.vtx_nu[ia]), nb = __ldg(&sc.vtx_nu[ib]), nc = __ldg(&sc.vtx_nu[ic]);
    const float4 ta = __ldg(&sc.vtx_tv[ia]), tb = __ldg(&sc.vtx_tv[ib]), tc = __ldg(&sc.vtx_tv[ic]);
    const float2 uv = (make_float2(na.w, ta.w) * W + make_float2(nb.w, tb.w) * U) + make_float2(nc.w, tc.w) * V;
    const float flip = __ldg(&sc.vtx_tw[ia]);
    const float4 nmap = tex2d(sc, dm.tex_normal, uv.x, uv.y), tcol = tex2d(sc, dm.tex_diffuse, uv.x, uv.y);
    float4 em = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.em_mode == 0 /* ENABLED */) { em = dm.mat.emissive * e.em_scale; em = em * tex2d(sc, dm.tex_emissive, uv.x, uv.y); }
    else if (e.em_mode == 2 /* OVERRIDE */) em = make_float4(e.em_r, e.em_g, e.em_b, e.em_scale) * e.em_scale;
    const float3 ln = xnormalize((f3(na) * W + f3(nb) * U) + f3(nc) * V);
    const float3 lt = xnormalize((f3(ta) * W + f3(tb) * U) + f3(tc) * V);
    const float3 nw = xnormalize(xform_vector(e.m, ln)), tw = xnormalize(xform_vector(e.m, lt));
    const float3 bw = cross(nw, tw) * flip;
    float3 nm = f3(nmap.x, nmap.y, nmap.z) * 2.f - f3(1.f);
    nm = xnormalize(nm);
    nm = xnormalize(f3(nm.x * tw.x + nm.y * bw.x + nm.z * nw.x, nm.x * tw.y + nm.y * bw.y + nm.z * nw.y, nm.x * tw.z + nm.y * bw.z + nm.z * nw.z));
    s.t = hit_t; s.normal = nm;
    if (em.x > 0.f || em.y > 0.f || em.z > 0.f) {
        const float mx = fmaxf(em.x, fmaxf(em.y, em.z)); const float inv = xdiv(1.0f, mx);
        s.mat.color = em * inv; s.flags |= SURF_EMISSIVE; return s;
    }
    s.pos = ro + rd * hit_t; s.incoming = rd; s.transport = throughput;
    if (tcol.w < 0.51f) { s.flags |= SURF_ALPHA; return s; }
    const float eta = xdiv(1.f, dm.mat.transmittance.w);
    s.tangent = tw; s.mat = dm.mat;
    const Shading mv(dm.mat);
    const float4 mr = tex2d(sc, dm.tex_mr, uv.x, uv.y);
    pack8(s.mat.params.x, mr.z * mv.metallic, 0);
    pack8(s.mat.params.x, mr.y * mv.roughness, 24);
    s.mat.color = tcol * dm.mat.color;
    const float4 cc = tex2d(sc, dm.tex_coat, uv.x, uv.y), ccr = tex2d(sc, dm.tex_coat_rough, uv.x, uv.y);
    const float4 tr = tex2d(sc, dm.tex_transmission, uv.x, uv.y), ti = tex2d(sc, dm.tex_tint, uv.x, uv.y);
    const float3 tint = f3(ti.x, ti.y, ti.z) * mv.tint;
    pack8(s.mat.params.z, mv.clearcoat * cc.x, 0);
    pack8(s.mat.params.z, mv.clearcoatgloss * (1.f - ccr.x), 8);
    s.mat.tint = f4(tint, s.mat.tint.w);
    pack8(s.mat.params.z, mv.transmission * tr.x, 16);
    s.mat.transmittance.w = eta;
    return s;
}

struct ShadowRayOut { float3 o, d, radiance; float tmax; };

// ShadeDirect (GPUShadeDirect.cu:42-153): one light from the CDF, uniform-ish point on it, unshadowed contribution.
// `seed` is the already hashed per-pixel stream (the volumetric march may have consumed numbers before).
LB_D bool nee_sample(const SceneView& sc, const Surface& s, uint32_t& seed, ShadowRayOut& out) {
    if (s.flags || sc.num_lights == 0u) return false;
    uint32_t li; float lpdf; cdf_get(sc, rand_f(seed), li, lpdf);
    const DevLight l = load_light(sc, li);
    const float u = rand_f(seed), v = rand_f(seed) * (1.f - u);
    const float3 point = l.p0 + ((l.p1 - l.p0) * u) + ((l.p2 - l.p0) * v);
    float3 dir = point - s.pos; const float dist = length(dir); dir /= dist;
    const float cos_in = fmaxf(dot(dir, s.normal), 0.f), cos_out = fmaxf(0.f, dot(l.normal, -dir));
    if (cos_in <= 0.f || dist <= 0.01f) return false;
    const float solid = (cos_out * l.area) / (dist * dist);
    float bpdf = 0.f; const float3 bsdf = bsdf_eval(s.mat, s.normal, s.tangent, -s.incoming, dir, bpdf);
    if (bpdf <= kBsdfEps) return false;
    float3 c = (bsdf / bpdf) * solid * cos_in * l.radiance;
    c *= ((1.f / lpdf) * s.transport);
    out.o = s.pos; out.d = dir; out.tmax = dist - 0.2f; out.radiance = c;
    return true;
}

struct BounceOut { float3 o, d, throughput; };

// ShadeIndirect (GPUShadeIndirect.cu:7-146)
LB_D bool bounce_sample(const Surface& s, uint32_t pixel_index, uint32_t seed_in, BounceOut& out) {
    uint32_t seed = wang_hash(seed_in + wang_hash(pixel_index));
    if (s.flags & SURF_ALPHA) { out.o = s.pos; out.d = s.incoming; out.throughput = s.transport; return true; }
    if (s.flags) return false;
    if (fabsf(dot(s.normal, s.incoming)) < 3.f * kBsdfEps) return false;
    float3 wi = f3(0.f); float pdf = 0.f; bool specular = false;
    const float r0 = rand_f(seed), r1 = rand_f(seed), r2 = rand_f(seed);
    const float3 bsdf = bsdf_sample(s.mat, s.normal, s.normal, s.tangent, -s.incoming, 1.f, r0, r1, r2, wi, pdf, specular);
    if (pdf <= kBsdfEps || isnan(pdf + bsdf.x + bsdf.y + bsdf.z)) return false;
    const float rr = specular ? 1.f : fminf(fmaxf(bsdf.x, fmaxf(bsdf.y, bsdf.z)), 1.f);
    const float rnd = rand_f(seed);
    if (rr < rnd) return false;
    float3 c = s.transport * (1.f / rr);
    c *= bsdf * fabsf(dot(s.normal, wi)) * (1.f / pdf);
    out.o = s.pos; out.d = wi; out.throughput = c;
    return true;
}

// Resample (ReSTIRKernels.cu:1259-1325): re-evaluate a light sample's unshadowed contribution at a pixel.
// Split in two so that a caller evaluating many samples against one pixel (RIS: 32, spatial reuse: up to 5) builds the BSDF
// context once, and so that the cheap geometric rejection can run ahead of the expensive BSDF evaluation (k_ris).
struct ResampleGeom { float3 dir; float solid, cos_in; };
// false = rejected (light below the horizon / facing away / too close): the sample's pdf becomes 0, nothing else changes
LB_D bool resample_geom(const float3& lpos, const float3& lnormal, float larea, const float3& ppos, const float3& pnormal, ResampleGeom& g) {
    float3 dir = lpos - ppos; const float dist = length(dir); dir /= dist;
    const float cos_in = fmaxf(dot(dir, pnormal), 0.f), cos_out = fmaxf(dot(lnormal, -dir), 0.f);
    if (cos_in <= 0 || cos_out <= 0 || dist <= 0.01f) return false;
    g.dir = dir; g.cos_in = cos_in; g.solid = (cos_out * larea) / (dist * dist);
    return true;
}
// MODE (what the caller has established about the material, lb_bsdf.cuh): 0 nothing, 1 isotropic roughness, 2 BsdfCtx::is_simple()
template <int MODE = 0>
LB_D void resample_shade(const BsdfCtx& ctx, const ResampleGeom& g, LightSample& out) {
    float pdf = 0.f; const float3 bsdf = MODE == 2 ? ctx.eval_simple(g.dir, pdf) : (MODE == 1 ? ctx.eval<true>(g.dir, pdf) : ctx.eval<false>(g.dir, pdf));
    const float added = pdf + bsdf.x + bsdf.y + bsdf.z;
    if (pdf <= kBsdfEps || isnan(added) || isinf(added)) { out.contribution = f3(0.f); out.pdf = 0; return; }
    const float3 c = (bsdf / pdf) * g.solid * g.cos_in * out.radiance;
    out.contribution = c; out.pdf = (c.x + c.y + c.z) / 3.f;
}
LB_D void resample(const LightSample& in, const float3& ppos, const float3& pnormal, const BsdfCtx& ctx, LightSample& out) {
    out = in;
    ResampleGeom g;
    if (!resample_geom(in.position, in.normal, in.area, ppos, pnormal, g)) { out.pdf = 0; return; }
#if LB_ISO_PER_LANE
    // per lane (temporal / spatial reuse, buffer merge): isotropic materials — nearly all — take the evaluation that has no anisotropic terms
    // instead of executing them predicated-off
    if (ctx.is_isotropic()) resample_shade<1>(ctx, g, out); else resample_shade<0>(ctx, g, out);
#else
    resample_shade(ctx, g, out);
#endif
}
LB_D BsdfCtx surface_ctx(const Surface& px) { return BsdfCtx(px.mat, px.normal, px.tangent, -px.incoming); }

} // namespace lb
