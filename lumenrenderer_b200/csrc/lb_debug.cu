// lb_debug.cu — debug taps used by the parity tests (lb_debug_* of include/lumen_b200.h): single-ray traces, stand-alone BSDF evaluation
// and sampling, read-back of surface / reservoir / hit records. Compiled in the EXACT arithmetic class (csrc/Makefile): the stand-alone
// BSDF evaluation is compared with the reference headers' golden vectors at 1e-5.
#include "lb_kernels.h"
#include "lb_trace.cuh"
#include "lb_shade.cuh"

namespace lb {

namespace {

// ------------------------------------------------------------------ debug taps (parity tests)
struct Hit20 { uint32_t inst, prim; float u, v, t; };

__global__ void k_debug_trace(BvhView bvh, const float* __restrict__ rays6, const float* __restrict__ tmaxs, uint32_t n, float tmin, float tmax, Hit20* hits, uint8_t* occ) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 o = f3(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]), d = f3(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]);
    HitInfo h;
    if (occ) { occ[i] = bvh8_trace<true>(bvh, o, d, tmin, tmaxs[i], h) ? 1 : 0; return; }
    if (bvh8_trace<false>(bvh, o, d, tmin, tmax, h)) hits[i] = Hit20{h.inst, h.prim, h.u, h.v, h.t};
    else hits[i] = Hit20{0u, 0u, 0.f, 0.f, -1.f};
}

__device__ Material unpack_mat24(const float* m) {
    Material p;
    p.color = make_float4(m[0], m[1], m[2], m[3]); p.transmittance = make_float4(m[4], m[5], m[6], m[7]); p.tint = make_float4(m[8], m[9], m[10], m[11]);
    p.emissive = make_float4(0.f, 0.f, 0.f, 0.f); p.params = make_uint4(0u, 0u, 0u, 0u);
    pack8(p.params.x, m[12], 0); pack8(p.params.x, m[13], 8); pack8(p.params.x, m[14], 16); pack8(p.params.x, m[15], 24);
    pack8(p.params.y, m[16], 0); pack8(p.params.y, m[17], 8); pack8(p.params.y, m[18], 16); pack8(p.params.y, m[19], 24);
    pack8(p.params.z, m[20], 0); pack8(p.params.z, m[21], 8); pack8(p.params.z, m[22], 16);
    return p;
}
__global__ void k_debug_bsdf(const float* __restrict__ mat24, const float* __restrict__ v12, uint32_t n, float* out, int sample) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Material m = unpack_mat24(mat24);
    const float* v = v12 + 12 * i;
    const float3 nrm = f3(v[0], v[1], v[2]), tan = f3(v[3], v[4], v[5]), wo = f3(v[6], v[7], v[8]);
    if (!sample) {
        const BsdfCtx c(m, nrm, tan, wo); const float3 wi = f3(v[9], v[10], v[11]);
        float pdf = 0.f; float3 b = c.eval(wi, pdf);
        // the lean evaluation RIS uses for materials without transmission / sheen / clear coat / anisotropy / subsurface must reproduce the general
        // one exactly in this unit (no contraction, IEEE division): a mismatch poisons the output, which fails the golden-vector tests
        if (c.is_simple()) {
            float p2 = 0.f; const float3 b2 = c.eval_simple(wi, p2);
            const bool same = __float_as_uint(b2.x + 0.f) == __float_as_uint(b.x + 0.f) && __float_as_uint(b2.y + 0.f) == __float_as_uint(b.y + 0.f) &&
                              __float_as_uint(b2.z + 0.f) == __float_as_uint(b.z + 0.f) && __float_as_uint(p2 + 0.f) == __float_as_uint(pdf + 0.f);
            if (!same) { b = f3(__int_as_float(0x7fc00000)); pdf = __int_as_float(0x7fc00000); }
        }
        if (c.is_isotropic()) {                      // likewise the evaluation without the anisotropic microfacet terms
            float p3 = 0.f; const float3 b3 = c.eval<true>(wi, p3);
            const bool same = __float_as_uint(b3.x + 0.f) == __float_as_uint(b.x + 0.f) && __float_as_uint(b3.y + 0.f) == __float_as_uint(b.y + 0.f) &&
                              __float_as_uint(b3.z + 0.f) == __float_as_uint(b.z + 0.f) && __float_as_uint(p3 + 0.f) == __float_as_uint(pdf + 0.f);
            if (!same && pdf == pdf) { b = f3(__int_as_float(0x7fc00000)); pdf = __int_as_float(0x7fc00000); }
        }
        out[4 * i] = b.x; out[4 * i + 1] = b.y; out[4 * i + 2] = b.z; out[4 * i + 3] = pdf;
    } else {
        float pdf = 0.f; bool spec = false; float3 wi = f3(0.f);
        const float3 b = bsdf_sample(m, nrm, nrm, tan, wo, 1.f, v[9], v[10], v[11], wi, pdf, spec);
        float* o = out + 8 * i; o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = wi.x; o[4] = wi.y; o[5] = wi.z; o[6] = pdf; o[7] = spec ? 1.f : 0.f;
    }
}
__global__ void k_debug_surface(const float4* __restrict__ planes, uint32_t npix, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    Surface s; surface_load(planes, npix, i, s);
    float* o = out + 24 * (size_t)i;
    o[0] = s.pos.x; o[1] = s.pos.y; o[2] = s.pos.z; o[3] = s.t; o[4] = s.normal.x; o[5] = s.normal.y; o[6] = s.normal.z; o[7] = (float)s.flags;
    o[8] = s.tangent.x; o[9] = s.tangent.y; o[10] = s.tangent.z; o[11] = 0; o[12] = s.incoming.x; o[13] = s.incoming.y; o[14] = s.incoming.z; o[15] = 0;
    o[16] = s.transport.x; o[17] = s.transport.y; o[18] = s.transport.z; o[19] = 0; o[20] = s.mat.color.x; o[21] = s.mat.color.y; o[22] = s.mat.color.z; o[23] = s.mat.color.w;
}
__global__ void k_debug_reservoirs(const float4* __restrict__ planes, uint32_t npix, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    Reservoir q; reservoir_load(planes, npix, i, q);
    float* o = out + 20 * (size_t)i;
    o[0] = q.weight_sum; o[1] = q.weight; o[2] = (float)q.count; o[3] = q.s.pdf; o[4] = q.s.position.x; o[5] = q.s.position.y; o[6] = q.s.position.z; o[7] = q.s.area;
    o[8] = q.s.normal.x; o[9] = q.s.normal.y; o[10] = q.s.normal.z; o[11] = 0; o[12] = q.s.radiance.x; o[13] = q.s.radiance.y; o[14] = q.s.radiance.z; o[15] = 0;
    o[16] = q.s.contribution.x; o[17] = q.s.contribution.y; o[18] = q.s.contribution.z; o[19] = 0;
}
__global__ void k_debug_hits(const uint4* __restrict__ hits, uint32_t n, Hit20* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 h = hits[i];
    const __half2 b = *reinterpret_cast<const __half2*>(&h.z);
    const float t = __uint_as_float(h.w);
    out[i] = t > 0.f ? Hit20{h.x, h.y, __low2float(b), __high2float(b), t} : Hit20{0u, 0u, 0.f, 0.f, -1.f};
}

} // namespace

void launch_debug_trace(const LaunchCfg& cfg, const BvhView& bvh, const float* rays6, const float* tmaxs, uint32_t n, float tmin, float tmax, void* hits20, uint8_t* occluded) {
    if (!n) return;
    k_debug_trace<<<grid_for(n, 128), 128, 0, cfg.stream>>>(bvh, rays6, tmaxs, n, tmin, tmax, (Hit20*)hits20, occluded); LB_LAUNCH_CHECK();
}
void launch_debug_bsdf(const LaunchCfg& cfg, const float* mat24, const float* v12, uint32_t n, float* out, bool sample) {
    if (!n) return;
    k_debug_bsdf<<<grid_for(n, 128), 128, 0, cfg.stream>>>(mat24, v12, n, out, sample ? 1 : 0); LB_LAUNCH_CHECK();
}
void launch_debug_surface(const LaunchCfg& cfg, const float4* planes, uint32_t npix, float* out24) {
    k_debug_surface<<<grid_for(npix, 256), 256, 0, cfg.stream>>>(planes, npix, out24); LB_LAUNCH_CHECK();
}
void launch_debug_reservoirs(const LaunchCfg& cfg, const float4* planes, uint32_t npix, float* out20) {
    k_debug_reservoirs<<<grid_for(npix, 256), 256, 0, cfg.stream>>>(planes, npix, out20); LB_LAUNCH_CHECK();
}
void launch_debug_hits(const LaunchCfg& cfg, const uint4* hits, uint32_t n, void* hits20) {
    k_debug_hits<<<grid_for(n, 256), 256, 0, cfg.stream>>>(hits, n, (Hit20*)hits20); LB_LAUNCH_CHECK();
}

} // namespace lb
