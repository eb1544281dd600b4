// C-callable wrapper over the UNMODIFIED reference BSDF headers, compiled for the host.
// Built by oracle/Makefile into oracle/_ref/libref_bsdf.so; used only by tests to pin the oracle.
#include "shim.h"
// Compile the DEVICE branches of the reference headers (ggxmdf.cuh:90,208 sincosf ordering; disney.cuh:175 / frosted.cuh:126 default
// `adjoint` parameter): that is the code the reference renders with. All system / CUDA headers are already included by shim.h.
#define __CUDACC__ 1
#include "CUDAKernels/RandomUtilities.cuh"
#include "CUDAKernels/disney.cuh"

extern "C" {

struct RefMaterial { float color[4], emissive[4], transmittance[4], tint[4]; unsigned params[4]; };
static_assert(sizeof(RefMaterial) == sizeof(MaterialData), "layout");

unsigned ref_wang_hash(unsigned s) { return WangHash(s); }
unsigned ref_random_int(unsigned* s) { return RandomInt(*s); }
float ref_random_float(unsigned* s) { return RandomFloat(*s); }

// packs the 13 byte-quantised parameters exactly like the reference setters do
void ref_pack_material(RefMaterial* out, const float* color4, const float* transmittance3, const float* tint3,
                       float luminance, float ior, float metallic, float subsurface, float specular, float roughness,
                       float spectint, float anisotropic, float sheen, float sheentint, float clearcoat,
                       float clearcoatgloss, float transmission)
{
    MaterialData m(0.f);
    m.SetColor(make_float4(color4[0], color4[1], color4[2], color4[3]));
    m.SetTransmittance(make_float3(transmittance3[0], transmittance3[1], transmittance3[2]));
    m.SetTint(make_float3(tint3[0], tint3[1], tint3[2]));
    m.SetLuminance(luminance); m.SetRefractiveIndex(ior);
    m.SetMetallic(metallic); m.SetSubSurface(subsurface); m.SetSpecular(specular); m.SetRoughness(roughness);
    m.SetSpecTint(spectint); m.SetAnisotropic(anisotropic); m.SetSheen(sheen); m.SetSheenTint(sheentint);
    m.SetClearCoat(clearcoat); m.SetClearCoatGloss(clearcoatgloss); m.SetTransmission(transmission);
    *reinterpret_cast<MaterialData*>(out) = m;
}

void ref_evaluate_bsdf(const RefMaterial* m, const float* iN, const float* iT, const float* wow, const float* wiw,
                       float* bsdf3, float* pdf)
{
    float p = 0.f;
    const float3 v = EvaluateBSDF(*reinterpret_cast<const MaterialData*>(m), make_float3(iN[0], iN[1], iN[2]),
        make_float3(iT[0], iT[1], iT[2]), make_float3(wow[0], wow[1], wow[2]), make_float3(wiw[0], wiw[1], wiw[2]), p);
    bsdf3[0] = v.x; bsdf3[1] = v.y; bsdf3[2] = v.z; *pdf = p;
}

void ref_sample_bsdf(const RefMaterial* m, const float* iN, const float* N, const float* iT, const float* wow,
                     float distance, float r0, float r1, float r2, float* bsdf3, float* wiw3, float* pdf, int* specular)
{
    float p = 0.f; bool spec = false; float3 wiw = make_float3(0.f);
    const float3 v = SampleBSDF(*reinterpret_cast<const MaterialData*>(m), make_float3(iN[0], iN[1], iN[2]),
        make_float3(N[0], N[1], N[2]), make_float3(iT[0], iT[1], iT[2]), make_float3(wow[0], wow[1], wow[2]),
        distance, r0, r1, r2, wiw, p, spec);
    bsdf3[0] = v.x; bsdf3[1] = v.y; bsdf3[2] = v.z; wiw3[0] = wiw.x; wiw3[1] = wiw.y; wiw3[2] = wiw.z;
    *pdf = p; *specular = spec ? 1 : 0;
}

}
