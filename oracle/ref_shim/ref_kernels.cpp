// ref_kernels.cpp — TEST INFRASTRUCTURE. C-callable wrapper over the reference's own radiance-deciding device code, compiled for the host
// IN PLACE by oracle/Makefile into oracle/_ref/libref_kernels.so:
//   Resample, CombineBiased, CombineUnbiased   LumenPT/src/CUDAKernels/ReSTIRKernels.cu:1123-1325
//   FillLightBagsInternal :343-370, PickPrimarySamplesInternal :402-522, GenerateShadowRay :546-582, SpatialNeighbourSamplingInternal :787-980,
//   CombineTemporalSamplesInternal :1015-1121, CombineReservoirBuffersInternal :1407-1436   (the whole __global__ bodies, same file)
//   ShadeIndirect (the whole __global__ body)  LumenPT/src/CUDAKernels/WaveFrontKernels/GPUShadeIndirect.cu:7-146
//   ShadeDirect   (the whole __global__ body)  LumenPT/src/CUDAKernels/WaveFrontKernels/GPUShadeDirect.cu:42-153
// The three translation units cannot be compiled whole (thrust, MemoryBuffer, OptiX device headers and two MSVC-only constructs in headers they
// include), so the Makefile cuts exactly those line ranges out of the reference files WHERE THEY LIE (sed -n 'a,bp') into
// oracle/_ref/gen/*.inc at build time — git-ignored build products, nothing of the reference is committed — and this file supplies the
// environment nvcc would have: the reference's own struct headers (included in place), blockIdx / blockDim / threadIdx, atomicAdd.
// Used by tests/golden/make_golden_kernels.py to record known answers that pin oracle/oracle.cpp's restatement of these functions.
#include <cstring>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_fp16.h>                 // host flavour first: the headers below re-include it under __CUDACC__
using std::isnan; using std::isinf;
#include <Optix/optix.h>               // host API only
#include <Cuda/cuda/helpers.h>
struct ref_dim3 { unsigned x, y, z; };
static thread_local ref_dim3 blockIdx, blockDim, threadIdx, gridDim;
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p += v; return o; }
#define __CUDACC__ 1
#include "CudaDefines.h"
#include "MaterialStructs.h"
#include <nanovdb/NanoVDB.h>
#include "WaveFrontDataStructs/SurfaceData.h"
#include "WaveFrontDataStructs/AtomicBuffer.h"
#include "WaveFrontDataStructs/IntersectionRayData.h"
#include "WaveFrontDataStructs/ShadowRayData.h"
#include "WaveFrontDataStructs/LightData.h"
#include "WaveFrontDataStructs/VolumetricData.h"
#include "ReSTIRData.h"
#include "CUDAKernels/RandomUtilities.cuh"
#include "CUDAKernels/disney.cuh"
using namespace WaveFront;
#define PIXEL_DATA_INDEX(PIXELX, PIXELY, WIDTH) ((PIXELY * WIDTH) + PIXELX)       // WaveFrontDataStructs.h:13
// declared by ReSTIRKernels.cuh (not includable: MemoryBuffer, thrust)
__device__ __inline__ void Resample(LightSample* a_Input, const WaveFront::SurfaceData* a_PixelData, LightSample* a_Output);
// GPUVolumetricShadeDirect.cu:8-101 marches the pixel's volume segment; the known answers here have no volume (entry == exit), for which
// the reference's function draws no random number and appends nothing — that early-out is all this stand-in reproduces
static inline void VolumetricShadeDirect(PixelIndex, const uint3, const VolumetricData*, AtomicBuffer<ShadowRayData>* const, const AtomicBuffer<TriangleLight>* const,
                                         unsigned int&, const CDF* const, cudaSurfaceObject_t) {}
#include "Half2.h"
#define RESERVOIR_INDEX(INTERSECTION_INDEX, DEPTH, MAX_DEPTH) (((INTERSECTION_INDEX) * (MAX_DEPTH)) + (DEPTH))      // ReSTIRKernels.cuh:17-18
#define PIXEL_INDEX(X, Y, WIDTH) (((Y) * (WIDTH) + (X)))
// hazard 1: the reference picks the light bag by the hardware SM id; the canonical choice (the alternative its own comment gives, :423) is the
// block index — which is what this stand-in returns, so the kernel below is run with 256-thread blocks
static inline uint32_t __mysmid() { return blockIdx.x; }
// a cudaSurfaceObject_t here is the address of a plain host image; the two surface accesses of the temporal kernel go to it
struct RefSurface { unsigned char* data; size_t pitch; };
template <class T> static inline void surf2Dread(T* out, cudaSurfaceObject_t s, int xbytes, int y, int) { const RefSurface* f = reinterpret_cast<const RefSurface*>(s); memcpy(out, f->data + (size_t)y * f->pitch + xbytes, sizeof(T)); }
// ShadeReservoirs (:618-665) adds contribution * weight / 3 to a half4 output image. The channels of this framework are fp32 (canonical choice
// 2), so the stand-in performs the same sum in fp32 into a float4 image; what the temporal kernel shades, and when, is the reference's.
static inline void ShadeReservoirs(Reservoir* a_Reservoirs, unsigned a_Width, unsigned a_InputX, unsigned a_InputY, unsigned a_OutputX, unsigned a_OutputY, cudaSurfaceObject_t a_OutputBuffer)
{
    constexpr auto numShadedSamples = ReSTIRSettings::numReservoirsPerPixel * (1 + (ReSTIRSettings::enableTemporal ? 1 : 0) + (ReSTIRSettings::enableSpatial ? 1 : 0));
    const RefSurface* f = reinterpret_cast<const RefSurface*>(a_OutputBuffer);
    float4* px = reinterpret_cast<float4*>(f->data + (size_t)a_OutputY * f->pitch) + a_OutputX;
    const Reservoir& r = a_Reservoirs[PIXEL_DATA_INDEX(a_InputX, a_InputY, a_Width)];
    if (r.weight > 0.f) { const float3 c = r.sample.unshadowedPathContribution * (r.weight / static_cast<float>(numShadedSamples)); px->x += c.x; px->y += c.y; px->z += c.z; }
}
#include "restir_device.inc"           // oracle/_ref/gen: ReSTIRKernels.cu:1123-1325
#include "restir_bags.inc"             // :343-370
#include "restir_ris.inc"              // :402-522
#include "restir_shadow_ray.inc"       // :546-582
#include "restir_spatial.inc"          // :787-980
#include "restir_temporal.inc"         // :1015-1121
#include "restir_combine_buffers.inc"  // :1407-1436
#include "shade_indirect.inc"          // oracle/_ref/gen: GPUShadeIndirect.cu:7-146
#include "shade_direct.inc"            // oracle/_ref/gen: GPUShadeDirect.cu:42-153

namespace {

// mat24: color4, transmittance3, ior, tint3, luminance, metallic, subsurface, specular, roughness, spectint, anisotropic, sheen, sheentint,
//        clearcoat, clearcoatgloss, transmission, pad — packed through the reference's own setters (MaterialStructs.h:84-260)
MaterialData make_material(const float* m)
{
    MaterialData d(0.f);
    d.SetColor(make_float4(m[0], m[1], m[2], m[3]));
    d.SetTransmittance(make_float3(m[4], m[5], m[6])); d.SetRefractiveIndex(m[7]);
    d.SetTint(make_float3(m[8], m[9], m[10])); d.SetLuminance(m[11]);
    d.SetMetallic(m[12]); d.SetSubSurface(m[13]); d.SetSpecular(m[14]); d.SetRoughness(m[15]);
    d.SetSpecTint(m[16]); d.SetAnisotropic(m[17]); d.SetSheen(m[18]); d.SetSheenTint(m[19]);
    d.SetClearCoat(m[20]); d.SetClearCoatGloss(m[21]); d.SetTransmission(m[22]);
    return d;
}
// surf44: position3, normal3, tangent3, incoming3, transport3, t, flags, pad3, mat24
SurfaceData make_surface(const float* s, unsigned px, unsigned py)
{
    SurfaceData d;
    memset(&d, 0, sizeof d);
    d.m_PixelIndex = PixelIndex{px, py};
    d.m_Position = make_float3(s[0], s[1], s[2]); d.m_Normal = make_float3(s[3], s[4], s[5]); d.m_GeometricNormal = d.m_Normal;
    d.m_Tangent = make_float3(s[6], s[7], s[8]); d.m_IncomingRayDirection = make_float3(s[9], s[10], s[11]);
    d.m_TransportFactor = make_float3(s[12], s[13], s[14]); d.m_IntersectionT = s[15];
    d.m_SurfaceFlags = static_cast<SurfaceFlags>((unsigned char)s[16]);
    d.m_MaterialData = make_material(s + 20);
    return d;
}
// sample14: radiance3, normal3, position3, area, contribution3, pdf
LightSample make_sample(const float* s)
{
    LightSample l;
    l.radiance = make_float3(s[0], s[1], s[2]); l.normal = make_float3(s[3], s[4], s[5]); l.position = make_float3(s[6], s[7], s[8]);
    l.area = s[9]; l.unshadowedPathContribution = make_float3(s[10], s[11], s[12]); l.solidAnglePdf = s[13];
    return l;
}
void put_sample(const LightSample& l, float* o)
{
    o[0] = l.radiance.x; o[1] = l.radiance.y; o[2] = l.radiance.z; o[3] = l.normal.x; o[4] = l.normal.y; o[5] = l.normal.z;
    o[6] = l.position.x; o[7] = l.position.y; o[8] = l.position.z; o[9] = l.area;
    o[10] = l.unshadowedPathContribution.x; o[11] = l.unshadowedPathContribution.y; o[12] = l.unshadowedPathContribution.z; o[13] = l.solidAnglePdf;
}
template <class T> AtomicBuffer<T>* make_buffer(unsigned cap)
{
    auto* b = static_cast<AtomicBuffer<T>*>(calloc(1, sizeof(AtomicBuffer<T>) + sizeof(T) * (cap + 1)));
    b->counter = 0; b->maxSize = cap;
    return b;
}

}

// ---- whole ReSTIR kernels over a w x h frame. surfs44 per pixel; reservoirs as reservoir17 per pixel; lights / CDF as in ref_kat_shade_direct.
static void put_reservoir(const Reservoir& r, float* o) { o[0] = r.weightSum; o[1] = (float)r.sampleCount; o[2] = r.weight; put_sample(r.sample, o + 3); }
static Reservoir make_reservoir(const float* r) { Reservoir q; q.weightSum = r[0]; q.sampleCount = (long long)r[1]; q.weight = r[2]; q.sample = make_sample(r + 3); return q; }
static SurfaceData* make_surfaces(const float* surfs44, unsigned w, unsigned h)
{
    SurfaceData* s = static_cast<SurfaceData*>(calloc((size_t)w * h, sizeof(SurfaceData)));
    for (unsigned i = 0; i < w * h; ++i) s[i] = make_surface(surfs44 + 44 * (size_t)i, i % w, i / w);
    return s;
}
template <class F> static void launch_1d(unsigned n, F&& f)          // <<<ceil(n / 256), 256>>>
{
    blockDim = {256, 1, 1}; gridDim = {(n + 255u) / 256u, 1, 1};
    for (unsigned b = 0; b < gridDim.x; ++b) for (unsigned t = 0; t < 256u; ++t) { blockIdx = {b, 0, 0}; threadIdx = {t, 0, 0}; f(); }
}

extern "C" {

void ref_kat_resample(const float* samples14, unsigned n, const float* surf44, float* out14)
{
    const SurfaceData px = make_surface(surf44, 0, 0);
    for (unsigned k = 0; k < n; ++k) { LightSample in = make_sample(samples14 + 14 * k), out; Resample(&in, &px, &out); put_sample(out, out14 + 14 * k); }
}

// reservoir17: weightSum, sampleCount, weight, sample14.  surfs44: the pixels the n reservoirs came from (CombineUnbiased only)
void ref_kat_combine(const float* res17, unsigned n, const float* surf44, const float* surfs44, unsigned seed, int unbiased, float* out17)
{
    const SurfaceData px = make_surface(surf44, 0, 0);
    Reservoir* in = static_cast<Reservoir*>(calloc(n, sizeof(Reservoir)));
    SurfaceData* from = static_cast<SurfaceData*>(calloc(n, sizeof(SurfaceData)));
    for (unsigned k = 0; k < n; ++k) {
        const float* r = res17 + 17 * k;
        in[k].weightSum = r[0]; in[k].sampleCount = (long long)r[1]; in[k].weight = r[2]; in[k].sample = make_sample(r + 3);
        if (unbiased) from[k] = make_surface(surfs44 + 44 * k, 0, 0);
    }
    Reservoir out;
    if (unbiased) CombineUnbiased(&out, &px, (int)n, in, from, seed); else CombineBiased(&out, (int)n, in, &px, seed);
    out17[0] = out.weightSum; out17[1] = (float)out.sampleCount; out17[2] = out.weight; put_sample(out.sample, out17 + 3);
    free(in); free(from);
}

// FillLightBags + PickPrimarySamples: bags_out = (light index recovered from p0, pdf) per entry, reservoirs_out = reservoir17 per pixel
void ref_kat_ris(const float* surfs44, unsigned w, unsigned h, unsigned bag_seed, unsigned ris_seed, const float* lights16, const float* cdf_weights, unsigned nlights,
                 float* bag_pdf_out, float* bag_p0x_out, float* reservoirs_out)
{
    SurfaceData* s = make_surfaces(surfs44, w, h);
    auto* lights = make_buffer<TriangleLight>(nlights);
    for (unsigned k = 0; k < nlights; ++k) {
        const float* l = lights16 + 16 * k; TriangleLight t;
        t.p0 = make_float3(l[0], l[1], l[2]); t.p1 = make_float3(l[3], l[4], l[5]); t.p2 = make_float3(l[6], l[7], l[8]);
        t.normal = make_float3(l[9], l[10], l[11]); t.radiance = make_float3(l[12], l[13], l[14]); t.area = l[15];
        lights->data[k] = t;
    }
    lights->counter = nlights;
    CDF* cdf = static_cast<CDF*>(calloc(1, sizeof(CDF) + sizeof(float) * (nlights + 1)));
    cdf->Reset();
    for (unsigned k = 0; k < nlights; ++k) cdf->Insert(cdf_weights[k]);
    const unsigned nb = ReSTIRSettings::numLightBags, per = ReSTIRSettings::numLightsPerBag;
    LightBagEntry* bags = static_cast<LightBagEntry*>(calloc((size_t)nb * per, sizeof(LightBagEntry)));
    launch_1d(nb * per, [&]() { FillLightBagsInternal(nb, per, cdf, bags, lights, bag_seed); });
    for (unsigned i = 0; i < nb * per; ++i) { bag_pdf_out[i] = bags[i].pdf; bag_p0x_out[i] = bags[i].light.p0.x; }
    Reservoir* res = static_cast<Reservoir*>(calloc((size_t)w * h, sizeof(Reservoir)));
    launch_1d(w * h, [&]() { PickPrimarySamplesInternal(bags, res, ReSTIRSettings::numPrimarySamples, w * h, nb, per, s, ris_seed); });
    for (unsigned i = 0; i < w * h; ++i) put_reservoir(res[i], reservoirs_out + 17 * (size_t)i);
    free(res); free(bags); free(cdf); free(lights); free(s);
}
// GenerateShadowRay: rays8 = reservoir index, origin3, direction3, distance, in append order; returns the count
unsigned ref_kat_visibility_rays(const float* surfs44, const float* reservoirs17, unsigned w, unsigned h, float* rays8)
{
    SurfaceData* s = make_surfaces(surfs44, w, h);
    Reservoir* res = static_cast<Reservoir*>(calloc((size_t)w * h, sizeof(Reservoir)));
    for (unsigned i = 0; i < w * h; ++i) res[i] = make_reservoir(reservoirs17 + 17 * (size_t)i);
    auto* out = make_buffer<RestirShadowRay>(w * h);
    launch_1d(w * h, [&]() { GenerateShadowRay(out, res, s, w * h); });
    for (unsigned k = 0; k < out->counter; ++k) { const RestirShadowRay& r = out->data[k]; float* o = rays8 + 8 * k;
        o[0] = (float)r.index; o[1] = r.origin.x; o[2] = r.origin.y; o[3] = r.origin.z; o[4] = r.direction.x; o[5] = r.direction.y; o[6] = r.direction.z; o[7] = r.distance; }
    const unsigned n = out->counter;
    free(out); free(res); free(s);
    return n;
}
// CombineTemporalSamplesInternal: cur/prev surfaces and reservoirs, half2 motion vectors given as floats; the current reservoirs are updated in
// place (returned in cur_out17) and the shading of the previous reservoirs is added to direct4 (fp32, see ShadeReservoirs above)
void ref_kat_temporal(const float* cur44, const float* prev44, const float* cur17, const float* prev17, const float* motion2, unsigned w, unsigned h, unsigned seed,
                      float* cur_out17, float* direct4)
{
    SurfaceData* sc = make_surfaces(cur44, w, h); SurfaceData* sp = make_surfaces(prev44, w, h);
    Reservoir* rc = static_cast<Reservoir*>(calloc((size_t)w * h, sizeof(Reservoir))); Reservoir* rp = static_cast<Reservoir*>(calloc((size_t)w * h, sizeof(Reservoir)));
    for (unsigned i = 0; i < w * h; ++i) { rc[i] = make_reservoir(cur17 + 17 * (size_t)i); rp[i] = make_reservoir(prev17 + 17 * (size_t)i); }
    half2* mv = static_cast<half2*>(calloc((size_t)w * h, sizeof(half2)));
    for (unsigned i = 0; i < w * h; ++i) mv[i] = __floats2half2_rn(motion2[2 * i], motion2[2 * i + 1]);
    RefSurface mvs{reinterpret_cast<unsigned char*>(mv), (size_t)w * sizeof(half2)}, outs{reinterpret_cast<unsigned char*>(direct4), (size_t)w * sizeof(float4)};
    launch_1d(w * h, [&]() { CombineTemporalSamplesInternal(rc, rp, sc, sp, seed, w * h, make_uint2(w, h), reinterpret_cast<cudaSurfaceObject_t>(&mvs), reinterpret_cast<cudaSurfaceObject_t>(&outs)); });
    for (unsigned i = 0; i < w * h; ++i) put_reservoir(rc[i], cur_out17 + 17 * (size_t)i);
    free(mv); free(rp); free(rc); free(sp); free(sc);
}
// SpatialNeighbourSamplingInternal, one iteration: out17 holds the previous content of the output buffer on entry (pixels with fewer than two
// similar neighbours are Reset(), which keeps the stored sample) and the result on return
void ref_kat_spatial(const float* surfs44, const float* in17, unsigned w, unsigned h, unsigned seed, float* out17)
{
    SurfaceData* s = make_surfaces(surfs44, w, h);
    Reservoir* in = static_cast<Reservoir*>(calloc((size_t)w * h, sizeof(Reservoir))); Reservoir* out = static_cast<Reservoir*>(calloc((size_t)w * h, sizeof(Reservoir)));
    for (unsigned i = 0; i < w * h; ++i) { in[i] = make_reservoir(in17 + 17 * (size_t)i); out[i] = make_reservoir(out17 + 17 * (size_t)i); }
    launch_1d(w * h, [&]() { SpatialNeighbourSamplingInternal(in, out, s, seed, make_uint2(w, h), w * h); });
    for (unsigned i = 0; i < w * h; ++i) put_reservoir(out[i], out17 + 17 * (size_t)i);
    free(out); free(in); free(s);
}
// CombineReservoirBuffersInternal: a17 <- CombineBiased(a17, b17) per pixel
void ref_kat_combine_buffers(const float* surfs44, float* a17, const float* b17, unsigned w, unsigned h, unsigned seed)
{
    SurfaceData* s = make_surfaces(surfs44, w, h);
    Reservoir* a = static_cast<Reservoir*>(calloc((size_t)w * h, sizeof(Reservoir))); Reservoir* b = static_cast<Reservoir*>(calloc((size_t)w * h, sizeof(Reservoir)));
    for (unsigned i = 0; i < w * h; ++i) { a[i] = make_reservoir(a17 + 17 * (size_t)i); b[i] = make_reservoir(b17 + 17 * (size_t)i); }
    launch_1d(w * h, [&]() { CombineReservoirBuffersInternal(a, b, s, w * h, seed); });
    for (unsigned i = 0; i < w * h; ++i) put_reservoir(a[i], a17 + 17 * (size_t)i);
    free(b); free(a); free(s);
}

// The ShadeIndirect kernel over a w x h grid of surfaces, one "thread" per pixel in launch order. rays11: px, py, origin3, direction3, contribution3.
unsigned ref_kat_shade_indirect(const float* surfs44, unsigned w, unsigned h, unsigned seed, float* rays11, unsigned cap)
{
    SurfaceData* s = static_cast<SurfaceData*>(calloc((size_t)w * h, sizeof(SurfaceData)));
    for (unsigned i = 0; i < w * h; ++i) s[i] = make_surface(surfs44 + 44 * (size_t)i, i % w, i / w);
    auto* out = make_buffer<IntersectionRayData>(w * h);
    blockDim = {1, 1, 1}; gridDim = {w, h, 1}; threadIdx = {0, 0, 0};
    for (unsigned y = 0; y < h; ++y) for (unsigned x = 0; x < w; ++x) { blockIdx = {x, y, 0}; ShadeIndirect(make_uint3(w, h, 0), s, out, seed); }
    const unsigned n = out->counter < cap ? out->counter : cap;
    for (unsigned k = 0; k < n; ++k) {
        const IntersectionRayData& r = out->data[k]; float* o = rays11 + 11 * k;
        o[0] = (float)r.m_PixelIndex.m_X; o[1] = (float)r.m_PixelIndex.m_Y; o[2] = r.m_Origin.x; o[3] = r.m_Origin.y; o[4] = r.m_Origin.z;
        o[5] = r.m_Direction.x; o[6] = r.m_Direction.y; o[7] = r.m_Direction.z; o[8] = r.m_Contribution.x; o[9] = r.m_Contribution.y; o[10] = r.m_Contribution.z;
    }
    const unsigned total = out->counter;
    free(out); free(s);
    return total;
}

// The ShadeDirect kernel over a w x h grid. lights16: p0, p1, p2, normal, radiance, area in CDF order; the CDF is built by CDF::Insert of
// cdf_weights (ReSTIRData.h:194-203). rays12: px, py, origin3, direction3, tmax, radiance3, channel.
unsigned ref_kat_shade_direct(const float* surfs44, unsigned w, unsigned h, unsigned seed, const float* lights16, const float* cdf_weights, unsigned nlights,
                              float* rays12, unsigned cap)
{
    SurfaceData* s = static_cast<SurfaceData*>(calloc((size_t)w * h, sizeof(SurfaceData)));
    for (unsigned i = 0; i < w * h; ++i) s[i] = make_surface(surfs44 + 44 * (size_t)i, i % w, i / w);
    VolumetricData* vol = static_cast<VolumetricData*>(calloc((size_t)w * h, sizeof(VolumetricData)));
    auto* lights = make_buffer<TriangleLight>(nlights);
    for (unsigned k = 0; k < nlights; ++k) {
        const float* l = lights16 + 16 * k; TriangleLight t;
        t.p0 = make_float3(l[0], l[1], l[2]); t.p1 = make_float3(l[3], l[4], l[5]); t.p2 = make_float3(l[6], l[7], l[8]);
        t.normal = make_float3(l[9], l[10], l[11]); t.radiance = make_float3(l[12], l[13], l[14]); t.area = l[15];
        lights->data[k] = t;
    }
    lights->counter = nlights;
    CDF* cdf = static_cast<CDF*>(calloc(1, sizeof(CDF) + sizeof(float) * (nlights + 1)));
    cdf->Reset();
    for (unsigned k = 0; k < nlights; ++k) cdf->Insert(cdf_weights[k]);
    auto* out = make_buffer<ShadowRayData>(w * h);
    auto* vout = make_buffer<ShadowRayData>(8);
    blockDim = {1, 1, 1}; gridDim = {w, h, 1}; threadIdx = {0, 0, 0};
    for (unsigned y = 0; y < h; ++y) for (unsigned x = 0; x < w; ++x) { blockIdx = {x, y, 0}; ShadeDirect(make_uint3(w, h, 1), s, vol, lights, seed, cdf, out, vout, 0); }
    const unsigned n = out->counter < cap ? out->counter : cap;
    for (unsigned k = 0; k < n; ++k) {
        const ShadowRayData& r = out->data[k]; float* o = rays12 + 12 * k;
        o[0] = (float)r.m_PixelIndex.m_X; o[1] = (float)r.m_PixelIndex.m_Y; o[2] = r.m_Origin.x; o[3] = r.m_Origin.y; o[4] = r.m_Origin.z;
        o[5] = r.m_Direction.x; o[6] = r.m_Direction.y; o[7] = r.m_Direction.z; o[8] = r.m_MaxDistance;
        o[9] = r.m_PotentialRadiance.x; o[10] = r.m_PotentialRadiance.y; o[11] = r.m_PotentialRadiance.z;
    }
    const unsigned total = out->counter;
    free(out); free(vout); free(cdf); free(lights); free(vol); free(s);
    return total;
}

}
