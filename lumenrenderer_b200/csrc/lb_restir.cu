// lb_restir.cu — ReSTIR direct lighting: light bags, RIS candidate generation, visibility, temporal and spatial reuse.
//
// Reference (under /root/reference/Lumen_Engine/LumenPT/src/):
//   pass order            Framework/ReSTIR.cpp:65-233
//   FillLightBags         CUDAKernels/ReSTIRKernels.cu:343-370
//   PickPrimarySamples    CUDAKernels/ReSTIRKernels.cu:402-522   (Reservoir::Update/UpdateWeight: Shaders/CppCommon/ReSTIRData.h:122-161)
//   GenerateShadowRay + ReSTIRRayGen + ShadeReservoirs   ReSTIRKernels.cu:546-665, Shaders/WaveFrontShaders.cu:181-216
//   temporal              ReSTIRKernels.cu:1015-1121      spatial ReSTIRKernels.cu:787-980
//   CombineBiased / CombineReservoirBuffers   ReSTIRKernels.cu:1200-1257, :1407-1436
// Canonical choices for the reference's non-deterministic spots (SURVEY hazards 1, 13, 14): the light bag is chosen per
// 256-pixel block from WangHash(seed + block) instead of the hardware SM id; Reservoir::Update receives its seed by value.
// B200 design: reservoirs and surfaces are SoA 16-byte planes; the visibility pass generates, traces and shades in ONE
// kernel (no 32-B ray round trip through HBM, no host read-back of a ray counter).
#include "lb_kernels.h"
#include "lb_trace.cuh"
#include "lb_shade.cuh"
#include <cfloat>

namespace lb {

namespace {

constexpr int kBlock = 256;
constexpr uint32_t kNumBags = 50, kLightsPerBag = 1000, kPrimarySamples = 32, kSpatialSamples = 5, kSpatialRadius = 30, kSpatialIterations = 2;   // ReSTIRData.h:34-56
constexpr float kSimilarCos = 0.72222222223f;

LB_D bool reservoir_update(Reservoir& r, const LightSample& s, float w, uint32_t seed /* by value */) {
    r.weight_sum += w; ++r.count;
    const float u = rand_f(seed);
    if (u <= (w / r.weight_sum)) { r.s = s; return true; }
    return false;
}
LB_D void reservoir_update_weight(Reservoir& r) {
    if (r.count == 0 || r.weight_sum <= 0.f) { r.weight = 0.f; return; }
    r.weight = (1.f / fmaxf(r.s.pdf, FLT_EPSILON)) * ((1.f / (float)r.count) * r.weight_sum);
}
LB_D bool similar(float d1, float d2, const float3& n1, const float3& n2) {
    const float pct = fabsf(d1 - d2) / ((d1 + d2) / 2.f);
    return pct < 0.10f && dot(n1, n2) > kSimilarCos;
}
// CombineBiased over two reservoirs
LB_D Reservoir combine_pair(const Reservoir& a, const Reservoir& b, const Surface& px, uint32_t seed) {
    Reservoir out = reservoir_zero(); int total = 0;
    {
        LightSample rs; resample(a.s, px, rs);
        reservoir_update(out, rs, (float)a.count * a.weight * rs.pdf, seed); total += a.count;
    }
    {
        LightSample rs; resample(b.s, px, rs);
        reservoir_update(out, rs, (float)b.count * b.weight * rs.pdf, seed); total += b.count;
    }
    out.count = total; reservoir_update_weight(out);
    return out;
}

__global__ void __launch_bounds__(kBlock) k_fill_bags(SceneView sc, uint2* __restrict__ bags, uint32_t a_seed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kNumBags * kLightsPerBag) return;
    uint32_t s = wang_hash(a_seed + wang_hash(i));
    const float r = rand_f(s);
    uint32_t li; float pdf; cdf_get(sc, r, li, pdf);
    bags[i] = make_uint2(li, __float_as_uint(pdf));
}

__global__ void __launch_bounds__(kBlock) k_ris(FrameView fv, SceneView sc, const uint2* __restrict__ bags, uint32_t seed) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const size_t np = fv.npix;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < fv.npix; i += stride) {
        if (surface_flags(fv.surf_cur, np, i)) { reservoir_store(fv.res_cur, np, i, reservoir_zero()); continue; }
        uint32_t bag_seed = wang_hash(seed + i / 256u);
        const int bag = (int)roundf((float)(kNumBags - 1u) * rand_f(bag_seed));
        const uint2* picked = bags + (size_t)bag * kLightsPerBag;
        Surface px; surface_load(fv.surf_cur, np, i, px);
        uint32_t s = wang_hash(seed + wang_hash(i));
        Reservoir fresh = reservoir_zero();
        for (uint32_t k = 0; k < kPrimarySamples; ++k) {
            const float r = rand_f(s);
            const uint2 be = __ldg(&picked[(int)roundf((float)(kLightsPerBag - 1u) * r)]);
            const DevLight l = load_light(sc, be.x);
            const float u = rand_f(s), v = rand_f(s) * (1.f - u);
            LightSample ls; ls.radiance = l.radiance; ls.normal = l.normal; ls.area = l.area; ls.contribution = f3(0.f); ls.pdf = 0.f;
            ls.position = l.p0 + ((l.p1 - l.p0) * u) + ((l.p2 - l.p0) * v);
            LightSample rs; resample(ls, px, rs);
            reservoir_update(fresh, rs, rs.pdf / __uint_as_float(be.y), s);
        }
        reservoir_update_weight(fresh);
        reservoir_store(fv.res_cur, np, i, fresh);
    }
}

// visibility of the reservoir's sample + shading of the survivor into DIRECT, one kernel
__global__ void __launch_bounds__(kBlock) k_visibility_shade(FrameView fv, BvhView bvh, uint32_t* ticket, float inv_shaded_count_denominator, unsigned long long* stat) {
    const uint32_t n = fv.npix;
    const size_t np = fv.npix;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t traced = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(ticket, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + lane;
        if (i < n) {
            float4 r0 = fv.res_cur[i];                         // weightSum, weight, count, pdf
            float weight = r0.y;
            const float4 sp = fv.surf_cur[i];
            const uint32_t flags = surface_flags(fv.surf_cur, np, i);
            if (!flags && weight > 0.f) {
                const float3 pos = f3(sp);
                float3 d = f3(fv.res_cur[np + i]) - pos; const float l = length(d); d /= l;
                ++traced;
                HitInfo h;
                if (bvh8_trace<true>(bvh, pos, d, 0.1f, l - 0.05f, h)) { weight = 0.f; r0.y = 0.f; fv.res_cur[i] = r0; }
            }
            if (weight > 0.f) {
                const float3 c = f3(fv.res_cur[4 * np + i]) * (weight / inv_shaded_count_denominator);
                float4 o = fv.channels[i]; o.x += c.x; o.y += c.y; o.z += c.z; fv.channels[i] = o;
            }
        }
        __syncwarp();
    }
    traced = __reduce_add_sync(0xFFFFFFFFu, traced);
    if (lane == 0 && traced) atomicAdd(stat, (unsigned long long)traced);
}

__global__ void __launch_bounds__(kBlock) k_temporal(FrameView fv, uint32_t seed, float shaded) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const size_t np = fv.npix;
    const int W = (int)fv.width, H = (int)fv.height;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < fv.npix; i += stride) {
        const int cy = (int)(i / fv.width), cx = (int)(i - (uint32_t)cy * fv.width);
        const float2 mvec = fv.motion[i];
        const int mx = (int)roundf((float)W * mvec.x), my = (int)roundf((float)H * mvec.y);
        const int ty = cy + my, tx = cx + mx; uint32_t ti = i;
        if (ty >= 0 && ty < H && tx >= 0 && tx < W) ti = (uint32_t)ty * fv.width + (uint32_t)tx;
        Surface sp; surface_load_geom(fv.surf_prev, np, ti, sp);
        if (sp.flags) continue;
        if (surface_flags(fv.surf_cur, np, i)) continue;
        Surface sc; surface_load(fv.surf_cur, np, i, sc);
        if (!similar(sp.t, sc.t, sp.normal, sc.normal)) continue;
        Reservoir prev, cur; reservoir_load(fv.res_prev, np, ti, prev); reservoir_load(fv.res_cur, np, i, cur);
        if (prev.weight > 0.f) {
            const float3 c = prev.s.contribution * (prev.weight / shaded);
            float4 o = fv.channels[i]; o.x += c.x; o.y += c.y; o.z += c.z; fv.channels[i] = o;
        }
        prev.count = min(prev.count, cur.count * 20);
        reservoir_store(fv.res_cur, np, i, combine_pair(prev, cur, sc, wang_hash(seed + i)));
    }
}

__global__ void __launch_bounds__(kBlock) k_spatial(FrameView fv, const float4* __restrict__ in, float4* __restrict__ out, uint32_t seed) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const size_t np = fv.npix;
    const int W = (int)fv.width, H = (int)fv.height;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < fv.npix; i += stride) {
        Surface sc; surface_load_geom(fv.surf_cur, np, i, sc);
        if (sc.flags) continue;
        uint32_t s = wang_hash(seed + i);
        const int y = (int)(i / fv.width), x = (int)(i - (uint32_t)y * fv.width);
        uint32_t nb[kSpatialSamples]; int count = 0;
#pragma unroll
        for (uint32_t k = 0; k < kSpatialSamples; ++k) {
            const int ny = (int)roundf((rand_f(s) * 2.f - 1.f) * (float)kSpatialRadius) + y;
            const int nx = (int)roundf((rand_f(s) * 2.f - 1.f) * (float)kSpatialRadius) + x;
            if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
            const uint32_t ni = (uint32_t)ny * fv.width + (uint32_t)nx;
            Surface sn; surface_load_geom(fv.surf_cur, np, ni, sn);
            if (sn.flags) continue;
            if (similar(sn.t, sc.t, sn.normal, sc.normal)) nb[count++] = ni;
        }
        if (count > 1) {
            Surface p0; surface_load(fv.surf_cur, np, nb[0], p0);      // resampled at the FIRST accepted neighbour
            Reservoir acc = reservoir_zero(); int total = 0;
            for (int k = 0; k < count; ++k) {
                Reservoir q; reservoir_load(in, np, nb[k], q);
                LightSample rs; resample(q.s, p0, rs);
                reservoir_update(acc, rs, (float)q.count * q.weight * rs.pdf, seed);   // kernel-wide seed by value (hazard 14)
                total += q.count;
            }
            acc.count = total; reservoir_update_weight(acc);
            reservoir_store(out, np, i, acc);
        } else {
            const float4 r0 = out[i];                                   // Reservoir::Reset keeps the stored sample
            out[i] = make_float4(0.f, 0.f, __int_as_float(0), r0.w);
        }
    }
}

__global__ void __launch_bounds__(kBlock) k_combine(FrameView fv, const float4* __restrict__ nbuf, uint32_t cseed) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const size_t np = fv.npix;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < fv.npix; i += stride) {
        if (surface_flags(fv.surf_cur, np, i)) continue;
        Surface sc; surface_load(fv.surf_cur, np, i, sc);
        Reservoir a, b; reservoir_load(fv.res_cur, np, i, a); reservoir_load(nbuf, np, i, b);
        reservoir_store(fv.res_cur, np, i, combine_pair(a, b, sc, wang_hash(cseed + i)));
    }
}

} // namespace

void launch_restir(const LaunchCfg& cfg, const FrameView& fv, const SceneView& sc, const BvhView& bvh, const RestirBuffers& rb, const RestirArgs& a, uint32_t& ticket) {
    if (sc.num_lights == 0u) return;
    cudaStream_t st = cfg.stream;
    const int grid = cfg.sms * 4;
    uint32_t seed = wang_hash(a.seed);
    k_fill_bags<<<grid_for(kNumBags * kLightsPerBag, kBlock), kBlock, 0, st>>>(sc, rb.bags, a.seed); LB_LAUNCH_CHECK();
    seed = wang_hash(seed);
    k_ris<<<grid, kBlock, 0, st>>>(fv, sc, rb.bags, seed); LB_LAUNCH_CHECK();
    const float shaded = 1.f + (a.temporal ? 1.f : 0.f) + (a.spatial ? 1.f : 0.f);
    k_visibility_shade<<<grid, kBlock, 0, st>>>(fv, bvh, &fv.counters[CNT_TICKET0 + ticket++], shaded, &fv.stats[STAT_VIS]); LB_LAUNCH_CHECK();
    if (a.temporal) {
        seed = wang_hash(seed);
        k_temporal<<<grid, kBlock, 0, st>>>(fv, seed, shaded); LB_LAUNCH_CHECK();
    }
    if (a.spatial) {
        seed = wang_hash(seed);
        const float4* from = fv.res_cur; float4* to = fv.res_tmp_a;
        for (uint32_t it = 0; it < kSpatialIterations; ++it) {
            k_spatial<<<grid, kBlock, 0, st>>>(fv, from, to, seed); LB_LAUNCH_CHECK();
            if (it == 0) { from = fv.res_tmp_a; to = fv.res_tmp_b; } else { const float4* t = from; from = to; to = const_cast<float4*>(t); }
        }
        k_visibility_shade<<<grid, kBlock, 0, st>>>(fv, bvh, &fv.counters[CNT_TICKET0 + ticket++], shaded, &fv.stats[STAT_VIS]); LB_LAUNCH_CHECK();
        k_combine<<<grid, kBlock, 0, st>>>(fv, from, wang_hash(seed)); LB_LAUNCH_CHECK();
    }
}

} // namespace lb
