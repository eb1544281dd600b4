// lb_volume.cu — volume kernels of the wavefront: bounds intersection per wave, volumetric shadow rays, delta tracking.
// See lb_volume.cuh for the reference citations and the definition of the two volume modes.
#include "lb_kernels.h"
#include "lb_trace.cuh"
#include "lb_volume.cuh"

namespace lb {

namespace {

constexpr int kBlock = 256;

// K5 + K26: nearest volume entry in front of the surface hit, per queue slot
__global__ void __launch_bounds__(kBlock) k_volume_extend(const float4* __restrict__ ro, const float4* __restrict__ rd, const uint4* __restrict__ hits, const uint32_t* __restrict__ count,
                                                           const DevVolume* __restrict__ vols, uint32_t nvol, float tmin, float tmax, float4* __restrict__ out) {
    const uint32_t n = *count, stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float ht = __uint_as_float(hits[i].w);
        const float far = ht > 0.f ? fminf(tmax, ht) : tmax;
        const VolHit v = volume_intersect(vols, nvol, f3(ro[i]), f3(rd[i]), tmin, far);
        out[i] = make_float4(v.t0, v.t1, v.density, __int_as_float(v.vinst));
    }
}

// volumetric shadow rays of the compat march: several per pixel and launch -> atomic adds (all carry the same constant radiance)
struct VolShadowJob {
    static constexpr bool kDeferDone = false;
    ShadowQueue q; float4* channels; size_t npix; float tmin;
    LB_D bool load(uint32_t i, float3& o, float3& d, float& t0, float& t1) const { const float4 o4 = q.o[i]; o = f3(o4); d = f3(q.d[i]); t0 = tmin; t1 = o4.w; return true; }
    LB_D void done(uint32_t i, bool occluded, const Tracer&) const {
        if (occluded) return;
        const float4 L = q.L[i];
        float* dst = reinterpret_cast<float*>(&channels[(size_t)__float_as_int(L.w) * npix + __float_as_uint(q.d[i].w)]);
        atomicAdd(dst, L.x); atomicAdd(dst + 1, L.y); atomicAdd(dst + 2, L.z);
    }
};
__global__ void __launch_bounds__(kBlock, 4) k_vol_shadow(BvhView bvh, ShadowQueue q, const uint32_t* __restrict__ count, uint32_t* ticket, float4* channels, size_t npix, float tmin, unsigned long long* stat, TraceTuning tune) {
    const uint32_t n = *count;
    VolShadowJob job{q, channels, npix, tmin};
    trace_queue<true>(bvh, n, ticket, job, tune);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stat, (unsigned long long)n);
}

// LB_VOLUME_DELTA: delta tracking through the nearest medium; a real collision replaces the surface interaction of this wave
// by an isotropic scattering event (NEE with ratio-tracked transmittance + continuation ray).
template <bool PRIMARY>
__global__ void __launch_bounds__(kBlock) k_volume_delta(FrameView fv, SceneView sc, int queue, ShadeArgs a) {
    const uint32_t n = fv.counters[queue ? CNT_RAYS_B : CNT_RAYS_A];
    const RayQueue in = fv.rays[queue], out = fv.rays[queue ^ 1];
    uint32_t* out_count = &fv.counters[queue ? CNT_RAYS_A : CNT_RAYS_B];
    uint4* hits = PRIMARY ? fv.primary_hits : fv.hits;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 vh = fv.vol_hits[i];
        const int vinst = __float_as_int(vh.w);
        if (vinst < 0) continue;
        const float4 o4 = in.o[i], d4 = in.d[i];
        const uint32_t pixel = __float_as_uint(d4.w);
        uint32_t seed = wang_hash((a.seed ^ 0x9e3779b9u) + pixel + fv.pix0);
        float ts;
        if (!delta_track(a.volumes[vinst], f3(o4), f3(d4), vh.x, vh.y, seed, ts)) continue;
        hits[i].w = __float_as_uint(-2.f);                       // consumed by the medium: the surface shader sees a miss
        const float3 p = f3(o4) + f3(d4) * ts;
        const float3 T = f3(in.T[i]);
        if (sc.num_lights) {
            uint32_t li; float lpdf; cdf_get(sc, rand_f(seed), li, lpdf);
            const DevLight l = load_light(sc, li);
            const float u = rand_f(seed), v = rand_f(seed) * (1.f - u);
            const float3 point = l.p0 + ((l.p1 - l.p0) * u) + ((l.p2 - l.p0) * v);
            float3 dir = point - p; const float dist = length(dir); dir /= dist;
            const float cos_out = fmaxf(0.f, dot(l.normal, -dir));
            if (dist > 0.01f && cos_out > 0.f) {
                const float solid = (cos_out * l.area) / (dist * dist);
                const float tr = ratio_transmittance(a.volumes, a.num_volumes, p, dir, 0.f, dist - 0.2f, seed);
                const float3 c = T * (kVolumeAlbedo * (0.25f * kInvPi) * solid * (1.f / lpdf) * tr) * l.radiance;
                const uint32_t slot = queue_append_slot(&fv.counters[CNT_SHADOW]);
                fv.shadow.o[slot] = f4(p, dist - 0.2f);
                fv.shadow.d[slot] = f4(dir, __uint_as_float(pixel));
                fv.shadow.L[slot] = f4(c, __int_as_float(a.nee_channel));
            }
        }
        if (a.do_bounce) {
            const float z = 1.f - 2.f * rand_f(seed);
            const float r = sqrtf(fmaxf(0.f, 1.f - z * z));
            const float phi = kTwoPi * rand_f(seed);
            float sp, cp; det_sincos(phi, sp, cp);                     // portable sin / cos: the scattered direction is bit-identical in the oracle
            const float3 dirn = f3(r * cp, r * sp, z);
            const uint32_t slot = queue_append_slot(out_count);
            out.o[slot] = f4(p, 0.f);
            out.d[slot] = f4(dirn, __uint_as_float(pixel));
            out.T[slot] = f4(T * kVolumeAlbedo, 0.f);
        }
    }
}

} // namespace

void launch_volume_extend(const LaunchCfg& cfg, const FrameView& fv, int queue, bool primary, const DevVolume* volumes, uint32_t num_volumes, float tmin, float tmax) {
    k_volume_extend<<<cfg.sms * 8, kBlock, 0, cfg.stream>>>(fv.rays[queue].o, fv.rays[queue].d, primary ? fv.primary_hits : fv.hits, &fv.counters[queue ? CNT_RAYS_B : CNT_RAYS_A],
        volumes, num_volumes, tmin, tmax, fv.vol_hits); LB_LAUNCH_CHECK();
}
void launch_volume_shadow(const LaunchCfg& cfg, const FrameView& fv, const BvhView& bvh, uint32_t ticket, float tmin) {
    k_vol_shadow<<<cfg.sms * 4, kBlock, 0, cfg.stream>>>(bvh, fv.vol_shadow, &fv.counters[CNT_VOL_SHADOW], &fv.counters[CNT_TICKET0 + ticket], fv.channels, fv.npix, tmin, &fv.stats[STAT_SHADOW], cfg.trace_any); LB_LAUNCH_CHECK();
}
void launch_volume_delta(const LaunchCfg& cfg, const FrameView& fv, const SceneView& sc, int queue, bool primary, const ShadeArgs& a) {
    if (primary) k_volume_delta<true><<<cfg.sms * 4, kBlock, 0, cfg.stream>>>(fv, sc, queue, a);
    else k_volume_delta<false><<<cfg.sms * 4, kBlock, 0, cfg.stream>>>(fv, sc, queue, a);
    LB_LAUNCH_CHECK();
}

} // namespace lb
