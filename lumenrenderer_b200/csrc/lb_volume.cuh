// lb_volume.cuh — participating media: volume-bounds intersection, the reference's fixed-step march (compat mode) and
// delta / ratio tracking over a dense density grid (the NanoVDB FloatGrid stand-in).
//
// Reference (under /root/reference/Lumen_Engine/LumenPT/src/):
//   volume bbox intersection + hit record   Shaders/volumetric_wavefront.cu:30-95, CUDAKernels/VolumetricKernels/GPUExtractVolumetricData.cu:7-41
//   5-step constant-density march           CUDAKernels/VolumetricKernels/GPUVolumetricShadeDirect.cu:8-101
// The reference never samples the grid values; delta tracking (Woodcock) with a per-grid majorant and
// ratio-tracked shadow transmittance is this implementation's LB_VOLUME_DELTA mode (north_star item 4). Its oracle twin is
// oracle/oracle.cpp (delta_track / ratio_transmittance) with the same RNG streams.
#pragma once
#include "lb_shade.cuh"
#include "lb_kernels.h"

namespace lb {

constexpr float kVolumeAlbedo = 0.8f;          // single-scattering albedo, isotropic phase
constexpr int kTrackingMaxSteps = 1024;

struct VolHit { float t0, t1, density; int vinst; };

LB_D VolHit vol_hit_none() { VolHit v; v.t0 = -1.f; v.t1 = -1.f; v.density = 0.f; v.vinst = -1; return v; }

// slab test of the ray against every volume instance's object-space box; nearest entry wins
LB_D VolHit volume_intersect(const DevVolume* __restrict__ vols, uint32_t nvol, const float3& ro, const float3& rd, float tmin, float tmax) {
    VolHit best = vol_hit_none();
    for (uint32_t k = 0; k < nvol; ++k) {
        const DevVolume& v = vols[k];
        const float3 o = xform_point(v.inv, ro), d = xform_vector(v.inv, rd);
        float t0 = tmin, t1 = tmax; bool ok = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!ok) break;
            const float inv = 1.0f / comp(d, a);
            float ta = (comp(v.lo, a) - comp(o, a)) * inv, tb = (comp(v.hi, a) - comp(o, a)) * inv;
            if (ta > tb) { const float s = ta; ta = tb; tb = s; }
            t0 = fmaxf(t0, ta); t1 = fminf(t1, tb); ok = t0 <= t1;
        }
        if (ok && (best.vinst < 0 || t0 < best.t0)) { best.t0 = t0; best.t1 = t1; best.density = v.instance_density; best.vinst = (int)k; }
    }
    return best;
}

// nearest-voxel density lookup in object space
LB_D float volume_density(const DevVolume& v, const float3& p) {
    if (!v.density) return 1.f;
    const float3 e = v.hi - v.lo;
    int x = (int)((p.x - v.lo.x) / e.x * (float)v.nx), y = (int)((p.y - v.lo.y) / e.y * (float)v.ny), z = (int)((p.z - v.lo.z) / e.z * (float)v.nz);
    x = x < 0 ? 0 : (x >= (int)v.nx ? (int)v.nx - 1 : x); y = y < 0 ? 0 : (y >= (int)v.ny ? (int)v.ny - 1 : y); z = z < 0 ? 0 : (z >= (int)v.nz ? (int)v.nz - 1 : z);
    return __ldg(&v.density[((size_t)z * v.ny + y) * v.nx + x]);
}

// delta tracking: true when a real collision happens before t1 (t_scatter in world ray units)
LB_D bool delta_track(const DevVolume& v, const float3& ro, const float3& rd, float t0, float t1, uint32_t& seed, float& t_scatter) {
    const float sigma_max = v.instance_density * v.majorant;
    if (!(sigma_max > 0.f)) return false;
    const float3 o = xform_point(v.inv, ro), d = xform_vector(v.inv, rd);
    float t = t0;
    for (int it = 0; it < kTrackingMaxSteps; ++it) {
        t -= logf(1.0f - rand_f(seed) * 0.99999994f) / sigma_max;
        if (t >= t1) return false;
        const float dens = v.instance_density * volume_density(v, o + d * t);
        if (rand_f(seed) * sigma_max < dens) { t_scatter = t; return true; }
    }
    return false;
}

// ratio-tracking estimate of the transmittance of the segment [tmin, tmax] through every volume instance
LB_D float ratio_transmittance(const DevVolume* __restrict__ vols, uint32_t nvol, const float3& ro, const float3& rd, float tmin, float tmax, uint32_t& seed) {
    float tr = 1.f;
    for (uint32_t k = 0; k < nvol; ++k) {
        const DevVolume& v = vols[k];
        const float3 o = xform_point(v.inv, ro), d = xform_vector(v.inv, rd);
        float t0 = tmin, t1 = tmax; bool ok = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!ok) break;
            const float inv = 1.0f / comp(d, a);
            float ta = (comp(v.lo, a) - comp(o, a)) * inv, tb = (comp(v.hi, a) - comp(o, a)) * inv;
            if (ta > tb) { const float s = ta; ta = tb; tb = s; }
            t0 = fmaxf(t0, ta); t1 = fminf(t1, tb); ok = t0 <= t1;
        }
        if (!ok) continue;
        const float sigma_max = v.instance_density * v.majorant;
        if (!(sigma_max > 0.f)) continue;
        if (!v.density) { tr *= expf(-sigma_max * (t1 - t0)); continue; }       // homogeneous: analytic
        float t = t0;
        for (int it = 0; it < kTrackingMaxSteps; ++it) {
            t -= logf(1.0f - rand_f(seed) * 0.99999994f) / sigma_max;
            if (t >= t1) break;
            const float dens = v.instance_density * volume_density(v, o + d * t);
            tr *= 1.0f - dens / sigma_max;
            if (tr <= 0.f) return 0.f;
        }
    }
    return tr;
}

} // namespace lb
