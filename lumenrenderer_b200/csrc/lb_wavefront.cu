// lb_wavefront.cu — the wavefront kernels: ray generation, extend, fused extract+NEE+bounce, shadow, merge.
//
// Reference path (under /root/reference/Lumen_Engine/LumenPT/src/):
//   K1  GeneratePrimaryRay        CUDAKernels/WaveFrontKernels/GPUGeneratePrimRay.cu:8-82
//   K2  extend (optixTrace)       Shaders/WaveFrontShaders.cu:42-112, closest-hit pack :301-328
//   K3  shadow rays + accumulate  Shaders/WaveFrontShaders.cu:114-179
//   K6  ExtractSurfaceData        CUDAKernels/WaveFrontKernels/GPUExtractSurfaceData.cu:8-228
//   K8  motion vectors            CUDAKernels/MotionVectors.cu:8-55
//   K9  ResolveDirectLightHits    CUDAKernels/WaveFrontKernels/GPUShadeDirect.cu:11-40
//   K10 ShadeDirect, K11 ShadeIndirect   GPUShadeDirect.cu:42-153, GPUShadeIndirect.cu:7-146
//   K12 MergeOutputChannels, K13 WriteToOutput   GPUMergeOutputChannels.cu:5-88, GPUShadingKernels.cu:28-56
// B200 design: SoA 16-byte planes, one fused shade kernel per wave (the 176-B SurfaceData only reaches HBM for the
// primary hit, where ReSTIR needs it), warp-aggregated queue appends, persistent grids reading device-side counts.
#include "lb_kernels.h"
#include "lb_trace.cuh"
#include "lb_shade.cuh"
#include "lb_volume.cuh"

namespace lb {

namespace {

constexpr int kBlock = 256;
#ifndef LB_SHADE_BLOCKS
#define LB_SHADE_BLOCKS 2         // resident blocks per SM the fused shade kernel is compiled for
#endif

// ------------------------------------------------------------------ K1 ray generation
__device__ __forceinline__ float halton(uint32_t index, uint32_t base) {
    ++index; float f = 1.f, r = 0.f;
    // IEEE division; halving is exact as a product (base is a compile-time 2 or 3 at both call sites)
    while (index > 0) { f = base == 2u ? f * 0.5f : xdiv(f, (float)base); r = r + f * (float)(index % base); index = index / base; }
    return r;
}

__global__ void __launch_bounds__(kBlock) k_raygen(FrameView fv, CameraBasis cam, uint32_t frame_count) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < fv.npix; i += stride) {
        const uint32_t sy = i / fv.width, sx = i - sy * fv.width;
        const uint32_t gi = i + fv.pix0;                    // the pixel's index in the full frame (== i unless this renderer is a row band)
        const float jx = halton(frame_count + gi, 2), jy = halton(frame_count + gi, 3);
        float dx = xdiv((float)sx + jx, (float)fv.width), dy = xdiv((float)(sy + fv.row0) + jy, (float)fv.full_height);
        dx = -(dx * 2.0f - 1.0f); dy = -(dy * 2.0f - 1.0f);
        const float3 d = f3(fmaf(dx, cam.U.x, fmaf(dy, cam.V.x, cam.W.x)), fmaf(dx, cam.U.y, fmaf(dy, cam.V.y, cam.W.y)), fmaf(dx, cam.U.z, fmaf(dy, cam.V.z, cam.W.z)));
        const float len2 = fmaf(d.x, d.x, fmaf(d.y, d.y, d.z * d.z));
        const float inv = xdiv(1.0f, xsqrt(len2));
        fv.rays[0].o[i] = f4(cam.eye, 0.f);
        fv.rays[0].d[i] = make_float4(d.x * inv, d.y * inv, d.z * inv, __uint_as_float(i));
        fv.rays[0].T[i] = make_float4(1.f, 1.f, 1.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { fv.counters[CNT_RAYS_A] = fv.npix; fv.counters[CNT_RAYS_B] = 0u; }
}

// ------------------------------------------------------------------ K2 extend: persistent warps, per-lane ray refill (trace_queue)
struct ExtendJob {
    static constexpr bool kDeferDone = false;
    const float4* __restrict__ ro; const float4* __restrict__ rd; uint4* __restrict__ hits; float tmin, tmax;
    LB_D bool load(uint32_t i, float3& o, float3& d, float& t0, float& t1) const { o = f3(ro[i]); d = f3(rd[i]); t0 = tmin; t1 = tmax; return true; }
    LB_D void done(uint32_t i, bool hit, const Tracer& tr) const {
        uint4 rec = make_uint4(0u, 0u, 0u, __float_as_uint(-1.f));
        if (hit) {
            const __half2 b = __floats2half2_rn(tr.bu, tr.bv);            // fp16 barycentrics, WaveFrontShaders.cu:318-321
            rec = make_uint4(tr.bi, tr.bp, *reinterpret_cast<const uint32_t*>(&b), __float_as_uint(tr.best));
        }
        hits[i] = rec;
    }
};
__global__ void __launch_bounds__(kBlock, 4) k_extend(BvhView bvh, const float4* __restrict__ ro, const float4* __restrict__ rd, const uint32_t* __restrict__ count,
                                                    uint32_t* ticket, uint4* __restrict__ hits, float tmin, float tmax, unsigned long long* stat, TraceTuning tune) {
    const uint32_t n = *count;
    ExtendJob job{ro, rd, hits, tmin, tmax};
    trace_queue<false>(bvh, n, ticket, job, tune);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stat, (unsigned long long)n);
}

// ------------------------------------------------------------------ fused shade: K6 (+K8, K9 at depth 0) + K10 + K11
// NEE = false: the instance for waves that draw no light sample (the primary wave when ReSTIR supplies the direct light) — the NEE and
// compat-volume code is not generated at all: a smaller kernel for the launch that touches every pixel
template <bool PRIMARY, bool NEE = true>
__global__ void __launch_bounds__(kBlock, LB_SHADE_BLOCKS) k_shade(FrameView fv, SceneView sc, int queue, ShadeArgs a) {
    const uint32_t n = fv.counters[queue ? CNT_RAYS_B : CNT_RAYS_A];
    const RayQueue in = fv.rays[queue], out = fv.rays[queue ^ 1];
    uint32_t* out_count = &fv.counters[queue ? CNT_RAYS_A : CNT_RAYS_B];
    const uint4* hits = PRIMARY ? fv.primary_hits : fv.hits;
    const uint32_t stride = gridDim.x * blockDim.x;
    const size_t np = fv.npix;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 o4 = in.o[i], d4 = in.d[i], T4 = in.T[i];
        const uint4 hr = hits[i];
        const uint32_t pixel = __float_as_uint(d4.w);
        const uint32_t gpixel = pixel + fv.pix0;            // full-frame index: what every random stream is keyed on
        const __half2 hb = *reinterpret_cast<const __half2*>(&hr.z);
        const Surface s = extract_surface(sc, f3(o4), f3(d4), f3(T4), hr.x, hr.y, __low2float(hb), __high2float(hb), __uint_as_float(hr.w));

        if (PRIMARY) {
            surface_store(fv.surf_cur, np, pixel, s);
            // K8 motion vector (fp16-rounded like the reference's half2 surface)
            float2 mv = make_float2(0.f, 0.f);
            if (s.t > 0.f) {
                const uint32_t y = pixel / fv.width, x = pixel - y * fv.width;
                const float cx = xdiv((float)x + 0.5f, (float)fv.width), cy = xdiv((float)(y + fv.row0) + 0.5f, (float)fv.full_height);
                const float* M = a.prev_view_proj;
                const float px = fmaf(M[0], s.pos.x, fmaf(M[1], s.pos.y, fmaf(M[2], s.pos.z, M[3])));
                const float py = fmaf(M[4], s.pos.x, fmaf(M[5], s.pos.y, fmaf(M[6], s.pos.z, M[7])));
                const float pw = fmaf(M[12], s.pos.x, fmaf(M[13], s.pos.y, fmaf(M[14], s.pos.z, M[15])));
                const float inv = xdiv(1.0f, pw);
                mv = make_float2(half_round((px * inv * 0.5f + 0.5f) - cx), half_round((py * inv * 0.5f + 0.5f) - cy));
            }
            fv.motion[pixel] = mv;
            // the four light channels start the frame here: K9 writes emissive primary hits into DIRECT
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            fv.channels[0 * np + pixel] = (s.flags & SURF_EMISSIVE) ? s.mat.color : zero;
            fv.channels[1 * np + pixel] = zero; fv.channels[2 * np + pixel] = zero; fv.channels[3 * np + pixel] = zero;
        }

        const bool owned = pixel >= fv.own_pix0 && pixel < fv.own_pix1;          // halo rows of a band: surface record only
        if (NEE && a.do_nee && owned) {
            uint32_t seed = wang_hash(a.seed + gpixel);
            if (a.num_volumes && a.volume_mode == 0 /* LB_VOLUME_COMPAT */ && sc.num_lights) {
                // VolumetricShadeDirect (GPUVolumetricShadeDirect.cu:8-101): 5 fixed steps, constant density per unit length, the grid is
                // never sampled; every step spawns a shadow ray of constant radiance 0.01; accumulated density becomes the alpha
                const float4 vh = fv.vol_hits[i];
                if (__float_as_int(vh.w) >= 0 && vh.y > vh.x) {
                    const float3 rd3 = f3(d4), entry = f3(o4) + rd3 * vh.x;
                    const float distance = vh.y - vh.x; float acc = 0.f; const float step = distance / 5;
                    float3 prev = entry; const float offset = rand_f(seed) * step;
                    for (int k = 0; k < 5 && acc < 1.0f && (float)k * step < distance; k++) {
                        const float ts = (float)k * step + offset; const float3 p = entry + rd3 * ts;
                        const float dprev = length(p - prev); prev = p;
                        uint32_t li; float lpdf; cdf_get(sc, rand_f(seed), li, lpdf);
                        const DevLight l = load_light(sc, li);
                        const float u = rand_f(seed), v = rand_f(seed) * (1.f - u);
                        const float3 point = l.p0 + ((l.p1 - l.p0) * u) + ((l.p2 - l.p0) * v);
                        float3 dir = point - p; const float ld = length(dir); dir /= ld;
                        const uint32_t slot = queue_append_slot(&fv.counters[CNT_VOL_SHADOW]);
                        fv.vol_shadow.o[slot] = f4(p, ld - 0.2f);
                        fv.vol_shadow.d[slot] = f4(dir, __uint_as_float(pixel));
                        fv.vol_shadow.L[slot] = make_float4(0.01f, 0.01f, 0.01f, __int_as_float(3));
                        acc += vh.z * dprev;
                    }
                    fv.channels[3 * np + pixel].w = acc;          // rgb of the channel is only ever touched by the shadow adds
                }
            }
            ShadowRayOut sr;
            bool ok = nee_sample(sc, s, seed, sr);
            if (ok && a.num_volumes && a.volume_mode == 1 /* LB_VOLUME_DELTA */) {
                uint32_t vseed = wang_hash((a.seed ^ 0x85ebca6bu) + gpixel);
                const float tr = ratio_transmittance(a.volumes, a.num_volumes, sr.o, sr.d, 0.01f, sr.tmax, vseed);
                sr.radiance *= tr;
            }
            if (ok) {
                const uint32_t slot = queue_append_slot(&fv.counters[CNT_SHADOW]);
                fv.shadow.o[slot] = f4(sr.o, sr.tmax);
                fv.shadow.d[slot] = f4(sr.d, __uint_as_float(pixel));
                fv.shadow.L[slot] = f4(sr.radiance, __int_as_float(a.nee_channel));
            }
        }
        if (a.do_bounce && owned) {
            BounceOut b;
            const bool ok = bounce_sample(s, gpixel, wang_hash(a.seed), b);
            if (ok) {
                const uint32_t slot = queue_append_slot(out_count);
                out.o[slot] = f4(b.o, 0.f);
                out.d[slot] = f4(b.d, __uint_as_float(pixel));
                out.T[slot] = f4(b.throughput, 0.f);
            }
        }
    }
}

// ------------------------------------------------------------------ K3 shadow rays: any-hit, unoccluded rays add their radiance
// One shadow ray per pixel per launch (one NEE sample per wave), so the fp32 read-modify-write below is race free —
// the reference's fp16 RMW is racy (SURVEY hazard 2).
struct ShadowJob {
    static constexpr bool kDeferDone = true;
    ShadowQueue q; float4* channels; size_t npix; float tmin;
    LB_D bool load(uint32_t i, float3& o, float3& d, float& t0, float& t1) const { const float4 o4 = q.o[i]; o = f3(o4); d = f3(q.d[i]); t0 = tmin; t1 = o4.w; return true; }
    LB_D void done(uint32_t i, bool occluded, const Tracer&) const {
        if (occluded) return;
        const float4 L = q.L[i];
        float4* dst = &channels[(size_t)__float_as_int(L.w) * npix + __float_as_uint(q.d[i].w)];
        float4 c = *dst; c.x += L.x; c.y += L.y; c.z += L.z; *dst = c;
    }
};
__global__ void __launch_bounds__(kBlock, 4) k_shadow(BvhView bvh, ShadowQueue q, const uint32_t* __restrict__ count, uint32_t* ticket,
                                                    float4* channels, size_t npix, float tmin, unsigned long long* stat, TraceTuning tune) {
    const uint32_t n = *count;
    ShadowJob job{q, channels, npix, tmin};
    trace_queue<true>(bvh, n, ticket, job, tune);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stat, (unsigned long long)n);
}

// ------------------------------------------------------------------ the late bounce waves as ONE launch (K2 + K6 + K10 + K3 + K11 per path)
// From the third wave on a frame holds few rays (C2: 1e5, then 2e4 of 3.7 M pixels): a wave's extend, shade and shadow launches each last as
// long as their slowest ray's chain of dependent node visits (~0.03 .. 0.1 ms), whatever their size, and every launch ends in a tail of idle
// SMs. Here a LANE follows one path from its queue entry to its end — closest hit, surface extraction, NEE sample + its shadow ray, bounce,
// next closest hit, ... — so the rest of the frame's bounce chain is one launch with no grid-wide step between the waves. Every stage is the
// device function the per-wave kernels call, seeded per wave exactly as the host seeds the launches (seed_d+1 = WangHash(seed_d)), and the
// fp32 adds into the pixel's INDIRECT channel happen in wave order: the image is bit-identical to the per-wave schedule.
struct TailArgs { uint32_t first_depth, max_depth, seed; float tmin, tmax; };
__global__ void __launch_bounds__(128, 4) k_tail(FrameView fv, SceneView sc, BvhView bvh, BvhView bvh_any, int queue, TailArgs a) {
    const uint32_t n = fv.counters[queue ? CNT_RAYS_B : CNT_RAYS_A];
    const RayQueue in = fv.rays[queue];
    const size_t np = fv.npix;
    uint32_t n_extend = 0u, n_shadow = 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = in.o[i], d4 = in.d[i], T4 = in.T[i];
        float3 o = f3(o4), d = f3(d4), T = f3(T4);
        const uint32_t pixel = __float_as_uint(d4.w), gpixel = pixel + fv.pix0;
        uint32_t seed = a.seed;
        for (uint32_t depth = a.first_depth; depth < a.max_depth; ++depth, seed = wang_hash(seed)) {
            HitInfo h; h.inst = 0u; h.prim = 0u; h.u = 0.f; h.v = 0.f; h.t = -1.f;
            const bool hit = bvh8_trace<false>(bvh, o, d, a.tmin, a.tmax, h);
            ++n_extend;
            // the hit record of the wavefront path carries fp16 barycentrics (WaveFrontShaders.cu:318-321)
            const Surface s = extract_surface(sc, o, d, T, h.inst, h.prim, half_round(h.u), half_round(h.v), hit ? h.t : -1.f);
            uint32_t nee_seed = wang_hash(seed + gpixel);
            ShadowRayOut sr;
            if (nee_sample(sc, s, nee_seed, sr)) {
                ++n_shadow;
                HitInfo unused;
                if (!bvh8_trace<true>(bvh_any, sr.o, sr.d, a.tmin, sr.tmax, unused)) {
                    float4* dst = &fv.channels[(size_t)1 /* LB_CHANNEL_INDIRECT */ * np + pixel];
                    float4 c = *dst; c.x += sr.radiance.x; c.y += sr.radiance.y; c.z += sr.radiance.z; *dst = c;
                }
            }
            if (depth + 1u >= a.max_depth) break;
            BounceOut b;
            if (!bounce_sample(s, gpixel, wang_hash(seed), b)) break;
            o = b.o; d = b.d; T = b.throughput;
        }
    }
    n_extend = __reduce_add_sync(0xFFFFFFFFu, n_extend); n_shadow = __reduce_add_sync(0xFFFFFFFFu, n_shadow);
    if ((threadIdx.x & 31u) == 0u) {
        if (n_extend) atomicAdd(&fv.stats[STAT_EXTEND], (unsigned long long)n_extend);
        if (n_shadow) atomicAdd(&fv.stats[STAT_SHADOW], (unsigned long long)n_shadow);
    }
}

// ------------------------------------------------------------------ K12 merge + progressive accumulate + K13 8-bit output
__device__ __forceinline__ unsigned char to_srgb8(float c) {
    c = clampf(c, 0.f, 1.f);
    const float s = c < 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
    const float x = clampf(s, 0.f, 1.f);
    const uint32_t v = (uint32_t)(x * 256.f);
    return (unsigned char)(v > 255u ? 255u : v);
}

__global__ void __launch_bounds__(kBlock) k_merge(FrameView fv, int blend, uint32_t blend_count) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const size_t np = fv.npix;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < fv.npix; i += stride) {
        const float4 d = fv.channels[i], in = fv.channels[np + i], sp = fv.channels[2 * np + i], vo = fv.channels[3 * np + i];
        float4 m = make_float4((d.x + in.x) + sp.x, (d.y + in.y) + sp.y, (d.z + in.z) + sp.z, (d.w + in.w) + sp.w);
        const float al = vo.w;
        m = make_float4(m.x * (1.0f - al) + vo.x * al, m.y * (1.0f - al) + vo.y * al, m.z * (1.0f - al) + vo.z * al, m.w * (1.0f - al) + vo.w * al);
        float4 c;
        if (blend) {
            float4 acc = blend_count ? fv.accum[i] + m : m;          // the first blended frame starts the sum (no cleared buffer needed)
            fv.accum[i] = acc;
            const float inv = xdiv(1.0f, (float)(blend_count + 1u));
            c = acc * inv;
        } else { c = m; fv.accum[i] = m; }
        fv.combined[i] = c;
        fv.ldr[i] = make_uchar4(to_srgb8(c.x), to_srgb8(c.y), to_srgb8(c.z), 255);
    }
}

__global__ void __launch_bounds__(kBlock) k_resolve(FrameView fv, float inv) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < fv.npix; i += stride) {
        const float4 c = fv.accum[i] * inv;
        fv.combined[i] = c;
        fv.ldr[i] = make_uchar4(to_srgb8(c.x), to_srgb8(c.y), to_srgb8(c.z), 255);
    }
}

// ------------------------------------------------------------------ G-buffer side outputs for a denoiser / upscaler behind the path
// depth: ExtractDepthDataGpu / ExtractNRD_DLSSdataGpu (GPUExtractDepthData.cu:6-72, GPUExtractNRD_DLSSdata.cu:6-89): t normalised by the
// camera's min/max render distance, 0 where t < 0. normal + roughness: the reference writes a half4 surface -> values rounded through
// fp16. albedo: m_MaterialData.m_Color (PrepareOptixDenoisingGPU, GPUPostProcessingEffects.cu:13-50).
__global__ void __launch_bounds__(kBlock) k_gbuffer(const float4* __restrict__ surf, uint32_t npix, float min_d, float max_d,
                                                     float* __restrict__ depth, float4* __restrict__ normal_rough, float4* __restrict__ albedo) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const size_t np = npix;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const float4 g = surf[surf_at(np, 1, i)];                                  // normal, signed depth
        float t = fabsf(g.w);
        if (depth) depth[i] = t < 0.f ? 0.f : xdiv(t - fminf(min_d, t), fmaxf(max_d, t) - fminf(min_d, t));
        if (normal_rough) {
            const float roughness = unpack8(__float_as_uint(surf[surf_at(np, 8, i)].x), 24);
            normal_rough[i] = make_float4(half_round(g.x), half_round(g.y), half_round(g.z), half_round(roughness));
        }
        if (albedo) albedo[i] = surf[surf_at(np, 5, i)];
    }
}

inline int persistent_grid(const LaunchCfg& cfg, int per_sm) { return cfg.sms * per_sm; }

} // namespace

void launch_raygen(const LaunchCfg& cfg, const FrameView& fv, const CameraBasis& cam, uint32_t frame_count) {
    k_raygen<<<persistent_grid(cfg, 8), kBlock, 0, cfg.stream>>>(fv, cam, frame_count); LB_LAUNCH_CHECK();
}
void launch_extend(const LaunchCfg& cfg, const FrameView& fv, const BvhView& bvh, int queue, uint32_t ticket, bool primary, float tmin, float tmax) {
    k_extend<<<persistent_grid(cfg, 4), kBlock, 0, cfg.stream>>>(bvh, fv.rays[queue].o, fv.rays[queue].d, &fv.counters[queue ? CNT_RAYS_B : CNT_RAYS_A],
        &fv.counters[CNT_TICKET0 + ticket], primary ? fv.primary_hits : fv.hits, tmin, tmax, &fv.stats[STAT_EXTEND], cfg.trace); LB_LAUNCH_CHECK();
}
void launch_shade(const LaunchCfg& cfg, const FrameView& fv, const SceneView& sc, int queue, const ShadeArgs& a) {
    if (a.depth == 0 && !a.do_nee) k_shade<true, false><<<persistent_grid(cfg, 4), kBlock, 0, cfg.stream>>>(fv, sc, queue, a);
    else if (a.depth == 0) k_shade<true><<<persistent_grid(cfg, 4), kBlock, 0, cfg.stream>>>(fv, sc, queue, a);
    else k_shade<false><<<persistent_grid(cfg, 4), kBlock, 0, cfg.stream>>>(fv, sc, queue, a);
    LB_LAUNCH_CHECK();
}
void launch_shadow(const LaunchCfg& cfg, const FrameView& fv, const BvhView& bvh, uint32_t ticket, float tmin) {
    k_shadow<<<persistent_grid(cfg, 4), kBlock, 0, cfg.stream>>>(bvh, fv.shadow, &fv.counters[CNT_SHADOW], &fv.counters[CNT_TICKET0 + ticket],
        fv.channels, fv.npix, tmin, &fv.stats[STAT_SHADOW], cfg.trace_any); LB_LAUNCH_CHECK();
}
void launch_tail(const LaunchCfg& cfg, const FrameView& fv, const SceneView& sc, const BvhView& bvh, const BvhView& bvh_any, int queue, uint32_t first_depth, uint32_t max_depth,
                 uint32_t seed, float tmin, float tmax) {
    k_tail<<<cfg.sms * 8, 128, 0, cfg.stream>>>(fv, sc, bvh, bvh_any, queue, TailArgs{first_depth, max_depth, seed, tmin, tmax}); LB_LAUNCH_CHECK();
}
void launch_merge(const LaunchCfg& cfg, const FrameView& fv, int blend, uint32_t blend_count) {
    k_merge<<<persistent_grid(cfg, 8), kBlock, 0, cfg.stream>>>(fv, blend, blend_count); LB_LAUNCH_CHECK();
}
void launch_gbuffer(const LaunchCfg& cfg, const float4* surf, uint32_t npix, float min_d, float max_d, float* depth, float4* normal_rough, float4* albedo) {
    k_gbuffer<<<persistent_grid(cfg, 8), kBlock, 0, cfg.stream>>>(surf, npix, min_d, max_d, depth, normal_rough, albedo); LB_LAUNCH_CHECK();
}
void launch_resolve(const LaunchCfg& cfg, const FrameView& fv, float inv_frames) {
    k_resolve<<<persistent_grid(cfg, 8), kBlock, 0, cfg.stream>>>(fv, inv_frames); LB_LAUNCH_CHECK();
}
#ifdef LB_TRACE_STATS
void dump_trace_stats_wavefront() {
    unsigned long long h[8]; cudaMemcpyFromSymbol(h, g_trace_stats, sizeof h);
    for (int a = 0; a < 2; ++a) if (h[4 * a + 2]) fprintf(stderr, "TRACE stats [wavefront %s]: rays %llu, node visits / ray %.2f, triangle tests / ray %.2f (%.2f pass the edge test)\n", a ? "any-hit" : "closest",
        h[4 * a + 2], (double)h[4 * a] / h[4 * a + 2], (double)h[4 * a + 1] / h[4 * a + 2], (double)h[4 * a + 3] / h[4 * a + 2]);
    memset(h, 0, sizeof h); cudaMemcpyToSymbol(g_trace_stats, h, sizeof h);
}
#endif
} // namespace lb
