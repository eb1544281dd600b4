// lb_scene.cu — device-side scene preparation: instance flattening, emissive triangle detection, light list + CDF.
//
// Reference behaviour (under /root/reference/Lumen_Engine/LumenPT/src/):
//   FindEmissivesGpu (serial <<<1,1>>> loop in the reference)   CUDAKernels/WaveFrontKernels/GPUEmissiveLookup.cu:13-109
//   BuildLightDataBuffer + BuildLightDataInstance               Framework/LightDataBuffer.cpp:37-125, CUDAKernels/WaveFrontKernels/GPUDataBufferKernels.cu:66-186
//   FillCDF (thrust::sort by mean radiance + inclusive_scan)    CUDAKernels/ReSTIRKernels.cu:49-130
// Canonical choices (SURVEY hazards 4): lights are compacted in (scene-table entry, triangle) order, sorted with a
// STABLE radix sort, and the prefix sum is a fixed-shape blocked scan (256-element blocks summed left to right,
// block totals summed left to right) so that the CDF is reproducible bit for bit.
#include "lb_kernels.h"
#include "lb_shade.cuh"
#include <cub/cub.cuh>

namespace lb {

namespace {

__device__ __forceinline__ uint32_t entry_of(const DevEntry* entries, uint32_t num_entries, uint32_t g) {
    uint32_t lo = 0, hi = num_entries - 1u;          // last entry with tri_offset <= g
    while (lo < hi) { const uint32_t mid = (lo + hi + 1u) >> 1; if (entries[mid].tri_offset <= g) lo = mid; else hi = mid - 1u; }
    return lo;
}

__global__ void k_flatten(ScenePrepIn in, DevTri* __restrict__ out) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < in.total_tris; g += stride) {
        const uint32_t e = entry_of(in.entries, in.num_entries, g);
        const DevEntry& en = in.entries[e];
        const uint32_t t = g - en.tri_offset;
        const uint32_t i0 = in.indices[en.index_base + 3u * t] + en.vertex_base, i1 = in.indices[en.index_base + 3u * t + 1u] + en.vertex_base, i2 = in.indices[en.index_base + 3u * t + 2u] + en.vertex_base;
        const float3 p0 = xform_point(en.m, f3(in.vtx_pos[i0])), p1 = xform_point(en.m, f3(in.vtx_pos[i1])), p2 = xform_point(en.m, f3(in.vtx_pos[i2]));
        DevTri tr; tr.v0 = f4(p0, __uint_as_float(e)); tr.v1 = f4(p1, __uint_as_float(t)); tr.v2 = f4(p2, 0.f);
        out[g] = tr;
    }
}

__global__ void k_find_emissives(SceneView sc, const DevPrimRange* __restrict__ prims, uint32_t num_prims, uint32_t total, uint8_t* __restrict__ flags, uint32_t* __restrict__ counts) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        uint32_t lo = 0, hi = num_prims - 1u;
        while (lo < hi) { const uint32_t mid = (lo + hi + 1u) >> 1; if (prims[mid].flag_offset <= g) lo = mid; else hi = mid - 1u; }
        const DevPrimRange pr = prims[lo];
        const uint32_t t = g - pr.flag_offset;
        const DevMaterial& m = sc.materials[pr.material];
        uint8_t f = 0;
        if (!(m.mat.emissive.x == 0.f && m.mat.emissive.y == 0.f && m.mat.emissive.z == 0.f)) {
            const uint32_t i0 = sc.indices[pr.index_base + 3u * t] + pr.vertex_base, i1 = sc.indices[pr.index_base + 3u * t + 1u] + pr.vertex_base, i2 = sc.indices[pr.index_base + 3u * t + 2u] + pr.vertex_base;
            const float2 c = ((make_float2(sc.vtx_nu[i0].w, sc.vtx_tv[i0].w) + make_float2(sc.vtx_nu[i1].w, sc.vtx_tv[i1].w)) + make_float2(sc.vtx_nu[i2].w, sc.vtx_tv[i2].w)) * (1.f / 3.f);
            const float4 e = m.mat.emissive * tex2d(sc, m.tex_emissive, c.x, c.y);
            if (e.x > 0.f || e.y > 0.f || e.z > 0.f) { f = 1; atomicAdd(&counts[lo], 1u); }
        }
        flags[g] = f;
    }
}

// one emissive triangle of (entry e, triangle t) -> TriangleLight; returns false when it does not emit
__device__ bool make_light(const SceneView& sc, const ScenePrepIn& in, const uint8_t* prim_flags, uint32_t g, DevLight& l) {
    const uint32_t e = entry_of(in.entries, in.num_entries, g);
    const DevEntry& en = in.entries[e];
    if (!en.lights_on) return false;
    const uint32_t t = g - en.tri_offset;
    if (!(en.em_mode == 2 || prim_flags[en.flag_offset + t])) return false;
    const uint32_t i0 = in.indices[en.index_base + 3u * t] + en.vertex_base, i1 = in.indices[en.index_base + 3u * t + 1u] + en.vertex_base, i2 = in.indices[en.index_base + 3u * t + 2u] + en.vertex_base;
    const DevMaterial& m = sc.materials[en.material];
    float4 em;
    if (en.em_mode == 0) {
        const float2 c = ((make_float2(sc.vtx_nu[i0].w, sc.vtx_tv[i0].w) + make_float2(sc.vtx_nu[i1].w, sc.vtx_tv[i1].w)) + make_float2(sc.vtx_nu[i2].w, sc.vtx_tv[i2].w)) * (1.f / 3.f);
        em = tex2d(sc, m.tex_emissive, c.x, c.y); em = em * (m.mat.emissive * en.em_scale);
    } else {
        em = make_float4(en.em_r, en.em_g, en.em_b, en.em_scale) * en.em_scale;
    }
    if (!(em.x > 0.f || em.y > 0.f || em.z > 0.f)) return false;
    l.p0 = xform_point(en.m, f3(in.vtx_pos[i0])); l.p1 = xform_point(en.m, f3(in.vtx_pos[i1])); l.p2 = xform_point(en.m, f3(in.vtx_pos[i2]));
    l.radiance = f3(em.x, em.y, em.z);
    l.normal = normalize(xform_vector(en.m, ((f3(sc.vtx_nu[i0]) + f3(sc.vtx_nu[i1])) + f3(sc.vtx_nu[i2])) * (1.f / 3.f)));
    const float3 a = l.p0 - l.p1, b = l.p0 - l.p2;
    const float cx = a.y * b.z - b.y * a.z, cy = a.x * b.z - b.x * a.z, cz = a.x * b.y - b.x * a.y;
    l.area = sqrtf(cx * cx + cy * cy + cz * cz) / 2.0f;
    return true;
}

__global__ void k_light_flags(SceneView sc, ScenePrepIn in, const uint8_t* __restrict__ prim_flags, uint32_t* __restrict__ flags) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < in.total_tris; g += stride) { DevLight l; flags[g] = make_light(sc, in, prim_flags, g, l) ? 1u : 0u; }
}
__global__ void k_light_emit(SceneView sc, ScenePrepIn in, const uint8_t* __restrict__ prim_flags, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ offsets,
                             DevLight* __restrict__ lights, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < in.total_tris; g += stride) {
        if (!flags[g]) continue;
        DevLight l; make_light(sc, in, prim_flags, g, l);
        const uint32_t k = offsets[g];
        lights[k] = l;
        keys[k] = __float_as_uint((l.radiance.x + l.radiance.y + l.radiance.z) / 3.f);   // positive floats order like their bit patterns
        vals[k] = k;
    }
}
__global__ void k_light_gather(const DevLight* __restrict__ src, const uint32_t* __restrict__ order, uint32_t n, DevLight* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[order[i]];
}
// blocked scan, pass 1: each thread owns one 256-element block and sums it left to right
__global__ void k_cdf_blocks(const DevLight* __restrict__ lights, uint32_t n, float* __restrict__ cdf, float* __restrict__ totals) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nb = (n + 255u) / 256u;
    if (b >= nb) return;
    float s = 0.f;
    const uint32_t end = min(n, (b + 1u) * 256u);
    for (uint32_t i = b * 256u; i < end; ++i) { const DevLight& l = lights[i]; s += (l.radiance.x + l.radiance.y + l.radiance.z) / 3.f; cdf[i] = s; }
    totals[b] = s;
}
// pass 2: running offsets of the block totals, left to right (single thread; <= 4096 blocks for the 1M-light cap)
__global__ void k_cdf_offsets(float* __restrict__ totals, uint32_t nb) {
    float run = 0.f;
    for (uint32_t b = 0; b < nb; ++b) { const float t = totals[b]; totals[b] = run; run = b ? run + t : t; }
}
__global__ void k_cdf_apply(float* __restrict__ cdf, uint32_t n, const float* __restrict__ offsets) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i < 256u) return;
    cdf[i] = offsets[i / 256u] + cdf[i];
}

} // namespace

void launch_flatten(const LaunchCfg& cfg, const ScenePrepIn& in, DevTri* out) {
    if (!in.total_tris) return;
    k_flatten<<<cfg.sms * 8, 256, 0, cfg.stream>>>(in, out); LB_LAUNCH_CHECK();
}

void launch_find_emissives(const LaunchCfg& cfg, const SceneView& sc, const DevPrimRange* prims, uint32_t num_prims, uint32_t total, uint8_t* flags, uint32_t* counts) {
    if (!total) return;
    k_find_emissives<<<cfg.sms * 8, 256, 0, cfg.stream>>>(sc, prims, num_prims, total, flags, counts); LB_LAUNCH_CHECK();
}

void build_lights(const LaunchCfg& cfg, const SceneView& sc, const ScenePrepIn& in, const uint8_t* prim_flags, LightBuild& out) {
    out.num_lights = 0; out.cdf_sum = 0.f;
    const uint32_t n = in.total_tris;
    if (!n) return;
    cudaStream_t s = cfg.stream;
    // scratch from the stream-ordered pool (StreamBuf): a dynamic scene rebuilds the light list after every instance move, and ten cudaMalloc /
    // cudaFree pairs (each a device-wide synchronisation) cost more than the kernels — up to 100 ms of host time on the 10 M-triangle C4
    StreamBuf<uint32_t> flags, offsets; StreamBuf<unsigned char> tmp;
    flags.reserve(n, s); offsets.reserve(n, s);
    k_light_flags<<<cfg.sms * 8, 256, 0, s>>>(sc, in, prim_flags, flags.p); LB_LAUNCH_CHECK();
    size_t tb = 0;
    LB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, flags.p, offsets.p, (int)n, s));
    tmp.reserve(tb, s);
    LB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flags.p, offsets.p, (int)n, s));
    uint32_t last_flag = 0, last_off = 0;
    LB_CUDA(cudaMemcpyAsync(&last_flag, flags.p + (n - 1), 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(cudaMemcpyAsync(&last_off, offsets.p + (n - 1), 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(cudaStreamSynchronize(s));
    const uint32_t nl = last_flag + last_off;
    if (!nl) return;
    StreamBuf<DevLight> unsorted; StreamBuf<uint32_t> keys, keys2, vals, order; StreamBuf<float> totals;
    unsorted.reserve(nl, s); keys.reserve(nl, s); keys2.reserve(nl, s); vals.reserve(nl, s); order.reserve(nl, s);
    out.lights.reserve(nl); out.cdf.reserve(nl);
    k_light_emit<<<cfg.sms * 8, 256, 0, s>>>(sc, in, prim_flags, flags.p, offsets.p, unsorted.p, keys.p, vals.p); LB_LAUNCH_CHECK();
    size_t sb = 0;
    LB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sb, keys.p, keys2.p, vals.p, order.p, (int)nl, 0, 32, s));
    StreamBuf<unsigned char> tmp2; tmp2.reserve(sb, s);
    LB_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, sb, keys.p, keys2.p, vals.p, order.p, (int)nl, 0, 32, s));
    k_light_gather<<<grid_for(nl, 256), 256, 0, s>>>(unsorted.p, order.p, nl, out.lights.p); LB_LAUNCH_CHECK();
    const uint32_t nb = (nl + 255u) / 256u;
    totals.reserve(nb, s);
    k_cdf_blocks<<<grid_for(nb, 64), 64, 0, s>>>(out.lights.p, nl, out.cdf.p, totals.p); LB_LAUNCH_CHECK();
    k_cdf_offsets<<<1, 1, 0, s>>>(totals.p, nb); LB_LAUNCH_CHECK();
    k_cdf_apply<<<grid_for(nl, 256), 256, 0, s>>>(out.cdf.p, nl, totals.p); LB_LAUNCH_CHECK();
    LB_CUDA(cudaMemcpyAsync(&out.cdf_sum, out.cdf.p + (nl - 1), 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(cudaStreamSynchronize(s));
    out.num_lights = nl;
}

} // namespace lb
