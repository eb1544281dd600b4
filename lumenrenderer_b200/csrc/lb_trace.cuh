// lb_trace.cuh — ray/triangle intersection and traversal of the compressed 8-wide BVH.
//
// Replaces the reference's optixTrace calls (LumenPT/src/Shaders/WaveFrontShaders.cu:63-76 closest hit,
// :128-140 shadow any-hit, :197-210 ReSTIR visibility any-hit; no culling, OPTIX_RAY_FLAG_NONE :70) — B200 has no
// RT cores. Hit definition (DESIGN.md "extend"): world-space triangles, the watertight test of Woop/Benthin/Wald
// 2013 with the operation order below, accept tmin < t < tmax, closest = smallest t with ties towards the
// smaller (instance, primitive); barycentrics (u, v) weight vertices 1 and 2 (optixGetTriangleBarycentrics).
// Accepted set and t are a pure function of (ray, triangle), so any conservative BVH returns the same hit.
#pragma once
#include "lb_device.cuh"

namespace lb {

struct RayShear { int kx, ky, kz; float sx, sy, sz; };

LB_D RayShear make_shear(const float3& d) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    RayShear r;
    r.kz = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    if (comp(d, r.kz) < 0.0f) { const int t = r.kx; r.kx = r.ky; r.ky = t; }
    const float dz = comp(d, r.kz);
    r.sx = comp(d, r.kx) / dz; r.sy = comp(d, r.ky) / dz; r.sz = 1.0f / dz;
    return r;
}

// true + (t, u, v) when the ray's line hits the triangle; the range test is the caller's.
LB_D bool tri_test(const float3& org, const RayShear& s, const float3& p0, const float3& p1, const float3& p2, float& t, float& u, float& v) {
    const float3 A = p0 - org, B = p1 - org, C = p2 - org;
    const float Akz = comp(A, s.kz), Bkz = comp(B, s.kz), Ckz = comp(C, s.kz);
    const float Ax = fmaf(-s.sx, Akz, comp(A, s.kx)), Ay = fmaf(-s.sy, Akz, comp(A, s.ky));
    const float Bx = fmaf(-s.sx, Bkz, comp(B, s.kx)), By = fmaf(-s.sy, Bkz, comp(B, s.ky));
    const float Cx = fmaf(-s.sx, Ckz, comp(C, s.kx)), Cy = fmaf(-s.sy, Ckz, comp(C, s.ky));
    // Edge functions with UNFUSED products (explicit _rn intrinsics are never contracted): for an edge shared by two triangles
    // the two evaluations are exact negations of each other — the watertightness property. An FMA would round only one product.
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {          // edge-on: redo the edge functions in double (rare)
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = (U + V) + W;
    if (det == 0.0f) return false;
    const float Az = s.sz * Akz, Bz = s.sz * Bkz, Cz = s.sz * Ckz;
    const float T = fmaf(U, Az, fmaf(V, Bz, W * Cz));
    t = T / det; u = V / det; v = W / det;
    return true;
}

struct HitInfo { uint32_t inst, prim; float u, v, t; };

constexpr int kTraceStack = 40;

LB_D float safe_rcp(float d) { return 1.0f / (fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); }

// Traversal of the compressed wide BVH: one ray per thread, node groups / triangle groups as in
// Ylitie et al. 2017. The octant trick orders children front to back (slot ^ octant, highest bit first).
template <bool ANY>
LB_D bool bvh8_trace(const BvhView& bvh, const float3& o, const float3& d, float tmin, float tmax, HitInfo& hit) {
    if (bvh.num_tris == 0 || !(tmax > tmin)) return false;
    const RayShear sh = make_shear(d);
    const float3 idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    // octant and near/far swizzle come from the sign of the INVERSE direction, so that a component of -0.0 (inverse -1e20)
    // is treated consistently by both
    const bool neg_x = idir.x < 0.f, neg_y = idir.y < 0.f, neg_z = idir.z < 0.f;
    const uint32_t octinv = (neg_x ? 0u : 4u) | (neg_y ? 0u : 2u) | (neg_z ? 0u : 1u);
    const uint32_t octinv4 = octinv * 0x01010101u;
    float best = tmax; bool found = false; uint32_t bi = 0, bp = 0; float bu = 0.f, bv = 0.f;

    uint2 stack[kTraceStack]; int sp = 0;
    uint2 ngroup = make_uint2(0u, 0x80000000u);
    uint2 tgroup = make_uint2(0u, 0u);
    for (;;) {
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t hits_imask = ngroup.y;
            const uint32_t child_bit = 31u - (uint32_t)__clz(hits_imask);
            const uint32_t child_base = ngroup.x;
            ngroup.y &= ~(1u << child_bit);
            if (ngroup.y > 0x00FFFFFFu && sp < kTraceStack) stack[sp++] = ngroup;
            const uint32_t slot = (child_bit - 24u) ^ (octinv & 7u);
            const uint32_t rel = __popc(hits_imask & ~(0xFFFFFFFFu << slot));
            const uint4* np = reinterpret_cast<const uint4*>(bvh.nodes + (child_base + rel));
            const uint4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3), q4 = __ldg(np + 4);

            const float ax = __uint_as_float((q0.w & 0xFFu) << 23) * idir.x;
            const float ay = __uint_as_float(((q0.w >> 8) & 0xFFu) << 23) * idir.y;
            const float az = __uint_as_float(((q0.w >> 16) & 0xFFu) << 23) * idir.z;
            const float ox = (__uint_as_float(q0.x) - o.x) * idir.x;
            const float oy = (__uint_as_float(q0.y) - o.y) * idir.y;
            const float oz = (__uint_as_float(q0.z) - o.z) * idir.z;
            uint32_t hitmask = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t meta4 = h ? q1.w : q1.z;
                const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xFFu;
                const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
                const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
                const uint32_t lox = h ? q2.y : q2.x, loy = h ? q2.w : q2.z, loz = h ? q3.y : q3.x;
                const uint32_t hix = h ? q3.w : q3.z, hiy = h ? q4.y : q4.x, hiz = h ? q4.w : q4.z;
                const uint32_t nx = neg_x ? hix : lox, fx = neg_x ? lox : hix;
                const uint32_t ny = neg_y ? hiy : loy, fy = neg_y ? loy : hiy;
                const uint32_t nz = neg_z ? hiz : loz, fz = neg_z ? loz : hiz;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int sft = 8 * j;
                    const float t0x = fmaf((float)((nx >> sft) & 0xFFu), ax, ox), t1x = fmaf((float)((fx >> sft) & 0xFFu), ax, ox);
                    const float t0y = fmaf((float)((ny >> sft) & 0xFFu), ay, oy), t1y = fmaf((float)((fy >> sft) & 0xFFu), ay, oy);
                    const float t0z = fmaf((float)((nz >> sft) & 0xFFu), az, oz), t1z = fmaf((float)((fz >> sft) & 0xFFu), az, oz);
                    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
                    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, best));
                    if (tn <= tf) {
                        const uint32_t bits = (child_bits4 >> sft) & 0xFFu;
                        const uint32_t idx = (bit_index4 >> sft) & 0xFFu;
                        hitmask |= bits << idx;
                    }
                }
            }
            ngroup.x = q1.x; ngroup.y = (hitmask & 0xFF000000u) | (q0.w >> 24);
            tgroup.x = q1.y; tgroup.y = hitmask & 0x00FFFFFFu;
        } else {
            tgroup = ngroup; ngroup = make_uint2(0u, 0u);
        }

        while (tgroup.y != 0u) {
            const uint32_t k = (uint32_t)__ffs(tgroup.y) - 1u;
            tgroup.y &= tgroup.y - 1u;
            const float4* tp = reinterpret_cast<const float4*>(bvh.tris + (tgroup.x + k));
            const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
            float t, u, v;
            if (!tri_test(o, sh, f3(v0), f3(v1), f3(v2), t, u, v)) continue;
            if (!(t > tmin)) continue;
            if (ANY) { if (t < tmax) return true; continue; }
            const uint32_t ti = __float_as_uint(v0.w), tpi = __float_as_uint(v1.w);
            const bool better = found ? (t < best || (t == best && (ti < bi || (ti == bi && tpi < bp)))) : (t < tmax);
            if (better) { found = true; best = t; bi = ti; bp = tpi; bu = u; bv = v; }
        }

        if (ngroup.y <= 0x00FFFFFFu) {
            if (sp == 0) break;
            ngroup = stack[--sp];
        }
    }
    if (ANY) return false;
    if (found) { hit.inst = bi; hit.prim = bp; hit.u = bu; hit.v = bv; hit.t = best; }
    return found;
}

} // namespace lb
