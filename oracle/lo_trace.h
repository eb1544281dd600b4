// CPU ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path.
//
// Ray/triangle intersection and a binned-SAH binary BVH for the extend / shadow / visibility passes.
// The reference delegates this arithmetic to closed-source OptiX 7.3 (optixTrace call sites
// /root/reference/Lumen_Engine/LumenPT/src/Shaders/WaveFrontShaders.cu:63-76,128-140,197-210), so there is
// no reference source to restate: parity is UNPINNED for this part. The canonical definition is
//   - world-space triangles (instance transform applied with lo::xform_point),
//   - the watertight test of Woop, Benthin, Wald 2013 with the exact operation order written below,
//   - accept tmin < t < tmax, no culling (OPTIX_RAY_FLAG_NONE, WaveFrontShaders.cu:70),
//   - closest hit = smallest t, ties broken towards the smaller (instanceId, primitiveIndex),
//   - barycentrics (u,v) = weights of vertex 1 and 2 (optixGetTriangleBarycentrics convention).
// Because the accepted set and t of every (ray, triangle) pair are a pure function of the ray and the three
// vertices, any conservative acceleration structure returns the same hit: the GPU's 8-wide BVH and this
// binary BVH are independent implementations of the same function.
#pragma once
#include "lo_math.h"
#include <vector>
#include <algorithm>
#include <cfloat>

namespace lo {

struct RayShear { int kx, ky, kz; float sx, sy, sz; };

static inline RayShear make_shear(const V3& d) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    RayShear r;
    r.kz = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    r.kx = (r.kz + 1) % 3; r.ky = (r.kx + 1) % 3;
    if (comp(d, r.kz) < 0.0f) std::swap(r.kx, r.ky);
    const float dz = comp(d, r.kz);
    r.sx = comp(d, r.kx) / dz; r.sy = comp(d, r.ky) / dz; r.sz = 1.0f / dz;
    return r;
}

// Returns true and (t, u, v) when the ray hits the triangle anywhere on the line (range test is the caller's).
static inline bool tri_test(const V3& org, const RayShear& s, const V3& p0, const V3& p1, const V3& p2, float& t, float& u, float& v) {
    const V3 A = p0 - org, B = p1 - org, C = p2 - org;
    const float Akz = comp(A, s.kz), Bkz = comp(B, s.kz), Ckz = comp(C, s.kz);
    const float Ax = fmaf(-s.sx, Akz, comp(A, s.kx)), Ay = fmaf(-s.sy, Akz, comp(A, s.ky));
    const float Bx = fmaf(-s.sx, Bkz, comp(B, s.kx)), By = fmaf(-s.sy, Bkz, comp(B, s.ky));
    const float Cx = fmaf(-s.sx, Ckz, comp(C, s.kx)), Cy = fmaf(-s.sy, Ckz, comp(C, s.ky));
    // Edge functions with UNFUSED products: for an edge shared by two triangles the two evaluations are then exact negations
    // of each other, which is what makes the test watertight (a fused multiply-add rounds only one product and breaks this).
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
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

struct Tri { V3 p0, p1, p2; uint32_t inst, prim; };
struct HitRec { uint32_t inst, prim; float u, v, t; };

struct Bvh2 {
    struct Node { V3 lo, hi; uint32_t left, count; };   // count > 0: leaf, left = first index into order
    std::vector<Node> nodes;
    std::vector<uint32_t> order;
    const std::vector<Tri>* tris = nullptr;

    void build(const std::vector<Tri>& t) {
        tris = &t; nodes.clear(); order.resize(t.size());
        for (size_t i = 0; i < t.size(); ++i) order[i] = (uint32_t)i;
        if (t.empty()) return;
        std::vector<V3> lo(t.size()), hi(t.size()), ce(t.size());
        for (size_t i = 0; i < t.size(); ++i) {
            lo[i] = v3(fminf(t[i].p0.x, fminf(t[i].p1.x, t[i].p2.x)), fminf(t[i].p0.y, fminf(t[i].p1.y, t[i].p2.y)), fminf(t[i].p0.z, fminf(t[i].p1.z, t[i].p2.z)));
            hi[i] = v3(fmaxf(t[i].p0.x, fmaxf(t[i].p1.x, t[i].p2.x)), fmaxf(t[i].p0.y, fmaxf(t[i].p1.y, t[i].p2.y)), fmaxf(t[i].p0.z, fmaxf(t[i].p1.z, t[i].p2.z)));
            ce[i] = (lo[i] + hi[i]) * 0.5f;
        }
        nodes.reserve(t.size() * 2);
        nodes.push_back({});
        struct Task { uint32_t node, first, count; };
        std::vector<Task> stack{{0u, 0u, (uint32_t)t.size()}};
        while (!stack.empty()) {
            const Task task = stack.back(); stack.pop_back();
            V3 blo = v3(FLT_MAX), bhi = v3(-FLT_MAX), clo = v3(FLT_MAX), chi = v3(-FLT_MAX);
            for (uint32_t i = task.first; i < task.first + task.count; ++i) {
                const uint32_t k = order[i];
                blo = v3(fminf(blo.x, lo[k].x), fminf(blo.y, lo[k].y), fminf(blo.z, lo[k].z));
                bhi = v3(fmaxf(bhi.x, hi[k].x), fmaxf(bhi.y, hi[k].y), fmaxf(bhi.z, hi[k].z));
                clo = v3(fminf(clo.x, ce[k].x), fminf(clo.y, ce[k].y), fminf(clo.z, ce[k].z));
                chi = v3(fmaxf(chi.x, ce[k].x), fmaxf(chi.y, ce[k].y), fmaxf(chi.z, ce[k].z));
            }
            // conservative padding: the slab test below is evaluated in floating point
            const V3 pad = v3(1e-5f) + v3(fmaxf(fabsf(blo.x), fabsf(bhi.x)), fmaxf(fabsf(blo.y), fabsf(bhi.y)), fmaxf(fabsf(blo.z), fabsf(bhi.z))) * 1e-5f;
            nodes[task.node].lo = blo - pad; nodes[task.node].hi = bhi + pad;
            const V3 ext = chi - clo;
            const int axis = (ext.x >= ext.y && ext.x >= ext.z) ? 0 : (ext.y >= ext.z ? 1 : 2);
            if (task.count <= 4 || comp(ext, axis) <= 0.0f) { nodes[task.node].left = task.first; nodes[task.node].count = task.count; continue; }
            constexpr int NB = 16;
            struct Bin { V3 lo, hi; uint32_t n; } bins[NB];
            for (auto& b : bins) { b.lo = v3(FLT_MAX); b.hi = v3(-FLT_MAX); b.n = 0; }
            const float c0 = comp(clo, axis), scale = NB / comp(ext, axis);
            auto bin_of = [&](uint32_t k) { int b = (int)((comp(ce[k], axis) - c0) * scale); return b < 0 ? 0 : (b >= NB ? NB - 1 : b); };
            for (uint32_t i = task.first; i < task.first + task.count; ++i) {
                const uint32_t k = order[i]; Bin& b = bins[bin_of(k)]; b.n++;
                b.lo = v3(fminf(b.lo.x, lo[k].x), fminf(b.lo.y, lo[k].y), fminf(b.lo.z, lo[k].z));
                b.hi = v3(fmaxf(b.hi.x, hi[k].x), fmaxf(b.hi.y, hi[k].y), fmaxf(b.hi.z, hi[k].z));
            }
            auto area = [](const V3& a, const V3& b) { const V3 e = b - a; return e.x < 0 ? 0.f : 2.f * (e.x * e.y + e.y * e.z + e.z * e.x); };
            float rightA[NB]; uint32_t rightN[NB];
            { V3 a = v3(FLT_MAX), b = v3(-FLT_MAX); uint32_t n = 0;
              for (int i = NB - 1; i > 0; --i) { a = v3(fminf(a.x, bins[i].lo.x), fminf(a.y, bins[i].lo.y), fminf(a.z, bins[i].lo.z)); b = v3(fmaxf(b.x, bins[i].hi.x), fmaxf(b.y, bins[i].hi.y), fmaxf(b.z, bins[i].hi.z)); n += bins[i].n; rightA[i] = area(a, b); rightN[i] = n; } }
            float best = FLT_MAX; int bestSplit = -1;
            { V3 a = v3(FLT_MAX), b = v3(-FLT_MAX); uint32_t n = 0;
              for (int i = 0; i < NB - 1; ++i) { a = v3(fminf(a.x, bins[i].lo.x), fminf(a.y, bins[i].lo.y), fminf(a.z, bins[i].lo.z)); b = v3(fmaxf(b.x, bins[i].hi.x), fmaxf(b.y, bins[i].hi.y), fmaxf(b.z, bins[i].hi.z)); n += bins[i].n;
                  if (n == 0 || rightN[i + 1] == 0) continue;
                  const float cost = area(a, b) * n + rightA[i + 1] * rightN[i + 1];
                  if (cost < best) { best = cost; bestSplit = i; } } }
            uint32_t mid;
            if (bestSplit < 0) { mid = task.first + task.count / 2;
                std::nth_element(order.begin() + task.first, order.begin() + mid, order.begin() + task.first + task.count, [&](uint32_t a, uint32_t b) { return comp(ce[a], axis) < comp(ce[b], axis); }); }
            else { mid = (uint32_t)(std::partition(order.begin() + task.first, order.begin() + task.first + task.count, [&](uint32_t k) { return bin_of(k) <= bestSplit; }) - order.begin()); }
            if (mid == task.first || mid == task.first + task.count) mid = task.first + task.count / 2;
            const uint32_t l = (uint32_t)nodes.size(); nodes.push_back({}); nodes.push_back({});
            nodes[task.node].left = l; nodes[task.node].count = 0;
            stack.push_back({l, task.first, mid - task.first});
            stack.push_back({l + 1, mid, task.first + task.count - mid});
        }
    }

    static inline bool slab(const Node& n, const V3& o, const V3& inv, float tmin, float tmax) {
        float t0 = (n.lo.x - o.x) * inv.x, t1 = (n.hi.x - o.x) * inv.x; if (t0 > t1) std::swap(t0, t1);
        float u0 = (n.lo.y - o.y) * inv.y, u1 = (n.hi.y - o.y) * inv.y; if (u0 > u1) std::swap(u0, u1);
        float w0 = (n.lo.z - o.z) * inv.z, w1 = (n.hi.z - o.z) * inv.z; if (w0 > w1) std::swap(w0, w1);
        // NaN (0 * inf) compares false and therefore never rejects: conservative
        const float tn = fmaxf(fmaxf(t0, u0), fmaxf(w0, tmin)), tf = fminf(fminf(t1, u1), fminf(w1, tmax));
        return !(tn > tf + fabsf(tf) * 1e-5f + 1e-6f);
    }

    // closest hit; returns false on miss
    bool closest(const V3& o, const V3& d, float tmin, float tmax, HitRec& out) const {
        if (nodes.empty()) return false;
        const RayShear sh = make_shear(d);
        const V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        bool found = false; float best = tmax; uint32_t bi = 0, bp = 0; float bu = 0, bv = 0;
        uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node& n = nodes[stack[--sp]];
            if (!slab(n, o, inv, tmin, best)) continue;
            if (n.count) {
                for (uint32_t i = n.left; i < n.left + n.count; ++i) {
                    const Tri& tr = (*tris)[order[i]]; float t, u, v;
                    if (!tri_test(o, sh, tr.p0, tr.p1, tr.p2, t, u, v)) continue;
                    if (!(t > tmin)) continue;
                    const bool better = found ? (t < best || (t == best && (tr.inst < bi || (tr.inst == bi && tr.prim < bp)))) : (t < tmax);
                    if (better) { found = true; best = t; bi = tr.inst; bp = tr.prim; bu = u; bv = v; }
                }
            } else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
        }
        if (found) out = {bi, bp, bu, bv, best};
        return found;
    }
    bool any(const V3& o, const V3& d, float tmin, float tmax) const {
        if (nodes.empty() || !(tmax > tmin)) return false;
        const RayShear sh = make_shear(d);
        const V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node& n = nodes[stack[--sp]];
            if (!slab(n, o, inv, tmin, tmax)) continue;
            if (n.count) {
                for (uint32_t i = n.left; i < n.left + n.count; ++i) {
                    const Tri& tr = (*tris)[order[i]]; float t, u, v;
                    if (tri_test(o, sh, tr.p0, tr.p1, tr.p2, t, u, v) && t > tmin && t < tmax) return true;
                }
            } else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
        }
        return false;
    }
};

// brute force versions (used by the tests to validate the BVH)
static inline bool closest_brute(const std::vector<Tri>& tris, const V3& o, const V3& d, float tmin, float tmax, HitRec& out) {
    const RayShear sh = make_shear(d);
    bool found = false; float best = tmax; uint32_t bi = 0, bp = 0; float bu = 0, bv = 0;
    for (const Tri& tr : tris) {
        float t, u, v;
        if (!tri_test(o, sh, tr.p0, tr.p1, tr.p2, t, u, v)) continue;
        if (!(t > tmin)) continue;
        const bool better = found ? (t < best || (t == best && (tr.inst < bi || (tr.inst == bi && tr.prim < bp)))) : (t < tmax);
        if (better) { found = true; best = t; bi = tr.inst; bp = tr.prim; bu = u; bv = v; }
    }
    if (found) out = {bi, bp, bu, bv, best};
    return found;
}

} // namespace lo
