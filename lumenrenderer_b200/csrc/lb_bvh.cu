// lb_bvh.cu — GPU acceleration-structure build for the extend / shadow / visibility kernels.
//
// Replaces optixAccelBuild (GAS per primitive + IAS per instance + scene IAS, LumenPT/src/Framework/OptixWrapper.cpp:46-131,
// PTScene.cpp:74-156). B200 design: with 180 GB of HBM the instances are flattened into world space (48 B per
// triangle), so there is ONE bounding hierarchy and no per-instance ray transform in the traversal loop.
//   1. per-triangle padded AABB + centroid bounds           (k_tri_bounds)
//   2. 63-bit Morton codes, CUB radix sort                   (k_morton, cub::DeviceRadixSort)
//   3. binary hierarchy over the Morton order, one of
//      a. PLOC (Meister & Bittner 2018, parallel locally-ordered clustering): every cluster looks at its 2*16 neighbours in
//         Morton order for the partner that minimises the surface area of the merged box; mutual nearest neighbours merge;
//         repeat until one cluster is left (k_ploc_nearest, k_ploc_merge, CUB scan, k_ploc_compact). SAH-quality, default.
//      b. binary radix tree, Karras 2012 (k_hierarchy) + bottom-up AABB refit with arrival counters (k_refit). Fastest build
//         (LB_BVH_BUILDER=lbvh), ~1.5x more node visits per ray.
//   4. level-synchronous greedy collapse into compressed 8-wide nodes (80 B, quantised child boxes, octant-ordered
//      slots, <= 3 triangles per leaf child), triangles re-emitted in node order            (k_collapse)
#include "lb_host.h"
#include <cub/cub.cuh>
#include <cfloat>
#include <vector>

namespace lb {

namespace {

#ifndef LB_LEAF_MAX
#define LB_LEAF_MAX 3            // triangles per leaf child (the node meta byte has 3 unary bits); A/B builds: 1, 2
#endif
constexpr uint32_t kLeafMax = LB_LEAF_MAX;

__device__ __forceinline__ int float_order(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// conservative padding: the slab test of the traversal is evaluated in floating point
__device__ __forceinline__ float pad_of(float lo, float hi) { return 1e-5f + fmaxf(fabsf(lo), fabsf(hi)) * 1e-5f; }

__global__ void k_tri_bounds(const DevTri* __restrict__ tris, uint32_t n, float4* __restrict__ lo, float4* __restrict__ hi, int* __restrict__ cbounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 c = f3(0.f); bool ok = i < n;
    if (ok) {
        const float3 a = f3(tris[i].v0), b = f3(tris[i].v1), cc = f3(tris[i].v2);
        float3 l = f3(fminf(a.x, fminf(b.x, cc.x)), fminf(a.y, fminf(b.y, cc.y)), fminf(a.z, fminf(b.z, cc.z)));
        float3 h = f3(fmaxf(a.x, fmaxf(b.x, cc.x)), fmaxf(a.y, fmaxf(b.y, cc.y)), fmaxf(a.z, fmaxf(b.z, cc.z)));
        c = (l + h) * 0.5f;
        const float3 p = f3(pad_of(l.x, h.x), pad_of(l.y, h.y), pad_of(l.z, h.z));
        l = l - p; h = h + p;
        lo[i] = f4(l, 0.f); hi[i] = f4(h, 0.f);
        ok = isfinite(c.x) && isfinite(c.y) && isfinite(c.z);
    }
    // warp reduce, then one atomic per warp and axis
    int mn[3] = {ok ? float_order(c.x) : INT_MAX, ok ? float_order(c.y) : INT_MAX, ok ? float_order(c.z) : INT_MAX};
    int mx[3] = {ok ? float_order(c.x) : INT_MIN, ok ? float_order(c.y) : INT_MIN, ok ? float_order(c.z) : INT_MIN};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        mn[k] = __reduce_min_sync(0xFFFFFFFFu, mn[k]);
        mx[k] = __reduce_max_sync(0xFFFFFFFFu, mx[k]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&cbounds[k], mn[k]); atomicMax(&cbounds[3 + k], mx[k]); }
    }
}

// ---- early split clipping (Ernst & Greiner 2007; the pre-pass PLOC is usually paired with): a triangle whose box is much larger than
// the scene's typical primitive — the floors, walls and arches of a real asset — enters the build as several REFERENCES, one per cell of a
// world-aligned grid its (clipped) area reaches, each with the box of the part inside the cell. The hierarchy then separates space across
// a large triangle instead of wrapping it in one box that overlaps half the scene. The leaf arrays hold the triangle once per reference;
// a ray that meets the same triangle through two leaves computes the same (t, u, v) twice, and the tie rule keeps the first — hits are a
// pure function of (ray, triangle set) as before. Boxes: the clipped polygon's box padded like every leaf box (the clip computes cut points
// in float: the padding is orders above that error, so neighbouring references overlap instead of leaving a gap).
constexpr int kSplitMaxCells = 512;
struct SplitParams { float cell; uint32_t enabled; };

__device__ __forceinline__ float axis_of(const float3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
__device__ __forceinline__ void set_axis(float3& v, int a, float x) { if (a == 0) v.x = x; else if (a == 1) v.y = x; else v.z = x; }
// Sutherland-Hodgman against one axis-aligned half space (keep x_a >= bound when `greater`, else x_a <= bound); returns the new vertex count
__device__ int clip_half_space(const float3* in, int n, float3* out, int a, float bound, bool greater) {
    int m = 0;
    for (int k = 0; k < n; ++k) {
        const float3 p = in[k], q = in[k + 1 == n ? 0 : k + 1];
        const float pa = axis_of(p, a), qa = axis_of(q, a);
        const bool pin = greater ? pa >= bound : pa <= bound, qin = greater ? qa >= bound : qa <= bound;
        if (pin) out[m++] = p;
        if (pin != qin) {
            const float t = (bound - pa) / (qa - pa);
            float3 x = f3(p.x + t * (q.x - p.x), p.y + t * (q.y - p.y), p.z + t * (q.z - p.z));
            set_axis(x, a, bound);
            out[m++] = x;
        }
    }
    return m;
}
// the part of triangle (a, b, c) inside the box [lo, hi]: false when (numerically) nothing is left; else its bounds
__device__ bool clipped_bounds(const float3& a, const float3& b, const float3& c, const float3& lo, const float3& hi, float3& blo, float3& bhi) {
    float3 p0[10], p1[10];
    p0[0] = a; p0[1] = b; p0[2] = c; int n = 3;
    for (int ax = 0; ax < 3 && n > 0; ++ax) {
        n = clip_half_space(p0, n, p1, ax, axis_of(lo, ax), true);
        if (n == 0) break;
        n = clip_half_space(p1, n, p0, ax, axis_of(hi, ax), false);
    }
    if (n == 0) return false;
    blo = p0[0]; bhi = p0[0];
    for (int k = 1; k < n; ++k) { blo = f3(fminf(blo.x, p0[k].x), fminf(blo.y, p0[k].y), fminf(blo.z, p0[k].z)); bhi = f3(fmaxf(bhi.x, p0[k].x), fmaxf(bhi.y, p0[k].y), fmaxf(bhi.z, p0[k].z)); }
    return true;
}
// the grid a triangle is cut on: the global cell size, doubled until the triangle's box spans at most kSplitMaxCells cells
struct SplitGrid { float cell; int i0[3], n[3]; bool split; };
__device__ SplitGrid split_grid(const float3& l, const float3& h, float cell0) {
    SplitGrid g; g.cell = cell0;
    const float longest = fmaxf(h.x - l.x, fmaxf(h.y - l.y, h.z - l.z));
    g.split = cell0 > 0.f && longest > cell0 && isfinite(longest);
    g.n[0] = g.n[1] = g.n[2] = 1; g.i0[0] = g.i0[1] = g.i0[2] = 0;
    if (!g.split) return g;
    for (int it = 0; it < 24; ++it) {
        long long total = 1;
        for (int a = 0; a < 3; ++a) {
            const float fl = floorf(axis_of(l, a) / g.cell), fh = floorf(axis_of(h, a) / g.cell);
            g.i0[a] = (int)fl; g.n[a] = (int)(fh - fl) + 1; total *= g.n[a];
        }
        if (total <= kSplitMaxCells) break;
        g.cell *= 2.f;
    }
    return g;
}
// MODE 0: count the references of every triangle; MODE 1: write them (boxes + owning triangle) at offset[i]
template <int MODE>
__global__ void k_split(const DevTri* __restrict__ tris, uint32_t n, SplitParams sp, const uint32_t* __restrict__ offset, uint32_t* __restrict__ counts,
                        float4* __restrict__ rlo, float4* __restrict__ rhi, uint32_t* __restrict__ ref_tri) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 a = f3(tris[i].v0), b = f3(tris[i].v1), c = f3(tris[i].v2);
    const float3 l = f3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
    const float3 h = f3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
    const float3 pad = f3(pad_of(l.x, h.x), pad_of(l.y, h.y), pad_of(l.z, h.z));
    const float3 tl = l - pad, th = h + pad;                    // the triangle's own leaf box (k_tri_bounds)
    uint32_t emitted = 0u; const uint32_t base = MODE ? offset[i] : 0u;
    const SplitGrid g = split_grid(l, h, sp.enabled ? sp.cell : 0.f);
    if (g.split) {
        for (int z = 0; z < g.n[2]; ++z) for (int y = 0; y < g.n[1]; ++y) for (int x = 0; x < g.n[0]; ++x) {
            const float3 cl = f3((float)(g.i0[0] + x) * g.cell, (float)(g.i0[1] + y) * g.cell, (float)(g.i0[2] + z) * g.cell);
            const float3 ch = f3((float)(g.i0[0] + x + 1) * g.cell, (float)(g.i0[1] + y + 1) * g.cell, (float)(g.i0[2] + z + 1) * g.cell);
            // the cell widened by the leaf padding before the clip: a point of the triangle on a cell border belongs to both neighbours
            const float3 cp = f3(pad_of(cl.x, ch.x), pad_of(cl.y, ch.y), pad_of(cl.z, ch.z));
            float3 bl, bh;
            if (!clipped_bounds(a, b, c, cl - cp, ch + cp, bl, bh)) continue;
            if (MODE) {
                const float3 bp = f3(pad_of(bl.x, bh.x), pad_of(bl.y, bh.y), pad_of(bl.z, bh.z));
                bl = bl - bp; bh = bh + bp;
                rlo[base + emitted] = make_float4(fmaxf(bl.x, tl.x), fmaxf(bl.y, tl.y), fmaxf(bl.z, tl.z), 0.f);
                rhi[base + emitted] = make_float4(fminf(bh.x, th.x), fminf(bh.y, th.y), fminf(bh.z, th.z), 0.f);
                ref_tri[base + emitted] = i;
            }
            ++emitted;
        }
    }
    if (emitted == 0u) {                                        // not split (or degenerate: nothing survived the clip): the triangle's own box
        if (MODE) { rlo[base] = f4(tl, 0.f); rhi[base] = f4(th, 0.f); ref_tri[base] = i; }
        emitted = 1u;
    }
    if (!MODE) counts[i] = emitted;
}
// centroid bounds of the references (what k_tri_bounds computes for triangles), for the Morton codes
__global__ void k_ref_bounds(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, int* __restrict__ cbounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < n;
    float3 c = f3(0.f);
    if (in) c = (f3(lo[i]) + f3(hi[i])) * 0.5f;
    const bool ok = in && isfinite(c.x) && isfinite(c.y) && isfinite(c.z);
    int mn[3] = {ok ? float_order(c.x) : INT_MAX, ok ? float_order(c.y) : INT_MAX, ok ? float_order(c.z) : INT_MAX};
    int mx[3] = {ok ? float_order(c.x) : INT_MIN, ok ? float_order(c.y) : INT_MIN, ok ? float_order(c.z) : INT_MIN};
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = __reduce_min_sync(0xFFFFFFFFu, mn[k]); mx[k] = __reduce_max_sync(0xFFFFFFFFu, mx[k]); }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&cbounds[k], mn[k]); atomicMax(&cbounds[3 + k], mx[k]); }
    }
}

__device__ __forceinline__ uint64_t spread21(uint64_t x) {
    x &= 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, const int* __restrict__ cbounds,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 bl = f3(order_float(cbounds[0]), order_float(cbounds[1]), order_float(cbounds[2]));
    const float3 bh = f3(order_float(cbounds[3]), order_float(cbounds[4]), order_float(cbounds[5]));
    const float3 c = (f3(lo[i]) + f3(hi[i])) * 0.5f;
    const float3 e = bh - bl;
    const float sx = e.x > 0.f ? 2097151.f / e.x : 0.f, sy = e.y > 0.f ? 2097151.f / e.y : 0.f, sz = e.z > 0.f ? 2097151.f / e.z : 0.f;
    const uint64_t qx = (uint64_t)fminf(fmaxf((c.x - bl.x) * sx, 0.f), 2097151.f);
    const uint64_t qy = (uint64_t)fminf(fmaxf((c.y - bl.y) * sy, 0.f), 2097151.f);
    const uint64_t qz = (uint64_t)fminf(fmaxf((c.z - bl.z) * sz, 0.f), 2097151.f);
    keys[i] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
    vals[i] = i;
}

// ---- binary radix tree. Node ids: internal i in [0, n-2]; leaf j is id (n-1)+j.
__device__ __forceinline__ int delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, uint2* __restrict__ children, uint32_t* __restrict__ parent, uint2* __restrict__ range) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do { t = (t + 1) >> 1; if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t; } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const uint32_t left = (first == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
    const uint32_t right = (last == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
    children[i] = make_uint2(left, right);
    range[i] = make_uint2((uint32_t)first, (uint32_t)last);
    parent[left] = (uint32_t)i; parent[right] = (uint32_t)i;
}

__global__ void k_refit(const uint32_t* __restrict__ sorted, const float4* __restrict__ tlo, const float4* __restrict__ thi, int n,
                        const uint2* __restrict__ children, const uint32_t* __restrict__ parent, float4* nlo, float4* nhi, uint32_t* count, uint32_t* flags) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t leaf = (uint32_t)(n - 1 + j);
    const uint32_t src = sorted[j];
    nlo[leaf] = tlo[src]; nhi[leaf] = thi[src]; count[leaf] = 1u;
    if (n == 1) return;
    uint32_t cur = parent[leaf];
    for (;;) {
        __threadfence();
        if (atomicAdd(&flags[cur], 1u) == 0u) return;      // first child to arrive stops; the second one owns the node
        __threadfence();
        const uint2 ch = children[cur];
        const float4 al = __ldcg(&nlo[ch.x]), ah = __ldcg(&nhi[ch.x]), bl = __ldcg(&nlo[ch.y]), bh = __ldcg(&nhi[ch.y]);
        nlo[cur] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.f);
        nhi[cur] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
        count[cur] = __ldcg(&count[ch.x]) + __ldcg(&count[ch.y]);
        if (cur == 0u) return;
        cur = parent[cur];
    }
}

// ---- PLOC. Clusters live in three parallel arrays in Morton order: node id, box lo, box hi.
// Neighbour search window of PLOC in Morton order, chosen per hierarchy (bvh_build's ploc_radius): 16 for the closest-hit hierarchy,
// 128 for the any-hit one (measured on C2, DESIGN.md "Two hierarchies").
constexpr int kPlocBlock = 256;

__global__ void k_ploc_init(const uint32_t* __restrict__ sorted, const float4* __restrict__ tlo, const float4* __restrict__ thi, uint32_t n,
                            uint32_t* __restrict__ cl, float4* __restrict__ clo, float4* __restrict__ chi, float4* __restrict__ nlo, float4* __restrict__ nhi, uint32_t* __restrict__ count) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t leaf = n - 1u + j, src = sorted[j];
    const float4 l = tlo[src], h = thi[src];
    cl[j] = leaf; clo[j] = l; chi[j] = h; nlo[leaf] = l; nhi[leaf] = h; count[leaf] = 1u;
}

// nearest neighbour of every cluster within +-kPlocRadius positions: smallest surface area of the merged box, ties to the smaller index
template <int kPlocRadius>
__global__ void __launch_bounds__(kPlocBlock) k_ploc_nearest(const float4* __restrict__ clo, const float4* __restrict__ chi, const uint32_t* __restrict__ m_in, uint32_t* __restrict__ nearest) {
    const uint32_t m = *m_in;                                   // the round's cluster count lives on the device: no host round trip per round
    __shared__ float4 slo[kPlocBlock + 2 * kPlocRadius], shi[kPlocBlock + 2 * kPlocRadius];
    const int first = (int)(blockIdx.x * kPlocBlock) - kPlocRadius;
    for (int t = threadIdx.x; t < kPlocBlock + 2 * kPlocRadius; t += kPlocBlock) {
        const int g = first + t;
        if (g >= 0 && g < (int)m) { slo[t] = clo[g]; shi[t] = chi[g]; }
    }
    __syncthreads();
    const int i = (int)(blockIdx.x * kPlocBlock + threadIdx.x);
    if (i >= (int)m) return;
    const float4 l = slo[threadIdx.x + kPlocRadius], h = shi[threadIdx.x + kPlocRadius];
    float best = FLT_MAX; int bj = -1;
    for (int d = -kPlocRadius; d <= kPlocRadius; ++d) {
        const int j = i + d;
        if (d == 0 || j < 0 || j >= (int)m) continue;
        const float4 ol = slo[threadIdx.x + kPlocRadius + d], oh = shi[threadIdx.x + kPlocRadius + d];
        const float ex = fmaxf(h.x, oh.x) - fminf(l.x, ol.x), ey = fmaxf(h.y, oh.y) - fminf(l.y, ol.y), ez = fmaxf(h.z, oh.z) - fminf(l.z, ol.z);
        const float a = ex * ey + ey * ez + ez * ex;
        if (a < best) { best = a; bj = j; }
    }
    nearest[i] = (uint32_t)bj;
}

// mutual nearest neighbours merge into a new binary node, which takes the place of the left partner
__global__ void k_ploc_merge(uint32_t* __restrict__ cl, float4* __restrict__ clo, float4* __restrict__ chi, const uint32_t* __restrict__ nearest, const uint32_t* __restrict__ m_in,
                             uint2* __restrict__ children, float4* __restrict__ nlo, float4* __restrict__ nhi, uint32_t* __restrict__ count, uint32_t* __restrict__ next_node, uint32_t* __restrict__ keep) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, m = *m_in;
    if (i >= m) return;
    const uint32_t j = nearest[i];
    if (j >= m || nearest[j] != i) { keep[i] = 1u; return; }
    if (i > j) { keep[i] = 0u; return; }
    const uint32_t a = cl[i], b = cl[j];
    const float4 al = clo[i], ah = chi[i], bl = clo[j], bh = chi[j];
    const uint32_t id = atomicAdd(next_node, 1u);
    const float4 l = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.f), h = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
    children[id] = make_uint2(a, b); nlo[id] = l; nhi[id] = h; count[id] = count[a] + count[b];
    cl[i] = id; clo[i] = l; chi[i] = h; keep[i] = 1u;
}

// The last rounds — at most kPlocTail clusters left — in ONE block: clusters in shared memory, the same nearest / merge / compact steps with
// block barriers instead of launches (the ~20 rounds from 1024 clusters down to the root were ~100 launches of a few microseconds of work).
// Same search window, same tie rules as k_ploc_nearest / k_ploc_merge: the tree is the one the launch-per-round loop builds.
constexpr int kPlocTail = 1024;
template <int kPlocRadius>
__global__ void __launch_bounds__(kPlocTail) k_ploc_tail(uint32_t* __restrict__ cl, const float4* __restrict__ clo, const float4* __restrict__ chi, uint32_t* __restrict__ m_io,
                                                         uint2* __restrict__ children, float4* __restrict__ nlo, float4* __restrict__ nhi, uint32_t* __restrict__ count,
                                                         uint32_t* __restrict__ next_node, uint32_t* __restrict__ rounds) {
    __shared__ uint32_t s_cl[kPlocTail], s_near[kPlocTail], s_warp[kPlocTail / 32];
    __shared__ float4 s_lo[kPlocTail], s_hi[kPlocTail];
    const int i = (int)threadIdx.x, lane = i & 31, warp = i >> 5;
    int m = (int)*m_io;
    if (i < m) { s_cl[i] = cl[i]; s_lo[i] = clo[i]; s_hi[i] = chi[i]; }
    __syncthreads();
    uint32_t done = 0u;
    while (m > 1) {
        float4 l = make_float4(0.f, 0.f, 0.f, 0.f), h = l; uint32_t id = 0u;
        if (i < m) {
            l = s_lo[i]; h = s_hi[i]; id = s_cl[i];
            float best = FLT_MAX; int bj = -1;
            const int d0 = max(-kPlocRadius, -i), d1 = min(kPlocRadius, m - 1 - i);
            for (int d = d0; d <= d1; ++d) {
                if (d == 0) continue;
                const float4 ol = s_lo[i + d], oh = s_hi[i + d];
                const float ex = fmaxf(h.x, oh.x) - fminf(l.x, ol.x), ey = fmaxf(h.y, oh.y) - fminf(l.y, ol.y), ez = fmaxf(h.z, oh.z) - fminf(l.z, ol.z);
                const float a = ex * ey + ey * ez + ez * ex;
                if (a < best) { best = a; bj = i + d; }
            }
            s_near[i] = (uint32_t)bj;
        }
        __syncthreads();
        bool keep = false;
        if (i < m) {
            const uint32_t j = s_near[i];
            if (j >= (uint32_t)m || s_near[j] != (uint32_t)i) keep = true;
            else if ((uint32_t)i < j) {
                const uint32_t b = s_cl[j]; const float4 bl = s_lo[j], bh = s_hi[j];
                const uint32_t node = atomicAdd(next_node, 1u);
                l = make_float4(fminf(l.x, bl.x), fminf(l.y, bl.y), fminf(l.z, bl.z), 0.f); h = make_float4(fmaxf(h.x, bh.x), fmaxf(h.y, bh.y), fmaxf(h.z, bh.z), 0.f);
                children[node] = make_uint2(id, b); nlo[node] = l; nhi[node] = h; count[node] = count[id] + count[b];
                id = node; keep = true;
            }
        }
        // order-preserving compaction: ballot inside the warp, running sum over the 32 warp totals
        const uint32_t vote = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_warp[warp] = (uint32_t)__popc(vote);
        __syncthreads();                                        // also: every read of s_cl / s_lo / s_hi of this round is done
        uint32_t base = 0u, total = 0u;
        for (int w = 0; w < kPlocTail / 32; ++w) { const uint32_t c = s_warp[w]; if (w < warp) base += c; total += c; }
        if (keep) { const uint32_t o = base + (uint32_t)__popc(vote & ((1u << lane) - 1u)); s_cl[o] = id; s_lo[o] = l; s_hi[o] = h; }
        __syncthreads();
        m = (int)total; ++done;
    }
    if (i == 0) { cl[0] = s_cl[0]; *m_io = 1u; *rounds += done; }
}

__global__ void k_ploc_compact(const uint32_t* __restrict__ cl, const float4* __restrict__ clo, const float4* __restrict__ chi, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ offset, const uint32_t* __restrict__ m_in,
                               uint32_t* __restrict__ cl_out, float4* __restrict__ clo_out, float4* __restrict__ chi_out, uint32_t* __restrict__ m_out, uint32_t* __restrict__ rounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, m = *m_in;
    if (i >= m) return;
    if (i == 0u && m > 1u) *rounds += 1u;                       // a round launched after the last merge only carries the root across
    if (keep[i]) { const uint32_t o = offset[i]; cl_out[o] = cl[i]; clo_out[o] = clo[i]; chi_out[o] = chi[i]; }
    if (i == m - 1u) *m_out = offset[i] + keep[i];
}

// Copies 1 and 2 of the leaf-ordered triangles with their coordinates rotated: component j of copy r is coordinate (j + r) mod 3 (the ids in
// the w lanes stay). The intersection test shears the triangle along the ray's dominant axis kz and needs its coordinates in the order
// (kz + 1, kz + 2, kz); a ray reads the copy r = (kz + 1) mod 3, where that order is (x, y, z) — no per-triangle component selection
// (it was 10 % of the traversal's instructions, profiles/r01_u_kernels.md). 96 B more per triangle; HBM has room.
__global__ void k_rotate_tris(DevTri* __restrict__ tris, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevTri t = tris[i];
    DevTri a, b;
    a.v0 = make_float4(t.v0.y, t.v0.z, t.v0.x, t.v0.w); a.v1 = make_float4(t.v1.y, t.v1.z, t.v1.x, t.v1.w); a.v2 = make_float4(t.v2.y, t.v2.z, t.v2.x, t.v2.w);
    b.v0 = make_float4(t.v0.z, t.v0.x, t.v0.y, t.v0.w); b.v1 = make_float4(t.v1.z, t.v1.x, t.v1.y, t.v1.w); b.v2 = make_float4(t.v2.z, t.v2.x, t.v2.y, t.v2.w);
    tris[(size_t)n + i] = a; tris[2 * (size_t)n + i] = b;
}

struct WorkItem { uint32_t bnode, wnode; };

__device__ __forceinline__ float half_area(const float4& lo, const float4& hi) {
    const float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
    return ex * ey + ey * ez + ez * ex;
}

// exponent byte e (biased) with 2^(e-127) * 255 >= extent
__device__ __forceinline__ uint32_t quant_exponent(float extent) {
    int e = 0;
    frexpf(extent / 255.f, &e);               // extent/255 = m * 2^e, m in [0.5, 1)  =>  2^e >= extent/255
    e += 127;
    return (uint32_t)max(1, min(254, e));
}

__device__ __forceinline__ uint32_t quant_down(float v, float p, float inv, float scale) { float q = floorf((v - p) * inv); q = fminf(fmaxf(q, 0.f), 255.f); while (q > 0.f && p + q * scale > v) q -= 1.f; return (uint32_t)q; }
__device__ __forceinline__ uint32_t quant_up(float v, float p, float inv, float scale) { float q = ceilf((v - p) * inv); q = fminf(fmaxf(q, 0.f), 255.f); while (q < 255.f && p + q * scale < v) q += 1.f; return (uint32_t)q; }

// ---- refit (bvh_refit): gather the moved triangles into leaf order, then recompute the boxes level by level from the deepest level up
__global__ void k_regather(const DevTri* __restrict__ tris_in, const uint32_t* __restrict__ perm, uint32_t n, DevTri* __restrict__ tris_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tris_out[i] = tris_in[perm[i]];
}
__global__ void k_refit_level(Bvh8Node* __restrict__ nodes, uint32_t first, uint32_t count, const DevTri* __restrict__ tris, float4* __restrict__ box_lo, float4* __restrict__ box_hi) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= count) return;
    const uint32_t id = first + w;
    Bvh8Node nd = nodes[id];
    const uint32_t imask = nd.q0.w >> 24, child_base = nd.q1.x, tri_base = nd.q1.y;
    float3 clo[8], chi[8]; bool used[8];
    float3 lo = f3(FLT_MAX), hi = f3(-FLT_MAX);
    for (int s = 0; s < 8; ++s) {
        const uint32_t m = ((s < 4 ? nd.q1.z : nd.q1.w) >> (8 * (s & 3))) & 0xFFu;
        used[s] = m != 0u;
        if (!used[s]) continue;
        float3 l, h;
        if (imask & (1u << s)) { const uint32_t c = child_base + __popc(imask & ((1u << s) - 1u)); l = f3(box_lo[c]); h = f3(box_hi[c]); }
        else {
            const uint32_t k = __popc(m >> 5), off = m & 31u;
            l = f3(FLT_MAX); h = f3(-FLT_MAX);
            for (uint32_t t = 0; t < k; ++t) {
                const DevTri tr = tris[tri_base + off + t];
                const float3 a = f3(tr.v0), b = f3(tr.v1), c = f3(tr.v2);
                float3 tl = f3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
                float3 th = f3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
                const float3 p = f3(pad_of(tl.x, th.x), pad_of(tl.y, th.y), pad_of(tl.z, th.z));          // the padding of k_tri_bounds
                tl = tl - p; th = th + p;
                l = f3(fminf(l.x, tl.x), fminf(l.y, tl.y), fminf(l.z, tl.z)); h = f3(fmaxf(h.x, th.x), fmaxf(h.y, th.y), fmaxf(h.z, th.z));
            }
        }
        clo[s] = l; chi[s] = h;
        lo = f3(fminf(lo.x, l.x), fminf(lo.y, l.y), fminf(lo.z, l.z)); hi = f3(fmaxf(hi.x, h.x), fmaxf(hi.y, h.y), fmaxf(hi.z, h.z));
    }
    box_lo[id] = f4(lo, 0.f); box_hi[id] = f4(hi, 0.f);
    const uint32_t ex = quant_exponent(hi.x - lo.x), ey = quant_exponent(hi.y - lo.y), ez = quant_exponent(hi.z - lo.z);
    const float sx = __uint_as_float(ex << 23), sy = __uint_as_float(ey << 23), sz = __uint_as_float(ez << 23);
    const float ix = 1.0f / sx, iy = 1.0f / sy, iz = 1.0f / sz;
    uint32_t qlo[3][8], qhi[3][8];
    for (int s = 0; s < 8; ++s) {
        for (int a = 0; a < 3; ++a) { qlo[a][s] = 255u; qhi[a][s] = 0u; }
        if (!used[s]) continue;
        qlo[0][s] = quant_down(clo[s].x, lo.x, ix, sx); qlo[1][s] = quant_down(clo[s].y, lo.y, iy, sy); qlo[2][s] = quant_down(clo[s].z, lo.z, iz, sz);
        qhi[0][s] = quant_up(chi[s].x, lo.x, ix, sx); qhi[1][s] = quant_up(chi[s].y, lo.y, iy, sy); qhi[2][s] = quant_up(chi[s].z, lo.z, iz, sz);
    }
    auto pack4 = [](const uint32_t* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); };
    nd.q0 = make_uint4(__float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), ex | (ey << 8) | (ez << 16) | (imask << 24));
    nd.q2 = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
    nd.q3 = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
    nd.q4 = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
    nodes[id] = nd;
}

__global__ void k_collapse(const WorkItem* __restrict__ items, const uint32_t* __restrict__ n_items_in, WorkItem* __restrict__ next, uint32_t* __restrict__ next_count, uint32_t* __restrict__ counters /* 0: nodes, 1: tris */,
                           int n, const uint2* __restrict__ children, const uint32_t* __restrict__ count, const float4* __restrict__ nlo, const float4* __restrict__ nhi,
                           const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ ref_tri, const DevTri* __restrict__ tris_in, Bvh8Node* __restrict__ nodes, DevTri* __restrict__ tris_out,
                           uint32_t* __restrict__ perm_out) {
    // the level's item count lives on the device (written by the previous level's launch): a fixed grid strides over it
    const uint32_t n_items = *n_items_in;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_items; w += gridDim.x * blockDim.x) {
    const WorkItem item = items[w];
    const uint32_t first_leaf = (uint32_t)(n - 1);
    auto tri_count = [&](uint32_t node) -> uint32_t { return count[node]; };

    uint32_t cand[8]; int nc = 0;
    if (tri_count(item.bnode) <= kLeafMax) cand[nc++] = item.bnode;
    else { const uint2 ch = children[item.bnode]; cand[nc++] = ch.x; cand[nc++] = ch.y; }
    while (nc < 8) {                      // open the largest openable child until the node is full
        int best = -1; float best_area = -1.f;
        for (int c = 0; c < nc; ++c) {
            if (tri_count(cand[c]) <= kLeafMax) continue;
            const float a = half_area(nlo[cand[c]], nhi[cand[c]]);
            if (a > best_area) { best_area = a; best = c; }
        }
        if (best < 0) break;
        const uint2 ch = children[cand[best]];
        cand[best] = ch.x; cand[nc++] = ch.y;
    }

    const float4 plo = nlo[item.bnode], phi = nhi[item.bnode];
    const float3 pc = f3((plo.x + phi.x) * 0.5f, (plo.y + phi.y) * 0.5f, (plo.z + phi.z) * 0.5f);
    // octant-ordered slots: slot bit (4,2,1) set <=> child lies towards +x,+y,+z of the parent centre
    float cost[8][8];
    for (int c = 0; c < nc; ++c) {
        const float4 l = nlo[cand[c]], h = nhi[cand[c]];
        const float3 dc = f3((l.x + h.x) * 0.5f - pc.x, (l.y + h.y) * 0.5f - pc.y, (l.z + h.z) * 0.5f - pc.z);
        for (int s = 0; s < 8; ++s)
            cost[c][s] = ((s & 4) ? dc.x : -dc.x) + ((s & 2) ? dc.y : -dc.y) + ((s & 1) ? dc.z : -dc.z);
    }
    int slot_of[8]; int child_at[8];
    for (int s = 0; s < 8; ++s) child_at[s] = -1;
    for (int c = 0; c < nc; ++c) slot_of[c] = -1;
    for (int round = 0; round < nc; ++round) {
        float bestv = -FLT_MAX; int bc = -1, bs = -1;
        for (int c = 0; c < nc; ++c) {
            if (slot_of[c] >= 0) continue;
            for (int s = 0; s < 8; ++s) {
                if (child_at[s] >= 0) continue;
                if (cost[c][s] > bestv) { bestv = cost[c][s]; bc = c; bs = s; }
            }
        }
        slot_of[bc] = bs; child_at[bs] = bc;
    }

    uint32_t n_inner = 0, n_leaf_tris = 0;
    for (int c = 0; c < nc; ++c) { const uint32_t k = tri_count(cand[c]); if (k <= kLeafMax) n_leaf_tris += k; else ++n_inner; }
    const uint32_t child_base = n_inner ? atomicAdd(&counters[0], n_inner) : 0u;
    const uint32_t tri_base = n_leaf_tris ? atomicAdd(&counters[1], n_leaf_tris) : 0u;
    const uint32_t next_base = n_inner ? atomicAdd(next_count, n_inner) : 0u;

    const uint32_t ex = quant_exponent(phi.x - plo.x), ey = quant_exponent(phi.y - plo.y), ez = quant_exponent(phi.z - plo.z);
    const float sx = __uint_as_float(ex << 23), sy = __uint_as_float(ey << 23), sz = __uint_as_float(ez << 23);
    const float ix = 1.0f / sx, iy = 1.0f / sy, iz = 1.0f / sz;

    uint32_t imask = 0, rank = 0, tri_off = 0;
    uint32_t meta[8], qlo[3][8], qhi[3][8];
    for (int s = 0; s < 8; ++s) {
        meta[s] = 0; for (int a = 0; a < 3; ++a) { qlo[a][s] = 255u; qhi[a][s] = 0u; }
        const int c = child_at[s];
        if (c < 0) continue;
        const uint32_t node = cand[c];
        const float4 l = nlo[node], h = nhi[node];
        qlo[0][s] = quant_down(l.x, plo.x, ix, sx); qlo[1][s] = quant_down(l.y, plo.y, iy, sy); qlo[2][s] = quant_down(l.z, plo.z, iz, sz);
        qhi[0][s] = quant_up(h.x, plo.x, ix, sx); qhi[1][s] = quant_up(h.y, plo.y, iy, sy); qhi[2][s] = quant_up(h.z, plo.z, iz, sz);
        const uint32_t k = tri_count(node);
        if (k <= kLeafMax) {
            meta[s] = (((1u << k) - 1u) << 5) | tri_off;
            // the (at most kLeafMax) triangles below this binary node, left to right
            uint32_t walk[kLeafMax + 1]; int wp = 0; walk[wp++] = node;
            while (wp > 0) {
                const uint32_t b = walk[--wp];
                if (b >= first_leaf) { const uint32_t ref = sorted[b - first_leaf], src = ref_tri ? ref_tri[ref] : ref; perm_out[tri_base + tri_off] = src; tris_out[tri_base + tri_off++] = tris_in[src]; }
                else { const uint2 ch = children[b]; walk[wp++] = ch.y; walk[wp++] = ch.x; }
            }
        } else {
            meta[s] = (1u << 5) | (24u + (uint32_t)s);
            imask |= 1u << s;
            next[next_base + rank] = WorkItem{node, child_base + rank};
            ++rank;
        }
    }
    auto pack4 = [](const uint32_t* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); };
    Bvh8Node out;
    out.q0 = make_uint4(__float_as_uint(plo.x), __float_as_uint(plo.y), __float_as_uint(plo.z), ex | (ey << 8) | (ez << 16) | (imask << 24));
    out.q1 = make_uint4(child_base, tri_base, pack4(meta), pack4(meta + 4));
    out.q2 = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
    out.q3 = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
    out.q4 = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
    nodes[item.wnode] = out;
    }
}

} // namespace

void bvh_build(cudaStream_t s, const DevTri* tris_in, uint32_t n_tris, DeviceBvh& out, BvhBuilder builder, int ploc_radius, float split_fraction) {
    out.num_nodes = 0; out.num_tris = 0; out.src_tris = 0; out.levels = 0; out.build_ms = 0.f; out.ploc_rounds = 0;
    if (n_tris == 0) return;
    static const bool timing = getenv("LB_BVH_TIMING") != nullptr;
    auto tick = [&](const char* what) { if (!timing) return; cudaStreamSynchronize(s); static double last = 0; timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; fprintf(stderr, "[bvh %u] %-10s +%.3f ms\n", n_tris, what, last ? now - last : 0.0); last = now; };
    tick("enter");
    cudaEvent_t e0, e1, e_split; LB_CUDA(cudaEventCreate(&e0)); LB_CUDA(cudaEventCreate(&e1)); LB_CUDA(cudaEventCreate(&e_split));
    const int B = 256;
    const int h_bounds[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};

    // ---- references: one per triangle, or several for the triangles early split clipping cuts up (see k_split). Their number decides the
    //      size of everything below, so this stage runs first, with its own scratch.
    StreamBuf<float4> tlo, thi, rlo, rhi; StreamBuf<int> cbounds; StreamBuf<uint32_t> ref_tri, split_counts, split_offsets; StreamBuf<unsigned char> cub_tmp;
    tlo.reserve(n_tris, s); thi.reserve(n_tris, s); cbounds.reserve(6, s);
    LB_CUDA(cudaEventRecord(e_split, s));
    LB_CUDA(cudaMemcpyAsync(cbounds.p, h_bounds, sizeof h_bounds, cudaMemcpyHostToDevice, s));
    k_tri_bounds<<<grid_for(n_tris, B), B, 0, s>>>(tris_in, n_tris, tlo.p, thi.p, cbounds.p); LB_LAUNCH_CHECK();
    uint32_t n = n_tris; const float4 *lo_p = tlo.p, *hi_p = thi.p; const uint32_t* ref_tri_p = nullptr;
    out.split_cell = 0.f;
    if (split_fraction > 0.f && n_tris > 1u) {
        int hb[6];
        LB_CUDA(cudaMemcpyAsync(hb, cbounds.p, sizeof hb, cudaMemcpyDeviceToHost, s)); LB_CUDA(cudaStreamSynchronize(s));
        auto of = [](int i) { const int u = i >= 0 ? i : i ^ 0x7FFFFFFF; float f; memcpy(&f, &u, 4); return f; };
        const float extent = std::max(of(hb[3]) - of(hb[0]), std::max(of(hb[4]) - of(hb[1]), of(hb[5]) - of(hb[2])));
        if (extent > 0.f && std::isfinite(extent)) {
            split_counts.reserve(n_tris, s); split_offsets.reserve(n_tris, s);
            size_t scan_bytes = 0;
            LB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, split_counts.p, split_offsets.p, (int)n_tris, s));
            cub_tmp.reserve(scan_bytes, s);
            SplitParams sp{extent * split_fraction, 1u};
            uint32_t total = n_tris;
            for (int attempt = 0; attempt < 10; ++attempt, sp.cell *= 2.f) {         // the references may at most double the leaf arrays
                k_split<0><<<grid_for(n_tris, B), B, 0, s>>>(tris_in, n_tris, sp, nullptr, split_counts.p, nullptr, nullptr, nullptr); LB_LAUNCH_CHECK();
                LB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, scan_bytes, split_counts.p, split_offsets.p, (int)n_tris, s));
                uint32_t last_off = 0, last_cnt = 0;
                LB_CUDA(cudaMemcpyAsync(&last_off, split_offsets.p + (n_tris - 1u), 4, cudaMemcpyDeviceToHost, s));
                LB_CUDA(cudaMemcpyAsync(&last_cnt, split_counts.p + (n_tris - 1u), 4, cudaMemcpyDeviceToHost, s));
                LB_CUDA(cudaStreamSynchronize(s));
                total = last_off + last_cnt;
                if ((uint64_t)total <= 2ull * n_tris) break;
            }
            if (total > n_tris && (uint64_t)total <= 2ull * n_tris) {
                n = total; out.split_cell = sp.cell;
                rlo.reserve(n, s); rhi.reserve(n, s); ref_tri.reserve(n, s);
                k_split<1><<<grid_for(n_tris, B), B, 0, s>>>(tris_in, n_tris, sp, split_offsets.p, nullptr, rlo.p, rhi.p, ref_tri.p); LB_LAUNCH_CHECK();
                LB_CUDA(cudaMemcpyAsync(cbounds.p, h_bounds, sizeof h_bounds, cudaMemcpyHostToDevice, s));
                k_ref_bounds<<<grid_for(n, B), B, 0, s>>>(rlo.p, rhi.p, n, cbounds.p); LB_LAUNCH_CHECK();
                lo_p = rlo.p; hi_p = rhi.p; ref_tri_p = ref_tri.p;
            }
        }
    }
    tick("split");
    timespec a0, a1; clock_gettime(CLOCK_MONOTONIC, &a0);

    const uint32_t n_binary = 2u * n - 1u;
    StreamBuf<float4> nlo, nhi; StreamBuf<uint64_t> keys, keys_sorted; StreamBuf<uint32_t> vals, sorted, count, counters;
    StreamBuf<uint2> children; StreamBuf<WorkItem> items_a, items_b;
    nlo.reserve(n_binary, s); nhi.reserve(n_binary, s); count.reserve(n_binary, s);
    keys.reserve(n, s); keys_sorted.reserve(n, s); vals.reserve(n, s); sorted.reserve(n, s); counters.reserve(4, s);
    children.reserve(n, s); items_a.reserve(n, s); items_b.reserve(n, s);
    tick("scratch");
    out.nodes.reserve(n); out.tris.reserve(3 * (size_t)n);          // three axis-rotated copies of the leaf-ordered triangles (k_rotate_tris)
    out.perm.reserve(n); out.box_lo.reserve(n); out.box_hi.reserve(n); out.level_start.clear(); out.refits = 0;

    tick("alloc");
    // build_ms is the device time of the build proper — what a re-commit of the scene costs. The allocations above are paid by the first
    // build of a renderer only (afterwards the buffers and the stream-ordered pool hold the memory); they are reported apart: on a fresh
    // process they take 3 - 10 ms of cudaMalloc / pool growth for C2, more than the build itself.
    clock_gettime(CLOCK_MONOTONIC, &a1); out.alloc_ms = (float)((a1.tv_sec - a0.tv_sec) * 1e3 + (a1.tv_nsec - a0.tv_nsec) * 1e-6);
    LB_CUDA(cudaEventRecord(e0, s));
    k_morton<<<grid_for(n, B), B, 0, s>>>(lo_p, hi_p, n, cbounds.p, keys.p, vals.p); LB_LAUNCH_CHECK();
    size_t tmp_bytes = 0;
    LB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_sorted.p, vals.p, sorted.p, (int)n, 0, 63, s));
    cub_tmp.reserve(tmp_bytes, s);
    LB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, keys.p, keys_sorted.p, vals.p, sorted.p, (int)n, 0, 63, s));

    tick("sort");
    uint32_t root = 0u;                                     // binary node ids: internal [0, n-2], leaf j = (n-1) + j
    if (builder == BvhBuilder::LBVH || n == 1) {
        StreamBuf<uint32_t> parent, flags; StreamBuf<uint2> range;
        parent.reserve(n_binary, s); flags.reserve(n, s); range.reserve(n, s);
        flags.zero();
        if (n > 1) { k_hierarchy<<<grid_for(n - 1, B), B, 0, s>>>(keys_sorted.p, (int)n, children.p, parent.p, range.p); LB_LAUNCH_CHECK(); }
        k_refit<<<grid_for(n, B), B, 0, s>>>(sorted.p, lo_p, hi_p, (int)n, children.p, parent.p, nlo.p, nhi.p, count.p, flags.p); LB_LAUNCH_CHECK();
        LB_CUDA(cudaStreamSynchronize(s));                  // the temporaries above are released here
    } else {
        // PLOC: the cluster arrays ping-pong through a flag / exclusive-scan / scatter compaction every round
        StreamBuf<uint32_t> cl[2], nearest, keep, offset, m_dev; StreamBuf<float4> clo[2], chi[2];
        for (int k = 0; k < 2; ++k) { cl[k].reserve(n, s); clo[k].reserve(n, s); chi[k].reserve(n, s); }
        nearest.reserve(n, s); keep.reserve(n, s); offset.reserve(n, s); m_dev.reserve(4, s);
        size_t scan_bytes = 0;
        LB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, keep.p, offset.p, (int)n, s));
        cub_tmp.reserve(scan_bytes, s);
        // m_dev: [0] / [2] cluster count before / after the round (the two slots alternate), [1] next internal node id, [3] rounds that merged.
        // The rounds are launched in batches with grids sized for the count at the start of the batch; the kernels read the actual count
        // from the device, and a round launched after the last merge just carries the root across. One host round trip per BATCH: the
        // first version read the count back after every round (~55 rounds x 2 hierarchies x ~100 us: 12 of the 16 ms of a C2 build).
        const uint32_t h_m[4] = {n, 0u, 0u, 0u};
        LB_CUDA(cudaMemcpyAsync(m_dev.p, h_m, sizeof h_m, cudaMemcpyHostToDevice, s));
        k_ploc_init<<<grid_for(n, B), B, 0, s>>>(sorted.p, lo_p, hi_p, n, cl[0].p, clo[0].p, chi[0].p, nlo.p, nhi.p, count.p); LB_LAUNCH_CHECK();
        uint32_t m = n; int cur = 0; uint32_t slot = 0u;
        constexpr int kRoundsPerBatch = 8;
        while (m > 1u) {
            if (m <= (uint32_t)kPlocTail) {                     // the rest in one block
                uint32_t* m_io = m_dev.p + slot;
                if (ploc_radius >= 128) k_ploc_tail<128><<<1, kPlocTail, 0, s>>>(cl[cur].p, clo[cur].p, chi[cur].p, m_io, children.p, nlo.p, nhi.p, count.p, m_dev.p + 1, m_dev.p + 3);
                else if (ploc_radius >= 64) k_ploc_tail<64><<<1, kPlocTail, 0, s>>>(cl[cur].p, clo[cur].p, chi[cur].p, m_io, children.p, nlo.p, nhi.p, count.p, m_dev.p + 1, m_dev.p + 3);
                else k_ploc_tail<16><<<1, kPlocTail, 0, s>>>(cl[cur].p, clo[cur].p, chi[cur].p, m_io, children.p, nlo.p, nhi.p, count.p, m_dev.p + 1, m_dev.p + 3);
                LB_LAUNCH_CHECK();
            } else
            for (int b = 0; b < kRoundsPerBatch; ++b) {
                const uint32_t* m_in = m_dev.p + slot; uint32_t* m_out = m_dev.p + (slot ^ 2u);
                if (ploc_radius >= 128) k_ploc_nearest<128><<<grid_for(m, kPlocBlock), kPlocBlock, 0, s>>>(clo[cur].p, chi[cur].p, m_in, nearest.p);
                else if (ploc_radius >= 64) k_ploc_nearest<64><<<grid_for(m, kPlocBlock), kPlocBlock, 0, s>>>(clo[cur].p, chi[cur].p, m_in, nearest.p);
                else k_ploc_nearest<16><<<grid_for(m, kPlocBlock), kPlocBlock, 0, s>>>(clo[cur].p, chi[cur].p, m_in, nearest.p);
                LB_LAUNCH_CHECK();
                k_ploc_merge<<<grid_for(m, B), B, 0, s>>>(cl[cur].p, clo[cur].p, chi[cur].p, nearest.p, m_in, children.p, nlo.p, nhi.p, count.p, m_dev.p + 1, keep.p); LB_LAUNCH_CHECK();
                // scanned over the batch's starting count: entries past the round's own count are stale flags of an earlier round and do not
                // reach the exclusive prefixes in front of them
                LB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, scan_bytes, keep.p, offset.p, (int)m, s));
                k_ploc_compact<<<grid_for(m, B), B, 0, s>>>(cl[cur].p, clo[cur].p, chi[cur].p, keep.p, offset.p, m_in, cl[cur ^ 1].p, clo[cur ^ 1].p, chi[cur ^ 1].p, m_out, m_dev.p + 3); LB_LAUNCH_CHECK();
                cur ^= 1; slot ^= 2u;
            }
            uint32_t h_back[4];
            LB_CUDA(cudaMemcpyAsync(h_back, m_dev.p, sizeof h_back, cudaMemcpyDeviceToHost, s));
            LB_CUDA(cudaMemcpyAsync(&root, cl[cur].p, sizeof root, cudaMemcpyDeviceToHost, s));
            LB_CUDA(cudaStreamSynchronize(s));
            const uint32_t m_new = h_back[slot];
            if (m_new >= m || m_new == 0u) throw CudaError("bvh_build: PLOC made no progress");
            m = m_new; out.ploc_rounds = h_back[3];
        }
    }

    tick("binary");
    // level-synchronous collapse of the binary hierarchy into 8-wide nodes. level_items[L] = work items of level L, appended by the launch of
    // level L - 1; the launches go out in batches over a fixed grid and the counts are read back once per batch (a launch for a level
    // past the last one finds 0 items)
    const uint32_t h_counters[4] = {1u, 0u, 0u, 0u};
    LB_CUDA(cudaMemcpyAsync(counters.p, h_counters, sizeof h_counters, cudaMemcpyHostToDevice, s));
    const WorkItem root_item{n == 1 ? 0u : root, 0u};
    LB_CUDA(cudaMemcpyAsync(items_a.p, &root_item, sizeof root_item, cudaMemcpyHostToDevice, s));
    constexpr uint32_t kMaxLevels = 64, kLevelsPerBatch = 12;
    StreamBuf<uint32_t> level_items; level_items.reserve(kMaxLevels + 1u, s);
    uint32_t h_items[kMaxLevels + 1u] = {1u};
    LB_CUDA(cudaMemcpyAsync(level_items.p, h_items, sizeof h_items, cudaMemcpyHostToDevice, s));
    WorkItem* cur = items_a.p; WorkItem* nxt = items_b.p;
    uint32_t h_c[4] = {0u, 0u, 0u, 0u};
    int sms = 148; { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const uint32_t collapse_grid = (uint32_t)std::min<uint64_t>(grid_for(n, 128), (uint64_t)sms * 16u);
    uint32_t launched = 0;
    for (;;) {
        for (uint32_t b = 0; b < kLevelsPerBatch && launched < kMaxLevels; ++b, ++launched) {
            k_collapse<<<collapse_grid, 128, 0, s>>>(cur, level_items.p + launched, nxt, level_items.p + launched + 1u, counters.p, (int)n, children.p, count.p, nlo.p, nhi.p, sorted.p, ref_tri_p, tris_in,
                                                     out.nodes.p, out.tris.p, out.perm.p);
            LB_LAUNCH_CHECK();
            std::swap(cur, nxt);
        }
        LB_CUDA(cudaMemcpyAsync(h_items, level_items.p, sizeof h_items, cudaMemcpyDeviceToHost, s));
        LB_CUDA(cudaMemcpyAsync(h_c, counters.p, sizeof h_c, cudaMemcpyDeviceToHost, s));
        LB_CUDA(cudaStreamSynchronize(s));
        if (h_items[launched] == 0u) break;
        if (launched >= kMaxLevels) throw CudaError("bvh_build: hierarchy deeper than 64 levels");
    }
    uint32_t level_first = 0;                               // the nodes of a level are a contiguous range: [level_first, level_first + items)
    for (uint32_t L = 0; L < kMaxLevels && h_items[L] != 0u; ++L) { out.level_start.push_back(level_first); level_first += h_items[L]; ++out.levels; }
    out.level_start.push_back(level_first);
    out.num_nodes = h_c[0]; out.num_tris = h_c[1];
    tick("collapse");
    if (out.num_tris == n) { k_rotate_tris<<<grid_for(n, B), B, 0, s>>>(out.tris.p, n); LB_LAUNCH_CHECK(); }
    LB_CUDA(cudaEventRecord(e1, s)); LB_CUDA(cudaEventSynchronize(e1));
    LB_CUDA(cudaEventElapsedTime(&out.build_ms, e0, e1));
    out.src_tris = n_tris;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e_split);
    tick("finish");
    if (out.num_tris != n) throw CudaError("bvh_build: triangle count mismatch after collapse");
    if (level_first != out.num_nodes) throw CudaError("bvh_build: level ranges do not cover the nodes");
}

void bvh_refit(cudaStream_t s, const DevTri* tris_in, uint32_t n, DeviceBvh& bvh) {
    if (n == 0 || bvh.src_tris != n || bvh.level_start.size() != (size_t)bvh.levels + 1u) throw CudaError("bvh_refit: the hierarchy was not built from this many triangles");
    cudaEvent_t e0, e1; LB_CUDA(cudaEventCreate(&e0)); LB_CUDA(cudaEventCreate(&e1));
    LB_CUDA(cudaEventRecord(e0, s));
    const int B = 256;
    const uint32_t nl = bvh.num_tris;                       // leaf entries (>= n when triangles were split into references: a refitted reference
                                                            // gets its whole triangle's box — correct, looser than the clipped box of the build)
    k_regather<<<grid_for(nl, B), B, 0, s>>>(tris_in, bvh.perm.p, nl, bvh.tris.p); LB_LAUNCH_CHECK();
    k_rotate_tris<<<grid_for(nl, B), B, 0, s>>>(bvh.tris.p, nl); LB_LAUNCH_CHECK();
    for (int level = (int)bvh.levels - 1; level >= 0; --level) {
        const uint32_t first = bvh.level_start[level], count = bvh.level_start[level + 1] - first;
        if (!count) continue;
        k_refit_level<<<grid_for(count, 128), 128, 0, s>>>(bvh.nodes.p, first, count, bvh.tris.p, bvh.box_lo.p, bvh.box_hi.p); LB_LAUNCH_CHECK();
    }
    LB_CUDA(cudaEventRecord(e1, s)); LB_CUDA(cudaEventSynchronize(e1));
    LB_CUDA(cudaEventElapsedTime(&bvh.refit_ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ++bvh.refits;
}

} // namespace lb
