// lb_trace.cuh — ray/triangle intersection and traversal of the compressed 8-wide BVH.
//
// Replaces the reference's optixTrace calls (LumenPT/src/Shaders/WaveFrontShaders.cu:63-76 closest hit,
// :128-140 shadow any-hit, :197-210 ReSTIR visibility any-hit; no culling, OPTIX_RAY_FLAG_NONE :70) — B200 has no
// RT cores. Hit definition (DESIGN.md "extend"): world-space triangles, the watertight test of Woop/Benthin/Wald
// 2013 with the operation order below, accept tmin < t < tmax, closest = smallest t with ties towards the
// smaller (instance, primitive); barycentrics (u, v) weight vertices 1 and 2 (optixGetTriangleBarycentrics).
// Accepted set and t are a pure function of (ray, triangle), so any conservative BVH and any traversal order return
// the same hit — which is what lets the warp scheduling below (ray refill, postponed triangle tests) change freely.
//
// Structure: `Tracer` is the per-lane traversal state machine (one node test or one triangle test per step);
// `bvh8_trace` drives it for a single ray; `trace_queue` drives 32 of them per warp over a device-side ray queue with
//   * per-lane refill: a lane whose ray finished takes the next queue entry as soon as enough lanes are idle, instead of the
//     whole warp waiting for its slowest ray (persistent threads, Aila & Laine 2009);
//   * warp-voted rounds: every round is either a node round or a triangle round, chosen by ballot, so the lanes that work
//     execute the same code; triangle tests are postponed until a quarter of the live lanes have one pending (Ylitie,
//     Karras, Laine 2017, "postponing"), a lane with only triangle work left swaps in node work from its stack meanwhile.
#pragma once
#include "lb_device.cuh"

namespace lb {

// Division used by the traversal. Exact class by default (hit records are bit-compared with the oracle). A translation unit whose rays
// only decide visibility of ReSTIR samples (lb_restir.cu: DIRECT radiance, tolerance class) defines LB_TRACE_TOLERANCE_CLASS before
// including this header and gets its own unit's division (approximate under --use_fast_math), as it always had.
#ifdef LB_TRACE_TOLERANCE_CLASS
LB_D float tdiv(float a, float b) { return a / b; }
#else
LB_D float tdiv(float a, float b) { return xdiv(a, b); }
#endif

// Shear of the watertight test (Woop, Benthin, Wald 2013) along the ray's dominant axis kz. `rot` = (kz + 1) mod 3 selects the triangle
// copy whose stored coordinate order is (kx, ky, kz) = (kz + 1, kz + 2, kz) (lb_bvh.cu k_rotate_tris), so nothing is selected per triangle.
// The reference formulation swaps kx and ky when the direction's kz component is negative (to keep the winding); the swap exchanges
// (Ax, Ay), (Bx, By), (Cx, Cy), which negates U, V, W, det and T EXACTLY (a - b == -(b - a) in IEEE arithmetic, the double fallback
// included) and leaves the sign tests, t = T / det, u = V / det and v = W / det bit-identical — so it is simply not done here, and the
// oracle, which does it, still agrees bit for bit (tests/test_gpu_trace.py).
struct RayShear { int rot; float sx, sy, sz; };

LB_D float3 rotate_axes(const float3& v, int rot) {              // (v[rot], v[rot + 1], v[rot + 2]) with indices mod 3
    return rot == 0 ? v : (rot == 1 ? f3(v.y, v.z, v.x) : f3(v.z, v.x, v.y));
}
LB_D RayShear make_shear(const float3& d) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    const int kz = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    RayShear r;
    r.rot = kz == 2 ? 0 : kz + 1;
    const float3 dr = rotate_axes(d, r.rot);                     // (d[kx], d[ky], d[kz])
    r.sx = tdiv(dr.x, dr.z); r.sy = tdiv(dr.y, dr.z); r.sz = tdiv(1.0f, dr.z);
    return r;
}

// true + (t, u, v) when the ray's line hits the triangle; the range test is the caller's. `org` and the vertices are in the rotated frame.
LB_D bool tri_test(const float3& org, const RayShear& s, const float3& p0, const float3& p1, const float3& p2, float& t, float& u, float& v) {
    const float3 A = p0 - org, B = p1 - org, C = p2 - org;
    const float Ax = fmaf(-s.sx, A.z, A.x), Ay = fmaf(-s.sy, A.z, A.y);
    const float Bx = fmaf(-s.sx, B.z, B.x), By = fmaf(-s.sy, B.z, B.y);
    const float Cx = fmaf(-s.sx, C.z, C.x), Cy = fmaf(-s.sy, C.z, C.y);
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
    const float Az = s.sz * A.z, Bz = s.sz * B.z, Cz = s.sz * C.z;
    const float T = fmaf(U, Az, fmaf(V, Bz, W * Cz));
    t = tdiv(T, det); u = tdiv(V, det); v = tdiv(W, det);
    return true;
}

struct HitInfo { uint32_t inst, prim; float u, v, t; };

#ifdef LB_TRACE_STATS
// debug build: node visits / triangle tests / rays / warp rounds per translation unit (each unit that includes this header has its own copy;
// lb::dump_trace_stats_<unit> prints and clears it)
static __device__ unsigned long long g_trace_stats[8];
#define LB_TSTAT(k) (++tstat[k])
#define LB_TSTAT_RAY(tr) (++(tr).tstat[2])
#else
#define LB_TSTAT(k) ((void)0)
#define LB_TSTAT_RAY(tr) ((void)0)
#endif

constexpr int kTraceStack = 64;

LB_D float safe_rcp(float d) { return tdiv(1.0f, fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); }

// 2^23 + (byte J of `word`) as a float: one PRMT builds the bit pattern 0x4B0000bb, no integer-to-float conversion (I2F runs
// on the quarter-rate XU pipe and was the top stall of the first version of this kernel). The 2^23 bias is folded into the
// slab offsets by the caller.
template <int J> LB_D float byte_biased(uint32_t word, uint32_t k4b) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(k4b), "n"(0x7440 | J));
    return __uint_as_float(r);
}

// Per-lane traversal state. A "group" is {base, mask}: mask > 0x00FFFFFF = node group {first child node, hit-children
// bits 31..24 | imask 7..0}, otherwise a triangle group {first triangle, 24 pending bits} (Ylitie et al. 2017).
// The octant trick orders children front to back (slot ^ octant, highest bit first).
struct Tracer {
    float3 o, idir, orot; RayShear sh;      // orot: origin in the rotated frame of the triangle copy this ray reads
    uint32_t tri_off;                       // first triangle of that copy
    float tmin, best;                        // best = upper end of the search interval: tmax until a closer hit is found (never changes for any-hit rays)
    uint32_t octinv;
    uint2 cur;
    int sp;
    bool found; uint32_t bi, bp; float bu, bv;
    bool dropped;               // set by push() on overflow; owned by the driver (not reset per ray)
#ifdef LB_TRACE_STATS
    uint32_t tstat[4] = {0u, 0u, 0u, 0u};      // node steps, triangle tests, rays, triangle tests that passed the edge test
#endif
    uint2* stack;               // kTraceStack entries of thread-local memory owned by the driver (kept out of this struct so that the
                                // scalar state above is promoted to registers)

    LB_D void begin(const BvhView& bvh, const float3& o_, const float3& d, float tmin_, float tmax_) {
        o = o_; tmin = tmin_; best = tmax_;
        sh = make_shear(d); orot = rotate_axes(o_, sh.rot); tri_off = (uint32_t)sh.rot * bvh.num_tris;
        idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
        // octant and near/far swizzle come from the sign of the INVERSE direction, so that a component of -0.0 (inverse -1e20)
        // is treated consistently by both
        octinv = (idir.x < 0.f ? 0u : 4u) | (idir.y < 0.f ? 0u : 2u) | (idir.z < 0.f ? 0u : 1u);
        found = false; bi = 0u; bp = 0u; bu = 0.f; bv = 0.f; sp = 0;
        cur = (bvh.num_tris == 0u || !(tmax_ > tmin_)) ? make_uint2(0u, 0u) : make_uint2(0u, 0x80000000u);
    }
    LB_D bool is_node() const { return cur.y > 0x00FFFFFFu; }
    LB_D bool is_tri() const { return cur.y != 0u && cur.y <= 0x00FFFFFFu; }
    // A full stack drops the group — a possibly wrong hit — so it is never silent: `dropped` reaches the frame's CNT_STACK_OVERFLOW counter
    // (lb_frame_counters "stack_overflows", asserted 0 by the tests), and the scene commit refuses hierarchies deeper than the stack can
    // hold (one pending sibling group per level + one transient entry: levels + 2 <= kTraceStack, lb_api.cu commit_scene).
    LB_D void push(const uint2& g) { if (sp < kTraceStack) stack[sp++] = g; else dropped = true; }
    // makes `cur` non-empty from the stack; false = traversal finished
    LB_D bool refill_group() {
        if (cur.y != 0u) return true;
        if (sp == 0) return false;
        cur = stack[--sp];
        return true;
    }
    // a lane holding only triangle work takes node work from the top of its stack instead (the triangles go back on the stack)
    LB_D void swap_in_node() {
        if (sp > 0 && sp <= kTraceStack - 16 && stack[sp - 1].y > 0x00FFFFFFu) {      // never let postponing be what fills the stack
        const uint2 t = stack[sp - 1]; stack[sp - 1] = cur; cur = t; }
    }

    // visit the nearest un-visited hit child of the current node group: fetch its 80-byte node, slab-test the 8 children
    LB_D void node_step(const BvhView& bvh) {
        LB_TSTAT(0);
        const uint32_t hits_imask = cur.y;
        const uint32_t child_bit = 31u - (uint32_t)__clz(hits_imask);
        const uint32_t child_base = cur.x;
        cur.y &= ~(1u << child_bit);
        if (cur.y > 0x00FFFFFFu) push(cur);
        const uint32_t slot = (child_bit - 24u) ^ (octinv & 7u);
        const uint32_t rel = __popc(hits_imask & ~(0xFFFFFFFFu << slot));
        const uint4* np = reinterpret_cast<const uint4*>(bvh.nodes + (child_base + rel));
        const uint4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3), q4 = __ldg(np + 4);

        const bool neg_x = idir.x < 0.f, neg_y = idir.y < 0.f, neg_z = idir.z < 0.f;
        const uint32_t octinv4 = octinv * 0x01010101u;
        const float ax = __uint_as_float((q0.w & 0xFFu) << 23) * idir.x;
        const float ay = __uint_as_float(((q0.w >> 8) & 0xFFu) << 23) * idir.y;
        const float az = __uint_as_float(((q0.w >> 16) & 0xFFu) << 23) * idir.z;
        // Plane distance t = q*a + (origin - o)*idir with q a byte. The byte arrives as X = 2^23 + q (byte_biased), so the offset
        // carries -2^23*a: t = X*a + b, b = fma(-2^23, a, (origin - o)*idir). Rounding b costs at most |a|/2 (half a quantisation
        // step); the near offsets are lowered and the far offsets raised by a whole step |a|, which keeps the test conservative —
        // a box is never culled that the exact test would enter (hits are decided by the triangle test alone).
        const float bx = fmaf(-8388608.0f, ax, (__uint_as_float(q0.x) - o.x) * idir.x);
        const float by = fmaf(-8388608.0f, ay, (__uint_as_float(q0.y) - o.y) * idir.y);
        const float bz = fmaf(-8388608.0f, az, (__uint_as_float(q0.z) - o.z) * idir.z);
        const float bxn = bx - fabsf(ax), bxf = bx + fabsf(ax);
        const float byn = by - fabsf(ay), byf = by + fabsf(ay);
        const float bzn = bz - fabsf(az), bzf = bz + fabsf(az);
        // 0x4B000000 kept in a REGISTER (num_tris < 2^31, so the OR adds nothing — but ptxas cannot fold it): PRMT takes only one
        // immediate, and with the constant as the immediate every one of the 48 PRMTs needed its selector moved into a register first
        const uint32_t k4b = 0x4B000000u | (bvh.num_tris >> 31);
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
#define LB_SLAB(J)                                                                                                         \
            {                                                                                                              \
                const float t0x = fmaf(byte_biased<J>(nx, k4b), ax, bxn), t1x = fmaf(byte_biased<J>(fx, k4b), ax, bxf);    \
                const float t0y = fmaf(byte_biased<J>(ny, k4b), ay, byn), t1y = fmaf(byte_biased<J>(fy, k4b), ay, byf);    \
                const float t0z = fmaf(byte_biased<J>(nz, k4b), az, bzn), t1z = fmaf(byte_biased<J>(fz, k4b), az, bzf);    \
                const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));                                                 \
                const float tf = fminf(fminf(t1x, t1y), fminf(t1z, best));                                                 \
                if (tn <= tf) hitmask |= ((child_bits4 >> (8 * J)) & 0xFFu) << ((bit_index4 >> (8 * J)) & 0xFFu);          \
            }
            LB_SLAB(0) LB_SLAB(1) LB_SLAB(2) LB_SLAB(3)
#undef LB_SLAB
        }
        const uint2 ngroup = make_uint2(q1.x, (hitmask & 0xFF000000u) | (q0.w >> 24));
        const uint2 tgroup = make_uint2(q1.y, hitmask & 0x00FFFFFFu);
        // leaf children of this node are tested before descending further (they are at least as near as its inner children's content)
        if (tgroup.y != 0u) { if (ngroup.y > 0x00FFFFFFu) push(ngroup); cur = tgroup; }
        else cur = ngroup.y > 0x00FFFFFFu ? ngroup : make_uint2(0u, 0u);
    }

    // test ONE pending triangle of the current triangle group. ANY: true = occluded (stop).
    template <bool ANY>
    LB_D bool tri_step(const BvhView& bvh) {
        LB_TSTAT(1);
        const uint32_t k = (uint32_t)__ffs(cur.y) - 1u;
        cur.y &= cur.y - 1u;
        const float4* tp = reinterpret_cast<const float4*>(bvh.tris + (tri_off + cur.x + k));
        const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
        float t, u, v;
        if (!tri_test(orot, sh, f3(v0), f3(v1), f3(v2), t, u, v)) return false;
        LB_TSTAT(3);
        if (!(t > tmin)) return false;
        if (ANY) return t < best;
        const uint32_t ti = __float_as_uint(v0.w), tpi = __float_as_uint(v1.w);
        const bool better = found ? (t < best || (t == best && (ti < bi || (ti == bi && tpi < bp)))) : (t < best);
        if (better) { found = true; best = t; bi = ti; bp = tpi; bu = u; bv = v; }
        return false;
    }
    LB_D void result(HitInfo& hit) const { hit.inst = bi; hit.prim = bp; hit.u = bu; hit.v = bv; hit.t = best; }
};

// one ray, one thread (debug taps, unit tests)
template <bool ANY>
LB_D bool bvh8_trace(const BvhView& bvh, const float3& o, const float3& d, float tmin, float tmax, HitInfo& hit) {
    uint2 stack_mem[kTraceStack];
    Tracer tr; tr.stack = stack_mem; tr.dropped = false; tr.begin(bvh, o, d, tmin, tmax);
    while (tr.refill_group()) {
        if (tr.is_node()) tr.node_step(bvh);
        else if (tr.template tri_step<ANY>(bvh)) return true;
    }
    if (tr.dropped && bvh.overflow) atomicAdd(bvh.overflow, 1u);
    if (ANY) return false;
    if (tr.found) tr.result(hit);
    return tr.found;
}


// A warp works through queue entries [0, n) handed out by a device ticket.
//   job.load(i, o, d, tmin, tmax) -> bool : fetch entry i; false = the entry needs no ray (job.done is still called, hit = false)
//   job.done(i, hit, tracer)              : consume the result (ANY: hit = occluded; else hit = tracer.found)
//   Job::kDeferDone                       : consume results when the lane is refilled (jobs whose done() waits for memory) or at once
// All 32 lanes of the warp must call this together.
template <bool ANY, class Job>
LB_D void trace_queue(const BvhView& bvh, uint32_t n, uint32_t* ticket, Job& job, const TraceTuning tune) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint2 stack_mem[kTraceStack];
    Tracer tr; tr.stack = stack_mem; tr.cur = make_uint2(0u, 0u); tr.sp = 0; tr.dropped = false;
    uint32_t item = 0u; bool live = false, exhausted = false;
    // A finished ray's result is consumed (job.done: for shadow / visibility rays a dependent load + read-modify-write) when its lane is
    // refilled, together with the other finished lanes of the warp, not at the moment it finishes: the warp then waits for that memory
    // round trip once per refill batch instead of once per ray.
    bool fin = false, fin_hit = false;
    for (;;) {
        // ---- refill idle lanes from the queue
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !live);
        if (idle != 0u && !exhausted && (idle == 0xFFFFFFFFu || __popc(idle) >= tune.refill_min)) {
            if (!live) {
                if (fin) { job.done(item, fin_hit, tr); fin = false; }
                uint32_t base = 0u;
                const int leader = __ffs(idle) - 1;
                if ((int)lane == leader) base = atomicAdd(ticket, (uint32_t)__popc(idle));
                base = __shfl_sync(idle, base, leader);
                item = base + (uint32_t)__popc(idle & lt_mask);
                if (item < n) {
                    float3 o, d; float tmin, tmax;
                    if (job.load(item, o, d, tmin, tmax)) { tr.begin(bvh, o, d, tmin, tmax); live = true; LB_TSTAT_RAY(tr); }
                    else job.done(item, false, tr);
                }
            }
            exhausted = __any_sync(0xFFFFFFFFu, !live && item >= n);
            if (!__any_sync(0xFFFFFFFFu, live)) { if (exhausted) break; continue; }
        } else if (idle == 0xFFFFFFFFu) break;                  // nothing live and the queue is exhausted
        // ---- node phase: every lane holding a node group visits one child
        if (live && tr.is_node()) tr.node_step(bvh);
        // ---- triangle phase, voted: one triangle per lane per iteration, for as long as enough lanes have triangles pending (or
        //      nobody has node work). Lanes left holding triangles put them back and take node work from their stack.
        uint32_t tri_mask = __ballot_sync(0xFFFFFFFFu, live && tr.is_tri());
        if (tri_mask != 0u) {
            const uint32_t node_mask = __ballot_sync(0xFFFFFFFFu, live && tr.is_node());
            const int live_n = 32 - __popc(idle);
            while (tri_mask != 0u && (node_mask == 0u || __popc(tri_mask) * tune.tri_quarter >= live_n)) {
                if (live && tr.is_tri()) {
                    if (tr.template tri_step<ANY>(bvh)) { if (Job::kDeferDone) { fin = true; fin_hit = true; } else job.done(item, true, tr); live = false; }
                }
                tri_mask = __ballot_sync(0xFFFFFFFFu, live && tr.is_tri());
            }
            if (live && tr.is_tri()) tr.swap_in_node();
        }
        // ---- every live lane holds a non-empty group for the next round, or retires
        if (live && !tr.refill_group()) { if (Job::kDeferDone) { fin = true; fin_hit = ANY ? false : tr.found; } else job.done(item, ANY ? false : tr.found, tr); live = false; }
    }
    if (fin) job.done(item, fin_hit, tr);
    if (tr.dropped && bvh.overflow) atomicAdd(bvh.overflow, 1u);
#ifdef LB_TRACE_STATS
    for (int k = 0; k < 4; ++k) { const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, tr.tstat[k]); if (lane == 0u && v) atomicAdd(&g_trace_stats[(ANY ? 4 : 0) + k], (unsigned long long)v); }
#endif
}

} // namespace lb
