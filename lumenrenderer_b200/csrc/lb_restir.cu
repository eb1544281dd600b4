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
#define LB_TRACE_TOLERANCE_CLASS          // visibility rays: occlusion of a reservoir sample, never a bit-compared hit record
#include "lb_trace.cuh"
#include "lb_shade.cuh"
#include <cfloat>
#include <cuda.h>                     // CUtensorMap (type only: the driver entry point is resolved at run time, lb_api.cu)

namespace lb {

namespace {

constexpr int kBlock = 256;
#ifndef LB_GATHER_BLOCKS
#define LB_GATHER_BLOCKS 2        // resident blocks per SM the gather kernels (temporal / spatial / combine) are compiled for
#endif
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
    const BsdfCtx ctx = surface_ctx(px);
    Reservoir out = reservoir_zero(); int total = 0;
#pragma unroll 1                                   // ONE inlined BSDF evaluation: the instruction footprint, not the trip count, is what costs here
    for (int k = 0; k < 2; ++k) {
        const Reservoir& q = k ? b : a;
        LightSample rs; resample(q.s, px.pos, px.normal, ctx, rs);
        reservoir_update(out, rs, (float)q.count * q.weight * rs.pdf, seed); total += q.count;
    }
    out.count = total; reservoir_update_weight(out);
    return out;
}

// CombineUnbiased over two reservoirs (ReSTIRKernels.cu:1123-1198; LbSettings::restir_unbiased — dead in the reference's shipped build): the
// selected sample is re-evaluated at the pixel each reservoir was generated at, and only reservoirs for whose pixel it has a non-zero target
// pdf count towards the normalisation
LB_D Reservoir combine_pair_unbiased(const Reservoir& a, const Reservoir& b, const Surface& px, const Surface& from_a, const Surface& from_b, uint32_t seed) {
    Reservoir out = reservoir_zero(); int total = 0;
    {
        const BsdfCtx ctx = surface_ctx(px);
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
            const Reservoir& q = k ? b : a;
            LightSample rs; resample(q.s, px.pos, px.normal, ctx, rs);
            reservoir_update(out, rs, (float)q.count * q.weight * rs.pdf, seed); total += q.count;
        }
    }
    out.count = total;
    int correction = 0;
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        const Surface& f = k ? from_b : from_a;
        const BsdfCtx ctx = surface_ctx(f);
        LightSample rs; resample(out.s, f.pos, f.normal, ctx, rs);
        if (rs.pdf > 0) correction += k ? b.count : a.count;
    }
    const float m = 1.f / fmaxf((float)correction, FLT_EPSILON);
    out.weight = (1.f / fmaxf(out.s.pdf, FLT_EPSILON)) * (m * out.weight_sum);
    return out;
}

// Pixel order of the gather kernels (temporal / spatial reuse): the image is cut into 32x8-pixel tiles walked in vertical strips
// 16 tiles (512 px) wide; the unit of work is one 32-pixel ROW of a tile, handed out to WARPS by a device ticket in tile order (row r of
// tile t is item 8 t + r). The warps in flight therefore always hold consecutive rows of a few consecutive tiles, whatever the grid /
// occupancy — a compact 2-D region whose +-30-pixel neighbour halo is mostly shared, instead of ~60 full image rows — so the gathers stay
// in L2 / L1, and no warp ever waits for another one (the first version handed whole tiles to blocks: two __syncthreads per tile and the
// block's slowest warp cost 0.9 stalled warps per issue, profiles/r01_u_kernels.md). A warp's own loads and stores are 512-byte coalesced.
constexpr uint32_t kTileW = 32, kTileH = 8, kStripTiles = 16;
struct TileWalk {
    uint32_t tiles_x, tiles_y, nitems, full;
    LB_D explicit TileWalk(const FrameView& fv) {
        tiles_x = (fv.width + kTileW - 1u) / kTileW; tiles_y = (fv.height + kTileH - 1u) / kTileH;
        nitems = tiles_x * tiles_y * kTileH; full = kStripTiles * tiles_y;
    }
    // pixel of this lane in row item `it`; false when it falls outside the image
    LB_D bool pixel(const FrameView& fv, uint32_t it, int& x, int& y) const {
        const uint32_t t = it / kTileH, row = it - t * kTileH;
        const uint32_t strip = t / full, r = t - strip * full;
        const uint32_t sw = min(kStripTiles, tiles_x - strip * kStripTiles);
        const uint32_t ty = r / sw, tx = strip * kStripTiles + (r - ty * sw);
        x = (int)(tx * kTileW + (threadIdx.x & 31u)); y = (int)(ty * kTileH + row);
        return (uint32_t)x < fv.width && (uint32_t)y < fv.height;
    }
    // next row item of this warp (warp-uniform); >= nitems when the image is exhausted
    LB_D uint32_t next(uint32_t* ticket) const {
        uint32_t it = 0u;
        if ((threadIdx.x & 31u) == 0u) it = atomicAdd(ticket, 1u);
        return __shfl_sync(0xFFFFFFFFu, it, 0);
    }
};

__global__ void __launch_bounds__(kBlock) k_fill_bags(SceneView sc, uint2* __restrict__ bags, uint32_t a_seed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kNumBags * kLightsPerBag) return;
    uint32_t s = wang_hash(a_seed + wang_hash(i));
    const float r = rand_f(s);
    uint32_t li; float pdf; cdf_get(sc, r, li, pdf);
    bags[i] = make_uint2(li, __float_as_uint(pdf));
}

// RIS over 32 bag candidates (the most expensive kernel of the reference's frame: 32 full BSDF evaluations per pixel).
//  * The pixel's BSDF context is built once (BsdfCtx).
//  * Phase A walks the 32 candidates with the whole warp converged and only decides which of them pass the geometric test
//    (light above the pixel's horizon and facing it). A candidate that fails contributes weight +0: it changes nothing but
//    the reservoir's sample count, so it needs no ordering and is accounted for by a popcount.
//  * Phase B evaluates the survivors in candidate order, one per lane per round: the lanes re-align on the expensive BSDF
//    evaluation, which the warp executes max-over-lanes(survivors) times instead of 32. Phase A leaves the xorshift state in
//    front of every candidate in shared memory, so every random number is the one the sequential loop of the reference
//    would have drawn.
//  * Between the phases the pixels are regrouped by survivor count, so that the 32 lanes of a warp run (nearly) the same number of rounds
//    (see k_ris below).
//  * All 256 pixels of a group share one light bag (canonical choice of hazard 1: bag = WangHash(seed + pixel / 256)). The groups are
//    counting-sorted by bag (k_ris_order); a block stages ONE bag — 1000 entries, each the full light record + its bag pdf, 88 KB — in
//    shared memory and works through that bag's groups. A candidate fetch is then 4-5 shared-memory reads instead of a dependent chain of random
//    global gathers (bag entry -> 64-byte light record): ncu showed the first version bound by the L1 data pipe
//    (l1tex__data_pipe_lsu_wavefronts 77 % of peak, ~8.5 wavefronts per request, profiles/r01_m_frame.md), not by instruction issue.
struct BagSmem {
    // phase A (conservative accept / reject of a candidate, 2 reads): bounding sphere of the light triangle, its plane
    float4* sph;    // centre.xyz, radius (+inf: "never reject": entries whose bag pdf is 0 or NaN take the ordered path)
    float4* pln;    // normal.xyz, min over the vertices of dot(normal, vertex)
    // phase B (the survivors: exact geometry, radiance)
    float4* g0;     // p0.xyz, p1.x
    float4* g1;     // p1.yz, p2.xy
    float4* g2;     // p2.z, radiance.xyz
    float2* pa;     // bag pdf, area
};
constexpr int kRisBlock = 512;                                  // ONE block per SM: 16 warps share one staged bag
constexpr int kRisRows = 2;                                     // 32-pixel rows per warp and iteration
constexpr int kRisSlots = kRisBlock * kRisRows;                 // pixels a block holds between phase A and phase B
#ifndef LB_RIS_GROUP_WARPS
#define LB_RIS_GROUP_WARPS 4
#endif
constexpr int kRisGroupWarps = LB_RIS_GROUP_WARPS;              // warps that sort their pixels together (a named barrier each: the groups of a block drift apart,
constexpr int kRisGroups = kRisBlock / 32 / kRisGroupWarps;     // so that one is in the shared-memory-bound phase A while another is in the MUFU-bound phase B)
constexpr int kRisGroupThreads = kRisGroupWarps * 32, kRisGroupSlots = kRisGroupThreads * kRisRows, kRisGroupRows = kRisGroupWarps * kRisRows;
static_assert(kRisGroups * kRisGroupWarps * 32 == kRisBlock && (kRisGroups <= 15 || kRisGroupWarps == 1), "groups tile the block; one hardware barrier each");
constexpr size_t kRisBagBytes = (size_t)kLightsPerBag * (5 * sizeof(float4) + sizeof(float2));
constexpr size_t kRisStateBytes = (size_t)kPrimarySamples * kRisSlots * sizeof(uint32_t);
// bag | candidate states [32][slots] | survivor mask [slots] | pixel [slots] | slots in ascending order of survivors (u16)
constexpr size_t kRisSmemBytes = kRisBagBytes + kRisStateBytes + (size_t)kRisSlots * (2 * sizeof(uint32_t) + sizeof(uint16_t));
static_assert(kRisSmemBytes + 1024u <= 227u * 1024u, "bag + candidate states + sort arrays must fit the 227 KB a block can have");

struct BagCandidate { LightSample ls; float bag_pdf; };
// geometry of the next candidate of stream `s` (position on the light, normal, area, bag pdf) + its bag slot and radiance
LB_D uint32_t draw_candidate_geom(const BagSmem& b, uint32_t& s, BagCandidate& c) {
    const float r = rand_f(s);
    const uint32_t slot = (uint32_t)(int)roundf((float)(kLightsPerBag - 1u) * r);
    const float4 a = b.g0[slot], bb = b.g1[slot], cc = b.g2[slot], pl = b.pln[slot]; const float2 pa = b.pa[slot];
    const float u = rand_f(s), v = rand_f(s) * (1.f - u);
    const float3 p0 = f3(a.x, a.y, a.z), p1 = f3(a.w, bb.x, bb.y), p2 = f3(bb.z, bb.w, cc.x);
    c.ls.normal = f3(pl); c.ls.area = pa.y; c.ls.contribution = f3(0.f); c.ls.pdf = 0.f;
    c.ls.radiance = f3(cc.y, cc.z, cc.w);
    c.ls.position = p0 + ((p1 - p0) * u) + ((p2 - p0) * v);
    c.bag_pdf = pa.x;
    return slot;
}
// Phase A's test: can the candidate at `slot` possibly pass Resample's geometric test (light above the pixel's horizon, facing it)? Decided
// for the WHOLE light triangle from its bounding sphere and plane, with margins 100x the rounding error of the exact test — a superset of
// what phase B accepts; a false survivor merely takes the ordered path with weight 0. Any point x of the triangle has
//   (x - p) . N  <=  (c - p) . N + r |N|      and      n . (p - x)  <=  n . p - min_vertices(n . v)   (n: the light's stored normal),
// the two quantities whose signs Resample tests (cos_in, cos_out, ReSTIRKernels.cu:1270-1281).
LB_D bool candidate_may_pass(const BagSmem& b, uint32_t slot, const float3& ppos, const float3& pnormal) {
    const float4 S = b.sph[slot], P = b.pln[slot];
    const float3 d = f3(S) - ppos;
    const float s_in = dot(d, pnormal) + S.w * 1.001f, eps_in = 1e-5f * (fabsf(d.x) + fabsf(d.y) + fabsf(d.z) + S.w);
    const float dn = dot(f3(P), ppos), s_out = dn - P.w, eps_out = 1e-5f * (fabsf(dn) + fabsf(P.w) + S.w);
    return !(s_in < -eps_in || s_out < -eps_out);                // NaN / inf anywhere: not rejected
}

extern __shared__ __align__(16) unsigned char ris_smem[];
#ifdef LB_RIS_STATS
__device__ unsigned long long g_ris_stats[8];
#endif

// canonical bag of a 256-pixel group (hazard 1): WangHash(seed + full-frame group index) -> one of the 50 bags
LB_D uint32_t bag_of_group(uint32_t seed, uint32_t group_base_pixel) {
    uint32_t bag_seed = wang_hash(seed + group_base_pixel / 256u);
    return (uint32_t)(int)roundf((float)(kNumBags - 1u) * rand_f(bag_seed));
}

// Pixel groups ordered by their bag (counting sort in one block; the order inside a bag is irrelevant — every pixel's result depends on
// its own group's bag only). Layout of `order`: [0, ngroups) = {group, bag} sorted by bag; [ngroups + b] = {first entry, entry count} of
// bag b; [ngroups + 64 + b].x = the bag's work ticket (zeroed here, every frame).
__global__ void __launch_bounds__(1024) k_ris_order(uint32_t seed, uint32_t npix, uint32_t pix0, uint2* __restrict__ order) {
    __shared__ uint32_t s_count[kNumBags], s_start[kNumBags];
    const uint32_t ngroups = (npix + 255u) / 256u;
    if (threadIdx.x < kNumBags) s_count[threadIdx.x] = 0u;
    __syncthreads();
    for (uint32_t g = threadIdx.x; g < ngroups; g += blockDim.x) atomicAdd(&s_count[bag_of_group(seed, g * 256u + pix0)], 1u);
    __syncthreads();
    if (threadIdx.x == 0u) {
        uint32_t run = 0u;
        for (uint32_t b = 0; b < kNumBags; ++b) { s_start[b] = run; order[ngroups + b] = make_uint2(run, s_count[b]); order[ngroups + 64u + b] = make_uint2(0u, 0u); run += s_count[b]; }
    }
    __syncthreads();
    for (uint32_t g = threadIdx.x; g < ngroups; g += blockDim.x) {
        const uint32_t bag = bag_of_group(seed, g * 256u + pix0);
        order[atomicAdd(&s_start[bag], 1u)] = make_uint2(g, bag);
    }
}

// Phase B of k_ris for one 32-pixel row: the survivors of every lane in candidate order, one per lane per round, the lanes aligned on the
// BSDF evaluation. MODE is voted by the warp: 2 = every pixel of the row has a material without transmission / sheen / clear coat / anisotropy
// / subsurface (lean evaluation), 1 = every pixel has isotropic roughness (the general evaluation without the anisotropic microfacet terms, which
// the compiler otherwise executes predicated-off), 0 = anything.
template <int MODE>
LB_D void ris_phase_b(const BagSmem& bag, const uint32_t (*s_state)[kRisSlots], uint32_t slot, const BsdfCtx& ctx, const Surface& px, uint32_t mask, uint32_t s0, Reservoir& fresh) {
    uint32_t sb = s0;
#ifdef LB_RIS_STATS
    if ((threadIdx.x & 31u) == 0u) atomicAdd(&g_ris_stats[5], 1ull);
    atomicAdd(&g_ris_stats[0], (unsigned long long)__popc(mask));
    { uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, (uint32_t)__popc(mask)); if ((threadIdx.x & 31u) == 0u) atomicAdd(&g_ris_stats[1], (unsigned long long)mx); }
#endif
    // The reservoir's Update replaces the whole 14-float sample when a candidate is accepted; here the loop only remembers WHICH candidate was
    // accepted last (+ the two things the evaluation produced: its pdf and contribution) and the sample's geometry is drawn again, once, after
    // the loop, from the candidate's saved xorshift state — the same numbers, 9 selects per round and 9 live registers less.
    int ksel = -1; float sel_pdf = 0.f; float3 sel_contribution = f3(0.f);
    while (__any_sync(0xFFFFFFFFu, mask != 0u)) {
        const bool active = mask != 0u;
        uint32_t kcur = 0u;
        BagCandidate c; ResampleGeom g; bool have_g = false;
        c.ls.radiance = f3(0.f); c.ls.normal = f3(0.f); c.ls.position = f3(0.f); c.ls.contribution = f3(0.f); c.ls.area = 0.f; c.ls.pdf = 0.f; c.bag_pdf = 1.f;
        g.dir = f3(0.f); g.solid = 0.f; g.cos_in = 0.f;
        if (active) {
            const uint32_t k = (uint32_t)__ffs(mask) - 1u; mask &= mask - 1u;
            kcur = k;
            sb = s_state[k][slot];
            draw_candidate_geom(bag, sb, c);
            have_g = resample_geom(c.ls.position, c.ls.normal, c.ls.area, px.pos, px.normal, g);
        }
        __syncwarp();
#ifdef LB_RIS_STATS
        if (have_g) atomicAdd(&g_ris_stats[2], 1ull);
        if (__any_sync(0xFFFFFFFFu, have_g) && (threadIdx.x & 31u) == 0u) atomicAdd(&g_ris_stats[3], 1ull);
        if ((threadIdx.x & 31u) == 0u) atomicAdd(&g_ris_stats[4], 1ull);
#endif
        if (have_g) resample_shade<MODE>(ctx, g, c.ls);
        if (active) {                                           // Reservoir::Update (ReSTIRData.h:122-141), the sample kept by index
            const float w = (have_g ? c.ls.pdf : 0.f) / c.bag_pdf;
            fresh.weight_sum += w; ++fresh.count;
            uint32_t su = sb;                                   // seed by value (hazard 14)
            if (rand_f(su) <= (w / fresh.weight_sum)) { ksel = (int)kcur; sel_pdf = c.ls.pdf; sel_contribution = c.ls.contribution; }
        }
        __syncwarp();
    }
    if (ksel >= 0) {
        BagCandidate c; uint32_t ss = s_state[ksel][slot];
        draw_candidate_geom(bag, ss, c);
        fresh.s = c.ls; fresh.s.pdf = sel_pdf; fresh.s.contribution = sel_contribution;
    }
}

// Work item = one 32-pixel row of a 256-pixel group, handed out by the work ticket of ONE bag: a block stages a bag (1000 entries with the
// full light record, 88 KB of shared memory) and works through that bag's rows 32 at a time — two per warp. Blocks start on bag
// (blockIdx mod 50) and, when their bag is exhausted, move on to the next bag that still has rows (work stealing; costs one more staging).
//
// Phase B costs one BSDF evaluation per lane and round, and a warp runs max-over-lanes(survivors) rounds. The survivor count of a pixel is
// close to Binomial(32, 1/2): measured on C2 15.8 survivors per pixel but 21.95 rounds per row — 28 % of the lanes of phase B idle
// (profiles/r02_n_ris_sorted.md). So the block does phase A for all its 1024 pixels first, counting-sorts them by survivor count in
// shared memory (33 bins), and hands phase B warps of 32 pixels with (nearly) EQUAL counts: warp w takes group w of the ascending order
// and then group 31 - w, so that every warp has about the same number of rounds in total and the barrier of the next iteration does not
// wait for the warp that drew the long groups. A pixel's result depends on its own candidates only, so the regrouping changes nothing.
__global__ void __launch_bounds__(kRisBlock, 1) k_ris(FrameView fv, SceneView sc, const uint2* __restrict__ bags, uint2* __restrict__ order, uint32_t seed, int allow_simple) {
    static_assert(kPrimarySamples == 32u && kRisRows % 2 == 0, "the survivor mask is one 32-bit word; a warp takes pairs of sets (a short one, a long one)");
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    const size_t np = fv.npix;
    BagSmem bag;
    bag.sph = reinterpret_cast<float4*>(ris_smem); bag.pln = bag.sph + kLightsPerBag; bag.g0 = bag.pln + kLightsPerBag; bag.g1 = bag.g0 + kLightsPerBag; bag.g2 = bag.g1 + kLightsPerBag;
    bag.pa = reinterpret_cast<float2*>(bag.g2 + kLightsPerBag);
    // xorshift state in front of every candidate, [candidate][slot]: phase B picks a survivor's stream up here instead of replaying
    // the draws of the candidates it skips (that replay loop, divergent by nature, was 12 % of the kernel's instructions)
    uint32_t (*s_state)[kRisSlots] = reinterpret_cast<uint32_t (*)[kRisSlots]>(ris_smem + kRisBagBytes);
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(ris_smem + kRisBagBytes + kRisStateBytes);
    uint32_t* s_pix = s_mask + kRisSlots;
    uint16_t* s_sorted = reinterpret_cast<uint16_t*>(s_pix + kRisSlots);
    __shared__ uint32_t s_next, s_row0[kRisGroups], s_bin[kRisGroups][kPrimarySamples + 2u];   // bin 0: slots without a pixel; bin 1 + c: pixels with c survivors
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t grp = warp / (uint32_t)kRisGroupWarps, gwarp = warp % (uint32_t)kRisGroupWarps, gtid = threadIdx.x % (uint32_t)kRisGroupThreads;
    const uint32_t slot_base = grp * (uint32_t)kRisGroupSlots;
    auto group_sync = [grp]() { if (kRisGroupWarps == 1) __syncwarp(); else asm volatile("bar.sync %0, %1;" ::"r"(1u + grp), "n"(kRisGroupThreads) : "memory"); };
    const uint32_t ngroups = (fv.npix + 255u) / 256u;
    const uint2* meta = order + ngroups;
    uint32_t* tickets = reinterpret_cast<uint32_t*>(order + ngroups + 64u);        // .x of entry b (stride 2 words)
    uint32_t cur = blockIdx.x % kNumBags;                       // block-uniform
    for (uint32_t visited = 0u; ; ) {
        // ---- next bag with rows left, starting at `cur` (warp 0 looks at all 50 tickets at once)
        __syncthreads();                                        // nobody reads the staged bag any more
        if (threadIdx.x < 32u) {
            bool open[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t b = lane + 32u * (uint32_t)h;
                open[h] = b < kNumBags && *reinterpret_cast<volatile uint32_t*>(&tickets[2u * b]) < meta[b].y * 8u;
            }
            const unsigned long long m = (unsigned long long)__ballot_sync(0xFFFFFFFFu, open[0]) | ((unsigned long long)__ballot_sync(0xFFFFFFFFu, open[1]) << 32);
            // first open bag at or after cur, cyclically
            const unsigned long long hi = m >> cur, lo = m & ((1ull << cur) - 1ull);
            const uint32_t nxt = hi ? cur + (uint32_t)__ffsll((long long)hi) - 1u : (lo ? (uint32_t)__ffsll((long long)lo) - 1u : kNone);
            if (lane == 0u) s_next = nxt;
        }
        __syncthreads();
        cur = s_next;
        if (cur == kNone || ++visited > 2u * kNumBags) break;
        const uint2* picked = bags + (size_t)cur * kLightsPerBag;
        for (uint32_t e = threadIdx.x; e < kLightsPerBag; e += kRisBlock) {
            const uint2 be = __ldg(&picked[e]);
            const float4* lp = reinterpret_cast<const float4*>(sc.lights + be.x);
            const float4 a = __ldg(lp), b = __ldg(lp + 1), c = __ldg(lp + 2), d = __ldg(lp + 3);
            const float3 p0 = f3(a.x, a.y, a.z), p1 = f3(a.w, b.x, b.y), p2 = f3(b.z, b.w, c.x), nrm = f3(c.y, c.z, c.w);
            const float3 ctr = (p0 + p1 + p2) * (1.f / 3.f);
            float rad = fmaxf(length(p0 - ctr), fmaxf(length(p1 - ctr), length(p2 - ctr))) * 1.0001f + 1e-6f;
            const float bag_pdf = __uint_as_float(be.y);
            // the light's normal is the transformed mean VERTEX normal (GPUDataBufferKernels.cu:150-156), not the triangle's own: n . x is not constant
            // over the triangle, so the plane offset is its minimum over the three vertices (n . x is linear in x)
            float4 plane = f4(nrm, fminf(dot(nrm, p0), fminf(dot(nrm, p1), dot(nrm, p2))));
            if (!(bag_pdf != 0.f && bag_pdf == bag_pdf)) { rad = __int_as_float(0x7f800000); plane.w = -__int_as_float(0x7f800000); }      // never rejected: ordered path
            bag.sph[e] = f4(ctr, rad); bag.pln[e] = plane;
            bag.g0[e] = a; bag.g1[e] = b; bag.g2[e] = make_float4(c.x, d.x, d.y, d.z); bag.pa[e] = make_float2(bag_pdf, d.w);
        }
        __syncthreads();                                        // the bag is staged
        const uint2 range = meta[cur];
        // ---- the bag's rows: every group of warps takes kRisGroupRows at a time, on its own
        for (;;) {
            group_sync();                                       // phase B of the group's previous iteration is over
            if (gtid == 0u) s_row0[grp] = atomicAdd(&tickets[2u * cur], (uint32_t)kRisGroupRows);
            if (gtid < kPrimarySamples + 2u) s_bin[grp][gtid] = 0u;
            group_sync();
            const uint32_t row0 = s_row0[grp];
            if (row0 >= range.y * 8u) break;                    // group-uniform
            // ---- phase A: which candidates of which pixel need the ordered path
            uint32_t key[kRisRows];
#pragma unroll
            for (uint32_t r = 0; r < (uint32_t)kRisRows; ++r) {
                const uint32_t it = row0 + gwarp * (uint32_t)kRisRows + r, slot = slot_base + r * (uint32_t)kRisGroupThreads + gtid;
                uint32_t i = kNone;
                if (it < range.y * 8u) {
                    i = __ldg(&order[range.x + (it >> 3)]).x * 256u + (it & 7u) * 32u + lane;
                    if (i >= fv.npix) i = kNone;
                }
                float3 ppos = f3(0.f), pnormal = f3(0.f);
                if (i != kNone) {
                    const Float8 s01 = ld2(fv.surf_cur + surf_pair(np, 0, i));      // position | flags, normal | depth
                    if (__float_as_uint(s01.a.w)) { reservoir_store(fv.res_cur, np, i, reservoir_zero()); i = kNone; }
                    else {
                        ppos = f3(s01.a); pnormal = f3(s01.b);
                        // the rest of the pixel's shading record is wanted in phase B, by whichever lane the sort gives the pixel to: ask L2 for
                        // it now (coalesced, no registers) — k_shade wrote it a millisecond ago and 500 MB of other planes have passed through L2 since
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(fv.surf_cur + surf_pair(np, 1, i)));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(fv.surf_cur + surf_pair(np, 2, i)));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(fv.surf_cur + surf_pair(np, 3, i)));
                    }
                }
                uint32_t mask = 0u;
                if (i != kNone) {
                    uint32_t sa = wang_hash(seed + wang_hash(i + fv.pix0));
#pragma unroll 2
                    for (uint32_t k = 0; k < kPrimarySamples; ++k) {
                        s_state[k][slot] = sa;
                        const float rr = rand_f(sa);
                        const uint32_t cslot = (uint32_t)(int)roundf((float)(kLightsPerBag - 1u) * rr);
                        rand_u32(sa); rand_u32(sa);                  // the candidate's u and v: only the stream position matters here
                        // ordered unless the update is provably a pure count increment: the light cannot pass the geometric test from anywhere on it
                        // (weight 0 / bag pdf, a bag pdf of 0 or NaN never lands here) and the acceptance draw is not the all-zero xorshift state
                        const bool ordered = candidate_may_pass(bag, cslot, ppos, pnormal) || sa == 0u;
                        mask |= (ordered ? 1u : 0u) << k;
                    }
                }
                s_mask[slot] = mask; s_pix[slot] = i;
                key[r] = i == kNone ? 0u : 1u + (uint32_t)__popc(mask);
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key[r]);
                if (lane == (uint32_t)__ffs(peers) - 1u) atomicAdd(&s_bin[grp][key[r]], (uint32_t)__popc(peers));
            }
            group_sync();
            // ---- counting sort of the group's slots by survivor count: exclusive scan of the 34 bins by its first warp, then a ranked scatter
            if (gwarp == 0u) {
                const uint32_t c0 = s_bin[grp][lane], c1 = lane < 2u ? s_bin[grp][32u + lane] : 0u;
                uint32_t inc = c0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if ((int)lane >= d) inc += t; }
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31), c32 = __shfl_sync(0xFFFFFFFFu, c1, 0);
                s_bin[grp][lane] = inc - c0;
                if (lane == 0u) s_bin[grp][32] = total; else if (lane == 1u) s_bin[grp][33] = total + c32;
            }
            group_sync();
#pragma unroll
            for (uint32_t r = 0; r < (uint32_t)kRisRows; ++r) {
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key[r]);
                const int leader = __ffs(peers) - 1;
                uint32_t base = 0u;
                if ((int)lane == leader) base = atomicAdd(&s_bin[grp][key[r]], (uint32_t)__popc(peers));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                s_sorted[slot_base + base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = (uint16_t)(slot_base + r * (uint32_t)kRisGroupThreads + gtid);
            }
            group_sync();
            // ---- phase B: survivors in candidate order, lanes aligned on the BSDF evaluation; 32 pixels of (nearly) equal survivor count per
            //      warp: warp w of the group takes set w of the ascending order, then set (last - w)
#pragma unroll 1
            for (uint32_t r = 0; r < (uint32_t)kRisRows; ++r) {
                const uint32_t set = (r & 1u) == 0u ? gwarp * (uint32_t)(kRisRows / 2) + r / 2u : (uint32_t)kRisGroupRows - 1u - (gwarp * (uint32_t)(kRisRows / 2) + r / 2u);
                const uint32_t slot = s_sorted[slot_base + set * 32u + lane];
                const uint32_t i = s_pix[slot], mask = s_mask[slot];
                const bool valid = i != kNone;
                if (!__any_sync(0xFFFFFFFFu, valid)) continue;
                Surface px; px.pos = f3(0.f); px.normal = f3(0.f); px.tangent = f3(0.f); px.incoming = f3(0.f); px.transport = f3(0.f); px.t = 0.f; px.flags = 0u;
                px.mat.color = make_float4(0.f, 0.f, 0.f, 0.f); px.mat.emissive = px.mat.color; px.mat.transmittance = px.mat.color; px.mat.tint = px.mat.color; px.mat.params = make_uint4(0u, 0u, 0u, 0u);
                if (valid) surface_load_shading(fv.surf_cur, np, i, px);
                const uint32_t s0 = wang_hash(seed + wang_hash(i + fv.pix0));
                Reservoir fresh = reservoir_zero();
                if (valid) fresh.count = (int)kPrimarySamples - __popc(mask);
                const BsdfCtx ctx = surface_ctx(px);
                if (__all_sync(0xFFFFFFFFu, allow_simple && (!valid || ctx.is_simple()))) ris_phase_b<2>(bag, s_state, slot, ctx, px, mask, s0, fresh);
                else if (__all_sync(0xFFFFFFFFu, allow_simple && (!valid || ctx.is_isotropic()))) ris_phase_b<1>(bag, s_state, slot, ctx, px, mask, s0, fresh);
                else ris_phase_b<0>(bag, s_state, slot, ctx, px, mask, s0, fresh);
                if (valid) {
                    reservoir_update_weight(fresh);
                    reservoir_store(fv.res_cur, np, i, fresh);
                }
            }
        }
    }
}

// visibility of the reservoir's sample + shading of the survivor into DIRECT, one kernel
struct VisibilityJob {
    static constexpr bool kDeferDone = true;
    FrameView fv; float shaded; uint32_t traced;
    float4 r0;                                                  // lane state between load and done: weightSum, weight, count, pdf
    LB_D bool load(uint32_t i, float3& o, float3& d, float& t0, float& t1) {
        const size_t np = fv.npix;
        const Float8 r01 = ld2(fv.res_cur + res_pair(np, 0, i));        // weights | sample position
        r0 = r01.a;
        const float4 sp = fv.surf_cur[surf_at(np, 0, i)];       // position, flags
        if (__float_as_uint(sp.w) || !(r0.y > 0.f)) return false;
        o = f3(sp);
        d = f3(r01.b) - o; const float l = length(d); d /= l;
        t0 = 0.1f; t1 = l - 0.05f;
        ++traced;
        return true;
    }
    LB_D void done(uint32_t i, bool occluded, const Tracer&) {
        if (occluded) { r0.y = 0.f; fv.res_cur[res_at(fv.npix, 0, i)] = r0; return; }
        if (r0.y > 0.f) {
            const float3 c = f3(fv.res_cur[res_at(fv.npix, 4, i)]) * (r0.y / shaded);
            float4 o = fv.channels[i]; o.x += c.x; o.y += c.y; o.z += c.z; fv.channels[i] = o;
        }
    }
};
__global__ void __launch_bounds__(kBlock, 4) k_visibility_shade(FrameView fv, BvhView bvh, uint32_t* ticket, float inv_shaded_count_denominator, unsigned long long* stat, TraceTuning tune) {
    VisibilityJob job{fv, inv_shaded_count_denominator, 0u, make_float4(0.f, 0.f, 0.f, 0.f)};
    trace_queue<true>(bvh, fv.npix, ticket, job, tune);
    const uint32_t traced = __reduce_add_sync(0xFFFFFFFFu, job.traced);
    if ((threadIdx.x & 31u) == 0u && traced) atomicAdd(stat, (unsigned long long)traced);
}

// ------------------------------------------------------------------ visibility rays binned by direction (RestirShadowRay queue, ReSTIRKernels.cu:546-582)
// The reservoir samples of neighbouring pixels point at different lights, so a warp that takes 32 consecutive pixels traces 32 rays that
// leave the same spot in 32 directions: its lanes walk different nodes (19 of 32 lanes active in the first version) and the warp lasts
// as long as its longest ray. The reference materialises the rays (32-byte RestirShadowRay, atomic append); here a pre-pass does the
// same — but per 64x32-pixel tile, counting-sorted in shared memory by the ray's DIRECTION (octahedral map, 32x32 bins in Morton order) —
// and appends the tile's rays, bin after bin, to one compact queue. Consecutive queue entries then share origin region and direction: a
// warp's 32 rays visit the same nodes. Occlusion is a pure function of (ray, triangle set), so the order changes no result.
constexpr uint32_t kBinTileW = 64, kBinTileH = 32, kBinTile = kBinTileW * kBinTileH, kBinPer = kBinTile / kBlock, kDirBins = 1024;
LB_D uint32_t spread5(uint32_t v) { v &= 31u; v = (v | (v << 4)) & 0x10Fu; v = (v | (v << 2)) & 0x133u; v = (v | (v << 1)) & 0x155u; return v; }
LB_D uint32_t direction_bin(const float3& d) {
    const float inv = 1.f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z));
    float px = d.x * inv, py = d.y * inv;
    if (d.z < 0.f) { const float ox = (1.f - fabsf(py)) * (px < 0.f ? -1.f : 1.f), oy = (1.f - fabsf(px)) * (py < 0.f ? -1.f : 1.f); px = ox; py = oy; }
    const uint32_t bx = (uint32_t)fminf(fmaxf((px * 0.5f + 0.5f) * 32.f, 0.f), 31.f), by = (uint32_t)fminf(fmaxf((py * 0.5f + 0.5f) * 32.f, 0.f), 31.f);
    return spread5(bx) | (spread5(by) << 1);
}
__global__ void __launch_bounds__(kBlock) k_vis_bin(FrameView fv, float4* __restrict__ ray_o, float4* __restrict__ ray_d, uint32_t* __restrict__ count) {
    __shared__ uint32_t s_hist[kDirBins];
    __shared__ uint32_t s_warp[kBlock / 32];
    __shared__ uint32_t s_base;
    const size_t np = fv.npix;
    const uint32_t tiles_x = (fv.width + kBinTileW - 1u) / kBinTileW, tiles_y = (fv.height + kBinTileH - 1u) / kBinTileH;
    for (uint32_t tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
        for (uint32_t b = threadIdx.x; b < kDirBins; b += kBlock) s_hist[b] = 0u;
        __syncthreads();
        float4 o4[kBinPer], d4[kBinPer]; uint32_t bin[kBinPer];
#pragma unroll
        for (uint32_t k = 0; k < kBinPer; ++k) {
            const uint32_t local = k * kBlock + threadIdx.x;
            const uint32_t x = tx * kBinTileW + (local & (kBinTileW - 1u)), y = ty * kBinTileH + local / kBinTileW;
            bin[k] = 0xFFFFFFFFu;
            if (x < fv.width && y < fv.height) {
                const uint32_t i = y * fv.width + x;
                const Float8 r01 = ld2(fv.res_cur + res_pair(np, 0, i));
                const float4 r0 = r01.a, sp = fv.surf_cur[surf_at(np, 0, i)];
                if (!__float_as_uint(sp.w) && r0.y > 0.f) {
                    const float3 o = f3(sp);
                    float3 d = f3(r01.b) - o; const float l = length(d); d /= l;
                    o4[k] = f4(o, l - 0.05f); d4[k] = f4(d, __uint_as_float(i));
                    bin[k] = direction_bin(d);
                    atomicAdd(&s_hist[bin[k]], 1u);
                }
            }
        }
        __syncthreads();
        // exclusive scan of the 1024 bin counts: 4 per thread, warp scan, scan of the 8 warp totals
        uint32_t c[4], run = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[j] = s_hist[threadIdx.x * 4u + j]; run += c[j]; }
        uint32_t incl = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off); if ((threadIdx.x & 31u) >= (uint32_t)off) incl += v; }
        if ((threadIdx.x & 31u) == 31u) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t before = 0u, total = 0u;
#pragma unroll
        for (uint32_t w = 0; w < kBlock / 32; ++w) { const uint32_t v = s_warp[w]; if (w < (threadIdx.x >> 5)) before += v; total += v; }
        if (threadIdx.x == 0u) s_base = atomicAdd(count, total);
        uint32_t excl = before + incl - run;
#pragma unroll
        for (int j = 0; j < 4; ++j) { s_hist[threadIdx.x * 4u + j] = excl; excl += c[j]; }
        __syncthreads();
        const uint32_t base = s_base;
#pragma unroll
        for (uint32_t k = 0; k < kBinPer; ++k) {
            if (bin[k] != 0xFFFFFFFFu) {
                const uint32_t at = base + atomicAdd(&s_hist[bin[k]], 1u);
                ray_o[at] = o4[k]; ray_d[at] = d4[k];
            }
        }
        __syncthreads();
    }
}
struct SortedVisibilityJob {
    static constexpr bool kDeferDone = true;
    FrameView fv; float shaded; const float4* __restrict__ ray_o; const float4* __restrict__ ray_d;
    uint32_t pixel;                                             // lane state between load and done
    LB_D bool load(uint32_t i, float3& o, float3& d, float& t0, float& t1) {
        const float4 o4 = ray_o[i], d4 = ray_d[i];
        o = f3(o4); d = f3(d4); t0 = 0.1f; t1 = o4.w; pixel = __float_as_uint(d4.w);
        return true;
    }
    LB_D void done(uint32_t, bool occluded, const Tracer&) {
        float* weight = reinterpret_cast<float*>(fv.res_cur + res_at(fv.npix, 0, pixel)) + 1;
        if (occluded) { *weight = 0.f; return; }
        const float3 c = f3(fv.res_cur[res_at(fv.npix, 4, pixel)]) * (*weight / shaded);
        float4 o = fv.channels[pixel]; o.x += c.x; o.y += c.y; o.z += c.z; fv.channels[pixel] = o;
    }
};
__global__ void __launch_bounds__(kBlock, 4) k_visibility_sorted(FrameView fv, BvhView bvh, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const uint32_t* __restrict__ count,
                                                                uint32_t* ticket, float shaded, unsigned long long* stat, TraceTuning tune) {
    const uint32_t n = *count;
    SortedVisibilityJob job{fv, shaded, ray_o, ray_d, 0u};
    trace_queue<true>(bvh, n, ticket, job, tune);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stat, (unsigned long long)n);
}

template <bool UNBIASED>
__global__ void __launch_bounds__(kBlock, LB_GATHER_BLOCKS) k_temporal(FrameView fv, uint32_t* ticket, uint32_t seed, float shaded) {
    const size_t np = fv.npix;
    const int W = (int)fv.width, H = (int)fv.height;
    const TileWalk tw(fv);
    for (uint32_t item = tw.next(ticket); item < tw.nitems; item = tw.next(ticket)) {
        int cx, cy;
        if (!tw.pixel(fv, item, cx, cy)) continue;
        const uint32_t i = (uint32_t)cy * fv.width + (uint32_t)cx;
        // two memory round trips per pixel instead of four: everything of the CURRENT pixel is requested together with its motion vector
        // (it does not depend on it), then the previous frame's similarity record and reservoir together. A pixel that turns out flagged or
        // dissimilar has read a few planes for nothing — rare, and the pass is bound by its round trips at 16 warps per SM.
        const float2 mvec = fv.motion[i];
        Surface sc; surface_load_shading(fv.surf_cur, np, i, sc);
        Reservoir prev, cur; reservoir_load(fv.res_cur, np, i, cur);
        const int mx = (int)roundf((float)W * mvec.x), my = (int)roundf((float)fv.full_height * mvec.y);
        const int ty = cy + my, tx = cx + mx; uint32_t ti = i;
        if (ty >= 0 && ty < H && tx >= 0 && tx < W) ti = (uint32_t)ty * fv.width + (uint32_t)tx;
        const SurfGeom gp = surface_geom(fv.surf_prev, np, ti);
        reservoir_load(fv.res_prev, np, ti, prev);
        if (gp.flagged || sc.flags) continue;                   // plane 1's sign bit == (flags != 0), lb_device.cuh
        if (!similar(gp.t, sc.t, gp.normal, sc.normal)) continue;
        if (prev.weight > 0.f) {
            const float3 c = prev.s.contribution * (prev.weight / shaded);
            float4 o = fv.channels[i]; o.x += c.x; o.y += c.y; o.z += c.z; fv.channels[i] = o;
        }
        prev.count = min(prev.count, cur.count * 20);
        if (UNBIASED) {
            Surface sp; surface_load_shading(fv.surf_prev, np, ti, sp);
            reservoir_store(fv.res_cur, np, i, combine_pair_unbiased(prev, cur, sc, sp, sc, wang_hash(seed + i + fv.pix0)));
        } else reservoir_store(fv.res_cur, np, i, combine_pair(prev, cur, sc, wang_hash(seed + i + fv.pix0)));
    }
}

// what spatial reuse reads of a neighbour's reservoir: 4 of its 5 planes (the stored contribution is re-evaluated)
struct ResProbe { float4 a, b, c, d; };
LB_D ResProbe res_probe(const float4* __restrict__ planes, size_t n, uint32_t i) {
    const Float8 ab = ld2(planes + res_pair(n, 0, i)), cd = ld2(planes + res_pair(n, 1, i));      // two 32-byte requests
    ResProbe p; p.a = ab.a; p.b = ab.b; p.c = cd.a; p.d = cd.b;
    return p;
}

// One pixel of SpatialNeighbourSamplingInternal (ReSTIRKernels.cu:787-980). `geom_at(nx, ny, index)` returns the similarity record (surface
// plane 1: normal, signed depth) of a pixel inside the image — from global memory (k_spatial) or from the tile staged in shared memory
// (k_spatial_tma).
// The two halves of a pixel: the probes (which of the five drawn neighbours are similar) and the merge of their reservoirs.
// spatial_probe returns the number of accepted neighbours in nb[] (pixel indices), or -1 for a pixel without a surface (nothing to do).
// `code` (optional) receives the accepted neighbours as offsets: bits 0-2 the count (7: the pixel has no surface), then 12 bits per neighbour
// in acceptance order, (dx + 32) | (dy + 32) << 6 with |dx|, |dy| <= 30 — see spatial_unpack.
template <class GeomAt>
LB_D int spatial_probe(const FrameView& fv, uint32_t seed, int x, int y, const GeomAt& geom_at, uint32_t nb[kSpatialSamples], unsigned long long* code = nullptr) {
    const int W = (int)fv.width, H = (int)fv.height;
    const uint32_t i = (uint32_t)y * fv.width + (uint32_t)x;
    // the pixel's own record and all five neighbour probes are requested together, before any is tested: one memory round trip (a pixel
    // without a surface has probed for nothing — rare)
    const float4 own = geom_at(x, y, i);
    uint32_t s = wang_hash(seed + i + fv.pix0);
    uint32_t ni[kSpatialSamples]; float4 ng[kSpatialSamples]; bool inside[kSpatialSamples];
#pragma unroll
    for (uint32_t k = 0; k < kSpatialSamples; ++k) {
        const int ny = (int)roundf((rand_f(s) * 2.f - 1.f) * (float)kSpatialRadius) + y;
        const int nx = (int)roundf((rand_f(s) * 2.f - 1.f) * (float)kSpatialRadius) + x;
        inside[k] = !(nx < 0 || nx >= W || ny < 0 || ny >= H);
        ni[k] = inside[k] ? (uint32_t)ny * fv.width + (uint32_t)nx : i;
        ng[k] = geom_at(inside[k] ? nx : x, inside[k] ? ny : y, ni[k]);
    }
    const SurfGeom gc = surf_geom_unpack(own);
    if (gc.flagged) { if (code) *code = 7ull; return -1; }
    int count = 0; unsigned long long packed = 0ull;
#pragma unroll
    for (uint32_t k = 0; k < kSpatialSamples; ++k) {
        const SurfGeom gn = surf_geom_unpack(ng[k]);
        if (inside[k] && !gn.flagged && similar(gn.t, gc.t, gn.normal, gc.normal)) {
            if (code) {
                const int ny = (int)(ni[k] / fv.width), nx = (int)(ni[k] - (uint32_t)ny * fv.width);
                packed |= (unsigned long long)((uint32_t)(nx - x + 32) | ((uint32_t)(ny - y + 32) << 6)) << (3 + 12 * count);
            }
            nb[count++] = ni[k];
        }
    }
    if (code) *code = packed | (unsigned long long)count;
    return count;
}
// the accepted neighbours of a pixel from the list its first spatial pass left: both passes run with the same seed (ReSTIR.cpp draws it once
// for the loop), so they draw the same five neighbours and accept the same ones — similarity is a function of the surfaces alone
LB_D int spatial_unpack(const FrameView& fv, unsigned long long code, int x, int y, uint32_t nb[kSpatialSamples]) {
    const int count = (int)(code & 7ull);
    if (count == 7) return -1;
#pragma unroll
    for (int k = 0; k < (int)kSpatialSamples; ++k) {
        const uint32_t f = (uint32_t)(code >> (3 + 12 * k)) & 0xFFFu;
        nb[k] = (uint32_t)(y + (int)(f >> 6) - 32) * fv.width + (uint32_t)(x + (int)(f & 63u) - 32);
    }
    return count;
}
template <bool UNBIASED>
LB_D void spatial_merge(const FrameView& fv, const float4* __restrict__ in, float4* __restrict__ out, uint32_t seed, uint32_t i, int count, const uint32_t nb[kSpatialSamples]) {
    const size_t np = fv.npix;
    const bool degenerate = seed == 0u;
    if (count > 1) {
        ResProbe cur = res_probe(in, np, nb[0]);
        Surface p0; surface_load_shading(fv.surf_cur, np, nb[0], p0);      // resampled at the FIRST accepted neighbour
        const BsdfCtx ctx = surface_ctx(p0);
        Reservoir acc = reservoir_zero(); int total = 0;
#pragma unroll 1
        for (int k = 0; k < count; ++k) {
            ResProbe nxt = cur;
            if (k + 1 < count) nxt = res_probe(in, np, nb[k + 1]);          // next neighbour's reservoir is in flight during this evaluation
            LightSample q; q.position = f3(cur.b); q.area = cur.b.w; q.normal = f3(cur.c); q.radiance = f3(cur.d); q.pdf = cur.a.w;
            // a geometrically rejected sample keeps its stored contribution, but it can only be selected when the acceptance draw is
            // exactly 0, i.e. for the all-zero xorshift state: only then is plane 4 fetched
            q.contribution = degenerate ? f3(in[res_at(np, 4, nb[k])]) : f3(0.f);
            const int qcount = __float_as_int(cur.a.z); const float qweight = cur.a.y;
            LightSample rs; resample(q, p0.pos, p0.normal, ctx, rs);
            reservoir_update(acc, rs, (float)qcount * qweight * rs.pdf, seed);   // kernel-wide seed by value (hazard 14)
            total += qcount;
            cur = nxt;
        }
        acc.count = total;
        if (!UNBIASED) reservoir_update_weight(acc);
        else {
            // the unbiased branch, ReSTIRKernels.cu:905-970: the selected sample re-evaluated at every accepted neighbour; the reference adds
            // the sample count of the OUTPUT buffer's stale reservoir of this pixel (a_ReservoirsOut[index].sampleCount, :951) — as written
            const int stale = __float_as_int(out[res_at(np, 0, i)].z);
            int correction = 0;
#pragma unroll 1
            for (int k = 0; k < count; ++k) {
                Surface pk; surface_load_shading(fv.surf_cur, np, nb[k], pk);
                const BsdfCtx ck = surface_ctx(pk);
                LightSample rs; resample(acc.s, pk.pos, pk.normal, ck, rs);
                if (rs.pdf > 0) correction += stale;
            }
            const float m = 1.f / fmaxf((float)correction, FLT_EPSILON);
            acc.weight = (1.f / fmaxf(acc.s.pdf, FLT_EPSILON)) * (m * acc.weight_sum);
        }
        reservoir_store(out, np, i, acc);
    } else {
        const float4 r0 = out[res_at(np, 0, i)];                    // Reservoir::Reset keeps the stored sample
        out[res_at(np, 0, i)] = make_float4(0.f, 0.f, __int_as_float(0), r0.w);
    }
}

template <bool UNBIASED, class GeomAt>
LB_D void spatial_pixel(const FrameView& fv, const float4* __restrict__ in, float4* __restrict__ out, uint32_t seed, int x, int y, const GeomAt& geom_at) {
    uint32_t nb[kSpatialSamples];
    const int count = spatial_probe(fv, seed, x, y, geom_at, nb);
    if (count >= 0) spatial_merge<UNBIASED>(fv, in, out, seed, (uint32_t)y * fv.width + (uint32_t)x, count, nb);
}

// PASS 0: probe, leave the accepted neighbours in nb_list, merge. PASS 1 (the second iteration): no probes — the list of pass 0, merge.
template <bool UNBIASED, int PASS>
__global__ void __launch_bounds__(kBlock, LB_GATHER_BLOCKS) k_spatial(FrameView fv, uint32_t* ticket, const float4* __restrict__ in, float4* __restrict__ out, uint32_t seed,
                                                                      unsigned long long* __restrict__ nb_list) {
    const TileWalk tw(fv);
    const float4* __restrict__ geom = fv.surf_cur + 1;                // plane 1 (normal, signed depth): the second half of pair 0
    auto geom_at = [geom](int, int, uint32_t index) { return geom[2u * (size_t)index]; };
    for (uint32_t item = tw.next(ticket); item < tw.nitems; item = tw.next(ticket)) {
        int x, y;
        if (!tw.pixel(fv, item, x, y)) continue;
        const uint32_t i = (uint32_t)y * fv.width + (uint32_t)x;
        uint32_t nb[kSpatialSamples]; int count;
        if (PASS == 0) { unsigned long long code; count = spatial_probe(fv, seed, x, y, geom_at, nb, &code); nb_list[i] = code; }
        else count = spatial_unpack(fv, nb_list[i], x, y, nb);
        if (count >= 0) spatial_merge<UNBIASED>(fv, in, out, seed, i, count, nb);
    }
}

// (Tried and dropped, r02: regrouping the pixels by accepted-neighbour count between the probes and the merge, as k_ris does with its
// survivors — the merge loop runs at 16 of 32 lanes. Warp instructions fell, the time rose 1.056 -> 1.117 ms for the two passes: the pass
// waits for its gathers, not for issue slots, and a regrouped warp gathers from 32 rows instead of one. profiles/r02_a_ab.md.)
// ---- the same pass with the neighbourhood's similarity records staged in shared memory by the TMA unit (north_star item 3).
// A block owns a 32x16-pixel tile at a time; every neighbour a pixel of the tile can draw lies within +-30 pixels, so ONE tensor copy
// (cp.async.bulk.tensor.3d over surface plane 1 — the second half of the 32-byte records of plane pair 0, so a tensor of {2 eight-byte
// elements, W records 32 bytes apart, H rows}: a box of 2 x 92 x 76 = 111 872 bytes, zero-filled outside the image) brings in everything the
// five similarity probes of all 512 pixels can touch; the probes — the first dependent step of every pixel — are then shared-memory reads (29 cycles) instead of L2 / DRAM gathers.
// One block of 16 warps per SM with TWO tile buffers: while the warps work on one tile (a row each), the TMA unit fills the other with the
// next tile the block drew from the ticket. The reservoirs of the ACCEPTED neighbours (4 planes) and the surface of the first one (8 planes)
// do not fit beside the tiles and stay global gathers. (The first version — a 3-D box with a 16-byte innermost extent, one buffer, two
// blocks per SM — moved the box as 6 992 16-byte requests and exposed every load: 2.36 ms against 1.05 ms, profiles/r02_j_spatial_tma.md.)
constexpr int kSpTileW = 32, kSpTileH = 16, kSpHalo = (int)kSpatialRadius, kSpBoxW = kSpTileW + 2 * kSpHalo, kSpBoxH = kSpTileH + 2 * kSpHalo;
constexpr int kSpBlock = 512;
constexpr uint32_t kSpTileBytes = (uint32_t)(kSpBoxW * kSpBoxH * sizeof(float4));
static_assert(2u * kSpTileBytes + 1024u <= 227u * 1024u, "two tile buffers of the TMA-staged spatial pass in one block");
static_assert(kSpBlock / 32 == kSpTileH, "16 warps = the 16 rows of a tile");
extern __shared__ __align__(128) unsigned char sp_smem[];

LB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
LB_D void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
LB_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
LB_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
LB_D void tma_load_3d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void __launch_bounds__(kSpBlock, 1) k_spatial_tma(FrameView fv, const __grid_constant__ CUtensorMap tmap, uint32_t* ticket, const float4* __restrict__ in, float4* __restrict__ out, uint32_t seed) {
    float4* const tiles[2] = {reinterpret_cast<float4*>(sp_smem), reinterpret_cast<float4*>(sp_smem + kSpTileBytes)};
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t s_tile[2];
    const uint32_t tiles_x = (fv.width + kSpTileW - 1u) / kSpTileW, tiles_y = (fv.height + kSpTileH - 1u) / kSpTileH, ntiles = tiles_x * tiles_y;
    constexpr uint32_t kStrip = 16u;                               // tiles are walked in vertical strips 16 tiles (512 px) wide, like TileWalk: the
    const uint32_t full = kStrip * tiles_y;                        // gathers of neighbouring blocks then share L2 lines
    auto origin = [&](uint32_t t, int& x0, int& y0) {
        const uint32_t strip = t / full, r = t - strip * full, sw = min(kStrip, tiles_x - strip * kStrip);
        const uint32_t ty = r / sw, tx = strip * kStrip + (r - ty * sw);
        x0 = (int)(tx * kSpTileW); y0 = (int)(ty * kSpTileH);
    };
    // thread 0 draws the next tile and starts its copy into buffer b (the buffer's previous tile is no longer read: callers synchronise first)
    auto fetch = [&](int b) {
        const uint32_t t = atomicAdd(ticket, 1u);
        s_tile[b] = t;
        if (t < ntiles) { int x0, y0; origin(t, x0, y0); mbar_expect_tx(&bar[b], kSpTileBytes); tma_load_3d(tiles[b], &tmap, &bar[b], 0, x0 - kSpHalo, y0 - kSpHalo); }
    };
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1u); mbar_init(&bar[1], 1u); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); fetch(0); }
    __syncthreads();
    uint32_t parity[2] = {0u, 0u};
    for (int b = 0;; b ^= 1) {
        const uint32_t t = s_tile[b];
        if (t >= ntiles) break;
        if (threadIdx.x == 0) fetch(b ^ 1);                        // the other buffer was released by the barrier that ended the previous iteration
        int x0, y0; origin(t, x0, y0);
        mbar_wait(&bar[b], parity[b]); parity[b] ^= 1u;
        const float4* tile = tiles[b]; const int bx = x0 - kSpHalo, by = y0 - kSpHalo;
        auto geom_at = [tile, bx, by](int nx, int ny, uint32_t) { return tile[(ny - by) * kSpBoxW + (nx - bx)]; };
        const int x = x0 + (int)(threadIdx.x & 31u), y = y0 + (int)(threadIdx.x >> 5);
        if ((uint32_t)x < fv.width && (uint32_t)y < fv.height) spatial_pixel<false>(fv, in, out, seed, x, y, geom_at);
        __syncthreads();                                           // everybody is done with buffer b, and s_tile[b ^ 1] is visible
    }
}

__global__ void __launch_bounds__(kBlock, LB_GATHER_BLOCKS) k_combine(FrameView fv, const float4* __restrict__ nbuf, uint32_t cseed) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const size_t np = fv.npix;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < fv.npix; i += stride) {
        Surface sc; surface_load_shading(fv.surf_cur, np, i, sc);      // one round trip: the flags arrive with the rest
        Reservoir a, b; reservoir_load(fv.res_cur, np, i, a); reservoir_load(nbuf, np, i, b);
        if (sc.flags) continue;
        reservoir_store(fv.res_cur, np, i, combine_pair(a, b, sc, wang_hash(cseed + i + fv.pix0)));
    }
}

} // namespace
#ifdef LB_RIS_STATS
void dump_ris_stats() {
    unsigned long long h[8]; cudaMemcpyFromSymbol(h, g_ris_stats, sizeof h);
    fprintf(stderr, "RIS stats: rows %llu, survivors/row %.2f (per lane %.2f), max-per-row %.2f, rounds/row %.2f, rounds with an exact pass %.2f, exact passes/row %.2f\n",
            h[5], (double)h[0] / h[5], (double)h[0] / h[5] / 32., (double)h[1] / h[5], (double)h[4] / h[5], (double)h[3] / h[5], (double)h[2] / h[5]);
    memset(h, 0, sizeof h); cudaMemcpyToSymbol(g_ris_stats, h, sizeof h);
}
#endif

void launch_restir(const LaunchCfg& cfg, const FrameView& fv, const SceneView& sc, const BvhView& bvh, const RestirBuffers& rb, const RestirArgs& a, uint32_t& ticket) {
    if (sc.num_lights == 0u) return;
    cudaStream_t st = cfg.stream;
    const int grid = cfg.sms * 4;
    auto lap = [&](const char* stage) { if (a.lap) a.lap(a.lap_user, stage); };
    uint32_t seed = wang_hash(a.seed);
    k_fill_bags<<<grid_for(kNumBags * kLightsPerBag, kBlock), kBlock, 0, st>>>(sc, rb.bags, a.seed); LB_LAUNCH_CHECK();
    seed = wang_hash(seed);
    LB_CUDA(cudaFuncSetAttribute(k_ris, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRisSmemBytes));      // per device; a host-side setting, no launch
    k_ris_order<<<1, 1024, 0, st>>>(seed, fv.npix, fv.pix0, rb.ris_order); LB_LAUNCH_CHECK();
    k_ris<<<cfg.sms, kRisBlock, kRisSmemBytes, st>>>(fv, sc, rb.bags, rb.ris_order, seed, a.ris_simple); LB_LAUNCH_CHECK();
    lap("restir_ris");
    const float shaded = 1.f + (a.temporal ? 1.f : 0.f) + (a.spatial ? 1.f : 0.f);
    int vis_pass = 0;
    auto visibility = [&]() {
        if (rb.vis_ray_o) {
            uint32_t* count = &fv.counters[vis_pass++ ? CNT_VIS2 : CNT_VIS];
            k_vis_bin<<<cfg.sms * 3, kBlock, 0, st>>>(fv, rb.vis_ray_o, rb.vis_ray_d, count); LB_LAUNCH_CHECK();
            k_visibility_sorted<<<grid, kBlock, 0, st>>>(fv, bvh, rb.vis_ray_o, rb.vis_ray_d, count, &fv.counters[CNT_TICKET0 + ticket++], shaded, &fv.stats[STAT_VIS], cfg.trace_any); LB_LAUNCH_CHECK();
        } else {
            k_visibility_shade<<<grid, kBlock, 0, st>>>(fv, bvh, &fv.counters[CNT_TICKET0 + ticket++], shaded, &fv.stats[STAT_VIS], cfg.trace_any); LB_LAUNCH_CHECK();
        }
        lap("restir_visibility");
    };
    visibility();
    if (a.temporal) {
        seed = wang_hash(seed);
        if (a.unbiased) k_temporal<true><<<cfg.sms * LB_GATHER_BLOCKS, kBlock, 0, st>>>(fv, &fv.counters[CNT_TICKET0 + ticket++], seed, shaded);
        else k_temporal<false><<<cfg.sms * LB_GATHER_BLOCKS, kBlock, 0, st>>>(fv, &fv.counters[CNT_TICKET0 + ticket++], seed, shaded);
        LB_LAUNCH_CHECK();
        lap("restir_temporal");
    }
    if (a.spatial) {
        seed = wang_hash(seed);
        const float4* from = fv.res_cur; float4* to = fv.res_tmp_a;
        for (uint32_t it = 0; it < kSpatialIterations; ++it) {
            if (a.unbiased) {
                if (it == 0) k_spatial<true, 0><<<cfg.sms * LB_GATHER_BLOCKS, kBlock, 0, st>>>(fv, &fv.counters[CNT_TICKET0 + ticket++], from, to, seed, rb.spatial_nb);
                else k_spatial<true, 1><<<cfg.sms * LB_GATHER_BLOCKS, kBlock, 0, st>>>(fv, &fv.counters[CNT_TICKET0 + ticket++], from, to, seed, rb.spatial_nb);
            }
            else if (rb.tmap_geom) {
                LB_CUDA(cudaFuncSetAttribute(k_spatial_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kSpTileBytes)));
                k_spatial_tma<<<cfg.sms, kSpBlock, 2 * kSpTileBytes, st>>>(fv, *static_cast<const CUtensorMap*>(rb.tmap_geom), &fv.counters[CNT_TICKET0 + ticket++], from, to, seed);
            } else if (it == 0) k_spatial<false, 0><<<cfg.sms * LB_GATHER_BLOCKS, kBlock, 0, st>>>(fv, &fv.counters[CNT_TICKET0 + ticket++], from, to, seed, rb.spatial_nb);
            else k_spatial<false, 1><<<cfg.sms * LB_GATHER_BLOCKS, kBlock, 0, st>>>(fv, &fv.counters[CNT_TICKET0 + ticket++], from, to, seed, rb.spatial_nb);
            LB_LAUNCH_CHECK();
            if (it == 0) { from = fv.res_tmp_a; to = fv.res_tmp_b; } else { const float4* t = from; from = to; to = const_cast<float4*>(t); }
        }
        lap("restir_spatial");
        visibility();
        k_combine<<<cfg.sms * LB_GATHER_BLOCKS, kBlock, 0, st>>>(fv, from, wang_hash(seed)); LB_LAUNCH_CHECK();
        lap("restir_combine");
    }
}

#ifdef LB_TRACE_STATS
void dump_trace_stats_restir() {
    unsigned long long h[8]; cudaMemcpyFromSymbol(h, g_trace_stats, sizeof h);
    for (int a = 0; a < 2; ++a) if (h[4 * a + 2]) fprintf(stderr, "TRACE stats [restir %s]: rays %llu, node visits / ray %.2f, triangle tests / ray %.2f (%.2f pass the edge test)\n", a ? "any-hit" : "closest",
        h[4 * a + 2], (double)h[4 * a] / h[4 * a + 2], (double)h[4 * a + 1] / h[4 * a + 2], (double)h[4 * a + 3] / h[4 * a + 2]);
    memset(h, 0, sizeof h); cudaMemcpyToSymbol(g_trace_stats, h, sizeof h);
}
#endif
} // namespace lb
