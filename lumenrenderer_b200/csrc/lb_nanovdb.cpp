// lb_nanovdb.cpp — NanoVDB file ingest in front of the path: .vndb / .nvdb -> what LumenRenderer::CreateVolume needs (the grid's world
// bounding box for the volume-bounds intersection, and its voxel values as the density field of the delta tracker). Host code only.
//
// Replaces, for NanoVDB files, PTVolume::Load (PT/Framework/PTVolume.cpp:47-108: `nanovdb::io::readGrid` for ".vndb") and the
// accessors the device side uses (Shaders/volumetric_wavefront.cu:66-92: `grid.worldBBox()`). The reference links the header-only
// NanoVDB it vendors (LumenPT/vendor/openvdb/nanovdb/nanovdb, ABI 29.3.0); none of it is included here — this file restates the
// published on-disk / in-memory format of that version:
//   file    util/IO.h:107-160   segment header {magic "NanoVDB0", version, gridCount, codec} + per grid 160 B of meta data + name,
//           then per grid the buffer (:301-352: raw, or codec ZIP = u64 size + one zlib stream; BLOSC is not supported)
//   buffer  NanoVDB.h:1890-1905 GridData 672 B · :2184-2190 TreeData 64 B (byte offsets of the node arrays, relative to the tree)
//           · :2394-2456 RootData 64 B + 32-B tiles {key = 3 x 21 bits of origin >> 12, childID, state, value}
//           · :2733-2766 upper (32^3) / lower (16^3) internal nodes {bbox, offset, flags, value mask, child mask, stats, table}
//           · :3022-3040 leaves (8^3) {bbox, flags, value mask, stats, 512 values}
//   lookup  RootNode::getValue / InternalNode::getValue / LeafNode::getValue (table offset = x-major, NanoVDB.h CoordToOffset)
// Only float grids are accepted (the reference casts to nanovdb::FloatGrid and nothing else). Every offset read from the file is
// bounds-checked: a malformed file is an error code, never a crash.
//
// Density convention (DESIGN.md "NanoVDB ingest"): a fog volume's values are the density; a level set (the reference's own
// Sandbox/assets/volume/Sphere.vndb is one) becomes a fog the way OpenVDB's sdfToFogVolume does: inside, -value / background
// clamped to 1; outside 0. Negative densities are clamped to 0. The reference itself never samples the values.
#include "../../include/lumen_b200.h"
#include "lb_png.h"
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;
int vfail(int code, const std::string& msg) { g_error = msg; return code; }
[[noreturn]] void bad(const std::string& msg) { throw std::runtime_error(msg); }

constexpr uint64_t kMagic = 0x304244566f6e614eull;                   // "NanoVDB0"
constexpr size_t kGridBytes = 672, kTreeBytes = 64, kRootBytes = 64, kTileBytes = 32;
constexpr size_t kUpperBytes = 32 + 4096 + 4096 + 16 + 16 + 32768 * 4;      // 139328: header, 2 masks, stats (+16 pad to 32), table
constexpr size_t kLowerBytes = 32 + 512 + 512 + 16 + 16 + 4096 * 4;         // 17472
constexpr size_t kLeafBytes = 16 + 64 + 16 + 512 * 4;                       // 2144
constexpr size_t kUpperMask = 32, kUpperChild = 32 + 4096, kUpperTable = kUpperBytes - 32768 * 4;
constexpr size_t kLowerMask = 32, kLowerChild = 32 + 512, kLowerTable = kLowerBytes - 4096 * 4;
constexpr size_t kLeafMask = 16, kLeafValues = 96;

template <class T> T rd(const uint8_t* p) { T v; memcpy(&v, p, sizeof(T)); return v; }

struct Tile { int32_t origin[3]; int32_t child; uint32_t state; float value; };

} // namespace

struct LbNanoVdbOpaque {
    std::vector<uint8_t> buf;            // the grid buffer, as nanovdb::GridHandle holds it
    LbNanoVdbInfo info{};
    size_t tree = 0, root = 0, uppers = 0, lowers = 0, leaves = 0;
    std::vector<Tile> tiles;

    const uint8_t* at(size_t off, size_t n) const { if (off > buf.size() || n > buf.size() - off) bad("node offset runs past the end of the grid buffer"); return buf.data() + off; }
    static bool bit(const uint8_t* mask, uint32_t n) { return (rd<uint64_t>(mask + 8 * (n >> 6)) >> (n & 63u)) & 1u; }

    void parse() {
        if (buf.size() < kGridBytes + kTreeBytes + kRootBytes) bad("grid buffer too small");
        const uint8_t* g = buf.data();
        if (rd<uint64_t>(g) != kMagic) bad("grid buffer: magic number mismatch");
        const uint32_t ver = rd<uint32_t>(g + 16);
        info.version[0] = ver >> 21; info.version[1] = (ver >> 10) & 0x7FFu; info.version[2] = ver & 0x3FFu;
        if (info.version[0] != 29u) bad("grid buffer: NanoVDB ABI " + std::to_string(info.version[0]) + " (the reference vendors ABI 29)");
        info.grid_bytes = rd<uint64_t>(g + 24);
        if (info.grid_bytes != buf.size()) bad("grid buffer: size field disagrees with the meta data");
        memcpy(info.name, g + 32, 255); info.name[255] = 0;
        // Map (264 B at 288): float mat[9], invMat[9], vec[3], taper, then the same in double
        const uint8_t* map = g + 288;
        for (int k = 0; k < 9; ++k) info.map_matrix[k] = rd<double>(map + 88 + 8 * k);
        for (int k = 0; k < 3; ++k) info.map_translation[k] = rd<double>(map + 88 + 8 * (18 + k));
        for (int k = 0; k < 3; ++k) { info.world_min[k] = rd<double>(g + 552 + 8 * k); info.world_max[k] = rd<double>(g + 576 + 8 * k); info.voxel_size[k] = rd<double>(g + 600 + 8 * k); }
        info.grid_class = rd<uint32_t>(g + 624); info.grid_type = rd<uint32_t>(g + 628);
        if (info.grid_type != LB_NANOVDB_TYPE_FLOAT) bad("grid type " + std::to_string(info.grid_type) + ": only float grids are supported (the reference reads nanovdb::FloatGrid)");
        tree = kGridBytes;
        const uint8_t* t = at(tree, kTreeBytes);
        uint64_t bytes[4]; for (int k = 0; k < 4; ++k) { bytes[k] = rd<uint64_t>(t + 8 * k); info.node_count[k] = rd<uint32_t>(t + 32 + 4 * k); }
        leaves = tree + bytes[0]; lowers = tree + bytes[1]; uppers = tree + bytes[2]; root = tree + bytes[3];
        const uint8_t* r = at(root, kRootBytes);
        for (int k = 0; k < 3; ++k) { info.index_min[k] = rd<int32_t>(r + 4 * k); info.index_max[k] = rd<int32_t>(r + 12 + 4 * k); }
        info.active_voxels = rd<uint64_t>(r + 24);
        const uint32_t ntiles = rd<uint32_t>(r + 32);
        info.background = rd<float>(r + 36); info.value_min = rd<float>(r + 40); info.value_max = rd<float>(r + 44);
        at(root + kRootBytes, (size_t)ntiles * kTileBytes);
        if (uppers != root + kRootBytes + (size_t)ntiles * kTileBytes) bad("tree layout: upper nodes do not follow the root tiles");
        at(uppers, (size_t)info.node_count[2] * kUpperBytes); at(lowers, (size_t)info.node_count[1] * kLowerBytes); at(leaves, (size_t)info.node_count[0] * kLeafBytes);
        tiles.resize(ntiles);
        for (uint32_t k = 0; k < ntiles; ++k) {
            const uint8_t* p = r + kRootBytes + (size_t)k * kTileBytes;
            const uint64_t key = rd<uint64_t>(p);
            Tile& tl = tiles[k];
            tl.origin[0] = (int32_t)(uint32_t)(((key >> 42) & 0x1FFFFFull) << 12); tl.origin[1] = (int32_t)(uint32_t)(((key >> 21) & 0x1FFFFFull) << 12); tl.origin[2] = (int32_t)(uint32_t)((key & 0x1FFFFFull) << 12);
            tl.child = rd<int32_t>(p + 8); tl.state = rd<uint32_t>(p + 12); tl.value = rd<float>(p + 16);
            if (tl.child >= (int32_t)info.node_count[2]) bad("root tile: child index out of range");
        }
        if (info.index_max[0] < info.index_min[0] || info.index_max[1] < info.index_min[1] || info.index_max[2] < info.index_min[2])
            for (int k = 0; k < 3; ++k) { info.index_min[k] = 0; info.index_max[k] = -1; }                                                // empty grid
    }

    // child of an internal node: reinterpret_cast<const ChildT*>(this + mOffset) + childID  (NanoVDB.h:2766)
    size_t child_of(size_t node, size_t node_bytes, size_t table, uint32_t n, size_t child_bytes) const {
        const int32_t off = rd<int32_t>(buf.data() + node + 24);
        const uint32_t id = rd<uint32_t>(buf.data() + node + table + 4 * (size_t)n);
        const int64_t pos = (int64_t)node + (int64_t)off * (int64_t)node_bytes + (int64_t)id * (int64_t)child_bytes;
        if (pos < 0) bad("child offset before the start of the grid buffer");
        at((size_t)pos, child_bytes);
        return (size_t)pos;
    }

    // ReadAccessor::getValue / isActive
    float value(int32_t i, int32_t j, int32_t k, bool* active) const {
        const int32_t key[3] = {i & ~4095, j & ~4095, k & ~4095};
        const Tile* tl = nullptr;
        for (const Tile& c : tiles) if (c.origin[0] == key[0] && c.origin[1] == key[1] && c.origin[2] == key[2]) { tl = &c; break; }
        if (!tl) { if (active) *active = false; return info.background; }
        if (tl->child < 0) { if (active) *active = tl->state != 0u; return tl->value; }
        const size_t up = uppers + (size_t)tl->child * kUpperBytes;
        const uint32_t nu = (((uint32_t)i & 4095u) >> 7 << 10) + (((uint32_t)j & 4095u) >> 7 << 5) + (((uint32_t)k & 4095u) >> 7);
        const uint8_t* u = buf.data() + up;
        if (!bit(u + kUpperChild, nu)) { if (active) *active = bit(u + kUpperMask, nu); return rd<float>(u + kUpperTable + 4 * (size_t)nu); }
        const size_t lo = child_of(up, kUpperBytes, kUpperTable, nu, kLowerBytes);
        const uint32_t nl = (((uint32_t)i & 127u) >> 3 << 8) + (((uint32_t)j & 127u) >> 3 << 4) + (((uint32_t)k & 127u) >> 3);
        const uint8_t* l = buf.data() + lo;
        if (!bit(l + kLowerChild, nl)) { if (active) *active = bit(l + kLowerMask, nl); return rd<float>(l + kLowerTable + 4 * (size_t)nl); }
        const size_t lf = child_of(lo, kLowerBytes, kLowerTable, nl, kLeafBytes);
        const uint32_t nv = (((uint32_t)i & 7u) << 6) + (((uint32_t)j & 7u) << 3) + ((uint32_t)k & 7u);
        const uint8_t* f = buf.data() + lf;
        if (active) *active = bit(f + kLeafMask, nv);
        return rd<float>(f + kLeafValues + 4 * (size_t)nv);
    }

    float to_density(float v) const {
        if (info.grid_class == LB_NANOVDB_CLASS_LEVEL_SET) {
            if (!(v < 0.f) || !(info.background > 0.f)) return 0.f;
            const float d = -v / info.background;
            return d < 1.f ? d : 1.f;
        }
        return v > 0.f ? v : 0.f;
    }

    // Top-down fill of the dense box index_min..index_max (x fastest): tiles and background first, leaves last.
    void dense(float* out, bool as_density) const {
        const int32_t* lo = info.index_min; const int32_t* hi = info.index_max;
        if (hi[0] < lo[0]) return;
        const size_t nx = (size_t)(hi[0] - lo[0]) + 1, ny = (size_t)(hi[1] - lo[1]) + 1, nz = (size_t)(hi[2] - lo[2]) + 1;
        const float bg = as_density ? to_density(info.background) : info.background;
        for (size_t n = 0, e = nx * ny * nz; n < e; ++n) out[n] = bg;
        // fill the part of the cube [o, o + dim) that lies inside the box with one value
        auto fill = [&](const int32_t o[3], int32_t dim, float v) {
            int64_t a[3], b[3];
            for (int c = 0; c < 3; ++c) { a[c] = std::max<int64_t>(o[c], lo[c]); b[c] = std::min<int64_t>((int64_t)o[c] + dim - 1, hi[c]); if (a[c] > b[c]) return; }
            const float w = as_density ? to_density(v) : v;
            for (int64_t z = a[2]; z <= b[2]; ++z) for (int64_t y = a[1]; y <= b[1]; ++y) {
                float* row = out + ((size_t)(z - lo[2]) * ny + (size_t)(y - lo[1])) * nx;
                for (int64_t x = a[0]; x <= b[0]; ++x) row[(size_t)(x - lo[0])] = w;
            }
        };
        // range of table entries of a node at origin o with children of size `step` that touch the box, per axis
        auto range = [&](const int32_t o[3], int32_t step, int32_t count, int32_t first[3], int32_t last[3]) {
            auto floor_div = [](int64_t a, int64_t b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
            for (int c = 0; c < 3; ++c) {
                first[c] = (int32_t)std::max<int64_t>(floor_div((int64_t)lo[c] - o[c], step), 0);
                last[c] = (int32_t)std::min<int64_t>(floor_div((int64_t)hi[c] - o[c], step), count - 1);
            }
        };
        for (const Tile& tl : tiles) {
            if (tl.child < 0) { fill(tl.origin, 4096, tl.value); continue; }
            const size_t up = uppers + (size_t)tl.child * kUpperBytes; const uint8_t* u = buf.data() + up;
            int32_t f2[3], l2[3]; range(tl.origin, 128, 32, f2, l2);
            for (int32_t ux = f2[0]; ux <= l2[0]; ++ux) for (int32_t uy = f2[1]; uy <= l2[1]; ++uy) for (int32_t uz = f2[2]; uz <= l2[2]; ++uz) {
                const uint32_t nu = ((uint32_t)ux << 10) + ((uint32_t)uy << 5) + (uint32_t)uz;
                const int32_t o1[3] = {tl.origin[0] + ux * 128, tl.origin[1] + uy * 128, tl.origin[2] + uz * 128};
                if (!bit(u + kUpperChild, nu)) { fill(o1, 128, rd<float>(u + kUpperTable + 4 * (size_t)nu)); continue; }
                const size_t lw = child_of(up, kUpperBytes, kUpperTable, nu, kLowerBytes); const uint8_t* l = buf.data() + lw;
                int32_t f1[3], l1[3]; range(o1, 8, 16, f1, l1);
                for (int32_t lx = f1[0]; lx <= l1[0]; ++lx) for (int32_t ly = f1[1]; ly <= l1[1]; ++ly) for (int32_t lz = f1[2]; lz <= l1[2]; ++lz) {
                    const uint32_t nl = ((uint32_t)lx << 8) + ((uint32_t)ly << 4) + (uint32_t)lz;
                    const int32_t o0[3] = {o1[0] + lx * 8, o1[1] + ly * 8, o1[2] + lz * 8};
                    if (!bit(l + kLowerChild, nl)) { fill(o0, 8, rd<float>(l + kLowerTable + 4 * (size_t)nl)); continue; }
                    const uint8_t* lf = buf.data() + child_of(lw, kLowerBytes, kLowerTable, nl, kLeafBytes);
                    for (int32_t x = 0; x < 8; ++x) { const int64_t gx = (int64_t)o0[0] + x; if (gx < lo[0] || gx > hi[0]) continue;
                        for (int32_t y = 0; y < 8; ++y) { const int64_t gy = (int64_t)o0[1] + y; if (gy < lo[1] || gy > hi[1]) continue;
                            for (int32_t z = 0; z < 8; ++z) { const int64_t gz = (int64_t)o0[2] + z; if (gz < lo[2] || gz > hi[2]) continue;
                                const float v = rd<float>(lf + kLeafValues + 4 * (size_t)((x << 6) + (y << 3) + z));
                                out[((size_t)(gz - lo[2]) * ny + (size_t)(gy - lo[1])) * nx + (size_t)(gx - lo[0])] = as_density ? to_density(v) : v;
                            } } }
                }
            }
        }
    }
};

namespace {

bool read_whole_file(const char* path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path, "rb"); if (!f) return false;
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    out.resize((size_t)n);
    const size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0; fclose(f);
    return got == (size_t)n;
}

// nanovdb::io::readGrid(is, n): walk the segments, find grid number `index`, decode its buffer
void open_bytes(const uint8_t* p, size_t size, uint32_t index, LbNanoVdbOpaque& g) {
    size_t pos = 0; uint32_t counter = 0, total = 0; bool found = false;
    while (pos + 16 <= size) {
        if (rd<uint64_t>(p + pos) != kMagic) bad(pos == 0 ? "magic number error: this is not a NanoVDB file" : "segment header: magic number mismatch");
        const uint32_t ver = rd<uint32_t>(p + pos + 8); const uint16_t count = rd<uint16_t>(p + pos + 12), codec = rd<uint16_t>(p + pos + 14);
        if ((ver >> 21) != 29u) bad("file written by NanoVDB ABI " + std::to_string(ver >> 21) + "; the reference vendors ABI 29 and rejects others (util/IO.h:464-475)");
        if (count == 0) bad("segment contains no grids");
        pos += 16;
        struct Meta { uint64_t grid_size, file_size; };
        std::vector<Meta> metas(count);
        for (uint16_t k = 0; k < count; ++k) {
            if (pos + 160 > size) bad("truncated file: grid meta data");
            metas[k].grid_size = rd<uint64_t>(p + pos); metas[k].file_size = rd<uint64_t>(p + pos + 8);
            const uint32_t name_size = rd<uint32_t>(p + pos + 136);
            if (name_size > size - pos - 160) bad("truncated file: grid name");
            pos += 160 + (size_t)name_size;
        }
        for (uint16_t k = 0; k < count; ++k, ++counter) {
            if (metas[k].file_size > size - pos) bad("truncated file: grid " + std::to_string(counter) + " needs " + std::to_string(metas[k].file_size) + " bytes");
            if (counter == index && !found) {
                found = true; g.info.codec = codec;
                if (metas[k].grid_size > ((uint64_t)1 << 40)) bad("grid size field is implausible");
                if (codec == 0) {
                    if (metas[k].file_size != metas[k].grid_size) bad("uncompressed grid: file size and grid size differ");
                    g.buf.assign(p + pos, p + pos + metas[k].grid_size);
                } else if (codec == 1) {                                   // ZIP: u64 byte count + one zlib stream (util/IO.h:308-323)
                    if (metas[k].file_size < 8) bad("truncated ZIP block");
                    const uint64_t zsize = rd<uint64_t>(p + pos);
                    if (zsize > metas[k].file_size - 8) bad("ZIP block larger than the grid's share of the file");
                    if (metas[k].grid_size > zsize * 1032u + 65536u) bad("ZIP codec: the grid size field exceeds what the block can inflate to");      // deflate expands at most ~1032 : 1
                    g.buf.reserve((size_t)metas[k].grid_size);
                    if (!lb::png::inflate(p + pos + 8, (size_t)zsize, g.buf, (size_t)metas[k].grid_size)) bad("ZIP codec: inflate failed");
                    if (g.buf.size() != metas[k].grid_size) bad("ZIP codec: decompressed size differs from the grid size");
                } else bad(codec == 2 ? "BLOSC-compressed NanoVDB files are not supported (re-save with codec NONE or ZIP)" : "unknown compression codec");
            }
            pos += (size_t)metas[k].file_size;
        }
        total += count;
    }
    if (pos == 0) bad("magic number error: this is not a NanoVDB file");
    if (!found) bad("grid index " + std::to_string(index) + " exceeds the grid count of the file (" + std::to_string(total) + ")");
    g.info.grid_count = total;
    g.parse();
}

bool ends_with(const std::string& s, const char* suffix) {
    const size_t n = strlen(suffix); if (s.size() < n) return false;
    for (size_t k = 0; k < n; ++k) if (tolower((unsigned char)s[s.size() - n + k]) != suffix[k]) return false;
    return true;
}

} // namespace

extern "C" {

LB_API const char* lb_nanovdb_last_error(void) { return g_error.c_str(); }

LB_API int lb_nanovdb_open_memory(const void* bytes, size_t size, uint32_t grid_index, LbNanoVdb* out) {
    if (!bytes || !out) return vfail(LB_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    try {
        std::unique_ptr<LbNanoVdbOpaque> g(new LbNanoVdbOpaque());
        open_bytes(static_cast<const uint8_t*>(bytes), size, grid_index, *g);
        *out = g.release();
        return LB_OK;
    } catch (const std::bad_alloc&) { return vfail(LB_ERR_OUT_OF_MEMORY, "out of memory");
    } catch (const std::exception& e) { const std::string m = e.what(); return vfail(m.find("not supported") != std::string::npos || m.find("only float") != std::string::npos ? LB_ERR_UNSUPPORTED : LB_ERR_INVALID_ARGUMENT, m); }
}
LB_API int lb_nanovdb_open(const char* path, uint32_t grid_index, LbNanoVdb* out) {
    if (!path || !out) return vfail(LB_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    std::vector<uint8_t> file;
    if (!read_whole_file(path, file)) return vfail(LB_ERR_INVALID_ARGUMENT, std::string("cannot read ") + path);
    return lb_nanovdb_open_memory(file.data(), file.size(), grid_index, out);
}
LB_API int lb_nanovdb_close(LbNanoVdb g) { delete g; return LB_OK; }
LB_API int lb_nanovdb_info(LbNanoVdb g, LbNanoVdbInfo* out) { if (!g || !out) return vfail(LB_ERR_INVALID_ARGUMENT, "null argument"); *out = g->info; return LB_OK; }
LB_API int lb_nanovdb_values(LbNanoVdb g, const int32_t* ijk3, uint32_t n, float* values, uint8_t* active) {
    if (!g || (n && (!ijk3 || !values))) return vfail(LB_ERR_INVALID_ARGUMENT, "null argument");
    try {
        for (uint32_t k = 0; k < n; ++k) { bool on = false; values[k] = g->value(ijk3[3 * k], ijk3[3 * k + 1], ijk3[3 * k + 2], &on); if (active) active[k] = on ? 1 : 0; }
        return LB_OK;
    } catch (const std::exception& e) { return vfail(LB_ERR_INVALID_ARGUMENT, e.what()); }
}
LB_API int lb_nanovdb_dense(LbNanoVdb g, int as_density, float* out, size_t capacity_floats) {
    if (!g || !out) return vfail(LB_ERR_INVALID_ARGUMENT, "null argument");
    const LbNanoVdbInfo& i = g->info;
    if (i.index_max[0] < i.index_min[0]) return LB_OK;
    // index_min / index_max come unchecked from the file's root bounding box: cap every extent (as lb_volume_create_nanovdb does) before the
    // product, which would otherwise wrap (2^32 x 2^32 x 1 -> 0) and let dense() write outside `out`
    uint64_t ext[3];
    for (int a = 0; a < 3; ++a) {
        if (i.index_max[a] < i.index_min[a]) return LB_OK;
        ext[a] = (uint64_t)((int64_t)i.index_max[a] - (int64_t)i.index_min[a]) + 1;
        if (ext[a] > 8192) return vfail(LB_ERR_OUT_OF_MEMORY, "dense box of the grid exceeds 8192 voxels along an axis");
    }
    const uint64_t need = ext[0] * ext[1] * ext[2];
    if (need > ((uint64_t)1 << 33)) return vfail(LB_ERR_OUT_OF_MEMORY, "dense box of the grid exceeds 2^33 voxels");
    if (capacity_floats < need) return vfail(LB_ERR_INVALID_ARGUMENT, "dense buffer too small: " + std::to_string(need) + " floats needed");
    try { g->dense(out, as_density != 0); return LB_OK; } catch (const std::exception& e) { return vfail(LB_ERR_INVALID_ARGUMENT, e.what()); }
}

// LumenRenderer::CreateVolume for a parsed NanoVDB grid: the volume's object-space box is the grid's world bounding box
// (volumetric_wavefront.cu:87 intersects grid.worldBBox() in the instance's object space), its density field the dense box of values.
LB_API int lb_volume_create_nanovdb(LbRenderer r, LbNanoVdb g, LbHandle* out) {
    if (!r || !g || !out) return vfail(LB_ERR_INVALID_ARGUMENT, "null argument");
    const LbNanoVdbInfo& i = g->info;
    if (i.index_max[0] < i.index_min[0]) return vfail(LB_ERR_INVALID_ARGUMENT, "the grid has no active voxels");
    // the dense box is axis-aligned in index space: the index -> world map must be scale + translation (what every grid the reference can
    // load has — NanoVDB itself assumes an affine map of uniform scale, NanoVDB.h:1886)
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) if (a != b && i.map_matrix[a * 3 + b] != 0.0) return vfail(LB_ERR_UNSUPPORTED, "grids with a rotated or sheared index-to-world map are not supported");
    const uint64_t dims[3] = {(uint64_t)(i.index_max[0] - (int64_t)i.index_min[0]) + 1, (uint64_t)(i.index_max[1] - (int64_t)i.index_min[1]) + 1, (uint64_t)(i.index_max[2] - (int64_t)i.index_min[2]) + 1};
    if (dims[0] > 8192 || dims[1] > 8192 || dims[2] > 8192 || dims[0] * dims[1] * dims[2] > ((uint64_t)1 << 33)) return vfail(LB_ERR_OUT_OF_MEMORY, "dense box of the grid exceeds 2^33 voxels");
    try {
        std::vector<float> dense((size_t)(dims[0] * dims[1] * dims[2]));
        g->dense(dense.data(), true);
        LbVolumeDesc d{}; d.density = dense.data(); d.nx = (uint32_t)dims[0]; d.ny = (uint32_t)dims[1]; d.nz = (uint32_t)dims[2];
        for (int a = 0; a < 3; ++a) {
            // voxel (i) covers [i, i + 1) in index space: the box of the stored values, mapped to world space (equals worldBBox() for the
            // grids NanoVDB's own builders write; computed from the map so that it is exactly the box the dense values tile)
            const double w0 = i.map_matrix[a * 3 + a] * (double)i.index_min[a] + i.map_translation[a], w1 = i.map_matrix[a * 3 + a] * ((double)i.index_max[a] + 1.0) + i.map_translation[a];
            d.bbox_min[a] = (float)std::min(w0, w1); d.bbox_max[a] = (float)std::max(w0, w1);
        }
        if (i.map_matrix[0] < 0.0 || i.map_matrix[4] < 0.0 || i.map_matrix[8] < 0.0) return vfail(LB_ERR_UNSUPPORTED, "grids with a mirrored index-to-world map are not supported");
        const int rc = lb_volume_create(r, &d, out);
        if (rc != LB_OK) g_error = lb_last_error();
        return rc;
    } catch (const std::bad_alloc&) { return vfail(LB_ERR_OUT_OF_MEMORY, "out of memory");
    } catch (const std::exception& e) { return vfail(LB_ERR_INVALID_ARGUMENT, e.what()); }
}

// LumenRenderer::CreateVolume(const std::string&) (LM/Renderer/LumenRenderer.h:168; PTVolume::Load dispatches on the extension,
// PT/Framework/PTVolume.cpp:69-107)
LB_API int lb_volume_create_file(LbRenderer r, const char* path, LbHandle* out) {
    if (!r || !path || !out) return vfail(LB_ERR_INVALID_ARGUMENT, "null argument");
    const std::string p(path);
    if (ends_with(p, ".vdb")) return vfail(LB_ERR_UNSUPPORTED, "OpenVDB .vdb files need the OpenVDB library to decode; convert to NanoVDB (.vndb / .nvdb) first");
    if (!ends_with(p, ".vndb") && !ends_with(p, ".nvdb")) return vfail(LB_ERR_UNSUPPORTED, "file type not compatible with volume loading (PTVolume.cpp:104-107): " + p);
    LbNanoVdb g = nullptr;
    int rc = lb_nanovdb_open(path, 0u, &g);
    if (rc != LB_OK) return rc;
    rc = lb_volume_create_nanovdb(r, g, out);
    lb_nanovdb_close(g);
    return rc;
}

} // extern "C"
