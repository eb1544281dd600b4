// ref_nanovdb.cpp — TEST INFRASTRUCTURE. Host program over the reference's own vendored NanoVDB (ABI 29.3.0,
// /root/reference/Lumen_Engine/LumenPT/vendor/openvdb/nanovdb/nanovdb/{NanoVDB.h,util/IO.h,util/Primitives.h}), compiled in
// place by oracle/Makefile into oracle/_ref/ref_nanovdb (nothing of the reference is copied into this repository).
// It is what PTVolume::Load does for a .vndb (PT/Framework/PTVolume.cpp:93-98: nanovdb::io::readGrid) plus the accessors the
// volumetric intersection program uses (Shaders/volumetric_wavefront.cu:66-92: grid.worldBBox()).
//
//   ref_nanovdb make <fog|ls> <radius> <voxel> <halfwidth> <cx> <cy> <cz> <out.vndb> [zip]   write a small fixture file
//   ref_nanovdb dump <in.vndb> <nsamples> <seed> <out.bin>                              read a file and dump what a reader must reproduce
//
// dump layout (little endian): u32 gridType, u32 gridClass, i32 indexBBox[6], f64 worldBBox[6], f64 voxelSize[3], f64 mat[9], f64 vec[3],
// u64 activeVoxels, f32 background, f32 min, f32 max, u32 nodeCount[4], u32 nsamples, then nsamples x {i32 i, j, k, f32 value, u32 active},
// then u64 n = dense voxel count of the index bbox and a double sum + u64 xor-fold of the float bit patterns of all dense values
// (x fastest, z slowest — the order of LbVolumeDesc::density).
#include <nanovdb/NanoVDB.h>
#include <nanovdb/util/IO.h>
#include <nanovdb/util/Primitives.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

static uint32_t xorshift(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string cmd = argv[1];
    try {
        if (cmd == "make" && (argc == 10 || argc == 11)) {
            const std::string kind = argv[2];
            const float radius = (float)atof(argv[3]), voxel = (float)atof(argv[4]), half = (float)atof(argv[5]);
            const nanovdb::Vec3d c(atof(argv[6]), atof(argv[7]), atof(argv[8]));
            nanovdb::GridHandle<> h = kind == "fog" ? nanovdb::createFogVolumeSphere<float>(radius, c, voxel, half, nanovdb::Vec3d(0), "sphere_fog")
                                                    : nanovdb::createLevelSetSphere<float>(radius, c, voxel, half, nanovdb::Vec3d(0), "sphere_ls");
            nanovdb::io::Codec codec = nanovdb::io::Codec::NONE;
#ifdef NANOVDB_USE_ZIP
            if (argc == 11 && std::string(argv[10]) == "zip") codec = nanovdb::io::Codec::ZIP;
#endif
            nanovdb::io::writeGrid(argv[9], h, codec);
            return 0;
        }
        if (cmd == "dump" && argc == 6) {
            auto h = nanovdb::io::readGrid(argv[2]);
            const nanovdb::FloatGrid* g = h.grid<float>();
            if (!g) { fprintf(stderr, "not a float grid\n"); return 3; }
            FILE* f = fopen(argv[5], "wb"); if (!f) return 4;
            auto put = [&](const void* p, size_t n) { fwrite(p, 1, n, f); };
            const uint32_t gt = (uint32_t)g->gridType(), gc = (uint32_t)g->gridClass(); put(&gt, 4); put(&gc, 4);
            const auto ib = g->indexBBox(); const int32_t ibb[6] = {ib.min()[0], ib.min()[1], ib.min()[2], ib.max()[0], ib.max()[1], ib.max()[2]}; put(ibb, 24);
            const auto wb = g->worldBBox(); const double wbb[6] = {wb.min()[0], wb.min()[1], wb.min()[2], wb.max()[0], wb.max()[1], wb.max()[2]}; put(wbb, 48);
            const auto vs = g->voxelSize(); const double vsz[3] = {vs[0], vs[1], vs[2]}; put(vsz, 24);
            // index -> world map probed through the public API: columns of the matrix = images of the unit vectors minus the image of 0
            const nanovdb::Vec3d o = g->indexToWorld(nanovdb::Vec3d(0.0));
            double mat[9];
            for (int c = 0; c < 3; ++c) { nanovdb::Vec3d e(0.0); e[c] = 1.0; const nanovdb::Vec3d w = g->indexToWorld(e); for (int r = 0; r < 3; ++r) mat[r * 3 + c] = w[r] - o[r]; }
            put(mat, 72); const double vec[3] = {o[0], o[1], o[2]}; put(vec, 24);
            const uint64_t av = g->activeVoxelCount(); put(&av, 8);
            const float bg = g->tree().background(); float mn, mx; g->tree().extrema(mn, mx); put(&bg, 4); put(&mn, 4); put(&mx, 4);
            const uint32_t nc[4] = {g->tree().nodeCount(0), g->tree().nodeCount(1), g->tree().nodeCount(2), 1u}; put(nc, 16);
            const uint32_t ns = (uint32_t)atoi(argv[3]); put(&ns, 4);
            uint32_t s = (uint32_t)strtoul(argv[4], nullptr, 10) | 1u;
            auto acc = g->getAccessor();
            for (uint32_t k = 0; k < ns; ++k) {
                int32_t c[3];
                for (int a = 0; a < 3; ++a) { const int lo = ibb[a] - 9, hi = ibb[3 + a] + 9; c[a] = lo + (int)(xorshift(s) % (uint32_t)(hi - lo + 1)); }
                const nanovdb::Coord ijk(c[0], c[1], c[2]);
                const float v = acc.getValue(ijk); const uint32_t on = acc.isActive(ijk) ? 1u : 0u;
                put(c, 12); put(&v, 4); put(&on, 4);
            }
            uint64_t n = 0, fold = 0; double sum = 0.0;
            for (int z = ibb[2]; z <= ibb[5]; ++z) for (int y = ibb[1]; y <= ibb[4]; ++y) for (int x = ibb[0]; x <= ibb[3]; ++x) {
                const float v = acc.getValue(nanovdb::Coord(x, y, z)); uint32_t b; memcpy(&b, &v, 4);
                fold = ((fold << 7) | (fold >> 57)) ^ b; sum += (double)v; ++n;
            }
            put(&n, 8); put(&sum, 8); put(&fold, 8);
            fclose(f);
            return 0;
        }
    } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 5; }
    fprintf(stderr, "usage: see the header of ref_nanovdb.cpp\n");
    return 2;
}
