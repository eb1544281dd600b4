// TEST INFRASTRUCTURE. Mutation fuzzer of the host-only asset readers, built with -fsanitize=address,undefined by tests/test_fuzz_loaders.py:
// usage: <exe> <iterations> <seed> <file>... — every mutated input must end in an error code or a usable document, never in a crash,
// an out-of-bounds access or a runaway allocation. The renderer entry points the readers call are stubbed.
#include "lumen_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
// renderer entry points lb_gltf_upload calls: not exercised by the fuzzer
extern "C" int lb_texture_create(LbRenderer, const uint8_t*, uint32_t, uint32_t, int, LbHandle* o) { *o = 0; return 0; }
extern "C" int lb_material_create(LbRenderer, const LbMaterialDesc*, LbHandle* o) { *o = 0; return 0; }
extern "C" int lb_primitive_create(LbRenderer, const LbPrimitiveDesc*, LbHandle* o) { *o = 0; return 0; }
extern "C" int lb_mesh_create(LbRenderer, const LbHandle*, uint32_t, LbHandle* o) { *o = 0; return 0; }
extern "C" int lb_scene_add_mesh_instance(LbRenderer, LbHandle, const float*, const LbEmissiveness*, LbHandle, LbHandle* o) { *o = 0; return 0; }
extern "C" const char* lb_last_error(void) { return ""; }
static uint32_t s = 777; static uint32_t rnd() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
int main(int argc, char** argv) {
    int total = 0, opened = 0; const int iters = atoi(argv[1]); s = (uint32_t)atoi(argv[2]) | 1u;
    for (int a = 3; a < argc; ++a) {
        FILE* f = fopen(argv[a], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        std::vector<unsigned char> orig(n); if (fread(orig.data(), 1, n, f) != (size_t)n) return 1; fclose(f);
        const bool glb = strstr(argv[a], ".glb") != nullptr;
        const bool ollad = strstr(argv[a], ".ollad") != nullptr;                  // binary cache file: byte mutations only
        const std::string tmp = std::string(getenv("FUZZ_TMP") ? getenv("FUZZ_TMP") : "/tmp") + "/fuzz_gltf" + (ollad ? ".ollad" : glb ? ".glb" : ".gltf");
        for (int it = 0; it < iters; ++it) {
            std::vector<unsigned char> b = orig;
            const int kind = ollad ? (rnd() % 2 ? 0 : 3) : rnd() % 5;
            if (kind == 0) b.resize(rnd() % (b.size() + 1));
            const int edits = 1 + rnd() % 6;
            for (int e = 0; e < edits && !b.empty(); ++e) {
                size_t at = rnd() % b.size();
                if (kind == 1) { static const char* tok[] = {"-1", "99999999", "0", "{}", "[]", "null", "1e308", "\"\"", ",", "}"}; const char* t = tok[rnd() % 10]; for (size_t k = 0; t[k] && at + k < b.size(); ++k) b[at + k] = (unsigned char)t[k]; }
                else if (kind == 2 && !glb && !ollad) { // digit tweak: hits counts, offsets, indices
                    for (size_t k = at; k < b.size() && k < at + 200; ++k) if (b[k] >= '0' && b[k] <= '9') { b[k] = (unsigned char)('0' + rnd() % 10); break; } }
                else b[at] = (unsigned char)rnd();
            }
            FILE* o = fopen(tmp.c_str(), "wb"); fwrite(b.data(), 1, b.size(), o); fclose(o);
            LbGltf g = nullptr; ++total;
            if (lb_gltf_open(tmp.c_str(), nullptr, nullptr, &g) == 0 && g) {
                ++opened;
                LbGltfInfo info; lb_gltf_info(g, &info);
                for (uint32_t m = 0; m < info.meshes; ++m) { uint32_t c = 0; lb_gltf_mesh_primitive_count(g, m, &c); for (uint32_t p = 0; p < c; ++p) { LbPrimitiveDesc d; lb_gltf_primitive(g, m, p, &d); } }
                for (uint32_t i = 0; i < info.materials; ++i) { LbMaterialDesc d; lb_gltf_material(g, i, &d); }
                LbHandle first; uint32_t cnt; lb_gltf_upload((LbRenderer)1, g, nullptr, &first, &cnt);
                if (it % 8 == 0) lb_gltf_save_ollad(g, (tmp + ".out").c_str());
                lb_gltf_close(g);
            }
        }
    }
    printf("fuzzed %d inputs, %d still opened\n", total, opened);
    return 0;
}
