// TEST INFRASTRUCTURE. Mutation fuzzer of the host-only asset readers, built with -fsanitize=address,undefined by tests/test_fuzz_loaders.py:
// usage: <exe> <iterations> <seed> <file>... — every mutated input must end in an error code or a usable document, never in a crash,
// an out-of-bounds access or a runaway allocation. The renderer entry points the readers call are stubbed.
#include "lumen_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
extern "C" int lb_volume_create(LbRenderer, const LbVolumeDesc*, LbHandle* out) { *out = 0; return 0; }
extern "C" const char* lb_last_error(void) { return ""; }
static uint32_t s = 12345; static uint32_t rnd() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
int main(int argc, char** argv) {
    int total = 0, opened = 0; const int iters = atoi(argv[1]); s = (uint32_t)atoi(argv[2]) | 1u;
    for (int a = 3; a < argc; ++a) {
        FILE* f = fopen(argv[a], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        std::vector<unsigned char> orig(n); fread(orig.data(), 1, n, f); fclose(f);
        for (int it = 0; it < iters; ++it) {
            std::vector<unsigned char> b = orig;
            const int kind = rnd() % 4;
            if (kind == 0) b.resize(rnd() % (b.size() + 1));
            const int edits = 1 + rnd() % 8;
            for (int e = 0; e < edits && !b.empty(); ++e) {
                size_t at = (rnd() % 3 == 0) ? rnd() % std::min<size_t>(b.size(), 1200) : rnd() % b.size();
                if (kind == 2 && at + 4 <= b.size()) { uint32_t v = rnd(); memcpy(&b[at & ~3u], &v, 4); } else b[at] = (unsigned char)rnd();
            }
            LbNanoVdb g = nullptr; ++total;
            if (lb_nanovdb_open_memory(b.data(), b.size(), rnd() % 2, &g) == 0 && g) {
                ++opened;
                LbNanoVdbInfo info; lb_nanovdb_info(g, &info);
                int32_t ijk[30]; for (int k = 0; k < 30; ++k) ijk[k] = (int32_t)(rnd() % 20000) - 10000;
                float v[10]; unsigned char on[10]; lb_nanovdb_values(g, ijk, 10, v, on);
                const long long dx = (long long)info.index_max[0] - info.index_min[0] + 1, dy = (long long)info.index_max[1] - info.index_min[1] + 1, dz = (long long)info.index_max[2] - info.index_min[2] + 1;
                if (dx > 0 && dy > 0 && dz > 0 && dx < 400 && dy < 400 && dz < 400) { std::vector<float> d((size_t)(dx * dy * dz)); lb_nanovdb_dense(g, it & 1, d.data(), d.size()); LbHandle h; lb_volume_create_nanovdb((LbRenderer)1, g, &h); }
                lb_nanovdb_close(g);
            }
        }
    }
    printf("fuzzed %d inputs, %d still opened\n", total, opened);
    return 0;
}
