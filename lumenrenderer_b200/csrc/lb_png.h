// lb_png.h — 8-bit RGBA PNG encoder for the output stage (host only, no dependencies).
//
// Replaces the screenshot path of the reference: WaveFrontRenderer::GetOutputTexturePixels (LumenPT/src/Framework/
// WaveFrontRenderer.cpp:1379-1394) feeding stbi_write_png(w, h, 4, pixels, 0) in Sandbox/src/OutputLayer.cpp:882-896.
// Same file contract (colour type 6, bit depth 8, non-interlaced, rows top to bottom as stored in the output buffer);
// the compressed bytes differ from stb's, the decoded pixels do not.
//
// Encoding: every scan line uses filter 1 (Sub) or 2 (Up), whichever has the smaller sum of absolute residuals; the
// filtered stream is deflated with ONE fixed-Huffman block and a greedy LZ77 matcher over a 4-byte hash table (window 32 KB).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace lb {
namespace png {

inline uint32_t crc32(const uint8_t* p, size_t n, uint32_t crc = 0u) {
    static uint32_t table[256]; static bool ready = false;
    if (!ready) { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; } ready = true; }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 255u] ^ (crc >> 8);
    return ~crc;
}
inline uint32_t adler32(const uint8_t* p, size_t n) {
    uint32_t a = 1u, b = 0u;
    while (n) { const size_t k = n < 5552 ? n : 5552; for (size_t i = 0; i < k; ++i) { a += p[i]; b += a; } a %= 65521u; b %= 65521u; p += k; n -= k; }
    return (b << 16) | a;
}

struct BitWriter {
    std::vector<uint8_t>& out; uint64_t acc = 0; int bits = 0;
    explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
    void put(uint32_t v, int n) { acc |= (uint64_t)v << bits; bits += n; while (bits >= 8) { out.push_back((uint8_t)acc); acc >>= 8; bits -= 8; } }
    void put_rev(uint32_t code, int n) { uint32_t r = 0; for (int i = 0; i < n; ++i) r |= ((code >> i) & 1u) << (n - 1 - i); put(r, n); }   // Huffman codes are MSB first
    void flush() { if (bits) { out.push_back((uint8_t)acc); acc = 0; bits = 0; } }
};
// fixed Huffman code of a literal/length symbol (RFC 1951, 3.2.6)
inline void put_symbol(BitWriter& w, uint32_t s) {
    if (s < 144) w.put_rev(0x30 + s, 8);
    else if (s < 256) w.put_rev(0x190 + (s - 144), 9);
    else if (s < 280) w.put_rev(s - 256, 7);
    else w.put_rev(0xC0 + (s - 280), 8);
}
inline void put_match(BitWriter& w, uint32_t len, uint32_t dist) {
    static const uint16_t lbase[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
    static const uint8_t lextra[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
    static const uint16_t dbase[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
    static const uint8_t dextra[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
    int li = 28; while (lbase[li] > len) --li;
    put_symbol(w, 257u + (uint32_t)li); if (lextra[li]) w.put(len - lbase[li], lextra[li]);
    int di = 29; while (dbase[di] > dist) --di;
    w.put_rev((uint32_t)di, 5); if (dextra[di]) w.put(dist - dbase[di], dextra[di]);
}
// zlib stream (RFC 1950) of `src`
inline void deflate(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
    out.push_back(0x78); out.push_back(0x01);
    BitWriter w(out);
    w.put(1u, 1); w.put(1u, 2);                              // final block, fixed Huffman
    constexpr uint32_t kHashBits = 15, kWindow = 32768, kMaxLen = 258;
    std::vector<int64_t> head((size_t)1 << kHashBits, -1);
    size_t i = 0;
    while (i < n) {
        uint32_t best_len = 0, best_dist = 0;
        if (i + 4 <= n) {
            uint32_t key; memcpy(&key, src + i, 4);
            const uint32_t h = (key * 2654435761u) >> (32 - kHashBits);
            const int64_t cand = head[h]; head[h] = (int64_t)i;
            if (cand >= 0 && i - (size_t)cand <= kWindow) {
                const size_t lim = n - i < kMaxLen ? n - i : kMaxLen; size_t l = 0;
                while (l < lim && src[(size_t)cand + l] == src[i + l]) ++l;
                if (l >= 4) { best_len = (uint32_t)l; best_dist = (uint32_t)(i - (size_t)cand); }
            }
        }
        if (best_len) { put_match(w, best_len, best_dist); i += best_len; }
        else { put_symbol(w, src[i]); ++i; }
    }
    put_symbol(w, 256u);
    w.flush();
    const uint32_t a = adler32(src, n);
    out.push_back((uint8_t)(a >> 24)); out.push_back((uint8_t)(a >> 16)); out.push_back((uint8_t)(a >> 8)); out.push_back((uint8_t)a);
}

inline void put_chunk(std::vector<uint8_t>& f, const char type[4], const uint8_t* data, size_t n) {
    const uint32_t len = (uint32_t)n;
    f.push_back((uint8_t)(len >> 24)); f.push_back((uint8_t)(len >> 16)); f.push_back((uint8_t)(len >> 8)); f.push_back((uint8_t)len);
    const size_t at = f.size();
    f.insert(f.end(), type, type + 4); if (n) f.insert(f.end(), data, data + n);
    const uint32_t c = crc32(f.data() + at, n + 4);
    f.push_back((uint8_t)(c >> 24)); f.push_back((uint8_t)(c >> 16)); f.push_back((uint8_t)(c >> 8)); f.push_back((uint8_t)c);
}

// rgba8: h rows of w pixels, top row first. Returns the complete file.
inline std::vector<uint8_t> encode_rgba8(const uint8_t* rgba8, uint32_t w, uint32_t h) {
    const size_t stride = (size_t)w * 4;
    std::vector<uint8_t> raw((stride + 1) * h), sub(stride), up(stride);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* row = rgba8 + (size_t)y * stride; const uint8_t* prev = y ? row - stride : nullptr;
        uint64_t cost_sub = 0, cost_up = 0;
        for (size_t x = 0; x < stride; ++x) {
            sub[x] = (uint8_t)(row[x] - (x >= 4 ? row[x - 4] : 0)); up[x] = (uint8_t)(row[x] - (prev ? prev[x] : 0));
            cost_sub += (uint64_t)abs((int)(int8_t)sub[x]); cost_up += (uint64_t)abs((int)(int8_t)up[x]);
        }
        uint8_t* dst = raw.data() + (size_t)y * (stride + 1);
        const bool use_up = cost_up < cost_sub;
        dst[0] = use_up ? 2 : 1; memcpy(dst + 1, use_up ? up.data() : sub.data(), stride);
    }
    std::vector<uint8_t> z; z.reserve(raw.size() / 2 + 64); deflate(raw.data(), raw.size(), z);
    std::vector<uint8_t> f; f.reserve(z.size() + 64);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    f.insert(f.end(), sig, sig + 8);
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 6, 0, 0, 0};
    put_chunk(f, "IHDR", ihdr, 13); put_chunk(f, "IDAT", z.data(), z.size()); put_chunk(f, "IEND", nullptr, 0);
    return f;
}

inline bool write_file(const char* path, const std::vector<uint8_t>& bytes) {
    FILE* fp = fopen(path, "wb");
    if (!fp) return false;
    const bool ok = fwrite(bytes.data(), 1, bytes.size(), fp) == bytes.size();
    return fclose(fp) == 0 && ok;
}

} // namespace png
} // namespace lb
