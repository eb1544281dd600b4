// lb_png.h — 8-bit RGBA PNG encoder for the output stage and PNG decoder for the glTF ingest (host only, no dependencies).
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


// ---------------------------------------------------------------- decoding (glTF ingest: the reference decodes with stbi_load_from_memory(..., 4),
// LumenPT/src/Tools/LumenPTModelConverter.cpp:121). PNG (plain or Adam7-interlaced) of every bit depth (1, 2, 4, 8, 16) and colour type, palette and colour-key transparency; output RGBA8.
struct BitReader {
    const uint8_t* p; size_t n, pos = 0; uint64_t acc = 0; int bits = 0;
    BitReader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
    uint32_t get(int k) { while (bits < k) { acc |= (uint64_t)(pos < n ? p[pos] : 0) << bits; ++pos; bits += 8; } const uint32_t v = (uint32_t)(acc & ((1ull << k) - 1)); acc >>= k; bits -= k; return v; }
    void align() { acc >>= (bits & 7); bits -= (bits & 7); }
};
struct Huffman {
    uint16_t count[16] = {0}; uint16_t symbol[288];
    void build(const uint8_t* lengths, int n) {
        for (int i = 0; i < 16; ++i) count[i] = 0;
        for (int i = 0; i < n; ++i) ++count[lengths[i]];
        count[0] = 0;
        uint16_t offs[16]; offs[1] = 0;
        for (int i = 1; i < 15; ++i) offs[i + 1] = offs[i] + count[i];
        for (int i = 0; i < n; ++i) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    }
    int decode(BitReader& br) const {                      // canonical code, one bit at a time (textures are decoded once per load)
        int code = 0, first = 0, index = 0;
        for (int len = 1; len <= 15; ++len) {
            code |= (int)br.get(1);
            const int c = count[len];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};
// `max_out`: the decompressed size the caller expects (PNG: rows x (stride + 1); NanoVDB: the grid size). A stream that produces more, or
// that keeps decoding after its input is exhausted (the bit reader pads with zeros), is rejected instead of growing the output without bound.
inline bool inflate(const uint8_t* src, size_t n, std::vector<uint8_t>& out, size_t max_out) {
    if (n < 6) return false;
    BitReader br(src + 2, n - 2);                           // zlib header: CMF, FLG
    static const uint16_t lbase[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
    static const uint8_t lextra[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
    static const uint16_t dbase[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
    static const uint8_t dextra[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
    for (;;) {
        const uint32_t final = br.get(1), type = br.get(2);
        if (type == 0) {
            br.align();
            const uint32_t len = br.get(16), nlen = br.get(16);
            if ((len ^ 0xFFFFu) != nlen) return false;
            if (out.size() + len > max_out || br.pos > br.n + 8) return false;
            for (uint32_t i = 0; i < len; ++i) out.push_back((uint8_t)br.get(8));
        } else if (type == 1 || type == 2) {
            Huffman lit, dist; uint8_t lengths[320];
            if (type == 1) {
                for (int i = 0; i < 288; ++i) lengths[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
                lit.build(lengths, 288);
                for (int i = 0; i < 30; ++i) lengths[i] = 5;
                dist.build(lengths, 30);
            } else {
                const int hlit = (int)br.get(5) + 257, hdist = (int)br.get(5) + 1, hclen = (int)br.get(4) + 4;
                static const uint8_t order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
                uint8_t cl[19] = {0};
                for (int i = 0; i < hclen; ++i) cl[order[i]] = (uint8_t)br.get(3);
                Huffman clh; clh.build(cl, 19);
                int i = 0;
                while (i < hlit + hdist) {
                    const int sym = clh.decode(br);
                    if (sym < 0) return false;
                    if (sym < 16) lengths[i++] = (uint8_t)sym;
                    else {
                        uint8_t prev = 0; int rep;
                        if (sym == 16) { if (i == 0) return false; prev = lengths[i - 1]; rep = 3 + (int)br.get(2); }
                        else if (sym == 17) rep = 3 + (int)br.get(3);
                        else rep = 11 + (int)br.get(7);
                        if (i + rep > hlit + hdist) return false;
                        while (rep--) lengths[i++] = prev;
                    }
                }
                lit.build(lengths, hlit); dist.build(lengths + hlit, hdist);
            }
            for (;;) {
                const int sym = lit.decode(br);
                if (sym < 0 || out.size() >= max_out + 1 || br.pos > br.n + 8) return false;
                if (sym < 256) out.push_back((uint8_t)sym);
                else if (sym == 256) break;
                else {
                    if (sym > 285) return false;
                    const uint32_t len = lbase[sym - 257] + br.get(lextra[sym - 257]);
                    const int ds = dist.decode(br);
                    if (ds < 0 || ds > 29) return false;
                    const uint32_t d = dbase[ds] + br.get(dextra[ds]);
                    if (d > out.size() || out.size() + len > max_out) return false;
                    const size_t from = out.size() - d;
                    for (uint32_t k = 0; k < len; ++k) out.push_back(out[from + k]);
                }
            }
        } else return false;
        if (final) break;
        if (br.pos > br.n + 8) return false;
    }
    return true;
}
inline bool is_png(const uint8_t* p, size_t n) { static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}; return n >= 8 && memcmp(p, sig, 8) == 0; }
// reverses the row filters of one (sub-)image: `in` = h rows of (filter byte + stride bytes), `out` = h rows of stride bytes
inline bool unfilter(const uint8_t* in, uint8_t* out, size_t stride, uint32_t h, size_t bpp) {
    for (uint32_t y = 0; y < h; ++y) {
        const int f = in[0]; ++in;
        uint8_t* row = out + (size_t)y * stride; const uint8_t* up = y ? row - stride : nullptr;
        if (f < 0 || f > 4) return false;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= bpp ? row[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int pred = 0;
            if (f == 1) pred = a; else if (f == 2) pred = b; else if (f == 3) pred = (a + b) >> 1;
            else if (f == 4) { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
            row[x] = (uint8_t)(in[x] + pred);
        }
        in += stride;
    }
    return true;
}
// false = not a PNG this decoder handles (damaged)
#if defined(__GNUC__) && !defined(__clang__)
__attribute__((optimize("no-tree-slp-vectorize")))            // g++ 13.3 -O2 dies in its SLP vectoriser on this function (compute_live_loop_exits)
#endif
inline bool decode_rgba8(const uint8_t* file, size_t n, std::vector<uint8_t>& rgba, uint32_t& w, uint32_t& h) {
    if (!is_png(file, n)) return false;
    size_t pos = 8; std::vector<uint8_t> idat, plte, trns; int depth = 0, colour = 0, lace = 0; w = h = 0;
    auto be = [&](size_t at) { return ((uint32_t)file[at] << 24) | ((uint32_t)file[at + 1] << 16) | ((uint32_t)file[at + 2] << 8) | file[at + 3]; };
    while (pos + 12 <= n) {
        const uint32_t len = be(pos); const uint8_t* type = file + pos + 4; const uint8_t* body = file + pos + 8;
        if (pos + 12 + (size_t)len > n) return false;
        if (!memcmp(type, "IHDR", 4) && len >= 13) { w = be(pos + 8); h = be(pos + 12); depth = body[8]; colour = body[9]; lace = body[12]; }
        else if (!memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!w || !h || lace > 1 || (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16)) return false;
    if (w > 65536u || h > 65536u || (uint64_t)w * h > ((uint64_t)1 << 28)) return false;      // a corrupt header must not turn into a giant allocation
    const int channels = colour == 0 ? 1 : colour == 2 ? 3 : colour == 3 ? 1 : colour == 4 ? 2 : colour == 6 ? 4 : 0;
    if (!channels || (colour == 3 && depth == 16) || (depth < 8 && colour != 0 && colour != 3)) return false;
    // sub-byte samples (grey and palette images of depth 1, 2, 4): rows are bit-packed, the filters work on whole bytes
    const size_t bits = (size_t)channels * depth, bpp = bits >= 8 ? bits / 8 : 1;
    // passes: one for a plain image, the seven Adam7 sub-images (each filtered on its own) for an interlaced one
    static const uint32_t x0[7] = {0, 4, 0, 2, 0, 1, 0}, y0[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
    struct Pass { uint32_t x0, y0, dx, dy, w, h; size_t stride, at; };
    std::vector<Pass> passes; size_t total = 0, plain = 0;
    for (int k = 0; k < (lace ? 7 : 1); ++k) {
        Pass q = lace ? Pass{x0[k], y0[k], dx[k], dy[k], (w - x0[k] + dx[k] - 1) / dx[k], (h - y0[k] + dy[k] - 1) / dy[k], 0, 0} : Pass{0, 0, 1, 1, w, h, 0, 0};
        if (!q.w || !q.h) continue;
        q.stride = ((size_t)q.w * bits + 7) / 8; q.at = plain;
        total += (q.stride + 1) * q.h; plain += q.stride * q.h;
        passes.push_back(q);
    }
    std::vector<uint8_t> raw; raw.reserve(total);
    if (!inflate(idat.data(), idat.size(), raw, total) || raw.size() < total) return false;
    std::vector<uint8_t> img(plain);
    const uint8_t* in = raw.data();
    for (size_t k = 0; k < passes.size(); ++k) {
        const Pass& q = passes[k];
        if (!unfilter(in, img.data() + q.at, q.stride, q.h, bpp)) return false;
        in += (q.stride + 1) * q.h;
    }
    rgba.resize((size_t)w * h * 4);
    const size_t step = depth == 16 ? 2 : 1;                  // 16-bit samples: the high byte (what stb's 8-bit interface returns)
    // tRNS of a grey / RGB image is a colour key (stb_image: stbi__compute_transparency on the 8-bit samples — sub-byte grey after scaling —
    // and stbi__compute_transparency16 on the 16-bit ones)
    static const int scale_of_depth[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    const bool keyed = (colour == 0 && trns.size() >= 2) || (colour == 2 && trns.size() >= 6);
    uint32_t key[3] = {0, 0, 0};
    if (keyed) for (int k = 0; k < (colour == 0 ? 1 : 3); ++k) {
        const uint32_t v16 = ((uint32_t)trns[2 * k] << 8) | trns[2 * k + 1];
        key[k] = depth == 16 ? v16 : (v16 & 255u) * (uint32_t)scale_of_depth[depth];
    }
    // one pixel: samples at `s` (sub-byte samples already extracted into s[0]) -> RGBA8 at `d`
    auto pixel = [&](const uint8_t* s, uint8_t* d) {
        const uint32_t f0 = depth == 16 ? ((uint32_t)s[0] << 8) | s[1] : s[0];
        if (colour == 0) {
            const uint8_t g = depth < 8 ? (uint8_t)(s[0] * scale_of_depth[depth]) : s[0];
            d[0] = d[1] = d[2] = g; d[3] = keyed && (depth == 16 ? f0 : (uint32_t)g) == key[0] ? 0 : 255;
        } else if (colour == 2) {
            const uint32_t f1 = depth == 16 ? ((uint32_t)s[2] << 8) | s[3] : s[1], f2 = depth == 16 ? ((uint32_t)s[4] << 8) | s[5] : s[2];
            d[0] = s[0]; d[1] = s[step]; d[2] = s[2 * step]; d[3] = keyed && f0 == key[0] && f1 == key[1] && f2 == key[2] ? 0 : 255;
        } else if (colour == 3) {
            const size_t k = s[0]; const bool in_table = 3 * k + 2 < plte.size();
            d[0] = in_table ? plte[3 * k] : 0; d[1] = in_table ? plte[3 * k + 1] : 0; d[2] = in_table ? plte[3 * k + 2] : 0; d[3] = k < trns.size() ? trns[k] : 255;
        } else if (colour == 4) { d[0] = d[1] = d[2] = s[0]; d[3] = s[step]; }
        else { d[0] = s[0]; d[1] = s[step]; d[2] = s[2 * step]; d[3] = s[3 * step]; }
    };
    for (size_t k = 0; k < passes.size(); ++k) {
        const Pass q = passes[k];
        for (uint32_t y = 0; y < q.h; ++y) {
            const uint8_t* row = img.data() + q.at + (size_t)y * q.stride;
            uint8_t* out = rgba.data() + ((size_t)(q.y0 + y * q.dy) * w + q.x0) * 4;
            for (uint32_t x = 0; x < q.w; ++x) {
                uint8_t sub = 0;
                if (depth < 8) { const size_t bit = (size_t)x * depth; sub = (uint8_t)((row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1)); }
                pixel(depth < 8 ? &sub : row + (size_t)x * bpp, out + (size_t)x * q.dx * 4);
            }
        }
    }
    return true;
}

inline bool write_file(const char* path, const std::vector<uint8_t>& bytes) {
    FILE* fp = fopen(path, "wb");
    if (!fp) return false;
    const bool ok = fwrite(bytes.data(), 1, bytes.size(), fp) == bytes.size();
    return fclose(fp) == 0 && ok;
}

} // namespace png
} // namespace lb
