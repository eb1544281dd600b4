// lb_jpeg.h — JPEG decoder of the asset ingest (baseline and progressive DCT, Huffman coding, 8 bit, 1 / 3 / 4 components, restart intervals).
//
// The reference decodes every glTF image with the stb_image it vendors (stbi_load_from_memory(..., 4),
// LumenPT/src/Tools/LumenPTModelConverter.cpp:105-131; Lumen/vendor/stb/stb_image.h v2.25). A JPEG file does not determine its pixels: the
// inverse DCT, the chroma up-sampling filter and the YCbCr -> RGB conversion are the decoder's choice, and texels feed the renderer's
// parity (surface records are bit-compared). This decoder therefore makes stb_image's choices, so its RGBA8 output is the one the reference
// renders with (tests/test_jpeg.py: every JPEG among the reference's assets against stb_image compiled in place, SHA-256 per file):
//   * entropy decoding and the progressive refinement passes are ITU-T T.81's (Annex F / G) — any conforming decoder yields the same
//     coefficients; coefficients are kept as 16-bit integers, dequantised by a wrapping 16-bit product as stb does;
//   * inverse DCT: the Loeffler-Ligtenberg-Moschytz integer butterfly (the "islow" variant of the IJG library, constants scaled by 2^12),
//     column pass rounded to 2 extra bits (+512 >> 10), row pass to 8 bits with the level shift folded in (+65536 + (128 << 17), >> 17);
//   * chroma up-sampling: the 3:1 triangle filter horizontally (h2v1), vertically (h1v2) or both (h2v2: 9-3-3-1, rounded once), pixel
//     replication for every other sampling ratio;
//   * colour: fixed-point BT.601 full range with 20 fractional bits (coefficients rounded to 12 bits first), the Cb term of green truncated
//     to 16 integer bits before the sum (stb's scalar path is written to agree with its SIMD path, so it is the specification);
//   * 4 components: Adobe CMYK (transform 0) and YCCK (transform 2) through the rounded 8x8 product; 3 components tagged RGB stay RGB.
// Malformed input is an error (false), never a crash: every read is bounds-checked, sizes are capped.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace lb {
namespace jpeg {

inline bool is_jpeg(const uint8_t* p, size_t n) { return n >= 3 && p[0] == 0xFF && p[1] == 0xD8 && p[2] == 0xFF; }

struct Huff {
    uint8_t bits[17] = {0};       // codes per length 1..16
    uint8_t vals[256] = {0};
    int32_t maxcode[18]; int32_t valptr[17]; int32_t mincode[17];
    bool ok = false;
    bool build() {                // T.81 Annex C (code generation) + F.2.2.3 (decoder tables)
        int32_t code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            if (code > (1 << l)) return false;                 // more codes of this length than the prefix space holds
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7FFFFFFF;
        if (k > 256) return false;
        return ok = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int w = 0, hgt = 0;            // samples of this component that belong to the image
    int w2 = 0, h2 = 0;            // allocated plane: whole MCUs
    int dc_pred = 0;
    std::vector<uint8_t> plane;    // w2 x h2 samples
    std::vector<int16_t> coeff;    // progressive: (w2 / 8) x (h2 / 8) blocks of 64, natural order
};

struct Decoder {
    const uint8_t* p; size_t n, pos = 0;
    // entropy-coded segment reader: bits are taken MSB first; a marker ends the segment, after which zeros are fed (as stb does)
    uint32_t bitbuf = 0; int bitcnt = 0; bool hit_marker = false; int marker = -1;
    uint16_t quant[4][64]; bool have_quant[4] = {false, false, false, false};
    Huff hdc[4], hac[4];
    Component comp[4]; int ncomp = 0;
    uint32_t width = 0, height = 0;
    bool progressive = false; int restart_interval = 0;
    int hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
    bool jfif = false; int adobe_transform = -1; int rgb_ids = 0;
    // scan state
    int scan_n = 0, order[4] = {0, 0, 0, 0}, ss = 0, se = 63, ah = 0, al = 0, eobrun = 0, todo = 0;

    Decoder(const uint8_t* data, size_t size) : p(data), n(size) { memset(quant, 0, sizeof quant); }

    int get8() { return pos < n ? p[pos++] : 0; }
    int get16() { const int a = get8(); return (a << 8) | get8(); }
    bool eof() const { return pos >= n; }

    // ---- bit reader
    void fill() {
        while (bitcnt <= 24) {
            int b = 0;
            if (!hit_marker) {
                b = get8();
                if (b == 0xFF) {
                    int c = get8();
                    while (c == 0xFF) c = get8();              // fill bytes
                    if (c != 0) { marker = c; hit_marker = true; b = 0; }
                }
            }
            bitbuf |= (uint32_t)b << (24 - bitcnt);
            bitcnt += 8;
        }
    }
    int bit() { if (bitcnt < 1) fill(); const int b = (int)(bitbuf >> 31); bitbuf <<= 1; --bitcnt; return b; }
    int bits(int k) { if (k == 0) return 0; if (bitcnt < k) fill(); const int v = (int)(bitbuf >> (32 - k)); bitbuf <<= k; bitcnt -= k; return v; }
    // T.81 F.2.2.1 EXTEND
    int receive_extend(int s) { if (s == 0) return 0; const int v = bits(s); return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }
    int decode(const Huff& h) {                                  // T.81 F.2.2.3 DECODE
        if (!h.ok) return -1;
        if (bitcnt < 16) fill();
        int32_t code = 0;
        for (int l = 1; l <= 16; ++l) {
            code = (code << 1) | (int32_t)((bitbuf >> (32 - l)) & 1u);
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) {
                bitbuf <<= l; bitcnt -= l;
                return h.vals[h.valptr[l] + (code - h.mincode[l])];
            }
        }
        return -1;
    }
    void reset_entropy() {
        bitbuf = 0; bitcnt = 0; hit_marker = false; marker = -1; eobrun = 0;
        for (Component& c : comp) c.dc_pred = 0;
        todo = restart_interval ? restart_interval : 0x7FFFFFFF;
    }

    // ---- tables and headers
    int next_marker() {
        if (marker >= 0) { const int m = marker; marker = -1; return m; }
        int x = get8();
        if (x != 0xFF) return -1;
        while (x == 0xFF) x = get8();
        return x;
    }
    bool dqt() {
        int L = get16() - 2;
        while (L > 0) {
            const int q = get8(), prec = q >> 4, t = q & 15;
            if ((prec != 0 && prec != 1) || t > 3) return false;
            static const uint8_t zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                           35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
            for (int i = 0; i < 64; ++i) quant[t][zz[i]] = (uint16_t)(prec ? get16() : get8());
            have_quant[t] = true;
            L -= prec ? 129 : 65;
        }
        return L == 0;
    }
    bool dht() {
        int L = get16() - 2;
        while (L > 0) {
            const int q = get8(), tc = q >> 4, th = q & 15;
            if (tc > 1 || th > 3) return false;
            Huff& h = tc ? hac[th] : hdc[th];
            int total = 0;
            for (int i = 1; i <= 16; ++i) { h.bits[i] = (uint8_t)get8(); total += h.bits[i]; }
            if (total > 256) return false;
            for (int i = 0; i < total; ++i) h.vals[i] = (uint8_t)get8();
            if (!h.build()) return false;
            L -= 17 + total;
        }
        return L == 0;
    }
    bool app_or_com(int m) {
        int L = get16();
        if (L < 2) return false;
        L -= 2;
        if (m == 0xE0 && L >= 5) { static const char tag[5] = {'J', 'F', 'I', 'F', 0}; bool ok = true; for (int i = 0; i < 5; ++i) ok &= get8() == (uint8_t)tag[i]; L -= 5; if (ok) jfif = true; }
        else if (m == 0xEE && L >= 12) {
            static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0}; bool ok = true; for (int i = 0; i < 6; ++i) ok &= get8() == (uint8_t)tag[i]; L -= 6;
            if (ok) { get8(); get16(); get16(); adobe_transform = get8(); L -= 6; }
        }
        if ((size_t)L > n - (pos < n ? pos : n)) { pos = n; return true; }
        pos += (size_t)L;
        return true;
    }
    bool marker_segment(int m) {
        if (m == 0xDD) { if (get16() != 4) return false; restart_interval = get16(); return true; }
        if (m == 0xDB) return dqt();
        if (m == 0xC4) return dht();
        if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) return app_or_com(m);
        return false;
    }
    bool frame_header() {
        const int Lf = get16(); if (Lf < 11) return false;
        if (get8() != 8) return false;                                            // 8-bit samples only (as stb)
        height = (uint32_t)get16(); width = (uint32_t)get16();
        if (!width || !height) return false;
        ncomp = get8();
        if (ncomp != 1 && ncomp != 3 && ncomp != 4) return false;
        if (Lf != 8 + 3 * ncomp) return false;
        if ((uint64_t)width * height > (1ull << 28)) return false;                 // 16384 x 16384
        rgb_ids = 0;
        for (int i = 0; i < ncomp; ++i) {
            Component& c = comp[i];
            c.id = get8();
            static const char rgb[3] = {'R', 'G', 'B'};
            if (ncomp == 3 && c.id == rgb[i]) ++rgb_ids;
            const int q = get8(); c.h = q >> 4; c.v = q & 15;
            if (!c.h || c.h > 4 || !c.v || c.v > 4) return false;
            c.tq = get8(); if (c.tq > 3) return false;
            if (c.h > hmax) hmax = c.h;
            if (c.v > vmax) vmax = c.v;
        }
        mcux = (int)((width + 8u * hmax - 1u) / (8u * hmax)); mcuy = (int)((height + 8u * vmax - 1u) / (8u * vmax));
        for (int i = 0; i < ncomp; ++i) {
            Component& c = comp[i];
            c.w = (int)((width * (uint32_t)c.h + hmax - 1) / hmax); c.hgt = (int)((height * (uint32_t)c.v + vmax - 1) / vmax);
            c.w2 = mcux * c.h * 8; c.h2 = mcuy * c.v * 8;
            c.plane.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) c.coeff.assign((size_t)c.w2 * c.h2, 0);
        }
        return true;
    }
    bool scan_header() {
        const int Ls = get16();
        scan_n = get8();
        if (scan_n < 1 || scan_n > 4 || scan_n > ncomp || Ls != 6 + 2 * scan_n) return false;
        for (int i = 0; i < scan_n; ++i) {
            const int id = get8(), q = get8();
            int which = 0;
            while (which < ncomp && comp[which].id != id) ++which;
            if (which == ncomp) return false;
            comp[which].td = q >> 4; comp[which].ta = q & 15;
            if (comp[which].td > 3 || comp[which].ta > 3) return false;
            order[i] = which;
        }
        ss = get8(); se = get8();
        const int a = get8(); ah = a >> 4; al = a & 15;
        if (progressive) { if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13) return false; }
        else { if (ss != 0 || ah != 0 || al != 0) return false; se = 63; }
        return true;
    }

    // ---- blocks
    static const uint8_t* zigzag() {
        static const uint8_t zz[64 + 16] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                            35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                            63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};      // a corrupt run may step past 63: lands on 63
        return zz;
    }
    bool block_baseline(int16_t* d, Component& c) {
        const uint16_t* q = quant[c.tq]; const uint8_t* zz = zigzag();
        const int t = decode(hdc[c.td]);
        if (t < 0 || t > 15) return false;
        memset(d, 0, 64 * sizeof(int16_t));
        c.dc_pred += receive_extend(t);
        d[0] = (int16_t)(c.dc_pred * q[0]);
        for (int k = 1; k < 64;) {
            const int rs = decode(hac[c.ta]);
            if (rs < 0) return false;
            const int s = rs & 15, r = rs >> 4;
            if (s == 0) { if (rs != 0xF0) break; k += 16; }
            else { k += r; const int z = zz[k++]; d[z] = (int16_t)(receive_extend(s) * q[z]); }
        }
        return true;
    }
    bool block_prog_dc(int16_t* d, Component& c) {
        if (se != 0) return false;
        if (ah == 0) {
            memset(d, 0, 64 * sizeof(int16_t));
            const int t = decode(hdc[c.td]);
            if (t < 0 || t > 15) return false;
            c.dc_pred += receive_extend(t);
            d[0] = (int16_t)(c.dc_pred * (1 << al));
        } else if (bit()) d[0] = (int16_t)(d[0] + (int16_t)(1 << al));
        return true;
    }
    bool block_prog_ac(int16_t* d, Component& c) {
        if (ss == 0) return false;
        const uint8_t* zz = zigzag();
        if (ah == 0) {                                                            // first pass over the band (T.81 G.1.2.2)
            if (eobrun) { --eobrun; return true; }
            for (int k = ss; k <= se;) {
                const int rs = decode(hac[c.ta]);
                if (rs < 0) return false;
                const int s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) { eobrun = (1 << r); if (r) eobrun += bits(r); --eobrun; break; }
                    k += 16;
                } else { k += r; const int z = zz[k++]; d[z] = (int16_t)(receive_extend(s) * (1 << al)); }
            }
            return true;
        }
        const int16_t one = (int16_t)(1 << al);                                   // refinement pass (G.1.2.3)
        auto refine = [&](int16_t& v) { if (bit() && (v & one) == 0) v = (int16_t)(v > 0 ? v + one : v - one); };
        if (eobrun) {
            --eobrun;
            for (int k = ss; k <= se; ++k) { int16_t& v = d[zz[k]]; if (v != 0) refine(v); }
            return true;
        }
        for (int k = ss; k <= se;) {
            const int rs = decode(hac[c.ta]);
            if (rs < 0) return false;
            int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (r < 15) { eobrun = (1 << r) - 1; if (r) eobrun += bits(r); r = 64; }      // end of band: only the non-zero history is refined
            } else {
                if (s != 1) return false;
                s = bit() ? one : -one;
            }
            while (k <= se) {
                int16_t& v = d[zz[k++]];
                if (v != 0) refine(v);
                else { if (r == 0) { v = (int16_t)s; break; } --r; }
            }
        }
        return true;
    }

    // ---- inverse DCT (LLM, 12-bit constants), 8x8 block of dequantised coefficients -> 8-bit samples
    static int fx(double x) { return (int)(x * 4096 + 0.5); }
    // 64-bit intermediates: identical to stb's `int` arithmetic for every stream a conforming encoder writes (|coefficient x quantiser| stays
    // far inside 32 bits), and defined — instead of a signed overflow — for the coefficients a corrupt stream can hold (found by the fuzzer)
    typedef long long i64;
    struct Idct1 { i64 x0, x1, x2, x3, t0, t1, t2, t3; };
    static Idct1 idct1(i64 s0, i64 s1, i64 s2, i64 s3, i64 s4, i64 s5, i64 s6, i64 s7) {
        static const i64 c0541 = fx(0.5411961f), c1847 = fx(-1.847759065f), c0765 = fx(0.765366865f), c1175 = fx(1.175875602f), c0298 = fx(0.298631336f),
                         c2053 = fx(2.053119869f), c3072 = fx(3.072711026f), c1501 = fx(1.501321110f), c0899 = fx(-0.899976223f), c2562 = fx(-2.562915447f),
                         c1961 = fx(-1.961570560f), c0390 = fx(-0.390180644f);
        Idct1 o;
        i64 p1 = (s2 + s6) * c0541;
        const i64 e2 = p1 + s6 * c1847, e3 = p1 + s2 * c0765;
        const i64 e0 = (s0 + s4) * 4096, e1 = (s0 - s4) * 4096;
        o.x0 = e0 + e3; o.x3 = e0 - e3; o.x1 = e1 + e2; o.x2 = e1 - e2;
        i64 t0 = s7, t1 = s5, t2 = s3, t3 = s1;
        i64 p3 = t0 + t2, p4 = t1 + t3; p1 = t0 + t3; i64 p2 = t1 + t2;
        const i64 p5 = (p3 + p4) * c1175;
        t0 *= c0298; t1 *= c2053; t2 *= c3072; t3 *= c1501;
        p1 = p5 + p1 * c0899; p2 = p5 + p2 * c2562; p3 *= c1961; p4 *= c0390;
        o.t3 = t3 + p1 + p4; o.t2 = t2 + p2 + p3; o.t1 = t1 + p2 + p4; o.t0 = t0 + p1 + p3;
        return o;
    }
    static uint8_t clamp8(i64 x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }
    static int wrap32(i64 x) { return (int)(uint32_t)(unsigned long long)x; }      // what stb's int would hold (two's complement)
    static void idct(uint8_t* out, int stride, const int16_t* d) {
        int v[64];
        for (int i = 0; i < 8; ++i) {                                             // columns
            const int16_t* c = d + i;
            if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
                const int dc = c[0] * 4;
                for (int r = 0; r < 8; ++r) v[r * 8 + i] = dc;
                continue;
            }
            Idct1 o = idct1(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]);
            o.x0 += 512; o.x1 += 512; o.x2 += 512; o.x3 += 512;
            v[0 * 8 + i] = wrap32((o.x0 + o.t3) >> 10); v[7 * 8 + i] = wrap32((o.x0 - o.t3) >> 10);
            v[1 * 8 + i] = wrap32((o.x1 + o.t2) >> 10); v[6 * 8 + i] = wrap32((o.x1 - o.t2) >> 10);
            v[2 * 8 + i] = wrap32((o.x2 + o.t1) >> 10); v[5 * 8 + i] = wrap32((o.x2 - o.t1) >> 10);
            v[3 * 8 + i] = wrap32((o.x3 + o.t0) >> 10); v[4 * 8 + i] = wrap32((o.x3 - o.t0) >> 10);
        }
        for (int r = 0; r < 8; ++r, out += stride) {                              // rows
            const int* w = v + r * 8;
            Idct1 o = idct1(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
            const i64 bias = 65536 + (128 << 17);
            o.x0 += bias; o.x1 += bias; o.x2 += bias; o.x3 += bias;
            out[0] = clamp8((o.x0 + o.t3) >> 17); out[7] = clamp8((o.x0 - o.t3) >> 17);
            out[1] = clamp8((o.x1 + o.t2) >> 17); out[6] = clamp8((o.x1 - o.t2) >> 17);
            out[2] = clamp8((o.x2 + o.t1) >> 17); out[5] = clamp8((o.x2 - o.t1) >> 17);
            out[3] = clamp8((o.x3 + o.t0) >> 17); out[4] = clamp8((o.x3 - o.t0) >> 17);
        }
    }

    // ---- one scan
    bool restart_check() {                 // after every MCU: true = go on, sets `stop` when the segment ended without a restart marker
        if (--todo > 0) return true;
        if (bitcnt < 24) fill();
        if (!(marker >= 0xD0 && marker <= 0xD7)) { stop = true; return true; }
        reset_entropy();
        return true;
    }
    bool stop = false;
    bool scan() {
        reset_entropy(); stop = false;
        int16_t blk[64];
        if (scan_n == 1) {                                                        // non-interleaved: the component's own block grid
            Component& c = comp[order[0]];
            const int bw = (c.w + 7) >> 3, bh = (c.hgt + 7) >> 3;
            for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
                if (!progressive) {
                    if (!have_quant[c.tq] && false) return false;
                    if (!block_baseline(blk, c)) return false;
                    idct(c.plane.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, blk);
                } else {
                    int16_t* d = c.coeff.data() + 64 * ((size_t)i + (size_t)j * (c.w2 / 8));
                    if (!(ss == 0 ? block_prog_dc(d, c) : block_prog_ac(d, c))) return false;
                }
                restart_check(); if (stop) return true;
            }
            return true;
        }
        for (int j = 0; j < mcuy; ++j) for (int i = 0; i < mcux; ++i) {           // interleaved MCUs
            for (int k = 0; k < scan_n; ++k) {
                Component& c = comp[order[k]];
                for (int y = 0; y < c.v; ++y) for (int x = 0; x < c.h; ++x) {
                    const int bx = i * c.h + x, by = j * c.v + y;
                    if (!progressive) {
                        if (!block_baseline(blk, c)) return false;
                        idct(c.plane.data() + (size_t)c.w2 * by * 8 + bx * 8, c.w2, blk);
                    } else {
                        if (ss != 0) return false;                                  // AC scans are never interleaved
                        if (!block_prog_dc(c.coeff.data() + 64 * ((size_t)bx + (size_t)by * (c.w2 / 8)), c)) return false;
                    }
                }
            }
            restart_check(); if (stop) return true;
        }
        return true;
    }
    void finish_progressive() {
        for (int n_ = 0; n_ < ncomp; ++n_) {
            Component& c = comp[n_];
            const int bw = (c.w + 7) >> 3, bh = (c.hgt + 7) >> 3; const uint16_t* q = quant[c.tq];
            for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
                int16_t* d = c.coeff.data() + 64 * ((size_t)i + (size_t)j * (c.w2 / 8));
                for (int k = 0; k < 64; ++k) d[k] = (int16_t)(d[k] * q[k]);
                idct(c.plane.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, d);
            }
        }
    }
    bool decode_image() {
        if (next_marker() != 0xD8) return false;
        int m = next_marker();
        while (!(m == 0xC0 || m == 0xC1 || m == 0xC2)) {
            if (m < 0) { if (eof()) return false; m = next_marker(); continue; }
            if (!marker_segment(m)) return false;
            m = next_marker();
            while (m < 0) { if (eof()) return false; m = next_marker(); }
        }
        progressive = m == 0xC2;
        if (!frame_header()) return false;
        m = next_marker();
        int guard = 0;
        while (m != 0xD9) {
            if (++guard > 100000) return false;
            if (m == 0xDA) {
                if (!scan_header() || !scan()) return false;
                if (marker < 0) {                                                  // the scan did not end on a marker: look for the next one
                    while (!eof()) { if (get8() == 0xFF) { marker = get8(); break; } }
                    if (marker < 0) break;                                         // stream ends without EOI: use what was decoded
                }
            } else if (m == 0xDC) { if (get16() != 4) return false; if ((uint32_t)get16() != height) return false; }
            else if (m < 0) { if (eof()) break; }
            else if (!marker_segment(m)) return false;
            m = next_marker();
        }
        if (progressive) finish_progressive();
        return true;
    }
};

// ---- up-sampling of one component row to full width (each returns the row to read: `buf` or the input itself)
inline const uint8_t* up_row(uint8_t* buf, const uint8_t* near_, const uint8_t* far_, int w, int hs, int vs) {
    if (hs == 1 && vs == 1) return near_;
    if (hs == 1 && vs == 2) { for (int i = 0; i < w; ++i) buf[i] = (uint8_t)((3 * near_[i] + far_[i] + 2) >> 2); return buf; }
    if (hs == 2 && vs == 1) {
        if (w == 1) { buf[0] = buf[1] = near_[0]; return buf; }
        buf[0] = near_[0]; buf[1] = (uint8_t)((near_[0] * 3 + near_[1] + 2) >> 2);
        int i = 1;
        for (; i < w - 1; ++i) { const int t = 3 * near_[i] + 2; buf[i * 2] = (uint8_t)((t + near_[i - 1]) >> 2); buf[i * 2 + 1] = (uint8_t)((t + near_[i + 1]) >> 2); }
        buf[i * 2] = (uint8_t)((near_[w - 2] * 3 + near_[w - 1] + 2) >> 2); buf[i * 2 + 1] = near_[w - 1];
        return buf;
    }
    if (hs == 2 && vs == 2) {
        if (w == 1) { buf[0] = buf[1] = (uint8_t)((3 * near_[0] + far_[0] + 2) >> 2); return buf; }
        int t1 = 3 * near_[0] + far_[0];
        buf[0] = (uint8_t)((t1 + 2) >> 2);
        for (int i = 1; i < w; ++i) {
            const int t0 = t1; t1 = 3 * near_[i] + far_[i];
            buf[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4); buf[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
        }
        buf[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
        return buf;
    }
    for (int i = 0; i < w; ++i) for (int j = 0; j < hs; ++j) buf[i * hs + j] = near_[i];
    return buf;
}

inline uint8_t mul8(uint8_t x, uint8_t y) { const unsigned t = (unsigned)x * y + 128u; return (uint8_t)((t + (t >> 8)) >> 8); }

inline void ycc_to_rgba(uint8_t* out, const uint8_t* y, const uint8_t* cb, const uint8_t* cr, uint32_t count) {
    const int kr = ((int)(1.40200f * 4096.0f + 0.5f)) << 8, kg1 = ((int)(0.71414f * 4096.0f + 0.5f)) << 8, kg2 = ((int)(0.34414f * 4096.0f + 0.5f)) << 8,
              kb = ((int)(1.77200f * 4096.0f + 0.5f)) << 8;
    for (uint32_t i = 0; i < count; ++i, out += 4) {
        const int yf = (y[i] << 20) + (1 << 19), r_ = cr[i] - 128, b_ = cb[i] - 128;
        int r = yf + r_ * kr;
        int g = yf + (r_ * -kg1) + (int)((unsigned)(b_ * -kg2) & 0xffff0000u);
        int b = yf + b_ * kb;
        r >>= 20; g >>= 20; b >>= 20;
        out[0] = Decoder::clamp8(r); out[1] = Decoder::clamp8(g); out[2] = Decoder::clamp8(b); out[3] = 255;
    }
}

// RGBA8 pixels of a JPEG file, as stbi_load_from_memory(..., 4) returns them. false = not a JPEG this decoder handles / corrupt.
inline bool decode_rgba8(const uint8_t* file, size_t n, std::vector<uint8_t>& rgba, uint32_t& w, uint32_t& h) {
    if (!is_jpeg(file, n)) return false;
    Decoder d(file, n);
    if (!d.decode_image()) return false;
    w = d.width; h = d.height;
    rgba.assign((size_t)w * h * 4, 255);
    const bool is_rgb = d.ncomp == 3 && (d.rgb_ids == 3 || (d.adobe_transform == 0 && !d.jfif));
    struct Up { int hs, vs, ystep, wl, ypos; const uint8_t* line0; const uint8_t* line1; std::vector<uint8_t> buf; } up[4];
    for (int k = 0; k < d.ncomp; ++k) {
        Component& c = d.comp[k];
        up[k].hs = d.hmax / c.h; up[k].vs = d.vmax / c.v; up[k].ystep = up[k].vs >> 1; up[k].wl = (int)((w + up[k].hs - 1) / up[k].hs); up[k].ypos = 0;
        up[k].line0 = up[k].line1 = c.plane.data();
        up[k].buf.assign((size_t)up[k].wl * up[k].hs + 8, 0);
    }
    const uint8_t* row[4] = {nullptr, nullptr, nullptr, nullptr};
    for (uint32_t j = 0; j < h; ++j) {
        uint8_t* out = rgba.data() + (size_t)j * w * 4;
        for (int k = 0; k < d.ncomp; ++k) {
            Up& u = up[k]; Component& c = d.comp[k];
            const bool bot = u.ystep >= (u.vs >> 1);
            row[k] = up_row(u.buf.data(), bot ? u.line1 : u.line0, bot ? u.line0 : u.line1, u.wl, u.hs, u.vs);
            if (++u.ystep >= u.vs) { u.ystep = 0; u.line0 = u.line1; if (++u.ypos < c.hgt) u.line1 += c.w2; }
        }
        if (d.ncomp == 3) {
            if (is_rgb) for (uint32_t i = 0; i < w; ++i) { out[4 * i] = row[0][i]; out[4 * i + 1] = row[1][i]; out[4 * i + 2] = row[2][i]; out[4 * i + 3] = 255; }
            else ycc_to_rgba(out, row[0], row[1], row[2], w);
        } else if (d.ncomp == 4) {
            if (d.adobe_transform == 0) for (uint32_t i = 0; i < w; ++i) { const uint8_t m = row[3][i]; out[4 * i] = mul8(row[0][i], m); out[4 * i + 1] = mul8(row[1][i], m); out[4 * i + 2] = mul8(row[2][i], m); out[4 * i + 3] = 255; }
            else {
                ycc_to_rgba(out, row[0], row[1], row[2], w);
                if (d.adobe_transform == 2) for (uint32_t i = 0; i < w; ++i) { const uint8_t m = row[3][i]; out[4 * i] = mul8(255 - out[4 * i], m); out[4 * i + 1] = mul8(255 - out[4 * i + 1], m); out[4 * i + 2] = mul8(255 - out[4 * i + 2], m); }
            }
        } else for (uint32_t i = 0; i < w; ++i) { out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = row[0][i]; out[4 * i + 3] = 255; }
    }
    return true;
}

} // namespace jpeg
} // namespace lb
