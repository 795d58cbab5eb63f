// jpeg_io.hpp -- baseline JPEG reader for the CLI (the reference reads its inputs with Imgcodecs.imread,
// PictureService.java:107-131; its own sample images album.jpg / haha.jpg / hkp.jpg are baseline 4:2:0 JPEGs).
//
// Scope: baseline sequential DCT (SOF0), Huffman coding, 8-bit samples, 1 (gray) or 3 (YCbCr) components, sampling factors
// 1x1 / 2x1 / 1x2 / 2x2 for the luma component with 1x1 chroma, restart intervals.  Not supported: progressive / arithmetic /
// lossless / 12-bit / CMYK files (read_bgr returns false; the CLI then reports the file as unreadable, as imread would for a
// corrupt file).
//
// The arithmetic is the integer pipeline every libjpeg-derived decoder (and therefore imread) uses by default, restated from
// its published description so that the pixels equal imread's bit for bit:
//   * inverse DCT: the "accurate integer" Loeffler-Ligtenberg-Moschytz 8x8 algorithm with 13-bit constants and 2 extra
//     bits kept between the column and the row pass;
//   * chroma upsampling: the "fancy" triangle filter (3/4 nearer + 1/4 farther sample, in each direction, alternating
//     rounding), edge samples replicated; chroma planes of at most two columns are replicated instead, as libjpeg does;
//   * colour conversion: R = Y + 1.40200 Cr, G = Y - 0.34414 Cb - 0.71414 Cr, B = Y + 1.77200 Cb in 16.16 fixed point.
// Checked against cv2.imread on generated fixtures (tests/golden/jpeg/) and on the reference's three images
// (tests/test_cli.py).  Host-side I/O only: nothing here runs on the segmentation path.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

namespace msegment {
namespace jpeg {

struct HuffTable {
    bool present = false;
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17] = {0}, maxcode[18] = {0}, valptr[17] = {0};
    void build()
    {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l];
            k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int wblocks = 0, hblocks = 0;         // allocated size in blocks (whole MCUs)
    int dw = 0, dh = 0;                   // real (downsampled) size in samples
    int pred = 0;
    std::vector<uint8_t> plane;           // wblocks*8 x hblocks*8 samples
};

struct Reader {
    const uint8_t* p;
    size_t n, pos = 0;
    uint32_t acc = 0;
    int nbits = 0;
    bool hit_marker = false;
    int marker = 0;
    explicit Reader(const std::vector<uint8_t>& d) : p(d.data()), n(d.size()) {}
    int byte() { return pos < n ? p[pos++] : -1; }
    int u16() { int a = byte(), b = byte(); return (a < 0 || b < 0) ? -1 : (a << 8) | b; }
    // entropy-coded segment: 0xFF00 is a stuffed 0xFF, any other 0xFFxx is a marker (bits are then padded with zeros)
    void fill()
    {
        while (nbits <= 24) {
            int b = 0;
            if (!hit_marker) {
                if (pos >= n) { hit_marker = true; marker = 0xD9; }
                else {
                    b = p[pos++];
                    if (b == 0xFF) {
                        int m = pos < n ? p[pos] : 0xD9;
                        if (m == 0) pos++;
                        else { hit_marker = true; marker = m; pos++; b = 0; }
                    }
                }
            }
            acc |= (uint32_t)b << (24 - nbits);
            nbits += 8;
        }
    }
    int bits(int k)
    {
        if (k == 0) return 0;
        if (nbits < k) fill();
        int v = (int)(acc >> (32 - k));
        acc <<= k;
        nbits -= k;
        return v;
    }
    void reset_bits() { acc = 0; nbits = 0; }
};

inline int decode_symbol(Reader& r, const HuffTable& t)
{
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | r.bits(1);
        if (t.maxcode[l] >= 0 && code <= t.maxcode[l] && code >= t.mincode[l]) return t.vals[t.valptr[l] + code - t.mincode[l]];
    }
    return -1;
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// accurate integer inverse DCT of one dequantised block -> 64 samples (0..255), row-major
inline void idct_islow(const int* coef, uint8_t* out, int stride)
{
    constexpr int CB = 13, P1 = 2;
    constexpr long F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299,
                   F1847 = 15137, F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
    long ws[64];
    auto descale = [](long x, int s) { return (x + (1L << (s - 1))) >> s; };
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < 8; i++) {
            long in[8];
            for (int k = 0; k < 8; k++) in[k] = pass == 0 ? coef[k * 8 + i] : ws[i * 8 + k];
            long z2 = in[2], z3 = in[6];
            long z1 = (z2 + z3) * F0541;
            long tmp2 = z1 + z3 * -F1847;
            long tmp3 = z1 + z2 * F0765;
            z2 = in[0]; z3 = in[4];
            long tmp0 = (z2 + z3) * (1L << CB);
            long tmp1 = (z2 - z3) * (1L << CB);
            long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
            z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
            long z4 = tmp1 + tmp3;
            long z5 = (z3 + z4) * F1175;
            tmp0 *= F0298; tmp1 *= F2053; tmp2 *= F3072; tmp3 *= F1501;
            z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            const long o[8] = {tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2,
                               tmp10 - tmp3};
            if (pass == 0) {
                for (int k = 0; k < 8; k++) ws[k * 8 + i] = descale(o[k], CB - P1);
            } else {
                for (int k = 0; k < 8; k++) {
                    long v = descale(o[k], CB + P1 + 3) + 128;
                    out[i * stride + k] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
                }
            }
        }
    }
}

inline bool read_bgr(const std::string& path, std::vector<uint8_t>& bgr, int& width, int& height)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::vector<uint8_t> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (data.size() < 4 || data[0] != 0xFF || data[1] != 0xD8) return false;
    static const int zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    Reader r(data);
    r.pos = 2;
    int qt[4][64] = {{0}};
    bool have_q[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    std::vector<Component> comp;
    int restart = 0, hmax = 1, vmax = 1;
    bool have_sof = false;
    for (;;) {
        int b = r.byte();
        if (b < 0) return false;
        if (b != 0xFF) continue;
        int m = r.byte();
        while (m == 0xFF) m = r.byte();
        if (m < 0) return false;
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) return false;                                    // end of image before any scan
        int len = r.u16();
        if (len < 2 || r.pos + (size_t)len - 2 > r.n) return false;
        size_t end = r.pos + (size_t)len - 2;
        if (m == 0xDB) {                                                // quantisation tables
            while (r.pos < end) {
                int pq = r.byte();
                int prec = pq >> 4, id = pq & 15;
                if (id > 3) return false;
                for (int i = 0; i < 64; i++) {
                    int v = prec ? r.u16() : r.byte();
                    if (v < 0) return false;
                    qt[id][zigzag[i]] = v;
                }
                have_q[id] = true;
            }
        } else if (m == 0xC4) {                                         // Huffman tables
            while (r.pos < end) {
                int tc = r.byte();
                int cls = tc >> 4, id = tc & 15;
                if (cls > 1 || id > 3) return false;
                HuffTable& t = cls ? ac[id] : dc[id];
                int total = 0;
                t.bits[0] = 0;
                for (int l = 1; l <= 16; l++) { t.bits[l] = (uint8_t)r.byte(); total += t.bits[l]; }
                if (total > 256) return false;
                for (int i = 0; i < total; i++) t.vals[i] = (uint8_t)r.byte();
                t.present = true;
                t.build();
            }
        } else if (m == 0xC0 || m == 0xC1) {                            // baseline / extended sequential, Huffman
            int prec = r.byte();
            height = r.u16();
            width = r.u16();
            int nc = r.byte();
            if (prec != 8 || width <= 0 || height <= 0 || (nc != 1 && nc != 3)) return false;
            comp.assign(nc, Component());
            for (int i = 0; i < nc; i++) {
                comp[i].id = r.byte();
                int hv = r.byte();
                comp[i].h = hv >> 4; comp[i].v = hv & 15;
                comp[i].tq = r.byte();
                if (comp[i].h < 1 || comp[i].h > 2 || comp[i].v < 1 || comp[i].v > 2 || comp[i].tq > 3) return false;
            }
            if (nc == 3 && (comp[1].h != 1 || comp[1].v != 1 || comp[2].h != 1 || comp[2].v != 1)) return false;
            if (nc == 1) { comp[0].h = 1; comp[0].v = 1; }              // a single component is never interleaved
            hmax = comp[0].h; vmax = comp[0].v;
            have_sof = true;
        } else if (m == 0xC2 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            return false;                                               // progressive / lossless / arithmetic: not supported
        } else if (m == 0xDD) {
            restart = r.u16();
        } else if (m == 0xDA) {                                         // start of scan: one interleaved scan expected
            if (!have_sof) return false;
            int ns = r.byte();
            if (ns != (int)comp.size()) return false;
            for (int i = 0; i < ns; i++) {
                int id = r.byte(), tt = r.byte();
                bool found = false;
                for (auto& c : comp)
                    if (c.id == id) { c.td = tt >> 4; c.ta = tt & 15; found = true; }
                if (!found) return false;
            }
            r.pos = end;
            break;
        }
        r.pos = end;
    }
    for (auto& c : comp)
        if (!have_q[c.tq] || c.td > 3 || c.ta > 3 || !dc[c.td].present || !ac[c.ta].present) return false;
    const int mcu_w = 8 * hmax, mcu_h = 8 * vmax;
    const int mcus_x = (width + mcu_w - 1) / mcu_w, mcus_y = (height + mcu_h - 1) / mcu_h;
    for (auto& c : comp) {
        c.wblocks = mcus_x * c.h;
        c.hblocks = mcus_y * c.v;
        c.dw = (width * c.h + hmax - 1) / hmax;
        c.dh = (height * c.v + vmax - 1) / vmax;
        c.plane.assign((size_t)c.wblocks * 8 * c.hblocks * 8, 0);
    }
    // ---- entropy decoding + inverse DCT, MCU by MCU
    r.reset_bits();
    int rst_left = restart;
    for (int my = 0; my < mcus_y; my++)
        for (int mx = 0; mx < mcus_x; mx++) {
            if (restart && rst_left == 0) {                             // restart marker: byte-align, reset the predictors
                r.reset_bits();
                if (!r.hit_marker) {                                    // the marker has not been consumed by the bit reader yet
                    while (r.pos + 1 < r.n && !(r.p[r.pos] == 0xFF && r.p[r.pos + 1] >= 0xD0 && r.p[r.pos + 1] <= 0xD7)) r.pos++;
                    r.pos += 2;
                }
                r.hit_marker = false;
                for (auto& c : comp) c.pred = 0;
                rst_left = restart;
            }
            for (auto& c : comp)
                for (int by = 0; by < c.v; by++)
                    for (int bx = 0; bx < c.h; bx++) {
                        int coef[64] = {0};
                        int s = decode_symbol(r, dc[c.td]);
                        if (s < 0 || s > 11) return false;
                        int diff = s ? extend(r.bits(s), s) : 0;
                        c.pred += diff;
                        coef[0] = c.pred * qt[c.tq][0];
                        for (int k = 1; k < 64;) {
                            int rs = decode_symbol(r, ac[c.ta]);
                            if (rs < 0) return false;
                            int run = rs >> 4, size = rs & 15;
                            if (size == 0) {
                                if (run == 15) { k += 16; continue; }
                                break;                                  // end of block
                            }
                            k += run;
                            if (k > 63) return false;
                            coef[zigzag[k]] = extend(r.bits(size), size) * qt[c.tq][zigzag[k]];
                            k++;
                        }
                        const int px = (mx * c.h + bx) * 8, py = (my * c.v + by) * 8;
                        idct_islow(coef, c.plane.data() + (size_t)py * c.wblocks * 8 + px, c.wblocks * 8);
                    }
            if (restart) rst_left--;
        }
    // ---- upsampling + colour conversion
    bgr.assign((size_t)width * height * 3, 0);
    if (comp.size() == 1) {
        const Component& y = comp[0];
        for (int i = 0; i < height; i++)
            for (int j = 0; j < width; j++) {
                uint8_t v = y.plane[(size_t)i * y.wblocks * 8 + j];
                uint8_t* o = &bgr[((size_t)i * width + j) * 3];
                o[0] = o[1] = o[2] = v;
            }
        return true;
    }
    // chroma planes at full resolution ("fancy" triangle filter; edge rows / columns replicated)
    std::vector<uint8_t> up[2];
    for (int ci = 1; ci <= 2; ci++) {
        const Component& c = comp[ci];
        const int pitch = c.wblocks * 8;
        std::vector<uint8_t>& o = up[ci - 1];
        o.assign((size_t)width * height, 0);
        auto S = [&](int yy, int xx) -> int {
            yy = yy < 0 ? 0 : (yy >= c.dh ? c.dh - 1 : yy);
            return c.plane[(size_t)yy * pitch + xx];
        };
        if (hmax == 1 && vmax == 1) {
            for (int i = 0; i < height; i++) memcpy(&o[(size_t)i * width], &c.plane[(size_t)i * pitch], (size_t)width);
        } else if (hmax == 2 && c.dw <= 2) {                            // too narrow for the triangle filter: plain replication
            for (int i = 0; i < height; i++)
                for (int j = 0; j < width; j++) o[(size_t)i * width + j] = (uint8_t)S(vmax == 2 ? i >> 1 : i, j >> 1);
        } else if (hmax == 2 && vmax == 1) {                            // h2v1: 3/4 + 1/4 horizontally
            for (int i = 0; i < height; i++)
                for (int x = 0; x < c.dw; x++) {
                    int cur = S(i, x);
                    int l = x > 0 ? S(i, x - 1) : cur, rr = x + 1 < c.dw ? S(i, x + 1) : cur;
                    int a = x == 0 ? cur : (3 * cur + l + 1) >> 2;
                    int b = x + 1 == c.dw ? cur : (3 * cur + rr + 2) >> 2;
                    if (2 * x < width) o[(size_t)i * width + 2 * x] = (uint8_t)a;
                    if (2 * x + 1 < width) o[(size_t)i * width + 2 * x + 1] = (uint8_t)b;
                }
        } else if (hmax == 1 && vmax == 2) {                            // h1v2: 3/4 + 1/4 vertically
            for (int i = 0; i < height; i++) {
                int yy = i >> 1, other = (i & 1) ? yy + 1 : yy - 1, bias = (i & 1) ? 2 : 1;
                for (int x = 0; x < width; x++) o[(size_t)i * width + x] = (uint8_t)((3 * S(yy, x) + S(other, x) + bias) >> 2);
            }
        } else {                                                        // h2v2
            for (int i = 0; i < height; i++) {
                int yy = i >> 1, other = (i & 1) ? yy + 1 : yy - 1;
                auto colsum = [&](int x) { return 3 * S(yy, x) + S(other, x); };
                for (int x = 0; x < c.dw; x++) {
                    int cur = colsum(x);
                    int a = x == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + colsum(x - 1) + 8) >> 4;
                    int b = x + 1 == c.dw ? (cur * 4 + 7) >> 4 : (cur * 3 + colsum(x + 1) + 7) >> 4;
                    if (2 * x < width) o[(size_t)i * width + 2 * x] = (uint8_t)a;
                    if (2 * x + 1 < width) o[(size_t)i * width + 2 * x + 1] = (uint8_t)b;
                }
            }
        }
    }
    // YCbCr -> BGR, 16.16 fixed point tables
    int cr_r[256], cb_b[256];
    long cr_g[256], cb_g[256];
    for (int i = 0; i < 256; i++) {
        long x = i - 128;
        cr_r[i] = (int)((91881L * x + 32768) >> 16);                    // FIX(1.40200)
        cb_b[i] = (int)((116130L * x + 32768) >> 16);                   // FIX(1.77200)
        cr_g[i] = -46802L * x;                                          // FIX(0.71414)
        cb_g[i] = -22554L * x + 32768;                                  // FIX(0.34414) + ONE_HALF
    }
    auto clamp = [](int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
    const Component& Y = comp[0];
    for (int i = 0; i < height; i++)
        for (int j = 0; j < width; j++) {
            int y = Y.plane[(size_t)i * Y.wblocks * 8 + j];
            int cb = up[0][(size_t)i * width + j], cr = up[1][(size_t)i * width + j];
            uint8_t* o = &bgr[((size_t)i * width + j) * 3];
            o[2] = clamp(y + cr_r[cr]);
            o[1] = clamp(y + (int)((cb_g[cb] + cr_g[cr]) >> 16));
            o[0] = clamp(y + cb_b[cb]);
        }
    return true;
}

}  // namespace jpeg
}  // namespace msegment
