// png_io.hpp -- minimal PNG codec for the console program (no zlib/libpng in the image).
//
// The reference writes its batch with Imgcodecs.imwrite(<...>.png) (PictureService.java:209-216) and reads its inputs with
// Imgcodecs.imread (PictureService.java:94-96; resources/images/guide.png is an 8-bit palette PNG).  File IO is outside
// the hot path; this header only makes the CLI's file formats match the reference's.
//   write: 8-bit gray or RGB, filter 0, zlib "stored" blocks (valid PNG, no compression).
//   read : non-interlaced PNG, bit depth 8 (gray, gray+alpha, RGB, RGBA) or palette (1/2/4/8 bit); alpha is dropped, as
//          imread's default IMREAD_COLOR does.  Full inflate (stored, fixed and dynamic Huffman blocks).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace msegment {
namespace png {

inline uint32_t crc32(const uint8_t* p, size_t n, uint32_t crc = 0)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}

inline void put32(std::vector<uint8_t>& v, uint32_t x)
{
    v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x));
}

inline void chunk(std::vector<uint8_t>& out, const char* tag, const std::vector<uint8_t>& body)
{
    put32(out, (uint32_t)body.size());
    size_t at = out.size();
    out.insert(out.end(), tag, tag + 4);
    out.insert(out.end(), body.begin(), body.end());
    put32(out, crc32(out.data() + at, out.size() - at));
}

// pixels: rows of `channels` (1 = gray, 3 = RGB) bytes, `stride` bytes apart.
inline bool write(const std::string& path, const uint8_t* pixels, size_t stride, int w, int h, int channels)
{
    if (w <= 0 || h <= 0 || (channels != 1 && channels != 3)) return false;
    const size_t row = (size_t)w * channels + 1;
    std::vector<uint8_t> raw(row * h);
    for (int y = 0; y < h; y++) {
        raw[y * row] = 0;                                               // filter type None
        memcpy(&raw[y * row + 1], pixels + (size_t)y * stride, row - 1);
    }
    std::vector<uint8_t> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);                               // zlib header, no preset dictionary
    uint32_t a = 1, b = 0;                                              // adler32
    for (size_t off = 0; off < raw.size() || off == 0;) {
        size_t n = raw.size() - off;
        if (n > 65535) n = 65535;
        z.push_back(off + n == raw.size() ? 1 : 0);                     // BFINAL, BTYPE=00 (stored)
        z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8));
        z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for (size_t i = off; i < off + n; i++) { a += raw[i]; if (a >= 65521) a -= 65521; b += a; if (b >= 65521) b -= 65521; }
        off += n;
        if (n == 0) break;
    }
    put32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)w); put32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(channels == 3 ? 2 : 0); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    return fclose(f) == 0 && ok;
}

// ---- inflate (RFC 1951) -------------------------------------------------------------------------------------------
struct BitReader {
    const uint8_t* p; size_t n, pos = 0; uint32_t acc = 0; int cnt = 0; bool bad = false;
    BitReader(const uint8_t* d, size_t len) : p(d), n(len) {}
    uint32_t bits(int k)
    {
        while (cnt < k) {
            if (pos >= n) { bad = true; return 0; }
            acc |= (uint32_t)p[pos++] << cnt; cnt += 8;
        }
        uint32_t v = k ? acc & ((1u << k) - 1) : 0;
        acc >>= k; cnt -= k;
        return v;
    }
    void align() { acc = 0; cnt = 0; }
};

struct Huffman {
    uint16_t count[16], symbol[288];
    void build(const uint8_t* len, int n)
    {
        memset(count, 0, sizeof(count));
        for (int i = 0; i < n; i++) count[len[i]]++;
        count[0] = 0;
        uint16_t offs[16]; offs[1] = 0;
        for (int i = 1; i < 15; i++) offs[i + 1] = offs[i] + count[i];
        for (int i = 0; i < n; i++) if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    }
    int decode(BitReader& br) const
    {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; l++) {
            code |= (int)br.bits(1);
            if (br.bad) return -1;
            int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};

inline bool inflate(const uint8_t* src, size_t n, std::vector<uint8_t>& out)
{
    static const uint16_t lbase[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
    static const uint16_t lext[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
    static const uint16_t dbase[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,
                                       8193,12289,16385,24577};
    static const uint16_t dext[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
    static const uint8_t order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
    BitReader br(src, n);
    for (int last = 0; !last;) {
        last = (int)br.bits(1);
        int type = (int)br.bits(2);
        if (br.bad) return false;
        if (type == 0) {
            br.align();
            if (br.pos + 4 > n) return false;
            uint32_t len = src[br.pos] | (src[br.pos + 1] << 8), nlen = src[br.pos + 2] | (src[br.pos + 3] << 8);
            br.pos += 4;
            if ((len ^ 0xFFFF) != nlen || br.pos + len > n) return false;
            out.insert(out.end(), src + br.pos, src + br.pos + len);
            br.pos += len;
            continue;
        }
        if (type == 3) return false;
        Huffman lit, dist;
        uint8_t lens[320];
        if (type == 1) {
            int i = 0;
            for (; i < 144; i++) lens[i] = 8;
            for (; i < 256; i++) lens[i] = 9;
            for (; i < 280; i++) lens[i] = 7;
            for (; i < 288; i++) lens[i] = 8;
            lit.build(lens, 288);
            for (i = 0; i < 30; i++) lens[i] = 5;
            dist.build(lens, 30);
        } else {
            int nl = (int)br.bits(5) + 257, nd = (int)br.bits(5) + 1, nc = (int)br.bits(4) + 4;
            if (br.bad || nl > 286 || nd > 30) return false;
            uint8_t cl[19] = {0};
            for (int i = 0; i < nc; i++) cl[order[i]] = (uint8_t)br.bits(3);
            Huffman ch;
            ch.build(cl, 19);
            for (int i = 0; i < nl + nd;) {
                int s = ch.decode(br);
                if (s < 0) return false;
                if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                int rep, val = 0;
                if (s == 16) { if (i == 0) return false; val = lens[i - 1]; rep = 3 + (int)br.bits(2); }
                else if (s == 17) rep = 3 + (int)br.bits(3);
                else rep = 11 + (int)br.bits(7);
                if (i + rep > nl + nd) return false;
                while (rep--) lens[i++] = (uint8_t)val;
            }
            lit.build(lens, nl);
            dist.build(lens + nl, nd);
        }
        for (;;) {
            int s = lit.decode(br);
            if (s < 0) return false;
            if (s < 256) { out.push_back((uint8_t)s); continue; }
            if (s == 256) break;
            s -= 257;
            if (s >= 29) return false;
            size_t len = lbase[s] + br.bits(lext[s]);
            int d = dist.decode(br);
            if (d < 0 || d >= 30) return false;
            size_t back = dbase[d] + br.bits(dext[d]);
            if (br.bad || back > out.size()) return false;
            size_t from = out.size() - back;
            for (size_t i = 0; i < len; i++) out.push_back(out[from + i]);
        }
    }
    return true;
}

// Decodes into interleaved BGR (OpenCV order), 3 bytes per pixel.
inline bool read_bgr(const std::string& path, std::vector<uint8_t>& bgr, int& w, int& h)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::vector<uint8_t> d;
    uint8_t buf[65536];
    for (size_t k; (k = fread(buf, 1, sizeof(buf), f)) > 0;) d.insert(d.end(), buf, buf + k);
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};
    if (d.size() < 8 || memcmp(d.data(), sig, 8)) return false;
    auto rd32 = [&](size_t o) { return (uint32_t)d[o] << 24 | (uint32_t)d[o + 1] << 16 | (uint32_t)d[o + 2] << 8 | d[o + 3]; };
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    w = h = 0;
    for (size_t o = 8; o + 12 <= d.size();) {
        uint32_t len = rd32(o);
        if (o + 12 + len > d.size()) return false;
        const char* tag = (const char*)&d[o + 4];
        const uint8_t* body = &d[o + 8];
        if (!memcmp(tag, "IHDR", 4) && len >= 13) {
            w = (int)rd32(o + 8); h = (int)rd32(o + 12); depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!memcmp(tag, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(tag, "IEND", 4)) break;
        o += 12 + len;
    }
    if (w <= 0 || h <= 0 || interlace || idat.size() < 6) return false;
    int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch || (ctype == 3 ? (depth != 1 && depth != 2 && depth != 4 && depth != 8) : depth != 8)) return false;
    std::vector<uint8_t> raw;
    if (!inflate(idat.data() + 2, idat.size() - 2, raw)) return false;    // skip the 2-byte zlib header; adler not checked
    const size_t rowb = ((size_t)w * ch * depth + 7) / 8, bpp = (ch * depth + 7) / 8;
    if (raw.size() < (rowb + 1) * h) return false;
    std::vector<uint8_t> prev(rowb, 0), cur(rowb);
    bgr.assign((size_t)w * h * 3, 0);
    for (int y = 0; y < h; y++) {
        const uint8_t* r = &raw[(rowb + 1) * y];
        int ft = r[0];
        for (size_t i = 0; i < rowb; i++) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0, x = r[1 + i];
            switch (ft) {
            case 0: break;
            case 1: x += a; break;
            case 2: x += b; break;
            case 3: x += (a + b) >> 1; break;
            case 4: {
                int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                x += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                break;
            }
            default: return false;
            }
            cur[i] = (uint8_t)x;
        }
        uint8_t* o = &bgr[(size_t)y * w * 3];
        for (int x = 0; x < w; x++) {
            if (ctype == 3) {
                int idx = depth == 8 ? cur[x] : (cur[(x * depth) >> 3] >> (8 - depth - ((x * depth) & 7))) & ((1 << depth) - 1);
                if ((size_t)idx * 3 + 2 >= plte.size()) return false;
                o[3 * x] = plte[idx * 3 + 2]; o[3 * x + 1] = plte[idx * 3 + 1]; o[3 * x + 2] = plte[idx * 3];
            } else if (ch <= 2) {
                o[3 * x] = o[3 * x + 1] = o[3 * x + 2] = cur[(size_t)x * ch];
            } else {
                o[3 * x] = cur[(size_t)x * ch + 2]; o[3 * x + 1] = cur[(size_t)x * ch + 1]; o[3 * x + 2] = cur[(size_t)x * ch];
            }
        }
        prev.swap(cur);
    }
    return true;
}

}  // namespace png
}  // namespace msegment
