// assets.cpp — scene / asset ingestion in the reference's formats.  See assets.h for the reference call sites.
// Self-contained: the reference links assimp and stb_image (not vendored in /root/reference, not installed here), so the formats are
// restated from their published specifications: RFC 1950 / 1951 (zlib / deflate), the PNG specification (filters, Adam7, tRNS),
// Radiance RGBE, Wavefront OBJ / MTL, glTF 2.0 (JSON + .bin / .glb / base64 data URIs).
#include "assets.h"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int code, const char* fmt, ...)
{
    char    buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

bool read_file(const std::string& path, std::vector<uint8_t>& out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
    fclose(f);
    return got == out.size();
}
std::string dir_of(const std::string& p)
{
    size_t k = p.find_last_of("/\\");
    return k == std::string::npos ? std::string() : p.substr(0, k + 1);
}
std::string ext_of(const std::string& p)
{
    size_t k = p.find_last_of('.');
    std::string e = k == std::string::npos ? std::string() : p.substr(k + 1);
    for (auto& c : e) c = (char)tolower((unsigned char)c);
    return e;
}
std::string resolve(const std::string& base_file, std::string rel)
{ // resolve_relative_path + the '\\' -> '/' replacement (mesh.cpp:56-84, 352)
    std::replace(rel.begin(), rel.end(), '\\', '/');
    if (!rel.empty() && rel[0] == '/') return rel;
    return dir_of(base_file) + rel;
}

// =====================================================================================================================
// inflate (RFC 1951) inside a zlib stream (RFC 1950)
// =====================================================================================================================
struct Bits {
    const uint8_t* p; size_t n, pos = 0; uint32_t buf = 0; int cnt = 0; bool bad = false;
    int bits(int need)
    {
        uint32_t v = buf;
        while (cnt < need)
        {
            if (pos >= n) { bad = true; return 0; }
            v |= (uint32_t)p[pos++] << cnt;
            cnt += 8;
        }
        buf = need < 32 ? v >> need : 0;
        cnt -= need;
        return (int)(v & ((need < 32 ? (1u << need) : 0u) - 1u));
    }
};
struct Huff { short count[16]; short symbol[288]; };
int huff_build(Huff& h, const short* length, int n)
{
    short offs[16];
    for (int l = 0; l < 16; l++) h.count[l] = 0;
    for (int s = 0; s < n; s++) h.count[length[s]]++;
    if (h.count[0] == n) return 0;
    int left = 1;
    for (int l = 1; l < 16; l++) { left <<= 1; left -= h.count[l]; if (left < 0) return left; }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h.count[l];
    for (int s = 0; s < n; s++) if (length[s] != 0) h.symbol[offs[length[s]]++] = (short)s;
    return left;
}
int huff_decode(Bits& b, const Huff& h)
{
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; l++)
    {
        code |= b.bits(1);
        if (b.bad) return -1;
        int count = h.count[l];
        if (code - count < first) return h.symbol[index + (code - first)];
        index += count; first += count; first <<= 1; code <<= 1;
    }
    return -1;
}
bool inflate_codes(Bits& b, std::vector<uint8_t>& out, const Huff& lc, const Huff& dc)
{
    static const short lens[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
    static const short lext[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
    static const short dists[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
    static const short dext[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
    for (;;)
    {
        int sym = huff_decode(b, lc);
        if (sym < 0) return false;
        if (sym < 256) out.push_back((uint8_t)sym);
        else if (sym == 256) return true;
        else
        {
            sym -= 257;
            if (sym >= 29) return false;
            int len = lens[sym] + b.bits(lext[sym]);
            int ds  = huff_decode(b, dc);
            if (ds < 0 || ds >= 30) return false;
            size_t dist = (size_t)dists[ds] + (size_t)b.bits(dext[ds]);
            if (b.bad || dist > out.size()) return false;
            size_t from = out.size() - dist;
            for (int i = 0; i < len; i++) out.push_back(out[from + i]);
        }
    }
}
bool zlib_inflate(const uint8_t* src, size_t n, std::vector<uint8_t>& out)
{
    if (n < 2 || (src[0] & 15) != 8 || ((src[0] << 8) + src[1]) % 31 != 0 || (src[1] & 32)) return false;
    Bits b { src + 2, n - 2 };
    Huff fixl, fixd;
    {
        short l[288];
        int   s = 0;
        for (; s < 144; s++) l[s] = 8;
        for (; s < 256; s++) l[s] = 9;
        for (; s < 280; s++) l[s] = 7;
        for (; s < 288; s++) l[s] = 8;
        huff_build(fixl, l, 288);
        for (s = 0; s < 30; s++) l[s] = 5;
        huff_build(fixd, l, 30);
    }
    int last;
    do
    {
        last     = b.bits(1);
        int type = b.bits(2);
        if (b.bad) return false;
        if (type == 0)
        {
            b.buf = 0; b.cnt = 0;
            if (b.pos + 4 > b.n) return false;
            unsigned len = b.p[b.pos] | (b.p[b.pos + 1] << 8), nlen = b.p[b.pos + 2] | (b.p[b.pos + 3] << 8);
            b.pos += 4;
            if ((len ^ 0xFFFFu) != nlen || b.pos + len > b.n) return false;
            out.insert(out.end(), b.p + b.pos, b.p + b.pos + len);
            b.pos += len;
        }
        else if (type == 1) { if (!inflate_codes(b, out, fixl, fixd)) return false; }
        else if (type == 2)
        {
            static const short order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
            short lengths[320];
            int   nlen = b.bits(5) + 257, ndist = b.bits(5) + 1, ncode = b.bits(4) + 4;
            if (b.bad || nlen > 286 || ndist > 30) return false;
            int idx = 0;
            for (; idx < ncode; idx++) lengths[order[idx]] = (short)b.bits(3);
            for (; idx < 19; idx++) lengths[order[idx]] = 0;
            Huff cl;
            if (huff_build(cl, lengths, 19) != 0) return false;
            idx = 0;
            while (idx < nlen + ndist)
            {
                int sym = huff_decode(b, cl);
                if (sym < 0) return false;
                if (sym < 16) lengths[idx++] = (short)sym;
                else
                {
                    int len = 0, rep;
                    if (sym == 16) { if (idx == 0) return false; len = lengths[idx - 1]; rep = 3 + b.bits(2); }
                    else if (sym == 17) rep = 3 + b.bits(3);
                    else rep = 11 + b.bits(7);
                    if (idx + rep > nlen + ndist) return false;
                    while (rep--) lengths[idx++] = (short)len;
                }
            }
            if (lengths[256] == 0) return false;
            Huff dl, dd;
            int  e = huff_build(dl, lengths, nlen);
            if (e < 0 || (e > 0 && nlen - dl.count[0] != 1)) return false;
            e = huff_build(dd, lengths + nlen, ndist);
            if (e < 0 || (e > 0 && ndist - dd.count[0] != 1)) return false;
            if (!inflate_codes(b, out, dl, dd)) return false;
        }
        else return false;
    } while (!last);
    return !b.bad;
}

// =====================================================================================================================
// PNG
// =====================================================================================================================
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
// unfilter `h` scanlines of `stride` bytes (+1 filter byte each); bpp = bytes per complete pixel (>= 1)
bool png_unfilter(const uint8_t* src, size_t avail, int h, size_t stride, int bpp, std::vector<uint8_t>& dst)
{
    if (avail < (stride + 1) * (size_t)h) return false;
    dst.assign(stride * (size_t)h, 0);
    for (int y = 0; y < h; y++)
    {
        const uint8_t  ft   = src[(stride + 1) * (size_t)y];
        const uint8_t* in   = src + (stride + 1) * (size_t)y + 1;
        uint8_t*       cur  = dst.data() + stride * (size_t)y;
        const uint8_t* prev = y ? cur - stride : nullptr;
        for (size_t i = 0; i < stride; i++)
        {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= (size_t)bpp) ? prev[i - bpp] : 0;
            int v = in[i];
            switch (ft)
            {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: return false;
            }
            cur[i] = (uint8_t)v;
        }
    }
    return true;
}

int png_decode(const uint8_t* d, size_t n, bool flip, int* W, int* H, int* C, uint8_t** out)
{
    static const uint8_t sig[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
    if (n < 8 || memcmp(d, sig, 8) != 0) return fail(HRA_ERR_FORMAT, "not a PNG file");
    size_t pos = 8;
    int    w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool have_ihdr = false, done = false;
    while (!done && pos + 12 <= n)
    {
        const uint32_t len = be32(d + pos);
        const uint8_t* ty  = d + pos + 4;
        if (pos + 12 + (size_t)len > n) return fail(HRA_ERR_FORMAT, "PNG: truncated chunk");
        const uint8_t* body = d + pos + 8;
        if (!memcmp(ty, "IHDR", 4))
        {
            if (len != 13) return fail(HRA_ERR_FORMAT, "PNG: bad IHDR");
            w = (int)be32(body); h = (int)be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            if (body[10] != 0 || body[11] != 0 || interlace > 1) return fail(HRA_ERR_FORMAT, "PNG: unknown compression / filter / interlace method");
            have_ihdr = true;
        }
        else if (!memcmp(ty, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(ty, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(ty, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(ty, "IEND", 4)) done = true;
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || w <= 0 || h <= 0 || (size_t)w * h > (1u << 28)) return fail(HRA_ERR_FORMAT, "PNG: missing IHDR or bad size");
    int src_ch;
    switch (ctype) { case 0: src_ch = 1; break; case 2: src_ch = 3; break; case 3: src_ch = 1; break; case 4: src_ch = 2; break; case 6: src_ch = 4; break;
                     default: return fail(HRA_ERR_FORMAT, "PNG: bad colour type %d", ctype); }
    const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8) : (depth == 8 || depth == 16);
    if (!depth_ok) return fail(HRA_ERR_FORMAT, "PNG: bit depth %d not allowed for colour type %d", depth, ctype);
    if (ctype == 3 && plte.size() < 3) return fail(HRA_ERR_FORMAT, "PNG: palette image without PLTE");
    std::vector<uint8_t> raw;
    if (!zlib_inflate(idat.data(), idat.size(), raw)) return fail(HRA_ERR_FORMAT, "PNG: corrupt deflate stream");
    const int bits_pp = depth * src_ch, bpp = std::max(1, bits_pp / 8);
    { // a header may claim any size: the decompressed data must cover it before anything of that size is allocated
        size_t need = 0;
        if (!interlace) need = (((size_t)w * bits_pp + 7) / 8 + 1) * (size_t)h;
        else
        {
            static const int xs[7] = { 0, 4, 0, 2, 0, 1, 0 }, ys[7] = { 0, 0, 4, 0, 2, 0, 1 }, dxs[7] = { 8, 8, 4, 4, 2, 2, 1 }, dys[7] = { 8, 8, 8, 4, 4, 2, 2 };
            for (int p = 0; p < 7; p++)
            {
                const int pw = (w - xs[p] + dxs[p] - 1) / dxs[p], ph = (h - ys[p] + dys[p] - 1) / dys[p];
                if (pw > 0 && ph > 0) need += (((size_t)pw * bits_pp + 7) / 8 + 1) * (size_t)ph;
            }
        }
        if (raw.size() < need) return fail(HRA_ERR_FORMAT, "PNG: image data too short for %d x %d", w, h);
    }
    // samples -> 16-bit values per channel, full image
    std::vector<uint16_t> samp((size_t)w * h * src_ch);
    auto unpack_pass = [&](const std::vector<uint8_t>& lines, int pw, int ph, int x0, int y0, int dx, int dy) {
        const size_t stride = ((size_t)pw * bits_pp + 7) / 8;
        for (int y = 0; y < ph; y++)
            for (int x = 0; x < pw; x++)
                for (int c = 0; c < src_ch; c++)
                {
                    const uint8_t* row = lines.data() + stride * (size_t)y;
                    const size_t   si  = (size_t)x * src_ch + c;
                    uint16_t       v;
                    if (depth == 8) v = row[si];
                    else if (depth == 16) v = (uint16_t)((row[2 * si] << 8) | row[2 * si + 1]);
                    else { const size_t bit = si * depth; v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1); }
                    samp[((size_t)(y0 + y * dy) * w + (x0 + x * dx)) * src_ch + c] = v;
                }
    };
    if (!interlace)
    {
        std::vector<uint8_t> lines;
        if (!png_unfilter(raw.data(), raw.size(), h, ((size_t)w * bits_pp + 7) / 8, bpp, lines)) return fail(HRA_ERR_FORMAT, "PNG: image data too short / bad filter");
        unpack_pass(lines, w, h, 0, 0, 1, 1);
    }
    else
    {
        static const int xs[7] = { 0, 4, 0, 2, 0, 1, 0 }, ys[7] = { 0, 0, 4, 0, 2, 0, 1 }, dxs[7] = { 8, 8, 4, 4, 2, 2, 1 }, dys[7] = { 8, 8, 8, 4, 4, 2, 2 };
        size_t off = 0;
        for (int p = 0; p < 7; p++)
        {
            const int pw = (w - xs[p] + dxs[p] - 1) / dxs[p], ph = (h - ys[p] + dys[p] - 1) / dys[p];
            if (pw <= 0 || ph <= 0) continue;
            const size_t stride = ((size_t)pw * bits_pp + 7) / 8;
            std::vector<uint8_t> lines;
            if (off > raw.size() || !png_unfilter(raw.data() + off, raw.size() - off, ph, stride, bpp, lines)) return fail(HRA_ERR_FORMAT, "PNG: interlaced data too short / bad filter");
            off += (stride + 1) * (size_t)ph;
            unpack_pass(lines, pw, ph, xs[p], ys[p], dxs[p], dys[p]);
        }
    }
    // stb_image channel rules: grey 1 (+tRNS: 2), grey+alpha 2, RGB 3 (+tRNS: 4), palette 3 (+tRNS: 4), RGBA 4; then Image::create_from_file
    // reloads 3-channel images as 4 (vk.cpp:163-168)
    int out_ch = ctype == 0 ? (trns.size() >= 2 ? 2 : 1) : ctype == 4 ? 2 : 4;
    uint8_t* o = (uint8_t*)malloc((size_t)w * h * out_ch);
    if (!o) return fail(HRA_ERR_IO, "out of memory");
    const int maxv = (1 << depth) - 1;
    auto to8 = [&](uint16_t v) -> uint8_t { return depth == 16 ? (uint8_t)(v >> 8) : depth == 8 ? (uint8_t)v : (uint8_t)(v * (255 / maxv)); };
    for (size_t i = 0; i < (size_t)w * h; i++)
    {
        const uint16_t* s = samp.data() + i * src_ch;
        const size_t    y = i / w, x = i % w;
        uint8_t*        q = o + ((flip ? (size_t)(h - 1 - y) : y) * w + x) * out_ch;
        if (ctype == 0)
        {
            q[0] = to8(s[0]);
            if (out_ch == 2) q[1] = s[0] == (uint16_t)((trns[0] << 8) | trns[1]) ? 0 : 255;
        }
        else if (ctype == 4) { q[0] = to8(s[0]); q[1] = to8(s[1]); }
        else if (ctype == 2)
        {
            q[0] = to8(s[0]); q[1] = to8(s[1]); q[2] = to8(s[2]); q[3] = 255;
            if (trns.size() >= 6 && s[0] == (uint16_t)((trns[0] << 8) | trns[1]) && s[1] == (uint16_t)((trns[2] << 8) | trns[3]) && s[2] == (uint16_t)((trns[4] << 8) | trns[5])) q[3] = 0;
        }
        else if (ctype == 6) { q[0] = to8(s[0]); q[1] = to8(s[1]); q[2] = to8(s[2]); q[3] = to8(s[3]); }
        else
        {
            const size_t pi = s[0];
            if (3 * pi + 2 >= plte.size()) { free(o); return fail(HRA_ERR_FORMAT, "PNG: palette index out of range"); }
            q[0] = plte[3 * pi]; q[1] = plte[3 * pi + 1]; q[2] = plte[3 * pi + 2];
            q[3] = pi < trns.size() ? trns[pi] : 255;
        }
    }
    *W = w; *H = h; *C = out_ch; *out = o;
    return HRA_OK;
}

// ---- PNG writer (stored deflate blocks) -----------------------------------------------------------------------------------------
uint32_t crc32_update(uint32_t c, const uint8_t* p, size_t n)
{
    static uint32_t table[256];
    static bool     init = false;
    if (!init)
    {
        for (uint32_t i = 0; i < 256; i++)
        {
            uint32_t v = i;
            for (int k = 0; k < 8; k++) v = (v & 1u) ? 0xEDB88320u ^ (v >> 1) : v >> 1;
            table[i] = v;
        }
        init = true;
    }
    c = ~c;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return ~c;
}
void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x); }
void png_chunk(std::vector<uint8_t>& out, const char* tag, const std::vector<uint8_t>& body)
{
    put_be32(out, (uint32_t)body.size());
    const size_t start = out.size();
    out.insert(out.end(), tag, tag + 4);
    out.insert(out.end(), body.begin(), body.end());
    put_be32(out, crc32_update(0, out.data() + start, out.size() - start));
}

// =====================================================================================================================
// Radiance .hdr (RGBE)
// =====================================================================================================================
int hdr_decode(const uint8_t* d, size_t n, bool flip, int* W, int* H, float** out)
{
    size_t pos = 0;
    auto   line = [&](std::string& s) -> bool {
        s.clear();
        if (pos >= n) return false;
        while (pos < n && d[pos] != '\n') s.push_back((char)d[pos++]);
        pos++;
        return true;
    };
    std::string s;
    if (!line(s) || (s != "#?RADIANCE" && s != "#?RGBE")) return fail(HRA_ERR_FORMAT, "HDR: missing #?RADIANCE signature");
    bool fmt = false;
    for (;;)
    {
        if (!line(s)) return fail(HRA_ERR_FORMAT, "HDR: truncated header");
        if (s.empty()) break;
        if (s == "FORMAT=32-bit_rle_rgbe") fmt = true;
    }
    if (!fmt) return fail(HRA_ERR_UNSUPPORTED, "HDR: unsupported format (only 32-bit_rle_rgbe)");
    if (!line(s)) return fail(HRA_ERR_FORMAT, "HDR: missing resolution line");
    int w = 0, h = 0;
    if (sscanf(s.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) return fail(HRA_ERR_UNSUPPORTED, "HDR: unsupported data layout (only -Y h +X w)");
    float* o = (float*)malloc((size_t)w * h * 4 * sizeof(float));
    if (!o) return fail(HRA_ERR_IO, "out of memory");
    auto conv = [](const uint8_t* p, float* q) { // stbi__hdr_convert
        if (p[3] != 0) { const float f = ldexpf(1.0f, (int)p[3] - (128 + 8)); q[0] = p[0] * f; q[1] = p[1] * f; q[2] = p[2] * f; }
        else q[0] = q[1] = q[2] = 0.0f;
        q[3] = 1.0f;
    };
    std::vector<uint8_t> scan((size_t)w * 4);
    for (int y = 0; y < h; y++)
    {
        float* row = o + (size_t)(flip ? h - 1 - y : y) * w * 4;
        bool   rle = false;
        if (w >= 8 && w < 32768 && pos + 4 <= n && d[pos] == 2 && d[pos + 1] == 2 && !(d[pos + 2] & 0x80))
        {
            if (((d[pos + 2] << 8) | d[pos + 3]) != w) { free(o); return fail(HRA_ERR_FORMAT, "HDR: scanline width mismatch"); }
            rle = true;
            pos += 4;
        }
        if (rle)
        {
            for (int c = 0; c < 4; c++)
            {
                int x = 0;
                while (x < w)
                {
                    if (pos >= n) { free(o); return fail(HRA_ERR_FORMAT, "HDR: truncated RLE data"); }
                    int cnt = d[pos++];
                    if (cnt > 128)
                    {
                        cnt -= 128;
                        if (pos >= n || x + cnt > w) { free(o); return fail(HRA_ERR_FORMAT, "HDR: corrupt RLE run"); }
                        const uint8_t v = d[pos++];
                        for (int i = 0; i < cnt; i++) scan[(size_t)(x++) * 4 + c] = v;
                    }
                    else
                    {
                        if (cnt == 0 || pos + cnt > n || x + cnt > w) { free(o); return fail(HRA_ERR_FORMAT, "HDR: corrupt RLE literal"); }
                        for (int i = 0; i < cnt; i++) scan[(size_t)(x++) * 4 + c] = d[pos++];
                    }
                }
            }
            for (int x = 0; x < w; x++) conv(scan.data() + (size_t)x * 4, row + (size_t)x * 4);
        }
        else
        {
            if (pos + (size_t)w * 4 > n) { free(o); return fail(HRA_ERR_FORMAT, "HDR: truncated pixel data"); }
            for (int x = 0; x < w; x++) conv(d + pos + (size_t)x * 4, row + (size_t)x * 4);
            pos += (size_t)w * 4;
        }
    }
    *W = w; *H = h; *out = o;
    return HRA_OK;
}

// =====================================================================================================================
// JSON (RFC 8259) — just enough of a DOM for glTF
// =====================================================================================================================
struct JVal {
    enum T { NUL, BOOL, NUM, STR, ARR, OBJ } t = NUL;
    double num = 0; bool b = false; std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const char* k) const
    {
        if (t != OBJ) return nullptr;
        for (auto& kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    double      number(const char* k, double def) const { const JVal* v = get(k); return v && v->t == NUM ? v->num : def; }
    long        integer(const char* k, long def) const { const JVal* v = get(k); return v && v->t == NUM ? (long)v->num : def; }
    std::string string(const char* k) const { const JVal* v = get(k); return v && v->t == STR ? v->str : std::string(); }
    const JVal* array(const char* k) const { const JVal* v = get(k); return v && v->t == ARR ? v : nullptr; }
};
struct JParser {
    const char* p; const char* e; bool ok = true; int depth = 0;
    void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool lit(const char* s) { size_t n = strlen(s); if ((size_t)(e - p) >= n && !memcmp(p, s, n)) { p += n; return true; } return false; }
    void utf8(std::string& s, unsigned cp)
    {
        if (cp < 0x80) s.push_back((char)cp);
        else if (cp < 0x800) { s.push_back((char)(0xC0 | (cp >> 6))); s.push_back((char)(0x80 | (cp & 63))); }
        else if (cp < 0x10000) { s.push_back((char)(0xE0 | (cp >> 12))); s.push_back((char)(0x80 | ((cp >> 6) & 63))); s.push_back((char)(0x80 | (cp & 63))); }
        else { s.push_back((char)(0xF0 | (cp >> 18))); s.push_back((char)(0x80 | ((cp >> 12) & 63))); s.push_back((char)(0x80 | ((cp >> 6) & 63))); s.push_back((char)(0x80 | (cp & 63))); }
    }
    bool str(std::string& s)
    {
        if (p >= e || *p != '"') return ok = false;
        p++;
        while (p < e && *p != '"')
        {
            if (*p == '\\')
            {
                if (++p >= e) return ok = false;
                switch (*p)
                {
                case 'n': s.push_back('\n'); break; case 't': s.push_back('\t'); break; case 'r': s.push_back('\r'); break;
                case 'b': s.push_back('\b'); break; case 'f': s.push_back('\f'); break;
                case 'u': {
                    if (e - p < 5) return ok = false;
                    unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                    p += 4;
                    if (cp >= 0xD800 && cp < 0xDC00 && e - p >= 7 && p[1] == '\\' && p[2] == 'u')
                    {
                        unsigned lo = (unsigned)strtoul(std::string(p + 3, p + 7).c_str(), nullptr, 16);
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        p += 6;
                    }
                    utf8(s, cp);
                    break;
                }
                default: s.push_back(*p);
                }
                p++;
            }
            else s.push_back(*p++);
        }
        if (p >= e) return ok = false;
        p++;
        return true;
    }
    bool val(JVal& v)
    {
        if (++depth > 200) return ok = false;
        ws();
        if (p >= e) return ok = false;
        if (*p == '{')
        {
            v.t = JVal::OBJ; p++; ws();
            if (p < e && *p == '}') { p++; depth--; return true; }
            for (;;)
            {
                ws();
                std::string k;
                if (!str(k)) return false;
                ws();
                if (p >= e || *p != ':') return ok = false;
                p++;
                v.obj.emplace_back(k, JVal());
                if (!val(v.obj.back().second)) return false;
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == '}') { p++; break; }
                return ok = false;
            }
        }
        else if (*p == '[')
        {
            v.t = JVal::ARR; p++; ws();
            if (p < e && *p == ']') { p++; depth--; return true; }
            for (;;)
            {
                v.arr.emplace_back();
                if (!val(v.arr.back())) return false;
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == ']') { p++; break; }
                return ok = false;
            }
        }
        else if (*p == '"') { v.t = JVal::STR; if (!str(v.str)) return false; }
        else if (lit("true")) { v.t = JVal::BOOL; v.b = true; }
        else if (lit("false")) { v.t = JVal::BOOL; v.b = false; }
        else if (lit("null")) v.t = JVal::NUL;
        else
        {
            char* end = nullptr;
            std::string tmp(p, std::min<size_t>((size_t)(e - p), 64));
            v.num = strtod(tmp.c_str(), &end);
            if (end == tmp.c_str()) return ok = false;
            v.t = JVal::NUM;
            p += end - tmp.c_str();
        }
        depth--;
        return true;
    }
};

bool base64_decode(const std::string& s, size_t start, std::vector<uint8_t>& out)
{
    uint32_t acc = 0; int nb = 0;
    for (size_t i = start; i < s.size(); i++)
    {
        const char c = s[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62;
        else if (c == '/' || c == '_') v = 63;
        else if (c == '=' || c == '\n' || c == '\r') continue;
        else return false;
        acc = (acc << 6) | (uint32_t)v;
        nb += 6;
        if (nb >= 8) { nb -= 8; out.push_back((uint8_t)((acc >> nb) & 0xFF)); }
    }
    return true;
}

// =====================================================================================================================
// meshes
// =====================================================================================================================
struct V3 { float x, y, z; };
inline V3    operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3    operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3    operator*(V3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3    cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline V3    normalize_safe(V3 a) { const float l = std::sqrt(dot(a, a)); return l > 0.0f ? a * (1.0f / l) : a; }

struct MaterialRec {
    hr_material m;
    std::string tex[5];
};
// one assimp mesh (aiMesh) before it is appended to the dw::Mesh arrays
struct RawMesh {
    std::vector<V3> pos, nrm, tan, bit;
    std::vector<float> uv; // 2 per vertex
    bool has_nrm = false, has_uv = false, has_tan = false;
    std::vector<uint32_t> idx;
    int material = -1; // index into the importer's material list (-1: the importer's default material)
};

} // namespace

struct hra_mesh {
    std::string path;
    std::vector<hr_vertex>   vertices;
    std::vector<uint32_t>    indices;
    std::vector<hra_submesh> submeshes;
    std::vector<MaterialRec> materials;
    std::vector<hr_material> material_consts;
    float mn[3] = { 0, 0, 0 }, mx[3] = { 0, 0, 0 };
};

struct hra_scene {
    struct Inst { const hra_mesh* mesh; float model[16]; };
    std::vector<Inst>        insts;
    std::vector<hr_vertex>   vertices;
    std::vector<uint32_t>    indices;
    std::vector<hr_instance> instances;
    std::vector<hr_material> materials;
    // material textures (hra_scene_textures): decoded images, their hr_texture views, one binding per global material
    std::vector<std::vector<uint8_t>>  tex_pixels;
    std::vector<hr_texture>            textures;
    std::vector<hr_material_textures>  bindings;
    std::string                        tex_warnings;
    bool finalized = false;
};

namespace {

MaterialRec default_material(float grey, float metallic)
{ // reference defaults when a key is missing: albedo (1,1,1,1) overwritten by the importer's diffuse colour, roughness 1, metallic 0 (mesh.cpp:309-312)
    MaterialRec r;
    memset(&r.m, 0, sizeof(r.m));
    r.m.albedo[0] = r.m.albedo[1] = r.m.albedo[2] = grey;
    r.m.albedo[3] = 1.0f;
    r.m.roughness = 1.0f;
    r.m.metallic  = metallic;
    return r;
}

// GenSmoothNormals: face normals (normalised) summed over all vertices of the sub-mesh that share a position, normalised
void gen_smooth_normals(RawMesh& m)
{
    struct Key { uint32_t a, b, c; bool operator<(const Key& o) const { return a != o.a ? a < o.a : (b != o.b ? b < o.b : c < o.c); } };
    auto key = [](V3 p) { Key k; memcpy(&k.a, &p.x, 4); memcpy(&k.b, &p.y, 4); memcpy(&k.c, &p.z, 4); if (p.x == 0.0f) k.a = 0; if (p.y == 0.0f) k.b = 0; if (p.z == 0.0f) k.c = 0; return k; };
    std::map<Key, V3> acc;
    for (size_t f = 0; f + 2 < m.idx.size(); f += 3)
    {
        const V3 p0 = m.pos[m.idx[f]], p1 = m.pos[m.idx[f + 1]], p2 = m.pos[m.idx[f + 2]];
        const V3 fn = normalize_safe(cross(p1 - p0, p2 - p0));
        for (int j = 0; j < 3; j++) { V3& a = acc[key(m.pos[m.idx[f + j]])]; a = a + fn; }
    }
    m.nrm.resize(m.pos.size());
    for (size_t i = 0; i < m.pos.size(); i++)
    {
        auto it = acc.find(key(m.pos[i]));
        m.nrm[i] = it == acc.end() ? V3 { 0, 0, 0 } : normalize_safe(it->second);
    }
    m.has_nrm = true;
}

// CalcTangentSpace: per-face tangent / bitangent from the UV deltas, accumulated per vertex, made orthogonal to the normal
void calc_tangents(RawMesh& m)
{
    m.tan.assign(m.pos.size(), V3 { 0, 0, 0 });
    m.bit.assign(m.pos.size(), V3 { 0, 0, 0 });
    for (size_t f = 0; f + 2 < m.idx.size(); f += 3)
    {
        const uint32_t i0 = m.idx[f], i1 = m.idx[f + 1], i2 = m.idx[f + 2];
        const V3 v = m.pos[i1] - m.pos[i0], w = m.pos[i2] - m.pos[i0];
        float sx = m.uv[2 * i1] - m.uv[2 * i0], sy = m.uv[2 * i1 + 1] - m.uv[2 * i0 + 1];
        float tx = m.uv[2 * i2] - m.uv[2 * i0], ty = m.uv[2 * i2 + 1] - m.uv[2 * i0 + 1];
        const float dir = (tx * sy - ty * sx) < 0.0f ? -1.0f : 1.0f;
        if (sx * ty == sy * tx) { sx = 0.0f; sy = 1.0f; tx = 1.0f; ty = 0.0f; }
        const V3 t = (w * sy - v * ty) * dir, b = (w * sx - v * tx) * dir;
        for (uint32_t i : { i0, i1, i2 }) { m.tan[i] = m.tan[i] + t; m.bit[i] = m.bit[i] + b; }
    }
    for (size_t i = 0; i < m.pos.size(); i++)
    {
        const V3 n = m.nrm[i];
        m.tan[i] = normalize_safe(m.tan[i] - n * dot(m.tan[i], n));
        m.bit[i] = normalize_safe(m.bit[i] - n * dot(m.bit[i], n));
    }
    m.has_tan = true;
}

// Mesh::load_from_disk after the importer ran (mesh.cpp:257-613): sub-mesh table, material table in first-use order, vertex /
// index arrays with the base vertex folded into the indices, extents
int assemble(hra_mesh* out, std::vector<RawMesh>& raws, const std::vector<MaterialRec>& imported, const MaterialRec& fallback)
{
    std::unordered_map<int, uint32_t> local_mat;
    uint32_t vertex_count = 0, index_count = 0;
    for (auto& r : raws)
    {
        if (r.pos.empty() || r.idx.size() < 3) continue;
        if (!r.has_nrm) gen_smooth_normals(r);
        if (!r.has_tan && r.has_uv) calc_tangents(r);
        hra_submesh sm;
        memset(&sm, 0, sizeof(sm));
        sm.index_count  = (uint32_t)(r.idx.size() / 3 * 3);
        sm.base_index   = index_count;
        sm.base_vertex  = 0; // folded into the indices (mesh.cpp:593-601)
        sm.vertex_count = (uint32_t)r.pos.size();
        auto it = local_mat.find(r.material);
        if (it == local_mat.end())
        {
            it = local_mat.emplace(r.material, (uint32_t)out->materials.size()).first;
            out->materials.push_back(r.material >= 0 && (size_t)r.material < imported.size() ? imported[r.material] : fallback);
        }
        sm.mat_idx = it->second;
        for (int a = 0; a < 3; a++) sm.min_extents[a] = sm.max_extents[a] = (&r.pos[0].x)[a];
        for (size_t k = 0; k < r.pos.size(); k++)
        {
            hr_vertex v;
            memset(&v, 0, sizeof(v));
            v.position[0] = r.pos[k].x; v.position[1] = r.pos[k].y; v.position[2] = r.pos[k].z; v.position[3] = (float)sm.mat_idx; // mesh.cpp:544
            v.normal[0] = r.nrm[k].x; v.normal[1] = r.nrm[k].y; v.normal[2] = r.nrm[k].z;
            if (r.has_tan)
            {
                V3 t = r.tan[k];
                const V3 b = r.bit[k];
                if (dot(cross(r.nrm[k], t), b) < 0.0f) t = t * -1.0f; // mesh.cpp:553-556
                v.tangent[0] = t.x; v.tangent[1] = t.y; v.tangent[2] = t.z;
                v.bitangent[0] = b.x; v.bitangent[1] = b.y; v.bitangent[2] = b.z;
            }
            if (r.has_uv) { v.tex_coord[0] = r.uv[2 * k]; v.tex_coord[1] = r.uv[2 * k + 1]; }
            for (int a = 0; a < 3; a++)
            {
                sm.min_extents[a] = std::min(sm.min_extents[a], v.position[a]);
                sm.max_extents[a] = std::max(sm.max_extents[a], v.position[a]);
            }
            out->vertices.push_back(v);
        }
        for (size_t k = 0; k < sm.index_count; k++)
        {
            if (r.idx[k] >= r.pos.size()) return fail(HRA_ERR_FORMAT, "%s: vertex index out of range", out->path.c_str());
            out->indices.push_back(vertex_count + r.idx[k]);
        }
        vertex_count += sm.vertex_count;
        index_count += sm.index_count;
        out->submeshes.push_back(sm);
    }
    if (out->submeshes.empty()) return fail(HRA_ERR_FORMAT, "%s: no triangles", out->path.c_str());
    for (int a = 0; a < 3; a++) { out->mn[a] = out->submeshes[0].min_extents[a]; out->mx[a] = out->submeshes[0].max_extents[a]; }
    for (auto& sm : out->submeshes)
        for (int a = 0; a < 3; a++) { out->mn[a] = std::min(out->mn[a], sm.min_extents[a]); out->mx[a] = std::max(out->mx[a], sm.max_extents[a]); }
    for (auto& m : out->materials) out->material_consts.push_back(m.m);
    return HRA_OK;
}

// ---- Wavefront OBJ / MTL ---------------------------------------------------------------------------------------------
std::vector<std::string> split_ws(const std::string& s)
{
    std::vector<std::string> r;
    size_t i = 0;
    while (i < s.size())
    {
        while (i < s.size() && isspace((unsigned char)s[i])) i++;
        size_t j = i;
        while (j < s.size() && !isspace((unsigned char)s[j])) j++;
        if (j > i) r.push_back(s.substr(i, j - i));
        i = j;
    }
    return r;
}
std::string rest_after(const std::string& line, const std::string& key)
{ // file names may contain spaces: everything after the keyword, trimmed
    size_t k = line.find(key);
    std::string r = line.substr(k + key.size());
    size_t a = r.find_first_not_of(" \t"), b = r.find_last_not_of(" \t\r");
    return a == std::string::npos ? std::string() : r.substr(a, b - a + 1);
}
void for_each_line(const std::vector<uint8_t>& data, const std::function<void(const std::string&)>& fn)
{
    std::string line;
    size_t i = 0;
    while (i <= data.size())
    {
        if (i == data.size() || data[i] == '\n')
        {
            while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
            if (!line.empty() && line.back() == '\\') line.pop_back(); // continuation
            else { if (!line.empty() && line[0] != '#') fn(line); line.clear(); }
        }
        else line.push_back((char)data[i]);
        i++;
    }
}

void load_mtl(const std::string& path, std::vector<MaterialRec>& mats, std::map<std::string, int>& names)
{
    std::vector<uint8_t> data;
    if (!read_file(path, data)) return; // assimp logs and carries on with default materials
    MaterialRec* cur = nullptr;
    for_each_line(data, [&](const std::string& line) {
        const auto tok = split_ws(line);
        if (tok.empty()) return;
        const std::string& k = tok[0];
        if (k == "newmtl")
        {
            const std::string name = rest_after(line, "newmtl");
            names[name] = (int)mats.size();
            mats.push_back(default_material(0.6f, 0.0f)); // ObjFile::Material default diffuse 0.6
            cur = &mats.back();
            return;
        }
        if (!cur) return;
        auto f = [&](size_t i, float def) { return i < tok.size() ? (float)atof(tok[i].c_str()) : def; };
        // texture statements: "<key> [-option args ...] file name" — the file name is the rest of the line (it may contain spaces)
        auto tex = [&]() { const std::string r = rest_after(line, k); return resolve(path, (!r.empty() && r[0] == '-') ? tok.back() : r); };
        if (k == "Kd") { cur->m.albedo[0] = f(1, 0.6f); cur->m.albedo[1] = f(2, cur->m.albedo[0]); cur->m.albedo[2] = f(3, cur->m.albedo[0]); }
        else if (k == "d") cur->m.albedo[3] = f(1, 1.0f);
        else if (k == "Tr") cur->m.albedo[3] = 1.0f - f(1, 0.0f);
        else if (k == "Pr") cur->m.roughness = f(1, 1.0f); // PBR extension of the MTL format (absent: the reference's defaults 1 / 0)
        else if (k == "Pm") cur->m.metallic = f(1, 0.0f);
        else if (k == "map_Kd") cur->tex[HRA_TEX_ALBEDO] = tex();
        else if (k == "map_Ns") cur->tex[HRA_TEX_ROUGHNESS] = tex();  // aiTextureType_SHININESS (mesh.cpp:404)
        else if (k == "map_Ka") cur->tex[HRA_TEX_METALLIC] = tex();   // aiTextureType_AMBIENT (mesh.cpp:430)
        else if (k == "map_Ke" || k == "map_emissive") cur->tex[HRA_TEX_EMISSIVE] = tex();
        else if (k == "norm" || k == "map_Kn") cur->tex[HRA_TEX_NORMAL] = tex();
        else if ((k == "map_Bump" || k == "map_bump" || k == "bump") && cur->tex[HRA_TEX_NORMAL].empty()) cur->tex[HRA_TEX_NORMAL] = tex(); // aiTextureType_HEIGHT (mesh.cpp:477)
        // Ke: the reference tests `if (material->Get(AI_MATKEY_COLOR_EMISSIVE, e))`, which is true only on FAILURE (mesh.cpp:455): constant emission stays 0
    });
}

int load_obj(const std::string& path, hra_mesh* out)
{
    std::vector<uint8_t> data;
    if (!read_file(path, data)) return fail(HRA_ERR_IO, "cannot read %s", path.c_str());
    std::vector<V3> P, N;
    std::vector<float> T;
    std::vector<MaterialRec> mats;
    std::map<std::string, int> mat_names;
    std::vector<RawMesh> raws;
    struct Key { int v, t, n; bool operator<(const Key& o) const { return v != o.v ? v < o.v : (t != o.t ? t < o.t : n < o.n); } };
    std::map<Key, uint32_t> dedupe;
    int  cur_mat = -1;
    bool need_new = true;
    int  err = HRA_OK;
    for_each_line(data, [&](const std::string& line) {
        if (err) return;
        const auto tok = split_ws(line);
        if (tok.empty()) return;
        const std::string& k = tok[0];
        auto f = [&](size_t i) { return i < tok.size() ? (float)atof(tok[i].c_str()) : 0.0f; };
        if (k == "v") P.push_back({ f(1), f(2), f(3) });
        else if (k == "vn") N.push_back({ f(1), f(2), f(3) });
        else if (k == "vt") { T.push_back(f(1)); T.push_back(1.0f - f(2)); } // aiProcess_FlipUVs
        else if (k == "o" || k == "g") need_new = true;
        else if (k == "mtllib") load_mtl(resolve(path, rest_after(line, "mtllib")), mats, mat_names);
        else if (k == "usemtl")
        {
            const std::string name = rest_after(line, "usemtl");
            auto it = mat_names.find(name);
            const int m = it == mat_names.end() ? -1 : it->second;
            if (m != cur_mat) { cur_mat = m; need_new = true; }
        }
        else if (k == "f")
        {
            if (tok.size() < 4) return;
            if (need_new || raws.empty()) { raws.emplace_back(); raws.back().material = cur_mat; dedupe.clear(); need_new = false; }
            RawMesh& r = raws.back();
            std::vector<uint32_t> corner;
            for (size_t i = 1; i < tok.size(); i++)
            {
                Key key { 0, 0, 0 };
                const char* s = tok[i].c_str();
                char* e = nullptr;
                key.v = (int)strtol(s, &e, 10);
                if (*e == '/') { s = e + 1; if (*s != '/') key.t = (int)strtol(s, &e, 10); else e = (char*)s; if (*e == '/') key.n = (int)strtol(e + 1, &e, 10); }
                if (key.v < 0) key.v = (int)P.size() + key.v + 1;
                if (key.t < 0) key.t = (int)(T.size() / 2) + key.t + 1;
                if (key.n < 0) key.n = (int)N.size() + key.n + 1;
                if (key.v < 1 || key.v > (int)P.size() || key.t > (int)(T.size() / 2) || key.n > (int)N.size()) { err = fail(HRA_ERR_FORMAT, "%s: face index out of range", path.c_str()); return; }
                auto it = dedupe.find(key);
                if (it == dedupe.end())
                {
                    it = dedupe.emplace(key, (uint32_t)r.pos.size()).first;
                    r.pos.push_back(P[key.v - 1]);
                    r.nrm.push_back(key.n ? N[key.n - 1] : V3 { 0, 0, 0 });
                    r.uv.push_back(key.t ? T[2 * (key.t - 1)] : 0.0f);
                    r.uv.push_back(key.t ? T[2 * (key.t - 1) + 1] : 0.0f);
                    if (r.pos.size() == 1) { r.has_nrm = key.n != 0; r.has_uv = key.t != 0; }
                    else { r.has_nrm = r.has_nrm && key.n != 0; r.has_uv = r.has_uv && key.t != 0; }
                }
                corner.push_back(it->second);
            }
            for (size_t i = 1; i + 1 < corner.size(); i++) { r.idx.push_back(corner[0]); r.idx.push_back(corner[i]); r.idx.push_back(corner[i + 1]); } // aiProcess_Triangulate (fan)
        }
    });
    if (err) return err;
    return assemble(out, raws, mats, default_material(0.6f, 0.0f));
}

// ---- glTF 2.0 -----------------------------------------------------------------------------------------------------------
struct Gltf {
    JVal root;
    std::string path;
    std::vector<std::vector<uint8_t>> buffers;
    std::vector<uint8_t> glb_bin;
    bool has_glb_bin = false;
};
int gltf_buffer(Gltf& g, size_t bi, const std::vector<uint8_t>** out)
{
    const JVal* bufs = g.root.array("buffers");
    if (!bufs || bi >= bufs->arr.size()) return fail(HRA_ERR_FORMAT, "%s: buffer %zu does not exist", g.path.c_str(), bi);
    if (g.buffers.size() < bufs->arr.size()) g.buffers.resize(bufs->arr.size());
    if (g.buffers[bi].empty())
    {
        const std::string uri = bufs->arr[bi].string("uri");
        if (uri.empty())
        {
            if (bi != 0 || !g.has_glb_bin) return fail(HRA_ERR_FORMAT, "%s: buffer %zu has no uri and there is no GLB binary chunk", g.path.c_str(), bi);
            g.buffers[bi] = g.glb_bin;
        }
        else if (uri.compare(0, 5, "data:") == 0)
        {
            const size_t k = uri.find("base64,");
            if (k == std::string::npos || !base64_decode(uri, k + 7, g.buffers[bi])) return fail(HRA_ERR_FORMAT, "%s: bad data URI in buffer %zu", g.path.c_str(), bi);
        }
        else
        {
            std::string dec; // percent-decoding of the few characters exporters escape
            for (size_t i = 0; i < uri.size(); i++)
                if (uri[i] == '%' && i + 2 < uri.size()) { dec.push_back((char)strtol(uri.substr(i + 1, 2).c_str(), nullptr, 16)); i += 2; }
                else dec.push_back(uri[i]);
            if (!read_file(resolve(g.path, dec), g.buffers[bi])) return fail(HRA_ERR_IO, "cannot read %s", resolve(g.path, dec).c_str());
        }
    }
    *out = &g.buffers[bi];
    return HRA_OK;
}
// read accessor `ai` as floats (ncomp per element, normalised integers converted) or as uint32 indices
int gltf_accessor(Gltf& g, long ai, int want_comp, std::vector<float>* fout, std::vector<uint32_t>* iout, size_t* count)
{
    const JVal* accs = g.root.array("accessors");
    if (!accs || ai < 0 || (size_t)ai >= accs->arr.size()) return fail(HRA_ERR_FORMAT, "%s: accessor %ld does not exist", g.path.c_str(), ai);
    const JVal& a = accs->arr[ai];
    if (a.get("sparse")) return fail(HRA_ERR_UNSUPPORTED, "%s: sparse accessors are not supported", g.path.c_str());
    const std::string type = a.string("type");
    const int ncomp = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4 : 0;
    if (!ncomp || (want_comp && ncomp < want_comp)) return fail(HRA_ERR_FORMAT, "%s: accessor %ld has type %s", g.path.c_str(), ai, type.c_str());
    const long ct = a.integer("componentType", 0);
    const int  csz = ct == 5120 || ct == 5121 ? 1 : ct == 5122 || ct == 5123 ? 2 : ct == 5125 || ct == 5126 ? 4 : 0;
    if (!csz) return fail(HRA_ERR_FORMAT, "%s: accessor %ld has component type %ld", g.path.c_str(), ai, ct);
    const size_t n = (size_t)a.integer("count", 0);
    const bool normalized = a.get("normalized") && a.get("normalized")->b;
    const long bv = a.integer("bufferView", -1);
    *count = n;
    const int out_comp = want_comp ? want_comp : ncomp;
    if (n > (1u << 28)) return fail(HRA_ERR_FORMAT, "%s: accessor %ld claims %zu elements", g.path.c_str(), ai, n);
    if (bv < 0)
    { // no bufferView: all zeros
        if (fout) fout->assign(n * out_comp, 0.0f);
        if (iout) iout->assign(n, 0u);
        return HRA_OK;
    }
    const JVal* views = g.root.array("bufferViews");
    if (!views || (size_t)bv >= views->arr.size()) return fail(HRA_ERR_FORMAT, "%s: bufferView %ld does not exist", g.path.c_str(), bv);
    const JVal& v = views->arr[bv];
    const std::vector<uint8_t>* buf = nullptr;
    int rc = gltf_buffer(g, (size_t)v.integer("buffer", 0), &buf);
    if (rc) return rc;
    const size_t elem = (size_t)csz * ncomp;
    size_t stride = (size_t)v.integer("byteStride", 0);
    if (!stride) stride = elem;
    const size_t off = (size_t)v.integer("byteOffset", 0) + (size_t)a.integer("byteOffset", 0);
    if (n && (off > buf->size() || stride > buf->size() || off + stride * (n - 1) + elem > buf->size())) return fail(HRA_ERR_FORMAT, "%s: accessor %ld reads past the end of its buffer", g.path.c_str(), ai);
    if (fout) fout->assign(n * out_comp, 0.0f); // sized only after the data was seen to exist
    if (iout) iout->assign(n, 0u);
    for (size_t i = 0; i < n; i++)
        for (int c = 0; c < (iout ? 1 : out_comp); c++)
        {
            const uint8_t* p = buf->data() + off + stride * i + (size_t)csz * c;
            double val;
            switch (ct)
            {
            case 5120: { int8_t x; memcpy(&x, p, 1); val = normalized ? std::max(x / 127.0, -1.0) : x; break; }
            case 5121: { val = normalized ? *p / 255.0 : *p; break; }
            case 5122: { int16_t x; memcpy(&x, p, 2); val = normalized ? std::max(x / 32767.0, -1.0) : x; break; }
            case 5123: { uint16_t x; memcpy(&x, p, 2); val = normalized ? x / 65535.0 : x; break; }
            case 5125: { uint32_t x; memcpy(&x, p, 4); val = x; break; }
            default:   { float x; memcpy(&x, p, 4); val = x; break; }
            }
            if (iout) (*iout)[i] = (uint32_t)val;
            else (*fout)[i * out_comp + c] = (float)val;
        }
    return HRA_OK;
}
std::string gltf_texture_path(const Gltf& g, const JVal* texinfo)
{
    if (!texinfo) return std::string();
    const long ti = texinfo->integer("index", -1);
    const JVal* texs = g.root.array("textures");
    if (ti < 0 || !texs || (size_t)ti >= texs->arr.size()) return std::string();
    const long src = texs->arr[ti].integer("source", -1);
    const JVal* imgs = g.root.array("images");
    if (src < 0 || !imgs || (size_t)src >= imgs->arr.size()) return std::string();
    const std::string uri = imgs->arr[src].string("uri");
    if (uri.empty() || uri.compare(0, 5, "data:") == 0) return std::string();
    return resolve(g.path, uri);
}

int load_gltf(const std::string& path, hra_mesh* out)
{
    std::vector<uint8_t> data;
    if (!read_file(path, data)) return fail(HRA_ERR_IO, "cannot read %s", path.c_str());
    Gltf g;
    g.path = path;
    const char* js = (const char*)data.data();
    size_t      jn = data.size();
    if (data.size() >= 12 && !memcmp(data.data(), "glTF", 4))
    { // GLB container: header (magic, version, length) + JSON chunk + optional BIN chunk
        uint32_t ver, total;
        memcpy(&ver, data.data() + 4, 4); memcpy(&total, data.data() + 8, 4);
        if (ver != 2 || total > data.size()) return fail(HRA_ERR_FORMAT, "%s: bad GLB header", path.c_str());
        size_t pos = 12;
        js = nullptr;
        while (pos + 8 <= total)
        {
            uint32_t clen, ctype;
            memcpy(&clen, data.data() + pos, 4); memcpy(&ctype, data.data() + pos + 4, 4);
            if (pos + 8 + clen > total) return fail(HRA_ERR_FORMAT, "%s: truncated GLB chunk", path.c_str());
            if (ctype == 0x4E4F534A && !js) { js = (const char*)data.data() + pos + 8; jn = clen; }
            else if (ctype == 0x004E4942 && !g.has_glb_bin) { g.glb_bin.assign(data.begin() + pos + 8, data.begin() + pos + 8 + clen); g.has_glb_bin = true; }
            pos += 8 + ((clen + 3) & ~3u);
        }
        if (!js) return fail(HRA_ERR_FORMAT, "%s: GLB without a JSON chunk", path.c_str());
    }
    JParser jp { js, js + jn };
    if (!jp.val(g.root) || g.root.t != JVal::OBJ) return fail(HRA_ERR_FORMAT, "%s: JSON syntax error near byte %zu", path.c_str(), (size_t)(jp.p - js));
    // materials (assimp's glTF2 importer: every material carries the pbrMetallicRoughness factors with the glTF defaults 1 / 1 / 1)
    std::vector<MaterialRec> mats;
    if (const JVal* jm = g.root.array("materials"))
        for (const JVal& m : jm->arr)
        {
            MaterialRec r = default_material(1.0f, 1.0f);
            if (const JVal* pbr = m.get("pbrMetallicRoughness"))
            {
                if (const JVal* bc = pbr->array("baseColorFactor"))
                    for (size_t i = 0; i < 4 && i < bc->arr.size(); i++) r.m.albedo[i] = (float)bc->arr[i].num;
                r.m.metallic  = (float)pbr->number("metallicFactor", 1.0);
                r.m.roughness = (float)pbr->number("roughnessFactor", 1.0);
                r.tex[HRA_TEX_ALBEDO] = gltf_texture_path(g, pbr->get("baseColorTexture"));
                const std::string mr = gltf_texture_path(g, pbr->get("metallicRoughnessTexture")); // roughness = .g, metallic = .b (mesh.cpp:420,447)
                r.tex[HRA_TEX_ROUGHNESS] = mr;
                r.tex[HRA_TEX_METALLIC]  = mr;
            }
            r.tex[HRA_TEX_NORMAL]   = gltf_texture_path(g, m.get("normalTexture"));
            r.tex[HRA_TEX_EMISSIVE] = gltf_texture_path(g, m.get("emissiveTexture"));
            // emissiveFactor: never read as a constant by the reference (mesh.cpp:455, see load_mtl)
            mats.push_back(r);
        }
    std::vector<RawMesh> raws;
    const JVal* meshes = g.root.array("meshes");
    if (!meshes) return fail(HRA_ERR_FORMAT, "%s: no meshes", path.c_str());
    for (const JVal& m : meshes->arr)
    {
        const JVal* prims = m.array("primitives");
        if (!prims) continue;
        for (const JVal& p : prims->arr)
        {
            const long mode = p.integer("mode", 4);
            if (mode != 4) return fail(HRA_ERR_UNSUPPORTED, "%s: primitive mode %ld (only triangle lists)", path.c_str(), mode);
            const JVal* attr = p.get("attributes");
            if (!attr || attr->integer("POSITION", -1) < 0) continue;
            RawMesh r;
            r.material = (int)p.integer("material", -1);
            std::vector<float> f;
            size_t n = 0, n2 = 0;
            int rc = gltf_accessor(g, attr->integer("POSITION", -1), 3, &f, nullptr, &n);
            if (rc) return rc;
            r.pos.resize(n);
            for (size_t i = 0; i < n; i++) r.pos[i] = { f[3 * i], f[3 * i + 1], f[3 * i + 2] };
            if (attr->integer("NORMAL", -1) >= 0)
            {
                if ((rc = gltf_accessor(g, attr->integer("NORMAL", -1), 3, &f, nullptr, &n2))) return rc;
                if (n2 != n) return fail(HRA_ERR_FORMAT, "%s: NORMAL count differs from POSITION count", path.c_str());
                r.nrm.resize(n);
                for (size_t i = 0; i < n; i++) r.nrm[i] = { f[3 * i], f[3 * i + 1], f[3 * i + 2] };
                r.has_nrm = true;
            }
            if (attr->integer("TEXCOORD_0", -1) >= 0)
            {
                if ((rc = gltf_accessor(g, attr->integer("TEXCOORD_0", -1), 2, &f, nullptr, &n2))) return rc;
                if (n2 != n) return fail(HRA_ERR_FORMAT, "%s: TEXCOORD_0 count differs from POSITION count", path.c_str());
                r.uv.resize(2 * n);
                for (size_t i = 0; i < n; i++) { r.uv[2 * i] = f[2 * i]; r.uv[2 * i + 1] = f[2 * i + 1]; } // assimp's glTF2 importer flips v on import and aiProcess_FlipUVs flips it back: the file's value
                r.has_uv = true;
            }
            else r.uv.assign(2 * n, 0.0f);
            if (attr->integer("TANGENT", -1) >= 0 && r.has_nrm)
            {
                if ((rc = gltf_accessor(g, attr->integer("TANGENT", -1), 4, &f, nullptr, &n2))) return rc;
                if (n2 != n) return fail(HRA_ERR_FORMAT, "%s: TANGENT count differs from POSITION count", path.c_str());
                r.tan.resize(n); r.bit.resize(n);
                for (size_t i = 0; i < n; i++)
                {
                    r.tan[i] = { f[4 * i], f[4 * i + 1], f[4 * i + 2] };
                    r.bit[i] = cross(r.nrm[i], r.tan[i]) * f[4 * i + 3]; // bitangent = cross(normal, tangent.xyz) * tangent.w (glTF 2.0 §3.7.2.1)
                }
                r.has_tan = true;
            }
            if (p.integer("indices", -1) >= 0)
            {
                if ((rc = gltf_accessor(g, p.integer("indices", -1), 0, nullptr, &r.idx, &n2))) return rc;
            }
            else { r.idx.resize(n); for (size_t i = 0; i < n; i++) r.idx[i] = (uint32_t)i; }
            raws.push_back(std::move(r));
        }
    }
    return assemble(out, raws, mats, default_material(1.0f, 1.0f));
}

} // namespace

// =====================================================================================================================
// C interface
// =====================================================================================================================
extern "C" {

const char* hra_last_error(void) { return g_err.c_str(); }

// no C++ exception may cross the C boundary (std::bad_alloc on absurd sizes, std::length_error ...)
#define HRA_GUARDED(expr)                                                                       \
    try { return (expr); }                                                                      \
    catch (const std::exception& e) { return fail(HRA_ERR_IO, "out of memory or internal error: %s", e.what()); }

int hra_image_load_memory(const uint8_t* bytes, size_t n, int flip_vertical, int* width, int* height, int* channels, uint8_t** data)
{
    if (!bytes || !width || !height || !channels || !data) return fail(HRA_ERR_INVALID_ARG, "hra_image_load_memory: null argument");
    HRA_GUARDED(png_decode(bytes, n, flip_vertical != 0, width, height, channels, data));
}
int hra_image_load(const char* path, int flip_vertical, int* width, int* height, int* channels, uint8_t** data)
{
    if (!path) return fail(HRA_ERR_INVALID_ARG, "hra_image_load: null path");
    std::vector<uint8_t> file;
    if (!read_file(path, file)) return fail(HRA_ERR_IO, "cannot read %s", path);
    return hra_image_load_memory(file.data(), file.size(), flip_vertical, width, height, channels, data);
}
int hra_image_loadf(const char* path, int flip_vertical, int* width, int* height, float** rgba)
{
    if (!path || !width || !height || !rgba) return fail(HRA_ERR_INVALID_ARG, "hra_image_loadf: null argument");
    std::vector<uint8_t> file;
    if (!read_file(path, file)) return fail(HRA_ERR_IO, "cannot read %s", path);
    HRA_GUARDED(hdr_decode(file.data(), file.size(), flip_vertical != 0, width, height, rgba));
}
void hra_image_free(void* data) { free(data); }

int hra_image_save_png(const char* path, int width, int height, int channels, const uint8_t* data)
{
    if (!path || !data || width <= 0 || height <= 0 || channels < 1 || channels > 4) return fail(HRA_ERR_INVALID_ARG, "hra_image_save_png: bad argument");
    static const uint8_t ctype[5] = { 0, 0, 4, 2, 6 };
    std::vector<uint8_t> out = { 137, 80, 78, 71, 13, 10, 26, 10 }, body;
    put_be32(body, (uint32_t)width); put_be32(body, (uint32_t)height);
    body.push_back(8); body.push_back(ctype[channels]); body.push_back(0); body.push_back(0); body.push_back(0);
    png_chunk(out, "IHDR", body);
    // scanlines with filter type 0, wrapped in a zlib stream of stored blocks (<= 65535 bytes each)
    const size_t stride = (size_t)width * channels;
    std::vector<uint8_t> raw;
    raw.reserve((stride + 1) * (size_t)height);
    for (int y = 0; y < height; y++) { raw.push_back(0); raw.insert(raw.end(), data + stride * (size_t)y, data + stride * (size_t)(y + 1)); }
    body.clear();
    body.push_back(0x78); body.push_back(0x01);
    uint32_t a = 1, b = 0; // adler32
    for (size_t pos = 0; pos < raw.size();)
    {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        body.push_back(pos + n == raw.size() ? 1 : 0);
        body.push_back((uint8_t)(n & 255)); body.push_back((uint8_t)(n >> 8)); body.push_back((uint8_t)(~n & 255)); body.push_back((uint8_t)((~n >> 8) & 255));
        body.insert(body.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = 0; i < n; i++) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
        pos += n;
    }
    put_be32(body, (b << 16) | a);
    png_chunk(out, "IDAT", body);
    png_chunk(out, "IEND", {});
    FILE* f = fopen(path, "wb");
    if (!f) return fail(HRA_ERR_IO, "cannot write %s", path);
    const size_t w = fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    return w == out.size() ? HRA_OK : fail(HRA_ERR_IO, "short write to %s", path);
}

int hra_bluenoise_load(const char* dir, uint8_t* sobol, uint8_t* sr, uint32_t* slots_loaded)
{
    if (!dir || !sobol || !sr) return fail(HRA_ERR_INVALID_ARG, "hra_bluenoise_load: null argument");
    std::string d = dir;
    if (!d.empty() && d.back() != '/') d.push_back('/');
    auto load_rgba = [&](const std::string& p, int w, int h, uint8_t* dst) -> int {
        int iw, ih, ic;
        uint8_t* px = nullptr;
        int rc = hra_image_load(p.c_str(), 0, &iw, &ih, &ic, &px);
        if (rc) return rc;
        if (iw != w || ih != h || ic != 4) { free(px); return fail(HRA_ERR_FORMAT, "%s: expected %d x %d RGB(A), got %d x %d with %d channels", p.c_str(), w, h, iw, ih, ic); }
        memcpy(dst, px, (size_t)w * h * 4);
        free(px);
        return HRA_OK;
    };
    int rc = load_rgba(d + "sobol_256_4d.png", 256, 1, sobol); // kSOBOL_TEXTURE, blue_noise.cpp:5
    if (rc) return rc;
    uint32_t mask = 0;
    for (int s = 0; s < 9; s++)
    { // kSCRAMBLING_RANKING_TEXTURES, blue_noise.cpp:9-19
        const std::string p = d + "scrambling_ranking_128x128_2d_" + std::to_string(1 << s) + "spp.png";
        FILE* f = fopen(p.c_str(), "rb");
        if (!f) { if (s == 0) return fail(HRA_ERR_IO, "cannot read %s", p.c_str()); continue; }
        fclose(f);
        if ((rc = load_rgba(p, 128, 128, sr + (size_t)s * 128 * 128 * 4))) return rc;
        mask |= 1u << s;
    }
    if (slots_loaded) *slots_loaded = mask;
    return HRA_OK;
}

int hra_brdf_lut_load(const char* path, uint16_t* out)
{
    if (!path || !out) return fail(HRA_ERR_INVALID_ARG, "hra_brdf_lut_load: null argument");
    std::vector<uint8_t> file;
    if (!read_file(path, file)) return fail(HRA_ERR_IO, "cannot read %s", path);
    const size_t want = 512ull * 512 * 2 * sizeof(uint16_t); // BRDF_LUT_SIZE 512, RG16F (brdf_preintegrate_lut.cpp:8-31)
    if (file.size() < want) return fail(HRA_ERR_FORMAT, "%s: %zu bytes, expected %zu", path, file.size(), want);
    memcpy(out, file.data(), want);
    return HRA_OK;
}

int hra_environment_constant(const char* hdr_path, float rgb[3])
{
    int w, h;
    float* px = nullptr;
    int rc = hra_image_loadf(hdr_path, 0, &w, &h, &px);
    if (rc) return rc;
    double acc[3] = { 0, 0, 0 }, wsum = 0;
    for (int y = 0; y < h; y++)
    {
        const double wgt = std::sin(3.14159265358979323846 * (y + 0.5) / h); // solid angle of an equirectangular row
        double row[3] = { 0, 0, 0 };
        for (int x = 0; x < w; x++)
            for (int c = 0; c < 3; c++) row[c] += px[((size_t)y * w + x) * 4 + c];
        for (int c = 0; c < 3; c++) acc[c] += wgt * row[c];
        wsum += wgt * w;
    }
    free(px);
    for (int c = 0; c < 3; c++) rgb[c] = (float)(acc[c] / wsum);
    return HRA_OK;
}

int hra_mesh_load(const char* path, hra_mesh** out)
{
    if (!path || !out) return fail(HRA_ERR_INVALID_ARG, "hra_mesh_load: null argument");
    hra_mesh* m = new hra_mesh();
    m->path = path;
    const std::string e = ext_of(path);
    int rc;
    try
    {
        if (e == "obj") rc = load_obj(path, m);
        else if (e == "gltf" || e == "glb") rc = load_gltf(path, m);
        else rc = fail(HRA_ERR_UNSUPPORTED, "%s: unsupported mesh format '%s' (obj, gltf, glb)", path, e.c_str());
    }
    catch (const std::exception& ex) { rc = fail(HRA_ERR_IO, "%s: out of memory or internal error: %s", path, ex.what()); }
    if (rc) { delete m; *out = nullptr; return rc; }
    *out = m;
    return HRA_OK;
}
void hra_mesh_destroy(hra_mesh* m) { delete m; }
void hra_mesh_counts(const hra_mesh* m, uint64_t* nv, uint64_t* ni, uint64_t* ns, uint64_t* nm)
{
    if (nv) *nv = m->vertices.size();
    if (ni) *ni = m->indices.size();
    if (ns) *ns = m->submeshes.size();
    if (nm) *nm = m->materials.size();
}
const hr_vertex*   hra_mesh_vertices(const hra_mesh* m) { return m->vertices.data(); }
const uint32_t*    hra_mesh_indices(const hra_mesh* m) { return m->indices.data(); }
const hr_material* hra_mesh_materials(const hra_mesh* m) { return m->material_consts.data(); }
int hra_mesh_submesh(const hra_mesh* m, uint32_t i, hra_submesh* out)
{
    if (!m || !out || i >= m->submeshes.size()) return fail(HRA_ERR_INVALID_ARG, "hra_mesh_submesh: index out of range");
    *out = m->submeshes[i];
    return HRA_OK;
}
void hra_mesh_extents(const hra_mesh* m, float mn[3], float mx[3]) { memcpy(mn, m->mn, 12); memcpy(mx, m->mx, 12); }
const char* hra_mesh_material_texture(const hra_mesh* m, uint32_t material, int kind)
{
    if (!m || material >= m->materials.size() || kind < 0 || kind > 4) return "";
    return m->materials[material].tex[kind].c_str();
}

hra_scene* hra_scene_create(void) { return new hra_scene(); }
void       hra_scene_destroy(hra_scene* s) { delete s; }
int hra_scene_add_instance(hra_scene* s, const hra_mesh* mesh, const float model16[16])
{
    if (!s || !mesh || !model16) return fail(HRA_ERR_INVALID_ARG, "hra_scene_add_instance: null argument");
    hra_scene::Inst in;
    in.mesh = mesh;
    memcpy(in.model, model16, 64);
    s->insts.push_back(in);
    s->finalized = false;
    return HRA_OK;
}
int hra_scene_finalize(hra_scene* s)
{
    if (!s || s->insts.empty()) return fail(HRA_ERR_INVALID_ARG, "hra_scene_finalize: no instances");
    s->vertices.clear(); s->indices.clear(); s->instances.clear(); s->materials.clear();
    s->tex_pixels.clear(); s->textures.clear(); s->bindings.clear(); s->tex_warnings.clear();
    std::map<std::pair<std::string, int>, int32_t> tex_index; // (path, sRGB) -> texture index (-1: unusable)
    auto use_texture = [&](const std::string& path, bool srgb) -> int32_t {
        if (path.empty()) return -1;
        auto key = std::make_pair(path, srgb ? 1 : 0);
        auto it  = tex_index.find(key);
        if (it != tex_index.end()) return it->second;
        int32_t idx = -1;
        int w = 0, h = 0, c = 0;
        uint8_t* px = nullptr;
        if (ext_of(path) != "png") s->tex_warnings += path + ": not a PNG (only PNG textures are decoded), constant used\n";
        else if (hra_image_load(path.c_str(), 0, &w, &h, &c, &px) != HRA_OK) s->tex_warnings += path + ": " + g_err + ", constant used\n";
        else
        {
            idx = (int32_t)s->tex_pixels.size();
            s->tex_pixels.emplace_back(px, px + (size_t)w * h * c);
            free(px);
            hr_texture t;
            t.width = w; t.height = h; t.channels = c; t.srgb = srgb ? 1 : 0; t.data = nullptr; // pointers are set once all images are in place
            s->textures.push_back(t);
        }
        tex_index[key] = idx;
        return idx;
    };
    struct Off { uint32_t vertex, index, material; };
    std::map<const hra_mesh*, Off> seen; // m_local_to_global_mesh_idx, ray_traced_scene.cpp:283-297
    for (const auto& in : s->insts)
    {
        auto it = seen.find(in.mesh);
        if (it == seen.end())
        {
            if (s->vertices.size() + in.mesh->vertices.size() > 0xFFFFFFFFull || s->indices.size() + in.mesh->indices.size() > 0xFFFFFFFFull)
                return fail(HRA_ERR_UNSUPPORTED, "hra_scene_finalize: more than 2^32 vertices or indices (hr_instance offsets are 32-bit)");
            Off o { (uint32_t)s->vertices.size(), (uint32_t)s->indices.size(), (uint32_t)s->materials.size() };
            s->vertices.insert(s->vertices.end(), in.mesh->vertices.begin(), in.mesh->vertices.end());
            s->indices.insert(s->indices.end(), in.mesh->indices.begin(), in.mesh->indices.end());
            s->materials.insert(s->materials.end(), in.mesh->material_consts.begin(), in.mesh->material_consts.end());
            const bool gltf = ext_of(in.mesh->path) != "obj";
            for (const auto& m : in.mesh->materials)
            { // Material::load: albedo sRGB (material.cpp:114), the rest linear; packed glTF image: roughness .g, metallic .b (mesh.cpp:411,437)
                hr_material_textures b;
                b.albedo            = use_texture(m.tex[HRA_TEX_ALBEDO], true);
                b.normal            = use_texture(m.tex[HRA_TEX_NORMAL], false);
                b.roughness         = use_texture(m.tex[HRA_TEX_ROUGHNESS], false);
                b.roughness_channel = gltf ? 1 : 0;
                b.metallic          = use_texture(m.tex[HRA_TEX_METALLIC], false);
                b.metallic_channel  = gltf ? 2 : 0;
                b.emissive          = use_texture(m.tex[HRA_TEX_EMISSIVE], false);
                s->bindings.push_back(b);
            }
            it = seen.emplace(in.mesh, o).first;
        }
        for (const auto& sm : in.mesh->submeshes)
        { // one draw per (instance, sub-mesh): its running index is the mesh id of the G-buffer pass (g_buffer.cpp:141-175)
            hr_instance hi;
            memcpy(hi.model, in.model, 64);
            hi.first_index  = it->second.index + sm.base_index;
            hi.index_count  = sm.index_count;
            hi.base_vertex  = it->second.vertex;
            hi.material_idx = it->second.material + sm.mat_idx;
            s->instances.push_back(hi);
        }
    }
    for (size_t i = 0; i < s->textures.size(); i++) s->textures[i].data = s->tex_pixels[i].data();
    s->finalized = true;
    return HRA_OK;
}
void hra_scene_texture_counts(const hra_scene* s, uint64_t* nt, uint64_t* nb)
{
    if (nt) *nt = s->textures.size();
    if (nb) *nb = s->bindings.size();
}
const hr_texture*           hra_scene_textures(const hra_scene* s) { return s->textures.data(); }
const hr_material_textures* hra_scene_material_textures(const hra_scene* s) { return s->bindings.data(); }
const char*                 hra_scene_texture_warnings(const hra_scene* s) { return s->tex_warnings.c_str(); }
void hra_scene_counts(const hra_scene* s, uint64_t* nv, uint64_t* ni, uint64_t* nin, uint64_t* nm)
{
    if (nv) *nv = s->vertices.size();
    if (ni) *ni = s->indices.size();
    if (nin) *nin = s->instances.size();
    if (nm) *nm = s->materials.size();
}
const hr_vertex*   hra_scene_vertices(const hra_scene* s) { return s->vertices.data(); }
const uint32_t*    hra_scene_indices(const hra_scene* s) { return s->indices.data(); }
const hr_instance* hra_scene_instances(const hra_scene* s) { return s->instances.data(); }
const hr_material* hra_scene_materials(const hra_scene* s) { return s->materials.data(); }

} // extern "C"
