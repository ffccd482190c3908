// rt_loader.cpp — OBJ + MTL + textures model-folder loader with the semantics of the reference's
// getTrianglesData_ (RayTracing/Assets/headers/mesh.h:279-613; citations M:line), re-hosted on
// std::filesystem (the reference needs <windows.h> for its directory listing, external/filesUtil/
// myFile.cpp:6) and with a small PNG decoder on zlib in place of stb_image (stb is vendored only
// inside the reference tree).  Checked against the reference's own loader, compiled from where it
// lies, in tests/test_loader_cpu.py: triangles, material table and decoded texture bytes identical.
//
// Kept quirks (they decide which material a triangle gets, so a drop-in must keep them):
//   * `materials[]` is filled in std::map order — library file name, then material name, with the
//     `_default_` library among them — while RTXTriangle.materialIndex is the running `newmtl`
//     counter in MTL-file order (M:321-324,369,455-462,602);
//   * texture index = position of the file name in the directory listing (sorted here), M:308-317;
//   * `Ke` with any positive component turns the material into a LIGHT whose strength is the Rec.601
//     luma of Ke (M:378-391); `map_Kd` makes it TEXTURE unless it is LIGHT / GLASS(_HIGHLIGHT)
//     (M:430-450); `GlassHighlight`, `EDGE_HIGHLIGHT` keys (M:416-429); `Ni` is ignored (M:399-415);
//   * faces must be triangles written with exactly three spaces on the line (M:501-506); vertex
//     forms v, v/vt, v/vt/vn, v//vn (M:516-590); (aTex,bTex,cTex) = (vt1, vt2, vt0) (M:602-606);
//   * the .obj used is the first one of the (sorted) listing (M:223-253).
// Differences, on purpose: fields the reference leaves uninitialised (UVs of untextured faces,
// unused Material floats, RTXTriangle.pad) are zero; errors are returned, not thrown as ints.
// Textures: PNG (all colour types, non-interlaced) and baseline JPEG without chroma subsampling — every
// texture the reference ships.
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "rt_host.h"

namespace fs = std::filesystem;

namespace {

thread_local std::string g_lerr;

std::vector<std::string> listFiles(const fs::path& dir) {  // getFilenamesInFolder, sorted
    std::vector<std::string> names;
    std::error_code ec;
    if (!fs::exists(dir, ec)) return names;
    for (const auto& e : fs::directory_iterator(dir, ec))
        if (e.is_regular_file()) names.push_back(e.path().filename().string());
    std::sort(names.begin(), names.end());
    return names;
}

// ------------------------------------------------------------------------------------------------ PNG
struct Image {
    int w = 0, h = 0, ch = 0;
    std::vector<uint8_t> px;
};
inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// 8-bit output with the channel conventions of stbi_load(path, &w, &h, &n, 0): gray 1, gray+alpha 2,
// RGB 3, RGBA 4, palette 3 (4 with tRNS), +1 channel for tRNS on gray / RGB, sub-byte gray scaled to
// 0..255, 16-bit samples reduced to their high byte.  Non-interlaced files only.
bool decodePng(const std::vector<uint8_t>& file, Image& out, std::string& err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 33 || memcmp(file.data(), sig, 8) != 0) { err = "not a PNG file"; return false; }
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool sawIhdr = false;
    while (pos + 8 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        if (pos + 12 + (size_t)len > file.size()) { err = "truncated PNG chunk"; return false; }
        const uint8_t* data = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) { err = "bad IHDR"; return false; }
            w = be32(data); h = be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            sawIhdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!memcmp(type, "tRNS", 4)) {
            trns.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!sawIhdr || w == 0 || h == 0 || w > 65536 || h > 65536) { err = "bad PNG header"; return false; }
    if (interlace) { err = "interlaced PNG not supported"; return false; }
    int fileCh;
    switch (ctype) {
        case 0: fileCh = 1; break;
        case 2: fileCh = 3; break;
        case 3: fileCh = 1; break;
        case 4: fileCh = 2; break;
        case 6: fileCh = 4; break;
        default: err = "bad PNG colour type"; return false;
    }
    if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16) ||
        (depth < 8 && ctype != 0 && ctype != 3) || (depth == 16 && ctype == 3)) { err = "bad PNG bit depth"; return false; }
    const size_t rowBytes = ((size_t)w * fileCh * depth + 7) / 8;
    const int bpp = std::max(1, fileCh * depth / 8);  // filter unit
    std::vector<uint8_t> raw((rowBytes + 1) * h);
    uLongf rawLen = (uLongf)raw.size();
    const int zr = uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || rawLen != raw.size()) { err = "PNG inflate failed"; return false; }
    // unfilter in place
    std::vector<uint8_t> img(rowBytes * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t ft = raw[(rowBytes + 1) * y];
        const uint8_t* src = &raw[(rowBytes + 1) * y + 1];
        uint8_t* dst = &img[rowBytes * y];
        const uint8_t* up = y ? &img[rowBytes * (y - 1)] : nullptr;
        for (size_t x = 0; x < rowBytes; x++) {
            const int a = x >= (size_t)bpp ? dst[x - bpp] : 0;
            const int b = up ? up[x] : 0;
            const int c = (up && x >= (size_t)bpp) ? up[x - bpp] : 0;
            int v = src[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: err = "bad PNG filter"; return false;
            }
            dst[x] = (uint8_t)v;
        }
    }
    // samples -> 8 bit per channel
    std::vector<uint8_t> s8((size_t)w * h * fileCh);
    if (depth == 8) {
        s8 = img;
    } else if (depth == 16) {
        for (size_t i = 0; i < s8.size(); i++) s8[i] = img[2 * i];  // high byte
    } else {
        static const int scaleTab[5] = {0, 0xff, 0x55, 0, 0x11};
        const int scale = ctype == 0 ? scaleTab[depth] : 1;  // palette indices are not scaled
        for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++) {
                const size_t bit = (size_t)x * depth;
                const uint8_t byte = img[rowBytes * y + bit / 8];
                const int v = (byte >> (8 - depth - (int)(bit % 8))) & ((1 << depth) - 1);
                s8[(size_t)y * w + x] = (uint8_t)(v * scale);
            }
    }
    // expand palette / tRNS
    if (ctype == 3) {
        if (palette.empty()) { err = "PNG palette missing"; return false; }
        const int outCh = trns.empty() ? 3 : 4;
        out.px.resize((size_t)w * h * outCh);
        for (size_t i = 0; i < (size_t)w * h; i++) {
            const size_t k = s8[i];
            for (int c = 0; c < 3; c++) out.px[i * outCh + c] = 3 * k + c < palette.size() ? palette[3 * k + c] : 0;
            if (outCh == 4) out.px[i * 4 + 3] = k < trns.size() ? trns[k] : 255;
        }
        out.ch = outCh;
    } else if (!trns.empty() && (ctype == 0 || ctype == 2) && trns.size() >= (size_t)fileCh * 2) {
        const int outCh = fileCh + 1;
        uint8_t key[3] = {0, 0, 0};
        for (int c = 0; c < fileCh; c++) {  // tRNS holds 16-bit values
            const int v16 = trns[2 * c] << 8 | trns[2 * c + 1];
            static const int scaleTab[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
            key[c] = depth == 16 ? (uint8_t)(v16 >> 8) : (uint8_t)((v16 & 255) * (ctype == 0 ? scaleTab[depth] : 1));
        }
        out.px.resize((size_t)w * h * outCh);
        for (size_t i = 0; i < (size_t)w * h; i++) {
            bool match = true;
            for (int c = 0; c < fileCh; c++) {
                out.px[i * outCh + c] = s8[i * fileCh + c];
                match = match && s8[i * fileCh + c] == key[c];
            }
            out.px[i * outCh + fileCh] = match ? 0 : 255;
        }
        out.ch = outCh;
    } else {
        out.px.swap(s8);
        out.ch = fileCh;
    }
    out.w = (int)w;
    out.h = (int)h;
    return true;
}

// ------------------------------------------------------------------------------------------------ JPEG
// Baseline sequential JPEG (SOF0, 8 bit, Huffman), 1 or 3 components without chroma subsampling — which is
// what every JPEG the reference ships is (RayTracing/Data/{robot,plants,toonHouse}/textures).  The reference
// decodes through stb_image 2.30 (external/stb, public domain); JPEG decoding is only specified up to the
// accuracy of the IDCT and of the colour conversion, so to hand the SAME texels to the renderer this decoder
// restates the arithmetic stb_image uses for those two steps: the Loeffler–Ligtenberg–Moschytz integer IDCT
// with 12-bit constants (column pass >> 10 after +512, row pass >> 17 after +65536 + (128 << 17)) and the
// 20-bit fixed-point YCbCr → RGB conversion.  tests/test_loader_cpu.py compares the result byte for byte
// with the reference's own loader.  Subsampled or progressive files are rejected with an error.
struct JpegDecoder {
    const uint8_t* d;
    size_t n, pos = 0;
    std::string err;
    int W = 0, H = 0, ncomp = 0;
    struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0; } comp[3];
    uint16_t qt[4][64] = {};
    struct Huff {
        bool present = false;
        uint8_t size[257];
        uint16_t code[256];
        uint8_t val[256];
        int maxcode[18];
        int delta[17];
        int count = 0;
    } huff[2][4];
    int restart = 0;
    uint32_t bitbuf = 0;
    int bitcnt = 0;
    bool hitMarker = false;
    uint8_t marker = 0;

    JpegDecoder(const uint8_t* data, size_t len) : d(data), n(len) {}
    bool fail(const char* m) { if (err.empty()) err = m; return false; }
    int u8() { return pos < n ? d[pos++] : 0; }
    int u16() { const int a = u8(); return a << 8 | u8(); }

    bool buildHuff(Huff& h, const uint8_t counts[16]) {
        int k = 0;
        for (int i = 0; i < 16; i++)
            for (int j = 0; j < counts[i]; j++) h.size[k++] = (uint8_t)(i + 1);
        h.size[k] = 0;
        h.count = k;
        int code = 0;
        k = 0;
        for (int j = 1; j <= 16; j++) {
            h.delta[j] = k - code;
            if (h.size[k] == j) {
                while (h.size[k] == j) h.code[k++] = (uint16_t)code++;
                if (code - 1 >= (1 << j)) return fail("bad JPEG code lengths");
            }
            h.maxcode[j] = code << (16 - j);
            code <<= 1;
        }
        h.maxcode[17] = 0x7fffffff;
        h.present = true;
        return true;
    }
    void fill() {
        while (bitcnt <= 24) {
            int b = 0;
            if (!hitMarker) {
                b = pos < n ? d[pos++] : 0;
                if (b == 0xff) {
                    int c = pos < n ? d[pos++] : 0;
                    while (c == 0xff) c = pos < n ? d[pos++] : 0;
                    if (c != 0) {
                        marker = (uint8_t)c;
                        hitMarker = true;
                        b = 0;
                    }
                }
            }
            bitbuf |= (uint32_t)b << (24 - bitcnt);
            bitcnt += 8;
        }
    }
    int decodeSym(const Huff& h) {
        if (bitcnt < 16) fill();
        const int top = (int)(bitbuf >> 16);
        int len = 1;
        while (len <= 16 && top >= h.maxcode[len]) len++;
        if (len > 16 || len > bitcnt) return -1;
        const int idx = (int)((bitbuf >> (32 - len)) & ((1u << len) - 1u)) + h.delta[len];
        if (idx < 0 || idx >= h.count) return -1;
        bitbuf <<= len;
        bitcnt -= len;
        return h.val[idx];
    }
    int receiveExtend(int nb) {
        if (nb == 0) return 0;
        if (bitcnt < nb) fill();
        const int v = (int)(bitbuf >> (32 - nb));
        bitbuf <<= nb;
        bitcnt -= nb;
        return v < (1 << (nb - 1)) ? v - (1 << nb) + 1 : v;  // EXTEND of ITU T.81 F.2.2.1
    }
    static uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

    // LL&M integer IDCT, 12-bit constants; data in natural order, already dequantised
    static void idct(uint8_t* out, int stride, const short data[64]) {
        auto f2f = [](double x) { return (int)(x * 4096 + 0.5); };
        static const int c0541 = f2f(0.5411961), c1847 = f2f(-1.847759065), c0765 = f2f(0.765366865),
                         c1175 = f2f(1.175875602), c0298 = f2f(0.298631336), c2053 = f2f(2.053119869),
                         c3072 = f2f(3.072711026), c1501 = f2f(1.501321110), c0899 = f2f(-0.899976223),
                         c2562 = f2f(-2.562915447), c1961 = f2f(-1.961570560), c0390 = f2f(-0.390180644);
        auto pass = [&](int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7, int& x0, int& x1, int& x2,
                        int& x3, int& t0, int& t1, int& t2, int& t3) {
            int p1 = (s2 + s6) * c0541;
            int a2 = p1 + s6 * c1847, a3 = p1 + s2 * c0765;
            int a0 = (s0 + s4) * 4096, a1 = (s0 - s4) * 4096;
            x0 = a0 + a3; x3 = a0 - a3; x1 = a1 + a2; x2 = a1 - a2;
            t0 = s7; t1 = s5; t2 = s3; t3 = s1;
            int p3 = t0 + t2, p4 = t1 + t3;
            p1 = t0 + t3;
            int p2 = t1 + t2;
            const int p5 = (p3 + p4) * c1175;
            t0 *= c0298; t1 *= c2053; t2 *= c3072; t3 *= c1501;
            p1 = p5 + p1 * c0899; p2 = p5 + p2 * c2562; p3 *= c1961; p4 *= c0390;
            t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;
        };
        int val[64];
        for (int i = 0; i < 8; i++) {
            const short* dd = data + i;
            int* v = val + i;
            if (dd[8] == 0 && dd[16] == 0 && dd[24] == 0 && dd[32] == 0 && dd[40] == 0 && dd[48] == 0 && dd[56] == 0) {
                const int dc = dd[0] * 4;
                v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dc;
            } else {
                int x0, x1, x2, x3, t0, t1, t2, t3;
                pass(dd[0], dd[8], dd[16], dd[24], dd[32], dd[40], dd[48], dd[56], x0, x1, x2, x3, t0, t1, t2, t3);
                x0 += 512; x1 += 512; x2 += 512; x3 += 512;
                v[0] = (x0 + t3) >> 10; v[56] = (x0 - t3) >> 10;
                v[8] = (x1 + t2) >> 10; v[48] = (x1 - t2) >> 10;
                v[16] = (x2 + t1) >> 10; v[40] = (x2 - t1) >> 10;
                v[24] = (x3 + t0) >> 10; v[32] = (x3 - t0) >> 10;
            }
        }
        for (int i = 0; i < 8; i++) {
            const int* v = val + 8 * i;
            uint8_t* o = out + (size_t)i * stride;
            int x0, x1, x2, x3, t0, t1, t2, t3;
            pass(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], x0, x1, x2, x3, t0, t1, t2, t3);
            const int bias = 65536 + (128 << 17);
            x0 += bias; x1 += bias; x2 += bias; x3 += bias;
            o[0] = clamp8((x0 + t3) >> 17); o[7] = clamp8((x0 - t3) >> 17);
            o[1] = clamp8((x1 + t2) >> 17); o[6] = clamp8((x1 - t2) >> 17);
            o[2] = clamp8((x2 + t1) >> 17); o[5] = clamp8((x2 - t1) >> 17);
            o[3] = clamp8((x3 + t0) >> 17); o[4] = clamp8((x3 - t0) >> 17);
        }
    }

    bool decode(Image& out) {
        static const uint8_t zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                           41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                           30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        if (n < 4 || d[0] != 0xff || d[1] != 0xd8) return fail("not a JPEG file");
        pos = 2;
        bool sawSof = false;
        for (;;) {
            int m = u8();
            while (m != 0xff && pos < n) m = u8();
            while (m == 0xff && pos < n) m = u8();
            if (pos >= n) return fail("JPEG ended before the scan");
            if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) continue;
            const int len = u16();
            if (len < 2 || pos + (size_t)len - 2 > n) return fail("bad JPEG segment");
            const size_t end = pos + (size_t)len - 2;
            if (m == 0xdb) {  // DQT
                while (pos < end) {
                    const int pq = u8();
                    const int t = pq & 15, wide = pq >> 4;
                    if (t > 3) return fail("bad DQT");
                    for (int i = 0; i < 64; i++) qt[t][zigzag[i]] = (uint16_t)(wide ? u16() : u8());
                }
            } else if (m == 0xc4) {  // DHT
                while (pos < end) {
                    const int tc = u8();
                    const int cls = tc >> 4, id = tc & 15;
                    if (cls > 1 || id > 3) return fail("bad DHT");
                    uint8_t counts[16];
                    int total = 0;
                    for (int i = 0; i < 16; i++) { counts[i] = (uint8_t)u8(); total += counts[i]; }
                    if (total > 256) return fail("bad DHT");
                    Huff& h = huff[cls][id];
                    if (!buildHuff(h, counts)) return false;
                    for (int i = 0; i < total; i++) h.val[i] = (uint8_t)u8();
                }
            } else if (m == 0xc0 || m == 0xc1) {  // SOF0 / SOF1 (extended sequential, Huffman)
                if (u8() != 8) return fail("only 8-bit JPEG");
                H = u16(); W = u16(); ncomp = u8();
                if (W <= 0 || H <= 0 || (ncomp != 1 && ncomp != 3)) return fail("unsupported JPEG layout");
                for (int i = 0; i < ncomp; i++) {
                    comp[i].id = u8();
                    const int hv = u8();
                    comp[i].h = hv >> 4; comp[i].v = hv & 15;
                    comp[i].tq = u8();
                    if (comp[i].tq > 3) return fail("bad SOF");
                }
                for (int i = 0; i < ncomp; i++)
                    if (comp[i].h != comp[0].h || comp[i].v != comp[0].v) return fail("chroma-subsampled JPEG not supported");
                sawSof = true;
            } else if (m == 0xc2) {
                return fail("progressive JPEG not supported");
            } else if (m == 0xdd) {
                restart = u16();
            } else if (m == 0xda) {  // SOS
                if (!sawSof) return fail("SOS before SOF");
                const int ns = u8();
                if (ns != ncomp) return fail("non-interleaved JPEG scan not supported");
                for (int i = 0; i < ns; i++) {
                    const int id = u8(), t = u8();
                    int k = 0;
                    while (k < ncomp && comp[k].id != id) k++;
                    if (k == ncomp) return fail("bad SOS");
                    comp[k].td = t >> 4; comp[k].ta = t & 15;
                    if (comp[k].td > 3 || comp[k].ta > 3 || !huff[0][comp[k].td].present || !huff[1][comp[k].ta].present)
                        return fail("missing Huffman table");
                }
                pos = end;
                break;
            }
            pos = end;
        }
        // ---- entropy-coded segment: one 8x8 block per component per MCU
        const int bw = (W + 7) / 8, bh = (H + 7) / 8;
        std::vector<uint8_t> plane[3];
        const int pw = bw * 8;
        for (int c = 0; c < ncomp; c++) plane[c].assign((size_t)pw * bh * 8, 0);
        int todo = restart ? restart : 0x7fffffff;
        for (int by = 0; by < bh; by++)
            for (int bx = 0; bx < bw; bx++) {
                for (int c = 0; c < ncomp; c++) {
                    short blk[64];
                    memset(blk, 0, sizeof blk);
                    const int t = decodeSym(huff[0][comp[c].td]);
                    if (t < 0 || t > 15) return fail("bad JPEG DC code");
                    const int diff = t ? receiveExtend(t) : 0;
                    comp[c].pred += diff;
                    blk[0] = (short)(comp[c].pred * qt[comp[c].tq][0]);
                    for (int k = 1; k < 64;) {
                        const int rs = decodeSym(huff[1][comp[c].ta]);
                        if (rs < 0) return fail("bad JPEG AC code");
                        const int r = rs >> 4, sz = rs & 15;
                        if (sz == 0) {
                            if (rs != 0xf0) break;  // EOB
                            k += 16;
                        } else {
                            k += r;
                            if (k > 63) return fail("bad JPEG run");
                            const int z = zigzag[k++];
                            blk[z] = (short)(receiveExtend(sz) * qt[comp[c].tq][z]);
                        }
                    }
                    idct(&plane[c][(size_t)by * 8 * pw + (size_t)bx * 8], pw, blk);
                }
                if (--todo <= 0) {  // restart interval: byte-align, expect RSTn, reset predictors
                    if (bitcnt < 24) fill();
                    if (!(hitMarker && marker >= 0xd0 && marker <= 0xd7)) {
                        if (!(by == bh - 1 && bx == bw - 1)) return fail("missing JPEG restart marker");
                    }
                    bitbuf = 0; bitcnt = 0; hitMarker = false;
                    for (int c = 0; c < ncomp; c++) comp[c].pred = 0;
                    todo = restart;
                }
            }
        // ---- output: grey, or YCbCr -> RGB in 20-bit fixed point
        out.w = W; out.h = H; out.ch = ncomp;
        out.px.resize((size_t)W * H * ncomp);
        const bool isRGB = ncomp == 3 && comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B';
        auto f2fix = [](float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; };
        const int kCrR = f2fix(1.40200f), kCrG = -f2fix(0.71414f), kCbG = -f2fix(0.34414f), kCbB = f2fix(1.77200f);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                uint8_t* o = &out.px[((size_t)y * W + x) * ncomp];
                const size_t i = (size_t)y * pw + x;
                if (ncomp == 1) {
                    o[0] = plane[0][i];
                } else if (isRGB) {
                    o[0] = plane[0][i]; o[1] = plane[1][i]; o[2] = plane[2][i];
                } else {
                    const int yf = (plane[0][i] << 20) + (1 << 19);
                    const int cr = plane[2][i] - 128, cb = plane[1][i] - 128;
                    const int r = yf + cr * kCrR;
                    const int g = yf + cr * kCrG + (int)(((unsigned)(cb * kCbG)) & 0xffff0000u);
                    const int b = yf + cb * kCbB;
                    o[0] = clamp8(r >> 20); o[1] = clamp8(g >> 20); o[2] = clamp8(b >> 20);
                }
            }
        return true;
    }
};

bool loadTexture(const fs::path& p, Image& img, std::string& err) {
    std::ifstream f(p, std::ios::binary);
    if (!f) { err = "cannot open " + p.string(); return false; }
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::string ext = p.extension().string();
    std::transform(ext.begin(), ext.end(), ext.begin(), ::tolower);
    if (ext == ".jpg" || ext == ".jpeg") {
        JpegDecoder jd(bytes.data(), bytes.size());
        if (!jd.decode(img)) { err = p.filename().string() + ": " + jd.err; return false; }
    } else if (ext == ".png") {
        if (!decodePng(bytes, img, err)) { err = p.filename().string() + ": " + err; return false; }
    } else {
        err = "texture " + p.filename().string() + ": only PNG and baseline JPEG are decoded by the host library";
        return false;
    }
    // stbi_set_flip_vertically_on_load(true), external/OpenGL/textureClass.cpp:65
    const size_t row = (size_t)img.w * img.ch;
    for (int y = 0; y < img.h / 2; y++)
        std::swap_ranges(img.px.begin() + row * y, img.px.begin() + row * (y + 1), img.px.begin() + row * (img.h - 1 - y));
    return true;
}

rt_material blankMaterial() {  // Material() of M:47 on zeroed storage
    rt_material m;
    memset(&m, 0, sizeof m);
    m.color[0] = m.color[1] = m.color[2] = 1.0f;
    m.textureIndex = -1;
    m.materialType = RT_MAT_DIFFUSE;
    return m;
}

std::vector<std::string> splitSpaces(const std::string& s) {  // split(line, ' '), M:156-175
    std::vector<std::string> out;
    std::string tok;
    for (char ch : s) {
        if (ch == ' ') {
            if (!tok.empty()) { out.push_back(tok); tok.clear(); }
        } else {
            tok += ch;
        }
    }
    if (!tok.empty()) out.push_back(tok);
    return out;
}

}  // namespace

extern "C" {

int rth_decode_png(const uint8_t* bytes, int64_t n, uint8_t* out, int64_t cap, int32_t* w, int32_t* h, int32_t* ch) {
    Image img;
    std::string err;
    if (!bytes || n <= 0 || !decodePng(std::vector<uint8_t>(bytes, bytes + n), img, err)) {
        rth_set_error(err.empty() ? "bad argument" : err.c_str());
        return RT_ERR_INVALID;
    }
    if (w) *w = img.w;
    if (h) *h = img.h;
    if (ch) *ch = img.ch;
    if (out) {
        if ((int64_t)img.px.size() > cap) { rth_set_error("output buffer too small"); return RT_ERR_INVALID; }
        memcpy(out, img.px.data(), img.px.size());
    }
    return RT_OK;
}

}  // extern "C"

// getTrianglesData_(folder, ...) — M:279-613.  The scene must be fresh (only the `_default_` material).
static int loadFolderImpl(rth_scene* s, const char* folder) {
    if (!s || !folder) { g_lerr = "bad argument"; return RT_ERR_INVALID; }
    if (rth_scene_triangle_count(s) != 0 || rth_scene_material_count(s) != 1) {
        g_lerr = "rth_load_model_folder needs a fresh scene";
        return RT_ERR_STATE;
    }
    const fs::path dir(folder);
    std::error_code ec;
    if (!fs::is_directory(dir, ec)) { g_lerr = std::string("not a directory: ") + folder; return RT_ERR_INVALID; }
    const std::vector<std::string> files = listFiles(dir);
    std::string objName;
    for (const auto& f : files)
        if (fs::path(f).extension() == ".obj") { objName = f; break; }  // findFirstObjFile, M:223-253
    if (objName.empty()) { g_lerr = "OBJ file not found"; return RT_ERR_INVALID; }

    // ---- textures (M:305-318)
    std::map<std::string, int> texFileToIndex;
    const std::vector<std::string> texNames = listFiles(dir / "textures");
    if ((int)texNames.size() > RT_MAX_TEXTURES) { g_lerr = "more than 5 textures"; return RT_ERR_INVALID; }
    for (size_t i = 0; i < texNames.size(); i++) {
        Image img;
        std::string err;
        if (!loadTexture(dir / "textures" / texNames[i], img, err)) { g_lerr = err; return RT_ERR_INVALID; }
        texFileToIndex[texNames[i]] = (int)i;
        if (rth_set_texture(s, (int)i, img.px.data(), img.w, img.h, img.ch) != RT_OK) { g_lerr = rth_last_error(); return RT_ERR_INVALID; }
    }

    // ---- MTL files (M:320-453)
    std::map<std::string, std::map<std::string, rt_material>> libToMtlMaps;
    {
        rt_material def = blankMaterial();
        def.index = 0;
        libToMtlMaps["_default_"]["_default_"] = def;
    }
    int matIndex = 0;
    for (const std::string& name : files) {
        const size_t dot = name.find('.');
        if (dot == std::string::npos) { g_lerr = "File extension not found: " + name; return RT_ERR_INVALID; }
        if (name.substr(dot + 1) != "mtl") continue;
        std::ifstream mtl(dir / name);
        if (!mtl.is_open()) continue;
        std::map<std::string, rt_material> nameToMtl;
        std::string mtlName, line;
        auto cur = [&]() -> rt_material& {
            auto it = nameToMtl.find(mtlName);
            if (it == nameToMtl.end()) it = nameToMtl.emplace(mtlName, blankMaterial()).first;  // operator[] of M:376
            return it->second;
        };
        while (std::getline(mtl, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            std::stringstream ss(line);
            std::string key;
            ss >> key;
            if (key == "newmtl") {
                ss >> mtlName;
                rt_material m = blankMaterial();
                m.index = ++matIndex;
                nameToMtl[mtlName] = m;
            } else if (key == "Kd") {
                float v[3] = {0, 0, 0};
                ss >> v[0] >> v[1] >> v[2];
                rt_material& m = cur();
                m.color[0] = v[0]; m.color[1] = v[1]; m.color[2] = v[2]; m.color[3] = 0.0f;
            } else if (key == "Ke") {
                float v[3] = {0, 0, 0};
                ss >> v[0] >> v[1] >> v[2];
                rt_material& m = cur();
                m.emissionColor[0] = v[0]; m.emissionColor[1] = v[1]; m.emissionColor[2] = v[2]; m.emissionColor[3] = 0.0f;
                if (v[0] > 0.0f || v[1] > 0.0f || v[2] > 0.0f) {
                    m.materialType = RT_MAT_LIGHT;  // makeLight, M:72-77
                    m.emissionStrength = 0.299f * v[0] + 0.587f * v[1] + 0.114f * v[2];
                } else {
                    m.emissionStrength = 0.0f;
                }
            } else if (key == "GlassHighlight") {
                rt_material& m = cur();
                if (m.materialType != RT_MAT_LIGHT) m.materialType = RT_MAT_GLASS_HIGHLIGHT;  // makeGlassHighlight keeps the colour
            } else if (key == "EDGE_HIGHLIGHT") {
                cur().isEdgeHighlight = 1;
            } else if (key == "map_Kd") {
                rt_material& m = cur();
                if (m.materialType != RT_MAT_LIGHT && m.materialType != RT_MAT_GLASS && m.materialType != RT_MAT_GLASS_HIGHLIGHT) {
                    std::string texName;
                    ss >> texName;
                    m.materialType = RT_MAT_TEXTURE;
                    auto it = texFileToIndex.find(texName);
                    if (it == texFileToIndex.end()) { g_lerr = "texture not found: " + texName; return RT_ERR_INVALID; }
                    m.textureIndex = it->second;
                }
            }
        }
        libToMtlMaps[name] = nameToMtl;
    }
    // ---- materials vector in map order (M:455-462); slot 0 of a fresh scene is replaced too
    std::vector<rt_material> materials;
    for (const auto& lib : libToMtlMaps)
        for (const auto& m : lib.second) materials.push_back(m.second);

    // ---- OBJ (M:464-610)
    std::ifstream obj(dir / objName);
    if (!obj.is_open()) { g_lerr = "Cannot find the OBJ path specified."; return RT_ERR_INVALID; }
    struct P3 { float x, y, z; };
    struct P2 { float x, y; };
    std::vector<P3> verts;
    std::vector<P2> texCoords;
    std::vector<rt_triangle> tris;
    std::string currentLib = "_default_", currentMtl = "_default_", line;
    while (std::getline(obj, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::stringstream ss(line);
        std::string type;
        ss >> type;
        if (type == "mtllib") ss >> currentLib;
        else if (type == "usemtl") ss >> currentMtl;
        else if (type == "v") {
            P3 v{0, 0, 0};
            ss >> v.x >> v.y >> v.z;
            verts.push_back(v);
        } else if (type == "vt") {
            P2 v{0, 0};
            ss >> v.x >> v.y;
            texCoords.push_back(v);
        } else if (type == "f") {
            if (std::count(line.begin(), line.end(), ' ') != 3) {
                g_lerr = "Invalid OBJ file, non-triangle face not supported. Line: " + line;
                return RT_ERR_INVALID;
            }
            const std::vector<std::string> face = splitSpaces(line);
            if (face.size() < 4) { g_lerr = "bad face line: " + line; return RT_ERR_INVALID; }
            const long slashes = std::count(face[1].begin(), face[1].end(), '/');
            const bool hasVt = slashes == 1 || (slashes == 2 && face[1].find("//") == std::string::npos);
            P3 p[3];
            P2 t[3] = {{0, 0}, {0, 0}, {0, 0}};
            for (int i = 0; i < 3; i++) {
                const std::string& tok = face[1 + i];
                const size_t s1 = tok.find('/');
                long vi, ti = 0;
                try {
                    vi = std::stol(tok.substr(0, s1));
                    if (hasVt) {
                        const size_t s2 = tok.find('/', s1 + 1);
                        ti = std::stol(tok.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1));
                    }
                } catch (...) { g_lerr = "bad face token: " + tok; return RT_ERR_INVALID; }
                if (vi < 1 || vi > (long)verts.size() || (hasVt && (ti < 1 || ti > (long)texCoords.size()))) {
                    g_lerr = "face index out of range: " + tok;
                    return RT_ERR_INVALID;
                }
                p[i] = verts[(size_t)vi - 1];
                if (hasVt) t[i] = texCoords[(size_t)ti - 1];
            }
            auto lib = libToMtlMaps.find(currentLib);
            if (lib == libToMtlMaps.end() || lib->second.find(currentMtl) == lib->second.end()) {
                g_lerr = "Requested material or library not found: " + currentLib + ", " + currentMtl;
                return RT_ERR_INVALID;
            }
            rt_triangle tr;
            memset(&tr, 0, sizeof tr);
            tr.a[0] = p[0].x; tr.a[1] = p[0].y; tr.a[2] = p[0].z;
            tr.b[0] = p[1].x; tr.b[1] = p[1].y; tr.b[2] = p[1].z;
            tr.c[0] = p[2].x; tr.c[1] = p[2].y; tr.c[2] = p[2].z;
            tr.aTex[0] = t[1].x; tr.aTex[1] = t[1].y;  // (vt1, vt2, vt0), M:602-606
            tr.bTex[0] = t[2].x; tr.bTex[1] = t[2].y;
            tr.cTex[0] = t[0].x; tr.cTex[1] = t[0].y;
            tr.materialIndex = lib->second.at(currentMtl).index;
            tris.push_back(tr);
        }
    }
    if (rth_scene_replace_materials(s, materials.data(), (int32_t)materials.size()) != RT_OK) { g_lerr = rth_last_error(); return RT_ERR_INVALID; }
    rth_add_triangles(s, tris.data(), (int64_t)tris.size());
    return RT_OK;
}

extern "C" {

int rth_load_model_folder(rth_scene* s, const char* folder) {
    const int rc = loadFolderImpl(s, folder);
    if (rc != RT_OK) rth_set_error(g_lerr.c_str());
    return rc;
}

}  // extern "C"
