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
// JPEG textures are not decoded yet (round 2) — such folders load through an RTSC file written by
// oracle/_ref/ref_host, which runs the reference's loader and stb.
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

bool loadTexture(const fs::path& p, Image& img, std::string& err) {
    std::ifstream f(p, std::ios::binary);
    if (!f) { err = "cannot open " + p.string(); return false; }
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::string ext = p.extension().string();
    std::transform(ext.begin(), ext.end(), ext.begin(), ::tolower);
    if (ext != ".png") {
        err = "texture " + p.filename().string() + ": only PNG is decoded by the host library (JPEG: load an RTSC file "
              "written by oracle/_ref/ref_host, or decode in the caller and use rth_set_texture)";
        return false;
    }
    if (!decodePng(bytes, img, err)) { err = p.filename().string() + ": " + err; return false; }
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
