// rt_host.cpp — host-side scene assembly, camera and output (pure C++17; no CUDA, no GL).
//
// Same operations and arithmetic as the reference's host code around the hot path (citations are
// relative to the reference tree; R = RayTracing/src/rayTracing.cpp, C = .../headers/camera.h,
// M = .../headers/mesh.h), rebuilt table-driven: every container is "8 box corners + an index
// table + optional light quads".  Compile with -ffp-contract=off so that vertex positions equal the
// reference's bit for bit (checked against the real reference code in tests/test_host_vs_ref.py).
#include "rt_host.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

struct V3 {
    float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 x, V3 y) {
    return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}
inline V3 normalize(V3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline void put3(float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
inline V3 get3(const float* p) { return {p[0], p[1], p[2]}; }

struct Texture {
    int w = 0, h = 0, ch = 0;
    std::vector<uint8_t> px;
};

rt_material blankMaterial() {  // Material() of M:47 on zeroed storage
    rt_material m;
    memset(&m, 0, sizeof m);
    m.color[0] = m.color[1] = m.color[2] = 1.0f;
    m.textureIndex = -1;
    m.materialType = RT_MAT_DIFFUSE;
    return m;
}

rt_triangle makeTri(int mat, V3 a, V3 b, V3 c) {
    rt_triangle t;
    memset(&t, 0, sizeof t);
    put3(t.a, a);
    put3(t.b, b);
    put3(t.c, c);
    t.materialIndex = mat;
    return t;
}

// 4x4 column-major matrix with glm 0.9.9.7 operation order (type_mat4x4.inl:536-648,
// ext/matrix_transform.inl:18-46)
struct V4 {
    float x, y, z, w;
};
inline V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
struct M4 {
    V4 c[4];
};
M4 identity() { return M4{{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}}; }
M4 rotate(const M4& m, float angle, V3 v) {
    const float c = std::cos(angle), s = std::sin(angle);
    const V3 axis = normalize(v);
    const V3 temp = axis * (1.0f - c);
    float R[3][3];
    R[0][0] = c + temp.x * axis.x;
    R[0][1] = temp.x * axis.y + s * axis.z;
    R[0][2] = temp.x * axis.z - s * axis.y;
    R[1][0] = temp.y * axis.x - s * axis.z;
    R[1][1] = c + temp.y * axis.y;
    R[1][2] = temp.y * axis.z + s * axis.x;
    R[2][0] = temp.z * axis.x + s * axis.y;
    R[2][1] = temp.z * axis.y - s * axis.x;
    R[2][2] = c + temp.z * axis.z;
    M4 r;
    for (int i = 0; i < 3; i++) r.c[i] = (m.c[0] * R[i][0] + m.c[1] * R[i][1]) + m.c[2] * R[i][2];
    r.c[3] = m.c[3];
    return r;
}
M4 mul(const M4& a, const M4& b) {
    M4 r;
    for (int i = 0; i < 4; i++) {
        const V4 bc = b.c[i];
        r.c[i] = ((a.c[0] * bc.x + a.c[1] * bc.y) + a.c[2] * bc.z) + a.c[3] * bc.w;
    }
    return r;
}
V4 mulv(const M4& m, V4 v) {
    const V4 add0 = m.c[0] * v.x + m.c[1] * v.y;
    const V4 add1 = m.c[2] * v.z + m.c[3] * v.w;
    return add0 + add1;
}

}  // namespace

struct rth_scene {
    std::vector<rt_triangle> tris;
    std::vector<rt_material> mats;
    Texture tex[RT_MAX_TEXTURES];
};

namespace {

struct Bounds {
    V3 mn{1e30f, 1e30f, 1e30f}, mx{-1e30f, -1e30f, -1e30f};
};
Bounds sceneBounds(const rth_scene& s) {  // BoundingBox::growToInclude over all triangles
    Bounds b;
    for (const auto& t : s.tris)
        for (const float* p : {t.a, t.b, t.c}) {
            b.mn.x = std::min(b.mn.x, p[0]); b.mn.y = std::min(b.mn.y, p[1]); b.mn.z = std::min(b.mn.z, p[2]);
            b.mx.x = std::max(b.mx.x, p[0]); b.mx.y = std::max(b.mx.y, p[1]); b.mx.z = std::max(b.mx.z, p[2]);
        }
    return b;
}

// Box corner i: bit0 → max X, bit1 → max Y, bit2 → MIN Z (R:467-477 ordering)
struct Box {
    float minX, maxX, minY, maxY, minZ, maxZ;
    V3 corner(int i) const {
        return {(i & 1) ? maxX : minX, (i & 2) ? maxY : minY, (i & 4) ? minZ : maxZ};
    }
};
// Padded enclosure shared by addCornellBox / addMirrorCornellBox / addSideLitCornellBox
// (R:455-465).  BoundingBox::size() returns the X extent in all three components (BVH.h:30-33),
// so every pad is a multiple of the X extent; `extentX` is handed back for the light sizes.
Box paddedBox(const rth_scene& s, float pad, float& extentX) {
    const Bounds b = sceneBounds(s);
    const float sx = b.mx.x - b.mn.x;
    extentX = sx;
    Box box;
    box.minX = b.mn.x - sx * pad;
    box.maxX = b.mx.x + sx * pad;
    box.minY = b.mn.y - sx * pad * 0.1f;
    box.maxY = b.mx.y + sx * pad;
    box.minZ = b.mn.z - sx * pad;
    box.maxZ = b.mx.z + sx * pad;
    return box;
}

void addIndexed(rth_scene& s, const V3* corners, const int (*idx)[3], int n, int mat) {
    for (int i = 0; i < n; i++)
        s.tris.push_back(makeTri(mat, corners[idx[i][0]], corners[idx[i][1]], corners[idx[i][2]]));
}
void addBoxWalls(rth_scene& s, const Box& box, const int (*idx)[3], const int* mats, int uniformMat) {
    V3 c[8];
    for (int i = 0; i < 8; i++) c[i] = box.corner(i);
    for (int i = 0; i < 12; i++)
        s.tris.push_back(makeTri(mats ? mats[i] : uniformMat, c[idx[i][0]], c[idx[i][1]], c[idx[i][2]]));
}

// wall index tables
const int kWallsCornell[12][3] = {{0, 3, 1}, {0, 2, 3}, {0, 5, 4}, {0, 1, 5}, {0, 6, 2}, {0, 4, 6},
                                  {7, 1, 3}, {7, 5, 1}, {7, 2, 6}, {7, 3, 2}, {7, 4, 5}, {7, 6, 4}};  // R:496-510
const int kWallsInward[12][3] = {{0, 3, 1}, {0, 2, 3}, {0, 5, 4}, {0, 1, 5}, {0, 6, 2}, {0, 4, 6},
                                 {1, 7, 5}, {1, 3, 7}, {2, 7, 3}, {2, 6, 7}, {4, 7, 6}, {4, 5, 7}};  // R:611-625, :780-794

}  // namespace

extern "C" {

const char* rth_last_error(void) { return g_err.c_str(); }
void rth_set_error(const char* msg) { g_err = msg ? msg : ""; }

rth_scene* rth_scene_create(void) {
    rth_scene* s = new rth_scene;
    rt_material def = blankMaterial();
    def.index = 0;
    s->mats.push_back(def);
    return s;
}
void rth_scene_destroy(rth_scene* s) { delete s; }
int64_t rth_scene_triangle_count(const rth_scene* s) { return (int64_t)s->tris.size(); }
const rt_triangle* rth_scene_triangles(const rth_scene* s) { return s->tris.data(); }
int32_t rth_scene_material_count(const rth_scene* s) { return (int32_t)s->mats.size(); }
const rt_material* rth_scene_materials(const rth_scene* s) { return s->mats.data(); }
int32_t rth_scene_texture_count(const rth_scene* s) {
    int n = 0;
    while (n < RT_MAX_TEXTURES && s->tex[n].w > 0) n++;
    return n;
}
const uint8_t* rth_scene_texture(const rth_scene* s, int32_t slot, int32_t* w, int32_t* h, int32_t* ch) {
    if (slot < 0 || slot >= RT_MAX_TEXTURES || s->tex[slot].w <= 0) return nullptr;
    if (w) *w = s->tex[slot].w;
    if (h) *h = s->tex[slot].h;
    if (ch) *ch = s->tex[slot].ch;
    return s->tex[slot].px.data();
}

int rth_scene_replace_materials(rth_scene* s, const rt_material* mats, int32_t count) {
    if (!s || !mats || count <= 0) { g_err = "rth_scene_replace_materials: bad argument"; return RT_ERR_INVALID; }
    s->mats.assign(mats, mats + count);
    return RT_OK;
}
int32_t rth_add_material(rth_scene* s, const rt_material* m) {
    s->mats.push_back(*m);
    return (int32_t)s->mats.size() - 1;
}
int32_t rth_add_diffuse(rth_scene* s, float r, float g, float b) {  // M:49-53
    rt_material m = blankMaterial();
    m.materialType = RT_MAT_DIFFUSE;
    m.color[0] = r; m.color[1] = g; m.color[2] = b; m.color[3] = 0.0f;
    return rth_add_material(s, &m);
}
int32_t rth_add_light(rth_scene* s, float r, float g, float b, float strength) {  // M:72-77
    rt_material m = blankMaterial();
    m.materialType = RT_MAT_LIGHT;
    m.emissionColor[0] = r; m.emissionColor[1] = g; m.emissionColor[2] = b;
    m.emissionStrength = strength;
    return rth_add_material(s, &m);
}
int32_t rth_add_specular(rth_scene* s, float r, float g, float b, float sr, float sg, float sb,
                         float smoothness, float prob) {  // M:63-70
    rt_material m = blankMaterial();
    m.materialType = RT_MAT_SPECULAR;
    m.color[0] = r; m.color[1] = g; m.color[2] = b; m.color[3] = 0.0f;
    m.specularColor[0] = sr; m.specularColor[1] = sg; m.specularColor[2] = sb;
    m.smoothness = smoothness;
    m.specularProbability = prob;
    return rth_add_material(s, &m);
}
int32_t rth_add_checker(rth_scene* s, float scale) {  // M:79-83
    rt_material m = blankMaterial();
    m.materialType = RT_MAT_CHECKER;
    m.checkerScale = scale;
    return rth_add_material(s, &m);
}
int32_t rth_add_glass(rth_scene* s, float r, float g, float b, float ri) {  // M:85-90
    rt_material m = blankMaterial();
    m.materialType = RT_MAT_GLASS;
    m.color[0] = r; m.color[1] = g; m.color[2] = b; m.color[3] = 0.0f;
    m.refractiveIndex = ri;
    return rth_add_material(s, &m);
}
int32_t rth_add_textured(rth_scene* s, int32_t texIndex) {  // M:98-102
    rt_material m = blankMaterial();
    m.materialType = RT_MAT_TEXTURE;
    m.textureIndex = texIndex;
    return rth_add_material(s, &m);
}
int32_t rth_add_fixed_materials(rth_scene* s) {  // R:1268-1283
    const int32_t red = rth_add_diffuse(s, 1.0f, 0.0f, 0.0f);
    rth_add_diffuse(s, 0.0f, 1.0f, 0.0f);
    rth_add_diffuse(s, 1.0f, 1.0f, 1.0f);
    rth_add_light(s, 1.0f, 1.0f, 1.0f, 15.0f);
    rth_add_specular(s, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f);
    return red;
}

int rth_add_triangles(rth_scene* s, const rt_triangle* tris, int64_t n) {
    s->tris.insert(s->tris.end(), tris, tris + n);
    return RT_OK;
}

int rth_add_cube(rth_scene* s, const float center[3], const float size[3], const float rotation[3],
                 int32_t material) {  // R:867-923
    const M4 rotX = rotate(identity(), rotation[0], {1, 0, 0});
    const M4 rotY = rotate(identity(), rotation[1], {0, 1, 0});
    const M4 rotZ = rotate(identity(), rotation[2], {0, 0, 1});
    const M4 rot = mul(mul(rotZ, rotY), rotX);
    const V3 half = get3(size) * 0.5f;
    V3 v[8];
    for (int i = 0; i < 8; i++) {  // same corner ordering as Box::corner
        const V3 local = {(i & 1) ? +half.x : -half.x, (i & 2) ? +half.y : -half.y,
                          (i & 4) ? -half.z : +half.z};
        const V4 r = mulv(rot, {local.x, local.y, local.z, 1.0f});
        v[i] = V3{r.x, r.y, r.z} + get3(center);
    }
    static const int faces[12][3] = {{0, 1, 3}, {0, 3, 2}, {1, 5, 7}, {1, 7, 3}, {5, 4, 6}, {5, 6, 7},
                                     {4, 0, 2}, {4, 2, 6}, {2, 3, 7}, {2, 7, 6}, {4, 5, 1}, {4, 1, 0}};
    addIndexed(*s, v, faces, 12, material);
    return RT_OK;
}

int rth_create_classic_cornell_box(rth_scene* s, float roomSize, int32_t red, int32_t green,
                                   int32_t white, int32_t light) {  // R:949-1041
    const float half = roomSize * 0.5f;
    const Box room{-half, +half, -half, +half, -half, +half};
    static const int walls[12][3] = {{0, 3, 1}, {0, 2, 3}, {4, 7, 6}, {4, 5, 7}, {0, 5, 4}, {0, 1, 5},
                                     {2, 6, 7}, {2, 7, 3}, {0, 4, 6}, {0, 6, 2}, {1, 3, 7}, {1, 7, 5}};
    const int mats[12] = {white, white, white, white, white, white, white, white, red, red, green, green};
    addBoxWalls(*s, room, walls, mats, 0);

    const float lightWidth = roomSize * (130.0f / 555.0f);
    const float lightDepth = roomSize * (105.0f / 555.0f);
    const float lightY = half - 0.001f;
    const V3 lc[4] = {{-lightWidth * 0.5f, lightY, +lightDepth * 0.5f},
                      {+lightWidth * 0.5f, lightY, +lightDepth * 0.5f},
                      {-lightWidth * 0.5f, lightY, -lightDepth * 0.5f},
                      {+lightWidth * 0.5f, lightY, -lightDepth * 0.5f}};
    static const int lidx[2][3] = {{0, 2, 1}, {1, 2, 3}};
    addIndexed(*s, lc, lidx, 2, light);

    const float boxScale = 165.0f / 555.0f;
    const float boxSize = roomSize * boxScale;
    const float deg2rad = 0.01745329251994329576923690768489f;  // glm::radians
    {
        const float c[3] = {half * 0.5f, -half + boxSize * 0.5f, -half * 0.3f};
        const float sz[3] = {boxSize, boxSize, boxSize};
        const float rot[3] = {0.0f, -18.0f * deg2rad, 0.0f};
        rth_add_cube(s, c, sz, rot, white);
    }
    {
        const float tallH = roomSize * (330.0f / 555.0f);
        const float c[3] = {-half * 0.3f, -half + tallH * 0.5f, -half * 0.6f};
        const float sz[3] = {boxSize, tallH, boxSize};
        const float rot[3] = {0.0f, 16.5f * deg2rad, 0.0f};
        rth_add_cube(s, c, sz, rot, white);
    }
    return RT_OK;
}

int rth_create_diverse_cornell_box(rth_scene* s, float roomSize, int32_t red, int32_t green,
                                   int32_t white, int32_t light, int32_t glass, int32_t mirror,
                                   int32_t checker, int32_t metal) {  // R:1071-1118
    rth_create_classic_cornell_box(s, roomSize, red, green, white, light);
    const float half = roomSize * 0.5f;
    struct Spec {
        float pos[3], size[3], rot[3];
        int mat;
    };
    const Spec cubes[] = {
        {{-half * 0.7f, -half + 0.1f, half * 0.6f}, {0.15f, 0.15f, 0.15f}, {0.0f, 0.785f, 0.0f}, glass},
        {{half * 0.6f, -half + 0.05f, -half * 0.4f}, {0.08f, 0.08f, 0.08f}, {0.2f, 0.5f, 0.3f}, mirror},
        {{0.0f, -half + 0.2f, -half * 0.7f}, {0.25f, 0.4f, 0.25f}, {0.0f, 0.0f, 0.1f}, checker},
        {{half * 0.3f, -half + 0.3f, half * 0.2f}, {0.1f, 0.6f, 0.1f}, {0.1f, 1.2f, 0.0f}, metal},
        {{-half * 0.2f, -half + 0.15f, -half * 0.2f}, {0.2f, 0.1f, 0.3f}, {0.5f, 0.0f, 0.2f}, glass},
        {{half * 0.8f, -half + 0.03f, half * 0.8f}, {0.05f, 0.05f, 0.05f}, {0.0f, 0.0f, 0.0f}, mirror},
        {{half * 0.75f, -half + 0.08f, half * 0.75f}, {0.06f, 0.06f, 0.06f}, {0.3f, 0.3f, 0.3f}, mirror},
        {{-half * 0.5f, -half + 0.02f, -half * 0.6f}, {0.3f, 0.04f, 0.3f}, {0.0f, 0.7f, 0.0f}, checker},
        {{half * 0.1f, -half + 0.25f, half * 0.5f}, {0.18f, 0.18f, 0.18f}, {0.6f, 0.4f, 0.8f}, metal},
    };
    for (const Spec& c : cubes) rth_add_cube(s, c.pos, c.size, c.rot, c.mat);
    return RT_OK;
}

int rth_add_cornell_box(rth_scene* s, float lightSize, float pad, int32_t light, int32_t lightEnabled) {
    float sx;  // R:453-547
    const Box box = paddedBox(*s, pad, sx);
    const float bsx = box.maxX - box.minX, bsz = box.maxZ - box.minZ;
    const float cx = (box.maxX + box.minX) / 2.0f, cz = (box.maxZ + box.minZ) / 2.0f;
    const float lMinX = cx - lightSize * bsx / 2.0f, lMaxX = cx + lightSize * bsx / 2.0f;
    const float lMinZ = cz - lightSize * bsz / 2.0f, lMaxZ = cz + lightSize * bsz / 2.0f;
    const float lY = box.maxY - 1e-3f;
    addBoxWalls(*s, box, kWallsCornell, nullptr, 0);
    if (lightEnabled) {
        const V3 lc[4] = {{lMinX, lY, lMaxZ}, {lMaxX, lY, lMaxZ}, {lMinX, lY, lMinZ}, {lMaxX, lY, lMinZ}};
        static const int lidx[4][3] = {{0, 3, 1}, {0, 2, 3}, {0, 1, 3}, {0, 3, 2}};  // both windings
        addIndexed(*s, lc, lidx, 4, light);
    }
    return RT_OK;
}

int rth_add_mirror_cornell_box(rth_scene* s, float lightSize, float pad, int32_t light, int32_t mirror) {
    float sx;  // R:569-664
    const Box box = paddedBox(*s, pad, sx);
    const float cx = (box.maxX + box.minX) / 2.0f, cz = (box.maxZ + box.minZ) / 2.0f;
    const float hs = lightSize * sx / 2.0f;
    const float y = box.maxY - 1e-3f;
    addBoxWalls(*s, box, kWallsInward, nullptr, mirror);
    const V3 lc[4] = {{cx - hs, y, cz + hs}, {cx + hs, y, cz + hs}, {cx - hs, y, cz - hs}, {cx + hs, y, cz - hs}};
    static const int lidx[2][3] = {{0, 1, 2}, {1, 3, 2}};
    addIndexed(*s, lc, lidx, 2, light);
    return RT_OK;
}

int rth_add_side_lit_cornell_box(rth_scene* s, float lightSize, float pad, int32_t light, int32_t wall,
                                 int32_t rotate) {  // R:690-847
    float sx;
    const Box box = paddedBox(*s, pad, sx);
    const float cy = (box.maxY + box.minY) / 2.0f;
    const float hs = lightSize * sx / 2.0f;
    const float off = 1e-3f;
    V3 first[4], second[4];
    if (!rotate) {
        const float cz = (box.maxZ + box.minZ) / 2.0f;
        const float xl = box.minX + off, xr = box.maxX - off;
        const V3 a[4] = {{xl, cy - hs, cz + hs}, {xl, cy + hs, cz + hs}, {xl, cy - hs, cz - hs}, {xl, cy + hs, cz - hs}};
        const V3 b[4] = {{xr, cy - hs, cz + hs}, {xr, cy + hs, cz + hs}, {xr, cy - hs, cz - hs}, {xr, cy + hs, cz - hs}};
        std::copy(a, a + 4, first);
        std::copy(b, b + 4, second);
    } else {
        const float cx = (box.maxX + box.minX) / 2.0f;
        const float zf = box.maxZ - off, zb = box.minZ + off;
        const V3 a[4] = {{cx - hs, cy - hs, zf}, {cx + hs, cy - hs, zf}, {cx - hs, cy + hs, zf}, {cx + hs, cy + hs, zf}};
        const V3 b[4] = {{cx - hs, cy - hs, zb}, {cx + hs, cy - hs, zb}, {cx - hs, cy + hs, zb}, {cx + hs, cy + hs, zb}};
        std::copy(a, a + 4, first);
        std::copy(b, b + 4, second);
    }
    static const int idxFirst[2][3] = {{0, 2, 1}, {1, 2, 3}};
    static const int idxSecond[2][3] = {{0, 1, 2}, {1, 3, 2}};
    addBoxWalls(*s, box, kWallsInward, nullptr, wall);
    addIndexed(*s, first, idxFirst, 2, light);
    addIndexed(*s, second, idxSecond, 2, light);
    return RT_OK;
}

int rth_add_sky_light_plane(rth_scene* s, int32_t light) {  // R:388-432
    const Bounds b = sceneBounds(*s);
    const float sx = b.mx.x - b.mn.x;
    const float planeY = b.mx.y + sx * 0.3f;
    const V3 c[4] = {{b.mn.x, planeY, b.mx.z}, {b.mx.x, planeY, b.mx.z}, {b.mn.x, planeY, b.mn.z}, {b.mx.x, planeY, b.mn.z}};
    static const int idx[2][3] = {{0, 3, 1}, {0, 2, 3}};
    addIndexed(*s, c, idx, 2, light);
    addIndexed(*s, c, idx, 2, light);  // inserted twice on purpose (R:428-431)
    return RT_OK;
}

int rth_add_displaced_sphere(rth_scene* s, int32_t n, const float center[3], float radius, float amp,
                             int32_t material) {
    if (n < 2) { g_err = "displaced sphere needs n >= 2"; return RT_ERR_INVALID; }
    const double PI_D = 3.14159265358979323846;
    std::vector<V3> v((size_t)(n + 1) * n);
    for (int i = 0; i <= n; i++) {
        // the first and last rows collapse onto the poles, so the surface is closed; the n triangles
        // per pole that degenerate to a segment have zero area and can never be hit (2n² in total)
        const double theta = PI_D * (double)i / (double)n;
        const double sinTheta = (i == 0 || i == n) ? 0.0 : std::sin(theta);
        for (int j = 0; j < n; j++) {
            const double phi = 2.0 * PI_D * (double)j / (double)n;
            const double r = (double)radius * (1.0 + (double)amp * std::sin(8.0 * theta) * std::sin(6.0 * phi));
            v[(size_t)i * n + j] = V3{(float)((double)center[0] + r * sinTheta * std::cos(phi)),
                                      (float)((double)center[1] + r * std::cos(theta)),
                                      (float)((double)center[2] + r * sinTheta * std::sin(phi))};
        }
    }
    s->tris.reserve(s->tris.size() + (size_t)2 * n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const int j1 = (j + 1) % n;
            const V3 p00 = v[(size_t)i * n + j], p01 = v[(size_t)i * n + j1];
            const V3 p10 = v[(size_t)(i + 1) * n + j], p11 = v[(size_t)(i + 1) * n + j1];
            const float u0 = (float)j / (float)n, u1 = (float)(j + 1) / (float)n;
            const float v0 = (float)i / (float)n, v1 = (float)(i + 1) / (float)n;
            // outward normal: (d/dphi) x (d/dtheta) points away from the centre
            rt_triangle t0 = makeTri(material, p00, p01, p11);
            rt_triangle t1 = makeTri(material, p00, p11, p10);
            // (aTex,bTex,cTex) hold the uv of (b, c, a) — the loader's rotation, mesh.h:602-606
            t0.aTex[0] = u1; t0.aTex[1] = v0; t0.bTex[0] = u1; t0.bTex[1] = v1; t0.cTex[0] = u0; t0.cTex[1] = v0;
            t1.aTex[0] = u1; t1.aTex[1] = v1; t1.bTex[0] = u0; t1.bTex[1] = v1; t1.cTex[0] = u0; t1.cTex[1] = v0;
            s->tris.push_back(t0);
            s->tris.push_back(t1);
        }
    return RT_OK;
}

int rth_set_texture(rth_scene* s, int32_t slot, const uint8_t* px, int32_t w, int32_t h, int32_t ch) {
    if (!s || slot < 0 || slot >= RT_MAX_TEXTURES || !px || w <= 0 || h <= 0 || ch < 1 || ch > 4) {
        g_err = "bad texture";
        return RT_ERR_INVALID;
    }
    Texture& t = s->tex[slot];
    t.w = w; t.h = h; t.ch = ch;
    t.px.assign(px, px + (size_t)w * h * ch);
    return RT_OK;
}

int rth_set_procedural_texture(rth_scene* s, int32_t slot, int32_t size) {
    if (size <= 1) { g_err = "bad size"; return RT_ERR_INVALID; }
    std::vector<uint8_t> px((size_t)size * size * 3);
    const int cell = std::max(1, size / 16);
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) {
            uint8_t* p = &px[((size_t)y * size + x) * 3];
            const bool on = (((x / cell) + (y / cell)) & 1) != 0;
            p[0] = (uint8_t)(60 + (x * 180) / (size - 1));
            p[1] = (uint8_t)(60 + (y * 180) / (size - 1));
            p[2] = on ? 230 : 70;
        }
    return rth_set_texture(s, slot, px.data(), size, size, 3);
}

// ------------------------------------------------------------------------------------------------ camera
static void cameraViewport(rth_camera* c, float mouseYOffset) {  // C:163-180
    c->zoom += mouseYOffset;
    const float h = std::tan(c->hfov / 2);
    const float viewportWidth = 2 * h / std::exp(c->zoom * c->zoomSensitivity);
    const float viewportHeight = viewportWidth / c->aspectRatio;
    const V3 right = get3(c->right), up = get3(c->up), front = get3(c->front);
    const V3 vr = (right * viewportWidth) * c->focusDistance;
    const V3 vu = (up * viewportHeight) * c->focusDistance;
    const V3 vf = (-front) * c->focusDistance;
    put3(c->viewportRight, vr);
    put3(c->viewportUp, vu);
    put3(c->viewportFront, vf);
    put3(c->pixelRight, vr / c->scrWidth);
    put3(c->pixelUp, vu / c->scrHeight);
    const float defocusRadius = c->focusDistance * std::tan(c->defocusAngle / 2.0f);
    put3(c->defocusDiskRight, right * defocusRadius);
    put3(c->defocusDiskUp, up * defocusRadius);
}
static void cameraBasis(rth_camera* c) {  // C:150-160
    // the unqualified cos/sin/exp of camera.h resolve to the float overloads (checked bit for bit
    // against the compiled reference in tests/test_oracle_cpu.py)
    V3 front;
    front.x = std::cos(c->yaw) * std::cos(c->pitch);
    front.y = std::sin(c->pitch);
    front.z = std::sin(c->yaw) * std::cos(c->pitch);
    front = normalize(front);
    const V3 right = normalize(cross(get3(c->worldUp), front));
    const V3 up = normalize(cross(front, right));
    put3(c->front, front);
    put3(c->right, right);
    put3(c->up, up);
    cameraViewport(c, 0.0f);
}
static inline float clampf(float x, float lo, float hi) { return std::max(std::min(x, hi), lo); }
static const float kPI = (float)(atan(1.0) * 4.0f);  // math_util.h:8

void rth_camera_init(rth_camera* c, int32_t width, int32_t height, float speed, const float pos[3],
                     float hfov, float pitch, float yaw, float focusDist, float defocusAngle, float zoom) {
    memset(c, 0, sizeof *c);
    c->scrWidth = (float)width;
    c->scrHeight = (float)height;
    c->aspectRatio = (float)width / height;
    c->speed = speed;
    put3(c->position, get3(pos));
    c->hfov = hfov; c->pitch = pitch; c->yaw = yaw;
    c->focusDistance = focusDist; c->defocusAngle = defocusAngle; c->zoom = zoom;
    c->worldUp[1] = 1.0f;
    // the reference ctor calls updateBasisVectors() before the sensitivities are assigned
    // (C:106-113), so the first viewport uses an indeterminate zoomSensitivity and is then
    // recomputed with 0.1; only the second result is observable.
    c->mouseSensitivity = 1.0f;
    c->zoomSensitivity = 0.1f;
    c->defocusSensitivity = 0.1f;
    cameraBasis(c);
    c->lastX = 0.0; c->lastY = 0.0;
    c->firstMouse = 1;
}
void rth_camera_keyboard(rth_camera* c, uint8_t bits, float dt) {  // C:121-147
    V3 pos = get3(c->position);
    const V3 front = get3(c->front), right = get3(c->right), worldUp = get3(c->worldUp);
    const V3 flat = normalize(V3{front.x, 0.0f, front.z});
    if (bits & RTH_FORWARD) pos = pos - (flat * c->speed) * dt;
    if (bits & RTH_BACKWARD) pos = pos + (flat * c->speed) * dt;
    if (bits & RTH_LEFT) pos = pos - (right * c->speed) * dt;
    if (bits & RTH_RIGHT) pos = pos + (right * c->speed) * dt;
    if (bits & RTH_UP) pos = pos + (worldUp * c->speed) * dt;
    if (bits & RTH_DOWN) pos = pos - (worldUp * c->speed) * dt;
    put3(c->position, pos);
    if (bits & RTH_DEFOCUS_UP) {
        c->defocusAngle += c->defocusSensitivity * dt;
        c->defocusAngle = clampf(c->defocusAngle, 0.0f, kPI / 2.0f);
        cameraViewport(c, 0.0f);
    }
    if (bits & RTH_DEFOCUS_DOWN) {
        c->defocusAngle -= c->defocusSensitivity * dt;
        c->defocusAngle = clampf(c->defocusAngle, 0.0f, kPI / 2.0f);
        cameraViewport(c, 0.0f);
    }
}
void rth_camera_mouse(rth_camera* c, double xpos, double ypos) {  // C:194-213
    if (c->firstMouse) {
        c->lastX = xpos;
        c->lastY = ypos;
        c->firstMouse = 0;
    }
    const float xoffset = (float)(xpos - c->lastX);
    const float yoffset = (float)(ypos - c->lastY);
    c->lastX = xpos;
    c->lastY = ypos;
    const float ez = std::exp(c->zoom * c->zoomSensitivity);
    c->pitch += yoffset / c->scrWidth * c->mouseSensitivity / ez;
    c->pitch = clampf(c->pitch, -kPI / 2.1f, kPI / 2.1f);
    c->yaw += xoffset / c->scrWidth * c->mouseSensitivity / ez;
    cameraBasis(c);
}
void rth_camera_scroll(rth_camera* c, float yOffset) { cameraViewport(c, yOffset); }
void rth_camera_update_uniforms(const rth_camera* c, rt_uniforms* u) {  // C:182-192
    auto put4 = [](float* d, const float* s3) { d[0] = s3[0]; d[1] = s3[1]; d[2] = s3[2]; d[3] = 0.0f; };
    put4(u->cameraPos, c->position);
    put4(u->viewportRight, c->viewportRight);
    put4(u->viewportUp, c->viewportUp);
    put4(u->viewportFront, c->viewportFront);
    put4(u->pixelRight, c->pixelRight);
    put4(u->pixelUp, c->pixelUp);
    put4(u->defocusDiskRight, c->defocusDiskRight);
    put4(u->defocusDiskUp, c->defocusDiskUp);
}

void rth_get_defaults(rth_defaults* d) {  // R:48-89
    memset(d, 0, sizeof *d);
    d->scr_width = 1000; d->scr_height = 1000;
    d->max_bounce_count = 10;
    d->num_rays_per_pixel = 5;
    d->rays_per_pixel_sensitivity = 10.0f;
    d->basic_shading = 1; d->basic_shading_shadow = 0; d->basic_shading_environmental_light = 0;
    d->light_position[0] = 10.0f; d->light_position[1] = 100.0f; d->light_position[2] = 1.0f;
    d->screenshot_basic_shading = 0; d->screenshot_environmental_light = 1;
    d->screenshot_max_bounce_count = 20; d->screenshot_rays_per_pixel = 64; d->screenshot_frames = 10;
    d->cornell_light_brightness = 15.0f; d->cornell_padding = 0.3f; d->cornell_light_size = 0.17f;
    d->max_speed = 10.0f;
    d->hfov = kPI / 6; d->pitch = 0.0f; d->yaw = kPI / 2.0f;
    d->focus_distance = 20.0f; d->defocus_angle = 0.0f; d->zoom = 1.0f;
    d->camera_pos[0] = 0.0f; d->camera_pos[1] = 5.0f; d->camera_pos[2] = 10.0f;
}

void rth_fill_interactive_uniforms(const rth_scene* s, const rth_camera* c, float numRaysPerPixel,
                                   uint32_t frameIndex, rt_uniforms* u) {  // R:1386-1400
    rth_defaults d;
    rth_get_defaults(&d);
    memset(u, 0, sizeof *u);
    u->numTextures = rth_scene_texture_count(s);
    u->width = (uint32_t)c->scrWidth;
    u->height = (uint32_t)c->scrHeight;
    u->numSpheres = 0;
    u->numTriangles = (int32_t)s->tris.size();
    u->basicShading = d.basic_shading;
    u->basicShadingShadow = d.basic_shading_shadow;
    u->basicShadingLightPosition[0] = d.light_position[0];
    u->basicShadingLightPosition[1] = d.light_position[1];
    u->basicShadingLightPosition[2] = d.light_position[2];
    u->environmentalLight = d.basic_shading_environmental_light;
    u->maxBounceCount = d.max_bounce_count;
    u->numRaysPerPixel = (int32_t)numRaysPerPixel;  // float → int truncation, R:1396
    u->frameIndex = frameIndex;
    rth_camera_update_uniforms(c, u);
}
void rth_fill_screenshot_uniforms(const rth_scene* s, const rth_camera* c, rt_uniforms* u) {
    rth_defaults d;  // R:146-152 on top of the last interactive block
    rth_get_defaults(&d);
    rth_fill_interactive_uniforms(s, c, d.num_rays_per_pixel, 0, u);
    u->basicShading = d.screenshot_basic_shading;
    u->environmentalLight = d.screenshot_environmental_light;
    u->maxBounceCount = d.screenshot_max_bounce_count;
    u->numRaysPerPixel = d.screenshot_rays_per_pixel;
    u->frameIndex = 0;
}
float rth_adjust_rays_per_pixel(float current, int32_t increase, float dt) {  // R:358-367
    current += (increase ? 10.0f : -10.0f) * dt;
    return clampf(current, 1.1f, 200.0f);
}

// ------------------------------------------------------------------------------------------------ PNG
static uint32_t crc32buf(uint32_t crc, const uint8_t* p, size_t n) { return (uint32_t)crc32(crc, p, (uInt)n); }
int rth_write_png(const char* path, int32_t w, int32_t h, int32_t channels, const uint8_t* pixels) {
    if (!path || !pixels || w <= 0 || h <= 0 || (channels != 1 && channels != 3 && channels != 4)) {
        g_err = "rth_write_png: bad argument";
        return RT_ERR_INVALID;
    }
    const size_t stride = (size_t)w * channels;
    std::vector<uint8_t> raw((stride + 1) * h);
    for (int y = 0; y < h; y++) {
        raw[(stride + 1) * y] = 0;  // filter: none
        memcpy(&raw[(stride + 1) * y + 1], pixels + stride * y, stride);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) {
        g_err = "rth_write_png: deflate failed";
        return RT_ERR_INVALID;
    }
    FILE* f = fopen(path, "wb");
    if (!f) { g_err = std::string("rth_write_png: cannot open ") + path; return RT_ERR_INVALID; }
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    bool wrote = fwrite(sig, 1, 8, f) == 8;
    auto chunk = [&](const char* type, const uint8_t* data, uint32_t len) {
        uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                          (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
        wrote = wrote && fwrite(hdr, 1, 8, f) == 8;
        if (len) wrote = wrote && fwrite(data, 1, len, f) == len;
        uint32_t c = crc32buf(0, hdr + 4, 4);
        if (len) c = crc32buf(c, data, len);
        uint8_t cb[4] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
        wrote = wrote && fwrite(cb, 1, 4, f) == 4;
    };
    const uint8_t colorType = channels == 1 ? 0 : (channels == 3 ? 2 : 6);
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                        (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h,
                        8, colorType, 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), (uint32_t)zlen);
    chunk("IEND", nullptr, 0);
    if (fclose(f) != 0 || !wrote) { g_err = std::string("rth_write_png: write failed: ") + path; return RT_ERR_INVALID; }
    return RT_OK;
}

// ------------------------------------------------------------------------------------------------ RTSC
static const int64_t kMagic = 0x43535452;
int rth_scene_save(const rth_scene* s, const char* path) {
    if (!s || !path) { g_err = "rth_scene_save: NULL argument"; return RT_ERR_INVALID; }
    FILE* f = fopen(path, "wb");
    if (!f) { g_err = std::string("cannot open ") + path; return RT_ERR_INVALID; }
    const int ntex = rth_scene_texture_count(s);
    int64_t hdr[8] = {kMagic, (int64_t)s->tris.size(), (int64_t)s->mats.size(), 0, 0, ntex, 0, 0};
    bool ok = fwrite(hdr, 8, 8, f) == 8 &&
              fwrite(s->tris.data(), sizeof(rt_triangle), s->tris.size(), f) == s->tris.size() &&
              fwrite(s->mats.data(), sizeof(rt_material), s->mats.size(), f) == s->mats.size();
    for (int i = 0; ok && i < ntex; i++) {
        int32_t th[4] = {s->tex[i].w, s->tex[i].h, s->tex[i].ch, 0};
        ok = fwrite(th, 4, 4, f) == 4 && fwrite(s->tex[i].px.data(), 1, s->tex[i].px.size(), f) == s->tex[i].px.size();
    }
    if (fclose(f) != 0 || !ok) { g_err = std::string("rth_scene_save: write failed: ") + path; return RT_ERR_INVALID; }
    return RT_OK;
}
int rth_scene_load(rth_scene* s, const char* path) {
    // The header is untrusted: every count is checked against the file size before anything is resized, the file is
    // read into temporaries, and the scene is only touched once the whole file has been accepted.
    if (!s || !path) { g_err = "rth_scene_load: NULL argument"; return RT_ERR_INVALID; }
    FILE* f = fopen(path, "rb");
    if (!f) { g_err = std::string("cannot open ") + path; return RT_ERR_INVALID; }
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
    try {
        if (fseek(f, 0, SEEK_END) != 0) { g_err = "cannot seek in RTSC file"; return RT_ERR_INVALID; }
        const int64_t fileSize = (int64_t)ftell(f);
        rewind(f);
        int64_t hdr[8];
        if (fread(hdr, 8, 8, f) != 8 || hdr[0] != kMagic) { g_err = "not an RTSC file"; return RT_ERR_INVALID; }
        const int64_t nTris = hdr[1], nMats = hdr[2], nNodes = hdr[3], nSorted = hdr[4], nTex = hdr[5];
        const int64_t lim = fileSize / 48 + 1;  // no section can hold more records than the file has bytes for
        if (nTris < 0 || nMats < 0 || nNodes < 0 || nSorted < 0 || nTex < 0 || nTris > lim || nMats > lim || nNodes > lim ||
            nSorted > lim || nTex > RT_MAX_TEXTURES ||
            64 + nTris * (int64_t)sizeof(rt_triangle) + nMats * (int64_t)sizeof(rt_material) + nNodes * 48 + nSorted * 80 > fileSize) {
            g_err = "RTSC header inconsistent with the file size";
            return RT_ERR_INVALID;
        }
        std::vector<rt_triangle> tris((size_t)nTris);
        std::vector<rt_material> mats((size_t)nMats);
        bool ok = fread(tris.data(), sizeof(rt_triangle), tris.size(), f) == tris.size() &&
                  fread(mats.data(), sizeof(rt_material), mats.size(), f) == mats.size();
        ok = ok && fseek(f, (long)(nNodes * 48 + nSorted * 80), SEEK_CUR) == 0;
        struct Tex { int32_t w = 0, h = 0, ch = 0; std::vector<uint8_t> px; } tex[RT_MAX_TEXTURES];
        for (int i = 0; ok && i < nTex; i++) {
            int32_t th[4];
            ok = fread(th, 4, 4, f) == 4;
            if (!ok) break;
            if (th[0] <= 0 || th[1] <= 0 || th[2] < 1 || th[2] > 4 || (int64_t)th[0] * th[1] * th[2] > fileSize) {
                g_err = "RTSC texture header out of range";
                return RT_ERR_INVALID;
            }
            tex[i].w = th[0]; tex[i].h = th[1]; tex[i].ch = th[2];
            tex[i].px.resize((size_t)th[0] * th[1] * th[2]);
            ok = fread(tex[i].px.data(), 1, tex[i].px.size(), f) == tex[i].px.size();
        }
        if (!ok) { g_err = "truncated RTSC file"; return RT_ERR_INVALID; }
        s->tris.swap(tris);
        s->mats.swap(mats);
        for (int i = 0; i < nTex; i++) {
            s->tex[i].w = tex[i].w; s->tex[i].h = tex[i].h; s->tex[i].ch = tex[i].ch;
            s->tex[i].px.swap(tex[i].px);
        }
        return RT_OK;
    } catch (const std::bad_alloc&) {
        g_err = "rth_scene_load: out of memory";
        return RT_ERR_OOM;
    } catch (const std::exception& e) {
        g_err = std::string("rth_scene_load: ") + e.what();
        return RT_ERR_INVALID;
    }
}

}  // extern "C"
